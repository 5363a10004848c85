// Furthest point sampling.  Semantics: libs/pointnet_lib/src/sampling_gpu.cu:93-209
// (kernel), :211-253 (launcher), cuda_utils.h:10-14 (opt_n_threads).
//
// The reference runs one block of T = opt_n_threads(n) threads per cloud, re-reads
// xyz and the running min-distance `temp` from global memory in each of the m-1
// rounds and reduces through a 10-level __syncthreads tree.  Here the block is
// persistent over the rounds with its points resident on chip (registers, or
// shared memory for xyz when 16 points per thread would not fit the register
// file), the per-round argmax is two redux.sync per warp plus ONE block barrier,
// and `temp` is written back once at the end.
//
// Tie-break: a thread keeps the lowest k among equal
// maxima in its residue class k = t (mod T); the reference tree then keeps, among
// equal per-thread maxima, the slot with the smallest bit-reversed thread id.
// We use the same T and the same thread->point map and reduce the pair
// (ordered(d2), ~bitrev(t)) with an integer max, which selects the same point.
#include "common.cuh"
#include "../../include/dcl_b200.h"
#include <cooperative_groups.h>
#include <math.h>

namespace cg = cooperative_groups;

namespace {

// cuda_utils.h:10-14 of the reference, restated.
inline int ref_opt_n_threads(int work_size) {
    const int pow_2 = (int)(log((double)work_size) / log(2.0));
    int t = 1 << pow_2;
    if (t > 1024) t = 1024;
    if (t < 1) t = 1;
    return t;
}

__device__ __forceinline__ uint32_t ordered_bits(float f) {
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

struct FpsShared {
    uint32_t u[2][32];
    uint32_t tk[2][32];
    int bi[2][32];
};

// Block-wide argmax of (best, slot tie-break); returns the winning point index to
// every thread.  One __syncthreads per call; `par` alternates the scratch buffer.
__device__ __forceinline__ int fps_block_argmax(FpsShared& sh, int par, float best, int besti, uint32_t tie_rank,
                                                bool active, int nwarps) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t u = active ? ordered_bits(best) : 0u;
    const uint32_t umax = __reduce_max_sync(0xffffffffu, u);
    const uint32_t tk = (active && u == umax) ? tie_rank : 0u;
    const uint32_t tmax = __reduce_max_sync(0xffffffffu, tk);
    if (tk == tmax && tk != 0u) {
        sh.u[par][warp] = umax;
        sh.tk[par][warp] = tmax;
        sh.bi[par][warp] = besti;
    }
    __syncthreads();
    const uint32_t wu = (lane < nwarps) ? sh.u[par][lane] : 0u;
    const uint32_t wt = (lane < nwarps) ? sh.tk[par][lane] : 0u;
    const int wi = (lane < nwarps) ? sh.bi[par][lane] : 0;
    const uint32_t gu = __reduce_max_sync(0xffffffffu, wu);
    const uint32_t wt2 = (lane < nwarps && wu == gu) ? wt : 0u;
    const uint32_t gt = __reduce_max_sync(0xffffffffu, wt2);
    const uint32_t src = __ffs(__ballot_sync(0xffffffffu, wt2 == gt && wt2 != 0u)) - 1;
    return __shfl_sync(0xffffffffu, wi, src);
}

// tie_rank: larger wins.  ~bitrev_L(t) restricted to L bits, +1 so that it is never 0.
__device__ __forceinline__ uint32_t fps_tie_rank(int t, int T) {
    const int L = 31 - __clz(T);
    const uint32_t rev = (L == 0) ? 0u : (__brev((uint32_t)t) >> (32 - L));
    return (uint32_t)T - rev;  // in [1, T]
}

// MODE 0: xyz + temp in registers (n <= T*PPT).  MODE 1: xyz in shared memory (SoA),
// temp in registers.
template <int PPT, int MODE>
__global__ void __launch_bounds__(1024, 1) fps_resident_kernel(int n, int m, int T, const float* __restrict__ dataset,
                                                               float* __restrict__ temp, int* __restrict__ idxs) {
    extern __shared__ __align__(16) float s_xyz[];  // MODE 1: x[n] y[n] z[n]
    __shared__ FpsShared sh;
    if (m <= 0) return;
    const int t = threadIdx.x;
    const bool active = t < T;
    const int nwarps = (blockDim.x + 31) >> 5;
    dataset += (size_t)blockIdx.x * n * 3;
    temp += (size_t)blockIdx.x * n;
    idxs += (size_t)blockIdx.x * m;

    float px[MODE == 0 ? PPT : 1], py[MODE == 0 ? PPT : 1], pz[MODE == 0 ? PPT : 1];
    float td[PPT];
    float* sx = s_xyz;
    float* sy = s_xyz + n;
    float* sz = s_xyz + 2 * n;
    if (MODE == 1) {
        for (int i = t; i < n * 3; i += blockDim.x) {
            const float v = dataset[i];
            const int k = i / 3, c = i - 3 * k;
            s_xyz[c * n + k] = v;
        }
    }
#pragma unroll
    for (int i = 0; i < PPT; ++i) {
        const int k = t + i * T;
        const bool ok = active && k < n;
        td[i] = ok ? temp[k] : 0.f;
        if (MODE == 0) {
            px[i] = ok ? dataset[k * 3 + 0] : 0.f;
            py[i] = ok ? dataset[k * 3 + 1] : 0.f;
            pz[i] = ok ? dataset[k * 3 + 2] : 0.f;
        }
    }
    if (t == 0) idxs[0] = 0;
    const uint32_t tie_rank = fps_tie_rank(t, T);
    __syncthreads();

    int old = 0;
    for (int j = 1; j < m; ++j) {
        float x1, y1, z1;
        if (MODE == 1) {
            x1 = sx[old];
            y1 = sy[old];
            z1 = sz[old];
        } else {
            x1 = __ldg(dataset + old * 3 + 0);
            y1 = __ldg(dataset + old * 3 + 1);
            z1 = __ldg(dataset + old * 3 + 2);
        }
        float best = -1.f;
        int besti = 0;
#pragma unroll
        for (int i = 0; i < PPT; ++i) {
            const int k = t + i * T;
            if (k < n) {
                float x2, y2, z2;
                if (MODE == 1) {
                    x2 = sx[k];
                    y2 = sy[k];
                    z2 = sz[k];
                } else {
                    x2 = px[i];
                    y2 = py[i];
                    z2 = pz[i];
                }
                const float d = dcl_dist2(x2, y2, z2, x1, y1, z1);
                const float d2 = fminf(d, td[i]);
                td[i] = d2;
                if (d2 > best) {
                    best = d2;
                    besti = k;
                }
            }
        }
        old = fps_block_argmax(sh, j & 1, best, besti, tie_rank, active, nwarps);
        if (t == 0) idxs[j] = old;
    }
    if (m > 1) {
#pragma unroll
        for (int i = 0; i < PPT; ++i) {
            const int k = t + i * T;
            if (active && k < n) temp[k] = td[i];
        }
    }
}

// Any n: xyz and temp stay in global memory (L1/L2-resident), as in the reference.
__global__ void __launch_bounds__(1024, 1) fps_streaming_kernel(int n, int m, int T, const float* __restrict__ dataset,
                                                                float* __restrict__ temp, int* __restrict__ idxs) {
    __shared__ FpsShared sh;
    if (m <= 0) return;
    const int t = threadIdx.x;
    const bool active = t < T;
    const int nwarps = (blockDim.x + 31) >> 5;
    dataset += (size_t)blockIdx.x * n * 3;
    temp += (size_t)blockIdx.x * n;
    idxs += (size_t)blockIdx.x * m;
    if (t == 0) idxs[0] = 0;
    const uint32_t tie_rank = fps_tie_rank(t, T);
    int old = 0;
    for (int j = 1; j < m; ++j) {
        const float x1 = dataset[old * 3 + 0], y1 = dataset[old * 3 + 1], z1 = dataset[old * 3 + 2];
        float best = -1.f;
        int besti = 0;
        if (active) {
            for (int k = t; k < n; k += T) {
                const float d = dcl_dist2(dataset[k * 3 + 0], dataset[k * 3 + 1], dataset[k * 3 + 2], x1, y1, z1);
                const float d2 = fminf(d, temp[k]);
                temp[k] = d2;
                if (d2 > best) {
                    best = d2;
                    besti = k;
                }
            }
        }
        old = fps_block_argmax(sh, j & 1, best, besti, tie_rank, active, nwarps);
        if (t == 0) idxs[j] = old;
    }
}

// ---------------------------------------------------------------------------------------------
// Cluster version (n >= 1024, i.e. the reference's T = 1024): CS CTAs of 1024 threads share one cloud, so
// CS times as many SMs work on the serial chain of m-1 rounds.  Thread `tid` of CTA `rank` owns the points
// k = tid + 1024*(rank + CS*i): every k it owns is in the reference thread's residue class (k mod 1024 = tid)
// and ascending in i, so the per-thread "first maximum" is the reference's.  Points and running distances
// live in registers.  Per round: sweep -> warp argmax (2 redux) -> one __syncthreads -> CTA argmax ->
// the CTA winner (key + coordinates, 20 B) is written into every CTA's mailbox through distributed shared
// memory -> cluster barrier -> every thread picks the cluster winner.  Key = (ordered(d2), T - bitrev(tid),
// ~k): maximal d2, then the reference tree's slot preference, then — for equal slots in different CTAs —
// the lowest k.
struct FpsMail {
    uint32_t u, lo;
    float x, y, z;
    uint32_t pad[3];
};

template <int PPT, int CS>
__global__ void __cluster_dims__(CS, 1, 1) __launch_bounds__(1024, 1)
    fps_cluster_kernel(int n, int m, const float* __restrict__ dataset, float* __restrict__ temp,
                       int* __restrict__ idxs) {
    __shared__ FpsMail s_warp[2][32];
    __shared__ FpsMail s_mail[2][CS];
    if (m <= 0) return;
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int cloud = blockIdx.x / CS;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    dataset += (size_t)cloud * n * 3;
    temp += (size_t)cloud * n;
    idxs += (size_t)cloud * m;

    float px[PPT], py[PPT], pz[PPT], td[PPT];
#pragma unroll
    for (int i = 0; i < PPT; ++i) {
        const int k = tid + 1024 * (rank + CS * i);
        const bool ok = k < n;
        px[i] = ok ? dataset[k * 3 + 0] : 0.f;
        py[i] = ok ? dataset[k * 3 + 1] : 0.f;
        pz[i] = ok ? dataset[k * 3 + 2] : 0.f;
        td[i] = ok ? temp[k] : -1.f;  // a padded slot can never beat best = -1 (strict '>')
    }
    const uint32_t tie_rank = fps_tie_rank(tid, 1024);  // [1, 1024]
    if (rank == 0 && tid == 0) idxs[0] = 0;
    float x1 = dataset[0], y1 = dataset[1], z1 = dataset[2];

    for (int j = 1; j < m; ++j) {
        const int par = j & 1;
        float best = -1.f, bx = 0.f, by = 0.f, bz = 0.f;
        int bestk = 0;
#pragma unroll
        for (int i = 0; i < PPT; ++i) {
            const float d = dcl_dist2(px[i], py[i], pz[i], x1, y1, z1);
            const float d2 = fminf(d, td[i]);
            td[i] = d2;
            if (d2 > best) {
                best = d2;
                bestk = tid + 1024 * (rank + CS * i);
                bx = px[i];
                by = py[i];
                bz = pz[i];
            }
        }
        // warp level
        const uint32_t u = ordered_bits(best);
        const uint32_t umax = __reduce_max_sync(0xffffffffu, u);
        const uint32_t lo = (u == umax) ? ((tie_rank << 21) | (0x1FFFFFu - (uint32_t)bestk)) : 0u;
        const uint32_t lomax = __reduce_max_sync(0xffffffffu, lo);
        if (lo == lomax && lo != 0u) {
            FpsMail& e = s_warp[par][warp];
            e.u = umax;
            e.lo = lomax;
            e.x = bx;
            e.y = by;
            e.z = bz;
        }
        __syncthreads();
        // CTA level (every warp redundantly; only warp 0 publishes)
        const FpsMail we = s_warp[par][lane];
        const uint32_t gu = __reduce_max_sync(0xffffffffu, we.u);
        const uint32_t wl = (we.u == gu) ? we.lo : 0u;
        const uint32_t gl = __reduce_max_sync(0xffffffffu, wl);
        if (warp == 0) {
            const int src = __ffs(__ballot_sync(0xffffffffu, wl == gl && wl != 0u)) - 1;
            const float wx = __shfl_sync(0xffffffffu, we.x, src), wy = __shfl_sync(0xffffffffu, we.y, src),
                        wz = __shfl_sync(0xffffffffu, we.z, src);
            if (lane < CS) {
                FpsMail* remote = cluster.map_shared_rank(&s_mail[par][rank], lane);
                remote->u = gu;
                remote->lo = gl;
                remote->x = wx;
                remote->y = wy;
                remote->z = wz;
            }
        }
        cluster.sync();
        // cluster level: CS entries, identical in every CTA
        uint32_t bu = 0u, bl = 0u;
#pragma unroll
        for (int r = 0; r < CS; ++r) {
            const FpsMail e = s_mail[par][r];
            if (e.u > bu || (e.u == bu && e.lo > bl)) {
                bu = e.u;
                bl = e.lo;
                x1 = e.x;
                y1 = e.y;
                z1 = e.z;
            }
        }
        if (rank == 0 && tid == 0) idxs[j] = (int)(0x1FFFFFu - (bl & 0x1FFFFFu));
    }
    if (m > 1) {
#pragma unroll
        for (int i = 0; i < PPT; ++i) {
            const int k = tid + 1024 * (rank + CS * i);
            if (k < n) temp[k] = td[i];
        }
    }
    cluster.sync();  // no CTA may exit while a peer can still write into its mailbox
}

template <int CS>
int launch_cluster(int b, int n, int m, const float* dataset, float* temp, int* idxs, cudaStream_t st) {
    const int ppt = DCL_DIVUP(n, 1024 * CS);
    if (ppt <= 1) fps_cluster_kernel<1, CS><<<b * CS, 1024, 0, st>>>(n, m, dataset, temp, idxs);
    else if (ppt <= 2) fps_cluster_kernel<2, CS><<<b * CS, 1024, 0, st>>>(n, m, dataset, temp, idxs);
    else if (ppt <= 4) fps_cluster_kernel<4, CS><<<b * CS, 1024, 0, st>>>(n, m, dataset, temp, idxs);
    else fps_cluster_kernel<8, CS><<<b * CS, 1024, 0, st>>>(n, m, dataset, temp, idxs);
    return dcl_launch_status();
}

template <int PPT, int MODE>
int launch_resident(int b, int n, int m, int T, int threads, const float* dataset, float* temp, int* idxs,
                    cudaStream_t st) {
    const size_t smem = MODE == 1 ? (size_t)n * 12 : 0;
    if (smem > 48 * 1024)
        cudaFuncSetAttribute(fps_resident_kernel<PPT, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    fps_resident_kernel<PPT, MODE><<<b, threads, smem, st>>>(n, m, T, dataset, temp, idxs);
    return dcl_launch_status();
}

}  // namespace

DCL_API int dcl_lib_furthest_point_sampling_kernel_launcher(int b, int n, int m, const float* dataset, float* temp,
                                                            int* idxs, void* stream) {
    DCL_RETURN_IF_BAD(b >= 0 && n >= 1 && m >= 0);
    if (b == 0 || m == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const int T = ref_opt_n_threads(n);
    const int threads = T < 32 ? 32 : T;
    const int ppt = DCL_DIVUP(n, T);
    // Clouds of >= 2048 points: a thread-block cluster per cloud (4 CTAs; 8 when the batch alone cannot fill
    // the SMs or the cloud is too large for 4), points in registers.
    if (T == 1024 && n >= 2048 && n <= 1024 * 8 * 8) {
        const bool use8 = (n > 1024 * 4 * 8) || (b * 4 < 96);
        return use8 ? launch_cluster<8>(b, n, m, dataset, temp, idxs, st)
                    : launch_cluster<4>(b, n, m, dataset, temp, idxs, st);
    }
    if (ppt <= 1) return launch_resident<1, 0>(b, n, m, T, threads, dataset, temp, idxs, st);
    if (ppt <= 2) return launch_resident<2, 0>(b, n, m, T, threads, dataset, temp, idxs, st);
    if (ppt <= 4) return launch_resident<4, 0>(b, n, m, T, threads, dataset, temp, idxs, st);
    if (ppt <= 8) return launch_resident<8, 0>(b, n, m, T, threads, dataset, temp, idxs, st);
    if (ppt <= 16 && (size_t)n * 12 <= 200 * 1024)
        return launch_resident<16, 1>(b, n, m, T, threads, dataset, temp, idxs, st);
    fps_streaming_kernel<<<b, threads, 0, st>>>(n, m, T, dataset, temp, idxs);
    return dcl_launch_status();
}
