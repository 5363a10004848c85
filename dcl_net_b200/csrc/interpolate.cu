// Batched (B,N,3) nearest-neighbour search and 3-point interpolation.
// Semantics follow libs/pointnet_lib/src/interpolate_gpu.cu of the reference:
//   knn_kernel_fast               :9-57    three_nn_kernel_fast            :81-124
//   three_interpolate_kernel_fast :149-169 three_interpolate_grad_kernel   :192-214
// Design: thread-per-query brute force kept (ascending candidate order is what
// defines the reference's tie-break), but candidates are streamed through shared
// memory by 1-D TMA bulk copies instead of being re-read from global by every
// warp, and each thread carries several queries so one shared-memory broadcast
// feeds several distance evaluations.
#include "common.cuh"
#include "tile_pipe.cuh"
#include "../../include/dcl_b200.h"
#include <math_constants.h>

#include <stdlib.h>

namespace {

constexpr int NN_THREADS = 128;
constexpr int NN_TILE_PTS = 1024;              // 12 KB per stage
constexpr int NN_TILE_FLOATS = NN_TILE_PTS * 3;

// Insert candidate (d,k) into the ascending triple.  Strict '<' everywhere: an
// equal distance never displaces an earlier index (reference :109-121).  The
// reference keeps the triple in double initialised to 1e40; every candidate is an
// fp32 value, so comparing in fp32 against +inf gives the same decisions and the
// same final cast ((float)1e40 == +inf).
__device__ __forceinline__ void nn3_insert(float d, int k, float& b1, float& b2, float& b3, int& i1, int& i2,
                                           int& i3) {
    if (d < b3) {
        if (d < b2) {
            b3 = b2;
            i3 = i2;
            if (d < b1) {
                b2 = b1;
                i2 = i1;
                b1 = d;
                i1 = k;
            } else {
                b2 = d;
                i2 = k;
            }
        } else {
            b3 = d;
            i3 = k;
        }
    }
}

template <int QPT>
__global__ void __launch_bounds__(NN_THREADS) three_nn_kernel(int n, int m, const float* __restrict__ unknown,
                                                              const float* __restrict__ known,
                                                              float* __restrict__ dist2, int* __restrict__ idx) {
    __shared__ __align__(16) float s_tile[2 * NN_TILE_FLOATS];
    __shared__ uint64_t s_bar[2];
    const int bs = blockIdx.y;
    unknown += (size_t)bs * n * 3;
    known += (size_t)bs * m * 3;
    dist2 += (size_t)bs * n * 3;
    idx += (size_t)bs * n * 3;

    const int q0 = (blockIdx.x * NN_THREADS + threadIdx.x) * QPT;
    float ux[QPT], uy[QPT], uz[QPT];
    float b1[QPT], b2[QPT], b3[QPT];
    int i1[QPT], i2[QPT], i3[QPT];
#pragma unroll
    for (int q = 0; q < QPT; ++q) {
        const int qi = min(q0 + q, n - 1);
        ux[q] = unknown[qi * 3 + 0];
        uy[q] = unknown[qi * 3 + 1];
        uz[q] = unknown[qi * 3 + 2];
        b1[q] = b2[q] = b3[q] = CUDART_INF_F;
        i1[q] = i2[q] = i3[q] = 0;
    }

    DclTilePipe<NN_TILE_FLOATS> pipe;
    pipe.init(s_tile, s_bar, known, m * 3);
    for (int t = 0; t < pipe.ntiles; ++t) {
        const int cnt = pipe.acquire(t) / 3;
        const float* tile = pipe.tile(t);
        const int kbase = t * NN_TILE_PTS;
#pragma unroll 4
        for (int j = 0; j < cnt; ++j) {
            const float x = tile[j * 3 + 0], y = tile[j * 3 + 1], z = tile[j * 3 + 2];
#pragma unroll
            for (int q = 0; q < QPT; ++q) {
                const float d = dcl_dist2(ux[q], uy[q], uz[q], x, y, z);
                nn3_insert(d, kbase + j, b1[q], b2[q], b3[q], i1[q], i2[q], i3[q]);
            }
        }
        pipe.release(t);
    }
#pragma unroll
    for (int q = 0; q < QPT; ++q) {
        const int qi = q0 + q;
        if (qi < n) {
            dist2[qi * 3 + 0] = b1[q];
            dist2[qi * 3 + 1] = b2[q];
            dist2[qi * 3 + 2] = b3[q];
            idx[qi * 3 + 0] = i1[q];
            idx[qi * 3 + 1] = i2[q];
            idx[qi * 3 + 2] = i3[q];
        }
    }
}

// k-NN with the list held in registers, right-aligned in a KMAX-slot array so
// that every access uses a compile-time index: slots [KMAX-k, KMAX) are live,
// the ones below hold -inf and stop the bubble.  A candidate enters at the last
// slot iff d < worst, then bubbles towards the front while strictly smaller than
// its predecessor -> equal distances stay in arrival (ascending index) order,
// which is the reference's "first j with d < best[j]" rule (:42-52).
template <int KMAX>
__global__ void __launch_bounds__(NN_THREADS) knn_reg_kernel(int n, int m, int k, const float* __restrict__ unknown,
                                                             const float* __restrict__ known,
                                                             float* __restrict__ dist2, int* __restrict__ idx) {
    __shared__ __align__(16) float s_tile[2 * NN_TILE_FLOATS];
    __shared__ uint64_t s_bar[2];
    const int bs = blockIdx.y;
    unknown += (size_t)bs * n * 3;
    known += (size_t)bs * m * 3;
    dist2 += (size_t)bs * n * k;
    idx += (size_t)bs * n * k;

    const int qi = blockIdx.x * NN_THREADS + threadIdx.x;
    const int qc = min(qi, n - 1);
    const float ux = unknown[qc * 3 + 0], uy = unknown[qc * 3 + 1], uz = unknown[qc * 3 + 2];
    float best[KMAX];
    int besti[KMAX];
#pragma unroll
    for (int j = 0; j < KMAX; ++j) {
        best[j] = (j >= KMAX - k) ? CUDART_INF_F : -CUDART_INF_F;
        besti[j] = 0;
    }

    DclTilePipe<NN_TILE_FLOATS> pipe;
    pipe.init(s_tile, s_bar, known, m * 3);
    for (int t = 0; t < pipe.ntiles; ++t) {
        const int cnt = pipe.acquire(t) / 3;
        const float* tile = pipe.tile(t);
        const int kbase = t * NN_TILE_PTS;
#pragma unroll 2
        for (int j = 0; j < cnt; ++j) {
            const float d = dcl_dist2(ux, uy, uz, tile[j * 3 + 0], tile[j * 3 + 1], tile[j * 3 + 2]);
            // Insertions are rare once the list has warmed up: make the branch warp-uniform so that the
            // (fully unrolled, ~6*KMAX instruction) bubble is skipped instead of executed predicated-off.
            if (!__any_sync(0xffffffffu, d < best[KMAX - 1])) continue;
            if (d < best[KMAX - 1]) {
                best[KMAX - 1] = d;
                besti[KMAX - 1] = kbase + j;
#pragma unroll
                for (int l = KMAX - 1; l > 0; --l) {
                    if (best[l] < best[l - 1]) {
                        const float tf = best[l];
                        best[l] = best[l - 1];
                        best[l - 1] = tf;
                        const int ti = besti[l];
                        besti[l] = besti[l - 1];
                        besti[l - 1] = ti;
                    }
                }
            }
        }
        pipe.release(t);
    }
    if (qi < n) {
#pragma unroll
        for (int j = 0; j < KMAX; ++j) {
            const int o = j - (KMAX - k);
            if (o >= 0) {
                dist2[(size_t)qi * k + o] = best[j];
                idx[(size_t)qi * k + o] = besti[j];
            }
        }
    }
}

// Any k (the reference allows up to 200): the list lives in the caller's output
// rows, which are private to the thread.
__global__ void __launch_bounds__(NN_THREADS) knn_big_kernel(int n, int m, int k, const float* __restrict__ unknown,
                                                             const float* __restrict__ known, float* dist2,
                                                             int* idx) {
    __shared__ __align__(16) float s_tile[2 * NN_TILE_FLOATS];
    __shared__ uint64_t s_bar[2];
    const int bs = blockIdx.y;
    unknown += (size_t)bs * n * 3;
    known += (size_t)bs * m * 3;
    dist2 += (size_t)bs * n * k;
    idx += (size_t)bs * n * k;
    const int qi = blockIdx.x * NN_THREADS + threadIdx.x;
    const int qc = min(qi, n - 1);
    const float ux = unknown[qc * 3 + 0], uy = unknown[qc * 3 + 1], uz = unknown[qc * 3 + 2];
    float* best = dist2 + (size_t)qc * k;
    int* besti = idx + (size_t)qc * k;
    if (qi < n) {
        for (int j = 0; j < k; ++j) {
            best[j] = CUDART_INF_F;
            besti[j] = 0;
        }
    }
    float worst = CUDART_INF_F;
    DclTilePipe<NN_TILE_FLOATS> pipe;
    pipe.init(s_tile, s_bar, known, m * 3);
    for (int t = 0; t < pipe.ntiles; ++t) {
        const int cnt = pipe.acquire(t) / 3;
        const float* tile = pipe.tile(t);
        const int kbase = t * NN_TILE_PTS;
        if (qi < n) {
            for (int j = 0; j < cnt; ++j) {
                const float d = dcl_dist2(ux, uy, uz, tile[j * 3 + 0], tile[j * 3 + 1], tile[j * 3 + 2]);
                if (d < worst) {
                    int l = k - 1;
                    while (l > 0 && d < best[l - 1]) {
                        best[l] = best[l - 1];
                        besti[l] = besti[l - 1];
                        --l;
                    }
                    best[l] = d;
                    besti[l] = kbase + j;
                    worst = best[k - 1];
                }
            }
        }
        pipe.release(t);
    }
}

// ---- warp-cooperative k-NN (k <= 32, known cloud resident in shared memory) ----
// A warp owns KW_Q queries at a time and scans 32 candidates per step (lane = candidate): coordinates are read once per
// step and tested against each query's current k-th distance.  The sorted list of a query is spread over the lanes
// (lane l = l-th neighbour); a candidate that beats the k-th distance is inserted where the sequential rule puts it —
// behind every entry with distance <= its own — by one ballot and two shuffles, candidates of a step in ascending
// order, each re-tested against the updated k-th distance: the list equals the reference's (interpolate_gpu.cu:42-52)
// element for element.  The thread-per-query kernel above keeps the list in one thread's registers and runs its
// ~6*KMAX-instruction bubble whenever ANY lane of the warp inserts, which with 32 independent queries per warp is
// nearly every candidate.  The CTA stages the whole known cloud once (m <= KW_MAX_M) and walks KW_ITERS query groups.
constexpr int KW_Q = 2;
constexpr int KW_WARPS = 8;
constexpr int KW_ITERS = 8;
constexpr int KW_THREADS = KW_WARPS * 32;
constexpr int KW_QPC = KW_WARPS * KW_Q * KW_ITERS;   // queries per CTA
constexpr int KW_MAX_M = 4096;                       // 48 KB of shared memory

__global__ void __launch_bounds__(KW_THREADS) knn_warp_kernel(int n, int m, int k, const float* __restrict__ unknown,
                                                              const float* __restrict__ known,
                                                              float* __restrict__ dist2, int* __restrict__ idx) {
    extern __shared__ __align__(16) float s_known[];
    __shared__ uint64_t s_bar;
    const int bs = blockIdx.y;
    unknown += (size_t)bs * n * 3;
    known += (size_t)bs * m * 3;
    dist2 += (size_t)bs * n * k;
    idx += (size_t)bs * n * k;
    const int nfl = m * 3;
    if (((((uintptr_t)known) & 15u) == 0) && ((nfl & 3) == 0)) {
        if (threadIdx.x == 0) {
            dcl_mbar_init(&s_bar, 1);
            dcl_fence_barrier_init();
            dcl_mbar_arrive_expect_tx(&s_bar, (uint32_t)nfl * 4u);
            dcl_bulk_g2s(s_known, known, (uint32_t)nfl * 4u, &s_bar);
        }
        __syncthreads();
        dcl_mbar_wait(&s_bar, 0);
    } else {
        for (int i = threadIdx.x; i < nfl; i += KW_THREADS) s_known[i] = __ldg(known + i);
        __syncthreads();
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool in_list = lane < k;
    for (int it = 0; it < KW_ITERS; ++it) {
        const int q0 = blockIdx.x * KW_QPC + (it * KW_WARPS + warp) * KW_Q;
        if (q0 >= n) break;
        float ux[KW_Q], uy[KW_Q], uz[KW_Q], bd[KW_Q], worst[KW_Q];
        int bi[KW_Q];
#pragma unroll
        for (int q = 0; q < KW_Q; ++q) {
            const int qc = min(q0 + q, n - 1);
            ux[q] = __ldg(unknown + qc * 3 + 0);
            uy[q] = __ldg(unknown + qc * 3 + 1);
            uz[q] = __ldg(unknown + qc * 3 + 2);
            bd[q] = CUDART_INF_F;
            bi[q] = 0;
            worst[q] = CUDART_INF_F;
        }
        for (int j0 = 0; j0 < m; j0 += 32) {
            const int j = j0 + lane;
            const bool in = j < m;
            const int jc = in ? j : m - 1;
            const float x = s_known[jc * 3 + 0], y = s_known[jc * 3 + 1], z = s_known[jc * 3 + 2];
#pragma unroll
            for (int q = 0; q < KW_Q; ++q) {
                const float d = dcl_dist2(ux[q], uy[q], uz[q], x, y, z);
                uint32_t mask = __ballot_sync(0xffffffffu, in && d < worst[q]);
                while (mask != 0u) {
                    const int src = __ffs(mask) - 1;
                    mask &= mask - 1u;
                    const float dc = __shfl_sync(0xffffffffu, d, src);
                    if (dc < worst[q]) {                                   // warp-uniform
                        // position = number of list entries with distance <= dc (equal distances keep arrival order)
                        const int p = __popc(__ballot_sync(0xffffffffu, in_list && bd[q] <= dc));
                        const float nd = __shfl_up_sync(0xffffffffu, bd[q], 1);
                        const int ni = __shfl_up_sync(0xffffffffu, bi[q], 1);
                        if (in_list && lane > p) {
                            bd[q] = nd;
                            bi[q] = ni;
                        } else if (lane == p) {
                            bd[q] = dc;
                            bi[q] = j0 + src;
                        }
                        worst[q] = __shfl_sync(0xffffffffu, bd[q], k - 1);
                    }
                }
            }
        }
#pragma unroll
        for (int q = 0; q < KW_Q; ++q) {
            if (q0 + q < n && in_list) {
                dist2[(size_t)(q0 + q) * k + lane] = bd[q];
                idx[(size_t)(q0 + q) * k + lane] = bi[q];
            }
        }
    }
}

// ------------------------------ interpolation ------------------------------
constexpr int IP_THREADS = 256;

// One CTA: batch bs, channel group [c0, c0+CG), query chunk.  The CG feature rows
// (contiguous CG*m floats) are staged in shared memory by one bulk copy; idx and
// weight are read once per CG channels instead of once per channel.
template <int CG>
__global__ void __launch_bounds__(IP_THREADS) three_interpolate_kernel(int c, int m, int n, int n_per_cta,
                                                                       const float* __restrict__ points,
                                                                       const int* __restrict__ idx,
                                                                       const float* __restrict__ weight,
                                                                       float* __restrict__ out) {
    extern __shared__ __align__(16) float s_rows[];
    __shared__ uint64_t s_bar;
    const int bs = blockIdx.z;
    const int c0 = blockIdx.y * CG;
    const int ncg = min(CG, c - c0);
    const float* src = points + ((size_t)bs * c + c0) * m;
    const int nfl = ncg * m;
    const bool tma = ((((uintptr_t)src) & 15u) == 0) && ((nfl & 3) == 0);
    if (tma) {
        if (threadIdx.x == 0) {
            dcl_mbar_init(&s_bar, 1);
            dcl_fence_barrier_init();
            dcl_mbar_arrive_expect_tx(&s_bar, (uint32_t)nfl * 4u);
            dcl_bulk_g2s(s_rows, src, (uint32_t)nfl * 4u, &s_bar);
        }
        __syncthreads();
        dcl_mbar_wait(&s_bar, 0);
    } else {
        for (int i = threadIdx.x; i < nfl; i += IP_THREADS) s_rows[i] = __ldg(src + i);
        __syncthreads();
    }
    idx += (size_t)bs * n * 3;
    weight += (size_t)bs * n * 3;
    out += ((size_t)bs * c + c0) * n;
    const int i_begin = blockIdx.x * n_per_cta;
    const int i_end = min(n, i_begin + n_per_cta);
    for (int i = i_begin + threadIdx.x; i < i_end; i += IP_THREADS) {
        const int j0 = idx[i * 3 + 0], j1 = idx[i * 3 + 1], j2 = idx[i * 3 + 2];
        const float w0 = weight[i * 3 + 0], w1 = weight[i * 3 + 1], w2 = weight[i * 3 + 2];
#pragma unroll
        for (int cc = 0; cc < CG; ++cc) {
            if (cc < ncg) {
                const float* row = s_rows + cc * m;
                dcl_st_stream_f1(out + (size_t)cc * n + i, dcl_interp3(w0, row[j0], w1, row[j1], w2, row[j2]));
            }
        }
    }
}

// Channel-interleaved form of the kernel above for groups of 8 channels: the staged rows are stored point-major,
// s_il[point][8 channels] = two 16-byte halves per point, so that one query's three gathers are 6 LDS.128 instead of
// 24 LDS.32 — the ~3.5-way bank conflicts of 32 random 4-byte gathers per instruction (ncu: 16 M conflict cycles per
// launch at the microbench shape, the kernel's limiter at 38 % of the HBM roofline) become the ~2.4-way conflicts of 8
// random 16-byte gathers per quarter-warp phase on a quarter of the instructions.  The two halves of a point are
// swizzled (half h of point j at 16-byte slot 2j + (h ^ ((j >> 2) & 1))) so that a phase's eight lanes spread over all
// eight slots mod 8, for the staging stores (consecutive j) as well as for the gathers.  Same fma order as above:
// bit-identical results.
__global__ void __launch_bounds__(IP_THREADS) three_interpolate_il8_kernel(int c, int m, int n, int n_per_cta,
                                                                           const float* __restrict__ points,
                                                                           const int* __restrict__ idx,
                                                                           const float* __restrict__ weight,
                                                                           float* __restrict__ out) {
    extern __shared__ __align__(16) float4 s_il[];     // [m][2]
    const int bs = blockIdx.z;
    const int c0 = blockIdx.y * 8;
    const int ncg = min(8, c - c0);
    const float* src = points + ((size_t)bs * c + c0) * m;
    for (int j = threadIdx.x; j < m; j += IP_THREADS) {
        float v[8];
#pragma unroll
        for (int cc = 0; cc < 8; ++cc) v[cc] = cc < ncg ? __ldg(src + (size_t)cc * m + j) : 0.f;
        const int sw = (j >> 2) & 1;
        s_il[2 * j + sw] = make_float4(v[0], v[1], v[2], v[3]);
        s_il[2 * j + (sw ^ 1)] = make_float4(v[4], v[5], v[6], v[7]);
    }
    __syncthreads();
    idx += (size_t)bs * n * 3;
    weight += (size_t)bs * n * 3;
    out += ((size_t)bs * c + c0) * n;
    const int i_begin = blockIdx.x * n_per_cta;
    const int i_end = min(n, i_begin + n_per_cta);
    for (int i = i_begin + threadIdx.x; i < i_end; i += IP_THREADS) {
        const int j0 = idx[i * 3 + 0], j1 = idx[i * 3 + 1], j2 = idx[i * 3 + 2];
        const float w0 = weight[i * 3 + 0], w1 = weight[i * 3 + 1], w2 = weight[i * 3 + 2];
        const int s0 = (j0 >> 2) & 1, s1 = (j1 >> 2) & 1, s2 = (j2 >> 2) & 1;
        const float4 a0 = s_il[2 * j0 + s0], a1 = s_il[2 * j1 + s1], a2 = s_il[2 * j2 + s2];
        const float4 b0 = s_il[2 * j0 + (s0 ^ 1)], b1 = s_il[2 * j1 + (s1 ^ 1)], b2 = s_il[2 * j2 + (s2 ^ 1)];
        const float r[8] = {dcl_interp3(w0, a0.x, w1, a1.x, w2, a2.x), dcl_interp3(w0, a0.y, w1, a1.y, w2, a2.y),
                            dcl_interp3(w0, a0.z, w1, a1.z, w2, a2.z), dcl_interp3(w0, a0.w, w1, a1.w, w2, a2.w),
                            dcl_interp3(w0, b0.x, w1, b1.x, w2, b2.x), dcl_interp3(w0, b0.y, w1, b1.y, w2, b2.y),
                            dcl_interp3(w0, b0.z, w1, b1.z, w2, b2.z), dcl_interp3(w0, b0.w, w1, b1.w, w2, b2.w)};
#pragma unroll
        for (int cc = 0; cc < 8; ++cc)
            if (cc < ncg) dcl_st_stream_f1(out + (size_t)cc * n + i, r[cc]);
    }
}

// ---- lane-per-channel backward (used when 32 transposed rows fit in shared memory) ----
// The CG-row backward above scatters into row[j] with 32 random j per warp instruction: ~3.5-way conflicting
// shared-memory atomics, three per element.  Here a CTA accumulates into acc_t[point][channel] (stride 33) and a
// warp works on blocks of 32 queries x 32 channels with lane = channel: the three scatter-adds of a query are
// conflict-free, the block's idx / weight rows are staged once and read back as broadcasts, and the gradient block
// is read as 128-byte rows of consecutive queries and transposed through a per-warp tile.  (The same layout was
// tried for the forward gather and lost to the CG-row kernel, 162 vs 121 us: one CTA of 16 warps per SM cannot hide
// the dependent shared-memory chain, while the conflicting gathers above run with several CTAs per SM.)
constexpr int IPL_THREADS = 512;
constexpr int IPL_STRIDE = 33;
constexpr int IPL_MAX_SMEM = 224 * 1024;
constexpr int IPL_WARP_FLOATS = 32 * IPL_STRIDE + 32 * 8;  // per warp: transpose tile + the block's (idx, weight) rows

__global__ void __launch_bounds__(IPL_THREADS) three_interpolate_grad_lane_kernel(int c, int n, int m, int n_per_cta,
                                                                                  const float* __restrict__ grad_out,
                                                                                  const int* __restrict__ idx,
                                                                                  const float* __restrict__ weight,
                                                                                  float* __restrict__ grad_points) {
    extern __shared__ __align__(16) float s_mem[];
    float* acc_t = s_mem;                                    // [m][33]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* tile = s_mem + (size_t)m * IPL_STRIDE + warp * IPL_WARP_FLOATS;
    float4* qrow = reinterpret_cast<float4*>(tile + 32 * IPL_STRIDE);
    const int bs = blockIdx.z, c0 = blockIdx.y * 32;
    const int ncg = min(32, c - c0);
    for (int i = threadIdx.x; i < m * IPL_STRIDE; i += IPL_THREADS) acc_t[i] = 0.f;
    __syncthreads();
    idx += (size_t)bs * n * 3;
    weight += (size_t)bs * n * 3;
    grad_out += ((size_t)bs * c + c0) * n;
    const int i_begin = blockIdx.x * n_per_cta;
    const int i_end = min(n, i_begin + n_per_cta);
    for (int i0 = i_begin + warp * 32; i0 < i_end; i0 += IPL_THREADS) {
        const bool live = i0 + lane < i_end;
        const int iq = min(i0 + lane, i_end - 1);
        const int mj0 = __ldg(idx + iq * 3 + 0), mj1 = __ldg(idx + iq * 3 + 1), mj2 = __ldg(idx + iq * 3 + 2);
        const float mw0 = __ldg(weight + iq * 3 + 0), mw1 = __ldg(weight + iq * 3 + 1), mw2 = __ldg(weight + iq * 3 + 2);
#pragma unroll 8
        for (int cc = 0; cc < 32; ++cc)
            tile[cc * IPL_STRIDE + lane] = (live && cc < ncg) ? __ldg(grad_out + (size_t)cc * n + i0 + lane) : 0.f;
        qrow[lane * 2 + 0] = make_float4(__int_as_float(mj0 * IPL_STRIDE), __int_as_float(mj1 * IPL_STRIDE),
                                         __int_as_float(mj2 * IPL_STRIDE), 0.f);
        qrow[lane * 2 + 1] = make_float4(mw0, mw1, mw2, 0.f);
        __syncwarp();
        const int nq = min(32, i_end - i0);
#pragma unroll 4
        for (int q = 0; q < nq; ++q) {
            const float4 jj = qrow[q * 2 + 0], ww = qrow[q * 2 + 1];
            const float g = tile[lane * IPL_STRIDE + q];
            atomicAdd(acc_t + __float_as_int(jj.x) + lane, __fmul_rn(g, ww.x));
            atomicAdd(acc_t + __float_as_int(jj.y) + lane, __fmul_rn(g, ww.y));
            atomicAdd(acc_t + __float_as_int(jj.z) + lane, __fmul_rn(g, ww.z));
        }
        __syncwarp();
    }
    __syncthreads();
    float* dst = grad_points + ((size_t)bs * c + c0) * m;
    for (int cc = warp; cc < ncg; cc += IPL_THREADS / 32) {
        for (int pt = lane; pt < m; pt += 32) {
            const float v = acc_t[pt * IPL_STRIDE + cc];
            if (gridDim.x == 1) dst[(size_t)cc * m + pt] += v;
            else if (v != 0.f) atomicAdd(dst + (size_t)cc * m + pt, v);
        }
    }
}

// query chunks per (batch, channel group) so that the grid is a whole number of waves of one CTA per SM
int ipl_chunks(int n, int groups_times_b) {
    int best = 1;
    double best_eff = 0.0;
    for (int ch = 1; ch <= 16; ++ch) {
        if ((long)ch * 256 > n && ch > 1) break;
        const long ctas = (long)ch * groups_times_b;
        const long waves = (ctas + 147) / 148;
        const double eff = (double)ctas / (double)(waves * 148);
        if (eff > best_eff + 0.02) {
            best_eff = eff;
            best = ch;
        }
    }
    return best;
}

size_t ipl_smem(int m) { return ((size_t)m * IPL_STRIDE + (size_t)(IPL_THREADS / 32) * IPL_WARP_FLOATS) * sizeof(float); }

// Rows too long for shared memory: gather straight from global / L2.
__global__ void __launch_bounds__(IP_THREADS) three_interpolate_gmem_kernel(int c, int m, int n,
                                                                            const float* __restrict__ points,
                                                                            const int* __restrict__ idx,
                                                                            const float* __restrict__ weight,
                                                                            float* __restrict__ out) {
    const int bs = blockIdx.z, cc = blockIdx.y;
    const int i = blockIdx.x * IP_THREADS + threadIdx.x;
    if (i >= n) return;
    const float* row = points + ((size_t)bs * c + cc) * m;
    const int* ip = idx + ((size_t)bs * n + i) * 3;
    const float* wp = weight + ((size_t)bs * n + i) * 3;
    out[((size_t)bs * c + cc) * n + i] = dcl_interp3(wp[0], row[ip[0]], wp[1], row[ip[1]], wp[2], row[ip[2]]);
}

// Backward: per-CTA shared-memory accumulation of CG rows, flushed once.
template <int CG>
__global__ void __launch_bounds__(IP_THREADS) three_interpolate_grad_kernel(int c, int n, int m, int n_per_cta,
                                                                            const float* __restrict__ grad_out,
                                                                            const int* __restrict__ idx,
                                                                            const float* __restrict__ weight,
                                                                            float* __restrict__ grad_points) {
    extern __shared__ __align__(16) float s_rows[];
    const int bs = blockIdx.z;
    const int c0 = blockIdx.y * CG;
    const int ncg = min(CG, c - c0);
    for (int i = threadIdx.x; i < ncg * m; i += IP_THREADS) s_rows[i] = 0.f;
    __syncthreads();
    idx += (size_t)bs * n * 3;
    weight += (size_t)bs * n * 3;
    grad_out += ((size_t)bs * c + c0) * n;
    const int i_begin = blockIdx.x * n_per_cta;
    const int i_end = min(n, i_begin + n_per_cta);
    for (int i = i_begin + threadIdx.x; i < i_end; i += IP_THREADS) {
        const int j0 = idx[i * 3 + 0], j1 = idx[i * 3 + 1], j2 = idx[i * 3 + 2];
        const float w0 = weight[i * 3 + 0], w1 = weight[i * 3 + 1], w2 = weight[i * 3 + 2];
#pragma unroll
        for (int cc = 0; cc < CG; ++cc) {
            if (cc < ncg) {
                const float g = __ldg(grad_out + (size_t)cc * n + i);
                float* row = s_rows + cc * m;
                atomicAdd(row + j0, __fmul_rn(g, w0));
                atomicAdd(row + j1, __fmul_rn(g, w1));
                atomicAdd(row + j2, __fmul_rn(g, w2));
            }
        }
    }
    __syncthreads();
    float* dst = grad_points + ((size_t)bs * c + c0) * m;
    if (gridDim.x == 1) {
        for (int i = threadIdx.x; i < ncg * m; i += IP_THREADS) dst[i] += s_rows[i];
    } else {
        for (int i = threadIdx.x; i < ncg * m; i += IP_THREADS) {
            const float v = s_rows[i];
            if (v != 0.f) atomicAdd(dst + i, v);
        }
    }
}

__global__ void __launch_bounds__(IP_THREADS) three_interpolate_grad_gmem_kernel(int c, int n, int m,
                                                                                 const float* __restrict__ grad_out,
                                                                                 const int* __restrict__ idx,
                                                                                 const float* __restrict__ weight,
                                                                                 float* __restrict__ grad_points) {
    const int bs = blockIdx.z, cc = blockIdx.y;
    const int i = blockIdx.x * IP_THREADS + threadIdx.x;
    if (i >= n) return;
    const float g = grad_out[((size_t)bs * c + cc) * n + i];
    const int* ip = idx + ((size_t)bs * n + i) * 3;
    const float* wp = weight + ((size_t)bs * n + i) * 3;
    float* row = grad_points + ((size_t)bs * c + cc) * m;
    atomicAdd(row + ip[0], __fmul_rn(g, wp[0]));
    atomicAdd(row + ip[1], __fmul_rn(g, wp[1]));
    atomicAdd(row + ip[2], __fmul_rn(g, wp[2]));
}

constexpr size_t SMEM_SOFT = 64 * 1024;   // keeps >= 3 CTAs per SM
constexpr size_t SMEM_HARD = 200 * 1024;

inline int pick_cg(int c, int m) {
    int cg = 8;
    while (cg > 1 && (size_t)cg * m * 4 > SMEM_SOFT) cg >>= 1;
    if ((size_t)cg * m * 4 > SMEM_HARD) return 0;
    while (cg > 1 && cg / 2 >= c) cg >>= 1;
    return cg;
}

// Split the query axis so that the grid has at least ~4 CTAs per SM.
inline int pick_n_per_cta(int n, int other_ctas) {
    const int want = 148 * 4;
    int split = DCL_DIVUP(want, other_ctas > 0 ? other_ctas : 1);
    if (split < 1) split = 1;
    int per = DCL_DIVUP(n, split);
    per = DCL_DIVUP(per, IP_THREADS) * IP_THREADS;
    if (per < IP_THREADS * 4) per = IP_THREADS * 4;
    return per;
}

template <typename K>
inline void allow_smem(K kernel, size_t bytes) {
    if (bytes > 48 * 1024) cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

}  // namespace

DCL_API int dcl_lib_three_nn_kernel_launcher_fast(int b, int n, int m, const float* unknown, const float* known,
                                                  float* dist2, int* idx, void* stream) {
    DCL_RETURN_IF_BAD(b >= 0 && n >= 0 && m >= 0);
    if (b == 0 || n == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    // Enough CTAs to fill 148 SMs with one query per thread?  Otherwise there is work to
    // amortise: carry 4 (or 2) queries per thread.
    const long ctas1 = (long)DCL_DIVUP(n, NN_THREADS) * b;
    if (ctas1 >= 148L * 8 * 4) {
        dim3 grid(DCL_DIVUP(n, NN_THREADS * 4), b);
        three_nn_kernel<4><<<grid, NN_THREADS, 0, st>>>(n, m, unknown, known, dist2, idx);
    } else if (ctas1 >= 148L * 8 * 2) {
        dim3 grid(DCL_DIVUP(n, NN_THREADS * 2), b);
        three_nn_kernel<2><<<grid, NN_THREADS, 0, st>>>(n, m, unknown, known, dist2, idx);
    } else {
        dim3 grid(DCL_DIVUP(n, NN_THREADS), b);
        three_nn_kernel<1><<<grid, NN_THREADS, 0, st>>>(n, m, unknown, known, dist2, idx);
    }
    return dcl_launch_status();
}

DCL_API int dcl_lib_knn_kernel_launcher_fast(int b, int n, int m, int k, const float* unknown, const float* known,
                                             float* dist2, int* idx, void* stream) {
    DCL_RETURN_IF_BAD(b >= 0 && n >= 0 && m >= 0 && k >= 1 && k <= 200);
    if (b == 0 || n == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    // warp-cooperative kernel when the known cloud fits in shared memory; DCL_KNN_THREAD=1 keeps the thread-per-query
    // kernels (A/B runs)
    static const bool force_thread = getenv("DCL_KNN_THREAD") != nullptr;
    if (!force_thread && k <= 32 && m >= 1 && m <= KW_MAX_M) {
        const size_t smem = (size_t)m * 3 * 4;
        if (smem > 40 * 1024) cudaFuncSetAttribute(knn_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        knn_warp_kernel<<<dim3(DCL_DIVUP(n, KW_QPC), b), KW_THREADS, smem, st>>>(n, m, k, unknown, known, dist2, idx);
        return dcl_launch_status();
    }
    dim3 grid(DCL_DIVUP(n, NN_THREADS), b);
    if (k <= 4)
        knn_reg_kernel<4><<<grid, NN_THREADS, 0, st>>>(n, m, k, unknown, known, dist2, idx);
    else if (k <= 8)
        knn_reg_kernel<8><<<grid, NN_THREADS, 0, st>>>(n, m, k, unknown, known, dist2, idx);
    else if (k <= 16)
        knn_reg_kernel<16><<<grid, NN_THREADS, 0, st>>>(n, m, k, unknown, known, dist2, idx);
    else if (k <= 32)
        knn_reg_kernel<32><<<grid, NN_THREADS, 0, st>>>(n, m, k, unknown, known, dist2, idx);
    else
        knn_big_kernel<<<grid, NN_THREADS, 0, st>>>(n, m, k, unknown, known, dist2, idx);
    return dcl_launch_status();
}

DCL_API int dcl_lib_three_interpolate_kernel_launcher_fast(int b, int c, int m, int n, const float* points,
                                                           const int* idx, const float* weight, float* out,
                                                           void* stream) {
    DCL_RETURN_IF_BAD(b >= 0 && c >= 0 && m >= 0 && n >= 0);
    if (b == 0 || c == 0 || n == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const int cg = pick_cg(c, m);
    if (cg == 0) {
        dim3 grid(DCL_DIVUP(n, IP_THREADS), c, b);
        three_interpolate_gmem_kernel<<<grid, IP_THREADS, 0, st>>>(c, m, n, points, idx, weight, out);
        return dcl_launch_status();
    }
    const int ngroups = DCL_DIVUP(c, cg);
    const int per = pick_n_per_cta(n, ngroups * b);
    dim3 grid(DCL_DIVUP(n, per), ngroups, b);
    const size_t smem = (size_t)cg * m * 4;
    // groups of 8 channels: the channel-interleaved kernel; DCL_INTERP_ROWS=1 keeps the row-staged one (A/B runs)
    static const bool force_rows = getenv("DCL_INTERP_ROWS") != nullptr;
    if (cg == 8 && !force_rows) {
        allow_smem(three_interpolate_il8_kernel, smem);
        three_interpolate_il8_kernel<<<grid, IP_THREADS, smem, st>>>(c, m, n, per, points, idx, weight, out);
        return dcl_launch_status();
    }
#define DCL_LAUNCH_IP(CG)                                                                                       \
    allow_smem(three_interpolate_kernel<CG>, smem);                                                             \
    three_interpolate_kernel<CG><<<grid, IP_THREADS, smem, st>>>(c, m, n, per, points, idx, weight, out)
    switch (cg) {
        case 8: DCL_LAUNCH_IP(8); break;
        case 4: DCL_LAUNCH_IP(4); break;
        case 2: DCL_LAUNCH_IP(2); break;
        default: DCL_LAUNCH_IP(1); break;
    }
#undef DCL_LAUNCH_IP
    return dcl_launch_status();
}

DCL_API int dcl_lib_three_interpolate_grad_kernel_launcher_fast(int b, int c, int n, int m, const float* grad_out,
                                                                const int* idx, const float* weight,
                                                                float* grad_points, void* stream) {
    DCL_RETURN_IF_BAD(b >= 0 && c >= 0 && m >= 0 && n >= 0);
    if (b == 0 || c == 0 || n == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    if (c >= 16 && m >= 4 && m % 4 == 0 && n >= 64 && ipl_smem(m) <= IPL_MAX_SMEM) {
        const int groups = DCL_DIVUP(c, 32);
        const int chunks = ipl_chunks(n, groups * b);
        const int per = DCL_DIVUP(DCL_DIVUP(n, chunks), 32) * 32;
        allow_smem(three_interpolate_grad_lane_kernel, ipl_smem(m));
        three_interpolate_grad_lane_kernel<<<dim3(DCL_DIVUP(n, per), groups, b), IPL_THREADS, ipl_smem(m), st>>>(
            c, n, m, per, grad_out, idx, weight, grad_points);
        return dcl_launch_status();
    }
    const int cg = pick_cg(c, m);
    if (cg == 0) {
        dim3 grid(DCL_DIVUP(n, IP_THREADS), c, b);
        three_interpolate_grad_gmem_kernel<<<grid, IP_THREADS, 0, st>>>(c, n, m, grad_out, idx, weight, grad_points);
        return dcl_launch_status();
    }
    const int ngroups = DCL_DIVUP(c, cg);
    const int per = pick_n_per_cta(n, ngroups * b);
    dim3 grid(DCL_DIVUP(n, per), ngroups, b);
    const size_t smem = (size_t)cg * m * 4;
#define DCL_LAUNCH_IPG(CG)                                                                                      \
    allow_smem(three_interpolate_grad_kernel<CG>, smem);                                                        \
    three_interpolate_grad_kernel<CG><<<grid, IP_THREADS, smem, st>>>(c, n, m, per, grad_out, idx, weight,      \
                                                                      grad_points)
    switch (cg) {
        case 8: DCL_LAUNCH_IPG(8); break;
        case 4: DCL_LAUNCH_IPG(4); break;
        case 2: DCL_LAUNCH_IPG(2); break;
        default: DCL_LAUNCH_IPG(1); break;
    }
#undef DCL_LAUNCH_IPG
    return dcl_launch_status();
}
