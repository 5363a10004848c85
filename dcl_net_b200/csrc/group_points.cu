// Neighbourhood gather / scatter-add:  grouping_operation and gather_operation.
// Semantics: libs/pointnet_lib/src/group_points_gpu.cu:47-66 (fwd), :8-25 (grad);
//            libs/pointnet_lib/src/sampling_gpu.cu:8-24 (gather fwd), :46-63 (grad).
// gather is grouping with nsample == 1, so both share these kernels.
//
// The reference launches one thread per output element and channel, with a random
// 4-byte global read each.  Here a CTA owns CG whole feature rows of one batch:
// the rows (contiguous CG*n floats) are staged in shared memory by one TMA bulk
// copy, gathers then hit shared-memory banks instead of L1 lines, the index
// stream is read once per CG channels with 128-bit loads and the output is
// written with 128-bit streaming stores.  Backward accumulates the CG rows in
// shared memory and flushes each row once.
#include "common.cuh"
#include "../../include/dcl_b200.h"

namespace {

constexpr int GP_THREADS = 256;
constexpr size_t GP_SMEM_SOFT = 64 * 1024;
constexpr size_t GP_SMEM_HARD = 200 * 1024;

__device__ __forceinline__ void stage_rows(float* s_rows, uint64_t* s_bar, const float* src, int nfl) {
    const bool tma = ((((uintptr_t)src) & 15u) == 0) && ((nfl & 3) == 0);
    if (tma) {
        if (threadIdx.x == 0) {
            dcl_mbar_init(s_bar, 1);
            dcl_fence_barrier_init();
            dcl_mbar_arrive_expect_tx(s_bar, (uint32_t)nfl * 4u);
            dcl_bulk_g2s(s_rows, src, (uint32_t)nfl * 4u, s_bar);
        }
        __syncthreads();
        dcl_mbar_wait(s_bar, 0);
    } else {
        for (int i = threadIdx.x; i < nfl; i += blockDim.x) s_rows[i] = __ldg(src + i);
        __syncthreads();
    }
}

// out[b, c, e] = points[b, c, idx[b, e]],  e in [0, E), E = npoints*nsample.
template <int CG>
__global__ void __launch_bounds__(GP_THREADS) group_fwd_kernel(int c, int n, int E, int e_per_cta,
                                                               const float* __restrict__ points,
                                                               const int* __restrict__ idx, float* __restrict__ out) {
    extern __shared__ __align__(16) float s_rows[];
    __shared__ uint64_t s_bar;
    const int bs = blockIdx.z;
    const int c0 = blockIdx.y * CG;
    const int ncg = min(CG, c - c0);
    stage_rows(s_rows, &s_bar, points + ((size_t)bs * c + c0) * n, ncg * n);
    idx += (size_t)bs * E;
    out += ((size_t)bs * c + c0) * E;
    const int e_begin = blockIdx.x * e_per_cta;
    const int e_end = min(E, e_begin + e_per_cta);
    const bool vec = ((E & 3) == 0) && ((((uintptr_t)idx) & 15u) == 0) && ((((uintptr_t)out) & 15u) == 0);
    if (vec) {
        for (int e = e_begin + threadIdx.x * 4; e < e_end; e += GP_THREADS * 4) {
            const int4 j = dcl_ld_stream_i4(idx + e);
#pragma unroll
            for (int cc = 0; cc < CG; ++cc) {
                if (cc < ncg) {
                    const float* row = s_rows + cc * n;
                    dcl_st_stream_f4(out + (size_t)cc * E + e, make_float4(row[j.x], row[j.y], row[j.z], row[j.w]));
                }
            }
        }
    } else {
        for (int e = e_begin + threadIdx.x; e < e_end; e += GP_THREADS) {
            const int j = idx[e];
#pragma unroll
            for (int cc = 0; cc < CG; ++cc)
                if (cc < ncg) out[(size_t)cc * E + e] = s_rows[cc * n + j];
        }
    }
}

__global__ void __launch_bounds__(GP_THREADS) group_fwd_gmem_kernel(int c, int n, int E,
                                                                    const float* __restrict__ points,
                                                                    const int* __restrict__ idx,
                                                                    float* __restrict__ out) {
    const int bs = blockIdx.z, cc = blockIdx.y;
    const int e = blockIdx.x * GP_THREADS + threadIdx.x;
    if (e >= E) return;
    out[((size_t)bs * c + cc) * E + e] = points[((size_t)bs * c + cc) * n + idx[(size_t)bs * E + e]];
}

// grad_points[b, c, idx[b, e]] += grad_out[b, c, e]
template <int CG>
__global__ void __launch_bounds__(GP_THREADS) group_grad_kernel(int c, int n, int E, int e_per_cta,
                                                                const float* __restrict__ grad_out,
                                                                const int* __restrict__ idx,
                                                                float* __restrict__ grad_points) {
    extern __shared__ __align__(16) float s_rows[];
    const int bs = blockIdx.z;
    const int c0 = blockIdx.y * CG;
    const int ncg = min(CG, c - c0);
    for (int i = threadIdx.x; i < ncg * n; i += GP_THREADS) s_rows[i] = 0.f;
    __syncthreads();
    idx += (size_t)bs * E;
    grad_out += ((size_t)bs * c + c0) * E;
    const int e_begin = blockIdx.x * e_per_cta;
    const int e_end = min(E, e_begin + e_per_cta);
    const bool vec = ((E & 3) == 0) && ((((uintptr_t)idx) & 15u) == 0) && ((((uintptr_t)grad_out) & 15u) == 0);
    if (vec) {
        for (int e = e_begin + threadIdx.x * 4; e < e_end; e += GP_THREADS * 4) {
            const int4 j = dcl_ld_stream_i4(idx + e);
#pragma unroll
            for (int cc = 0; cc < CG; ++cc) {
                if (cc < ncg) {
                    const float4 g = dcl_ld_stream_f4(grad_out + (size_t)cc * E + e);
                    float* row = s_rows + cc * n;
                    atomicAdd(row + j.x, g.x);
                    atomicAdd(row + j.y, g.y);
                    atomicAdd(row + j.z, g.z);
                    atomicAdd(row + j.w, g.w);
                }
            }
        }
    } else {
        for (int e = e_begin + threadIdx.x; e < e_end; e += GP_THREADS) {
            const int j = idx[e];
#pragma unroll
            for (int cc = 0; cc < CG; ++cc)
                if (cc < ncg) atomicAdd(s_rows + cc * n + j, grad_out[(size_t)cc * E + e]);
        }
    }
    __syncthreads();
    float* dst = grad_points + ((size_t)bs * c + c0) * n;
    if (gridDim.x == 1) {
        for (int i = threadIdx.x; i < ncg * n; i += GP_THREADS) dst[i] += s_rows[i];
    } else {
        for (int i = threadIdx.x; i < ncg * n; i += GP_THREADS) {
            const float v = s_rows[i];
            if (v != 0.f) atomicAdd(dst + i, v);
        }
    }
}

__global__ void __launch_bounds__(GP_THREADS) group_grad_gmem_kernel(int c, int n, int E,
                                                                     const float* __restrict__ grad_out,
                                                                     const int* __restrict__ idx,
                                                                     float* __restrict__ grad_points) {
    const int bs = blockIdx.z, cc = blockIdx.y;
    const int e = blockIdx.x * GP_THREADS + threadIdx.x;
    if (e >= E) return;
    atomicAdd(grad_points + ((size_t)bs * c + cc) * n + idx[(size_t)bs * E + e],
              grad_out[((size_t)bs * c + cc) * E + e]);
}

inline int pick_cg(int c, int n) {
    int cg = 8;
    while (cg > 1 && (size_t)cg * n * 4 > GP_SMEM_SOFT) cg >>= 1;
    if ((size_t)cg * n * 4 > GP_SMEM_HARD) return 0;
    while (cg > 1 && cg / 2 >= c) cg >>= 1;
    return cg;
}

inline int pick_e_per_cta(int E, long other_ctas) {
    const long want = 148L * 4;
    long split = DCL_DIVUP(want, other_ctas > 0 ? other_ctas : 1);
    if (split < 1) split = 1;
    long per = DCL_DIVUP((long)E, split);
    const long quantum = GP_THREADS * 4;
    per = DCL_DIVUP(per, quantum) * quantum;
    if (per < quantum * 2) per = quantum * 2;
    return (int)per;
}

template <typename K>
inline void allow_smem(K kernel, size_t bytes) {
    if (bytes > 48 * 1024) cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

int group_fwd(int b, int c, int n, int E, const float* points, const int* idx, float* out, cudaStream_t st) {
    if (b == 0 || c == 0 || E == 0) return 0;
    const int cg = pick_cg(c, n);
    if (cg == 0) {
        dim3 grid(DCL_DIVUP(E, GP_THREADS), c, b);
        group_fwd_gmem_kernel<<<grid, GP_THREADS, 0, st>>>(c, n, E, points, idx, out);
        return dcl_launch_status();
    }
    const int ngroups = DCL_DIVUP(c, cg);
    const int per = pick_e_per_cta(E, (long)ngroups * b);
    dim3 grid(DCL_DIVUP(E, per), ngroups, b);
    const size_t smem = (size_t)cg * n * 4;
#define DCL_LAUNCH(CG)                        \
    allow_smem(group_fwd_kernel<CG>, smem); \
    group_fwd_kernel<CG><<<grid, GP_THREADS, smem, st>>>(c, n, E, per, points, idx, out)
    switch (cg) {
        case 8: DCL_LAUNCH(8); break;
        case 4: DCL_LAUNCH(4); break;
        case 2: DCL_LAUNCH(2); break;
        default: DCL_LAUNCH(1); break;
    }
#undef DCL_LAUNCH
    return dcl_launch_status();
}

int group_grad(int b, int c, int n, int E, const float* grad_out, const int* idx, float* grad_points,
               cudaStream_t st) {
    if (b == 0 || c == 0 || E == 0) return 0;
    const int cg = pick_cg(c, n);
    if (cg == 0) {
        dim3 grid(DCL_DIVUP(E, GP_THREADS), c, b);
        group_grad_gmem_kernel<<<grid, GP_THREADS, 0, st>>>(c, n, E, grad_out, idx, grad_points);
        return dcl_launch_status();
    }
    const int ngroups = DCL_DIVUP(c, cg);
    const int per = pick_e_per_cta(E, (long)ngroups * b);
    dim3 grid(DCL_DIVUP(E, per), ngroups, b);
    const size_t smem = (size_t)cg * n * 4;
#define DCL_LAUNCH(CG)                         \
    allow_smem(group_grad_kernel<CG>, smem); \
    group_grad_kernel<CG><<<grid, GP_THREADS, smem, st>>>(c, n, E, per, grad_out, idx, grad_points)
    switch (cg) {
        case 8: DCL_LAUNCH(8); break;
        case 4: DCL_LAUNCH(4); break;
        case 2: DCL_LAUNCH(2); break;
        default: DCL_LAUNCH(1); break;
    }
#undef DCL_LAUNCH
    return dcl_launch_status();
}

}  // namespace

DCL_API int dcl_lib_group_points_kernel_launcher_fast(int b, int c, int n, int npoints, int nsample,
                                                      const float* points, const int* idx, float* out,
                                                      void* stream) {
    DCL_RETURN_IF_BAD(b >= 0 && c >= 0 && n >= 0 && npoints >= 0 && nsample >= 0);
    DCL_RETURN_IF_BAD((long)npoints * nsample < (1L << 31));
    return group_fwd(b, c, n, npoints * nsample, points, idx, out, (cudaStream_t)stream);
}

DCL_API int dcl_lib_group_points_grad_kernel_launcher_fast(int b, int c, int n, int npoints, int nsample,
                                                           const float* grad_out, const int* idx,
                                                           float* grad_points, void* stream) {
    DCL_RETURN_IF_BAD(b >= 0 && c >= 0 && n >= 0 && npoints >= 0 && nsample >= 0);
    DCL_RETURN_IF_BAD((long)npoints * nsample < (1L << 31));
    return group_grad(b, c, n, npoints * nsample, grad_out, idx, grad_points, (cudaStream_t)stream);
}

DCL_API int dcl_lib_gather_points_kernel_launcher_fast(int b, int c, int n, int npoints, const float* points,
                                                       const int* idx, float* out, void* stream) {
    DCL_RETURN_IF_BAD(b >= 0 && c >= 0 && n >= 0 && npoints >= 0);
    return group_fwd(b, c, n, npoints, points, idx, out, (cudaStream_t)stream);
}

DCL_API int dcl_lib_gather_points_grad_kernel_launcher_fast(int b, int c, int n, int npoints, const float* grad_out,
                                                            const int* idx, float* grad_points, void* stream) {
    DCL_RETURN_IF_BAD(b >= 0 && c >= 0 && n >= 0 && npoints >= 0);
    return group_grad(b, c, n, npoints, grad_out, idx, grad_points, (cudaStream_t)stream);
}
