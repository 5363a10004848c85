// Batched 3x3 pose solve.
//  * dcl_svd3_project  — the SVD tail of ortho9d2matrix (models/DCL_Net.py:15-36,
//    models/refiner.py:35-56; normalize_vector utils/transform3D.py:16-21):
//        M = [x y z] (columns, each v/(|v|+1e-8));  U,S,V = svd(M);
//        R = U diag(1,1,det(U V^T)) V^T
//  * dcl_weighted_kabsch — confidence-weighted Kabsch (north_star part 3; not in the
//    reference): warp-per-instance weighted covariance, same projection.
//  * dcl_pose_compose — stage-2 composition, tools/test_YCBV_stage2.py:222-225.
//
// The projection does not need U, S, V individually.  With a one-sided Jacobi
// iteration  A = M V  (V a product of proper rotations, det V = +1) the columns of A
// converge to sigma_i u_i.  Let c be the column with the smallest norm and (a,b,c) a
// cyclic permutation of (0,1,2); then  d * u_c = u_a x u_b  with d = det(U V^T), hence
//        R = u_a v_a^T + u_b v_b^T + (u_a x u_b) v_c^T
// — no sorting, no explicit determinant, and the smallest singular direction (the
// ill-conditioned one) is never normalised.  The iteration runs in fp64 so that the
// result is the exact projection to well below the reference's own fp32 LAPACK error.
#include "common.cuh"
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include "../../include/dcl_b200.h"

namespace {

struct Mat3d {
    double m[3][3];
};

__device__ __forceinline__ void jacobi_rotate(double A[3][3], double V[3][3], int p, int q) {
    double alpha = 0, beta = 0, gamma = 0;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        alpha += A[r][p] * A[r][p];
        beta += A[r][q] * A[r][q];
        gamma += A[r][p] * A[r][q];
    }
    if (gamma == 0.0 || fabs(gamma) <= 1e-300) return;
    if (fabs(gamma) <= 1e-17 * sqrt(alpha * beta)) return;
    const double zeta = (beta - alpha) / (2.0 * gamma);
    const double tt = copysign(1.0, zeta) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
    const double cs = 1.0 / sqrt(1.0 + tt * tt);
    const double sn = cs * tt;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const double ap = A[r][p], aq = A[r][q];
        A[r][p] = cs * ap - sn * aq;
        A[r][q] = sn * ap + cs * aq;
        const double vp = V[r][p], vq = V[r][q];
        V[r][p] = cs * vp - sn * vq;
        V[r][q] = sn * vp + cs * vq;
    }
}

// R (row-major) = projection of M (row-major) onto SO(3) as defined above.
__device__ void project_so3(const double Min[3][3], float* __restrict__ R) {
    double A[3][3], V[3][3];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            A[r][c] = Min[r][c];
            V[r][c] = (r == c) ? 1.0 : 0.0;
        }
    for (int sweep = 0; sweep < 12; ++sweep) {
        jacobi_rotate(A, V, 0, 1);
        jacobi_rotate(A, V, 0, 2);
        jacobi_rotate(A, V, 1, 2);
        // |a_p . a_q| summed over the three column pairs against the total squared norm
        double g01 = 0, g02 = 0, g12 = 0, diag = 0;
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            g01 += A[r][0] * A[r][1];
            g02 += A[r][0] * A[r][2];
            g12 += A[r][1] * A[r][2];
            diag += A[r][0] * A[r][0] + A[r][1] * A[r][1] + A[r][2] * A[r][2];
        }
        const double off = fabs(g01) + fabs(g02) + fabs(g12);
        if (off <= 1e-15 * diag) break;
    }
    double nrm[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) nrm[c] = A[0][c] * A[0][c] + A[1][c] * A[1][c] + A[2][c] * A[2][c];
    int cmin = 0;
    if (nrm[1] < nrm[cmin]) cmin = 1;
    if (nrm[2] < nrm[cmin]) cmin = 2;
    const int ia = (cmin + 1) % 3, ib = (cmin + 2) % 3;
    const double ra = rsqrt(nrm[ia] > 0 ? nrm[ia] : 1.0), rb = rsqrt(nrm[ib] > 0 ? nrm[ib] : 1.0);
    double ua[3], ub[3], uc[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        ua[r] = A[r][ia] * ra;
        ub[r] = A[r][ib] * rb;
    }
    uc[0] = ua[1] * ub[2] - ua[2] * ub[1];
    uc[1] = ua[2] * ub[0] - ua[0] * ub[2];
    uc[2] = ua[0] * ub[1] - ua[1] * ub[0];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c)
            R[r * 3 + c] = (float)(ua[r] * V[c][ia] + ub[r] * V[c][ib] + uc[r] * V[c][cmin]);
}

__global__ void __launch_bounds__(128) svd3_project_kernel(int b, const float* __restrict__ in9, int normalize,
                                                           float* __restrict__ R) {
    const int i = blockIdx.x * 128 + threadIdx.x;
    if (i >= b) return;
    const float* v = in9 + (size_t)i * 9;
    double M[3][3];
    if (normalize) {
        // fp32 exactly as torch evaluates it: sqrt(sum(v^2)) + 1e-8, then v / mag.
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float a = v[c * 3 + 0], bb = v[c * 3 + 1], cc = v[c * 3 + 2];
            const float ss = __fadd_rn(__fadd_rn(__fmul_rn(a, a), __fmul_rn(bb, bb)), __fmul_rn(cc, cc));
            const float mag = __fadd_rn(__fsqrt_rn(ss), 1e-8f);
            M[0][c] = (double)__fdiv_rn(a, mag);
            M[1][c] = (double)__fdiv_rn(bb, mag);
            M[2][c] = (double)__fdiv_rn(cc, mag);
        }
    } else {
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c) M[r][c] = (double)v[r * 3 + c];
    }
    project_so3(M, R + (size_t)i * 9);
}

// One warp per instance.  Deterministic: fixed lane->point map and a fixed shuffle tree.
__global__ void __launch_bounds__(128) weighted_kabsch_kernel(int b, int n, const float* __restrict__ src,
                                                              const float* __restrict__ dst,
                                                              const float* __restrict__ w, float* __restrict__ R,
                                                              float* __restrict__ t) {
    const int inst = blockIdx.x * 4 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (inst >= b) return;
    src += (size_t)inst * n * 3;
    dst += (size_t)inst * n * 3;
    w += (size_t)inst * n;
    // pass 1: weighted means
    double acc[7] = {0, 0, 0, 0, 0, 0, 0};
    for (int i = lane; i < n; i += 32) {
        const double wi = w[i];
        acc[0] += wi;
        acc[1] += wi * src[i * 3 + 0];
        acc[2] += wi * src[i * 3 + 1];
        acc[3] += wi * src[i * 3 + 2];
        acc[4] += wi * dst[i * 3 + 0];
        acc[5] += wi * dst[i * 3 + 1];
        acc[6] += wi * dst[i * 3 + 2];
    }
#pragma unroll
    for (int k = 0; k < 7; ++k)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
    const double inv = acc[0] != 0.0 ? 1.0 / acc[0] : 0.0;
    const double pm[3] = {acc[1] * inv, acc[2] * inv, acc[3] * inv};
    const double qm[3] = {acc[4] * inv, acc[5] * inv, acc[6] * inv};
    // pass 2: M = H^T = sum w (q - qm)(p - pm)^T   (row index from dst, column from src)
    double h[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = lane; i < n; i += 32) {
        const double wi = w[i];
        const double p0 = src[i * 3 + 0] - pm[0], p1 = src[i * 3 + 1] - pm[1], p2 = src[i * 3 + 2] - pm[2];
        const double q0 = wi * (dst[i * 3 + 0] - qm[0]), q1 = wi * (dst[i * 3 + 1] - qm[1]),
                     q2 = wi * (dst[i * 3 + 2] - qm[2]);
        h[0] += q0 * p0; h[1] += q0 * p1; h[2] += q0 * p2;
        h[3] += q1 * p0; h[4] += q1 * p1; h[5] += q1 * p2;
        h[6] += q2 * p0; h[7] += q2 * p1; h[8] += q2 * p2;
    }
#pragma unroll
    for (int k = 0; k < 9; ++k)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) h[k] += __shfl_xor_sync(0xffffffffu, h[k], o);
    if (lane == 0) {
        double M[3][3];
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c) M[r][c] = h[r * 3 + c];
        float Rl[9];
        project_so3(M, Rl);
#pragma unroll
        for (int k = 0; k < 9; ++k) R[(size_t)inst * 9 + k] = Rl[k];
#pragma unroll
        for (int r = 0; r < 3; ++r)
            t[(size_t)inst * 3 + r] =
                (float)(qm[r] - ((double)Rl[r * 3 + 0] * pm[0] + (double)Rl[r * 3 + 1] * pm[1] +
                                 (double)Rl[r * 3 + 2] * pm[2]));
    }
}

// t <- R dt + t ; R <- R dR ; out[b, ch, i] = sum_r (p[b,i,r] - t[r]) R[r][ch]   (channel-major)
// fp32 with the operation order of torch (matmul rows accumulate left to right).
__global__ void __launch_bounds__(256) pose_compose_kernel(int n, float* __restrict__ R, float* __restrict__ t,
                                                           const float* __restrict__ dR,
                                                           const float* __restrict__ dt,
                                                           const float* __restrict__ points_in,
                                                           float* __restrict__ points_out_cm,
                                                           int64_t out_batch_stride,
                                                           unsigned char* __restrict__ points_out_pm, int update,
                                                           int pm_fmt) {
    __shared__ float sR[9], sT[3];
    const int bs = blockIdx.y;
    if (threadIdx.x < 12) {
        const float* Rb = R + (size_t)bs * 9;
        const float* tb = t + (size_t)bs * 3;
        if (update) {
            const float* dRb = dR + (size_t)bs * 9;
            const float* dtb = dt + (size_t)bs * 3;
            if (threadIdx.x < 9) {
                const int r = threadIdx.x / 3, c = threadIdx.x % 3;
                sR[threadIdx.x] = __fmaf_rn(Rb[r * 3 + 2], dRb[6 + c],
                                            __fmaf_rn(Rb[r * 3 + 1], dRb[3 + c], __fmul_rn(Rb[r * 3 + 0], dRb[c])));
            } else {
                const int r = threadIdx.x - 9;
                const float rd = __fmaf_rn(Rb[r * 3 + 2], dtb[2],
                                           __fmaf_rn(Rb[r * 3 + 1], dtb[1], __fmul_rn(Rb[r * 3 + 0], dtb[0])));
                sT[r] = __fadd_rn(rd, tb[r]);
            }
        } else {
            if (threadIdx.x < 9) sR[threadIdx.x] = Rb[threadIdx.x];
            else sT[threadIdx.x - 9] = tb[threadIdx.x - 9];
        }
    }
    __syncthreads();
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i < n) {
        const float* p = points_in + ((size_t)bs * n + i) * 3;
        const float x = __fsub_rn(p[0], sT[0]), y = __fsub_rn(p[1], sT[1]), z = __fsub_rn(p[2], sT[2]);
        float v[3];
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) v[ch] = __fmaf_rn(z, sR[6 + ch], __fmaf_rn(y, sR[3 + ch], __fmul_rn(x, sR[ch])));
        if (points_out_cm != nullptr) {
            float* o = points_out_cm + (size_t)bs * out_batch_stride + i;
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) o[(size_t)ch * n] = v[ch];
        }
        if (points_out_pm != nullptr && pm_fmt == 1) {
            // PM16 image with 32 channels per row: channels 0-2 = fp16(xyz), 3-5 = fp16(xyz - fp16(xyz)); the caller
            // zeroed the image once, only the first 8-channel unit of the row is rewritten
            const size_t r = (size_t)bs * n + i;
            unsigned char* d = points_out_pm + (r >> 7) * 8192 + ((r & 127) >> 3) * 512 + (r & 7) * 16;
            __half hi[3], lo[3];
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) {
                hi[ch] = __float2half_rn(fminf(fmaxf(v[ch], -65504.f), 65504.f));
                lo[ch] = __float2half_rn(v[ch] - __half2float(hi[ch]));
            }
            const __half2 w0 = __halves2half2(hi[0], hi[1]), w1 = __halves2half2(hi[2], lo[0]),
                          w2 = __halves2half2(lo[1], lo[2]);
            *reinterpret_cast<uint4*>(d) = make_uint4(*reinterpret_cast<const uint32_t*>(&w0),
                                                      *reinterpret_cast<const uint32_t*>(&w1),
                                                      *reinterpret_cast<const uint32_t*>(&w2), 0u);
        } else if (points_out_pm != nullptr) {
            // point-major bf16 hi / lo image with 32 channels per row (pm_gemm.cu), channels 0-2 = xyz; the caller
            // zeroed the image once, only the first 8-channel unit of the row is rewritten
            const size_t r = (size_t)bs * n + i;
            unsigned char* d = points_out_pm + (r >> 7) * 16384 + ((r & 127) >> 3) * 512 + (r & 7) * 16;
            uint32_t h[2], l[2];
            {
                const __nv_bfloat162 hh = __floats2bfloat162_rn(v[0], v[1]);
                h[0] = *reinterpret_cast<const uint32_t*>(&hh);
                const __nv_bfloat162 ll = __floats2bfloat162_rn(v[0] - __uint_as_float(h[0] << 16),
                                                                v[1] - __uint_as_float(h[0] & 0xffff0000u));
                l[0] = *reinterpret_cast<const uint32_t*>(&ll);
                const __nv_bfloat162 h2 = __floats2bfloat162_rn(v[2], 0.f);
                h[1] = *reinterpret_cast<const uint32_t*>(&h2);
                const __nv_bfloat162 l2 = __floats2bfloat162_rn(v[2] - __uint_as_float(h[1] << 16), 0.f);
                l[1] = *reinterpret_cast<const uint32_t*>(&l2);
            }
            *reinterpret_cast<uint4*>(d) = make_uint4(h[0], h[1], 0u, 0u);
            *reinterpret_cast<uint4*>(d + 8192) = make_uint4(l[0], l[1], 0u, 0u);
        }
    }
    // every CTA of this batch item has read R,t above; only CTA 0 commits the update, after the
    // grid-wide read is guaranteed by doing it in a second launch (see host code).
}

__global__ void pose_commit_kernel(int b, float* __restrict__ R, float* __restrict__ t,
                                   const float* __restrict__ dR, const float* __restrict__ dt) {
    const int bs = blockIdx.x * blockDim.x + threadIdx.x;
    if (bs >= b) return;
    float Rb[9], tb[3], Rn[9], tn[3];
#pragma unroll
    for (int k = 0; k < 9; ++k) Rb[k] = R[(size_t)bs * 9 + k];
#pragma unroll
    for (int k = 0; k < 3; ++k) tb[k] = t[(size_t)bs * 3 + k];
    const float* dRb = dR + (size_t)bs * 9;
    const float* dtb = dt + (size_t)bs * 3;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
#pragma unroll
        for (int c = 0; c < 3; ++c)
            Rn[r * 3 + c] = __fmaf_rn(Rb[r * 3 + 2], dRb[6 + c],
                                      __fmaf_rn(Rb[r * 3 + 1], dRb[3 + c], __fmul_rn(Rb[r * 3 + 0], dRb[c])));
        const float rd =
            __fmaf_rn(Rb[r * 3 + 2], dtb[2], __fmaf_rn(Rb[r * 3 + 1], dtb[1], __fmul_rn(Rb[r * 3 + 0], dtb[0])));
        tn[r] = __fadd_rn(rd, tb[r]);
    }
#pragma unroll
    for (int k = 0; k < 9; ++k) R[(size_t)bs * 9 + k] = Rn[k];
#pragma unroll
    for (int k = 0; k < 3; ++k) t[(size_t)bs * 3 + k] = tn[k];
}

}  // namespace

DCL_API int dcl_svd3_project(int b, const float* in9, int normalize_columns, float* R, void* stream) {
    DCL_RETURN_IF_BAD(b >= 0);
    if (b == 0) return 0;
    svd3_project_kernel<<<DCL_DIVUP(b, 128), 128, 0, (cudaStream_t)stream>>>(b, in9, normalize_columns, R);
    return dcl_launch_status();
}

DCL_API int dcl_weighted_kabsch(int b, int n, const float* src, const float* dst, const float* w, float* R, float* t,
                                void* stream) {
    DCL_RETURN_IF_BAD(b >= 0 && n >= 0);
    if (b == 0) return 0;
    weighted_kabsch_kernel<<<DCL_DIVUP(b, 4), 128, 0, (cudaStream_t)stream>>>(b, n, src, dst, w, R, t);
    return dcl_launch_status();
}

DCL_API int dcl_pose_compose(int b, int n, float* R, float* t, const float* dR, const float* dt,
                             const float* points_in, float* points_out_cm, int64_t out_batch_stride, void* stream) {
    return dcl_pose_compose_pm(b, n, R, t, dR, dt, points_in, points_out_cm, out_batch_stride, nullptr, stream);
}

static int pose_compose_any(int b, int n, float* R, float* t, const float* dR, const float* dt, const float* points_in,
                            float* points_out_cm, int64_t out_batch_stride, void* points_out_pm, int pm_fmt,
                            void* stream);

DCL_API int dcl_pose_compose_pm(int b, int n, float* R, float* t, const float* dR, const float* dt,
                                const float* points_in, float* points_out_cm, int64_t out_batch_stride,
                                void* points_out_pm, void* stream) {
    return pose_compose_any(b, n, R, t, dR, dt, points_in, points_out_cm, out_batch_stride, points_out_pm, 0, stream);
}

DCL_API int dcl_pose_compose_pm16(int b, int n, float* R, float* t, const float* dR, const float* dt,
                                  const float* points_in, float* points_out_cm, int64_t out_batch_stride,
                                  void* points_out_pm16, void* stream) {
    return pose_compose_any(b, n, R, t, dR, dt, points_in, points_out_cm, out_batch_stride, points_out_pm16, 1, stream);
}

static int pose_compose_any(int b, int n, float* R, float* t, const float* dR, const float* dt, const float* points_in,
                            float* points_out_cm, int64_t out_batch_stride, void* points_out_pm, int pm_fmt,
                            void* stream) {
    DCL_RETURN_IF_BAD(b >= 0 && n >= 0);
    DCL_RETURN_IF_BAD(points_out_pm == nullptr || (((long)b * n) % 128 == 0 && ((uintptr_t)points_out_pm & 15u) == 0));
    if (b == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const int update = (dR != nullptr && dt != nullptr) ? 1 : 0;
    int launched = 0;
    if (n > 0 && points_in != nullptr && (points_out_cm != nullptr || points_out_pm != nullptr)) {
        ++launched;
        dim3 grid(DCL_DIVUP(n, 256), b);
        pose_compose_kernel<<<grid, 256, 0, st>>>(n, R, t, dR, dt, points_in, points_out_cm, out_batch_stride,
                                                  reinterpret_cast<unsigned char*>(points_out_pm), update, pm_fmt);
    }
    if (update) {
        ++launched;
        pose_commit_kernel<<<DCL_DIVUP(b, 128), 128, 0, st>>>(b, R, t, dR, dt);
    }
    return dcl_launch_status(launched);
}
