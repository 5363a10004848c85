// Pose regressors on the pooled feature: regressor_rot / regressor_trans of models/DCL_Net.py:139-151,230-235
// (and regressor_rot2 / regressor_trans2 of models/refiner.py:66-77): Head_MultiLayerPerceptron
// [d_in -> d_h1 -> d_h2 -> d_out], Conv1d(k=1) + ReLU on a (b, d_in, 1) tensor, i.e. three tiny dense layers per
// instance.  The reference runs each as a cuDNN convolution with a single output position (~50 us apiece on B200);
// here one CTA per (instance, head) keeps the activations in shared memory and streams the fp32 weights once:
// warp per output row, lanes stride the input with 128-bit loads, shuffle reduction.  fp32 FMA throughout.
#include "common.cuh"
#include "../../include/dcl_b200.h"

namespace {

constexpr int PH_THREADS = 256;
constexpr int PH_MAX_IN = 1024;

__device__ __forceinline__ float warp_dot(const float* __restrict__ w, const float* x, int n, int lane) {
    float acc = 0.f;
    if ((n & 127) == 0) {
        const float4* w4 = reinterpret_cast<const float4*>(w);
        const float4* x4 = reinterpret_cast<const float4*>(x);
        for (int i = lane; i < (n >> 2); i += 32) {
            const float4 a = __ldg(w4 + i), b = x4[i];
            acc = __fmaf_rn(a.x, b.x, acc);
            acc = __fmaf_rn(a.y, b.y, acc);
            acc = __fmaf_rn(a.z, b.z, acc);
            acc = __fmaf_rn(a.w, b.w, acc);
        }
    } else {
        for (int i = lane; i < n; i += 32) acc = __fmaf_rn(__ldg(w + i), x[i], acc);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    return acc;
}

__global__ void __launch_bounds__(PH_THREADS) pose_head_kernel(dcl_pose_head_mlp h0, dcl_pose_head_mlp h1,
                                                               const float* __restrict__ pooled,
                                                               float* __restrict__ out0, float* __restrict__ out1) {
    __shared__ __align__(16) float s_x[PH_MAX_IN];
    __shared__ __align__(16) float s_a[PH_MAX_IN];
    __shared__ __align__(16) float s_b[PH_MAX_IN];
    const dcl_pose_head_mlp& h = blockIdx.y == 0 ? h0 : h1;
    float* out = blockIdx.y == 0 ? out0 : out1;
    const int inst = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < h.d_in; i += PH_THREADS) s_x[i] = pooled[(size_t)inst * h.d_in + i];
    __syncthreads();
    for (int o = warp; o < h.d_h1; o += PH_THREADS / 32) {
        const float v = warp_dot(h.w1 + (size_t)o * h.d_in, s_x, h.d_in, lane);
        if (lane == 0) s_a[o] = fmaxf(v + h.b1[o], 0.f);
    }
    __syncthreads();
    for (int o = warp; o < h.d_h2; o += PH_THREADS / 32) {
        const float v = warp_dot(h.w2 + (size_t)o * h.d_h1, s_a, h.d_h1, lane);
        if (lane == 0) s_b[o] = fmaxf(v + h.b2[o], 0.f);
    }
    __syncthreads();
    for (int o = warp; o < h.d_out; o += PH_THREADS / 32) {
        const float v = warp_dot(h.w3 + (size_t)o * h.d_h2, s_b, h.d_h2, lane);
        if (lane == 0) out[(size_t)inst * h.d_out + o] = v + h.b3[o];
    }
}

bool head_ok(const dcl_pose_head_mlp& h) {
    return h.w1 && h.b1 && h.w2 && h.b2 && h.w3 && h.b3 && h.d_in > 0 && h.d_in <= PH_MAX_IN && h.d_h1 > 0 &&
           h.d_h1 <= PH_MAX_IN && h.d_h2 > 0 && h.d_h2 <= PH_MAX_IN && h.d_out > 0;
}

}  // namespace

DCL_API int dcl_pose_head(int b, const float* pooled, const dcl_pose_head_mlp* rot_head,
                          const dcl_pose_head_mlp* trans_head, float* out_rot, float* out_trans, void* stream) {
    DCL_RETURN_IF_BAD(b >= 0 && pooled != nullptr && rot_head != nullptr && trans_head != nullptr);
    DCL_RETURN_IF_BAD(head_ok(*rot_head) && head_ok(*trans_head) && rot_head->d_in == trans_head->d_in);
    DCL_RETURN_IF_BAD(out_rot != nullptr && out_trans != nullptr && (((uintptr_t)pooled) & 15u) == 0);
    if (b == 0) return 0;
    dim3 grid(b, 2);
    pose_head_kernel<<<grid, PH_THREADS, 0, (cudaStream_t)stream>>>(*rot_head, *trans_head, pooled, out_rot, out_trans);
    return dcl_launch_status();
}
