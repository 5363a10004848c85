// Pose regressors on the pooled feature: regressor_rot / regressor_trans of models/DCL_Net.py:139-151,230-235
// (and regressor_rot2 / regressor_trans2 of models/refiner.py:66-77): Head_MultiLayerPerceptron
// [d_in -> d_h1 -> d_h2 -> d_out], Conv1d(k=1) + ReLU on a (b, d_in, 1) tensor, i.e. three small dense layers
// applied to every instance of the batch.  The reference runs each as a cuDNN convolution with a single output
// position (~50 us apiece on B200).  Here each layer of BOTH heads is one launch of a small-batch dense kernel:
// a CTA stages up to 32 instances' input vectors in shared memory once, each warp owns one output row whose
// fp32 weights it streams exactly once with 128-bit loads and applies to all staged instances (one accumulator
// per instance), finishing with a shuffle reduction.  Weights are read once per layer instead of once per
// instance.  fp32 FMA throughout.
#include "common.cuh"
#include "../../include/dcl_b200.h"
#include <math_constants.h>

namespace {

constexpr int PH_THREADS = 256;
constexpr int PH_ROWS = PH_THREADS / 32;  // output rows per CTA
constexpr int PH_INST = 32;               // instances per CTA
constexpr int PH_MAX_IN = 1024;

struct DenseJob {
    const float* x;     // (b, k) input of this head
    const float* w;     // (o, k)
    const float* bias;  // (o)
    float* y;           // (b, o)
    int k, o, relu;
};

// grid = (row chunks, heads, instance chunks)
__global__ void __launch_bounds__(PH_THREADS) dense_small_batch_kernel(DenseJob j0, DenseJob j1, int b) {
    extern __shared__ __align__(16) float s_x[];  // PH_INST x k
    const DenseJob& j = blockIdx.y == 0 ? j0 : j1;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row = blockIdx.x * PH_ROWS + warp;
    const int i0 = blockIdx.z * PH_INST;
    const int ni = min(PH_INST, b - i0);
    if (blockIdx.x * PH_ROWS >= j.o) return;  // this head has fewer rows than the other one
    const int k = j.k;
    // Stage the instances' input vectors: they are contiguous, so one bulk copy (TMA) brings all of them; a
    // per-thread load loop pays one L2 round trip per few elements (it was 3/4 of this kernel's time).
    __shared__ uint64_t s_bar;
    const float* xsrc = j.x + (size_t)i0 * k;
    const uint32_t xbytes = (uint32_t)ni * (uint32_t)k * 4u;
    const bool bulk = ((((uintptr_t)xsrc) & 15u) == 0) && ((xbytes & 15u) == 0) && xbytes < (1u << 20);
    if (bulk) {
        if (threadIdx.x == 0) {
            dcl_mbar_init(&s_bar, 1);
            dcl_fence_barrier_init();
            dcl_mbar_arrive_expect_tx(&s_bar, xbytes);
            dcl_bulk_g2s(s_x, xsrc, xbytes, &s_bar);
        }
    } else {
        for (int idx = threadIdx.x; idx < ni * k; idx += PH_THREADS) s_x[idx] = xsrc[idx];
    }
    // the weight row does not depend on the staged inputs: fetch it while they land
    const float* wr = j.w + (size_t)min(row, j.o - 1) * k;
    const bool vec = (k & 127) == 0 && k <= PH_MAX_IN;
    float4 wreg[PH_MAX_IN / 128];
    if (vec) {
        const float4* w4 = reinterpret_cast<const float4*>(wr);
#pragma unroll
        for (int t = 0; t < PH_MAX_IN / 128; ++t)
            wreg[t] = (t < (k >> 7)) ? __ldg(w4 + t * 32 + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncthreads();  // s_bar initialised / cooperative staging complete
    if (bulk) dcl_mbar_wait(&s_bar, 0);
    if (row >= j.o) return;
    float acc[PH_INST];
#pragma unroll
    for (int i = 0; i < PH_INST; ++i) acc[i] = 0.f;
    if (vec) {
#pragma unroll
        for (int t = 0; t < PH_MAX_IN / 128; ++t) {
            if (t < (k >> 7)) {
                const float4 a = wreg[t];
                const int c = t * 32 + lane;
#pragma unroll
                for (int i = 0; i < PH_INST; ++i) {
                    if (i < ni) {
                        const float4 xv = reinterpret_cast<const float4*>(s_x + (size_t)i * k)[c];
                        acc[i] = __fmaf_rn(a.x, xv.x, acc[i]);
                        acc[i] = __fmaf_rn(a.y, xv.y, acc[i]);
                        acc[i] = __fmaf_rn(a.z, xv.z, acc[i]);
                        acc[i] = __fmaf_rn(a.w, xv.w, acc[i]);
                    }
                }
            }
        }
    } else {
        for (int c = lane; c < k; c += 32) {
            const float a = __ldg(wr + c);
#pragma unroll
            for (int i = 0; i < PH_INST; ++i)
                if (i < ni) acc[i] = __fmaf_rn(a, s_x[(size_t)i * k + c], acc[i]);
        }
    }
    const float bias = j.bias[row];
#pragma unroll
    for (int i = 0; i < PH_INST; ++i) {
        float v = acc[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0 && i < ni) {
            v += bias;
            j.y[(size_t)(i0 + i) * j.o + row] = j.relu ? fmaxf(v, 0.f) : v;
        }
    }
}

// Confidence weights (models/DCL_Net.py:219-220): conf = sigmoid(cat([conf_1, conf_2], dim=2)), conf_softmax =
// softmax(conf, dim=2) over the 2n correspondences of an instance.  One CTA per instance; the logits come from the
// per-row dot epilogue of dcl_pm_gemm without their bias, which is added here.  Outputs: conf (b, 2n) and the two
// halves of conf_softmax as flat per-row weights (b*n each), the layout the pooled fuser epilogue reads.
constexpr int CW_THREADS = 256;
__global__ void __launch_bounds__(CW_THREADS) conf_weights_kernel(int n, const float* __restrict__ logit_1,
                                                                  const float* __restrict__ logit_2,
                                                                  const float* __restrict__ bias_1,
                                                                  const float* __restrict__ bias_2,
                                                                  float* __restrict__ conf, float* __restrict__ w1,
                                                                  float* __restrict__ w2) {
    __shared__ float s_red[CW_THREADS / 32];
    __shared__ float s_bcast;
    const int bi = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float b1 = bias_1[0], b2 = bias_2[0];
    const float* l1 = logit_1 + (size_t)bi * n;
    const float* l2 = logit_2 + (size_t)bi * n;
    float* crow = conf + (size_t)bi * 2 * n;
    // Up to CW_PER values per thread stay in registers between the three phases (2n <= CW_THREADS * CW_PER covers
    // n <= 2048); longer rows re-read what pass 1 wrote.
    constexpr int CW_PER = 16;
    const int total_n = 2 * n;
    const bool in_regs = total_n <= CW_THREADS * CW_PER;
    float cv[CW_PER];
    // pass 1: sigmoid, row maximum
    float mx = -CUDART_INF_F;
    if (in_regs) {
#pragma unroll
        for (int k = 0; k < CW_PER; ++k) {
            const int i = tid + k * CW_THREADS;
            float c = -CUDART_INF_F;
            if (i < total_n) {
                const float x = (i < n) ? __fadd_rn(__ldg(l1 + i), b1) : __fadd_rn(__ldg(l2 + i - n), b2);
                c = 1.0f / (1.0f + expf(-x));
                crow[i] = c;
            }
            cv[k] = c;
            mx = fmaxf(mx, c);
        }
    } else {
        for (int i = tid; i < total_n; i += CW_THREADS) {
            const float x = (i < n) ? __fadd_rn(l1[i], b1) : __fadd_rn(l2[i - n], b2);
            const float c = 1.0f / (1.0f + expf(-x));
            crow[i] = c;
            mx = fmaxf(mx, c);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane == 0) s_red[warp] = mx;
    __syncthreads();
    if (tid == 0) {
        float m = s_red[0];
        for (int w = 1; w < CW_THREADS / 32; ++w) m = fmaxf(m, s_red[w]);
        s_bcast = m;
    }
    __syncthreads();
    mx = s_bcast;
    // pass 2: sum of exp(c - max)
    float sum = 0.f;
    if (in_regs) {
#pragma unroll
        for (int k = 0; k < CW_PER; ++k) {
            cv[k] = expf(cv[k] - mx);          // exp(-inf) = 0 for the padding slots
            sum += cv[k];
        }
    } else {
        for (int i = tid; i < total_n; i += CW_THREADS) sum += expf(crow[i] - mx);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    __syncthreads();
    if (lane == 0) s_red[warp] = sum;
    __syncthreads();
    if (tid == 0) {
        float t = 0.f;
        for (int w = 0; w < CW_THREADS / 32; ++w) t += s_red[w];
        s_bcast = t;
    }
    __syncthreads();
    const float total = s_bcast;
    if (in_regs) {
#pragma unroll
        for (int k = 0; k < CW_PER; ++k) {
            const int i = tid + k * CW_THREADS;
            if (i < total_n) {
                const float w = cv[k] / total;
                if (i < n) w1[(size_t)bi * n + i] = w;
                else w2[(size_t)bi * n + i - n] = w;
            }
        }
    } else {
        for (int i = tid; i < total_n; i += CW_THREADS) {
            const float w = expf(crow[i] - mx) / total;
            if (i < n) w1[(size_t)bi * n + i] = w;
            else w2[(size_t)bi * n + i - n] = w;
        }
    }
}

bool head_ok(const dcl_pose_head_mlp& h) {
    return h.w1 && h.b1 && h.w2 && h.b2 && h.w3 && h.b3 && h.d_in > 0 && h.d_in <= PH_MAX_IN && h.d_h1 > 0 &&
           h.d_h1 <= PH_MAX_IN && h.d_h2 > 0 && h.d_h2 <= PH_MAX_IN && h.d_out > 0;
}

int launch_layer(const DenseJob& a, const DenseJob& c, int b, cudaStream_t st) {
    const int kmax = a.k > c.k ? a.k : c.k, omax = a.o > c.o ? a.o : c.o;
    const size_t smem = (size_t)PH_INST * kmax * sizeof(float);
    if (smem > 48 * 1024)
        cudaFuncSetAttribute(dense_small_batch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    dim3 grid(DCL_DIVUP(omax, PH_ROWS), 2, DCL_DIVUP(b, PH_INST));
    dense_small_batch_kernel<<<grid, PH_THREADS, smem, st>>>(a, c, b);
    return dcl_launch_status();
}

}  // namespace

DCL_API size_t dcl_pose_head_workspace_bytes(int b, const dcl_pose_head_mlp* rot_head,
                                             const dcl_pose_head_mlp* trans_head) {
    if (b < 0 || rot_head == nullptr || trans_head == nullptr) return 0;
    return (size_t)b * (rot_head->d_h1 + rot_head->d_h2 + trans_head->d_h1 + trans_head->d_h2) * sizeof(float) + 64;
}

DCL_API int dcl_pose_head(int b, const float* pooled, const dcl_pose_head_mlp* rot_head,
                          const dcl_pose_head_mlp* trans_head, float* out_rot, float* out_trans, void* workspace,
                          size_t workspace_bytes, void* stream) {
    DCL_RETURN_IF_BAD(b >= 0 && pooled != nullptr && rot_head != nullptr && trans_head != nullptr);
    DCL_RETURN_IF_BAD(head_ok(*rot_head) && head_ok(*trans_head) && rot_head->d_in == trans_head->d_in);
    DCL_RETURN_IF_BAD(out_rot != nullptr && out_trans != nullptr && (((uintptr_t)pooled) & 15u) == 0);
    DCL_RETURN_IF_BAD(workspace != nullptr && (((uintptr_t)workspace) & 15u) == 0 &&
                      workspace_bytes >= dcl_pose_head_workspace_bytes(b, rot_head, trans_head));
    if (b == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const dcl_pose_head_mlp& r = *rot_head;
    const dcl_pose_head_mlp& t = *trans_head;
    float* r1 = reinterpret_cast<float*>(workspace);
    float* r2 = r1 + (size_t)b * r.d_h1;
    float* t1 = r2 + (size_t)b * r.d_h2;
    float* t2 = t1 + (size_t)b * t.d_h1;
    int e = launch_layer({pooled, r.w1, r.b1, r1, r.d_in, r.d_h1, 1}, {pooled, t.w1, t.b1, t1, t.d_in, t.d_h1, 1}, b, st);
    if (e) return e;
    e = launch_layer({r1, r.w2, r.b2, r2, r.d_h1, r.d_h2, 1}, {t1, t.w2, t.b2, t2, t.d_h1, t.d_h2, 1}, b, st);
    if (e) return e;
    return launch_layer({r2, r.w3, r.b3, out_rot, r.d_h2, r.d_out, 0}, {t2, t.w3, t.b3, out_trans, t.d_h2, t.d_out, 0}, b,
                        st);
}

DCL_API int dcl_conf_weights(int b, int n, const float* logit_1, const float* logit_2, const float* bias_1,
                             const float* bias_2, float* conf, float* w1, float* w2, void* stream) {
    DCL_RETURN_IF_BAD(b >= 0 && n > 0 && logit_1 && logit_2 && bias_1 && bias_2 && conf && w1 && w2);
    if (b == 0) return 0;
    conf_weights_kernel<<<b, CW_THREADS, 0, (cudaStream_t)stream>>>(n, logit_1, logit_2, bias_1, bias_2, conf, w1, w2);
    return dcl_launch_status();
}
