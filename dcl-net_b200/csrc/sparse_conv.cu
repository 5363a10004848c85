// Sparse 3x3x3 convolution of the two towers (SparseConv3d / SubMConv3d of Backbone_SPCONV, models/Modules.py:100-159)
// on tcgen05, output-stationary.
//
// Reference (libs/spconv/include/spconv/spconv_ops.h:253-349): for each of the 27 kernel offsets, gather the input
// rows of that offset's pairs into a buffer, torch::mm with W[k] (fp32 cuBLAS), scatter-add into the output —
// 27 x (gather kernel + GEMM + atomic scatter) launches per layer, 16 layers per forward.
//
// Here: one launch per layer for both towers.  A CTA owns up to four tiles of 128 OUTPUT rows and one n-tile of the
// output channels; its accumulators stay in TMEM (4 x NT fp32 columns) while it walks the 27 offsets, so there is no
// scatter and no atomics: out[r] = sum_k in[nbr[r,k]] W[k] with the rulebook row nbr[r, 0..26] (sparse_index.cu).
// The reduction axis is the "virtual channel" kv = k*CIN + c (27*CIN, padded to a multiple of 64): a pipeline stage
// is 64 virtual channels — 4 offsets at CIN = 16, half an offset at CIN = 128 —
//     A stage  128 rows x 64 fp16 (16 KB): gathered straight from global memory into the UMMA K-major core-matrix
//              layout by cp.async (16 bytes = 8 channels of one input row per copy, zero-filled for absent voxels),
//     W stage  [hi | lo] x NT x 64 fp16: one TMA bulk copy, shared by the CTA's four row tiles,
// and 4 K-steps x 2 MMAs (activations rounded once to fp16, weights as fp16 hi + lo: the format of the whole
// inference path).  Stages whose offsets no row of the CTA uses (anymask) are skipped.  Epilogue: + shift (folded
// BatchNorm1d) -> ReLU -> fp16 operand rows of the next layer and / or fp32 rows.
//
// Warps: 0-3 gather (one thread per output row) and afterwards run the epilogue (TMEM lane = row), 4 issues the
// MMAs, 5 streams the weights.
#include "common.cuh"
#include "umma.cuh"
#include "../../include/dcl_b200.h"
#include <cuda_fp16.h>

namespace {

constexpr int SC_BM = 128;
constexpr int SC_KC = 64;            // virtual channels per stage
constexpr int SC_MT = 4;             // row tiles per CTA
constexpr int SC_NA = 6;             // A-stage ring
constexpr int SC_NW = 3;             // W-stage ring
constexpr int SC_THREADS = 192;
constexpr int SC_A_BYTES = SC_BM * SC_KC * 2;   // 16384
constexpr int SC_LOOKAHEAD = 4;      // gather stages in flight per thread (< SC_NA)

template <int NT>
struct ScCfg {
    static constexpr int W_HALF = NT * SC_KC * 2;
    static constexpr int W_BYTES = 2 * W_HALF;
    static constexpr int OFF_W = SC_NA * SC_A_BYTES;
    static constexpr int OFF_BAR = OFF_W + SC_NW * W_BYTES;
    static constexpr int SMEM_BYTES = OFF_BAR + 256;
    static constexpr int TMEM_COLS = (SC_MT * NT) < 32 ? 32 : (SC_MT * NT);   // 64, 128, 256, 512: powers of two
    static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
    static_assert(TMEM_COLS <= 512, "TMEM budget");
};

struct ScTower {
    const __half* in16;
    const int* nbr;
    const unsigned int* anymask;
    const int* total_ptr;      // &offsets[s_out][B]
    const unsigned char* w;    // packed weights
    const float* shift;
    __half* out16;
    float* out32;
    int cap_out;
};
struct ScArgs {
    int cin, cout, nstages;    // nstages = ceil(27*cin / 64)
    ScTower tw[2];
};

__device__ __forceinline__ void sc_cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void sc_cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void sc_cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// bit mask over stages: stage s is needed when any of the kernel offsets it covers is used by a row of this CTA
__device__ __forceinline__ unsigned long long sc_stage_mask(unsigned int kmask, int cin, int nstages) {
    unsigned long long m = 0ull;
    for (int s = 0; s < nstages; ++s) {
        const int k_lo = (s * SC_KC) / cin, k_hi = min(26, (s * SC_KC + SC_KC - 1) / cin);
        unsigned int span = 0;
        for (int k = k_lo; k <= k_hi; ++k) span |= 1u << k;
        if (k_lo <= 26 && (kmask & span)) m |= 1ull << s;
    }
    return m;
}

template <int NT>
__global__ void __launch_bounds__(SC_THREADS, 1) sparse_conv3_kernel(const __grid_constant__ ScArgs args) {
    using Cfg = ScCfg<NT>;
    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);
    uint64_t* a_full = bars;                 // [SC_NA] 128 gather threads
    uint64_t* a_empty = a_full + SC_NA;      // [SC_NA] one commit
    uint64_t* w_full = a_empty + SC_NA;      // [SC_NW] TMA bytes
    uint64_t* w_empty = w_full + SC_NW;      // [SC_NW] one commit
    uint64_t* acc_full = w_empty + SC_NW;    // [1]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

    const ScTower& tw = args.tw[blockIdx.z];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int total = min(*tw.total_ptr, tw.cap_out);
    const int ntiles = (total + SC_BM - 1) / SC_BM;
    const int tile0 = blockIdx.x * SC_MT;
    if (tile0 >= ntiles) return;                       // uniform for the CTA: nothing was allocated yet
    const int nt_mine = min(SC_MT, ntiles - tile0);
    const int nti = blockIdx.y;
    const int cin = args.cin, cout = args.cout, nstages = args.nstages;

    unsigned int kmask = 0;
    for (int t = 0; t < nt_mine; ++t) kmask |= tw.anymask[tile0 + t];
    const unsigned long long smask = sc_stage_mask(kmask, cin, nstages);

    if (threadIdx.x == 0) {
        for (int i = 0; i < SC_NA; ++i) {
            dcl_mbar_init(a_full + i, 128);
            dcl_mbar_init(a_empty + i, 1);
        }
        for (int i = 0; i < SC_NW; ++i) {
            dcl_mbar_init(w_full + i, 1);
            dcl_mbar_init(w_empty + i, 1);
        }
        dcl_mbar_init(acc_full, 1);
        dcl_fence_barrier_init();
    }
    if (warp == 4) tc_alloc(tmem_slot, Cfg::TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < 4) {
        // ===================== gather: thread = output row of each tile =====================
        const int row = threadIdx.x;
        const uint32_t a_row = (uint32_t)(row >> 3) * (SC_KC / 8) * 128u + (uint32_t)(row & 7) * 16u;
        const uint32_t sA = dcl_smem_u32(smem);
        int issued = 0, signalled = 0;
        int pend_buf[SC_LOOKAHEAD + 1];
#pragma unroll
        for (int i = 0; i <= SC_LOOKAHEAD; ++i) pend_buf[i] = 0;
        for (int s = 0; s < nstages; ++s) {
            if (!((smask >> s) & 1ull)) continue;
            for (int t = 0; t < nt_mine; ++t) {
                const int buf = issued % SC_NA;
                if (issued >= SC_NA) dcl_mbar_wait(a_empty + buf, (uint32_t)(((issued / SC_NA) - 1) & 1));
                const int r = (tile0 + t) * SC_BM + row;
                const int* nb = tw.nbr + (size_t)r * 32;
                const uint32_t dst = sA + buf * SC_A_BYTES + a_row;
#pragma unroll
                for (int j = 0; j < SC_KC / 8; ++j) {
                    const int kv = s * SC_KC + j * 8;
                    const int k = kv / cin, c0 = kv - k * cin;
                    int src_row = -1;
                    if (k < 27 && r < total) src_row = __ldg(nb + k);
                    const __half* src = tw.in16 + (size_t)(src_row < 0 ? 0 : src_row) * cin + c0;
                    sc_cp_async16(dst + j * 128, src, src_row < 0 ? 0u : 16u);
                }
                sc_cp_commit();
                pend_buf[issued % (SC_LOOKAHEAD + 1)] = buf;
                ++issued;
                if (issued - signalled > SC_LOOKAHEAD) {
                    sc_cp_wait<SC_LOOKAHEAD>();          // the oldest outstanding stage of this thread has landed
                    dcl_fence_proxy_async();             // generic-proxy writes -> visible to the tensor core
                    dcl_mbar_arrive(a_full + pend_buf[signalled % (SC_LOOKAHEAD + 1)]);
                    ++signalled;
                }
            }
        }
        sc_cp_wait<0>();
        dcl_fence_proxy_async();
        while (signalled < issued) {
            dcl_mbar_arrive(a_full + pend_buf[signalled % (SC_LOOKAHEAD + 1)]);
            ++signalled;
        }
        // ===================== epilogue: TMEM lane = output row =====================
        dcl_mbar_wait(acc_full, 0);
        tc_fence_after();
        const uint32_t t_lane = (uint32_t)(warp * 32) << 16;
        for (int t = 0; t < nt_mine; ++t) {
            const int r = (tile0 + t) * SC_BM + row;
#pragma unroll 1
            for (int cc = 0; cc < NT / 16; ++cc) {
                uint32_t v[16];
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                    "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                    : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                      "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                    : "r"(tmem_base + t_lane + t * NT + cc * 16)
                    : "memory");
                tc_wait_ld();
                if (r < total) {
                    const int col0 = nti * NT + cc * 16;
                    float y[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i)
                        y[i] = fmaxf(__uint_as_float(v[i]) + __ldg(tw.shift + col0 + i), 0.f);
                    if (tw.out32 != nullptr) {
                        float4* o = reinterpret_cast<float4*>(tw.out32 + (size_t)r * cout + col0);
#pragma unroll
                        for (int q = 0; q < 4; ++q) o[q] = make_float4(y[4 * q], y[4 * q + 1], y[4 * q + 2], y[4 * q + 3]);
                    }
                    if (tw.out16 != nullptr) {
                        uint32_t h[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            const __half2 hh = __floats2half2_rn(fminf(y[2 * e], 65504.f), fminf(y[2 * e + 1], 65504.f));
                            h[e] = *reinterpret_cast<const uint32_t*>(&hh);
                        }
                        uint4* o = reinterpret_cast<uint4*>(tw.out16 + (size_t)r * cout + col0);
                        o[0] = make_uint4(h[0], h[1], h[2], h[3]);
                        o[1] = make_uint4(h[4], h[5], h[6], h[7]);
                    }
                }
            }
        }
        tc_fence_before();
    } else if (warp == 4) {
        // ===================== MMA issuer =====================
        if (dcl_elect_one()) {
            constexpr uint32_t idesc = (1u << 4) | ((uint32_t)(NT >> 3) << 17) | ((uint32_t)(SC_BM >> 4) << 24);  // f16 x f16 -> f32
            constexpr uint32_t SBO = (SC_KC / 8) * 128;
            const uint64_t dA0 = umma_desc(dcl_smem_u32(smem), 128, SBO);
            const uint64_t dW0 = umma_desc(dcl_smem_u32(smem + Cfg::OFF_W), 128, SBO);
            int ai = 0, wi = 0;
            bool first = true;
            for (int s = 0; s < nstages; ++s) {
                if (!((smask >> s) & 1ull)) continue;
                const int wb = wi % SC_NW;
                dcl_mbar_wait(w_full + wb, (uint32_t)((wi / SC_NW) & 1));
                const uint64_t dWh = dW0 + (uint64_t)((wb * Cfg::W_BYTES) >> 4);
                const uint64_t dWl = dWh + (uint64_t)(Cfg::W_HALF >> 4);
                for (int t = 0; t < nt_mine; ++t, ++ai) {
                    const int ab = ai % SC_NA;
                    dcl_mbar_wait(a_full + ab, (uint32_t)((ai / SC_NA) & 1));
                    tc_fence_after();
                    const uint64_t dA = dA0 + (uint64_t)((ab * SC_A_BYTES) >> 4);
                    const uint32_t tacc = tmem_base + t * NT;
#pragma unroll
                    for (int ks = 0; ks < SC_KC / 16; ++ks) {
                        const uint64_t off = (uint64_t)((ks * 256) >> 4);
                        tc_mma_bf16(tacc, dA + off, dWh + off, idesc, (first && ks == 0) ? 0u : 1u);
                        tc_mma_bf16(tacc, dA + off, dWl + off, idesc, 1u);
                    }
                    tc_commit(a_empty + ab);
                }
                first = false;
                tc_commit(w_empty + wb);
                ++wi;
            }
            tc_commit(acc_full);
        }
    } else {
        // ===================== weight stream =====================
        if (dcl_elect_one()) {
            const unsigned char* w = tw.w + (size_t)nti * nstages * Cfg::W_BYTES;
            int wi = 0;
            for (int s = 0; s < nstages; ++s) {
                if (!((smask >> s) & 1ull)) continue;
                const int wb = wi % SC_NW;
                if (wi >= SC_NW) dcl_mbar_wait(w_empty + wb, (uint32_t)(((wi / SC_NW) - 1) & 1));
                dcl_mbar_arrive_expect_tx(w_full + wb, Cfg::W_BYTES);
                dcl_bulk_g2s(smem + Cfg::OFF_W + wb * Cfg::W_BYTES, w + (size_t)s * Cfg::W_BYTES, Cfg::W_BYTES, w_full + wb);
                ++wi;
            }
        }
    }
    __syncwarp();
    __syncthreads();
    if (warp == 4) {
        tc_fence_after();
        tc_dealloc(tmem_base, Cfg::TMEM_COLS);
    }
}

template <int NT>
int sc_launch(const ScArgs& args, int ntowers, int max_cap, cudaStream_t st) {
    using Cfg = ScCfg<NT>;
    cudaError_t e = cudaFuncSetAttribute(sparse_conv3_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
    const int groups = DCL_DIVUP(max_cap / SC_BM, SC_MT);
    dim3 grid(groups, args.cout / NT, ntowers);
    sparse_conv3_kernel<NT><<<grid, SC_THREADS, Cfg::SMEM_BYTES, st>>>(args);
    return dcl_launch_status();
}

}  // namespace

DCL_API int dcl_spb_conv3(int b, int cin_pad, int cout, int ntowers, const dcl_spb_conv* convs, void* stream) {
    DCL_RETURN_IF_BAD(b > 0 && ntowers >= 1 && ntowers <= 2 && convs != nullptr);
    DCL_RETURN_IF_BAD(cin_pad == 16 || cin_pad == 32 || cin_pad == 64 || cin_pad == 128);
    DCL_RETURN_IF_BAD(cout == 16 || cout == 32 || cout == 64 || cout == 128 || cout == 256);
    ScArgs args;
    args.cin = cin_pad;
    args.cout = cout;
    args.nstages = DCL_DIVUP(27 * cin_pad, SC_KC);
    int max_cap = 0;
    for (int t = 0; t < ntowers; ++t) {
        const dcl_spb_conv& c = convs[t];
        DCL_RETURN_IF_BAD(c.in16 != nullptr && c.nbr != nullptr && c.anymask != nullptr && c.offsets_out != nullptr &&
                          c.w != nullptr && c.shift != nullptr && (c.out16 != nullptr || c.out32 != nullptr));
        DCL_RETURN_IF_BAD(c.cap_out > 0 && c.cap_out % SC_BM == 0);
        DCL_RETURN_IF_BAD(((((uintptr_t)c.in16) | ((uintptr_t)c.w) | ((uintptr_t)c.out16) | ((uintptr_t)c.out32)) & 15u) == 0);
        args.tw[t] = {reinterpret_cast<const __half*>(c.in16), c.nbr, c.anymask, c.offsets_out + b,
                      reinterpret_cast<const unsigned char*>(c.w), c.shift, reinterpret_cast<__half*>(c.out16), c.out32,
                      c.cap_out};
        if (c.cap_out > max_cap) max_cap = c.cap_out;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const int nt = cout < 128 ? cout : 128;
    if (nt == 16) return sc_launch<16>(args, ntowers, max_cap, st);
    if (nt == 32) return sc_launch<32>(args, ntowers, max_cap, st);
    if (nt == 64) return sc_launch<64>(args, ntowers, max_cap, st);
    return sc_launch<128>(args, ntowers, max_cap, st);
}
