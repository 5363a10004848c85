// Point-major tensor-core GEMM for the pointwise MLP stacks of the FDA section
// (models/DCL_Net.py:56-151: the disengage Conv3d1x1+BN+ReLU stacks; models/Modules.py:173-201:
//  Head_MultiLayerPerceptron = Conv1d(k=1) -> ReLU -> [BatchNorm1d]).  A 1x1 convolution over points is
//      Y[r, o] = act( sum_i X[r, i] * W[o, i] + bias[o] )        r = point (row), i/o = channels
// The reference runs them as fp32 cuDNN/cuBLAS SIMT kernels; they are ~95 % of the stage-1 step.
//
// Here they run on tcgen05 with the same bf16 hi/lo operand split as the FDA kernel (3 MMAs per product,
// fp32 accumulation in TMEM => fp32-faithful results) and a fused epilogue
//      bias -> ReLU -> optional per-channel affine (eval-mode BatchNorm after the ReLU)
// that emits, as requested per problem:
//   * the next layer's operand directly ("PM image", below) — activations never round-trip through fp32,
//   * an fp32 channel-major (B, C, N) tensor (what the FDA kernel / the caller consume),
//   * per-warp partial sums of  row_weight[r] * Y[r, :]  (the confidence-weighted pooling of
//     models/DCL_Net.py:228), reduced later in a fixed order => deterministic,
//   * a per-row dot product  sum_o Y[r,o] * dot_w[o]  (the final 128 -> 1 layer of the confidence heads).
// Eval-mode BN *before* the ReLU (the disengage blocks) is folded into W and bias on the host.
//
// PM image of an (R x C) activation, R % 128 == 0, C % 32 == 0: blobs of (128 rows x 32 channels), each blob
// = bf16 hi image (8 KB) then bf16 lo image (8 KB), each image in the UMMA K-major no-swizzle layout
//   byte(r, c, half) = ((r/128)*(C/32) + c/32)*16384 + half*8192 + ((r%128)/8)*512 + ((c%32)/8)*128 + (r%8)*16 + (c%8)*2
// so one blob is one pipeline stage of the A operand, moved by ONE TMA bulk copy (LBO 128 B, SBO 512 B).
// Packed weights: per (n-tile of NT output channels, k-block of 32 input channels) one blob
//   [hi (NT x 32) | lo (NT x 32)], same layout with NT rows — one bulk copy per stage.
//
// Second operand format ("PM16", a_fmt / out_fmt == 1): activations stored ONCE-rounded in fp16 (2 bytes per element,
// blobs of 128 rows x 32 channels = 8 KB, same core-matrix layout), weights split into fp16 hi + lo: every product is
// X*Whi + X*Wlo — 2 MMAs instead of 3 and half the activation traffic.  What the rounding costs was measured against
// the tolerances of the path (tools/emulate_precision.py -> profiles/r02_precision_emulation_*.json; GPU parity
// tests): features ~4e-5 relative (bar 1e-3), poses ~1e-5 deg (bar 0.01 deg).  Values are clamped to the fp16 range.
//
// CTA = 128 rows x NT output channels; warp 0 TMA producer, warp 1 MMA issuer / TMEM owner, warps 2-5 epilogue.
// Two CTAs are resident per SM (TMEM NT <= 256 columns each), so one CTA's epilogue overlaps the other's MMAs.
#include "common.cuh"
#include "umma.cuh"
#include "../../include/dcl_b200.h"
#include <cuda_fp16.h>
#include <stdlib.h>

namespace {

constexpr int GM_BM = 128;
constexpr int GM_BK = 32;
constexpr int GM_THREADS = 192;    // simple kernel: TMA warp, MMA warp, 4 epilogue warps
constexpr int GM_P_EPI = 256;      // persistent kernel: 8 epilogue warps (two threads per row)
constexpr int GM_P_THREADS = 64 + GM_P_EPI;
constexpr int GM_A_BLOB = GM_BM * GM_BK * 4;  // 16384: bf16 hi image + bf16 lo image
constexpr int GM_A16_BLOB = GM_BM * GM_BK * 2;  // 8192: one fp16 image (PM16)
constexpr int GM_MAX_PROBLEMS = 8;
// operand formats (dcl_pm_gemm_problem.a_fmt / out_fmt)
constexpr int FMT_BF16X2 = 0, FMT_F16 = 1;
template <int FMT> struct GmFmt {
    static constexpr int A_BLOB = FMT == FMT_F16 ? GM_A16_BLOB : GM_A_BLOB;
};
__host__ __device__ constexpr uint32_t umma_idesc_f16(int M, int N, bool b_mn_major = false) {
    // kind::f16, A and B fp16 (format 0), fp32 accumulator
    return (1u << 4) | ((b_mn_major ? 1u : 0u) << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// two fp32 -> packed fp16x2, clamped to the finite fp16 range
__device__ __forceinline__ uint32_t pack_f16x2(float x0, float x1) {
    x0 = fminf(fmaxf(x0, -65504.f), 65504.f);
    x1 = fminf(fmaxf(x1, -65504.f), 65504.f);
    const __half2 h = __floats2half2_rn(x0, x1);
    return *reinterpret_cast<const uint32_t*>(&h);
}

struct PmGemmBatch {
    dcl_pm_gemm_problem p[GM_MAX_PROBLEMS];
};

template <int NT, int STAGES, int FMT = FMT_BF16X2>
struct GmCfg {
    static constexpr int A_BLOB = GmFmt<FMT>::A_BLOB;
    static constexpr int B_BLOB = NT * GM_BK * 4;
    static constexpr int STAGE_BYTES = A_BLOB + B_BLOB;
    static constexpr int OFF_BAR = STAGES * STAGE_BYTES;
    static constexpr int SMEM_BYTES = OFF_BAR + 128;
};

// Epilogue of one (128 x NT) tile: TMEM accumulator -> bias / ReLU / affine -> requested outputs.
// Called by the four epilogue warps (128 threads; quad = warp & 3 selects the TMEM lane quarter) after the
// accumulator is complete.  ncu on the first version showed ~40 instructions per output column and warp (per-element
// null checks, constant-bank loads with a dynamic problem index, scalar bf16 conversions) — the epilogue, not the
// MMAs, paced the kernel.  Now: the problem descriptor is copied to registers once, the per-column parameters
// (bias, post-scale, post-shift, dot vector; neutral values when absent) are staged in shared memory by the epilogue
// warps and read back with 128-bit broadcast loads, the per-element math is branch-free (FADD, FMNMX, FFMA) and the
// hi/lo split converts two values per instruction.
struct GmColParams {           // per output column of the tile, in shared memory
    float bias[256], scale[256], shift[256], dotw[256];
};

// EPI epilogue threads: 128 (one per row) or 256 (two per row: warps w and w+4 share a TMEM lane quadrant and split
// the tile's columns in halves — the layers with few k-blocks are paced by their epilogue, not by their MMAs).
template <int EPI>
__device__ __forceinline__ void gm_epi_barrier() { asm volatile("bar.sync 1, %0;" ::"n"(EPI) : "memory"); }

template <int NT, int EPI>
__device__ __forceinline__ void gm_epilogue_tile(const dcl_pm_gemm_problem& pr_in, int mt, int nti, uint32_t tmem_acc,
                                                 int ewarp, int lane, GmColParams& cp, float* s_dot, int inst = 0) {
    dcl_pm_gemm_problem pr = pr_in;  // registers, not repeated constant-bank loads with a dynamic index
    if (inst != 0)  // split-K slice `inst` of a strided-batch problem (weight-gradient launches): its own fp32 output
        pr.out_cm = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(pr.out_cm) + (size_t)inst * pr.out_cm_inst_stride);
    const int quad = ewarp & 3;            // ewarp = warp index - 2; (ewarp + 2) % 4 is the TMEM quadrant this warp may read
    const int half = ewarp >> 2;           // which half of the columns (always 0 with 128 epilogue threads)
    constexpr int HALVES = EPI / 128;
    const int row = ((ewarp + 2) & 3) * 32 + lane;
    const int tid = ewarp * 32 + lane;
    (void)quad;
    const size_t r_glob = (size_t)mt * GM_BM + row;
    const uint32_t t_lane = (uint32_t)(((ewarp + 2) & 3) * 32) << 16;
    const int cout = pr.cout;
    // stage the tile's column parameters (previous tile's readers are past this point: barrier first)
    gm_epi_barrier<EPI>();
    for (int c = tid; c < NT; c += EPI) {
        const int col = nti * NT + c;
        cp.bias[c] = pr.bias != nullptr ? __ldg(pr.bias + col) : 0.f;
        cp.scale[c] = pr.post_scale != nullptr ? __ldg(pr.post_scale + col) : 1.f;
        cp.shift[c] = pr.post_shift != nullptr ? __ldg(pr.post_shift + col) : 0.f;
        cp.dotw[c] = pr.dot_w != nullptr ? __ldg(pr.dot_w + col) : 0.f;
    }
    gm_epi_barrier<EPI>();
    const float relu_floor = pr.relu ? 0.f : -3.402823466e+38f;
    const float rw = pr.pool_w != nullptr ? __ldg(pr.pool_w + r_glob) : 0.f;
    size_t cm_base = 0;
    if (pr.out_cm != nullptr) {
        const size_t inst = r_glob / pr.rows_per_inst, within = r_glob - inst * pr.rows_per_inst;
        cm_base = inst * (size_t)cout * pr.rows_per_inst + within;
    }
    float dot = 0.f;
    constexpr int CH = NT / 32 / HALVES;  // 32-column chunks per thread
#pragma unroll 1
    for (int cc = half * CH; cc < (half + 1) * CH; ++cc) {
        uint32_t v[32];
        DCL_TMEM_LD32(tmem_acc + t_lane + cc * 32, v);
        tc_wait_ld();
        const int col0 = nti * NT + cc * 32;
        float y[32];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const float4 bb = reinterpret_cast<const float4*>(cp.bias + cc * 32)[q];
            const float4 sc = reinterpret_cast<const float4*>(cp.scale + cc * 32)[q];
            const float4 sh = reinterpret_cast<const float4*>(cp.shift + cc * 32)[q];
            y[q * 4 + 0] = __fmaf_rn(fmaxf(__uint_as_float(v[q * 4 + 0]) + bb.x, relu_floor), sc.x, sh.x);
            y[q * 4 + 1] = __fmaf_rn(fmaxf(__uint_as_float(v[q * 4 + 1]) + bb.y, relu_floor), sc.y, sh.y);
            y[q * 4 + 2] = __fmaf_rn(fmaxf(__uint_as_float(v[q * 4 + 2]) + bb.z, relu_floor), sc.z, sh.z);
            y[q * 4 + 3] = __fmaf_rn(fmaxf(__uint_as_float(v[q * 4 + 3]) + bb.w, relu_floor), sc.w, sh.w);
        }
        if (pr.dot_out != nullptr) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const float4 dw = reinterpret_cast<const float4*>(cp.dotw + cc * 32)[q];
                dot = __fmaf_rn(y[q * 4 + 0], dw.x, dot);
                dot = __fmaf_rn(y[q * 4 + 1], dw.y, dot);
                dot = __fmaf_rn(y[q * 4 + 2], dw.z, dot);
                dot = __fmaf_rn(y[q * 4 + 3], dw.w, dot);
            }
        }
        if (pr.out_fmt == FMT_F16 && (pr.out_pm != nullptr || pr.out_v != nullptr)) {
            // PM16: the 32 values once-rounded to fp16, as four 8-channel (16-byte) units
            uint4 fq[4];
#pragma unroll
            for (int ch = 0; ch < 4; ++ch)
                fq[ch] = make_uint4(pack_f16x2(y[ch * 8 + 0], y[ch * 8 + 1]), pack_f16x2(y[ch * 8 + 2], y[ch * 8 + 3]),
                                    pack_f16x2(y[ch * 8 + 4], y[ch * 8 + 5]), pack_f16x2(y[ch * 8 + 6], y[ch * 8 + 7]));
            if (pr.out_pm != nullptr) {
                unsigned char* blob = reinterpret_cast<unsigned char*>(pr.out_pm) +
                                      ((size_t)mt * (cout / 32) + col0 / 32) * GM_A16_BLOB + (row >> 3) * 512 + (row & 7) * 16;
#pragma unroll
                for (int ch = 0; ch < 4; ++ch) *reinterpret_cast<uint4*>(blob + ch * 128) = fq[ch];
            }
            if (pr.out_v != nullptr) {
                // fp16 value image of the fused FDA kernel: chunks of 16 keys x v_rows channels, one image per chunk
                const size_t vchunk = (size_t)pr.v_rows * 32;
                unsigned char* d = reinterpret_cast<unsigned char*>(pr.out_v) + (r_glob >> 4) * vchunk +
                                   (size_t)((pr.v_row0 + col0) >> 3) * 256 + ((r_glob >> 3) & 1) * 128 + (r_glob & 7) * 16;
#pragma unroll
                for (int ch = 0; ch < 4; ++ch) *reinterpret_cast<uint4*>(d + ch * 256) = fq[ch];
            }
        }
        const bool pm_v_bf16 = pr.out_fmt == FMT_BF16X2;
        if ((pm_v_bf16 && (pr.out_pm != nullptr || pr.out_v != nullptr)) || pr.out_qk != nullptr) {
            // bf16 hi / lo halves of the 32 values, as four 8-channel (16-byte) units each
            uint4 hq[4], lq[4];
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) {
                uint32_t h[4], l[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) split2_bf16(y[ch * 8 + 2 * e], y[ch * 8 + 2 * e + 1], h[e], l[e]);
                hq[ch] = make_uint4(h[0], h[1], h[2], h[3]);
                lq[ch] = make_uint4(l[0], l[1], l[2], l[3]);
            }
            if (pm_v_bf16 && pr.out_pm != nullptr) {
                unsigned char* blob = reinterpret_cast<unsigned char*>(pr.out_pm) +
                                      ((size_t)mt * (cout / 32) + col0 / 32) * GM_A_BLOB + (row >> 3) * 512 + (row & 7) * 16;
#pragma unroll
                for (int ch = 0; ch < 4; ++ch) {
                    *reinterpret_cast<uint4*>(blob + ch * 128) = hq[ch];
                    *reinterpret_cast<uint4*>(blob + GM_A_BLOB / 2 + ch * 128) = lq[ch];
                }
            }
            if (pr.out_qk != nullptr) {
                // FDA query / key operand image (fda.cu): tiles of T rows x cout channels, hi image then lo image,
                // element (r, c) at (r/8)*(cout/8)*128 + (c/8)*128 + (r%8)*16 + (c%8)*2 bytes
                const int T = pr.qk_tile_rows;
                const size_t half = (size_t)T * cout * 2;
                const int rt = (int)(r_glob % T);
                unsigned char* d = reinterpret_cast<unsigned char*>(pr.out_qk) + (r_glob / T) * 2 * half +
                                   (size_t)(rt >> 3) * (cout / 8) * 128 + (rt & 7) * 16 + (col0 / 8) * 128;
#pragma unroll
                for (int ch = 0; ch < 4; ++ch) {
                    *reinterpret_cast<uint4*>(d + ch * 128) = hq[ch];
                    *reinterpret_cast<uint4*>(d + half + ch * 128) = lq[ch];
                }
            }
            if (pm_v_bf16 && pr.out_v != nullptr) {
                // FDA value image (fda.cu): chunks of 16 keys x v_rows value channels, hi then lo; this row is a key,
                // its columns are value channels v_row0 + col: 8 channels of one key = one 16-byte unit at
                // (vrow/8)*256 + ((key%16)/8)*128 + (key%8)*16
                const size_t vhalf = (size_t)pr.v_rows * 32;
                unsigned char* d = reinterpret_cast<unsigned char*>(pr.out_v) + (r_glob >> 4) * 2 * vhalf +
                                   (size_t)((pr.v_row0 + col0) >> 3) * 256 + ((r_glob >> 3) & 1) * 128 + (r_glob & 7) * 16;
#pragma unroll
                for (int ch = 0; ch < 4; ++ch) {
                    *reinterpret_cast<uint4*>(d + ch * 256) = hq[ch];
                    *reinterpret_cast<uint4*>(d + vhalf + ch * 256) = lq[ch];
                }
            }
        }
        if (pr.out_cm != nullptr) {
            float* o = pr.out_cm + cm_base + (size_t)col0 * pr.rows_per_inst;
#pragma unroll
            for (int i = 0; i < 32; ++i) o[(size_t)i * pr.rows_per_inst] = y[i];
        }
        if (pr.pool_out != nullptr) {
            float pv[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) pv[i] = y[i] * rw;
            // transpose-reduce: afterwards pv[0] on lane L = sum over the warp's 32 rows of column L
#pragma unroll
            for (int st = 16; st >= 1; st >>= 1) {
                const bool upper = (lane & st) != 0;
#pragma unroll
                for (int i = 0; i < st; ++i) {
                    const float send = upper ? pv[i] : pv[i + st];
                    const float keep = upper ? pv[i + st] : pv[i];
                    pv[i] = keep + __shfl_xor_sync(0xffffffffu, send, st);
                }
            }
            pr.pool_out[(r_glob >> 5) * cout + col0 + lane] = pv[0];
        }
    }
    if (pr.dot_out != nullptr) {
        if (HALVES == 1) {
            pr.dot_out[r_glob] = dot;
        } else {
            // the two threads of a row add their halves (lower columns first, as a single thread would)
            if (half == 1) s_dot[row] = dot;
            gm_epi_barrier<EPI>();
            if (half == 0) pr.dot_out[r_glob] = dot + s_dot[row];
        }
    }
}

template <int NT, int STAGES, int FMT>
__global__ void __launch_bounds__(GM_THREADS, 2) pm_gemm_kernel(const __grid_constant__ PmGemmBatch batch, int nprob) {
    using Cfg = GmCfg<NT, STAGES, FMT>;
    constexpr int A_BLOB = Cfg::A_BLOB;
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(16) GmColParams s_colp;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);
    uint64_t* empty = full + STAGES;
    uint64_t* acc_full = empty + STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

    // grid = (problems, n-tiles, m-tiles): CTAs that read the same A blobs (same m-tile: the other n-tiles of a
    // layer, the other problems sharing the input) are launched next to each other, so each blob comes from DRAM
    // once and is an L2 hit for the rest (with the m-tile fastest, ncu showed 3x the unique bytes read from DRAM).
    const dcl_pm_gemm_problem& pr = batch.p[blockIdx.x % nprob];
    const int inst = blockIdx.x / nprob;   // slice of a strided-batch problem (0 unless inst_count > 1)
    const int mt = blockIdx.z, nti = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int KB = pr.kb_total;

    if (threadIdx.x == 0) {
        for (int i = 0; i < STAGES; ++i) {
            dcl_mbar_init(full + i, 1);
            dcl_mbar_init(empty + i, 1);
        }
        dcl_mbar_init(acc_full, 1);
        dcl_fence_barrier_init();
    }
    if (warp == 1) tc_alloc(tmem_slot, NT);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (dcl_elect_one()) {
            const unsigned char* a0 = reinterpret_cast<const unsigned char*>(pr.a0) + (size_t)mt * pr.kb0 * A_BLOB +
                                      (size_t)inst * pr.a_inst_stride;
            const unsigned char* a1 = reinterpret_cast<const unsigned char*>(pr.a1) +
                                      (size_t)mt * (KB - pr.kb0) * A_BLOB;
            const unsigned char* w = reinterpret_cast<const unsigned char*>(pr.w) + (size_t)nti * KB * Cfg::B_BLOB +
                                     (size_t)inst * pr.w_inst_stride;
            for (int kb = 0; kb < KB; ++kb) {
                const int s = kb % STAGES;
                if (kb >= STAGES) dcl_mbar_wait(empty + s, (uint32_t)(((kb / STAGES) - 1) & 1));
                unsigned char* dst = smem + s * Cfg::STAGE_BYTES;
                dcl_mbar_arrive_expect_tx(full + s, Cfg::STAGE_BYTES);
                const unsigned char* asrc = (kb < pr.kb0) ? a0 + (size_t)kb * A_BLOB
                                                          : a1 + (size_t)(kb - pr.kb0) * A_BLOB;
                dcl_bulk_g2s(dst, asrc, A_BLOB, full + s);
                dcl_bulk_g2s(dst + A_BLOB, w + (size_t)kb * Cfg::B_BLOB, Cfg::B_BLOB, full + s);
            }
        }
    } else if (warp == 1) {
        if (dcl_elect_one()) {
            constexpr uint32_t idesc = FMT == FMT_F16 ? umma_idesc_f16(GM_BM, NT) : umma_idesc_bf16(GM_BM, NT);
            for (int kb = 0; kb < KB; ++kb) {
                const int s = kb % STAGES;
                dcl_mbar_wait(full + s, (uint32_t)((kb / STAGES) & 1));
                tc_fence_after();
                const uint32_t a = dcl_smem_u32(smem + s * Cfg::STAGE_BYTES);
                const uint32_t b = a + A_BLOB;
#pragma unroll
                for (int ks = 0; ks < GM_BK / 16; ++ks) {
                    const uint32_t off = ks * 256;
                    if constexpr (FMT == FMT_F16) {
                        // X (fp16, once-rounded) * (W_hi + W_lo): two MMAs
                        const uint64_t dA = umma_desc(a + off, 128, 512);
                        tc_mma_bf16(tmem_base, dA, umma_desc(b + off, 128, 512), idesc, (kb == 0 && ks == 0) ? 0u : 1u);
                        tc_mma_bf16(tmem_base, dA, umma_desc(b + Cfg::B_BLOB / 2 + off, 128, 512), idesc, 1u);
                    } else {
                        mma_split3(tmem_base, a + off, a + A_BLOB / 2 + off, b + off, b + Cfg::B_BLOB / 2 + off, 128, 512,
                                   128, 512, idesc, kb == 0 && ks == 0);
                    }
                }
                tc_commit(empty + s);
            }
            tc_commit(acc_full);
        }
    } else {
        // ===================== epilogue =====================
        dcl_mbar_wait(acc_full, 0);
        tc_fence_after();
        gm_epilogue_tile<NT, 128>(pr, mt, nti, tmem_base, warp - 2, lane, s_colp, nullptr, inst);
        tc_fence_before();
    }
    __syncwarp();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tc_dealloc(tmem_base, NT);
    }
}

// ------------------------------------------------------------------ persistent, cluster-of-2 variant
// Same math and epilogue; what changes is how the tensor pipe is kept fed (ncu on the kernel above: tensor pipe
// active ~45 %; a 2-stage ring per CTA cannot cover the L2 latency and 62 B/clk/SM of operand traffic is at the
// limit of what L2 delivers):
//   * one persistent CTA per SM walks a static list of tiles; a deep ring (192 KB of stages) runs ahead across
//     tile boundaries; two TMEM accumulators alternate so the epilogue of tile i overlaps the MMAs of tile i+1;
//   * CTAs are paired in a cluster: the pair works on two m-tiles of the same (problem, n-tile), so both need the
//     same weight blobs at the same time — CTA 0 fetches the hi half, CTA 1 the lo half, each MULTICASTS its half
//     into both CTAs' shared memory (cp.async.bulk ... .multicast::cluster).  Operand traffic per SM drops from
//     48 to 32 KB per k-block.  A stage is refilled only after BOTH CTAs' MMAs released it: tcgen05.commit is
//     multicast to the pair's `empty` barriers (2 arrivals per phase).

template <int NT, int STAGES>
struct GmPCfg {
    static constexpr int B_BLOB = NT * GM_BK * 4;
    static constexpr int STAGE_BYTES = GM_A_BLOB + B_BLOB;
    static constexpr int OFF_BAR = STAGES * STAGE_BYTES;
    static constexpr int SMEM_BYTES = OFF_BAR + 256;
    static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
    static_assert(2 * NT <= 512, "TMEM budget");
};

template <int NT, int STAGES>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GM_P_THREADS, 1)
    pm_gemm_cluster_kernel(const __grid_constant__ PmGemmBatch batch, int nprob, int ntiles_n, int npairs_m) {
    using Cfg = GmPCfg<NT, STAGES>;
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(16) GmColParams s_colp;
    __shared__ float s_dot[GM_BM];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);
    uint64_t* empty = full + STAGES;
    uint64_t* acc_full = empty + STAGES;   // [2]
    uint64_t* acc_empty = acc_full + 2;    // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = dcl_cluster_ctarank();
    const int cluster_id = blockIdx.x >> 1, nclusters = gridDim.x >> 1;
    const int total_units = npairs_m * ntiles_n * nprob;   // unit = (m-tile pair, n-tile, problem); problem fastest

    if (threadIdx.x == 0) {
        for (int i = 0; i < STAGES; ++i) {
            dcl_mbar_init(full + i, 1);
            dcl_mbar_init(empty + i, 2);
        }
        for (int i = 0; i < 2; ++i) {
            dcl_mbar_init(acc_full + i, 1);
            dcl_mbar_init(acc_empty + i, GM_P_EPI);
        }
        dcl_fence_barrier_init();
    }
    if (warp == 1) tc_alloc(tmem_slot, 2 * NT);
    tc_fence_before();
    __syncthreads();
    dcl_cluster_sync();  // the peer's barriers are initialised before anything is multicast at them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (dcl_elect_one()) {
            int it = 0;
            for (int u = cluster_id; u < total_units; u += nclusters) {
                const int prob = u % nprob, nti = (u / nprob) % ntiles_n, mt = 2 * (u / (nprob * ntiles_n)) + (int)rank;
                const dcl_pm_gemm_problem& pr = batch.p[prob];
                const int KB = pr.kb_total;
                const unsigned char* a0 = reinterpret_cast<const unsigned char*>(pr.a0) + (size_t)mt * pr.kb0 * GM_A_BLOB;
                const unsigned char* a1 = reinterpret_cast<const unsigned char*>(pr.a1) +
                                          (size_t)mt * (KB - pr.kb0) * GM_A_BLOB;
                const unsigned char* w = reinterpret_cast<const unsigned char*>(pr.w) + (size_t)nti * KB * Cfg::B_BLOB +
                                         rank * (Cfg::B_BLOB / 2);
                for (int kb = 0; kb < KB; ++kb, ++it) {
                    const int s = it % STAGES;
                    if (it >= STAGES) dcl_mbar_wait(empty + s, (uint32_t)(((it / STAGES) - 1) & 1));
                    unsigned char* dst = smem + s * Cfg::STAGE_BYTES;
                    dcl_mbar_arrive_expect_tx(full + s, Cfg::STAGE_BYTES);
                    const unsigned char* asrc = (kb < pr.kb0) ? a0 + (size_t)kb * GM_A_BLOB
                                                              : a1 + (size_t)(kb - pr.kb0) * GM_A_BLOB;
                    dcl_bulk_g2s(dst, asrc, GM_A_BLOB, full + s);
                    dcl_bulk_g2s_mcast(dst + GM_A_BLOB + rank * (Cfg::B_BLOB / 2), w + (size_t)kb * Cfg::B_BLOB,
                                       Cfg::B_BLOB / 2, full + s, (uint16_t)0x3);
                }
            }
        }
    } else if (warp == 1) {
        if (dcl_elect_one()) {
            constexpr uint32_t idesc = umma_idesc_bf16(GM_BM, NT);
            const uint64_t desc0 = umma_desc(dcl_smem_u32(smem), 128, 512);
            int it = 0, tl = 0;
            for (int u = cluster_id; u < total_units; u += nclusters, ++tl) {
                const int KB = batch.p[u % nprob].kb_total;
                const int acc = tl & 1;
                if (tl >= 2) dcl_mbar_wait(acc_empty + acc, (uint32_t)(((tl >> 1) - 1) & 1));
                tc_fence_after();
                const uint32_t tacc = tmem_base + acc * NT;
                for (int kb = 0; kb < KB; ++kb, ++it) {
                    const int s = it % STAGES;
                    dcl_mbar_wait(full + s, (uint32_t)((it / STAGES) & 1));
                    tc_fence_after();
                    // descriptors differ only in their start-address field
                    const uint64_t dAh = desc0 + (uint64_t)((s * Cfg::STAGE_BYTES) >> 4);
                    const uint64_t dAl = dAh + (uint64_t)((GM_A_BLOB / 2) >> 4);
                    const uint64_t dBh = dAh + (uint64_t)(GM_A_BLOB >> 4);
                    const uint64_t dBl = dBh + (uint64_t)((Cfg::B_BLOB / 2) >> 4);
#pragma unroll
                    for (int ks = 0; ks < GM_BK / 16; ++ks) {
                        const uint64_t off = (uint64_t)((ks * 256) >> 4);
                        tc_mma_bf16(tacc, dAh + off, dBh + off, idesc, (kb == 0 && ks == 0) ? 0u : 1u);
                        tc_mma_bf16(tacc, dAh + off, dBl + off, idesc, 1u);
                        tc_mma_bf16(tacc, dAl + off, dBh + off, idesc, 1u);
                    }
                    tc_commit_mcast(empty + s, (uint16_t)0x3);
                }
                tc_commit(acc_full + acc);
            }
        }
    } else {
        int tl = 0;
        for (int u = cluster_id; u < total_units; u += nclusters, ++tl) {
            const int prob = u % nprob, nti = (u / nprob) % ntiles_n, mt = 2 * (u / (nprob * ntiles_n)) + (int)rank;
            const int acc = tl & 1;
            dcl_mbar_wait(acc_full + acc, (uint32_t)((tl >> 1) & 1));
            tc_fence_after();
            gm_epilogue_tile<NT, GM_P_EPI>(batch.p[prob], mt, nti, tmem_base + acc * NT, warp - 2, lane, s_colp, s_dot);
            tc_fence_before();
            dcl_mbar_arrive(acc_empty + acc);
        }
    }
    __syncwarp();
    __syncthreads();
    dcl_cluster_sync();  // no CTA leaves while its peer may still multicast into it
    if (warp == 1) {
        tc_fence_after();
        tc_dealloc(tmem_base, 2 * NT);
    }
}

// ------------------------------------------------------------------ persistent CTA-pair variant (cta_group::2)
// The pair runs every k-step as ONE M = 256 tcgen05.mma issued by the leader: each CTA stages its own A blob and only
// HALF of the W blob (output rows [r NT/2, (r+1) NT/2), hi and lo), the tensor cores of both SMs read both halves.
// Against the multicast kernel above this takes the ingest per SM and k-block from 48 KB to 32 KB, the operand
// reads from shared memory with it, and lets the ring hold STAGES2 = 6/9/12 stages.  Hand-offs as in fda.cu's pair
// kernel (conventions pinned by dcl_debug_umma_pair_gemm): the peer's MMA warp relays "my stage landed" to the
// leader's full barriers, every tcgen05.commit is multicast to both CTAs, epilogue warps of both CTAs release the
// accumulator with one CTA-scope arrive per warp on the leader's barrier.
template <int NT, int STAGES, int FMT = FMT_BF16X2>
struct GmP2Cfg {
    static constexpr int A_BLOB = GmFmt<FMT>::A_BLOB;
    static constexpr int B_BLOB = NT * GM_BK * 4;          // whole W blob in global memory: [hi | lo]
    static constexpr int B_HALF_ROWS = B_BLOB / 4;         // this CTA's rows of the hi (or lo) image
    static constexpr int STAGE_BYTES = A_BLOB + B_BLOB / 2;
    static constexpr int OFF_BAR = STAGES * STAGE_BYTES;
    static constexpr int SMEM_BYTES = OFF_BAR + 512;
    static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
    static_assert(2 * NT <= 512, "TMEM budget");
    static_assert(2 * STAGES + 4 <= 60, "barrier area");
};

template <int NT, int STAGES, int FMT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GM_P_THREADS, 1)
    pm_gemm_pair_kernel(const __grid_constant__ PmGemmBatch batch, int nprob, int ntiles_n, int npairs_m, int ninst) {
    using Cfg = GmP2Cfg<NT, STAGES, FMT>;
    constexpr int A_BLOB = Cfg::A_BLOB;
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(16) GmColParams s_colp;
    __shared__ float s_dot[GM_BM];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);
    uint64_t* empty = full + STAGES;
    uint64_t* acc_full = empty + STAGES;   // [2]
    uint64_t* acc_empty = acc_full + 2;    // [2] (the leader's are the ones waited on)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = dcl_cluster_ctarank();
    const bool leader = rank == 0;
    const int cluster_id = blockIdx.x >> 1, nclusters = gridDim.x >> 1;
    // unit = (m-tile pair, slice, n-tile, problem), problem fastest: CTAs running side by side read the same A blobs
    // (the problems / n-tiles sharing an input); `slice` = the split-K instance of a strided-batch problem
    const int total_units = npairs_m * ntiles_n * nprob * ninst;

    if (threadIdx.x == 0) {
        for (int i = 0; i < STAGES; ++i) {
            dcl_mbar_init(full + i, leader ? 2 : 1);       // leader: own producer + the peer's relay
            dcl_mbar_init(empty + i, 1);                   // one multicast commit per phase
        }
        for (int i = 0; i < 2; ++i) {
            dcl_mbar_init(acc_full + i, 1);
            dcl_mbar_init(acc_empty + i, 2 * (GM_P_EPI / 32));  // one arrive per epilogue warp of both CTAs
        }
        dcl_fence_barrier_init();
    }
    __syncthreads();
    dcl_cluster_sync();  // both CTAs' barriers exist before anything is signalled at them
    if (warp == 1) tc2_alloc(tmem_slot, 2 * NT);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (dcl_elect_one()) {
            int it = 0;
            for (int u = cluster_id; u < total_units; u += nclusters) {
                const int prob = u % nprob, nti = (u / nprob) % ntiles_n, inst = (u / (nprob * ntiles_n)) % ninst;
                const int mt = 2 * (u / (nprob * ntiles_n * ninst)) + (int)rank;
                const dcl_pm_gemm_problem& pr = batch.p[prob];
                const int KB = pr.kb_total;
                const unsigned char* a0 = reinterpret_cast<const unsigned char*>(pr.a0) + (size_t)mt * pr.kb0 * A_BLOB +
                                          (size_t)inst * pr.a_inst_stride;
                const unsigned char* a1 = reinterpret_cast<const unsigned char*>(pr.a1) +
                                          (size_t)mt * (KB - pr.kb0) * A_BLOB;
                const unsigned char* w = reinterpret_cast<const unsigned char*>(pr.w) + (size_t)nti * KB * Cfg::B_BLOB +
                                         rank * Cfg::B_HALF_ROWS + (size_t)inst * pr.w_inst_stride;
                for (int kb = 0; kb < KB; ++kb, ++it) {
                    const int s = it % STAGES;
                    if (it >= STAGES) dcl_mbar_wait(empty + s, (uint32_t)(((it / STAGES) - 1) & 1));
                    unsigned char* dst = smem + s * Cfg::STAGE_BYTES;
                    dcl_mbar_arrive_expect_tx(full + s, Cfg::STAGE_BYTES);
                    const unsigned char* asrc = (kb < pr.kb0) ? a0 + (size_t)kb * A_BLOB
                                                              : a1 + (size_t)(kb - pr.kb0) * A_BLOB;
                    const unsigned char* wsrc = w + (size_t)kb * Cfg::B_BLOB;
                    dcl_bulk_g2s(dst, asrc, A_BLOB, full + s);
                    dcl_bulk_g2s(dst + A_BLOB, wsrc, Cfg::B_HALF_ROWS, full + s);                          // hi rows
                    dcl_bulk_g2s(dst + A_BLOB + Cfg::B_HALF_ROWS, wsrc + Cfg::B_BLOB / 2, Cfg::B_HALF_ROWS,
                                 full + s);                                                                   // lo rows
                }
            }
        }
    } else if (warp == 1) {
        if (leader && dcl_elect_one()) {
            constexpr uint32_t idesc = FMT == FMT_F16 ? umma_idesc_f16(2 * GM_BM, NT) : umma_idesc_bf16(2 * GM_BM, NT);
            const uint64_t desc0 = umma_desc(dcl_smem_u32(smem), 128, 512);
            int it = 0, tl = 0;
            for (int u = cluster_id; u < total_units; u += nclusters, ++tl) {
                const int KB = batch.p[u % nprob].kb_total;
                const int acc = tl & 1;
                if (tl >= 2) dcl_mbar_wait(acc_empty + acc, (uint32_t)(((tl >> 1) - 1) & 1));
                tc_fence_after();
                const uint32_t tacc = tmem_base + acc * NT;
                for (int kb = 0; kb < KB; ++kb, ++it) {
                    const int s = it % STAGES;
                    dcl_mbar_wait(full + s, (uint32_t)((it / STAGES) & 1));
                    tc_fence_after();
                    const uint64_t dAh = desc0 + (uint64_t)((s * Cfg::STAGE_BYTES) >> 4);
                    const uint64_t dAl = dAh + (uint64_t)((A_BLOB / 2) >> 4);   // bf16 hi/lo format only
                    const uint64_t dBh = dAh + (uint64_t)(A_BLOB >> 4);
                    const uint64_t dBl = dBh + (uint64_t)(Cfg::B_HALF_ROWS >> 4);
#pragma unroll
                    for (int ks = 0; ks < GM_BK / 16; ++ks) {
                        const uint64_t off = (uint64_t)((ks * 256) >> 4);
                        tc2_mma_bf16(tacc, dAh + off, dBh + off, idesc, (kb == 0 && ks == 0) ? 0u : 1u);
                        tc2_mma_bf16(tacc, dAh + off, dBl + off, idesc, 1u);
                        if constexpr (FMT == FMT_BF16X2) tc2_mma_bf16(tacc, dAl + off, dBh + off, idesc, 1u);
                    }
                    tc2_commit_mcast(empty + s, (uint16_t)0x3);
                }
                tc2_commit_mcast(acc_full + acc, (uint16_t)0x3);
            }
        } else if (!leader && dcl_elect_one()) {
            // relay: "my operands of this stage have landed", in ring order (= the leader's consumption order)
            int it = 0;
            for (int u = cluster_id; u < total_units; u += nclusters) {
                const int KB = batch.p[u % nprob].kb_total;
                for (int kb = 0; kb < KB; ++kb, ++it) {
                    const int s = it % STAGES;
                    dcl_mbar_wait(full + s, (uint32_t)((it / STAGES) & 1));
                    dcl_mbar_arrive_remote(full + s, 0);
                }
            }
        }
    } else {
        int tl = 0;
        for (int u = cluster_id; u < total_units; u += nclusters, ++tl) {
            const int prob = u % nprob, nti = (u / nprob) % ntiles_n, inst = (u / (nprob * ntiles_n)) % ninst;
            const int mt = 2 * (u / (nprob * ntiles_n * ninst)) + (int)rank;
            const int acc = tl & 1;
            dcl_mbar_wait(acc_full + acc, (uint32_t)((tl >> 1) & 1));
            tc_fence_after();
            gm_epilogue_tile<NT, GM_P_EPI>(batch.p[prob], mt, nti, tmem_base + acc * NT, warp - 2, lane, s_colp, s_dot,
                                           inst);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) dcl_mbar_arrive_remote(acc_empty + acc, 0);
        }
    }
    __syncwarp();
    __syncthreads();
    dcl_cluster_sync();  // the leader's MMAs read the peer's shared memory until the last commit
    if (warp == 1) {
        tc_fence_after();
        tc2_dealloc(tmem_base, 2 * NT);
    }
}

template <int NT, int STAGES, int FMT>
int launch_gemm_pair(const PmGemmBatch& batch, int nprob, int rows, int cout, cudaStream_t st, int ninst) {
    using Cfg = GmP2Cfg<NT, STAGES, FMT>;
    static int num_sms = 0;
    if (num_sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    }
    cudaError_t e = cudaFuncSetAttribute(pm_gemm_pair_kernel<NT, STAGES, FMT>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
    const int ntiles_n = cout / NT, npairs_m = rows / (2 * GM_BM);
    const int units = npairs_m * ntiles_n * nprob * ninst;
    int clusters = num_sms / 2;
    if (clusters > units) clusters = units;
    pm_gemm_pair_kernel<NT, STAGES, FMT><<<2 * clusters, GM_P_THREADS, Cfg::SMEM_BYTES, st>>>(batch, nprob, ntiles_n,
                                                                                             npairs_m, ninst);
    return dcl_launch_status();
}

template <int NT, int STAGES>
int launch_gemm_cluster(const PmGemmBatch& batch, int nprob, int rows, int cout, cudaStream_t st) {
    using Cfg = GmPCfg<NT, STAGES>;
    static int num_sms = 0;
    if (num_sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    }
    cudaError_t e = cudaFuncSetAttribute(pm_gemm_cluster_kernel<NT, STAGES>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
    const int ntiles_n = cout / NT, npairs_m = rows / (2 * GM_BM);
    const int units = npairs_m * ntiles_n * nprob;
    int clusters = num_sms / 2;
    if (clusters > units) clusters = units;
    pm_gemm_cluster_kernel<NT, STAGES><<<2 * clusters, GM_P_THREADS, Cfg::SMEM_BYTES, st>>>(batch, nprob, ntiles_n,
                                                                                        npairs_m);
    return dcl_launch_status();
}

// ------------------------------------------------------------------ packing helpers
// fp32 row-major (rows x c, row stride `ld`) -> PM image.  Thread = (row, chunk of 8 channels).
__global__ void __launch_bounds__(256) pm_pack_rows_kernel(int rows, int c, int ld, const float* __restrict__ src,
                                                           unsigned char* __restrict__ dst, int fmt) {
    const long g = (long)blockIdx.x * 256 + threadIdx.x;
    const int nchunk = c / 8;
    if (g >= (long)rows * nchunk) return;
    // consecutive threads -> consecutive rows of the same chunk: 8 lanes write one 128-B run
    const int r = (int)(g % rows), ch = (int)(g / rows);
    const float* s = src + (size_t)r * ld + ch * 8;
    if (fmt == FMT_F16) {
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = __ldg(s + e);
        unsigned char* d = dst + ((size_t)(r / 128) * (c / 32) + ch / 4) * GM_A16_BLOB + ((r % 128) >> 3) * 512 +
                           (ch & 3) * 128 + (r & 7) * 16;
        *reinterpret_cast<uint4*>(d) = make_uint4(pack_f16x2(v[0], v[1]), pack_f16x2(v[2], v[3]), pack_f16x2(v[4], v[5]),
                                                  pack_f16x2(v[6], v[7]));
        return;
    }
    __nv_bfloat16 h[8], l[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) split_bf16(__ldg(s + e), h[e], l[e]);
    unsigned char* d = dst + ((size_t)(r / 128) * (c / 32) + ch / 4) * GM_A_BLOB + ((r % 128) >> 3) * 512 + (ch & 3) * 128 +
                       (r & 7) * 16;
    *reinterpret_cast<uint4*>(d) = make_uint4(pack2(h[0], h[1]), pack2(h[2], h[3]), pack2(h[4], h[5]), pack2(h[6], h[7]));
    *reinterpret_cast<uint4*>(d + GM_A_BLOB / 2) =
        make_uint4(pack2(l[0], l[1]), pack2(l[2], l[3]), pack2(l[4], l[5]), pack2(l[6], l[7]));
}

// fp32 channel-major (b, c, n) -> PM image of the (b*n x c) activation.  Thread = (row, chunk); reads coalesce
// along n for a fixed channel.
__global__ void __launch_bounds__(256) pm_pack_cm_kernel(int b, int c, int n, const float* __restrict__ src,
                                                         unsigned char* __restrict__ dst, int fmt) {
    const long g = (long)blockIdx.x * 256 + threadIdx.x;
    const long rows = (long)b * n;
    const int nchunk = c / 8;
    if (g >= rows * nchunk) return;
    const long r = g % rows;
    const int ch = (int)(g / rows);
    const int inst = (int)(r / n), within = (int)(r % n);
    const float* s = src + ((size_t)inst * c + ch * 8) * n + within;
    if (fmt == FMT_F16) {
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = __ldg(s + (size_t)e * n);
        unsigned char* d = dst + ((size_t)(r / 128) * (c / 32) + ch / 4) * GM_A16_BLOB + ((r % 128) >> 3) * 512 +
                           (ch & 3) * 128 + (r & 7) * 16;
        *reinterpret_cast<uint4*>(d) = make_uint4(pack_f16x2(v[0], v[1]), pack_f16x2(v[2], v[3]), pack_f16x2(v[4], v[5]),
                                                  pack_f16x2(v[6], v[7]));
        return;
    }
    __nv_bfloat16 h[8], l[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) split_bf16(__ldg(s + (size_t)e * n), h[e], l[e]);
    unsigned char* d = dst + ((size_t)(r / 128) * (c / 32) + ch / 4) * GM_A_BLOB + ((r % 128) >> 3) * 512 + (ch & 3) * 128 +
                       (r & 7) * 16;
    *reinterpret_cast<uint4*>(d) = make_uint4(pack2(h[0], h[1]), pack2(h[2], h[3]), pack2(h[4], h[5]), pack2(h[6], h[7]));
    *reinterpret_cast<uint4*>(d + GM_A_BLOB / 2) =
        make_uint4(pack2(l[0], l[1]), pack2(l[2], l[3]), pack2(l[4], l[5]), pack2(l[6], l[7]));
}

// PM image -> fp32 row-major (hi + lo); tests / inspection.
__global__ void __launch_bounds__(256) pm_unpack_kernel(int rows, int c, const unsigned char* __restrict__ src,
                                                        float* __restrict__ dst, int fmt) {
    const long g = (long)blockIdx.x * 256 + threadIdx.x;
    if (g >= (long)rows * c) return;
    const int r = (int)(g / c), ch = (int)(g % c);
    if (fmt == FMT_F16) {
        const unsigned char* s16 = src + ((size_t)(r / 128) * (c / 32) + ch / 32) * GM_A16_BLOB + ((r % 128) >> 3) * 512 +
                                   ((ch % 32) >> 3) * 128 + (r & 7) * 16 + (ch & 7) * 2;
        dst[g] = __half2float(*reinterpret_cast<const __half*>(s16));
        return;
    }
    const unsigned char* s = src + ((size_t)(r / 128) * (c / 32) + ch / 32) * GM_A_BLOB + ((r % 128) >> 3) * 512 +
                             ((ch % 32) >> 3) * 128 + (r & 7) * 16 + (ch & 7) * 2;
    const float hi = __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(s));
    const float lo = __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(s + GM_A_BLOB / 2));
    dst[g] = hi + lo;
}

// out[inst, col] = sum over the `parts` per-warp partials of an instance, in index order (deterministic), then —
// when a second set of partials is given — over that set, continuing the same running sum (so one launch equals
// two accumulating launches bit for bit); accumulate != 0 starts from what is already there.
__global__ void __launch_bounds__(256) pm_pool_reduce_kernel(int insts, int cout, int parts,
                                                             const float* __restrict__ partials,
                                                             const float* __restrict__ partials2,
                                                             float* __restrict__ out, int accumulate) {
    const int g = blockIdx.x * 256 + threadIdx.x;
    if (g >= insts * cout) return;
    const int inst = g / cout, col = g % cout;
    float acc = accumulate ? out[g] : 0.f;
    // the additions stay in index order; the loads of a batch of 16 are independent of the running sum, so they are
    // all in flight together whatever `parts` is (a plain loop serialises load -> add when it is not unrolled)
    auto add_set = [&](const float* p) {
        for (int i0 = 0; i0 < parts; i0 += 16) {
            float v[16];
#pragma unroll
            for (int u = 0; u < 16; ++u) v[u] = (i0 + u < parts) ? __ldg(p + (size_t)(i0 + u) * cout) : 0.f;
#pragma unroll
            for (int u = 0; u < 16; ++u)
                if (i0 + u < parts) acc += v[u];
        }
    };
    add_set(partials + (size_t)inst * parts * cout + col);
    if (partials2 != nullptr) add_set(partials2 + (size_t)inst * parts * cout + col);
    out[g] = acc;
}

template <int NT, int STAGES, int FMT>
int launch_gemm(const PmGemmBatch& batch, int nprob, int rows, int cout, cudaStream_t st, int ninst) {
    using Cfg = GmCfg<NT, STAGES, FMT>;
    cudaError_t e = cudaFuncSetAttribute(pm_gemm_kernel<NT, STAGES, FMT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
    dim3 grid(nprob * ninst, cout / NT, rows / GM_BM);
    pm_gemm_kernel<NT, STAGES, FMT><<<grid, GM_THREADS, Cfg::SMEM_BYTES, st>>>(batch, nprob);
    return dcl_launch_status();
}

}  // namespace

DCL_API int dcl_pm_gemm(int nproblems, const dcl_pm_gemm_problem* problems, int rows, void* stream) {
    DCL_RETURN_IF_BAD(nproblems >= 1 && nproblems <= GM_MAX_PROBLEMS && problems != nullptr);
    DCL_RETURN_IF_BAD(rows > 0 && rows % GM_BM == 0 && rows / GM_BM <= 65535);
    PmGemmBatch batch;
    const int cout = problems[0].cout, nt = problems[0].nt, a_fmt = problems[0].a_fmt;
    const int ninst = problems[0].inst_count > 1 ? problems[0].inst_count : 1;
    DCL_RETURN_IF_BAD((nt == 64 || nt == 128 || nt == 256) && cout % nt == 0);
    DCL_RETURN_IF_BAD((long)nproblems * ninst <= 65535);
    DCL_RETURN_IF_BAD(a_fmt == FMT_BF16X2 || a_fmt == FMT_F16);
    for (int i = 0; i < nproblems; ++i) {
        const dcl_pm_gemm_problem& p = problems[i];
        DCL_RETURN_IF_BAD(p.cout == cout && p.nt == nt && p.kb_total >= 1 && p.kb0 >= 0 && p.kb0 <= p.kb_total);
        DCL_RETURN_IF_BAD(p.a_fmt == a_fmt && (p.out_fmt == FMT_BF16X2 || p.out_fmt == FMT_F16));
        // strided-batch problems (split-K slices with their own fp32 output): one operand image per slice, fp32 output only
        DCL_RETURN_IF_BAD((p.inst_count > 1 ? p.inst_count : 1) == ninst);
        DCL_RETURN_IF_BAD(ninst == 1 || (p.kb0 == p.kb_total && p.out_cm != nullptr && p.out_pm == nullptr &&
                                         p.pool_out == nullptr && p.dot_out == nullptr && p.out_qk == nullptr &&
                                         p.out_v == nullptr && p.a_inst_stride % 16 == 0 && p.w_inst_stride % 16 == 0 &&
                                         p.out_cm_inst_stride % 4 == 0));
        DCL_RETURN_IF_BAD(p.a0 != nullptr && p.w != nullptr && (p.kb0 == p.kb_total || p.a1 != nullptr));
        DCL_RETURN_IF_BAD((p.post_scale == nullptr) == (p.post_shift == nullptr));
        DCL_RETURN_IF_BAD(p.out_cm == nullptr || (p.rows_per_inst > 0 && p.rows_per_inst % 32 == 0 && rows % p.rows_per_inst == 0));
        DCL_RETURN_IF_BAD(p.pool_out == nullptr || p.pool_w != nullptr);
        DCL_RETURN_IF_BAD(p.dot_out == nullptr || (p.dot_w != nullptr && cout == nt));
        DCL_RETURN_IF_BAD(p.out_qk == nullptr || ((p.qk_tile_rows == 64 || p.qk_tile_rows == 128) && cout % 8 == 0 &&
                                                  ((uintptr_t)p.out_qk & 15u) == 0));
        DCL_RETURN_IF_BAD(p.out_v == nullptr || (p.v_row0 >= 0 && p.v_row0 % 8 == 0 && p.v_rows >= p.v_row0 + cout &&
                                                 p.v_rows % 8 == 0 && rows % 16 == 0 && ((uintptr_t)p.out_v & 15u) == 0));
        DCL_RETURN_IF_BAD(((((uintptr_t)p.a0) | ((uintptr_t)p.a1) | ((uintptr_t)p.w) | ((uintptr_t)p.out_pm)) & 15u) == 0);
        batch.p[i] = p;
    }
    cudaStream_t st = (cudaStream_t)stream;
    // Persistent kernels when the m-tiles pair up: the CTA-pair (cta_group::2) kernel by default,
    // DCL_PM_GEMM_MCAST=1 selects the multicast one (bf16 hi/lo operands only), DCL_PM_GEMM_SIMPLE=1 the simple
    // kernel (A/B runs).
    static const bool force_simple = getenv("DCL_PM_GEMM_SIMPLE") != nullptr;
    static const bool force_mcast = getenv("DCL_PM_GEMM_MCAST") != nullptr;
    const bool paired = (rows / GM_BM) % 2 == 0;
    if (a_fmt == FMT_F16) {
        if (!force_simple && paired) {
            if (nt == 256) return launch_gemm_pair<256, 8, FMT_F16>(batch, nproblems, rows, cout, st, ninst);
            if (nt == 128) return launch_gemm_pair<128, 12, FMT_F16>(batch, nproblems, rows, cout, st, ninst);
            return launch_gemm_pair<64, 12, FMT_F16>(batch, nproblems, rows, cout, st, ninst);
        }
        if (nt == 256) return launch_gemm<256, 2, FMT_F16>(batch, nproblems, rows, cout, st, ninst);
        if (nt == 128) return launch_gemm<128, 4, FMT_F16>(batch, nproblems, rows, cout, st, ninst);
        return launch_gemm<64, 6, FMT_F16>(batch, nproblems, rows, cout, st, ninst);
    }
    if (!force_simple && !force_mcast && paired) {
        if (nt == 256) return launch_gemm_pair<256, 6, FMT_BF16X2>(batch, nproblems, rows, cout, st, ninst);
        if (nt == 128) return launch_gemm_pair<128, 8, FMT_BF16X2>(batch, nproblems, rows, cout, st, ninst);
        return launch_gemm_pair<64, 9, FMT_BF16X2>(batch, nproblems, rows, cout, st, ninst);
    }
    if (!force_simple && paired && ninst == 1) {
        if (nt == 256) return launch_gemm_cluster<256, 4>(batch, nproblems, rows, cout, st);
        if (nt == 128) return launch_gemm_cluster<128, 6>(batch, nproblems, rows, cout, st);
        return launch_gemm_cluster<64, 8>(batch, nproblems, rows, cout, st);
    }
    if (nt == 256) return launch_gemm<256, 2, FMT_BF16X2>(batch, nproblems, rows, cout, st, ninst);
    if (nt == 128) return launch_gemm<128, 3, FMT_BF16X2>(batch, nproblems, rows, cout, st, ninst);
    return launch_gemm<64, 4, FMT_BF16X2>(batch, nproblems, rows, cout, st, ninst);
}

DCL_API int dcl_pm_pack_rows(int rows, int c, int ld, const float* src, void* dst_pm, int fmt, void* stream) {
    DCL_RETURN_IF_BAD(rows > 0 && rows % 128 == 0 && c > 0 && c % 32 == 0 && ld >= c && (fmt == 0 || fmt == 1));
    DCL_RETURN_IF_BAD((((uintptr_t)dst_pm) & 15u) == 0);
    const long total = (long)rows * (c / 8);
    pm_pack_rows_kernel<<<(unsigned)DCL_DIVUP(total, 256L), 256, 0, (cudaStream_t)stream>>>(
        rows, c, ld, src, reinterpret_cast<unsigned char*>(dst_pm), fmt);
    return dcl_launch_status();
}

DCL_API int dcl_pm_pack_cm(int b, int c, int n, const float* src, void* dst_pm, int fmt, void* stream) {
    DCL_RETURN_IF_BAD(b > 0 && n > 0 && ((long)b * n) % 128 == 0 && c > 0 && c % 32 == 0 && (fmt == 0 || fmt == 1));
    DCL_RETURN_IF_BAD((((uintptr_t)dst_pm) & 15u) == 0);
    const long total = (long)b * n * (c / 8);
    pm_pack_cm_kernel<<<(unsigned)DCL_DIVUP(total, 256L), 256, 0, (cudaStream_t)stream>>>(
        b, c, n, src, reinterpret_cast<unsigned char*>(dst_pm), fmt);
    return dcl_launch_status();
}

DCL_API int dcl_pm_unpack(int rows, int c, const void* src_pm, float* dst, int fmt, void* stream) {
    DCL_RETURN_IF_BAD(rows > 0 && rows % 128 == 0 && c > 0 && c % 32 == 0 && (fmt == 0 || fmt == 1));
    const long total = (long)rows * c;
    pm_unpack_kernel<<<(unsigned)DCL_DIVUP(total, 256L), 256, 0, (cudaStream_t)stream>>>(
        rows, c, reinterpret_cast<const unsigned char*>(src_pm), dst, fmt);
    return dcl_launch_status();
}

DCL_API int dcl_pm_pool_reduce(int insts, int cout, int parts, const float* partials, const float* partials2,
                               float* out, int accumulate, void* stream) {
    DCL_RETURN_IF_BAD(insts > 0 && cout > 0 && parts > 0 && partials != nullptr && out != nullptr);
    pm_pool_reduce_kernel<<<DCL_DIVUP(insts * cout, 256), 256, 0, (cudaStream_t)stream>>>(insts, cout, parts, partials,
                                                                                        partials2, out, accumulate);
    return dcl_launch_status();
}
