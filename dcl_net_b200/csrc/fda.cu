// Fused Feature Disengagement & Alignment (FDA) correspondence kernel.
//
// Reference (unfused, fp32, PyTorch library calls):
//   models/Modules.py:166-169   A = softmax(bmm(RI_2^T, RI_1), dim=1); RE_embed = bmm(RE_2, A)
//   models/DCL_Net.py:213,215   F_*_m = bmm(RI_2, A)
// i.e. per instance an (m x n) similarity of c-channel features, a softmax over the m
// (key) axis for every query column, and two products with that matrix.  The reference
// writes the (b,m,n) matrix to HBM and re-reads it three times.
//
// Here: one flash-style kernel.  A CTA owns 128 queries of one instance and streams the
// keys in blocks of 64:
//     S  = Q K^T            tcgen05.mma, fp32 accumulator in TMEM (2 x 64 columns, ping-pong)
//     P  = exp2(S*log2e - m) online softmax in registers (lazy rescale), written to smem
//     O += P [RE_2 ; RI_2]^T tcgen05.mma, fp32 accumulator in TMEM (p + c columns)
// The similarity matrix never leaves the SM.
//
// Precision.  north_star asks for bf16 operands / fp32 accumulation AND 1e-3 relative
// agreement with the fp32 reference.  Logits are unscaled dot products of post-ReLU
// features, so plain bf16 operands (2^-9) miss the tolerance.  Every operand x is split
// x = hi + lo (both bf16) and every product is evaluated as hi*hi + hi*lo + lo*hi
// (3 MMAs, error ~2^-17): fp32-faithful results from the bf16 tensor pipe.
//
// Operand staging.  The operands reach the kernel as bf16 hi/lo tile images laid out in HBM exactly as the UMMA
// no-swizzle ("interleave") shared-memory image, so every tile moves with 1-D TMA bulk copies (cp.async.bulk /
// UBLKCP) completing on an mbarrier.  In the inference path the disengage GEMMs write these images in their
// epilogue (pm_gemm.cu: out_qk / out_v); for the reference's fp32 channel-major interface a pre-pass
// (fda_pack_kernel) makes them.  Query / key tiles are K-major: a tile of R rows x K elements is a grid of
// 8x8-element "core matrices" (8 rows x 16 B, 128 B contiguous);
//               LBO = byte distance between core matrices adjacent in K,
//               SBO = byte distance between core matrices adjacent in the row direction
// (cute::UMMA::make_umma_desc<Major::K>, INTERLEAVE: ((8,n),2):((1,SBO),LBO) in uint128).  The value chunks are
// MN-major (a 16-byte unit = 8 value rows of one key; LBO = distance between 8-key groups, SBO = between 8-row
// chunks), which a point-major producer writes with 16-byte stores.
//
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = MMA issuer + TMEM owner (in the CTA-pair kernel the
// peer's warp 1 relays its load completions to the leader), warps 2-9 = softmax / correction / epilogue, two
// threads per query row (warps w and w+4 touch TMEM lanes 32*(w%4)..+31 and split the columns).
#include "common.cuh"
#include "umma.cuh"
#include "../../include/dcl_b200.h"
#include <cuda_fp16.h>
#include <math_constants.h>
#include <cstdlib>

namespace {

constexpr int QT = 128;  // queries per CTA (UMMA M)
constexpr int KB = 64;   // keys per block (UMMA N of the S product, K of the O product)
constexpr int KS = 16;   // keys per V chunk = one UMMA K step
constexpr int FDA_THREADS = 320;  // TMA warp, MMA warp, eight softmax warps
constexpr int FDA_P = 256;
constexpr float LOG2E = 1.4426950408889634f;
constexpr float RESCALE_TH = 8.0f;  // lazy rescale threshold (log2 units)

template <int C>
struct FdaCfg {
    static constexpr bool PV16 = false;               // P and V as bf16 hi/lo pairs (3 MMAs per P V product)
    static constexpr int VROWS = FDA_P + C;           // value rows: RE_2 then RI_2
    static constexpr int NK = (C == 64) ? 2 : 1;      // K block ring depth
    static constexpr int NV = (C == 64) ? 4 : 2;      // V chunk ring depth
    static constexpr int NP = 2;                      // P buffers (the o_done pairing below assumes 2)
    static constexpr int Q_HALF = QT * C * 2;         // bytes of the hi (or lo) image
    static constexpr int K_HALF = KB * C * 2;
    static constexpr int V_HALF = VROWS * KS * 2;
    static constexpr int P_HALF = QT * KB * 2;
    static constexpr int Q_BYTES = 2 * Q_HALF, K_BYTES = 2 * K_HALF, V_BYTES = 2 * V_HALF, P_BYTES = 2 * P_HALF;
    static constexpr int OFF_Q = 0;
    static constexpr int OFF_K = OFF_Q + Q_BYTES;
    static constexpr int OFF_V = OFF_K + NK * K_BYTES;
    static constexpr int OFF_P = OFF_V + NV * V_BYTES;
    static constexpr int OFF_BAR = OFF_P + NP * P_BYTES;
    static constexpr int OFF_XCH = OFF_BAR + 256;      // 6 x 128 floats exchanged between the two threads of a row
    static constexpr int SMEM_BYTES = OFF_XCH + 6 * QT * 4;
    static constexpr int S_COL = VROWS;               // TMEM: O at [0,VROWS), S ping-pong after it
    static constexpr int TMEM_COLS = 512;
    static_assert(VROWS + 2 * KB <= 512, "TMEM budget");
    static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
    // operand strides (bytes) of the packed images
    static constexpr int QK_LBO = 128, QK_SBO = (C / 8) * 128;
    static constexpr int V_LBO = 128, V_SBO = 256;
    static constexpr int P_LBO = 128, P_SBO = (KB / 8) * 128;
};

// ------------------------------------------------------------------ pack pre-pass
// Element (row, k) of an operand tile with `kchunks` 8-element chunks along K lives at
//   (row/8)*kchunks*64 + (k/8)*64 + (row%8)*8 + (k%8)      [bf16 elements]
// Q tile: rows = queries (128), K = channels.  K tile: rows = keys (64), K = channels.
// V chunk (16 keys): MN-major for the P V product — 16-byte units of 8 value rows of one key at
//   (vrow/8)*128 + ((key%16)/8)*64 + (key%8)*8 + (vrow%8)   [bf16 elements]
// which is also what a point-major producer (one thread per key) writes with 16-byte stores (dcl_pm_gemm out_v).
template <int C>
__global__ void __launch_bounds__(256) fda_pack_kernel(int n, int m, const float* __restrict__ RI_1,
                                                       const float* __restrict__ RI_2,
                                                       const float* __restrict__ RE_2,
                                                       __nv_bfloat16* __restrict__ Qp,
                                                       __nv_bfloat16* __restrict__ Kp,
                                                       __nv_bfloat16* __restrict__ Vp, int pv_fmt) {
    using Cfg = FdaCfg<C>;
    const int bs = blockIdx.y;
    const int section = blockIdx.z;  // 0: Q, 1: K, 2: V
    const int tid = blockIdx.x * 256 + threadIdx.x;
    if (section == 0 || section == 1) {
        // thread = (row, channel chunk); row fastest so global reads (fixed channel) coalesce
        const int rows_total = section == 0 ? n : m;
        const int tile_rows = section == 0 ? QT : KB;
        const int row = tid % rows_total, cchunk = tid / rows_total;
        if (cchunk >= C / 8) return;
        const float* src = (section == 0 ? RI_1 : RI_2) + ((size_t)bs * C + cchunk * 8) * rows_total + row;
        __nv_bfloat16 hi[8], lo[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) split_bf16(__ldg(src + (size_t)i * rows_total), hi[i], lo[i]);
        const int tile = row / tile_rows, r = row % tile_rows;
        const size_t half = (size_t)tile_rows * C;  // elements in one hi (or lo) image
        __nv_bfloat16* dst = (section == 0 ? Qp : Kp) + ((size_t)bs * (rows_total / tile_rows) + tile) * 2 * half +
                             (size_t)(r / 8) * (C / 8) * 64 + cchunk * 64 + (r % 8) * 8;
        *reinterpret_cast<uint4*>(dst) =
            make_uint4(pack2(hi[0], hi[1]), pack2(hi[2], hi[3]), pack2(hi[4], hi[5]), pack2(hi[6], hi[7]));
        *reinterpret_cast<uint4*>(dst + half) =
            make_uint4(pack2(lo[0], lo[1]), pack2(lo[2], lo[3]), pack2(lo[4], lo[5]), pack2(lo[6], lo[7]));
    } else {
        // thread = (key, chunk of 8 value rows); key fastest so the global reads (fixed value row) coalesce and each
        // group of 8 lanes writes 128 contiguous bytes.  The value image is MN-major for the P V product: a 16-byte
        // unit holds 8 value rows of ONE key.
        const int key = tid % m, nchunk = tid / m;
        if (nchunk >= Cfg::VROWS / 8) return;
        __nv_bfloat16 hi[8], lo[8];
        float val[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int vrow = nchunk * 8 + i;
            const float* src = (vrow < FDA_P) ? RE_2 + ((size_t)bs * FDA_P + vrow) * m
                                              : RI_2 + ((size_t)bs * C + (vrow - FDA_P)) * m;
            val[i] = __ldg(src + key);
            split_bf16(val[i], hi[i], lo[i]);
        }
        const size_t half = (size_t)Cfg::VROWS * KS;  // elements in one hi (or lo) image
        if (pv_fmt == 1) {
            // one fp16 image per chunk (values rounded once, clamped to the fp16 range)
            uint32_t w[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const __half2 hh = __floats2half2_rn(fminf(fmaxf(val[2 * e], -65504.f), 65504.f),
                                                     fminf(fmaxf(val[2 * e + 1], -65504.f), 65504.f));
                w[e] = *reinterpret_cast<const uint32_t*>(&hh);
            }
            __nv_bfloat16* d16 = Vp + ((size_t)bs * (m / KS) + key / KS) * half + (size_t)nchunk * 128 +
                                 ((key % KS) / 8) * 64 + (key % 8) * 8;
            *reinterpret_cast<uint4*>(d16) = make_uint4(w[0], w[1], w[2], w[3]);
            return;
        }
        __nv_bfloat16* dst = Vp + ((size_t)bs * (m / KS) + key / KS) * 2 * half + (size_t)nchunk * 128 +
                             ((key % KS) / 8) * 64 + (key % 8) * 8;
        *reinterpret_cast<uint4*>(dst) =
            make_uint4(pack2(hi[0], hi[1]), pack2(hi[2], hi[3]), pack2(hi[4], hi[5]), pack2(hi[6], hi[7]));
        *reinterpret_cast<uint4*>(dst + half) =
            make_uint4(pack2(lo[0], lo[1]), pack2(lo[2], lo[3]), pack2(lo[4], lo[5]), pack2(lo[6], lo[7]));
    }
}

// ------------------------------------------------------------------ timeline trace (bring-up / profiling)
// When a trace buffer is installed (dcl_debug_fda_set_trace), CTA (0,0) of the FDA kernels stamps clock64() at the
// hand-off points of its three roles: trace[role * 1024 + block * 8 + event].
__device__ long long* g_fda_trace = nullptr;
__device__ __forceinline__ long long fda_globaltimer() {
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// life-cycle stamps (globaltimer, ns) of CTA (0,0) [slot 0] and of the last leader CTA of the grid [slot 1]:
// trace[3 * 1024 + slot * 8 + event], events 0 entry, 1 set-up done, 2 main loop done, 3 epilogue done, 4 exit.
__device__ __forceinline__ long long* fda_life_slot() {
    long long* g = g_fda_trace;
    if (g == nullptr) return nullptr;
    if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) return g + 3 * 1024;
    if (blockIdx.x == ((gridDim.x - 1) & ~1u) && blockIdx.y == gridDim.y - 1 && blockIdx.z == gridDim.z - 1)
        return g + 3 * 1024 + 8;
    return nullptr;
}
__device__ __forceinline__ void fda_life(long long* slot, int ev) {
    if (slot != nullptr) slot[ev] = fda_globaltimer();
}
__device__ __forceinline__ void fda_stamp(long long* tr, int role, int j, int ev) {
    if (tr != nullptr && j < 128) tr[role * 1024 + j * 8 + ev] = clock64();
}

// Where a launch delivers its two results.  Each of RE_embed (256 channels) and RI_embed (C channels) can go out
// as the reference's fp32 channel-major tensor, as a point-major bf16 hi/lo image (dcl_pm_gemm's activation format,
// pm_gemm.cu) ready to be the next layer's A operand, or both; null pointers are skipped.
struct FdaOut {
    float* re_cm;
    float* ri_cm;
    unsigned char* re_pm;
    unsigned char* ri_pm;
    float* lse;
};

// One launch can serve several independent problems of equal shape (grid.z = job): the two directions of the dual
// FDA of models/DCL_Net.py:206-215 then share their partial last waves.
constexpr int FDA_MAX_JOBS = 2;
struct FdaJob {
    const __nv_bfloat16 *Qp, *Kp, *Vp;
    FdaOut out;
};
struct FdaJobs {
    FdaJob j[FDA_MAX_JOBS];
};

// ------------------------------------------------------------------ softmax / correction / epilogue warps
// PAIR: the barriers the MMA issuer waits on live in the leader CTA of the pair; every warp announces itself there
// with one cluster-scope arrive (local or remote).  Otherwise every thread arrives on its own CTA's barrier.
template <bool PAIR>
__device__ __forceinline__ void fda_arrive(uint64_t* bar, int lane) {
    if constexpr (PAIR) {
        __syncwarp();
        if (lane == 0) dcl_mbar_arrive_remote(bar, 0);
    } else {
        dcl_mbar_arrive(bar);
    }
}

// Eight warps, two threads per query row: warps w and w+4 own the same TMEM lane quadrant and split the 64 keys of a
// block (and the O columns in the rescale / epilogue) in halves, so every scheduler holds two softmax warps and the
// S -> P latency, which paces the tensor pipe, halves.  The two threads of a row agree on the running maximum through
// a double-buffered shared-memory slot and a 64-thread named barrier; their partial row sums meet in the epilogue.
template <class Cfg, bool PAIR>
__device__ __forceinline__ void fda_softmax_warps(unsigned char* smem, uint32_t tmem_base, int warp, int lane, int qt,
                                                  int bs, int n, int NB, uint64_t* s_full, uint64_t* s_empty,
                                                  uint64_t* o_done, uint64_t* p_full, const FdaOut& out) {
    constexpr int C = Cfg::VROWS - FDA_P;
    constexpr int HK = KB / 2;                   // keys per thread and block
    constexpr int OC = Cfg::VROWS / 32;          // 32-column chunks of O
    long long* tr = (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && warp == 2 && lane == 0) ? g_fda_trace : nullptr;
    long long* life = (warp == 2 && lane == 0) ? fda_life_slot() : nullptr;
    const int quad = warp & 3;
    const int half = (warp - 2) >> 2;
    const int row = quad * 32 + lane;  // query row within the tile == TMEM lane
    const uint32_t t_lane = (uint32_t)(quad * 32) << 16;
    float* xch = reinterpret_cast<float*>(smem + Cfg::OFF_XCH);  // [parity][half][row] partial maxima, then [half][row] sums
    const uint32_t pair_bar = 1u + (uint32_t)quad;
    float m_ref = -CUDART_INF_F, l = 0.f;
    unsigned char* p_row_base = smem + Cfg::OFF_P + (row >> 3) * Cfg::P_SBO + (row & 7) * 16 + half * (HK / 8) * Cfg::P_LBO;
    for (int j = 0; j < NB; ++j) {
        const int sb = j & 1;
        fda_stamp(tr, 1, j, 0);
        dcl_mbar_wait(s_full + sb, (uint32_t)((j >> 1) & 1));
        fda_stamp(tr, 1, j, 1);
        tc_fence_after();
        uint32_t sv[HK];
        DCL_TMEM_LD32(tmem_base + t_lane + Cfg::S_COL + sb * KB + half * HK, sv);
        tc_wait_ld();
        tc_fence_before();
        fda_arrive<PAIR>(s_empty + sb, lane);
        fda_stamp(tr, 1, j, 2);

        float mx0 = __uint_as_float(sv[0]), mx1 = __uint_as_float(sv[1]);
#pragma unroll
        for (int i = 2; i < HK; i += 2) {
            mx0 = fmaxf(mx0, __uint_as_float(sv[i]));
            mx1 = fmaxf(mx1, __uint_as_float(sv[i + 1]));
        }
        float mx = fmaxf(mx0, mx1);
        xch[(sb * 2 + half) * QT + row] = mx;
        asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
        mx = fmaxf(mx, xch[(sb * 2 + (half ^ 1)) * QT + row]) * LOG2E;
        float alpha = 1.f;
        bool need = false;
        if (j == 0) {
            m_ref = mx;
        } else if (mx > m_ref + RESCALE_TH) {
            alpha = ex2_approx(m_ref - mx);
            m_ref = mx;
            need = true;
        }
        // P buffer j%NP was last read by PV(j-NP)
        fda_stamp(tr, 1, j, 3);
        if (j >= 2) dcl_mbar_wait(o_done + (j & 1), (uint32_t)(((j >> 1) - 1) & 1));
        fda_stamp(tr, 1, j, 4);
        {
            unsigned char* pd = p_row_base + (j % Cfg::NP) * Cfg::P_BYTES;
            // The row sum is taken over the weights the tensor core will actually see (hi + lo), so that
            // numerator and denominator of the softmax carry the same rounding.
            float sum0 = 0.f, sum1 = 0.f;
            const float neg_m = -m_ref;
#pragma unroll
            for (int kc = 0; kc < HK / 8; ++kc) {
                uint32_t h[4], lw[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float p0 = ex2_approx(__fmaf_rn(__uint_as_float(sv[kc * 8 + 2 * e]), LOG2E, neg_m));
                    const float p1 = ex2_approx(__fmaf_rn(__uint_as_float(sv[kc * 8 + 2 * e + 1]), LOG2E, neg_m));
                    if constexpr (Cfg::PV16) {
                        // p <= 2^RESCALE_TH = 256 between rescales: inside the fp16 range
                        const __half2 hp = __floats2half2_rn(p0, p1);
                        h[e] = *reinterpret_cast<const uint32_t*>(&hp);
                        const float2 back = __half22float2(hp);
                        sum0 += back.x;
                        sum1 += back.y;
                    } else {
                        split2_bf16(p0, p1, h[e], lw[e]);
                        sum0 += __uint_as_float(h[e] << 16) + __uint_as_float(lw[e] << 16);
                        sum1 += __uint_as_float(h[e] & 0xffff0000u) + __uint_as_float(lw[e] & 0xffff0000u);
                    }
                }
                *reinterpret_cast<uint4*>(pd + kc * Cfg::P_LBO) = make_uint4(h[0], h[1], h[2], h[3]);
                if constexpr (!Cfg::PV16)
                    *reinterpret_cast<uint4*>(pd + Cfg::P_HALF + kc * Cfg::P_LBO) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
            }
            l = __fmaf_rn(l, alpha, sum0 + sum1);
        }
        if (j >= 1) {
            if (__any_sync(0xffffffffu, need)) {
                // O is rescaled between PV(j-1) and PV(j); this thread takes its half of the columns
                dcl_mbar_wait(o_done + ((j - 1) & 1), (uint32_t)(((j - 1) >> 1) & 1));
                tc_fence_after();
#pragma unroll 1
                for (int cc = half * (OC / 2); cc < (half + 1) * (OC / 2); ++cc) {
                    uint32_t ov[32];
                    const uint32_t ta = tmem_base + t_lane + cc * 32;
                    DCL_TMEM_LD32(ta, ov);
                    tc_wait_ld();
#pragma unroll
                    for (int i = 0; i < 32; ++i) ov[i] = __float_as_uint(__uint_as_float(ov[i]) * alpha);
                    DCL_TMEM_ST32(ta, ov);
                }
                tc_wait_st();
            }
        }
        dcl_fence_proxy_async();
        tc_fence_before();
        fda_arrive<PAIR>(p_full + (j % Cfg::NP), lane);
        fda_stamp(tr, 1, j, 5);
    }
    fda_life(life, 2);
    // ---- epilogue: O / l -> global; the two threads of a row first add up their partial sums
    float* lx = xch + 4 * QT;
    lx[half * QT + row] = l;
    asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
    l += lx[(half ^ 1) * QT + row];
    dcl_mbar_wait(o_done + ((NB - 1) & 1), (uint32_t)(((NB - 1) >> 1) & 1));
    tc_fence_after();
    const float inv_l = 1.0f / l;
    const int qglob = qt * QT + row;
    const size_t tile = (size_t)bs * (n / QT) + qt;  // 128-row tile of the (b*n)-row point-major images
    const uint32_t pm_row = (uint32_t)(row >> 3) * 512u + (uint32_t)(row & 7) * 16u;
#pragma unroll 1
    for (int cc = half * (OC / 2); cc < (half + 1) * (OC / 2); ++cc) {
        uint32_t ov[32];
        DCL_TMEM_LD32(tmem_base + t_lane + cc * 32, ov);
        tc_wait_ld();
        const bool is_re = cc < FDA_P / 32;
        const int kb = is_re ? cc : cc - FDA_P / 32;  // 32-channel block within its tensor
        float y[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) y[i] = __uint_as_float(ov[i]) * inv_l;
        float* cm = is_re ? out.re_cm : out.ri_cm;
        if (cm != nullptr) {
            // channel-major: consecutive lanes = consecutive queries, 512 contiguous bytes per channel and CTA
            float* o = cm + ((size_t)bs * (is_re ? FDA_P : C) + kb * 32) * n + qglob;
#pragma unroll
            for (int i = 0; i < 32; ++i) o[(size_t)i * n] = y[i];
        }
        unsigned char* pm = is_re ? out.re_pm : out.ri_pm;
        if constexpr (Cfg::PV16) {
            if (pm != nullptr) {
                // PM16: one whole 8 KB blob (128 rows x 32 channels, fp16) per CTA and iteration
                unsigned char* blob = pm + (tile * ((is_re ? FDA_P : C) / 32) + kb) * 8192 + pm_row;
#pragma unroll
                for (int ch = 0; ch < 4; ++ch) {
                    uint32_t h[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float a0 = fminf(fmaxf(y[ch * 8 + 2 * e], -65504.f), 65504.f);
                        const float a1 = fminf(fmaxf(y[ch * 8 + 2 * e + 1], -65504.f), 65504.f);
                        const __half2 hh = __floats2half2_rn(a0, a1);
                        h[e] = *reinterpret_cast<const uint32_t*>(&hh);
                    }
                    *reinterpret_cast<uint4*>(blob + ch * 128) = make_uint4(h[0], h[1], h[2], h[3]);
                }
            }
        } else if (pm != nullptr) {
            // one whole 16 KB blob (128 rows x 32 channels, hi | lo) per CTA and iteration
            unsigned char* blob = pm + (tile * ((is_re ? FDA_P : C) / 32) + kb) * 16384 + pm_row;
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) {
                uint32_t h[4], lw[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) split2_bf16(y[ch * 8 + 2 * e], y[ch * 8 + 2 * e + 1], h[e], lw[e]);
                *reinterpret_cast<uint4*>(blob + ch * 128) = make_uint4(h[0], h[1], h[2], h[3]);
                *reinterpret_cast<uint4*>(blob + 8192 + ch * 128) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
            }
        }
    }
    if (half == 0 && out.lse != nullptr) out.lse[(size_t)bs * n + qglob] = m_ref * (1.0f / LOG2E) + logf(l);
    tc_fence_before();
    fda_life(life, 3);
}

// ------------------------------------------------------------------ main kernel
template <int C>
__global__ void __launch_bounds__(FDA_THREADS, 1) fda_fwd_kernel(int n, int m,
                                                                 const __grid_constant__ FdaJobs jobs) {
    using Cfg = FdaCfg<C>;
    const FdaJob& job = jobs.j[blockIdx.z];
    const __nv_bfloat16 *Qp = job.Qp, *Kp = job.Kp, *Vp = job.Vp;
    const FdaOut& out = job.out;
    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);
    uint64_t* q_full = bars + 0;
    uint64_t* k_full = bars + 1;   // [2]
    uint64_t* k_empty = bars + 3;  // [2]
    uint64_t* s_full = bars + 5;   // [2]
    uint64_t* s_empty = bars + 7;  // [2]
    uint64_t* o_done = bars + 9;   // [2]: PV(j) commits to o_done[j & 1], so a waiter is never more than one phase behind
    uint64_t* p_full = bars + 11;  // [NP] (<= 2)
    uint64_t* v_full = bars + 13;  // [NV] (<= 4)
    uint64_t* v_empty = bars + 17; // [NV]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 21);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int qt = blockIdx.x, bs = blockIdx.y;
    const int NB = m / KB;  // key blocks
    long long* life0 = threadIdx.x == 0 ? fda_life_slot() : nullptr;
    fda_life(life0, 0);

    if (threadIdx.x == 0) {
        dcl_mbar_init(q_full, 1);
        for (int i = 0; i < 2; ++i) {
            dcl_mbar_init(k_full + i, 1);
            dcl_mbar_init(k_empty + i, 1);
            dcl_mbar_init(s_full + i, 1);
            dcl_mbar_init(s_empty + i, 256);
        }
        dcl_mbar_init(o_done, 1);
        dcl_mbar_init(o_done + 1, 1);
        for (int i = 0; i < Cfg::NP; ++i) dcl_mbar_init(p_full + i, 256);
        for (int i = 0; i < Cfg::NV; ++i) {
            dcl_mbar_init(v_full + i, 1);
            dcl_mbar_init(v_empty + i, 1);
        }
        dcl_fence_barrier_init();
    }
    if (warp == 1) tc_alloc(tmem_slot, Cfg::TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    fda_life(life0, 1);

    const uint32_t sQ = dcl_smem_u32(smem + Cfg::OFF_Q);
    const uint32_t sK = dcl_smem_u32(smem + Cfg::OFF_K);
    const uint32_t sV = dcl_smem_u32(smem + Cfg::OFF_V);
    const uint32_t sP = dcl_smem_u32(smem + Cfg::OFF_P);

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (dcl_elect_one()) {
            const unsigned char* gQ =
                reinterpret_cast<const unsigned char*>(Qp) + ((size_t)bs * (n / QT) + qt) * Cfg::Q_BYTES;
            const unsigned char* gK = reinterpret_cast<const unsigned char*>(Kp) + (size_t)bs * NB * Cfg::K_BYTES;
            const unsigned char* gV =
                reinterpret_cast<const unsigned char*>(Vp) + (size_t)bs * (m / KS) * Cfg::V_BYTES;
            long long* trp = (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) ? g_fda_trace : nullptr;
            dcl_mbar_arrive_expect_tx(q_full, Cfg::Q_BYTES);
            dcl_bulk_g2s(smem + Cfg::OFF_Q, gQ, Cfg::Q_BYTES, q_full);
            auto load_k = [&](int j) {
                const int s = j % Cfg::NK;
                if (j >= Cfg::NK) dcl_mbar_wait(k_empty + s, (uint32_t)((j / Cfg::NK - 1) & 1));
                dcl_mbar_arrive_expect_tx(k_full + s, Cfg::K_BYTES);
                dcl_bulk_g2s(smem + Cfg::OFF_K + s * Cfg::K_BYTES, gK + (size_t)j * Cfg::K_BYTES, Cfg::K_BYTES,
                             k_full + s);
            };
            for (int j = 0; j < Cfg::NK && j < NB; ++j) load_k(j);
            int vi = 0;  // running V chunk counter
            for (int j = 0; j < NB; ++j) {
                // K block j+NK reuses the slot of block j, free once S(j) has completed; S(j) is issued ahead of
                // PV(j-1), so this wait never depends on the V chunks issued below.
                if (j + Cfg::NK < NB) load_k(j + Cfg::NK);
                for (int ks = 0; ks < KB / KS; ++ks, ++vi) {
                    const int s = vi % Cfg::NV;
                    const int use = vi / Cfg::NV;
                    if (ks == 0) fda_stamp(trp, 2, j, 0);
                    if (use >= 1) dcl_mbar_wait(v_empty + s, (uint32_t)((use - 1) & 1));
                    if (ks == 0) fda_stamp(trp, 2, j, 1);
                    if (ks == 3) fda_stamp(trp, 2, j, 2);
                    dcl_mbar_arrive_expect_tx(v_full + s, Cfg::V_BYTES);
                    dcl_bulk_g2s(smem + Cfg::OFF_V + s * Cfg::V_BYTES, gV + (size_t)vi * Cfg::V_BYTES, Cfg::V_BYTES,
                                 v_full + s);
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (dcl_elect_one()) {
            long long* tr = (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) ? g_fda_trace : nullptr;
            constexpr uint32_t idesc_s = umma_idesc_bf16(QT, KB);
            constexpr uint32_t idesc_o1 = umma_idesc_bf16(QT, 256, true);   // value image is MN-major
            constexpr uint32_t idesc_o2 = umma_idesc_bf16(QT, C, true);
            const uint32_t tO = tmem_base;
            // Descriptors differ only in their 14-bit start-address field: build one per operand image and bump it.
            const uint64_t dQh = umma_desc(sQ, Cfg::QK_LBO, Cfg::QK_SBO);
            const uint64_t dQl = umma_desc(sQ + Cfg::Q_HALF, Cfg::QK_LBO, Cfg::QK_SBO);
            const uint64_t dK0 = umma_desc(sK, Cfg::QK_LBO, Cfg::QK_SBO);
            const uint64_t dV0 = umma_desc(sV, Cfg::V_LBO, Cfg::V_SBO);
            const uint64_t dP0 = umma_desc(sP, Cfg::P_LBO, Cfg::P_SBO);
            auto issue_s = [&](int j) {
                const int s = j % Cfg::NK;
                fda_stamp(tr, 0, j, 0);
                dcl_mbar_wait(k_full + s, (uint32_t)((j / Cfg::NK) & 1));
                fda_stamp(tr, 0, j, 1);
                if (j >= 2) dcl_mbar_wait(s_empty + (j & 1), (uint32_t)(((j >> 1) - 1) & 1));
                fda_stamp(tr, 0, j, 2);
                tc_fence_after();
                const uint64_t dKh = dK0 + (uint64_t)((s * Cfg::K_BYTES) >> 4);
                const uint64_t dKl = dKh + (uint64_t)(Cfg::K_HALF >> 4);
                const uint32_t tS = tmem_base + Cfg::S_COL + (j & 1) * KB;
#pragma unroll
                for (int kk = 0; kk < C / 16; ++kk) {
                    const uint64_t off = (uint64_t)((kk * 2 * Cfg::QK_LBO) >> 4);  // two K chunks per UMMA
                    tc_mma_bf16(tS, dQh + off, dKh + off, idesc_s, kk == 0 ? 0u : 1u);
                    tc_mma_bf16(tS, dQh + off, dKl + off, idesc_s, 1u);
                    tc_mma_bf16(tS, dQl + off, dKh + off, idesc_s, 1u);
                }
                tc_commit(s_full + (j & 1));
                tc_commit(k_empty + s);
            };
            dcl_mbar_wait(q_full, 0);
            issue_s(0);
            int vi = 0;
            for (int j = 0; j < NB; ++j) {
                if (j + 1 < NB) issue_s(j + 1);
                const int ps = j % Cfg::NP;
                fda_stamp(tr, 0, j, 3);
                dcl_mbar_wait(p_full + ps, (uint32_t)((j / Cfg::NP) & 1));
                fda_stamp(tr, 0, j, 4);
                tc_fence_after();
                const uint64_t dPh = dP0 + (uint64_t)((ps * Cfg::P_BYTES) >> 4);
#pragma unroll
                for (int ks = 0; ks < KB / KS; ++ks, ++vi) {
                    const int s = vi % Cfg::NV;
                    dcl_mbar_wait(v_full + s, (uint32_t)((vi / Cfg::NV) & 1));
                    if (ks == 0) fda_stamp(tr, 0, j, 6);
                    if (ks == 3) fda_stamp(tr, 0, j, 7);
                    tc_fence_after();
                    const uint64_t dVh = dV0 + (uint64_t)((s * Cfg::V_BYTES) >> 4);
                    const uint64_t dVl = dVh + (uint64_t)(Cfg::V_HALF >> 4);
                    const uint64_t dAh = dPh + (uint64_t)((ks * 2 * Cfg::P_LBO) >> 4);
                    const uint64_t dAl = dAh + (uint64_t)(Cfg::P_HALF >> 4);
                    const uint32_t acc = (j == 0 && ks == 0) ? 0u : 1u;
                    // value rows [0,256) -> O columns [0,256)
                    tc_mma_bf16(tO, dAh, dVh, idesc_o1, acc);
                    tc_mma_bf16(tO, dAh, dVl, idesc_o1, 1u);
                    tc_mma_bf16(tO, dAl, dVh, idesc_o1, 1u);
                    // value rows [256,256+C) -> O columns [256,256+C)
                    constexpr uint64_t v2 = (uint64_t)(((256 / 8) * Cfg::V_SBO) >> 4);
                    tc_mma_bf16(tO + 256, dAh, dVh + v2, idesc_o2, acc);
                    tc_mma_bf16(tO + 256, dAh, dVl + v2, idesc_o2, 1u);
                    tc_mma_bf16(tO + 256, dAl, dVh + v2, idesc_o2, 1u);
                    tc_commit(v_empty + s);
                }
                tc_commit(o_done + (j & 1));
                fda_stamp(tr, 0, j, 5);
            }
        }
    } else {
        fda_softmax_warps<Cfg, false>(smem, tmem_base, warp, lane, qt, bs, n, NB, s_full, s_empty, o_done, p_full, out);
    }
    __syncwarp();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tc_dealloc(tmem_base, Cfg::TMEM_COLS);
    }
    fda_life(life0, 4);
}

// ------------------------------------------------------------------ CTA-pair kernel
// Two query tiles of one instance share every K block and V chunk: the pair runs each product as ONE M=256
// tcgen05.mma.cta_group::2 issued by the leader (cluster rank 0); CTA r stages only its half of the B operand —
// keys [32r, 32r+32) of a block, value rows [128r, 128r+128) and [256 + rC/2, 256 + (r+1)C/2) of a chunk — so the
// L2 -> shared-memory traffic per SM halves (the single-CTA kernel is bound by it: 2.1 MB per CTA against ~43 B/clk/SM
// of L2 bandwidth) and the freed shared memory deepens the V ring.  Conventions pinned by dcl_debug_umma_pair_gemm.
//   * Operand loads are issued by each CTA for its own half; the peer's MMA warp relays "my half landed"
//     to the leader's full barriers (count 2 there) in the order the leader consumes them.
//   * tcgen05.commit multicasts every completion (S ready, K/V slot free, PV done) to both CTAs.
//   * Softmax warps of both CTAs announce "S read" / "P written" with one cluster-scope arrive per warp on the
//     leader's barriers (count 8).
// PV16_: the P V products run on once-rounded fp16 operands — P = exp2(..) in (0,1] and the values (activations that
// are held in fp16 everywhere on the inference path) — as ONE MMA each instead of three; the value image then holds
// one fp16 image per 16-key chunk and the point-major outputs are PM16 images.  The logits (Q K^T) keep split
// operands.  Measured cost: profiles/r02_precision_emulation_*.json ("P single fp16": 3e-5 relative on the features).
template <int C, bool PV16_ = false>
struct FdaPairCfg {
    static constexpr bool PV16 = PV16_;
    static constexpr int VROWS = FDA_P + C;
    static constexpr int VH = VROWS / 2;               // value rows staged per CTA
    static constexpr int NK = 2;
    static constexpr int NV = PV16_ ? 8 : ((C == 64) ? 8 : 5);
    static constexpr int NP = 2;
    static constexpr int Q_HALF = QT * C * 2;
    static constexpr int K_HALF = (KB / 2) * C * 2;    // hi (or lo) image of this CTA's 32 keys
    static constexpr int V_HALF = VH * KS * 2;         // hi (or lo) image of this CTA's value rows, one chunk
    static constexpr int V_PART_A = 128 * KS * 2;      // rows of the N=256 product come first, then the N=C product's
    static constexpr int P_HALF = QT * KB * 2;
    static constexpr int PV_IMAGES = PV16_ ? 1 : 2;    // images per P buffer / V chunk
    static constexpr int Q_BYTES = 2 * Q_HALF, K_BYTES = 2 * K_HALF, V_BYTES = PV_IMAGES * V_HALF,
                         P_BYTES = PV_IMAGES * P_HALF;
    static constexpr int G_K_HALF = KB * C * 2, G_K_BYTES = 2 * G_K_HALF;         // images written by fda_pack_kernel
    static constexpr int G_V_HALF = VROWS * KS * 2, G_V_BYTES = PV_IMAGES * G_V_HALF;
    static constexpr int OFF_Q = 0;
    static constexpr int OFF_K = OFF_Q + Q_BYTES;
    static constexpr int OFF_V = OFF_K + NK * K_BYTES;
    static constexpr int OFF_P = OFF_V + NV * V_BYTES;
    static constexpr int OFF_BAR = OFF_P + NP * P_BYTES;
    static constexpr int OFF_XCH = OFF_BAR + 512;
    static constexpr int SMEM_BYTES = OFF_XCH + 6 * QT * 4;
    static constexpr int S_COL = VROWS;
    static constexpr int TMEM_COLS = 512;
    static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
    static constexpr int QK_LBO = 128, QK_SBO = (C / 8) * 128;
    static constexpr int V_LBO = 128, V_SBO = 256;
    static constexpr int P_LBO = 128, P_SBO = (KB / 8) * 128;
};

template <int C, bool PV16>
__global__ void __launch_bounds__(FDA_THREADS, 1) fda_pair_kernel(int n, int m,
                                                                  const __grid_constant__ FdaJobs jobs) {
    using Cfg = FdaPairCfg<C, PV16>;
    const FdaJob& job = jobs.j[blockIdx.z];
    const __nv_bfloat16 *Qp = job.Qp, *Kp = job.Kp, *Vp = job.Vp;
    const FdaOut& out = job.out;
    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);
    uint64_t* q_full = bars + 0;
    uint64_t* k_full = bars + 1;    // [2]
    uint64_t* k_empty = bars + 3;   // [2]
    uint64_t* s_full = bars + 5;    // [2]
    uint64_t* s_empty = bars + 7;   // [2]
    uint64_t* o_done = bars + 9;    // [2]
    uint64_t* p_full = bars + 11;   // [2]
    uint64_t* v_full = bars + 13;   // [NV] (<= 8)
    uint64_t* v_empty = bars + 21;  // [NV]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 29);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = dcl_cluster_ctarank();
    const bool leader = rank == 0;
    const int qt = blockIdx.x, bs = blockIdx.y;  // cluster = blocks (2i, 2i+1) along x
    const int NB = m / KB;
    long long* life0 = threadIdx.x == 0 ? fda_life_slot() : nullptr;
    fda_life(life0, 0);

    if (threadIdx.x == 0) {
        const uint32_t both = leader ? 2u : 1u;  // leader: own producer + the peer's relay
        dcl_mbar_init(q_full, both);
        for (int i = 0; i < 2; ++i) {
            dcl_mbar_init(k_full + i, both);
            dcl_mbar_init(k_empty + i, 1);
            dcl_mbar_init(s_full + i, 1);
            dcl_mbar_init(s_empty + i, 16);
            dcl_mbar_init(o_done + i, 1);
            dcl_mbar_init(p_full + i, 16);
        }
        for (int i = 0; i < Cfg::NV; ++i) {
            dcl_mbar_init(v_full + i, both);
            dcl_mbar_init(v_empty + i, 1);
        }
        dcl_fence_barrier_init();
    }
    __syncthreads();
    dcl_cluster_sync();  // both CTAs' barriers exist before anything is signalled at them
    if (warp == 1) tc2_alloc(tmem_slot, Cfg::TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    fda_life(life0, 1);

    const uint32_t sQ = dcl_smem_u32(smem + Cfg::OFF_Q);
    const uint32_t sK = dcl_smem_u32(smem + Cfg::OFF_K);
    const uint32_t sV = dcl_smem_u32(smem + Cfg::OFF_V);
    const uint32_t sP = dcl_smem_u32(smem + Cfg::OFF_P);

    if (warp == 0) {
        // ===================== TMA producer (each CTA: its Q tile, its halves of K and V) =====================
        if (dcl_elect_one()) {
            const unsigned char* gQ =
                reinterpret_cast<const unsigned char*>(Qp) + ((size_t)bs * (n / QT) + qt) * Cfg::Q_BYTES;
            const unsigned char* gK = reinterpret_cast<const unsigned char*>(Kp) + (size_t)bs * NB * Cfg::G_K_BYTES +
                                      rank * Cfg::K_HALF;
            const unsigned char* gV =
                reinterpret_cast<const unsigned char*>(Vp) + (size_t)bs * (m / KS) * Cfg::G_V_BYTES;
            const unsigned char* gVa = gV + (size_t)rank * Cfg::V_PART_A;                       // rows [128r, +128)
            const unsigned char* gVb = gV + (size_t)(256 + rank * (C / 2)) * (KS * 2);           // rows [256 + rC/2, ..)
            constexpr uint32_t V_PART_B = Cfg::V_HALF - Cfg::V_PART_A;
            long long* trp = (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) ? g_fda_trace : nullptr;
            dcl_mbar_arrive_expect_tx(q_full, Cfg::Q_BYTES);
            dcl_bulk_g2s(smem + Cfg::OFF_Q, gQ, Cfg::Q_BYTES, q_full);
            auto load_k = [&](int j) {
                const int s = j % Cfg::NK;
                if (j >= Cfg::NK) dcl_mbar_wait(k_empty + s, (uint32_t)((j / Cfg::NK - 1) & 1));
                dcl_mbar_arrive_expect_tx(k_full + s, Cfg::K_BYTES);
                unsigned char* dst = smem + Cfg::OFF_K + s * Cfg::K_BYTES;
                const unsigned char* src = gK + (size_t)j * Cfg::G_K_BYTES;
                dcl_bulk_g2s(dst, src, Cfg::K_HALF, k_full + s);
                dcl_bulk_g2s(dst + Cfg::K_HALF, src + Cfg::G_K_HALF, Cfg::K_HALF, k_full + s);
            };
            for (int j = 0; j < Cfg::NK && j < NB; ++j) load_k(j);
            int vi = 0;
            for (int j = 0; j < NB; ++j) {
                if (j + Cfg::NK < NB) load_k(j + Cfg::NK);
                for (int ks = 0; ks < KB / KS; ++ks, ++vi) {
                    const int s = vi % Cfg::NV;
                    const int use = vi / Cfg::NV;
                    if (ks == 0) fda_stamp(trp, 2, j, 0);
                    if (use >= 1) dcl_mbar_wait(v_empty + s, (uint32_t)((use - 1) & 1));
                    if (ks == 0) fda_stamp(trp, 2, j, 1);
                    if (ks == 3) fda_stamp(trp, 2, j, 2);
                    dcl_mbar_arrive_expect_tx(v_full + s, Cfg::V_BYTES);
                    unsigned char* dst = smem + Cfg::OFF_V + s * Cfg::V_BYTES;
                    const size_t g = (size_t)vi * Cfg::G_V_BYTES;
                    dcl_bulk_g2s(dst, gVa + g, Cfg::V_PART_A, v_full + s);
                    dcl_bulk_g2s(dst + Cfg::V_PART_A, gVb + g, V_PART_B, v_full + s);
                    if constexpr (!PV16) {
                        dcl_bulk_g2s(dst + Cfg::V_HALF, gVa + g + Cfg::G_V_HALF, Cfg::V_PART_A, v_full + s);
                        dcl_bulk_g2s(dst + Cfg::V_HALF + Cfg::V_PART_A, gVb + g + Cfg::G_V_HALF, V_PART_B, v_full + s);
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (leader && dcl_elect_one()) {
            // ===================== MMA issuer (leader only) =====================
            long long* tr = (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) ? g_fda_trace : nullptr;
            constexpr uint32_t idesc_s = umma_idesc_bf16(2 * QT, KB);
            // value image is MN-major; PV16: fp16 x fp16 (a_format = b_format = 0)
            constexpr uint32_t f16_mask = PV16 ? ~((1u << 7) | (1u << 10)) : ~0u;
            constexpr uint32_t idesc_o1 = umma_idesc_bf16(2 * QT, 256, true) & f16_mask;
            constexpr uint32_t idesc_o2 = umma_idesc_bf16(2 * QT, C, true) & f16_mask;
            const uint32_t tO = tmem_base;
            const uint64_t dQh = umma_desc(sQ, Cfg::QK_LBO, Cfg::QK_SBO);
            const uint64_t dQl = dQh + (uint64_t)(Cfg::Q_HALF >> 4);
            const uint64_t dK0 = umma_desc(sK, Cfg::QK_LBO, Cfg::QK_SBO);
            const uint64_t dV0 = umma_desc(sV, Cfg::V_LBO, Cfg::V_SBO);
            const uint64_t dP0 = umma_desc(sP, Cfg::P_LBO, Cfg::P_SBO);
            auto issue_s = [&](int j) {
                const int s = j % Cfg::NK;
                fda_stamp(tr, 0, j, 0);
                dcl_mbar_wait_cluster(k_full + s, (uint32_t)((j / Cfg::NK) & 1));
                fda_stamp(tr, 0, j, 1);
                if (j >= 2) dcl_mbar_wait_cluster(s_empty + (j & 1), (uint32_t)(((j >> 1) - 1) & 1));
                fda_stamp(tr, 0, j, 2);
                tc_fence_after();
                const uint64_t dKh = dK0 + (uint64_t)((s * Cfg::K_BYTES) >> 4);
                const uint64_t dKl = dKh + (uint64_t)(Cfg::K_HALF >> 4);
                const uint32_t tS = tmem_base + Cfg::S_COL + (j & 1) * KB;
#pragma unroll
                for (int kk = 0; kk < C / 16; ++kk) {
                    const uint64_t off = (uint64_t)((kk * 2 * Cfg::QK_LBO) >> 4);
                    tc2_mma_bf16(tS, dQh + off, dKh + off, idesc_s, kk == 0 ? 0u : 1u);
                    tc2_mma_bf16(tS, dQh + off, dKl + off, idesc_s, 1u);
                    tc2_mma_bf16(tS, dQl + off, dKh + off, idesc_s, 1u);
                }
                tc2_commit_mcast(s_full + (j & 1), (uint16_t)0x3);
                tc2_commit_mcast(k_empty + s, (uint16_t)0x3);
            };
            dcl_mbar_wait_cluster(q_full, 0);
            issue_s(0);
            int vi = 0;
            for (int j = 0; j < NB; ++j) {
                if (j + 1 < NB) issue_s(j + 1);
                const int ps = j & 1;
                fda_stamp(tr, 0, j, 3);
                dcl_mbar_wait_cluster(p_full + ps, (uint32_t)((j >> 1) & 1));
                fda_stamp(tr, 0, j, 4);
                tc_fence_after();
                const uint64_t dPh = dP0 + (uint64_t)((ps * Cfg::P_BYTES) >> 4);
#pragma unroll
                for (int ks = 0; ks < KB / KS; ++ks, ++vi) {
                    const int s = vi % Cfg::NV;
                    dcl_mbar_wait_cluster(v_full + s, (uint32_t)((vi / Cfg::NV) & 1));
                    if (ks == 0) fda_stamp(tr, 0, j, 6);
                    if (ks == 3) fda_stamp(tr, 0, j, 7);
                    tc_fence_after();
                    const uint64_t dVh = dV0 + (uint64_t)((s * Cfg::V_BYTES) >> 4);
                    const uint64_t dVl = dVh + (uint64_t)(Cfg::V_HALF >> 4);
                    const uint64_t dAh = dPh + (uint64_t)((ks * 2 * Cfg::P_LBO) >> 4);
                    const uint64_t dAl = dAh + (uint64_t)(Cfg::P_HALF >> 4);
                    const uint32_t acc = (j == 0 && ks == 0) ? 0u : 1u;
                    constexpr uint64_t v2 = (uint64_t)(Cfg::V_PART_A >> 4);
                    tc2_mma_bf16(tO, dAh, dVh, idesc_o1, acc);
                    if constexpr (!PV16) {
                        tc2_mma_bf16(tO, dAh, dVl, idesc_o1, 1u);
                        tc2_mma_bf16(tO, dAl, dVh, idesc_o1, 1u);
                    }
                    tc2_mma_bf16(tO + 256, dAh, dVh + v2, idesc_o2, acc);
                    if constexpr (!PV16) {
                        tc2_mma_bf16(tO + 256, dAh, dVl + v2, idesc_o2, 1u);
                        tc2_mma_bf16(tO + 256, dAl, dVh + v2, idesc_o2, 1u);
                    }
                    tc2_commit_mcast(v_empty + s, (uint16_t)0x3);
                }
                tc2_commit_mcast(o_done + (j & 1), (uint16_t)0x3);
                fda_stamp(tr, 0, j, 5);
            }
        } else if (!leader && dcl_elect_one()) {
            // ===================== relay (peer): "my half landed", in the leader's consumption order =====================
            dcl_mbar_wait(q_full, 0);
            dcl_mbar_arrive_remote(q_full, 0);
            auto relay_k = [&](int j) {
                const int s = j % Cfg::NK;
                dcl_mbar_wait(k_full + s, (uint32_t)((j / Cfg::NK) & 1));
                dcl_mbar_arrive_remote(k_full + s, 0);
            };
            relay_k(0);
            int vi = 0;
            for (int j = 0; j < NB; ++j) {
                if (j + 1 < NB) relay_k(j + 1);
                for (int ks = 0; ks < KB / KS; ++ks, ++vi) {
                    const int s = vi % Cfg::NV;
                    dcl_mbar_wait(v_full + s, (uint32_t)((vi / Cfg::NV) & 1));
                    dcl_mbar_arrive_remote(v_full + s, 0);
                }
            }
        }
    } else {
        fda_softmax_warps<Cfg, true>(smem, tmem_base, warp, lane, qt, bs, n, NB, s_full, s_empty, o_done, p_full, out);
    }
    __syncwarp();
    __syncthreads();
    dcl_cluster_sync();  // the leader's MMAs read the peer's shared memory until the last commit
    if (warp == 1) {
        tc_fence_after();
        tc2_dealloc(tmem_base, Cfg::TMEM_COLS);
    }
    fda_life(life0, 4);
}

// ------------------------------------------------------------------ attention map (train / inspection)
// A[b, mi, ni] = exp( sum_c RI_2[b,c,mi] * RI_1[b,c,ni] - lse[b,ni] ), fp32 SIMT, 64x64 tiles.
__global__ void __launch_bounds__(256) fda_attention_map_kernel(int c, int n, int m, const float* __restrict__ RI_1,
                                                                const float* __restrict__ RI_2,
                                                                const float* __restrict__ lse,
                                                                float* __restrict__ A) {
    __shared__ float sA[16][64 + 1];  // keys   (c-slab x m-tile)
    __shared__ float sB[16][64 + 1];  // queries (c-slab x n-tile)
    const int bs = blockIdx.z;
    const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    float acc[4][4] = {};
    const float* k_base = RI_2 + (size_t)bs * c * m;
    const float* q_base = RI_1 + (size_t)bs * c * n;
    for (int c0 = 0; c0 < c; c0 += 16) {
        for (int i = threadIdx.x; i < 16 * 64; i += 256) {
            const int cc = i >> 6, x = i & 63;
            sA[cc][x] = (m0 + x < m && c0 + cc < c) ? k_base[(size_t)(c0 + cc) * m + m0 + x] : 0.f;
            sB[cc][x] = (n0 + x < n && c0 + cc < c) ? q_base[(size_t)(c0 + cc) * n + n0 + x] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int cc = 0; cc < 16; ++cc) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                a[i] = sA[cc][ty * 4 + i];
                b[i] = sB[cc][tx * 4 + i];
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int jx = 0; jx < 4; ++jx) acc[i][jx] = __fmaf_rn(a[i], b[jx], acc[i][jx]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int mi = m0 + ty * 4 + i;
        if (mi >= m) continue;
#pragma unroll
        for (int jx = 0; jx < 4; ++jx) {
            const int ni = n0 + tx * 4 + jx;
            if (ni < n) A[((size_t)bs * m + mi) * n + ni] = expf(acc[i][jx] - lse[(size_t)bs * n + ni]);
        }
    }
}


// ------------------------------------------------------------------ UMMA probe (tests / bring-up)
// D (128 x N, fp32) = A (128 x K) * B (N x K)^T with the same packing, descriptors, split
// product and TMEM read-back as the main kernel, in one CTA and without any pipeline.
// `swap_lbo_sbo` exchanges the two stride fields of the descriptors (bring-up knob that
// pins down the descriptor semantics on hardware; the product uses 0).
__global__ void __launch_bounds__(128, 1) umma_probe_kernel(int N, int K, const float* __restrict__ A,
                                                            const float* __restrict__ B, float* __restrict__ D,
                                                            int swap_lbo_sbo) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int a_half = 128 * K * 2, b_half = N * K * 2;
    unsigned char* sA = smem;
    unsigned char* sB = smem + 2 * a_half;
    const int kch = K / 8;
    for (int i = threadIdx.x; i < 128 * kch; i += 128) {
        const int r = i % 128, kc = i / 128;
        __nv_bfloat16 h[8], l[8];
        for (int e = 0; e < 8; ++e) split_bf16(A[(size_t)r * K + kc * 8 + e], h[e], l[e]);
        unsigned char* d = sA + ((r >> 3) * kch + kc) * 128 + (r & 7) * 16;
        *reinterpret_cast<uint4*>(d) = make_uint4(pack2(h[0], h[1]), pack2(h[2], h[3]), pack2(h[4], h[5]), pack2(h[6], h[7]));
        *reinterpret_cast<uint4*>(d + a_half) = make_uint4(pack2(l[0], l[1]), pack2(l[2], l[3]), pack2(l[4], l[5]), pack2(l[6], l[7]));
    }
    const bool b_mn = (swap_lbo_sbo & 2) != 0;
    swap_lbo_sbo &= 1;
    const int nch = N / 8;
    if (b_mn) {
        // B^T image: rows = k (the contraction index), 8-element chunks along n — the "point-major" image of a
        // (K points x N channels) activation, consumed as an MN-major B operand.
        for (int i = threadIdx.x; i < K * nch; i += 128) {
            const int k = i % K, nc = i / K;
            __nv_bfloat16 h[8], l[8];
            for (int e = 0; e < 8; ++e) split_bf16(B[(size_t)(nc * 8 + e) * K + k], h[e], l[e]);
            unsigned char* d = sB + ((k >> 3) * nch + nc) * 128 + (k & 7) * 16;
            *reinterpret_cast<uint4*>(d) = make_uint4(pack2(h[0], h[1]), pack2(h[2], h[3]), pack2(h[4], h[5]), pack2(h[6], h[7]));
            *reinterpret_cast<uint4*>(d + b_half) = make_uint4(pack2(l[0], l[1]), pack2(l[2], l[3]), pack2(l[4], l[5]), pack2(l[6], l[7]));
        }
    }
    for (int i = threadIdx.x; i < (b_mn ? 0 : N * kch); i += 128) {
        const int r = i % N, kc = i / N;
        __nv_bfloat16 h[8], l[8];
        for (int e = 0; e < 8; ++e) split_bf16(B[(size_t)r * K + kc * 8 + e], h[e], l[e]);
        unsigned char* d = sB + ((r >> 3) * kch + kc) * 128 + (r & 7) * 16;
        *reinterpret_cast<uint4*>(d) = make_uint4(pack2(h[0], h[1]), pack2(h[2], h[3]), pack2(h[4], h[5]), pack2(h[6], h[7]));
        *reinterpret_cast<uint4*>(d + b_half) = make_uint4(pack2(l[0], l[1]), pack2(l[2], l[3]), pack2(l[4], l[5]), pack2(l[6], l[7]));
    }
    if (threadIdx.x == 0) {
        dcl_mbar_init(&bar, 1);
        dcl_fence_barrier_init();
    }
    if ((threadIdx.x >> 5) == 0) tc_alloc(&tmem_slot, 256);
    dcl_fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;
    if (threadIdx.x == 0) {
        uint32_t lbo = 128, sbo = (uint32_t)kch * 128;
        if (swap_lbo_sbo) { const uint32_t t = lbo; lbo = sbo; sbo = t; }
        const uint32_t idesc = umma_idesc_bf16(128, N, b_mn);
        const uint32_t a0 = dcl_smem_u32(sA), b0 = dcl_smem_u32(sB);
        // MN-major B: LBO = stride between 8-row K groups, SBO = stride between 8-element chunks along N.
        uint32_t b_lbo = b_mn ? (uint32_t)nch * 128 : lbo, b_sbo = b_mn ? 128u : sbo;
        if (b_mn && swap_lbo_sbo) { const uint32_t t = b_lbo; b_lbo = b_sbo; b_sbo = t; }
        if (b_mn) { lbo = 128; sbo = (uint32_t)kch * 128; }
        for (int kk = 0; kk < K / 16; ++kk) {
            const uint32_t off = kk * 256;
            const uint32_t boff = b_mn ? kk * 2 * (uint32_t)nch * 128 : off;
            mma_split3(tmem_base, a0 + off, a0 + a_half + off, b0 + boff, b0 + b_half + boff, lbo, sbo, b_lbo, b_sbo,
                       idesc, kk == 0);
        }
        tc_commit(&bar);
    }
    dcl_mbar_wait(&bar, 0);
    tc_fence_after();
    const int quad = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row = quad * 32 + lane;
    for (int cc = 0; cc < N / 32; ++cc) {
        uint32_t ov[32];
        DCL_TMEM_LD32(tmem_base + ((uint32_t)(quad * 32) << 16) + cc * 32, ov);
        tc_wait_ld();
        for (int i = 0; i < 32; ++i) D[(size_t)row * N + cc * 32 + i] = __uint_as_float(ov[i]);
    }
    tc_fence_before();
    __syncthreads();
    if ((threadIdx.x >> 5) == 0) {
        tc_fence_after();
        tc_dealloc(tmem_base, 256);
    }
}

// ------------------------------------------------------------------ CTA-pair UMMA probe (tests / bring-up)
// D (256 x N) = A (256 x K) * B (N x K)^T issued once by the leader CTA of a 2-CTA cluster with cta_group::2:
// CTA r holds A rows [128r, 128r+128) and B rows [r N/2, (r+1) N/2) in its own shared memory, and finds rows
// [128r, 128r+128) of D (all N columns) in its own TMEM.  mode 0: the pair meets at a cluster barrier before the
// MMA; mode 1: the peer instead announces its operands with a remote mbarrier arrive on the leader's barrier (the
// hand-off a pipelined kernel uses).
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
    umma_pair_probe_kernel(int N, int K, const float* __restrict__ A, const float* __restrict__ B,
                           float* __restrict__ D, int mode) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bar_done, bar_ready;
    __shared__ uint32_t tmem_slot;
    const uint32_t rank = dcl_cluster_ctarank();
    const int nh = N / 2;
    const int a_half = 128 * K * 2, b_half = nh * K * 2;
    unsigned char* sA = smem;
    unsigned char* sB = smem + 2 * a_half;
    const int kch = K / 8;
    const float* Ar = A + (size_t)rank * 128 * K;
    const float* Br = B + (size_t)rank * nh * K;
    for (int i = threadIdx.x; i < 128 * kch; i += 128) {
        const int r = i % 128, kc = i / 128;
        uint32_t h[4], l[4];
        for (int e = 0; e < 4; ++e)
            split2_bf16(Ar[(size_t)r * K + kc * 8 + 2 * e], Ar[(size_t)r * K + kc * 8 + 2 * e + 1], h[e], l[e]);
        unsigned char* d = sA + ((r >> 3) * kch + kc) * 128 + (r & 7) * 16;
        *reinterpret_cast<uint4*>(d) = make_uint4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<uint4*>(d + a_half) = make_uint4(l[0], l[1], l[2], l[3]);
    }
    for (int i = threadIdx.x; i < nh * kch; i += 128) {
        const int r = i % nh, kc = i / nh;
        uint32_t h[4], l[4];
        for (int e = 0; e < 4; ++e)
            split2_bf16(Br[(size_t)r * K + kc * 8 + 2 * e], Br[(size_t)r * K + kc * 8 + 2 * e + 1], h[e], l[e]);
        unsigned char* d = sB + ((r >> 3) * kch + kc) * 128 + (r & 7) * 16;
        *reinterpret_cast<uint4*>(d) = make_uint4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<uint4*>(d + b_half) = make_uint4(l[0], l[1], l[2], l[3]);
    }
    if (threadIdx.x == 0) {
        dcl_mbar_init(&bar_done, 1);
        dcl_mbar_init(&bar_ready, 1);
        dcl_fence_barrier_init();
    }
    dcl_fence_proxy_async();
    __syncthreads();
    dcl_cluster_sync();  // barriers of both CTAs exist before anything is signalled at them
    if ((threadIdx.x >> 5) == 0) tc2_alloc(&tmem_slot, 256);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;
    if (mode == 0) {
        dcl_cluster_sync();
    } else if (rank == 1 && threadIdx.x == 0) {
        dcl_mbar_arrive_remote(&bar_ready, 0);
    }
    if (rank == 0 && threadIdx.x == 0) {
        if (mode != 0) dcl_mbar_wait_cluster(&bar_ready, 0);
        tc_fence_after();
        const uint32_t idesc = umma_idesc_bf16(256, N);
        const uint64_t dAh = umma_desc(dcl_smem_u32(sA), 128, (uint32_t)kch * 128);
        const uint64_t dAl = dAh + (uint64_t)(a_half >> 4);
        const uint64_t dBh = umma_desc(dcl_smem_u32(sB), 128, (uint32_t)kch * 128);
        const uint64_t dBl = dBh + (uint64_t)(b_half >> 4);
        for (int kk = 0; kk < K / 16; ++kk) {
            const uint64_t off = (uint64_t)(kk * 256) >> 4;
            tc2_mma_bf16(tmem_base, dAh + off, dBh + off, idesc, kk == 0 ? 0u : 1u);
            tc2_mma_bf16(tmem_base, dAh + off, dBl + off, idesc, 1u);
            tc2_mma_bf16(tmem_base, dAl + off, dBh + off, idesc, 1u);
        }
        tc2_commit_mcast(&bar_done, (uint16_t)0x3);
    }
    dcl_mbar_wait(&bar_done, 0);
    tc_fence_after();
    const int quad = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row = quad * 32 + lane;
    for (int cc = 0; cc < N / 32; ++cc) {
        uint32_t ov[32];
        DCL_TMEM_LD32(tmem_base + ((uint32_t)(quad * 32) << 16) + cc * 32, ov);
        tc_wait_ld();
        for (int i = 0; i < 32; ++i) D[((size_t)rank * 128 + row) * N + cc * 32 + i] = __uint_as_float(ov[i]);
    }
    tc_fence_before();
    __syncthreads();
    dcl_cluster_sync();
    if ((threadIdx.x >> 5) == 0) {
        tc_fence_after();
        tc2_dealloc(tmem_base, 256);
    }
}

// ------------------------------------------------------------------ A-from-TMEM UMMA probe (bring-up)
// D (128 x N) = bf16(A) (128 x K) * bf16(B)^T with the A operand read from tensor memory (the "TS" form of
// tcgen05.mma) — the building block for keeping a hidden activation tile on chip between two GEMM layers.
// variant 0: A stored two bf16 per 32-bit TMEM column (K = 16 -> 8 columns per MMA); variant 1: one bf16 per column.
// Not used by the product yet; tools/probe_umma_ts.py reports which convention the hardware follows.
__global__ void __launch_bounds__(128, 1) umma_ts_probe_kernel(int N, int K, const float* __restrict__ A,
                                                               const float* __restrict__ B, float* __restrict__ D,
                                                               int variant) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    unsigned char* sB = smem;
    const int kch = K / 8;
    for (int i = threadIdx.x; i < N * kch; i += 128) {
        const int r = i % N, kc = i / N;
        uint32_t h[4], l[4];
        for (int e = 0; e < 4; ++e)
            split2_bf16(B[(size_t)r * K + kc * 8 + 2 * e], B[(size_t)r * K + kc * 8 + 2 * e + 1], h[e], l[e]);
        *reinterpret_cast<uint4*>(sB + ((r >> 3) * kch + kc) * 128 + (r & 7) * 16) = make_uint4(h[0], h[1], h[2], h[3]);
    }
    if (threadIdx.x == 0) {
        dcl_mbar_init(&bar, 1);
        dcl_fence_barrier_init();
    }
    if ((threadIdx.x >> 5) == 0) tc_alloc(&tmem_slot, 512);
    dcl_fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;
    const int quad = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row = quad * 32 + lane;
    const uint32_t t_lane = (uint32_t)(quad * 32) << 16;
    const uint32_t a_col = 256;
    // this thread's row of A into its TMEM lane
    const int words = variant == 0 ? K / 2 : K;
    for (int w0 = 0; w0 < words; w0 += 32) {
        uint32_t v[32];
        for (int j = 0; j < 32; ++j) {
            const int wj = w0 + j;
            uint32_t h = 0, l = 0;
            if (wj < words) {
                if (variant == 0) split2_bf16(A[(size_t)row * K + 2 * wj], A[(size_t)row * K + 2 * wj + 1], h, l);
                else { split2_bf16(A[(size_t)row * K + wj], 0.f, h, l); h &= 0xffffu; }
            }
            v[j] = h;
        }
        DCL_TMEM_ST32(tmem_base + t_lane + a_col + w0, v);
    }
    tc_wait_st();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (threadIdx.x == 0) {
        const uint32_t idesc = umma_idesc_bf16(128, N);
        const uint64_t dB = umma_desc(dcl_smem_u32(sB), 128, (uint32_t)kch * 128);
        const uint32_t cols_per_step = variant == 0 ? 8u : 16u;
        for (int kk = 0; kk < K / 16; ++kk) {
            const uint32_t ta = tmem_base + a_col + kk * cols_per_step;
            const uint64_t db = dB + (uint64_t)((kk * 256) >> 4);
            const uint32_t acc = kk == 0 ? 0u : 1u;
            asm volatile(
                "{\n\t.reg .pred p;\n\t"
                "setp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                ::"r"(tmem_base), "r"(ta), "l"(db), "r"(idesc), "r"(acc)
                : "memory");
        }
        tc_commit(&bar);
    }
    dcl_mbar_wait(&bar, 0);
    tc_fence_after();
    for (int cc = 0; cc < N / 32; ++cc) {
        uint32_t ov[32];
        DCL_TMEM_LD32(tmem_base + t_lane + cc * 32, ov);
        tc_wait_ld();
        for (int i = 0; i < 32; ++i) D[(size_t)row * N + cc * 32 + i] = __uint_as_float(ov[i]);
    }
    tc_fence_before();
    __syncthreads();
    if ((threadIdx.x >> 5) == 0) {
        tc_fence_after();
        tc_dealloc(tmem_base, 512);
    }
}

template <int C>
struct FdaWs {
    __nv_bfloat16 *Qp, *Kp, *Vp;
    FdaWs(void* workspace, int b, int n, int m) {
        using Cfg = FdaCfg<C>;
        unsigned char* ws = reinterpret_cast<unsigned char*>(workspace);
        const size_t q_bytes = (size_t)b * (n / QT) * Cfg::Q_BYTES;
        const size_t k_bytes = (size_t)b * (m / KB) * Cfg::K_BYTES;
        Qp = reinterpret_cast<__nv_bfloat16*>(ws);
        Kp = reinterpret_cast<__nv_bfloat16*>(ws + q_bytes);
        Vp = reinterpret_cast<__nv_bfloat16*>(ws + q_bytes + k_bytes);
    }
};

template <int C>
int fda_pack_launch(int b, int n, int m, const float* RI_1, const float* RI_2, const float* RE_2, void* workspace,
                    int pv_fmt, cudaStream_t st) {
    using Cfg = FdaCfg<C>;
    FdaWs<C> w(workspace, b, n, m);
    const int work_q = n * (C / 8), work_k = m * (C / 8), work_v = m * (Cfg::VROWS / 8);
    int work = work_q > work_k ? work_q : work_k;
    if (work_v > work) work = work_v;
    dim3 grid(DCL_DIVUP(work, 256), b, 3);
    fda_pack_kernel<C><<<grid, 256, 0, st>>>(n, m, RI_1, RI_2, RE_2, w.Qp, w.Kp, w.Vp, pv_fmt);
    return dcl_launch_status();
}

template <int C>
FdaJobs fda_jobs(int njobs, const dcl_fda_job* in, int b, int n, int m) {
    FdaJobs jobs = {};
    for (int i = 0; i < njobs; ++i) {
        FdaWs<C> w(in[i].workspace, b, n, m);
        jobs.j[i].Qp = w.Qp;
        jobs.j[i].Kp = w.Kp;
        jobs.j[i].Vp = w.Vp;
        jobs.j[i].out = {in[i].RE_embed, in[i].RI_embed, reinterpret_cast<unsigned char*>(in[i].RE_pm),
                         reinterpret_cast<unsigned char*>(in[i].RI_pm), in[i].lse};
    }
    return jobs;
}

template <int C>
int fda_main_launch(int njobs, const dcl_fda_job* in, int b, int n, int m, cudaStream_t st) {
    using Cfg = FdaCfg<C>;
    const FdaJobs jobs = fda_jobs<C>(njobs, in, b, n, m);
    cudaError_t e = cudaFuncSetAttribute(fda_fwd_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
    dim3 grid(n / QT, b, njobs);
    fda_fwd_kernel<C><<<grid, FDA_THREADS, Cfg::SMEM_BYTES, st>>>(n, m, jobs);
    return dcl_launch_status();
}

// CTA-pair variant: needs an even number of query tiles per instance.
template <int C, bool PV16>
int fda_pair_launch(int njobs, const dcl_fda_job* in, int b, int n, int m, cudaStream_t st) {
    using Cfg = FdaPairCfg<C, PV16>;
    const FdaJobs jobs = fda_jobs<C>(njobs, in, b, n, m);
    cudaError_t e = cudaFuncSetAttribute(fda_pair_kernel<C, PV16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(n / QT, b, njobs);
    cfg.blockDim = dim3(FDA_THREADS);
    cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    e = cudaLaunchKernelEx(&cfg, fda_pair_kernel<C, PV16>, n, m, jobs);
    if (e != cudaSuccess) return (int)e;
    return dcl_launch_status();
}

bool fda_use_pair(int n) {
    static const bool forced_single = getenv("DCL_FDA_SINGLE") != nullptr;  // A/B switch for benchmarking
    return !forced_single && (n / QT) % 2 == 0;
}

bool fda_shape_ok(int b, int c, int p, int n, int m) {
    return b >= 0 && (c == 64 || c == 128) && p == FDA_P && n > 0 && m > 0 && n % QT == 0 && m % KB == 0;
}

}  // namespace

DCL_API size_t dcl_fda_workspace_bytes(int b, int c, int p, int n, int m) {
    if (b < 0 || n < 0 || m < 0 || p != FDA_P || (c != 64 && c != 128)) return 0;
    // Q: b*n*c hi+lo bf16; K: b*m*c; V: b*m*(p+c)
    return (size_t)b * ((size_t)n * c + (size_t)m * c + (size_t)m * (p + c)) * 4 + 1024;
}

DCL_API int dcl_fda_pack_fmt(int b, int c, int p, int n, int m, const float* RI_1, const float* RI_2,
                             const float* RE_2, void* workspace, size_t workspace_bytes, int pv_fmt, void* stream) {
    DCL_RETURN_IF_BAD(fda_shape_ok(b, c, p, n, m) && (pv_fmt == 0 || pv_fmt == 1));
    DCL_RETURN_IF_BAD(workspace != nullptr && (((uintptr_t)workspace) & 127u) == 0);
    DCL_RETURN_IF_BAD(workspace_bytes >= dcl_fda_workspace_bytes(b, c, p, n, m));
    DCL_RETURN_IF_BAD(((((uintptr_t)RE_2) | ((uintptr_t)RI_2)) & 15u) == 0);
    if (b == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    if (c == 64) return fda_pack_launch<64>(b, n, m, RI_1, RI_2, RE_2, workspace, pv_fmt, st);
    return fda_pack_launch<128>(b, n, m, RI_1, RI_2, RE_2, workspace, pv_fmt, st);
}

DCL_API int dcl_fda_pack(int b, int c, int p, int n, int m, const float* RI_1, const float* RI_2, const float* RE_2,
                         void* workspace, size_t workspace_bytes, void* stream) {
    return dcl_fda_pack_fmt(b, c, p, n, m, RI_1, RI_2, RE_2, workspace, workspace_bytes, 0, stream);
}

DCL_API int dcl_fda_fwd_packed_jobs_fmt(int njobs, const dcl_fda_job* jobs, int b, int c, int p, int n, int m,
                                        size_t workspace_bytes, int pv_fmt, void* stream) {
    DCL_RETURN_IF_BAD(njobs >= 1 && njobs <= FDA_MAX_JOBS && jobs != nullptr && fda_shape_ok(b, c, p, n, m));
    // the fp16 P V form exists in the CTA-pair kernel only: an even number of 128-query tiles per instance
    DCL_RETURN_IF_BAD(pv_fmt == 0 || (pv_fmt == 1 && (n / QT) % 2 == 0));
    DCL_RETURN_IF_BAD(workspace_bytes >= dcl_fda_workspace_bytes(b, c, p, n, m));
    for (int i = 0; i < njobs; ++i) {
        DCL_RETURN_IF_BAD(jobs[i].workspace != nullptr && (((uintptr_t)jobs[i].workspace) & 127u) == 0);
        DCL_RETURN_IF_BAD((((uintptr_t)jobs[i].RE_pm) & 15u) == 0 && (((uintptr_t)jobs[i].RI_pm) & 15u) == 0);
    }
    if (b == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    if (pv_fmt == 1) {
        if (c == 64) return fda_pair_launch<64, true>(njobs, jobs, b, n, m, st);
        return fda_pair_launch<128, true>(njobs, jobs, b, n, m, st);
    }
    if (fda_use_pair(n)) {
        if (c == 64) return fda_pair_launch<64, false>(njobs, jobs, b, n, m, st);
        return fda_pair_launch<128, false>(njobs, jobs, b, n, m, st);
    }
    if (c == 64) return fda_main_launch<64>(njobs, jobs, b, n, m, st);
    return fda_main_launch<128>(njobs, jobs, b, n, m, st);
}

DCL_API int dcl_fda_fwd_packed_jobs(int njobs, const dcl_fda_job* jobs, int b, int c, int p, int n, int m,
                                    size_t workspace_bytes, void* stream) {
    return dcl_fda_fwd_packed_jobs_fmt(njobs, jobs, b, c, p, n, m, workspace_bytes, 0, stream);
}

DCL_API int dcl_fda_fwd_packed_pm(int b, int c, int p, int n, int m, float* RE_embed, float* RI_embed, void* RE_pm,
                                  void* RI_pm, float* lse_out, void* workspace, size_t workspace_bytes,
                                  void* stream) {
    const dcl_fda_job job = {workspace, RE_embed, RI_embed, RE_pm, RI_pm, lse_out};
    return dcl_fda_fwd_packed_jobs(1, &job, b, c, p, n, m, workspace_bytes, stream);
}

DCL_API int dcl_fda_fwd_packed(int b, int c, int p, int n, int m, float* RE_embed, float* RI_embed, float* lse_out,
                               void* workspace, size_t workspace_bytes, void* stream) {
    DCL_RETURN_IF_BAD(RE_embed != nullptr && RI_embed != nullptr);
    return dcl_fda_fwd_packed_pm(b, c, p, n, m, RE_embed, RI_embed, nullptr, nullptr, lse_out, workspace,
                                 workspace_bytes, stream);
}

DCL_API int dcl_fda_align_fwd(int b, int c, int p, int n, int m, const float* RI_1, const float* RI_2,
                              const float* RE_2, float* RE_embed, float* RI_embed, float* lse_out, void* workspace,
                              size_t workspace_bytes, void* stream) {
    int e = dcl_fda_pack(b, c, p, n, m, RI_1, RI_2, RE_2, workspace, workspace_bytes, stream);
    if (e) return e;
    return dcl_fda_fwd_packed(b, c, p, n, m, RE_embed, RI_embed, lse_out, workspace, workspace_bytes, stream);
}

DCL_API int dcl_fda_attention_map(int b, int c, int n, int m, const float* RI_1, const float* RI_2,
                                  const float* lse, float* A, void* stream) {
    DCL_RETURN_IF_BAD(b >= 0 && c > 0 && n > 0 && m > 0);
    if (b == 0) return 0;
    dim3 grid(DCL_DIVUP(n, 64), DCL_DIVUP(m, 64), b);
    fda_attention_map_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(c, n, m, RI_1, RI_2, lse, A);
    return dcl_launch_status();
}

DCL_API int dcl_debug_umma_gemm(int N, int K, const float* A, const float* B, float* D, int swap_lbo_sbo,
                                void* stream) {
    DCL_RETURN_IF_BAD(N >= 32 && N <= 256 && N % 32 == 0 && K >= 16 && K % 16 == 0 && swap_lbo_sbo >= 0 && swap_lbo_sbo < 4);
    const size_t smem = (size_t)(128 + N) * K * 4;
    DCL_RETURN_IF_BAD(smem <= 200 * 1024);
    cudaError_t e = cudaFuncSetAttribute(umma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    umma_probe_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(N, K, A, B, D, swap_lbo_sbo);
    return dcl_launch_status();
}

DCL_API int dcl_debug_umma_pair_gemm(int N, int K, const float* A, const float* B, float* D, int mode, void* stream) {
    DCL_RETURN_IF_BAD(N >= 32 && N <= 256 && N % 32 == 0 && K >= 16 && K % 16 == 0 && mode >= 0 && mode < 2);
    const size_t smem = (size_t)(128 + N / 2) * K * 4;
    DCL_RETURN_IF_BAD(smem <= 200 * 1024);
    cudaError_t e =
        cudaFuncSetAttribute(umma_pair_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    umma_pair_probe_kernel<<<2, 128, smem, (cudaStream_t)stream>>>(N, K, A, B, D, mode);
    return dcl_launch_status();
}

DCL_API int dcl_debug_fda_set_trace(long long* device_buffer) {
    cudaError_t e = cudaMemcpyToSymbol(g_fda_trace, &device_buffer, sizeof(device_buffer));
    return (int)e;
}

DCL_API int dcl_fda_workspace_layout(int b, int c, int p, int n, int m, size_t* offsets3) {
    DCL_RETURN_IF_BAD(offsets3 != nullptr && fda_shape_ok(b, c, p, n, m));
    const size_t q_bytes = (size_t)b * n * c * 4, k_bytes = (size_t)b * m * c * 4;
    offsets3[0] = 0;
    offsets3[1] = q_bytes;
    offsets3[2] = q_bytes + k_bytes;
    return 0;
}

DCL_API int dcl_debug_umma_ts_gemm(int N, int K, const float* A, const float* B, float* D, int variant, void* stream) {
    DCL_RETURN_IF_BAD(N >= 32 && N <= 256 && N % 32 == 0 && K >= 16 && K % 16 == 0 && K <= 256 && variant >= 0 &&
                      variant < 2);
    const size_t smem = (size_t)N * K * 2;
    DCL_RETURN_IF_BAD(smem <= 200 * 1024);
    cudaError_t e = cudaFuncSetAttribute(umma_ts_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    umma_ts_probe_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(N, K, A, B, D, variant);
    return dcl_launch_status();
}
