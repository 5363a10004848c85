// Fused backward of one FDA direction (models/Modules.py:166-169 under autograd: bmm -> softmax -> two bmm), the
// (M x N) attention matrix and its gradient never leaving the chip.  With S[m,n] = sum_c K[c,m] Q[c,n],
// A = softmax_m S, E = V A, I = K A (values Vc = [V; K], their gradients gc = [gE; gI]):
//      dA = Vc^T gc          D[n] = sum_m A dA = <gE, E>[n] + <gI, I>[n]     (from the forward's outputs)
//      dS = A o (dA - D)     dVc = gc A^T        dK = Q dS^T        dQ = K dS
// A is rebuilt from the forward's per-query log-sum-exp: exp(S - lse), S recomputed with the same split-bf16 product.
//
// One kernel template, three roles — each is "a stationary 128-row tile against streamed 64-row blocks", every
// product a bf16 hi/lo split tcgen05.mma (3 MMAs, fp32 accumulate in TMEM):
//      ROLE_DQ  tile = 128 queries, blocks = keys:    T1 = S,  T2 = dA,   W = dS,    ACC(128 x C)  += W K^T-image
//      ROLE_DK  tile = 128 keys,    blocks = queries: T1 = S^T, T2 = dA^T, W = dS^T,  ACC(128 x C)  += W Q^T-image
//      ROLE_DV  tile = 128 keys,    blocks = queries: T1 = S^T,           W = A^T,   ACC(128 x VC) += W gc^T-image
// Warp 0 streams operands (1-D TMA bulk copies of PM-image blobs: csrc/train_ops.cu writes the images), warp 1
// issues the MMAs, warps 2-5 (thread = tile row = TMEM lane) turn T1 / T2 into the hi/lo image of W in shared memory.
// T1/T2 are double-buffered in TMEM so the products of block j+1 run under the pointwise pass of block j.
// Operands: "K-images" (PM image of the (points x channels) matrix; contraction over channels) and per-instance
// "T-images" (PM image of the (channels x points) matrix; contraction over points), see include/dcl_b200.h.
#include "common.cuh"
#include "umma.cuh"
#include "../../include/dcl_b200.h"

namespace {

constexpr int FB_ST = 128;     // stationary rows per CTA
constexpr int FB_BLK = 64;     // streamed rows per block
constexpr int FB_THREADS = 192;
constexpr int FB_BLOB = 16384;
constexpr float FB_LOG2E = 1.4426950408889634f;
constexpr int ROLE_DQ = 0, ROLE_DK = 1, ROLE_DV = 2;

struct FbJob {
    const unsigned char* stat1;   // K-image of the stationary side, C wide   (DQ: Q;  DK/DV: K)
    const unsigned char* str1;    // K-image of the streamed side, C wide     (DQ: K;  DK/DV: Q)
    const unsigned char* stat2;   // K-image, VC wide                         (DQ: gc; DK: Vc)
    const unsigned char* str2;    //                                          (DQ: Vc; DK: gc)
    const unsigned char* zt;      // T-images of the accumulated operand      (DQ: K;  DK: Q;  DV: gc)
    const float* lse;             // (b, n_queries)
    const float* dsum;            // (b, n_queries)  D
    float* out;                   // (b, ACC_W, n_stat)
};
struct FbJobs { FbJob j[2]; };

template <int C, int ROLE>
struct FbCfg {
    static constexpr int VC = 256 + C;
    static constexpr bool HAS_T2 = ROLE != ROLE_DV;
    static constexpr int KB1 = C / 32;
    static constexpr int KB2 = HAS_T2 ? VC / 32 : 0;
    static constexpr int STAGE = FB_BLOB + FB_BLOB / 2;         // A blob (128 rows) + 64 streamed rows, hi and lo
    static constexpr int NS = ROLE == ROLE_DV ? 3 : 4;
    static constexpr int NW = ROLE == ROLE_DV ? 1 : 2;
    static constexpr int NZ = ROLE == ROLE_DV ? 1 : 2;
    static constexpr int ACC_W = ROLE == ROLE_DV ? VC : C;
    static constexpr int ZT = (ACC_W + 127) / 128;              // 128-row tiles of the accumulated operand
    static constexpr int Z_ROWS_PAD = ZT * 128;                 // rows of the T-image per instance
    static constexpr int Z_STAGE = ZT * 2 * FB_BLOB;            // per block: ZT tiles x 2 k-blocks of 32 points
    static constexpr int W_BYTES = FB_ST * FB_BLK * 4;          // hi image + lo image
    static constexpr int TW = HAS_T2 ? 128 : 64;                // TMEM columns per T buffer
    static constexpr int ACC_COL = 2 * TW;
    static constexpr int OFF_W = NS * STAGE;
    static constexpr int OFF_Z = OFF_W + NW * W_BYTES;
    static constexpr int OFF_VEC = OFF_Z + NZ * Z_STAGE;        // [2 buffers][lse | D][64]
    static constexpr int OFF_BAR = OFF_VEC + 2 * 2 * FB_BLK * 4;
    static constexpr int SMEM_BYTES = OFF_BAR + 256;
    static_assert(ACC_COL + ACC_W <= 512, "TMEM budget");
    static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
    __host__ __device__ static constexpr int tile_rows(int t) { return ACC_W - 128 * t >= 128 ? 128 : ACC_W - 128 * t; }
};

template <int C, int ROLE>
__global__ void __launch_bounds__(FB_THREADS, 1) fda_bwd_kernel(int n_stat, int n_str, const __grid_constant__ FbJobs jobs) {
    using Cfg = FbCfg<C, ROLE>;
    extern __shared__ __align__(1024) unsigned char smem[];
    const FbJob& job = jobs.j[blockIdx.z];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);
    uint64_t* full = bars;                       // [NS]
    uint64_t* empty = full + Cfg::NS;            // [NS]
    uint64_t* z_full = empty + Cfg::NS;          // [2]
    uint64_t* z_empty = z_full + 2;              // [2]
    uint64_t* t_full = z_empty + 2;              // [2]
    uint64_t* t_empty = t_full + 2;              // [2]
    uint64_t* w_full = t_empty + 2;              // [2]
    uint64_t* w_empty = w_full + 2;              // [2]
    uint64_t* acc_full = w_empty + 2;            // [1]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile = blockIdx.x, bs = blockIdx.y;
    const int NB = n_str / FB_BLK;

    if (threadIdx.x == 0) {
        for (int i = 0; i < Cfg::NS; ++i) {
            dcl_mbar_init(full + i, 1);
            dcl_mbar_init(empty + i, 1);
        }
        for (int i = 0; i < 2; ++i) {
            dcl_mbar_init(z_full + i, 1);
            dcl_mbar_init(z_empty + i, 1);
            dcl_mbar_init(t_full + i, 1);
            dcl_mbar_init(t_empty + i, 128);
            dcl_mbar_init(w_full + i, 128);
            dcl_mbar_init(w_empty + i, 1);
        }
        dcl_mbar_init(acc_full, 1);
        dcl_fence_barrier_init();
    }
    if (warp == 1) tc_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (dcl_elect_one()) {
            const size_t rt_stat = (size_t)bs * (n_stat / 128) + tile;     // 128-row tile of the stationary K-images
            int it = 0;
            for (int j = 0; j < NB; ++j) {
                const size_t rt_str = (size_t)bs * (n_str / 128) + (j >> 1);
                const int half = j & 1;
                for (int kb = 0; kb < Cfg::KB1 + Cfg::KB2; ++kb, ++it) {
                    const int s = it % Cfg::NS;
                    if (it >= Cfg::NS) dcl_mbar_wait(empty + s, (uint32_t)(((it / Cfg::NS) - 1) & 1));
                    unsigned char* dst = smem + s * Cfg::STAGE;
                    const bool first = kb < Cfg::KB1;
                    const int k = first ? kb : kb - Cfg::KB1;
                    const int kbs = first ? Cfg::KB1 : Cfg::KB2;
                    const unsigned char* a = (first ? job.stat1 : job.stat2) + (rt_stat * kbs + k) * FB_BLOB;
                    const unsigned char* b = (first ? job.str1 : job.str2) + (rt_str * kbs + k) * FB_BLOB + half * 4096;
                    dcl_mbar_arrive_expect_tx(full + s, Cfg::STAGE);
                    dcl_bulk_g2s(dst, a, FB_BLOB, full + s);
                    dcl_bulk_g2s(dst + FB_BLOB, b, 4096, full + s);
                    dcl_bulk_g2s(dst + FB_BLOB + 4096, b + FB_BLOB / 2, 4096, full + s);
                }
                // the accumulated operand of block j: ZT row tiles x 2 k-blocks of its T-image
                const int zs = j % Cfg::NZ;
                if (j >= Cfg::NZ) dcl_mbar_wait(z_empty + zs, (uint32_t)(((j / Cfg::NZ) - 1) & 1));
                uint32_t bytes = 0;
#pragma unroll
                for (int t = 0; t < Cfg::ZT; ++t) bytes += 2 * 2 * Cfg::tile_rows(t) * 64;
                dcl_mbar_arrive_expect_tx(z_full + zs, bytes);
                const unsigned char* zimg = job.zt + (size_t)bs * Cfg::Z_ROWS_PAD * n_str * 4;
#pragma unroll
                for (int t = 0; t < Cfg::ZT; ++t) {
#pragma unroll
                    for (int kb = 0; kb < 2; ++kb) {
                        const unsigned char* src = zimg + ((size_t)t * (n_str / 32) + 2 * j + kb) * FB_BLOB;
                        unsigned char* dst = smem + Cfg::OFF_Z + zs * Cfg::Z_STAGE + (t * 2 + kb) * FB_BLOB;
                        const uint32_t hb = Cfg::tile_rows(t) * 64;
                        dcl_bulk_g2s(dst, src, hb, z_full + zs);
                        dcl_bulk_g2s(dst + FB_BLOB / 2, src + FB_BLOB / 2, hb, z_full + zs);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (dcl_elect_one()) {
            constexpr uint32_t idesc_t = umma_idesc_bf16(FB_ST, FB_BLK);
            const uint32_t s0 = dcl_smem_u32(smem);
            int it = 0;
            auto issue_t = [&](int j) {
                const int tb = j & 1;
                if (j >= 2) dcl_mbar_wait(t_empty + tb, (uint32_t)(((j >> 1) - 1) & 1));
                tc_fence_after();
                for (int kb = 0; kb < Cfg::KB1 + Cfg::KB2; ++kb, ++it) {
                    const int s = it % Cfg::NS;
                    dcl_mbar_wait(full + s, (uint32_t)((it / Cfg::NS) & 1));
                    tc_fence_after();
                    const bool first = kb < Cfg::KB1;
                    const uint32_t tcol = tmem_base + tb * Cfg::TW + (first ? 0 : 64);
                    const bool start = first ? kb == 0 : kb == Cfg::KB1;
                    const uint32_t a = s0 + s * Cfg::STAGE, b = a + FB_BLOB;
#pragma unroll
                    for (int ks = 0; ks < 2; ++ks)
                        mma_split3(tcol, a + ks * 256, a + FB_BLOB / 2 + ks * 256, b + ks * 256, b + 4096 + ks * 256, 128, 512,
                                   128, 512, idesc_t, start && ks == 0);
                    tc_commit(empty + s);
                }
                tc_commit(t_full + tb);
            };
            issue_t(0);
            for (int j = 0; j < NB; ++j) {
                if (j + 1 < NB) issue_t(j + 1);
                const int wb = j % Cfg::NW, zs = j % Cfg::NZ;
                dcl_mbar_wait(w_full + wb, (uint32_t)((j / Cfg::NW) & 1));
                dcl_mbar_wait(z_full + zs, (uint32_t)((j / Cfg::NZ) & 1));
                tc_fence_after();
                const uint32_t w = s0 + Cfg::OFF_W + wb * Cfg::W_BYTES;
                const uint32_t z = s0 + Cfg::OFF_Z + zs * Cfg::Z_STAGE;
#pragma unroll
                for (int ks = 0; ks < FB_BLK / 16; ++ks) {
#pragma unroll
                    for (int t = 0; t < Cfg::ZT; ++t) {
                        const uint32_t zb = z + (t * 2 + (ks >> 1)) * FB_BLOB + (ks & 1) * 256;
                        mma_split3(tmem_base + Cfg::ACC_COL + t * 128, w + ks * 256, w + Cfg::W_BYTES / 2 + ks * 256, zb,
                                   zb + FB_BLOB / 2, 128, 1024, 128, 512, umma_idesc_bf16(FB_ST, Cfg::tile_rows(t)),
                                   j == 0 && ks == 0);
                    }
                }
                tc_commit(w_empty + wb);
                tc_commit(z_empty + zs);
            }
            tc_commit(acc_full);
        }
    } else {
        // ===================== pointwise pass: T1 (, T2) -> W =====================
        const int ew = warp - 2;
        const int row = ((ew + 2) & 3) * 32 + lane;                  // tile row == TMEM lane (warp w reads quadrant w % 4)
        const uint32_t t_lane = (uint32_t)(((ew + 2) & 3) * 32) << 16;
        const int tid = ew * 32 + lane;
        float* vec = reinterpret_cast<float*>(smem + Cfg::OFF_VEC);
        const size_t qbase = (size_t)bs * (ROLE == ROLE_DQ ? n_stat : n_str);
        float my_lse = 0.f, my_d = 0.f;
        if (ROLE == ROLE_DQ) {
            my_lse = __ldg(job.lse + qbase + tile * FB_ST + row) * FB_LOG2E;
            my_d = __ldg(job.dsum + qbase + tile * FB_ST + row);
        }
        for (int j = 0; j < NB; ++j) {
            const int tb = j & 1, wb = j % Cfg::NW;
            float* vb = vec + tb * 2 * FB_BLK;
            if (ROLE != ROLE_DQ) {
                // per-query vectors of the streamed block (columns of T1 / T2)
                if (tid < FB_BLK) vb[tid] = __ldg(job.lse + qbase + j * FB_BLK + tid) * FB_LOG2E;
                else vb[tid] = __ldg(job.dsum + qbase + j * FB_BLK + tid - FB_BLK);
                asm volatile("bar.sync 1, 128;" ::: "memory");
            }
            dcl_mbar_wait(t_full + tb, (uint32_t)((j >> 1) & 1));
            tc_fence_after();
            if (j >= Cfg::NW) dcl_mbar_wait(w_empty + wb, (uint32_t)(((j / Cfg::NW) - 1) & 1));
            unsigned char* wrow = smem + Cfg::OFF_W + wb * Cfg::W_BYTES + (row >> 3) * 1024 + (row & 7) * 16;
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                uint32_t sv[32], dv[32];
                DCL_TMEM_LD32(tmem_base + t_lane + tb * Cfg::TW + hh * 32, sv);
                if (Cfg::HAS_T2) DCL_TMEM_LD32(tmem_base + t_lane + tb * Cfg::TW + 64 + hh * 32, dv);
                tc_wait_ld();
#pragma unroll
                for (int kc = 0; kc < 4; ++kc) {
                    uint32_t h[4], l[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        float w2[2];
#pragma unroll
                        for (int q = 0; q < 2; ++q) {
                            const int col = kc * 8 + 2 * e + q;
                            const float ls = ROLE == ROLE_DQ ? my_lse : vb[hh * 32 + col];
                            const float p = ex2_approx(__fmaf_rn(__uint_as_float(sv[col]), FB_LOG2E, -ls));
                            if (Cfg::HAS_T2) {
                                const float dd = ROLE == ROLE_DQ ? my_d : vb[FB_BLK + hh * 32 + col];
                                w2[q] = p * (__uint_as_float(dv[col]) - dd);
                            } else {
                                w2[q] = p;
                            }
                        }
                        split2_bf16(w2[0], w2[1], h[e], l[e]);
                    }
                    unsigned char* d = wrow + (hh * 4 + kc) * 128;
                    *reinterpret_cast<uint4*>(d) = make_uint4(h[0], h[1], h[2], h[3]);
                    *reinterpret_cast<uint4*>(d + Cfg::W_BYTES / 2) = make_uint4(l[0], l[1], l[2], l[3]);
                }
            }
            tc_fence_before();
            dcl_mbar_arrive(t_empty + tb);
            dcl_fence_proxy_async();
            dcl_mbar_arrive(w_full + wb);
        }
        // ---- epilogue: ACC (128 x ACC_W) -> out (b, ACC_W, n_stat), lane = row => coalesced along the points
        dcl_mbar_wait(acc_full, 0);
        tc_fence_after();
        float* o = job.out + (size_t)bs * Cfg::ACC_W * n_stat + (size_t)tile * FB_ST + row;
#pragma unroll 1
        for (int cc = 0; cc < Cfg::ACC_W / 32; ++cc) {
            uint32_t v[32];
            DCL_TMEM_LD32(tmem_base + t_lane + Cfg::ACC_COL + cc * 32, v);
            tc_wait_ld();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[(size_t)(cc * 32 + i) * n_stat] = __uint_as_float(v[i]);
        }
        tc_fence_before();
    }
    __syncwarp();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tc_dealloc(tmem_base, 512);
    }
}

template <int C, int ROLE>
int fb_launch(int b, int n_stat, int n_str, const FbJobs& jobs, int njobs, cudaStream_t st) {
    using Cfg = FbCfg<C, ROLE>;
    cudaError_t e = cudaFuncSetAttribute(fda_bwd_kernel<C, ROLE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
    fda_bwd_kernel<C, ROLE><<<dim3(n_stat / FB_ST, b, njobs), FB_THREADS, Cfg::SMEM_BYTES, st>>>(n_stat, n_str, jobs);
    return dcl_launch_status();
}

template <int C>
int fb_run(int njobs, const dcl_fda_bwd_job* jobs, int b, int n, int m, cudaStream_t st) {
    FbJobs dq, dk, dv;
    for (int i = 0; i < njobs; ++i) {
        const dcl_fda_bwd_job& j = jobs[i];
        auto u8 = [](const void* p) { return reinterpret_cast<const unsigned char*>(p); };
        dq.j[i] = FbJob{u8(j.q_k), u8(j.k_k), u8(j.g_k), u8(j.v_k), u8(j.k_t), j.lse, j.dsum, j.d_q};
        dk.j[i] = FbJob{u8(j.k_k), u8(j.q_k), u8(j.v_k), u8(j.g_k), u8(j.q_t), j.lse, j.dsum, j.d_k};
        dv.j[i] = FbJob{u8(j.k_k), u8(j.q_k), nullptr, nullptr, u8(j.g_t), j.lse, j.dsum, j.d_v};
    }
    int e = fb_launch<C, ROLE_DQ>(b, n, m, dq, njobs, st);
    if (e != 0) return e;
    e = fb_launch<C, ROLE_DK>(b, m, n, dk, njobs, st);
    if (e != 0) return e;
    return fb_launch<C, ROLE_DV>(b, m, n, dv, njobs, st);
}

}  // namespace

DCL_API int dcl_fda_bwd(int njobs, const dcl_fda_bwd_job* jobs, int b, int c, int p, int n, int m, void* stream) {
    DCL_RETURN_IF_BAD(njobs >= 1 && njobs <= 2 && jobs != nullptr && b > 0 && b <= 65535 && p == 256);
    DCL_RETURN_IF_BAD((c == 64 || c == 128) && n > 0 && n % 128 == 0 && m > 0 && m % 128 == 0);
    for (int i = 0; i < njobs; ++i) {
        const dcl_fda_bwd_job& j = jobs[i];
        DCL_RETURN_IF_BAD(j.q_k && j.k_k && j.g_k && j.v_k && j.q_t && j.k_t && j.g_t && j.lse && j.dsum && j.d_q &&
                          j.d_k && j.d_v);
        DCL_RETURN_IF_BAD(((((uintptr_t)j.q_k) | ((uintptr_t)j.k_k) | ((uintptr_t)j.g_k) | ((uintptr_t)j.v_k) |
                            ((uintptr_t)j.q_t) | ((uintptr_t)j.k_t) | ((uintptr_t)j.g_t)) & 15u) == 0);
    }
    cudaStream_t st = (cudaStream_t)stream;
    return c == 64 ? fb_run<64>(njobs, jobs, b, n, m, st) : fb_run<128>(njobs, jobs, b, n, m, st);
}
