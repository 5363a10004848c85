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
constexpr int SC_THREADS = 192;
constexpr int SC_A_BYTES = SC_BM * SC_KC * 2;   // 16384

// MT = row tiles per CTA (they share every W stage): 4 where a layer has thousands of tiles, 1 for the deep levels,
// whose few hundred tiles must spread over all SMs.  Ring depths: NT = 128 keeps one CTA per SM with deep rings;
// narrower layers use shallower rings so that two CTAs share an SM (one's epilogue under the other's main loop).
template <int NT, int MT>
struct ScCfg {
    static constexpr int NA = NT == 128 ? 6 : 4;          // A-stage ring
    static constexpr int NW = NT == 128 ? 3 : 2;          // W-stage ring
    static constexpr int W_HALF = NT * SC_KC * 2;
    static constexpr int W_BYTES = 2 * W_HALF;
    static constexpr int OFF_W = NA * SC_A_BYTES;
    static constexpr int OFF_BAR = OFF_W + NW * W_BYTES;
    static constexpr int SMEM_BYTES = OFF_BAR + 256;
    static constexpr int TMEM_COLS = (MT * NT) < 32 ? 32 : (MT * NT);   // powers of two: NT and MT are
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
    // .ca: the 16-byte requests of the lanes that share a 32-byte sector merge in L1's tag stage (with .cg every
    // lane pulled its own sector from L2: twice the bytes, ncu sectors/request 25 instead of 16)
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}

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

// CS = CTAs per cluster.  CS > 1 (the wide, deep layers, which are bound by streaming W from L2): the CTAs of a
// cluster own consecutive tile groups and walk the same stage list; each fetches 1/CS of every W stage and
// multicasts it into all of them, and a W buffer is refilled only after every CTA's MMAs have released it
// (tcgen05.commit multicast, CS arrivals per phase).  A CTA of an active cluster that has no tiles of its own still
// loads and releases its share.
template <int NT, int MT, int CS>
__global__ void __launch_bounds__(SC_THREADS, 1) sparse_conv3_kernel(const __grid_constant__ ScArgs args) {
    using Cfg = ScCfg<NT, MT>;
    constexpr int SC_NA = Cfg::NA, SC_NW = Cfg::NW, SC_MT = MT;
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
    const int ctile0 = (blockIdx.x / CS) * CS * SC_MT;     // first tile of the cluster
    if (ctile0 >= ntiles) return;                      // uniform for the cluster: nothing was allocated yet
    const int nt_mine = max(0, min(SC_MT, ntiles - tile0));
    const int nti = blockIdx.y;
    const int cin = args.cin, cout = args.cout, nstages = args.nstages;
    const uint32_t crank = CS > 1 ? dcl_cluster_ctarank() : 0u;

    unsigned int kmask = 0;                            // kernel offsets used by any row of the cluster's tiles
    for (int t = ctile0; t < min(ntiles, ctile0 + CS * SC_MT); ++t) kmask |= tw.anymask[t];
    const unsigned long long smask = sc_stage_mask(kmask, cin, nstages);

    if (threadIdx.x == 0) {
        for (int i = 0; i < SC_NA; ++i) {
            dcl_mbar_init(a_full + i, 128);
            dcl_mbar_init(a_empty + i, 1);
        }
        for (int i = 0; i < SC_NW; ++i) {
            dcl_mbar_init(w_full + i, 1);
            dcl_mbar_init(w_empty + i, CS);
        }
        dcl_mbar_init(acc_full, 1);
        dcl_fence_barrier_init();
    }
    if (warp == 4 && nt_mine > 0) tc_alloc(tmem_slot, Cfg::TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    if constexpr (CS > 1) dcl_cluster_sync();          // the peers' barriers exist before anything is multicast at them
    tc_fence_after();
    const uint32_t tmem_base = nt_mine > 0 ? *tmem_slot : 0u;

    if (warp < 4) {
      if (nt_mine > 0) {
        // ===================== gather =====================
        // A warp instruction covers 8 rows x 4 chunks of 16 bytes: lane%8 = row within an 8-row group (the eight
        // 16-byte slots of a core-matrix row block: conflict-free shared-memory writes), lane/8 = chunk, so that each
        // input row is read 64 contiguous bytes at a time (two instructions per row) instead of 16.
        const int rl = lane & 7, jc = lane >> 3;
        const uint32_t sA = dcl_smem_u32(smem);
        // rulebook entries of one (stage, tile): this thread's 4 rows x 2 chunks.  They are fetched ONE iteration ahead
        // of the copies that need them, so that the gather never waits on a dependent global load.
        auto load_rows = [&](int s, int t, int (&rows)[8]) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int k = (s * SC_KC + (jc + 4 * h) * 8) / cin;
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    const int rt = warp * 32 + g * 8 + rl;
                    const int r = (tile0 + t) * SC_BM + rt;
                    rows[g * 2 + h] = (k < 27 && r < total) ? __ldg(tw.nbr + ((size_t)(tile0 + t) * 32 + k) * 128 + rt) : -1;
                }
            }
        };
        auto next_stage = [&](int s) {
            ++s;
            while (s < nstages && !((smask >> s) & 1ull)) ++s;
            return s;
        };
        int issued = 0;
        int s = next_stage(-1), t = 0;
        int cur[8], nxt[8];
        if (s < nstages) load_rows(s, 0, cur);
        while (s < nstages) {
            int s2 = s, t2 = t + 1;
            if (t2 == nt_mine) {
                t2 = 0;
                s2 = next_stage(s);
            }
            if (s2 < nstages) load_rows(s2, t2, nxt);
            const int buf = issued % SC_NA;
            if (issued >= SC_NA) dcl_mbar_wait(a_empty + buf, (uint32_t)(((issued / SC_NA) - 1) & 1));
            const uint32_t dst0 = sA + buf * SC_A_BYTES + (uint32_t)rl * 16u;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int kv = s * SC_KC + (jc + 4 * h) * 8;
                const int c0 = kv - (kv / cin) * cin;
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    const int src_row = cur[g * 2 + h];
                    const __half* src = tw.in16 + (size_t)(src_row < 0 ? 0 : src_row) * cin + c0;
                    sc_cp_async16(dst0 + (uint32_t)(warp * 4 + g) * (SC_KC / 8) * 128u + (uint32_t)(jc + 4 * h) * 128u, src,
                                  src_row < 0 ? 0u : 16u);
                }
            }
            // the stage's barrier receives this thread's arrival when its copies have landed (no wait, no stall: the
            // thread runs ahead until the ring is full) — the hand-off CUTLASS's sm100 cp.async mainloop uses
            asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(dcl_smem_u32(a_full + buf)) : "memory");
            ++issued;
#pragma unroll
            for (int i = 0; i < 8; ++i) cur[i] = nxt[i];
            s = s2;
            t = t2;
        }
        // ===================== epilogue: TMEM lane = output row =====================
        dcl_mbar_wait(acc_full, 0);
        tc_fence_after();
        const uint32_t t_lane = (uint32_t)(warp * 32) << 16;
        const int row = threadIdx.x;
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
      }
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
                if constexpr (CS > 1) tc_commit_mcast(w_empty + wb, (uint16_t)((1u << CS) - 1u));
                else tc_commit(w_empty + wb);
                ++wi;
            }
            if (nt_mine > 0) tc_commit(acc_full);
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
                if constexpr (CS > 1) {
                    constexpr uint32_t SLICE = Cfg::W_BYTES / CS;
                    dcl_bulk_g2s_mcast(smem + Cfg::OFF_W + wb * Cfg::W_BYTES + crank * SLICE,
                                       w + (size_t)s * Cfg::W_BYTES + crank * SLICE, SLICE, w_full + wb,
                                       (uint16_t)((1u << CS) - 1u));
                } else {
                    dcl_bulk_g2s(smem + Cfg::OFF_W + wb * Cfg::W_BYTES, w + (size_t)s * Cfg::W_BYTES, Cfg::W_BYTES,
                                 w_full + wb);
                }
                ++wi;
            }
        }
    }
    __syncwarp();
    __syncthreads();
    if constexpr (CS > 1) dcl_cluster_sync();          // no CTA leaves while a peer may still multicast into it
    if (warp == 4 && nt_mine > 0) {
        tc_fence_after();
        tc_dealloc(tmem_base, Cfg::TMEM_COLS);
    }
}

template <int NT, int MT>
int sc_launch(const ScArgs& args, int ntowers, int max_cap, cudaStream_t st) {
    using Cfg = ScCfg<NT, MT>;
    constexpr int CS = NT >= 128 ? 4 : 1;          // W multicast where W streaming is what binds
    cudaError_t e = cudaFuncSetAttribute(sparse_conv3_kernel<NT, MT, CS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
    int groups = DCL_DIVUP(max_cap / SC_BM, MT);
    groups = DCL_DIVUP(groups, CS) * CS;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(groups, args.cout / NT, ntowers);
    cfg.blockDim = dim3(SC_THREADS);
    cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CS;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    e = cudaLaunchKernelEx(&cfg, sparse_conv3_kernel<NT, MT, CS>, args);
    if (e != cudaSuccess) return (int)e;
    return dcl_launch_status();
}

template <int NT>
int sc_launch_nt(const ScArgs& args, int ntowers, int max_cap, int tiles_total, cudaStream_t st) {
    // row tiles per CTA: every W stage is fetched once per CTA, so more tiles per CTA = less L2 traffic (the deep,
    // wide layers are bound by exactly that), as long as there are still CTAs for every SM
    if constexpr (NT * 4 <= 512) {
        if (tiles_total >= 148 * 4) return sc_launch<NT, 4>(args, ntowers, max_cap, st);
    }
    if (tiles_total >= 148) return sc_launch<NT, 2>(args, ntowers, max_cap, st);
    return sc_launch<NT, 1>(args, ntowers, max_cap, st);
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
    const int nt = cout;                      // one n-tile: the gathered rows are read once for all output channels
    int tiles_total = 0;
    for (int t = 0; t < ntowers; ++t) tiles_total += convs[t].cap_out / SC_BM;
    if (nt == 16) return sc_launch_nt<16>(args, ntowers, max_cap, tiles_total, st);
    if (nt == 32) return sc_launch_nt<32>(args, ntowers, max_cap, tiles_total, st);
    if (nt == 64) return sc_launch_nt<64>(args, ntowers, max_cap, tiles_total, st);
    if (nt == 128) return sc_launch_nt<128>(args, ntowers, max_cap, tiles_total, st);
    return sc_launch_nt<256>(args, ntowers, max_cap, tiles_total, st);
}
