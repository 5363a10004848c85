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
// Warps: 0-7 gather and afterwards run the epilogue (TMEM lane = row, two warps per lane quadrant), 8 issues the
// MMAs, 9 streams the weights and the CTA's rulebook tiles.
#include "common.cuh"
#include "umma.cuh"
#include "../../include/dcl_b200.h"
#include <cuda_fp16.h>

namespace {

constexpr int SC_BM = 128;
constexpr int SC_KC = 64;            // virtual channels per stage
constexpr int SC_GW = 8;             // gather / epilogue warps
constexpr int SC_THREADS = (SC_GW + 2) * 32;   // + MMA issuer + weight stream
constexpr int SC_A_BYTES = SC_BM * SC_KC * 2;   // 16384

// MT = row tiles per CTA (they share every W stage): 4 where a layer has thousands of tiles, 1 for the deep levels,
// whose few hundred tiles must spread over all SMs.  Ring depths: NT = 128 keeps one CTA per SM with deep rings;
// narrower layers use shallower rings so that two CTAs share an SM (one's epilogue under the other's main loop).
constexpr int SC_TAB_BYTES = 27 * SC_BM * 4;    // rulebook of one tile: 27 offsets x 128 rows (contiguous in nbr[tile])

template <int NT, int MT>
struct ScCfg {
    static constexpr int NA = (NT == 256 || (NT == 128 && MT > 3)) ? 4 : 6;   // A-stage ring
    static constexpr int NW = (NT == 128 && MT <= 2) ? 3 : 2;                 // W-stage ring
    static constexpr int W_HALF = NT * SC_KC * 2;
    static constexpr int W_BYTES = 2 * W_HALF;
    static constexpr int OFF_W = NA * SC_A_BYTES;
    static constexpr int OFF_TAB = OFF_W + NW * W_BYTES;               // the CTA's MT rulebook tiles, loaded once
    static constexpr int OFF_BAR = OFF_TAB + MT * SC_TAB_BYTES;
    static constexpr int SMEM_BYTES = OFF_BAR + 256;
    static constexpr int sc_pow2(int v) { return v <= 32 ? 32 : v <= 64 ? 64 : v <= 128 ? 128 : v <= 256 ? 256 : 512; }
    static constexpr int TMEM_COLS = sc_pow2(MT * NT);
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

// Timeline trace (bring-up / profiling): when a buffer is installed (dcl_debug_spconv_set_trace), CTA (0,0,0) stamps
// clock64() per role and iteration: trace[role * 512 + iteration], roles 0 = gather after waiting for its A buffer,
// 1 = gather after issuing the copies, 2 = MMA after the A stage landed, 3 = MMA after issuing, 4 = W issued,
// 5 = MMA after the W stage landed.
__device__ long long* g_sc_trace = nullptr;
__device__ __forceinline__ void sc_stamp(long long* tr, int role, int it) {
    if (tr != nullptr && it < 512) tr[role * 512 + it] = clock64();
}

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
    uint64_t* a_full = bars;                 // [SC_NA] one arrival per gather thread
    uint64_t* a_empty = a_full + SC_NA;      // [SC_NA] one commit
    uint64_t* w_full = a_empty + SC_NA;      // [SC_NW] TMA bytes
    uint64_t* w_empty = w_full + SC_NW;      // [SC_NW] one commit
    uint64_t* acc_full = w_empty + SC_NW;    // [1]
    uint64_t* tab_full = acc_full + 1;       // [1] rulebook tiles landed
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tab_full + 1);

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
            dcl_mbar_init(a_full + i, SC_GW * 32);
            dcl_mbar_init(a_empty + i, 1);
        }
        for (int i = 0; i < SC_NW; ++i) {
            dcl_mbar_init(w_full + i, 1);
            dcl_mbar_init(w_empty + i, CS);
        }
        dcl_mbar_init(acc_full, 1);
        dcl_mbar_init(tab_full, 1);
        dcl_fence_barrier_init();
    }
    if (warp == SC_GW && nt_mine > 0) tc_alloc(tmem_slot, Cfg::TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    if constexpr (CS > 1) dcl_cluster_sync();          // the peers' barriers exist before anything is multicast at them
    tc_fence_after();
    const uint32_t tmem_base = nt_mine > 0 ? *tmem_slot : 0u;

    if (warp < SC_GW) {
      if (nt_mine > 0) {
        // ===================== gather (8 warps) =====================
        // A warp instruction covers 8 rows x 4 chunks of 16 bytes: lane%8 = row within an 8-row group (the eight
        // 16-byte slots of a core-matrix row block: conflict-free shared-memory writes), lane/8 = chunk, so that each
        // input row is read 64 contiguous bytes at a time.  Warp w owns row groups 2w and 2w+1 of every tile.  The
        // rulebook entries come from the CTA's tiles of the table in shared memory (one bulk copy per tile at kernel
        // start).  History (tools/trace_spconv.py): with 4 gather warps, runtime divisions and the table in global
        // memory the gather warps' own instruction stream took ~950 cycles per 16 KB stage — not L2, not the tensor
        // pipe — which is what every layer ran at.
        const int rl = lane & 7, jc = lane >> 3;
        const uint32_t sA = dcl_smem_u32(smem);
        const int* tab = reinterpret_cast<const int*>(smem + Cfg::OFF_TAB);
        long long* tr = (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && threadIdx.x == 0) ? g_sc_trace : nullptr;
        const int cin_shift = 31 - __clz(cin), cin_mask = cin - 1;          // cin is a power of two
        dcl_mbar_wait(tab_full, 0);
        int issued = 0;
        for (int s = 0; s < nstages; ++s) {
            if (!((smask >> s) & 1ull)) continue;
            int kk[2], cc0[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int kv = s * SC_KC + (jc + 4 * h) * 8;
                kk[h] = kv >> cin_shift;
                cc0[h] = kv & cin_mask;
            }
            for (int t = 0; t < nt_mine; ++t) {
                const int buf = issued % SC_NA;
                if (issued >= SC_NA) dcl_mbar_wait(a_empty + buf, (uint32_t)(((issued / SC_NA) - 1) & 1));
                sc_stamp(tr, 0, issued);
                const uint32_t dst0 = sA + buf * SC_A_BYTES + (uint32_t)rl * 16u;
#pragma unroll
                for (int gi = 0; gi < 2; ++gi) {
                    const int grp = warp * 2 + gi;
                    const int rt = grp * 8 + rl;
                    const bool live = (tile0 + t) * SC_BM + rt < total;
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int src_row = (live && kk[h] < 27) ? tab[(t * 27 + kk[h]) * SC_BM + rt] : -1;
                        const __half* src = tw.in16 + (((size_t)(src_row < 0 ? 0 : src_row)) << cin_shift) + cc0[h];
                        sc_cp_async16(dst0 + (uint32_t)grp * (SC_KC / 8) * 128u + (uint32_t)(jc + 4 * h) * 128u, src,
                                      src_row < 0 ? 0u : 16u);
                    }
                }
                // the stage's barrier receives this thread's arrival when its copies have landed (no wait, no stall:
                // the thread runs ahead until the ring is full) — the hand-off CUTLASS's sm100 cp.async mainloop uses
                asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(dcl_smem_u32(a_full + buf)) : "memory");
                sc_stamp(tr, 1, issued);
                ++issued;
            }
        }
        // ===================== epilogue: TMEM lane = output row; warps w and w+4 split the columns =====================
        dcl_mbar_wait(acc_full, 0);
        tc_fence_after();
        const int quad = warp & 3, half = warp >> 2;
        const uint32_t t_lane = (uint32_t)(quad * 32) << 16;
        const int row = quad * 32 + lane;
        constexpr int CH = NT / 8 / 2;          // 8-column chunks per thread
        for (int t = 0; t < nt_mine; ++t) {
            const int r = (tile0 + t) * SC_BM + row;
#pragma unroll 1
            for (int cc = half * CH; cc < (half + 1) * CH; ++cc) {
                uint32_t v[8];
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                    : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                    : "r"(tmem_base + t_lane + t * NT + cc * 8)
                    : "memory");
                tc_wait_ld();
                if (r < total) {
                    const int col0 = nti * NT + cc * 8;
                    const float4 s0 = __ldg(reinterpret_cast<const float4*>(tw.shift + col0));
                    const float4 s1 = __ldg(reinterpret_cast<const float4*>(tw.shift + col0) + 1);
                    float y[8];
                    y[0] = fmaxf(__uint_as_float(v[0]) + s0.x, 0.f);
                    y[1] = fmaxf(__uint_as_float(v[1]) + s0.y, 0.f);
                    y[2] = fmaxf(__uint_as_float(v[2]) + s0.z, 0.f);
                    y[3] = fmaxf(__uint_as_float(v[3]) + s0.w, 0.f);
                    y[4] = fmaxf(__uint_as_float(v[4]) + s1.x, 0.f);
                    y[5] = fmaxf(__uint_as_float(v[5]) + s1.y, 0.f);
                    y[6] = fmaxf(__uint_as_float(v[6]) + s1.z, 0.f);
                    y[7] = fmaxf(__uint_as_float(v[7]) + s1.w, 0.f);
                    if (tw.out32 != nullptr) {
                        float4* o = reinterpret_cast<float4*>(tw.out32 + (size_t)r * cout + col0);
                        o[0] = make_float4(y[0], y[1], y[2], y[3]);
                        o[1] = make_float4(y[4], y[5], y[6], y[7]);
                    }
                    if (tw.out16 != nullptr) {
                        uint32_t h[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const __half2 hh = __floats2half2_rn(fminf(y[2 * e], 65504.f), fminf(y[2 * e + 1], 65504.f));
                            h[e] = *reinterpret_cast<const uint32_t*>(&hh);
                        }
                        *reinterpret_cast<uint4*>(tw.out16 + (size_t)r * cout + col0) = make_uint4(h[0], h[1], h[2], h[3]);
                    }
                }
            }
        }
        tc_fence_before();
      }
    } else if (warp == SC_GW) {
        // ===================== MMA issuer =====================
        if (dcl_elect_one()) {
            constexpr uint32_t idesc = (1u << 4) | ((uint32_t)(NT >> 3) << 17) | ((uint32_t)(SC_BM >> 4) << 24);  // f16 x f16 -> f32
            constexpr uint32_t SBO = (SC_KC / 8) * 128;
            const uint64_t dA0 = umma_desc(dcl_smem_u32(smem), 128, SBO);
            const uint64_t dW0 = umma_desc(dcl_smem_u32(smem + Cfg::OFF_W), 128, SBO);
            int ai = 0, wi = 0;
            bool first = true;
            long long* tr = (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) ? g_sc_trace : nullptr;
            for (int s = 0; s < nstages; ++s) {
                if (!((smask >> s) & 1ull)) continue;
                const int wb = wi % SC_NW;
                dcl_mbar_wait(w_full + wb, (uint32_t)((wi / SC_NW) & 1));
                sc_stamp(tr, 5, wi);
                const uint64_t dWh = dW0 + (uint64_t)((wb * Cfg::W_BYTES) >> 4);
                const uint64_t dWl = dWh + (uint64_t)(Cfg::W_HALF >> 4);
                for (int t = 0; t < nt_mine; ++t, ++ai) {
                    const int ab = ai % SC_NA;
                    dcl_mbar_wait(a_full + ab, (uint32_t)((ai / SC_NA) & 1));
                    sc_stamp(tr, 2, ai);
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
                    sc_stamp(tr, 3, ai);
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
            if (nt_mine > 0) {
                dcl_mbar_arrive_expect_tx(tab_full, (uint32_t)nt_mine * SC_TAB_BYTES);
                for (int t = 0; t < nt_mine; ++t)
                    dcl_bulk_g2s(smem + Cfg::OFF_TAB + t * SC_TAB_BYTES, tw.nbr + (size_t)(tile0 + t) * 32 * SC_BM,
                                 SC_TAB_BYTES, tab_full);
            }
            const unsigned char* w = tw.w + (size_t)nti * nstages * Cfg::W_BYTES;
            int wi = 0;
            for (int s = 0; s < nstages; ++s) {
                if (!((smask >> s) & 1ull)) continue;
                const int wb = wi % SC_NW;
                if (wi >= SC_NW) dcl_mbar_wait(w_empty + wb, (uint32_t)(((wi / SC_NW) - 1) & 1));
                if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) sc_stamp(g_sc_trace, 4, wi);
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
    if (warp == SC_GW && nt_mine > 0) {
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
    // Row tiles per CTA (MT): the tiles of a CTA share every W stage, and — what matters more for these short
    // kernels — the CTA count should not spill into a nearly empty second wave: MT = tiles / 148 rounded up, capped by
    // TMEM (MT * NT <= 512 columns) and shared memory.  tiles_total comes from the buffer capacities, which are
    // planned ~1.3x above the real row counts.
    constexpr int MT_MAX = NT == 256 ? 2 : NT == 128 ? 4 : 6;
    const int est_tiles = tiles_total * 10 / 13;
    int mt = DCL_DIVUP(est_tiles, 148);
    if (mt > MT_MAX) mt = MT_MAX;
    if (mt <= 1) return sc_launch<NT, 1>(args, ntowers, max_cap, st);
    if (mt == 2) return sc_launch<NT, 2>(args, ntowers, max_cap, st);
    if constexpr (MT_MAX >= 4) {
        if (mt == 3) return sc_launch<NT, 3>(args, ntowers, max_cap, st);
        if (mt == 4) return sc_launch<NT, 4>(args, ntowers, max_cap, st);
    }
    if constexpr (MT_MAX >= 6) return sc_launch<NT, 6>(args, ntowers, max_cap, st);
    return sc_launch<NT, MT_MAX>(args, ntowers, max_cap, st);
}

}  // namespace

DCL_API int dcl_debug_spconv_set_trace(long long* device_buffer) {
    return (int)cudaMemcpyToSymbol(g_sc_trace, &device_buffer, sizeof(device_buffer));
}

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
