// Flat, batch-id-aware three_nn / three_interpolate — the variants DCL-Net's
// Network.forward actually calls (models/Modules.py:213-251).
// Semantics: libs/pointnet_sp/src/interpolate_gpu.cu:9-56 (three_nn),
//            :80-102 (three_interpolate), :124-146 (grad).
//
// three_nn in the reference is O(n*m): every query scans every known row and skips
// the ones whose batch id differs.  Two implementations here, bit-identical:
//  * full scan (reference signature, no scratch): known rows streamed through shared
//    memory by TMA bulk copies;
//  * segmented (caller-provided scratch): known rows are bucketed by batch id on the
//    device (histogram -> scan -> scatter), each query scans only its bucket.  The
//    scatter is not order-preserving, so candidates are ranked by the explicit
//    lexicographic key (d, original index) — which is exactly the order the
//    reference's ascending scan with strict '<' produces.
// three_interpolate uses one lane per 4 channels of a row so that both the three
// feature-row reads and the output write are coalesced 128-bit accesses (the
// reference strides by C between lanes).
#include "common.cuh"
#include <cuda_fp16.h>
#include "tile_pipe.cuh"
#include "umma.cuh"
#include "../../include/dcl_b200.h"
#include <math_constants.h>

namespace {

constexpr int SP_THREADS = 128;
constexpr int SP_TILE_ROWS = 1024;  // 16 KB per stage
constexpr int SP_TILE_FLOATS = SP_TILE_ROWS * 4;

__device__ __forceinline__ void nn3_insert_strict(float d, int k, float& b1, float& b2, float& b3, int& i1, int& i2,
                                                  int& i3) {
    if (d < b3) {
        if (d < b2) {
            b3 = b2;
            i3 = i2;
            if (d < b1) {
                b2 = b1;
                i2 = i1;
                b1 = d;
                i1 = k;
            } else {
                b2 = d;
                i2 = k;
            }
        } else {
            b3 = d;
            i3 = k;
        }
    }
}

// (d,k) < (b,i) lexicographically
__device__ __forceinline__ bool lex_lt(float d, int k, float b, int i) { return d < b || (d == b && k < i); }

__device__ __forceinline__ void nn3_insert_lex(float d, int k, float& b1, float& b2, float& b3, int& i1, int& i2,
                                               int& i3) {
    if (lex_lt(d, k, b3, i3)) {
        if (lex_lt(d, k, b2, i2)) {
            b3 = b2;
            i3 = i2;
            if (lex_lt(d, k, b1, i1)) {
                b2 = b1;
                i2 = i1;
                b1 = d;
                i1 = k;
            } else {
                b2 = d;
                i2 = k;
            }
        } else {
            b3 = d;
            i3 = k;
        }
    }
}

// ---------------------------------------------------------------- full scan
__global__ void __launch_bounds__(SP_THREADS) sp_three_nn_scan_kernel(int n, int m, const float* __restrict__ unknown,
                                                                      const float* __restrict__ known,
                                                                      float* __restrict__ dist2,
                                                                      int* __restrict__ idx) {
    __shared__ __align__(16) float s_tile[2 * SP_TILE_FLOATS];
    __shared__ uint64_t s_bar[2];
    const int qi = blockIdx.x * SP_THREADS + threadIdx.x;
    const int qc = min(qi, n - 1);
    const float4 u = reinterpret_cast<const float4*>(unknown)[qc];
    float b1 = CUDART_INF_F, b2 = CUDART_INF_F, b3 = CUDART_INF_F;
    int i1 = 0, i2 = 0, i3 = 0;
    DclTilePipe<SP_TILE_FLOATS> pipe;
    pipe.init(s_tile, s_bar, known, m * 4);
    for (int t = 0; t < pipe.ntiles; ++t) {
        const int cnt = pipe.acquire(t) / 4;
        const float4* tile = reinterpret_cast<const float4*>(pipe.tile(t));
        const int kbase = t * SP_TILE_ROWS;
#pragma unroll 4
        for (int j = 0; j < cnt; ++j) {
            const float4 k4 = tile[j];
            if (k4.x != u.x) continue;
            const float d = dcl_dist2(u.y, u.z, u.w, k4.y, k4.z, k4.w);
            nn3_insert_strict(d, kbase + j, b1, b2, b3, i1, i2, i3);
        }
        pipe.release(t);
    }
    if (qi < n) {
        dist2[qi * 3 + 0] = b1;
        dist2[qi * 3 + 1] = b2;
        dist2[qi * 3 + 2] = b3;
        idx[qi * 3 + 0] = i1;
        idx[qi * 3 + 1] = i2;
        idx[qi * 3 + 2] = i3;
    }
}

// ---------------------------------------------------------------- segmented
// Workspace layout (ints unless noted):
//   [0]              flag: 1 if some known batch id is not an integer in [0, MAXB)
//   [1]              max batch id seen (+1), i.e. number of buckets in use
//   [4 .. 4+MAXB]    bucket offsets (MAXB+1 entries; counts before the scan)
//   [.. +MAXB]       scatter cursors
//   then 16-B aligned float4 sorted[m] = (x, y, z, bits(original index))
constexpr int WS_FLAG = 0, WS_NB = 1, WS_OFF = 4;
constexpr int WS_CUR = WS_OFF + DCL_SP_MAX_BATCH + 4;
constexpr int WS_HDR_INTS = WS_CUR + DCL_SP_MAX_BATCH;  // multiple of 4 -> sorted[] is 16-B aligned

__device__ __forceinline__ bool batch_id_ok(float b, int& ib) {
    ib = (int)b;
    return (b == (float)ib) && ib >= 0 && ib < DCL_SP_MAX_BATCH;
}

// How the m known rows are handed over: (m,4) float bxyz rows, or — fusing Ops_tensor2points
// (models/Modules.py:204-211) — (m,4) int voxel indices (b,ix,iy,iz) whose centres are
//   ((float(i) * ext) + offset) + 0.5*ext      evaluated in fp32 in exactly that order, as torch does.
struct KnownRows {
    const float* rows;
    const int* vox;
    float ext[3], off[3], half[3];
    // first voxel index of row k (voxel form only): rows that share it share their first centre coordinate
    __device__ __forceinline__ int slab(int k) const { return __ldg(vox + 4 * (size_t)k + 1); }
    __device__ __forceinline__ float4 get(int k) const {
        if (rows != nullptr) return __ldg(reinterpret_cast<const float4*>(rows) + k);
        const int4 v = __ldg(reinterpret_cast<const int4*>(vox) + k);
        return make_float4((float)v.x, __fadd_rn(__fadd_rn(__fmul_rn((float)v.y, ext[0]), off[0]), half[0]),
                           __fadd_rn(__fadd_rn(__fmul_rn((float)v.z, ext[1]), off[1]), half[1]),
                           __fadd_rn(__fadd_rn(__fmul_rn((float)v.w, ext[2]), off[2]), half[2]));
    }
};

constexpr int SP_SBINS = 1024;  // batch ids below this are aggregated per block in shared memory

__global__ void __launch_bounds__(256) sp_bucket_hist_kernel(int m, KnownRows kr, int* __restrict__ ws) {
    __shared__ int s_cnt[SP_SBINS];
    __shared__ int s_max;
    for (int i = threadIdx.x; i < SP_SBINS; i += 256) s_cnt[i] = 0;
    if (threadIdx.x == 0) s_max = 0;
    __syncthreads();
    const int k = blockIdx.x * 256 + threadIdx.x;
    if (k < m) {
        int ib;
        if (batch_id_ok(kr.get(k).x, ib)) {
            if (ib < SP_SBINS) atomicAdd(s_cnt + ib, 1);
            else atomicAdd(ws + WS_OFF + ib, 1);
            atomicMax(&s_max, ib + 1);
        } else {
            atomicOr(ws + WS_FLAG, 1);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < SP_SBINS; i += 256)
        if (s_cnt[i] != 0) atomicAdd(ws + WS_OFF + i, s_cnt[i]);
    if (threadIdx.x == 0 && s_max != 0) atomicMax(ws + WS_NB, s_max);
}

// Single CTA: exclusive scan of the bucket counts in place; off[nb] = total.
__global__ void __launch_bounds__(1024) sp_bucket_scan_kernel(int* __restrict__ ws) {
    __shared__ int s_warp[32];
    __shared__ int s_carry;
    const int nb = ws[WS_NB];
    int* off = ws + WS_OFF;
    int* cur = ws + WS_CUR;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < nb; base += 1024) {
        const int i = base + threadIdx.x;
        const int v = (i < nb) ? off[i] : 0;
        int x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, x, o);
            if ((threadIdx.x & 31) >= o) x += y;
        }
        if ((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = x;
        __syncthreads();
        if (threadIdx.x < 32) {
            int w = s_warp[threadIdx.x];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int y = __shfl_up_sync(0xffffffffu, w, o);
                if (threadIdx.x >= o) w += y;
            }
            s_warp[threadIdx.x] = w;
        }
        __syncthreads();
        const int warp_prefix = (threadIdx.x >= 32) ? s_warp[(threadIdx.x >> 5) - 1] : 0;
        const int incl = x + warp_prefix + s_carry;
        if (i < nb) {
            off[i] = incl - v;
            cur[i] = 0;
        }
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) off[nb] = s_carry;
}

__global__ void __launch_bounds__(256) sp_bucket_scatter_kernel(int m, KnownRows kr, int* __restrict__ ws,
                                                                float4* __restrict__ sorted) {
    __shared__ int s_cnt[SP_SBINS];  // per-block count, then the block's base inside the bucket
    for (int i = threadIdx.x; i < SP_SBINS; i += 256) s_cnt[i] = 0;
    __syncthreads();
    const int k = blockIdx.x * 256 + threadIdx.x;
    float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
    int ib = -1, local = 0;
    if (k < m) {
        r = kr.get(k);
        if (!batch_id_ok(r.x, ib)) ib = -1;
        else if (ib < SP_SBINS) local = atomicAdd(s_cnt + ib, 1);
        else local = atomicAdd(ws + WS_CUR + ib, 1);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < SP_SBINS; i += 256)
        if (s_cnt[i] != 0) s_cnt[i] = atomicAdd(ws + WS_CUR + i, s_cnt[i]);
    __syncthreads();
    if (ib >= 0) {
        const int pos = ws[WS_OFF + ib] + local + (ib < SP_SBINS ? s_cnt[ib] : 0);
        sorted[pos] = make_float4(r.y, r.z, r.w, __int_as_float(k));
    }
}

// 3-NN of one query by a group of LPQ consecutive lanes: lane `sub` scans candidates sub, sub+LPQ, ... of the
// query's bucket (coalesced 16-B loads), then the LPQ partial triples are merged by a shuffle butterfly.  The
// lexicographic key (d, original index) makes the result independent of who saw which candidate in which order.
// All 32 lanes of the warp must call this (full-mask shuffles); afterwards every lane of a group holds the result.
constexpr int LPQ = 8;
constexpr int SP_SEG_TILE_FLOATS = 512 * 4;  // 512 bucket entries (8 KB) per stage
// Slab mode (voxel levels, gx > 0): the buckets are keyed by (batch id, first voxel index), i.e. an instance's
// entries are grouped into gx slabs that share one centre coordinate.  A query walks the slabs outwards from its
// own and stops as soon as the squared distance to the next slab's plane exceeds what it already holds — every
// candidate of such a slab is at least that far away — which cuts the candidates per query from the whole instance
// (~1100 at the finest level) to a few slabs.  The instance's entries (<= SP_SLAB_CAP) sit in shared memory.
constexpr int SP_SLAB_CAP = 2048;
constexpr int SP_SLAB_MAX_GX = 128;
constexpr int SP_STILE_FLOATS = SP_SLAB_CAP * 4 > 2 * SP_SEG_TILE_FLOATS ? SP_SLAB_CAP * 4 : 2 * SP_SEG_TILE_FLOATS;

template <int LPQ_>
__device__ __forceinline__ void sp_group_search_t(const int* __restrict__ ws, const float4* __restrict__ sorted,
                                                const KnownRows& kr, int m, bool valid, float4 u,
                                                int sub, float& b1, float& b2, float& b3, int& i1, int& i2, int& i3,
                                                int gx = 0) {
    b1 = b2 = b3 = CUDART_INF_F;
    i1 = i2 = i3 = 0x7fffffff;  // sentinels lose every tie; mapped back to the reference's 0 at the end
    if (ws[WS_FLAG] != 0) {
        // Some batch id is not a small non-negative integer: buckets are unusable, scan everything
        // (float equality on the id, as the reference does).
        if (valid) {
            for (int k = sub; k < m; k += LPQ_) {
                const float4 c = kr.get(k);
                if (c.x != u.x) continue;
                const float d = dcl_dist2(u.y, u.z, u.w, c.y, c.z, c.w);
                if (d < CUDART_INF_F) nn3_insert_lex(d, k, b1, b2, b3, i1, i2, i3);  // inf / NaN never enter
            }
        }
    } else {
        int my_b = -1;
        const bool has_bucket = valid && batch_id_ok(u.x, my_b) && my_b < ws[WS_NB];
        // Common case: every query of the CTA belongs to the same batch item (clouds are stored batch-major).
        // Then the bucket is streamed ONCE per CTA through shared memory by TMA bulk copies (2-stage ring) and
        // all 16 query groups scan it there; ncu showed the per-group global loads stalled on the long
        // scoreboard two thirds of the time.  Mixed CTAs take the per-group global path below.
        __shared__ __align__(16) float s_tile[SP_STILE_FLOATS];
        __shared__ uint64_t s_bar[2];
        __shared__ int s_b0;
        __shared__ int s_slab[SP_SLAB_MAX_GX + 1];
        const int gxe = gx > 0 ? gx : 1;  // buckets per batch id
        if (threadIdx.x == 0) s_b0 = has_bucket ? my_b : -1;
        __syncthreads();
        const int b0 = s_b0;
        const bool uniform = __syncthreads_and((!valid || (has_bucket && my_b == b0)) ? 1 : 0) && b0 >= 0;
        const int ubeg = uniform ? ws[WS_OFF + b0 * gxe] : 0, uend = uniform ? ws[WS_OFF + (b0 + 1) * gxe] : 0;
        if (uniform && gx > 0 && gx <= SP_SLAB_MAX_GX && uend - ubeg <= SP_SLAB_CAP) {
            // ---- slab walk over the instance staged in shared memory
            const uint32_t bytes = (uint32_t)(uend - ubeg) * 16u;
            if (threadIdx.x == 0) {
                dcl_mbar_init(&s_bar[0], 1);
                dcl_fence_barrier_init();
                if (bytes != 0) {
                    dcl_mbar_arrive_expect_tx(&s_bar[0], bytes);
                    dcl_bulk_g2s(s_tile, sorted + ubeg, bytes, &s_bar[0]);
                }
            }
            for (int i = threadIdx.x; i <= gx; i += blockDim.x) s_slab[i] = ws[WS_OFF + b0 * gx + i] - ubeg;
            __syncthreads();
            if (bytes != 0) dcl_mbar_wait(&s_bar[0], 0);
            if (valid && bytes != 0) {
                const float4* inst = reinterpret_cast<const float4*>(s_tile);
                const unsigned gmask = ((1u << LPQ_) - 1u) << ((threadIdx.x & 31) & ~(LPQ_ - 1));
                // squared distance to the plane of slab s; its centre coordinate is formed exactly as KnownRows::get does
                auto slab_d2 = [&](int sidx) {
                    const float cx = __fadd_rn(__fadd_rn(__fmul_rn((float)sidx, kr.ext[0]), kr.off[0]), kr.half[0]);
                    const float dx = __fsub_rn(u.y, cx);
                    return __fmul_rn(dx, dx);
                };
                int lo = (int)floorf(__fdiv_rn(__fsub_rn(u.y, kr.off[0]), kr.ext[0]));
                lo = max(0, min(gx - 1, lo));
                int hi = lo + 1;
                float bound = CUDART_INF_F;
                while (lo >= 0 || hi < gx) {
                    const float dl = lo >= 0 ? slab_d2(lo) : CUDART_INF_F;
                    const float dh = hi < gx ? slab_d2(hi) : CUDART_INF_F;
                    const bool take_lo = hi >= gx || (lo >= 0 && dl <= dh);
                    const float dmin = take_lo ? dl : dh;
                    // every unvisited candidate is at least dmin (1 - 3 ulp) away: nothing closer than `bound` is left
                    if (!(dmin <= __fmul_rn(bound, 1.000001f))) break;
                    const int sidx = take_lo ? lo-- : hi++;
                    const int sb = s_slab[sidx], se = s_slab[sidx + 1];
                    for (int j = sb + sub; j < se; j += LPQ_) {
                        const float4 c = inst[j];
                        const float d = dcl_dist2(u.y, u.z, u.w, c.x, c.y, c.z);
                        if (!(d > b3) && d < CUDART_INF_F)
                            nn3_insert_lex(d, __float_as_int(c.w), b1, b2, b3, i1, i2, i3);
                    }
                    if (se > sb) {
                        // the group's third-best distance is at most any of its lanes' third-best
                        float t = b3;
#pragma unroll
                        for (int o = LPQ_ / 2; o > 0; o >>= 1) t = fminf(t, __shfl_xor_sync(gmask, t, o));
                        bound = t;
                    }
                }
            }
            __syncthreads();  // s_tile / s_bar are reused by the next level
        } else if (uniform) {
            const int beg = ubeg, end = uend;
            DclTilePipe<SP_SEG_TILE_FLOATS> pipe;
            pipe.init(s_tile, s_bar, reinterpret_cast<const float*>(sorted + beg), (end - beg) * 4);
            for (int t = 0; t < pipe.ntiles; ++t) {
                const int cnt = pipe.acquire(t) / 4;
                const float4* tile = reinterpret_cast<const float4*>(pipe.tile(t));
                if (valid) {
#pragma unroll 4
                    for (int j = sub; j < cnt; j += LPQ_) {
                        const float4 c = tile[j];
                        const float d = dcl_dist2(u.y, u.z, u.w, c.x, c.y, c.z);
                        if (!(d > b3) && d < CUDART_INF_F)
                            nn3_insert_lex(d, __float_as_int(c.w), b1, b2, b3, i1, i2, i3);
                    }
                }
                pipe.release(t);
            }
        } else if (has_bucket) {
            const int beg = ws[WS_OFF + my_b * gxe], end = ws[WS_OFF + (my_b + 1) * gxe];
#pragma unroll 8
            for (int j = beg + sub; j < end; j += LPQ_) {
                const float4 c = __ldg(sorted + j);
                const float d = dcl_dist2(u.y, u.z, u.w, c.x, c.y, c.z);
                // one compare rejects almost every candidate; ties (d == b3) and the first three go the slow way
                if (!(d > b3) && d < CUDART_INF_F) nn3_insert_lex(d, __float_as_int(c.w), b1, b2, b3, i1, i2, i3);
            }
        }
    }
#pragma unroll
    for (int o = LPQ_ / 2; o > 0; o >>= 1) {
        const float ob1 = __shfl_xor_sync(0xffffffffu, b1, o), ob2 = __shfl_xor_sync(0xffffffffu, b2, o),
                    ob3 = __shfl_xor_sync(0xffffffffu, b3, o);
        const int oi1 = __shfl_xor_sync(0xffffffffu, i1, o), oi2 = __shfl_xor_sync(0xffffffffu, i2, o),
                  oi3 = __shfl_xor_sync(0xffffffffu, i3, o);
        nn3_insert_lex(ob1, oi1, b1, b2, b3, i1, i2, i3);
        nn3_insert_lex(ob2, oi2, b1, b2, b3, i1, i2, i3);
        nn3_insert_lex(ob3, oi3, b1, b2, b3, i1, i2, i3);
    }
    if (i1 == 0x7fffffff) i1 = 0;  // unfilled slots: (inf, 0) like the reference
    if (i2 == 0x7fffffff) i2 = 0;
    if (i3 == 0x7fffffff) i3 = 0;
}
__device__ __forceinline__ void sp_group_search(const int* __restrict__ ws, const float4* __restrict__ sorted,
                                                const KnownRows& kr, int m, bool valid, float4 u,
                                                int sub, float& b1, float& b2, float& b3, int& i1, int& i2, int& i3) {
    sp_group_search_t<LPQ>(ws, sorted, kr, m, valid, u, sub, b1, b2, b3, i1, i2, i3, 0);
}


__global__ void __launch_bounds__(SP_THREADS) sp_three_nn_seg_kernel(int n, int m, const float* __restrict__ unknown,
                                                                     KnownRows kr,
                                                                     const int* __restrict__ ws,
                                                                     const float4* __restrict__ sorted,
                                                                     float* __restrict__ dist2,
                                                                     int* __restrict__ idx) {
    const int qi = blockIdx.x * (SP_THREADS / LPQ) + threadIdx.x / LPQ;
    const int sub = threadIdx.x % LPQ;
    const bool valid = qi < n;
    const float4 u = reinterpret_cast<const float4*>(unknown)[valid ? qi : (n - 1)];
    float b1, b2, b3;
    int i1, i2, i3;
    sp_group_search(ws, sorted, kr, m, valid, u, sub, b1, b2, b3, i1, i2, i3);
    if (valid && sub == 0) {
        dist2[qi * 3 + 0] = b1;
        dist2[qi * 3 + 1] = b2;
        dist2[qi * 3 + 2] = b3;
        idx[qi * 3 + 0] = i1;
        idx[qi * 3 + 1] = i2;
        idx[qi * 3 + 2] = i3;
    }
}

// ------------------------------------------------------------ interpolation
// out[i, c] = fma(w2,f2, fma(w0,f0, w1*f1)); one thread per (row, 4 channels).
__global__ void __launch_bounds__(256) sp_interp_v4_kernel(int c4, int n, const float4* __restrict__ points,
                                                           const int* __restrict__ idx,
                                                           const float* __restrict__ weight,
                                                           float4* __restrict__ out) {
    const long g = (long)blockIdx.x * 256 + threadIdx.x;
    if (g >= (long)n * c4) return;
    const int i = (int)(g / c4), cc = (int)(g - (long)i * c4);
    const int j0 = __ldg(idx + i * 3), j1 = __ldg(idx + i * 3 + 1), j2 = __ldg(idx + i * 3 + 2);
    const float w0 = __ldg(weight + i * 3), w1 = __ldg(weight + i * 3 + 1), w2 = __ldg(weight + i * 3 + 2);
    const float4 f0 = __ldg(points + (size_t)j0 * c4 + cc);
    const float4 f1 = __ldg(points + (size_t)j1 * c4 + cc);
    const float4 f2 = __ldg(points + (size_t)j2 * c4 + cc);
    float4 o;
    o.x = dcl_interp3(w0, f0.x, w1, f1.x, w2, f2.x);
    o.y = dcl_interp3(w0, f0.y, w1, f1.y, w2, f2.y);
    o.z = dcl_interp3(w0, f0.z, w1, f1.z, w2, f2.z);
    o.w = dcl_interp3(w0, f0.w, w1, f1.w, w2, f2.w);
    out[(size_t)i * c4 + cc] = o;
}

__global__ void __launch_bounds__(256) sp_interp_v1_kernel(int c, int n, const float* __restrict__ points,
                                                           const int* __restrict__ idx,
                                                           const float* __restrict__ weight,
                                                           float* __restrict__ out) {
    const long g = (long)blockIdx.x * 256 + threadIdx.x;
    if (g >= (long)n * c) return;
    const int i = (int)(g / c), cc = (int)(g - (long)i * c);
    const int j0 = idx[i * 3], j1 = idx[i * 3 + 1], j2 = idx[i * 3 + 2];
    out[(size_t)i * c + cc] =
        dcl_interp3(weight[i * 3], points[(size_t)j0 * c + cc], weight[i * 3 + 1], points[(size_t)j1 * c + cc],
                    weight[i * 3 + 2], points[(size_t)j2 * c + cc]);
}

// grad_points[idx[i,j], c] += grad_out[i, c] * w[i, j]
__global__ void __launch_bounds__(256) sp_interp_grad_v4_kernel(int c4, int n, const float4* __restrict__ grad_out,
                                                                const int* __restrict__ idx,
                                                                const float* __restrict__ weight,
                                                                float4* __restrict__ grad_points) {
    const long g = (long)blockIdx.x * 256 + threadIdx.x;
    if (g >= (long)n * c4) return;
    const int i = (int)(g / c4), cc = (int)(g - (long)i * c4);
    const float4 go = __ldg(grad_out + (size_t)i * c4 + cc);
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const int r = __ldg(idx + i * 3 + j);
        const float w = __ldg(weight + i * 3 + j);
        const float4 v = make_float4(__fmul_rn(go.x, w), __fmul_rn(go.y, w), __fmul_rn(go.z, w), __fmul_rn(go.w, w));
        atomicAdd(grad_points + (size_t)r * c4 + cc, v);  // red.global.add.v4.f32 (sm_90+)
    }
}

__global__ void __launch_bounds__(256) sp_interp_grad_v1_kernel(int c, int n, const float* __restrict__ grad_out,
                                                                const int* __restrict__ idx,
                                                                const float* __restrict__ weight,
                                                                float* __restrict__ grad_points) {
    const long g = (long)blockIdx.x * 256 + threadIdx.x;
    if (g >= (long)n * c) return;
    const int i = (int)(g / c), cc = (int)(g - (long)i * c);
    const float go = grad_out[(size_t)i * c + cc];
#pragma unroll
    for (int j = 0; j < 3; ++j)
        atomicAdd(grad_points + (size_t)idx[i * 3 + j] * c + cc, __fmul_rn(go, weight[i * 3 + j]));
}

// ------------------------------------------------ fused search + interpolation
// A group of LPQ lanes owns one query: segmented 3-NN (above), then weights as models/Modules.py:222-224
//   dist = sqrt(d2); r = 1/(dist+1e-8); w = r / sum(r)      (sum associated as torch's CUDA reduction does
//                                                             for a row of three: (r0 + r2) + r1)
// and the group's lanes spread over the channels of the output row (float4 each).
__global__ void __launch_bounds__(SP_THREADS) sp_nn_interp_fused_kernel(int n, int m, int c,
                                                                        const float* __restrict__ unknown,
                                                                        KnownRows kr,
                                                                        const int* __restrict__ ws,
                                                                        const float4* __restrict__ sorted,
                                                                        const float* __restrict__ feats,
                                                                        float* __restrict__ out, int out_stride,
                                                                        int out_col0, int vec_ok) {
    const int qi = blockIdx.x * (SP_THREADS / LPQ) + threadIdx.x / LPQ;
    const int sub = threadIdx.x % LPQ;
    const bool valid = qi < n;
    const float4 u = reinterpret_cast<const float4*>(unknown)[valid ? qi : (n - 1)];
    float b1, b2, b3;
    int j0, j1, j2;
    sp_group_search(ws, sorted, kr, m, valid, u, sub, b1, b2, b3, j0, j1, j2);
    if (!valid) return;
    const float r0 = __fdiv_rn(1.0f, __fadd_rn(__fsqrt_rn(b1), 1e-8f));
    const float r1 = __fdiv_rn(1.0f, __fadd_rn(__fsqrt_rn(b2), 1e-8f));
    const float r2 = __fdiv_rn(1.0f, __fadd_rn(__fsqrt_rn(b3), 1e-8f));
    const float norm = __fadd_rn(__fadd_rn(r0, r2), r1);
    const float a0 = __fdiv_rn(r0, norm), a1 = __fdiv_rn(r1, norm), a2 = __fdiv_rn(r2, norm);
    float* orow = out + (size_t)qi * out_stride + out_col0;
    if (vec_ok) {
        const float4* f0 = reinterpret_cast<const float4*>(feats + (size_t)j0 * c);
        const float4* f1 = reinterpret_cast<const float4*>(feats + (size_t)j1 * c);
        const float4* f2 = reinterpret_cast<const float4*>(feats + (size_t)j2 * c);
        for (int cc = sub; cc < (c >> 2); cc += LPQ) {
            const float4 x0 = __ldg(f0 + cc), x1 = __ldg(f1 + cc), x2 = __ldg(f2 + cc);
            float4 o;
            o.x = dcl_interp3(a0, x0.x, a1, x1.x, a2, x2.x);
            o.y = dcl_interp3(a0, x0.y, a1, x1.y, a2, x2.y);
            o.z = dcl_interp3(a0, x0.z, a1, x1.z, a2, x2.z);
            o.w = dcl_interp3(a0, x0.w, a1, x1.w, a2, x2.w);
            reinterpret_cast<float4*>(orow)[cc] = o;
        }
    } else {
        for (int cc = sub; cc < c; cc += LPQ)
            orow[cc] = dcl_interp3(a0, feats[(size_t)j0 * c + cc], a1, feats[(size_t)j1 * c + cc], a2,
                                   feats[(size_t)j2 * c + cc]);
    }
}

// Same, writing a PM image (see pm_gemm.cu): each lane of the group produces chunks of 8 channels and stores their
// bf16 hi / lo halves straight into the operand layout of the first disengage GEMM.
__global__ void __launch_bounds__(SP_THREADS) sp_nn_interp_fused_pm_kernel(int n, int m, int c,
                                                                           const float* __restrict__ unknown,
                                                                           KnownRows kr,
                                                                           const int* __restrict__ ws,
                                                                           const float4* __restrict__ sorted,
                                                                           const float* __restrict__ feats,
                                                                           unsigned char* __restrict__ out_pm,
                                                                           int c_total, int out_col0) {
    const int qi = blockIdx.x * (SP_THREADS / LPQ) + threadIdx.x / LPQ;
    const int sub = threadIdx.x % LPQ;
    const bool valid = qi < n;
    const float4 u = reinterpret_cast<const float4*>(unknown)[valid ? qi : (n - 1)];
    float b1, b2, b3;
    int j0, j1, j2;
    sp_group_search(ws, sorted, kr, m, valid, u, sub, b1, b2, b3, j0, j1, j2);
    if (!valid) return;
    const float r0 = __fdiv_rn(1.0f, __fadd_rn(__fsqrt_rn(b1), 1e-8f));
    const float r1 = __fdiv_rn(1.0f, __fadd_rn(__fsqrt_rn(b2), 1e-8f));
    const float r2 = __fdiv_rn(1.0f, __fadd_rn(__fsqrt_rn(b3), 1e-8f));
    const float norm = __fadd_rn(__fadd_rn(r0, r2), r1);
    const float a0 = __fdiv_rn(r0, norm), a1 = __fdiv_rn(r1, norm), a2 = __fdiv_rn(r2, norm);
    const float4* f0 = reinterpret_cast<const float4*>(feats + (size_t)j0 * c);
    const float4* f1 = reinterpret_cast<const float4*>(feats + (size_t)j1 * c);
    const float4* f2 = reinterpret_cast<const float4*>(feats + (size_t)j2 * c);
    unsigned char* row_base = out_pm + (size_t)(qi / 128) * (c_total / 32) * 16384 + ((qi % 128) >> 3) * 512 + (qi & 7) * 16;
    for (int c8 = sub; c8 < (c >> 3); c8 += LPQ) {
        float o[8];
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
            const float4 x0 = __ldg(f0 + c8 * 2 + hh), x1 = __ldg(f1 + c8 * 2 + hh), x2 = __ldg(f2 + c8 * 2 + hh);
            o[hh * 4 + 0] = dcl_interp3(a0, x0.x, a1, x1.x, a2, x2.x);
            o[hh * 4 + 1] = dcl_interp3(a0, x0.y, a1, x1.y, a2, x2.y);
            o[hh * 4 + 2] = dcl_interp3(a0, x0.z, a1, x1.z, a2, x2.z);
            o[hh * 4 + 3] = dcl_interp3(a0, x0.w, a1, x1.w, a2, x2.w);
        }
        __nv_bfloat16 h[8], l[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) split_bf16(o[e], h[e], l[e]);
        const int chunk = (out_col0 >> 3) + c8;
        unsigned char* d = row_base + (size_t)(chunk >> 2) * 16384 + (chunk & 3) * 128;
        *reinterpret_cast<uint4*>(d) = make_uint4(pack2(h[0], h[1]), pack2(h[2], h[3]), pack2(h[4], h[5]), pack2(h[6], h[7]));
        *reinterpret_cast<uint4*>(d + 8192) = make_uint4(pack2(l[0], l[1]), pack2(l[2], l[3]), pack2(l[4], l[5]), pack2(l[6], l[7]));
    }
}

KnownRows rows_from_float(const float* known) {
    KnownRows kr = {};
    kr.rows = known;
    return kr;
}

KnownRows rows_from_voxels(const int* vox, const float* ext3, const float* off3) {
    KnownRows kr = {};
    kr.vox = vox;
    for (int i = 0; i < 3; ++i) {
        kr.ext[i] = ext3[i];
        kr.off[i] = off3[i];
        kr.half[i] = 0.5f * ext3[i];
    }
    return kr;
}

int build_buckets(int m, const KnownRows& kr, int* ws, float4* sorted, cudaStream_t st) {
    // only the header words that are read before being written need clearing
    cudaError_t e = cudaMemsetAsync(ws, 0, (size_t)(WS_OFF + DCL_SP_MAX_BATCH + 4) * sizeof(int), st);
    if (e != cudaSuccess) return (int)e;
    if (m > 0) sp_bucket_hist_kernel<<<DCL_DIVUP(m, 256), 256, 0, st>>>(m, kr, ws);
    sp_bucket_scan_kernel<<<1, 1024, 0, st>>>(ws);
    if (m > 0) sp_bucket_scatter_kernel<<<DCL_DIVUP(m, 256), 256, 0, st>>>(m, kr, ws, sorted);
    return dcl_launch_status(m > 0 ? 3 : 1);
}

// ------------------------------------------------ all pyramid levels of one tower in two launches
// Ops_GetPointFeat_spconv (models/Modules.py:227-251) interpolates the same query points from four voxel levels.
// One launch builds all the levels' buckets — a cluster of 8 CTAs per level, whose per-CTA histograms meet through
// distributed shared memory, so there is no global header to clear and no separate scan / scatter pass — and one
// launch runs search + interpolation for every level (a CTA keeps its 16 queries and walks the levels).
constexpr int SPL_MAX_LEVELS = 8;
constexpr int SPB_THREADS = 512, SPB_CLUSTER = 8, SPB_ITEMS = 10;
constexpr int SPL_SBINS = 2048;                        // buckets per level: batch ids x slabs
constexpr int SPL_HDR_INTS = WS_OFF + SPL_SBINS + 4;  // flag, nb, offsets[SPL_SBINS + 1]; multiple of 4

struct SpLevelDev {
    int m, c, out_col0;
    int gx;  // > 0: buckets keyed by (batch id, first voxel index in [0, gx)); 0: by batch id only
    KnownRows kr;
    int* ws;
    float4* sorted;
    const float* feats;
};
constexpr int SPL_MAX_TOWERS = 2;
struct SpTowerDev {  // one set of query points and the pyramid levels [lv0, lv1) it interpolates from
    int n, c_total, lv0, lv1;
    const float* unknown;
    unsigned char* out_pm;
    int out_fmt;  // 0: PM image (bf16 hi/lo, 16 KB blobs); 1: PM16 image (fp16, 8 KB blobs)
};
struct SpLevelBatch {
    int nlevels, ntowers;
    SpTowerDev tw[SPL_MAX_TOWERS];
    SpLevelDev lv[SPL_MAX_LEVELS];
};

__global__ void __cluster_dims__(SPB_CLUSTER, 1, 1) __launch_bounds__(SPB_THREADS, 1)
    sp_bucket_build_cluster_kernel(const __grid_constant__ SpLevelBatch batch) {
    __shared__ int s_cnt[SPL_SBINS];   // this CTA's histogram; the other CTAs of the cluster read it through DSMEM
    __shared__ int s_base[SPL_SBINS];  // where this CTA's entries of bucket b start in sorted[]
    __shared__ int s_cur[SPL_SBINS];   // bucket totals, then scatter cursors
    __shared__ int s_meta[2];          // [0] an id outside the bucket range, [1] max batch id + 1
    __shared__ int s_warp[SPB_THREADS / 32];
    const SpLevelDev& lv = batch.lv[blockIdx.y];
    const uint32_t rank = dcl_cluster_ctarank();
    const int tid = threadIdx.x, lane = tid & 31;
    for (int i = tid; i < SPL_SBINS; i += SPB_THREADS) s_cnt[i] = 0;
    if (tid < 2) s_meta[tid] = 0;
    __syncthreads();
    const int m = lv.m;
    const int gxe = lv.gx > 0 ? lv.gx : 1;
    const int slice = DCL_DIVUP(m, SPB_CLUSTER);
    const int k0 = (int)rank * slice, k1 = min(m, k0 + slice);
    // Slices of up to SPB_ITEMS points per thread (m <= 40960 per level) stay in registers between the two passes
    // and all their loads are in flight at once; longer slices are re-read in pass 2.
    const bool in_regs = (k1 - k0) <= SPB_THREADS * SPB_ITEMS;
    float4 held[SPB_ITEMS];
    int held_b[SPB_ITEMS], held_s[SPB_ITEMS];
    if (in_regs) {
#pragma unroll
        for (int t = 0; t < SPB_ITEMS; ++t) {
            const int k = k0 + t * SPB_THREADS + tid;
            held[t] = (k < k1) ? lv.kr.get(k) : make_float4(-1.f, 0.f, 0.f, 0.f);
            held_s[t] = (k < k1 && lv.gx > 0) ? lv.kr.slab(k) : 0;
        }
    }
    // pass 1: histogram of this CTA's slice; lanes with the same bucket (the usual case: clouds are stored
    // batch-major) elect one lane to add their count
    // bucket of a point: batch id * slabs + slab; -1 (and the level's flag) when either is out of range
    auto classify = [&](float bf, int slab, bool live) {
        int ib = -1;
        if (!live) return -1;
        if (!batch_id_ok(bf, ib) || slab < 0 || slab >= gxe || ib >= SPL_SBINS / gxe) {
            s_meta[0] = 1;
            return -1;
        }
        return ib * gxe + slab;
    };
    auto count_one = [&](int key) {
        const unsigned peers = __match_any_sync(0xffffffffu, key);
        if (key >= 0 && lane == __ffs(peers) - 1) {
            atomicAdd(s_cnt + key, __popc(peers));
            atomicMax(&s_meta[1], key / gxe + 1);
        }
    };
    if (in_regs) {
#pragma unroll
        for (int t = 0; t < SPB_ITEMS; ++t) {
            held_b[t] = classify(held[t].x, held_s[t], k0 + t * SPB_THREADS + tid < k1);
            count_one(held_b[t]);
        }
    } else {
        for (int kb = k0; kb < k1; kb += SPB_THREADS) {
            const int k = kb + tid;
            const bool live = k < k1;
            count_one(classify(live ? lv.kr.get(k).x : 0.f, (live && lv.gx > 0) ? lv.kr.slab(k) : 0, live));
        }
    }
    __syncthreads();
    dcl_cluster_sync();
    // combine: bucket totals and the number of entries the lower-ranked CTAs put into each bucket (only the
    // buckets in use: every CTA first learns the largest batch id any of them saw)
    int flag = 0, nb = 0;
#pragma unroll
    for (uint32_t r = 0; r < SPB_CLUSTER; ++r) {
        flag |= dcl_ld_dsmem_s32(&s_meta[0], r);
        nb = max(nb, dcl_ld_dsmem_s32(&s_meta[1], r));
    }
    for (int b = tid; b < SPL_SBINS; b += SPB_THREADS) {
        int total = 0, before = 0;
        if (b < nb * gxe) {
#pragma unroll
            for (uint32_t r = 0; r < SPB_CLUSTER; ++r) {
                const int cnt = dcl_ld_dsmem_s32(s_cnt + b, r);
                total += cnt;
                before += (r < rank) ? cnt : 0;
            }
        }
        s_cur[b] = total;
        s_base[b] = before;
    }
    __syncthreads();
    dcl_cluster_sync();  // nobody reads a peer's shared memory after this point
    // exclusive scan of the totals (SPL_SBINS / SPB_THREADS consecutive buckets per thread)
    {
        constexpr int PER = SPL_SBINS / SPB_THREADS;
        int v[PER], x = 0;
#pragma unroll
        for (int i = 0; i < PER; ++i) {
            v[i] = s_cur[PER * tid + i];
            x += v[i];
        }
        const int mine = x;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) s_warp[tid >> 5] = x;
        __syncthreads();
        if (tid < 32) {
            int w = (tid < SPB_THREADS / 32) ? s_warp[tid] : 0;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int y = __shfl_up_sync(0xffffffffu, w, o);
                if (tid >= o) w += y;
            }
            if (tid < SPB_THREADS / 32) s_warp[tid] = w;
        }
        __syncthreads();
        int run = x - mine + ((tid >= 32) ? s_warp[(tid >> 5) - 1] : 0);
#pragma unroll
        for (int i = 0; i < PER; ++i) {
            s_base[PER * tid + i] += run;
            if (rank == 0) lv.ws[WS_OFF + PER * tid + i] = run;
            run += v[i];
        }
        if (rank == 0) {
            if (tid == SPB_THREADS - 1) lv.ws[WS_OFF + SPL_SBINS] = run;
            if (tid == 0) {
                lv.ws[WS_FLAG] = flag;
                lv.ws[WS_NB] = nb;
            }
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < PER; ++i) s_cur[PER * tid + i] = 0;
        __syncthreads();
    }
    // pass 2: scatter this CTA's slice
    auto place_one = [&](const float4& r, int ib, int k) {
        const unsigned peers = __match_any_sync(0xffffffffu, ib);
        const int leader = __ffs(peers) - 1;
        int first = 0;
        if (ib >= 0 && lane == leader) first = atomicAdd(s_cur + ib, __popc(peers));
        first = __shfl_sync(0xffffffffu, first, leader);
        if (ib >= 0) {
            const int pos = s_base[ib] + first + __popc(peers & ((1u << lane) - 1u));
            lv.sorted[pos] = make_float4(r.y, r.z, r.w, __int_as_float(k));
        }
    };
    if (in_regs) {
#pragma unroll
        for (int t = 0; t < SPB_ITEMS; ++t) place_one(held[t], held_b[t], k0 + t * SPB_THREADS + tid);
    } else {
        for (int kb = k0; kb < k1; kb += SPB_THREADS) {
            const int k = kb + tid;
            const bool live = k < k1;
            const float4 r = live ? lv.kr.get(k) : make_float4(0.f, 0.f, 0.f, 0.f);
            place_one(r, classify(r.x, (live && lv.gx > 0) ? lv.kr.slab(k) : 0, live), k);
        }
    }
}

// Lanes per query in the multi-level kernel: the slab walk leaves few candidates per query, so the per-query work
// every lane of a group repeats (merge butterfly, interpolation weights) weighs more than the scan; four lanes
// halve it against the eight of the whole-instance kernels.
constexpr int SPL_LPQ = 4;
__global__ void __launch_bounds__(SP_THREADS)
    sp_nn_interp_levels_pm_kernel(const __grid_constant__ SpLevelBatch batch) {
    // grid = (query groups, towers)
    const SpTowerDev& tw = batch.tw[blockIdx.y];
    const int n = tw.n, c_total = tw.c_total;
    if (blockIdx.x * (SP_THREADS / SPL_LPQ) >= n) return;
    const float* __restrict__ unknown = tw.unknown;
    unsigned char* __restrict__ out_pm = tw.out_pm;
    const int qi = blockIdx.x * (SP_THREADS / SPL_LPQ) + threadIdx.x / SPL_LPQ;
    const int sub = threadIdx.x % SPL_LPQ;
    const bool valid = qi < n;
    const float4 u = reinterpret_cast<const float4*>(unknown)[valid ? qi : (n - 1)];
    const bool f16 = tw.out_fmt == 1;
    const size_t blob = f16 ? 8192 : 16384;
    unsigned char* row_base = out_pm + (size_t)(qi / 128) * (c_total / 32) * blob + ((qi % 128) >> 3) * 512 + (qi & 7) * 16;
    for (int li = tw.lv0; li < tw.lv1; ++li) {
        const SpLevelDev& lv = batch.lv[li];
        if (lv.m == 0) continue;  // nothing to interpolate from (the reference would gather row 0 of an empty tensor)
        float b1, b2, b3;
        int j0, j1, j2;
        sp_group_search_t<SPL_LPQ>(lv.ws, lv.sorted, lv.kr, lv.m, valid, u, sub, b1, b2, b3, j0, j1, j2, lv.gx);
        if (valid) {
            const int c = lv.c;
            const float r0 = __fdiv_rn(1.0f, __fadd_rn(__fsqrt_rn(b1), 1e-8f));
            const float r1 = __fdiv_rn(1.0f, __fadd_rn(__fsqrt_rn(b2), 1e-8f));
            const float r2 = __fdiv_rn(1.0f, __fadd_rn(__fsqrt_rn(b3), 1e-8f));
            const float norm = __fadd_rn(__fadd_rn(r0, r2), r1);
            const float a0 = __fdiv_rn(r0, norm), a1 = __fdiv_rn(r1, norm), a2 = __fdiv_rn(r2, norm);
            const float4* f0 = reinterpret_cast<const float4*>(lv.feats + (size_t)j0 * c);
            const float4* f1 = reinterpret_cast<const float4*>(lv.feats + (size_t)j1 * c);
            const float4* f2 = reinterpret_cast<const float4*>(lv.feats + (size_t)j2 * c);
#pragma unroll 2
            for (int c8 = sub; c8 < (c >> 3); c8 += SPL_LPQ) {
                float o[8];
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                    const float4 x0 = __ldg(f0 + c8 * 2 + hh), x1 = __ldg(f1 + c8 * 2 + hh), x2 = __ldg(f2 + c8 * 2 + hh);
                    o[hh * 4 + 0] = dcl_interp3(a0, x0.x, a1, x1.x, a2, x2.x);
                    o[hh * 4 + 1] = dcl_interp3(a0, x0.y, a1, x1.y, a2, x2.y);
                    o[hh * 4 + 2] = dcl_interp3(a0, x0.z, a1, x1.z, a2, x2.z);
                    o[hh * 4 + 3] = dcl_interp3(a0, x0.w, a1, x1.w, a2, x2.w);
                }
                uint32_t h[4], l[4];
                const int chunk = (lv.out_col0 >> 3) + c8;
                unsigned char* d = row_base + (size_t)(chunk >> 2) * blob + (chunk & 3) * 128;
                if (f16) {
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const __half2 hh = __floats2half2_rn(fminf(fmaxf(o[2 * e], -65504.f), 65504.f),
                                                             fminf(fmaxf(o[2 * e + 1], -65504.f), 65504.f));
                        h[e] = *reinterpret_cast<const uint32_t*>(&hh);
                    }
                    *reinterpret_cast<uint4*>(d) = make_uint4(h[0], h[1], h[2], h[3]);
                } else {
#pragma unroll
                    for (int e = 0; e < 4; ++e) split2_bf16(o[2 * e], o[2 * e + 1], h[e], l[e]);
                    *reinterpret_cast<uint4*>(d) = make_uint4(h[0], h[1], h[2], h[3]);
                    *reinterpret_cast<uint4*>(d + 8192) = make_uint4(l[0], l[1], l[2], l[3]);
                }
            }
        }
    }
}

size_t sp_level_ws_bytes(int m) {
    return (size_t)SPL_HDR_INTS * sizeof(int) + (size_t)(m > 0 ? m : 0) * sizeof(float4) + 16;
}

}  // namespace

DCL_API int dcl_sp_three_nn_kernel_launcher_fast(int n, int m, const float* unknown, const float* known,
                                                 float* dist2, int* idx, void* stream) {
    DCL_RETURN_IF_BAD(n >= 0 && m >= 0);
    DCL_RETURN_IF_BAD(((uintptr_t)unknown & 15u) == 0);
    if (n == 0) return 0;
    sp_three_nn_scan_kernel<<<DCL_DIVUP(n, SP_THREADS), SP_THREADS, 0, (cudaStream_t)stream>>>(
        n, m, unknown, known, dist2, idx);
    return dcl_launch_status();
}

DCL_API size_t dcl_sp_three_nn_workspace_bytes(int n, int m) {
    (void)n;
    return (size_t)WS_HDR_INTS * sizeof(int) + (size_t)(m > 0 ? m : 0) * sizeof(float4) + 16;
}

DCL_API int dcl_sp_three_nn_segmented(int n, int m, const float* unknown, const float* known, float* dist2, int* idx,
                                      void* workspace, size_t workspace_bytes, void* stream) {
    DCL_RETURN_IF_BAD(n >= 0 && m >= 0 && workspace != nullptr);
    DCL_RETURN_IF_BAD(workspace_bytes >= dcl_sp_three_nn_workspace_bytes(n, m));
    DCL_RETURN_IF_BAD(((uintptr_t)workspace & 15u) == 0 && ((uintptr_t)unknown & 15u) == 0 &&
                      ((uintptr_t)known & 15u) == 0);
    if (n == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    int* ws = (int*)workspace;
    float4* sorted = (float4*)(ws + WS_HDR_INTS);
    const KnownRows kr = rows_from_float(known);
    int err = build_buckets(m, kr, ws, sorted, st);
    if (err) return err;
    const int grid = DCL_DIVUP(n, SP_THREADS / LPQ);
    sp_three_nn_seg_kernel<<<grid, SP_THREADS, 0, st>>>(n, m, unknown, kr, ws, sorted, dist2, idx);
    return dcl_launch_status();
}

DCL_API int dcl_sp_three_interpolate_kernel_launcher_fast(int c, int m, int n, const float* points, const int* idx,
                                                          const float* weight, float* out, void* stream) {
    DCL_RETURN_IF_BAD(c >= 0 && m >= 0 && n >= 0);
    if (c == 0 || n == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const bool vec = ((c & 3) == 0) && ((((uintptr_t)points) & 15u) == 0) && ((((uintptr_t)out) & 15u) == 0);
    if (vec) {
        const long total = (long)n * (c >> 2);
        sp_interp_v4_kernel<<<(unsigned)DCL_DIVUP(total, 256L), 256, 0, st>>>(c >> 2, n, (const float4*)points, idx,
                                                                              weight, (float4*)out);
    } else {
        const long total = (long)n * c;
        sp_interp_v1_kernel<<<(unsigned)DCL_DIVUP(total, 256L), 256, 0, st>>>(c, n, points, idx, weight, out);
    }
    return dcl_launch_status();
}

DCL_API int dcl_sp_three_interpolate_grad_kernel_launcher_fast(int c, int n, int m, const float* grad_out,
                                                               const int* idx, const float* weight,
                                                               float* grad_points, void* stream) {
    DCL_RETURN_IF_BAD(c >= 0 && m >= 0 && n >= 0);
    if (c == 0 || n == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const bool vec =
        ((c & 3) == 0) && ((((uintptr_t)grad_out) & 15u) == 0) && ((((uintptr_t)grad_points) & 15u) == 0);
    if (vec) {
        const long total = (long)n * (c >> 2);
        sp_interp_grad_v4_kernel<<<(unsigned)DCL_DIVUP(total, 256L), 256, 0, st>>>(
            c >> 2, n, (const float4*)grad_out, idx, weight, (float4*)grad_points);
    } else {
        const long total = (long)n * c;
        sp_interp_grad_v1_kernel<<<(unsigned)DCL_DIVUP(total, 256L), 256, 0, st>>>(c, n, grad_out, idx, weight,
                                                                                   grad_points);
    }
    return dcl_launch_status();
}

DCL_API int dcl_sp_nn_interpolate_fused(int n, int m, int c, const float* unknown, const float* known,
                                        const float* feats, float* out, int out_stride, int out_col0,
                                        void* workspace, size_t workspace_bytes, void* stream) {
    DCL_RETURN_IF_BAD(n >= 0 && m >= 0 && c >= 0 && workspace != nullptr && out_stride >= out_col0 + c);
    DCL_RETURN_IF_BAD(workspace_bytes >= dcl_sp_three_nn_workspace_bytes(n, m));
    DCL_RETURN_IF_BAD(((uintptr_t)workspace & 15u) == 0 && ((uintptr_t)unknown & 15u) == 0 &&
                      ((uintptr_t)known & 15u) == 0);
    if (n == 0 || c == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    int* ws = (int*)workspace;
    float4* sorted = (float4*)(ws + WS_HDR_INTS);
    const KnownRows kr = rows_from_float(known);
    int err = build_buckets(m, kr, ws, sorted, st);
    if (err) return err;
    const int vec_ok = ((c & 3) == 0) && ((out_stride & 3) == 0) && ((out_col0 & 3) == 0) &&
                       ((((uintptr_t)feats) & 15u) == 0) && ((((uintptr_t)out) & 15u) == 0);
    sp_nn_interp_fused_kernel<<<DCL_DIVUP(n, SP_THREADS / LPQ), SP_THREADS, 0, st>>>(
        n, m, c, unknown, kr, ws, sorted, feats, out, out_stride, out_col0, vec_ok);
    return dcl_launch_status();
}

DCL_API int dcl_sp_nn_interpolate_fused_pm(int n, int m, int c, const float* unknown, const float* known,
                                           const float* feats, void* out_pm, int c_total, int out_col0,
                                           void* workspace, size_t workspace_bytes, void* stream) {
    DCL_RETURN_IF_BAD(n > 0 && n % 128 == 0 && m >= 0 && c > 0 && c % 8 == 0 && out_col0 % 8 == 0);
    DCL_RETURN_IF_BAD(c_total % 32 == 0 && c_total >= out_col0 + c && workspace != nullptr && out_pm != nullptr);
    DCL_RETURN_IF_BAD(workspace_bytes >= dcl_sp_three_nn_workspace_bytes(n, m));
    DCL_RETURN_IF_BAD(((uintptr_t)workspace & 15u) == 0 && ((uintptr_t)unknown & 15u) == 0 &&
                      ((uintptr_t)known & 15u) == 0 && ((uintptr_t)feats & 15u) == 0 && ((uintptr_t)out_pm & 15u) == 0);
    cudaStream_t st = (cudaStream_t)stream;
    int* ws = (int*)workspace;
    float4* sorted = (float4*)(ws + WS_HDR_INTS);
    const KnownRows kr = rows_from_float(known);
    int err = build_buckets(m, kr, ws, sorted, st);
    if (err) return err;
    sp_nn_interp_fused_pm_kernel<<<DCL_DIVUP(n, SP_THREADS / LPQ), SP_THREADS, 0, st>>>(
        n, m, c, unknown, kr, ws, sorted, feats, reinterpret_cast<unsigned char*>(out_pm), c_total, out_col0);
    return dcl_launch_status();
}

DCL_API int dcl_sp_nn_interpolate_vox_pm(int n, int m, int c, const float* unknown, const int* vox_indices,
                                         const float* voxel_extent3, const float* offset3, const float* feats,
                                         void* out_pm, int c_total, int out_col0, void* workspace,
                                         size_t workspace_bytes, void* stream) {
    DCL_RETURN_IF_BAD(n > 0 && n % 128 == 0 && m >= 0 && c > 0 && c % 8 == 0 && out_col0 % 8 == 0);
    DCL_RETURN_IF_BAD(c_total % 32 == 0 && c_total >= out_col0 + c && workspace != nullptr && out_pm != nullptr);
    DCL_RETURN_IF_BAD(voxel_extent3 != nullptr && offset3 != nullptr);
    DCL_RETURN_IF_BAD(workspace_bytes >= dcl_sp_three_nn_workspace_bytes(n, m));
    DCL_RETURN_IF_BAD(((uintptr_t)workspace & 15u) == 0 && ((uintptr_t)unknown & 15u) == 0 &&
                      ((uintptr_t)vox_indices & 15u) == 0 && ((uintptr_t)feats & 15u) == 0 &&
                      ((uintptr_t)out_pm & 15u) == 0);
    cudaStream_t st = (cudaStream_t)stream;
    int* ws = (int*)workspace;
    float4* sorted = (float4*)(ws + WS_HDR_INTS);
    const KnownRows kr = rows_from_voxels(vox_indices, voxel_extent3, offset3);
    int err = build_buckets(m, kr, ws, sorted, st);
    if (err) return err;
    sp_nn_interp_fused_pm_kernel<<<DCL_DIVUP(n, SP_THREADS / LPQ), SP_THREADS, 0, st>>>(
        n, m, c, unknown, kr, ws, sorted, feats, reinterpret_cast<unsigned char*>(out_pm), c_total, out_col0);
    return dcl_launch_status();
}

DCL_API size_t dcl_sp_levels_workspace_bytes(int nlevels, const dcl_sp_level* levels) {
    if (nlevels < 0 || nlevels > SPL_MAX_LEVELS || (nlevels > 0 && levels == nullptr)) return 0;
    size_t total = 0;
    for (int i = 0; i < nlevels; ++i) total += sp_level_ws_bytes(levels[i].m);
    return total;
}

DCL_API int dcl_sp_nn_interpolate_towers_pm(int ntowers, const dcl_sp_tower* towers, void* workspace,
                                            size_t workspace_bytes, void* stream) {
    DCL_RETURN_IF_BAD(ntowers >= 1 && ntowers <= SPL_MAX_TOWERS && towers != nullptr && workspace != nullptr);
    DCL_RETURN_IF_BAD(((uintptr_t)workspace & 15u) == 0);
    SpLevelBatch batch = {};
    batch.ntowers = ntowers;
    unsigned char* w = reinterpret_cast<unsigned char*>(workspace);
    size_t need = 0;
    int nmax = 0, li = 0;
    for (int t = 0; t < ntowers; ++t) {
        const dcl_sp_tower& tin = towers[t];
        DCL_RETURN_IF_BAD(tin.n > 0 && tin.n % 128 == 0 && tin.nlevels >= 1 && tin.levels != nullptr &&
                          li + tin.nlevels <= SPL_MAX_LEVELS);
        DCL_RETURN_IF_BAD(tin.c_total % 32 == 0 && tin.out_pm != nullptr && tin.unknown != nullptr);
        DCL_RETURN_IF_BAD(((uintptr_t)tin.unknown & 15u) == 0 && ((uintptr_t)tin.out_pm & 15u) == 0);
        SpTowerDev& tw = batch.tw[t];
        tw.n = tin.n;
        tw.c_total = tin.c_total;
        tw.unknown = tin.unknown;
        tw.out_pm = reinterpret_cast<unsigned char*>(tin.out_pm);
        DCL_RETURN_IF_BAD(tin.out_fmt == 0 || tin.out_fmt == 1);
        tw.out_fmt = tin.out_fmt;
        tw.lv0 = li;
        for (int i = 0; i < tin.nlevels; ++i, ++li) {
            const dcl_sp_level& in = tin.levels[i];
            DCL_RETURN_IF_BAD(in.m >= 0 && in.c > 0 && in.c % 8 == 0 && in.out_col0 >= 0 && in.out_col0 % 8 == 0 &&
                              tin.c_total >= in.out_col0 + in.c);
            DCL_RETURN_IF_BAD(in.m == 0 || (in.vox_indices != nullptr && in.feats != nullptr));
            DCL_RETURN_IF_BAD(((uintptr_t)in.vox_indices & 15u) == 0 && ((uintptr_t)in.feats & 15u) == 0);
            need += sp_level_ws_bytes(in.m);
            DCL_RETURN_IF_BAD(need <= workspace_bytes);
            SpLevelDev& lv = batch.lv[li];
            lv.m = in.m;
            lv.c = in.c;
            lv.out_col0 = in.out_col0;
            lv.gx = (in.grid_x > 0 && in.grid_x <= SP_SLAB_MAX_GX) ? in.grid_x : 0;
            lv.kr = rows_from_voxels(in.vox_indices, in.voxel_extent, in.offset);
            lv.ws = reinterpret_cast<int*>(w);
            lv.sorted = reinterpret_cast<float4*>(w + (size_t)SPL_HDR_INTS * sizeof(int));
            lv.feats = in.feats;
            w += sp_level_ws_bytes(in.m);
        }
        tw.lv1 = li;
        nmax = tin.n > nmax ? tin.n : nmax;
    }
    batch.nlevels = li;
    cudaStream_t st = (cudaStream_t)stream;
    sp_bucket_build_cluster_kernel<<<dim3(SPB_CLUSTER, batch.nlevels), SPB_THREADS, 0, st>>>(batch);
    sp_nn_interp_levels_pm_kernel<<<dim3(DCL_DIVUP(nmax, SP_THREADS / SPL_LPQ), ntowers), SP_THREADS, 0, st>>>(batch);
    return dcl_launch_status(2);
}

DCL_API int dcl_sp_nn_interpolate_levels_pm(int n, const float* unknown, int nlevels, const dcl_sp_level* levels,
                                            void* out_pm, int c_total, void* workspace, size_t workspace_bytes,
                                            void* stream) {
    DCL_RETURN_IF_BAD(nlevels >= 1 && nlevels <= SPL_MAX_LEVELS && levels != nullptr);
    const dcl_sp_tower tower = {n, c_total, nlevels, unknown, out_pm, levels, 0};
    return dcl_sp_nn_interpolate_towers_pm(1, &tower, workspace, workspace_bytes, stream);
}
