// Training path of the pointwise MLP stacks (models/DCL_Net.py:56-151, models/Modules.py:58-97,173-201 in train
// mode; the reference runs them as cuDNN/cuBLAS calls under autograd).  The three GEMMs of every layer
//      forward  U  = X W^T (+ bias, ReLU)             dcl_pm_gemm on the PM image of X
//      dgrad    dX = dZ W                             dcl_pm_gemm on the PM image of dZ, "weights" = packed W^T
//      wgrad    dW = dZ^T X = sum_b dZ_b^T X_b        dcl_pm_gemm, strided batch: per instance b the TRANSPOSED images
//                                                     of dZ_b (rows = channels, K = points) and of X_b; a PM image with
//                                                     128-row tiles is byte-identical to packed weights of n-tile 128
// run on the tcgen05 kernel of pm_gemm.cu with bf16 hi/lo operands (3 MMAs per product: fp32-faithful, no loss
// scaling needed for the gradients).  This file holds what surrounds them — everything HBM-bound and elementwise:
//   * tr_tile_kernel: one pass over an fp32 activation (or gradient) that applies the layer's pointwise transform
//     (train-mode BatchNorm affine, ReLU, their backward) and emits the operand images the GEMMs read: the PM image,
//     the per-instance transposed images, optionally the fp32 tensor and per-tile column sums (bias gradient).
//     A CTA owns a tile of 32 channels x 128 points staged in shared memory, so that both images leave as whole
//     16-byte units in runs of >= 512 contiguous bytes whatever the layout of the source.
//   * tr_bn_stats_kernel / tr_bn_bwd_reduce_kernel: per-channel batch statistics and the two sums of the BatchNorm
//     backward, accumulated in fp64 per thread and reduced in a fixed order (deterministic).
//   * tr_pack_weights_kernel: fp32 weights (or their transpose, zero-padded) -> packed bf16 hi/lo blobs.
#include "common.cuh"
#include "umma.cuh"
#include "../../include/dcl_b200.h"

namespace {

constexpr int TR_MAX_ITEMS = 8;
constexpr int TR_TC = 32;          // channels per tile
constexpr int TR_TP = 128;         // points per tile
constexpr int TR_LD = TR_TP + 4;   // shared-memory row stride (floats): 16-byte aligned rows, conflict-free float4 accesses
constexpr int TR_BLOB = 16384;     // PM blob: bf16 hi image (8 KB) + lo image (8 KB) of 128 rows x 32 channels

struct TrTileBatch { dcl_tr_tile it[TR_MAX_ITEMS]; };
struct TrBnBatch { dcl_tr_bn it[TR_MAX_ITEMS]; };
struct TrBnBwdBatch { dcl_tr_bn_bwd it[TR_MAX_ITEMS]; };
struct TrWpackBatch { dcl_tr_wpack it[TR_MAX_ITEMS]; };
struct TrColsumBatch { dcl_tr_colsum it[TR_MAX_ITEMS]; };

__device__ __forceinline__ uint4 pack8_hi_lo(const float* v, uint4& lo) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) split2_bf16(v[2 * e], v[2 * e + 1], h[e], l[e]);
    lo = make_uint4(l[0], l[1], l[2], l[3]);
    return make_uint4(h[0], h[1], h[2], h[3]);
}

// ------------------------------------------------------------------ tile transform + operand images
// the pointwise transform of one element; `cg` = its channel
struct TrChan { float sc, sh, mean, rstd, k1, k2; };
__device__ __forceinline__ TrChan tr_chan(const dcl_tr_tile& it, int cg, float inv_cnt) {
    TrChan p = {1.f, 0.f, 0.f, 1.f, 0.f, 0.f};
    if (it.mode != DCL_TR_COPY && it.mode != DCL_TR_BWD_RELU) {
        p.sc = __ldg(it.scale + cg);
        p.sh = __ldg(it.shift + cg);
    }
    if (it.mode >= DCL_TR_BWD_BN_RELU) {
        p.mean = __ldg(it.mean + cg);
        p.rstd = __ldg(it.rstd + cg);
        p.k1 = __ldg(it.s1 + cg) * inv_cnt;
        p.k2 = __ldg(it.s2 + cg) * inv_cnt;
    }
    return p;
}
__device__ __forceinline__ float tr_apply(int mode, float x, float u, const TrChan& p) {
    switch (mode) {
        case DCL_TR_COPY: return x;
        case DCL_TR_AFFINE: return __fmaf_rn(x, p.sc, p.sh);
        case DCL_TR_AFFINE_RELU: return fmaxf(__fmaf_rn(x, p.sc, p.sh), 0.f);
        case DCL_TR_BWD_RELU: return u > 0.f ? x : 0.f;
        case DCL_TR_BWD_BN_RELU: {
            const float g = __fmaf_rn(u, p.sc, p.sh) > 0.f ? x : 0.f;
            return p.sc * (g - p.k1 - (u - p.mean) * p.rstd * p.k2);
        }
        default:  // DCL_TR_BWD_RELU_BN
            return u > 0.f ? p.sc * (x - p.k1 - (u - p.mean) * p.rstd * p.k2) : 0.f;
    }
}

// VEC: 0 = scalar loads (any strides), 1 = channel-major source read as float4 along the points, 2 = point-major
// source read as float4 along the channels (16-byte aligned rows; checked on the host).
template <int VEC>
__global__ void __launch_bounds__(256) tr_tile_kernel(const __grid_constant__ TrTileBatch batch) {
    __shared__ __align__(16) float tile[TR_TC * TR_LD];
    const dcl_tr_tile& it = batch.it[blockIdx.z];
    const int tiles_per_inst = it.n / TR_TP;
    const int inst = blockIdx.x / tiles_per_inst, n0 = (blockIdx.x % tiles_per_inst) * TR_TP;
    const int c0 = blockIdx.y * TR_TC;
    if (c0 >= it.c || inst >= it.b) return;      // items of a launch may differ in width / batch
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int mode = it.mode;
    const float inv_cnt = 1.f / ((float)it.b * (float)it.n);

    // ---- load + transform -> tile[ch][pt]
    // x has element strides (x_sb, x_sc, x_sn); u (backward modes) is a contiguous (b, c, n) tensor
    const float* xb = it.x + (size_t)inst * it.x_sb + (size_t)c0 * it.x_sc + (size_t)n0 * it.x_sn;
    const float* ub = it.u != nullptr ? it.u + ((size_t)inst * it.c + c0) * it.n + n0 : nullptr;
    const bool need_u = mode >= DCL_TR_BWD_RELU;
    if (VEC == 1) {
        float4 xv[4], uv[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {            // thread -> (channel t/32 + 8j, points 4*(t%32) .. +3)
            const int ch = (t >> 5) + 8 * j, pt = (t & 31) * 4;
            xv[j] = dcl_ld_stream_f4(xb + (size_t)ch * it.x_sc + pt);
            uv[j] = need_u ? dcl_ld_stream_f4(ub + (size_t)ch * it.n + pt) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int ch = (t >> 5) + 8 * j, pt = (t & 31) * 4;
            const TrChan p = tr_chan(it, c0 + ch, inv_cnt);
            float4 y;
            y.x = tr_apply(mode, xv[j].x, uv[j].x, p);
            y.y = tr_apply(mode, xv[j].y, uv[j].y, p);
            y.z = tr_apply(mode, xv[j].z, uv[j].z, p);
            y.w = tr_apply(mode, xv[j].w, uv[j].w, p);
            *reinterpret_cast<float4*>(tile + ch * TR_LD + pt) = y;
        }
    } else if (VEC == 2) {
        float4 xv[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {            // thread -> (point t/8 + 32j, channels 4*(t%8) .. +3); forward modes only
            const int pt = (t >> 3) + 32 * j, ch = (t & 7) * 4;
            xv[j] = dcl_ld_stream_f4(xb + (size_t)pt * it.x_sn + ch);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int pt = (t >> 3) + 32 * j, ch = (t & 7) * 4;
            const float v[4] = {xv[j].x, xv[j].y, xv[j].z, xv[j].w};
#pragma unroll
            for (int e = 0; e < 4; ++e)
                tile[(ch + e) * TR_LD + pt] = tr_apply(mode, v[e], 0.f, tr_chan(it, c0 + ch + e, inv_cnt));
        }
    } else {
#pragma unroll 4
        for (int j = 0; j < (TR_TC * TR_TP) / 256; ++j) {
            int ch, pt;
            const int idx = t + 256 * j;
            if (it.x_sc == 1) {       // point-major source: 32 consecutive channels per point
                ch = idx & 31;
                pt = idx >> 5;
            } else {                  // channel-major source: 128 consecutive points per channel
                ch = idx >> 7;
                pt = idx & 127;
            }
            const float x = __ldg(xb + (size_t)ch * it.x_sc + (size_t)pt * it.x_sn);
            const float u = need_u ? __ldg(ub + (size_t)ch * it.n + pt) : 0.f;
            tile[ch * TR_LD + pt] = tr_apply(mode, x, u, tr_chan(it, c0 + ch, inv_cnt));
        }
    }
    __syncthreads();

    // ---- fp32 channel-major copy of the result
    if (it.out_cm != nullptr) {
        float* o = it.out_cm + ((size_t)inst * it.c + c0) * it.n + n0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int ch = (t >> 5) + 8 * j, pt = (t & 31) * 4;
            *reinterpret_cast<float4*>(o + (size_t)ch * it.n + pt) = *reinterpret_cast<const float4*>(tile + ch * TR_LD + pt);
        }
    }
    // ---- PM image of the (b*n x c) matrix: this tile is exactly one blob
    if (it.out_k != nullptr) {
        const int kcols = it.k_cols > 0 ? it.k_cols : it.c;
        unsigned char* blob = reinterpret_cast<unsigned char*>(it.out_k) +
                              ((size_t)((size_t)inst * it.n + n0) / 128 * (kcols / 32) + it.k_col0 / 32 + blockIdx.y) * TR_BLOB;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int uu = t + 256 * h;                     // unit = (row r, 8-channel chunk q); byte offset uu*16
            const int r = (uu >> 5) * 8 + (uu & 7), q = (uu >> 3) & 3;
            float v[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = tile[(q * 8 + e) * TR_LD + r];
            uint4 lo;
            const uint4 hi = pack8_hi_lo(v, lo);
            *reinterpret_cast<uint4*>(blob + uu * 16) = hi;
            *reinterpret_cast<uint4*>(blob + TR_BLOB / 2 + uu * 16) = lo;
        }
    }
    // ---- transposed image of instance `inst`: PM image of the (t_rows x n) matrix [channel][point]
    if (it.out_t != nullptr) {
        const int row0 = it.t_row0 + c0;                    // first of this tile's 32 channel rows
        const int grp = it.t_group > 1 ? it.t_group : 1;
        const int kbs = grp * (it.n / 32);
        unsigned char* img = reinterpret_cast<unsigned char*>(it.out_t) + (size_t)(inst / grp) * it.t_rows * kbs * 128 +
                             ((size_t)(row0 / 128) * kbs + (inst % grp) * (it.n / 32) + n0 / 32) * TR_BLOB +
                             ((row0 % 128) / 8) * 512;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int uu = t + 256 * h;                     // unit = (k-block kbl, row group rgl, 8-point chunk q8, row e)
            const int kbl = uu >> 7, rgl = (uu >> 5) & 3, q8 = (uu >> 3) & 3, e = uu & 7;
            const float* src = tile + (rgl * 8 + e) * TR_LD + kbl * 32 + q8 * 8;
            const float4 v0 = *reinterpret_cast<const float4*>(src), v1 = *reinterpret_cast<const float4*>(src + 4);
            const float v[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
            uint4 lo;
            const uint4 hi = pack8_hi_lo(v, lo);
            unsigned char* d = img + (size_t)kbl * TR_BLOB + rgl * 512 + q8 * 128 + e * 16;
            *reinterpret_cast<uint4*>(d) = hi;
            *reinterpret_cast<uint4*>(d + TR_BLOB / 2) = lo;
        }
    }
    // ---- per-tile column sums (bias gradient partials), fixed order
    if (it.col_partial != nullptr) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int ch = warp * 4 + k;
            float s = (tile[ch * TR_LD + lane] + tile[ch * TR_LD + lane + 32]) +
                      (tile[ch * TR_LD + lane + 64] + tile[ch * TR_LD + lane + 96]);
#pragma unroll
            for (int o = 16; o >= 1; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (lane == 0) it.col_partial[(size_t)blockIdx.x * it.c + c0 + ch] = s;
        }
    }
}

// ------------------------------------------------------------------ block reduction of two fp64 sums (fixed order)
__device__ __forceinline__ void tr_block_reduce2(double& a, double& b, double* s_a, double* s_b) {
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) {
        s_a[warp] = a;
        s_b[warp] = b;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double ta = 0.0, tb = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) {
            ta += s_a[w];
            tb += s_b[w];
        }
        a = ta;
        b = tb;
    }
}

// Train-mode BatchNorm statistics of channel blockIdx.x of a contiguous (b, c, n) tensor.
__global__ void __launch_bounds__(256) tr_bn_stats_kernel(const __grid_constant__ TrBnBatch batch) {
    __shared__ double s_a[8], s_b[8];
    const dcl_tr_bn& it = batch.it[blockIdx.y];
    const int ch = blockIdx.x;
    if (ch >= it.c) return;
    double sum = 0.0, sq = 0.0;
    const int n4 = it.n / 4;
    for (int inst = 0; inst < it.b; ++inst) {
        const float4* row = reinterpret_cast<const float4*>(it.u + ((size_t)inst * it.c + ch) * it.n);
        for (int i = threadIdx.x; i < n4; i += 256) {
            const float4 v = __ldg(row + i);
            sum += (double)v.x + (double)v.y + (double)v.z + (double)v.w;
            sq += (double)v.x * v.x + (double)v.y * v.y + (double)v.z * v.z + (double)v.w * v.w;
        }
    }
    tr_block_reduce2(sum, sq, s_a, s_b);
    if (threadIdx.x == 0) {
        const double cnt = (double)it.b * it.n;
        const double mean = sum / cnt;
        double var = sq / cnt - mean * mean;
        var = var > 0.0 ? var : 0.0;
        const float rstd = (float)(1.0 / sqrt(var + (double)it.eps));
        const float g = it.gamma != nullptr ? it.gamma[ch] : 1.f, bt = it.beta != nullptr ? it.beta[ch] : 0.f;
        const float scale = g * rstd;
        it.mean[ch] = (float)mean;
        it.rstd[ch] = rstd;
        it.scale[ch] = scale;
        it.shift[ch] = bt - (float)mean * scale;
        if (it.running_mean != nullptr) {
            const float m = it.momentum;
            it.running_mean[ch] = (1.f - m) * it.running_mean[ch] + m * (float)mean;
            it.running_var[ch] = (1.f - m) * it.running_var[ch] + m * (float)(var * cnt / (cnt - 1.0));
        }
    }
}

// The two per-channel sums of the BatchNorm backward:  s1 = sum g (= d beta),  s2 = sum g * xhat (= d gamma), where
// g = dY masked by the ReLU that follows the BatchNorm (mode DCL_TR_BWD_BN_RELU) or dY itself (DCL_TR_BWD_RELU_BN).
__global__ void __launch_bounds__(256) tr_bn_bwd_reduce_kernel(const __grid_constant__ TrBnBwdBatch batch) {
    __shared__ double s_a[8], s_b[8];
    const dcl_tr_bn_bwd& it = batch.it[blockIdx.y];
    const int ch = blockIdx.x;
    if (ch >= it.c) return;
    const float mean = it.mean[ch], rstd = it.rstd[ch], sc = it.scale[ch], sh = it.shift[ch];
    const bool masked = it.mode == DCL_TR_BWD_BN_RELU;
    double s1 = 0.0, s2 = 0.0;
    const bool vec = it.n % 4 == 0 && it.dy_sb % 4 == 0 && it.dy_sc % 4 == 0 && ((((uintptr_t)it.dy) | ((uintptr_t)it.u)) & 15u) == 0;
    for (int inst = 0; inst < it.b; ++inst) {
        const float* urow = it.u + ((size_t)inst * it.c + ch) * it.n;
        const float* drow = it.dy + (size_t)inst * it.dy_sb + (size_t)ch * it.dy_sc;
        if (vec) {
            for (int i = threadIdx.x; i < it.n / 4; i += 256) {
                const float4 u4 = __ldg(reinterpret_cast<const float4*>(urow) + i);
                const float4 g4 = __ldg(reinterpret_cast<const float4*>(drow) + i);
                const float uu[4] = {u4.x, u4.y, u4.z, u4.w}, gg[4] = {g4.x, g4.y, g4.z, g4.w};
                float a1 = 0.f, a2 = 0.f;      // four terms in fp32, then into the fp64 running sums
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float g = (masked && !(__fmaf_rn(uu[e], sc, sh) > 0.f)) ? 0.f : gg[e];
                    a1 += g;
                    a2 = __fmaf_rn(g, (uu[e] - mean) * rstd, a2);
                }
                s1 += (double)a1;
                s2 += (double)a2;
            }
        } else {
            for (int i = threadIdx.x; i < it.n; i += 256) {
                const float u = __ldg(urow + i);
                float g = __ldg(drow + i);
                if (masked && !(__fmaf_rn(u, sc, sh) > 0.f)) g = 0.f;
                s1 += (double)g;
                s2 += (double)g * (double)((u - mean) * rstd);
            }
        }
    }
    tr_block_reduce2(s1, s2, s_a, s_b);
    if (threadIdx.x == 0) {
        it.s1[ch] = (float)s1;
        it.s2[ch] = (float)s2;
    }
}

// fp32 weights (rows x cols, row-major; transpose != 0: the source holds the matrix transposed, cols x rows) ->
// packed bf16 hi/lo blobs of the (rows_pad x k_pad) matrix with n-tile nt (include/dcl_b200.h: packed weights);
// elements outside (rows x cols) are zero.  Thread = one 16-byte unit (row o, 8 consecutive k).
__global__ void __launch_bounds__(256) tr_pack_weights_kernel(const __grid_constant__ TrWpackBatch batch) {
    const dcl_tr_wpack& it = batch.it[blockIdx.y];
    const int kchunks = it.k_pad / 8;
    const long g = (long)blockIdx.x * 256 + threadIdx.x;
    if (g >= (long)it.rows_pad * kchunks) return;
    const int o = (int)(g % it.rows_pad), kc = (int)(g / it.rows_pad);
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const int k = kc * 8 + e;
        float x = 0.f;
        if (o < it.rows && k < it.cols)
            x = it.transpose ? __ldg(it.src + (size_t)k * it.rows + o) : __ldg(it.src + (size_t)o * it.cols + k);
        v[e] = x;
    }
    uint4 lo;
    const uint4 hi = pack8_hi_lo(v, lo);
    const int nt = it.nt;
    const int k_total = it.k_total > 0 ? it.k_total : it.k_pad;
    unsigned char* d = reinterpret_cast<unsigned char*>(it.dst) +
                       ((size_t)(o / nt) * (k_total / 32) + it.k_col0 / 32 + kc / 4) * ((size_t)nt * 128) +
                       ((o % nt) / 8) * 512 + (kc & 3) * 128 + (o & 7) * 16;
    *reinterpret_cast<uint4*>(d) = hi;
    *reinterpret_cast<uint4*>(d + (size_t)nt * 64) = lo;
}

// out[c] = sum over `parts` rows of partial[parts][c], lane = channel (coalesced 128-byte rows), the rows dealt to the
// CTA's 8 warps round-robin and the 8 warp sums added in warp order: a fixed order, hence deterministic.
__global__ void __launch_bounds__(256) tr_colsum_kernel(const __grid_constant__ TrColsumBatch batch) {
    __shared__ float s_part[8][33];
    const dcl_tr_colsum& it = batch.it[blockIdx.y];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ch = blockIdx.x * 32 + lane;
    if (blockIdx.x * 32 >= it.c) return;
    float acc = 0.f;
    if (ch < it.c) {
#pragma unroll 4
        for (int p = warp; p < it.parts; p += 8) acc += __ldg(it.partial + (size_t)p * it.c + ch);
    }
    s_part[warp][lane] = acc;
    __syncthreads();
    if (warp == 0 && ch < it.c) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += s_part[w][lane];
        it.out[ch] = t;
    }
}

}  // namespace

DCL_API int dcl_tr_tile_pass(int nitems, const dcl_tr_tile* items, void* stream) {
    DCL_RETURN_IF_BAD(nitems >= 1 && nitems <= TR_MAX_ITEMS && items != nullptr);
    TrTileBatch batch;
    int max_tiles = 0, max_cb = 0;
    for (int i = 0; i < nitems; ++i) {
        const dcl_tr_tile& it = items[i];
        DCL_RETURN_IF_BAD(it.x != nullptr && it.b > 0 && it.c > 0 && it.c % TR_TC == 0 && it.n > 0 && it.n % TR_TP == 0);
        DCL_RETURN_IF_BAD(it.mode >= DCL_TR_COPY && it.mode <= DCL_TR_BWD_RELU_BN);
        const bool affine = it.mode != DCL_TR_COPY && it.mode != DCL_TR_BWD_RELU;
        DCL_RETURN_IF_BAD(!affine || (it.scale != nullptr && it.shift != nullptr));
        DCL_RETURN_IF_BAD(it.mode < DCL_TR_BWD_RELU || it.u != nullptr);
        DCL_RETURN_IF_BAD(it.mode < DCL_TR_BWD_BN_RELU ||
                          (it.mean != nullptr && it.rstd != nullptr && it.s1 != nullptr && it.s2 != nullptr));
        DCL_RETURN_IF_BAD(it.out_t == nullptr || (it.t_rows % 128 == 0 && it.t_row0 >= 0 && it.t_row0 % TR_TC == 0 &&
                                                  it.t_row0 + it.c <= it.t_rows));
        DCL_RETURN_IF_BAD(((((uintptr_t)it.out_k) | ((uintptr_t)it.out_t)) & 15u) == 0);
        DCL_RETURN_IF_BAD(it.out_cm == nullptr || (((uintptr_t)it.out_cm) & 15u) == 0);
        DCL_RETURN_IF_BAD(it.t_group <= 1 || it.b % it.t_group == 0);
        DCL_RETURN_IF_BAD(it.k_cols == 0 ? it.k_col0 == 0
                                         : (it.k_cols % 32 == 0 && it.k_col0 >= 0 && it.k_col0 % 32 == 0 &&
                                            it.k_col0 + it.c <= it.k_cols));
        batch.it[i] = it;
        const int tiles = it.b * (it.n / TR_TP);
        max_tiles = tiles > max_tiles ? tiles : max_tiles;
        max_cb = it.c / TR_TC > max_cb ? it.c / TR_TC : max_cb;
    }
    // vector path of the loads: all items channel-major with 16-byte aligned rows (1), all point-major forward
    // transforms with aligned rows (2), scalar otherwise
    bool v1 = true, v2 = true;
    for (int i = 0; i < nitems; ++i) {
        const dcl_tr_tile& it = items[i];
        const bool al = (((uintptr_t)it.x) & 15u) == 0 && it.x_sb % 4 == 0;
        v1 = v1 && al && it.x_sn == 1 && it.x_sc % 4 == 0 && (it.u == nullptr || (((uintptr_t)it.u) & 15u) == 0);
        v2 = v2 && al && it.x_sc == 1 && it.x_sn % 4 == 0 && it.mode < DCL_TR_BWD_RELU;
        v1 = v1 && (it.out_cm == nullptr || (((uintptr_t)it.out_cm) & 15u) == 0);
        v2 = v2 && (it.out_cm == nullptr || (((uintptr_t)it.out_cm) & 15u) == 0);
    }
    const dim3 grid(max_tiles, max_cb, nitems);
    cudaStream_t st = (cudaStream_t)stream;
    if (v1) tr_tile_kernel<1><<<grid, 256, 0, st>>>(batch);
    else if (v2) tr_tile_kernel<2><<<grid, 256, 0, st>>>(batch);
    else tr_tile_kernel<0><<<grid, 256, 0, st>>>(batch);
    return dcl_launch_status();
}

DCL_API int dcl_tr_bn_stats(int nitems, const dcl_tr_bn* items, void* stream) {
    DCL_RETURN_IF_BAD(nitems >= 1 && nitems <= TR_MAX_ITEMS && items != nullptr);
    TrBnBatch batch;
    int max_c = 0;
    for (int i = 0; i < nitems; ++i) {
        const dcl_tr_bn& it = items[i];
        DCL_RETURN_IF_BAD(it.u != nullptr && it.b > 0 && it.c > 0 && it.n > 0 && it.n % 4 == 0 && (long)it.b * it.n > 1);
        DCL_RETURN_IF_BAD(it.mean != nullptr && it.rstd != nullptr && it.scale != nullptr && it.shift != nullptr);
        DCL_RETURN_IF_BAD((it.running_mean == nullptr) == (it.running_var == nullptr));
        DCL_RETURN_IF_BAD((((uintptr_t)it.u) & 15u) == 0);
        batch.it[i] = it;
        max_c = it.c > max_c ? it.c : max_c;
    }
    tr_bn_stats_kernel<<<dim3(max_c, nitems), 256, 0, (cudaStream_t)stream>>>(batch);
    return dcl_launch_status();
}

DCL_API int dcl_tr_bn_bwd_reduce(int nitems, const dcl_tr_bn_bwd* items, void* stream) {
    DCL_RETURN_IF_BAD(nitems >= 1 && nitems <= TR_MAX_ITEMS && items != nullptr);
    TrBnBwdBatch batch;
    int max_c = 0;
    for (int i = 0; i < nitems; ++i) {
        const dcl_tr_bn_bwd& it = items[i];
        DCL_RETURN_IF_BAD(it.dy != nullptr && it.u != nullptr && it.b > 0 && it.c > 0 && it.n > 0);
        DCL_RETURN_IF_BAD(it.mode == DCL_TR_BWD_BN_RELU || it.mode == DCL_TR_BWD_RELU_BN);
        DCL_RETURN_IF_BAD(it.mean != nullptr && it.rstd != nullptr && it.scale != nullptr && it.shift != nullptr &&
                          it.s1 != nullptr && it.s2 != nullptr);
        batch.it[i] = it;
        max_c = it.c > max_c ? it.c : max_c;
    }
    tr_bn_bwd_reduce_kernel<<<dim3(max_c, nitems), 256, 0, (cudaStream_t)stream>>>(batch);
    return dcl_launch_status();
}

DCL_API int dcl_tr_pack_weights(int nitems, const dcl_tr_wpack* items, void* stream) {
    DCL_RETURN_IF_BAD(nitems >= 1 && nitems <= TR_MAX_ITEMS && items != nullptr);
    TrWpackBatch batch;
    long max_units = 0;
    for (int i = 0; i < nitems; ++i) {
        const dcl_tr_wpack& it = items[i];
        DCL_RETURN_IF_BAD(it.src != nullptr && it.dst != nullptr && it.rows > 0 && it.cols > 0);
        DCL_RETURN_IF_BAD((it.nt == 64 || it.nt == 128 || it.nt == 256) && it.rows_pad >= it.rows &&
                          it.rows_pad % it.nt == 0 && it.k_pad >= it.cols && it.k_pad % 32 == 0);
        DCL_RETURN_IF_BAD((((uintptr_t)it.dst) & 15u) == 0);
        DCL_RETURN_IF_BAD(it.k_total == 0 ? it.k_col0 == 0
                                          : (it.k_total % 32 == 0 && it.k_col0 >= 0 && it.k_col0 % 32 == 0 &&
                                             it.k_col0 + it.k_pad <= it.k_total));
        batch.it[i] = it;
        const long units = (long)it.rows_pad * (it.k_pad / 8);
        max_units = units > max_units ? units : max_units;
    }
    tr_pack_weights_kernel<<<dim3((unsigned)DCL_DIVUP(max_units, 256L), nitems), 256, 0, (cudaStream_t)stream>>>(batch);
    return dcl_launch_status();
}

DCL_API int dcl_tr_colsum_reduce(int nitems, const dcl_tr_colsum* items, void* stream) {
    DCL_RETURN_IF_BAD(nitems >= 1 && nitems <= TR_MAX_ITEMS && items != nullptr);
    TrColsumBatch batch;
    int max_c = 0;
    for (int i = 0; i < nitems; ++i) {
        const dcl_tr_colsum& it = items[i];
        DCL_RETURN_IF_BAD(it.partial != nullptr && it.out != nullptr && it.parts > 0 && it.c > 0);
        batch.it[i] = it;
        max_c = it.c > max_c ? it.c : max_c;
    }
    tr_colsum_kernel<<<dim3(DCL_DIVUP(max_c, 32), nitems), 256, 0, (cudaStream_t)stream>>>(batch);
    return dcl_launch_status();
}
