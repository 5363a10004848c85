// Voxel sets, voxelisation and rulebooks of the two sparse-conv towers, built on the
// device with no host round trip.
//
// Reference: the dataloader hashes every point into a voxel on the host (libs/pointgroup_ops voxelize_idx,
// voxelize.cpp:57-163), Network.forward mean-pools the point features (voxelize.cu:10-23), and every SparseConv3d /
// SubMConv3d / SparseAvgPool3d of Backbone_SPCONV (models/Modules.py:100-159) builds its index pairs with a dense
// 64^3-per-instance int grid, atomics and a torch::_unique sort (libs/spconv indice.cu.h:24-220, spconv_ops.h:27-136).
//
// Here a voxel set of one instance is a BIT GRID: G*G rows of G bits (row = (x,y), bit = z), 32 KB at G = 64.  The
// operators the towers need are bit-parallel on it:
//     SparseConv3d k3 s1 p1   output set = 3x3x3 dilation              (OR of 9 rows, then r | r<<1 | r>>1)
//     SubMConv3d              output set = input set
//     SparseAvgPool3d k3 s2   output set = dilation sampled at even z   (OR of 9 rows, shifts, compress even bits)
// and the reference's row order (sorted by linear index, spconv_ops.h:122) is the rank of a bit: a per-row exclusive
// prefix of popcounts plus the popcount below the bit.  One CTA per instance builds all nine sets of a tower
// (spb_build_sets_kernel); rows are then packed densely over the batch (offset of an instance = sum of the counts
// before it) and the neighbour table ("rulebook") of every conv / pool is 27 rank queries per output row.
//
// Set s of a tower: 0 = occupied voxels (64^3), then per pyramid level l = 0..3: 2l+1 = conv-out set (dilated, grid
// 64 >> l), 2l+2 = pool-out set (grid 32 >> l).  Op j = 3l + {0,1,2}: conv (out 2l+1, in 2l), subm (out = in = 2l+1),
// pool (out 2l+2, in 2l+1).
#include "common.cuh"
#include "../../include/dcl_b200.h"
#include <cuda_fp16.h>

namespace {

constexpr int SPB_THREADS = 1024;
constexpr int SPB_NSETS = DCL_SPB_NSETS;   // 9
constexpr int SPB_NOPS = DCL_SPB_NOPS;     // 12
constexpr int SPB_MAX_PTS = 4096;          // points per instance handled by one CTA

__host__ __device__ constexpr int spb_grid(int s) { return 64 >> (s >> 1); }   // 64,64,32,32,16,16,8,8,4
// rows of set s start here inside an instance's row block
__host__ __device__ inline int spb_row_base(int s) {
    int base = 0;
    for (int i = 0; i < s; ++i) base += spb_grid(i) * spb_grid(i);
    return base;
}
constexpr int SPB_ROWS_PER_INST = 4096 * 2 + 1024 * 2 + 256 * 2 + 64 * 2 + 16;  // 10896

__device__ __forceinline__ unsigned long long spb_low_mask(int g) {
    return g >= 64 ? ~0ull : ((1ull << g) - 1ull);
}
__device__ __forceinline__ unsigned long long spb_compress_even(unsigned long long x) {
    x &= 0x5555555555555555ull;
    x = (x | (x >> 1)) & 0x3333333333333333ull;
    x = (x | (x >> 2)) & 0x0f0f0f0f0f0f0f0full;
    x = (x | (x >> 4)) & 0x00ff00ff00ff00ffull;
    x = (x | (x >> 8)) & 0x0000ffff0000ffffull;
    x = (x | (x >> 16)) & 0x00000000ffffffffull;
    return x;
}

// exclusive scan of popcounts of rows[0..n) (n <= 4096, a multiple of 16 or smaller than the block) by the whole
// block; returns the total.  prefix may be shared or global memory.
__device__ int spb_block_prefix(const unsigned long long* rows, int n, int* prefix, int* s_warp) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int per = (n + SPB_THREADS - 1) / SPB_THREADS;   // <= 4
    int local[4], sum = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = tid * per + i;
        local[i] = (i < per && r < n) ? __popcll(rows[r]) : 0;
        sum += local[i];
    }
    int inc = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += v;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        int w = s_warp[lane];
        int winc = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, winc, d);
            if (lane >= d) winc += v;
        }
        s_warp[lane] = winc - w;          // exclusive warp offsets
        if (lane == 31) s_warp[32] = winc;  // total
    }
    __syncthreads();
    int run = s_warp[warp] + inc - sum;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = tid * per + i;
        if (i < per && r < n) prefix[r] = run;
        run += local[i];
    }
    const int total = s_warp[32];
    __syncthreads();
    return total;
}

// in-place bitonic sort of n (power of two, <= 4096) unsigned keys in shared memory, ascending
__device__ void spb_bitonic(unsigned int* keys, int n) {
    for (int k = 2; k <= n; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < n; i += SPB_THREADS) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const unsigned int a = keys[i], b = keys[ixj];
                    const bool up = (i & k) == 0;
                    if ((a > b) == up) {
                        keys[i] = b;
                        keys[ixj] = a;
                    }
                }
            }
            __syncthreads();
        }
    }
}

struct SpbTowerDev {
    const float* points;     // (B*n_per, 3) or null
    const float* rgb;        // (B*n_per, 3) or null (zeros)
    const int* coords;       // (B*n_per, 4) bxyz voxel coordinates, used when points == null
    unsigned long long* rows;  // [B][SPB_ROWS_PER_INST]
    int* prefix;             // [B][SPB_ROWS_PER_INST]
    int* counts;             // [SPB_NSETS][B]
    // voxelisation outputs, slot layout: instance b owns rows [b*n_per, (b+1)*n_per)
    __half* feat16;          // (B*n_per, 16) conv-0 operand rows in SORTED voxel order: [1, rgb, xyz, lo(rgb), lo(xyz), 0..]
    float* feat32;           // (B*n_per, 7)  mean-voxelised features, FIRST-APPEARANCE order (reference layout)
    int* occupied;           // (B*n_per, 4)  voxel coordinates, first-appearance order
    int* p2v;                // (B*n_per)     point -> voxel (first-appearance id, local to the instance)
    int* v2p_sorted;         // (B*n_per)     point indices (local) grouped by voxel (first-appearance order), ascending
    int* v2p_start;          // (B*n_per)     start of each voxel's group in v2p_sorted (local); count = next start - start
    int* errors;             // [0] += points outside the grid
};
struct SpbBuildArgs {
    int B, n_per, ntowers;
    float half_extent, unit;
    SpbTowerDev tw[2];
};

// ------------------------------------------------------------------ kernel A: all voxel sets of one instance
__global__ void __launch_bounds__(SPB_THREADS, 1) spb_build_sets_kernel(const __grid_constant__ SpbBuildArgs args) {
    extern __shared__ __align__(16) unsigned char spb_smem[];
    unsigned long long* bufA = reinterpret_cast<unsigned long long*>(spb_smem);            // 4096 rows
    unsigned long long* bufB = bufA + 4096;                                                   // 4096 rows
    int* s_prefix = reinterpret_cast<int*>(bufB + 4096);                                      // 4096
    unsigned int* s_keys = reinterpret_cast<unsigned int*>(s_prefix + 4096);                  // SPB_MAX_PTS
    unsigned int* s_vox = s_keys + SPB_MAX_PTS;                                               // SPB_MAX_PTS: voxel linear id of point i
    int* s_first = reinterpret_cast<int*>(s_vox + SPB_MAX_PTS);                               // SPB_MAX_PTS: per voxel rank: start in sorted list
    int* s_fa = s_first + SPB_MAX_PTS;                                                        // SPB_MAX_PTS: rank -> first-appearance id
    __shared__ int s_warp[40];
    __shared__ int s_bad;

    const SpbTowerDev& tw = args.tw[blockIdx.y];
    const int b = blockIdx.x, n_per = args.n_per, tid = threadIdx.x;
    unsigned long long* g_rows = tw.rows + (size_t)b * SPB_ROWS_PER_INST;
    int* g_prefix = tw.prefix + (size_t)b * SPB_ROWS_PER_INST;

    for (int i = tid; i < 4096; i += SPB_THREADS) bufA[i] = 0ull;
    if (tid == 0) s_bad = 0;
    __syncthreads();
    // ---- S0: mark the voxel of every point.  (p + 0.5*extent) / unit in fp32, truncated, as the dataloader does
    // (YCBV/dataloader_test_YCBV.py:177); a point outside the 64^3 grid is skipped and counted.
    for (int i = tid; i < n_per; i += SPB_THREADS) {
        int ix, iy, iz;
        const size_t p = (size_t)b * n_per + i;
        if (tw.points != nullptr) {
            ix = (int)__fdiv_rn(__fadd_rn(tw.points[p * 3 + 0], args.half_extent), args.unit);
            iy = (int)__fdiv_rn(__fadd_rn(tw.points[p * 3 + 1], args.half_extent), args.unit);
            iz = (int)__fdiv_rn(__fadd_rn(tw.points[p * 3 + 2], args.half_extent), args.unit);
        } else {
            ix = tw.coords[p * 4 + 1];
            iy = tw.coords[p * 4 + 2];
            iz = tw.coords[p * 4 + 3];
        }
        const bool ok = ((unsigned)ix < 64u) && ((unsigned)iy < 64u) && ((unsigned)iz < 64u);
        s_vox[i] = ok ? (unsigned)((ix * 64 + iy) * 64 + iz) : 0xffffffffu;
        if (ok) atomicOr(&bufA[ix * 64 + iy], 1ull << iz);
        else atomicAdd(&s_bad, 1);
    }
    __syncthreads();
    int total = spb_block_prefix(bufA, 4096, s_prefix, s_warp);
    for (int i = tid; i < 4096; i += SPB_THREADS) {
        g_rows[i] = bufA[i];
        g_prefix[i] = s_prefix[i];
    }
    if (tid == 0) {
        tw.counts[0 * args.B + b] = total;
        if (s_bad > 0 && tw.errors != nullptr) atomicAdd(tw.errors, s_bad);
    }
    const int n_vox = total;

    // ---- voxelisation: group the points by voxel rank, keep point order inside a voxel (voxelize.cpp:96-107)
    int npow = 1;
    while (npow < n_per) npow <<= 1;
    for (int i = tid; i < npow; i += SPB_THREADS) {
        unsigned int key = 0xffffffffu;
        if (i < n_per && s_vox[i] != 0xffffffffu) {
            const unsigned int v = s_vox[i];
            const int row = v >> 6, bit = v & 63;
            const int rank = s_prefix[row] + __popcll(bufA[row] & ((1ull << bit) - 1ull));
            key = ((unsigned)rank << 12) | (unsigned)i;      // n_per <= 4096, rank < 4096
            s_vox[i] = (unsigned)rank;                        // from here on: voxel RANK of point i
        }
        s_keys[i] = key;
    }
    __syncthreads();
    spb_bitonic(s_keys, npow);
    // segment starts: s_first[rank] = position of the voxel's first (lowest-index) point in the sorted list
    for (int i = tid; i < npow; i += SPB_THREADS) {
        const unsigned int key = s_keys[i];
        if (key != 0xffffffffu && (i == 0 || (s_keys[i - 1] >> 12) != (key >> 12))) s_first[key >> 12] = i;
    }
    __syncthreads();
    // first-appearance numbering (voxelize.cpp:100-103): order the voxels by the index of their first point
    unsigned int* s_keys2 = reinterpret_cast<unsigned int*>(bufB);   // bufB is free until the dilation below
    int vpow = 1;
    while (vpow < n_vox) vpow <<= 1;
    for (int r = tid; r < vpow; r += SPB_THREADS)
        s_keys2[r] = r < n_vox ? (((s_keys[s_first[r]] & 0xfffu) << 12) | (unsigned)r) : 0xffffffffu;
    __syncthreads();
    spb_bitonic(s_keys2, vpow);
    for (int f = tid; f < n_vox; f += SPB_THREADS) s_fa[s_keys2[f] & 0xfffu] = f;
    __syncthreads();
    {
        const size_t slot = (size_t)b * n_per;
        // per voxel (thread = rank): mean of [1, rgb, xyz] over its points IN POINT ORDER, multiplier 1/n first
        // (voxelize.cu:15-21: out += multiplier * inp); coordinates of the first point; the CSR form of output_map
        for (int r = tid; r < n_vox; r += SPB_THREADS) {
            const int start = s_first[r];
            int cnt = 0;
            while (start + cnt < npow && (s_keys[start + cnt] >> 12) == (unsigned)r) ++cnt;
            const float mult = __fdiv_rn(1.0f, (float)cnt);
            float acc[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            for (int j = 0; j < cnt; ++j) {
                const size_t p = slot + (s_keys[start + j] & 0xfffu);
                float f[7];
                f[0] = 1.0f;
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    f[1 + c] = tw.rgb != nullptr ? tw.rgb[p * 3 + c] : 0.f;
                    f[4 + c] = tw.points != nullptr ? tw.points[p * 3 + c] : 0.f;
                }
#pragma unroll
                for (int c = 0; c < 7; ++c) acc[c] = __fadd_rn(acc[c], __fmul_rn(mult, f[c]));
            }
            const int fa = s_fa[r];
            if (tw.feat32 != nullptr) {
#pragma unroll
                for (int c = 0; c < 7; ++c) tw.feat32[(slot + fa) * 7 + c] = acc[c];
            }
            if (tw.feat16 != nullptr) {
                // conv-0 operand row (sorted order): value channels 0-6, fp16 remainders of rgb / xyz in 7-12
                __half h[16];
#pragma unroll
                for (int c = 0; c < 7; ++c) h[c] = __float2half_rn(acc[c]);
#pragma unroll
                for (int c = 1; c < 7; ++c) h[6 + c] = __float2half_rn(acc[c] - __half2float(h[c]));
                h[13] = h[14] = h[15] = __float2half_rn(0.f);
                uint4* dst = reinterpret_cast<uint4*>(tw.feat16 + (slot + r) * 16);
                dst[0] = *reinterpret_cast<const uint4*>(&h[0]);
                dst[1] = *reinterpret_cast<const uint4*>(&h[8]);
            }
        }
        // voxel coordinates in first-appearance order
        if (tw.occupied != nullptr) {
            for (int row = tid; row < 4096; row += SPB_THREADS) {
                unsigned long long bits = bufA[row];
                int r = s_prefix[row];
                while (bits) {
                    const int z = __ffsll((long long)bits) - 1;
                    bits &= bits - 1ull;
                    int* o = tw.occupied + (slot + s_fa[r]) * 4;
                    o[0] = b;
                    o[1] = row >> 6;
                    o[2] = row & 63;
                    o[3] = z;
                    ++r;
                }
            }
        }
        if (tw.p2v != nullptr)
            for (int i = tid; i < n_per; i += SPB_THREADS)
                tw.p2v[slot + i] = s_vox[i] == 0xffffffffu ? -1 : s_fa[s_vox[i]];
        // the point list grouped by voxel in first-appearance order: sort (fa, point) keys
        if (tw.v2p_sorted != nullptr) {
            __syncthreads();
            for (int i = tid; i < npow; i += SPB_THREADS) {
                const unsigned int key = s_keys[i];
                s_keys[i] = key == 0xffffffffu ? key : (((unsigned)s_fa[key >> 12] << 12) | (key & 0xfffu));
            }
            __syncthreads();
            spb_bitonic(s_keys, npow);
            for (int i = tid; i < n_per; i += SPB_THREADS) {
                const unsigned int key = s_keys[i];
                tw.v2p_sorted[slot + i] = key == 0xffffffffu ? -1 : (int)(key & 0xfffu);
                if (key != 0xffffffffu && (i == 0 || (s_keys[i - 1] >> 12) != (key >> 12)))
                    tw.v2p_start[slot + (key >> 12)] = i;
            }
        }
    }
    __syncthreads();

    // ---- the conv-out / pool-out chain.  cur = bit grid of the set just finished (G rows per side)
    unsigned long long* cur = bufA;
    unsigned long long* nxt = bufB;
    int G = 64;
    for (int level = 0; level < 4; ++level) {
        // conv-out set 2*level+1: 3x3x3 dilation of cur on the same grid
        for (int row = tid; row < G * G; row += SPB_THREADS) {
            const int x = row / G, y = row - x * G;
            unsigned long long t = 0ull;
            for (int dx = -1; dx <= 1; ++dx) {
                const int xx = x + dx;
                if ((unsigned)xx >= (unsigned)G) continue;
                for (int dy = -1; dy <= 1; ++dy) {
                    const int yy = y + dy;
                    if ((unsigned)yy >= (unsigned)G) continue;
                    t |= cur[xx * G + yy];
                }
            }
            nxt[row] = (t | (t << 1) | (t >> 1)) & spb_low_mask(G);
        }
        __syncthreads();
        {
            const int s = 2 * level + 1, base = spb_row_base(s);
            total = spb_block_prefix(nxt, G * G, s_prefix, s_warp);
            for (int i = tid; i < G * G; i += SPB_THREADS) {
                g_rows[base + i] = nxt[i];
                g_prefix[base + i] = s_prefix[i];
            }
            if (tid == 0) tw.counts[s * args.B + b] = total;
        }
        __syncthreads();
        // pool-out set 2*level+2: kernel 3, stride 2, padding 1 on the dilated set: out o <- in 2o-1 .. 2o+1
        const int H = G / 2;
        for (int row = tid; row < H * H; row += SPB_THREADS) {
            const int ox = row / H, oy = row - ox * H;
            unsigned long long t = 0ull;
            for (int dx = -1; dx <= 1; ++dx) {
                const int xx = 2 * ox + dx;
                if ((unsigned)xx >= (unsigned)G) continue;
                for (int dy = -1; dy <= 1; ++dy) {
                    const int yy = 2 * oy + dy;
                    if ((unsigned)yy >= (unsigned)G) continue;
                    t |= nxt[xx * G + yy];
                }
            }
            const unsigned long long u = t | (t << 1) | (t >> 1);   // bit j: any of j-1, j, j+1
            cur[row] = spb_compress_even(u) & spb_low_mask(H);       // bit oz <- bit 2*oz
        }
        __syncthreads();
        {
            const int s = 2 * level + 2, base = spb_row_base(s);
            total = spb_block_prefix(cur, H * H, s_prefix, s_warp);
            for (int i = tid; i < H * H; i += SPB_THREADS) {
                g_rows[base + i] = cur[i];
                g_prefix[base + i] = s_prefix[i];
            }
            if (tid == 0) tw.counts[s * args.B + b] = total;
        }
        __syncthreads();
        G = H;
    }
}

// ------------------------------------------------------------------ kernel B: offsets, coordinates of every row
struct SpbEmitTower {
    const unsigned long long* rows;
    const int* prefix;
    const int* counts;      // [SPB_NSETS][B]
    int* offsets;           // [SPB_NSETS][B + 1] exclusive scan over the batch (last = total rows)
    int* indices[SPB_NSETS];  // (cap_s, 4) int32 bxyz per set (null: not wanted); rows >= total get batch id = B
    int cap[SPB_NSETS];
    int* errors;            // [1] |= 1 << s when set s overflows its capacity
};
struct SpbEmitArgs {
    int B, ntowers;
    SpbEmitTower tw[2];
};

__global__ void __launch_bounds__(256) spb_emit_indices_kernel(const __grid_constant__ SpbEmitArgs args) {
    __shared__ int s_red[8];
    const SpbEmitTower& tw = args.tw[blockIdx.z];
    const int b = blockIdx.x, s = blockIdx.y, B = args.B, tid = threadIdx.x;
    // offset of this instance = sum of the counts of the instances before it
    int part = 0;
    for (int i = tid; i < b; i += 256) part += tw.counts[s * B + i];
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) part += __shfl_xor_sync(0xffffffffu, part, d);
    if ((tid & 31) == 0) s_red[tid >> 5] = part;
    __syncthreads();
    int off = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) off += s_red[w];
    const int cnt = tw.counts[s * B + b];
    const int cap = tw.cap[s];
    if (tid == 0) {
        tw.offsets[s * (B + 1) + b] = off;
        if (b == B - 1) {
            tw.offsets[s * (B + 1) + B] = off + cnt;
            if (off + cnt > cap && tw.errors != nullptr) atomicOr(tw.errors + 1, 1 << s);
        }
    }
    int* ind = tw.indices[s];
    if (ind == nullptr) return;
    const int G = spb_grid(s);
    const unsigned long long* rows = tw.rows + (size_t)b * SPB_ROWS_PER_INST + spb_row_base(s);
    const int* prefix = tw.prefix + (size_t)b * SPB_ROWS_PER_INST + spb_row_base(s);
    for (int row = tid; row < G * G; row += 256) {
        unsigned long long bits = rows[row];
        int r = off + prefix[row];
        const int x = row / G, y = row - x * G;
        while (bits) {
            const int z = __ffsll((long long)bits) - 1;
            bits &= bits - 1ull;
            if (r < cap) reinterpret_cast<int4*>(ind)[r] = make_int4(b, x, y, z);
            ++r;
        }
    }
    // the last instance also marks the unused tail of the buffer: batch id B = "no instance"
    if (b == B - 1)
        for (int r = off + cnt + tid; r < cap; r += 256) reinterpret_cast<int4*>(ind)[r] = make_int4(B, 0, 0, 0);
}

// ------------------------------------------------------------------ kernel C: rulebooks
// nbr[op][tile][32][128] (tile = 128 consecutive output rows): slot k in 0..26 = input row feeding the tile's row
// through kernel offset k = (k0*3+k1)*3+k2 (input position = out*stride - 1 + k, geometry.h:24-86), -1 when that
// voxel is not in the input set; slot 27 = number of valid entries (the pool's divisor, summaryRF.cu:27-41).  anymask[op][tile] = OR over the tile's 128
// rows of their 27-bit validity masks, so the convolution skips kernel offsets no row of a tile uses.
struct SpbRuleTower {
    const unsigned long long* rows;
    const int* prefix;
    const int* offsets;          // [SPB_NSETS][B+1]
    const int* indices[SPB_NSETS];
    int* nbr[SPB_NOPS];          // (cap_out, 32)
    unsigned int* anymask[SPB_NOPS];  // (cap_out / 128)
    int cap[SPB_NSETS];
    int in0_slot;                // op 0's input rows live in slots of this many rows per instance (voxeliser output)
};
struct SpbRuleArgs {
    int B, ntowers;
    int tile_end[2 * SPB_NOPS];   // flat grid: blocks [tile_end[j-1], tile_end[j]) are the tiles of (tower, op) = (j / 12, j % 12)
    SpbRuleTower tw[2];
};

__global__ void __launch_bounds__(256) spb_rulebook_kernel(const __grid_constant__ SpbRuleArgs args) {
    // block = one tile of 128 output rows of one op, in three phases of INDEPENDENT accesses (the first version walked
    // its rows one warp-iteration at a time through three dependent global loads and was latency-bound):
    //   1. the tile's 128 row coordinates -> shared memory;
    //   2. the 128 x 9 (row, dx, dy) bit rows of the input set and their prefixes -> shared memory (each serves
    //      the three dz of its column);
    //   3. the 128 x 27 rank queries from shared memory into the tile's table, which is written out TRANSPOSED —
    //      nbr[tile][k][row in tile] — so that the consumers read, per kernel offset, 128 consecutive entries;
    //      slot 27 = number of valid entries; the tile's validity mask is OR-ed in shared memory (no atomics).
    __shared__ int4 s_ind[128];
    __shared__ unsigned long long s_bits[128][9];
    __shared__ int s_pre[128][9];
    __shared__ int s_base[128];
    __shared__ unsigned int s_valid[128];
    __shared__ unsigned int s_any;
    // flat grid over the tiles of every (tower, op): a (tiles, ops, towers) grid sized by the largest set launched
    // four times as many blocks as there are tiles, most of which exited at once
    int j = 0;
    while (j < 2 * SPB_NOPS - 1 && (int)blockIdx.x >= args.tile_end[j]) ++j;
    const SpbRuleTower& tw = args.tw[j / SPB_NOPS];
    const int op = j % SPB_NOPS, B = args.B;
    const int level = op / 3, kind = op - 3 * level;              // 0 conv, 1 subm, 2 pool
    const int s_out = kind == 2 ? 2 * level + 2 : 2 * level + 1;
    const int s_in = kind == 0 ? 2 * level : 2 * level + 1;
    const int stride = kind == 2 ? 2 : 1;
    const int total = min(tw.offsets[s_out * (B + 1) + B], tw.cap[s_out]);
    const int tile = (int)blockIdx.x - (j == 0 ? 0 : args.tile_end[j - 1]);
    if (tile * 128 >= total) return;
    const int tid = threadIdx.x;
    const int Gin = spb_grid(s_in);
    const int in_base = spb_row_base(s_in);
    const bool slots = op == 0 && tw.in0_slot > 0;
    if (tid == 0) s_any = 0;
    if (tid < 128) {
        const int r = tile * 128 + tid;
        int4 o = make_int4(-1, 0, 0, 0);
        if (r < total) o = reinterpret_cast<const int4*>(tw.indices[s_out])[r];
        s_ind[tid] = o;
        s_valid[tid] = 0;
        s_base[tid] = o.x < 0 ? 0 : (slots ? o.x * tw.in0_slot : tw.offsets[s_in * (B + 1) + o.x]);
    }
    __syncthreads();
    for (int e = tid; e < 128 * 9; e += 256) {
        const int rt = e / 9, d = e - rt * 9;
        const int4 o = s_ind[rt];
        unsigned long long bits = 0ull;
        int pre = 0;
        if (o.x >= 0) {
            const int x = o.y * stride - 1 + d / 3, y = o.z * stride - 1 + d % 3;
            if ((unsigned)x < (unsigned)Gin && (unsigned)y < (unsigned)Gin) {
                const size_t rb = (size_t)o.x * SPB_ROWS_PER_INST + in_base + x * Gin + y;
                bits = tw.rows[rb];
                pre = tw.prefix[rb];
            }
        }
        s_bits[rt][d] = bits;
        s_pre[rt][d] = pre;
    }
    __syncthreads();
    int* nbr = tw.nbr[op] + (size_t)tile * 32 * 128;
    for (int e = tid; e < 27 * 128; e += 256) {
        const int k = e >> 7, rt = e & 127;                        // consecutive threads = consecutive rows: coalesced
        const int4 o = s_ind[rt];
        int val = -1;
        if (o.x >= 0) {
            const int z = o.w * stride - 1 + k % 3;
            const unsigned long long bits = s_bits[rt][k / 3];
            if ((unsigned)z < (unsigned)Gin && ((bits >> z) & 1ull))
                val = s_base[rt] + s_pre[rt][k / 3] + __popcll(bits & ((1ull << z) - 1ull));
        }
        nbr[e] = val;
        if (val >= 0) atomicOr(&s_valid[rt], 1u << k);
    }
    __syncthreads();
    if (tid < 128) {
        const unsigned int v = s_valid[tid];
        nbr[27 * 128 + tid] = __popc(v);
        if (v) atomicOr(&s_any, v);
    }
    __syncthreads();
    if (tid == 0) tw.anymask[op][tile] = s_any;
}

// ------------------------------------------------------------------ SparseAvgPool3d (k3, s2, p1, use_gs = False)
// out[o] = sum over the kernel offsets, ascending, of in[i] / rf[o]   (avgpool.cu:44: out = out + in / rf)
struct SpbPoolTower {
    const float* in;           // (rows_in, c) fp32
    const int* nbr;            // (cap_out / 128, 32, 128)
    const int* offsets_out;    // &offsets[s_out * (B+1)]  (element B = total)
    float* out32;              // (cap_out, c)
    __half* out16;             // (cap_out, c) the same rounded once to fp16 (operand rows of the next conv), or null
    int cap_out;
};
struct SpbPoolArgs {
    int B, c, ntowers;
    SpbPoolTower tw[2];
};

// x / d, correctly rounded, for a divisor that is a small positive integer and a finite x whose quotient is normal or
// zero: q0 = RN(x * rc) with rc = RN(1/d), one exact residual, one correction (Markstein).  IEEE division
// (__fdiv_rn) gives the same bits but takes its slow path for every zero numerator — half of all post-ReLU features.
__device__ __forceinline__ float spb_div(float x, float d, float rc) {
    const float q0 = __fmul_rn(x, rc);
    const float e = __fmaf_rn(-q0, d, x);
    return __fmaf_rn(e, rc, q0);
}

__global__ void __launch_bounds__(256) spb_avgpool_kernel(const __grid_constant__ SpbPoolArgs args) {
    const SpbPoolTower& tw = args.tw[blockIdx.y];
    const int c4 = args.c >> 2;
    const int total = min(tw.offsets_out[args.B], tw.cap_out);
    const long work = (long)total * c4;
    for (long g = (long)blockIdx.x * 256 + threadIdx.x; g < work; g += (long)gridDim.x * 256) {
        const int r = (int)(g / c4), q = (int)(g - (long)r * c4);
        const int* ncol = tw.nbr + (size_t)(r >> 7) * 32 * 128 + (r & 127);    // entry k of row r at ncol[k * 128]
        const float rf = (float)__ldg(ncol + 27 * 128);
        const float rc = __frcp_rn(rf);
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        // additions in ascending kernel-offset order (avgpool.cu:44 runs one launch per offset)
#pragma unroll 1
        for (int i = 0; i < 7; ++i) {
            int nb[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) nb[j] = 4 * i + j < 27 ? __ldg(ncol + (4 * i + j) * 128) : -1;
            float4 v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j)
                v[j] = (nb[j] >= 0 && 4 * i + j < 27)
                           ? __ldg(reinterpret_cast<const float4*>(tw.in + (size_t)nb[j] * args.c) + q)
                           : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (nb[j] < 0 || 4 * i + j >= 27) continue;
                acc.x = __fadd_rn(acc.x, spb_div(v[j].x, rf, rc));
                acc.y = __fadd_rn(acc.y, spb_div(v[j].y, rf, rc));
                acc.z = __fadd_rn(acc.z, spb_div(v[j].z, rf, rc));
                acc.w = __fadd_rn(acc.w, spb_div(v[j].w, rf, rc));
            }
        }
        reinterpret_cast<float4*>(tw.out32 + (size_t)r * args.c)[q] = acc;
        if (tw.out16 != nullptr) {
            const __half2 a = __floats2half2_rn(fminf(fmaxf(acc.x, -65504.f), 65504.f), fminf(fmaxf(acc.y, -65504.f), 65504.f));
            const __half2 bq = __floats2half2_rn(fminf(fmaxf(acc.z, -65504.f), 65504.f), fminf(fmaxf(acc.w, -65504.f), 65504.f));
            uint2 w;
            w.x = *reinterpret_cast<const uint32_t*>(&a);
            w.y = *reinterpret_cast<const uint32_t*>(&bq);
            reinterpret_cast<uint2*>(tw.out16 + (size_t)r * args.c)[q] = w;
        }
    }
}

// ------------------------------------------------------------------ voxelisation (mean) with an explicit rule matrix
// libs/pointgroup_ops voxelize.cu:10-23 (mode 4): out[m, :] = sum_i (1/n_m) * feats[rule[m, 1+i], :], in rule order.
__global__ void __launch_bounds__(256) voxelize_mean_kernel(int m, int width, int c, const float* __restrict__ feats,
                                                            const int* __restrict__ rules, float* __restrict__ out) {
    const long g = (long)blockIdx.x * 256 + threadIdx.x;
    if (g >= (long)m * c) return;
    const int row = (int)(g / c), ch = (int)(g - (long)row * c);
    const int* r = rules + (size_t)row * width;
    const int n = r[0];
    const float mult = n > 0 ? __fdiv_rn(1.0f, (float)n) : 1.0f;
    float acc = 0.f;
    for (int i = 1; i <= n; ++i) acc = __fadd_rn(acc, __fmul_rn(mult, __ldg(feats + (size_t)r[i] * c + ch)));
    out[g] = acc;
}

}  // namespace

// =================================================================================== C-ABI
DCL_API size_t dcl_spb_rows_per_instance(void) { return SPB_ROWS_PER_INST; }

DCL_API int dcl_spb_build_sets(int b, int n_per, int ntowers, const dcl_spb_tower_in* towers, float unit, int grid,
                               void* stream) {
    DCL_RETURN_IF_BAD(b > 0 && n_per > 0 && n_per <= SPB_MAX_PTS && ntowers >= 1 && ntowers <= 2 && towers != nullptr);
    DCL_RETURN_IF_BAD(grid == 64 && unit > 0.f);
    SpbBuildArgs args;
    args.B = b;
    args.n_per = n_per;
    args.ntowers = ntowers;
    args.unit = unit;
    // total extent * 0.5 exactly as the dataloader forms it: float(unit * limit) * 0.5 (dataloader_test_YCBV.py:177)
    args.half_extent = (float)((double)unit * grid) * 0.5f;
    for (int t = 0; t < ntowers; ++t) {
        const dcl_spb_tower_in& in = towers[t];
        DCL_RETURN_IF_BAD((in.points != nullptr || in.coords != nullptr) && in.rows != nullptr && in.prefix != nullptr &&
                          in.counts != nullptr);
        SpbTowerDev& d = args.tw[t];
        d.points = in.points;
        d.rgb = in.rgb;
        d.coords = in.coords;
        d.rows = reinterpret_cast<unsigned long long*>(in.rows);
        d.prefix = in.prefix;
        d.counts = in.counts;
        d.feat16 = reinterpret_cast<__half*>(in.feat16);
        d.feat32 = in.feat32;
        d.occupied = in.occupied;
        d.p2v = in.p2v;
        d.v2p_sorted = in.v2p_sorted;
        d.v2p_start = in.v2p_start;
        d.errors = in.errors;
        DCL_RETURN_IF_BAD(in.v2p_sorted == nullptr || in.v2p_start != nullptr);
    }
    const size_t smem = 2 * 4096 * 8 + 4096 * 4 + 4 * SPB_MAX_PTS * 4;
    cudaError_t e = cudaFuncSetAttribute(spb_build_sets_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    spb_build_sets_kernel<<<dim3(b, ntowers), SPB_THREADS, smem, (cudaStream_t)stream>>>(args);
    return dcl_launch_status();
}

DCL_API int dcl_spb_emit(int b, int ntowers, const dcl_spb_tower_sets* towers, int in0_slot, void* stream) {
    DCL_RETURN_IF_BAD(b > 0 && ntowers >= 1 && ntowers <= 2 && towers != nullptr);
    cudaStream_t st = (cudaStream_t)stream;
    SpbEmitArgs ea;
    SpbRuleArgs ra;
    ea.B = ra.B = b;
    ea.ntowers = ra.ntowers = ntowers;
    int max_cap = 0;
    for (int t = 0; t < ntowers; ++t) {
        const dcl_spb_tower_sets& in = towers[t];
        DCL_RETURN_IF_BAD(in.rows != nullptr && in.prefix != nullptr && in.counts != nullptr && in.offsets != nullptr);
        SpbEmitTower& e = ea.tw[t];
        SpbRuleTower& r = ra.tw[t];
        e.rows = r.rows = reinterpret_cast<const unsigned long long*>(in.rows);
        e.prefix = r.prefix = in.prefix;
        e.counts = in.counts;
        e.offsets = in.offsets;
        r.offsets = in.offsets;
        e.errors = in.errors;
        r.in0_slot = in0_slot;
        for (int s = 0; s < SPB_NSETS; ++s) {
            e.indices[s] = in.indices[s];
            r.indices[s] = in.indices[s];
            e.cap[s] = r.cap[s] = in.cap[s];
            DCL_RETURN_IF_BAD(in.cap[s] >= 0 && (in.cap[s] % 128 == 0));
        }
        for (int op = 0; op < SPB_NOPS; ++op) {
            r.nbr[op] = in.nbr[op];
            r.anymask[op] = in.anymask[op];
            const int level = op / 3, kind = op % 3;
            const int s_out = kind == 2 ? 2 * level + 2 : 2 * level + 1;
            DCL_RETURN_IF_BAD(in.nbr[op] != nullptr && in.anymask[op] != nullptr && in.indices[s_out] != nullptr);
            if (in.cap[s_out] > max_cap) max_cap = in.cap[s_out];
        }
    }
    spb_emit_indices_kernel<<<dim3(b, SPB_NSETS, ntowers), 256, 0, st>>>(ea);
    int blocks = 0;
    for (int j = 0; j < 2 * SPB_NOPS; ++j) {
        const int t = j / SPB_NOPS, op = j % SPB_NOPS;
        const int level = op / 3, kind = op % 3;
        if (t < ntowers) blocks += towers[t].cap[kind == 2 ? 2 * level + 2 : 2 * level + 1] / 128;
        ra.tile_end[j] = blocks;
    }
    (void)max_cap;
    if (blocks > 0) spb_rulebook_kernel<<<blocks, 256, 0, st>>>(ra);
    return dcl_launch_status(2);
}

DCL_API int dcl_spb_avgpool(int b, int c, int ntowers, const dcl_spb_pool* pools, void* stream) {
    DCL_RETURN_IF_BAD(b > 0 && c > 0 && c % 4 == 0 && ntowers >= 1 && ntowers <= 2 && pools != nullptr);
    SpbPoolArgs args;
    args.B = b;
    args.c = c;
    args.ntowers = ntowers;
    int max_cap = 0;
    for (int t = 0; t < ntowers; ++t) {
        const dcl_spb_pool& p = pools[t];
        DCL_RETURN_IF_BAD(p.in != nullptr && p.nbr != nullptr && p.offsets_out != nullptr && p.out32 != nullptr &&
                          p.cap_out > 0);
        args.tw[t] = {p.in, p.nbr, p.offsets_out, p.out32, reinterpret_cast<__half*>(p.out16), p.cap_out};
        if (p.cap_out > max_cap) max_cap = p.cap_out;
    }
    long work = (long)max_cap * (c / 4);
    int blocks = (int)((work + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    spb_avgpool_kernel<<<dim3(blocks, ntowers), 256, 0, (cudaStream_t)stream>>>(args);
    return dcl_launch_status();
}

DCL_API int dcl_voxelize_mean(int m, int width, int c, const float* feats, const int* rules, float* out, void* stream) {
    DCL_RETURN_IF_BAD(m >= 0 && width >= 2 && c > 0 && feats != nullptr && rules != nullptr && out != nullptr);
    if (m == 0) return 0;
    const long total = (long)m * c;
    voxelize_mean_kernel<<<(unsigned)DCL_DIVUP(total, 256L), 256, 0, (cudaStream_t)stream>>>(m, width, c, feats, rules, out);
    return dcl_launch_status();
}
