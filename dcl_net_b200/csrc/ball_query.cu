// Ball query.  Semantics: libs/pointnet_lib/src/ball_query_gpu.cu:9-45 —
// per centre, scan k = 0..n-1 ascending, collect the first `nsample` indices with
// d2 < radius*radius (strict, fp32); on the first hit fill every slot with it; rows
// without a hit are left as the caller provided them (zeros).
//
// The reference has every warp re-read xyz[k] from global for each of its 32
// centres and writes idx with a stride of nsample.  Here the candidate tiles are
// streamed once per CTA through shared memory (TMA bulk copies, double-buffered),
// a CTA stops as soon as all of its centres are full, and the index rows are
// assembled in shared memory and written back coalesced.
#include "common.cuh"
#include "tile_pipe.cuh"
#include "../../include/dcl_b200.h"
#include <stdlib.h>

namespace {

constexpr int BQ_THREADS = 128;
constexpr int BQ_TILE_PTS = 1024;
constexpr int BQ_TILE_FLOATS = BQ_TILE_PTS * 3;

template <bool STAGE_OUT>
__global__ void __launch_bounds__(BQ_THREADS) ball_query_kernel(int n, int m, float radius, int nsample,
                                                                const float* __restrict__ new_xyz,
                                                                const float* __restrict__ xyz,
                                                                int* __restrict__ idx) {
    __shared__ __align__(16) float s_tile[2 * BQ_TILE_FLOATS];
    __shared__ uint64_t s_bar[2];
    extern __shared__ __align__(16) int s_out[];  // STAGE_OUT: BQ_THREADS * nsample
    const int bs = blockIdx.y;
    const int pt = blockIdx.x * BQ_THREADS + threadIdx.x;
    const bool valid = pt < m;
    new_xyz += (size_t)bs * m * 3;
    xyz += (size_t)bs * n * 3;
    idx += (size_t)bs * m * nsample;

    const float radius2 = __fmul_rn(radius, radius);
    const int pc = valid ? pt : (m - 1);
    const float cx = new_xyz[pc * 3 + 0], cy = new_xyz[pc * 3 + 1], cz = new_xyz[pc * 3 + 2];
    // Row slot l of this thread lives at s_out[l * BQ_THREADS + tid] (conflict-free) or in global.
    int* my_row = STAGE_OUT ? (s_out + threadIdx.x) : (idx + (size_t)pc * nsample);
    const int row_stride = STAGE_OUT ? BQ_THREADS : 1;

    int cnt = 0;
    bool done = !valid || nsample <= 0;
    DclTilePipe<BQ_TILE_FLOATS> pipe;
    pipe.init(s_tile, s_bar, xyz, n * 3);
    for (int t = 0; t < pipe.ntiles; ++t) {
        const int tc = pipe.acquire(t) / 3;
        const float* tile = pipe.tile(t);
        const int kbase = t * BQ_TILE_PTS;
        // No `break` inside the scan: with independent thread scheduling a divergent exit keeps the lanes of a
        // warp apart for the rest of the loop (measured: ~8x the instructions).  Lanes that are full just stop
        // recording; the warp leaves a chunk early only when all of its lanes are done (a uniform branch).
        for (int j0 = 0; j0 < tc; j0 += 32) {
            if (__all_sync(0xffffffffu, done)) break;
            const int jend = min(tc, j0 + 32);
            for (int j = j0; j < jend; ++j) {
                const float d2 = dcl_dist2(cx, cy, cz, tile[j * 3 + 0], tile[j * 3 + 1], tile[j * 3 + 2]);
                if (!done && d2 < radius2) {
                    const int k = kbase + j;
                    if (cnt == 0) {
                        for (int l = 0; l < nsample; ++l) my_row[l * row_stride] = k;
                    }
                    my_row[cnt * row_stride] = k;
                    ++cnt;
                    done = cnt >= nsample;
                }
                __syncwarp();
            }
        }
        const int all_done = __syncthreads_and(done ? 1 : 0);
        if (all_done) {
            pipe.drain_n(t + 1, 1);
            break;
        }
        pipe.release_nosync(t);
    }
    if (STAGE_OUT) {
        __syncthreads();
        // Thread r owns row r; a row is written only if it had a hit (cnt > 0).  Publish the
        // hit flags, then let consecutive threads write consecutive words.
        __shared__ int s_hit[BQ_THREADS];
        s_hit[threadIdx.x] = (valid && cnt > 0) ? 1 : 0;
        __syncthreads();
        const int row0 = blockIdx.x * BQ_THREADS;
        const int rows = min(BQ_THREADS, m - row0);
        const int total = rows * nsample;
        int* dst = idx + (size_t)row0 * nsample;
        for (int w = threadIdx.x; w < total; w += BQ_THREADS) {
            const int r = w / nsample, l = w - r * nsample;
            if (s_hit[r]) dst[w] = s_out[l * BQ_THREADS + r];
        }
    }
}


// ------------------------------------------------------------------ warp-cooperative form
// One warp scans for BQW_Q centres at once, 32 candidates per step (lane = candidate): a candidate's coordinates are
// read once (three conflict-free shared-memory loads, stride 3 words) and tested against the warp's BQW_Q centres; hits
// are rare (the expected ball holds a handful of the n points), so the common step is loads + 7 instructions per centre
// + one vote.  A hit step places the indices in scan order: slot = hits so far + popc(ballot below this lane), which is
// exactly the order of the reference's ascending scan; the first hit of a centre is remembered and, as in the
// reference, fills the slots the scan never reaches.  Against the thread-per-centre kernel above: 32x the threads
// (b*m warps instead of b*m threads — that kernel ran at ~10 % occupancy at the microbench shape and was bound by the
// latency of its dependent shared-memory loads), and no per-candidate __syncwarp.
constexpr int BQW_Q = 4;            // centres per warp
constexpr int BQW_WARPS = 8;        // warps per CTA
constexpr int BQW_THREADS = BQW_WARPS * 32;

__global__ void __launch_bounds__(BQW_THREADS) ball_query_warp_kernel(int n, int m, float radius, int nsample,
                                                                      const float* __restrict__ new_xyz,
                                                                      const float* __restrict__ xyz,
                                                                      int* __restrict__ idx) {
    __shared__ __align__(16) float s_tile[2 * BQ_TILE_FLOATS];
    __shared__ uint64_t s_bar[2];
    const int bs = blockIdx.y;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int q0 = (blockIdx.x * BQW_WARPS + warp) * BQW_Q;     // first centre of this warp
    new_xyz += (size_t)bs * m * 3;
    xyz += (size_t)bs * n * 3;
    idx += (size_t)bs * m * nsample;
    const float radius2 = __fmul_rn(radius, radius);
    float cx[BQW_Q], cy[BQW_Q], cz[BQW_Q];
    int cnt[BQW_Q], first[BQW_Q];
#pragma unroll
    for (int q = 0; q < BQW_Q; ++q) {
        const int pc = min(q0 + q, m - 1);
        cx[q] = __ldg(new_xyz + pc * 3 + 0);
        cy[q] = __ldg(new_xyz + pc * 3 + 1);
        cz[q] = __ldg(new_xyz + pc * 3 + 2);
        cnt[q] = (q0 + q < m) ? 0 : nsample;          // centres past the end count as full
        first[q] = 0;
    }
    const uint32_t lt_mask = (1u << lane) - 1u;
    DclTilePipe<BQ_TILE_FLOATS> pipe;
    pipe.init(s_tile, s_bar, xyz, n * 3);
    bool warp_done = false;
    for (int t = 0; t < pipe.ntiles; ++t) {
        const int tc = pipe.acquire(t) / 3;
        const float* tile = pipe.tile(t);
        const int kbase = t * BQ_TILE_PTS;
        if (!warp_done) {
            for (int j0 = 0; j0 < tc; j0 += 32) {
                const int j = j0 + lane;
                const bool in = j < tc;
                const int jc = in ? j : tc - 1;
                const float x = tile[jc * 3 + 0], y = tile[jc * 3 + 1], z = tile[jc * 3 + 2];
                bool any_open = false;
#pragma unroll
                for (int q = 0; q < BQW_Q; ++q) {
                    const bool hit = in && dcl_dist2(cx[q], cy[q], cz[q], x, y, z) < radius2;
                    const uint32_t mask = __ballot_sync(0xffffffffu, hit);
                    if (mask != 0u && cnt[q] < nsample) {
                        if (cnt[q] == 0) first[q] = kbase + j0 + (__ffs(mask) - 1);
                        const int pos = cnt[q] + __popc(mask & lt_mask);
                        if (hit && pos < nsample) idx[(size_t)(q0 + q) * nsample + pos] = kbase + j;
                        cnt[q] = min(nsample, cnt[q] + __popc(mask));
                    }
                    any_open = any_open || cnt[q] < nsample;
                }
                if (!any_open) {
                    warp_done = true;
                    break;
                }
            }
        }
        // a CTA leaves the scan when all of its warps are full
        const int all_done = __syncthreads_and(warp_done ? 1 : 0);
        if (all_done) {
            pipe.drain_n(t + 1, 1);
            break;
        }
        pipe.release_nosync(t);
    }
    // slots the scan did not reach repeat the first hit (the reference fills the whole row with it at the first hit)
#pragma unroll
    for (int q = 0; q < BQW_Q; ++q) {
        if (q0 + q < m && cnt[q] > 0)
            for (int l = cnt[q] + lane; l < nsample; l += 32) idx[(size_t)(q0 + q) * nsample + l] = first[q];
    }
}

}  // namespace

DCL_API int dcl_lib_ball_query_kernel_launcher_fast(int b, int n, int m, float radius, int nsample,
                                                    const float* new_xyz, const float* xyz, int* idx,
                                                    void* stream) {
    DCL_RETURN_IF_BAD(b >= 0 && n >= 0 && m >= 0 && nsample >= 0);
    if (b == 0 || m == 0 || nsample == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    // warp-cooperative kernel by default; DCL_BALL_QUERY_THREAD=1 keeps the thread-per-centre kernel (A/B runs)
    static const bool force_thread = getenv("DCL_BALL_QUERY_THREAD") != nullptr;
    if (!force_thread) {
        dim3 wgrid(DCL_DIVUP(m, BQW_WARPS * BQW_Q), b);
        ball_query_warp_kernel<<<wgrid, BQW_THREADS, 0, st>>>(n, m, radius, nsample, new_xyz, xyz, idx);
        return dcl_launch_status();
    }
    dim3 grid(DCL_DIVUP(m, BQ_THREADS), b);
    const size_t smem = (size_t)BQ_THREADS * nsample * sizeof(int);
    if (smem <= 64 * 1024) {
        if (smem > 20 * 1024)
            cudaFuncSetAttribute(ball_query_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        ball_query_kernel<true><<<grid, BQ_THREADS, smem, st>>>(n, m, radius, nsample, new_xyz, xyz, idx);
    } else {
        ball_query_kernel<false><<<grid, BQ_THREADS, 0, st>>>(n, m, radius, nsample, new_xyz, xyz, idx);
    }
    return dcl_launch_status();
}
