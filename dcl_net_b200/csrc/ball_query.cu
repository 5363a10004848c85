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

}  // namespace

DCL_API int dcl_lib_ball_query_kernel_launcher_fast(int b, int n, int m, float radius, int nsample,
                                                    const float* new_xyz, const float* xyz, int* idx,
                                                    void* stream) {
    DCL_RETURN_IF_BAD(b >= 0 && n >= 0 && m >= 0 && nsample >= 0);
    if (b == 0 || m == 0 || nsample == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
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
