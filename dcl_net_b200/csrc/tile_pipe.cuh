// Two-stage shared-memory ring that streams one contiguous float array through
// a CTA.  Fast path: 1-D TMA bulk copies (cp.async.bulk -> UBLKCP) completing on
// mbarriers, issued by thread 0 one tile ahead of the consumers.  Fallback when
// the source is not 16-B aligned / sized: cooperative ld.global -> st.shared.
//
// Usage (all threads of the CTA, uniformly):
//     pipe.init(buf, bars, src, total_floats);      // includes a __syncthreads
//     for (t = 0; t < pipe.ntiles; ++t) {
//         int cnt = pipe.acquire(t);  const float* tile = pipe.tile(t);
//         ... read tile[0..cnt) ...
//         pipe.release(t);                           // includes a __syncthreads
//     }
// A CTA that wants to stop early calls pipe.drain(t_next) (uniformly) first so
// that no bulk copy is still landing in its shared memory when it exits.
#pragma once
#include "common.cuh"

template <int TILE_FLOATS>
struct DclTilePipe {
    static_assert(TILE_FLOATS % 4 == 0, "tile must be a multiple of 16 bytes");
    float* buf;
    uint64_t* bars;
    const float* src;
    int total;
    int ntiles;
    bool tma;

    __device__ __forceinline__ int count(int t) const {
        const int rem = total - t * TILE_FLOATS;
        return rem < TILE_FLOATS ? rem : TILE_FLOATS;
    }
    __device__ __forceinline__ const float* tile(int t) const { return buf + (t & 1) * TILE_FLOATS; }

    __device__ __forceinline__ void issue(int t) {
        const int s = t & 1;
        const uint32_t bytes = (uint32_t)count(t) * 4u;
        dcl_mbar_arrive_expect_tx(&bars[s], bytes);
        dcl_bulk_g2s(buf + s * TILE_FLOATS, src + (size_t)t * TILE_FLOATS, bytes, &bars[s]);
    }

    __device__ __forceinline__ void init(float* smem_buf, uint64_t* smem_bars, const float* gsrc, int total_floats) {
        buf = smem_buf;
        bars = smem_bars;
        src = gsrc;
        total = total_floats;
        ntiles = DCL_DIVUP(total_floats, TILE_FLOATS);
        tma = ((((uintptr_t)gsrc) & 15u) == 0) && ((total_floats & 3) == 0);
        if (tma) {
            if (threadIdx.x == 0) {
                dcl_mbar_init(&bars[0], 1);
                dcl_mbar_init(&bars[1], 1);
                dcl_fence_barrier_init();
            }
            __syncthreads();
            if (threadIdx.x == 0) {
                if (ntiles > 0) issue(0);
                if (ntiles > 1) issue(1);
            }
        }
    }

    __device__ __forceinline__ int acquire(int t) {
        const int cnt = count(t);
        if (tma) {
            dcl_mbar_wait(&bars[t & 1], (uint32_t)((t >> 1) & 1));
        } else {
            float* dst = buf + (t & 1) * TILE_FLOATS;
            const float* s = src + (size_t)t * TILE_FLOATS;
            for (int i = threadIdx.x; i < cnt; i += blockDim.x) dst[i] = __ldg(s + i);
            __syncthreads();
        }
        return cnt;
    }

    __device__ __forceinline__ void release(int t) {
        __syncthreads();
        release_nosync(t);
    }
    // For callers that already executed a CTA-wide barrier after their last read of tile t.
    __device__ __forceinline__ void release_nosync(int t) {
        if (tma && threadIdx.x == 0 && t + 2 < ntiles) issue(t + 2);
    }

    // Wait for every bulk copy that has been issued but not yet consumed, given
    // that tiles [0, t_next) were acquired AND released.
    __device__ __forceinline__ void drain(int t_next) { drain_n(t_next, 2); }
    // Same, when only `n_outstanding` tiles starting at t_next have been issued.
    __device__ __forceinline__ void drain_n(int t_next, int n_outstanding) {
        if (tma && threadIdx.x == 0) {
            for (int t = t_next; t < ntiles && t < t_next + n_outstanding; ++t)
                dcl_mbar_wait(&bars[t & 1], (uint32_t)((t >> 1) & 1));
        }
        __syncthreads();
    }
};
