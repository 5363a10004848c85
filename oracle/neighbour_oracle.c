/*
 * TEST INFRASTRUCTURE — NOT PART OF THE PRODUCT.
 *
 * CPU restatement (plain C, single thread) of the reference's CUDA-only point-cloud
 * neighbourhood operators, used only as the checker by tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline leg.  Nothing under dcl-net_b200/ may call into it.
 *
 * Every function cites the reference kernel it follows.  Floating-point expressions use
 * fmaf() in the exact contraction order nvcc (-O2, default -fmad=true, sm_100a) gives
 * the reference sources — checked in the SASS of oracle/_ref (see oracle/README.md):
 *     d2  = fmaf(dz,dz, fmaf(dx,dx, dy*dy))
 *     out = fmaf(w2,f2, fmaf(w0,f0, w1*f1))
 *
 * Parity status: pinned against the reference's own kernels (oracle/_ref, built from
 * /root/reference unmodified) on the GPU box by tests/test_gpu_neighbour_ops.py, and
 * against the fixtures those kernels produced (tests/golden/neighbour_ref_*.npz).
 *
 * Build: gcc -O2 -ffp-contract=off -shared -fPIC (oracle/build_oracle.py).
 * -ffp-contract=off matters: the contraction order is spelled out by hand.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORACLE_API __attribute__((visibility("default")))

static inline float dist2f(float ax, float ay, float az, float bx, float by, float bz) {
    const float dx = ax - bx, dy = ay - by, dz = az - bz;
    return fmaf(dz, dz, fmaf(dx, dx, dy * dy));
}

/* libs/pointnet_lib/src/cuda_utils.h:10-14 */
ORACLE_API int oracle_opt_n_threads(int work_size) {
    const int pow_2 = (int)(log((double)work_size) / log(2.0));
    int t = 1 << pow_2;
    if (t > 1024) t = 1024;
    if (t < 1) t = 1;
    return t;
}

/* libs/pointnet_lib/src/sampling_gpu.cu:93-209.  Simulates the block literally: T
 * threads, per-thread strided scan with strict '>', then the shared-memory tree of
 * __update (:86-91) with strides T/2 .. 1 (left slot kept unless right is greater). */
ORACLE_API void oracle_furthest_point_sampling(int b, int n, int m, const float* dataset, float* temp, int* idxs) {
    if (m <= 0) return;
    const int T = oracle_opt_n_threads(n);
    float* dists = (float*)malloc(sizeof(float) * (size_t)T);
    int* dists_i = (int*)malloc(sizeof(int) * (size_t)T);
    for (int bi = 0; bi < b; ++bi) {
        const float* xyz = dataset + (size_t)bi * n * 3;
        float* tmp = temp + (size_t)bi * n;
        int* out = idxs + (size_t)bi * m;
        int old = 0;
        out[0] = old;
        for (int j = 1; j < m; ++j) {
            const float x1 = xyz[old * 3 + 0], y1 = xyz[old * 3 + 1], z1 = xyz[old * 3 + 2];
            for (int tid = 0; tid < T; ++tid) {
                int besti = 0;
                float best = -1.f;
                for (int k = tid; k < n; k += T) {
                    const float d = dist2f(xyz[k * 3 + 0], xyz[k * 3 + 1], xyz[k * 3 + 2], x1, y1, z1);
                    const float d2 = fminf(d, tmp[k]);
                    tmp[k] = d2;
                    besti = d2 > best ? k : besti;
                    best = d2 > best ? d2 : best;
                }
                dists[tid] = best;
                dists_i[tid] = besti;
            }
            for (int s = T / 2; s >= 1; s >>= 1) {
                for (int tid = 0; tid < s; ++tid) {
                    const float v1 = dists[tid], v2 = dists[tid + s];
                    const int i1 = dists_i[tid], i2 = dists_i[tid + s];
                    dists[tid] = fmaxf(v1, v2);
                    dists_i[tid] = v2 > v1 ? i2 : i1;
                }
            }
            old = dists_i[0];
            out[j] = old;
        }
    }
    free(dists);
    free(dists_i);
}

/* sampling_gpu.cu:8-24 */
ORACLE_API void oracle_gather_points(int b, int c, int n, int m, const float* points, const int* idx, float* out) {
    for (int bi = 0; bi < b; ++bi)
        for (int ci = 0; ci < c; ++ci)
            for (int j = 0; j < m; ++j)
                out[((size_t)bi * c + ci) * m + j] = points[((size_t)bi * c + ci) * n + idx[(size_t)bi * m + j]];
}

/* sampling_gpu.cu:46-63 (atomicAdd order is unspecified there; here ascending j) */
ORACLE_API void oracle_gather_points_grad(int b, int c, int n, int m, const float* grad_out, const int* idx,
                                          float* grad_points) {
    for (int bi = 0; bi < b; ++bi)
        for (int ci = 0; ci < c; ++ci)
            for (int j = 0; j < m; ++j)
                grad_points[((size_t)bi * c + ci) * n + idx[(size_t)bi * m + j]] +=
                    grad_out[((size_t)bi * c + ci) * m + j];
}

/* libs/pointnet_lib/src/ball_query_gpu.cu:9-45 */
ORACLE_API void oracle_ball_query(int b, int n, int m, float radius, int nsample, const float* new_xyz,
                                  const float* xyz, int* idx) {
    const float radius2 = radius * radius;
    for (int bi = 0; bi < b; ++bi) {
        for (int pt = 0; pt < m; ++pt) {
            const float* c = new_xyz + ((size_t)bi * m + pt) * 3;
            const float* p = xyz + (size_t)bi * n * 3;
            int* row = idx + ((size_t)bi * m + pt) * nsample;
            int cnt = 0;
            for (int k = 0; k < n; ++k) {
                const float d2 = dist2f(c[0], c[1], c[2], p[k * 3 + 0], p[k * 3 + 1], p[k * 3 + 2]);
                if (d2 < radius2) {
                    if (cnt == 0)
                        for (int l = 0; l < nsample; ++l) row[l] = k;
                    row[cnt] = k;
                    ++cnt;
                    if (cnt >= nsample) break;
                }
            }
        }
    }
}

/* libs/pointnet_lib/src/group_points_gpu.cu:47-66 */
ORACLE_API void oracle_group_points(int b, int c, int n, int npoints, int nsample, const float* points,
                                    const int* idx, float* out) {
    const size_t E = (size_t)npoints * nsample;
    for (int bi = 0; bi < b; ++bi)
        for (int ci = 0; ci < c; ++ci)
            for (size_t e = 0; e < E; ++e)
                out[((size_t)bi * c + ci) * E + e] = points[((size_t)bi * c + ci) * n + idx[(size_t)bi * E + e]];
}

/* group_points_gpu.cu:8-25 */
ORACLE_API void oracle_group_points_grad(int b, int c, int n, int npoints, int nsample, const float* grad_out,
                                         const int* idx, float* grad_points) {
    const size_t E = (size_t)npoints * nsample;
    for (int bi = 0; bi < b; ++bi)
        for (int ci = 0; ci < c; ++ci)
            for (size_t e = 0; e < E; ++e)
                grad_points[((size_t)bi * c + ci) * n + idx[(size_t)bi * E + e]] +=
                    grad_out[((size_t)bi * c + ci) * E + e];
}

/* libs/pointnet_lib/src/interpolate_gpu.cu:81-124.  best* are double initialised to
 * 1e40 as in the reference; the stores narrow to float (1e40 -> +inf). */
ORACLE_API void oracle_three_nn(int b, int n, int m, const float* unknown, const float* known, float* dist2,
                                int* idx) {
    for (int bi = 0; bi < b; ++bi) {
        const float* kn = known + (size_t)bi * m * 3;
        for (int pt = 0; pt < n; ++pt) {
            const float* u = unknown + ((size_t)bi * n + pt) * 3;
            double best1 = 1e40, best2 = 1e40, best3 = 1e40;
            int besti1 = 0, besti2 = 0, besti3 = 0;
            for (int k = 0; k < m; ++k) {
                const float d = dist2f(u[0], u[1], u[2], kn[k * 3 + 0], kn[k * 3 + 1], kn[k * 3 + 2]);
                if (d < best1) {
                    best3 = best2; besti3 = besti2;
                    best2 = best1; besti2 = besti1;
                    best1 = d; besti1 = k;
                } else if (d < best2) {
                    best3 = best2; besti3 = besti2;
                    best2 = d; besti2 = k;
                } else if (d < best3) {
                    best3 = d; besti3 = k;
                }
            }
            float* dd = dist2 + ((size_t)bi * n + pt) * 3;
            int* ii = idx + ((size_t)bi * n + pt) * 3;
            dd[0] = (float)best1; dd[1] = (float)best2; dd[2] = (float)best3;
            ii[0] = besti1; ii[1] = besti2; ii[2] = besti3;
        }
    }
}

/* interpolate_gpu.cu:9-57 */
ORACLE_API void oracle_knn(int b, int n, int m, int k, const float* unknown, const float* known, float* dist2,
                           int* idx) {
    double best[200];
    int besti[200];
    for (int bi = 0; bi < b; ++bi) {
        const float* kn = known + (size_t)bi * m * 3;
        for (int pt = 0; pt < n; ++pt) {
            const float* u = unknown + ((size_t)bi * n + pt) * 3;
            for (int i = 0; i < k; ++i) {
                best[i] = 1e40;
                besti[i] = 0;
            }
            for (int i = 0; i < m; ++i) {
                const float d = dist2f(u[0], u[1], u[2], kn[i * 3 + 0], kn[i * 3 + 1], kn[i * 3 + 2]);
                for (int j = 0; j < k; ++j) {
                    if (d < best[j]) {
                        for (int l = k - 1; l > j; --l) {
                            best[l] = best[l - 1];
                            besti[l] = besti[l - 1];
                        }
                        best[j] = d;
                        besti[j] = i;
                        break;
                    }
                }
            }
            for (int i = 0; i < k; ++i) {
                idx[((size_t)bi * n + pt) * k + i] = besti[i];
                dist2[((size_t)bi * n + pt) * k + i] = (float)best[i];
            }
        }
    }
}

/* interpolate_gpu.cu:149-169 */
ORACLE_API void oracle_three_interpolate(int b, int c, int m, int n, const float* points, const int* idx,
                                         const float* weight, float* out) {
    for (int bi = 0; bi < b; ++bi)
        for (int ci = 0; ci < c; ++ci) {
            const float* row = points + ((size_t)bi * c + ci) * m;
            for (int i = 0; i < n; ++i) {
                const int* ii = idx + ((size_t)bi * n + i) * 3;
                const float* w = weight + ((size_t)bi * n + i) * 3;
                out[((size_t)bi * c + ci) * n + i] = fmaf(w[2], row[ii[2]], fmaf(w[0], row[ii[0]], w[1] * row[ii[1]]));
            }
        }
}

/* interpolate_gpu.cu:192-214 */
ORACLE_API void oracle_three_interpolate_grad(int b, int c, int n, int m, const float* grad_out, const int* idx,
                                              const float* weight, float* grad_points) {
    for (int bi = 0; bi < b; ++bi)
        for (int ci = 0; ci < c; ++ci) {
            float* row = grad_points + ((size_t)bi * c + ci) * m;
            for (int i = 0; i < n; ++i) {
                const int* ii = idx + ((size_t)bi * n + i) * 3;
                const float* w = weight + ((size_t)bi * n + i) * 3;
                const float g = grad_out[((size_t)bi * c + ci) * n + i];
                row[ii[0]] += g * w[0];
                row[ii[1]] += g * w[1];
                row[ii[2]] += g * w[2];
            }
        }
}

/* libs/pointnet_sp/src/interpolate_gpu.cu:9-56 */
ORACLE_API void oracle_sp_three_nn(int n, int m, const float* unknown, const float* known, float* dist2, int* idx) {
    for (int pt = 0; pt < n; ++pt) {
        const float* u = unknown + (size_t)pt * 4;
        double best1 = 1e40, best2 = 1e40, best3 = 1e40;
        int besti1 = 0, besti2 = 0, besti3 = 0;
        for (int k = 0; k < m; ++k) {
            const float* kn = known + (size_t)k * 4;
            if (kn[0] != u[0]) continue;
            const float d = dist2f(u[1], u[2], u[3], kn[1], kn[2], kn[3]);
            if (d < best1) {
                best3 = best2; besti3 = besti2;
                best2 = best1; besti2 = besti1;
                best1 = d; besti1 = k;
            } else if (d < best2) {
                best3 = best2; besti3 = besti2;
                best2 = d; besti2 = k;
            } else if (d < best3) {
                best3 = d; besti3 = k;
            }
        }
        dist2[pt * 3 + 0] = (float)best1; dist2[pt * 3 + 1] = (float)best2; dist2[pt * 3 + 2] = (float)best3;
        idx[pt * 3 + 0] = besti1; idx[pt * 3 + 1] = besti2; idx[pt * 3 + 2] = besti3;
    }
}

/* Model (not a restatement of the reference) of the product's slab-walk search (csrc/sp_interpolate.cu,
 * sp_group_search_t, slab mode): the known points are voxel centres ((i * ext) + off) + 0.5 * ext formed in fp32; an
 * instance's voxels are grouped into slabs of equal first index; a query visits slabs outwards from its own and stops
 * when the squared distance to the nearest unvisited slab plane exceeds 1.000001 x its current third-best; candidates
 * are ranked by the key (d, original index).  tests/test_cpu_oracle.py checks on the CPU that this returns exactly what
 * the reference-order scan oracle_sp_three_nn returns, i.e. that the pruning rule and the key are sound, ties included.
 * vox: (m,4) int32 (b, ix, iy, iz); unknown: (n,4) float bxyz. */
static int lex_lt(float d, int k, float bd, int bk) { return d < bd || (d == bd && k < bk); }
static void insert_lex(float d, int k, float* bd, int* bk) {
    if (!lex_lt(d, k, bd[2], bk[2])) return;
    if (lex_lt(d, k, bd[1], bk[1])) {
        bd[2] = bd[1]; bk[2] = bk[1];
        if (lex_lt(d, k, bd[0], bk[0])) { bd[1] = bd[0]; bk[1] = bk[0]; bd[0] = d; bk[0] = k; }
        else { bd[1] = d; bk[1] = k; }
    } else { bd[2] = d; bk[2] = k; }
}
ORACLE_API void oracle_sp_three_nn_slab_model(int n, int m, int gx, const float* unknown, const int* vox,
                                              const float* ext, const float* off, float* dist2, int* idx,
                                              long* visited_out) {
    long visited = 0;
    const float half0 = 0.5f * ext[0], half1 = 0.5f * ext[1], half2 = 0.5f * ext[2];
    for (int pt = 0; pt < n; ++pt) {
        const float* u = unknown + (size_t)pt * 4;
        float bd[3] = {INFINITY, INFINITY, INFINITY};
        int bk[3] = {0x7fffffff, 0x7fffffff, 0x7fffffff};
        int lo = (int)floorf((u[1] - off[0]) / ext[0]);
        if (lo < 0) lo = 0;
        if (lo > gx - 1) lo = gx - 1;
        int hi = lo + 1;
        while (lo >= 0 || hi < gx) {
            float dl = INFINITY, dh = INFINITY;
            if (lo >= 0) { const float cx = (((float)lo * ext[0]) + off[0]) + half0; const float dx = u[1] - cx; dl = dx * dx; }
            if (hi < gx) { const float cx = (((float)hi * ext[0]) + off[0]) + half0; const float dx = u[1] - cx; dh = dx * dx; }
            const int take_lo = hi >= gx || (lo >= 0 && dl <= dh);
            const float dmin = take_lo ? dl : dh;
            if (!(dmin <= bd[2] * 1.000001f)) break;
            const int s = take_lo ? lo-- : hi++;
            for (int k = 0; k < m; ++k) {   /* the slab's members, in any order */
                const int* v = vox + (size_t)k * 4;
                if ((float)v[0] != u[0] || v[1] != s) continue;
                const float cx = (((float)v[1] * ext[0]) + off[0]) + half0;
                const float cy = (((float)v[2] * ext[1]) + off[1]) + half1;
                const float cz = (((float)v[3] * ext[2]) + off[2]) + half2;
                const float d = dist2f(u[1], u[2], u[3], cx, cy, cz);
                ++visited;
                if (d < INFINITY) insert_lex(d, k, bd, bk);
            }
        }
        for (int j = 0; j < 3; ++j) {
            dist2[pt * 3 + j] = bd[j];
            idx[pt * 3 + j] = bk[j] == 0x7fffffff ? 0 : bk[j];
        }
    }
    if (visited_out) *visited_out = visited;
}

/* libs/pointnet_sp/src/interpolate_gpu.cu:80-102 */
ORACLE_API void oracle_sp_three_interpolate(int c, int m, int n, const float* points, const int* idx,
                                            const float* weight, float* out) {
    (void)m;
    for (int i = 0; i < n; ++i) {
        const int* ii = idx + (size_t)i * 3;
        const float* w = weight + (size_t)i * 3;
        for (int ci = 0; ci < c; ++ci)
            out[(size_t)i * c + ci] =
                fmaf(w[2], points[(size_t)ii[2] * c + ci],
                     fmaf(w[0], points[(size_t)ii[0] * c + ci], w[1] * points[(size_t)ii[1] * c + ci]));
    }
}

/* libs/pointnet_sp/src/interpolate_gpu.cu:124-146 */
ORACLE_API void oracle_sp_three_interpolate_grad(int c, int n, int m, const float* grad_out, const int* idx,
                                                 const float* weight, float* grad_points) {
    (void)m;
    for (int i = 0; i < n; ++i) {
        const int* ii = idx + (size_t)i * 3;
        const float* w = weight + (size_t)i * 3;
        for (int ci = 0; ci < c; ++ci) {
            const float g = grad_out[(size_t)i * c + ci];
            grad_points[(size_t)ii[0] * c + ci] += g * w[0];
            grad_points[(size_t)ii[1] * c + ci] += g * w[1];
            grad_points[(size_t)ii[2] * c + ci] += g * w[2];
        }
    }
}
