/*
 * dcl_b200.h — C-ABI of the B200-native DCL-Net hot path (libdcl_b200.so).
 *
 * Every entry point is `extern "C"`, takes plain device pointers + sizes + a
 * `cudaStream_t` (passed as void*), never allocates or frees, keeps no state,
 * launches asynchronously on the given stream and returns a `cudaError_t`
 * as int (0 == cudaSuccess).  The reference instead prints and calls
 * exit(-1) on a launch failure (libs/pointnet_lib/src/sampling_gpu.cu:248-252).
 *
 * Group 1 mirrors, one-to-one, the launchers the reference's pybind wrappers
 * call (same argument order and meaning; only the name gets a dcl_lib_ /
 * dcl_sp_ prefix because the two reference libraries define same-named
 * symbols).  Group 2 has no reference counterpart at the C level: it replaces
 * PyTorch library calls the reference makes from models/*.py.
 *
 * All tensors are dense, row-major, fp32 / int32, resident on the current
 * device.
 */
#ifndef DCL_B200_H
#define DCL_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ABI version of this header; bumped on any signature change. */
#define DCL_B200_ABI_VERSION 5
int dcl_b200_abi_version(void);
/* Compiled-for architecture as an integer (100 for sm_100a). */
int dcl_b200_arch(void);
/* Diagnostic: number of kernels this library has launched so far in the process. */
unsigned long long dcl_b200_launch_count(void);

/* ------------------------------------------------------------------------- */
/* Group 1a: libs/pointnet_lib (batched (B,N,3) clouds)                        */
/* ------------------------------------------------------------------------- */

/* replaces furthest_point_sampling_kernel_launcher
 * (libs/pointnet_lib/src/sampling_gpu.h:26-27, sampling_gpu.cu:211-253).
 * dataset (b,n,3); temp (b,n) pre-filled by the caller (reference: 1e10),
 * left holding the final min-distances; idxs (b,m) int32, idxs[:,0] = 0.
 * Tie-break identical to the reference's shared-memory tree reduction. */
int dcl_lib_furthest_point_sampling_kernel_launcher(int b, int n, int m,
    const float* dataset, float* temp, int* idxs, void* stream);

/* replaces gather_points_kernel_launcher_fast (sampling_gpu.h:12-13).
 * points (b,c,n), idx (b,npoints) -> out (b,c,npoints). */
int dcl_lib_gather_points_kernel_launcher_fast(int b, int c, int n, int npoints,
    const float* points, const int* idx, float* out, void* stream);

/* replaces gather_points_grad_kernel_launcher_fast (sampling_gpu.h:19-20).
 * grad_out (b,c,npoints), idx (b,npoints); accumulates into grad_points (b,c,n)
 * (caller zeroes it, libs/pointnet_lib/pointnet2_utils.py:70). */
int dcl_lib_gather_points_grad_kernel_launcher_fast(int b, int c, int n, int npoints,
    const float* grad_out, const int* idx, float* grad_points, void* stream);

/* replaces ball_query_kernel_launcher_fast (definition ball_query_gpu.cu:48-49;
 * pointer order is (new_xyz, xyz) as in the definition, not the header).
 * new_xyz (b,m,3), xyz (b,n,3) -> idx (b,m,nsample); rows without a hit are
 * left untouched (caller zeroes idx, pointnet2_utils.py:261). */
int dcl_lib_ball_query_kernel_launcher_fast(int b, int n, int m, float radius, int nsample,
    const float* new_xyz, const float* xyz, int* idx, void* stream);

/* replaces group_points_kernel_launcher_fast (group_points_gpu.h:13-14).
 * points (b,c,n), idx (b,npoints,nsample) -> out (b,c,npoints,nsample). */
int dcl_lib_group_points_kernel_launcher_fast(int b, int c, int n, int npoints, int nsample,
    const float* points, const int* idx, float* out, void* stream);

/* replaces group_points_grad_kernel_launcher_fast (group_points_gpu.h:19-20). */
int dcl_lib_group_points_grad_kernel_launcher_fast(int b, int c, int n, int npoints, int nsample,
    const float* grad_out, const int* idx, float* grad_points, void* stream);

/* replaces three_nn_kernel_launcher_fast (libs/pointnet_lib/src/interpolate_gpu.h:16-17).
 * unknown (b,n,3), known (b,m,3) -> dist2 (b,n,3) squared, idx (b,n,3). */
int dcl_lib_three_nn_kernel_launcher_fast(int b, int n, int m,
    const float* unknown, const float* known, float* dist2, int* idx, void* stream);

/* replaces knn_kernel_launcher_fast (interpolate_gpu.h:22-23); 1 <= k <= 200. */
int dcl_lib_knn_kernel_launcher_fast(int b, int n, int m, int k,
    const float* unknown, const float* known, float* dist2, int* idx, void* stream);

/* replaces three_interpolate_kernel_launcher_fast (interpolate_gpu.h:28-29).
 * points (b,c,m), idx/weight (b,n,3) -> out (b,c,n). */
int dcl_lib_three_interpolate_kernel_launcher_fast(int b, int c, int m, int n,
    const float* points, const int* idx, const float* weight, float* out, void* stream);

/* replaces three_interpolate_grad_kernel_launcher_fast (interpolate_gpu.h:33-34).
 * grad_out (b,c,n) -> accumulates into grad_points (b,c,m). */
int dcl_lib_three_interpolate_grad_kernel_launcher_fast(int b, int c, int n, int m,
    const float* grad_out, const int* idx, const float* weight, float* grad_points, void* stream);

/* ------------------------------------------------------------------------- */
/* Group 1b: libs/pointnet_sp (flat, batch id in column 0)                     */
/* ------------------------------------------------------------------------- */

/* replaces three_nn_kernel_launcher_fast (libs/pointnet_sp/src/interpolate_gpu.h:16-17).
 * unknown (n,4) bxyz, known (m,4) bxyz -> dist2 (n,3), idx (n,3) into `known`.
 * Same O(n*m) scan as the reference, tiled through shared memory. */
int dcl_sp_three_nn_kernel_launcher_fast(int n, int m,
    const float* unknown, const float* known, float* dist2, int* idx, void* stream);

/* Segmented variant of the same op: known rows are bucketed by batch id into
 * caller-provided scratch, each query scans only its own bucket.  Results are
 * bit-identical to dcl_sp_three_nn_kernel_launcher_fast.  `workspace` must hold
 * dcl_sp_three_nn_workspace_bytes(n,m) bytes.  Batch ids that are not integers
 * in [0, DCL_SP_MAX_BATCH) make the call fall back (on device, no host sync) to
 * the full scan. */
#define DCL_SP_MAX_BATCH 65536
size_t dcl_sp_three_nn_workspace_bytes(int n, int m);
int dcl_sp_three_nn_segmented(int n, int m,
    const float* unknown, const float* known, float* dist2, int* idx,
    void* workspace, size_t workspace_bytes, void* stream);

/* replaces three_interpolate_kernel_launcher_fast (pointnet_sp interpolate_gpu.h:22-23).
 * points (m,c), idx/weight (n,3) -> out (n,c). */
int dcl_sp_three_interpolate_kernel_launcher_fast(int c, int m, int n,
    const float* points, const int* idx, const float* weight, float* out, void* stream);

/* replaces three_interpolate_grad_kernel_launcher_fast (pointnet_sp interpolate_gpu.h:27-28).
 * grad_out (n,c) -> accumulates into grad_points (m,c). */
int dcl_sp_three_interpolate_grad_kernel_launcher_fast(int c, int n, int m,
    const float* grad_out, const int* idx, const float* weight, float* grad_points, void* stream);

/* Fusion of models/Modules.py:213-226 (Ops_nearest_neighbor_interpolate):
 * three_nn -> sqrt -> w = (1/(d+1e-8))/sum -> three_interpolate, writing into a
 * column slice of a wider (n, out_stride) row-major output (the torch.cat of
 * models/Modules.py:250).  Uses the segmented search; same workspace. */
int dcl_sp_nn_interpolate_fused(int n, int m, int c,
    const float* unknown, const float* known, const float* feats,
    float* out, int out_stride, int out_col0,
    void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------- */
/* Group 2: FDA correspondence head and pose solve (replace torch library calls) */
/* ------------------------------------------------------------------------- */

/* Fused Aligner (models/Modules.py:166-169) + confidence product
 * (models/DCL_Net.py:213,215), one direction:
 *   A        = softmax over the m axis of  RI_2^T RI_1      (b, m, n), never stored
 *   RE_embed = RE_2 A                                       (b, p, n)
 *   RI_embed = RI_2 A                                       (b, c, n)
 * RI_1 (b,c,n) queries, RI_2 (b,c,m) keys, RE_2 (b,p,m) values; fp32, channel-major
 * exactly as the reference holds them.  c in {64,128}, p == 256, n and m multiples
 * of 128 / 64.  The contraction runs on tcgen05 with bf16 hi/lo-split operands and
 * fp32 TMEM accumulation.  `workspace` holds the packed operands:
 * dcl_fda_workspace_bytes(b,c,p,n,m).  lse_out (b,n) optional (may be NULL):
 * per-query log-sum-exp, enough to rebuild A. */
size_t dcl_fda_workspace_bytes(int b, int c, int p, int n, int m);
int dcl_fda_align_fwd(int b, int c, int p, int n, int m,
    const float* RI_1, const float* RI_2, const float* RE_2,
    float* RE_embed, float* RI_embed, float* lse_out,
    void* workspace, size_t workspace_bytes, void* stream);

/* The two halves of dcl_fda_align_fwd, for callers that want to time or reuse them:
 * dcl_fda_pack converts the fp32 operands into the bf16 hi/lo tile images in `workspace`;
 * dcl_fda_fwd_packed runs the fused tcgen05 kernel on a packed workspace. */
int dcl_fda_pack(int b, int c, int p, int n, int m,
    const float* RI_1, const float* RI_2, const float* RE_2,
    void* workspace, size_t workspace_bytes, void* stream);
int dcl_fda_fwd_packed(int b, int c, int p, int n, int m,
    float* RE_embed, float* RI_embed, float* lse_out,
    void* workspace, size_t workspace_bytes, void* stream);
/* Up to two independent problems of equal shape in ONE launch (the two directions of the dual FDA,
 * models/DCL_Net.py:206-215, then share their partial last waves).  Fields as in dcl_fda_fwd_packed_pm;
 * every job has its own workspace of workspace_bytes. */
typedef struct dcl_fda_job {
    void* workspace;
    float* RE_embed;
    float* RI_embed;
    void* RE_pm;
    void* RI_pm;
    float* lse;
} dcl_fda_job;
int dcl_fda_fwd_packed_jobs(int njobs, const dcl_fda_job* jobs, int b, int c, int p, int n, int m,
    size_t workspace_bytes, void* stream);
/* The same with a choice of format for the P V products (pv_fmt):
 *   0  P and the values as bf16 hi/lo pairs, 3 MMAs per product (what every entry point above runs);
 *   1  P = exp2(..) and the values rounded once to fp16, ONE MMA per product; needs (n/128) even.  The value image
 *      then holds one fp16 image per 16-key chunk (chunk stride (256+c)*32 bytes instead of twice that; same
 *      element map), and RE_pm / RI_pm are written as PM16 images.  The logits keep split operands either way.
 * dcl_fda_pack_fmt is dcl_fda_pack writing the value image in that format. */
int dcl_fda_fwd_packed_jobs_fmt(int njobs, const dcl_fda_job* jobs, int b, int c, int p, int n, int m,
    size_t workspace_bytes, int pv_fmt, void* stream);
int dcl_fda_pack_fmt(int b, int c, int p, int n, int m,
    const float* RI_1, const float* RI_2, const float* RE_2,
    void* workspace, size_t workspace_bytes, int pv_fmt, void* stream);
/* Byte offsets of the query, key and value operand images inside the workspace (offsets[3]), for producers
 * that write them directly (dcl_pm_gemm_problem.out_qk / out_v) instead of calling dcl_fda_pack:
 *   query image: per 128 queries  [hi: 128 x c | lo], element (r,ch) at (r/8)*(c/8)*128 + (ch/8)*128 + (r%8)*16 + (ch%8)*2
 *   key image:   per 64 keys      [hi:  64 x c | lo], same element map
 *   value image: per 16 keys      [hi: (256+c) value rows x 16 keys | lo], element (vrow,key) at
 *                (vrow/8)*256 + ((key%16)/8)*128 + (key%8)*16 + (vrow%8)*2      (bytes; all bf16, value = hi + lo) */
int dcl_fda_workspace_layout(int b, int c, int p, int n, int m, size_t* offsets3);

/* dcl_fda_fwd_packed with a choice of output formats: each of RE_embed / RI_embed goes out
 * as the reference's fp32 channel-major tensor (RE_embed, RI_embed), as a point-major
 * bf16 hi/lo image over the b*n query rows (RE_pm: 256 channels, RI_pm: c channels — the
 * activation format of dcl_pm_gemm, so the fuser / confidence MLPs of
 * models/DCL_Net.py:216-228 read the aligned features without a repack), or both.
 * NULL outputs are skipped. */
int dcl_fda_fwd_packed_pm(int b, int c, int p, int n, int m,
    float* RE_embed, float* RI_embed, void* RE_pm, void* RI_pm, float* lse_out,
    void* workspace, size_t workspace_bytes, void* stream);

/* Materialise the attention map A (b,m,n) of models/Modules.py:167 from the
 * inputs and the lse of dcl_fda_align_fwd (train / inspection path only). */
int dcl_fda_attention_map(int b, int c, int n, int m,
    const float* RI_1, const float* RI_2, const float* lse, float* A, void* stream);

/* replaces ortho9d2matrix's torch.svd + det + diag_embed + matmuls
 * (models/DCL_Net.py:15-36, models/refiner.py:35-56).
 * in9 (b,9): when normalize_columns != 0, rows are [x_raw | y_raw | z_raw] and each
 * 3-vector is scaled by 1/(|v|+1e-8) (utils/transform3D.py:18-20) and used as a
 * COLUMN of M; when 0, in9 is a row-major 3x3 M used as is.
 * out R (b,3,3) = U diag(1,1,det(U V^T)) V^T. */
int dcl_svd3_project(int b, const float* in9, int normalize_columns, float* R, void* stream);

/* Confidence-weighted Kabsch (BASELINE.json north_star part 3; no reference
 * counterpart — SURVEY.md D2/a15).  src,dst (b,n,3), w (b,n) ->
 * R (b,3,3), t (b,3) minimising sum_i w_i |R src_i + t - dst_i|^2. */
int dcl_weighted_kabsch(int b, int n, const float* src, const float* dst, const float* w,
    float* R, float* t, void* stream);

/* Stage-2 pose composition (tools/test_YCBV_stage2.py:210,222-225):
 *   t' = R dt + t ;  R' = R dR ;  out[b, ch, i] = sum_r (points_in[b,i,r] - t'[r]) R'[r][ch]
 * R (b,3,3) and t (b,3) are updated in place; dR/dt may both be NULL (no update: only
 * canonicalise with the current pose, test_YCBV_stage2.py:210).  The canonicalised cloud
 * is written CHANNEL-MAJOR, i.e. already transposed as the refiner input wants it
 * (test_YCBV_stage2.py:212,225): channel ch of instance b starts at
 * points_out_cm + b*out_batch_stride + ch*n, so it can be the first three channels of
 * the (b, 3+256, n) refiner input.  points_in/points_out_cm may be NULL (pose update only). */
int dcl_pose_compose(int b, int n, float* R, float* t, const float* dR, const float* dt,
    const float* points_in, float* points_out_cm, int64_t out_batch_stride, void* stream);
/* Same, additionally (or instead: points_out_cm may be NULL) writing the canonicalised cloud as a
 * point-major bf16 hi/lo image with 32 channels per row (channels 0-2 = xyz, the rest must have been
 * zeroed once by the caller; (b*n) % 128 == 0): the second operand block of the refiner's first layer
 * on tensor cores (models/refiner.py:61-63 via dcl_pm_gemm, X = [F_Xo_p image | this image]). */
int dcl_pose_compose_pm(int b, int n, float* R, float* t, const float* dR, const float* dt,
    const float* points_in, float* points_out_cm, int64_t out_batch_stride,
    void* points_out_pm, void* stream);
/* The same writing a PM16 image (fp16, 32 channels per row): channels 0-2 = fp16(xyz), channels 3-5 = fp16 of the
 * remainder xyz - fp16(xyz) — coordinates keep ~22 bits; the consumer's weight columns for channels 3-5 repeat those
 * of channels 0-2 (X W^T = hi W^T + lo W^T). */
int dcl_pose_compose_pm16(int b, int n, float* R, float* t, const float* dR, const float* dt,
    const float* points_in, float* points_out_cm, int64_t out_batch_stride,
    void* points_out_pm16, void* stream);

/* ------------------------------------------------------------------------- */
/* Group 3: pointwise MLP stacks on tensor cores (replace cuDNN/cuBLAS calls)     */
/* ------------------------------------------------------------------------- */

/* "PM image" of an (R x C) activation (R % 128 == 0, C % 32 == 0; 4*R*C bytes): blobs of 128 rows x 32
 * channels, each blob = bf16 hi image (8 KB) then bf16 lo image (8 KB), value = hi + lo:
 *   byte(r,c,half) = ((r/128)*(C/32) + c/32)*16384 + half*8192 + ((r%128)/8)*512 + ((c%32)/8)*128 + (r%8)*16 + (c%8)*2
 * "PM16 image" (format DCL_PM_FMT_F16; 2*R*C bytes): the same blobs holding ONE fp16 image (8 KB) of the value
 * rounded once to fp16 (clamped to +-65504):
 *   byte(r,c) = ((r/128)*(C/32) + c/32)*8192 + ((r%128)/8)*512 + ((c%32)/8)*128 + (r%8)*16 + (c%8)*2
 * Its packed weights have the layout below with fp16 hi / lo halves instead of bf16 ones.  The inference path uses
 * PM16 for every activation (2 MMAs per product instead of 3, half the bytes); what the single rounding costs
 * against the path's tolerances is measured in profiles/r02_precision_emulation_*.json and by the parity tests.
 * Packed weights of a (cout x cin) layer, n-tile width nt in {64,128,256} (cout % nt == 0, cin % 32 == 0):
 *   byte(o,i,half) = ((o/nt)*(cin/32) + i/32)*(nt*128) + half*(nt*64) + ((o%nt)/8)*512 + ((i%32)/8)*128 + (o%8)*16 + (i%8)*2
 *
 * One problem:  Y = act(X W^T + bias) with X = [a0 | a1] concatenated along channels (a0 contributes kb0
 * k-blocks of 32 channels and must be exactly that wide; a1 the remaining kb_total - kb0), i.e. the
 * Conv3d/Conv1d 1x1 (+folded BatchNorm) + ReLU layers of models/DCL_Net.py:56-151 and the
 * Conv1d -> ReLU -> BatchNorm layers of models/Modules.py:173-201 (post_scale/post_shift = eval-mode BN
 * applied after the activation).  Outputs, each optional: out_pm (PM image of Y), out_cm (fp32, (R/rows_per_inst,
 * cout, rows_per_inst) channel-major as the reference holds activations), pool_out ((R/32) x cout partial sums over
 * each 32-row group of pool_w[r] * Y[r,:], for the confidence-weighted pooling of models/DCL_Net.py:228). */
typedef struct dcl_pm_gemm_problem {
    const void* a0;
    const void* a1;
    int kb0;
    int kb_total;
    const void* w;
    const float* bias;
    const float* post_scale;
    const float* post_shift;
    int relu;
    int cout;
    int nt;
    void* out_pm;
    float* out_cm;
    int rows_per_inst;
    const float* pool_w;
    float* pool_out;
    const float* dot_w;   /* [cout]; needs cout == nt */
    float* dot_out;       /* [R]: sum_o Y[r,o] * dot_w[o] — a trailing cout -> 1 layer without its bias */
    /* Y written straight into the operand images of the fused FDA kernel (dcl_fda_workspace_layout), so the
     * disengage layers feed the attention without the dcl_fda_pack pass: out_qk = query (qk_tile_rows 128) or
     * key (64) image, Y being the whole RI tensor (cout == the FDA's c); out_v = value image, Y's columns
     * becoming value rows [v_row0, v_row0 + cout) of v_rows (= 256 + c) — RE_2 at row 0, RI_2 at row 256. */
    void* out_qk;
    int qk_tile_rows;
    void* out_v;
    int v_row0;
    int v_rows;
    /* Operand formats (DCL_PM_FMT_*), see "PM16 image" above.  a_fmt: format of a0 / a1 AND of the packed weights
     * (0: bf16 hi/lo images, 3 MMAs per product; 1: fp16 activations rounded once + fp16 hi/lo weights, 2 MMAs);
     * equal for all problems of a launch.  out_fmt: format of out_pm and out_v (out_qk is always bf16 hi/lo: the
     * logits of the FDA softmax keep split operands). */
    int a_fmt;
    int out_fmt;
    /* Strided batch (the weight-gradient launches of the training path, dcl_tr_*): inst_count > 1 turns the problem
     * into inst_count independent slices s = 0..inst_count-1 reading a0 + s*a_inst_stride and w + s*w_inst_stride
     * (bytes) and writing out_cm + s*out_cm_inst_stride (bytes): split-K over instances, one fp32 partial per slice.
     * Needs kb0 == kb_total and out_cm as the only output; equal inst_count for all problems of a launch. */
    int inst_count;
    long long a_inst_stride;
    long long w_inst_stride;
    long long out_cm_inst_stride;
} dcl_pm_gemm_problem;
#define DCL_PM_FMT_BF16X2 0
#define DCL_PM_FMT_F16 1

/* Up to 8 problems with equal (cout, nt) over the same number of rows in ONE launch (grid.z = problem). */
int dcl_pm_gemm(int nproblems, const dcl_pm_gemm_problem* problems, int rows, void* stream);
/* fp32 row-major (rows x c, leading dimension ld) -> PM image (fmt 0) or PM16 image (fmt 1). */
int dcl_pm_pack_rows(int rows, int c, int ld, const float* src, void* dst_pm, int fmt, void* stream);
/* fp32 channel-major (b, c, n) -> PM / PM16 image of the (b*n x c) activation. */
int dcl_pm_pack_cm(int b, int c, int n, const float* src, void* dst_pm, int fmt, void* stream);
/* PM / PM16 image -> fp32 row-major (rows x c). */
int dcl_pm_unpack(int rows, int c, const void* src_pm, float* dst, int fmt, void* stream);
/* out[inst, :] (+)= sum of `parts` consecutive partial rows per instance, in index order, then
 * (partials2 != NULL) of the second set's, continuing the same running sum. */
int dcl_pm_pool_reduce(int insts, int cout, int parts, const float* partials, const float* partials2,
    float* out, int accumulate, void* stream);
/* dcl_sp_nn_interpolate_fused writing columns [out_col0, out_col0+c) of a PM image with c_total channels
 * (n % 128 == 0, c % 8 == 0, out_col0 % 8 == 0) instead of an fp32 matrix. */
int dcl_sp_nn_interpolate_fused_pm(int n, int m, int c,
    const float* unknown, const float* known, const float* feats,
    void* out_pm, int c_total, int out_col0,
    void* workspace, size_t workspace_bytes, void* stream);

/* The two pose regressors (models/DCL_Net.py:139-151,230-235; models/refiner.py:66-77) on the pooled feature
 * (b, d_in): per head three dense layers d_in -> d_h1 -> d_h2 -> d_out with ReLU after the first two
 * (Conv1d(k=1) on a (b, d_in, 1) tensor).  Weights are the Conv1d tensors as they are, fp32 row-major
 * (d_out x d_in); every width <= 1024.  out_rot (b, rot_head->d_out), out_trans (b, trans_head->d_out). */
typedef struct dcl_pose_head_mlp {
    const float* w1; const float* b1;
    const float* w2; const float* b2;
    const float* w3; const float* b3;
    int d_in, d_h1, d_h2, d_out;
} dcl_pose_head_mlp;
size_t dcl_pose_head_workspace_bytes(int b, const dcl_pose_head_mlp* rot_head, const dcl_pose_head_mlp* trans_head);
int dcl_pose_head(int b, const float* pooled, const dcl_pose_head_mlp* rot_head,
    const dcl_pose_head_mlp* trans_head, float* out_rot, float* out_trans,
    void* workspace, size_t workspace_bytes, void* stream);

/* dcl_sp_nn_interpolate_fused_pm with Ops_tensor2points (models/Modules.py:204-211) fused in: the known rows are the
 * sparse tensor's int32 (m,4) indices (b,ix,iy,iz); their centres ((float(i)*ext)+offset)+0.5*ext are formed in the
 * kernels in exactly torch's fp32 evaluation order.  voxel_extent3 / offset3 are HOST pointers to 3 floats. */
int dcl_sp_nn_interpolate_vox_pm(int n, int m, int c,
    const float* unknown, const int* vox_indices, const float* voxel_extent3, const float* offset3,
    const float* feats, void* out_pm, int c_total, int out_col0,
    void* workspace, size_t workspace_bytes, void* stream);

/* replaces models/DCL_Net.py:219-220: conf = sigmoid(cat([conf_1, conf_2], dim=2));
 * conf_softmax = softmax(conf, dim=2).  logit_1 / logit_2 (b,n) are the outputs of the last
 * regressor_conf / regressor_conf_bi layer WITHOUT its bias (bias_1 / bias_2: one device float
 * each).  conf (b,2n); w1 / w2 (b*n) = conf_softmax[:, :n] / [:, n:]. */
int dcl_conf_weights(int b, int n, const float* logit_1, const float* logit_2,
    const float* bias_1, const float* bias_2, float* conf, float* w1, float* w2, void* stream);

/* Nearest-point distance from every point of cloud a (b,n,3) to cloud bpts (b,m,3):
 * min_dist[b,i] = min_j ||a_i - bpts_j|| (Euclidean), argmin (b,n; may be NULL) = the lowest such j.
 * Replaces the B x N x M x 3 broadcast of CD_Dis (models/DCL_Net.py:307-311, models/refiner.py:129-133:
 * 0.5 * (min over dim 2 + min over dim 1) = two calls with the clouds swapped) and of the ADD-S metric
 * (tools/test_YCBV_stage1.py:188: mean over n of min_dist). */
int dcl_nearest_dist(int b, int n, int m, const float* a, const float* bpts,
    float* min_dist, int* argmin, void* stream);

/* All pyramid levels of one tower at once (Ops_GetPointFeat_spconv.forward,
 * models/Modules.py:227-251, calls Ops_nearest_neighbor_interpolate once per level with the
 * same query points): one launch builds every level's batch buckets, one launch searches
 * and interpolates every level into its column range of the same point-major image.
 * Same results, bit for bit, as nlevels calls of dcl_sp_nn_interpolate_vox_pm.
 * voxel_extent / offset are host values; vox_indices (m,4) int32 and feats (m,c) fp32 are
 * device pointers.  nlevels <= 8.  More than 2048 (batch id, slab) buckets take the full-scan
 * fallback; a level
 * with m == 0 leaves its columns untouched. */
typedef struct dcl_sp_level {
    int m, c, out_col0;
    int grid_x;   /* first voxel indices lie in [0, grid_x) (the level's grid size; <= 128): lets the
                   * search walk slabs of equal first coordinate outwards from the query and stop early.
                   * 0 disables it; an index outside the range only costs the full-scan fallback. */
    const int* vox_indices;
    float voxel_extent[3];
    float offset[3];
    const float* feats;
} dcl_sp_level;
/* One tower = one set of n query points (n,4 bxyz) and its levels; both towers of the network
 * (observed cloud / template cloud) go in ONE pair of launches, <= 8 levels in total.  The
 * workspace must hold the sum of dcl_sp_levels_workspace_bytes over the towers. */
typedef struct dcl_sp_tower {
    int n, c_total, nlevels;
    const float* unknown;
    void* out_pm;
    const dcl_sp_level* levels;
    int out_fmt;   /* DCL_PM_FMT_BF16X2 (PM image) or DCL_PM_FMT_F16 (PM16 image) */
} dcl_sp_tower;
int dcl_sp_nn_interpolate_towers_pm(int ntowers, const dcl_sp_tower* towers,
    void* workspace, size_t workspace_bytes, void* stream);
size_t dcl_sp_levels_workspace_bytes(int nlevels, const dcl_sp_level* levels);
int dcl_sp_nn_interpolate_levels_pm(int n, const float* unknown, int nlevels,
    const dcl_sp_level* levels, void* out_pm, int c_total,
    void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------- */
/* Group 4: the input side — voxelisation and the sparse-conv towers   */
/* ------------------------------------------------------------------------- */
/* Replaces, on the device and without a host round trip:
 *   pointgroup_ops.voxelization_idx  (libs/pointgroup_ops/src/voxelize/voxelize.cpp:10-163, a CPU hash map in the
 *                                     dataloader) and pointgroup_ops.voxelization mode 4 (voxelize.cu:10-31),
 *   spconv get_indice_pairs          (libs/spconv/include/spconv/spconv_ops.h:27-136, indice.cu.h:24-220) for the
 *                                     SparseConv3d / SubMConv3d / SparseAvgPool3d of Backbone_SPCONV
 *                                     (models/Modules.py:100-159),
 *   spconv indice_conv / indice_avgpool (spconv_ops.h:253-349, src/spconv/avgpool.cu).
 * A voxel set of an instance is a bit grid (G*G rows of G bits); a tower has 9 sets — 0: occupied voxels (64^3),
 * 2l+1: output set of level l's SparseConv3d (= input / output set of its SubMConv3d; grid 64>>l), 2l+2: output set of
 * level l's SparseAvgPool3d (grid 32>>l) — and 12 ops, op 3l+{0,1,2} = level l's {SparseConv3d, SubMConv3d, pool}.
 * Rows of a set are packed over the batch in the reference's order (batch, then linear voxel index). */
#define DCL_SPB_NSETS 9
#define DCL_SPB_NOPS 12
/* 64-bit rows (and int prefixes) per instance over the 9 sets: 2*64^2 + 2*32^2 + 2*16^2 + 2*8^2 + 4^2. */
size_t dcl_spb_rows_per_instance(void);

/* One tower's inputs and voxelisation outputs.  points (b*n_per,3) fp32 [+ rgb (b*n_per,3)], or voxel coordinates
 * coords (b*n_per,4) int32 bxyz when points == NULL.  rows / prefix: [b][dcl_spb_rows_per_instance()] uint64 / int32;
 * counts: [9][b].  Voxelisation outputs use SLOTS of n_per rows per instance and may be NULL:
 *   feat16     (b*n_per,16) fp16: operand rows of the first convolution, voxels in SORTED order (rank in the bit grid):
 *              channels 0-6 = mean of [1, rgb, xyz] over the voxel's points in point order with the 1/n multiplier
 *              applied first (voxelize.cu:15-21), channels 7-12 = fp16 remainders of channels 1-6, 13-15 zero;
 *   feat32     (b*n_per,7) fp32 the same means, voxels in FIRST-APPEARANCE order (the reference's numbering,
 *              voxelize.cpp:96-107), as are occupied (b*n_per,4) int32 bxyz, p2v (b*n_per) (voxel of each point, local
 *              to the instance), v2p_sorted (b*n_per) (local point indices grouped by voxel, ascending) and
 *              v2p_start (b*n_per) (start of each voxel's group).
 * errors[0] += number of points outside the grid (skipped). */
typedef struct dcl_spb_tower_in {
    const float* points;
    const float* rgb;
    const int* coords;
    void* rows;
    int* prefix;
    int* counts;
    void* feat16;
    float* feat32;
    int* occupied;
    int* p2v;
    int* v2p_sorted;
    int* v2p_start;
    int* errors;
} dcl_spb_tower_in;
/* All nine voxel sets of every instance of up to two towers in ONE launch (one CTA per instance and tower).
 * unit = voxel edge (0.006), grid must be 64, n_per <= 4096. */
int dcl_spb_build_sets(int b, int n_per, int ntowers, const dcl_spb_tower_in* towers, float unit, int grid,
    void* stream);

/* Row coordinates and rulebooks.  offsets: [9][b+1] (exclusive scan of counts over the batch; element b = rows of
 * the set); indices[s]: (cap[s],4) int32 bxyz, rows past the total get batch id = b (cap % 128 == 0);
 * nbr[op]: (cap_out/128, 32, 128) int32, tile-transposed — nbr[tile][k][row % 128], k in 0..26: input row feeding
 * output row tile*128 + row%128 through kernel offset k = (k0*3+k1)*3+k2 (input voxel = out*stride - 1 + k,
 * geometry.h:24-86) or -1; slot 27: number of valid entries;
 * anymask[op]: (cap_out/128) 27-bit OR of a tile's validity masks.  in0_slot > 0: the input rows of op 0 are
 * dcl_spb_tower_in.feat16's slots (instance*in0_slot + rank).  errors[1] |= 1<<s when set s exceeds cap[s]. */
typedef struct dcl_spb_tower_sets {
    const void* rows;
    const int* prefix;
    const int* counts;
    int* offsets;
    int* indices[DCL_SPB_NSETS];
    int cap[DCL_SPB_NSETS];
    int* nbr[DCL_SPB_NOPS];
    unsigned int* anymask[DCL_SPB_NOPS];
    int* errors;
} dcl_spb_tower_sets;
int dcl_spb_emit(int b, int ntowers, const dcl_spb_tower_sets* towers, int in0_slot, void* stream);

/* One sparse convolution (3x3x3) of up to two towers in one launch, output-stationary on tensor cores:
 *   out[r, :] = relu( sum_k in[nbr[r,k], :] W[k] + shift )        (BatchNorm1d folded: W scaled, shift = bias)
 * in16: (rows_in, cin_pad) fp16 operand rows, cin_pad in {16,32,64,128}; w: packed fp16 hi/lo weights
 * [stage of 64 virtual channels kv = k*cin_pad + c][hi | lo][cout x 64] (K-major core matrices, see dcl_net_b200/backbone.py:pack_conv_weight);
 * cout in {16,32,64,128,256}; total rows read from offsets_out[b].  out16 (rows, cout) fp16 and / or out32 fp32. */
typedef struct dcl_spb_conv {
    const void* in16;
    const int* nbr;
    const unsigned int* anymask;
    const int* offsets_out;
    const void* w;
    const float* shift;
    void* out16;
    float* out32;
    int cap_out;
} dcl_spb_conv;
int dcl_spb_conv3(int b, int cin_pad, int cout, int ntowers, const dcl_spb_conv* convs, void* stream);

/* SparseAvgPool3d(k3, s2, p1, use_gs=False): out[r] = sum over kernel offsets k ascending of in[nbr(r,k)] / nbr(r,27)
 * (src/spconv/avgpool.cu:44, summaryRF.cu:27-41); in (rows_in, c) fp32 -> out32 (cap_out, c) fp32 [+ out16 fp16]. */
typedef struct dcl_spb_pool {
    const float* in;
    const int* nbr;
    const int* offsets_out;
    float* out32;
    void* out16;
    int cap_out;
} dcl_spb_pool;
int dcl_spb_avgpool(int b, int c, int ntowers, const dcl_spb_pool* pools, void* stream);

/* replaces voxelize_fp_cuda (libs/pointgroup_ops/src/voxelize/voxelize.cu:10-31) for mode 4 (mean):
 * feats (n,c), rules (m,width) int32 rows [count, i_1 .. i_count, ...] -> out (m,c), summed in rule order. */
int dcl_voxelize_mean(int m, int width, int c, const float* feats, const int* rules, float* out, void* stream);

/* ------------------------------------------------------------------------- */
/* Group 5: training path of the pointwise MLP stacks (BASELINE.json configs[4])  */
/* ------------------------------------------------------------------------- */
/* Replaces, in train mode, the cuDNN/cuBLAS calls autograd makes for the Conv3d/Conv1d(k=1) + BatchNorm + ReLU
 * layers of models/DCL_Net.py:56-151 and models/Modules.py:58-97,173-201 (forward and backward), stepped by
 * tools/train_YCBV_stage1.py:168-191.  Every layer is three dcl_pm_gemm calls on bf16 hi/lo operand images
 * (forward; dgrad dX = dZ W with the packed transpose of W as "weights"; wgrad dW = sum_b dZ_b^T X_b as a strided
 * batch over the per-instance TRANSPOSED images, split-K over instances) plus the HBM-bound passes below.
 * Layer kinds: conv(+bias); conv+bias -> ReLU; conv -> BN -> ReLU (disengage blocks); conv+bias -> ReLU -> BN
 * (neck fusers).  U = what the GEMM epilogue writes (fp32 (b,c,n): conv output after bias / ReLU where they follow
 * the conv directly); BN scale = gamma*rstd, shift = beta - mean*scale from the batch statistics. */
#define DCL_TR_COPY         0   /* y = x */
#define DCL_TR_AFFINE       1   /* y = x*scale + shift                      (forward: ReLU -> BN layers)   */
#define DCL_TR_AFFINE_RELU  2   /* y = relu(x*scale + shift)                (forward: BN -> ReLU layers)   */
#define DCL_TR_BWD_RELU     3   /* dZ = dY * [u > 0]                        (conv+bias -> ReLU)            */
#define DCL_TR_BWD_BN_RELU  4   /* dZ = scale*(g - s1/cnt - xhat*s2/cnt), g = dY*[u*scale+shift > 0]       */
#define DCL_TR_BWD_RELU_BN  5   /* dZ = [u > 0] * scale*(dY - s1/cnt - xhat*s2/cnt)                        */
/* One pass over an activation (or gradient): the pointwise transform `mode`, then any of the operand images the
 * GEMMs read.  x: fp32 with element strides (x_sb, x_sc, x_sn) over (instance, channel, point) — channel-major
 * (b,c,n) tensors and their channel slices, or point-major (b*n, c) matrices (x_sc == 1); u: contiguous (b,c,n)
 * (backward modes); xhat = (u - mean)*rstd, cnt = b*n.  c % 32 == 0, n % 128 == 0.  Outputs (NULL = skipped):
 *   out_k       PM image of the (b*n x c) result (A operand of the forward / dgrad GEMM); with k_cols > 0 the result
 *               becomes columns [k_col0, k_col0+c) of a (b*n x k_cols) image (k_col0 % 32 == 0): concatenations;
 *   out_t       per instance the PM image of the (t_rows x n) matrix [channel][point], this tensor's channels at rows
 *               [t_row0, t_row0+c) (t_rows % 128 == 0, t_row0 % 32 == 0; rows never written must be zeroed by the
 *               caller once) — operands of the wgrad GEMM; instance stride t_rows*n*4 bytes;
 *   out_cm      the result as fp32 (b,c,n);
 *   col_partial (b*n/128, c) sums of the result over each 128-point tile (bias gradient; reduce with
 *               dcl_pm_pool_reduce). */
typedef struct dcl_tr_tile {
    const float* x;
    const float* u;
    long long x_sb, x_sc, x_sn;
    int b, c, n;
    int mode;
    const float* scale;
    const float* shift;
    const float* mean;
    const float* rstd;
    const float* s1;
    const float* s2;
    void* out_k;
    void* out_t;
    int t_row0, t_rows;
    float* out_cm;
    float* col_partial;
    int k_col0, k_cols;
    int t_group;   /* > 1: out_t holds one image per GROUP of t_group consecutive instances, (t_rows x t_group*n), the
                    * instances side by side along the points (fewer, longer split-K slices for the wgrad GEMM) */
} dcl_tr_tile;
/* Up to 8 items per launch. */
int dcl_tr_tile_pass(int nitems, const dcl_tr_tile* items, void* stream);
/* Train-mode BatchNorm statistics (torch.nn.BatchNorm1d/3d forward in training): per channel of the contiguous
 * (b,c,n) tensor u the batch mean and rstd = 1/sqrt(biased var + eps), scale / shift as above, and — when
 * running_mean is given — the running statistics update with `momentum` (unbiased variance), in place. */
typedef struct dcl_tr_bn {
    const float* u;
    int b, c, n;
    const float* gamma;
    const float* beta;
    float eps, momentum;
    float* running_mean;
    float* running_var;
    float* mean;
    float* rstd;
    float* scale;
    float* shift;
} dcl_tr_bn;
int dcl_tr_bn_stats(int nitems, const dcl_tr_bn* items, void* stream);
/* BatchNorm backward sums per channel: s1 = sum g (= d beta), s2 = sum g*xhat (= d gamma); g as in the modes
 * DCL_TR_BWD_BN_RELU / DCL_TR_BWD_RELU_BN.  dy: fp32 (b,c,n) with instance / channel strides dy_sb / dy_sc. */
typedef struct dcl_tr_bn_bwd {
    const float* dy;
    const float* u;
    long long dy_sb, dy_sc;
    int b, c, n;
    int mode;
    const float* mean;
    const float* rstd;
    const float* scale;
    const float* shift;
    float* s1;
    float* s2;
} dcl_tr_bn_bwd;
int dcl_tr_bn_bwd_reduce(int nitems, const dcl_tr_bn_bwd* items, void* stream);
/* fp32 weights -> packed bf16 hi/lo blobs of dcl_pm_gemm (n-tile nt): the (rows x cols) matrix src (row-major), or —
 * transpose != 0 — the transpose of the (cols x rows) matrix src (the dgrad "weights" W^T), zero-padded to
 * (rows_pad x k_pad); rows_pad % nt == 0, k_pad % 32 == 0. */
typedef struct dcl_tr_wpack {
    const float* src;
    void* dst;
    int rows, cols;
    int rows_pad, k_pad;
    int nt;
    int transpose;
    int k_col0, k_total;   /* k_total > 0: the matrix becomes columns [k_col0, k_col0+k_pad) of a (rows_pad x k_total)
                            * packed matrix (k_col0 % 32 == 0): weights concatenated along the reduction axis */
} dcl_tr_wpack;
int dcl_tr_pack_weights(int nitems, const dcl_tr_wpack* items, void* stream);
/* out[c] = sum over the `parts` rows of partial (parts x c), in a fixed order: the bias gradients from the per-tile
 * column sums of dcl_tr_tile_pass (col_partial).  Up to 8 items per launch. */
typedef struct dcl_tr_colsum {
    const float* partial;
    float* out;
    int parts, c;
} dcl_tr_colsum;
int dcl_tr_colsum_reduce(int nitems, const dcl_tr_colsum* items, void* stream);

/* Fused backward of the FDA (models/Modules.py:166-169 under autograd; forward = dcl_fda_align_fwd): with
 * S = RI_2^T RI_1, A = softmax_m S, RE_embed = RE_2 A, RI_embed = RI_2 A and the output gradients gE (b,p,n), gI (b,c,n):
 *   d_q (b,c,n)      = d RI_1                 = RI_2 dS,          dS = A o (dA - D),  dA = [RE_2; RI_2]^T [gE; gI]
 *   d_k (b,c,m)      = d RI_2 (key role)      = RI_1 dS^T
 *   d_v (b,p+c,m)    = [d RE_2; d RI_2 (value role)] = [gE; gI] A^T
 * with A rebuilt on chip from the forward's lse (b,n) and D = dsum (b,n) = sum_p gE o RE_embed + sum_c gI o RI_embed;
 * no (m x n) matrix is written to memory.  Operands are bf16 hi/lo PM images written by dcl_tr_tile_pass:
 * "_k" = PM image of the (b*points x channels) matrix, "_t" = per-instance PM images of the (channels x points)
 * matrix with the channel rows padded to a multiple of 128 (zero rows):
 *   q_k, q_t: RI_1 (c wide);  k_k, k_t: RI_2;  v_k: [RE_2 | RI_2] (p+c wide);  g_k, g_t: [gE | gI] (p+c wide).
 * c in {64,128}, p = 256, n % 128 == 0, m % 128 == 0.  Up to two jobs of equal shape per call (the two directions). */
typedef struct dcl_fda_bwd_job {
    const void* q_k;
    const void* k_k;
    const void* g_k;
    const void* v_k;
    const void* q_t;
    const void* k_t;
    const void* g_t;
    const float* lse;
    const float* dsum;
    float* d_q;
    float* d_k;
    float* d_v;
} dcl_fda_bwd_job;
int dcl_fda_bwd(int njobs, const dcl_fda_bwd_job* jobs, int b, int c, int p, int n, int m, void* stream);

/* ------------------------------------------------------------------------- */
/* Bring-up / test hook                                                       */
/* ------------------------------------------------------------------------- */
/* D (128 x N) = A (128 x K) B^T (B is N x K), all row-major fp32, through exactly the
 * operand packing, UMMA descriptors, hi/lo split product and TMEM read-back of
 * dcl_fda_align_fwd, in a single CTA.  N % 32 == 0, N <= 256, K % 16 == 0.
 * swap_lbo_sbo must be 0 (1 exchanges the descriptor stride fields; used once on
 * hardware to pin the descriptor semantics). */
int dcl_debug_umma_gemm(int N, int K, const float* A, const float* B, float* D,
    int swap_lbo_sbo, void* stream);

/* D (256 x N) = A (256 x K) B^T issued by the leader of a 2-CTA cluster as one M=256
 * tcgen05.mma.cta_group::2 per K step: CTA r holds A rows [128r,128r+128) and B rows
 * [rN/2,(r+1)N/2).  mode 0: cluster barrier before the MMA; mode 1: the peer hands its
 * operands over with a remote mbarrier arrive.  Pins the CTA-pair conventions. */
int dcl_debug_umma_pair_gemm(int N, int K, const float* A, const float* B, float* D,
    int mode, void* stream);

/* D (128 x N) = bf16(A) bf16(B)^T with the A operand read from TENSOR memory (the "TS" form of
 * tcgen05.mma), single bf16 product.  variant 0: two bf16 per 32-bit TMEM column; 1: one per column.
 * Bring-up probe for keeping a hidden activation tile on chip between two layers; not used by the
 * product (tools/probe_umma_ts.py). */
int dcl_debug_umma_ts_gemm(int N, int K, const float* A, const float* B, float* D,
    int variant, void* stream);

/* Installs (or, with NULL, removes) a device buffer of 4*1024 int64 into which CTA (0,0) of
 * the FDA kernels stamps clock64() at its role hand-offs: [role*1024 + key_block*8 + event],
 * roles 0 = MMA issuer, 1 = softmax warp 2, 2 = TMA producer; [3*1024 + slot*8 + event] holds
 * globaltimer life-cycle stamps of the first and the last CTA (tools/trace_fda.py). */
int dcl_debug_fda_set_trace(long long* device_buffer);

/* Installs (or, with NULL, removes) a device buffer of 6*512 int64 into which CTA (0,0,0) of the sparse convolution
 * stamps clock64() per role and pipeline iteration (csrc/sparse_conv.cu; tools/trace_spconv.py). */
int dcl_debug_spconv_set_trace(long long* device_buffer);

#ifdef __cplusplus
}
#endif
#endif /* DCL_B200_H */
