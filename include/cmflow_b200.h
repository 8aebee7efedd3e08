/*
 * cmflow_b200.h -- C ABI of libcmflow_b200.so (B200 / sm_100a).
 *
 * Plain pointers, sizes and a stream; no torch types.  Every entry point returns 0 on success or a
 * non-zero CMF_ERR_* code (the reference prints to stderr and calls exit(-1) instead, e.g.
 * lib/src/ball_query_gpu.cu:62-66 -- a library must not kill its host).  cmf_last_error() returns a
 * thread-local human-readable message for the last failure.  All device pointers must be contiguous
 * fp32 / int32 buffers with the layouts stated; `stream` is a cudaStream_t passed as void* (0 = the
 * legacy default stream).  Nothing here allocates device memory except cmf_model_* (workspace owned
 * by the handle) and nothing synchronises the stream except the *_host entry point.
 *
 * Part 1 replaces, one for one, the launchers the reference's pybind module `pointnet2_cuda` binds
 * (lib/src/pointnet2_api.cpp:11-24); file:line of the replaced launcher prototype is given per entry.
 * Part 2 is the model-path operators of utils/model_utils/radarflow_util.py and models/cmflow.py.
 * Part 3 is the whole-forward engine behind models/cmflow.py:171-197 / models/cmflow_t.py:185-211.
 * Paths are relative to the upstream tree (Toytiny/CMFlow @ 16a095a).
 */
#ifndef CMFLOW_B200_H
#define CMFLOW_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CMF_OK 0
#define CMF_ERR_INVALID 1   /* bad argument (null pointer, negative size, k out of range ...) */
#define CMF_ERR_CUDA 2      /* a CUDA runtime call or kernel launch failed                     */
#define CMF_ERR_NOMEM 3     /* workspace allocation failed                                     */
#define CMF_ERR_STATE 4     /* handle used before weights were uploaded, etc.                  */

const char *cmf_last_error(void);
const char *cmf_version(void);                 /* "cmflow_b200 x.y.z sm_100a" */
int cmf_device_check(void);                    /* CMF_OK iff the current device is compute capability 10.x */

/* ------------------------------------------------------------------------------------------------
 * Part 1 -- pointnet2_cuda operator set
 * ---------------------------------------------------------------------------------------------- */

/* replaces ball_query_kernel_launcher_fast (lib/src/ball_query_gpu.h:12-13, ball_query_gpu.cu:48-67).
 * new_xyz (B,M,3), xyz (B,N,3) -> idx (B,M,nsample) int32, PRE-ZEROED by the caller
 * (lib/pointnet2_utils.py:246). First nsample indices in index order with d2 < radius^2 (strict),
 * remainder padded with the first hit, rows without a hit untouched. */
int cmf_ball_query(int b, int n, int m, float radius, int nsample,
                   const float *new_xyz, const float *xyz, int *idx, void *stream);

/* replaces group_points_kernel_launcher_fast (lib/src/group_points_gpu.h:13-14).
 * points (B,C,N), idx (B,npoints,nsample) -> out (B,C,npoints,nsample) */
int cmf_group_points(int b, int c, int n, int npoints, int nsample,
                     const float *points, const int *idx, float *out, void *stream);

/* replaces group_points_grad_kernel_launcher_fast (lib/src/group_points_gpu.h:19-20).
 * grad_out (B,C,npoints,nsample), idx -> grad_points (B,C,N), accumulated (caller zeroes). */
int cmf_group_points_grad(int b, int c, int n, int npoints, int nsample,
                          const float *grad_out, const int *idx, float *grad_points, void *stream);

/* replaces gather_points_kernel_launcher_fast (lib/src/sampling_gpu.h:12-13).
 * points (B,C,N), idx (B,npoints) -> out (B,C,npoints) */
int cmf_gather_points(int b, int c, int n, int npoints,
                      const float *points, const int *idx, float *out, void *stream);

/* replaces gather_points_grad_kernel_launcher_fast (lib/src/sampling_gpu.h:19-20). */
int cmf_gather_points_grad(int b, int c, int n, int npoints,
                           const float *grad_out, const int *idx, float *grad_points, void *stream);

/* replaces furthest_point_sampling_kernel_launcher (lib/src/sampling_gpu.h:26-27).
 * dataset (B,N,3), temp (B,N) in/out (caller fills 1e10, lib/pointnet2_utils.py:26), idxs (B,M) int32.
 * Tie-breaking reproduces the reference's block-size-dependent tree (sampling_gpu.cu:86-209). */
int cmf_furthest_point_sampling(int b, int n, int m,
                                const float *dataset, float *temp, int *idxs, void *stream);

/* replaces knn_kernel_launcher_fast (lib/src/interpolate_gpu.h:19-20, interpolate_gpu.cu:9-57).
 * unknown (B,N,3) queries, known (B,M,3) candidates -> dist2 (B,N,k) f32 SQUARED distances,
 * idx (B,N,k) int32; ascending, ties keep the lower index; 1 <= k <= 200 as in the reference. */
int cmf_knn(int b, int n, int m, int k, const float *unknown, const float *known,
            float *dist2, int *idx, void *stream);

/* replaces three_nn_kernel_launcher_fast (lib/src/interpolate_gpu.h:13-14). */
int cmf_three_nn(int b, int n, int m, const float *unknown, const float *known,
                 float *dist2, int *idx, void *stream);

/* replaces three_interpolate_kernel_launcher_fast (lib/src/interpolate_gpu.h:26-27).
 * points (B,C,M), idx (B,N,3), weight (B,N,3) -> out (B,C,N) */
int cmf_three_interpolate(int b, int c, int m, int n, const float *points, const int *idx,
                          const float *weight, float *out, void *stream);

/* replaces three_interpolate_grad_kernel_launcher_fast (lib/src/interpolate_gpu.h:33-34). */
int cmf_three_interpolate_grad(int b, int c, int n, int m, const float *grad_out, const int *idx,
                               const float *weight, float *grad_points, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Part 2 -- model-path operators
 * ---------------------------------------------------------------------------------------------- */

/* knn_point + square_distance (utils/model_utils/radarflow_util.py:88-99, 8-30): the k nearest
 * candidates of every query under the reference's EXPANDED float32 distance
 *   d = max(((-2*dot(q,x)) + |q|^2) + |x|^2, 0),  dot = fma(qz,xz, fma(qy,xy, qx*xx))
 * xyz (B,N,3) candidates, new_xyz (B,S,3) queries -> idx (B,S,k) int32 ascending by (d, index)
 * (torch.topk(sorted=False) leaves the order unspecified), optional dist (B,S,k) or NULL. 1<=k<=32. */
int cmf_knn_point(int b, int n, int s, int k, const float *xyz, const float *new_xyz,
                  int *idx, float *dist, void *stream);

/* Kernel-density estimate used by the soft chamfer loss (utils/util.py:172-182 compute_density_loss, called at losses/radar_loss.py:39-40):
 * density (B,N) = mean over the M candidates of exp(-d2 / (2 bw^2)) / (2.5 bw), d2 = the clamped expanded-form squared distance of
 * utils/util.py:148-170.  xyz1 (B,N,3) queries, xyz2 (B,M,3) candidates.  No (B,N,M) matrix is materialised. */
int cmf_kde_density(int b, int n, int m, const float *xyz1, const float *xyz2, float bandwidth, float *density, void *stream);

/* Multi-radius ball query of a cloud against itself, all four CMFlow scales in ONE pass over the
 * candidates (models/cmflow.py:21-22,35-36: r = 2,4,8,16; K = 4,8,16,32; QueryAndGroup.forward,
 * lib/pointnet2_utils.py:277).  xyz_planar (B,3,N) (the model's input layout) -> idx (B,N,60) int32:
 * per point the 4 rows concatenated [K=4 | K=8 | K=16 | K=32], each with cmf_ball_query semantics. */
int cmf_ball_query_ms(int b, int n, const float *xyz_planar, int *idx60, void *stream);

/* CMFlow.WeightedKabsch + refine_with_transform (models/cmflow.py:96-169, 112-125), one launch:
 * pc1 (B,3,N), flow (B,3,N), score (B,N) (stat_cls), eps added to the score before normalising
 * (1e-4 for CMFlow, cmflow.py:105; 0 for CMFlow-T, cmflow_t.py:119), stat_thres ->
 * trans (B,4,4), sf_agg (B,3,N), mask (B,N) uint8.  3x3 SVD in fp64 Jacobi; reproduces the
 * reference's row-2 flip of V (cmflow.py:162). */
int cmf_kabsch_refine(int b, int n, const float *pc1, const float *flow, const float *score,
                      float eps, float stat_thres, float *trans, float *sf_agg, uint8_t *mask, void *stream);

/* Weighted Kabsch alone (models/cmflow.py:128-169): A,Bp (B,3,N), W (B,N) normalised weights -> trans (B,4,4). */
int cmf_weighted_kabsch(int b, int n, const float *A, const float *Bp, const float *W, float *trans, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Part 3 -- whole-forward engine (models/cmflow.py:171-197, models/cmflow_t.py:185-211)
 * ---------------------------------------------------------------------------------------------- */

typedef struct cmf_model cmf_model;

/* Number of floats of the packed-weight blob for (temporal ? CMFlow_T : CMFlow); layout documented in
 * cmflow_b200/weights.py (BatchNorm folded into conv scale/bias in float64 on the host). */
size_t cmf_model_blob_floats(int temporal);

/* Create an engine on the current device (every later call on the engine switches to that device for its own duration and
 * restores the caller's).  `blob` is a HOST pointer to cmf_model_blob_floats() floats;
 * it is copied to the device.  stat_thres as models/cmflow.py:18. */
int cmf_model_create(cmf_model **out, const float *blob, size_t blob_floats, int temporal, float stat_thres);
void cmf_model_destroy(cmf_model *m);

/* RaFlow (models/raflow.py:11-164; the third model of models/model.py:21-27).  Same backbone as CMFlow -- pack the RaFlow state_dict with
 * cmflow_b200/weights.py:pack(..., raflow=True) into a non-temporal blob -- but no motion head, and the scene-flow refinement of
 * raflow.py:79-117 instead of the weighted Kabsch head.  cmf_model_set_raflow() switches an engine created with cmf_model_create(..., 0, ...)
 * to that model (rigid_thres: configs.yaml:29 / raflow.py:16; rigid_pcs: raflow.py:17 = 0.25); afterwards only
 * cmf_model_forward_raflow() may be called.  interval (B) seconds between the frames (main_util.py:139); outputs as RaFlow.forward returns
 * them: output (B,3,N) initial flow, sf_agg (B,3,N), pre_trans (B,4,4), mask_s (B,N) uint8.  Device pointers; enqueues on `stream`. */
int cmf_model_set_raflow(cmf_model *m, float rigid_thres, float rigid_pcs);
int cmf_model_forward_raflow(cmf_model *m, int b, int n,
                             const float *pc1, const float *pc2, const float *ft1, const float *ft2, const float *interval,
                             float *output, float *sf_agg, float *pre_trans, uint8_t *mask_s, void *stream);
/* The refinement alone (raflow.py:79-156) on a given initial flow: one CTA per pair, fp64 moments, 3x3 Jacobi SVD. */
int cmf_raflow_refine(int b, int n, const float *pc1, const float *ft1, const float *flow, const float *interval,
                      float rigid_thres, float rigid_pcs, float *sf_agg, float *trans, uint8_t *mask_s, void *stream);

/* Workspace bytes the engine holds for the largest (B,N) seen so far (diagnostic). */
size_t cmf_model_workspace_bytes(const cmf_model *m);

/* Number of (b, n, mode) shapes whose kernel sequence cmf_model_forward_host currently replays as a CUDA graph (opt-in: environment
 * CMF_HOST_GRAPH=1; second call of a shape onwards, on a capturable -- i.e. non-legacy-default -- stream; otherwise eager launches).  Diagnostic.
 * A cmf_model is not thread-safe: one host thread per engine at a time. */
int cmf_model_host_graphs(const cmf_model *m);

/* Number of kernels one forward launches (for bench.py's gpu_launches). */
int cmf_model_launches_per_forward(const cmf_model *m);

/* Diagnostics of a hung launch.  The tensor-core kernels bound every mbarrier wait by a wall-clock watchdog (30 s) that traps instead of
 * hanging the GPU; the resulting launch failure is sticky for the process.  Before trapping, the waiting thread records where it was:
 * out4 = {0 if no watchdog fired | 1 + kernel family (0 one-CTA GEMM, 1 CTA-pair GEMM, 2 fused set-conv #2, 3 chain kernels),
 * block << 32 | thread, grid << 32 | block size, nanoseconds waited}.  Callable after the failure (the record lives in host memory). */
int cmf_watchdog_read(unsigned long long *out4);

/* Device-resident forward. pc1,pc2,ft1,ft2 (B,3,N) fp32.  gfeat_prev (B,256) or NULL (CMFlow-T only;
 * NULL = zeros, cmflow_t.py:97-98).  Outputs: sf_agg (B,3,N), stat_cls (B,N) [the reference's (B,1,N)],
 * pre_trans (B,4,4), mask (B,N) uint8, gfeat_out (B,256) (CMFlow-T only, else may be NULL).
 * Enqueues on `stream`, does not synchronise. */
int cmf_model_forward(cmf_model *m, int b, int n,
                      const float *pc1, const float *pc2, const float *ft1, const float *ft2,
                      const float *gfeat_prev,
                      float *sf_agg, float *stat_cls, float *pre_trans, uint8_t *mask, float *gfeat_out,
                      void *stream);

/* General device-resident form.  n1 / n2 = points of cloud 1 / cloud 2 (n2 = 0 means n1): the reference's evaluation loop feeds
 * un-resampled clouds of different sizes, one pair at a time (dataset/vod.py:92-93 resamples only when training; main.py:203), and
 * FeatureCorrelator.forward handles N1 != N2 (radarflow_util.py:185-237).  pc2, ft2 are (B,3,n2); every output is per point of cloud 1.
 * label_m (B,n1) device floats or NULL: mode='train' with pseudo motion labels (models/cmflow.py:181-182, cmflow_t.py:196-197) -- the
 * labels replace the predicted scores in the ego-motion head's weights and in the refinement mask (cmflow.py:188); stat_cls still
 * returns the network's own scores.  Inference arithmetic otherwise (BatchNorm running statistics; no gradients). */
int cmf_model_forward2(cmf_model *m, int b, int n1, int n2,
                       const float *pc1, const float *pc2, const float *ft1, const float *ft2,
                       const float *gfeat_prev, const float *label_m,
                       float *sf_agg, float *stat_cls, float *pre_trans, uint8_t *mask, float *gfeat_out,
                       void *stream);
int cmf_model_forward_raflow2(cmf_model *m, int b, int n1, int n2,
                              const float *pc1, const float *pc2, const float *ft1, const float *ft2, const float *interval,
                              float *output, float *sf_agg, float *pre_trans, uint8_t *mask_s, void *stream);

/* Same with HOST buffers (pinned for full speed): copies inputs H2D, runs the forward, copies the four
 * outputs D2H and synchronises `stream`.  This is the end-to-end call bench.py times as `e2e`. */
int cmf_model_forward_host(cmf_model *m, int b, int n,
                           const float *pc1, const float *pc2, const float *ft1, const float *ft2,
                           const float *gfeat_prev,
                           float *sf_agg, float *stat_cls, float *pre_trans, uint8_t *mask, float *gfeat_out,
                           void *stream);

int cmf_model_forward_host2(cmf_model *m, int b, int n1, int n2,
                            const float *pc1, const float *pc2, const float *ft1, const float *ft2,
                            const float *gfeat_prev,
                            float *sf_agg, float *stat_cls, float *pre_trans, uint8_t *mask, float *gfeat_out,
                            void *stream);

/* Pipelined host form: the engine owns two staging slots (device buffers + copy streams).  cmf_model_submit_host(slot, ...) enqueues the
 * H2D copies of this call on the engine's upload stream, the kernels on `stream` (after the upload) and the D2H copies on the engine's
 * download stream (after the kernels) and returns without waiting; cmf_model_wait_host(slot) blocks until that call's outputs are in the
 * host buffers.  Alternating slots 0,1 overlaps the upload of call i+1 and the download of call i-1 with the kernels of call i.
 * Host buffers must be pinned for the copies to be asynchronous, and stay untouched until the slot has been waited on. */
int cmf_model_submit_host(cmf_model *m, int slot, int b, int n1, int n2,
                          const float *pc1, const float *pc2, const float *ft1, const float *ft2,
                          const float *gfeat_prev,
                          float *sf_agg, float *stat_cls, float *pre_trans, uint8_t *mask, float *gfeat_out,
                          void *stream);
int cmf_model_wait_host(cmf_model *m, int slot);

/* Arithmetic mode of the big 1x1-conv GEMMs: 0 = strict fp32 FMA (parity build, default), 1 = tcgen05 tensor cores with
 * 3xTF32 split precision (fp32 accumulate in TMEM; ~2^-22 relative per product), 2 = tcgen05 with 3xFP16 split precision
 * (same 22-bit split at twice the tensor rate; operands scaled by exact powers of two chosen per weight row / per frame pair
 * from measured maxima).  The default can also be chosen with the environment variable CMF_MODE=fp32|tf32x3|fp16x3 read at
 * cmf_model_create(). */
int cmf_model_set_mode(cmf_model *m, int mode);
int cmf_model_get_mode(const cmf_model *m);

/* Per-category device timers.  When enabled, every launch of the next forwards is bracketed by CUDA events on
 * the launching stream; cmf_model_read_profile() waits for the last forward and returns, per category
 * (cmf_model_profile_categories() of them, named by cmf_model_profile_name()), the summed device time in ms,
 * the launch count and the algorithmic work (FLOPs for gemm_* categories, 0 otherwise) of that forward. */
int cmf_model_set_profiling(cmf_model *m, int enable);
int cmf_model_profile_categories(void);
const char *cmf_model_profile_name(int cat);
int cmf_model_read_profile(cmf_model *m, float *ms, int *launches, double *work);

/* Debug taps (device pointers into the workspace of the last forward; NULL if not produced):
 * "E" (B,N,800) = [f1 256 | cor 512 | ft 3 | zero pad 29] | "f2" (B,N,256) | "g1","g2","gp" (B,256) | "prop" (B,N,256)
 * | "flow" (B,3,N) | "bq1","bq2" (B,N,60) int32 | "knn12","knn11" (B,N,8) int32 | "P" (B,N,2048) | "cost1","u1","u2" (B,N,512). */
const void *cmf_model_tap(const cmf_model *m, const char *name);

/* Test doorway: Out[c][m] = act(sum_k W[m][k] X[c][k] + bias[m]) through the tcgen05 3xTF32 GEMM alone.
 * scratch_tiles: cmf_test_tc_tiled_floats(M,K) floats of device scratch; ldx must cover K rounded up to 32 (zero padded). */
int cmf_test_tc_gemm(int M, int K, long long cols, const float *W, int ldw, const float *X, int ldx,
                     const float *bias, int act, float *Out, int ldo, float *scratch_tiles, void *stream);
size_t cmf_test_tc_tiled_floats(int M, int K);
/* Same with the operand format chosen: fmt 0 = 3xTF32, 1 = 3xFP16 (power-of-two scaled).  For fmt 1, amax_in (device, one float per group of
 * cols_per_pair consecutive rows of X: an upper bound on |X| in that group) selects the per-group activation scale (NULL = unscaled);
 * amax_out (device, uint32 bit patterns, zeroed by the caller) receives max|Out| per group, or NULL. */
int cmf_test_tc_gemm_fmt(int fmt, int M, int K, long long cols, const float *W, int ldw, const float *X, int ldx,
                         const float *bias, int act, float *Out, int ldo, float *scratch_tiles,
                         int cols_per_pair, const float *amax_in, unsigned int *amax_out, void *stream);
/* Instrumented runs of cmf_test_tc_gemm: device buffer long long[grid][8] receiving per-role barrier wait cycles (NULL = off). */
void cmf_test_tc_set_dbg(long long *dbg);

/* ------------------------------------------------------------------------------------------------
 * Part 4 -- evaluation metrics of the reference's eval loop as device reductions
 * (utils/eval_util.py:42-117, utils/odometry_util.py:34-142; callers main_util.py:175-193).
 * Each call ADDS one batch's sums into a caller-zeroed device array of doubles; cmflow_b200/eval_util.py
 * turns sums into the reference's dictionaries (and all-reduces them across ranks first when sharded).
 * ---------------------------------------------------------------------------------------------- */

/* eval_scene_flow: pc (B,3,N), pred / labels (B,N,3), mask (B,N) float (1 = static).  sums11 += {points, sum error, #accs, #accr,
 * sum re_error, sum re_error[mask==0], #(mask==0), sum re_error[mask==1], #(mask==1), #sas, #ras}; radar resolution as args.radar_res. */
int cmf_eval_scene_flow_sums(int b, int n, const float *pc, const float *pred, const float *labels, const float *mask,
                             double r_res, double theta_res, double phi_res, double *sums11, void *stream);
/* eval_motion_seg: counts4 += {tp, tn, fp, fn} over `total` points (pre, gt float 0/1). */
int cmf_eval_motion_seg_counts(long long total, const float *pre, const float *gt, double *counts4, void *stream);
/* eval_trans_RPE: sums3 += {pairs, sum |translation of gt^-1 pred|, sum rotation angle of gt^-1 pred in degrees}; (B,4,4) each. */
int cmf_eval_rpe_sums(int b, const float *gt_trans, const float *pred_trans, double *sums3, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* CMFLOW_B200_H */
