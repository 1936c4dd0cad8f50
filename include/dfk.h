/*
 * dfk.h -- C ABI of libdfk.so: B200-native (sm_100a) replacement for the dense
 * alignment hot path of DeepFactors' libdf_cuda.so (sources/cuda).
 *
 * The reference exports C++ class templates over Sophus/Eigen/VisionCore types
 * (not a C ABI); this header is the plain-pointer core a maintainer binds from
 * thin wrappers (see INTEGRATION.md and the headers under include/df/ for the C++ facade that
 * reproduces df::SfmAligner / df::SE3Aligner on top of it).  Every entry point
 * cites the reference interface it replaces; citations are file:line into
 * jczarnowski/DeepFactors @ bffc78a.
 *
 * Conventions
 *   pose      float[7] in Sophus::SE3f::data() order: unit quaternion (x,y,z,w),
 *             translation (x,y,z).
 *   DfkImage  pitched 2-D device view, the stand-in for vc::Image2DView /
 *             vc::Buffer2DView: element (x,y) at (char*)ptr + y*pitch_bytes +
 *             x*elem_size.  `width` counts PIXELS for every kind of image:
 *               scalar images (img, dpt, std, valid, prx_orig): 1 float / pixel
 *               grad1: 2 floats / pixel (gx,gy), Eigen::Matrix<float,1,2>
 *               prx_jac: code_size contiguous floats / pixel
 *                        (sources/core/mapping/keyframe.h:52, dense_sfm.h:150;
 *                        the reference views it as a (W*CS) x H float image)
 *   results   JtJ is the packed upper triangle, row major: (i,j), i<=j, at
 *             i*NP - i*(i-1)/2 + (j-i); column order [pose0 t(3) w(3) | pose1
 *             t(3) w(3) | code(C)] (dense_sfm.h:163-177).  Jtr is NOT negated,
 *             residual is the raw sum of squared weighted residuals
 *             (photometric_factor.cpp:105-106,275-282 do the sign flip/rescale).
 *   errors    every call returns a DfkStatus; dfk_last_error(handle) returns the
 *             message the reference would have thrown (launch_utils.h:26-32
 *             vc::CUDAException, cu_sfmaligner.cpp:171-173 std::runtime_error).
 *   threads   a handle is not re-entrant; different handles are independent
 *             (parameters are kernel arguments, not a process-global __constant__
 *             as in cu_sfmaligner.cpp:34,111).
 */
#ifndef DFK_H_
#define DFK_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DFK_VERSION 104

typedef enum {
  DFK_OK = 0,
  DFK_ERR_INVALID_ARG = 1, /* glog CHECK failures of the reference (cu_sfmaligner.cpp:190-191) */
  DFK_ERR_CUDA = 2,        /* vc::CUDAException (launch_utils.h:26-32) */
  DFK_ERR_UNSUPPORTED = 3, /* code size / layout this build has no kernel for */
  DFK_ERR_NOMEM = 4
} DfkStatus;

typedef struct DfkContext* DfkHandle;

/* vc::Image2DView<T, TargetDeviceCUDA> stand-in (VisionCore, not in tree; call sites dense_sfm.h:141-147) */
typedef struct {
  void* ptr;
  size_t pitch_bytes;
  uint32_t width;  /* pixels */
  uint32_t height; /* rows   */
} DfkImage;

/* df::PinholeCamera<float> (sources/common/algorithm/pinhole_camera.h:43, _impl.h:30-31) */
typedef struct {
  float fx, fy, u0, v0;
  float width, height;
} DfkCamera;

/* df::DenseSfmParams (sources/common/algorithm/dense_sfm.h:36-43), same defaults */
typedef struct {
  float huber_delta; /* 0.1  */
  float ocl_th;      /* 1000, unused by the reference */
  float avg_dpt;     /* 2.0  */
  float min_dpt;     /* 0.0  */
  int32_t valid_border; /* 2 */
} DfkDenseSfmParams;

/* df::SfmAlignerParams (sources/cuda/cu_sfmaligner.h:41-48).  The launch-shape members are
 * accepted for source compatibility and validated like the reference (threads % 32 == 0,
 * blocks <= 1024; cu_sfmaligner.cpp:187-203) but the kernels size their own grids. */
typedef struct {
  DfkDenseSfmParams sfmparams;
  int32_t step_threads; /* 32  */
  int32_t step_blocks;  /* 11  */
  int32_t eval_threads; /* 224 */
  int32_t eval_blocks;  /* 66  */
} DfkSfmAlignerParams;

/* How the (6+C)x(6+C) Gauss-Newton Gram is accumulated.
 *   DFK_GRAM_FP32     CUDA-core FFMA, fp32 products and sums (any supported code size)
 *   DFK_GRAM_TF32X3   tcgen05 tensor cores, split-precision tf32 (hi*hi + lo*hi + hi*lo),
 *                     fp32 accumulate in TMEM; code size >= 32 only
 *   DFK_GRAM_AUTO     tensor cores where available for the code size, else FP32 */
typedef enum { DFK_GRAM_AUTO = 0, DFK_GRAM_FP32 = 1, DFK_GRAM_TF32X3 = 2 } DfkGramMode;

/* ------------------------------------------------------------------ lifetime / config */

/* Replaces SfmAligner::SfmAligner / SE3Aligner::SE3Aligner (cu_sfmaligner.cpp:102-114,
 * cu_se3aligner.cpp:119-120): owns the reduction scratch (bscratch_) and a stream.
 * device < 0 => current device. */
DfkStatus dfk_create(int device, DfkHandle* out);
DfkStatus dfk_destroy(DfkHandle h);
/* cudaStream_t to launch on; NULL is the legacy default stream (the one the reference uses,
 * with cudaDeviceSynchronize after every launch: launch_utils.h:28).  A new handle launches on a
 * private non-blocking stream; dfk_use_own_stream returns to it. */
DfkStatus dfk_set_stream(DfkHandle h, void* cuda_stream);
DfkStatus dfk_use_own_stream(DfkHandle h);
/* The RunStep kernels are persistent: one launch fills every SM for its whole duration, so a kernel that arrives on
 * another stream meanwhile (the collective of the previous step on a multi-GPU window, bench.py) finds no room, starts
 * when the first CTAs retire and then holds SMs the NEXT step's grid was sized for.  num_sms > 0 sizes the grids for
 * that many SMs and leaves the rest to the concurrent kernel; 0 = all SMs (default).  Results do not depend on it
 * beyond the summation order of the per-CTA partials (deterministic for a given limit).  No reference counterpart: the
 * reference synchronises the device after every launch (launch_utils.h:28). */
DfkStatus dfk_set_sm_limit(DfkHandle h, int num_sms);
void* dfk_get_stream(DfkHandle h);
DfkStatus dfk_synchronize(DfkHandle h);
const char* dfk_last_error(DfkHandle h);
const char* dfk_status_string(DfkStatus s);
int dfk_version(void);
/* 1 if this build has a RunStep kernel for the code size (reference: only 32, cu_sfmaligner.cpp:209) */
int dfk_sfm_supports_code_size(int code_size);

/* Measurement hooks (no reference equivalent; the reference times with std::clock around
 * synchronous calls, sources/common/timing.h:28-45).  With profiling on, every launch of the
 * dominant kernel (the per-tile warp+Gram kernel of RunStep) is bracketed by CUDA events on the
 * launching stream.  dfk_get_profile synchronizes the stream, returns the summed device time
 * of those launches and their count plus the number of ALL kernels this handle launched since the
 * last call, and resets the counters. */
DfkStatus dfk_set_profiling(DfkHandle h, int enabled);
DfkStatus dfk_get_profile(DfkHandle h, double* main_kernel_ms, uint64_t* main_kernel_launches,
                          uint64_t* total_kernel_launches);

/* SfmAligner ctor params + SetEvalThreadsBlocks/SetStepThreadsBlocks (cu_sfmaligner.cpp:187-203) */
DfkStatus dfk_sfm_set_params(DfkHandle h, const DfkSfmAlignerParams* p);
DfkStatus dfk_sfm_get_params(DfkHandle h, DfkSfmAlignerParams* p);
DfkStatus dfk_sfm_set_gram_mode(DfkHandle h, DfkGramMode m);
/* SE3Aligner::SetHuberDelta (cu_se3aligner.h:72), default 0.1 (:85) */
DfkStatus dfk_se3_set_huber_delta(DfkHandle h, float v);

/* ------------------------------------------------------------------ SfmAligner */

/* SfmAligner<float,CS>::RunStep (cu_sfmaligner.h:76-86, cu_sfmaligner.cpp:149-185).
 * code0/std0 are accepted and ignored exactly as the reference kernel ignores them
 * (dense_sfm.h:56-67,133-201); both may be NULL.  valid0 is in/out: set to 1 where a
 * correspondence is valid, never cleared (dense_sfm.h:161).  Synchronous; results land in
 * host memory: JtJ[NP(NP+1)/2], Jtr[NP], NP = 12 + code_size. */
DfkStatus dfk_sfm_run_step(DfkHandle h, const float pose0[7], const float pose1[7],
                           const float* code0, int code_size, const DfkCamera* cam,
                           const DfkImage* img0, const DfkImage* img1, const DfkImage* dpt0,
                           const DfkImage* std0, const DfkImage* valid0, const DfkImage* prx0_jac,
                           const DfkImage* grad1,
                           float* JtJ, float* Jtr, float* residual, uint64_t* inliers);

/* SfmAligner<float,CS>::EvaluateError (cu_sfmaligner.h:67-74, cu_sfmaligner.cpp:120-147):
 * border 1 / min_dpt 0 (dense_sfm.h:91), Huber-weighted squared error + inlier count.
 * std0/grad1 feed only the dead uncertainty weight (dense_sfm.h:100-107); may be NULL. */
DfkStatus dfk_sfm_evaluate_error(DfkHandle h, const float pose0[7], const float pose1[7],
                                 const DfkCamera* cam, const DfkImage* img0, const DfkImage* img1,
                                 const DfkImage* dpt0, const DfkImage* std0, const DfkImage* grad1,
                                 float* residual, uint64_t* inliers);

/* One (keyframe, frame, pyramid level) evaluation of a batch: the unit PhotometricFactor
 * creates per pair and level (sources/core/mapping/df_work.cpp:211-225). */
typedef struct {
  float pose0[7];
  float pose1[7];
  DfkCamera cam;
  DfkImage img0, img1, dpt0, valid0, prx0_jac, grad1;
  /* Optional fused depth decode (PhotometricFactor::UpdateDepthMaps + RunAlignmentStep in one pass,
   * photometric_factor.cpp:229,331-341): when `code` is not NULL, every pixel's depth is decoded first,
   *   dpt0(x,y) = avg_dpt / (prx_orig(x,y) + prx0_jac(x,y,:) . code) - avg_dpt      (warping.h:30-69),
   * written to dpt0 (an OUTPUT in this mode, exactly what dfk_update_depth writes, bit for bit) and used for the warp,
   * so the code Jacobian is read from HBM once instead of twice.  `code` is a HOST pointer to code_size floats. */
  DfkImage prx_orig;
  const float* code;
} DfkSfmWorkItem;

/* floats per device result record for code size C: [JtJ packed | Jtr | residual | inliers(u32 bits)] */
#define DFK_SFM_NP(C) (12 + (C))
#define DFK_SFM_RECORD_FLOATS(C) (DFK_SFM_NP(C) * (DFK_SFM_NP(C) + 1) / 2 + DFK_SFM_NP(C) + 2)

/* Batched RunStep: n work items in one persistent launch (the reference evaluates them one
 * call at a time from ISAM2, mapper.cpp:518-519).  `items` is a HOST array; `records_dev`
 * is DEVICE memory, n * DFK_SFM_RECORD_FLOATS(code_size) floats.  Asynchronous on the
 * handle's stream (no host sync, no D2H). */
DfkStatus dfk_sfm_run_step_batch(DfkHandle h, const DfkSfmWorkItem* items, int n, int code_size,
                                 float* records_dev);
/* Same, then copies the records to host memory and synchronizes. */
DfkStatus dfk_sfm_run_step_batch_host(DfkHandle h, const DfkSfmWorkItem* items, int n, int code_size,
                                      float* records_host);

/* ------------------------------------------------------------------ streaming evaluation from HOST memory
 *
 * The reference's inputs live in host/device mirrored pyramids that are uploaded lazily, one synchronous copy at a
 * time, on first GPU use (sources/cuda/synced_pyramid.h:118-126,178-198).  This is the same hand-over as ONE pipelined
 * call: dfk_sfm_stream_submit takes work items whose image views point at HOST memory (pinned for full speed), uploads
 * them on a copy stream into one of `depth` device slots, evaluates them (dfk_sfm_run_step_batch) on the handle's stream
 * as soon as the upload has landed and sends the result records back -- all asynchronous, so the upload of submission
 * k+1 overlaps the evaluation of submission k.  dfk_sfm_stream_wait blocks until the records of a ticket are in host
 * memory.  Tickets must be waited for in order; at most `depth` submissions may be outstanding.
 * valid0 of a streamed item is device scratch (the mask is neither uploaded nor returned); with the fused depth decode
 * (code != NULL) prx_orig is uploaded instead of dpt0 and the decoded depth stays on the device.
 */
typedef struct DfkSfmStream DfkSfmStream;
DfkStatus dfk_sfm_stream_create(DfkHandle h, int code_size, int max_items, size_t max_bytes_per_submit, int depth,
                                DfkSfmStream** out);
DfkStatus dfk_sfm_stream_destroy(DfkHandle h, DfkSfmStream* s);
DfkStatus dfk_sfm_stream_submit(DfkHandle h, DfkSfmStream* s, const DfkSfmWorkItem* host_items, int n,
                                uint64_t* ticket);
DfkStatus dfk_sfm_stream_wait(DfkHandle h, DfkSfmStream* s, uint64_t ticket, float* records_host);

/* ------------------------------------------------------------------ keyframe window (block-sparse normal equations)
 *
 * What the factor graph does with the RunStep results of a window of keyframes: every (pair, level) result is one
 * PhotometricFactor (sources/core/mapping/df_work.cpp:211-225) whose linearize() slices the (12+C)^2 Hessian into the
 * blocks G11 G12 G13 G22 G23 G33 / g1 g2 g3 of a HessianFactor over (pose0, pose1, code0) with g = -Jtr and the
 * residual rescaled to res / inliers * W * H (sources/core/gtsam/photometric_factor.cpp:105-161, 275-282); the solver
 * then adds the factors of the window into one system.  dfk_window_assemble does that sum ON THE DEVICE, straight from
 * the record buffer of dfk_sfm_run_step_batch, into a packed block-sparse buffer -- the one buffer a multi-GPU
 * Gauss-Newton step all-reduces (pairs shard across GPUs, every rank assembles its own pairs into the same layout).
 *
 * Variables: keyframe k owns [pose_k (6) | code_k (C)], B = 6 + C.  Buffer layout (fp32):
 *   K diagonal blocks  B x B, row-major, full symmetric
 *   K gradients        B            (g = -sum Jtr)
 *   P coupling blocks  B x 6, row-major: rows = [pose0 | code0] of the pair's keyframe k0, columns = pose1 of its k1
 *   2 scalars          f = sum of rescaled residuals over items with overlap, total inliers (as a float)
 * Deterministic: every output element is summed by one thread in item order (a gather, no float atomics).
 */
typedef struct DfkWindow DfkWindow;
typedef struct {
  int32_t num_keyframes;
  int32_t num_pairs;
  int32_t num_items;       /* records per evaluation: one per (pair, level) */
  int32_t code_size;
  const int32_t* pair_k0;  /* [num_pairs] keyframe (pose0 / code0) of every pair      (HOST arrays, copied) */
  const int32_t* pair_k1;  /* [num_pairs] frame (pose1)                                                      */
  const int32_t* item_pair;   /* [num_items] pair of every record                                            */
  const int32_t* item_width;  /* [num_items] level size, for the residual rescale                            */
  const int32_t* item_height;
} DfkWindowDesc;
DfkStatus dfk_window_create(DfkHandle h, const DfkWindowDesc* desc, DfkWindow** out);
DfkStatus dfk_window_destroy(DfkHandle h, DfkWindow* w);
/* floats of the block-sparse buffer: K*(B*B + B) + P*6*B + 2 */
size_t dfk_window_floats(const DfkWindow* w);
/* records_dev: num_items records as written by dfk_sfm_run_step_batch (DEVICE).  window_dev: dfk_window_floats()
 * floats (DEVICE), fully overwritten.  Asynchronous on the handle's stream, one launch. */
DfkStatus dfk_window_assemble(DfkHandle h, const DfkWindow* w, const float* records_dev, float* window_dev);

/* ------------------------------------------------------------------ SE3Aligner */

/* SE3Aligner<float>::RunStep (cu_se3aligner.h:65-70, cu_se3aligner.cpp:153-176):
 * JtJ[21] packed upper 6x6, Jtr[6]. */
DfkStatus dfk_se3_run_step(DfkHandle h, const float se3[7], const DfkCamera* cam,
                           const DfkImage* img0, const DfkImage* img1, const DfkImage* dpt0,
                           const DfkImage* grad1,
                           float* JtJ, float* Jtr, float* residual, uint64_t* inliers);

/* One pyramid level of a tracking problem: keyframe image/depth (img0, dpt0), live frame image/gradient (img1,
 * grad1), the level's camera and the Gauss-Newton iteration count (TrackerConfig::iterations_per_level,
 * core/system/camera_tracker.h:45-50). */
typedef struct DfkTrackLevel {
  DfkCamera cam;
  DfkImage img0, img1, dpt0, grad1;
  int iterations;
} DfkTrackLevel;

/* CameraTracker::TrackFrame (core/system/camera_tracker.cpp:42-69): coarse-to-fine Gauss-Newton on pose_ck.
 * levels[0] is the finest level; iteration runs from levels[num_levels-1] down to levels[0].  Per iteration the
 * reference does SE3Aligner::RunStep + cudaDeviceSynchronize + a 120-byte D2H + a host 6x6 LDLT + retraction
 * (update = -JtJ.ldlt().solve(Jtr); t += update.head<3>(); so3 = exp(update.tail<3>()) * so3); here every iteration
 * is ONE launch whose last block solves the 6x6 system and retracts the pose in device memory, all iterations are
 * enqueued back to back and there is a single read-back at the end.
 *   pose_ck         in/out, (qx,qy,qz,qw,tx,ty,tz)
 *   inlier_fraction inliers / area and error = residual / inliers of the last evaluated system (the reference
 *   error           records them on the last iteration of level 0, :65-69); error = +inf when inliers == 0
 *   last_system     optional, 29 floats [JtJ packed upper 21 | Jtr 6 | residual | inliers (u32 bits)]
 *   history         optional, history_capacity x 36 floats: per iteration the 29 floats above + the pose (7) they
 *                   were evaluated at
 * An iteration whose system is not positive definite (e.g. zero inliers) leaves the pose untouched. */
DfkStatus dfk_se3_track(DfkHandle h, float pose_ck[7], const DfkTrackLevel* levels, int num_levels,
                        float* inlier_fraction, float* error, float* last_system, float* history,
                        int history_capacity);

/* SE3Aligner<float>::Warp (cu_se3aligner.h:58-63, cu_se3aligner.cpp:125-151): renders img1
 * into frame 0 (img2, 0 where invalid); residual = SIGNED sum(img0 - sampled) (:106). */
DfkStatus dfk_se3_warp(DfkHandle h, const float se3[7], const DfkCamera* cam,
                       const DfkImage* img0, const DfkImage* img1, const DfkImage* dpt0,
                       const DfkImage* img2, float* residual, uint64_t* inliers);

/* ------------------------------------------------------------------ DepthAligner */

/* DepthAligner<float,CS>::RunStep (sources/cuda/cu_depthaligner.h:46-49, cu_depthaligner.cpp:32-113): aligns the depth
 * decoded from `code` (HOST, code_size floats) to target_dpt; every pixel counts.  JtJ[CS(CS+1)/2] packed upper,
 * Jtr[CS], residual = sum diff^2, inliers = W*H.  The reference hard-codes avg_dpt = 2 in this kernel (:44); here it
 * is the handle's DenseSfmParams::avg_dpt (same default).  Synchronous. */
DfkStatus dfk_depth_run_step(DfkHandle h, const float* code, int code_size, const DfkImage* target_dpt,
                             const DfkImage* prx_orig, const DfkImage* prx_jac,
                             float* JtJ, float* Jtr, float* residual, uint64_t* inliers);

/* ------------------------------------------------------------------ sparse keypoint factor */

/* ReprojectionFactor::linearize (sources/core/gtsam/reprojection_factor.cpp:157-269): the Jacobian rows of a keypoint
 * reprojection factor, gathered on the device from the keyframe's level-0 proximity / code-Jacobian buffers instead of
 * mirroring the whole pyramid to the host (kf_->pyr_jac.GetCpuLevel(0), :193).
 *   query_xy / train_xy   HOST, 2 floats per match: matched keypoints in the keyframe / in the frame (:183-186)
 *   rows                  HOST out, (2 * num_matches) x (13 + code_size), row-major:
 *                         [dErr/dPose0 (6) | dErr/dPose1 (6) | dErr/dCode0 (C) | b (1)], already multiplied by the
 *                         Cauchy weight (m_estimators.h:43-48, parameter `cauchy_delta` = the factor's huber_delta_) and
 *                         divided by sigma -- the blocks of gtsam::JacobianFactor(keys, Ab) (:255-268); zero rows for a
 *                         match whose point falls behind the camera (:204-212)
 *   total_err             out, sum of squared UNWEIGHTED reprojection errors (total_err_, :242,258)
 * avg_dpt is hard-coded to 2 there (:171); here it is the handle's DenseSfmParams::avg_dpt.  Synchronous. */
DfkStatus dfk_reprojection_linearize(DfkHandle h, const float pose0[7], const float pose1[7], const float* code0,
                                     int code_size, const DfkCamera* cam, const DfkImage* prx_orig,
                                     const DfkImage* prx_jac, int num_matches, const float* query_xy,
                                     const float* train_xy, float cauchy_delta, float sigma, float* rows,
                                     float* total_err);

/* SparseGeometricFactor::linearize (sources/core/gtsam/sparse_geometric_factor.cpp:157-271): the Jacobian rows of the
 * sparse depth-consistency factor between two keyframes, evaluated on the device from the keyframes' level-0 proximity /
 * code-Jacobian buffers and keyframe 1's depth gradient instead of host mirrors of all five (:181-183, :207-209, :220).
 *   points_xy    HOST, 2 ints per point: the sampled pixels of keyframe 0 (UniformSampler, uniform_sampler.h:28-32)
 *   dpt_grad1    kf1->dpt_grad: SobelGradients of keyframe 1's level-0 depth (mapper.cpp:998-1000), 2 floats per pixel
 *   rows         HOST out, num_points x (13 + 2 * code_size), row-major:
 *                [dErr/dPose0 (6) | dErr/dPose1 (6) | dErr/dCode0 (C) | dErr/dCode1 (C) | b], already multiplied by the
 *                Huber weight (DenseSfm_RobustLoss, dense_sfm.h:47-50) -- the blocks of gtsam::JacobianFactor(keys, Ab)
 *                (:260-270); zero rows for points whose correspondence is invalid (:190-198)
 *   num_valid    out (may be NULL): rows that are not all zero
 * avg_dpt is hard-coded to 2 there (:166); here it is the handle's DenseSfmParams::avg_dpt.  Synchronous. */
DfkStatus dfk_sparse_geometric_linearize(DfkHandle h, const float pose0[7], const float pose1[7], const float* code0,
                                         const float* code1, int code_size, const DfkCamera* cam, const DfkImage* prx0_orig,
                                         const DfkImage* prx0_jac, const DfkImage* prx1_orig, const DfkImage* prx1_jac,
                                         const DfkImage* dpt_grad1, int num_points, const int* points_xy, float huber_delta,
                                         float* rows, int* num_valid);

/* ------------------------------------------------------------------ cu_image_proc free functions */

/* df::UpdateDepth (cu_image_proc.h:41-44, cu_image_proc.cpp:248-277):
 * dpt = avg/(prx_orig + prx_jac . code) - avg.  code is HOST memory. Asynchronous. */
DfkStatus dfk_update_depth(DfkHandle h, const float* code, int code_size, const DfkImage* prx_orig,
                           const DfkImage* prx_jac, float avg_dpt, const DfkImage* dpt_out);
/* df::SobelGradients (cu_image_proc.h:27-29, cu_image_proc.cpp:57-113). Asynchronous. */
DfkStatus dfk_sobel_gradients(DfkHandle h, const DfkImage* img, const DfkImage* grad);
/* df::GaussianBlurDown (cu_image_proc.h:31-33, cu_image_proc.cpp:134-184). Asynchronous. */
DfkStatus dfk_gaussian_blur_down(DfkHandle h, const DfkImage* in, const DfkImage* out);
/* The image half of Frame::FillPyramids / BuildKeyframe (core/mapping/frame.h:80-94, mapper.cpp:935-949):
 * imgs[0] is the input; imgs[l] = GaussianBlurDown(imgs[l-1]) and, when grads != NULL, grads[l] =
 * SobelGradients(imgs[l]) for every level.  2*levels-1 launches enqueued back to back on the handle's stream, no host
 * synchronization (the reference synchronizes after each and re-uploads the kernel taps with cudaMemcpyToSymbol,
 * cu_image_proc.cpp:103-112,174-183). */
DfkStatus dfk_build_image_pyramid(DfkHandle h, const DfkImage* imgs, const DfkImage* grads, int levels);
/* df::SquaredError (cu_image_proc.h:35-39, cu_image_proc.cpp:190-242). Synchronous. */
DfkStatus dfk_squared_error(DfkHandle h, const DfkImage* a, const DfkImage* b, float* out);

#ifdef __cplusplus
}
#endif
#endif /* DFK_H_ */
