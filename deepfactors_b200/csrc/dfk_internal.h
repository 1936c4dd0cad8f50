// dfk_internal.h -- shared between the translation units of libdfk.so (not installed).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "dfk.h"

namespace dfk {

// ----------------------------------------------------------------------------------------------
// Device-side description of one (keyframe, frame, level) evaluation.  Built on the host by
// dfk_api.cu from a DfkSfmWorkItem: the relative pose and its two 6x6 Jacobians are host work in
// the reference too (cu_sfmaligner.cpp:164-166).
// ----------------------------------------------------------------------------------------------
struct SfmItemDev {
  // pose_10 = pose1^-1 * pose0 : quaternion (x,y,z,w), translation, and the same rotation as a
  // row-major 3x3 (used only for derivative terms, never for the validity chain)
  float q[4];
  float t[3];
  float R[9];
  // camera (pinhole_camera.h:43) + validity window (pinhole_camera_impl.h:102-108)
  float fx, fy, u0, v0;
  float border;  // (float)valid_border
  float ulim;    // width  - border  (float arithmetic as in the reference)
  float vlim;    // height - border
  float min_dpt, avg_dpt, huber_delta;
  // buffers; pitches in floats
  const float* img0;
  const float* img1;
  const float* dpt0;
  float* valid0;
  const float* jac;
  const float* grad1;
  uint32_t img0_pitch, img1_pitch, dpt0_pitch, valid0_pitch, jac_pitch, grad1_pitch;
  uint32_t width, height;
  uint32_t num_pixels;
  // tiling
  uint32_t tile_begin;  // first global tile index of this item
  uint32_t num_tiles;
  uint32_t perm_mul;    // tile k of the item is processed as (k * perm_mul) % num_tiles  (k * perm_mul < 2^32)
  uint32_t mag_tiles;   // floor(2^32 / num_tiles): division by multiply-high + one correction step
  uint32_t mag_width;   // floor(2^32 / width)
  // partial-sum bookkeeping: CTA c (first_cta <= c < first_cta+num_ctas) writes slot
  // partial_begin + (c - first_cta)
  uint32_t first_cta, num_ctas, partial_begin;
  uint32_t flags;  // bit0: bulk-copy (TMA) eligible, bit1: grad1 rows are 8-byte aligned
  // fused depth decode (ITEM_FLAG_FUSED_DEPTH): `dpt0` then points at prx_orig (what the tile loader stages), the decoded
  // depth is written to dpt_out; code = code_size floats in device scratch
  float* dpt_out;
  uint32_t dpt_out_pitch;
  const float* code;
  // normalised ray table of the item's camera level (device memory cached by the handle): xn[0..width), then yn[0..height)
  // (tensor-core kernel)
  const float* ray_tab;
  // relative-pose Jacobians (warping.h:120-134), row-major 6x6; used by the finalize kernel
  float P0[36];
  float P1[36];
};

enum : uint32_t { ITEM_FLAG_BULK = 1u, ITEM_FLAG_GRAD_ALIGNED = 2u, ITEM_FLAG_FUSED_DEPTH = 4u };

// Geometry of the fp32 Gram kernel, shared by host planning code and the kernel.
template <int C>
struct SfmCfg {
  static constexpr int NF = C + 7;               // features: code(C) | a(6) | r(1)
  static constexpr int NFP = (NF + 7) & ~7;      // padded to 8
  static constexpr int NB = NFP / 8;             // 8x8 blocks per side
  static constexpr int NBLK = NB * (NB + 1) / 2;  // upper-triangular blocks
  static constexpr int PARTIAL_FLOATS = NFP * NFP + 8;  // G (row major NFP x NFP) | inliers(u32) | pad
};

constexpr int kTilePixels = 256;    // fp32 kernel
constexpr int kTcTilePixels = 128;  // tensor-core kernel
// tensor-core partial: rows = TMEM lanes that carry data (32 code-h, 32 code-l, 7 + 1 pose-h, 7 + 1 pose-l),
// columns = B features (32 code, 7 pose/residual, 1 pad)
constexpr int kTcRows = 80;
constexpr int kTcCols = 40;
// stored column-major with the row dimension padded to the 96 TMEM lanes of the three operand warps, so that a
// warp's 32 lanes (rows) touch 32 consecutive floats per column: coalesced st / red
constexpr int kTcRowsPad = 96;
constexpr int kTcPartialFloats = kTcRowsPad * kTcCols + 8;

struct SfmLaunchPlan {
  int num_items = 0;
  int num_tiles = 0;
  int num_ctas = 0;
  int num_partials = 0;
  int max_ctas_per_item = 0;
};

// dfk_sfm_fp32.cu
cudaError_t launch_sfm_fp32(int code_size, const SfmItemDev* items_dev, const SfmLaunchPlan& plan,
                            float* partials_dev, float* records_dev, cudaStream_t stream,
                            cudaEvent_t ev_start = nullptr, cudaEvent_t ev_stop = nullptr);
cudaError_t launch_sfm_finalize(int code_size, bool tc, const SfmItemDev* items_dev, int num_items,
                                const float* partials_dev, float* records_dev, cudaStream_t stream);
// dfk_sfm_wide.cu : C = 64 / 128 (thread-owned 8x8 blocks); partial format = the fp32 kernel's
constexpr int sfm_wide_tile_pixels(int code_size) { return code_size >= 128 ? 64 : 128; }
bool sfm_wide_supported(int code_size);
cudaError_t launch_sfm_wide(int code_size, const SfmItemDev* items_dev, const SfmLaunchPlan& plan,
                            float* partials_dev, cudaStream_t stream, cudaEvent_t ev_start = nullptr,
                            cudaEvent_t ev_stop = nullptr);
// dfk_sfm_tc.cu
bool sfm_tc_supported(int code_size);
cudaError_t launch_sfm_tc(const SfmItemDev* items_dev, const SfmLaunchPlan& plan, bool build_ray_tables,
                          float* partials_dev, cudaStream_t stream, cudaEvent_t ev_start = nullptr,
                          cudaEvent_t ev_stop = nullptr);
size_t sfm_partial_floats(int code_size);
// resident CTAs per SM of the fp32 kernel: at C = 8 a CTA is 11 warps and ~60 KB of shared memory, two fit (the front-end
// is latency-bound, so the second CTA nearly doubles the throughput); from C = 16 on the register budget allows one
constexpr int sfm_fp32_ctas_per_sm(int code_size) { return code_size <= 8 ? 2 : 1; }
bool sfm_fp32_supported(int code_size);
int sfm_max_ctas();  // grid size of the persistent kernel on the current device

// dfk_simple.cu : single-pass reductions with a last-block finalize
struct PixelCam {
  float q[4], t[3];
  float fx, fy, u0, v0, border, ulim, vlim, min_dpt;
};
struct View {
  const float* ptr;
  uint32_t pitch;  // floats
};
cudaError_t launch_se3_step(const PixelCam& pc, float huber_delta, int width, int height, View img0, View img1,
                            View dpt0, View grad1, bool grad_aligned, float* scratch, unsigned int* counter,
                            float* out_dev /*29 floats: 21 JtJ, 6 Jtr, res, inliers bits*/, cudaStream_t s,
                            float* pose_dev = nullptr /*tracking mode: pose read from / updated in device memory*/,
                            float* history_dev = nullptr /*36 floats: the 29 above + the pose they were evaluated at*/);
cudaError_t launch_eval_error(const PixelCam& pc, float huber_delta, int width, int height, View img0, View img1,
                              View dpt0, float* scratch, unsigned int* counter, float* out_dev /*2*/, cudaStream_t s);
cudaError_t launch_warp(const PixelCam& pc, int width, int height, View img0, View img1, View dpt0, float* img2,
                        uint32_t img2_pitch, float* scratch, unsigned int* counter, float* out_dev /*2*/,
                        cudaStream_t s);
cudaError_t launch_update_depth(const float* code_dev, int code_size, int width, int height, View prx_orig, View jac,
                                float avg_dpt, float* dpt, uint32_t dpt_pitch, cudaStream_t s);
cudaError_t launch_sobel(int width, int height, View img, float* grad, uint32_t grad_pitch, cudaStream_t s);
cudaError_t launch_blur_down(int in_w, int in_h, View in, int out_w, int out_h, float* out, uint32_t out_pitch,
                             cudaStream_t s);
cudaError_t launch_squared_error(int width, int height, View a, View b, float* scratch, unsigned int* counter,
                                 float* out_dev, cudaStream_t s);
// dfk_window.cu : block-sparse window assembly (gather over CSR lists built on the host by dfk_window_create)
struct WindowDev {
  int num_keyframes, num_pairs, num_items, code_size;
  const int* kf0_ptr;     // [K+1] items whose keyframe (k0) is k ...
  const int* kf0_items;   // ... in item order
  const int* kf1_ptr;     // [K+1] items whose frame (k1) is k
  const int* kf1_items;
  const int* pair_ptr;    // [P+1] items of pair p (its levels)
  const int* pair_items;
  const float* item_area; // [n] W * H of the item's level
};
cudaError_t launch_window_assemble(const WindowDev& w, const float* records_dev, float* out_dev, cudaStream_t stream);

// dfk_depth.cu : DepthAligner::RunStep
size_t depth_partial_floats(int code_size);
cudaError_t launch_depth_step(const float* code_dev, int code_size, int width, int height, View tgt, View prx_orig,
                              View jac, float avg_dpt, float* scratch /*blocks * depth_partial_floats*/,
                              unsigned int* counter, float* out_dev /*C(C+1)/2 + C + 2*/, int blocks, cudaStream_t s);

// dfk_sparse.cu : ReprojectionFactor::linearize rows
struct SparsePose {
  float q[4], t[3], R[9];      // pose_10 = pose1^-1 * pose0
  float P0[36], P1[36];        // pose10_J_pose0 / pose10_J_pose1, row-major 6x6
  float fx, fy, u0, v0;
};
cudaError_t launch_reprojection_rows(const SparsePose& sp, const float* code_dev, int code_size, View prx_orig, View jac,
                                     int width, int height, int num_matches, const float* query_dev, const float* train_dev,
                                     float cauchy_delta, float sigma, float avg_dpt, float* rows_dev, float* err2_dev,
                                     cudaStream_t s);

cudaError_t launch_sparse_geometric_rows(const SparsePose& sp, float cam_w, float cam_h, const float* code0_dev,
                                         const float* code1_dev, int code_size, View prx0, View jac0, View prx1, View jac1,
                                         View grad1, int width, int height, int num_points, const int* points_dev,
                                         float huber_delta, float avg_dpt, float* rows_dev, cudaStream_t s);

constexpr int kSimpleMaxBlocks = 1024;
constexpr int kSimpleScratchFloats = kSimpleMaxBlocks * 32;

}  // namespace dfk
