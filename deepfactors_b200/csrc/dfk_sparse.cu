// dfk_sparse.cu -- the sparse factors on the device: ReprojectionFactor::linearize
// (sources/core/gtsam/reprojection_factor.cpp:157-269) and SparseGeometricFactor::linearize (second half of the file).
//
// The reference evaluates this sparse keypoint factor on the CPU and, to read the code Jacobian at <= a few thousand
// keypoints, forces a device -> host mirror of the keyframe's WHOLE level-0 code-Jacobian pyramid
// (kf_->pyr_jac.GetCpuLevel(0), :193; 39 MB at 640x480, C = 32).  Here the rows are gathered where the data lives: one
// thread per match reads its C-float Jacobian row and proximity, decodes the depth, warps the keypoint, and writes the
// two rows of the JacobianFactor [ dErr/dPose0 (6) | dErr/dPose1 (6) | dErr/dCode0 (C) | b (1) ]; ~1 MB goes back.
//
// Per match i (query keypoint in the keyframe, train keypoint in the frame):
//   (xi, yi)  = integer pixel of the query (the reference indexes with (int)query.x / implicit size_t conversions, :194-195,
//               and FindCorrespondence takes std::size_t x, y, warping.h:206)
//   dpt0      = DepthFromCode(c0, prx_J_cde, prx_0code, avg_dpt = 2)                       :198, warping.h:52-69
//   corr      = FindCorrespondence(xi, yi, dpt0, cam, pose10, 1, 0, check_bounds = false)  :199-200  (valid <=> Z > 0)
//   invalid   -> zero rows                                                                   :204-212
//   J_cde     = FindCorrespondenceJacobianCode (2 x C)                                       :217-218, warping.h:294-313
//   J_pose10  = FindCorrespondenceJacobianPose (2 x 6);  J_pose0/1 = J_pose10 * pose10_J_pose0/1   :221-229
//   diff      = pix1(train) - corr.pix1 ; err = |diff| ; w = CauchyWeight(err, huber_delta)  :232-239, m_estimators.h:43-48
//   rows *= w ; total_err += err^2 ; rows /= sigma                                           :242-253
#include <cuda_runtime.h>
#include <stdint.h>

#include "dfk_geom.cuh"
#include "dfk_internal.h"

namespace dfk {

namespace {

template <int C>
__global__ void __launch_bounds__(128)
reprojection_rows_kernel(SparsePose sp, const float* __restrict__ code, View prx_orig, View jac, int width, int height,
                         int num_matches, const float2* __restrict__ query, const float2* __restrict__ train,
                         float cauchy_delta, float sigma, float avg_dpt, float* __restrict__ rows, float* __restrict__ err2)
{
  constexpr int RW = 13 + C;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= num_matches) return;
  float* r0 = rows + (size_t)(2 * i) * RW;
  float* r1 = r0 + RW;
  const float2 q = query[i], tr = train[i];
  const int xi = (int)q.x, yi = (int)q.y;
  bool valid = xi >= 0 && yi >= 0 && xi < width && yi < height;  // the reference would read out of bounds
  float dpt0 = 0.f, X = 0.f, Y = 0.f, Z = 0.f, px = 0.f, py = 0.f, pz = 0.f, xn = 0.f, yn = 0.f;
  const float* jr = nullptr;
  if (valid) {
    jr = jac.ptr + (size_t)yi * jac.pitch + (size_t)xi * C;
    float dot = 0.0f;
    for (int k = 0; k < C; ++k) dot += __ldg(jr + k) * code[k];  // (prx_J_cde * code)(0), left to right
    const float prx = __ldg(prx_orig.ptr + (size_t)yi * prx_orig.pitch + xi) + dot;
    dpt0 = avg_dpt / prx - avg_dpt;
    // Reproject + se3 * pt (quaternion rotate as Sophus does)
    xn = ((float)xi - sp.u0) / sp.fx;
    yn = ((float)yi - sp.v0) / sp.fy;
    const float P0 = xn * dpt0, P1 = yn * dpt0, P2 = dpt0;
    float uv0 = sp.q[1] * P2 - sp.q[2] * P1, uv1 = sp.q[2] * P0 - sp.q[0] * P2, uv2 = sp.q[0] * P1 - sp.q[1] * P0;
    uv0 += uv0; uv1 += uv1; uv2 += uv2;
    px = (P0 + sp.q[3] * uv0) + (sp.q[1] * uv2 - sp.q[2] * uv1);
    py = (P1 + sp.q[3] * uv1) + (sp.q[2] * uv0 - sp.q[0] * uv2);
    pz = (P2 + sp.q[3] * uv2) + (sp.q[0] * uv1 - sp.q[1] * uv0);
    X = px + sp.t[0]; Y = py + sp.t[1]; Z = pz + sp.t[2];
    valid = Z > 0.0f;  // depth > min_dpt (0); bounds are not checked (check_bounds = false)
  }
  if (!valid) {
    for (int k = 0; k < RW; ++k) { r0[k] = 0.0f; r1[k] = 0.0f; }
    err2[i] = 0.0f;
    return;
  }
  const float u = sp.fx * X / Z + sp.u0, v = sp.fy * Y / Z + sp.v0;  // Project
  // ProjectPointJacobian
  const float c00 = sp.fx / Z, c02 = -(sp.fx * X) / Z / Z, c11 = sp.fy / Z, c12 = -(sp.fy * Y) / Z / Z;
  // corr_J_pose10 = dCam * [I | -hat(R pt)]
  const float A0[6] = {c00, 0.f, c02, c02 * py, c00 * pz - c02 * px, -(c00 * py)};
  const float A1[6] = {0.f, c11, c12, c12 * py - c11 * pz, -(c12 * px), c11 * px};
  // pix1_J_dpt = dCam * R * (xn, yn, 1);  dpt_J_prx = -avg / prx^2
  const float q0 = sp.R[0] * xn + sp.R[1] * yn + sp.R[2];
  const float q1 = sp.R[3] * xn + sp.R[4] * yn + sp.R[5];
  const float q2 = sp.R[6] * xn + sp.R[7] * yn + sp.R[8];
  const float pr = avg_dpt / (avg_dpt + dpt0);
  const float dJ = -avg_dpt / (pr * pr);
  const float jd0 = (c00 * q0 + c02 * q2) * dJ, jd1 = (c11 * q1 + c12 * q2) * dJ;
  const float d0 = tr.x - u, d1 = tr.y - v;
  const float err = sqrtf(d0 * d0 + d1 * d1);
  // CauchyWeight(x, delta): a = delta / x; abs(a) / sqrt(2) * sqrt(log(1 + 1 / a / a))
  const float a = cauchy_delta / err;
  const float w = fabsf(a) / sqrtf(2.0f) * sqrtf(logf(1.0f + 1.0f / a / a));
  const float ws = w / sigma;
  for (int j = 0; j < 6; ++j) {
    float s00 = 0.f, s01 = 0.f, s10 = 0.f, s11 = 0.f;
    for (int k = 0; k < 6; ++k) {
      s00 += A0[k] * sp.P0[k * 6 + j];
      s01 += A0[k] * sp.P1[k * 6 + j];
      s10 += A1[k] * sp.P0[k * 6 + j];
      s11 += A1[k] * sp.P1[k * 6 + j];
    }
    r0[j] = s00 * w / sigma; r0[6 + j] = s01 * w / sigma;
    r1[j] = s10 * w / sigma; r1[6 + j] = s11 * w / sigma;
  }
  for (int k = 0; k < C; ++k) {
    const float jc = __ldg(jr + k);
    r0[12 + k] = jd0 * jc * w / sigma;
    r1[12 + k] = jd1 * jc * w / sigma;
  }
  (void)ws;
  r0[12 + C] = d0 * w / sigma;
  r1[12 + C] = d1 * w / sigma;
  err2[i] = err * err;
}

// ---------------------------------------------------------------------------------------------------------------------
// SparseGeometricFactor::linearize (sources/core/gtsam/sparse_geometric_factor.cpp:157-271): one row per sampled pixel of
// keyframe 0 -- the depth keyframe 1 decodes at the (nearest-neighbour) correspondence against the depth of the warped
// point.  The reference runs it on the CPU over host mirrors of BOTH keyframes' level-0 proximity / code-Jacobian
// pyramids and of kf1's depth gradient (:181-183, :207-209, :220).  One thread per point:
//   dpt0   = DepthFromCode(c0, prx0_J_cde, prx0_0code, avg_dpt)                                   :186
//   corr   = FindCorrespondence(pt, dpt0, cam, pose10)  (border 1, min_dpt 0, bounds checked)     :187-198
//   dpt1_p = corr.tpt.z ; pix1_nn = (int) corr.pix1 ; dpt1 = DepthFromCode(c1, kf1 @ pix1_nn)      :201-210
//   err    = dpt1 - dpt1_p                                                                         :213
//   J_pose0/1 = ( TransformJacobianPose.row(2) - dpt_grad * corr_J_pose10 ) * pose10_J_pose0/1     :223-238
//   J_cde0 = (R ray).z * DepthJacobianPrx(dpt0) * prx0_J_cde - dpt_grad * corr_J_cde0              :241-246
//   J_cde1 = -DepthJacobianPrx(dpt1) * prx1_J_cde                                                  :249
//   everything * HuberWeight(err, huber_delta)                                                     :252-258
// The decode and the validity chain use round-to-nearest intrinsics in the reference's operation order (as the dense
// kernels do), so the set of valid rows and the nearest-neighbour pixels are those of the CPU evaluation.
template <int C>
__global__ void __launch_bounds__(128)
sparse_geometric_rows_kernel(SparsePose sp, float cam_w, float cam_h, const float* __restrict__ code0,
                             const float* __restrict__ code1, View prx0, View jac0, View prx1, View jac1, View grad1, int width,
                             int height, int num_points, const int2* __restrict__ points, float huber_delta, float avg_dpt,
                             float* __restrict__ rows)
{
  constexpr int RW = 13 + 2 * C;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= num_points) return;
  float* r = rows + (size_t)i * RW;
  const int2 pt = points[i];
  bool valid = pt.x >= 0 && pt.y >= 0 && pt.x < width && pt.y < height;  // the reference would read out of bounds
  Warped w;
  w.valid = false;
  float dpt0 = 0.0f;
  const float* jr0 = nullptr;
  if (valid) {
    jr0 = jac0.ptr + (size_t)pt.y * jac0.pitch + (size_t)pt.x * C;
    float dot = 0.0f;
    for (int k = 0; k < C; ++k) dot = __fadd_rn(dot, __fmul_rn(__ldg(jr0 + k), code0[k]));  // (prx_J_cde * code)(0), left to right
    dpt0 = prx_to_depth(__fadd_rn(__ldg(prx0.ptr + (size_t)pt.y * prx0.pitch + pt.x), dot), avg_dpt);
    w = warp_pixel((float)pt.x, (float)pt.y, dpt0, sp.q, sp.t, sp.fx, sp.fy, sp.u0, sp.v0, 1.0f, __fsub_rn(cam_w, 1.0f),
                   __fsub_rn(cam_h, 1.0f), 0.0f);
    valid = w.valid;
  }
  if (!valid) {
    for (int k = 0; k < RW; ++k) r[k] = 0.0f;
    return;
  }
  const int nx = (int)w.u, ny = (int)w.v;  // pix1.cast<int>()
  const float* jr1 = jac1.ptr + (size_t)ny * jac1.pitch + (size_t)nx * C;
  float dot1 = 0.0f;
  for (int k = 0; k < C; ++k) dot1 = __fadd_rn(dot1, __fmul_rn(__ldg(jr1 + k), code1[k]));
  const float dpt1 = prx_to_depth(__fadd_rn(__ldg(prx1.ptr + (size_t)ny * prx1.pitch + nx), dot1), avg_dpt);
  const float err = __fsub_rn(dpt1, w.tz);
  const float g0 = __ldg(grad1.ptr + (size_t)ny * grad1.pitch + 2 * nx), g1 = __ldg(grad1.ptr + (size_t)ny * grad1.pitch + 2 * nx + 1);
  const float X = w.tx, Y = w.ty, Z = w.tz;
  const float c00 = sp.fx / Z, c02 = -(sp.fx * X) / Z / Z, c11 = sp.fy / Z, c12 = -(sp.fy * Y) / Z / Z;  // ProjectPointJacobian
  const float A0[6] = {c00, 0.f, c02, c02 * w.py, c00 * w.pz - c02 * w.px, -(c00 * w.py)};  // corr_J_pose10 = dCam [I | -hat(R pt)]
  const float A1[6] = {0.f, c11, c12, c12 * w.py - c11 * w.pz, -(c12 * w.px), c11 * w.px};
  const float T2[6] = {0.f, 0.f, 1.f, w.py, -w.px, 0.f};  // row 2 of TransformJacobianPose
  float B[6];
#pragma unroll
  for (int k = 0; k < 6; ++k) B[k] = T2[k] - (g0 * A0[k] + g1 * A1[k]);
  const float hw = huber_weight(err, huber_delta);
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      s0 += B[k] * sp.P0[k * 6 + j];
      s1 += B[k] * sp.P1[k * 6 + j];
    }
    r[j] = s0 * hw;
    r[6 + j] = s1 * hw;
  }
  // pix1_J_dpt = dCam * R * ray ; (R ray).z ; dpt_J_prx = -avg / prx^2
  const float q0 = sp.R[0] * w.xn + sp.R[1] * w.yn + sp.R[2];
  const float q1 = sp.R[3] * w.xn + sp.R[4] * w.yn + sp.R[5];
  const float q2 = sp.R[6] * w.xn + sp.R[7] * w.yn + sp.R[8];
  const float pr0 = avg_dpt / (avg_dpt + dpt0), dJ0 = -avg_dpt / (pr0 * pr0);
  const float jd0 = (c00 * q0 + c02 * q2) * dJ0, jd1 = (c11 * q1 + c12 * q2) * dJ0;
  const float e0 = (q2 * dJ0 - (g0 * jd0 + g1 * jd1)) * hw;
  const float pr1 = avg_dpt / (avg_dpt + dpt1), dJ1 = -avg_dpt / (pr1 * pr1);
  const float e1 = -dJ1 * hw;
  for (int k = 0; k < C; ++k) {
    r[12 + k] = e0 * __ldg(jr0 + k);
    r[12 + C + k] = e1 * __ldg(jr1 + k);
  }
  r[12 + 2 * C] = err * hw;
}

}  // namespace

cudaError_t launch_reprojection_rows(const SparsePose& sp, const float* code_dev, int code_size, View prx_orig, View jac,
                                     int width, int height, int num_matches, const float* query_dev, const float* train_dev,
                                     float cauchy_delta, float sigma, float avg_dpt, float* rows_dev, float* err2_dev,
                                     cudaStream_t s)
{
  const int blocks = (num_matches + 127) / 128;
  const float2* q = reinterpret_cast<const float2*>(query_dev);
  const float2* t = reinterpret_cast<const float2*>(train_dev);
#define DFK_SP(CS)                                                                                                      \
  case CS:                                                                                                              \
    reprojection_rows_kernel<CS><<<blocks, 128, 0, s>>>(sp, code_dev, prx_orig, jac, width, height, num_matches, q, t,   \
                                                        cauchy_delta, sigma, avg_dpt, rows_dev, err2_dev);             \
    break;
  switch (code_size) {
    DFK_SP(8)
    DFK_SP(16)
    DFK_SP(32)
    DFK_SP(64)
    DFK_SP(128)
    default: return cudaErrorInvalidValue;
  }
#undef DFK_SP
  return cudaGetLastError();
}

cudaError_t launch_sparse_geometric_rows(const SparsePose& sp, float cam_w, float cam_h, const float* code0_dev,
                                         const float* code1_dev, int code_size, View prx0, View jac0, View prx1, View jac1,
                                         View grad1, int width, int height, int num_points, const int* points_dev,
                                         float huber_delta, float avg_dpt, float* rows_dev, cudaStream_t s)
{
  const int blocks = (num_points + 127) / 128;
  const int2* pts = reinterpret_cast<const int2*>(points_dev);
#define DFK_SG(CS)                                                                                                        \
  case CS:                                                                                                                \
    sparse_geometric_rows_kernel<CS><<<blocks, 128, 0, s>>>(sp, cam_w, cam_h, code0_dev, code1_dev, prx0, jac0, prx1, jac1, \
                                                            grad1, width, height, num_points, pts, huber_delta, avg_dpt,  \
                                                            rows_dev);                                                    \
    break;
  switch (code_size) {
    DFK_SG(8)
    DFK_SG(16)
    DFK_SG(32)
    DFK_SG(64)
    DFK_SG(128)
    default: return cudaErrorInvalidValue;
  }
#undef DFK_SG
  return cudaGetLastError();
}

}  // namespace dfk
