// oracle/ref_driver.cpp -- builds oracle/_ref/libdfk_ref.so: the REFERENCE's own per-pixel headers
//   /root/reference/sources/common/algorithm/{dense_sfm,warping,pinhole_camera(_impl),lucas_kanade_se3,m_estimators}.h
//   /root/reference/sources/cuda/{reduction_items,kernel_utils}.h
// compiled unmodified from where they lie (no reference source is copied into this repository), against the stand-in
// third-party headers under oracle/shim/ (Eigen, Sophus, VisionCore and OpenCV are not installed in this image), and
// driven by the host loop of the reference's own GPU-vs-CPU test (tests/ut_sfmaligner.cpp:297-315).
//
// TEST INFRASTRUCTURE ONLY: tests/test_oracle_ref.py checks the hand-written oracle (oracle/dfk_oracle_impl.inc) against
// this library; bench.py may time it as the CPU baseline ("kind": "reference").  Nothing in the product links it.
// What this pins: every line of the reference's own math (DenseSfm, DenseSfm_EvaluateError, LucasKanadeSE3,
// FindCorrespondence + Jacobians, RelativePose + Jacobians, PinholeCamera, HuberWeight, DepthFromCode).  What it cannot
// pin: the conventions restated in oracle/shim/ (getBilinear, quaternion rotate, SO3 product) -- see the shim headers.
//
// The C entry points mirror oracle/dfk_oracle.h (pitches in ELEMENTS of float) so the test can call both alike.
#include <chrono>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

#include <Eigen/Core>
#include <sophus/se3.hpp>
#include <VisionCore/Buffers/Image2D.hpp>

#include "dense_sfm.h"
#include "lucas_kanade_se3.h"
#include "warping.h"

namespace {

using Grad = Eigen::Matrix<float, 1, 2>;
using ImgView = vc::Image2DView<float, vc::TargetHost>;
using GradView = vc::Image2DView<Grad, vc::TargetHost>;

template <typename T>
Sophus::SE3<T> make_se3(const T p[7])
{
  return Sophus::SE3<T>(Sophus::SO3<T>(p[0], p[1], p[2], p[3]), Eigen::Matrix<T, 3, 1>(p[4], p[5], p[6]));
}

template <typename T>
void store_se3(const Sophus::SE3<T>& s, T p[7])
{
  p[0] = s.so3().x(); p[1] = s.so3().y(); p[2] = s.so3().z(); p[3] = s.so3().w();
  p[4] = s.translation()[0]; p[5] = s.translation()[1]; p[6] = s.translation()[2];
}

ImgView view(const float* p, int w, int h, size_t pitch_floats)
{
  return ImgView(const_cast<float*>(p), (size_t)w, (size_t)h, pitch_floats * sizeof(float));
}

struct Cam6 {
  float fx, fy, u0, v0, width, height;
};
struct Params5 {
  float huber_delta, ocl_th, avg_dpt, min_dpt;
  int valid_border;
};

// lucas_kanade_se3.h:74 writes `ReductionItem::HessianType(J.transpose())` without `typename`: fine for nvcc's device
// front end (the only place the reference instantiates it), ill-formed for g++ when HessianType is a type.  The item type
// is a template parameter of LucasKanadeSE3, so the host build passes this derived item whose `HessianType` IS a
// non-type: a static function building the very same packed matrix.  The reference header stays untouched.
struct LkItem : df::JTJJrReductionItem<float, 6> {
  using Base = df::JTJJrReductionItem<float, 6>;
  static Base::HessianType HessianType(const Base::JacobianType& v) { return Base::HessianType(v); }
};

template <int CS>
void sfm_run_step(const float* pose0, const float* pose1, const Cam6* c, int width, int height, const float* img0,
                  size_t img0_pitch, const float* img1, size_t img1_pitch, const float* dpt0, size_t dpt0_pitch,
                  float* valid0, size_t valid0_pitch, const float* jac, size_t jac_pitch, const float* grad1,
                  size_t grad1_pitch, const Params5* prm, float* JtJ, float* Jtr, float* residual, uint64_t* inliers)
{
  using Item = df::JTJJrReductionItem<float, 12 + CS>;
  const Sophus::SE3f p0 = make_se3(pose0), p1 = make_se3(pose1);
  const df::PinholeCamera<float> cam(c->fx, c->fy, c->u0, c->v0, c->width, c->height);
  df::DenseSfmParams sp;
  sp.huber_delta = prm->huber_delta; sp.ocl_th = prm->ocl_th; sp.avg_dpt = prm->avg_dpt; sp.min_dpt = prm->min_dpt;
  sp.valid_border = prm->valid_border;
  // tests/ut_sfmaligner.cpp:297-300 (== cu_sfmaligner.cpp:164-166)
  Eigen::Matrix<float, 6, 6> pose10_J_pose0;
  Eigen::Matrix<float, 6, 6> pose10_J_pose1;
  const Sophus::SE3f pose_10 = df::RelativePose(p1, p0, pose10_J_pose1, pose10_J_pose0);
  Eigen::Matrix<float, CS, 1> code;
  code.setZero();  // dead inside DenseSfm (dense_sfm.h:133-201 never reads it)
  const ImgView vimg0 = view(img0, width, height, img0_pitch), vimg1 = view(img1, width, height, img1_pitch);
  const ImgView vdpt0 = view(dpt0, width, height, dpt0_pitch);
  const ImgView vstd0 = view(dpt0, width, height, dpt0_pitch);  // std0 only feeds the dead uncertainty weight (:66)
  ImgView vvalid0 = view(valid0, width, height, valid0_pitch);
  const ImgView vjac(const_cast<float*>(jac), (size_t)width * CS, (size_t)height, jac_pitch * sizeof(float));
  const GradView vgrad(reinterpret_cast<Grad*>(const_cast<float*>(grad1)), (size_t)width, (size_t)height,
                       grad1_pitch * sizeof(float));
  Item result;
  // tests/ut_sfmaligner.cpp:303-315: x outer, y inner
  for (int x = 0; x < width; x += 1)
    for (int y = 0; y < height; y += 1)
      df::DenseSfm<float, CS, vc::TargetHost>(x, y, pose_10, pose10_J_pose0, pose10_J_pose1, code, cam, vimg0, vimg1,
                                              vdpt0, vstd0, vvalid0, vjac, vgrad, sp, result);
  constexpr int NP = 12 + CS;
  for (int k = 0; k < NP * (NP + 1) / 2; ++k) JtJ[k] = result.JtJ.coeff()(k);
  for (int k = 0; k < NP; ++k) Jtr[k] = result.Jtr(k);
  *residual = result.residual;
  *inliers = result.inliers;
}

template <int CS>
void update_depth(const float* code_in, int width, int height, const float* prx_orig, size_t prx_pitch,
                  const float* jac, size_t jac_pitch, float avg_dpt, float* dpt_out, size_t dpt_pitch)
{
  Eigen::Matrix<float, CS, 1> code;
  for (int k = 0; k < CS; ++k) code(k) = code_in[k];
  const ImgView vprx = view(prx_orig, width, height, prx_pitch);
  const ImgView vjac(const_cast<float*>(jac), (size_t)width * CS, (size_t)height, jac_pitch * sizeof(float));
  ImgView vout = view(dpt_out, width, height, dpt_pitch);
  // body of kernel_update_depth (cuda/cu_image_proc.cpp:255-263), every pixel once
  for (int y = 0; y < height; ++y)
    for (int x = 0; x < width; ++x) {
      Eigen::Map<const Eigen::Matrix<float, 1, CS>> tmp(&vjac(x * CS, y));
      const Eigen::Matrix<float, 1, CS> prx_J_cde(tmp);
      vout(x, y) = df::DepthFromCode(code, prx_J_cde, vprx(x, y), avg_dpt);
    }
}

// Body of kernel_depthaligner_run_step (cuda/cu_depthaligner.cpp:46-65) run for every pixel once, row-major: the kernel
// is a __device__ lambda inside a .cpp compiled as CUDA, so the statements are repeated here around the reference's own
// df::DepthFromCode / df::DepthJacobianPrx (warping.h) and its reduction item.  avg_dpt = 2 is hard-coded there (:44).
template <int CS>
void depth_run_step(const float* code_in, int width, int height, const float* tgt, size_t tgt_pitch, const float* prx_orig,
                    size_t prx_pitch, const float* jac, size_t jac_pitch, float avg_dpt, float* JtJ, float* Jtr,
                    float* residual, uint64_t* inliers)
{
  using Item = df::JTJJrReductionItem<float, CS>;
  Eigen::Matrix<float, CS, 1> code;
  for (int k = 0; k < CS; ++k) code(k) = code_in[k];
  const ImgView vtgt = view(tgt, width, height, tgt_pitch), vprx = view(prx_orig, width, height, prx_pitch);
  const ImgView vjac(const_cast<float*>(jac), (size_t)width * CS, (size_t)height, jac_pitch * sizeof(float));
  Item sum;
  for (int y = 0; y < height; ++y)
    for (int x = 0; x < width; ++x) {
      Eigen::Map<const Eigen::Matrix<float, 1, CS>> tmp(&vjac(x * CS, y));
      const Eigen::Matrix<float, 1, CS> prx_J_cde(tmp);
      float dpt = df::DepthFromCode(code, prx_J_cde, vprx(x, y), avg_dpt);
      float diff = vtgt(x, y) - dpt;
      Eigen::Matrix<float, 1, CS> J = -2 * abs(diff) * df::DepthJacobianPrx(dpt, avg_dpt) * prx_J_cde;
      sum.inliers += 1;
      sum.residual += diff * diff;
      sum.Jtr += J.transpose() * diff;
      sum.JtJ += typename Item::HessianType(J.transpose());
    }
  for (int k = 0; k < CS * (CS + 1) / 2; ++k) JtJ[k] = sum.JtJ.coeff()(k);
  for (int k = 0; k < CS; ++k) Jtr[k] = sum.Jtr(k);
  *residual = sum.residual;
  *inliers = sum.inliers;
}

// ReprojectionFactor::linearize (core/gtsam/reprojection_factor.cpp:175-258): the factor's loop body around the
// reference's own warping.h / m_estimators.h functions (the factor itself needs GTSAM + OpenCV keypoints, which are not
// here): every statement between the lookup of the keypoints and the fill of Ab, in the reference's order.
template <int CS>
float reprojection_rows(const float* pose0, const float* pose1, const float* code_in, const Cam6* c, int width, int height,
                        const float* prx_orig, size_t prx_pitch, const float* jac, size_t jac_pitch, int num_matches,
                        const float* query_xy, const float* train_xy, float huber_delta_, float sigma_, float avg_dpt,
                        float* rows)
{
  using Scalar = float;
  const int RW = 13 + CS;
  const Sophus::SE3f p0 = make_se3(pose0), p1 = make_se3(pose1);
  const df::PinholeCamera<float> cam_(c->fx, c->fy, c->u0, c->v0, c->width, c->height);
  Eigen::Matrix<Scalar, CS, 1> c0;
  for (int k = 0; k < CS; ++k) c0(k) = code_in[k];
  const ImgView prx_img = view(prx_orig, width, height, prx_pitch);
  ImgView prx0_jac_img(const_cast<float*>(jac), (size_t)width * CS, (size_t)height, jac_pitch * sizeof(float));
  Scalar total_err = 0;
  for (int i = 0; i < num_matches; ++i) {
    float* r0 = rows + (size_t)(2 * i) * RW;
    float* r1 = r0 + RW;
    for (int k = 0; k < RW; ++k) { r0[k] = 0; r1[k] = 0; }
    const float qx = query_xy[2 * i], qy = query_xy[2 * i + 1];
    if ((int)qx < 0 || (int)qy < 0 || (int)qx >= width || (int)qy >= height) continue;
    Eigen::Matrix<Scalar, 2, 1> pix1(train_xy[2 * i], train_xy[2 * i + 1]);
    Eigen::Matrix<Scalar, 6, 6> pose10_J_pose0;
    Eigen::Matrix<Scalar, 6, 6> pose10_J_pose1;
    Sophus::SE3f pose10 = df::RelativePose(p1, p0, pose10_J_pose1, pose10_J_pose0);
    Eigen::Map<const Eigen::Matrix<Scalar, 1, CS>> tmp(&prx0_jac_img((int)qx * CS, (int)qy));
    const Eigen::Matrix<Scalar, 1, CS> prx_J_cde(tmp);
    Scalar prx_0code = prx_img(qx, qy);
    Scalar dpt0 = df::DepthFromCode(c0, prx_J_cde, prx_0code, avg_dpt);
    df::Correspondence<Scalar> corr = df::FindCorrespondence(qx, qy, dpt0, cam_, pose10, 1.f, 0.f, false);
    if (not corr.valid) continue;
    Eigen::Matrix<Scalar, 2, CS> corr_J_cde;
    df::FindCorrespondenceJacobianCode(corr, dpt0, cam_, pose10, prx_J_cde, avg_dpt, corr_J_cde);
    Eigen::Matrix<Scalar, 2, 6> corr_J_pose10;
    corr_J_pose10 = df::FindCorrespondenceJacobianPose(corr, dpt0, cam_, pose10);
    Eigen::Matrix<Scalar, 2, 6> err_J_pose0 = corr_J_pose10 * pose10_J_pose0;
    Eigen::Matrix<Scalar, 2, 6> err_J_pose1 = corr_J_pose10 * pose10_J_pose1;
    Eigen::Matrix<Scalar, 2, CS> err_J_cde = corr_J_cde;
    Eigen::Matrix<Scalar, 2, 1> diff = pix1 - corr.pix1;
    Scalar err = diff.norm();
    Scalar hbr_wgt = df::CauchyWeight(err, huber_delta_);
    err_J_pose0 *= hbr_wgt;
    err_J_pose1 *= hbr_wgt;
    err_J_cde *= hbr_wgt;
    diff *= hbr_wgt;
    total_err += err * err;
    err_J_pose0 = err_J_pose0 / sigma_;
    err_J_pose1 = err_J_pose1 / sigma_;
    err_J_cde = err_J_cde / sigma_;
    diff = diff / sigma_;
    for (int j = 0; j < 6; ++j) {
      r0[j] = err_J_pose0(0, j); r1[j] = err_J_pose0(1, j);
      r0[6 + j] = err_J_pose1(0, j); r1[6 + j] = err_J_pose1(1, j);
    }
    for (int k = 0; k < CS; ++k) { r0[12 + k] = err_J_cde(0, k); r1[12 + k] = err_J_cde(1, k); }
    r0[12 + CS] = diff(0);
    r1[12 + CS] = diff(1);
  }
  return total_err;
}

// SparseGeometricFactor::linearize's rows (core/gtsam/sparse_geometric_factor.cpp:157-271) over the reference's own
// warping.h / dense_sfm.h / pinhole_camera.h functions (the factor itself needs GTSAM and the Keyframe class): every
// statement between the lookup of the point and the fill of Ab, in the reference's order.  dpt_grad1 = kf1->dpt_grad
// (mapper.cpp:998-1000: SobelGradients of the keyframe's level-0 depth), 2 floats per pixel.
template <int CS>
int sparse_geometric_rows(const float* pose0, const float* pose1, const float* code0_in, const float* code1_in, const Cam6* c,
                          int width, int height, const float* prx0_orig, size_t prx0_pitch, const float* jac0, size_t jac0_pitch,
                          const float* prx1_orig, size_t prx1_pitch, const float* jac1, size_t jac1_pitch,
                          const float* dpt_grad1, size_t grad_pitch, int num_points, const int* points_xy, float huber_delta_,
                          float avg_dpt, float* rows)
{
  using Scalar = float;
  const int RW = 13 + 2 * CS;
  const Sophus::SE3f p0 = make_se3(pose0), p1 = make_se3(pose1);
  const df::PinholeCamera<float> cam_(c->fx, c->fy, c->u0, c->v0, c->width, c->height);
  Eigen::Matrix<Scalar, CS, 1> c0, c1;
  for (int k = 0; k < CS; ++k) { c0(k) = code0_in[k]; c1(k) = code1_in[k]; }
  const ImgView prx0_img = view(prx0_orig, width, height, prx0_pitch);
  const ImgView prx1_img = view(prx1_orig, width, height, prx1_pitch);
  ImgView prx0_jac_img(const_cast<float*>(jac0), (size_t)width * CS, (size_t)height, jac0_pitch * sizeof(float));
  ImgView prx1_jac_img(const_cast<float*>(jac1), (size_t)width * CS, (size_t)height, jac1_pitch * sizeof(float));
  int valid_rows = 0;
  for (int i = 0; i < num_points; ++i) {
    float* r = rows + (size_t)i * RW;
    for (int k = 0; k < RW; ++k) r[k] = 0;
    const int ptx = points_xy[2 * i], pty = points_xy[2 * i + 1];
    if (ptx < 0 || pty < 0 || ptx >= width || pty >= height) continue;
    Eigen::Matrix<Scalar, 6, 6> pose10_J_pose0;
    Eigen::Matrix<Scalar, 6, 6> pose10_J_pose1;
    Sophus::SE3f pose10 = df::RelativePose(p1, p0, pose10_J_pose1, pose10_J_pose0);
    Eigen::Map<const Eigen::Matrix<Scalar, 1, CS>> tmp0(&prx0_jac_img(ptx * CS, pty));
    const Eigen::Matrix<Scalar, 1, CS> prx0_J_cde(tmp0);
    Scalar prx0_0code = prx0_img(ptx, pty);
    Scalar dpt0 = df::DepthFromCode(c0, prx0_J_cde, prx0_0code, avg_dpt);
    df::Correspondence<Scalar> corr = df::FindCorrespondence((std::size_t)ptx, (std::size_t)pty, dpt0, cam_, pose10);
    if (!cam_.PixelValid(corr.pix1) || !corr.valid) continue;
    Scalar dpt1_p = corr.tpt(2);
    const int nn0 = (int)corr.pix1(0), nn1 = (int)corr.pix1(1);  // pix1.cast<int>()
    Eigen::Map<const Eigen::Matrix<Scalar, 1, CS>> tmp1(&prx1_jac_img(nn0 * CS, nn1));
    const Eigen::Matrix<Scalar, 1, CS> prx1_J_cde(tmp1);
    Scalar prx1_0code = prx1_img(nn0, nn1);
    Scalar dpt1 = df::DepthFromCode(c1, prx1_J_cde, prx1_0code, avg_dpt);
    Scalar err = dpt1 - dpt1_p;
    Eigen::Matrix<Scalar, 1, 2> dpt_grad;
    dpt_grad(0, 0) = dpt_grad1[(size_t)nn1 * grad_pitch + 2 * nn0];
    dpt_grad(0, 1) = dpt_grad1[(size_t)nn1 * grad_pitch + 2 * nn0 + 1];
    Eigen::Matrix<Scalar, 2, 6> corr_J_pose10 = df::FindCorrespondenceJacobianPose(corr, dpt0, cam_, pose10);
    Eigen::Matrix<Scalar, 3, 6> tpt_J_pose0 = df::TransformJacobianPose(corr.pt, pose10) * pose10_J_pose0;
    const Eigen::Matrix<Scalar, 3, 6>& tpt_J_pose0_c = tpt_J_pose0;
    Eigen::Matrix<Scalar, 1, 6> dpt1p_J_pose0 = tpt_J_pose0_c.template block<1, 6>(2, 0);
    Eigen::Matrix<Scalar, 1, 6> g_pose10 = dpt_grad * corr_J_pose10;
    Eigen::Matrix<Scalar, 1, 6> err_J_pose0 = dpt1p_J_pose0 - g_pose10 * pose10_J_pose0;
    Eigen::Matrix<Scalar, 3, 6> tpt_J_pose1 = df::TransformJacobianPose(corr.pt, pose10) * pose10_J_pose1;
    const Eigen::Matrix<Scalar, 3, 6>& tpt_J_pose1_c = tpt_J_pose1;
    Eigen::Matrix<Scalar, 1, 6> dpt1p_J_pose1 = tpt_J_pose1_c.template block<1, 6>(2, 0);
    Eigen::Matrix<Scalar, 1, 6> err_J_pose1 = dpt1p_J_pose1 - g_pose10 * pose10_J_pose1;
    Eigen::Matrix<Scalar, 2, CS> corr_J_cde0;
    df::FindCorrespondenceJacobianCode(corr, dpt0, cam_, pose10, prx0_J_cde, avg_dpt, corr_J_cde0);
    Eigen::Matrix<Scalar, 3, 3> trans_J_pt = df::TransformJacobianPoint(corr.pt, pose10);
    Eigen::Matrix<Scalar, 3, 1> trans_J_dpt = trans_J_pt * cam_.ReprojectDepthJacobian(corr.pix0, dpt0);
    Eigen::Matrix<Scalar, 3, CS> trans_J_cde = (trans_J_dpt * df::DepthJacobianPrx(dpt0, avg_dpt)) * prx0_J_cde;
    const Eigen::Matrix<Scalar, 3, CS>& trans_J_cde_c = trans_J_cde;  // the shim's non-const block() is an assignable view
    Eigen::Matrix<Scalar, 1, CS> err_J_cde0 = trans_J_cde_c.template block<1, CS>(2, 0) - dpt_grad * corr_J_cde0;
    Eigen::Matrix<Scalar, 1, CS> err_J_cde1 = prx1_J_cde * (-df::DepthJacobianPrx(dpt1, avg_dpt));
    Scalar hbr_wgt = df::DenseSfm_RobustLoss(err, huber_delta_);
    err *= hbr_wgt;
    err_J_pose0 *= hbr_wgt;
    err_J_pose1 *= hbr_wgt;
    err_J_cde0 *= hbr_wgt;
    err_J_cde1 *= hbr_wgt;
    for (int j = 0; j < 6; ++j) { r[j] = err_J_pose0(0, j); r[6 + j] = err_J_pose1(0, j); }
    for (int k = 0; k < CS; ++k) { r[12 + k] = err_J_cde0(0, k); r[12 + CS + k] = err_J_cde1(0, k); }
    r[12 + 2 * CS] = err;
    ++valid_rows;
  }
  return valid_rows;
}

}  // namespace

extern "C" {

int dfkr_supports_code_size(int cs) { return cs == 8 || cs == 16 || cs == 32 || cs == 64 || cs == 128; }

// warping.h:98-137; jac_a / jac_b 6x6 row-major
void dfkr_relative_pose_f(const float a[7], const float b[7], float ab[7], float jac_a[36], float jac_b[36])
{
  Eigen::Matrix<float, 6, 6> ja, jb;
  const Sophus::SE3f r = df::RelativePose(make_se3(a), make_se3(b), ja, jb);
  store_se3(r, ab);
  for (int i = 0; i < 6; ++i)
    for (int j = 0; j < 6; ++j) {
      if (jac_a) jac_a[i * 6 + j] = ja(i, j);
      if (jac_b) jac_b[i * 6 + j] = jb(i, j);
    }
}
void dfkr_relative_pose_d(const double a[7], const double b[7], double ab[7], double jac_a[36], double jac_b[36])
{
  Eigen::Matrix<double, 6, 6> ja, jb;
  const Sophus::SE3d r = df::RelativePose(make_se3(a), make_se3(b), ja, jb);
  store_se3(r, ab);
  for (int i = 0; i < 6; ++i)
    for (int j = 0; j < 6; ++j) {
      if (jac_a) jac_a[i * 6 + j] = ja(i, j);
      if (jac_b) jac_b[i * 6 + j] = jb(i, j);
    }
}

// FindCorrespondence + its Jacobians for one (integer) pixel, double: out[0]=valid, out[1..2]=pix1,
// out[3..14]=corresp_J_pose (2x6 row major), out[15..16]=pix1_J_prx   (same packing as dfko_probe_pixel_d)
void dfkr_probe_pixel_d(size_t x, size_t y, double dpt, const Cam6* c, const double pose[7], int border, double min_dpt,
                        double avg_dpt, double out[17])
{
  const df::PinholeCamera<double> cam(c->fx, c->fy, c->u0, c->v0, c->width, c->height);
  const Sophus::SE3d se3 = make_se3(pose);
  for (int i = 0; i < 17; ++i) out[i] = 0;
  // check_bounds = false keeps pt / tpt filled for out-of-image pixels (the Jacobians only need tpt)
  const df::Correspondence<double> bounded = df::FindCorrespondence(x, y, dpt, cam, se3, border, min_dpt);
  const df::Correspondence<double> corr = df::FindCorrespondence(x, y, dpt, cam, se3, border, min_dpt, false);
  out[0] = bounded.valid ? 1.0 : 0.0;
  if (!corr.valid) return;
  out[1] = corr.pix1[0];
  out[2] = corr.pix1[1];
  const Eigen::Matrix<double, 2, 6> A = df::FindCorrespondenceJacobianPose(corr, dpt, cam, se3);
  for (int k = 0; k < 6; ++k) {
    out[3 + k] = A(0, k);
    out[9 + k] = A(1, k);
  }
  Eigen::Matrix<double, 2, 1> pJ;
  df::FindCorrespondenceJacobianPrx(corr, dpt, cam, se3, avg_dpt, pJ);
  out[15] = pJ(0);
  out[16] = pJ(1);
}

float dfkr_huber_weight_f(float x, float delta) { return df::HuberWeight(x, delta); }

// SfmAligner::RunStep's math on the host: the loop of tests/ut_sfmaligner.cpp:297-315.  Returns 0, or -1 for a code size
// that is not instantiated here.
int dfkr_sfm_run_step_f(const float pose0[7], const float pose1[7], int code_size, const Cam6* cam, int width,
                        int height, const float* img0, size_t img0_pitch, const float* img1, size_t img1_pitch,
                        const float* dpt0, size_t dpt0_pitch, float* valid0, size_t valid0_pitch,
                        const float* prx0_jac, size_t jac_pitch, const float* grad1, size_t grad1_pitch,
                        const Params5* params, float* JtJ, float* Jtr, float* residual, uint64_t* inliers)
{
#define DFKR_CASE(CS)                                                                                                  \
  case CS:                                                                                                             \
    sfm_run_step<CS>(pose0, pose1, cam, width, height, img0, img0_pitch, img1, img1_pitch, dpt0, dpt0_pitch, valid0,   \
                     valid0_pitch, prx0_jac, jac_pitch, grad1, grad1_pitch, params, JtJ, Jtr, residual, inliers);      \
    return 0;
  switch (code_size) {
    DFKR_CASE(8)
    DFKR_CASE(16)
    DFKR_CASE(32)
    DFKR_CASE(64)
    DFKR_CASE(128)
    default: return -1;
  }
#undef DFKR_CASE
}

// SfmAligner::EvaluateError's math (cu_sfmaligner.cpp:120-147, dense_sfm.h:79-119), x outer / y inner
void dfkr_sfm_evaluate_error_f(const float pose0[7], const float pose1[7], const Cam6* c, int width, int height,
                               const float* img0, size_t img0_pitch, const float* img1, size_t img1_pitch,
                               const float* dpt0, size_t dpt0_pitch, const float* grad1, size_t grad1_pitch,
                               const Params5* prm, float* residual, uint64_t* inliers)
{
  const df::PinholeCamera<float> cam(c->fx, c->fy, c->u0, c->v0, c->width, c->height);
  df::DenseSfmParams sp;
  sp.huber_delta = prm->huber_delta; sp.ocl_th = prm->ocl_th; sp.avg_dpt = prm->avg_dpt; sp.min_dpt = prm->min_dpt;
  sp.valid_border = prm->valid_border;
  const Sophus::SE3f pose_10 = df::RelativePose(make_se3(pose1), make_se3(pose0));  // cu_sfmaligner.cpp:131
  const ImgView vimg0 = view(img0, width, height, img0_pitch), vimg1 = view(img1, width, height, img1_pitch);
  const ImgView vdpt0 = view(dpt0, width, height, dpt0_pitch);
  const GradView vgrad(reinterpret_cast<Grad*>(const_cast<float*>(grad1)), (size_t)width, (size_t)height,
                       grad1_pitch * sizeof(float));
  df::CorrespondenceReductionItem<float> result;
  for (int x = 0; x < width; ++x)
    for (int y = 0; y < height; ++y)
      df::DenseSfm_EvaluateError<float, 32, vc::TargetHost>(x, y, pose_10, cam, vimg0, vimg1, vdpt0, vdpt0, vgrad, sp,
                                                            result);
  *residual = result.residual;
  *inliers = result.inliers;
}

// SE3Aligner::RunStep's math (cu_se3aligner.cpp:37-59, lucas_kanade_se3.h:41-77), x outer / y inner
void dfkr_se3_run_step_f(const float se3[7], const Cam6* c, int width, int height, const float* img0, size_t img0_pitch,
                         const float* img1, size_t img1_pitch, const float* dpt0, size_t dpt0_pitch,
                         const float* grad1, size_t grad1_pitch, float huber_delta, float* JtJ, float* Jtr,
                         float* residual, uint64_t* inliers)
{
  using Item = LkItem;
  const df::PinholeCamera<float> cam(c->fx, c->fy, c->u0, c->v0, c->width, c->height);
  const Sophus::SE3f pose = make_se3(se3);
  const ImgView vimg0 = view(img0, width, height, img0_pitch), vimg1 = view(img1, width, height, img1_pitch);
  const ImgView vdpt0 = view(dpt0, width, height, dpt0_pitch);
  const GradView vgrad(reinterpret_cast<Grad*>(const_cast<float*>(grad1)), (size_t)width, (size_t)height,
                       grad1_pitch * sizeof(float));
  Item sum;
  for (int x = 0; x < width; ++x)
    for (int y = 0; y < height; ++y)
      sum += df::LucasKanadeSE3<float, vc::TargetHost, Sophus::SE3f, df::PinholeCamera<float>, Item>(
          x, y, pose, cam, vimg0, vimg1, vdpt0, vgrad, huber_delta);
  for (int k = 0; k < 21; ++k) JtJ[k] = sum.JtJ.coeff()(k);
  for (int k = 0; k < 6; ++k) Jtr[k] = sum.Jtr(k);
  *residual = sum.residual;
  *inliers = sum.inliers;
}

// CPU baseline in THROUGHPUT mode (bench.py --impl reference): `nthreads` threads, each running the reference's own
// single-threaded host loop (dfkr_sfm_run_step_f above, x outer / y inner as ut_sfmaligner.cpp:303-315) over the whole
// pyramid of one pair `evals_per_thread` times on the same read-only inputs.  Same level struct as
// oracle/dfk_oracle.h (DfkoLevel).  Returns the wall time in seconds; -1 for an uninstantiated code size.
struct RefLevel {
  Cam6 cam;
  int width, height;
  const float* img0; size_t img0_pitch;
  const float* img1; size_t img1_pitch;
  const float* dpt0; size_t dpt0_pitch;
  const float* prx0_jac; size_t jac_pitch;
  const float* grad1; size_t grad1_pitch;
};
double dfkr_sfm_throughput_f(int nthreads, int evals_per_thread, const float pose0[7], const float pose1[7],
                             int code_size, int nlevels, const RefLevel* levels, const Params5* params, float* rec_out)
{
  if (!dfkr_supports_code_size(code_size)) return -1.0;
  if (nthreads < 1) nthreads = 1;
  const int NP = 12 + code_size, NH = NP * (NP + 1) / 2, REC = NH + NP + 2;
  std::vector<float> recs((size_t)nthreads * REC, 0.0f);
  std::vector<std::vector<float>> valid((size_t)nthreads);
  auto worker = [&](int t) {
    float* rec = recs.data() + (size_t)t * REC;
    for (int e = 0; e < evals_per_thread; ++e)
      for (int l = nlevels - 1; l >= 0; --l) {
        const RefLevel& L = levels[l];
        valid[t].assign((size_t)L.width * L.height, 0.0f);  // DenseSfm writes valid0 (dense_sfm.h:161): thread-private
        uint64_t inl = 0;
        dfkr_sfm_run_step_f(pose0, pose1, code_size, &L.cam, L.width, L.height, L.img0, L.img0_pitch, L.img1,
                            L.img1_pitch, L.dpt0, L.dpt0_pitch, valid[t].data(), (size_t)L.width, L.prx0_jac, L.jac_pitch,
                            L.grad1, L.grad1_pitch, params, rec, rec + NH, rec + NH + NP, &inl);
        rec[NH + NP + 1] = (float)inl;
      }
  };
  const auto t0 = std::chrono::steady_clock::now();
  std::vector<std::thread> th;
  for (int t = 1; t < nthreads; ++t) th.emplace_back(worker, t);
  worker(0);
  for (auto& x : th) x.join();
  const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  if (rec_out) std::memcpy(rec_out, recs.data(), sizeof(float) * REC);
  return dt;
}

// DepthAligner::RunStep's math (cu_depthaligner.cpp:32-71)
int dfkr_depth_run_step_f(const float* code, int code_size, int width, int height, const float* tgt, size_t tgt_pitch,
                          const float* prx_orig, size_t prx_pitch, const float* prx_jac, size_t jac_pitch, float avg_dpt,
                          float* JtJ, float* Jtr, float* residual, uint64_t* inliers)
{
#define DFKR_DCASE(CS)                                                                                                   \
  case CS:                                                                                                               \
    depth_run_step<CS>(code, width, height, tgt, tgt_pitch, prx_orig, prx_pitch, prx_jac, jac_pitch, avg_dpt, JtJ, Jtr,  \
                       residual, inliers);                                                                              \
    return 0;
  switch (code_size) {
    DFKR_DCASE(8)
    DFKR_DCASE(16)
    DFKR_DCASE(32)
    DFKR_DCASE(64)
    default: return -1;
  }
#undef DFKR_DCASE
}

// ReprojectionFactor::linearize's rows (core/gtsam/reprojection_factor.cpp:157-269); returns total_err, -1 if the code
// size is not instantiated
float dfkr_reprojection_rows_f(const float pose0[7], const float pose1[7], const float* code, int code_size, const Cam6* cam,
                               int width, int height, const float* prx_orig, size_t prx_pitch, const float* prx_jac,
                               size_t jac_pitch, int num_matches, const float* query_xy, const float* train_xy,
                               float cauchy_delta, float sigma, float avg_dpt, float* rows)
{
  switch (code_size) {
    case 8: return reprojection_rows<8>(pose0, pose1, code, cam, width, height, prx_orig, prx_pitch, prx_jac, jac_pitch, num_matches, query_xy, train_xy, cauchy_delta, sigma, avg_dpt, rows);
    case 32: return reprojection_rows<32>(pose0, pose1, code, cam, width, height, prx_orig, prx_pitch, prx_jac, jac_pitch, num_matches, query_xy, train_xy, cauchy_delta, sigma, avg_dpt, rows);
    default: return -1.0f;
  }
}

// SparseGeometricFactor::linearize's rows; returns the number of non-zero (valid) rows, -1 if the code size is not
// instantiated
int dfkr_sparse_geometric_rows_f(const float pose0[7], const float pose1[7], const float* code0, const float* code1,
                                 int code_size, const Cam6* cam, int width, int height, const float* prx0_orig,
                                 size_t prx0_pitch, const float* jac0, size_t jac0_pitch, const float* prx1_orig,
                                 size_t prx1_pitch, const float* jac1, size_t jac1_pitch, const float* dpt_grad1,
                                 size_t grad_pitch, int num_points, const int* points_xy, float huber_delta, float avg_dpt,
                                 float* rows)
{
  switch (code_size) {
    case 8: return sparse_geometric_rows<8>(pose0, pose1, code0, code1, cam, width, height, prx0_orig, prx0_pitch, jac0, jac0_pitch, prx1_orig, prx1_pitch, jac1, jac1_pitch, dpt_grad1, grad_pitch, num_points, points_xy, huber_delta, avg_dpt, rows);
    case 32: return sparse_geometric_rows<32>(pose0, pose1, code0, code1, cam, width, height, prx0_orig, prx0_pitch, jac0, jac0_pitch, prx1_orig, prx1_pitch, jac1, jac1_pitch, dpt_grad1, grad_pitch, num_points, points_xy, huber_delta, avg_dpt, rows);
    default: return -1;
  }
}

// UpdateDepth's math (cu_image_proc.cpp:248-264, warping.h:52-69)
int dfkr_update_depth_f(const float* code, int code_size, int width, int height, const float* prx_orig,
                        size_t prx_pitch, const float* prx_jac, size_t jac_pitch, float avg_dpt, float* dpt_out,
                        size_t dpt_pitch)
{
  switch (code_size) {
    case 8: update_depth<8>(code, width, height, prx_orig, prx_pitch, prx_jac, jac_pitch, avg_dpt, dpt_out, dpt_pitch); return 0;
    case 16: update_depth<16>(code, width, height, prx_orig, prx_pitch, prx_jac, jac_pitch, avg_dpt, dpt_out, dpt_pitch); return 0;
    case 32: update_depth<32>(code, width, height, prx_orig, prx_pitch, prx_jac, jac_pitch, avg_dpt, dpt_out, dpt_pitch); return 0;
    case 64: update_depth<64>(code, width, height, prx_orig, prx_pitch, prx_jac, jac_pitch, avg_dpt, dpt_out, dpt_pitch); return 0;
    case 128: update_depth<128>(code, width, height, prx_orig, prx_pitch, prx_jac, jac_pitch, avg_dpt, dpt_out, dpt_pitch); return 0;
    default: return -1;
  }
}

}  // extern "C"
