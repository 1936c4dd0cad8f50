// dfk_facade.h -- header-only C++ facade that reproduces the reference's aligner API surface
// on top of the C ABI of libdfk.so (include/dfk.h).
//
// Same namespaces, class names, method names, argument order and result member names as
//   df::SfmAligner<Scalar,CS>      sources/cuda/cu_sfmaligner.h:50-97
//   df::SE3Aligner<Scalar>         sources/cuda/cu_se3aligner.h:38-86
//   df::SfmAlignerParams           sources/cuda/cu_sfmaligner.h:41-48
//   df::DenseSfmParams             sources/common/algorithm/dense_sfm.h:36-43
//   df::JTJJrReductionItem<T,NP>   sources/cuda/reduction_items.h:77-143
//   df::CorrespondenceReductionItem<T>   sources/cuda/reduction_items.h:35-71
//   df::UpdateDepth / SobelGradients / GaussianBlurDown / SquaredError   sources/cuda/cu_image_proc.h:27-44
// (jczarnowski/DeepFactors @ bffc78a).
//
// The reference's argument types come from Sophus, Eigen and VisionCore, none of which is a
// dependency of this repository.  The facade is therefore DUCK-TYPED: every method is a template
// over the argument types and only uses the members the real types have
//   pose      .data()                          -> const float*  (Sophus::SE3f: quaternion xyzw, translation)
//   code      .data()                          -> const float*  (Eigen::Matrix<float,CS,1>)
//   camera    .fx() .fy() .u0() .v0() .width() .height()       (df::PinholeCamera<float>)
//   image     .ptr() .pitch() .width() .height()               (vc::Image2DView / vc::Buffer2DView;
//                                                               pitch in BYTES, width in elements)
// so a build that does have those libraries passes its own objects straight through (see
// INTEGRATION.md), and a build without them uses the minimal stand-ins of df/dfk_standins.h.
// Results are returned by value like the reference; dense access goes through
// JtJ.toDenseMatrix(i, j) / JtJ.coeff() instead of an Eigen expression.
#ifndef DFK_FACADE_H_
#define DFK_FACADE_H_

#include <array>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../dfk.h"

namespace df
{

// ---------------------------------------------------------------------------------------------
// errors: the reference throws vc::CUDAException from CudaCheckLastError (launch_utils.h:26-32) and
// std::runtime_error (cu_sfmaligner.cpp:171-173); glog CHECKs abort.  Here everything is an exception.
// ---------------------------------------------------------------------------------------------
class CUDAException : public std::runtime_error
{
public:
  CUDAException(DfkStatus st, const std::string& msg) : std::runtime_error(msg), status(st) {}
  DfkStatus status;
};

namespace detail
{
inline void Check(DfkHandle h, DfkStatus st)
{
  if (st == DFK_OK) return;
  const char* m = dfk_last_error(h);
  throw CUDAException(st, (m && *m) ? std::string(m) : std::string(dfk_status_string(st)));
}

template <typename ImageT>
inline DfkImage View(const ImageT& img, unsigned floats_per_px)
{
  DfkImage v;
  v.ptr = const_cast<void*>(static_cast<const void*>(img.ptr()));
  v.pitch_bytes = img.pitch();
  // the reference views the code Jacobian as a (W*CS) x H float image (keyframe.h:52); gradients are
  // Eigen::Matrix<float,1,2> pixels.  `width` of the C ABI counts pixels.
  v.width = static_cast<uint32_t>(img.width() / (floats_per_px > 2 ? floats_per_px : 1));
  v.height = static_cast<uint32_t>(img.height());
  return v;
}

template <typename CamT>
inline DfkCamera Cam(const CamT& cam)
{
  return DfkCamera{static_cast<float>(cam.fx()), static_cast<float>(cam.fy()), static_cast<float>(cam.u0()),
                   static_cast<float>(cam.v0()), static_cast<float>(cam.width()), static_cast<float>(cam.height())};
}

struct HandleDeleter {
  void operator()(DfkContext* h) const { dfk_destroy(h); }
};
using HandlePtr = std::unique_ptr<DfkContext, HandleDeleter>;

inline HandlePtr MakeHandle()
{
  DfkHandle h = nullptr;
  DfkStatus st = dfk_create(-1, &h);
  if (st != DFK_OK) throw CUDAException(st, "dfk_create failed: " + std::string(dfk_status_string(st)));
  // Ordering contract of the drop-in facade = the reference's: every launch goes to the LEGACY DEFAULT stream
  // (cu_sfmaligner.cpp:175-179 launches without a stream argument), so inputs an integrator produced asynchronously on
  // stream 0 (uploads, pyramid / network kernels) are complete before an aligner reads them.  A handle's own stream is
  // non-blocking and NOT ordered against stream 0; callers that want it (or their own stream) say so with SetStream.
  dfk_set_stream(h, nullptr);
  return HandlePtr(h);
}
}  // namespace detail

// ---------------------------------------------------------------------------------------------
// vc::types::SquareUpperTriangularMatrix<Scalar,NP> stand-in (VisionCore, not in tree): packed upper
// triangle, row major.
// ---------------------------------------------------------------------------------------------
template <typename Scalar, int NP>
struct SquareUpperTriangularMatrix {
  static constexpr int Size = NP * (NP + 1) / 2;
  std::array<Scalar, Size> coeff_{};

  std::array<Scalar, Size>& coeff() { return coeff_; }
  const std::array<Scalar, Size>& coeff() const { return coeff_; }
  static constexpr int Index(int i, int j) { return i * NP - (i * (i - 1)) / 2 + (j - i); }
  // element of the full symmetric matrix (toDenseMatrix()(i,j) in the reference)
  Scalar toDenseMatrix(int i, int j) const { return i <= j ? coeff_[Index(i, j)] : coeff_[Index(j, i)]; }
  // full symmetric NP x NP, row major
  std::vector<Scalar> toDenseMatrix() const
  {
    std::vector<Scalar> m(static_cast<size_t>(NP) * NP);
    for (int i = 0; i < NP; ++i)
      for (int j = 0; j < NP; ++j) m[static_cast<size_t>(i) * NP + j] = toDenseMatrix(i, j);
    return m;
  }
  SquareUpperTriangularMatrix& operator+=(const SquareUpperTriangularMatrix& o)
  {
    for (int i = 0; i < Size; ++i) coeff_[i] += o.coeff_[i];
    return *this;
  }
};

// reduction_items.h:35-71
template <typename Scalar>
struct CorrespondenceReductionItem {
  CorrespondenceReductionItem() : residual(0), inliers(0) {}
  CorrespondenceReductionItem operator+(const CorrespondenceReductionItem& rhs) const
  {
    CorrespondenceReductionItem r(*this);
    r += rhs;
    return r;
  }
  CorrespondenceReductionItem& operator+=(const CorrespondenceReductionItem& rhs)
  {
    residual += rhs.residual;
    inliers += rhs.inliers;
    return *this;
  }
  Scalar residual;
  std::size_t inliers;
};

// reduction_items.h:77-143
template <typename Scalar, int NP>
struct JTJJrReductionItem {
  typedef SquareUpperTriangularMatrix<Scalar, NP> HessianType;
  typedef std::array<Scalar, NP> JacobianType;
  JTJJrReductionItem() : residual(0), inliers(0) { Jtr.fill(Scalar(0)); }
  JTJJrReductionItem operator+(const JTJJrReductionItem& rhs) const
  {
    JTJJrReductionItem r(*this);
    r += rhs;
    return r;
  }
  JTJJrReductionItem& operator+=(const JTJJrReductionItem& rhs)
  {
    JtJ += rhs.JtJ;
    for (int i = 0; i < NP; ++i) Jtr[i] += rhs.Jtr[i];
    residual += rhs.residual;
    inliers += rhs.inliers;
    return *this;
  }
  HessianType JtJ;
  JacobianType Jtr;
  Scalar residual;
  std::size_t inliers;
};

// dense_sfm.h:36-43
struct DenseSfmParams {
  float huber_delta = 0.1f;
  float ocl_th = 1000;  // disabled by default
  float avg_dpt = 2.0f;
  float min_dpt = 0.0f;
  int valid_border = 2;
};

// cu_sfmaligner.h:41-48
struct SfmAlignerParams {
  DenseSfmParams sfmparams;
  int step_threads = 32;  // NOTE: threads must be a multiple of 32!
  int step_blocks = 11;
  int eval_threads = 224;
  int eval_blocks = 66;
};

// ---------------------------------------------------------------------------------------------
// df::SfmAligner<Scalar,CS>  (cu_sfmaligner.h:50-97)
// ---------------------------------------------------------------------------------------------
template <typename Scalar, int CS>
class SfmAligner
{
  static_assert(sizeof(Scalar) == sizeof(float), "only float is instantiated (cu_sfmaligner.cpp:209)");

public:
  typedef std::shared_ptr<SfmAligner<Scalar, CS>> Ptr;
  typedef JTJJrReductionItem<Scalar, 12 + CS> ReductionItem;
  typedef CorrespondenceReductionItem<Scalar> ErrorReductionItem;

  explicit SfmAligner(SfmAlignerParams params = SfmAlignerParams()) : params_(params), h_(detail::MakeHandle())
  {
    Upload();
  }
  virtual ~SfmAligner() {}

  template <typename SE3T, typename CamT, typename ImageBuffer, typename GradBuffer>
  ErrorReductionItem EvaluateError(const SE3T& pose0, const SE3T& pose1, const CamT& cam, const ImageBuffer& img0,
                                   const ImageBuffer& img1, const ImageBuffer& dpt0, const ImageBuffer& std0,
                                   const GradBuffer& grad1)
  {
    const DfkCamera c = detail::Cam(cam);
    const DfkImage i0 = detail::View(img0, 1), i1 = detail::View(img1, 1), d0 = detail::View(dpt0, 1);
    const DfkImage s0 = detail::View(std0, 1), g1 = detail::View(grad1, 2);
    ErrorReductionItem r;
    uint64_t inl = 0;
    detail::Check(h_.get(), dfk_sfm_evaluate_error(h_.get(), pose0.data(), pose1.data(), &c, &i0, &i1, &d0, &s0, &g1,
                                                   &r.residual, &inl));
    r.inliers = static_cast<std::size_t>(inl);
    return r;
  }

  template <typename SE3T, typename CodeT, typename CamT, typename ImageBuffer, typename GradBuffer>
  ReductionItem RunStep(const SE3T& pose0, const SE3T& pose1, const CodeT& code0, const CamT& cam,
                        const ImageBuffer& img0, const ImageBuffer& img1, const ImageBuffer& dpt0,
                        const ImageBuffer& std0, ImageBuffer& valid0, const ImageBuffer& prx0_jac,
                        const GradBuffer& grad1)
  {
    const DfkCamera c = detail::Cam(cam);
    const DfkImage i0 = detail::View(img0, 1), i1 = detail::View(img1, 1), d0 = detail::View(dpt0, 1);
    const DfkImage s0 = detail::View(std0, 1), v0 = detail::View(valid0, 1), g1 = detail::View(grad1, 2);
    const DfkImage jc = detail::View(prx0_jac, CS);
    ReductionItem r;
    uint64_t inl = 0;
    detail::Check(h_.get(), dfk_sfm_run_step(h_.get(), pose0.data(), pose1.data(), code0.data(), CS, &c, &i0, &i1, &d0,
                                              &s0, &v0, &jc, &g1, r.JtJ.coeff().data(), r.Jtr.data(), &r.residual,
                                              &inl));
    r.inliers = static_cast<std::size_t>(inl);
    return r;
  }

  // UpdateDepth + RunStep in one launch (what PhotometricFactor does back to back: UpdateDepthMaps,
  // photometric_factor.cpp:331-341, then RunAlignmentStep :267-274): dpt0 is decoded from prx_orig0 + prx0_jac * code0
  // inside the kernel, written to dpt0 (an output here) and used for the warp; the code Jacobian is read once.
  template <typename SE3T, typename CodeT, typename CamT, typename ImageBuffer, typename GradBuffer>
  ReductionItem RunStepDecodeDepth(const SE3T& pose0, const SE3T& pose1, const CodeT& code0, const CamT& cam,
                                   const ImageBuffer& img0, const ImageBuffer& img1, const ImageBuffer& prx_orig0,
                                   ImageBuffer& dpt0, ImageBuffer& valid0, const ImageBuffer& prx0_jac,
                                   const GradBuffer& grad1)
  {
    DfkSfmWorkItem w{};
    for (int k = 0; k < 7; ++k) {
      w.pose0[k] = pose0.data()[k];
      w.pose1[k] = pose1.data()[k];
    }
    w.cam = detail::Cam(cam);
    w.img0 = detail::View(img0, 1); w.img1 = detail::View(img1, 1); w.dpt0 = detail::View(dpt0, 1);
    w.valid0 = detail::View(valid0, 1); w.prx0_jac = detail::View(prx0_jac, CS); w.grad1 = detail::View(grad1, 2);
    w.prx_orig = detail::View(prx_orig0, 1);
    w.code = code0.data();
    std::vector<float> rec(DFK_SFM_RECORD_FLOATS(CS));
    detail::Check(h_.get(), dfk_sfm_run_step_batch_host(h_.get(), &w, 1, CS, rec.data()));
    ReductionItem r;
    constexpr int NP = 12 + CS, NH = NP * (NP + 1) / 2;
    for (int k = 0; k < NH; ++k) r.JtJ.coeff().data()[k] = rec[k];
    for (int k = 0; k < NP; ++k) r.Jtr.data()[k] = rec[NH + k];
    r.residual = rec[NH + NP];
    uint32_t bits;
    std::memcpy(&bits, &rec[NH + NP + 1], 4);
    r.inliers = bits;
    return r;
  }

  void SetEvalThreadsBlocks(int threads, int blocks)
  {
    SfmAlignerParams p = params_;
    p.eval_threads = threads;
    p.eval_blocks = blocks;
    Upload(p);
  }
  void SetStepThreadsBlocks(int threads, int blocks)
  {
    SfmAlignerParams p = params_;
    p.step_threads = threads;
    p.step_blocks = blocks;
    Upload(p);
  }

  DfkHandle handle() const { return h_.get(); }  // extension: batched / asynchronous entry points of dfk.h
  // extension: launch on `stream` (a cudaStream_t; nullptr = the legacy default stream, the facade's default) or on the
  // handle's private non-blocking stream.  The caller then owns the ordering against its producers.
  void SetStream(void* stream) { detail::Check(h_.get(), dfk_set_stream(h_.get(), stream)); }
  void UseOwnStream() { detail::Check(h_.get(), dfk_use_own_stream(h_.get())); }

private:
  void Upload() { Upload(params_); }
  void Upload(const SfmAlignerParams& p)
  {
    DfkSfmAlignerParams c;
    c.sfmparams = DfkDenseSfmParams{p.sfmparams.huber_delta, p.sfmparams.ocl_th, p.sfmparams.avg_dpt,
                                    p.sfmparams.min_dpt, p.sfmparams.valid_border};
    c.step_threads = p.step_threads;
    c.step_blocks = p.step_blocks;
    c.eval_threads = p.eval_threads;
    c.eval_blocks = p.eval_blocks;
    detail::Check(h_.get(), dfk_sfm_set_params(h_.get(), &c));  // throws on threads % 32 != 0 / blocks > 1024
    params_ = p;
  }

  static const int max_blocks = 1024;
  SfmAlignerParams params_;
  detail::HandlePtr h_;
};

// ---------------------------------------------------------------------------------------------
// df::DepthAligner<Scalar, CS>  (cu_depthaligner.h:38-54)
// ---------------------------------------------------------------------------------------------
template <typename Scalar, int CS>
class DepthAligner
{
  static_assert(sizeof(Scalar) == sizeof(float), "only float is instantiated (cu_depthaligner.cpp:118)");

public:
  typedef std::shared_ptr<DepthAligner<Scalar, CS>> Ptr;
  typedef JTJJrReductionItem<Scalar, CS> ReductionItem;

  DepthAligner() : h_(detail::MakeHandle()) {}
  virtual ~DepthAligner() {}

  // cu_depthaligner.cpp:83-113; CodeT = Eigen::Matrix<Scalar,CS,1> (anything with data()), ImageBuffer = vc::Image2DView
  template <typename CodeT, typename ImageBuffer>
  ReductionItem RunStep(const CodeT& code, const ImageBuffer& target_dpt, const ImageBuffer& prx_orig,
                        const ImageBuffer& prx_jac)
  {
    const DfkImage t = detail::View(target_dpt, 1), p = detail::View(prx_orig, 1), j = detail::View(prx_jac, CS);
    ReductionItem r;
    uint64_t inl = 0;
    detail::Check(h_.get(), dfk_depth_run_step(h_.get(), code.data(), CS, &t, &p, &j, r.JtJ.coeff().data(), r.Jtr.data(),
                                               &r.residual, &inl));
    r.inliers = static_cast<std::size_t>(inl);
    return r;
  }
  DfkHandle handle() const { return h_.get(); }
  void SetStream(void* stream) { detail::Check(h_.get(), dfk_set_stream(h_.get(), stream)); }

private:
  detail::HandlePtr h_;
};

// ---------------------------------------------------------------------------------------------
// df::SE3Aligner<Scalar>  (cu_se3aligner.h:38-86)
// ---------------------------------------------------------------------------------------------
template <typename Scalar>
class SE3Aligner
{
  static_assert(sizeof(Scalar) == sizeof(float), "only float is instantiated (cu_se3aligner.cpp:179)");

public:
  typedef std::shared_ptr<SE3Aligner<Scalar>> Ptr;
  typedef JTJJrReductionItem<Scalar, 6> ReductionItem;
  typedef CorrespondenceReductionItem<Scalar> CorrespondenceItem;

  SE3Aligner() : h_(detail::MakeHandle()) {}
  virtual ~SE3Aligner() {}

  // Renders img0 from image img1 at pose T_01
  template <typename SE3T, typename CamT, typename ImageBuffer>
  CorrespondenceItem Warp(const SE3T& se3, const CamT& cam, const ImageBuffer& img0, const ImageBuffer& img1,
                          const ImageBuffer& dpt0, ImageBuffer& img2)
  {
    const DfkCamera c = detail::Cam(cam);
    const DfkImage i0 = detail::View(img0, 1), i1 = detail::View(img1, 1), d0 = detail::View(dpt0, 1);
    const DfkImage i2 = detail::View(img2, 1);
    CorrespondenceItem r;
    uint64_t inl = 0;
    detail::Check(h_.get(), dfk_se3_warp(h_.get(), se3.data(), &c, &i0, &i1, &d0, &i2, &r.residual, &inl));
    r.inliers = static_cast<std::size_t>(inl);
    return r;
  }

  template <typename SE3T, typename CamT, typename ImageBuffer, typename GradBuffer>
  ReductionItem RunStep(const SE3T& se3, const CamT& cam, const ImageBuffer& img0, const ImageBuffer& img1,
                        const ImageBuffer& dpt0, const GradBuffer& grad1)
  {
    const DfkCamera c = detail::Cam(cam);
    const DfkImage i0 = detail::View(img0, 1), i1 = detail::View(img1, 1), d0 = detail::View(dpt0, 1);
    const DfkImage g1 = detail::View(grad1, 2);
    ReductionItem r;
    uint64_t inl = 0;
    detail::Check(h_.get(), dfk_se3_run_step(h_.get(), se3.data(), &c, &i0, &i1, &d0, &g1, r.JtJ.coeff().data(),
                                              r.Jtr.data(), &r.residual, &inl));
    r.inliers = static_cast<std::size_t>(inl);
    return r;
  }

  DfkHandle handle() const { return h_.get(); }
  void SetStream(void* stream) { detail::Check(h_.get(), dfk_set_stream(h_.get(), stream)); }  // see SfmAligner::SetStream
  void UseOwnStream() { detail::Check(h_.get(), dfk_use_own_stream(h_.get())); }
  void SetHuberDelta(float val)
  {
    huber_delta_ = val;
    detail::Check(h_.get(), dfk_se3_set_huber_delta(h_.get(), val));
  }

  // The loop of CameraTracker::TrackFrame (core/system/camera_tracker.cpp:48-70) as ONE call: for level =
  // levels-1 .. 0, iterations_per_level[level] times { RunStep; update = -JtJ.ldlt().solve(Jtr); t += update.head<3>();
  // so3 = exp(update.tail<3>()) * so3 } with the solve and the retraction done on the device and a single read-back.
  // The pyramids are anything indexable by level (vc::RuntimeBufferPyramidManaged::operator[], std::vector of
  // views, SyncedBufferPyramid::GetGpuLevel results collected in a vector); camera_pyr[level] is the level's camera.
  // pose_ck is updated in place; returns {inliers / area, residual / inliers} of the last evaluated system, the two
  // numbers TrackFrame stores in inliers_ / error_ (:65-69).
  template <typename SE3T, typename CamPyr, typename ImagePyr0, typename ImagePyr1, typename DepthPyr, typename GradPyr>
  std::pair<float, float> TrackLevels(SE3T& pose_ck, const CamPyr& camera_pyr, const ImagePyr0& pyr_img0,
                                      const ImagePyr1& pyr_img1, const DepthPyr& pyr_dpt0, const GradPyr& pyr_grad1,
                                      const std::vector<int>& iterations_per_level)
  {
    std::vector<DfkTrackLevel> lv(iterations_per_level.size());
    for (std::size_t l = 0; l < lv.size(); ++l) {
      lv[l].cam = detail::Cam(camera_pyr[l]);
      lv[l].img0 = detail::View(pyr_img0[l], 1);
      lv[l].img1 = detail::View(pyr_img1[l], 1);
      lv[l].dpt0 = detail::View(pyr_dpt0[l], 1);
      lv[l].grad1 = detail::View(pyr_grad1[l], 2);
      lv[l].iterations = iterations_per_level[l];
    }
    float frac = 0.f, err = 0.f;
    detail::Check(h_.get(), dfk_se3_track(h_.get(), pose_ck.data(), lv.data(), static_cast<int>(lv.size()), &frac, &err,
                                          nullptr, nullptr, 0));
    return std::make_pair(frac, err);
  }

private:
  detail::HandlePtr h_;
  float huber_delta_ = 0.1f;
};

// ---------------------------------------------------------------------------------------------
// cu_image_proc.h:27-44 free functions.  They use one process-wide handle (default stream semantics
// of the reference); pass an explicit handle to run them on another stream.
// ---------------------------------------------------------------------------------------------
namespace detail
{
inline DfkHandle DefaultHandle()
{
  static HandlePtr h = MakeHandle();
  return h.get();
}
}  // namespace detail

template <typename ImageBuf, typename GradBuf>
void SobelGradients(const ImageBuf& img, GradBuf& grad)
{
  const DfkImage i = detail::View(img, 1), g = detail::View(grad, 2);
  detail::Check(detail::DefaultHandle(), dfk_sobel_gradients(detail::DefaultHandle(), &i, &g));
  detail::Check(detail::DefaultHandle(), dfk_synchronize(detail::DefaultHandle()));  // launch_utils.h:28
}

template <typename ImageBuf>
void GaussianBlurDown(const ImageBuf& in, ImageBuf& out)
{
  const DfkImage i = detail::View(in, 1), o = detail::View(out, 1);
  detail::Check(detail::DefaultHandle(), dfk_gaussian_blur_down(detail::DefaultHandle(), &i, &o));
  detail::Check(detail::DefaultHandle(), dfk_synchronize(detail::DefaultHandle()));
}

template <typename ImageBuf>
float SquaredError(const ImageBuf& buf1, const ImageBuf& buf2)
{
  const DfkImage a = detail::View(buf1, 1), b = detail::View(buf2, 1);
  float out = 0.f;
  detail::Check(detail::DefaultHandle(), dfk_squared_error(detail::DefaultHandle(), &a, &b, &out));
  return out;
}

template <typename CodeT, typename ImageBuf>
void UpdateDepth(const CodeT& code, const ImageBuf& prx_orig, const ImageBuf& prx_jac, float avg_dpt, ImageBuf& dpt_out)
{
  const int cs = static_cast<int>(code.size());
  const DfkImage p = detail::View(prx_orig, 1), d = detail::View(dpt_out, 1);
  const DfkImage j = detail::View(prx_jac, static_cast<unsigned>(cs));
  detail::Check(detail::DefaultHandle(),
                dfk_update_depth(detail::DefaultHandle(), code.data(), cs, &p, &j, avg_dpt, &d));
  detail::Check(detail::DefaultHandle(), dfk_synchronize(detail::DefaultHandle()));
}

}  // namespace df

#endif  // DFK_FACADE_H_
