// dfk_standins.h -- minimal stand-ins for the third-party types of the reference's aligner API, for
// builds that do not have Sophus / Eigen / VisionCore (this repository does not).  They expose exactly
// the members df/dfk_facade.h uses, with the semantics of the real types:
//   df::standin::SE3            Sophus::SE3f        data() -> quaternion (x,y,z,w), translation (x,y,z)
//   df::standin::Code<CS>       Eigen::Matrix<float,CS,1>   data(), size()
//   df::standin::PinholeCamera  df::PinholeCamera<float> (sources/common/algorithm/pinhole_camera.h:43)
//   df::standin::Image2DView<T> vc::Image2DView<T, TargetDeviceCUDA>   ptr(), pitch() [bytes], width(), height()
#ifndef DFK_STANDINS_H_
#define DFK_STANDINS_H_

#include <array>
#include <cmath>
#include <cstddef>

namespace df
{
namespace standin
{

struct SE3 {
  float d[7] = {0, 0, 0, 1, 0, 0, 0};
  const float* data() const { return d; }
  float* data() { return d; }
  // Sophus::SE3f(SO3f::exp(omega), trs)
  static SE3 FromRotTrs(const float omega[3], const float trs[3])
  {
    SE3 p;
    const float th2 = omega[0] * omega[0] + omega[1] * omega[1] + omega[2] * omega[2];
    const float th = std::sqrt(th2);
    float imag, real;
    if (th < 1e-10f) {
      imag = 0.5f - th2 / 48.0f;
      real = 1.0f - th2 / 8.0f;
    } else {
      imag = std::sin(0.5f * th) / th;
      real = std::cos(0.5f * th);
    }
    p.d[0] = imag * omega[0]; p.d[1] = imag * omega[1]; p.d[2] = imag * omega[2]; p.d[3] = real;
    p.d[4] = trs[0]; p.d[5] = trs[1]; p.d[6] = trs[2];
    return p;
  }
  SE3 inverse() const
  {
    SE3 r;
    const float q[4] = {-d[0], -d[1], -d[2], d[3]};
    const float v[3] = {-d[4], -d[5], -d[6]};
    float uv[3] = {q[1] * v[2] - q[2] * v[1], q[2] * v[0] - q[0] * v[2], q[0] * v[1] - q[1] * v[0]};
    for (float& u : uv) u += u;
    r.d[0] = q[0]; r.d[1] = q[1]; r.d[2] = q[2]; r.d[3] = q[3];
    r.d[4] = v[0] + q[3] * uv[0] + (q[1] * uv[2] - q[2] * uv[1]);
    r.d[5] = v[1] + q[3] * uv[1] + (q[2] * uv[0] - q[0] * uv[2]);
    r.d[6] = v[2] + q[3] * uv[2] + (q[0] * uv[1] - q[1] * uv[0]);
    return r;
  }
};

template <int CS>
using Code = std::array<float, CS>;

class PinholeCamera
{
public:
  PinholeCamera(float fx, float fy, float u0, float v0, float width, float height)
      : fx_(fx), fy_(fy), u0_(u0), v0_(v0), width_(width), height_(height) {}
  float fx() const { return fx_; }
  float fy() const { return fy_; }
  float u0() const { return u0_; }
  float v0() const { return v0_; }
  float width() const { return width_; }
  float height() const { return height_; }

private:
  float fx_, fy_, u0_, v0_, width_, height_;
};

// non-owning pitched 2-D view; T = float, or a 2-float pixel for gradients
template <typename T>
class Image2DView
{
public:
  Image2DView() = default;
  Image2DView(T* ptr, std::size_t pitch_bytes, std::size_t width, std::size_t height)
      : ptr_(ptr), pitch_(pitch_bytes), width_(width), height_(height) {}
  T* ptr() { return ptr_; }
  const T* ptr() const { return ptr_; }
  std::size_t pitch() const { return pitch_; }
  std::size_t width() const { return width_; }
  std::size_t height() const { return height_; }
  std::size_t area() const { return width_ * height_; }

private:
  T* ptr_ = nullptr;
  std::size_t pitch_ = 0, width_ = 0, height_ = 0;
};

struct Grad {
  float gx, gy;
};

}  // namespace standin
}  // namespace df

#endif  // DFK_STANDINS_H_
