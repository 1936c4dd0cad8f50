// oracle/shim/sophus/se3.hpp -- minimal stand-in for Sophus::SO3 / Sophus::SE3 (strasdat/Sophus @ d0b7315, not in this
// image), the subset the reference's hot-path headers use.  TEST INFRASTRUCTURE ONLY (see Eigen/Core here).
// Restated semantics:
//   SO3 holds a unit quaternion (x, y, z, w); SO3 * p = Eigen::QuaternionBase::_transformVector(p):
//       uv = q.vec x p;  uv += uv;  p + w*uv + q.vec x uv          (this decides pixel validity bit for bit)
//   SO3 * SO3 = quaternion product (Eigen's operator*, "a * b" with a applied last), re-normalised only in exp/ctor
//   SO3::matrix() = Eigen::QuaternionBase::toRotationMatrix()
//   SO3::inverse() = conjugate;  SE3::inverse() = (R^-1, R^-1 * (-t));  SE3 * SE3 = (Ra*Rb, ta + Ra*tb)
//   SO3::exp(w) = quaternion (sin(theta/2)/theta * w, cos(theta/2)) with the small-angle Taylor branch
//   SE3 data layout: quaternion (x,y,z,w) then translation -- Sophus::SE3f::data() order, what the C ABI passes.
#ifndef DFK_SHIM_SOPHUS_SE3_
#define DFK_SHIM_SOPHUS_SE3_

#include <Eigen/Core>

#include <cmath>

namespace Sophus {

template <typename T>
class SO3 {
 public:
  using Point = Eigen::Matrix<T, 3, 1>;
  using Transformation = Eigen::Matrix<T, 3, 3>;
  static constexpr int DoF = 3;

  SO3() : x_(0), y_(0), z_(0), w_(1) {}
  SO3(T x, T y, T z, T w) : x_(x), y_(y), z_(z), w_(w) {}

  T x() const { return x_; }
  T y() const { return y_; }
  T z() const { return z_; }
  T w() const { return w_; }

  static Transformation hat(const Point& o)
  {
    Transformation m;
    m << T(0), -o[2], o[1], o[2], T(0), -o[0], -o[1], o[0], T(0);
    return m;
  }

  static SO3 exp(const Point& omega)
  {
    const T theta_sq = omega[0] * omega[0] + omega[1] * omega[1] + omega[2] * omega[2];
    const T theta = std::sqrt(theta_sq);
    const T half_theta = T(0.5) * theta;
    T imag, real;
    if (theta < T(1e-10)) {  // Sophus::Constants<T>::epsilon()
      const T theta_po4 = theta_sq * theta_sq;
      imag = T(0.5) - T(1.0 / 48.0) * theta_sq + T(1.0 / 3840.0) * theta_po4;
      real = T(1) - T(0.5) * theta_sq + T(1.0 / 384.0) * theta_po4;
    } else {
      imag = std::sin(half_theta) / theta;
      real = std::cos(half_theta);
    }
    return SO3(imag * omega[0], imag * omega[1], imag * omega[2], real);
  }

  SO3 inverse() const { return SO3(-x_, -y_, -z_, w_); }

  // Eigen quaternion product
  SO3 operator*(const SO3& b) const
  {
    const SO3& a = *this;
    return SO3(a.w_ * b.x_ + a.x_ * b.w_ + a.y_ * b.z_ - a.z_ * b.y_,
               a.w_ * b.y_ + a.y_ * b.w_ + a.z_ * b.x_ - a.x_ * b.z_,
               a.w_ * b.z_ + a.z_ * b.w_ + a.x_ * b.y_ - a.y_ * b.x_,
               a.w_ * b.w_ - a.x_ * b.x_ - a.y_ * b.y_ - a.z_ * b.z_);
  }

  // Eigen::QuaternionBase::_transformVector
  Point operator*(const Point& v) const
  {
    const Point qv(x_, y_, z_);
    Point uv = qv.cross(v);
    uv += uv;
    return v + uv * w_ + qv.cross(uv);
  }

  // Eigen::QuaternionBase::toRotationMatrix
  Transformation matrix() const
  {
    Transformation res;
    const T tx = T(2) * x_, ty = T(2) * y_, tz = T(2) * z_;
    const T twx = tx * w_, twy = ty * w_, twz = tz * w_;
    const T txx = tx * x_, txy = ty * x_, txz = tz * x_;
    const T tyy = ty * y_, tyz = tz * y_, tzz = tz * z_;
    res(0, 0) = T(1) - (tyy + tzz);
    res(0, 1) = txy - twz;
    res(0, 2) = txz + twy;
    res(1, 0) = txy + twz;
    res(1, 1) = T(1) - (txx + tzz);
    res(1, 2) = tyz - twx;
    res(2, 0) = txz - twy;
    res(2, 1) = tyz + twx;
    res(2, 2) = T(1) - (txx + tyy);
    return res;
  }

 private:
  T x_, y_, z_, w_;
};

template <typename T>
class SE3 {
 public:
  using SO3Type = SO3<T>;
  using Point = Eigen::Matrix<T, 3, 1>;
  static constexpr int DoF = 6;

  SE3() : so3_(), t_(Point::Zero()) {}
  SE3(const SO3<T>& so3, const Point& t) : so3_(so3), t_(t) {}

  SO3<T>& so3() { return so3_; }
  const SO3<T>& so3() const { return so3_; }
  Point& translation() { return t_; }
  const Point& translation() const { return t_; }

  SE3 inverse() const
  {
    const SO3<T> inv = so3_.inverse();
    return SE3(inv, inv * (t_ * T(-1)));
  }
  SE3 operator*(const SE3& o) const { return SE3(so3_ * o.so3_, t_ + so3_ * o.t_); }
  Point operator*(const Point& p) const { return so3_ * p + t_; }

 private:
  SO3<T> so3_;
  Point t_;
};

using SE3f = SE3<float>;
using SE3d = SE3<double>;
using SO3f = SO3<float>;
using SO3d = SO3<double>;

}  // namespace Sophus

#endif
