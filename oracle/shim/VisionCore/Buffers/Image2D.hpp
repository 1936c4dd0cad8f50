// oracle/shim/VisionCore/Buffers/Image2D.hpp -- stand-in (TEST INFRASTRUCTURE ONLY) for vc::Image2DView<T, Target>:
// a non-owning pitched 2-D view.  Restated semantics (VisionCore Buffer2D.hpp, not in this image):
//   operator()(x, y)           element x of row y, rows `pitch` BYTES apart
//   getBilinear<TR>(pix)       ix = floor(u), iy = floor(v), fx = u - ix, fy = v - iy;
//                              lerp( lerp(row(iy)[ix], row(iy)[ix+1], fx), lerp(row(iy+1)[ix], row(iy+1)[ix+1], fx), fy )
//                              with lerp(a, b, t) = a + t * (b - a); integer coordinates at pixel centres, no clamping
//                              (call sites dense_sfm.h:95,103,167,180; lucas_kanade_se3.h:57,61)
#ifndef DFK_SHIM_VC_IMAGE2D_
#define DFK_SHIM_VC_IMAGE2D_

#include <Eigen/Core>

#include <cmath>
#include <cstddef>

#include "../Platform.hpp"

namespace vc {

template <typename T, typename Target>
class Image2DView {
 public:
  using ValueType = T;
  Image2DView() : ptr_(nullptr), w_(0), h_(0), pitch_(0) {}
  Image2DView(T* ptr, std::size_t w, std::size_t h, std::size_t pitch_bytes) : ptr_(ptr), w_(w), h_(h), pitch_(pitch_bytes) {}

  std::size_t width() const { return w_; }
  std::size_t height() const { return h_; }
  std::size_t pitch() const { return pitch_; }
  std::size_t area() const { return w_ * h_; }
  T* ptr() { return ptr_; }
  const T* ptr() const { return ptr_; }
  T* rowPtr(std::size_t y) { return reinterpret_cast<T*>(reinterpret_cast<unsigned char*>(ptr_) + y * pitch_); }
  const T* rowPtr(std::size_t y) const
  {
    return reinterpret_cast<const T*>(reinterpret_cast<const unsigned char*>(ptr_) + y * pitch_);
  }
  T& operator()(std::size_t x, std::size_t y) { return rowPtr(y)[x]; }
  const T& operator()(std::size_t x, std::size_t y) const { return rowPtr(y)[x]; }

  template <typename TR, typename S>
  TR getBilinear(const Eigen::Matrix<S, 2, 1>& p) const
  {
    return getBilinear<TR>(p(0), p(1));
  }
  template <typename TR>
  TR getBilinear(float u, float v) const
  {
    const float ix = floorf(u);
    const float iy = floorf(v);
    const float fx = u - ix;
    const float fy = v - iy;
    const T* bl = rowPtr((std::size_t)iy) + (std::size_t)ix;
    const T* tl = rowPtr((std::size_t)iy + 1) + (std::size_t)ix;
    return lerp<TR>(lerp<TR>(bl[0], bl[1], fx), lerp<TR>(tl[0], tl[1], fx), fy);
  }

 private:
  template <typename TR, typename A>
  static TR lerp(const A& a, const A& b, float t)
  {
    return TR(a + (b - a) * t);
  }
  T* ptr_;
  std::size_t w_, h_, pitch_;
};

template <typename T, typename Target>
using Buffer2DView = Image2DView<T, Target>;

// only named (never instantiated) on the path: default template argument of RenderDpt (warping.h:71)
template <typename T, typename Target>
class Image2DManaged {
 public:
  using ViewT = Image2DView<T, Target>;
};

}  // namespace vc

#endif
