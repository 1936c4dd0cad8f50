// oracle/shim/VisionCore/Types/SquareUpperTriangularMatrix.hpp -- stand-in (TEST INFRASTRUCTURE ONLY).
// vc::types::SquareUpperTriangularMatrix<T, N>: the N(N+1)/2 coefficients of the upper triangle, row by row
// ((0,0) (0,1) ... (0,N-1) (1,1) ...), held in an Eigen column vector `coeff()`; constructing it from a vector v gives
// the upper triangle of v v^T; toDenseMatrix() mirrors it into a full symmetric matrix
// (call sites: sources/cuda/reduction_items.h:80,113-116; dense_sfm.h:199; lucas_kanade_se3.h:72;
//  core/gtsam/photometric_factor.cpp:105).
#ifndef DFK_SHIM_VC_SUTM_
#define DFK_SHIM_VC_SUTM_

#include <Eigen/Core>

namespace vc {
namespace types {

template <typename T, int N>
class SquareUpperTriangularMatrix {
 public:
  static constexpr int NumCoeffs = N * (N + 1) / 2;
  using CoeffType = Eigen::Matrix<T, NumCoeffs, 1>;
  using DenseMatrixType = Eigen::Matrix<T, N, N>;

  SquareUpperTriangularMatrix() {}
  explicit SquareUpperTriangularMatrix(const Eigen::Matrix<T, N, 1>& v)
  {
    int k = 0;
    for (int i = 0; i < N; ++i)
      for (int j = i; j < N; ++j) c_(k++) = v(i) * v(j);
  }
  static SquareUpperTriangularMatrix Zero()
  {
    SquareUpperTriangularMatrix m;
    m.c_ = CoeffType::Zero();
    return m;
  }
  CoeffType& coeff() { return c_; }
  const CoeffType& coeff() const { return c_; }
  SquareUpperTriangularMatrix& operator+=(const SquareUpperTriangularMatrix& o)
  {
    c_ += o.c_;
    return *this;
  }
  SquareUpperTriangularMatrix operator+(const SquareUpperTriangularMatrix& o) const
  {
    SquareUpperTriangularMatrix r = *this;
    r += o;
    return r;
  }
  DenseMatrixType toDenseMatrix() const
  {
    DenseMatrixType m;
    int k = 0;
    for (int i = 0; i < N; ++i)
      for (int j = i; j < N; ++j) {
        m(i, j) = c_(k);
        m(j, i) = c_(k);
        ++k;
      }
    return m;
  }

 private:
  CoeffType c_;
};

}  // namespace types
}  // namespace vc

#endif
