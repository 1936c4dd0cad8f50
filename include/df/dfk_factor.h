// dfk_factor.h -- the consumer side of the hot path, in C++: what df::PhotometricFactor does with an aligner result
// and how a window of such results becomes one set of normal equations.  Host-only, header-only, no GTSAM / Eigen.
//
//   LinearizePhotometric   PhotometricFactor::linearize + RunAlignmentStep's post-processing
//                          (sources/core/gtsam/photometric_factor.cpp:84-181, 223-293): the residual rescale
//                          res / inliers * W * H (:275-282; +inf when there is no overlap), JtJ cast to double, Jtr negated
//                          (:105-106), and the slicing into the HessianFactor blocks G11 G12 G13 G22 G23 G33 / g1 g2 g3
//                          (:126-161) in the aligner's column order [pose0 | pose1 | code0].
//   WindowSystem           the block-sparse -> dense normal equations of a keyframe window: variables
//                          [pose_k (6) | code_k (C)] per keyframe; a pair (k0 -> k1) adds its pose0 / code0 blocks to
//                          keyframe k0's diagonal block, pose1 to k1's, and the pose0-pose1 / pose1-code0 couplings off the
//                          diagonal.  The buffer is what one NCCL all-reduce sums across ranks.
//   LinearizeReprojection / LinearizeSparseGeometric
//                          the Jacobian rows of the two sparse factors (reprojection_factor.cpp:157-269,
//                          sparse_geometric_factor.cpp:157-271), evaluated on the device from the keyframes' GPU buffers;
//                          the caller copies the row blocks into gtsam::VerticalBlockMatrix Ab(0..) as the reference does.
// deepfactors_b200/factors.py is the Python mirror; tests/cpp/factor_test.cpp checks the two against each other.
#ifndef DFK_FACTOR_H_
#define DFK_FACTOR_H_

#include <cstddef>
#include <limits>
#include <vector>

#include "dfk_facade.h"

namespace df
{

// row-major dense blocks, double like the reference's cast (photometric_factor.cpp:105)
template <int CS>
struct PhotometricBlocks {
  std::vector<double> G11, G12, G13, G22, G23, G33;  // 6x6, 6x6, 6xCS, 6x6, 6xCS, CSxCS
  std::vector<double> g1, g2, g3;                    // 6, 6, CS   (= -Jtr blocks)
  double f = 0.0;                                    // rescaled residual energy
};

template <int CS>
PhotometricBlocks<CS> LinearizePhotometric(const JTJJrReductionItem<float, 12 + CS>& sys, int width, int height)
{
  constexpr int NP = 12 + CS;
  PhotometricBlocks<CS> b;
  auto block = [&](int r0, int nr, int c0, int nc) {
    std::vector<double> m(static_cast<std::size_t>(nr) * nc);
    for (int r = 0; r < nr; ++r)
      for (int c = 0; c < nc; ++c) m[static_cast<std::size_t>(r) * nc + c] = static_cast<double>(sys.JtJ.toDenseMatrix(r0 + r, c0 + c));
    return m;
  };
  b.G11 = block(0, 6, 0, 6);
  b.G12 = block(0, 6, 6, 6);
  b.G13 = block(0, 6, 12, CS);
  b.G22 = block(6, 6, 6, 6);
  b.G23 = block(6, 6, 12, CS);
  b.G33 = block(12, CS, 12, CS);
  b.g1.resize(6);
  b.g2.resize(6);
  b.g3.resize(CS);
  for (int k = 0; k < 6; ++k) {
    b.g1[k] = -static_cast<double>(sys.Jtr[k]);
    b.g2[k] = -static_cast<double>(sys.Jtr[6 + k]);
  }
  for (int k = 0; k < CS; ++k) b.g3[k] = -static_cast<double>(sys.Jtr[12 + k]);
  static_assert(NP == 12 + CS, "column order [pose0 | pose1 | code0]");
  b.f = sys.inliers > 0 ? static_cast<double>(sys.residual) / static_cast<double>(sys.inliers) * width * height
                        : std::numeric_limits<double>::infinity();
  return b;
}

// Dense normal equations of a window of `num_keyframes` keyframes.
template <int CS>
class WindowSystem
{
public:
  static constexpr int Block = 6 + CS;
  explicit WindowSystem(int num_keyframes)
      : n_(num_keyframes), H_(static_cast<std::size_t>(dim()) * dim(), 0.0), g_(dim(), 0.0), f_(0.0)
  {
  }
  int dim() const { return n_ * Block; }
  double& H(int r, int c) { return H_[static_cast<std::size_t>(r) * dim() + c]; }
  double H(int r, int c) const { return H_[static_cast<std::size_t>(r) * dim() + c]; }
  std::vector<double>& H() { return H_; }
  std::vector<double>& g() { return g_; }
  const std::vector<double>& g() const { return g_; }
  double f() const { return f_; }

  // pair (k0 -> k1): keyframe k0 is warped into frame k1 (pose0 / code0 belong to k0, pose1 to k1)
  void Add(int k0, int k1, const JTJJrReductionItem<float, 12 + CS>& sys, int width, int height)
  {
    const int off[3] = {k0 * Block, k1 * Block, k0 * Block + 6};  // pose0, pose1, code0
    const int loc[3] = {0, 6, 12};
    const int len[3] = {6, 6, CS};
    for (int a = 0; a < 3; ++a) {
      for (int r = 0; r < len[a]; ++r) {
        g_[off[a] + r] -= static_cast<double>(sys.Jtr[loc[a] + r]);
        for (int b = 0; b < 3; ++b)
          for (int c = 0; c < len[b]; ++c)
            H(off[a] + r, off[b] + c) += static_cast<double>(sys.JtJ.toDenseMatrix(loc[a] + r, loc[b] + c));
      }
    }
    if (sys.inliers > 0) f_ += static_cast<double>(sys.residual) / static_cast<double>(sys.inliers) * width * height;
  }

private:
  int n_;
  std::vector<double> H_, g_;
  double f_;
};

// contiguous, balanced shard of the pair list for `rank` (sizes differ by at most one): pairs shard across GPUs with
// no data-path collective, the window buffers of the ranks are summed by one all-reduce
// Rows of a JacobianFactor, row-major; `width` floats per row, the last one is b.
struct SparseRows {
  std::vector<float> rows;
  int num_rows = 0;
  int width = 0;
  float total_err = 0.0f;  // ReprojectionFactor: sum of squared unweighted errors (total_err_, reprojection_factor.cpp:242)
  int num_valid = 0;       // SparseGeometricFactor: rows that are not all zero
  const float* row(int i) const { return rows.data() + static_cast<std::size_t>(i) * width; }
};

// ReprojectionFactor::linearize (reprojection_factor.cpp:157-269): 2 rows per match,
// [dErr/dPose0 (6) | dErr/dPose1 (6) | dErr/dCode0 (CS) | b].  query_xy / train_xy: 2 floats per match (host);
// prx_orig / prx_jac: the keyframe's level-0 GPU views (kf->pyr_prx_orig.GetGpuLevel(0), kf->pyr_jac.GetGpuLevel(0)).
template <int CS, typename SE3T, typename CodeT, typename CamT, typename ImageBuffer>
SparseRows LinearizeReprojection(DfkHandle h, const SE3T& pose0, const SE3T& pose1, const CodeT& code0, const CamT& cam,
                                 const ImageBuffer& prx_orig, const ImageBuffer& prx_jac, int num_matches,
                                 const float* query_xy, const float* train_xy, float huber_delta, float sigma)
{
  SparseRows out;
  out.num_rows = 2 * num_matches;
  out.width = 13 + CS;
  out.rows.assign(static_cast<std::size_t>(out.num_rows) * out.width, 0.0f);
  const DfkCamera c = detail::Cam(cam);
  const DfkImage p = detail::View(prx_orig, 1), j = detail::View(prx_jac, CS);
  detail::Check(h, dfk_reprojection_linearize(h, pose0.data(), pose1.data(), code0.data(), CS, &c, &p, &j, num_matches,
                                              query_xy, train_xy, huber_delta, sigma, out.rows.data(), &out.total_err));
  return out;
}

// SparseGeometricFactor::linearize (sparse_geometric_factor.cpp:157-271): 1 row per sampled point,
// [dErr/dPose0 (6) | dErr/dPose1 (6) | dErr/dCode0 (CS) | dErr/dCode1 (CS) | b].  points_xy: 2 ints per point (host);
// the image arguments are the two keyframes' level-0 GPU views and kf1's depth gradient (2 floats per pixel).
template <int CS, typename SE3T, typename CodeT, typename CamT, typename ImageBuffer, typename GradBuffer>
SparseRows LinearizeSparseGeometric(DfkHandle h, const SE3T& pose0, const SE3T& pose1, const CodeT& code0, const CodeT& code1,
                                    const CamT& cam, const ImageBuffer& prx0_orig, const ImageBuffer& prx0_jac,
                                    const ImageBuffer& prx1_orig, const ImageBuffer& prx1_jac, const GradBuffer& dpt_grad1,
                                    int num_points, const int* points_xy, float huber_delta)
{
  SparseRows out;
  out.num_rows = num_points;
  out.width = 13 + 2 * CS;
  out.rows.assign(static_cast<std::size_t>(out.num_rows) * out.width, 0.0f);
  const DfkCamera c = detail::Cam(cam);
  const DfkImage p0 = detail::View(prx0_orig, 1), j0 = detail::View(prx0_jac, CS), p1 = detail::View(prx1_orig, 1),
                 j1 = detail::View(prx1_jac, CS), g1 = detail::View(dpt_grad1, 2);
  detail::Check(h, dfk_sparse_geometric_linearize(h, pose0.data(), pose1.data(), code0.data(), code1.data(), CS, &c, &p0, &j0,
                                                  &p1, &j1, &g1, num_points, points_xy, huber_delta, out.rows.data(),
                                                  &out.num_valid));
  return out;
}

inline void ShardPairs(std::size_t num_pairs, int world_size, int rank, std::size_t* begin, std::size_t* end)
{
  *begin = (num_pairs * static_cast<std::size_t>(rank)) / static_cast<std::size_t>(world_size);
  *end = (num_pairs * static_cast<std::size_t>(rank + 1)) / static_cast<std::size_t>(world_size);
}

}  // namespace df

#endif  // DFK_FACTOR_H_
