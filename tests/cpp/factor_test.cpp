// factor_test.cpp -- host-only check of include/df/dfk_factor.h (PhotometricFactor block slicing + window assembly,
// photometric_factor.cpp:84-181,275-282).  Fills a few aligner results from a fixed LCG, prints the assembled window
// and the blocks of the first result as JSON; tests/test_factors_cpp.py recomputes both with deepfactors_b200/factors.py.
#include <cstdint>
#include <cstdio>
#include <vector>

#include "df/dfk_factor.h"

// Compile-time check of the sparse-factor wrappers (they need a device to RUN: tests/test_gpu_parity.py drives the same
// C entry points): stand-ins with the accessor names of vc::Image2DView / df::PinholeCamera / Sophus::SE3f / Eigen vectors.
namespace
{
struct ViewStandIn {
  const float* ptr() const { return nullptr; }
  std::size_t pitch() const { return 0; }
  std::size_t width() const { return 0; }
  std::size_t height() const { return 0; }
};
struct CamStandIn {
  float fx() const { return 1; }
  float fy() const { return 1; }
  float u0() const { return 0; }
  float v0() const { return 0; }
  float width() const { return 0; }
  float height() const { return 0; }
};
struct VecStandIn {
  const float* data() const { return nullptr; }
};
using ReprojFn = df::SparseRows (*)(DfkHandle, const VecStandIn&, const VecStandIn&, const VecStandIn&, const CamStandIn&,
                                    const ViewStandIn&, const ViewStandIn&, int, const float*, const float*, float, float);
using GeomFn = df::SparseRows (*)(DfkHandle, const VecStandIn&, const VecStandIn&, const VecStandIn&, const VecStandIn&,
                                  const CamStandIn&, const ViewStandIn&, const ViewStandIn&, const ViewStandIn&,
                                  const ViewStandIn&, const ViewStandIn&, int, const int*, float);
volatile ReprojFn g_reproj = &df::LinearizeReprojection<8, VecStandIn, VecStandIn, CamStandIn, ViewStandIn>;
volatile GeomFn g_geom = &df::LinearizeSparseGeometric<8, VecStandIn, VecStandIn, CamStandIn, ViewStandIn, ViewStandIn>;
}  // namespace

int main()
{
  constexpr int CS = 8, NP = 12 + CS;
  uint32_t s = 2024u;
  auto rnd = [&]() {
    s = s * 1664525u + 1013904223u;
    return static_cast<float>((s >> 8) & 0xffff) / 65535.0f - 0.5f;
  };
  const int pairs[5][2] = {{0, 1}, {1, 2}, {2, 0}, {0, 2}, {1, 0}};
  const int sizes[5][2] = {{640, 480}, {320, 240}, {160, 120}, {80, 60}, {640, 480}};
  std::vector<df::JTJJrReductionItem<float, NP>> items(5);
  for (int i = 0; i < 5; ++i) {
    for (auto& c : items[i].JtJ.coeff()) c = rnd();
    for (auto& c : items[i].Jtr) c = rnd();
    items[i].residual = rnd() + 1.0f;
    items[i].inliers = (i == 3) ? 0 : static_cast<std::size_t>(1000 + 17 * i);  // one result without overlap
  }
  df::WindowSystem<CS> win(3);
  for (int i = 0; i < 5; ++i) win.Add(pairs[i][0], pairs[i][1], items[i], sizes[i][0], sizes[i][1]);
  const auto blk = df::LinearizePhotometric<CS>(items[0], 640, 480);
  const auto none = df::LinearizePhotometric<CS>(items[3], 80, 60);

  std::printf("{\"dim\": %d, \"f\": %.17g, \"no_overlap_is_inf\": %s, \"f0\": %.17g,\n \"H\": [", win.dim(), win.f(),
              (none.f > 1e300) ? "true" : "false", blk.f);
  for (int r = 0; r < win.dim(); ++r)
    for (int c = 0; c < win.dim(); ++c) std::printf("%s%.17g", (r || c) ? "," : "", win.H(r, c));
  std::printf("],\n \"g\": [");
  for (int r = 0; r < win.dim(); ++r) std::printf("%s%.17g", r ? "," : "", win.g()[r]);
  auto dump = [](const char* name, const std::vector<double>& v) {
    std::printf("],\n \"%s\": [", name);
    for (std::size_t k = 0; k < v.size(); ++k) std::printf("%s%.17g", k ? "," : "", v[k]);
  };
  dump("G11", blk.G11); dump("G12", blk.G12); dump("G13", blk.G13); dump("G22", blk.G22); dump("G23", blk.G23);
  dump("G33", blk.G33); dump("g1", blk.g1); dump("g2", blk.g2); dump("g3", blk.g3);
  std::printf("]}\n");
  std::size_t b0, e0, b1, e1;
  df::ShardPairs(5, 2, 0, &b0, &e0);
  df::ShardPairs(5, 2, 1, &b1, &e1);
  return (b0 == 0 && e0 == b1 && e1 == 5) ? 0 : 1;
}
