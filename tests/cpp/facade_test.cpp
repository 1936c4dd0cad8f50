// facade_test.cpp -- the reference's own GPU-vs-CPU test, restated through the drop-in facade:
// tests/ut_sfmaligner.cpp:235-327 FullJacobianCompareWithCpu (GPU RunStep vs the host loop over
// df::DenseSfm; inliers exactly equal, dense H within tolerance) and tests/ut_se3aligner.cpp RunStep,
// on synthetic data (the reference's network outputs are not reproducible here).
// Build: see tests/cpp/Makefile.  Needs a GPU to run; compiling it is part of the CPU build check.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "df/dfk_facade.h"
#include "df/dfk_standins.h"
#include "dfk_oracle.h"

using namespace df::standin;

template <typename T>
struct DeviceImage {  // vc::Image2DManaged stand-in
  T* ptr = nullptr;
  size_t pitch = 0, w = 0, h = 0;
  DeviceImage(size_t w_, size_t h_) : w(w_), h(h_)
  {
    if (cudaMallocPitch((void**)&ptr, &pitch, w * sizeof(T), h) != cudaSuccess) { std::puts("cudaMallocPitch failed"); std::exit(2); }
    cudaMemset2D(ptr, pitch, 0, w * sizeof(T), h);
  }
  ~DeviceImage() { cudaFree(ptr); }
  void copyFrom(const T* host) { cudaMemcpy2D(ptr, pitch, host, w * sizeof(T), w * sizeof(T), h, cudaMemcpyHostToDevice); }
  Image2DView<T> view() { return Image2DView<T>(ptr, pitch, w, h); }
};

#define EXPECT(c)                                                        \
  do {                                                                   \
    if (!(c)) { std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #c); return 1; } \
  } while (0)

int main()
{
  constexpr int CS = 32;
  const int W = 320, H = 240;
  // synthetic inputs (same recipe as deepfactors_b200/synth.py, simplified)
  std::vector<float> img0(W * H), img1(W * H), dpt0(W * H), jac((size_t)W * H * CS), grad((size_t)W * H * 2), zeros(W * H, 0.f);
  unsigned s = 12345u;
  auto rnd = [&]() { s = s * 1664525u + 1013904223u; return ((s >> 8) & 0xffff) / 65535.0f - 0.5f; };
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x) {
      img0[y * W + x] = 0.5f + 0.25f * std::sin(x / 9.0f) * std::cos(y / 7.0f);
      img1[y * W + x] = 0.5f + 0.25f * std::sin(x / 9.0f + 0.4f) * std::cos(y / 7.0f - 0.2f);
      const float prx = 0.4f + 0.1f * std::sin(x / 20.0f) * std::cos(y / 25.0f);
      dpt0[y * W + x] = 2.0f / prx - 2.0f;
      for (int k = 0; k < CS; ++k) jac[((size_t)y * W + x) * CS + k] = 0.05f * rnd();
    }
  dfko_sobel_gradients_f(W, H, img1.data(), W, grad.data(), 2 * W);

  DeviceImage<float> d_img0(W, H), d_img1(W, H), d_dpt0(W, H), d_std0(W, H), d_vld0(W, H), d_jac((size_t)W * CS, H);
  DeviceImage<Grad> d_grad(W, H);
  d_img0.copyFrom(img0.data()); d_img1.copyFrom(img1.data()); d_dpt0.copyFrom(dpt0.data()); d_jac.copyFrom(jac.data());
  d_grad.copyFrom(reinterpret_cast<const Grad*>(grad.data()));

  // poses of ut_sfmaligner.cpp:254-264
  const float rot[3] = {0.1f, 0.1f, 0.0f}, trs[3] = {-0.5f, -0.5f, 0.0f};
  SE3 pose0, pose1 = SE3::FromRotTrs(rot, trs).inverse();
  Code<CS> code{};
  PinholeCamera cam(W / 2 / 0.5773502691896257f, H / 2 / 0.41421356237309503f, W / 2, H / 2, W, H);  // testing_utils.h:34-40

  df::SfmAlignerParams params;
  params.sfmparams.huber_delta = 0.5f;  // ut_sfmaligner.cpp:69
  df::SfmAligner<float, CS> aligner(params);
  auto i0 = d_img0.view(), i1 = d_img1.view(), dp = d_dpt0.view(), sd = d_std0.view(), vl = d_vld0.view(), jc = d_jac.view();
  auto gr = d_grad.view();
  auto gpu = aligner.RunStep(pose0, pose1, code, cam, i0, i1, dp, sd, vl, jc, gr);

  // CPU result: the oracle's DenseSfm loop in the reference order (x outer, y inner)
  constexpr int NP = 12 + CS;
  std::vector<double> H64(NP * (NP + 1) / 2), g64(NP);
  double res64 = 0;
  uint64_t inl64 = 0, inl32 = 0;
  std::vector<float> H32(NP * (NP + 1) / 2), g32(NP);
  float res32 = 0;
  DfkoCamera ocam{cam.fx(), cam.fy(), cam.u0(), cam.v0(), cam.width(), cam.height()};
  DfkoSfmParams oprm{0.5f, 1000.f, 2.0f, 0.0f, 2};
  dfko_sfm_run_step_f(pose0.data(), pose1.data(), CS, &ocam, W, H, img0.data(), W, img1.data(), W, dpt0.data(), W, nullptr, 0,
                      jac.data(), (size_t)W * CS, grad.data(), 2 * W, &oprm, 0, H32.data(), g32.data(), &res32, &inl32);
  dfko_sfm_run_step_d(pose0.data(), pose1.data(), CS, &ocam, W, H, img0.data(), W, img1.data(), W, dpt0.data(), W, nullptr, 0,
                      jac.data(), (size_t)W * CS, grad.data(), 2 * W, &oprm, 0, H64.data(), g64.data(), &res64, &inl64);
  EXPECT(gpu.inliers == inl32);  // ut_sfmaligner.cpp:320
  EXPECT(gpu.inliers != 0);
  double scale = 0, err = 0;
  for (size_t k = 0; k < H64.size(); ++k) scale = std::fmax(scale, std::fabs(H64[k]));
  for (size_t k = 0; k < H64.size(); ++k) err = std::fmax(err, std::fabs(gpu.JtJ.coeff()[k] - H64[k]));
  std::printf("SfmAligner::RunStep inliers=%zu  max|H_gpu-H_f64|/max|H| = %.3e\n", gpu.inliers, err / scale);
  EXPECT(inl64 != inl32 || err <= 2e-5 * scale);
  EXPECT(gpu.JtJ.toDenseMatrix(3, 20) == gpu.JtJ.toDenseMatrix(20, 3));

  auto ev = aligner.EvaluateError(pose0, pose1, cam, i0, i1, dp, sd, gr);
  EXPECT(ev.inliers >= gpu.inliers);

  df::SE3Aligner<float> se3;
  auto r6 = se3.RunStep(pose1, cam, i0, i1, dp, gr);
  double J6[21], g6[6], res6;
  uint64_t inl6;
  dfko_se3_run_step_d(pose1.data(), &ocam, W, H, img0.data(), W, img1.data(), W, dpt0.data(), W, grad.data(), 2 * W, 0.1f, J6, g6,
                      &res6, &inl6);
  float J6f[21], g6f[6], res6f;
  uint64_t inl6f;
  dfko_se3_run_step_f(pose1.data(), &ocam, W, H, img0.data(), W, img1.data(), W, dpt0.data(), W, grad.data(), 2 * W, 0.1f, J6f, g6f,
                      &res6f, &inl6f);
  EXPECT(r6.inliers == inl6f);

  // fused depth decode: with a zero code the decoded depth is avg/prx_orig - avg == dpt0, so the fused call must
  // reproduce the plain RunStep (same inliers; same sums up to the 1-ulp difference between the two depth formulas)
  {
    std::vector<float> prx(W * H);
    for (int i = 0; i < W * H; ++i) prx[i] = 2.0f / (dpt0[i] + 2.0f);
    DeviceImage<float> d_prx(W, H), d_dout(W, H), d_vld2(W, H);
    d_prx.copyFrom(prx.data());
    auto po = d_prx.view(), dout = d_dout.view(), v2 = d_vld2.view();
    auto fused = aligner.RunStepDecodeDepth(pose0, pose1, code, cam, i0, i1, po, dout, v2, jc, gr);
    EXPECT(std::llabs((long long)fused.inliers - (long long)gpu.inliers) <= 8);
    EXPECT(std::fabs(fused.residual - gpu.residual) <= 1e-3f * gpu.residual);
  }

  // CameraTracker::TrackFrame's loop as one device-side call: a single level, 8 Gauss-Newton iterations starting at
  // pose1 must not increase the mean Huber-weighted residual, and moves the pose
  {
    SE3 trk = pose1;
    std::vector<PinholeCamera> cams(1, cam);
    std::vector<Image2DView<float>> k0(1, i0), f1(1, i1), kd(1, dp);
    std::vector<Image2DView<Grad>> g1(1, gr);
    const auto stats = se3.TrackLevels(trk, cams, k0, f1, kd, g1, std::vector<int>(1, 8));
    EXPECT(stats.first > 0.3f && stats.first <= 1.0f);
    EXPECT(stats.second <= r6.residual / (float)r6.inliers * 1.0001f);
    float moved = 0.f;
    for (int k = 0; k < 7; ++k) moved += std::fabs(trk.data()[k] - pose1.data()[k]);
    EXPECT(moved > 1e-4f);
  }

  bool threw = false;
  try {
    aligner.SetStepThreadsBlocks(33, 11);  // CHECK_EQ(threads % 32, 0) in the reference
  } catch (const std::exception&) {
    threw = true;
  }
  EXPECT(threw);
  std::puts("FACADE_TEST_OK");
  return 0;
}
