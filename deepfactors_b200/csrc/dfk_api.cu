// dfk_api.cu -- C ABI of libdfk.so (see include/dfk.h): argument validation, host-side SE3
// algebra (relative pose + Jacobians, as the reference does on the host in
// cu_sfmaligner.cpp:164-166), launch planning, scratch ownership, error reporting.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <new>
#include <string>
#include <vector>

#include "dfk.h"
#include "dfk_internal.h"

using namespace dfk;

struct DfkContext {
  int device = 0;
  int num_sms = 1;
  int sm_limit = 0;  // dfk_set_sm_limit: SMs the persistent step kernels may occupy (0 = all)
  cudaStream_t own_stream = nullptr;
  cudaStream_t stream = nullptr;
  std::string err;
  DfkSfmAlignerParams params;
  DfkGramMode gram_mode = DFK_GRAM_AUTO;
  float se3_huber_delta = 0.1f;  // cu_se3aligner.h:85

  float* simple_scratch = nullptr;   // kSimpleScratchFloats
  unsigned int* counter = nullptr;   // 1 (self-resetting ticket)
  float* out_dev = nullptr;          // 32 floats
  float* out_host = nullptr;         // pinned, 32 floats
  float* code_dev = nullptr;         // 256 floats
  float* track_dev = nullptr;        // dfk_se3_track: [pose 8 | per-iteration history 36 each]
  size_t track_cap = 0;              // floats
  float* track_host = nullptr;       // pinned mirror of track_dev + the last system (32)
  size_t track_host_cap = 0;

  float* sparse_dev = nullptr;       // dfk_reprojection_linearize: [query | train | rows | err2]
  size_t sparse_cap = 0;
  float* sparse_host = nullptr;      // pinned mirror
  size_t sparse_host_cap = 0;
  SfmItemDev* items_dev = nullptr;
  size_t items_cap = 0;
  float* partials_dev = nullptr;
  size_t partials_cap = 0;  // floats
  // normalised ray tables of the tensor-core kernel: they depend on (fx, u0, width, fy, v0, height) only, so they are
  // built once per camera level and reused by every later call (one launch less per evaluation in steady state)
  struct RayTab {
    float fx, fy, u0, v0;
    uint32_t w, h;
    float* dev;
  };
  std::vector<RayTab> ray_cache;
  bool ray_miss = false;  // build_items found a camera without a table: run the table kernel this call
  float* codes_dev = nullptr;  // fused depth decode: code_size floats per work item
  size_t codes_cap = 0;
  std::vector<float> codes_host;
  float* records_dev = nullptr;
  size_t records_cap = 0;  // floats
  float* records_host = nullptr;  // pinned
  size_t records_host_cap = 0;
  std::vector<SfmItemDev> items_host;

  // measurement hooks (dfk_set_profiling / dfk_get_profile)
  bool profiling = false;
  std::vector<cudaEvent_t> ev_pool;  // pairs: [2k] start, [2k+1] stop
  size_t ev_used = 0;                // number of pairs recorded since the last read
  double ev_ms_accum = 0.0;          // time of pairs already drained
  uint64_t ev_count_accum = 0;
  uint64_t launches = 0;
};

// pipelined host -> device -> host evaluation (dfk_sfm_stream_*)
struct DfkSfmStream {
  int device = 0, code_size = 0, max_items = 0, depth = 0;
  size_t max_bytes = 0;
  cudaStream_t copy_stream = nullptr;
  struct Slot {
    unsigned char* dev = nullptr;   // staged inputs (+ valid0 / decoded-depth scratch)
    size_t cap = 0;
    float* rec_dev = nullptr;
    float* rec_host = nullptr;      // pinned
    cudaEvent_t uploaded = nullptr, done = nullptr;
    int n = 0;
    uint64_t ticket = 0;
    bool busy = false;
  };
  std::vector<Slot> slots;
  std::vector<DfkSfmWorkItem> dev_items;  // scratch of submit()
  uint64_t next_ticket = 0, next_wait = 0;
};

// CSR adjacency of a keyframe window on the device (dfk_window_create)
struct DfkWindow {
  int device = 0;
  WindowDev dev{};
  int* ints = nullptr;      // one allocation: kf0_ptr | kf0_items | kf1_ptr | kf1_items | pair_ptr | pair_items
  float* areas = nullptr;
  size_t floats = 0;
};

namespace {

DfkStatus fail(DfkHandle h, DfkStatus s, const std::string& msg)
{
  if (h) h->err = msg;
  return s;
}

// out-of-memory exit of an extern "C" entry point (never throws itself)
DfkStatus oom(DfkHandle h) noexcept
{
  if (h) {
    try {
      h->err = "out of host memory";
    } catch (...) {
    }
  }
  return DFK_ERR_NOMEM;
}

DfkStatus cuda_fail(DfkHandle h, cudaError_t e, const char* what)
{
  // message format of vc::CUDAException thrown from CudaCheckLastError (launch_utils.h:26-32)
  std::string m = std::string(what) + ": " + cudaGetErrorString(e);
  cudaGetLastError();  // clear sticky-less errors
  return fail(h, DFK_ERR_CUDA, m);
}

#define DFK_CUDA(h, call, what)                            \
  do {                                                     \
    cudaError_t e__ = (call);                              \
    if (e__ != cudaSuccess) return cuda_fail(h, e__, what); \
  } while (0)

struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev)
  {
    cudaGetDevice(&prev);
    if (prev != dev) cudaSetDevice(dev);
    else prev = -1;
  }
  ~DeviceGuard()
  {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

// ---------------------------------------------------------------------------- SE3 algebra (fp32)
// Eigen QuaternionBase::_transformVector
void quat_rotate(const float q[4], const float v[3], float out[3])
{
  float uv0 = q[1] * v[2] - q[2] * v[1];
  float uv1 = q[2] * v[0] - q[0] * v[2];
  float uv2 = q[0] * v[1] - q[1] * v[0];
  uv0 += uv0; uv1 += uv1; uv2 += uv2;
  out[0] = (v[0] + q[3] * uv0) + (q[1] * uv2 - q[2] * uv1);
  out[1] = (v[1] + q[3] * uv1) + (q[2] * uv0 - q[0] * uv2);
  out[2] = (v[2] + q[3] * uv2) + (q[0] * uv1 - q[1] * uv0);
}

void quat_mul(const float a[4], const float b[4], float o[4])
{
  const float w = a[3] * b[3] - a[0] * b[0] - a[1] * b[1] - a[2] * b[2];
  const float x = a[3] * b[0] + a[0] * b[3] + a[1] * b[2] - a[2] * b[1];
  const float y = a[3] * b[1] + a[1] * b[3] + a[2] * b[0] - a[0] * b[2];
  const float z = a[3] * b[2] + a[2] * b[3] + a[0] * b[1] - a[1] * b[0];
  o[0] = x; o[1] = y; o[2] = z; o[3] = w;
}

void quat_to_matrix(const float q[4], float R[9])
{
  const float tx = 2.f * q[0], ty = 2.f * q[1], tz = 2.f * q[2];
  const float twx = tx * q[3], twy = ty * q[3], twz = tz * q[3];
  const float txx = tx * q[0], txy = ty * q[0], txz = tz * q[0];
  const float tyy = ty * q[1], tyz = tz * q[1], tzz = tz * q[2];
  R[0] = 1.f - (tyy + tzz); R[1] = txy - twz;         R[2] = txz + twy;
  R[3] = txy + twz;         R[4] = 1.f - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy;         R[7] = tyz + twx;         R[8] = 1.f - (txx + tyy);
}

// warping.h:98-137 RelativePose(pose_a, pose_b, jac_a, jac_b): pose_ab = a^-1 * b
void relative_pose(const float a[7], const float b[7], float ab[7], float* jac_a, float* jac_b)
{
  const float qi[4] = {-a[0], -a[1], -a[2], a[3]};
  const float nta[3] = {-a[4], -a[5], -a[6]};
  float ti[3], tmp[3];
  quat_rotate(qi, nta, ti);
  quat_mul(qi, b, ab);
  quat_rotate(qi, b + 4, tmp);
  ab[4] = ti[0] + tmp[0]; ab[5] = ti[1] + tmp[1]; ab[6] = ti[2] + tmp[2];
  if (!jac_a && !jac_b) return;
  float Ra[9], RaT[9];
  quat_to_matrix(a, Ra);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) RaT[i * 3 + j] = Ra[j * 3 + i];
  if (jac_a) {
    const float d[3] = {a[4] - b[4], a[5] - b[5], a[6] - b[6]};
    float v[3];
    for (int i = 0; i < 3; ++i) v[i] = RaT[i * 3 + 0] * d[0] + RaT[i * 3 + 1] * d[1] + RaT[i * 3 + 2] * d[2];
    const float hat[9] = {0, -v[2], v[1], v[2], 0, -v[0], -v[1], v[0], 0};
    for (int i = 0; i < 36; ++i) jac_a[i] = 0.f;
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) {
        float s = 0.f;
        for (int k = 0; k < 3; ++k) s += hat[i * 3 + k] * RaT[k * 3 + j];
        jac_a[i * 6 + j] = -RaT[i * 3 + j];
        jac_a[i * 6 + 3 + j] = -s;
        jac_a[(3 + i) * 6 + 3 + j] = -RaT[i * 3 + j];
      }
  }
  if (jac_b) {
    for (int i = 0; i < 36; ++i) jac_b[i] = 0.f;
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) {
        jac_b[i * 6 + j] = RaT[i * 3 + j];
        jac_b[(3 + i) * 6 + 3 + j] = RaT[i * 3 + j];
      }
  }
}

// ---------------------------------------------------------------------------- validation helpers
bool img_ok(const DfkImage* im, uint32_t w, uint32_t h, uint32_t floats_per_px)
{
  return im && im->ptr && im->width == w && im->height == h && (im->pitch_bytes % 4 == 0) &&
         im->pitch_bytes >= (size_t)w * floats_per_px * 4;
}

// The validity window comes from the camera (PixelValid: u < cam.width - border, pinhole_camera_impl.h:102-108) while the
// bilinear taps index the images: a camera larger than the level it is used with (e.g. a level-0 camera with level-1
// buffers) would read outside them.  The reference has no such check (it would read out of bounds); here it is an
// argument error.
bool cam_ok(const DfkCamera* cam, uint32_t w, uint32_t h)
{
  return cam && cam->width <= (float)w && cam->height <= (float)h && cam->width >= 0.0f && cam->height >= 0.0f;
}

View view_of(const DfkImage* im) { return View{static_cast<const float*>(im->ptr), (uint32_t)(im->pitch_bytes / 4)}; }

bool aligned(const void* p, size_t a) { return reinterpret_cast<uintptr_t>(p) % a == 0; }

PixelCam make_pixel_cam(const float pose[7], const DfkCamera* cam, int border, float min_dpt)
{
  PixelCam pc;
  for (int i = 0; i < 4; ++i) pc.q[i] = pose[i];
  for (int i = 0; i < 3; ++i) pc.t[i] = pose[4 + i];
  pc.fx = cam->fx; pc.fy = cam->fy; pc.u0 = cam->u0; pc.v0 = cam->v0;
  pc.border = (float)border;
  pc.ulim = cam->width - (float)border;   // PixelValid: x < width_ - border (pinhole_camera_impl.h:107)
  pc.vlim = cam->height - (float)border;
  pc.min_dpt = min_dpt;
  return pc;
}

uint32_t gcd_u32(uint32_t a, uint32_t b)
{
  while (b) {
    const uint32_t t = a % b;
    a = b;
    b = t;
  }
  return a;
}

// stride for the in-item tile permutation: ~golden-ratio of the tile count, coprime with it
uint32_t perm_multiplier(uint32_t n)
{
  static const bool no_perm = []() { const char* e = getenv("DFK_NO_PERM"); return e && e[0] == '1'; }();
  if (no_perm) return 1;
  if (n <= 2 || n > 65535u) return 1;  // keeps k * perm_mul below 2^32 on the device
  uint32_t m = (uint32_t)((double)n * 0.6180339887498949);
  if (m < 1) m = 1;
  while (gcd_u32(m, n) != 1) ++m;
  return m % n == 0 ? 1 : m;
}

template <typename T>
cudaError_t ensure(T** ptr, size_t* cap, size_t need)
{
  if (*cap >= need) return cudaSuccess;
  const size_t old_cap = *cap;
  // cudaFree synchronises the device, so work still running on the stream has finished with the old buffer
  if (*ptr) cudaFree(*ptr);
  *ptr = nullptr;
  *cap = 0;
  size_t n = std::max(need, old_cap * 2);
  cudaError_t e = cudaMalloc(reinterpret_cast<void**>(ptr), n * sizeof(T));
  if (e == cudaSuccess) *cap = n;
  return e;
}

DfkStatus build_items(DfkHandle h, const DfkSfmWorkItem* items, int n, int code_size, int tile_px, int max_ctas,
                      bool want_ray_tabs, const float* codes_dev, SfmLaunchPlan* plan)
{
  h->ray_miss = false;
  const DfkDenseSfmParams& sp = h->params.sfmparams;
  h->items_host.resize(n);
  uint32_t tile_cursor = 0;
  for (int i = 0; i < n; ++i) {
    const DfkSfmWorkItem& w = items[i];
    SfmItemDev& d = h->items_host[i];
    const uint32_t W = w.img0.width, H = w.img0.height;
    if (W == 0 || H == 0) return fail(h, DFK_ERR_INVALID_ARG, "[SfmAligner::RunStep] empty image");
    if (!img_ok(&w.img0, W, H, 1) || !img_ok(&w.img1, W, H, 1) || !img_ok(&w.dpt0, W, H, 1) ||
        !img_ok(&w.valid0, W, H, 1) || !img_ok(&w.prx0_jac, W, H, code_size) || !img_ok(&w.grad1, W, H, 2))
      return fail(h, DFK_ERR_INVALID_ARG,
                  "[SfmAligner::RunStep] inconsistent image views (size, pitch or null pointer) in work item " +
                      std::to_string(i));
    if (!cam_ok(&w.cam, W, H))
      return fail(h, DFK_ERR_INVALID_ARG,
                  "[SfmAligner::RunStep] camera viewport larger than the image views in work item " + std::to_string(i));
    float p10[7];
    relative_pose(w.pose1, w.pose0, p10, d.P1, d.P0);  // RelativePose(pose1, pose0, J_pose1, J_pose0)
    for (int k = 0; k < 4; ++k) d.q[k] = p10[k];
    for (int k = 0; k < 3; ++k) d.t[k] = p10[4 + k];
    quat_to_matrix(p10, d.R);
    d.fx = w.cam.fx; d.fy = w.cam.fy; d.u0 = w.cam.u0; d.v0 = w.cam.v0;
    d.border = (float)sp.valid_border;
    d.ulim = w.cam.width - (float)sp.valid_border;
    d.vlim = w.cam.height - (float)sp.valid_border;
    d.min_dpt = sp.min_dpt; d.avg_dpt = sp.avg_dpt; d.huber_delta = sp.huber_delta;
    d.img0 = (const float*)w.img0.ptr; d.img1 = (const float*)w.img1.ptr; d.dpt0 = (const float*)w.dpt0.ptr;
    d.valid0 = (float*)w.valid0.ptr; d.jac = (const float*)w.prx0_jac.ptr; d.grad1 = (const float*)w.grad1.ptr;
    d.img0_pitch = (uint32_t)(w.img0.pitch_bytes / 4); d.img1_pitch = (uint32_t)(w.img1.pitch_bytes / 4);
    d.dpt0_pitch = (uint32_t)(w.dpt0.pitch_bytes / 4); d.valid0_pitch = (uint32_t)(w.valid0.pitch_bytes / 4);
    d.jac_pitch = (uint32_t)(w.prx0_jac.pitch_bytes / 4); d.grad1_pitch = (uint32_t)(w.grad1.pitch_bytes / 4);
    d.dpt_out = nullptr; d.dpt_out_pitch = 0; d.code = nullptr;
    const bool fused = (w.code != nullptr);
    if (fused) {
      // UpdateDepth + RunStep in one pass: the tile loader stages prx_orig where it would stage dpt0, the front-end
      // decodes the depth (bit for bit what dfk_update_depth computes) and writes it to dpt0
      if (!img_ok(&w.prx_orig, W, H, 1))
        return fail(h, DFK_ERR_INVALID_ARG,
                    "[SfmAligner::RunStep] fused depth decode: inconsistent prx_orig view in work item " + std::to_string(i));
      d.dpt_out = (float*)w.dpt0.ptr;
      d.dpt_out_pitch = d.dpt0_pitch;
      d.dpt0 = (const float*)w.prx_orig.ptr;
      d.dpt0_pitch = (uint32_t)(w.prx_orig.pitch_bytes / 4);
      d.code = codes_dev + (size_t)i * code_size;
      memcpy(h->codes_host.data() + (size_t)i * code_size, w.code, sizeof(float) * code_size);
    }
    d.width = W; d.height = H; d.num_pixels = W * H;
    d.num_tiles = (d.num_pixels + tile_px - 1) / tile_px;
    d.ray_tab = nullptr;
    if (want_ray_tabs) {
      for (const auto& r : h->ray_cache)
        if (r.fx == d.fx && r.fy == d.fy && r.u0 == d.u0 && r.v0 == d.v0 && r.w == W && r.h == H) {
          d.ray_tab = r.dev;
          break;
        }
      if (!d.ray_tab) {
        if (h->ray_cache.size() >= 256) {  // a caller cycling through cameras: start over (cudaFree synchronises)
          for (auto& r : h->ray_cache) cudaFree(r.dev);
          h->ray_cache.clear();
        }
        DfkContext::RayTab r{d.fx, d.fy, d.u0, d.v0, W, H, nullptr};
        if (cudaMalloc((void**)&r.dev, sizeof(float) * ((size_t)W + H)) != cudaSuccess)
          return fail(h, DFK_ERR_CUDA, "[SfmAligner::RunStep] scratch allocation failed");
        h->ray_cache.push_back(r);
        d.ray_tab = r.dev;
        h->ray_miss = true;
      }
    }
    d.tile_begin = tile_cursor;
    tile_cursor += d.num_tiles;
    d.perm_mul = perm_multiplier(d.num_tiles);
    d.mag_tiles = (uint32_t)((1ull << 32) / d.num_tiles);
    d.mag_width = (uint32_t)((1ull << 32) / W);
    d.flags = 0;
    const bool bulk = (W % 4 == 0) && aligned(d.img0, 16) && aligned(d.dpt0, 16) && aligned(d.jac, 16) &&
                      (d.img0_pitch % 4 == 0) && (d.dpt0_pitch % 4 == 0) && (d.jac_pitch % 4 == 0) &&
                      (code_size % 4 == 0);
    if (bulk) d.flags |= ITEM_FLAG_BULK;
    if (aligned(d.grad1, 8) && d.grad1_pitch % 2 == 0) d.flags |= ITEM_FLAG_GRAD_ALIGNED;
    if (fused) d.flags |= ITEM_FLAG_FUSED_DEPTH;
  }
  const int T = (int)tile_cursor;
  int G = std::min(max_ctas, T);
  if (G < 1) G = 1;
  plan->num_items = n;
  plan->num_tiles = T;
  plan->num_ctas = G;
  // which CTAs touch which item (CTA c owns global tiles [c*T/G, (c+1)*T/G))
  uint32_t partial_cursor = 0;
  int c = 0;
  int max_per_item = 0;
  for (int i = 0; i < n; ++i) {
    SfmItemDev& d = h->items_host[i];
    const long long tb = d.tile_begin, te = tb + d.num_tiles;
    while ((long long)(c + 1) * T / G <= tb) ++c;  // first CTA whose range ends after tb
    int first = c, last = c;
    while ((long long)(last + 1) * T / G < te) ++last;
    d.first_cta = (uint32_t)first;
    d.num_ctas = (uint32_t)(last - first + 1);
    d.partial_begin = partial_cursor;
    partial_cursor += d.num_ctas;
    max_per_item = std::max(max_per_item, (int)d.num_ctas);
  }
  plan->num_partials = (int)partial_cursor;
  plan->max_ctas_per_item = max_per_item;
  return DFK_OK;
}

constexpr size_t kMaxEventPairs = 8192;

// drains recorded event pairs into the accumulators (synchronizes the stream)
DfkStatus drain_events(DfkHandle h)
{
  if (h->ev_used == 0) return DFK_OK;
  DFK_CUDA(h, cudaStreamSynchronize(h->stream), "profiling: stream synchronize failed");
  for (size_t k = 0; k < h->ev_used; ++k) {
    float ms = 0.f;
    DFK_CUDA(h, cudaEventElapsedTime(&ms, h->ev_pool[2 * k], h->ev_pool[2 * k + 1]), "profiling: event read failed");
    h->ev_ms_accum += ms;
  }
  h->ev_count_accum += h->ev_used;
  h->ev_used = 0;
  return DFK_OK;
}

DfkStatus profile_events(DfkHandle h, cudaEvent_t* e0, cudaEvent_t* e1)
{
  if (h->ev_used == kMaxEventPairs) {
    DfkStatus st = drain_events(h);
    if (st != DFK_OK) return st;
  }
  while (h->ev_pool.size() < 2 * (h->ev_used + 1)) {
    cudaEvent_t e;
    DFK_CUDA(h, cudaEventCreate(&e), "profiling: event creation failed");
    h->ev_pool.push_back(e);
  }
  *e0 = h->ev_pool[2 * h->ev_used];
  *e1 = h->ev_pool[2 * h->ev_used + 1];
  h->ev_used += 1;
  return DFK_OK;
}

DfkStatus run_batch(DfkHandle h, const DfkSfmWorkItem* items, int n, int code_size, float* records_dev)
{
  if (!h) return DFK_ERR_INVALID_ARG;
  if (!items || n <= 0 || !records_dev) return fail(h, DFK_ERR_INVALID_ARG, "[SfmAligner::RunStep] null/empty batch");
  const bool wide = sfm_wide_supported(code_size);
  if (!sfm_fp32_supported(code_size) && !wide)
    return fail(h, DFK_ERR_UNSUPPORTED,
                "[SfmAligner::RunStep] no kernel instantiated for code size " + std::to_string(code_size));
  bool tc = (h->gram_mode == DFK_GRAM_TF32X3) || (h->gram_mode == DFK_GRAM_AUTO && sfm_tc_supported(code_size));
  if (tc) {
    // the tensor-core kernel gathers grad1 with 8-byte loads; odd layouts go to the fp32 kernel (AUTO) or fail (forced)
    bool grads_ok = true;
    for (int i = 0; i < n && grads_ok; ++i)
      grads_ok = items[i].grad1.ptr && aligned(items[i].grad1.ptr, 8) && (items[i].grad1.pitch_bytes % 8 == 0);
    if (!grads_ok) {
      if (h->gram_mode == DFK_GRAM_TF32X3)
        return fail(h, DFK_ERR_UNSUPPORTED, "[SfmAligner::RunStep] tensor-core path needs 8-byte aligned grad1 rows");
      tc = false;
    }
  }
  if (tc && !sfm_tc_supported(code_size))
    return fail(h, DFK_ERR_UNSUPPORTED,
                "[SfmAligner::RunStep] tensor-core Gram path is not instantiated for code size " +
                    std::to_string(code_size));
  DeviceGuard guard(h->device);
  SfmLaunchPlan plan;
  const int tile_px = tc ? kTcTilePixels : (wide ? sfm_wide_tile_pixels(code_size) : kTilePixels);
  bool any_fused = false;
  for (int i = 0; i < n; ++i) any_fused = any_fused || items[i].code != nullptr;
  if (any_fused) {
    DFK_CUDA(h, ensure(&h->codes_dev, &h->codes_cap, (size_t)n * code_size), "[SfmAligner::RunStep] scratch allocation failed");
    h->codes_host.assign((size_t)n * code_size, 0.0f);
  }
  const int ctas_per_sm = tc ? 2 : (wide ? 1 : sfm_fp32_ctas_per_sm(code_size));
  const int sms = (h->sm_limit > 0 && h->sm_limit < h->num_sms) ? h->sm_limit : h->num_sms;
  DfkStatus st = build_items(h, items, n, code_size, tile_px, ctas_per_sm * sms,
                             tc, h->codes_dev, &plan);
  if (st != DFK_OK) return st;
  if (any_fused)
    DFK_CUDA(h, cudaMemcpyAsync(h->codes_dev, h->codes_host.data(), sizeof(float) * (size_t)n * code_size,
                                cudaMemcpyHostToDevice, h->stream),
             "[SfmAligner::RunStep] code upload failed");
  const size_t pfloats = tc ? (size_t)kTcPartialFloats : sfm_partial_floats(code_size);
  DFK_CUDA(h, ensure(&h->items_dev, &h->items_cap, (size_t)n), "[SfmAligner::RunStep] scratch allocation failed");
  DFK_CUDA(h, ensure(&h->partials_dev, &h->partials_cap, (size_t)plan.num_partials * pfloats),
           "[SfmAligner::RunStep] scratch allocation failed");
  DFK_CUDA(h, cudaMemcpyAsync(h->items_dev, h->items_host.data(), sizeof(SfmItemDev) * n, cudaMemcpyHostToDevice,
                              h->stream),
           "[SfmAligner::RunStep] work list upload failed");
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  if (h->profiling) {
    DfkStatus ps = profile_events(h, &ev0, &ev1);
    if (ps != DFK_OK) return ps;
  }
  if (tc) {
    DFK_CUDA(h, launch_sfm_tc(h->items_dev, plan, h->ray_miss, h->partials_dev, h->stream, ev0, ev1),
             "[SfmAligner::RunStep] kernel launch failed");
    if (h->ray_miss) h->launches += 1;  // ray-table kernel
  } else if (wide) {
    DFK_CUDA(h, launch_sfm_wide(code_size, h->items_dev, plan, h->partials_dev, h->stream, ev0, ev1),
             "[SfmAligner::RunStep] kernel launch failed");
  } else {
    DFK_CUDA(h, launch_sfm_fp32(code_size, h->items_dev, plan, h->partials_dev, records_dev, h->stream, ev0, ev1),
             "[SfmAligner::RunStep] kernel launch failed");
  }
  DFK_CUDA(h, launch_sfm_finalize(code_size, tc, h->items_dev, n, h->partials_dev, records_dev, h->stream),
           "[SfmAligner::RunStep] kernel launch failed");
  h->launches += 2;  // step kernel + finalize kernel
  return DFK_OK;
}

}  // namespace

// ================================================================================== C ABI
extern "C" {

int dfk_version(void) { return DFK_VERSION; }

const char* dfk_status_string(DfkStatus s)
{
  switch (s) {
    case DFK_OK: return "ok";
    case DFK_ERR_INVALID_ARG: return "invalid argument";
    case DFK_ERR_CUDA: return "CUDA error";
    case DFK_ERR_UNSUPPORTED: return "unsupported";
    case DFK_ERR_NOMEM: return "out of memory";
  }
  return "unknown";
}

int dfk_sfm_supports_code_size(int code_size)
{
  return (sfm_fp32_supported(code_size) || sfm_wide_supported(code_size)) ? 1 : 0;
}

DfkStatus dfk_create(int device, DfkHandle* out)
{
  try {
    if (!out) return DFK_ERR_INVALID_ARG;
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) return DFK_ERR_CUDA;
    if (device < 0) {
      if (cudaGetDevice(&device) != cudaSuccess) return DFK_ERR_CUDA;
    }
    if (device >= count) return DFK_ERR_INVALID_ARG;
    DfkContext* h = new (std::nothrow) DfkContext();
    if (!h) return DFK_ERR_NOMEM;
    h->device = device;
    h->params.sfmparams = DfkDenseSfmParams{0.1f, 1000.f, 2.0f, 0.0f, 2};
    h->params.step_threads = 32; h->params.step_blocks = 11; h->params.eval_threads = 224; h->params.eval_blocks = 66;
    DeviceGuard guard(device);
    bool ok = true;
    ok = ok && cudaDeviceGetAttribute(&h->num_sms, cudaDevAttrMultiProcessorCount, device) == cudaSuccess;
    ok = ok && cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking) == cudaSuccess;
    h->stream = h->own_stream;
    ok = ok && cudaMalloc((void**)&h->simple_scratch, sizeof(float) * kSimpleScratchFloats) == cudaSuccess;
    ok = ok && cudaMalloc((void**)&h->counter, sizeof(unsigned int)) == cudaSuccess;
    ok = ok && cudaMalloc((void**)&h->out_dev, sizeof(float) * 32) == cudaSuccess;
    ok = ok && cudaMalloc((void**)&h->code_dev, sizeof(float) * 256) == cudaSuccess;
    ok = ok && cudaMallocHost((void**)&h->out_host, sizeof(float) * 32) == cudaSuccess;
    ok = ok && cudaMemset(h->counter, 0, sizeof(unsigned int)) == cudaSuccess;
    if (!ok) {
      dfk_destroy(h);
      return DFK_ERR_CUDA;
    }
    *out = h;
    return DFK_OK;
  } catch (...) {
    return DFK_ERR_NOMEM;
  }
}

DfkStatus dfk_destroy(DfkHandle h)
{
  try {
    if (!h) return DFK_OK;
    DeviceGuard guard(h->device);
    if (h->own_stream) cudaStreamSynchronize(h->own_stream);
    cudaFree(h->simple_scratch); cudaFree(h->counter); cudaFree(h->out_dev); cudaFree(h->code_dev);
    cudaFree(h->track_dev);
    cudaFree(h->codes_dev);
    if (h->track_host) cudaFreeHost(h->track_host);
    cudaFree(h->items_dev); cudaFree(h->partials_dev); cudaFree(h->records_dev);
    for (auto& r : h->ray_cache) cudaFree(r.dev);
    if (h->out_host) cudaFreeHost(h->out_host);
    if (h->records_host) cudaFreeHost(h->records_host);
    for (cudaEvent_t e : h->ev_pool) cudaEventDestroy(e);
    if (h->own_stream) cudaStreamDestroy(h->own_stream);
    delete h;
    return DFK_OK;
  } catch (...) {  // std::bad_alloc / std::length_error from host containers must not cross the C ABI
    return oom(h);
  }
}

DfkStatus dfk_set_stream(DfkHandle h, void* cuda_stream)
{
  try {
    if (!h) return DFK_ERR_INVALID_ARG;
    h->stream = static_cast<cudaStream_t>(cuda_stream);
    return DFK_OK;
  } catch (...) {  // std::bad_alloc / std::length_error from host containers must not cross the C ABI
    return oom(h);
  }
}

DfkStatus dfk_use_own_stream(DfkHandle h)
{
  try {
    if (!h) return DFK_ERR_INVALID_ARG;
    h->stream = h->own_stream;
    return DFK_OK;
  } catch (...) {  // std::bad_alloc / std::length_error from host containers must not cross the C ABI
    return oom(h);
  }
}

DfkStatus dfk_set_sm_limit(DfkHandle h, int num_sms)
{
  if (!h) return DFK_ERR_INVALID_ARG;
  try {
    if (num_sms < 0) return fail(h, DFK_ERR_INVALID_ARG, "[dfk_set_sm_limit] num_sms < 0");
    h->sm_limit = num_sms;
    return DFK_OK;
  } catch (const std::bad_alloc&) {
    return oom(h);
  } catch (...) {
    return DFK_ERR_CUDA;
  }
}

void* dfk_get_stream(DfkHandle h) { return h ? h->stream : nullptr; }

DfkStatus dfk_synchronize(DfkHandle h)
{
  try {
    if (!h) return DFK_ERR_INVALID_ARG;
    DeviceGuard guard(h->device);
    DFK_CUDA(h, cudaStreamSynchronize(h->stream), "stream synchronize failed");
    return DFK_OK;
  } catch (...) {  // std::bad_alloc / std::length_error from host containers must not cross the C ABI
    return oom(h);
  }
}

const char* dfk_last_error(DfkHandle h) { return h ? h->err.c_str() : "null handle"; }

DfkStatus dfk_set_profiling(DfkHandle h, int enabled)
{
  try {
    if (!h) return DFK_ERR_INVALID_ARG;
    h->profiling = enabled != 0;
    return DFK_OK;
  } catch (...) {  // std::bad_alloc / std::length_error from host containers must not cross the C ABI
    return oom(h);
  }
}

DfkStatus dfk_get_profile(DfkHandle h, double* main_kernel_ms, uint64_t* main_kernel_launches,
                          uint64_t* total_kernel_launches)
{
  try {
    if (!h) return DFK_ERR_INVALID_ARG;
    DeviceGuard guard(h->device);
    DfkStatus st = drain_events(h);
    if (st != DFK_OK) return st;
    if (main_kernel_ms) *main_kernel_ms = h->ev_ms_accum;
    if (main_kernel_launches) *main_kernel_launches = h->ev_count_accum;
    if (total_kernel_launches) *total_kernel_launches = h->launches;
    h->ev_ms_accum = 0.0;
    h->ev_count_accum = 0;
    h->launches = 0;
    return DFK_OK;
  } catch (...) {  // std::bad_alloc / std::length_error from host containers must not cross the C ABI
    return oom(h);
  }
}

DfkStatus dfk_sfm_set_params(DfkHandle h, const DfkSfmAlignerParams* p)
{
  try {
    if (!h || !p) return DFK_ERR_INVALID_ARG;
    // CHECK_EQ(threads % 32, 0), CHECK_LE(blocks, max_blocks) (cu_sfmaligner.cpp:187-203)
    if (p->step_threads % 32 != 0 || p->eval_threads % 32 != 0)
      return fail(h, DFK_ERR_INVALID_ARG, "threads must be a multiple of 32!");
    if (p->step_blocks > 1024 || p->eval_blocks > 1024) return fail(h, DFK_ERR_INVALID_ARG, "blocks must be less than 1024");
    if (p->sfmparams.valid_border < 1)
      return fail(h, DFK_ERR_INVALID_ARG, "valid_border must be >= 1 (bilinear sampling reads ix+1, iy+1)");
    h->params = *p;
    return DFK_OK;
  } catch (...) {  // std::bad_alloc / std::length_error from host containers must not cross the C ABI
    return oom(h);
  }
}

DfkStatus dfk_sfm_get_params(DfkHandle h, DfkSfmAlignerParams* p)
{
  try {
    if (!h || !p) return DFK_ERR_INVALID_ARG;
    *p = h->params;
    return DFK_OK;
  } catch (...) {  // std::bad_alloc / std::length_error from host containers must not cross the C ABI
    return oom(h);
  }
}

DfkStatus dfk_sfm_set_gram_mode(DfkHandle h, DfkGramMode m)
{
  try {
    if (!h) return DFK_ERR_INVALID_ARG;
    if (m != DFK_GRAM_AUTO && m != DFK_GRAM_FP32 && m != DFK_GRAM_TF32X3)
      return fail(h, DFK_ERR_INVALID_ARG, "unknown gram mode");
    h->gram_mode = m;
    return DFK_OK;
  } catch (...) {  // std::bad_alloc / std::length_error from host containers must not cross the C ABI
    return oom(h);
  }
}

DfkStatus dfk_se3_set_huber_delta(DfkHandle h, float v)
{
  try {
    if (!h) return DFK_ERR_INVALID_ARG;
    h->se3_huber_delta = v;
    return DFK_OK;
  } catch (...) {  // std::bad_alloc / std::length_error from host containers must not cross the C ABI
    return oom(h);
  }
}

DfkStatus dfk_sfm_run_step_batch(DfkHandle h, const DfkSfmWorkItem* items, int n, int code_size, float* records_dev)
{
  try {
    return run_batch(h, items, n, code_size, records_dev);
  } catch (...) {  // std::bad_alloc / std::length_error from host containers must not cross the C ABI
    return oom(h);
  }
}

DfkStatus dfk_sfm_run_step_batch_host(DfkHandle h, const DfkSfmWorkItem* items, int n, int code_size,
                                      float* records_host)
{
  try {
    if (!h) return DFK_ERR_INVALID_ARG;
    if (!records_host || n <= 0) return fail(h, DFK_ERR_INVALID_ARG, "[SfmAligner::RunStep] null/empty batch");
    if (!dfk_sfm_supports_code_size(code_size))
      return fail(h, DFK_ERR_UNSUPPORTED,
                  "[SfmAligner::RunStep] no kernel instantiated for code size " + std::to_string(code_size));
    DeviceGuard guard(h->device);
    const size_t rec = (size_t)DFK_SFM_RECORD_FLOATS(code_size);
    DFK_CUDA(h, ensure(&h->records_dev, &h->records_cap, rec * n), "[SfmAligner::RunStep] scratch allocation failed");
    if (h->records_host_cap < rec * n) {
      if (h->records_host) cudaFreeHost(h->records_host);
      h->records_host = nullptr;
      h->records_host_cap = 0;
      DFK_CUDA(h, cudaMallocHost((void**)&h->records_host, rec * n * sizeof(float)),
               "[SfmAligner::RunStep] pinned allocation failed");
      h->records_host_cap = rec * n;
    }
    DfkStatus st = run_batch(h, items, n, code_size, h->records_dev);
    if (st != DFK_OK) return st;
    DFK_CUDA(h, cudaMemcpyAsync(h->records_host, h->records_dev, rec * n * sizeof(float), cudaMemcpyDeviceToHost,
                                h->stream),
             "[SfmAligner::RunStep] result download failed");
    DFK_CUDA(h, cudaStreamSynchronize(h->stream), "[SfmAligner::RunStep] kernel launch failed");
    memcpy(records_host, h->records_host, rec * n * sizeof(float));
    return DFK_OK;
  } catch (...) {  // std::bad_alloc / std::length_error from host containers must not cross the C ABI
    return oom(h);
  }
}

DfkStatus dfk_sfm_run_step(DfkHandle h, const float pose0[7], const float pose1[7], const float* /*code0*/,
                           int code_size, const DfkCamera* cam, const DfkImage* img0, const DfkImage* img1,
                           const DfkImage* dpt0, const DfkImage* /*std0*/, const DfkImage* valid0,
                           const DfkImage* prx0_jac, const DfkImage* grad1, float* JtJ, float* Jtr, float* residual,
                           uint64_t* inliers)
{
  try {
    if (!h) return DFK_ERR_INVALID_ARG;
    if (!pose0 || !pose1 || !cam || !img0 || !img1 || !dpt0 || !valid0 || !prx0_jac || !grad1 || !JtJ || !Jtr ||
        !residual || !inliers)
      return fail(h, DFK_ERR_INVALID_ARG, "[SfmAligner::RunStep] null argument");
    DfkSfmWorkItem w{};  // no fused depth decode: code = NULL
    memcpy(w.pose0, pose0, sizeof(w.pose0));
    memcpy(w.pose1, pose1, sizeof(w.pose1));
    w.cam = *cam;
    w.img0 = *img0; w.img1 = *img1; w.dpt0 = *dpt0; w.valid0 = *valid0; w.prx0_jac = *prx0_jac; w.grad1 = *grad1;
    const int NP = 12 + code_size;
    const int NH = NP * (NP + 1) / 2;
    std::vector<float> rec((size_t)DFK_SFM_RECORD_FLOATS(code_size));
    DfkStatus st = dfk_sfm_run_step_batch_host(h, &w, 1, code_size, rec.data());
    if (st != DFK_OK) return st;
    memcpy(JtJ, rec.data(), sizeof(float) * NH);
    memcpy(Jtr, rec.data() + NH, sizeof(float) * NP);
    *residual = rec[NH + NP];
    uint32_t bits;
    memcpy(&bits, &rec[NH + NP + 1], 4);
    *inliers = bits;
    return DFK_OK;
  } catch (...) {  // std::bad_alloc / std::length_error from host containers must not cross the C ABI
    return oom(h);
  }
}

static DfkStatus fetch_out(DfkHandle h, int nfloats, const char* what)
{
  try {
    h->launches += 1;
    DFK_CUDA(h, cudaMemcpyAsync(h->out_host, h->out_dev, sizeof(float) * nfloats, cudaMemcpyDeviceToHost, h->stream),
             what);
    DFK_CUDA(h, cudaStreamSynchronize(h->stream), what);
    return DFK_OK;
  } catch (...) {  // std::bad_alloc / std::length_error from host containers must not cross the C ABI
    return oom(h);
  }
}

DfkStatus dfk_sfm_evaluate_error(DfkHandle h, const float pose0[7], const float pose1[7], const DfkCamera* cam,
                                 const DfkImage* img0, const DfkImage* img1, const DfkImage* dpt0,
                                 const DfkImage* /*std0*/, const DfkImage* /*grad1*/, float* residual,
                                 uint64_t* inliers)
{
  try {
    if (!h) return DFK_ERR_INVALID_ARG;
    if (!pose0 || !pose1 || !cam || !img0 || !img1 || !dpt0 || !residual || !inliers)
      return fail(h, DFK_ERR_INVALID_ARG, "[SfmAligner::EvaluateError] null argument");
    const uint32_t W = img0->width, H = img0->height;
    if (W == 0 || H == 0 || !img_ok(img0, W, H, 1) || !img_ok(img1, W, H, 1) || !img_ok(dpt0, W, H, 1))
      return fail(h, DFK_ERR_INVALID_ARG, "[SfmAligner::EvaluateError] inconsistent image views");
    if (!cam_ok(cam, W, H)) return fail(h, DFK_ERR_INVALID_ARG, "[SfmAligner::EvaluateError] camera viewport larger than the image views");
    DeviceGuard guard(h->device);
    float p10[7];
    relative_pose(pose1, pose0, p10, nullptr, nullptr);  // cu_sfmaligner.cpp:131
    // DenseSfm_EvaluateError uses FindCorrespondence defaults: border 1, min_dpt 0 (dense_sfm.h:91)
    const PixelCam pc = make_pixel_cam(p10, cam, 1, 0.0f);
    DFK_CUDA(h, launch_eval_error(pc, h->params.sfmparams.huber_delta, (int)W, (int)H, view_of(img0), view_of(img1),
                                  view_of(dpt0), h->simple_scratch, h->counter, h->out_dev, h->stream),
             "[SfmAligner::EvaluateError] kernel launch failed");
    DfkStatus st = fetch_out(h, 2, "[SfmAligner::EvaluateError] kernel launch failed");
    if (st != DFK_OK) return st;
    *residual = h->out_host[0];
    uint32_t bits;
    memcpy(&bits, &h->out_host[1], 4);
    *inliers = bits;
    return DFK_OK;
  } catch (...) {  // std::bad_alloc / std::length_error from host containers must not cross the C ABI
    return oom(h);
  }
}

DfkStatus dfk_se3_run_step(DfkHandle h, const float se3[7], const DfkCamera* cam, const DfkImage* img0,
                           const DfkImage* img1, const DfkImage* dpt0, const DfkImage* grad1, float* JtJ, float* Jtr,
                           float* residual, uint64_t* inliers)
{
  try {
    if (!h) return DFK_ERR_INVALID_ARG;
    if (!se3 || !cam || !img0 || !img1 || !dpt0 || !grad1 || !JtJ || !Jtr || !residual || !inliers)
      return fail(h, DFK_ERR_INVALID_ARG, "[SE3Aligner::RunStep] null argument");
    const uint32_t W = img0->width, H = img0->height;
    if (W == 0 || H == 0 || !img_ok(img0, W, H, 1) || !img_ok(img1, W, H, 1) || !img_ok(dpt0, W, H, 1) ||
        !img_ok(grad1, W, H, 2))
      return fail(h, DFK_ERR_INVALID_ARG, "[SE3Aligner::RunStep] inconsistent image views");
    if (!cam_ok(cam, W, H)) return fail(h, DFK_ERR_INVALID_ARG, "[SE3Aligner::RunStep] camera viewport larger than the image views");
    DeviceGuard guard(h->device);
    const PixelCam pc = make_pixel_cam(se3, cam, 1, 0.0f);  // lucas_kanade_se3.h:52 defaults
    const View g = view_of(grad1);
    const bool galigned = aligned(g.ptr, 8) && g.pitch % 2 == 0;
    DFK_CUDA(h, launch_se3_step(pc, h->se3_huber_delta, (int)W, (int)H, view_of(img0), view_of(img1), view_of(dpt0), g,
                                galigned, h->simple_scratch, h->counter, h->out_dev, h->stream),
             "[SE3Aligner::RunStep] Kernel launch failed");
    DfkStatus st = fetch_out(h, 29, "[SE3Aligner::RunStep] Kernel launch failed");
    if (st != DFK_OK) return st;
    memcpy(JtJ, h->out_host, sizeof(float) * 21);
    memcpy(Jtr, h->out_host + 21, sizeof(float) * 6);
    *residual = h->out_host[27];
    uint32_t bits;
    memcpy(&bits, &h->out_host[28], 4);
    *inliers = bits;
    return DFK_OK;
  } catch (...) {  // std::bad_alloc / std::length_error from host containers must not cross the C ABI
    return oom(h);
  }
}

DfkStatus dfk_se3_track(DfkHandle h, float pose_ck[7], const DfkTrackLevel* levels, int num_levels,
                        float* inlier_fraction, float* error, float* last_system, float* history, int history_capacity)
{
  try {
    if (!h) return DFK_ERR_INVALID_ARG;
    if (!pose_ck || !levels || num_levels <= 0)
      return fail(h, DFK_ERR_INVALID_ARG, "[CameraTracker::TrackFrame] null argument / no pyramid levels");
    int total_iters = 0;
    for (int l = 0; l < num_levels; ++l) {
      const DfkTrackLevel& L = levels[l];
      const uint32_t W = L.img0.width, H = L.img0.height;
      if (L.iterations < 0 || W == 0 || H == 0 || !img_ok(&L.img0, W, H, 1) || !img_ok(&L.img1, W, H, 1) ||
          !img_ok(&L.dpt0, W, H, 1) || !img_ok(&L.grad1, W, H, 2) || !cam_ok(&L.cam, W, H))
        return fail(h, DFK_ERR_INVALID_ARG,
                    "[CameraTracker::TrackFrame] inconsistent image views / camera larger than them / negative iteration count at level " +
                        std::to_string(l));
      total_iters += L.iterations;
    }
    if (history && history_capacity < total_iters)
      return fail(h, DFK_ERR_INVALID_ARG, "[CameraTracker::TrackFrame] history buffer too small");
    DeviceGuard guard(h->device);
    const size_t nfloat = 8 + 36 * (size_t)std::max(total_iters, 1);
    DFK_CUDA(h, ensure(&h->track_dev, &h->track_cap, nfloat), "[CameraTracker::TrackFrame] scratch allocation failed");
    if (h->track_host_cap < nfloat + 32) {
      if (h->track_host) cudaFreeHost(h->track_host);
      h->track_host = nullptr;
      h->track_host_cap = 0;
      DFK_CUDA(h, cudaMallocHost((void**)&h->track_host, sizeof(float) * (nfloat + 32)),
               "[CameraTracker::TrackFrame] pinned allocation failed");
      h->track_host_cap = nfloat + 32;
    }
    // the pose goes to the device once; every iteration reads it there and its last block writes the update
    memcpy(h->track_host, pose_ck, sizeof(float) * 7);
    DFK_CUDA(h, cudaMemcpyAsync(h->track_dev, h->track_host, sizeof(float) * 7, cudaMemcpyHostToDevice, h->stream),
             "[CameraTracker::TrackFrame] pose upload failed");
    DFK_CUDA(h, cudaMemsetAsync(h->out_dev, 0, sizeof(float) * 32, h->stream), "[CameraTracker::TrackFrame] memset failed");
    int it = 0;
    uint32_t last_area = 0;
    for (int l = num_levels - 1; l >= 0; --l) {  // coarse to fine (camera_tracker.cpp:48)
      const DfkTrackLevel& L = levels[l];
      const PixelCam pc = make_pixel_cam(pose_ck, &L.cam, 1, 0.0f);  // q/t are overridden by the device pose
      const View g = view_of(&L.grad1);
      const bool galigned = aligned(g.ptr, 8) && g.pitch % 2 == 0;
      for (int k = 0; k < L.iterations; ++k, ++it) {
        DFK_CUDA(h, launch_se3_step(pc, h->se3_huber_delta, (int)L.img0.width, (int)L.img0.height, view_of(&L.img0),
                                    view_of(&L.img1), view_of(&L.dpt0), g, galigned, h->simple_scratch, h->counter,
                                    h->out_dev, h->stream, h->track_dev, h->track_dev + 8 + 36 * (size_t)it),
                 "[CameraTracker::TrackFrame] kernel launch failed");
        h->launches += 1;
        last_area = L.img0.width * L.img0.height;
      }
    }
    // one read-back: final pose, the last evaluated system, the per-iteration history
    float* host_sys = h->track_host + nfloat;
    DFK_CUDA(h, cudaMemcpyAsync(h->track_host, h->track_dev, sizeof(float) * (8 + 36 * (size_t)total_iters),
                                cudaMemcpyDeviceToHost, h->stream),
             "[CameraTracker::TrackFrame] read-back failed");
    DFK_CUDA(h, cudaMemcpyAsync(host_sys, h->out_dev, sizeof(float) * 29, cudaMemcpyDeviceToHost, h->stream),
             "[CameraTracker::TrackFrame] read-back failed");
    DFK_CUDA(h, cudaStreamSynchronize(h->stream), "[CameraTracker::TrackFrame] stream synchronize failed");
    memcpy(pose_ck, h->track_host, sizeof(float) * 7);
    uint32_t inl = 0;
    memcpy(&inl, &host_sys[28], 4);
    // camera_tracker.cpp:65-69: inliers_ / error_ are recorded at the LAST ITERATION OF LEVEL 0 only; with no level-0
    // iteration the reference keeps its previous values, so the outputs are left untouched then.  (Levels run coarse to
    // fine, so the last evaluated system is level 0's last iteration whenever level 0 iterates at all.)
    if (levels[0].iterations > 0) {
      if (inlier_fraction) *inlier_fraction = last_area ? (float)inl / (float)last_area : 0.0f;
      if (error) *error = inl != 0 ? host_sys[27] / (float)inl : INFINITY;
    }
    if (last_system) memcpy(last_system, host_sys, sizeof(float) * 29);
    if (history) memcpy(history, h->track_host + 8, sizeof(float) * 36 * (size_t)total_iters);
    return DFK_OK;
  } catch (...) {  // std::bad_alloc / std::length_error from host containers must not cross the C ABI
    return oom(h);
  }
}

DfkStatus dfk_se3_warp(DfkHandle h, const float se3[7], const DfkCamera* cam, const DfkImage* img0,
                       const DfkImage* img1, const DfkImage* dpt0, const DfkImage* img2, float* residual,
                       uint64_t* inliers)
{
  try {
    if (!h) return DFK_ERR_INVALID_ARG;
    if (!se3 || !cam || !img0 || !img1 || !dpt0 || !img2 || !residual || !inliers)
      return fail(h, DFK_ERR_INVALID_ARG, "[SE3Aligner::Warp] null argument");
    const uint32_t W = img0->width, H = img0->height;
    if (W == 0 || H == 0 || !img_ok(img0, W, H, 1) || !img_ok(img1, W, H, 1) || !img_ok(dpt0, W, H, 1) ||
        !img_ok(img2, W, H, 1))
      return fail(h, DFK_ERR_INVALID_ARG, "[SE3Aligner::Warp] inconsistent image views");
    if (!cam_ok(cam, W, H)) return fail(h, DFK_ERR_INVALID_ARG, "[SE3Aligner::Warp] camera viewport larger than the image views");
    DeviceGuard guard(h->device);
    // depth <= 0 -> skip ; PixelValid(pix1, 1) (cu_se3aligner.cpp:89-97)
    const PixelCam pc = make_pixel_cam(se3, cam, 1, 0.0f);
    DFK_CUDA(h, launch_warp(pc, (int)W, (int)H, view_of(img0), view_of(img1), view_of(dpt0), (float*)img2->ptr,
                            (uint32_t)(img2->pitch_bytes / 4), h->simple_scratch, h->counter, h->out_dev, h->stream),
             "[SE3Aligner::Warp] Kernel launch failed (kernel_warp_calculate)");
    DfkStatus st = fetch_out(h, 2, "[SE3Aligner::Warp] Kernel launch failed (kernel_finalize_reduction)");
    if (st != DFK_OK) return st;
    *residual = h->out_host[0];
    uint32_t bits;
    memcpy(&bits, &h->out_host[1], 4);
    *inliers = bits;
    return DFK_OK;
  } catch (...) {  // std::bad_alloc / std::length_error from host containers must not cross the C ABI
    return oom(h);
  }
}

DfkStatus dfk_update_depth(DfkHandle h, const float* code, int code_size, const DfkImage* prx_orig,
                           const DfkImage* prx_jac, float avg_dpt, const DfkImage* dpt_out)
{
  try {
    if (!h) return DFK_ERR_INVALID_ARG;
    if (!code || !prx_orig || !prx_jac || !dpt_out) return fail(h, DFK_ERR_INVALID_ARG, "[UpdateDepth] null argument");
    if (code_size < 1 || code_size > 256) return fail(h, DFK_ERR_UNSUPPORTED, "[UpdateDepth] code size out of range");
    const uint32_t W = dpt_out->width, H = dpt_out->height;
    if (W == 0 || H == 0 || !img_ok(prx_orig, W, H, 1) || !img_ok(prx_jac, W, H, code_size) || !img_ok(dpt_out, W, H, 1))
      return fail(h, DFK_ERR_INVALID_ARG, "[UpdateDepth] inconsistent image views");
    DeviceGuard guard(h->device);
    DFK_CUDA(h, cudaMemcpyAsync(h->code_dev, code, sizeof(float) * code_size, cudaMemcpyHostToDevice, h->stream),
             "[UpdateDepth] code upload failed");
    DFK_CUDA(h, launch_update_depth(h->code_dev, code_size, (int)W, (int)H, view_of(prx_orig), view_of(prx_jac),
                                    avg_dpt, (float*)dpt_out->ptr, (uint32_t)(dpt_out->pitch_bytes / 4), h->stream),
             "[UpdateDepth] kernel launch failed");
    h->launches += 1;
    return DFK_OK;
  } catch (...) {  // std::bad_alloc / std::length_error from host containers must not cross the C ABI
    return oom(h);
  }
}

DfkStatus dfk_sobel_gradients(DfkHandle h, const DfkImage* img, const DfkImage* grad)
{
  try {
    if (!h) return DFK_ERR_INVALID_ARG;
    if (!img || !grad) return fail(h, DFK_ERR_INVALID_ARG, "[SobelGradients] null argument");
    const uint32_t W = img->width, H = img->height;
    if (W == 0 || H == 0 || !img_ok(img, W, H, 1) || !img_ok(grad, W, H, 2))
      return fail(h, DFK_ERR_INVALID_ARG, "[SobelGradients] inconsistent image views");
    DeviceGuard guard(h->device);
    DFK_CUDA(h, launch_sobel((int)W, (int)H, view_of(img), (float*)grad->ptr, (uint32_t)(grad->pitch_bytes / 4),
                             h->stream),
             "Kernel launch failed (kernel_sobel_gradients)");
    h->launches += 1;
    return DFK_OK;
  } catch (...) {  // std::bad_alloc / std::length_error from host containers must not cross the C ABI
    return oom(h);
  }
}

DfkStatus dfk_gaussian_blur_down(DfkHandle h, const DfkImage* in, const DfkImage* out)
{
  try {
    if (!h) return DFK_ERR_INVALID_ARG;
    if (!in || !out) return fail(h, DFK_ERR_INVALID_ARG, "[GaussianBlurDown] null argument");
    if (in->width == 0 || in->height == 0 || out->width == 0 || out->height == 0 ||
        !img_ok(in, in->width, in->height, 1) || !img_ok(out, out->width, out->height, 1))
      return fail(h, DFK_ERR_INVALID_ARG, "[GaussianBlurDown] inconsistent image views");
    DeviceGuard guard(h->device);
    DFK_CUDA(h, launch_blur_down((int)in->width, (int)in->height, view_of(in), (int)out->width, (int)out->height,
                                 (float*)out->ptr, (uint32_t)(out->pitch_bytes / 4), h->stream),
             "Kernel launch failed (kernel_gaussian_blur_down)");
    h->launches += 1;
    return DFK_OK;
  } catch (...) {  // std::bad_alloc / std::length_error from host containers must not cross the C ABI
    return oom(h);
  }
}

DfkStatus dfk_build_image_pyramid(DfkHandle h, const DfkImage* imgs, const DfkImage* grads, int levels)
{
  try {
    if (!h) return DFK_ERR_INVALID_ARG;
    if (!imgs || levels <= 0) return fail(h, DFK_ERR_INVALID_ARG, "[BuildImagePyramid] null argument / no levels");
    for (int l = 1; l < levels; ++l) {
      DfkStatus st = dfk_gaussian_blur_down(h, &imgs[l - 1], &imgs[l]);
      if (st != DFK_OK) return st;
    }
    if (grads)
      for (int l = 0; l < levels; ++l) {
        DfkStatus st = dfk_sobel_gradients(h, &imgs[l], &grads[l]);
        if (st != DFK_OK) return st;
      }
    return DFK_OK;
  } catch (...) {  // std::bad_alloc / std::length_error from host containers must not cross the C ABI
    return oom(h);
  }
}

DfkStatus dfk_squared_error(DfkHandle h, const DfkImage* a, const DfkImage* b, float* out)
{
  try {
    if (!h) return DFK_ERR_INVALID_ARG;
    if (!a || !b || !out) return fail(h, DFK_ERR_INVALID_ARG, "[SquaredError] null argument");
    const uint32_t W = a->width, H = a->height;
    if (W == 0 || H == 0 || !img_ok(a, W, H, 1) || !img_ok(b, W, H, 1))
      return fail(h, DFK_ERR_INVALID_ARG, "[SquaredError] inconsistent image views");
    DeviceGuard guard(h->device);
    DFK_CUDA(h, launch_squared_error((int)W, (int)H, view_of(a), view_of(b), h->simple_scratch, h->counter, h->out_dev,
                                     h->stream),
             "[SquaredError] kernel launch failed");
    DfkStatus st = fetch_out(h, 1, "[SquaredError] kernel launch failed");
    if (st != DFK_OK) return st;
    *out = h->out_host[0];
    return DFK_OK;
  } catch (...) {  // std::bad_alloc / std::length_error from host containers must not cross the C ABI
    return oom(h);
  }
}

DfkStatus dfk_window_create(DfkHandle h, const DfkWindowDesc* d, DfkWindow** out)
{
  try {
    if (!h) return DFK_ERR_INVALID_ARG;
    if (!d || !out) return fail(h, DFK_ERR_INVALID_ARG, "[Window] null argument");
    *out = nullptr;
    const int K = d->num_keyframes, P = d->num_pairs, n = d->num_items;
    if (K <= 0 || P <= 0 || n <= 0 || !d->pair_k0 || !d->pair_k1 || !d->item_pair || !d->item_width || !d->item_height)
      return fail(h, DFK_ERR_INVALID_ARG, "[Window] empty window / null index array");
    if (!dfk_sfm_supports_code_size(d->code_size))
      return fail(h, DFK_ERR_UNSUPPORTED, "[Window] no RunStep kernel for code size " + std::to_string(d->code_size));
    for (int p = 0; p < P; ++p)
      if (d->pair_k0[p] < 0 || d->pair_k0[p] >= K || d->pair_k1[p] < 0 || d->pair_k1[p] >= K)
        return fail(h, DFK_ERR_INVALID_ARG, "[Window] pair " + std::to_string(p) + " names a keyframe outside the window");
    for (int i = 0; i < n; ++i)
      if (d->item_pair[i] < 0 || d->item_pair[i] >= P || d->item_width[i] <= 0 || d->item_height[i] <= 0)
        return fail(h, DFK_ERR_INVALID_ARG, "[Window] record " + std::to_string(i) + " names a pair outside the window");
    // CSR lists in item order (the summation order of the gather kernel)
    std::vector<int> kf0_ptr(K + 1, 0), kf1_ptr(K + 1, 0), pair_ptr(P + 1, 0);
    for (int i = 0; i < n; ++i) {
      const int p = d->item_pair[i];
      kf0_ptr[d->pair_k0[p] + 1] += 1;
      kf1_ptr[d->pair_k1[p] + 1] += 1;
      pair_ptr[p + 1] += 1;
    }
    for (int k = 0; k < K; ++k) { kf0_ptr[k + 1] += kf0_ptr[k]; kf1_ptr[k + 1] += kf1_ptr[k]; }
    for (int p = 0; p < P; ++p) pair_ptr[p + 1] += pair_ptr[p];
    std::vector<int> kf0_items(n), kf1_items(n), pair_items(n);
    {
      std::vector<int> c0(kf0_ptr.begin(), kf0_ptr.end() - 1), c1(kf1_ptr.begin(), kf1_ptr.end() - 1),
          cp(pair_ptr.begin(), pair_ptr.end() - 1);
      for (int i = 0; i < n; ++i) {
        const int p = d->item_pair[i];
        kf0_items[c0[d->pair_k0[p]]++] = i;
        kf1_items[c1[d->pair_k1[p]]++] = i;
        pair_items[cp[p]++] = i;
      }
    }
    std::vector<int> blob;
    blob.reserve((size_t)2 * (K + 1) + (P + 1) + 3 * (size_t)n);
    const size_t o_kf0p = 0;
    blob.insert(blob.end(), kf0_ptr.begin(), kf0_ptr.end());
    const size_t o_kf0i = blob.size();
    blob.insert(blob.end(), kf0_items.begin(), kf0_items.end());
    const size_t o_kf1p = blob.size();
    blob.insert(blob.end(), kf1_ptr.begin(), kf1_ptr.end());
    const size_t o_kf1i = blob.size();
    blob.insert(blob.end(), kf1_items.begin(), kf1_items.end());
    const size_t o_pp = blob.size();
    blob.insert(blob.end(), pair_ptr.begin(), pair_ptr.end());
    const size_t o_pi = blob.size();
    blob.insert(blob.end(), pair_items.begin(), pair_items.end());
    std::vector<float> areas(n);
    for (int i = 0; i < n; ++i) areas[i] = (float)d->item_width[i] * (float)d->item_height[i];

    DeviceGuard guard(h->device);
    DfkWindow* w = new (std::nothrow) DfkWindow();
    if (!w) return oom(h);
    w->device = h->device;
    cudaError_t e = cudaMalloc((void**)&w->ints, blob.size() * sizeof(int));
    if (e == cudaSuccess) e = cudaMalloc((void**)&w->areas, areas.size() * sizeof(float));
    if (e == cudaSuccess) e = cudaMemcpy(w->ints, blob.data(), blob.size() * sizeof(int), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(w->areas, areas.data(), areas.size() * sizeof(float), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
      cudaFree(w->ints);
      cudaFree(w->areas);
      delete w;
      return cuda_fail(h, e, "[Window] index upload failed");
    }
    w->dev.num_keyframes = K; w->dev.num_pairs = P; w->dev.num_items = n; w->dev.code_size = d->code_size;
    w->dev.kf0_ptr = w->ints + o_kf0p; w->dev.kf0_items = w->ints + o_kf0i;
    w->dev.kf1_ptr = w->ints + o_kf1p; w->dev.kf1_items = w->ints + o_kf1i;
    w->dev.pair_ptr = w->ints + o_pp; w->dev.pair_items = w->ints + o_pi;
    w->dev.item_area = w->areas;
    const size_t B = 6 + (size_t)d->code_size;
    w->floats = (size_t)K * (B * B + B) + (size_t)P * 6 * B + 2;
    *out = w;
    return DFK_OK;
  } catch (...) {
    return oom(h);
  }
}

DfkStatus dfk_window_destroy(DfkHandle h, DfkWindow* w)
{
  (void)h;
  if (!w) return DFK_OK;
  DeviceGuard guard(w->device);
  cudaFree(w->ints);
  cudaFree(w->areas);
  delete w;
  return DFK_OK;
}

size_t dfk_window_floats(const DfkWindow* w) { return w ? w->floats : 0; }

DfkStatus dfk_window_assemble(DfkHandle h, const DfkWindow* w, const float* records_dev, float* window_dev)
{
  try {
    if (!h) return DFK_ERR_INVALID_ARG;
    if (!w || !records_dev || !window_dev) return fail(h, DFK_ERR_INVALID_ARG, "[Window] null argument");
    if (w->device != h->device) return fail(h, DFK_ERR_INVALID_ARG, "[Window] window and handle live on different devices");
    DeviceGuard guard(h->device);
    DFK_CUDA(h, launch_window_assemble(w->dev, records_dev, window_dev, h->stream), "[Window] kernel launch failed");
    h->launches += 1;
    return DFK_OK;
  } catch (...) {
    return oom(h);
  }
}

// ---------------------------------------------------------------------------------------------- streaming from host
DfkStatus dfk_sfm_stream_create(DfkHandle h, int code_size, int max_items, size_t max_bytes, int depth, DfkSfmStream** out)
{
  try {
    if (!h) return DFK_ERR_INVALID_ARG;
    if (!out || max_items <= 0 || max_bytes == 0 || depth < 1 || depth > 16)
      return fail(h, DFK_ERR_INVALID_ARG, "[SfmStream] bad argument (1 <= depth <= 16, max_items > 0, max_bytes > 0)");
    *out = nullptr;
    if (!dfk_sfm_supports_code_size(code_size))
      return fail(h, DFK_ERR_UNSUPPORTED, "[SfmStream] no RunStep kernel for code size " + std::to_string(code_size));
    DeviceGuard guard(h->device);
    DfkSfmStream* s = new (std::nothrow) DfkSfmStream();
    if (!s) return oom(h);
    s->device = h->device; s->code_size = code_size; s->max_items = max_items; s->depth = depth;
    // initial slot size (a hint): staged images are padded to a 256-byte row pitch and offset, scratch images (valid0,
    // decoded depth) ride along; a submission that needs more grows its slot
    s->max_bytes = max_bytes + max_bytes / 4 + (size_t)max_items * 8 * 4096;
    s->slots.resize(depth);
    s->dev_items.resize(max_items);
    const size_t rec = (size_t)DFK_SFM_RECORD_FLOATS(code_size) * max_items * sizeof(float);
    cudaError_t e = cudaStreamCreateWithFlags(&s->copy_stream, cudaStreamNonBlocking);
    for (int k = 0; k < depth && e == cudaSuccess; ++k) {
      DfkSfmStream::Slot& sl = s->slots[k];
      e = cudaMalloc((void**)&sl.dev, s->max_bytes);
      if (e == cudaSuccess) sl.cap = s->max_bytes;
      if (e == cudaSuccess) e = cudaMalloc((void**)&sl.rec_dev, rec);
      if (e == cudaSuccess) e = cudaMallocHost((void**)&sl.rec_host, rec);
      if (e == cudaSuccess) e = cudaEventCreateWithFlags(&sl.uploaded, cudaEventDisableTiming);
      if (e == cudaSuccess) e = cudaEventCreateWithFlags(&sl.done, cudaEventDisableTiming);
    }
    if (e != cudaSuccess) {
      dfk_sfm_stream_destroy(h, s);
      return cuda_fail(h, e, "[SfmStream] allocation failed");
    }
    *out = s;
    return DFK_OK;
  } catch (...) {
    return oom(h);
  }
}

DfkStatus dfk_sfm_stream_destroy(DfkHandle h, DfkSfmStream* s)
{
  (void)h;
  if (!s) return DFK_OK;
  DeviceGuard guard(s->device);
  if (s->copy_stream) cudaStreamSynchronize(s->copy_stream);
  for (auto& sl : s->slots) {
    if (sl.done) { cudaEventSynchronize(sl.done); cudaEventDestroy(sl.done); }
    if (sl.uploaded) cudaEventDestroy(sl.uploaded);
    cudaFree(sl.dev);
    cudaFree(sl.rec_dev);
    if (sl.rec_host) cudaFreeHost(sl.rec_host);
  }
  if (s->copy_stream) cudaStreamDestroy(s->copy_stream);
  delete s;
  return DFK_OK;
}

DfkStatus dfk_sfm_stream_submit(DfkHandle h, DfkSfmStream* s, const DfkSfmWorkItem* items, int n, uint64_t* ticket)
{
  try {
    if (!h) return DFK_ERR_INVALID_ARG;
    if (!s || !items || !ticket || n <= 0) return fail(h, DFK_ERR_INVALID_ARG, "[SfmStream] null argument / empty submission");
    if (n > s->max_items) return fail(h, DFK_ERR_INVALID_ARG, "[SfmStream] more work items than the stream was created for");
    if (s->device != h->device) return fail(h, DFK_ERR_INVALID_ARG, "[SfmStream] stream and handle live on different devices");
    DfkSfmStream::Slot& sl = s->slots[s->next_ticket % (uint64_t)s->depth];
    if (sl.busy)
      return fail(h, DFK_ERR_INVALID_ARG,
                  "[SfmStream] " + std::to_string(s->depth) + " submissions outstanding: wait for ticket " +
                      std::to_string(sl.ticket) + " first");
    DeviceGuard guard(h->device);
    cudaError_t err = cudaSuccess;
    // one image: host view -> 256-byte aligned, 256-byte pitched device view inside the slot.  upload == false: device
    // scratch.  Two passes over the items: sizes first (the slot grows if it has to), then the copies.
    for (int pass = 0; pass < 2; ++pass) {
      size_t cursor = 0;
      auto stage = [&](const DfkImage& src, uint32_t floats_per_px, bool upload) -> DfkImage {
        DfkImage d{};
        const size_t row = (size_t)src.width * floats_per_px * sizeof(float);
        // rows that are already 16-byte multiples stay dense on the device (what the bulk-copy loaders of the kernels
        // need), so a dense host image travels as ONE linear copy: per-row DMA descriptors cost ~30 % of the PCIe rate on
        // the small pyramid levels; odd widths get a 256-byte pitch and a 2-D copy
        const size_t pitch = (row % 16 == 0) ? row : ((row + 255) & ~(size_t)255);
        cursor = (cursor + 255) & ~(size_t)255;
        d.ptr = sl.dev + cursor;
        d.pitch_bytes = pitch;
        d.width = src.width;
        d.height = src.height;
        cursor += pitch * src.height;
        if (pass == 1 && upload && err == cudaSuccess) {
          if (pitch == row && src.pitch_bytes == row)
            err = cudaMemcpyAsync(d.ptr, src.ptr, row * src.height, cudaMemcpyHostToDevice, s->copy_stream);
          else
            err = cudaMemcpy2DAsync(d.ptr, pitch, src.ptr, src.pitch_bytes, row, src.height, cudaMemcpyHostToDevice,
                                    s->copy_stream);
        }
        return d;
      };
      for (int i = 0; i < n; ++i) {
        const DfkSfmWorkItem& w = items[i];
        const uint32_t W = w.img0.width, H = w.img0.height;
        const bool fused = w.code != nullptr;
        if (pass == 0 && (W == 0 || H == 0 || !img_ok(&w.img0, W, H, 1) || !img_ok(&w.img1, W, H, 1) ||
                          !img_ok(&w.prx0_jac, W, H, s->code_size) || !img_ok(&w.grad1, W, H, 2) ||
                          (fused ? !img_ok(&w.prx_orig, W, H, 1) : !img_ok(&w.dpt0, W, H, 1))))
          return fail(h, DFK_ERR_INVALID_ARG, "[SfmStream] inconsistent host image views in work item " + std::to_string(i));
        DfkSfmWorkItem& d = s->dev_items[i];
        d = w;
        d.img0 = stage(w.img0, 1, true);
        d.img1 = stage(w.img1, 1, true);
        d.prx0_jac = stage(w.prx0_jac, (uint32_t)s->code_size, true);
        d.grad1 = stage(w.grad1, 2, true);
        const DfkImage scalar = w.img0;  // geometry of a scalar scratch image
        d.valid0 = stage(scalar, 1, false);
        if (fused) {
          d.prx_orig = stage(w.prx_orig, 1, true);
          d.dpt0 = stage(scalar, 1, false);  // the decoded depth stays on the device
        } else {
          d.dpt0 = stage(w.dpt0, 1, true);
        }
      }
      if (pass == 0 && cursor > sl.cap) {  // the slot is idle (not busy): its memory can be replaced
        cudaFree(sl.dev);
        sl.dev = nullptr;
        sl.cap = 0;
        DFK_CUDA(h, cudaMalloc((void**)&sl.dev, cursor), "[SfmStream] slot allocation failed");
        sl.cap = cursor;
      }
    }
    if (err != cudaSuccess) return cuda_fail(h, err, "[SfmStream] upload failed");
    DFK_CUDA(h, cudaEventRecord(sl.uploaded, s->copy_stream), "[SfmStream] event record failed");
    DFK_CUDA(h, cudaStreamWaitEvent(h->stream, sl.uploaded, 0), "[SfmStream] stream wait failed");
    DfkStatus st = run_batch(h, s->dev_items.data(), n, s->code_size, sl.rec_dev);
    if (st != DFK_OK) return st;
    DFK_CUDA(h, cudaMemcpyAsync(sl.rec_host, sl.rec_dev, (size_t)DFK_SFM_RECORD_FLOATS(s->code_size) * n * sizeof(float),
                                cudaMemcpyDeviceToHost, h->stream),
             "[SfmStream] result download failed");
    DFK_CUDA(h, cudaEventRecord(sl.done, h->stream), "[SfmStream] event record failed");
    // the NEXT use of this slot's staging memory (depth submissions later) is host-ordered behind wait(ticket); the copy
    // stream itself must not run ahead of the evaluation that still reads the slot it is about to overwrite
    sl.busy = true;
    sl.n = n;
    sl.ticket = s->next_ticket;
    *ticket = s->next_ticket++;
    return DFK_OK;
  } catch (...) {
    return oom(h);
  }
}

DfkStatus dfk_sfm_stream_wait(DfkHandle h, DfkSfmStream* s, uint64_t ticket, float* records_host)
{
  try {
    if (!h) return DFK_ERR_INVALID_ARG;
    if (!s || !records_host) return fail(h, DFK_ERR_INVALID_ARG, "[SfmStream] null argument");
    DfkSfmStream::Slot& sl = s->slots[ticket % (uint64_t)s->depth];
    if (!sl.busy || sl.ticket != ticket || ticket != s->next_wait)
      return fail(h, DFK_ERR_INVALID_ARG, "[SfmStream] tickets must be waited for once, in submission order");
    DeviceGuard guard(h->device);
    DFK_CUDA(h, cudaEventSynchronize(sl.done), "[SfmStream] kernel launch failed");
    memcpy(records_host, sl.rec_host, (size_t)DFK_SFM_RECORD_FLOATS(s->code_size) * sl.n * sizeof(float));
    sl.busy = false;
    s->next_wait += 1;
    return DFK_OK;
  } catch (...) {
    return oom(h);
  }
}

DfkStatus dfk_depth_run_step(DfkHandle h, const float* code, int code_size, const DfkImage* target_dpt,
                             const DfkImage* prx_orig, const DfkImage* prx_jac, float* JtJ, float* Jtr, float* residual,
                             uint64_t* inliers)
{
  try {
    if (!h) return DFK_ERR_INVALID_ARG;
    if (!code || !target_dpt || !prx_orig || !prx_jac || !JtJ || !Jtr || !residual || !inliers)
      return fail(h, DFK_ERR_INVALID_ARG, "[DepthAligner::RunStep] null argument");
    // CHECK_EQ(codesize, CS) (cu_depthaligner.cpp:90-91): the code size must be one this build instantiates
    if (!(code_size == 8 || code_size == 16 || code_size == 32 || code_size == 64 || code_size == 128))
      return fail(h, DFK_ERR_UNSUPPORTED,
                  "DepthAligner used with a different code size than it was compiled for: " + std::to_string(code_size));
    const uint32_t W = target_dpt->width, H = target_dpt->height;
    if (W == 0 || H == 0 || !img_ok(target_dpt, W, H, 1) || !img_ok(prx_orig, W, H, 1) || !img_ok(prx_jac, W, H, code_size))
      return fail(h, DFK_ERR_INVALID_ARG, "[DepthAligner::RunStep] inconsistent image views");
    DeviceGuard guard(h->device);
    const int area = (int)(W * H);
    const int blocks = std::max(1, std::min(2 * h->num_sms, (area + 63) / 64));
    const size_t NH = (size_t)code_size * (code_size + 1) / 2, REC = NH + code_size + 2;
    DFK_CUDA(h, ensure(&h->partials_dev, &h->partials_cap, (size_t)blocks * depth_partial_floats(code_size)),
             "[DepthAligner::RunStep] scratch allocation failed");
    DFK_CUDA(h, ensure(&h->records_dev, &h->records_cap, REC), "[DepthAligner::RunStep] scratch allocation failed");
    if (h->records_host_cap < REC) {
      if (h->records_host) cudaFreeHost(h->records_host);
      h->records_host = nullptr;
      h->records_host_cap = 0;
      DFK_CUDA(h, cudaMallocHost((void**)&h->records_host, REC * sizeof(float)), "[DepthAligner::RunStep] pinned allocation failed");
      h->records_host_cap = REC;
    }
    DFK_CUDA(h, cudaMemcpyAsync(h->code_dev, code, sizeof(float) * code_size, cudaMemcpyHostToDevice, h->stream),
             "[DepthAligner::RunStep] code upload failed");
    DFK_CUDA(h, launch_depth_step(h->code_dev, code_size, (int)W, (int)H, view_of(target_dpt), view_of(prx_orig),
                                  view_of(prx_jac), h->params.sfmparams.avg_dpt, h->partials_dev, h->counter, h->records_dev,
                                  blocks, h->stream),
             "[DepthAligner::RunStep] kernel launch failed");
    h->launches += 1;
    DFK_CUDA(h, cudaMemcpyAsync(h->records_host, h->records_dev, REC * sizeof(float), cudaMemcpyDeviceToHost, h->stream),
             "[DepthAligner::RunStep] kernel launch failed");
    DFK_CUDA(h, cudaStreamSynchronize(h->stream), "[DepthAligner::RunStep] kernel launch failed");
    memcpy(JtJ, h->records_host, sizeof(float) * NH);
    memcpy(Jtr, h->records_host + NH, sizeof(float) * code_size);
    *residual = h->records_host[NH + code_size];
    uint32_t bits;
    memcpy(&bits, &h->records_host[NH + code_size + 1], 4);
    *inliers = bits;
    return DFK_OK;
  } catch (...) {
    return oom(h);
  }
}

DfkStatus dfk_reprojection_linearize(DfkHandle h, const float pose0[7], const float pose1[7], const float* code0,
                                     int code_size, const DfkCamera* cam, const DfkImage* prx_orig, const DfkImage* prx_jac,
                                     int num_matches, const float* query_xy, const float* train_xy, float cauchy_delta,
                                     float sigma, float* rows, float* total_err)
{
  try {
    if (!h) return DFK_ERR_INVALID_ARG;
    if (!pose0 || !pose1 || !code0 || !cam || !prx_orig || !prx_jac || !query_xy || !train_xy || !rows || !total_err)
      return fail(h, DFK_ERR_INVALID_ARG, "[ReprojectionFactor::linearize] null argument");
    if (!(code_size == 8 || code_size == 16 || code_size == 32 || code_size == 64 || code_size == 128))
      return fail(h, DFK_ERR_UNSUPPORTED, "[ReprojectionFactor::linearize] code size not instantiated: " + std::to_string(code_size));
    if (num_matches <= 0 || !(sigma > 0.0f))
      return fail(h, DFK_ERR_INVALID_ARG, "[ReprojectionFactor::linearize] no matches / non-positive sigma");
    const uint32_t W = prx_orig->width, H = prx_orig->height;
    if (W == 0 || H == 0 || !img_ok(prx_orig, W, H, 1) || !img_ok(prx_jac, W, H, code_size))
      return fail(h, DFK_ERR_INVALID_ARG, "[ReprojectionFactor::linearize] inconsistent image views");
    DeviceGuard guard(h->device);
    const size_t M = (size_t)num_matches, RW = 13 + (size_t)code_size;
    const size_t n_in = 4 * M, n_out = 2 * M * RW + M;
    DFK_CUDA(h, ensure(&h->sparse_dev, &h->sparse_cap, n_in + n_out), "[ReprojectionFactor::linearize] scratch allocation failed");
    if (h->sparse_host_cap < n_in + n_out) {
      if (h->sparse_host) cudaFreeHost(h->sparse_host);
      h->sparse_host = nullptr;
      h->sparse_host_cap = 0;
      DFK_CUDA(h, cudaMallocHost((void**)&h->sparse_host, (n_in + n_out) * sizeof(float)),
               "[ReprojectionFactor::linearize] pinned allocation failed");
      h->sparse_host_cap = n_in + n_out;
    }
    SparsePose sp;
    float p10[7];
    relative_pose(pose1, pose0, p10, sp.P1, sp.P0);  // RelativePose(p1, p0, pose10_J_pose1, pose10_J_pose0), :189-190
    for (int k = 0; k < 4; ++k) sp.q[k] = p10[k];
    for (int k = 0; k < 3; ++k) sp.t[k] = p10[4 + k];
    quat_to_matrix(p10, sp.R);
    sp.fx = cam->fx; sp.fy = cam->fy; sp.u0 = cam->u0; sp.v0 = cam->v0;
    memcpy(h->sparse_host, query_xy, 2 * M * sizeof(float));
    memcpy(h->sparse_host + 2 * M, train_xy, 2 * M * sizeof(float));
    float* d_query = h->sparse_dev;
    float* d_train = d_query + 2 * M;
    float* d_rows = d_train + 2 * M;
    float* d_err2 = d_rows + 2 * M * RW;
    DFK_CUDA(h, cudaMemcpyAsync(h->code_dev, code0, sizeof(float) * code_size, cudaMemcpyHostToDevice, h->stream),
             "[ReprojectionFactor::linearize] code upload failed");
    DFK_CUDA(h, cudaMemcpyAsync(d_query, h->sparse_host, n_in * sizeof(float), cudaMemcpyHostToDevice, h->stream),
             "[ReprojectionFactor::linearize] match upload failed");
    DFK_CUDA(h, launch_reprojection_rows(sp, h->code_dev, code_size, view_of(prx_orig), view_of(prx_jac), (int)W, (int)H,
                                         num_matches, d_query, d_train, cauchy_delta, sigma, h->params.sfmparams.avg_dpt,
                                         d_rows, d_err2, h->stream),
             "[ReprojectionFactor::linearize] kernel launch failed");
    h->launches += 1;
    DFK_CUDA(h, cudaMemcpyAsync(h->sparse_host + n_in, d_rows, n_out * sizeof(float), cudaMemcpyDeviceToHost, h->stream),
             "[ReprojectionFactor::linearize] result download failed");
    DFK_CUDA(h, cudaStreamSynchronize(h->stream), "[ReprojectionFactor::linearize] kernel launch failed");
    memcpy(rows, h->sparse_host + n_in, 2 * M * RW * sizeof(float));
    float tot = 0.0f;  // Scalar total_err accumulated in match order (:179,242)
    const float* e2 = h->sparse_host + n_in + 2 * M * RW;
    for (size_t i = 0; i < M; ++i) tot += e2[i];
    *total_err = tot;
    return DFK_OK;
  } catch (...) {
    return oom(h);
  }
}

DfkStatus dfk_sparse_geometric_linearize(DfkHandle h, const float pose0[7], const float pose1[7], const float* code0,
                                         const float* code1, int code_size, const DfkCamera* cam, const DfkImage* prx0_orig,
                                         const DfkImage* prx0_jac, const DfkImage* prx1_orig, const DfkImage* prx1_jac,
                                         const DfkImage* dpt_grad1, int num_points, const int* points_xy, float huber_delta,
                                         float* rows, int* num_valid)
{
  try {
    if (!h) return DFK_ERR_INVALID_ARG;
    if (!pose0 || !pose1 || !code0 || !code1 || !cam || !prx0_orig || !prx0_jac || !prx1_orig || !prx1_jac || !dpt_grad1 ||
        !points_xy || !rows)
      return fail(h, DFK_ERR_INVALID_ARG, "[SparseGeometricFactor::linearize] null argument");
    if (!(code_size == 8 || code_size == 16 || code_size == 32 || code_size == 64 || code_size == 128))
      return fail(h, DFK_ERR_UNSUPPORTED, "[SparseGeometricFactor::linearize] code size not instantiated: " + std::to_string(code_size));
    if (num_points <= 0 || !(huber_delta > 0.0f))
      return fail(h, DFK_ERR_INVALID_ARG, "[SparseGeometricFactor::linearize] no points / non-positive huber delta");
    const uint32_t W = prx0_orig->width, H = prx0_orig->height;
    if (W == 0 || H == 0 || !img_ok(prx0_orig, W, H, 1) || !img_ok(prx0_jac, W, H, code_size) || !img_ok(prx1_orig, W, H, 1) ||
        !img_ok(prx1_jac, W, H, code_size) || !img_ok(dpt_grad1, W, H, 2))
      return fail(h, DFK_ERR_INVALID_ARG, "[SparseGeometricFactor::linearize] inconsistent image views");
    if (!cam_ok(cam, W, H))  // the nearest-neighbour lookups in keyframe 1 index with the camera's validity window
      return fail(h, DFK_ERR_INVALID_ARG, "[SparseGeometricFactor::linearize] camera larger than the image views");
    DeviceGuard guard(h->device);
    const size_t M = (size_t)num_points, RW = 13 + 2 * (size_t)code_size;
    const size_t n_in = 2 * M, n_out = M * RW;  // ints and floats are both 4 bytes
    DFK_CUDA(h, ensure(&h->sparse_dev, &h->sparse_cap, n_in + n_out), "[SparseGeometricFactor::linearize] scratch allocation failed");
    if (h->sparse_host_cap < n_in + n_out) {
      if (h->sparse_host) cudaFreeHost(h->sparse_host);
      h->sparse_host = nullptr;
      h->sparse_host_cap = 0;
      DFK_CUDA(h, cudaMallocHost((void**)&h->sparse_host, (n_in + n_out) * sizeof(float)),
               "[SparseGeometricFactor::linearize] pinned allocation failed");
      h->sparse_host_cap = n_in + n_out;
    }
    SparsePose sp;
    float p10[7];
    relative_pose(pose1, pose0, p10, sp.P1, sp.P0);  // RelativePose(p1, p0, pose10_J_pose1, pose10_J_pose0), :176-178
    for (int k = 0; k < 4; ++k) sp.q[k] = p10[k];
    for (int k = 0; k < 3; ++k) sp.t[k] = p10[4 + k];
    quat_to_matrix(p10, sp.R);
    sp.fx = cam->fx; sp.fy = cam->fy; sp.u0 = cam->u0; sp.v0 = cam->v0;
    memcpy(h->sparse_host, points_xy, 2 * M * sizeof(int));
    float codes[256] = {0};  // code_dev holds 256 floats: code0 at 0, code1 at 128
    memcpy(codes, code0, sizeof(float) * code_size);
    memcpy(codes + 128, code1, sizeof(float) * code_size);
    int* d_points = reinterpret_cast<int*>(h->sparse_dev);
    float* d_rows = h->sparse_dev + n_in;
    DFK_CUDA(h, cudaMemcpyAsync(h->code_dev, codes, sizeof(codes), cudaMemcpyHostToDevice, h->stream),
             "[SparseGeometricFactor::linearize] code upload failed");
    DFK_CUDA(h, cudaMemcpyAsync(d_points, h->sparse_host, n_in * sizeof(float), cudaMemcpyHostToDevice, h->stream),
             "[SparseGeometricFactor::linearize] point upload failed");
    DFK_CUDA(h, launch_sparse_geometric_rows(sp, cam->width, cam->height, h->code_dev, h->code_dev + 128, code_size,
                                             view_of(prx0_orig), view_of(prx0_jac), view_of(prx1_orig), view_of(prx1_jac),
                                             view_of(dpt_grad1), (int)W, (int)H, num_points, d_points, huber_delta,
                                             h->params.sfmparams.avg_dpt, d_rows, h->stream),
             "[SparseGeometricFactor::linearize] kernel launch failed");
    h->launches += 1;
    DFK_CUDA(h, cudaMemcpyAsync(h->sparse_host + n_in, d_rows, n_out * sizeof(float), cudaMemcpyDeviceToHost, h->stream),
             "[SparseGeometricFactor::linearize] result download failed");
    DFK_CUDA(h, cudaStreamSynchronize(h->stream), "[SparseGeometricFactor::linearize] kernel launch failed");
    memcpy(rows, h->sparse_host + n_in, n_out * sizeof(float));
    if (num_valid) {
      int nv = 0;
      for (size_t i = 0; i < M; ++i) {
        const float* r = rows + i * RW;
        bool any = false;
        for (size_t k = 0; k < RW && !any; ++k) any = r[k] != 0.0f;
        nv += any ? 1 : 0;
      }
      *num_valid = nv;
    }
    return DFK_OK;
  } catch (...) {
    return oom(h);
  }
}

}  // extern "C"
