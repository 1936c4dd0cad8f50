// dfk_sfm_fp32.cu -- SfmAligner::RunStep hot path, fp32 CUDA-core Gram variant (sm_100a).
//
// Replaces kernel_step_calculate + DenseSfm + runReductions/finalizeReduction +
// kernel_finalize_reduction of the reference (sources/cuda/cu_sfmaligner.cpp:40-70,
// sources/common/algorithm/dense_sfm.h:133-201, sources/cuda/kernel_utils.h:51-69).
//
// Design (see DESIGN.md "fp32 Gram kernel"):
//   * ONE persistent launch evaluates a whole list of (pair, level) items.  The pixel stream
//     of every item is cut into tiles of 256 linear pixels; CTA c owns a contiguous range of
//     the global tile sequence (static => bitwise reproducible results), tiles inside an item
//     are visited in a strided order so that spatially clustered invalid regions balance out.
//   * Per tile the TMA engine (cp.async.bulk, 1-D row segments) stages the code-Jacobian rows
//     (C contiguous floats per pixel), img0 and dpt0 into a 3-deep shared-memory ring.
//   * 4 "front-end" warps (one thread per pixel) run the exact-order validity chain, gather
//     img1/grad1 bilinearly, form the reduced row  m = w*[ e*jc (C) | a (6) | diff (1) ]  and
//     write it, compacted to valid pixels, K-major into a double-buffered tile M[feature][pixel].
//   * NBLK "Gram" warps each own one 8x8 block of the upper triangle of G = sum m^T m
//     ((7+C)^2, the (6+C) reduced system + gradient + energy of SURVEY Appendix A) with lanes
//     striding over pixels; 64 register accumulators per thread, operands via conflict-free
//     LDS.  Accumulators are reduce-scattered across lanes only when the CTA leaves an item.
//   * A second, wide kernel sums the per-CTA partials in fixed order and expands the reduced
//     system to the reference's (12+C) layout with the host-computed relative-pose Jacobians:
//     JtJ = E^T G E,  E = [[P0,P1,0],[0,0,I]].
#include <cuda_runtime.h>
#include <stdint.h>

#include "dfk_async.cuh"
#include "dfk_geom.cuh"
#include "dfk_internal.h"
#include "dfk_tile_stage.cuh"

namespace dfk {

namespace {

constexpr int kStages = 3;
constexpr int kFeWarps = 8;
constexpr int kFeThreads = kFeWarps * 32;
static_assert(kTilePixels == kFeThreads, "one front-end thread per tile pixel");
constexpr int kMaxTab = 1024;
constexpr int kMaxCode = 32;  // largest code size of this kernel (fused depth decode keeps the code in shared memory)  // per-item tables of the normalised ray coordinates (columns / rows)

struct TileMeta {
  int nvalid;
  int item_changed;  // 1 if this tile starts a new item for this CTA
  int slot;          // partial slot of the tile's item
  int pad;
};

using ItemSmem = StagedItem<kMaxCode>;  // dfk_tile_stage.cuh

template <int C>
struct Smem {
  using Cfg = SfmCfg<C>;
  alignas(128) float jc[kStages][kTilePixels * C];
  alignas(16) float img0[kStages][kTilePixels];
  alignas(16) float dpt0[kStages][kTilePixels];
  alignas(16) float M[2][Cfg::NFP * kTilePixels];
  float xn_tab[kMaxTab];  // (x - u0) / fx, IEEE, per column of the current item
  float yn_tab[kMaxTab];  // (y - v0) / fy per row
  alignas(8) uint64_t full_tma[kStages];
  uint64_t m_full[2];
  uint64_t m_empty[2];
  TileMeta meta[2];
  ItemSmem item;
  int cnt[kFeWarps];  // valid counts per front-end warp
};

// reduce-scatter of 64 per-lane accumulators: afterwards lane l holds the warp-wide sums of
// entries 2l and 2l+1 in acc[0], acc[1].  62 shuffles; fixed order => deterministic.
__device__ __forceinline__ void reduce_scatter64(float (&acc)[64], int lane)
{
#pragma unroll
  for (int m = 16, n = 64; m >= 1; m >>= 1, n >>= 1) {
    const bool up = (lane & m) != 0;
#pragma unroll
    for (int i = 0; i < n / 2; ++i) {
      const float send = up ? acc[i] : acc[i + n / 2];
      const float keep = up ? acc[i + n / 2] : acc[i];
      acc[i] = keep + __shfl_xor_sync(0xffffffffu, send, m);
    }
  }
}

template <int C>
__global__ void __launch_bounds__((kFeWarps + SfmCfg<C>::NBLK) * 32, sfm_fp32_ctas_per_sm(C))
sfm_step_fp32_kernel(const SfmItemDev* __restrict__ items, int num_items, int num_tiles, float* __restrict__ partials)
{
  using Cfg = SfmCfg<C>;
  constexpr int NFP = Cfg::NFP;
  constexpr int NB = Cfg::NB;
  constexpr int NBLK = Cfg::NBLK;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  Smem<C>& sm = *reinterpret_cast<Smem<C>*>(smem_raw);

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const int cta = blockIdx.x;
  const int G = gridDim.x;
  const int g_lo = (int)(((long long)cta * num_tiles) / G);
  const int g_hi = (int)(((long long)(cta + 1) * num_tiles) / G);

  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) mbar_init(&sm.full_tma[s], 1);
    for (int b = 0; b < 2; ++b) {
      mbar_init(&sm.m_full[b], kFeWarps);  // one arrival per front-end warp (every arrival wakes the waiters)
      mbar_init(&sm.m_empty[b], NBLK);
    }
    mbar_fence_init();
  }
  __syncthreads();
  if (g_lo >= g_hi) return;

  if (warp < kFeWarps) {
    // ========================================================================= front-end
    // item cursors: `it` for the tile being processed, `it_pf` for the prefetcher
    int it = 0;
    while (it + 1 < num_items && (uint32_t)g_lo >= items[it].tile_begin + items[it].num_tiles) ++it;
    int it_pf = it;
    uint32_t tma_phase_bits = 0;  // bit s = parity to wait for on stage s
    int cur_item = -1;

    // prologue: prefetch the first kStages tiles
    if (tid == 0) {
      for (int j = 0; j < kStages && g_lo + j < g_hi; ++j) {
        const int g = g_lo + j;
        while ((uint32_t)g >= items[it_pf].tile_begin + items[it_pf].num_tiles) ++it_pf;
        if (items[it_pf].flags & ITEM_FLAG_BULK) issue_tile_loads<C, kTilePixels>(sm, items, it_pf, g, j);
      }
    }

    for (int g = g_lo, i = 0; g < g_hi; ++g, ++i) {
      const int st = i % kStages;
      const int buf = i & 1;
      while ((uint32_t)g >= items[it].tile_begin + items[it].num_tiles) ++it;
      const bool changed = (it != cur_item);
      if (changed) {
        named_bar_sync(1, kFeThreads);  // everyone finished reading the previous item's params
        load_item(sm.item, items[it], tid, kFeThreads, cta);
        load_code(sm.item, items[it], C, tid, kFeThreads);
        cur_item = it;
        {
          // normalised ray tables (Reproject's IEEE divisions hoisted out of the pixel loop)
          const SfmItemDev& src = items[it];
          if (src.width <= kMaxTab && src.height <= kMaxTab) {
            for (uint32_t x = tid; x < src.width; x += kFeThreads) sm.xn_tab[x] = ray_coord((float)x, src.u0, src.fx);
            for (uint32_t y = tid; y < src.height; y += kFeThreads) sm.yn_tab[y] = ray_coord((float)y, src.v0, src.fy);
          }
        }
        named_bar_sync(1, kFeThreads);
      }
      const ItemSmem& I = sm.item;
      const uint32_t k = (uint32_t)g - I.tile_begin;
      const uint32_t tau = (uint32_t)(((uint64_t)k * I.perm_mul) % I.num_tiles);
      const uint32_t p0 = tau * kTilePixels;
      const uint32_t n = min((uint32_t)kTilePixels, I.num_pixels - p0);
      const bool bulk = (I.flags & ITEM_FLAG_BULK) != 0;
      if (bulk) {
        mbar_wait(&sm.full_tma[st], (tma_phase_bits >> st) & 1u);
        tma_phase_bits ^= (1u << st);
      } else {
        coop_tile_loads<C, kFeThreads>(sm, I, p0, n, st, tid);
        named_bar_sync(1, kFeThreads);
      }

      // ---- geometry for this thread's pixel --------------------------------------------------
      float feat[8];  // s, wa0..5, wr
      bool ok = false;
      const uint32_t s = tid;
      if (s < n) {
        const uint32_t p = p0 + s;
        const uint32_t y = p / I.width, x = p - y * I.width;
        float d = sm.dpt0[st][s];
        if (I.flags & ITEM_FLAG_FUSED_DEPTH) {
          // the stage holds prx_orig: decode the depth exactly as update_depth_kernel does and publish it
          const float4* row = reinterpret_cast<const float4*>(&sm.jc[st][s * C]);
          const float4* cod = reinterpret_cast<const float4*>(I.code);
          float part[C / 4];
#pragma unroll
          for (int k4 = 0; k4 < C / 4; ++k4) part[k4] = chunk_dot(row[k4], cod[k4]);
          d = prx_to_depth(__fadd_rn(d, butterfly_sum<C / 4>(part)), I.avg_dpt);
          I.dpt_out[(size_t)y * I.dpt_out_pitch + x] = d;
        }
        const bool tab = (I.width <= kMaxTab) && (I.height <= kMaxTab);
        const float xn = tab ? sm.xn_tab[x] : ray_coord((float)x, I.u0, I.fx);
        const float yn = tab ? sm.yn_tab[y] : ray_coord((float)y, I.v0, I.fy);
        const Warped w = warp_ray(xn, yn, d, I.q, I.t, I.fx, I.fy, I.u0, I.v0, I.border, I.ulim, I.vlim, I.min_dpt);
        if (w.valid) {
          ok = true;
          I.valid0[(size_t)y * I.valid0_pitch + x] = 1.0f;  // dense_sfm.h:161
          int ix, iy;
          float fu, fv, gx, gy;
          bilin_setup(w.u, w.v, ix, iy, fu, fv);
          sample_grad(I.grad1, I.grad1_pitch, (I.flags & ITEM_FLAG_GRAD_ALIGNED) != 0, ix, iy, fu, fv, gx, gy);
          const float i1 = sample_scalar(I.img1, I.img1_pitch, ix, iy, fu, fv);
          float a[6], c00, c02, c11, c12;
          pose_jacobian_row(w, I.fx, I.fy, gx, gy, a, c00, c02, c11, c12);
          const float e = prx_jacobian(w, I.R, d, I.avg_dpt, gx, gy, c00, c02, c11, c12);
          const float diff = sm.img0[st][s] - i1;
          const float hw = huber_weight(diff, I.huber_delta);
          feat[0] = hw * e;
#pragma unroll
          for (int j = 0; j < 6; ++j) feat[1 + j] = hw * a[j];
          feat[7] = hw * diff;
        }
      }

      // ---- compaction: valid pixels first --------------------------------------------------
      const unsigned bal = __ballot_sync(0xffffffffu, ok);
      const int before = __popc(bal & ((1u << lane) - 1u));
      const int rank = ok ? before : (lane - before);  // rank among the valid / invalid lanes of this warp
      if (lane == 0) sm.cnt[warp] = __popc(bal);
      // the M buffer we are about to overwrite must have been drained by the Gram warps
      mbar_wait(&sm.m_empty[buf], ((i >> 1) & 1u) ^ 1u);
      named_bar_sync(1, kFeThreads);
      int nvalid = 0, base_valid = 0, base_invalid = 0;
#pragma unroll
      for (int w2 = 0; w2 < kFeWarps; ++w2) {
        const int cvalid = sm.cnt[w2];
        if (w2 == warp) {
          base_valid = nvalid;
          base_invalid = 32 * w2 - nvalid;
        }
        nvalid += cvalid;
      }
      const int padded = (nvalid + 31) & ~31;
      float* Mb = sm.M[buf];
      if (ok) {
        const int idx = base_valid + rank;
        const float sc = feat[0];
        if constexpr (C % 4 == 0 && C >= 4) {
          constexpr int NV = C / 4;
          const int rot = (NV >= 8) ? lane : (lane / (8 / (NV < 8 ? NV : 8)));
          const float4* src = reinterpret_cast<const float4*>(&sm.jc[st][s * C]);
#pragma unroll
          for (int k4 = 0; k4 < NV; ++k4) {
            const int kk4 = (k4 + rot) % NV;
            const float4 v = src[kk4];
            Mb[(kk4 * 4 + 0) * kTilePixels + idx] = sc * v.x;
            Mb[(kk4 * 4 + 1) * kTilePixels + idx] = sc * v.y;
            Mb[(kk4 * 4 + 2) * kTilePixels + idx] = sc * v.z;
            Mb[(kk4 * 4 + 3) * kTilePixels + idx] = sc * v.w;
          }
        } else {
#pragma unroll
          for (int kk = 0; kk < C; ++kk) Mb[kk * kTilePixels + idx] = sc * sm.jc[st][s * C + kk];
        }
#pragma unroll
        for (int j = 0; j < 7; ++j) Mb[(C + j) * kTilePixels + idx] = feat[1 + j];
#pragma unroll
        for (int j = C + 7; j < NFP; ++j) Mb[j * kTilePixels + idx] = 0.0f;
      } else {
        const int idx = nvalid + base_invalid + rank;
        if (idx < padded) {
#pragma unroll
          for (int j = 0; j < NFP; ++j) Mb[j * kTilePixels + idx] = 0.0f;
        }
      }
      if (tid == 0) {
        sm.meta[buf].nvalid = nvalid;
        sm.meta[buf].item_changed = changed ? 1 : 0;
        sm.meta[buf].slot = (int)I.slot;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm.m_full[buf]);  // release: M tile + meta visible to the Gram warps
      named_bar_sync(1, kFeThreads);    // all front-end threads are done with ring stage `st`
      if (tid == 0) {
        const int gn = g + kStages;
        if (gn < g_hi) {
          while ((uint32_t)gn >= items[it_pf].tile_begin + items[it_pf].num_tiles) ++it_pf;
          if (items[it_pf].flags & ITEM_FLAG_BULK) issue_tile_loads<C, kTilePixels>(sm, items, it_pf, gn, st);
        }
      }
    }
  } else if (warp < kFeWarps + NBLK) {
    // ========================================================================= Gram warps
    const int b = warp - kFeWarps;
    // block index -> (bi, bj), bi <= bj, row-major over the upper triangle
    int bi = 0, rem = b;
    while (rem >= NB - bi) {
      rem -= NB - bi;
      ++bi;
    }
    const int bj = bi + rem;
    float acc[64];
#pragma unroll
    for (int e = 0; e < 64; ++e) acc[e] = 0.0f;
    unsigned int inliers = 0;
    int cur_slot = -1;

    auto flush = [&](int slot) {
      reduce_scatter64(acc, lane);
      float* P = partials + (size_t)slot * Cfg::PARTIAL_FLOATS;
      const int e0 = 2 * lane;  // entries e0, e0+1 of the 8x8 block: row e0/8, cols e0%8, e0%8+1
      float2 v = make_float2(acc[0], acc[1]);
      *reinterpret_cast<float2*>(&P[(8 * bi + (e0 >> 3)) * NFP + 8 * bj + (e0 & 7)]) = v;
      if (b == 0 && lane == 0) reinterpret_cast<unsigned int*>(P)[NFP * NFP] = inliers;
#pragma unroll
      for (int e = 0; e < 64; ++e) acc[e] = 0.0f;
      inliers = 0;
    };

    for (int g = g_lo, i = 0; g < g_hi; ++g, ++i) {
      const int buf = i & 1;
      mbar_wait(&sm.m_full[buf], (i >> 1) & 1u);
      const TileMeta meta = sm.meta[buf];
      if (meta.item_changed) {
        if (cur_slot >= 0) flush(cur_slot);
        cur_slot = meta.slot;
      }
      inliers += (unsigned)meta.nvalid;
      // 32-bit shared-window addresses + explicit ld.shared keep the loop at 64 accumulators + 12 operands
      const uint32_t mr = smem_u32(sm.M[buf]) + 4u * ((8 * bi) * kTilePixels + lane);
      const uint32_t mc = smem_u32(sm.M[buf]) + 4u * ((8 * bj) * kTilePixels + lane);
      const int steps = (meta.nvalid + 31) >> 5;
      for (int s = 0; s < steps; ++s) {
        float r[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] = lds_f32(mr + 4u * (j * kTilePixels + s * 32));
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          float c[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) c[k] = (bi == bj) ? r[4 * h + k] : lds_f32(mc + 4u * ((4 * h + k) * kTilePixels + s * 32));
#pragma unroll
          for (int j = 0; j < 8; ++j)
#pragma unroll
            for (int k = 0; k < 4; ++k) acc[j * 8 + 4 * h + k] = fmaf(r[j], c[k], acc[j * 8 + 4 * h + k]);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm.m_empty[buf]);
    }
    if (cur_slot >= 0) flush(cur_slot);
  }
}

template <int C>
cudaError_t launch_impl(const SfmItemDev* items_dev, const SfmLaunchPlan& plan, float* partials_dev,
                        float* records_dev, cudaStream_t stream, cudaEvent_t ev_start, cudaEvent_t ev_stop)
{
  using Cfg = SfmCfg<C>;
  const size_t smem = sizeof(Smem<C>);
  cudaError_t err = cudaFuncSetAttribute(sfm_step_fp32_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem);
  if (err != cudaSuccess) return err;
  const int threads = (kFeWarps + Cfg::NBLK) * 32;
  if (ev_start) cudaEventRecord(ev_start, stream);
  sfm_step_fp32_kernel<C><<<plan.num_ctas, threads, smem, stream>>>(items_dev, plan.num_items, plan.num_tiles,
                                                                    partials_dev);
  if (ev_stop) cudaEventRecord(ev_stop, stream);
  err = cudaGetLastError();
  if (err != cudaSuccess) return err;
  (void)records_dev;
  return cudaSuccess;
}

}  // namespace

bool sfm_fp32_supported(int code_size) { return code_size == 8 || code_size == 16 || code_size == 32; }

size_t sfm_partial_floats(int code_size)
{
  switch (code_size) {
    case 8: return SfmCfg<8>::PARTIAL_FLOATS;
    case 16: return SfmCfg<16>::PARTIAL_FLOATS;
    case 32: return SfmCfg<32>::PARTIAL_FLOATS;
    case 64: return SfmCfg<64>::PARTIAL_FLOATS;
    case 128: return SfmCfg<128>::PARTIAL_FLOATS;
    default: return 0;
  }
}

int sfm_max_ctas()
{
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return sms > 0 ? sms : 1;
}

cudaError_t launch_sfm_fp32(int code_size, const SfmItemDev* items_dev, const SfmLaunchPlan& plan,
                            float* partials_dev, float* records_dev, cudaStream_t stream, cudaEvent_t ev_start,
                            cudaEvent_t ev_stop)
{
  switch (code_size) {
    case 8: return launch_impl<8>(items_dev, plan, partials_dev, records_dev, stream, ev_start, ev_stop);
    case 16: return launch_impl<16>(items_dev, plan, partials_dev, records_dev, stream, ev_start, ev_stop);
    case 32: return launch_impl<32>(items_dev, plan, partials_dev, records_dev, stream, ev_start, ev_stop);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace dfk
