// dfk_sfm_wide.cu -- SfmAligner::RunStep hot path for the wide code sizes (C = 64, 128), fp32 CUDA-core Gram.
//
// Same contract as dfk_sfm_fp32.cu (kernel_step_calculate + DenseSfm + the two-stage reduction of the
// reference: sources/cuda/cu_sfmaligner.cpp:40-70, sources/common/algorithm/dense_sfm.h:133-201,
// sources/cuda/kernel_utils.h:51-69) for the code sizes the reference declares but cannot launch (its
// per-thread 1 x (12+C) item does not fit at C >= 64: cu_sfmaligner.cpp:170-173, instantiations commented
// out at :210-211).
//
// What differs from the C <= 32 kernel: the reduced system G = sum m^T m has (7+C)^2 entries -- 45 / 153
// upper 8x8 blocks at C = 64 / 128 -- so a block is owned by a THREAD (64 register accumulators), not by a
// warp.  KS threads share one block and split the tile's compacted pixels round-robin; their partial sums
// meet in a fixed shuffle order when the CTA leaves an item, so results are bitwise reproducible.
//   * persistent CTAs, static tile ranges, TMA (cp.async.bulk) staging of jac / img0 / dpt0 rows: as in
//     the C <= 32 kernel, with a tile of 128 (C = 64) or 64 (C = 128) pixels so two stages fit.
//   * front-end: one thread per pixel; exact-order validity chain, gathers, reduced row
//       m = w * [ e*jc (C) | a (6) | diff ]   written pixel-major and compacted into M[pixel][NFP].
//   * Gram threads: for every valid pixel of their split, 2+2 LDS.128 and 64 FMA.
// This is the first correct path for these sizes: CUDA-core bound (9.8 kFMA per pixel at C = 128), the
// tcgen05 formulation of dfk_sfm_tc.cu is the follow-up (DESIGN.md "what comes next").
#include <cuda_runtime.h>
#include <stdint.h>

#include "dfk_async.cuh"
#include "dfk_geom.cuh"
#include "dfk_internal.h"
#include "dfk_tile_stage.cuh"

namespace dfk {

namespace {

constexpr int kStages = 2;
constexpr int kMaxCode = 128;  // largest code size of this kernel (fused depth decode keeps the code in shared memory)

template <int C>
struct WideCfg {
  using Cfg = SfmCfg<C>;
  static constexpr int TILE = sfm_wide_tile_pixels(C);
  static constexpr int KS = (C >= 128) ? 2 : 8;  // threads per 8x8 block (power of two, <= 32)
  static constexpr int FE_THREADS = TILE;
  static constexpr int FE_WARPS = TILE / 32;
  static constexpr int GRAM_THREADS = ((Cfg::NBLK * KS + 31) / 32) * 32;
  static constexpr int GRAM_WARPS = GRAM_THREADS / 32;
  static constexpr int THREADS = FE_THREADS + GRAM_THREADS;
};

struct TileMeta {
  int nvalid;
  int item_changed;
  int slot;
  int pad;
};

using ItemSmem = StagedItem<kMaxCode>;  // dfk_tile_stage.cuh

template <int C>
struct Smem {
  using W = WideCfg<C>;
  alignas(128) float jc[kStages][W::TILE * C];
  alignas(16) float M[2][W::TILE * SfmCfg<C>::NFP];
  alignas(16) float img0[kStages][W::TILE];
  alignas(16) float dpt0[kStages][W::TILE];
  alignas(8) uint64_t full_tma[kStages];
  uint64_t m_full[2];
  uint64_t m_empty[2];
  TileMeta meta[2];
  ItemSmem item;
  int cnt[W::FE_WARPS];
};

template <int C>
__global__ void __launch_bounds__(WideCfg<C>::THREADS, 1)
sfm_step_wide_kernel(const SfmItemDev* __restrict__ items, int num_items, int num_tiles, float* __restrict__ partials)
{
  using Cfg = SfmCfg<C>;
  using W = WideCfg<C>;
  constexpr int NFP = Cfg::NFP;
  constexpr int NB = Cfg::NB;
  constexpr int NBLK = Cfg::NBLK;
  constexpr int TILE = W::TILE;
  constexpr int KS = W::KS;
  constexpr int FE = W::FE_THREADS;
  constexpr int NV = C / 4;  // float4 chunks per code-Jacobian row
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
      mbar_init(&sm.m_full[b], W::FE_WARPS);  // one arrival per front-end warp (every arrival wakes the waiters)
      mbar_init(&sm.m_empty[b], W::GRAM_WARPS);
    }
    mbar_fence_init();
  }
  __syncthreads();
  if (g_lo >= g_hi) return;

  if (tid < FE) {
    // ========================================================================= front-end
    int it = 0;
    while (it + 1 < num_items && (uint32_t)g_lo >= items[it].tile_begin + items[it].num_tiles) ++it;
    int it_pf = it;
    uint32_t tma_phase_bits = 0;
    int cur_item = -1;

    if (tid == 0) {
      for (int j = 0; j < kStages && g_lo + j < g_hi; ++j) {
        const int g = g_lo + j;
        while ((uint32_t)g >= items[it_pf].tile_begin + items[it_pf].num_tiles) ++it_pf;
        if (items[it_pf].flags & ITEM_FLAG_BULK) issue_tile_loads<C, WideCfg<C>::TILE>(sm, items, it_pf, g, j);
      }
    }

    for (int g = g_lo, i = 0; g < g_hi; ++g, ++i) {
      const int st = i % kStages;
      const int buf = i & 1;
      while ((uint32_t)g >= items[it].tile_begin + items[it].num_tiles) ++it;
      const bool changed = (it != cur_item);
      if (changed) {
        named_bar_sync(1, FE);
        load_item(sm.item, items[it], tid, W::FE_THREADS, cta);
        load_code(sm.item, items[it], C, tid, FE);
        cur_item = it;
        named_bar_sync(1, FE);
      }
      const ItemSmem& I = sm.item;
      const uint32_t k = (uint32_t)g - I.tile_begin;
      const uint32_t tau = (uint32_t)(((uint64_t)k * I.perm_mul) % I.num_tiles);
      const uint32_t p0 = tau * TILE;
      const uint32_t n = min((uint32_t)TILE, I.num_pixels - p0);
      if (I.flags & ITEM_FLAG_BULK) {
        mbar_wait(&sm.full_tma[st], (tma_phase_bits >> st) & 1u);
        tma_phase_bits ^= (1u << st);
      } else {
        coop_tile_loads<C, WideCfg<C>::FE_THREADS>(sm, I, p0, n, st, tid);
        named_bar_sync(1, FE);
      }

      float feat[8];  // s = w*e, w*a[0..5], w*diff
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
          float part[NV];
#pragma unroll
          for (int k4 = 0; k4 < NV; ++k4) part[k4] = chunk_dot(row[k4], cod[k4]);
          d = prx_to_depth(__fadd_rn(d, butterfly_sum<NV>(part)), I.avg_dpt);
          I.dpt_out[(size_t)y * I.dpt_out_pitch + x] = d;
        }
        const float xn = ray_coord((float)x, I.u0, I.fx);
        const float yn = ray_coord((float)y, I.v0, I.fy);
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

      // ---- compaction: the tile's valid pixels become rows 0 .. nvalid-1 of M ------------------------
      const unsigned bal = __ballot_sync(0xffffffffu, ok);
      const int rank = __popc(bal & ((1u << lane) - 1u));
      if (lane == 0) sm.cnt[warp] = __popc(bal);
      mbar_wait(&sm.m_empty[buf], ((i >> 1) & 1u) ^ 1u);  // M[buf] drained by the Gram threads (tile i-2)
      named_bar_sync(1, FE);
      int nvalid = 0, base_valid = 0;
#pragma unroll
      for (int w2 = 0; w2 < W::FE_WARPS; ++w2) {
        if (w2 == warp) base_valid = nvalid;
        nvalid += sm.cnt[w2];
      }
      if (ok) {
        const int idx = base_valid + rank;
        const float sc = feat[0];
        // rotated float4 order: lane l touches chunk (k4 + l) % NV => reads (row stride 4C bytes) and writes
        // (row stride 4*NFP bytes, NFP % 32 == 8) are both free of bank conflicts
        const float4* src = reinterpret_cast<const float4*>(&sm.jc[st][s * C]);
        float4* dst = reinterpret_cast<float4*>(&sm.M[buf][idx * NFP]);
#pragma unroll 8
        for (int k4 = 0; k4 < NV; ++k4) {
          const int kk4 = (k4 + lane) % NV;
          const float4 v = src[kk4];
          dst[kk4] = make_float4(sc * v.x, sc * v.y, sc * v.z, sc * v.w);
        }
        dst[NV] = make_float4(feat[1], feat[2], feat[3], feat[4]);
        dst[NV + 1] = make_float4(feat[5], feat[6], feat[7], 0.0f);
      }
      if (tid == 0) {
        sm.meta[buf].nvalid = nvalid;
        sm.meta[buf].item_changed = changed ? 1 : 0;
        sm.meta[buf].slot = (int)I.slot;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm.m_full[buf]);
      named_bar_sync(1, FE);  // all front-end threads are done with ring stage `st`
      if (tid == 0) {
        const int gn = g + kStages;
        if (gn < g_hi) {
          while ((uint32_t)gn >= items[it_pf].tile_begin + items[it_pf].num_tiles) ++it_pf;
          if (items[it_pf].flags & ITEM_FLAG_BULK) issue_tile_loads<C, WideCfg<C>::TILE>(sm, items, it_pf, gn, st);
        }
      }
    }
  } else {
    // ========================================================================= Gram threads
    const int gt = tid - FE;
    const int braw = gt / KS;
    const int ks = gt % KS;
    const bool active = braw < NBLK;
    const int b = active ? braw : 0;
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
      // the KS threads of a block sit in adjacent lanes: butterfly in fixed order, split 0 stores
#pragma unroll
      for (int e = 0; e < 64; ++e) {
        float v = acc[e];
#pragma unroll
        for (int m = 1; m < KS; m <<= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
        acc[e] = v;
      }
      float* P = partials + (size_t)slot * Cfg::PARTIAL_FLOATS;
      if (active && ks == 0) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          float4* row = reinterpret_cast<float4*>(&P[(8 * bi + r) * NFP + 8 * bj]);
          row[0] = make_float4(acc[8 * r + 0], acc[8 * r + 1], acc[8 * r + 2], acc[8 * r + 3]);
          row[1] = make_float4(acc[8 * r + 4], acc[8 * r + 5], acc[8 * r + 6], acc[8 * r + 7]);
        }
        if (braw == 0) reinterpret_cast<unsigned int*>(P)[NFP * NFP] = inliers;
      }
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
      const float4* Mr = reinterpret_cast<const float4*>(&sm.M[buf][8 * bi]);
      const float4* Mc = reinterpret_cast<const float4*>(&sm.M[buf][8 * bj]);
#pragma unroll 1
      for (int p = ks; p < meta.nvalid; p += KS) {
        const float4 r0 = Mr[p * (NFP / 4)], r1 = Mr[p * (NFP / 4) + 1];
        const float4 c0 = Mc[p * (NFP / 4)], c1 = Mc[p * (NFP / 4) + 1];
        const float r[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
        const float c[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
#pragma unroll
        for (int j = 0; j < 8; ++j)
#pragma unroll
          for (int k = 0; k < 8; ++k) acc[8 * j + k] = fmaf(r[j], c[k], acc[8 * j + k]);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm.m_empty[buf]);
    }
    if (cur_slot >= 0) flush(cur_slot);
  }
}

template <int C>
cudaError_t launch_impl(const SfmItemDev* items_dev, const SfmLaunchPlan& plan, float* partials_dev,
                        cudaStream_t stream, cudaEvent_t ev_start, cudaEvent_t ev_stop)
{
  const size_t smem = sizeof(Smem<C>);
  cudaError_t err = cudaFuncSetAttribute(sfm_step_wide_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem);
  if (err != cudaSuccess) return err;
  if (ev_start) cudaEventRecord(ev_start, stream);
  sfm_step_wide_kernel<C><<<plan.num_ctas, WideCfg<C>::THREADS, smem, stream>>>(items_dev, plan.num_items,
                                                                               plan.num_tiles, partials_dev);
  if (ev_stop) cudaEventRecord(ev_stop, stream);
  return cudaGetLastError();
}

}  // namespace

bool sfm_wide_supported(int code_size) { return code_size == 64 || code_size == 128; }

cudaError_t launch_sfm_wide(int code_size, const SfmItemDev* items_dev, const SfmLaunchPlan& plan,
                            float* partials_dev, cudaStream_t stream, cudaEvent_t ev_start, cudaEvent_t ev_stop)
{
  switch (code_size) {
    case 64: return launch_impl<64>(items_dev, plan, partials_dev, stream, ev_start, ev_stop);
    case 128: return launch_impl<128>(items_dev, plan, partials_dev, stream, ev_start, ev_stop);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace dfk
