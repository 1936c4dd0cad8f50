// dfk_tile_stage.cuh -- what the fp32 and the wide RunStep kernels share around their tile ring: the per-item parameter
// block a CTA keeps in shared memory, and the two ways a 1-D tile of an item (tile k of the item is processed as
// (k * perm_mul) % num_tiles, dfk_internal.h) gets into a ring stage: row segments through the bulk-copy engine
// (cp.async.bulk, one issuing thread), or a cooperative copy for items whose buffers are not 16-byte friendly.
// The ring itself (jc / img0 / dpt0 stages + full_tma barriers) lives in each kernel's Smem<C>.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "dfk_async.cuh"
#include "dfk_internal.h"

namespace dfk {

// per-item parameters the front-end needs, copied to shared memory when the CTA enters an item
template <int MAXCODE>
struct StagedItem {
  float q[4];
  float t[3];
  float R[9];
  float fx, fy, u0, v0, border, ulim, vlim, min_dpt, avg_dpt, huber_delta;
  const float* img0;
  const float* img1;
  const float* dpt0;
  float* valid0;
  const float* jac;
  const float* grad1;
  uint32_t img0_pitch, img1_pitch, dpt0_pitch, valid0_pitch, jac_pitch, grad1_pitch;
  uint32_t width, height, num_pixels, tile_begin, num_tiles, perm_mul, flags, slot;
  float* dpt_out;  // fused depth decode (ITEM_FLAG_FUSED_DEPTH): decoded depth goes here, dpt0 stages prx_orig
  uint32_t dpt_out_pitch;
  alignas(16) float code[MAXCODE];
};

// field-wise copy by a few of the `nthreads` calling threads (small, once per item); callers sync afterwards
template <int MAXCODE>
__device__ __forceinline__ void load_item(StagedItem<MAXCODE>& dst, const SfmItemDev& src, int tid, int nthreads, int cta)
{
  if (tid < 4) dst.q[tid] = src.q[tid];
  if (tid < 3) dst.t[tid] = src.t[tid];
  if (tid < 9) dst.R[tid] = src.R[tid];
  if (tid == 32 % nthreads) {
    dst.fx = src.fx; dst.fy = src.fy; dst.u0 = src.u0; dst.v0 = src.v0;
    dst.border = src.border; dst.ulim = src.ulim; dst.vlim = src.vlim;
    dst.min_dpt = src.min_dpt; dst.avg_dpt = src.avg_dpt; dst.huber_delta = src.huber_delta;
  }
  if (tid == 64 % nthreads) {
    dst.img0 = src.img0; dst.img1 = src.img1; dst.dpt0 = src.dpt0; dst.valid0 = src.valid0;
    dst.jac = src.jac; dst.grad1 = src.grad1;
    dst.img0_pitch = src.img0_pitch; dst.img1_pitch = src.img1_pitch; dst.dpt0_pitch = src.dpt0_pitch;
    dst.valid0_pitch = src.valid0_pitch; dst.jac_pitch = src.jac_pitch; dst.grad1_pitch = src.grad1_pitch;
  }
  if (tid == 96 % nthreads) {
    dst.width = src.width; dst.height = src.height; dst.num_pixels = src.num_pixels;
    dst.tile_begin = src.tile_begin; dst.num_tiles = src.num_tiles; dst.perm_mul = src.perm_mul;
    dst.flags = src.flags;
    dst.slot = src.partial_begin + (uint32_t)cta - src.first_cta;
    dst.dpt_out = src.dpt_out; dst.dpt_out_pitch = src.dpt_out_pitch;
  }
}

// fused depth decode: the item's latent code -> shared memory (callers sync afterwards)
template <int MAXCODE>
__device__ __forceinline__ void load_code(StagedItem<MAXCODE>& dst, const SfmItemDev& src, int code_size, int tid, int nthreads)
{
  if (src.flags & ITEM_FLAG_FUSED_DEPTH)
    for (int k = tid; k < code_size; k += nthreads) dst.code[k] = __ldg(src.code + k);
}

// Issue the bulk copies of global tile g (item `it`, TILE pixels per tile) into ring stage `st`.  One thread.
template <int C, int TILE, class SmemT>
__device__ __forceinline__ void issue_tile_loads(SmemT& sm, const SfmItemDev* __restrict__ items, int it, int g, int st)
{
  const SfmItemDev& I = items[it];
  const uint32_t k = (uint32_t)g - I.tile_begin;
  const uint32_t tau = (uint32_t)(((uint64_t)k * I.perm_mul) % I.num_tiles);
  const uint32_t p0 = tau * TILE;
  const uint32_t n = min((uint32_t)TILE, I.num_pixels - p0);
  const uint32_t W = I.width;
  uint32_t y = p0 / W;
  uint32_t x = p0 - y * W;
  mbar_arrive_expect_tx(&sm.full_tma[st], n * (C + 2) * 4u);
  uint32_t slot = 0;
  while (slot < n) {
    const uint32_t seg = min(W - x, n - slot);
    bulk_g2s(&sm.jc[st][slot * C], I.jac + (size_t)y * I.jac_pitch + (size_t)x * C, seg * C * 4u, &sm.full_tma[st]);
    bulk_g2s(&sm.img0[st][slot], I.img0 + (size_t)y * I.img0_pitch + x, seg * 4u, &sm.full_tma[st]);
    bulk_g2s(&sm.dpt0[st][slot], I.dpt0 + (size_t)y * I.dpt0_pitch + x, seg * 4u, &sm.full_tma[st]);
    slot += seg;
    x = 0;
    ++y;
  }
}

// cooperative (non-TMA) staging by NT threads for items whose buffers are not 16-byte friendly
template <int C, int NT, class SmemT, class ItemT>
__device__ __forceinline__ void coop_tile_loads(SmemT& sm, const ItemT& I, uint32_t p0, uint32_t n, int st, int tid)
{
  const uint32_t W = I.width;
  for (uint32_t s = tid; s < n; s += NT) {
    const uint32_t p = p0 + s;
    const uint32_t y = p / W, x = p - y * W;
    sm.img0[st][s] = __ldg(I.img0 + (size_t)y * I.img0_pitch + x);
    sm.dpt0[st][s] = __ldg(I.dpt0 + (size_t)y * I.dpt0_pitch + x);
  }
  for (uint32_t e = tid; e < n * C; e += NT) {
    const uint32_t s = e / C, kk = e - s * C;
    const uint32_t p = p0 + s;
    const uint32_t y = p / W, x = p - y * W;
    sm.jc[st][e] = __ldg(I.jac + (size_t)y * I.jac_pitch + (size_t)x * C + kk);
  }
}

}  // namespace dfk
