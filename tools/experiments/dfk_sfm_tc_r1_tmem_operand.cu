// dfk_sfm_tc.cu -- SfmAligner::RunStep hot path, tcgen05 tensor-core Gram variant (sm_100a, C = 32).
//
// Same contract as dfk_sfm_fp32.cu (replaces kernel_step_calculate + DenseSfm + the two-kernel
// reduction of sources/cuda/cu_sfmaligner.cpp:40-70,149-185, dense_sfm.h:133-201), different engine
// for the reduced Gram  G = sum_p m_p^T m_p,  m = w*[ e*jc (32) | a (6) | diff (1) ]  (39 features):
//
//   Split precision ("3xTF32" folded into ONE MMA): every feature value v is split exactly into
//   h = the bits the tensor core keeps (fp32 -> tf32 is a truncation of the low 13 mantissa bits on
//   this hardware, measured by tools/umma_probe.cu) and l = v - h.  With A = [h rows ; l rows] (78 of
//   M = 128 rows) and B = h (39 of N = 48 columns), one tcgen05.mma.kind::tf32 per 8 pixels yields
//   HH = sum h h^T and LH = sum l h^T;  G = HH + LH + LH^T  drops only the l*l terms (~2^-22).
//
//   Per CTA (512 threads, 2 CTAs / SM, 256 TMEM columns each; register budgets by setmaxnreg):
//     warps 0-7   front-end : two groups of 4 warps that alternate tiles; one thread per pixel of a 128-pixel tile:
//                             (optional depth decode,) exact-order validity chain, bilinear gathers, Jacobian row,
//                             Huber.  Each warp owns one 32-pixel block: valid pixels are compacted warp-locally and
//                             the staged code-Jacobian row of a valid pixel is scaled by s = w*e and moved to its rank
//                             IN PLACE in the ring stage; w*a[6], w*diff go to shared memory (feat).
//     warps 8-10, operand   : two groups (A, B) of 3 warps, lane = feature row; group g builds blocks g and g+2 of
//           12-14             every tile.  ow 0: h of the 32 code features (the raw scaled values; the tensor core
//                             truncates), also written K-major to shared memory as B; ow 1: l of the code features;
//                             ow 2: h and l of the 7 pose/residual features.  A goes registers -> TMEM with
//                             tcgen05.st.32x32b.x32 (lane = row, column = pixel); the code rows are read with one
//                             conflict-free LDS per pixel (lane = code dimension).  Group A also drains the chains.
//     warp 11     control   : lane 0 issues the MMAs (A from TMEM, B from shared memory through a K-major no-swizzle
//                             descriptor), the tcgen05.commit arrivals, and allocates TMEM.
//     warp 15     producer  : lane 0 issues the cp.async.bulk copies of a tile as soon as its ring stage is free.
//   The fp32 accumulator in TMEM adds with truncation (measured: ~ -2^-24 relative per k-step), so a chain is cut every
//   kFlushTiles tiles: operand group A pulls the finished chain out of TMEM (tcgen05.ld) and adds it in round-to-nearest
//   fp32 to the CTA's partial in global memory (single writer per address, program order => reproducible).
//   Experiment switches (-DDFK_EXP_NOGEOM / NOCOMPACT / NOOPBUILD / NOMMA / NODRAIN: wrong results, informative
//   times) and phase timers (-DDFK_TC_TIMERS + env DFK_TC_DEBUG=1) are kept for the roofline accounting in DESIGN.md.
//
// Tile staging (cp.async.bulk row segments into a 4-deep ring), the static tile->CTA assignment, the
// in-item tile permutation, the per-CTA partials and the wide deterministic finalize are those of the
// fp32 kernel.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "dfk_async.cuh"
#include "dfk_geom.cuh"
#include "dfk_internal.h"
#include "dfk_tcgen05.cuh"

namespace dfk {

namespace {

constexpr int C = 32;
constexpr int TILE = kTcTilePixels;  // 128
constexpr int FEAT_STRIDE = TILE + 4;  // floats per feature row in shared memory (see Smem::feat)
constexpr int HALF = 64;
constexpr int STAGES = 4;
constexpr int FE_GROUPS = 2;         // front-end group g handles the CTA's tiles i with i % 2 == g
constexpr int FE_THREADS = 128;      // per group: one thread per pixel of a tile
constexpr int OP_THREADS = 256;       // warps 8-10: operand group A, 11: control, 12-14: operand group B, 15: TMA producer
// The operand / control / producer warps take the HIGH warp ids: the sub-core arbiter favours higher warp ids, and
// the short operand pipeline must not queue behind the eight front-end warps.
constexpr int THREADS = OP_THREADS + FE_GROUPS * FE_THREADS;
constexpr int NB = 48;           // MMA N (39 used)
constexpr int MM = 128;          // MMA M (78 used)
#ifndef DFK_FLUSH_TILES
#define DFK_FLUSH_TILES 8
#endif
constexpr int kFlushTiles = DFK_FLUSH_TILES;  // TMEM accumulation chain length (tiles)
constexpr uint32_t TMEM_COLS = 256;
constexpr uint32_t A_COL = 0;    // [0,128): two 64-column halves of A
constexpr uint32_t D_COL = 128;  // [128,176), [176,224): two accumulators
constexpr uint32_t B_SBO = (HALF / 4) * 128;                 // 2048 B between 8-row groups
constexpr uint32_t B_HALF_BYTES = (NB / 8) * B_SBO;           // 12288 B
constexpr int JC_STAGE_FLOATS = (TILE + 1) * C;               // +1: an all-zero row for padded pixels

struct TileMeta {
  int nv[4];         // valid pixels of the four 32-pixel blocks (each block compacted on its own)
  int item_changed;
  int slot;
  int pad[2];
};

struct ItemSmem {
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
  const float* ray_tab;
  float* dpt_out;  // fused depth decode: where the decoded depth goes (dpt0 then stages prx_orig)
  uint32_t img0_pitch, img1_pitch, dpt0_pitch, valid0_pitch, jac_pitch, grad1_pitch, dpt_out_pitch;
  uint32_t width, height, num_pixels, tile_begin, num_tiles, perm_mul, flags, slot, mag_tiles, mag_width;
  alignas(128) float code[C];  // fused depth decode: the latent code of the item (128-byte aligned: chunk addresses are formed by xor)
};

struct Smem {
  alignas(128) float jc[STAGES][JC_STAGE_FLOATS];
  alignas(128) unsigned char B[2][B_HALF_BYTES];
  alignas(16) float img0[STAGES][TILE];
  alignas(16) float dpt0[STAGES][TILE];
  // K-major per-pixel scalars of the compacted pixels: s, wa0..5, wr.  Rows are padded by one float4 so that the pose
  // operand warp, whose lanes read the SAME pixel chunk of 7 different rows, hits 7 different bank groups
  alignas(16) float feat[2][8][FEAT_STRIDE];
  alignas(16) int sid[2][TILE];        // slot (row of the jc stage) of each compacted pixel
  alignas(8) uint64_t tma_full[STAGES];
  uint64_t stage_empty[STAGES];  // the operand warps are done with the ring stage
  uint64_t feat_full[2];
  uint64_t feat_empty[2];
  uint64_t a_full[2];
  uint64_t a_empty[2];
  uint64_t d_full[2];
  uint64_t d_empty[2];
  TileMeta meta[2];
  ItemSmem item[FE_GROUPS];
  uint32_t tmem_base;
};

__device__ __forceinline__ void load_item(ItemSmem& dst, const SfmItemDev& src, int tid, int cta)
{
  if (tid < 4) dst.q[tid] = src.q[tid];
  if (tid < 3) dst.t[tid] = src.t[tid];
  if (tid < 9) dst.R[tid] = src.R[tid];
  if (tid == 32) {
    dst.fx = src.fx; dst.fy = src.fy; dst.u0 = src.u0; dst.v0 = src.v0;
    dst.border = src.border; dst.ulim = src.ulim; dst.vlim = src.vlim;
    dst.min_dpt = src.min_dpt; dst.avg_dpt = src.avg_dpt; dst.huber_delta = src.huber_delta;
  }
  if (tid == 64) {
    dst.img0 = src.img0; dst.img1 = src.img1; dst.dpt0 = src.dpt0; dst.valid0 = src.valid0;
    dst.jac = src.jac; dst.grad1 = src.grad1; dst.ray_tab = src.ray_tab;
    dst.img0_pitch = src.img0_pitch; dst.img1_pitch = src.img1_pitch; dst.dpt0_pitch = src.dpt0_pitch;
    dst.valid0_pitch = src.valid0_pitch; dst.jac_pitch = src.jac_pitch; dst.grad1_pitch = src.grad1_pitch;
    dst.dpt_out = src.dpt_out; dst.dpt_out_pitch = src.dpt_out_pitch;
  }
  if (tid >= 64 && tid < 64 + C && (src.flags & ITEM_FLAG_FUSED_DEPTH)) dst.code[tid - 64] = __ldg(src.code + (tid - 64));
  if (tid == 96) {
    dst.width = src.width; dst.height = src.height; dst.num_pixels = src.num_pixels;
    dst.tile_begin = src.tile_begin; dst.num_tiles = src.num_tiles; dst.perm_mul = src.perm_mul;
    dst.flags = src.flags;
    dst.mag_tiles = src.mag_tiles;
    dst.mag_width = src.mag_width;
    dst.slot = src.partial_begin + (uint32_t)cta - src.first_cta;
  }
}

// with_scalars: also stage img0 / dpt0 (only the fused depth decode reads them from the stage; otherwise the front-end
// threads fetch their own pixel with two coalesced loads long before the tile lands)
__device__ __forceinline__ void issue_tile_loads(Smem& sm, const SfmItemDev* __restrict__ items, int it, int g, int st)
{
  const SfmItemDev& I = items[it];
  const bool with_scalars = (I.flags & ITEM_FLAG_FUSED_DEPTH) != 0;
  const uint32_t k = (uint32_t)g - I.tile_begin;
  const uint32_t tau = (uint32_t)(((uint64_t)k * I.perm_mul) % I.num_tiles);
  const uint32_t p0 = tau * TILE;
  const uint32_t n = min((uint32_t)TILE, I.num_pixels - p0);
  const uint32_t W = I.width;
  uint32_t y = p0 / W;
  uint32_t x = p0 - y * W;
  mbar_arrive_expect_tx(&sm.tma_full[st], n * (C + (with_scalars ? 2 : 0)) * 4u);
  uint32_t slot = 0;
  while (slot < n) {
    const uint32_t seg = min(W - x, n - slot);
    bulk_g2s(&sm.jc[st][slot * C], I.jac + (size_t)y * I.jac_pitch + (size_t)x * C, seg * C * 4u, &sm.tma_full[st]);
    if (with_scalars) {
      bulk_g2s(&sm.img0[st][slot], I.img0 + (size_t)y * I.img0_pitch + x, seg * 4u, &sm.tma_full[st]);
      bulk_g2s(&sm.dpt0[st][slot], I.dpt0 + (size_t)y * I.dpt0_pitch + x, seg * 4u, &sm.tma_full[st]);
    }
    slot += seg;
    x = 0;
    ++y;
  }
}

__device__ __forceinline__ void coop_tile_loads(Smem& sm, const ItemSmem& I, uint32_t p0, uint32_t n, int st, int ft)
{
  const uint32_t W = I.width;
  for (uint32_t s = ft; s < n; s += FE_THREADS) {
    const uint32_t p = p0 + s;
    const uint32_t y = p / W, x = p - y * W;
    sm.img0[st][s] = __ldg(I.img0 + (size_t)y * I.img0_pitch + x);
    sm.dpt0[st][s] = __ldg(I.dpt0 + (size_t)y * I.dpt0_pitch + x);
  }
  for (uint32_t e = ft; e < n * C; e += FE_THREADS) {
    const uint32_t s = e / C, kk = e - s * C;
    const uint32_t p = p0 + s;
    const uint32_t y = p / W, x = p - y * W;
    sm.jc[st][e] = __ldg(I.jac + (size_t)y * I.jac_pitch + (size_t)x * C + kk);
  }
}

// a / b and a % b through the precomputed mag = floor(2^32 / b): multiply-high, one correction step
__device__ __forceinline__ uint32_t div_magic(uint32_t a, uint32_t b, uint32_t mag, uint32_t& rem)
{
  uint32_t q = __umulhi(a, mag);
  uint32_t r = a - q * b;
  if (r >= b) {
    ++q;
    r -= b;
  }
  rem = r;
  return q;
}
__device__ __forceinline__ float tf32_trunc(float v) { return __uint_as_float(__float_as_uint(v) & 0xffffe000u); }

// ---- optional phase timers (clock64 sums per role), enabled with the env var DFK_TC_DEBUG=1 ----------
// Compiled in only with -DDFK_TC_TIMERS (they cost ~40 instructions per tile and role).
__device__ unsigned long long g_dbg[16];
#ifdef DFK_TC_TIMERS
struct Tmr {
  long long t;
  bool on;
  __device__ __forceinline__ void start() { if (on) t = clock64(); }
  __device__ __forceinline__ void lap(unsigned long long& acc) { if (on) { const long long n = clock64(); acc += (unsigned long long)(n - t); t = n; } }
};
#else
struct Tmr {
  long long t;
  bool on;
  __device__ __forceinline__ void start() {}
  __device__ __forceinline__ void lap(unsigned long long&) {}
};
#endif

// chain bookkeeping shared (by construction) between the control thread and the operand warps
struct ChainState {
  int e = -1;              // current chain index
  int tiles_in_chain = 0;
  __device__ __forceinline__ bool starts_chain(int i, int item_changed) const
  {
    return i == 0 || item_changed != 0 || tiles_in_chain == kFlushTiles;
  }
};

__global__ void __launch_bounds__(THREADS, 2)
sfm_step_tc_kernel(const SfmItemDev* __restrict__ items, int num_items, int num_tiles, float* __restrict__ partials,
                   int dbg)
{
  extern __shared__ __align__(128) unsigned char smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const int cta = blockIdx.x;
  const int G = gridDim.x;
  const int g_lo = (int)(((long long)cta * num_tiles) / G);
  const int g_hi = (int)(((long long)(cta + 1) * num_tiles) / G);
  const int ntiles = g_hi - g_lo;

  // ---- one-time setup ---------------------------------------------------------------------------
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&sm.tma_full[s], 1);
      mbar_init(&sm.stage_empty[s], 6);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&sm.feat_full[b], FE_THREADS / 32);  // one arrival per front-end warp: every arrival wakes the waiters
      mbar_init(&sm.feat_empty[b], 7);  // 2 x 3 operand warps + the control thread (it reads meta[b])
      mbar_init(&sm.a_full[b], 6);
      mbar_init(&sm.a_empty[b], 1);
      mbar_init(&sm.d_full[b], 1);
      mbar_init(&sm.d_empty[b], 3);
    }
    mbar_fence_init();
  }
  // zero row of every jc stage and the whole B buffer (rows 39..47 are never written again)
  for (int s = 0; s < STAGES; ++s)
    if (tid < C) sm.jc[s][TILE * C + tid] = 0.0f;
  for (int e = tid; e < (int)(2 * B_HALF_BYTES / 4); e += THREADS) reinterpret_cast<float*>(sm.B)[e] = 0.0f;
  if (warp == 11) {
    tmem_alloc(&sm.tmem_base, TMEM_COLS);
    tmem_relinquish();
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = sm.tmem_base;

  // register budget per role (warpgroup granularity): operand / control warps are lean, the front-end is not
  if (warp < 8) {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 72;");
  } else {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
  }

  if (ntiles > 0) {
    if (warp < 8) {
      // ======================================================================= front-end groups
      const int grp = warp >> 2;                // 0 / 1
      const int ft = tid & (FE_THREADS - 1);    // 0..127 = pixel slot
      const int fwarp = warp & 3;
      const uint32_t bar_id = 1 + grp;
      ItemSmem& I = sm.item[grp];
      int it = 0;
      uint32_t tma_phase_bits = 0;
      int cur_item = -1;
      uint32_t item_lo = 0, item_hi = 0;  // global tile range of the item in shared memory
#ifdef DFK_TC_TIMERS
      Tmr tm{0, dbg != 0 && ft == 0};
#else
      Tmr tm{0, false};
#endif
      unsigned long long t_tma = 0, t_geo = 0, t_fe_wait = 0, t_fe_write = 0, t_issue_fe = 0;
      for (int i = grp; i < ntiles; i += FE_GROUPS) {
        const int g = g_lo + i;
        const int st = i % STAGES;
        const int fb = grp;  // == i & 1
        if ((uint32_t)g >= item_hi || cur_item < 0) {
          while ((uint32_t)g >= items[it].tile_begin + items[it].num_tiles) ++it;
          named_bar_sync(bar_id, FE_THREADS);
          load_item(I, items[it], ft, cta);
          cur_item = it;
          named_bar_sync(bar_id, FE_THREADS);
          item_lo = I.tile_begin;
          item_hi = I.tile_begin + I.num_tiles;
        }
        // the tile sequence enters a new item here (relative to tile i-1, which the other group handles)
        const bool seq_changed = (i == 0) || ((uint32_t)(g - 1) < item_lo);
        const uint32_t k = (uint32_t)g - item_lo;
        uint32_t tau;
        div_magic(k * I.perm_mul, I.num_tiles, I.mag_tiles, tau);  // host guarantees k * perm_mul < 2^32
        const uint32_t p0 = tau * TILE;
        const uint32_t n = min((uint32_t)TILE, I.num_pixels - p0);
        const bool bulk = (I.flags & ITEM_FLAG_BULK) != 0;
        const uint32_t s = ft;
        // tile origin (uniform) by one division, then this thread's pixel by wrap-around
        uint32_t x0;
        const uint32_t y0 = div_magic(p0, I.width, I.mag_width, x0);
        uint32_t pxx = x0 + (s < n ? s : 0u), py = y0;
        while (pxx >= I.width) {
          pxx -= I.width;
          ++py;
        }
        const float xn = __ldg(I.ray_tab + pxx);              // in flight while the tile lands
        const float yn = __ldg(I.ray_tab + I.width + py);
        // The tile's code-Jacobian rows are needed only after the geometry (compaction), so a bulk-staged tile is waited
        // for THERE: this thread's own dpt0 / img0 come straight from global memory (coalesced, issued now).  The fused
        // depth decode reads the rows first thing and keeps the early wait.
        const bool early = !bulk || (I.flags & ITEM_FLAG_FUSED_DEPTH) != 0;
        float d_g = 0.0f, i0_g = 0.0f;
        if (!early && s < n) {
          d_g = __ldg(I.dpt0 + (size_t)py * I.dpt0_pitch + pxx);
          i0_g = __ldg(I.img0 + (size_t)py * I.img0_pitch + pxx);
        }
        tm.start();
        if (bulk) {
          if (early) {
            mbar_wait_parked(&sm.tma_full[st], (tma_phase_bits >> st) & 1u);
            tma_phase_bits ^= (1u << st);
          }
        } else {
          // stage st was last read by the operand warps of tile i-4 (stage_empty / feat_empty completed)
          coop_tile_loads(sm, I, p0, n, st, ft);
          named_bar_sync(bar_id, FE_THREADS);
        }
        tm.lap(t_tma);

        float feat[8];
        bool ok = false;
#ifdef DFK_EXP_NOGEOM
        if (s < n) {
          ok = (s & 3u) != 0u;
          const float d = sm.dpt0[st][s];
#pragma unroll
          for (int j = 0; j < 8; ++j) feat[j] = 0.001f * (float)(j + 1) * d + xn * yn;
          if (ok) I.valid0[(size_t)py * I.valid0_pitch + pxx] = 1.0f;
        }
        if (false) {
          const uint32_t y = py, x = pxx;
#else
        if (s < n) {
          const uint32_t y = py, x = pxx;
#endif
          float d = early ? sm.dpt0[st][s] : d_g;
          if (I.flags & ITEM_FLAG_FUSED_DEPTH) {
            // the stage holds prx_orig: decode the depth from this pixel's code-Jacobian row (same arithmetic as
            // update_depth_kernel: chunk fma chains + xor-butterfly; register j holds chunk j ^ (lane & 7), which the
            // butterfly does not care about), publish it, and carry on with it
            const uint32_t src = smem_u32(&sm.jc[st][s * C]) + ((uint32_t)(lane & 7) << 4);
            const uint32_t cod = smem_u32(I.code) + ((uint32_t)(lane & 7) << 4);
            float part[C / 4];
#pragma unroll
            for (int k4 = 0; k4 < C / 4; ++k4) {
              float4 v, c;
              asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(src ^ ((uint32_t)k4 << 4)));
              asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(c.x), "=f"(c.y), "=f"(c.z), "=f"(c.w) : "r"(cod ^ ((uint32_t)k4 << 4)));
              part[k4] = chunk_dot(v, c);
            }
            d = prx_to_depth(__fadd_rn(d, butterfly_sum<C / 4>(part)), I.avg_dpt);
            I.dpt_out[(size_t)y * I.dpt_out_pitch + x] = d;
          }
          const Warped w = warp_ray(xn, yn, d, I.q, I.t, I.fx, I.fy, I.u0, I.v0, I.border, I.ulim, I.vlim, I.min_dpt);
          if (w.valid) {
            ok = true;
            I.valid0[(size_t)y * I.valid0_pitch + x] = 1.0f;  // dense_sfm.h:161
            int ix, iy;
            float fu, fv, gx, gy;
            bilin_setup(w.u, w.v, ix, iy, fu, fv);
            sample_grad(I.grad1, I.grad1_pitch, true, ix, iy, fu, fv, gx, gy);  // the API guarantees 8-byte rows here
            const float i1 = sample_scalar(I.img1, I.img1_pitch, ix, iy, fu, fv);
            float a[6], c00, c02, c11, c12;
            pose_jacobian_row(w, I.fx, I.fy, gx, gy, a, c00, c02, c11, c12);
            const float e = prx_jacobian(w, I.R, d, I.avg_dpt, gx, gy, c00, c02, c11, c12);
            const float diff = (early ? sm.img0[st][s] : i0_g) - i1;
            const float hw = huber_weight(diff, I.huber_delta);
            feat[0] = hw * e;
#pragma unroll
            for (int j = 0; j < 6; ++j) feat[1 + j] = hw * a[j];
            feat[7] = hw * diff;
          }
        }
        // ---- warp-local compaction: this warp owns the 32-pixel block `fwarp` of the tile -----------
        const unsigned bal = __ballot_sync(0xffffffffu, ok);
        const int rank = __popc(bal & ((1u << lane) - 1u));
        const int nvb = __popc(bal);
        const int padded = (nvb + 7) & ~7;
        // In-place compaction of the staged code-Jacobian rows of the block, scaled by s = w*e on the way:
        // row 32*fwarp + r becomes the r-th VALID pixel's  s * jc[0..31]  (what the operand warps feed to the
        // tensor core).  Rotated float4 order keeps reads (row = slot) and writes (row = rank) free of bank
        // conflicts; all rows are read into registers before any is overwritten (same warp => __syncwarp).
        if (bulk && !early) {  // now the rows are needed
          mbar_wait_parked(&sm.tma_full[st], (tma_phase_bits >> st) & 1u);
          tma_phase_bits ^= (1u << st);
        }
        float4 rowv[C / 4];
#ifdef DFK_EXP_NOCOMPACT
        if (false) {
#else
        if (ok) {
#endif
          // rows are 128-byte aligned: chunk (k4 ^ (lane & 7)) of row s  ==  (row address + (lane & 7) * 16) ^ (k4 * 16)
          const uint32_t src = smem_u32(&sm.jc[st][s * C]) + ((uint32_t)(lane & 7) << 4);
#pragma unroll
          for (int k4 = 0; k4 < C / 4; ++k4) {
            const uint32_t addr = src ^ ((uint32_t)k4 << 4);
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                         : "=f"(rowv[k4].x), "=f"(rowv[k4].y), "=f"(rowv[k4].z), "=f"(rowv[k4].w)
                         : "r"(addr));
          }
        }
        __syncwarp();
        tm.lap(t_geo);
        // feat[fb] of tile i-2 must have been consumed by the operand warps
        mbar_wait_parked(&sm.feat_empty[fb], ((i >> 1) & 1u) ^ 1u);
        tm.lap(t_fe_wait);
        const int blk0 = 32 * fwarp;
        if (ok) {
          const int c = blk0 + rank;
          const float sc = feat[0];
          const uint32_t dst = smem_u32(&sm.jc[st][c * C]) + ((uint32_t)(lane & 7) << 4);
#ifndef DFK_EXP_NOCOMPACT
#pragma unroll
          for (int k4 = 0; k4 < C / 4; ++k4) {
            const float4 v = rowv[k4];
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(dst ^ ((uint32_t)k4 << 4)), "f"(sc * v.x),
                         "f"(sc * v.y), "f"(sc * v.z), "f"(sc * v.w)
                         : "memory");
          }
#else
          (void)sc; (void)dst;
#endif
#pragma unroll
          for (int f = 1; f < 8; ++f) sm.feat[fb][f][c] = feat[f];
        }
        // pad the block's list to a multiple of 8 with "pixels" that contribute exactly zero
        if (lane < padded - nvb) {
          const int c = blk0 + nvb + lane;
          float4* dst = reinterpret_cast<float4*>(&sm.jc[st][c * C]);
#pragma unroll
          for (int k4 = 0; k4 < C / 4; ++k4) dst[k4] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int f = 1; f < 8; ++f) sm.feat[fb][f][c] = 0.0f;
        }
        if (lane == 0) sm.meta[fb].nv[fwarp] = nvb;
        if (ft == 0) {
          sm.meta[fb].item_changed = seq_changed ? 1 : 0;
          sm.meta[fb].slot = (int)I.slot;
        }
        __syncwarp();  // the warp's rows / feat entries / meta are ordered before lane 0's release
        if (lane == 0) mbar_arrive(&sm.feat_full[fb]);
        tm.lap(t_fe_write);
      }
      if (tm.on) {
        atomicAdd(&g_dbg[0], t_tma); atomicAdd(&g_dbg[1], t_geo); atomicAdd(&g_dbg[2], t_fe_wait);
        atomicAdd(&g_dbg[3], t_fe_write); atomicAdd(&g_dbg[13], t_issue_fe);
      }
    } else if (warp == 11) {
      // ======================================================================= control warp
      if (lane == 0) {
        const uint32_t idesc = make_idesc_tf32(MM, NB);
        ChainState ch;
        bool first = true;
#ifdef DFK_TC_TIMERS
        Tmr tm{0, dbg != 0};
#else
        Tmr tm{0, false};
#endif
        unsigned long long t_afull = 0, t_issue = 0;
        for (int i = 0; i < ntiles; ++i) {
          const int fb = i & 1;
          TileMeta meta{};
          for (int h = 0; h < 2; ++h) {
            tm.start();
            mbar_wait_parked(&sm.a_full[h], i & 1u);
            tc_fence_after();
            tm.lap(t_afull);
            if (h == 0) {
              meta = sm.meta[fb];
              mbar_arrive(&sm.feat_empty[fb]);  // meta[fb] may now be overwritten (once the operand warps agree)
              if (ch.starts_chain(i, meta.item_changed)) {
                if (i > 0) umma_commit(&sm.d_full[ch.e & 1]);
                ch.e += 1;
                ch.tiles_in_chain = 0;
                first = true;
                const int use = ch.e >> 1;  // n-th use of this accumulator buffer
                if (use >= 1) {
                  mbar_wait_parked(&sm.d_empty[ch.e & 1], (use - 1) & 1u);
                  tc_fence_after();
                }
              }
              ch.tiles_in_chain += 1;
            }
            const uint32_t d_addr = tbase + D_COL + NB * (ch.e & 1);
            const uint64_t bdesc0 = make_smem_desc_kmajor_noswizzle(smem_u32(sm.B[h]), 128, B_SBO);
#pragma unroll
            for (int bb = 0; bb < 2; ++bb) {  // the two 32-pixel blocks of the half
              const int nk = (meta.nv[2 * h + bb] + 7) >> 3;
              for (int ks = 0; ks < nk; ++ks) {
                const int kc = 4 * bb + ks;  // 8-pixel k-step inside the half
#ifndef DFK_EXP_NOMMA
                umma_tf32_ts(d_addr, tbase + A_COL + HALF * h + 8 * kc, bdesc0 + (uint64_t)((kc * 256) >> 4), idesc, !first);
#endif
                first = false;
              }
            }
            umma_commit(&sm.a_empty[h]);
            tm.lap(t_issue);
          }
        }
        umma_commit(&sm.d_full[ch.e & 1]);
        if (tm.on) { atomicAdd(&g_dbg[4], t_afull); atomicAdd(&g_dbg[5], t_issue); }
      }
    } else if (warp == 15) {
      // ======================================================================= TMA producer (one thread)
      // Tile j is issued as soon as its ring stage is free (tile j-4 consumed), i.e. up to three tiles ahead of
      // the operand warps; issuing bulk copies costs hundreds of cycles apiece, so it lives on its own warp.
      if (lane == 0) {
        int it_pf = 0;
        for (int j = 0; j < ntiles; ++j) {
          const int g = g_lo + j;
          if (j >= STAGES) mbar_wait_parked(&sm.stage_empty[j % STAGES], ((j / STAGES) - 1) & 1u);
          while ((uint32_t)g >= items[it_pf].tile_begin + items[it_pf].num_tiles) ++it_pf;
          if (items[it_pf].flags & ITEM_FLAG_BULK) issue_tile_loads(sm, items, it_pf, g, j % STAGES);
        }
      }
    } else if ((warp & 3) != 3) {
      // ======================================================================= operand warps
      // group A (warps 0-2) builds half 0 of every tile and drains the accumulators; group B (warps 4-6)
      // builds half 1.  ow: 0 = code-h (+B), 1 = code-l, 2 = pose/residual h+l.
      const int ogrp = (warp - 8) >> 2;
      const int ow = warp & 3;
      const uint32_t lane_taddr = tbase + ((uint32_t)(ow * 32) << 16);
      const int row = ow * 32 + lane;  // TMEM lane == row of the partial
      ChainState ch;
      int chain_valid = 0;        // valid pixels accumulated into the current chain
      int cur_slot = -1;
      bool slot_fresh = true;     // the current item's partial has not been written yet by this CTA
      unsigned int inliers = 0;   // of the current item (warp 0 reports)
      // deferred drain of a finished chain
      bool pend = false;
      int pend_e = 0, pend_valid = 0, pend_slot = 0;
      bool pend_fresh = false, pend_item_end = false;
      unsigned int pend_inliers = 0;

      // Move a finished chain TMEM -> the CTA's partial in global memory (single writer, fixed order).
      // fresh: first chain of the item in this CTA (store), else read-modify-write in round-to-nearest fp32.
      auto drain = [&](int e, int valid, int slot, bool fresh, bool item_end, unsigned int inl) {
        const int b = e & 1, use = e >> 1;
        float* P = partials + (size_t)slot * kTcPartialFloats;
        mbar_wait(&sm.d_full[b], use & 1u);
        tc_fence_after();
        // three passes of 16 accumulator columns keep the register footprint small.  The first chain of an
        // item in this CTA stores, later chains add with fire-and-forget red.global.add.f32: this thread is the
        // only writer of its row and issues its updates in program order, so the sum order is fixed.
#ifdef DFK_EXP_NODRAIN
        if (false) {
#else
        if (valid > 0 || fresh) {
#endif
#pragma unroll 1
          for (int pass = 0; pass < 3; ++pass) {
            const int nq = pass < 2 ? 4 : (kTcCols - 32) / 4;  // float4 per pass (columns 40..47 are padding)
            uint32_t v[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = 0u;
            if (valid > 0) {
              tmem_ld_x16(lane_taddr + D_COL + NB * b + 16 * pass, v);
              tmem_wait_ld();
            }
            // column-major partial: this lane's row at column j is P[j * kTcRowsPad + row] -> a warp writes 128
            // contiguous bytes per column
            float* dcol = P + (16 * pass) * kTcRowsPad + row;
            if (fresh) {
#pragma unroll
              for (int j = 0; j < 16; ++j)
                if (j < 4 * nq) __stcg(dcol + j * kTcRowsPad, __uint_as_float(v[j]));
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j)
                if (j < 4 * nq)
                  asm volatile("red.global.add.f32 [%0], %1;" ::"l"(dcol + j * kTcRowsPad), "f"(__uint_as_float(v[j])) : "memory");
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.d_empty[b]);
        if (item_end && ow == 0 && lane == 0) reinterpret_cast<unsigned int*>(P)[kTcRowsPad * kTcCols] = inl;
      };

#ifdef DFK_TC_TIMERS
      Tmr tm{0, dbg != 0 && warp == 8 && lane == 0};
#else
      Tmr tm{0, false};
#endif
      const bool is_a = (ogrp == 0);
      unsigned long long t_ffull = 0, t_aempty = 0, t_build = 0, t_sync = 0, t_drain = 0, t_total = 0, t_misc = 0;
#ifdef DFK_TC_TIMERS
      const long long t_begin = tm.on ? clock64() : 0;
#else
      const long long t_begin = 0;
#endif
      for (int i = 0; i < ntiles; ++i) {
        const int st = i % STAGES;
        const int fb = i & 1;
        tm.start();
        mbar_wait(&sm.feat_full[fb], (i >> 1) & 1u);
        tm.lap(t_ffull);
        const TileMeta meta = sm.meta[fb];
        const int tile_valid = meta.nv[0] + meta.nv[1] + meta.nv[2] + meta.nv[3];
        if (is_a && ch.starts_chain(i, meta.item_changed)) {
          if (i > 0) {
            pend = true;
            pend_e = ch.e;
            pend_valid = chain_valid;
            pend_item_end = meta.item_changed != 0;
            pend_slot = cur_slot;
            pend_fresh = slot_fresh;
            pend_inliers = inliers;
            slot_fresh = false;
          }
          ch.e += 1;
          ch.tiles_in_chain = 0;
          chain_valid = 0;
          if (meta.item_changed) {
            cur_slot = meta.slot;
            slot_fresh = true;
            inliers = 0;
          }
        }
        ch.tiles_in_chain += 1;
        chain_valid += tile_valid;
        inliers += (unsigned)tile_valid;

        // plain (non-volatile) shared-memory accesses: the compiler is free to overlap the loads of
        // several chunks; the mbarrier waits / fences around the loops carry the "memory" clobbers
        tm.lap(t_misc);
        const float* __restrict__ vrow = sm.jc[st] + lane;  // compacted, pre-scaled rows: vrow[c * C]
        const float4* __restrict__ featp = reinterpret_cast<const float4*>(sm.feat[fb]);
#pragma unroll 1
        for (int h = 0; h < 2; ++h) {
          const int nv = sm.meta[fb].nv[2 * h + ogrp];  // this group's block of the half (shared memory: no local-array indexing)
          // A/B half h was last read by the MMAs of tile i-1
          tm.start();
          mbar_wait(&sm.a_empty[h], (i & 1u) ^ 1u);
          tc_fence_after();
          tm.lap(t_aempty);
          unsigned char* bh = sm.B[h];
          // each half = two 32-pixel blocks; operand group g builds block g of the half: 32 row loads, one
          // 32-column tcgen05.st (registers -> TMEM lanes), and for the h rows the K-major B tile
#ifdef DFK_EXP_NOOPBUILD
          if (false) {
#else
          if (nv > 0) {
#endif
            const int c0 = HALF * h + 32 * ogrp;  // first compacted pixel of the block
            const uint32_t a_taddr = lane_taddr + A_COL + c0;
            uint32_t v[32];
            if (ow < 2) {
              const float* src = vrow + c0 * C;
              float val[32];
#pragma unroll
              for (int j = 0; j < 32; ++j) val[j] = src[j * C];
              if (ow == 0) {
                // B rows = features (this lane), 8 k-chunks of 4 pixels, 128 B apart
                float4* brow = reinterpret_cast<float4*>(bh + (uint32_t)(lane >> 3) * B_SBO + (uint32_t)(lane & 7) * 16u) +
                               8 * (8 * ogrp);
#pragma unroll
                for (int q = 0; q < 8; ++q)
                  brow[8 * q] = make_float4(val[4 * q], val[4 * q + 1], val[4 * q + 2], val[4 * q + 3]);
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(val[j]);
              } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(val[j] - tf32_trunc(val[j]));
              }
            } else {
              // pose / residual features: lanes 0-6 = h of feature 1+lane, lanes 7-13 = l of feature 1+(lane-7)
              const int f = 1 + (lane < 7 ? lane : (lane < 14 ? lane - 7 : 0));
              const float4* fp = featp + f * (FEAT_STRIDE / 4) + (c0 >> 2);
              float4 x[8];
#pragma unroll
              for (int q = 0; q < 8; ++q) x[q] = fp[q];
              if (lane < 7) {
                const uint32_t brow_i = 32u + (uint32_t)lane;
                float4* brow = reinterpret_cast<float4*>(bh + (brow_i >> 3) * B_SBO + (brow_i & 7u) * 16u) + 8 * (8 * ogrp);
#pragma unroll
                for (int q = 0; q < 8; ++q) brow[8 * q] = x[q];
              }
#pragma unroll
              for (int q = 0; q < 8; ++q) {
                float e0 = x[q].x, e1 = x[q].y, e2 = x[q].z, e3 = x[q].w;
                if (lane >= 7) {
                  e0 -= tf32_trunc(e0); e1 -= tf32_trunc(e1); e2 -= tf32_trunc(e2); e3 -= tf32_trunc(e3);
                }
                if (lane >= 14) { e0 = 0.f; e1 = 0.f; e2 = 0.f; e3 = 0.f; }
                v[4 * q] = __float_as_uint(e0); v[4 * q + 1] = __float_as_uint(e1);
                v[4 * q + 2] = __float_as_uint(e2); v[4 * q + 3] = __float_as_uint(e3);
              }
            }
            tmem_st_x32(a_taddr, v);
          }
          tm.lap(t_build);
          tmem_wait_st();
          if (ow != 1) fence_proxy_async_smem();  // the code-l warp wrote TMEM only, no B rows
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&sm.a_full[h]);
          tm.lap(t_sync);
        }
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&sm.feat_empty[fb]);
          mbar_arrive(&sm.stage_empty[st]);
        }
        tm.start();

        if (is_a && pend) {  // the chain that ended before this tile: its MMAs completed long ago
          drain(pend_e, pend_valid, pend_slot, pend_fresh, pend_item_end, pend_inliers);
          pend = false;
        }
        tm.lap(t_drain);
      }
      if (is_a) drain(ch.e, chain_valid, cur_slot, slot_fresh, true, inliers);
      if (tm.on) {
        t_total = (unsigned long long)(clock64() - t_begin);
        (void)t_begin;
        atomicAdd(&g_dbg[6], t_ffull); atomicAdd(&g_dbg[7], t_aempty); atomicAdd(&g_dbg[8], t_build);
        atomicAdd(&g_dbg[9], t_sync); atomicAdd(&g_dbg[10], t_drain); atomicAdd(&g_dbg[11], t_total);
        atomicAdd(&g_dbg[12], (unsigned long long)ntiles); atomicAdd(&g_dbg[14], t_misc);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 11) tmem_dealloc(tbase, TMEM_COLS);
}

}  // namespace

// normalised ray tables of an item: xn[x] = (x - u0)/fx for x < W, then yn[y] = (y - v0)/fy for y < H
__global__ void sfm_ray_tables_kernel(const SfmItemDev* __restrict__ items, float* __restrict__ tabs)
{
  (void)tabs;  // every item carries the address of its own table inside the scratch buffer
  const SfmItemDev& I = items[blockIdx.x];
  float* dst = const_cast<float*>(I.ray_tab);
  for (uint32_t x = threadIdx.x; x < I.width; x += blockDim.x) dst[x] = ray_coord((float)x, I.u0, I.fx);
  for (uint32_t y = threadIdx.x; y < I.height; y += blockDim.x) dst[I.width + y] = ray_coord((float)y, I.v0, I.fy);
}

bool sfm_tc_supported(int code_size) { return code_size == 32; }

size_t sfm_tc_smem_bytes() { return sizeof(Smem); }

cudaError_t launch_sfm_tc(const SfmItemDev* items_dev, const SfmLaunchPlan& plan, bool build_ray_tables,
                          float* partials_dev, cudaStream_t stream, cudaEvent_t ev_start, cudaEvent_t ev_stop)
{
  const size_t smem = sizeof(Smem);
  static const cudaError_t attr_err =  // once per process, not once per launch
      cudaFuncSetAttribute(sfm_step_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem));
  cudaError_t err = attr_err;
  if (err != cudaSuccess) return err;
  if (build_ray_tables) {  // only when the work list names a camera level the handle has no table for yet
    sfm_ray_tables_kernel<<<plan.num_items, 256, 0, stream>>>(items_dev, nullptr);
    err = cudaGetLastError();
    if (err != cudaSuccess) return err;
  }
  static const int dbg = []() { const char* e = getenv("DFK_TC_DEBUG"); return (e && e[0] == '1') ? 1 : 0; }();
  if (dbg) {
    unsigned long long z[16] = {0};
    cudaMemcpyToSymbolAsync(g_dbg, z, sizeof(z), 0, cudaMemcpyHostToDevice, stream);
  }
  if (ev_start) cudaEventRecord(ev_start, stream);
  sfm_step_tc_kernel<<<plan.num_ctas, THREADS, smem, stream>>>(items_dev, plan.num_items, plan.num_tiles, partials_dev,
                                                              dbg);
  if (ev_stop) cudaEventRecord(ev_stop, stream);
  if (dbg) {
    unsigned long long v[16];
    cudaStreamSynchronize(stream);
    cudaMemcpyFromSymbol(v, g_dbg, sizeof(v));
    const double nt = v[12] ? (double)v[12] : 1.0;
    fprintf(stderr,
            "[dfk tc dbg] ctas=%d tiles=%d | per tile cycles: FE(g0+g1 thread0) tma_wait %.0f geom %.0f feat_empty_wait %.0f "
            "write %.0f tma_issue %.0f | CTRL a_full_wait %.0f issue %.0f | OP feat_full_wait %.0f a_empty_wait %.0f build %.0f sync %.0f "
            "drain %.0f misc %.0f total %.0f\n",
            plan.num_ctas, plan.num_tiles, v[0] / nt, v[1] / nt, v[2] / nt, v[3] / nt, v[13] / nt, v[4] / nt, v[5] / nt, v[6] / nt,
            v[7] / nt, v[8] / nt, v[9] / nt, v[10] / nt, v[14] / nt, v[11] / nt);
  }
  return cudaGetLastError();
}

}  // namespace dfk
