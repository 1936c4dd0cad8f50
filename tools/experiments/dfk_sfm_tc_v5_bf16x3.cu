// ARCHIVED EXPERIMENT (round 2) -- not built.  Three-term bf16 split ([h;m;l] x [h;m], kind::f16, K = 16, MN-major
// SWIZZLE_128B operands) of the direct-operand kernel.  Parity-green (max rel. error of JtJ vs fp64 2.0e-6, inliers exact,
// 30/30 sfm GPU tests) but SLOWER than the production tf32 [h;l] x h version: 0.212 ms vs 0.148 ms on the pair8 batch.
// ncu (profiles/README.md, round 2): +22 % executed instructions (the 3-way split costs ~5.5 instructions per value
// against 2), the front-end no longer fits its 64-register cap (ptxas ignores setmaxnreg when allocating: 56-208 bytes
// of spills = 2.4 M local-memory requests per launch, which eat the shared-memory wavefronts the smaller operands
// save), one front-end warp less (the 120 accumulator rows need a fourth drain warp), and 2.7 x larger partials
// (128 x 80 accumulator).  With the MMAs removed it still takes 0.207 ms, i.e. the loss is in the front-end.
// It needs kTcRows = 120, kTcCols = 80, kTcRowsPad = 128 in dfk_internal.h and the 8-term finalize
// G = hh + (hm + mh) + mm + (lh + lh^T) + (lm + lm^T) with the feature-slot maps given in the header comment below.
// dfk_sfm_tc.cu -- SfmAligner::RunStep hot path, tcgen05 tensor-core Gram variant (sm_100a, C = 32).
//
// Same contract as dfk_sfm_fp32.cu (replaces kernel_step_calculate + DenseSfm + the two-kernel
// reduction of sources/cuda/cu_sfmaligner.cpp:40-70,149-185, dense_sfm.h:133-201), different engine
// for the reduced Gram  G = sum_p m_p^T m_p,  m = w*[ e*jc (32) | a (6) | diff (1) ]  (39 features):
//
//   Split precision, three bf16 terms folded into ONE MMA: every feature value v is split exactly into
//   h = bf16(v), m = bf16(v - h), l = bf16(v - h - m)  (8 + 8 + 8 mantissa bits; the two subtractions are exact in fp32).
//   With A = [h ; m ; l] (120 of M = 128 rows) and B = [h ; m] (80 columns), one tcgen05.mma.kind::f16 per 16 pixels
//   yields hh, hm, mh, mm, lh, lm in fp32;  G = hh + hm + mh + mm + (lh + lh^T) + (lm + lm^T)  drops only the m*l and
//   l*l terms (~2^-24).  Against the round-1 tf32 split ([h;l] x h, 4-byte operands, K = 8) this halves the number of
//   MMAs and cuts the operand bytes the tensor core fetches from shared memory from 22 KB to 13 KB per 32 pixels, and
//   the operand stores from 10 KB to 8 KB -- the kernel is bound by the SM's shared-memory / L1 data path, which the
//   operand fetch shares with the front-end's loads and stores (tools/umma_ss_bw_probe.cu).
//
//   Data path ("direct operands"): both MMA operands are read from shared memory in the PIXEL-major ("MN-major")
//   layout a per-pixel front-end writes naturally -- canonical SWIZZLE_128B, atoms of 64 features x 8 pixels (1 KB): pixel
//   row r = 128 contiguous bytes whose 16-byte chunk c is stored at chunk position c ^ r (tools/umma_probe_bf16mn.cu).
//   There is no operand-building stage, no TMEM A operand, no transposition and no tile staging: a front-end warp goes
//   from global memory to a finished operand block on its own and hands it to the MMA issuer with one mbarrier arrival.
//
//   Per CTA (512 threads, 2 CTAs / SM, 256 TMEM columns each; register budgets by setmaxnreg):
//     warps 0-10  front-end : warp w owns the CTA's 32-pixel blocks j = w, w+11, ... (a 128-pixel tile = 4 blocks).
//                             One thread per pixel: (optional depth decode,) exact-order validity chain, bilinear
//                             gathers, Jacobian row, Huber -> s = w*e, w*a[6], w*diff.  Then the block's code-Jacobian
//                             rows are read straight from global memory, COALESCED (8 lanes = the eight 16-byte chunks of
//                             one pixel, 4 pixels per instruction; the rows were prefetched into L2 when the block
//                             started), scaled by the pixel's s (one shuffle), split into h / m / l and stored into the
//                             operand slot; each thread adds its own pixel's 7 pose/residual values.  Invalid pixels
//                             contribute exact zeros; 16-pixel groups without a valid pixel are skipped (no loads, no MMA).
//     warp 11     control   : walks the blocks in order: waits for the slot, issues one MMA per non-empty 16-pixel group
//                             (A = the slot's two atoms: M = 128; B = atom 0 and the first 16 features of atom 1: N = 80),
//                             commits the slot back to its next user, cuts the chains and publishes their records to the
//                             drain warps.  It also allocates TMEM.
//     warps 12-15 drain     : pull a finished accumulation chain out of TMEM (tcgen05.ld) and add it in round-to-nearest
//                             fp32 to the CTA's partial in global memory (single writer per address, program order).
//   Operand slot (8 KB): MN atom a at a * 4096; inside, K atom q (pixels 8q..8q+7) at q * 1024, pixel row r at r * 128,
//   16-byte chunk c at (c ^ r) * 16.  Feature slots (2 bytes each):
//     atom 0:  chunk c = [ code-h 4c..4c+3 | code-m 4c..4c+3 ]   (one 16-byte store per thread and 4 features)
//     atom 1:  [ pose-h 0-7 | pose-m 8-15 | code-l 16-47 | pose-l 48-55 | zero 56-63 ]
//   Accumulator rows (TMEM lanes) = slots of atom 0 then atom 1; columns = slots of atom 0 then slots 0-15 of atom 1.
//   The fp32 accumulator in TMEM adds with truncation (measured: ~ -2^-24 relative per k-step), so a chain is cut every
//   kFlushTiles tiles.
//
// The static tile->CTA assignment, the in-item tile permutation, the per-CTA partials and the wide deterministic
// finalize are those of the fp32 kernel.  Earlier data paths (round 1: TMA-staged tiles + operand warps transposing into
// a TMEM A operand; the tf32 [h;l] x h version of this kernel) are kept under tools/experiments/ with their measurements.
#include <cuda_bf16.h>
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
constexpr int NFE = 11;              // front-end warps
constexpr int THREADS = 512;         // 11 front-end warps, control, 4 drain warps
constexpr int CTRL_WARP = 11;
#ifndef DFK_TC_SLOTS
#define DFK_TC_SLOTS 9
#endif
constexpr int NSLOT = DFK_TC_SLOTS;  // operand slots; block j uses slot j % NSLOT
constexpr uint32_t ATOM_BYTES = 1024;             // 64 features x 8 pixels, bf16
constexpr uint32_t MN_STRIDE = 4 * ATOM_BYTES;    // the 4 K atoms of one MN atom are contiguous
constexpr uint32_t SLOT_BYTES = 2 * MN_STRIDE;    // 8192
constexpr int NB = kTcCols;      // MMA N = 80
constexpr int MM = 128;          // MMA M (120 used)
#ifndef DFK_TC_LOAD_BATCH
#define DFK_TC_LOAD_BATCH 2
#endif
#ifndef DFK_FLUSH_TILES
#define DFK_FLUSH_TILES 8
#endif
constexpr int kFlushTiles = DFK_FLUSH_TILES;  // TMEM accumulation chain length (tiles)
constexpr uint32_t TMEM_COLS = 256;
constexpr uint32_t D_COL = 0;  // [0,80), [80,160): two accumulators

struct SlotMeta {
  int mask;          // bit g: the 8-pixel group g of the block has a valid pixel
  int nvalid;
  int item_changed;  // the block opens a tile that is the first of its item in this CTA's sequence
  int pslot;         // partial slot of the item
};

struct ChainRec {
  int valid;         // valid pixels accumulated into the chain
  int pslot;
  int fresh;         // first chain of the item in this CTA: store, do not add
  int item_end;
  unsigned int inliers;
  int last;
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
  float* dpt_out;  // fused depth decode: where the decoded depth goes (dpt0 then is prx_orig)
  uint32_t img0_pitch, img1_pitch, dpt0_pitch, valid0_pitch, jac_pitch, grad1_pitch, dpt_out_pitch;
  uint32_t width, height, num_pixels, tile_begin, num_tiles, perm_mul, flags, slot, mag_tiles, mag_width;
  alignas(16) float code[C];  // fused depth decode: the latent code of the item
};

struct Smem {
  alignas(1024) unsigned char op[NSLOT][SLOT_BYTES];
  alignas(128) ItemSmem item[NFE];
  alignas(8) uint64_t full[NSLOT];
  // done[w]: the MMAs of the block that used front-end warp w's NEXT slot before it have completed.  One barrier per
  // waiting warp, not per slot: a parity wait can only tell a phase from its neighbours, and with several independent
  // producers per slot a per-slot barrier can be two phases away from the one a producer needs
  uint64_t done[NFE];
  uint64_t d_full[2];
  uint64_t d_empty[2];
  alignas(16) SlotMeta meta[NSLOT];  // read with one 16-byte load
  alignas(16) ChainRec chain[2];
  uint32_t tmem_base;
};

// one warp copies an item description to its private shared-memory copy
__device__ __forceinline__ void load_item(ItemSmem& dst, const SfmItemDev& src, int lane, int cta)
{
  if (lane < 4) dst.q[lane] = src.q[lane];
  if (lane >= 4 && lane < 7) dst.t[lane - 4] = src.t[lane - 4];
  if (lane >= 8 && lane < 17) dst.R[lane - 8] = src.R[lane - 8];
  if (lane == 17) {
    dst.fx = src.fx; dst.fy = src.fy; dst.u0 = src.u0; dst.v0 = src.v0;
    dst.border = src.border; dst.ulim = src.ulim; dst.vlim = src.vlim;
    dst.min_dpt = src.min_dpt; dst.avg_dpt = src.avg_dpt; dst.huber_delta = src.huber_delta;
  }
  if (lane == 18) {
    dst.img0 = src.img0; dst.img1 = src.img1; dst.dpt0 = src.dpt0; dst.valid0 = src.valid0;
    dst.jac = src.jac; dst.grad1 = src.grad1; dst.ray_tab = src.ray_tab;
    dst.dpt_out = src.dpt_out; dst.dpt_out_pitch = src.dpt_out_pitch;
  }
  if (lane == 19) {
    dst.img0_pitch = src.img0_pitch; dst.img1_pitch = src.img1_pitch; dst.dpt0_pitch = src.dpt0_pitch;
    dst.valid0_pitch = src.valid0_pitch; dst.jac_pitch = src.jac_pitch; dst.grad1_pitch = src.grad1_pitch;
  }
  if (lane == 20) {
    dst.width = src.width; dst.height = src.height; dst.num_pixels = src.num_pixels;
    dst.tile_begin = src.tile_begin; dst.num_tiles = src.num_tiles; dst.perm_mul = src.perm_mul;
    dst.flags = src.flags;
    dst.mag_tiles = src.mag_tiles;
    dst.mag_width = src.mag_width;
    dst.slot = src.partial_begin + (uint32_t)cta - src.first_cta;
  }
  if (src.flags & ITEM_FLAG_FUSED_DEPTH) dst.code[lane] = __ldg(src.code + lane);
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
// v0, v1 -> packed bf16 pair (v0 in the low half = lower address), round to nearest
__device__ __forceinline__ uint32_t pack_bf16(float v0, float v1)
{
  const __nv_bfloat162 p = __floats2bfloat162_rn(v0, v1);
  return *reinterpret_cast<const uint32_t*>(&p);
}
__device__ __forceinline__ float bf16_lo(uint32_t p) { return __uint_as_float(p << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t p) { return __uint_as_float(p & 0xffff0000u); }
// exact three-way split of two values: h = bf16(v), m = bf16(v - h), l = bf16(v - h - m), each as a packed pair
__device__ __forceinline__ void split3(float v0, float v1, uint32_t& h, uint32_t& m, uint32_t& l)
{
  h = pack_bf16(v0, v1);
  const float r0 = v0 - bf16_lo(h), r1 = v1 - bf16_hi(h);
  m = pack_bf16(r0, r1);
  l = pack_bf16(r0 - bf16_lo(m), r1 - bf16_hi(m));
}
__device__ __forceinline__ void sts128u(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d)
{
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d));
}
__device__ __forceinline__ void sts64u(uint32_t addr, uint32_t a, uint32_t b)
{
  asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(a), "r"(b));
}

// -DDFK_TC_WATCHDOG: a wait that gives up after ~1 s, says who was waiting for what, and traps (bring-up aid: a
// protocol error becomes a CUDA error with a message instead of a hung device)
#ifdef DFK_TC_WATCHDOG
__device__ __noinline__ void wd_wait(uint64_t* bar, uint32_t parity, int what, int idx)
{
  for (long long spin = 0; spin < 2000000; ++spin)
    if (mbar_try_wait(bar, parity)) return;
  if ((threadIdx.x & 31) == 0)
    printf("[dfk tc watchdog] cta %d warp %d stuck: wait %d index %d parity %u\n", (int)blockIdx.x, (int)(threadIdx.x >> 5), what,
           idx, parity);
  __trap();
}
#define TC_WAIT(bar, parity, what, idx) wd_wait(bar, parity, what, idx)
#else
#define TC_WAIT(bar, parity, what, idx) mbar_wait(bar, parity)
#endif


// 16 bytes of a code-Jacobian row; rows of items without the BULK flag are only 4-byte aligned
__device__ __forceinline__ float4 load_chunk(const float* __restrict__ p, bool aligned16)
{
  if (aligned16) return __ldg(reinterpret_cast<const float4*>(p));
  return make_float4(__ldg(p), __ldg(p + 1), __ldg(p + 2), __ldg(p + 3));
}

__global__ void __launch_bounds__(THREADS, 2)
sfm_step_tc_kernel(const SfmItemDev* __restrict__ items, int num_items, int num_tiles, float* __restrict__ partials)
{
  extern __shared__ unsigned char smem_raw[];
  // the swizzle pattern of the operand atoms is a function of the shared-memory address bits: 1024-byte aligned ring
  Smem& sm = *reinterpret_cast<Smem*>(smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));
  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const int cta = blockIdx.x;
  const int G = gridDim.x;
  const int g_lo = (int)(((long long)cta * num_tiles) / G);
  const int g_hi = (int)(((long long)(cta + 1) * num_tiles) / G);
  const int ntiles = g_hi - g_lo;
  const int nblocks = 4 * ntiles;
  (void)num_items;

  // ---- one-time setup ---------------------------------------------------------------------------
  if (tid == 0) {
    for (int s = 0; s < NSLOT; ++s) {
      mbar_init(&sm.full[s], 1);
    }
    for (int w = 0; w < NFE; ++w) mbar_init(&sm.done[w], 1);
    for (int b = 0; b < 2; ++b) {
      mbar_init(&sm.d_full[b], 2);   // the tcgen05.commit of the chain's MMAs + the control thread's release of the record
      mbar_init(&sm.d_empty[b], 4);
    }
    mbar_fence_init();
  }
  // the never-written part of the ring (feature slots 56-63 of atom 1) must hold zeros
  for (int e = tid; e < (int)(NSLOT * SLOT_BYTES / 16); e += THREADS)
    reinterpret_cast<float4*>(&sm.op[0][0])[e] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (warp == CTRL_WARP) {
    tmem_alloc(&sm.tmem_base, TMEM_COLS);
    tmem_relinquish();
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = sm.tmem_base;

  // register budget per role (warpgroup granularity): 12 x 32 x 72 + 4 x 32 x 40 = 32768 = half the register file
  if (warp <= CTRL_WARP) {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 72;");
  } else {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
  }

  if (ntiles > 0) {
    if (warp < NFE) {
      // ======================================================================= front-end warps
      ItemSmem& I = sm.item[warp];
      int it = -1;  // the item in shared memory
      uint32_t done_phase = 0;
      for (int j = warp; j < nblocks; j += NFE) {
        const int i = j >> 2, b = j & 3;
        const int g = g_lo + i;
        if (it < 0 || (uint32_t)g >= I.tile_begin + I.num_tiles) {
          if (it < 0) it = 0;
          while ((uint32_t)g >= items[it].tile_begin + items[it].num_tiles) ++it;
          __syncwarp();
          load_item(I, items[it], lane, cta);
          __syncwarp();
        }
        const uint32_t item_lo = I.tile_begin;
        // the tile sequence enters a new item with this block
        const bool seq_changed = (b == 0) && ((i == 0) || ((uint32_t)(g - 1) < item_lo));
        const uint32_t k = (uint32_t)g - item_lo;
        uint32_t tau;
        div_magic(k * I.perm_mul, I.num_tiles, I.mag_tiles, tau);  // host guarantees k * perm_mul < 2^32
        const uint32_t p0 = tau * TILE;
        const uint32_t n = min((uint32_t)TILE, I.num_pixels - p0);
        const uint32_t s = 32u * (uint32_t)b + (uint32_t)lane;  // pixel slot in the tile
        const bool inb = s < n;
        const bool blk_live = 32u * (uint32_t)b < n;  // else: a block past the end of the item's last tile
        const uint32_t W = I.width;
        const bool a16 = (I.flags & ITEM_FLAG_BULK) != 0;
        const bool fused = (I.flags & ITEM_FLAG_FUSED_DEPTH) != 0;
        // block origin (uniform) by one division, then this thread's pixel by wrap-around; lanes past the end of the
        // tile shadow the block's first pixel (their loads stay in bounds, their contribution is zero)
        uint32_t x0;
        const uint32_t y0 = div_magic(blk_live ? p0 + 32u * (uint32_t)b : p0, W, I.mag_width, x0);
        uint32_t pxx = x0 + (inb ? (uint32_t)lane : 0u), py = y0;
        while (pxx >= W) {
          pxx -= W;
          ++py;
        }
        const uint32_t joff = py * I.jac_pitch + pxx * C;  // this pixel's code-Jacobian row (floats)
        float feat[8];
        bool ok = false;
        if (blk_live) {
#ifndef DFK_TC_NOPREFETCH
          asm volatile("prefetch.global.L2 [%0];" ::"l"(I.jac + joff));
#endif
          const float xn = __ldg(I.ray_tab + pxx);
          const float yn = __ldg(I.ray_tab + W + py);
          float d = __ldg(I.dpt0 + (size_t)py * I.dpt0_pitch + pxx);
          const float i0 = __ldg(I.img0 + (size_t)py * I.img0_pitch + pxx);
          if (fused) {
            // dpt0 is prx_orig: decode the depth from the pixel's code-Jacobian row -- same arithmetic as
            // update_depth_kernel (chunk fma chains, then the xor-butterfly over the 8 chunk sums, here ACROSS the 8
            // lanes that hold the chunks of one pixel), publish it, and carry on with it
            const float4 cc = *reinterpret_cast<const float4*>(&I.code[4 * (lane & 7)]);
            float mine = 0.0f;
#pragma unroll
            for (int i8 = 0; i8 < 8; ++i8) {
              const uint32_t offk = __shfl_sync(0xffffffffu, joff, 4 * i8 + (lane >> 3));
              float p = chunk_dot(load_chunk(I.jac + offk + 4 * (lane & 7), a16), cc);
              p = __fadd_rn(p, __shfl_xor_sync(0xffffffffu, p, 4));
              p = __fadd_rn(p, __shfl_xor_sync(0xffffffffu, p, 2));
              p = __fadd_rn(p, __shfl_xor_sync(0xffffffffu, p, 1));
              const float got = __shfl_sync(0xffffffffu, p, 8 * (lane & 3));  // pixel 4 i8 + (lane & 3)
              if ((lane >> 2) == i8) mine = got;
            }
            d = prx_to_depth(__fadd_rn(d, mine), I.avg_dpt);
            if (inb) I.dpt_out[(size_t)py * I.dpt_out_pitch + pxx] = d;
          }
          if (inb) {
            const Warped w = warp_ray(xn, yn, d, I.q, I.t, I.fx, I.fy, I.u0, I.v0, I.border, I.ulim, I.vlim, I.min_dpt);
            if (w.valid) {
              ok = true;
              I.valid0[(size_t)py * I.valid0_pitch + pxx] = 1.0f;  // dense_sfm.h:161
              int ix, iy;
              float fu, fv, gx, gy;
              bilin_setup(w.u, w.v, ix, iy, fu, fv);
#ifdef DFK_EXP_NOGATHER  // experiment (wrong results): sample at the pixel itself -> coalesced taps
              ix = (int)pxx < (int)W - 1 ? (int)pxx : (int)W - 2;
              iy = (int)py < (int)I.height - 1 ? (int)py : (int)I.height - 2;
#endif
              sample_grad(I.grad1, I.grad1_pitch, true, ix, iy, fu, fv, gx, gy);  // the API guarantees 8-byte rows here
              const float i1 = sample_scalar(I.img1, I.img1_pitch, ix, iy, fu, fv);
              float a[6], c00, c02, c11, c12;
              pose_jacobian_row(w, I.fx, I.fy, gx, gy, a, c00, c02, c11, c12);
              const float e = prx_jacobian(w, I.R, d, I.avg_dpt, gx, gy, c00, c02, c11, c12);
              const float diff = i0 - i1;
              const float hw = huber_weight(diff, I.huber_delta);
              feat[0] = hw * e;
#pragma unroll
              for (int f = 0; f < 6; ++f) feat[1 + f] = hw * a[f];
              feat[7] = hw * diff;
            }
          }
        }
        if (!ok) {
#pragma unroll
          for (int f = 0; f < 8; ++f) feat[f] = 0.0f;
        }
        const unsigned bal = __ballot_sync(0xffffffffu, ok);
        const int nvb = __popc(bal);
        const int mask = ((bal & 0xffffu) ? 1 : 0) | ((bal & 0xffff0000u) ? 2 : 0);  // non-empty 16-pixel groups
        // ---- the operand slot: free once the MMAs of block j - NSLOT have completed (the control thread commits them
        // to this warp's barrier; one commit per block of this warp, in order) ------------------------------------------
        const int slot = j % NSLOT;
        if (j >= NSLOT) {
          TC_WAIT(&sm.done[warp], done_phase, 0, j);
          done_phase ^= 1u;
        }
        if (nvb > 0) {
          const uint32_t sbase = smem_u32(&sm.op[slot][0]);
          // code features: iteration i8 covers pixels 4 i8 .. 4 i8 + 3; 8 lanes hold the eight 16-byte chunks of one
          // pixel's row (512 contiguous bytes per warp load).  Lane groups 0-7 / 8-15 / 16-23 / 24-31 take pixels
          // 0 / 2 / 1 / 3, so that the two rows one half-warp stores 8 bytes per lane to (code-l) sit in different banks
          const int r = ((lane >> 3) & 1) * 2 + (lane >> 4);
          const int c = lane & 7;
          // one 16-pixel group (= one k-step of the MMA) at a time, its four row loads in two batches of two (registers)
#pragma unroll
          for (int gq = 0; gq < 2; ++gq) {
            if (!((mask >> gq) & 1)) continue;  // no valid pixel in the group: no loads, no stores, no MMA
#pragma unroll
            for (int ub = 0; ub < 4; ub += DFK_TC_LOAD_BATCH) {
              float4 v[DFK_TC_LOAD_BATCH];
#pragma unroll
              for (int u = 0; u < DFK_TC_LOAD_BATCH; ++u) {
                const int kp = 16 * gq + 4 * (ub + u) + r;
                const uint32_t offk = __shfl_sync(0xffffffffu, joff, kp);
                v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
#ifdef DFK_EXP_NOJC  // experiment (wrong results): no code-Jacobian traffic
                if ((bal >> kp) & 1u) v[u] = make_float4(0.1f * (float)offk, 0.2f, 0.3f, 0.4f);
#else
                if ((bal >> kp) & 1u) v[u] = load_chunk(I.jac + offk + 4 * c, a16);
#endif
              }
#pragma unroll
              for (int u = 0; u < DFK_TC_LOAD_BATCH; ++u) {
                const int i8 = 4 * gq + ub + u;
                const float sk = __shfl_sync(0xffffffffu, feat[0], 4 * i8 + r);
                const int rr = 4 * (i8 & 1) + r;  // pixel row inside its K atom (i8 >> 1)
                const uint32_t row = sbase + (uint32_t)(i8 >> 1) * ATOM_BYTES + (uint32_t)rr * 128u;
                uint32_t h0, m0, l0, h1, m1, l1;
                split3(sk * v[u].x, sk * v[u].y, h0, m0, l0);
                split3(sk * v[u].z, sk * v[u].w, h1, m1, l1);
                sts128u(row + (uint32_t)((c ^ rr) << 4), h0, h1, m0, m1);  // atom 0, chunk c: h 4c..4c+3 | m 4c..4c+3
                // atom 1, code-l slots 16 + 4c ..: byte 32 + 8c of the row = chunk 2 + c/2, half c & 1
                sts64u(row + MN_STRIDE + (uint32_t)((((2 + (c >> 1)) ^ rr) << 4) | ((c & 1) << 3)), l0, l1);
              }
            }
          }
          // pose / residual features: this thread's pixel = K position `lane`: chunks 0 (h), 1 (m), 6 (l) of its atom-1 row;
          // one term at a time (the residuals replace the values in place) keeps the register footprint small
          {
            const int rr = lane & 7;
            const uint32_t row = sbase + MN_STRIDE + (uint32_t)(lane >> 3) * ATOM_BYTES + (uint32_t)rr * 128u;
            float f0 = feat[1], f1 = feat[2], f2 = feat[3], f3 = feat[4], f4 = feat[5], f5 = feat[6], f6 = feat[7];
#pragma unroll
            for (int term = 0; term < 3; ++term) {
              const uint32_t p0 = pack_bf16(f0, f1), p1 = pack_bf16(f2, f3), p2 = pack_bf16(f4, f5), p3 = pack_bf16(f6, 0.0f);
              const int chunk = term == 0 ? 0 : (term == 1 ? 1 : 6);
              sts128u(row + (uint32_t)((chunk ^ rr) << 4), p0, p1, p2, p3);
              if (term < 2) {
                f0 -= bf16_lo(p0); f1 -= bf16_hi(p0); f2 -= bf16_lo(p1); f3 -= bf16_hi(p1);
                f4 -= bf16_lo(p2); f5 -= bf16_hi(p2); f6 -= bf16_lo(p3);
              }
            }
          }
          fence_proxy_async_smem();  // generic-proxy writes -> visible to the MMA's operand fetch
        }
        if (lane == 0) {
          sm.meta[slot].mask = mask;
          sm.meta[slot].nvalid = nvb;
          sm.meta[slot].item_changed = seq_changed ? 1 : 0;
          sm.meta[slot].pslot = (int)I.slot;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.full[slot]);
      }
    } else if (warp == CTRL_WARP) {
      // ======================================================================= control warp
      // The whole warp walks the blocks (uniform control flow: waits, meta, chain bookkeeping are done redundantly by all
      // lanes, which keeps the loop free of divergence); one elected lane issues the MMAs, the commits and the records.
      // This loop is serial per CTA, so its length bounds the kernel: counters instead of divisions, descriptors advanced
      // by constants, one 16-byte load for the slot meta.
      {
        const bool leader = elect_one_sync();
#ifdef DFK_EXP_SMALLMMA  // experiment (wrong results): same issue pattern, a fraction of the operand fetch
        const uint32_t idesc = make_idesc_bf16(64, 16) | kIdescAMnMajor | kIdescBMnMajor;
#else
        const uint32_t idesc = make_idesc_bf16(MM, NB) | kIdescAMnMajor | kIdescBMnMajor;
#endif
        const uint32_t desc_hi = (ATOM_BYTES >> 4) | (1u << 14) | (2u << 29);  // SBO | version 1 | SWIZZLE_128B
        const uint32_t ring_lo = (smem_u32(&sm.op[0][0]) & 0x3ffffu) >> 4;
        int e = -1, tiles_in_chain = 0;
        uint32_t first = 1;
        int chain_valid = 0, cur_pslot = -1;
        bool slot_fresh = true;
        unsigned int inliers = 0;
        // close chain e: publish its record, then let the drain warps go once its MMAs have completed
        auto close_chain = [&](bool item_end, bool last) {
          if (leader) {
            ChainRec& r = sm.chain[e & 1];
            r.valid = chain_valid;
            r.pslot = cur_pslot;
            r.fresh = slot_fresh ? 1 : 0;
            r.item_end = item_end ? 1 : 0;
            r.inliers = inliers;
            r.last = last ? 1 : 0;
            umma_commit(&sm.d_full[e & 1]);
            mbar_arrive(&sm.d_full[e & 1]);
          }
          slot_fresh = false;
        };
        int slot = 0, next_w = NSLOT % NFE;
        uint32_t full_phase = 0;
        for (int j = 0; j < nblocks; ++j) {
          TC_WAIT(&sm.full[slot], full_phase, 1, j);
          tc_fence_after();
          const int4 meta = *reinterpret_cast<const int4*>(&sm.meta[slot]);  // mask, nvalid, item_changed, pslot
          if ((j & 3) == 0) {
            if (j == 0 || meta.z != 0 || tiles_in_chain == kFlushTiles) {
              if (j > 0) close_chain(meta.z != 0, false);
              e += 1;
              tiles_in_chain = 0;
              chain_valid = 0;
              first = 1;
              const int use = e >> 1;  // n-th use of this accumulator buffer (and of its record)
              if (use >= 1) {
                TC_WAIT(&sm.d_empty[e & 1], (uint32_t)(use - 1) & 1u, 2, e);
                tc_fence_after();
              }
              if (j == 0 || meta.z != 0) {
                cur_pslot = meta.w;
                slot_fresh = true;
                inliers = 0;
              }
            }
            tiles_in_chain += 1;
          }
          chain_valid += meta.y;
          inliers += (unsigned)meta.y;
          if (leader) {
            const uint32_t d_addr = tbase + D_COL + NB * (uint32_t)(e & 1);
            const uint32_t lo = ring_lo + (uint32_t)slot * (SLOT_BYTES >> 4);
            const uint32_t dlo = lo | ((MN_STRIDE >> 4) << 16);  // A and B: same start, same LBO; B is the first 80 rows
#ifdef DFK_EXP_NOMMA  // experiment (wrong results): no tensor-core work, no operand fetch
            if (meta.x == 3) {
              first = 0;
            } else
#endif
            if (meta.x == 3) {  // the common case: two k-steps back to back
              umma_bf16_ss_x2(d_addr, dlo, desc_hi, idesc, first ^ 1u, (2 * ATOM_BYTES) >> 4);
              first = 0;
            } else {
#pragma unroll
              for (int gq = 0; gq < 2; ++gq) {
                if ((meta.x >> gq) & 1) {
                  const uint64_t d = ((uint64_t)desc_hi << 32) | (dlo + (uint32_t)gq * ((2 * ATOM_BYTES) >> 4));  // two K atoms per k-step
                  umma_bf16_ss(d_addr, d, d, idesc, first == 0);
                  first = 0;
                }
              }
            }
            umma_commit(&sm.done[next_w]);  // the slot's next user: block j + NSLOT
          } else if (meta.x != 0) {
            first = 0;
          }
          if (++slot == NSLOT) {
            slot = 0;
            full_phase ^= 1u;
          }
          if (++next_w == NFE) next_w = 0;
        }
        close_chain(true, true);
      }
    } else {
      // ======================================================================= drain warps (TMEM lanes 0..127)
      const int ow = warp & 3;  // the lane quarter this warp may access
      const uint32_t lane_taddr = tbase + ((uint32_t)(ow * 32) << 16);
      const int row = ow * 32 + lane;  // TMEM lane == row of the partial
      for (int e = 0;; ++e) {
        const int b = e & 1, use = e >> 1;
        TC_WAIT(&sm.d_full[b], (uint32_t)use & 1u, 3, e);
        tc_fence_after();
        const ChainRec rec = sm.chain[b];
        float* P = partials + (size_t)rec.pslot * kTcPartialFloats;
        // five passes of 16 accumulator columns keep the register footprint small.  The first chain of an item in
        // this CTA stores, later chains add with fire-and-forget red.global.add.f32: this thread is the only writer
        // of its row and issues its updates in program order, so the sum order is fixed.
#ifdef DFK_EXP_NODRAIN  // experiment (wrong results): chains are not moved to global memory
        if (false) {
#else
        if (rec.valid > 0 || rec.fresh) {
#endif
#pragma unroll 1
          for (int pass = 0; pass < kTcCols / 16; ++pass) {
            uint32_t v[16];
#pragma unroll
            for (int c = 0; c < 16; ++c) v[c] = 0u;
            if (rec.valid > 0) {
              tmem_ld_x16(lane_taddr + D_COL + NB * b + 16 * pass, v);
              tmem_wait_ld();
            }
            // column-major partial: this lane's row at column c is P[c * kTcRowsPad + row] -> a warp writes 128
            // contiguous bytes per column
            float* dcol = P + (16 * pass) * kTcRowsPad + row;
            if (rec.fresh) {
#pragma unroll
              for (int c = 0; c < 16; ++c)
                __stcg(dcol + c * kTcRowsPad, __uint_as_float(v[c]));
            } else {
#pragma unroll
              for (int c = 0; c < 16; ++c)
                asm volatile("red.global.add.f32 [%0], %1;" ::"l"(dcol + c * kTcRowsPad), "f"(__uint_as_float(v[c])) : "memory");
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.d_empty[b]);
        if (rec.item_end && ow == 0 && lane == 0) reinterpret_cast<unsigned int*>(P)[kTcRowsPad * kTcCols] = rec.inliers;
        if (rec.last) break;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == CTRL_WARP) tmem_dealloc(tbase, TMEM_COLS);
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

size_t sfm_tc_smem_bytes() { return sizeof(Smem) + 1024; }

cudaError_t launch_sfm_tc(const SfmItemDev* items_dev, const SfmLaunchPlan& plan, bool build_ray_tables,
                          float* partials_dev, cudaStream_t stream, cudaEvent_t ev_start, cudaEvent_t ev_stop)
{
  const size_t smem = sizeof(Smem) + 1024;  // + the slack the kernel aligns the operand ring with
  static const cudaError_t attr_err =  // once per process, not once per launch
      cudaFuncSetAttribute(sfm_step_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(Smem) + 1024));
  cudaError_t err = attr_err;
  if (err != cudaSuccess) return err;
  if (build_ray_tables) {  // only when the work list names a camera level the handle has no table for yet
    sfm_ray_tables_kernel<<<plan.num_items, 256, 0, stream>>>(items_dev, nullptr);
    err = cudaGetLastError();
    if (err != cudaSuccess) return err;
  }
  if (ev_start) cudaEventRecord(ev_start, stream);
  sfm_step_tc_kernel<<<plan.num_ctas, THREADS, smem, stream>>>(items_dev, plan.num_items, plan.num_tiles, partials_dev);
  if (ev_stop) cudaEventRecord(ev_stop, stream);
  return cudaGetLastError();
}

}  // namespace dfk
