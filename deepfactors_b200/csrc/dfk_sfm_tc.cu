// dfk_sfm_tc.cu -- SfmAligner::RunStep hot path, tcgen05 tensor-core Gram variant (sm_100a, C = 32).
//
// Same contract as dfk_sfm_fp32.cu (replaces kernel_step_calculate + DenseSfm + the two-kernel
// reduction of sources/cuda/cu_sfmaligner.cpp:40-70,149-185, dense_sfm.h:133-201), different engine
// for the reduced Gram  G = sum_p m_p^T m_p,  m = w*[ e*jc (32) | a (6) | diff (1) ]  (39 features):
//
//   Split precision ("3xTF32" folded into ONE MMA): every feature value v is split exactly into
//   h = the bits the tensor core keeps (fp32 -> tf32 is a truncation of the low 13 mantissa bits on
//   this hardware, measured by tools/umma_probe.cu) and l = v - h.  With A = [h rows ; l rows] and B = h,
//   one tcgen05.mma.kind::tf32 per 8 pixels yields HH = sum h h^T and LH = sum l h^T;
//   G = HH + LH + LH^T  drops only the l*l terms (~2^-22).
//
//   Round-2 data path ("direct operands"): both MMA operands are read from shared memory in the PIXEL-major
//   ("MN-major") layout a per-pixel front-end writes naturally -- canonical layout SWIZZLE_128B_BASE32B, the one the
//   hardware defines for transposed 32-bit operands (tools/umma_probe_mn.cu: atoms of 32 features x 4 pixels, the
//   32-byte chunk c of pixel row r stored at chunk position c ^ r).  There is no operand-building stage, no TMEM A
//   operand, no transposition and no tile staging any more: a front-end warp goes from global memory to a finished
//   operand block on its own and hands it to the MMA issuer with one mbarrier arrival.
//
//   Per CTA (512 threads, 2 CTAs / SM, 128 TMEM columns each; the drain / control warpgroup gives registers back with
//   setmaxnreg -- ptxas allocates for the launch bound, 64 registers, whatever setmaxnreg.inc says afterwards):
//     warps 0-11  front-end : warp w owns the CTA's 32-pixel blocks j = w, w+12, ... (a 128-pixel tile = 4 blocks).
//                             One thread per pixel: (optional depth decode,) exact-order validity chain, bilinear
//                             gathers, Jacobian row, Huber -> s = w*e, w*a[6], w*diff.  Then the block's code-Jacobian
//                             rows are read straight from global memory, COALESCED (lane = 16-byte chunk lane&7 of pixel
//                             4i + lane/8; the rows were prefetched into L2 when the block started), scaled by the
//                             pixel's s (one shuffle), split into h / l and stored as the code-h and code-l atoms of
//                             an operand slot; each thread adds its own pixel's 7 pose/residual values (h and l) to the
//                             third atom.  Invalid pixels contribute exact zeros; 8-pixel groups without a valid
//                             pixel are skipped altogether (no loads, no MMA).
//     warps 12-14 drain     : pull a finished accumulation chain out of TMEM (tcgen05.ld) and add it in round-to-nearest
//                             fp32 to the CTA's partial in global memory (single writer per address, program order).
//     warp 15     control   : the whole warp walks the blocks in order (uniform control flow), one elected lane issues:
//                             waits for the slot, one MMA per non-empty 8-pixel group (A = the slot's atoms 0,1,2 (+1
//                             ignored atom: M = 128), B = atoms 0 and 2: N = 48), commits the slot to the warp that
//                             uses it next, cuts the chains and publishes their records to the drain warps.  It also
//                             allocates TMEM.  This loop is serial per CTA: ~60 instructions per block.
//   Operand slots: DFK_TC_SLOTS = 7 of 12 KB, block j uses slot j % 7 (an odd count rotates every warp through every slot;
//   more slots cost L1 capacity the gathers want: 5 / 6 / 7 / 8 slots = 0.152 / 0.158 / 0.148 / 0.172 ms).  Atom a (a = 0 code-h, 1 code-l, 2 pose: h at features 0-7, l at 8-15, rest zero) at
//   a * 4096; inside, K atom q (pixels 4q..4q+3) at q * 512, pixel row r at r * 128, 32-byte chunk c at (c ^ r) * 32.
//   Accumulator rows (TMEM lanes): 0-31 code-h, 32-63 code-l, 64-71 pose-h, 72-79 pose-l; columns 0-31 code, 32-39 pose.
//   The fp32 accumulator in TMEM adds with truncation (measured: ~ -2^-24 relative per k-step), so a chain is cut every
//   kFlushTiles tiles.
//
// The static tile->CTA assignment, the in-item tile permutation, the per-CTA partials and the wide deterministic
// finalize are those of the fp32 kernel.  The round-1 kernel (TMA-staged tiles, operand warps transposing into a TMEM
// A operand), two intermediate redesigns and the three-term bf16 split of this kernel are kept under tools/experiments/
// with their measurements (profiles/README.md).
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
constexpr int NFE = 12;              // front-end warps
constexpr int THREADS = 512;         // 12 front-end warps, 3 drain warps, control
#ifndef DFK_TC_SLOTS
#define DFK_TC_SLOTS 7
#endif
constexpr int NSLOT = DFK_TC_SLOTS;  // operand slots; block j uses slot j % NSLOT
constexpr uint32_t ATOM_BYTES = 512;              // 32 features x 4 pixels
constexpr uint32_t MN_STRIDE = 8 * ATOM_BYTES;    // the 8 K atoms of one MN atom are contiguous
constexpr uint32_t SLOT_BYTES = 3 * MN_STRIDE;    // 12288
constexpr int NB = 48;           // MMA N (40 used)
constexpr int MM = 128;          // MMA M (80 used)
#ifndef DFK_FLUSH_TILES
#define DFK_FLUSH_TILES 8
#endif
constexpr int kFlushTiles = DFK_FLUSH_TILES;  // TMEM accumulation chain length (tiles)
constexpr uint32_t TMEM_COLS = 128;
constexpr uint32_t D_COL = 0;  // [0,48), [48,96): two accumulators

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
  // >= 4 KB follow the ring: the (ignored) fourth MN atom of an M = 128 operand in the last slot reads into them
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
static_assert(sizeof(ItemSmem) * NFE >= MN_STRIDE, "the ring's tail pad");

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
__device__ __forceinline__ float tf32_trunc(float v) { return __uint_as_float(__float_as_uint(v) & 0xffffe000u); }

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

__device__ __forceinline__ void sts128(uint32_t addr, float a, float b, float c, float d)
{
  // no "memory" clobber: the front-end's only plain shared-memory accesses are reads of its item copy and the slot
  // meta (written after the proxy fence, which does carry the clobber); volatile keeps the stores ordered with the fence
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d));
}

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
      mbar_init(&sm.d_empty[b], 3);
    }
    mbar_fence_init();
  }
  // the never-written parts of the ring (pose atom chunks 2, 3) must hold zeros
  for (int e = tid; e < (int)(NSLOT * SLOT_BYTES / 16); e += THREADS)
    reinterpret_cast<float4*>(&sm.op[0][0])[e] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (warp == 15) {
    tmem_alloc(&sm.tmem_base, TMEM_COLS);
    tmem_relinquish();
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = sm.tmem_base;

  // register budget per role (warpgroup granularity): 12 x 32 x 72 + 4 x 32 x 40 = 32768 = half the register file
  if (warp < NFE) {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 72;");
  } else {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
  }

  if (ntiles > 0) {
    if (warp < NFE) {
      // ======================================================================= front-end warps
      ItemSmem& I = sm.item[warp];
      int it = 0;
      int cur_item = -1;
      uint32_t item_lo = 0, item_hi = 0;  // global tile range of the item in shared memory
      uint32_t done_phase = 0;
      for (int j = warp; j < nblocks; j += NFE) {
        const int i = j >> 2, b = j & 3;
        const int g = g_lo + i;
        if (cur_item < 0 || (uint32_t)g >= item_hi) {
          while ((uint32_t)g >= items[it].tile_begin + items[it].num_tiles) ++it;
          __syncwarp();
          load_item(I, items[it], lane, cta);
          __syncwarp();
          cur_item = it;
          item_lo = I.tile_begin;
          item_hi = I.tile_begin + I.num_tiles;
        }
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
        const float* __restrict__ jac = I.jac;
        const uint32_t joff = py * I.jac_pitch + pxx * C;  // this pixel's code-Jacobian row (floats)
        float feat[8];
        bool ok = false;
        if (blk_live) {
#ifndef DFK_TC_NOPREFETCH
          asm volatile("prefetch.global.L2 [%0];" ::"l"(jac + joff));
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
              float p = chunk_dot(load_chunk(jac + offk + 4 * (lane & 7), a16), cc);
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
        const int mask = ((bal & 0xffu) ? 1 : 0) | ((bal & 0xff00u) ? 2 : 0) | ((bal & 0xff0000u) ? 4 : 0) |
                         ((bal & 0xff000000u) ? 8 : 0);
        // ---- the operand slot: free once the MMAs of block j - NSLOT have completed (the control thread commits them
        // to this warp's barrier; one commit per block of this warp, in order) ------------------------------------------
        const int slot = j % NSLOT;
        if (j >= NSLOT) {
          TC_WAIT(&sm.done[warp], done_phase, 0, j);
          done_phase ^= 1u;
        }
        if (nvb > 0) {
          const uint32_t sbase = smem_u32(&sm.op[slot][0]);
          // code atoms: iteration i8 covers the K atom of pixels 4 i8 .. 4 i8 + 3; this lane holds the 16-byte chunk
          // lane & 7 of pixel 4 i8 + lane / 8 -> 512 contiguous bytes per warp load
          const int r = lane >> 3;
          const uint32_t dst0 = sbase + (uint32_t)r * 128u + (uint32_t)(((((lane & 7) >> 1) ^ r) << 5) | ((lane & 1) << 4));
          float4 v[8];
#pragma unroll
          for (int i8 = 0; i8 < 8; ++i8) {
            const int kp = 4 * i8 + r;
            const uint32_t offk = __shfl_sync(0xffffffffu, joff, kp);
            v[i8] = make_float4(0.f, 0.f, 0.f, 0.f);
#ifdef DFK_EXP_NOJC  // experiment (wrong results): no code-Jacobian traffic
            if ((bal >> kp) & 1u) v[i8] = make_float4(0.1f * (float)offk, 0.2f, 0.3f, 0.4f);
#else
            if ((bal >> kp) & 1u) v[i8] = load_chunk(jac + offk + 4 * (lane & 7), a16);
#endif
          }
#pragma unroll
          for (int i8 = 0; i8 < 8; ++i8) {
            const float sk = __shfl_sync(0xffffffffu, feat[0], 4 * i8 + r);
            if ((mask >> (i8 >> 1)) & 1) {
              const float h0 = sk * v[i8].x, h1 = sk * v[i8].y, h2 = sk * v[i8].z, h3 = sk * v[i8].w;
              const uint32_t dst = dst0 + (uint32_t)i8 * ATOM_BYTES;
#ifdef DFK_EXP_NOSTORE  // experiment (wrong results): one store instead of two
              sts128(dst, h0 + tf32_trunc(h1), h1, h2 - tf32_trunc(h3), h3);
#else
              sts128(dst, h0, h1, h2, h3);  // the tensor core truncates: h rows carry the raw values
              sts128(dst + MN_STRIDE, h0 - tf32_trunc(h0), h1 - tf32_trunc(h1), h2 - tf32_trunc(h2), h3 - tf32_trunc(h3));
#endif
            }
          }
          // pose atom: this thread's pixel = K position `lane`: h of the 7 pose / residual values in chunk 0, l in chunk 1.
          // Lanes of odd K atoms write the two 16-byte halves of a chunk in the opposite order: the 8 lanes of a
          // quarter warp (two K atoms x four rows) then hit 8 different bank groups
          {
            const int rp = lane & 3;
            const bool odd = (lane & 4) != 0;
            const uint32_t row = sbase + 2u * MN_STRIDE + (uint32_t)(lane >> 2) * ATOM_BYTES + (uint32_t)rp * 128u;
            const uint32_t ch = row + ((uint32_t)rp << 5) + (odd ? 16u : 0u);         // chunk 0 ^ rp
            const uint32_t cl = row + ((uint32_t)(rp ^ 1) << 5) + (odd ? 16u : 0u);   // chunk 1 ^ rp
            const float a0 = feat[1], a1 = feat[2], a2 = feat[3], a3 = feat[4], b0 = feat[5], b1 = feat[6], b2 = feat[7];
            const float la0 = a0 - tf32_trunc(a0), la1 = a1 - tf32_trunc(a1), la2 = a2 - tf32_trunc(a2), la3 = a3 - tf32_trunc(a3);
            const float lb0 = b0 - tf32_trunc(b0), lb1 = b1 - tf32_trunc(b1), lb2 = b2 - tf32_trunc(b2);
            sts128(ch, odd ? b0 : a0, odd ? b1 : a1, odd ? b2 : a2, odd ? 0.0f : a3);
            sts128(ch ^ 16u, odd ? a0 : b0, odd ? a1 : b1, odd ? a2 : b2, odd ? a3 : 0.0f);
            sts128(cl, odd ? lb0 : la0, odd ? lb1 : la1, odd ? lb2 : la2, odd ? 0.0f : la3);
            sts128(cl ^ 16u, odd ? la0 : lb0, odd ? la1 : lb1, odd ? la2 : lb2, odd ? la3 : 0.0f);
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
    } else if (warp == 15) {
      // ======================================================================= control warp
      // The whole warp walks the blocks (uniform control flow: waits, meta, chain bookkeeping are done redundantly by all
      // lanes, which keeps the loop free of divergence); one elected lane issues the MMAs, the commits and the records.
      // This loop is serial per CTA, so its length bounds the kernel: counters instead of divisions, descriptors advanced
      // by constants, one 16-byte load for the slot meta.
      {
        const bool leader = elect_one_sync();
#ifdef DFK_EXP_SMALLMMA  // experiment (wrong results): same issue pattern, a fraction of the operand fetch
        const uint32_t idesc = make_idesc_tf32(64, 16) | kIdescAMnMajor | kIdescBMnMajor;
#else
        const uint32_t idesc = make_idesc_tf32(MM, NB) | kIdescAMnMajor | kIdescBMnMajor;
#endif
        const uint32_t desc_hi = (ATOM_BYTES >> 4) | (1u << 14) | (1u << 29);  // SBO | version 1 | SWIZZLE_128B_BASE32B
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
#ifdef DFK_EXP_NOMMA
            if (meta.x == 15) {
              first = 0;
            } else
#endif
            if (meta.x == 15) {  // the common case: four k-steps back to back
              umma_tf32_ss_x4(d_addr, lo | ((MN_STRIDE >> 4) << 16), lo | ((2 * MN_STRIDE >> 4) << 16), desc_hi, idesc, first ^ 1u,
                              (2 * ATOM_BYTES) >> 4);
              first = 0;
            } else {
#pragma unroll
              for (int gq = 0; gq < 4; ++gq) {
                if ((meta.x >> gq) & 1) {
                  const uint32_t adv = (uint32_t)gq * ((2 * ATOM_BYTES) >> 4);  // two K atoms per k-step
#ifndef DFK_EXP_NOMMA  // experiment (wrong results): no tensor-core work, no operand fetch
                  umma_tf32_ss(d_addr, ((uint64_t)desc_hi << 32) | (lo + adv) | ((MN_STRIDE >> 4) << 16),
                               ((uint64_t)desc_hi << 32) | (lo + adv) | ((2 * MN_STRIDE >> 4) << 16), idesc, first == 0);
#endif
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
      // ======================================================================= drain warps (TMEM lanes 0..95)
      const int ow = warp & 3;  // 0, 1, 2 = the lane quarter this warp may access
      const uint32_t lane_taddr = tbase + ((uint32_t)(ow * 32) << 16);
      const int row = ow * 32 + lane;  // TMEM lane == row of the partial
      for (int e = 0;; ++e) {
        const int b = e & 1, use = e >> 1;
        TC_WAIT(&sm.d_full[b], (uint32_t)use & 1u, 3, e);
        tc_fence_after();
        const ChainRec rec = sm.chain[b];
        float* P = partials + (size_t)rec.pslot * kTcPartialFloats;
        // three passes of 16 accumulator columns keep the register footprint small.  The first chain of an item in
        // this CTA stores, later chains add with fire-and-forget red.global.add.f32: this thread is the only writer
        // of its row and issues its updates in program order, so the sum order is fixed.
        if (rec.valid > 0 || rec.fresh) {
#pragma unroll 1
          for (int pass = 0; pass < 3; ++pass) {
            const int ncol = pass < 2 ? 16 : kTcCols - 32;  // columns 40..47 are padding
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
                if (c < ncol) __stcg(dcol + c * kTcRowsPad, __uint_as_float(v[c]));
            } else {
#pragma unroll
              for (int c = 0; c < 16; ++c)
                if (c < ncol)
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
  if (warp == 15) tmem_dealloc(tbase, TMEM_COLS);
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
