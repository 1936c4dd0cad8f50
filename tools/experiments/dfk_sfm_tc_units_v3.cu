// dfk_sfm_tc.cu -- SfmAligner::RunStep hot path, tcgen05 tensor-core Gram variant (sm_100a, C = 32).  Round-2 design.
//
// Replaces kernel_step_calculate + DenseSfm + the two-kernel reduction of the reference
// (sources/cuda/cu_sfmaligner.cpp:40-70,149-185, sources/common/algorithm/dense_sfm.h:133-201).  The reduced Gram
//     G = sum_p m_p^T m_p,   m = w*[ e*jc (32) | a (6) | diff (1) ] , plus the constant feature 1 of every valid pixel
// (40 features; G[39][39] IS the inlier count, exact in fp32) runs on the 5th-generation tensor cores with split
// precision folded into ONE MMA: every feature value v is split exactly into h = the bits kind::tf32 keeps (a truncation
// of the low 13 mantissa bits on this hardware, measured by tools/umma_probe.cu) and l = v - h; A = [h rows ; l rows]
// (80 of M = 128 rows, in TMEM), B = h (40 of N = 48 columns, K-major in shared memory); one tcgen05.mma.kind::tf32 per
// 8 pixels gives HH = sum h h^T and LH = sum l h^T, and G = HH + LH + LH^T drops only the l*l terms (~2^-22).  The
// finalize kernel (dfk_sfm_finalize.cu) recombines and expands to the reference's (12+C) layout.
//
// Work decomposition: a BLOCK is 32 pixels of one image row (lane = pixel), a UNIT is two consecutive blocks (the
// granule of every hand-off), a PATCH is a 32-pixel-wide strip x ~32 rows; the patches of an item are visited in a
// golden-ratio permuted order (balances clustered invalid regions over the CTAs), the rows of a patch consecutively
// (neighbouring rows share their bilinear taps in L1).  CTA c owns a contiguous range of the global unit sequence
// (static => bitwise reproducible results); items hold an even number of blocks, so a unit never straddles two items.
//
// Measured facts that shape the roles (profiles/README.md, round 2): a hand-off (mbarrier wait or release-arrive) costs
// 150-350 cycles on a loaded SM, so a serial stage that synchronises per 32 pixels cannot run faster than ~800 cycles
// per block even with nothing to do; one warp streams the ~150-instruction operand build of a block in ~350 cycles.
// Hence: hand-offs per UNIT, two operand trios working on alternate units, no hand-off the control thread can avoid.
//
// Roles per CTA (512 threads, 2 CTAs / SM, 256 TMEM columns each):
//   warps 0-7   front-end : each warp processes whole blocks on its own (block b -> warp b % 8), one thread per pixel:
//                           coalesced loads of dpt0 / img0, exact-order validity chain, bilinear gathers of img1 /
//                           grad1, Jacobian row, Huber weight -> nine scalars per pixel (s = w*e, w*a[6], w*diff, 1)
//                           into a feat slot in shared memory.  Never touches the code Jacobian (unless it decodes depth).
//   warp 11     producer  : lane 0 issues the cp.async.bulk copies (TMA engine, SASS UBLKCP) of the blocks' 32 x 128-byte
//                           code-Jacobian row segments into a 16-block ring, up to 8 units ahead of their use.
//   warps 12-14 trio A,   : operand builders of the even / odd units; lane = feature row.  h warp: 32 conflict-free LDS
//   warps 8-10  trio B      of the raw rows per block, v = s * jc, tcgen05.st into TMEM lanes 0-31 (the tensor core
//                           truncates v to h) and v K-major into shared memory as B;  l warp: the same loads,
//                           l = v - trunc(v) into lanes 32-63;  p warp: h and l of the 8 pose / residual / count
//                           features (lanes 64-79) + their B rows.  Trio A also drains the accumulation chains.
//   warp 15     control   : lane 0 issues 4 tcgen05.mma (M128 N48 K8, A from TMEM, B from shared memory) per non-empty
//                           block and the tcgen05.commit arrivals; allocates TMEM.
// The fp32 accumulator in TMEM adds with truncation (measured ~ -2^-24 relative per k-step), so an accumulation chain
// is cut every 16 units (<= 128 k-steps) and at item boundaries: trio A pulls the finished chain out of TMEM
// (tcgen05.ld) and adds it in round-to-nearest fp32 to the CTA's partial in global memory (single writer per address,
// program order => reproducible).  Blocks without a valid pixel cost a validity test and nothing else.
//
// Phase timers: -DDFK_TC_TIMERS + env DFK_TC_DEBUG=1 prints per-role cycle sums per block.
#include <cuda_runtime.h>
#include <stddef.h>
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
constexpr int BLK = kTcBlockPixels;  // 32
constexpr int NFE = 8;               // front-end warps 0..7
constexpr int W_TRIO_B = 8;          // warps 8, 9, 10: h, l, p of the odd units  (TMEM lane quarters 0, 1, 2)
constexpr int W_PROD = 11;           // lane 0: bulk-copy producer
constexpr int W_TRIO_A = 12;         // warps 12, 13, 14: h, l, p of the even units (TMEM lane quarters 0, 1, 2)
constexpr int W_CTRL = 15;           // lane 0: MMA issue
constexpr int NITEM = NFE + 3;       // item heads in shared memory: FE warps, producer, trio-A h warp, control
constexpr int NRU = 8;               // code-Jacobian ring: units (two 4 KB stages each)
constexpr int NFU = 8;               // feat slots: units
constexpr int THREADS = 512;
constexpr int NB = 48;               // MMA N (40 used)
constexpr int MM = 128;              // MMA M (80 used)
#ifndef DFK_CHAIN_UNITS
#define DFK_CHAIN_UNITS 16
#endif
constexpr int kChainUnits = DFK_CHAIN_UNITS;  // TMEM accumulation chain length: 16 units = 32 blocks = 128 k-steps
constexpr uint32_t TMEM_COLS = 256;
constexpr uint32_t A_COL = 0;                 // [0, 128): trio t builds unit u into columns [64 t, 64 t + 64)
constexpr uint32_t D_COL = 128;               // two accumulators of NB columns
constexpr uint32_t B_SBO = (BLK / 4) * 128;   // 1024 B between 8-row groups
constexpr uint32_t B_BLOCK_BYTES = (NB / 8) * B_SBO;  // 6144 B per block
constexpr int NFEAT = 9;                      // s, wa0..5, wr, 1
constexpr int FEAT_STRIDE = BLK + 4;          // floats per feature row: the p warp's lanes read the same pixel chunk of 8
                                              // rows -> 8 different bank groups
static_assert(D_COL + 2 * NB <= TMEM_COLS, "TMEM budget");

struct BlockMeta {
  int nv;             // valid pixels of the block (0: nothing to build / multiply)
  int n;              // pixels of the block inside the image (0..32)
  int bulk;           // its code-Jacobian rows travel through the ring (else the operand warps read global memory)
  int pad;
  const float* jrow;  // global address of the block's code-Jacobian row segment
  long long pad2;
};

// what a role keeps of its current item (shared memory, one copy per role instance)
struct ItemHead {
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
  float* dpt_out;
  const float* code;
  uint32_t img0_pitch, img1_pitch, dpt0_pitch, valid0_pitch, jac_pitch, grad1_pitch, dpt_out_pitch;
  uint32_t width, height, blk_begin, num_blocks, perm_mul, flags, pslot;
  uint32_t ph, ngroups, npatches, mag_ph, mag_np, mag_ng;
};

struct Smem {
  alignas(128) float jc[2 * NRU][BLK * C];                // 64 KB
  alignas(128) unsigned char B[2][2][B_BLOCK_BYTES];      // [trio][block of the unit]  24 KB
  alignas(16) float feat[2 * NFU][NFEAT][FEAT_STRIDE];    // 20.25 KB
  alignas(16) BlockMeta meta[2 * NFU];
  alignas(16) ItemHead item[NITEM];
  alignas(8) uint64_t tma_full[NRU];     // 2 arrivals (producer, one per block) + transaction bytes
  uint64_t stage_empty[NRU];             // h and l warp of the consuming trio
  uint64_t feat_full[NFU];               // the two front-end warps of the unit
  uint64_t feat_empty[NFU];              // the three warps of the consuming trio
  uint64_t ab_full[2];                   // the three warps of the trio
  uint64_t ab_empty[2];                  // tcgen05.commit
  uint64_t d_full[2];                    // tcgen05.commit
  uint64_t d_empty[2];                   // the three warps of trio A
  uint32_t umeta[2];                     // per trio: bit j = block j of the unit has valid pixels
  uint32_t chain_nz[2];                  // per accumulator: the chain issued at least one MMA
  uint32_t tmem_base;
};
static_assert(sizeof(Smem) <= 115712, "two CTAs per SM");

// a / b and a % b through the precomputed mag (floor(2^32 / b), 0xffffffff for b == 1): multiply-high + one correction
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

__device__ __forceinline__ void load_item_head(ItemHead& dst, const SfmItemDev& src, int lane, int cta)
{
  if (lane < 4) dst.q[lane] = src.q[lane];
  if (lane < 3) dst.t[lane] = src.t[lane];
  if (lane >= 8 && lane < 17) dst.R[lane - 8] = src.R[lane - 8];
  if (lane == 17) {
    dst.fx = src.fx; dst.fy = src.fy; dst.u0 = src.u0; dst.v0 = src.v0;
    dst.border = src.border; dst.ulim = src.ulim; dst.vlim = src.vlim;
    dst.min_dpt = src.min_dpt; dst.avg_dpt = src.avg_dpt; dst.huber_delta = src.huber_delta;
  }
  if (lane == 18) {
    dst.img0 = src.img0; dst.img1 = src.img1; dst.dpt0 = src.dpt0; dst.valid0 = src.valid0;
    dst.jac = src.jac; dst.grad1 = src.grad1; dst.dpt_out = src.dpt_out; dst.code = src.code;
  }
  if (lane == 19) {
    dst.img0_pitch = src.img0_pitch; dst.img1_pitch = src.img1_pitch; dst.dpt0_pitch = src.dpt0_pitch;
    dst.valid0_pitch = src.valid0_pitch; dst.jac_pitch = src.jac_pitch; dst.grad1_pitch = src.grad1_pitch;
    dst.dpt_out_pitch = src.dpt_out_pitch;
  }
  if (lane == 20) {
    dst.width = src.width; dst.height = src.height; dst.blk_begin = src.tile_begin; dst.num_blocks = src.num_tiles;
    dst.perm_mul = src.perm_mul; dst.flags = src.flags;
    dst.pslot = src.partial_begin + (uint32_t)cta - src.first_cta;
    dst.ph = src.tc_ph; dst.ngroups = src.tc_ngroups; dst.npatches = src.tc_npatches;
    dst.mag_ph = src.tc_mag_ph; dst.mag_np = src.tc_mag_np; dst.mag_ng = src.tc_mag_ng;
  }
}

// block k of an item -> first pixel (x0, y); null: padding block (below the last row of its patch, or the one that
// makes the item's block count even)
__device__ __forceinline__ bool decode_block(const ItemHead& I, uint32_t k, uint32_t& x0, uint32_t& y)
{
  uint32_t r, pid, grp;
  const uint32_t qd = div_magic(k, I.ph, I.mag_ph, r);
  if (qd >= I.npatches) {
    x0 = 0;
    y = 0;
    return false;
  }
  div_magic(qd * I.perm_mul, I.npatches, I.mag_np, pid);  // host guarantees qd * perm_mul < 2^32
  const uint32_t strip = div_magic(pid, I.ngroups, I.mag_ng, grp);
  x0 = strip * BLK;
  y = grp * I.ph + r;
  return y < I.height;
}

// ---- optional phase timers (clock64 sums per role), enabled with the env var DFK_TC_DEBUG=1 ----------
__device__ unsigned long long g_dbg[32];
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

// Position of a unit in the CTA's sequence: which item it belongs to and which accumulation chain.  Chains are cut at
// item boundaries and every kChainUnits units inside an item -- a function of the position alone, so the control thread
// and trio A walk the same chain sequence without talking to each other.
struct SeqWalker {
  const SfmItemDev* items;
  int gu_lo;         // first global unit of the CTA
  int it = 0;        // current item
  int item_first = 0, item_end = 0;  // CTA-local unit range of the current item (clipped below at 0)
  int e = -1;        // chain index of the current unit
  int chain_first = 0;
  bool have = false;
  __device__ __forceinline__ SeqWalker(const SfmItemDev* p, int g) : items(p), gu_lo(g) {}
  // advance to unit u (u must not decrease); returns true if u starts a chain
  __device__ __forceinline__ bool step(int u)
  {
    bool starts = false;
    if (!have || u >= item_end) {
      const uint32_t g = (uint32_t)(2 * (gu_lo + u));
      while (g >= items[it].tile_begin + items[it].num_tiles) ++it;
      item_first = max(0, (int)(items[it].tile_begin / 2) - gu_lo);
      item_end = (int)((items[it].tile_begin + items[it].num_tiles) / 2) - gu_lo;
      have = true;
      starts = true;
    } else if (u - chain_first >= kChainUnits) {
      starts = true;
    }
    if (starts) {
      e += 1;
      chain_first = u;
    }
    return starts;
  }
  __device__ __forceinline__ bool chain_is_first_of_item() const { return chain_first == item_first; }
};

__global__ void __launch_bounds__(THREADS, 2)
sfm_step_tc_kernel(const SfmItemDev* __restrict__ items, int num_items, int num_units, float* __restrict__ partials,
                   int dbg)
{
  extern __shared__ __align__(128) unsigned char smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const int cta = blockIdx.x;
  const int G = gridDim.x;
  const int gu_lo = (int)(((long long)cta * num_units) / G);
  const int gu_hi = (int)(((long long)(cta + 1) * num_units) / G);
  const int nunits = gu_hi - gu_lo;
  const int nblk = 2 * nunits;
  const int g_lo = 2 * gu_lo;
  (void)num_items;

  // ---- one-time setup ---------------------------------------------------------------------------
  if (tid == 0) {
    for (int s = 0; s < NRU; ++s) {
      mbar_init(&sm.tma_full[s], 2);
      mbar_init(&sm.stage_empty[s], 2);
    }
    for (int s = 0; s < NFU; ++s) {
      mbar_init(&sm.feat_full[s], 2);
      mbar_init(&sm.feat_empty[s], 3);
    }
    for (int t = 0; t < 2; ++t) {
      mbar_init(&sm.ab_full[t], 3);
      mbar_init(&sm.ab_empty[t], 1);
      mbar_init(&sm.d_full[t], 1);
      mbar_init(&sm.d_empty[t], 3);
    }
    mbar_fence_init();
  }
  // B rows 40..47 are never written again: zero the whole buffer once
  for (int e = tid; e < (int)(sizeof(sm.B) / 4); e += THREADS) reinterpret_cast<float*>(sm.B)[e] = 0.0f;
  if (warp == W_CTRL) {
    tmem_alloc(&sm.tmem_base, TMEM_COLS);
    tmem_relinquish();
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = sm.tmem_base;

  if (nunits > 0) {
    if (warp < NFE) {
      // ======================================================================= front-end warps
      ItemHead& I = sm.item[warp];
      int it = 0;
      bool have_item = false;
      uint32_t item_lo = 0, item_hi = 0;
      uint32_t last_x0 = 0xffffffffu;
      float xn = 0.f;
#ifdef DFK_TC_TIMERS
      Tmr tm{0, dbg != 0 && lane == 0};
#else
      Tmr tm{0, false};
#endif
      unsigned long long t_head = 0, t_geo = 0, t_gather = 0, t_wait = 0, t_write = 0;
      for (int b = warp; b < nblk; b += NFE) {
        tm.start();
        const uint32_t g = (uint32_t)(g_lo + b);
        if (!have_item || g >= item_hi) {
          while (g >= items[it].tile_begin + items[it].num_tiles) ++it;
          __syncwarp();
          load_item_head(I, items[it], lane, cta);
          __syncwarp();
          have_item = true;
          item_lo = I.blk_begin;
          item_hi = I.blk_begin + I.num_blocks;
          last_x0 = 0xffffffffu;
        }
        uint32_t x0, y;
        const bool real = decode_block(I, g - item_lo, x0, y);
        const int u = b >> 1;
        const int fu = u % NFU;
        const int fs = 2 * fu + (b & 1);
        const uint32_t n = real ? min((uint32_t)BLK, I.width - x0) : 0u;
        const bool bulk = (I.flags & ITEM_FLAG_BULK) != 0;
        const float* jrow = I.jac + (size_t)y * I.jac_pitch + (size_t)x0 * C;
        tm.lap(t_head);

        float feat[NFEAT];
#pragma unroll
        for (int j = 0; j < NFEAT; ++j) feat[j] = 0.0f;
        unsigned bal = 0u;
        if (real) {
          const bool in = (uint32_t)lane < n;
          const uint32_t x = x0 + (in ? (uint32_t)lane : 0u);
          float d = __ldcs(I.dpt0 + (size_t)y * I.dpt0_pitch + x);
          const float i0 = __ldcs(I.img0 + (size_t)y * I.img0_pitch + x);
          if (x0 != last_x0) {  // Reproject's (x - u0) / fx depends on the column only: once per strip
            xn = ray_coord((float)x, I.u0, I.fx);
            last_x0 = x0;
          }
          const float yn = ray_coord((float)y, I.v0, I.fy);
          if (I.flags & ITEM_FLAG_FUSED_DEPTH) {
            // dpt0 points at prx_orig: decode the depth from this pixel's code-Jacobian row with the arithmetic of
            // update_depth_kernel (chunk fma chains + xor-butterfly over the chunk sums), publish it, carry on with it
            const float* rowf = jrow + (size_t)(in ? lane : 0) * C;
            const float4* cod = reinterpret_cast<const float4*>(I.code);  // device scratch, 128-byte aligned per item
            float part[C / 4];
            if (bulk) {  // 16-byte aligned rows
              const float4* row = reinterpret_cast<const float4*>(rowf);
#pragma unroll
              for (int k4 = 0; k4 < C / 4; ++k4) part[k4] = chunk_dot(__ldg(row + k4), __ldg(cod + k4));
            } else {
#pragma unroll
              for (int k4 = 0; k4 < C / 4; ++k4)
                part[k4] = chunk_dot(make_float4(__ldg(rowf + 4 * k4), __ldg(rowf + 4 * k4 + 1), __ldg(rowf + 4 * k4 + 2),
                                                 __ldg(rowf + 4 * k4 + 3)),
                                     __ldg(cod + k4));
            }
            d = prx_to_depth(__fadd_rn(d, butterfly_sum<C / 4>(part)), I.avg_dpt);
            if (in) I.dpt_out[(size_t)y * I.dpt_out_pitch + x] = d;
          }
          const Warped w = warp_ray(xn, yn, d, I.q, I.t, I.fx, I.fy, I.u0, I.v0, I.border, I.ulim, I.vlim, I.min_dpt);
          const bool ok = in && w.valid;
          bal = __ballot_sync(0xffffffffu, ok);
          tm.lap(t_geo);
#ifdef DFK_EXP_NOGATHER
          if (ok) {
            I.valid0[(size_t)y * I.valid0_pitch + x] = 1.0f;
#pragma unroll
            for (int j = 0; j < 8; ++j) feat[j] = 0.001f * (float)(j + 1) * d + w.u * 1e-4f + i0;
            feat[8] = 1.0f;
          }
          if (false) {
#else
          if (ok) {
#endif
            I.valid0[(size_t)y * I.valid0_pitch + x] = 1.0f;  // dense_sfm.h:161
            int ix, iy;
            float fu_, fv_, gx, gy;
            bilin_setup(w.u, w.v, ix, iy, fu_, fv_);
            sample_grad(I.grad1, I.grad1_pitch, true, ix, iy, fu_, fv_, gx, gy);  // the API guarantees 8-byte rows here
            const float i1 = sample_scalar(I.img1, I.img1_pitch, ix, iy, fu_, fv_);
            float a[6], c00, c02, c11, c12;
            pose_jacobian_row(w, I.fx, I.fy, gx, gy, a, c00, c02, c11, c12);
            const float e = prx_jacobian(w, I.R, d, I.avg_dpt, gx, gy, c00, c02, c11, c12);
            const float diff = i0 - i1;
            const float hw = huber_weight(diff, I.huber_delta);
            feat[0] = hw * e;
#pragma unroll
            for (int j = 0; j < 6; ++j) feat[1 + j] = hw * a[j];
            feat[7] = hw * diff;
            feat[8] = 1.0f;  // inliers += 1 (dense_sfm.h:196) as a Gram entry
          }
          tm.lap(t_gather);
        }
        const int nv = __popc(bal);
        // the slot's previous unit (u - NFU) must have been consumed by its trio
        if (lane == 0 && u >= NFU) mbar_wait(&sm.feat_empty[fu], ((uint32_t)(u / NFU) - 1u) & 1u);
        __syncwarp();
        tm.lap(t_wait);
        if (nv > 0) {
#pragma unroll
          for (int f = 0; f < NFEAT; ++f) sm.feat[fs][f][lane] = feat[f];
        }
        if (lane == 0) {
          BlockMeta m;
          m.nv = nv;
          m.n = (int)n;
          m.bulk = bulk ? 1 : 0;
          m.pad = 0;
          m.jrow = jrow;
          m.pad2 = 0;
          sm.meta[fs] = m;
        }
        __syncwarp();  // the warp's feat entries / meta are ordered before lane 0's release
        if (lane == 0) mbar_arrive(&sm.feat_full[fu]);
        tm.lap(t_write);
      }
      if (tm.on) {
        atomicAdd(&g_dbg[0], t_head); atomicAdd(&g_dbg[1], t_geo); atomicAdd(&g_dbg[2], t_gather);
        atomicAdd(&g_dbg[3], t_wait); atomicAdd(&g_dbg[4], t_write);
      }
    } else if (warp == W_PROD) {
      // ======================================================================= bulk-copy producer (one thread)
      if (lane == 0) {
        ItemHead& I = sm.item[NFE];
        int it = 0;
        bool have_item = false;
        uint32_t item_lo = 0, item_hi = 0;
        for (int u = 0; u < nunits; ++u) {
          const int ru = u % NRU;
          if (u >= NRU) mbar_wait(&sm.stage_empty[ru], ((uint32_t)(u / NRU) - 1u) & 1u);
#pragma unroll 1
          for (int j = 0; j < 2; ++j) {
            const uint32_t g = (uint32_t)(g_lo + 2 * u + j);
            if (!have_item || g >= item_hi) {
              while (g >= items[it].tile_begin + items[it].num_tiles) ++it;
              const SfmItemDev& S = items[it];
              I.jac = S.jac; I.jac_pitch = S.jac_pitch; I.width = S.width; I.height = S.height; I.flags = S.flags;
              I.perm_mul = S.perm_mul; I.ph = S.tc_ph; I.ngroups = S.tc_ngroups; I.npatches = S.tc_npatches;
              I.mag_ph = S.tc_mag_ph; I.mag_np = S.tc_mag_np; I.mag_ng = S.tc_mag_ng;
              item_lo = S.tile_begin;
              item_hi = S.tile_begin + S.num_tiles;
              have_item = true;
            }
            uint32_t x0, y;
            const bool real = decode_block(I, g - item_lo, x0, y);
            if (real && (I.flags & ITEM_FLAG_BULK)) {
              const uint32_t n = min((uint32_t)BLK, I.width - x0);
              mbar_arrive_expect_tx(&sm.tma_full[ru], n * (uint32_t)(C * 4));
              bulk_g2s(&sm.jc[2 * ru + j][0], I.jac + (size_t)y * I.jac_pitch + (size_t)x0 * C, n * (uint32_t)(C * 4),
                       &sm.tma_full[ru]);
            } else {
              mbar_arrive(&sm.tma_full[ru]);  // keeps the phase of the ring slot in step with the unit count
            }
          }
        }
      }
    } else if (warp == W_CTRL) {
      // ======================================================================= control warp (one thread)
      if (lane == 0) {
        const uint32_t idesc = make_idesc_tf32(MM, NB);
        SeqWalker seq(items, gu_lo);
        bool first = true;
        bool had_mma = false;
#ifdef DFK_TC_TIMERS
        Tmr tm{0, dbg != 0};
#else
        Tmr tm{0, false};
#endif
        unsigned long long t_ab = 0, t_issue = 0, t_dempty = 0;
        for (int u = 0; u < nunits; ++u) {
          const int t = u & 1;
          tm.start();
          const int e_prev = seq.e;
          if (seq.step(u)) {
            if (u > 0) {
              sm.chain_nz[e_prev & 1] = had_mma ? 1u : 0u;
              __threadfence_block();
              umma_commit(&sm.d_full[e_prev & 1]);
            }
            first = true;
            had_mma = false;
            const int use = seq.e >> 1;  // n-th use of this accumulator buffer
            if (use >= 1) {
              mbar_wait(&sm.d_empty[seq.e & 1], (uint32_t)(use - 1) & 1u);
              tc_fence_after();
            }
            tm.lap(t_dempty);
          }
          mbar_wait(&sm.ab_full[t], (uint32_t)(u >> 1) & 1u);
          tc_fence_after();
          const uint32_t flags = sm.umeta[t];
          tm.lap(t_ab);
          const uint32_t d_addr = tbase + D_COL + NB * (seq.e & 1);
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            if (flags & (1u << j)) {
              const uint32_t a_addr = tbase + A_COL + 64u * (uint32_t)t + 32u * (uint32_t)j;
              const uint64_t bdesc0 = make_smem_desc_kmajor_noswizzle(smem_u32(sm.B[t][j]), 128, B_SBO);
#pragma unroll
              for (int ks = 0; ks < BLK / 8; ++ks) {  // 8-pixel k-steps: two 128-byte core-matrix columns of B each
#ifndef DFK_EXP_NOOP
                umma_tf32_ts(d_addr, a_addr + 8u * ks, bdesc0 + (uint64_t)((ks * 256) >> 4), idesc, !first);
#endif
                first = false;
              }
              had_mma = true;
            }
          }
          umma_commit(&sm.ab_empty[t]);
          tm.lap(t_issue);
        }
        sm.chain_nz[seq.e & 1] = had_mma ? 1u : 0u;
        __threadfence_block();
        umma_commit(&sm.d_full[seq.e & 1]);
        if (tm.on) {
          atomicAdd(&g_dbg[7], t_ab); atomicAdd(&g_dbg[8], t_issue); atomicAdd(&g_dbg[9], t_dempty);
        }
      }
    } else {
      // ======================================================================= operand trios (A: 12-14, B: 8-10)
      const int trio = (warp >= W_TRIO_A) ? 0 : 1;             // trio A builds the even units, trio B the odd ones
      const int ow = warp - (trio == 0 ? W_TRIO_A : W_TRIO_B);  // 0: h, 1: l, 2: p  == TMEM lane quarter
      const uint32_t lane_taddr = tbase + ((uint32_t)(ow * 32) << 16);
      const int row = ow * 32 + lane;  // TMEM lane == row of the partial
      SeqWalker seq(items, gu_lo);     // trio A: chain bookkeeping for the drains
      int drained = 0;                 // chains drained so far (trio A)
      // descriptors of the chains that started but are not drained yet (the walker is at most a few chains ahead)
      int ch_slot[8];
      bool ch_fresh[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) { ch_slot[k] = 0; ch_fresh[k] = false; }

      // Move a finished chain TMEM -> the CTA's partial in global memory (single writer, fixed order).
      // fresh: first chain of the item in this CTA (store), else fire-and-forget red.global.add.f32 in program order.
      auto drain = [&](int e) {
        const int bb = e & 1, use = e >> 1;
        const int slot = ch_slot[e & 7];
        const bool fresh = ch_fresh[e & 7];
        float* P = partials + (size_t)slot * kTcPartialFloats;
        mbar_wait(&sm.d_full[bb], (uint32_t)use & 1u);
        tc_fence_after();
        const bool nz = sm.chain_nz[bb] != 0u;
#ifdef DFK_EXP_NOOP
        if (fresh) {
#else
        if (nz || fresh) {
#endif
#pragma unroll 1
          for (int pass = 0; pass < 3; ++pass) {
            const int nq = pass < 2 ? 4 : (kTcCols - 32) / 4;  // float4 per pass (columns 40..47 are padding)
            uint32_t v[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = 0u;
#ifndef DFK_EXP_NOOP
            if (nz) {
              tmem_ld_x16(lane_taddr + D_COL + NB * bb + 16 * pass, v);
              tmem_wait_ld();
            }
#endif
            // column-major partial: this lane's row at column j is P[j * kTcRowsPad + row]
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
        if (lane == 0) mbar_arrive(&sm.d_empty[bb]);
      };

#ifdef DFK_TC_TIMERS
      Tmr tm{0, dbg != 0 && trio == 0 && ow == 0 && lane == 0};
#else
      Tmr tm{0, false};
#endif
      unsigned long long t_ffull = 0, t_abempty = 0, t_tma = 0, t_build = 0, t_sync = 0, t_drain = 0, t_total = 0;
#ifdef DFK_TC_TIMERS
      const long long t_begin = tm.on ? clock64() : 0;
#endif
      int walked = 0;  // trio A: units the chain walker has passed
      for (int u = trio; u < nunits; u += 2) {
        const int ru = u % NRU;
        const int fu = u % NFU;
        tm.start();
        mbar_wait(&sm.feat_full[fu], (uint32_t)(u / NFU) & 1u);
        tm.lap(t_ffull);
        const BlockMeta m0 = sm.meta[2 * fu], m1 = sm.meta[2 * fu + 1];
        const uint32_t flags = (m0.nv > 0 ? 1u : 0u) | (m1.nv > 0 ? 2u : 0u);
        const bool any_bulk = (m0.bulk | m1.bulk) != 0;
        // The trio's A columns / B rows / umeta word were last read by the control thread for unit u - 2.  Waited for
        // even when this unit is empty: the trio must never run two ab_full phases ahead of the control thread.
        if (u >= 2) mbar_wait(&sm.ab_empty[trio], ((uint32_t)(u >> 1) - 1u) & 1u);
        tc_fence_after();
        tm.lap(t_abempty);
        if (ow < 2 && any_bulk) mbar_wait(&sm.tma_full[ru], (uint32_t)(u / NRU) & 1u);  // also for an empty unit: its copies must land before the ring slot is handed back
        tm.lap(t_tma);
#ifndef DFK_EXP_NOOP
#pragma unroll 1
        for (int j = 0; j < 2; ++j) {
          if (!(flags & (1u << j))) continue;
          const BlockMeta& meta = j ? m1 : m0;
          const int fs = 2 * fu + j;
          const uint32_t a_taddr = lane_taddr + A_COL + 64u * (uint32_t)trio + 32u * (uint32_t)j;
          unsigned char* bslot = sm.B[trio][j];
          if (ow < 2) {
            const float4* s4p = reinterpret_cast<const float4*>(sm.feat[fs][0]);
            const float* jcs = sm.jc[2 * ru + j] + lane;        // raw rows: jcs[p * C]
            const float* jcg = meta.jrow + lane;                // non-bulk items: straight from global memory
            float4* brow = reinterpret_cast<float4*>(bslot + (uint32_t)(lane >> 3) * B_SBO + (uint32_t)(lane & 7) * 16u);
#pragma unroll
            for (int half = 0; half < 2; ++half) {
              float4 s4[4];
#pragma unroll
              for (int q = 0; q < 4; ++q) s4[q] = s4p[4 * half + q];  // broadcast reads of s
              float val[16];
              if (meta.bulk) {
#pragma unroll
                for (int p = 0; p < 16; ++p) val[p] = jcs[(16 * half + p) * C];
              } else {
#pragma unroll
                for (int p = 0; p < 16; ++p) val[p] = (16 * half + p) < meta.n ? __ldg(jcg + (16 * half + p) * C) : 0.0f;
              }
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const float sc[4] = {s4[q].x, s4[q].y, s4[q].z, s4[q].w};
#pragma unroll
                for (int c4 = 0; c4 < 4; ++c4)
                  // s == 0 (invalid pixel, or one past the image edge whose ring row is stale): exactly zero, whatever
                  // the row holds
                  val[4 * q + c4] = sc[c4] != 0.0f ? sc[c4] * val[4 * q + c4] : 0.0f;
              }
              uint32_t v[16];
              if (ow == 0) {
                // B rows = features (this lane), k-chunks of 4 pixels, 128 B apart
#pragma unroll
                for (int q = 0; q < 4; ++q)
                  brow[8 * (4 * half + q)] = make_float4(val[4 * q], val[4 * q + 1], val[4 * q + 2], val[4 * q + 3]);
#pragma unroll
                for (int p = 0; p < 16; ++p) v[p] = __float_as_uint(val[p]);
              } else {
#pragma unroll
                for (int p = 0; p < 16; ++p) v[p] = __float_as_uint(val[p] - tf32_trunc(val[p]));
              }
              tmem_st_x16(a_taddr + 16u * half, v);
            }
          } else {
            // pose / residual / count features: lanes 0-7 = h of feature 1+lane, lanes 8-15 = l of feature 1+(lane-8)
            const int f = 1 + (lane & 7);
            const float4* fp = reinterpret_cast<const float4*>(sm.feat[fs][f]);
            float4 x[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) x[q] = fp[q];
            if (lane < 8) {
              const uint32_t brow_i = 32u + (uint32_t)lane;
              float4* brow = reinterpret_cast<float4*>(bslot + (brow_i >> 3) * B_SBO + (brow_i & 7u) * 16u);
#pragma unroll
              for (int q = 0; q < 8; ++q) brow[8 * q] = x[q];
            }
            uint32_t v[32];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              float e0 = x[q].x, e1 = x[q].y, e2 = x[q].z, e3 = x[q].w;
              if (lane >= 8) {
                e0 -= tf32_trunc(e0); e1 -= tf32_trunc(e1); e2 -= tf32_trunc(e2); e3 -= tf32_trunc(e3);
              }
              if (lane >= 16) { e0 = 0.f; e1 = 0.f; e2 = 0.f; e3 = 0.f; }
              v[4 * q] = __float_as_uint(e0); v[4 * q + 1] = __float_as_uint(e1);
              v[4 * q + 2] = __float_as_uint(e2); v[4 * q + 3] = __float_as_uint(e3);
            }
            tmem_st_x32(a_taddr, v);
          }
        }
#endif
        tm.lap(t_build);
        if (flags) {
          tmem_wait_st();
          if (ow != 1) fence_proxy_async_smem();  // the l warp wrote TMEM only, no B rows
          tc_fence_before();
        }
        if (ow == 0 && lane == 0) sm.umeta[trio] = flags;  // read by the control thread after the trio's three arrivals
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&sm.ab_full[trio]);
          mbar_arrive(&sm.feat_empty[fu]);
          if (ow < 2) mbar_arrive(&sm.stage_empty[ru]);
        }
        tm.lap(t_sync);
        if (trio == 0) {
          // chains that ended before the unit after this one: walk the sequence up to u + 1 (so that a chain starting at the
          // other trio's next unit is seen as early as possible), drain everything that ended
          const int upto = min(u + 1, nunits - 1);
          for (; walked <= upto; ++walked) {
            if (seq.step(walked)) {
              ch_slot[seq.e & 7] = (int)(items[seq.it].partial_begin + (uint32_t)cta - items[seq.it].first_cta);
              ch_fresh[seq.e & 7] = seq.chain_is_first_of_item();
            }
          }
          // chain seq.e contains unit `upto`; every chain before it has ended.  The last chain that ended may still be
          // waiting for the control thread to reach unit `upto`: if that is the other trio's unit, leave it for the next
          // round instead of stalling this trio behind it.
          const int ended = seq.e;  // chains [drained, ended) are complete
          const int safe = (upto > u) ? ((seq.chain_first == upto) ? ended - 1 : ended) : ended;
          while (drained < safe) drain(drained++);
        }
        tm.lap(t_drain);
      }
      if (trio == 0) {
        for (; walked < nunits; ++walked) {
          if (seq.step(walked)) {
            ch_slot[seq.e & 7] = (int)(items[seq.it].partial_begin + (uint32_t)cta - items[seq.it].first_cta);
            ch_fresh[seq.e & 7] = seq.chain_is_first_of_item();
          }
        }
        while (drained <= seq.e) drain(drained++);
      }
      if (tm.on) {
#ifdef DFK_TC_TIMERS
        t_total = (unsigned long long)(clock64() - t_begin);
#endif
        atomicAdd(&g_dbg[10], t_ffull); atomicAdd(&g_dbg[11], t_abempty); atomicAdd(&g_dbg[12], t_tma);
        atomicAdd(&g_dbg[13], t_build); atomicAdd(&g_dbg[14], t_sync); atomicAdd(&g_dbg[15], t_drain);
        atomicAdd(&g_dbg[16], t_total); atomicAdd(&g_dbg[17], (unsigned long long)nblk);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == W_CTRL) tmem_dealloc(tbase, TMEM_COLS);
}

}  // namespace

bool sfm_tc_supported(int code_size) { return code_size == 32; }

size_t sfm_tc_smem_bytes() { return sizeof(Smem); }

cudaError_t launch_sfm_tc(const SfmItemDev* items_dev, const SfmLaunchPlan& plan, float* partials_dev,
                          cudaStream_t stream, cudaEvent_t ev_start, cudaEvent_t ev_stop)
{
  const size_t smem = sizeof(Smem);
  static const cudaError_t attr_err =
      cudaFuncSetAttribute(sfm_step_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem));
  if (attr_err != cudaSuccess) return attr_err;
  static const int dbg = []() { const char* e = getenv("DFK_TC_DEBUG"); return (e && e[0] == '1') ? 1 : 0; }();
  if (dbg) {
    unsigned long long z[32] = {0};
    cudaMemcpyToSymbolAsync(g_dbg, z, sizeof(z), 0, cudaMemcpyHostToDevice, stream);
  }
  if (ev_start) cudaEventRecord(ev_start, stream);
  sfm_step_tc_kernel<<<plan.num_ctas, THREADS, smem, stream>>>(items_dev, plan.num_items, plan.num_tiles / 2, partials_dev,
                                                              dbg);
  if (ev_stop) cudaEventRecord(ev_stop, stream);
  if (dbg) {
    unsigned long long v[32];
    cudaStreamSynchronize(stream);
    cudaMemcpyFromSymbol(v, g_dbg, sizeof(v));
    const double nb = v[17] ? (double)v[17] : 1.0;  // blocks (summed over CTAs)
    fprintf(stderr,
            "[dfk tc dbg] ctas=%d blocks=%d | cycles per block: FE(lane 0 of 8 warps) head %.0f geom %.0f gather %.0f "
            "slot_wait %.0f write %.0f | CTRL ab_wait %.0f issue %.0f d_empty %.0f | trio A h: feat_wait %.0f ab_empty %.0f "
            "tma %.0f build %.0f sync %.0f drain %.0f total %.0f\\n",
            plan.num_ctas, plan.num_tiles, v[0] / nb, v[1] / nb, v[2] / nb, v[3] / nb, v[4] / nb, v[7] / nb, v[8] / nb,
            v[9] / nb, v[10] / nb, v[11] / nb, v[12] / nb, v[13] / nb, v[14] / nb, v[15] / nb, v[16] / nb);
  }
  return cudaGetLastError();
}

}  // namespace dfk
