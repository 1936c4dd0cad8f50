// dfk_sfm_tc.cu -- SfmAligner::RunStep hot path, tcgen05 tensor-core Gram variant (sm_100a, C = 32).  Round-2 design.
//
// Replaces kernel_step_calculate + DenseSfm + the two-kernel reduction of the reference
// (sources/cuda/cu_sfmaligner.cpp:40-70,149-185, sources/common/algorithm/dense_sfm.h:133-201).  The reduced Gram
//     G = sum_p m_p^T m_p,   m = w*[ e*jc (32) | a (6) | diff (1) ]        (39 features)
// runs on the 5th-generation tensor cores with split precision folded into ONE MMA: every feature value v is split
// exactly into h = the bits kind::tf32 keeps (a truncation of the low 13 mantissa bits on this hardware, measured by
// tools/umma_probe.cu) and l = v - h; A = [h rows ; l rows] (78 of M = 128 rows, in TMEM), B = h (39 of N = 48 columns,
// K-major in shared memory); one tcgen05.mma.kind::tf32 per 8 pixels gives HH = sum h h^T and LH = sum l h^T and
// G = HH + LH + LH^T drops only the l*l terms (~2^-22).  The finalize kernel (dfk_sfm_finalize.cu) recombines.
//
// Work decomposition: a BLOCK is 32 pixels of one image row (lane = pixel), a PATCH is a 32-pixel-wide strip x ~32
// rows; the patches of an item are visited in a golden-ratio permuted order (balances clustered invalid regions over
// the CTAs), the rows of a patch consecutively (neighbouring rows share their bilinear taps in L1).  CTA c owns a
// contiguous range of the global block sequence (static => bitwise reproducible results).
//
// Roles per CTA (512 threads, 2 CTAs / SM, 256 TMEM columns each):
//   warps 0-11  front-end : each warp processes whole blocks on its own (block b -> warp b % 12), one thread per
//                           pixel: coalesced loads of dpt0 / img0, exact-order validity chain, bilinear gathers of
//                           img1 / grad1, Jacobian row, Huber weight -> eight scalars per pixel (s = w*e, w*a[6],
//                           w*diff) into a feat slot in shared memory.  No compaction, no touching of the code
//                           Jacobian.  Lane 0 also issues the block's cp.async.bulk (TMA engine, SASS UBLKCP) of the
//                           32 x 128-byte code-Jacobian row segment into a 16-deep ring, ~12 blocks ahead of its use.
//   warp 12/13  operand h/l: lane = code feature.  Per block 32 conflict-free LDS of the raw rows, v = s * jc, and
//                           tcgen05.st (registers -> TMEM lanes 0-31: v, the tensor core truncates it to h; lanes
//                           32-63: l = v - trunc(v)); the h warp also writes v K-major to shared memory as B.
//   warp 14     operand p : h and l of the 7 pose/residual features (TMEM lanes 64-77) + their B rows.
//   warp 15     control   : lane 0 issues 4 tcgen05.mma (M128 N48 K8, A from TMEM, B from shared memory) per block and
//                           the tcgen05.commit arrivals; allocates TMEM.
// The fp32 accumulator in TMEM adds with truncation (measured ~ -2^-24 relative per k-step), so an accumulation chain
// is cut every 32 non-empty blocks: the operand warps pull the finished chain out of TMEM (tcgen05.ld) and add it in
// round-to-nearest fp32 to the CTA's partial in global memory (single writer per address, program order).
// Blocks without a valid pixel (outside the image overlap) cost a validity test and nothing else.
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
constexpr int NFE = 12;              // front-end warps (warps 0..11)
constexpr int NST = 16;              // code-Jacobian ring stages (one block each)
constexpr int NFS = 16;              // feat slots
constexpr int NAB = 4;               // operand slots: A (TMEM, 32 columns each) / B (shared memory)
constexpr int THREADS = 512;
constexpr int W_OPH = 12, W_CTRL = 15;  // warps 12, 13, 14: operand h, l, pose (TMEM lane quarters 0, 1, 2)
constexpr int NB = 48;               // MMA N (39 used)
constexpr int MM = 128;              // MMA M (78 used)
#ifndef DFK_FLUSH_BLOCKS
#define DFK_FLUSH_BLOCKS 32
#endif
constexpr int kFlushBlocks = DFK_FLUSH_BLOCKS;  // TMEM accumulation chain length (non-empty blocks)
constexpr uint32_t TMEM_COLS = 256;
constexpr uint32_t A_COL = 0;                 // [0, 32*NAB)
constexpr uint32_t D_COL = 32 * NAB;          // two accumulators of NB columns
constexpr uint32_t B_SBO = (BLK / 4) * 128;   // 1024 B between 8-row groups
constexpr uint32_t B_SLOT_BYTES = (NB / 8) * B_SBO;  // 6144 B
constexpr int FEAT_STRIDE = BLK + 4;          // floats per feature row of a feat slot: the pose operand warp's lanes read
                                              // the same pixel chunk of 7 rows -> 7 different bank groups
static_assert(D_COL + 2 * NB <= TMEM_COLS, "TMEM budget");

struct BlockMeta {
  int nv;            // valid pixels of the block (0: nothing to build / multiply)
  int item_changed;  // the block sequence enters a new item here
  int pslot;         // partial slot of (item, CTA)
  int n;             // pixels of the block inside the image (1..32)
  const float* jrow; // global address of the block's code-Jacobian row segment (non-bulk items read it directly)
  int bulk;
  int pad;
};

// what a front-end warp keeps of its current item (shared memory, one copy per warp)
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
  alignas(128) float jc[NST][BLK * C];            // 64 KB
  alignas(128) unsigned char B[NAB][B_SLOT_BYTES]; // 24 KB
  alignas(16) float feat[NFS][8][FEAT_STRIDE];    // s, wa0..5, wr  (18 KB)
  alignas(16) BlockMeta meta[NFS];
  alignas(16) ItemHead item[NFE];
  alignas(8) uint64_t tma_full[NST];
  uint64_t stage_empty[NST];
  uint64_t feat_full[NFS];
  uint64_t feat_empty[NFS];
  uint64_t ab_full[NAB];
  uint64_t ab_empty[NAB];
  uint64_t d_full[2];
  uint64_t d_empty[2];
  uint32_t tmem_base;
};

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

// ---- optional phase timers (clock64 sums per role), enabled with the env var DFK_TC_DEBUG=1 ----------
__device__ unsigned long long g_dbg[24];
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
  int e = -1;               // current chain index
  int blocks_in_chain = 0;  // non-empty blocks
  __device__ __forceinline__ bool starts_chain(int b, int item_changed) const
  {
    return b == 0 || item_changed != 0 || blocks_in_chain >= kFlushBlocks;
  }
};

__global__ void __launch_bounds__(THREADS, 2)
sfm_step_tc_kernel(const SfmItemDev* __restrict__ items, int num_items, int num_blocks, float* __restrict__ partials,
                   int dbg)
{
  extern __shared__ __align__(128) unsigned char smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const int cta = blockIdx.x;
  const int G = gridDim.x;
  const int g_lo = (int)(((long long)cta * num_blocks) / G);
  const int g_hi = (int)(((long long)(cta + 1) * num_blocks) / G);
  const int nblk = g_hi - g_lo;
  (void)num_items;

  // ---- one-time setup ---------------------------------------------------------------------------
  if (tid == 0) {
    for (int s = 0; s < NST; ++s) {
      mbar_init(&sm.tma_full[s], 1);
      mbar_init(&sm.stage_empty[s], 2);  // operand warps h and l
    }
    for (int s = 0; s < NFS; ++s) {
      mbar_init(&sm.feat_full[s], 1);
      mbar_init(&sm.feat_empty[s], 4);  // three operand warps + the control thread
    }
    for (int s = 0; s < NAB; ++s) {
      mbar_init(&sm.ab_full[s], 3);
      mbar_init(&sm.ab_empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&sm.d_full[b], 1);
      mbar_init(&sm.d_empty[b], 3);
    }
    mbar_fence_init();
  }
  // B rows 39..47 are never written again: zero the whole buffer once
  for (int e = tid; e < (int)(NAB * B_SLOT_BYTES / 4); e += THREADS) reinterpret_cast<float*>(sm.B)[e] = 0.0f;
  if (warp == W_CTRL) {
    tmem_alloc(&sm.tmem_base, TMEM_COLS);
    tmem_relinquish();
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = sm.tmem_base;

  if (nblk > 0) {
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
      unsigned long long t_head = 0, t_geo = 0, t_gather = 0, t_wait = 0, t_write = 0, t_tmaw = 0;
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
        const bool seq_changed = (b == 0) || (g == item_lo);
        // ---- block -> (x0, y) ---------------------------------------------------------------------
        const uint32_t k = g - item_lo;
        uint32_t r, pid, grp;
        const uint32_t qd = div_magic(k, I.ph, I.mag_ph, r);
        div_magic(qd * I.perm_mul, I.npatches, I.mag_np, pid);  // host guarantees qd * perm_mul < 2^32
        const uint32_t strip = div_magic(pid, I.ngroups, I.mag_ng, grp);
        const uint32_t x0 = strip * BLK;
        const uint32_t y = grp * I.ph + r;
        const bool null = y >= I.height;
        const int st = b % NST;
        const int fs = b % NFS;
        const uint32_t W = I.width;
        const uint32_t n = null ? 0u : min((uint32_t)BLK, W - x0);
        const bool bulk = (I.flags & ITEM_FLAG_BULK) != 0;
        const float* jrow = I.jac + (size_t)y * I.jac_pitch + (size_t)x0 * C;
        // ---- the block's code-Jacobian rows: one bulk copy into ring stage st, consumed ~NFE blocks later ------------
        bool issued = true;
        if (lane == 0) {
          const bool free_now = (b < NST) || mbar_try_wait(&sm.stage_empty[st], ((uint32_t)(b / NST) - 1u) & 1u);
          if (free_now) {
            if (bulk && !null) {
              mbar_arrive_expect_tx(&sm.tma_full[st], n * (uint32_t)(C * 4));
              bulk_g2s(&sm.jc[st][0], jrow, n * (uint32_t)(C * 4), &sm.tma_full[st]);
            } else {
              mbar_arrive(&sm.tma_full[st]);  // keeps the phase of the stage in step with the block count
            }
          } else {
            issued = false;
          }
        }
        tm.lap(t_head);

        float feat[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) feat[j] = 0.0f;
        unsigned bal = 0u;
        if (!null) {
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
          if (ok) {
            I.valid0[(size_t)y * I.valid0_pitch + x] = 1.0f;  // dense_sfm.h:161
            int ix, iy;
            float fu, fv, gx, gy;
            bilin_setup(w.u, w.v, ix, iy, fu, fv);
            sample_grad(I.grad1, I.grad1_pitch, true, ix, iy, fu, fv, gx, gy);  // the API guarantees 8-byte rows here
            const float i1 = sample_scalar(I.img1, I.img1_pitch, ix, iy, fu, fv);
            float a[6], c00, c02, c11, c12;
            pose_jacobian_row(w, I.fx, I.fy, gx, gy, a, c00, c02, c11, c12);
            const float e = prx_jacobian(w, I.R, d, I.avg_dpt, gx, gy, c00, c02, c11, c12);
            const float diff = i0 - i1;
            const float hw = huber_weight(diff, I.huber_delta);
            feat[0] = hw * e;
#pragma unroll
            for (int j = 0; j < 6; ++j) feat[1 + j] = hw * a[j];
            feat[7] = hw * diff;
          }
          tm.lap(t_gather);
        }
        const int nv = __popc(bal);
        // the slot's previous block (b - NFS) must have been consumed by the operand warps and the control thread
        if (lane == 0) {
          if (!issued) {
            mbar_wait(&sm.stage_empty[st], ((uint32_t)(b / NST) - 1u) & 1u);
            if (bulk && !null) {
              mbar_arrive_expect_tx(&sm.tma_full[st], n * (uint32_t)(C * 4));
              bulk_g2s(&sm.jc[st][0], jrow, n * (uint32_t)(C * 4), &sm.tma_full[st]);
            } else {
              mbar_arrive(&sm.tma_full[st]);
            }
          }
          tm.lap(t_tmaw);
          if (b >= NFS) mbar_wait(&sm.feat_empty[fs], ((uint32_t)(b / NFS) - 1u) & 1u);
        }
        __syncwarp();
        tm.lap(t_wait);
        if (nv > 0) {
#pragma unroll
          for (int f = 0; f < 8; ++f) sm.feat[fs][f][lane] = feat[f];
        }
        if (lane == 0) {
          BlockMeta m;
          m.nv = nv;
          m.item_changed = seq_changed ? 1 : 0;
          m.pslot = (int)I.pslot;
          m.n = (int)n;
          m.jrow = jrow;
          m.bulk = bulk ? 1 : 0;
          m.pad = 0;
          sm.meta[fs] = m;
        }
        __syncwarp();  // the warp's feat entries / meta are ordered before lane 0's release
        if (lane == 0) mbar_arrive(&sm.feat_full[fs]);
        tm.lap(t_write);
      }
      if (tm.on) {
        atomicAdd(&g_dbg[0], t_head); atomicAdd(&g_dbg[1], t_geo); atomicAdd(&g_dbg[2], t_gather);
        atomicAdd(&g_dbg[3], t_wait); atomicAdd(&g_dbg[4], t_write); atomicAdd(&g_dbg[5], t_tmaw);
      }
    } else if (warp == W_CTRL) {
      // ======================================================================= control warp (one thread)
      if (lane == 0) {
        const uint32_t idesc = make_idesc_tf32(MM, NB);
        ChainState ch;
        bool first = true;
        int na = 0;  // non-empty blocks so far
#ifdef DFK_TC_TIMERS
        Tmr tm{0, dbg != 0};
#else
        Tmr tm{0, false};
#endif
        unsigned long long t_ff = 0, t_ab = 0, t_issue = 0, t_dempty = 0;
        for (int b = 0; b < nblk; ++b) {
          const int fs = b % NFS;
          tm.start();
          mbar_wait(&sm.feat_full[fs], (uint32_t)(b / NFS) & 1u);
          const int nv = sm.meta[fs].nv;
          const int item_changed = sm.meta[fs].item_changed;
          mbar_arrive(&sm.feat_empty[fs]);
          tm.lap(t_ff);
          if (ch.starts_chain(b, item_changed)) {
            if (b > 0) umma_commit(&sm.d_full[ch.e & 1]);
            ch.e += 1;
            ch.blocks_in_chain = 0;
            first = true;
            const int use = ch.e >> 1;  // n-th use of this accumulator buffer
            if (use >= 1) {
              mbar_wait(&sm.d_empty[ch.e & 1], (uint32_t)(use - 1) & 1u);
              tc_fence_after();
            }
            tm.lap(t_dempty);
          }
          if (nv > 0) {
            ch.blocks_in_chain += 1;
            const int ab = na % NAB;
            mbar_wait(&sm.ab_full[ab], (uint32_t)(na / NAB) & 1u);
            tc_fence_after();
            tm.lap(t_ab);
            const uint32_t d_addr = tbase + D_COL + NB * (ch.e & 1);
            const uint32_t a_addr = tbase + A_COL + 32u * (uint32_t)ab;
            const uint64_t bdesc0 = make_smem_desc_kmajor_noswizzle(smem_u32(sm.B[ab]), 128, B_SBO);
#pragma unroll
            for (int ks = 0; ks < BLK / 8; ++ks) {  // 8-pixel k-steps: two 128-byte core-matrix columns of B each
              umma_tf32_ts(d_addr, a_addr + 8u * ks, bdesc0 + (uint64_t)((ks * 256) >> 4), idesc, !first);
              first = false;
            }
            umma_commit(&sm.ab_empty[ab]);
            ++na;
            tm.lap(t_issue);
          }
        }
        umma_commit(&sm.d_full[ch.e & 1]);
        if (tm.on) {
          atomicAdd(&g_dbg[6], t_ff); atomicAdd(&g_dbg[7], t_ab); atomicAdd(&g_dbg[8], t_issue);
          atomicAdd(&g_dbg[9], t_dempty);
        }
      }
    } else {
      // ======================================================================= operand warps (12: h, 13: l, 14: pose)
      const int ow = warp - W_OPH;  // 0 / 1 / 2 == TMEM lane quarter
      const uint32_t lane_taddr = tbase + ((uint32_t)(ow * 32) << 16);
      const int row = ow * 32 + lane;  // TMEM lane == row of the partial
      ChainState ch;
      int na = 0;
      int chain_valid = 0;        // valid pixels accumulated into the current chain
      int cur_slot = -1;
      bool slot_fresh = true;     // the current item's partial has not been written yet by this CTA
      unsigned int inliers = 0;   // of the current item (warp h reports)
      // deferred drain of a finished chain
      bool pend = false;
      int pend_e = 0, pend_valid = 0, pend_slot = 0;
      bool pend_fresh = false, pend_item_end = false;
      unsigned int pend_inliers = 0;

      // Move a finished chain TMEM -> the CTA's partial in global memory (single writer, fixed order).
      // fresh: first chain of the item in this CTA (store), else fire-and-forget red.global.add.f32 in program order.
      auto drain = [&](int e, int valid, int slot, bool fresh, bool item_end, unsigned int inl) {
        const int bb = e & 1, use = e >> 1;
        float* P = partials + (size_t)slot * kTcPartialFloats;
        mbar_wait(&sm.d_full[bb], (uint32_t)use & 1u);
        tc_fence_after();
        if (valid > 0 || fresh) {
#pragma unroll 1
          for (int pass = 0; pass < 3; ++pass) {
            const int nq = pass < 2 ? 4 : (kTcCols - 32) / 4;  // float4 per pass (columns 40..47 are padding)
            uint32_t v[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = 0u;
            if (valid > 0) {
              tmem_ld_x16(lane_taddr + D_COL + NB * bb + 16 * pass, v);
              tmem_wait_ld();
            }
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
        if (item_end && ow == 0 && lane == 0) reinterpret_cast<unsigned int*>(P)[kTcRowsPad * kTcCols] = inl;
      };

#ifdef DFK_TC_TIMERS
      Tmr tm{0, dbg != 0 && ow == 0 && lane == 0};
#else
      Tmr tm{0, false};
#endif
      unsigned long long t_ffull = 0, t_abempty = 0, t_tma = 0, t_build = 0, t_sync = 0, t_drain = 0, t_total = 0;
#ifdef DFK_TC_TIMERS
      const long long t_begin = tm.on ? clock64() : 0;
#endif
      for (int b = 0; b < nblk; ++b) {
        const int st = b % NST;
        const int fs = b % NFS;
        tm.start();
        mbar_wait(&sm.feat_full[fs], (uint32_t)(b / NFS) & 1u);
        tm.lap(t_ffull);
        const BlockMeta meta = sm.meta[fs];
        if (ch.starts_chain(b, meta.item_changed)) {
          if (b > 0) {
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
          ch.blocks_in_chain = 0;
          chain_valid = 0;
          if (meta.item_changed || b == 0) {
            cur_slot = meta.pslot;
            slot_fresh = true;
            inliers = 0;
          }
        }
        chain_valid += meta.nv;
        inliers += (unsigned)meta.nv;

        if (meta.nv > 0) {
          ch.blocks_in_chain += 1;
          const int ab = na % NAB;
          // A/B slot `ab` was last read by the MMAs of non-empty block na - NAB
          if (na >= NAB) mbar_wait(&sm.ab_empty[ab], ((uint32_t)(na / NAB) - 1u) & 1u);
          tc_fence_after();
          tm.lap(t_abempty);
          const uint32_t a_taddr = lane_taddr + A_COL + 32u * (uint32_t)ab;
          unsigned char* bslot = sm.B[ab];
          if (ow < 2) {
            if (meta.bulk) mbar_wait(&sm.tma_full[st], (uint32_t)(b / NST) & 1u);
            tm.lap(t_tma);
            const float4* s4p = reinterpret_cast<const float4*>(sm.feat[fs][0]);
            const float* jcs = sm.jc[st] + lane;                // raw rows: jcs[p * C]
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
                for (int j = 0; j < 16; ++j) val[j] = jcs[(16 * half + j) * C];
              } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) val[j] = (16 * half + j) < meta.n ? __ldg(jcg + (16 * half + j) * C) : 0.0f;
              }
              uint32_t v[16];
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const float sc[4] = {s4[q].x, s4[q].y, s4[q].z, s4[q].w};
#pragma unroll
                for (int c4 = 0; c4 < 4; ++c4) {
                  // s == 0 (invalid pixel, or one past the image edge whose stage row is stale): exactly zero, whatever
                  // the row holds
                  const float x = sc[c4] != 0.0f ? sc[c4] * val[4 * q + c4] : 0.0f;
                  val[4 * q + c4] = x;
                  v[4 * q + c4] = __float_as_uint(ow == 0 ? x : x - tf32_trunc(x));
                }
              }
              if (ow == 0) {
                // B rows = features (this lane), k-chunks of 4 pixels, 128 B apart
#pragma unroll
                for (int q = 0; q < 4; ++q)
                  brow[8 * (4 * half + q)] = make_float4(val[4 * q], val[4 * q + 1], val[4 * q + 2], val[4 * q + 3]);
              }
              tmem_st_x16(a_taddr + 16u * half, v);
            }
          } else {
            tm.lap(t_tma);
            // pose / residual features: lanes 0-6 = h of feature 1+lane, lanes 7-13 = l of feature 1+(lane-7)
            const int f = 1 + (lane < 7 ? lane : (lane < 14 ? lane - 7 : 0));
            const float4* fp = reinterpret_cast<const float4*>(sm.feat[fs][f]);
            float4 x[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) x[q] = fp[q];
            if (lane < 7) {
              const uint32_t brow_i = 32u + (uint32_t)lane;
              float4* brow = reinterpret_cast<float4*>(bslot + (brow_i >> 3) * B_SBO + (brow_i & 7u) * 16u);
#pragma unroll
              for (int q = 0; q < 8; ++q) brow[8 * q] = x[q];
            }
            uint32_t v[32];
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
            tmem_st_x32(a_taddr, v);
          }
          tm.lap(t_build);
          tmem_wait_st();
          if (ow != 1) fence_proxy_async_smem();  // the l warp wrote TMEM only, no B rows
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&sm.ab_full[ab]);
          ++na;
          tm.lap(t_sync);
        }
        // an empty block still owns a bulk copy in flight into its ring stage: the stage (and the phase of its barrier) may
        // only be handed back once that copy has landed
        if (ow < 2 && meta.nv == 0 && meta.bulk) mbar_wait(&sm.tma_full[st], (uint32_t)(b / NST) & 1u);
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&sm.feat_empty[fs]);
          if (ow < 2) mbar_arrive(&sm.stage_empty[st]);
        }
        tm.start();
        if (pend) {  // the chain that ended before this block: its MMAs were issued long ago
          drain(pend_e, pend_valid, pend_slot, pend_fresh, pend_item_end, pend_inliers);
          pend = false;
        }
        tm.lap(t_drain);
      }
      drain(ch.e, chain_valid, cur_slot, slot_fresh, true, inliers);
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
    unsigned long long z[24] = {0};
    cudaMemcpyToSymbolAsync(g_dbg, z, sizeof(z), 0, cudaMemcpyHostToDevice, stream);
  }
  if (ev_start) cudaEventRecord(ev_start, stream);
  sfm_step_tc_kernel<<<plan.num_ctas, THREADS, smem, stream>>>(items_dev, plan.num_items, plan.num_tiles, partials_dev,
                                                              dbg);
  if (ev_stop) cudaEventRecord(ev_stop, stream);
  if (dbg) {
    unsigned long long v[24];
    cudaStreamSynchronize(stream);
    cudaMemcpyFromSymbol(v, g_dbg, sizeof(v));
    const double nb = v[17] ? (double)v[17] : 1.0;  // blocks (summed over CTAs)
    fprintf(stderr,
            "[dfk tc dbg] ctas=%d blocks=%d | cycles per block: FE(lane 0 of 12 warps) head %.0f geom %.0f gather %.0f tma_wait %.0f "
            "slot_wait %.0f write %.0f | CTRL feat_wait %.0f ab_wait %.0f issue %.0f d_empty %.0f | OPh feat_wait %.0f ab_empty %.0f "
            "tma %.0f build %.0f sync %.0f drain %.0f total %.0f\n",
            plan.num_ctas, plan.num_tiles, v[0] / nb, v[1] / nb, v[2] / nb, v[5] / nb, v[3] / nb, v[4] / nb, v[6] / nb, v[7] / nb,
            v[8] / nb, v[9] / nb, v[10] / nb, v[11] / nb, v[12] / nb, v[13] / nb, v[14] / nb, v[15] / nb, v[16] / nb);
  }
  return cudaGetLastError();
}

}  // namespace dfk
