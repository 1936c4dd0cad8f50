// umma_probe_mn.cu -- probe of tcgen05.mma kind::tf32 with BOTH operands in shared memory (SS form), in particular
// MN-major operands with the 128-byte swizzle: the layout a pixel-major feature tile has naturally (row = pixel = K
// index, 32 contiguous floats = 32 features = one MN block, chunk j of row r at 16-byte position j ^ (r & 7)).
// One feature matrix F[feature][pixel] is laid out in shared memory in one of three ways and used as A (features
// 0..127) and B (features 0..N-1); D = A * B^T over KT pixels is compared with a CPU reference (tf32-truncated inputs).
//   layout 0: MN-major, SWIZZLE_128B   addr(f,k) = (f/32)*BLK + k*128 + (((f%32)/4) ^ (k&7))*16 + (f%4)*4
//   layout 1: MN-major, no swizzle     addr(f,k) = (f/4)*(KT*16) + (k/8)*128 + (k%8)*16 + (f%4)*4
//   layout 2: K-major,  no swizzle     addr(f,k) = (f/8)*(KT*32) + (k/4)*128 + (f%8)*16 + (k%4)*4
//   layout 3: K-major,  SWIZZLE_128B   addr(f,k) = (k/32)*KBLK + (f/8)*1024 + (f%8)*128 + (((k%32)/4) ^ (f%8))*16 + (k%4)*4
//             (only features 0..79 are stored; K blocks 10 KB apart, so D rows 80..127 are junk by construction)
// Measured on B200 (sm_100a, CUDA 12.9): the SS form works for K-major operands, no swizzle and SWIZZLE_128B alike
// (error 1e-5 = tf32 truncation), including start addresses advanced by 32 bytes per k-step inside the swizzle atom, K
// blocks 10 KB apart and LBO = 0 / 16 / 1024 (ignored); every MN-major variant (a_major = b_major = 1) WRITES ZEROS.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -o umma_probe_mn umma_probe_mn.cu
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../deepfactors_b200/csrc/dfk_async.cuh"
#include "../deepfactors_b200/csrc/dfk_tcgen05.cuh"

using namespace dfk;

constexpr int M = 128, NF = 128, KT = 64, NMAX = 96;
constexpr uint32_t BLK_BYTES = KT * 128;
constexpr uint32_t TILE_BYTES = NF * KT * 4;
constexpr uint32_t KBLK3 = 80 * 128;  // rows 80..127 of a K block overlap the next block (junk rows of D)

struct Params {
  int layout;         // 0,1,2 (see above)
  uint32_t lbo, sbo;  // descriptor fields (bytes)
  uint32_t kstep;     // descriptor start-address advance per k-step of 8 pixels (bytes)
  uint32_t ltype;     // descriptor layout type
  uint32_t idesc;
  int n;
  uint32_t lbo_b;     // 0: B uses the A descriptor; else B's own LBO (B's n-th MN atom = feature atom n * lbo_b / lbo)
  uint32_t base_off;  // byte offset of the tile inside the 1024-byte aligned window
  int m64;            // 1: M = 64 with A = feature atoms 0 and 2 (rows 0-31 = features 0-31, rows 32-63 = features 64-95)
};

__host__ __device__ inline uint32_t feat_addr(int layout, int f, int k)
{
  if (layout == 0) return (f / 32) * BLK_BYTES + k * 128 + ((((f % 32) / 4) ^ (k & 7)) * 16) + (f % 4) * 4;
  if (layout == 1) return (f / 4) * (KT * 16) + (k / 8) * 128 + (k % 8) * 16 + (f % 4) * 4;
  if (layout == 2) return (f / 8) * (KT * 32) + (k / 4) * 128 + (f % 8) * 16 + (k % 4) * 4;
  // layout 4: MN-major SWIZZLE_128B_BASE32B (layout type 1): atoms of 32 features x 4 pixels (512 B): pixel k%4 = 128-byte
  // row, 32-byte chunk (f%32)/8 of the row at position ((f%32)/8) ^ (k%4); K atoms 512 B apart, MN atoms BLK_BYTES apart
  if (layout == 4) return (f / 32) * BLK_BYTES + (k / 4) * 512 + (k % 4) * 128 + ((((f % 32) / 8) ^ (k % 4)) * 32) + (f % 8) * 4;
  // layout 3: K-major SWIZZLE_128B, K blocks of 32 pixels: rows = features (128 B = 32 pixels), 8-row groups of 1024 B,
  // chunk (k%32)/4 of row f at position ((k%32)/4) ^ (f%8); K block q at q * KBLK3
  return (k / 32) * KBLK3 + (f / 8) * 1024 + (f % 8) * 128 + ((((k % 32) / 4) ^ (f % 8)) * 16) + (k % 4) * 4;
}

__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout)
{
  return (uint64_t)((addr & 0x3ffffu) >> 4) | ((uint64_t)((lbo >> 4) & 0x3fffu) << 16) |
         ((uint64_t)((sbo >> 4) & 0x3fffu) << 32) | (1ull << 46) | ((uint64_t)layout << 61);
}


// F: [NF][KT]
__global__ void __launch_bounds__(128) probe_kernel(const float* __restrict__ F, float* __restrict__ D, Params P)
{
  extern __shared__ unsigned char smem[];
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(8) uint64_t bar;
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t s0 = smem_u32(smem);
  const uint32_t sbase = ((s0 + 1023u) & ~1023u) + P.base_off;  // shared-window address, 1024-byte aligned (+ offset)

  if (warp == 0) {
    tmem_alloc(&tmem_base_s, 128);
    tmem_relinquish();
  }
  if (tid == 0) {
    mbar_init(&bar, 1);
    mbar_fence_init();
  }
  for (int e = tid; e < NF * KT; e += 128) {
    const int f = e / KT, k = e % KT;
    if (P.layout == 3 && f >= 80) continue;
    const uint32_t a = sbase + feat_addr(P.layout, f, k);
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(F[e]) : "memory");
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tmem_base_s;
  const uint32_t lane_addr = tbase + ((uint32_t)(warp * 32) << 16);
  {  // sentinel in the accumulator: shows whether the MMA wrote anything
    uint32_t v[8];
    for (int j = 0; j < 8; ++j) v[j] = __float_as_uint(7.0f);
    for (int c = 0; c < NMAX; c += 8) tmem_st_x8(lane_addr + c, v);
    tmem_wait_st();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (tid == 0) {
    for (int ks = 0; ks < KT / 8; ++ks) {
      const uint32_t start = (P.layout == 3) ? sbase + (ks / 4) * KBLK3 + (ks % 4) * 32 : sbase + ks * P.kstep;
      const uint64_t ad = make_desc(start, P.lbo, P.sbo, P.ltype);
      const uint64_t bd = P.lbo_b ? make_desc(start, P.lbo_b, P.sbo, P.ltype) : ad;
      umma_tf32_ss(tbase, ad, bd, P.idesc, ks > 0);  // B = the first N features of the same tile (or atoms 0, 2 with lbo_b = 2 lbo)
    }
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  for (int c = 0; c < NMAX; c += 16) {
    uint32_t v[16];
    tmem_ld_x16(lane_addr + c, v);
    tmem_wait_ld();
    for (int j = 0; j < 16; ++j) D[tid * NMAX + c + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tbase, 128);
}

static float tf32_trunc(float x)
{
  uint32_t u;
  memcpy(&u, &x, 4);
  u &= 0xffffe000u;
  memcpy(&x, &u, 4);
  return x;
}

int main()
{
  const size_t nf = (size_t)NF * KT;
  float* hF = (float*)malloc(nf * 4);
  srand(7);
  for (size_t i = 0; i < nf; ++i) hF[i] = (float)(rand() % 2001 - 1000) / 1000.0f;
  static double ref[M][NMAX];
  static double ref2[M][NF];
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < NF; ++n) {
      double s = 0;
      for (int k = 0; k < KT; ++k) s += (double)tf32_trunc(hF[m * KT + k]) * (double)tf32_trunc(hF[n * KT + k]);
      ref2[m][n] = s;
    }
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < NMAX; ++n) {
      double s = 0;
      for (int k = 0; k < KT; ++k) s += (double)tf32_trunc(hF[m * KT + k]) * (double)tf32_trunc(hF[n * KT + k]);
      ref[m][n] = s;
    }
  float *dF, *dD;
  cudaMalloc(&dF, nf * 4);
  cudaMalloc(&dD, M * NMAX * 4);
  cudaMemcpy(dF, hF, nf * 4, cudaMemcpyHostToDevice);
  const size_t smem = TILE_BYTES + 4096;
  cudaError_t e = cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) printf("attr: %s\n", cudaGetErrorString(e));
  const uint32_t MN = (1u << 15) | (1u << 16);
  struct Var { const char* name; Params p; } vars[] = {
      {"K-major noswz SS (sanity) N=48", {2, 128, KT * 32, 256, 0, make_idesc_tf32(M, 48), 48}},
      {"MN SW128 LBO=blk SBO=1024 N=48", {0, BLK_BYTES, 1024, 1024, 2, make_idesc_tf32(M, 48) | MN, 48}},
      {"MN SW128 LBO=blk SBO=1024 N=64", {0, BLK_BYTES, 1024, 1024, 2, make_idesc_tf32(M, 64) | MN, 64}},
      {"MN SW128 LBO=blk SBO=1024 N=32", {0, BLK_BYTES, 1024, 1024, 2, make_idesc_tf32(M, 32) | MN, 32}},
      {"MN SW128 LBO=1024 SBO=blk N=64", {0, 1024, BLK_BYTES, 1024, 2, make_idesc_tf32(M, 64) | MN, 64}},
      {"MN SW128 LBO=blk SBO=blk N=64", {0, BLK_BYTES, BLK_BYTES, 1024, 2, make_idesc_tf32(M, 64) | MN, 64}},
      {"MN noswz LBO=128 SBO=KT*16 N=48", {1, 128, KT * 16, 128, 0, make_idesc_tf32(M, 48) | MN, 48}},
      {"MN noswz LBO=KT*16 SBO=128 N=48", {1, KT * 16, 128, 128, 0, make_idesc_tf32(M, 48) | MN, 48}},
      {"K-major SW128 LBO=16 SBO=1024 N=48 (rows<80)", {3, 16, 1024, 0, 2, make_idesc_tf32(M, 48), 48}},
      {"K-major SW128 LBO=0 SBO=1024 N=48 (rows<80)", {3, 0, 1024, 0, 2, make_idesc_tf32(M, 48), 48}},
      {"K-major SW128 LBO=1024 SBO=1024 N=48", {3, 1024, 1024, 0, 2, make_idesc_tf32(M, 48), 48}},
      {"MN SW128 layout, K-major flags (control)", {0, BLK_BYTES, 1024, 1024, 2, make_idesc_tf32(M, 48), 48}},
      {"MN SW128_32B LBO=blk SBO=512 N=64", {4, BLK_BYTES, 512, 1024, 1, make_idesc_tf32(M, 64) | MN, 64}},
      {"MN SW128_32B LBO=blk SBO=512 N=48", {4, BLK_BYTES, 512, 1024, 1, make_idesc_tf32(M, 48) | MN, 48}},
      {"MN SW128_32B LBO=blk SBO=512 N=32", {4, BLK_BYTES, 512, 1024, 1, make_idesc_tf32(M, 32) | MN, 32}},
      {"MN SW128_32B LBO=512 SBO=blk N=64", {4, 512, BLK_BYTES, 1024, 1, make_idesc_tf32(M, 64) | MN, 64}},
      {"MN SW128_32B N=48, B atoms 0,2 (lbo_b=2blk)", {4, BLK_BYTES, 512, 1024, 1, make_idesc_tf32(M, 48) | MN, 48, 2 * BLK_BYTES}},
      {"MN SW128_32B N=64, B atoms 0,2 (lbo_b=2blk)", {4, BLK_BYTES, 512, 1024, 1, make_idesc_tf32(M, 64) | MN, 64, 2 * BLK_BYTES}},
      {"MN SW128_32B N=48 base+512", {4, BLK_BYTES, 512, 1024, 1, make_idesc_tf32(M, 48) | MN, 48, 0, 512}},
      {"MN SW128_32B N=48 base+128", {4, BLK_BYTES, 512, 1024, 1, make_idesc_tf32(M, 48) | MN, 48, 0, 128}},
      {"MN SW128_32B M=64 A atoms 0,2  N=80 B atoms 0,1,2", {4, 2 * BLK_BYTES, 512, 1024, 1, make_idesc_tf32(64, 80) | MN, 80, BLK_BYTES, 0, 1}},
      {"MN SW128_32B A only MN (B K-major flag)", {4, BLK_BYTES, 512, 1024, 1, make_idesc_tf32(M, 48) | (1u << 15), 48}},
  };
  static float hD[M * NMAX];
  for (const Var& v : vars) {
    cudaMemset(dD, 0, sizeof(hD));
    probe_kernel<<<1, 128, smem>>>(dF, dD, v.p);
    e = cudaGetLastError();
    if (e != cudaSuccess) { printf("%-44s launch error %s\n", v.name, cudaGetErrorString(e)); continue; }
    e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%-44s CUDA error %s\n", v.name, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(hD, dD, sizeof(hD), cudaMemcpyDeviceToHost);
    double err_all = 0, err_32 = 0, err_r32 = 0;
    int sentinels = 0;
    for (int m = 0; m < (v.p.layout == 3 ? 80 : M); ++m)
      for (int n = 0; n < v.p.n; ++n) {
        const int nb = (v.p.lbo_b && n >= 32) ? n + 32 : n;  // B atom 1 = feature atom 2
        const double d = fabs((double)hD[m * NMAX + n] - ref2[m][nb]);
        if (hD[m * NMAX + n] == 7.0f) ++sentinels;
        if (d > err_all) err_all = d;
        if (n < 32 && d > err_32) err_32 = d;
        if (n < 32 && m < 32 && d > err_r32) err_r32 = d;
      }
    if (v.p.m64) {  // where did the 64 rows go?  match every TMEM lane against the reference rows
      printf("%-44s lane -> A row:", v.name);
      for (int lane = 0; lane < 128; ++lane) {
        int best = -1;
        for (int r = 0; r < 64 && best < 0; ++r) {
          const int fa = r < 32 ? r : 32 + r;
          double e = 0;
          for (int n = 0; n < v.p.n; ++n) e = fmax(e, fabs((double)hD[lane * NMAX + n] - ref2[fa][n]));
          if (e < 1e-4) best = r;
        }
        if (best >= 0 && (lane % 16 == 0 || best % 16 != lane % 16)) printf(" %d:%d", lane, best);
      }
      printf("\n");
      continue;
    }
    printf("%-44s err all=%.2e cols<32=%.2e 32x32=%.2e untouched=%d  D00=%.4f/%.4f D[1][0]=%.4f/%.4f D[40][35]=%.4f/%.4f D[70][5]=%.4f/%.4f\n",
           v.name, err_all, err_32, err_r32, sentinels, hD[0], ref[0][0], hD[NMAX], ref[1][0], hD[40 * NMAX + 35],
           ref[40][35], hD[70 * NMAX + 5], ref[70][5]);
  }
  return 0;
}
