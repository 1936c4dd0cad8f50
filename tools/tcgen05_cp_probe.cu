// tcgen05_cp_probe.cu -- probe for round 2: can `tcgen05.cp` move the A operand of the Gram
// from shared memory to TMEM, so that no operand warps (LDS + tcgen05.st) are needed?
//   source  : feature-major tile F[feature row 0..127][pixel 0..KT-1] in shared memory, in one of two layouts
//               layout 0: K-major, no swizzle      addr(f,k) = (f/8)*(KT*32) + (k/4)*128 + (f%8)*16 + (k%4)*4
//               layout 1: K-major, SWIZZLE_128B    addr(f,k) = (k/32)*KBLK + (f/8)*1024 + (f%8)*128 + (((k%32)/4) ^ (f%8))*16 + (k%4)*4
//   copy    : per 8 pixels one `tcgen05.cp.cta_group::1.128x256b [tmem + 8*ks], desc` (128 lanes x 8 fp32 columns)
//   check   : tcgen05.ld of the 128 x KT block == F, then one TS-form MMA chain over it (A from TMEM, B = the same tile
//             through the matching K-major descriptor) == F F^T restricted to 48 columns.
// Measured on B200 (sm_100a, CUDA 12.9): both layouts copy exactly (0 mismatches of 8192 words) and the TS-form MMA chain
// over the copied A gives max|D - ref| = 1.1e-5 (tf32 truncation) -- tcgen05.cp takes the same K-major descriptors as the
// MMA operands, SWIZZLE_128B included, with 32-byte k-step advances inside the swizzle atom.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -o tcgen05_cp_probe tcgen05_cp_probe.cu
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../deepfactors_b200/csrc/dfk_async.cuh"
#include "../deepfactors_b200/csrc/dfk_tcgen05.cuh"

using namespace dfk;

constexpr int M = 128, N = 48, KT = 64;
constexpr uint32_t KBLK = 128 * 128;  // bytes per 32-pixel K block in layout 1 (128 rows x 128 B)

__host__ __device__ inline uint32_t feat_addr(int layout, int f, int k)
{
  if (layout == 0) return (f / 8) * (KT * 32) + (k / 4) * 128 + (f % 8) * 16 + (k % 4) * 4;
  return (k / 32) * KBLK + (f / 8) * 1024 + (f % 8) * 128 + ((((k % 32) / 4) ^ (f % 8)) * 16) + (k % 4) * 4;
}

__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout)
{
  return (uint64_t)((addr & 0x3ffffu) >> 4) | ((uint64_t)((lbo >> 4) & 0x3fffu) << 16) |
         ((uint64_t)((sbo >> 4) & 0x3fffu) << 32) | (1ull << 46) | ((uint64_t)layout << 61);
}

__device__ __forceinline__ void tmem_cp_128x256b(uint32_t taddr, uint64_t desc)
{
  asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(taddr), "l"(desc) : "memory");
}

__global__ void __launch_bounds__(128) probe_kernel(const float* __restrict__ F, float* __restrict__ Acopy,
                                                    float* __restrict__ D, int layout)
{
  extern __shared__ unsigned char smem[];
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(8) uint64_t bar;
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t sbase = (smem_u32(smem) + 1023u) & ~1023u;

  if (warp == 0) {
    tmem_alloc(&tmem_base_s, 128);
    tmem_relinquish();
  }
  if (tid == 0) {
    mbar_init(&bar, 1);
    mbar_fence_init();
  }
  for (int e = tid; e < M * KT; e += 128) {
    const int f = e / KT, k = e % KT;
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(sbase + feat_addr(layout, f, k)), "f"(F[e]) : "memory");
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tmem_base_s;
  const uint32_t a_col = 0, d_col = 64;
  const uint32_t lbo = layout == 0 ? 128u : 16u;
  const uint32_t sbo = layout == 0 ? (uint32_t)(KT * 32) : 1024u;
  const uint32_t ltype = layout == 0 ? 0u : 2u;
  if (tid == 0) {
    // A: smem -> TMEM, 8 pixels (256 bit) per instruction
    for (int ks = 0; ks < KT / 8; ++ks) {
      const uint32_t start = layout == 0 ? sbase + ks * 256 : sbase + (ks / 4) * KBLK + (ks % 4) * 32;
      tmem_cp_128x256b(tbase + a_col + 8 * ks, make_desc(start, lbo, sbo, ltype));
    }
    // D = A (TMEM) * B^T (smem, first N rows of the same tile)
    const uint32_t idesc = make_idesc_tf32(M, N);
    for (int ks = 0; ks < KT / 8; ++ks) {
      const uint32_t start = layout == 0 ? sbase + ks * 256 : sbase + (ks / 4) * KBLK + (ks % 4) * 32;
      umma_tf32_ts(tbase + d_col, tbase + a_col + 8 * ks, make_desc(start, lbo, sbo, ltype), idesc, ks > 0);
    }
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  const uint32_t lane_addr = tbase + ((uint32_t)(warp * 32) << 16);
  for (int c = 0; c < KT; c += 16) {
    uint32_t v[16];
    tmem_ld_x16(lane_addr + a_col + c, v);
    tmem_wait_ld();
    for (int j = 0; j < 16; ++j) Acopy[tid * KT + c + j] = __uint_as_float(v[j]);
  }
  for (int c = 0; c < N; c += 16) {
    uint32_t v[16];
    tmem_ld_x16(lane_addr + d_col + c, v);
    tmem_wait_ld();
    for (int j = 0; j < 16; ++j) D[tid * N + c + j] = __uint_as_float(v[j]);
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
  static float hF[M * KT], hA[M * KT], hD[M * N];
  srand(11);
  for (int i = 0; i < M * KT; ++i) hF[i] = (float)(rand() % 2001 - 1000) / 1000.0f;
  float *dF, *dA, *dD;
  cudaMalloc(&dF, sizeof(hF));
  cudaMalloc(&dA, sizeof(hA));
  cudaMalloc(&dD, sizeof(hD));
  cudaMemcpy(dF, hF, sizeof(hF), cudaMemcpyHostToDevice);
  const size_t smem = 2 * KBLK + 2048;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  for (int layout = 0; layout < 2; ++layout) {
    cudaMemset(dA, 0, sizeof(hA));
    cudaMemset(dD, 0, sizeof(hD));
    probe_kernel<<<1, 128, smem>>>(dF, dA, dD, layout);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      printf("layout %d: CUDA error %s\n", layout, cudaGetErrorString(e));
      return 1;
    }
    cudaMemcpy(hA, dA, sizeof(hA), cudaMemcpyDeviceToHost);
    cudaMemcpy(hD, dD, sizeof(hD), cudaMemcpyDeviceToHost);
    int bad_copy = 0;
    for (int i = 0; i < M * KT; ++i) bad_copy += (hA[i] != hF[i]);
    double err = 0;
    for (int m = 0; m < M; ++m)
      for (int n = 0; n < N; ++n) {
        double s = 0;
        for (int k = 0; k < KT; ++k) s += (double)tf32_trunc(hF[m * KT + k]) * (double)tf32_trunc(hF[n * KT + k]);
        err = fmax(err, fabs(s - (double)hD[m * N + n]));
      }
    printf("layout %d (%s): tcgen05.cp copy mismatches %d / %d ; A[5][3]=%.4f want %.4f ; max|D - ref| = %.3e\n", layout,
           layout ? "K-major SWIZZLE_128B" : "K-major no swizzle", bad_copy, M * KT, hA[5 * KT + 3], hF[5 * KT + 3], err);
  }
  return 0;
}
