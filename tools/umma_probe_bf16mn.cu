// umma_probe_bf16mn.cu -- tcgen05.mma kind::f16 (bf16 inputs, fp32 accumulate), both operands in shared memory in the
// MN-major (pixel-major) SWIZZLE_128B layout: atom = 64 features x 8 pixels (1 KB), pixel row r = 128 contiguous bytes
// whose 16-byte chunk c sits at chunk position c ^ r.  A = features 0..127 (two atoms, LBO apart), B = features
// 0..N-1 (N = 80: atom 0 and the first 16 features of atom 1); K = 16 pixels per instruction = two K atoms (SBO apart).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -o umma_probe_bf16mn umma_probe_bf16mn.cu
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../deepfactors_b200/csrc/dfk_async.cuh"
#include "../deepfactors_b200/csrc/dfk_tcgen05.cuh"

using namespace dfk;

constexpr int NF = 128, KT = 32, NMAX = 96;
constexpr uint32_t LBO = (KT / 8) * 1024, SBO = 1024;

__host__ __device__ inline uint32_t feat_addr(int f, int k)
{
  return (f / 64) * LBO + (k / 8) * SBO + (k % 8) * 128 + ((((f % 64) / 8) ^ (k % 8)) * 16) + (f % 8) * 2;
}
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout)
{
  return (uint64_t)((addr & 0x3ffffu) >> 4) | ((uint64_t)((lbo >> 4) & 0x3fffu) << 16) |
         ((uint64_t)((sbo >> 4) & 0x3fffu) << 32) | (1ull << 46) | ((uint64_t)layout << 61);
}
// kind::f16: [4,6) c = F32 (1) | [7,10) a = BF16 (1) | [10,13) b = BF16 (1) | [15] a MN-major | [16] b MN-major | N>>3 | M>>4
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N)
{
  return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__global__ void __launch_bounds__(128) probe_kernel(const float* __restrict__ F, float* __restrict__ D, int n, int m)
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
  for (int e = tid; e < NF * KT; e += 128) {
    const int f = e / KT, k = e % KT;
    const __nv_bfloat16 b = __float2bfloat16_rn(F[e]);
    const unsigned short u = *reinterpret_cast<const unsigned short*>(&b);
    asm volatile("st.shared.u16 [%0], %1;" ::"r"(sbase + feat_addr(f, k)), "h"(u) : "memory");
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tmem_base_s;
  const uint32_t lane_addr = tbase + ((uint32_t)(warp * 32) << 16);
  if (tid == 0) {
    const uint32_t idesc = make_idesc_bf16(m, n);
    for (int ks = 0; ks < KT / 16; ++ks) {
      const uint64_t ad = make_desc(sbase + ks * 2 * SBO, LBO, SBO, 2);
      asm volatile(
          "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
          "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tbase),
          "l"(ad), "l"(ad), "r"(idesc), "r"((uint32_t)(ks > 0))
          : "memory");
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

static float bf16_rn(float x)
{
  uint32_t u;
  memcpy(&u, &x, 4);
  u = (u + 0x7fffu + ((u >> 16) & 1u)) & 0xffff0000u;
  memcpy(&x, &u, 4);
  return x;
}

int main()
{
  const size_t nf = (size_t)NF * KT;
  float* hF = (float*)malloc(nf * 4);
  srand(7);
  for (size_t i = 0; i < nf; ++i) hF[i] = (float)(rand() % 2001 - 1000) / 1000.0f;
  static double ref[NF][NF];
  for (int a = 0; a < NF; ++a)
    for (int b = 0; b < NF; ++b) {
      double s = 0;
      for (int k = 0; k < KT; ++k) s += (double)bf16_rn(hF[a * KT + k]) * (double)bf16_rn(hF[b * KT + k]);
      ref[a][b] = s;
    }
  float *dF, *dD;
  cudaMalloc(&dF, nf * 4);
  cudaMalloc(&dD, 128 * NMAX * 4);
  cudaMemcpy(dF, hF, nf * 4, cudaMemcpyHostToDevice);
  const size_t smem = NF * KT * 2 + 2048;
  static float hD[128 * NMAX];
  const int ns[] = {64, 80, 48};
  for (int n : ns) {
    cudaMemset(dD, 0, sizeof(hD));
    probe_kernel<<<1, 128, smem>>>(dF, dD, n, 128);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("N=%d CUDA error %s\n", n, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(hD, dD, sizeof(hD), cudaMemcpyDeviceToHost);
    double err = 0;
    for (int a = 0; a < 128; ++a)
      for (int b = 0; b < n; ++b) err = fmax(err, fabs((double)hD[a * NMAX + b] - ref[a][b]));
    printf("bf16 MN-major SW128 M=128 N=%d K=16x2: max err %.3e  D[0][0] %.5f/%.5f D[100][70] %.5f/%.5f\n", n, err, hD[0], ref[0][0],
           hD[100 * NMAX + 70], ref[100][70]);
  }
  return 0;
}
