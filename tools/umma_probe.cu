// umma_probe.cu -- standalone probe of the tcgen05 pieces the tensor-core Gram kernel relies on:
//   * A operand (M=128 x K) in TMEM written with tcgen05.st.32x32b (lane = row, column = k)
//   * B operand (N=48 x K) in shared memory, K-major, no swizzle (8x16B core matrices, LBO/SBO)
//   * tcgen05.mma.cta_group::1.kind::tf32 M=128 N=48 K=8, fp32 accumulate in TMEM, commit -> mbarrier
//   * tcgen05.ld of the accumulator
// and of two numerical facts: how fp32 bit patterns are converted to tf32 (truncate vs round) and how
// the fp32 accumulation rounds.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -o umma_probe umma_probe.cu
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../deepfactors_b200/csrc/dfk_async.cuh"
#include "../deepfactors_b200/csrc/dfk_tcgen05.cuh"

using namespace dfk;

constexpr int M = 128, N = 48, KT = 64;  // KT = total K (8 k-steps of 8)

__global__ void __launch_bounds__(128) probe_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                    float* __restrict__ D, int rounds)
{
  __shared__ __align__(128) float Bs[(N / 8) * (KT / 4) * 32];  // 6 row groups x 16 k-chunks x 128 B
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(8) uint64_t bar;
  const int tid = threadIdx.x, warp = tid >> 5;

  if (warp == 0) {
    tmem_alloc(&tmem_base_s, 128);
    tmem_relinquish();
  }
  if (tid == 0) {
    mbar_init(&bar, 1);
    mbar_fence_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tmem_base_s;
  const uint32_t a_col = 0, d_col = 64;

  // A: thread t owns row t -> lane t, columns a_col .. a_col+63
  const uint32_t lane_addr = tbase + ((uint32_t)(warp * 32) << 16);
  for (int c = 0; c < KT; c += 8) {
    uint32_t v[8];
    for (int j = 0; j < 8; ++j) v[j] = __float_as_uint(A[tid * KT + c + j]);
    tmem_st_x8(lane_addr + a_col + c, v);
  }
  // B: row n (< 48), k -> (n/8)*SBO + (k/4)*128 + (n%8)*16 + (k%4)*4 bytes, SBO = (KT/4)*128
  constexpr uint32_t SBO = (KT / 4) * 128;
  if (tid < N) {
    for (int k = 0; k < KT; ++k) {
      const uint32_t off = (tid / 8) * SBO + (k / 4) * 128 + (tid % 8) * 16 + (k % 4) * 4;
      Bs[off / 4] = B[tid * KT + k];
    }
  }
  fence_proxy_async_smem();
  tmem_wait_st();
  tc_fence_before();
  __syncthreads();
  if (tid == 0) {
    tc_fence_after();
    const uint32_t idesc = make_idesc_tf32(M, N);
    const uint64_t bdesc0 = make_smem_desc_kmajor_noswizzle(smem_u32(Bs), /*lbo*/ 128, /*sbo*/ SBO);
    for (int r = 0; r < rounds; ++r)
      for (int ks = 0; ks < KT / 8; ++ks) {
        const uint64_t bdesc = bdesc0 + (uint64_t)((ks * 256) >> 4);
        umma_tf32_ts(tbase + d_col, tbase + a_col + ks * 8, bdesc, idesc, (r | ks) != 0);
      }
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  uint32_t out[48];
  tmem_ld_x16(lane_addr + d_col, out);
  tmem_ld_x16(lane_addr + d_col + 16, out + 16);
  tmem_ld_x16(lane_addr + d_col + 32, out + 32);
  tmem_wait_ld();
  for (int j = 0; j < N; ++j) D[tid * N + j] = __uint_as_float(out[j]);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tbase, 128);
}

static float trunc_tf32(float x)
{
  uint32_t u;
  memcpy(&u, &x, 4);
  u &= 0xffffe000u;
  memcpy(&x, &u, 4);
  return x;
}
static float rna_tf32(float x)
{
  uint32_t u;
  memcpy(&u, &x, 4);
  u += 0x1000u;
  u &= 0xffffe000u;
  memcpy(&x, &u, 4);
  return x;
}

int main()
{
  float *hA = (float*)malloc(M * KT * 4), *hB = (float*)malloc(N * KT * 4), *hD = (float*)malloc(M * N * 4);
  srand(7);
  for (int i = 0; i < M * KT; ++i) hA[i] = (float)rand() / RAND_MAX * 2.f - 1.f;
  for (int i = 0; i < N * KT; ++i) hB[i] = (float)rand() / RAND_MAX * 2.f - 1.f;
  float *dA, *dB, *dD;
  cudaMalloc(&dA, M * KT * 4);
  cudaMalloc(&dB, N * KT * 4);
  cudaMalloc(&dD, M * N * 4);
  cudaMemcpy(dA, hA, M * KT * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB, N * KT * 4, cudaMemcpyHostToDevice);
  for (int rounds = 1; rounds <= 64; rounds *= 64) {
    cudaMemset(dD, 0, M * N * 4);
    probe_kernel<<<1, 128>>>(dA, dB, dD, rounds);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      printf("CUDA error: %s\n", cudaGetErrorString(e));
      return 1;
    }
    cudaMemcpy(hD, dD, M * N * 4, cudaMemcpyDeviceToHost);
    double e_trunc = 0, e_rna = 0, e_full = 0, scale = 0;
    for (int i = 0; i < M; ++i)
      for (int j = 0; j < N; ++j) {
        double st = 0, sr = 0, sf = 0;
        for (int k = 0; k < KT; ++k) {
          st += (double)trunc_tf32(hA[i * KT + k]) * trunc_tf32(hB[j * KT + k]);
          sr += (double)rna_tf32(hA[i * KT + k]) * rna_tf32(hB[j * KT + k]);
          sf += (double)hA[i * KT + k] * hB[j * KT + k];
        }
        st *= rounds; sr *= rounds; sf *= rounds;
        const double d = hD[i * N + j];
        e_trunc = fmax(e_trunc, fabs(d - st));
        e_rna = fmax(e_rna, fabs(d - sr));
        e_full = fmax(e_full, fabs(d - sf));
        scale = fmax(scale, fabs(sf));
      }
    printf("rounds=%d  max|D| ~ %.3f   max err vs trunc-tf32 inputs: %.3e   vs rna-tf32 inputs: %.3e   vs fp32 inputs: %.3e\n",
           rounds, scale, e_trunc, e_rna, e_full);
    printf("   D[0][0..3] = %.6f %.6f %.6f %.6f ; D[127][47] = %.6f\n", hD[0], hD[1], hD[2], hD[3], hD[127 * N + 47]);
  }
  // accumulation rounding: all-positive terms, many rounds
  for (int i = 0; i < M * KT; ++i) hA[i] = trunc_tf32(0.5f + (float)rand() / RAND_MAX);
  for (int i = 0; i < N * KT; ++i) hB[i] = trunc_tf32(0.5f + (float)rand() / RAND_MAX);
  cudaMemcpy(dA, hA, M * KT * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB, N * KT * 4, cudaMemcpyHostToDevice);
  const int rounds = 512;  // 32768 positive terms per entry
  probe_kernel<<<1, 128>>>(dA, dB, dD, rounds);
  cudaDeviceSynchronize();
  cudaMemcpy(hD, dD, M * N * 4, cudaMemcpyDeviceToHost);
  double worst = 0, mean = 0;
  for (int i = 0; i < M; ++i)
    for (int j = 0; j < N; ++j) {
      double s = 0;
      for (int k = 0; k < KT; ++k) s += (double)hA[i * KT + k] * hB[j * KT + k];
      s *= rounds;
      const double rel = (hD[i * N + j] - s) / s;
      worst = fmax(worst, fabs(rel));
      mean += rel;
    }
  printf("accumulate %d positive terms: worst rel err %.3e, mean signed rel err %.3e (negative => truncating adds)\n",
         rounds * KT, worst, mean / (M * N));
  printf("PROBE_DONE\n");
  return 0;
}
