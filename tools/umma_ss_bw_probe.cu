// umma_ss_bw_probe.cu -- how fast does tcgen05.mma kind::tf32 run with BOTH operands in shared memory (SS form, MN-major
// SWIZZLE_128B_BASE32B, M = 128, N = 48, K = 8), and does its operand fetch compete with ordinary LDS / STS traffic?
// One CTA per SM (grid = #SMs), 4 + NW warps: warp 0 lane 0 issues NMMA back-to-back MMAs over a ring of operand blocks and
// waits for the commit; warps 4.. (NW of them) run a loop of STS.128 + LDS.128 on a private region (mode & 1) for as
// long as the MMAs run.  Reported: cycles per MMA alone, cycles per MMA with the LSU traffic, LSU bytes / cycle alone
// (mode 2: no MMAs, fixed iteration count) and together.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -o umma_ss_bw_probe umma_ss_bw_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../deepfactors_b200/csrc/dfk_async.cuh"
#include "../deepfactors_b200/csrc/dfk_tcgen05.cuh"

using namespace dfk;

constexpr int NW = 8;  // LSU warps

__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout)
{
  return (uint64_t)((addr & 0x3ffffu) >> 4) | ((uint64_t)((lbo >> 4) & 0x3fffu) << 16) |
         ((uint64_t)((sbo >> 4) & 0x3fffu) << 32) | (1ull << 46) | ((uint64_t)layout << 61);
}

// mode bit 0: MMAs, bit 1: LSU traffic.  a_in_tmem: use the TS form (A from TMEM columns 64..) for comparison
__global__ void __launch_bounds__(128 + 32 * NW) bw_kernel(int mode, int nmma, int m_rows, int n_cols, int a_in_tmem, int nacc,
                                                           unsigned long long* out)
{
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(8) uint64_t bar;
  __shared__ volatile int stop;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t sbase = (smem_u32(smem) + 1023u) & ~1023u;
  if (warp == 0) {
    tmem_alloc(&tmem_base_s, 256);
    tmem_relinquish();
  }
  if (tid == 0) {
    mbar_init(&bar, 1);
    mbar_fence_init();
    stop = 0;
  }
  for (int e = tid; e < 48 * 1024 / 4; e += blockDim.x)
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(sbase + 4 * e), "f"(1.0f / (float)(1 + (e & 1023))) : "memory");
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tmem_base_s;
  unsigned long long t_mma = 0, lsu_ops = 0, t_lsu = 0;
  if (warp == 0) {
    if (lane == 0 && (mode & 1)) {
      const uint32_t idesc = make_idesc_tf32(m_rows, n_cols) | (a_in_tmem ? (1u << 16) : ((1u << 15) | (1u << 16)));
      const long long t0 = clock64();
      for (int i = 0; i < nmma; ++i) {
        // operand blocks of 12 KB (3 MN atoms x 8 K atoms x 512 B), four of them, K-steps of 1024 B inside
        const uint32_t blk = sbase + (uint32_t)((i >> 2) & 3) * 12288u + (uint32_t)(i & 3) * 1024u;
        const uint64_t ad = make_desc(blk, 4096, 512, 1);
        const uint64_t bd = make_desc(blk, 8192, 512, 1);
        if (a_in_tmem) {
          asm volatile(
              "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
              "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tbase + 48 * (uint32_t)(i % nacc)),
              "r"(tbase + 192 + 8 * (i & 7)), "l"(bd), "r"(idesc), "r"((uint32_t)(i >= nacc))
              : "memory");
        } else {
          umma_tf32_ss(tbase + 48 * (uint32_t)(i % nacc), ad, bd, idesc, i >= nacc);
        }
      }
      umma_commit(&bar);
      mbar_wait(&bar, 0);
      t_mma = (unsigned long long)(clock64() - t0);
      stop = 1;
    }
  } else if (warp >= 4 && (mode & 2)) {
    // private 4 KB region per warp above the operand blocks; conflict-free 16-byte accesses
    const uint32_t mine = sbase + 49152u + (uint32_t)(warp - 4) * 4096u + (uint32_t)lane * 16u;
    float4 v = make_float4(1.f, 2.f, 3.f, 4.f), acc = make_float4(0.f, 0.f, 0.f, 0.f);
    const long long t0 = clock64();
    int it = 0;
    const int fixed = (mode & 1) ? (1 << 30) : nmma;  // alone: a fixed number of iterations
    while (it < fixed) {
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(mine + (uint32_t)r * 512u), "f"(v.x), "f"(v.y), "f"(v.z),
                     "f"(v.w)
                     : "memory");
        float4 q;
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(q.x), "=f"(q.y), "=f"(q.z), "=f"(q.w) : "r"(mine + (uint32_t)((r + 3) & 7) * 512u));
        acc.x += q.x; acc.y += q.y; acc.z += q.z; acc.w += q.w;
      }
      ++it;
      if ((mode & 1) && stop) break;
    }
    t_lsu = (unsigned long long)(clock64() - t0);
    lsu_ops = (unsigned long long)it * 16ull;  // warp-wide 512-byte accesses
    if (acc.x == 123.456f) out[7] = 1;
  }
  __syncthreads();
  if (blockIdx.x == 0) {
    if (tid == 0) { out[0] = t_mma; }
    if (tid == 128) { out[1] = t_lsu; out[2] = lsu_ops; }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tbase, 256);
}

int main()
{
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  unsigned long long* d;
  cudaMalloc(&d, 64);
  const size_t smem = 49152 + NW * 4096 + 2048;
  cudaFuncSetAttribute(bw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int nmma = 20000;
  struct Cfg { const char* name; int mode, m, n, ts, nacc; } cfgs[] = {
      {"SS M128 N48 1 acc", 1, 128, 48, 0, 1}, {"SS M128 N48 2 acc", 1, 128, 48, 0, 2}, {"SS M128 N48 4 acc", 1, 128, 48, 0, 4},
      {"SS M64 N80 1 acc(80 cols!)", 1, 64, 80, 0, 1}, {"SS M64 N48 2 acc", 1, 64, 48, 0, 2}, {"SS M64 N48 4 acc", 1, 64, 48, 0, 4},
      {"TS M128 N48 1 acc", 1, 128, 48, 1, 1}, {"TS M128 N48 2 acc", 1, 128, 48, 1, 2}, {"TS M128 N48 4 acc", 1, 128, 48, 1, 4},
      {"SS M128 N48 4 acc + LSU", 3, 128, 48, 0, 4}, {"LSU alone", 2, 128, 48, 0, 1},
  };
  for (const Cfg& c : cfgs) {
    unsigned long long h[8] = {0};
    cudaMemset(d, 0, 64);
    bw_kernel<<<sms, 128 + 32 * NW, smem>>>(c.mode, nmma, c.m, c.n, c.ts, c.nacc, d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: %s\n", c.name, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
    printf("%-30s", c.name);
    if (c.mode & 1) printf(" %.1f cycles / MMA", (double)h[0] / nmma);
    if (c.mode & 2) printf("  LSU: %.1f B/cycle/SM (%d warps, %.0f cycles)", (double)h[2] * 512.0 * NW / (double)h[1], NW, (double)h[1]);
    printf("\n");
  }
  return 0;
}
