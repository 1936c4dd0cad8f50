// dfk_tcgen05.cuh -- inline-PTX wrappers for the 5th-generation tensor-core path (sm_100a):
// TMEM allocation, tcgen05.st / tcgen05.ld, tcgen05.mma kind::tf32 with A in TMEM and B in shared
// memory, tcgen05.commit, the tcgen05 fences, and the descriptor encodings.
// Field layouts follow the PTX ISA "tcgen05" chapter as restated in the vendored CUTLASS headers
// (cute/arch/mma_sm100_desc.hpp, mma_sm100_umma.hpp, copy_sm100.hpp, tmem_allocator_sm100.hpp) --
// read for the bit layouts only; no CUTLASS code is compiled into this library.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "dfk_async.cuh"

namespace dfk {

// ---------------------------------------------------------------------------------- TMEM allocation
// one full warp; writes the base address (lane 0, column c) to *dst_smem.  ncols: power of two >= 32
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols)
{
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish()
{
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols)
{
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// ---------------------------------------------------------------------------------- fences
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// generic-proxy shared-memory writes -> visible to the async proxy (tcgen05.mma reading B through a descriptor)
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---------------------------------------------------------------------------------- TMEM <-> registers
// 32x32b shape: thread i of the warp accesses lane (quarter_base + i), `n` consecutive 32-bit columns.
// taddr = tmem_base + (lane_base << 16) + column, lane_base = 32 * (warp_id % 4).
__device__ __forceinline__ void tmem_st_x8(uint32_t taddr, const uint32_t (&v)[8])
{
  // no "memory" clobber: the instruction touches registers and TMEM only, so ordinary shared-memory
  // loads of the next chunk may be scheduled across it (ordering vs. tcgen05.wait::st comes from volatile)
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(v[0]),
               "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]));
}
__device__ __forceinline__ void tmem_st_x16(uint32_t taddr, const uint32_t (&v)[16])
{
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]));
}
__device__ __forceinline__ void tmem_st_x32(uint32_t taddr, const uint32_t (&v)[32])
{
  asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31]));
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld_x16(uint32_t taddr, uint32_t* v)
{
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, "
      "[%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---------------------------------------------------------------------------------- descriptors
// Instruction descriptor, kind::tf32, fp32 accumulate, A and B K-major, dense:
//   [4,6) c_format = 1 (F32) | [7,10) a_format = 2 (TF32) | [10,13) b_format = 2 (TF32)
//   [15] a_major = 0 | [16] b_major = 0 | [17,23) N >> 3 | [24,29) M >> 4
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N)
{
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// Shared-memory matrix descriptor, K-major operand, no swizzle ("interleave"): the operand is a grid of
// 8-row x 16-byte core matrices stored as 128 contiguous bytes; `lbo` = byte distance between core
// matrices adjacent in K, `sbo` = byte distance between core matrices adjacent in M/N.
//   [0,14) addr >> 4 | [16,30) lbo >> 4 | [32,46) sbo >> 4 | [46,48) version = 1 | [61,64) layout = 0
__device__ __forceinline__ uint64_t make_smem_desc_kmajor_noswizzle(uint32_t smem_addr, uint32_t lbo, uint32_t sbo)
{
  return (uint64_t)((smem_addr & 0x3ffffu) >> 4) | ((uint64_t)((lbo >> 4) & 0x3fffu) << 16) |
         ((uint64_t)((sbo >> 4) & 0x3fffu) << 32) | (1ull << 46);
}

// MN-major ("transposed") 32-bit operands: instruction-descriptor bits, and the one canonical shared-memory layout the
// hardware defines for them -- layout type 1 = SWIZZLE_128B_BASE32B (tools/umma_probe_mn.cu; SWIZZLE_NONE / _128B with
// the MN-major bits set write zeros).  Atom = 32 elements along M/N x 4 along K (512 B, 512-byte aligned): K row r is
// 128 contiguous bytes whose 32-byte chunk c sits at chunk position c ^ r.  `lbo` = byte distance between atoms
// adjacent in M/N, `sbo` = between atoms adjacent in K (one K = 8 instruction spans two).
constexpr uint32_t kIdescAMnMajor = 1u << 15;
constexpr uint32_t kIdescBMnMajor = 1u << 16;
__device__ __forceinline__ uint64_t make_smem_desc_mn_sw128_32b(uint32_t smem_addr, uint32_t lbo, uint32_t sbo)
{
  return (uint64_t)((smem_addr & 0x3ffffu) >> 4) | ((uint64_t)((lbo >> 4) & 0x3fffu) << 16) |
         ((uint64_t)((sbo >> 4) & 0x3fffu) << 32) | (1ull << 46) | (1ull << 61);
}

// ---------------------------------------------------------------------------------- MMA + commit
// D[tmem] (+)= A[tmem] * B[smem]^T ; one elected thread issues.
__device__ __forceinline__ void umma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                             bool accumulate)
{
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate)
      : "memory");
}

// D[tmem] (+)= A[smem] * B[smem]^T ; both operands through shared-memory descriptors
__device__ __forceinline__ void umma_tf32_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                             bool accumulate)
{
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate)
      : "memory");
}

// four k-steps of one operand block: descriptors {a_lo | hi}, {b_lo | hi}, start addresses advanced by `step` (16-byte
// units) per k-step; the first MMA accumulates iff `accumulate`, the rest always
__device__ __forceinline__ void umma_tf32_ss_x4(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t hi, uint32_t idesc,
                                                uint32_t accumulate, uint32_t step)
{
  asm volatile(
      "{\n\t"
      ".reg .pred p, t;\n\t"
      ".reg .b64 a, b, s;\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "setp.eq.b32 t, 0, 0;\n\t"
      "mov.b64 a, {%1, %3};\n\t"
      "mov.b64 b, {%2, %3};\n\t"
      "cvt.u64.u32 s, %6;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], a, b, %4, p;\n\t"
      "add.u64 a, a, s;\n\t"
      "add.u64 b, b, s;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], a, b, %4, t;\n\t"
      "add.u64 a, a, s;\n\t"
      "add.u64 b, b, s;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], a, b, %4, t;\n\t"
      "add.u64 a, a, s;\n\t"
      "add.u64 b, b, s;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], a, b, %4, t;\n\t"
      "}\n" ::"r"(d_tmem),
      "r"(a_lo), "r"(b_lo), "r"(hi), "r"(idesc), "r"(accumulate), "r"(step)
      : "memory");
}

// kind::f16 with bf16 inputs, fp32 accumulate:  [4,6) c = F32 (1) | [7,10) a = BF16 (1) | [10,13) b = BF16 (1) | N >> 3 | M >> 4
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N)
{
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, bool accumulate)
{
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate)
      : "memory");
}
// two k-steps of one operand block, A and B through the same descriptor {lo | hi}, start advanced by `step` (16-byte
// units); the first MMA accumulates iff `accumulate`, the second always
__device__ __forceinline__ void umma_bf16_ss_x2(uint32_t d_tmem, uint32_t lo, uint32_t hi, uint32_t idesc, uint32_t accumulate,
                                                uint32_t step)
{
  asm volatile(
      "{\n\t"
      ".reg .pred p, t;\n\t"
      ".reg .b64 a, s;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "setp.eq.b32 t, 0, 0;\n\t"
      "mov.b64 a, {%1, %2};\n\t"
      "cvt.u64.u32 s, %5;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], a, a, %3, p;\n\t"
      "add.u64 a, a, s;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], a, a, %3, t;\n\t"
      "}\n" ::"r"(d_tmem),
      "r"(lo), "r"(hi), "r"(idesc), "r"(accumulate), "r"(step)
      : "memory");
}

__device__ __forceinline__ bool elect_one_sync()
{
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

// arrives (count 1) on `bar` when every tcgen05 operation this thread issued so far has completed
__device__ __forceinline__ void umma_commit(uint64_t* bar)
{
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

}  // namespace dfk
