// tcgen05 / TMEM / mbarrier helpers shared by the tensor-core LinkPredictor kernels (sm_100a).
#pragma once
#include <cuda_bf16.h>

#include "eps_common.cuh"

namespace eps {

constexpr int TC_THREADS = 256;
constexpr int TC_BM = 128;

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}

// ---- UMMA descriptors (cute/arch/mma_sm100_desc.hpp field layout) -------------------------
// shared-memory matrix descriptor, K-major, SWIZZLE_128B: 8-row groups 1024 B apart
__device__ __forceinline__ uint64_t umma_smem_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);        // start address      bits [0,14)
  d |= (uint64_t)1 << 16;                          // leading byte off.  bits [16,30)  (unused for SW128 K-major)
  d |= (uint64_t)(1024u >> 4) << 32;               // stride byte off.   bits [32,46)
  d |= (uint64_t)1 << 46;                          // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                          // layout type SWIZZLE_128B
  return d;
}
// instruction descriptor, kind::f16: D = f32, A = B = bf16, both K-major, M x N tile
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void umma_bf16_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                             uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
      :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t mbar_saddr) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
               :: "r"(mbar_saddr) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void mbar_init(uint32_t saddr, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(saddr), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t saddr, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred P1;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "DONE:\n\t}\n"
      :: "r"(saddr), "r"(parity) : "memory");
}

// 32 consecutive fp32 accumulator columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float *v) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n\t"
      "tcgen05.wait::ld.sync.aligned;\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 b = __floats2bfloat162_rn(lo, hi);  // .x = lo (low 16 bits)
  return *reinterpret_cast<uint32_t *>(&b);
}

// A-operand element of the first layer: bf16( bf16(h_u) * bf16(h_v) ).  The embeddings are rounded
// to bf16 FIRST (the hot path gathers them from a bf16 copy of h: half the bytes per pair), then
// multiplied; the product of two bf16 values is exact in fp32, so HMUL2.BF16 rounds exactly once.
// Every tensor-core kernel (tc / tc2 / tc3, fp32 or bf16 source) uses this, so a score does not
// depend on which kernel or which source format produced it.
__device__ __forceinline__ uint32_t mul_bf16x2(uint32_t a, uint32_t b) {
  __nv_bfloat162 r = __hmul2(*reinterpret_cast<const __nv_bfloat162 *>(&a), *reinterpret_cast<const __nv_bfloat162 *>(&b));
  return *reinterpret_cast<uint32_t *>(&r);
}
__device__ __forceinline__ uint32_t hadamard_bf16x2(float a0, float a1, float b0, float b1) {
  return mul_bf16x2(pack_bf16x2(a0, a1), pack_bf16x2(b0, b1));
}

// byte offset of the 16-byte chunk holding elements k..k+7 (k % 8 == 0) of row r inside a K-major
// SWIZZLE_128B tile with `rows` rows: 64-element K blocks, 128-byte rows, chunk index ^= r & 7
__host__ __device__ __forceinline__ uint32_t sw128_chunk_off(int rows, int r, int k) {
  const int kb = k >> 6, chunk = (k & 63) >> 3;
  return (uint32_t)kb * (uint32_t)rows * 128u + (uint32_t)r * 128u + (uint32_t)((chunk ^ (r & 7)) << 4);
}


}  // namespace eps
