// tcgen05 / TMEM / mbarrier helpers shared by the tensor-core LinkPredictor kernels (sm_100a).
#pragma once
#include <cuda_fp16.h>

#include "eps_common.cuh"

namespace eps {

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
// instruction descriptor, kind::f16: D = f32 (bits [4,6) = 1), A = B = fp16 (a_format bits [7,10) = 0,
// b_format bits [10,13) = 0; bf16 would be 1), both K-major, M x N tile
__host__ __device__ constexpr uint32_t umma_idesc_f16(int M, int N) {
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void mbar_init(uint32_t saddr, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(saddr), "r"(count) : "memory");
}
// ---- operand format of the tensor-core arm: IEEE fp16 (11-bit significand), fp32 accumulate ----
// fp16 runs at the same tcgen05 rate as bf16 and carries three more significand bits: measured on the ppa-shape
// filter model the score deviation from the fp32 arm drops 8x, which is what lets the tensor-core scores serve
// as a PREFILTER for the exact fp32 list (filter_step.py).  fp16's narrow range is handled by a power-of-two
// scale computed on the device from worst-case bounds (tc_scale_kernel): every operand — embedding products,
// hidden activations, bias terms — is carried multiplied by S = hscale^2, so nothing can overflow, and ReLU's
// positive homogeneity makes the scale drop out exactly at the output layer.
struct TcScale {      // written by tc_scale_kernel, read by every kernel of the arm
  float hscale;       // 2^a: embedding rows are converted as fp16(h * hscale)
  float S;            // hscale^2: scale of the Hadamard products, the biases and all hidden activations
  float invS;         // 1 / S, folded into the output layer's weights
  float hmax;         // max |h| the scale was derived from (diagnostic)
};

__device__ __forceinline__ uint32_t pack_f16x2(float lo, float hi) {
  uint32_t d;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));   // .x = lo (low 16 bits)
  return d;
}

// A-operand element of the first layer: fp16( fp16(h_u * hscale) * fp16(h_v * hscale) ).  The embeddings are
// rounded to fp16 FIRST (the hot path gathers them from an fp16 copy of h: half the bytes per pair), then
// multiplied with one rounding (HMUL2).  Every path of the arm (fp16 table or fp32 source) uses this, so a
// score does not depend on which source format produced it.
__device__ __forceinline__ uint32_t mul_f16x2(uint32_t a, uint32_t b) {
  __half2 r = __hmul2(*reinterpret_cast<const __half2 *>(&a), *reinterpret_cast<const __half2 *>(&b));
  return *reinterpret_cast<uint32_t *>(&r);
}
__device__ __forceinline__ uint32_t hadamard_f16x2(float a0, float a1, float b0, float b1, float hscale) {
  return mul_f16x2(pack_f16x2(a0 * hscale, a1 * hscale), pack_f16x2(b0 * hscale, b1 * hscale));
}

// byte offset of the 16-byte chunk holding elements k..k+7 (k % 8 == 0) of row r inside a K-major
// SWIZZLE_128B tile with `rows` rows: 64-element K blocks, 128-byte rows, chunk index ^= r & 7
__host__ __device__ __forceinline__ uint32_t sw128_chunk_off(int rows, int r, int k) {
  const int kb = k >> 6, chunk = (k & 63) >> 3;
  return (uint32_t)kb * (uint32_t)rows * 128u + (uint32_t)r * 128u + (uint32_t)((chunk ^ (r & 7)) << 4);
}


}  // namespace eps
