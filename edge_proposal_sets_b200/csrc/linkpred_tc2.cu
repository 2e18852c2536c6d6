// K2 (tensor-core arm, CTA-pair variant) — tcgen05.mma.cta_group::2 over a cluster of two CTAs.
//
// Same math as linkpred_tc.cu (replaces /root/reference/models.py:478-485,506).  Why a CTA pair:
// with H = 256 one hidden layer's bf16 weights are 128 KB, so a 3-layer LinkPredictor (two hidden
// GEMMs: collab / ppa / twitch / fb defaults, models.py:712-744) cannot keep both matrices in one
// SM's 227 KB next to the activation tile.  cta_group::2 splits the N (output-feature) dimension of
// the B operand across the two SMs of a TPC: each CTA keeps rows [r*H/2, (r+1)*H/2) of EVERY hidden
// layer's weights resident for the whole kernel (2 x 64 KB), gathers its own 128 candidate pairs,
// and one thread of the leader CTA issues M = 256 MMAs that read both SMs' operands and write each
// SM's 128 x H fp32 accumulator into its own TMEM.  No weight bytes move after the prologue.
#include <cooperative_groups.h>
#include <stdlib.h>

#include "tc_common.cuh"

namespace eps {

namespace cg = cooperative_groups;

__device__ __forceinline__ void umma_bf16_ss_2cta(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                                  uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}\n"
      :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u) : "memory");
}
__device__ __forceinline__ void umma_commit_2cta(uint32_t mbar_saddr) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      :: "r"(mbar_saddr), "h"((uint16_t)3) : "memory");
}

// hidden-layer weights -> per-CTA-half bf16 SWIZZLE_128B images: [layer][half][kblock][H/2 rows][128 B]
__global__ void pack_weights_halves_kernel(MlpParams prm, int H, int nhidden, uint8_t *img) {
  const int chunks_per_row = H / 8, HH = H / 2;
  const int total = nhidden * H * chunks_per_row;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int l = i / (H * chunks_per_row);
    const int rem = i - l * H * chunks_per_row;
    const int n = rem / chunks_per_row, c = rem - n * chunks_per_row;
    const float *w = prm.W[l] + (size_t)n * H + c * 8;
    const float4 a = *reinterpret_cast<const float4 *>(w), b = *reinterpret_cast<const float4 *>(w + 4);
    uint4 o;
    o.x = pack_bf16x2(a.x, a.y); o.y = pack_bf16x2(a.z, a.w);
    o.z = pack_bf16x2(b.x, b.y); o.w = pack_bf16x2(b.z, b.w);
    const int half = n / HH, nn = n - half * HH;
    *reinterpret_cast<uint4 *>(img + (size_t)l * H * H * 2 + (size_t)half * HH * H * 2 +
                               sw128_chunk_off(HH, nn, c * 8)) = o;
  }
}

template <int H>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC_THREADS, 1)
linkpred_tc2_kernel(const float *__restrict__ h, const int *__restrict__ pu, const int *__restrict__ pv,
                    long long M, const MlpParams prm, int L, int apply_sigmoid,
                    const uint8_t *__restrict__ wimg, float *__restrict__ score) {
  static_assert(H % 64 == 0 && H >= 64 && H <= 256, "H in {64,128,192,256}");
  constexpr int HH = H / 2;
  constexpr int A_BYTES = TC_BM * H * 2;
  constexpr int WH_BYTES = HH * H * 2;               // one layer, this CTA's half
  constexpr uint32_t TMEM_COLS = H <= 64 ? 64 : (H <= 128 ? 128 : 256);
  constexpr uint32_t IDESC = umma_idesc_bf16(2 * TC_BM, H);   // M = 256 across the pair
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t *sA = smem;
  uint8_t *sW = smem + A_BYTES;                       // [nhidden][WH_BYTES]
  const int nhidden = L - 1;
  float *sBias = reinterpret_cast<float *>(sW + (size_t)nhidden * WH_BYTES);   // [nhidden][H]
  float *sWlast = sBias + nhidden * H;                // [H]
  float *sPart = sWlast + H;                          // [2][128]
  __shared__ __align__(8) uint64_t mbar_mma;
  __shared__ uint32_t tmem_base_slot;
  cg::cluster_group cluster = cg::this_cluster();
  const uint32_t cta_rank = cluster.block_rank();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;"
                 :: "r"(smem_u32(&tmem_base_slot)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    mbar_init(smem_u32(&mbar_mma), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // resident operands: this CTA's half of every hidden layer, biases, output layer
  for (int l = 0; l < nhidden; ++l) {
    const uint4 *src = reinterpret_cast<const uint4 *>(wimg + (size_t)l * H * H * 2 + (size_t)cta_rank * WH_BYTES);
    uint4 *dst = reinterpret_cast<uint4 *>(sW + (size_t)l * WH_BYTES);
    for (int i = tid; i < WH_BYTES / 16; i += TC_THREADS) dst[i] = __ldg(src + i);
  }
  for (int i = tid; i < nhidden * H; i += TC_THREADS) sBias[i] = __ldg(prm.b[i / H] + (i % H));
  for (int i = tid; i < H; i += TC_THREADS) sWlast[i] = __ldg(prm.W[L - 1] + i);
  const float b_last = __ldg(prm.b[L - 1]);
  fence_async_smem();
  tc_fence_before();
  cluster.sync();
  tc_fence_after();
  const uint32_t tmem_acc = tmem_base_slot;
  const uint32_t sA_addr = smem_u32(sA), sW_addr = smem_u32(sW), mbar_addr = smem_u32(&mbar_mma);
  uint32_t phase = 0;

  const long long npair_tiles = (M + 2 * TC_BM - 1) / (2 * TC_BM);
  const long long nclusters = gridDim.x / 2, cluster_id = blockIdx.x / 2;
  for (long long tile = cluster_id; tile < npair_tiles; tile += nclusters) {
    const long long p0 = tile * (2 * TC_BM) + (long long)cta_rank * TC_BM;
    const int rows = (int)max(0ll, min((long long)TC_BM, M - p0));
    // ---- gather + Hadamard -> bf16 A tile (this CTA's 128 pairs) ----
    for (int r = warp; r < TC_BM; r += TC_THREADS / 32) {
      if (r < rows) {
        const float *hu = h + (size_t)__ldg(pu + p0 + r) * H;
        const float *hv = h + (size_t)__ldg(pv + p0 + r) * H;
        for (int c = lane; c < H / 8; c += 32) {
          const float4 a0 = __ldg(reinterpret_cast<const float4 *>(hu) + 2 * c);
          const float4 a1 = __ldg(reinterpret_cast<const float4 *>(hu) + 2 * c + 1);
          const float4 b0 = __ldg(reinterpret_cast<const float4 *>(hv) + 2 * c);
          const float4 b1 = __ldg(reinterpret_cast<const float4 *>(hv) + 2 * c + 1);
          uint4 o;
          o.x = hadamard_bf16x2(a0.x, a0.y, b0.x, b0.y); o.y = hadamard_bf16x2(a0.z, a0.w, b0.z, b0.w);
          o.z = hadamard_bf16x2(a1.x, a1.y, b1.x, b1.y); o.w = hadamard_bf16x2(a1.z, a1.w, b1.z, b1.w);
          *reinterpret_cast<uint4 *>(sA + sw128_chunk_off(TC_BM, r, c * 8)) = o;
        }
      } else {
        for (int c = lane; c < H / 8; c += 32)
          *reinterpret_cast<uint4 *>(sA + sw128_chunk_off(TC_BM, r, c * 8)) = make_uint4(0, 0, 0, 0);
      }
    }
    float part = 0.f;
    for (int l = 0; l < nhidden; ++l) {
      fence_async_smem();   // A tile (generic-proxy stores) -> visible to the tensor cores of both SMs
      tc_fence_before();    // this thread's TMEM reads of the previous layer are complete
      cluster.sync();       // both CTAs' A tiles are in place, both accumulators are free
      if (cta_rank == 0 && tid == 0) {
        tc_fence_after();
#pragma unroll
        for (int kb = 0; kb < H / 64; ++kb) {
#pragma unroll
          for (int k16 = 0; k16 < 4; ++k16) {
            const uint64_t ad = umma_smem_desc(sA_addr + kb * (TC_BM * 128) + k16 * 32);
            const uint64_t bd = umma_smem_desc(sW_addr + l * WH_BYTES + kb * (HH * 128) + k16 * 32);
            umma_bf16_ss_2cta(tmem_acc, ad, bd, IDESC, (kb | k16) ? 1u : 0u);
          }
        }
        umma_commit_2cta(mbar_addr);   // arrives on mbar_mma of BOTH CTAs when the MMAs retire
      }
      mbar_wait(mbar_addr, phase);
      phase ^= 1;
      tc_fence_after();
      const int row = (warp & 3) * 32 + lane;
      const int chalf = warp >> 2;
      const bool last_hidden = l == nhidden - 1;
      const float *bias = sBias + l * H;
#pragma unroll 1
      for (int c0 = chalf * (H / 2); c0 < (chalf + 1) * (H / 2); c0 += 32) {
        float v[32];
        tmem_ld32(tmem_acc + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)c0, v);
        if (last_hidden) {
#pragma unroll
          for (int j = 0; j < 32; ++j) part = fmaf(fmaxf(v[j] + bias[c0 + j], 0.f), sWlast[c0 + j], part);
        } else {
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            uint4 o;
            o.x = pack_bf16x2(fmaxf(v[j + 0] + bias[c0 + j + 0], 0.f), fmaxf(v[j + 1] + bias[c0 + j + 1], 0.f));
            o.y = pack_bf16x2(fmaxf(v[j + 2] + bias[c0 + j + 2], 0.f), fmaxf(v[j + 3] + bias[c0 + j + 3], 0.f));
            o.z = pack_bf16x2(fmaxf(v[j + 4] + bias[c0 + j + 4], 0.f), fmaxf(v[j + 5] + bias[c0 + j + 5], 0.f));
            o.w = pack_bf16x2(fmaxf(v[j + 6] + bias[c0 + j + 6], 0.f), fmaxf(v[j + 7] + bias[c0 + j + 7], 0.f));
            *reinterpret_cast<uint4 *>(sA + sw128_chunk_off(TC_BM, row, c0 + j)) = o;
          }
        }
      }
      if (last_hidden) sPart[chalf * TC_BM + row] = part;
    }
    __syncthreads();
    if (tid < TC_BM && tid < rows) {
      float s = sPart[tid] + sPart[TC_BM + tid] + b_last;
      score[p0 + tid] = apply_sigmoid ? sigmoidf_ref(s) : s;
    }
    // the next tile's gather overwrites sA: the last layer's MMAs (which read it) have retired
    // (mbar wait above) in BOTH CTAs only after the next cluster.sync(); the peer may still be
    // waiting on its copy of the barrier, but it never reads this CTA's sA outside an MMA.
  }
  tc_fence_before();
  cluster.sync();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" :: "r"(tmem_acc), "r"(TMEM_COLS) : "memory");
  }
}

template <int H>
static int tc2_launch_h(const float *h, const int *pu, const int *pv, long long M, const MlpParams &prm, int L,
                        int apply_sigmoid, float *score, uint8_t *img, cudaStream_t stream) {
  const int nhidden = L - 1;
  const size_t smem = 1024 + (size_t)TC_BM * H * 2 + (size_t)nhidden * (H / 2) * H * 2 +
                      sizeof(float) * ((size_t)nhidden * H + H + 2 * TC_BM);
  if (smem > 227 * 1024) {
    set_error("eps_linkpred_mlp: %d hidden layers of H=%d do not fit the CTA-pair kernel (%zu B smem)", nhidden, H, smem);
    return EPS_ERR_UNSUPPORTED;
  }
  auto kern = linkpred_tc2_kernel<H>;
  EPS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long npair_tiles = (M + 2 * TC_BM - 1) / (2 * TC_BM);
  const int clusters = (int)std::min<long long>(npair_tiles, (long long)(sm_count() / 2));
  kern<<<2 * clusters, TC_THREADS, smem, stream>>>(h, pu, pv, M, prm, L, apply_sigmoid, img, score);
  EPS_LAUNCH_CHECK();
  return EPS_OK;
}

int linkpred_tc2_launch(const float *h, const void *h_bf16, int H, const int *pu, const int *pv, long long M,
                        const MlpParams &prm, int L, int apply_sigmoid, float *score, uint8_t *img, int n,
                        int *tile_order, bool prepared, cudaStream_t stream) {
  const int total = (L - 1) * H * (H / 8);
  if (!prepared) {                                  // EPS_MLP_REUSE_WORKSPACE: the images of an earlier call are still there
    pack_weights_halves_kernel<<<(total + 255) / 256, 256, 0, stream>>>(prm, H, L - 1, img);
    EPS_LAUNCH_CHECK();
  }
  const char *variant = getenv("EPS_TC_VARIANT");   // "2": force this (unpipelined) kernel
  if (!(variant && variant[0] == '2')) {
    const int st = h_bf16 ? linkpred_tc3_launch(h_bf16, 1, H, pu, pv, M, prm, L, apply_sigmoid, score, img, n,
                                                tile_order, stream)
                          : linkpred_tc3_launch(h, 0, H, pu, pv, M, prm, L, apply_sigmoid, score, img, n,
                                                tile_order, stream);
    if (st != EPS_ERR_UNSUPPORTED) return st;       // pipelined kernel ran (or failed for real)
  }
  if (H == 64) return tc2_launch_h<64>(h, pu, pv, M, prm, L, apply_sigmoid, score, img, stream);
  if (H == 128) return tc2_launch_h<128>(h, pu, pv, M, prm, L, apply_sigmoid, score, img, stream);
  return tc2_launch_h<256>(h, pu, pv, M, prm, L, apply_sigmoid, score, img, stream);
}

}  // namespace eps
