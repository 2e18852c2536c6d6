// K2 (tensor-core arm) — fused gather + Hadamard + LinkPredictor MLP on tcgen05 / TMEM.
//
// Replaces /root/reference/models.py:506 (two row gathers) + models.py:478-485 (LinkPredictor),
// i.e. index_select x2, mul, (cuBLAS SGEMM + bias + relu) x (L-1), GEMV, sigmoid — 2L+3 launches
// and 2*B*H*4 bytes of materialised gathers per batch in the reference.
//
// One CTA owns a tile of 128 candidate pairs (UMMA M = 128, cta_group::1):
//   gather    h[u], h[v] rows (fp32, 128-bit loads), multiply in fp32, round to bf16 and store the
//             A tile straight into shared memory in the UMMA canonical K-major SWIZZLE_128B layout;
//   MMA       one elected thread issues tcgen05.mma.kind::f16 (bf16 x bf16 -> fp32) over K = H in
//             steps of 16 against the layer's weight matrix, which sits in shared memory as a
//             pre-swizzled bf16 image (packed once per call by pack_weights_kernel); the 128 x H
//             fp32 accumulator lives in TMEM;
//   epilogue  every thread owns one accumulator row (TMEM lane): tcgen05.ld 32 columns at a time,
//             + bias, ReLU, then either bf16 -> shared memory as the next layer's A operand, or —
//             for the last hidden layer — the H -> 1 output layer as a running dot product in
//             registers, + bias, sigmoid, one coalesced 4-byte store per pair.
// Arithmetic: bf16 operands, fp32 products/accumulation/bias/sigmoid.  Tolerance vs the fp32
// reference path is stated in tests/test_gpu_mlp_tc.py and DESIGN.md.
#include <stdlib.h>

#include "tc_common.cuh"

namespace eps {

// fp32 [out,in] weights of the hidden layers -> bf16 swizzled images, one H*H*2-byte image per layer
__global__ void pack_weights_kernel(MlpParams prm, int H, int nhidden, uint8_t *img) {
  const int chunks_per_row = H / 8;
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
    *reinterpret_cast<uint4 *>(img + (size_t)l * H * H * 2 + sw128_chunk_off(H, n, c * 8)) = o;
  }
}

template <int H>
__global__ void __launch_bounds__(TC_THREADS, 1)
linkpred_tc_kernel(const float *__restrict__ h, const int *__restrict__ pu, const int *__restrict__ pv,
                   long long M, const MlpParams prm, int L, int apply_sigmoid,
                   const uint8_t *__restrict__ wimg, float *__restrict__ score) {
  static_assert(H % 64 == 0 && H >= 64 && H <= 256, "tensor-core arm: H in {64,128,192,256}");
  constexpr int A_BYTES = TC_BM * H * 2;
  constexpr int W_BYTES = H * H * 2;
  constexpr uint32_t TMEM_COLS = H <= 64 ? 64 : (H <= 128 ? 128 : 256);
  constexpr uint32_t IDESC = umma_idesc_bf16(TC_BM, H);
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t *sA = smem;
  uint8_t *sW = smem + A_BYTES;
  float *sBias = reinterpret_cast<float *>(smem + A_BYTES + W_BYTES);  // [(L-1)][H]
  float *sWlast = sBias + (EPS_MAX_MLP_LAYERS - 1) * H;                // [H]
  float *sPart = sWlast + H;                                           // [2][128]
  __shared__ __align__(8) uint64_t mbar_mma;
  __shared__ uint32_t tmem_base_slot;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nhidden = L - 1;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 :: "r"(smem_u32(&tmem_base_slot)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    mbar_init(smem_u32(&mbar_mma), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = tid; i < nhidden * H; i += TC_THREADS) sBias[i] = __ldg(prm.b[i / H] + (i % H));
  for (int i = tid; i < H; i += TC_THREADS) sWlast[i] = __ldg(prm.W[L - 1] + i);
  const float b_last = __ldg(prm.b[L - 1]);
  int resident_layer = -1;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_acc = tmem_base_slot;
  const uint32_t sA_addr = smem_u32(sA), sW_addr = smem_u32(sW), mbar_addr = smem_u32(&mbar_mma);
  uint32_t phase = 0;

  const long long ntiles = (M + TC_BM - 1) / TC_BM;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long p0 = tile * TC_BM;
    const int rows = (int)min((long long)TC_BM, M - p0);
    // ---- gather + Hadamard -> bf16 A tile (swizzled) ----
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
      // ---- weights of layer l -> shared memory (stay resident while only one hidden layer) ----
      if (resident_layer != l) {
        const uint4 *src = reinterpret_cast<const uint4 *>(wimg + (size_t)l * W_BYTES);
        uint4 *dst = reinterpret_cast<uint4 *>(sW);
        for (int i = tid; i < W_BYTES / 16; i += TC_THREADS) dst[i] = __ldg(src + i);
        resident_layer = l;
      }
      fence_async_smem();   // generic-proxy smem writes (A tile, weights) -> visible to the tensor core
      __syncthreads();
      if (tid == 0) {
        tc_fence_after();
#pragma unroll
        for (int kb = 0; kb < H / 64; ++kb) {
#pragma unroll
          for (int k16 = 0; k16 < 4; ++k16) {
            const uint64_t ad = umma_smem_desc(sA_addr + kb * (TC_BM * 128) + k16 * 32);
            const uint64_t bd = umma_smem_desc(sW_addr + kb * (H * 128) + k16 * 32);
            umma_bf16_ss(tmem_acc, ad, bd, IDESC, (kb | k16) ? 1u : 0u);
          }
        }
        umma_commit(mbar_addr);   // implies tcgen05.fence::before_thread_sync
      }
      mbar_wait(mbar_addr, phase);
      phase ^= 1;
      tc_fence_after();
      // ---- epilogue: thread owns accumulator row (lane quadrant of its warp), half of the columns ----
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
      tc_fence_before();   // TMEM reads done before the next MMA (after the barrier) overwrites
      if (last_hidden) sPart[chalf * TC_BM + row] = part;
    }
    __syncthreads();
    if (tid < TC_BM && tid < rows) {
      float s = sPart[tid] + sPart[TC_BM + tid] + b_last;
      score[p0 + tid] = apply_sigmoid ? sigmoidf_ref(s) : s;
    }
    __syncthreads();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_acc), "r"(TMEM_COLS) : "memory");
  }
}

// The pipelined kernel gathers from a bf16 copy of h (half the bytes per pair) when the pair list
// is long enough to amortise writing it; short lists read the caller's fp32 rows and round them in
// registers — the SAME rounding, so the scores do not depend on this choice.
bool linkpred_tc_uses_table(int n, long long M) { return M >= 2ll * n; }

static size_t tc_img_bytes(int H, int L) {
  return ((size_t)std::max(L - 1, 0) * H * H * 2 + 255) & ~(size_t)255;
}

size_t linkpred_tc_workspace_bytes(int n, int H, int L, long long M) {
  // weight images | bf16 embedding table | tile schedule of the pipelined kernel (one int per 256 pairs)
  return 256 + tc_img_bytes(H, L) + (linkpred_tc_uses_table(n, M) ? (((size_t)n * H * 2 + 255) & ~(size_t)255) : 0) +
         (size_t)((M + 255) / 256) * 4 + 256;
}

template <int H>
static int tc_launch_h(const float *h, const int *pu, const int *pv, long long M, const MlpParams &prm, int L,
                       int apply_sigmoid, float *score, uint8_t *img, cudaStream_t stream) {
  const size_t smem = 1024 + (size_t)TC_BM * H * 2 + (size_t)H * H * 2 +
                      sizeof(float) * ((EPS_MAX_MLP_LAYERS - 1) * H + H + 2 * TC_BM);
  auto kern = linkpred_tc_kernel<H>;
  EPS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long ntiles = (M + TC_BM - 1) / TC_BM;
  const int grid = (int)std::min<long long>(ntiles, (long long)sm_count());
  kern<<<grid, TC_THREADS, smem, stream>>>(h, pu, pv, M, prm, L, apply_sigmoid, img, score);
  EPS_LAUNCH_CHECK();
  return EPS_OK;
}

int linkpred_tc_launch(const float *h, int n, int H, const int *pu, const int *pv, long long M,
                       const MlpParams &prm, int L, int apply_sigmoid, float *score, void *workspace,
                       size_t workspace_bytes, bool prepared, cudaStream_t stream) {
  if (L < 2 || !(H == 64 || H == 128 || H == 256)) {
    set_error("eps_linkpred_mlp: the tcgen05 arm needs num_layers >= 2 and H in {64,128,256} (got L=%d H=%d); "
              "use EPS_MLP_FP32", L, H);
    return EPS_ERR_UNSUPPORTED;
  }
  if (!workspace || workspace_bytes < linkpred_tc_workspace_bytes(n, H, L, M)) {
    set_error("eps_linkpred_mlp: workspace too small");
    return EPS_ERR_WORKSPACE;
  }
  uint8_t *img = (uint8_t *)workspace + 256;
  // CTA-pair kernel (all hidden-layer weights resident) unless EPS_TC_VARIANT=1 asks for the
  // single-CTA kernel (kept for A/B measurements)
  const char *variant = getenv("EPS_TC_VARIANT");
  if (!(variant && variant[0] == '1') && sm_count() >= 2) {
    void *table = nullptr;
    uint8_t *tail = img + ((tc_img_bytes(H, L) + 255) & ~(size_t)255);
    if (linkpred_tc_uses_table(n, M) && !(variant && variant[0] == '2')) {
      table = tail;
      tail += ((size_t)n * H * 2 + 255) & ~(size_t)255;
      if (!prepared) {                              // EPS_MLP_REUSE_WORKSPACE: the table of an earlier call is still there
        const int st = h_to_bf16_launch(h, (long long)n * H, table, stream);
        if (st != EPS_OK) return st;
      }
    }
    return linkpred_tc2_launch(h, table, H, pu, pv, M, prm, L, apply_sigmoid, score, img, n, (int *)tail, prepared, stream);
  }
  const int total = (L - 1) * H * (H / 8);
  pack_weights_kernel<<<(total + 255) / 256, 256, 0, stream>>>(prm, H, L - 1, img);
  EPS_LAUNCH_CHECK();
  if (H == 64) return tc_launch_h<64>(h, pu, pv, M, prm, L, apply_sigmoid, score, img, stream);
  if (H == 128) return tc_launch_h<128>(h, pu, pv, M, prm, L, apply_sigmoid, score, img, stream);
  return tc_launch_h<256>(h, pu, pv, M, prm, L, apply_sigmoid, score, img, stream);
}

}  // namespace eps
