// K2 (tensor-core arm, pipelined CTA-pair kernel) — the production variant.
//
// Same math as linkpred_tc.cu / linkpred_tc2.cu (replaces /root/reference/models.py:478-485,506).
// A cluster of two CTAs (tcgen05 cta_group::2) keeps its halves of ALL hidden-layer weights resident
// in shared memory and runs three warp-specialised roles per CTA, connected by mbarriers:
//
//   producers (8 warps)  gather h[u], h[v] (fp32, 128-bit loads, next chunk prefetched in registers),
//                        multiply, round to bf16 and fill a 3-stage ring of 128 x 32 K-chunks
//                        (K-major SWIZZLE_64B) for the first layer;
//   MMA issuer (1 lane,  waits for a ring stage from BOTH CTAs, issues M=256 x N=H x K=16 UMMAs
//   leader CTA only)     into one of two TMEM accumulator slots, releases stages / publishes
//                        accumulators with multicast tcgen05.commit; later layers read their A operand
//                        (the previous layer's activations) from a resident 128 x H bf16 tile;
//   epilogue (4 warps)   thread-per-row: tcgen05.ld, + bias, ReLU, then either bf16 -> the activation
//                        tile for the next layer, or the fused H -> 1 output layer + sigmoid.
//
// The two accumulator slots let the tensor pipe start the next GEMM (next layer, or next tile's first
// layer) while the epilogue drains the previous one; the ring decouples the HBM/L2 gather from both.
#include <cooperative_groups.h>
#include <stdlib.h>

#include "tc_common.cuh"

namespace eps {

namespace cg = cooperative_groups;

constexpr int P_EPI_WARPS = 4;
constexpr int P_RING = 3;                                          // ring stages == producer groups
constexpr int P_GROUP_WARPS = 4;                                   // warps per producer group
constexpr int P_PROD_WARPS = P_RING * P_GROUP_WARPS;               // 12
constexpr int P_IDS_WARP = P_EPI_WARPS + 1;                        // warp 5 prefetches pair ids
constexpr int P_FIRST_PROD_WARP = P_EPI_WARPS + 2;
constexpr int P_THREADS = (P_EPI_WARPS + 2 + P_PROD_WARPS) * 32;   // 576
constexpr int P_CHUNK_K = 32;
constexpr int P_STAGE_BYTES = TC_BM * P_CHUNK_K * 2;               // 8 KB

__device__ __forceinline__ uint64_t umma_smem_desc_sw64(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(512u >> 4) << 32;                // 8 rows x 64 B
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)4 << 61;                          // SWIZZLE_64B
  return d;
}
__device__ __forceinline__ uint32_t sw64_chunk_off(int r, int sub) {
  return (uint32_t)r * 64u + (uint32_t)((sub ^ ((r >> 1) & 3)) << 4);
}
__device__ __forceinline__ void umma_bf16_ss_2cta_p(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                                    uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}\n"
      :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u) : "memory");
}
__device__ __forceinline__ void umma_commit_mc(uint32_t mbar_saddr) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      :: "r"(mbar_saddr), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void mbar_arrive_on_cta(uint32_t local_saddr, uint32_t target_cta) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_saddr), "r"(target_cta));
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" :: "r"(r) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t saddr, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred P1;\n\t"
      "LAB_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE;\n\t"
      "bra LAB_WAIT;\n\t"
      "DONE:\n\t}\n"
      :: "r"(saddr), "r"(parity) : "memory");
}

struct PipeBarriers {
  uint64_t full[P_RING];     // producers (both CTAs) -> MMA issuer       (waited in the leader)
  uint64_t empty[P_RING];    // MMA commit -> producers                    (multicast, both CTAs)
  uint64_t acc_full[2];      // MMA commit -> epilogue                     (multicast, both CTAs)
  uint64_t acc_free[2];      // epilogue (both CTAs) -> MMA issuer         (waited in the leader)
  uint64_t a2_full;          // epilogue (both CTAs) -> MMA issuer         (waited in the leader)
  uint64_t ids_full[2];      // ids warp -> producers                      (CTA-local)
  uint64_t ids_empty[2];     // producers -> ids warp                      (CTA-local)
};

template <int H>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(P_THREADS, 1)
linkpred_tc3_kernel(const float *__restrict__ h, const int *__restrict__ pu, const int *__restrict__ pv,
                    long long M, const MlpParams prm, int L, int apply_sigmoid,
                    const uint8_t *__restrict__ wimg, float *__restrict__ score, int tune) {
  static_assert(H % 64 == 0 && H >= 64 && H <= 256, "H in {64,128,192,256}");
  constexpr int HH = H / 2;
  constexpr int WH_BYTES = HH * H * 2;
  constexpr int A2_BYTES = TC_BM * H * 2;
  constexpr int NCHUNK = H / P_CHUNK_K;
  constexpr uint32_t TMEM_COLS = 2 * H <= 128 ? 128 : (2 * H <= 256 ? 256 : 512);
  constexpr uint32_t IDESC = umma_idesc_bf16(2 * TC_BM, H);
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int nhidden = L - 1;
  uint8_t *sW = smem;                                                   // [nhidden][WH_BYTES]
  uint8_t *sRing = sW + (size_t)nhidden * WH_BYTES;                     // [P_RING][8 KB]
  uint8_t *sA2 = sRing + P_RING * P_STAGE_BYTES;                        // 128 x H bf16 (nhidden >= 2)
  float *sBias = reinterpret_cast<float *>(sA2 + (nhidden >= 2 ? A2_BYTES : 0));
  float *sWlast = sBias + nhidden * H;
  int2 *sIds = reinterpret_cast<int2 *>(sWlast + H);                    // [2][128] (u, v) of this CTA's rows
  __shared__ __align__(8) PipeBarriers bars;
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
    for (int i = 0; i < P_RING; ++i) {
      mbar_init(smem_u32(&bars.full[i]), 2 * P_GROUP_WARPS);
      mbar_init(smem_u32(&bars.empty[i]), 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&bars.acc_full[i]), 1);
      mbar_init(smem_u32(&bars.acc_free[i]), 2 * P_EPI_WARPS);
    }
    mbar_init(smem_u32(&bars.a2_full), 2 * P_EPI_WARPS);
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&bars.ids_full[i]), 1);
      mbar_init(smem_u32(&bars.ids_empty[i]), P_PROD_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int l = 0; l < nhidden; ++l) {
    const uint4 *src = reinterpret_cast<const uint4 *>(wimg + (size_t)l * H * H * 2 + (size_t)cta_rank * WH_BYTES);
    uint4 *dst = reinterpret_cast<uint4 *>(sW + (size_t)l * WH_BYTES);
    for (int i = tid; i < WH_BYTES / 16; i += P_THREADS) dst[i] = __ldg(src + i);
  }
  for (int i = tid; i < nhidden * H; i += P_THREADS) sBias[i] = __ldg(prm.b[i / H] + (i % H));
  for (int i = tid; i < H; i += P_THREADS) sWlast[i] = __ldg(prm.W[L - 1] + i);
  const float b_last = __ldg(prm.b[L - 1]);
  fence_async_smem();
  tc_fence_before();
  cluster.sync();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  const long long npair_tiles = (M + 2 * TC_BM - 1) / (2 * TC_BM);
  const long long nclusters = gridDim.x / 2, cluster_id = blockIdx.x / 2;

  if (warp < P_EPI_WARPS) {
    // =============================== EPILOGUE ===============================
    uint32_t acph = 0;            // bit s = phase parity of acc_full[s] (kept in a register)
    uint32_t seq = 0;
    const int row = warp * 32 + lane;
    for (long long tile = cluster_id; tile < npair_tiles; tile += nclusters) {
      const long long p0 = tile * (2 * TC_BM) + (long long)cta_rank * TC_BM;
      for (int l = 0; l < nhidden; ++l, ++seq) {
        const uint32_t slot = seq & 1;
        mbar_wait_cluster(smem_u32(&bars.acc_full[slot]), (acph >> slot) & 1u);
        acph ^= 1u << slot;
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + slot * H;
        const float *bias = sBias + l * H;
        if (l < nhidden - 1) {
#pragma unroll 1
          for (int c0 = 0; c0 < H; c0 += 32) {
            float v[32];
            tmem_ld32(taddr + (uint32_t)c0, v);
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              uint4 o;
              o.x = pack_bf16x2(fmaxf(v[j + 0] + bias[c0 + j + 0], 0.f), fmaxf(v[j + 1] + bias[c0 + j + 1], 0.f));
              o.y = pack_bf16x2(fmaxf(v[j + 2] + bias[c0 + j + 2], 0.f), fmaxf(v[j + 3] + bias[c0 + j + 3], 0.f));
              o.z = pack_bf16x2(fmaxf(v[j + 4] + bias[c0 + j + 4], 0.f), fmaxf(v[j + 5] + bias[c0 + j + 5], 0.f));
              o.w = pack_bf16x2(fmaxf(v[j + 6] + bias[c0 + j + 6], 0.f), fmaxf(v[j + 7] + bias[c0 + j + 7], 0.f));
              *reinterpret_cast<uint4 *>(sA2 + sw128_chunk_off(TC_BM, row, c0 + j)) = o;
            }
          }
          fence_async_smem();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            mbar_arrive_on_cta(smem_u32(&bars.a2_full), 0);
            mbar_arrive_on_cta(smem_u32(&bars.acc_free[slot]), 0);
          }
        } else {
          float part = 0.f;
#pragma unroll 1
          for (int c0 = 0; c0 < H; c0 += 32) {
            float v[32];
            tmem_ld32(taddr + (uint32_t)c0, v);
#pragma unroll
            for (int j = 0; j < 32; ++j) part = fmaf(fmaxf(v[j] + bias[c0 + j], 0.f), sWlast[c0 + j], part);
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_on_cta(smem_u32(&bars.acc_free[slot]), 0);
          if (p0 + row < M) {
            const float s = part + b_last;
            score[p0 + row] = apply_sigmoid ? sigmoidf_ref(s) : s;
          }
        }
      }
    }
  } else if (warp == P_EPI_WARPS) {
    // =============================== MMA ISSUER (leader CTA, one lane) ===============================
    if (cta_rank == 0 && lane == 0) {   // lanes 1..31 wait at the __syncwarp below (keeps the warp
                                        // converged for the aligned cluster barrier at the end)
      uint32_t stage = 0, fph = 0, a2ph = 0, seq = 0;
      uint32_t afph = 0;
      const uint32_t sW_addr = smem_u32(sW), sRing_addr = smem_u32(sRing), sA2_addr = smem_u32(sA2);
      for (long long tile = cluster_id; tile < npair_tiles; tile += nclusters) {
        for (int l = 0; l < nhidden; ++l, ++seq) {
          const uint32_t slot = seq & 1;
          mbar_wait_cluster(smem_u32(&bars.acc_free[slot]), ((afph >> slot) & 1u) ^ 1u);   // first use passes
          afph ^= 1u << slot;
          tc_fence_after();
          const uint32_t d = tmem_base + slot * H;
          if (l == 0) {
            for (int c = 0; c < NCHUNK; ++c) {
              mbar_wait_cluster(smem_u32(&bars.full[stage]), fph);
              tc_fence_after();
#pragma unroll
              for (int k16 = 0; k16 < P_CHUNK_K / 16; ++k16) {
                const int k = c * P_CHUNK_K + k16 * 16;
                const uint64_t ad = umma_smem_desc_sw64(sRing_addr + stage * P_STAGE_BYTES + k16 * 32);
                const uint64_t bd = umma_smem_desc(sW_addr + (k >> 6) * (HH * 128) + ((k & 63) >> 4) * 32);
                umma_bf16_ss_2cta_p(d, ad, bd, IDESC, (c | k16) ? 1u : 0u);
              }
              umma_commit_mc(smem_u32(&bars.empty[stage]));
              if (++stage == P_RING) { stage = 0; fph ^= 1; }
            }
          } else {
            mbar_wait_cluster(smem_u32(&bars.a2_full), a2ph);
            a2ph ^= 1;
            tc_fence_after();
#pragma unroll
            for (int kb = 0; kb < H / 64; ++kb) {
#pragma unroll
              for (int k16 = 0; k16 < 4; ++k16) {
                const uint64_t ad = umma_smem_desc(sA2_addr + kb * (TC_BM * 128) + k16 * 32);
                const uint64_t bd = umma_smem_desc(sW_addr + l * WH_BYTES + kb * (HH * 128) + k16 * 32);
                umma_bf16_ss_2cta_p(d, ad, bd, IDESC, (kb | k16) ? 1u : 0u);
              }
            }
          }
          umma_commit_mc(smem_u32(&bars.acc_full[slot]));
        }
      }
    }
    __syncwarp();
  } else if (warp == P_IDS_WARP) {
    // =============================== PAIR-ID PREFETCH ===============================
    // one tile ahead of the producers: (u, v) of this CTA's 128 rows -> shared memory, so the row
    // gathers never wait on a dependent index load
    long long tl = 0;
    for (long long tile = cluster_id; tile < npair_tiles; tile += nclusters, ++tl) {
      const int slot = (int)(tl & 1);
      mbar_wait_cluster(smem_u32(&bars.ids_empty[slot]), (uint32_t)(((tl >> 1) & 1) ^ 1));
      const long long p0 = tile * (2 * TC_BM) + (long long)cta_rank * TC_BM;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int r = lane + 32 * q;
        int2 id = make_int2(-1, -1);
        if (p0 + r < M) { id.x = __ldg(pu + p0 + r); id.y = __ldg(pv + p0 + r); }
        sIds[slot * TC_BM + r] = id;
        // pull the whole 1 KB embedding rows of the NEXT tile into L2 now (one DRAM page visit per
        // row instead of eight chunk-sized ones later); the producers' loads then hit L2
        const int vprev = __shfl_up_sync(FULL, id.y, 1);
        if ((tune & 1) && id.x >= 0) {
          const char *ru = reinterpret_cast<const char *>(h + (size_t)id.x * H);
#pragma unroll
          for (int b = 0; b < H * 4; b += 128) asm volatile("prefetch.global.L2 [%0];" :: "l"(ru + b));
          if (lane == 0 || vprev != id.y) {                  // runs of equal v: prefetch each row once
            const char *rv = reinterpret_cast<const char *>(h + (size_t)id.y * H);
#pragma unroll
            for (int b = 0; b < H * 4; b += 128) asm volatile("prefetch.global.L2 [%0];" :: "l"(rv + b));
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive_on_cta(smem_u32(&bars.ids_full[slot]), cta_rank);
    }
  } else {
    // =============================== PRODUCERS ===============================
    // Three independent groups of four warps; group g owns ring stage g and produces chunks
    // g, g+3, g+6, ... of the flattened chunk stream of this CTA's tiles.  A warp has ONE batch of
    // loads in flight at a time (issue 16 x LDG.128, wait for the stage to drain, convert, store,
    // arrive), so there is no scoreboard coupling between batches; the three groups run out of
    // phase and keep three chunks (96 KB per SM) in flight.
    // Lane mapping inside a group (128 threads): 8 consecutive lanes cover one 128-byte row chunk
    // (32 fp32) so a warp instruction reads 4 whole cache lines, and rows that share v (the
    // common case inside a run of the column-major candidate order) coalesce into one request.
    const int pw = warp - P_FIRST_PROD_WARP;                 // 0..11
    const int group = pw / P_GROUP_WARPS;                    // == ring stage
    const int t = (pw % P_GROUP_WARPS) * 32 + lane;          // 0..127
    const int l8 = t & 7;
    const int rg = t >> 3;                                   // 0..15 ; rows rg + 16 q, q < 8
    long long my_tiles = 0;
    if (cluster_id < npair_tiles) my_tiles = (npair_tiles - cluster_id + nclusters - 1) / nclusters;
    const long long total = my_tiles * NCHUNK;
    uint8_t *dst = sRing + group * P_STAGE_BYTES;
    const uint32_t empty_addr = smem_u32(&bars.empty[group]), full_addr = smem_u32(&bars.full[group]);
    long long cur_tl = -1;
    int idu[8], idv[8];
    uint32_t n_use = 0, reuse = 0;
    for (long long i = group; i < total; i += P_RING, ++n_use) {
      const long long tl = i / NCHUNK;
      const int c = (int)(i - tl * NCHUNK);
      if (tl != cur_tl) {
        if (cur_tl >= 0) {                                   // done with the previous tile's ids
          __syncwarp();
          if (lane == 0) mbar_arrive_on_cta(smem_u32(&bars.ids_empty[cur_tl & 1]), cta_rank);
        }
        mbar_wait_cluster(smem_u32(&bars.ids_full[tl & 1]), (uint32_t)((tl >> 1) & 1));
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int2 id = sIds[(tl & 1) * TC_BM + rg + 16 * q];
          idu[q] = id.x; idv[q] = id.y;
        }
        // rows of one thread that repeat the previous row's v (runs of equal v in the column-major
        // candidate order) skip the h[v] load and copy the register at consume time
        reuse = 0;
        if (tune & 2) {
#pragma unroll
          for (int q = 1; q < 8; ++q)
            if (idu[q] >= 0 && idv[q] == idv[q - 1]) reuse |= 1u << q;
        }
        cur_tl = tl;
      }
      const int koff = c * P_CHUNK_K + l8 * 4;
      float4 xu[8], xv[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        if (idu[q] >= 0) {
          xu[q] = __ldg(reinterpret_cast<const float4 *>(h + (size_t)idu[q] * H + koff));
          if (!((reuse >> q) & 1u))
            xv[q] = __ldg(reinterpret_cast<const float4 *>(h + (size_t)idv[q] * H + koff));
        }
      }
      mbar_wait_cluster(empty_addr, (n_use & 1) ^ 1);        // the MMAs that read this stage retired
#pragma unroll
      for (int q = 1; q < 8; ++q)
        if ((reuse >> q) & 1u) xv[q] = xv[q - 1];            // after the loads have landed: no load-to-load dependency
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int r = rg + 16 * q;
        uint2 o = make_uint2(0u, 0u);
        if (idu[q] >= 0) {
          o.x = pack_bf16x2(xu[q].x * xv[q].x, xu[q].y * xv[q].y);
          o.y = pack_bf16x2(xu[q].z * xv[q].z, xu[q].w * xv[q].w);
        }
        *reinterpret_cast<uint2 *>(dst + sw64_chunk_off(r, l8 >> 1) + (l8 & 1) * 8) = o;
      }
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive_on_cta(full_addr, 0);
    }
  }
  tc_fence_before();
  cluster.sync();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

template <int H>
static int tc3_launch_h(const float *h, const int *pu, const int *pv, long long M, const MlpParams &prm, int L,
                        int apply_sigmoid, float *score, uint8_t *img, cudaStream_t stream) {
  const int nhidden = L - 1;
  const size_t smem = 1024 + (size_t)nhidden * (H / 2) * H * 2 + (size_t)P_RING * P_STAGE_BYTES +
                      (nhidden >= 2 ? (size_t)TC_BM * H * 2 : 0) + sizeof(float) * ((size_t)nhidden * H + H) +
                      2 * TC_BM * sizeof(int2);
  if (smem > 227 * 1024) return EPS_ERR_UNSUPPORTED;   // caller falls back to linkpred_tc2 / tc
  auto kern = linkpred_tc3_kernel<H>;
  EPS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long npair_tiles = (M + 2 * TC_BM - 1) / (2 * TC_BM);
  const int clusters = (int)std::min<long long>(npair_tiles, (long long)(sm_count() / 2));
  const char *tn = getenv("EPS_TC3_TUNE");   // bit0: L2 row prefetch by the id warp, bit1: h[v] register reuse
  const int tune = tn ? atoi(tn) : 0;
  kern<<<2 * clusters, P_THREADS, smem, stream>>>(h, pu, pv, M, prm, L, apply_sigmoid, img, score, tune);
  EPS_LAUNCH_CHECK();
  return EPS_OK;
}

// expects the per-half weight images of pack_weights_halves_kernel (linkpred_tc2.cu) in `img`
int linkpred_tc3_launch(const float *h, int H, const int *pu, const int *pv, long long M, const MlpParams &prm,
                        int L, int apply_sigmoid, float *score, uint8_t *img, cudaStream_t stream) {
  if (H == 64) return tc3_launch_h<64>(h, pu, pv, M, prm, L, apply_sigmoid, score, img, stream);
  if (H == 128) return tc3_launch_h<128>(h, pu, pv, M, prm, L, apply_sigmoid, score, img, stream);
  return tc3_launch_h<256>(h, pu, pv, M, prm, L, apply_sigmoid, score, img, stream);
}

}  // namespace eps
