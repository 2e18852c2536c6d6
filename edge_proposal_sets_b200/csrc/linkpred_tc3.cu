// K2 (tensor-core arm, pipelined CTA-pair kernel) — the production variant.
//
// Same math as linkpred_tc.cu / linkpred_tc2.cu (replaces /root/reference/models.py:478-485,506).
// A cluster of two CTAs (tcgen05 cta_group::2) keeps its halves of ALL hidden-layer weights resident
// in shared memory and runs three warp-specialised roles per CTA, connected by mbarriers:
//
//   producers (12 warps) gather h[u], h[v] from a bf16 copy of h (128-bit loads, the next chunk's
//                        loads in flight while the current one is converted), multiply (HMUL2.BF16)
//                        and fill a ring of 128 x 32 K-chunks (K-major SWIZZLE_64B, 5 stages at H = 256);
//   MMA issuer (1 lane,  waits for a ring stage from BOTH CTAs, issues M=256 x N=H x K=16 UMMAs
//   leader CTA only)     into one of two TMEM accumulator slots, releases stages / publishes
//                        accumulators with multicast tcgen05.commit; later layers read their A operand
//                        (the previous layer's activations) from a 3-slot ring of 64-column K-blocks;
//   epilogue (4 warps)   thread-per-row: tcgen05.ld, + bias, ReLU, then either bf16 -> the activation
//                        tile for the next layer — published per 64-column K-block, so the next
//                        layer's MMAs start while the rest of the tile is still being converted —
//                        or the fused H -> 1 output layer + sigmoid.
//
// The two accumulator slots let the tensor pipe start the next GEMM (next layer, or next tile's first
// layer) while the epilogue drains the previous one; the ring decouples the HBM/L2 gather from both.
#include <cooperative_groups.h>
#include <stdlib.h>

#include "tc_common.cuh"

namespace eps {

namespace cg = cooperative_groups;

constexpr int P_EPI_WARPS = 4;
constexpr int P_MAX_RING = 8;                                      // first-layer ring stages (8 KB each), chosen at launch
constexpr int P_A2_SLOTS = 3;                                      // activation K-block ring (16 KB each)
constexpr int P_GROUP_WARPS = 4;                                   // warps per producer group
constexpr int P_IDS_WARP = P_EPI_WARPS + 1;                        // warp 5 prefetches pair ids
constexpr int P_FIRST_PROD_WARP = P_EPI_WARPS + 2;
// NG producer groups of four warps each: 3 -> 576 threads (96 registers), 2 -> 448 threads (144 registers)
constexpr int p_threads(int ng) { return (P_EPI_WARPS + 2 + ng * P_GROUP_WARPS) * 32; }
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
  uint64_t full[P_MAX_RING];   // producers (both CTAs) -> MMA issuer       (waited in the leader)
  uint64_t empty[P_MAX_RING];  // MMA commit -> producers                    (multicast, both CTAs)
  uint64_t acc_full[2];        // MMA commit -> epilogue                     (multicast, both CTAs)
  uint64_t acc_free[2];        // epilogue (both CTAs) -> MMA issuer         (waited in the leader)
  uint64_t a2_full[P_A2_SLOTS];   // epilogue (both CTAs) -> MMA issuer: a 64-column K-block of activations is in place
  uint64_t a2_empty[P_A2_SLOTS];  // MMA commit -> epilogue                 (multicast, both CTAs)
  uint64_t ids_full[2];        // ids warp -> producers                      (CTA-local)
  uint64_t ids_empty[2];       // producers -> ids warp                      (CTA-local)
  uint32_t tmem_base_slot;
};

__device__ __forceinline__ uint32_t cvt_relu_bf16x2(float lo, float hi) {
  uint32_t d;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));   // max(x, 0) then RN to bf16
  return d;
}
__device__ __forceinline__ float2 add_f32x2(float2 a, float2 b) {   // FADD2: two fp32 adds, one issue slot
  unsigned long long pa, pb, pr;
  asm("mov.b64 %0, {%1, %2};" : "=l"(pa) : "f"(a.x), "f"(a.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(pb) : "f"(b.x), "f"(b.y));
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(pr) : "l"(pa), "l"(pb));
  float2 r;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(pr));
  return r;
}
__device__ __forceinline__ float2 fma_f32x2(float2 a, float2 b, float2 c) {
  unsigned long long pa, pb, pc, pr;
  asm("mov.b64 %0, {%1, %2};" : "=l"(pa) : "f"(a.x), "f"(a.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(pb) : "f"(b.x), "f"(b.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(pc) : "f"(c.x), "f"(c.y));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(pr) : "l"(pa), "l"(pb), "l"(pc));
  float2 r;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(pr));
  return r;
}

template <int H, bool HB /* h is a bf16 table (else fp32) */, int NG /* producer groups */>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(p_threads(NG), 1)
linkpred_tc3_kernel(const void *__restrict__ h, const int *__restrict__ pu, const int *__restrict__ pv,
                    long long M, const MlpParams prm, int L, int apply_sigmoid,
                    const uint8_t *__restrict__ wimg, float *__restrict__ score, int tune, int ring) {
  static_assert(H % 64 == 0 && H >= 64 && H <= 256, "H in {64,128,192,256}");
  constexpr int HH = H / 2;
  constexpr int WH_BYTES = HH * H * 2;
  constexpr int A2_SLOT_BYTES = TC_BM * 128;                  // 128 rows x 64 K bf16, SWIZZLE_128B
  constexpr int NCHUNK = H / P_CHUNK_K;
  constexpr int NKB = H / 64;
  constexpr int P_THREADS = p_threads(NG);
  constexpr int P_PROD_WARPS = NG * P_GROUP_WARPS;
  constexpr uint32_t TMEM_COLS = 2 * H <= 128 ? 128 : (2 * H <= 256 ? 256 : 512);
  constexpr uint32_t IDESC = umma_idesc_bf16(2 * TC_BM, H);
  // ALL shared memory is dynamic and laid out by hand (no alignment pad: the window itself is 1 KB
  // aligned — checked below — and every swizzled region starts at a multiple of 1 KB).  With H = 256
  // and two hidden layers the resident weights take 128 KB; the rest is two rings: `ring` stages of
  // the first layer's A operand (8 KB each, as many as fit) and P_A2_SLOTS K-blocks of activations.
  extern __shared__ __align__(1024) uint8_t smem[];
  const int nhidden = L - 1;
  uint8_t *sW = smem;                                                   // [nhidden][WH_BYTES]
  uint8_t *sRing = sW + (size_t)nhidden * WH_BYTES;                     // [ring][8 KB]
  uint8_t *sA2 = sRing + (size_t)ring * P_STAGE_BYTES;                  // [P_A2_SLOTS][16 KB] (nhidden >= 2)
  float *sBias = reinterpret_cast<float *>(sA2 + (nhidden >= 2 ? P_A2_SLOTS * A2_SLOT_BYTES : 0));   // [nhidden][H]
  float *sWlast = sBias + nhidden * H;                                  // [H]
  int2 *sIds = reinterpret_cast<int2 *>(sWlast + H);                    // [2][128] (u, v) of this CTA's rows
  PipeBarriers &bars = *reinterpret_cast<PipeBarriers *>(sIds + 2 * TC_BM);
  cg::cluster_group cluster = cg::this_cluster();
  const uint32_t cta_rank = cluster.block_rank();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if ((smem_u32(smem) & 1023u) != 0u) __trap();

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;"
                 :: "r"(smem_u32(&bars.tmem_base_slot)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    for (int i = 0; i < P_MAX_RING; ++i) {
      mbar_init(smem_u32(&bars.full[i]), 2 * P_GROUP_WARPS);
      mbar_init(smem_u32(&bars.empty[i]), 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&bars.acc_full[i]), 1);
      mbar_init(smem_u32(&bars.acc_free[i]), 2 * P_EPI_WARPS);
    }
    for (int i = 0; i < P_A2_SLOTS; ++i) {
      mbar_init(smem_u32(&bars.a2_full[i]), 2 * P_EPI_WARPS);
      mbar_init(smem_u32(&bars.a2_empty[i]), 1);
    }
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
  const uint32_t tmem_base = bars.tmem_base_slot;

  const long long npair_tiles = (M + 2 * TC_BM - 1) / (2 * TC_BM);
  const long long nclusters = gridDim.x / 2, cluster_id = blockIdx.x / 2;

  if (warp < P_EPI_WARPS) {
    // =============================== EPILOGUE ===============================
    // Two fp32 adds per FADD2, ReLU fused into the bf16 pack (cvt.rn.relu.bf16x2), biases read with
    // warp-uniform 128-bit shared loads: a hidden-layer element costs ~1.4 issue slots.
    uint32_t acph = 0;            // bit s = phase parity of acc_full[s] (kept in a register)
    uint32_t seq = 0, a2g = 0;    // a2g: activation K-blocks produced so far (slot = a2g % P_A2_SLOTS)
    const int row = warp * 32 + lane;
    for (long long tile = cluster_id; tile < npair_tiles; tile += nclusters) {
      const long long p0 = tile * (2 * TC_BM) + (long long)cta_rank * TC_BM;
      for (int l = 0; l < nhidden; ++l, ++seq) {
        const uint32_t slot = seq & 1;
        mbar_wait_cluster(smem_u32(&bars.acc_full[slot]), (acph >> slot) & 1u);
        acph ^= 1u << slot;
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + slot * H;
        const float4 *b4 = reinterpret_cast<const float4 *>(sBias + l * H);
        if (l < nhidden - 1) {
#pragma unroll 1
          for (int kb = 0; kb < NKB; ++kb, ++a2g) {
            // K-block kb of the next layer's A operand goes to slot a2g % 3 of the activation ring:
            // wait until the MMAs that read the slot's previous contents have retired
            const uint32_t a2s = a2g % P_A2_SLOTS;
            mbar_wait_cluster(smem_u32(&bars.a2_empty[a2s]), ((a2g / P_A2_SLOTS) & 1u) ^ 1u);
            uint8_t *dstrow = sA2 + a2s * A2_SLOT_BYTES + row * 128;
#pragma unroll
            for (int half = 0; half < 2; ++half) {
              const int c0 = kb * 64 + half * 32;
              float v[32];
              tmem_ld32(taddr + (uint32_t)c0, v);
#pragma unroll
              for (int j = 0; j < 32; j += 8) {
                const float4 ba = b4[(c0 + j) >> 2], bc = b4[((c0 + j) >> 2) + 1];
                const float2 t0 = add_f32x2(make_float2(v[j + 0], v[j + 1]), make_float2(ba.x, ba.y));
                const float2 t1 = add_f32x2(make_float2(v[j + 2], v[j + 3]), make_float2(ba.z, ba.w));
                const float2 t2 = add_f32x2(make_float2(v[j + 4], v[j + 5]), make_float2(bc.x, bc.y));
                const float2 t3 = add_f32x2(make_float2(v[j + 6], v[j + 7]), make_float2(bc.z, bc.w));
                uint4 o;
                o.x = cvt_relu_bf16x2(t0.x, t0.y); o.y = cvt_relu_bf16x2(t1.x, t1.y);
                o.z = cvt_relu_bf16x2(t2.x, t2.y); o.w = cvt_relu_bf16x2(t3.x, t3.y);
                const int chunk = half * 4 + (j >> 3);
                *reinterpret_cast<uint4 *>(dstrow + ((chunk ^ (row & 7)) << 4)) = o;
              }
            }
            // the K-block is complete: the tensor pipe starts the next layer on it while the
            // remaining columns of this tile are still being converted
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive_on_cta(smem_u32(&bars.a2_full[a2s]), 0);
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_on_cta(smem_u32(&bars.acc_free[slot]), 0);
        } else {
          const float4 *w4 = reinterpret_cast<const float4 *>(sWlast);
          float2 part = make_float2(0.f, 0.f);
#pragma unroll 1
          for (int c0 = 0; c0 < H; c0 += 32) {
            float v[32];
            tmem_ld32(taddr + (uint32_t)c0, v);
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 ba = b4[(c0 + j) >> 2], wa = w4[(c0 + j) >> 2];
              float2 t0 = add_f32x2(make_float2(v[j + 0], v[j + 1]), make_float2(ba.x, ba.y));
              float2 t1 = add_f32x2(make_float2(v[j + 2], v[j + 3]), make_float2(ba.z, ba.w));
              t0.x = fmaxf(t0.x, 0.f); t0.y = fmaxf(t0.y, 0.f);
              t1.x = fmaxf(t1.x, 0.f); t1.y = fmaxf(t1.y, 0.f);
              part = fma_f32x2(t0, make_float2(wa.x, wa.y), part);
              part = fma_f32x2(t1, make_float2(wa.z, wa.w), part);
            }
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_on_cta(smem_u32(&bars.acc_free[slot]), 0);
          if (p0 + row < M) {
            const float s = (part.x + part.y) + b_last;
            score[p0 + row] = apply_sigmoid ? sigmoidf_ref(s) : s;
          }
        }
      }
    }
  } else if (warp == P_EPI_WARPS) {
    // =============================== MMA ISSUER (leader CTA, one lane) ===============================
    if (cta_rank == 0 && lane == 0) {   // lanes 1..31 wait at the __syncwarp below (keeps the warp
                                        // converged for the aligned cluster barrier at the end)
      uint32_t ci = 0, a2g = 0, seq = 0;    // ci: chunks consumed so far (stage = ci % ring); a2g: activation K-blocks
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
            for (int c = 0; c < NCHUNK; ++c, ++ci) {
              const uint32_t stage = ci % (uint32_t)ring;
              mbar_wait_cluster(smem_u32(&bars.full[stage]), (ci / (uint32_t)ring) & 1u);
              tc_fence_after();
#pragma unroll
              for (int k16 = 0; k16 < P_CHUNK_K / 16; ++k16) {
                const int k = c * P_CHUNK_K + k16 * 16;
                const uint64_t ad = umma_smem_desc_sw64(sRing_addr + stage * P_STAGE_BYTES + k16 * 32);
                const uint64_t bd = umma_smem_desc(sW_addr + (k >> 6) * (HH * 128) + ((k & 63) >> 4) * 32);
                umma_bf16_ss_2cta_p(d, ad, bd, IDESC, (c | k16) ? 1u : 0u);
              }
              umma_commit_mc(smem_u32(&bars.empty[stage]));
            }
          } else {
#pragma unroll 1
            for (int kb = 0; kb < NKB; ++kb, ++a2g) {
              const uint32_t a2s = a2g % P_A2_SLOTS;
              mbar_wait_cluster(smem_u32(&bars.a2_full[a2s]), (a2g / P_A2_SLOTS) & 1u);   // K-block kb is in place
              tc_fence_after();
#pragma unroll
              for (int k16 = 0; k16 < 4; ++k16) {
                const uint64_t ad = umma_smem_desc(sA2_addr + a2s * A2_SLOT_BYTES + k16 * 32);
                const uint64_t bd = umma_smem_desc(sW_addr + l * WH_BYTES + kb * (HH * 128) + k16 * 32);
                umma_bf16_ss_2cta_p(d, ad, bd, IDESC, (kb | k16) ? 1u : 0u);
              }
              umma_commit_mc(smem_u32(&bars.a2_empty[a2s]));      // slot reusable once these MMAs retire
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
        // pull the whole embedding rows of the NEXT tile into L2 now (one DRAM page visit per row
        // instead of eight chunk-sized ones later); the producers' loads then hit L2
        const int vprev = __shfl_up_sync(FULL, id.y, 1);
        if ((tune & 1) && id.x >= 0) {
          constexpr int RB = H * (HB ? 2 : 4);
          const char *ru = reinterpret_cast<const char *>(h) + (size_t)id.x * RB;
#pragma unroll
          for (int b = 0; b < RB; b += 128) asm volatile("prefetch.global.L2 [%0];" :: "l"(ru + b));
          if (lane == 0 || vprev != id.y) {                  // runs of equal v: prefetch each row once
            const char *rv = reinterpret_cast<const char *>(h) + (size_t)id.y * RB;
#pragma unroll
            for (int b = 0; b < RB; b += 128) asm volatile("prefetch.global.L2 [%0];" :: "l"(rv + b));
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive_on_cta(smem_u32(&bars.ids_full[slot]), cta_rank);
    }
  } else {
    // =============================== PRODUCERS ===============================
    // Three independent groups of four warps; group g produces chunks g, g+3, g+6, ... of the
    // flattened chunk stream of this CTA's tiles (a chunk = 128 rows x 32 K) into ring stage
    // (chunk % ring) — more stages than groups, so a group never waits on the stage it has just filled.
    // Lane mapping inside a group (128 threads): 4 consecutive lanes cover one row's chunk, each lane
    // 8 K-elements = one 16-byte unit of the swizzled stage; a warp instruction covers 8 rows, and
    // rows that share v (the common case inside a run of the column-major candidate order) coalesce
    // into one request.
    //   HB (bf16 copy of h, the hot path): 2 x LDG.128 per row (u, v) -> 8 loads per chunk, and the
    //       loads of the NEXT chunk are issued before the current one is converted (two register
    //       buffers), so every producer thread keeps 8-16 x 16 B in flight (~75 KB per SM);
    //   fp32 source (small pair lists, no table): 4 x LDG.128 per row, one chunk in flight.
    const int pw = warp - P_FIRST_PROD_WARP;                 // 0..11
    const int group = pw / P_GROUP_WARPS;
    const int t = (pw % P_GROUP_WARPS) * 32 + lane;          // 0..127
    const int l4 = t & 3;
    const int rg = t >> 2;                                   // 0..31 ; rows rg + 32 q, q < 4
    long long my_tiles = 0;
    if (cluster_id < npair_tiles) my_tiles = (npair_tiles - cluster_id + nclusters - 1) / nclusters;
    const long long total = my_tiles * NCHUNK;
    const char *hbase = reinterpret_cast<const char *>(h);
    constexpr int ROW_BYTES = H * (HB ? 2 : 4);
    constexpr int LPR = HB ? 1 : 2;                          // LDG.128 per row and operand
    long long cur_tl = -1;
    int idu[4], idv[4];
    struct Buf { uint4 xu[4][LPR], xv[4][LPR]; uint32_t valid; };

    // Move this warp's position in the pair-id pipeline to tile `tl`: release every tile left behind
    // (also tiles this group has no chunk in — H = 64 has 2 chunks per tile for 3 groups), waiting
    // for each tile's ids to have been published first so that a release can never be counted
    // towards an earlier phase of the same slot.
    auto advance_to = [&](long long tl) {
      while (cur_tl < tl) {
        if (cur_tl >= 0) {
          __syncwarp();
          if (lane == 0) mbar_arrive_on_cta(smem_u32(&bars.ids_empty[cur_tl & 1]), cta_rank);
        }
        ++cur_tl;
        if (cur_tl < my_tiles)
          mbar_wait_cluster(smem_u32(&bars.ids_full[cur_tl & 1]), (uint32_t)((cur_tl >> 1) & 1));
      }
    };
    auto issue = [&](Buf &b, long long i) {
      const long long tl = i / NCHUNK;
      const int c = (int)(i - tl * NCHUNK);
      if (tl != cur_tl) {
        advance_to(tl);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int2 id = sIds[(tl & 1) * TC_BM + rg + 32 * q];
          idu[q] = id.x; idv[q] = id.y;
        }
      }
      const int boff = (c * P_CHUNK_K + l4 * 8) * (HB ? 2 : 4);
      b.valid = 0;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (idu[q] >= 0) {
          b.valid |= 1u << q;
          const uint4 *pu4 = reinterpret_cast<const uint4 *>(hbase + (size_t)idu[q] * ROW_BYTES + boff);
          const uint4 *pv4 = reinterpret_cast<const uint4 *>(hbase + (size_t)idv[q] * ROW_BYTES + boff);
#pragma unroll
          for (int j = 0; j < LPR; ++j) { b.xu[q][j] = __ldg(pu4 + j); b.xv[q][j] = __ldg(pv4 + j); }
        }
      }
    };
    auto consume = [&](Buf &b, long long i) {
      const uint32_t stage = (uint32_t)(i % ring);
      uint8_t *dst = sRing + stage * P_STAGE_BYTES;
      mbar_wait_cluster(smem_u32(&bars.empty[stage]), (uint32_t)(((i / ring) & 1) ^ 1));   // the MMAs that read this stage retired
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int r = rg + 32 * q;
        uint4 o = make_uint4(0u, 0u, 0u, 0u);
        if ((b.valid >> q) & 1u) {
          if (HB) {
            o.x = mul_bf16x2(b.xu[q][0].x, b.xv[q][0].x); o.y = mul_bf16x2(b.xu[q][0].y, b.xv[q][0].y);
            o.z = mul_bf16x2(b.xu[q][0].z, b.xv[q][0].z); o.w = mul_bf16x2(b.xu[q][0].w, b.xv[q][0].w);
          } else {
            const uint4 a0 = b.xu[q][0], a1 = b.xu[q][LPR - 1], c0 = b.xv[q][0], c1 = b.xv[q][LPR - 1];
            o.x = hadamard_bf16x2(__uint_as_float(a0.x), __uint_as_float(a0.y), __uint_as_float(c0.x), __uint_as_float(c0.y));
            o.y = hadamard_bf16x2(__uint_as_float(a0.z), __uint_as_float(a0.w), __uint_as_float(c0.z), __uint_as_float(c0.w));
            o.z = hadamard_bf16x2(__uint_as_float(a1.x), __uint_as_float(a1.y), __uint_as_float(c1.x), __uint_as_float(c1.y));
            o.w = hadamard_bf16x2(__uint_as_float(a1.z), __uint_as_float(a1.w), __uint_as_float(c1.z), __uint_as_float(c1.w));
          }
        }
        *reinterpret_cast<uint4 *>(dst + sw64_chunk_off(r, l4)) = o;
      }
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive_on_cta(smem_u32(&bars.full[stage]), 0);
    };

    long long i = group;
    if (HB) {
      Buf A, B;
      if (i < total) issue(A, i);
      while (i < total) {
        if (i + NG < total) issue(B, i + NG);
        consume(A, i);
        i += NG;
        if (i >= total) break;
        if (i + NG < total) issue(A, i + NG);
        consume(B, i);
        i += NG;
      }
    } else {
      Buf A;
      for (; i < total; i += NG) { issue(A, i); consume(A, i); }
    }
    advance_to(my_tiles);                                    // release the tiles this group had no chunk in
  }
  tc_fence_before();
  cluster.sync();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

template <int H, bool HB, int NG>
static int tc3_launch_g(const void *h, const int *pu, const int *pv, long long M, const MlpParams &prm, int L,
                        int apply_sigmoid, float *score, uint8_t *img, cudaStream_t stream) {
  const int nhidden = L - 1;
  const size_t fixed = (size_t)nhidden * (H / 2) * H * 2 + (nhidden >= 2 ? (size_t)P_A2_SLOTS * TC_BM * 128 : 0) +
                       sizeof(float) * ((size_t)nhidden * H + H) + 2 * TC_BM * sizeof(int2) + sizeof(PipeBarriers);
  const size_t budget = 227 * 1024;
  if (fixed + (size_t)(NG + 1) * P_STAGE_BYTES > budget) return EPS_ERR_UNSUPPORTED;   // caller falls back to linkpred_tc2 / tc
  int ring = (int)std::min<size_t>((budget - fixed) / P_STAGE_BYTES, (size_t)P_MAX_RING);
  const char *rg = getenv("EPS_TC3_RING");    // cap the ring depth (A/B measurements)
  if (rg && atoi(rg) >= NG + 1) ring = std::min(ring, atoi(rg));
  const size_t smem = fixed + (size_t)ring * P_STAGE_BYTES;
  auto kern = linkpred_tc3_kernel<H, HB, NG>;
  EPS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long npair_tiles = (M + 2 * TC_BM - 1) / (2 * TC_BM);
  const int clusters = (int)std::min<long long>(npair_tiles, (long long)(sm_count() / 2));
  const char *tn = getenv("EPS_TC3_TUNE");   // bit0: L2 row prefetch by the id warp (default on)
  const int tune = tn ? atoi(tn) : 1;
  kern<<<2 * clusters, p_threads(NG), smem, stream>>>(h, pu, pv, M, prm, L, apply_sigmoid, img, score, tune, ring);
  EPS_LAUNCH_CHECK();
  return EPS_OK;
}

template <int H, bool HB>
static int tc3_launch_h(const void *h, const int *pu, const int *pv, long long M, const MlpParams &prm, int L,
                        int apply_sigmoid, float *score, uint8_t *img, cudaStream_t stream) {
  const char *g = getenv("EPS_TC3_GROUPS");   // producer groups: 3 (default) or 2 (A/B measurements)
  if (g && g[0] == '2') return tc3_launch_g<H, HB, 2>(h, pu, pv, M, prm, L, apply_sigmoid, score, img, stream);
  return tc3_launch_g<H, HB, 3>(h, pu, pv, M, prm, L, apply_sigmoid, score, img, stream);
}

// fp32 embeddings -> bf16 table (round to nearest even), 8 elements per thread
__global__ void __launch_bounds__(256) h_to_bf16_kernel(const float *__restrict__ h, long long n8, uint4 *__restrict__ out) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += stride) {
    const float4 a = __ldg(reinterpret_cast<const float4 *>(h) + 2 * i);
    const float4 b = __ldg(reinterpret_cast<const float4 *>(h) + 2 * i + 1);
    uint4 o;
    o.x = pack_bf16x2(a.x, a.y); o.y = pack_bf16x2(a.z, a.w);
    o.z = pack_bf16x2(b.x, b.y); o.w = pack_bf16x2(b.z, b.w);
    out[i] = o;
  }
}

int h_to_bf16_launch(const float *h, long long elems, void *out, cudaStream_t stream) {
  const long long n8 = elems / 8;
  if (n8 == 0) return EPS_OK;
  const int grid = (int)std::min<long long>((n8 + 255) / 256, (long long)sm_count() * 16);
  h_to_bf16_kernel<<<grid, 256, 0, stream>>>(h, n8, reinterpret_cast<uint4 *>(out));
  EPS_LAUNCH_CHECK();
  return EPS_OK;
}

// expects the per-half weight images of pack_weights_halves_kernel (linkpred_tc2.cu) in `img`;
// h_is_bf16: `h` is the bf16 table written by h_to_bf16_launch, else the caller's fp32 matrix
int linkpred_tc3_launch(const void *h, int h_is_bf16, int H, const int *pu, const int *pv, long long M,
                        const MlpParams &prm, int L, int apply_sigmoid, float *score, uint8_t *img,
                        cudaStream_t stream) {
#define EPS_TC3(HV)                                                                                       \
  return h_is_bf16 ? tc3_launch_h<HV, true>(h, pu, pv, M, prm, L, apply_sigmoid, score, img, stream)      \
                   : tc3_launch_h<HV, false>(h, pu, pv, M, prm, L, apply_sigmoid, score, img, stream)
  if (H == 64) { EPS_TC3(64); }
  if (H == 128) { EPS_TC3(128); }
  EPS_TC3(256);
#undef EPS_TC3
}

}  // namespace eps
