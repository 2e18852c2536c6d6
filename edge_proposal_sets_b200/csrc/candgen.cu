// K6 — 2-hop candidate enumeration on the GPU.
//
// Replaces /root/reference/filter.py:96-109: the reference forms A@A on ONE CPU thread
// (torch_sparse spspmm), drops the diagonal, zeroes known edges through a scipy masked assignment
// and walks the CSC matrix, materialising 24 bytes of int64/float temporaries per candidate.
// Here each owner node v (= all_edges[:,1], the outer key of the reference's column-major order)
// is handled by one CTA: the union of N(k), k in N(v), is built as a bitmap over node ids in
// shared memory (n/8 bytes: 72 KB for ogbl-ppa), N(v) and v are cleared from it, and the set bits
// are emitted in ascending u — which IS the reference order (sorted by (v, u)); no sort needed.
// Two passes over the same owner range: count -> (host prefix sum) -> fill.  Owners are handed
// out dynamically; neighbour lists are walked as one flattened coalesced stream per 32 lists.
#include "eps_common.cuh"

namespace eps {

constexpr int CG_THREADS = 256;

template <bool FILL>
__global__ void __launch_bounds__(CG_THREADS)
twohop_kernel(const int *__restrict__ rowptr, const int *__restrict__ col, int n, int v_lo, int v_hi,
              const long long *__restrict__ offsets, unsigned int *__restrict__ counts,
              int *__restrict__ pair_u, int *__restrict__ pair_v, unsigned int *owner_counter) {
  extern __shared__ uint32_t sm[];
  const int W = (n + 31) >> 5;       // bitmap words
  const int W2 = (W + 31) >> 5;      // one bit per non-zero bitmap word
  uint32_t *bm = sm;
  uint32_t *bm2 = sm + W;
  __shared__ int s_owner;
  __shared__ uint32_t s_scan[CG_THREADS / 32];
  __shared__ uint32_t s_carry;
  const int tid = threadIdx.x, lane = lane_id(), warp = tid >> 5;
  constexpr int NW = CG_THREADS / 32;
  for (int w = tid; w < W + W2; w += CG_THREADS) sm[w] = 0;
  for (;;) {
    __syncthreads();
    if (tid == 0) s_owner = v_lo + (int)atomicAdd(owner_counter, 1u);
    __syncthreads();
    const int v = s_owner;
    if (v >= v_hi) break;
    const int vs = __ldg(rowptr + v), ve = __ldg(rowptr + v + 1);
    // ---- mark N(N(v)) ----
    for (int base = vs + warp * 32; base < ve; base += NW * 32) {
      int s = 0, len = 0;
      if (base + lane < ve) {
        const int k = __ldg(col + base + lane);
        s = __ldg(rowptr + k);
        len = __ldg(rowptr + k + 1) - s;
      }
      int pin = len;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        int t = __shfl_up_sync(FULL, pin, d);
        if (lane >= d) pin += t;
      }
      const int pex = pin - len;
      const int total = __shfl_sync(FULL, pin, 31);
      for (int j = 0; j < total; j += 32) {
        const int p = j + lane;
        int lo = 0;
#pragma unroll
        for (int step = 16; step >= 1; step >>= 1) {
          int t = __shfl_sync(FULL, pin, lo + step - 1);
          if (t <= p) lo += step;
        }
        const int s_t = __shfl_sync(FULL, s, lo);
        const int pe_t = __shfl_sync(FULL, pex, lo);
        if (p < total) {
          const int u = __ldg(col + s_t + (p - pe_t));
          const uint32_t bit = 1u << (u & 31);
          const int w = u >> 5;
          if (!(bm[w] & bit)) {                       // cheap pre-test: most bits are already set
            const uint32_t old = atomicOr(&bm[w], bit);
            if (old == 0) atomicOr(&bm2[w >> 5], 1u << (w & 31));
          }
        }
      }
    }
    __syncthreads();
    // ---- drop known edges and the diagonal (filter.py:100,103) ----
    for (int p = vs + tid; p < ve; p += CG_THREADS) {
      const int k = __ldg(col + p);
      atomicAnd(&bm[k >> 5], ~(1u << (k & 31)));
    }
    if (tid == 0) atomicAnd(&bm[v >> 5], ~(1u << (v & 31)));
    __syncthreads();
    // ---- count / emit in ascending u; thread t owns bitmap words [32*(c+t), 32*(c+t)+32) ----
    uint32_t total_cnt = 0;
    long long out_base = 0;
    if (FILL) out_base = offsets[v - v_lo];
    if (tid == 0) s_carry = 0;
    for (int c = 0; c < W2; c += CG_THREADS) {
      const int g = c + tid;
      uint32_t m2 = (g < W2) ? bm2[g] : 0u;
      uint32_t cnt = 0;
      {
        uint32_t m = m2;
        while (m) {
          const int b = __ffs(m) - 1;
          m &= m - 1;
          cnt += __popc(bm[g * 32 + b]);
        }
      }
      if (FILL) {
        // block exclusive scan of cnt
        uint32_t inc = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          uint32_t t = __shfl_up_sync(FULL, inc, d);
          if (lane >= d) inc += t;
        }
        if (lane == 31) s_scan[warp] = inc;
        __syncthreads();
        uint32_t wbase = 0, tot = 0;
#pragma unroll
        for (int w = 0; w < NW; ++w) {
          const uint32_t x = s_scan[w];
          if (w < warp) wbase += x;
          tot += x;
        }
        const uint32_t carry = s_carry;
        long long o = out_base + carry + wbase + (inc - cnt);
        uint32_t m = m2;
        while (m) {
          const int b = __ffs(m) - 1;
          m &= m - 1;
          const int w = g * 32 + b;
          uint32_t bits = bm[w];
          bm[w] = 0;
          while (bits) {
            const int q = __ffs(bits) - 1;
            bits &= bits - 1;
            pair_u[o] = (w << 5) + q;
            pair_v[o] = v;
            ++o;
          }
        }
        if (g < W2) bm2[g] = 0;
        __syncthreads();
        if (tid == 0) s_carry = carry + tot;
      } else {
        uint32_t m = m2;
        while (m) {
          const int b = __ffs(m) - 1;
          m &= m - 1;
          bm[g * 32 + b] = 0;
        }
        if (g < W2) bm2[g] = 0;
        total_cnt += cnt;
      }
    }
    if (!FILL) {
#pragma unroll
      for (int o = 16; o; o >>= 1) total_cnt += __shfl_xor_sync(FULL, total_cnt, o);
      if (lane == 0) s_scan[warp] = total_cnt;
      __syncthreads();
      if (tid == 0) {
        uint32_t t = 0;
        for (int w = 0; w < NW; ++w) t += s_scan[w];
        counts[v - v_lo] = t;
      }
    }
  }
}

}  // namespace eps

extern "C" size_t eps_twohop_workspace_bytes(void) { return 256; }

extern "C" int eps_twohop_candidates(const int32_t *rowptr, const int32_t *col, int32_t n,
                                     int32_t v_lo, int32_t v_hi, const int64_t *offsets,
                                     uint32_t *counts, int32_t *pair_u, int32_t *pair_v,
                                     void *workspace, size_t workspace_bytes, void *stream_) {
  using namespace eps;
  cudaStream_t stream = (cudaStream_t)stream_;
  EPS_CHECK_ARG(rowptr && col, "null graph pointer");
  EPS_CHECK_ARG(n > 0 && v_lo >= 0 && v_hi <= n && v_lo <= v_hi, "bad owner range");
  const bool fill = offsets != nullptr;
  EPS_CHECK_ARG(fill ? (pair_u && pair_v) : (counts != nullptr), "missing output pointer");
  if (v_lo == v_hi) return EPS_OK;
  const int sms = sm_count();
  if (sms <= 0) { set_error("eps_twohop_candidates: no CUDA device"); return EPS_ERR_CUDA; }
  if (!workspace || workspace_bytes < eps_twohop_workspace_bytes()) {
    set_error("eps_twohop_candidates: workspace too small");
    return EPS_ERR_WORKSPACE;
  }
  const int W = (n + 31) / 32, W2 = (W + 31) / 32;
  const size_t smem = (size_t)(W + W2) * 4;
  if (smem > 200 * 1024) {
    set_error("eps_twohop_candidates: n=%d needs a %zu-byte bitmap (> 200 KB of shared memory)", n, smem);
    return EPS_ERR_UNSUPPORTED;
  }
  EPS_CUDA(cudaMemsetAsync(workspace, 0, 4, stream));
  auto kern = fill ? twohop_kernel<true> : twohop_kernel<false>;
  EPS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int occ = 0;
  EPS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, CG_THREADS, smem));
  if (occ < 1) occ = 1;
  const int grid = (int)std::min<long long>((long long)(v_hi - v_lo), (long long)sms * occ);
  kern<<<grid, CG_THREADS, smem, stream>>>(rowptr, col, n, v_lo, v_hi, (const long long *)offsets,
                                           counts, pair_u, pair_v, (unsigned int *)workspace);
  EPS_LAUNCH_CHECK();
  return EPS_OK;
}
