// K6+K3 fused — 2-hop candidate enumeration WITH Common-Neighbour / Adamic-Adar / RA scores.
//
// Replaces /root/reference/filter.py:96-109 followed by the per-batch scoring loop
// filter.py:113-142 for the heuristic filter models ('simple', 'adamic', 'adamic_ogb',
// 'resource_allocation').  The reference forms A@A to enumerate candidates, THROWS THE VALUES AWAY
// (filter.py:108-109) and then re-derives them pair by pair (models.py:536-554,
// adamic_utils.py:20-23).  But the values are the scores: every 2-path v - k - u contributes one
// common neighbour k to the pair (u, v), so
//     CN(u, v) = #2-paths,      AA / RA (u, v) = sum over the 2-paths of w_k .
// Walking the 2-paths of an owner v costs  sum_{k in N(v)} deg(k)  list elements, whereas scoring
// its candidates one by one (K3) costs  sum_{u in cand(v)} deg(u)  — ~100x more on the ogbl-ppa
// shape (mean CN ~ 1.4, mean degree of a candidate ~ 150).
//
// One CTA per owner v (handed out dynamically), everything in shared memory but the outputs:
//   1. mark     bitmap of U = union of N(k), k in N(v)              (flattened coalesced walk)
//   2. clear    N(v) and v                                           (filter.py:100,103)
//   3. rank     per-word exclusive popcount prefix (u16 inside a 32-word block + u32 per block),
//               and emission of pair_u / pair_v in ascending u — the reference's column-major order
//   4. score    second walk over the same 2-paths: a set bit u -> rank(u) -> one RED.ADD into the
//               compact accumulator of that candidate (count as int32, weight as 64-bit fixed point,
//               eps_common.cuh) — integer atomics, hence exact and order-independent
//   5. cleanup  zero the touched bitmap words
// A tiny element-wise pass then turns the fixed-point sums into fp32 scores (+ sigmoid).
// Results are bit-identical to eps_cn_aa on the same pairs.
//
// Where an owner's output starts:
//   * two-pass (eps_twohop_scored): from the caller's prefix sum of a separate count pass
//     (eps_twohop_candidates) — the count pass repeats step 1 for every owner;
//   * ONE-PASS (eps_twohop_onepass): every owner writes into a PADDED slot whose offset is the prefix sum
//     of a cheap per-owner upper bound (min(#2-paths, n-1-deg), host-side torch ops), and records its
//     real count; a single-block scan turns the counts into compact offsets and the finalize pass —
//     which has to read every accumulator anyway — moves (u, score, count) to the compact position and
//     regenerates v from the owner id.  No count pass, no inter-CTA dependency.
#include "eps_common.cuh"

namespace eps {

constexpr int TS_THREADS = 512;           // two-pass entry point
constexpr int TS_DEFAULT_THREADS = 512;   // one-pass entry point (see eps_twohop_onepass)
constexpr int TS_LONG = 96;   // lists at least this long are streamed warp-wide without a search
constexpr int TS_DEPTH = 8;   // ... with this many loads in flight per lane (a hub owner's walk is one CTA's latency chain)

// Visit every element of the neighbour lists of N(v).  A warp takes LG lists at a time (LG = 32, 16, ... 1,
// warp-uniform), and the groups are HANDED OUT through a shared-memory counter: with a static round-robin an owner of
// average degree (74 on the ppa shape) gave some warps two groups and the others one, and a warp that drew a hub's
// list kept the other fifteen waiting — 46 % of all stall samples sat at the two barriers behind the walks
// (ncu source counters, profiles/round2_e_twohop.md).  The caller picks the largest LG that still leaves ~4 groups
// per warp to balance with.  `counter` is zero when the walk starts (the caller resets it behind a barrier).
// f.visit(u, p) is called once per 2-path v - k - u (p = index of u in `col`, i.e. inside N(k)) after
// f.select_slot*(lane that holds k's metadata).
template <typename F>
__device__ __forceinline__ void walk_two_paths(const int *__restrict__ rowptr, const int *__restrict__ col,
                                               int vs, int ve, int lane, int LG, int *counter, F f) {
  for (;;) {
    int grp = 0;
    if (lane == 0) grp = atomicAdd(counter, 1);
    grp = __shfl_sync(FULL, grp, 0);
    const int base = vs + grp * LG;
    if (base >= ve) break;
    int k = -1, s = 0, len = 0;
    if (lane < LG && base + lane < ve) {
      k = __ldg(col + base + lane);
      s = __ldg(rowptr + k);
      len = __ldg(rowptr + k + 1) - s;
    }
    f.load_slot(k, base + lane);
    // ---- long lists: the whole warp streams one list, TS_DEPTH loads in flight per lane ----
    unsigned longmask = __ballot_sync(FULL, len >= TS_LONG);
    while (longmask) {
      const int b = __ffs(longmask) - 1;
      longmask &= longmask - 1;
      const int sb = __shfl_sync(FULL, s, b);
      const int lb = __shfl_sync(FULL, len, b);
      f.select_slot(b);
      const int *__restrict__ lp = col + sb;
      for (int off = 0; off < lb; off += 32 * TS_DEPTH) {
        int u[TS_DEPTH];
#pragma unroll
        for (int q = 0; q < TS_DEPTH; ++q) {
          const int p = off + q * 32 + lane;
          u[q] = p < lb ? __ldg(lp + p) : -1;
        }
#pragma unroll
        for (int q = 0; q < TS_DEPTH; ++q)
          if (u[q] >= 0) f.visit(u[q], sb + off + q * 32 + lane);
      }
    }
    // ---- short lists: one flattened sequence ----
    const int slen = len >= TS_LONG ? 0 : len;
    int pin = slen;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      int t = __shfl_up_sync(FULL, pin, d);
      if (lane >= d) pin += t;
    }
    const int pex = pin - slen;
    const int total = __shfl_sync(FULL, pin, 31);
    for (int j = 0; j < total; j += 32) {
      const int p = j + lane;
      int lo = 0;
      for (int step = LG >> 1; step >= 1; step >>= 1) {      // first slot whose inclusive prefix exceeds p
        int t = __shfl_sync(FULL, pin, lo + step - 1);
        if (t <= p) lo += step;
      }
      const int s_t = __shfl_sync(FULL, s, lo);
      const int pe_t = __shfl_sync(FULL, pex, lo);
      f.select_slot_lane(lo);
      if (p < total) f.visit(__ldg(col + s_t + (p - pe_t)), s_t + (p - pe_t));
    }
  }
}

struct MarkVisitor {
  uint32_t *bm, *bm2;
  __device__ __forceinline__ void load_slot(int, int) {}
  __device__ __forceinline__ void select_slot(int) {}
  __device__ __forceinline__ void select_slot_lane(int) {}
  __device__ __forceinline__ void visit(int u, int) {
    const uint32_t bit = 1u << (u & 31);
    const int w = u >> 5;
    if (!(bm[w] & bit)) {                          // cheap pre-test: most bits are already set
      const uint32_t old = atomicOr(&bm[w], bit);
      if (old == 0) atomicOr(&bm2[w >> 5], 1u << (w & 31));
    }
  }
};

// Weighted adjacency (HAS_VAL, collab): the term of the 2-path v - k - u is a_u * (a_v * w_k) (or a_u * a_v
// without a weight table) with a_v = A[v,k] (the value next to k in N(v)) and a_u = A[k,u] (next to u in
// N(k)) — the same fp32 products as eps_cn_aa forms from A[u,k]; the caller guarantees A[k,u] == A[u,k]
// bit for bit (candidates.py checks it once per graph).
template <bool HAS_W, bool WANT_CN, bool HAS_VAL>
struct ScoreVisitor {
  const uint32_t *bm, *blk;
  const uint16_t *pre;
  const float *__restrict__ wtable;
  const float *__restrict__ val;
  unsigned long long *acc;   // + out_base already applied
  int *cn;
  unsigned long long fx_slot = 0ull, fx = 0ull;
  float sv_slot = 0.f, sv = 0.f;
  __device__ __forceinline__ void load_slot(int k, int pos) {
    if (HAS_VAL) {
      const float a_v = k >= 0 ? __ldg(val + pos) : 0.f;
      sv_slot = (HAS_W && k >= 0) ? __fmul_rn(a_v, __ldg(wtable + k)) : a_v;
    } else if (HAS_W) {
      fx_slot = k >= 0 ? to_fixed(__ldg(wtable + k)) : 0ull;
    }
  }
  __device__ __forceinline__ void select_slot(int b) {           // warp-uniform slot
    if (HAS_VAL) sv = __shfl_sync(FULL, sv_slot, b);
    else if (HAS_W) fx = __shfl_sync(FULL, fx_slot, b);
  }
  __device__ __forceinline__ void select_slot_lane(int lo) {     // per-lane slot
    if (HAS_VAL) sv = __shfl_sync(FULL, sv_slot, lo);
    else if (HAS_W) fx = __shfl_sync(FULL, fx_slot, lo);
  }
  __device__ __forceinline__ void visit(int u, int p) {
    const int w = u >> 5, b = u & 31;
    const uint32_t word = bm[w];
    if ((word >> b) & 1u) {
      const uint32_t idx = blk[w >> 5] + pre[w] + __popc(word & ((1u << b) - 1u));
      if (HAS_VAL) atomicAdd(acc + idx, to_fixed(__fmul_rn(__ldg(val + p), sv)));
      else if (HAS_W) atomicAdd(acc + idx, fx);
      if (WANT_CN) atomicAdd(cn + idx, 1);
    }
  }
};

template <bool HAS_W, bool WANT_CN, bool ONEPASS, bool HAS_VAL = false, int THREADS = TS_THREADS>
__global__ void __launch_bounds__(THREADS)
twohop_score_kernel(const int *__restrict__ rowptr, const int *__restrict__ col, const float *__restrict__ val,
                    const float *__restrict__ wtable, int n, int v_lo, int v_hi,
                    const long long *__restrict__ offsets, int *__restrict__ pair_u,
                    int *__restrict__ pair_v, unsigned long long *__restrict__ acc, int *__restrict__ cn,
                    unsigned int *owner_counter, unsigned int *__restrict__ counts_out,
                    const int *__restrict__ owner_order, int n_order) {
  extern __shared__ uint32_t sm[];
  const int W = (n + 31) >> 5;       // bitmap words
  const int W2 = (W + 31) >> 5;      // blocks of 32 words
  uint32_t *bm = sm;
  uint32_t *bm2 = bm + W;            // one bit per non-zero bitmap word
  uint32_t *blk = bm2 + W2;          // #candidates before block g
  uint16_t *pre = reinterpret_cast<uint16_t *>(blk + W2);   // #candidates before word w inside its block
  __shared__ int s_owner;
  __shared__ int s_walk;             // next list group of the current walk
  __shared__ uint32_t s_scan[THREADS / 32];
  __shared__ uint32_t s_carry;
  const int tid = threadIdx.x, lane = lane_id(), warp = tid >> 5;
  constexpr int NW = THREADS / 32;
  for (int w = tid; w < W + W2; w += THREADS) sm[w] = 0;
  for (;;) {
    __syncthreads();
    if (tid == 0) {
      // owners are handed out HEAVIEST FIRST when the caller supplies an order (one-pass entry point): a hub owner
      // costs ~40 average owners, and drawn late by one CTA it kept the other 295 waiting — SMs were busy 48 % of the
      // kernel's duration (profiles/round2_e_twohop.md)
      const int t = (int)atomicAdd(owner_counter, 1u);
      s_owner = (owner_order && t < n_order) ? v_lo + owner_order[t] : v_lo + t;
      s_walk = 0;
    }
    __syncthreads();
    const int v = s_owner;
    if (v >= v_hi) break;
    const int vs = __ldg(rowptr + v), ve = __ldg(rowptr + v + 1);
    int LG = 32;                                           // lists per group: ~4 groups per warp to balance with
    while (LG > 1 && (ve - vs + LG - 1) / LG < 4 * NW) LG >>= 1;
    const long long out_base = offsets[v - v_lo];
    if (offsets[v - v_lo + 1] == out_base) continue;   // no (room for) candidates (uniform): bitmap untouched
    // ---- 1. mark ----
    walk_two_paths(rowptr, col, vs, ve, lane, LG, &s_walk, MarkVisitor{bm, bm2});
    __syncthreads();
    // ---- 2. clear known edges and the diagonal ----
    for (int p = vs + tid; p < ve; p += THREADS) {
      const int k = __ldg(col + p);
      atomicAnd(&bm[k >> 5], ~(1u << (k & 31)));
    }
    if (tid == 0) atomicAnd(&bm[v >> 5], ~(1u << (v & 31)));
    if (tid == 0) { s_carry = 0; s_walk = 0; }        // the first walk is over (barrier above): reset for the second
    __syncthreads();
    // ---- 3a. rank: thread t owns block c + t (32 bitmap words) ----
    for (int c = 0; c < W2; c += THREADS) {
      const int g = c + tid;
      const uint32_t m2 = (g < W2) ? bm2[g] : 0u;
      uint32_t cnt = 0;
      {
        uint32_t m = m2;
        while (m) {
          const int b = __ffs(m) - 1;
          m &= m - 1;
          const int w = g * 32 + b;
          pre[w] = (uint16_t)cnt;
          cnt += __popc(bm[w]);
        }
      }
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
      if (g < W2) blk[g] = carry + wbase + (inc - cnt);
      __syncthreads();
      if (tid == 0) s_carry = carry + tot;
    }
    __syncthreads();
    if (ONEPASS) {
      const uint32_t total = s_carry;
      const bool fits = (long long)total <= offsets[v - v_lo + 1] - out_base;
      // padded slot: record the real count (a violated bound poisons N instead of overrunning the slot)
      if (tid == 0) counts_out[v - v_lo] = fits ? total : 0xffffffffu;
      if (!fits) {
        for (int g = tid; g < W2; g += THREADS) {
          uint32_t m = bm2[g];
          while (m) {
            const int b = __ffs(m) - 1;
            m &= m - 1;
            bm[g * 32 + b] = 0;
          }
          bm2[g] = 0;
        }
        continue;
      }
    }
    // ---- 3b. emit pair_u / pair_v in ascending u ----
    for (int g = tid; g < W2; g += THREADS) {
      uint32_t m = bm2[g];
      long long o = out_base + blk[g];
      while (m) {
        const int b = __ffs(m) - 1;
        m &= m - 1;
        const int w = g * 32 + b;
        uint32_t bits = bm[w];
        while (bits) {
          const int q = __ffs(bits) - 1;
          bits &= bits - 1;
          pair_u[o] = (w << 5) + q;
          if (!ONEPASS) pair_v[o] = v;                 // one-pass: v is regenerated by the compaction
          ++o;
        }
      }
    }
    // ---- 4. score: second walk, one integer RED per 2-path that lands on a candidate ----
    if (HAS_W || WANT_CN || HAS_VAL) {
      ScoreVisitor<HAS_W, WANT_CN, HAS_VAL> sv{bm, blk, pre, wtable, val,
                                               (HAS_W || HAS_VAL) ? acc + out_base : nullptr,
                                               WANT_CN ? cn + out_base : nullptr};
      walk_two_paths(rowptr, col, vs, ve, lane, LG, &s_walk, sv);
    }
    __syncthreads();
    // ---- 5. cleanup ----
    for (int c = 0; c < W2; c += THREADS) {
      const int g = c + tid;
      if (g < W2) {
        uint32_t m = bm2[g];
        while (m) {
          const int b = __ffs(m) - 1;
          m &= m - 1;
          bm[g * 32 + b] = 0;
        }
        bm2[g] = 0;
      }
    }
  }
}

__global__ void __launch_bounds__(256)
twohop_finalize_kernel(const unsigned long long *__restrict__ acc, const int *__restrict__ cn,
                       long long N, int flags, float *__restrict__ score) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += stride) {
    float sc = acc ? from_fixed(acc[i]) : (float)cn[i];
    if (flags & EPS_CN_SIGMOID) sc = sigmoidf_ref(sc);
    score[i] = sc;
  }
}

// Owner order for the one-pass kernel, heaviest first: bucket b = 63 - floor(log2(slot size + 1)), so bucket 0 holds the
// largest slots.  PASS 0 counts, PASS 1 turns the 64 counts into bases (one warp pair), PASS 2 scatters (order inside a
// bucket is whatever the atomics give: the outputs do not depend on the order owners are processed in).
template <int PASS>
__global__ void __launch_bounds__(256)
owner_bucket_kernel(const long long *__restrict__ boff, int n_own, unsigned int *__restrict__ buckets, int *__restrict__ order) {
  if (PASS == 1) {
    __shared__ unsigned int c[64];
    const int t = threadIdx.x;
    c[t] = buckets[t];
    __syncthreads();
    if (t == 0) {
      unsigned int run = 0;
      for (int b = 0; b < 64; ++b) { const unsigned int x = c[b]; c[b] = run; run += x; }
    }
    __syncthreads();
    buckets[t] = c[t];
    return;
  }
  const int stride = gridDim.x * blockDim.x;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_own; i += stride) {
    const unsigned long long sz = (unsigned long long)(boff[i + 1] - boff[i]) + 1ull;
    const int b = __clzll((long long)sz);                 // 63 - floor(log2(sz)); sz >= 1
    if (PASS == 0) atomicAdd(&buckets[b], 1u);
    else order[atomicAdd(&buckets[b], 1u)] = i;
  }
}

// single-block exclusive scan of the per-owner counts (uint32 -> int64 offsets[n_own + 1])
__global__ void __launch_bounds__(1024)
owner_scan_kernel(const unsigned int *__restrict__ counts, int n_own, long long *__restrict__ offsets) {
  __shared__ unsigned long long part[1024];
  const int t = threadIdx.x;
  const int per = (n_own + 1023) / 1024;
  const int lo = min(n_own, t * per), hi = min(n_own, lo + per);
  unsigned long long sum = 0;
  for (int i = lo; i < hi; ++i) sum += counts[i];
  part[t] = sum;
  __syncthreads();
  for (int o = 1; o < 1024; o <<= 1) {
    unsigned long long y = 0;
    if (t >= o) y = part[t - o];
    __syncthreads();
    part[t] += y;
    __syncthreads();
  }
  unsigned long long run = part[t] - sum;
  for (int i = lo; i < hi; ++i) { offsets[i] = (long long)run; run += counts[i]; }
  if (t == 1023) offsets[n_own] = (long long)part[1023];
}

// padded slot of owner v -> compact position; fixed-point sums -> fp32 scores on the way.
// Work is cut by COMPACT position (tiles of CP_TILE candidates), not by owner, so a hub with 10^5
// candidates does not serialise on one CTA; the owner of a position is found by binary search over the
// compact offsets of the (few) owners the tile touches.
constexpr int CP_THREADS = 256, CP_PER = 8, CP_TILE = CP_THREADS * CP_PER;

__global__ void __launch_bounds__(CP_THREADS)
twohop_compact_kernel(const long long *__restrict__ pad_off, const long long *__restrict__ cmp_off,
                      int v_lo, int n_own, long long cap, const int *__restrict__ pad_u,
                      const unsigned long long *__restrict__ acc, const int *__restrict__ cn, int flags,
                      int *__restrict__ pair_u, int *__restrict__ pair_v, float *__restrict__ score,
                      int *__restrict__ count) {
  const long long N = cmp_off[n_own];
  if (N > cap) return;     // a violated bound poisoned N: the caller reports it, nothing is moved
  __shared__ int s_lo, s_hi;
  const long long ntiles = (N + CP_TILE - 1) / CP_TILE;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long base = tile * CP_TILE, end = min(base + (long long)CP_TILE, N);
    __syncthreads();
    if (threadIdx.x < 2) {
      // largest o with cmp_off[o] <= x  (x = first / last position of the tile)
      const long long x = threadIdx.x == 0 ? base : end - 1;
      int lo = 0, hi = n_own;                       // invariant: cmp_off[lo] <= x < cmp_off[hi]
      while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (cmp_off[mid] <= x) lo = mid; else hi = mid;
      }
      if (threadIdx.x == 0) s_lo = lo; else s_hi = lo;
    }
    __syncthreads();
    const int o_lo = s_lo, o_hi = s_hi;
    long long src[CP_PER];
    int own[CP_PER];
#pragma unroll
    for (int q = 0; q < CP_PER; ++q) {
      const long long i = base + q * CP_THREADS + threadIdx.x;
      int lo = o_lo, hi = o_hi + 1;
      if (i < end) {
        while (hi - lo > 1) {
          const int mid = (lo + hi) >> 1;
          if (cmp_off[mid] <= i) lo = mid; else hi = mid;
        }
      }
      own[q] = lo;
      src[q] = i < end ? pad_off[lo] + (i - cmp_off[lo]) : -1;
    }
    int uu[CP_PER], cc[CP_PER];
    unsigned long long aa[CP_PER];
#pragma unroll
    for (int q = 0; q < CP_PER; ++q) {
      uu[q] = 0; cc[q] = 0; aa[q] = 0ull;
      if (src[q] >= 0) {
        uu[q] = pad_u[src[q]];
        if (acc) aa[q] = acc[src[q]];
        if (cn) cc[q] = cn[src[q]];
      }
    }
#pragma unroll
    for (int q = 0; q < CP_PER; ++q) {
      if (src[q] < 0) continue;
      const long long i = base + q * CP_THREADS + threadIdx.x;
      pair_u[i] = uu[q];
      pair_v[i] = v_lo + own[q];
      if (score) {
        float sc = acc ? from_fixed(aa[q]) : (float)cc[q];
        if (flags & EPS_CN_SIGMOID) sc = sigmoidf_ref(sc);
        score[i] = sc;
      }
      if (count) count[i] = cc[q];
    }
  }
}

}  // namespace eps

extern "C" size_t eps_twohop_scored_workspace_bytes(int64_t N) {
  return 256 + (size_t)(N > 0 ? N : 0) * 8;
}

extern "C" int eps_twohop_scored(const int32_t *rowptr, const int32_t *col, const float *wtable,
                                 int32_t n, int32_t v_lo, int32_t v_hi, const int64_t *offsets,
                                 int64_t N, int flags, int32_t *pair_u, int32_t *pair_v, float *score,
                                 int32_t *count, void *workspace, size_t workspace_bytes,
                                 void *stream_) {
  using namespace eps;
  cudaStream_t stream = (cudaStream_t)stream_;
  EPS_CHECK_ARG(rowptr && col && offsets, "null graph or offsets pointer");
  EPS_CHECK_ARG(n > 0 && v_lo >= 0 && v_hi <= n && v_lo <= v_hi && N >= 0, "bad owner range or N");
  if (v_lo == v_hi || N == 0) return EPS_OK;
  EPS_CHECK_ARG(pair_u && pair_v && (score || count), "missing output pointer");
  const int sms = sm_count();
  if (sms <= 0) { set_error("eps_twohop_scored: no CUDA device"); return EPS_ERR_CUDA; }
  if (!workspace || workspace_bytes < eps_twohop_scored_workspace_bytes(N)) {
    set_error("eps_twohop_scored: workspace too small");
    return EPS_ERR_WORKSPACE;
  }
  const int W = (n + 31) / 32, W2 = (W + 31) / 32;
  const size_t smem = (size_t)(W + 2 * W2) * 4 + (size_t)W * 2 + 16;
  if (smem > 200 * 1024) {
    set_error("eps_twohop_scored: n=%d needs %zu bytes of shared memory (> 200 KB)", n, smem);
    return EPS_ERR_UNSUPPORTED;
  }
  unsigned long long *acc = nullptr;
  int *cn = count;
  uint8_t *ws = (uint8_t *)workspace;
  EPS_CUDA(cudaMemsetAsync(ws, 0, 4, stream));
  if (wtable) {
    acc = (unsigned long long *)(ws + 256);
    EPS_CUDA(cudaMemsetAsync(acc, 0, (size_t)N * 8, stream));
  } else if (!cn) {
    cn = (int *)(ws + 256);          // CN scores only: count lives in the workspace
  }
  if (cn) EPS_CUDA(cudaMemsetAsync(cn, 0, (size_t)N * 4, stream));
  void (*kern)(const int *, const int *, const float *, const float *, int, int, int, const long long *, int *,
               int *, unsigned long long *, int *, unsigned int *, unsigned int *, const int *, int);
  if (wtable) kern = cn ? twohop_score_kernel<true, true, false> : twohop_score_kernel<true, false, false>;
  else kern = twohop_score_kernel<false, true, false>;
  EPS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int occ = 0;
  EPS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, TS_THREADS, smem));
  if (occ < 1) occ = 1;
  const int grid = (int)std::min<long long>((long long)(v_hi - v_lo), (long long)sms * occ);
  kern<<<grid, TS_THREADS, smem, stream>>>(rowptr, col, nullptr, wtable, n, v_lo, v_hi,
                                           (const long long *)offsets, pair_u, pair_v, acc, cn,
                                           (unsigned int *)ws, nullptr, nullptr, 0);
  EPS_LAUNCH_CHECK();
  if (score) {
    const int fgrid = (int)std::min<long long>((N + 255) / 256, (long long)sms * 8);
    twohop_finalize_kernel<<<fgrid, 256, 0, stream>>>(acc, cn, (long long)N, flags, score);
    EPS_LAUNCH_CHECK();
  }
  return EPS_OK;
}

// ---------------------------------------------------------------------------------------------
// one-pass entry: enumerate (+ score) the owner range without a count pass
// ---------------------------------------------------------------------------------------------
static inline size_t up256(size_t x) { return (x + 255) & ~(size_t)255; }

extern "C" size_t eps_twohop_onepass_workspace_bytes(int64_t cap, int32_t n_owners) {
  const size_t c = (size_t)(cap > 0 ? cap : 0), o = (size_t)(n_owners > 0 ? n_owners : 0);
  // ticket | work buckets | per-owner counts | owner order | padded u | padded CN counts | padded fixed-point sums
  return 512 + 2 * up256(o * 4) + up256(c * 4) + up256(c * 4) + up256(c * 8);
}

extern "C" int eps_twohop_onepass(const int32_t *rowptr, const int32_t *col, const float *val,
                                  const float *wtable, int32_t n, int32_t v_lo, int32_t v_hi, const int64_t *bound_offsets,
                                  int64_t cap, int flags, int32_t *pair_u, int32_t *pair_v, float *score,
                                  int32_t *count, int64_t *offsets_out, void *workspace,
                                  size_t workspace_bytes, void *stream_) {
  using namespace eps;
  cudaStream_t stream = (cudaStream_t)stream_;
  EPS_CHECK_ARG(rowptr && col && bound_offsets && offsets_out, "null graph, bound_offsets or offsets_out pointer");
  EPS_CHECK_ARG(n > 0 && v_lo >= 0 && v_hi <= n && v_lo < v_hi && cap >= 0, "bad owner range or cap");
  EPS_CHECK_ARG(cap == 0 || (pair_u && pair_v), "missing pair output pointer");
  EPS_CHECK_ARG(!(wtable && !score), "wtable given but no score output");
  EPS_CHECK_ARG(!(val && !score), "edge values given but no score output");
  const int sms = sm_count();
  if (sms <= 0) { set_error("eps_twohop_onepass: no CUDA device"); return EPS_ERR_CUDA; }
  const int n_own = v_hi - v_lo;
  if (!workspace || workspace_bytes < eps_twohop_onepass_workspace_bytes(cap, n_own)) {
    set_error("eps_twohop_onepass: workspace too small");
    return EPS_ERR_WORKSPACE;
  }
  const int W = (n + 31) / 32, W2 = (W + 31) / 32;
  const size_t smem = (size_t)(W + 2 * W2) * 4 + (size_t)W * 2 + 16;
  if (smem > 200 * 1024) {
    set_error("eps_twohop_onepass: n=%d needs %zu bytes of shared memory (> 200 KB)", n, smem);
    return EPS_ERR_UNSUPPORTED;
  }
  uint8_t *ws = (uint8_t *)workspace;
  unsigned int *buckets = (unsigned int *)(ws + 256);
  unsigned int *counts = (unsigned int *)(ws + 512);
  int *order = (int *)(ws + 512 + up256((size_t)n_own * 4));
  int *pad_u = (int *)(ws + 512 + 2 * up256((size_t)n_own * 4));
  int *pad_cn = (int *)((uint8_t *)pad_u + up256((size_t)cap * 4));
  unsigned long long *pad_acc = (unsigned long long *)((uint8_t *)pad_cn + up256((size_t)cap * 4));
  EPS_CUDA(cudaMemsetAsync(ws, 0, 512 + up256((size_t)n_own * 4), stream));   // ticket + buckets + counts
  const bool want_score = score != nullptr || count != nullptr;
  unsigned long long *acc = nullptr;
  int *cn = nullptr;
  if (wtable || val) {
    acc = pad_acc;
    EPS_CUDA(cudaMemsetAsync(acc, 0, (size_t)cap * 8, stream));
    if (count) cn = pad_cn;
  } else if (want_score) {
    cn = pad_cn;
  }
  if (cn) EPS_CUDA(cudaMemsetAsync(cn, 0, (size_t)cap * 4, stream));
  void (*kern)(const int *, const int *, const float *, const float *, int, int, int, const long long *, int *,
               int *, unsigned long long *, int *, unsigned int *, unsigned int *, const int *, int);
  // CTA size.  Measured (gpurun_out/r31, fused phase per step): 1024 threads help only where the bitmap limits
  // the kernel to 2 CTAs per SM (ppa 12.6 -> 11.6 ms) and hurt where 512-thread CTAs already fill the SM
  // (collab 4.6 -> 7.7 ms; ddi 8.1 -> 8.2): 512 stays the default, EPS_TS_THREADS=1024 overrides (A/B runs)
  int threads = TS_DEFAULT_THREADS;
  if (const char *e = getenv("EPS_TS_THREADS")) threads = atoi(e) == 1024 ? 1024 : 512;
#define EPS_TS_PICK(T)                                                                                                  \
  do {                                                                                                                  \
    if (!want_score) kern = twohop_score_kernel<false, false, true, false, T>;                                         \
    else if (val && wtable) kern = cn ? twohop_score_kernel<true, true, true, true, T> : twohop_score_kernel<true, false, true, true, T>;   \
    else if (val) kern = cn ? twohop_score_kernel<false, true, true, true, T> : twohop_score_kernel<false, false, true, true, T>;           \
    else if (wtable) kern = cn ? twohop_score_kernel<true, true, true, false, T> : twohop_score_kernel<true, false, true, false, T>;        \
    else kern = twohop_score_kernel<false, true, true, false, T>;                                                      \
  } while (0)
  if (threads == 1024) EPS_TS_PICK(1024); else EPS_TS_PICK(512);
#undef EPS_TS_PICK
  EPS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int occ = 0;
  EPS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, smem));
  if (occ < 1) occ = 1;
  const int grid = (int)std::min<long long>((long long)n_own, (long long)sms * occ);
  {  // heaviest-first hand-out: counting sort of the owners by the power of two of their slot size (the bound)
    const int ogrid = std::min((n_own + 255) / 256, sms * 8);
    owner_bucket_kernel<0><<<ogrid, 256, 0, stream>>>((const long long *)bound_offsets, n_own, buckets, order);
    owner_bucket_kernel<1><<<1, 64, 0, stream>>>((const long long *)bound_offsets, n_own, buckets, order);
    owner_bucket_kernel<2><<<ogrid, 256, 0, stream>>>((const long long *)bound_offsets, n_own, buckets, order);
    EPS_LAUNCH_CHECK();
  }
  kern<<<grid, threads, smem, stream>>>(rowptr, col, val, wtable, n, v_lo, v_hi,
                                           (const long long *)bound_offsets, pad_u, nullptr, acc, cn,
                                           (unsigned int *)ws, counts, order, n_own);
  EPS_LAUNCH_CHECK();
  owner_scan_kernel<<<1, 1024, 0, stream>>>(counts, n_own, (long long *)offsets_out);
  EPS_LAUNCH_CHECK();
  if (cap > 0) {
    const int cgrid = (int)std::min<long long>((cap + CP_TILE - 1) / CP_TILE, (long long)sms * 8);
    twohop_compact_kernel<<<cgrid, CP_THREADS, 0, stream>>>((const long long *)bound_offsets,
                                                     (const long long *)offsets_out, v_lo, n_own, (long long)cap,
                                                     pad_u, acc,
                                                     cn, flags, pair_u, pair_v, score, count);
    EPS_LAUNCH_CHECK();
  }
  return EPS_OK;
}
