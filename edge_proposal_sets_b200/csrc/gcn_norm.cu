// K1b — GCN normalisation of the adjacency, built once per graph on the device.
//
// Replaces torch_geometric's gcn_norm as called from GCNConv.forward (reached from
// /root/reference/models.py:183,186 on EVERY forward because the reference's GCNConv has
// cached=False): fill_diag(adj, 1) — the diagonal is SET to 1, inserted where missing —,
// deg = rowsum, dinv = deg^-1/2 (inf -> 0), value = (w * dinv[row]) * dinv[col] with two fp32
// roundings in that order.  Two passes over the CSR (one warp per row), a prefix sum in between:
//   count: dinv[i], new row length (deg + 1 if the row has no diagonal entry)
//   fill : columns with the diagonal merged at its sorted position, normalised values
#include "eps_common.cuh"

namespace eps {

__global__ void __launch_bounds__(256)
gcn_norm_count_kernel(const int *__restrict__ rowptr, const int *__restrict__ col,
                      const float *__restrict__ val, int n, float *__restrict__ dinv,
                      int *__restrict__ newlen) {
  const int lane = lane_id();
  const int warps = (gridDim.x * blockDim.x) >> 5;
  for (int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < n; i += warps) {
    const int s = __ldg(rowptr + i), e = __ldg(rowptr + i + 1);
    float sum = 0.f;   // sum of off-diagonal weights; integer-valued in every reference dataset
    int has = 0;
    for (int p = s + lane; p < e; p += 32) {
      const int c = __ldg(col + p);
      if (c == i) has = 1;
      else sum += val ? __ldg(val + p) : 1.f;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      sum += __shfl_xor_sync(FULL, sum, o);
      has |= __shfl_xor_sync(FULL, has, o);
    }
    if (lane == 0) {
      const float deg = sum + 1.f;                     // diagonal set to 1
      const float d = 1.0f / sqrtf(deg);
      dinv[i] = isinf(d) ? 0.f : d;
      newlen[i] = (e - s) + (has ? 0 : 1);
    }
  }
}

__global__ void __launch_bounds__(256)
gcn_norm_fill_kernel(const int *__restrict__ rowptr, const int *__restrict__ col,
                     const float *__restrict__ val, int n, const float *__restrict__ dinv,
                     const int *__restrict__ rowptr2, int *__restrict__ col2, float *__restrict__ val2) {
  const int lane = lane_id();
  const int warps = (gridDim.x * blockDim.x) >> 5;
  for (int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < n; i += warps) {
    const int s = __ldg(rowptr + i), e = __ldg(rowptr + i + 1);
    const int s2 = __ldg(rowptr2 + i);
    const bool missing = (__ldg(rowptr2 + i + 1) - s2) != (e - s);
    const float di = __ldg(dinv + i);
    int lt = 0;   // entries with column < i seen so far (only needed when the diagonal is missing)
    for (int base = s; base < e; base += 32) {
      const int p = base + lane;
      int c = 0x7fffffff;
      float w = 1.f;
      if (p < e) { c = __ldg(col + p); if (val) w = __ldg(val + p); }
      const unsigned below = __ballot_sync(FULL, p < e && c < i);
      if (p < e) {
        if (c == i) w = 1.f;                            // fill_diag: set, not add
        const int shift = (missing && c > i) ? 1 : 0;
        col2[s2 + (p - s) + shift] = c;
        val2[s2 + (p - s) + shift] = __fmul_rn(__fmul_rn(w, di), __ldg(dinv + c));
      }
      lt += __popc(below);
    }
    if (missing && lane == 0) {
      col2[s2 + lt] = i;
      val2[s2 + lt] = __fmul_rn(__fmul_rn(1.f, di), di);
    }
  }
}

}  // namespace eps

extern "C" int eps_gcn_norm_count(const int32_t *rowptr, const int32_t *col, const float *val, int32_t n,
                                  float *dinv, int32_t *newlen, void *stream_) {
  using namespace eps;
  EPS_CHECK_ARG(n >= 0, "bad n");
  if (n == 0) return EPS_OK;
  EPS_CHECK_ARG(rowptr && col && dinv && newlen, "null pointer");
  const int sms = sm_count();
  if (sms <= 0) { set_error("eps_gcn_norm_count: no CUDA device"); return EPS_ERR_CUDA; }
  const int grid = (int)std::min<long long>(((long long)n + 7) / 8, (long long)sms * 8);
  gcn_norm_count_kernel<<<grid, 256, 0, (cudaStream_t)stream_>>>(rowptr, col, val, n, dinv, newlen);
  EPS_LAUNCH_CHECK();
  return EPS_OK;
}

extern "C" int eps_gcn_norm_fill(const int32_t *rowptr, const int32_t *col, const float *val, int32_t n,
                                 const float *dinv, const int32_t *rowptr2, int32_t *col2, float *val2,
                                 void *stream_) {
  using namespace eps;
  EPS_CHECK_ARG(n >= 0, "bad n");
  if (n == 0) return EPS_OK;
  EPS_CHECK_ARG(rowptr && col && dinv && rowptr2 && col2 && val2, "null pointer");
  const int sms = sm_count();
  if (sms <= 0) { set_error("eps_gcn_norm_fill: no CUDA device"); return EPS_ERR_CUDA; }
  const int grid = (int)std::min<long long>(((long long)n + 7) / 8, (long long)sms * 8);
  gcn_norm_fill_kernel<<<grid, 256, 0, (cudaStream_t)stream_>>>(rowptr, col, val, n, dinv, rowptr2, col2, val2);
  EPS_LAUNCH_CHECK();
  return EPS_OK;
}
