// K1 — CSR SpMM for GCNConv / SAGEConv neighbour aggregation.
//
// Replaces torch_sparse's spmm kernel reached from /root/reference/models.py:183,186 (GCNConv,
// reduce='add' with normalised edge values) and :436,439 (SAGEConv, reduce='mean', no values).
//
// One warp per output row, rows handed out dynamically (atomic counter) so the power-law degree
// tail does not serialise a static partition.  The warp reads the row's (col, val) 32 at a time
// with one coalesced load, broadcasts them by shuffle and gathers the X rows with 128-bit loads:
// for F = 256 every neighbour is one fully coalesced 1 KB read (2 x float4 per lane).  Four
// neighbour rows are kept in flight per lane, but the accumulation itself stays a strict
// left-to-right fmaf chain in ascending column order — the order of torch_sparse's kernel
// (SURVEY A.3), so the result does not depend on launch geometry.
// Epilogue fused: mean division (SAGE), + bias, ReLU.
#include "eps_common.cuh"

namespace eps {

constexpr int SPMM_THREADS = 256;

template <bool HAS_VAL, int VEC /*float4 per lane per feature tile*/>
__device__ __forceinline__ void spmm_row_vec(const int *__restrict__ col, const float *__restrict__ val,
                                             const float *__restrict__ X, float *__restrict__ Y,
                                             int row, int start, int end, int F, int f0, int reduce,
                                             const float *__restrict__ bias, int relu) {
  // this lane owns features f0 + (q*32 + lane)*4 .. +3, q < VEC
  const int lane = lane_id();
  float4 acc[VEC];
  bool act[VEC];
#pragma unroll
  for (int q = 0; q < VEC; ++q) {
    acc[q] = make_float4(0.f, 0.f, 0.f, 0.f);
    act[q] = f0 + (q * 32 + lane) * 4 < F;
  }
  for (int base = start; base < end; base += 32) {
    int c = 0;
    float v = 1.f;
    if (base + lane < end) {
      c = __ldg(col + base + lane);
      if (HAS_VAL) v = __ldg(val + base + lane);
    }
    const int cnt = min(32, end - base);
    int t = 0;
    for (; t + 4 <= cnt; t += 4) {
      int cc[4];
      float vv[4];
      float4 x[4][VEC];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        cc[u] = __shfl_sync(FULL, c, t + u);
        vv[u] = __shfl_sync(FULL, v, t + u);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float4 *xr = reinterpret_cast<const float4 *>(X + (size_t)cc[u] * F + f0);
#pragma unroll
        for (int q = 0; q < VEC; ++q)
          x[u][q] = act[q] ? __ldg(xr + q * 32 + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
#pragma unroll
        for (int q = 0; q < VEC; ++q) {
          acc[q].x = fmaf(vv[u], x[u][q].x, acc[q].x);
          acc[q].y = fmaf(vv[u], x[u][q].y, acc[q].y);
          acc[q].z = fmaf(vv[u], x[u][q].z, acc[q].z);
          acc[q].w = fmaf(vv[u], x[u][q].w, acc[q].w);
        }
      }
    }
    for (; t < cnt; ++t) {
      const int cc = __shfl_sync(FULL, c, t);
      const float vv = __shfl_sync(FULL, v, t);
      const float4 *xr = reinterpret_cast<const float4 *>(X + (size_t)cc * F + f0);
#pragma unroll
      for (int q = 0; q < VEC; ++q) {
        if (act[q]) {
          const float4 x = __ldg(xr + q * 32 + lane);
          acc[q].x = fmaf(vv, x.x, acc[q].x);
          acc[q].y = fmaf(vv, x.y, acc[q].y);
          acc[q].z = fmaf(vv, x.z, acc[q].z);
          acc[q].w = fmaf(vv, x.w, acc[q].w);
        }
      }
    }
  }
  const float denom = (float)max(end - start, 1);
#pragma unroll
  for (int q = 0; q < VEC; ++q) {
    if (!act[q]) continue;
    const int f = f0 + (q * 32 + lane) * 4;
    float4 r = acc[q];
    if (reduce == EPS_REDUCE_MEAN) {
      r.x = __fdiv_rn(r.x, denom); r.y = __fdiv_rn(r.y, denom);
      r.z = __fdiv_rn(r.z, denom); r.w = __fdiv_rn(r.w, denom);
    }
    if (bias) {
      const float4 b = __ldg(reinterpret_cast<const float4 *>(bias + f));
      r.x = __fadd_rn(r.x, b.x); r.y = __fadd_rn(r.y, b.y);
      r.z = __fadd_rn(r.z, b.z); r.w = __fadd_rn(r.w, b.w);
    }
    if (relu) {
      r.x = fmaxf(r.x, 0.f); r.y = fmaxf(r.y, 0.f); r.z = fmaxf(r.z, 0.f); r.w = fmaxf(r.w, 0.f);
    }
    *reinterpret_cast<float4 *>(Y + (size_t)row * F + f) = r;
  }
}

// scalar path for F % 4 != 0 (rows are then not 16-byte aligned): lane owns features f0+lane+32q
template <bool HAS_VAL>
__device__ __forceinline__ void spmm_row_scalar(const int *__restrict__ col, const float *__restrict__ val,
                                                const float *__restrict__ X, float *__restrict__ Y,
                                                int row, int start, int end, int F, int f0, int reduce,
                                                const float *__restrict__ bias, int relu) {
  constexpr int Q = 4;  // 128 features per pass
  const int lane = lane_id();
  float acc[Q];
#pragma unroll
  for (int q = 0; q < Q; ++q) acc[q] = 0.f;
  for (int base = start; base < end; base += 32) {
    int c = 0;
    float v = 1.f;
    if (base + lane < end) {
      c = __ldg(col + base + lane);
      if (HAS_VAL) v = __ldg(val + base + lane);
    }
    const int cnt = min(32, end - base);
    for (int t = 0; t < cnt; ++t) {
      const int cc = __shfl_sync(FULL, c, t);
      const float vv = __shfl_sync(FULL, v, t);
      const float *xr = X + (size_t)cc * F + f0;
#pragma unroll
      for (int q = 0; q < Q; ++q) {
        const int f = f0 + q * 32 + lane;
        if (f < F) acc[q] = fmaf(vv, __ldg(xr + q * 32 + lane), acc[q]);
      }
    }
  }
  const float denom = (float)max(end - start, 1);
#pragma unroll
  for (int q = 0; q < Q; ++q) {
    const int f = f0 + q * 32 + lane;
    if (f >= F) continue;
    float r = acc[q];
    if (reduce == EPS_REDUCE_MEAN) r = __fdiv_rn(r, denom);
    if (bias) r = __fadd_rn(r, __ldg(bias + f));
    if (relu) r = fmaxf(r, 0.f);
    Y[(size_t)row * F + f] = r;
  }
}

template <bool HAS_VAL, bool VECTOR>
__global__ void __launch_bounds__(SPMM_THREADS)
spmm_csr_kernel(const int *__restrict__ rowptr, const int *__restrict__ col,
                const float *__restrict__ val, const float *__restrict__ X, float *__restrict__ Y,
                int n_rows, int F, int reduce, const float *__restrict__ bias, int relu,
                unsigned int *row_counter) {
  const int lane = lane_id();
  for (;;) {
    int row = 0;
    if (lane == 0) row = (int)atomicAdd(row_counter, 1u);
    row = __shfl_sync(FULL, row, 0);
    if (row >= n_rows) break;
    const int start = __ldg(rowptr + row), end = __ldg(rowptr + row + 1);
    if (VECTOR) {
      int f0 = 0;
      for (; f0 + 256 <= F; f0 += 256)
        spmm_row_vec<HAS_VAL, 2>(col, val, X, Y, row, start, end, F, f0, reduce, bias, relu);
      if (f0 + 128 < F)
        spmm_row_vec<HAS_VAL, 2>(col, val, X, Y, row, start, end, F, f0, reduce, bias, relu);
      else if (f0 < F)
        spmm_row_vec<HAS_VAL, 1>(col, val, X, Y, row, start, end, F, f0, reduce, bias, relu);
    } else {
      for (int f0 = 0; f0 < F; f0 += 128)
        spmm_row_scalar<HAS_VAL>(col, val, X, Y, row, start, end, F, f0, reduce, bias, relu);
    }
  }
}

}  // namespace eps

extern "C" size_t eps_spmm_workspace_bytes(void) { return 256; }

extern "C" int eps_spmm_csr_f32(const int32_t *rowptr, const int32_t *col, const float *val,
                                const float *X, float *Y, int32_t n_rows, int32_t F, int reduce,
                                const float *bias, int relu, void *workspace,
                                size_t workspace_bytes, void *stream_) {
  using namespace eps;
  cudaStream_t stream = (cudaStream_t)stream_;
  EPS_CHECK_ARG(rowptr && col && X && Y, "null pointer");
  EPS_CHECK_ARG(X != Y, "X and Y must not alias");
  EPS_CHECK_ARG(n_rows >= 0 && F >= 1, "bad shape");
  EPS_CHECK_ARG(reduce == EPS_REDUCE_SUM || reduce == EPS_REDUCE_MEAN, "bad reduce");
  if (n_rows == 0) return EPS_OK;
  const int sms = sm_count();
  if (sms <= 0) { set_error("eps_spmm_csr_f32: no CUDA device"); return EPS_ERR_CUDA; }
  if (!workspace || workspace_bytes < eps_spmm_workspace_bytes()) {
    set_error("eps_spmm_csr_f32: workspace too small");
    return EPS_ERR_WORKSPACE;
  }
  EPS_CUDA(cudaMemsetAsync(workspace, 0, 4, stream));
  const bool vec = (F % 4 == 0) && (((uintptr_t)X | (uintptr_t)Y | (uintptr_t)bias) % 16 == 0);
  const int warps_per_block = SPMM_THREADS / 32;
  const long long want = ((long long)n_rows + warps_per_block - 1) / warps_per_block;
  const int grid = (int)std::min<long long>(want, (long long)sms * 8);
  unsigned int *ctr = (unsigned int *)workspace;
#define EPS_SPMM(HV, VC)                                                                       \
  spmm_csr_kernel<HV, VC><<<grid, SPMM_THREADS, 0, stream>>>(rowptr, col, val, X, Y, n_rows, F, \
                                                             reduce, bias, relu, ctr)
  if (val) { if (vec) EPS_SPMM(true, true); else EPS_SPMM(true, false); }
  else     { if (vec) EPS_SPMM(false, true); else EPS_SPMM(false, false); }
#undef EPS_SPMM
  EPS_LAUNCH_CHECK();
  return EPS_OK;
}
