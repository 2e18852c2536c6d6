// K7 — the LinkPredictor's input in TRAINING mode: z0[b,:] = h[u_b,:] * h[v_b,:] and its backward.
//
// Replaces /root/reference/models.py:506 (h[edges[0]], h[edges[1]]: two index_select launches that materialise
// 2 x B x H floats) + the x_i * x_j of models.py:479 in the forward of /root/reference/train_and_eval.py:60-66, and in
// the backward the mul-backward + two index_select backwards (index_add with atomics) autograd runs for them.
//   forward   one pass: both rows are gathered with 128-bit loads and multiplied in registers, B x H floats written;
//   backward  dh[u_b,:] += dz[b,:] * h[v_b,:],  dh[v_b,:] += dz[b,:] * h[u_b,:]   (fp32 RED.ADD, like ATen's index_add)
// The dense layers behind z0 stay cuBLAS GEMMs (plain library GEMMs); the scoring path's fused kernel is K2.
#include "eps_common.cuh"

namespace eps {

constexpr int PH_THREADS = 256;

// one warp per pair, lanes stride the row in float4
template <bool BWD>
__global__ void __launch_bounds__(PH_THREADS)
pair_hadamard_kernel(const float *__restrict__ h, int H, const int *__restrict__ pu, const int *__restrict__ pv,
                     long long M, float *__restrict__ z /* fwd: out [M,H]; bwd: dz [M,H] */, float *__restrict__ dh) {
  const int lane = lane_id();
  const long long warp = ((long long)blockIdx.x * PH_THREADS + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * PH_THREADS) >> 5;
  const int H4 = H >> 2;
  for (long long b = warp; b < M; b += nwarps) {
    const int u = __ldg(pu + b), v = __ldg(pv + b);
    const float4 *hu = reinterpret_cast<const float4 *>(h + (size_t)u * H);
    const float4 *hv = reinterpret_cast<const float4 *>(h + (size_t)v * H);
    float4 *zr = reinterpret_cast<float4 *>(z + (size_t)b * H);
    for (int c = lane; c < H4; c += 32) {
      const float4 a = __ldg(hu + c), w = __ldg(hv + c);
      if (!BWD) {
        zr[c] = make_float4(__fmul_rn(a.x, w.x), __fmul_rn(a.y, w.y), __fmul_rn(a.z, w.z), __fmul_rn(a.w, w.w));
      } else {
        const float4 g = zr[c];
        float *du = dh + (size_t)u * H + 4 * c, *dv = dh + (size_t)v * H + 4 * c;
        atomicAdd(du + 0, g.x * w.x); atomicAdd(du + 1, g.y * w.y); atomicAdd(du + 2, g.z * w.z); atomicAdd(du + 3, g.w * w.w);
        atomicAdd(dv + 0, g.x * a.x); atomicAdd(dv + 1, g.y * a.y); atomicAdd(dv + 2, g.z * a.z); atomicAdd(dv + 3, g.w * a.w);
      }
    }
  }
}

}  // namespace eps

extern "C" int eps_pair_hadamard_f32(const float *h, int32_t n, int32_t H, const int32_t *pair_u, const int32_t *pair_v,
                                     int64_t M, float *out, void *stream) {
  using namespace eps;
  EPS_CHECK_ARG(n > 0 && H > 0 && (H & 3) == 0 && M >= 0, "bad n, H (multiple of 4) or M");
  if (M == 0) return EPS_OK;
  EPS_CHECK_ARG(h && pair_u && pair_v && out, "null pointer");
  const int sms = sm_count();
  if (sms <= 0) { set_error("eps_pair_hadamard_f32: no CUDA device"); return EPS_ERR_CUDA; }
  const int grid = (int)std::min<long long>((M * 32 + PH_THREADS - 1) / PH_THREADS, (long long)sms * 16);
  pair_hadamard_kernel<false><<<grid, PH_THREADS, 0, (cudaStream_t)stream>>>(h, H, pair_u, pair_v, M, out, nullptr);
  EPS_LAUNCH_CHECK();
  return EPS_OK;
}

extern "C" int eps_pair_hadamard_bwd_f32(const float *h, int32_t n, int32_t H, const int32_t *pair_u,
                                         const int32_t *pair_v, int64_t M, const float *dz, float *dh, void *stream) {
  using namespace eps;
  EPS_CHECK_ARG(n > 0 && H > 0 && (H & 3) == 0 && M >= 0, "bad n, H (multiple of 4) or M");
  if (M == 0) return EPS_OK;
  EPS_CHECK_ARG(h && pair_u && pair_v && dz && dh, "null pointer");
  const int sms = sm_count();
  if (sms <= 0) { set_error("eps_pair_hadamard_bwd_f32: no CUDA device"); return EPS_ERR_CUDA; }
  const int grid = (int)std::min<long long>((M * 32 + PH_THREADS - 1) / PH_THREADS, (long long)sms * 16);
  pair_hadamard_kernel<true><<<grid, PH_THREADS, 0, (cudaStream_t)stream>>>(h, H, pair_u, pair_v, M, const_cast<float *>(dz), dh);
  EPS_LAUNCH_CHECK();
  return EPS_OK;
}
