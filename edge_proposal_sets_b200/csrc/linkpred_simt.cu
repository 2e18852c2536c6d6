// K2 (fp32 arm) — fused gather + Hadamard + LinkPredictor MLP in fp32 on the CUDA cores.
//
// Replaces /root/reference/models.py:506 (h[edges[0]], h[edges[1]] gathers) and
// models.py:478-485 (mul, (Linear, ReLU) x (L-1), Linear(H->1), sigmoid), which the reference
// runs as 2L+3 separate ATen / cuBLAS launches per batch.  This arm keeps the reference's fp32
// arithmetic (fp32 products, fp32 accumulate, k ascending) and is the precision baseline the
// tcgen05 arm (linkpred_tc.cu) is measured against; it is bounded by the fp32 FFMA rate, not by
// HBM.  One CTA owns a tile of 32 pairs; activations ping-pong between two shared-memory
// buffers, weights stream through L1/L2 (256 KB per layer, L2-resident).  A thread computes TWO output
// features for all 32 rows: every broadcast LDS.128 of an activation piece then feeds 8 FFMAs instead of 4 —
// with one feature per thread the shared-memory pipe (32 wavefronts per 128 FFMA) was as busy as the FFMA pipe.
// Each (row, feature) sum is still ONE fmaf chain in ascending k: the scores are bit-identical to round 1.
#include "eps_common.cuh"

namespace eps {

constexpr int MLP_THREADS = 128;
constexpr int MLP_BM = 32;

__global__ void __launch_bounds__(MLP_THREADS)
linkpred_fp32_kernel(const float *__restrict__ h, int H, const int *__restrict__ pu,
                     const int *__restrict__ pv, long long M, const MlpParams prm, int L,
                     int apply_sigmoid, float *__restrict__ score) {
  const float *const *Wd = prm.W;
  const float *const *bd = prm.b;
  extern __shared__ __align__(16) float smem[];
  float *Z0 = smem;
  float *Z1 = smem + MLP_BM * H;
  const int tid = threadIdx.x, lane = lane_id(), warp = tid >> 5;
  const long long ntiles = (M + MLP_BM - 1) / MLP_BM;
  const int H4 = H >> 2;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long p0 = tile * MLP_BM;
    const int rows = (int)min((long long)MLP_BM, M - p0);
    // gather + Hadamard:  Z0[r,:] = h[u_r,:] * h[v_r,:]
    for (int r = warp; r < MLP_BM; r += MLP_THREADS / 32) {
      float4 *dst = reinterpret_cast<float4 *>(Z0 + r * H);
      if (r < rows) {
        const float4 *hu = reinterpret_cast<const float4 *>(h + (size_t)pu[p0 + r] * H);
        const float4 *hv = reinterpret_cast<const float4 *>(h + (size_t)pv[p0 + r] * H);
        for (int c = lane; c < H4; c += 32) {
          const float4 a = __ldg(hu + c), b = __ldg(hv + c);
          dst[c] = make_float4(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y), __fmul_rn(a.z, b.z),
                               __fmul_rn(a.w, b.w));
        }
      } else {
        for (int c = lane; c < H4; c += 32) dst[c] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    __syncthreads();
    float *zin = Z0, *zout = Z1;
    for (int l = 0; l < L - 1; ++l) {
      const float *W = Wd[l];
      const float *b = bd[l];
      for (int j = tid; j < H; j += 2 * MLP_THREADS) {
        const int j1 = j + MLP_THREADS;                       // second feature of this thread (if it exists)
        const bool two = j1 < H;
        float acc0[MLP_BM], acc1[MLP_BM];
#pragma unroll
        for (int r = 0; r < MLP_BM; ++r) { acc0[r] = 0.f; acc1[r] = 0.f; }
        const float4 *wrow0 = reinterpret_cast<const float4 *>(W + (size_t)j * H);
        const float4 *wrow1 = reinterpret_cast<const float4 *>(W + (size_t)(two ? j1 : j) * H);
        for (int k4 = 0; k4 < H4; ++k4) {
          const float4 w0 = __ldg(wrow0 + k4), w1 = __ldg(wrow1 + k4);
#pragma unroll
          for (int r = 0; r < MLP_BM; ++r) {
            const float4 a = *reinterpret_cast<const float4 *>(zin + r * H + k4 * 4);
            acc0[r] = fmaf(a.x, w0.x, acc0[r]);
            acc0[r] = fmaf(a.y, w0.y, acc0[r]);
            acc0[r] = fmaf(a.z, w0.z, acc0[r]);
            acc0[r] = fmaf(a.w, w0.w, acc0[r]);
            acc1[r] = fmaf(a.x, w1.x, acc1[r]);
            acc1[r] = fmaf(a.y, w1.y, acc1[r]);
            acc1[r] = fmaf(a.z, w1.z, acc1[r]);
            acc1[r] = fmaf(a.w, w1.w, acc1[r]);
          }
        }
        const float b0 = __ldg(b + j), b1 = __ldg(b + (two ? j1 : j));
#pragma unroll
        for (int r = 0; r < MLP_BM; ++r) {
          zout[r * H + j] = fmaxf(__fadd_rn(acc0[r], b0), 0.f);
          if (two) zout[r * H + j1] = fmaxf(__fadd_rn(acc1[r], b1), 0.f);
        }
      }
      __syncthreads();
      float *t = zin; zin = zout; zout = t;
    }
    // last layer H -> 1, one warp per row
    const float *W = Wd[L - 1];
    const float bl = __ldg(bd[L - 1]);
    for (int r = warp; r < rows; r += MLP_THREADS / 32) {
      float s = 0.f;
      for (int k = lane; k < H; k += 32) s = fmaf(zin[r * H + k], __ldg(W + k), s);
#pragma unroll
      for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(FULL, s, o);
      if (lane == 0) {
        s = __fadd_rn(s, bl);
        score[p0 + r] = apply_sigmoid ? sigmoidf_ref(s) : s;
      }
    }
    __syncthreads();
  }
}

int linkpred_fp32_launch(const float *h, int H, const int *pu, const int *pv, long long M,
                         const MlpParams &prm, int L, int apply_sigmoid, float *score,
                         cudaStream_t stream) {
  const size_t smem = (size_t)2 * MLP_BM * H * sizeof(float);
  EPS_CUDA(cudaFuncSetAttribute(linkpred_fp32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)smem));
  int occ = 0;
  EPS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, linkpred_fp32_kernel, MLP_THREADS, smem));
  if (occ < 1) occ = 1;
  const long long ntiles = (M + MLP_BM - 1) / MLP_BM;
  const int grid = (int)std::min<long long>(ntiles, (long long)sm_count() * occ);
  linkpred_fp32_kernel<<<grid, MLP_THREADS, smem, stream>>>(h, H, pu, pv, M, prm, L,
                                                            apply_sigmoid, score);
  EPS_LAUNCH_CHECK();
  return EPS_OK;
}

}  // namespace eps
