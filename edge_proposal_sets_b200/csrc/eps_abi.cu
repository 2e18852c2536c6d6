// C-ABI plumbing shared by all kernels: error slot, device query, K2 dispatcher.
#include <stdarg.h>

#include "eps_common.cuh"

namespace eps {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int sm_count() {
  static thread_local int cached_dev = -1;
  static thread_local int cached_sms = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return 0; }
  if (dev != cached_dev) {
    int sms = 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) {
      cudaGetLastError();
      return 0;
    }
    cached_dev = dev;
    cached_sms = sms;
  }
  return cached_sms;
}

}  // namespace eps

extern "C" int eps_version(void) { return EPS_VERSION; }
extern "C" const char *eps_last_error(void) { return eps::g_err; }

extern "C" size_t eps_linkpred_workspace_bytes(int32_t n, int32_t H, int32_t L, int64_t M, int precision) {
  precision &= ~EPS_MLP_REUSE_WORKSPACE;
  if (precision == EPS_MLP_TC_BF16) return eps::linkpred_tc_workspace_bytes(n, H, L, M);
  return 256;
}

extern "C" int eps_linkpred_mlp(const float *h, int32_t n, int32_t H, const int32_t *pair_u,
                                const int32_t *pair_v, int64_t M, const float *const *W_h,
                                const float *const *b_h, int32_t L, int precision,
                                int apply_sigmoid, float *score, void *workspace,
                                size_t workspace_bytes, void *stream_) {
  using namespace eps;
  cudaStream_t stream = (cudaStream_t)stream_;
  EPS_CHECK_ARG(n > 0 && M >= 0, "bad n or M");
  if (M == 0) return EPS_OK;
  EPS_CHECK_ARG(h && pair_u && pair_v && W_h && b_h && score, "null pointer");
  EPS_CHECK_ARG(L >= 1 && L <= EPS_MAX_MLP_LAYERS, "num_layers out of range [1,8]");
  EPS_CHECK_ARG(H >= 4 && H % 4 == 0 && H <= 512, "hidden size must be a multiple of 4, <= 512");
  EPS_CHECK_ARG(((uintptr_t)h) % 16 == 0, "h must be 16-byte aligned");
  if (M == 0) return EPS_OK;
  if (sm_count() <= 0) { set_error("eps_linkpred_mlp: no CUDA device"); return EPS_ERR_CUDA; }
  MlpParams prm;
  for (int l = 0; l < EPS_MAX_MLP_LAYERS; ++l) {
    prm.W[l] = l < L ? W_h[l] : nullptr;
    prm.b[l] = l < L ? b_h[l] : nullptr;
    if (l < L) {
      EPS_CHECK_ARG(prm.W[l] && prm.b[l], "null weight or bias pointer");
      EPS_CHECK_ARG(((uintptr_t)prm.W[l]) % 16 == 0, "weights must be 16-byte aligned");
    }
  }
  const bool prepared = (precision & EPS_MLP_REUSE_WORKSPACE) != 0;
  precision &= ~EPS_MLP_REUSE_WORKSPACE;
  if (precision == EPS_MLP_FP32)
    return linkpred_fp32_launch(h, H, pair_u, pair_v, M, prm, L, apply_sigmoid, score, stream);
  if (precision == EPS_MLP_TC_BF16)
    return linkpred_tc_launch(h, n, H, pair_u, pair_v, M, prm, L, apply_sigmoid, score, workspace,
                              workspace_bytes, prepared, stream);
  set_error("eps_linkpred_mlp: unknown precision %d", precision);
  return EPS_ERR_INVALID;
}
