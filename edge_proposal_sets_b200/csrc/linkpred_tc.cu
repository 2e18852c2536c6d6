// K2 (tensor-core arm) — placeholder until the tcgen05 kernel lands.
#include "eps_common.cuh"
namespace eps {
size_t linkpred_tc_workspace_bytes(int H, int L) { return 256; }
int linkpred_tc_launch(const float *, int, int, const int *, const int *, long long, const MlpParams &,
                       int, int, float *, void *, size_t, cudaStream_t) {
  set_error("eps_linkpred_mlp: EPS_MLP_TC_BF16 not built");
  return EPS_ERR_UNSUPPORTED;
}
}  // namespace eps
