// K5 — multi-GPU proposal-set merge: one ncclAllGather of the per-GPU top-k rows + a final
// K4 select on every rank.
//
// The reference is single-GPU (no torch.distributed / NCCL call anywhere in /root/reference);
// this is the only exchange step the sharded filter needs: candidates are owned by
// contiguous ranges of v = all_edges[:,1], every GPU scores its own range and keeps a local
// top-k; since rank order == owner order == candidate-index order, the position of a row in the
// gathered [world, k_local] array preserves the global tie order inside equal scores, so the same
// "score desc, position asc" select reproduces the single-GPU proposal list bit for bit.
//
// NCCL is resolved with dlopen at first use (the copy torch already loaded, libnccl.so.2), so the
// library has no link-time NCCL dependency and loads on hosts without it.
#include <dlfcn.h>
#include <string.h>

#include "eps_common.cuh"

namespace eps {

typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;  // ncclSuccess == 0
constexpr int kNcclFloat = 7;  // ncclFloat32

struct NcclApi {
  ncclResult_t (*GetUniqueId)(ncclUniqueId *);
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int);
  ncclResult_t (*AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t);
  ncclResult_t (*CommDestroy)(ncclComm_t);
  const char *(*GetErrorString)(ncclResult_t);
  bool ok = false;
};

static NcclApi &nccl() {
  static NcclApi api;
  static bool tried = false;
  if (tried) return api;
  tried = true;
  void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) return api;
  api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(h, "ncclGetUniqueId");
  api.CommInitRank = (decltype(api.CommInitRank))dlsym(h, "ncclCommInitRank");
  api.AllGather = (decltype(api.AllGather))dlsym(h, "ncclAllGather");
  api.CommDestroy = (decltype(api.CommDestroy))dlsym(h, "ncclCommDestroy");
  api.GetErrorString = (decltype(api.GetErrorString))dlsym(h, "ncclGetErrorString");
  api.ok = api.GetUniqueId && api.CommInitRank && api.AllGather && api.CommDestroy;
  return api;
}

#define EPS_NCCL(call)                                                                   \
  do {                                                                                   \
    ncclResult_t r__ = (call);                                                           \
    if (r__ != 0) {                                                                      \
      eps::set_error("%s: %s -> %s", __func__, #call,                                    \
                     nccl().GetErrorString ? nccl().GetErrorString(r__) : "nccl error"); \
      return EPS_ERR_NCCL;                                                               \
    }                                                                                    \
  } while (0)

struct Comm {
  ncclComm_t comm;
  int world, rank;
};

__global__ void extract_score_kernel(const float *__restrict__ rows, long long m, float *__restrict__ score) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < m) score[i] = rows[i * 3 + 2];
}

__global__ void gather_rows_kernel(const float *__restrict__ rows, const uint32_t *__restrict__ idx,
                                   long long k, float *__restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < k) {
    const size_t p = idx[i];
    out[i * 3 + 0] = rows[p * 3 + 0];
    out[i * 3 + 1] = rows[p * 3 + 1];
    out[i * 3 + 2] = rows[p * 3 + 2];
  }
}

static inline size_t a256(size_t x) { return (x + 255) & ~(size_t)255; }

}  // namespace eps

extern "C" int eps_comm_unique_id(void *out128_h) {
  using namespace eps;
  EPS_CHECK_ARG(out128_h != nullptr, "null output");
  if (!nccl().ok) { set_error("eps_comm_unique_id: libnccl.so.2 not found"); return EPS_ERR_NCCL; }
  ncclUniqueId id;
  EPS_NCCL(nccl().GetUniqueId(&id));
  memcpy(out128_h, &id, 128);
  return EPS_OK;
}

extern "C" int eps_comm_init(const void *id128_h, int world, int rank, void **comm_out) {
  using namespace eps;
  EPS_CHECK_ARG(id128_h && comm_out, "null pointer");
  EPS_CHECK_ARG(world >= 1 && rank >= 0 && rank < world, "bad world/rank");
  if (!nccl().ok) { set_error("eps_comm_init: libnccl.so.2 not found"); return EPS_ERR_NCCL; }
  ncclUniqueId id;
  memcpy(&id, id128_h, 128);
  Comm *c = new Comm{nullptr, world, rank};
  ncclResult_t r = nccl().CommInitRank(&c->comm, world, id, rank);
  if (r != 0) {
    set_error("eps_comm_init: ncclCommInitRank -> %s", nccl().GetErrorString ? nccl().GetErrorString(r) : "?");
    delete c;
    return EPS_ERR_NCCL;
  }
  *comm_out = c;
  return EPS_OK;
}

extern "C" int eps_comm_destroy(void *comm) {
  using namespace eps;
  if (!comm) return EPS_OK;
  Comm *c = (Comm *)comm;
  if (nccl().ok && c->comm) nccl().CommDestroy(c->comm);
  delete c;
  return EPS_OK;
}

extern "C" size_t eps_topk_merge_workspace_bytes(int world, int64_t k_local, int64_t k) {
  using namespace eps;
  const int64_t m = (int64_t)world * k_local;
  if (m <= 0 || k <= 0) return 256;
  return a256((size_t)m * 12) + a256((size_t)m * 4) + a256((size_t)k * 4) + eps_topk_workspace_bytes(m, k);
}

// local_k3: this rank's [k_local,3] fp32 rows (u, v, score), sorted, padded with score = -inf rows
// if the rank owns fewer than k_local candidates.  out_k3: [k,3] global proposal list (all ranks).
extern "C" int eps_topk_merge_allgather(void *comm, const float *local_k3, int64_t k_local, int64_t k,
                                        float *out_k3, void *workspace, size_t workspace_bytes,
                                        void *stream_) {
  using namespace eps;
  cudaStream_t stream = (cudaStream_t)stream_;
  EPS_CHECK_ARG(comm && local_k3 && out_k3, "null pointer");
  Comm *c = (Comm *)comm;
  const int64_t m = (int64_t)c->world * k_local;
  EPS_CHECK_ARG(k_local >= 1 && k >= 1 && k <= m, "need 1 <= k <= world*k_local");
  if (!workspace || workspace_bytes < eps_topk_merge_workspace_bytes(c->world, k_local, k)) {
    set_error("eps_topk_merge_allgather: workspace too small");
    return EPS_ERR_WORKSPACE;
  }
  char *ws = (char *)workspace;
  float *gathered = (float *)ws; ws += a256((size_t)m * 12);
  float *score = (float *)ws; ws += a256((size_t)m * 4);
  uint32_t *idx = (uint32_t *)ws; ws += a256((size_t)k * 4);
  EPS_NCCL(nccl().AllGather(local_k3, gathered, (size_t)k_local * 3, kNcclFloat, c->comm, stream));
  extract_score_kernel<<<(unsigned)((m + 255) / 256), 256, 0, stream>>>(gathered, m, score);
  EPS_LAUNCH_CHECK();
  int st = eps_topk_f32(score, m, k, idx, nullptr, ws, eps_topk_workspace_bytes(m, k), stream);
  if (st != EPS_OK) return st;
  gather_rows_kernel<<<(unsigned)((k + 255) / 256), 256, 0, stream>>>(gathered, idx, k, out_k3);
  EPS_LAUNCH_CHECK();
  return EPS_OK;
}
