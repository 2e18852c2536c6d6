// Microbenchmark: issue / service rate of cp.async.bulk.tensor.2d ... tile::gather4 with small boxes (the K2 gather:
// 4 rows x 64 B or 4 rows x 128 B of an fp16 embedding table), as a function of the number of issuing warps per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/micro/bin/tma_gather4_bw tools/micro/tma_gather4_bw.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// every warp: `iters` rounds of (expect_tx, 32 lanes x 1 gather4 [ALL lanes] or LPW lanes, wait)
template <int BOXB /* bytes per row piece */>
__global__ void __launch_bounds__(512) k(const __grid_constant__ CUtensorMap tmap, const int *rows, int nrows_mask, int iters,
                                         int lanes_per_warp, long long *clocks) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar[16];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  if (lane == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bar[warp])));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  uint8_t *dst = smem + (size_t)warp * 32 * 4 * BOXB;       // one "stage" per warp
  const uint32_t b = smem_u32(&bar[warp]);
  uint32_t ph = 0;
  int base = (blockIdx.x * nw + warp) * 4096 + lane * 4;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    const int r0 = rows[(base + 0) & nrows_mask], r1 = rows[(base + 1) & nrows_mask], r2 = rows[(base + 2) & nrows_mask],
              r3 = rows[(base + 3) & nrows_mask];
    base += 128;
    if (lane == 0)
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(b), "r"(lanes_per_warp * 4 * BOXB) : "memory");
    __syncwarp();
    if (lane < lanes_per_warp)
      asm volatile(
          "cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
          :: "r"(smem_u32(dst) + lane * 4 * BOXB), "l"(reinterpret_cast<uint64_t>(&tmap)), "r"((it & 3) * (BOXB / 2)), "r"(r0), "r"(r1),
             "r"(r2), "r"(r3), "r"(b) : "memory");
    asm volatile(
        "{\n\t.reg .pred P1;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t@P1 bra D;\n\tbra W;\n\tD:\n\t}\n"
        :: "r"(b), "r"(ph) : "memory");
    ph ^= 1;
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) clocks[blockIdx.x] = t1 - t0;
}

typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                             const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int BOXB>
static void run(EncodeFn enc, void *table, int n, int H, const int *rows, int mask, int warps, int lpw, long long *clk) {
  CUtensorMap tm;
  const cuuint64_t gdim[2] = {(cuuint64_t)H, (cuuint64_t)n}, gstr[1] = {(cuuint64_t)H * 2};
  const cuuint32_t box[2] = {BOXB / 2, 1}, es[2] = {1, 1};
  CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, table, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   BOXB == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return; }
  const int iters = 400;
  const size_t smem = (size_t)warps * 32 * 4 * BOXB;
  cudaFuncSetAttribute(k<BOXB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  k<BOXB><<<148, warps * 32, smem>>>(tm, rows, mask, 4, lpw, clk);
  k<BOXB><<<148, warps * 32, smem>>>(tm, rows, mask, iters, lpw, clk);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
  const double g4 = (double)iters * warps * lpw;
  printf("n=%7d box=%3d B  warps=%2d lanes/warp=%2d  %s  %.1f clk per gather4 per SM  -> %.1f B/clk/SM\n", n, BOXB, warps, lpw,
         cudaGetErrorString(e), avg / g4, g4 * 4 * BOXB / avg);
}

int main() {
  void *fp = nullptr; cudaDriverEntryPointQueryResult q;
  cudaFree(0);
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q) != cudaSuccess || !fp) { printf("no encode\n"); return 1; }
  EncodeFn enc = (EncodeFn)fp;
  const int H = 256, NB = 576289;
  void *table; cudaMalloc(&table, (size_t)NB * H * 2); cudaMemset(table, 0, (size_t)NB * H * 2);
  const int NR = 1 << 22;
  std::vector<int> hr(NR);
  long long *clk; cudaMalloc(&clk, 8 * 148);
  int *rows; cudaMalloc(&rows, NR * 4);
  for (int n : {4096, NB}) {
    srand(1); for (int i = 0; i < NR; ++i) hr[i] = (int)(((long long)rand() * 32768 + rand()) % n);
    cudaMemcpy(rows, hr.data(), NR * 4, cudaMemcpyHostToDevice);
    for (int warps : {1, 2, 4, 8, 16}) run<64>(enc, table, n, H, rows, NR - 1, warps, 32, clk);
    for (int warps : {1, 4, 8}) run<128>(enc, table, n, H, rows, NR - 1, warps, 32, clk);
    run<64>(enc, table, n, H, rows, NR - 1, 8, 4, clk);
    run<64>(enc, table, n, H, rows, NR - 1, 16, 2, clk);
    run<64>(enc, table, n, H, rows, NR - 1, 16, 1, clk);
  }
  return 0;
}
