// Microbenchmark: raw tcgen05.ld (TMEM -> registers) throughput on one SM, for the epilogue of K2.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/tmem_ld_bw tools/micro/tmem_ld_bw.cu && /tmp/tmem_ld_bw
// Variants: load width (x16 / x32 / x64), loads in flight per warp before tcgen05.wait::ld (1 / 2 / 4),
// warps per TMEM lane quarter (1 / 2).  Reports bytes per clock per SM (all quarters together).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int X>
__device__ __forceinline__ void ld_issue(uint32_t taddr, uint32_t *r);
template <>
__device__ __forceinline__ void ld_issue<16>(uint32_t taddr, uint32_t *r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                 "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr) : "memory");
}
template <>
__device__ __forceinline__ void ld_issue<32>(uint32_t taddr, uint32_t *r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                 "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
                 "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
                 "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
               : "r"(taddr) : "memory");
}
__device__ __forceinline__ void ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// each warp of quarter q reads its share of `cols` columns `iters` times; INFL loads of X columns in flight
template <int X, int INFL, int WPQ>
__global__ void __launch_bounds__(128 * WPQ) bw_kernel(int iters, int cols, long long *clocks, uint32_t *sink) {
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, q = warp & 3, part = warp >> 2;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = tmem_slot + ((uint32_t)(q * 32) << 16);
  const int my_cols = cols / WPQ, c0 = part * my_cols;
  uint32_t acc = 0;
  uint32_t buf[INFL][X];
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    for (int c = 0; c < my_cols; c += X * INFL) {
#pragma unroll
      for (int f = 0; f < INFL; ++f) ld_issue<X>(base + (uint32_t)(c0 + c + f * X), buf[f]);
      ld_wait();
#pragma unroll
      for (int f = 0; f < INFL; ++f)
#pragma unroll
        for (int j = 0; j < X; ++j) acc ^= buf[f][j];
    }
  }
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) clocks[blockIdx.x] = t1 - t0;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_slot), "r"(512) : "memory");
}

template <int X, int INFL, int WPQ>
static void run(const char *name) {
  long long *clk; uint32_t *sink;
  cudaMalloc(&clk, 8 * 148); cudaMalloc(&sink, 4 * 148 * 256);
  const int iters = 200, cols = 256;
  bw_kernel<X, INFL, WPQ><<<148, 128 * WPQ>>>(2, cols, clk, sink);
  bw_kernel<X, INFL, WPQ><<<148, 128 * WPQ>>>(iters, cols, clk, sink);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
  const double bytes = (double)iters * cols * 128 * 4;
  printf("%-44s %s  clocks %lld  -> %.1f B/clk/SM  (%.0f clk per 128x256 fp32 accumulator)\n", name, cudaGetErrorString(e), h[0],
         bytes / h[0], (double)h[0] / iters);
  cudaFree(clk); cudaFree(sink);
}

int main() {
  run<16, 1, 1>("x16, 1 in flight, 1 warp/quarter");
  run<16, 2, 1>("x16, 2 in flight, 1 warp/quarter");
  run<16, 4, 1>("x16, 4 in flight, 1 warp/quarter");
  run<32, 1, 1>("x32, 1 in flight, 1 warp/quarter");
  run<32, 2, 1>("x32, 2 in flight, 1 warp/quarter");
  run<32, 4, 1>("x32, 4 in flight, 1 warp/quarter");
  run<16, 2, 2>("x16, 2 in flight, 2 warps/quarter");
  run<32, 1, 2>("x32, 1 in flight, 2 warps/quarter");
  run<32, 2, 2>("x32, 2 in flight, 2 warps/quarter");
  return 0;
}
