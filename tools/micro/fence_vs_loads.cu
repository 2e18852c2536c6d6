// Microbenchmark: does fence.proxy.async (MEMBAR.ALL.CTA + FENCE.VIEW.ASYNC.S) wait for the thread's OUTSTANDING
// global loads?  One warp issues a global load that misses every cache, then stores to shared memory and fences;
// the clocks around the fence are compared with the same sequence without a load in flight.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void k(const float *big, long long stride, int with_load, long long *out, float *sink) {
  __shared__ float sm[64];
  const int lane = threadIdx.x;
  long long t_f = 0, t_u = 0;
  float acc = 0.f;
  for (int it = 0; it < 64; ++it) {
    float v = 0.f;
    const float *p = big + ((long long)(it * 32 + lane) * stride);
    if (with_load) v = __ldcs(p);                      // in flight, result not used before the fence
    sm[lane] = (float)it;
    const long long t0 = clock64();
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    const long long t1 = clock64();
    acc += v;                                          // first use of the load
    const long long t2 = clock64();
    t_f += t1 - t0;
    t_u += t2 - t1;
    __syncwarp();
  }
  if (lane == 0) { out[0] = t_f / 64; out[1] = t_u / 64; }
  sink[lane] = acc + sm[lane];
}

int main() {
  float *big, *sink; long long *out;
  const long long stride = 1 << 16;                    // 256 KB apart: every access a fresh DRAM page
  cudaMalloc(&big, (size_t)64 * 32 * stride * 4 + 4096); cudaMalloc(&sink, 256); cudaMalloc(&out, 16);
  cudaMemset(big, 0, (size_t)64 * 32 * stride * 4);
  for (int w = 0; w < 2; ++w) {
    k<<<1, 32>>>(big, stride, w, out, sink);
    cudaDeviceSynchronize();
    long long h[2]; cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
    printf("%s: fence %lld clk, first use of the load after the fence %lld clk\n", w ? "load in flight " : "no load        ", h[0], h[1]);
  }
  return 0;
}
