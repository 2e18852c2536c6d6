// Microbenchmark: which primitive gathers 64-byte / 128-byte pieces of random fp16 embedding rows fastest on one SM?
//   LDG.128.nc -> registers, LDGSTS (cp.async .cg / .ca) -> shared memory; W warps per SM, K 16-byte units in flight per thread.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/micro/bin/gather_bw tools/micro/gather_bw.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// MODE 0: ld.global.nc.v4 -> regs; 1: cp.async.cg 16; 2: cp.async.ca 16.  PIECE = bytes per row piece (64 / 128 / 512).
template <int MODE, int PIECE, int K>
__global__ void __launch_bounds__(1024) k(const char *table, int row_bytes, const int *rows, int mask, int iters, long long *clocks,
                                          unsigned *sink) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  constexpr int LPR = PIECE / 16;            // lanes per row piece
  constexpr int RPI = 32 / LPR;              // rows per warp instruction
  const int rsub = lane / LPR, unit = lane % LPR;
  int base = (blockIdx.x * nw + warp) * 8192;
  uint8_t *dst = smem + (size_t)warp * K * 512 * 2;
  unsigned acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {
      uint4 v[K];
#pragma unroll
      for (int j = 0; j < K; ++j) {
        const int r = __ldg(rows + ((base + j * RPI + rsub) & mask));
        const uint4 *p = reinterpret_cast<const uint4 *>(table + (size_t)r * row_bytes + ((it & 3) * PIECE) % row_bytes + unit * 16);
        asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v[j].x), "=r"(v[j].y), "=r"(v[j].z), "=r"(v[j].w) : "l"(p));
      }
#pragma unroll
      for (int j = 0; j < K; ++j) acc ^= v[j].x ^ v[j].y ^ v[j].z ^ v[j].w;
    } else {
      uint8_t *d = dst + (it & 1) * K * 512;
#pragma unroll
      for (int j = 0; j < K; ++j) {
        const int r = __ldg(rows + ((base + j * RPI + rsub) & mask));
        const char *p = table + (size_t)r * row_bytes + ((it & 3) * PIECE) % row_bytes + unit * 16;
        if (MODE == 1) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(smem_u32(d + j * 512 + lane * 16)), "l"(p) : "memory");
        else asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" :: "r"(smem_u32(d + j * 512 + lane * 16)), "l"(p) : "memory");
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
      asm volatile("cp.async.wait_group 1;" ::: "memory");      // the previous iteration's group
    }
    base += K * RPI;
  }
  if (MODE != 0) asm volatile("cp.async.wait_group 0;" ::: "memory");
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) clocks[blockIdx.x] = t1 - t0;
  if (MODE != 0) acc = dst[lane];
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int MODE, int PIECE, int K>
static void run(const char *name, const char *table, int row_bytes, const int *rows, int mask, int warps, long long *clk, unsigned *sink) {
  const int iters = 300;
  const size_t smem = MODE ? (size_t)warps * K * 512 * 2 : 0;
  cudaFuncSetAttribute(k<MODE, PIECE, K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  k<MODE, PIECE, K><<<148, warps * 32, smem>>>(table, row_bytes, rows, mask, 4, clk, sink);
  k<MODE, PIECE, K><<<148, warps * 32, smem>>>(table, row_bytes, rows, mask, iters, clk, sink);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
  const double bytes = (double)iters * warps * K * 512;
  printf("%-12s piece=%3d B  K=%2d  warps=%2d  %s  %.1f B/clk/SM  (%.0f clk per warp instruction per SM)\n", name, PIECE, K, warps,
         cudaGetErrorString(e), bytes / avg, avg / ((double)iters * warps * K));
}

int main(int argc, char **argv) {
  const int H = 256, NB = 576289, row_bytes = H * 2;
  char *table; cudaMalloc(&table, (size_t)NB * row_bytes); cudaMemset(table, 1, (size_t)NB * row_bytes);
  const int NR = 1 << 23;
  std::vector<int> hr(NR);
  long long *clk; cudaMalloc(&clk, 8 * 148);
  unsigned *sink; cudaMalloc(&sink, 4 * 148 * 1024);
  int *rows; cudaMalloc(&rows, (size_t)NR * 4);
  for (int pass = 0; pass < 2; ++pass) {
    srand(1);
    const int n = pass == 0 ? 65536 : NB;     // 32 MB (L2 resident) / 295 MB
    for (int i = 0; i < NR; ++i) hr[i] = (int)(((long long)rand() * 32768 + rand()) % n);
    cudaMemcpy(rows, hr.data(), (size_t)NR * 4, cudaMemcpyHostToDevice);
    printf("---- random rows of a %d-row table (%d MB)\n", n, (int)((size_t)n * row_bytes >> 20));
    for (int warps : {4, 8, 16}) {
      run<0, 64, 4>("LDG.nc", table, row_bytes, rows, NR - 1, warps, clk, sink);
      run<0, 64, 8>("LDG.nc", table, row_bytes, rows, NR - 1, warps, clk, sink);
      run<0, 128, 8>("LDG.nc", table, row_bytes, rows, NR - 1, warps, clk, sink);
      run<0, 512, 8>("LDG.nc", table, row_bytes, rows, NR - 1, warps, clk, sink);
    }
    for (int warps : {1, 4, 8}) {
      run<1, 64, 8>("LDGSTS.cg", table, row_bytes, rows, NR - 1, warps, clk, sink);
      run<1, 128, 8>("LDGSTS.cg", table, row_bytes, rows, NR - 1, warps, clk, sink);
      run<1, 512, 8>("LDGSTS.cg", table, row_bytes, rows, NR - 1, warps, clk, sink);
      run<2, 64, 8>("LDGSTS.ca", table, row_bytes, rows, NR - 1, warps, clk, sink);
      run<2, 128, 8>("LDGSTS.ca", table, row_bytes, rows, NR - 1, warps, clk, sink);
    }
    run<1, 64, 16>("LDGSTS.cg", table, row_bytes, rows, NR - 1, 1, clk, sink);
    run<1, 64, 16>("LDGSTS.cg", table, row_bytes, rows, NR - 1, 2, clk, sink);
    run<1, 128, 16>("LDGSTS.cg", table, row_bytes, rows, NR - 1, 2, clk, sink);
  }
  return 0;
}
