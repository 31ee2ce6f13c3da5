// Same-address atomicAdd throughput (one atomic per warp per trip, result consumed) against the same
// loop without contention (one counter per warp).  nvcc -arch=sm_100a -O3 -o atomic_rate atomic_rate.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(int *counters, int stride_ints, int trips, int *sink, int work) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int *c = counters + (size_t)warp * stride_ints;
  int acc = 0;
  float f = threadIdx.x;
  for (int t = 0; t < trips; t++) {
    for (int w = 0; w < work; w++) f = f * 1.0001f + 0.5f;  // stand-in for the shading arithmetic
    int pos = 0;
    if ((threadIdx.x & 31) == 0) pos = atomicAdd(c, 26);
    pos = __shfl_sync(0xffffffffu, pos, 0);
    acc += pos;
  }
  if (acc == 0x7fffffff || f == 12345.f) sink[0] = acc;
}
int main() {
  int *d, *sink;
  const int blocks = 148 * 8, threads = 128, warps = blocks * threads / 32;
  cudaMalloc(&d, (size_t)warps * 128);
  cudaMalloc(&sink, 4);
  cudaMemset(d, 0, (size_t)warps * 128);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int work : {0, 200, 600}) {
    for (int stride : {0, 32}) {
      const int trips = 256;
      k<<<blocks, threads>>>(d, stride, trips, sink, work);
      cudaEventRecord(e0);
      k<<<blocks, threads>>>(d, stride, trips, sink, work);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      const double n = (double)warps * trips;
      printf("work %3d FMAs/trip, %s: %.3f ms for %.0f warp atomics -> %.2f G atomics/s (%.2f ns each)\n", work,
             stride ? "one counter per warp" : "ONE counter      ", ms, n, n / ms / 1e6, ms * 1e6 / n);
    }
  }
  return 0;
}
