// Issue rate of FFMA vs FFMA2 (fma.rn.f32x2) on sm_100a: 8 independent accumulator chains per thread,
// 1024 threads per block, one block per SM.  Prints warp instructions per clock per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_rate ffma2_rate.cu && ./ffma2_rate
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(1024) rate(float *out, float a, float b, int iters, long long *clk) {
  float x[8], y[8];
  for (int k = 0; k < 8; k++) {
    x[k] = threadIdx.x * 0.001f + k;
    y[k] = threadIdx.x * 0.002f - k;
  }
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int k = 0; k < 8; k++) {
      if (MODE == 0) {
        x[k] = fmaf(x[k], a, b);
        y[k] = fmaf(y[k], a, b);
      } else {
        unsigned long long v, aa, bb;
        asm("mov.b64 %0, {%1, %2};" : "=l"(v) : "f"(x[k]), "f"(y[k]));
        asm("mov.b64 %0, {%1, %1};" : "=l"(aa) : "f"(a));
        asm("mov.b64 %0, {%1, %1};" : "=l"(bb) : "f"(b));
        asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(v) : "l"(v), "l"(aa), "l"(bb));
        asm("mov.b64 {%0, %1}, %2;" : "=f"(x[k]), "=f"(y[k]) : "l"(v));
      }
    }
  }
  long long t1 = clock64();
  float s = 0;
  for (int k = 0; k < 8; k++) s += x[k] + y[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

int main() {
  float *out;
  long long *clk, h[148];
  cudaMalloc(&out, 148 * 1024 * 4);
  cudaMalloc(&clk, 148 * 8);
  const int iters = 20000;
  for (int mode = 0; mode < 2; mode++) {
    for (int rep = 0; rep < 2; rep++) {
      if (mode == 0) rate<0><<<148, 1024>>>(out, 1.0001f, 0.5f, iters, clk);
      else rate<1><<<148, 1024>>>(out, 1.0001f, 0.5f, iters, clk);
      cudaDeviceSynchronize();
    }
    cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
    const double fma_per_thread = (double)iters * 16;
    const double warp_instr = fma_per_thread / (mode == 0 ? 1 : 2) * 32;  // per SM (32 warps)
    printf("%s: %.0f clocks, %.2f warp-instr/clk/SM, %.1f FMA lanes/clk/SM\n", mode == 0 ? "FFMA " : "FFMA2", (double)h[0],
           warp_instr / h[0], fma_per_thread * 1024 / h[0]);
  }
  return 0;
}
