// dev microbench: fp32 FMA issue rate, scalar FFMA vs packed fma.rn.f32x2 (sm_100a), 8 independent chains per thread
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE> __global__ void __launch_bounds__(256) k(float* out, int iters, float a, float b) {
  float x[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = threadIdx.x * 0.001f + i;
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {
#pragma unroll
      for (int i = 0; i < 16; ++i) x[i] = fmaf(x[i], a, b);
    } else {
#pragma unroll
      for (int i = 0; i < 16; i += 2) {
        unsigned long long v, va, vb;
        asm("mov.b64 %0, {%1, %2};" : "=l"(v) : "f"(x[i]), "f"(x[i + 1]));
        asm("mov.b64 %0, {%1, %1};" : "=l"(va) : "f"(a));
        asm("mov.b64 %0, {%1, %1};" : "=l"(vb) : "f"(b));
        asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(v) : "l"(v), "l"(va), "l"(vb));
        asm("mov.b64 {%0, %1}, %2;" : "=f"(x[i]), "=f"(x[i + 1]) : "l"(v));
      }
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE> void run(const char* name, float* out) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  const int iters = 4096, grid = 148 * 8;
  k<MODE><<<grid, 256>>>(out, 16, 1.0001f, 0.5f);
  cudaEventRecord(a); k<MODE><<<grid, 256>>>(out, iters, 1.0001f, 0.5f); cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  const double fmas = (double)grid * 256 * 16.0 * iters;
  printf("%-10s %8.3f ms  %7.2f TFMA/s  (%5.1f FMA/clk/SM @1.92 GHz)  %s\n", name, ms, fmas / ms / 1e9, fmas / (ms * 1e-3) / 148 / 1.92e9, cudaGetErrorString(cudaGetLastError()));
}
int main() { float* out; cudaMalloc(&out, 148 * 8 * 256 * 4); run<0>("FFMA", out); run<1>("FFMA2", out); run<0>("FFMA", out); run<1>("FFMA2", out); return 0; }
