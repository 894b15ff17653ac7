// dev microbench: per-SM load throughput of Phase-A-like patterns, L2-resident data, 256 threads, 1 CTA/SM, no compute
#include <cstdio>
#include <cuda_runtime.h>
// MODE 0: kernel pattern float2 (x0 = task%40)      1: float4 pairs (xp = task%20)
//      2: float2, 32-column groups (aligned 256 B)   3: flat streaming float4      4: flat streaming float2
//      5: kernel pattern float2 but rows padded to 1664 B (13 lines)
#ifndef LOADFN
#define LOADFN __ldcg
#endif
template <int MODE> __global__ void __launch_bounds__(256, 1) pat(const float2* __restrict__ in, int n_items, float* out) {
  extern __shared__ float2 sm[];
  float acc = 0.f;
  for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
    const float2* img = in + (size_t)(item >> 1) * 41600;
    if (MODE == 0 || MODE == 2 || MODE == 5) {
      const int GW = (MODE == 2) ? 32 : 40, RS = (MODE == 5) ? 208 : 200;
      for (int k = 0; k < 4; ++k) {
        const int task = threadIdx.x + k * 256;
        const int g = task / GW, x0 = task % GW;
        if (g >= 25) break;
        const float2* p = img + g * RS + x0;
        float2 v[40];
#pragma unroll
        for (int i = 0; i < 5; ++i)
#pragma unroll
          for (int j = 0; j < 8; ++j) v[i * 8 + j] = LOADFN(p + j * 25 * RS + i * GW);
#pragma unroll
        for (int e = 0; e < 40; ++e) acc += v[e].x + v[e].y;
      }
    } else if (MODE == 1) {
      for (int k = 0; k < 2; ++k) {
        const int task = threadIdx.x + k * 256;
        const int g = task / 20, xp = task % 20;
        if (g >= 25) break;
        const float4* p = reinterpret_cast<const float4*>(img + g * 200 + 2 * xp);
        float4 v[40];
#pragma unroll
        for (int i = 0; i < 5; ++i)
#pragma unroll
          for (int j = 0; j < 8; ++j) v[i * 8 + j] = LOADFN(p + (j * 25 * 200 + i * 40) / 2);
#pragma unroll
        for (int e = 0; e < 40; ++e) acc += v[e].x + v[e].y + v[e].z + v[e].w;
      }
    } else if (MODE == 3) {
      const float4* p = reinterpret_cast<const float4*>(img);
      for (int k = 0; k < 2; ++k) {
        float4 v[40];
#pragma unroll
        for (int e = 0; e < 40; ++e) { const int idx = (k * 40 + e) * 256 + threadIdx.x; v[e] = idx < 20000 ? LOADFN(p + idx) : make_float4(0, 0, 0, 0); }
#pragma unroll
        for (int e = 0; e < 40; ++e) acc += v[e].x + v[e].y + v[e].z + v[e].w;
      }
    } else {
      for (int k = 0; k < 4; ++k) {
        float2 v[40];
#pragma unroll
        for (int e = 0; e < 40; ++e) { const int idx = (k * 40 + e) * 256 + threadIdx.x; v[e] = idx < 40000 ? LOADFN(img + idx) : make_float2(0, 0); }
#pragma unroll
        for (int e = 0; e < 40; ++e) acc += v[e].x + v[e].y;
      }
    }
    sm[threadIdx.x] = make_float2(acc, acc);
  }
  if (acc == 123.456f) out[0] = acc;
}
template <int MODE> void run(const char* name, const float2* d, float* out, int imgs, int smem = 172800) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  const int items = imgs * 2;
  cudaFuncSetAttribute(pat<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  pat<MODE><<<148, 256, smem>>>(d, items, out);
  cudaEventRecord(a);
  for (int r = 0; r < 5; ++r) pat<MODE><<<148, 256, smem>>>(d, items, out);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b); ms /= 5;
  const double c = ms * 1e-3 * 1.92e9 / ((items + 147) / 148);
  printf("%-34s smem=%6d imgs=%3d : %7.1f us  %7.0f cycles/item  %6.1f B/clk/SM  %s\n", name, smem, imgs, ms * 1e3, c, 320000.0 / c, cudaGetErrorString(cudaGetLastError()));
}
int main() {
  const int n_img = 600; float2* d; float* out;
  cudaMalloc(&d, (size_t)n_img * 41600 * 8); cudaMalloc(&out, 4); cudaMemset(d, 0, (size_t)n_img * 41600 * 8);
  for (int smem : {172800, 131072, 65536, 4096}) {
    run<0>("kernel pattern float2", d, out, 600, smem);
    run<1>("pair pattern float4", d, out, 600, smem);
    run<3>("flat streaming float4", d, out, 600, smem);
  }
  return 0;
}
