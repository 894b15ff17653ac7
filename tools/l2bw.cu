// dev microbench: L2->SM read bandwidth (buffer resident in L2) and HBM read bandwidth, various access widths
#include <cstdio>
#include <cuda_runtime.h>
template <int VEC> __global__ void rd(const float* __restrict__ p, size_t n_vec, int reps, float* out) {
  float acc = 0.f;
  for (int r = 0; r < reps; ++r) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += (size_t)gridDim.x * blockDim.x) {
      if (VEC == 4) { float4 v = __ldcg(reinterpret_cast<const float4*>(p) + i); acc += v.x + v.y + v.z + v.w; }
      else if (VEC == 2) { float2 v = __ldcg(reinterpret_cast<const float2*>(p) + i); acc += v.x + v.y; }
      else { acc += __ldcg(p + i); }
    }
  }
  if (acc == 123.456f) out[0] = acc;
}
template <int VEC> void run(const char* name, const float* d, size_t bytes, int reps, int blocks, int threads, float* out) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  size_t n_vec = bytes / (4 * VEC);
  rd<VEC><<<blocks, threads>>>(d, n_vec, 1, out);
  cudaEventRecord(a); rd<VEC><<<blocks, threads>>>(d, n_vec, reps, out); cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  printf("%-28s %7.1f MB x%3d  blocks %4d x %4d : %8.1f GB/s\n", name, bytes / 1e6, reps, blocks, threads, bytes * (double)reps / ms / 1e6);
}
int main() {
  float *d, *out; size_t big = (size_t)2 << 30;
  cudaMalloc(&d, big); cudaMalloc(&out, 4); cudaMemset(d, 0, big);
  for (int threads : {256, 512, 1024}) {
    run<4>("L2 32MB float4", d, 32u << 20, 50, 148 * (2048 / threads), threads, out);
    run<2>("L2 32MB float2", d, 32u << 20, 50, 148 * (2048 / threads), threads, out);
  }
  run<4>("L2 32MB float4 1cta/SM 512", d, 32u << 20, 50, 148, 512, out);
  run<2>("L2 32MB float2 1cta/SM 512", d, 32u << 20, 50, 148, 512, out);
  run<2>("L2 32MB float2 1cta/SM 256", d, 32u << 20, 50, 148, 256, out);
  run<4>("L2 64MB float4", d, 64u << 20, 30, 148 * 8, 256, out);
  run<4>("L2 96MB float4", d, 96u << 20, 20, 148 * 8, 256, out);
  run<4>("HBM 2GB float4", d, big, 2, 148 * 8, 256, out);
  run<2>("HBM 2GB float2", d, big, 2, 148 * 8, 256, out);
  return 0;
}
