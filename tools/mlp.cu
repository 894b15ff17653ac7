// dev microbench: per-SM load throughput at low occupancy vs loads in flight per thread (LDG.64/128) and TMA bulk copies
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int VEC, int U> __global__ void __launch_bounds__(512, 1) rd(const float* __restrict__ p, size_t n_vec, int reps, float* out) {
  float acc = 0.f;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (int r = 0; r < reps; ++r) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + (U - 1) * stride < n_vec; i += U * stride) {
      if (VEC == 4) { float4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) v[u] = __ldcg(reinterpret_cast<const float4*>(p) + i + u * stride);
#pragma unroll
        for (int u = 0; u < U; ++u) acc += v[u].x + v[u].y + v[u].z + v[u].w;
      } else { float2 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) v[u] = __ldcg(reinterpret_cast<const float2*>(p) + i + u * stride);
#pragma unroll
        for (int u = 0; u < U; ++u) acc += v[u].x + v[u].y;
      }
    }
  }
  if (acc == 123.456f) out[0] = acc;
}
// TMA: one thread issues bulk copies of CHUNK bytes into a ring of NSLOT slots; everyone just waits (no consumption)
template <int CHUNK, int NSLOT> __global__ void __launch_bounds__(128, 1) tma_rd(const char* p, size_t bytes_per_cta, int reps, float* out) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) unsigned long long bar[NSLOT];
  const unsigned sbase = (unsigned)__cvta_generic_to_shared(smem);
  if (threadIdx.x == 0) {
    for (int s = 0; s < NSLOT; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((unsigned)__cvta_generic_to_shared(&bar[s])));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  __syncthreads();
  const char* src = p + (size_t)blockIdx.x * bytes_per_cta;
  const int n_chunks = (int)(bytes_per_cta / CHUNK) * reps;
  if (threadIdx.x == 0) {
    int issued = 0, done = 0; unsigned phase[NSLOT]; for (int s = 0; s < NSLOT; ++s) phase[s] = 0;
    while (done < n_chunks) {
      while (issued < n_chunks && issued - done < NSLOT) {
        const int s = issued % NSLOT;
        const unsigned b = (unsigned)__cvta_generic_to_shared(&bar[s]);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(CHUNK));
        asm volatile("cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(sbase + s * CHUNK), "l"(src + (size_t)(issued % (bytes_per_cta / CHUNK)) * CHUNK), "r"(CHUNK), "r"(b) : "memory");
        ++issued;
      }
      const int s = done % NSLOT;
      const unsigned b = (unsigned)__cvta_generic_to_shared(&bar[s]);
      unsigned ok = 0;
      while (!ok) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(b), "r"(phase[s]) : "memory");
      phase[s] ^= 1; ++done;
    }
  }
  __syncthreads();
  if (smem[threadIdx.x] == 77 && out) out[1] = 1.f;
}
template <int VEC, int U> void run(const char* name, const float* d, size_t bytes, int reps, int threads, float* out) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  size_t n_vec = bytes / (4 * VEC);
  rd<VEC, U><<<148, threads>>>(d, n_vec, 1, out);
  cudaEventRecord(a); rd<VEC, U><<<148, threads>>>(d, n_vec, reps, out); cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  double gbs = bytes * (double)reps / ms / 1e6;
  printf("%-10s vec%d U=%2d thr=%4d inflight/SM=%6.1f KB : %8.1f GB/s  (%5.1f B/clk/SM @1.92GHz)\n", name, VEC * 4, U, threads, threads * U * VEC * 4 / 1024.0, gbs, gbs / 148 / 1.92);
}
template <int CHUNK, int NSLOT> void run_tma(const char* name, const char* d, size_t bytes, int reps, float* out) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  size_t per = bytes / 148 / CHUNK * CHUNK;
  cudaFuncSetAttribute(tma_rd<CHUNK, NSLOT>, cudaFuncAttributeMaxDynamicSharedMemorySize, CHUNK * NSLOT);
  tma_rd<CHUNK, NSLOT><<<148, 128, CHUNK * NSLOT>>>(d, per, 1, out);
  cudaEventRecord(a); tma_rd<CHUNK, NSLOT><<<148, 128, CHUNK * NSLOT>>>(d, per, reps, out); cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  double gbs = per * 148.0 * reps / ms / 1e6;
  printf("%-10s TMA chunk=%5d slots=%2d inflight/SM=%6.1f KB : %8.1f GB/s  (%5.1f B/clk/SM) %s\n", name, CHUNK, NSLOT, CHUNK * NSLOT / 1024.0, gbs, gbs / 148 / 1.92, cudaGetErrorString(cudaGetLastError()));
}
int main() {
  float *d, *out; size_t big = (size_t)2 << 30;
  cudaMalloc(&d, big); cudaMalloc(&out, 8); cudaMemset(d, 0, big);
  const size_t l2 = 32u << 20;
  run<2, 1>("L2", d, l2, 40, 512, out); run<2, 4>("L2", d, l2, 40, 512, out); run<2, 8>("L2", d, l2, 40, 512, out); run<2, 16>("L2", d, l2, 40, 512, out); run<2, 32>("L2", d, l2, 40, 512, out);
  run<4, 1>("L2", d, l2, 40, 512, out); run<4, 4>("L2", d, l2, 40, 512, out); run<4, 8>("L2", d, l2, 40, 512, out); run<4, 16>("L2", d, l2, 40, 512, out);
  run<2, 8>("L2", d, l2, 40, 256, out); run<2, 16>("L2", d, l2, 40, 256, out); run<2, 32>("L2", d, l2, 40, 256, out);
  run<4, 8>("L2", d, l2, 40, 256, out); run<4, 16>("L2", d, l2, 40, 256, out);
  run<2, 4>("HBM", d, big, 2, 512, out); run<2, 16>("HBM", d, big, 2, 512, out); run<2, 32>("HBM", d, big, 2, 512, out);
  run<4, 4>("HBM", d, big, 2, 512, out); run<4, 16>("HBM", d, big, 2, 512, out);
  run<4, 16>("HBM", d, big, 2, 256, out);
  run_tma<1600, 8>("L2", (const char*)d, l2, 40, out); run_tma<1600, 32>("L2", (const char*)d, l2, 40, out);
  run_tma<12800, 4>("L2", (const char*)d, l2, 40, out); run_tma<12800, 8>("L2", (const char*)d, l2, 40, out);
  run_tma<1600, 32>("HBM", (const char*)d, big, 2, out); run_tma<12800, 4>("HBM", (const char*)d, big, 2, out); run_tma<12800, 8>("HBM", (const char*)d, big, 2, out);
  run_tma<12800, 16>("HBM", (const char*)d, big, 2, out);
  return 0;
}
