// dev microbench: TMA (cp.async.bulk) row copies into a shared-memory ring with a producer thread and
// consumer warps (full/empty mbarriers), 1 CTA/SM, 172 KB of smem reserved like the real kernel.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* b, int n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(n)); }
__device__ __forceinline__ void mbar_expect(unsigned long long* b, unsigned bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(unsigned long long* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ void mbar_wait(unsigned long long* b, unsigned parity) {
  asm volatile("{ .reg .pred p; WAIT_%=: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1; @p bra DONE_%=; bra WAIT_%=; DONE_%=: }" ::"r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// slot = ROWS rows of 1600 B gathered from rows {g + 25 j} (the Phase A group), NSLOT slots
template <int ROWS, int NSLOT> __global__ void __launch_bounds__(288, 1) ring(const char* img_base, int n_items, float* out) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) unsigned long long full[NSLOT], empty[NSLOT];
  const int tid = threadIdx.x;
  if (tid == 0) { for (int s = 0; s < NSLOT; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 256); } asm volatile("fence.mbarrier_init.release.cluster;"); }
  __syncthreads();
  constexpr int GROUPS = 200 / ROWS;          // slots per image
  float acc = 0.f;
  if (tid >= 256) {                           // producer warp
    if (tid == 256) {
      int it = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const char* img = img_base + (size_t)(item >> 1) * 320000;
        for (int g = 0; g < GROUPS; ++g, ++it) {
          const int s = it % NSLOT; const unsigned ph = (it / NSLOT) & 1;
          if (it >= NSLOT) mbar_wait(&empty[s], ph ^ 1);
          mbar_expect(&full[s], ROWS * 1600);
          for (int j = 0; j < ROWS; ++j) tma_load(smem + (size_t)s * ROWS * 1600 + j * 1600, img + (size_t)(g + GROUPS * j) * 1600, 1600, &full[s]);
        }
      }
    }
  } else {                                    // consumers: wait, touch a little, release
    int it = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x)
      for (int g = 0; g < GROUPS; ++g, ++it) {
        const int s = it % NSLOT; const unsigned ph = (it / NSLOT) & 1;
        mbar_wait(&full[s], ph);
        acc += reinterpret_cast<const float*>(smem + (size_t)s * ROWS * 1600)[tid];
        mbar_arrive(&empty[s]);
      }
  }
  if (acc == 123.456f) out[0] = acc;
}
template <int ROWS, int NSLOT> void run(const char* d, float* out, int imgs) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  const int items = imgs * 2, smem = ROWS * NSLOT * 1600 + 120000;     // + ballast so only 1 CTA/SM fits, like the real kernel
  cudaFuncSetAttribute(ring<ROWS, NSLOT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  ring<ROWS, NSLOT><<<148, 288, smem>>>(d, items, out);
  cudaEventRecord(a);
  for (int r = 0; r < 5; ++r) ring<ROWS, NSLOT><<<148, 288, smem>>>(d, items, out);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b); ms /= 5;
  const double c = ms * 1e-3 * 1.92e9 / ((items + 147) / 148);
  printf("TMA ring rows/slot=%d slots=%d (%5.1f KB in flight) imgs=%3d : %7.1f us %7.0f cycles/item %6.1f B/clk/SM %s\n", ROWS, NSLOT, ROWS * NSLOT * 1.6, imgs, ms * 1e3, c,
         320000.0 / c, cudaGetErrorString(cudaGetLastError()));
}
int main() {
  char* d; float* out; cudaMalloc(&d, (size_t)600 * 320000); cudaMalloc(&out, 4); cudaMemset(d, 0, (size_t)600 * 320000);
  for (int imgs : {600, 148}) { run<8, 2>(d, out, imgs); run<8, 4>(d, out, imgs); run<4, 8>(d, out, imgs); run<8, 6>(d, out, imgs); }
  return 0;
}
