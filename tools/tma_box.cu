// dev microbench: per-SM ingest of the Phase A access pattern through TMA TENSOR-MAP boxes (cp.async.bulk.tensor.4d).
// The coil image (200 x 200 complex) is described as a 4-D tensor (x = 200, g = 25, j = 8, image) of 8-byte elements, so ONE
// instruction fetches the 8 rows {g + 25 j} of a Phase A task group: a 12.8 KB box (BOXW = 200) or one 40-column group of
// it (BOXW = 40, 2.56 KB).  One producer thread per CTA keeps a ring of NSLOT boxes in flight (mbarrier full/empty
// pairs); 256 consumer threads wait for a box, read their share of it from shared memory and release it.  One CTA per SM,
// 148 CTAs, shared memory padded to the real kernel's footprint.  Prints bytes per clock and SM for images streamed from HBM
// (600 distinct images = 192 MB) and from L2 (40 images re-read).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_bin/tma_box tools/tma_box.cu && tools/_bin/tma_box
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* b, int n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(n)); }
__device__ __forceinline__ void mbar_expect(unsigned long long* b, unsigned bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(unsigned long long* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ void mbar_wait(unsigned long long* b, unsigned parity) {
  asm volatile("{ .reg .pred p; WAIT_%=: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1; @p bra DONE_%=; bra WAIT_%=; DONE_%=: }" ::"r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_box(void* dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, unsigned long long* bar) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
               ::"r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar)) : "memory");
}

template <int BOXW, int NSLOT> __global__ void __launch_bounds__(288, 1)
ring(const __grid_constant__ CUtensorMap map, int n_images, int wrap, float* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) unsigned long long full[NSLOT], empty[NSLOT];
  constexpr int BOX_BYTES = BOXW * 8 * 8, GROUPS = 25 * (200 / BOXW);
  const int tid = threadIdx.x;
  if (tid == 0) { for (int s = 0; s < NSLOT; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 256); } asm volatile("fence.mbarrier_init.release.cluster;"); }
  __syncthreads();
  float acc = 0.f;
  if (tid >= 256) {
    if (tid == 256) {
      int it = 0;
      for (int img = blockIdx.x; img < n_images; img += gridDim.x)
        for (int g = 0; g < GROUPS; ++g, ++it) {
          const int s = it % NSLOT; const unsigned ph = (it / NSLOT) & 1;
          if (it >= NSLOT) mbar_wait(&empty[s], ph ^ 1);
          mbar_expect(&full[s], BOX_BYTES);
          tma_box(smem + (size_t)s * BOX_BYTES, &map, (g / 25) * BOXW, g % 25, 0, img % wrap, &full[s]);
        }
    }
  } else {
    int it = 0;
    for (int img = blockIdx.x; img < n_images; img += gridDim.x)
      for (int g = 0; g < GROUPS; ++g, ++it) {
        const int s = it % NSLOT; const unsigned ph = (it / NSLOT) & 1;
        mbar_wait(&full[s], ph);
        const float2* p = reinterpret_cast<const float2*>(smem + (size_t)s * BOX_BYTES);
        for (int e = tid; e < BOX_BYTES / 8; e += 256) { const float2 v = p[e]; acc += v.x + v.y; }   // consume the whole box
        mbar_arrive(&empty[s]);
      }
  }
  if (acc == 123.456f) out[0] = acc;
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                             CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int BOXW, int NSLOT> void run(EncodeFn enc, char* d, float* out, int imgs, int wrap) {
  CUtensorMap map;
  const cuuint64_t dims[4] = {200, 25, 8, (cuuint64_t)wrap};
  const cuuint64_t strides[3] = {1600, 40000, 320000};
  const cuuint32_t box[4] = {BOXW, 1, 8, 1}, estr[4] = {1, 1, 1, 1};
  CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_UINT64, 4, d, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed: %d\n", (int)r); return; }
  constexpr int BOX_BYTES = BOXW * 64;
  const int smem = NSLOT * BOX_BYTES + 1024;
  cudaFuncSetAttribute(ring<BOXW, NSLOT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  // ballast kernel attribute: keep one CTA per SM like the real kernel (it needs ~190 KB for B); here 1 CTA/SM through launch_bounds + grid = 148
  ring<BOXW, NSLOT><<<148, 288, smem>>>(map, imgs, wrap, out);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  cudaEventRecord(a);
  for (int rep = 0; rep < 5; ++rep) ring<BOXW, NSLOT><<<148, 288, smem>>>(map, imgs, wrap, out);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b); ms /= 5;
  const double rounds = (imgs + 147) / 148, cyc = ms * 1e-3 * 1.9e9 / rounds;
  printf("box %3d cols x 8 rows (%5.1f KB) x %2d slots = %5.1f KB in flight, %4d images (%s): %7.1f us  %7.0f cycles/image  %6.1f B/clk/SM  %s\n", BOXW, BOX_BYTES / 1024.0,
         NSLOT, NSLOT * BOX_BYTES / 1024.0, imgs, wrap < 148 ? "L2-resident" : "HBM", ms * 1e3, cyc, 320000.0 / cyc, cudaGetErrorString(cudaGetLastError()));
}

int main() {
  EncodeFn enc = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &q) != cudaSuccess || !enc) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
  char* d; float* out;
  cudaMalloc(&d, (size_t)600 * 320000); cudaMalloc(&out, 4); cudaMemset(d, 0, (size_t)600 * 320000);
  for (int wrap : {600, 40}) {
    const int imgs = 592;
    run<200, 2>(enc, d, out, imgs, wrap); run<200, 3>(enc, d, out, imgs, wrap); run<200, 6>(enc, d, out, imgs, wrap); run<200, 12>(enc, d, out, imgs, wrap);
    run<40, 8>(enc, d, out, imgs, wrap); run<40, 16>(enc, d, out, imgs, wrap); run<40, 32>(enc, d, out, imgs, wrap); run<40, 64>(enc, d, out, imgs, wrap);
  }
  return 0;
}
