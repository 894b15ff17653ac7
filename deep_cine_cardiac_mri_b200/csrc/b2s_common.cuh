// Shared host-side helpers of the C-ABI translation units.
#pragma once
#include <cuda_runtime.h>
#include <atomic>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/b200sense.h"

namespace b2s {

void set_error(const char* fmt, ...);          // defined in b2s_abi.cu (thread-local text)

inline int fail(int code, const char* what) { set_error("%s", what); return code; }

extern std::atomic<unsigned long long> g_kernel_launches;
extern std::atomic<int> g_fused_path;          // b2s_abi.cu; kernel family of the fused plan sizes (b2s_set_fused_path)
extern std::atomic<int> g_sm_reserve;          // b2s_abi.cu; SMs the persistent kernels leave free (b2s_set_sm_reserve)   // b2s_abi.cu; kernels launched through this library

inline int check_launch(const char* what, int n_kernels = 1) {
  g_kernel_launches.fetch_add((unsigned long long)n_kernels, std::memory_order_relaxed);
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("%s: %s", what, cudaGetErrorString(e)); return B2S_ECUDA; }
  return B2S_OK;
}

#define B2S_CUDA(call)                                                              \
  do {                                                                              \
    const cudaError_t e_ = (call);                                                  \
    if (e_ != cudaSuccess) { ::b2s::set_error("%s: %s", #call, cudaGetErrorString(e_)); return B2S_ECUDA; } \
  } while (0)

// torch.fft normalisation -> scalar applied once in the epilogue
inline float norm_scale(int h, int w, int inverse, int norm) {
  const double n = (double)h * (double)w;
  if (norm == B2S_NORM_ORTHO) return (float)(1.0 / sqrt(n));
  if ((norm == B2S_NORM_BACKWARD && inverse) || (norm == B2S_NORM_FORWARD && !inverse)) return (float)(1.0 / n);
  return 1.f;
}

inline bool bad_norm(int norm) { return norm < 0 || norm > 2; }

// generic (any size) centred 2-D FFT, two passes through global memory (b2s_generic.cu)
int generic_fft2(const float* in, float* out, int64_t n_images, int h, int w, int inverse, float scale,
                 cudaStream_t st);

// strip-streamed fused kernels (b2s_strip.cu), h == w in {200, 256}.  `*unavailable` = 1 (and B2S_OK) when the
// launch could not be made (stream capturing before the workspace exists): the caller uses the on-chip kernels.
int strip_fft2c(int h, const float* in, float* out, int64_t n, int inverse, float scale, cudaStream_t st, int* unavailable);
int strip_expand(int h, const float* image, const float* sens, float* kspace, const float* ref, const uint8_t* mask,
                 const float* v, int mode, int t, int c, int64_t n, float scale, cudaStream_t st, int* unavailable);
int strip_reduce(int h, const float* kspace, const float* mult, float* out, const uint8_t* mask, const float* v,
                 int weight_mode, int over_frames, int t, int c, int64_t n, float scale, cudaStream_t st, int* unavailable);
int strip_ifft_weighted(int h, const float* kspace, float* y, const uint8_t* mask, const float* v, int weight_mode, int c,
                        int64_t n, float scale, cudaStream_t st, int* unavailable);

// element-wise helpers used by the non-fused fallbacks (b2s_pointwise.cu)
int launch_expand_product(const float* image, const float* sens, float* out, int b, int t, int c,
                          int64_t hw, cudaStream_t st);
int launch_kspace_epilogue(float* k, const float* ref, const uint8_t* mask, const float* v, int mode,
                           int64_t n_bt, int c, int h, int w, cudaStream_t st);
int launch_row_weight(const float* k, float* out, const uint8_t* mask, const float* v, int wmode,
                      int64_t n_bt, int c, int h, int w, cudaStream_t st);
int launch_coil_reduce(const float* y, const float* mult, float* out, int over_frames, int b, int t,
                       int c, int64_t hw, cudaStream_t st);

}  // namespace b2s
