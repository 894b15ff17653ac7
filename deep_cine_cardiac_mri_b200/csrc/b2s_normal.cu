// Launcher of the on-chip CineNet normal operator (normal_warp.cuh); other
// heights are composed from the expand / reduce operators by the Python layer.
#include "b2s_common.cuh"
#include "normal_warp.cuh"

using namespace b2s;

template <class P> static int launch_warp(const NormalArgs& a, long long items, cudaStream_t st) {
  B2S_CUDA(cudaFuncSetAttribute(normal_warp_kernel<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, P::SMEM_BYTES));
  const unsigned blocks = (unsigned)((items + P::ITEMS - 1) / P::ITEMS);
  normal_warp_kernel<P><<<blocks, P::NT, P::SMEM_BYTES, st>>>(a, items);
  return check_launch("normal_warp_kernel");
}

static int launch_normal(const float* x, const float* sens, const uint8_t* mask, const float* v, float* out, int mode,
                         const float* ssq, const float* bref, int b, int t, int c, int h, int w, void* stream, float* dot_part = nullptr) {
  if (!x || !sens || !mask || !v || !out || b < 0 || t < 0 || c < 0 || (mode >= 1 && (!ssq || !bref)))
    return fail(B2S_EINVAL, "b2s_normal_op: bad argument");
  NormalArgs a; a.x = (const cfloat*)x; a.sens = (const cfloat*)sens; a.mask = mask; a.vptr = v; a.out = (cfloat*)out;
  a.T = t; a.C = c; a.W = w; a.mode = mode; a.ssq = ssq; a.bref = (const cfloat*)bref; a.dot_part = dot_part;
  if ((h != 200 && h != 256) || w % 4 != 0 || w <= 0)
    return fail(B2S_EUNSUPPORTED, "b2s_normal_op: needs h in {200, 256} and w % 4 == 0");
  const long long items = (long long)b * t * (w / 4);
  if (items == 0) return B2S_OK;
  if (items > 0x7fffffffLL) return fail(B2S_EUNSUPPORTED, "b2s_normal_op: too many frames");
  const cudaStream_t st = (cudaStream_t)stream;
  // small launches (fewer work items than half the warps the GPU holds, e.g. one 15-frame slice per call): two warps share an
  // item, each takes half of the coils, the partial coil sums are added in a fixed order
  int sms = 148;
  { int dev = 0; if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); }
  if (h == 200 && w == 200 && c >= 2 && items * 2 <= (long long)sms * 12) return launch_warp<NormalWarpPlan<200, 200, 4, 3, 2>>(a, items, st);
  if (h == 200) return w == 200 ? launch_warp<NormalWarpPlan<200, 200, 3, 4>>(a, items, st) : launch_warp<NormalWarpPlan<200, 0, 3, 4>>(a, items, st);
  return w == 256 ? launch_warp<NormalWarpPlan<256, 256, 4, 2>>(a, items, st) : launch_warp<NormalWarpPlan<256, 0, 4, 2>>(a, items, st);
}

extern "C" int b2s_normal_op(const float* x, const float* sens, const uint8_t* mask, const float* v, float* out,
                             int b, int t, int c, int h, int w, void* stream) {
  return launch_normal(x, sens, mask, v, out, 0, nullptr, nullptr, b, t, c, h, w, stream);
}

extern "C" int b2s_normal_dc(const float* x, const float* sens, const uint8_t* mask, const float* v, const float* ssq,
                             const float* bref, float* out, int b, int t, int c, int h, int w, void* stream) {
  return launch_normal(x, sens, mask, v, out, 1, ssq, bref, b, t, c, h, w, stream);
}

extern "C" int b2s_normal_dc_abs(const float* x, const float* sens, const uint8_t* mask, const float* v, const float* ssq,
                                 const float* bref, float* out_abs, int b, int t, int c, int h, int w, void* stream) {
  return launch_normal(x, sens, mask, v, out_abs, 2, ssq, bref, b, t, c, h, w, stream);
}

// H x together with the per-item partial sums of <x, H x> (b * t * w / 4 floats): one CG iteration needs no separate dot
// kernel for <p, H p> (cinenet.py:159); the partials are summed in a fixed order by b2s_cg_update.
extern "C" int b2s_normal_op_dot(const float* x, const float* sens, const uint8_t* mask, const float* v, float* out, float* dot_partials,
                                 int b, int t, int c, int h, int w, void* stream) {
  if (!dot_partials) return fail(B2S_EINVAL, "b2s_normal_op_dot: null pointer");
  return launch_normal(x, sens, mask, v, out, 0, nullptr, nullptr, b, t, c, h, w, stream, dot_partials);
}
