// Launcher of the on-chip CineNet normal operator (normal_core.cuh); other
// heights are composed from the expand / reduce operators by the Python layer.
#include "b2s_common.cuh"
#include "normal_core.cuh"

using namespace b2s;

static int launch_normal(const float* x, const float* sens, const uint8_t* mask, const float* v, float* out, int mode,
                         const float* ssq, const float* bref, int b, int t, int c, int h, int w, void* stream) {
  if (!x || !sens || !mask || !v || !out || b < 0 || t < 0 || c < 0 || (mode == 1 && (!ssq || !bref)))
    return fail(B2S_EINVAL, "b2s_normal_op: bad argument");
  typedef NormalPlan<200, 20> P;
  if (h != P::H || w % P::XC != 0) return fail(B2S_EUNSUPPORTED, "b2s_normal_op: needs h == 200 and w % 20 == 0");
  const long long blocks = (long long)b * t * (w / P::XC);
  if (blocks == 0) return B2S_OK;
  if (blocks > 0x7fffffffLL) return fail(B2S_EUNSUPPORTED, "b2s_normal_op: too many frames");
  NormalArgs a; a.x = (const cfloat*)x; a.sens = (const cfloat*)sens; a.mask = mask; a.vptr = v; a.out = (cfloat*)out;
  a.T = t; a.C = c; a.W = w; a.mode = mode; a.ssq = ssq; a.bref = (const cfloat*)bref;
  B2S_CUDA(cudaFuncSetAttribute(normal_op_kernel<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, P::SMEM_BYTES));
  normal_op_kernel<P><<<(unsigned)blocks, P::NT, P::SMEM_BYTES, (cudaStream_t)stream>>>(a);
  return check_launch("normal_op_kernel");
}

extern "C" int b2s_normal_op(const float* x, const float* sens, const uint8_t* mask, const float* v, float* out,
                             int b, int t, int c, int h, int w, void* stream) {
  return launch_normal(x, sens, mask, v, out, 0, nullptr, nullptr, b, t, c, h, w, stream);
}

extern "C" int b2s_normal_dc(const float* x, const float* sens, const uint8_t* mask, const float* v, const float* ssq,
                             const float* bref, float* out, int b, int t, int c, int h, int w, void* stream) {
  return launch_normal(x, sens, mask, v, out, 1, ssq, bref, b, t, c, h, w, stream);
}
