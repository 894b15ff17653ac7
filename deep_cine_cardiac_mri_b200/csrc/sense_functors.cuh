// Prologue / epilogue functors that fuse the SENSE pointwise work into the
// first and last register stage of the 2-D FFT core (fft2_core.cuh).
//
//   prologues  ProPlain    x                                   (fft2c / ifft2c, utils/fftc.py:59-110)
//              ProExpand   S_c * x_t                           (sens_expand, models/varnet.py:181-185)
//              ProKspace   w(ky) * k,  w = a + b*mask[ky]      (sens_reduce, varnet.py:187-194; masked
//                                                               BackwardOperator, xpdnet.py:161-164)
//   epilogues  EpiPlain    store
//              EpiKspace   plain | k*m | soft-DC blend | k*m - ref
//                          (cinenet.py:129, varnet.py:281-282, xpdnet.py:295-298)
//              EpiReduce   out[o] += conj(mult[m]) * y         (coil sum varnet.py:192-194, or the
//                                                               frame sum of the sens-map gradient)
//
// "image" = linear index ((b*T + t)*C + c) of one H x W coil image.
#pragma once
#include "fft2_core.cuh"

namespace b2s {

#if defined(__CUDA_ARCH__)
__device__ __forceinline__ void red_add2(cfloat* p, float a, float b) {
  asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(a), "f"(b) : "memory");
}
#else
inline void red_add2(cfloat* p, float a, float b) { p->x += a; p->y += b; }
#endif

struct NoPre {};

// ------------------------------ prologues ---------------------------------- //
template <int W, bool INV> struct ProPlain {
  const cfloat* in; long long image_stride;
  struct Ctx { const cfloat* p; };
  B2S_HD Ctx ctx(long long image) const { Ctx c; c.p = in + image * image_stride; return c; }
  B2S_HD float row_weight(const Ctx&, int) const { return 1.f; }
  B2S_HD void load(const Ctx& c, int y, int x, float, float& re, float& im) const {
    const cfloat v = c.p[y * W + x];
    re = INV ? v.y : v.x; im = INV ? v.x : v.y;
  }
};

template <int W> struct ProExpand {
  const cfloat* img; const cfloat* sens; int T, C; long long hw;
  struct Ctx { const cfloat* a; const cfloat* s; };
  B2S_HD Ctx ctx(long long image) const {
    const long long c = image % C, bt = image / C, b = bt / T;
    Ctx k; k.a = img + bt * hw; k.s = sens + (b * C + c) * hw; return k;
  }
  B2S_HD float row_weight(const Ctx&, int) const { return 1.f; }
  B2S_HD void load(const Ctx& c, int y, int x, float, float& re, float& im) const {
    const cfloat a = c.a[y * W + x], s = c.s[y * W + x];
    re = a.x * s.x - a.y * s.y; im = a.x * s.y + a.y * s.x;
  }
};

// WMODE 0: w = 1 ; 1: w = mask[ky] ; 2: w = 1 - eta*mask[ky], eta = v/(1+v), v read from device memory
template <int W, int WMODE> struct ProKspace {
  const cfloat* k; const uint8_t* mask; const float* vptr; int C, H; long long hw;
  struct Ctx { const cfloat* p; const uint8_t* m; float wa, wb; };
  B2S_HD Ctx ctx(long long image) const {
    Ctx c; c.p = k + image * hw; c.m = WMODE ? mask + (image / C) * H : nullptr;
    c.wa = 1.f; c.wb = 0.f;
    if (WMODE == 1) { c.wa = 0.f; c.wb = 1.f; }
    if (WMODE == 2) { const float v = *vptr; c.wb = -v / (1.f + v); }
    return c;
  }
  B2S_HD float row_weight(const Ctx& c, int y) const {
    return WMODE ? c.wa + c.wb * (float)c.m[y] : 1.f;
  }
  B2S_HD void load(const Ctx& c, int y, int x, float w, float& re, float& im) const {
    const cfloat v = c.p[y * W + x];                 // inverse transform: feed swapped
    if (WMODE) { re = v.y * w; im = v.x * w; } else { re = v.y; im = v.x; }
  }
};

// ------------------------------ epilogues ---------------------------------- //
template <int W, bool INV> struct EpiPlain {
  cfloat* out; long long image_stride;
  struct Ctx { cfloat* p; };
  template <int G> using Pre = NoPre;
  B2S_HD Ctx ctx(long long image) const { Ctx c; c.p = out + image * image_stride; return c; }
  template <int G> B2S_HD void prefetch(const Ctx&, int, int, NoPre&) const {}
  template <int G> B2S_HD void store(const Ctx& c, int ky, int kx, float re, float im, const NoPre&, int) const {
    c.p[ky * W + kx] = INV ? make_c(im, re) : make_c(re, im);
  }
};

// MODE 0: k ; 1: k*m + 0.0 ; 2: (1-m) k + m (k + v ref)/(1+v) ; 3: k*m - ref
template <int W, int MODE> struct EpiKspace {
  cfloat* out; const cfloat* ref; const uint8_t* mask; const float* vptr; int C, H; long long hw;
  struct Ctx { cfloat* p; const cfloat* r; const uint8_t* m; float v; };
  template <int G> struct PreT { cfloat r[(MODE >= 2) ? G : 1]; uint8_t m[(MODE >= 1) ? G : 1]; };
  template <int G> using Pre = PreT<G>;
  B2S_HD Ctx ctx(long long image) const {
    Ctx c; c.p = out + image * hw;
    c.r = (MODE >= 2) ? ref + image * hw : nullptr;
    c.m = (MODE >= 1) ? mask + (image / C) * H : nullptr;
    c.v = (MODE == 2) ? *vptr : 0.f;
    return c;
  }
  template <int G> B2S_HD void prefetch(const Ctx& c, int m0, int kx, PreT<G>& pre) const {
    if (MODE >= 1) {
#pragma unroll
      for (int k = 0; k < G; ++k) pre.m[k] = c.m[m0 + 8 * k];
    }
    if (MODE >= 2) {
#pragma unroll
      for (int k = 0; k < G; ++k) {
        // DC only needs ref on sampled rows; the residual needs it everywhere
        if (MODE == 3 || pre.m[k]) pre.r[k] = c.r[(m0 + 8 * k) * W + kx];
        else pre.r[k] = make_c(0.f, 0.f);
      }
    }
  }
  template <int G> B2S_HD void store(const Ctx& c, int ky, int kx, float re, float im, const PreT<G>& pre, int k) const {
    if (MODE == 1) { if (!pre.m[k]) { re = 0.f; im = 0.f; } }
    if (MODE == 2) {
      if (pre.m[k]) { re = (re + c.v * pre.r[k].x) / (1.f + c.v); im = (im + c.v * pre.r[k].y) / (1.f + c.v); }
    }
    if (MODE == 3) {
      if (!pre.m[k]) { re = 0.f; im = 0.f; }
      re -= pre.r[k].x; im -= pre.r[k].y;
    }
    c.p[ky * W + kx] = make_c(re, im);
  }
};

// out[(b,t,c) . ostride] += conj(mult[(b,t,c) . mstride]) * ifft(k);  zero stride = reduced dim
template <int W> struct EpiReduce {
  cfloat* out; const cfloat* mult; int T, C;
  long long os_b, os_t, os_c, ms_b, ms_t, ms_c;
  struct Ctx { cfloat* o; const cfloat* m; };
  template <int G> struct PreT { cfloat s[G]; };
  template <int G> using Pre = PreT<G>;
  B2S_HD Ctx ctx(long long image) const {
    const long long c = image % C, bt = image / C, b = bt / T, t = bt % T;
    Ctx k; k.o = out + b * os_b + t * os_t + c * os_c; k.m = mult + b * ms_b + t * ms_t + c * ms_c;
    return k;
  }
  template <int G> B2S_HD void prefetch(const Ctx& c, int m0, int kx, PreT<G>& pre) const {
#pragma unroll
    for (int k = 0; k < G; ++k) pre.s[k] = c.m[(m0 + 8 * k) * W + kx];
  }
  template <int G> B2S_HD void store(const Ctx& c, int ky, int kx, float re, float im, const PreT<G>& pre, int k) const {
    const float yr = im, yi = re;                      // swap back (inverse transform)
    const cfloat s = pre.s[k];
    red_add2(c.o + ky * W + kx, yr * s.x + yi * s.y, yi * s.x - yr * s.y);
  }
};

}  // namespace b2s
