// CPU emulation of the fused FFT kernels: compiles the very same host/device
// phase functions with g++ and runs them sequentially.  Used by the not-gpu
// tests to check the kernel index/twiddle math against the oracle without a GPU.
#include <cstdint>
#include "sense_functors.cuh"
#include "fft2_kernel.cuh"
#include "fft2_whole.cuh"
#include "fft2_packed.cuh"

using namespace b2s;
typedef Plan<200, 200, 256, 1, 2> P200;    // half split
typedef Plan<200, 200, 256, 2, 2> P200V;   // half split, 128-bit accesses in Phases A and C
typedef Plan<200, 200, 256, 2, 2, 1> P200W; // half split, 128-bit Phase A only (what sens_expand ships)
typedef Plan<200, 200, 128, 1, 4> P200Q;   // quarter split (2 CTAs/SM on the device)
typedef Plan<256, 256, 256, 1, 4> P256;
typedef Plan<256, 256, 256, 2, 4, 1> P256W;  // 128-bit Phase A (what sens_expand ships at 256 x 256)

static float norm_scale(int h, int w, int inverse, int norm) {
  // norm: 0 "backward", 1 "ortho", 2 "forward" (torch.fft semantics)
  const double n = (double)h * w;
  double s = 1.0;
  if (norm == 1) s = 1.0 / sqrt(n);
  else if ((norm == 0 && inverse) || (norm == 2 && !inverse)) s = 1.0 / n;
  return (float)s;
}

static int g_variant = 0;   // 200x200: 0 half split, 1 half split with 128-bit accesses, 2 quarter split, 3 paired (cluster), 4 wide Phase A,
                            // 5 half split with the soft-DC row fix-up epilogue (EpiDCFix), 6 / 7 whole image (parking lot) with a
                            // 5- / 2-step load queue (7: soft DC through the row fix-up), 8 / 9 the same with packed two-transform arithmetic
template <class P, class Pro, class Epi> static void whole_if(const Pro& pro, const Epi& epi, float scale, long long n, int qd) {
  if constexpr (P::H == 200 && P::FOLD == 2 && P::NC == 1) {
    if (qd == 5) fft2_whole_emulate<P, Pro, Epi, 5, 1>(pro, epi, scale, n); else fft2_whole_emulate<P, Pro, Epi, 2, 2>(pro, epi, scale, n);
  } else fft2_half_emulate<P>(pro, epi, scale, n);
}
template <class P, class Pro, class Epi> static void packed_if(const Pro& pro, const Epi& epi, float scale, long long n, int qd) {
  if constexpr (P::H == 200 && P::FOLD == 2 && P::NC == 1) {
    typedef PackPlan200<256> PP;
    if (qd == 5) fft2_packed_emulate<PP, Pro, Epi, 5, 1>(pro, epi, scale, n); else fft2_packed_emulate<PP, Pro, Epi, 2, 2>(pro, epi, scale, n);
  } else fft2_half_emulate<P>(pro, epi, scale, n);
}
#define EMULATE(P, pro, epi, scale, n) do { if (g_variant >= 8 && g_variant <= 12) { packed_if<P>(pro, epi, scale, n, g_variant == 8 ? 5 : 2); break; } if (g_variant == 6 || g_variant == 7) { whole_if<P>(pro, epi, scale, n, g_variant == 6 ? 5 : 2); break; } if (g_variant == 3 && P::FOLD == 2 && P::NC == 1) fft2_pair_emulate_if<P>(pro, epi, scale, n); else fft2_half_emulate<P>(pro, epi, scale, n); } while (0)
template <class P, class Pro, class Epi> static void fft2_pair_emulate_if(const Pro& pro, const Epi& epi, float scale, long long n) {
  if constexpr (P::FOLD == 2 && P::NC == 1) fft2_pair_emulate<P>(pro, epi, scale, n); else fft2_half_emulate<P>(pro, epi, scale, n);
}

template <class P> static int t_fft2c(const float* in, float* out, long long n_images, int inverse, int norm) {
  constexpr int H = P::H, W = P::W;
  const float scale = norm_scale(H, W, inverse, norm) * centre_sign<P>();
  const long long hw = (long long)H * W;
  if (inverse) {
    ProPlain<H, W, true> pro{(const cfloat*)in, hw};
    EpiPlain<H, W, true> epi{(cfloat*)out, hw};
    EMULATE(P, pro, epi, scale, n_images);
  } else {
    ProPlain<H, W, false> pro{(const cfloat*)in, hw};
    EpiPlain<H, W, false> epi{(cfloat*)out, hw};
    EMULATE(P, pro, epi, scale, n_images);
  }
  return 0;
}

template <class P> static int t_expand(const float* img, const float* sens, float* kout, const float* ref, const uint8_t* mask,
                                       const float* v, int mode, int b, int t, int c, int norm) {
  constexpr int H = P::H, W = P::W;
  const float scale = norm_scale(H, W, 0, norm) * centre_sign<P>();
  const long long hw = (long long)H * W, n = (long long)b * t * c;
  ProExpand<H, W> pro{(const cfloat*)img, (const cfloat*)sens, t, c, hw};
#define RUN(M) { EpiKspace<H, W, M> epi{(cfloat*)kout, (const cfloat*)ref, mask, v, c, hw}; EMULATE(P, pro, epi, scale, n); }
  if (mode == 2 && (g_variant >= 10 && g_variant <= 12)) {   // packed whole-image kernel, staged soft DC (11: capacity 3 + 3 rows -> general path, 12: 14 + 12 rows -> rows in the holes of B)
    if constexpr (P::H == 200 && P::FOLD == 2 && P::NC == 1) {
      EpiDCStage<H, W> epi{(cfloat*)kout, (const cfloat*)ref, mask, v, c, hw};
      fft2_packed_emulate<PackPlan200<256>, ProExpand<H, W>, EpiDCStage<H, W>, 2, 2>(pro, epi, scale, n, g_variant == 11 ? 3 : g_variant == 12 ? 14 : 0);
      return 0;
    }
  }
  if (mode == 2 && (g_variant == 5 || g_variant == 7 || g_variant == 9)) {
    EpiDCFix<H, W> epi{(cfloat*)kout, (const cfloat*)ref, mask, v, c, hw, 1};
    if (g_variant == 9) packed_if<P>(pro, epi, scale, n, 2); else if (g_variant == 7) whole_if<P>(pro, epi, scale, n, 2); else fft2_half_emulate<P>(pro, epi, scale, n);
    return 0;
  }
  if (mode == 0) RUN(0) else if (mode == 1) RUN(1) else if (mode == 2) RUN(2) else if (mode == 3) RUN(3) else return 1;
#undef RUN
  return 0;
}

template <class P> static int t_reduce(const float* k, const float* mult, float* out, const uint8_t* mask, const float* v, int wmode,
                                       int over_frames, int b, int t, int c, int norm) {
  constexpr int H = P::H, W = P::W;
  const float scale = norm_scale(H, W, 1, norm) * centre_sign<P>();
  const long long hw = (long long)H * W, n = (long long)b * t * c;
  EpiReduce<H, W> epi;
  epi.out = (cfloat*)out; epi.mult = (const cfloat*)mult; epi.T = t; epi.C = c;
  if (!over_frames) {   // out (b,t,h,w) = sum_c conj(S[b,c]) y[b,t,c]
    epi.os_b = t * hw; epi.os_t = hw; epi.os_c = 0; epi.ms_b = c * hw; epi.ms_t = 0; epi.ms_c = hw;
    for (long long i = 0; i < (long long)b * t * hw * 2; ++i) out[i] = 0.f;
  } else {              // out (b,c,h,w) = sum_t conj(X[b,t]) y[b,t,c]
    epi.os_b = c * hw; epi.os_t = 0; epi.os_c = hw; epi.ms_b = t * hw; epi.ms_t = hw; epi.ms_c = 0;
    for (long long i = 0; i < (long long)b * c * hw * 2; ++i) out[i] = 0.f;
  }
#define RUN(M) { ProKspace<H, W, M> pro{(const cfloat*)k, mask, v, c, hw}; EMULATE(P, pro, epi, scale, n); }
  if (wmode == 0) RUN(0) else if (wmode == 1) RUN(1) else if (wmode == 2) RUN(2) else return 1;
#undef RUN
  return 0;
}

#define DISPATCH(FN, ...)                                                          \
  if (h == 200 && w == 200) return g_variant == 2 ? FN<P200Q>(__VA_ARGS__) : g_variant == 1 ? FN<P200V>(__VA_ARGS__) : g_variant == 4 ? FN<P200W>(__VA_ARGS__) : FN<P200>(__VA_ARGS__); \
  if (h == 256 && w == 256) return g_variant == 4 ? FN<P256W>(__VA_ARGS__) : FN<P256>(__VA_ARGS__);                          \
  return 2;

extern "C" {

void emu_set_variant(int v) { g_variant = v; }

int emu_fft2c(const float* in, float* out, long long n_images, int h, int w, int inverse, int norm) {
  DISPATCH(t_fft2c, in, out, n_images, inverse, norm)
}

int emu_sens_expand(const float* img, const float* sens, float* kout, const float* ref, const uint8_t* mask,
                    const float* v, int mode, int b, int t, int c, int h, int w, int norm) {
  DISPATCH(t_expand, img, sens, kout, ref, mask, v, mode, b, t, c, norm)
}

int emu_sens_reduce(const float* k, const float* mult, float* out, const uint8_t* mask, const float* v, int wmode,
                    int over_frames, int b, int t, int c, int h, int w, int norm) {
  DISPATCH(t_reduce, k, mult, out, mask, v, wmode, over_frames, b, t, c, norm)
}

}  // extern "C"

#include "normal_warp.cuh"
// warp-private on-chip normal operator (the product kernel of b2s_normal_op / b2s_normal_dc), executed lane by lane:
// mode 0 = normal operator, mode 1 = image-domain VarNet cascade; fixed == 1 takes the compile-time-width plan
extern "C" int emu_normal_warp(const float* x, const float* sens, const uint8_t* mask, const float* v, const float* ssq,
                               const float* bref, float* out, int mode, int b, int t, int c, int h, int w, int fixed) {
  if ((h != 200 && h != 256) || w % 4) return 2;
  if (fixed && w != h) return 2;
  NormalArgs a; a.x = (const cfloat*)x; a.sens = (const cfloat*)sens; a.mask = mask; a.vptr = v; a.out = (cfloat*)out;
  a.T = t; a.C = c; a.W = w; a.mode = mode; a.ssq = ssq; a.bref = (const cfloat*)bref; a.dot_part = nullptr;
  const long long n = (long long)b * t;
  if (h == 200 && fixed == 2) normal_warp_emulate<NormalWarpPlan<200, 200, 4, 3, 2>>(a, n);        // two warps per item, half of the coils each
  else if (h == 200) { if (fixed) normal_warp_emulate<NormalWarpPlan<200, 200, 3, 4>>(a, n); else normal_warp_emulate<NormalWarpPlan<200, 0, 3, 4>>(a, n); }
  else          { if (fixed) normal_warp_emulate<NormalWarpPlan<256, 256, 4, 2>>(a, n); else normal_warp_emulate<NormalWarpPlan<256, 0, 4, 2>>(a, n); }
  return 0;
}

// every generated codelet against a direct DFT (double), returns the worst relative error
template <int N> static double codelet_err() {
  float re[N], im[N]; double xr[N], xi[N];
  for (int j = 0; j < N; ++j) { xr[j] = sin(1.0 + j * 0.7) + 0.1 * j; xi[j] = cos(2.0 + j * 1.3); re[j] = (float)xr[j]; im[j] = (float)xi[j]; }
  Dft<N>::run(re, im);
  double err = 0, mx = 0;
  for (int k = 0; k < N; ++k) {
    double sr = 0, si = 0;
    for (int j = 0; j < N; ++j) { const double a = -2.0 * M_PI * j * k / N; sr += xr[j] * cos(a) - xi[j] * sin(a); si += xr[j] * sin(a) + xi[j] * cos(a); }
    err = fmax(err, hypot(sr - re[k], si - im[k])); mx = fmax(mx, hypot(sr, si));
  }
  return err / mx;
}
extern "C" double emu_codelet_worst_error() {
  double e = 0;
  e = fmax(e, codelet_err<2>()); e = fmax(e, codelet_err<3>()); e = fmax(e, codelet_err<4>()); e = fmax(e, codelet_err<5>());
  e = fmax(e, codelet_err<8>()); e = fmax(e, codelet_err<10>()); e = fmax(e, codelet_err<12>()); e = fmax(e, codelet_err<15>());
  e = fmax(e, codelet_err<16>()); e = fmax(e, codelet_err<20>()); e = fmax(e, codelet_err<24>()); e = fmax(e, codelet_err<25>());
  e = fmax(e, codelet_err<30>()); e = fmax(e, codelet_err<32>()); e = fmax(e, codelet_err<40>());
  return e;
}
