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
// "image" = linear index ((b*T + t)*C + c) of one H x W coil image.  Every
// functor hands the core ONE pointer set per task (`task_ptr`); all further
// addressing is compile-time constant offsets, so loads/stores carry immediates.
// `load<NC>` / `store<NC>` move NC adjacent complex values with one 8*NC-byte access.
// `l2_prefetch` issues bulk L2 prefetches for data a later phase / work item reads.
#pragma once
#include "fft2_core.cuh"

namespace b2s {

#if defined(__CUDA_ARCH__)
template <int NC> __device__ __forceinline__ void red_add(cfloat* p, const float* re, const float* im);
template <> __device__ __forceinline__ void red_add<1>(cfloat* p, const float* re, const float* im) {
  asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(re[0]), "f"(im[0]) : "memory");
}
template <> __device__ __forceinline__ void red_add<2>(cfloat* p, const float* re, const float* im) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(re[0]), "f"(im[0]), "f"(re[1]), "f"(im[1]) : "memory");
}
// one instruction prefetches `bytes` (multiple of 16) contiguous bytes into L2
__device__ __forceinline__ void l2_prefetch_bulk(const void* p, unsigned bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}
#else
template <int NC> inline void red_add(cfloat* p, const float* re, const float* im) {
  for (int n = 0; n < NC; ++n) { p[n].x += re[n]; p[n].y += im[n]; }
}
inline void l2_prefetch_bulk(const void*, unsigned) {}
#endif

// Streaming global loads bypass L1 (ld.global.cg): with ~200 KB of the SM's 256 KB used as shared
// memory the L1 is only a few hundred lines, and an allocating load can only be in flight while it
// owns an L1 line - that capped memory-level parallelism at ~16 KB per SM (measured, profiles/).
#if defined(__CUDA_ARCH__)
template <int NC> __device__ __forceinline__ cvec<NC> ldv(const cfloat* p);
template <> __device__ __forceinline__ cvec<1> ldv<1>(const cfloat* p) {
  const float2 t = __ldcg(reinterpret_cast<const float2*>(p));
  cvec<1> r; r.v[0].x = t.x; r.v[0].y = t.y; return r;
}
template <> __device__ __forceinline__ cvec<2> ldv<2>(const cfloat* p) {
  const float4 t = __ldcg(reinterpret_cast<const float4*>(p));
  cvec<2> r; r.v[0].x = t.x; r.v[0].y = t.y; r.v[1].x = t.z; r.v[1].y = t.w; return r;
}
// L2 cache-policy hints: operands that are re-read many times per launch (sens maps: once per frame, the
// coil-combined image: once per coil) are kept with evict_last, the k-space streams (read or written once)
// are marked evict_first so that they do not push those operands out of L2.
__device__ __forceinline__ unsigned long long l2_keep() { unsigned long long p; asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p)); return p; }
__device__ __forceinline__ unsigned long long l2_stream() { unsigned long long p; asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p)); return p; }
template <int NC> __device__ __forceinline__ cvec<NC> ldv_pol(const cfloat* p, unsigned long long pol);
template <> __device__ __forceinline__ cvec<1> ldv_pol<1>(const cfloat* p, unsigned long long pol) {
  cvec<1> r;
  asm volatile("ld.global.L1::no_allocate.L2::cache_hint.v2.f32 {%0, %1}, [%2], %3;" : "=f"(r.v[0].x), "=f"(r.v[0].y) : "l"(p), "l"(pol));
  return r;
}
template <> __device__ __forceinline__ cvec<2> ldv_pol<2>(const cfloat* p, unsigned long long pol) {
  cvec<2> r;
  asm volatile("ld.global.L1::no_allocate.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;"
               : "=f"(r.v[0].x), "=f"(r.v[0].y), "=f"(r.v[1].x), "=f"(r.v[1].y) : "l"(p), "l"(pol));
  return r;
}
template <int NC> __device__ __forceinline__ void stv_pol(cfloat* p, const cvec<NC>& v, unsigned long long pol);
template <> __device__ __forceinline__ void stv_pol<1>(cfloat* p, const cvec<1>& v, unsigned long long pol) {
  asm volatile("st.global.L2::cache_hint.v2.f32 [%0], {%1, %2}, %3;" ::"l"(p), "f"(v.v[0].x), "f"(v.v[0].y), "l"(pol) : "memory");
}
template <> __device__ __forceinline__ void stv_pol<2>(cfloat* p, const cvec<2>& v, unsigned long long pol) {
  asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(p), "f"(v.v[0].x), "f"(v.v[0].y), "f"(v.v[1].x), "f"(v.v[1].y), "l"(pol) : "memory");
}
template <int NC> __device__ __forceinline__ cvec<NC> ldv_keep(const cfloat* p) { return ldv_pol<NC>(p, l2_keep()); }
template <int NC> __device__ __forceinline__ cvec<NC> ldv_stream(const cfloat* p) { return ldv_pol<NC>(p, l2_stream()); }
template <int NC> __device__ __forceinline__ void stv_stream(cfloat* p, const cvec<NC>& v) { stv_pol<NC>(p, v, l2_stream()); }
#else
template <int NC> inline cvec<NC> ldv(const cfloat* p) { return *reinterpret_cast<const cvec<NC>*>(p); }
template <int NC> inline cvec<NC> ldv_keep(const cfloat* p) { return ldv<NC>(p); }
template <int NC> inline cvec<NC> ldv_stream(const cfloat* p) { return ldv<NC>(p); }
template <int NC> inline void stv_stream(cfloat* p, const cvec<NC>& v) { *reinterpret_cast<cvec<NC>*>(p) = v; }
#endif
template <int NC> B2S_HD void stv(cfloat* p, const cvec<NC>& v) { *reinterpret_cast<cvec<NC>*>(p) = v; }

// `bytes` contiguous bytes in chunks of 32 KB, one chunk per thread
B2S_HD void prefetch_span(const void* base, long long bytes, int tid) {
  const long long chunk = 32768;
  const long long off = (long long)tid * chunk;
  if (off < bytes) l2_prefetch_bulk((const char*)base + off, (unsigned)((bytes - off) < chunk ? (bytes - off) : chunk));
}

// Sampled rows of the reference k-space of `image` into L2, one 128-byte line per instruction through the ordinary
// load/store path (the per-SM bulk-copy unit needs ~600 cycles per 1600-byte row whatever its size: 50 rows per image
// would outlast the image).  Called one image AHEAD of their use, with evict_last so that the streamed output does
// not push them out again.
template <int H, int W>
B2S_HD void prefetch_sampled_rows(const cfloat* ref_image, const uint8_t* mask_row, int tid, int nt) {
#if defined(__CUDA_ARCH__)
  constexpr int LPR = (W * 8 + 127) / 128;
  for (int e = tid; e < H * LPR; e += nt) {
    const int y = e / LPR, l = e - y * LPR;
    if (mask_row[y]) asm volatile("prefetch.global.L2::evict_last [%0];" ::"l"((const char*)(ref_image + (long long)y * W) + 128 * l));
  }
#else
  (void)ref_image; (void)mask_row; (void)tid; (void)nt;
#endif
}

// ------------------------------ prologues ---------------------------------- //
template <int H, int W, bool INV> struct ProPlain {
  const cfloat* in; long long image_stride;
  struct Ctx { const cfloat* p; };
  typedef const cfloat* Ptr;
  B2S_HD Ctx ctx(long long image) const { Ctx c; c.p = in + image * image_stride; return c; }
  // ~one task (32-40 loads) in flight per thread; the paired kernel holds both parities, so it keeps fewer
  template <int R, bool PAIRED = false, int NC = 1> static constexpr int qdepth() { return (PAIRED || NC > 1) ? 2 : (R <= 5 ? R : R / 2); }
  template <int NC> struct Unit { cvec<NC> a[8]; };
  // raw loads of the 8 rows g + G*j of one column group (row stride RS = G*W elements)
  template <int NC, int RS> B2S_HD void fetch(const Ctx& c, int, int off, Unit<NC>& u) const {
#ifndef B2S_NOLOAD
    const cfloat* p = c.p + off;
#pragma unroll
    for (int j = 0; j < 8; ++j) u.a[j] = ldv<NC>(p + j * RS);
#else   /* dev-only: compute floor of Phase A without memory traffic */
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
      for (int n = 0; n < NC; ++n) u.a[j].v[n] = make_c((float)(off + j), (float)(off - n));
#endif
  }
  template <int NC> B2S_HD void value(const Unit<NC>& u, int j, float* re, float* im) const {
#pragma unroll
    for (int n = 0; n < NC; ++n) { re[n] = INV ? u.a[j].v[n].y : u.a[j].v[n].x; im[n] = INV ? u.a[j].v[n].x : u.a[j].v[n].y; }
  }
  B2S_HD void l2_prefetch(long long image, int tid) const { prefetch_span(in + image * image_stride, (long long)H * W * 8, tid); }
};

template <int H, int W> struct ProExpand {
  const cfloat* img; const cfloat* sens; int T, C; long long hw;
  struct Ctx { const cfloat* a; const cfloat* s; };
  struct Ptr { const cfloat* a; const cfloat* s; };
  B2S_HD Ctx ctx(long long image) const {
    const long long c = image % C, bt = image / C, b = bt / T;
    Ctx k; k.a = img + bt * hw; k.s = sens + (b * C + c) * hw; return k;
  }
  // steps of raw loads in flight per thread (one step = 8 rows x {x, S} = 32 registers per column): 2 with 64-bit loads;
  // with 128-bit loads 1 at 200 x 200 (R = 5; 2 spill) and 2 at 256 x 256 (R = 8).  A 4-step queue at 200 x 200 (loop
  // trip of 4 tasks, PhaseA::TT) wins operator microbenchmarks on random data (187 -> 181 us) but loses inside the
  // 12-cascade pipeline on cine data (184.8 -> 188.8 us per launch, profiles/r1_analysis.md), so it is not used.
  template <int R, bool PAIRED = false, int NC = 1> static constexpr int qdepth() { return PAIRED ? 1 : (NC > 1 ? (R == 5 ? 1 : 2) : 2); }   // x 16 loads in flight per thread
  template <int NC> struct Unit { cvec<NC> a[8], s[8]; };
  template <int NC, int RS> B2S_HD void fetch(const Ctx& c, int, int off, Unit<NC>& u) const {
    const cfloat* pa = c.a + off; const cfloat* ps = c.s + off;
#pragma unroll
    for (int j = 0; j < 8; ++j) { u.a[j] = ldv_keep<NC>(pa + j * RS); u.s[j] = ldv_keep<NC>(ps + j * RS); }
  }
  template <int NC> B2S_HD void value(const Unit<NC>& u, int j, float* re, float* im) const {
#pragma unroll
    for (int n = 0; n < NC; ++n) {
      const cfloat a = u.a[j].v[n], s = u.s[j].v[n];
      re[n] = a.x * s.x - a.y * s.y;
      im[n] = a.x * s.y + a.y * s.x;
    }
  }
  B2S_HD void l2_prefetch(long long, int) const {}             // image and maps are L2 resident
};

// WMODE 0: w = 1 ; 1: w = mask[ky] ; 2: w = 1 - eta*mask[ky], eta = v/(1+v), v read from device memory
template <int H, int W, int WMODE> struct ProKspace {
  const cfloat* k; const uint8_t* mask; const float* vptr; int C; long long hw;
  struct Ctx { const cfloat* p; const uint8_t* m; float wa, wb; };
  typedef const cfloat* Ptr;
  B2S_HD Ctx ctx(long long image) const {
    Ctx c; c.p = k + image * hw; c.m = WMODE ? mask + (image / C) * H : nullptr;
    c.wa = 1.f; c.wb = 0.f;
    if (WMODE == 1) { c.wa = 0.f; c.wb = 1.f; }
    if (WMODE == 2) { const float v = *vptr; c.wb = -v / (1.f + v); }
    return c;
  }
  template <int R, bool PAIRED = false, int NC = 1> static constexpr int qdepth() { return (PAIRED || NC > 1) ? 2 : (R <= 5 ? R : R / 2); }
  template <int NC> struct Unit { cvec<NC> a[8]; float w[WMODE ? 8 : 1]; };
  template <int NC, int RS> B2S_HD void fetch(const Ctx& c, int g, int off, Unit<NC>& u) const {
    const cfloat* p = c.p + off;
#pragma unroll
    for (int j = 0; j < 8; ++j) u.a[j] = ldv<NC>(p + j * RS);
    if (WMODE) {
#pragma unroll
      for (int j = 0; j < 8; ++j) u.w[j] = c.wa + c.wb * (float)c.m[g + j * (RS / W)];
    }
  }
  template <int NC> B2S_HD void value(const Unit<NC>& u, int j, float* re, float* im) const {
#pragma unroll
    for (int n = 0; n < NC; ++n) {                       // inverse transform: feed swapped
      if (WMODE) { re[n] = u.a[j].v[n].y * u.w[j]; im[n] = u.a[j].v[n].x * u.w[j]; }
      else { re[n] = u.a[j].v[n].y; im[n] = u.a[j].v[n].x; }
    }
  }
  B2S_HD void l2_prefetch(long long image, int tid) const { prefetch_span(k + image * hw, (long long)H * W * 8, tid); }
};

// ------------------------------ epilogues ---------------------------------- //
template <int H, int W, bool INV> struct EpiPlain {
  static constexpr bool FIXUP = false;
  cfloat* out; long long image_stride;
  struct Ctx { cfloat* p; };
  typedef cfloat* Ptr;
  B2S_HD Ctx ctx(long long image, const uint8_t*) const { Ctx c; c.p = out + image * image_stride; return c; }
  B2S_HD void stage_mask(long long, uint8_t*, int, int) const {}
  B2S_HD Ptr task_ptr(const Ctx& c, int m, int kx) const { return c.p + m * W + kx; }
  template <int G, int NC> struct Pre {};
  template <int G, int NC> B2S_HD void prefetch(Ptr, Pre<G, NC>&) const {}
  template <int G, int NC> B2S_HD void store(Ptr p, int k, const float* re, const float* im, const Pre<G, NC>&) const {
    cvec<NC> v;
#pragma unroll
    for (int n = 0; n < NC; ++n) v.v[n] = INV ? make_c(im[n], re[n]) : make_c(re[n], im[n]);
#ifdef B2S_NOSTORE  /* dev-only: compute floor of Phase C */
    if (v.v[0].x == 123.456f)
#endif
    stv<NC>(p + 8 * k * W, v);
  }
  B2S_HD void l2_prefetch(long long, int, int, int) const {}
  B2S_HD void l2_prefetch_ahead(long long, int, int) const {}
};

// MODE 0: k ; 1: k*m + 0.0 ; 2: (1-m) k + m (k + v ref)/(1+v) ; 3: k*m - ref
template <int H, int W, int MODE> struct EpiKspace {
  static constexpr bool FIXUP = false;
  cfloat* out; const cfloat* ref; const uint8_t* mask; const float* vptr; int C; long long hw;
  struct Ctx { cfloat* p; const cfloat* r; const uint8_t* m; float v, inv1v; };
  struct Ptr { cfloat* p; const cfloat* r; const uint8_t* m; float v, inv1v; };
  // the (b,t) mask row of this item, staged once per item into shared memory (mrow)
  B2S_HD void stage_mask(long long image, uint8_t* mrow, int tid, int nt) const {
    if (MODE >= 1) { const uint8_t* src = mask + (image / C) * H; for (int y = tid; y < H; y += nt) mrow[y] = src[y]; }
  }
  B2S_HD Ctx ctx(long long image, const uint8_t* mrow) const {
    Ctx c; c.p = out + image * hw;
    c.r = (MODE >= 2) ? ref + image * hw : nullptr;
    c.m = (MODE >= 1) ? mrow : nullptr;
    c.v = (MODE == 2) ? *vptr : 0.f;
    c.inv1v = 1.f / (1.f + c.v);
    return c;
  }
  B2S_HD Ptr task_ptr(const Ctx& c, int m, int kx) const {
    Ptr t; const int off = m * W + kx;
    t.p = c.p + off; t.r = (MODE >= 2) ? c.r + off : nullptr; t.m = (MODE >= 1) ? c.m + m : nullptr; t.v = c.v; t.inv1v = c.inv1v;
    return t;
  }
  template <int G, int NC> struct Pre { cvec<NC> r[(MODE >= 2) ? G : 1]; unsigned mbits; };
  template <int G, int NC> B2S_HD void prefetch(const Ptr& t, Pre<G, NC>& pre) const {
    pre.mbits = 0xffffffffu;
    if (MODE >= 1) {
      pre.mbits = 0u;
#pragma unroll
      for (int k = 0; k < G; ++k) pre.mbits |= (t.m[8 * k] ? 1u : 0u) << k;
    }
    if (MODE >= 2) {
#pragma unroll
      for (int k = 0; k < G; ++k) {                        // DC needs ref on sampled rows only
        if (MODE == 2 && H == 256) {                       // defined on every path: a write under a predicate alone keeps the previous
#pragma unroll                                             // task's 2*G*NC registers alive across the whole persistent loop (measured: helps
          for (int n = 0; n < NC; ++n) pre.r[k].v[n] = make_c(0.f, 0.f);   // the 256 x 256 plan, costs 4 us at 200 x 200)
        }
        if (MODE == 3 || ((pre.mbits >> k) & 1u)) pre.r[k] = ldv_stream<NC>(t.r + 8 * k * W);
      }
    }
  }
  template <int G, int NC> B2S_HD void store(const Ptr& t, int k, const float* re_in, const float* im_in, const Pre<G, NC>& pre) const {
    const bool mk = (pre.mbits >> k) & 1u;
    cvec<NC> o;
    if (MODE <= 1) {
#pragma unroll
      for (int n = 0; n < NC; ++n) o.v[n] = mk ? make_c(re_in[n], im_in[n]) : make_c(0.f, 0.f);
    } else if (MODE == 2) {
      if (mk) {
        const cvec<NC> r = pre.r[k];
#pragma unroll
        for (int n = 0; n < NC; ++n)
          o.v[n] = make_c((re_in[n] + t.v * r.v[n].x) * t.inv1v, (im_in[n] + t.v * r.v[n].y) * t.inv1v);
      } else {
#pragma unroll
        for (int n = 0; n < NC; ++n) o.v[n] = make_c(re_in[n], im_in[n]);
      }
    } else {
      const cvec<NC> r = pre.r[k];
#pragma unroll
      for (int n = 0; n < NC; ++n)
        o.v[n] = mk ? make_c(re_in[n] - r.v[n].x, im_in[n] - r.v[n].y) : make_c(0.f - r.v[n].x, 0.f - r.v[n].y);
    }
    stv_stream<NC>(t.p + 8 * k * W, o);
  }
  // the reference k-space this item will blend with (whole image: the sibling half needs the rest)
  // DC (MODE 2) blends on sampled rows only: warm L2 with just those rows of this work item (ky = q mod
  // fold), W*8 bytes each; MODE 3 reads every row.
  // whole-image kernels: the NEXT image's sampled reference rows (MODE 2), see prefetch_sampled_rows
  B2S_HD void l2_prefetch_ahead(long long image, int tid, int nt) const {
    if (MODE == 2) prefetch_sampled_rows<H, W>(ref + image * hw, mask + (image / C) * H, tid, nt);
    if (MODE == 3) prefetch_span(ref + image * hw, (long long)H * W * 8, tid);
  }
  B2S_HD void l2_prefetch(long long image, int q, int fold, int tid) const {
    if (MODE == 3) { if (q == 0) prefetch_span(ref + image * hw, (long long)H * W * 8, tid); }
    if (MODE == 2) {
      const int y = fold * tid + q;
      if (y < H && mask[(image / C) * H + y]) l2_prefetch_bulk(ref + image * hw + (long long)y * W, W * 8);
    }
  }
};

#ifndef B2S_FIX_U
#define B2S_FIX_U 10
#endif
// Soft-DC blend (varnet.py:281-282) as a ROW FIX-UP instead of a predicated epilogue.  Phase C stores the plain
// transform (no mask look-ups, no predicated reference loads, 50 registers fewer); once every warp of the CTA has
// left that Phase C (the kernel calls `fixup` behind the next work item's Phase A barrier) the CTA re-reads only the
// SAMPLED rows of what it just wrote (still in L2) together with the same rows of the reference k-space, blends them
// and writes them back: 3 x 25 % of the image in fully coalesced 128-bit accesses with ten loads in flight per
// thread, instead of 25 predicated 64-bit loads per Phase C task that sit in the dependency chain of its stores.
// `stage_rows` lists, per work item, the sampled rows of its residue class (ky == q mod FOLD).
template <int H, int W> struct EpiDCFix {
  static constexpr bool FIXUP = true;
  static constexpr int V4 = W / 2;                       // 128-bit accesses per row
  cfloat* out; const cfloat* ref; const uint8_t* mask; const float* vptr; int C; long long hw; int pf;
  struct Ctx { cfloat* p; };
  typedef cfloat* Ptr;
  B2S_HD Ctx ctx(long long image, const uint8_t*) const { Ctx c; c.p = out + image * hw; return c; }
  // aux layout: [0, H) the image's mask row; two row lists (slots) of up to H/2 sampled rows each (or ONE list of up to
  // H rows in slot 0); [CNT + 4*slot] their counts
  static constexpr int LIST = (H + 15) / 16 * 16, HALF = LIST / 2, CNT = 2 * LIST, AUX_BYTES = 2 * LIST + 16;
  // the item's mask row into aux[0, H): called by every thread BEFORE Phase A (the load latency hides behind it)
  B2S_HD void stage_mask_row(long long image, uint8_t* aux, int tid, int nt) const {
    const uint8_t* src = mask + (image / C) * H;
    for (int y = tid; y < H; y += nt) aux[y] = src[y];
  }
  // compaction of the staged row by ONE warp (lanes tid0 .. tid0 + 31) after a barrier: rows q, q + FOLD, ... -> list `slot`
  template <int FOLD> B2S_HD void stage_rows(int q, uint8_t* aux, int tid, int tid0, int slot = 0) const {
    constexpr int NR = H / FOLD, IT = (NR + 31) / 32;
    uint8_t* list = aux + LIST + slot * HALF;
#if defined(__CUDA_ARCH__)
    if (tid >= tid0 && tid < tid0 + 32) {
      const int lane = tid - tid0;
      int n = 0;
#pragma unroll
      for (int i = 0; i < IT; ++i) {
        const int r = 32 * i + lane;
        const bool on = (r < NR) && aux[q + FOLD * r];
        const unsigned bits = __ballot_sync(0xffffffffu, on);
        if (on) list[n + __popc(bits & ((1u << lane) - 1u))] = (uint8_t)(q + FOLD * r);
        n += __popc(bits);
      }
      if (lane == 0) reinterpret_cast<int*>(aux + CNT)[slot] = n;
    }
#else
    if (tid == tid0) {
      int n = 0;
      for (int y = q; y < H; y += FOLD) if (aux[y]) list[n++] = (uint8_t)y;
      reinterpret_cast<int*>(aux + CNT)[slot] = n;
    }
    (void)IT;
#endif
  }
  B2S_HD void stage_mask(long long, uint8_t*, int, int) const {}
  B2S_HD Ptr task_ptr(const Ctx& c, int m, int kx) const { return c.p + m * W + kx; }
  template <int G, int NC> struct Pre {};
  template <int G, int NC> B2S_HD void prefetch(Ptr, Pre<G, NC>&) const {}
  template <int G, int NC> B2S_HD void store(Ptr p, int k, const float* re, const float* im, const Pre<G, NC>&) const {
    cvec<NC> v;
#pragma unroll
    for (int n = 0; n < NC; ++n) v.v[n] = make_c(re[n], im[n]);
    stv<NC>(p + 8 * k * W, v);
  }
  // blend the sampled rows listed in aux (written by stage_rows for this image/q)
  // (not inlined on the device: its 80-100 registers of loads in flight then do not weigh on the allocation of the
  // persistent loop it is called from)
#if defined(__CUDACC__)
  template <int NT> __host__ __device__ __noinline__ void fixup(long long image, const uint8_t* aux, int tid, int slot = 0) const {
#else
  template <int NT> void fixup(long long image, const uint8_t* aux, int tid, int slot = 0) const {
#endif
    constexpr int U = B2S_FIX_U;                          // 128-bit load PAIRS in flight per thread (the phase owns the whole
                                                          // register file: nothing else is live between Phases A and B)
    if (pf & 16) return;                                  // dev: timing without the fix-up
    const int total = reinterpret_cast<const int*>(aux + CNT)[slot] * V4;
    const uint8_t* list = aux + LIST + slot * HALF;
    cvec<2>* o = reinterpret_cast<cvec<2>*>(out + image * hw);
    const cvec<2>* r = reinterpret_cast<const cvec<2>*>(ref + image * hw);
    const float v = *vptr, inv1v = 1.f / (1.f + v);
    for (int e0 = tid; e0 < total; e0 += NT * U) {
      cvec<2> a[U], b[U]; int off[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int e = e0 + u * NT;
        off[u] = -1;
        if (e < total) {
          const int i = e / V4;
          off[u] = (int)list[i] * V4 + (e - i * V4);
          a[u] = ldv<2>(reinterpret_cast<const cfloat*>(o + off[u]));
          b[u] = (pf & 32) ? a[u] : ldv_stream<2>(reinterpret_cast<const cfloat*>(r + off[u]));
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (off[u] >= 0) {
          cvec<2> z;
#pragma unroll
          for (int n = 0; n < 2; ++n)
            z.v[n] = make_c((a[u].v[n].x + v * b[u].v[n].x) * inv1v, (a[u].v[n].y + v * b[u].v[n].y) * inv1v);
          if (!(pf & 64) || z.v[0].x == 123.456f) stv_stream<2>(reinterpret_cast<cfloat*>(o + off[u]), z);
        }
      }
    }
  }
  B2S_HD void l2_prefetch_ahead(long long image, int tid, int nt) const {
    if ((pf & 15) == 3) prefetch_sampled_rows<H, W>(ref + image * hw, mask + (image / C) * H, tid, nt);
  }
  B2S_HD void l2_prefetch(long long image, int q, int fold, int tid) const {
    if ((pf & 15) == 1) {
      const int y = fold * tid + q;
      if (y < H && mask[(image / C) * H + y]) l2_prefetch_bulk(ref + image * hw + (long long)y * W, W * 8);
    }
#if defined(__CUDA_ARCH__)
    if ((pf & 15) == 2) {                                        // per-thread 128-byte line prefetches: W*8/128 lines per row
      constexpr int LPR = (W * 8 + 127) / 128;
      for (int e = tid; e < (H / fold) * LPR; e += 256) {
        const int r = e / LPR, l = e - r * LPR, y = fold * r + q;
        if (mask[(image / C) * H + y]) asm volatile("prefetch.global.L2 [%0];" ::"l"((const char*)(ref + image * hw + (long long)y * W) + 128 * l));
      }
    }
#endif
  }
};

// out[(b,t,c) . ostride] += conj(mult[(b,t,c) . mstride]) * ifft(k);  zero stride = reduced dim
template <int H, int W> struct EpiReduce {
  static constexpr bool FIXUP = false;
  cfloat* out; const cfloat* mult; int T, C;
  long long os_b, os_t, os_c, ms_b, ms_t, ms_c;
  struct Ctx { cfloat* o; const cfloat* m; };
  struct Ptr { cfloat* o; const cfloat* m; };
  B2S_HD Ctx ctx(long long image, const uint8_t*) const {
    const long long c = image % C, bt = image / C, b = bt / T, t = bt % T;
    Ctx k; k.o = out + b * os_b + t * os_t + c * os_c; k.m = mult + b * ms_b + t * ms_t + c * ms_c;
    return k;
  }
  B2S_HD void stage_mask(long long, uint8_t*, int, int) const {}
  B2S_HD Ptr task_ptr(const Ctx& c, int m, int kx) const { Ptr t; t.o = c.o + m * W + kx; t.m = c.m + m * W + kx; return t; }
  template <int G, int NC> struct Pre { cvec<NC> s[G]; };
  template <int G, int NC> B2S_HD void prefetch(const Ptr& t, Pre<G, NC>& pre) const {
#pragma unroll
    for (int k = 0; k < G; ++k) pre.s[k] = ldv_keep<NC>(t.m + 8 * k * W);
  }
  template <int G, int NC> B2S_HD void store(const Ptr& t, int k, const float* re, const float* im, const Pre<G, NC>& pre) const {
    const cvec<NC> s = pre.s[k];
    float ar[NC], ai[NC];
#pragma unroll
    for (int n = 0; n < NC; ++n) {
      const float yr = im[n], yi = re[n];                  // swap back (inverse transform)
      ar[n] = yr * s.v[n].x + yi * s.v[n].y;
      ai[n] = yi * s.v[n].x - yr * s.v[n].y;
    }
    red_add<NC>(t.o + 8 * k * W, ar, ai);
  }
  B2S_HD void l2_prefetch(long long, int, int, int) const {}
  B2S_HD void l2_prefetch_ahead(long long, int, int) const {}
};

}  // namespace b2s
