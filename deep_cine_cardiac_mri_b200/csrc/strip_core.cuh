// Strip-streamed fused centred 2-D FFT ("two passes through L2, one pass through HBM").
//
// The half-split kernel (fft2_core.cuh) keeps one image per SM on chip and is bound by what a single resident
// CTA can ingest from L2.  This kernel gives up on-chip residency of the *image* and keeps only the *intermediate*
// on chip - in the 126 MB L2 instead of shared memory:
//
//   pass R  sub-unit = 4 consecutive rows of one image (4*W*8 contiguous bytes).  W-point transform per row in two
//           steps (W = N1*N2): N1-point register codelet over stride-N2 elements, twiddle (table also carries the
//           centring signs and the scale), exchange through the warp's private shared-memory buffer, then the N2 =
//           A*B points of one (row, k1) belong to ONE lane which runs A-point codelets, constant twiddles and
//           B-point codelets in place.  The prologue functor supplies the element (identity | S_c * x_t |
//           row-weighted k-space); results go to a scratch ring in global memory whose layout is strip-major, so
//           pass C reads contiguous blocks.
//   pass C  sub-unit = a strip of 4 columns of one image.  H-point transform per column, same two steps; the epilogue
//           functor consumes the result (store | mask / soft-DC blend / residual | conj(S)-multiply + coil sum).
//
// One persistent kernel runs both passes.  WARPS are the unit of execution: after the table build there is no CTA
// barrier; each warp draws tickets (one sub-unit each) from an atomic counter, issues the global loads of its next
// sub-unit into registers before it runs the second step of the current one, and signals completion with
// red.release.  The ticket order puts pass C of image p - LAG behind pass R of image p, so the scratch ring (NSLOT
// images, 80 x 320 KB = 26 MB at 200 x 200) is written and re-read while still L2-resident and the k-space crosses HBM
// once per direction.  Dependencies are per-image counters (release/acquire): pass C of image i waits for its H/4
// pass-R sub-units; pass R of image i waits until pass C of image i - NSLOT (the previous owner of the ring slot) is
// finished.  Tickets are handed out in order and a sub-unit only ever waits for smaller tickets, i.e. for sub-units
// that are already running: no deadlock; waits are bounded and flag `status` instead of hanging the GPU.
//
// Centring (utils/fftc.py:59-110) for even sizes: fftc(x)[k] = (-1)^(N/2) (-1)^k DFT((-1)^n x[n])[k].
// With n = N2*n1 + n2 and k = k1 + N1*k2 (N1 even): (-1)^n1 (N2 odd) is a rotation of the N1-point
// codelet's outputs by N1/2, (-1)^(n2 + k1) goes into the twiddle table, which also carries the scale.
// The inverse transform runs the forward machinery on re/im-swapped data.
//
// Measured on B200 (profiles/r1_strip_experiment.md): the memory side works (DRAM traffic = algorithmic bytes), but
// the formulation spends 20-60 % more instructions than the on-chip kernels and loses to them; it is selectable
// (b2s_set_fused_path / B2S_PATH=strip) and covered by the same parity tests, not the default.
#pragma once
#include <stdint.h>
#include "codelets.cuh"
#include "fft2_core.cuh"       // cfloat, twiddle()
#include "strip_twiddles.cuh"
#include "sense_functors.cuh"  // ldv_pol / stv_pol / l2_keep / l2_stream / red_add / l2_prefetch_bulk

namespace b2s {

constexpr int STRIP_NV = 4;    // vectors (rows in pass R, columns in pass C) per warp sub-unit
constexpr int STRIP_TV = 4;    // vectors per ticket (= one warp sub-unit)
constexpr int STRIP_WARPS = 4; // warps per CTA (they only share the twiddle tables)

// One dimension's plan: N = N1*N2 points, N2 = A*B for the in-thread second step; shared-memory pitches of the exchange buffer X[v][k1][n2] for the row-pass / column-pass lane orders
// (brute-forced conflict-free for 64-bit accesses, tools/strip_banks.py).
template <int N_> struct StripDim;
template <> struct StripDim<200> {
  static constexpr int N = 200, N1 = 8, N2 = 25, A = 5, B = 5;
  static constexpr int PK_R = 26, PV_R = 217, PK_C = 25, PV_C = 204;
};
template <> struct StripDim<256> {
  static constexpr int N = 256, N1 = 16, N2 = 16, A = 4, B = 4;
  static constexpr int PK_R = 17, PV_R = 272, PK_C = 17, PV_C = 276;
};

struct StripArgs {
  cfloat* scratch;        // ring: NSLOT images, strip-major [slot][strip][row][NV]
  int* head;              // ticket counter
  int* status;            // != 0: a dependency wait timed out (bug indicator; results invalid)
  int* done_r;            // [n_images] pass-R units finished per image
  int* done_c;            // [n_images] pass-C units finished per image
  int n_images, lag, nslot;
  float scale;
};

#if defined(__CUDACC__)

#if !defined(__CUDA_ARCH__)
// host compilation pass of nvcc: the device helpers of sense_functors.cuh do not exist; these stubs only
// make the __global__ template parse (they are never called on the host)
template <int NC> inline cvec<NC> ldv_pol(const cfloat* p, unsigned long long) { return *reinterpret_cast<const cvec<NC>*>(p); }
template <int NC> inline void stv_pol(cfloat* p, const cvec<NC>& v, unsigned long long) { *reinterpret_cast<cvec<NC>*>(p) = v; }
inline unsigned long long l2_keep() { return 0; }
inline unsigned long long l2_stream() { return 0; }
#endif

// scratch reads: L2 only (.cg) - a ring slot is rewritten by other SMs during the launch, L1 must never serve it
__device__ __forceinline__ cvec<1> ld_scratch(const cfloat* p, unsigned long long pol) {
  cvec<1> r;
#if defined(__CUDA_ARCH__)
  asm volatile("ld.global.cg.L2::cache_hint.v2.f32 {%0, %1}, [%2], %3;" : "=f"(r.v[0].x), "=f"(r.v[0].y) : "l"(p), "l"(pol));
#else
  r = *reinterpret_cast<const cvec<1>*>(p); (void)pol;
#endif
  return r;
}

__device__ __forceinline__ int ld_acquire(const int* p) {
  int v = 0;
#if defined(__CUDA_ARCH__)
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
#endif
  return v;
}
__device__ __forceinline__ void red_release(int* p, int v) {
#if defined(__CUDA_ARCH__)
  asm volatile("red.release.gpu.global.add.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
#endif
}
// thread 0 spins until *p >= want (bounded: a timeout flags `status` instead of hanging the GPU)
__device__ __forceinline__ void wait_count(const int* p, int want, int* status) {
  unsigned spins = 0;
  while (ld_acquire(p) < want) {
    __nanosleep(64);
    if (++spins > (1u << 22)) { atomicExch(status, 1); break; }
  }
}

// ------------------------------------------------------------------------------------------------ //
// prologues (pass R input).  fetch(): raw global loads of element `off` (= row*W + col) of the image into
// RAWN register slots; value(): the complex number fed to the transform (re/im-swapped for inverse transforms).
// ------------------------------------------------------------------------------------------------ //
template <bool INV> struct SProPlain {
  const cfloat* in; long long image_stride;
  struct Ctx { const cfloat* p; };
  static constexpr int RAWN = 1;
  static constexpr bool STREAM_IN = true;
  __device__ __forceinline__ Ctx ctx(long long image) const { Ctx c; c.p = in + image * image_stride; return c; }
  __device__ __forceinline__ void fetch(const Ctx& c, int off, cfloat* raw, unsigned long long pol) const { raw[0] = ldv_pol<1>(c.p + off, pol).v[0]; }
  __device__ __forceinline__ float row_weight(const Ctx&, int) const { return 1.f; }
  __device__ __forceinline__ void value(const cfloat* raw, float, float& re, float& im) const {
    re = INV ? raw[0].y : raw[0].x; im = INV ? raw[0].x : raw[0].y;
  }
};

// S_c * x_t  (models/varnet.py:181-185)
struct SProExpand {
  const cfloat* img; const cfloat* sens; int T, C; long long hw;
  struct Ctx { const cfloat* a; const cfloat* s; };
  static constexpr int RAWN = 2;
  static constexpr bool STREAM_IN = false;     // image and maps are re-read: keep them in L2
  __device__ __forceinline__ Ctx ctx(long long image) const {
    const long long c = image % C, bt = image / C, b = bt / T;
    Ctx k; k.a = img + bt * hw; k.s = sens + (b * C + c) * hw; return k;
  }
  __device__ __forceinline__ void fetch(const Ctx& c, int off, cfloat* raw, unsigned long long pol) const {
    raw[0] = ldv_pol<1>(c.a + off, pol).v[0]; raw[1] = ldv_pol<1>(c.s + off, pol).v[0];
  }
  __device__ __forceinline__ float row_weight(const Ctx&, int) const { return 1.f; }
  __device__ __forceinline__ void value(const cfloat* raw, float, float& re, float& im) const {
    const cfloat a = raw[0], s = raw[1];
    re = a.x * s.x - a.y * s.y;
    im = a.x * s.y + a.y * s.x;
  }
};

// w(ky) * k with w = 1 | mask | 1 - eta*mask (sens_reduce, masked BackwardOperator, DC backward); inverse
template <int WMODE> struct SProKspace {
  const cfloat* k; const uint8_t* mask; const float* vptr; int C, H; long long hw;
  struct Ctx { const cfloat* p; const uint8_t* m; float wa, wb; };
  static constexpr int RAWN = 1;
  static constexpr bool STREAM_IN = true;
  __device__ __forceinline__ Ctx ctx(long long image) const {
    Ctx c; c.p = k + image * hw; c.m = WMODE ? mask + (image / C) * H : nullptr;
    c.wa = 1.f; c.wb = 0.f;
    if (WMODE == 1) { c.wa = 0.f; c.wb = 1.f; }
    if (WMODE == 2) { const float v = *vptr; c.wb = -v / (1.f + v); }
    return c;
  }
  __device__ __forceinline__ void fetch(const Ctx& c, int off, cfloat* raw, unsigned long long pol) const { raw[0] = ldv_pol<1>(c.p + off, pol).v[0]; }
  __device__ __forceinline__ float row_weight(const Ctx& c, int row) const { return WMODE ? c.wa + c.wb * (float)c.m[row] : 1.f; }
  __device__ __forceinline__ void value(const cfloat* raw, float w, float& re, float& im) const {
    if (WMODE) { re = raw[0].y * w; im = raw[0].x * w; } else { re = raw[0].y; im = raw[0].x; }
  }
};

// ------------------------------------------------------------------------------------------------ //
// epilogues (pass C output).  A lane finishes NOUT results of column `col` at a time, rows
// row0 + RS*j (j = 0..NOUT-1): pre() issues the loads the group needs (one group ahead of the codelet that
// produces it), fin() combines and stores.
// ------------------------------------------------------------------------------------------------ //
template <bool INV> struct SEpiPlain {
  cfloat* out; long long image_stride; int W;
  struct Ctx { cfloat* p; };
  template <int NOUT> struct Pre {};
  __device__ __forceinline__ Ctx ctx(long long image) const { Ctx c; c.p = out + image * image_stride; return c; }
  __device__ __forceinline__ void warm(long long, int, int, int) const {}
  template <int RS, int NOUT> __device__ __forceinline__ void pre(const Ctx&, int, int, Pre<NOUT>&) const {}
  template <int RS, int NOUT>
  __device__ __forceinline__ void fin(const Ctx& c, int row0, int col, const float (&re)[NOUT], const float (&im)[NOUT], const Pre<NOUT>&) const {
    const unsigned long long pol = l2_stream();
    cfloat* p = c.p + (long long)row0 * W + col;
#pragma unroll
    for (int j = 0; j < NOUT; ++j) {
      cvec<1> o; o.v[0] = INV ? make_c(im[j], re[j]) : make_c(re[j], im[j]);
      stv_pol<1>(p + (long long)j * RS * W, o, pol);
    }
  }
};

// MODE 0: k ; 1: k*m + 0.0 ; 2: (1-m) k + m (k + v ref)/(1+v) ; 3: k*m - ref
// (cinenet.py:129, varnet.py:281-282, xpdnet.py:295-298)
template <int MODE> struct SEpiKspace {
  cfloat* out; const cfloat* ref; const uint8_t* mask; const float* vptr; int C, H, W; long long hw;
  struct Ctx { cfloat* p; const cfloat* r; const uint8_t* m; float v, inv1v; };
  template <int NOUT> struct Pre { cfloat r[(MODE >= 2) ? NOUT : 1]; unsigned mbits; };
  __device__ __forceinline__ Ctx ctx(long long image) const {
    Ctx c; c.p = out + image * hw;
    c.r = (MODE >= 2) ? ref + image * hw : nullptr;
    c.m = (MODE >= 1) ? mask + (image / C) * H : nullptr;
    c.v = (MODE == 2) ? *vptr : 0.f;
    c.inv1v = 1.f / (1.f + c.v);
    return c;
  }
  // called by the pass-R sub-unit that covers rows [row0, row0 + n) of `image`: warm L2 with the reference rows
  // pass C will blend with (whole rows: full DRAM bursts instead of the 32-byte pieces pass C reads)
  __device__ __forceinline__ void warm(long long image, int row0, int n, int lane) const {
    if (MODE >= 2 && lane < n) {
      const int y = row0 + lane;
      if (MODE == 3 || mask[(image / C) * H + y]) l2_prefetch_bulk(ref + image * hw + (long long)y * W, (unsigned)W * 8u);
    }
  }
  template <int RS, int NOUT> __device__ __forceinline__ void pre(const Ctx& c, int row0, int col, Pre<NOUT>& q) const {
    q.mbits = 0xffffffffu;
    if (MODE >= 1) {
      q.mbits = 0u;
#pragma unroll
      for (int j = 0; j < NOUT; ++j) q.mbits |= (c.m[row0 + RS * j] ? 1u : 0u) << j;
    }
    if (MODE >= 2) {
      const unsigned long long pol = l2_stream();
      const cfloat* p = c.r + (long long)row0 * W + col;
#pragma unroll
      for (int j = 0; j < NOUT; ++j)
        if (MODE == 3 || ((q.mbits >> j) & 1u)) q.r[j] = ldv_pol<1>(p + (long long)j * RS * W, pol).v[0];
    }
  }
  template <int RS, int NOUT>
  __device__ __forceinline__ void fin(const Ctx& c, int row0, int col, const float (&re)[NOUT], const float (&im)[NOUT], const Pre<NOUT>& q) const {
    const unsigned long long pol = l2_stream();
    cfloat* p = c.p + (long long)row0 * W + col;
#pragma unroll
    for (int j = 0; j < NOUT; ++j) {
      const bool mk = (q.mbits >> j) & 1u;
      cvec<1> o;
      if (MODE <= 1) o.v[0] = mk ? make_c(re[j], im[j]) : make_c(0.f, 0.f);
      else if (MODE == 2) o.v[0] = mk ? make_c((re[j] + c.v * q.r[j].x) * c.inv1v, (im[j] + c.v * q.r[j].y) * c.inv1v) : make_c(re[j], im[j]);
      else o.v[0] = mk ? make_c(re[j] - q.r[j].x, im[j] - q.r[j].y) : make_c(0.f - q.r[j].x, 0.f - q.r[j].y);
      stv_pol<1>(p + (long long)j * RS * W, o, pol);
    }
  }
};

// out[(b,t,c) . ostride] += conj(mult[(b,t,c) . mstride]) * y ; zero stride = reduced dim (varnet.py:192-194)
struct SEpiReduce {
  cfloat* out; const cfloat* mult; int T, C, W;
  long long os_b, os_t, os_c, ms_b, ms_t, ms_c;
  struct Ctx { cfloat* o; const cfloat* m; };
  template <int NOUT> struct Pre { cfloat s[NOUT]; };
  __device__ __forceinline__ Ctx ctx(long long image) const {
    const long long c = image % C, bt = image / C, b = bt / T, t = bt % T;
    Ctx k; k.o = out + b * os_b + t * os_t + c * os_c; k.m = mult + b * ms_b + t * ms_t + c * ms_c;
    return k;
  }
  __device__ __forceinline__ void warm(long long, int, int, int) const {}
  template <int RS, int NOUT> __device__ __forceinline__ void pre(const Ctx& c, int row0, int col, Pre<NOUT>& q) const {
    const unsigned long long pol = l2_keep();
    const cfloat* p = c.m + (long long)row0 * W + col;
#pragma unroll
    for (int j = 0; j < NOUT; ++j) q.s[j] = ldv_pol<1>(p + (long long)j * RS * W, pol).v[0];
  }
  template <int RS, int NOUT>
  __device__ __forceinline__ void fin(const Ctx& c, int row0, int col, const float (&re)[NOUT], const float (&im)[NOUT], const Pre<NOUT>& q) const {
    cfloat* p = c.o + (long long)row0 * W + col;
#pragma unroll
    for (int j = 0; j < NOUT; ++j) {
      const float yr = im[j], yi = re[j];                  // swap back (inverse transform)
      float ar[1], ai[1];
      ar[0] = yr * q.s[j].x + yi * q.s[j].y;
      ai[0] = yi * q.s[j].x - yr * q.s[j].y;
      red_add<1>(p + (long long)j * RS * W, ar, ai);
    }
  }
};

// ------------------------------------------------------------------------------------------------ //
// tables: T[k1][n2] = scale * (-1)^(n2 + k1) * exp(-2 pi i n2 k1 / N)
// ------------------------------------------------------------------------------------------------ //
template <class D> __device__ __forceinline__ void strip_build_table(cfloat* t, float scale, int tid, int nt) {
  for (int e = tid; e < D::N; e += nt) {
    const int k1 = e / D::N2, n2 = e - k1 * D::N2;
    cfloat w = twiddle(n2 * k1, D::N);
    const float s = ((n2 + k1) & 1) ? -scale : scale;
    t[e] = make_c(w.x * s, w.y * s);
  }
}

// A warp's current work: one sub-unit = STRIP_NV vectors.  kind 0: pass R rows, 1: pass C columns, -1: none.
struct WUnit { int kind, image, blk; };

// Step 1 of a sub-unit, split in two so that the global loads of the NEXT sub-unit are in flight while the
// warp runs step 2 of the current one (registers are the landing zone: PF complex slots per lane).
//   ROWPASS : element (v, n) of the sub-unit comes from the prologue at image offset (row0 + v)*W + n; lanes run
//             along n2 (global contiguity).  !ROWPASS: from the scratch strip at n*NV + v; lanes run along v.
//   PRE rounds of tasks are resident in `pf`; later rounds are loaded into the slots of consumed ones.
template <class D, bool ROWPASS, class Pro, int PF>
struct StripStep1 {
  static constexpr int N1 = D::N1, N2 = D::N2, NV = STRIP_NV;
  static constexpr int TASKS = NV * N2, ROUNDS = (TASKS + 31) / 32;
  static constexpr int RAWN = ROWPASS ? Pro::RAWN : 1;
  static constexpr int PRE0 = PF / (N1 * RAWN), PRE = PRE0 > ROUNDS ? ROUNDS : PRE0;
  static constexpr int PK = ROWPASS ? D::PK_R : D::PK_C, PV = ROWPASS ? D::PV_R : D::PV_C;
  static constexpr int SH = (N2 & 1) ? N1 / 2 : 0;
  static_assert(PRE >= 1, "prefetch buffer too small for one round");

  static __device__ __forceinline__ void decode(int task, int& v, int& n2) {
    if (ROWPASS) { v = task / N2; n2 = task - v * N2; } else { n2 = task / NV; v = task - n2 * NV; }
  }
  static __device__ __forceinline__ void load_round(const Pro& pro, const typename Pro::Ctx& ctx, const cfloat* src, int row0, int W,
                                                    int lane, int rr, cfloat* slot, unsigned long long pol) {
    const int task = lane + 32 * rr;
    if (task < TASKS) {
      int v, n2; decode(task, v, n2);
      if (ROWPASS) {
        const int off = (row0 + v) * W + n2;
#pragma unroll
        for (int n1 = 0; n1 < N1; ++n1) pro.fetch(ctx, off + N2 * n1, slot + n1 * RAWN, pol);
      } else {
        const cfloat* p = src + n2 * NV + v;
#pragma unroll
        for (int n1 = 0; n1 < N1; ++n1) slot[n1] = ld_scratch(p + N2 * n1 * NV, pol).v[0];
      }
    }
  }
  static __device__ __forceinline__ unsigned long long policy() {
    return ROWPASS ? (Pro::STREAM_IN ? l2_stream() : l2_keep()) : l2_keep();
  }
  // rounds [0, PRE)
  static __device__ __forceinline__ void issue(const Pro& pro, const typename Pro::Ctx& ctx, const cfloat* src, int row0, int W,
                                               int lane, cfloat (&pf)[PF]) {
    const unsigned long long pol = policy();
#pragma unroll
    for (int rr = 0; rr < PRE; ++rr) load_round(pro, ctx, src, row0, W, lane, rr, pf + rr * N1 * RAWN, pol);
  }
  // consume every round (loading rounds >= PRE on the way), write X
  static __device__ __forceinline__ void run(const Pro& pro, const typename Pro::Ctx& ctx, const cfloat* src, int row0, int W,
                                             int lane, cfloat (&pf)[PF], cfloat* X, const cfloat* T) {
    const unsigned long long pol = policy();
#pragma unroll
    for (int rr = 0; rr < ROUNDS; ++rr) {
      cfloat* slot = pf + (rr % PRE) * N1 * RAWN;
      const int task = lane + 32 * rr;
      if (task < TASKS) {
        int v, n2; decode(task, v, n2);
        float re[N1], im[N1];
        if (ROWPASS) {
          const float w = pro.row_weight(ctx, row0 + v);
#pragma unroll
          for (int n1 = 0; n1 < N1; ++n1) pro.value(slot + n1 * RAWN, w, re[n1], im[n1]);
        } else {
#pragma unroll
          for (int n1 = 0; n1 < N1; ++n1) { re[n1] = slot[n1].x; im[n1] = slot[n1].y; }
        }
        if (rr + PRE < ROUNDS) load_round(pro, ctx, src, row0, W, lane, rr + PRE, slot, pol);
        Dft<N1>::run(re, im);
        cfloat* dst = X + v * PV + n2;
        const cfloat* tw = T + n2;
#pragma unroll
        for (int k1 = 0; k1 < N1; ++k1) {
          const int s = (k1 + SH) % N1;
          const cfloat t = tw[k1 * N2];
          dst[k1 * PK] = make_c(re[s] * t.x - im[s] * t.y, re[s] * t.y + im[s] * t.x);
        }
      } else if (rr + PRE < ROUNDS) {
        load_round(pro, ctx, src, row0, W, lane, rr + PRE, slot, pol);
      }
    }
  }
};

// First half of step 2: the N2 = A*B points of one (vector, k1) row of X belong to ONE lane, which runs the
// A-point codelets and the constant second-level twiddles in place in its row (no synchronisation needed).
template <class D>
__device__ __forceinline__ void strip_step2a(cfloat* row) {
  constexpr int A = D::A, B = D::B, N2 = D::N2;
#pragma unroll
  for (int b = 0; b < B; ++b) {
    float re[A], im[A];
#pragma unroll
    for (int j = 0; j < A; ++j) { const cfloat c = row[B * j + b]; re[j] = c.x; im[j] = c.y; }
    Dft<A>::run(re, im);
#pragma unroll
    for (int ka = 0; ka < A; ++ka) {
      const int e = (b * ka) % N2;
      float r = re[ka], i = im[ka];
      if (e != 0) { const float2 t = StripTw<N2>::get(e); const float rr = r * t.x - i * t.y; i = r * t.y + i * t.x; r = rr; }
      row[B * ka + b] = make_c(r, i);
    }
  }
}
// second half: B-point codelet of group ka; outputs k2 = ka + A*kb, kb = 0..B-1
template <class D>
__device__ __forceinline__ void strip_step2b(const cfloat* row, int ka, float (&re)[D::B], float (&im)[D::B]) {
#pragma unroll
  for (int b = 0; b < D::B; ++b) { const cfloat c = row[D::B * ka + b]; re[b] = c.x; im[b] = c.y; }
  Dft<D::B>::run(re, im);
}

// ------------------------------------------------------------------------------------------------ //
// the kernel
// ------------------------------------------------------------------------------------------------ //
// DW / DH: StripDim of the row pass (W points) and of the column pass (H points).
template <class DW, class DH, class Pro, class Epi, int MINB>
__global__ void __launch_bounds__(32 * STRIP_WARPS, MINB)
strip_fft2_kernel(const Pro pro, const Epi epi, const StripArgs a) {
  constexpr int NV = STRIP_NV, TV = STRIP_TV, W = DW::N, H = DH::N, PF = 32;
  static_assert(DW::N1 % 2 == 0 && DH::N1 % 2 == 0, "N1 must be even (centring folded into the tables)");
  static_assert(W % TV == 0 && H % TV == 0 && TV == NV, "bad plan");
  static_assert(DW::N1 % NV == 0, "pass-R output column k1 + N1*k2 -> strip k1/NV + (N1/NV)*k2, lane column k1 % NV");
  constexpr int RB = H / TV;          // pass-R tickets per image
  constexpr int SB = W / TV;          // pass-C tickets per image
  constexpr int UPP = RB + SB;        // tickets per position
  constexpr int XR = NV * DW::PV_R, XC = NV * DH::PV_C, XE = XR > XC ? XR : XC;
  typedef StripStep1<DW, true, Pro, PF> S1R;
  typedef StripStep1<DH, false, Pro, PF> S1C;

  __shared__ __align__(16) cfloat Xs[STRIP_WARPS][XE];
  __shared__ __align__(16) cfloat TW[W];
  __shared__ __align__(16) cfloat TH[H];
  const int tid = threadIdx.x, lane = tid & 31;
  cfloat* X = Xs[tid >> 5];

  strip_build_table<DW>(TW, 1.f, tid, 32 * STRIP_WARPS);
  strip_build_table<DH>(TH, a.scale, tid, 32 * STRIP_WARPS);
  __syncthreads();                    // the only CTA-wide barrier: from here on every warp is on its own

  const long long total = (long long)(a.n_images + a.lag) * UPP;
  const long long img_elems = (long long)H * W;
  int my_next = 0;                    // lane 0: the ticket this warp takes next (fetched one ticket ahead)
  if (lane == 0) my_next = atomicAdd(a.head, 1);

  // lane 0: block until *p >= want (bounded - a timeout flags `status` instead of hanging the GPU)
  auto wait_for = [&](const int* p, int want) { wait_count(p, want, a.status); };
  // next valid ticket
  auto advance = [&]() -> WUnit {
    WUnit n;
    for (;;) {
      const int u = __shfl_sync(0xffffffffu, my_next, 0);
      if (u >= total) { n.kind = -1; return n; }
      if (lane == 0) my_next = atomicAdd(a.head, 1);
      const int p = u / UPP, r = u - p * UPP;
      if (r < RB) { n.kind = 0; n.image = p; n.blk = r; }
      else { n.kind = 1; n.image = p - a.lag; n.blk = r - RB; }
      if (n.image >= 0 && n.image < a.n_images) return n;
    }
  };
  auto scratch_in = [&](const WUnit& c) -> const cfloat* {     // the NV-column strip a pass-C sub-unit reads
    return a.scratch + (long long)(c.image % a.nslot) * img_elems + (long long)c.blk * H * NV;
  };
  // loads of sub-unit `c` into pf; pass C needs its image's pass R complete: returns false (nothing issued) if
  // `may_block` is false and it is not
  cfloat pf[PF];
  auto issue = [&](const WUnit& c, bool may_block) -> bool {
    if (c.kind == 0) {
      S1R::issue(pro, pro.ctx(c.image), nullptr, c.blk * NV, W, lane, pf);
      return true;
    }
    int ok = 1;
    if (lane == 0) { if (may_block) wait_for(a.done_r + c.image, RB); else ok = ld_acquire(a.done_r + c.image) >= RB; }
    ok = __shfl_sync(0xffffffffu, ok, 0);
    __syncwarp();
    if (!ok) return false;
    S1C::issue(pro, typename Pro::Ctx(), scratch_in(c), 0, W, lane, pf);
    return true;
  };

  WUnit cur = advance();
  if (cur.kind >= 0) issue(cur, true);

#pragma unroll 1
  while (cur.kind >= 0) {
    // ---------------- step 1 of `cur` (its loads are in flight or have landed) ----------------
    if (cur.kind == 0) {
      const int row0 = cur.blk * NV;
      epi.warm(cur.image, row0, NV, lane);
      S1R::run(pro, pro.ctx(cur.image), nullptr, row0, W, lane, pf, X, TW);
    } else {
      S1C::run(pro, typename Pro::Ctx(), scratch_in(cur), 0, W, lane, pf, X, TH);
    }
    __syncwarp();
    // ---------------- signals and the next sub-unit's loads ----------------
    if (lane == 0) {
      if (cur.kind == 1) red_release(a.done_c + cur.image, 1);  // scratch strip consumed
      // pass R: the ring slot must have been drained by pass C of its previous owner before we overwrite it
      if (cur.kind == 0 && cur.image >= a.nslot) wait_for(a.done_c + (cur.image - a.nslot), SB);
    }
    __syncwarp();
    const WUnit nxt = advance();
    const bool issued = nxt.kind >= 0 ? issue(nxt, false) : true;
    // ---------------- step 2 of `cur` ----------------
    if (cur.kind == 0) {
      constexpr int N1 = DW::N1, A = DW::A, B = DW::B, R2 = (NV * N1 + 31) / 32;
      const unsigned long long pol = l2_keep();
      cfloat* slot = a.scratch + (long long)(cur.image % a.nslot) * img_elems;
      const int row0 = cur.blk * NV;
#pragma unroll
      for (int r2 = 0; r2 < R2; ++r2) {
        const int task = lane + 32 * r2;
        const int v = task / N1, k1 = task - v * N1;            // lanes along k1: NV contiguous scratch elements
        cfloat* row = X + v * DW::PV_R + k1 * DW::PK_R;
        cfloat* dst = slot + ((long long)(k1 / NV) * H + (row0 + v)) * NV + (k1 % NV);
        strip_step2a<DW>(row);
#pragma unroll
        for (int ka = 0; ka < A; ++ka) {
          float re[B], im[B];
          strip_step2b<DW>(row, ka, re, im);
#pragma unroll
          for (int kb = 0; kb < B; ++kb) {
            cvec<1> o; o.v[0] = make_c(re[kb], im[kb]);
            stv_pol<1>(dst + (long long)(N1 / NV) * (ka + A * kb) * H * NV, o, pol);
          }
        }
      }
      __syncwarp();
      if (lane == 0) red_release(a.done_r + cur.image, 1);     // every lane's scratch stores precede the release
    } else {
      constexpr int N1 = DH::N1, A = DH::A, B = DH::B, R2 = (NV * N1 + 31) / 32;
      const typename Epi::Ctx ectx = epi.ctx(cur.image);
#pragma unroll
      for (int r2 = 0; r2 < R2; ++r2) {
        const int task = lane + 32 * r2;
        const int k1 = task / NV, v = task - k1 * NV;           // lanes along v: NV contiguous output columns
        cfloat* row = X + v * DH::PV_C + k1 * DH::PK_C;
        const int col = cur.blk * NV + v;
        typename Epi::template Pre<B> q0, q1;
        epi.template pre<N1 * A, B>(ectx, k1, col, q0);
        strip_step2a<DH>(row);
#pragma unroll
        for (int ka = 0; ka < A; ++ka) {
          float re[B], im[B];
          strip_step2b<DH>(row, ka, re, im);
          if (ka + 1 < A) epi.template pre<N1 * A, B>(ectx, k1 + N1 * (ka + 1), col, (ka & 1) ? q0 : q1);
          epi.template fin<N1 * A, B>(ectx, k1 + N1 * ka, col, re, im, (ka & 1) ? q1 : q0);
        }
      }
    }
    __syncwarp();                                               // X is free again
    if (!issued) issue(nxt, true);
    cur = nxt;
  }
}

#endif  // __CUDACC__

}  // namespace b2s
