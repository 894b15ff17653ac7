// CineNet normal operator  H x = A^H M A x + v x  (models/cinenet.py:121-133, recurrent_cinenet.py:74-86) and the
// image-domain VarNet cascade (varnet.py:253-282), k-space kept on chip - WARP-PRIVATE formulation.
//
// The mask selects k-space rows only (data/subsample.py:146-151), so F^H M F = (F_h^H M F_h) (x) I_w: no transform
// along w is needed, every image column is independent and the c * K bytes of intermediate k-space never exist:
//
//     H x = sum_c conj(S_c) . [ F_h^H M F_h (S_c . x) ]  +  v x .
//
// Decomposition of the work:
//
//   * one work item = 4 adjacent columns of one frame, owned by ONE WARP for all coils.  Lane (m, xl) = (lane / 4, lane % 4)
//     owns rows {m + 8 i} of column x0 + xl: its x values, its S_c values, its share of the coil sum and its radix-G
//     transforms live in its own registers / thread-private shared-memory slots.  A warp access touches 8 rows x 32
//     contiguous bytes = 8 full sectors; with a compile-time width every row address is base + immediate.
//   * H = 8 G.  Forward: radix-G over the lane's rows (register codelet), twiddle, length-8 transform over m giving bins
//     g + G k2; the mask is applied there and the inverse starts in the same registers (length-8 over {g + G j},
//     twiddle), then radix-G back to rows m + 8 k.  The length-8 step is the only inter-thread exchange and goes
//     through a warp-private buffer, so the coil loop runs on __syncwarp() alone - no CTA barrier: the resident warps
//     of an SM drift apart and cover each other's exchange and prefetch latencies.
//   * the exchange buffer holds QUADS {re_a, re_b, im_a, im_b} of two adjacent g (a = 2p, b = 2p + 1): every access is
//     128-bit, and the length-8 step - forward, mask, inverse, twiddle - transforms both g of a quad at once in packed
//     fp32 (FADD2 / FMUL2 / FFMA2), half the issue slots of the scalar form.  G odd: the last quad's b half is a dummy.
//   * centring signs (-1)^m and the 1/H scale are folded into the two twiddle tables; the mask row becomes a per-warp
//     table of packed 0/1 factors in the order the length-8 step reads them.
//   * S_c is loaded straight into registers at the top of a coil (a cp.async landing buffer one coil ahead measured
//     slower: its shared-memory wavefronts cost more than the latency the other resident warps cover anyway); CTAs are
//     small (3-4 warps, 2-4 per SM) so that prologues and epilogues of different CTAs overlap.
//
// Plans: H in {200 (G = 25), 256 (G = 32)}; W compile-time (200, 256) or 0 = run-time width (any multiple of 4).
// Emulated on the CPU by normal_warp_emulate (tests/host_emul).
#pragma once
#include "fft2_core.cuh"
#include "packed.cuh"

namespace b2s {

// mode 0: out = A^H M A x + v x                        (HOperator, cinenet.py:121-133)
// mode 1: out = ssq . x - eta (A^H M A x - bref)        (one VarNet cascade in the image domain:
//         A^H[ DC(A x, ref) ] with ssq = sum_c |S_c|^2, bref = A^H ref, eta = v/(1+v); varnet.py:253-282
//         followed by the next cascade's / the model's sens_reduce, varnet.py:150-151,253)
// mode 2: as mode 1, but the LAST cascade of an inference: writes |.| (complex_abs of the final sens_reduce, varnet.py:150-151)
//         as one float per pixel into `out` viewed as float (b,t,h,w)
struct NormalArgs {
  const cfloat* x; const cfloat* sens; const uint8_t* mask; const float* vptr; cfloat* out;
  int T, C, W;
  int mode; const float* ssq; const cfloat* bref;
  float* dot_part;     // mode 0, optional: dot_part[item] = sum over the item's pixels of Re<x, H x> (CG's <p, H p>, cinenet.py:159)
};

struct alignas(16) nquad { float ra, rb, ia, ib; };

template <int H_, int W_, int WARPS_, int CTAS_, int SPLIT_ = 1> struct NormalWarpPlan {
  static constexpr int H = H_, G = H_ / 8, XC = 4, WFIX = W_;
  static constexpr int WARPS = WARPS_, NT = 32 * WARPS_, CTAS = CTAS_;
  static constexpr int SPLIT = SPLIT_;                        // warps that share one work item, each taking 1/SPLIT of the coils (small launches:
  static constexpr int ITEMS = WARPS_ / SPLIT_;               // fills the GPU when b*t*w/4 items are fewer than the resident warps); items per CTA
  static_assert(WARPS_ % SPLIT_ == 0, "whole items per CTA");   // CTAS: resident CTAs per SM the registers are budgeted for
  static constexpr int NP = (G + 1) / 2;                      // quads (pairs of adjacent g) per (m, xl)
  static constexpr int EPQ = 8 * XC + 4;                      // quads per pair-block of E; 36 = 4 mod 8: step 2 conflict-free
  // per warp, in 8-byte units
  static constexpr int E_OFF = 0;                             // E[p][m][xl]   quads 0..31 of block p; the 4 padding quads hold
  static constexpr int MKQ = 8 * XC;                          // MK[p][k2] = {mask(2p + G k2), mask(2p+1 + G k2)} as floats, k2 = 0..7
  static constexpr int X_OFF = E_OFF + 2 * NP * EPQ;          // x_t   slots [i][lane]
  static constexpr int WARP_ELEMS = X_OFF + H_ * XC;
  // per CTA
  static constexpr int TWPQ = 9;                              // quads per pair-block of the twiddle tables (odd pitch)
  static constexpr int TW1_OFF = WARPS_ * WARP_ELEMS;         // TW1[p][m] = (-1)^m w^(m g)        {c_a, c_b, s_a, s_b}
  static constexpr int TW2_OFF = TW1_OFF + 2 * NP * TWPQ;     // TW2[p][m] = (-1)^m w^(m g) / H
  static constexpr int SMEM_ELEMS = TW2_OFF + 2 * NP * TWPQ;
  static constexpr int SMEM_BYTES = SMEM_ELEMS * 8;
  static constexpr int TASKS2 = NP * XC;                      // (pair, xl) packed length-8 transforms per coil and warp
  static constexpr int ROUNDS2 = (TASKS2 + 31) / 32;
  static_assert(H_ % 8 == 0 && (W_ % 4) == 0, "H = 8 G, 4-column groups");
};

template <class P> B2S_HD int nw_width(const NormalArgs& a) { return P::WFIX ? P::WFIX : a.W; }

// twiddle tables (once per CTA)
template <class P> B2S_HD void nw_build_tables(cfloat* smem, int tid, int nthreads) {
  nquad* t1 = reinterpret_cast<nquad*>(smem + P::TW1_OFF);
  nquad* t2 = reinterpret_cast<nquad*>(smem + P::TW2_OFF);
  const float scale = 1.f / (float)P::H;                      // ortho forward * ortho inverse along h
  for (int n = tid; n < 8 * P::NP; n += nthreads) {
    const int p = n >> 3, m = n & 7;
    const cfloat wa = twiddle(m * (2 * p), P::H);
    const cfloat wb = (2 * p + 1 < P::G) ? twiddle(m * (2 * p + 1), P::H) : make_c(0.f, 0.f);
    const float sg = (m & 1) ? -1.f : 1.f;
    nquad q; q.ra = wa.x * sg; q.rb = wb.x * sg; q.ia = wa.y * sg; q.ib = wb.y * sg;
    t1[p * P::TWPQ + m] = q;
    q.ra *= scale; q.rb *= scale; q.ia *= scale; q.ib *= scale;
    t2[p * P::TWPQ + m] = q;
  }
}

// once per item: this lane's rows of x_t and its share of the packed mask factors
template <class P>
B2S_HD void nw_stage(const NormalArgs& a, cfloat* ws, long long bt, int x0, int lane) {
  const int m = lane >> 2, xl = lane & 3;
  const int w = nw_width<P>(a);
  const cfloat* xp = a.x + (bt * P::H + m) * (long long)w + x0 + xl;
#pragma unroll
  for (int i = 0; i < P::G; ++i) ws[P::X_OFF + 32 * i + lane] = xp[(size_t)i * 8 * w];
  const uint8_t* mrow = a.mask + bt * P::H;
  f2* mk = reinterpret_cast<f2*>(ws + P::E_OFF);
  for (int n = lane; n < 8 * P::NP; n += 32) {
    const int p = n >> 3, k2 = n & 7, g = 2 * p;
    mk[2 * (p * P::EPQ + P::MKQ) + k2] = make_f2(mrow[g + P::G * k2] ? 1.f : 0.f, (g + 1 < P::G && mrow[g + 1 + P::G * k2]) ? 1.f : 0.f);
  }
}

// step 1: p = S_c x, radix-G over this lane's rows, twiddle (with the input checkerboard), E[p][m][xl]
template <class P>
B2S_HD void nw_step1(const NormalArgs& a, cfloat* ws, const cfloat* smem, const cfloat* sp, int lane, cfloat (&sv)[P::G]) {
  constexpr int G = P::G, NP = P::NP;
  const int m = lane >> 2;
  float re[2 * NP], im[2 * NP];
  const int w = nw_width<P>(a);
#pragma unroll
  for (int i = 0; i < G; ++i) sv[i] = sp[(size_t)i * 8 * w];   // S_c straight into registers: the other resident warps cover the latency
  {
    float pr[G], pi[G];
#pragma unroll
    for (int i = 0; i < G; ++i) {
      const cfloat xv = ws[P::X_OFF + 32 * i + lane];
      pr[i] = xv.x * sv[i].x - xv.y * sv[i].y;
      pi[i] = xv.x * sv[i].y + xv.y * sv[i].x;
    }
    Dft<G>::run(pr, pi);
#pragma unroll
    for (int g = 0; g < G; ++g) { re[g] = pr[g]; im[g] = pi[g]; }
    if (G & 1) { re[G] = 0.f; im[G] = 0.f; }
  }
  const nquad* t1 = reinterpret_cast<const nquad*>(smem + P::TW1_OFF);
  nquad* e = reinterpret_cast<nquad*>(ws + P::E_OFF);
#pragma unroll
  for (int p = 0; p < NP; ++p) {
    const nquad w = t1[p * P::TWPQ + m];
    const f2 r = make_f2(re[2 * p], re[2 * p + 1]), i2 = make_f2(im[2 * p], im[2 * p + 1]);
    const f2 wc = make_f2(w.ra, w.rb), wsn = make_f2(w.ia, w.ib);
    const f2 o_r = vsub(vmul2(r, wc), vmul2(i2, wsn));
    const f2 o_i = vfma2(r, wsn, vmul2(i2, wc));
    nquad q; q.ra = o_r.x; q.rb = o_r.y; q.ia = o_i.x; q.ib = o_i.y;
    e[p * P::EPQ + lane] = q;
  }
}

// step 2 (packed, both g of a quad): length-8 transform over m -> bins g + G k2, mask, inverse length-8 over {g + G j},
// twiddle (with the output checkerboard and 1/H); in place
template <class P>
B2S_HD void nw_step2(cfloat* ws, const cfloat* smem, int task) {
  const int p = task >> 2, xl = task & 3;
  nquad* e = reinterpret_cast<nquad*>(ws + P::E_OFF) + p * P::EPQ + xl;
  f2 re[8], im[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { const nquad v = e[4 * j]; re[j] = make_f2(v.ra, v.rb); im[j] = make_f2(v.ia, v.ib); }
  dft8(re, im);
  const nquad* mk = e - xl + P::MKQ;                       // this pair's mask factors, two k2 per 128-bit access
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const nquad v = mk[k];
    const f2 m0 = make_f2(v.ra, v.rb), m1 = make_f2(v.ia, v.ib);
    re[2 * k] = vmul2(re[2 * k], m0); im[2 * k] = vmul2(im[2 * k], m0);
    re[2 * k + 1] = vmul2(re[2 * k + 1], m1); im[2 * k + 1] = vmul2(im[2 * k + 1], m1);
  }
  dft8(im, re);                                           // inverse = forward on swapped data
  const nquad* t2 = reinterpret_cast<const nquad*>(smem + P::TW2_OFF) + p * P::TWPQ;
#pragma unroll
  for (int m = 0; m < 8; ++m) {
    const nquad w = t2[m];
    const f2 wc = make_f2(w.ra, w.rb), wsn = make_f2(w.ia, w.ib);
    // still in the swapped domain: value = (im, re)
    const f2 o_r = vsub(vmul2(im[m], wc), vmul2(re[m], wsn));
    const f2 o_i = vfma2(im[m], wsn, vmul2(re[m], wc));
    nquad q; q.ra = o_r.x; q.rb = o_r.y; q.ia = o_i.x; q.ib = o_i.y;
    e[4 * m] = q;
  }
}

// step 3: radix-G over g -> rows m + 8k (swapped domain), un-swap, conj(S_c), accumulate
template <class P>
B2S_HD void nw_step3(const cfloat* ws, int lane, const cfloat (&sv)[P::G], float (&accr)[P::G], float (&acci)[P::G]) {
  constexpr int G = P::G, NP = P::NP;
  float re[G], im[G];
  const nquad* e = reinterpret_cast<const nquad*>(ws + P::E_OFF);
#pragma unroll
  for (int p = 0; p < NP; ++p) {
    const nquad v = e[p * P::EPQ + lane];
    re[2 * p] = v.ra; im[2 * p] = v.ia;
    if (2 * p + 1 < G) { re[2 * p + 1] = v.rb; im[2 * p + 1] = v.ib; }
  }
  Dft<G>::run(re, im);
#pragma unroll
  for (int k = 0; k < G; ++k) {
    const float yr = im[k], yi = re[k];                   // un-swap
    accr[k] += yr * sv[k].x + yi * sv[k].y;
    acci[k] += yi * sv[k].x - yr * sv[k].y;
  }
}

// the coil sum starts at 0 (normal operator) or at -bref (image-domain cascade: A^H M A x - bref is what the epilogue needs, and
// loading bref here puts its HBM latency behind the x / S_0 loads instead of at the end of the item)
template <class P>
B2S_HD void nw_init_acc(const NormalArgs& a, long long bt, int x0, int lane, float (&accr)[P::G], float (&acci)[P::G]) {
  if (a.mode == 0) {
#pragma unroll
    for (int k = 0; k < P::G; ++k) { accr[k] = 0.f; acci[k] = 0.f; }
  } else {
    const int w = nw_width<P>(a);
    const cfloat* bp = a.bref + (bt * P::H + (lane >> 2)) * (long long)w + x0 + (lane & 3);
#pragma unroll
    for (int k = 0; k < P::G; ++k) { const cfloat br = bp[(size_t)k * 8 * w]; accr[k] = -br.x; acci[k] = -br.y; }
  }
}

// returns this lane's share of <x, H x> (mode 0; 0 otherwise)
template <class P>
B2S_HD float nw_finish(const NormalArgs& a, const cfloat* ws, long long bt, int x0, int lane, const float (&accr)[P::G],
                       const float (&acci)[P::G], float v) {
  constexpr int G = P::G;
  const int m = lane >> 2, xl = lane & 3;
  const int w = nw_width<P>(a);
  const long long hw = (long long)P::H * w;
  const long long pix0 = (long long)m * w + x0 + xl;
  cfloat* op = a.out + bt * hw + pix0;
  const float eta = v / (1.f + v);
  float dsum = 0.f;
  if (a.mode == 0) {
#pragma unroll
    for (int k = 0; k < G; ++k) {
      const cfloat xv = ws[P::X_OFF + 32 * k + lane];
      const float hr = accr[k] + v * xv.x, hi = acci[k] + v * xv.y;
      op[(size_t)k * 8 * w] = make_c(hr, hi);
      dsum += xv.x * hr + xv.y * hi;
    }
  } else {
    const float* dp = a.ssq + (bt / a.T) * hw + pix0;
    float* mp = reinterpret_cast<float*>(a.out) + bt * hw + pix0;
#pragma unroll
    for (int k = 0; k < G; ++k) {
      const cfloat xv = ws[P::X_OFF + 32 * k + lane];
      const float d = dp[(size_t)k * 8 * w];
      const float re = d * xv.x - eta * accr[k], im = d * xv.y - eta * acci[k];      // acc = A^H M A x - bref (nw_init_acc)
      if (a.mode == 1) op[(size_t)k * 8 * w] = make_c(re, im);
      else mp[(size_t)k * 8 * w] = sqrtf(re * re + im * im);           // complex_abs, utils/math.py:41-56
    }
  }
  return dsum;
}

#if defined(__CUDACC__)
template <class P>
__global__ void __launch_bounds__(P::NT, P::CTAS) normal_warp_kernel(const NormalArgs a, long long n_items) {
  extern __shared__ __align__(16) unsigned char b2s_smem_raw[];
  cfloat* smem = reinterpret_cast<cfloat*>(b2s_smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  nw_build_tables<P>(smem, tid, P::NT);
  __syncthreads();                                        // the only CTA barrier
  const int part = warp % P::SPLIT;
  const long long item = (long long)blockIdx.x * P::ITEMS + warp / P::SPLIT;
  if (item >= n_items) return;                            // (all SPLIT warps of an item leave together)
  cfloat* ws = smem + warp * P::WARP_ELEMS;
  const int w = nw_width<P>(a);
  const int groups = w / P::XC;
  const long long bt = item / groups;
  const int x0 = (int)(item - bt * groups) * P::XC;
  const size_t hw = (size_t)P::H * w;
  const int c0 = (int)((long long)part * a.C / P::SPLIT), c1 = (int)((long long)(part + 1) * a.C / P::SPLIT);   // this warp's coils
  const cfloat* sp = a.sens + ((size_t)(bt / a.T) * a.C + c0) * hw + (size_t)(lane >> 2) * w + x0 + (lane & 3);
  nw_stage<P>(a, ws, bt, x0, lane);
  float accr[P::G], acci[P::G];
  cfloat sv[P::G];
  if (part == 0) nw_init_acc<P>(a, bt, x0, lane, accr, acci);
  else {
#pragma unroll
    for (int k = 0; k < P::G; ++k) { accr[k] = 0.f; acci[k] = 0.f; }
  }
  __syncwarp();                                           // mask factors visible to the warp
#pragma unroll 1
  for (int c = c0; c < c1; ++c) {
    nw_step1<P>(a, ws, smem, sp, lane, sv);
    sp += hw;
    __syncwarp();
#pragma unroll 1
    for (int r = 0; r < P::ROUNDS2; ++r) {
      const int task = lane + 32 * r;
      if (task < P::TASKS2) nw_step2<P>(ws, smem, task);
    }
    __syncwarp();
    nw_step3<P>(ws, lane, sv, accr, acci);
    __syncwarp();
  }
  if (P::SPLIT > 1) {                                     // partial coil sums of the item's other warps -> warp `part == 0` (fixed order)
    if (part != 0) {
#pragma unroll
      for (int k = 0; k < P::G; ++k) ws[P::E_OFF + 32 * k + lane] = make_c(accr[k], acci[k]);   // (the exchange buffer is free now)
    }
    asm volatile("bar.sync %0, %1;" ::"r"(1 + warp / P::SPLIT), "r"(32 * P::SPLIT) : "memory");   // named barrier of this item's warps
    if (part != 0) return;
    for (int q = 1; q < P::SPLIT; ++q) {
      const cfloat* other = ws + q * P::WARP_ELEMS + P::E_OFF;
#pragma unroll
      for (int k = 0; k < P::G; ++k) { const cfloat v = other[32 * k + lane]; accr[k] += v.x; acci[k] += v.y; }
    }
  }
  float dsum = nw_finish<P>(a, ws, bt, x0, lane, accr, acci, *a.vptr);
  if (a.dot_part) {                                       // fixed-order butterfly over the warp: bit-reproducible partials
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) dsum += __shfl_xor_sync(0xffffffffu, dsum, o);
    if (lane == 0) a.dot_part[item] = dsum;
  }
}
#endif

template <class P>
void normal_warp_emulate(const NormalArgs& a, long long n_bt) {
  cfloat* smem = new cfloat[P::SMEM_ELEMS];
  cfloat* ws = smem;                                      // warp 0's slice
  float (*accr)[P::G] = new float[32][P::G];
  float (*acci)[P::G] = new float[32][P::G];
  cfloat (*sv)[P::G] = new cfloat[32][P::G];
  nw_build_tables<P>(smem, 0, 1);
  const int w = nw_width<P>(a);
  const int groups = w / P::XC;
  const size_t hw = (size_t)P::H * w;
  for (long long item = 0; item < n_bt * groups; ++item) {
    const long long bt = item / groups;
    const int x0 = (int)(item - bt * groups) * P::XC;
    const cfloat* sp[32];
    for (int lane = 0; lane < 32; ++lane) {
      sp[lane] = a.sens + (size_t)(bt / a.T) * a.C * hw + (size_t)(lane >> 2) * w + x0 + (lane & 3);
      nw_stage<P>(a, ws, bt, x0, lane);
      nw_init_acc<P>(a, bt, x0, lane, accr[lane], acci[lane]);
    }
    float (*pr)[P::G] = new float[32][P::G];
    float (*pi)[P::G] = new float[32][P::G];
    for (int part = 0; part < P::SPLIT; ++part) {           // the item's warps one after the other, partial sums added in warp order
      const int c0 = (int)((long long)part * a.C / P::SPLIT), c1 = (int)((long long)(part + 1) * a.C / P::SPLIT);
      float (*ar)[P::G] = part == 0 ? accr : pr;
      float (*ai)[P::G] = part == 0 ? acci : pi;
      if (part > 0) for (int lane = 0; lane < 32; ++lane) for (int k = 0; k < P::G; ++k) { ar[lane][k] = 0.f; ai[lane][k] = 0.f; }
      for (int c = c0; c < c1; ++c) {
        for (int lane = 0; lane < 32; ++lane) nw_step1<P>(a, ws, smem, sp[lane] + (size_t)c * hw, lane, sv[lane]);
        for (int task = 0; task < P::TASKS2; ++task) nw_step2<P>(ws, smem, task);
        for (int lane = 0; lane < 32; ++lane) nw_step3<P>(ws, lane, sv[lane], ar[lane], ai[lane]);
      }
      if (part > 0) for (int lane = 0; lane < 32; ++lane) for (int k = 0; k < P::G; ++k) { accr[lane][k] += pr[lane][k]; acci[lane][k] += pi[lane][k]; }
    }
    delete[] pr; delete[] pi;
    float dsum = 0.f;
    for (int lane = 0; lane < 32; ++lane) dsum += nw_finish<P>(a, ws, bt, x0, lane, accr[lane], acci[lane], *a.vptr);
    if (a.dot_part) a.dot_part[item] = dsum;
  }
  delete[] smem; delete[] accr; delete[] acci; delete[] sv;
}

}  // namespace b2s
