// CineNet normal operator  H x = A^H M A x + v x  (models/cinenet.py:121-133,
// recurrent_cinenet.py:74-86) with the k-space kept on chip.
//
// The mask selects k-space ROWS only (data/subsample.py:146-151), so it commutes
// with the transform along w and  F^H M F = (F_h^H M F_h) (x) I_w :
//
//     H x = sum_c conj(S_c) . [ F_h^H M F_h (S_c . x) ]  +  v x .
//
// No FFT along w is needed at all, every image column is independent, and the
// c * K bytes of intermediate k-space of the reference never exist.  One CTA
// owns XC adjacent columns of one frame, loops over the coils and keeps the
// coil sum in registers (deterministic, no atomics).
//
// Length-H transform, H = 8*G (G = 25):  forward  = radix-G over rows {m + 8 i}
// (thread (m, x) owns exactly those rows, the same rows its accumulator holds),
// twiddle, radix-8 over m giving bins k = g + G*k2;  the mask is applied there
// and the inverse starts immediately in the same registers with the radix-8 over
// {g + G*j}, twiddle, then radix-G giving rows m + 8*k.  Two shared-memory
// exchanges per coil.  Centring: input/output signs (-1)^y = (-1)^m.
#pragma once
#include "fft2_core.cuh"

namespace b2s {

template <int H_, int XC_> struct NormalPlan {
  static constexpr int H = H_, G = H_ / 8, XC = XC_;
  static constexpr int NT = 8 * XC_;
  static constexpr int E_ELEMS = H_ * XC_;           // exchange buffer [g][m][x]
  static constexpr int X_OFF = E_ELEMS;              // this CTA's columns of x_t, [row][x] (read by every coil)
  static constexpr int L_OFF = X_OFF + H_ * XC_;     // landing buffer of the NEXT coil's S values (cp.async, thread-private slots)
  static constexpr int TW_OFF = L_OFF + H_ * XC_;    // w_H^n, n in [0,H)
  static constexpr int SMEM_ELEMS = TW_OFF + H_;
  static constexpr int SMEM_BYTES = SMEM_ELEMS * 8 + H_;   // + mask row (uint8)
  static constexpr int TASKS2 = G * XC_;
  static_assert(H_ % 8 == 0, "H must be a multiple of 8");
};

// mode 0: out = A^H M A x + v x                        (HOperator, cinenet.py:121-133)
// mode 1: out = ssq . x - eta (A^H M A x - bref)        (one VarNet cascade in the image domain:
//         A^H[ DC(A x, ref) ] with ssq = sum_c |S_c|^2, bref = A^H ref, eta = v/(1+v); varnet.py:253-282
//         followed by the next cascade's / the model's sens_reduce, varnet.py:150-151,253)
struct NormalArgs {
  const cfloat* x; const cfloat* sens; const uint8_t* mask; const float* vptr; cfloat* out;
  int T, C, W;
  int mode; const float* ssq; const cfloat* bref;
};

// step 0 (once per CTA): this thread's rows {m + 8i} of x_t into shared memory
template <class P>
B2S_HD void normal_stage_x(const NormalArgs& a, cfloat* smem, long long bt, int x0, int tid) {
  constexpr int G = P::G, XC = P::XC;
  const int m = tid / XC, xl = tid - m * XC;
  const cfloat* xp = a.x + bt * (long long)P::H * a.W + x0 + xl;
#pragma unroll
  for (int i = 0; i < G; ++i) smem[P::X_OFF + (m + 8 * i) * XC + xl] = xp[(long long)(m + 8 * i) * a.W];
}

// asynchronous copy of one 8-byte element global -> shared (cp.async); a plain copy on the host
B2S_HD void async_copy8(cfloat* dst_smem, const cfloat* src) {
#if defined(__CUDA_ARCH__)
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(dst_smem)), "l"(src) : "memory");
#else
  *dst_smem = *src;
#endif
}
B2S_HD void async_wait_all() {
#if defined(__CUDA_ARCH__)
  asm volatile("cp.async.wait_all;" ::: "memory");
#endif
}

// this thread's 25 values of S_c into its own slots of the landing buffer (read back only by this thread)
template <class P>
B2S_HD void normal_prefetch_s(const NormalArgs& a, cfloat* smem, long long bt, int c, int x0, int tid) {
  constexpr int G = P::G, XC = P::XC;
  const int m = tid / XC, xl = tid - m * XC;
  const long long b = bt / a.T;
  const cfloat* sp = a.sens + (b * a.C + c) * (long long)P::H * a.W + x0 + xl;
#pragma unroll
  for (int i = 0; i < G; ++i) async_copy8(smem + P::L_OFF + (m + 8 * i) * XC + xl, sp + (long long)(m + 8 * i) * a.W);
}

// step 1: p = S_c x, radix-G over this thread's rows, twiddle, store E[g][m][xl]; S_c values stay in sv
template <class P>
B2S_HD void normal_step1(const NormalArgs& a, cfloat* smem, long long bt, int c, int x0, int tid, cfloat (&sv)[P::G]) {
  constexpr int G = P::G, XC = P::XC;
  const int m = tid / XC, xl = tid - m * XC;
  float re[G], im[G];
  const float sg = (m & 1) ? -1.f : 1.f;
  async_wait_all();                                   // S_c landed (issued one coil ago)
#pragma unroll
  for (int i = 0; i < G; ++i) sv[i] = smem[P::L_OFF + (m + 8 * i) * XC + xl];
  if (c + 1 < a.C) normal_prefetch_s<P>(a, smem, bt, c + 1, x0, tid);   // S_{c+1} streams in behind this coil's math
#pragma unroll
  for (int i = 0; i < G; ++i) {
    const cfloat xv = smem[P::X_OFF + (m + 8 * i) * XC + xl];
    re[i] = (xv.x * sv[i].x - xv.y * sv[i].y) * sg;
    im[i] = (xv.x * sv[i].y + xv.y * sv[i].x) * sg;
  }
  Dft<G>::run(re, im);
  const cfloat* tw = smem + P::TW_OFF;
#pragma unroll
  for (int g = 0; g < G; ++g) {
    const cfloat w = tw[m * g];                       // m*g <= 7*(G-1) < H
    smem[(g * 8 + m) * XC + xl] = make_c(re[g] * w.x - im[g] * w.y, re[g] * w.y + im[g] * w.x);
  }
}

// step 2: radix-8 over m -> bins g + G*k2, mask, inverse radix-8 over {g + G*j}, twiddle, in place
template <class P>
B2S_HD void normal_step2(cfloat* smem, const uint8_t* mrow, int task) {
  constexpr int G = P::G, XC = P::XC;
  const int g = task / XC, xl = task - g * XC;
  float re[8], im[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { const cfloat v = smem[(g * 8 + j) * XC + xl]; re[j] = v.x; im[j] = v.y; }
  dft8(re, im);
#pragma unroll
  for (int k2 = 0; k2 < 8; ++k2) {
    const float mk = mrow[g + G * k2] ? 1.f : 0.f;
    re[k2] *= mk; im[k2] *= mk;
  }
  dft8(im, re);                                        // inverse = forward on swapped data
  const cfloat* tw = smem + P::TW_OFF;
#pragma unroll
  for (int m = 0; m < 8; ++m) {
    const cfloat w = tw[m * g];
    // still in the swapped domain: value = (im, re)
    smem[(g * 8 + m) * XC + xl] = make_c(im[m] * w.x - re[m] * w.y, im[m] * w.y + re[m] * w.x);
  }
}

// step 3: radix-G over g -> rows m + 8k (swapped domain), un-swap, sign, conj(S_c), accumulate
template <class P>
B2S_HD void normal_step3(const cfloat* smem, int tid, const cfloat (&sv)[P::G],
                         float (&accr)[P::G], float (&acci)[P::G], float scale) {
  constexpr int G = P::G, XC = P::XC;
  const int m = tid / XC, xl = tid - m * XC;
  float re[G], im[G];
#pragma unroll
  for (int g = 0; g < G; ++g) { const cfloat v = smem[(g * 8 + m) * XC + xl]; re[g] = v.x; im[g] = v.y; }
  Dft<G>::run(re, im);
  const float s = (m & 1) ? -scale : scale;
#pragma unroll
  for (int k = 0; k < G; ++k) {
    const float yr = im[k] * s, yi = re[k] * s;        // un-swap
    accr[k] += yr * sv[k].x + yi * sv[k].y;
    acci[k] += yi * sv[k].x - yr * sv[k].y;
  }
}

template <class P>
B2S_HD void normal_finish(const NormalArgs& a, const cfloat* smem, long long bt, int x0, int tid, const float (&accr)[P::G],
                          const float (&acci)[P::G], float v) {
  constexpr int G = P::G, XC = P::XC;
  const int m = tid / XC, xl = tid - m * XC, x = x0 + xl;
  const long long hw = (long long)P::H * a.W;
  const float eta = v / (1.f + v);
#pragma unroll
  for (int k = 0; k < G; ++k) {
    const long long pix = (long long)(m + 8 * k) * a.W + x;
    const long long off = bt * hw + pix;
    const cfloat xv = smem[P::X_OFF + (m + 8 * k) * XC + xl];
    if (a.mode == 0) {
      a.out[off] = make_c(accr[k] + v * xv.x, acci[k] + v * xv.y);
    } else {
      const float d = a.ssq[(bt / a.T) * hw + pix];
      const cfloat br = a.bref[off];
      a.out[off] = make_c(d * xv.x - eta * (accr[k] - br.x), d * xv.y - eta * (acci[k] - br.y));
    }
  }
}

#if defined(__CUDACC__)
template <class P>
__global__ void __launch_bounds__(P::NT, 2) normal_op_kernel(const NormalArgs a) {
  extern __shared__ __align__(16) unsigned char b2s_smem_raw[];
  cfloat* smem = reinterpret_cast<cfloat*>(b2s_smem_raw);
  uint8_t* mrow = reinterpret_cast<uint8_t*>(smem + P::SMEM_ELEMS);
  const int tid = threadIdx.x;
  const int chunks = a.W / P::XC;
  const long long bt = blockIdx.x / chunks;
  const int x0 = (blockIdx.x % chunks) * P::XC;
  for (int n = tid; n < P::H; n += P::NT) { smem[P::TW_OFF + n] = twiddle(n, P::H); mrow[n] = a.mask[bt * P::H + n]; }
  normal_stage_x<P>(a, smem, bt, x0, tid);            // each thread only ever reads back its own rows
  normal_prefetch_s<P>(a, smem, bt, 0, x0, tid);
  float accr[P::G], acci[P::G];
  cfloat sv[P::G];
#pragma unroll
  for (int k = 0; k < P::G; ++k) { accr[k] = 0.f; acci[k] = 0.f; }
  const float scale = 1.f / (float)P::H;              // ortho forward * ortho inverse along h
  __syncthreads();
#pragma unroll 1
  for (int c = 0; c < a.C; ++c) {
    normal_step1<P>(a, smem, bt, c, x0, tid, sv);
    __syncthreads();
    for (int task = tid; task < P::TASKS2; task += P::NT) normal_step2<P>(smem, mrow, task);
    __syncthreads();
    normal_step3<P>(smem, tid, sv, accr, acci, scale);
    __syncthreads();
  }
  normal_finish<P>(a, smem, bt, x0, tid, accr, acci, *a.vptr);
}
#endif

template <class P>
void normal_op_emulate(const NormalArgs& a, long long n_bt) {
  cfloat* smem = new cfloat[P::SMEM_ELEMS];
  uint8_t* mrow = new uint8_t[P::H];
  float (*accr)[P::G] = new float[P::NT][P::G];
  float (*acci)[P::G] = new float[P::NT][P::G];
  cfloat (*sv)[P::G] = new cfloat[P::NT][P::G];
  const int chunks = a.W / P::XC;
  for (long long blk = 0; blk < n_bt * chunks; ++blk) {
    const long long bt = blk / chunks;
    const int x0 = (int)(blk % chunks) * P::XC;
    for (int n = 0; n < P::H; ++n) { smem[P::TW_OFF + n] = twiddle(n, P::H); mrow[n] = a.mask[bt * P::H + n]; }
    for (int tid = 0; tid < P::NT; ++tid) for (int k = 0; k < P::G; ++k) { accr[tid][k] = 0.f; acci[tid][k] = 0.f; }
    for (int tid = 0; tid < P::NT; ++tid) { normal_stage_x<P>(a, smem, bt, x0, tid); normal_prefetch_s<P>(a, smem, bt, 0, x0, tid); }
    for (int c = 0; c < a.C; ++c) {
      for (int tid = 0; tid < P::NT; ++tid) normal_step1<P>(a, smem, bt, c, x0, tid, sv[tid]);
      for (int tid = 0; tid < P::NT; ++tid)
        for (int task = tid; task < P::TASKS2; task += P::NT) normal_step2<P>(smem, mrow, task);
      for (int tid = 0; tid < P::NT; ++tid) normal_step3<P>(smem, tid, sv[tid], accr[tid], acci[tid], 1.f / (float)P::H);
    }
    for (int tid = 0; tid < P::NT; ++tid) normal_finish<P>(a, smem, bt, x0, tid, accr[tid], acci[tid], *a.vptr);
  }
  delete[] smem; delete[] mrow; delete[] accr; delete[] acci; delete[] sv;
}

}  // namespace b2s
