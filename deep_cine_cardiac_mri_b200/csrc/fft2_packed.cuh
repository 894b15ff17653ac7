// Packed whole-image fused centred 2-D FFT, 200 x 200 (the dataset crop, data/mri_data.py:273-277).
//
// Two ideas on top of the half-split design of fft2_core.cuh:
//
// 1. ONE work item = ONE coil image (fft2_whole.cuh): every input element is loaded once and every S*x product formed
//    once; the half of the intermediate that does not fit in shared memory is parked in the SM's tensor memory (TMEM).
//
// 2. TWO transforms per thread in packed fp32 (packed.cuh): every butterfly of Phases A (second half), B and C works
//    on an aligned register pair holding the same element of two independent sub-transforms, so it is ONE FADD2 /
//    FMUL2 / FFMA2 instead of two scalar instructions.  The pairs are
//        (m-block 2*rp, m-block 2*rp + 1)   i.e. two of the four radix-8 output residues kept per parity,
//    because that pairing survives all three phases: Phase A finishes both m-blocks of a pair with the same row
//    twiddle / radix-5 / column twiddle sequence, Phase B transforms the same row segment of both, Phase C runs the
//    25-point transform down the same column of both.  Shared memory therefore holds B as 16-byte quads
//    {re_a, re_b, im_a, im_b}: every shared-memory access is 128-bit and lands directly in two register pairs.
//    Scalar arithmetic remains only where values are (un)paired for free: the first butterflies on freshly loaded data
//    (prologue product, radix-8 folds, the last radix-2 stage that regroups (parity 0, parity 1) into
//    (m-block a, m-block b)) and the final scale multiply of Phase C in front of the 64-bit global stores.
//
// Same arithmetic, same order of operations per element as the scalar kernels up to fp32 rounding of fused
// multiply-adds; parity tests are shared (tests/test_gpu_parity.py, tests/host_emul).
#pragma once
#include "fft2_whole.cuh"

namespace b2s {

struct alignas(16) cquad { float ra, rb, ia, ib; };          // one element of B: two complex values (a, b)
struct alignas(16) u32x4 { uint32_t x, y, z, w; };
B2S_HD cquad make_q(f2 re, f2 im) { cquad q; q.ra = re.x; q.rb = re.y; q.ia = im.x; q.ib = im.y; return q; }

template <int NT_ = 256> struct PackPlan200 {
  static constexpr int H = 200, W = 200, G = 25, R = 5, X0 = 40;
  static constexpr int SEG = 45;                             // k1-segment pitch: column kx = k1 + 5 k2 lives at k1*SEG + k2 through Phases
                                                             // B AND C; 45 = 5 mod 8 makes Phase C's 128-bit reads (lanes along kx) conflict-free
  static constexpr int PITCH = 225;                          // quads per row pair; odd: Phase B (lanes along rows) conflict-free
  static constexpr int NT = NT_;
  static constexpr int ROWS = 2 * G;                         // row pairs of B: (rp, g)
  static constexpr int B_QUADS = ROWS * PITCH;
  static constexpr int TW_OFF = B_QUADS;                     // TWP[W]   {c, c, s, s}
  static constexpr int TH_OFF = TW_OFF + W;                  // THP[2 sx][2 q][G][2 rp]  {re_a, re_b, im_a, im_b}
  static constexpr int SMEM_QUADS = TH_OFF + 8 * G;
  static constexpr int MASK_BYTES = (H + 15) / 16 * 16;
  static constexpr int AUX_BYTES = 2 * MASK_BYTES + 16;      // as Derived<P>::AUX_BYTES (EpiDCFix layout)
  static constexpr int SMEM_BYTES = SMEM_QUADS * 16 + 2 * AUX_BYTES + 16;   // + TMEM address slot and one mbarrier
  static constexpr int TASKS_A = G * X0, TPT = (TASKS_A + NT - 1) / NT;
  static constexpr int TASKS_B = ROWS * R;
  static constexpr int TASKS_C = 2 * W;
  static constexpr int PARK = 2 * R * 4;                     // floats parked per task: parity 1, two row pairs
  static_assert(R * SEG <= PITCH && W <= PITCH && (PITCH & 1), "pitch");
  static_assert(TPT * PARK <= 256, "parking lot: 256 TMEM columns per thread");
  static_assert(TASKS_B <= NT, "Phase B is one round");
};

// staged soft-DC epilogue (EpiDCStage below): extra shared memory behind the TMEM slot
template <class Epi, class = void> struct PkStaged { static constexpr bool value = false; };
template <class Epi> struct PkStaged<Epi, decltype((void)Epi::STAGED)> { static constexpr bool value = Epi::STAGED; };
template <class P> struct PkStage {
  static constexpr int TAB_BYTES = P::H * 16 + 16 + 8 * 32 + 16;                 // DcRow per output row, 16 bytes of zeros, slot bytes + flags
  static constexpr int MAX_SMEM = 232448;                                        // 227 KB: the most one CTA can have
  static constexpr int CAP = (MAX_SMEM - P::SMEM_BYTES - TAB_BYTES) / (P::W * 8);   // rows of the staging buffer
  static constexpr int BYTES = TAB_BYTES + CAP * P::W * 8;
  // ... plus HOLE_ROWS more rows in the padding of B itself: each k1 segment of a row pair is SEG = 45 quads of which 40
  // are used, i.e. 5 holes of 80 bytes (10 complex) per row pair; staged row h of that kind occupies the 20 holes of
  // row pairs 4h .. 4h+3: element kx at quad ((4h + kx/50) * PITCH + ((kx/10) % 5) * SEG + X0) + 8 (kx % 10) bytes
  static constexpr int HOLE = P::SEG - P::X0;                                    // quads per hole
  static constexpr int HOLE_CPX = 2 * HOLE;                                      // complex values per hole
  static constexpr int HOLES_PER_ROW = P::W / HOLE_CPX;                          // 20
  static constexpr int RP_PER_ROW = HOLES_PER_ROW / P::R;                        // 4 row pairs per staged row
  static constexpr int HOLE_ROWS = P::ROWS / RP_PER_ROW;                         // 12
  static constexpr int HOLE_STRIDE = RP_PER_ROW * P::PITCH * 16;                 // bytes between consecutive hole rows
  static constexpr int CAP_ALL = CAP + HOLE_ROWS;
  static_assert(CAP >= 16 && CAP_ALL < 0xfe && P::W % HOLE_CPX == 0 && HOLES_PER_ROW % P::R == 0, "staging buffer");
  // byte offset (from the start of B) of element kx of hole row 0
  static B2S_HD int hole_base(int kx) {
    const int id = kx / HOLE_CPX;
    return ((id / P::R) * P::PITCH + (id % P::R) * P::SEG + P::X0) * 16 + (kx % HOLE_CPX) * 8;
  }
};
template <class P, class Epi> struct PkSmem { static constexpr int BYTES = P::SMEM_BYTES + (PkStaged<Epi>::value ? PkStage<P>::BYTES : 0); };

// output row residue (mod 8) of m-block r of parity q (G odd: the (-1)^(G j) part of the input checkerboard relabels by 4)
B2S_HD int pk_m_of(int r, int q) { return (2 * r + q + 4) & 7; }

template <class P> B2S_HD void pk_build_tables(cquad* smem, int tid, int nthreads) {
  for (int n = tid; n < P::W; n += nthreads) {
    const cfloat t = twiddle(n, P::W);
    cquad q; q.ra = t.x; q.rb = t.x; q.ia = t.y; q.ib = t.y;
    smem[P::TW_OFF + n] = q;
  }
  for (int e = tid; e < 8 * P::G; e += nthreads) {
    const int rp = e & 1, g = (e >> 1) % P::G, q = (e / (2 * P::G)) & 1, sxb = e / (4 * P::G);
    cfloat a = twiddle(g * pk_m_of(2 * rp, q), P::H), b = twiddle(g * pk_m_of(2 * rp + 1, q), P::H);
    const float s = (((g & 1) != 0) != (sxb != 0)) ? -1.f : 1.f;      // (-1)^g (-1)^x of the input checkerboard
    cquad v; v.ra = a.x * s; v.rb = b.x * s; v.ia = a.y * s; v.ib = b.y * s;
    smem[P::TH_OFF + e] = v;
  }
}

// --------------------------------------------------------------------------- //
// Phase A
// --------------------------------------------------------------------------- //
template <class P, class Pro, int QD_, int TT_> struct PkPhaseA {
  static constexpr int G = P::G, R = P::R, X0 = P::X0, NT = P::NT, TPT = P::TPT, PARK = P::PARK;
  static constexpr int STEPS = TPT * R;
  static constexpr int QD = QD_, TT = TT_;
  static_assert(STEPS % QD == 0 && (TT * R) % QD == 0 && TPT % TT == 0, "queue depth must divide one trip's steps");
  typedef typename Pro::template Unit<1> Unit;
  struct Queue { Unit u[QD]; };

  static B2S_HD bool task_of(int tid, int k, int& g, int& x0) {
    const int task = tid + k * NT;
    g = task / X0; x0 = task - g * X0;
    return task < P::TASKS_A;
  }
  static B2S_HD void issue(const Pro& pro, const typename Pro::Ctx& ctx, int tid, int s, Unit& u) {
    int g, x0;
    if (!task_of(tid, s / R, g, x0)) return;
    pro.template fetch<1, G * P::W>(ctx, g, g * P::W + x0 + X0 * (s % R), u);
  }
  static B2S_HD void prefill(const Pro& pro, const typename Pro::Ctx& ctx, int tid, Queue& q) {
#pragma unroll
    for (int s = 0; s < QD; ++s) issue(pro, ctx, tid, s, q.u[s]);
  }

  template <bool SYNC_FIRST, class Park>
  static B2S_HD void run(const Pro& pro, const typename Pro::Ctx& ctx, const typename Pro::Ctx& next, bool has_next,
                         cquad* smem, int tid, Queue& qu, const Park& park) {
    const float h = 0.70710678118654752440f;
    f2 ar[2][2][R], ai[2][2][R];            // [parity q][row pair rp][column-group index i] = (m-block 2rp, m-block 2rp+1)
#pragma unroll 1
    for (int kp = 0; kp < TPT; kp += TT)
#pragma unroll
    for (int u = 0; u < TT * R; ++u) {
      const int s = kp * R + u;
      const int k = kp + u / R, i = u % R, slot = u % QD;
      int g, x0;
      const bool valid = task_of(tid, k, g, x0);
      {
        // radix-8 DIF over the 8 rows.  Folds (scalar, on the freshly produced prologue values): even outputs come
        // from e_j = a_j + a_{j+4}, odd outputs from o_j = (a_j - a_{j+4}) w8^j; both go through the same radix-4,
        // so they are paired (e_j, o_j) for its first stage (packed) ...
        f2 pr[4], pi[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float a_r[1], a_i[1], b_r[1], b_i[1];
          pro.template value<1>(qu.u[slot], j, a_r, a_i);
          pro.template value<1>(qu.u[slot], j + 4, b_r, b_i);
          const float er = a_r[0] + b_r[0], ei = a_i[0] + b_i[0];
          float o_r, o_i;
          if (j == 0)      { o_r = a_r[0] - b_r[0]; o_i = a_i[0] - b_i[0]; }
          else if (j == 1) { const float dr = a_r[0] - b_r[0], di = a_i[0] - b_i[0]; o_r = (dr + di) * h; o_i = (di - dr) * h; }
          else if (j == 2) { o_r = a_i[0] - b_i[0]; o_i = b_r[0] - a_r[0]; }
          else             { const float dr = a_r[0] - b_r[0], di = a_i[0] - b_i[0]; o_r = (di - dr) * h; o_i = (dr + di) * (-h); }
          pr[j] = make_f2(er, o_r); pi[j] = make_f2(ei, o_i);
        }
        const f2 t0r = vadd(pr[0], pr[2]), t0i = vadd(pi[0], pi[2]);
        const f2 t1r = vsub(pr[0], pr[2]), t1i = vsub(pi[0], pi[2]);
        const f2 t2r = vadd(pr[1], pr[3]), t2i = vadd(pi[1], pi[3]);
        const f2 t3r = vsub(pr[1], pr[3]), t3i = vsub(pi[1], pi[3]);
        // ... and its last stage (scalar) regroups the results as (output r, output r + 1) of ONE parity:
        // y0 = t0 + t2, y1 = t1 - i t3 | y2 = t0 - t2, y3 = t1 + i t3
        ar[0][0][i] = make_f2(t0r.x + t2r.x, t1r.x + t3i.x); ai[0][0][i] = make_f2(t0i.x + t2i.x, t1i.x - t3r.x);
        ar[0][1][i] = make_f2(t0r.x - t2r.x, t1r.x - t3i.x); ai[0][1][i] = make_f2(t0i.x - t2i.x, t1i.x + t3r.x);
        ar[1][0][i] = make_f2(t0r.y + t2r.y, t1r.y + t3i.y); ai[1][0][i] = make_f2(t0i.y + t2i.y, t1i.y - t3r.y);
        ar[1][1][i] = make_f2(t0r.y - t2r.y, t1r.y - t3i.y); ai[1][1][i] = make_f2(t0i.y - t2i.y, t1i.y + t3r.y);
      }
      // ---- refill the slot with step s + QD (of this image, else of the next one)
      if (s + QD < STEPS) issue(pro, ctx, tid, s + QD, qu.u[slot]);
      else if (has_next) issue(pro, next, tid, s + QD - STEPS, qu.u[slot]);
      if (SYNC_FIRST && u == R - 1) {          // first write into B of this image: every warp must have left the
        if (kp == 0) B2S_CTA_SYNC();           // previous image's last Phase C
      }
      // ---- last column group of the task: row twiddles, radix-R over the column groups, column twiddles (all packed)
      if (i == R - 1) {
        int gg = valid ? g : 0, x = valid ? x0 : 0;
        B2S_OPAQUE(gg);
        B2S_OPAQUE(x);
        f2 twr[R], twi[R];
#pragma unroll
        for (int k1 = 1; k1 < R; ++k1) { const cquad t = smem[P::TW_OFF + (x * k1) % P::W]; twr[k1] = make_f2(t.ra, t.rb); twi[k1] = make_f2(t.ia, t.ib); }
        float pk[PARK];
#pragma unroll
        for (int q = 0; q < 2; ++q)
#pragma unroll
          for (int rp = 0; rp < 2; ++rp) {
            const cquad th = smem[P::TH_OFF + ((((x & 1) * 2 + q) * G + gg) << 1) + rp];
            const f2 thr = make_f2(th.ra, th.rb), thi = make_f2(th.ia, th.ib);
            f2 nr[R], ni[R];
#pragma unroll
            for (int ii = 0; ii < R; ++ii) {
              nr[ii] = vsub(vmul2(ar[q][rp][ii], thr), vmul2(ai[q][rp][ii], thi));
              ni[ii] = vadd(vmul2(ar[q][rp][ii], thi), vmul2(ai[q][rp][ii], thr));
            }
            Dft<R>::run(nr, ni);
#pragma unroll
            for (int k1 = 0; k1 < R; ++k1) {
              f2 o_r = nr[k1], o_i = ni[k1];
              if (k1 > 0) {
                o_r = vsub(vmul2(nr[k1], twr[k1]), vmul2(ni[k1], twi[k1]));
                o_i = vadd(vmul2(nr[k1], twi[k1]), vmul2(ni[k1], twr[k1]));
              }
              if (q == 0) {
                if (valid) smem[(rp * G + gg) * P::PITCH + x + k1 * P::SEG] = make_q(o_r, o_i);
              } else {
                float* d = pk + (rp * R + k1) * 4;
                d[0] = o_r.x; d[1] = o_r.y; d[2] = o_i.x; d[3] = o_i.y;
              }
            }
          }
        park.template store<PARK>(k * PARK, pk);   // unconditional: the TMEM store is warp-collective
      }
    }
  }

  // parked parity-1 quads -> B (the places Phase A writes parity 0 to)
  template <class Park>
  static B2S_HD void unpark(cquad* smem, int tid, const Park& park) {
    park.wait_st();
#pragma unroll 1
    for (int k = 0; k < TPT; ++k) {
      float pk[PARK];
      park.template load<PARK>(k * PARK, pk);
      park.wait_ld();
      int g, x0;
      if (task_of(tid, k, g, x0)) {
#pragma unroll
        for (int rp = 0; rp < 2; ++rp)
#pragma unroll
          for (int k1 = 0; k1 < R; ++k1) {
            const float* d = pk + (rp * R + k1) * 4;
            cquad v; v.ra = d[0]; v.rb = d[1]; v.ia = d[2]; v.ib = d[3];
            smem[(rp * G + g) * P::PITCH + x0 + k1 * P::SEG] = v;
          }
      }
    }
  }
};

// --------------------------------------------------------------------------- //
// Phase B: X0-point packed codelet along w for (row pair, k1), IN PLACE in the thread's own k1 segment (the natural
// kx order is never materialised: Phase C addresses column kx = k1 + R k2 at k1*SEG + k2), so the phase has no
// inter-thread dependency and needs no barrier of its own; one round
// --------------------------------------------------------------------------- //
template <class P> B2S_HD void pk_phase_b(cquad* smem, int tid) {
  const int k1 = tid / P::ROWS;
  const int row = tid - k1 * P::ROWS;
  if (tid >= P::TASKS_B) return;
  cquad* seg = smem + row * P::PITCH + k1 * P::SEG;
  f2 re[P::X0], im[P::X0];
#pragma unroll
  for (int n = 0; n < P::X0; ++n) { const cquad v = seg[n]; re[n] = make_f2(v.ra, v.rb); im[n] = make_f2(v.ia, v.ib); }
  Dft<P::X0>::run(re, im);
#pragma unroll
  for (int k2 = 0; k2 < P::X0; ++k2) seg[k2] = make_q(re[k2], im[k2]);
}

// --------------------------------------------------------------------------- //
// Phase C: G-point packed codelet along h for (row pair, kx); the two results of a pair are output rows
// m + 8k and m + 2 + 8k of column kx and leave through the (scalar) epilogue functor
// --------------------------------------------------------------------------- //
template <class P, class Epi>
B2S_HD void pk_phase_c(const Epi& epi, const typename Epi::Ctx& ctx, const cquad* smem, int q, int task, float scale) {
  constexpr int G = P::G;
  const int rp = task / P::W, kx = task - rp * P::W;
  const int m = pk_m_of(2 * rp, q);                        // pk_m_of(2 * rp + 1, q) == m + 2
  const typename Epi::Ptr tpa = epi.task_ptr(ctx, m, kx), tpb = epi.task_ptr(ctx, m + 2, kx);
  typename Epi::template Pre<G, 1> prea, preb;
  epi.template prefetch<G, 1>(tpa, prea);
  epi.template prefetch<G, 1>(tpb, preb);
  f2 ur[G], ui[G];
  const cquad* src = smem + (rp * G) * P::PITCH + (kx % P::R) * P::SEG + kx / P::R;
#pragma unroll
  for (int g = 0; g < G; ++g) { const cquad v = src[g * P::PITCH]; ur[g] = make_f2(v.ra, v.rb); ui[g] = make_f2(v.ia, v.ib); }
  Dft<G>::run(ur, ui);
  const float s = ((q + kx) & 1) ? -scale : scale;         // (-1)^(ky+kx), ky = q mod 2
#pragma unroll
  for (int k = 0; k < G; ++k) {
    float re[1], im[1];
    re[0] = ur[k].x * s; im[0] = ui[k].x * s;
    epi.template store<G, 1>(tpa, k, re, im, prea);
    re[0] = ur[k].y * s; im[0] = ui[k].y * s;
    epi.template store<G, 1>(tpb, k, re, im, preb);
  }
}

// Phase C, two tasks per thread in ONE basic block: thread kx < W transforms column kx of BOTH row pairs, so that the
// two independent 25-point dependency chains (and the stores of the first with the arithmetic of the second) can be
// interleaved by the scheduler - with two warps per scheduler the phase is latency-bound, not issue-bound.
template <class P, class Epi>
B2S_HD void pk_phase_c_dual(const Epi& epi, const typename Epi::Ctx& ctx, const cquad* smem, int q, int kx, float scale) {
  constexpr int G = P::G;
  f2 ur[2][G], ui[2][G];
  const cquad* src = smem + (kx % P::R) * P::SEG + kx / P::R;
#pragma unroll
  for (int rp = 0; rp < 2; ++rp)
#pragma unroll
    for (int g = 0; g < G; ++g) { const cquad v = src[(rp * G + g) * P::PITCH]; ur[rp][g] = make_f2(v.ra, v.rb); ui[rp][g] = make_f2(v.ia, v.ib); }
  Dft<G>::run(ur[0], ui[0]);
  Dft<G>::run(ur[1], ui[1]);
  const float s = ((q + kx) & 1) ? -scale : scale;
#pragma unroll
  for (int rp = 0; rp < 2; ++rp) {
    const int m = pk_m_of(2 * rp, q);
    const typename Epi::Ptr tpa = epi.task_ptr(ctx, m, kx), tpb = epi.task_ptr(ctx, m + 2, kx);
    typename Epi::template Pre<G, 1> prea, preb;
    epi.template prefetch<G, 1>(tpa, prea);
    epi.template prefetch<G, 1>(tpb, preb);
#pragma unroll
    for (int k = 0; k < G; ++k) {
      float re[1], im[1];
      re[0] = ur[rp][k].x * s; im[0] = ui[rp][k].x * s;
      epi.template store<G, 1>(tpa, k, re, im, prea);
      re[0] = ur[rp][k].y * s; im[0] = ui[rp][k].y * s;
      epi.template store<G, 1>(tpb, k, re, im, preb);
    }
  }
}

// --------------------------------------------------------------------------- //
// Soft data consistency (varnet.py:281-282) with the reference rows STAGED in shared memory.
//
// The blend (1-m) k + m (k + v ref)/(1+v) needs the reference k-space on the sampled rows only (a quarter of them).
// Fetching it from Phase C costs ~100 registers of loads in flight per thread (spills) or exposed memory latency,
// a predicated blend per row costs a branch per row (the unrolled loop no longer fits the instruction cache), and a
// fix-up pass re-reads what was just stored.  Here Phase C runs ONE branch-free statement per output row,
//       out = (as * sigma) * u + beta * row[kx * stride],
// driven by a per-row table in shared memory: {as, stride, row}
//       unsampled row :  as = scale            stride 0   row -> 8 bytes of zeros
//       sampled row   :  as = scale / (1+v)    stride 8   row -> the reference row, staged in shared memory
// (beta = v/(1+v)).  The reference rows of a parity are copied into the staging buffer asynchronously (16-byte
// cp.async, global -> shared without registers) while Phase B runs.  The buffer holds CAP rows (a parity has 25 +- 3
// sampled rows at 4x acceleration); `row` is a GENERIC pointer, so a sampled row that did not fit simply points at the
// reference k-space in global memory - correct for any mask (fully sampled included), just slower for those rows.
// --------------------------------------------------------------------------- //
struct alignas(16) DcRow { float as; int stride; const char* row; };
static_assert(sizeof(DcRow) == 16, "one 128-bit shared-memory access per output row");

template <int H, int W> struct EpiDCStage {
  static constexpr bool FIXUP = false, STAGED = true;
  static B2S_HD int tab_index(int y) { return (y & 7) * (H / 8) + (y >> 3); }   // the H/8 rows of an m-block are contiguous
  static constexpr int SLOT_NONE = 0xff, OVF = 8 * 32;                          // slot bytes: one 32-byte line per m-block, then 2 overflow flags
  static B2S_HD int slot_index(int y) { return (y & 7) * 32 + (y >> 3); }
  cfloat* out; const cfloat* ref; const uint8_t* mask; const float* vptr; int C; long long hw;
  // every thread, before Phase A: this image's mask row -> aux[0, H)
  B2S_HD void stage_mask_row(long long image, uint8_t* aux, int tid, int nt) const {
    const uint8_t* src = mask + (image / C) * H;
    for (int y = tid; y < H; y += nt) aux[y] = src[y];
  }
  B2S_HD DcRow entry(long long image, int y, bool on, int slot, int cap, float scale, float inv1v, const cfloat* stage, const char* zero) const {
    DcRow e;
    e.as = on ? scale * inv1v : scale;
    e.stride = on ? 8 : 0;
    e.row = !on ? zero : (slot < cap ? reinterpret_cast<const char*>(stage + slot * W)
                                     : reinterpret_cast<const char*>(ref + image * hw + (long long)y * W));
    return e;
  }
  // host emulation of stage_issue
  template <class ST> void stage_host(long long image, int q, int cap, int cap_all, const uint8_t* aux, DcRow* tab, uint8_t* slots, cfloat* stage,
                                      void* bbase, const char* zero, float scale) const {
    const float v = *vptr, inv1v = 1.f / (1.f + v);
    int n = 0;
    for (int y = q; y < H; y += 2) {
      const bool on = aux[y] != 0;
      if (on && n < cap) for (int x = 0; x < W; ++x) stage[n * W + x] = ref[image * hw + (long long)y * W + x];
      else if (on && n < cap_all)
        for (int x = 0; x < W; ++x)
          *reinterpret_cast<cfloat*>(reinterpret_cast<char*>(bbase) + (n - cap) * ST::HOLE_STRIDE + ST::hole_base(x)) = ref[image * hw + (long long)y * W + x];
      tab[tab_index(y)] = entry(image, y, on, n, cap, scale, inv1v, stage, zero);
      slots[slot_index(y)] = (uint8_t)(on ? (n < cap_all ? n : cap_all) : SLOT_NONE);
      n += on;
    }
    slots[OVF + q] = (uint8_t)(n > cap_all);
  }
#if defined(__CUDACC__)
  // EVERY warp (NWARP == 8): the 100 rows of parity q are four ballot words; each warp repeats the four ballots (the
  // slot of a row is its rank among the sampled rows), warps 0-3 write the table entries of word `warp`, and warps w and
  // w + 4 share the copies of word (w & 3): 1600-byte rows as 100 x 16-byte cp.async per row, one per lane and trip.
  // (The per-SM bulk-copy unit needs ~600 cycles per 1600-byte cp.async.bulk - measured - and would not finish behind
  // Phase B.)  The caller completes the copies with cp.async.wait_all in front of the barrier that precedes Phase C.
  template <class ST, int NWARP> __device__ __forceinline__ void stage_issue(long long image, int q, const uint8_t* aux, DcRow* tab, uint8_t* slots,
                                                                             cfloat* stage, const void* bbase, const char* zero, float scale, int warp, int lane) const {
    constexpr int CAP = ST::CAP, CAP_ALL = ST::CAP_ALL;
    static_assert(NWARP == 8 && H / 2 <= 128, "four ballot words, two warps per word");
    constexpr int NR = H / 2, CHUNKS = W * 8 / 16;
    const float v = *vptr, inv1v = 1.f / (1.f + v);
    const int word = warp & 3;
    unsigned mine = 0;
    int base = 0, total = 0;                                 // sampled rows in the words before `word` / in the parity
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = 32 * i + lane;
      const bool on = (r < NR) && aux[q + 2 * r];
      const unsigned bits = __ballot_sync(0xffffffffu, on);
      if (i < word) base += __popc(bits);
      if (i == word) mine = bits;
      total += __popc(bits);
    }
    {
      const int r = 32 * word + lane;
      const bool on = (mine >> lane) & 1u;
      if (warp < 4 && r < NR) {
        const int y = q + 2 * r;
        const int sl = base + __popc(mine & ((1u << lane) - 1u));
        if (total > CAP_ALL) tab[tab_index(y)] = entry(image, y, on, sl, CAP, scale, inv1v, stage, zero);   // (general path only)
        slots[slot_index(y)] = (uint8_t)(on ? (sl < CAP_ALL ? sl : CAP_ALL) : SLOT_NONE);   // (CAP_ALL = sampled but not staged)
      }
      if (warp == 0 && lane == 0) slots[OVF + q] = (uint8_t)(total > CAP_ALL);
    }
    const char* src = reinterpret_cast<const char*>(ref + image * hw);
    unsigned b = mine;
    int sl = base, rank = 0;
    while (b) {
      const int bit = __ffs(b) - 1;
      b &= b - 1;
      if (sl < CAP_ALL && (rank & 1) == (warp >> 2)) {
        const char* s = src + (long long)(q + 2 * (32 * word + bit)) * (W * 8);
        if (sl < CAP) {
          const unsigned d = (unsigned)__cvta_generic_to_shared(stage + sl * W);
          for (int c = lane; c < CHUNKS; c += 32)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d + 16u * c), "l"(s + 16 * c) : "memory");
        } else {                                              // a row in the padding holes of B (see PkStage)
          const unsigned d = (unsigned)__cvta_generic_to_shared(bbase) + (unsigned)((sl - CAP) * ST::HOLE_STRIDE);
          for (int c = lane; c < CHUNKS; c += 32)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d + (unsigned)ST::hole_base(2 * c)), "l"(s + 16 * c) : "memory");
        }
      }
      ++sl; ++rank;
    }
  }
#endif
  B2S_HD void l2_prefetch_ahead(long long, int, int) const {}
  // thread y < H holds the mask byte of row y of `image` (fetched an image ahead): it pulls its own row into L2
  B2S_HD void l2_prefetch_row(long long image, int y, uint8_t on) const {
#if defined(__CUDA_ARCH__)
    if (on && y < H) {
      const char* p = reinterpret_cast<const char*>(ref + image * hw + (long long)y * W);
#pragma unroll
      for (int l = 0; l < (W * 8 + 127) / 128; ++l) asm volatile("prefetch.global.L2::evict_last [%0];" ::"l"(p + 128 * l));
    }
#else
    (void)image; (void)y; (void)on;
#endif
  }
};

// (re, im) at a SHARED-memory address if `on`, else (0, 0); one predicated LDS, no branch
B2S_HD cfloat lds_c_if(const char* p, bool on) {
  cfloat r = make_c(0.f, 0.f);
#if defined(__CUDA_ARCH__)
  asm volatile("{ .reg .pred p; setp.ne.u32 p, %3, 0; @p ld.shared.v2.f32 {%0, %1}, [%2]; }"
               : "+f"(r.x), "+f"(r.y) : "r"((unsigned)__cvta_generic_to_shared(p)), "r"((unsigned)on));
#else
  if (on) r = *reinterpret_cast<const cfloat*>(p);
#endif
  return r;
}

// Phase C with the staged soft-DC epilogue, every sampled row of the parity staged (the common case): the staging slots
// of the task's 2 x G rows come into registers with four 128-bit loads, and each row is
//   a = sampled ? s/(1+v) : s ;  out = u * a ;  sampled: out += beta * stage[slot][kx]
// with the load and the two FMAs predicated - no branch, no table access per row.
template <class P, int HH, int WW>
B2S_HD void pk_phase_c_dc_fast(const EpiDCStage<HH, WW>& epi, long long image, const cquad* smem, const uint8_t* slots, const cfloat* stage,
                               int q, int task, float scale, float inv1v, float beta, int cap0) {
  typedef PkStage<P> ST;
  constexpr int G = P::G, W = P::W;
  typedef EpiDCStage<HH, WW> E;
  static_assert(G <= 32, "one 32-byte line of slots per m-block");
  const int rp = task / W, kx = task - rp * W;
  const int m = pk_m_of(2 * rp, q);
  cfloat* po = epi.out + image * epi.hw + m * W + kx;
  uint32_t sw[2][8];
#pragma unroll
  for (int ab = 0; ab < 2; ++ab) {
    const u32x4* ps = reinterpret_cast<const u32x4*>(slots + (m + 2 * ab) * 32);
    const u32x4 lo = ps[0], hi = ps[1];
    sw[ab][0] = lo.x; sw[ab][1] = lo.y; sw[ab][2] = lo.z; sw[ab][3] = lo.w; sw[ab][4] = hi.x; sw[ab][5] = hi.y; sw[ab][6] = hi.z; sw[ab][7] = hi.w;
  }
  f2 ur[G], ui[G];
  const cquad* src = smem + (rp * G) * P::PITCH + (kx % P::R) * P::SEG + kx / P::R;
#pragma unroll
  for (int g = 0; g < G; ++g) { const cquad vv = src[g * P::PITCH]; ur[g] = make_f2(vv.ra, vv.rb); ui[g] = make_f2(vv.ia, vv.ib); }
  Dft<G>::run(ur, ui);
  const float s0 = ((q + kx) & 1) ? -scale : scale, s1 = s0 * inv1v;
  const char* sk = reinterpret_cast<const char*>(stage + kx);                                      // staged row sl < cap0: + sl * W * 8
  const char* hk = reinterpret_cast<const char*>(smem) + ST::hole_base(kx) - cap0 * ST::HOLE_STRIDE;   // row in the holes of B: + sl * HOLE_STRIDE
#pragma unroll
  for (int k = 0; k < G; ++k) {
#pragma unroll
    for (int ab = 0; ab < 2; ++ab) {
      const int sl = (int)((sw[ab][k >> 2] >> ((k & 3) * 8)) & 0xffu);
      const bool on = sl != E::SLOT_NONE;
      const float a = on ? s1 : s0;
      float re = (ab ? ur[k].y : ur[k].x) * a, im = (ab ? ui[k].y : ui[k].x) * a;
      // predicated 64-bit shared-memory load (zeros for an unsampled row): straight-line code - as C++ `if` this became
      // a branch with a convergence barrier per row and the unrolled loop no longer fitted the instruction cache
      const cfloat r = lds_c_if(sl < cap0 ? sk + sl * (W * 8) : hk + sl * ST::HOLE_STRIDE, on);
      re = fmaf(r.x, beta, re); im = fmaf(r.y, beta, im);
      cvec<1> o; o.v[0] = make_c(re, im);
      stv_stream<1>(po + (2 * ab + 8 * k) * W, o);
    }
  }
}

// Phase C with the staged soft-DC epilogue, general (some sampled rows of the parity did not fit the staging buffer):
// one table entry per row, whose generic pointer leads to shared OR global memory
template <class P, int HH, int WW>
B2S_HD void pk_phase_c_dc(const EpiDCStage<HH, WW>& epi, long long image, const cquad* smem, const DcRow* tab,
                          int q, int task, float beta) {
  constexpr int G = P::G, W = P::W;
  const int rp = task / W, kx = task - rp * W;
  const int m = pk_m_of(2 * rp, q);
  cfloat* po = epi.out + image * epi.hw + m * W + kx;
  f2 ur[G], ui[G];
  const cquad* src = smem + (rp * G) * P::PITCH + (kx % P::R) * P::SEG + kx / P::R;
#pragma unroll
  for (int g = 0; g < G; ++g) { const cquad vv = src[g * P::PITCH]; ur[g] = make_f2(vv.ra, vv.rb); ui[g] = make_f2(vv.ia, vv.ib); }
  Dft<G>::run(ur, ui);
  const float sigma = ((q + kx) & 1) ? -1.f : 1.f;         // (-1)^(ky+kx), ky = q mod 2
  const DcRow* ta = tab + m * G;                           // rows m + 8k; rows m + 2 + 8k follow 2 G entries later
#pragma unroll
  for (int k = 0; k < G; ++k) {
#pragma unroll
    for (int ab = 0; ab < 2; ++ab) {
      const DcRow e = ta[2 * ab * G + k];
      const cfloat r = *reinterpret_cast<const cfloat*>(e.row + kx * e.stride);
      const float a = e.as * sigma;
      cvec<1> o;
      o.v[0] = make_c(fmaf(r.x, beta, (ab ? ur[k].y : ur[k].x) * a), fmaf(r.y, beta, (ab ? ui[k].y : ui[k].x) * a));
      stv_stream<1>(po + (2 * ab + 8 * k) * W, o);
    }
  }
}

#if defined(__CUDACC__)
template <class P, class Pro, class Epi, int QD, int TT, bool CARRY, bool REVERSE = false>
__global__ void __launch_bounds__(P::NT, 1)
fft2_packed_kernel(const Pro pro, const Epi epi, const float scale, const int n_images, const int n_total, const int ahead_only, const int cdual) {
  extern __shared__ __align__(16) unsigned char b2s_smem_raw[];
  cquad* smem = reinterpret_cast<cquad*>(b2s_smem_raw);
  uint8_t* mrow = reinterpret_cast<uint8_t*>(smem + P::SMEM_QUADS);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mrow + 2 * P::AUX_BYTES);
  DcRow* dc_tab = reinterpret_cast<DcRow*>(tmem_slot + 4);                                   // (staged DC only)
  const char* dc_zero = reinterpret_cast<const char*>(dc_tab + P::H);
  uint8_t* dc_slots = reinterpret_cast<uint8_t*>(dc_tab + P::H + 1);
  cfloat* dc_stage = reinterpret_cast<cfloat*>(dc_slots + 8 * 32 + 16);
  const int tid = threadIdx.x;
  if constexpr (PkStaged<Epi>::value) { if (tid < 4) reinterpret_cast<float*>(dc_tab + P::H)[tid] = 0.f; }
  if (tid < 32) {                                            // 512 TMEM columns as the parking lot (one CTA per SM)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(tmem_slot)), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  pk_build_tables<P>(smem, tid, P::NT);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  ParkTmem park;
  {
    const int warp = tid >> 5;                               // lanes 32*(warp % 4) .., columns 256*(warp / 4) ..
    park.base = tmem_base + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)(256 * (warp >> 2));
  }
#ifdef B2S_PHASE_TIMING
  long long tprev = clock64();
#endif
  typedef PkPhaseA<P, Pro, QD, TT> PA;
  typename PA::Queue queue;
  auto image_of = [&](int n) -> long long { return REVERSE ? n_total - 1 - n : n; };
  if (CARRY && (int)blockIdx.x < n_images) PA::prefill(pro, pro.ctx(image_of(blockIdx.x)), tid, queue);
  long long prev_image = -1;
  if ((int)blockIdx.x < n_images) epi.l2_prefetch_ahead(image_of(blockIdx.x), tid, P::NT);
  uint8_t mask_next = 0;
  if constexpr (PkStaged<Epi>::value) { if ((int)blockIdx.x < n_images && tid < P::H) mask_next = epi.mask[(image_of(blockIdx.x) / epi.C) * P::H + tid]; }

#pragma unroll 1
  for (int n = blockIdx.x; n < n_images; n += gridDim.x) {
    const long long image = image_of(n);
    const int next = n + (int)gridDim.x;
    const bool has_next = next < n_images;
    const long long next_image = has_next ? image_of(next) : image;
    if (!CARRY) PA::prefill(pro, pro.ctx(image), tid, queue);
    if constexpr (Epi::FIXUP) epi.stage_mask_row(image, mrow + P::AUX_BYTES, tid, P::NT);
    if constexpr (PkStaged<Epi>::value) {                    // this image's mask row was fetched an image ago (no load latency in front of Phase A)
      if (tid < P::H) (mrow + P::AUX_BYTES)[tid] = mask_next;
      if (has_next && tid < P::H) mask_next = epi.mask[(next_image / epi.C) * P::H + tid];
    }
    (void)ahead_only;
    PA::template run<true>(pro, pro.ctx(image), pro.ctx(next_image), CARRY && has_next, smem, tid, queue, park);
    __syncthreads();
    B2S_TICK(0);
    if constexpr (PkStaged<Epi>::value)                      // reference rows of parity 0 -> staging buffer (behind Phase B)
      epi.template stage_issue<PkStage<P>, P::NT / 32>(image, 0, mrow + P::AUX_BYTES, dc_tab, dc_slots, dc_stage, smem, dc_zero, scale, tid >> 5, tid & 31);
    else epi.stage_mask(image, mrow, tid, P::NT);
    if constexpr (Epi::FIXUP) {
      // every Phase C (parity 1) store of the previous image has been issued (barrier inside Phase A): blend its
      // sampled odd rows (list slot 1), then list this image's rows (one warp; visible after the next barrier)
      if (prev_image >= 0) epi.template fixup<P::NT>(prev_image, mrow + P::AUX_BYTES, tid, 1);
      prev_image = image;
      B2S_TICK(4);
    }
    if (has_next) { pro.l2_prefetch(next_image, tid); epi.l2_prefetch_ahead(next_image, tid, P::NT); }
    if constexpr (PkStaged<Epi>::value) { if (has_next && ahead_only) epi.l2_prefetch_row(next_image, tid, mask_next); }
    if constexpr (!PkStaged<Epi>::value) { if (ahead_only == 0) epi.l2_prefetch(image, 0, 1, tid); }

#pragma unroll 1
    for (int q = 0; q < 2; ++q) {
      if (q == 1) {
        __syncthreads();                                     // every Phase C (q = 0) read of B is done
        if constexpr (Epi::FIXUP) { epi.template fixup<P::NT>(image, mrow + P::AUX_BYTES, tid, 0); B2S_TICK(4); }   // its even rows, just stored
        if constexpr (PkStaged<Epi>::value)                  // the buffer is free again: reference rows of parity 1 (behind unpark + Phase B)
          epi.template stage_issue<PkStage<P>, P::NT / 32>(image, 1, mrow + P::AUX_BYTES, dc_tab, dc_slots, dc_stage, smem, dc_zero, scale, tid >> 5, tid & 31);
        PA::unpark(smem, tid, park);
        __syncthreads();
        B2S_TICK(5);
      }
      pk_phase_b<P>(smem, tid);
      if constexpr (PkStaged<Epi>::value) asm volatile("cp.async.wait_all;" ::: "memory");   // this thread's staged rows have landed
      __syncthreads();
      B2S_TICK(1);
      if constexpr (Epi::FIXUP) {                            // (after the barrier: the fix-up above may still have been reading slot 1)
        if (q == 0) { epi.template stage_rows<2>(0, mrow + P::AUX_BYTES, tid, P::NT - 32, 0); epi.template stage_rows<2>(1, mrow + P::AUX_BYTES, tid, P::NT - 64, 1); }
      }
      if constexpr (PkStaged<Epi>::value) {
        const float v = *epi.vptr, inv1v = 1.f / (1.f + v), beta = v * inv1v;
        if (dc_slots[Epi::OVF + q] == 0) {
          for (int task = tid; task < P::TASKS_C; task += P::NT) pk_phase_c_dc_fast<P>(epi, image, smem, dc_slots, dc_stage, q, task, scale, inv1v, beta, PkStage<P>::CAP);
        } else {
          for (int task = tid; task < P::TASKS_C; task += P::NT) pk_phase_c_dc<P>(epi, image, smem, dc_tab, q, task, beta);
        }
      } else {
        const typename Epi::Ctx ectx = epi.ctx(image, mrow);
        if (cdual) { if (tid < P::W) pk_phase_c_dual<P>(epi, ectx, smem, q, tid, scale); }
        else for (int task = tid; task < P::TASKS_C; task += P::NT) pk_phase_c<P>(epi, ectx, smem, q, task, scale);
      }
      B2S_TICK(3);
    }
  }
  if constexpr (Epi::FIXUP) {
    __syncthreads();
    if (prev_image >= 0) epi.template fixup<P::NT>(prev_image, mrow + P::AUX_BYTES, tid, 1);
  }
  __syncthreads();
  if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
}
#endif

// Sequential emulation (tests/host_emul)
template <class P, class Pro, class Epi, int QD, int TT>
void fft2_packed_emulate(const Pro& pro, const Epi& epi, float scale, long long n_images, int emu_stage_cap = 0) {
  typedef PkPhaseA<P, Pro, QD, TT> PA;
  cquad* smem = new cquad[P::SMEM_QUADS];
  uint8_t* mrow = new uint8_t[2 * P::AUX_BYTES];
  typename PA::Queue* queues = new typename PA::Queue[P::NT];
  float* lot = new float[(size_t)P::NT * 256];
  DcRow* dc_tab = new DcRow[P::H];
  uint8_t* dc_slots = new uint8_t[8 * 32 + 16];
  const float dc_zero[4] = {0.f, 0.f, 0.f, 0.f};
  cfloat* dc_stage = new cfloat[(size_t)PkStage<P>::CAP * P::W];
  for (int i = 0; i < P::SMEM_QUADS; ++i) { cquad z; z.ra = z.rb = z.ia = z.ib = 0.f; smem[i] = z; }
  for (int tid = 0; tid < P::NT; ++tid) pk_build_tables<P>(smem, tid, P::NT);
  if (n_images > 0) for (int tid = 0; tid < P::NT; ++tid) PA::prefill(pro, pro.ctx(0), tid, queues[tid]);
  for (long long image = 0; image < n_images; ++image) {
    const bool has_next = image + 1 < n_images;
    if constexpr (PkStaged<Epi>::value) { for (int tid = 0; tid < P::NT; ++tid) epi.stage_mask_row(image, mrow + P::AUX_BYTES, tid, P::NT); }
    else { for (int tid = 0; tid < P::NT; ++tid) epi.stage_mask(image, mrow, tid, P::NT); }
    for (int tid = 0; tid < P::NT; ++tid) {
      ParkHost park{lot + (size_t)tid * 256};
      PA::template run<false>(pro, pro.ctx(image), pro.ctx(has_next ? image + 1 : image), has_next, smem, tid, queues[tid], park);
    }
    for (int q = 0; q < 2; ++q) {
      if (q == 1) for (int tid = 0; tid < P::NT; ++tid) { ParkHost park{lot + (size_t)tid * 256}; PA::unpark(smem, tid, park); }
      for (int tid = 0; tid < P::NT; ++tid) pk_phase_b<P>(smem, tid);
      if constexpr (PkStaged<Epi>::value) {
        // (emu_stage_cap = c: c rows in the staging buffer + c in the holes, so that small test masks reach every path)
        const int cap = emu_stage_cap > 0 && emu_stage_cap < PkStage<P>::CAP ? emu_stage_cap : PkStage<P>::CAP;
        const int cap_all = emu_stage_cap > 0 && emu_stage_cap < PkStage<P>::HOLE_ROWS ? 2 * emu_stage_cap : cap + PkStage<P>::HOLE_ROWS;
        epi.template stage_host<PkStage<P>>(image, q, cap, cap_all, mrow + P::AUX_BYTES, dc_tab, dc_slots, dc_stage, smem, reinterpret_cast<const char*>(dc_zero), scale);
        const float v = *epi.vptr, inv1v = 1.f / (1.f + v), beta = v * inv1v;
        for (int tid = 0; tid < P::NT; ++tid)
          for (int task = tid; task < P::TASKS_C; task += P::NT) {
            if (dc_slots[Epi::OVF + q] == 0) pk_phase_c_dc_fast<P>(epi, image, smem, dc_slots, dc_stage, q, task, scale, inv1v, beta, cap);
            else pk_phase_c_dc<P>(epi, image, smem, dc_tab, q, task, beta);
          }
      } else {
        const typename Epi::Ctx ectx = epi.ctx(image, mrow);
        for (int tid = 0; tid < P::NT; ++tid)
          for (int task = tid; task < P::TASKS_C; task += P::NT) pk_phase_c<P>(epi, ectx, smem, q, task, scale);
      }
    }
    if constexpr (Epi::FIXUP) {
      for (int tid = 0; tid < P::NT; ++tid) epi.stage_mask_row(image, mrow + P::AUX_BYTES, tid, P::NT);
      for (int tid = 0; tid < P::NT; ++tid) { epi.template stage_rows<2>(0, mrow + P::AUX_BYTES, tid, P::NT - 32, 0); epi.template stage_rows<2>(1, mrow + P::AUX_BYTES, tid, P::NT - 64, 1); }
      for (int tid = 0; tid < P::NT; ++tid) { epi.template fixup<P::NT>(image, mrow + P::AUX_BYTES, tid, 0); epi.template fixup<P::NT>(image, mrow + P::AUX_BYTES, tid, 1); }
    }
  }
  delete[] dc_stage;
  delete[] dc_slots;
  delete[] dc_tab;
  delete[] lot;
  delete[] queues;
  delete[] mrow;
  delete[] smem;
}

}  // namespace b2s
