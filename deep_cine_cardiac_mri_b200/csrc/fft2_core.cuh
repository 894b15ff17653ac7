// Fused centred 2-D FFT core ("half-split" design) — host/device phase functions.
//
// One work item produces HALF of the output rows of one H x W coil image (output
// rows ky == q mod 2) from ALL of its input rows, entirely on chip:
//
//   Phase A  load 8 rows {g + G*j} x R column groups {x0 + X0*i} per task straight
//            from global memory into registers with 128-bit accesses (NC = 2 adjacent
//            columns per thread; prologue functor = identity | S*x | k-space row
//            weight), radix-8 DIF butterfly over the 8 rows pruned to the 4 outputs
//            of parity q, row twiddle, radix-R butterfly over the R columns, column
//            twiddle, store to the shared buffer B.
//   Phase B  X0-point register codelet along w per (row, k1); rewrites the row in
//            natural kx order.
//   Phase C  G-point register codelet along h per (m-block, kx pair); results leave
//            through the epilogue functor with 128-bit accesses (plain store | mask /
//            soft-DC blend with the reference k-space | conj(S)-multiply + coil sum).
//
// The centring (fftshift/ifftshift of utils/fftc.py:59-110) is folded in: for
// even sizes fft2c(x) = chk * FFT2(chk * x) with chk = (-1)^(y+x); the input
// signs become an index relabel of the radix-8 outputs / a sign on the thread's
// twiddles, the output signs go into the per-thread scale.  The inverse
// transform runs the same forward machinery on re/im-swapped data.
//
// Everything here is `B2S_HD` so tests/host_emul can execute the exact phase
// code on the CPU (one phase = one loop over all thread ids).
#pragma once
#include <stdint.h>
#include "codelets.cuh"

// CTA barrier inside host/device phase code: the host emulation runs one phase for all thread ids in
// turn, so there it is a no-op.
#define B2S_OPAQUE(v) asm volatile("" : "+r"(v))
#if defined(__CUDA_ARCH__)
#define B2S_CTA_SYNC() __syncthreads()
#else
#define B2S_CTA_SYNC() ((void)0)
#endif

namespace b2s {

struct alignas(8) cfloat { float x, y; };
// NC adjacent complex values moved with one (64- or 128-bit) access
template <int NC> struct alignas(8 * NC) cvec { cfloat v[NC]; };

B2S_HD cfloat make_c(float a, float b) { cfloat r; r.x = a; r.y = b; return r; }

// exp(-2 pi i n / N) evaluated in double (called a few hundred times per CTA)
B2S_HD cfloat twiddle(int n, int N) {
  n %= N;
  if (n < 0) n += N;
  double s, c;
#if defined(__CUDA_ARCH__)
  sincospi(-2.0 * (double)n / (double)N, &s, &c);
#else
  const double a = -2.0 * 3.14159265358979323846264338327950288 * (double)n / (double)N;
  s = sin(a); c = cos(a);
#endif
  return make_c((float)c, (float)s);
}

// --------------------------------------------------------------------------- //
// plans
// --------------------------------------------------------------------------- //
template <int H_, int W_, int NT_ = 256, int NC_ = 1, int FOLD_ = 2, int NCC_ = NC_> struct Plan;

// 200 x 200 (dataset crop, data/mri_data.py:273-277): h = 8*25, w = 5*40.
// FOLD 2: half split, 170 KB of shared memory, one 256-thread CTA per SM.
// FOLD 4: quarter split, 85 KB, two 128-thread CTAs per SM (their phases overlap) at the price of 4x L2 reads.
template <int NT_, int NC_, int FOLD_, int NCC_> struct Plan<200, 200, NT_, NC_, FOLD_, NCC_> {
  static constexpr int H = 200, W = 200;
  static constexpr int G = 25;        // Phase C codelet size, H = 8*G
  static constexpr int R = 5;         // Phase A column radix, W = R*X0
  static constexpr int X0 = 40;       // Phase B codelet size
  static constexpr int SEG = 41;      // padded k1-segment pitch (complex) written by Phase A
  static constexpr int PITCH = 213;   // row pitch of B (complex); 213 = 5 mod 16 keeps Phase B conflict-free
  static constexpr int NT = NT_;      // threads per CTA
  static constexpr int NC = NC_;      // adjacent columns per thread in Phase A (global access = 8*NC bytes)
  static constexpr int NCC = NCC_;    // ... and in Phase C
  static constexpr int FOLD = FOLD_;  // work items per image: item q produces output rows ky == q (mod FOLD)
  static constexpr int CTAS = (FOLD_ == 4) ? 2 : 1;   // resident CTAs per SM
};

// 256 x 256 (BASELINE.json configs[4]): h = 8*32, w = 8*32.  Half an image (128 x 256 complex) does not fit
// in shared memory, so the image is split in four: item q keeps 2 of the 8 radix-8 outputs (ky == q mod 4).
template <int NT_, int NC_, int FOLD_, int NCC_> struct Plan<256, 256, NT_, NC_, FOLD_, NCC_> {
  static constexpr int H = 256, W = 256;
  static constexpr int G = 32, R = 8, X0 = 32;
  static constexpr int SEG = 33;      // 8 segments of 32 (+1 pad)
  static constexpr int PITCH = 265;   // odd: lanes along rows stay conflict-free in Phase B
  static constexpr int NT = NT_, NC = NC_, NCC = NCC_;
  static constexpr int FOLD = 4;
  static constexpr int CTAS = 1;
  static_assert(FOLD_ == 4, "256 x 256 only fits as a quarter split");
};

template <class P> struct Derived {
  static constexpr int NKEEP = 8 / P::FOLD;                     // radix-8 outputs (m-blocks) kept per item
  static constexpr int ROWS = NKEEP * P::G;                     // rows of B
  static constexpr int SG = P::G & 1;                           // (-1)^(G j) relabel
  static constexpr int B_ELEMS = ROWS * P::PITCH;               // complex elements
  static constexpr int TW_OFF = B_ELEMS;                        // TW[W]
  static constexpr int TH_OFF = TW_OFF + P::W;                  // TH[FOLD][G][NKEEP] (every q)
  static constexpr int SMEM_ELEMS = TH_OFF + 8 * P::G;
  static constexpr int MASK_BYTES = (P::H + 15) / 16 * 16;      // this item's mask row (uint8) after the tables
  static constexpr int AUX_BYTES = 2 * MASK_BYTES + 16;         // row-fix-up epilogues: mask row + sampled-row list + count,
  static constexpr int SMEM_BYTES = SMEM_ELEMS * 8 + 2 * AUX_BYTES;   // double-buffered (current and previous work item)
  static constexpr int XP = P::X0 / P::NC;                      // column groups per row group
  static constexpr int TASKS_A = P::G * XP;
  static constexpr int KXP = P::W / P::NCC;
  static constexpr int TASKS_C = NKEEP * KXP;
  static constexpr int RPR = (P::NT / P::R) < ROWS ? (P::NT / P::R) : ROWS;   // rows per Phase-B round
  static constexpr int ROUNDS_B = (ROWS + RPR - 1) / RPR;
  static_assert(P::H == 8 * P::G && P::W == P::R * P::X0, "bad plan");
  static_assert((P::X0 & 1) == 0, "X0 must be even (sign folding)");
  static_assert(P::X0 % P::NC == 0 && P::W % P::NCC == 0, "NC must divide X0, NCC must divide W");
  static_assert(P::R * P::SEG <= P::PITCH && P::W <= P::PITCH, "pitch too small");
};

// output row residue (mod 8) handled by m-block r of item q: radix-8 output m' = q + FOLD*r, relabelled by
// the (-1)^(G j) part of the input checkerboard
template <class P> B2S_HD int m_of(int r, int q) { return (P::FOLD * r + q + 4 * Derived<P>::SG) & 7; }

// --------------------------------------------------------------------------- //
// tables (per CTA, in shared memory)
// --------------------------------------------------------------------------- //
template <class P> B2S_HD void build_tables(cfloat* smem, int tid, int nthreads) {
  using D = Derived<P>;
  for (int n = tid; n < P::W; n += nthreads) smem[D::TW_OFF + n] = twiddle(n, P::W);
  for (int e = tid; e < 8 * P::G; e += nthreads) {
    const int q = e / (P::G * D::NKEEP), g = (e / D::NKEEP) % P::G, r = e % D::NKEEP;
    cfloat t = twiddle(g * m_of<P>(r, q), P::H);
    if (g & 1) { t.x = -t.x; t.y = -t.y; }                      // (-1)^g of the input checkerboard
    smem[D::TH_OFF + e] = t;
  }
}

// --------------------------------------------------------------------------- //
// Phase A — software-pipelined over "steps".
//
// A thread owns TPT tasks (g, x0) per work item; a task is R steps (one step = the 8 rows
// {g + G*j} of column group i).  The raw global loads of step s + QD are issued right after
// step s has been consumed, into a register queue of QD slots that stays live across Phase B,
// Phase C and the work-item boundary: global/L2 latency is hidden behind the register
// codelets of all three phases instead of being exposed at the head of every Phase A
// (shared memory is full, registers are the only landing zone).
// --------------------------------------------------------------------------- //
template <class P, class Pro> struct PhaseA {
  using D = Derived<P>;
  static constexpr int G = P::G, R = P::R, X0 = P::X0, NC = P::NC, NT = P::NT;
  static constexpr int TPT = (D::TASKS_A + NT - 1) / NT;       // tasks per thread and item
  static constexpr int STEPS = TPT * R;
  static constexpr int NK = D::NKEEP;
  static constexpr int QD = Pro::template qdepth<R, false, NC>();   // steps in flight
  static constexpr int TT = ((2 * R) % QD == 0) ? 2 : 4;            // tasks per loop trip (queue slots stay compile-time)
  static_assert(STEPS % QD == 0 && (TT * R) % QD == 0 && TPT % TT == 0, "queue depth must divide one trip's steps");
  typedef typename Pro::template Unit<NC> Unit;
  struct Queue { Unit u[QD]; };

  static B2S_HD bool task_of(int tid, int k, int& g, int& x0) {
    const int task = tid + k * NT;
    g = task / D::XP; x0 = (task - g * D::XP) * NC;
    return task < D::TASKS_A;
  }

  // raw loads of step s (task s / R, column group s % R) of the item described by ctx
  static B2S_HD void issue(const Pro& pro, const typename Pro::Ctx& ctx, int tid, int s, Unit& u) {
    int g, x0;
    if (!task_of(tid, s / R, g, x0)) return;
    pro.template fetch<NC, G * P::W>(ctx, g, g * P::W + x0 + X0 * (s % R), u);
  }

  static B2S_HD void prefill(const Pro& pro, const typename Pro::Ctx& ctx, int tid, Queue& q) {
#pragma unroll
    for (int s = 0; s < QD; ++s) issue(pro, ctx, tid, s, q.u[s]);
  }

  // consume the queue for one work item (half q), refilling it from this item and then the next
  template <bool SYNC_FIRST = false>
  static B2S_HD void run(const Pro& pro, const typename Pro::Ctx& ctx, const typename Pro::Ctx& next, bool has_next,
                         cfloat* smem, int q, int tid, Queue& qu) {
    const float h = 0.70710678118654752440f;
    float ur[NC][NK][R], ui[NC][NK][R];     // [column][m-block r][column-group index i]
#pragma unroll 1
    for (int kp = 0; kp < TPT; kp += TT)    // TT tasks per trip: queue slots stay compile-time, code stays small
#pragma unroll
    for (int u = 0; u < TT * R; ++u) {
      const int s = kp * R + u;
      const int k = kp + u / R, i = u % R, slot = u % QD;
      int g, x0;
      const bool valid = task_of(tid, k, g, x0);
      // ---- consume step s: fold the row pairs (j, j+4), W8 twiddles, (pruned) radix-4 over j
      const int q0 = q & 1;                   // parity of the kept radix-8 outputs
      float fr[NC][4], fi[NC][4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float ar[NC], ai[NC], br[NC], bi[NC];
        pro.template value<NC>(qu.u[slot], j, ar, ai);
        pro.template value<NC>(qu.u[slot], j + 4, br, bi);
#pragma unroll
        for (int n = 0; n < NC; ++n) {
          if (q0 == 0) {
            fr[n][j] = ar[n] + br[n]; fi[n][j] = ai[n] + bi[n];
          } else {
            const float dr = ar[n] - br[n], di = ai[n] - bi[n];
            if (j == 0)      { fr[n][j] = dr;              fi[n][j] = di; }
            else if (j == 1) { fr[n][j] = (dr + di) * h;   fi[n][j] = (di - dr) * h; }
            else if (j == 2) { fr[n][j] = di;              fi[n][j] = -dr; }
            else             { fr[n][j] = (di - dr) * h;   fi[n][j] = -(dr + di) * h; }
          }
        }
      }
#pragma unroll
      for (int n = 0; n < NC; ++n) {
        if (NK == 4) {                         // half split: all four outputs of this parity
          dft4(fr[n], fi[n]);
#pragma unroll
          for (int r = 0; r < NK; ++r) { ur[n][r][i] = fr[n][r]; ui[n][r][i] = fi[n][r]; }
        } else {                               // quarter split: radix-4 outputs q1 and q1 + 2 only
          const int q1 = (q >> 1) & 1;
          float hr[2], hi[2];
          if (q1 == 0) {
            hr[0] = fr[n][0] + fr[n][2]; hi[0] = fi[n][0] + fi[n][2];
            hr[1] = fr[n][1] + fr[n][3]; hi[1] = fi[n][1] + fi[n][3];
          } else {                             // (f_j - f_{j+2}) * w4^j, w4 = -i
            hr[0] = fr[n][0] - fr[n][2]; hi[0] = fi[n][0] - fi[n][2];
            hr[1] = fi[n][1] - fi[n][3]; hi[1] = -(fr[n][1] - fr[n][3]);
          }
          ur[n][0][i] = hr[0] + hr[1]; ui[n][0][i] = hi[0] + hi[1];
          ur[n][NK - 1][i] = hr[0] - hr[1]; ui[n][NK - 1][i] = hi[0] - hi[1];
        }
      }
      // ---- refill the slot with step s + QD (of this item, else of the next one)
      if (s + QD < STEPS) issue(pro, ctx, tid, s + QD, qu.u[slot]);
      else if (has_next) issue(pro, next, tid, s + QD - STEPS, qu.u[slot]);
      // ---- last column group of the task: row twiddles, radix-R over i, column twiddles, store
      if (SYNC_FIRST && u == R - 1) {          // first write into B of this item: every warp must have left
        if (kp == 0) B2S_CTA_SYNC();           // the previous item's Phase C (its ragged last round overlaps
      }                                        // the loads issued above instead of idling 7 of 8 warps)
      if (i == R - 1 && valid) {
        B2S_OPAQUE(g);                                // (same: table and B addresses stay per-task work)
        cfloat th[NK];
#pragma unroll
        for (int r = 0; r < NK; ++r) th[r] = smem[D::TH_OFF + (q * G + g) * NK + r];
#pragma unroll
        for (int n = 0; n < NC; ++n) {
          int x = x0 + n;
          B2S_OPAQUE(x);                              // keep the twiddle addresses out of the persistent
                                                      // loop's invariants (ptxas hoists and then spills them)
          const float sx = (x & 1) ? -1.f : 1.f;      // column parity of the input checkerboard
          cfloat tw[R];
#pragma unroll
          for (int k1 = 1; k1 < R; ++k1) tw[k1] = smem[D::TW_OFF + (x * k1) % P::W];
#pragma unroll
          for (int r = 0; r < NK; ++r) {
            const float tx = th[r].x * sx, ty = th[r].y * sx;
#pragma unroll
            for (int ii = 0; ii < R; ++ii) {
              const float a = ur[n][r][ii], b = ui[n][r][ii];
              ur[n][r][ii] = a * tx - b * ty;
              ui[n][r][ii] = a * ty + b * tx;
            }
            Dft<R>::run(ur[n][r], ui[n][r]);
            cfloat* dst = smem + (r * G + g) * P::PITCH + x;
            dst[0] = make_c(ur[n][r][0], ui[n][r][0]);
#pragma unroll
            for (int k1 = 1; k1 < R; ++k1) {
              const float a = ur[n][r][k1], b = ui[n][r][k1];
              dst[k1 * P::SEG] = make_c(a * tw[k1].x - b * tw[k1].y, a * tw[k1].y + b * tw[k1].x);
            }
          }
        }
      }
    }
  }
};

// --------------------------------------------------------------------------- //
// Phase A, paired (2-CTA cluster, half split only).
//
// The two CTAs of a cluster work on the SAME image; CTA `rank` owns output parity q = rank (its B buffer,
// Phases B and C are unchanged).  In Phase A each CTA loads only HALF of the columns (x0 in
// [rank*X0/2, (rank+1)*X0/2)) of every row, so every input element - and for sens_expand every S*x
// product - is fetched/computed once per image instead of once per half: the L2 -> SM ingest, which bounds
// the unpaired kernel, is halved.  From its loads a thread forms the folds of BOTH parities
// (a_j + a_{j+4} and (a_j - a_{j+4}) w8^j), finishes both, and stores parity `rank` into its own B and the
// other parity into the partner's B through distributed shared memory (`remote` = the partner's buffer
// mapped into this CTA's address space; plain generic stores).
// --------------------------------------------------------------------------- //
template <class P, class Pro> struct PhaseA2 {
  using D = Derived<P>;
  static constexpr int G = P::G, R = P::R, X0 = P::X0, NC = P::NC, NT = P::NT;
  static_assert(P::FOLD == 2 && X0 % (2 * NC) == 0, "paired Phase A needs the half split");
  static constexpr int XH = X0 / 2 / NC;                        // column groups per row group and CTA
  static constexpr int TASKS = G * XH;
  static constexpr int TPT = (TASKS + NT - 1) / NT;
  static constexpr int STEPS = TPT * R;
  static constexpr int QD = Pro::template qdepth<R, true, NC>();
  static_assert(STEPS % QD == 0 && (2 * R) % QD == 0 && TPT % 2 == 0, "queue depth must divide two tasks' steps");
  typedef typename Pro::template Unit<NC> Unit;
  struct Queue { Unit u[QD]; };

  static B2S_HD bool task_of(int tid, int k, int rank, int& g, int& x0) {
    const int task = tid + k * NT;
    g = task / XH; x0 = (rank * XH + (task - g * XH)) * NC;
    return task < TASKS;
  }
  static B2S_HD void issue(const Pro& pro, const typename Pro::Ctx& ctx, int tid, int rank, int s, Unit& u) {
    int g, x0;
    if (!task_of(tid, s / R, rank, g, x0)) return;
    pro.template fetch<NC, G * P::W>(ctx, g, g * P::W + x0 + X0 * (s % R), u);
  }
  static B2S_HD void prefill(const Pro& pro, const typename Pro::Ctx& ctx, int tid, int rank, Queue& q) {
#pragma unroll
    for (int s = 0; s < QD; ++s) issue(pro, ctx, tid, rank, s, q.u[s]);
  }

  // `own` = this CTA's shared memory (tables + B of parity `rank`), `remote` = the partner's B
  static B2S_HD void run(const Pro& pro, const typename Pro::Ctx& ctx, const typename Pro::Ctx& next, bool has_next,
                         cfloat* own, cfloat* remote, int rank, int tid, Queue& qu) {
    const float h = 0.70710678118654752440f;
    float ur[2][NC][4][R], ui[2][NC][4][R];   // [parity][column][m-block r][column-group index i]
#pragma unroll 1
    for (int kp = 0; kp < TPT; kp += 2)
#pragma unroll
    for (int u = 0; u < 2 * R; ++u) {
      const int s = kp * R + u;
      const int k = kp + u / R, i = u % R, slot = u % QD;
      int g, x0;
      const bool valid = task_of(tid, k, rank, g, x0);
#pragma unroll
      for (int n = 0; n < NC; ++n) {
        float er[4], ei[4], orr[4], oi[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float ar[NC], ai[NC], br[NC], bi[NC];
          pro.template value<NC>(qu.u[slot], j, ar, ai);
          pro.template value<NC>(qu.u[slot], j + 4, br, bi);
          er[j] = ar[n] + br[n]; ei[j] = ai[n] + bi[n];
          const float dr = ar[n] - br[n], di = ai[n] - bi[n];
          if (j == 0)      { orr[j] = dr;              oi[j] = di; }
          else if (j == 1) { orr[j] = (dr + di) * h;   oi[j] = (di - dr) * h; }
          else if (j == 2) { orr[j] = di;              oi[j] = -dr; }
          else             { orr[j] = (di - dr) * h;   oi[j] = -(dr + di) * h; }
        }
        dft4(er, ei);
        dft4(orr, oi);
#pragma unroll
        for (int r = 0; r < 4; ++r) { ur[0][n][r][i] = er[r]; ui[0][n][r][i] = ei[r]; ur[1][n][r][i] = orr[r]; ui[1][n][r][i] = oi[r]; }
      }
      if (s + QD < STEPS) issue(pro, ctx, tid, rank, s + QD, qu.u[slot]);
      else if (has_next) issue(pro, next, tid, rank, s + QD - STEPS, qu.u[slot]);
      if (i == R - 1 && valid) {
#pragma unroll
        for (int n = 0; n < NC; ++n) {
          const int x = x0 + n;
          const float sx = (x & 1) ? -1.f : 1.f;
          cfloat tw[R];
#pragma unroll
          for (int k1 = 1; k1 < R; ++k1) tw[k1] = own[D::TW_OFF + (x * k1) % P::W];
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            cfloat* base = (q == rank) ? own : remote;
#pragma unroll
            for (int r = 0; r < 4; ++r) {
              const cfloat th = own[D::TH_OFF + (q * G + g) * 4 + r];
              const float tx = th.x * sx, ty = th.y * sx;
#pragma unroll
              for (int ii = 0; ii < R; ++ii) {
                const float a = ur[q][n][r][ii], b = ui[q][n][r][ii];
                ur[q][n][r][ii] = a * tx - b * ty;
                ui[q][n][r][ii] = a * ty + b * tx;
              }
              Dft<R>::run(ur[q][n][r], ui[q][n][r]);
              cfloat* dst = base + (r * G + g) * P::PITCH + x;
              dst[0] = make_c(ur[q][n][r][0], ui[q][n][r][0]);
#pragma unroll
              for (int k1 = 1; k1 < R; ++k1) {
                const float a = ur[q][n][r][k1], b = ui[q][n][r][k1];
                dst[k1 * P::SEG] = make_c(a * tw[k1].x - b * tw[k1].y, a * tw[k1].y + b * tw[k1].x);
              }
            }
          }
        }
      }
    }
  }
};

// --------------------------------------------------------------------------- //
// Phase B: read (segment layout) -> X0-point codelet ; write (natural layout)
// --------------------------------------------------------------------------- //
template <class P> struct PhaseBRegs { float re[P::X0], im[P::X0]; int row, k1; bool valid; };

template <class P> B2S_HD void phase_b_read(const cfloat* smem, int round, int local, PhaseBRegs<P>& s) {
  using D = Derived<P>;
  s.k1 = local / D::RPR;
  s.row = round * D::RPR + (local - s.k1 * D::RPR);
  s.valid = (s.k1 < P::R) && (s.row < D::ROWS);
  if (!s.valid) return;
  const cfloat* src = smem + s.row * P::PITCH + s.k1 * P::SEG;
#pragma unroll
  for (int n = 0; n < P::X0; ++n) { const cfloat v = src[n]; s.re[n] = v.x; s.im[n] = v.y; }
  Dft<P::X0>::run(s.re, s.im);
}

template <class P> B2S_HD void phase_b_write(cfloat* smem, const PhaseBRegs<P>& s) {
  if (!s.valid) return;
  cfloat* dst = smem + s.row * P::PITCH + s.k1;
#pragma unroll
  for (int k2 = 0; k2 < P::X0; ++k2) dst[P::R * k2] = make_c(s.re[k2], s.im[k2]);
}

// --------------------------------------------------------------------------- //
// Phase C
// --------------------------------------------------------------------------- //
template <class P, class Epi>
B2S_HD void phase_c(const Epi& epi, const typename Epi::Ctx& ctx, const cfloat* smem, int q, int task,
                    float scale) {
  using D = Derived<P>;
  constexpr int G = P::G, NC = P::NCC;
  const int r = task / D::KXP, kx = (task - r * D::KXP) * NC;
  const int m = m_of<P>(r, q);
  const typename Epi::Ptr tp = epi.task_ptr(ctx, m, kx);   // rows m + 8*k: constant offsets 8*k*W
  typename Epi::template Pre<G, NC> pre;
  epi.template prefetch<G, NC>(tp, pre);                   // epilogue operands in flight behind the codelet
  float ur[NC][G], ui[NC][G];
#pragma unroll
  for (int n = 0; n < NC; ++n) {
    const cfloat* src = smem + (r * G) * P::PITCH + kx + n;
#pragma unroll
    for (int g = 0; g < G; ++g) { const cfloat v = src[g * P::PITCH]; ur[n][g] = v.x; ui[n][g] = v.y; }
    Dft<G>::run(ur[n], ui[n]);
  }
#pragma unroll
  for (int k = 0; k < G; ++k) {
    float re[NC], im[NC];
#pragma unroll
    for (int n = 0; n < NC; ++n) {
      const float s = ((q + kx + n) & 1) ? -scale : scale;    // (-1)^(ky+kx), ky = q mod 2
      re[n] = ur[n][k] * s; im[n] = ui[n][k] * s;
    }
    epi.template store<G, NC>(tp, k, re, im, pre);
  }
}

// global sign (-1)^(H/2 + W/2) of the centred transform for even sizes
template <class P> B2S_HD float centre_sign() { return (((P::H / 2) + (P::W / 2)) & 1) ? -1.f : 1.f; }

}  // namespace b2s
