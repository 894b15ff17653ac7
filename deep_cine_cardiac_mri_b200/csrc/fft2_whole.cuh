// Whole-image variant of the fused centred 2-D FFT (200 x 200): one work item = one coil image.
//
// The half-split kernel (fft2_kernel.cuh) makes every SM ingest each image twice, because only HALF of the
// intermediate (the radix-8 outputs of one parity, 160 KB) fits in shared memory next to nothing else.  On B200
// there is a second on-chip memory that a CUDA-core kernel normally leaves idle: the 256 KB of TENSOR MEMORY
// (TMEM, 128 lanes x 512 columns x 32 bit).  It is only reachable through tcgen05.st / tcgen05.ld, lane l of warp w
// can only touch TMEM lane 32*(w % 4) + l - which is exactly what a per-thread parking lot needs.  So:
//
//   Phase A   loads every input element ONCE (and forms every S*x product once), runs the FULL radix-8 over the
//             rows, and finishes both parities: parity 0 goes to the shared buffer B as before, parity 1 (40 floats
//             per task, 160 per thread, 160 KB per image) is PARKED in the thread's own TMEM columns.
//   B, C      on parity 0 (unchanged code).
//   unpark    each thread reads its 160 floats back from TMEM and writes them where Phase A would have put them.
//   B, C      on parity 1.
//
// No tensor-core instruction is issued; TMEM is used purely as 160 KB of extra scratch-pad.  Host emulation
// (tests/host_emul) parks in a plain per-thread array.
#pragma once
#include "fft2_kernel.cuh"

namespace b2s {

// --------------------------------------------------------------------------- //
// parking lot
// --------------------------------------------------------------------------- //
#if defined(__CUDACC__)
struct ParkTmem {
  uint32_t base;                                   // TMEM address of this thread's column 0 (lane bits included)
  // 8 consecutive columns of the calling lane's TMEM row; warp-collective (.sync.aligned)
  __device__ __forceinline__ void st8(int col, const float* v) const {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"r"(base + (uint32_t)col), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
                   "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])),
                   "r"(__float_as_uint(v[7])) : "memory");
  }
  __device__ __forceinline__ void ld8(int col, float* v) const {
    uint32_t r0, r1, r2, r3, r4, r5, r6, r7;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3), "=r"(r4), "=r"(r5), "=r"(r6), "=r"(r7) : "r"(base + (uint32_t)col) : "memory");
    v[0] = __uint_as_float(r0); v[1] = __uint_as_float(r1); v[2] = __uint_as_float(r2); v[3] = __uint_as_float(r3);
    v[4] = __uint_as_float(r4); v[5] = __uint_as_float(r5); v[6] = __uint_as_float(r6); v[7] = __uint_as_float(r7);
  }
  template <int N> __device__ __forceinline__ void store(int col, const float (&v)[N]) const {
    static_assert(N % 8 == 0, "park in multiples of 8 columns");
#pragma unroll
    for (int i = 0; i < N; i += 8) st8(col + i, v + i);
  }
  template <int N> __device__ __forceinline__ void load(int col, float (&v)[N]) const {
#pragma unroll
    for (int i = 0; i < N; i += 8) ld8(col + i, v + i);
  }
  __device__ __forceinline__ void wait_st() const { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
  __device__ __forceinline__ void wait_ld() const { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
};
#endif
struct ParkHost {
  float* p;                                        // this thread's parking array
  template <int N> void store(int col, const float (&v)[N]) const { for (int i = 0; i < N; ++i) p[col + i] = v[i]; }
  template <int N> void load(int col, float (&v)[N]) const { for (int i = 0; i < N; ++i) v[i] = p[col + i]; }
  void wait_st() const {}
  void wait_ld() const {}
};

// --------------------------------------------------------------------------- //
// Phase A, whole image: full radix-8 over the rows, both parities finished per task
// --------------------------------------------------------------------------- //
template <class P, class Pro, int QD_, int TT_> struct PhaseAW {
  using D = Derived<P>;
  static constexpr int G = P::G, R = P::R, X0 = P::X0, NT = P::NT;
  static_assert(P::FOLD == 2 && P::NC == 1, "whole-image Phase A: half-sized B, one column per thread");
  static constexpr int TASKS = G * X0;
  static constexpr int TPT = (TASKS + NT - 1) / NT;        // tasks per thread and image
  static constexpr int STEPS = TPT * R;
  static constexpr int QD = QD_, TT = TT_;
  static constexpr int PARK = 2 * 4 * R;                   // floats parked per task (the parity-1 half)
  static_assert(PARK % 8 == 0 && TPT * PARK <= 256, "parking lot: 256 TMEM columns per thread (8 warps share 4 lane quarters)");
  static_assert(STEPS % QD == 0 && (TT * R) % QD == 0 && TPT % TT == 0, "queue depth must divide one trip's steps");
  typedef typename Pro::template Unit<1> Unit;
  struct Queue { Unit u[QD]; };

  static B2S_HD bool task_of(int tid, int k, int& g, int& x0) {
    const int task = tid + k * NT;
    g = task / X0; x0 = task - g * X0;
    return task < TASKS;
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

  // row twiddle, radix-R over the column groups, column twiddles of one (parity, m-block) of a task
  static B2S_HD void finish(float (&ar)[R], float (&ai)[R], const cfloat th, float sx, const cfloat (&tw)[R]) {
    const float tx = th.x * sx, ty = th.y * sx;
#pragma unroll
    for (int ii = 0; ii < R; ++ii) {
      const float a = ar[ii], b = ai[ii];
      ar[ii] = a * tx - b * ty;
      ai[ii] = a * ty + b * tx;
    }
    Dft<R>::run(ar, ai);
#pragma unroll
    for (int k1 = 1; k1 < R; ++k1) {
      const float a = ar[k1], b = ai[k1];
      ar[k1] = a * tw[k1].x - b * tw[k1].y;
      ai[k1] = a * tw[k1].y + b * tw[k1].x;
    }
  }

  template <bool SYNC_FIRST, class Park>
  static B2S_HD void run(const Pro& pro, const typename Pro::Ctx& ctx, const typename Pro::Ctx& next, bool has_next,
                         cfloat* smem, int tid, Queue& qu, const Park& park) {
    const float h = 0.70710678118654752440f;
    float ur[2][4][R], ui[2][4][R];         // [parity][m-block r][column-group index i]
#pragma unroll 1
    for (int kp = 0; kp < TPT; kp += TT)
#pragma unroll
    for (int u = 0; u < TT * R; ++u) {
      const int s = kp * R + u;
      const int k = kp + u / R, i = u % R, slot = u % QD;
      int g, x0;
      const bool valid = task_of(tid, k, g, x0);
      // ---- consume step s: full radix-8 DIF over the 8 rows: even outputs from a_j + a_{j+4}, odd outputs from
      //      (a_j - a_{j+4}) w8^j, then one radix-4 each
      {
        float er[4], ei[4], orr[4], oi[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float ar[1], ai[1], br[1], bi[1];
          pro.template value<1>(qu.u[slot], j, ar, ai);
          pro.template value<1>(qu.u[slot], j + 4, br, bi);
          er[j] = ar[0] + br[0]; ei[j] = ai[0] + bi[0];
          const float dr = ar[0] - br[0], di = ai[0] - bi[0];
          if (j == 0)      { orr[j] = dr;              oi[j] = di; }
          else if (j == 1) { orr[j] = (dr + di) * h;   oi[j] = (di - dr) * h; }
          else if (j == 2) { orr[j] = di;              oi[j] = -dr; }
          else             { orr[j] = (di - dr) * h;   oi[j] = -(dr + di) * h; }
        }
        dft4(er, ei);
        dft4(orr, oi);
#pragma unroll
        for (int r = 0; r < 4; ++r) { ur[0][r][i] = er[r]; ui[0][r][i] = ei[r]; ur[1][r][i] = orr[r]; ui[1][r][i] = oi[r]; }
      }
      // ---- refill the slot with step s + QD (of this image, else of the next one)
      if (s + QD < STEPS) issue(pro, ctx, tid, s + QD, qu.u[slot]);
      else if (has_next) issue(pro, next, tid, s + QD - STEPS, qu.u[slot]);
      if (SYNC_FIRST && u == R - 1) {          // first write into B of this image: every warp must have left the
        if (kp == 0) B2S_CTA_SYNC();           // previous image's last Phase C
      }
      // ---- last column group of the task: finish both parities
      if (i == R - 1) {
        int gg = valid ? g : 0, x = valid ? x0 : 0;
        B2S_OPAQUE(gg);
        B2S_OPAQUE(x);
        const float sx = (x & 1) ? -1.f : 1.f;        // column parity of the input checkerboard
        cfloat tw[R];
        tw[0] = make_c(1.f, 0.f);
#pragma unroll
        for (int k1 = 1; k1 < R; ++k1) tw[k1] = smem[D::TW_OFF + (x * k1) % P::W];
        // parity 0 -> shared memory
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          finish(ur[0][r], ui[0][r], smem[D::TH_OFF + (0 * G + gg) * 4 + r], sx, tw);
          if (valid) {
            cfloat* dst = smem + (r * G + gg) * P::PITCH + x;
#pragma unroll
            for (int k1 = 0; k1 < R; ++k1) dst[k1 * P::SEG] = make_c(ur[0][r][k1], ui[0][r][k1]);
          }
        }
        // parity 1 -> parking lot (unconditional: the TMEM store is warp-collective)
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          finish(ur[1][r], ui[1][r], smem[D::TH_OFF + (1 * G + gg) * 4 + r], sx, tw);
        }
        float pk[PARK];
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
          for (int k1 = 0; k1 < R; ++k1) { pk[(r * R + k1) * 2] = ur[1][r][k1]; pk[(r * R + k1) * 2 + 1] = ui[1][r][k1]; }
        park.template store<PARK>(k * PARK, pk);
      }
    }
  }

  // parked parity-1 values -> B (same places Phase A writes parity 0 to); two tasks in flight
  template <class Park>
  static B2S_HD void unpark(cfloat* smem, int tid, const Park& park) {
    park.wait_st();
#pragma unroll 1
    for (int k = 0; k < TPT; ++k) {
      float pk[PARK];
      park.template load<PARK>(k * PARK, pk);
      park.wait_ld();
      int g, x0;
      if (task_of(tid, k, g, x0)) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          cfloat* dst = smem + (r * G + g) * P::PITCH + x0;
#pragma unroll
          for (int k1 = 0; k1 < R; ++k1) dst[k1 * P::SEG] = make_c(pk[(r * R + k1) * 2], pk[(r * R + k1) * 2 + 1]);
        }
      }
    }
  }
};

#if defined(__CUDACC__)
// Persistent whole-image kernel: CTA b processes images b, b + gridDim, ...
template <class P> struct WholeSmem { static constexpr int BYTES = Derived<P>::SMEM_BYTES + 16; };   // + the TMEM address slot

template <class P, class Pro, class Epi, int QD, int TT, bool CARRY, bool REVERSE = false>
__global__ void __launch_bounds__(P::NT, 1)
fft2_whole_kernel(const Pro pro, const Epi epi, const float scale, const int n_images, const int n_total, const int, const int) {
  using D = Derived<P>;
  extern __shared__ __align__(16) unsigned char b2s_smem_raw[];
  cfloat* smem = reinterpret_cast<cfloat*>(b2s_smem_raw);
  uint8_t* mrow = reinterpret_cast<uint8_t*>(smem + D::SMEM_ELEMS);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mrow + 2 * D::AUX_BYTES);
  const int tid = threadIdx.x;

  // 512 columns of tensor memory as parking lot (one CTA per SM: nothing else can want them)
  if (tid < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(tmem_slot)), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  build_tables<P>(smem, tid, P::NT);
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
  typedef PhaseAW<P, Pro, QD, TT> PA;
  typename PA::Queue queue;
  auto image_of = [&](int n) -> long long { return REVERSE ? n_total - 1 - n : n; };
  if (CARRY && (int)blockIdx.x < n_images) PA::prefill(pro, pro.ctx(image_of(blockIdx.x)), tid, queue);
  long long prev_image = -1;

#pragma unroll 1
  for (int n = blockIdx.x; n < n_images; n += gridDim.x) {
    const long long image = image_of(n);
    const int next = n + (int)gridDim.x;
    const bool has_next = next < n_images;
    const long long next_image = has_next ? image_of(next) : image;
    if (!CARRY) PA::prefill(pro, pro.ctx(image), tid, queue);
    if constexpr (Epi::FIXUP) epi.stage_mask_row(image, mrow + D::AUX_BYTES, tid, P::NT);
    PA::template run<true>(pro, pro.ctx(image), pro.ctx(next_image), CARRY && has_next, smem, tid, queue, park);
    __syncthreads();
    B2S_TICK(0);
    epi.stage_mask(image, mrow, tid, P::NT);
    if constexpr (Epi::FIXUP) {
      // all Phase C stores of the previous image have been issued (barrier inside Phase A): blend its sampled rows
      if (prev_image >= 0) epi.template fixup<P::NT>(prev_image, mrow + D::AUX_BYTES, tid);
      prev_image = image;
      B2S_TICK(4);
    }
    if (has_next) pro.l2_prefetch(next_image, tid);
    epi.l2_prefetch(image, 0, 1, tid);

#pragma unroll 1
    for (int q = 0; q < 2; ++q) {
      if (q == 1) {
        __syncthreads();                                     // every Phase C (q = 0) read of B is done
        PA::unpark(smem, tid, park);
        __syncthreads();
        B2S_TICK(5);
      }
#pragma unroll 1
      for (int round = 0; round < D::ROUNDS_B; ++round) {
        PhaseBRegs<P> s;
        phase_b_read<P>(smem, round, tid, s);
        __syncthreads();
        B2S_TICK(1);
        if constexpr (Epi::FIXUP) {                          // the list of THIS image's sampled rows (one buffer: the
          if (q == 0 && round == 0) epi.template stage_rows<1>(0, mrow + D::AUX_BYTES, tid, P::NT - 32);   // fix-up above is done)
        }
        phase_b_write<P>(smem, s);
        __syncthreads();
        B2S_TICK(2);
      }
      {
        const typename Epi::Ctx ectx = epi.ctx(image, mrow);
        for (int task = tid; task < D::TASKS_C; task += P::NT) phase_c<P>(epi, ectx, smem, q, task, scale);
      }
      B2S_TICK(3);
    }
  }
  if constexpr (Epi::FIXUP) {
    __syncthreads();
    if (prev_image >= 0) epi.template fixup<P::NT>(prev_image, mrow + D::AUX_BYTES, tid);
  }
  __syncthreads();
  if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
}
#endif

// Sequential emulation (tests/host_emul): one "CTA" walks every image; the parking lot is a per-thread array.
template <class P, class Pro, class Epi, int QD, int TT>
void fft2_whole_emulate(const Pro& pro, const Epi& epi, float scale, long long n_images) {
  using D = Derived<P>;
  typedef PhaseAW<P, Pro, QD, TT> PA;
  cfloat* smem = new cfloat[D::SMEM_ELEMS];
  uint8_t* mrow = new uint8_t[2 * D::AUX_BYTES];
  PhaseBRegs<P>* regs = new PhaseBRegs<P>[P::NT];
  typename PA::Queue* queues = new typename PA::Queue[P::NT];
  float* lot = new float[(size_t)P::NT * 256];
  for (int i = 0; i < D::SMEM_ELEMS; ++i) smem[i] = make_c(0.f, 0.f);
  for (int tid = 0; tid < P::NT; ++tid) build_tables<P>(smem, tid, P::NT);
  if (n_images > 0) for (int tid = 0; tid < P::NT; ++tid) PA::prefill(pro, pro.ctx(0), tid, queues[tid]);
  for (long long image = 0; image < n_images; ++image) {
    const bool has_next = image + 1 < n_images;
    for (int tid = 0; tid < P::NT; ++tid) epi.stage_mask(image, mrow, tid, P::NT);
    const typename Epi::Ctx ectx = epi.ctx(image, mrow);
    for (int tid = 0; tid < P::NT; ++tid) {
      ParkHost park{lot + (size_t)tid * 256};
      PA::template run<false>(pro, pro.ctx(image), pro.ctx(has_next ? image + 1 : image), has_next, smem, tid, queues[tid], park);
    }
    for (int q = 0; q < 2; ++q) {
      if (q == 1) for (int tid = 0; tid < P::NT; ++tid) { ParkHost park{lot + (size_t)tid * 256}; PA::unpark(smem, tid, park); }
      for (int round = 0; round < D::ROUNDS_B; ++round) {
        for (int tid = 0; tid < P::NT; ++tid) phase_b_read<P>(smem, round, tid, regs[tid]);
        for (int tid = 0; tid < P::NT; ++tid) phase_b_write<P>(smem, regs[tid]);
      }
      for (int tid = 0; tid < P::NT; ++tid)
        for (int task = tid; task < D::TASKS_C; task += P::NT) phase_c<P>(epi, ectx, smem, q, task, scale);
    }
    if constexpr (Epi::FIXUP) {
      for (int tid = 0; tid < P::NT; ++tid) epi.stage_mask_row(image, mrow + D::AUX_BYTES, tid, P::NT);
      for (int tid = 0; tid < P::NT; ++tid) epi.template stage_rows<1>(0, mrow + D::AUX_BYTES, tid, P::NT - 32);
      for (int tid = 0; tid < P::NT; ++tid) epi.template fixup<P::NT>(image, mrow + D::AUX_BYTES, tid);
    }
  }
  delete[] lot;
  delete[] queues;
  delete[] regs;
  delete[] mrow;
  delete[] smem;
}

}  // namespace b2s
