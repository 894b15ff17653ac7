// Persistent kernel wrapper (device) and sequential emulation (host) of the
// half-split fused 2-D FFT.  Work item = (coil image, half q); CTA b processes
// items b, b + gridDim, ... so the two halves of an image run on neighbouring
// CTAs at the same time and share the image through L2.
#pragma once
#include "fft2_core.cuh"
#if defined(__CUDACC__)
#include <cooperative_groups.h>
#endif

namespace b2s {

#if defined(__CUDACC__)
#ifdef B2S_PHASE_TIMING
__device__ unsigned long long g_phase_cycles[8];
#define B2S_TICK(slot) do { if (tid == 0) { const long long t_ = clock64(); atomicAdd(&g_phase_cycles[slot], (unsigned long long)(t_ - tprev)); tprev = t_; } } while (0)
#else
#define B2S_TICK(slot) do { } while (0)
#endif

// CARRY: keep the Phase A load queue alive across Phases B/C and the item boundary (hides the head
// latency of every Phase A; costs QD*16 registers during Phase C, so only epilogues without operand
// prefetch use it).
template <class P, class Pro, class Epi, bool CARRY, bool REVERSE = false>
__global__ void __launch_bounds__(P::NT, P::CTAS)
fft2_half_kernel(const Pro pro, const Epi epi, const float scale, const int n_items, const int item0) {
  using D = Derived<P>;
  extern __shared__ __align__(16) unsigned char b2s_smem_raw[];
  cfloat* smem = reinterpret_cast<cfloat*>(b2s_smem_raw);
  uint8_t* mrow = reinterpret_cast<uint8_t*>(smem + D::SMEM_ELEMS);
  const int tid = threadIdx.x;

  build_tables<P>(smem, tid, P::NT);
  __syncthreads();

#ifdef B2S_PHASE_TIMING
  long long tprev = clock64();
#endif
  typedef PhaseA<P, Pro> PA;
  typename PA::Queue queue;
  // REVERSE: walk the images from the last to the first (a template parameter: as a run-time argument it cost the
  // forward kernels 2 % through register allocation).  A kernel that consumes what the previous kernel just
  // streamed out (sens_reduce after sens_expand + DC: 192 MB through a 126 MB L2) then starts with the part that is
  // still cached instead of the part that was evicted first.
  const int n_img = n_items / P::FOLD;
  auto image_of = [&](int item) -> long long { const int n = item / P::FOLD; return REVERSE ? n_img - 1 - n : n; };
  if (CARRY && item0 + (int)blockIdx.x < n_items) PA::prefill(pro, pro.ctx(image_of(item0 + blockIdx.x)), tid, queue);
  static_assert(!Epi::FIXUP || 2 * D::MASK_BYTES + 16 <= D::AUX_BYTES, "aux buffer too small");
  int cur = 0;                                             // aux buffer of the current item (row fix-up epilogues)
  long long prev_image = -1;

#pragma unroll 1
  for (int item = item0 + blockIdx.x; item < n_items; item += gridDim.x) {
    const long long image = image_of(item);
    const int q = item % P::FOLD;
    const int next = item + (int)gridDim.x;
    const bool has_next = next < n_items;
    const long long next_image = has_next ? image_of(next) : image;
    if (!CARRY) PA::prefill(pro, pro.ctx(image), tid, queue);
    if constexpr (Epi::FIXUP) epi.stage_mask_row(image, mrow + cur * D::AUX_BYTES, tid, P::NT);
    // run<true>: the barrier that protects B (and the mask rows) from the previous item's Phase C sits
    // inside, just before the first write into B, so the first loads of this item are already in flight
    PA::template run<true>(pro, pro.ctx(image), pro.ctx(next_image), CARRY && has_next, smem, q, tid, queue);
    __syncthreads();
    B2S_TICK(0);
    epi.stage_mask(image, mrow, tid, P::NT);               // visible to Phase C through the barriers below
    if constexpr (Epi::FIXUP) {
      // every warp has passed the barrier inside Phase A, i.e. has issued all Phase C stores of the previous item:
      // blend its sampled rows now (its row list sits in the other aux buffer), and list this item's rows
      epi.template stage_rows<P::FOLD>(q, mrow + cur * D::AUX_BYTES, tid, P::NT - 32);   // last warp: not the one with the ragged Phase C round
      if (prev_image >= 0) epi.template fixup<P::NT>(prev_image, mrow + (cur ^ 1) * D::AUX_BYTES, tid);
      prev_image = image;
      cur ^= 1;
      B2S_TICK(4);
    }
    // warm L2 with the rest of the next item while this SM is busy with register codelets (Phases B, C)
    if (has_next && (next % P::FOLD) == 0) pro.l2_prefetch(next_image, tid);
    epi.l2_prefetch(image, q, P::FOLD, tid);                        // what Phase C will read (issued here, not before
                                                           // Phase A: bulk prefetches compete with its demand loads)

#pragma unroll 1
    for (int round = 0; round < D::ROUNDS_B; ++round) {
      PhaseBRegs<P> s;
      phase_b_read<P>(smem, round, tid, s);
      __syncthreads();
      B2S_TICK(1);
      phase_b_write<P>(smem, s);
      __syncthreads();
      B2S_TICK(2);
    }

    {
      // (requesting a task's epilogue operands one task ahead, or behind the last Phase B write, was
      // measured slower: sens_reduce 158 vs 140 us - the extra live registers cost more than the latency)
      const typename Epi::Ctx ectx = epi.ctx(image, mrow);
      for (int task = tid; task < D::TASKS_C; task += P::NT) phase_c<P>(epi, ectx, smem, q, task, scale);
    }
    B2S_TICK(3);                                           // no barrier here: see run<true>
  }
  if constexpr (Epi::FIXUP) {
    __syncthreads();                                       // last item: its Phase C stores and row list
    if (prev_image >= 0) epi.template fixup<P::NT>(prev_image, mrow + (cur ^ 1) * D::AUX_BYTES, tid);
  }
}

// Paired variant: a cluster of two CTAs per image (rank = output parity), see PhaseA2 in fft2_core.cuh.
template <class P, class Pro, class Epi, bool CARRY>
__global__ void __launch_bounds__(P::NT, 1)
fft2_pair_kernel(const Pro pro, const Epi epi, const float scale, const int n_images) {
  using D = Derived<P>;
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  extern __shared__ __align__(16) unsigned char b2s_smem_raw[];
  cfloat* smem = reinterpret_cast<cfloat*>(b2s_smem_raw);
  uint8_t* mrow = reinterpret_cast<uint8_t*>(smem + D::SMEM_ELEMS);
  const int tid = threadIdx.x;
  const int rank = (int)cluster.block_rank();                  // == output parity q
  const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  cfloat* remote = cluster.map_shared_rank(smem, rank ^ 1);    // the partner's shared memory (DSMEM)

  build_tables<P>(smem, tid, P::NT);
  typedef PhaseA2<P, Pro> PA;
  typename PA::Queue queue;
  if (CARRY && pair < n_images) PA::prefill(pro, pro.ctx(pair), tid, rank, queue);
  cluster.sync();                                              // tables visible; partner resident
#ifdef B2S_PHASE_TIMING
  long long tprev = clock64();
#endif
#pragma unroll 1
  for (int image = pair; image < n_images; image += n_pairs) {
    const int next = image + n_pairs;
    const bool has_next = next < n_images;
    epi.stage_mask(image, mrow, tid, P::NT);
    if (!CARRY) PA::prefill(pro, pro.ctx(image), tid, rank, queue);
    PA::run(pro, pro.ctx(image), pro.ctx(has_next ? next : image), CARRY && has_next, smem, remote, rank, tid, queue);
    cluster.sync();                                            // both halves of both B buffers written
    B2S_TICK(0);
    if (has_next && rank == 0) pro.l2_prefetch(next, tid);
    epi.l2_prefetch(image, rank, 2, tid);

#pragma unroll 1
    for (int round = 0; round < D::ROUNDS_B; ++round) {
      PhaseBRegs<P> s;
      phase_b_read<P>(smem, round, tid, s);
      __syncthreads();
      B2S_TICK(1);
      phase_b_write<P>(smem, s);
      __syncthreads();
      B2S_TICK(2);
    }
    {
      const typename Epi::Ctx ectx = epi.ctx(image, mrow);
      for (int task = tid; task < D::TASKS_C; task += P::NT) phase_c<P>(epi, ectx, smem, rank, task, scale);
    }
    cluster.sync();                                            // partner done reading its B before we write into it
    B2S_TICK(3);
  }
}
#endif

// Sequential execution of the same phases (tests/host_emul): a phase boundary
// is a barrier, so running each phase for all thread ids in turn is equivalent.
template <class P, class Pro, class Epi>
void fft2_half_emulate(const Pro& pro, const Epi& epi, float scale, long long n_images) {
  using D = Derived<P>;
  typedef PhaseA<P, Pro> PA;
  cfloat* smem = new cfloat[D::SMEM_ELEMS];
  uint8_t* mrow = new uint8_t[2 * D::AUX_BYTES];
  PhaseBRegs<P>* regs = new PhaseBRegs<P>[P::NT];
  typename PA::Queue* queues = new typename PA::Queue[P::NT];
  for (int i = 0; i < D::SMEM_ELEMS; ++i) smem[i] = make_c(0.f, 0.f);
  for (int tid = 0; tid < P::NT; ++tid) build_tables<P>(smem, tid, P::NT);
  const long long n_items = P::FOLD * n_images;
  if (n_items > 0) for (int tid = 0; tid < P::NT; ++tid) PA::prefill(pro, pro.ctx(0), tid, queues[tid]);
  for (long long item = 0; item < n_items; ++item) {          // one "CTA" walks every item (gridDim = 1)
    const long long image = item / P::FOLD;
    const int q = (int)(item % P::FOLD);
    const bool has_next = item + 1 < n_items;
    for (int tid = 0; tid < P::NT; ++tid) epi.stage_mask(image, mrow, tid, P::NT);
    const typename Epi::Ctx ectx = epi.ctx(image, mrow);
    for (int tid = 0; tid < P::NT; ++tid)
      PA::run(pro, pro.ctx(image), pro.ctx(has_next ? ((item + 1) / P::FOLD) : image), has_next, smem, q, tid, queues[tid]);
    for (int round = 0; round < D::ROUNDS_B; ++round) {
      for (int tid = 0; tid < P::NT; ++tid) phase_b_read<P>(smem, round, tid, regs[tid]);
      for (int tid = 0; tid < P::NT; ++tid) phase_b_write<P>(smem, regs[tid]);
    }
    for (int tid = 0; tid < P::NT; ++tid)
      for (int task = tid; task < D::TASKS_C; task += P::NT) phase_c<P>(epi, ectx, smem, q, task, scale);
    if constexpr (Epi::FIXUP) {                                // (the device defers this behind the next item's Phase A)
      for (int tid = 0; tid < P::NT; ++tid) epi.stage_mask_row(image, mrow, tid, P::NT);
      for (int tid = 0; tid < P::NT; ++tid) epi.template stage_rows<P::FOLD>(q, mrow, tid, P::NT - 32);
      for (int tid = 0; tid < P::NT; ++tid) epi.template fixup<P::NT>(image, mrow, tid);
    }
  }
  delete[] queues;
  delete[] regs;
  delete[] mrow;
  delete[] smem;
}


// Pair emulation: the two CTAs of a cluster run phase by phase, `remote` is simply the other array.
template <class P, class Pro, class Epi>
void fft2_pair_emulate(const Pro& pro, const Epi& epi, float scale, long long n_images) {
  using D = Derived<P>;
  typedef PhaseA2<P, Pro> PA;
  cfloat* smem[2]; uint8_t* mrow[2];
  PhaseBRegs<P>* regs = new PhaseBRegs<P>[P::NT];
  typename PA::Queue* queues[2];
  for (int r = 0; r < 2; ++r) {
    smem[r] = new cfloat[D::SMEM_ELEMS]; mrow[r] = new uint8_t[D::MASK_BYTES]; queues[r] = new typename PA::Queue[P::NT];
    for (int i = 0; i < D::SMEM_ELEMS; ++i) smem[r][i] = make_c(0.f, 0.f);
    for (int tid = 0; tid < P::NT; ++tid) build_tables<P>(smem[r], tid, P::NT);
    if (n_images > 0) for (int tid = 0; tid < P::NT; ++tid) PA::prefill(pro, pro.ctx(0), tid, r, queues[r][tid]);
  }
  for (long long image = 0; image < n_images; ++image) {
    const bool has_next = image + 1 < n_images;
    for (int r = 0; r < 2; ++r) {
      for (int tid = 0; tid < P::NT; ++tid) epi.stage_mask(image, mrow[r], tid, P::NT);
      for (int tid = 0; tid < P::NT; ++tid)
        PA::run(pro, pro.ctx(image), pro.ctx(has_next ? image + 1 : image), has_next, smem[r], smem[r ^ 1], r, tid, queues[r][tid]);
    }
    for (int r = 0; r < 2; ++r) {
      const typename Epi::Ctx ectx = epi.ctx(image, mrow[r]);
      for (int round = 0; round < D::ROUNDS_B; ++round) {
        for (int tid = 0; tid < P::NT; ++tid) phase_b_read<P>(smem[r], round, tid, regs[tid]);
        for (int tid = 0; tid < P::NT; ++tid) phase_b_write<P>(smem[r], regs[tid]);
      }
      for (int tid = 0; tid < P::NT; ++tid)
        for (int task = tid; task < D::TASKS_C; task += P::NT) phase_c<P>(epi, ectx, smem[r], r, task, scale);
    }
  }
  for (int r = 0; r < 2; ++r) { delete[] smem[r]; delete[] mrow[r]; delete[] queues[r]; }
  delete[] regs;
}

}  // namespace b2s
