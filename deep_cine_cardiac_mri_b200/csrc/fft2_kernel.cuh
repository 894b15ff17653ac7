// Persistent kernel wrapper (device) and sequential emulation (host) of the
// half-split fused 2-D FFT.  Work item = (coil image, half q); CTA b processes
// items b, b + gridDim, ... so the two halves of an image run on neighbouring
// CTAs at the same time and share the image through L2.
#pragma once
#include "fft2_core.cuh"

namespace b2s {

#if defined(__CUDACC__)
#ifdef B2S_PHASE_TIMING
__device__ unsigned long long g_phase_cycles[8];
#define B2S_TICK(slot) do { if (tid == 0) { const long long t_ = clock64(); atomicAdd(&g_phase_cycles[slot], (unsigned long long)(t_ - tprev)); tprev = t_; } } while (0)
#else
#define B2S_TICK(slot) do { } while (0)
#endif

template <class P, class Pro, class Epi>
__global__ void __launch_bounds__(P::NT, 1)
fft2_half_kernel(const Pro pro, const Epi epi, const float scale, const int n_items, const unsigned stagger_ns) {
  using D = Derived<P>;
  extern __shared__ __align__(16) unsigned char b2s_smem_raw[];
  cfloat* smem = reinterpret_cast<cfloat*>(b2s_smem_raw);
  const int tid = threadIdx.x;

  build_tables<P>(smem, tid, P::NT);
  if (stagger_ns && tid == 0) {
    // de-phase the persistent CTAs: identical work items would otherwise keep every SM in the
    // same (load | compute | store) phase at the same time and serialise HBM against the math
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    const unsigned long long wait = (unsigned long long)(blockIdx.x & 3u) * stagger_ns;
    do { __nanosleep(256); asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1)); } while (t1 - t0 < wait);
  }
  __syncthreads();

#ifdef B2S_PHASE_TIMING
  long long tprev = clock64();
#endif
#pragma unroll 1
  for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
    const long long image = item >> 1;
    const int q = item & 1;
    epi.l2_prefetch(image, q, tid);                        // what Phase C will read
    {
      const typename Pro::Ctx pctx = pro.ctx(image);
      for (int task = tid; task < D::TASKS_A; task += P::NT) phase_a<P>(pro, pctx, smem, q, task);
    }
    __syncthreads();
    B2S_TICK(0);

    const int next = item + (int)gridDim.x;                // warm L2 with the next item's input
    if (next < n_items && !(next & 1)) pro.l2_prefetch(next >> 1, tid);

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
      const typename Epi::Ctx ectx = epi.ctx(image);
      for (int task = tid; task < D::TASKS_C; task += P::NT) phase_c<P>(epi, ectx, smem, q, task, scale);
    }
    __syncthreads();                                       // B is rewritten by the next item's Phase A
    B2S_TICK(3);
  }
}
#endif

// Sequential execution of the same phases (tests/host_emul): a phase boundary
// is a barrier, so running each phase for all thread ids in turn is equivalent.
template <class P, class Pro, class Epi>
void fft2_half_emulate(const Pro& pro, const Epi& epi, float scale, long long n_images) {
  using D = Derived<P>;
  cfloat* smem = new cfloat[D::SMEM_ELEMS];
  PhaseBRegs<P>* regs = new PhaseBRegs<P>[P::NT];
  for (int i = 0; i < D::SMEM_ELEMS; ++i) smem[i] = make_c(0.f, 0.f);
  for (int tid = 0; tid < P::NT; ++tid) build_tables<P>(smem, tid, P::NT);
  for (long long item = 0; item < 2 * n_images; ++item) {
    const long long image = item >> 1;
    const int q = (int)(item & 1);
    const typename Pro::Ctx pctx = pro.ctx(image);
    const typename Epi::Ctx ectx = epi.ctx(image);
    for (int tid = 0; tid < P::NT; ++tid)
      for (int task = tid; task < D::TASKS_A; task += P::NT) phase_a<P>(pro, pctx, smem, q, task);
    for (int round = 0; round < D::ROUNDS_B; ++round) {
      for (int tid = 0; tid < P::NT; ++tid) phase_b_read<P>(smem, round, tid, regs[tid]);
      for (int tid = 0; tid < P::NT; ++tid) phase_b_write<P>(smem, regs[tid]);
    }
    for (int tid = 0; tid < P::NT; ++tid)
      for (int task = tid; task < D::TASKS_C; task += P::NT) phase_c<P>(epi, ectx, smem, q, task, scale);
  }
  delete[] regs;
  delete[] smem;
}

}  // namespace b2s
