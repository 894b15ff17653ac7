// Kernel wrapper (device) and sequential emulation (host) of the half-split
// fused 2-D FFT.  Grid = 2 CTAs per coil image (q = blockIdx & 1).
#pragma once
#include "fft2_core.cuh"

namespace b2s {

#if defined(__CUDACC__)
template <class P, class Pro, class Epi>
__global__ void __launch_bounds__(P::NT, 1)
fft2_half_kernel(const Pro pro, const Epi epi, const float scale) {
  using D = Derived<P>;
  extern __shared__ __align__(16) unsigned char b2s_smem_raw[];
  cfloat* smem = reinterpret_cast<cfloat*>(b2s_smem_raw);
  const int tid = threadIdx.x;
  const long long image = blockIdx.x >> 1;
  const int q = blockIdx.x & 1;

  build_tables<P>(smem, q, tid, P::NT);
  const typename Pro::Ctx pctx = pro.ctx(image);
  const typename Epi::Ctx ectx = epi.ctx(image);
  __syncthreads();

  for (int task = tid; task < D::TASKS_A; task += P::NT) phase_a<P>(pro, pctx, smem, q, task);
  __syncthreads();

#pragma unroll 1
  for (int round = 0; round < D::ROUNDS_B; ++round) {
    PhaseBRegs<P> s;
    phase_b_read<P>(smem, round, tid, s);
    __syncthreads();
    phase_b_write<P>(smem, s);
    __syncthreads();
  }

  for (int task = tid; task < D::TASKS_C; task += P::NT) phase_c<P>(epi, ectx, smem, q, task, scale);
}
#endif

// Sequential execution of the same phases (tests/host_emul): a phase boundary
// is a barrier, so running each phase for all thread ids in turn is equivalent.
template <class P, class Pro, class Epi>
void fft2_half_emulate(const Pro& pro, const Epi& epi, float scale, long long n_images) {
  using D = Derived<P>;
  cfloat* smem = new cfloat[D::SMEM_ELEMS];
  PhaseBRegs<P>* regs = new PhaseBRegs<P>[P::NT];
  for (long long item = 0; item < 2 * n_images; ++item) {
    const long long image = item >> 1;
    const int q = (int)(item & 1);
    for (int i = 0; i < D::SMEM_ELEMS; ++i) smem[i] = make_c(0.f, 0.f);
    for (int tid = 0; tid < P::NT; ++tid) build_tables<P>(smem, q, tid, P::NT);
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
