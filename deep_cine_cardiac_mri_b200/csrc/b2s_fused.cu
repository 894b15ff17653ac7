// Fused single-pass SENSE operators on the plan sizes (200 x 200) and their
// composition from the generic kernels for every other size.
#include <stdlib.h>
#include <type_traits>
#include "b2s_common.cuh"
#include "sense_functors.cuh"
#include "fft2_kernel.cuh"
#include "fft2_whole.cuh"     // (parking-lot types; the scalar whole-image kernel itself is experimental)
#include "fft2_packed.cuh"

using namespace b2s;

// Which kernel family serves the fused plan sizes (b2s_set_fused_path, include/b200sense.h):
//   0 auto (default)  1 strip-streamed (experimental builds)  2 half/quarter-split only  3 packed whole-image wherever it exists

namespace {

typedef Plan<200, 200, 256, 1, 2> P200H;   // half split: 8 warps x 255 registers, 1 CTA/SM
typedef Plan<256, 256, 256, 1, 4> P256;    // quarter split, 1 CTA/SM
typedef Plan<200, 200, 256, 2, 2, 1> P200W; // half split, 128-bit Phase A loads (two columns per thread)
typedef Plan<256, 256, 256, 2, 4, 1> P256W; // quarter split, 128-bit Phase A loads
typedef PackPlan200<256> PK200;             // packed whole-image kernel (fft2_packed.cuh)
#ifdef B2S_EXPERIMENTS
typedef Plan<200, 200, 128, 1, 4> P200Q;   // quarter split: 2 CTAs x 4 warps per SM
static int env_int(const char* name, int dflt) { const char* e = getenv(name); return e ? atoi(e) : dflt; }
#endif

struct DeviceInfo { int sms; };
static int device_info(DeviceInfo& d) {
  int dev = 0;
  B2S_CUDA(cudaGetDevice(&dev));
  static std::atomic<int> sm_count[64];       // immutable per-device facts (zero-initialised statics)
  if (dev < 0 || dev >= 64) return fail(B2S_EUNSUPPORTED, "device index >= 64");
  if (!sm_count[dev].load()) { int n = 0; B2S_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev)); sm_count[dev].store(n); }
  d.sms = sm_count[dev].load();
  return B2S_OK;
}
static int grid_slots(const DeviceInfo& d) { const int r = g_sm_reserve.load(); return d.sms - r > 0 ? d.sms - r : 1; }   // see b2s_set_sm_reserve

// Opt-in to > 48 KB of dynamic shared memory, once per kernel and device.  `TAG` makes the flag array unique per KERNEL
// (kernels of different plans share a function-pointer type, so the pointer type alone is not a key); setting the
// attribute twice is harmless.
template <class TAG, class K> static int allow_smem(K kern, int bytes) {
  int dev = 0;
  B2S_CUDA(cudaGetDevice(&dev));
  static std::atomic<bool> configured[64];
  if (!configured[dev & 63].load()) {
    B2S_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    configured[dev & 63].store(true);
  }
  return B2S_OK;
}
template <class... T> struct KernelTag {};

// REVERSE: the kernel walks the images last-to-first (sens_reduce: it usually follows the kernel that wrote them, and
// the end of a 192 MB stream is what is still in the 126 MB L2).  `first_image`: images before it are skipped (they
// were handled by the whole-image kernel; with REVERSE the LAST first_image images are the ones skipped).
template <class P, class Pro, class Epi, bool CARRY = false, bool REVERSE = false>
int launch_fused(const Pro& pro, const Epi& epi, float scale, int64_t n_images, cudaStream_t st, int64_t first_image = 0) {
  if (n_images <= first_image) return B2S_OK;
  if (P::FOLD * n_images > 0x7fffffffLL) return fail(B2S_EUNSUPPORTED, "too many images for one launch");
  auto kern = fft2_half_kernel<P, Pro, Epi, CARRY, REVERSE>;
  DeviceInfo d;
  if (int rc = device_info(d)) return rc;
  if (int rc = allow_smem<KernelTag<P, Pro, Epi, std::integral_constant<int, CARRY + 2 * REVERSE>>>(kern, Derived<P>::SMEM_BYTES)) return rc;
  const int n_items = (int)(P::FOLD * n_images), item0 = (int)(P::FOLD * first_image);
  const int slots = grid_slots(d) * P::CTAS;
  const unsigned grid = (unsigned)(n_items - item0 < slots ? n_items - item0 : slots);   // persistent: P::CTAS CTAs per SM
  kern<<<grid, P::NT, Derived<P>::SMEM_BYTES, st>>>(pro, epi, scale, n_items, item0);
  return check_launch("fft2_half_kernel");
}

// Packed whole-image kernel (fft2_packed.cuh: one image per CTA, second half parked in tensor memory, two transforms
// per thread in packed fp32) on as many images as fill whole rounds of the persistent grid; a remainder that fits one
// round of half items (2 * rem <= CTAs) goes through the half-split kernel behind it with the epilogue `tail_epi` (a
// ragged last round of whole images would idle most SMs for a full image time).
template <class P, class Pro, class Epi, int QD, int TT, bool CARRY, bool REVERSE, class TailEpi>
int launch_packed(const Pro& pro, const Epi& epi, const TailEpi& tail_epi, float scale, int64_t n_images, cudaStream_t st) {
  if (n_images <= 0) return B2S_OK;
  if (2 * n_images > 0x7fffffffLL) return fail(B2S_EUNSUPPORTED, "too many images for one launch");
  auto kern = fft2_packed_kernel<PK200, Pro, Epi, QD, TT, CARRY, REVERSE>;
  constexpr int SMEM = PkSmem<PK200, Epi>::BYTES;
  DeviceInfo d;
  if (int rc = device_info(d)) return rc;
  if (int rc = allow_smem<KernelTag<PK200, Pro, Epi, std::integral_constant<int, QD * 100 + TT * 10 + CARRY + 2 * REVERSE>>>(kern, SMEM)) return rc;
  const int slots = grid_slots(d);
  int64_t n_whole = n_images;
  const int64_t rem = n_images % slots;
  if (n_images > slots && rem > 0 && 2 * rem <= slots) n_whole = n_images - rem;
  const unsigned grid = (unsigned)(n_whole < slots ? n_whole : slots);
  kern<<<grid, PK200::NT, SMEM, st>>>(pro, epi, scale, (int)n_whole, (int)n_images, 1, 0);
  const int rc = check_launch("fft2_packed_kernel");
  if (rc || n_whole == n_images) return rc;
  return launch_fused<P, Pro, TailEpi, CARRY, REVERSE>(pro, tail_epi, scale, n_images, st, n_whole);
}

// Measured cost model (B200, microseconds per round of the 148-CTA persistent grid, profiles/r2_analysis.md): the
// packed whole-image kernel processes an image in less SM time, but its work items are twice as coarse, so a launch
// whose image count is not a multiple of the grid pays a tail (remainder through the half-split kernel behind it).
struct KernelCost { float whole_round, half_round, tail; };
static bool packed_is_faster(int64_t n, int slots, const KernelCost& k) {
  const int path = g_fused_path.load();
  if (path == 2) return false;
  if (path == 3) return true;
  const int64_t rem = n % slots;
  const float whole = (float)(n / slots) * k.whole_round + (rem == 0 ? 0.f : (n > slots && 2 * rem <= slots ? k.tail : k.whole_round));
  const float half = (float)((2 * n + slots - 1) / slots) * k.half_round;
  return whole < half;
}

}  // namespace

// ---- per-plan launch helpers -------------------------------------------------------------------
template <class P>
int plan_fft2c(const float* in, float* out, int64_t n_images, int inverse, float scale, cudaStream_t st) {
  constexpr int H = P::H, W = P::W;
  const long long hw = (long long)H * W;
  const float s = scale * centre_sign<P>();
  // (the packed whole-image kernel ties with the half split for the plain transform and for sens_reduce at full rounds
  // and loses with a remainder - profiles/r2_analysis.md - so they stay on the half split)
  if (inverse) {
    ProPlain<H, W, true> pro{(const cfloat*)in, hw};
    EpiPlain<H, W, true> epi{(cfloat*)out, hw};
    return launch_fused<P, ProPlain<H, W, true>, EpiPlain<H, W, true>, true>(pro, epi, s, n_images, st);
  }
  ProPlain<H, W, false> pro{(const cfloat*)in, hw};
  EpiPlain<H, W, false> epi{(cfloat*)out, hw};
  return launch_fused<P, ProPlain<H, W, false>, EpiPlain<H, W, false>, true>(pro, epi, s, n_images, st);
}

template <class P>
int plan_expand(const float* image, const float* sens, float* kspace, const float* ref, const uint8_t* mask,
                const float* v, int mode, int t, int c, int64_t n, float scale, cudaStream_t st) {
  constexpr int H = P::H, W = P::W;
  const long long hw = (long long)H * W;
  const float s = scale * centre_sign<P>();
  ProExpand<H, W> pro{(const cfloat*)image, (const cfloat*)sens, t, c, hw};
#define B2S_RUN(M)                                                                    \
  {                                                                                   \
    EpiKspace<H, W, M> epi{(cfloat*)kspace, (const cfloat*)ref, mask, v, c, hw};      \
    return launch_fused<P>(pro, epi, s, n, st);                                       \
  }
  switch (mode) { case 0: B2S_RUN(0) case 1: B2S_RUN(1) case 2: B2S_RUN(2) default: B2S_RUN(3) }
#undef B2S_RUN
}

// sens_expand at 200 x 200 on the packed whole-image kernel (remainder images: half split, same epilogue semantics)
static int packed_expand(const float* image, const float* sens, float* kspace, const float* ref, const uint8_t* mask,
                         const float* v, int mode, int t, int c, int64_t n, float scale, cudaStream_t st) {
  constexpr int H = 200, W = 200;
  const long long hw = (long long)H * W;
  const float s = scale * centre_sign<P200H>();
  typedef ProExpand<H, W> PR;
  PR pro{(const cfloat*)image, (const cfloat*)sens, t, c, hw};
#define B2S_RUN(M)                                                                    \
  {                                                                                   \
    EpiKspace<H, W, M> epi{(cfloat*)kspace, (const cfloat*)ref, mask, v, c, hw};      \
    return launch_packed<P200H, PR, EpiKspace<H, W, M>, 2, 2, false, false>(pro, epi, epi, s, n, st); \
  }
  if (mode == 2) {        // soft DC: reference rows staged in shared memory (EpiDCStage); tail images: fused epilogue
    EpiDCStage<H, W> epi{(cfloat*)kspace, (const cfloat*)ref, mask, v, c, hw};
    EpiKspace<H, W, 2> tail{(cfloat*)kspace, (const cfloat*)ref, mask, v, c, hw};
    return launch_packed<P200H, PR, EpiDCStage<H, W>, 2, 2, false, false>(pro, epi, tail, s, n, st);
  }
  switch (mode) { case 0: B2S_RUN(0) case 1: B2S_RUN(1) default: B2S_RUN(3) }
#undef B2S_RUN
}

template <class P>
int plan_reduce(const float* kspace, const float* mult, float* out, const uint8_t* mask, const float* v,
                int weight_mode, int over_frames, int t, int c, int64_t n, float scale, cudaStream_t st) {
  constexpr int H = P::H, W = P::W;
  const long long hw = (long long)H * W;
  const float s = scale * centre_sign<P>();
  EpiReduce<H, W> epi;
  epi.out = (cfloat*)out; epi.mult = (const cfloat*)mult; epi.T = t; epi.C = c;
  if (!over_frames) { epi.os_b = t * hw; epi.os_t = hw; epi.os_c = 0; epi.ms_b = c * hw; epi.ms_t = 0; epi.ms_c = hw; }
  else              { epi.os_b = c * hw; epi.os_t = 0; epi.os_c = hw; epi.ms_b = t * hw; epi.ms_t = hw; epi.ms_c = 0; }
#define B2S_RUN(M)                                                      \
  {                                                                     \
    ProKspace<H, W, M> pro{(const cfloat*)kspace, mask, v, c, hw};      \
    return launch_fused<P, ProKspace<H, W, M>, EpiReduce<H, W>, false, true>(pro, epi, s, n, st); \
  }
  switch (weight_mode) { case 0: B2S_RUN(0) case 1: B2S_RUN(1) default: B2S_RUN(2) }
#undef B2S_RUN
}

// Deterministic variant of plan_reduce: weighted inverse transform of every coil image into `y` (no float
// atomics); the caller then sums the coils in a fixed order (launch_coil_reduce).
template <class P>
int plan_ifft_weighted(const float* kspace, float* y, const uint8_t* mask, const float* v, int weight_mode,
                       int c, int64_t n, float scale, cudaStream_t st) {
  constexpr int H = P::H, W = P::W;
  const long long hw = (long long)H * W;
  const float s = scale * centre_sign<P>();
  EpiPlain<H, W, true> epi{(cfloat*)y, hw};
#define B2S_RUN(M)                                                      \
  {                                                                     \
    ProKspace<H, W, M> pro{(const cfloat*)kspace, mask, v, c, hw};      \
    return launch_fused<P>(pro, epi, s, n, st);                         \
  }
  switch (weight_mode) { case 0: B2S_RUN(0) case 1: B2S_RUN(1) default: B2S_RUN(2) }
#undef B2S_RUN
}

static inline int plan_id(int h, int w) { return (h == 200 && w == 200) ? 1 : (h == 256 && w == 256) ? 2 : 0; }

static inline bool aligned16(const void* a, const void* b = nullptr, const void* c2 = nullptr) {
  return (((uintptr_t)a | (uintptr_t)b | (uintptr_t)c2) & 15) == 0;
}

#ifdef B2S_EXPERIMENTS
namespace {
#include "b2s_fused_experiments.inc"
}
#endif

extern "C" int b2s_has_fused_plan(int h, int w) { return plan_id(h, w) ? 1 : 0; }

extern "C" size_t b2s_scratch_bytes(int b, int t, int c, int h, int w) {
  if (b2s_has_fused_plan(h, w)) return 0;
  return (size_t)b * t * c * h * w * 2 * sizeof(float);
}

extern "C" int b2s_fft2c(const float* in, float* out, int64_t n_images, int h, int w, int inverse,
                         int norm, void* stream) {
  if (h <= 0 || w <= 0 || n_images < 0 || bad_norm(norm)) return fail(B2S_EINVAL, "b2s_fft2c: bad argument");
  if (n_images == 0) return B2S_OK;
  if (!in || !out) return fail(B2S_EINVAL, "b2s_fft2c: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const float scale = norm_scale(h, w, inverse, norm);
#ifdef B2S_EXPERIMENTS
  { int rc = 0; if (experimental_fft2c(in, out, n_images, h, w, inverse, scale, st, &rc)) return rc; }
#endif
  switch (plan_id(h, w)) {
    case 1: return plan_fft2c<P200H>(in, out, n_images, inverse, scale, st);
    case 2: return plan_fft2c<P256>(in, out, n_images, inverse, scale, st);
    default: return generic_fft2(in, out, n_images, h, w, inverse, scale, st);
  }
}

extern "C" int b2s_sens_expand(const float* image, const float* sens, float* kspace, const float* ref,
                               const uint8_t* mask, const float* v, int mode, int b, int t, int c,
                               int h, int w, int norm, void* scratch, size_t scratch_bytes, void* stream) {
  (void)scratch; (void)scratch_bytes;
  if (b < 0 || t < 0 || c < 0 || bad_norm(norm) || mode < 0 || mode > 3)
    return fail(B2S_EINVAL, "b2s_sens_expand: bad argument");
  if ((int64_t)b * t * c == 0) return B2S_OK;
  if (!image || !sens || !kspace) return fail(B2S_EINVAL, "b2s_sens_expand: null pointer");
  if ((mode >= 1 && !mask) || (mode >= 2 && !ref) || (mode == 2 && !v))
    return fail(B2S_EINVAL, "b2s_sens_expand: mode needs mask/ref/v");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t n = (int64_t)b * t * c;
  const float scale = norm_scale(h, w, 0, norm);
#ifdef B2S_EXPERIMENTS
  { int rc = 0; if (experimental_expand(image, sens, kspace, ref, mask, v, mode, t, c, n, h, w, scale, st, &rc)) return rc; }
#endif
  // 128-bit Phase A loads need 16-byte aligned rows (a tensor view at an odd complex offset is only 8-byte aligned)
  const bool wide_ok = aligned16(image, sens);
  switch (plan_id(h, w)) {
    case 1: {
      DeviceInfo d;
      if (int rc = device_info(d)) return rc;
      // per round of the persistent grid, microseconds: {packed whole image, half item, tail launch}
      const KernelCost cost = (mode == 2) ? KernelCost{39.5f, 20.8f, 28.f} : KernelCost{27.8f, 17.0f, 22.f};
      if ((mode != 2 || aligned16(ref)) && packed_is_faster(n, grid_slots(d), cost))
        return packed_expand(image, sens, kspace, ref, mask, v, mode, t, c, n, scale, st);
      return (wide_ok && mode != 2) ? plan_expand<P200W>(image, sens, kspace, ref, mask, v, mode, t, c, n, scale, st)
                                    : plan_expand<P200H>(image, sens, kspace, ref, mask, v, mode, t, c, n, scale, st);
    }
    case 2: return wide_ok ? plan_expand<P256W>(image, sens, kspace, ref, mask, v, mode, t, c, n, scale, st)
                           : plan_expand<P256>(image, sens, kspace, ref, mask, v, mode, t, c, n, scale, st);
    default: break;
  }
  // generic sizes: S*x -> kspace, FFT in place, epilogue in place
  int rc = launch_expand_product(image, sens, kspace, b, t, c, (int64_t)h * w, st);
  if (rc) return rc;
  rc = generic_fft2(kspace, kspace, n, h, w, 0, scale, st);
  if (rc) return rc;
  if (mode == 0) return B2S_OK;
  return launch_kspace_epilogue(kspace, ref, mask, v, mode, (int64_t)b * t, c, h, w, st);
}

extern "C" int b2s_sens_reduce(const float* kspace, const float* mult, float* out, const uint8_t* mask,
                               const float* v, int weight_mode, int over_frames, int b, int t, int c,
                               int h, int w, int norm, void* scratch, size_t scratch_bytes, void* stream) {
  const int deterministic = (weight_mode & B2S_REDUCE_DETERMINISTIC) ? 1 : 0;
  weight_mode &= ~B2S_REDUCE_DETERMINISTIC;
  if (b < 0 || t < 0 || c < 0 || bad_norm(norm) || weight_mode < 0 || weight_mode > 2)
    return fail(B2S_EINVAL, "b2s_sens_reduce: bad argument");
  if ((over_frames ? (int64_t)b * c : (int64_t)b * t) == 0) return B2S_OK;
  if (!out || ((int64_t)b * t * c > 0 && (!kspace || !mult))) return fail(B2S_EINVAL, "b2s_sens_reduce: null pointer");
  if ((weight_mode >= 1 && !mask) || (weight_mode == 2 && !v))
    return fail(B2S_EINVAL, "b2s_sens_reduce: weight mode needs mask/v");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t n = (int64_t)b * t * c;
  const int64_t hw = (int64_t)h * w;
  const float scale = norm_scale(h, w, 1, norm);
  const int64_t out_images = over_frames ? (int64_t)b * c : (int64_t)b * t;
  if (plan_id(h, w) && deterministic && n > 0) {
    const size_t need = (size_t)n * hw * 2 * sizeof(float);
    if (!scratch || scratch_bytes < need) return fail(B2S_EINVAL, "b2s_sens_reduce: deterministic mode needs b*t*c*h*w*8 scratch bytes");
    float* y = (float*)scratch;
    const int rc = plan_id(h, w) == 2 ? plan_ifft_weighted<P256>(kspace, y, mask, v, weight_mode, c, n, scale, st)
                                      : plan_ifft_weighted<P200H>(kspace, y, mask, v, weight_mode, c, n, scale, st);
    if (rc) return rc;
    return launch_coil_reduce(y, mult, out, over_frames, b, t, c, hw, st);
  }
  if (plan_id(h, w)) {
    B2S_CUDA(cudaMemsetAsync(out, 0, (size_t)out_images * hw * 2 * sizeof(float), st));
    if (n == 0) return B2S_OK;
#ifdef B2S_EXPERIMENTS
    { int rc = 0; if (experimental_reduce(kspace, mult, out, mask, v, weight_mode, over_frames, t, c, n, h, w, scale, st, &rc)) return rc; }
#endif
    if (plan_id(h, w) == 2) return plan_reduce<P256>(kspace, mult, out, mask, v, weight_mode, over_frames, t, c, n, scale, st);
    return plan_reduce<P200H>(kspace, mult, out, mask, v, weight_mode, over_frames, t, c, n, scale, st);
  }
  // generic sizes: (row weight) -> IFFT into scratch -> conj-multiply + reduce
  const size_t need = (size_t)n * hw * 2 * sizeof(float);
  if (n > 0 && (!scratch || scratch_bytes < need)) return fail(B2S_EINVAL, "b2s_sens_reduce: scratch too small (see b2s_scratch_bytes)");
  float* y = (float*)scratch;
  int rc = B2S_OK;
  const float* src = kspace;
  if (weight_mode) {
    rc = launch_row_weight(kspace, y, mask, v, weight_mode, (int64_t)b * t, c, h, w, st);
    if (rc) return rc;
    src = y;
  }
  if (n > 0) {
    rc = generic_fft2(src, y, n, h, w, 1, scale, st);
    if (rc) return rc;
  }
  return launch_coil_reduce(y, mult, out, over_frames, b, t, c, hw, st);
}

extern "C" size_t b2s_dc_step_ws_bytes(int b, int t, int c, int h, int w) {
  const size_t K = (size_t)b * t * c * h * w * 8, I = (size_t)b * t * h * w * 8, S = (size_t)b * c * h * w * 8;
  const size_t M = ((size_t)b * t * h + 255) / 256 * 256;
  return 3 * K + I + S + M + 256 + b2s_scratch_bytes(b, t, c, h, w);
}

extern "C" int b2s_dc_step_host(const float* kspace_host, const float* ref_host, const float* sens_host,
                                const uint8_t* mask_host, float v_value, float* out_host, int b, int t,
                                int c, int h, int w, void* ws, size_t ws_bytes, void* stream) {
  if (!kspace_host || !ref_host || !sens_host || !mask_host || !out_host || !ws)
    return fail(B2S_EINVAL, "b2s_dc_step_host: null pointer");
  if (ws_bytes < b2s_dc_step_ws_bytes(b, t, c, h, w)) return fail(B2S_EINVAL, "b2s_dc_step_host: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t K = (size_t)b * t * c * h * w * 8, I = (size_t)b * t * h * w * 8, S = (size_t)b * c * h * w * 8;
  const size_t M = ((size_t)b * t * h + 255) / 256 * 256;
  char* p = (char*)ws;
  float* dk = (float*)p; p += K;
  float* dref = (float*)p; p += K;
  float* dout = (float*)p; p += K;
  float* dimg = (float*)p; p += I;
  float* dsens = (float*)p; p += S;
  uint8_t* dmask = (uint8_t*)p; p += M;
  float* dv = (float*)p; p += 256;
  void* scratch = p;
  const size_t sb = b2s_scratch_bytes(b, t, c, h, w);
  B2S_CUDA(cudaMemcpyAsync(dk, kspace_host, K, cudaMemcpyHostToDevice, st));
  B2S_CUDA(cudaMemcpyAsync(dref, ref_host, K, cudaMemcpyHostToDevice, st));
  B2S_CUDA(cudaMemcpyAsync(dsens, sens_host, S, cudaMemcpyHostToDevice, st));
  B2S_CUDA(cudaMemcpyAsync(dmask, mask_host, (size_t)b * t * h, cudaMemcpyHostToDevice, st));
  B2S_CUDA(cudaMemcpyAsync(dv, &v_value, sizeof(float), cudaMemcpyHostToDevice, st));
  int rc = b2s_sens_reduce(dk, dsens, dimg, nullptr, nullptr, 0, 0, b, t, c, h, w, B2S_NORM_ORTHO, scratch, sb, st);
  if (rc) return rc;
  rc = b2s_sens_expand(dimg, dsens, dout, dref, dmask, dv, B2S_EXPAND_DC, b, t, c, h, w, B2S_NORM_ORTHO, scratch, sb, st);
  if (rc) return rc;
  B2S_CUDA(cudaMemcpyAsync(out_host, dout, K, cudaMemcpyDeviceToHost, st));
  return B2S_OK;
}

#ifdef B2S_PHASE_TIMING
extern "C" int b2s_debug_phase_cycles(unsigned long long* out_host, int reset) {
  B2S_CUDA(cudaDeviceSynchronize());
  B2S_CUDA(cudaMemcpyFromSymbol(out_host, b2s::g_phase_cycles, 8 * sizeof(unsigned long long)));
  if (reset) { unsigned long long z[8] = {0}; B2S_CUDA(cudaMemcpyToSymbol(b2s::g_phase_cycles, z, sizeof(z))); }
  return B2S_OK;
}
#endif
