// Launchers of the strip-streamed fused 2-D FFT kernels (strip_core.cuh) and their per-(device, stream)
// workspace: ticket counter, per-image dependency counters and the L2-resident scratch ring.
#include <stdlib.h>
#include <map>
#include <mutex>
#include "b2s_common.cuh"
#include "strip_core.cuh"

using namespace b2s;

namespace {

struct StripWs {
  char* base = nullptr;
  size_t bytes = 0;
  int cap_images = 0;
  int nslot = 0;
  long long hw = 0;
  int* status() const { return (int*)base; }
  int* head() const { return (int*)(base + 256); }
  int* done_r() const { return head() + 64; }
  int* done_c() const { return done_r() + cap_images; }
  size_t counter_bytes() const { return (size_t)(64 + 2 * (size_t)cap_images) * sizeof(int); }
  cfloat* scratch() const { return (cfloat*)(base + 256 + ((counter_bytes() + 255) / 256) * 256); }
};

std::mutex g_ws_mutex;
std::map<std::pair<int, cudaStream_t>, StripWs> g_ws;

int env_int(const char* name, int dflt) { const char* e = getenv(name); return e ? atoi(e) : dflt; }

// LAG: how many images pass C trails pass R in the ticket order; NSLOT: images in the scratch ring.
// 148 SMs x 16 warps hold ~2400 tickets (24 images' worth at 100 tickets per image) in flight at most; pass C of an
// image is handed out `lag` images after its pass R, and the ring must hold lag + in-flight images.
int strip_lag() { static int v = env_int("B2S_STRIP_LAG", 40); return v < 0 ? 0 : v; }
int strip_nslot() { static int v = env_int("B2S_STRIP_NSLOT", 80); return v; }

int alloc_ws(int64_t n_images, long long hw, int nslot, StripWs& out) {
  int cap = 4096;
  while (cap < n_images) cap *= 2;
  StripWs n;
  n.cap_images = cap; n.nslot = nslot; n.hw = hw;
  const size_t cb = ((n.counter_bytes() + 255) / 256) * 256;
  n.bytes = 256 + cb + (size_t)nslot * (size_t)hw * sizeof(cfloat);
  B2S_CUDA(cudaMalloc((void**)&n.base, n.bytes));
  B2S_CUDA(cudaMemset(n.base, 0, 256 + cb));
  out = n;
  return B2S_OK;
}

std::map<int, StripWs> g_spare;   // one unassigned workspace per device, handed to a stream first seen while capturing

// returns B2S_OK and fills `ws`, or an error; `*unavailable` = 1 when the stream is capturing, has no
// workspace yet and no spare fits (cudaMalloc is illegal during capture): the caller falls back.
int get_ws(cudaStream_t st, int64_t n_images, long long hw, StripWs& ws, int* unavailable) {
  *unavailable = 0;
  int dev = 0;
  B2S_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lock(g_ws_mutex);
  StripWs& w = g_ws[std::make_pair(dev, st)];
  const int nslot = strip_nslot() > strip_lag() ? strip_nslot() : strip_lag() + 1;
  auto fits = [&](const StripWs& x) { return x.base && x.cap_images >= n_images && x.hw >= hw && x.nslot == nslot; };
  if (!fits(w)) {
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &cs) != cudaSuccess) { cudaGetLastError(); }
    if (cs != cudaStreamCaptureStatusNone) {
      StripWs& sp = g_spare[dev];
      if (w.base || !fits(sp)) { *unavailable = 1; return B2S_OK; }
      w = sp; sp = StripWs();
    } else {
      if (w.base) { B2S_CUDA(cudaFree(w.base)); w = StripWs(); }
      const int rc = alloc_ws(n_images, hw, nslot, w);
      if (rc) return rc;
    }
  }
  ws = w;
  {
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &cs) != cudaSuccess) { cudaGetLastError(); }
    StripWs& sp = g_spare[dev];
    if (cs == cudaStreamCaptureStatusNone && !fits(sp)) {
      if (sp.base) { B2S_CUDA(cudaFree(sp.base)); sp = StripWs(); }
      const int rc = alloc_ws(n_images, hw, nslot, sp);
      if (rc) return rc;
    }
  }
  return B2S_OK;
}

template <class DW, class DH, class Pro, class Epi>
int launch_strip(const Pro& pro, const Epi& epi, float scale, int64_t n_images, cudaStream_t st, int* unavailable) {
  *unavailable = 0;
  if (n_images <= 0) return B2S_OK;
  constexpr int NT = 32 * STRIP_WARPS, MINB = 4;          // 16 autonomous warps x <= 128 registers per SM
  constexpr int UPP = DH::N / STRIP_TV + DW::N / STRIP_TV;
  const int lag = (int)(strip_lag() < n_images ? strip_lag() : n_images);
  if ((n_images + lag) * (long long)UPP > 0x7fffffffLL) return fail(B2S_EUNSUPPORTED, "too many images for one launch");
  StripWs ws;
  const int rc = get_ws(st, n_images, (long long)DW::N * DH::N, ws, unavailable);
  if (rc || *unavailable) return rc;
  auto kern = strip_fft2_kernel<DW, DH, Pro, Epi, MINB>;
  int dev = 0;
  B2S_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) return fail(B2S_EUNSUPPORTED, "device index >= 64");
  static std::atomic<int> slots[64];                      // resident CTAs per device for this instantiation
  if (!slots[dev].load()) {
    int sms = 0, per_sm = 0;
    B2S_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    B2S_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    B2S_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, NT, 0));
    if (per_sm < 1) return fail(B2S_ECUDA, "strip kernel does not fit on an SM");
    slots[dev].store(sms * per_sm);
  }
  const long long total = (n_images + lag) * (long long)UPP;                       // tickets; one warp per ticket at a time
  const long long want = (total + STRIP_WARPS - 1) / STRIP_WARPS;
  const unsigned grid = (unsigned)(want < slots[dev].load() ? want : slots[dev].load());
  B2S_CUDA(cudaMemsetAsync(ws.head(), 0, ws.counter_bytes(), st));
  StripArgs a;
  a.scratch = ws.scratch(); a.head = ws.head(); a.status = ws.status(); a.done_r = ws.done_r(); a.done_c = ws.done_c();
  a.n_images = (int)n_images; a.lag = lag; a.nslot = ws.nslot; a.scale = scale;
  kern<<<grid, NT, 0, st>>>(pro, epi, a);
  return check_launch("strip_fft2_kernel");
}

template <int N> float strip_sign() { return ((N / 2) & 1) ? -1.f : 1.f; }

}  // namespace

namespace b2s {

// (selected by b2s_set_fused_path(1) / B2S_STRIP=1 in experimental builds: b2s_fused_experiments.inc)

template <int H, int W>
int strip_fft2c_t(const float* in, float* out, int64_t n, int inverse, float scale, cudaStream_t st, int* un) {
  typedef StripDim<W> DW; typedef StripDim<H> DH;
  const long long hw = (long long)H * W;
  const float s = scale * strip_sign<H>() * strip_sign<W>();
  if (inverse) return launch_strip<DW, DH>(SProPlain<true>{(const cfloat*)in, hw}, SEpiPlain<true>{(cfloat*)out, hw, W}, s, n, st, un);
  return launch_strip<DW, DH>(SProPlain<false>{(const cfloat*)in, hw}, SEpiPlain<false>{(cfloat*)out, hw, W}, s, n, st, un);
}

template <int H, int W>
int strip_expand_t(const float* image, const float* sens, float* kspace, const float* ref, const uint8_t* mask,
                   const float* v, int mode, int t, int c, int64_t n, float scale, cudaStream_t st, int* un) {
  typedef StripDim<W> DW; typedef StripDim<H> DH;
  const long long hw = (long long)H * W;
  const float s = scale * strip_sign<H>() * strip_sign<W>();
  SProExpand pro{(const cfloat*)image, (const cfloat*)sens, t, c, hw};
#define B2S_RUN(M) return launch_strip<DW, DH>(pro, SEpiKspace<M>{(cfloat*)kspace, (const cfloat*)ref, mask, v, c, H, W, hw}, s, n, st, un);
  switch (mode) { case 0: B2S_RUN(0) case 1: B2S_RUN(1) case 2: B2S_RUN(2) default: B2S_RUN(3) }
#undef B2S_RUN
}

template <int H, int W>
int strip_reduce_t(const float* kspace, const float* mult, float* out, const uint8_t* mask, const float* v,
                   int weight_mode, int over_frames, int t, int c, int64_t n, float scale, cudaStream_t st, int* un) {
  typedef StripDim<W> DW; typedef StripDim<H> DH;
  const long long hw = (long long)H * W;
  const float s = scale * strip_sign<H>() * strip_sign<W>();
  SEpiReduce epi;
  epi.out = (cfloat*)out; epi.mult = (const cfloat*)mult; epi.T = t; epi.C = c; epi.W = W;
  if (!over_frames) { epi.os_b = t * hw; epi.os_t = hw; epi.os_c = 0; epi.ms_b = c * hw; epi.ms_t = 0; epi.ms_c = hw; }
  else              { epi.os_b = c * hw; epi.os_t = 0; epi.os_c = hw; epi.ms_b = t * hw; epi.ms_t = hw; epi.ms_c = 0; }
#define B2S_RUN(M) return launch_strip<DW, DH>(SProKspace<M>{(const cfloat*)kspace, mask, v, c, H, hw}, epi, s, n, st, un);
  switch (weight_mode) { case 0: B2S_RUN(0) case 1: B2S_RUN(1) default: B2S_RUN(2) }
#undef B2S_RUN
}

template <int H, int W>
int strip_ifft_weighted_t(const float* kspace, float* y, const uint8_t* mask, const float* v, int weight_mode, int c,
                          int64_t n, float scale, cudaStream_t st, int* un) {
  typedef StripDim<W> DW; typedef StripDim<H> DH;
  const long long hw = (long long)H * W;
  const float s = scale * strip_sign<H>() * strip_sign<W>();
  SEpiPlain<true> epi{(cfloat*)y, hw, W};
#define B2S_RUN(M) return launch_strip<DW, DH>(SProKspace<M>{(const cfloat*)kspace, mask, v, c, H, hw}, epi, s, n, st, un);
  switch (weight_mode) { case 0: B2S_RUN(0) case 1: B2S_RUN(1) default: B2S_RUN(2) }
#undef B2S_RUN
}

int strip_fft2c(int h, const float* in, float* out, int64_t n, int inverse, float scale, cudaStream_t st, int* un) {
  return h == 200 ? strip_fft2c_t<200, 200>(in, out, n, inverse, scale, st, un) : strip_fft2c_t<256, 256>(in, out, n, inverse, scale, st, un);
}
int strip_expand(int h, const float* image, const float* sens, float* kspace, const float* ref, const uint8_t* mask,
                 const float* v, int mode, int t, int c, int64_t n, float scale, cudaStream_t st, int* un) {
  return h == 200 ? strip_expand_t<200, 200>(image, sens, kspace, ref, mask, v, mode, t, c, n, scale, st, un)
                  : strip_expand_t<256, 256>(image, sens, kspace, ref, mask, v, mode, t, c, n, scale, st, un);
}
int strip_reduce(int h, const float* kspace, const float* mult, float* out, const uint8_t* mask, const float* v,
                 int weight_mode, int over_frames, int t, int c, int64_t n, float scale, cudaStream_t st, int* un) {
  return h == 200 ? strip_reduce_t<200, 200>(kspace, mult, out, mask, v, weight_mode, over_frames, t, c, n, scale, st, un)
                  : strip_reduce_t<256, 256>(kspace, mult, out, mask, v, weight_mode, over_frames, t, c, n, scale, st, un);
}
int strip_ifft_weighted(int h, const float* kspace, float* y, const uint8_t* mask, const float* v, int weight_mode, int c,
                        int64_t n, float scale, cudaStream_t st, int* un) {
  return h == 200 ? strip_ifft_weighted_t<200, 200>(kspace, y, mask, v, weight_mode, c, n, scale, st, un)
                  : strip_ifft_weighted_t<256, 256>(kspace, y, mask, v, weight_mode, c, n, scale, st, un);
}

}  // namespace b2s

// 0 = no dependency wait ever timed out on any workspace of this process (synchronises the device)
extern "C" int b2s_debug_strip_status(void) {
  if (cudaDeviceSynchronize() != cudaSuccess) return -1;
  std::lock_guard<std::mutex> lock(g_ws_mutex);
  int dev = 0, bad = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return -1;
  for (auto& kv : g_ws) {
    if (kv.first.first != dev || !kv.second.base) continue;
    int s = 0;
    if (cudaMemcpy(&s, kv.second.status(), sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    bad |= s;
  }
  return bad;
}
