// Version / error plumbing of the C ABI (include/b200sense.h).
#include <stdarg.h>
#include "b2s_common.cuh"

namespace b2s {
static thread_local char g_err[512] = "";
std::atomic<unsigned long long> g_kernel_launches{0};
std::atomic<int> g_sm_reserve{0};
std::atomic<int> g_fused_path{0};
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace b2s

extern "C" int b2s_version(void) { return 100; }
extern "C" const char* b2s_last_error(void) { return b2s::g_err; }
extern "C" int b2s_set_sm_reserve(int n_sms) {
  if (n_sms < 0 || n_sms > 64) { b2s::set_error("b2s_set_sm_reserve: 0 <= n_sms <= 64"); return B2S_EINVAL; }
  b2s::g_sm_reserve.store(n_sms);
  return B2S_OK;
}
extern "C" unsigned long long b2s_launch_count(int reset) {
  return reset ? b2s::g_kernel_launches.exchange(0) : b2s::g_kernel_launches.load();
}

// Kernel family behind the fused plan sizes: 0 / -1 auto (default), 1 strip-streamed (experimental builds only),
// 2 half/quarter-split kernels only, 3 packed whole-image kernel wherever it exists.  A test / measurement knob.
extern "C" int b2s_set_fused_path(int path) {
  if (path == -1) path = 0;
  if (path < 0 || path > 3) { b2s::set_error("b2s_set_fused_path: path must be -1/0 (auto), 1 (strip), 2 (half split) or 3 (packed)"); return B2S_EINVAL; }
#ifndef B2S_EXPERIMENTS
  if (path == 1) { b2s::set_error("b2s_set_fused_path: the strip-streamed kernels are only in experimental builds (make EXPERIMENTS=1)"); return B2S_EUNSUPPORTED; }
#endif
  b2s::g_fused_path.store(path);
  return B2S_OK;
}
#ifndef B2S_EXPERIMENTS
extern "C" int b2s_debug_strip_status(void) { return 0; }
#endif
