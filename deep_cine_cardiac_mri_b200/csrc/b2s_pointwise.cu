// Memory-bound element-wise / small-reduction kernels of the SENSE path:
// stand-alone DC blend (+backward), the fastMRI complex helpers, RSS, the
// SensitivityModel pre/post ops, the xfyf temporal head/tail and the CG vector
// kernels.  All are grid-stride, 64/128-bit vectorised, one pass over HBM.
#include "b2s_common.cuh"
#include "fft2_core.cuh"
#include "codelets.cuh"

using namespace b2s;

namespace {

constexpr int NT = 256;
inline unsigned grid_for(long long n, int per_block = NT, long long cap = 148LL * 16) {
  long long g = (n + per_block - 1) / per_block;
  if (g < 1) g = 1;
  if (g > cap) g = cap;
  return (unsigned)g;
}
#define GRID_STRIDE(i, n) for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < (n); i += (long long)gridDim.x * blockDim.x)

__device__ __forceinline__ float block_sum(float v) {
  __shared__ float red[32];
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) red[wid] = v;
  __syncthreads();
  v = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.f;
  if (wid == 0) for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;   // valid in thread 0
}

// ---------------- generic-size fallbacks of the fused operators ---------------- //
__global__ void expand_product_kernel(const cfloat* img, const cfloat* sens, cfloat* out, int T, int C,
                                      long long hw, long long n) {
  GRID_STRIDE(i, n) {
    const long long pix = i % hw, im = i / hw, c = im % C, bt = im / C, b = bt / T;
    const cfloat a = img[bt * hw + pix], s = sens[(b * C + c) * hw + pix];
    out[i] = make_c(a.x * s.x - a.y * s.y, a.x * s.y + a.y * s.x);
  }
}

__global__ void kspace_epilogue_kernel(cfloat* k, const cfloat* ref, const uint8_t* mask, const float* vptr,
                                       int mode, int C, int H, int W, long long n) {
  const float v = (mode == 2) ? *vptr : 0.f;
  GRID_STRIDE(i, n) {
    const long long row = i / W, y = row % H, bt = row / H / C;
    const bool m = mask[bt * H + y] != 0;
    cfloat z = k[i];
    if (mode == 1) { if (!m) z = make_c(0.f, 0.f); }
    else if (mode == 2) { if (m) { const cfloat r = ref[i]; z = make_c((z.x + v * r.x) / (1.f + v), (z.y + v * r.y) / (1.f + v)); } }
    else { const cfloat r = ref[i]; if (!m) z = make_c(0.f, 0.f); z = make_c(z.x - r.x, z.y - r.y); }
    k[i] = z;
  }
}

__global__ void row_weight_kernel(const cfloat* k, cfloat* out, const uint8_t* mask, const float* vptr,
                                  int wmode, int C, int H, int W, long long n) {
  float wa = 0.f, wb = 1.f;
  if (wmode == 2) { const float v = *vptr; wa = 1.f; wb = -v / (1.f + v); }
  GRID_STRIDE(i, n) {
    const long long row = i / W, y = row % H, bt = row / H / C;
    const float wgt = wa + wb * (float)mask[bt * H + y];
    const cfloat z = k[i];
    out[i] = make_c(z.x * wgt + 0.f, z.y * wgt + 0.f);      // (+ 0.0 as the reference: no negative zeros, transforms.py:90)
  }
}

__global__ void coil_reduce_kernel(const cfloat* y, const cfloat* mult, cfloat* out, int over_frames, int T,
                                   int C, long long hw, long long n_out) {
  GRID_STRIDE(i, n_out) {
    const long long pix = i % hw, oi = i / hw;     // oi = b*T+t (coil sum) or b*C+c (frame sum)
    float ar = 0.f, ai = 0.f;
    if (!over_frames) {
      const long long b = oi / T;
      for (int c = 0; c < C; ++c) {
        const cfloat v = y[(oi * C + c) * hw + pix], s = mult[(b * C + c) * hw + pix];
        ar += v.x * s.x + v.y * s.y; ai += v.y * s.x - v.x * s.y;
      }
    } else {
      const long long b = oi / C, c = oi % C;
      for (int t = 0; t < T; ++t) {
        const cfloat v = y[((b * T + t) * C + c) * hw + pix], s = mult[(b * T + t) * hw + pix];
        ar += v.x * s.x + v.y * s.y; ai += v.y * s.x - v.x * s.y;
      }
    }
    out[i] = make_c(ar, ai);
  }
}

// ------------------------- sparse upload of masked k-space ------------------ //
// One warp per k-space row: sampled rows (mask = 1) are read straight from pinned host memory (UVA: the
// host pointer is valid on the device) and land in the dense device tensor, unsampled rows are written as
// zeros.  Only the sampled quarter of the k-space crosses PCIe, and no host thread packs anything.
template <class V>
__global__ void upload_rows_kernel(const V* __restrict__ src, const uint8_t* __restrict__ mask, V* __restrict__ dst, int C, int H,
                                   int WV, long long n_rows) {
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  for (long long row = (long long)blockIdx.x * wpb + (threadIdx.x >> 5); row < n_rows; row += (long long)gridDim.x * wpb) {
    const long long bt = row / ((long long)C * H);
    const int y = (int)(row % H);
    const bool m = mask[bt * H + y] != 0;
    const V* s = src + row * WV;
    V* d = dst + row * WV;
    V zero; memset(&zero, 0, sizeof(V));
    for (int i0 = 0; i0 < WV; i0 += 128) {                 // 4 independent host reads per lane in flight
      V v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) { const int i = i0 + lane + 32 * u; v[u] = (m && i < WV) ? s[i] : zero; }
#pragma unroll
      for (int u = 0; u < 4; ++u) { const int i = i0 + lane + 32 * u; if (i < WV) d[i] = v[u]; }
    }
  }
}

// ------------------------------- DC blend ---------------------------------- //
// V = float4 (two complex per thread, even w) or cfloat (odd w)
template <class V> struct VecOps;
template <> struct VecOps<float4> {
  static constexpr int N = 4;
  static __device__ __forceinline__ void get(const float4& v, float* f) { f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w; }
  static __device__ __forceinline__ float4 put(const float* f) { return make_float4(f[0], f[1], f[2], f[3]); }
};
template <> struct VecOps<cfloat> {
  static constexpr int N = 2;
  static __device__ __forceinline__ void get(const cfloat& v, float* f) { f[0] = v.x; f[1] = v.y; }
  static __device__ __forceinline__ cfloat put(const float* f) { return make_c(f[0], f[1]); }
};

template <class V>
__global__ void dc_blend_kernel(const V* k, const V* ref, const uint8_t* mask, const float* vptr, V* out, int C, int H,
                                int WV, long long nv) {
  const float v = *vptr;
  GRID_STRIDE(i, nv) {
    const long long row = i / WV, y = row % H, bt = row / H / C;
    float z[VecOps<V>::N], r[VecOps<V>::N];
    VecOps<V>::get(k[i], z);
    if (mask[bt * H + y]) {
      VecOps<V>::get(ref[i], r);
#pragma unroll
      for (int e = 0; e < VecOps<V>::N; ++e) z[e] = (z[e] + v * r[e]) / (1.f + v);
    }
    out[i] = VecOps<V>::put(z);
  }
}

template <class V>
__global__ void dc_blend_bwd_kernel(const V* g, const V* outv, const V* ref, const uint8_t* mask, const float* vptr,
                                    V* gk, V* gref, float* gv, int C, int H, int WV, long long nv) {
  const float v = *vptr, eta = v / (1.f + v), inv1 = 1.f / (1.f + v);
  float acc = 0.f;
  GRID_STRIDE(i, nv) {
    const long long row = i / WV, y = row % H, bt = row / H / C;
    float gg[VecOps<V>::N], t[VecOps<V>::N];
    VecOps<V>::get(g[i], gg);
    const bool m = mask[bt * H + y] != 0;
    const float a = m ? (1.f - eta) : 1.f, bb = m ? eta : 0.f;
    if (gk) {
#pragma unroll
      for (int e = 0; e < VecOps<V>::N; ++e) t[e] = gg[e] * a;
      gk[i] = VecOps<V>::put(t);
    }
    if (gref) {
#pragma unroll
      for (int e = 0; e < VecOps<V>::N; ++e) t[e] = gg[e] * bb;
      gref[i] = VecOps<V>::put(t);
    }
    if (gv && m) {
      float r[VecOps<V>::N], o[VecOps<V>::N];
      VecOps<V>::get(ref[i], r); VecOps<V>::get(outv[i], o);
#pragma unroll
      for (int e = 0; e < VecOps<V>::N; ++e) acc += gg[e] * (r[e] - o[e]) * inv1;
    }
  }
  if (gv) {                                   // ordered two-stage sum: partials here, dot_final_kernel after
    acc = block_sum(acc);
    if (threadIdx.x == 0) gv[B2S_DC_BWD_GV_FLOATS - 1024 + blockIdx.x] = acc;
  }
}

// ---------------------------- complex helpers ------------------------------ //
struct MulDims { int nd; long long shape[6], sa[6], sb[6]; };

__global__ void complex_mul_kernel(const cfloat* a, const cfloat* b, cfloat* out, MulDims d, int conj_b, long long n) {
  GRID_STRIDE(i, n) {
    long long rem = i, oa = 0, ob = 0;
#pragma unroll
    for (int k = 5; k >= 0; --k) {
      if (k < d.nd) { const long long idx = rem % d.shape[k]; rem /= d.shape[k]; oa += idx * d.sa[k]; ob += idx * d.sb[k]; }
    }
    const cfloat x = a[oa]; cfloat y = b[ob];
    if (conj_b) y.y = -y.y;
    out[i] = make_c(x.x * y.x - x.y * y.y, x.x * y.y + x.y * y.x);
  }
}

__global__ void complex_conj_kernel(const cfloat* in, cfloat* out, long long n) {
  GRID_STRIDE(i, n) { const cfloat v = in[i]; out[i] = make_c(v.x, -v.y); }
}

__global__ void complex_abs_kernel(const cfloat* in, float* out, long long n, int squared) {
  GRID_STRIDE(i, n) { const cfloat v = in[i]; const float s = v.x * v.x + v.y * v.y; out[i] = squared ? s : sqrtf(s); }
}

__global__ void rss_kernel(const float* in, float* out, long long outer, long long r, long long inner, int is_complex) {
  const long long n = outer * inner;
  GRID_STRIDE(i, n) {
    const long long o = i / inner, p = i % inner;
    float acc = 0.f;
    if (is_complex) {
      const cfloat* z = reinterpret_cast<const cfloat*>(in);
      for (long long k = 0; k < r; ++k) { const cfloat v = z[(o * r + k) * inner + p]; acc += v.x * v.x + v.y * v.y; }
    } else {
      for (long long k = 0; k < r; ++k) { const float v = in[(o * r + k) * inner + p]; acc += v * v; }
    }
    out[i] = sqrtf(acc);
  }
}

// ------------------------- SensitivityModel pre/post ----------------------- //
// ACS window exactly as models/varnet.py:64-68 on frame 0 of each batch element.
__device__ __forceinline__ void acs_window(const uint8_t* m0, int H, int& pad, int& nlf) {
  const int cent = H / 2;
  int left = -1, right = H;
  for (int y = 0; y < cent; ++y) if (m0[y] == 0) left = y;              // last zero in [:cent]
  for (int y = H - 1; y >= cent; --y) if (m0[y] == 0) right = y;        // first zero in [cent:]
  nlf = right - left;
  pad = (H - nlf + 1) / 2;
}

__global__ void acs_mean_kernel(const cfloat* k, const uint8_t* mask, cfloat* out, int32_t* window_out, int T, int C,
                                int H, int W) {
  // blockIdx.y = batch element; the window is computed once per block
  __shared__ int s_pad, s_nlf;
  const int b = blockIdx.y;
  if (threadIdx.x == 0) {
    int pad, nlf; acs_window(mask + (long long)b * T * H, H, pad, nlf);
    s_pad = pad; s_nlf = nlf;
    if (window_out && blockIdx.x == 0) { window_out[2 * b] = pad; window_out[2 * b + 1] = nlf; }
  }
  __syncthreads();
  const int pad = s_pad, nlf = s_nlf;
  const long long hw = (long long)H * W, per_b = (long long)C * hw;
  const float invT = 1.f / (float)T;
  GRID_STRIDE(i, per_b) {
    const long long c = i / hw, yx = i % hw;
    const int y = (int)(yx / W);
    cfloat acc = make_c(0.f, 0.f);
    if (y >= pad && y < pad + nlf) {
      for (int t = 0; t < T; ++t) { const cfloat v = k[(((long long)b * T + t) * C + c) * hw + yx]; acc.x += v.x; acc.y += v.y; }
      acc.x *= invT; acc.y *= invT;
    }
    out[(long long)b * per_b + i] = acc;
  }
}

__global__ void rss_normalize_kernel(const cfloat* in, cfloat* out, int C, long long hw, long long n_pix) {
  GRID_STRIDE(i, n_pix) {
    const long long b = i / hw, p = i % hw;
    float acc = 0.f;
    for (int c = 0; c < C; ++c) { const cfloat v = in[(b * C + c) * hw + p]; acc += v.x * v.x + v.y * v.y; }
    const float r = sqrtf(acc);
    for (int c = 0; c < C; ++c) { const cfloat v = in[(b * C + c) * hw + p]; out[(b * C + c) * hw + p] = make_c(v.x / r, v.y / r); }
  }
}

__global__ void rss_normalize_bwd_kernel(const cfloat* g, const cfloat* in, cfloat* gin, int C, long long hw, long long n_pix) {
  GRID_STRIDE(i, n_pix) {
    const long long b = i / hw, p = i % hw;
    float ss = 0.f, dot = 0.f;
    for (int c = 0; c < C; ++c) {
      const cfloat v = in[(b * C + c) * hw + p], gg = g[(b * C + c) * hw + p];
      ss += v.x * v.x + v.y * v.y; dot += v.x * gg.x + v.y * gg.y;
    }
    const float r = sqrtf(ss), f = dot / (ss * r), ir = 1.f / r;
    for (int c = 0; c < C; ++c) {
      const cfloat v = in[(b * C + c) * hw + p], gg = g[(b * C + c) * hw + p];
      gin[(b * C + c) * hw + p] = make_c(gg.x * ir - v.x * f, gg.y * ir - v.y * f);
    }
  }
}

// ----------------------------- temporal head/tail -------------------------- //
// one thread per pixel; the t samples of blockDim pixels live in shared memory [t][blockDim]
__global__ void temporal_kernel(const cfloat* in, const cfloat* mean_in, cfloat* out, cfloat* mean_out, int T, long long hw,
                                long long n_pix, int xf, int post) {
  extern __shared__ __align__(16) unsigned char raw[];
  cfloat* tile = reinterpret_cast<cfloat*>(raw);                 // [T][blockDim]
  cfloat* tw = tile + (size_t)T * blockDim.x;                    // [T] forward twiddles
  for (int i = threadIdx.x; i < T; i += blockDim.x) tw[i] = twiddle(i, T);
  __syncthreads();
  const int s_in = (T + 1) / 2, s_out = T / 2;
  const float sc = xf ? rsqrtf((float)T) : 1.f;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_pix) return;
  const long long b = i / hw, p = i % hw;
  const cfloat* src = in + b * T * hw + p;
  cfloat* dst = out + b * T * hw + p;
  cfloat mu = make_c(0.f, 0.f);
  if (!post) {
    for (int t = 0; t < T; ++t) { const cfloat v = src[(long long)t * hw]; tile[t * blockDim.x + threadIdx.x] = v; mu.x += v.x; mu.y += v.y; }
    mu.x /= (float)T; mu.y /= (float)T;
    mean_out[i] = mu;
  } else {
    mu = mean_in[i];
    for (int t = 0; t < T; ++t) tile[t * blockDim.x + threadIdx.x] = src[(long long)t * hw];
  }
  if (!xf) {
    for (int t = 0; t < T; ++t) {
      const cfloat v = tile[t * blockDim.x + threadIdx.x];
      dst[(long long)t * hw] = post ? make_c(v.x + mu.x, v.y + mu.y) : make_c(v.x - mu.x, v.y - mu.y);
    }
    return;
  }
  for (int kk = 0; kk < T; ++kk) {
    int kp = kk - s_out; if (kp < 0) kp += T;
    float ar = 0.f, ai = 0.f;
    int ph = (s_in * kp) % T;                                    // ((j + s_in) * kp) mod T, j = 0
    for (int j = 0; j < T; ++j) {
      cfloat v = tile[j * blockDim.x + threadIdx.x];
      if (!post) { v.x -= mu.x; v.y -= mu.y; }
      const cfloat wv = tw[ph];
      const float wy = post ? -wv.y : wv.y;                      // inverse = conjugate twiddles
      ar += v.x * wv.x - v.y * wy; ai += v.x * wy + v.y * wv.x;
      ph += kp; if (ph >= T) ph -= T;
    }
    dst[(long long)kk * hw] = post ? make_c(ar * sc + mu.x, ai * sc + mu.y) : make_c(ar * sc, ai * sc);
  }
}

// Register-codelet version for the common frame counts: each thread holds its pixel's T samples,
// the centring rolls are compile-time index maps, the inverse uses the re/im swap.
template <int T>
__global__ void __launch_bounds__(128) temporal_codelet_kernel(const cfloat* __restrict__ in, const cfloat* __restrict__ mean_in,
                                                               cfloat* __restrict__ out, cfloat* __restrict__ mean_out,
                                                               long long hw, long long n_pix, int xf, int post) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_pix) return;
  const long long b = i / hw, p = i % hw;
  const cfloat* src = in + b * T * hw + p;
  cfloat* dst = out + b * T * hw + p;
  constexpr int S_IN = (T + 1) / 2, S_OUT = T / 2;
  float re[T], im[T];
  cfloat mu = make_c(0.f, 0.f);
  cfloat v[T];
#pragma unroll
  for (int t = 0; t < T; ++t) v[t] = src[(long long)t * hw];
  if (!post) {
#pragma unroll
    for (int t = 0; t < T; ++t) { mu.x += v[t].x; mu.y += v[t].y; }
    mu.x *= (1.f / (float)T); mu.y *= (1.f / (float)T);
    mean_out[i] = mu;
#pragma unroll
    for (int t = 0; t < T; ++t) { v[t].x -= mu.x; v[t].y -= mu.y; }
  } else {
    mu = mean_in[i];
  }
  if (!xf) {
#pragma unroll
    for (int t = 0; t < T; ++t) dst[(long long)t * hw] = post ? make_c(v[t].x + mu.x, v[t].y + mu.y) : v[t];
    return;
  }
  // z[e] = x[(e - S_IN) mod T]
#pragma unroll
  for (int e = 0; e < T; ++e) {
    const int t = (e - S_IN + T) % T;
    re[e] = post ? v[t].y : v[t].x;               // inverse: swapped
    im[e] = post ? v[t].x : v[t].y;
  }
  Dft<T>::run(re, im);
  const float sc = rsqrtf((float)T);
#pragma unroll
  for (int kk = 0; kk < T; ++kk) {                 // Y[kk] = X[(kk - S_OUT) mod T]
    const int k = (kk - S_OUT + T) % T;
    const float a = (post ? im[k] : re[k]) * sc, c = (post ? re[k] : im[k]) * sc;
    dst[(long long)kk * hw] = post ? make_c(a + mu.x, c + mu.y) : make_c(a, c);
  }
}

template <int T>
static int launch_temporal_codelet(const float* in, const float* mean_in, float* out, float* mean_out, long long hw,
                                   long long n_pix, int xf, int post, cudaStream_t st) {
  temporal_codelet_kernel<T><<<(unsigned)((n_pix + 127) / 128), 128, 0, st>>>((const cfloat*)in, (const cfloat*)mean_in, (cfloat*)out,
                                                                            (cfloat*)mean_out, hw, n_pix, xf, post);
  return check_launch("temporal_codelet_kernel");
}

// --------------------------------- CG kernels ------------------------------ //
__global__ void dot_partial_kernel(const float* a, const float* b, float* partial, long long n) {
  float acc = 0.f;
  GRID_STRIDE(i, n) acc += a[i] * b[i];
  acc = block_sum(acc);
  if (threadIdx.x == 0) partial[blockIdx.x] = acc;
}
__global__ void dot_final_kernel(const float* partial, float* out, int n) {
  float acc = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) acc += partial[i];
  acc = block_sum(acc);
  if (threadIdx.x == 0) out[0] = acc;
}
__global__ void axpy_ratio_kernel(float* y, const float* x, const float* num, const float* den, float sign, long long n) {
  const float a = sign * (num[0] / den[0]);
  GRID_STRIDE(i, n) y[i] = y[i] + a * x[i];
}
__global__ void xpay_ratio_kernel(float* p, const float* r, const float* num, const float* den, long long n) {
  const float bta = num[0] / den[0];
  GRID_STRIDE(i, n) p[i] = r[i] + bta * p[i];
}
// sum of n partials in a fixed order, result broadcast to the whole block
__device__ __forceinline__ float block_total(const float* __restrict__ part, int n) {
  __shared__ float tot;
  float acc = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) acc += part[i];
  acc = block_sum(acc);
  if (threadIdx.x == 0) tot = acc;
  __syncthreads();
  return tot;
}
// one CG iteration after d = H p (cinenet.py:159-164): alpha = rs_old / <p, d>;  x += alpha p;  r -= alpha d;  partials of <r, r>
__global__ void cg_update_kernel(const float* __restrict__ p, const float* __restrict__ d, float* __restrict__ x, float* __restrict__ r,
                                 const float* __restrict__ pd_part, int n_pd, const float* __restrict__ rs_old, float* __restrict__ rr_part, long long n) {
  const float alpha = rs_old[0] / block_total(pd_part, n_pd);
  float acc = 0.f;
  GRID_STRIDE(i, n) {
    x[i] = x[i] + alpha * p[i];
    const float rn = r[i] - alpha * d[i];
    r[i] = rn;
    acc += rn * rn;
  }
  acc = block_sum(acc);
  if (threadIdx.x == 0) rr_part[blockIdx.x] = acc;
}
// ... and its second half (cinenet.py:165-167): rs_new = <r, r>;  p = r + (rs_new / rs_old) p
__global__ void cg_direction_kernel(float* __restrict__ p, const float* __restrict__ r, const float* __restrict__ rr_part, int n_rr,
                                    const float* __restrict__ rs_old, float* __restrict__ rs_new, long long n) {
  const float rn = block_total(rr_part, n_rr);
  const float beta = rn / rs_old[0];
  if (blockIdx.x == 0 && threadIdx.x == 0) rs_new[0] = rn;
  GRID_STRIDE(i, n) p[i] = r[i] + beta * p[i];
}
__global__ void axpby_kernel(const float* a, const float* b, const float* v, float scale, float* out, long long n) {
  const float s = v ? v[0] * scale : scale;
  GRID_STRIDE(i, n) out[i] = a[i] + s * b[i];
}

}  // namespace

namespace b2s {

int launch_expand_product(const float* image, const float* sens, float* out, int b, int t, int c, int64_t hw, cudaStream_t st) {
  const long long n = (long long)b * t * c * hw;
  if (n == 0) return B2S_OK;
  expand_product_kernel<<<grid_for(n), NT, 0, st>>>((const cfloat*)image, (const cfloat*)sens, (cfloat*)out, t, c, hw, n);
  return check_launch("expand_product_kernel");
}
int launch_kspace_epilogue(float* k, const float* ref, const uint8_t* mask, const float* v, int mode, int64_t n_bt, int c, int h, int w, cudaStream_t st) {
  const long long n = n_bt * c * h * w;
  if (n == 0) return B2S_OK;
  kspace_epilogue_kernel<<<grid_for(n), NT, 0, st>>>((cfloat*)k, (const cfloat*)ref, mask, v, mode, c, h, w, n);
  return check_launch("kspace_epilogue_kernel");
}
int launch_row_weight(const float* k, float* out, const uint8_t* mask, const float* v, int wmode, int64_t n_bt, int c, int h, int w, cudaStream_t st) {
  const long long n = n_bt * c * h * w;
  if (n == 0) return B2S_OK;
  row_weight_kernel<<<grid_for(n), NT, 0, st>>>((const cfloat*)k, (cfloat*)out, mask, v, wmode, c, h, w, n);
  return check_launch("row_weight_kernel");
}
int launch_coil_reduce(const float* y, const float* mult, float* out, int over_frames, int b, int t, int c, int64_t hw, cudaStream_t st) {
  const long long n_out = (over_frames ? (long long)b * c : (long long)b * t) * hw;
  if (n_out == 0) return B2S_OK;
  coil_reduce_kernel<<<grid_for(n_out), NT, 0, st>>>((const cfloat*)y, (const cfloat*)mult, (cfloat*)out, over_frames, t, c, hw, n_out);
  return check_launch("coil_reduce_kernel");
}

}  // namespace b2s

// apply_mask (data/transforms.py:66-92): out = k * m + 0.0 with the (b,t,h) row mask broadcast over coils and columns
extern "C" int b2s_apply_mask(const float* kspace, const uint8_t* mask, float* out, int64_t n_bt, int c, int h, int w, void* stream) {
  if (n_bt < 0 || c < 0 || h < 0 || w < 0) return fail(B2S_EINVAL, "b2s_apply_mask: bad argument");
  if (n_bt * c * (long long)h * w == 0) return B2S_OK;
  if (!kspace || !mask || !out) return fail(B2S_EINVAL, "b2s_apply_mask: null pointer");
  return launch_row_weight(kspace, out, mask, nullptr, 1, n_bt, c, h, w, (cudaStream_t)stream);
}

extern "C" int b2s_upload_rows(const float* kspace_host, const uint8_t* mask, float* kspace_dev, int64_t n_bt, int c, int h,
                               int w, void* stream) {
  if (n_bt < 0 || c < 0 || h < 0 || w < 0) return fail(B2S_EINVAL, "b2s_upload_rows: bad argument");
  const long long n_rows = n_bt * c * (long long)h;
  if (n_rows == 0 || w == 0) return B2S_OK;
  if (!kspace_host || !mask || !kspace_dev) return fail(B2S_EINVAL, "b2s_upload_rows: null pointer");
  cudaPointerAttributes attr;
  if (cudaPointerGetAttributes(&attr, kspace_host) != cudaSuccess || attr.type != cudaMemoryTypeHost || !attr.devicePointer) {
    cudaGetLastError();
    return fail(B2S_EINVAL, "b2s_upload_rows: kspace_host must be pinned (page-locked, device-accessible) host memory");
  }
  const float* src = (const float*)attr.devicePointer;
  cudaStream_t st = (cudaStream_t)stream;
  // With SMs reserved (b2s_set_sm_reserve) the upload runs as that many 1024-thread CTAs on exactly those SMs, beside
  // the persistent compute kernels of another stream; otherwise it spreads over the whole GPU.
  const int reserve = g_sm_reserve.load();
  const unsigned grid = reserve > 0 ? (unsigned)reserve : 148u * 8u;
  const unsigned nt = reserve > 0 ? 1024u : (unsigned)NT;
  if (w % 2 == 0 && ((uintptr_t)src % 16 == 0) && ((uintptr_t)kspace_dev % 16 == 0))
    upload_rows_kernel<float4><<<grid, nt, 0, st>>>((const float4*)src, mask, (float4*)kspace_dev, c, h, w / 2, n_rows);
  else
    upload_rows_kernel<float2><<<grid, nt, 0, st>>>((const float2*)src, mask, (float2*)kspace_dev, c, h, w, n_rows);
  return check_launch("upload_rows_kernel");
}

extern "C" int b2s_dc_blend(const float* kspace, const float* ref, const uint8_t* mask, const float* v, float* out,
                            int64_t n_bt, int c, int h, int w, void* stream) {
  const long long n = n_bt * c * h * (long long)w;
  if (n == 0) return B2S_OK;
  if (!kspace || !ref || !mask || !v || !out) return fail(B2S_EINVAL, "b2s_dc_blend: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  if (w % 2 == 0)
    dc_blend_kernel<float4><<<grid_for(n / 2), NT, 0, st>>>((const float4*)kspace, (const float4*)ref, mask, v, (float4*)out, c, h, w / 2, n / 2);
  else
    dc_blend_kernel<cfloat><<<grid_for(n), NT, 0, st>>>((const cfloat*)kspace, (const cfloat*)ref, mask, v, (cfloat*)out, c, h, w, n);
  return check_launch("dc_blend_kernel");
}

extern "C" int b2s_dc_blend_bwd(const float* g, const float* out, const float* ref, const uint8_t* mask, const float* v,
                                float* gk, float* gref, float* gv, int64_t n_bt, int c, int h, int w, void* stream) {
  const long long n = n_bt * c * h * (long long)w;
  if (n == 0) return B2S_OK;
  if (!g || !mask || !v || (gv && (!out || !ref))) return fail(B2S_EINVAL, "b2s_dc_blend_bwd: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  const long long nv = (w % 2 == 0) ? n / 2 : n;
  const unsigned grid = grid_for(nv, NT, gv ? 1024 : 148LL * 16);
  if (w % 2 == 0)
    dc_blend_bwd_kernel<float4><<<grid, NT, 0, st>>>((const float4*)g, (const float4*)out, (const float4*)ref, mask, v,
                                                     (float4*)gk, (float4*)gref, gv, c, h, w / 2, nv);
  else
    dc_blend_bwd_kernel<cfloat><<<grid, NT, 0, st>>>((const cfloat*)g, (const cfloat*)out, (const cfloat*)ref, mask, v,
                                                     (cfloat*)gk, (cfloat*)gref, gv, c, h, w, nv);
  if (!gv) return check_launch("dc_blend_bwd_kernel");
  dot_final_kernel<<<1, NT, 0, st>>>(gv + B2S_DC_BWD_GV_FLOATS - 1024, gv, (int)grid);
  return check_launch("dc_blend_bwd kernels", 2);
}

extern "C" int b2s_complex_mul(const float* a, const float* b, float* out, int ndim, const int64_t* shape,
                               const int64_t* stride_a, const int64_t* stride_b, int conj_b, void* stream) {
  if (ndim < 0 || ndim > 6 || (ndim && (!shape || !stride_a || !stride_b))) return fail(B2S_EINVAL, "b2s_complex_mul: bad argument");
  MulDims d; d.nd = ndim; long long n = 1;
  for (int k = 0; k < 6; ++k) { d.shape[k] = 1; d.sa[k] = 0; d.sb[k] = 0; }
  for (int k = 0; k < ndim; ++k) { d.shape[k] = shape[k]; d.sa[k] = stride_a[k]; d.sb[k] = stride_b[k]; n *= shape[k]; }
  if (n == 0) return B2S_OK;
  if (!a || !b || !out) return fail(B2S_EINVAL, "b2s_complex_mul: null pointer");
  complex_mul_kernel<<<grid_for(n), NT, 0, (cudaStream_t)stream>>>((const cfloat*)a, (const cfloat*)b, (cfloat*)out, d, conj_b, n);
  return check_launch("complex_mul_kernel");
}

extern "C" int b2s_complex_conj(const float* in, float* out, int64_t n, void* stream) {
  if (n == 0) return B2S_OK;
  if (!in || !out) return fail(B2S_EINVAL, "b2s_complex_conj: null pointer");
  complex_conj_kernel<<<grid_for(n), NT, 0, (cudaStream_t)stream>>>((const cfloat*)in, (cfloat*)out, n);
  return check_launch("complex_conj_kernel");
}

extern "C" int b2s_complex_abs(const float* in, float* out, int64_t n, int squared, void* stream) {
  if (n == 0) return B2S_OK;
  if (!in || !out) return fail(B2S_EINVAL, "b2s_complex_abs: null pointer");
  complex_abs_kernel<<<grid_for(n), NT, 0, (cudaStream_t)stream>>>((const cfloat*)in, out, n, squared);
  return check_launch("complex_abs_kernel");
}

extern "C" int b2s_rss(const float* in, float* out, int64_t outer, int64_t r, int64_t inner, int is_complex, void* stream) {
  if (outer * inner == 0) return B2S_OK;
  if (!in || !out) return fail(B2S_EINVAL, "b2s_rss: null pointer");
  rss_kernel<<<grid_for(outer * inner), NT, 0, (cudaStream_t)stream>>>(in, out, outer, r, inner, is_complex);
  return check_launch("rss_kernel");
}

extern "C" int b2s_acs_mean(const float* kspace, const uint8_t* mask, float* out, int32_t* window_out, int b, int t, int c,
                            int h, int w, void* stream) {
  if (!kspace || !mask || !out) return fail(B2S_EINVAL, "b2s_acs_mean: null pointer");
  const long long per_b = (long long)c * h * w;
  if (per_b == 0 || b == 0) return B2S_OK;
  if (t < 1 || b > 65535) return fail(B2S_EUNSUPPORTED, "b2s_acs_mean: need t >= 1 and b <= 65535");
  const dim3 grid(grid_for(per_b, NT, 148 * 4), (unsigned)b);
  acs_mean_kernel<<<grid, NT, 0, (cudaStream_t)stream>>>((const cfloat*)kspace, mask, (cfloat*)out, window_out, t, c, h, w);
  return check_launch("acs_mean_kernel");
}

extern "C" int b2s_rss_normalize(const float* in, float* out, int b, int c, int64_t hw, void* stream) {
  if (!in || !out) return fail(B2S_EINVAL, "b2s_rss_normalize: null pointer");
  const long long n = (long long)b * hw;
  if (n == 0) return B2S_OK;
  rss_normalize_kernel<<<grid_for(n), NT, 0, (cudaStream_t)stream>>>((const cfloat*)in, (cfloat*)out, c, hw, n);
  return check_launch("rss_normalize_kernel");
}

extern "C" int b2s_rss_normalize_bwd(const float* g, const float* in, float* gin, int b, int c, int64_t hw, void* stream) {
  if (!g || !in || !gin) return fail(B2S_EINVAL, "b2s_rss_normalize_bwd: null pointer");
  const long long n = (long long)b * hw;
  if (n == 0) return B2S_OK;
  rss_normalize_bwd_kernel<<<grid_for(n), NT, 0, (cudaStream_t)stream>>>((const cfloat*)g, (const cfloat*)in, (cfloat*)gin, c, hw, n);
  return check_launch("rss_normalize_bwd_kernel");
}

static int launch_temporal(const float* in, const float* mean_in, float* out, float* mean_out, int b, int t, int64_t hw,
                           int xf, int post, void* stream) {
  if (t < 1 || t > 192) return fail(B2S_EUNSUPPORTED, "temporal length must be in [1, 192]");
  const long long n_pix = (long long)b * hw;
  if (n_pix == 0) return B2S_OK;
  switch (t) {                                             // register codelets for the usual frame counts
#define B2S_T(N) case N: return launch_temporal_codelet<N>(in, mean_in, out, mean_out, hw, n_pix, xf, post, (cudaStream_t)stream);
    B2S_T(12) B2S_T(15) B2S_T(16) B2S_T(20) B2S_T(24) B2S_T(25) B2S_T(30) B2S_T(32)
#undef B2S_T
    default: break;
  }
  int block = (int)(5632 / t) / 32 * 32;                   // t * block * 8 B + tw <= 48 KB
  if (block > 128) block = 128;
  if (block < 32) block = 32;
  const size_t smem = ((size_t)t * block + t) * sizeof(cfloat);
  const long long blocks = (n_pix + block - 1) / block;
  temporal_kernel<<<(unsigned)blocks, block, smem, (cudaStream_t)stream>>>((const cfloat*)in, (const cfloat*)mean_in, (cfloat*)out,
                                                                           (cfloat*)mean_out, t, hw, n_pix, xf, post);
  return check_launch("temporal_kernel");
}

extern "C" int b2s_temporal_pre(const float* image, float* x, float* mean, int b, int t, int64_t hw, int xf, void* stream) {
  if (!image || !x || !mean) return fail(B2S_EINVAL, "b2s_temporal_pre: null pointer");
  return launch_temporal(image, nullptr, x, mean, b, t, hw, xf, 0, stream);
}

extern "C" int b2s_temporal_post(const float* x, const float* mean, float* out, int b, int t, int64_t hw, int xf, void* stream) {
  if (!x || !mean || !out) return fail(B2S_EINVAL, "b2s_temporal_post: null pointer");
  return launch_temporal(x, mean, out, nullptr, b, t, hw, xf, 1, stream);
}

extern "C" int b2s_dot(const float* a, const float* b, float* out, int64_t n, float* scratch, void* stream) {
  if (!a || !b || !out || !scratch) return fail(B2S_EINVAL, "b2s_dot: null pointer");
  const unsigned g = grid_for(n, NT, 1024);
  dot_partial_kernel<<<g, NT, 0, (cudaStream_t)stream>>>(a, b, scratch, n);
  dot_final_kernel<<<1, NT, 0, (cudaStream_t)stream>>>(scratch, out, (int)g);
  return check_launch("dot kernels", 2);
}

extern "C" int b2s_axpy_ratio(float* y, const float* x, const float* num, const float* den, float sign, int64_t n, void* stream) {
  if (!y || !x || !num || !den) return fail(B2S_EINVAL, "b2s_axpy_ratio: null pointer");
  if (n == 0) return B2S_OK;
  axpy_ratio_kernel<<<grid_for(n), NT, 0, (cudaStream_t)stream>>>(y, x, num, den, sign, n);
  return check_launch("axpy_ratio_kernel");
}

extern "C" int b2s_xpay_ratio(float* p, const float* r, const float* num, const float* den, int64_t n, void* stream) {
  if (!p || !r || !num || !den) return fail(B2S_EINVAL, "b2s_xpay_ratio: null pointer");
  if (n == 0) return B2S_OK;
  xpay_ratio_kernel<<<grid_for(n), NT, 0, (cudaStream_t)stream>>>(p, r, num, den, n);
  return check_launch("xpay_ratio_kernel");
}

extern "C" int b2s_axpby(const float* a, const float* b, const float* v, float scale, float* out, int64_t n, void* stream) {
  if (!a || !b || !out) return fail(B2S_EINVAL, "b2s_axpby: null pointer");
  if (n == 0) return B2S_OK;
  axpby_kernel<<<grid_for(n), NT, 0, (cudaStream_t)stream>>>(a, b, v, scale, out, n);
  return check_launch("axpby_kernel");
}

// Fused CG iteration (after d = H p with the <p, d> partials of b2s_normal_op_dot): two launches instead of six
extern "C" int b2s_cg_blocks(int64_t n) { return (int)grid_for(n, NT, 1024); }

extern "C" int b2s_cg_update(const float* p, const float* d, float* x, float* r, const float* pd_partials, int n_pd, const float* rs_old,
                             float* rr_partials, int64_t n, void* stream) {
  if (!p || !d || !x || !r || !pd_partials || !rs_old || !rr_partials || n_pd <= 0) return fail(B2S_EINVAL, "b2s_cg_update: bad argument");
  if (n == 0) return B2S_OK;
  cg_update_kernel<<<grid_for(n, NT, 1024), NT, 0, (cudaStream_t)stream>>>(p, d, x, r, pd_partials, n_pd, rs_old, rr_partials, n);
  return check_launch("cg_update_kernel");
}

extern "C" int b2s_cg_direction(float* p, const float* r, const float* rr_partials, const float* rs_old, float* rs_new, int64_t n, void* stream) {
  if (!p || !r || !rr_partials || !rs_old || !rs_new || rs_old == rs_new) return fail(B2S_EINVAL, "b2s_cg_direction: bad argument");
  if (n == 0) return B2S_OK;
  cg_direction_kernel<<<grid_for(n, NT, 1024), NT, 0, (cudaStream_t)stream>>>(p, r, rr_partials, (int)grid_for(n, NT, 1024), rs_old, rs_new, n);
  return check_launch("cg_direction_kernel");
}
