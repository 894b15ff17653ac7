// Regulariser-side layout glue of the xf / yf planes.
//
// Between the temporal head and tail of xfyf_transform the reference turns the image series x (b,t,h,w,2) into the two
// plane stacks its 2-D U-Nets run on and back (models/varnet.py:215-232), and NormUnet.forward wraps each U-Net call in
// complex_to_chan_dim / norm / pad ... unpad / unnorm / chan_complex_to_last_dim (models/denoisers/norm_unet.py:48-114):
// two clones, four permute+view copies, a reshape copy, mean, std, normalise, F.pad per plane stack on the way in, the
// mirror image on the way out and the 0.5 (xf + yf) average - about 25 I-sized eager kernels per cascade.  Here:
//
//   b2s_planes_stats   group statistics of NormUnet.norm for both stacks: mean and unbiased std of the real and of the
//                      imaginary parts of every x-f plane (b, y) over (t, x) and of every y-f plane (b, x) over (t, y),
//                      accumulated in double, reduced in a fixed order (bit-reproducible)
//   b2s_planes_pack    x -> xf (b*h, 2, wp, tp) and yf (b*w, 2, hp, tp): the U-Nets' NCHW inputs, normalised
//                      ((x - mean) / std, NormUnet.norm) and zero-padded to multiples of 16 (NormUnet.pad).  The x-f
//                      kernel holds a whole (b, y) plane in shared memory, so it computes that plane's statistics,
//                      writes the plane, and leaves per-row partial sums of every column for the y-f statistics - x
//                      is read twice in total.  Without statistics (CineNet's plain Unet, cinenet.py:193-196) it is
//                      the bare permutation
//   b2s_planes_unpack  U-Net outputs -> 0.5 * (unnorm(unpad(xf)) + unnorm(unpad(yf))) as (b,t,h,w,2)
//
// All transposes go through shared-memory tiles so that both the global reads and the global writes are contiguous
// runs (>= 64 bytes).  The tensors are I-sized (19 MB at b4 t15 200x200) and L2-resident between the kernels.
#include "b2s_common.cuh"

using namespace b2s;

namespace {

constexpr int NT = 256;
constexpr int U = 8;        // independent global loads in flight per thread in the tile fills

struct PlaneDims {
  int B, T, H, W;          // x (B,T,H,W,2)
  int HP, WP, TP;          // padded plane sizes: xf (B*H, 2, WP, TP), yf (B*W, 2, HP, TP)
  int ph0, pw0, pt0;       // leading pads
};

// ----------------------------------------------------------------------------------------------------------------- //
// statistics.  stats layout: [plane][ch][2] = {mean, std} (float)
// ----------------------------------------------------------------------------------------------------------------- //
__device__ __forceinline__ void finish_stats(double s, double ss, double n, float* out) {
  const double mean = s / n;
  double var = (ss - n * mean * mean) / (n - 1.0);          // unbiased, torch.std default (norm_unet.py:66)
  if (var < 0.0) var = 0.0;
  out[0] = (float)mean;
  out[1] = (float)sqrt(var);
}

// x-f planes: one CTA per (b, y); thread parity = channel (re / im)
__global__ void __launch_bounds__(NT) stats_rows_kernel(const float* __restrict__ x, float* __restrict__ stats, PlaneDims d) {
  __shared__ double red[2][NT];
  const int by = blockIdx.x, b = by / d.H, y = by - b * d.H;
  const int row = 2 * d.W;
  double s = 0.0, ss = 0.0;
  for (int t = 0; t < d.T; ++t) {
    const float* p = x + (((size_t)b * d.T + t) * d.H + y) * row;
    for (int i = threadIdx.x; i < row; i += NT) { const double v = p[i]; s += v; ss += v * v; }   // NT even: parity fixed
  }
  red[0][threadIdx.x] = s; red[1][threadIdx.x] = ss;
  __syncthreads();
  if (threadIdx.x < 2) {
    double a = 0.0, c = 0.0;
    for (int i = threadIdx.x; i < NT; i += 2) { a += red[0][i]; c += red[1][i]; }
    finish_stats(a, c, (double)d.T * d.W, stats + ((size_t)by * 2 + threadIdx.x) * 2);
  }
}

// y-f planes: one CTA per (b, 8-column tile): thread = (y slice of 16, column, channel)
__global__ void __launch_bounds__(NT) stats_cols_kernel(const float* __restrict__ x, float* __restrict__ stats, PlaneDims d) {
  __shared__ double red[2][NT];
  const int tiles = (d.W + 7) / 8;
  const int b = blockIdx.x / tiles, x0 = (blockIdx.x - b * tiles) * 8;
  const int e = threadIdx.x & 15, ys = threadIdx.x >> 4;     // e = column * 2 + channel
  const int col = x0 + (e >> 1);
  double s = 0.0, ss = 0.0;
  if (col < d.W)
    for (int t = 0; t < d.T; ++t)
      for (int y = ys; y < d.H; y += 16) {
        const double v = x[((((size_t)b * d.T + t) * d.H + y) * d.W + col) * 2 + (e & 1)];
        s += v; ss += v * v;
      }
  red[0][threadIdx.x] = s; red[1][threadIdx.x] = ss;
  __syncthreads();
  if (threadIdx.x < 16 && col < d.W) {
    double a = 0.0, c = 0.0;
    for (int i = 0; i < 16; ++i) { a += red[0][i * 16 + e]; c += red[1][i * 16 + e]; }
    finish_stats(a, c, (double)d.T * d.H, stats + (((size_t)b * d.W + col) * 2 + (e & 1)) * 2);
  }
}

// ----------------------------------------------------------------------------------------------------------------- //
// pack
// ----------------------------------------------------------------------------------------------------------------- //
// x-f stack: one CTA per (b, y).  tile[t][2 W (+1)] <- T contiguous 8W-byte rows; out plane (2, WP, TP) contiguous.
// With NORM: the plane's own statistics (thread parity = channel) and, for the y-f statistics, this row's partial sums
// over t of every (column, channel): part[(b, y)][2 W] = {sum, sum of squares} (double).
template <bool NORM>
__global__ void __launch_bounds__(NT) pack_xf_kernel(const float* __restrict__ x, float* __restrict__ stats, double2* __restrict__ part,
                                                     float* __restrict__ out, PlaneDims d) {
  extern __shared__ float tile[];
  __shared__ double red[2][NT];
  __shared__ float ms[4];
  const int by = blockIdx.x, b = by / d.H, y = by - b * d.H;
  const int row = 2 * d.W, pitch = row + 1;
  double s = 0.0, ss = 0.0;
  {
    // all T rows of this plane as one index space, U loads in flight per thread (row is even and NT is even: the
    // parity of a thread's elements - its channel - is fixed)
    const int n = d.T * row;
    const float* p0 = x + ((size_t)b * d.T * d.H + y) * row;
    const size_t tstride = (size_t)d.H * row;
    for (int base = threadIdx.x; base < n; base += NT * U) {
      float v[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int idx = base + u * NT, t = idx / row, i = idx - t * row;
        v[u] = idx < n ? p0[t * tstride + i] : 0.f;
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int idx = base + u * NT, t = idx / row, i = idx - t * row;
        if (idx < n) { tile[t * pitch + i] = v[u]; if (NORM) { s += (double)v[u]; ss += (double)v[u] * (double)v[u]; } }
      }
    }
  }
  if (NORM) { red[0][threadIdx.x] = s; red[1][threadIdx.x] = ss; }
  __syncthreads();
  if (NORM) {
    if (threadIdx.x < 2) {
      double a = 0.0, c = 0.0;
      for (int i = threadIdx.x; i < NT; i += 2) { a += red[0][i]; c += red[1][i]; }
      float* st = stats + ((size_t)by * 2 + threadIdx.x) * 2;
      finish_stats(a, c, (double)d.T * d.W, st);
      ms[threadIdx.x * 2] = st[0]; ms[threadIdx.x * 2 + 1] = st[1];
    }
    for (int i = threadIdx.x; i < row; i += NT) {
      double a = 0.0, c = 0.0;
      for (int t = 0; t < d.T; ++t) { const double v = tile[t * pitch + i]; a += v; c += v * v; }
      part[(size_t)by * row + i] = make_double2(a, c);
    }
    __syncthreads();
  }
  const int plane = d.WP * d.TP;
  float* o = out + (size_t)by * 2 * plane;
  // a thread keeps its plane column tp and walks down the (channel, xp) rows: no divisions in the loop
  const int tp = threadIdx.x % d.TP, r0 = threadIdx.x / d.TP, rstep = NT / d.TP, t = tp - d.pt0;
  const bool tin = t >= 0 && t < d.T && r0 < rstep;          // (NT % TP != 0: the last partial group of threads idles)
  if (rstep > 0) {
    for (int ch = 0; ch < 2; ++ch) {
      const float mean = NORM ? ms[ch * 2] : 0.f, sd = NORM ? ms[ch * 2 + 1] : 1.f;
      for (int xp = r0; xp < d.WP; xp += rstep) {
        if (r0 >= rstep) break;
        const int xx = xp - d.pw0;
        float v = 0.f;
        if (tin && xx >= 0 && xx < d.W) { v = tile[t * pitch + 2 * xx + ch]; if (NORM) v = (v - mean) / sd; }
        o[(size_t)ch * plane + xp * d.TP + tp] = v;
      }
    }
  }
}

// y-f statistics from the per-row partial sums: one CTA per (b, 16 (column, channel) entries), 16 y slices per entry,
// fixed summation order
__global__ void __launch_bounds__(NT) col_stats_kernel(const double2* __restrict__ part, float* __restrict__ stats, PlaneDims d) {
  __shared__ double red[2][NT];
  const int row = 2 * d.W, groups = (row + 15) / 16;
  const int b = blockIdx.x / groups, e = (blockIdx.x - b * groups) * 16 + (threadIdx.x & 15), ys = threadIdx.x >> 4;
  double a = 0.0, c = 0.0;
  if (e < row)
    for (int y = ys; y < d.H; y += 16) { const double2 v = part[((size_t)b * d.H + y) * row + e]; a += v.x; c += v.y; }
  red[0][threadIdx.x] = a; red[1][threadIdx.x] = c;
  __syncthreads();
  if (threadIdx.x < 16 && e < row) {
    a = 0.0; c = 0.0;
    for (int k = 0; k < 16; ++k) { a += red[0][k * 16 + threadIdx.x]; c += red[1][k * 16 + threadIdx.x]; }
    finish_stats(a, c, (double)d.T * d.H, stats + ((size_t)b * row + e) * 2);      // [(b, x)][ch][2] == [b][e][2]
  }
}

// y-f stack: one CTA per (b, 8-column tile, chunk of YC padded rows).  tile[t][y][16 (+1)] <- 64-byte runs;
// out: for each (column, channel) the chunk's (yp, tp) block is contiguous
__global__ void __launch_bounds__(NT) pack_yf_kernel(const float* __restrict__ x, const float* __restrict__ stats, float* __restrict__ out, PlaneDims d, int YC) {
  extern __shared__ float tile[];
  const int tiles = (d.W + 7) / 8, chunks = (d.HP + YC - 1) / YC;
  int id = blockIdx.x;
  const int ck = id % chunks; id /= chunks;
  const int x0 = (id % tiles) * 8, b = id / tiles;
  const int yp0 = ck * YC, nyp = min(YC, d.HP - yp0);
  const int e = threadIdx.x & 15, seg = threadIdx.x >> 4;
  const int col = x0 + (e >> 1);
  for (int t = 0; t < d.T; ++t) {
    const float* pt = x + (((size_t)b * d.T + t) * d.H) * d.W * 2 + (size_t)col * 2 + (e & 1);
    for (int yl0 = seg; yl0 < nyp; yl0 += (NT / 16) * U) {
      float v[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int yl = yl0 + u * (NT / 16), y = yp0 + yl - d.ph0;
        v[u] = (yl < nyp && y >= 0 && y < d.H && col < d.W) ? pt[(size_t)y * d.W * 2] : 0.f;
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int yl = yl0 + u * (NT / 16);
        if (yl < nyp) tile[(t * nyp + yl) * 17 + e] = v[u];
      }
    }
  }
  __syncthreads();
  const int plane = d.HP * d.TP;
  // a thread keeps its plane column tp and walks down the rows of one (column, channel) block after the other
  const int tp = threadIdx.x % d.TP, r0 = threadIdx.x / d.TP, rstep = NT / d.TP, t = tp - d.pt0;
  const bool tin = t >= 0 && t < d.T;
  if (rstep > 0 && r0 < rstep) {
    for (int ee = 0; ee < 16; ++ee) {
      const int c2 = x0 + (ee >> 1), ch = ee & 1;
      if (c2 >= d.W) break;
      float mean = 0.f, sd = 1.f;
      if (stats) { const float* st = stats + (((size_t)b * d.W + c2) * 2 + ch) * 2; mean = st[0]; sd = st[1]; }
      float* o = out + ((size_t)(b * d.W + c2) * 2 + ch) * plane + (size_t)yp0 * d.TP + tp;
      for (int yl = r0; yl < nyp; yl += rstep) {
        const int y = yp0 + yl - d.ph0;
        float v = 0.f;
        if (tin && y >= 0 && y < d.H) { v = tile[(t * nyp + yl) * 17 + ee]; if (stats) v = (v - mean) / sd; }
        o[(size_t)yl * d.TP] = v;
      }
    }
  }
}

// ----------------------------------------------------------------------------------------------------------------- //
// unpack: one CTA per (b, y).  A[ch][xp][TP (+1)] <- the (b, y) x-f plane (contiguous); Bf[x][ch][TP (+1)] <- row y + ph0
// of every (b, x) y-f plane (TP-float runs); out rows (b, t, y, :, :) contiguous
// ----------------------------------------------------------------------------------------------------------------- //
__global__ void __launch_bounds__(NT) unpack_kernel(const float* __restrict__ uxf, const float* __restrict__ uyf, const float* __restrict__ sxf,
                                                    const float* __restrict__ syf, float* __restrict__ out, PlaneDims d) {
  extern __shared__ float sm[];
  const int by = blockIdx.x, b = by / d.H, y = by - b * d.H;
  const int tpp = d.TP + 1;
  float* A = sm;                                  // 2 * WP * tpp
  float* Bf = A + 2 * d.WP * tpp;                 // 2 * W * tpp
  float* St = Bf + 2 * d.W * tpp;                 // y-f statistics of this b: [x][ch][2]
  const int planex = d.WP * d.TP, planey = d.HP * d.TP;
  const float* px = uxf + (size_t)by * 2 * planex;
  {
    // a thread keeps its plane column tp: rows of the x-f plane (2 WP of them), then the TP-float runs of the y-f planes
    // (2 W of them), U loads in flight, no divisions in the loops
    const int tp = threadIdx.x % d.TP, r0 = threadIdx.x / d.TP, rstep = NT / d.TP;
    const int na = 2 * d.WP, nb = 2 * d.W;
    const float* py = uyf + (size_t)(b * d.W) * 2 * planey + (size_t)(y + d.ph0) * d.TP + tp;
    if (r0 < rstep) {
      for (int base = r0; base < na + nb; base += rstep * U) {
        float v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int r = base + u * rstep;
          if (r < na) v[u] = px[r * d.TP + tp];
          else if (r < na + nb) v[u] = py[(size_t)(r - na) * planey];                 // r - na = x * 2 + ch
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int r = base + u * rstep;
          if (r < na) A[r * tpp + tp] = v[u];
          else if (r < na + nb) Bf[(r - na) * tpp + tp] = v[u];
        }
      }
    }
  }
  if (syf) for (int i = threadIdx.x; i < 4 * d.W; i += NT) St[i] = syf[(size_t)b * d.W * 4 + i];
  float mx[2] = {0.f, 0.f}, sx[2] = {1.f, 1.f};
  if (sxf) for (int ch = 0; ch < 2; ++ch) { mx[ch] = sxf[((size_t)by * 2 + ch) * 2]; sx[ch] = sxf[((size_t)by * 2 + ch) * 2 + 1]; }
  __syncthreads();
  const int row = 2 * d.W;
  for (int t = 0; t < d.T; ++t) {
    float* o = out + (((size_t)b * d.T + t) * d.H + y) * row;
    for (int i = threadIdx.x; i < row; i += NT) {
      const int xx = i >> 1, ch = i & 1;
      float a = A[(ch * d.WP + xx + d.pw0) * tpp + t + d.pt0];
      float c = Bf[i * tpp + t + d.pt0];
      if (sxf) a = a * sx[ch] + mx[ch];           // NormUnet.unnorm (norm_unet.py:70-73)
      if (syf) c = c * St[i * 2 + 1] + St[i * 2];
      o[i] = 0.5f * (a + c);                      // varnet.py:232
    }
  }
}

int check_dims(const char* what, int b, int t, int h, int w, int hp, int wp, int tp, int ph0, int pw0, int pt0) {
  if (b < 0 || t <= 0 || h <= 0 || w <= 0 || hp < h || wp < w || tp < t || tp > NT || ph0 < 0 || pw0 < 0 || pt0 < 0 || ph0 + h > hp || pw0 + w > wp || pt0 + t > tp)
    return fail(B2S_EINVAL, what);
  return B2S_OK;
}

}  // namespace

extern "C" int b2s_planes_stats(const float* x, float* stats_xf, float* stats_yf, int b, int t, int h, int w, void* stream) {
  if (!x || !stats_xf || !stats_yf) return fail(B2S_EINVAL, "b2s_planes_stats: null pointer");
  if (b < 0 || t <= 0 || h <= 0 || w <= 0 || (long long)t * w < 2 || (long long)t * h < 2) return fail(B2S_EINVAL, "b2s_planes_stats: bad shape");
  if (b == 0) return B2S_OK;
  PlaneDims d{b, t, h, w, h, w, t, 0, 0, 0};
  stats_rows_kernel<<<(unsigned)(b * h), NT, 0, (cudaStream_t)stream>>>(x, stats_xf, d);
  stats_cols_kernel<<<(unsigned)(b * ((w + 7) / 8)), NT, 0, (cudaStream_t)stream>>>(x, stats_yf, d);
  return check_launch("planes_stats", 2);
}

extern "C" size_t b2s_planes_scratch_bytes(int b, int t, int h, int w) {
  (void)t;
  return (size_t)(b > 0 ? b : 0) * (size_t)h * (size_t)w * 2 * sizeof(double2);
}

extern "C" int b2s_planes_pack(const float* x, float* stats_xf, float* stats_yf, float* xf, float* yf, int b, int t, int h, int w,
                               int hp, int wp, int tp, int ph0, int pw0, int pt0, void* scratch, size_t scratch_bytes, void* stream) {
  if (!x || !xf || !yf || ((stats_xf == nullptr) != (stats_yf == nullptr))) return fail(B2S_EINVAL, "b2s_planes_pack: bad pointer");
  if (int rc = check_dims("b2s_planes_pack: bad shape", b, t, h, w, hp, wp, tp, ph0, pw0, pt0)) return rc;
  const bool norm = stats_xf != nullptr;
  if (norm && ((long long)t * w < 2 || (long long)t * h < 2)) return fail(B2S_EINVAL, "b2s_planes_pack: planes of fewer than 2 samples have no std");
  if (norm && (!scratch || scratch_bytes < b2s_planes_scratch_bytes(b, t, h, w) || ((uintptr_t)scratch & 15)))
    return fail(B2S_EINVAL, "b2s_planes_pack: scratch too small or not 16-byte aligned (b2s_planes_scratch_bytes)");
  if (b == 0) return B2S_OK;
  PlaneDims d{b, t, h, w, hp, wp, tp, ph0, pw0, pt0};
  const cudaStream_t st = (cudaStream_t)stream;
  const size_t smx = (size_t)t * (2 * w + 1) * 4;
  int yc = (int)(48 * 1024 / ((size_t)t * 17 * 4));
  if (yc < 1) yc = 1;
  if (yc > hp) yc = hp;
  const size_t smy = (size_t)t * yc * 17 * 4;
  if (smx > 200 * 1024) return fail(B2S_EUNSUPPORTED, "b2s_planes_pack: t * w too large for one shared-memory tile");
  B2S_CUDA(cudaFuncSetAttribute(pack_xf_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smx));
  B2S_CUDA(cudaFuncSetAttribute(pack_xf_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smx));
  B2S_CUDA(cudaFuncSetAttribute(pack_yf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smy));
  int launches = 2;
  if (norm) {
    pack_xf_kernel<true><<<(unsigned)(b * h), NT, smx, st>>>(x, stats_xf, (double2*)scratch, xf, d);
    col_stats_kernel<<<(unsigned)(b * ((2 * w + 15) / 16)), NT, 0, st>>>((const double2*)scratch, stats_yf, d);
    ++launches;
  } else {
    pack_xf_kernel<false><<<(unsigned)(b * h), NT, smx, st>>>(x, nullptr, nullptr, xf, d);
  }
  const int chunks = (hp + yc - 1) / yc;
  pack_yf_kernel<<<(unsigned)(b * ((w + 7) / 8) * chunks), NT, smy, st>>>(x, stats_yf, yf, d, yc);
  return check_launch("planes_pack", launches);
}

extern "C" int b2s_planes_unpack(const float* uxf, const float* uyf, const float* stats_xf, const float* stats_yf, float* out, int b, int t, int h,
                                 int w, int hp, int wp, int tp, int ph0, int pw0, int pt0, void* stream) {
  if (!uxf || !uyf || !out || ((stats_xf == nullptr) != (stats_yf == nullptr))) return fail(B2S_EINVAL, "b2s_planes_unpack: bad pointer");
  if (int rc = check_dims("b2s_planes_unpack: bad shape", b, t, h, w, hp, wp, tp, ph0, pw0, pt0)) return rc;
  if (b == 0) return B2S_OK;
  PlaneDims d{b, t, h, w, hp, wp, tp, ph0, pw0, pt0};
  const size_t sm = ((size_t)2 * wp * (tp + 1) + (size_t)2 * w * (tp + 1) + (size_t)4 * w) * 4;
  if (sm > 200 * 1024) return fail(B2S_EUNSUPPORTED, "b2s_planes_unpack: plane too large for one shared-memory tile");
  B2S_CUDA(cudaFuncSetAttribute(unpack_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
  unpack_kernel<<<(unsigned)(b * h), NT, sm, (cudaStream_t)stream>>>(uxf, uyf, stats_xf, stats_yf, out, d);
  return check_launch("planes_unpack");
}
