// Time-averaged SSIM (training loss and test metric) and the NMSE / PSNR error statistics on the GPU.
//
// Reference: utils/losses.py:25-58 (SSIMLoss: per frame, five 7x7 uniform "valid" convolutions of X, Y, XX,
// YY, XY, sample covariance, S = (2 ux uy + C1)(2 vxy + C2) / ((ux^2 + uy^2 + C1)(vx + vy + C2)), loss =
// mean_t (1 - mean S); data_range = Y.max() per frame through a host round trip) and utils/evaluate.py:11-42
// (nmse, psnr, ssim = skimage's structural_similarity defaults, the same formula with one data_range per
// volume).  Here: one tiled kernel per direction (window sums separable in shared memory, no intermediate
// maps in HBM), data_range read from device memory (no host sync), ordered two-stage reductions
// (bit-reproducible).  Window sums are taken on tile-shifted values x - x0, y - y0 (x0, y0 = the tile's
// first pixel): means and (co)variances are shift-invariant in exact arithmetic, and in fp32 the shift
// removes most of the cancellation in uxx - ux^2 that the reference's own fp32 convolutions suffer from.
#include "b2s_common.cuh"

namespace b2s {
namespace {

constexpr int WIN = 7, HALO = WIN - 1, TS = 32, NT = 256;
constexpr float INV_NP = 1.f / (WIN * WIN), COV_NORM = (float)(WIN * WIN) / (WIN * WIN - 1);

__device__ __forceinline__ float block_sum256(float v, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = 0.f;
  if (threadIdx.x < 32) {
    r = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
  }
  __syncthreads();
  return r;                      // valid in warp 0
}

struct Stats { float ux, uy, uxx, uyy, uxy; };   // window means of the shifted values

// S and (optionally) its partial derivatives w.r.t. the shifted window means ux', uxx', uxy'
__device__ __forceinline__ float ssim_point(const Stats& s, float px, float py, float c1, float c2, float* g1, float* g2,
                                            float* g3) {
  const float mx = px + s.ux, my = py + s.uy;                  // true means
  const float vx = COV_NORM * (s.uxx - s.ux * s.ux), vy = COV_NORM * (s.uyy - s.uy * s.uy);
  const float vxy = COV_NORM * (s.uxy - s.ux * s.uy);
  const float a1 = 2.f * mx * my + c1, a2 = 2.f * vxy + c2, b1 = mx * mx + my * my + c1, b2 = vx + vy + c2;
  const float invd = 1.f / (b1 * b2);
  const float S = a1 * a2 * invd;
  if (g1) {
    const float da1 = 2.f * my, da2 = -2.f * COV_NORM * s.uy, db1 = 2.f * mx, db2 = -2.f * COV_NORM * s.ux;
    *g1 = (da1 * a2 + a1 * da2) * invd - S * (db1 / b1 + db2 / b2);     // dS/dux'
    *g2 = -S * COV_NORM / b2;                                            // dS/duxx'
    *g3 = 2.f * COV_NORM * a1 * invd;                                    // dS/duxy'
  }
  return S;
}

// loads an RxR region (origin (y0, x0), zero outside the image) of x and y, shifted by the pivots
template <int R>
__device__ __forceinline__ void load_region(const float* __restrict__ x, const float* __restrict__ y, int h, int w, int y0,
                                            int x0, float px, float py, float (*sx)[R + 1], float (*sy)[R + 1]) {
  for (int i = threadIdx.x; i < R * R; i += NT) {
    const int r = i / R, c = i - r * R, yy = y0 + r, xx = x0 + c;
    const bool in = yy >= 0 && yy < h && xx >= 0 && xx < w;
    sx[r][c] = in ? x[(long long)yy * w + xx] - px : 0.f;
    sy[r][c] = in ? y[(long long)yy * w + xx] - py : 0.f;
  }
}

// horizontal 7-sums of the five products: rows R, output columns C = R - HALO
template <int R>
__device__ __forceinline__ void hsum5(const float (*sx)[R + 1], const float (*sy)[R + 1], float (*h5)[R][R - HALO + 1]) {
  constexpr int C = R - HALO;
  for (int i = threadIdx.x; i < R * C; i += NT) {
    const int r = i / C, c = i - r * C;
    float a = 0.f, b = 0.f, aa = 0.f, bb = 0.f, ab = 0.f;
#pragma unroll
    for (int k = 0; k < WIN; ++k) {
      const float u = sx[r][c + k], v = sy[r][c + k];
      a += u; b += v; aa = fmaf(u, u, aa); bb = fmaf(v, v, bb); ab = fmaf(u, v, ab);
    }
    h5[0][r][c] = a; h5[1][r][c] = b; h5[2][r][c] = aa; h5[3][r][c] = bb; h5[4][r][c] = ab;
  }
}

template <int R>
__device__ __forceinline__ Stats vsum5(const float (*h5)[R][R - HALO + 1], int r, int c) {
  float a = 0.f, b = 0.f, aa = 0.f, bb = 0.f, ab = 0.f;
#pragma unroll
  for (int k = 0; k < WIN; ++k) { a += h5[0][r + k][c]; b += h5[1][r + k][c]; aa += h5[2][r + k][c]; bb += h5[3][r + k][c]; ab += h5[4][r + k][c]; }
  Stats s; s.ux = a * INV_NP; s.uy = b * INV_NP; s.uxx = aa * INV_NP; s.uyy = bb * INV_NP; s.uxy = ab * INV_NP;
  return s;
}

__device__ __forceinline__ float pivot(const float* img, int h, int w, int y0, int x0) {
  const int yy = min(max(y0, 0), h - 1), xx = min(max(x0, 0), w - 1);
  return img[(long long)yy * w + xx];
}

// ---- forward: partial[img * tiles + tile] = sum of S over the tile's valid window positions
__global__ void __launch_bounds__(NT) ssim_fwd_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                                      const float* __restrict__ dr, int dr_stride, int T, int h, int w,
                                                      float k1, float k2, float* __restrict__ partial) {
  constexpr int R = TS + HALO;
  __shared__ float sx[R][R + 1], sy[R][R + 1], h5[5][R][TS + 1], red[8];
  const int img = blockIdx.z, oy0 = blockIdx.y * TS, ox0 = blockIdx.x * TS, oh = h - HALO, ow = w - HALO;
  const float* xi = x + (long long)img * h * w;
  const float* yi = y + (long long)img * h * w;
  const float range = dr[(img % T) * dr_stride];
  const float c1 = (k1 * range) * (k1 * range), c2 = (k2 * range) * (k2 * range);
  const float px = pivot(xi, h, w, oy0, ox0), py = pivot(yi, h, w, oy0, ox0);
  load_region<R>(xi, yi, h, w, oy0, ox0, px, py, sx, sy);
  __syncthreads();
  hsum5<R>(sx, sy, h5);
  __syncthreads();
  float acc = 0.f;
  for (int i = threadIdx.x; i < TS * TS; i += NT) {
    const int r = i / TS, c = i - r * TS;
    if (oy0 + r < oh && ox0 + c < ow) acc += ssim_point(vsum5<R>(h5, r, c), px, py, c1, c2, nullptr, nullptr, nullptr);
  }
  acc = block_sum256(acc, red);
  if (threadIdx.x == 0) partial[(long long)img * (gridDim.x * gridDim.y) + blockIdx.y * gridDim.x + blockIdx.x] = acc;
}

// out[t] = mean over batch and window positions of frame t (fixed summation order), out[T] = mean_t (1 - out[t])
__global__ void __launch_bounds__(NT) ssim_final_kernel(const float* __restrict__ partial, int B, int T, int tiles,
                                                        float inv_count, float* __restrict__ out) {
  __shared__ float red[8];
  float loss = 0.f;
  for (int t = 0; t < T; ++t) {
    float acc = 0.f;
    for (int i = threadIdx.x; i < B * tiles; i += NT) acc += partial[((long long)(i / tiles) * T + t) * tiles + (i % tiles)];
    acc = block_sum256(acc, red);
    if (threadIdx.x == 0) { const float m = acc * inv_count; out[t] = m; loss += 1.f - m; }
  }
  if (threadIdx.x == 0) out[T] = loss / (float)T;
}

// ---- backward: gx = gout * d loss / d x
__global__ void __launch_bounds__(NT) ssim_bwd_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                                      const float* __restrict__ dr, int dr_stride, const float* __restrict__ gout,
                                                      int T, int h, int w, float k1, float k2, float coef, float* __restrict__ gx) {
  constexpr int R = TS + 2 * HALO, M = TS + HALO;      // input region edge, window-position region edge
  extern __shared__ float dyn[];
  float (*sx)[R + 1] = reinterpret_cast<float (*)[R + 1]>(dyn);
  float (*sy)[R + 1] = sx + R;
  float (*h5)[R][M + 1] = reinterpret_cast<float (*)[R][M + 1]>(sy + R);          // 5 planes; reused below
  float (*mp)[M][M + 1] = reinterpret_cast<float (*)[M][M + 1]>(&h5[5][0][0]);     // 3 derivative maps
  float (*hm)[M][TS + 1] = reinterpret_cast<float (*)[M][TS + 1]>(&h5[0][0][0]);   // their horizontal sums (aliases h5)
  const int img = blockIdx.z, iy0 = blockIdx.y * TS, ix0 = blockIdx.x * TS, oh = h - HALO, ow = w - HALO;
  const float* xi = x + (long long)img * h * w;
  const float* yi = y + (long long)img * h * w;
  const float range = dr[(img % T) * dr_stride];
  const float c1 = (k1 * range) * (k1 * range), c2 = (k2 * range) * (k2 * range);
  const float px = pivot(xi, h, w, iy0, ix0), py = pivot(yi, h, w, iy0, ix0);
  load_region<R>(xi, yi, h, w, iy0 - HALO, ix0 - HALO, px, py, sx, sy);
  __syncthreads();
  {  // horizontal sums: R rows x M columns (window column positions ix0-6 .. ix0+31)
    for (int i = threadIdx.x; i < R * M; i += NT) {
      const int r = i / M, c = i - r * M;
      float a = 0.f, b = 0.f, aa = 0.f, bb = 0.f, ab = 0.f;
#pragma unroll
      for (int k = 0; k < WIN; ++k) {
        const float u = sx[r][c + k], v = sy[r][c + k];
        a += u; b += v; aa = fmaf(u, u, aa); bb = fmaf(v, v, bb); ab = fmaf(u, v, ab);
      }
      h5[0][r][c] = a; h5[1][r][c] = b; h5[2][r][c] = aa; h5[3][r][c] = bb; h5[4][r][c] = ab;
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < M * M; i += NT) {       // derivative maps at window positions (oy, ox)
    const int r = i / M, c = i - r * M, oy = iy0 - HALO + r, ox = ix0 - HALO + c;
    float g1 = 0.f, g2 = 0.f, g3 = 0.f;
    if (oy >= 0 && oy < oh && ox >= 0 && ox < ow) {
      float a = 0.f, b = 0.f, aa = 0.f, bb = 0.f, ab = 0.f;
#pragma unroll
      for (int k = 0; k < WIN; ++k) { a += h5[0][r + k][c]; b += h5[1][r + k][c]; aa += h5[2][r + k][c]; bb += h5[3][r + k][c]; ab += h5[4][r + k][c]; }
      Stats s; s.ux = a * INV_NP; s.uy = b * INV_NP; s.uxx = aa * INV_NP; s.uyy = bb * INV_NP; s.uxy = ab * INV_NP;
      ssim_point(s, px, py, c1, c2, &g1, &g2, &g3);
    }
    mp[0][r][c] = g1; mp[1][r][c] = g2; mp[2][r][c] = g3;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < M * TS; i += NT) {       // transposed box filter, horizontal part
    const int r = i / TS, c = i - r * TS;
    float a = 0.f, b = 0.f, d = 0.f;
#pragma unroll
    for (int k = 0; k < WIN; ++k) { a += mp[0][r][c + k]; b += mp[1][r][c + k]; d += mp[2][r][c + k]; }
    hm[0][r][c] = a; hm[1][r][c] = b; hm[2][r][c] = d;
  }
  __syncthreads();
  const float g = gout[0] * coef;
  for (int i = threadIdx.x; i < TS * TS; i += NT) {      // vertical part + chain rule
    const int r = i / TS, c = i - r * TS, yy = iy0 + r, xx = ix0 + c;
    if (yy < h && xx < w) {
      float a = 0.f, b = 0.f, d = 0.f;
#pragma unroll
      for (int k = 0; k < WIN; ++k) { a += hm[0][r + k][c]; b += hm[1][r + k][c]; d += hm[2][r + k][c]; }
      gx[(long long)img * h * w + (long long)yy * w + xx] = g * (a + 2.f * sx[r + HALO][c + HALO] * b + sy[r + HALO][c + HALO] * d);
    }
  }
}
constexpr int BWD_R = TS + 2 * HALO, BWD_M = TS + HALO;
constexpr size_t BWD_SMEM = sizeof(float) * (2 * BWD_R * (BWD_R + 1) + 5 * BWD_R * (BWD_M + 1) + 3 * BWD_M * (BWD_M + 1));

// ---- per-frame maximum over batch and pixels (data_range = Y.max() of losses.py:35)
__global__ void __launch_bounds__(NT) frame_max_kernel(const float* __restrict__ y, int B, int T, long long hw, float* __restrict__ out) {
  __shared__ float red[8];
  const int t = blockIdx.x;
  float m = -INFINITY;
  for (int b = 0; b < B; ++b) {
    const float* p = y + ((long long)b * T + t) * hw;
    for (long long i = threadIdx.x; i < hw; i += NT) m = fmaxf(m, p[i]);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) { for (int i = 1; i < NT / 32; ++i) m = fmaxf(m, red[i]); out[t] = m; }
}

// ---- error statistics: partial sums of (gt-pred)^2, gt^2 and max gt per block, then one ordered pass
__global__ void __launch_bounds__(NT) err_partial_kernel(const float* __restrict__ gt, const float* __restrict__ pred, long long n,
                                                         float* __restrict__ scratch) {
  __shared__ float red[8];
  float se = 0.f, sg = 0.f, mx = -INFINITY;
  for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < n; i += (long long)gridDim.x * NT) {
    const float g = gt[i], d = g - pred[i];
    se = fmaf(d, d, se); sg = fmaf(g, g, sg); mx = fmaxf(mx, g);
  }
  se = block_sum256(se, red);
  sg = block_sum256(sg, red);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < NT / 32; ++i) mx = fmaxf(mx, red[i]);
    scratch[3 * blockIdx.x] = se; scratch[3 * blockIdx.x + 1] = sg; scratch[3 * blockIdx.x + 2] = mx;
  }
}
__global__ void __launch_bounds__(NT) err_final_kernel(const float* __restrict__ scratch, int nblk, float n, float* __restrict__ out) {
  __shared__ float red[8];
  float se = 0.f, sg = 0.f, mx = -INFINITY;
  for (int i = threadIdx.x; i < nblk; i += NT) { se += scratch[3 * i]; sg += scratch[3 * i + 1]; mx = fmaxf(mx, scratch[3 * i + 2]); }
  se = block_sum256(se, red);
  sg = block_sum256(sg, red);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < NT / 32; ++i) mx = fmaxf(mx, red[i]);
    out[0] = se; out[1] = sg; out[2] = mx; out[3] = n;
  }
}

inline bool bad_ssim_shape(int b, int t, int h, int w) { return b < 0 || t < 0 || h < WIN || w < WIN; }
inline int tiles_of(int n) { return (n + TS - 1) / TS; }

}  // namespace
}  // namespace b2s

using namespace b2s;

extern "C" size_t b2s_ssim_scratch_floats(int b, int t, int h, int w) {
  if (bad_ssim_shape(b, t, h, w)) return 0;
  return (size_t)b * t * tiles_of(h - HALO) * tiles_of(w - HALO);
}

extern "C" int b2s_ssim_fwd(const float* x, const float* y, const float* data_range, int dr_stride, int b, int t, int h,
                            int w, int win, float k1, float k2, float* out, float* scratch, void* stream) {
  if (win != WIN) return fail(B2S_EUNSUPPORTED, "b2s_ssim_fwd: only win_size 7 is built");
  if (bad_ssim_shape(b, t, h, w) || (dr_stride != 0 && dr_stride != 1)) return fail(B2S_EINVAL, "b2s_ssim_fwd: bad argument");
  if ((long long)b * t == 0) return B2S_OK;
  if (!x || !y || !data_range || !out || !scratch) return fail(B2S_EINVAL, "b2s_ssim_fwd: null pointer");
  if ((long long)b * t > 65535) return fail(B2S_EUNSUPPORTED, "b2s_ssim_fwd: more than 65535 frames per call");
  cudaStream_t st = (cudaStream_t)stream;
  const int ty = tiles_of(h - HALO), tx = tiles_of(w - HALO);
  ssim_fwd_kernel<<<dim3(tx, ty, b * t), NT, 0, st>>>(x, y, data_range, dr_stride, t, h, w, k1, k2, scratch);
  const float inv_count = 1.f / ((float)b * (float)(h - HALO) * (float)(w - HALO));
  ssim_final_kernel<<<1, NT, 0, st>>>(scratch, b, t, tx * ty, inv_count, out);
  return check_launch("ssim forward kernels", 2);
}

extern "C" int b2s_ssim_bwd(const float* x, const float* y, const float* data_range, int dr_stride, const float* gout,
                            int b, int t, int h, int w, int win, float k1, float k2, float* gx, void* stream) {
  if (win != WIN) return fail(B2S_EUNSUPPORTED, "b2s_ssim_bwd: only win_size 7 is built");
  if (bad_ssim_shape(b, t, h, w) || (dr_stride != 0 && dr_stride != 1)) return fail(B2S_EINVAL, "b2s_ssim_bwd: bad argument");
  if ((long long)b * t == 0) return B2S_OK;
  if (!x || !y || !data_range || !gout || !gx) return fail(B2S_EINVAL, "b2s_ssim_bwd: null pointer");
  if ((long long)b * t > 65535) return fail(B2S_EUNSUPPORTED, "b2s_ssim_bwd: more than 65535 frames per call");
  cudaStream_t st = (cudaStream_t)stream;
  static std::atomic<bool> configured[64];
  int dev = 0;
  B2S_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) return fail(B2S_EUNSUPPORTED, "device index >= 64");
  if (!configured[dev].load()) {
    B2S_CUDA(cudaFuncSetAttribute(ssim_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BWD_SMEM));
    configured[dev].store(true);
  }
  // loss = 1/T sum_t (1 - 1/(B oh ow) sum S); every window weight is 1/49
  const float coef = -1.f / ((float)t * (float)b * (float)(h - HALO) * (float)(w - HALO)) * INV_NP;
  ssim_bwd_kernel<<<dim3(tiles_of(w), tiles_of(h), b * t), NT, BWD_SMEM, st>>>(x, y, data_range, dr_stride, gout, t, h, w, k1, k2, coef, gx);
  return check_launch("ssim_bwd_kernel");
}

extern "C" int b2s_frame_max(const float* y, float* out, int b, int t, int64_t hw, void* stream) {
  if (b < 0 || t < 0 || hw < 0) return fail(B2S_EINVAL, "b2s_frame_max: bad argument");
  if (t == 0) return B2S_OK;
  if (!y || !out) return fail(B2S_EINVAL, "b2s_frame_max: null pointer");
  frame_max_kernel<<<t, NT, 0, (cudaStream_t)stream>>>(y, b, t, hw, out);
  return check_launch("frame_max_kernel");
}

extern "C" int b2s_err_stats(const float* gt, const float* pred, int64_t n, float* out, float* scratch, void* stream) {
  if (n < 0) return fail(B2S_EINVAL, "b2s_err_stats: bad argument");
  if (!gt || !pred || !out || !scratch) return fail(B2S_EINVAL, "b2s_err_stats: null pointer");
  long long nb = (n + NT - 1) / NT;
  const int nblk = (int)(nb < 1 ? 1 : (nb > 1024 ? 1024 : nb));
  err_partial_kernel<<<nblk, NT, 0, (cudaStream_t)stream>>>(gt, pred, n, scratch);
  err_final_kernel<<<1, NT, 0, (cudaStream_t)stream>>>(scratch, nblk, (float)n, out);
  return check_launch("err_stats kernels", 2);
}
