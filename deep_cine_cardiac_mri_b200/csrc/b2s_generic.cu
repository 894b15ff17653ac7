// Any-size centred FFTs (shared-memory Stockham, mixed radix 2/3/4/5 + direct
// prime stages): the functional-API fallback for shapes without a fused plan
// (utils/fftc.py:59-110 on arbitrary h x w) and the temporal fft1c / ifft1c over
// t = 15..30 frames (utils/fftc.py:5-56; xpdnet.py:466,500 shift order).
//
// One CTA transforms LPB lines held in shared memory.  A "line" l is addressed
// as base = (l / inner) * outer_stride + (l % inner) * inner_stride with element
// stride `es` (all in complex elements), which covers rows, columns and the
// (outer, n, inner) temporal layout.  The centring rolls are folded into the
// global load / store indices; the inverse runs the forward stages on re/im
// swapped data.
#include "b2s_common.cuh"
#include "codelets.cuh"
#include "fft2_core.cuh"

using namespace b2s;

namespace {

constexpr int MAX_FAC = 24;

struct LineParams {
  int n, nfac, fac[MAX_FAC];
  int shift_in, shift_out, inverse, lpb, l_fastest;
  float scale;
  long long n_lines, inner, outer_stride, inner_stride, es;
};

template <int P>
__device__ __forceinline__ void stage_small(const cfloat* __restrict__ src, cfloat* __restrict__ dst,
                                            const cfloat* __restrict__ tw, int n, int Ns, int j) {
  const int m = n / P, k = j % Ns, step = n / (Ns * P);
  float re[P], im[P];
#pragma unroll
  for (int t = 0; t < P; ++t) {
    const cfloat v = src[j + t * m];
    if (t == 0) { re[t] = v.x; im[t] = v.y; }
    else {
      const cfloat w = tw[(t * k * step) % n];
      re[t] = v.x * w.x - v.y * w.y; im[t] = v.x * w.y + v.y * w.x;
    }
  }
  Dft<P>::run(re, im);
  const int j0 = (j - k) * P + k;
#pragma unroll
  for (int u = 0; u < P; ++u) dst[j0 + u * Ns] = make_c(re[u], im[u]);
}

__device__ __forceinline__ void stage_generic(const cfloat* __restrict__ src, cfloat* __restrict__ dst,
                                              const cfloat* __restrict__ tw, int n, int Ns, int p, int j) {
  const int m = n / p, k = j % Ns;
  const long long step = n / (Ns * p);
  const int j0 = (j - k) * p + k;
  for (int u = 0; u < p; ++u) {
    float ar = 0.f, ai = 0.f;
    for (int t = 0; t < p; ++t) {
      const cfloat v = src[j + t * m];
      const cfloat w = tw[(int)(((long long)t * k * step + (long long)t * u * m) % n)];
      ar += v.x * w.x - v.y * w.y; ai += v.x * w.y + v.y * w.x;
    }
    dst[j0 + u * Ns] = make_c(ar, ai);
  }
}

__global__ void __launch_bounds__(256) line_fft_kernel(const cfloat* in, cfloat* out,
                                                       const LineParams p) {
  extern __shared__ __align__(16) unsigned char raw[];
  cfloat* buf0 = reinterpret_cast<cfloat*>(raw);
  cfloat* buf1 = buf0 + (size_t)p.lpb * p.n;
  cfloat* tw = buf1 + (size_t)p.lpb * p.n;
  const int n = p.n, tid = threadIdx.x, nt = blockDim.x;
  const long long line0 = (long long)blockIdx.x * p.lpb;
  const int lines = (int)min((long long)p.lpb, p.n_lines - line0);

  for (int i = tid; i < n; i += nt) tw[i] = twiddle(i, n);

  // load (ifftshift folded in: position e of the rolled sequence is x[(e - shift_in) mod n])
  for (int idx = tid; idx < lines * n; idx += nt) {
    int l, e;
    if (p.l_fastest) { l = idx % lines; e = idx / lines; } else { e = idx % n; l = idx / n; }
    const long long line = line0 + l;
    const long long base = (line / p.inner) * p.outer_stride + (line % p.inner) * p.inner_stride;
    int srcpos = e - p.shift_in; if (srcpos < 0) srcpos += n;
    const cfloat v = in[base + (long long)srcpos * p.es];
    buf0[l * n + e] = p.inverse ? make_c(v.y, v.x) : v;
  }
  __syncthreads();

  cfloat* src = buf0; cfloat* dst = buf1;
  int Ns = 1;
  for (int s = 0; s < p.nfac; ++s) {
    const int radix = p.fac[s], m = n / radix;
    for (int idx = tid; idx < lines * m; idx += nt) {
      const int l = idx / m, j = idx - l * m;
      const cfloat* a = src + l * n; cfloat* b = dst + l * n;
      switch (radix) {
        case 2: stage_small<2>(a, b, tw, n, Ns, j); break;
        case 3: stage_small<3>(a, b, tw, n, Ns, j); break;
        case 4: stage_small<4>(a, b, tw, n, Ns, j); break;
        case 5: stage_small<5>(a, b, tw, n, Ns, j); break;
        default: stage_generic(a, b, tw, n, Ns, radix, j); break;
      }
    }
    __syncthreads();
    cfloat* tmp = src; src = dst; dst = tmp;
    Ns *= radix;
  }

  // store (fftshift folded in: X[k'] lands at (k' + shift_out) mod n)
  for (int idx = tid; idx < lines * n; idx += nt) {
    int l, e;
    if (p.l_fastest) { l = idx % lines; e = idx / lines; } else { e = idx % n; l = idx / n; }
    const long long line = line0 + l;
    const long long base = (line / p.inner) * p.outer_stride + (line % p.inner) * p.inner_stride;
    int kp = e - p.shift_out; if (kp < 0) kp += n;              // source bin for destination e
    const cfloat v = src[l * n + kp];
    out[base + (long long)e * p.es] = p.inverse ? make_c(v.y * p.scale, v.x * p.scale)
                                                : make_c(v.x * p.scale, v.y * p.scale);
  }
}

int factorize(int n, int* fac) {
  int nf = 0;
  while (n % 4 == 0) { fac[nf++] = 4; n /= 4; }
  while (n % 2 == 0) { fac[nf++] = 2; n /= 2; }
  while (n % 3 == 0) { fac[nf++] = 3; n /= 3; }
  while (n % 5 == 0) { fac[nf++] = 5; n /= 5; }
  for (int p = 7; (long long)p * p <= n; p += 2)
    while (n % p == 0) { fac[nf++] = p; n /= p; }
  if (n > 1) fac[nf++] = n;
  return nf;
}

int launch_lines(const float* in, float* out, int n, long long n_lines, long long inner,
                 long long outer_stride, long long inner_stride, long long es, int inverse, float scale,
                 int shift_in, int shift_out, cudaStream_t st) {
  if (n_lines <= 0) return B2S_OK;
  if (n < 1 || n > 2048) return fail(B2S_EUNSUPPORTED, "FFT length must be in [1, 2048]");
  LineParams p;
  p.n = n; p.nfac = (n == 1) ? 0 : factorize(n, p.fac);
  p.shift_in = ((shift_in % n) + n) % n; p.shift_out = ((shift_out % n) + n) % n;
  p.inverse = inverse; p.scale = scale;
  int lpb = 2048 / n; if (lpb < 1) lpb = 1; if (lpb > 16) lpb = 16;
  p.lpb = lpb; p.l_fastest = (es != 1);
  p.n_lines = n_lines; p.inner = inner; p.outer_stride = outer_stride; p.inner_stride = inner_stride; p.es = es;
  const long long blocks = (n_lines + lpb - 1) / lpb;
  if (blocks > 0x7fffffffLL) return fail(B2S_EUNSUPPORTED, "too many lines for one launch");
  const size_t smem = ((size_t)2 * lpb * n + n) * sizeof(cfloat);
  line_fft_kernel<<<(unsigned)blocks, 256, smem, st>>>((const cfloat*)in, (cfloat*)out, p);
  return check_launch("line_fft_kernel");
}

}  // namespace

namespace b2s {

int generic_fft2(const float* in, float* out, int64_t n_images, int h, int w, int inverse, float scale,
                 cudaStream_t st) {
  // pass 1: rows (along w); pass 2: columns (along h), in place on `out`
  int rc = launch_lines(in, out, w, n_images * h, 1, w, 0, 1, inverse, 1.f, (w + 1) / 2, w / 2, st);
  if (rc) return rc;
  return launch_lines(out, out, h, n_images * w, w, (long long)h * w, 1, w, inverse, scale, (h + 1) / 2, h / 2, st);
}

}  // namespace b2s

extern "C" int b2s_fft1c(const float* in, float* out, int64_t outer, int n, int64_t inner, int inverse,
                         int norm, int shift_in, int shift_out, void* stream) {
  if (outer < 0 || inner < 0 || n < 1 || bad_norm(norm)) return fail(B2S_EINVAL, "b2s_fft1c: bad argument");
  if (outer * inner == 0) return B2S_OK;
  if (!in || !out) return fail(B2S_EINVAL, "b2s_fft1c: null pointer");
  float scale = 1.f;
  if (norm == B2S_NORM_ORTHO) scale = (float)(1.0 / sqrt((double)n));
  else if ((norm == B2S_NORM_BACKWARD && inverse) || (norm == B2S_NORM_FORWARD && !inverse)) scale = 1.f / (float)n;
  return launch_lines(in, out, n, outer * inner, inner, (long long)n * inner, 1, inner, inverse, scale,
                      shift_in, shift_out, (cudaStream_t)stream);
}
