/* b200sense — C ABI of the B200-native SENSE / data-consistency operators.
 *
 * Drop-in boundary for the hot path of f78bono/deep-cine-cardiac-mri.  The
 * reference has no FFI layer: its boundary is the Python functional API
 * (reconstruction/utils/__init__.py:1-25) plus the block methods listed below.
 * Each entry point names the reference code it replaces (paths relative to the
 * reference checkout).  INTEGRATION.md shows the ctypes binding.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name ends in `_host`;
 *   - complex tensors are float32 with a trailing (re, im) pair, contiguous;
 *   - k-space  (b,t,c,h,w,2) · image (b,t,h,w,2) · sens (b,c,h,w,2) ·
 *     mask (b,t,h) uint8 (the reference's (b,t,1,h,1,1) flattened) ·
 *     `v` = softplus(lambda) as ONE float in device memory (no host sync);
 *   - the caller owns and pre-allocates all buffers, scratch included;
 *   - `stream` is a cudaStream_t; calls only enqueue work (graph-capturable);
 *   - `norm`: 0 "backward" (None), 1 "ortho", 2 "forward" (torch.fft meaning);
 *   - return 0 on success, B2S_E* otherwise; b2s_last_error() gives the text.
 *     No entry point ever falls back to a CPU or library path.
 */
#ifndef B200SENSE_H
#define B200SENSE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2S_OK 0
#define B2S_EINVAL 1       /* bad argument (null pointer, bad mode)            */
#define B2S_EUNSUPPORTED 2 /* shape outside what the kernels implement         */
#define B2S_ECUDA 3        /* CUDA runtime error (text in b2s_last_error)      */

#define B2S_NORM_BACKWARD 0
#define B2S_NORM_ORTHO 1
#define B2S_NORM_FORWARD 2

/* epilogue of b2s_sens_expand */
#define B2S_EXPAND_PLAIN 0    /* k = F(S x)                      varnet.py:181-185          */
#define B2S_EXPAND_MASK 1     /* k*m + 0.0                       cinenet.py:127-129, xpdnet.py:129-131 */
#define B2S_EXPAND_DC 2       /* (1-m)k + m(k + v ref)/(1+v)     varnet.py:281-282, recurrent_varnet.py:86-89 */
#define B2S_EXPAND_RESIDUAL 3 /* k*m - ref                       xpdnet.py:295-298 + 386-403 */

/* input row weight of b2s_sens_reduce */
#define B2S_REDUCE_PLAIN 0    /* A^H k                           varnet.py:187-194          */
#define B2S_REDUCE_MASK 1     /* A^H (k*m)                       xpdnet.py:161-167          */
#define B2S_REDUCE_DCGRAD 2   /* A^H (k*(1 - v/(1+v) m))         backward of B2S_EXPAND_DC  */
/* OR into weight_mode: sum the coils in a fixed order instead of with float atomics (run-to-run
 * bit-identical results, for the reference's Trainer(deterministic=True), train_test_varnet.py:292).
 * Costs one extra K-sized round trip; needs b*t*c*h*w*8 bytes of scratch for every shape. */
#define B2S_REDUCE_DETERMINISTIC 0x10

/* State.  Every operator entry point below is a pure function of its arguments (device pointers, sizes, stream): the library
 * keeps no per-call or per-tensor state and may be called concurrently from several host threads on different streams.
 * The only mutable process-wide settings are the two launch-configuration knobs b2s_set_sm_reserve and b2s_set_fused_path
 * (atomics, read once per launch; meant to be set at start-up, they never change results, only which SMs / which kernel
 * family a launch uses), the launch counter, and the thread-local text of b2s_last_error. */
int b2s_version(void);
const char* b2s_last_error(void);
/* number of CUDA kernels this library has launched so far (host-side counter; reset != 0 zeroes it) */
unsigned long long b2s_launch_count(int reset);
/* Leave `n_sms` SMs free: the persistent fused kernels (one CTA per SM) launch that many CTAs fewer, and
 * b2s_upload_rows runs as `n_sms` 1024-thread CTAs on them - the sparse upload of the next batch then overlaps the
 * compute of the current one instead of queueing behind its statically strided grids.  0 (default) = use every SM. */
int b2s_set_sm_reserve(int n_sms);
/* Diagnostic of the strip-streamed kernels (experimental builds, `make EXPERIMENTS=1`): synchronises the current device
 * and returns 0 when no inter-CTA dependency wait ever timed out on it (always 0 in product builds). */
int b2s_debug_strip_status(void);
/* Kernel family behind the fused plan sizes (200x200, 256x256) of b2s_fft2c / b2s_sens_expand / b2s_sens_reduce - a test
 * and measurement knob, process-wide, every family computes the same operators (same parity tests):
 *   0 or -1  automatic (default): per launch, the measured cost model of csrc/b2s_fused.cu picks between
 *   2        the half-split (200x200) / quarter-split (256x256) on-chip kernels (csrc/fft2_core.cuh) and
 *   3        the packed whole-image kernel (200x200 sens_expand: one image per CTA, second half of the intermediate
 *            parked in tensor memory, two transforms per thread in packed fp32 - csrc/fft2_packed.cuh);
 *   1        strip-streamed kernels (two passes through an L2-resident ring, csrc/strip_core.cuh) - experimental builds
 *            only, B2S_EUNSUPPORTED otherwise. */
int b2s_set_fused_path(int path);
/* 1 if (h,w) runs on the fused single-pass kernels, 0 if on the generic two-pass ones */
int b2s_has_fused_plan(int h, int w);
/* bytes of scratch b2s_sens_expand / b2s_sens_reduce need for this shape (0 for fused plans) */
size_t b2s_scratch_bytes(int b, int t, int c, int h, int w);

/* fft2c / ifft2c — utils/fftc.py:59-83, 86-110.  n_images centred 2-D transforms of h x w. */
int b2s_fft2c(const float* in, float* out, int64_t n_images, int h, int w, int inverse, int norm,
              void* stream);

/* fft1c / ifft1c — utils/fftc.py:5-29, 32-56, and the XPDNet XF variant xpdnet.py:466,500.
 * Layout (outer, n, inner, 2), transform over n (the callers' permuted views
 * varnet.py:211-213,236-238 are this layout with inner = h*w).
 * shift_in / shift_out: circular roll applied before / after the transform
 * (fft1c: (n+1)/2, n/2 ; xpdnet.py:466: n/2, (n+1)/2). */
int b2s_fft1c(const float* in, float* out, int64_t outer, int n, int64_t inner, int inverse, int norm,
              int shift_in, int shift_out, void* stream);

/* A — sens_expand fused with its consumer: varnet.py:181-185 (+281-282), cinenet.py:106-110 (+129),
 * xpdnet.py:119-133, recurrent_varnet.py:65-69.  image (b,t,h,w,2), sens (b,c,h,w,2) ->
 * kspace (b,t,c,h,w,2).  ref / mask / v may be NULL when the mode does not use them. */
int b2s_sens_expand(const float* image, const float* sens, float* kspace, const float* ref,
                    const uint8_t* mask, const float* v, int mode, int b, int t, int c, int h, int w,
                    int norm, void* scratch, size_t scratch_bytes, void* stream);

/* A^H — sens_reduce: varnet.py:187-194, cinenet.py:112-119, xpdnet.py:152-167, recurrent_*.py.
 * over_frames == 0: out (b,t,h,w,2) = sum_c conj(mult[b,c]) * ifft2c(w(ky) k[b,t,c]),  mult = sens
 * over_frames == 1: out (b,c,h,w,2) = sum_t conj(mult[b,t]) * ifft2c(w(ky) k[b,t,c]),  mult = image
 *                   (gradient w.r.t. the sensitivity maps). */
int b2s_sens_reduce(const float* kspace, const float* mult, float* out, const uint8_t* mask,
                    const float* v, int weight_mode, int over_frames, int b, int t, int c, int h,
                    int w, int norm, void* scratch, size_t scratch_bytes, void* stream);

/* apply_mask - data/transforms.py:66-92 (the multiplication; the mask itself comes from the host-side mask functions,
 * data/subsample.py:75-215, restated in deep_cine_cardiac_mri_b200/masks.py): out = kspace * m + 0.0, mask (n_bt, h) uint8
 * broadcast over coils and columns. */
int b2s_apply_mask(const float* kspace, const uint8_t* mask, float* out, int64_t n_bt, int c, int h, int w,
                   void* stream);

/* Stand-alone soft data-consistency blend — varnet.py:281-282. n_bt = b*t. */
int b2s_dc_blend(const float* kspace, const float* ref, const uint8_t* mask, const float* v,
                 float* out, int64_t n_bt, int c, int h, int w, void* stream);
/* Its backward: gk = g(1 - eta m), gref = g eta m (either may be NULL),
 * gv[0] = sum g m (ref - out)/(1+v) (ordered two-stage sum, no float atomics); gv (may be NULL)
 * points to B2S_DC_BWD_GV_FLOATS floats: element 0 is the result, the rest is workspace. */
#define B2S_DC_BWD_GV_FLOATS 1032
int b2s_dc_blend_bwd(const float* g, const float* out, const float* ref, const uint8_t* mask,
                     const float* v, float* gk, float* gref, float* gv, int64_t n_bt, int c, int h,
                     int w, void* stream);

/* utils/math.py:5-25 complex_mul with broadcasting: element strides (in complex
 * elements, 0 = broadcast) over up to 6 leading dims; out is contiguous. conj_b: multiply by conj(b). */
int b2s_complex_mul(const float* a, const float* b, float* out, int ndim, const int64_t* shape,
                    const int64_t* stride_a, const int64_t* stride_b, int conj_b, void* stream);
/* utils/math.py:28-45 / 48-62 / 65-79 on n contiguous complex elements */
int b2s_complex_conj(const float* in, float* out, int64_t n, void* stream);
int b2s_complex_abs(const float* in, float* out, int64_t n, int squared, void* stream);
/* utils/coil_combine.py:5-18 / 21-34: sqrt(sum over r of x^2) on (outer, r, inner[,2]) */
int b2s_rss(const float* in, float* out, int64_t outer, int64_t r, int64_t inner, int is_complex,
            void* stream);

/* SensitivityModel pre — varnet.py:64-71 / xpdnet.py:75-82: device-side ACS window from frame 0 of
 * the mask, mean over t, rows outside the window zeroed (transforms.py:95-108).
 * kspace (b,t,c,h,w,2) -> out (b,c,h,w,2).  window_out (optional, 2 ints per batch: pad, nlf). */
int b2s_acs_mean(const float* kspace, const uint8_t* mask, float* out, int32_t* window_out, int b,
                 int t, int c, int h, int w, void* stream);
/* SensitivityModel post — varnet.py:58-59: x / rss_complex(x, coil dim) on (b,c,hw,2) */
int b2s_rss_normalize(const float* in, float* out, int b, int c, int64_t hw, void* stream);
/* backward of rss_normalize: gin = (g - x Re<g,x>_c / rss^2) / rss */
int b2s_rss_normalize_bwd(const float* g, const float* in, float* gin, int b, int c, int64_t hw,
                          void* stream);

/* xfyf_transform head/tail — varnet.py:202-213, 232-241; cinenet.py:180-191, 210-219.
 * pre:  image (b,t,hw,2) -> x = fft1c_t(image - mean_t) (xf != 0) or image - mean_t ; mean (b,hw,2)
 * post: x (b,t,hw,2), mean -> ifft1c_t(x) + mean (xf != 0) or x + mean */
int b2s_temporal_pre(const float* image, float* x, float* mean, int b, int t, int64_t hw, int xf,
                     void* stream);
int b2s_temporal_post(const float* x, const float* mean, float* out, int b, int t, int64_t hw, int xf,
                      void* stream);

/* Regulariser-side layout glue of the x-f / y-f planes — varnet.py:215-232 (xfyf_transform's permute/view pairs and
 * the 0.5 (xf + yf) average) and denoisers/norm_unet.py:48-114 (NormUnet.complex_to_chan_dim / norm / pad and
 * unpad / unnorm / chan_complex_to_last_dim around the untouched U-Net); cinenet.py:193-212 without statistics.
 * x (b,t,h,w,2).  stats_xf (b*h,2,2) / stats_yf (b*w,2,2) = {mean, unbiased std} of the real and imaginary parts of
 * every x-f plane (b,y) over (t,x) and every y-f plane (b,x) over (t,y) — NormUnet.norm's groups. */
int b2s_planes_stats(const float* x, float* stats_xf, float* stats_yf, int b, int t, int h, int w, void* stream);
/* xf (b*h,2,wp,tp), yf (b*w,2,hp,tp): the U-Nets' NCHW inputs, zero-padded: plane row r holds image column/row
 * r - pw0 / r - ph0, plane column q holds frame q - pt0.  stats_xf / stats_yf both non-NULL: OUTPUTS, the group statistics
 * are computed here (as b2s_planes_stats) and the planes are (x - mean)/std; `scratch` >= b2s_planes_scratch_bytes bytes,
 * 16-byte aligned.  Both NULL: bare permutation (cinenet.py:193-196), no scratch. */
size_t b2s_planes_scratch_bytes(int b, int t, int h, int w);
int b2s_planes_pack(const float* x, float* stats_xf, float* stats_yf, float* xf, float* yf,
                    int b, int t, int h, int w, int hp, int wp, int tp, int ph0, int pw0, int pt0,
                    void* scratch, size_t scratch_bytes, void* stream);
/* out (b,t,h,w,2) = 0.5 * (unnorm(unpad(uxf)) + unnorm(unpad(uyf))) from the U-Net outputs (same layouts as above). */
int b2s_planes_unpack(const float* uxf, const float* uyf, const float* stats_xf, const float* stats_yf, float* out,
                      int b, int t, int h, int w, int hp, int wp, int tp, int ph0, int pw0, int pt0, void* stream);

/* CineNet normal operator and CG — cinenet.py:121-171, recurrent_cinenet.py:74-124.
 * H x = A^H M A x + v x with the k-space kept on chip: because the mask only selects rows,
 * F_w cancels and H x = sum_c conj(S_c) * (F_h^H M F_h)(S_c x) + v x.  x, out (b,t,h,w,2).
 * h in {200, 256}, w % 4 == 0 (B2S_EUNSUPPORTED otherwise: compose b2s_sens_expand / b2s_sens_reduce). */
int b2s_normal_op(const float* x, const float* sens, const uint8_t* mask, const float* v, float* out,
                  int b, int t, int c, int h, int w, void* stream);
/* One VarNet cascade's data-consistency step entirely in the image domain (k-space never exists):
 *   out = A^H[ DC(A x, ref) ] = ssq . x - v/(1+v) (A^H M A x - bref),   ssq = sum_c |S_c|^2 (b,h,w),
 *   bref = A^H ref (b,t,h,w,2).  Replaces varnet.py:257 + 281-282 of cascade n together with the
 *   sens_reduce of cascade n+1 (varnet.py:253) / of VarNet.forward (varnet.py:150-151). */
int b2s_normal_dc(const float* x, const float* sens, const uint8_t* mask, const float* v, const float* ssq,
                  const float* bref, float* out, int b, int t, int c, int h, int w, void* stream);
/* The last cascade of an inference with the final magnitude fused in: out_abs (b,t,h,w) = | b2s_normal_dc(...) |,
 * i.e. complex_abs(sens_reduce(kspace_pred)) of VarNet.forward (varnet.py:150-151, utils/math.py:41-56). */
int b2s_normal_dc_abs(const float* x, const float* sens, const uint8_t* mask, const float* v, const float* ssq,
                      const float* bref, float* out_abs, int b, int t, int c, int h, int w, void* stream);
/* H x as b2s_normal_op, plus dot_partials[b*t*w/4]: per work item the partial sum of <x, H x> (real inner product over
 * the float pairs) - CG's <p, H p> (cinenet.py:159) without a separate dot kernel.  Fixed summation order. */
int b2s_normal_op_dot(const float* x, const float* sens, const uint8_t* mask, const float* v, float* out,
                      float* dot_partials, int b, int t, int c, int h, int w, void* stream);
/* One CG iteration after d = H p (cinenet.py:159-167) in two launches, all scalars on the device, fixed summation order:
 *   b2s_cg_update:    alpha = rs_old / sum(pd_partials[0..n_pd));  x += alpha p;  r -= alpha d;  rr_partials[b2s_cg_blocks(n)] = partials of <r, r>
 *   b2s_cg_direction: rs_new = sum(rr_partials);  p = r + (rs_new / rs_old) p          (n = floats per vector) */
int b2s_cg_blocks(int64_t n);
int b2s_cg_update(const float* p, const float* d, float* x, float* r, const float* pd_partials, int n_pd,
                  const float* rs_old, float* rr_partials, int64_t n, void* stream);
int b2s_cg_direction(float* p, const float* r, const float* rr_partials, const float* rs_old, float* rs_new,
                     int64_t n, void* stream);
/* CG scalar/vector kernels with alpha, beta kept in device memory (no .item() syncs):
 * dot: out[0] = <a,b> over n floats (deterministic two-stage; scratch >= 1024 floats) */
int b2s_dot(const float* a, const float* b, float* out, int64_t n, float* scratch, void* stream);
/* y = y + (sign * num[0]/den[0]) * x */
int b2s_axpy_ratio(float* y, const float* x, const float* num, const float* den, float sign,
                   int64_t n, void* stream);
/* p = r + (num[0]/den[0]) * p */
int b2s_xpay_ratio(float* p, const float* r, const float* num, const float* den, int64_t n,
                   void* stream);
/* out = a + v[0]*b  (rhs and H x assembly) ; v may be NULL with scale used instead */
int b2s_axpby(const float* a, const float* b, const float* v, float scale, float* out, int64_t n,
              void* stream);

/* End-to-end convenience with HOST buffers (pinned or pageable): one VarNet-style DC cascade
 * k_next = DC(A(A^H k), ref) for a batch; copies in, runs, copies out on `stream`.
 * Used for the e2e measurement; device workspace `ws` of b2s_dc_step_ws_bytes() bytes. */
size_t b2s_dc_step_ws_bytes(int b, int t, int c, int h, int w);
int b2s_dc_step_host(const float* kspace_host, const float* ref_host, const float* sens_host,
                     const uint8_t* mask_host, float v_value, float* out_host, int b, int t, int c,
                     int h, int w, void* ws, size_t ws_bytes, void* stream);

/* Sparse upload of a masked k-space (what data/transforms.py:66-92 apply_mask leaves: unsampled rows are
 * zero).  kspace_host: PINNED host memory (n_bt,c,h,w,2), read by the GPU over PCIe (UVA); only rows with
 * mask (n_bt,h) != 0 are transferred, the others are written as zeros into kspace_dev.  mask: device. */
int b2s_upload_rows(const float* kspace_host, const uint8_t* mask, float* kspace_dev, int64_t n_bt, int c,
                    int h, int w, void* stream);

/* ---- training loss and test metrics -------------------------------------------- *
 * Time-averaged SSIM - utils/losses.py:25-58 (SSIMLoss.forward) and utils/evaluate.py:25-42 (ssim):
 * x, y (b,t,h,w) float32; `win` x `win` uniform window (only 7 is built), "valid" positions, sample
 * covariance; data_range[(frame index t) * dr_stride] read from device memory (dr_stride 1: one value per
 * frame t, as the loss takes Y.max() per frame; 0: one value for the whole volume, as the metric does).
 * out[0..t-1] = mean S of frame t over batch and window positions, out[t] = mean_t (1 - out[t]) = the loss.
 * scratch: b2s_ssim_scratch_floats() floats.  Ordered sums: bit-reproducible. */
size_t b2s_ssim_scratch_floats(int b, int t, int h, int w);
int b2s_ssim_fwd(const float* x, const float* y, const float* data_range, int dr_stride, int b, int t,
                 int h, int w, int win, float k1, float k2, float* out, float* scratch, void* stream);
/* gx = gout[0] * d out[t] / d x  (the loss's gradient w.r.t. the prediction; y and data_range are constants,
 * as in the reference where data_range is rebuilt from a Python float, losses.py:35) */
int b2s_ssim_bwd(const float* x, const float* y, const float* data_range, int dr_stride,
                 const float* gout, int b, int t, int h, int w, int win, float k1, float k2, float* gx,
                 void* stream);
/* out[t] = max over batch and pixels of frame t of y (b,t,hw) - the loss's data_range (losses.py:35) */
int b2s_frame_max(const float* y, float* out, int b, int t, int64_t hw, void* stream);
/* out[0] = sum (gt-pred)^2, out[1] = sum gt^2, out[2] = max gt, out[3] = n  over n floats (ordered two-stage
 * sums; scratch >= 3072 floats): NMSE = out[0]/out[1], PSNR = 10 log10(maxval^2 n / out[0]) -
 * utils/evaluate.py:6-22 */
int b2s_err_stats(const float* gt, const float* pred, int64_t n, float* out, float* scratch, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200SENSE_H */
