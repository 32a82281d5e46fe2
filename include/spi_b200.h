/* spi_b200.h -- C ABI of libspi_b200.so (sm_100a kernels for the SPI inversion hot path).
 *
 * Conventions (restating the reference plugin conventions of SURVEY.md §8b.1 as a plain C ABI):
 *   - every pointer is a DEVICE pointer unless named `*_strides` / `*_out` (host); shapes are plain ints;
 *   - inputs are borrowed, outputs are caller-allocated (the Python shim allocates them with torch);
 *   - work is enqueued on `stream`; no host synchronisation, no internal threads;
 *   - return value: 0 = ok, -1 = invalid argument (Python raises RuntimeError with spi_last_error(), as
 *     TORCH_CHECK does in the reference), -2 = no specialised kernel (soft error: the reference's
 *     `return_code = -1` of filtered_lrelu.cpp:53-56), -3 = CUDA launch failure;
 *   - dtype codes: 0 = float32, 1 = float16, 2 = float64; index math is 64-bit where tensors can exceed INT_MAX.
 *   - a NULL pointer means "operand absent" (the reference passes empty tensors, bias_act.py:138).
 */
#ifndef SPI_B200_H
#define SPI_B200_H

#include <cuda_runtime.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- library state ---------------------------------------------------------------------------------- */
const char* spi_last_error(void);
unsigned long long spi_launch_count(void);   /* kernels launched by this library since the last reset */
void spi_reset_launch_count(void);
int spi_abi_version(void);                   /* 3 for this header; bumped whenever a signature below changes */

/* ---- L1 operators: replace the pybind plugins built by eg3d/torch_utils/custom_ops.py:61 ------------- */

/* bias_act plugin op: eg3d/torch_utils/ops/bias_act.cpp:36-93 (kernel bias_act.cu:28-147).
 * grad = 0/1/2; act = 1..9 (linear, relu, lrelu, tanh, sigmoid, elu, selu, softplus, swish; bias_act.py:23-33);
 * clamp < 0 disables clamping; step_b = x.stride(dim), size_b = b.numel(); x must be dense. */
int spi_bias_act(const void* x, const void* b, const void* xref, const void* yref, const void* dy, void* y,
                 long long numel, int size_b, int step_b, int dtype, int grad, int act, float alpha, float gain,
                 float clamp, cudaStream_t stream);

/* SynthesisLayer epilogue (networks_stylegan2.py:320-329) in one pass: y = clamp(act(x + noise[h,w]*strength + b[c])*gain).
 * noise: fp32 [hw]; noise_strength: device scalar; channels_last_c = C for channels-last x, 0 for NCHW. */
int spi_bias_act_noise(const void* x, const void* b, void* y, const float* noise, const float* noise_strength, long long numel,
                       int size_b, int step_b, int hw, int channels_last_c, int dtype, int act, float alpha, float gain,
                       float clamp, cudaStream_t stream);

/* The tail of an up-sampling SynthesisLayer in one pass: conv2d_resample.py:117-119 (4x4 FIR after the stride-2 transposed
 * convolution) followed by networks_stylegan2.py:320-329 (noise, bias, activation, gain, clamp):
 *   y = clamp(act(upfirdn2d(x, f[4x4], pad, fir_gain) + noise[h,w]*strength + b[c]) * gain)
 * x, y: channels-last fp32 [n, c, h, w] given by element strides (c % 4 == 0, 16-byte aligned); act 1 (linear) or 3 (lrelu);
 * noise (fp32 [out_h*out_w]) may be NULL; clamp < 0 = off.  Same arithmetic and order as spi_upfirdn2d + spi_bias_act_noise. */
int spi_blur4_bias_act_noise(const float* x, const float* f, float* y, const float* b, const float* noise,
                             const float* noise_strength, int n, int c, int in_h, int in_w, const long long* x_strides,
                             const long long* y_strides, int padx0, int padx1, int pady0, int pady1, int flip, float fir_gain,
                             int act, float alpha, float gain, float clamp, cudaStream_t stream);

/* Activation gradient and its reductions in one pass (what `bias_act(grad=1)` followed by the sums of `SynthesisLayer`'s bias /
 * noise_strength gradients computes, eg3d/torch_utils/ops/bias_act.py:150-187 + autograd's sums): dx = act'(yref) * dy, db[c] = sum over pixels
 * of dx (NULL: not wanted), dstrength = sum dx * noise[pixel % hw] (NULL: not wanted).  act 1 linear / 2 relu / 3 lrelu; channels-last rows
 * of c floats (c % 4 == 0, c <= 1024), 16-byte aligned.  db / dstrength are overwritten. */
int spi_bias_act_grad_reduce(const float* dy, const float* yref, float* dx, long long numel, int c, int hw, int act, float alpha, float gain,
                             float clamp, const float* noise, float* db, float* dstrength, cudaStream_t stream);
/* Gradient reductions of that epilogue in one pass over dx (channels-last fp32 [pixels, C]): db[c] = sum dx (bias_act.py:166),
 * dpix[h,w] = sum_{n,c} dx (gradient of the noise term), dstrength = sum dpix*noise.  Any output may be NULL; pixels = n * hw. */
int spi_epilogue_grad_reduce(const float* dx, long long pixels, int c, int hw, const float* noise, float* db, float* dpix,
                             float* dstrength, cudaStream_t stream);

/* upfirdn2d plugin op: eg3d/torch_utils/ops/upfirdn2d.cpp:20-105 (kernels upfirdn2d.cu:33-204).
 * f: fp32 [fh, fw]; strides in elements, order (n, c, h, w); y is [n, c, out_h, out_w] with
 * out = (in*up + pad0 + pad1 - f + down) / down (upfirdn2d.cpp:49-50). */
int spi_upfirdn2d(const void* x, const float* f, void* y, int dtype, int n, int c, int in_h, int in_w,
                  const long long* x_strides, const long long* y_strides, int fh, int fw, int upx, int upy, int downx,
                  int downy, int padx0, int padx1, int pady0, int pady1, int flip, float gain, cudaStream_t stream);

/* filtered_lrelu plugin op: eg3d/torch_utils/ops/filtered_lrelu.cpp:20-213 (kernel filtered_lrelu.cu:144).
 * fu/fd: fp32 rank-2 filters (separable 1-D filters are expanded by the caller); s: uint8 sign tensor
 * [n, c, s_h, s_w_bytes] (4 elements per byte) or NULL; sign_mode 0 none / 1 write / 2 read (+ offsets sx, sy). */
int spi_filtered_lrelu_sign_shape(int yh, int yw, int down, int fdh, int fdw, int* sh_out, int* sw_bytes_out);
int spi_filtered_lrelu(const void* x, void* y, const void* b, unsigned char* s, const float* fu, const float* fd, int dtype,
                       int n, int c, int xh, int xw, const long long* x_strides, const long long* y_strides, int fuh, int fuw,
                       int fdh, int fdw, int up, int down, int px0, int px1, int py0, int py1, int s_h, int s_w_bytes, int sx,
                       int sy, float gain, float slope, float clamp, int flip, int sign_mode, cudaStream_t stream);
/* filtered_lrelu_act_ plugin op: filtered_lrelu.cpp:217-293 (kernel filtered_lrelu.cu:1110); in place on x. */
int spi_filtered_lrelu_act(void* x, unsigned char* s, int dtype, int n, int c, int h, int w, const long long* x_strides,
                           int s_h, int s_w_bytes, int sx, int sy, float gain, float slope, float clamp, int sign_mode,
                           cudaStream_t stream);

/* ---- fused volume renderer: replaces eg3d/training/volumetric_rendering/{renderer,ray_marcher,ray_sampler}.py
 *      and OSGDecoder (eg3d/training/triplane.py:112-135) ---------------------------------------------- */

/* ImportanceRenderer.forward (renderer.py:88-140) fused with RaySampler.forward (ray_sampler.py:24-61).
 * planes: channels-last [n, plane_h, plane_w, 96] with batch stride plane_batch_stride floats (0 = one tri-plane set shared
 * by all n views: the backbone output does not depend on the camera); origins/dirs [n, R, 3]; jitter [n, R, dc]; u [n*R, df];
 * decoder tensors as stored in the state dict (decoder.net.0/2.{weight,bias}), lr_mul = decoder_lr_mul; outputs feat [n, R, 32],
 * depth [n, R] (clamped to the global sample-depth range, ray_marcher.py:49-50), wsum [n, R];
 * depths_all [n, R, dc+df] is the sorted merged depth list the backward pass needs; sigma_all optional (NULL);
 * minmax: 2 ints of scratch that the backward pass re-reads. */
int spi_render_forward(const float* planes, const float* origins, const float* dirs, const float* jitter, const float* u,
                       const float* w1, const float* b1, const float* w2, const float* b2, float lr_mul, float* feat,
                       float* depth, float* wsum, float* depths_all, float* sigma_all, int* minmax, int n, int rays_per_image,
                       long long plane_batch_stride, int plane_h, int plane_w, int dc, int df, float ray_start, float ray_end, float box_warp, int disparity,
                       cudaStream_t stream);
/* Backward of the above w.r.t. planes (accumulated into g_planes, may be NULL; batch stride grad_batch_stride, which may be a full
 * plane set even when the planes themselves are shared by all views: per-view gradient planes keep the views' REDs apart) and, when the sc_* buffers are
 * given, the per-sample rows [S,32] [S,64] [S,64] [S,36] from which the decoder weight gradients are formed
 * (dW1 = dpre^T f, dW2 = dout^T hid; S = n*R*(dc+df)). */
int spi_render_backward(const float* planes, const float* origins, const float* dirs, const float* depths_all, const int* minmax,
                        const float* w1, const float* b1, const float* w2, const float* b2, float lr_mul, const float* g_feat,
                        const float* g_depth, float* g_planes, float* sc_f, float* sc_hid, float* sc_dpre, float* sc_dout, int n,
                        int rays_per_image, long long plane_batch_stride, long long grad_batch_stride, int plane_h, int plane_w, int dc, int df,
                        float box_warp, cudaStream_t stream);
/* Same pair with the forward pass KEEPING its decoder activations for the backward pass (tcgen05 kernels; 180 GB of HBM make the
 * ~0.5 GB per image cheaper than re-gathering 12 texel lines per sample and re-running layers 1-2): sv_h [S,64] hidden layer,
 * sv_o [S,36] pre-activation outputs incl. bias (32 colours, sigma, 3 pad), sv_f [S,32] gathered features (NULL unless decoder
 * weight gradients are wanted), sv_src [n,R,dc+df] uint8 storage index of every merged sample; S = n*R*(dc+df), rows in storage
 * order (coarse i -> i, importance j -> dc + j).  The backward pass then writes sc_dpre / sc_dout in the same row order, so that
 * dW1 = sc_dpre^T sv_f and dW2 = sc_dout^T sv_h.  spi_render_keeps_activations(dc, df) = 1 when these entry points apply. */
int spi_render_keeps_activations(int dc, int df);
int spi_render_forward_keep(const float* planes, const float* origins, const float* dirs, const float* jitter, const float* u,
                            const float* w1, const float* b1, const float* w2, const float* b2, float lr_mul, float* feat, float* depth,
                            float* wsum, float* depths_all, int* minmax, int n, int rays_per_image, long long plane_batch_stride,
                            int plane_h, int plane_w, int dc, int df, float ray_start, float ray_end, float box_warp, int disparity,
                            float* sv_h, float* sv_o, float* sv_f, unsigned char* sv_src, cudaStream_t stream);
int spi_render_backward_kept(const float* planes, const float* origins, const float* dirs, const float* depths_all, const int* minmax,
                             const float* w1, const float* b1, const float* w2, const float* b2, float lr_mul, const float* g_feat,
                             const float* g_depth, float* g_planes, float* sc_dpre, float* sc_dout, int n, int rays_per_image,
                             long long plane_batch_stride, long long grad_batch_stride, int plane_h, int plane_w, int dc, int df, float box_warp,
                             const float* sv_h, const float* sv_o, const unsigned char* sv_src, cudaStream_t stream);
/* ImportanceRenderer.run_model (renderer.py:142-149) on arbitrary points: coords [n, m, 3] -> rgb [n, m, 32],
 * sigma [n, m]; used by TriPlaneGenerator.sample / sample_mixed (triplane.py:91-102). */
int spi_points_forward(const float* planes, const float* coords, const float* w1, const float* b1, const float* w2,
                       const float* b2, float lr_mul, float* rgb, float* sigma, int n, int m, int plane_h, int plane_w, float box_warp,
                       cudaStream_t stream);
int spi_points_backward(const float* planes, const float* coords, const float* w1, const float* b1, const float* w2,
                        const float* b2, float lr_mul, const float* g_rgb, const float* g_sigma, float* g_planes, float* sc_f, float* sc_hid,
                        float* sc_dpre, float* sc_dout, int n, int m, int plane_h, int plane_w, float box_warp,
                        cudaStream_t stream);
/* Stand-alone stages (bit-exact index parity). */
int spi_ray_sampler(const float* cam, int n, int res, float* origins, float* dirs, cudaStream_t stream);      /* ray_sampler.py:24 */
int spi_ray_march(const float* colors, const float* sigma, const float* depths, int rays, int d, int c, float* rgb,
                  float* depth, float* weights, int* minmax, cudaStream_t stream);                            /* ray_marcher.py:25 */
int spi_sample_importance(const float* depths, const float* weights, const float* u, int rays, int dc, int df, float* fine,
                          int* inds, float* cdf, cudaStream_t stream);                                        /* renderer.py:194-253 */
int spi_inverse_cdf(const float* bins, const float* cdf, const float* u, int rays, int ncdf, int nbins, int df, float* fine,
                    int* inds, cudaStream_t stream);                                                          /* renderer.py:241-253 */
int spi_unify_samples(const float* depths_coarse, const float* depths_fine, int rays, int dc, int df, int* perm,
                      float* sorted, cudaStream_t stream);                                                    /* renderer.py:157-167 */

/* ---- modulated-convolution weight preparation: replaces the elementwise chain of modulated_conv2d
 *      (eg3d/training/networks_stylegan2.py:58-68) -------------------------------------------------------- */
/* out[n,o,i,k] = W[o,i,k]*s[n,i] (* rsqrt(sum_{i,k}(W s)^2 + 1e-8) when demodulate); dcoef [n,o] saved for backward.
 * layout of out / grad_out: 0 = [n][o][i][k], 1 = [n][o][k][i] (channels-last conv weight), 2 = [n][i][k][o] (channels-last
 * weight of the stride-2 transposed conv). */
int spi_modulate_weights(const float* weight, const float* styles, float* out, float* dcoef, int n, int o, int i, int kk,
                         int demodulate, int layout, cudaStream_t stream);
int spi_modulate_weights_backward(const float* weight, const float* styles, const float* dcoef, const float* grad_out,
                                  float* grad_weight, float* grad_styles, int n, int o, int i, int kk, int demodulate,
                                  int layout, cudaStream_t stream);

/* The same for every modulated convolution of a network at once (one launch each; per-layer HOST tables of `layers` <= 32 entries):
 * forward (layouts 0 / 1 only), backward (grad_styles[l] must be ZERO on entry: the caller packs them into one zero-filled buffer), and
 * the [g][o][taps][i] -> [g][i][taps'][o] re-layouts the data-gradient convolutions read (spi_conv_weight_transpose for many tensors). */
int spi_modulate_weights_many(int layers, const float* const* weight, const float* const* styles, float* const* out, float* const* dcoef,
                              const int* n, const int* o, const int* i, const int* kk, const int* demodulate, const int* layout,
                              cudaStream_t stream);
int spi_modulate_weights_backward_many(int layers, const float* const* weight, const float* const* styles, const float* const* dcoef,
                                       const float* const* grad_out, float* const* grad_weight, float* const* grad_styles, const int* n,
                                       const int* o, const int* i, const int* kk, const int* demodulate, const int* layout,
                                       cudaStream_t stream);
int spi_conv_weight_transpose_many(int layers, const float* const* w, float* const* wt, const int* g, const int* o, const int* taps,
                                   const int* i, const int* reverse, cudaStream_t stream);

/* ---- style bank: all affine layers of a synthesis network in one launch (FullyConnectedLayer(w_dim, in_channels, bias_init=1)
 *      inside every SynthesisLayer / ToRGBLayer, eg3d/training/networks_stylegan2.py:282,316,352,357-358) ----------------- */
/* out[l][n,i] = ogain[l] * (wgain[l] * sum_k ws[n, widx[l], k] * W[l][i,k] + bgain[l] * b[l][i]).  ws has element strides
 * (ws_sn, ws_sl, 1); the per-layer tables (W, b, out, I, widx, gains; layers <= 32) are HOST arrays read during the call. */
int spi_style_bank_forward(const float* ws, long long ws_sn, long long ws_sl, int n, int k, int layers, const float* const* W,
                           const float* const* b, float* const* out, const int* I, const int* widx, const float* wgain,
                           const float* bgain, const float* ogain, cudaStream_t stream);
/* ds[l] [n][I[l]] (NULL = zero); dW[l] [I[l]][k] / db[l] [I[l]] (NULL = not wanted) overwritten; dws (NULL = not wanted) contiguous
 * [n][num_ws][k], overwritten. */
int spi_style_bank_backward(const float* ws, long long ws_sn, long long ws_sl, int n, int k, int layers, const float* const* W,
                            const float* const* ds, float* const* dW, float* const* db, const int* I, const int* widx,
                            const float* wgain, const float* bgain, const float* ogain, float* dws, int num_ws, cudaStream_t stream);

/* ---- depth-guided 3-D warp: replaces rotate() (spi/utils/rotate.py:92-116) --------------------------- */
/* cameras [n, 25]; depths [n, 1, depth_res, depth_res]; image [n, 3, res, res]; mask [n, 1, res, res] or NULL;
 * *_bs = batch strides in elements of the source tensors (0 broadcasts one source over n views, replacing the
 * `.repeat(rot_bs, ...)` copies of rot_bbox_cx_coach.py:93-99). */
int spi_rotate(const float* target_camera, const float* target_depth, const float* src_image, const float* src_camera,
               const float* src_depth, const float* src_mask, float* out_rgb, float* out_mask, int n, int res, int depth_res,
               long long src_camera_bs, long long src_depth_bs, long long src_image_bs, long long src_mask_bs, float eps,
               cudaStream_t stream);

/* ---- optimiser / streaming helpers -------------------------------------------------------------------- */
/* torch.optim.Adam step (base_coach.py:134, *_projector.py:58) over flat fp32 arenas; hyper (device, optional)
 * = {lr, 1-beta1^t, 1-beta2^t}; zero_grad != 0 clears the gradient arena in the same pass; skip_if_le (device scalar,
 * optional): the step is a no-op when *skip_if_le <= skip_threshold -- the early exit of rot_bbox_cx_coach.py:148-151
 * (`if loss_lpips <= threshold: break` BEFORE `optimizer.step()`) without a host synchronisation. */
int spi_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, long long n, float lr, float beta1,
                  float beta2, float eps, int step, const float* hyper, int zero_grad, const float* skip_if_le,
                  float skip_threshold, cudaStream_t stream);
/* Same update over a device table of `count` rows {param, grad, exp_avg, exp_avg_sq, numel} (5 x int64, device pointers): the
 * gradients are read where autograd left them instead of being accumulated into a gradient arena first.  A row may be a chunk of
 * a tensor (the caller splits large tensors so that the rows are balanced); blocks_per_row CTAs of 256 threads serve each row. */
int spi_adam_step_multi(const void* table, int count, int blocks_per_row, float lr, float beta1, float beta2, float eps, int step,
                        const float* hyper, const float* skip_if_le, float skip_threshold, cudaStream_t stream);
/* F.interpolate(..., (H/2, W/2), mode='bilinear'|'area') at exactly half size (lpips.py:38-39, bbox_cx_loss.py:161-163,
 * w_projector.py:50,83); contiguous [planes, 2*out_h, 2*out_w] -> [planes, out_h, out_w]; backward != 0 runs the adjoint. */
int spi_downsample2x(const float* x, float* y, long long planes, int out_h, int out_w, int backward, cudaStream_t stream);

/* ---- dense convolution on tcgen05 tensor cores (opt-in engine; the default engine of the dense layers is cuDNN) ---------
 * Implicit GEMM for the stride-1 'same' convolutions behind `_conv2d_wrapper` (eg3d/torch_utils/ops/conv2d_resample.py:30-43)
 * as issued by `modulated_conv2d` (eg3d/training/networks_stylegan2.py:34-91), SuperresolutionHybrid8XDC
 * (eg3d/training/superresolution.py:279-290) and the VGG taps (spi/criteria/lpips/networks.py:53-63).
 * x [n,h,w,ci], y [n,h,w,co] channels-last fp32; w [g][co][kh][kw][ci] with g = n when per_sample else 1; correlation with
 * zero padding (kh-1)/2; TF32 operands (round-to-nearest on the TMA load unless flags & 1), fp32 accumulation in tensor memory.
 * Optional fused epilogue = SynthesisLayer tail (networks_stylegan2.py:320-329) / bias_act (bias_act.py:54):
 *   y = clamp(act(acc + noise[h,w]*strength + bias[c]) * gain); act 0 linear, 1 relu, 2 lrelu(slope); clamp < 0 = off.
 * spi_conv2d_tc_supported: ci, co multiples of 32, kh = kw in {1, 3}, w >= 16, h >= 8.
 * spi_conv2d_tc_error: synchronises and returns non-zero if a pipeline barrier timed out since the last call (never hangs). */
int spi_conv2d_tc_supported(int h, int w, int ci, int co, int kh, int kw);
int spi_conv2d_tc(const float* x, const float* w, float* y, int n, int h, int wd, int ci, int co, int kh, int kw, int per_sample,
                  const float* bias, const float* noise, const float* noise_strength, int act, float slope, float gain, float clamp,
                  int flags, cudaStream_t stream);
int spi_conv2d_tc_error(void);
/* same flag for every tcgen05 kernel of the library (convolution, fused renderer): synchronises, returns and clears it */
int spi_tc_error(void);
/* w [g][o][taps][i] -> wt [g][i][taps reversed][o]: weights of the data-gradient convolution (F.conv2d backward w.r.t. input). */
int spi_conv_weight_flip_transpose(const float* w, float* wt, int g, int o, int taps, int i, cudaStream_t stream);

/* ---- conv engine, second generation (spi_b200/csrc/conv_tc2.cu): halo-patch implicit GEMM on tcgen05 + TMEM + TMA ----------
 * Replaces the cuDNN calls behind `_conv2d_wrapper` (eg3d/torch_utils/ops/conv2d_resample.py:30-43, :114-128) for the three
 * convolution forms the generator, the super-resolution module and the VGG extractors issue, forward and data-gradient:
 *   spi_conv2d_tc2               stride-1 'same' correlation, k = 1 or 3 (fused SynthesisLayer / bias_act epilogue as spi_conv2d_tc)
 *   spi_conv_transpose2d_s2_tc2  stride-2 transposed 3x3 convolution, no padding: [n,h,w,ci] -> [n,2h+1,2w+1,co]
 *                                (y[2iy+ky, 2ix+kx, o] += x[iy,ix,i] * w[g][o][ky*3+kx][i]; conv2d_resample.py:117)
 *   spi_conv2d_s2_tc2            stride-2 3x3 correlation, no padding: [n,2h+1,2w+1,ci] -> [n,h,w,co] (data gradient of the former)
 * All tensors channels-last fp32, w [g][co][taps][ci] with g = n when per_sample else 1; ci, co multiples of 32.
 * flags bit 0: raw fp32 bits (truncation) instead of round-to-nearest TF32 on load.  Pipeline time-outs raise spi_tc_error(). */
int spi_conv_tc2_supported(int ci, int co);
int spi_conv2d_tc2(const float* x, const float* w, float* y, int n, int h, int wd, int ci, int co, int k, int per_sample,
                   const float* bias, const float* noise, const float* noise_strength, int act, float slope, float gain, float clamp,
                   int flags, cudaStream_t stream);
int spi_conv_transpose2d_s2_tc2(const float* x, const float* w, float* y, int n, int h, int wd, int ci, int co, int per_sample, int flags,
                                cudaStream_t stream);
int spi_conv2d_s2_tc2(const float* x, const float* w, float* y, int n, int h, int wd, int ci, int co, int per_sample, int flags,
                      cudaStream_t stream);
/* Number of Cin splits the call of form 0 (spi_conv2d_tc2) / 1 (spi_conv_transpose2d_s2_tc2) / 2 (spi_conv2d_s2_tc2) will use; > 1: the
 * output is accumulated with reduce-adds and the entry point zero-fills it first -- unless flags bit 9 (512) says the caller hands in
 * memory that is already zero (one fill per iteration for all such outputs instead of one per call).  The weight-gradient, RGB and
 * modulate-backward entry points take the same promise as mode + 4, which + 4 and layout + 8.  No device work. */
int spi_conv_tc2_splits(int form, int n, int h, int wd, int ci, int co, int k, int per_sample, int epilogue, int flags);
/* Weight gradients of the same convolutions (spi_b200/csrc/conv_wgrad_tc2.cu; replaces cuDNN's wgrad behind the autograd of
 * `_conv2d_wrapper`, conv2d_resample.py:30-43).  mode 0: y = spi_conv2d_tc2(x, w), dy [n,h,w,co] -> dw [g][co][k*k][ci];
 * mode 1: y = spi_conv_transpose2d_s2_tc2(x, w), dy [n,2h+1,2w+1,co] -> dw TRANSPOSED [g][ci][9][co].  g = n when per_sample (one
 * gradient per image) else 1 (summed over the batch); dw is overwritten; partial sums are combined with fp32 reduce-adds. */
int spi_conv_wgrad_tc2(const float* x, const float* dy, float* dw, int n, int h, int wd, int ci, int co, int k, int per_sample, int mode,
                       cudaStream_t stream);
/* Tall-skinny reduction over `rows` rows (a multiple of 8) on the same kernel: out[m][c] = sum_r u[r][m] * v[r][c] and, when usum is not
 * NULL, usum[m] = sum_r u[r][m].  u [rows][cu], v [rows][cv] row-major fp32, cu / cv multiples of 4 and >= 32; TF32 operands, fp32
 * accumulation; out [cu][cv] and usum [cu] are overwritten.  Replaces the two GEMMs + column sums behind the decoder gradients of the
 * renderer (OSGDecoder, eg3d/training/triplane.py:112-135, as autograd differentiates it). */
int spi_rows_outer_sum(const float* u, const float* v, long long rows, int cu, int cv, float* out, float* usum, cudaStream_t stream);
/* 1x1 convolution onto at most 4 output channels (the RGB heads of the super-resolution ToRGB layers, networks_stylegan2.py:503-518):
 * a streaming op (1.5 FLOP/byte), plain coalesced kernels (spi_b200/csrc/conv_rgb.cu).  x [n][pixels][ci] (ci a multiple of 128, <= 512),
 * y / dy [n][pixels][co], w / dw [g][co][ci], g = n when per_sample else 1.
 * which = 0: y = x w^T (a = x, b = w, c = y); 1: dx = dy w (a = dy, b = w, c = dx); 2: dw = dy^T x (a = x, b = dy, c = dw, overwritten). */
int spi_conv1x1_rgb_supported(int ci, int co);
int spi_conv1x1_rgb(int which, const float* a, const float* b, float* c, long long pixels, int n, int ci, int co, int per_sample,
                    cudaStream_t stream);
/* w [g][o][taps][i] -> wt [g][i][taps'][o], taps' reversed when `reverse` (stride-1 data gradient) else kept (stride-2 forms). */
int spi_conv_weight_transpose(const float* w, float* wt, int g, int o, int taps, int i, int reverse, cudaStream_t stream);

/* ---- mirror-view contextual loss, spi/criteria/bbox_cx_loss.py (spi_b200/csrc/boxcx.cu) ---------------------------------------------
 * spi_roi_align: torchvision.ops.roi_align(input, rois, output_size=pooled, spatial_scale=1, sampling_ratio=-1, aligned=False) as called at
 *   bbox_cx_loss.py:47-57; in [n,c,h,w] fp32 with element strides[4], rois [k][5] = (batch index, x1, y1, x2, y2), out [k,c,pooled,pooled]
 *   contiguous.  spi_roi_align_backward ACCUMULATES the gradient into gin (same strides; the caller zeroes it).
 * spi_cx_rows_forward: S [b,m,n] cosine similarities -> the chain of bbox_cx_loss.py:116-131,176: d = 1 - S, relative distance to the row
 *   minimum (+1e-5, clamp +-10), w = exp((1 - d~) / band_width), cx = w / rowsum, colmax[b,j] = max_i cx[b,i,j]; stats [b,m,2] keeps the
 *   row minimum and row sum for the backward pass.  n <= 2048.
 * spi_cx_rows_backward: gcol [b,n] = dL/dcolmax -> dS [b,m,n]. */
int spi_roi_align(const float* in, const float* rois, float* out, int n, int c, int h, int w, const long long* strides, int k, int pooled,
                  cudaStream_t stream);
int spi_roi_align_backward(const float* gout, const float* rois, float* gin, int n, int c, int h, int w, const long long* strides, int k, int pooled,
                           cudaStream_t stream);
int spi_cx_rows_forward(const float* S, int b, int m, int n, float band_width, float* stats, float* colmax, cudaStream_t stream);
int spi_cx_rows_backward(const float* S, int b, int m, int n, float band_width, const float* stats, const float* colmax, const float* gcol,
                         float* dS, cudaStream_t stream);

/* ---- stage-1 noise-buffer regulariser + re-normalisation: spi/training/projectors/mirror_projector.py:107-115,128-131 (same loops
 *      in w_projector.py / w_plus_projector.py), all buffers in one launch each.
 * table: device array of `count` records {float* x; long long out_off; int size; int pad} describing square fp32 [size,size]
 * buffers (size a power of two <= 256).  forward: partial[b] = sum over the average-pool pyramid (down to size <= 8) of
 * mean(x*roll(x,1,W))^2 + mean(x*roll(x,1,H))^2; stats[b][8][2] keeps the per-level means for backward.
 * backward: (out_base + out_off_b)[i,j] = gout[0] * d(sum_b partial[b]) / d x_b[i,j].  renorm: x -= mean(x); x *= rsqrt(mean(x^2)). */
int spi_noise_reg_forward(const void* table, int count, int max_size, float* partial, float* stats, cudaStream_t stream);
int spi_noise_reg_backward(const void* table, int count, int max_size, const float* stats, const float* gout, float* out_base,
                           cudaStream_t stream);
int spi_noise_renorm(const void* table, int count, cudaStream_t stream);

/* ---- one LPIPS feature tap: spi/criteria/lpips/lpips.py:50-71 + utils.normalize_activation, fused.
 * x: raw VGG features of the generated image, channels-last fp32 [n, hw, c]; yn: unit-normalised target features [ny, hw, c]
 * (ny = n or 1, constant); lin: 1x1 lin-layer weights [c].  forward ACCUMULATES out[0] += sum_n mean_hw sum_c lin_c (xn_c - yn_c)^2
 * with xn = x / (sqrt(sum_c x^2) + 1e-10); backward writes dx = gout[0] * d(tap)/dx.  sample_weight [n] (NULL = ones) weights the sum
 * over n: the two views of the mirror projector (`mirror_projector.py:100-104`: lpips(img, target) + w_m * lpips(img_m, target_m)) share
 * one pass of the VGG trunk that way. */
int spi_lpips_tap_forward(const float* x, const float* yn, const float* lin, int n, int hw, int c, int ny, float* out,
                          const float* sample_weight, cudaStream_t stream);
int spi_lpips_tap_backward(const float* x, const float* yn, const float* lin, int n, int hw, int c, int ny, const float* gout, float* dx,
                           const float* sample_weight, cudaStream_t stream);

/* out[c] = sum over rows of x[row, c]; x row-major [rows, cols] fp32, cols a multiple of 4 (<= 128).  Used for the decoder bias
 * gradients db1 = sum dpre, db2 = sum dout over the per-sample rows of spi_render_backward (triplane.py:123-135). */
int spi_column_sums(const float* x, long long rows, int cols, float* out, cudaStream_t stream);

/* 2x2 / stride-2 max pooling of channels-last fp32 activations [n, h, w, c] (VGG feature extractors of the losses,
 * spi/criteria/lpips/networks.py:75-80, bbox_cx_loss.py:79-87).  backward = 0: out [n, h/2, w/2, c] = max over the window;
 * backward = 1: out [n, h, w, c] = dy routed to the first maximum of every window (torch's arg-max rule), zeros elsewhere. */
int spi_maxpool2x2(const float* x, const float* dy, float* out, int n, int h, int w, int c, int backward, cudaStream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* SPI_B200_H */
