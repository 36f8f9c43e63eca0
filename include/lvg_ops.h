/*
 * lvg_ops.h -- C ABI of liblvg_ops.so, the sm_100a operator library behind the
 * torch_utils.ops drop-in (bias_act, upfirdn2d, filtered_lrelu, conv2d, fma).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name ends in _host;
 *   - the caller owns all buffers (outputs are allocated by the host side,
 *     e.g. torch.empty) -- the library never allocates device memory that
 *     outlives a call, and keeps no mutable global device state, so calls are
 *     safe from any thread on any stream;
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream);
 *     calls only enqueue work and never synchronise;
 *   - strides are in ELEMENTS, shapes are [N, C, H, W] order;
 *   - return 0 = launched, LVG_UNSUPPORTED (-1) = no kernel for this
 *     configuration (the caller may compose other entry points, mirroring the
 *     reference plugin's return code, filtered_lrelu.cpp:53-57), >0 = argument
 *     or CUDA error; lvg_last_error() then describes it (thread local).
 *
 * Each entry point names the reference plugin function it replaces
 * (paths relative to the NVlabs/long-video-gan tree).
 */
#ifndef LVG_OPS_H_
#define LVG_OPS_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LVG_ABI_VERSION 1

#define LVG_OK            0
#define LVG_UNSUPPORTED (-1)
#define LVG_ERR_ARG       1
#define LVG_ERR_CUDA      2

/* element types */
#define LVG_F32 0
#define LVG_F16 1
#define LVG_F64 2

/* activation codes: same numbering as `cuda_idx` in torch_utils/ops/bias_act.py:21-31 */
#define LVG_ACT_LINEAR   1
#define LVG_ACT_RELU     2
#define LVG_ACT_LRELU    3
#define LVG_ACT_TANH     4
#define LVG_ACT_SIGMOID  5
#define LVG_ACT_ELU      6
#define LVG_ACT_SELU     7
#define LVG_ACT_SOFTPLUS 8
#define LVG_ACT_SWISH    9

int         lvg_abi_version(void);
const char* lvg_last_error(void);
/* compile-time facts about the build: "sm_100a;..." */
const char* lvg_build_info(void);
/* number of kernels this library has launched in the calling process so far */
int64_t     lvg_launch_count(void);

/*
 * bias_act -- replaces bias_act_plugin.bias_act (torch_utils/ops/bias_act.cpp:32-90,
 * kernel bias_act.cu:23-147).
 *   grad = 0: y = clamp(act(x + b) * gain)
 *   grad = 1: y = x * gain * act'(.)   , zero where yref is outside (-clamp, clamp)
 *   grad = 2: y = x * dy * gain * act''(.), same masking
 * x, xref, yref, dy, y are dense buffers of n elements in one common layout;
 * the bias of element i is b[(i / step_b) % size_b]. NULL = operand absent.
 * clamp < 0 disables clamping.
 */
int lvg_bias_act(const void* x, const void* b, const void* xref, const void* yref,
                 const void* dy, void* y, int dtype, int64_t n, int64_t size_b,
                 int64_t step_b, int grad, int act, float alpha, float gain,
                 float clamp, void* stream);

/*
 * bias_act backward with the bias-gradient reduction fused in: as grad = 1
 * above, and additionally db[c] += sum of y over all elements whose bias index
 * is c (fp32 accumulators, db_f32 has size_b floats and must be zeroed by the
 * caller). Saves the extra read of dx that `dx.sum(...)` costs
 * (torch_utils/ops/bias_act.py:169-170).
 */
int lvg_bias_act_grad_db(const void* dy_in, const void* b, const void* xref,
                         const void* yref, void* dx, float* db_f32, int dtype,
                         int64_t n, int64_t size_b, int64_t step_b, int act,
                         float alpha, float gain, float clamp, void* stream);
/*
 * relu / lrelu with 2-bit codes instead of a saved output. The gradient of these activations depends on the
 * forward output only through "was it positive" and "was it saturated by the clamp"; the forward pass can
 * emit exactly those two bits per element (bit 0 = not positive, bit 1 = clamped) and the backward pass then
 * reads dy + codes (2 n s + n / 4 bytes) instead of dy + y (3 n s) -- same results as lvg_bias_act(grad = 1) /
 * lvg_bias_act_grad_db with yref = y (the reference always re-reads y: bias_act.py:160-172, bias_act.cu:60-118).
 * The codes buffer is opaque (tile-local word order of the kernels), 8-byte aligned, lvg_bias_act_codes_bytes(dtype, n)
 * bytes long (about n / 4), and only meaningful to lvg_bias_act_bwd_codes for a dy with the memory layout of x.
 * fwd: y and codes from x (+ b). bwd: dx from dy and codes; db_f32 != NULL also accumulates the bias gradient
 * (zero-initialised fp32 [size_b], bias along a non-contiguous dimension). Returns -1 when the activation is not
 * relu / lrelu, the type not fp32 / fp16, n not a multiple of the 16-byte pack or an operand unaligned.
 */
int lvg_bias_act_fwd_codes(const void* x, const void* b, void* y, void* codes, int dtype, int64_t n,
                           int64_t size_b, int64_t step_b, int act, float alpha, float gain, float clamp,
                           void* stream);
int lvg_bias_act_bwd_codes(const void* dy, const void* codes, void* dx, float* db_f32, int dtype, int64_t n,
                           int64_t size_b, int64_t step_b, int act, float alpha, float gain, float clamp,
                           void* stream);
int64_t lvg_bias_act_codes_bytes(int dtype, int64_t n);

/*
 * upfirdn2d -- replaces upfirdn2d_plugin.upfirdn2d (torch_utils/ops/upfirdn2d.cpp:16-98,
 * kernels upfirdn2d.cu:29-200). Full 2-D filter f[fh][fw] (float32, element
 * strides f_stride_y / f_stride_x). Output size per axis:
 *   out = (in * up + pad0 + pad1 - ftaps + down) / down      (upfirdn2d.cpp:35-36)
 * so y_shape carries the caller's choice of pad1. Any x / y strides.
 */
int lvg_upfirdn2d(const void* x, const float* f, void* y, int dtype,
                  const int64_t x_shape[4], const int64_t x_stride[4],
                  const int64_t y_shape[4], const int64_t y_stride[4],
                  int fw, int fh, int64_t f_stride_x, int64_t f_stride_y,
                  int upx, int upy, int downx, int downy, int padx0, int pady0,
                  int flip, float gain, void* stream);

/*
 * Separable upfirdn2d in ONE launch: horizontal pass with fx[fw], vertical
 * pass with fy[fh], the intermediate stays in shared memory. Replaces the two
 * chained plugin calls of torch_utils/ops/upfirdn2d.py:244-245 (and their HBM
 * round trip). fx == NULL or fy == NULL means "no filter along that axis"
 * (1 tap of weight 1). `gain` is applied once.
 * Returns LVG_UNSUPPORTED for shapes the tiled kernel does not cover
 * (caller then issues two lvg_upfirdn2d calls).
 */
int lvg_upfirdn2d_sep(const void* x, const float* fx, const float* fy, void* y,
                      int dtype, const int64_t x_shape[4], const int64_t x_stride[4],
                      const int64_t y_shape[4], const int64_t y_stride[4],
                      int fw, int fh, int upx, int upy, int downx, int downy,
                      int padx0, int pady0, int flip, float gain, void* stream);

/*
 * filtered_lrelu -- replaces filtered_lrelu_plugin.filtered_lrelu
 * (torch_utils/ops/filtered_lrelu.cpp:16-209, kernel filtered_lrelu.cu:139-1099):
 *   y = downfir( clamp( lrelu( upfir(x + b) * up^2 * gain ) ) )
 * fu / fd: float32, separable when f?_h == 0 (f?_w taps used on both axes),
 * else full [f?_h][f?_w] row-major. Sign tensor: uint8 [N][C][s_h][s_wbytes],
 * 2 bits per up-sampled sample, 4 samples per byte along x (bit0 = negative,
 * bit1 = clamped; filtered_lrelu.cu:494-519). Exactly one of the modes:
 *   write_signs != 0 : so written (forward with gradients)
 *   si != NULL       : signs read at offset (sx, sy) instead of evaluating lrelu/clamp (backward)
 *   neither          : plain forward
 * Returns LVG_UNSUPPORTED when no fused kernel covers (up, down, filter sizes).
 */
int lvg_filtered_lrelu(const void* x, const float* fu, const float* fd, const void* b,
                       const uint8_t* si, void* y, uint8_t* so, int dtype,
                       const int64_t x_shape[4], const int64_t x_stride[4],
                       const int64_t y_shape[4], const int64_t y_stride[4],
                       int fu_w, int fu_h, int fd_w, int fd_h, int up, int down,
                       int px0, int py0, int s_h, int s_wbytes, int sx, int sy,
                       float gain, float slope, float clamp, int flip,
                       int write_signs, void* stream);

/* 0 if lvg_filtered_lrelu has a fused kernel for this configuration, else LVG_UNSUPPORTED */
int lvg_filtered_lrelu_supported(int dtype, int fu_w, int fu_h, int fd_w, int fd_h,
                                 int up, int down);

/*
 * In-place gain * lrelu * clamp with sign write / read -- replaces
 * filtered_lrelu_plugin.filtered_lrelu_act_ (filtered_lrelu.cpp:213-290,
 * kernel filtered_lrelu.cu:1105-1211). x is modified in place.
 */
int lvg_filtered_lrelu_act(void* x, const uint8_t* si, uint8_t* so, int dtype,
                           const int64_t x_shape[4], const int64_t x_stride[4],
                           int s_h, int s_wbytes, int sx, int sy, float gain,
                           float slope, float clamp, int write_signs, void* stream);

/*
 * fma -- out = a * b + c with numpy-style broadcasting (torch_utils/ops/fma.py:15-25).
 * All operands are described on the common broadcast shape (rank <= 6):
 * a stride of 0 marks a broadcast dimension. out is dense in `shape` order.
 */
int lvg_fma(const void* a, const void* b, const void* c, void* out, int dtype,
            int rank, const int64_t shape[6], const int64_t a_stride[6],
            const int64_t b_stride[6], const int64_t c_stride[6], void* stream);

/*
 * Grouped 2-D convolution (cross-correlation, like torch.nn.functional.conv2d)
 * on tcgen05 tensor cores -- the kernel behind conv2d_gradfix.conv2d
 * (torch_utils/ops/conv2d_gradfix.py:37-40) for the per-sample-weight
 * "modulated" convolutions (model/generator_sres.py:63-65) and the
 * discriminator convolutions (conv2d_resample.py:29-41).
 *   x [N][G*Cin][H][W]  w [G*Cout][Cin][kh][kw]  y [N][G*Cout][Ho][Wo]
 * NCHW-contiguous fp16 operands, fp32 accumulation in tensor memory, stride 1, 3x3 or 1x1,
 * any batch n (samples share the weights of their group). `workspace` holds the
 * re-tiled weights (lvg_conv2d_fprop_workspace bytes, 16-byte aligned).
 * Returns LVG_UNSUPPORTED outside the covered envelope.
 */
int lvg_conv2d_fprop(const void* x, const void* w, void* y, int dtype,
                     int n, int groups, int cin, int cout, int h, int wd,
                     int kh, int kw, int stride, int pad_h, int pad_w,
                     void* workspace, int64_t workspace_bytes, void* stream);
int64_t lvg_conv2d_fprop_workspace(int dtype, int n, int groups, int cin, int cout,
                                   int h, int wd, int kh, int kw, int stride,
                                   int pad_h, int pad_w);
/*
 * Gradient of the convolution above with respect to its input (same kernel, the weights
 * repacked channel-transposed and spatially mirrored, padding k-1-pad): dx [N][G*Cin][H][W] from
 * dy [N][G*Cout][Ho][Wo]. The argument list describes the FORWARD convolution. Same workspace.
 */
int lvg_conv2d_dgrad(const void* dy, const void* w, void* dx, int dtype,
                     int n, int groups, int cin, int cout, int h, int wd,
                     int kh, int kw, int stride, int pad_h, int pad_w,
                     void* workspace, int64_t workspace_bytes, void* stream);
/*
 * Gradient of the same convolution with respect to its weights: dw [G*Cout][Cin][kh][kw] (fp16, summed
 * over the N samples) from x [N][G*Cin][H][W] and dy [N][G*Cout][Ho][Wo]. Replaces the weight-gradient
 * leg of conv2d_gradfix (conv2d_gradfix.py:119-141 -> aten::convolution_backward / cuDNN). The argument
 * list describes the FORWARD convolution. No workspace. Returns -1 outside fp16 / stride 1 / 3x3, 1x1.
 */
int lvg_conv2d_wgrad(const void* x, const void* dy, void* dw, int dtype,
                     int n, int groups, int cin, int cout, int h, int wd,
                     int kh, int kw, int stride, int pad_h, int pad_w, void* stream);

/*
 * Grouped 1-D / 2-D / 3-D convolution (cross-correlation) on tcgen05 tensor cores, TMA-fed (csrc/conv_igemm.cu): one
 * engine for conv2d_gradfix.conv2d / conv_transpose2d (conv2d_gradfix.py:37-45), the F.conv3d calls of the low-res
 * networks (generator_lres.py:119,578; discriminator_lres.py:172) and the F.conv1d calls of the low-res discriminator
 * (discriminator_lres.py:108-127).
 *   x [N][G*Cin][T][H][W]   w [G*Cout][Cin][kt][kh][kw]   y [N][G*Cout][To][Ho][Wo]      (dense; 2-D: T = kt = 1)
 * kh*kw <= 9, kt <= 7; `stride` (1..4) applies to H and W (T: 1) -- a strided forward pass computes the stride-1 result and
 * stores every stride-th row / column, its gradients spread dy over that lattice. dtype LVG_F16: fp16 operands, fp32 accumulation, fp16 result. dtype LVG_F32: operands
 * split into bf16 hi + lo halves, hi*hi + lo*hi + hi*lo accumulated in fp32 (fp32-grade result, relative error
 * ~2^-16; the reference runs these layers in strict fp32, train_lres.py:269). Optional fused epilogue on fprop:
 * act 0 = none, 1 = (y + bias[co]) * gain clamped, 2 = lrelu(y + bias[co], alpha) * gain clamped (bias_act semantics,
 * bias_act.py:52-86; clamp < 0 = none; bias indexed g*Cout + co, may be NULL).
 * dgrad / wgrad: gradients with respect to x / w; their argument lists describe the FORWARD convolution.
 * `workspace`: lvg_convnd_workspace (fprop, dgrad) or lvg_convnd_wgrad_workspace bytes, 16-byte aligned, holds the
 * re-tiled operands (and the split-K partial sums of wgrad). Return LVG_UNSUPPORTED (-1 from the size queries) outside
 * the envelope.
 */
int64_t lvg_convnd_workspace(int dtype, int n, int groups, int cin, int cout, int t, int h, int wd,
                             int kt, int kh, int kw, int pad_t, int pad_h, int pad_w);
int lvg_convnd_fprop(const void* x, const void* w, void* y, int dtype, int n, int groups, int cin, int cout,
                     int t, int h, int wd, int kt, int kh, int kw, int pad_t, int pad_h, int pad_w, int stride,
                     const float* bias, int act, float alpha, float gain, float clamp,
                     void* workspace, int64_t workspace_bytes, void* stream);
int lvg_convnd_dgrad(const void* dy, const void* w, void* dx, int dtype, int n, int groups, int cin, int cout,
                     int t, int h, int wd, int kt, int kh, int kw, int pad_t, int pad_h, int pad_w, int stride,
                     void* workspace, int64_t workspace_bytes, void* stream);
int64_t lvg_convnd_wgrad_workspace(int dtype, int n, int groups, int cin, int cout, int t, int h, int wd,
                                   int kt, int kh, int kw, int pad_t, int pad_h, int pad_w);
int lvg_convnd_wgrad(const void* x, const void* dy, void* dw, int dtype, int n, int groups, int cin, int cout,
                     int t, int h, int wd, int kt, int kh, int kw, int pad_t, int pad_h, int pad_w, int stride,
                     void* workspace, int64_t workspace_bytes, void* stream);
/*
 * Introspection: the tiling lvg_convnd_fprop (mode 0) / lvg_convnd_dgrad (mode 1) launches with, as 48 ints -- wgroups, cout
 * (rows of the GEMM), mt, kc, nblk, nimg, lo_blk, to, ho, wo, kt, kh, kw, pad_t, pad_h, pad_w, tt, th, wt, wtb, thb, frame_px,
 * ncols, n0, epi_warps, nbuf, tiles_x, tiles_y, tiles_t, total_tiles, ks, stages, a_resident, a_stage, b_step, b_bytes, b_box,
 * stage_bytes, ostride, hos, wos, 0..., [47] = 1 when the call takes the streaming 1x1x1 kernels instead -- host arithmetic only
 * (no device needed): tests/test_igemm_emul.py replays the kernel's addressing with it on the CPU. The argument list describes
 * the FORWARD convolution in both modes.
 */
int lvg_convnd_plan(int mode, int dtype, int n, int groups, int cin, int cout, int t, int h, int wd, int kt, int kh, int kw,
                    int pad_t, int pad_h, int pad_w, int stride, int* out, int out_len);
/*
 * Introspection: the tiling lvg_convnd_wgrad launches with, as 32 ints -- split, cpad_a, cpad_b, nt, ntiles, mt, nsplit,
 * ablk, khc, nseg, ps, rh, stages, a_stage, b_stage, stage_bytes, tail_bytes, smem, seg_w[4], seg_x0[4], pointwise, mrows, 0... --
 * host arithmetic only (no device needed): tests/test_wgrad_emul.py replays the kernel's addressing with it on the CPU.
 */
int lvg_convnd_wgrad_plan(int dtype, int n, int groups, int cin, int cout, int t, int h, int wd, int kt, int kh, int kw,
                          int pad_t, int pad_h, int pad_w, int* out, int out_len);
/*
 * Both gradients of one convolution call (what autograd asks of F.conv3d / conv2d_gradfix in a first-order backward pass,
 * conv2d_gradfix.py:118-141): dx as lvg_convnd_dgrad, dw as lvg_convnd_wgrad, with dy re-tiled ONCE for the two kernels
 * (the separate entry points re-tile it once each). Argument list = the FORWARD convolution; `workspace`:
 * lvg_convnd_backward_workspace bytes.
 */
int64_t lvg_convnd_backward_workspace(int dtype, int n, int groups, int cin, int cout, int t, int h, int wd,
                                      int kt, int kh, int kw, int pad_t, int pad_h, int pad_w);
int lvg_convnd_backward(const void* x, const void* dy, const void* w, void* dx, void* dw, int dtype, int n, int groups,
                        int cin, int cout, int t, int h, int wd, int kt, int kh, int kw, int pad_t, int pad_h, int pad_w,
                        int stride, void* workspace, int64_t workspace_bytes, void* stream);

/*
 * Depthwise long FIR along the last axis (cross-correlation, no padding), fp32:
 *   y[n][g][t] = sum_k w[g][k] * x[n][g][t + k],   x [n][groups][lin], w [groups][k], y [n][groups][lin - k + 1]
 * -- BlurredNoise.blur of the low-res generator: F.conv1d(noise, blur_filters [128, 1, 5000], groups = 128)
 * (model/generator_lres.py:378-387). Leading zero taps of each filter are skipped. `workspace`:
 * lvg_fir1d_depthwise_workspace(groups) bytes. Forward only (the input is noise, the filters are buffers).
 */
int64_t lvg_fir1d_depthwise_workspace(int groups);
int lvg_fir1d_depthwise(const float* x, const float* w, float* y, int n, int groups, int lin, int k,
                        void* workspace, int64_t workspace_bytes, void* stream);

/*
 * Post-processing of an all-reduced flat gradient buffer, in place and in one pass:
 *   g = g * scale;  NaN -> 0, +inf -> +limit, -inf -> -limit  (finite values untouched)
 * -- the `/ world_size * gain` + `nan_to_num(nan=0, posinf=1e5, neginf=-1e5)` tail of
 * utils.sync_grads (utils.py:116-124) without its extra passes over the buffer. fp32 only.
 */
int lvg_grad_postprocess(float* g, int64_t n, float scale, float limit, void* stream);

/*
 * Fused tail of a training update over flat fp32 buffers, one pass, in place:
 *   g' = nan_to_num(g * grad_scale, nan = 0, +-inf = +-grad_limit)      only if grad_limit > 0 (utils.py:120-121)
 *   m  = lerp(m, g', 1 - beta1);  v = v * beta2 + (1 - beta2) * g'^2
 *   p  = p - lr / (1 - beta1^step) * m / (sqrt(v) / sqrt(1 - beta2^step) + eps)
 *        -- torch.optim.Adam.step() as the reference configures it (model/video_gan_lres.py:84-85: no weight decay,
 *           no amsgrad), `step` counting from 1
 *   p_ema = lerp(p_ema, p, 1 - ema_beta)                                 only if p_ema != NULL
 *        -- update_G_ema (model/video_gan_lres.py:208-214)
 * write_grad != 0 stores g' back (what utils.sync_grads leaves in .grad). Replaces ~300 per-tensor launches.
 */
int lvg_adam_step(float* p, float* g, float* m, float* v, float* p_ema, int64_t n, float lr, float beta1, float beta2,
                  float eps, int64_t step, float grad_scale, float grad_limit, int write_grad, float ema_beta, void* stream);

/* a = torch.lerp(a, b, weight) over flat fp32 buffers (the buffers of update_G_ema) */
int lvg_lerp(float* a, const float* b, int64_t n, float weight, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LVG_OPS_H_ */
