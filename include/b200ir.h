/* b200ir.h — C ABI of libb200ir.so: the B200 (sm_100a) kernels behind the GFPGANv1OCR forward pass
 * and the pyblur degradation of ChuRuaNh0/Image_Restoration (Car_Plate-Restoration).
 *
 * This is the drop-in native boundary.  The reference's own native boundary on this path is two
 * pybind11 modules:
 *   fused_act_ext.fused_bias_act(input, bias, refer, act, grad, alpha, scale)
 *       -> basicsr/ops/fused_act/src/fused_bias_act.cpp:14-26, kernel fused_bias_act_kernel.cu:20-50
 *   upfirdn2d_ext.upfirdn2d(input, kernel, up_x, up_y, down_x, down_y, pad_x0, pad_x1, pad_y0, pad_y1)
 *       -> basicsr/ops/upfirdn2d/src/upfirdn2d.cpp:13-24, kernels upfirdn2d_kernel.cu:51-208
 * plus the library calls (F.conv2d / F.conv_transpose2d / F.linear / F.interpolate) made from
 * basicsr/archs/stylegan2_ocr_arch.py and basicsr/archs/gfpganv1_ocr_arch.py, and scipy.signal.convolve2d /
 * cv2.resize made from pyblur/ and basicsr/data/ffhq_degradation_dataset.py.  Each entry point below names
 * the reference call it replaces.
 *
 * Conventions (same as the reference ops: asynchronous, ordered on the caller's stream, no hidden sync):
 *   - plain pointers and sizes only; all pointers are DEVICE pointers unless a name ends in _host;
 *   - the caller owns every buffer (inputs, outputs, workspace); nothing is allocated or freed here;
 *   - `stream` is a cudaStream_t passed as void*; 0 is the legacy default stream;
 *   - every function returns 0 on success, non-zero on error; b200ir_last_error() gives the message of the
 *     last failing call on the calling thread;
 *   - activations are NHWC fp16 (`__half`), i.e. [B][H][W][C] with C contiguous, unless stated otherwise;
 *   - there is no CPU fallback: on a machine without an sm_100 GPU every compute entry point returns an error.
 */
#ifndef B200IR_H_
#define B200IR_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200IR_ABI_VERSION 3   /* 3: b200ir_degrade_crop.mask_mode (appended), blur_mode 5, b200ir_degrade_full_ex */
#define B200IR_MAX_TAPS 16
#define B200IR_MAX_VIEWS 4

const char* b200ir_last_error(void);
int b200ir_abi_version(void);
/* Number of kernels this library has launched since load (all entry points, all threads). */
uint64_t b200ir_launch_count(void);
/* 0 if the current device can run the sm_100a kernels, else non-zero (+ last_error). */
int b200ir_device_check(void);

/* ------------------------------------------------------------------------------------------------
 * Implicit-GEMM convolution on tcgen05 tensor cores (TMA-fed, TMEM accumulators).
 *
 * Replaces, on the GFPGANv1OCR path: F.conv2d of EqualConv2d.forward (stylegan2_ocr_arch.py:639-648),
 * the grouped F.conv2d / F.conv_transpose2d of ModulatedConv2d.forward (:261-277), F.conv2d of
 * ConvUpLayer.forward (gfpganv1_ocr_arch.py:192-198), F.linear of EqualLinear.forward
 * (stylegan2_ocr_arch.py:165-175, as a 1x1 conv over a [B,1,1,K] view), with the FusedLeakyReLU
 * (fused_bias_act_kernel.cu:27-48), demodulation (:253-257), noise injection (:326-330) and the residual
 * merges of ResBlock / ResUpBlock (:733, gfpganv1_ocr_arch.py:224) fused into the epilogue.
 *
 * GEMM view: M = output positions (b, y, x) over extents (m_b, m_h, m_w), tiled (tile_b, tile_h, tile_w) with
 * tile_b*tile_h*tile_w == 128; N = cout tiled by block_n; K = num_taps * cin.
 * Tap t reads input view tap_view[t] at (b, y + tap_dy[t], x + tap_dx[t], :); out-of-range reads are zero
 * (TMA out-of-bounds fill), which implements the zero padding of every conv on this path.
 * Weights are fp16 [cout][num_taps*cin] (K contiguous, tap-major).
 *
 * Epilogue, per output element (fp32 math):
 *   v = acc * (demod ? demod[b*cout + n] : 1)
 *     + (noise ? noise_gain[0] * noise[b*noise_stride_b + yo*out_w_full + xo] : 0) + (bias ? bias[n] : 0)
 *   if (act) v = (v > 0 ? v : 0.2 v) * sqrt(2)
 *   if (res_mode == 1) v = (v + res[b, yo, xo, n]) * res_scale
 *   if (res_mode == 2) v = (v + bilinear_up2(res_lowres)[b, yo, xo, n]) * res_scale   (align_corners=False)
 *   if (rgb_w)  rgb_part[nt][b][o][yo][xo] = sum_{n in N-tile nt} v * rgb_w[b][o][n]   (o = 0..2; ToRGB fused, see below)
 *   if (out_scale) v *= out_scale[b*cout + n]      (modulation of the NEXT modulated conv, stylegan2_ocr_arch.py:247-251)
 *   if (!no_store) out[b, yo, xo, out_c_off + n] = v      with yo = y*out_y_mul + out_y_off, xo = x*out_x_mul + out_x_off
 *
 * Fused ToRGB (stylegan2_ocr_arch.py:357-374): the 1x1 modulated conv to 3 channels reads exactly the tensor this
 * conv produces, so each epilogue thread (one output position, block_n channels) accumulates the three dot products
 * with rgb_w[b][o][n] = w[o][n]/sqrt(cout) * s_rgb[b][n] and writes them as fp32 NCHW partial planes, one per N-tile
 * (rgb_part is [cout/block_n][m_b][3][rgb_h][rgb_w_px]); b200ir_rgb_combine adds bias, the partials and the up-sampled skip.
 */
typedef struct {
  const void* ptr; /* fp16, element (b,y,x,c) at ptr[b*stride_b + y*stride_h + x*stride_w + c] */
  int32_t c, w, h, b;
  int64_t stride_w, stride_h, stride_b; /* in elements; multiples of 8 */
} b200ir_view;

typedef struct {
  b200ir_view a[B200IR_MAX_VIEWS];
  int32_t num_views;
  const void* weight; /* fp16 [cout][num_taps*cin] */
  int32_t cin, cout;
  int32_t num_taps;
  int8_t tap_view[B200IR_MAX_TAPS], tap_dx[B200IR_MAX_TAPS], tap_dy[B200IR_MAX_TAPS];
  int32_t m_w, m_h, m_b;
  int32_t tile_w, tile_h, tile_b;
  int32_t block_n; /* 16..256, multiple of 16, divides cout */
  /* output */
  void* out;
  int32_t out_fp32; /* 0: fp16, 1: fp32 */
  int64_t out_stride_x, out_stride_y, out_stride_b; /* elements */
  int32_t out_c_off;
  int32_t out_x_mul, out_x_off, out_y_mul, out_y_off;
  /* epilogue */
  const float* bias;       /* [cout] or NULL */
  const float* demod;      /* [m_b][cout] or NULL */
  const float* noise;      /* fp32 plane(s) indexed [b*noise_stride_b + yo*noise_stride_y + xo] or NULL */
  const float* noise_gain; /* device scalar (StyleConv.weight) */
  int64_t noise_stride_b, noise_stride_y;
  int32_t act;      /* 0 none, 1 leaky-relu(0.2)*sqrt(2) (FusedLeakyReLU), 2 max(v, act_slope*v) */
  int32_t res_mode; /* 0 none, 1 same resolution, 2 bilinear x2 of a half-resolution tensor */
  const void* res;  /* fp16 NHWC */
  int64_t res_stride_x, res_stride_y, res_stride_b;
  int32_t res_w, res_h; /* extents of the residual tensor (for the clamp of mode 2) */
  float res_scale;
  int32_t max_ctas; /* 0: one CTA per SM */
  int32_t row_mode; /* 0: generic tiles; 1: allow the row-sliding variant (3x3 stride 1, tile 128x1x1, cout 16/32/64,
                       weights resident in shared memory; picked only when eligible and the batch gives enough work
                       items); 2: take it whenever eligible, whatever the batch */
  const float* out_scale; /* [m_b][cout] or NULL */
  const float* rgb_w;     /* [m_b][3][cout] or NULL */
  float* rgb_part;        /* [cout/block_n][m_b][3][rgb_h][rgb_w_px] fp32, required with rgb_w */
  int32_t rgb_w_px, rgb_h;
  int32_t no_store;       /* 1: do not write `out` (the tensor is only consumed by the fused ToRGB) */
  /* plain nn.Conv2d networks (options/*.yml archs: MSRResNet / EDSR / RCAN, basicsr/archs/arch_util.py):
   * act == 2: v = max(v, act_slope * v) without the sqrt(2) gain (nn.ReLU: slope 0, nn.LeakyReLU(0.1): slope 0.1);
   * res_mul != 0: the residual merge is v * res_scale + res * res_mul (ResidualBlockNoBN: identity + out * res_scale,
   *   arch_util.py:90-93) instead of (v + res) * res_scale;
   * ps_r = 2 or 3: nn.PixelShuffle(ps_r) fused into the store (arch_util.py:96-109, srresnet_arch.py:60-64): the
   *   weight rows are packed in (dy, dx, c) order with block_n = cout / ps_r^2, N-tile t = dy * ps_r + dx writes its
   *   block_n channels to pixel (y * ps_r + dy, x * ps_r + dx); out strides describe the up-sampled tensor. */
  float act_slope;
  float res_mul;
  int32_t ps_r;
  /* Sub-pixel phases as GEMM columns (the stride-2 transposed conv of an upsampling StyleConv as ONE conv:
   * column n belongs to phase n / ps_c (0: ps_c = block_n, the plain pixel shuffle above) and channel n % ps_c;
   * demod_c != 0: the demod table has demod_c channels per image and is indexed with n % demod_c;
   * use_tap_mask: N-tile t executes only the taps in tap_mask[t] (its weight block for the others is all zero). */
  int32_t ps_c, demod_c, use_tap_mask;
  uint32_t tap_mask[8];
  /* ConvUpLayer folded (bilinear x2 + 3x3 conv as one conv over the low-resolution, replicate-padded input; phases =
   * column blocks, ps_r = 2, res_mode = 2): surplus of the folded form on the outermost output ring, subtracted from
   * the accumulator before bias / activation.  fp32, [m_b][2*m_w][ps_c] (top / bottom rows) and [m_b][2*m_h][ps_c]
   * (left / right columns), corner terms already folded into top / bottom.  All four or none. */
  const float* corr_top;
  const float* corr_bot;
  const float* corr_left;
  const float* corr_right;
  /* w_per_image != 0: every image has its own weight matrix -- weight is fp16 [m_b][cout][num_taps*cin] and image b contracts
   * with block b (tile_b must be 1).  The style term of the perceptual loss sends dL/dGram_b back through F_b this way
   * (a 1x1 conv whose matrix differs per image, losses.py:330-356). */
  int32_t w_per_image;
} b200ir_conv_desc;

int b200ir_conv_igemm(const b200ir_conv_desc* d, void* stream);

/* Launch plans (one-time *_plan_create / destroy for the TMA descriptors): everything
 * b200ir_conv_igemm derives from the descriptor on every call -- validation, up to five cuTensorMapEncodeTiled calls,
 * tile / ring sizing, epilogue-profile selection -- is done once and kept in an opaque host object; plan_launch only
 * enqueues the kernel.  A plan is bound to the pointers, shapes and device of its descriptor (the tensor maps hold the
 * base addresses): re-create it when a buffer moves.  Launches of one plan may be issued from any stream, including under
 * CUDA-graph capture.  destroy(NULL) is a no-op. */
typedef struct b200ir_conv_plan b200ir_conv_plan;
int b200ir_conv_plan_create(const b200ir_conv_desc* d, b200ir_conv_plan** plan);
int b200ir_conv_plan_launch(const b200ir_conv_plan* plan, void* stream);
void b200ir_conv_plan_destroy(b200ir_conv_plan* plan);

/* Weight packing (SURVEY.md 8(b) `pack_weights`): state_dict layout fp32 [cout][cin][kh][kw] (EqualConv2d.weight,
 * stylegan2_ocr_arch.py:629-637; nn.Conv2d.weight) -> the fp16 GEMM operand of b200ir_conv_igemm, times `scale` (the
 * equalised-lr factor 1 / sqrt(cin * kh * kw), stylegan2_ocr_arch.py:631, :640):
 *   mode 0 (forward):        out[co][(i*kw + j)*cin + ci]                       = w[co][ci][i][j] * scale   ([cout][taps*cin])
 *   mode 1 (input gradient): out[ci][((kh-1-i)*kw + (kw-1-j))*cout + co]        = w[co][ci][i][j] * scale   ([cin][taps*cout]:
 *            the stride-1 'same' conv is its own adjoint up to this flip / transpose, so dgrad runs on conv_igemm)
 * cin_pad >= cin (mode 0) pads the input-channel block with zeros up to cin_pad (0 = cin). */
int b200ir_pack_weights(const float* w, void* out, int cout, int cin, int kh, int kw, float scale, int mode, int cin_pad,
                        void* stream);

/* ------------------------------------------------------------------------------------------------
 * Memory-bound stages (128-bit NHWC access).  FIR = outer([1,3,3,1])/64 (make_resample_kernel,
 * stylegan2_ocr_arch.py:26-40) evaluated as upfirdn2d does (upfirdn2d.py:162-192; zero padding).
 */

/* conv_body_first: ConvLayer(3, cout, 1) = EqualConv2d 1x1 + FusedLeakyReLU (stylegan2_ocr_arch.py:658-705).
 * x: fp32 NCHW [B][3][H][W]; w: fp32 [cout][3] (already scaled by 1/sqrt(3)); bias fp32 [cout]; out NHWC fp16. */
int b200ir_first_conv(const float* x, const float* w, const float* bias, void* out, int B, int H, int W, int cout,
                      void* stream);

/* Weight gradient of a 3x3 stride-1 'same' convolution (F.conv2d(x, W, padding=1), the EqualConv2d / ConvUpLayer convs:
 * stylegan2_ocr_arch.py:639-648, gfpganv1_ocr_arch.py:192): dW[co][kh][kw][ci] = sum_{b,y,x} dy[b][y][x][co] *
 * x[b][y+kh-1][x+kw-1][ci].  x NHWC fp16 [B][H][W][cin], dy NHWC fp16 [B][H][W][cout], dw fp32 [cout][9][cin] (the
 * layout of the packed forward weights; the equalised-lr scale is the caller's), overwritten.  cin % 16 == 0,
 * cout % 8 == 0 (channel blocks are 128 x 64/128; ragged blocks are zero-filled by TMA).  tcgen05 GEMM over pixels with MN-major operands, split over pixel ranges, fp32 atomic reduction. */
int b200ir_conv_wgrad(const void* x, const void* dy, float* dw, int B, int H, int W, int cin, int cout, void* stream);

/* The same kernel over a strided view of x and a subset of the nine taps:
 *   dW[co][kh][kw][ci] = sum_{b,y,x} dy[b][y][x][co] * xv[b][y+kh-1][x+kw-1][ci]   for the taps with bit kh*3+kw of tap_mask
 * set (zero outside the view's extent, which may differ from H x W; taps outside the mask are left zero).  This covers
 *   - 1x1 convs (EqualConv2d k = 1: ResBlock / ResUpBlock skips, condition heads):  tap_mask = 1 << 4;
 *   - the stride-2 3x3 convs of ResBlock.conv2 (stylegan2_ocr_arch.py:685-697, F.conv2d(stride=2) over the FIR-smoothed
 *     buffer p): one call per pixel phase (ry, rx) of p, the view's strides stepping two pixels / two rows from
 *     p + (ry * row + rx) * cin; kernel element (2*sy + ry, 2*sx + rx) of the conv comes back at tap (sy + 1, sx + 1).
 * xv->c = cin (multiple of 16), xv->b = B, cout % 8 == 0. */
int b200ir_conv_wgrad_view(const b200ir_view* x, const void* dy, float* dw, int B, int H, int W, int cout,
                           uint32_t tap_mask, void* stream);

/* Backward of FusedLeakyReLU fused with the bias gradient (basicsr/ops/fused_act/fused_act.py:30-63; device code
 * fused_bias_act_kernel.cu:20-50 with act = 3, grad = 1, and the grad_input.sum over batch and pixels of
 * FusedLeakyReLUFunctionBackward.forward):  dz = dy * scale * (y > 0 ? 1 : slope),  dbias[c] = sum_p dz[p][c].
 * dy, y (the saved forward output, the reference's `out`), dz: NHWC fp16 [n_pix][C]; dbias fp32 [C], overwritten, may be
 * NULL (ScaledLeakyReLU, stylegan2_ocr_arch.py:604-606).  C % 8 == 0.  dz may alias dy.  y == NULL means
 * "no activation" (every element takes the scale branch) and dz == NULL skips the store: with scale = 1 that is the plain
 * bias gradient of a conv without activation (dbias[c] = sum_p dy[p][c]). */
int b200ir_lrelu_bias_bwd(const void* dy, const void* y, void* dz, float* dbias, int64_t n_pix, int C, float slope,
                          float scale, void* stream);

/* Adjoint of b200ir_fir_pad22 (upfirdn2d(x, k, pad = (2, 2)), upfirdn2d.py:153-192: the blur in front of the stride-2 3x3
 * conv of ResBlock.conv2) = upfirdn2d(d, k, pad = (1, 1)): in [B][in_h][in_w][C] fp16 with the valid (H+1) x (W+1) region
 * at the origin (the raw output of the transposed conv that is the stride-2 conv's input gradient) -> out [B][H][W][C].
 * TMA streaming kernel; C % 32 == 0. */
int b200ir_fir_pad11(const void* in, void* out, int B, int H, int W, int C, int in_h, int in_w, void* stream);

/* Adjoint of b200ir_fir_down2 (FIR pad (1, 1) + even-position sampling in front of the 1x1 skip conv of ResBlock):
 * d [B][h][w][C] fp16 -> out [B][2h][2w][C] = add + upfirdn2d(d, k, up = 2, pad = (2, 1)) (k normalised to sum 1, no
 * up-sampling gain); add may be NULL or alias out. */
int b200ir_fir_down2_adjoint(const void* d, const void* add, void* out, int B, int h, int w, int C, void* stream);

/* Adjoint of b200ir_bilinear_up2 (F.interpolate x2 bilinear, align_corners=False, of ConvUpLayer, gfpganv1_ocr_arch.py:190):
 * d [B][2h][2w][C] fp16 -> out [B][h][w][C] = scale * sum of the 4 x 4 taps (.25, .75, .75, .25)^2 at rows / columns
 * clamp(2i - 1 + t, 0, 2n - 1). */
int b200ir_bilinear_up2_adjoint(const void* d, void* out, int B, int h, int w, int C, float scale, void* stream);

/* Weight gradient of b200ir_first_conv (conv_body_first, the 1x1 EqualConv2d over the 3-channel input image): x fp32 NCHW
 * [B][3][H][W], dz NHWC fp16 [B][H][W][cout] (after b200ir_lrelu_bias_bwd) -> dw fp32 [cout][3], overwritten (gradient
 * w.r.t. the scaled weights the forward entry point takes).  cout = 8 * a divisor of 256. */
int b200ir_first_conv_wgrad(const float* x, const void* dz, float* dw, int B, int H, int W, int cout, void* stream);

/* Backward of b200ir_minibatch_stddev (stylegan2_arch.py:791-801): x NHWC fp16 [B][P][C] (the layer's input), dcat NHWC
 * fp16 [B][P][c_pad] (gradient of its output: C feature channels + the statistic channel + padding), ds fp32 [B / group]
 * (ds[m] = sum over the group's samples and pixels of dcat[.., C]) -> dx NHWC fp16 [B][P][C].  C % 8 == 0, group <= 8. */
int b200ir_minibatch_stddev_bwd(const void* x, const void* dcat, const float* ds, void* dx, int B, int P, int C, int c_pad,
                                int group, void* stream);

/* One torch.optim.Adam step (no amsgrad; optimizer_g / optimizer_d of basicsr/models/gfpgan_model.py:217-248) over a flat
 * fp32 parameter buffer, fused with the gradient scaling of the data-parallel average (grad_scale = 1 / world) and,
 * when ema != NULL, with the EMA update of BaseModel.model_ema (basicsr/models/base_model.py:50-57):
 *   g = grad * grad_scale + weight_decay * p;  m = b1 m + (1 - b1) g;  v = b2 v + (1 - b2) g^2
 *   p -= lr / (1 - b1^step) * m / (sqrt(v) / sqrt(1 - b2^step) + eps);  ema = decay * ema + (1 - decay) * p
 * All buffers n fp32 elements, 16-byte aligned; step counts from 1.  A non-finite g (an overflowed fp16 activation
 * gradient under a static loss scale) is treated as 0 for that element instead of turning p, m and v into NaN. */
int b200ir_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr, float beta1,
                     float beta2, float eps, float weight_decay, int step, float grad_scale, float* ema, float ema_decay,
                     void* stream);

/* ------------------------------------------------------------------------------------------------
 * Training step (GFPGANModel.optimize_parameters, basicsr/models/gfpgan_model.py:494-691): adjoints of the frozen StyleGAN2
 * decoder's pointwise stages and the losses.  With fix_decoder=True the decoder has no weight gradients; l_g needs
 * d(image)/d(style_code) and d(image)/d(conditions).  Per ModulatedConv2d (stylegan2_ocr_arch.py:239-279), with
 * x' = x * s[b,ci], y = conv(x', W/sqrt(cin k^2)) * d[b,co], d = rsqrt(scale2 * sum_ci s^2 wsq[co,ci] + 1e-8):
 *   dx' = dgrad(dy * d)  (b200ir_conv_igemm with adjoint weights),  dx = dx' * s,
 *   ds[b,ci] = sum_p x * dx'  -  scale2 * s[b,ci] * sum_co dd[b,co] * d[b,co]^2 * wsq[co,ci],   dd[b,co] = sum_p dy * y.
 * Gradients of activations are NHWC fp16 and carry the caller's loss scale; tables are fp32. */

/* SFT (gfpganv1_ocr_arch.py:118-125) + modulation of the next conv (stylegan2_ocr_arch.py:247-251), training forward
 * (the inference path fuses both into b200ir_upfir_act; training keeps the pre-SFT tensor `a` for the backward pass):
 *   out[b,p,c] = (c < C - c_sft ? a : a * scale[b,p,c-(C-c_sft)] + shift[...]) * (s_next ? s_next[b*C+c] : 1)
 * a, out NHWC fp16 [B][P][C]; scale / shift NHWC fp16 [B][P][c_sft] or both NULL. */
int b200ir_sft_mod(const void* a, const void* scale, const void* shift, int c_sft, const float* s_next, void* out, int B,
                   int64_t P, int C, void* stream);
/* Backward of b200ir_sft_mod and the direct style gradient of the conv that consumed `out`: g NHWC fp16 [B][P][C] =
 * gradient w.r.t. out (the dgrad GEMM's result); with o the SFT output recomputed from a, scale, shift:
 *   ds[b][c] += sum_p g * o   (fp32 [B][C], the caller zeroes it; NULL = skip);   do = g * s_next;
 *   da = do on the plain channels, do * scale on the SFT channels (accumulate != 0: added to the existing da; NULL = skip);
 *   dscale = do * a,  dshift = do   (NHWC fp16 [B][P][c_sft]).
 * a_stride_b: batch stride of `a` in elements (0: one [P][C] tensor for all images -- ConstantInput). */
int b200ir_sft_mod_bwd(const void* g, const void* a, int64_t a_stride_b, const void* scale, const void* shift, int c_sft,
                       const float* s_next, void* da, int accumulate, void* dscale, void* dshift, float* ds, int B, int64_t P,
                       int C, void* stream);
/* Backward of the StyleConv tail (noise injection + FusedLeakyReLU, stylegan2_ocr_arch.py:323-333) with the demodulation
 * reduction: a = saved output = lrelu(y + gain * noise + bias) * sqrt 2;  dz = da * sqrt 2 * (a > 0 ? 1 : 0.2);  y is
 * reconstructed from a;  dd[b][c] += sum_p dz * y (fp32 [B][C], caller zeroes; NULL = skip);
 * out = dz * (oscale ? oscale[b*C+c] : 1) * mul   (the demodulation d, and the gain 4 of the up-sampling FIR). out may alias da. */
int b200ir_style_act_bwd(const void* da, const void* a, const float* noise, int64_t noise_stride_b, const float* noise_gain,
                         const float* bias, const float* oscale, float mul, void* out, float* dd, int B, int64_t P, int C,
                         void* stream);
/* Backward of ToRGB's modulated 1x1 conv (stylegan2_ocr_arch.py:357-374; no demodulation): drgb fp32 NCHW [B][3][P],
 * a NHWC fp16 [B][P][C] (its input), w fp32 [3][C] pre-scaled by 1/sqrt(C), s fp32 [B][C] or NULL:
 *   t = sum_o drgb[b][o][p] * w[o][c];   da (+)= s * t;   ds[b][c] += sum_p a * t. */
int b200ir_to_rgb_bwd(const float* drgb, const void* a, const float* w, const float* s, void* da, int accumulate, float* ds,
                      int B, int64_t P, int C, void* stream);
/* fix_decoder = false (the shipped training YAMLs train the StyleGAN2 decoder too): the same two kernels with the extra
 * per-image reductions the decoder's own parameters need --
 *   style_act_bwd_params: db[b][c] += sum_p dz (StyleConv.activate.bias), dn[b][c] += sum_p dz * noise (StyleConv.weight, the
 *     noise gain; stylegan2_ocr_arch.py:316-333), fp32 [B][C], caller zeroes;
 *   to_rgb_bwd_params: R[b][o][c] += sum_p drgb[b][o][p] * a[b][p][c] (fp32 [B][3][C], caller zeroes): the 1x1 weight gradient
 *     before its modulation, dw[o][c] = sum_b s[b][c] R[b][o][c] / sqrt(C);
 *   plane_sums: out[c] += sum_{b,p} x[b][c][p] over fp32 NCHW planes (ToRGB.bias). */
int b200ir_style_act_bwd_params(const void* da, const void* a, const float* noise, int64_t noise_stride_b,
                                const float* noise_gain, const float* bias, const float* oscale, float mul, void* out, float* dd,
                                float* db, float* dn, int B, int64_t P, int C, void* stream);
int b200ir_to_rgb_bwd_params(const float* drgb, const void* a, const float* w, const float* s, void* da, int accumulate, float* ds,
                             float* R, int B, int64_t P, int C, void* stream);
int b200ir_plane_sums(const float* x, float* out, int B, int Cn, int64_t P, void* stream);
/* Folding the per-image tables over the batch into parameter gradients (fix_decoder = false):
 *   table_colsum: out[j] = scale * sum_b in[b][j] * (mul ? mul[b][j % m] : 1), in fp32 or fp16 [B][n] (in_f16), mul fp32 [B][m]
 *     -- activate.bias / modulation.bias / noise gain (plain sums), ToRGB weight (R weighted by s), ConstantInput (the gradient
 *     of the modulated constant, NHWC fp16, weighted by s; stylegan2_ocr_arch.py:287-301);
 *   mod_linear_wgrad: dw[ci][f] = wscale * sum_b ds[b][ci] * latent[b][lat_idx][f] (modulation EqualLinear weight, :233-234);
 *   modconv_wgrad: dw[co][ci][k] = scale * G - scale^2 * W[co][ci][k] * sum_b dd[b][co] d[b][co]^2 s[b][ci]^2 in the reference
 *     layout [cout][cin][taps]; G = b200ir_conv_wgrad's tap-major result [cout][taps][cin] on (modulated input, dy * d), or
 *     [cin][taps][cout] when `transposed` (the up-sampling conv: its GEMM runs with input and output roles exchanged). The second
 *     term is the weight's path through the demodulation table (:253-257). */
/* Weight gradient of a low-channel 3x3 conv computed on pixel-folded views: b200ir_conv_wgrad is called with x, dy read as
 * [B][H][W/f][f*C] (f horizontally adjacent pixels = f*C channels) and returns G fp32 [f*cout][9][f*cin]; this folds G back to
 * dw fp32 [cout][9][cin]: tap (kh, kw) = sum of the blocks (s_o, s_i, kw') with f * (kw' - 1) + s_i - s_o = kw - 1. */
int b200ir_wgrad_unfold(const float* G, float* dw, int f, int cin, int cout, void* stream);
int b200ir_table_colsum(const void* in, int in_f16, const float* mul, int m, float scale, float* out, int B, int64_t n,
                        void* stream);
int b200ir_mod_linear_wgrad(const float* ds, const float* latent, float wscale, float* dw, int L, int F, int lat_idx, int B,
                            int cin, void* stream);
int b200ir_modconv_wgrad(const float* G, int transposed, const float* W, const float* s, const float* dd, const float* d,
                         float scale, float* dw, int B, int cin, int cout, int taps, void* stream);
/* Adjoint of UpFirDnUpsample on the RGB skip (upfirdn2d(skip, FIR*4, up=2, pad=(2,1)), stylegan2_ocr_arch.py:43-69):
 * d fp32 [planes][2h][2w] -> out fp32 [planes][h][w]. */
int b200ir_rgb_up_adjoint(const float* d, float* out, int planes, int h, int w, void* stream);
/* ds[b][ci] -= scale2 * s[b][ci] * sum_co dd[b][co] * d[b][co]^2 * wsq[co][ci]  (the style gradient through the
 * demodulation table, stylegan2_ocr_arch.py:253-257); in place on ds. */
int b200ir_demod_bwd(float* ds, const float* s, const float* dd, const float* d, const float* wsq, float scale2, int B, int cin,
                     int cout, void* stream);
/* Modulation EqualLinear w.r.t. the latent (stylegan2_ocr_arch.py:229-234): dlat[b][lat_idx][f] += wscale * sum_ci ds[b][ci] *
 * w[ci][f];  dlat fp32 [B][L][F]. */
int b200ir_mod_linear_bwd(const float* ds, const float* w, float wscale, float* dlat, int L, int F, int lat_idx, int B, int cin,
                          void* stream);
/* Input gradient of b200ir_first_conv (the discriminator's conv_body.0 passes d(score)/d(image) back to net_g,
 * gfpgan_model.py:549-552): dz NHWC fp16 [B][H][W][cout] (after b200ir_lrelu_bias_bwd), w fp32 [cout][3] (scaled) ->
 * dx fp32 NCHW [B][3][H][W] (accumulate != 0: added). */
int b200ir_first_conv_dgrad(const void* dz, const float* w, float* dx, int accumulate, int B, int H, int W, int cout,
                            void* stream);
/* The U-Net's toRGB heads (gfpganv1_ocr_arch.py:377-378) are computed with the three channels padded to cpad for the GEMM
 * kernels: head NHWC fp16 [B][P][cpad] -> rgb fp32 NCHW [B][3][P] (what the reference returns in out_rgbs), and the adjoint. */
int b200ir_head_to_nchw(const void* head, float* rgb, int B, int64_t P, int cpad, void* stream);
int b200ir_nchw_to_head(const float* drgb, void* dhead, int B, int64_t P, int cpad, void* stream);
/* Per-image Gram-type reduction on the weight-gradient GEMM (b200ir_conv_wgrad_view's kernel without the sum over the batch):
 * out[b][co][ci] = sum_{y,x} dy[b][y][x][co] * x[b][y][x][ci];  x NHWC fp16 [B][H][W][cin], dy [B][H][W][cout], out fp32
 * [B][cout][cin], overwritten.  With dy = x this is the un-normalised Gram matrix of PerceptualLoss._gram_mat
 * (losses.py:343-356), all images in one launch.  cin % 16 == 0, cout % 8 == 0. */
int b200ir_gram_batched(const void* x, const void* dy, float* out, int B, int H, int W, int cin, int cout, void* stream);

/* Perceptual loss (PerceptualLoss, basicsr/losses/losses.py:250-356 over VGGFeatureExtractor, basicsr/archs/vgg_arch.py:56-160):
 * the VGG19 convs run on b200ir_conv_igemm (nn.Conv2d 3x3 + bias + ReLU = act 2, slope 0), their input gradients on the same
 * kernel with adjoint weights, the Gram matrices of the style term on b200ir_conv_wgrad_view; these are the remaining stages.
 * maxpool2_relu: nn.ReLU + nn.MaxPool2d(2, 2) on a pre-activation tensor z NHWC fp16 [B][H][W][C] -> [B][H/2][W/2][C]
 *   (the tapped layers are read BEFORE their ReLU, so the conv stores z and this kernel applies both);
 * maxpool2_relu_bwd: dz = add + dpool routed to the first maximum of each window when it is positive (add: the loss
 *   gradient at the tapped layer, may be NULL; dpool may be NULL for the top layer);
 * l1_loss_f16: b200ir_l1_loss on fp16 tensors with an fp16 gradient (n % 8 == 0). */
int b200ir_maxpool2_relu(const void* z, void* out, int B, int H, int W, int C, void* stream);
int b200ir_maxpool2_relu_bwd(const void* z, const void* dpool, const void* add, void* dz, int B, int H, int W, int C,
                             void* stream);
int b200ir_l1_loss_f16(const void* x, const void* t, int64_t n, float weight, float grad_scale, float* loss, void* grad,
                       void* stream);

/* R1 penalty of the discriminator (r1_penalty, basicsr/losses/losses.py:492-506; gfpgan_model.py:683-689): the double
 * backward is assembled from the primal backward signals and a forward-mode (tangent) pass along v = grad_x sum D(x), see
 * image_restoration_b200/r1.py.  These are its three non-GEMM pieces:
 * sum_squares: out[0] += scale * sum_i x[i]^2  (fp32; the penalty's value);
 * minibatch_stddev_jvp: tangent of b200ir_minibatch_stddev's output along t (NHWC fp16 [B][P][C], the tangent of its input x):
 *   tcat [B][P][c_pad] = [ t | ts[m] | 0 ],  ts[m] = 1/(C P) sum_{p,c} sum_g (x_g - mu) t_g / (G sigma)   (ts: fp32 [B/group] work);
 * minibatch_stddev_hvp: q = Hessian of (a . s)(x) times t, a fp32 [B/group] = d f / d s[m]:
 *   q[b][p][c] = a[m]/(C P) * ((t_g - tbar)/(G sigma) - (x_g - mu) * dot/(G^2 sigma^3)),  dot = sum_g (x_g - mu) t_g. */
int b200ir_sum_squares(const float* x, int64_t n, float scale, float* out, void* stream);
int b200ir_minibatch_stddev_jvp(const void* x, const void* t, float* ts, void* tcat, int B, int P, int C, int c_pad, int group,
                                void* stream);
int b200ir_minibatch_stddev_hvp(const void* x, const void* t, const float* a, void* q, int B, int P, int C, int group,
                                void* stream);

/* L1Loss(loss_weight, reduction='mean') (basicsr/losses/losses.py:81-106; l_g_pix gfpgan_model.py:521, pyramid :532-536)
 * forward + gradient in one pass: loss[0] += weight / n * sum |x - t|;  grad[i] = grad_scale * weight / n * sign(x - t).
 * fp32 arrays; loss is a device scalar the caller zeroes; grad may be NULL. */
int b200ir_l1_loss(const float* x, const float* t, int64_t n, float weight, float grad_scale, float* loss, float* grad,
                   void* stream);
/* GANLoss('wgan_softplus') (losses.py:404-419, 438-470): loss[0] += weight / n * sum softplus(sign * pred) with sign = -1
 * for target_is_real;  dpred = grad_scale * weight / n * sign * sigmoid(sign * pred).  pred / dpred fp16, element i at
 * [i * stride]; dpred may be NULL. */
int b200ir_softplus_loss(const void* pred, int n, int stride, float sign, float weight, float grad_scale, float* loss,
                         void* dpred, void* stream);

/* Minibatch standard deviation of StyleGAN2Discriminator.forward (basicsr/archs/stylegan2_arch.py:791-801), stddev_feat = 1:
 * x NHWC fp16 [B][P][C]; group = min(B, stddev_group) must divide B; s fp32 [B / group] (work buffer);
 * out NHWC fp16 [B][P][c_pad] = concat(x, stddev channel, zero padding up to c_pad): the input of final_conv. */
int b200ir_minibatch_stddev(const void* x, float* s, void* out, int B, int P, int C, int c_pad, int group, void* stream);

/* Pre / post of the serving scripts (api.py:96-105, inference.py:61-71):
 * u8_to_input  = img2tensor(img / 255., bgr2rgb=True, float32=True) + normalize(mean .5, std .5)
 *                (basicsr/utils/img_util.py:9-35): uint8 HWC [B][H][W][3] -> fp32 NCHW [B][3][H][W] in [-1, 1];
 * image_to_u8  = tensor2img(out, rgb2bgr=True, min_max=(-1, 1)) (img_util.py:38-94): clamp, (x+1)/2, *255, round half
 *                to even, uint8 HWC.  swap_rb != 0 reverses the channel order (BGR images, as cv2 delivers them). */
int b200ir_u8_to_input(const uint8_t* img, float* x, int B, int H, int W, int swap_rb, void* stream);
/* the same for a float image in [0, 1] (the GT side of FFHQDegradationDataset.__getitem__,
 * ffhq_degradation_dataset.py:288, :310): fp32 HWC [B][H][W][3] -> fp32 NCHW, (x - 0.5) / 0.5 */
int b200ir_f32_to_input(const float* img, float* x, int B, int H, int W, int swap_rb, void* stream);
int b200ir_image_to_u8(const float* x, uint8_t* img, int B, int H, int W, int swap_rb, void* stream);

/* Helpers of the folded ConvUpLayer (b200ir_conv_desc.corr_*): replicate_border fills the one-pixel ring of an NHWC fp16
 * buffer [B][h+2][w+2][C] from its interior (the clamp of F.interpolate(..., 'bilinear'), gfpganv1_ocr_arch.py:190);
 * upfold_corners removes the doubly counted corner taps from the top / bottom correction rows (wc fp32 [4][cout][cin] =
 * W[:, :, 0, 0], W[:, :, 0, 2], W[:, :, 2, 0], W[:, :, 2, 2] of the 3x3 conv; top / bot fp32 [B][2w][cout]). */
int b200ir_replicate_border(void* t, int B, int h, int w, int C, void* stream);
int b200ir_upfold_corners(const void* tp, const float* wc, float* top, float* bot, int B, int h, int w, int cin, int cout,
                          void* stream);

/* Tiled full-frame inference (BASELINE config 4; the reference has no tiling code -- api_plate_oto.py:376-401 resizes
 * the whole image -- so the contract is this header): frame fp32 [C][H][W]; tiles fp32 [nty*ntx][C][T][T] at rows ty[]
 * and columns tx[] (device int32 arrays, ascending, first 0, last H-T / W-T).  Blend weights are separable linear
 * ramps min(1, (d+1)/(overlap+1)) with d the distance to a tile edge that is not a frame edge; the result is the
 * weight-normalised sum. */
int b200ir_tiles_gather(const float* frame, float* tiles, int C, int H, int W, int T, const int32_t* ty,
                        const int32_t* tx, int nty, int ntx, void* stream);
int b200ir_tiles_blend(const float* tiles, float* frame, int C, int H, int W, int T, int overlap, const int32_t* ty,
                       const int32_t* tx, int nty, int ntx, void* stream);

/* UpFirDnSmooth before a stride-2 3x3 conv (pad (2,2), ConvLayer downsample=True, k=3):
 * in [B][H][W][C] -> out rows/cols 0..H / 0..W of a [B][out_h][out_w][C] buffer (out_h >= H+1, out_w >= W+1). */
int b200ir_fir_pad22(const void* in, void* out, int B, int H, int W, int C, int out_h, int out_w, void* stream);

/* UpFirDnSmooth (pad (1,1)) followed by the stride-2 sampling of the 1x1 skip conv (ResBlock.skip):
 * in [B][H][W][C] -> out [B][H/2][W/2][C]. */
int b200ir_fir_down2(const void* in, void* out, int B, int H, int W, int C, void* stream);

/* F.interpolate(scale_factor=2, mode='bilinear', align_corners=False) (gfpganv1_ocr_arch.py:190). */
int b200ir_bilinear_up2(const void* in, void* out, int B, int h, int w, int C, void* stream);

/* out = a + b (fp16, n elements, n % 8 == 0): the U-Net skip add, gfpganv1_ocr_arch.py:368. */
int b200ir_add(const void* a, const void* b, void* out, int64_t n, void* stream);

/* Tail of an upsampling StyleConv (stylegan2_ocr_arch.py:261-267,323-333) + SFT (gfpganv1_ocr_arch.py:118-125):
 * raw: demodulated transposed-conv output, [B][raw_h][raw_w][C] buffer holding (2h+1)x(2w+1) valid samples;
 * y = FIR*4 with pad (1,1) -> [B][2h][2w][C]; y += noise_gain*noise + bias; y = lrelu(y)*sqrt2;
 * channels c >= C - c_sft: y = y*scale[..., c-(C-c_sft)] + shift[...]  (scale/shift NHWC fp16 with c_sft channels);
 * if s_next: y *= s_next[b*C + c]   (modulation of the next conv, stylegan2_ocr_arch.py:247-251). */
int b200ir_upfir_act(const void* raw, void* out, int B, int h2, int w2, int C, int raw_h, int raw_w, const float* noise,
                     int64_t noise_stride_b, const float* noise_gain, const float* bias, const void* scale,
                     const void* shift, int c_sft, const float* s_next, void* stream);

/* ToRGB.forward (stylegan2_ocr_arch.py:357-374): rgb[b,o,y,x] = sum_c w[o][c]*s[b][c]*x[b,y,x,c] + bias[o]
 *   (+ upfirdn2d(skip, FIR*4, up=2, pad=(2,1)) when skip != NULL, skip is fp32 NCHW [B][3][h/2][w/2]).
 * w: fp32 [3][C] pre-scaled by 1/sqrt(C); s: fp32 [B][C] or NULL (plain EqualConv2d toRGB, gfpganv1_ocr_arch.py:290-292).
 * rgb: fp32 NCHW [B][3][h][w].  If xs_out: xs_out = x * s_next[b][c] (NHWC fp16; input of the next modulated conv). */
int b200ir_to_rgb(const void* x, int B, int h, int w, int C, const float* wrgb, const float* s, const float* bias,
                  const float* skip, float* rgb, const float* s_next, void* xs_out, void* stream);

/* Second half of the fused ToRGB (see b200ir_conv_igemm): rgb[b][o][y][x] = bias[o] + sum_t part[t][b][o][y][x]
 *   (+ upfirdn2d(skip, FIR*4, up=2, pad=(2,1)) when skip != NULL; skip fp32 NCHW [B][3][h/2][w/2]).
 * rgb_wmod builds the per-image weights the conv epilogue uses: wm[b][o][c] = w[o][c] * s[b][c]. */
int b200ir_rgb_combine(const float* part, int n_parts, const float* bias, const float* skip, float* rgb, int B, int h,
                       int w, void* stream);
int b200ir_rgb_wmod(const float* w, const float* s, float* wm, int B, int C, void* stream);

/* ConstantInput.forward (stylegan2_ocr_arch.py:389-391) fused with the first modulation:
 * out[b,p,c] = cst[p*C + c] * s[b*C + c];  cst NHWC fp16 [P][C]. */
int b200ir_modulate_const(const void* cst, const float* s, void* out, int B, int P, int C, void* stream);

/* EqualLinear modulation (stylegan2_ocr_arch.py:247, 165-175): s[b][i] = sum_f w[i][f]*latent[b][lat_idx][f]*wscale + bias[i].
 * latent fp32 [B][L][F]. */
int b200ir_mod_linear(const float* latent, int L, int F, int lat_idx, const float* w, const float* bias, float wscale,
                      float* s, int B, int cin, void* stream);

/* Demodulation coefficients (stylegan2_ocr_arch.py:253-257): d[b][o] = rsqrt(scale2 * sum_i s[b][i]^2 * wsq[o][i] + 1e-8),
 * wsq[o][i] = sum_k W[o][i][k]^2 (fp32). */
int b200ir_demod(const float* s, const float* wsq, float scale2, float* d, int B, int cin, int cout, void* stream);

/* All modulation linears / demodulation tables of one forward in a single launch each (same arithmetic as
 * b200ir_mod_linear / b200ir_demod); `layers_dev` is a DEVICE array of n_layers records. */
typedef struct {
  const float* w;    /* [cin][F] */
  const float* bias; /* [cin] */
  float* s;          /* out [B][cin] */
  int32_t lat_idx, cin;
} b200ir_mod_layer;
typedef struct {
  const float* s;   /* [B][cin] */
  const float* wsq; /* [cout][cin] */
  float* d;         /* out [B][cout] */
  float scale2;
  int32_t cin, cout;
} b200ir_demod_layer;
int b200ir_mod_linear_multi(const float* latent, int L, int F, const b200ir_mod_layer* layers_dev, int n_layers,
                            int max_cin, float wscale, int B, void* stream);
int b200ir_demod_multi(const b200ir_demod_layer* layers_dev, int n_layers, int max_cout, int B, void* stream);

/* Style MLP (StyleGAN2OCRGenerator.style_mlp, stylegan2_ocr_arch.py:12-23,424-430; the input_is_latent=False path):
 * out = MLP(NormStyleCode(z)), n_layers x [EqualLinear(F, F, lr_mul) + fused leaky-ReLU]; z, out fp32 [B][F];
 * w fp32 [n_layers][F][F] and bias fp32 [n_layers][F] as stored in the state_dict (not pre-scaled). */
int b200ir_style_mlp(const float* z, const float* w, const float* bias, float* out, int B, int F, int n_layers,
                     float lr_mul, void* stream);

/* NHWC fp16 [B][P][C] -> fp32 matrix [B][P*C] is a reinterpretation; this converts fp32 NCHW image batches to the
 * caller-facing layout when needed: out_nchw[b][c][p] = in_nhwc[b][p][c] (fp16 -> fp32). */
int b200ir_nhwc_to_nchw_f32(const void* in, float* out, int B, int P, int C, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Plain-conv SR networks of options/*.yml (MSRResNet srresnet_arch.py:8-68, EDSR edsr_arch.py:8-72, RCAN
 * rcan_arch.py:8-135): the 3x3 convs, ReLU / LeakyReLU(0.1), residual merges and nn.PixelShuffle run in
 * b200ir_conv_igemm (act == 2, res_mul, ps_r); these are the memory-bound stages around them.
 */

/* fp32 NCHW [B][C][H*s][W*s] -> NHWC fp16 [B][H][W][Cpad], (x - sub[c]) * mul, zero padding channels (edsr_arch.py:62:
 * (x - mean) * img_range; sub may be NULL).  unshuffle = s > 1 fuses pixel_unshuffle (arch_util.py:185-201, RRDBNet
 * scale 2 / 1): output channel c*s*s + dy*s + dx <- x[b][c][y*s+dy][x*s+dx]; H, W are the OUTPUT extents. */
int b200ir_nchw_to_nhwc_pad(const float* x, void* out, int B, int C, int H, int W, int Cpad, const float* sub, float mul,
                            int unshuffle, void* stream);
/* F.interpolate(scale_factor=2, mode='nearest') on NHWC fp16 (rrdbnet_arch.py:118-119). */
int b200ir_nearest_up2(const void* in, void* out, int B, int h, int w, int C, void* stream);
/* conv_last output fp32 NHWC [B][H][W][Cpad] -> fp32 NCHW [B][C][H][W]: y * mul + add[c] (edsr_arch.py:69:
 * x / img_range + mean) + F.interpolate(base, scale_factor=scale, mode='bilinear', align_corners=False)
 * (srresnet_arch.py:66-67; base fp32 NCHW [B][C][H/scale][W/scale] or NULL). */
int b200ir_sr_output(const float* y, float* out, int B, int C, int H, int W, int Cpad, float mul, const float* add,
                     const float* base, int scale, void* stream);
/* RCAN channel attention (rcan_arch.py:8-24): nn.AdaptiveAvgPool2d(1) of an NHWC fp16 tensor -> mean fp32 [B][C];
 * att = sigmoid(W2 relu(W1 mean + b1) + b2) with W1 [Cs][C], W2 [C][Cs]; RCAB tail (rcan_arch.py:43-45)
 * out = x * att[b][c] * res_scale + identity (att NULL = 1: the RRDB merge out * 0.2 + x, rrdbnet_arch.py:59-63;
 * identity / out may be channel slices of wider NHWC buffers: pixel strides in elements). */
int b200ir_channel_mean(const void* x, float* mean, int B, int HW, int C, void* stream);
int b200ir_ca_mlp(const float* mean, const float* w1, const float* b1, const float* w2, const float* b2, float* att,
                  int B, int C, int Cs, void* stream);
int b200ir_ca_scale_add(const void* x, const float* att, const void* identity, void* out, float res_scale, int B, int HW,
                        int C, int64_t id_stride, int64_t out_stride, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Fused degradation (pyblur blur -> cv2.resize down -> Gaussian noise -> clip -> cv2.resize up -> round/clip ->
 * normalize), one launch per batch of crops.  Replaces random_pyblur (degradations.py:363-366; pyblur/*.py
 * convolve2d(mode='same', fillvalue=255).astype(uint8)), the two cv2.resize(INTER_LINEAR) calls and
 * random_add_gaussian_noise + round/clip/normalize of FFHQDegradationDataset.__getitem__
 * (ffhq_degradation_dataset.py:244-272,307-311; degradations.py:660-669).
 *
 * gt:     uint8 [B][H][W][3] (the uint8 image random_pyblur builds: np.array(img*255, dtype=uint8))
 * taps:   fp64 [B][kmax][kmax] blur kernels, zero padded, centred; ksize[b] odd (0 = no blur)
 * taps_f64: int32 [B] (NULL = all 0): 1 = the reference's convolve2d ran in float64 for this crop (box / disk / line
 *         kernels are float64 under NumPy 2), 0 = in float32 (psf kernels; every kernel under the pinned NumPy 1.23).
 *         The blur reproduces scipy's summation tree in that type: the truncated uint8 image is bit-identical.
 * lr_w/lr_h: int32 [B] low-resolution size per crop; noise: fp32 [B][lr_hmax][lr_wmax][3] (already sigma/255-scaled)
 * out:    fp32 NCHW [B][3][H][W], (x-0.5)/0.5-normalised, channel order reversed if bgr2rgb.
 * blur_u8_out / blur_f32_out (optional, parity checks): the full blurred image [B][H][W][3] as pyblur returns it
 *         (uint8, truncated) and the fp32 convolution result before the truncation.
 */
int b200ir_degrade(const uint8_t* gt, const double* taps, const int32_t* ksize, const int32_t* taps_f64, int kmax,
                   const int32_t* lr_w,
                   const int32_t* lr_h, const float* noise, int lr_wmax, int lr_hmax, float* out, uint8_t* blur_u8_out,
                   float* blur_f32_out, int B, int H, int W, int bgr2rgb, void* stream);

/* ------------------------------------------------------------------------------------------------
 * The whole LQ synthesis of FFHQDegradationDataset.__getitem__ (ffhq_degradation_dataset.py:242-311) in one launch,
 * one CTA per crop:  blur (random_mixed_kernels, degradations.py:419-523) -> cv2.resize down (:255-256) ->
 * random_add_gaussian_noise (degradations.py:660-669) -> random_add_jpg_compression (degradations.py:876-910:
 * cv2.imencode / imdecode, i.e. libjpeg-turbo baseline 4:2:0 with the islow DCT; the lossless entropy coding is
 * skipped, the rest is the same integer arithmetic) -> cv2.resize up (:272) -> color_jitter (:90-95) ->
 * cv2.cvtColor(BGR2GRAY) (:283-285) -> color_jitter_pt (torchvision adjust_*; :187-207, :290-296) -> round / clip /
 * normalize (:307-311).
 * Random draws stay on the host (image_restoration_b200.degradation.sample_params mirrors the reference's order of
 * random / np.random calls); this entry point is deterministic in its arguments.
 * random_mask (:153-187, :299-303; `random_mask: true` in the dataset options) is the mask argument of
 * b200ir_degrade_full_ex: the shapes are drawn on the host with the reference's own calls (random rectangles / halves,
 * cv2.line / circle / ellipse), the kernel applies them.
 * Not covered: 'pyblur_motion' / 'random_cover' call RandomMotion / RandomCover, which the reference's pyblur package does
 * not define (pyblur/pyblur/__init__.py) -- they raise NameError in the reference itself.
 */
typedef struct b200ir_degrade_crop {
  int32_t blur_mode;    /* 0 none; 1 'pyblur': scipy convolve2d on the uint8 image, fill 255, truncated to uint8;
                           2 cv2.filter2D on image/255 ('iso', 'aniso', 'motion', 'average'): correlation,
                           BORDER_REFLECT_101, fp32 sum in kernel order (OpenCV itself switches to a DFT for kernels
                           >= 11x11, so its result differs in the last bits);
                           3 'median': cv2.medianBlur(uint8 image, ksize), BORDER_REPLICATE (degradations.py:353-355);
                           4 'bilateral': cv2.bilateralFilter(uint8 image, ksize, sigma, sigma) (degradations.py:358-361);
                             taps = the space weights exp(-r^2 / (2 sigma^2)) inside the radius, zero outside;
                           5 'bicubic': torchvision Resize((h // 4, w // 4), BICUBIC) and back on the PIL image
                             (degradations.py:379-385) = Pillow's 8-bit ImagingResample twice, restated as the same integer
                             arithmetic (needs bicubic_scratch of b200ir_degrade_full_ex; taps unused) */
  int32_t ksize;        /* odd extent of the kernel inside its kmax x kmax block */
  int32_t blur_f64;     /* blur_mode 1 only: 1 = the reference's convolve2d ran in float64 (box / disk / line kernels are
                           float64 under NumPy 2), 0 = in float32 (psf kernels).  The kernel reproduces scipy's summation
                           tree in that type, so the truncated uint8 result is bit-identical */
  int32_t lr_w, lr_h;   /* int(w // scale), int(h // scale) */
  int32_t jpeg_quality; /* int(quality); 0 = no JPEG stage */
  int32_t gray;         /* 1: BGR2GRAY tiled to 3 channels */
  float jitter[3];      /* color_jitter shift per channel in the image's channel order; all 0 = off */
  float bilateral_sigma; /* blur_mode 4: sigmaColor = sigmaSpace */
  int32_t cj_count;      /* color_jitter_pt (ffhq_degradation_dataset.py:187-207, :290-296): number of adjustments, 0 = off */
  int32_t cj_order[4];   /* the drawn order: 0 brightness, 1 contrast, 2 saturation, 3 hue */
  float cj_factor[4];    /* the factor of each step as float32 (hue: the shift) */
  float cj_one_minus[4]; /* (float)(1.0 - factor), the second blend weight as torchvision's _blend evaluates it */
  int32_t mask_mode;     /* random_mask (ffhq_degradation_dataset.py:153-187, :299-303), needs the mask argument of
                            b200ir_degrade_full_ex: 0 off; 1 regular / half masks: masked pixels become 1.0; 2 irregular
                            mask: the whole image goes through np.array(img * 255.0, uint8) (truncation) and the drawn pixels
                            become 255 */
} b200ir_degrade_crop;

/* gt uint8 [B][H][W][3] (channel order B,G,R as the reference holds images).  gt_f32 (optional): the float image
 * img_gt in [0, 1] when it is NOT on the 8-bit grid (the dataset resizes every GT image with cv2.resize,
 * ffhq_degradation_dataset.py:230): the filter2D kinds and the no-blur case then read the float values, and gt is a
 * caller-owned scratch buffer that this call first fills with np.array(img * 255.0, dtype=uint8) for the pyblur /
 * median / bilateral kinds (degradations.py:353-366); taps fp64 [B][kmax][kmax] centred, zero
 * padded (blur_mode 2 rounds them to fp32 as cv2.filter2D does); crops: device array [B]; noise fp32 [B][lr_hmax][lr_wmax][3] already scaled by sigma/255, or NULL;
 * out fp32 NCHW [B][3][H][W] normalised to [-1, 1] (channels reversed when bgr2rgb); lr_out (optional, parity aid):
 * the low-resolution image after noise / JPEG, fp32 [B][lr_hmax][lr_wmax][3]. */
int b200ir_degrade_full(uint8_t* gt, const float* gt_f32, const double* taps, int kmax, const b200ir_degrade_crop* crops,
                        const float* noise, int lr_wmax, int lr_hmax, float* out, float* lr_out, int B, int H, int W,
                        int bgr2rgb, void* stream);
/* The same with random_mask and the 'bicubic' kind: mask uint8 [B][H][W] (non-zero = masked; rows of crops with mask_mode 0
 * are ignored), or NULL; bicubic_scratch uint8 [B][H][W][3], caller-owned, required when any crop has blur_mode 5 (the Pillow
 * round trip of those crops is written there by a first launch and read by the main one), else NULL. */
int b200ir_degrade_full_ex(uint8_t* gt, const float* gt_f32, const double* taps, int kmax, const b200ir_degrade_crop* crops,
                           const float* noise, int lr_wmax, int lr_hmax, const uint8_t* mask, uint8_t* bicubic_scratch,
                           float* out, float* lr_out, int B, int H, int W, int bgr2rgb, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200IR_H_ */
