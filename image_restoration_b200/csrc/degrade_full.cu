// Full per-crop degradation chain of FFHQDegradationDataset.__getitem__ (ffhq_degradation_dataset.py:242-311), one CTA
// per crop, everything between the GT crop and the normalised LQ tensor in ONE launch:
//   blur        random_mixed_kernels (degradations.py:419-523): 'iso' / 'aniso' / 'motion' / 'average' -> cv2.filter2D on
//               the float image (correlation, BORDER_REFLECT_101, no quantisation); 'pyblur' -> scipy convolve2d on the
//               uint8 image, fill 255, truncated to uint8 (pyblur/*.py)
//   down        cv2.resize(INTER_LINEAR) to (lr_w, lr_h)                      (:255-256)
//   noise       clip(img + noise, 0, 1)                                       (degradations.py:660-669)
//   JPEG        add_jpg_compression (degradations.py:876-892): cv2.imencode / imdecode = libjpeg-turbo baseline 4:2:0,
//               islow DCT, quality-scaled Annex K tables, fancy up-sampling; entropy coding is lossless and skipped.
//               Integer arithmetic, restated in oracle/jpeg_oracle.py (bit-exact against cv2 in the build container).
//   up          cv2.resize(INTER_LINEAR) back to (W, H)                       (:272)
//   jitter      clip(img + shift[c], 0, 1)                                    (:90-95, :280-281)
//   gray        cv2.cvtColor(BGR2GRAY) tiled to 3 channels                    (:283-285)
//   jitter (pt) color_jitter_pt: torchvision adjust_brightness / contrast / saturation / hue in a drawn order (:187-207, :290-296)
//   tail        clamp(round(x*255), 0, 255)/255, (x-0.5)/0.5, BGR->RGB, NCHW  (:288, :307-311)
// The blur is evaluated only at the 2x2 source pixels each low-resolution pixel interpolates (4 lanes per LR pixel, combined
// with shuffles); the GT crop is staged in shared memory once.
#include "blur_taps.cuh"
#include "host_common.h"

namespace b200ir {

static constexpr int kDfThreads = 512;

struct DfAxis {
  int i0, i1;
  float w0, w1;
};
// cv2.resize(INTER_LINEAR) tap for destination index d, resize from `src_n` to `dst_n` samples, AS EXECUTED for float
// images in the build container (opencv-python 4.13 dispatches to Intel IPP): source coordinate (d + 0.5) * src/dst - 0.5
// in float64, fraction rounded to fp32, indices clamped; each pass is fma(S1 - S0, f, S0), horizontal pass first.
// Restated and pinned bit-exactly against cv2.resize in oracle/degrade_full_oracle.py::resize_linear.
// (OpenCV's own C++ path rounds the coordinate to fp32 before taking the fraction and uses S0*(1-f) + S1*f: it differs
// from IPP by up to 1.3e-5 on noise images.)
__device__ __forceinline__ DfAxis df_linear_tap(int d, int src_n, int dst_n) {
  const double scale = (double)src_n / (double)dst_n;
  const double f = (d + 0.5) * scale - 0.5;
  const double fl = floor(f);
  const int s = (int)fl;
  DfAxis a;
  a.i0 = min(max(s, 0), src_n - 1);
  a.i1 = min(max(s + 1, 0), src_n - 1);
  a.w1 = (float)(f - fl);
  a.w0 = 0.f;  // unused
  return a;
}
__device__ __forceinline__ float df_lerp(float s0, float s1, float f) { return __fmaf_rn(__fsub_rn(s1, s0), f, s0); }

// ------------------------------------------------------------------------------------------ libjpeg "islow" DCT
#define DF_DESCALE(x, n) (((x) + (1 << ((n)-1))) >> (n))
static constexpr int kF0298 = 2446, kF0390 = 3196, kF0541 = 4433, kF0765 = 6270, kF0899 = 7373, kF1175 = 9633,
                     kF1501 = 12299, kF1847 = 15137, kF1961 = 16069, kF2053 = 16819, kF2562 = 20995, kF3072 = 25172;

// jfdctint.c jpeg_fdct_islow, one pass over 8 samples (pass 1: rows, results scaled up by 2^PASS1_BITS; pass 2: columns)
template <bool kFirst>
__device__ __forceinline__ void fdct8(int (&d)[8]) {
  const int t0 = d[0] + d[7], t7 = d[0] - d[7], t1 = d[1] + d[6], t6 = d[1] - d[6];
  const int t2 = d[2] + d[5], t5 = d[2] - d[5], t3 = d[3] + d[4], t4 = d[3] - d[4];
  const int t10 = t0 + t3, t13 = t0 - t3, t11 = t1 + t2, t12 = t1 - t2;
  constexpr int sh = kFirst ? 11 : 15;  // CONST_BITS -/+ PASS1_BITS
  if (kFirst) {
    d[0] = (t10 + t11) << 2;
    d[4] = (t10 - t11) << 2;
  } else {
    d[0] = DF_DESCALE(t10 + t11, 2);
    d[4] = DF_DESCALE(t10 - t11, 2);
  }
  int z1 = (t12 + t13) * kF0541;
  d[2] = DF_DESCALE(z1 + t13 * kF0765, sh);
  d[6] = DF_DESCALE(z1 + t12 * (-kF1847), sh);
  z1 = t4 + t7;
  int z2 = t5 + t6, z3 = t4 + t6, z4 = t5 + t7;
  const int z5 = (z3 + z4) * kF1175;
  const int a4 = t4 * kF0298, a5 = t5 * kF2053, a6 = t6 * kF3072, a7 = t7 * kF1501;
  z1 *= -kF0899;
  z2 *= -kF2562;
  z3 = z3 * (-kF1961) + z5;
  z4 = z4 * (-kF0390) + z5;
  d[7] = DF_DESCALE(a4 + z1 + z3, sh);
  d[5] = DF_DESCALE(a5 + z2 + z4, sh);
  d[3] = DF_DESCALE(a6 + z2 + z3, sh);
  d[1] = DF_DESCALE(a7 + z1 + z4, sh);
}

// jidctint.c jpeg_idct_islow, one pass (pass 1: columns; pass 2: rows with the final descale)
template <bool kFirst>
__device__ __forceinline__ void idct8(int (&c)[8]) {
  int z2 = c[2], z3 = c[6];
  int z1 = (z2 + z3) * kF0541;
  const int e2 = z1 + z3 * (-kF1847);
  const int e3 = z1 + z2 * kF0765;
  const int e0 = (c[0] + c[4]) << 13, e1 = (c[0] - c[4]) << 13;
  const int t10 = e0 + e3, t13 = e0 - e3, t11 = e1 + e2, t12 = e1 - e2;
  int o0 = c[7], o1 = c[5], o2 = c[3], o3 = c[1];
  z1 = o0 + o3;
  z2 = o1 + o2;
  z3 = o0 + o2;
  int z4 = o1 + o3;
  const int z5 = (z3 + z4) * kF1175;
  o0 *= kF0298;
  o1 *= kF2053;
  o2 *= kF3072;
  o3 *= kF1501;
  z1 *= -kF0899;
  z2 *= -kF2562;
  z3 = z3 * (-kF1961) + z5;
  z4 = z4 * (-kF0390) + z5;
  o0 += z1 + z3;
  o1 += z2 + z4;
  o2 += z2 + z3;
  o3 += z1 + z4;
  constexpr int sh = kFirst ? 11 : 18;  // CONST_BITS - PASS1_BITS | CONST_BITS + PASS1_BITS + 3
  c[0] = DF_DESCALE(t10 + o3, sh);
  c[7] = DF_DESCALE(t10 - o3, sh);
  c[1] = DF_DESCALE(t11 + o2, sh);
  c[6] = DF_DESCALE(t11 - o2, sh);
  c[2] = DF_DESCALE(t12 + o1, sh);
  c[5] = DF_DESCALE(t12 - o1, sh);
  c[3] = DF_DESCALE(t13 + o0, sh);
  c[4] = DF_DESCALE(t13 - o0, sh);
}

__constant__ uint8_t kStdLuma[64] = {16, 11, 10, 16, 24,  40,  51,  61,  12, 12, 14, 19, 26,  58,  60,  55,
                                     14, 13, 16, 24, 40,  57,  69,  56,  14, 17, 22, 29, 51,  87,  80,  62,
                                     18, 22, 37, 56, 68,  109, 103, 77,  24, 35, 55, 64, 81,  104, 113, 92,
                                     49, 64, 78, 87, 103, 121, 120, 101, 72, 92, 95, 98, 112, 100, 103, 99};
__constant__ uint8_t kStdChroma[64] = {17, 18, 24, 47, 99, 99, 99, 99, 18, 21, 26, 66, 99, 99, 99, 99,
                                       24, 26, 56, 99, 99, 99, 99, 99, 47, 66, 99, 99, 99, 99, 99, 99,
                                       99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99,
                                       99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99};

// jccolor.c rgb_ycc_convert (FIX(x) = (int)(x * 65536 + 0.5))
__device__ __forceinline__ void rgb2ycc(int r, int g, int b, int& y, int& cb, int& cr) {
  y = (19595 * r + 38470 * g + 7471 * b + 32768) >> 16;
  cb = (-11059 * r - 21709 * g + 32768 * b + (128 << 16) + 32767) >> 16;
  cr = (32768 * r - 27439 * g - 5329 * b + (128 << 16) + 32767) >> 16;
}
__device__ __forceinline__ int clamp255(int v) { return min(max(v, 0), 255); }

// shared-memory layout
struct DfLayout {
  size_t row, col, urow, ucol, lut, cw, qt, lr, yp, cbp, crp, gt, total;
};
__host__ __device__ inline size_t df_align(size_t x) { return (x + 15) & ~(size_t)15; }
__host__ __device__ inline DfLayout df_layout(int kmax, int lr_wmax, int lr_hmax, int H, int W, bool stage) {
  DfLayout l;
  const int wp = (lr_wmax + 15) / 16 * 16, hp = (lr_hmax + 15) / 16 * 16;
  l.row = df_align((size_t)kmax * kmax * sizeof(DfTap));
  l.col = l.row + (size_t)lr_hmax * sizeof(DfAxis);
  l.urow = l.col + (size_t)lr_wmax * sizeof(DfAxis);
  l.ucol = l.urow + (size_t)H * sizeof(DfAxis);
  l.lut = df_align(l.ucol + (size_t)W * sizeof(DfAxis));
  l.cw = l.lut + 256 * sizeof(float);   // bilateral colour weights, 3 * 256 entries
  l.qt = l.cw + 768 * sizeof(float);
  l.lr = l.qt + 128 * sizeof(int);
  l.yp = df_align(l.lr + (size_t)lr_hmax * lr_wmax * 3 * sizeof(float));
  l.cbp = l.yp + (size_t)hp * wp * sizeof(int);
  l.crp = l.cbp + (size_t)(hp / 2) * (wp / 2) * sizeof(int);
  l.gt = df_align(l.crp + (size_t)(hp / 2) * (wp / 2) * sizeof(int));
  l.total = l.gt + (stage ? (size_t)H * W * 3 : 0);
  return l;
}

// ---- torchvision.transforms.functional colour adjustments on one float RGB pixel (functional_tensor: _blend,
// rgb_to_grayscale, _rgb2hsv, _hsv2rgb), every product / sum / quotient rounded separately in fp32 as the tensor ops are.
__device__ __forceinline__ float df_gray_tv(float r, float g, float b) {
  return __fadd_rn(__fadd_rn(__fmul_rn(0.2989f, r), __fmul_rn(0.587f, g)), __fmul_rn(0.114f, b));
}
__device__ __forceinline__ float df_blend(float a, float bb, float ratio, float one_minus) {
  return fminf(fmaxf(__fadd_rn(__fmul_rn(ratio, a), __fmul_rn(one_minus, bb)), 0.f), 1.f);
}
__device__ __forceinline__ void df_hue(float hue, float& r, float& g, float& b) {
  // _rgb2hsv
  const float maxc = fmaxf(fmaxf(r, g), b), minc = fminf(fminf(r, g), b);
  const bool eqc = maxc == minc;
  const float cr = __fsub_rn(maxc, minc);
  const float sat = __fdiv_rn(cr, eqc ? 1.f : maxc);
  const float div = eqc ? 1.f : cr;
  const float rc = __fdiv_rn(__fsub_rn(maxc, r), div), gc = __fdiv_rn(__fsub_rn(maxc, g), div),
              bc = __fdiv_rn(__fsub_rn(maxc, b), div);
  const float hr = (maxc == r) ? __fsub_rn(bc, gc) : 0.f;
  const float hg = ((maxc == g) && (maxc != r)) ? __fsub_rn(__fadd_rn(2.f, rc), bc) : 0.f;
  const float hb = ((maxc != g) && (maxc != r)) ? __fsub_rn(__fadd_rn(4.f, gc), rc) : 0.f;
  float h = __fadd_rn(__fadd_rn(hr, hg), hb);
  h = fmodf(__fadd_rn(__fdiv_rn(h, 6.f), 1.f), 1.f);
  // h = (h + hue_factor) % 1.0 (torch.remainder: the sign follows the divisor)
  h = fmodf(__fadd_rn(h, hue), 1.f);
  if (h != 0.f && h < 0.f) h = __fadd_rn(h, 1.f);
  // _hsv2rgb
  const float v = maxc;
  const float h6 = __fmul_rn(h, 6.f);
  const float fi = floorf(h6);
  const float f = __fsub_rn(h6, fi);
  const int i = ((int)fi) % 6;
  const float p = fminf(fmaxf(__fmul_rn(v, __fsub_rn(1.f, sat)), 0.f), 1.f);
  const float q = fminf(fmaxf(__fmul_rn(v, __fsub_rn(1.f, __fmul_rn(sat, f))), 0.f), 1.f);
  const float t = fminf(fmaxf(__fmul_rn(v, __fsub_rn(1.f, __fmul_rn(sat, __fsub_rn(1.f, f)))), 0.f), 1.f);
  r = i == 0 ? v : i == 1 ? q : i == 2 ? p : i == 3 ? p : i == 4 ? t : v;
  g = i == 0 ? t : i == 1 ? v : i == 2 ? v : i == 3 ? q : i == 4 ? p : p;
  b = i == 0 ? p : i == 1 ? p : i == 2 ? t : i == 3 ? v : i == 4 ? v : q;
}
// op: 0 brightness, 1 contrast (mean = mean gray of the image), 2 saturation, 3 hue (factor = hue shift)
__device__ __forceinline__ void df_color_op(int op, float factor, float one_minus, float mean, float& r, float& g,
                                            float& b) {
  if (op == 0) {
    r = df_blend(r, 0.f, factor, one_minus);
    g = df_blend(g, 0.f, factor, one_minus);
    b = df_blend(b, 0.f, factor, one_minus);
  } else if (op == 1) {
    r = df_blend(r, mean, factor, one_minus);
    g = df_blend(g, mean, factor, one_minus);
    b = df_blend(b, mean, factor, one_minus);
  } else if (op == 2) {
    const float gr = df_gray_tv(r, g, b);
    r = df_blend(r, gr, factor, one_minus);
    g = df_blend(g, gr, factor, one_minus);
    b = df_blend(b, gr, factor, one_minus);
  } else {
    df_hue(factor, r, g, b);
  }
}

// cv2.medianBlur(uint8 image, k) at (y, x), BORDER_REPLICATE: exact median of the k*k window per channel, found by a
// radix-4 search on the value (four passes over the window, three thresholds per pass).
template <bool kInterior>
__device__ __forceinline__ void df_median3(const uint8_t* __restrict__ img, int H, int W, int y, int x, int rad,
                                           int (&med)[3]) {
  const int need = ((2 * rad + 1) * (2 * rad + 1)) / 2 + 1;  // 1-based rank of the median
  int lo[3] = {0, 0, 0};
#pragma unroll 1
  for (int shift = 6; shift >= 0; shift -= 2) {
    int t1[3], c1[3] = {0, 0, 0}, c2[3] = {0, 0, 0}, c3[3] = {0, 0, 0};
    const int step = 1 << shift;
#pragma unroll
    for (int c = 0; c < 3; ++c) t1[c] = lo[c] + step;
    for (int dy = -rad; dy <= rad; ++dy) {
      const int iy = kInterior ? y + dy : min(max(y + dy, 0), H - 1);
      const uint8_t* row = img + iy * W * 3;
      for (int dx = -rad; dx <= rad; ++dx) {
        const int ix = kInterior ? x + dx : min(max(x + dx, 0), W - 1);
        const uint8_t* px = row + ix * 3;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const int v = px[c];
          c1[c] += v < t1[c];
          c2[c] += v < t1[c] + step;
          c3[c] += v < t1[c] + 2 * step;
        }
      }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) lo[c] += (c1[c] >= need) ? 0 : (c2[c] >= need) ? step : (c3[c] >= need) ? 2 * step : 3 * step;
  }
  med[0] = lo[0];
  med[1] = lo[1];
  med[2] = lo[2];
}

// cv2.bilateralFilter(uint8 image, d, sigma, sigma) at (y, x), BORDER_REFLECT_101 (bilateral_filter.simd.hpp, 8-bit three
// channel path): taps inside the radius in row-major order, weight = space[k] * colour[|db| + |dg| + |dr|] (fp32),
// sum = fma(value, weight, sum), result cvRound(sum / wsum).  nz holds the space weights (built on the host exactly as
// OpenCV does), cw the colour table.
template <bool kInterior>
__device__ __forceinline__ void df_bilateral3(const uint8_t* __restrict__ img, int H, int W, int y, int x,
                                              const DfTap* __restrict__ nz, int n_nz, const float* __restrict__ cw,
                                              int (&out)[3]) {
  const uint8_t* cp = img + (y * W + x) * 3;
  const int c0 = cp[0], c1 = cp[1], c2 = cp[2];
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, ws = 0.f;
  for (int k = 0; k < n_nz; ++k) {
    const DfTap t = nz[k];
    int iy = y + t.dy, ix = x + t.dx;
    if (!kInterior) {
      iy = iy < 0 ? -iy : (iy >= H ? 2 * H - 2 - iy : iy);
      ix = ix < 0 ? -ix : (ix >= W ? 2 * W - 2 - ix : ix);
    }
    const uint8_t* px = img + (iy * W + ix) * 3;
    const int b = px[0], g = px[1], r = px[2];
    const float w = __fmul_rn((float)t.w, cw[abs(b - c0) + abs(g - c1) + abs(r - c2)]);
    s0 = __fmaf_rn((float)b, w, s0);
    s1 = __fmaf_rn((float)g, w, s1);
    s2 = __fmaf_rn((float)r, w, s2);
    ws = __fadd_rn(ws, w);
  }
  out[0] = min(max((int)rintf(__fdiv_rn(s0, ws)), 0), 255);
  out[1] = min(max((int)rintf(__fdiv_rn(s1, ws)), 0), 255);
  out[2] = min(max((int)rintf(__fdiv_rn(s2, ws)), 0), 255);
}

template <bool kStage>
__global__ void __launch_bounds__(kDfThreads, 1)
degrade_full_kernel(const uint8_t* __restrict__ gt, const float* __restrict__ gt_f32, const double* __restrict__ taps_all,
                    int kmax,
                    const b200ir_degrade_crop* __restrict__ crops, const float* __restrict__ noise, int lr_wmax,
                    int lr_hmax, float* __restrict__ out, float* __restrict__ lr_out, int H, int W, int bgr2rgb,
                    const uint8_t* __restrict__ mask, const uint8_t* __restrict__ alt) {
  extern __shared__ __align__(16) uint8_t smem[];
  __shared__ int s_nnz;
  const int b = blockIdx.x;
  const int tid = threadIdx.x;
  const b200ir_degrade_crop cp = crops[b];
  const int lw = cp.lr_w, lh = cp.lr_h, ksz = cp.ksize, mode = cp.blur_mode;
  const DfLayout lay = df_layout(kmax, lr_wmax, lr_hmax, H, W, kStage);
  DfTap* s_nz = reinterpret_cast<DfTap*>(smem);
  DfAxis* s_row = reinterpret_cast<DfAxis*>(smem + lay.row);
  DfAxis* s_col = reinterpret_cast<DfAxis*>(smem + lay.col);
  DfAxis* s_urow = reinterpret_cast<DfAxis*>(smem + lay.urow);
  DfAxis* s_ucol = reinterpret_cast<DfAxis*>(smem + lay.ucol);
  float* s_lut = reinterpret_cast<float*>(smem + lay.lut);
  float* s_cw = reinterpret_cast<float*>(smem + lay.cw);
  int* s_qt = reinterpret_cast<int*>(smem + lay.qt);
  float* s_lr = reinterpret_cast<float*>(smem + lay.lr);
  int* s_y = reinterpret_cast<int*>(smem + lay.yp);
  int* s_cb = reinterpret_cast<int*>(smem + lay.cbp);
  int* s_cr = reinterpret_cast<int*>(smem + lay.crp);
  uint8_t* s_gt = smem + lay.gt;
  // blur_mode 5 ('bicubic'): the Pillow round trip was evaluated by pil_bicubic_kernel into `alt`; from here on the crop is a
  // no-blur crop whose source is that uint8 image (np.array(blur2, float32) / 255.0, degradations.py:379-385)
  const bool pre_blurred = mode == 5 && alt != nullptr;
  const uint8_t* g_img = (pre_blurred ? alt : gt) + (size_t)b * H * W * 3;
  const float* g_imgf = (gt_f32 != nullptr && !pre_blurred) ? gt_f32 + (size_t)b * H * W * 3 : nullptr;  // float GT: used
                                                                                  // where the reference keeps float values

  // ---- set-up: compacted taps (one warp, kernel order), resize taps, u8/255 table, quantisation tables, GT staging
  if (tid < 32) {
    const int count = df_compact_taps(taps_all + (size_t)b * kmax * kmax, kmax, ksz, mode == 4 ? 2 : (mode == 3 ? 0 : mode),
                                      s_nz, tid);
    if (tid == 0) s_nnz = count;
  }
  if (mode == 4) {  // colour weights: (float)exp(i * i * -0.5 / sigma^2), evaluated in double as OpenCV does
    const double gc = -0.5 / ((double)cp.bilateral_sigma * (double)cp.bilateral_sigma);
    for (int i = tid; i < 768; i += kDfThreads) s_cw[i] = (float)exp((double)(i * i) * gc);
  }
  for (int i = tid; i < lh; i += kDfThreads) s_row[i] = df_linear_tap(i, H, lh);
  for (int i = tid; i < lw; i += kDfThreads) s_col[i] = df_linear_tap(i, W, lw);
  for (int i = tid; i < H; i += kDfThreads) s_urow[i] = df_linear_tap(i, lh, H);
  for (int i = tid; i < W; i += kDfThreads) s_ucol[i] = df_linear_tap(i, lw, W);
  if (tid < 256) s_lut[tid] = __fdiv_rn((float)tid, 255.f);
  if (tid < 128 && cp.jpeg_quality > 0) {  // jcparam.c jpeg_quality_scaling + jpeg_add_quant_table (baseline)
    int q = min(max(cp.jpeg_quality, 1), 100);
    const int scale = q < 50 ? 5000 / q : 200 - 2 * q;
    const int base = tid < 64 ? kStdLuma[tid] : kStdChroma[tid - 64];
    s_qt[tid] = min(max((base * scale + 50) / 100, 1), 255);
  }
  if (kStage) {
    const int n16 = (H * W * 3) >> 4;
    const uint4* src = reinterpret_cast<const uint4*>(g_img);
    uint4* dst = reinterpret_cast<uint4*>(s_gt);
    for (int i = tid; i < n16; i += kDfThreads) dst[i] = __ldg(src + i);
  }
  __syncthreads();
  const uint8_t* img = kStage ? s_gt : g_img;
  const int n_nz = s_nnz;
  const int rad = (ksz - 1) >> 1;
  const bool do_blur = mode != 0 && mode != 5 && ksz > 0 && (n_nz > 0 || mode == 3);

  // ---- 1. blur at the 2x2 source pixels of every LR pixel (4 lanes per LR pixel), down-resize, noise, clip
  {
    const int total = lh * lw * 4;
    const int rounds = (total + kDfThreads - 1) / kDfThreads;
    for (int rd = 0; rd < rounds; ++rd) {
      const int it = rd * kDfThreads + tid;
      const bool live = it < total;
      const int e = live ? it : 0;
      const int q = e & 3, px = e >> 2;
      const int lx = px % lw, ly = px / lw;
      const DfAxis ry = s_row[ly], rx = s_col[lx];
      const int y = (q & 2) ? ry.i1 : ry.i0;
      const int x = (q & 1) ? rx.i1 : rx.i0;
      float v[3];
      if (do_blur) {
        if (mode == 1) {
          if (cp.blur_f64) df_blur3_at<double, 1>(img, s_lut, H, W, y, x, rad, s_nz, n_nz, v);
          else df_blur3_at<float, 1>(img, s_lut, H, W, y, x, rad, s_nz, n_nz, v);
        } else if (mode == 2) {
          df_blur3_at<float, 2>(img, s_lut, H, W, y, x, rad, s_nz, n_nz, v, g_imgf);
        } else {
          const bool inside = y >= rad && y + rad < H && x >= rad && x + rad < W;
          int u[3];
          if (mode == 3) {
            if (inside) df_median3<true>(img, H, W, y, x, rad, u);
            else df_median3<false>(img, H, W, y, x, rad, u);
          } else {
            if (inside) df_bilateral3<true>(img, H, W, y, x, s_nz, n_nz, s_cw, u);
            else df_bilateral3<false>(img, H, W, y, x, s_nz, n_nz, s_cw, u);
          }
          v[0] = s_lut[u[0]];
          v[1] = s_lut[u[1]];
          v[2] = s_lut[u[2]];
        }
      } else if (g_imgf != nullptr) {
        const float* p = g_imgf + (y * W + x) * 3;
        v[0] = p[0];
        v[1] = p[1];
        v[2] = p[2];
      } else {
        const uint8_t* p = img + (y * W + x) * 3;
        v[0] = s_lut[p[0]];
        v[1] = s_lut[p[1]];
        v[2] = s_lut[p[2]];
      }
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float p01 = __shfl_down_sync(0xffffffffu, v[c], 1);
        const float p10 = __shfl_down_sync(0xffffffffu, v[c], 2);
        const float p11 = __shfl_down_sync(0xffffffffu, v[c], 3);
        if (live && q == 0) {
          float r = df_lerp(df_lerp(v[c], p01, rx.w1), df_lerp(p10, p11, rx.w1), ry.w1);
          if (noise != nullptr) r = __fadd_rn(r, noise[(((size_t)b * lr_hmax + ly) * lr_wmax + lx) * 3 + c]);
          s_lr[px * 3 + c] = fminf(fmaxf(r, 0.f), 1.f);
        }
      }
    }
  }
  __syncthreads();

  // ---- 2. JPEG round trip on the LR image (channel order of the image is B, G, R)
  if (cp.jpeg_quality > 0) {
    const int wp = (lw + 15) / 16 * 16, hp = (lh + 15) / 16 * 16;
    const int cwp = wp / 2, chp = hp / 2;
    const int cw = (lw + 1) / 2, ch = (lh + 1) / 2;
    // 2a. 8-bit conversion (cv::saturate_cast: round half to even), colour conversion, edge replication, chroma 2x2
    for (int it = tid; it < hp * wp; it += kDfThreads) {
      const int x = it % wp, y = it / wp;
      const float* p = s_lr + (min(y, lh - 1) * lw + min(x, lw - 1)) * 3;
      const int bb = (int)rintf(__fmul_rn(p[0], 255.f)), gg = (int)rintf(__fmul_rn(p[1], 255.f)),
                rr = (int)rintf(__fmul_rn(p[2], 255.f));
      int yy, cb, cr;
      rgb2ycc(rr, gg, bb, yy, cb, cr);
      s_y[it] = yy - 128;
    }
    for (int it = tid; it < chp * cwp; it += kDfThreads) {
      const int cx = it % cwp, cy = it / cwp;
      const int cyc = min(cy, ch - 1);  // rows below the image repeat the last DOWN-SAMPLED row (jcprepct.c)
      int sb = 0, sr = 0;
#pragma unroll
      for (int dy = 0; dy < 2; ++dy)
#pragma unroll
        for (int dx = 0; dx < 2; ++dx) {
          const float* p = s_lr + (min(2 * cyc + dy, lh - 1) * lw + min(2 * cx + dx, lw - 1)) * 3;
          const int bb = (int)rintf(__fmul_rn(p[0], 255.f)), gg = (int)rintf(__fmul_rn(p[1], 255.f)),
                    rr = (int)rintf(__fmul_rn(p[2], 255.f));
          int yy, cb, cr;
          rgb2ycc(rr, gg, bb, yy, cb, cr);
          sb += cb;
          sr += cr;
        }
      const int bias = (cx & 1) ? 2 : 1;  // jcsample.c h2v2_downsample: alternating 1, 2
      s_cb[it] = ((sb + bias) >> 2) - 128;
      s_cr[it] = ((sr + bias) >> 2) - 128;
    }
    __syncthreads();
    // 2b. per 8x8 block (8 lanes each): forward DCT, quantise, de-quantise, inverse DCT -- in place
    const int nby = (hp / 8) * (wp / 8), nbc = (chp / 8) * (cwp / 8);
    const int nblocks = nby + 2 * nbc;
    const int lane8 = tid & 7;
    for (int blk = tid >> 3; blk < (nblocks + 63) / 64 * 64; blk += kDfThreads / 8) {
      const bool live = blk < nblocks;
      int* plane = s_y;
      int pw = wp, bi = live ? blk : 0;
      const int* qt = s_qt;
      if (bi >= nby) {
        bi -= nby;
        plane = s_cb;
        pw = cwp;
        qt = s_qt + 64;
        if (bi >= nbc) {
          bi -= nbc;
          plane = s_cr;
        }
      }
      const int bpr = pw / 8;
      int* base = plane + (bi / bpr) * 8 * pw + (bi % bpr) * 8;
      int d[8];
      if (live) {
#pragma unroll
        for (int k = 0; k < 8; ++k) d[k] = base[lane8 * pw + k];  // row lane8
        fdct8<true>(d);
#pragma unroll
        for (int k = 0; k < 8; ++k) base[lane8 * pw + k] = d[k];
      }
      __syncwarp();
      if (live) {
#pragma unroll
        for (int k = 0; k < 8; ++k) d[k] = base[k * pw + lane8];  // column lane8
        fdct8<false>(d);
#pragma unroll
        for (int k = 0; k < 8; ++k) {  // jcdctmgr.c quantize (divisor = table << 3), then de-quantise
          const int qv = qt[k * 8 + lane8];
          const int dv = qv << 3;
          const int a = abs(d[k]);
          int r = (a + (dv >> 1)) / dv;
          r = d[k] < 0 ? -r : r;
          d[k] = r * qv;
        }
        idct8<true>(d);
#pragma unroll
        for (int k = 0; k < 8; ++k) base[k * pw + lane8] = d[k];
      }
      __syncwarp();
      if (live) {
#pragma unroll
        for (int k = 0; k < 8; ++k) d[k] = base[lane8 * pw + k];
        idct8<false>(d);
#pragma unroll
        for (int k = 0; k < 8; ++k) base[lane8 * pw + k] = clamp255(d[k] + 128);
      }
    }
    __syncthreads();
    // 2c. fancy chroma up-sampling (jdsample.c h2v2_fancy_upsample) + jdcolor.c ycc_rgb_convert -> LR image / 255
    for (int it = tid; it < lh * lw; it += kDfThreads) {
      const int x = it % lw, y = it / lw;
      const int cx = x >> 1, cy = y >> 1;
      const int cyf = (y & 1) ? min(cy + 1, ch - 1) : max(cy - 1, 0);  // farther row
      const int cxf = (x & 1) ? min(cx + 1, cw - 1) : max(cx - 1, 0);  // farther column
      const int rnd = (x & 1) ? 7 : 8;
      int cval[2];
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const int* pl = k ? s_cr : s_cb;
        const int s_this = 3 * pl[cy * cwp + cx] + pl[cyf * cwp + cx];
        const int s_far = 3 * pl[cy * cwp + cxf] + pl[cyf * cwp + cxf];
        cval[k] = (3 * s_this + s_far + rnd) >> 4;
      }
      const int yy = s_y[y * wp + x];
      const int xb = cval[0] - 128, xr = cval[1] - 128;
      const int rr = clamp255(yy + ((91881 * xr + 32768) >> 16));
      const int gg = clamp255(yy + ((-22554 * xb + 32768 - 46802 * xr) >> 16));
      const int bb = clamp255(yy + ((116130 * xb + 32768) >> 16));
      s_lr[it * 3 + 0] = s_lut[bb];
      s_lr[it * 3 + 1] = s_lut[gg];
      s_lr[it * 3 + 2] = s_lut[rr];
    }
    __syncthreads();
  }
  if (lr_out != nullptr) {  // parity aid: the LR image after noise / JPEG, [B][lr_hmax][lr_wmax][3]
    for (int it = tid; it < lh * lw * 3; it += kDfThreads) {
      const int c = it % 3, lx = (it / 3) % lw, ly = it / (3 * lw);
      lr_out[(((size_t)b * lr_hmax + ly) * lr_wmax + lx) * 3 + c] = s_lr[it];
    }
  }

  // ---- 3. up-resize, colour jitter, gray, colour jitter (torchvision ops), 8-bit grid, normalise, NCHW
  float* o = out + (size_t)b * 3 * H * W;
  const float j0 = cp.jitter[0], j1 = cp.jitter[1], j2 = cp.jitter[2];
  const bool jit = j0 != 0.f || j1 != 0.f || j2 != 0.f;
  auto pixel = [&](int it, float (&v)[3]) {  // image after :272-285, channel order B, G, R
    const int x = it % W, y = it / W;
    const DfAxis ry = s_urow[y], rx = s_ucol[x];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float p00 = s_lr[(ry.i0 * lw + rx.i0) * 3 + c], p01 = s_lr[(ry.i0 * lw + rx.i1) * 3 + c];
      const float p10 = s_lr[(ry.i1 * lw + rx.i0) * 3 + c], p11 = s_lr[(ry.i1 * lw + rx.i1) * 3 + c];
      v[c] = df_lerp(df_lerp(p00, p01, rx.w1), df_lerp(p10, p11, rx.w1), ry.w1);
    }
    if (jit) {
      v[0] = fminf(fmaxf(__fadd_rn(v[0], j0), 0.f), 1.f);
      v[1] = fminf(fmaxf(__fadd_rn(v[1], j1), 0.f), 1.f);
      v[2] = fminf(fmaxf(__fadd_rn(v[2], j2), 0.f), 1.f);
    }
    if (cp.gray) {  // OpenCV's float BGR2GRAY as executed in the build container: fma(r, .299, fma(b, .114, g * .587))
      const float g = __fmaf_rn(v[2], 0.299f, __fmaf_rn(v[0], 0.114f, __fmul_rn(v[1], 0.587f)));
      v[0] = v[1] = v[2] = g;
    }
  };
  // color_jitter_pt (ffhq_degradation_dataset.py:187-207): the four torchvision adjustments in the drawn order; the
  // contrast step blends with the mean gray level of the whole image at that point, so the pixels before it are
  // evaluated once more in a first pass that only accumulates that mean.
  const int n_ops = cp.cj_count;
  float cj_mean = 0.f;
  if (n_ops > 0) {
    int k_con = -1;
    for (int k = 0; k < n_ops; ++k)
      if (cp.cj_order[k] == 1) k_con = k;
    if (k_con >= 0) {
      double part = 0.0;
      for (int it = tid; it < H * W; it += kDfThreads) {
        float v[3];
        pixel(it, v);
        float r = v[2], g = v[1], bl = v[0];
        for (int k = 0; k < k_con; ++k) df_color_op(cp.cj_order[k], cp.cj_factor[k], cp.cj_one_minus[k], 0.f, r, g, bl);
        part += (double)df_gray_tv(r, g, bl);
      }
      __shared__ double s_red[kDfThreads / 32];
      __shared__ float s_mean;
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) part += __shfl_xor_sync(0xffffffffu, part, off);
      if ((tid & 31) == 0) s_red[tid >> 5] = part;
      __syncthreads();
      if (tid == 0) {
        double t = 0.0;
        for (int i = 0; i < kDfThreads / 32; ++i) t += s_red[i];
        s_mean = (float)(t / (double)(H * W));
      }
      __syncthreads();
      cj_mean = s_mean;
    }
  }
  const uint8_t* mask_b = (mask != nullptr && cp.mask_mode != 0) ? mask + (size_t)b * H * W : nullptr;
  for (int it = tid; it < H * W; it += kDfThreads) {
    float v[3];
    pixel(it, v);
    if (n_ops > 0) {
      float r = v[2], g = v[1], bl = v[0];
      for (int k = 0; k < n_ops; ++k) df_color_op(cp.cj_order[k], cp.cj_factor[k], cp.cj_one_minus[k], cj_mean, r, g, bl);
      v[2] = r;
      v[1] = g;
      v[0] = bl;
    }
    // random_mask (ffhq_degradation_dataset.py:153-187, :299-303) on the clamped image: the regular / half kinds set the masked
    // pixels to 1.0; the irregular kind first takes np.array(img * 255.0, uint8) of the WHOLE image (truncation) and draws 255
    const bool masked = mask_b != nullptr && mask_b[it] != 0;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float t = fminf(fmaxf(v[c], 0.f), 1.f);
      if (cp.mask_mode == 2)
        t = truncf(__fmul_rn(t, 255.f));
      else
        t = fminf(fmaxf(rintf(__fmul_rn(t, 255.f)), 0.f), 255.f);
      if (masked) t = 255.f;
      t = __fdiv_rn(t, 255.f);
      t = __fdiv_rn(__fsub_rn(t, 0.5f), 0.5f);
      const int co = bgr2rgb ? 2 - c : c;
      o[(size_t)co * H * W + it] = t;
    }
  }
}

// ---------------------------------------------------------------------------------------------- 'bicubic' kind
// degradations.bicubic (degradations.py:379-385): Image.fromarray(uint8 image) -> torchvision Resize((h // 4, w // 4), BICUBIC)
// -> Resize((h, w), BICUBIC), i.e. two Pillow ImagingResample calls on an 8-bit RGB image.  Pillow's 8-bit path is integer
// arithmetic (src/libImaging/Resample.c): per output index the window [xmin, xmin + n) and the coefficients
// (int)(+-0.5 + w / sum(w) * 2^22) of the Keys cubic (a = -0.5) evaluated in double, support 2 * max(scale, 1); horizontal pass
// first, every pass rounds with + 2^21 >> 22 and clamps to [0, 255].  The coefficient tables are built here, in fp64 with
// explicitly rounded operations (the same IEEE sequence as the C code), so the result is Pillow's bit for bit.
struct PilAxis {
  int xmin, n, off;  // window start, length, offset of its coefficients
};
__device__ __forceinline__ double pil_cubic(double x) {
  const double a = -0.5;
  if (x < 0.0) x = -x;
  if (x < 1.0) return __dadd_rn(__dmul_rn(__dmul_rn(__dsub_rn(__dmul_rn(a + 2.0, x), a + 3.0), x), x), 1.0);
  if (x < 2.0) return __dmul_rn(__dsub_rn(__dmul_rn(__dadd_rn(__dmul_rn(__dsub_rn(x, 5.0), x), 8.0), x), 4.0), a);
  return 0.0;
}
// precompute_coeffs + normalize_coeffs_8bpc for one output index (one thread)
__device__ void pil_axis_coeffs(int in_size, int out_size, int xx, int ksize, PilAxis* ax, int* kk) {
  const double scale = __ddiv_rn((double)(float)in_size, (double)out_size);
  const double fscale = scale < 1.0 ? 1.0 : scale;
  const double support = __dmul_rn(2.0, fscale);
  const double center = __dmul_rn((double)xx + 0.5, scale);
  const double ss = __ddiv_rn(1.0, fscale);
  int xmin = (int)__dadd_rn(__dsub_rn(center, support), 0.5);
  if (xmin < 0) xmin = 0;
  int xmax = (int)__dadd_rn(__dadd_rn(center, support), 0.5);
  if (xmax > in_size) xmax = in_size;
  const int n = xmax - xmin;
  double ww = 0.0;
  for (int x = 0; x < n; ++x) ww = __dadd_rn(ww, pil_cubic(__dmul_rn(__dadd_rn(__dsub_rn((double)(x + xmin), center), 0.5), ss)));
  int* k = kk + xx * ksize;
  for (int x = 0; x < ksize; ++x) {
    int c = 0;
    if (x < n) {
      double w = pil_cubic(__dmul_rn(__dadd_rn(__dsub_rn((double)(x + xmin), center), 0.5), ss));
      if (ww != 0.0) w = __ddiv_rn(w, ww);
      c = (w < 0) ? (int)__dadd_rn(-0.5, __dmul_rn(w, 4194304.0)) : (int)__dadd_rn(0.5, __dmul_rn(w, 4194304.0));
    }
    k[x] = c;
  }
  ax[xx].xmin = xmin;
  ax[xx].n = n;
  ax[xx].off = xx * ksize;
}
__host__ __device__ inline int pil_ksize(int in_size, int out_size) {
  const double scale = (double)(float)in_size / (double)out_size;
  const double support = 2.0 * (scale < 1.0 ? 1.0 : scale);
  int c = (int)support;
  if ((double)c < support) ++c;  // ceil
  return c * 2 + 1;
}
__device__ __forceinline__ uint8_t pil_clip8(int v) { return (uint8_t)min(max(v >> 22, 0), 255); }

constexpr int kPilThreads = 512;
// one CTA per crop; crops whose blur_mode is not 5 leave at once.  src / dst uint8 [B][H][W][3]
__global__ void __launch_bounds__(kPilThreads) pil_bicubic_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst,
                                                                  const b200ir_degrade_crop* __restrict__ crops, int H, int W) {
  const int b = blockIdx.x, tid = threadIdx.x;
  if (crops[b].blur_mode != 5) return;
  extern __shared__ __align__(16) uint8_t psm[];
  const int h4 = H / 4, w4 = W / 4;
  const int kdw = pil_ksize(W, w4), kdh = pil_ksize(H, h4), kuw = pil_ksize(w4, W), kuh = pil_ksize(h4, H);
  // tables: [down W | down H | up W | up H]
  PilAxis* a_dw = reinterpret_cast<PilAxis*>(psm);
  PilAxis* a_dh = a_dw + w4;
  PilAxis* a_uw = a_dh + h4;
  PilAxis* a_uh = a_uw + W;
  int* k_dw = reinterpret_cast<int*>(a_uh + H);
  int* k_dh = k_dw + w4 * kdw;
  int* k_uw = k_dh + h4 * kdh;
  int* k_uh = k_uw + W * kuw;
  uint8_t* t_a = reinterpret_cast<uint8_t*>(k_uh + H * kuh);   // [H][w4][3]  horizontal pass of the down-scaling
  uint8_t* t_s = t_a + (size_t)H * w4 * 3;                     // [h4][w4][3] the low-resolution image
  uint8_t* t_b = t_s + (size_t)h4 * w4 * 3;                    // [h4][W][3]  horizontal pass of the up-scaling
  for (int i = tid; i < w4; i += kPilThreads) pil_axis_coeffs(W, w4, i, kdw, a_dw, k_dw);
  for (int i = tid; i < h4; i += kPilThreads) pil_axis_coeffs(H, h4, i, kdh, a_dh, k_dh);
  for (int i = tid; i < W; i += kPilThreads) pil_axis_coeffs(w4, W, i, kuw, a_uw, k_uw);
  for (int i = tid; i < H; i += kPilThreads) pil_axis_coeffs(h4, H, i, kuh, a_uh, k_uh);
  __syncthreads();
  const uint8_t* img = src + (size_t)b * H * W * 3;
  uint8_t* o = dst + (size_t)b * H * W * 3;
  for (int it = tid; it < H * w4; it += kPilThreads) {  // H x W -> H x w4
    const int y = it / w4, xx = it - y * w4;
    const PilAxis ax = a_dw[xx];
    int s0 = 1 << 21, s1 = 1 << 21, s2 = 1 << 21;
    const uint8_t* p = img + ((size_t)y * W + ax.xmin) * 3;
    for (int x = 0; x < ax.n; ++x) {
      const int c = k_dw[ax.off + x];
      s0 += p[3 * x] * c; s1 += p[3 * x + 1] * c; s2 += p[3 * x + 2] * c;
    }
    uint8_t* q = t_a + (size_t)it * 3;
    q[0] = pil_clip8(s0); q[1] = pil_clip8(s1); q[2] = pil_clip8(s2);
  }
  __syncthreads();
  for (int it = tid; it < h4 * w4; it += kPilThreads) {  // H x w4 -> h4 x w4
    const int yy = it / w4, x = it - yy * w4;
    const PilAxis ax = a_dh[yy];
    int s0 = 1 << 21, s1 = 1 << 21, s2 = 1 << 21;
    for (int y = 0; y < ax.n; ++y) {
      const int c = k_dh[ax.off + y];
      const uint8_t* p = t_a + ((size_t)(ax.xmin + y) * w4 + x) * 3;
      s0 += p[0] * c; s1 += p[1] * c; s2 += p[2] * c;
    }
    uint8_t* q = t_s + (size_t)it * 3;
    q[0] = pil_clip8(s0); q[1] = pil_clip8(s1); q[2] = pil_clip8(s2);
  }
  __syncthreads();
  for (int it = tid; it < h4 * W; it += kPilThreads) {  // h4 x w4 -> h4 x W
    const int y = it / W, xx = it - y * W;
    const PilAxis ax = a_uw[xx];
    int s0 = 1 << 21, s1 = 1 << 21, s2 = 1 << 21;
    const uint8_t* p = t_s + ((size_t)y * w4 + ax.xmin) * 3;
    for (int x = 0; x < ax.n; ++x) {
      const int c = k_uw[ax.off + x];
      s0 += p[3 * x] * c; s1 += p[3 * x + 1] * c; s2 += p[3 * x + 2] * c;
    }
    uint8_t* q = t_b + (size_t)it * 3;
    q[0] = pil_clip8(s0); q[1] = pil_clip8(s1); q[2] = pil_clip8(s2);
  }
  __syncthreads();
  for (int it = tid; it < H * W; it += kPilThreads) {  // h4 x W -> H x W
    const int yy = it / W, x = it - yy * W;
    const PilAxis ax = a_uh[yy];
    int s0 = 1 << 21, s1 = 1 << 21, s2 = 1 << 21;
    for (int y = 0; y < ax.n; ++y) {
      const int c = k_uh[ax.off + y];
      const uint8_t* p = t_b + ((size_t)(ax.xmin + y) * W + x) * 3;
      s0 += p[0] * c; s1 += p[1] * c; s2 += p[2] * c;
    }
    uint8_t* q = o + (size_t)it * 3;
    q[0] = pil_clip8(s0); q[1] = pil_clip8(s1); q[2] = pil_clip8(s2);
  }
}

static size_t pil_smem_bytes(int H, int W) {
  const int h4 = H / 4, w4 = W / 4;
  const size_t tabs = (size_t)(w4 + h4 + W + H) * sizeof(PilAxis) +
                      ((size_t)w4 * pil_ksize(W, w4) + (size_t)h4 * pil_ksize(H, h4) + (size_t)W * pil_ksize(w4, W) +
                       (size_t)H * pil_ksize(h4, H)) * sizeof(int);
  return tabs + (size_t)H * w4 * 3 + (size_t)h4 * w4 * 3 + (size_t)h4 * W * 3 + 16;
}

// np.array(img * 255.0, dtype=np.uint8) of random_pyblur / median_blur / bilateral_blur (degradations.py:353-366):
// fp32 product, truncation toward zero.
__global__ void gt_to_u8_kernel(const float* __restrict__ f, uint8_t* __restrict__ u, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) u[i] = (uint8_t)(int)fminf(fmaxf(__fmul_rn(f[i], 255.f), 0.f), 255.f);
}

}  // namespace b200ir

using namespace b200ir;

extern "C" int b200ir_degrade_full_ex(uint8_t* gt, const float* gt_f32, const double* taps, int kmax,
                                      const b200ir_degrade_crop* crops, const float* noise, int lr_wmax, int lr_hmax,
                                      const uint8_t* mask, uint8_t* bicubic_scratch, float* out, float* lr_out, int B, int H,
                                      int W, int bgr2rgb, void* stream);

extern "C" int b200ir_degrade_full(uint8_t* gt, const float* gt_f32, const double* taps, int kmax,
                                   const b200ir_degrade_crop* crops,
                                   const float* noise, int lr_wmax, int lr_hmax, float* out, float* lr_out, int B, int H,
                                   int W, int bgr2rgb, void* stream) {
  return b200ir_degrade_full_ex(gt, gt_f32, taps, kmax, crops, noise, lr_wmax, lr_hmax, nullptr, nullptr, out, lr_out, B, H, W,
                                bgr2rgb, stream);
}

extern "C" int b200ir_degrade_full_ex(uint8_t* gt, const float* gt_f32, const double* taps, int kmax,
                                      const b200ir_degrade_crop* crops, const float* noise, int lr_wmax, int lr_hmax,
                                      const uint8_t* mask, uint8_t* bicubic_scratch, float* out, float* lr_out, int B, int H,
                                      int W, int bgr2rgb, void* stream) {
  B200IR_REQUIRE(gt && taps && crops && out, "degrade_full: null pointer");
  B200IR_REQUIRE(B > 0 && H > 1 && W > 1 && kmax > 0 && (kmax & 1) && lr_wmax > 0 && lr_hmax > 0,
                 "degrade_full: bad sizes (kmax must be odd)");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  int dev = 0, smem_optin = 0;
  if (cudaGetDevice(&dev) != cudaSuccess ||
      cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess) {
    set_error("degrade_full: no CUDA device");
    return 1;
  }
  if (gt_f32 != nullptr) {
    const size_t n = (size_t)B * H * W * 3;
    gt_to_u8_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(gt_f32, gt, n);
    if (check_launch("degrade_full(gt -> uint8)")) return 1;
  }
  if (bicubic_scratch != nullptr) {  // crops of the 'bicubic' kind (blur_mode 5): Pillow round trip of the uint8 image
    B200IR_REQUIRE(H >= 8 && W >= 8, "degrade_full: the bicubic kind needs an image of at least 8x8");
    const size_t psm = pil_smem_bytes(H, W);
    B200IR_REQUIRE(psm <= (size_t)smem_optin, "degrade_full: bicubic intermediates of a %dx%d image do not fit shared memory", H, W);
    cudaFuncSetAttribute(pil_bicubic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psm);
    pil_bicubic_kernel<<<B, kPilThreads, psm, st>>>(gt, bicubic_scratch, crops, H, W);
    if (check_launch("degrade_full(bicubic)")) return 1;
  }
  const size_t base = df_layout(kmax, lr_wmax, lr_hmax, H, W, false).total;
  const size_t staged = df_layout(kmax, lr_wmax, lr_hmax, H, W, true).total;
  B200IR_REQUIRE(base <= (size_t)smem_optin, "degrade_full: low-resolution image %dx%d does not fit shared memory",
                 lr_wmax, lr_hmax);
  const bool stage = staged <= (size_t)smem_optin && (H * W * 3) % 16 == 0 && (reinterpret_cast<uintptr_t>(gt) & 15) == 0 &&
                     (reinterpret_cast<uintptr_t>(bicubic_scratch) & 15) == 0;
  if (stage) {
    cudaFuncSetAttribute(degrade_full_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)staged);
    degrade_full_kernel<true><<<B, kDfThreads, staged, st>>>(gt, gt_f32, taps, kmax, crops, noise, lr_wmax, lr_hmax, out, lr_out,
                                                            H, W, bgr2rgb, mask, bicubic_scratch);
  } else {
    cudaFuncSetAttribute(degrade_full_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)base);
    degrade_full_kernel<false><<<B, kDfThreads, base, st>>>(gt, gt_f32, taps, kmax, crops, noise, lr_wmax, lr_hmax, out, lr_out,
                                                           H, W, bgr2rgb, mask, bicubic_scratch);
  }
  return check_launch("degrade_full");
}
