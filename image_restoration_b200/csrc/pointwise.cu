// Memory-bound stages of the GFPGANv1OCR forward pass: NHWC fp16, 8 channels (128 bit) per thread access.
// Reference semantics: see include/b200ir.h; numerics follow upfirdn2d.py:162-192 (zero padding, FIR
// outer([1,3,3,1])/64), F.interpolate bilinear align_corners=False, fused_bias_act_kernel.cu:27-48.
#include "host_common.h"

namespace b200ir {

static constexpr float kSqrt2 = 1.4142135623730951f;
static constexpr int kPwThreads = 256;

struct H8 {
  float v[8];
};

__device__ __forceinline__ H8 ld8(const __half* p) {
  uint4 q = __ldg(reinterpret_cast<const uint4*>(p));
  const __half2* h = reinterpret_cast<const __half2*>(&q);
  H8 r;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 f = __half22float2(h[i]);
    r.v[2 * i] = f.x;
    r.v[2 * i + 1] = f.y;
  }
  return r;
}
__device__ __forceinline__ void st8(__half* p, const H8& r) {
  uint4 q;
  __half2* h = reinterpret_cast<__half2*>(&q);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(r.v[2 * i], r.v[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = q;
}
__device__ __forceinline__ float lrelu_s(float x) { return (x > 0.f ? x : 0.2f * x) * kSqrt2; }

static inline int grid_for(long long n) {
  long long g = (n + kPwThreads - 1) / kPwThreads;
  return (int)g;
}

// ------------------------------------------------------------------------------------------ first conv
__global__ void first_conv_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                  const float* __restrict__ bias, __half* __restrict__ out, int B, int HW, int cout) {
  extern __shared__ float sw[];  // [cout*3] weights, [cout] bias
  for (int i = threadIdx.x; i < cout * 3; i += blockDim.x) sw[i] = w[i];
  for (int i = threadIdx.x; i < cout; i += blockDim.x) sw[cout * 3 + i] = bias[i];
  __syncthreads();
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)B * HW) return;
  const int b = (int)(idx / HW);
  const int p = (int)(idx % HW);
  const float* xp = x + (long long)b * 3 * HW + p;
  const float r = __ldg(xp), g = __ldg(xp + HW), bl = __ldg(xp + 2 * HW);
  __half* op = out + idx * cout;
  for (int c0 = 0; c0 < cout; c0 += 8) {
    H8 o;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float* wc = sw + (c0 + j) * 3;
      o.v[j] = lrelu_s(wc[0] * r + wc[1] * g + wc[2] * bl + sw[cout * 3 + c0 + j]);
    }
    st8(op + c0, o);
  }
}

// ------------------------------------------------------------------------------------------ FIR family
// in [B][H][W][C]; out position (y, x) = sum_{a,b} k[a]k[b]/64 * in[y*sy + a - py][x*sx + b - px], zero outside.
__global__ void fir4_kernel(const __half* __restrict__ in, __half* __restrict__ out, int B, int H, int W, int C,
                            int OH, int OW, int out_h, int out_w, int step, int pad) {
  const int cg = C >> 3;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)B * OH * OW * cg;
  if (idx >= total) return;
  const int c = (int)(idx % cg) * 8;
  long long r = idx / cg;
  const int x = (int)(r % OW);
  r /= OW;
  const int y = (int)(r % OH);
  const int b = (int)(r / OH);
  const float k[4] = {0.125f, 0.375f, 0.375f, 0.125f};
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int iy = y * step + a - pad;
    if (iy < 0 || iy >= H) continue;
#pragma unroll
    for (int bb = 0; bb < 4; ++bb) {
      const int ix = x * step + bb - pad;
      if (ix < 0 || ix >= W) continue;
      const H8 v = ld8(in + (((long long)b * H + iy) * W + ix) * C + c);
      const float wgt = k[a] * k[bb];
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += wgt * v.v[j];
    }
  }
  H8 o;
#pragma unroll
  for (int j = 0; j < 8; ++j) o.v[j] = acc[j];
  st8(out + (((long long)b * out_h + y) * out_w + x) * C + c, o);
}

// ------------------------------------------------------------------------------------------ bilinear x2
__global__ void bilinear_up2_kernel(const __half* __restrict__ in, __half* __restrict__ out, int B, int h, int w,
                                    int C) {
  const int cg = C >> 3;
  const int OH = 2 * h, OW = 2 * w;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)B * OH * OW * cg;
  if (idx >= total) return;
  const int c = (int)(idx % cg) * 8;
  long long r = idx / cg;
  const int x = (int)(r % OW);
  r /= OW;
  const int y = (int)(r % OH);
  const int b = (int)(r / OH);
  const int ky = y >> 1, kx = x >> 1;
  int ya, yb, xa, xb;
  float wy0, wy1, wx0, wx1;
  if (y & 1) { ya = ky; yb = min(ky + 1, h - 1); wy0 = 0.75f; wy1 = 0.25f; }
  else       { ya = max(ky - 1, 0); yb = ky; wy0 = 0.25f; wy1 = 0.75f; }
  if (x & 1) { xa = kx; xb = min(kx + 1, w - 1); wx0 = 0.75f; wx1 = 0.25f; }
  else       { xa = max(kx - 1, 0); xb = kx; wx0 = 0.25f; wx1 = 0.75f; }
  const __half* base = in + (long long)b * h * w * C + c;
  const H8 v00 = ld8(base + ((long long)ya * w + xa) * C);
  const H8 v01 = ld8(base + ((long long)ya * w + xb) * C);
  const H8 v10 = ld8(base + ((long long)yb * w + xa) * C);
  const H8 v11 = ld8(base + ((long long)yb * w + xb) * C);
  H8 o;
#pragma unroll
  for (int j = 0; j < 8; ++j)
    o.v[j] = wy0 * (wx0 * v00.v[j] + wx1 * v01.v[j]) + wy1 * (wx0 * v10.v[j] + wx1 * v11.v[j]);
  st8(out + idx * 8, o);
}

__global__ void add_kernel(const __half* __restrict__ a, const __half* __restrict__ b, __half* __restrict__ out,
                           long long n8) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n8) return;
  const H8 x = ld8(a + idx * 8), y = ld8(b + idx * 8);
  H8 o;
#pragma unroll
  for (int j = 0; j < 8; ++j) o.v[j] = x.v[j] + y.v[j];
  st8(out + idx * 8, o);
}

// ------------------------------------------------------------------------------------------ upsample-StyleConv tail
__global__ void upfir_act_kernel(const __half* __restrict__ raw, __half* __restrict__ out, int B, int h2, int w2, int C,
                                 int raw_h, int raw_w, const float* __restrict__ noise, long long noise_sb,
                                 const float* __restrict__ noise_gain, const float* __restrict__ bias,
                                 const __half* __restrict__ scale, const __half* __restrict__ shift, int c_sft,
                                 const float* __restrict__ s_next) {
  const int cg = C >> 3;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)B * h2 * w2 * cg;
  if (idx >= total) return;
  const int c = (int)(idx % cg) * 8;
  long long r = idx / cg;
  const int x = (int)(r % w2);
  r /= w2;
  const int y = (int)(r % h2);
  const int b = (int)(r / h2);
  // FIR*4 with pad (1,1) over the (2h+1)x(2w+1) = (h2+1)x(w2+1) valid raw samples
  const float k[4] = {0.25f, 0.75f, 0.75f, 0.25f};
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int iy = y + a - 1;
    if (iy < 0 || iy > h2) continue;
#pragma unroll
    for (int bb = 0; bb < 4; ++bb) {
      const int ix = x + bb - 1;
      if (ix < 0 || ix > w2) continue;
      const H8 v = ld8(raw + (((long long)b * raw_h + iy) * raw_w + ix) * C + c);
      const float wgt = k[a] * k[bb];
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += wgt * v.v[j];
    }
  }
  float nz = 0.f;
  if (noise != nullptr) nz = __ldg(noise_gain) * __ldg(noise + b * noise_sb + (long long)y * w2 + x);
  H8 o;
#pragma unroll
  for (int j = 0; j < 8; ++j) o.v[j] = lrelu_s(acc[j] + nz + __ldg(bias + c + j));
  const int c_keep = C - c_sft;
  if (scale != nullptr && c >= c_keep) {
    const long long po = (((long long)b * h2 + y) * w2 + x) * c_sft + (c - c_keep);
    const H8 sc = ld8(scale + po), sh = ld8(shift + po);
#pragma unroll
    for (int j = 0; j < 8; ++j) o.v[j] = o.v[j] * sc.v[j] + sh.v[j];
  }
  if (s_next != nullptr) {
#pragma unroll
    for (int j = 0; j < 8; ++j) o.v[j] *= __ldg(s_next + (long long)b * C + c + j);
  }
  st8(out + idx * 8, o);
}

// ------------------------------------------------------------------------------------------ toRGB
// A group of `lanes` consecutive lanes handles one pixel; each lane covers 8 channels per iteration.
__global__ void to_rgb_kernel(const __half* __restrict__ x, int B, int h, int w, int C, const float* __restrict__ wrgb,
                              const float* __restrict__ s, const float* __restrict__ bias,
                              const float* __restrict__ skip, float* __restrict__ rgb,
                              const float* __restrict__ s_next, __half* __restrict__ xs_out, int lanes) {
  const long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long pix = gtid / lanes;
  const int sub = (int)(gtid % lanes);
  const long long npix = (long long)B * h * w;
  const bool active = pix < npix;
  const long long pc = active ? pix : (npix - 1);
  const int b = (int)(pc / ((long long)h * w));
  float a0 = 0.f, a1 = 0.f, a2 = 0.f;
  for (int c = sub * 8; c < C; c += lanes * 8) {
    H8 v = ld8(x + pc * C + c);
    float m[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) m[j] = (s != nullptr) ? v.v[j] * __ldg(s + (long long)b * C + c + j) : v.v[j];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      a0 += m[j] * __ldg(wrgb + c + j);
      a1 += m[j] * __ldg(wrgb + C + c + j);
      a2 += m[j] * __ldg(wrgb + 2 * C + c + j);
    }
    if (xs_out != nullptr && active) {
      H8 o;
#pragma unroll
      for (int j = 0; j < 8; ++j) o.v[j] = v.v[j] * __ldg(s_next + (long long)b * C + c + j);
      st8(xs_out + pc * C + c, o);
    }
  }
  for (int off = lanes >> 1; off > 0; off >>= 1) {
    a0 += __shfl_xor_sync(0xffffffffu, a0, off);
    a1 += __shfl_xor_sync(0xffffffffu, a1, off);
    a2 += __shfl_xor_sync(0xffffffffu, a2, off);
  }
  if (!active || sub != 0) return;
  const int p = (int)(pc % ((long long)h * w));
  const int y = p / w, xq = p % w;
  float o[3] = {a0 + __ldg(bias), a1 + __ldg(bias + 1), a2 + __ldg(bias + 2)};
  if (skip != nullptr) {
    // upfirdn2d(skip, FIR*4, up=2, pad=(2,1)): even 2k -> .25*s[k-1] + .75*s[k]; odd 2k+1 -> .75*s[k] + .25*s[k+1];
    // zero (not clamped) outside.
    const int hh = h >> 1, ww = w >> 1;
    int ya, yb, xa, xb;
    float wy0, wy1, wx0, wx1;
    if (y & 1) { ya = y >> 1; yb = ya + 1; wy0 = 0.75f; wy1 = 0.25f; }
    else       { yb = y >> 1; ya = yb - 1; wy0 = 0.25f; wy1 = 0.75f; }
    if (xq & 1) { xa = xq >> 1; xb = xa + 1; wx0 = 0.75f; wx1 = 0.25f; }
    else        { xb = xq >> 1; xa = xb - 1; wx0 = 0.25f; wx1 = 0.75f; }
    if (ya < 0) wy0 = 0.f;
    if (yb >= hh) wy1 = 0.f;
    if (xa < 0) wx0 = 0.f;
    if (xb >= ww) wx1 = 0.f;
    ya = max(ya, 0); yb = min(yb, hh - 1); xa = max(xa, 0); xb = min(xb, ww - 1);
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
      const float* sp = skip + ((long long)b * 3 + ch) * hh * ww;
      o[ch] += wy0 * (wx0 * __ldg(sp + ya * ww + xa) + wx1 * __ldg(sp + ya * ww + xb)) +
               wy1 * (wx0 * __ldg(sp + yb * ww + xa) + wx1 * __ldg(sp + yb * ww + xb));
    }
  }
#pragma unroll
  for (int ch = 0; ch < 3; ++ch) rgb[((long long)b * 3 + ch) * h * w + p] = o[ch];
}

// ------------------------------------------------------------------------------------------ style path
__global__ void modulate_const_kernel(const __half* __restrict__ cst, const float* __restrict__ s,
                                      __half* __restrict__ out, int B, int P, int C) {
  const int cg = C >> 3;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)B * P * cg) return;
  const int c = (int)(idx % cg) * 8;
  long long r = idx / cg;
  const int p = (int)(r % P);
  const int b = (int)(r / P);
  const H8 v = ld8(cst + (long long)p * C + c);
  H8 o;
#pragma unroll
  for (int j = 0; j < 8; ++j) o.v[j] = v.v[j] * __ldg(s + (long long)b * C + c + j);
  st8(out + idx * 8, o);
}

// one warp per (b, i): dot over F
__global__ void mod_linear_kernel(const float* __restrict__ latent, int L, int F, int lat_idx,
                                  const float* __restrict__ w, const float* __restrict__ bias, float wscale,
                                  float* __restrict__ s, int B, int cin) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= B * cin) return;
  const int b = warp / cin, i = warp % cin;
  const float* lp = latent + ((long long)b * L + lat_idx) * F;
  const float* wp = w + (long long)i * F;
  float acc = 0.f;
  for (int f = lane; f < F; f += 32) acc += __ldg(wp + f) * __ldg(lp + f);
  for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
  if (lane == 0) s[(long long)b * cin + i] = acc * wscale + __ldg(bias + i);
}

// one warp per (b, o): sum_i s^2 * wsq
__global__ void demod_kernel(const float* __restrict__ s, const float* __restrict__ wsq, float scale2,
                             float* __restrict__ d, int B, int cin, int cout) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= B * cout) return;
  const int b = warp / cout, o = warp % cout;
  const float* sp = s + (long long)b * cin;
  const float* wp = wsq + (long long)o * cin;
  float acc = 0.f;
  for (int i = lane; i < cin; i += 32) {
    const float sv = __ldg(sp + i);
    acc += sv * sv * __ldg(wp + i);
  }
  for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
  if (lane == 0) d[(long long)b * cout + o] = rsqrtf(scale2 * acc + 1e-8f);
}

__global__ void nhwc_to_nchw_f32_kernel(const __half* __restrict__ in, float* __restrict__ out, int B, int P, int C) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)B * P * C) return;
  const int p = (int)(idx % P);
  long long r = idx / P;
  const int c = (int)(r % C);
  const int b = (int)(r / C);
  out[idx] = __half2float(in[((long long)b * P + p) * C + c]);
}

}  // namespace b200ir

using namespace b200ir;
#define STREAM reinterpret_cast<cudaStream_t>(stream)

extern "C" int b200ir_first_conv(const float* x, const float* w, const float* bias, void* out, int B, int H, int W,
                                 int cout, void* stream) {
  B200IR_REQUIRE(x && w && bias && out, "first_conv: null pointer");
  B200IR_REQUIRE(cout % 8 == 0 && cout <= 512, "first_conv: cout=%d", cout);
  const long long n = (long long)B * H * W;
  first_conv_kernel<<<grid_for(n), kPwThreads, cout * 4 * sizeof(float), STREAM>>>(x, w, bias, (__half*)out, B, H * W,
                                                                                  cout);
  return check_launch("first_conv");
}

extern "C" int b200ir_fir_pad22(const void* in, void* out, int B, int H, int W, int C, int out_h, int out_w,
                                void* stream) {
  B200IR_REQUIRE(in && out && C % 8 == 0 && out_h >= H + 1 && out_w >= W + 1, "fir_pad22: bad arguments");
  const long long n = (long long)B * (H + 1) * (W + 1) * (C / 8);
  fir4_kernel<<<grid_for(n), kPwThreads, 0, STREAM>>>((const __half*)in, (__half*)out, B, H, W, C, H + 1, W + 1, out_h,
                                                      out_w, 1, 2);
  return check_launch("fir_pad22");
}

extern "C" int b200ir_fir_down2(const void* in, void* out, int B, int H, int W, int C, void* stream) {
  B200IR_REQUIRE(in && out && C % 8 == 0 && H % 2 == 0 && W % 2 == 0, "fir_down2: bad arguments");
  const long long n = (long long)B * (H / 2) * (W / 2) * (C / 8);
  fir4_kernel<<<grid_for(n), kPwThreads, 0, STREAM>>>((const __half*)in, (__half*)out, B, H, W, C, H / 2, W / 2, H / 2,
                                                      W / 2, 2, 1);
  return check_launch("fir_down2");
}

extern "C" int b200ir_bilinear_up2(const void* in, void* out, int B, int h, int w, int C, void* stream) {
  B200IR_REQUIRE(in && out && C % 8 == 0, "bilinear_up2: bad arguments");
  const long long n = (long long)B * 4 * h * w * (C / 8);
  bilinear_up2_kernel<<<grid_for(n), kPwThreads, 0, STREAM>>>((const __half*)in, (__half*)out, B, h, w, C);
  return check_launch("bilinear_up2");
}

extern "C" int b200ir_add(const void* a, const void* b, void* out, int64_t n, void* stream) {
  B200IR_REQUIRE(a && b && out && n % 8 == 0, "add: bad arguments");
  add_kernel<<<grid_for(n / 8), kPwThreads, 0, STREAM>>>((const __half*)a, (const __half*)b, (__half*)out, n / 8);
  return check_launch("add");
}

extern "C" int b200ir_upfir_act(const void* raw, void* out, int B, int h2, int w2, int C, int raw_h, int raw_w,
                                const float* noise, int64_t noise_stride_b, const float* noise_gain, const float* bias,
                                const void* scale, const void* shift, int c_sft, const float* s_next, void* stream) {
  B200IR_REQUIRE(raw && out && bias && C % 8 == 0 && raw_h >= h2 + 1 && raw_w >= w2 + 1, "upfir_act: bad arguments");
  B200IR_REQUIRE(noise == nullptr || noise_gain != nullptr, "upfir_act: noise without gain");
  B200IR_REQUIRE((scale == nullptr) == (shift == nullptr), "upfir_act: scale/shift must come together");
  B200IR_REQUIRE(scale == nullptr || (c_sft % 8 == 0 && c_sft <= C && (C - c_sft) % 8 == 0), "upfir_act: c_sft=%d",
                 c_sft);
  const long long n = (long long)B * h2 * w2 * (C / 8);
  upfir_act_kernel<<<grid_for(n), kPwThreads, 0, STREAM>>>((const __half*)raw, (__half*)out, B, h2, w2, C, raw_h, raw_w,
                                                           noise, noise_stride_b, noise_gain, bias,
                                                           (const __half*)scale, (const __half*)shift, c_sft, s_next);
  return check_launch("upfir_act");
}

extern "C" int b200ir_to_rgb(const void* x, int B, int h, int w, int C, const float* wrgb, const float* s,
                             const float* bias, const float* skip, float* rgb, const float* s_next, void* xs_out,
                             void* stream) {
  B200IR_REQUIRE(x && wrgb && bias && rgb && C % 8 == 0, "to_rgb: bad arguments");
  B200IR_REQUIRE((xs_out == nullptr) || (s_next != nullptr), "to_rgb: xs_out without s_next");
  B200IR_REQUIRE(skip == nullptr || (h % 2 == 0 && w % 2 == 0), "to_rgb: skip needs even size");
  int lanes = C / 8;
  if (lanes > 32) lanes = 32;
  int pw = 1;
  while (pw * 2 <= lanes) pw *= 2;
  lanes = pw;
  const long long n = (long long)B * h * w * lanes;
  to_rgb_kernel<<<grid_for(n), kPwThreads, 0, STREAM>>>((const __half*)x, B, h, w, C, wrgb, s, bias, skip, rgb, s_next,
                                                        (__half*)xs_out, lanes);
  return check_launch("to_rgb");
}

extern "C" int b200ir_modulate_const(const void* cst, const float* s, void* out, int B, int P, int C, void* stream) {
  B200IR_REQUIRE(cst && s && out && C % 8 == 0, "modulate_const: bad arguments");
  const long long n = (long long)B * P * (C / 8);
  modulate_const_kernel<<<grid_for(n), kPwThreads, 0, STREAM>>>((const __half*)cst, s, (__half*)out, B, P, C);
  return check_launch("modulate_const");
}

extern "C" int b200ir_mod_linear(const float* latent, int L, int F, int lat_idx, const float* w, const float* bias,
                                 float wscale, float* s, int B, int cin, void* stream) {
  B200IR_REQUIRE(latent && w && bias && s && lat_idx >= 0 && lat_idx < L, "mod_linear: bad arguments");
  const long long n = (long long)B * cin * 32;
  mod_linear_kernel<<<grid_for(n), kPwThreads, 0, STREAM>>>(latent, L, F, lat_idx, w, bias, wscale, s, B, cin);
  return check_launch("mod_linear");
}

extern "C" int b200ir_demod(const float* s, const float* wsq, float scale2, float* d, int B, int cin, int cout,
                            void* stream) {
  B200IR_REQUIRE(s && wsq && d, "demod: bad arguments");
  const long long n = (long long)B * cout * 32;
  demod_kernel<<<grid_for(n), kPwThreads, 0, STREAM>>>(s, wsq, scale2, d, B, cin, cout);
  return check_launch("demod");
}

extern "C" int b200ir_nhwc_to_nchw_f32(const void* in, float* out, int B, int P, int C, void* stream) {
  B200IR_REQUIRE(in && out, "nhwc_to_nchw_f32: bad arguments");
  const long long n = (long long)B * P * C;
  nhwc_to_nchw_f32_kernel<<<grid_for(n), kPwThreads, 0, STREAM>>>((const __half*)in, out, B, P, C);
  return check_launch("nhwc_to_nchw_f32");
}
