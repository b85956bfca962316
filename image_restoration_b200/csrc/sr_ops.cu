// Memory-bound stages of the plain-conv SR networks the reference's options/*.yml select (MSRResNet / EDSR / RCAN,
// basicsr/archs/{srresnet,edsr,rcan}_arch.py): input re-layout with mean shift, output assembly with the bilinear base
// image, and RCAN's channel attention (global average pool -> 1x1 -> ReLU -> 1x1 -> sigmoid -> scale, rcan_arch.py:8-24).
// The 3x3 convolutions, their activations, residual merges and nn.PixelShuffle run in b200ir_conv_igemm.
#include "host_common.h"
#include "ptx.cuh"

namespace b200ir {

static constexpr int kSrThreads = 256;
static inline int sr_grid(long long n) { return (int)((n + kSrThreads - 1) / kSrThreads); }
#define STREAM reinterpret_cast<cudaStream_t>(stream)

// x fp32 NCHW [B][C][H*s][W*s] -> NHWC fp16 [B][H][W][Cpad]: (x - sub[c]) * mul, zero padding channels.  s > 1 fuses
// arch_util.pixel_unshuffle (:185-201): output channel c*s*s + dy*s + dx reads x[b][c][y*s + dy][x*s + dx].
__global__ void nchw_to_nhwc_pad_kernel(const float* __restrict__ x, __half* __restrict__ out, int B, int C, int H, int W,
                                        int Cpad, const float* __restrict__ sub, float mul, int s) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // one thread per (pixel, 8-channel group)
  const int cg = Cpad >> 3;
  const int HW = H * W;
  if (idx >= (long long)B * HW * cg) return;
  const int g = (int)(idx % cg);
  const long long pix = idx / cg;
  const int b = (int)(pix / HW);
  const int p = (int)(pix % HW);
  const int yy = p / W, xx = p % W;
  const int ss = s * s;
  __half2 h[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float v[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int co = g * 8 + 2 * j + e;
      v[e] = 0.f;
      if (co < C * ss) {
        const int c = co / ss, dy = (co % ss) / s, dx = co % s;
        const long long src = (((long long)b * C + c) * (H * s) + (yy * s + dy)) * (W * s) + xx * s + dx;
        v[e] = (__ldg(x + src) - (sub ? __ldg(sub + c) : 0.f)) * mul;
      }
    }
    h[j] = f2h2_sat(v[0], v[1]);
  }
  *reinterpret_cast<uint4*>(out + idx * 8) = *reinterpret_cast<uint4*>(h);
}

// F.interpolate(scale_factor=2, mode='nearest') on NHWC fp16 (rrdbnet_arch.py:118-119): out[b][y][x] = in[b][y/2][x/2]
__global__ void nearest_up2_kernel(const __half* __restrict__ in, __half* __restrict__ out, int B, int h, int w, int C) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // one thread per (input pixel, 8 channels)
  const int cg = C >> 3;
  if (idx >= (long long)B * h * w * cg) return;
  const int g = (int)(idx % cg);
  long long r = idx / cg;
  const int xx = (int)(r % w);
  r /= w;
  const int yy = (int)(r % h);
  const int b = (int)(r / h);
  const uint4 q = __ldg(reinterpret_cast<const uint4*>(in) + idx);
  __half* op = out + (((long long)b * 2 * h + 2 * yy) * (2 * w) + 2 * xx) * C + g * 8;
  *reinterpret_cast<uint4*>(op) = q;
  *reinterpret_cast<uint4*>(op + C) = q;
  *reinterpret_cast<uint4*>(op + (long long)2 * w * C) = q;
  *reinterpret_cast<uint4*>(op + (long long)2 * w * C + C) = q;
}

// F.interpolate(scale_factor=r, mode='bilinear', align_corners=False) sample of a [h][w] plane at output (y, x)
__device__ __forceinline__ float bilinear_sample(const float* __restrict__ plane, int h, int w, int y, int x, float inv_r) {
  const float sy = fmaxf(((float)y + 0.5f) * inv_r - 0.5f, 0.f);
  const float sx = fmaxf(((float)x + 0.5f) * inv_r - 0.5f, 0.f);
  const int y0 = min((int)sy, h - 1), x0 = min((int)sx, w - 1);
  const int y1 = min(y0 + 1, h - 1), x1 = min(x0 + 1, w - 1);
  const float ty = sy - (float)y0, tx = sx - (float)x0;
  const float a = plane[y0 * w + x0], b = plane[y0 * w + x1], c = plane[y1 * w + x0], d = plane[y1 * w + x1];
  return (1.f - ty) * ((1.f - tx) * a + tx * b) + ty * ((1.f - tx) * c + tx * d);
}

// y fp32 NHWC [B][H][W][Cpad] (conv_last output) -> out fp32 NCHW [B][C][H][W]:
//   out = y * mul + add[c]  (+ bilinear x`scale` up-sampling of base [B][C][H/scale][W/scale] when base != NULL)
__global__ void sr_output_kernel(const float* __restrict__ y, float* __restrict__ out, int B, int C, int H, int W, int Cpad,
                                 float mul, const float* __restrict__ add, const float* __restrict__ base, int scale) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)B * H * W) return;
  const int b = (int)(idx / ((long long)H * W));
  const int p = (int)(idx % ((long long)H * W));
  const int yy = p / W, xx = p % W;
  for (int c = 0; c < C; ++c) {
    float v = y[idx * Cpad + c] * mul + (add ? __ldg(add + c) : 0.f);
    if (base != nullptr) {
      const int h = H / scale, w = W / scale;
      v += bilinear_sample(base + ((long long)b * C + c) * h * w, h, w, yy, xx, 1.f / (float)scale);
    }
    out[((long long)b * C + c) * H * W + p] = v;
  }
}

// nn.AdaptiveAvgPool2d(1) over an NHWC fp16 tensor: mean[b][c] (fp32).  grid (C/8, B); each thread strides over pixels.
__global__ void channel_mean_kernel(const __half* __restrict__ x, float* __restrict__ mean, int HW, int C) {
  __shared__ float red[kSrThreads][8];
  const int g = blockIdx.x, b = blockIdx.y;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const __half* xp = x + (long long)b * HW * C + g * 8;
  for (int p = threadIdx.x; p < HW; p += blockDim.x) {
    const uint4 q = __ldg(reinterpret_cast<const uint4*>(xp + (long long)p * C));
    const __half2* h = reinterpret_cast<const __half2*>(&q);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = __half22float2(h[j]);
      acc[2 * j] += f.x;
      acc[2 * j + 1] += f.y;
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) red[threadIdx.x][j] = acc[j];
  __syncthreads();
  for (int s = blockDim.x >> 1; s > 0; s >>= 1) {
    if (threadIdx.x < s) {
#pragma unroll
      for (int j = 0; j < 8; ++j) red[threadIdx.x][j] += red[threadIdx.x + s][j];
    }
    __syncthreads();
  }
  if (threadIdx.x < 8) mean[(long long)b * C + g * 8 + threadIdx.x] = red[0][threadIdx.x] / (float)HW;
}

// attention = sigmoid(W2 relu(W1 mean + b1) + b2): one block per image; W1 [Cs][C], W2 [C][Cs]
__global__ void ca_mlp_kernel(const float* __restrict__ mean, const float* __restrict__ w1, const float* __restrict__ b1,
                              const float* __restrict__ w2, const float* __restrict__ b2, float* __restrict__ att, int C,
                              int Cs) {
  extern __shared__ float sm[];  // [C] mean, [Cs] hidden
  float* s_mean = sm;
  float* s_hid = sm + C;
  const int b = blockIdx.x;
  for (int c = threadIdx.x; c < C; c += blockDim.x) s_mean[c] = mean[(long long)b * C + c];
  __syncthreads();
  for (int j = threadIdx.x; j < Cs; j += blockDim.x) {
    float a = b1[j];
    for (int c = 0; c < C; ++c) a += w1[j * C + c] * s_mean[c];
    s_hid[j] = fmaxf(a, 0.f);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float a = b2[c];
    for (int j = 0; j < Cs; ++j) a += w2[c * Cs + j] * s_hid[j];
    att[(long long)b * C + c] = 1.f / (1.f + __expf(-a));
  }
}

// RCAB tail (rcan_arch.py:43-45): out = x * att[b][c] * res_scale + identity (fp16 NHWC; att == NULL: 1, which is the
// RRDB merge out * 0.2 + x of rrdbnet_arch.py:59-63).  identity / out may be channel slices of wider NHWC buffers
// (pixel strides id_stride / out_stride in elements); x is dense.
__global__ void ca_scale_add_kernel(const __half* __restrict__ x, const float* __restrict__ att,
                                    const __half* __restrict__ identity, __half* __restrict__ out, float res_scale, int B,
                                    int HW, int C, long long id_stride, long long out_stride) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int cg = C >> 3;
  if (idx >= (long long)B * HW * cg) return;
  const int g = (int)(idx % cg);
  const long long pix = idx / cg;
  const int b = (int)(pix / HW);
  const uint4 qx = __ldg(reinterpret_cast<const uint4*>(x) + idx);
  const uint4 qi = __ldg(reinterpret_cast<const uint4*>(identity + pix * id_stride + g * 8));
  const __half2* hx = reinterpret_cast<const __half2*>(&qx);
  const __half2* hi = reinterpret_cast<const __half2*>(&qi);
  __half2 ho[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 fx = __half22float2(hx[j]), fi = __half22float2(hi[j]);
    const float a0 = (att ? __ldg(att + (long long)b * C + g * 8 + 2 * j) : 1.f) * res_scale;
    const float a1 = (att ? __ldg(att + (long long)b * C + g * 8 + 2 * j + 1) : 1.f) * res_scale;
    ho[j] = f2h2_sat(fx.x * a0 + fi.x, fx.y * a1 + fi.y);
  }
  *reinterpret_cast<uint4*>(out + pix * out_stride + g * 8) = *reinterpret_cast<uint4*>(ho);
}

}  // namespace b200ir

using namespace b200ir;

extern "C" int b200ir_nchw_to_nhwc_pad(const float* x, void* out, int B, int C, int H, int W, int Cpad, const float* sub,
                                       float mul, int unshuffle, void* stream) {
  const int s = unshuffle > 1 ? unshuffle : 1;
  B200IR_REQUIRE(x && out && C > 0 && Cpad >= C * s * s && Cpad % 8 == 0, "nchw_to_nhwc_pad: bad arguments");
  const long long n = (long long)B * H * W * (Cpad / 8);
  nchw_to_nhwc_pad_kernel<<<sr_grid(n), kSrThreads, 0, STREAM>>>(x, (__half*)out, B, C, H, W, Cpad, sub, mul, s);
  return check_launch("nchw_to_nhwc_pad");
}

extern "C" int b200ir_nearest_up2(const void* in, void* out, int B, int h, int w, int C, void* stream) {
  B200IR_REQUIRE(in && out && C % 8 == 0, "nearest_up2: bad arguments");
  const long long n = (long long)B * h * w * (C / 8);
  nearest_up2_kernel<<<sr_grid(n), kSrThreads, 0, STREAM>>>((const __half*)in, (__half*)out, B, h, w, C);
  return check_launch("nearest_up2");
}

extern "C" int b200ir_sr_output(const float* y, float* out, int B, int C, int H, int W, int Cpad, float mul,
                                const float* add, const float* base, int scale, void* stream) {
  B200IR_REQUIRE(y && out && C > 0 && Cpad >= C, "sr_output: bad arguments");
  B200IR_REQUIRE(base == nullptr || (scale >= 1 && H % scale == 0 && W % scale == 0), "sr_output: scale=%d", scale);
  sr_output_kernel<<<sr_grid((long long)B * H * W), kSrThreads, 0, STREAM>>>(y, out, B, C, H, W, Cpad, mul, add, base,
                                                                             scale);
  return check_launch("sr_output");
}

extern "C" int b200ir_channel_mean(const void* x, float* mean, int B, int HW, int C, void* stream) {
  B200IR_REQUIRE(x && mean && C % 8 == 0 && HW > 0, "channel_mean: bad arguments");
  channel_mean_kernel<<<dim3(C / 8, B), kSrThreads, 0, STREAM>>>((const __half*)x, mean, HW, C);
  return check_launch("channel_mean");
}

extern "C" int b200ir_ca_mlp(const float* mean, const float* w1, const float* b1, const float* w2, const float* b2,
                             float* att, int B, int C, int Cs, void* stream) {
  B200IR_REQUIRE(mean && w1 && b1 && w2 && b2 && att && C > 0 && Cs > 0, "ca_mlp: bad arguments");
  ca_mlp_kernel<<<B, 128, (C + Cs) * sizeof(float), STREAM>>>(mean, w1, b1, w2, b2, att, C, Cs);
  return check_launch("ca_mlp");
}

extern "C" int b200ir_ca_scale_add(const void* x, const float* att, const void* identity, void* out, float res_scale,
                                   int B, int HW, int C, int64_t id_stride, int64_t out_stride, void* stream) {
  B200IR_REQUIRE(x && identity && out && C % 8 == 0 && id_stride % 8 == 0 && out_stride % 8 == 0 && id_stride >= C &&
                     out_stride >= C,
                 "ca_scale_add: bad arguments");
  const long long n = (long long)B * HW * (C / 8);
  ca_scale_add_kernel<<<sr_grid(n), kSrThreads, 0, STREAM>>>((const __half*)x, att, (const __half*)identity, (__half*)out,
                                                            res_scale, B, HW, C, id_stride, out_stride);
  return check_launch("ca_scale_add");
}
