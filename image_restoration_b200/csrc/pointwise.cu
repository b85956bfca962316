// Memory-bound stages of the GFPGANv1OCR forward pass: NHWC fp16, 8 channels (128 bit) per thread access.
// Reference semantics: see include/b200ir.h; numerics follow upfirdn2d.py:162-192 (zero padding, FIR
// outer([1,3,3,1])/64), F.interpolate bilinear align_corners=False, fused_bias_act_kernel.cu:27-48.
#include "host_common.h"
#include "ptx.cuh"

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
  for (int i = 0; i < 4; ++i) h[i] = f2h2_sat(r.v[2 * i], r.v[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = q;
}
__device__ __forceinline__ float lrelu_s(float x) { return (x > 0.f ? x : 0.2f * x) * kSqrt2; }

static inline int grid_for(long long n) {
  long long g = (n + kPwThreads - 1) / kPwThreads;
  return (int)g;
}

// ------------------------------------------------------------------------------------------ first conv
// One thread per pixel (coalesced fp32 NCHW reads, weights broadcast from shared memory).  A variant with cout / 8
// threads per pixel (fully contiguous 16-byte stores per warp) measured slower (94 vs 68 us at 64x128x384: the 64-bit
// index arithmetic and the 4x narrower input loads cost more than the partial-line stores).
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

// Same arithmetic, output staged through shared memory for cout = 16 / 32 / 64: each thread parks the 2 * cout bytes of
// its pixel as 16-byte chunks (chunk index XOR-swizzled so that the eight threads of a quarter warp hit eight different
// bank groups), then the CTA writes its 256 pixels as one contiguous run of 16-byte chunks: full 128-byte lines per
// warp store instead of 32 partial lines.  Four rounds of 256 pixels per CTA with all input loads issued up front.
constexpr int kFcPix = 4;  // pixels per thread (rounds of 256 pixels per CTA)
template <int kCpp>  // 16-byte chunks per pixel = cout / 8: 2, 4, 8 or 16
__global__ void __launch_bounds__(kPwThreads) first_conv_staged_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                                       const float* __restrict__ bias, __half* __restrict__ out,
                                                                       long long n_pix, int HW, int cout) {
  extern __shared__ __align__(16) uint8_t fc_smem[];
  uint4* stage = reinterpret_cast<uint4*>(fc_smem);                       // [256 pixels][cout / 8] chunks
  float* sw = reinterpret_cast<float*>(fc_smem + kPwThreads * cout * 2);  // [cout*3] weights, [cout] bias
  for (int i = threadIdx.x; i < cout * 3; i += blockDim.x) sw[i] = w[i];
  for (int i = threadIdx.x; i < cout; i += blockDim.x) sw[cout * 3 + i] = bias[i];
  __syncthreads();
  constexpr int cpp = kCpp;
  // shifts: q / cpp, px / (8 / cpp).  cout = 128 (16 chunks, 256 bytes per pixel: the discriminator's first layer): every pixel
  // starts on the same bank group, the XOR with the low pixel bits spreads a quarter warp over all eight of them
  constexpr int cs = kCpp == 2 ? 1 : (kCpp == 4 ? 2 : (kCpp == 8 ? 3 : 4)), gs = cs >= 3 ? 0 : 3 - cs;
  const int t = threadIdx.x;
  const int swz = (t >> gs) & (cpp - 1);
  const long long base0 = (long long)blockIdx.x * (kPwThreads * kFcPix);
  float r[kFcPix], g[kFcPix], bl[kFcPix];
#pragma unroll
  // image index / pixel of this thread's first pixel by one 32-bit division (n_pix < 2^31 is checked by the launcher); the
  // other rounds step from it
  unsigned img = (unsigned)(base0 + t) / (unsigned)HW, pix = (unsigned)(base0 + t) - img * (unsigned)HW;
#pragma unroll
  for (int k = 0; k < kFcPix; ++k) {                                      // all 3 * kFcPix loads in flight before any use
    r[k] = g[k] = bl[k] = 0.f;
    if (base0 + k * kPwThreads + t < n_pix) {
      const float* xp = x + ((size_t)img * 3) * HW + pix;
      r[k] = __ldg(xp), g[k] = __ldg(xp + HW), bl[k] = __ldg(xp + 2 * HW);
    }
    pix += kPwThreads;
    while (pix >= (unsigned)HW) pix -= HW, ++img;
  }
#pragma unroll
  for (int k = 0; k < kFcPix; ++k) {
    const long long base = base0 + k * kPwThreads;
    if (base >= n_pix) break;
    if (k) __syncthreads();                                               // previous round's copy-out done
#pragma unroll
    for (int j = 0; j < cpp; ++j) {
      uint4 q;
      __half2* h = reinterpret_cast<__half2*>(&q);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        // same expression as first_conv_kernel: the two kernels (and earlier builds) round identically
        const float* wc = sw + (j * 8 + 2 * c) * 3;
        const float* bc = sw + cout * 3 + j * 8 + 2 * c;
        h[c] = f2h2_sat(lrelu_s(wc[0] * r[k] + wc[1] * g[k] + wc[2] * bl[k] + bc[0]),
                        lrelu_s(wc[3] * r[k] + wc[4] * g[k] + wc[5] * bl[k] + bc[1]));
      }
      stage[(t << cs) + (j ^ swz)] = q;
    }
    __syncthreads();
    const long long rem = n_pix - base;
    const int chunks = (int)(rem < kPwThreads ? rem : kPwThreads) * cpp;
    uint4* dst = reinterpret_cast<uint4*>(out) + base * cpp;
    for (int q = t; q < chunks; q += kPwThreads) {
      const int px = q >> cs, j = q & (cpp - 1);
      __stcs(dst + q, stage[(px << cs) + (j ^ ((px >> gs) & (cpp - 1)))]);
    }
  }
}

// ------------------------------------------------------------------------------------------ optimiser step
// torch.optim.Adam (no amsgrad, the optimizer_g / optimizer_d of gfpgan_model.py:217-248, betas (0, 0.99)) over one flat
// fp32 buffer: four elements per thread, 16-byte accesses; grad_scale folds the all-reduce average (1 / world) in.
//   g = grad * grad_scale + wd * p;  m = b1 * m + (1 - b1) * g;  v = b2 * v + (1 - b2) * g * g
//   p -= (lr / (1 - b1^t)) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
__global__ void adam_step_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                 float* __restrict__ v, long long n, float b1, float b2, float eps, float wd,
                                 float step_size, float inv_bc2_sqrt, float grad_scale, float* __restrict__ ema,
                                 float ema_decay) {
  const long long i4 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i4 >= n) return;
  float pp[4], gg[4], mm[4], vv[4], ee[4];
  const bool full = i4 + 4 <= n;
  if (full) {
    *reinterpret_cast<float4*>(pp) = *reinterpret_cast<const float4*>(p + i4);
    *reinterpret_cast<float4*>(gg) = *reinterpret_cast<const float4*>(g + i4);
    *reinterpret_cast<float4*>(mm) = *reinterpret_cast<const float4*>(m + i4);
    *reinterpret_cast<float4*>(vv) = *reinterpret_cast<const float4*>(v + i4);
    if (ema) *reinterpret_cast<float4*>(ee) = *reinterpret_cast<const float4*>(ema + i4);
  } else {
    for (int k = 0; k < 4; ++k)
      if (i4 + k < n) {
        pp[k] = p[i4 + k]; gg[k] = g[i4 + k]; mm[k] = m[i4 + k]; vv[k] = v[i4 + k];
        if (ema) ee[k] = ema[i4 + k];
      }
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    float gr = gg[k] * grad_scale + wd * pp[k];
    // a non-finite gradient element (an fp16 activation gradient that overflowed under the static loss scale) must not
    // poison the parameter and its moments for good: it is dropped, the element is left where it is (bit test: this file
    // builds with fast-math).  GFPGANTrainer counts such elements and lowers its loss scale (train.py).
    if ((__float_as_uint(gr) & 0x7f800000u) == 0x7f800000u) gr = 0.f;
    mm[k] = b1 * mm[k] + (1.f - b1) * gr;
    vv[k] = b2 * vv[k] + (1.f - b2) * gr * gr;
    const float denom = __fsqrt_rn(vv[k]) * inv_bc2_sqrt + eps;          // IEEE sqrt / divide (the file builds with fast-math)
    pp[k] = pp[k] - step_size * __fdiv_rn(mm[k], denom);
    if (ema) ee[k] = ema_decay * ee[k] + (1.f - ema_decay) * pp[k];      // model_ema (base_model.py:50-57)
  }
  if (full) {
    *reinterpret_cast<float4*>(p + i4) = *reinterpret_cast<float4*>(pp);
    *reinterpret_cast<float4*>(m + i4) = *reinterpret_cast<float4*>(mm);
    *reinterpret_cast<float4*>(v + i4) = *reinterpret_cast<float4*>(vv);
    if (ema) *reinterpret_cast<float4*>(ema + i4) = *reinterpret_cast<float4*>(ee);
  } else {
    for (int k = 0; k < 4; ++k)
      if (i4 + k < n) {
        p[i4 + k] = pp[k]; m[i4 + k] = mm[k]; v[i4 + k] = vv[k];
        if (ema) ema[i4 + k] = ee[k];
      }
  }
}

// ------------------------------------------------------------------------------------------ minibatch stddev
// StyleGAN2Discriminator.forward (stylegan2_arch.py:791-801): group = min(B, stddev_group); the batch is viewed as
// (group, M = B / group, C, h, w); s[m] = mean over (c, h, w) of sqrt(var over the group (biased) + 1e-8); sample b gets
// s[b % M] as one extra channel.  x NHWC fp16 [B][P][C]; one block per m reduces P * C positions.
__global__ void mbstd_reduce_kernel(const __half* __restrict__ x, float* __restrict__ s, int M, int group, int PC) {
  const int m = blockIdx.x;
  float acc = 0.f;
  for (int i = threadIdx.x; i < PC; i += blockDim.x) {
    float v[8], mean = 0.f;
    for (int g = 0; g < group; ++g) {
      v[g] = __half2float(x[((long long)g * M + m) * PC + i]);
      mean += v[g];
    }
    mean /= (float)group;
    float var = 0.f;
    for (int g = 0; g < group; ++g) var += (v[g] - mean) * (v[g] - mean);
    acc += sqrtf(var / (float)group + 1e-8f);
  }
  __shared__ float red[32];
  for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[i];
    s[m] = t / (float)PC;
  }
}
// out [B][P][c_pad] = concat(x, s[b % M], zeros): the input of final_conv with its channel count padded to the MMA K step
__global__ void mbstd_concat_kernel(const __half* __restrict__ x, const float* __restrict__ s, __half* __restrict__ out,
                                    int B, int P, int C, int c_pad, int M) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)B * P * c_pad) return;
  const int c = (int)(idx % c_pad);
  const long long bp = idx / c_pad;
  const int b = (int)(bp / P);
  out[idx] = c < C ? x[bp * C + c] : (c == C ? __float2half_rn(s[b % M]) : __float2half_rn(0.f));
}

// ------------------------------------------------------------------------------------------ uint8 image I/O
// img2tensor + normalize (basicsr/utils/img_util.py:9-35, api.py:96-101): uint8 HWC (BGR when swap) ->
// fp32 NCHW RGB, x/255 then (x - mean)/std with mean = std = 0.5.
__global__ void u8_to_input_kernel(const uint8_t* __restrict__ img, float* __restrict__ x, int B, int HW, int swap) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)B * HW) return;
  const int b = (int)(idx / HW);
  const int p = (int)(idx % HW);
  const uint8_t* ip = img + idx * 3;
  float* xp = x + (long long)b * 3 * HW + p;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float v = __fdiv_rn((float)ip[swap ? 2 - c : c], 255.f);  // IEEE division: same bits as torch / numpy
    xp[(long long)c * HW] = (v - 0.5f) * 2.f;                        // (v - mean) / std, exact for std = 0.5
  }
}

// img2tensor(bgr2rgb) + normalize(mean .5, std .5) of a float image in [0, 1] (the GT side of
// FFHQDegradationDataset.__getitem__, ffhq_degradation_dataset.py:288, :310): fp32 HWC -> fp32 NCHW, (x - 0.5) / 0.5.
__global__ void f32_to_input_kernel(const float* __restrict__ img, float* __restrict__ x, int B, int HW, int swap) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)B * HW) return;
  const int b = (int)(idx / HW);
  const int p = (int)(idx % HW);
  const float* ip = img + idx * 3;
  float* xp = x + (long long)b * 3 * HW + p;
#pragma unroll
  for (int c = 0; c < 3; ++c) xp[(long long)c * HW] = __fdiv_rn(__fsub_rn(ip[swap ? 2 - c : c], 0.5f), 0.5f);
}

// tensor2img (img_util.py:38-94) with min_max = (-1, 1): clamp -> (x+1)/2 -> *255 -> round half to even -> uint8, CHW RGB ->
// HWC (BGR when swap).
__global__ void image_to_u8_kernel(const float* __restrict__ x, uint8_t* __restrict__ img, int B, int HW, int swap) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)B * HW) return;
  const int b = (int)(idx / HW);
  const int p = (int)(idx % HW);
  const float* xp = x + (long long)b * 3 * HW + p;
  uint8_t* op = img + idx * 3;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float v = fminf(fmaxf(xp[(long long)c * HW], -1.f), 1.f);
    v = (v + 1.f) * 0.5f;  // (x - min) / (max - min)
    op[swap ? 2 - c : c] = (uint8_t)__float2int_rn(__fmul_rn(v, 255.0f));
  }
}

// ------------------------------------------------------------------------------------------ folded ConvUpLayer helpers
// Replicate padding of an NHWC fp16 buffer [B][h+2][w+2][C] whose interior has been written: the ring takes the nearest
// interior pixel (the clamped indices of F.interpolate(..., align_corners=False), gfpganv1_ocr_arch.py:190).
__global__ void replicate_border_kernel(__half* __restrict__ t, int B, int h, int w, int C) {
  const int cg = C >> 3;
  const int ring = 2 * (w + 2) + 2 * h;  // top row, bottom row, left / right columns of the interior rows
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)B * ring * cg) return;
  const int g = (int)(idx % cg);
  long long r = idx / cg;
  const int q = (int)(r % ring);
  const int b = (int)(r / ring);
  int Y, X;
  if (q < w + 2) { Y = 0; X = q; }
  else if (q < 2 * (w + 2)) { Y = h + 1; X = q - (w + 2); }
  else { const int k = q - 2 * (w + 2); Y = 1 + (k >> 1); X = (k & 1) ? w + 1 : 0; }
  const int ys = min(max(Y, 1), h), xs = min(max(X, 1), w);
  __half* base = t + (long long)b * (h + 2) * (w + 2) * C + g * 8;
  *reinterpret_cast<uint4*>(base + ((long long)Y * (w + 2) + X) * C) =
      *reinterpret_cast<const uint4*>(base + ((long long)ys * (w + 2) + xs) * C);
}

// Corner add-back of the folded ConvUpLayer correction: the row and the column surplus both contain the term of the
// corner tap, so top[b][0] -= W[0][0] t[b,0,0], top[b][OW-1] -= W[0][2] t[b,0,w-1], bot[b][0] -= W[2][0] t[b,h-1,0],
// bot[b][OW-1] -= W[2][2] t[b,h-1,w-1].  wc fp32 [4][cout][cin]; tp = the replicate-padded input [B][h+2][w+2][cin].
__global__ void upfold_corner_kernel(const __half* __restrict__ tp, const float* __restrict__ wc, float* __restrict__ top,
                                     float* __restrict__ bot, int h, int w, int cin, int cout) {
  // block = (image, corner); one warp per output channel (lanes stride over cin: coalesced weight rows)
  const int b = blockIdx.x, corner = blockIdx.y;
  const int Y = (corner & 2) ? h : 1, X = (corner & 1) ? w : 1;
  const __half* tv = tp + (((long long)b * (h + 2) + Y) * (w + 2) + X) * cin;
  float* dst = ((corner & 2) ? bot : top) + ((long long)b * 2 * w + ((corner & 1) ? 2 * w - 1 : 0)) * cout;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  for (int co = warp; co < cout; co += nwarps) {
    const float* wr = wc + ((long long)corner * cout + co) * cin;
    float acc = 0.f;
    for (int ci = lane; ci < cin; ci += 32) acc = fmaf(__ldg(wr + ci), __half2float(tv[ci]), acc);
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if (lane == 0) dst[co] -= acc;
  }
}

// ------------------------------------------------------------------------------------------ tiled full-frame inference
// BASELINE config 4: a frame is cut into overlapping T x T tiles (positions ty[], tx[]), the tiles go through the network
// as one batch, and are blended back with separable linear-ramp weights (weight 1 on the frame border side of a border
// tile, ramp (d+1)/(ov+1) over the first / last `ov` pixels elsewhere), normalised by the weight sum.
__device__ __forceinline__ float tile_ramp(int i, int T, int ov, bool at_lo, bool at_hi) {
  float w = 1.f;
  if (!at_lo) w = fminf(w, (float)(i + 1) / (float)(ov + 1));
  if (!at_hi) w = fminf(w, (float)(T - i) / (float)(ov + 1));
  return w;
}

__global__ void tiles_gather_kernel(const float* __restrict__ frame, float* __restrict__ tiles, int C, int H, int W, int T,
                                    const int* __restrict__ ty, const int* __restrict__ tx, int nty, int ntx) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long per_tile = (long long)C * T * T;
  if (idx >= per_tile * nty * ntx) return;
  const int t = (int)(idx / per_tile);
  long long r = idx % per_tile;
  const int c = (int)(r / (T * T));
  r %= (long long)T * T;
  const int y = (int)(r / T), x = (int)(r % T);
  const int y0 = ty[t / ntx], x0 = tx[t % ntx];
  tiles[idx] = frame[((long long)c * H + y0 + y) * W + x0 + x];
}

__global__ void tiles_blend_kernel(const float* __restrict__ tiles, float* __restrict__ frame, int C, int H, int W, int T,
                                   int ov, const int* __restrict__ ty, const int* __restrict__ tx, int nty, int ntx) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)H * W) return;
  const int y = (int)(idx / W), x = (int)(idx % W);
  float num[4] = {0.f, 0.f, 0.f, 0.f}, den = 0.f;
  for (int iy = 0; iy < nty; ++iy) {
    const int y0 = ty[iy];
    if (y < y0 || y >= y0 + T) continue;
    const float wy = tile_ramp(y - y0, T, ov, y0 == 0, y0 + T == H);
    for (int ix = 0; ix < ntx; ++ix) {
      const int x0 = tx[ix];
      if (x < x0 || x >= x0 + T) continue;
      const float w = wy * tile_ramp(x - x0, T, ov, x0 == 0, x0 + T == W);
      const float* tp = tiles + ((long long)(iy * ntx + ix) * C * T + (y - y0)) * T + (x - x0);
      for (int c = 0; c < C; ++c) num[c] += w * tp[(long long)c * T * T];
      den += w;
    }
  }
  for (int c = 0; c < C; ++c) frame[((long long)c * H + y) * W + x] = num[c] / den;
}

// ------------------------------------------------------------------------------------------ FIR family
// 4x4 FIR outer(k,k), k = kscale*[1,3,3,1], evaluated for FOUR horizontally adjacent outputs per thread (8 channels):
// 4 rows x 7 columns of 16-byte loads feed 4 outputs (7 loads per output instead of 16); separable per row.
//   out(y, x0+o) = sum_{a,b} k[a] k[b] * in(y + a - pad, x0 + o + b - pad), zero outside [0,Hv) x [0,Wv).
__device__ __forceinline__ void fir_quad(const __half* __restrict__ base, int pitch_w, int C, int Hv, int Wv, int y_in0,
                                         int x_in0, float kscale, float (&acc)[4][8]) {
  const float k[4] = {kscale, 3.f * kscale, 3.f * kscale, kscale};
#pragma unroll
  for (int o = 0; o < 4; ++o)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[o][j] = 0.f;
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int iy = y_in0 + a;
    if (iy < 0 || iy >= Hv) continue;
    float hrow[4][8];
#pragma unroll
    for (int o = 0; o < 4; ++o)
#pragma unroll
      for (int j = 0; j < 8; ++j) hrow[o][j] = 0.f;
    const __half* rowp = base + (long long)iy * pitch_w * C;
#pragma unroll
    for (int col = 0; col < 7; ++col) {
      const int ix = x_in0 + col;
      if (ix < 0 || ix >= Wv) continue;
      const H8 v = ld8(rowp + (long long)ix * C);
#pragma unroll
      for (int o = 0; o < 4; ++o) {
        const int bb = col - o;
        if (bb >= 0 && bb < 4) {
#pragma unroll
          for (int j = 0; j < 8; ++j) hrow[o][j] += k[bb] * v.v[j];
        }
      }
    }
#pragma unroll
    for (int o = 0; o < 4; ++o)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[o][j] += k[a] * hrow[o][j];
  }
}

// UpFirDnSmooth with pad (2,2): in [B][H][W][C] -> out rows 0..H, cols 0..W of [B][out_h][out_w][C]
__global__ void fir_pad22_kernel(const __half* __restrict__ in, __half* __restrict__ out, int B, int H, int W, int C,
                                 int out_h, int out_w) {
  const int cg = C >> 3;
  const int OW = W + 1, OH = H + 1;
  const int xg = (OW + 3) >> 2;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)B * OH * xg * cg) return;
  const int c = (int)(idx % cg) * 8;
  long long r = idx / cg;
  const int x0 = (int)(r % xg) * 4;
  r /= xg;
  const int y = (int)(r % OH);
  const int b = (int)(r / OH);
  float acc[4][8];
  fir_quad(in + (long long)b * H * W * C + c, W, C, H, W, y - 2, x0 - 2, 0.125f, acc);
  __half* op = out + (((long long)b * out_h + y) * out_w + x0) * C + c;
#pragma unroll
  for (int o = 0; o < 4; ++o) {
    if (x0 + o < OW) {
      H8 v;
#pragma unroll
      for (int j = 0; j < 8; ++j) v.v[j] = acc[o][j];
      st8(op + (long long)o * C, v);
    }
  }
}

// pad (1,1) + stride-2 sampling (ResBlock.skip): out(y,x) = sum k[a]k[b] in(2y + a - 1, 2x + b - 1)
__global__ void fir_down2_kernel(const __half* __restrict__ in, __half* __restrict__ out, int B, int H, int W, int C) {
  const int cg = C >> 3;
  const int OH = H >> 1, OW = W >> 1;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)B * OH * OW * cg) return;
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
    const int iy = 2 * y + a - 1;
    if (iy < 0 || iy >= H) continue;
#pragma unroll
    for (int bb = 0; bb < 4; ++bb) {
      const int ix = 2 * x + bb - 1;
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
  st8(out + idx * 8, o);
}

// ------------------------------------------------------------------------------------------ bilinear x2
// One thread produces the 2x2 outputs fed by input pixel (k,l): 3x3 loads per 4 outputs.
__global__ void bilinear_up2_kernel(const __half* __restrict__ in, __half* __restrict__ out, int B, int h, int w,
                                    int C) {
  const int cg = C >> 3;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)B * h * w * cg) return;
  const int c = (int)(idx % cg) * 8;
  long long r = idx / cg;
  const int l = (int)(r % w);
  r /= w;
  const int k = (int)(r % h);
  const int b = (int)(r / h);
  const int ys[3] = {max(k - 1, 0), k, min(k + 1, h - 1)};
  const int xs[3] = {max(l - 1, 0), l, min(l + 1, w - 1)};
  const __half* base = in + (long long)b * h * w * C + c;
  float lo[3][8], hi[3][8];  // horizontal pass: lo -> output col 2l, hi -> output col 2l+1
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const __half* rp = base + (long long)ys[a] * w * C;
    const H8 v0 = ld8(rp + (long long)xs[0] * C), v1 = ld8(rp + (long long)xs[1] * C), v2 = ld8(rp + (long long)xs[2] * C);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      lo[a][j] = 0.25f * v0.v[j] + 0.75f * v1.v[j];
      hi[a][j] = 0.75f * v1.v[j] + 0.25f * v2.v[j];
    }
  }
  const int OW = 2 * w;
  __half* op = out + (((long long)b * 2 * h + 2 * k) * OW + 2 * l) * C + c;
  H8 o;
#pragma unroll
  for (int j = 0; j < 8; ++j) o.v[j] = 0.25f * lo[0][j] + 0.75f * lo[1][j];
  st8(op, o);
#pragma unroll
  for (int j = 0; j < 8; ++j) o.v[j] = 0.25f * hi[0][j] + 0.75f * hi[1][j];
  st8(op + C, o);
#pragma unroll
  for (int j = 0; j < 8; ++j) o.v[j] = 0.75f * lo[1][j] + 0.25f * lo[2][j];
  st8(op + (long long)OW * C, o);
#pragma unroll
  for (int j = 0; j < 8; ++j) o.v[j] = 0.75f * hi[1][j] + 0.25f * hi[2][j];
  st8(op + (long long)OW * C + C, o);
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
// Four adjacent outputs per thread (w2 is a multiple of 4 on this path).
__global__ void upfir_act_kernel(const __half* __restrict__ raw, __half* __restrict__ out, int B, int h2, int w2, int C,
                                 int raw_h, int raw_w, const float* __restrict__ noise, long long noise_sb,
                                 const float* __restrict__ noise_gain, const float* __restrict__ bias,
                                 const __half* __restrict__ scale, const __half* __restrict__ shift, int c_sft,
                                 const float* __restrict__ s_next) {
  const int cg = C >> 3;
  const int xg = w2 >> 2;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)B * h2 * xg * cg) return;
  const int c = (int)(idx % cg) * 8;
  long long r = idx / cg;
  const int x0 = (int)(r % xg) * 4;
  r /= xg;
  const int y = (int)(r % h2);
  const int b = (int)(r / h2);
  float acc[4][8];
  // FIR*4 (k = [1,3,3,1]/4 per axis) with pad (1,1) over the (h2+1) x (w2+1) valid raw samples
  fir_quad(raw + (long long)b * raw_h * raw_w * C + c, raw_w, C, h2 + 1, w2 + 1, y - 1, x0 - 1, 0.25f, acc);
  float bs[8], sn[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    bs[j] = __ldg(bias + c + j);
    sn[j] = (s_next != nullptr) ? __ldg(s_next + (long long)b * C + c + j) : 1.f;
  }
  const float g = (noise != nullptr) ? __ldg(noise_gain) : 0.f;
  const int c_keep = C - c_sft;
  const bool sft = (scale != nullptr) && (c >= c_keep);
#pragma unroll
  for (int o = 0; o < 4; ++o) {
    const int x = x0 + o;
    const float nz = (noise != nullptr) ? g * __ldg(noise + b * noise_sb + (long long)y * w2 + x) : 0.f;
    H8 v;
#pragma unroll
    for (int j = 0; j < 8; ++j) v.v[j] = lrelu_s(acc[o][j] + nz + bs[j]);
    const long long pix = ((long long)b * h2 + y) * w2 + x;
    if (sft) {
      const long long po = pix * c_sft + (c - c_keep);
      const H8 sc = ld8(scale + po), sh = ld8(shift + po);
#pragma unroll
      for (int j = 0; j < 8; ++j) v.v[j] = v.v[j] * sc.v[j] + sh.v[j];
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) v.v[j] *= sn[j];
    st8(out + pix * C + c, v);
  }
}

// ------------------------------------------------------------------------------------------ toRGB
// grid (pixel blocks, B): a block works on kRgbPix pixels of ONE image, so every thread keeps its channel group's
// modulated RGB weights (w[o][c]*s[b][c]) and next-layer modulation in registers across all its pixels.
// `lanes` consecutive lanes share a pixel (8 channels each, G = C / (8*lanes) groups per lane).
static constexpr int kRgbPix = 256;
template <int G>
__global__ void to_rgb_kernel(const __half* __restrict__ x, int HW, int w, int C, const float* __restrict__ wrgb,
                              const float* __restrict__ s, const float* __restrict__ bias,
                              const float* __restrict__ skip, float* __restrict__ rgb,
                              const float* __restrict__ s_next, __half* __restrict__ xs_out, int lanes) {
  const int b = blockIdx.y;
  const int sub = threadIdx.x % lanes;
  const int pslot = threadIdx.x / lanes;
  const int pstep = blockDim.x / lanes;
  float wm[G][3][8], sn[G][8];
#pragma unroll
  for (int g = 0; g < G; ++g) {
    const int c = (g * lanes + sub) * 8;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float sv = (s != nullptr) ? __ldg(s + (long long)b * C + c + j) : 1.f;
      wm[g][0][j] = __ldg(wrgb + c + j) * sv;
      wm[g][1][j] = __ldg(wrgb + C + c + j) * sv;
      wm[g][2][j] = __ldg(wrgb + 2 * C + c + j) * sv;
      sn[g][j] = (s_next != nullptr) ? __ldg(s_next + (long long)b * C + c + j) : 1.f;
    }
  }
  const int h = HW / w;
  const int p_end = min(HW, (int)(blockIdx.x + 1) * kRgbPix);
  for (int p0 = blockIdx.x * kRgbPix; p0 < p_end; p0 += pstep) {
    const int p = p0 + pslot;
    const bool active = p < p_end;
    const long long pix = (long long)b * HW + (active ? p : p_end - 1);
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
#pragma unroll
    for (int g = 0; g < G; ++g) {
      const int c = (g * lanes + sub) * 8;
      const H8 v = ld8(x + pix * C + c);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        a0 += v.v[j] * wm[g][0][j];
        a1 += v.v[j] * wm[g][1][j];
        a2 += v.v[j] * wm[g][2][j];
      }
      if (xs_out != nullptr && active) {
        H8 o;
#pragma unroll
        for (int j = 0; j < 8; ++j) o.v[j] = v.v[j] * sn[g][j];
        st8(xs_out + pix * C + c, o);
      }
    }
    for (int off = lanes >> 1; off > 0; off >>= 1) {
      a0 += __shfl_xor_sync(0xffffffffu, a0, off);
      a1 += __shfl_xor_sync(0xffffffffu, a1, off);
      a2 += __shfl_xor_sync(0xffffffffu, a2, off);
    }
    if (!active || sub != 0) continue;
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
    for (int ch = 0; ch < 3; ++ch) rgb[((long long)b * 3 + ch) * HW + p] = o[ch];
  }
}

// second half of the fused ToRGB: bias + partial planes written by the conv epilogues + up-sampled skip
// Four horizontally adjacent pixels per thread (w % 4 == 0): 16-byte loads of the partial planes and stores of the
// result; the up-sampled skip of the quad needs skip columns x/2 - 1 .. x/2 + 2 of two rows.
__global__ void rgb_combine4_kernel(const float* __restrict__ part, int n_parts, const float* __restrict__ bias,
                                    const float* __restrict__ skip, float* __restrict__ rgb, int B, int h, int w) {
  const int wq = w >> 2;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long HW = (long long)h * w;
  if (idx >= (long long)B * h * wq) return;
  // 32-bit index arithmetic (the launcher keeps B * h * w / 4 below 2^31): 64-bit divisions by run-time values cost
  // more instructions than the rest of the thread
  const unsigned i32 = (unsigned)idx;
  const unsigned r = i32 / (unsigned)wq;
  const int xq = (int)(i32 - r * (unsigned)wq);
  const int b = (int)(r / (unsigned)h);
  const int y = (int)(r - (unsigned)b * (unsigned)h);
  const int x0 = xq * 4;
  const long long p = (long long)y * w + x0;
  float o[3][4];
#pragma unroll
  for (int ch = 0; ch < 3; ++ch) {
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int t = 0; t < n_parts; ++t) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(part + (((long long)t * B + b) * 3 + ch) * HW + p));
      a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
    }
    const float bb = __ldg(bias + ch);
    o[ch][0] = a.x + bb; o[ch][1] = a.y + bb; o[ch][2] = a.z + bb; o[ch][3] = a.w + bb;
  }
  if (skip != nullptr) {
    const int hh = h >> 1, ww = w >> 1;
    int ya, yb;
    float wy0, wy1;
    if (y & 1) { ya = y >> 1; yb = ya + 1; wy0 = 0.75f; wy1 = 0.25f; }
    else       { yb = y >> 1; ya = yb - 1; wy0 = 0.25f; wy1 = 0.75f; }
    if (ya < 0) wy0 = 0.f;
    if (yb >= hh) wy1 = 0.f;
    ya = max(ya, 0); yb = min(yb, hh - 1);
    const int xs = x0 >> 1;  // skip columns xs-1, xs, xs+1, xs+2
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
      const float* sp = skip + ((long long)b * 3 + ch) * hh * ww;
      float c[4];
      // pixel 0: cols xs-1, xs | pixel 1: xs, xs+1 | pixel 2: xs, xs+1 | pixel 3: xs+1, xs+2; c[k] = 0 outside the skip image
      float ra[4], rb[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int xc = xs - 1 + k;
        const int xcl = min(max(xc, 0), ww - 1);
        const float m = (xc >= 0 && xc < ww) ? 1.f : 0.f;
        ra[k] = __ldg(sp + ya * ww + xcl);
        rb[k] = __ldg(sp + yb * ww + xcl);
        c[k] = m;
      }
      // even pixel 2k: .25 * s[k-1] + .75 * s[k]; odd pixel 2k+1: .75 * s[k] + .25 * s[k+1]
      o[ch][0] += wy0 * (0.25f * c[0] * ra[0] + 0.75f * c[1] * ra[1]) + wy1 * (0.25f * c[0] * rb[0] + 0.75f * c[1] * rb[1]);
      o[ch][1] += wy0 * (0.75f * c[1] * ra[1] + 0.25f * c[2] * ra[2]) + wy1 * (0.75f * c[1] * rb[1] + 0.25f * c[2] * rb[2]);
      o[ch][2] += wy0 * (0.25f * c[1] * ra[1] + 0.75f * c[2] * ra[2]) + wy1 * (0.25f * c[1] * rb[1] + 0.75f * c[2] * rb[2]);
      o[ch][3] += wy0 * (0.75f * c[2] * ra[2] + 0.25f * c[3] * ra[3]) + wy1 * (0.75f * c[2] * rb[2] + 0.25f * c[3] * rb[3]);
    }
  }
#pragma unroll
  for (int ch = 0; ch < 3; ++ch)
    *reinterpret_cast<float4*>(rgb + ((long long)b * 3 + ch) * HW + p) = make_float4(o[ch][0], o[ch][1], o[ch][2], o[ch][3]);
}

__global__ void rgb_combine_kernel(const float* __restrict__ part, int n_parts, const float* __restrict__ bias,
                                   const float* __restrict__ skip, float* __restrict__ rgb, int B, int h, int w) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long HW = (long long)h * w;
  if (idx >= B * HW) return;
  const int b = (int)(idx / HW);
  const int p = (int)(idx % HW);
  const int y = p / w, xq = p % w;
  float o[3];
#pragma unroll
  for (int ch = 0; ch < 3; ++ch) {
    float a = 0.f;
    for (int t = 0; t < n_parts; ++t) a += __ldg(part + (((long long)t * B + b) * 3 + ch) * HW + p);
    o[ch] = a + __ldg(bias + ch);
  }
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
  for (int ch = 0; ch < 3; ++ch) rgb[((long long)b * 3 + ch) * HW + p] = o[ch];
}

__global__ void rgb_wmod_kernel(const float* __restrict__ w, const float* __restrict__ s, float* __restrict__ wm, int B,
                                int C) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * 3 * C) return;
  const int c = idx % C;
  const int o = (idx / C) % 3;
  const int b = idx / (3 * C);
  wm[idx] = __ldg(w + o * C + c) * (s != nullptr ? __ldg(s + (long long)b * C + c) : 1.f);
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

// All modulation linears / demodulation tables of a forward in one launch each.  Both are small row-dot products
//   out[b][o] = post( sum_c W[o][c] * f(X[b][c]) )
// grid (output blocks of kRdOut, layer, batch groups of kRdBatch).  A block stages f(X) of its kRdBatch images in shared
// memory as [c][kRdBatch]; each thread owns one output o and walks over c: one weight load + kRdBatch/4 broadcast
// LDS.128 + kRdBatch FMAs per c, no cross-lane reduction.  (The first version re-read a weight row per image and spent
// ~90 us per launch on L2 traffic.)
static constexpr int kRdBatch = 16;
static constexpr int kRdOut = 128;
static constexpr int kRdMaxC = 512;

template <bool SQUARE>
__device__ __forceinline__ void rowdot_block(const float* __restrict__ W, int n_out, int C, const float* X,
                                             long long x_stride, int B, float* acc_out /*[kRdBatch]*/, int o,
                                             float* sx /*[C][kRdBatch]*/) {
  const int b0 = blockIdx.z * kRdBatch;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {  // coalesced over c, kRdBatch independent loads in flight
    float v[kRdBatch];
#pragma unroll
    for (int j = 0; j < kRdBatch; ++j) v[j] = (b0 + j < B) ? X[(long long)(b0 + j) * x_stride + c] : 0.f;
#pragma unroll
    for (int q = 0; q < kRdBatch / 4; ++q) {
      float4 x4 = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
      if (SQUARE) x4 = make_float4(x4.x * x4.x, x4.y * x4.y, x4.z * x4.z, x4.w * x4.w);
      reinterpret_cast<float4*>(sx + c * kRdBatch)[q] = x4;
    }
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < kRdBatch; ++j) acc_out[j] = 0.f;
  if (o >= n_out) return;
  const float4* wp = reinterpret_cast<const float4*>(W + (long long)o * C);  // C % 4 == 0 (checked by the launcher)
#pragma unroll 2
  for (int c4 = 0; c4 < C / 4; ++c4) {
    const float4 w4 = __ldg(wp + c4);
    const float wv[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float4* xp = reinterpret_cast<const float4*>(sx + (4 * c4 + k) * kRdBatch);
#pragma unroll
      for (int q = 0; q < kRdBatch / 4; ++q) {
        const float4 x4 = xp[q];
        acc_out[4 * q] = fmaf(wv[k], x4.x, acc_out[4 * q]);
        acc_out[4 * q + 1] = fmaf(wv[k], x4.y, acc_out[4 * q + 1]);
        acc_out[4 * q + 2] = fmaf(wv[k], x4.z, acc_out[4 * q + 2]);
        acc_out[4 * q + 3] = fmaf(wv[k], x4.w, acc_out[4 * q + 3]);
      }
    }
  }
}

__global__ void __launch_bounds__(kRdOut) mod_linear_multi_kernel(const float* __restrict__ latent, int L, int F,
                                                                  const b200ir_mod_layer* layers, float wscale, int B) {
  __shared__ __align__(16) float sx[kRdMaxC * kRdBatch];
  const b200ir_mod_layer ly = layers[blockIdx.y];
  if (blockIdx.x * kRdOut >= ly.cin) return;
  const int o = blockIdx.x * kRdOut + threadIdx.x;
  float acc[kRdBatch];
  rowdot_block<false>(ly.w, ly.cin, F, latent + (long long)ly.lat_idx * F, (long long)L * F, B, acc, o, sx);
  if (o >= ly.cin) return;
  const float bias = __ldg(ly.bias + o);
#pragma unroll
  for (int j = 0; j < kRdBatch; ++j) {
    const int b = blockIdx.z * kRdBatch + j;
    if (b < B) ly.s[(long long)b * ly.cin + o] = acc[j] * wscale + bias;
  }
}

__global__ void __launch_bounds__(kRdOut) demod_multi_kernel(const b200ir_demod_layer* layers, int B) {
  __shared__ __align__(16) float sx[kRdMaxC * kRdBatch];
  const b200ir_demod_layer ly = layers[blockIdx.y];
  if (blockIdx.x * kRdOut >= ly.cout) return;
  const int o = blockIdx.x * kRdOut + threadIdx.x;
  float acc[kRdBatch];
  rowdot_block<true>(ly.wsq, ly.cout, ly.cin, ly.s, ly.cin, B, acc, o, sx);
  if (o >= ly.cout) return;
#pragma unroll
  for (int j = 0; j < kRdBatch; ++j) {
    const int b = blockIdx.z * kRdBatch + j;
    if (b < B) ly.d[(long long)b * ly.cout + o] = rsqrtf(ly.scale2 * acc[j] + 1e-8f);
  }
}

// Style MLP of StyleGAN2OCRGenerator (stylegan2_ocr_arch.py:12-23, 424-430; used when input_is_latent=False):
// NormStyleCode x * rsqrt(mean(x^2) + 1e-8), then num_mlp x [EqualLinear(F, F, lr_mul) + fused leaky-ReLU]:
//   y = lrelu(x W^T * (lr_mul / sqrt(F)) + b * lr_mul, 0.2) * sqrt(2).   One block per style vector; w [n_layers][F][F].
__global__ void style_mlp_kernel(const float* __restrict__ z, const float* __restrict__ w, const float* __restrict__ bias,
                                 float* __restrict__ out, int F, int n_layers, float lr_mul) {
  extern __shared__ float sv[];  // [2][F]
  float* cur = sv;
  float* nxt = sv + F;
  __shared__ float red[32];
  const int b = blockIdx.x;
  float ss = 0.f;
  for (int i = threadIdx.x; i < F; i += blockDim.x) {
    const float v = z[(long long)b * F + i];
    cur[i] = v;
    ss += v * v;
  }
  for (int off = 16; off > 0; off >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, off);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
  __syncthreads();
  float tot = 0.f;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) tot += red[i];
  const float nrm = rsqrtf(tot / (float)F + 1e-8f);
  for (int i = threadIdx.x; i < F; i += blockDim.x) cur[i] *= nrm;
  __syncthreads();
  const float wscale = lr_mul * rsqrtf((float)F);
  for (int l = 0; l < n_layers; ++l) {
    const float* wl = w + (long long)l * F * F;
    for (int o = threadIdx.x; o < F; o += blockDim.x) {
      float acc = 0.f;
      for (int k = 0; k < F; ++k) acc = fmaf(__ldg(wl + (long long)o * F + k), cur[k], acc);
      nxt[o] = lrelu_s(acc * wscale + __ldg(bias + l * F + o) * lr_mul);
    }
    __syncthreads();
    float* t = cur;
    cur = nxt;
    nxt = t;
  }
  for (int i = threadIdx.x; i < F; i += blockDim.x) out[(long long)b * F + i] = cur[i];
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
  if ((cout == 16 || cout == 32 || cout == 64 || cout == 128) && n < (1LL << 31)) {
    const size_t smem = (size_t)kPwThreads * cout * 2 + cout * 4 * sizeof(float);  // <= 33 KB (cout = 128: 66 KB, opt-in)
    const int grid = grid_for((n + kFcPix - 1) / kFcPix);
    if (cout == 128) {
      static bool configured = false;
      if (!configured) {
        if (cudaFuncSetAttribute(first_conv_staged_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) !=
            cudaSuccess) {
          set_error("first_conv: cudaFuncSetAttribute failed");
          return 1;
        }
        configured = true;
      }
      first_conv_staged_kernel<16><<<grid, kPwThreads, smem, STREAM>>>(x, w, bias, (__half*)out, n, H * W, cout);
    } else if (cout == 16)
      first_conv_staged_kernel<2><<<grid, kPwThreads, smem, STREAM>>>(x, w, bias, (__half*)out, n, H * W, cout);
    else if (cout == 32)
      first_conv_staged_kernel<4><<<grid, kPwThreads, smem, STREAM>>>(x, w, bias, (__half*)out, n, H * W, cout);
    else
      first_conv_staged_kernel<8><<<grid, kPwThreads, smem, STREAM>>>(x, w, bias, (__half*)out, n, H * W, cout);
    return check_launch("first_conv");
  }
  first_conv_kernel<<<grid_for(n), kPwThreads, cout * 4 * sizeof(float), STREAM>>>(x, w, bias, (__half*)out, B, H * W,
                                                                                  cout);
  return check_launch("first_conv");
}

extern "C" int b200ir_replicate_border(void* t, int B, int h, int w, int C, void* stream) {
  B200IR_REQUIRE(t && C % 8 == 0 && h > 0 && w > 0, "replicate_border: bad arguments");
  const long long n = (long long)B * (2 * (w + 2) + 2 * h) * (C / 8);
  replicate_border_kernel<<<grid_for(n), kPwThreads, 0, STREAM>>>((__half*)t, B, h, w, C);
  return check_launch("replicate_border");
}

extern "C" int b200ir_upfold_corners(const void* tp, const float* wc, float* top, float* bot, int B, int h, int w, int cin,
                                     int cout, void* stream) {
  B200IR_REQUIRE(tp && wc && top && bot && B > 0, "upfold_corners: bad arguments");
  upfold_corner_kernel<<<dim3(B, 4), 256, 0, STREAM>>>((const __half*)tp, wc, top, bot, h, w, cin, cout);
  return check_launch("upfold_corners");
}

extern "C" int b200ir_tiles_gather(const float* frame, float* tiles, int C, int H, int W, int T, const int32_t* ty,
                                   const int32_t* tx, int nty, int ntx, void* stream) {
  B200IR_REQUIRE(frame && tiles && ty && tx && C > 0 && T > 0 && T <= H && T <= W && nty > 0 && ntx > 0,
                 "tiles_gather: bad arguments");
  tiles_gather_kernel<<<grid_for((long long)C * T * T * nty * ntx), kPwThreads, 0, STREAM>>>(frame, tiles, C, H, W, T, ty,
                                                                                           tx, nty, ntx);
  return check_launch("tiles_gather");
}

extern "C" int b200ir_tiles_blend(const float* tiles, float* frame, int C, int H, int W, int T, int overlap,
                                  const int32_t* ty, const int32_t* tx, int nty, int ntx, void* stream) {
  B200IR_REQUIRE(frame && tiles && ty && tx && C > 0 && C <= 4 && T > 0 && overlap >= 0 && nty > 0 && ntx > 0,
                 "tiles_blend: bad arguments");
  tiles_blend_kernel<<<grid_for((long long)H * W), kPwThreads, 0, STREAM>>>(tiles, frame, C, H, W, T, overlap, ty, tx,
                                                                            nty, ntx);
  return check_launch("tiles_blend");
}

extern "C" int b200ir_u8_to_input(const uint8_t* img, float* x, int B, int H, int W, int swap_rb, void* stream) {
  B200IR_REQUIRE(img && x && B > 0 && H > 0 && W > 0, "u8_to_input: bad arguments");
  u8_to_input_kernel<<<grid_for((long long)B * H * W), kPwThreads, 0, STREAM>>>(img, x, B, H * W, swap_rb);
  return check_launch("u8_to_input");
}

extern "C" int b200ir_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                                float beta1, float beta2, float eps, float weight_decay, int step, float grad_scale,
                                float* ema, float ema_decay, void* stream) {
  B200IR_REQUIRE(param && grad && exp_avg && exp_avg_sq && n > 0 && step >= 1, "adam_step: bad arguments");
  B200IR_REQUIRE(((reinterpret_cast<uintptr_t>(param) | reinterpret_cast<uintptr_t>(grad) |
                   reinterpret_cast<uintptr_t>(exp_avg) | reinterpret_cast<uintptr_t>(exp_avg_sq) |
                   reinterpret_cast<uintptr_t>(ema)) & 15) == 0, "adam_step: buffers must be 16-byte aligned");
  const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
  const float step_size = (float)((double)lr / bc1), inv_bc2_sqrt = (float)(1.0 / sqrt(bc2));
  adam_step_kernel<<<grid_for((n + 3) / 4), kPwThreads, 0, STREAM>>>(param, grad, exp_avg, exp_avg_sq, n, beta1, beta2, eps,
                                                                    weight_decay, step_size, inv_bc2_sqrt, grad_scale, ema,
                                                                    ema_decay);
  return check_launch("adam_step");
}

extern "C" int b200ir_minibatch_stddev(const void* x, float* s, void* out, int B, int P, int C, int c_pad, int group,
                                       void* stream) {
  B200IR_REQUIRE(x && s && out && B > 0 && P > 0 && C > 0 && c_pad > C, "minibatch_stddev: bad arguments");
  B200IR_REQUIRE(group >= 1 && group <= 8 && B % group == 0, "minibatch_stddev: batch %d is not divisible by group %d", B,
                 group);
  const int M = B / group;
  mbstd_reduce_kernel<<<M, 256, 0, STREAM>>>((const __half*)x, s, M, group, P * C);
  if (check_launch("minibatch_stddev(reduce)")) return 1;
  mbstd_concat_kernel<<<grid_for((long long)B * P * c_pad), kPwThreads, 0, STREAM>>>((const __half*)x, s, (__half*)out, B, P,
                                                                                    C, c_pad, M);
  return check_launch("minibatch_stddev(concat)");
}

extern "C" int b200ir_f32_to_input(const float* img, float* x, int B, int H, int W, int swap_rb, void* stream) {
  B200IR_REQUIRE(img && x && B > 0 && H > 0 && W > 0, "f32_to_input: bad arguments");
  f32_to_input_kernel<<<grid_for((long long)B * H * W), kPwThreads, 0, STREAM>>>(img, x, B, H * W, swap_rb);
  return check_launch("f32_to_input");
}

extern "C" int b200ir_image_to_u8(const float* x, uint8_t* img, int B, int H, int W, int swap_rb, void* stream) {
  B200IR_REQUIRE(img && x && B > 0 && H > 0 && W > 0, "image_to_u8: bad arguments");
  image_to_u8_kernel<<<grid_for((long long)B * H * W), kPwThreads, 0, STREAM>>>(x, img, B, H * W, swap_rb);
  return check_launch("image_to_u8");
}

extern "C" int b200ir_fir_pad22(const void* in, void* out, int B, int H, int W, int C, int out_h, int out_w,
                                void* stream) {
  B200IR_REQUIRE(in && out && C % 8 == 0 && out_h >= H + 1 && out_w >= W + 1, "fir_pad22: bad arguments");
  {
    FirLaunch a = {};
    a.in = (const __half*)in; a.B = B; a.Hv = H; a.Wv = W;
    a.in_sw = C; a.in_sh = (long long)W * C; a.in_sb = (long long)H * W * C;
    a.C = C; a.OH = H + 1; a.OW = W + 1; a.pad = 2; a.kscale = 0.125f;
    a.out = (__half*)out; a.out_sy = (long long)out_w * C; a.out_sb = (long long)out_h * out_w * C;
    a.post = false;
    const int r = fir_stream_launch(a, STREAM, "fir_pad22");
    if (r >= 0) return r;
  }
  const long long n = (long long)B * (H + 1) * ((W + 4) / 4) * (C / 8);
  fir_pad22_kernel<<<grid_for(n), kPwThreads, 0, STREAM>>>((const __half*)in, (__half*)out, B, H, W, C, out_h, out_w);
  return check_launch("fir_pad22");
}

extern "C" int b200ir_fir_down2(const void* in, void* out, int B, int H, int W, int C, void* stream) {
  B200IR_REQUIRE(in && out && C % 8 == 0 && H % 2 == 0 && W % 2 == 0, "fir_down2: bad arguments");
  {
    const int r = resample_stream_launch(true, (const __half*)in, (__half*)out, B, H, W, C, STREAM);
    if (r >= 0) return r;
  }
  const long long n = (long long)B * (H / 2) * (W / 2) * (C / 8);
  fir_down2_kernel<<<grid_for(n), kPwThreads, 0, STREAM>>>((const __half*)in, (__half*)out, B, H, W, C);
  return check_launch("fir_down2");
}

extern "C" int b200ir_bilinear_up2(const void* in, void* out, int B, int h, int w, int C, void* stream) {
  B200IR_REQUIRE(in && out && C % 8 == 0, "bilinear_up2: bad arguments");
  {
    const int r = resample_stream_launch(false, (const __half*)in, (__half*)out, B, h, w, C, STREAM);
    if (r >= 0) return r;
  }
  const long long n = (long long)B * h * w * (C / 8);
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
  {
    FirLaunch a = {};
    a.in = (const __half*)raw; a.B = B; a.Hv = h2 + 1; a.Wv = w2 + 1;
    a.in_sw = C; a.in_sh = (long long)raw_w * C; a.in_sb = (long long)raw_h * raw_w * C;
    a.C = C; a.OH = h2; a.OW = w2; a.pad = 1; a.kscale = 0.25f;
    a.out = (__half*)out; a.out_sy = (long long)w2 * C; a.out_sb = (long long)h2 * w2 * C;
    a.post = true;
    a.noise = noise; a.noise_sb = noise_stride_b; a.noise_gain = noise_gain; a.bias = bias;
    a.scale = (const __half*)scale; a.shift = (const __half*)shift; a.c_sft = c_sft; a.s_next = s_next;
    const int r = fir_stream_launch(a, STREAM, "upfir_act");
    if (r >= 0) return r;
  }
  B200IR_REQUIRE(w2 % 4 == 0, "upfir_act: w2=%d must be a multiple of 4", w2);
  const long long n = (long long)B * h2 * (w2 / 4) * (C / 8);
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
  const int G = C / (8 * lanes);
  B200IR_REQUIRE(G * lanes * 8 == C && (G == 1 || G == 2 || G == 3), "to_rgb: C=%d unsupported", C);
  dim3 grid((h * w + kRgbPix - 1) / kRgbPix, B);
  if (G == 1)
    to_rgb_kernel<1><<<grid, kPwThreads, 0, STREAM>>>((const __half*)x, h * w, w, C, wrgb, s, bias, skip, rgb, s_next,
                                                      (__half*)xs_out, lanes);
  else if (G == 2)
    to_rgb_kernel<2><<<grid, kPwThreads, 0, STREAM>>>((const __half*)x, h * w, w, C, wrgb, s, bias, skip, rgb, s_next,
                                                      (__half*)xs_out, lanes);
  else
    to_rgb_kernel<3><<<grid, kPwThreads, 0, STREAM>>>((const __half*)x, h * w, w, C, wrgb, s, bias, skip, rgb, s_next,
                                                      (__half*)xs_out, lanes);
  return check_launch("to_rgb");
}

extern "C" int b200ir_rgb_combine(const float* part, int n_parts, const float* bias, const float* skip, float* rgb,
                                  int B, int h, int w, void* stream) {
  B200IR_REQUIRE(part && bias && rgb && n_parts >= 1 && (skip == nullptr || (h % 2 == 0 && w % 2 == 0)),
                 "rgb_combine: bad arguments");
  if (w % 4 == 0 && ((reinterpret_cast<uintptr_t>(part) | reinterpret_cast<uintptr_t>(rgb)) & 15) == 0 &&
      (long long)B * h * (w / 4) < (1LL << 31))
    rgb_combine4_kernel<<<grid_for((long long)B * h * (w / 4)), kPwThreads, 0, STREAM>>>(part, n_parts, bias, skip, rgb, B,
                                                                                        h, w);
  else
    rgb_combine_kernel<<<grid_for((long long)B * h * w), kPwThreads, 0, STREAM>>>(part, n_parts, bias, skip, rgb, B, h, w);
  return check_launch("rgb_combine");
}

extern "C" int b200ir_rgb_wmod(const float* w, const float* s, float* wm, int B, int C, void* stream) {
  B200IR_REQUIRE(w && wm && B > 0 && C > 0, "rgb_wmod: bad arguments");
  rgb_wmod_kernel<<<grid_for((long long)B * 3 * C), kPwThreads, 0, STREAM>>>(w, s, wm, B, C);
  return check_launch("rgb_wmod");
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

extern "C" int b200ir_mod_linear_multi(const float* latent, int L, int F, const b200ir_mod_layer* layers_dev,
                                       int n_layers, int max_cin, float wscale, int B, void* stream) {
  B200IR_REQUIRE(latent && layers_dev && n_layers > 0 && max_cin > 0 && F <= kRdMaxC && F % 4 == 0,
                 "mod_linear_multi: bad arguments (F=%d)", F);
  dim3 grid((max_cin + kRdOut - 1) / kRdOut, n_layers, (B + kRdBatch - 1) / kRdBatch);
  mod_linear_multi_kernel<<<grid, kRdOut, 0, STREAM>>>(latent, L, F, layers_dev, wscale, B);
  return check_launch("mod_linear_multi");
}

extern "C" int b200ir_demod_multi(const b200ir_demod_layer* layers_dev, int n_layers, int max_cout, int B,
                                  void* stream) {
  B200IR_REQUIRE(layers_dev && n_layers > 0 && max_cout > 0, "demod_multi: bad arguments");
  dim3 grid((max_cout + kRdOut - 1) / kRdOut, n_layers, (B + kRdBatch - 1) / kRdBatch);
  demod_multi_kernel<<<grid, kRdOut, 0, STREAM>>>(layers_dev, B);
  return check_launch("demod_multi");
}

extern "C" int b200ir_style_mlp(const float* z, const float* w, const float* bias, float* out, int B, int F, int n_layers,
                                float lr_mul, void* stream) {
  B200IR_REQUIRE(z && w && bias && out && B > 0 && F > 0 && n_layers >= 0, "style_mlp: bad arguments");
  style_mlp_kernel<<<B, 256, 2 * F * sizeof(float), STREAM>>>(z, w, bias, out, F, n_layers, lr_mul);
  return check_launch("style_mlp");
}

extern "C" int b200ir_nhwc_to_nchw_f32(const void* in, float* out, int B, int P, int C, void* stream) {
  B200IR_REQUIRE(in && out, "nhwc_to_nchw_f32: bad arguments");
  const long long n = (long long)B * P * C;
  nhwc_to_nchw_f32_kernel<<<grid_for(n), kPwThreads, 0, STREAM>>>((const __half*)in, out, B, P, C);
  return check_launch("nhwc_to_nchw_f32");
}
