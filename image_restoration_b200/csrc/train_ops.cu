// Memory-bound kernels of the training step (BASELINE config 5) that the forward path does not have:
// the adjoints of the StyleGAN2 decoder's pointwise stages (SFT, modulation, noise + FusedLeakyReLU, ToRGB and its skip
// up-sampling) together with the per-sample channel reductions that give the style gradients, the input gradient of the
// first 1x1 conv (the discriminator must pass d(score)/d(image) back to the generator), the L1 / softplus losses with
// their gradients, and the weight packing entry point.  Reference semantics: include/b200ir.h next to each entry point
// (stylegan2_ocr_arch.py:239-279, 323-333, 357-374; gfpganv1_ocr_arch.py:108-129; losses/losses.py:81-106, 404-419).
//
// Layout as everywhere on the path: activations and their gradients NHWC fp16, 8 channels (16 bytes) per thread access;
// tables (styles, demodulation, their gradients) fp32 [B][C].  The reductions follow lrelu_bias_bwd_kernel (backward.cu):
// every thread owns one group of 8 channels and a strided set of pixels of ONE image, partial sums stay in registers until
// one shared-memory reduction and one fp32 atomic per (CTA, channel).
#include "host_common.h"
#include "ptx.cuh"

namespace b200ir {

constexpr int kTrThreads = 256;
constexpr float kTrSqrt2 = 1.4142135623730951f;

__device__ __forceinline__ void tr_unpack(const uint4& q, float* f) {
  const __half2* h = reinterpret_cast<const __half2*>(&q);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 x = __half22float2(h[i]);
    f[2 * i] = x.x;
    f[2 * i + 1] = x.y;
  }
}
__device__ __forceinline__ uint4 tr_pack(const float* f) {
  uint4 q;
  __half2* h = reinterpret_cast<__half2*>(&q);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = f2h2_sat(f[2 * i], f[2 * i + 1]);
  return q;
}

// Thread layout of the per-image reductions: grid (pixel chunks, channel chunks, B); inside a CTA `groups` channel groups
// x `lanes` pixel lanes.
struct TrLayout {
  int groups, chunks, lanes;
  unsigned grid_x;
};
static TrLayout tr_layout(int C, long long P, int sms, int B) {
  TrLayout l;
  const int row_groups = C / 8;
  l.groups = kTrThreads;
  while (row_groups % l.groups) l.groups >>= 1;
  l.chunks = row_groups / l.groups;
  l.lanes = kTrThreads / l.groups;
  long long gx = (P + l.lanes - 1) / l.lanes;
  // about one wave of 8 resident CTAs per SM over the whole launch, at least one CTA per image and channel chunk
  long long cap = (8LL * sms + (long long)l.chunks * B - 1) / ((long long)l.chunks * B);
  if (cap < 1) cap = 1;
  if (gx > cap) gx = cap;
  l.grid_x = (unsigned)gx;
  return l;
}

// sums acc[8] over the pixel lanes of the CTA and adds the result to dst[c] (c = channel inside this CTA's chunk)
__device__ __forceinline__ void tr_reduce_add(const float* acc, float* dst, int groups, int lanes, float (*part)[9]) {
#pragma unroll
  for (int j = 0; j < 8; ++j) part[threadIdx.x][j] = acc[j];
  __syncthreads();
  for (int c = threadIdx.x; c < groups * 8; c += kTrThreads) {
    const int cg = c >> 3, cj = c & 7;
    float s = 0.f;
    for (int l = 0; l < lanes; ++l) s += part[l * groups + cg][cj];
    atomicAdd(dst + c, s);
  }
}

// ------------------------------------------------------------------------------------------ SFT + modulation, forward
__global__ void __launch_bounds__(kTrThreads) sft_mod_kernel(const uint4* __restrict__ a, const uint4* __restrict__ scale,
                                                             const uint4* __restrict__ shift, int sft_groups,
                                                             const float* __restrict__ s_next, uint4* __restrict__ out,
                                                             long long n, long long P, int row_groups) {
  for (long long idx = (long long)blockIdx.x * kTrThreads + threadIdx.x; idx < n; idx += (long long)gridDim.x * kTrThreads) {
    // 32-bit index arithmetic (the launcher keeps n below 2^32): 64-bit divisions per element would make this ALU-bound
    const unsigned pix = (unsigned)idx / (unsigned)row_groups;
    const int g = (int)((unsigned)idx - pix * (unsigned)row_groups);
    const long long b = pix / (unsigned)P;
    float v[8];
    tr_unpack(__ldcs(a + idx), v);
    const int gs = g - (row_groups - sft_groups);
    if (scale != nullptr && gs >= 0) {
      float sc[8], sh[8];
      tr_unpack(__ldcs(scale + (long long)pix * sft_groups + gs), sc);
      tr_unpack(__ldcs(shift + (long long)pix * sft_groups + gs), sh);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = fmaf(v[j], sc[j], sh[j]);
    }
    if (s_next != nullptr) {
      const float4 s0 = __ldg(reinterpret_cast<const float4*>(s_next + (b * row_groups + g) * 8));
      const float4 s1 = __ldg(reinterpret_cast<const float4*>(s_next + (b * row_groups + g) * 8) + 1);
      v[0] *= s0.x; v[1] *= s0.y; v[2] *= s0.z; v[3] *= s0.w;
      v[4] *= s1.x; v[5] *= s1.y; v[6] *= s1.z; v[7] *= s1.w;
    }
    out[idx] = tr_pack(v);
  }
}

// ------------------------------------------------------------------------------------------ SFT + modulation, backward
// g = gradient w.r.t. the modulated conv input (the dgrad GEMM's output).  o = SFT output (recomputed from a, scale, shift):
//   ds[b][c] += sum_p g * o;   do = g * s_next;   da = do (* scale on the SFT channels);   dscale = do * a;   dshift = do
__global__ void __launch_bounds__(kTrThreads) sft_mod_bwd_kernel(const uint4* __restrict__ g, const uint4* __restrict__ a,
                                                                 long long a_sb_groups, const uint4* __restrict__ scale,
                                                                 const uint4* __restrict__ shift, int sft_groups,
                                                                 const float* __restrict__ s_next, uint4* __restrict__ da,
                                                                 int accumulate, uint4* __restrict__ dscale,
                                                                 uint4* __restrict__ dshift, float* __restrict__ ds, long long P,
                                                                 int groups, int row_groups) {
  __shared__ float part[kTrThreads][9];
  const int gi = threadIdx.x % groups, lane = threadIdx.x / groups, lanes = kTrThreads / groups;
  const int grp = blockIdx.y * groups + gi;  // channel group inside the row
  const long long b = blockIdx.z;
  const int gs = grp - (row_groups - sft_groups);
  const bool sft = scale != nullptr && gs >= 0;
  float sn[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) sn[j] = (s_next != nullptr) ? __ldg(s_next + (b * row_groups + grp) * 8 + j) : 1.f;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  for (long long p = (long long)blockIdx.x * lanes + lane; p < P; p += (long long)gridDim.x * lanes) {
    const long long pix = b * P + p;
    const long long i = pix * row_groups + grp;
    float gv[8], av[8];
    tr_unpack(__ldcs(g + i), gv);
    tr_unpack(__ldg(a + b * a_sb_groups + p * row_groups + grp), av);
    float dov[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) dov[j] = gv[j] * sn[j];
    if (sft) {
      float sc[8], sh[8], t[8];
      tr_unpack(__ldcs(scale + pix * sft_groups + gs), sc);
      tr_unpack(__ldcs(shift + pix * sft_groups + gs), sh);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = fmaf(gv[j], fmaf(av[j], sc[j], sh[j]), acc[j]);
#pragma unroll
      for (int j = 0; j < 8; ++j) t[j] = dov[j] * av[j];
      dscale[pix * sft_groups + gs] = tr_pack(t);
      dshift[pix * sft_groups + gs] = tr_pack(dov);
#pragma unroll
      for (int j = 0; j < 8; ++j) dov[j] *= sc[j];
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = fmaf(gv[j], av[j], acc[j]);
    }
    if (da != nullptr) {
      if (accumulate) {
        float old[8];
        tr_unpack(da[i], old);
#pragma unroll
        for (int j = 0; j < 8; ++j) dov[j] += old[j];
      }
      da[i] = tr_pack(dov);
    }
  }
  if (ds != nullptr) tr_reduce_add(acc, ds + (b * row_groups + (long long)blockIdx.y * groups) * 8, groups, lanes, part);
}

// ------------------------------------------------------------------------------------------ StyleConv tail, backward
// a = lrelu(y + gain * noise + bias) * sqrt 2 (saved output);  dz = da * sqrt 2 * (a > 0 ? 1 : 0.2);
// y reconstructed from a;  dd[b][c] += sum_p dz * y;  out = dz * oscale[b][c] * mul
// kParams (fix_decoder = false): also db[b][c] += sum_p dz (activate.bias) and dn[b][c] += sum_p dz * noise (StyleConv.weight)
template <bool kParams>
__global__ void __launch_bounds__(kTrThreads) style_act_bwd_kernel(const uint4* __restrict__ da, const uint4* __restrict__ a,
                                                                   const float* __restrict__ noise, long long noise_sb,
                                                                   const float* __restrict__ noise_gain,
                                                                   const float* __restrict__ bias,
                                                                   const float* __restrict__ oscale, float mul,
                                                                   uint4* __restrict__ out, float* __restrict__ dd,
                                                                   float* __restrict__ db, float* __restrict__ dn, long long P,
                                                                   int groups, int row_groups) {
  __shared__ float part[kTrThreads][9];
  const int gi = threadIdx.x % groups, lane = threadIdx.x / groups, lanes = kTrThreads / groups;
  const int grp = blockIdx.y * groups + gi;
  const long long b = blockIdx.z;
  float bs[8], os[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    bs[j] = (bias != nullptr) ? __ldg(bias + grp * 8 + j) : 0.f;
    os[j] = ((oscale != nullptr) ? __ldg(oscale + (b * row_groups + grp) * 8 + j) : 1.f) * mul;
  }
  const float gain = (noise != nullptr) ? __ldg(noise_gain) : 0.f;
  float acc[8], accb[8], accn[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = accb[j] = accn[j] = 0.f;
  for (long long p = (long long)blockIdx.x * lanes + lane; p < P; p += (long long)gridDim.x * lanes) {
    const long long i = (b * P + p) * row_groups + grp;
    float dv[8], av[8], o[8];
    tr_unpack(__ldcs(da + i), dv);
    tr_unpack(__ldcs(a + i), av);
    const float nraw = (noise != nullptr) ? __ldg(noise + b * noise_sb + p) : 0.f;
    const float nz = gain * nraw;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const bool pos = av[j] > 0.f;
      const float dz = dv[j] * (pos ? kTrSqrt2 : 0.2f * kTrSqrt2);
      const float y = av[j] * (pos ? (1.f / kTrSqrt2) : (1.f / (0.2f * kTrSqrt2))) - nz - bs[j];
      acc[j] = fmaf(dz, y, acc[j]);
      if (kParams) {
        accb[j] += dz;
        accn[j] = fmaf(dz, nraw, accn[j]);
      }
      o[j] = dz * os[j];
    }
    out[i] = tr_pack(o);
  }
  const long long tab = (b * row_groups + (long long)blockIdx.y * groups) * 8;
  if (dd != nullptr) tr_reduce_add(acc, dd + tab, groups, lanes, part);
  if (kParams) {
    __syncthreads();
    tr_reduce_add(accb, db + tab, groups, lanes, part);
    __syncthreads();
    tr_reduce_add(accn, dn + tab, groups, lanes, part);
  }
}

// ------------------------------------------------------------------------------------------ ToRGB, backward
// t[c] = sum_o drgb[b][o][p] * w[o][c];  da (+)= s[b][c] * t;  ds[b][c] += sum_p a * t
// kParams (fix_decoder = false): also R[b][o][c] += sum_p drgb[b][o][p] * a[b][p][c]  (weight gradient of the 1x1 conv before
// its modulation: dw[o][c] = sum_b s[b][c] R[b][o][c])
template <bool kParams>
__global__ void __launch_bounds__(kTrThreads) to_rgb_bwd_kernel(const float* __restrict__ drgb, const uint4* __restrict__ a,
                                                                const float* __restrict__ w, const float* __restrict__ s,
                                                                uint4* __restrict__ da, int accumulate, float* __restrict__ ds,
                                                                float* __restrict__ R, long long P, int groups, int row_groups) {
  __shared__ float part[kTrThreads][9];
  const int gi = threadIdx.x % groups, lane = threadIdx.x / groups, lanes = kTrThreads / groups;
  const int grp = blockIdx.y * groups + gi;
  const long long b = blockIdx.z;
  const int C = row_groups * 8;
  float w0[8], w1[8], w2[8], sv[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    w0[j] = __ldg(w + grp * 8 + j);
    w1[j] = __ldg(w + C + grp * 8 + j);
    w2[j] = __ldg(w + 2 * C + grp * 8 + j);
    sv[j] = (s != nullptr) ? __ldg(s + (b * row_groups + grp) * 8 + j) : 1.f;
  }
  float acc[8], r0[8], r1[8], r2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = r0[j] = r1[j] = r2[j] = 0.f;
  const float* dr = drgb + b * 3 * P;
  for (long long p = (long long)blockIdx.x * lanes + lane; p < P; p += (long long)gridDim.x * lanes) {
    const long long i = (b * P + p) * row_groups + grp;
    const float d0 = __ldg(dr + p), d1 = __ldg(dr + P + p), d2 = __ldg(dr + 2 * P + p);
    float av[8], o[8];
    tr_unpack(__ldcs(a + i), av);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float t = fmaf(d0, w0[j], fmaf(d1, w1[j], d2 * w2[j]));
      acc[j] = fmaf(av[j], t, acc[j]);
      o[j] = t * sv[j];
      if (kParams) {
        r0[j] = fmaf(d0, av[j], r0[j]);
        r1[j] = fmaf(d1, av[j], r1[j]);
        r2[j] = fmaf(d2, av[j], r2[j]);
      }
    }
    if (da != nullptr) {
      if (accumulate) {
        float old[8];
        tr_unpack(da[i], old);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] += old[j];
      }
      da[i] = tr_pack(o);
    }
  }
  if (ds != nullptr) tr_reduce_add(acc, ds + (b * row_groups + (long long)blockIdx.y * groups) * 8, groups, lanes, part);
  if (kParams) {
    float* Rb = R + b * 3 * C + (long long)blockIdx.y * groups * 8;
    __syncthreads();
    tr_reduce_add(r0, Rb, groups, lanes, part);
    __syncthreads();
    tr_reduce_add(r1, Rb + C, groups, lanes, part);
    __syncthreads();
    tr_reduce_add(r2, Rb + 2 * C, groups, lanes, part);
  }
}

// out[c] += sum over b and p of x[b][c][p] (fp32 NCHW planes): bias gradient of ToRGB (to_rgb.bias, stylegan2_ocr_arch.py:349)
__global__ void __launch_bounds__(kTrThreads) plane_sums_kernel(const float* __restrict__ x, float* __restrict__ out, int Cn,
                                                                long long P) {
  __shared__ float red[kTrThreads / 32];
  const long long plane = blockIdx.y;             // b * Cn + c
  float acc = 0.f;
  const float* xp = x + plane * P;
  for (long long i = (long long)blockIdx.x * kTrThreads + threadIdx.x; i < P; i += (long long)gridDim.x * kTrThreads) acc += xp[i];
  for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float sacc = 0.f;
    for (int i = 0; i < kTrThreads / 32; ++i) sacc += red[i];
    atomicAdd(out + (plane % Cn), sacc);
  }
}

// ------------------------------------------------------------------------------------------ decoder parameter gradients
// (fix_decoder = false).  The per-image tables the pointwise adjoints leave behind are folded over the batch here.
// out[j] = scale * sum_b in[b][j] * (mul ? mul[b][j % m] : 1): bias gradients (sum of db / ds over the batch), the ToRGB
// weight (R[b][o][c] weighted by s[b][c]), ConstantInput (g0[b][p][c] weighted by s[b][c], fp16 input), the noise gain.
template <typename T>
__global__ void __launch_bounds__(kTrThreads) table_colsum_kernel(const T* __restrict__ in, const float* __restrict__ mul, int m,
                                                                  float scale, float* __restrict__ out, int B, long long n) {
  const long long j = (long long)blockIdx.x * kTrThreads + threadIdx.x;
  if (j >= n) return;
  const int jm = (mul != nullptr) ? (int)(j % m) : 0;
  float acc = 0.f;
  for (int b = 0; b < B; ++b) {
    const float v = (float)in[(long long)b * n + j];
    acc = (mul != nullptr) ? fmaf(v, __ldg(mul + (long long)b * m + jm), acc) : acc + v;
  }
  out[j] = scale * acc;
}

// dw[ci][f] = wscale * sum_b ds[b][ci] * latent[b][lat_idx][f]: weight gradient of a modulation EqualLinear
// (stylegan2_ocr_arch.py:233-234); one CTA per ci.
__global__ void __launch_bounds__(kTrThreads) mod_linear_wgrad_kernel(const float* __restrict__ ds, const float* __restrict__ lat,
                                                                      float wscale, float* __restrict__ dw, int L, int F,
                                                                      int lat_idx, int B, int cin) {
  const int ci = blockIdx.x;
  for (int f = threadIdx.x; f < F; f += kTrThreads) {
    float acc = 0.f;
    for (int b = 0; b < B; ++b) acc = fmaf(__ldg(ds + (long long)b * cin + ci), __ldg(lat + ((long long)b * L + lat_idx) * F + f), acc);
    dw[(long long)ci * F + f] = wscale * acc;
  }
}

// Weight gradient of ModulatedConv2d (stylegan2_ocr_arch.py:239-279) assembled in the reference layout:
//   dw[co][ci][k] = scale * G[co][k][ci] - scale^2 * W[co][ci][k] * sum_b dd[b][co] d[b][co]^2 s[b][ci]^2
// G = the tap-major result of the weight-gradient GEMM on (modulated input, dy * d); `transposed`: G is [ci][k][co] (the
// up-sampling conv, whose GEMM runs with the roles of input and output exchanged).  The second term is the dependence of
// the demodulation table d on W.  One CTA per co.
__global__ void __launch_bounds__(kTrThreads) modconv_wgrad_kernel(const float* __restrict__ G, int transposed,
                                                                   const float* __restrict__ W, const float* __restrict__ s,
                                                                   const float* __restrict__ dd, const float* __restrict__ d,
                                                                   float scale, float* __restrict__ dw, int B, int cin, int cout,
                                                                   int taps) {
  extern __shared__ float t[];  // [B]
  const int co = blockIdx.x;
  for (int b = threadIdx.x; b < B; b += kTrThreads) {
    const float dv = d[(long long)b * cout + co];
    t[b] = dd[(long long)b * cout + co] * dv * dv;
  }
  __syncthreads();
  for (int ci = threadIdx.x; ci < cin; ci += kTrThreads) {
    float m = 0.f;
    for (int b = 0; b < B; ++b) {
      const float sv = __ldg(s + (long long)b * cin + ci);
      m = fmaf(t[b], sv * sv, m);
    }
    m *= scale * scale;
    for (int k = 0; k < taps; ++k) {
      const float g = transposed ? G[((long long)ci * taps + k) * cout + co] : G[((long long)co * taps + k) * cin + ci];
      const long long o = ((long long)co * cin + ci) * taps + k;
      dw[o] = scale * g - W[o] * m;
    }
  }
}

// Low-channel weight gradients run on pixel-folded views (ops.conv_wgrad): f horizontally adjacent pixels of x and dy are read as
// f * C channels of one pixel, so a 32 -> 32 layer becomes a 128 -> 128 GEMM over a quarter of the pixels (the tensor pipe
// charges a full 128 x 64 tile whatever the channel count).  G[(s_o, co)][kh][kw'][(s_i, ci)] then holds every pair of
// sub-pixels; the original tap kw collects the pairs with f * (kw' - 1) + s_i - s_o = kw - 1.
__global__ void __launch_bounds__(kTrThreads) wgrad_unfold_kernel(const float* __restrict__ G, float* __restrict__ dw, int f,
                                                                  int cin, int cout) {
  const int n = cout * 9 * cin;
  const int idx = blockIdx.x * kTrThreads + threadIdx.x;
  if (idx >= n) return;
  const int ci = idx % cin, tap = (idx / cin) % 9, co = idx / (9 * cin);
  const int kh = tap / 3, d = tap % 3 - 1;
  const long long ldg = 9LL * f * cin;      // elements per folded output channel
  float acc = 0.f;
  for (int so = 0; so < f; ++so) {
    const int t = so + d;
    int dp, si;
    if (t >= 0 && t < f) {
      dp = 0, si = t;
    } else if (t == f) {
      dp = 1, si = 0;
    } else {
      dp = -1, si = f - 1;
    }
    acc += G[(long long)(so * cout + co) * ldg + (long long)(kh * 3 + dp + 1) * f * cin + si * cin + ci];
  }
  dw[idx] = acc;
}

// adjoint of upfirdn2d(skip, FIR * 4, up = 2, pad = (2, 1)) on fp32 planes: per axis
//   out[k] = .25 d[2k-1] + .75 d[2k] + .75 d[2k+1] + .25 d[2k+2]   (d = 0 outside)
__global__ void __launch_bounds__(kTrThreads) rgb_up_adjoint_kernel(const float* __restrict__ d, float* __restrict__ out,
                                                                    long long n, int h, int w) {
  const long long idx = (long long)blockIdx.x * kTrThreads + threadIdx.x;
  if (idx >= n) return;
  const int x = (int)(idx % w);
  const int y = (int)((idx / w) % h);
  const long long pl = idx / ((long long)w * h);
  const float* dp = d + pl * 4 * h * w;
  const int H2 = 2 * h, W2 = 2 * w;
  float s = 0.f;
#pragma unroll
  for (int ty = 0; ty < 4; ++ty) {
    const int yy = 2 * y - 1 + ty;
    if (yy < 0 || yy >= H2) continue;
    const float wy = (ty == 0 || ty == 3) ? 0.25f : 0.75f;
#pragma unroll
    for (int tx = 0; tx < 4; ++tx) {
      const int xx = 2 * x - 1 + tx;
      if (xx < 0 || xx >= W2) continue;
      s += wy * ((tx == 0 || tx == 3) ? 0.25f : 0.75f) * __ldg(dp + (long long)yy * W2 + xx);
    }
  }
  out[idx] = s;
}

// ------------------------------------------------------------------------------------------ style gradients (tables)
// ds[b][ci] -= scale2 * s[b][ci] * sum_co dd[b][co] * d[b][co]^2 * wsq[co][ci]   (one CTA per image)
__global__ void __launch_bounds__(kTrThreads) demod_bwd_kernel(float* __restrict__ ds, const float* __restrict__ s,
                                                               const float* __restrict__ dd, const float* __restrict__ d,
                                                               const float* __restrict__ wsq, float scale2, int cin, int cout) {
  extern __shared__ float t[];  // [cout]
  const long long b = blockIdx.x;
  for (int co = threadIdx.x; co < cout; co += kTrThreads) {
    const float dv = d[b * cout + co];
    t[co] = dd[b * cout + co] * dv * dv;
  }
  __syncthreads();
  for (int ci = threadIdx.x; ci < cin; ci += kTrThreads) {
    float acc = 0.f;
    for (int co = 0; co < cout; ++co) acc = fmaf(t[co], __ldg(wsq + (long long)co * cin + ci), acc);
    ds[b * cin + ci] -= scale2 * s[b * cin + ci] * acc;
  }
}

// dlat[b][lat_idx][f] += wscale * sum_ci ds[b][ci] * w[ci][f]   (one CTA per image)
__global__ void __launch_bounds__(kTrThreads) mod_linear_bwd_kernel(const float* __restrict__ ds, const float* __restrict__ w,
                                                                    float wscale, float* __restrict__ dlat, int L, int F,
                                                                    int lat_idx, int cin) {
  extern __shared__ float t[];  // [cin]
  const long long b = blockIdx.x;
  for (int ci = threadIdx.x; ci < cin; ci += kTrThreads) t[ci] = ds[b * cin + ci];
  __syncthreads();
  for (int f = threadIdx.x; f < F; f += kTrThreads) {
    float acc = 0.f;
    for (int ci = 0; ci < cin; ++ci) acc = fmaf(t[ci], __ldg(w + (long long)ci * F + f), acc);
    dlat[(b * L + lat_idx) * F + f] += wscale * acc;
  }
}

// ------------------------------------------------------------------------------------------ first conv, input gradient
// dx[b][k][p] (+)= sum_c dz[b][p][c] * w[c][k], k = 0..2: one thread per pixel (coalesced fp32 NCHW stores)
__global__ void __launch_bounds__(kTrThreads) first_conv_dgrad_kernel(const uint4* __restrict__ dz, const float* __restrict__ w,
                                                                      float* __restrict__ dx, int accumulate, long long n_pix,
                                                                      int HW, int cout) {
  extern __shared__ float sw[];  // [cout * 3]
  for (int i = threadIdx.x; i < cout * 3; i += kTrThreads) sw[i] = w[i];
  __syncthreads();
  const long long idx = (long long)blockIdx.x * kTrThreads + threadIdx.x;
  if (idx >= n_pix) return;
  const int groups = cout / 8;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f;
  for (int g = 0; g < groups; ++g) {
    float v[8];
    tr_unpack(__ldcs(dz + idx * groups + g), v);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float* wc = sw + (g * 8 + j) * 3;
      a0 = fmaf(v[j], wc[0], a0);
      a1 = fmaf(v[j], wc[1], a1);
      a2 = fmaf(v[j], wc[2], a2);
    }
  }
  const long long b = idx / HW, p = idx - b * HW;
  float* o = dx + b * 3 * HW + p;
  if (accumulate) {
    a0 += o[0];
    a1 += o[HW];
    a2 += o[2 * (long long)HW];
  }
  o[0] = a0;
  o[HW] = a1;
  o[2 * (long long)HW] = a2;
}

// ------------------------------------------------------------------------------------------ toRGB heads <-> fp32 NCHW
// head [B][P][cpad] fp16 (channels 0..2 = the image) -> rgb fp32 [B][3][P], and the adjoint (zero fill of channels 3..)
__global__ void __launch_bounds__(kTrThreads) head_to_nchw_kernel(const __half* __restrict__ head, float* __restrict__ rgb,
                                                                  long long n_pix, long long P, int cpad) {
  const long long idx = (long long)blockIdx.x * kTrThreads + threadIdx.x;
  if (idx >= n_pix) return;
  const long long b = idx / P, p = idx - b * P;
  const uint2 q = *reinterpret_cast<const uint2*>(head + idx * cpad);
  const __half2* h = reinterpret_cast<const __half2*>(&q);
  const float2 f0 = __half22float2(h[0]), f1 = __half22float2(h[1]);
  float* o = rgb + b * 3 * P + p;
  o[0] = f0.x;
  o[P] = f0.y;
  o[2 * P] = f1.x;
}
__global__ void __launch_bounds__(kTrThreads) nchw_to_head_kernel(const float* __restrict__ drgb, __half* __restrict__ dhead,
                                                                  long long n_pix, long long P, int cpad) {
  const long long idx = (long long)blockIdx.x * kTrThreads + threadIdx.x;
  if (idx >= n_pix) return;
  const long long b = idx / P, p = idx - b * P;
  const float* d = drgb + b * 3 * P + p;
  uint4 q = make_uint4(0u, 0u, 0u, 0u);
  __half2* h = reinterpret_cast<__half2*>(&q);
  h[0] = f2h2_sat(d[0], d[P]);
  h[1] = f2h2_sat(d[2 * P], 0.f);
  uint4* o = reinterpret_cast<uint4*>(dhead + idx * cpad);
  o[0] = q;
  for (int k = 1; k < cpad / 8; ++k) o[k] = make_uint4(0u, 0u, 0u, 0u);
}

// ------------------------------------------------------------------------------------------ losses
// L1Loss(reduction='mean') * weight: loss[0] += weight / n * sum |x - t|;  grad = gscale * sign(x - t) (0 at ties, as torch)
__global__ void __launch_bounds__(kTrThreads) l1_loss_kernel(const float* __restrict__ x, const float* __restrict__ t,
                                                             long long n, float lscale, float gscale, float* __restrict__ loss,
                                                             float* __restrict__ grad) {
  __shared__ float red[kTrThreads / 32];
  float acc = 0.f;
  for (long long i = (long long)blockIdx.x * kTrThreads + threadIdx.x; i < n; i += (long long)gridDim.x * kTrThreads) {
    const float d = __ldcs(x + i) - __ldcs(t + i);
    acc += fabsf(d);
    if (grad != nullptr) grad[i] = d > 0.f ? gscale : (d < 0.f ? -gscale : 0.f);
  }
  for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int i = 0; i < kTrThreads / 32; ++i) s += red[i];
    atomicAdd(loss, s * lscale);
  }
}

// GANLoss('wgan_softplus'): loss[0] += weight / n * sum softplus(sign * pred);  dpred = gscale * sign * sigmoid(sign * pred)
__global__ void __launch_bounds__(kTrThreads) softplus_loss_kernel(const __half* __restrict__ pred, int n, int stride, float sign,
                                                                   float lscale, float gscale, float* __restrict__ loss,
                                                                   __half* __restrict__ dpred) {
  __shared__ float red[kTrThreads / 32];
  float acc = 0.f;
  for (int i = threadIdx.x; i < n; i += kTrThreads) {
    const float v = sign * __half2float(pred[(long long)i * stride]);
    acc += fmaxf(v, 0.f) + log1pf(expf(-fabsf(v)));  // softplus without overflow
    if (dpred != nullptr) dpred[(long long)i * stride] = __float2half_rn(gscale * sign / (1.f + expf(-v)));
  }
  for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int i = 0; i < kTrThreads / 32; ++i) s += red[i];
    atomicAdd(loss, s * lscale);
  }
}

// ------------------------------------------------------------------------------------------ perceptual loss pieces (VGG19)
// nn.ReLU + nn.MaxPool2d(2, 2) of the VGG feature extractor (vgg_arch.py:100-125) on a PRE-activation tensor z (the layers the
// perceptual loss taps are read before their ReLU, losses.py:250-356): out = max(0, max over the 2 x 2 window).
__global__ void __launch_bounds__(kTrThreads) maxpool2_relu_kernel(const uint4* __restrict__ z, uint4* __restrict__ out,
                                                                   long long n, int oh, int ow, int groups) {
  for (long long idx = (long long)blockIdx.x * kTrThreads + threadIdx.x; idx < n; idx += (long long)gridDim.x * kTrThreads) {
    const unsigned pix = (unsigned)idx / (unsigned)groups;      // 32-bit index arithmetic: the launcher keeps n below 2^32
    const int g = (int)((unsigned)idx - pix * (unsigned)groups);
    const unsigned r = pix / (unsigned)ow;
    const int x = (int)(pix - r * (unsigned)ow);
    const long long b = r / (unsigned)oh;
    const int y = (int)(r - (unsigned)b * (unsigned)oh);
    const long long base = ((b * 2 * oh + 2 * y) * (2 * ow) + 2 * x) * groups + g;
    float m[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) m[j] = 0.f;
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        float v[8];
        tr_unpack(__ldcs(z + base + ((long long)dy * 2 * ow + dx) * groups), v);
#pragma unroll
        for (int j = 0; j < 8; ++j) m[j] = fmaxf(m[j], v[j]);
      }
    out[idx] = tr_pack(m);
  }
}

// backward: dz = add + (dpool routed to the first maximum of its window, if that maximum is positive); one thread per
// pooled position and channel group writes the four dz entries of its window.
__global__ void __launch_bounds__(kTrThreads) maxpool2_relu_bwd_kernel(const uint4* __restrict__ z, const uint4* __restrict__ dpool,
                                                                       const uint4* __restrict__ add, uint4* __restrict__ dz,
                                                                       long long n, int oh, int ow, int groups) {
  for (long long idx = (long long)blockIdx.x * kTrThreads + threadIdx.x; idx < n; idx += (long long)gridDim.x * kTrThreads) {
    const unsigned pix = (unsigned)idx / (unsigned)groups;      // 32-bit index arithmetic: the launcher keeps n below 2^32
    const int g = (int)((unsigned)idx - pix * (unsigned)groups);
    const unsigned r = pix / (unsigned)ow;
    const int x = (int)(pix - r * (unsigned)ow);
    const long long b = r / (unsigned)oh;
    const int y = (int)(r - (unsigned)b * (unsigned)oh);
    const long long base = ((b * 2 * oh + 2 * y) * (2 * ow) + 2 * x) * groups + g;
    float v[4][8], d[8];
    if (dpool != nullptr) {
      tr_unpack(__ldcs(dpool + idx), d);
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) d[j] = 0.f;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) tr_unpack(__ldcs(z + base + ((long long)(k >> 1) * 2 * ow + (k & 1)) * groups), v[k]);
    int arg[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float m = v[0][j];
      int a = 0;
#pragma unroll
      for (int k = 1; k < 4; ++k)
        if (v[k][j] > m) m = v[k][j], a = k;   // first maximum in row-major order, as nn.MaxPool2d
      arg[j] = m > 0.f ? a : -1;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const long long o = base + ((long long)(k >> 1) * 2 * ow + (k & 1)) * groups;
      float w[8];
      if (add != nullptr) {
        tr_unpack(__ldcs(add + o), w);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) w[j] = 0.f;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) w[j] += (arg[j] == k) ? d[j] : 0.f;
      dz[o] = tr_pack(w);
    }
  }
}

// L1Loss(mean) * weight on fp16 tensors (VGG features): loss[0] += weight / n * sum |x - t|; grad (fp16) = gscale * sign(x - t)
__global__ void __launch_bounds__(kTrThreads) l1_loss_f16_kernel(const uint4* __restrict__ x, const uint4* __restrict__ t,
                                                                 long long n8, float lscale, float gscale, float* __restrict__ loss,
                                                                 uint4* __restrict__ grad) {
  __shared__ float red[kTrThreads / 32];
  float acc = 0.f;
  for (long long i = (long long)blockIdx.x * kTrThreads + threadIdx.x; i < n8; i += (long long)gridDim.x * kTrThreads) {
    float a[8], b[8], gq[8];
    tr_unpack(__ldcs(x + i), a);
    tr_unpack(__ldcs(t + i), b);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float d = a[j] - b[j];
      acc += fabsf(d);
      gq[j] = d > 0.f ? gscale : (d < 0.f ? -gscale : 0.f);
    }
    if (grad != nullptr) grad[i] = tr_pack(gq);
  }
  for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int i = 0; i < kTrThreads / 32; ++i) s += red[i];
    atomicAdd(loss, s * lscale);
  }
}

// ------------------------------------------------------------------------------------------ R1 penalty pieces
// sum of squares of an fp32 array: out[0] += scale * sum x^2   (|grad_x D|^2 of the R1 penalty, losses.py:492-506)
__global__ void __launch_bounds__(kTrThreads) sum_squares_kernel(const float* __restrict__ x, long long n, float scale,
                                                                 float* __restrict__ out) {
  __shared__ float red[kTrThreads / 32];
  float acc = 0.f;
  for (long long i = (long long)blockIdx.x * kTrThreads + threadIdx.x; i < n; i += (long long)gridDim.x * kTrThreads) {
    const float v = __ldcs(x + i);
    acc = fmaf(v, v, acc);
  }
  for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int i = 0; i < kTrThreads / 32; ++i) s += red[i];
    atomicAdd(out, s * scale);
  }
}

// Minibatch standard deviation (stylegan2_arch.py:791-801; forward: b200ir_minibatch_stddev), first and second derivative
// along a tangent t.  Sample b = g * M + m belongs to statistic m; per (m, p, c): mu = mean_g x, sigma = sqrt(var_g x + 1e-8),
// tbar = mean_g t, dot = sum_g (x_g - mu) t_g.
//   JVP:  ts[m] = 1 / (C P) * sum_{p,c} dot / (G sigma)
//   HVP:  q[b][p][c] = a[m] / (C P) * ( (t_g - tbar) / (G sigma) - (x_g - mu) * dot / (G^2 sigma^3) )   (Hessian of a.s times t)
// One thread per (m, p, 8 channels), the group (<= 8 samples) in registers, as mbstd_bwd_kernel.
template <bool kHvp>
__global__ void __launch_bounds__(kTrThreads) mbstd_tangent_kernel(const uint4* __restrict__ x, const uint4* __restrict__ t,
                                                                   const float* __restrict__ a, float* __restrict__ ts,
                                                                   uint4* __restrict__ q, int M, int P, int groups, int group) {
  const long long n = (long long)M * P * groups;
  const float inv_cp = 1.f / ((float)groups * 8.f * (float)P);
  const float inv_g = 1.f / (float)group;
  // JVP partial sums are flushed per thread with one fp32 atomic (the layer runs at 4 x 4r pixels: a few thousand threads)
  for (long long idx = (long long)blockIdx.x * kTrThreads + threadIdx.x; idx < n; idx += (long long)gridDim.x * kTrThreads) {
    const unsigned r = (unsigned)idx / (unsigned)groups;
    const int g8 = (int)((unsigned)idx - r * (unsigned)groups);
    const int m = (int)(r / (unsigned)P), p = (int)(r - (unsigned)m * (unsigned)P);
    float xv[8][8], tv[8][8], mu[8], tb[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) mu[k] = tb[k] = 0.f;
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      if (g < group) {
        const long long row = ((long long)(g * M + m) * P + p) * groups + g8;
        tr_unpack(__ldg(x + row), xv[g]);
        tr_unpack(__ldg(t + row), tv[g]);
#pragma unroll
        for (int k = 0; k < 8; ++k) mu[k] += xv[g][k], tb[k] += tv[g][k];
      }
    }
    float jv = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      mu[k] *= inv_g;
      tb[k] *= inv_g;
      float var = 0.f, dot = 0.f;
#pragma unroll
      for (int g = 0; g < 8; ++g)
        if (g < group) {
          const float d = xv[g][k] - mu[k];
          var = fmaf(d, d, var);
          dot = fmaf(d, tv[g][k], dot);
        }
      const float rs = rsqrtf(var * inv_g + 1e-8f);  // 1 / sigma
      if (kHvp) {
        const float c1 = rs * inv_g, c2 = dot * inv_g * inv_g * rs * rs * rs;
#pragma unroll
        for (int g = 0; g < 8; ++g)
          if (g < group) tv[g][k] = (tv[g][k] - tb[k]) * c1 - (xv[g][k] - mu[k]) * c2;
      } else {
        jv = fmaf(dot * inv_g, rs, jv);
      }
    }
    if (kHvp) {
      const float am = __ldg(a + m) * inv_cp;
#pragma unroll
      for (int g = 0; g < 8; ++g)
        if (g < group) {
          float o[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) o[k] = tv[g][k] * am;
          q[((long long)(g * M + m) * P + p) * groups + g8] = tr_pack(o);
        }
    } else {
      atomicAdd(ts + m, jv * inv_cp);
    }
  }
}

// tcat[b][p][:] = [ t[b][p][0..C) | ts[m(b)] | zeros up to c_pad ]: the tangent of b200ir_minibatch_stddev's output
__global__ void __launch_bounds__(kTrThreads) mbstd_tangent_cat_kernel(const __half* __restrict__ t, const float* __restrict__ ts,
                                                                       __half* __restrict__ tcat, long long n_pix, int P, int C,
                                                                       int c_pad, int M) {
  const int pad_groups = c_pad / 8;
  const long long n = n_pix * pad_groups;
  for (long long idx = (long long)blockIdx.x * kTrThreads + threadIdx.x; idx < n; idx += (long long)gridDim.x * kTrThreads) {
    const long long pix = idx / pad_groups;
    const int g8 = (int)(idx - pix * pad_groups);
    const int b = (int)(pix / P);
    float o[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int c = g8 * 8 + k;
      o[k] = c < C ? __half2float(t[pix * C + c]) : (c == C ? ts[b % M] : 0.f);
    }
    *reinterpret_cast<uint4*>(tcat + pix * c_pad + g8 * 8) = tr_pack(o);
  }
}

// ------------------------------------------------------------------------------------------ weight packing
__global__ void __launch_bounds__(kTrThreads) pack_weights_kernel(const float* __restrict__ w, __half* __restrict__ out, int cout,
                                                                  int cin, int kh, int kw, float scale, int mode, int cin_pad) {
  const int taps = kh * kw;
  const long long n = (mode == 0) ? (long long)cout * taps * cin_pad : (long long)cin * taps * cout;
  for (long long idx = (long long)blockIdx.x * kTrThreads + threadIdx.x; idx < n; idx += (long long)gridDim.x * kTrThreads) {
    float v = 0.f;
    if (mode == 0) {
      const int ci = (int)(idx % cin_pad);
      const int t = (int)((idx / cin_pad) % taps);
      const int co = (int)(idx / ((long long)cin_pad * taps));
      if (ci < cin) v = w[((long long)co * cin + ci) * taps + t];
    } else {
      const int co = (int)(idx % cout);
      const int t = (int)((idx / cout) % taps);
      const int ci = (int)(idx / ((long long)cout * taps));
      v = w[((long long)co * cin + ci) * taps + (taps - 1 - t)];  // flipped tap: (kh-1-i)*kw + (kw-1-j) = taps-1 - (i*kw+j)
    }
    out[idx] = __float2half_rn(v * scale);
  }
}

}  // namespace b200ir

using namespace b200ir;
#define STREAM reinterpret_cast<cudaStream_t>(stream)

static inline unsigned tr_grid(long long n, int sms, int per_sm) {
  long long g = (n + kTrThreads - 1) / kTrThreads;
  const long long cap = (long long)sms * per_sm;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (unsigned)g;
}

extern "C" int b200ir_sft_mod(const void* a, const void* scale, const void* shift, int c_sft, const float* s_next, void* out,
                              int B, int64_t P, int C, void* stream) {
  B200IR_REQUIRE(a && out && B > 0 && P > 0 && C > 0 && C % 8 == 0, "sft_mod: bad arguments");
  B200IR_REQUIRE((scale == nullptr) == (shift == nullptr), "sft_mod: scale / shift must come together");
  B200IR_REQUIRE(scale == nullptr || (c_sft > 0 && c_sft % 8 == 0 && c_sft <= C), "sft_mod: c_sft=%d", c_sft);
  const int sms = num_sms();
  if (sms == 0) return 1;
  const long long n = (long long)B * P * (C / 8);
  B200IR_REQUIRE(n < (1LL << 32), "sft_mod: too many elements");
  sft_mod_kernel<<<tr_grid(n, sms, 16), kTrThreads, 0, STREAM>>>((const uint4*)a, (const uint4*)scale, (const uint4*)shift,
                                                                 scale ? c_sft / 8 : 0, s_next, (uint4*)out, n, P, C / 8);
  return check_launch("sft_mod");
}

extern "C" int b200ir_sft_mod_bwd(const void* g, const void* a, int64_t a_stride_b, const void* scale, const void* shift,
                                  int c_sft, const float* s_next, void* da, int accumulate, void* dscale, void* dshift,
                                  float* ds, int B, int64_t P, int C, void* stream) {
  B200IR_REQUIRE(g && a && B > 0 && P > 0 && C > 0 && C % 8 == 0 && a_stride_b % 8 == 0, "sft_mod_bwd: bad arguments");
  B200IR_REQUIRE((scale == nullptr) == (shift == nullptr), "sft_mod_bwd: scale / shift must come together");
  B200IR_REQUIRE(scale == nullptr || (c_sft > 0 && c_sft % 8 == 0 && c_sft <= C && dscale && dshift),
                 "sft_mod_bwd: c_sft=%d needs dscale / dshift", c_sft);
  B200IR_REQUIRE(da || ds || dscale, "sft_mod_bwd: nothing to compute");
  const int sms = num_sms();
  if (sms == 0) return 1;
  const TrLayout l = tr_layout(C, P, sms, B);
  sft_mod_bwd_kernel<<<dim3(l.grid_x, l.chunks, B), kTrThreads, 0, STREAM>>>(
      (const uint4*)g, (const uint4*)a, a_stride_b / 8, (const uint4*)scale, (const uint4*)shift, scale ? c_sft / 8 : 0, s_next,
      (uint4*)da, accumulate, (uint4*)dscale, (uint4*)dshift, ds, P, l.groups, C / 8);
  return check_launch("sft_mod_bwd");
}

extern "C" int b200ir_style_act_bwd(const void* da, const void* a, const float* noise, int64_t noise_stride_b,
                                    const float* noise_gain, const float* bias, const float* oscale, float mul, void* out,
                                    float* dd, int B, int64_t P, int C, void* stream) {
  B200IR_REQUIRE(da && a && out && B > 0 && P > 0 && C > 0 && C % 8 == 0, "style_act_bwd: bad arguments");
  B200IR_REQUIRE(noise == nullptr || noise_gain != nullptr, "style_act_bwd: noise without gain");
  const int sms = num_sms();
  if (sms == 0) return 1;
  const TrLayout l = tr_layout(C, P, sms, B);
  style_act_bwd_kernel<false><<<dim3(l.grid_x, l.chunks, B), kTrThreads, 0, STREAM>>>(
      (const uint4*)da, (const uint4*)a, noise, noise_stride_b, noise_gain, bias, oscale, mul, (uint4*)out, dd, nullptr, nullptr,
      P, l.groups, C / 8);
  return check_launch("style_act_bwd");
}

extern "C" int b200ir_style_act_bwd_params(const void* da, const void* a, const float* noise, int64_t noise_stride_b,
                                           const float* noise_gain, const float* bias, const float* oscale, float mul, void* out,
                                           float* dd, float* db, float* dn, int B, int64_t P, int C, void* stream) {
  B200IR_REQUIRE(da && a && out && db && dn && B > 0 && P > 0 && C > 0 && C % 8 == 0, "style_act_bwd_params: bad arguments");
  B200IR_REQUIRE(noise == nullptr || noise_gain != nullptr, "style_act_bwd_params: noise without gain");
  const int sms = num_sms();
  if (sms == 0) return 1;
  const TrLayout l = tr_layout(C, P, sms, B);
  style_act_bwd_kernel<true><<<dim3(l.grid_x, l.chunks, B), kTrThreads, 0, STREAM>>>(
      (const uint4*)da, (const uint4*)a, noise, noise_stride_b, noise_gain, bias, oscale, mul, (uint4*)out, dd, db, dn, P,
      l.groups, C / 8);
  return check_launch("style_act_bwd_params");
}

extern "C" int b200ir_to_rgb_bwd(const float* drgb, const void* a, const float* w, const float* s, void* da, int accumulate,
                                 float* ds, int B, int64_t P, int C, void* stream) {
  B200IR_REQUIRE(drgb && a && w && (da || ds) && B > 0 && P > 0 && C > 0 && C % 8 == 0, "to_rgb_bwd: bad arguments");
  const int sms = num_sms();
  if (sms == 0) return 1;
  const TrLayout l = tr_layout(C, P, sms, B);
  to_rgb_bwd_kernel<false><<<dim3(l.grid_x, l.chunks, B), kTrThreads, 0, STREAM>>>(drgb, (const uint4*)a, w, s, (uint4*)da,
                                                                                  accumulate, ds, nullptr, P, l.groups, C / 8);
  return check_launch("to_rgb_bwd");
}

extern "C" int b200ir_to_rgb_bwd_params(const float* drgb, const void* a, const float* w, const float* s, void* da, int accumulate,
                                        float* ds, float* R, int B, int64_t P, int C, void* stream) {
  B200IR_REQUIRE(drgb && a && w && R && B > 0 && P > 0 && C > 0 && C % 8 == 0, "to_rgb_bwd_params: bad arguments");
  const int sms = num_sms();
  if (sms == 0) return 1;
  const TrLayout l = tr_layout(C, P, sms, B);
  to_rgb_bwd_kernel<true><<<dim3(l.grid_x, l.chunks, B), kTrThreads, 0, STREAM>>>(drgb, (const uint4*)a, w, s, (uint4*)da,
                                                                                 accumulate, ds, R, P, l.groups, C / 8);
  return check_launch("to_rgb_bwd_params");
}

extern "C" int b200ir_plane_sums(const float* x, float* out, int B, int Cn, int64_t P, void* stream) {
  B200IR_REQUIRE(x && out && B > 0 && Cn > 0 && P > 0 && (long long)B * Cn <= 65535, "plane_sums: bad arguments");
  long long gx = (P + kTrThreads * 8 - 1) / (kTrThreads * 8);
  if (gx > 64) gx = 64;
  plane_sums_kernel<<<dim3((unsigned)gx, (unsigned)(B * Cn)), kTrThreads, 0, STREAM>>>(x, out, Cn, P);
  return check_launch("plane_sums");
}

extern "C" int b200ir_table_colsum(const void* in, int in_f16, const float* mul, int m, float scale, float* out, int B, int64_t n,
                                   void* stream) {
  B200IR_REQUIRE(in && out && B > 0 && n > 0 && (mul == nullptr || m > 0), "table_colsum: bad arguments");
  const unsigned grid = (unsigned)((n + kTrThreads - 1) / kTrThreads);
  if (in_f16)
    table_colsum_kernel<__half><<<grid, kTrThreads, 0, STREAM>>>((const __half*)in, mul, m, scale, out, B, n);
  else
    table_colsum_kernel<float><<<grid, kTrThreads, 0, STREAM>>>((const float*)in, mul, m, scale, out, B, n);
  return check_launch("table_colsum");
}

extern "C" int b200ir_mod_linear_wgrad(const float* ds, const float* latent, float wscale, float* dw, int L, int F, int lat_idx,
                                       int B, int cin, void* stream) {
  B200IR_REQUIRE(ds && latent && dw && B > 0 && cin > 0 && F > 0 && lat_idx >= 0 && lat_idx < L, "mod_linear_wgrad: bad arguments");
  mod_linear_wgrad_kernel<<<cin, kTrThreads, 0, STREAM>>>(ds, latent, wscale, dw, L, F, lat_idx, B, cin);
  return check_launch("mod_linear_wgrad");
}

extern "C" int b200ir_modconv_wgrad(const float* G, int transposed, const float* W, const float* s, const float* dd,
                                    const float* d, float scale, float* dw, int B, int cin, int cout, int taps, void* stream) {
  B200IR_REQUIRE(G && W && s && dd && d && dw && B > 0 && B <= 8192 && cin > 0 && cout > 0 && taps > 0,
                 "modconv_wgrad: bad arguments");
  modconv_wgrad_kernel<<<cout, kTrThreads, B * sizeof(float), STREAM>>>(G, transposed, W, s, dd, d, scale, dw, B, cin, cout, taps);
  return check_launch("modconv_wgrad");
}

extern "C" int b200ir_wgrad_unfold(const float* G, float* dw, int f, int cin, int cout, void* stream) {
  B200IR_REQUIRE(G && dw && f >= 1 && f <= 8 && cin > 0 && cout > 0, "wgrad_unfold: bad arguments");
  const int n = cout * 9 * cin;
  wgrad_unfold_kernel<<<(n + kTrThreads - 1) / kTrThreads, kTrThreads, 0, STREAM>>>(G, dw, f, cin, cout);
  return check_launch("wgrad_unfold");
}

extern "C" int b200ir_rgb_up_adjoint(const float* d, float* out, int planes, int h, int w, void* stream) {
  B200IR_REQUIRE(d && out && planes > 0 && h > 0 && w > 0, "rgb_up_adjoint: bad arguments");
  const long long n = (long long)planes * h * w;
  rgb_up_adjoint_kernel<<<(unsigned)((n + kTrThreads - 1) / kTrThreads), kTrThreads, 0, STREAM>>>(d, out, n, h, w);
  return check_launch("rgb_up_adjoint");
}

extern "C" int b200ir_demod_bwd(float* ds, const float* s, const float* dd, const float* d, const float* wsq, float scale2, int B,
                                int cin, int cout, void* stream) {
  B200IR_REQUIRE(ds && s && dd && d && wsq && B > 0 && cin > 0 && cout > 0 && cout <= 8192, "demod_bwd: bad arguments");
  demod_bwd_kernel<<<B, kTrThreads, cout * sizeof(float), STREAM>>>(ds, s, dd, d, wsq, scale2, cin, cout);
  return check_launch("demod_bwd");
}

extern "C" int b200ir_mod_linear_bwd(const float* ds, const float* w, float wscale, float* dlat, int L, int F, int lat_idx, int B,
                                     int cin, void* stream) {
  B200IR_REQUIRE(ds && w && dlat && B > 0 && cin > 0 && cin <= 8192 && F > 0 && lat_idx >= 0 && lat_idx < L,
                 "mod_linear_bwd: bad arguments");
  mod_linear_bwd_kernel<<<B, kTrThreads, cin * sizeof(float), STREAM>>>(ds, w, wscale, dlat, L, F, lat_idx, cin);
  return check_launch("mod_linear_bwd");
}

extern "C" int b200ir_first_conv_dgrad(const void* dz, const float* w, float* dx, int accumulate, int B, int H, int W, int cout,
                                       void* stream) {
  B200IR_REQUIRE(dz && w && dx && B > 0 && H > 0 && W > 0 && cout > 0 && cout % 8 == 0 && cout <= 1024,
                 "first_conv_dgrad: bad arguments");
  const long long n = (long long)B * H * W;
  first_conv_dgrad_kernel<<<(unsigned)((n + kTrThreads - 1) / kTrThreads), kTrThreads, cout * 3 * sizeof(float), STREAM>>>(
      (const uint4*)dz, w, dx, accumulate, n, H * W, cout);
  return check_launch("first_conv_dgrad");
}

extern "C" int b200ir_head_to_nchw(const void* head, float* rgb, int B, int64_t P, int cpad, void* stream) {
  B200IR_REQUIRE(head && rgb && B > 0 && P > 0 && cpad >= 8 && cpad % 8 == 0, "head_to_nchw: bad arguments");
  const long long n = (long long)B * P;
  head_to_nchw_kernel<<<(unsigned)((n + kTrThreads - 1) / kTrThreads), kTrThreads, 0, STREAM>>>((const __half*)head, rgb, n, P,
                                                                                               cpad);
  return check_launch("head_to_nchw");
}

extern "C" int b200ir_nchw_to_head(const float* drgb, void* dhead, int B, int64_t P, int cpad, void* stream) {
  B200IR_REQUIRE(drgb && dhead && B > 0 && P > 0 && cpad >= 8 && cpad % 8 == 0, "nchw_to_head: bad arguments");
  const long long n = (long long)B * P;
  nchw_to_head_kernel<<<(unsigned)((n + kTrThreads - 1) / kTrThreads), kTrThreads, 0, STREAM>>>(drgb, (__half*)dhead, n, P,
                                                                                               cpad);
  return check_launch("nchw_to_head");
}

extern "C" int b200ir_l1_loss(const float* x, const float* t, int64_t n, float weight, float grad_scale, float* loss, float* grad,
                              void* stream) {
  B200IR_REQUIRE(x && t && loss && n > 0, "l1_loss: bad arguments");
  const int sms = num_sms();
  if (sms == 0) return 1;
  l1_loss_kernel<<<tr_grid(n, sms, 8), kTrThreads, 0, STREAM>>>(x, t, n, weight / (float)n, grad_scale * weight / (float)n, loss,
                                                                grad);
  return check_launch("l1_loss");
}

extern "C" int b200ir_softplus_loss(const void* pred, int n, int stride, float sign, float weight, float grad_scale, float* loss,
                                    void* dpred, void* stream) {
  B200IR_REQUIRE(pred && loss && n > 0 && stride > 0 && (sign == 1.f || sign == -1.f), "softplus_loss: bad arguments");
  softplus_loss_kernel<<<1, kTrThreads, 0, STREAM>>>((const __half*)pred, n, stride, sign, weight / (float)n,
                                                     grad_scale * weight / (float)n, loss, (__half*)dpred);
  return check_launch("softplus_loss");
}

extern "C" int b200ir_maxpool2_relu(const void* z, void* out, int B, int H, int W, int C, void* stream) {
  B200IR_REQUIRE(z && out && B > 0 && H > 0 && W > 0 && H % 2 == 0 && W % 2 == 0 && C > 0 && C % 8 == 0,
                 "maxpool2_relu: bad arguments (H, W even; C %% 8 == 0)");
  const int sms = num_sms();
  if (sms == 0) return 1;
  const long long n = (long long)B * (H / 2) * (W / 2) * (C / 8);
  B200IR_REQUIRE(n < (1LL << 32), "maxpool2_relu: too many elements");
  maxpool2_relu_kernel<<<tr_grid(n, sms, 16), kTrThreads, 0, STREAM>>>((const uint4*)z, (uint4*)out, n, H / 2, W / 2, C / 8);
  return check_launch("maxpool2_relu");
}

extern "C" int b200ir_maxpool2_relu_bwd(const void* z, const void* dpool, const void* add, void* dz, int B, int H, int W, int C,
                                        void* stream) {
  B200IR_REQUIRE(z && dz && (dpool || add) && B > 0 && H > 0 && W > 0 && H % 2 == 0 && W % 2 == 0 && C > 0 && C % 8 == 0,
                 "maxpool2_relu_bwd: bad arguments");
  const int sms = num_sms();
  if (sms == 0) return 1;
  const long long n = (long long)B * (H / 2) * (W / 2) * (C / 8);
  B200IR_REQUIRE(n < (1LL << 32), "maxpool2_relu_bwd: too many elements");
  maxpool2_relu_bwd_kernel<<<tr_grid(n, sms, 16), kTrThreads, 0, STREAM>>>((const uint4*)z, (const uint4*)dpool, (const uint4*)add,
                                                                          (uint4*)dz, n, H / 2, W / 2, C / 8);
  return check_launch("maxpool2_relu_bwd");
}

extern "C" int b200ir_l1_loss_f16(const void* x, const void* t, int64_t n, float weight, float grad_scale, float* loss, void* grad,
                                  void* stream) {
  B200IR_REQUIRE(x && t && loss && n > 0 && n % 8 == 0, "l1_loss_f16: bad arguments (n %% 8 == 0)");
  const int sms = num_sms();
  if (sms == 0) return 1;
  l1_loss_f16_kernel<<<tr_grid(n / 8, sms, 8), kTrThreads, 0, STREAM>>>((const uint4*)x, (const uint4*)t, n / 8, weight / (float)n,
                                                                       grad_scale * weight / (float)n, loss, (uint4*)grad);
  return check_launch("l1_loss_f16");
}

extern "C" int b200ir_sum_squares(const float* x, int64_t n, float scale, float* out, void* stream) {
  B200IR_REQUIRE(x && out && n > 0, "sum_squares: bad arguments");
  const int sms = num_sms();
  if (sms == 0) return 1;
  sum_squares_kernel<<<tr_grid(n, sms, 8), kTrThreads, 0, STREAM>>>(x, n, scale, out);
  return check_launch("sum_squares");
}

extern "C" int b200ir_minibatch_stddev_jvp(const void* x, const void* t, float* ts, void* tcat, int B, int P, int C, int c_pad,
                                           int group, void* stream) {
  B200IR_REQUIRE(x && t && ts && tcat && B > 0 && P > 0 && C > 0 && C % 8 == 0 && c_pad > C && c_pad % 8 == 0,
                 "minibatch_stddev_jvp: bad arguments");
  B200IR_REQUIRE(group >= 1 && group <= 8 && B % group == 0, "minibatch_stddev_jvp: batch %d is not divisible by group %d", B,
                 group);
  const int sms = num_sms();
  if (sms == 0) return 1;
  const int M = B / group;
  const long long n = (long long)M * P * (C / 8);
  B200IR_REQUIRE(n < (1LL << 31), "minibatch_stddev_jvp: too many elements");
  if (cudaMemsetAsync(ts, 0, sizeof(float) * M, STREAM) != cudaSuccess) {
    set_error("minibatch_stddev_jvp: cudaMemsetAsync failed");
    return 1;
  }
  mbstd_tangent_kernel<false><<<tr_grid(n, sms, 8), kTrThreads, 0, STREAM>>>((const uint4*)x, (const uint4*)t, nullptr, ts,
                                                                            nullptr, M, P, C / 8, group);
  if (check_launch("minibatch_stddev_jvp")) return 1;
  const long long n2 = (long long)B * P * (c_pad / 8);
  mbstd_tangent_cat_kernel<<<tr_grid(n2, sms, 8), kTrThreads, 0, STREAM>>>((const __half*)t, ts, (__half*)tcat, (long long)B * P,
                                                                          P, C, c_pad, M);
  return check_launch("minibatch_stddev_jvp(cat)");
}

extern "C" int b200ir_minibatch_stddev_hvp(const void* x, const void* t, const float* a, void* q, int B, int P, int C, int group,
                                           void* stream) {
  B200IR_REQUIRE(x && t && a && q && B > 0 && P > 0 && C > 0 && C % 8 == 0, "minibatch_stddev_hvp: bad arguments");
  B200IR_REQUIRE(group >= 1 && group <= 8 && B % group == 0, "minibatch_stddev_hvp: batch %d is not divisible by group %d", B,
                 group);
  const int sms = num_sms();
  if (sms == 0) return 1;
  const int M = B / group;
  const long long n = (long long)M * P * (C / 8);
  B200IR_REQUIRE(n < (1LL << 31), "minibatch_stddev_hvp: too many elements");
  mbstd_tangent_kernel<true><<<tr_grid(n, sms, 8), kTrThreads, 0, STREAM>>>((const uint4*)x, (const uint4*)t, a, nullptr,
                                                                           (uint4*)q, M, P, C / 8, group);
  return check_launch("minibatch_stddev_hvp");
}

extern "C" int b200ir_pack_weights(const float* w, void* out, int cout, int cin, int kh, int kw, float scale, int mode,
                                   int cin_pad, void* stream) {
  B200IR_REQUIRE(w && out && cout > 0 && cin > 0 && kh > 0 && kw > 0 && (mode == 0 || mode == 1), "pack_weights: bad arguments");
  if (cin_pad == 0) cin_pad = cin;
  B200IR_REQUIRE(cin_pad >= cin && (mode == 0 || cin_pad == cin), "pack_weights: cin_pad=%d", cin_pad);
  const int sms = num_sms();
  if (sms == 0) return 1;
  const long long n = (mode == 0) ? (long long)cout * kh * kw * cin_pad : (long long)cin * kh * kw * cout;
  pack_weights_kernel<<<tr_grid(n, sms, 16), kTrThreads, 0, STREAM>>>(w, (__half*)out, cout, cin, kh, kw, scale, mode, cin_pad);
  return check_launch("pack_weights");
}
