// Backward of FusedLeakyReLU (basicsr/ops/fused_act/fused_act.py:30-63 -> fused_bias_act_kernel.cu:20-50 with act = 3,
// grad = 1) fused with the bias gradient (the sum over batch and pixels that FusedLeakyReLUFunctionBackward takes with
// grad_input.sum(dim)):  dz = dy * scale * (y > 0 ? 1 : slope),  dbias[c] = sum_p dz[p][c].
// The forward output y stands in for the reference's `out` (the sign of y is the sign of the pre-activation).
// HBM-bound: two 16-bit streams in, one out (6 bytes per element); every thread owns one group of 8 channels (16-byte
// accesses) and a strided set of pixels, so the bias partials stay in registers until one shared-memory reduction and
// one fp32 atomic per (CTA, channel).
#include "host_common.h"

namespace b200ir {

constexpr int kBwThreads = 256;

__global__ void __launch_bounds__(kBwThreads) lrelu_bias_bwd_kernel(const uint4* __restrict__ dy, const uint4* __restrict__ y,
                                                                    uint4* __restrict__ dz, float* __restrict__ dbias,
                                                                    long long n_pix, int groups, int row_groups, float slope,
                                                                    float scale) {
  __shared__ float part[kBwThreads][9];  // padded rows: the column sums below walk rows with stride `groups`
  const int g = threadIdx.x % groups;                   // channel group (8 channels) of this thread
  const int lane = threadIdx.x / groups;                // pixel lane inside the CTA
  const int lanes = kBwThreads / groups;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  const float pos = scale, neg = scale * slope;
  for (long long p = (long long)blockIdx.x * lanes + lane; p < n_pix; p += (long long)gridDim.x * lanes) {
    const long long i = p * row_groups + blockIdx.y * groups + g;   // blockIdx.y: chunk of `groups` channel groups
    const uint4 a = __ldcs(dy + i);
    const uint4 b = y ? __ldcs(y + i) : make_uint4(0x3C003C00u, 0x3C003C00u, 0x3C003C00u, 0x3C003C00u);  // no y: all 1.0
    uint4 o;
    const __half2* ah = reinterpret_cast<const __half2*>(&a);
    const __half2* bh = reinterpret_cast<const __half2*>(&b);
    __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 d = __half22float2(ah[j]), r = __half22float2(bh[j]);
      const float v0 = d.x * (r.x > 0.f ? pos : neg), v1 = d.y * (r.y > 0.f ? pos : neg);
      oh[j] = __floats2half2_rn(v0, v1);
      acc[2 * j] += v0;
      acc[2 * j + 1] += v1;
    }
    if (dz) __stcs(dz + i, o);
  }
  if (dbias == nullptr) return;
#pragma unroll
  for (int j = 0; j < 8; ++j) part[threadIdx.x][j] = acc[j];
  __syncthreads();
  for (int c = threadIdx.x; c < groups * 8; c += kBwThreads) {
    const int cg = c >> 3, cj = c & 7;
    float s = 0.f;
    for (int l = 0; l < lanes; ++l) s += part[l * groups + cg][cj];
    atomicAdd(dbias + blockIdx.y * groups * 8 + c, s);
  }
}

// Adjoint of fir_down2 (upfirdn2d(x, k, pad = (1, 1)) sampled at the even positions: the FIR + stride-2 front of the 1x1
// skip conv of ResBlock, stylegan2_ocr_arch.py:685-697 with kernel_size = 1): zero-stuffing x2 followed by the same 4-tap
// FIR, i.e. upfirdn2d(d, k, up = 2, pad = (2, 1)) without UpFirDnUpsample's gain of 4.  Per axis
//   out[2i] = (3 d[i] + d[i-1]) / 8,   out[2i+1] = (3 d[i] + d[i+1]) / 8,   d = 0 outside.
// One thread per (OUTPUT pixel, 8 channels): the 2 x 2 input neighbourhood it depends on (L1 / L2 hits: every input is
// shared by 9 outputs), the optional addend and one 16-byte store.  ~40 registers, so the SM runs at full occupancy; the
// first version computed a 2 x 2 output patch from a 3 x 3 input patch per thread (~100 registers, 2 CTAs per SM) and was
// latency-bound at ~40 % of the HBM rate on the training step's largest tensors (torch profiler, B = 256).
__global__ void __launch_bounds__(kBwThreads) fir_down2_adjoint_kernel(const uint4* __restrict__ d, const uint4* __restrict__ add,
                                                                       uint4* __restrict__ out, int B, int h, int w, int groups) {
  // grid: x = chunks of one output row's (X, channel group) pairs, y = output rows (b, Y): the index arithmetic per element is
  // one 32-bit division (a 64-bit division per element made the first per-pixel version ALU-bound at half the HBM rate)
  const int W2 = 2 * w, H2 = 2 * h;
  const unsigned row_elems = (unsigned)W2 * (unsigned)groups;
  for (unsigned row = blockIdx.y; row < (unsigned)B * (unsigned)H2; row += gridDim.y) {
    const int b = (int)(row / (unsigned)H2);
    const int Y = (int)(row - (unsigned)b * (unsigned)H2);
    const unsigned e = blockIdx.x * kBwThreads + threadIdx.x;
    if (e >= row_elems) continue;
    const int X = (int)(e / (unsigned)groups);
    const int g = (int)(e - (unsigned)X * (unsigned)groups);
    const long long idx = (long long)row * row_elems + e;
    const int i = Y >> 1, j = X >> 1;
    const int ni = (Y & 1) ? i + 1 : i - 1, nj = (X & 1) ? j + 1 : j - 1;   // the neighbour that gets weight 1/8 on this phase
    const bool vy = ni >= 0 && ni < h, vx = nj >= 0 && nj < w;
    const uint4* base = d + (long long)b * h * w * groups + g;
    const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
    const uint4 q11 = __ldg(base + ((long long)i * w + j) * groups);
    const uint4 qn1 = vy ? __ldg(base + ((long long)ni * w + j) * groups) : zero;
    const uint4 q1n = vx ? __ldg(base + ((long long)i * w + nj) * groups) : zero;
    const uint4 qnn = (vy && vx) ? __ldg(base + ((long long)ni * w + nj) * groups) : zero;
    float acc[8];
    if (add) {
      const uint4 q = __ldcs(add + idx);
      const __half2* hq = reinterpret_cast<const __half2*>(&q);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 f = __half22float2(hq[k]);
        acc[2 * k] = f.x;
        acc[2 * k + 1] = f.y;
      }
    } else {
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] = 0.f;
    }
    const __half2* h11 = reinterpret_cast<const __half2*>(&q11);
    const __half2* hn1 = reinterpret_cast<const __half2*>(&qn1);
    const __half2* h1n = reinterpret_cast<const __half2*>(&q1n);
    const __half2* hnn = reinterpret_cast<const __half2*>(&qnn);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 a = __half22float2(h11[k]), bq = __half22float2(hn1[k]), c = __half22float2(h1n[k]), e = __half22float2(hnn[k]);
      // same expression (and rounding) as the patch version: (9 v11 + 3 (vn1 + v1n) + vnn) / 64
      acc[2 * k] += (9.f * a.x + 3.f * (bq.x + c.x) + e.x) * (1.f / 64.f);
      acc[2 * k + 1] += (9.f * a.y + 3.f * (bq.y + c.y) + e.y) * (1.f / 64.f);
    }
    uint4 q;
    __half2* hq = reinterpret_cast<__half2*>(&q);
#pragma unroll
    for (int k = 0; k < 4; ++k) hq[k] = __floats2half2_rn(acc[2 * k], acc[2 * k + 1]);
    __stcs(out + idx, q);
  }
}

// Tiled variant for the large tensors of the training step: a CTA stages the (8 + 2) x (16 + 2) input pixels x 64 channels its
// 16 x 32 output tile depends on in shared memory once (the per-pixel kernel above reads every input 16 times through L1 / L2
// and ran at ~3.3 TB/s of algorithmic traffic on [256, 128, 384, 128]: L2-bandwidth-bound), then every thread produces 16
// outputs: four 16-byte shared-memory reads, the optional addend, one 16-byte store.  Same expression and rounding per output.
// GP = channel groups (of 8) per CTA: 8, or 4 for 32-channel tensors (then the tile is twice as wide: all 256 threads busy).
constexpr int kFdTi = 8;
template <int GP>
__global__ void __launch_bounds__(kBwThreads) fir_down2_adjoint_tiled_kernel(const uint4* __restrict__ d, const uint4* __restrict__ add,
                                                                             uint4* __restrict__ out, int h, int w, int groups,
                                                                             int chunks) {
  constexpr int kFdTj = 128 / GP;
  __shared__ uint4 tile[(kFdTi + 2) * (kFdTj + 2) * GP];
  const int chunk = blockIdx.x % chunks, tj = blockIdx.x / chunks, ti = blockIdx.y, b = blockIdx.z;
  const int g0 = chunk * GP, ng = min(GP, groups - g0);
  const int i0 = ti * kFdTi - 1, j0 = tj * kFdTj - 1;      // input coordinates of tile[0][0]
  const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
  const uint4* db = d + (long long)b * h * w * groups + g0;
  for (int e = threadIdx.x; e < (kFdTi + 2) * (kFdTj + 2) * GP; e += kBwThreads) {
    const int g = e % GP, pix = e / GP;
    const int r = pix / (kFdTj + 2), c = pix - r * (kFdTj + 2);
    const int i = i0 + r, j = j0 + c;
    tile[e] = (g < ng && i >= 0 && i < h && j >= 0 && j < w) ? __ldg(db + ((long long)i * w + j) * groups + g) : zero;
  }
  __syncthreads();
  const int g = threadIdx.x % GP, xl = threadIdx.x / GP;   // 2 * kFdTj output columns x GP channel groups per output row
  if (g >= ng) return;
  const int X = tj * 2 * kFdTj + xl;
  if (X >= 2 * w) return;
  const int jl = (xl >> 1) + 1, njl = (xl & 1) ? jl + 1 : jl - 1;
  const long long img = (long long)b * 4 * h * w * groups;
#pragma unroll 4
  for (int yl = 0; yl < 2 * kFdTi; ++yl) {
    const int Y = ti * 2 * kFdTi + yl;
    if (Y >= 2 * h) break;
    const int il = (yl >> 1) + 1, nil = (yl & 1) ? il + 1 : il - 1;
    const uint4 q11 = tile[(il * (kFdTj + 2) + jl) * GP + g], qn1 = tile[(nil * (kFdTj + 2) + jl) * GP + g];
    const uint4 q1n = tile[(il * (kFdTj + 2) + njl) * GP + g], qnn = tile[(nil * (kFdTj + 2) + njl) * GP + g];
    const long long idx = img + ((long long)Y * 2 * w + X) * groups + g0 + g;
    float acc[8];
    if (add) {
      const uint4 q = __ldcs(add + idx);
      const __half2* hq = reinterpret_cast<const __half2*>(&q);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 f = __half22float2(hq[k]);
        acc[2 * k] = f.x;
        acc[2 * k + 1] = f.y;
      }
    } else {
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] = 0.f;
    }
    const __half2* h11 = reinterpret_cast<const __half2*>(&q11);
    const __half2* hn1 = reinterpret_cast<const __half2*>(&qn1);
    const __half2* h1n = reinterpret_cast<const __half2*>(&q1n);
    const __half2* hnn = reinterpret_cast<const __half2*>(&qnn);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 a = __half22float2(h11[k]), bq = __half22float2(hn1[k]), c = __half22float2(h1n[k]), e = __half22float2(hnn[k]);
      acc[2 * k] += (9.f * a.x + 3.f * (bq.x + c.x) + e.x) * (1.f / 64.f);
      acc[2 * k + 1] += (9.f * a.y + 3.f * (bq.y + c.y) + e.y) * (1.f / 64.f);
    }
    uint4 q;
    __half2* hq = reinterpret_cast<__half2*>(&q);
#pragma unroll
    for (int k = 0; k < 4; ++k) hq[k] = __floats2half2_rn(acc[2 * k], acc[2 * k + 1]);
    __stcs(out + idx, q);
  }
}

// Adjoint of b200ir_bilinear_up2 (F.interpolate(scale_factor=2, mode='bilinear', align_corners=False) of ConvUpLayer,
// gfpganv1_ocr_arch.py:190): per axis hi[2i] = .75 lo[i] + .25 lo[max(i-1, 0)], hi[2i+1] = .75 lo[i] + .25 lo[min(i+1, n-1)],
// so  dlo[i] = sum_t (.25, .75, .75, .25)[t] * dhi[clamp(2i - 1 + t, 0, 2n - 1)]  (the clamp returns the border rows'
// replicated taps to the row that was replicated).  One thread per (low-resolution pixel, 8 channels): 4 x 4 gather.
__global__ void __launch_bounds__(kBwThreads) bilinear_up2_adjoint_kernel(const uint4* __restrict__ d, uint4* __restrict__ out,
                                                                          int B, int h, int w, int groups, float scale) {
  const long long n = (long long)B * h * w * groups;
  for (long long idx = (long long)blockIdx.x * kBwThreads + threadIdx.x; idx < n; idx += (long long)gridDim.x * kBwThreads) {
    unsigned r = (unsigned)idx / (unsigned)groups;   // 32-bit index arithmetic: the launcher keeps n below 2^31
    const int g = (int)((unsigned)idx - r * (unsigned)groups);
    const unsigned r2 = r / (unsigned)w;
    const int j = (int)(r - r2 * (unsigned)w);
    const int b = (int)(r2 / (unsigned)h);
    const int i = (int)(r2 - (unsigned)b * (unsigned)h);
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = 0.f;
#pragma unroll
    for (int ty = 0; ty < 4; ++ty) {
      const int y = min(max(2 * i - 1 + ty, 0), 2 * h - 1);
      const float wy = (ty == 0 || ty == 3) ? 0.25f : 0.75f;
#pragma unroll
      for (int tx = 0; tx < 4; ++tx) {
        const int x = min(max(2 * j - 1 + tx, 0), 2 * w - 1);
        const float wt = wy * ((tx == 0 || tx == 3) ? 0.25f : 0.75f);
        const uint4 q = __ldg(d + (((long long)b * 2 * h + y) * (2 * w) + x) * groups + g);
        const __half2* hq = reinterpret_cast<const __half2*>(&q);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 f = __half22float2(hq[k]);
          acc[2 * k] += wt * f.x;
          acc[2 * k + 1] += wt * f.y;
        }
      }
    }
    uint4 q;
    __half2* hq = reinterpret_cast<__half2*>(&q);
#pragma unroll
    for (int k = 0; k < 4; ++k) hq[k] = __floats2half2_rn(acc[2 * k] * scale, acc[2 * k + 1] * scale);
    out[idx] = q;
  }
}

// Weight gradient of conv_body_first (ConvLayer(3, C, 1): EqualConv2d 1x1 over the fp32 NCHW input image,
// stylegan2_ocr_arch.py:658-705): dw[c][k] = sum_p dz[p][c] * x[b(p)][k][hw(p)], k = 0..2.  Same thread layout as
// lrelu_bias_bwd_kernel (one group of 8 channels per thread, strided pixels), 24 partial sums per thread.
__global__ void __launch_bounds__(kBwThreads) first_conv_wgrad_kernel(const float* __restrict__ x, const uint4* __restrict__ dz,
                                                                      float* __restrict__ dw, long long n_pix, int HW,
                                                                      int groups) {
  __shared__ float part[kBwThreads][25];
  const int g = threadIdx.x % groups, lane = threadIdx.x / groups, lanes = kBwThreads / groups;
  float acc[8][3];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j][0] = acc[j][1] = acc[j][2] = 0.f;
  for (long long p = (long long)blockIdx.x * lanes + lane; p < n_pix; p += (long long)gridDim.x * lanes) {
    const uint4 a = __ldcs(dz + p * groups + g);
    const float* xp = x + (p / HW) * 3 * HW + (p % HW);
    const float x0 = __ldg(xp), x1 = __ldg(xp + HW), x2 = __ldg(xp + 2 * HW);
    const __half2* ah = reinterpret_cast<const __half2*>(&a);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 d = __half22float2(ah[j]);
      acc[2 * j][0] += d.x * x0, acc[2 * j][1] += d.x * x1, acc[2 * j][2] += d.x * x2;
      acc[2 * j + 1][0] += d.y * x0, acc[2 * j + 1][1] += d.y * x1, acc[2 * j + 1][2] += d.y * x2;
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j)
#pragma unroll
    for (int k = 0; k < 3; ++k) part[threadIdx.x][j * 3 + k] = acc[j][k];
  __syncthreads();
  for (int e = threadIdx.x; e < groups * 24; e += kBwThreads) {  // e = channel * 3 + k inside this CTA's channel range
    const int cg = e / 24, r = e % 24;
    float sum = 0.f;
    for (int l = 0; l < lanes; ++l) sum += part[l * groups + cg][r];
    atomicAdd(dw + e, sum);
  }
}

// Backward of the minibatch standard-deviation layer of the discriminator (stylegan2_arch.py:791-801; forward:
// b200ir_minibatch_stddev).  Sample b = g * M + m belongs to statistic m (M = B / group); s[m] = mean_{p,c} sigma[m][p][c],
// sigma = sqrt(var_g(x) + 1e-8), is broadcast into channel C of every member of the group.  With ds[m] the summed gradient
// of that channel:  dx[b][p][c] = dcat[b][p][c] + ds[m] * (x[b][p][c] - mu[m][p][c]) / (C * P * group * sigma[m][p][c]).
// One thread per (m, p, 8 channels), the group (<= 8 samples) in registers.
__global__ void __launch_bounds__(kBwThreads) mbstd_bwd_kernel(const uint4* __restrict__ x, const __half* __restrict__ dcat,
                                                               const float* __restrict__ ds, uint4* __restrict__ dx, int M, int P,
                                                               int groups, int c_pad, int group) {
  const long long n = (long long)M * P * groups;
  const float inv_n = 1.f / ((float)groups * 8.f * (float)P * (float)group);
  for (long long idx = (long long)blockIdx.x * kBwThreads + threadIdx.x; idx < n; idx += (long long)gridDim.x * kBwThreads) {
    const unsigned r = (unsigned)idx / (unsigned)groups;
    const int g8 = (int)((unsigned)idx - r * (unsigned)groups);
    const int m = (int)(r / (unsigned)P), p = (int)(r - (unsigned)m * (unsigned)P);
    float v[8][8], mu[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) mu[k] = 0.f;
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      if (g < group) {
        const uint4 q = __ldg(x + ((long long)(g * M + m) * P + p) * groups + g8);
        const __half2* hq = reinterpret_cast<const __half2*>(&q);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 f = __half22float2(hq[k]);
          v[g][2 * k] = f.x, v[g][2 * k + 1] = f.y;
          mu[2 * k] += f.x, mu[2 * k + 1] += f.y;
        }
      }
    }
    float coef[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      mu[k] /= (float)group;
      float var = 0.f;
#pragma unroll
      for (int g = 0; g < 8; ++g)
        if (g < group) var += (v[g][k] - mu[k]) * (v[g][k] - mu[k]);
      coef[k] = ds[m] * inv_n * rsqrtf(var / (float)group + 1e-8f);
    }
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      if (g < group) {
        const long long row = (long long)(g * M + m) * P + p;
        const uint4 dq = *reinterpret_cast<const uint4*>(dcat + row * c_pad + g8 * 8);
        const __half2* dh = reinterpret_cast<const __half2*>(&dq);
        uint4 o;
        __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 d = __half22float2(dh[k]);
          oh[k] = __floats2half2_rn(d.x + coef[2 * k] * (v[g][2 * k] - mu[2 * k]),
                                    d.y + coef[2 * k + 1] * (v[g][2 * k + 1] - mu[2 * k + 1]));
        }
        dx[row * groups + g8] = o;
      }
    }
  }
}

}  // namespace b200ir

using namespace b200ir;

extern "C" int b200ir_lrelu_bias_bwd(const void* dy, const void* y, void* dz, float* dbias, int64_t n_pix, int C, float slope,
                                     float scale, void* stream) {
  B200IR_REQUIRE(n_pix >= 0 && C > 0 && C % 8 == 0, "lrelu_bias_bwd: C=%d must be a multiple of 8", C);
  B200IR_REQUIRE(n_pix == 0 || (dy && (dz || dbias)), "lrelu_bias_bwd: null pointer");
  const int sms = num_sms();
  if (sms == 0) return 1;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (dbias && cudaMemsetAsync(dbias, 0, sizeof(float) * C, st) != cudaSuccess) {
    set_error("lrelu_bias_bwd: cudaMemsetAsync failed");
    return 1;
  }
  if (n_pix == 0) return 0;
  // channel groups (8 channels each) are cut into chunks whose size divides the CTA (every thread keeps one group)
  const int row_groups = C / 8;
  int groups = kBwThreads;
  while (row_groups % groups) groups >>= 1;
  const int chunks = row_groups / groups, lanes = kBwThreads / groups;
  long long grid = (n_pix + lanes - 1) / lanes;
  const long long cap = (8LL * sms + chunks - 1) / chunks;  // 8 resident CTAs of 256 threads per SM, one wave
  if (grid > cap) grid = cap;
  lrelu_bias_bwd_kernel<<<dim3((unsigned)grid, (unsigned)chunks), kBwThreads, 0, st>>>(
      (const uint4*)dy, (const uint4*)y, (uint4*)dz, dbias, n_pix, groups, row_groups, slope, scale);
  return check_launch("lrelu_bias_bwd");
}

extern "C" int b200ir_fir_pad11(const void* in, void* out, int B, int H, int W, int C, int in_h, int in_w, void* stream) {
  B200IR_REQUIRE(in && out && B > 0 && H > 0 && W > 0 && in_h >= H + 1 && in_w >= W + 1, "fir_pad11: bad arguments");
  FirLaunch a = {};
  a.in = (const __half*)in; a.B = B; a.Hv = H + 1; a.Wv = W + 1;
  a.in_sw = C; a.in_sh = (long long)in_w * C; a.in_sb = (long long)in_h * in_w * C;
  a.C = C; a.OH = H; a.OW = W; a.pad = 1; a.kscale = 0.125f;
  a.out = (__half*)out; a.out_sy = (long long)W * C; a.out_sb = (long long)H * W * C;
  a.post = false;
  const int r = fir_stream_launch(a, reinterpret_cast<cudaStream_t>(stream), "fir_pad11");
  if (r >= 0) return r;
  set_error("fir_pad11: C=%d must be a multiple of 32 and the input 16-byte aligned (TMA streaming kernel only)", C);
  return 1;
}

extern "C" int b200ir_fir_down2_adjoint(const void* d, const void* add, void* out, int B, int h, int w, int C, void* stream) {
  B200IR_REQUIRE(d && out && B > 0 && h > 0 && w > 0 && C > 0 && C % 8 == 0 && (long long)B * h * w * 4 < (1LL << 31),
                 "fir_down2_adjoint: bad arguments");
  const int sms = num_sms();
  if (sms == 0) return 1;
  const int groups = C / 8;
  if (h >= kFdTi && w >= 32 && B <= 65535 && (h + kFdTi - 1) / kFdTi <= 65535) {   // large levels: shared-memory tiles
    const int gp = (groups % 8 == 0) ? 8 : 4;
    const int tj = 128 / gp, chunks = (groups + gp - 1) / gp;
    const dim3 grid((unsigned)(((w + tj - 1) / tj) * chunks), (unsigned)((h + kFdTi - 1) / kFdTi), (unsigned)B);
    if (gp == 8)
      fir_down2_adjoint_tiled_kernel<8><<<grid, kBwThreads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
          (const uint4*)d, (const uint4*)add, (uint4*)out, h, w, groups, chunks);
    else
      fir_down2_adjoint_tiled_kernel<4><<<grid, kBwThreads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
          (const uint4*)d, (const uint4*)add, (uint4*)out, h, w, groups, chunks);
    return check_launch("fir_down2_adjoint");
  }
  const unsigned row_elems = (unsigned)(2 * w) * (unsigned)(C / 8);
  const long long rows = (long long)B * 2 * h;
  const dim3 grid((row_elems + kBwThreads - 1) / kBwThreads, (unsigned)(rows < 65535 ? rows : 65535));
  fir_down2_adjoint_kernel<<<grid, kBwThreads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      (const uint4*)d, (const uint4*)add, (uint4*)out, B, h, w, C / 8);
  return check_launch("fir_down2_adjoint");
}

extern "C" int b200ir_bilinear_up2_adjoint(const void* d, void* out, int B, int h, int w, int C, float scale, void* stream) {
  B200IR_REQUIRE(d && out && B > 0 && h > 0 && w > 0 && C > 0 && C % 8 == 0 && (long long)B * h * w * (C / 8) < (1LL << 31),
                 "bilinear_up2_adjoint: bad arguments");
  const int sms = num_sms();
  if (sms == 0) return 1;
  const long long n = (long long)B * h * w * (C / 8);
  long long grid = (n + kBwThreads - 1) / kBwThreads;
  if (grid > 16LL * sms) grid = 16LL * sms;
  bilinear_up2_adjoint_kernel<<<(int)grid, kBwThreads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      (const uint4*)d, (uint4*)out, B, h, w, C / 8, scale);
  return check_launch("bilinear_up2_adjoint");
}

extern "C" int b200ir_first_conv_wgrad(const float* x, const void* dz, float* dw, int B, int H, int W, int cout, void* stream) {
  B200IR_REQUIRE(x && dz && dw && B > 0 && H > 0 && W > 0 && cout > 0 && cout % 8 == 0 && kBwThreads % (cout / 8) == 0,
                 "first_conv_wgrad: cout=%d must be 8 * a divisor of %d", cout, kBwThreads);
  const int sms = num_sms();
  if (sms == 0) return 1;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (cudaMemsetAsync(dw, 0, sizeof(float) * cout * 3, st) != cudaSuccess) {
    set_error("first_conv_wgrad: cudaMemsetAsync failed");
    return 1;
  }
  const long long n_pix = (long long)B * H * W;
  const int groups = cout / 8, lanes = kBwThreads / groups;
  long long grid = (n_pix + lanes - 1) / lanes;
  if (grid > 4LL * sms) grid = 4LL * sms;
  first_conv_wgrad_kernel<<<(int)grid, kBwThreads, 0, st>>>(x, (const uint4*)dz, dw, n_pix, H * W, groups);
  return check_launch("first_conv_wgrad");
}

extern "C" int b200ir_minibatch_stddev_bwd(const void* x, const void* dcat, const float* ds, void* dx, int B, int P, int C,
                                           int c_pad, int group, void* stream) {
  B200IR_REQUIRE(x && dcat && ds && dx && B > 0 && P > 0 && C > 0 && C % 8 == 0 && c_pad > C && c_pad % 8 == 0,
                 "minibatch_stddev_bwd: bad arguments");
  B200IR_REQUIRE(group >= 1 && group <= 8 && B % group == 0, "minibatch_stddev_bwd: batch %d is not divisible by group %d", B,
                 group);
  const int sms = num_sms();
  if (sms == 0) return 1;
  const int M = B / group;
  const long long n = (long long)M * P * (C / 8);
  B200IR_REQUIRE(n < (1LL << 31), "minibatch_stddev_bwd: too many elements");
  long long grid = (n + kBwThreads - 1) / kBwThreads;
  if (grid > 8LL * sms) grid = 8LL * sms;
  mbstd_bwd_kernel<<<(int)grid, kBwThreads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      (const uint4*)x, (const __half*)dcat, ds, (uint4*)dx, M, P, C / 8, c_pad, group);
  return check_launch("minibatch_stddev_bwd");
}
