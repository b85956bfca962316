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
                                                                    long long n_pix, int groups, float slope, float scale) {
  __shared__ float part[kBwThreads][9];  // padded rows: the column sums below walk rows with stride `groups`
  const int g = threadIdx.x % groups;                   // channel group (8 channels) of this thread
  const int lane = threadIdx.x / groups;                // pixel lane inside the CTA
  const int lanes = kBwThreads / groups;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  const float pos = scale, neg = scale * slope;
  for (long long p = (long long)blockIdx.x * lanes + lane; p < n_pix; p += (long long)gridDim.x * lanes) {
    const long long i = p * groups + g;
    const uint4 a = __ldcs(dy + i), b = __ldcs(y + i);
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
    __stcs(dz + i, o);
  }
  if (dbias == nullptr) return;
#pragma unroll
  for (int j = 0; j < 8; ++j) part[threadIdx.x][j] = acc[j];
  __syncthreads();
  for (int c = threadIdx.x; c < groups * 8; c += kBwThreads) {
    const int cg = c >> 3, cj = c & 7;
    float s = 0.f;
    for (int l = 0; l < lanes; ++l) s += part[l * groups + cg][cj];
    atomicAdd(dbias + c, s);
  }
}

}  // namespace b200ir

using namespace b200ir;

extern "C" int b200ir_lrelu_bias_bwd(const void* dy, const void* y, void* dz, float* dbias, int64_t n_pix, int C, float slope,
                                     float scale, void* stream) {
  B200IR_REQUIRE(n_pix >= 0 && C > 0 && C % 8 == 0 && kBwThreads % (C / 8) == 0,
                 "lrelu_bias_bwd: C=%d must be 8 * a divisor of %d", C, kBwThreads);
  B200IR_REQUIRE(n_pix == 0 || (dy && y && dz), "lrelu_bias_bwd: null pointer");
  const int sms = num_sms();
  if (sms == 0) return 1;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (dbias && cudaMemsetAsync(dbias, 0, sizeof(float) * C, st) != cudaSuccess) {
    set_error("lrelu_bias_bwd: cudaMemsetAsync failed");
    return 1;
  }
  if (n_pix == 0) return 0;
  const int groups = C / 8, lanes = kBwThreads / groups;
  long long grid = (n_pix + lanes - 1) / lanes;
  if (grid > 8LL * sms) grid = 8LL * sms;  // 8 resident CTAs of 256 threads per SM, one wave
  lrelu_bias_bwd_kernel<<<(int)grid, kBwThreads, 0, st>>>((const uint4*)dy, (const uint4*)y, (uint4*)dz, dbias, n_pix, groups,
                                                          slope, scale);
  return check_launch("lrelu_bias_bwd");
}
