// TMA-fed streaming resamplers for NHWC fp16 tensors (same producer-warp / smem-ring / register-window structure as
// fir_tma.cu):
//   DOWN2: UpFirDnSmooth (pad (1,1), taps outer([1,3,3,1])/64) followed by the stride-2 sampling of the 1x1 skip conv of
//          ResBlock (stylegan2_ocr_arch.py:116-121,726-727; upfirdn2d.py:162-192): out(y,x) = sum k[a]k[b] in(2y+a-1, 2x+b-1)
//   UP2:   F.interpolate(scale_factor=2, mode='bilinear', align_corners=False) of ConvUpLayer (gfpganv1_ocr_arch.py:190):
//          out(2k) = .25 in(k-1) + .75 in(k), out(2k+1) = .75 in(k) + .25 in(k+1), indices clamped to the tensor.
// HBM-bound.  A work item is (image, column strip, chunk of 64 / 32 channels, chunk of rows); a producer warp streams the
// item's input rows through shared-memory stages with cp.async.bulk.tensor (out-of-bounds rows / columns arrive as
// zeros: exactly the zero padding DOWN2 needs; UP2 replaces them by the clamped neighbour), 8 consumer warps filter
// horizontally out of shared memory (128-bit, conflict-free) and vertically over a register window, and write every
// output once with 128-bit stores.  The direct kernels in pointwise.cu read every input 4x (DOWN2) / 9x (UP2) through L1.
#include <string.h>

#include "host_common.h"
#include "ptx.cuh"

namespace b200ir {

static constexpr int kRsConsumers = 256;
static constexpr int kRsThreads = kRsConsumers + 32;
static constexpr int kRsMaxStages = 8;
static constexpr int kRsRows = 4;  // input rows per stage

struct alignas(64) RsParams {
  CUtensorMap tmap_in;
  int B, H, W, C;  // input extents
  int OH, OW;      // output extents
  int strips, chunks, rchunks, R, KS, num_items;
  int NS, stage_bytes;
  __half* out;
};

struct RsItem {
  int b, strip, chunk, y0;
};
__device__ __forceinline__ RsItem rs_decode(const RsParams& p, int item) {
  RsItem t;
  t.chunk = item % p.chunks;
  item /= p.chunks;
  t.strip = item % p.strips;
  item /= p.strips;
  t.y0 = (item % p.rchunks) * p.R;
  t.b = item / p.rchunks;
  return t;
}

__device__ __forceinline__ void rs_unpack8(const uint4& q, float* f) {
  const __half2* h = reinterpret_cast<const __half2*>(&q);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 x = __half22float2(h[i]);
    f[2 * i] = x.x;
    f[2 * i + 1] = x.y;
  }
}
__device__ __forceinline__ uint4 rs_lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void rs_store8(__half* p, const float* v) {
  uint4 o;
  __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
  for (int e = 0; e < 4; ++e) oh[e] = f2h2_sat(v[2 * e], v[2 * e + 1]);
  *reinterpret_cast<uint4*>(p) = o;
}

// DOWN = true : strips / row chunks are counted in OUTPUT pixels; stage k holds input rows 2*y0 - 1 + 4k .. +3 and input
//               columns 2*x0 - 1 .. 2*x0 + 2*TW (IW = 2*TW + 2)
// DOWN = false: (UP2) strips / row chunks are counted in INPUT pixels; stage k holds input rows y0 - 1 + 4k .. +3 and input
//               columns x0 - 1 .. x0 + TW (IW = TW + 2)
template <int CC, bool DOWN>
__global__ void __launch_bounds__(kRsThreads, 2) resample_stream_kernel(const __grid_constant__ RsParams p) {
  constexpr int CG = CC / 8;
  constexpr int TW = kRsConsumers / CG;
  constexpr int IW = DOWN ? 2 * TW + 2 : TW + 2;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* base = smem_raw + (((raw_addr + 127u) & ~127u) - raw_addr);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(base + p.NS * p.stage_bytes);
  uint64_t* empty_bar = full_bar + kRsMaxStages;
  const int tid = threadIdx.x;
  if (tid == 0) {
    for (int i = 0; i < p.NS; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], kRsConsumers);
    }
    fence_barrier_init();
  }
  __syncthreads();
  const int my_items = (p.num_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  if (tid >= kRsConsumers) {
    if (tid == kRsConsumers) {  // ---------------- producer
      tma_prefetch_desc(&p.tmap_in);
      int slot = 0;
      uint32_t phase = 0;
      for (int it = 0; it < my_items; ++it) {
        const RsItem t = rs_decode(p, blockIdx.x + it * gridDim.x);
        const int cx = (DOWN ? 2 * t.strip * TW : t.strip * TW) - 1;
        const int cy = (DOWN ? 2 * t.y0 : t.y0) - 1;
        for (int k = 0; k < p.KS; ++k) {
          mbar_wait(&empty_bar[slot], phase ^ 1u);
          mbar_arrive_expect_tx(&full_bar[slot], p.stage_bytes);
          tma_load_4d(base + slot * p.stage_bytes, &p.tmap_in, &full_bar[slot], t.chunk * CC, cx, cy + k * kRsRows, t.b);
          if (++slot == p.NS) {
            slot = 0;
            phase ^= 1u;
          }
        }
      }
    }
    return;
  }

  // ---------------- consumers: thread = (8-channel group cg, column slot xq)
  const int cg = tid % CG;
  const int xq = tid / CG;
  const uint32_t thr_in = smem_u32(base) + ((DOWN ? 2 * xq : xq) * CC + cg * 8) * 2;  // first input column this thread reads
  float pa[8], pb[8];  // DOWN: horizontally filtered rows 2, 3 of the previous stage; UP2: h_lo / h_hi of the previous row
#pragma unroll
  for (int e = 0; e < 8; ++e) pa[e] = pb[e] = 0.f;
  int slot = 0;
  uint32_t phase = 0;
  for (int it = 0; it < my_items; ++it) {
    const RsItem t = rs_decode(p, blockIdx.x + it * gridDim.x);
    const int c = t.chunk * CC + cg * 8;
    const int x = t.strip * TW + xq;  // DOWN: output column; UP2: input column
    const bool xv = x < (DOWN ? p.OW : p.W);
    for (int k = 0; k < p.KS; ++k) {
      mbar_wait(&full_bar[slot], phase);
      const uint32_t st = thr_in + slot * p.stage_bytes;
      if (DOWN) {
        // horizontal pass of the four input rows: taps (1,3,3,1) over input columns 2x-1 .. 2x+2
        float hf[kRsRows][8];
#pragma unroll
        for (int r = 0; r < kRsRows; ++r) {
          float f[4][8];
#pragma unroll
          for (int col = 0; col < 4; ++col) rs_unpack8(rs_lds128(st + (r * IW + col) * CC * 2), f[col]);
#pragma unroll
          for (int e = 0; e < 8; ++e) hf[r][e] = fmaf(3.f, f[1][e] + f[2][e], f[0][e] + f[3][e]);
        }
        const int oa = t.y0 + 2 * k;  // aligned output row: input rows 0..3 of this stage
        const int ob = oa - 1;        // straddling output row: rows 2, 3 of the previous stage + rows 0, 1 of this one
        const int y_end = min(t.y0 + p.R, p.OH);
        if (xv) {
          float v[8];
          if (k > 0 && ob < y_end) {
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = fmaf(3.f, pb[e] + hf[0][e], pa[e] + hf[1][e]) * (1.f / 64.f);
            rs_store8(p.out + (((long long)t.b * p.OH + ob) * p.OW + x) * p.C + c, v);
          }
          if (oa < y_end) {
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = fmaf(3.f, hf[1][e] + hf[2][e], hf[0][e] + hf[3][e]) * (1.f / 64.f);
            rs_store8(p.out + (((long long)t.b * p.OH + oa) * p.OW + x) * p.C + c, v);
          }
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          pa[e] = hf[2][e];
          pb[e] = hf[3][e];
        }
      } else {
        // UP2: input row n = y0 - 1 + 4k + r; with the previous row it yields output rows 2n-1 and 2n
        const bool x_lo = (x == 0), x_hi = (x == p.W - 1);
#pragma unroll
        for (int r = 0; r < kRsRows; ++r) {
          const int n = t.y0 - 1 + k * kRsRows + r;
          float f[3][8];
#pragma unroll
          for (int col = 0; col < 3; ++col) rs_unpack8(rs_lds128(st + (r * IW + col) * CC * 2), f[col]);
          float hl[8], hh[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float left = x_lo ? f[1][e] : f[0][e];   // clamp: in(-1) := in(0)
            const float right = x_hi ? f[1][e] : f[2][e];  //        in(W) := in(W-1)
            hl[e] = 0.25f * left + 0.75f * f[1][e];
            hh[e] = 0.75f * f[1][e] + 0.25f * right;
          }
          if (xv && n >= t.y0 && n <= min(t.y0 + p.R, p.H)) {
            // rows clamp the same way: above the first row the previous row is the row itself, below the last row the
            // current row is the previous one
            const bool top = (n == 0), bot = (n >= p.H);
            __half* o = p.out + (((long long)t.b * p.OH + (2 * n - 1)) * p.OW + 2 * x) * p.C + c;
            float v[8];
            if (n > t.y0) {  // output row 2n-1 = .75 in(n-1) + .25 in(n)
#pragma unroll
              for (int e = 0; e < 8; ++e) v[e] = 0.75f * pa[e] + 0.25f * (bot ? pa[e] : hl[e]);
              rs_store8(o, v);
#pragma unroll
              for (int e = 0; e < 8; ++e) v[e] = 0.75f * pb[e] + 0.25f * (bot ? pb[e] : hh[e]);
              rs_store8(o + p.C, v);
            }
            if (n < min(t.y0 + p.R, p.H)) {  // output row 2n = .25 in(n-1) + .75 in(n)
              o += (long long)p.OW * p.C;
#pragma unroll
              for (int e = 0; e < 8; ++e) v[e] = 0.25f * (top ? hl[e] : pa[e]) + 0.75f * hl[e];
              rs_store8(o, v);
#pragma unroll
              for (int e = 0; e < 8; ++e) v[e] = 0.25f * (top ? hh[e] : pb[e]) + 0.75f * hh[e];
              rs_store8(o + p.C, v);
            }
          }
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            pa[e] = hl[e];
            pb[e] = hh[e];
          }
        }
      }
      mbar_arrive(&empty_bar[slot]);
      if (++slot == p.NS) {
        slot = 0;
        phase ^= 1u;
      }
    }
  }
}

template <int CC, bool DOWN>
static int rs_launch_variant(const __half* in, __half* out, int B, int H, int W, int C, cudaStream_t st, const char* what) {
  constexpr int CG = CC / 8, TW = kRsConsumers / CG, IW = DOWN ? 2 * TW + 2 : TW + 2;
  RsParams p;
  memset(&p, 0, sizeof(p));
  p.B = B; p.H = H; p.W = W; p.C = C;
  p.OH = DOWN ? H / 2 : 2 * H;
  p.OW = DOWN ? W / 2 : 2 * W;
  const int span_w = DOWN ? p.OW : W, span_h = DOWN ? p.OH : H;  // extents the strips / row chunks are counted in
  p.strips = (span_w + TW - 1) / TW;
  p.chunks = C / CC;
  const int sms = num_sms();
  if (sms == 0) return 1;
  int R = 32;
  while (R > 8 && (long long)B * p.strips * p.chunks * ((span_h + R - 1) / R) < 4LL * sms) R /= 2;
  if (R > span_h) R = span_h;
  R = (span_h + (span_h + R - 1) / R - 1) / ((span_h + R - 1) / R);  // equal row chunks
  if (DOWN) R = (R + 1) & ~1;                                          // the stride-2 window advances two rows per stage
  p.R = R;
  p.rchunks = (span_h + R - 1) / R;
  // DOWN: R/2 aligned outputs per ... every stage yields two output rows, plus one stage for the last straddling row;
  // UP2: rows y0-1 .. y0+R
  p.KS = DOWN ? (R + 1) / 2 + 1 : (R + 2 + kRsRows - 1) / kRsRows;
  p.num_items = B * p.strips * p.chunks * p.rchunks;
  p.stage_bytes = kRsRows * IW * CC * 2;
  static_assert((kRsRows * IW * CC * 2) % 128 == 0, "stages must stay 128-byte aligned");
  const int smem_max = smem_optin();
  int ns = (smem_max / 2 - 512) / p.stage_bytes;  // two CTAs per SM
  if (ns > 4) ns = 4;
  B200IR_REQUIRE(ns >= 2, "%s: stage of %d bytes does not fit", what, p.stage_bytes);
  p.NS = ns;
  p.out = out;
  {
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
    cuuint32_t box[4] = {(cuuint32_t)CC, (cuuint32_t)IW, (cuuint32_t)kRsRows, 1u};
    if (encode_map(&p.tmap_in, in, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE, what)) return 1;
  }
  const int smem_bytes = p.NS * p.stage_bytes + 256 + 128;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(resample_stream_kernel<CC, DOWN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         smem_max);
    if (e != cudaSuccess) {
      set_error("%s: cudaFuncSetAttribute: %s", what, cudaGetErrorString(e));
      return 1;
    }
    configured = true;
  }
  int grid = 2 * sms;
  if (grid > p.num_items) grid = p.num_items;
  resample_stream_kernel<CC, DOWN><<<grid, kRsThreads, smem_bytes, st>>>(p);
  return check_launch(what);
}

// returns -1 when the shape is not eligible (caller uses the direct kernels of pointwise.cu)
int resample_stream_launch(bool down, const __half* in, __half* out, int B, int H, int W, int C, cudaStream_t st) {
  if (C % 32 != 0 || (reinterpret_cast<uintptr_t>(in) & 15) != 0 || H < 2 || W < 2) return -1;
  if (down && (H % 2 || W % 2)) return -1;
  if (C % 64 == 0)
    return down ? rs_launch_variant<64, true>(in, out, B, H, W, C, st, "fir_down2")
                : rs_launch_variant<64, false>(in, out, B, H, W, C, st, "bilinear_up2");
  return down ? rs_launch_variant<32, true>(in, out, B, H, W, C, st, "fir_down2")
              : rs_launch_variant<32, false>(in, out, B, H, W, C, st, "bilinear_up2");
}

}  // namespace b200ir
