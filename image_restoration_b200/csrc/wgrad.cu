// Weight gradient of a 3x3 stride-1 'same' convolution (the EqualConv2d / ConvUpLayer convs of the U-Net and of the SFT
// heads; first backward kernel of the training-step row) as a tcgen05 GEMM whose contraction runs over
// PIXELS:
//     dW[co][kh][kw][ci] = sum_{b,y,x} dy[b][y][x][co] * x[b][y + kh - 1][x + kw - 1][ci]
// Both operands arrive exactly as the forward kernel's activation tiles do -- TMA boxes of (64 channels, 32 x 4 pixels)
// with the 128-byte swizzle, i.e. one 128-byte row per pixel -- but are handed to the tensor core as MN-MAJOR matrices
// (rows = K = pixels, the 64 contiguous channels = M resp. N): no transposed copy of the activations exists anywhere.
// Zero padding of the conv = TMA out-of-bounds fill of the shifted x box; pixels outside the image contribute zeros on
// the dy side as well, so ragged tiles need no masks.
//   CTA = (128 output channels) x (128 or 64 input channels) x (one kernel row kh: 3 accumulators of 128 x N fp32 in
//         TMEM) x (one contiguous range of pixel tiles: split-K); warp 0 = TMA producer, warp 1 = MMA issuer, warps 2-5
//         = epilogue (TMEM -> fp32 atomic adds into dW, which the entry point zeroes first).
// The x tile carries a one-pixel halo left and right (box 34 x 4 pixels) and is loaded ONCE per stage: the three kw taps
// read it through descriptors whose start address is shifted by kw pixel rows of 128 bytes (the K dimension of an
// MN-major operand is linear in shared memory: 128 bytes per pixel, 8-pixel atoms 1024 bytes apart; the swizzle XOR uses
// absolute address bits, so unaligned starts work with the base-offset field at 0, as in the forward row kernel).
// MMAs are issued per image row of the tile (32 pixels = two K = 16 steps) because dy (pitch 32) and x (pitch 34) only
// share the pixel order inside a row.
#include <stdlib.h>

#include "host_common.h"
#include "ptx.cuh"

namespace b200ir {

constexpr int kWgThreads = 192;
constexpr int kWgTileW = 32, kWgTileH = 4;                 // 128 pixels = K of one stage
constexpr uint32_t kWgBoxDy = 128 * 128;                   // bytes of one dy box (64 ch x 32 x 4 px)
constexpr uint32_t kWgBoxX = (kWgTileW + 2) * kWgTileH * 128;  // x box with halo (64 ch x 34 x 4 px) = 17 KB
constexpr uint32_t kWgTmemCols = 512;                      // 3 accumulators of up to 128 columns, power of two

template <int kNC>  // 64-channel chunks of the input-channel block: N = 64 * kNC
struct WgCfg {
  static constexpr uint32_t kStage = 2 * kWgBoxDy + kNC * kWgBoxX;  // dy: 2 boxes (128 co) | x: kNC boxes
  static constexpr int kStages = kNC == 2 ? 3 : 4;
};

struct WgradParams {
  CUtensorMap tmap_x, tmap_dy;
  float* dw;
  int cin, cout;
  int tiles_x, tiles_y, num_tiles, splits, ci_blocks;
  uint32_t tap_mask;   // bit kh * 3 + kw: taps to compute (the others stay zero)
  int num_kh;          // kernel rows with at least one tap in the mask
  int kh_list[3];
  // output addressing: dw[img * img_stride + co * co_stride + (kh * 3 + kw) * tap_stride + ci].  The weight gradient sums over
  // the batch (per_image = 0, one image slot); the per-image Gram matrices of the style loss do not (per_image = 1, grid.y = B,
  // compact [B][cout][cin] output of the single centre tap).
  int per_image;
  long long img_stride, co_stride, tap_stride;
};

// Shared-memory matrix descriptor of an MN-major operand with the 128-byte swizzle (canonical layout, in 16-byte units:
// ((8, n), (8, k)) : ((1, LBO), (8, SBO)) -- 64 MN-elements contiguous per K row, 8 K rows per 1024-byte atom, SBO
// between K atoms, LBO between 64-element MN chunks).
__device__ __forceinline__ uint64_t make_mnmajor_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= 1ull << 46;
  d |= 2ull << 61;  // SWIZZLE_128B
  return d;
}

template <int kNC>
__global__ void __launch_bounds__(kWgThreads, 1) conv_wgrad_kernel(const __grid_constant__ WgradParams p) {
  constexpr uint32_t kWgStage = WgCfg<kNC>::kStage;
  constexpr int kWgStages = WgCfg<kNC>::kStages;
  extern __shared__ __align__(1024) uint8_t wg_smem[];
  __shared__ __align__(8) uint64_t full_bar[4], empty_bar[4], acc_bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  int u = blockIdx.x;
  const int split = u % p.splits;
  u /= p.splits;
  const int kh = p.kh_list[u % p.num_kh];
  u /= p.num_kh;
  const uint32_t kw_mask = (p.tap_mask >> (kh * 3)) & 7u;
  const int ci_blk = u % p.ci_blocks;
  const int co_blk = u / p.ci_blocks;
  const int img = blockIdx.y;  // 0 unless per_image
  const int nt = p.per_image ? p.tiles_x * p.tiles_y : p.num_tiles;
  const int t_base = p.per_image ? img * nt : 0;
  const int t0 = t_base + (int)((long long)split * nt / p.splits);
  const int t1 = t_base + (int)((long long)(split + 1) * nt / p.splits);
  if (t0 >= t1) return;  // CTA-uniform

  uint8_t* ring = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(wg_smem) + 1023) & ~(uintptr_t)1023);
  if (threadIdx.x == 0) {
    for (int i = 0; i < kWgStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    mbar_init(&acc_bar, 1);
    fence_barrier_init();
    tma_prefetch_desc(&p.tmap_x);
    tma_prefetch_desc(&p.tmap_dy);
  }
  if (warp == 1) {
    tmem_alloc(&tmem_slot, kWgTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;

  if (warp == 0) {
    if (lane == 0) {  // ---------------- TMA producer
      int stage = 0;
      uint32_t phase = 0;
      const int per_img = p.tiles_x * p.tiles_y;
      for (int t = t0; t < t1; ++t) {
        const int b = t / per_img, r = t - b * per_img;
        const int y0 = (r / p.tiles_x) * kWgTileH, x0 = (r % p.tiles_x) * kWgTileW;
        mbar_wait_parked(&empty_bar[stage], phase ^ 1u);
        uint8_t* sa = ring + stage * kWgStage;
        mbar_arrive_expect_tx(&full_bar[stage], kWgStage);
        tma_load_4d(sa, &p.tmap_dy, &full_bar[stage], co_blk * 128, x0, y0, b);
        tma_load_4d(sa + kWgBoxDy, &p.tmap_dy, &full_bar[stage], co_blk * 128 + 64, x0, y0, b);
#pragma unroll
        for (int c = 0; c < kNC; ++c)
          tma_load_4d(sa + 2 * kWgBoxDy + c * kWgBoxX, &p.tmap_x, &full_bar[stage], (ci_blk * kNC + c) * 64, x0 - 1,
                      y0 + kh - 1, b);
        if (++stage == kWgStages) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {  // ---------------- MMA issuer
      // instruction descriptor: fp16 x fp16 -> fp32, M = 128, N = 64 * kNC, A and B MN-major (bits 15 / 16)
      const uint32_t idesc = make_idesc_f16(128, 64 * kNC, false) | (1u << 15) | (1u << 16);
      int stage = 0;
      uint32_t phase = 0;
      for (int t = t0; t < t1; ++t) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t sa = smem_u32(ring + stage * kWgStage);
        const uint32_t sx = sa + 2 * kWgBoxDy;
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          if (!((kw_mask >> kw) & 1u)) continue;
#pragma unroll
          for (int h = 0; h < kWgTileH; ++h) {
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {  // 16 pixels of image row h per MMA; x shifted by kw pixels inside its halo
              const uint64_t da = make_mnmajor_desc(sa + (h * kWgTileW + ks * 16) * 128, kWgBoxDy, 1024);
              const uint64_t db = make_mnmajor_desc(sx + (h * (kWgTileW + 2) + kw + ks * 16) * 128, kWgBoxX, 1024);
              umma_f16(tmem_base + kw * (64 * kNC), da, db, idesc, (t > t0 || h > 0 || ks > 0) ? 1u : 0u);
            }
          }
        }
        umma_commit(&empty_bar[stage]);
        if (++stage == kWgStages) {
          stage = 0;
          phase ^= 1u;
        }
      }
      umma_commit(&acc_bar);
    }
  } else {
    // ---------------- epilogue: warp w reads TMEM lanes 32 * (w % 4) ... + 31 = output channels of this block
    mbar_wait_parked(&acc_bar, 0);
    tc_fence_after();
    const int q = warp & 3;
    const int co = co_blk * 128 + q * 32 + lane;
    const uint32_t trow = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const int ci0 = ci_blk * (64 * kNC);
    for (int kw = 0; kw < 3; ++kw) {
      if (!((kw_mask >> kw) & 1u)) continue;
      float* dst = p.dw + (long long)img * p.img_stride + (long long)co * p.co_stride + (long long)(kh * 3 + kw) * p.tap_stride + ci0;
#pragma unroll
      for (int c16 = 0; c16 < 4 * kNC; ++c16) {
        if (ci0 + c16 * 16 >= p.cin) break;  // warp-uniform: channels past cin were zero-filled by TMA
        uint32_t v[16];
        tmem_ld16(trow + kw * (64 * kNC) + c16 * 16, v);
        tmem_ld_wait16(v);
        if (co < p.cout) {
#pragma unroll
          for (int j = 0; j < 16; ++j) atomicAdd(dst + c16 * 16 + j, __uint_as_float(v[j]));
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kWgTmemCols);
  }
}

}  // namespace b200ir

using namespace b200ir;

static int wgrad_launch(const b200ir_view* xv, const void* dy, float* dw, int B, int H, int W, int cout, uint32_t tap_mask,
                        bool per_image, void* stream) {
  B200IR_REQUIRE(xv && xv->ptr && dy && dw && B > 0 && H > 0 && W > 0 && xv->b == B, "conv_wgrad: bad arguments");
  const int cin = xv->c;
  tap_mask &= 0x1FFu;
  B200IR_REQUIRE(tap_mask != 0, "conv_wgrad: empty tap mask");
  B200IR_REQUIRE(cin > 0 && cin % 16 == 0 && cout > 0 && cout % 8 == 0,
                 "conv_wgrad: cin=%d must be a multiple of 16 and cout=%d of 8", cin, cout);
  B200IR_REQUIRE(((reinterpret_cast<uintptr_t>(xv->ptr) | reinterpret_cast<uintptr_t>(dy)) & 15) == 0 &&
                     xv->stride_w % 8 == 0 && xv->stride_h % 8 == 0 && xv->stride_b % 8 == 0,
                 "conv_wgrad: operands must be 16-byte aligned, view strides multiples of 8 elements");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int sms = num_sms();
  if (sms <= 0) return 1;
  WgradParams p = {};
  {
    cuuint32_t box[4] = {64u, (cuuint32_t)kWgTileW, (cuuint32_t)kWgTileH, 1u};
    cuuint32_t box_x[4] = {64u, (cuuint32_t)kWgTileW + 2u, (cuuint32_t)kWgTileH, 1u};
    cuuint64_t dx[4] = {(cuuint64_t)cin, (cuuint64_t)xv->w, (cuuint64_t)xv->h, (cuuint64_t)B};
    cuuint64_t sx[3] = {(cuuint64_t)xv->stride_w * 2, (cuuint64_t)xv->stride_h * 2, (cuuint64_t)xv->stride_b * 2};
    if (encode_map(&p.tmap_x, xv->ptr, 4, dx, sx, box_x, CU_TENSOR_MAP_SWIZZLE_128B, "conv_wgrad(x)")) return 1;
    cuuint64_t dd[4] = {(cuuint64_t)cout, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t sd[3] = {(cuuint64_t)cout * 2, (cuuint64_t)W * cout * 2, (cuuint64_t)H * W * cout * 2};
    if (encode_map(&p.tmap_dy, dy, 4, dd, sd, box, CU_TENSOR_MAP_SWIZZLE_128B, "conv_wgrad(dy)")) return 1;
  }
  p.dw = dw;
  p.cin = cin;
  p.cout = cout;
  p.tap_mask = tap_mask;
  for (int kh = 0; kh < 3; ++kh)
    if ((tap_mask >> (kh * 3)) & 7u) p.kh_list[p.num_kh++] = kh;
  p.tiles_x = (W + kWgTileW - 1) / kWgTileW;
  p.tiles_y = (H + kWgTileH - 1) / kWgTileH;
  p.num_tiles = B * p.tiles_x * p.tiles_y;
  // Channel blocks are 128 (cout) x 64 or 128 (cin); the TMA boxes of a ragged last block are zero-filled past the
  // tensor's channel extent and the epilogue skips those rows / columns, so small layers (32, 64 channels) run on the
  // same kernel with part of the MMA idle.
  const int nc = (cin > 64) ? 2 : 1;
  p.ci_blocks = (cin + 64 * nc - 1) / (64 * nc);
  const int units = ((cout + 127) / 128) * p.ci_blocks * p.num_kh;
  // One wave of CTAs (one per SM: the ring takes the whole shared memory) and at least ~16 pixel tiles per CTA, so that
  // the fixed cost of a CTA -- 128 x N x 3 fp32 atomics into dW -- stays below its MMA time (measured sweep:
  // tools/sweep_wgrad.py; two waves were 20-45 % slower on every layer of the B = 64 step).
  p.per_image = per_image ? 1 : 0;
  if (per_image) {  // compact per-image output of the single tap in the mask: [B][cout][cin]
    p.img_stride = (long long)cout * cin;
    p.co_stride = cin;
    p.tap_stride = 0;
  } else {
    p.img_stride = 0;
    p.co_stride = 9LL * cin;
    p.tap_stride = cin;
  }
  const int tiles_cta = per_image ? p.tiles_x * p.tiles_y : p.num_tiles;  // pixel tiles one (block, image) unit owns
  int splits = per_image ? sms / (units * B) : sms / units;
  if (splits > tiles_cta / 16) splits = tiles_cta / 16;
  {
    // tuning switch (tools/sweep_wgrad.py), read once: a positive integer forces the split count; 1 = deterministic
    // (one CTA per dW block, no cross-CTA atomics: the fp32 summation order is then fixed run to run)
    static int forced = -1;
    if (forced < 0) {
      const char* e = getenv("B200IR_WGRAD_SPLITS");
      const int v = e ? atoi(e) : 0;
      forced = v > 0 ? v : 0;
    }
    if (forced > 0) splits = forced;
  }
  if (splits > tiles_cta) splits = tiles_cta;
  if (splits < 1) splits = 1;
  p.splits = splits;
  const size_t dw_bytes = per_image ? (size_t)B * cout * cin * sizeof(float) : (size_t)cout * 9 * cin * sizeof(float);
  if (cudaMemsetAsync(dw, 0, dw_bytes, st) != cudaSuccess) {
    set_error("conv_wgrad: cudaMemsetAsync failed");
    return 1;
  }
  auto launch = [&](auto kernel, int smem) -> int {
    if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) {
      set_error("conv_wgrad: cudaFuncSetAttribute failed");
      return 1;
    }
    kernel<<<dim3(units * splits, per_image ? B : 1), kWgThreads, smem, st>>>(p);
    return 0;
  };
  if (nc == 2) {
    if (launch(conv_wgrad_kernel<2>, WgCfg<2>::kStages * WgCfg<2>::kStage + 1024)) return 1;
  } else {
    if (launch(conv_wgrad_kernel<1>, WgCfg<1>::kStages * WgCfg<1>::kStage + 1024)) return 1;
  }
  return check_launch("conv_wgrad");
}

extern "C" int b200ir_conv_wgrad_view(const b200ir_view* xv, const void* dy, float* dw, int B, int H, int W, int cout,
                                      uint32_t tap_mask, void* stream) {
  return wgrad_launch(xv, dy, dw, B, H, W, cout, tap_mask, false, stream);
}

extern "C" int b200ir_gram_batched(const void* x, const void* dy, float* out, int B, int H, int W, int cin, int cout,
                                   void* stream) {
  B200IR_REQUIRE(x && dy && out && B > 0 && B <= 65535 && H > 0 && W > 0 && cin > 0, "gram_batched: bad arguments");
  b200ir_view v = {};
  v.ptr = x;
  v.c = cin, v.w = W, v.h = H, v.b = B;
  v.stride_w = cin, v.stride_h = (int64_t)W * cin, v.stride_b = (int64_t)H * W * cin;
  return wgrad_launch(&v, dy, out, B, H, W, cout, 1u << 4, true, stream);
}

extern "C" int b200ir_conv_wgrad(const void* x, const void* dy, float* dw, int B, int H, int W, int cin, int cout,
                                 void* stream) {
  B200IR_REQUIRE(x && B > 0 && H > 0 && W > 0 && cin > 0, "conv_wgrad: bad arguments");
  b200ir_view v = {};
  v.ptr = x;
  v.c = cin, v.w = W, v.h = H, v.b = B;
  v.stride_w = cin, v.stride_h = (int64_t)W * cin, v.stride_b = (int64_t)H * W * cin;
  return b200ir_conv_wgrad_view(&v, dy, dw, B, H, W, cout, 0x1FFu, stream);
}
