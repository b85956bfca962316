// Implicit-GEMM convolution for sm_100a: TMA box loads (im2col by shifted boxes with out-of-bounds zero fill)
// -> 128B/64B-swizzled shared memory -> tcgen05.mma (M=128, N=block_n, K=16, fp16 x fp16 -> fp32 in TMEM)
// -> tcgen05.ld epilogue with bias / demodulation / noise / leaky-ReLU / residual fused.
//
// Persistent, warp-specialised: warp 0 = TMA producer, warp 1 = MMA issuer (+ TMEM owner), warps 2..5 =
// epilogue (one TMEM lane quarter each).  Two TMEM accumulator stages let the epilogue of tile i overlap the
// main loop of tile i+1.
//
// Reference semantics implemented here: see include/b200ir.h (b200ir_conv_igemm).
#pragma once
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "host_common.h"
#include "ptx.cuh"

namespace b200ir {

static constexpr int kBlockM = 128;
// warp 0 producer, warp 1 MMA issuer, then kEpiGroups groups of 4 epilogue warps (one warp per TMEM lane quarter) that
// drain tiles round-robin.  The specialised epilogues (<= 105 registers) run three groups: on the low-K layers the
// epilogue of a tile takes longer than its main loop, and the third group closes most of that gap.  The run-time generic
// epilogue (168 registers) stays at two so that it fits the register file.
template <int EPI>
struct EpiCfg {
  static constexpr int kGroups = (EPI >= 0) ? 3 : 2;
  static constexpr int kThreads = 64 + 128 * kGroups;
  // row kernel: a fourth epilogue group was measured and buys nothing (the cout = 64 layers are bound by the shared-
  // memory pipe -- MMA operand reads plus the epilogue's table reads -- not by epilogue latency)
  static constexpr int kRowGroups = kGroups;
  static constexpr int kRowThreads = 64 + 128 * kRowGroups;
};
static constexpr int kRowTable = 5 * 64;  // floats per epilogue group in the row kernel (block_n <= 64)
static constexpr int kMaxEpiGroups = 3;
static constexpr int kDemodTable = 2560;  // floats per epilogue group: per-tile tables [demod | out_scale | rgb_w x3]
static constexpr int kMaxBias = 512;
static constexpr int kMaxStages = 8;
static constexpr int kMaxAccStages = 16;  // TMEM accumulator ring: as many 128 x block_n tiles as fit in 512 columns

struct alignas(64) ConvParams {
  CUtensorMap tmap_a[B200IR_MAX_VIEWS];
  CUtensorMap tmap_b;
  int tiles_w, tiles_h, tiles_b, tiles_n, num_tiles;
  int tile_w, tile_h, tile_b;
  int block_n, block_k, k_chunks, num_taps;
  int m_w, m_h, m_b;
  int stages;
  int b_resident;  // generic kernel: all weight tiles (num_taps * k_chunks of block_n rows) stay in shared memory, the
                   // stage ring carries activations only (single N-tile layers whose weights fit next to >= 3 stages)
  uint32_t idesc;
  uint32_t idesc_n[3];  // row mode: instruction descriptors for N = 1, 2, 3 x block_n
  uint32_t tmem_cols;
  int acc_stages, acc_shift;
  // row mode (conv_row_kernel): 3x3 stride-1 conv, tile = 128 consecutive pixels of one row, weights resident in
  // shared memory, each input row segment loaded once (with a 1-pixel halo) and reused for 3 kw shifts x 3 output rows
  int row_R, row_chunks, row_items, row_slots, row_slot_bytes, row_w_bytes, desc_mode;
  int row_balanced;  // 1: every CTA takes an equal share of the (image, segment, output row) list instead of whole items
  int smem_demod;  // 1: per-tile demod table staged in shared memory
  int smem_aux;    // 1: out_scale / rgb_w tables staged behind it (4 more tables of tile_b * block_n floats)
  int st256;       // 1: fp16 output rows are 32-byte aligned -> 256-bit stores
  float act_gain;  // sqrt(2) when act is set (folded into the bias / demod / noise terms), else 1; the specialised
                   // epilogues (epi >= 0) also fold the residual scale into it
  float slope;     // leaky-ReLU slope (0.2), 1.0 when the layer has no activation: v = max(v, slope * v)
  int epi;         // index into kEpiProfiles (compile-time specialised epilogue) or -1 for the run-time generic one
  int8_t tap_view[B200IR_MAX_TAPS], tap_dx[B200IR_MAX_TAPS], tap_dy[B200IR_MAX_TAPS];
  // epilogue
  void* out;
  int out_fp32;
  long long out_sx, out_sy, out_sb;
  int out_c_off, out_x_mul, out_x_off, out_y_mul, out_y_off;
  int cout;
  const float* bias;
  const float* demod;
  const float* noise;
  const float* noise_gain;
  long long noise_sb, noise_sy;
  int act;
  int res_mode;
  const __half* res;
  long long res_sx, res_sy, res_sb;
  int res_w, res_h;
  float res_scale;
  float res_mul;  // weight of the residual term: v * res_scale + res * res_mul (legacy (v + res) * s: res_mul = s)
  int ps_r;       // pixel-shuffle factor of the store (0: off)
  int ps_shift;   // log2(ps_c) when ps_c is a power of two, else -1
  int ps_c;       // channels per sub-pixel phase: output column n belongs to phase n / ps_c, channel n % ps_c
  int bias_c;     // entries of the bias vector (cout, or ps_c when the phases share one bias: folded ConvUpLayer)
  int demod_c;    // demod / out_scale tables are indexed with channel n % demod_c (phases share one table row)
  // up-fold border corrections (see epilogue_upfold): fp32 [m_b][2*m_w][ps_c] (top, bottom) / [m_b][2*m_h][ps_c]
  const float* corr_top;
  const float* corr_bot;
  const float* corr_left;
  const float* corr_right;
  uint32_t tap_mask[8];  // per N-tile: taps to execute (bit t); tiles whose weight block for a tap is all zero skip it
  const float* out_scale;
  const float* rgb_w;
  float* rgb_part;
  long long rgb_plane, rgb_image;  // rgb_h*rgb_w_px, m_b*3*rgb_plane
  int rgb_w_px;
  int no_store;
  int dbg_skip_epi;  // profiling aid (env B200IR_DBG_SKIP_EPI): epilogue only recycles the accumulators
  int row_peek;      // row kernel: the MMA warp peeks at the next row's barriers inside its MMA sequence (env B200IR_ROW_PEEK)
  int epi_wait_ns;   // back-off of the epilogue's accumulator-full wait (env B200IR_EPI_WAIT_NS; 0 = parked try_wait)
  int epi_pipe;      // fast epilogues without global operands: keep the next chunk's TMEM load in flight (host heuristic)
  int epi_split;     // 256-column accumulators (two ring stages): every tile is drained as two 128-column halves and the
                     // half-tiles go round-robin over all three epilogue groups (otherwise the third group idles)
  int w_img_rows;    // per-image weights: rows of the weight matrix per image (= cout), 0 = one matrix for all images
  // CTA pairs (conv_igemm_kernel<.., kPair = true>): two CTAs share one 256 x block_n MMA per K step
  CUtensorMap tmap_b2;  // weight tile of block_n / 2 rows (each CTA of a pair stages half of B)
  int pair;             // 1: launched as clusters of two
  int pair_tiles;       // ceil(M-tiles / 2) * tiles_n work items of a pair
  uint32_t idesc_pair;  // instruction descriptor with M = 256
};

struct TileCoord {
  int x0, y0, b0, n0;
};

__device__ __forceinline__ TileCoord decode_tile(const ConvParams& p, int tile) {
  TileCoord t;
  int n_tile = tile % p.tiles_n;
  int m = tile / p.tiles_n;
  int xw = m % p.tiles_w;
  m /= p.tiles_w;
  int yh = m % p.tiles_h;
  int bb = m / p.tiles_h;
  t.x0 = xw * p.tile_w;
  t.y0 = yh * p.tile_h;
  t.b0 = bb * p.tile_b;
  t.n0 = n_tile * p.block_n;
  return t;
}

__device__ __forceinline__ void unpack_half8(const uint4& q, float* f) {
  const __half2* h = reinterpret_cast<const __half2*>(&q);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 x = __half22float2(h[i]);
    f[2 * i] = x.x;
    f[2 * i + 1] = x.y;
  }
}

// Per-thread addressing of one output position, computed (and the noise value fetched) BEFORE the accumulator is
// ready so that none of it sits on the MMA -> epilogue critical path.
struct EpiRow {
  long long out_off;
  long long chan_off;  // b * cout + n0: row of the per-image tables (out_scale)
  long long rgbw_off;  // b * 3 * cout + n0
  long long rgb_off;   // offset of (n-tile, b, 0, yo, xo) in rgb_part
  float nz;
  const __half* r00;
  const __half* r01;
  const __half* r10;
  const __half* r11;
  float wy0, wy1, wx0, wx1;
};

__device__ __forceinline__ EpiRow epi_setup(const ConvParams& p, int x, int y, int b, int n0, bool valid, float gain) {
  EpiRow r;
  int xo = x * p.out_x_mul + p.out_x_off;
  int yo = y * p.out_y_mul + p.out_y_off;
  int nch = n0;  // channel offset of this N-tile in the output
  if (p.ps_r) {  // nn.PixelShuffle / transposed-conv phases fused into the store: base of sub-pixel (0, 0), channel 0;
    xo = x * p.ps_r;  // the per-chunk phase offset is added by ps_offset()
    yo = y * p.ps_r;
    nch = 0;
  }
  r.out_off = (long long)b * p.out_sb + (long long)yo * p.out_sy + (long long)xo * p.out_sx + p.out_c_off + nch;
  r.chan_off = (long long)b * p.cout + n0;
  r.rgbw_off = (long long)b * 3 * p.cout + n0;
  r.rgb_off = (long long)(n0 / p.block_n) * p.rgb_image + (long long)b * 3 * p.rgb_plane + (long long)yo * p.rgb_w_px + xo;
  r.nz = 0.f;
  if (valid && p.noise != nullptr) r.nz = __ldg(p.noise + b * p.noise_sb + yo * p.noise_sy + xo);  // scaled by the caller later
  r.r00 = r.r01 = r.r10 = r.r11 = nullptr;
  r.wy0 = r.wy1 = r.wx0 = r.wx1 = 0.f;
  if (valid && p.res_mode == 1) {
    r.r00 = p.res + (long long)b * p.res_sb + (long long)yo * p.res_sy + (long long)xo * p.res_sx + n0;
  } else if (valid && p.res_mode == 2) {
    // F.interpolate(scale 2, bilinear, align_corners=False): even 2k -> .25*x[k-1] + .75*x[k], odd 2k+1 ->
    // .75*x[k] + .25*x[k+1], indices clamped to the tensor.
    const int ky = yo >> 1, kx = xo >> 1;
    int ya, yb, xa, xb;
    if (yo & 1) { ya = ky; yb = min(ky + 1, p.res_h - 1); r.wy0 = 0.75f; r.wy1 = 0.25f; }
    else        { ya = max(ky - 1, 0); yb = ky; r.wy0 = 0.25f; r.wy1 = 0.75f; }
    if (xo & 1) { xa = kx; xb = min(kx + 1, p.res_w - 1); r.wx0 = 0.75f; r.wx1 = 0.25f; }
    else        { xa = max(kx - 1, 0); xb = kx; r.wx0 = 0.25f; r.wx1 = 0.75f; }
    const __half* rb = p.res + (long long)b * p.res_sb + n0;
    r.r00 = rb + (long long)ya * p.res_sy + (long long)xa * p.res_sx;
    r.r01 = rb + (long long)ya * p.res_sy + (long long)xb * p.res_sx;
    r.r10 = rb + (long long)yb * p.res_sy + (long long)xa * p.res_sx;
    r.r11 = rb + (long long)yb * p.res_sy + (long long)xb * p.res_sx;
  }
  return r;
}

// Output offset (elements) of GEMM column n = n0 + c0 relative to the sub-pixel (0,0) / channel 0 base when the store is
// a pixel shuffle: phase t = n / ps_c -> sub-pixel (t / r, t % r), channel n % ps_c.  0 when ps_r == 0 (then the base
// already contains n0 and the caller adds c0).
__device__ __forceinline__ long long ps_offset(const ConvParams& p, int n0, int c0) {
  const int n = n0 + c0;
  const int t = (p.ps_shift >= 0) ? (n >> p.ps_shift) : (n / p.ps_c);  // ps_c is a power of two on the hot paths
  const int ty = (p.ps_r == 2) ? (t >> 1) : (t / p.ps_r);
  const int tx = t - ty * p.ps_r;
  return (long long)ty * p.out_sy + (long long)tx * p.out_sx + (n - t * p.ps_c) - c0;
}

// Drains this thread's row of one 128 x block_n accumulator tile, 16 columns at a time (columns c_begin, c_begin +
// c_step, ...: two warps share a TMEM lane quarter).  Per chunk the TMEM load is issued first, the operands that do not
// depend on it (bias / demod from shared memory, residual from global) are fetched while it is in flight.
//   s_bias  : shared memory, bias[n0 ...] (zeros when the layer has no bias)
//   s_demod : shared memory, demod[b][n0 ...] for this row's image, or nullptr
//   g_demod : global fallback for demod (used when the per-tile table does not fit), or nullptr
//   s_aux   : shared memory, [out_scale | rgb_w[0] | rgb_w[1] | rgb_w[2]] rows of this row's image, aux_stride floats
//             apart, or nullptr (then out_scale / rgb_w come from global memory: with ~227 KB of dynamic shared memory
//             there is almost no L1 left, so every such load is an L2 round trip -- measured as the top stall)
// All shared-memory tables are passed as 32-bit shared addresses and read with ld.shared: selecting between a shared
// and a global POINTER makes the compiler emit generic loads, which take the long L1TEX path (measured).
__device__ __forceinline__ float4 lds_f4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}

//   s_bias  : shared address of bias[n0 ...] (zeros when the layer has no bias); 0 -> g_bias (global, wide layers)
//   s_demod : shared address of demod[b][n0 ...] for this row's image, or 0
//   s_aux   : shared address of the [out_scale | rgb_w[0] | rgb_w[1] | rgb_w[2]] rows, aux_stride floats apart, or 0
__device__ __forceinline__ void epilogue_tile(const ConvParams& p, uint32_t taddr, uint64_t* full_bar,
                                              uint32_t full_phase, EpiRow r, bool valid, float gain, uint32_t s_bias,
                                              const float* g_bias, uint32_t s_demod, const float* g_demod,
                                              uint32_t s_aux, int aux_stride, int c_begin, int c_step, int n0) {
  mbar_wait_backoff(full_bar, full_phase, p.epi_wait_ns);
  tc_fence_after();
  if (p.dbg_skip_epi) return;
  r.nz *= gain;  // the noise load was issued in epi_setup, long before this first use
  float rgb_acc[3] = {0.f, 0.f, 0.f};
  for (int c0 = c_begin; c0 < p.block_n; c0 += c_step) {
    uint32_t raw[16];
    tmem_ld16(taddr + c0, raw);
    float4 bs[4], dm[4];
    if (s_bias != 0) {
#pragma unroll
      for (int j = 0; j < 4; ++j) bs[j] = lds_f4(s_bias + (c0 + 4 * j) * 4);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) bs[j] = __ldg(reinterpret_cast<const float4*>(g_bias + c0) + j);
    }
    if (s_demod != 0) {
#pragma unroll
      for (int j = 0; j < 4; ++j) dm[j] = lds_f4(s_demod + (c0 + 4 * j) * 4);
    } else if (g_demod != nullptr && valid) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        dm[j] = __ldg(reinterpret_cast<const float4*>(g_demod + c0) + j);
        dm[j].x *= p.act_gain; dm[j].y *= p.act_gain; dm[j].z *= p.act_gain; dm[j].w *= p.act_gain;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) dm[j] = make_float4(p.act_gain, p.act_gain, p.act_gain, p.act_gain);
    }
    uint4 ra[2], rb[2], rc[2], rd[2];
    if (valid && p.res_mode != 0) {
      ra[0] = __ldg(reinterpret_cast<const uint4*>(r.r00 + c0));
      ra[1] = __ldg(reinterpret_cast<const uint4*>(r.r00 + c0) + 1);
      if (p.res_mode == 2) {
        rb[0] = __ldg(reinterpret_cast<const uint4*>(r.r01 + c0));
        rb[1] = __ldg(reinterpret_cast<const uint4*>(r.r01 + c0) + 1);
        rc[0] = __ldg(reinterpret_cast<const uint4*>(r.r10 + c0));
        rc[1] = __ldg(reinterpret_cast<const uint4*>(r.r10 + c0) + 1);
        rd[0] = __ldg(reinterpret_cast<const uint4*>(r.r11 + c0));
        rd[1] = __ldg(reinterpret_cast<const uint4*>(r.r11 + c0) + 1);
      }
    }
    tmem_ld_wait16(raw);
    if (valid) {
      // bias / demod / noise arrive pre-multiplied by the activation gain (sqrt 2) when act is set, so the
      // leaky-ReLU is just max(v, 0.2 v)
      float v[16];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        v[4 * j] = __uint_as_float(raw[4 * j]) * dm[j].x + (bs[j].x + r.nz);
        v[4 * j + 1] = __uint_as_float(raw[4 * j + 1]) * dm[j].y + (bs[j].y + r.nz);
        v[4 * j + 2] = __uint_as_float(raw[4 * j + 2]) * dm[j].z + (bs[j].z + r.nz);
        v[4 * j + 3] = __uint_as_float(raw[4 * j + 3]) * dm[j].w + (bs[j].w + r.nz);
      }
      if (p.act) {
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], p.slope * v[j]);
      }
      if (p.res_mode == 1) {
        float f[16];
        unpack_half8(ra[0], f);
        unpack_half8(ra[1], f + 8);
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = v[j] * p.res_scale + f[j] * p.res_mul;
      } else if (p.res_mode == 2) {
        float fa[16], fb[16], fc[16], fd[16];
        unpack_half8(ra[0], fa); unpack_half8(ra[1], fa + 8);
        unpack_half8(rb[0], fb); unpack_half8(rb[1], fb + 8);
        unpack_half8(rc[0], fc); unpack_half8(rc[1], fc + 8);
        unpack_half8(rd[0], fd); unpack_half8(rd[1], fd + 8);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float up = r.wy0 * (r.wx0 * fa[j] + r.wx1 * fb[j]) + r.wy1 * (r.wx0 * fc[j] + r.wx1 * fd[j]);
          v[j] = v[j] * p.res_scale + up * p.res_mul;
        }
      }
      if (p.rgb_w != nullptr) {
        if (s_aux != 0) {
#pragma unroll
          for (int o = 0; o < 3; ++o) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float4 w4 = lds_f4(s_aux + ((1 + o) * aux_stride + c0 + 4 * j) * 4);
              rgb_acc[o] = fmaf(v[4 * j], w4.x, rgb_acc[o]);
              rgb_acc[o] = fmaf(v[4 * j + 1], w4.y, rgb_acc[o]);
              rgb_acc[o] = fmaf(v[4 * j + 2], w4.z, rgb_acc[o]);
              rgb_acc[o] = fmaf(v[4 * j + 3], w4.w, rgb_acc[o]);
            }
          }
        } else {
#pragma unroll
          for (int o = 0; o < 3; ++o) {
            const float4* wp = reinterpret_cast<const float4*>(p.rgb_w + r.rgbw_off + (long long)o * p.cout + c0);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float4 w4 = __ldg(wp + j);
              rgb_acc[o] = fmaf(v[4 * j], w4.x, rgb_acc[o]);
              rgb_acc[o] = fmaf(v[4 * j + 1], w4.y, rgb_acc[o]);
              rgb_acc[o] = fmaf(v[4 * j + 2], w4.z, rgb_acc[o]);
              rgb_acc[o] = fmaf(v[4 * j + 3], w4.w, rgb_acc[o]);
            }
          }
        }
      }
      if (p.out_scale != nullptr) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 s4 = (s_aux != 0) ? lds_f4(s_aux + (c0 + 4 * j) * 4)
                                         : __ldg(reinterpret_cast<const float4*>(p.out_scale + r.chan_off + c0) + j);
          v[4 * j] *= s4.x; v[4 * j + 1] *= s4.y; v[4 * j + 2] *= s4.z; v[4 * j + 3] *= s4.w;
        }
      }
      if (p.no_store) {
      } else if (p.out_fp32) {
        const long long pso = p.ps_r ? ps_offset(p, n0, c0) : 0;
        float4* op = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + r.out_off + c0 + pso);
#pragma unroll
        for (int j = 0; j < 4; ++j) op[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
      } else {
        uint32_t pk[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          __half2 h = f2h2_sat(v[2 * j], v[2 * j + 1]);
          pk[j] = *reinterpret_cast<uint32_t*>(&h);
        }
        const long long pso = p.ps_r ? ps_offset(p, n0, c0) : 0;
        __half* op = reinterpret_cast<__half*>(p.out) + r.out_off + c0 + pso;
        if (p.st256) {  // one full 32-byte sector per thread and instruction (no partial-sector writes in L2)
          asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(op), "r"(pk[0]), "r"(pk[1]),
                       "r"(pk[2]), "r"(pk[3]), "r"(pk[4]), "r"(pk[5]), "r"(pk[6]), "r"(pk[7])
                       : "memory");
        } else {
          reinterpret_cast<uint4*>(op)[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          reinterpret_cast<uint4*>(op)[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
        }
      }
    }
  }
  if (p.rgb_w != nullptr && valid) {
#pragma unroll
    for (int o = 0; o < 3; ++o) p.rgb_part[r.rgb_off + o * p.rgb_plane] = rgb_acc[o];
  }
}

// ------------------------------------------------------------------------------------------ specialised epilogues
// The generic epilogue above tests every feature flag per 16-column chunk (measured: ~550 SASS instructions per chunk,
// which made the epilogue -- not the MMA main loop -- the pace setter of every layer with K <= 1152).  The layers of
// the network use six feature combinations; each gets an epilogue with the flags as template constants.
enum : int { F_DEMOD = 1, F_NOISE = 2, F_RES1 = 4, F_RES2 = 8, F_RGB = 16, F_NOSTORE = 32, F_UPFOLD = 64 };
static constexpr int kNumEpiProfiles = 7;
__host__ __device__ constexpr int epi_profile_flags(int i) {
  return i == 0   ? 0
         : i == 1 ? F_RES1
         : i == 2 ? F_RES2
         : i == 3 ? F_DEMOD
         : i == 4 ? (F_DEMOD | F_NOISE | F_RGB)
         : i == 5 ? (F_DEMOD | F_NOISE | F_RGB | F_NOSTORE)
                  : (F_RES2 | F_UPFOLD);
}

struct FastRow {
  __half* out;  // output row of this position, first channel of the N-tile
  float* rgb;   // partial ToRGB plane element of this position (channel 0)
  float nz;
  const __half* r00;
  const __half* r01;
  const __half* r10;
  const __half* r11;
  float w00, w01, w10, w11;  // bilinear weights x residual scale
};

template <int F>
__device__ __forceinline__ FastRow fast_setup(const ConvParams& p, int x, int y, int b, int n0, bool valid) {
  FastRow r;
  int xo = x * p.out_x_mul + p.out_x_off;
  int yo = y * p.out_y_mul + p.out_y_off;
  int nch = n0;
  if (p.ps_r) {  // see epi_setup
    xo = x * p.ps_r;
    yo = y * p.ps_r;
    nch = 0;
  }
  r.out = reinterpret_cast<__half*>(p.out) + (long long)b * p.out_sb + (long long)yo * p.out_sy +
          (long long)xo * p.out_sx + p.out_c_off + nch;
  r.rgb = nullptr;
  if (F & F_RGB)
    r.rgb = p.rgb_part + (long long)(n0 / p.block_n) * p.rgb_image + (long long)b * 3 * p.rgb_plane +
            (long long)yo * p.rgb_w_px + xo;
  r.nz = 0.f;
  if ((F & F_NOISE) && valid) r.nz = __ldg(p.noise + b * p.noise_sb + yo * p.noise_sy + xo);
  r.r00 = r.r01 = r.r10 = r.r11 = nullptr;
  r.w00 = r.w01 = r.w10 = r.w11 = 0.f;
  if ((F & F_RES1) && valid) {
    r.r00 = p.res + (long long)b * p.res_sb + (long long)yo * p.res_sy + (long long)xo * p.res_sx + n0;
    r.w00 = p.res_mul;
  }
  if ((F & F_RES2) && valid) {
    const int ky = yo >> 1, kx = xo >> 1;
    int ya, yb, xa, xb;
    float wy0, wy1, wx0, wx1;
    if (yo & 1) { ya = ky; yb = min(ky + 1, p.res_h - 1); wy0 = 0.75f; wy1 = 0.25f; }
    else        { ya = max(ky - 1, 0); yb = ky; wy0 = 0.25f; wy1 = 0.75f; }
    if (xo & 1) { xa = kx; xb = min(kx + 1, p.res_w - 1); wx0 = 0.75f; wx1 = 0.25f; }
    else        { xa = max(kx - 1, 0); xb = kx; wx0 = 0.25f; wx1 = 0.75f; }
    const __half* rb = p.res + (long long)b * p.res_sb + n0;
    r.r00 = rb + (long long)ya * p.res_sy + (long long)xa * p.res_sx;
    r.r01 = rb + (long long)ya * p.res_sy + (long long)xb * p.res_sx;
    r.r10 = rb + (long long)yb * p.res_sy + (long long)xa * p.res_sx;
    r.r11 = rb + (long long)yb * p.res_sy + (long long)xb * p.res_sx;
    wy0 *= p.res_mul;
    wy1 *= p.res_mul;
    r.w00 = wy0 * wx0; r.w01 = wy0 * wx1; r.w10 = wy1 * wx0; r.w11 = wy1 * wx1;
  }
  return r;
}

// fp16 NHWC output with 32-byte aligned rows, bias (and demod / aux tables) in shared memory, gains pre-folded:
//   v = acc * (demod*g | g) + (bias*g + noise*gain*g);  v = max(v, slope*v);  v += res * w (w carries the residual scale)
// The epilogue as the two-CTAs-per-SM row kernels use it: one TMEM load per chunk, whole tile, no pipelining state (60-72
// registers; see epilogue_fast for the general form).
template <int F>
__device__ __forceinline__ void epilogue_fast_lean(const ConvParams& p, uint32_t taddr, uint64_t* full_bar,
                                              uint32_t full_phase, const FastRow& r, bool valid, float gain,
                                              uint32_t s_bias, uint32_t s_demod, uint32_t s_aux, int aux_stride,
                                              int n0) {
  mbar_wait_backoff(full_bar, full_phase, p.epi_wait_ns);
  tc_fence_after();
  if (p.dbg_skip_epi) return;
  const float nz = (F & F_NOISE) ? r.nz * gain : 0.f;
  const float ag = p.act_gain, slope = p.slope;
  float rgb0 = 0.f, rgb1 = 0.f, rgb2 = 0.f;
#pragma unroll 1
  for (int c0 = 0; c0 < p.block_n; c0 += 16) {
    uint32_t raw[16];
    tmem_ld16(taddr + c0, raw);
    float4 bs[4], dm[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) bs[j] = lds_f4(s_bias + (c0 + 4 * j) * 4);
    if (F & F_DEMOD) {
#pragma unroll
      for (int j = 0; j < 4; ++j) dm[j] = lds_f4(s_demod + (c0 + 4 * j) * 4);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) dm[j] = make_float4(ag, ag, ag, ag);
    }
    uint4 ra[2], rb[2], rc[2], rd[2];
    if ((F & (F_RES1 | F_RES2)) && valid) {
      ra[0] = __ldg(reinterpret_cast<const uint4*>(r.r00 + c0));
      ra[1] = __ldg(reinterpret_cast<const uint4*>(r.r00 + c0) + 1);
      if (F & F_RES2) {
        rb[0] = __ldg(reinterpret_cast<const uint4*>(r.r01 + c0));
        rb[1] = __ldg(reinterpret_cast<const uint4*>(r.r01 + c0) + 1);
        rc[0] = __ldg(reinterpret_cast<const uint4*>(r.r10 + c0));
        rc[1] = __ldg(reinterpret_cast<const uint4*>(r.r10 + c0) + 1);
        rd[0] = __ldg(reinterpret_cast<const uint4*>(r.r11 + c0));
        rd[1] = __ldg(reinterpret_cast<const uint4*>(r.r11 + c0) + 1);
      }
    }
    tmem_ld_wait16(raw);
    if (!valid) continue;
    float v[16];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      v[4 * j] = fmaf(__uint_as_float(raw[4 * j]), dm[j].x, bs[j].x + nz);
      v[4 * j + 1] = fmaf(__uint_as_float(raw[4 * j + 1]), dm[j].y, bs[j].y + nz);
      v[4 * j + 2] = fmaf(__uint_as_float(raw[4 * j + 2]), dm[j].z, bs[j].z + nz);
      v[4 * j + 3] = fmaf(__uint_as_float(raw[4 * j + 3]), dm[j].w, bs[j].w + nz);
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], slope * v[j]);
    if (F & F_RES1) {
      float f[16];
      unpack_half8(ra[0], f);
      unpack_half8(ra[1], f + 8);
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = fmaf(f[j], r.w00, v[j]);
    }
    if (F & F_RES2) {
      float f[16];
      unpack_half8(ra[0], f); unpack_half8(ra[1], f + 8);
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = fmaf(f[j], r.w00, v[j]);
      unpack_half8(rb[0], f); unpack_half8(rb[1], f + 8);
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = fmaf(f[j], r.w01, v[j]);
      unpack_half8(rc[0], f); unpack_half8(rc[1], f + 8);
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = fmaf(f[j], r.w10, v[j]);
      unpack_half8(rd[0], f); unpack_half8(rd[1], f + 8);
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = fmaf(f[j], r.w11, v[j]);
    }
    if (F & F_RGB) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 w0 = lds_f4(s_aux + (aux_stride + c0 + 4 * j) * 4);
        const float4 w1 = lds_f4(s_aux + (2 * aux_stride + c0 + 4 * j) * 4);
        const float4 w2 = lds_f4(s_aux + (3 * aux_stride + c0 + 4 * j) * 4);
        rgb0 = fmaf(v[4 * j], w0.x, rgb0); rgb1 = fmaf(v[4 * j], w1.x, rgb1); rgb2 = fmaf(v[4 * j], w2.x, rgb2);
        rgb0 = fmaf(v[4 * j + 1], w0.y, rgb0); rgb1 = fmaf(v[4 * j + 1], w1.y, rgb1); rgb2 = fmaf(v[4 * j + 1], w2.y, rgb2);
        rgb0 = fmaf(v[4 * j + 2], w0.z, rgb0); rgb1 = fmaf(v[4 * j + 2], w1.z, rgb1); rgb2 = fmaf(v[4 * j + 2], w2.z, rgb2);
        rgb0 = fmaf(v[4 * j + 3], w0.w, rgb0); rgb1 = fmaf(v[4 * j + 3], w1.w, rgb1); rgb2 = fmaf(v[4 * j + 3], w2.w, rgb2);
      }
      if (!(F & F_NOSTORE)) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 s4 = lds_f4(s_aux + (c0 + 4 * j) * 4);
          v[4 * j] *= s4.x; v[4 * j + 1] *= s4.y; v[4 * j + 2] *= s4.z; v[4 * j + 3] *= s4.w;
        }
      }
    }
    if (!(F & F_NOSTORE)) {
      uint32_t pk[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        __half2 h = f2h2_sat(v[2 * j], v[2 * j + 1]);
        pk[j] = *reinterpret_cast<uint32_t*>(&h);
      }
      const __half* op = r.out + c0 + (p.ps_r ? ps_offset(p, n0, c0) : 0);
      asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(op), "r"(pk[0]), "r"(pk[1]),
                   "r"(pk[2]), "r"(pk[3]), "r"(pk[4]), "r"(pk[5]), "r"(pk[6]), "r"(pk[7])
                   : "memory");
    }
  }
  if ((F & F_RGB) && valid) {
    r.rgb[0] = rgb0;
    r.rgb[p.rgb_plane] = rgb1;
    r.rgb[2 * p.rgb_plane] = rgb2;
  }
}

// kAllowPipe = false compiles the pipelined variant out: its second TMEM register buffer costs ~25 registers, and the
// row kernel's two-CTAs-per-SM configuration (cout <= 32, profiles 0 / 1) only fits with <= 73 registers per thread.
template <int F, bool kAllowPipe = true>
__device__ __forceinline__ void epilogue_fast(const ConvParams& p, uint32_t taddr, uint64_t* full_bar,
                                              uint32_t full_phase, const FastRow& r, bool valid, float gain,
                                              uint32_t s_bias, uint32_t s_demod, uint32_t s_aux, int aux_stride,
                                              int n0, int c_begin, int c_end) {
  mbar_wait_backoff(full_bar, full_phase, p.epi_wait_ns);
  tc_fence_after();
  if (p.dbg_skip_epi) return;
  const float nz = (F & F_NOISE) ? r.nz * gain : 0.f;
  const float ag = p.act_gain, slope = p.slope;
  const bool act = slope != 1.f;  // slope 1: no activation (max(v, v)); saves two instructions per element
  const int block_n = c_end;      // columns [c_begin, c_end) of the tile (the whole tile unless the epilogue is split)
  float rgb0 = 0.f, rgb1 = 0.f, rgb2 = 0.f;
  // pixel-shuffle store (transposed conv, folded ConvUpLayer): sub-pixel base and channel of the current chunk, advanced
  // by 16 channels per chunk instead of recomputing the phase split of n0 + c0 every time
  const int ps_r = p.ps_r, ps_c = p.ps_c;
  int ps_ch = 0, ps_t = 0;
  const __half* ps_base = r.out;
  if (ps_r) {
    const int nb = n0 + c_begin;
    ps_t = (p.ps_shift >= 0) ? (nb >> p.ps_shift) : (nb / ps_c);
    ps_ch = nb - ps_t * ps_c;
    const int ty = ps_t / ps_r;
    ps_base = r.out + (long long)ty * p.out_sy + (long long)(ps_t - ty * ps_r) * p.out_sx;
  }
  // everything after the accumulator chunk has arrived: scale / bias / activation / residual / ToRGB / store
  auto finish = [&](int c0, const uint32_t (&raw)[16], const uint4 (&ra)[2], const uint4 (&rb)[2], const uint4 (&rc)[2],
                    const uint4 (&rd)[2]) {
    if (valid) {
      float4 bs[4], dm[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) bs[j] = lds_f4(s_bias + (c0 + 4 * j) * 4);
      if (F & F_DEMOD) {
#pragma unroll
        for (int j = 0; j < 4; ++j) dm[j] = lds_f4(s_demod + (c0 + 4 * j) * 4);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) dm[j] = make_float4(ag, ag, ag, ag);
      }
      float v[16];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        v[4 * j] = fmaf(__uint_as_float(raw[4 * j]), dm[j].x, bs[j].x + nz);
        v[4 * j + 1] = fmaf(__uint_as_float(raw[4 * j + 1]), dm[j].y, bs[j].y + nz);
        v[4 * j + 2] = fmaf(__uint_as_float(raw[4 * j + 2]), dm[j].z, bs[j].z + nz);
        v[4 * j + 3] = fmaf(__uint_as_float(raw[4 * j + 3]), dm[j].w, bs[j].w + nz);
      }
      if (act) {
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], slope * v[j]);
      }
      if (F & F_RES1) {
        float f[16];
        unpack_half8(ra[0], f);
        unpack_half8(ra[1], f + 8);
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = fmaf(f[j], r.w00, v[j]);
      }
      if (F & F_RES2) {
        float f[16];
        unpack_half8(ra[0], f); unpack_half8(ra[1], f + 8);
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = fmaf(f[j], r.w00, v[j]);
        unpack_half8(rb[0], f); unpack_half8(rb[1], f + 8);
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = fmaf(f[j], r.w01, v[j]);
        unpack_half8(rc[0], f); unpack_half8(rc[1], f + 8);
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = fmaf(f[j], r.w10, v[j]);
        unpack_half8(rd[0], f); unpack_half8(rd[1], f + 8);
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = fmaf(f[j], r.w11, v[j]);
      }
      if (F & F_RGB) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 w0 = lds_f4(s_aux + (aux_stride + c0 + 4 * j) * 4);
          const float4 w1 = lds_f4(s_aux + (2 * aux_stride + c0 + 4 * j) * 4);
          const float4 w2 = lds_f4(s_aux + (3 * aux_stride + c0 + 4 * j) * 4);
          rgb0 = fmaf(v[4 * j], w0.x, rgb0); rgb1 = fmaf(v[4 * j], w1.x, rgb1); rgb2 = fmaf(v[4 * j], w2.x, rgb2);
          rgb0 = fmaf(v[4 * j + 1], w0.y, rgb0); rgb1 = fmaf(v[4 * j + 1], w1.y, rgb1); rgb2 = fmaf(v[4 * j + 1], w2.y, rgb2);
          rgb0 = fmaf(v[4 * j + 2], w0.z, rgb0); rgb1 = fmaf(v[4 * j + 2], w1.z, rgb1); rgb2 = fmaf(v[4 * j + 2], w2.z, rgb2);
          rgb0 = fmaf(v[4 * j + 3], w0.w, rgb0); rgb1 = fmaf(v[4 * j + 3], w1.w, rgb1); rgb2 = fmaf(v[4 * j + 3], w2.w, rgb2);
        }
        if (!(F & F_NOSTORE)) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 s4 = lds_f4(s_aux + (c0 + 4 * j) * 4);
            v[4 * j] *= s4.x; v[4 * j + 1] *= s4.y; v[4 * j + 2] *= s4.z; v[4 * j + 3] *= s4.w;
          }
        }
      }
      if (!(F & F_NOSTORE)) {
        uint32_t pk[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          __half2 h = f2h2_sat(v[2 * j], v[2 * j + 1]);
          pk[j] = *reinterpret_cast<uint32_t*>(&h);
        }
        const __half* op = ps_r ? ps_base + ps_ch : r.out + c0;
        asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(op), "r"(pk[0]), "r"(pk[1]),
                     "r"(pk[2]), "r"(pk[3]), "r"(pk[4]), "r"(pk[5]), "r"(pk[6]), "r"(pk[7])
                     : "memory");
      }
    }
    if (ps_r) {  // next chunk: 16 channels on, next sub-pixel when the channel block is complete
      ps_ch += 16;
      if (ps_ch >= ps_c) {
        ps_ch = 0;
        ++ps_t;
        const int ty = ps_t / ps_r;
        ps_base = r.out + (long long)ty * p.out_sy + (long long)(ps_t - ty * ps_r) * p.out_sx;
      }
    }
  };
  if ((F & (F_RES1 | F_RES2)) || !kAllowPipe || !p.epi_pipe) {
    // residual variants: the residual row(s) of a chunk are fetched from global memory while its TMEM load is in flight
#pragma unroll 1
    for (int c0 = c_begin; c0 < block_n; c0 += 16) {
      uint32_t raw[16];
      tmem_ld16(taddr + c0, raw);
      uint4 ra[2], rb[2], rc[2], rd[2];
      if ((F & (F_RES1 | F_RES2)) && valid) {
        ra[0] = __ldg(reinterpret_cast<const uint4*>(r.r00 + c0));
        ra[1] = __ldg(reinterpret_cast<const uint4*>(r.r00 + c0) + 1);
        if (F & F_RES2) {
          rb[0] = __ldg(reinterpret_cast<const uint4*>(r.r01 + c0));
          rb[1] = __ldg(reinterpret_cast<const uint4*>(r.r01 + c0) + 1);
          rc[0] = __ldg(reinterpret_cast<const uint4*>(r.r10 + c0));
          rc[1] = __ldg(reinterpret_cast<const uint4*>(r.r10 + c0) + 1);
          rd[0] = __ldg(reinterpret_cast<const uint4*>(r.r11 + c0));
          rd[1] = __ldg(reinterpret_cast<const uint4*>(r.r11 + c0) + 1);
        }
      }
      tmem_ld_wait16(raw);
      finish(c0, raw, ra, rb, rc, rd);
    }
  } else if constexpr (kAllowPipe) {
    // no global operands: the TMEM load of chunk k + 1 is in flight while chunk k is scaled, activated and stored (two
    // register buffers; tcgen05.wait::ld covers every outstanding load, so the next one is issued right after the wait)
    uint32_t raw_a[16], raw_b[16];
    const uint4 none[2] = {};
    tmem_ld16(taddr + c_begin, raw_a);
#pragma unroll 1
    for (int c0 = c_begin; c0 < block_n; c0 += 32) {
      tmem_ld_wait16(raw_a);
      const bool second = c0 + 16 < block_n;
      if (second) tmem_ld16(taddr + c0 + 16, raw_b);
      finish(c0, raw_a, none, none, none, none);
      if (second) {
        tmem_ld_wait16(raw_b);
        if (c0 + 32 < block_n) tmem_ld16(taddr + c0 + 32, raw_a);
        finish(c0 + 16, raw_b, none, none, none, none);
      }
    }
  }
  if ((F & F_RGB) && valid) {
    r.rgb[0] = rgb0;
    r.rgb[p.rgb_plane] = rgb1;
    r.rgb[2 * p.rgb_plane] = rgb2;
  }
}

// ------------------------------------------------------------------------------------------ ConvUpLayer, folded
// ConvUpLayer.forward (gfpganv1_ocr_arch.py:188-202) = conv3x3(F.interpolate(x, 2, 'bilinear')): bilinear x2 is linear and
// fixed, so for every output phase (py, px) the pair collapses into ONE 3x3 conv over the LOW-resolution input with
// weights sum_{kh,kw} W[kh,kw] A[py+kh][dy] A[px+kw][dx] (A = the bilinear tap matrix; same MMA work as the conv on the
// up-sampled grid, but no up-sampled tensor, and the four phases are column blocks: N = 4*cout).  The low-resolution
// input is replicate-padded, which reproduces the clamped interpolation everywhere; the only difference to the reference
// is the outermost output ring, where the reference's conv sees ZERO padding of the up-sampled tensor while the folded
// form sees interpolated values.  That surplus is four 1-D convolutions of the border rows / columns (computed by small
// GEMMs of the same kernel, corner terms added back) and is subtracted here before bias and activation.  Verified in
// fp64 against F.conv2d(F.interpolate(...)): max |diff| 1e-14 (tests/test_upfold_cpu.py).
// Per 16-column chunk: phase t = column / ps_c -> output pixel (2y + t/2, 2x + t%2); the ResUpBlock skip
// (gfpganv1_ocr_arch.py:224, 1x1 conv commuted to low resolution) is sampled bilinearly for that pixel.
__device__ __forceinline__ void epilogue_upfold(const ConvParams& p, uint32_t taddr, uint64_t* full_bar,
                                                uint32_t full_phase, int x, int y, int b, int n0, bool valid,
                                                uint32_t s_bias, int c_begin, int c_end) {
  mbar_wait_backoff(full_bar, full_phase, p.epi_wait_ns);
  tc_fence_after();
  if (p.dbg_skip_epi) return;
  const float ag = p.act_gain, slope = p.slope;
  const int OH = 2 * p.m_h, OW = 2 * p.m_w;
  __half* out_b = reinterpret_cast<__half*>(p.out) + (long long)b * p.out_sb + p.out_c_off;
  const __half* res_b = p.res + (long long)b * p.res_sb;
#pragma unroll 1
  for (int c0 = c_begin; c0 < c_end; c0 += 16) {
    uint32_t raw[16];
    tmem_ld16(taddr + c0, raw);
    const int n = n0 + c0;
    const int t = n >> p.ps_shift;
    const int cc = n - (t << p.ps_shift);
    const int yo = 2 * y + (t >> 1), xo = 2 * x + (t & 1);
    float4 bs[4];  // the four phases share one bias vector of ps_c entries: s_bias points at entry n0 % ps_c
#pragma unroll
    for (int j = 0; j < 4; ++j) bs[j] = lds_f4(s_bias + (cc - (n0 & (p.ps_c - 1)) + 4 * j) * 4);
    // bilinear sample of the low-resolution skip at (yo, xo) (align_corners=False, clamped)
    uint4 ra[2], rb[2], rc[2], rd[2];
    float w00 = 0.f, w01 = 0.f, w10 = 0.f, w11 = 0.f;
    float cr[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) cr[j] = 0.f;
    if (valid) {
      const int ky = yo >> 1, kx = xo >> 1;
      int ya, yb, xa, xb;
      float wy0, wy1, wx0, wx1;
      if (yo & 1) { ya = ky; yb = min(ky + 1, p.res_h - 1); wy0 = 0.75f; wy1 = 0.25f; }
      else        { ya = max(ky - 1, 0); yb = ky; wy0 = 0.25f; wy1 = 0.75f; }
      if (xo & 1) { xa = kx; xb = min(kx + 1, p.res_w - 1); wx0 = 0.75f; wx1 = 0.25f; }
      else        { xa = max(kx - 1, 0); xb = kx; wx0 = 0.25f; wx1 = 0.75f; }
      const __half* r00 = res_b + (long long)ya * p.res_sy + (long long)xa * p.res_sx + cc;
      const __half* r01 = res_b + (long long)ya * p.res_sy + (long long)xb * p.res_sx + cc;
      const __half* r10 = res_b + (long long)yb * p.res_sy + (long long)xa * p.res_sx + cc;
      const __half* r11 = res_b + (long long)yb * p.res_sy + (long long)xb * p.res_sx + cc;
      ra[0] = __ldg(reinterpret_cast<const uint4*>(r00)); ra[1] = __ldg(reinterpret_cast<const uint4*>(r00) + 1);
      rb[0] = __ldg(reinterpret_cast<const uint4*>(r01)); rb[1] = __ldg(reinterpret_cast<const uint4*>(r01) + 1);
      rc[0] = __ldg(reinterpret_cast<const uint4*>(r10)); rc[1] = __ldg(reinterpret_cast<const uint4*>(r10) + 1);
      rd[0] = __ldg(reinterpret_cast<const uint4*>(r11)); rd[1] = __ldg(reinterpret_cast<const uint4*>(r11) + 1);
      wy0 *= p.res_mul;
      wy1 *= p.res_mul;
      w00 = wy0 * wx0; w01 = wy0 * wx1; w10 = wy1 * wx0; w11 = wy1 * wx1;
      // surplus of the folded form on the outermost ring
      if (yo == 0 || yo == OH - 1) {
        const float* cp = (yo == 0 ? p.corr_top : p.corr_bot) + ((long long)b * OW + xo) * p.ps_c + cc;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 c4 = __ldg(reinterpret_cast<const float4*>(cp) + j);
          cr[4 * j] += c4.x; cr[4 * j + 1] += c4.y; cr[4 * j + 2] += c4.z; cr[4 * j + 3] += c4.w;
        }
      }
      if (xo == 0 || xo == OW - 1) {
        const float* cp = (xo == 0 ? p.corr_left : p.corr_right) + ((long long)b * OH + yo) * p.ps_c + cc;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 c4 = __ldg(reinterpret_cast<const float4*>(cp) + j);
          cr[4 * j] += c4.x; cr[4 * j + 1] += c4.y; cr[4 * j + 2] += c4.z; cr[4 * j + 3] += c4.w;
        }
      }
    }
    tmem_ld_wait16(raw);
    if (!valid) continue;
    float v[16];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      v[4 * j] = fmaf(__uint_as_float(raw[4 * j]) - cr[4 * j], ag, bs[j].x);
      v[4 * j + 1] = fmaf(__uint_as_float(raw[4 * j + 1]) - cr[4 * j + 1], ag, bs[j].y);
      v[4 * j + 2] = fmaf(__uint_as_float(raw[4 * j + 2]) - cr[4 * j + 2], ag, bs[j].z);
      v[4 * j + 3] = fmaf(__uint_as_float(raw[4 * j + 3]) - cr[4 * j + 3], ag, bs[j].w);
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], slope * v[j]);
    float f[16];
    unpack_half8(ra[0], f); unpack_half8(ra[1], f + 8);
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = fmaf(f[j], w00, v[j]);
    unpack_half8(rb[0], f); unpack_half8(rb[1], f + 8);
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = fmaf(f[j], w01, v[j]);
    unpack_half8(rc[0], f); unpack_half8(rc[1], f + 8);
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = fmaf(f[j], w10, v[j]);
    unpack_half8(rd[0], f); unpack_half8(rd[1], f + 8);
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = fmaf(f[j], w11, v[j]);
    uint32_t pk[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      __half2 h = f2h2_sat(v[2 * j], v[2 * j + 1]);
      pk[j] = *reinterpret_cast<uint32_t*>(&h);
    }
    const __half* op = out_b + (long long)yo * p.out_sy + (long long)xo * p.out_sx + cc;
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(op), "r"(pk[0]), "r"(pk[1]), "r"(pk[2]),
                 "r"(pk[3]), "r"(pk[4]), "r"(pk[5]), "r"(pk[6]), "r"(pk[7])
                 : "memory");
  }
}

// One tile through the epilogue this kernel instantiation was compiled for: EPI = index into epi_profile_flags (flags are
// template constants) or -1 for the run-time generic one.  Every (block_k, EPI) pair is its own kernel, compiled in its
// own translation unit (conv_epi*.cu): with all variants inlined behind a switch the register allocator spilled
// loop-carried state of the tile loops (a local-memory load per tile showed up as the second largest stall).
template <int EPI, bool kAllowPipe = true>
__device__ __forceinline__ void epilogue_one(const ConvParams& p, uint32_t taddr, uint64_t* full_bar, uint32_t full_phase,
                                             int x, int y, int b, int n0, bool valid, float gain, uint32_t s_bias,
                                             uint32_t s_dm, const float* g_dm, uint32_t s_aux, int aux_stride, int c_begin,
                                             int c_end) {
  if constexpr (EPI >= 0 && (epi_profile_flags(EPI < 0 ? 0 : EPI) & F_UPFOLD) != 0) {
    epilogue_upfold(p, taddr, full_bar, full_phase, x, y, b, n0, valid, s_bias, c_begin, c_end);
  } else if constexpr (EPI >= 0) {
    constexpr int F = epi_profile_flags(EPI);
    const FastRow fr = fast_setup<F>(p, x, y, b, n0, valid);
    if constexpr (kAllowPipe)
      epilogue_fast<F, true>(p, taddr, full_bar, full_phase, fr, valid, gain, s_bias, s_dm, s_aux, aux_stride, n0, c_begin, c_end);
    else
      epilogue_fast_lean<F>(p, taddr, full_bar, full_phase, fr, valid, gain, s_bias, s_dm, s_aux, aux_stride, n0);
  } else {
    const EpiRow r = epi_setup(p, x, y, b, n0, valid, gain);
    epilogue_tile(p, taddr, full_bar, full_phase, r, valid, gain, s_bias, p.bias + n0, s_dm, g_dm, s_aux, aux_stride, 0, 16,
                  n0);
  }
}

}  // namespace b200ir

// Device code of the two tcgen05 convolution kernels (included by conv_igemm.cu after ConvParams / epilogue_tile).
//
// Both kernels are persistent and warp-specialised: warp 0 = TMA producer, warp 1 = MMA issuer (+ TMEM owner),
// warps 2..5 = epilogue.  The producer and MMA warps run warp-uniform control flow and let ONE elected lane issue the
// TMA / tcgen05 instructions: a single thread issues dependent instructions only every few cycles, so those loops are
// kept as short as possible (descriptor words precomputed, only 32-bit adds per MMA, no divisions) — measured: with
// ~1000 scalar instructions per output tile the issuing thread, not the tensor pipe or memory, set the pace.

namespace b200ir {

// hi word of a K-major shared-memory descriptor (SBO, version, swizzle layout); lo word = (addr >> 4)
__device__ __forceinline__ uint32_t desc_hi_word(uint32_t row_bytes) {
  return static_cast<uint32_t>(make_kmajor_desc(0, row_bytes) >> 32);
}
__device__ __forceinline__ uint64_t desc64(uint32_t lo, uint32_t hi) {
  return (static_cast<uint64_t>(hi) << 32) | lo;
}

struct KernelSmem {
  uint8_t* base;        // 1024-byte aligned
  uint64_t* full_bar;   // [kMaxStages]
  uint64_t* empty_bar;  // [kMaxStages]
  uint64_t* tmem_full;  // [kMaxAccStages]
  uint64_t* tmem_empty; // [kMaxAccStages]
  uint64_t* w_bar;      // row mode: resident weights landed
  uint32_t* tmem_slot;
  float* bias;          // [kMaxBias] bias of all output channels (zeros when the layer has none)
  float* demod;         // [2][kDemodTable] per-tile demodulation table, double buffered
};

// bytes after the pipeline buffers: barriers + TMEM slot + bias + demod tables
static constexpr int kTailBytes = 512 + kMaxBias * 4 + kMaxEpiGroups * kDemodTable * 4;

// the two epilogue warp groups (4 warps each) drain alternate tiles; each group syncs on its own named barrier
__device__ __forceinline__ void epi_group_sync(int group) {
  asm volatile("bar.sync %0, 128;" ::"r"(group + 1) : "memory");
}

__device__ __forceinline__ KernelSmem carve_smem(uint8_t* smem_raw, uint32_t data_bytes) {
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t pad = ((raw_addr + 1023u) & ~1023u) - raw_addr;
  KernelSmem s;
  s.base = smem_raw + pad;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s.base + data_bytes);
  s.full_bar = bars;
  s.empty_bar = bars + kMaxStages;
  s.tmem_full = bars + 2 * kMaxStages;
  s.tmem_empty = bars + 2 * kMaxStages + kMaxAccStages;
  s.w_bar = bars + 2 * kMaxStages + 2 * kMaxAccStages;
  s.tmem_slot = reinterpret_cast<uint32_t*>(s.w_bar + 1);
  s.bias = reinterpret_cast<float*>(s.base + data_bytes + 512);
  s.demod = s.bias + kMaxBias;
  return s;
}

template <bool kPair = false>
__device__ __forceinline__ uint32_t kernel_prologue(const ConvParams& p, const KernelSmem& s, int nslots, int warp) {
  if (threadIdx.x == 0) {
    for (int i = 0; i < nslots; ++i) {
      mbar_init(&s.full_bar[i], 1);
      mbar_init(&s.empty_bar[i], 1);
    }
    for (int i = 0; i < p.acc_stages; ++i) {
      mbar_init(&s.tmem_full[i], 1);
      // pair: the epilogue threads of both CTAs release the leader's MMA; split epilogue: two groups drain one tile
      mbar_init(&s.tmem_empty[i], (kPair ? 256 : 128) * (p.epi_split ? 2 : 1));
    }
    mbar_init(s.w_bar, 1);
    fence_barrier_init();
  }
  // bias table: all output channels when they fit; all zeros (any kMaxBias-periodic window is valid) without a bias
  if (p.bias_c <= kMaxBias || p.bias == nullptr)
    for (int i = threadIdx.x; i < min(p.bias_c, kMaxBias); i += blockDim.x)
      s.bias[i] = (p.bias != nullptr) ? p.bias[i] * p.act_gain : 0.f;
  if (kPair) cluster_sync_all();  // the peer's barriers exist before any remote arrive, multicast commit or pair TMA load
  if (warp == 1) {
    if (kPair) {
      tmem_alloc_pair(s.tmem_slot, p.tmem_cols);
      tmem_relinquish_pair();
    } else {
      tmem_alloc(s.tmem_slot, p.tmem_cols);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  return *s.tmem_slot;
}

template <bool kPair = false>
__device__ __forceinline__ void kernel_epilogue(const ConvParams& p, uint32_t tmem_base, int warp) {
  tc_fence_before();
  if (kPair) cluster_sync_all();  // neither CTA leaves (or frees TMEM) while the other may still signal it or read its tiles
  else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    if (kPair) tmem_dealloc_pair(tmem_base, p.tmem_cols);
    else tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

// ================================================================================================ generic tiles
// kPair: clusters of two CTAs (the two SMs of a TPC) work on two vertically adjacent M-tiles of the same N-tile with ONE
// tcgen05.mma.cta_group::2 (M = 256) per K step, issued by the leader (cluster rank 0): each CTA stages its own 128 A rows and
// half of the B tile, so the weight traffic per CTA (L2 -> shared memory and shared memory -> tensor core) halves -- the
// operand fetch, not the tensor pipe, paces the N = 128 layers.  Barriers: the leader's full barrier counts
// the bytes of both CTAs' loads, its commits arrive on both CTAs' empty / accumulator-full barriers, and the epilogue threads
// of both CTAs release the leader's accumulator-empty barrier.  Weights are never resident in this variant.
template <int kBlockK, int EPI, bool kPair = false>
__global__ void __launch_bounds__(EpiCfg<EPI>::kThreads, 1) conv_igemm_kernel(const __grid_constant__ ConvParams p) {
  constexpr int kGroups = EpiCfg<EPI>::kGroups;
  extern __shared__ uint8_t smem_raw[];
  constexpr uint32_t row_bytes = kBlockK * 2;
  constexpr uint32_t a_bytes = kBlockM * row_bytes;
  constexpr int k_steps = kBlockK / 16;
  const uint32_t b_bytes = (kPair ? p.block_n / 2 : p.block_n) * row_bytes;
  const uint32_t stage_bytes = (!kPair && p.b_resident) ? a_bytes : a_bytes + b_bytes;
  const uint32_t w_bytes = (!kPair && p.b_resident) ? p.num_taps * p.k_chunks * b_bytes : 0u;  // resident weights sit in front
  const KernelSmem s = carve_smem(smem_raw, w_bytes + p.stages * stage_bytes);
  uint8_t* const ring = s.base + w_bytes;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = kPair ? cluster_ctarank() : 0u;
  // work items: tiles, or (pair of M-tiles, N-tile) items of which this CTA takes the M-tile of its rank
  const int item0 = kPair ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int item_step = kPair ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int num_items = kPair ? p.pair_tiles : p.num_tiles;
  auto tile_of = [&](int item) {
    return kPair ? (2 * (item / p.tiles_n) + (int)rank) * p.tiles_n + item % p.tiles_n : item;
  };
  const uint32_t tmem_base = kernel_prologue<kPair>(p, s, p.stages, warp);
  if (warp == 0) {
    // ---------------- TMA producer
    if (lane == 0) {
      for (int v = 0; v < B200IR_MAX_VIEWS; ++v) tma_prefetch_desc(&p.tmap_a[v]);
      tma_prefetch_desc(&p.tmap_b);
      if (kPair) tma_prefetch_desc(&p.tmap_b2);
      if (!kPair && p.b_resident) {  // one N-tile: every CTA needs the same weights for every tile, load them once
        mbar_arrive_expect_tx(s.w_bar, w_bytes);
        for (int kb = 0; kb < p.num_taps * p.k_chunks; ++kb)
          tma_load_2d(s.base + kb * b_bytes, &p.tmap_b, s.w_bar, kb * kBlockK, 0);
      }
    }
    __syncwarp();
    int stage = 0;
    uint32_t phase = 0;
    for (int item = item0; item < num_items; item += item_step) {
      const int tile = tile_of(item);
      const TileCoord t = decode_tile(p, tile);
      const uint32_t mask = p.tap_mask[(tile % p.tiles_n) & 7];
      int kb = 0;
      for (int tap = 0; tap < p.num_taps; ++tap) {
        if (!((mask >> tap) & 1u)) {
          kb += p.k_chunks;
          continue;
        }
        const int view = p.tap_view[tap];
        const int cx = t.x0 + p.tap_dx[tap];
        const int cy = t.y0 + p.tap_dy[tap];
        for (int kc = 0; kc < p.k_chunks; ++kc, ++kb) {
          mbar_wait_parked(&s.empty_bar[stage], phase ^ 1u);
          if (elect_one()) {
            uint8_t* sa = ring + stage * stage_bytes;
            if (kPair) {
              // the leader's barrier of this stage collects the bytes of both CTAs (its own arrive announces them all)
              if (rank == 0) mbar_arrive_expect_tx(&s.full_bar[stage], 2 * stage_bytes);
              const uint32_t bar = mapa_rank(smem_u32(&s.full_bar[stage]), 0);
              tma_load_4d_pair(sa, &p.tmap_a[view], bar, kc * kBlockK, cx, cy, t.b0);
              tma_load_2d_pair(sa + a_bytes, &p.tmap_b2, bar, kb * kBlockK, t.n0 + (int)rank * (p.block_n / 2));
            } else {
              mbar_arrive_expect_tx(&s.full_bar[stage], stage_bytes);
              tma_load_4d(sa, &p.tmap_a[view], &s.full_bar[stage], kc * kBlockK, cx, cy, t.b0);
              if (!p.b_resident) tma_load_2d(sa + a_bytes, &p.tmap_b, &s.full_bar[stage], kb * kBlockK, t.n0 + t.b0 * p.w_img_rows);
            }
          }
          __syncwarp();
          if (++stage == p.stages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1 && (!kPair || rank == 0)) {
    // ---------------- MMA issuer (pair: the leader CTA only)
    const uint32_t hi = desc_hi_word(row_bytes);
    const uint32_t base_lo = smem_u32(ring) >> 4;
    const uint32_t w_lo = smem_u32(s.base) >> 4;
    const uint32_t stage_lo = stage_bytes >> 4;
    if (!kPair && p.b_resident) {
      mbar_wait(s.w_bar, 0);
      tc_fence_after();
    }
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int item = item0; item < num_items; item += item_step, ++it) {
      const int tile = tile_of(item);
      const int acc = it & (p.acc_stages - 1);
      mbar_wait(&s.tmem_empty[acc], ((it >> p.acc_shift) & 1) ^ 1u);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + acc * p.block_n;
      const int num_kb = __popc(p.tap_mask[(tile % p.tiles_n) & 7] & ((1u << p.num_taps) - 1u)) * p.k_chunks;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&s.full_bar[stage], phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t a_lo = base_lo + stage * stage_lo;
          // resident weights: every tap is executed (no tap mask), so the k-block index is the running kb
          const uint32_t b_lo = (!kPair && p.b_resident) ? w_lo + kb * (b_bytes >> 4) : a_lo + (a_bytes >> 4);
          if (kPair) {
#pragma unroll
            for (int k = 0; k < k_steps; ++k)
              umma_f16_pair(tmem_d, desc64(a_lo + 2 * k, hi), desc64(b_lo + 2 * k, hi), p.idesc_pair, (k > 0 || kb > 0) ? 1u : 0u);
            umma_commit_pair(&s.empty_bar[stage]);
            if (kb == num_kb - 1) umma_commit_pair(&s.tmem_full[acc]);
          } else {
#pragma unroll
            for (int k = 0; k < k_steps; ++k)
              umma_f16(tmem_d, desc64(a_lo + 2 * k, hi), desc64(b_lo + 2 * k, hi), p.idesc, (k > 0 || kb > 0) ? 1u : 0u);
            umma_commit(&s.empty_bar[stage]);
            if (kb == num_kb - 1) umma_commit(&s.tmem_full[acc]);
          }
        }
        __syncwarp();
        if (++stage == p.stages) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  } else if (warp >= 2) {
    // ---------------- epilogue: two groups of 4 warps drain alternate tiles; TMEM lane quarter = warp % 4
    const int q = warp & 3;
    const int group = (warp - 2) >> 2;
    const int gt = (threadIdx.x - 64) & 127;
    const int row = q * 32 + lane;
    const int xx = row % p.tile_w;
    const int yy = (row / p.tile_w) % p.tile_h;
    const int bi = row / (p.tile_w * p.tile_h);
    const float gain = (p.noise != nullptr) ? __ldg(p.noise_gain) * p.act_gain : 0.f;
    int it = group;
    int last_key = -1;
    // a group may wait at most one phase ahead on an accumulator's mbarrier, so no more groups than accumulator stages
    // take part (block_n = 256 has two stages: the third group idles there, those layers are main-loop bound anyway)
    const bool split = kGroups == 3 && p.epi_split;
    const int ngroups = split ? 3 : min(kGroups, p.acc_stages);
    // unit u of this group: a tile (u = it), or half a tile when the epilogue is split (u = 2 * it + half)
    const int my_items = (num_items > item0) ? (num_items - item0 + item_step - 1) / item_step : 0;
    const int units = split ? 2 * my_items : my_items;
    for (int u = (group < ngroups) ? group : units; u < units; u += ngroups) {
      it = split ? (u >> 1) : u;
      const int c_begin = split ? (u & 1) * (p.block_n >> 1) : 0;
      const int c_end = split ? c_begin + (p.block_n >> 1) : p.block_n;
      const int item = item0 + it * item_step;
      const int tile = tile_of(item);
      const int acc = it & (p.acc_stages - 1);
      const TileCoord t = decode_tile(p, tile);
      const int x = t.x0 + xx, y = t.y0 + yy, b = t.b0 + bi;
      const bool valid = (x < p.m_w) && (y < p.m_h) && (b < p.m_b);
      uint32_t s_dm = 0, s_aux = 0;
      const float* g_dm = nullptr;
      const int tab_n = p.tile_b * p.block_n;
      if ((p.demod != nullptr && p.smem_demod) || p.smem_aux) {
        float* tab = s.demod + group * kDemodTable;
        // the tables depend on (first image, N-tile) only: consecutive tiles of a group usually share both, so the
        // staging (global loads + two group barriers, ~10k cycles when done per tile) is skipped when the key repeats
        const int key = t.b0 * p.tiles_n + (tile % p.tiles_n);
        if (key != last_key) {
        last_key = key;
        epi_group_sync(group);  // previous tile of this group fully drained
        // thread gt owns columns gt, gt + 128 of every image row of the table (block_n <= 256): channel indices are
        // computed once, the loads of the tile_b images are independent
        for (int col = gt; col < p.block_n; col += 128) {
          const int n = t.n0 + col;
          const int nd = n % p.demod_c;
#pragma unroll 4
          for (int j = 0; j < p.tile_b; ++j) {
            const int bb = t.b0 + j;
            const bool in = bb < p.m_b;
            const int i = j * p.block_n + col;
            if (p.smem_demod) tab[i] = in ? __ldg(p.demod + (long long)bb * p.demod_c + nd) * p.act_gain : 0.f;
            if (p.smem_aux) {
              tab[tab_n + i] = (in && p.out_scale != nullptr) ? __ldg(p.out_scale + (long long)bb * p.cout + n) : 1.f;
#pragma unroll
              for (int o = 0; o < 3; ++o)
                tab[(2 + o) * tab_n + i] =
                    (in && p.rgb_w != nullptr) ? __ldg(p.rgb_w + ((long long)bb * 3 + o) * p.cout + n) : 0.f;
            }
          }
        }
        epi_group_sync(group);
        }
        if (p.smem_demod) s_dm = smem_u32(tab + bi * p.block_n);
        if (p.smem_aux) s_aux = smem_u32(tab + tab_n + bi * p.block_n);
      }
      if (p.demod != nullptr && !p.smem_demod) g_dm = p.demod + (long long)b * p.demod_c + t.n0 % p.demod_c;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * p.block_n;
      // wide layers (cout > kMaxBias): bias from global memory (act == 0), or the zero table when there is none
      const uint32_t s_bias = (p.bias_c <= kMaxBias) ? smem_u32(s.bias + t.n0 % p.bias_c)
                                                     : (p.bias == nullptr ? smem_u32(s.bias + t.n0 % kMaxBias) : 0u);
      epilogue_one<EPI>(p, taddr, &s.tmem_full[acc], (it >> p.acc_shift) & 1, x, y, b, t.n0, valid, gain, s_bias, s_dm,
                        g_dm, s_aux, tab_n, c_begin, c_end);
      tc_fence_before();
      if (kPair && rank != 0) mbar_arrive_cluster(mapa_rank(smem_u32(&s.tmem_empty[acc]), 0));
      else mbar_arrive(&s.tmem_empty[acc]);
    }
  }
  kernel_epilogue<kPair>(p, tmem_base, warp);
}

// ================================================================================================ row mode
// 3x3 stride-1 conv for low channel counts at high resolution.  A work item is (image b, 128-pixel column segment,
// chunk of R output rows).  Every input row segment (130 pixels: 1-pixel halo each side, TMA zero fill outside the
// image) is loaded ONCE into a ring slot and consumed ONCE: input row i of an item contributes to the output rows
// i-2 (kh = 2), i-1 (kh = 1) and i (kh = 0), whose accumulators sit side by side in the TMEM ring, so one MMA with
// N = 3*cout and the B tile [W(kh=2) | W(kh=1) | W(kh=0)] updates all three.  Per input row that is 3 (kw) x Cin/16
// MMAs of N = 3*cout instead of 9 x Cin/16 MMAs of N = cout: the A operand (128 x 16 fp16 = 4 KB per MMA) is read from
// shared memory 3x less often, and with N <= 64 that read, not the tensor pipe, set the MMA rate (measured: 70 cycles
// per N=64 MMA against a 32-cycle pipe floor).  The three kw taps read the row through descriptors whose start address
// is shifted by kw pixels (measured on B200: the swizzle XOR uses absolute shared-memory address bits, so row-shifted
// start addresses work with the matrix-base-offset field left at 0 -- tools/try_row_mode.py).
// The newest target (output row i) has no partial sum yet: the first MMA of an input row is split so that this slice
// runs with accumulate = 0; a target run that wraps around the TMEM ring is split in two as well.
// All 9 tap weight tiles stay resident in shared memory for the whole kernel, grouped per (kw, k-chunk) in kh order 2,1,0.
struct RowItem {
  int b, seg, y0, rows_out;
};
__device__ __forceinline__ RowItem decode_item(const ConvParams& p, int item) {
  RowItem r;
  const int chunk = item % p.row_chunks;
  const int rest = item / p.row_chunks;
  r.seg = rest % p.tiles_w;
  r.b = rest / p.tiles_w;
  r.y0 = chunk * p.row_R;
  r.rows_out = min(p.row_R, p.m_h - r.y0);
  return r;
}

// Work items of one CTA.  Classic: items blockIdx.x, blockIdx.x + gridDim.x, ... of the uniform (image, segment, row chunk)
// grid.  Balanced (p.row_balanced): the output rows of all (image, segment) pairs form one list, CTA k takes the k-th of
// gridDim.x equal ranges and walks it pair by pair -- when the uniform grid does not divide by the CTA count (64 x 192
// levels at batch 64: 128 items on 148 SMs) the longest CTA gets ~13 % fewer rows.
struct RowIter {
  int item, step, last;       // classic
  long long r, r_end;         // balanced: next global output row, end of this CTA's range
  __device__ __forceinline__ RowIter(const ConvParams& p) {
    item = blockIdx.x; step = gridDim.x; last = p.row_items;
    const long long total = (long long)p.m_b * p.tiles_w * p.m_h;
    r = total * blockIdx.x / gridDim.x;
    r_end = total * (blockIdx.x + 1) / gridDim.x;
  }
  __device__ __forceinline__ bool next(const ConvParams& p, RowItem& w) {
    if (!p.row_balanced) {
      if (item >= last) return false;
      w = decode_item(p, item);
      item += step;
      return true;
    }
    if (r >= r_end) return false;
    const int pair = (int)(r / p.m_h);
    w.y0 = (int)(r - (long long)pair * p.m_h);
    const long long pair_end = (long long)(pair + 1) * p.m_h;
    w.rows_out = (int)((pair_end < r_end ? pair_end : r_end) - r);
    w.seg = pair % p.tiles_w;
    w.b = pair / p.tiles_w;
    r += w.rows_out;
    return true;
  }
};

// The instantiations the host may run as two CTAs per SM (build_conv_launch: `dual`) are held to the register budget of
// two resident CTAs by their launch bounds (66 / 76 registers without it after the barrier peek: the residual profile no
// longer fitted twice).
template <int kBlockK, int EPI>
__global__ void __launch_bounds__(EpiCfg<EPI>::kRowThreads, (kBlockK <= 32 && (EPI == 0 || EPI == 1)) ? 2 : 1)
conv_row_kernel(const __grid_constant__ ConvParams p) {
  constexpr int kGroups = EpiCfg<EPI>::kRowGroups;
  extern __shared__ uint8_t smem_raw[];
  constexpr uint32_t row_bytes = kBlockK * 2;
  constexpr int k_steps = kBlockK / 16;
  constexpr uint32_t slot_bytes = 136 * row_bytes;
  const uint32_t wtile_bytes = p.block_n * row_bytes;  // one (tap, kc) weight tile
  const KernelSmem s = carve_smem(smem_raw, p.row_w_bytes + p.row_slots * slot_bytes);
  uint8_t* s_w = s.base;
  uint8_t* s_ring = s.base + p.row_w_bytes;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int kc_n = p.k_chunks;
  const int nslots = p.row_slots;
  const uint32_t tmem_base = kernel_prologue(p, s, nslots, warp);

  if (warp == 0) {
    // ---------------- TMA producer: resident weights once, then one box per (input row, kc)
    if (lane == 0) {
      tma_prefetch_desc(&p.tmap_a[0]);
      tma_prefetch_desc(&p.tmap_b);
      mbar_arrive_expect_tx(s.w_bar, 9 * kc_n * wtile_bytes);
      const int cin = kc_n * kBlockK;
      for (int kw = 0; kw < 3; ++kw)
        for (int kc = 0; kc < kc_n; ++kc)
          for (int khi = 0; khi < 3; ++khi)  // kh = 2 - khi: ascending output row inside the merged B tile
            tma_load_2d(s_w + ((kw * kc_n + kc) * 3 + khi) * wtile_bytes, &p.tmap_b, s.w_bar,
                        ((2 - khi) * 3 + kw) * cin + kc * kBlockK, 0);
    }
    __syncwarp();
    int slot = 0;
    uint32_t phase = 0;
    RowItem w;
    for (RowIter iter(p); iter.next(p, w);) {
      const int cx = w.seg * 128 - 1;
      for (int r = 0; r < w.rows_out + 2; ++r) {
        for (int kc = 0; kc < kc_n; ++kc) {
          mbar_wait_parked(&s.empty_bar[slot], phase ^ 1u);
          if (elect_one()) {
            mbar_arrive_expect_tx(&s.full_bar[slot], 130 * row_bytes);
            tma_load_4d(s_ring + slot * slot_bytes, &p.tmap_a[0], &s.full_bar[slot], kc * kBlockK, cx, w.y0 - 1 + r,
                        w.b);
          }
          __syncwarp();
          if (++slot == nslots) {
            slot = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer.  Warp-uniform control flow, one elected lane issues (the compiler recognises the
    // elect.sync pattern and emits plain UTCHMMA; a lane == 0 branch makes it wrap every MMA in an election loop).
    // The issuing thread executes dependent scalar instructions only every ~5 cycles and nothing else bounds the
    // low-channel layers (measured on 32 -> 32 channels: tensor pipe 20 % busy, no wait on data or accumulators), so
    // the common row -- three targets, the newest one fresh, no wrap of the accumulator ring -- is straight-line code.
    mbar_wait(s.w_bar, 0);
    tc_fence_after();
    const uint32_t hi = desc_hi_word(row_bytes);
    const uint32_t w_lo = smem_u32(s_w) >> 4;
    const uint32_t ring_lo = smem_u32(s_ring) >> 4;
    const uint32_t wtile_lo = wtile_bytes >> 4;
    constexpr uint32_t slot_lo = slot_bytes >> 4;
    constexpr uint32_t px_lo = row_bytes >> 4;  // one pixel (one smem row) in descriptor units
    const int ring = p.acc_stages;               // accumulator slots (power of two)
    const uint32_t id1 = p.idesc_n[0], id2 = p.idesc_n[1], id3 = p.idesc_n[2];
    const uint32_t bn = p.block_n;
    int slot = 0;                                // ring slot of (current input row, kc 0)
    uint32_t phase = 0;
    int it_base = 0;                             // output rows issued before this item
    // The scalar work between the last MMA of one input row and the first MMA of the next (two barrier polls, slot
    // arithmetic, election) is not hidden by the tensor pipe once its short queue has drained, and the issuing thread is
    // what bounds these layers.  So the common row peeks at the NEXT row's barriers (non-blocking test_wait) in the middle
    // of its own MMA sequence, while the kw = 0 MMAs are still executing; the next iteration then skips the polls.
    bool pf_full = false, pf_empty = false;
    const int peek = p.row_peek;
    RowItem w;
    for (RowIter iter(p); iter.next(p, w);) {
      for (int i = 0; i < w.rows_out + 2; ++i) {
        const int row_slot = slot;
        for (int kc = 0; kc < kc_n; ++kc) {  // this input row has landed (all its k-chunks)
          if (!pf_full) mbar_wait(&s.full_bar[slot], phase);
          pf_full = false;
          if (++slot == nslots) {
            slot = 0;
            phase ^= 1u;
          }
        }
        const bool fresh = i < w.rows_out;  // output row i receives its first contribution (kh = 0) from this input row
        if (fresh && !pf_empty) {
          const int it = it_base + i;
          mbar_wait(&s.tmem_empty[it & (ring - 1)], ((it >> p.acc_shift) & 1) ^ 1u);
        }
        pf_empty = false;
        tc_fence_after();
        const int sa = (it_base + i - 2) & (ring - 1);
        if (kc_n == 1 && i >= 2 && fresh && sa <= ring - 3) {
          const uint32_t d = tmem_base + sa * bn;
          const uint32_t a_lo = ring_lo + row_slot * slot_lo;
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < k_steps; ++k) {
              const uint64_t a_desc = desc64(a_lo + 2 * k, hi);
              const uint32_t b_lo = w_lo + 2 * k;
              if (k == 0) {
                umma_f16_fixed<true>(d, a_desc, desc64(b_lo, hi), id2);
                umma_f16_fixed<false>(d + 2 * bn, a_desc, desc64(b_lo + 2 * wtile_lo, hi), id1);
              } else {
                umma_f16_fixed<true>(d, a_desc, desc64(b_lo, hi), id3);
              }
            }
          }
          __syncwarp();
          if (peek) {  // slot / phase already point at the next input row
            pf_full = mbar_test_wait(&s.full_bar[slot], phase);
            if (i + 1 < w.rows_out) {
              const int it1 = it_base + i + 1;
              pf_empty = mbar_test_wait(&s.tmem_empty[it1 & (ring - 1)], ((it1 >> p.acc_shift) & 1) ^ 1u);
            }
          }
          if (elect_one()) {
#pragma unroll
            for (int kw = 1; kw < 3; ++kw) {
#pragma unroll
              for (int k = 0; k < k_steps; ++k) {
                const uint64_t a_desc = desc64(a_lo + kw * px_lo + 2 * k, hi);
                const uint32_t b_lo = w_lo + kw * 3 * wtile_lo + 2 * k;
                umma_f16_fixed<true>(d, a_desc, desc64(b_lo, hi), id3);
              }
            }
            umma_commit(&s.empty_bar[row_slot]);
            umma_commit(&s.tmem_full[sa]);
          }
          __syncwarp();
          continue;
        }
        if (elect_one()) {
          const int j_lo = max(i - 2, 0), j_hi = min(i, w.rows_out - 1);
          // one accumulating run over output rows [ja, jb] (split where the accumulator ring wraps)
          auto run = [&](int ja, int jb, uint64_t a_desc, uint32_t tile_lo, uint32_t accumulate) {
            const int n = jb - ja + 1;
            const int s0 = (it_base + ja) & (ring - 1);
            const int n1 = min(n, ring - s0);
            const uint32_t b_lo = tile_lo + (ja - (i - 2)) * wtile_lo;
            umma_f16(tmem_base + s0 * bn, a_desc, desc64(b_lo, hi), p.idesc_n[n1 - 1], accumulate);
            if (n1 < n) umma_f16(tmem_base, a_desc, desc64(b_lo + n1 * wtile_lo, hi), p.idesc_n[n - n1 - 1], accumulate);
          };
          int sl = row_slot;
          for (int kc = 0; kc < kc_n; ++kc) {
            const uint32_t a_lo = ring_lo + sl * slot_lo;
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
              const uint32_t tile_lo = w_lo + (kw * kc_n + kc) * 3 * wtile_lo;
#pragma unroll
              for (int k = 0; k < k_steps; ++k) {
                const uint64_t a_desc = desc64(a_lo + kw * px_lo + 2 * k, hi);
                if (fresh && kc == 0 && kw == 0 && k == 0) {
                  if (j_hi > j_lo) run(j_lo, j_hi - 1, a_desc, tile_lo + 2 * k, 1u);
                  run(j_hi, j_hi, a_desc, tile_lo + 2 * k, 0u);
                } else {
                  run(j_lo, j_hi, a_desc, tile_lo + 2 * k, 1u);
                }
              }
            }
            umma_commit(&s.empty_bar[sl]);  // this (row, kc) slot is consumed once
            if (++sl == nslots) sl = 0;
          }
          if (i >= 2) umma_commit(&s.tmem_full[(it_base + i - 2) & (ring - 1)]);  // output row i-2 is complete
        }
        __syncwarp();
      }
      it_base += w.rows_out;
    }
  } else {
    // ---------------- epilogue: two groups of 4 warps drain alternate output rows
    const int q = warp & 3;
    const int group = (warp - 2) >> 2;
    const int gt = (threadIdx.x - 64) & 127;
    const int row = q * 32 + lane;
    const float gain = (p.noise != nullptr) ? __ldg(p.noise_gain) * p.act_gain : 0.f;
    const int ngroups = min(kGroups, p.acc_stages);  // see conv_igemm_kernel (the ring has 8 or 16 stages here)
    int it = 0;
    int mine = (group < ngroups) ? group : 0x7fffffff;  // next output row this group drains (rows go round the groups;
                                                        // a running counter: `it % ngroups` was a run-time division
                                                        // per row and warp, ~12 % of the kernel's instructions)
    RowItem w;
    for (RowIter iter(p); iter.next(p, w);) {
      const int x = w.seg * 128 + row;
      const bool valid = x < p.m_w;
      uint32_t s_dm = 0, s_aux = 0;
      if (p.demod != nullptr || p.smem_aux) {  // one image per item: each group stages that image's table rows once
        float* tab = s.demod + group * kRowTable;
        epi_group_sync(group);
        for (int i = gt; i < p.block_n; i += 128) {
          const long long ch = (long long)w.b * p.cout + i;
          if (p.demod != nullptr) tab[i] = __ldg(p.demod + ch) * p.act_gain;
          if (p.smem_aux) {
            tab[p.block_n + i] = (p.out_scale != nullptr) ? __ldg(p.out_scale + ch) : 1.f;
#pragma unroll
            for (int o = 0; o < 3; ++o)
              tab[(2 + o) * p.block_n + i] =
                  (p.rgb_w != nullptr) ? __ldg(p.rgb_w + ((long long)w.b * 3 + o) * p.cout + i) : 0.f;
          }
        }
        epi_group_sync(group);
        if (p.demod != nullptr) s_dm = smem_u32(tab);
        if (p.smem_aux) s_aux = smem_u32(tab + p.block_n);
      }
      for (int j = 0; j < w.rows_out; ++j, ++it) {
        if (it != mine) continue;
        mine += ngroups;
        const int acc = it & (p.acc_stages - 1);
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * p.block_n;
        // 32-channel inputs with the plain profiles may run two CTAs per SM (build_conv_launch: `dual`): keep them lean
        epilogue_one<EPI, !(kBlockK <= 32 && (EPI == 0 || EPI == 1))>(p, taddr, &s.tmem_full[acc], (it >> p.acc_shift) & 1, x,
                                                                     w.y0 + j, w.b, 0, valid, gain, smem_u32(s.bias), s_dm,
                                                                     nullptr, s_aux, p.block_n, 0, p.block_n);
        tc_fence_before();
        mbar_arrive(&s.tmem_empty[acc]);
      }
    }
  }
  kernel_epilogue(p, tmem_base, warp);
}


// ------------------------------------------------------------------------------------------ launch (per EPI)
// Defined as a template here, explicitly instantiated once per EPI in conv_epi*.cu; conv_igemm.cu calls through
// launch_conv_variant's extern declarations.
template <int EPI>
static int launch_pair(const ConvParams& p, int grid, int smem_bytes, int smem_max, cudaStream_t st) {
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(conv_igemm_kernel<64, EPI, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max);
    if (e != cudaSuccess) {
      set_error("conv: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return 1;
    }
    configured = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3((unsigned)EpiCfg<EPI>::kThreads);
  cfg.dynamicSmemBytes = (size_t)smem_bytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  const cudaError_t e = cudaLaunchKernelEx(&cfg, conv_igemm_kernel<64, EPI, true>, p);
  if (e != cudaSuccess) {
    set_error("conv_igemm (CTA pairs): %s", cudaGetErrorString(e));
    return 1;
  }
  return check_launch("conv_igemm(pair)");
}

template <int kBlockK, int EPI, bool ROW>
static int launch_one(const ConvParams& p, int grid, int smem_bytes, int smem_max, cudaStream_t st) {
  static bool configured = false;
  if (ROW) {
    if (!configured) {
      cudaError_t e = cudaFuncSetAttribute(conv_row_kernel<kBlockK, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max);
      // the two-CTA-per-SM configuration needs the full shared-memory carve-out, not the smallest one that fits one CTA
      if (e == cudaSuccess)
        e = cudaFuncSetAttribute(conv_row_kernel<kBlockK, EPI>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                 cudaSharedmemCarveoutMaxShared);
      if (e != cudaSuccess) {
        set_error("conv: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        return 1;
      }
      configured = true;
    }
    conv_row_kernel<kBlockK, EPI><<<grid, EpiCfg<EPI>::kRowThreads, smem_bytes, st>>>(p);
    return check_launch("conv_row");
  }
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(conv_igemm_kernel<kBlockK, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max);
    if (e != cudaSuccess) {
      set_error("conv: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return 1;
    }
    configured = true;
  }
  conv_igemm_kernel<kBlockK, EPI><<<grid, EpiCfg<EPI>::kThreads, smem_bytes, st>>>(p);
  return check_launch("conv_igemm");
}

// block_k 16 exists only with the generic epilogue (narrow test nets); the host never selects EPI >= 0 for it
template <int EPI>
int launch_conv_variant(const ConvParams& p, int block_k, bool row, int grid, int smem_bytes, int smem_max,
                        cudaStream_t st) {
  if (p.pair && !row && block_k == 64) return launch_pair<EPI>(p, grid, smem_bytes, smem_max, st);
  if (row) {
    if (block_k == 64) return launch_one<64, EPI, true>(p, grid, smem_bytes, smem_max, st);
    if (block_k == 32) return launch_one<32, EPI, true>(p, grid, smem_bytes, smem_max, st);
    if constexpr (EPI < 0) return launch_one<16, EPI, true>(p, grid, smem_bytes, smem_max, st);
  } else {
    if (block_k == 64) return launch_one<64, EPI, false>(p, grid, smem_bytes, smem_max, st);
    if (block_k == 32) return launch_one<32, EPI, false>(p, grid, smem_bytes, smem_max, st);
    if constexpr (EPI < 0) return launch_one<16, EPI, false>(p, grid, smem_bytes, smem_max, st);
  }
  set_error("conv: no kernel for block_k=%d with epilogue profile %d", block_k, EPI);
  return 1;
}

}  // namespace b200ir
