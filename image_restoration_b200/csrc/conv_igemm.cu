// Implicit-GEMM convolution for sm_100a: TMA box loads (im2col by shifted boxes with out-of-bounds zero fill)
// -> 128B/64B-swizzled shared memory -> tcgen05.mma (M=128, N=block_n, K=16, fp16 x fp16 -> fp32 in TMEM)
// -> tcgen05.ld epilogue with bias / demodulation / noise / leaky-ReLU / residual fused.
//
// Persistent, warp-specialised: warp 0 = TMA producer, warp 1 = MMA issuer (+ TMEM owner), warps 2..5 =
// epilogue (one TMEM lane quarter each).  Two TMEM accumulator stages let the epilogue of tile i overlap the
// main loop of tile i+1.
//
// Reference semantics implemented here: see include/b200ir.h (b200ir_conv_igemm).
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "host_common.h"
#include "ptx.cuh"

namespace b200ir {

static constexpr int kBlockM = 128;
static constexpr int kEpiWarps = 8;                         // two per TMEM lane quarter
static constexpr int kEpiThreads = kEpiWarps * 32;
static constexpr int kThreads = 64 + kEpiThreads;       // warp 0 producer, warp 1 MMA, warps 2.. epilogue
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
  uint32_t idesc;
  uint32_t idesc_n[3];  // row mode: instruction descriptors for N = 1, 2, 3 x block_n
  uint32_t tmem_cols;
  int acc_stages, acc_shift;
  // row mode (conv_row_kernel): 3x3 stride-1 conv, tile = 128 consecutive pixels of one row, weights resident in
  // shared memory, each input row segment loaded once (with a 1-pixel halo) and reused for 3 kw shifts x 3 output rows
  int row_R, row_chunks, row_items, row_slots, row_slot_bytes, row_w_bytes, desc_mode;
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
  const float* out_scale;
  const float* rgb_w;
  float* rgb_part;
  long long rgb_plane, rgb_image;  // rgb_h*rgb_w_px, m_b*3*rgb_plane
  int rgb_w_px;
  int no_store;
  int dbg_skip_epi;  // profiling aid (env B200IR_DBG_SKIP_EPI): epilogue only recycles the accumulators
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
  const int xo = x * p.out_x_mul + p.out_x_off;
  const int yo = y * p.out_y_mul + p.out_y_off;
  r.out_off = (long long)b * p.out_sb + (long long)yo * p.out_sy + (long long)xo * p.out_sx + p.out_c_off + n0;
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
                                              uint32_t s_aux, int aux_stride, int c_begin, int c_step) {
  mbar_wait(full_bar, full_phase);
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
        for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], 0.2f * v[j]);
      }
      if (p.res_mode == 1) {
        float f[16];
        unpack_half8(ra[0], f);
        unpack_half8(ra[1], f + 8);
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = (v[j] + f[j]) * p.res_scale;
      } else if (p.res_mode == 2) {
        float fa[16], fb[16], fc[16], fd[16];
        unpack_half8(ra[0], fa); unpack_half8(ra[1], fa + 8);
        unpack_half8(rb[0], fb); unpack_half8(rb[1], fb + 8);
        unpack_half8(rc[0], fc); unpack_half8(rc[1], fc + 8);
        unpack_half8(rd[0], fd); unpack_half8(rd[1], fd + 8);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float up = r.wy0 * (r.wx0 * fa[j] + r.wx1 * fb[j]) + r.wy1 * (r.wx0 * fc[j] + r.wx1 * fd[j]);
          v[j] = (v[j] + up) * p.res_scale;
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
        float4* op = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + r.out_off + c0);
#pragma unroll
        for (int j = 0; j < 4; ++j) op[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
      } else {
        uint32_t pk[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          __half2 h = __floats2half2_rn(v[2 * j], v[2 * j + 1]);
          pk[j] = *reinterpret_cast<uint32_t*>(&h);
        }
        __half* op = reinterpret_cast<__half*>(p.out) + r.out_off + c0;
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
enum : int { F_DEMOD = 1, F_NOISE = 2, F_RES1 = 4, F_RES2 = 8, F_RGB = 16, F_NOSTORE = 32 };
static constexpr int kNumEpiProfiles = 6;
__host__ __device__ constexpr int epi_profile_flags(int i) {
  return i == 0   ? 0
         : i == 1 ? F_RES1
         : i == 2 ? F_RES2
         : i == 3 ? F_DEMOD
         : i == 4 ? (F_DEMOD | F_NOISE | F_RGB)
                  : (F_DEMOD | F_NOISE | F_RGB | F_NOSTORE);
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
  const int xo = x * p.out_x_mul + p.out_x_off;
  const int yo = y * p.out_y_mul + p.out_y_off;
  r.out = reinterpret_cast<__half*>(p.out) + (long long)b * p.out_sb + (long long)yo * p.out_sy +
          (long long)xo * p.out_sx + p.out_c_off + n0;
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
    r.w00 = p.res_scale;
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
    wy0 *= p.res_scale;
    wy1 *= p.res_scale;
    r.w00 = wy0 * wx0; r.w01 = wy0 * wx1; r.w10 = wy1 * wx0; r.w11 = wy1 * wx1;
  }
  return r;
}

// fp16 NHWC output with 32-byte aligned rows, bias (and demod / aux tables) in shared memory, gains pre-folded:
//   v = acc * (demod*g | g) + (bias*g + noise*gain*g);  v = max(v, slope*v);  v += res * w (w carries the residual scale)
template <int F>
__device__ __forceinline__ void epilogue_fast(const ConvParams& p, uint32_t taddr, uint64_t* full_bar,
                                              uint32_t full_phase, const FastRow& r, bool valid, float gain,
                                              uint32_t s_bias, uint32_t s_demod, uint32_t s_aux, int aux_stride) {
  mbar_wait(full_bar, full_phase);
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
        __half2 h = __floats2half2_rn(v[2 * j], v[2 * j + 1]);
        pk[j] = *reinterpret_cast<uint32_t*>(&h);
      }
      asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(r.out + c0), "r"(pk[0]), "r"(pk[1]),
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

// One tile of either kernel through the epilogue the host selected (uniform switch, once per tile).
#define B200IR_EPI_CASE(I)                                                                                        \
  case I: {                                                                                                       \
    constexpr int F = epi_profile_flags(I);                                                                       \
    const FastRow fr = fast_setup<F>(p, x, y, b, n0, valid);                                                      \
    epilogue_fast<F>(p, taddr, full_bar, full_phase, fr, valid, gain, s_bias, s_dm, s_aux, aux_stride);           \
  } break;

__device__ __forceinline__ void epilogue_dispatch(const ConvParams& p, uint32_t taddr, uint64_t* full_bar,
                                                  uint32_t full_phase, int x, int y, int b, int n0, bool valid,
                                                  float gain, uint32_t s_bias, uint32_t s_dm, const float* g_dm,
                                                  uint32_t s_aux, int aux_stride) {
  switch (p.epi) {
    B200IR_EPI_CASE(0)
    B200IR_EPI_CASE(1)
    B200IR_EPI_CASE(2)
    B200IR_EPI_CASE(3)
    B200IR_EPI_CASE(4)
    B200IR_EPI_CASE(5)
    default: {
      const EpiRow r = epi_setup(p, x, y, b, n0, valid, gain);
      epilogue_tile(p, taddr, full_bar, full_phase, r, valid, gain, s_bias, p.bias + n0, s_dm, g_dm, s_aux, aux_stride,
                    0, 16);
    } break;
  }
}
#undef B200IR_EPI_CASE

}  // namespace b200ir

#include "conv_kernels.cuh"

namespace b200ir {

// ------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* sym = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || sym == nullptr) {
    set_error("cuTensorMapEncodeTiled unavailable: %s", cudaGetErrorString(e));
    return nullptr;
  }
  fn = reinterpret_cast<EncodeTiledFn>(sym);
  return fn;
}

int encode_map(CUtensorMap* m, const void* ptr, int rank, const cuuint64_t* dims, const cuuint64_t* strides_b,
               const cuuint32_t* box, CUtensorMapSwizzle swz, const char* what) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return 1;
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, rank, const_cast<void*>(ptr), dims, strides_b, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(%s) failed with CUresult %d (dims %llu,%llu,%llu,%llu box %u,%u,%u,%u)", what,
              (int)r, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
              (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 3 ? dims[3] : 0), box[0],
              rank > 1 ? box[1] : 0, rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0);
    return 1;
  }
  return 0;
}

static int g_num_sms = 0;
static int g_smem_optin = 0;

static int init_device_info() {
  if (g_num_sms) return 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    set_error("no CUDA device");
    return 1;
  }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) {
    set_error("cudaGetDeviceProperties failed");
    return 1;
  }
  if (prop.major != 10) {
    set_error("device is sm_%d%d; libb200ir needs sm_100 (B200) and has no fallback", prop.major, prop.minor);
    return 1;
  }
  g_num_sms = prop.multiProcessorCount;
  g_smem_optin = (int)prop.sharedMemPerBlockOptin;
  return 0;
}

int device_check_impl() { return init_device_info(); }
int num_sms() { return init_device_info() ? 0 : g_num_sms; }
int smem_optin() { return init_device_info() ? 0 : g_smem_optin; }

}  // namespace b200ir

using namespace b200ir;

template <typename K>
static int configure_smem(K kernel, int slot) {
  static bool done[8] = {false, false, false, false, false, false, false, false};
  if (done[slot]) return 0;
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, g_smem_optin);
  if (e != cudaSuccess) {
    set_error("conv: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    return 1;
  }
  done[slot] = true;
  return 0;
}

extern "C" int b200ir_conv_igemm(const b200ir_conv_desc* d, void* stream) {
  B200IR_REQUIRE(d != nullptr, "conv_igemm: null desc");
  if (init_device_info()) return 1;
  B200IR_REQUIRE(d->tile_w > 0 && d->tile_h > 0 && d->tile_b > 0 && d->tile_w * d->tile_h * d->tile_b == kBlockM,
                 "conv_igemm: tile %dx%dx%d must cover exactly 128 positions", d->tile_b, d->tile_h, d->tile_w);
  B200IR_REQUIRE(d->tile_w <= 256 && d->tile_h <= 256 && d->tile_b <= 256, "conv_igemm: tile extent > 256");
  B200IR_REQUIRE(d->cin >= 16 && d->cin % 16 == 0, "conv_igemm: cin=%d must be a multiple of 16", d->cin);
  B200IR_REQUIRE(d->block_n >= 16 && d->block_n <= 256 && d->block_n % 16 == 0 && d->cout % d->block_n == 0,
                 "conv_igemm: block_n=%d invalid for cout=%d", d->block_n, d->cout);
  B200IR_REQUIRE(d->num_taps >= 1 && d->num_taps <= B200IR_MAX_TAPS, "conv_igemm: num_taps=%d", d->num_taps);
  B200IR_REQUIRE(d->num_views >= 1 && d->num_views <= B200IR_MAX_VIEWS, "conv_igemm: num_views=%d", d->num_views);
  B200IR_REQUIRE((d->out != nullptr || d->no_store) && d->weight != nullptr, "conv_igemm: null out/weight");
  B200IR_REQUIRE(d->m_w > 0 && d->m_h > 0 && d->m_b > 0, "conv_igemm: empty M extents");
  B200IR_REQUIRE((d->out_c_off % 8) == 0 && (d->out_stride_x % 8) == 0, "conv_igemm: output not 16B aligned");
  B200IR_REQUIRE(d->res_mode >= 0 && d->res_mode <= 2, "conv_igemm: res_mode");
  B200IR_REQUIRE(d->res_mode == 0 || d->res != nullptr, "conv_igemm: res_mode set but res is NULL");
  B200IR_REQUIRE(d->noise == nullptr || d->noise_gain != nullptr, "conv_igemm: noise without noise_gain");
  B200IR_REQUIRE(d->rgb_w == nullptr || (d->rgb_part != nullptr && d->rgb_w_px > 0 && d->rgb_h > 0),
                 "conv_igemm: rgb_w needs rgb_part and the plane extents");
  B200IR_REQUIRE(!d->no_store || d->rgb_w != nullptr, "conv_igemm: no_store without a fused ToRGB leaves no output");

  ConvParams p;
  memset(&p, 0, sizeof(p));
  p.block_k = (d->cin % 64 == 0) ? 64 : ((d->cin % 32 == 0) ? 32 : 16);
  p.k_chunks = d->cin / p.block_k;
  const CUtensorMapSwizzle swz = (p.block_k == 64)   ? CU_TENSOR_MAP_SWIZZLE_128B
                                 : (p.block_k == 32) ? CU_TENSOR_MAP_SWIZZLE_64B
                                                     : CU_TENSOR_MAP_SWIZZLE_32B;
  for (int v = 0; v < B200IR_MAX_VIEWS; ++v) {
    const b200ir_view& a = d->a[v < d->num_views ? v : 0];
    B200IR_REQUIRE(a.ptr != nullptr && a.c >= d->cin, "conv_igemm: view %d invalid", v);
    B200IR_REQUIRE((reinterpret_cast<uintptr_t>(a.ptr) & 15) == 0 && a.stride_w % 8 == 0 && a.stride_h % 8 == 0 &&
                       a.stride_b % 8 == 0,
                   "conv_igemm: view %d not 16B aligned", v);
    cuuint64_t dims[4] = {(cuuint64_t)a.c, (cuuint64_t)a.w, (cuuint64_t)a.h, (cuuint64_t)a.b};
    cuuint64_t strides[3] = {(cuuint64_t)a.stride_w * 2, (cuuint64_t)a.stride_h * 2, (cuuint64_t)a.stride_b * 2};
    cuuint32_t box[4] = {(cuuint32_t)p.block_k, (cuuint32_t)d->tile_w, (cuuint32_t)d->tile_h, (cuuint32_t)d->tile_b};
    if (encode_map(&p.tmap_a[v], a.ptr, 4, dims, strides, box, swz, "activation")) return 1;
  }
  {
    const int k_total = d->num_taps * d->cin;
    cuuint64_t dims[2] = {(cuuint64_t)k_total, (cuuint64_t)d->cout};
    cuuint64_t strides[1] = {(cuuint64_t)k_total * 2};
    cuuint32_t box[2] = {(cuuint32_t)p.block_k, (cuuint32_t)d->block_n};
    if (encode_map(&p.tmap_b, d->weight, 2, dims, strides, box, swz, "weight")) return 1;
  }
  p.tile_w = d->tile_w; p.tile_h = d->tile_h; p.tile_b = d->tile_b;
  p.tiles_w = (d->m_w + d->tile_w - 1) / d->tile_w;
  p.tiles_h = (d->m_h + d->tile_h - 1) / d->tile_h;
  p.tiles_b = (d->m_b + d->tile_b - 1) / d->tile_b;
  p.tiles_n = d->cout / d->block_n;
  p.num_tiles = p.tiles_w * p.tiles_h * p.tiles_b * p.tiles_n;
  p.block_n = d->block_n;
  p.num_taps = d->num_taps;
  p.m_w = d->m_w; p.m_h = d->m_h; p.m_b = d->m_b;
  for (int t = 0; t < d->num_taps; ++t) {
    B200IR_REQUIRE(d->tap_view[t] >= 0 && d->tap_view[t] < d->num_views, "conv_igemm: tap %d view", t);
    p.tap_view[t] = d->tap_view[t]; p.tap_dx[t] = d->tap_dx[t]; p.tap_dy[t] = d->tap_dy[t];
  }
  p.idesc = make_idesc_f16(kBlockM, d->block_n, false);
  p.smem_demod = (d->demod != nullptr && d->tile_b * d->block_n <= kDemodTable) ? 1 : 0;
  p.smem_aux = ((d->out_scale != nullptr || d->rgb_w != nullptr) && 5 * d->tile_b * d->block_n <= kDemodTable) ? 1 : 0;
  p.st256 = (!d->out_fp32 && d->out_c_off % 16 == 0 && d->out_stride_x % 16 == 0 && d->out_stride_y % 16 == 0 &&
             d->out_stride_b % 16 == 0 && (reinterpret_cast<uintptr_t>(d->out) & 31) == 0) ? 1 : 0;
  B200IR_REQUIRE(d->cout <= kMaxBias || (d->bias != nullptr && !d->act),
                 "conv_igemm: cout=%d > %d needs a bias vector and no activation", d->cout, kMaxBias);
  // accumulator ring depth: the epilogue latency of a tile is hidden behind the main loops of the next
  // (acc_stages - 1) tiles; small tiles (short main loops) need a deeper ring
  int acc_stages = 2, acc_shift = 1;
  while (acc_stages < kMaxAccStages && 2 * acc_stages * d->block_n <= 512) {
    acc_stages *= 2;
    ++acc_shift;
  }
  p.acc_stages = acc_stages;
  p.acc_shift = acc_shift;
  uint32_t cols = 32;
  while (cols < (uint32_t)(acc_stages * d->block_n)) cols <<= 1;
  p.tmem_cols = cols;

  const int row_bytes = p.block_k * 2;
  const int stage_bytes = kBlockM * row_bytes + d->block_n * row_bytes;
  const int tail = kTailBytes;
  int stages = (g_smem_optin - 1024 - tail) / stage_bytes;
  if (stages > kMaxStages) stages = kMaxStages;
  B200IR_REQUIRE(stages >= 2, "conv_igemm: not enough shared memory for 2 stages");
  p.stages = stages;
  const int smem_bytes = stages * stage_bytes + tail + 1024;

  p.out = d->out; p.out_fp32 = d->out_fp32;
  p.out_sx = d->out_stride_x; p.out_sy = d->out_stride_y; p.out_sb = d->out_stride_b;
  p.out_c_off = d->out_c_off;
  p.out_x_mul = d->out_x_mul; p.out_x_off = d->out_x_off; p.out_y_mul = d->out_y_mul; p.out_y_off = d->out_y_off;
  p.cout = d->cout;
  p.bias = d->bias; p.demod = d->demod; p.noise = d->noise; p.noise_gain = d->noise_gain;
  p.noise_sb = d->noise_stride_b; p.noise_sy = d->noise_stride_y;
  p.act = d->act; p.act_gain = d->act ? 1.4142135623730951f : 1.f; p.res_mode = d->res_mode; p.res = reinterpret_cast<const __half*>(d->res);
  p.res_sx = d->res_stride_x; p.res_sy = d->res_stride_y; p.res_sb = d->res_stride_b;
  p.res_w = d->res_w; p.res_h = d->res_h; p.res_scale = d->res_scale;
  {
    static int dbg = -1;
    if (dbg < 0) dbg = (getenv("B200IR_DBG_SKIP_EPI") != nullptr) ? 1 : 0;
    p.dbg_skip_epi = dbg;

  }
  // ---- specialised epilogue selection
  {
    const bool fp16_fast = !d->out_fp32 && (d->no_store || p.st256) && d->cout <= kMaxBias;
    int flags = 0;
    bool ok = fp16_fast;
    if (d->demod != nullptr) { flags |= F_DEMOD; ok = ok && p.smem_demod; }
    if (d->noise != nullptr) flags |= F_NOISE;
    if (d->res_mode == 1) flags |= F_RES1;
    if (d->res_mode == 2) flags |= F_RES2;
    if (d->rgb_w != nullptr) { flags |= F_RGB; ok = ok && p.smem_aux; }
    else if (d->out_scale != nullptr) ok = false;
    if (d->no_store) flags |= F_NOSTORE;
    p.epi = -1;
    for (int i = 0; ok && i < kNumEpiProfiles; ++i)
      if (epi_profile_flags(i) == flags) p.epi = i;
    static int force_generic = -1;
    if (force_generic < 0) force_generic = (getenv("B200IR_GENERIC_EPI") != nullptr) ? 1 : 0;
    if (force_generic) p.epi = -1;
    p.slope = d->act ? 0.2f : 1.f;
    if (p.epi >= 0 && d->res_mode != 0) p.act_gain *= d->res_scale;
  }
  p.out_scale = d->out_scale; p.rgb_w = d->rgb_w; p.rgb_part = d->rgb_part; p.no_store = d->no_store;
  p.rgb_w_px = d->rgb_w_px; p.rgb_plane = (long long)d->rgb_h * d->rgb_w_px;
  p.rgb_image = (long long)d->m_b * 3 * p.rgb_plane;

  // ---- row mode eligibility: plain 3x3 stride-1 conv on one view, 128-pixel row tiles, weights fit in smem
  {
    bool row_ok = d->num_views == 1 && d->num_taps == 9 && d->tile_w == 128 && d->tile_h == 1 && d->tile_b == 1 &&
                  d->block_n == d->cout && d->out_x_mul == 1 && d->out_y_mul == 1 && d->out_x_off == 0 &&
                  d->out_y_off == 0 && d->row_mode != 0 &&
                  (d->cout == 16 || d->cout == 32 || d->cout == 64);  // N = 3*cout <= 256, ring of 512/cout slots
    for (int t = 0; row_ok && t < 9; ++t)
      row_ok = d->tap_view[t] == 0 && d->tap_dx[t] == (t % 3) - 1 && d->tap_dy[t] == (t / 3) - 1;
    if (row_ok) {
      const int w_bytes = 9 * p.k_chunks * d->block_n * row_bytes;
      const int slot_bytes = 136 * row_bytes;
      const int tail_r = kTailBytes;
      int slots = (g_smem_optin - 1024 - tail_r - w_bytes) / slot_bytes;
      if (slots > kMaxStages) slots = kMaxStages;
      if (slots >= 2 * p.k_chunks) {
        // re-encode the activation map with the 130-pixel halo box
        const b200ir_view& a = d->a[0];
        cuuint64_t dims[4] = {(cuuint64_t)a.c, (cuuint64_t)a.w, (cuuint64_t)a.h, (cuuint64_t)a.b};
        cuuint64_t strides[3] = {(cuuint64_t)a.stride_w * 2, (cuuint64_t)a.stride_h * 2, (cuuint64_t)a.stride_b * 2};
        cuuint32_t box[4] = {(cuuint32_t)p.block_k, 130u, 1u, 1u};
        if (encode_map(&p.tmap_a[0], a.ptr, 4, dims, strides, box, swz, "activation(row)")) return 1;
        int R = d->m_h;
        while ((long long)d->m_b * p.tiles_w * ((d->m_h + R - 1) / R) < 4LL * g_num_sms && R > 8) R = (R + 1) / 2;
        p.row_R = R;
        p.row_chunks = (d->m_h + R - 1) / R;
        p.row_items = d->m_b * p.tiles_w * p.row_chunks;
        // too little parallelism at small batch: generic tiles are faster (row_mode == 2 forces the variant: tests)
        row_ok = p.row_items >= 2 * g_num_sms || d->row_mode == 2;
        p.row_slots = slots;
        p.row_slot_bytes = slot_bytes;
        p.row_w_bytes = w_bytes;
        p.desc_mode = 0;
        for (int n = 1; n <= 3; ++n) p.idesc_n[n - 1] = make_idesc_f16(kBlockM, n * d->block_n, false);
        const int smem_row = w_bytes + slots * slot_bytes + tail_r + 1024;
        if (row_ok) {
          int grid_r = p.row_items < g_num_sms ? p.row_items : g_num_sms;
          if (d->max_ctas > 0 && grid_r > d->max_ctas) grid_r = d->max_ctas;
          cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
          if (p.block_k == 64) {
            if (configure_smem(conv_row_kernel<64>, 3)) return 1;
            conv_row_kernel<64><<<grid_r, kThreads, smem_row, st>>>(p);
          } else if (p.block_k == 32) {
            if (configure_smem(conv_row_kernel<32>, 4)) return 1;
            conv_row_kernel<32><<<grid_r, kThreads, smem_row, st>>>(p);
          } else {
            if (configure_smem(conv_row_kernel<16>, 5)) return 1;
            conv_row_kernel<16><<<grid_r, kThreads, smem_row, st>>>(p);
          }
          return check_launch("conv_row");
        }
        // not taken: restore the generic activation map (box = tile)
        cuuint32_t box_g[4] = {(cuuint32_t)p.block_k, (cuuint32_t)d->tile_w, (cuuint32_t)d->tile_h, (cuuint32_t)d->tile_b};
        if (encode_map(&p.tmap_a[0], a.ptr, 4, dims, strides, box_g, swz, "activation")) return 1;
      }
    }
  }

  int grid = p.num_tiles < g_num_sms ? p.num_tiles : g_num_sms;
  if (d->max_ctas > 0 && grid > d->max_ctas) grid = d->max_ctas;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (p.block_k == 64) {
    if (configure_smem(conv_igemm_kernel<64>, 0)) return 1;
    conv_igemm_kernel<64><<<grid, kThreads, smem_bytes, st>>>(p);
  } else if (p.block_k == 32) {
    if (configure_smem(conv_igemm_kernel<32>, 1)) return 1;
    conv_igemm_kernel<32><<<grid, kThreads, smem_bytes, st>>>(p);
  } else {
    if (configure_smem(conv_igemm_kernel<16>, 2)) return 1;
    conv_igemm_kernel<16><<<grid, kThreads, smem_bytes, st>>>(p);
  }
  return check_launch("conv_igemm");
}
