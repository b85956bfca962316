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
#include <string.h>

#include "host_common.h"
#include "ptx.cuh"

namespace b200ir {

static constexpr int kBlockM = 128;
static constexpr int kThreads = 192;
static constexpr int kMaxStages = 8;
static constexpr int kAccStages = 2;

struct alignas(64) ConvParams {
  CUtensorMap tmap_a[B200IR_MAX_VIEWS];
  CUtensorMap tmap_b;
  int tiles_w, tiles_h, tiles_b, tiles_n, num_tiles;
  int tile_w, tile_h, tile_b;
  int block_n, block_k, k_chunks, num_taps;
  int m_w, m_h, m_b;
  int stages;
  uint32_t idesc;
  uint32_t tmem_cols;
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

__device__ __forceinline__ void load_half16(const __half* ptr, float (&f)[16]) {
  const uint4* q = reinterpret_cast<const uint4*>(ptr);
  uint4 a = __ldg(q), b = __ldg(q + 1);
  const __half2* ha = reinterpret_cast<const __half2*>(&a);
  const __half2* hb = reinterpret_cast<const __half2*>(&b);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 x = __half22float2(ha[i]);
    float2 y = __half22float2(hb[i]);
    f[2 * i] = x.x;
    f[2 * i + 1] = x.y;
    f[8 + 2 * i] = y.x;
    f[8 + 2 * i + 1] = y.y;
  }
}

__global__ void __launch_bounds__(kThreads, 1) conv_igemm_kernel(const __grid_constant__ ConvParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t pad = ((raw_addr + 1023u) & ~1023u) - raw_addr;
  uint8_t* smem = smem_raw + pad;

  const uint32_t row_bytes = p.block_k * 2;
  const uint32_t a_bytes = kBlockM * row_bytes;
  const uint32_t b_bytes = p.block_n * row_bytes;
  const uint32_t stage_bytes = a_bytes + b_bytes;

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.stages * stage_bytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kMaxStages;
  uint64_t* tmem_full = bars + 2 * kMaxStages;
  uint64_t* tmem_empty = bars + 2 * kMaxStages + kAccStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kMaxStages + 2 * kAccStages);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < p.stages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < kAccStages; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 128);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, p.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int num_kb = p.num_taps * p.k_chunks;

  if (warp == 0) {
    // ===================== TMA producer (one lane) =====================
    if (lane == 0) {
      for (int v = 0; v < B200IR_MAX_VIEWS; ++v) tma_prefetch_desc(&p.tmap_a[v]);
      tma_prefetch_desc(&p.tmap_b);
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        const TileCoord t = decode_tile(p, tile);
        for (int tap = 0; tap < p.num_taps; ++tap) {
          const int view = p.tap_view[tap];
          const int cx = t.x0 + p.tap_dx[tap];
          const int cy = t.y0 + p.tap_dy[tap];
          for (int kc = 0; kc < p.k_chunks; ++kc) {
            mbar_wait(&empty_bar[stage], phase ^ 1u);
            uint8_t* sa = smem + stage * stage_bytes;
            uint8_t* sb = sa + a_bytes;
            mbar_arrive_expect_tx(&full_bar[stage], stage_bytes);
            tma_load_4d(sa, &p.tmap_a[view], &full_bar[stage], kc * p.block_k, cx, cy, t.b0);
            tma_load_2d(sb, &p.tmap_b, &full_bar[stage], (tap * p.k_chunks + kc) * p.block_k, t.n0);
            if (++stage == p.stages) {
              stage = 0;
              phase ^= 1u;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one lane) =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      const int k_steps = p.block_k / 16;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * p.block_n;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * stage_bytes);
          const uint32_t sb = sa + a_bytes;
          for (int k = 0; k < k_steps; ++k) {
            const uint64_t da = make_kmajor_desc(sa + k * 32, row_bytes);
            const uint64_t db = make_kmajor_desc(sb + k * 32, row_bytes);
            umma_f16(tmem_d, da, db, p.idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);
          if (kb == num_kb - 1) umma_commit(&tmem_full[acc]);
          if (++stage == p.stages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else {
    // ===================== epilogue (4 warps, TMEM lane quarter = warp % 4) =====================
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int xx = row % p.tile_w;
    const int yy = (row / p.tile_w) % p.tile_h;
    const int bi = row / (p.tile_w * p.tile_h);
    const float gain = (p.noise != nullptr) ? __ldg(p.noise_gain) : 0.f;
    int it = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const TileCoord t = decode_tile(p, tile);
      const int x = t.x0 + xx, y = t.y0 + yy, b = t.b0 + bi;
      const bool valid = (x < p.m_w) && (y < p.m_h) && (b < p.m_b);
      const int xo = x * p.out_x_mul + p.out_x_off;
      const int yo = y * p.out_y_mul + p.out_y_off;
      const long long out_off = (long long)b * p.out_sb + (long long)yo * p.out_sy + (long long)xo * p.out_sx +
                                p.out_c_off + t.n0;
      float nz = 0.f;
      if (valid && p.noise != nullptr) nz = gain * __ldg(p.noise + b * p.noise_sb + yo * p.noise_sy + xo);
      // residual source addressing
      const __half* r00 = nullptr;
      const __half* r01 = nullptr;
      const __half* r10 = nullptr;
      const __half* r11 = nullptr;
      float wy0 = 0.f, wy1 = 0.f, wx0 = 0.f, wx1 = 0.f;
      if (valid && p.res_mode == 1) {
        r00 = p.res + (long long)b * p.res_sb + (long long)yo * p.res_sy + (long long)xo * p.res_sx + t.n0;
      } else if (valid && p.res_mode == 2) {
        // F.interpolate(scale 2, bilinear, align_corners=False): even 2k -> .25*x[k-1] + .75*x[k], odd 2k+1 ->
        // .75*x[k] + .25*x[k+1], indices clamped to the tensor.
        const int ky = yo >> 1, kx = xo >> 1;
        int ya, yb, xa, xb;
        if (yo & 1) { ya = ky; yb = min(ky + 1, p.res_h - 1); wy0 = 0.75f; wy1 = 0.25f; }
        else        { ya = max(ky - 1, 0); yb = ky; wy0 = 0.25f; wy1 = 0.75f; }
        if (xo & 1) { xa = kx; xb = min(kx + 1, p.res_w - 1); wx0 = 0.75f; wx1 = 0.25f; }
        else        { xa = max(kx - 1, 0); xb = kx; wx0 = 0.25f; wx1 = 0.75f; }
        const __half* rb = p.res + (long long)b * p.res_sb + t.n0;
        r00 = rb + (long long)ya * p.res_sy + (long long)xa * p.res_sx;
        r01 = rb + (long long)ya * p.res_sy + (long long)xb * p.res_sx;
        r10 = rb + (long long)yb * p.res_sy + (long long)xa * p.res_sx;
        r11 = rb + (long long)yb * p.res_sy + (long long)xb * p.res_sx;
      }

      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * p.block_n;
      for (int c0 = 0; c0 < p.block_n; c0 += 16) {
        uint32_t raw[16];
        tmem_ld16(taddr + c0, raw);
        tmem_ld_wait();
        if (valid) {
          float v[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(raw[j]);
          const int n = t.n0 + c0;
          if (p.demod != nullptr) {
            const float4* dp = reinterpret_cast<const float4*>(p.demod + (long long)b * p.cout + n);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              float4 d = __ldg(dp + j);
              v[4 * j] *= d.x; v[4 * j + 1] *= d.y; v[4 * j + 2] *= d.z; v[4 * j + 3] *= d.w;
            }
          }
          if (p.bias != nullptr) {
            const float4* bp = reinterpret_cast<const float4*>(p.bias + n);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              float4 d = __ldg(bp + j);
              v[4 * j] += d.x; v[4 * j + 1] += d.y; v[4 * j + 2] += d.z; v[4 * j + 3] += d.w;
            }
          }
          if (p.noise != nullptr) {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] += nz;
          }
          if (p.act) {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = (v[j] > 0.f ? v[j] : 0.2f * v[j]) * 1.4142135623730951f;
          }
          if (p.res_mode == 1) {
            float r[16];
            load_half16(r00 + c0, r);
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = (v[j] + r[j]) * p.res_scale;
          } else if (p.res_mode == 2) {
            float ra[16], rb[16], rc[16], rd[16];
            load_half16(r00 + c0, ra);
            load_half16(r01 + c0, rb);
            load_half16(r10 + c0, rc);
            load_half16(r11 + c0, rd);
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float up = wy0 * (wx0 * ra[j] + wx1 * rb[j]) + wy1 * (wx0 * rc[j] + wx1 * rd[j]);
              v[j] = (v[j] + up) * p.res_scale;
            }
          }
          if (p.out_fp32) {
            float4* op = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + out_off + c0);
#pragma unroll
            for (int j = 0; j < 4; ++j) op[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          } else {
            uint32_t pk[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              __half2 h = __floats2half2_rn(v[2 * j], v[2 * j + 1]);
              pk[j] = *reinterpret_cast<uint32_t*>(&h);
            }
            uint4* op = reinterpret_cast<uint4*>(reinterpret_cast<__half*>(p.out) + out_off + c0);
            op[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            op[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&tmem_empty[acc]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

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

static int encode_map(CUtensorMap* m, const void* ptr, int rank, const cuuint64_t* dims, const cuuint64_t* strides_b,
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

}  // namespace b200ir

using namespace b200ir;

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
  B200IR_REQUIRE(d->out != nullptr && d->weight != nullptr, "conv_igemm: null out/weight");
  B200IR_REQUIRE(d->m_w > 0 && d->m_h > 0 && d->m_b > 0, "conv_igemm: empty M extents");
  B200IR_REQUIRE((d->out_c_off % 8) == 0 && (d->out_stride_x % 8) == 0, "conv_igemm: output not 16B aligned");
  B200IR_REQUIRE(d->res_mode >= 0 && d->res_mode <= 2, "conv_igemm: res_mode");
  B200IR_REQUIRE(d->res_mode == 0 || d->res != nullptr, "conv_igemm: res_mode set but res is NULL");
  B200IR_REQUIRE(d->noise == nullptr || d->noise_gain != nullptr, "conv_igemm: noise without noise_gain");

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
  uint32_t cols = 32;
  while (cols < (uint32_t)(kAccStages * d->block_n)) cols <<= 1;
  p.tmem_cols = cols;

  const int row_bytes = p.block_k * 2;
  const int stage_bytes = kBlockM * row_bytes + d->block_n * row_bytes;
  const int tail = (2 * kMaxStages + 2 * kAccStages) * 8 + 16;
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
  p.act = d->act; p.res_mode = d->res_mode; p.res = reinterpret_cast<const __half*>(d->res);
  p.res_sx = d->res_stride_x; p.res_sy = d->res_stride_y; p.res_sb = d->res_stride_b;
  p.res_w = d->res_w; p.res_h = d->res_h; p.res_scale = d->res_scale;

  static int configured_smem = 0;
  if (smem_bytes > configured_smem) {
    cudaError_t e = cudaFuncSetAttribute(conv_igemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, g_smem_optin);
    if (e != cudaSuccess) {
      set_error("conv_igemm: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return 1;
    }
    configured_smem = g_smem_optin;
  }
  int grid = p.num_tiles < g_num_sms ? p.num_tiles : g_num_sms;
  if (d->max_ctas > 0 && grid > d->max_ctas) grid = d->max_ctas;
  conv_igemm_kernel<<<grid, kThreads, smem_bytes, reinterpret_cast<cudaStream_t>(stream)>>>(p);
  return check_launch("conv_igemm");
}
