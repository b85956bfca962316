// Streaming 4x4 FIR kernels (upfirdn2d with outer([1,3,3,1]) taps) for NHWC fp16 tensors, fed by TMA.
//
// Replaces, on the GFPGANv1OCR path, upfirdn2d (basicsr/ops/upfirdn2d/upfirdn2d.py:162-192) in its two stride-1 uses:
//   * UpFirDnSmooth(pad (2,2)) before the stride-2 3x3 conv of ResBlock (stylegan2_ocr_arch.py:116-121,685-697);
//   * the FIR*4 (pad (1,1)) after the stride-2 transposed conv of an upsampling StyleConv (:261-267), fused with
//     noise + FusedLeakyReLU (:323-333), the SFT affine (gfpganv1_ocr_arch.py:118-125) and the modulation of the
//     next conv (stylegan2_ocr_arch.py:247-251).
//
// HBM-bound.  A work item is (image, column strip of TW outputs, chunk of CC channels, chunk of R output rows).  A
// producer warp streams the item's input rows through a ring of shared-memory stages with cp.async.bulk.tensor (the
// zero padding of upfirdn2d is the TMA out-of-bounds fill; the SFT scale/shift tiles ride in the same stage), 8 consumer
// warps run the separable filter: one horizontal pass per input row out of shared memory (128-bit, conflict-free), the
// vertical pass over a register window of the last 4 filtered rows.  Every input element is read from L2/HBM once per
// item (+3 halo rows per R), every output written once with 128-bit stores.
#include <string.h>

#include "host_common.h"
#include "ptx.cuh"

namespace b200ir {

static constexpr int kFirConsumers = 256;
static constexpr int kFirThreads = kFirConsumers + 32;
static constexpr int kFirMaxStages = 8;

struct alignas(64) FirParams {
  CUtensorMap tmap_in, tmap_scale, tmap_shift;
  int B, OH, OW, C, pad;
  int strips, chunks, rchunks, R, KS, num_items;
  int NS, stage_bytes, in_bytes, sft_bytes;
  float kscale2;  // (per-axis scale)^2, times sqrt(2) when the activation follows
  __half* out;
  long long out_sb, out_sy;  // elements
  const float* noise;
  long long noise_sb;
  const float* noise_gain;
  const float* bias;
  const float* s_next;
  int c_keep, has_sft;
};

__device__ __forceinline__ void unpack8(const uint4& q, float* f) {
  const __half2* h = reinterpret_cast<const __half2*>(&q);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 x = __half22float2(h[i]);
    f[2 * i] = x.x;
    f[2 * i + 1] = x.y;
  }
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}

struct FirItem {
  int b, strip, chunk, y0;
};
__device__ __forceinline__ FirItem fir_decode(const FirParams& p, int item) {
  FirItem t;
  t.chunk = item % p.chunks;
  item /= p.chunks;
  t.strip = item % p.strips;
  item /= p.strips;
  t.y0 = (item % p.rchunks) * p.R;
  t.b = item / p.rchunks;
  return t;
}

// CC: channels per chunk (TMA box inner extent); XP: horizontally adjacent outputs per thread; SR: input rows per stage
template <int CC, int XP, int SR, bool POST>
__global__ void __launch_bounds__(kFirThreads, 1) fir_stream_kernel(const __grid_constant__ FirParams p) {
  constexpr int CG = CC / 8;
  constexpr int XQ = kFirConsumers / CG;
  constexpr int TW = XQ * XP;
  constexpr int IW = TW + 3;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t pad_bytes = ((raw_addr + 127u) & ~127u) - raw_addr;
  uint8_t* base = smem_raw + pad_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(base + p.NS * p.stage_bytes);
  uint64_t* empty_bar = full_bar + kFirMaxStages;
  const int tid = threadIdx.x;
  if (tid == 0) {
    for (int i = 0; i < p.NS; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], kFirConsumers);
    }
    fence_barrier_init();
  }
  __syncthreads();
  const int my_items = (p.num_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  if (tid >= kFirConsumers) {
    // ---------------- producer warp
    if (tid == kFirConsumers) {
      tma_prefetch_desc(&p.tmap_in);
      if (POST && p.has_sft) {
        tma_prefetch_desc(&p.tmap_scale);
        tma_prefetch_desc(&p.tmap_shift);
      }
      int slot = 0;
      uint32_t phase = 0;
      for (int it = 0; it < my_items; ++it) {
        const FirItem t = fir_decode(p, blockIdx.x + it * gridDim.x);
        const int c0 = t.chunk * CC;
        const bool sft = POST && p.has_sft && (c0 + CC > p.c_keep);
        const int cs0 = max(c0, p.c_keep) - p.c_keep;
        const uint32_t bytes = p.in_bytes + (sft ? 2 * p.sft_bytes : 0);
        for (int k = 0; k < p.KS; ++k) {
          mbar_wait(&empty_bar[slot], phase ^ 1u);
          uint8_t* st = base + slot * p.stage_bytes;
          mbar_arrive_expect_tx(&full_bar[slot], bytes);
          tma_load_4d(st, &p.tmap_in, &full_bar[slot], c0, t.strip * TW - p.pad, t.y0 - p.pad + k * SR, t.b);
          if (sft) {
            tma_load_4d(st + p.in_bytes, &p.tmap_scale, &full_bar[slot], cs0, t.strip * TW, t.y0 + k * SR - 3, t.b);
            tma_load_4d(st + p.in_bytes + p.sft_bytes, &p.tmap_shift, &full_bar[slot], cs0, t.strip * TW,
                        t.y0 + k * SR - 3, t.b);
          }
          if (++slot == p.NS) {
            slot = 0;
            phase ^= 1u;
          }
        }
      }
    }
    return;
  }

  // ---------------- consumers
  const int cg = tid % CG;
  const int xq = tid / CG;
  const int xl = xq * XP;  // first output column of this thread inside the strip
  const uint32_t base_addr = smem_u32(base);
  float win[3][XP][8];  // horizontally filtered rows y-3, y-2, y-1 (unscaled taps 1,3,3,1)
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int j = 0; j < XP; ++j)
#pragma unroll
      for (int e = 0; e < 8; ++e) win[a][j][e] = 0.f;
  const float gain = (POST && p.noise != nullptr) ? __ldg(p.noise_gain) * 1.4142135623730951f : 0.f;
  int slot = 0;
  uint32_t phase = 0;
  for (int it = 0; it < my_items; ++it) {
    const FirItem t = fir_decode(p, blockIdx.x + it * gridDim.x);
    const int c0 = t.chunk * CC;
    const int c = c0 + cg * 8;
    const int x0 = t.strip * TW + xl;
    const int y_end = min(t.y0 + p.R, p.OH);
    float bs[8], sn[8];
    bool sft = false;
    uint32_t sft_off = 0;  // byte offset of this thread's 8 channels inside a scale/shift pixel
    uint32_t sft_pix = 0;  // bytes per scale/shift pixel in shared memory
    if (POST) {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        bs[e] = __ldg(p.bias + c + e) * 1.4142135623730951f;
        sn[e] = (p.s_next != nullptr) ? __ldg(p.s_next + (long long)t.b * p.C + c + e) : 1.f;
      }
      if (p.has_sft && c >= p.c_keep) {
        sft = true;
        sft_off = (uint32_t)(c - max(c0, p.c_keep)) * 2u;
        sft_pix = (uint32_t)(c0 + CC - max(c0, p.c_keep)) * 2u;
      }
    }
    __half* out_b = p.out + (long long)t.b * p.out_sb + c;
    for (int k = 0; k < p.KS; ++k) {
      // noise of the output rows this stage completes (independent of the TMA data: fetched before the wait)
      float nz[SR][XP];
      if (POST) {
#pragma unroll
        for (int r = 0; r < SR; ++r) {
          const int oy = t.y0 + k * SR + r - 3;
#pragma unroll
          for (int j = 0; j < XP; ++j) {
            nz[r][j] = 0.f;
            if (p.noise != nullptr && oy >= t.y0 && oy < y_end && x0 + j < p.OW)
              nz[r][j] = gain * __ldg(p.noise + t.b * p.noise_sb + (long long)oy * p.OW + x0 + j);
          }
        }
      }
      mbar_wait(&full_bar[slot], phase);
      const uint32_t st = base_addr + slot * p.stage_bytes;
#pragma unroll
      for (int r = 0; r < SR; ++r) {
        // horizontal pass of input row r: taps (1,3,3,1) over columns xl + j .. xl + j + 3
        float h[XP][8];
#pragma unroll
        for (int j = 0; j < XP; ++j)
#pragma unroll
          for (int e = 0; e < 8; ++e) h[j][e] = 0.f;
        const uint32_t row_addr = st + ((r * IW + xl) * CC + cg * 8) * 2;
#pragma unroll
        for (int col = 0; col < XP + 3; ++col) {
          const uint4 q = lds128(row_addr + col * CC * 2);
          float f[8];
          unpack8(q, f);
#pragma unroll
          for (int j = 0; j < XP; ++j) {
            const int tap = col - j;
            if (tap == 0 || tap == 3) {
#pragma unroll
              for (int e = 0; e < 8; ++e) h[j][e] += f[e];
            } else if (tap == 1 || tap == 2) {
#pragma unroll
              for (int e = 0; e < 8; ++e) h[j][e] = fmaf(3.f, f[e], h[j][e]);
            }
          }
        }
        const int oy = t.y0 + k * SR + r - 3;
        if (oy >= t.y0 && oy < y_end) {
#pragma unroll
          for (int j = 0; j < XP; ++j) {
            if (x0 + j < p.OW) {
              float v[8];
#pragma unroll
              for (int e = 0; e < 8; ++e)
                v[e] = ((win[0][j][e] + h[j][e]) + 3.f * (win[1][j][e] + win[2][j][e])) * p.kscale2;
              if (POST) {
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                  const float a = v[e] + (nz[r][j] + bs[e]);  // kscale2 / noise / bias carry the sqrt(2) gain
                  v[e] = fmaxf(a, 0.2f * a);
                }
                if (sft) {
                  const uint32_t sa = st + p.in_bytes + (r * TW + xl + j) * sft_pix + sft_off;
                  float sc[8], sh[8];
                  unpack8(lds128(sa), sc);
                  unpack8(lds128(sa + p.sft_bytes), sh);
#pragma unroll
                  for (int e = 0; e < 8; ++e) v[e] = fmaf(v[e], sc[e], sh[e]);
                }
#pragma unroll
                for (int e = 0; e < 8; ++e) v[e] *= sn[e];
              }
              uint4 o;
              __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
              for (int e = 0; e < 4; ++e) oh[e] = __floats2half2_rn(v[2 * e], v[2 * e + 1]);
              *reinterpret_cast<uint4*>(out_b + (long long)oy * p.out_sy + (long long)(x0 + j) * p.C) = o;
            }
          }
        }
#pragma unroll
        for (int j = 0; j < XP; ++j)
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            win[0][j][e] = win[1][j][e];
            win[1][j][e] = win[2][j][e];
            win[2][j][e] = h[j][e];
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


template <int CC, int XP, int SR, bool POST>
static int fir_launch_variant(const FirLaunch& a, cudaStream_t st, const char* what) {
  constexpr int CG = CC / 8, XQ = kFirConsumers / CG, TW = XQ * XP, IW = TW + 3;
  FirParams p;
  memset(&p, 0, sizeof(p));
  p.B = a.B; p.OH = a.OH; p.OW = a.OW; p.C = a.C; p.pad = a.pad;
  p.strips = (a.OW + TW - 1) / TW;
  p.chunks = a.C / CC;
  const int sms = num_sms();
  if (sms == 0) return 1;
  int R = 32;
  while (R > 8 && (long long)a.B * p.strips * p.chunks * ((a.OH + R - 1) / R) < 4LL * sms) R /= 2;
  if (R > a.OH) R = a.OH;
  p.R = R;
  p.rchunks = (a.OH + R - 1) / R;
  p.KS = (R + 3 + SR - 1) / SR;
  p.num_items = a.B * p.strips * p.chunks * p.rchunks;
  p.in_bytes = SR * IW * CC * 2;
  static_assert((SR * IW * CC * 2) % 128 == 0, "stage sub-buffers must stay 128-byte aligned");
  p.has_sft = (POST && a.scale != nullptr) ? 1 : 0;
  p.c_keep = a.C - a.c_sft;
  int cs = 0;
  if (p.has_sft) {
    cs = a.c_sft < CC ? a.c_sft : CC;
    B200IR_REQUIRE(a.c_sft % cs == 0 && p.c_keep % cs == 0 && (cs * 2) % 16 == 0, "%s: c_sft=%d unsupported (C=%d)",
                   what, a.c_sft, a.C);
    p.sft_bytes = SR * TW * cs * 2;
  }
  p.stage_bytes = p.in_bytes + 2 * p.sft_bytes;
  const int smem_max = smem_optin();
  int ns = (smem_max - 256 - 128) / p.stage_bytes;
  if (ns > kFirMaxStages) ns = kFirMaxStages;
  B200IR_REQUIRE(ns >= 2, "%s: stage of %d bytes does not fit twice in shared memory", what, p.stage_bytes);
  // two co-resident CTAs (16 consumer warps per SM) when the register budget of this variant allows it
  int ctas_per_sm = 1;
  if (ns >= 4) {
    const int ns2 = ns / 2 > 4 ? 4 : ns / 2;
    int occ = 0;
    cudaFuncSetAttribute(fir_stream_kernel<CC, XP, SR, POST>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fir_stream_kernel<CC, XP, SR, POST>, kFirThreads,
                                                      ns2 * p.stage_bytes + 256 + 128) == cudaSuccess && occ >= 2) {
      ns = ns2;
      ctas_per_sm = 2;
    }
  }
  p.NS = ns;
  p.kscale2 = a.kscale * a.kscale * (POST ? 1.4142135623730951f : 1.f);
  p.out = a.out; p.out_sb = a.out_sb; p.out_sy = a.out_sy;
  p.noise = a.noise; p.noise_sb = a.noise_sb; p.noise_gain = a.noise_gain; p.bias = a.bias; p.s_next = a.s_next;
  {
    cuuint64_t dims[4] = {(cuuint64_t)a.C, (cuuint64_t)a.Wv, (cuuint64_t)a.Hv, (cuuint64_t)a.B};
    cuuint64_t strides[3] = {(cuuint64_t)a.in_sw * 2, (cuuint64_t)a.in_sh * 2, (cuuint64_t)a.in_sb * 2};
    cuuint32_t box[4] = {(cuuint32_t)CC, (cuuint32_t)IW, (cuuint32_t)SR, 1u};
    if (encode_map(&p.tmap_in, a.in, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE, what)) return 1;
  }
  if (p.has_sft) {
    cuuint64_t dims[4] = {(cuuint64_t)a.c_sft, (cuuint64_t)a.OW, (cuuint64_t)a.OH, (cuuint64_t)a.B};
    cuuint64_t strides[3] = {(cuuint64_t)a.c_sft * 2, (cuuint64_t)a.OW * a.c_sft * 2,
                             (cuuint64_t)a.OH * a.OW * a.c_sft * 2};
    cuuint32_t box[4] = {(cuuint32_t)cs, (cuuint32_t)TW, (cuuint32_t)SR, 1u};
    if (encode_map(&p.tmap_scale, a.scale, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE, what)) return 1;
    if (encode_map(&p.tmap_shift, a.shift, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE, what)) return 1;
  }
  const int smem_bytes = p.NS * p.stage_bytes + 256 + 128;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(fir_stream_kernel<CC, XP, SR, POST>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max);
    if (e != cudaSuccess) {
      set_error("%s: cudaFuncSetAttribute: %s", what, cudaGetErrorString(e));
      return 1;
    }
    configured = true;
  }
  int grid = ctas_per_sm * sms;
  if (grid > p.num_items) grid = p.num_items;
  fir_stream_kernel<CC, XP, SR, POST><<<grid, kFirThreads, smem_bytes, st>>>(p);
  return check_launch(what);
}

// strips * TW / OW: fraction of the x slots that fall on real outputs (inverse); smaller is better
static inline double fir_waste(int ow, int tw) { return (double)((ow + tw - 1) / tw) * tw / ow; }

// returns -1 when the shape is not eligible for the streaming kernels (caller falls back to the direct kernels)
int fir_stream_launch(const FirLaunch& a, cudaStream_t st, const char* what) {
  if (a.C % 32 != 0 || (reinterpret_cast<uintptr_t>(a.in) & 15) != 0 || a.in_sw % 8 || a.in_sh % 8 || a.in_sb % 8)
    return -1;
  if (a.post && a.scale != nullptr) {
    const int cc = (a.C % 64 == 0) ? 64 : 32;
    const int cs = a.c_sft < cc ? a.c_sft : cc;
    if (cs <= 0 || a.c_sft % cs || (a.C - a.c_sft) % cs || cs % 8) return -1;
  }
  if (a.C % 64 == 0) {
    const bool xp2 = fir_waste(a.OW, 64) <= fir_waste(a.OW, 32);
    if (a.post) return xp2 ? fir_launch_variant<64, 2, 2, true>(a, st, what) : fir_launch_variant<64, 1, 4, true>(a, st, what);
    return xp2 ? fir_launch_variant<64, 2, 4, false>(a, st, what) : fir_launch_variant<64, 1, 4, false>(a, st, what);
  }
  const bool xp2 = fir_waste(a.OW, 128) <= fir_waste(a.OW, 64);
  if (a.post) return xp2 ? fir_launch_variant<32, 2, 2, true>(a, st, what) : fir_launch_variant<32, 1, 4, true>(a, st, what);
  return xp2 ? fir_launch_variant<32, 2, 4, false>(a, st, what) : fir_launch_variant<32, 1, 4, false>(a, st, what);
}

}  // namespace b200ir
