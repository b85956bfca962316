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
#include <stdlib.h>
#include <string.h>

#include "host_common.h"
#include "ptx.cuh"

namespace b200ir {

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

// horizontally filtered row (unscaled taps 1,3,3,1) of this thread's XP outputs x 8 channels
template <int XP>
struct FirRow {
  float v[XP][8];
};

// CC: channels per chunk (TMA box inner extent); XP: horizontally adjacent outputs per thread; SR: input rows per stage;
// NT: consumer threads.  SFT (when present) is uniform per chunk: the host requires c_keep % CC == 0.
template <int CC, int XP, int SR, int NT, bool POST>
__global__ void __launch_bounds__(NT + 32, (NT == 256 && !(POST && XP == 2)) ? 2 : 1) fir_stream_kernel(const __grid_constant__ FirParams p) {
  constexpr int CG = CC / 8;
  constexpr int XQ = NT / CG;
  constexpr int TW = XQ * XP;
  constexpr int IW = TW + 3;
  static_assert(SR == 2 || SR == 4, "the register window rotates with period 4");
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
      mbar_init(&empty_bar[i], NT);
    }
    fence_barrier_init();
  }
  __syncthreads();
  const int my_items = (p.num_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  if (tid >= NT) {
    // ---------------- producer warp
    if (tid == NT) {
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
        const bool sft = POST && p.has_sft && (c0 >= p.c_keep);
        const uint32_t bytes = p.in_bytes + (sft ? 2 * p.sft_bytes : 0);
        for (int k = 0; k < p.KS; ++k) {
          mbar_wait(&empty_bar[slot], phase ^ 1u);
          uint8_t* st = base + slot * p.stage_bytes;
          mbar_arrive_expect_tx(&full_bar[slot], bytes);
          tma_load_4d(st, &p.tmap_in, &full_bar[slot], c0, t.strip * TW - p.pad, t.y0 - p.pad + k * SR, t.b);
          if (sft) {
            tma_load_4d(st + p.in_bytes, &p.tmap_scale, &full_bar[slot], c0 - p.c_keep, t.strip * TW,
                        t.y0 + k * SR - 3, t.b);
            tma_load_4d(st + p.in_bytes + p.sft_bytes, &p.tmap_shift, &full_bar[slot], c0 - p.c_keep, t.strip * TW,
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
  const uint32_t thr_in = smem_u32(base) + (xl * CC + cg * 8) * 2;       // this thread's first input pixel, row 0
  const uint32_t thr_sft = smem_u32(base) + p.in_bytes + (xl * CC + cg * 8) * 2;
  FirRow<XP> w0, w1, w2, w3;  // rotating window of the last four horizontally filtered rows
#pragma unroll
  for (int j = 0; j < XP; ++j)
#pragma unroll
    for (int e = 0; e < 8; ++e) w0.v[j][e] = w1.v[j][e] = w2.v[j][e] = w3.v[j][e] = 0.f;
  const float gain = (POST && p.noise != nullptr) ? __ldg(p.noise_gain) * 1.4142135623730951f : 0.f;
  const float ks2 = p.kscale2;
  int slot = 0;
  uint32_t phase = 0;
  int g = 0;  // stages consumed so far (window rotation phase for SR == 2)
  for (int it = 0; it < my_items; ++it) {
    const FirItem t = fir_decode(p, blockIdx.x + it * gridDim.x);
    const int c0 = t.chunk * CC;
    const int c = c0 + cg * 8;
    const int x0 = t.strip * TW + xl;
    const int y_end = min(t.y0 + p.R, p.OH);
    bool xv[XP];
    bool any_x = false;
#pragma unroll
    for (int j = 0; j < XP; ++j) {
      xv[j] = x0 + j < p.OW;
      any_x = any_x || xv[j];
    }
    if (!any_x) {
      // ragged last strip (OW = 385, 193, 97 ...): threads past the image only keep the stage ring moving, so a strip
      // that holds a single column costs its (mostly out-of-bounds, i.e. free) TMA traffic and little else
      for (int k = 0; k < p.KS; ++k, ++g) {
        mbar_wait(&full_bar[slot], phase);
        mbar_arrive(&empty_bar[slot]);
        if (++slot == p.NS) {
          slot = 0;
          phase ^= 1u;
        }
      }
      continue;
    }
    float bs[8], sn[8];
    const bool sft = POST && p.has_sft && (c0 >= p.c_keep);
    if (POST) {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        bs[e] = __ldg(p.bias + c + e) * 1.4142135623730951f;
        sn[e] = (p.s_next != nullptr) ? __ldg(p.s_next + (long long)t.b * p.C + c + e) : 1.f;
      }
    }
    // row pointers advance by one output row per input row; the first three rows of an item only prime the window
    __half* out_row = p.out + (long long)t.b * p.out_sb + (long long)(t.y0 - 3) * p.out_sy + (long long)x0 * p.C + c;
    const float* nz_row = (POST && p.noise != nullptr) ? p.noise + t.b * p.noise_sb + (long long)(t.y0 - 3) * p.OW + x0
                                                       : nullptr;
    int oy = t.y0 - 3;
    for (int k = 0; k < p.KS; ++k, ++g) {
      // noise of the output rows this stage completes (independent of the TMA data: fetched before the wait)
      float nz[SR][XP];
#pragma unroll
      for (int r = 0; r < SR; ++r)
#pragma unroll
        for (int j = 0; j < XP; ++j) {
          nz[r][j] = 0.f;
          if (POST && nz_row != nullptr && oy + r >= t.y0 && oy + r < y_end && xv[j])
            nz[r][j] = gain * __ldg(nz_row + (long long)r * p.OW + j);
        }
      mbar_wait(&full_bar[slot], phase);
      const uint32_t st_off = slot * p.stage_bytes;

      // one input row: horizontal pass into D, then (if the output row exists) vertical pass over A, B, C, D
      auto do_row = [&](const FirRow<XP>& A, const FirRow<XP>& B, const FirRow<XP>& C, FirRow<XP>& D, int r) {
        float f[XP + 3][8];
#pragma unroll
        for (int col = 0; col < XP + 3; ++col) unpack8(lds128(thr_in + st_off + (r * IW + col) * CC * 2), f[col]);
#pragma unroll
        for (int j = 0; j < XP; ++j)
#pragma unroll
          for (int e = 0; e < 8; ++e) D.v[j][e] = fmaf(3.f, f[j + 1][e] + f[j + 2][e], f[j][e] + f[j + 3][e]);
        const int y = oy + r;
        if (y >= t.y0 && y < y_end) {
#pragma unroll
          for (int j = 0; j < XP; ++j) {
            if (xv[j]) {
              float v[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) v[e] = fmaf(3.f, B.v[j][e] + C.v[j][e], A.v[j][e] + D.v[j][e]);
              if (POST) {
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                  const float a = fmaf(v[e], ks2, nz[r][j] + bs[e]);  // ks2 / noise / bias carry the sqrt(2) gain
                  v[e] = fmaxf(a, 0.2f * a);
                }
                if (sft) {
                  const uint32_t sa = thr_sft + st_off + ((r * TW + j) * CC) * 2;
                  float sc[8], sh[8];
                  unpack8(lds128(sa), sc);
                  unpack8(lds128(sa + p.sft_bytes), sh);
#pragma unroll
                  for (int e = 0; e < 8; ++e) v[e] = fmaf(v[e], sc[e], sh[e]);
                }
#pragma unroll
                for (int e = 0; e < 8; ++e) v[e] *= sn[e];
              } else {
#pragma unroll
                for (int e = 0; e < 8; ++e) v[e] *= ks2;
              }
              uint4 o;
              __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
              for (int e = 0; e < 4; ++e) oh[e] = f2h2_sat(v[2 * e], v[2 * e + 1]);
              *reinterpret_cast<uint4*>(out_row + (long long)r * p.out_sy + (long long)j * p.C) = o;
            }
          }
        }
      };
      if (SR == 4) {
        do_row(w0, w1, w2, w3, 0);
        do_row(w1, w2, w3, w0, 1);
        do_row(w2, w3, w0, w1, 2);
        do_row(w3, w0, w1, w2, 3);
      } else if (g & 1) {
        do_row(w2, w3, w0, w1, 0);
        do_row(w3, w0, w1, w2, 1);
      } else {
        do_row(w0, w1, w2, w3, 0);
        do_row(w1, w2, w3, w0, 1);
      }
      mbar_arrive(&empty_bar[slot]);
      if (++slot == p.NS) {
        slot = 0;
        phase ^= 1u;
      }
      oy += SR;
      out_row += (long long)SR * p.out_sy;
      if (nz_row != nullptr) nz_row += (long long)SR * p.OW;
    }
  }
}

template <int CC, int XP, int SR, int NT, bool POST>
static int fir_launch_variant(const FirLaunch& a, cudaStream_t st, const char* what) {
  constexpr int CG = CC / 8, XQ = NT / CG, TW = XQ * XP, IW = TW + 3;
  constexpr int kThreads = NT + 32;
  FirParams p;
  memset(&p, 0, sizeof(p));
  p.B = a.B; p.OH = a.OH; p.OW = a.OW; p.C = a.C; p.pad = a.pad;
  p.strips = (a.OW + TW - 1) / TW;
  p.chunks = a.C / CC;
  const int sms = num_sms();
  if (sms == 0) return 1;
  int R = 32;
  if (const char* e = getenv("B200IR_FIR_R")) R = atoi(e) > 0 ? atoi(e) : R;
  while (R > 8 && (long long)a.B * p.strips * p.chunks * ((a.OH + R - 1) / R) < 4LL * sms) R /= 2;
  if (R > a.OH) R = a.OH;
  R = (a.OH + (a.OH + R - 1) / R - 1) / ((a.OH + R - 1) / R);  // equal row chunks (OH = 33 -> 17 + 16, not 32 + 1)
  p.R = R;
  p.rchunks = (a.OH + R - 1) / R;
  p.KS = (R + 3 + SR - 1) / SR;
  p.num_items = a.B * p.strips * p.chunks * p.rchunks;
  p.in_bytes = SR * IW * CC * 2;
  static_assert((SR * IW * CC * 2) % 128 == 0, "stage sub-buffers must stay 128-byte aligned");
  p.has_sft = (POST && a.scale != nullptr) ? 1 : 0;
  p.c_keep = a.C - a.c_sft;
  if (p.has_sft) {
    B200IR_REQUIRE(a.c_sft % CC == 0 && p.c_keep % CC == 0, "%s: c_sft=%d unsupported (C=%d)", what, a.c_sft, a.C);
    p.sft_bytes = SR * TW * CC * 2;
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
    cudaFuncSetAttribute(fir_stream_kernel<CC, XP, SR, NT, POST>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fir_stream_kernel<CC, XP, SR, NT, POST>, kThreads,
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
    cuuint32_t box[4] = {(cuuint32_t)CC, (cuuint32_t)TW, (cuuint32_t)SR, 1u};
    if (encode_map(&p.tmap_scale, a.scale, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE, what)) return 1;
    if (encode_map(&p.tmap_shift, a.shift, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE, what)) return 1;
  }
  const int smem_bytes = p.NS * p.stage_bytes + 256 + 128;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(fir_stream_kernel<CC, XP, SR, NT, POST>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max);
    if (e != cudaSuccess) {
      set_error("%s: cudaFuncSetAttribute: %s", what, cudaGetErrorString(e));
      return 1;
    }
    configured = true;
  }
  int grid = ctas_per_sm * sms;
  if (grid > p.num_items) grid = p.num_items;
  fir_stream_kernel<CC, XP, SR, NT, POST><<<grid, kThreads, smem_bytes, st>>>(p);
  return check_launch(what);
}

// cost of covering ow outputs with strips of tw: full strips count fully, a ragged last strip by its real width plus a
// fixed share (its idle threads skip the arithmetic, its out-of-bounds TMA traffic is free); smaller is better
static inline double fir_waste(int ow, int tw) {
  const int full = ow / tw, rem = ow - full * tw;
  return ((double)full * tw + (rem ? rem + 0.15 * tw : 0.0)) / ow;
}

template <int CC, bool POST>
static int fir_dispatch(const FirLaunch& a, cudaStream_t st, const char* what) {
  // candidates: (XP, NT) -> TW = NT / (CC/8) * XP; POST stages carry scale/shift tiles, so they use SR = 2 when XP = 2
  constexpr int CG = CC / 8;
  int force_xp = 0, force_nt = 0;
  if (const char* e = getenv("B200IR_FIR_XP")) force_xp = atoi(e);
  if (const char* e = getenv("B200IR_FIR_NT")) force_nt = atoi(e);
  int best_xp = 0, best_nt = 0;
  double best = 1e30;
  for (int nt = 384; nt >= 256; nt -= 128)
    for (int xp = 2; xp >= 1; --xp) {
      if ((force_xp && xp != force_xp) || (force_nt && nt != force_nt)) continue;
      // preference at equal waste: more consumer warps, then two outputs per thread (fewer conversions per output)
      const double w = fir_waste(a.OW, nt / CG * xp);
      if (w < best - 1e-9) {
        best = w;
        best_xp = xp;
        best_nt = nt;
      }
    }
  if (best_nt == 384)
    return best_xp == 2 ? fir_launch_variant<CC, 2, POST ? 2 : 4, 384, POST>(a, st, what)
                        : fir_launch_variant<CC, 1, 4, 384, POST>(a, st, what);
  return best_xp == 2 ? fir_launch_variant<CC, 2, POST ? 2 : 4, 256, POST>(a, st, what)
                      : fir_launch_variant<CC, 1, 4, 256, POST>(a, st, what);
}

// returns -1 when the shape is not eligible for the streaming kernels (caller falls back to the direct kernels)
int fir_stream_launch(const FirLaunch& a, cudaStream_t st, const char* what) {
  if (a.C % 32 != 0 || (reinterpret_cast<uintptr_t>(a.in) & 15) != 0 || a.in_sw % 8 || a.in_sh % 8 || a.in_sb % 8)
    return -1;
  const bool sft = a.post && a.scale != nullptr;
  if (sft && (a.c_sft % 32 != 0 || (a.C - a.c_sft) % 32 != 0)) return -1;
  // 64-channel chunks unless the SFT boundary (or C) only aligns to 32: SFT must be uniform per chunk
  const bool cc64 = a.C % 64 == 0 && (!sft || (a.c_sft % 64 == 0 && (a.C - a.c_sft) % 64 == 0));
  if (cc64) return a.post ? fir_dispatch<64, true>(a, st, what) : fir_dispatch<64, false>(a, st, what);
  return a.post ? fir_dispatch<32, true>(a, st, what) : fir_dispatch<32, false>(a, st, what);
}

}  // namespace b200ir
