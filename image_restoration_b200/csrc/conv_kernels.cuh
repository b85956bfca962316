// Device code of the two tcgen05 convolution kernels (included by conv_igemm.cu after ConvParams / epilogue_tile).
//
// Both kernels are persistent and warp-specialised: warp 0 = TMA producer, warp 1 = MMA issuer (+ TMEM owner),
// warps 2..5 = epilogue.  The producer and MMA warps run warp-uniform control flow and let ONE elected lane issue the
// TMA / tcgen05 instructions: a single thread issues dependent instructions only every few cycles, so those loops are
// kept as short as possible (descriptor words precomputed, only 32-bit adds per MMA, no divisions) — measured: with
// ~1000 scalar instructions per output tile the issuing thread, not the tensor pipe or memory, set the pace.
#pragma once

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
static constexpr int kTailBytes = 512 + kMaxBias * 4 + 2 * kDemodTable * 4;

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

__device__ __forceinline__ uint32_t kernel_prologue(const ConvParams& p, const KernelSmem& s, int nslots, int warp) {
  if (threadIdx.x == 0) {
    for (int i = 0; i < nslots; ++i) {
      mbar_init(&s.full_bar[i], 1);
      mbar_init(&s.empty_bar[i], 1);
    }
    for (int i = 0; i < p.acc_stages; ++i) {
      mbar_init(&s.tmem_full[i], 1);
      mbar_init(&s.tmem_empty[i], 128);
    }
    mbar_init(s.w_bar, 1);
    fence_barrier_init();
  }
  if (p.cout <= kMaxBias)
    for (int i = threadIdx.x; i < p.cout; i += kThreads)
      s.bias[i] = (p.bias != nullptr) ? p.bias[i] * p.act_gain : 0.f;
  if (warp == 1) {
    tmem_alloc(s.tmem_slot, p.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  return *s.tmem_slot;
}

__device__ __forceinline__ void kernel_epilogue(const ConvParams& p, uint32_t tmem_base, int warp) {
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

// ================================================================================================ generic tiles
template <int kBlockK>
__global__ void __launch_bounds__(kThreads, 1) conv_igemm_kernel(const __grid_constant__ ConvParams p) {
  extern __shared__ uint8_t smem_raw[];
  constexpr uint32_t row_bytes = kBlockK * 2;
  constexpr uint32_t a_bytes = kBlockM * row_bytes;
  constexpr int k_steps = kBlockK / 16;
  const uint32_t b_bytes = p.block_n * row_bytes;
  const uint32_t stage_bytes = a_bytes + b_bytes;
  const KernelSmem s = carve_smem(smem_raw, p.stages * stage_bytes);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t tmem_base = kernel_prologue(p, s, p.stages, warp);
  const int num_kb = p.num_taps * p.k_chunks;

  if (warp == 0) {
    // ---------------- TMA producer
    if (lane == 0) {
      for (int v = 0; v < B200IR_MAX_VIEWS; ++v) tma_prefetch_desc(&p.tmap_a[v]);
      tma_prefetch_desc(&p.tmap_b);
    }
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      const TileCoord t = decode_tile(p, tile);
      int kb = 0;
      for (int tap = 0; tap < p.num_taps; ++tap) {
        const int view = p.tap_view[tap];
        const int cx = t.x0 + p.tap_dx[tap];
        const int cy = t.y0 + p.tap_dy[tap];
        for (int kc = 0; kc < p.k_chunks; ++kc, ++kb) {
          mbar_wait(&s.empty_bar[stage], phase ^ 1u);
          if (elect_one()) {
            uint8_t* sa = s.base + stage * stage_bytes;
            mbar_arrive_expect_tx(&s.full_bar[stage], stage_bytes);
            tma_load_4d(sa, &p.tmap_a[view], &s.full_bar[stage], kc * kBlockK, cx, cy, t.b0);
            tma_load_2d(sa + a_bytes, &p.tmap_b, &s.full_bar[stage], kb * kBlockK, t.n0);
          }
          __syncwarp();
          if (++stage == p.stages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer
    const uint32_t hi = desc_hi_word(row_bytes);
    const uint32_t base_lo = smem_u32(s.base) >> 4;
    const uint32_t stage_lo = stage_bytes >> 4;
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
      const int acc = it & (p.acc_stages - 1);
      mbar_wait(&s.tmem_empty[acc], ((it >> p.acc_shift) & 1) ^ 1u);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + acc * p.block_n;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&s.full_bar[stage], phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t a_lo = base_lo + stage * stage_lo;
          const uint32_t b_lo = a_lo + (a_bytes >> 4);
#pragma unroll
          for (int k = 0; k < k_steps; ++k)
            umma_f16(tmem_d, desc64(a_lo + 2 * k, hi), desc64(b_lo + 2 * k, hi), p.idesc, (k > 0 || kb > 0) ? 1u : 0u);
          umma_commit(&s.empty_bar[stage]);
          if (kb == num_kb - 1) umma_commit(&s.tmem_full[acc]);
        }
        __syncwarp();
        if (++stage == p.stages) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  } else {
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
    for (int tile = blockIdx.x + group * gridDim.x; tile < p.num_tiles; tile += 2 * gridDim.x, it += 2) {
      const int acc = it & (p.acc_stages - 1);
      const TileCoord t = decode_tile(p, tile);
      const int x = t.x0 + xx, y = t.y0 + yy, b = t.b0 + bi;
      const bool valid = (x < p.m_w) && (y < p.m_h) && (b < p.m_b);
      uint32_t s_dm = 0, s_aux = 0;
      const float* g_dm = nullptr;
      const int tab_n = p.tile_b * p.block_n;
      if ((p.demod != nullptr && p.smem_demod) || p.smem_aux) {
        float* tab = s.demod + group * kDemodTable;
        epi_group_sync(group);  // previous tile of this group fully drained
        for (int i = gt; i < tab_n; i += 128) {
          const int bb = t.b0 + i / p.block_n;
          const long long ch = (long long)bb * p.cout + t.n0 + (i % p.block_n);
          const bool in = bb < p.m_b;
          if (p.smem_demod) tab[i] = in ? __ldg(p.demod + ch) * p.act_gain : 0.f;
          if (p.smem_aux) {
            tab[tab_n + i] = (in && p.out_scale != nullptr) ? __ldg(p.out_scale + ch) : 1.f;
#pragma unroll
            for (int o = 0; o < 3; ++o)
              tab[(2 + o) * tab_n + i] =
                  (in && p.rgb_w != nullptr) ? __ldg(p.rgb_w + ((long long)bb * 3 + o) * p.cout + t.n0 + (i % p.block_n)) : 0.f;
          }
        }
        epi_group_sync(group);
        if (p.smem_demod) s_dm = smem_u32(tab + bi * p.block_n);
        if (p.smem_aux) s_aux = smem_u32(tab + tab_n + bi * p.block_n);
      }
      if (p.demod != nullptr && !p.smem_demod) g_dm = p.demod + (long long)b * p.cout + t.n0;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * p.block_n;
      const uint32_t s_bias = (p.cout <= kMaxBias) ? smem_u32(s.bias + t.n0) : 0u;  // wide layers: global, act == 0
      epilogue_dispatch(p, taddr, &s.tmem_full[acc], (it >> p.acc_shift) & 1, x, y, b, t.n0, valid, gain, s_bias, s_dm,
                        g_dm, s_aux, tab_n);
      tc_fence_before();
      mbar_arrive(&s.tmem_empty[acc]);
    }
  }
  kernel_epilogue(p, tmem_base, warp);
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

template <int kBlockK>
__global__ void __launch_bounds__(kThreads, 1) conv_row_kernel(const __grid_constant__ ConvParams p) {
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
    for (int item = blockIdx.x; item < p.row_items; item += gridDim.x) {
      const RowItem w = decode_item(p, item);
      const int cx = w.seg * 128 - 1;
      for (int r = 0; r < w.rows_out + 2; ++r) {
        for (int kc = 0; kc < kc_n; ++kc) {
          mbar_wait(&s.empty_bar[slot], phase ^ 1u);
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
    // ---------------- MMA issuer
    mbar_wait(s.w_bar, 0);
    tc_fence_after();
    const uint32_t hi = desc_hi_word(row_bytes);
    const uint32_t w_lo = smem_u32(s_w) >> 4;
    const uint32_t ring_lo = smem_u32(s_ring) >> 4;
    const uint32_t wtile_lo = wtile_bytes >> 4;
    constexpr uint32_t slot_lo = slot_bytes >> 4;
    constexpr uint32_t px_lo = row_bytes >> 4;  // one pixel (one smem row) in descriptor units
    const int ring = p.acc_stages;               // accumulator slots (power of two)
    int slot = 0;                                // ring slot of (current input row, kc 0)
    uint32_t phase = 0;
    int it_base = 0;                             // output rows issued before this item
    for (int item = blockIdx.x; item < p.row_items; item += gridDim.x) {
      const RowItem w = decode_item(p, item);
      for (int i = 0; i < w.rows_out + 2; ++i) {
        const int row_slot = slot;
        for (int kc = 0; kc < kc_n; ++kc) {  // this input row has landed (all its k-chunks)
          mbar_wait(&s.full_bar[slot], phase);
          if (++slot == nslots) {
            slot = 0;
            phase ^= 1u;
          }
        }
        const bool fresh = i < w.rows_out;  // output row i receives its first contribution (kh = 0) from this input row
        if (fresh) {
          const int it = it_base + i;
          mbar_wait(&s.tmem_empty[it & (ring - 1)], ((it >> p.acc_shift) & 1) ^ 1u);
        }
        tc_fence_after();
        if (elect_one()) {
          const int j_lo = max(i - 2, 0), j_hi = min(i, w.rows_out - 1);
          // one accumulating run over output rows [ja, jb] (split where the accumulator ring wraps)
          auto run = [&](int ja, int jb, uint64_t a_desc, uint32_t tile_lo, uint32_t accumulate) {
            const int n = jb - ja + 1;
            const int sa = (it_base + ja) & (ring - 1);
            const int n1 = min(n, ring - sa);
            const uint32_t b_lo = tile_lo + (ja - (i - 2)) * wtile_lo;
            umma_f16(tmem_base + sa * p.block_n, a_desc, desc64(b_lo, hi), p.idesc_n[n1 - 1], accumulate);
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
    int it = 0;
    for (int item = blockIdx.x; item < p.row_items; item += gridDim.x) {
      const RowItem w = decode_item(p, item);
      const int x = w.seg * 128 + row;
      const bool valid = x < p.m_w;
      uint32_t s_dm = 0, s_aux = 0;
      if (p.demod != nullptr || p.smem_aux) {  // one image per item: each group stages that image's table rows once
        float* tab = s.demod + group * kDemodTable;
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
        if ((it & 1) != group) continue;
        const int acc = it & (p.acc_stages - 1);
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * p.block_n;
        epilogue_dispatch(p, taddr, &s.tmem_full[acc], (it >> p.acc_shift) & 1, x, w.y0 + j, w.b, 0, valid, gain,
                          smem_u32(s.bias), s_dm, nullptr, s_aux, p.block_n);
        tc_fence_before();
        mbar_arrive(&s.tmem_empty[acc]);
      }
    }
  }
  kernel_epilogue(p, tmem_base, warp);
}

}  // namespace b200ir
