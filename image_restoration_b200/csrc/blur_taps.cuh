// Blur taps shared by degrade.cu and degrade_full.cu: the compacted list of non-zero taps of one crop's kernel, carrying
// the summation tree of the reference's library call, and the typed (fp32 / fp64) accumulation over it.
#pragma once
#include <cstdint>

namespace b200ir {

// one non-zero blur tap: weight, source offset, and where it sits in the reference's summation tree
struct DfTap {
  double w;
  short dy, dx;
  int grp;  // (group id << 2) | 2 * last-of-group | first-of-group
};
__device__ __forceinline__ float df_mul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float df_add(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ double df_mul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double df_add(double a, double b) { return __dadd_rn(a, b); }


// Compacts the non-zero taps of the centred ksz x ksz kernel inside its kmax x kmax block (one warp, kernel order) into
// s_nz and marks the first / last tap of every summation group.  mode 1: convolution (flipped offsets), columns grouped
// in fours as scipy.signal.convolve2d sums them; mode 2: correlation, every tap its own group.  Returns the tap count
// (valid in every lane).  Call with the 32 lanes of one warp.
__device__ __forceinline__ int df_compact_taps(const double* __restrict__ tp, int kmax, int ksz, int mode, DfTap* s_nz,
                                               int lane) {
  const int cm = (kmax - 1) >> 1, r = (ksz - 1) >> 1;
  int count = 0;
  const int span = 2 * r + 1;
  const int blocked = span & ~3;  // columns summed in blocks of four by convolve2d
  for (int base = 0; mode != 0 && ksz > 0 && base < span * span; base += 32) {
    const int e = base + lane;
    const int ir = e / span, jr = e % span;
    const int i = cm - r + ir, j = cm - r + jr;
    const double t = (e < span * span) ? tp[i * kmax + j] : 0.0;
    const unsigned m = __ballot_sync(0xffffffffu, t != 0.0);
    if (t != 0.0) {
      DfTap z;
      z.dy = (short)(mode == 1 ? cm - i : i - cm);
      z.dx = (short)(mode == 1 ? cm - j : j - cm);
      z.w = t;
      const int gid = (mode == 1 && jr < blocked) ? ir * 64 + (jr >> 2) : ir * 64 + 32 + jr;
      z.grp = gid << 2;
      s_nz[count + __popc(m & ((1u << lane) - 1u))] = z;
    }
    count += __popc(m);
  }
  __syncwarp();
  for (int base = 0; base < count; base += 32) {
    const int k = base + lane;
    int bits = 0, gid = 0;
    if (k < count) {
      gid = s_nz[k].grp >> 2;
      const int prev = k > 0 ? (s_nz[k - 1].grp >> 2) : -1;
      const int next = k + 1 < count ? (s_nz[k + 1].grp >> 2) : -1;
      bits = (prev != gid ? 1 : 0) | (next != gid ? 2 : 0);
    }
    __syncwarp();
    if (k < count) s_nz[k].grp = (gid << 2) | bits;
    __syncwarp();
  }
  return count;
}

// Blur of the three channels at (y, x), in the arithmetic type T the reference's library call used.
// kMode 1: scipy.signal.convolve2d on the uint8 values, fill 255 outside (pyblur).  scipy 1.18 (this container; pinned
//   1.9.3) walks the kernel rows in ascending order and, inside a row, adds blocks of four columns as
//   ((p0 + p1) + p2) + p3 to the running sum, then the remaining columns one by one -- products and sums rounded
//   separately, in float64 when the kernel is float64 (box / disk / line under NumPy 2) and float32 when it is float32
//   (psf).  The tap list carries that tree (first / last of group), so the result is bit-identical; zero taps add
//   exact zeros and are dropped.
// kMode 2: cv2.filter2D on value/255, BORDER_REFLECT_101: every tap is its own group (plain running sum, fp32).
template <typename T, int kMode, bool kInterior>
__device__ __forceinline__ void df_blur3(const uint8_t* __restrict__ img, const float* __restrict__ lut, int H, int W, int y,
                                         int x, const DfTap* __restrict__ nz, int n_nz, T (&s)[3],
                                         const float* __restrict__ imgf = nullptr) {
  s[0] = s[1] = s[2] = (T)0;
  T g0 = (T)0, g1 = (T)0, g2 = (T)0;
  for (int k = 0; k < n_nz; ++k) {
    const DfTap t = nz[k];
    const T w = (T)t.w;
    int iy = y + t.dy, ix = x + t.dx;
    T v0, v1, v2;
    if (kMode == 1) {
      v0 = v1 = v2 = (T)255;
      if (kInterior || (iy >= 0 && iy < H && ix >= 0 && ix < W)) {
        const uint8_t* px = img + (iy * W + ix) * 3;
        v0 = (T)px[0];
        v1 = (T)px[1];
        v2 = (T)px[2];
      }
    } else {
      if (!kInterior) {
        iy = iy < 0 ? -iy : (iy >= H ? 2 * H - 2 - iy : iy);
        ix = ix < 0 ? -ix : (ix >= W ? 2 * W - 2 - ix : ix);
      }
      if (imgf != nullptr) {  // float GT image (not on the 8-bit grid): filter2D works on the float values
        const float* pf = imgf + (iy * W + ix) * 3;
        v0 = (T)pf[0];
        v1 = (T)pf[1];
        v2 = (T)pf[2];
      } else {
        const uint8_t* px = img + (iy * W + ix) * 3;
        v0 = (T)lut[px[0]];
        v1 = (T)lut[px[1]];
        v2 = (T)lut[px[2]];
      }
    }
    const T p0 = df_mul(w, v0), p1 = df_mul(w, v1), p2 = df_mul(w, v2);
    if (t.grp & 1) {
      g0 = p0;
      g1 = p1;
      g2 = p2;
    } else {
      g0 = df_add(g0, p0);
      g1 = df_add(g1, p1);
      g2 = df_add(g2, p2);
    }
    if (t.grp & 2) {
      s[0] = df_add(s[0], g0);
      s[1] = df_add(s[1], g1);
      s[2] = df_add(s[2], g2);
    }
  }
}

template <typename T, int kMode>
__device__ __forceinline__ void df_blur3_at(const uint8_t* __restrict__ img, const float* __restrict__ lut, int H, int W,
                                            int y, int x, int rad, const DfTap* __restrict__ nz, int n_nz, float (&v)[3],
                                            const float* __restrict__ imgf = nullptr) {
  T s[3];
  if (y >= rad && y + rad < H && x >= rad && x + rad < W) df_blur3<T, kMode, true>(img, lut, H, W, y, x, nz, n_nz, s, imgf);
  else df_blur3<T, kMode, false>(img, lut, H, W, y, x, nz, n_nz, s, imgf);
  if (kMode == 1) {  // .astype(uint8) (truncation; the sum is inside [0, 255] up to rounding), then / 255
#pragma unroll
    for (int c = 0; c < 3; ++c) v[c] = lut[(int)fmin(fmax((double)s[c], 0.0), 255.0)];
  } else {
#pragma unroll
    for (int c = 0; c < 3; ++c) v[c] = (float)s[c];
  }
}

// pyblur blur of one pixel, raw sums (before the uint8 truncation) in the reference's arithmetic type
template <typename T>
__device__ __forceinline__ void df_pyblur_raw(const uint8_t* __restrict__ img, int H, int W, int y, int x, int rad,
                                              const DfTap* __restrict__ nz, int n_nz, T (&s)[3]) {
  if (y >= rad && y + rad < H && x >= rad && x + rad < W) df_blur3<T, 1, true>(img, nullptr, H, W, y, x, nz, n_nz, s);
  else df_blur3<T, 1, false>(img, nullptr, H, W, y, x, nz, n_nz, s);
}


}  // namespace b200ir
