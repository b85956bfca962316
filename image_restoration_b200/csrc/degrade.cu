// Fused per-crop degradation: pyblur blur -> cv2.resize(INTER_LINEAR) down -> + Gaussian noise -> clip ->
// cv2.resize(INTER_LINEAR) up -> clamp/round to the 8-bit grid -> normalize to [-1,1] -> NCHW fp32.
// One CTA per crop; the blur is evaluated only at the (separable) set of source rows/cols the down-resize samples.
//
// Reference semantics (see include/b200ir.h for file:line):
//   blur   : scipy.signal.convolve2d(img_f32[:,:,c], K, mode='same', boundary='fill', fillvalue=255).astype(uint8)
//            out[y,x] = sum_{i,j} K[i,j] * img[y + ci - i, x + cj - j], accumulated in scipy's summation tree and in the
//            type scipy computes in (blur_taps.cuh: float64 for float64 kernels, float32 for float32 ones), then
//            truncation to uint8 -- bit-identical -- then /255 in fp32 (degradations.py:363-366)
//   resize : cv2.resize(INTER_LINEAR) on float32 as the build container's OpenCV runs it (Intel IPP): float64 source
//            coordinate, fp32 fraction, fma(S1 - S0, f, S0), horizontal pass then vertical pass
//   noise  : clip(img + noise, 0, 1) (degradations.py:660-669; noise already scaled by sigma/255)
//   tail   : clamp(round(x*255), 0, 255)/255 then (x-0.5)/0.5 (ffhq_degradation_dataset.py:307-311), BGR->RGB
#include "blur_taps.cuh"
#include "host_common.h"

namespace b200ir {

static constexpr int kDegThreads = 512;

struct ResizeAxis {
  int i0, i1;
  float w0, w1;
};

// shared-memory layout (every region 16-byte aligned):
//   non-zero taps (kmax^2) | row taps (lr_hmax) | col taps (lr_wmax) | up-resize row taps (H) | up-resize col taps (W) |
//   sampled blur (2lh*2lw*3 u8) | GT u8, later reused for the LR image (lh*lw*3 f32): the GT crop is dead once the
//   blur samples exist
struct DegLayout {
  size_t row, col, urow, ucol, lr, samp, gt;
};
__host__ __device__ inline size_t align16(size_t x) { return (x + 15) & ~(size_t)15; }
__host__ __device__ inline DegLayout deg_layout(int kmax, int lr_wmax, int lr_hmax, int H, int W) {
  DegLayout l;
  l.row = align16((size_t)kmax * kmax * sizeof(DfTap));
  l.col = l.row + (size_t)lr_hmax * sizeof(ResizeAxis);
  l.urow = l.col + (size_t)lr_wmax * sizeof(ResizeAxis);
  l.ucol = l.urow + (size_t)H * sizeof(ResizeAxis);
  l.samp = align16(l.ucol + (size_t)W * sizeof(ResizeAxis));
  l.gt = align16(l.samp + (size_t)2 * lr_hmax * 2 * lr_wmax * 3);
  l.lr = l.gt;  // aliases the staged GT crop
  return l;
}

// cv2.resize(INTER_LINEAR) tap for destination index d, resize from `src_n` to `dst_n` samples, AS EXECUTED for float
// images in the build container (opencv-python 4.13 dispatches to Intel IPP): source coordinate (d + 0.5) * src/dst - 0.5
// in float64, fraction rounded to fp32, indices clamped; each pass is fma(S1 - S0, f, S0), horizontal pass first.
// Restated and pinned bit-exactly against cv2.resize in oracle/degrade_full_oracle.py::resize_linear.
// (OpenCV's own C++ path rounds the coordinate to fp32 before taking the fraction and uses S0*(1-f) + S1*f: it differs
// from IPP by up to 1.3e-5 on noise images.)
__device__ __forceinline__ ResizeAxis cv_linear_tap(int d, int src_n, int dst_n) {
  const double scale = (double)src_n / (double)dst_n;
  const double f = (d + 0.5) * scale - 0.5;
  const double fl = floor(f);
  const int s = (int)fl;
  ResizeAxis a;
  a.i0 = min(max(s, 0), src_n - 1);
  a.i1 = min(max(s + 1, 0), src_n - 1);
  a.w1 = (float)(f - fl);
  a.w0 = 0.f;  // unused
  return a;
}
__device__ __forceinline__ float cv_lerp(float s0, float s1, float f) { return __fmaf_rn(__fsub_rn(s1, s0), f, s0); }

__device__ __forceinline__ uint8_t trunc_u8(double s) {
  s = fmin(fmax(s, 0.0), 255.0);
  return (uint8_t)s;  // truncation toward zero, as ndarray.astype(uint8)
}

// pyblur blur of the three channels at (y, x) in the reference's arithmetic type; raw sums returned as double (exact
// for both types)
__device__ __forceinline__ void blur3(const uint8_t* __restrict__ img, int H, int W, int y, int x, int rad, int f64,
                                      const DfTap* __restrict__ nz, int n_nz, double (&s)[3]) {
  if (f64) {
    df_pyblur_raw<double>(img, H, W, y, x, rad, nz, n_nz, s);
  } else {
    float t[3];
    df_pyblur_raw<float>(img, H, W, y, x, rad, nz, n_nz, t);
    s[0] = t[0];
    s[1] = t[1];
    s[2] = t[2];
  }
}

template <bool kStage>
__global__ void __launch_bounds__(kDegThreads, 1)
degrade_kernel(const uint8_t* __restrict__ gt, const double* __restrict__ taps_all, const int* __restrict__ ksize,
               const int* __restrict__ taps_f64, int kmax, const int* __restrict__ lr_w, const int* __restrict__ lr_h, const float* __restrict__ noise,
               int lr_wmax, int lr_hmax, float* __restrict__ out, int H, int W, int bgr2rgb) {
  extern __shared__ __align__(16) uint8_t smem[];
  __shared__ int s_nnz;
  const int b = blockIdx.x;
  const int tid = threadIdx.x;
  const int lw = lr_w[b], lh = lr_h[b];
  const int ksz = ksize[b];
  const DegLayout lay = deg_layout(kmax, lr_wmax, lr_hmax, H, W);
  DfTap* s_nz = reinterpret_cast<DfTap*>(smem);
  ResizeAxis* s_row = reinterpret_cast<ResizeAxis*>(smem + lay.row);
  ResizeAxis* s_col = reinterpret_cast<ResizeAxis*>(smem + lay.col);
  ResizeAxis* s_urow = reinterpret_cast<ResizeAxis*>(smem + lay.urow);
  ResizeAxis* s_ucol = reinterpret_cast<ResizeAxis*>(smem + lay.ucol);
  float* s_lr = reinterpret_cast<float*>(smem + lay.lr);
  uint8_t* s_samp = smem + lay.samp;
  uint8_t* s_gt = smem + lay.gt;
  const uint8_t* g_img = gt + (size_t)b * H * W * 3;

  // non-zero taps with their summation groups, compacted by one warp (blur_taps.cuh)
  if (tid < 32) {
    const int count = df_compact_taps(taps_all + (size_t)b * kmax * kmax, kmax, ksz, ksz > 0 ? 1 : 0, s_nz, tid);
    if (tid == 0) s_nnz = count;
  }
  for (int i = tid; i < lh; i += kDegThreads) s_row[i] = cv_linear_tap(i, H, lh);
  for (int i = tid; i < lw; i += kDegThreads) s_col[i] = cv_linear_tap(i, W, lw);
  for (int i = tid; i < H; i += kDegThreads) s_urow[i] = cv_linear_tap(i, lh, H);
  for (int i = tid; i < W; i += kDegThreads) s_ucol[i] = cv_linear_tap(i, lw, W);
  if (kStage) {
    const int n16 = (H * W * 3) >> 4;  // H*W*3 is a multiple of 16 whenever W % 16 == 0 (checked on the host)
    const uint4* src = reinterpret_cast<const uint4*>(g_img);
    uint4* dst = reinterpret_cast<uint4*>(s_gt);
    for (int i = tid; i < n16; i += kDegThreads) dst[i] = __ldg(src + i);
  }
  __syncthreads();
  const uint8_t* img = kStage ? s_gt : g_img;
  const int n_nz = s_nnz;
  const int rad = (ksz - 1) >> 1;
  const int f64 = taps_f64 != nullptr ? taps_f64[b] : 0;

  // 1a. blurred uint8 samples on the separable grid {row i0/i1} x {col i0/i1}: one thread per sample, three channels
  const int nr = 2 * lh, nc = 2 * lw;
  for (int it = tid; it < nr * nc; it += kDegThreads) {
    const int q = it % nc;
    const int r = it / nc;
    const int y = (r & 1) ? s_row[r >> 1].i1 : s_row[r >> 1].i0;
    const int x = (q & 1) ? s_col[q >> 1].i1 : s_col[q >> 1].i0;
    uint8_t* d = s_samp + it * 3;
    if (ksz > 0) {
      double s3[3];
      blur3(img, H, W, y, x, rad, f64, s_nz, n_nz, s3);
      d[0] = trunc_u8(s3[0]);
      d[1] = trunc_u8(s3[1]);
      d[2] = trunc_u8(s3[2]);
    } else {
      const uint8_t* px = img + (y * W + x) * 3;
      d[0] = px[0];
      d[1] = px[1];
      d[2] = px[2];
    }
  }
  __syncthreads();
  // 1b. down-resize (horizontal then vertical), add noise, clip
  for (int it = tid; it < lh * lw * 3; it += kDegThreads) {
    const int c = it % 3;
    const int lx = (it / 3) % lw;
    const int ly = it / (3 * lw);
    const ResizeAxis ry = s_row[ly], rx = s_col[lx];
    const float p00 = __fdiv_rn((float)s_samp[((2 * ly) * nc + 2 * lx) * 3 + c], 255.f);
    const float p01 = __fdiv_rn((float)s_samp[((2 * ly) * nc + 2 * lx + 1) * 3 + c], 255.f);
    const float p10 = __fdiv_rn((float)s_samp[((2 * ly + 1) * nc + 2 * lx) * 3 + c], 255.f);
    const float p11 = __fdiv_rn((float)s_samp[((2 * ly + 1) * nc + 2 * lx + 1) * 3 + c], 255.f);
    float v = cv_lerp(cv_lerp(p00, p01, rx.w1), cv_lerp(p10, p11, rx.w1), ry.w1);
    if (noise != nullptr) v = __fadd_rn(v, noise[(((size_t)b * lr_hmax + ly) * lr_wmax + lx) * 3 + c]);
    s_lr[it] = fminf(fmaxf(v, 0.f), 1.f);
  }
  __syncthreads();
  // 2. up-resize to (H, W), 8-bit grid, normalize, NCHW (optionally BGR->RGB)
  float* o = out + (size_t)b * 3 * H * W;
  for (int it = tid; it < H * W; it += kDegThreads) {
    const int x = it % W, y = it / W;
    const ResizeAxis ry = s_urow[y], rx = s_ucol[x];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float p00 = s_lr[(ry.i0 * lw + rx.i0) * 3 + c], p01 = s_lr[(ry.i0 * lw + rx.i1) * 3 + c];
      const float p10 = s_lr[(ry.i1 * lw + rx.i0) * 3 + c], p11 = s_lr[(ry.i1 * lw + rx.i1) * 3 + c];
      float v = cv_lerp(cv_lerp(p00, p01, rx.w1), cv_lerp(p10, p11, rx.w1), ry.w1);
      v = fminf(fmaxf(v, 0.f), 1.f);
      v = fminf(fmaxf(rintf(__fmul_rn(v, 255.f)), 0.f), 255.f);
      v = __fdiv_rn(v, 255.f);
      v = __fdiv_rn(__fsub_rn(v, 0.5f), 0.5f);
      const int co = bgr2rgb ? 2 - c : c;
      o[(size_t)co * H * W + it] = v;
    }
  }
}

// Full-resolution blur dump for parity tests (uint8 after truncation and/or fp32 before it).
// parity aid: the full blurred image as pyblur returns it.  grid (ceil(H*W / 256), B); every block compacts the taps of
// its crop again (cheap next to H*W*taps).
__global__ void blur_full_kernel(const uint8_t* __restrict__ gt, const double* __restrict__ taps_all,
                                 const int* __restrict__ ksize, const int* __restrict__ taps_f64, int kmax,
                                 uint8_t* __restrict__ blur_u8, float* __restrict__ blur_f32, int H, int W) {
  extern __shared__ __align__(16) uint8_t smem[];
  __shared__ int s_nnz;
  DfTap* s_nz = reinterpret_cast<DfTap*>(smem);
  const int b = blockIdx.y;
  const int ksz = ksize[b];
  if (threadIdx.x < 32) {
    const int count = df_compact_taps(taps_all + (size_t)b * kmax * kmax, kmax, ksz, ksz > 0 ? 1 : 0, s_nz, threadIdx.x);
    if (threadIdx.x == 0) s_nnz = count;
  }
  __syncthreads();
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= H * W) return;
  const int x = p % W, y = p / W;
  const uint8_t* img = gt + (size_t)b * H * W * 3;
  double s3[3];
  if (ksz > 0) {
    blur3(img, H, W, y, x, (ksz - 1) >> 1, taps_f64 != nullptr ? taps_f64[b] : 0, s_nz, s_nnz, s3);
  } else {
    for (int c = 0; c < 3; ++c) s3[c] = (double)img[(size_t)p * 3 + c];
  }
  const size_t o = ((size_t)b * H * W + p) * 3;
  for (int c = 0; c < 3; ++c) {
    if (blur_f32) blur_f32[o + c] = (float)s3[c];
    if (blur_u8) blur_u8[o + c] = trunc_u8(s3[c]);
  }
}

}  // namespace b200ir

using namespace b200ir;

extern "C" int b200ir_degrade(const uint8_t* gt, const double* taps, const int32_t* ksize, const int32_t* taps_f64,
                              int kmax, const int32_t* lr_w,
                              const int32_t* lr_h, const float* noise, int lr_wmax, int lr_hmax, float* out,
                              uint8_t* blur_u8_out, float* blur_f32_out, int B, int H, int W, int bgr2rgb,
                              void* stream) {
  B200IR_REQUIRE(gt && taps && ksize && lr_w && lr_h && out, "degrade: null pointer");
  B200IR_REQUIRE(B > 0 && H > 0 && W > 0 && kmax > 0 && (kmax & 1) && lr_wmax > 0 && lr_hmax > 0,
                 "degrade: bad sizes (kmax must be odd)");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  int dev = 0, smem_optin = 0;
  if (cudaGetDevice(&dev) != cudaSuccess ||
      cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess) {
    set_error("degrade: no CUDA device");
    return 1;
  }
  const size_t lr_bytes = (size_t)lr_hmax * lr_wmax * 3 * 4;
  const size_t gt_off = deg_layout(kmax, lr_wmax, lr_hmax, H, W).gt;
  const size_t base = gt_off + lr_bytes;  // without staging the tail only holds the LR image
  const size_t staged = gt_off + ((size_t)H * W * 3 > lr_bytes ? (size_t)H * W * 3 : lr_bytes);
  const bool stage = (staged <= (size_t)smem_optin) && ((H * W * 3) % 16 == 0) &&
                     ((reinterpret_cast<uintptr_t>(gt) & 15) == 0);
  B200IR_REQUIRE(base <= (size_t)smem_optin, "degrade: low-resolution image %dx%d does not fit shared memory", lr_wmax,
                 lr_hmax);
  if (stage) {
    cudaFuncSetAttribute(degrade_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)staged);
    degrade_kernel<true><<<B, kDegThreads, staged, st>>>(gt, taps, ksize, taps_f64, kmax, lr_w, lr_h, noise, lr_wmax, lr_hmax,
                                                        out, H, W, bgr2rgb);
  } else {
    cudaFuncSetAttribute(degrade_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)base);
    degrade_kernel<false><<<B, kDegThreads, base, st>>>(gt, taps, ksize, taps_f64, kmax, lr_w, lr_h, noise, lr_wmax, lr_hmax,
                                                        out, H, W, bgr2rgb);
  }
  if (check_launch("degrade")) return 1;
  if (blur_u8_out || blur_f32_out) {
    const size_t tap_bytes = (size_t)kmax * kmax * sizeof(DfTap);
    cudaFuncSetAttribute(blur_full_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tap_bytes);
    blur_full_kernel<<<dim3((unsigned)((H * W + 255) / 256), (unsigned)B), 256, tap_bytes, st>>>(
        gt, taps, ksize, taps_f64, kmax, blur_u8_out, blur_f32_out, H, W);
    if (check_launch("degrade(blur dump)")) return 1;
  }
  return 0;
}
