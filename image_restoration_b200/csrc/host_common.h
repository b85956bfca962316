// Host-side helpers shared by the C-ABI translation units: error string, launch counter, checks.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

#include "../../include/b200ir.h"

namespace b200ir {

void set_error(const char* fmt, ...);
extern std::atomic<uint64_t> g_launches;

inline int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: launch failed: %s", what, cudaGetErrorString(e));
    return 1;
  }
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return 0;
}

// fp16 tiled tensor map (cuTensorMapEncodeTiled through the runtime's driver entry point); returns non-zero + last_error
int encode_map(CUtensorMap* m, const void* ptr, int rank, const cuuint64_t* dims, const cuuint64_t* strides_b,
               const cuuint32_t* box, CUtensorMapSwizzle swz, const char* what);
int num_sms();     // 0 (+ last_error) when the device is not sm_100
int smem_optin();  // max dynamic shared memory per block

// arguments of the TMA-fed streaming FIR kernels (fir_tma.cu)
struct FirLaunch {
  const __half* in;
  int B, Hv, Wv;                 // valid input extent
  long long in_sw, in_sh, in_sb; // element strides of the input buffer
  int C, OH, OW, pad;
  float kscale;                  // per-axis tap scale (taps are kscale*[1,3,3,1])
  __half* out;
  long long out_sy, out_sb;
  bool post;
  const float* noise;
  long long noise_sb;
  const float* noise_gain;
  const float* bias;
  const __half* scale;
  const __half* shift;
  int c_sft;
  const float* s_next;
};
// returns -1 when the shape is not eligible (caller uses the direct kernels), else 0 / 1 like every launcher
int fir_stream_launch(const FirLaunch& a, cudaStream_t st, const char* what);

// TMA-fed streaming resamplers (resample_tma.cu): FIR pad (1,1) + stride 2 (down) / bilinear x2 (up); -1 = not eligible
int resample_stream_launch(bool down, const __half* in, __half* out, int B, int H, int W, int C, cudaStream_t st);

#define B200IR_REQUIRE(cond, ...)    \
  do {                               \
    if (!(cond)) {                   \
      b200ir::set_error(__VA_ARGS__); \
      return 1;                      \
    }                                \
  } while (0)

}  // namespace b200ir
