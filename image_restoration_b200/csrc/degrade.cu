// Fused per-crop degradation kernel (pyblur blur -> resize down -> noise -> clip -> resize up -> round -> normalize).
#include "host_common.h"

using namespace b200ir;

extern "C" int b200ir_degrade(const uint8_t* gt, const float* taps, const int32_t* ksize, int kmax, const int32_t* lr_w,
                              const int32_t* lr_h, const float* noise, int lr_wmax, int lr_hmax, float* out,
                              uint8_t* blur_out, int B, int H, int W, int bgr2rgb, void* stream) {
  set_error("degrade: not built yet");
  return 1;
}
