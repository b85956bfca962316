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

#define B200IR_REQUIRE(cond, ...)    \
  do {                               \
    if (!(cond)) {                   \
      b200ir::set_error(__VA_ARGS__); \
      return 1;                      \
    }                                \
  } while (0)

}  // namespace b200ir
