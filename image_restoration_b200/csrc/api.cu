// C-ABI glue shared by all translation units: error string, launch counter, version, device check.
#include <stdarg.h>

#include "host_common.h"

namespace b200ir {

static thread_local char g_err[512] = "";
std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int device_check_impl();

}  // namespace b200ir

extern "C" const char* b200ir_last_error(void) { return b200ir::g_err; }
extern "C" int b200ir_abi_version(void) { return B200IR_ABI_VERSION; }
extern "C" uint64_t b200ir_launch_count(void) { return b200ir::g_launches.load(); }
extern "C" int b200ir_device_check(void) { return b200ir::device_check_impl(); }
