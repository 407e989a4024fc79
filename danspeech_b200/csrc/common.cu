#include "common.cuh"

namespace dsb {
thread_local char g_err[512] = {0};
std::atomic<uint64_t> g_launches{0};

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
}  // namespace dsb

extern "C" const char* dsb_last_error(void) { return dsb::g_err; }
extern "C" int dsb_abi_version(void) { return 1; }
extern "C" uint64_t dsb_kernel_launch_count(void) { return dsb::g_launches.load(); }
