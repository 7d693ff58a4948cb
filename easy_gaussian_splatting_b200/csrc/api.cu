// ABI version + thread-local error text shared by all translation units.
#include <atomic>

#include "egs_common.cuh"

namespace egs {
char* error_buffer() {
  static thread_local char buf[512] = {0};
  return buf;
}
int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(error_buffer(), 512, fmt, ap);
  va_end(ap);
  return code;
}
static std::atomic<long long> g_kernel_launches{0};
void note_kernel_launches(int n) { g_kernel_launches.fetch_add(n, std::memory_order_relaxed); }
}  // namespace egs

extern "C" int egs_abi_version(void) { return EGS_ABI_VERSION; }
extern "C" const char* egs_last_error_string(void) { return egs::error_buffer(); }
extern "C" int64_t egs_kernel_launch_count(void) { return (int64_t)egs::g_kernel_launches.load(std::memory_order_relaxed); }
