// ABI version + thread-local error text shared by all translation units.
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
}  // namespace egs

extern "C" int egs_abi_version(void) { return EGS_ABI_VERSION; }
extern "C" const char* egs_last_error_string(void) { return egs::error_buffer(); }
