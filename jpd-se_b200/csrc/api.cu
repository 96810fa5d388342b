// C-ABI bookkeeping: version + thread-local error message.
#include <cstdarg>
#include <cstdio>

#include "common.cuh"

namespace jpdse {

char* last_error_buffer() {
  static thread_local char buf[512] = {0};
  return buf;
}

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(last_error_buffer(), 512, fmt, ap);
  va_end(ap);
  return code;
}

}  // namespace jpdse

extern "C" int jpdse_abi_version(void) { return 4; }
extern "C" const char* jpdse_last_error(void) { return jpdse::last_error_buffer(); }
