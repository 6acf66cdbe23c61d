// Library-level entry points of the C ABI: version and the thread-local error text.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace b2s {
static thread_local char g_error[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}
}  // namespace b2s

extern "C" {
int b2s_version(void) { return B2S_VERSION; }
const char* b2s_last_error(void) { return b2s::g_error; }
}
