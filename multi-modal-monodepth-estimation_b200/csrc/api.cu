// Library-wide entry points: version and thread-local error text.
#include <stdarg.h>
#include <string.h>
#include "common.cuh"
#include "../../include/b200swin.h"

namespace b200swin {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace b200swin

extern "C" int b200swin_version(void) { return 100; }
extern "C" const char* b200swin_last_error(void) { return b200swin::g_err; }
