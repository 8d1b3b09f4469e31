// Error plumbing and version of the C-ABI (include/km_b200.h).
#include "km_common.cuh"
#include <stdarg.h>
#include <string.h>

static thread_local char g_err[512] = "";

void km_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" const char* km_last_error(void) { return g_err; }
extern "C" int km_version(void) { return 100; }
