#include "host_util.h"

#include <cstdarg>
#include <cstdio>

static thread_local char g_err[512] = "";

int agarcl_set_error(int status, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return status;
}

extern "C" const char* agarcl_last_error(void) { return g_err; }
extern "C" const char* agarcl_version(void) { return "agarcl_b200 0.1 (sm_100a)"; }
