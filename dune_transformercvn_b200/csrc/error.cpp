// Thread-local error text behind tcvn_last_error(); ABI version.
#include <stdarg.h>
#include <atomic>
#include <stdio.h>

#include "../../include/tcvn.h"

namespace tcvn {
static thread_local char g_error[512] = "";
static std::atomic<long long> g_launches{0};

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
  return code;
}
}  // namespace tcvn

extern "C" long long tcvn_launch_count(void) { return tcvn::g_launches.load(std::memory_order_relaxed); }
extern "C" int tcvn_abi_version(void) { return TCVN_ABI_VERSION; }
extern "C" const char* tcvn_last_error(void) { return tcvn::g_error; }
