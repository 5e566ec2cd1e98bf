// Thread-local error text behind tcvn_last_error(); ABI version.
#include <stdarg.h>
#include <atomic>
#include <stdint.h>
#include <stdio.h>

#include "../../include/tcvn.h"

namespace tcvn {
static thread_local char g_error[512] = "";
static std::atomic<long long> g_launches{0};
static thread_local const unsigned long long* g_seed_offset = nullptr;

const unsigned long long* seed_offset_ptr() { return g_seed_offset; }

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

// Device-resident seed offset for the counter-hash random streams (dropout masks, pixel noise): every launch made from
// THIS host thread until the next call adds *device_ptr to the seed it was given.  A captured CUDA graph of the training
// step bakes kernel arguments in; with the step counter in device memory a replay still draws fresh masks (the forward
// and the backward of one step read the same value, so their masks agree).  NULL switches it off.
extern "C" int tcvn_set_seed_offset(const uint64_t* device_ptr) {
  tcvn::g_seed_offset = reinterpret_cast<const unsigned long long*>(device_ptr);
  return TCVN_OK;
}
