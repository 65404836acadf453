// api.cu — library-wide state of libpcoe.so: thread-local error text, launch counter, version.
#include "common.cuh"
#include <atomic>
#include <string.h>

namespace pcoe {

static thread_local char g_error[512] = "";
static std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof g_error, fmt, ap);
  va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

}  // namespace pcoe

extern "C" int pcoe_version(void) { return PCOE_VERSION; }
extern "C" const char* pcoe_last_error(void) { return pcoe::g_error; }
extern "C" uint64_t pcoe_launch_count(void) { return pcoe::g_launches.load(std::memory_order_relaxed); }
