// api.cu — library-wide state of libpcoe.so: thread-local error text, launch counter, version.
#include "common.cuh"
#include <atomic>
#include <string.h>
#include <map>
#include <mutex>
#include <string>
#include <vector>

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

// ---- per-kernel timing ------------------------------------------------------------------------
struct ProfRec { const char* what; cudaEvent_t a, b; };
static std::mutex g_prof_mu;
static bool g_prof_on = false;
static std::vector<ProfRec> g_prof;

int profile_begin(const char* what, cudaStream_t st) {
  if (!g_prof_on) return -1;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  ProfRec r{what, nullptr, nullptr};
  if (cudaEventCreate(&r.a) != cudaSuccess || cudaEventCreate(&r.b) != cudaSuccess) return -1;
  cudaEventRecord(r.a, st);
  g_prof.push_back(r);
  return (int)g_prof.size() - 1;
}

void profile_end(int slot, cudaStream_t st) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  if (slot >= 0 && slot < (int)g_prof.size()) cudaEventRecord(g_prof[slot].b, st);
}

bool profile_enabled() { return g_prof_on; }

static std::mutex g_aux_mu;
void aux_lock() { g_aux_mu.lock(); }
void aux_unlock() { g_aux_mu.unlock(); }

AuxStream* aux_stream() {
  static AuxStream tab[64];
  static bool made[64];
  static std::mutex mu;
  if (g_prof_on) return nullptr;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  std::lock_guard<std::mutex> lk(mu);
  if (!made[dev]) {
    AuxStream a{};
    if (cudaStreamCreateWithFlags(&a.s, cudaStreamNonBlocking) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    for (auto& e : a.ev)
      if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    tab[dev] = a;
    made[dev] = true;
  }
  return &tab[dev];
}

}  // namespace pcoe

extern "C" int pcoe_profile_enable(int on) {
  std::lock_guard<std::mutex> lk(pcoe::g_prof_mu);
  pcoe::g_prof_on = on != 0;
  return PCOE_OK;
}

extern "C" int pcoe_profile_report(char* buf, size_t cap) {
  using namespace pcoe;
  if (!buf || cap == 0) return fail(PCOE_ERR_NULL, "profile_report: buffer is NULL");
  std::lock_guard<std::mutex> lk(g_prof_mu);
  std::map<std::string, std::pair<int, double>> agg;
  for (auto& r : g_prof) {
    float ms = 0.f;
    cudaEventSynchronize(r.b);
    if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) {
      auto& e = agg[r.what];
      e.first += 1;
      e.second += ms;
    }
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
  }
  g_prof.clear();
  size_t off = 0;
  buf[0] = 0;
  for (auto& kv : agg) {
    int n = snprintf(buf + off, cap - off, "%s,%d,%.6f\n", kv.first.c_str(), kv.second.first, kv.second.second);
    if (n < 0 || (size_t)n >= cap - off) break;
    off += (size_t)n;
  }
  return PCOE_OK;
}

extern "C" int pcoe_version(void) { return PCOE_VERSION; }
extern "C" const char* pcoe_last_error(void) { return pcoe::g_error; }
extern "C" uint64_t pcoe_launch_count(void) { return pcoe::g_launches.load(std::memory_order_relaxed); }
