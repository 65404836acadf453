// common.cuh — shared helpers of libpcoe (error reporting, launch accounting, small device utils).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include "pcoe.h"

namespace pcoe {

// thread-local error text + process-wide launch counter (defined in api.cu)
void set_error(const char* fmt, ...);
void count_launch(int n = 1);

inline int fail(pcoe_status st, const char* fmt, ...) {
  char buf[480];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  set_error("%s", buf);
  return (int)st;
}

// Optional per-kernel CUDA-event timing (pcoe_profile_enable): begin/end events around a launch.
int profile_begin(const char* what, cudaStream_t st);   // returns slot or -1 when disabled
void profile_end(int slot, cudaStream_t st);
bool profile_enabled();

// One auxiliary stream + a few events per device (created on first use, never destroyed): lets one C-ABI call run
// independent kernels side by side (fork: record on the caller's stream, wait on aux; join: the reverse).  Works
// under stream capture (the dependencies become graph edges).  Disabled while per-kernel profiling is on, so that
// the CUDA-event timings stay those of kernels running alone.
struct AuxStream { cudaStream_t s; cudaEvent_t ev[4]; };
AuxStream* aux_stream();   // nullptr when unavailable / profiling
// The auxiliary stream and its events are per-device state shared by every caller: a fork/join sequence is enqueued
// under this lock (host-side only, microseconds), so concurrent host threads cannot interleave their event
// record / wait pairs.
void aux_lock();
void aux_unlock();
struct AuxGuard {
  bool held;
  explicit AuxGuard(bool take) : held(take) { if (held) aux_lock(); }
  ~AuxGuard() { if (held) aux_unlock(); }
  AuxGuard(const AuxGuard&) = delete;
  AuxGuard& operator=(const AuxGuard&) = delete;
};

// Checks the launch that was just enqueued (cudaPeekAtLastError is legal during graph capture).
inline int check_launch(const char* what) {
  count_launch();
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    cudaGetLastError();  // clear
    return fail(PCOE_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
  }
  return PCOE_OK;
}

// Brackets one kernel launch: `LaunchScope ls("name", st); kernel<<<...>>>(...); return ls.done();`
struct LaunchScope {
  const char* what;
  cudaStream_t st;
  int slot;
  LaunchScope(const char* w, cudaStream_t s) : what(w), st(s), slot(profile_begin(w, s)) {}
  int done() {
    if (slot >= 0) profile_end(slot, st);
    return check_launch(what);
  }
};

#define PCOE_CUDA(call)                                                                  \
  do {                                                                                   \
    cudaError_t e_ = (call);                                                             \
    if (e_ != cudaSuccess)                                                               \
      return ::pcoe::fail(PCOE_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e_));       \
  } while (0)

#define PCOE_TRY(expr)        \
  do {                        \
    int rc_ = (expr);         \
    if (rc_ != PCOE_OK) return rc_; \
  } while (0)

constexpr int kNumSMs = 148;  // B200

__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
__host__ __device__ inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Squared distance with the operation order and roundings of torch.sum((a - b) ** 2, -1):
// three rounded subtractions, three rounded squares, ((x+y)+z).  No FMA contraction.
__device__ __forceinline__ float sqdist_rn(float ax, float ay, float az, float bx, float by,
                                           float bz) {
  float dx = __fsub_rn(ax, bx), dy = __fsub_rn(ay, by), dz = __fsub_rn(az, bz);
  return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

}  // namespace pcoe
