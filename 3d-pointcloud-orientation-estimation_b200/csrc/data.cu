// data.cu — device-side input pipeline: the dataset's clouds live in HBM as one ragged fp32 array
// (points [total,3] + offsets [nclouds+1]); one launch builds a training batch by resampling every selected cloud to
// `num` points.  Replaces the per-sample host work of the reference's loaders:
//   sample_pts(arr, num) = arr[np.random.choice(len(arr), num, replace=len(arr) < num)]
//   (dataloader_multi_peak_vonMises.py:21-26, dataloader_8dir_sampled.py:13-15, dataloader_single_peak_vonMises.py:12-14)
// Same distribution, different random stream (numpy's legacy MT19937 permutation cannot be replayed on the device at
// this rate):
//   n >= num: a uniform random subset WITHOUT replacement = the `num` points with the smallest pseudo-random keys
//             key(i) = hi32(mix64(k0 + i)), ties broken by the lower index; found exactly by a 4-pass radix select over the
//             keys (recomputed on the fly, never stored) and written in ascending source order;
//   n <  num: `num` independent uniform draws WITH replacement, i_j = floor(u_j * n), u_j = mix64(k0 + j) / 2^64.
// k0 = mix64(seed ^ mix64(draw)) with draw = base_draw + counter * B + b: a CUDA-graph replay that increments the device
// counter draws fresh subsets.  Integer work end to end: oracle/data.py reproduces the indices bit for bit.
#include "common.cuh"

namespace pcoe {

__host__ __device__ __forceinline__ uint64_t mix64(uint64_t z) {   // splitmix64 finaliser
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

constexpr int kRsThreads = 512;
constexpr int kRsPer = 4;                      // consecutive points per thread in the compaction sweep

// inclusive block scan of one int per thread (kRsThreads threads); returns the exclusive prefix, *total = block sum
__device__ __forceinline__ int block_exscan(int v, int* warp_sums, int* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int o = __shfl_up_sync(0xFFFFFFFFu, inc, d);
    if (lane >= d) inc += o;
  }
  if (lane == 31) warp_sums[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    int w = lane < kRsThreads / 32 ? warp_sums[lane] : 0;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int o = __shfl_up_sync(0xFFFFFFFFu, w, d);
      if (lane >= d) w += o;
    }
    if (lane < kRsThreads / 32) warp_sums[lane] = w;     // inclusive warp totals
  }
  __syncthreads();
  const int base = warp ? warp_sums[warp - 1] : 0;
  *total = warp_sums[kRsThreads / 32 - 1];
  __syncthreads();
  return base + inc - v;
}

__global__ void __launch_bounds__(kRsThreads)
resample_clouds_kernel(const float* __restrict__ points, const int64_t* __restrict__ offsets,
                       const int32_t* __restrict__ cloud_ids, int num, uint64_t seed, uint64_t base_draw,
                       const uint64_t* __restrict__ counter_dev, float* __restrict__ out_xyz,
                       int32_t* __restrict__ out_idx) {
  __shared__ int hist[256];
  __shared__ int warp_sums[kRsThreads / 32];
  __shared__ uint32_t s_prefix;
  __shared__ int s_need;
  const int b = blockIdx.x, tid = threadIdx.x;
  const int c = cloud_ids[b];
  const int64_t o0 = offsets[c];
  const int n = (int)(offsets[c + 1] - o0);
  const uint64_t draw = base_draw + (counter_dev ? *counter_dev : 0ull) * gridDim.x + (uint64_t)b;
  const uint64_t k0 = mix64(seed ^ mix64(draw));
  const float* src = points + (size_t)o0 * 3;
  float* dst = out_xyz + (size_t)b * num * 3;
  int32_t* dsti = out_idx ? out_idx + (size_t)b * num : nullptr;

  if (n <= 0) {   // empty cloud: the reference would fail later; write zeros / -1
    for (int j = tid; j < num; j += kRsThreads) {
      dst[3 * j] = dst[3 * j + 1] = dst[3 * j + 2] = 0.f;
      if (dsti) dsti[j] = -1;
    }
    return;
  }
  if (n < num) {  // with replacement
    for (int j = tid; j < num; j += kRsThreads) {
      const int i = (int)__umul64hi(mix64(k0 + (uint64_t)j), (uint64_t)n);
      dst[3 * j] = __ldg(src + 3 * i); dst[3 * j + 1] = __ldg(src + 3 * i + 1); dst[3 * j + 2] = __ldg(src + 3 * i + 2);
      if (dsti) dsti[j] = i;
    }
    return;
  }

  // ---- radix select: T = the num-th smallest key; need = how many of the keys equal to T are taken
  if (tid == 0) { s_prefix = 0u; s_need = num; }
  uint32_t mask = 0u;
  for (int pass = 0; pass < 4; ++pass) {
    const int shift = 24 - 8 * pass;
    for (int h = tid; h < 256; h += kRsThreads) hist[h] = 0;
    __syncthreads();
    const uint32_t prefix = s_prefix;
    for (int i = tid; i < n; i += kRsThreads) {
      const uint32_t key = (uint32_t)(mix64(k0 + (uint64_t)i) >> 32);
      if ((key & mask) == prefix) atomicAdd(&hist[(key >> shift) & 255u], 1);
    }
    __syncthreads();
    if (tid < 32) {   // warp 0: 8 bins per lane, find the bin where the running count reaches `need`
      int loc[8], s = 0;
#pragma unroll
      for (int u = 0; u < 8; ++u) { loc[u] = hist[tid * 8 + u]; s += loc[u]; }
      int inc = s;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int o = __shfl_up_sync(0xFFFFFFFFu, inc, d);
        if (tid >= d) inc += o;
      }
      const int need = s_need;
      const int before = inc - s;
      const bool mine = before < need && need <= inc;      // exactly one lane
      if (mine) {
        int run = before;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          if (run < need && need <= run + loc[u]) {
            s_prefix = prefix | ((uint32_t)(tid * 8 + u) << shift);
            s_need = need - run;
            break;
          }
          run += loc[u];
        }
      }
    }
    mask |= 255u << shift;
    __syncthreads();
  }
  const uint32_t T = s_prefix;
  const int need_eq = s_need;

  // ---- ordered compaction: keys < T, plus the first need_eq (lowest index) keys == T
  int out_base = 0, eq_base = 0;
  for (int i0 = 0; i0 < n; i0 += kRsThreads * kRsPer) {
    const int i = i0 + tid * kRsPer;
    bool lt[kRsPer], eq[kRsPer];
    int neq = 0;
#pragma unroll
    for (int u = 0; u < kRsPer; ++u) {
      const uint32_t key = i + u < n ? (uint32_t)(mix64(k0 + (uint64_t)(i + u)) >> 32) : 0xFFFFFFFFu;
      lt[u] = i + u < n && key < T;
      eq[u] = i + u < n && key == T;
      neq += eq[u];
    }
    int eq_tot;
    int eq_rank = eq_base + block_exscan(neq, warp_sums, &eq_tot);
    bool sel[kRsPer];
    int nsel = 0;
#pragma unroll
    for (int u = 0; u < kRsPer; ++u) {
      sel[u] = lt[u] || (eq[u] && eq_rank < need_eq);
      eq_rank += eq[u];
      nsel += sel[u];
    }
    int sel_tot;
    int pos = out_base + block_exscan(nsel, warp_sums, &sel_tot);
#pragma unroll
    for (int u = 0; u < kRsPer; ++u) {
      if (!sel[u]) continue;
      const int ii = i + u;
      dst[3 * pos] = __ldg(src + 3 * ii); dst[3 * pos + 1] = __ldg(src + 3 * ii + 1); dst[3 * pos + 2] = __ldg(src + 3 * ii + 2);
      if (dsti) dsti[pos] = ii;
      ++pos;
    }
    out_base += sel_tot;
    eq_base += eq_tot;
  }
}

}  // namespace pcoe

using namespace pcoe;

extern "C" int pcoe_resample_clouds_f32(const float* points, const int64_t* offsets, int nclouds, const int32_t* cloud_ids,
                                        int B, int num, uint64_t seed, uint64_t base_draw, const uint64_t* counter_dev,
                                        float* out_xyz, int32_t* out_idx, void* stream) {
  if (B <= 0 || num <= 0 || nclouds <= 0) return fail(PCOE_ERR_BAD_SHAPE, "resample_clouds: B=%d num=%d nclouds=%d", B, num, nclouds);
  if (!points || !offsets || !cloud_ids || !out_xyz) return fail(PCOE_ERR_NULL, "resample_clouds: NULL pointer");
  LaunchScope ls("resample_clouds_kernel", (cudaStream_t)stream);
  resample_clouds_kernel<<<B, kRsThreads, 0, (cudaStream_t)stream>>>(points, offsets, cloud_ids, num, seed, base_draw, counter_dev,
                                                                     out_xyz, out_idx);
  return ls.done();
}
