// grouping.cu — neighbour search around sampled centroids.
//   pcoe_knn_f32         k nearest neighbours  (models/base.py:20-35: square_distance + topk)
//   pcoe_ball_query_f32  radius query          (PointNet++Demo.py:49-70)
// Both keep one cloud (SoA) in shared memory per CTA and give one warp to one centroid at a time;
// the (B,S,N) distance matrix of the reference is never materialised.
#include "common.cuh"
#include <math.h>

namespace pcoe {

constexpr int kGroupWarps = 8;         // warps per CTA
constexpr int kCentroidsPerWarp = 4;   // most centroids a warp walks through sequentially
// Centroids per warp of the ball-query kernels: up to kCentroidsPerWarp (the cloud is staged once per CTA), fewer when the
// grid would not fill the GPU - at 32 clouds x 128 centroids of 8192 points, 4 per warp gave 128 CTAs on 148 SMs
// (ncu: 9 % warps active, 121 us); 1 per warp gives 512 CTAs.
static int ball_cpw(int B, int S) {
  const long long tasks = (long long)B * S;
  long long c = tasks / ((long long)kGroupWarps * 2 * kNumSMs);
  return (int)(c < 1 ? 1 : (c > kCentroidsPerWarp ? kCentroidsPerWarp : c));
}

__device__ __forceinline__ void load_cloud_soa(const float* __restrict__ cloud, int N, float* sx,
                                               float* sy, float* sz) {
  // N*3 floats, 16-byte vectorised when the cloud base allows it (always for 4 | 3N and aligned B).
  const int total = N * 3;
  if ((((uintptr_t)cloud) & 15) == 0) {
    // four points = twelve floats = three 16-byte loads per thread and step: no division, all loads of a step in flight
    const float4* c4 = reinterpret_cast<const float4*>(cloud);
    for (int t = threadIdx.x; t < N / 4; t += blockDim.x) {
      const float4 a = __ldg(c4 + 3 * t), b = __ldg(c4 + 3 * t + 1), c = __ldg(c4 + 3 * t + 2);
      const int p = 4 * t;
      if ((N & 3) == 0) {   // sy, sz 16-byte aligned: one conflict-free 16-byte store per coordinate
        *reinterpret_cast<float4*>(sx + p) = make_float4(a.x, a.w, b.z, c.y);
        *reinterpret_cast<float4*>(sy + p) = make_float4(a.y, b.x, b.w, c.z);
        *reinterpret_cast<float4*>(sz + p) = make_float4(a.z, b.y, c.x, c.w);
      } else {
        sx[p] = a.x;     sy[p] = a.y;     sz[p] = a.z;
        sx[p + 1] = a.w; sy[p + 1] = b.x; sz[p + 1] = b.y;
        sx[p + 2] = b.z; sy[p + 2] = b.w; sz[p + 2] = c.x;
        sx[p + 3] = c.y; sy[p + 3] = c.z; sz[p + 3] = c.w;
      }
    }
    for (int i = (N / 4) * 12 + threadIdx.x; i < total; i += blockDim.x) {
      int p = i / 3, c = i - p * 3;
      (c == 0 ? sx : (c == 1 ? sy : sz))[p] = __ldg(cloud + i);
    }
  } else {
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
      int p = i / 3, c = i - p * 3;
      (c == 0 ? sx : (c == 1 ? sy : sz))[p] = __ldg(cloud + i);
    }
  }
}

// kNN by selection, one warp per centroid.  The cloud is walked in chunks of 1024 points: every lane
// holds 32 squared distances of the chunk in registers (point = chunk + slot*32 + lane, so the
// shared-memory reads are conflict free) next to its share of the running K-best list.  The K-th
// smallest distance T of (list U chunk) is found by a 31-step bitwise search on the fp32 bit pattern
// (distances are >= +0, so unsigned order == float order; one REDUX per step), then the new list =
// every entry below T plus the lowest-index entries equal to T.  No sorting, no serial insertion.
// Rows come out in ascending point index (the reference's topk(sorted=False) order is unspecified).
template <int KR>   // list slots per lane, K <= 32*KR
__global__ void __launch_bounds__(1024)
knn_kernel(const float* __restrict__ xyz, const float* __restrict__ new_xyz, int N, int S, int K,
           int32_t* __restrict__ out_idx) {
  extern __shared__ float smem_f[];
  float* sx = smem_f;
  float* sy = sx + N;
  float* sz = sy + N;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nwarps = blockDim.x >> 5;
  uint32_t* list_d = reinterpret_cast<uint32_t*>(sz + N) + warp * (32 * KR);
  int32_t* list_i = reinterpret_cast<int32_t*>(sz + N + nwarps * 32 * KR) + warp * (32 * KR);
  constexpr uint32_t kInf = 0x7F800000u;
  const unsigned lt_mask = (1u << lane) - 1u;

  const int b = blockIdx.y;
  load_cloud_soa(xyz + (size_t)b * N * 3, N, sx, sy, sz);
  __syncthreads();

  // one centroid per warp (high occupancy hides the REDUX / shared-memory latency of the search)
  for (int s = blockIdx.x * nwarps + warp; s < S; s += gridDim.x * nwarps) {
    const float* c = new_xyz + ((size_t)b * S + s) * 3;
    const float cx = __ldg(c), cy = __ldg(c + 1), cz = __ldg(c + 2);
    uint32_t ld[KR];
    int32_t li[KR];
#pragma unroll
    for (int j = 0; j < KR; ++j) { ld[j] = kInf; li[j] = -1; }
    uint32_t kth = kInf;   // current K-th smallest distance (bits); kInf while the list is not full

    for (int base = 0; base < N; base += 1024) {
      uint32_t d[32];
      bool any = false;
#pragma unroll
      for (int q = 0; q < 32; ++q) {
        const int i = base + q * 32 + lane;
        uint32_t v = kInf;
        if (i < N) v = __float_as_uint(sqdist_rn(sx[i], sy[i], sz[i], cx, cy, cz));
        // later chunks: an entry that is not below the current K-th can never enter (ties keep the lower index)
        if (v >= kth) v = kInf;
        d[q] = v;
        any |= v != kInf;
      }
      if (!__any_sync(0xFFFFFFFFu, any)) continue;

      // T = K-th smallest of list U chunk
      uint32_t T = 0;
#pragma unroll 1
      for (int bit = 30; bit >= 0; --bit) {
        const uint32_t cand = T | (1u << bit);
        int cnt = 0;
#pragma unroll
        for (int q = 0; q < 32; ++q) cnt += d[q] < cand;
#pragma unroll
        for (int j = 0; j < KR; ++j) cnt += ld[j] < cand;
        if (__reduce_add_sync(0xFFFFFFFFu, cnt) < K) T = cand;
      }
      int less = 0;
#pragma unroll
      for (int q = 0; q < 32; ++q) less += d[q] < T;
#pragma unroll
      for (int j = 0; j < KR; ++j) less += ld[j] < T;
      const int need_eq = K - __reduce_add_sync(0xFFFFFFFFu, less);   // >= 1

      // rebuild the list in ascending index order: old list entries first (lower indices), then the chunk
      int out_n = 0, eq_n = 0;
#pragma unroll
      for (int j = 0; j < KR; ++j) {
        const bool eq = ld[j] == T && T != kInf;
        const unsigned meq = __ballot_sync(0xFFFFFFFFu, eq);
        const bool sel = ld[j] < T || (eq && eq_n + __popc(meq & lt_mask) < need_eq);
        const unsigned msel = __ballot_sync(0xFFFFFFFFu, sel);
        if (sel) { const int pos = out_n + __popc(msel & lt_mask); list_d[pos] = ld[j]; list_i[pos] = li[j]; }
        out_n += __popc(msel);
        eq_n += __popc(meq);
      }
#pragma unroll
      for (int q = 0; q < 32; ++q) {
        const bool eq = d[q] == T && T != kInf;
        const unsigned meq = __ballot_sync(0xFFFFFFFFu, eq);
        const bool sel = d[q] < T || (eq && eq_n + __popc(meq & lt_mask) < need_eq);
        const unsigned msel = __ballot_sync(0xFFFFFFFFu, sel);
        if (sel) { const int pos = out_n + __popc(msel & lt_mask); list_d[pos] = d[q]; list_i[pos] = base + q * 32 + lane; }
        out_n += __popc(msel);
        eq_n += __popc(meq);
      }
      __syncwarp();
#pragma unroll
      for (int j = 0; j < KR; ++j) {
        const int e = j * 32 + lane;
        ld[j] = e < out_n ? list_d[e] : kInf;
        li[j] = e < out_n ? list_i[e] : -1;
      }
      __syncwarp();
      kth = out_n >= K ? T : kInf;
    }
    int32_t* o = out_idx + ((size_t)b * S + s) * K;
#pragma unroll
    for (int j = 0; j < KR; ++j) {
      const int e = j * 32 + lane;
      if (e < K) o[e] = li[j];
    }
  }
}

// K <= 32: candidate pre-selection.  The maximum over the 32 lanes of each lane's smallest distance bounds the
// K-th smallest distance of the chunk from above (32 elements are <= it), so only the ~10 % of the chunk below
// that bound - compacted in ascending point index into a per-warp shared-memory list together with the running
// K-best list - enter the 31-step bitwise search for the exact K-th value; the search then compares <= kCandPerLane
// registers per lane instead of 33.  Same selection rule and output order as knn_kernel (exact, ties keep the
// lowest indices); a chunk whose candidate set does not fit falls back to searching all 32 registers.
#ifndef KNN_MINB
#define KNN_MINB 4
#endif
constexpr int kCandPerLane = 6;                    // candidate capacity per warp = 32 * kCandPerLane
// WARPS = warps per CTA (8 / 16 / 32 by cloud size).  KNN_MINB = resident 8-warp CTAs per SM the register cap is set
// for.  Measured on B200 (SA1 + SA2 grouping of the bench step): 4 CTAs (64 registers, 124 B of spills) 52.4 us,
// 3 CTAs (80 registers, 36 B) 54.4 us, 2 CTAs (128 registers, no spills) 56.2 us - the kernel is issue-bound
// (ncu: 76 % issue slots active) and occupancy beats spill-freedom.
template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32, WARPS == 8 ? KNN_MINB : (WARPS == 16 ? 2 : 1))
knn_k32_kernel(const float* __restrict__ xyz, const float* __restrict__ new_xyz, int N, int S, int K,
               int32_t* __restrict__ out_idx) {
  extern __shared__ float smem_f[];
  float* sx = smem_f;
  float* sy = sx + N;
  float* sz = sy + N;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nwarps = blockDim.x >> 5;
  constexpr int kCap = 32 * kCandPerLane;
  uint32_t* cand_d = reinterpret_cast<uint32_t*>(sz + N) + warp * kCap;
  int32_t* cand_i = reinterpret_cast<int32_t*>(sz + N + nwarps * kCap) + warp * kCap;
  constexpr uint32_t kInf = 0x7F800000u;
  const unsigned lt_mask = (1u << lane) - 1u;

  const int b = blockIdx.y;
  load_cloud_soa(xyz + (size_t)b * N * 3, N, sx, sy, sz);
  __syncthreads();

  for (int s = blockIdx.x * nwarps + warp; s < S; s += gridDim.x * nwarps) {
    const float* c = new_xyz + ((size_t)b * S + s) * 3;
    const float cx = __ldg(c), cy = __ldg(c + 1), cz = __ldg(c + 2);
    uint32_t ld = kInf;        // running K-best list: one slot per lane, ascending point index
    int32_t li = -1;
    int list_n = 0;
    uint32_t kth = kInf;

    for (int base = 0; base < N; base += 1024) {
      uint32_t d[32];
      uint32_t lane_min = kInf;
#pragma unroll
      for (int q = 0; q < 32; ++q) {
        const int i = base + q * 32 + lane;
        uint32_t v = kInf;
        if (i < N) v = __float_as_uint(sqdist_rn(sx[i], sy[i], sz[i], cx, cy, cz));
        if (v >= kth) v = kInf;      // cannot enter: not below the current K-th (ties keep the lower index)
        d[q] = v;
        lane_min = min(lane_min, v);
      }
      if (__reduce_min_sync(0xFFFFFFFFu, lane_min) == kInf) continue;   // nothing below the current K-th
      // bound: first chunk - 32 elements are <= max(lane minima); later chunks - everything below kth is a candidate
      const uint32_t bound = list_n < K ? __reduce_max_sync(0xFFFFFFFFu, lane_min) : kInf - 1;

      // candidates in ascending index: the old list first, then the chunk
      int cn = 0;
      {
        const bool in = ld != kInf;
        const unsigned m = __ballot_sync(0xFFFFFFFFu, in);
        if (in) { const int pos = __popc(m & lt_mask); cand_d[pos] = ld; cand_i[pos] = li; }
        cn = __popc(m);
      }
#pragma unroll
      for (int q = 0; q < 32; ++q) {
        const bool in = d[q] <= bound && d[q] != kInf;   // kInf = out of range or already excluded
        const unsigned m = __ballot_sync(0xFFFFFFFFu, in);
        const int pos = cn + __popc(m & lt_mask);
        if (in && pos < kCap) { cand_d[pos] = d[q]; cand_i[pos] = base + q * 32 + lane; }
        cn += __popc(m);
      }
      __syncwarp();

      uint32_t T = 0;
      int need_eq;
      if (cn <= kCap) {
        uint32_t cd[kCandPerLane];
#pragma unroll
        for (int j = 0; j < kCandPerLane; ++j) cd[j] = j * 32 + lane < cn ? cand_d[j * 32 + lane] : kInf;
#pragma unroll 1
        for (int bit = 30; bit >= 0; --bit) {
          const uint32_t cand = T | (1u << bit);
          int cnt = 0;
#pragma unroll
          for (int j = 0; j < kCandPerLane; ++j) cnt += cd[j] < cand;
          if (__reduce_add_sync(0xFFFFFFFFu, cnt) < K) T = cand;
        }
        int less = 0;
#pragma unroll
        for (int j = 0; j < kCandPerLane; ++j) less += cd[j] < T;
        need_eq = K - __reduce_add_sync(0xFFFFFFFFu, less);
        // new list = candidates below T + the lowest-index candidates equal to T
        int out_n = 0, eq_n = 0;
        uint32_t nd = kInf;
        int32_t ni = -1;
#pragma unroll
        for (int j = 0; j < kCandPerLane; ++j) {
          const bool eq = cd[j] == T && T != kInf;
          const unsigned meq = __ballot_sync(0xFFFFFFFFu, eq);
          const bool sel = cd[j] < T || (eq && eq_n + __popc(meq & lt_mask) < need_eq);
          const unsigned msel = __ballot_sync(0xFFFFFFFFu, sel);
          // lane `pos` of the new list receives this entry: exchange through the (now consumed) candidate slots
          const int pos = out_n + __popc(msel & lt_mask);
          const int32_t idx = j * 32 + lane < cn ? cand_i[j * 32 + lane] : -1;
          __syncwarp();
          if (sel) { cand_d[pos] = cd[j]; cand_i[pos] = idx; }
          out_n += __popc(msel);
          eq_n += __popc(meq);
        }
        __syncwarp();
        nd = lane < out_n ? cand_d[lane] : kInf;
        ni = lane < out_n ? cand_i[lane] : -1;
        __syncwarp();
        ld = nd; li = ni; list_n = out_n;
      } else {
        // overflow (a very clustered chunk): search all 32 registers + the list, as knn_kernel does
#pragma unroll 1
        for (int bit = 30; bit >= 0; --bit) {
          const uint32_t cand = T | (1u << bit);
          int cnt = ld < cand;
#pragma unroll
          for (int q = 0; q < 32; ++q) cnt += d[q] < cand;
          if (__reduce_add_sync(0xFFFFFFFFu, cnt) < K) T = cand;
        }
        int less = ld < T;
#pragma unroll
        for (int q = 0; q < 32; ++q) less += d[q] < T;
        need_eq = K - __reduce_add_sync(0xFFFFFFFFu, less);
        int out_n = 0, eq_n = 0;
        {
          const bool eq = ld == T && T != kInf;
          const unsigned meq = __ballot_sync(0xFFFFFFFFu, eq);
          const bool sel = ld < T || (eq && __popc(meq & lt_mask) < need_eq);
          const unsigned msel = __ballot_sync(0xFFFFFFFFu, sel);
          if (sel) { const int pos = __popc(msel & lt_mask); cand_d[pos] = ld; cand_i[pos] = li; }
          out_n = __popc(msel);
          eq_n = __popc(meq);
        }
#pragma unroll
        for (int q = 0; q < 32; ++q) {
          const bool eq = d[q] == T && T != kInf;
          const unsigned meq = __ballot_sync(0xFFFFFFFFu, eq);
          const bool sel = d[q] < T || (eq && eq_n + __popc(meq & lt_mask) < need_eq);
          const unsigned msel = __ballot_sync(0xFFFFFFFFu, sel);
          if (sel) { const int pos = out_n + __popc(msel & lt_mask); cand_d[pos] = d[q]; cand_i[pos] = base + q * 32 + lane; }
          out_n += __popc(msel);
          eq_n += __popc(meq);
        }
        __syncwarp();
        ld = lane < out_n ? cand_d[lane] : kInf;
        li = lane < out_n ? cand_i[lane] : -1;
        list_n = out_n;
        __syncwarp();
      }
      kth = list_n >= K ? T : kInf;
    }
    int32_t* o = out_idx + ((size_t)b * S + s) * K;
    if (lane < K) o[lane] = li;
  }
}

// Ball query: ascending-index scan, the first `nsample` hits, padded with the first hit.
__global__ void __launch_bounds__(kGroupWarps * 32)
ball_query_kernel(const float* __restrict__ xyz, const float* __restrict__ new_xyz, int N, int S,
                  int nsample, float r2, int32_t* __restrict__ out_idx, int cpw) {
  extern __shared__ float smem_f[];
  float* sx = smem_f;
  float* sy = sx + N;
  float* sz = sy + N;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = blockIdx.y;
  load_cloud_soa(xyz + (size_t)b * N * 3, N, sx, sy, sz);
  __syncthreads();

  const int s_begin = (blockIdx.x * kGroupWarps + warp) * cpw;
  for (int s = s_begin; s < min(s_begin + cpw, S); ++s) {
    const float* c = new_xyz + ((size_t)b * S + s) * 3;
    const float cx = __ldg(c), cy = __ldg(c + 1), cz = __ldg(c + 2);
    int32_t* o = out_idx + ((size_t)b * S + s) * nsample;
    int cnt = 0, first = N;
    for (int base = 0; base < N && cnt < nsample; base += 32) {
      const int i = base + lane;
      bool hit = false;
      if (i < N) hit = !(sqdist_rn(sx[i], sy[i], sz[i], cx, cy, cz) > r2);  // "> r^2" is outside
      const unsigned m = __ballot_sync(0xFFFFFFFFu, hit);
      if (m) {
        if (cnt == 0) first = base + __ffs(m) - 1;
        const int slot = cnt + __popc(m & ((1u << lane) - 1u));
        if (hit && slot < nsample) o[slot] = i;
        cnt += __popc(m);
      }
    }
    for (int e = min(cnt, nsample) + lane; e < nsample; e += 32) o[e] = first;
  }
}

// square_distance (models/base.py:20-27): dist[b,i,j] = -2 <src_i, dst_j> + |src_i|^2 + |dst_j|^2, the reference's
// expanded form (it can come out slightly negative, as there).  One thread = one src row x 4 consecutive dst columns:
// the (B,N,M) matrix is written once with 16-byte stores, which is all the HBM traffic there is (write-bound).
__global__ void __launch_bounds__(256)
square_distance_kernel(const float* __restrict__ src, const float* __restrict__ dst, int N, int M, int C,
                       float* __restrict__ out) {
  const int b = blockIdx.z, i = blockIdx.y;
  const int j0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (j0 >= M) return;
  const float* s = src + ((size_t)b * N + i) * C;
  const float* d = dst + ((size_t)b * M + j0) * C;
  float dot[4] = {0.f, 0.f, 0.f, 0.f}, nd[4] = {0.f, 0.f, 0.f, 0.f}, ns = 0.f;
  const int nj = min(4, M - j0);
  for (int c = 0; c < C; ++c) {
    const float sv = __ldg(s + c);
    ns = fmaf(sv, sv, ns);
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (u < nj) {
        const float dv = __ldg(d + (size_t)u * C + c);
        dot[u] = fmaf(sv, dv, dot[u]);
        nd[u] = fmaf(dv, dv, nd[u]);
      }
  }
  float r[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) r[u] = (-2.f * dot[u] + ns) + nd[u];
  float* o = out + ((size_t)b * N + i) * M + j0;
  if (nj == 4 && ((((size_t)b * N + i) * M + j0) & 3) == 0) {
    *reinterpret_cast<float4*>(o) = make_float4(r[0], r[1], r[2], r[3]);
  } else {
    for (int u = 0; u < nj; ++u) o[u] = r[u];
  }
}

// Multi-scale grouping: up to kMaxScales radii in ONE pass over the staged cloud.  The squared distance of a point to
// the centroid is computed once and compared with every radius; each scale keeps its own hit count, so the result of
// scale r is exactly what ball_query_kernel gives for (radius_r, nsample_r).  The scan stops when every scale is full.
constexpr int kMaxScales = 4;
struct BallScales {
  int n;
  float r2[kMaxScales];
  int nsample[kMaxScales];
  int32_t* out[kMaxScales];
};
__global__ void __launch_bounds__(kGroupWarps * 32)
ball_query_multi_kernel(const float* __restrict__ xyz, const float* __restrict__ new_xyz, int N, int S, BallScales sc, int cpw) {
  extern __shared__ float smem_f[];
  float* sx = smem_f;
  float* sy = sx + N;
  float* sz = sy + N;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = blockIdx.y;
  load_cloud_soa(xyz + (size_t)b * N * 3, N, sx, sy, sz);
  __syncthreads();

  const unsigned lt_mask = (1u << lane) - 1u;
  const int s_begin = (blockIdx.x * kGroupWarps + warp) * cpw;
  for (int s = s_begin; s < min(s_begin + cpw, S); ++s) {
    const float* c = new_xyz + ((size_t)b * S + s) * 3;
    const float cx = __ldg(c), cy = __ldg(c + 1), cz = __ldg(c + 2);
    int cnt[kMaxScales], first[kMaxScales];
#pragma unroll
    for (int r = 0; r < kMaxScales; ++r) { cnt[r] = 0; first[r] = N; }
    for (int base = 0; base < N; base += 32) {
      bool open = false;
#pragma unroll
      for (int r = 0; r < kMaxScales; ++r) open |= r < sc.n && cnt[r] < sc.nsample[r];
      if (!open) break;
      const int i = base + lane;
      const float d2 = i < N ? sqdist_rn(sx[i], sy[i], sz[i], cx, cy, cz) : 0.f;
#pragma unroll
      for (int r = 0; r < kMaxScales; ++r) {
        if (r >= sc.n || cnt[r] >= sc.nsample[r]) continue;
        const bool hit = i < N && !(d2 > sc.r2[r]);
        const unsigned m = __ballot_sync(0xFFFFFFFFu, hit);
        if (m) {
          if (cnt[r] == 0) first[r] = base + __ffs(m) - 1;
          const int slot = cnt[r] + __popc(m & lt_mask);
          if (hit && slot < sc.nsample[r]) sc.out[r][((size_t)b * S + s) * sc.nsample[r] + slot] = i;
          cnt[r] += __popc(m);
        }
      }
    }
#pragma unroll
    for (int r = 0; r < kMaxScales; ++r) {
      if (r >= sc.n) continue;
      int32_t* o = sc.out[r] + ((size_t)b * S + s) * sc.nsample[r];
      for (int e = min(cnt[r], sc.nsample[r]) + lane; e < sc.nsample[r]; e += 32) o[e] = first[r];
    }
  }
}

// C == 3 (every call site of the reference): one thread = 4 src rows x 4 consecutive dst points.  The 12 dst floats are
// three 16-byte loads, reused for the 4 rows; 16 outputs leave as four 16-byte stores.
__global__ void __launch_bounds__(256)
square_distance3_kernel(const float* __restrict__ src, const float* __restrict__ dst, int N, int M, float* __restrict__ out) {
  const int b = blockIdx.z, i0 = blockIdx.y * 4;
  const int j0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (j0 >= M) return;
  const float* d = dst + ((size_t)b * M + j0) * 3;
  float dv[12];
  if (j0 + 4 <= M && ((((size_t)b * M + j0) * 3) & 3) == 0) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(d)), bq = __ldg(reinterpret_cast<const float4*>(d) + 1),
                 c = __ldg(reinterpret_cast<const float4*>(d) + 2);
    dv[0] = a.x; dv[1] = a.y; dv[2] = a.z; dv[3] = a.w; dv[4] = bq.x; dv[5] = bq.y; dv[6] = bq.z; dv[7] = bq.w;
    dv[8] = c.x; dv[9] = c.y; dv[10] = c.z; dv[11] = c.w;
  } else {
#pragma unroll
    for (int e = 0; e < 12; ++e) dv[e] = j0 + e / 3 < M ? __ldg(d + e) : 0.f;
  }
  float nd[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) nd[u] = fmaf(dv[3 * u + 2], dv[3 * u + 2], fmaf(dv[3 * u + 1], dv[3 * u + 1], dv[3 * u] * dv[3 * u]));
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int i = i0 + r;
    if (i >= N) break;
    const float* sp = src + ((size_t)b * N + i) * 3;
    const float sx = __ldg(sp), sy = __ldg(sp + 1), sz = __ldg(sp + 2);
    const float ns = fmaf(sz, sz, fmaf(sy, sy, sx * sx));
    float o4[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float dot = fmaf(sz, dv[3 * u + 2], fmaf(sy, dv[3 * u + 1], sx * dv[3 * u]));
      o4[u] = (-2.f * dot + ns) + nd[u];
    }
    float* o = out + ((size_t)b * N + i) * M + j0;
    if (j0 + 4 <= M && ((((size_t)b * N + i) * M + j0) & 3) == 0) {
      *reinterpret_cast<float4*>(o) = make_float4(o4[0], o4[1], o4[2], o4[3]);
    } else {
      for (int u = 0; u < 4 && j0 + u < M; ++u) o[u] = o4[u];
    }
  }
}

static int group_smem(int N, int K, size_t* smem, int warps = kGroupWarps) {
  *smem = (size_t)N * 3 * sizeof(float) + (size_t)warps * ((K + 31) / 32 * 32) * 8;
  return *smem <= 200 * 1024;
}

}  // namespace pcoe

using namespace pcoe;

extern "C" int pcoe_knn_f32(const float* xyz, const float* new_xyz, int B, int N, int S, int K,
                            int32_t* out_idx, void* stream) {
  if (B <= 0 || N <= 0 || S <= 0 || K <= 0)
    return fail(PCOE_ERR_BAD_SHAPE, "knn: B=%d N=%d S=%d K=%d", B, N, S, K);
  if (K > N) return fail(PCOE_ERR_BAD_SHAPE, "knn: K=%d exceeds N=%d (topk would raise)", K, N);
  if (K > 128) return fail(PCOE_ERR_UNSUPPORTED, "knn: K=%d > 128", K);
  if (!xyz || !new_xyz || !out_idx) return fail(PCOE_ERR_NULL, "knn: NULL pointer");
  size_t smem;
  const int warps = N <= 2048 ? 8 : (N <= 4096 ? 16 : 32);   // big clouds: fewer CTAs re-stage the cloud
  if (!group_smem(N, K, &smem, warps)) return fail(PCOE_ERR_UNSUPPORTED, "knn: N=%d does not fit shared memory", N);
  dim3 grid(ceil_div(S, warps), B);
#define PCOE_KNN(KR_)                                                                                       \
  {                                                                                                         \
    if (smem > 48 * 1024)                                                                                   \
      PCOE_CUDA(cudaFuncSetAttribute(knn_kernel<KR_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    LaunchScope ls("knn_kernel", (cudaStream_t)stream);                                                     \
    knn_kernel<KR_><<<grid, warps * 32, smem, (cudaStream_t)stream>>>(xyz, new_xyz, N, S, K, out_idx); \
    return ls.done();                                                                                       \
  }
  if (K <= 32) {   // candidate pre-selection kernel
    smem = (size_t)N * 3 * sizeof(float) + (size_t)warps * 32 * kCandPerLane * 8;
    if (smem > 200 * 1024) return fail(PCOE_ERR_UNSUPPORTED, "knn: N=%d does not fit shared memory", N);
    LaunchScope ls("knn_kernel", (cudaStream_t)stream);
#define PCOE_KNN32(W_)                                                                                              \
    {                                                                                                               \
      if (smem > 48 * 1024)                                                                                         \
        PCOE_CUDA(cudaFuncSetAttribute(knn_k32_kernel<W_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
      knn_k32_kernel<W_><<<grid, W_ * 32, smem, (cudaStream_t)stream>>>(xyz, new_xyz, N, S, K, out_idx);             \
    }
    if (warps == 8) PCOE_KNN32(8) else if (warps == 16) PCOE_KNN32(16) else PCOE_KNN32(32)
#undef PCOE_KNN32
    return ls.done();
  }
  if (K <= 64) PCOE_KNN(2) else PCOE_KNN(4)
#undef PCOE_KNN
}

extern "C" int pcoe_ball_query_f32(const float* xyz, const float* new_xyz, int B, int N, int S,
                                   int nsample, double radius, int32_t* out_idx, void* stream) {
  if (B <= 0 || N <= 0 || S <= 0 || nsample <= 0)
    return fail(PCOE_ERR_BAD_SHAPE, "ball_query: B=%d N=%d S=%d nsample=%d", B, N, S, nsample);
  if (!xyz || !new_xyz || !out_idx) return fail(PCOE_ERR_NULL, "ball_query: NULL pointer");
  size_t smem;
  if (!group_smem(N, 0, &smem))
    return fail(PCOE_ERR_UNSUPPORTED, "ball_query: N=%d does not fit shared memory", N);
  if (smem > 48 * 1024)
    PCOE_CUDA(cudaFuncSetAttribute(ball_query_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const float r2 = (float)(radius * radius);  // Python double r**2, cast to the tensor dtype
  const int cpw = ball_cpw(B, S);
  dim3 grid(ceil_div(S, kGroupWarps * cpw), B);
  LaunchScope ls("ball_query_kernel", (cudaStream_t)stream);
  ball_query_kernel<<<grid, kGroupWarps * 32, smem, (cudaStream_t)stream>>>(xyz, new_xyz, N, S, nsample, r2, out_idx, cpw);
  return ls.done();
}

extern "C" int pcoe_square_distance_f32(const float* src, const float* dst, int B, int N, int M, int C, float* out,
                                        void* stream) {
  if (B <= 0 || N <= 0 || M <= 0 || C <= 0) return fail(PCOE_ERR_BAD_SHAPE, "square_distance: B=%d N=%d M=%d C=%d", B, N, M, C);
  if (N > 65535 || B > 65535) return fail(PCOE_ERR_UNSUPPORTED, "square_distance: B=%d / N=%d exceed the grid limits", B, N);
  if (!src || !dst || !out) return fail(PCOE_ERR_NULL, "square_distance: NULL pointer");
  const int threads = M >= 1024 ? 256 : (M >= 256 ? 64 : 32);
  LaunchScope ls("square_distance_kernel", (cudaStream_t)stream);
  if (C == 3) {
    dim3 grid(ceil_div(ceil_div(M, 4), threads), ceil_div(N, 4), B);
    square_distance3_kernel<<<grid, threads, 0, (cudaStream_t)stream>>>(src, dst, N, M, out);
  } else {
    dim3 grid(ceil_div(ceil_div(M, 4), threads), N, B);
    square_distance_kernel<<<grid, threads, 0, (cudaStream_t)stream>>>(src, dst, N, M, C, out);
  }
  return ls.done();
}

extern "C" int pcoe_ball_query_multi_f32(const float* xyz, const float* new_xyz, int B, int N, int S, int nscales,
                                         const double* radius_host, const int* nsample_host, int32_t* const* out_idx_host,
                                         void* stream) {
  if (B <= 0 || N <= 0 || S <= 0) return fail(PCOE_ERR_BAD_SHAPE, "ball_query_multi: B=%d N=%d S=%d", B, N, S);
  if (nscales < 1 || nscales > kMaxScales)
    return fail(PCOE_ERR_UNSUPPORTED, "ball_query_multi: %d scales (1..%d supported)", nscales, kMaxScales);
  if (!xyz || !new_xyz || !radius_host || !nsample_host || !out_idx_host) return fail(PCOE_ERR_NULL, "ball_query_multi: NULL pointer");
  BallScales sc{};
  sc.n = nscales;
  for (int r = 0; r < nscales; ++r) {
    if (nsample_host[r] <= 0) return fail(PCOE_ERR_BAD_SHAPE, "ball_query_multi: nsample[%d]=%d", r, nsample_host[r]);
    if (!out_idx_host[r]) return fail(PCOE_ERR_NULL, "ball_query_multi: out_idx[%d] is NULL", r);
    sc.r2[r] = (float)(radius_host[r] * radius_host[r]);
    sc.nsample[r] = nsample_host[r];
    sc.out[r] = out_idx_host[r];
  }
  size_t smem;
  if (!group_smem(N, 0, &smem)) return fail(PCOE_ERR_UNSUPPORTED, "ball_query_multi: N=%d does not fit shared memory", N);
  if (smem > 48 * 1024)
    PCOE_CUDA(cudaFuncSetAttribute(ball_query_multi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int cpw = ball_cpw(B, S);
  dim3 grid(ceil_div(S, kGroupWarps * cpw), B);
  LaunchScope ls("ball_query_multi_kernel", (cudaStream_t)stream);
  ball_query_multi_kernel<<<grid, kGroupWarps * 32, smem, (cudaStream_t)stream>>>(xyz, new_xyz, N, S, sc, cpw);
  return ls.done();
}
