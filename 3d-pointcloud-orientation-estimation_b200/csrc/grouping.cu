// grouping.cu — neighbour search around sampled centroids.
//   pcoe_knn_f32         k nearest neighbours  (models/base.py:20-35: square_distance + topk)
//   pcoe_ball_query_f32  radius query          (PointNet++Demo.py:49-70)
// Both keep one cloud (SoA) in shared memory per CTA and give one warp to one centroid at a time;
// the (B,S,N) distance matrix of the reference is never materialised.
#include "common.cuh"
#include <math.h>

namespace pcoe {

constexpr int kGroupWarps = 8;         // warps per CTA
constexpr int kCentroidsPerWarp = 4;   // centroids a warp walks through sequentially

__device__ __forceinline__ void load_cloud_soa(const float* __restrict__ cloud, int N, float* sx,
                                               float* sy, float* sz) {
  // N*3 floats, 16-byte vectorised when the cloud base allows it (always for 4 | 3N and aligned B).
  const int total = N * 3;
  if ((((uintptr_t)cloud) & 15) == 0) {
    const float4* c4 = reinterpret_cast<const float4*>(cloud);
    for (int q = threadIdx.x; q < total / 4; q += blockDim.x) {
      float4 v = __ldg(c4 + q);
      float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        int i = q * 4 + u, p = i / 3, c = i - p * 3;
        (c == 0 ? sx : (c == 1 ? sy : sz))[p] = vv[u];
      }
    }
    for (int i = (total / 4) * 4 + threadIdx.x; i < total; i += blockDim.x) {
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

// kNN: per-warp sorted list (ascending distance) in shared memory, K/32 entries per lane.
// A candidate enters only if it is strictly closer than the current K-th entry; points are
// visited in ascending index so equal distances keep the lower index (matches a stable top-k).
__global__ void __launch_bounds__(kGroupWarps * 32)
knn_kernel(const float* __restrict__ xyz, const float* __restrict__ new_xyz, int N, int S, int K,
           int32_t* __restrict__ out_idx) {
  extern __shared__ float smem_f[];
  float* sx = smem_f;
  float* sy = sx + N;
  float* sz = sy + N;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* list_d = sz + N + warp * K;
  int32_t* list_i = reinterpret_cast<int32_t*>(sz + N + kGroupWarps * K) + warp * K;

  const int b = blockIdx.y;
  load_cloud_soa(xyz + (size_t)b * N * 3, N, sx, sy, sz);
  __syncthreads();

  const int s_begin = (blockIdx.x * kGroupWarps + warp) * kCentroidsPerWarp;
  for (int s = s_begin; s < min(s_begin + kCentroidsPerWarp, S); ++s) {
    const float* c = new_xyz + ((size_t)b * S + s) * 3;
    const float cx = __ldg(c), cy = __ldg(c + 1), cz = __ldg(c + 2);
    for (int e = lane; e < K; e += 32) { list_d[e] = INFINITY; list_i[e] = -1; }
    __syncwarp();
    float kth = INFINITY;
    int filled = 0;  // number of real entries (<= K)
    for (int base = 0; base < N; base += 32) {
      const int i = base + lane;
      float d = INFINITY;
      if (i < N) d = sqdist_rn(sx[i], sy[i], sz[i], cx, cy, cz);
      unsigned pass = __ballot_sync(0xFFFFFFFFu, i < N && (d < kth || filled < K));
      while (pass) {
        const int src = __ffs(pass) - 1;
        pass &= pass - 1;
        const float dv = __shfl_sync(0xFFFFFFFFu, d, src);
        const int iv = base + src;
        if (!(dv < kth || filled < K)) continue;  // threshold moved since the ballot
        // position = number of entries with distance <= dv (they all have a lower index)
        int pos = 0;
        for (int e0 = 0; e0 < K; e0 += 32) {
          int e = e0 + lane;
          pos += __popc(__ballot_sync(0xFFFFFFFFu, e < K && list_d[e] <= dv));
        }
        // shift [pos, K-1) up by one, highest chunk first so reads precede overwrites
        for (int e0 = ((K - 1) / 32) * 32; e0 >= 0; e0 -= 32) {
          int e = e0 + lane;
          float pd = 0.f; int pi = 0;
          bool mv = e < K && e > pos;
          if (mv) { pd = list_d[e - 1]; pi = list_i[e - 1]; }
          __syncwarp();
          if (mv) { list_d[e] = pd; list_i[e] = pi; }
          __syncwarp();
        }
        if (lane == 0) { list_d[pos] = dv; list_i[pos] = iv; }
        __syncwarp();
        if (filled < K) ++filled;
        kth = (filled == K) ? list_d[K - 1] : INFINITY;
      }
    }
    int32_t* o = out_idx + ((size_t)b * S + s) * K;
    for (int e = lane; e < K; e += 32) o[e] = list_i[e];
    __syncwarp();
  }
}

// Ball query: ascending-index scan, the first `nsample` hits, padded with the first hit.
__global__ void __launch_bounds__(kGroupWarps * 32)
ball_query_kernel(const float* __restrict__ xyz, const float* __restrict__ new_xyz, int N, int S,
                  int nsample, float r2, int32_t* __restrict__ out_idx) {
  extern __shared__ float smem_f[];
  float* sx = smem_f;
  float* sy = sx + N;
  float* sz = sy + N;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = blockIdx.y;
  load_cloud_soa(xyz + (size_t)b * N * 3, N, sx, sy, sz);
  __syncthreads();

  const int s_begin = (blockIdx.x * kGroupWarps + warp) * kCentroidsPerWarp;
  for (int s = s_begin; s < min(s_begin + kCentroidsPerWarp, S); ++s) {
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

static int group_smem(int N, int K, size_t* smem) {
  *smem = (size_t)N * 3 * sizeof(float) + (size_t)kGroupWarps * K * 8;
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
  if (!group_smem(N, K, &smem)) return fail(PCOE_ERR_UNSUPPORTED, "knn: N=%d does not fit shared memory", N);
  if (smem > 48 * 1024)
    PCOE_CUDA(cudaFuncSetAttribute(knn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(ceil_div(S, kGroupWarps * kCentroidsPerWarp), B);
  LaunchScope ls("knn_kernel", (cudaStream_t)stream);
  knn_kernel<<<grid, kGroupWarps * 32, smem, (cudaStream_t)stream>>>(xyz, new_xyz, N, S, K, out_idx);
  return ls.done();
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
  dim3 grid(ceil_div(S, kGroupWarps * kCentroidsPerWarp), B);
  LaunchScope ls("ball_query_kernel", (cudaStream_t)stream);
  ball_query_kernel<<<grid, kGroupWarps * 32, smem, (cudaStream_t)stream>>>(xyz, new_xyz, N, S, nsample, r2, out_idx);
  return ls.done();
}
