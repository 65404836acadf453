// sampling.cu — centroid sampling for the set-abstraction layer.
//   pcoe_fps_f32           farthest-point sampling, one CTA per cloud (PointNet++Demo.py:8-29)
//   pcoe_gather_points_f32 index_points with a (B,S) index (models/base.py:4-14)
//   pcoe_random_subset     device-side randperm(N)[:S] (models/pointnet_pp_8dir.py:28 semantics)
#include "common.cuh"

namespace pcoe {

// ---------------------------------------------------------------------------------------------
// FPS.  The cloud lives in shared memory (SoA) for the centroid broadcast and in registers for the
// per-thread distance update; the running minimum distance never leaves registers.  Each of the S
// dependent iterations costs one block barrier: a warp arg-max is two REDUX instructions (max of
// the distance bit pattern, then min index among the lanes that hold it), warp winners go through
// a double-buffered shared array and every warp reduces those redundantly.
// ---------------------------------------------------------------------------------------------
template <int PPT, bool SMEM_CLOUD>
__global__ void __launch_bounds__(1024)
fps_kernel(const float* __restrict__ xyz, int N, int S, const int32_t* __restrict__ start_idx,
           int32_t* __restrict__ out_idx, float* __restrict__ out_xyz) {
  extern __shared__ float smem_f[];
  __shared__ uint32_t s_val[2][32];
  __shared__ uint32_t s_idx[2][32];

  const int b = blockIdx.x, t = threadIdx.x, T = blockDim.x;
  const int lane = t & 31, warp = t >> 5, nwarps = T >> 5;
  const float* cloud = xyz + (size_t)b * N * 3;
  float* sx = smem_f;
  float* sy = sx + N;
  float* sz = sy + N;

  if (SMEM_CLOUD) {
    for (int i = t; i < N * 3; i += T) {  // coalesced read of the AoS cloud, SoA in smem
      float v = __ldg(cloud + i);
      int p = i / 3, c = i - p * 3;
      (c == 0 ? sx : (c == 1 ? sy : sz))[p] = v;
    }
    __syncthreads();
  }

  float px[PPT], py[PPT], pz[PPT], dmin[PPT];
#pragma unroll
  for (int j = 0; j < PPT; ++j) {
    int i = t + j * T;
    if (i < N) {
      if (SMEM_CLOUD) { px[j] = sx[i]; py[j] = sy[i]; pz[j] = sz[i]; }
      else { px[j] = __ldg(cloud + 3 * i); py[j] = __ldg(cloud + 3 * i + 1); pz[j] = __ldg(cloud + 3 * i + 2); }
      dmin[j] = 1e10f;
    } else {
      px[j] = py[j] = pz[j] = 0.f;
      dmin[j] = 0.f;  // never beats a real point: ties go to the lowest index
    }
  }

  int cur = start_idx ? start_idx[b] : 0;
  cur = min(max(cur, 0), N - 1);
  int buf = 0;
  for (int it = 0; it < S; ++it) {
    float cx, cy, cz;
    if (SMEM_CLOUD) { cx = sx[cur]; cy = sy[cur]; cz = sz[cur]; }
    else { cx = __ldg(cloud + 3 * cur); cy = __ldg(cloud + 3 * cur + 1); cz = __ldg(cloud + 3 * cur + 2); }
    if (t == 0) {
      out_idx[(size_t)b * S + it] = cur;
      if (out_xyz) {
        float* o = out_xyz + ((size_t)b * S + it) * 3;
        o[0] = cx; o[1] = cy; o[2] = cz;
      }
    }
    if (it == S - 1) break;

    uint32_t best_v = 0u, best_i = 0xFFFFFFFFu;
#pragma unroll
    for (int j = 0; j < PPT; ++j) {
      int i = t + j * T;
      float d = sqdist_rn(px[j], py[j], pz[j], cx, cy, cz);
      float m = (d < dmin[j]) ? d : dmin[j];  // mask = dist < distance; distance[mask] = dist[mask]
      dmin[j] = m;
      uint32_t bits = __float_as_uint(m);     // m >= 0: bit pattern is order preserving
      bool valid = i < N;
      if (valid && (bits > best_v || best_i == 0xFFFFFFFFu)) { best_v = bits; best_i = (uint32_t)i; }
    }
    uint32_t wv = __reduce_max_sync(0xFFFFFFFFu, best_v);
    uint32_t wi = __reduce_min_sync(0xFFFFFFFFu, best_v == wv ? best_i : 0xFFFFFFFFu);
    if (nwarps > 1) {
      if (lane == 0) { s_val[buf][warp] = wv; s_idx[buf][warp] = wi; }
      __syncthreads();
      uint32_t v = lane < nwarps ? s_val[buf][lane] : 0u;
      uint32_t ix = lane < nwarps ? s_idx[buf][lane] : 0xFFFFFFFFu;
      wv = __reduce_max_sync(0xFFFFFFFFu, v);
      wi = __reduce_min_sync(0xFFFFFFFFu, v == wv ? ix : 0xFFFFFFFFu);
      buf ^= 1;
    }
    cur = (int)wi;
  }
}

template <int PPT>
static int launch_fps(const float* xyz, int B, int N, int S, const int32_t* start, int32_t* out_idx,
                      float* out_xyz, int threads, cudaStream_t st) {
  size_t smem = (size_t)N * 3 * sizeof(float);
  LaunchScope ls("fps_kernel", st);
  if (smem <= 200 * 1024) {
    auto k = fps_kernel<PPT, true>;
    if (smem > 48 * 1024)
      PCOE_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<B, threads, smem, st>>>(xyz, N, S, start, out_idx, out_xyz);
  } else {
    fps_kernel<PPT, false><<<B, threads, 0, st>>>(xyz, N, S, start, out_idx, out_xyz);
  }
  return ls.done();
}

// ---------------------------------------------------------------------------------------------
__global__ void gather_points_kernel(const float* __restrict__ src, int N, int C,
                                     const int32_t* __restrict__ idx, int S, float* __restrict__ out,
                                     size_t total) {
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (size_t)gridDim.x * blockDim.x) {
    int c = (int)(e % C);
    size_t bs = e / C;
    int b = (int)(bs / S);
    int i = idx[bs];
    i = min(max(i, 0), N - 1);
    out[e] = __ldg(src + ((size_t)b * N + i) * C + c);
  }
}

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al. 2011), counter = (i, cloud, offset), key = seed.
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
    uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += W0; key.y += W1;
  }
  return ctr;
}

// One CTA per cloud: identity permutation in shared memory, S Fisher-Yates swaps by one thread
// (the swaps are a dependent chain), the S random offsets are drawn in parallel beforehand.
// xyz != nullptr: the selected points are gathered as well (out_xyz [B,S,3]) - saves the gather launch that follows.
__global__ void random_subset_kernel(int N, int S, uint64_t seed, uint64_t offset,
                                     const uint64_t* __restrict__ offset_dev, int32_t* __restrict__ out_idx,
                                     const float* __restrict__ xyz, float* __restrict__ out_xyz) {
  if (offset_dev) offset += *offset_dev;
  extern __shared__ int32_t s_perm[];
  int32_t* s_j = s_perm + N;
  const int b = blockIdx.x, t = threadIdx.x, T = blockDim.x;
  for (int i = t; i < N; i += T) s_perm[i] = i;
  for (int i = t; i < S; i += T) {
    uint4 r = philox4x32_10(make_uint4((uint32_t)i, (uint32_t)b, (uint32_t)offset, (uint32_t)(offset >> 32)),
                            make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
    uint32_t span = (uint32_t)(N - i);
    s_j[i] = i + (int32_t)(((uint64_t)r.x * span) >> 32);
  }
  __syncthreads();
  if (t == 0) {
    for (int i = 0; i < S; ++i) {
      int j = s_j[i];
      int32_t a = s_perm[i], c = s_perm[j];
      s_perm[i] = c; s_perm[j] = a;
    }
  }
  __syncthreads();
  for (int i = t; i < S; i += T) out_idx[(size_t)b * S + i] = s_perm[i];
  if (xyz)
    for (int e = t; e < S * 3; e += T) {
      const int i = e / 3, c = e - i * 3;
      out_xyz[(size_t)b * S * 3 + e] = __ldg(xyz + ((size_t)b * N + s_perm[i]) * 3 + c);
    }
}

}  // namespace pcoe

using namespace pcoe;

extern "C" int pcoe_fps_f32(const float* xyz, int B, int N, int S, const int32_t* start_idx,
                            int32_t* out_idx, float* out_xyz, void* stream) {
  if (B <= 0 || N <= 0 || S <= 0) return fail(PCOE_ERR_BAD_SHAPE, "fps: B=%d N=%d S=%d", B, N, S);
  if (!xyz || !out_idx) return fail(PCOE_ERR_NULL, "fps: xyz/out_idx is NULL");
  if (N > 16384) return fail(PCOE_ERR_UNSUPPORTED, "fps: N=%d > 16384 points per cloud", N);
  cudaStream_t st = (cudaStream_t)stream;
  int threads = ((N + 3) / 4 + 31) / 32 * 32;
  threads = threads < 32 ? 32 : (threads > 1024 ? 1024 : threads);
  int ppt = (N + threads - 1) / threads;
  if (ppt <= 1) return launch_fps<1>(xyz, B, N, S, start_idx, out_idx, out_xyz, threads, st);
  if (ppt <= 2) return launch_fps<2>(xyz, B, N, S, start_idx, out_idx, out_xyz, threads, st);
  if (ppt <= 4) return launch_fps<4>(xyz, B, N, S, start_idx, out_idx, out_xyz, threads, st);
  if (ppt <= 8) return launch_fps<8>(xyz, B, N, S, start_idx, out_idx, out_xyz, threads, st);
  return launch_fps<16>(xyz, B, N, S, start_idx, out_idx, out_xyz, threads, st);
}

extern "C" int pcoe_gather_points_f32(const float* src, int B, int N, int C, const int32_t* idx,
                                      int S, float* out, void* stream) {
  if (B <= 0 || N <= 0 || C <= 0 || S <= 0)
    return fail(PCOE_ERR_BAD_SHAPE, "gather_points: B=%d N=%d C=%d S=%d", B, N, C, S);
  if (!src || !idx || !out) return fail(PCOE_ERR_NULL, "gather_points: NULL pointer");
  size_t total = (size_t)B * S * C;
  int blocks = (int)((total + 255) / 256);
  if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
  LaunchScope ls("gather_points_kernel", (cudaStream_t)stream);
  gather_points_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(src, N, C, idx, S, out, total);
  return ls.done();
}

extern "C" int pcoe_random_subset_xyz(int B, int N, int S, uint64_t seed, uint64_t offset, const uint64_t* offset_dev,
                                      int32_t* out_idx, const float* xyz, float* out_xyz, void* stream);
extern "C" int pcoe_random_subset(int B, int N, int S, uint64_t seed, uint64_t offset,
                                  const uint64_t* offset_dev, int32_t* out_idx, void* stream) {
  return pcoe_random_subset_xyz(B, N, S, seed, offset, offset_dev, out_idx, nullptr, nullptr, stream);
}

extern "C" int pcoe_random_subset_xyz(int B, int N, int S, uint64_t seed, uint64_t offset, const uint64_t* offset_dev,
                                      int32_t* out_idx, const float* xyz, float* out_xyz, void* stream) {
  if (B <= 0 || N <= 0 || S <= 0 || S > N)
    return fail(PCOE_ERR_BAD_SHAPE, "random_subset: B=%d N=%d S=%d", B, N, S);
  if (!out_idx) return fail(PCOE_ERR_NULL, "random_subset: out_idx is NULL");
  if ((xyz == nullptr) != (out_xyz == nullptr)) return fail(PCOE_ERR_NULL, "random_subset: xyz and out_xyz go together");
  size_t smem = ((size_t)N + S) * sizeof(int32_t);
  if (smem > 200 * 1024) return fail(PCOE_ERR_UNSUPPORTED, "random_subset: N=%d too large", N);
  if (smem > 48 * 1024)
    PCOE_CUDA(cudaFuncSetAttribute(random_subset_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  LaunchScope ls("random_subset_kernel", (cudaStream_t)stream);
  random_subset_kernel<<<B, 128, smem, (cudaStream_t)stream>>>(N, S, seed, offset, offset_dev, out_idx, xyz, out_xyz);
  return ls.done();
}
