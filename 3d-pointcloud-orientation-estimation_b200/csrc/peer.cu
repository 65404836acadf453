// peer.cu — gradient all-reduce over NVLink peer memory (one node, one process per GPU).
//
// The data-parallel exchange of this path is ONE sum over the flat fp32 gradient buffer (5.9 MB) per step.  At that size
// a collective is latency-bound, not bandwidth-bound, so the ranks run one small kernel each over SYMMETRIC memory
// (every rank maps every peer's buffer: torch.distributed._symmetric_memory supplies the mappings, this file the
// kernel).  Two-shot, in place, with the second shot PUSHED so that two flag exchanges are the only synchronisation:
//     flag A: "my gradients are complete" (written to every peer by one thread at kernel start; every CTA polls its own
//             rank's pad, no grid barrier)
//     rank r: for each element of slice r: load it from every rank in rank order, add, STORE the sum into every
//             rank's buffer (the only reader of element i of anybody's slice r is the thread that overwrites it)
//     flag B: "my pushes have landed" (per-CTA system fence + arrival count; the last CTA of the rank writes the flag
//             and is the only one that waits for the peers' B flags; the other CTAs retire at once)
// When a rank's kernel completes, every peer has finished reading and writing that rank's buffer.  Flags are epochs that
// only grow (device-side launch counter), so the launch replays from a CUDA graph.  With an NVSwitch multicast mapping
// of the buffer (mc != NULL) the loop is multimem.ld_reduce (the switch adds the ranks' copies) + multimem.st (the
// switch broadcasts the sum): one load and one store per element instead of `world` of each.
#include "common.cuh"

namespace pcoe {

constexpr int kMaxPeers = 8;
struct PeerCtx {
  float* buf[kMaxPeers];
  uint32_t* pad[kMaxPeers];      // pad[p][16 * phase + src]: epoch flag written by rank src on rank p
  float* mc;                     // multicast mapping of the buffer (NVSwitch) or nullptr
  int rank, world;
};

__device__ __forceinline__ void st_relaxed_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_relaxed_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 mc_ld_reduce(const float* p) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void mc_st(float* p, const float4& v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};"
               ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

__device__ __forceinline__ uint64_t global_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
constexpr uint64_t kPeerTimeoutNs = 60ull * 1000000000ull;

// one thread: wait until every rank's flag of `phase` on THIS rank's pad has reached `epoch`.  A peer that never arrives
// (a rank died, or the ranks made different sequences of calls) must not hang the GPU: after kPeerTimeoutNs the kernel
// reports and traps, which fails every later CUDA call of the process.
__device__ __forceinline__ void wait_flags(const PeerCtx& c, int phase, uint32_t epoch) {
  uint64_t t0 = 0;
  for (int p = 0; p < c.world; ++p) {
    uint32_t polls = 0;
    while ((int32_t)(ld_relaxed_sys(c.pad[c.rank] + 16 * phase + p) - epoch) < 0) {
      if ((++polls & 0xFFFu) == 0u) {
        const uint64_t now = global_ns();
        if (t0 == 0) t0 = now;
        else if (now - t0 > kPeerTimeoutNs) {
          printf("pcoe peer_allreduce: rank %d waited 60 s for rank %d (phase %d, epoch %u): aborting\n", c.rank, p, phase, epoch);
          __trap();
        }
      }
    }
  }
  __threadfence_system();
}
__device__ __forceinline__ void post_flags(const PeerCtx& c, int phase, uint32_t epoch) {
  for (int p = 0; p < c.world; ++p) st_relaxed_sys(c.pad[p] + 16 * phase + c.rank, epoch);
}

constexpr int kPeerUnroll = 4;

// ctl (this rank's own device memory, zeroed once by the caller): [0] launches completed, [4] CTAs arrived at flag B
template <bool MC>
__global__ void __launch_bounds__(256)
peer_allreduce_kernel(PeerCtx c, size_t off, size_t n4 /* float4 elements */, uint32_t* __restrict__ ctl) {
  const uint32_t epoch = ctl[0] + 1u;                           // (the last CTA bumps ctl[0] after everyone has arrived)
  const size_t slice = (n4 + c.world - 1) / c.world;
  const size_t gstride = (size_t)gridDim.x * blockDim.x;
  const size_t gtid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;

  // ---- flag A.  The kernel boundary before this launch made this rank's gradients visible in its L2 (the coherence
  // point peers read through), so one relaxed store per peer announces them.
  if (threadIdx.x == 0) {
    if (blockIdx.x == 0) { __threadfence_system(); post_flags(c, 0, epoch); }
    wait_flags(c, 0, epoch);
  }
  __syncthreads();

  const size_t lo = slice * c.rank, hi = min(n4, lo + slice);
  if constexpr (MC) {
    float4* mc = reinterpret_cast<float4*>(c.mc + off);
    for (size_t i0 = lo + gtid; i0 < hi; i0 += gstride * kPeerUnroll) {
      float4 v[kPeerUnroll];
#pragma unroll
      for (int u = 0; u < kPeerUnroll; ++u) {
        const size_t i = i0 + (size_t)u * gstride;
        if (i < hi) v[u] = mc_ld_reduce(reinterpret_cast<const float*>(mc + i));
      }
#pragma unroll
      for (int u = 0; u < kPeerUnroll; ++u) {
        const size_t i = i0 + (size_t)u * gstride;
        if (i < hi) mc_st(reinterpret_cast<float*>(mc + i), v[u]);
      }
    }
  } else {
    for (size_t i0 = lo + gtid; i0 < hi; i0 += gstride * kPeerUnroll) {
      float4 acc[kPeerUnroll];
#pragma unroll
      for (int u = 0; u < kPeerUnroll; ++u) acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int p = 0; p < c.world; ++p) {                        // fixed rank order: one thread owns the element, every
        const float4* src = reinterpret_cast<const float4*>(c.buf[p] + off);   // rank receives the same bits
        float4 v[kPeerUnroll];
#pragma unroll
        for (int u = 0; u < kPeerUnroll; ++u) {
          const size_t i = i0 + (size_t)u * gstride;
          v[u] = i < hi ? __ldcg(src + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < kPeerUnroll; ++u) { acc[u].x += v[u].x; acc[u].y += v[u].y; acc[u].z += v[u].z; acc[u].w += v[u].w; }
      }
      for (int d = 0; d < c.world; ++d) {                        // push, starting at the own copy
        float4* dst = reinterpret_cast<float4*>(c.buf[(c.rank + d) % c.world] + off);
#pragma unroll
        for (int u = 0; u < kPeerUnroll; ++u) {
          const size_t i = i0 + (size_t)u * gstride;
          if (i < hi) __stcg(dst + i, acc[u]);
        }
      }
    }
  }

  // ---- flag B
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();                                     // this CTA's pushes have been performed at every GPU
    const uint32_t arrived = atomicAdd(&ctl[4], 1u) + 1u;
    if (arrived == gridDim.x) {                                 // last CTA of the rank: handshake, then retire the launch
      __threadfence_system();
      post_flags(c, 1, epoch);
      wait_flags(c, 1, epoch);
      ctl[4] = 0u;
      ctl[0] = epoch;
    }
  }
}

}  // namespace pcoe

using namespace pcoe;

extern "C" int pcoe_peer_allreduce_f32(const void* const* bufs_host, const void* const* pads_host, const void* mc_buf,
                                       int rank, int world, size_t offset, size_t n, uint32_t* ctl_dev, int max_ctas,
                                       void* stream) {
  if (world < 1 || world > kMaxPeers || rank < 0 || rank >= world)
    return fail(PCOE_ERR_BAD_SHAPE, "peer_allreduce: rank=%d world=%d (1..%d ranks)", rank, world, kMaxPeers);
  if (!bufs_host || !pads_host || !ctl_dev) return fail(PCOE_ERR_NULL, "peer_allreduce: NULL pointer");
  if ((offset & 3) || (n & 3)) return fail(PCOE_ERR_BAD_SHAPE, "peer_allreduce: offset=%zu / n=%zu must be multiples of 4 floats", offset, n);
  if (n == 0) return PCOE_OK;
  PeerCtx c{};
  c.rank = rank; c.world = world; c.mc = (float*)mc_buf;
  for (int p = 0; p < world; ++p) {
    if (!bufs_host[p] || !pads_host[p]) return fail(PCOE_ERR_NULL, "peer_allreduce: buffer / pad of rank %d is NULL", p);
    c.buf[p] = (float*)bufs_host[p];
    c.pad[p] = (uint32_t*)pads_host[p];
  }
  const size_t n4 = n / 4;
  const int cap = max_ctas > 0 ? max_ctas : 64;
  int ctas = (int)((n4 / world + 256 * kPeerUnroll - 1) / (256 * kPeerUnroll));
  ctas = ctas < 1 ? 1 : (ctas > cap ? cap : ctas);
  LaunchScope ls("peer_allreduce_kernel", (cudaStream_t)stream);
  if (mc_buf) peer_allreduce_kernel<true><<<ctas, 256, 0, (cudaStream_t)stream>>>(c, offset, n4, ctl_dev);
  else        peer_allreduce_kernel<false><<<ctas, 256, 0, (cudaStream_t)stream>>>(c, offset, n4, ctl_dev);
  return ls.done();
}
