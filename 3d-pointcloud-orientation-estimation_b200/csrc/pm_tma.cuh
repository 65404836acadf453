// pm_tma.cuh — inference-only pointwise-MLP layers with TMA-fed tcgen05 operands (the conv1d(k=1) + BatchNorm(eval) +
// ReLU stacks of the vanilla PointNet, models/pointnet.py:24-27,55-58,93,102-104).
//
// In eval mode the BatchNorm affine of a layer is known before the layer runs, so - unlike a training step, where the
// consumer must wait for the batch statistics - the EPILOGUE of layer l can finish the activation (affine, ReLU), split
// it into bf16 planes and write it in exactly the shared-memory image the tensor core reads (64-channel chunks of
// [64 ch x 128 pts] SWIZZLE_128B parts, one per plane).  Layer l+1 then has no producer warps at all: one thread issues
// a bulk async copy (cp.async.bulk, the TMA engine; SASS UBLKCP) per tile straight into the operand ring, completion
// lands on an mbarrier (complete_tx), the MMA warp issues, the epilogue warps drain TMEM.  fp32 activations never exist
// in HBM; the CUDA cores only touch an element once, in the epilogue that produces it.
//
//   image(tile t, chunk k of nk, plane p)  at byte  ((t * nk + k) * NP + p) * 16384,      NP = 2 (hi + lo: 16 bits)
//   channel-major part [64 ch x 128 pts]:  cm_off(64, c, pt / 8) + (pt % 8) * 2           (B operand, MN-major)
//   point-major   part [128 pts x 64 ch]:  sw128_off(pt, c)                               (B operand, K-major; packed inputs)
//
// Roles (320 threads): warps 0-7 epilogue (TMEM lane = output channel), warp 8 loader (one elected thread), warp 9 MMA
// issue.  CTA x owns output-channel block x % ncb (its weight slice stays resident in shared memory, both planes) and the
// tiles x / ncb, x / ncb + gridDim.x / ncb, ...
#pragma once
#include "sa_tc6.cuh"

namespace pcoe {
namespace pm {

using namespace v4;
using v5::kPart;

constexpr int kPmThreads = 320;
constexpr int kNP = 2;
constexpr uint32_t kOpB = kNP * kPart;      // one 64-channel operand chunk, both planes: 32 KB
constexpr int kPmStages = 4;

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(tc::smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(tc::smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst, uint32_t src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }   // the 8 epilogue warps

// ---- epilogues -------------------------------------------------------------------------------------------------
// last layer: per-32-point-block extreme of the raw accumulator (max where gamma >= 0, min otherwise: the affine is
// monotone per channel); pm_pool_kernel reduces the blocks of a cloud and applies the affine once
struct PoolPm {
  float* __restrict__ ymax;
  float* __restrict__ ymin;
  const float* __restrict__ gamma;
  int C;
  int c;
  bool want_max;
  __host__ __device__ __forceinline__ uint32_t stage_bytes() const { return 0; }
  __device__ __forceinline__ void init(int ch) { c = ch; want_max = c < C ? !signbit(gamma[c]) : true; }
  __device__ __forceinline__ void tile_begin(int) {}
  __device__ __forceinline__ void block(float (&v)[32], int tile, int j, bool valid, uint32_t, int) {
    if (c >= C || !valid) return;
    float e = v[0];
    if (want_max) {
#pragma unroll
      for (int i = 1; i < 32; ++i) e = fmaxf(e, v[i]);
      ymax[(size_t)(tile * 4 + j) * C + c] = e;
    } else {
#pragma unroll
      for (int i = 1; i < 32; ++i) e = fminf(e, v[i]);
      ymin[(size_t)(tile * 4 + j) * C + c] = e;
    }
  }
  __device__ __forceinline__ void tile_end(int, int, uint32_t) {}
  __device__ __forceinline__ void finish() {}
};

// hidden layer: z = act(scale * y + shift) -> two bf16 planes -> the next layer's channel-major operand image, staged in
// shared memory (conflict-free 16-byte stores) and written with one bulk async copy per 64-channel chunk
struct StorePlanesPm {
  uint8_t* __restrict__ oimg;
  const float* __restrict__ scale;
  const float* __restrict__ shift;
  int C, relu;
  int c;
  float sc, sh;
  __host__ __device__ __forceinline__ uint32_t stage_bytes() const { return (C < 128 ? 1u : 2u) * kOpB; }
  __device__ __forceinline__ void init(int ch) {
    c = ch;
    sc = c < C ? scale[c] : 0.f; sh = c < C ? shift[c] : 0.f;
  }
  __device__ __forceinline__ void tile_begin(int) {      // the previous tile's bulk store has finished READING the staging buffer
    if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    epi_bar_sync();
  }
  __device__ __forceinline__ void block(float (&v)[32], int, int j, bool valid, uint32_t sOut, int cb) {
    if (c >= C) return;
    const int cl = c - cb * 128, kc = cl >> 6, r = cl & 63;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float z[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const float t = fmaf(v[8 * q + u], sc, sh);
        z[u] = valid ? (relu ? fmaxf(t, 0.f) : t) : 0.f;
      }
      v6::split_store8<kNP>(z, sOut + (uint32_t)kc * kOpB, cm_off(64, r, j * 4 + q));
    }
  }
  __device__ __forceinline__ void tile_end(int tile, int cb, uint32_t sOut) {
    tc::fence_proxy_async();
    epi_bar_sync();
    if (threadIdx.x == 0) {
      const int nko = C >> 6, kc0 = cb * 2, n = min(2, nko - kc0);
      for (int kc = 0; kc < n; ++kc)
        bulk_s2g(oimg + ((size_t)tile * nko + kc0 + kc) * kOpB, sOut + (uint32_t)kc * kOpB, kOpB);
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
  }
  __device__ __forceinline__ void finish() {
    if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
};

// ---- the layer kernel --------------------------------------------------------------------------------------------
// Y^T[128 ch x 128 pts] = W[128 x 64 nk] * X^T per (tile, channel block); X arrives by bulk async copy.
// smem: [W: nk x 32 KB][X ring: nst x nk x 32 KB][output staging][ - ]
template <class Epi, bool CHMAJOR>
__global__ void __launch_bounds__(kPmThreads, 1)
pm_layer_kernel(const uint8_t* __restrict__ ximg, int nk, const __nv_bfloat16* __restrict__ Wp, size_t wps, int Kp, Epi epi,
                int M, int ncb, int nst) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem0 = (tc::smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ uint64_t full[kPmStages], empty[kPmStages], tmem_full[2], tmem_empty[2];
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (warp == 0) tc::tmem_alloc<256>(&tmem_base);
  if (tid == 0) {
    for (int s = 0; s < kPmStages; ++s) { tc::mbar_init(&full[s], 1); tc::mbar_init(&empty[s], 1); }
    for (int b = 0; b < 2; ++b) { tc::mbar_init(&tmem_full[b], 1); tc::mbar_init(&tmem_empty[b], kEpiThreads); }
  }
  const uint32_t tbytes = (uint32_t)nk * kOpB;
  const uint32_t sW = smem0, sX = sW + tbytes, sOut = sX + (uint32_t)nst * tbytes;
  const int cb = blockIdx.x % ncb, t0 = blockIdx.x / ncb, tstep = gridDim.x / ncb;
  const int ntiles = (M + kPts - 1) / kPts;
  if (tid < 256) {     // resident weight slice of this channel block, both planes
    for (int k = 0; k < nk; ++k)
#pragma unroll
      for (int p = 0; p < kNP; ++p) {
        uint4 w[4];
        v6::wload_k1(Wp + (size_t)p * wps, Kp, cb * 128, k * 64, tid, w);
        v6::wstore_k1(sW + (uint32_t)k * kOpB + (uint32_t)p * kPart, tid, w);
      }
    epi.init(cb * 128 + (warp & 3) * 32 + lane);
  }
  tc::fence_proxy_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = tmem_base;

  if (warp < 8) {
    const int eq = warp & 3, eh = warp >> 2;
    int i = 0;
    for (int tile = t0; tile < ntiles; tile += tstep, ++i) {
      const int b = i & 1, u = i >> 1, m0 = tile * kPts;
      epi.tile_begin(tile);
      tc::mbar_wait(&tmem_full[b], (uint32_t)(u & 1));
      tc::fence_after_sync();
#pragma unroll 1
      for (int j = eh * 2; j < eh * 2 + 2; ++j) {
        float v[32];
        tc::tmem_ld32(tmem + ((uint32_t)(eq * 32) << 16) + (uint32_t)(b * kPts + j * 32), v);
        epi.block(v, tile, j, m0 + j * 32 < M, sOut, cb);
      }
      tc::fence_before_sync();
      mbar_arrive_relaxed(&tmem_empty[b]);
      epi.tile_end(tile, cb, sOut);
    }
    epi.finish();
  } else if (warp == 8) {
    if (lane == 0) {
      int i = 0;
      for (int tile = t0; tile < ntiles; tile += tstep, ++i) {
        const int s = i % nst;
        if (i >= nst) tc::mbar_wait(&empty[s], (uint32_t)((i / nst - 1) & 1));
        mbar_expect_tx(&full[s], tbytes);
        const uint8_t* src = ximg + (size_t)tile * tbytes;
        for (int k = 0; k < nk; ++k)
          bulk_g2s(sX + (uint32_t)s * tbytes + (uint32_t)k * kOpB, src + (size_t)k * kOpB, kOpB, &full[s]);
      }
    }
  } else {   // warp 9: MMA issue, warp-uniform loop, one elected lane issues
    const uint32_t tm = tc::uniform_u32(tmem_base);
    const uint32_t idesc = tc::make_idesc_bf16(128, kPts, false, CHMAJOR);
    int i = 0;
    for (int tile = t0; tile < ntiles; tile += tstep, ++i) {
      const int s = i % nst, b = i & 1, u = i >> 1;
      tc::mbar_wait(&full[s], (uint32_t)((i / nst) & 1));
      if (u > 0) tc::mbar_wait(&tmem_empty[b], (uint32_t)((u - 1) & 1));
      tc::fence_after_sync();
      const uint32_t sXs = sX + (uint32_t)s * tbytes;
      for (int k = 0; k < nk; ++k)
        for (int q = 0; q < 4; ++q) {
          const uint64_t ad = tc::make_desc_sw128(sW + (uint32_t)k * kOpB + (uint32_t)q * 32, 16, 1024);
          const uint64_t bd = CHMAJOR ? tc::make_desc_sw128(sXs + (uint32_t)k * kOpB + (uint32_t)q * 2048, 8192, 1024)
                                      : tc::make_desc_sw128(sXs + (uint32_t)k * kOpB + (uint32_t)q * 32, 16, 1024);
          v6::mma_planes<kNP>(tm + (uint32_t)(b * kPts), ad, bd, idesc, k > 0 || q > 0);
        }
      tc::mma_commit_warp(&empty[s]);
      tc::mma_commit_warp(&tmem_full[b]);
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc<256>(tmem);
}

// ---- image writers for the first layer ------------------------------------------------------------------------------
// xyz stacks: z1 = act(scale * (W1 x) + shift), W1 [C1 x 3], written as the channel-major image of layer 2's operand.
// 1024 threads per (tile, 64-channel chunk); a warp covers 4 channel rows x 8 point chunks = 512 contiguous image bytes.
__global__ void __launch_bounds__(256)
pm_first_xyz_kernel(const float* __restrict__ xyz, int M, const float* __restrict__ W, const float* __restrict__ scale,
                    const float* __restrict__ shift, int C1, int relu, uint8_t* __restrict__ img) {
  const int nko = C1 >> 6;
  const size_t gt = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  const size_t gw = gt >> 5;                        // global warp: 32 warps per (tile, chunk)
  const int w = (int)(gw & 31);
  const size_t tk = gw >> 5;
  const int kc = (int)(tk % nko);
  const int tile = (int)(tk / nko);
  if ((size_t)tile * kPts >= (size_t)M) return;
  const int h = w >> 4, g = (w >> 1) & 7, r = (w & 1) * 4 + (lane >> 3), js = lane & 7;
  const int cl = 8 * g + r, c = kc * 64 + cl, chunk = (js ^ r) + 8 * h;
  const float w0 = W[c * 3], w1 = W[c * 3 + 1], w2 = W[c * 3 + 2], sc = scale[c], sh = shift[c];
  float z[8];
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    const int m = tile * kPts + chunk * 8 + u;
    float t = 0.f;
    if (m < M) {
      const float* p = xyz + (size_t)m * 3;
      t = fmaf(fmaf(w2, __ldg(p + 2), fmaf(w1, __ldg(p + 1), w0 * __ldg(p))), sc, sh);
      if (relu) t = fmaxf(t, 0.f);
    }
    z[u] = t;
  }
  uint4 hi, lo;
  v6::split8(z, hi, lo);
  uint8_t* dst = img + ((size_t)tile * nko + kc) * kOpB + cm_off(64, cl, chunk);
  *reinterpret_cast<uint4*>(dst) = hi;
  *reinterpret_cast<uint4*>(dst + kPart) = lo;
}

// feature stacks: point-major fp32 rows x [M, D] (D a multiple of 64) -> point-major operand image [tile][D/64][plane]
// (the input itself, no activation).  One thread = one point x 8 channels (two 16-byte loads, one 16-byte store per plane).
__global__ void __launch_bounds__(256)
pm_pack_rows_kernel(const float* __restrict__ x, int M, int D, uint8_t* __restrict__ img) {
  const int nk = D >> 6, upr = D >> 3;              // 8-channel units per row
  const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t Mld = ((size_t)M + kPts - 1) / kPts * kPts;
  if (e >= Mld * upr) return;
  const int ju = (int)(e % upr);
  const size_t m = e / upr;
  float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (m < (size_t)M) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(x + m * D + ju * 8)), b = __ldg(reinterpret_cast<const float4*>(x + m * D + ju * 8) + 1);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
  uint4 hi, lo;
  v6::split8(v, hi, lo);
  const int tile = (int)(m >> 7), p = (int)(m & 127), k = ju >> 3, j = ju & 7;
  uint8_t* dst = img + ((size_t)tile * nk + k) * kOpB + tc::sw128_off(p, j * 8);
  *reinterpret_cast<uint4*>(dst) = hi;
  *reinterpret_cast<uint4*>(dst + kPart) = lo;
}

}  // namespace pm
}  // namespace pcoe
