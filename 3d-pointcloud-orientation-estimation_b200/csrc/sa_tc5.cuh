// sa_tc5.cuh — K-chunk-streamed tcgen05 kernels for WIDE set-abstraction layers (bf16 mode) whose
// weight matrices do not fit shared memory: SA3 of the reference models (259 -> 256 -> 512 -> 1024 on
// B*32 rows, models/pointnet_pp_8dir.py:67).  Same orientation, HBM layouts, producers and epilogues
// as sa_tc4.cuh (channels on the TMEM lanes, tile-blocked channel-major activations, one zero-padded
// bf16 weight image [Rp][Kp] per layer); what changes is the decomposition:
//   * the row count is small (2048 rows at 64 clouds) and the channel counts are large, so the grid
//     runs over (128-point tile) x (128-channel block of the output) - and over split-K for wgrad -
//     instead of persistent CTAs that keep the whole weight matrix resident;
//   * the contraction is streamed through a ring of shared-memory stages in chunks of 64 channels
//     (forward / dgrad) or 128 points (wgrad): per chunk the 256 producer threads build BOTH MMA
//     operands (weight slice copied from the L2-resident image, activation slice transformed on the
//     fly), one thread issues the tcgen05.mma group, tcgen05.commit frees the stage.
//
//   tc5_fwd_kernel    Y^T[128 ch x 128 pts]  = W[128 x Cin]      * X^T[Cin x 128]       grid (tiles, Cout/128)
//   tc5_dgrad_kernel  dX^T[128 ch x 128 pts] = W^T[128 x Cout]   * dY^T[Cout x 128]     grid (tiles, Cin/128)
//        PT variant   dX[128 pts x 128 ch]   = dY[128 x Cout]    * W[Cout x 128]        (layer-1 scatter epilogue)
//   tc5_wgrad_kernel  dW[128 x nq]          += dY^T[128 x pts]   * X[pts x nq]          grid (Cout/128, q blocks, splits)
//
// CTA = 17 warps as in v4: warps 0-7 epilogue, 8-15 producers, warp 16 MMA issue.
#pragma once
#include "sa_tc4.cuh"

namespace pcoe {
namespace v5 {

using namespace v4;

constexpr int kMaxStages5 = 6;
constexpr uint32_t kPart = 16384;        // one 64-channel x 128-point (or 128-row x 64-k) operand part

// layer-1 input with features, point-major, streamed in 64-channel blocks: block kb < D/64 holds
// feats[:, 64kb .. 64kb+63], block D/64 holds [xyz - centroid (3) | zeros (13)] (channel order
// [feats | xyz], as the v4 weight images of layer 1).  256 threads: thread g owns the 8-channel unit
// g & 7 of rows (g >> 3) + 32 i.
struct GatherFeat5 {
  static constexpr bool kChMajor = false;
  struct Raw { float4 a[4], b[4]; };
  GatherBase gb;
  const float* __restrict__ feats;
  int D;
  __host__ __device__ __forceinline__ int nblocks() const { return D / 64 + 1; }
  __host__ __device__ __forceinline__ int block_k(int kb) const { return kb < D / 64 ? 64 : 16; }
  __host__ __device__ __forceinline__ int nconst() const { return 0; }
  __device__ __forceinline__ void init(float*, int, int) {}
  __device__ __forceinline__ void load64(int g, int m0, int kb, Raw& r) const {
    if (kb < D / 64) {
      const int j = g & 7;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int row = m0 + (g >> 3) + 32 * i;
        r.a[i] = r.b[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row < gb.M) {
          const float4* src = reinterpret_cast<const float4*>(feats + (size_t)gb.point_of(row) * D + kb * 64 + j * 8);
          r.a[i] = __ldg(src);
          r.b[i] = __ldg(src + 1);
        }
      }
    } else {
      float v[3] = {0.f, 0.f, 0.f}, c[3] = {0.f, 0.f, 0.f};
      if (g < kPts && m0 + g < gb.M) gb.load_xyz_raw(m0 + g, gb.point_of(m0 + g), v, c);
      r.a[0] = make_float4(v[0], v[1], v[2], 0.f);
      r.b[0] = make_float4(c[0], c[1], c[2], 0.f);
    }
  }
  __device__ __forceinline__ void store64(int g, int kb, const Raw& r, uint32_t saddr) const {
    if (kb < D / 64) {
      const int j = g & 7;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float v[8] = {r.a[i].x, r.a[i].y, r.a[i].z, r.a[i].w, r.b[i].x, r.b[i].y, r.b[i].z, r.b[i].w};
        tc::sts128(saddr + tc::sw128_off((g >> 3) + 32 * i, j * 8), tc::pack8_bf16(v));
      }
    } else if (g < kPts) {
      const float v[8] = {gb.centred(r.a[0].x, r.b[0].x), gb.centred(r.a[0].y, r.b[0].y), gb.centred(r.a[0].z, r.b[0].z),
                          0.f, 0.f, 0.f, 0.f, 0.f};
      tc::sts128(saddr + tc::sw128_off(g, 0), tc::pack8_bf16(v));
      tc::sts128(saddr + tc::sw128_off(g, 8), make_uint4(0u, 0u, 0u, 0u));
    }
  }
};

struct Barriers5 {
  uint64_t full[kMaxStages5];
  uint64_t empty[kMaxStages5];
  uint64_t tmem_full;
};

// Two-deep register software pipeline over W work units: the global loads of unit w+1 are in flight
// while unit w is transformed and stored to shared memory.
template <class Raw, class LoadF, class StoreF>
__device__ __forceinline__ void unit_pipeline(int W, LoadF load, StoreF store) {
  if (W <= 0) return;
  Raw ra, rb;
  load(0, ra);
  for (int w = 0; w < W; w += 2) {
    if (w + 1 < W) load(w + 1, rb);
    store(w, ra);
    if (w + 1 < W) {
      if (w + 2 < W) load(w + 2, ra);
      store(w + 1, rb);
    }
  }
}

// weight slice [128 rows x 64 k] of the image -> K-major SWIZZLE_128B part (forward A operand)
__device__ __forceinline__ void wload_k(const __nv_bfloat16* __restrict__ Wb, int Kp, int row0, int k0, int g, uint4 (&w)[4]) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
    w[i] = __ldg(reinterpret_cast<const uint4*>(Wb + (size_t)(row0 + (g >> 3) + 32 * i) * Kp + k0 + (g & 7) * 8));
}
__device__ __forceinline__ void wstore_k(uint32_t saddr, int g, const uint4 (&w)[4]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) tc::sts128(saddr + tc::sw128_off((g >> 3) + 32 * i, (g & 7) * 8), w[i]);
}
// weight slice [64 k rows x 128 columns] -> MN-major part: 2 blocks of [64 rows x 64 columns]
__device__ __forceinline__ void wload_mn(const __nv_bfloat16* __restrict__ Wb, int Kp, int row0, int col0, int g, uint4 (&w)[4]) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
    w[i] = __ldg(reinterpret_cast<const uint4*>(Wb + (size_t)(row0 + (g >> 4) + 16 * i) * Kp + col0 + (g & 15) * 8));
}
__device__ __forceinline__ void wstore_mn(uint32_t saddr, int g, const uint4 (&w)[4]) {
  const int j = g & 15;
#pragma unroll
  for (int i = 0; i < 4; ++i)
    tc::sts128(saddr + (uint32_t)(j >> 3) * 8192u + tc::sw128_off((g >> 4) + 16 * i, (j & 7) * 8), w[i]);
}

#define PCOE_V5_PROLOGUE(NBAR_STAGES)                                                              \
  extern __shared__ uint8_t smem_raw[];                                                            \
  const uint32_t smem0 = (tc::smem_u32(smem_raw) + 1023u) & ~1023u;                                \
  uint8_t* smem_gen = smem_raw + (smem0 - tc::smem_u32(smem_raw));                                 \
  __shared__ Barriers5 bar;                                                                        \
  __shared__ uint32_t tmem_base;                                                                   \
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;                                   \
  if (warp == 0) tc::tmem_alloc<128>(&tmem_base);                                                  \
  if (tid == 0) {                                                                                  \
    for (int s = 0; s < (NBAR_STAGES); ++s) { tc::mbar_init(&bar.full[s], kProdThreads); tc::mbar_init(&bar.empty[s], 1); } \
    tc::mbar_init(&bar.tmem_full, 1);                                                              \
  }

// ---------------------------------------------------------------------------------------------
// forward: one (tile, 128-channel block) per CTA.  Stage = [W part 16 KB][activation part 16 KB].
// ---------------------------------------------------------------------------------------------
template <class Prod, class Epi>
__global__ void __launch_bounds__(kThreads, 1)
tc5_fwd_kernel(Prod prod, const __nv_bfloat16* __restrict__ Wb, int Kp, Epi epi, int M, int nst) {
  PCOE_V5_PROLOGUE(nst)
  const int tile = blockIdx.x, cb = blockIdx.y, m0 = tile * kPts;
  const uint32_t sS = smem0, sbytes = 2 * kPart;
  float* csm = reinterpret_cast<float*>(smem_gen + (size_t)nst * sbytes);
  prod.init(csm, tid, kThreads);
  const int eq = warp & 3, eh = (warp >> 2) & 1;
  epi.init(csm + prod.nconst(), cb * 128 + eq * 32 + lane);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = tmem_base;
  int nk;
  if constexpr (Prod::kChMajor) nk = prod.kext() / 64; else nk = prod.nblocks();

  if (warp < 8) {
    tc::mbar_wait(&bar.tmem_full, 0u);
    tc::fence_after_sync();
    const uint32_t stg = sS;                                   // every stage is free once the accumulator is complete
#pragma unroll 1
    for (int j = eh * 2; j < eh * 2 + 2; ++j) {
      float v[32];
      tc::tmem_ld32(tmem + ((uint32_t)(eq * 32) << 16) + (uint32_t)(j * 32), v);
      epi.block(v, tile, j, m0 + j * 32 < M, stg - (uint32_t)cb * 32768u);
    }
    tc::fence_before_sync();
    if (Epi::kStage && epi.tile_dst(0) != nullptr) {
      stage_copy_out(epi.tile_dst(tile) + (size_t)cb * 32768, stg, 32768, 1, tid == 0);
      if (tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
    epi.finish();
  } else if (warp < 16) {
    const int g = tid - kEpiThreads;
    union RawU { typename Prod::Raw a; uint4 w[4]; __device__ RawU() {} };
    unit_pipeline<RawU>(2 * nk,
        [&](int w, RawU& r) {
          const int k = w >> 1;
          if ((w & 1) == 0) wload_k(Wb, Kp, cb * 128, k * 64, g, r.w);
          else if constexpr (Prod::kChMajor) prod.load(g, kProdThreads, m0, k, NoIdx{}, r.a);
          else prod.load64(g, m0, k, r.a);
        },
        [&](int w, const RawU& r) {
          const int k = w >> 1, s = k % nst;
          const uint32_t st = sS + (uint32_t)s * sbytes;
          if ((w & 1) == 0) {
            if (k >= nst) tc::mbar_wait(&bar.empty[s], (uint32_t)((k / nst - 1) & 1));
            wstore_k(st, g, r.w);
          } else {
            if constexpr (Prod::kChMajor) prod.store(g, kProdThreads, m0, k, r.a, st + kPart, 64, -64 * k);
            else prod.store64(g, k, r.a, st + kPart);
            tc::fence_proxy_async();
            mbar_arrive(&bar.full[s]);
          }
        });
  } else {   // warp 16: MMA issue, warp-uniform loop, one elected lane issues
    const uint32_t tmem = tc::uniform_u32(tmem_base);
    const uint32_t idesc = tc::make_idesc_bf16(128, kPts, false, Prod::kChMajor);
    for (int k = 0; k < nk; ++k) {
      const int s = k % nst;
      tc::mbar_wait(&bar.full[s], (uint32_t)((k / nst) & 1));
      tc::fence_after_sync();
      const uint32_t sA = sS + (uint32_t)s * sbytes, sB = sA + kPart;
      int kk = 64;
      if constexpr (!Prod::kChMajor) kk = prod.block_k(k);
      for (int q = 0; q < kk / 16; ++q)
        tc::mma_bf16_warp(tmem, tc::make_desc_sw128(sA + (uint32_t)q * 32, 16, 1024),
                     Prod::kChMajor ? tc::make_desc_sw128(sB + (uint32_t)q * 2048, 8192, 1024)
                                    : tc::make_desc_sw128(sB + (uint32_t)q * 32, 16, 1024),
                     idesc, k > 0 || q > 0);
      tc::mma_commit_warp(&bar.empty[s]);
    }
    tc::mma_commit_warp(&bar.tmem_full);
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc<128>(tmem);
}

// ---------------------------------------------------------------------------------------------
// dgrad: one (tile, 128-channel block of the layer INPUT) per CTA; contraction over the layer's
// output channels in chunks of 64.  Stage = [W part: 64 k rows x 128 input channels][dy part].
// PT: operands swapped, D[128 points x 128 input channels], point-on-lane epilogue (layer-1 scatter).
// ---------------------------------------------------------------------------------------------
template <class PProd, class Epi, bool PT>
__global__ void __launch_bounds__(kThreads, 1)
tc5_dgrad_kernel(PProd pp, const __nv_bfloat16* __restrict__ Wb, int Kp, Epi epi, int M, int nst) {
  PCOE_V5_PROLOGUE(nst)
  const int tile = blockIdx.x, cb = blockIdx.y, m0 = tile * kPts;
  const uint32_t sS = smem0, sbytes = 2 * kPart;
  float* csm = reinterpret_cast<float*>(smem_gen + (size_t)nst * sbytes);
  pp.init(csm, tid, kThreads);
  const int eq = warp & 3, eh = (warp >> 2) & 1;
  epi.init(csm + pp.nconst(), cb * 128 + eq * 32 + lane);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = tmem_base;
  const int nk = pp.C / 64;

  if (warp < 8) {
    tc::mbar_wait(&bar.tmem_full, 0u);
    tc::fence_after_sync();
    if constexpr (!PT) {
      const uint32_t stg = sS;
#pragma unroll 1
      for (int j = eh * 2; j < eh * 2 + 2; ++j) {
        float v[32];
        tc::tmem_ld32(tmem + ((uint32_t)(eq * 32) << 16) + (uint32_t)(j * 32), v);
        epi.block(v, tile, j, m0 + j * 32 < M, stg - (uint32_t)cb * 32768u);
      }
      tc::fence_before_sync();
      stage_copy_out(epi.tile_dst(tile) + (size_t)cb * 32768, stg, 32768, 1, tid == 0);
      if (tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    } else {
      const int row = m0 + eq * 32 + lane;
#pragma unroll 1
      for (int cbk = eh * 32; cbk < 128; cbk += 64) {
        float v[32];
        tc::tmem_ld32(tmem + ((uint32_t)(eq * 32) << 16) + (uint32_t)cbk, v);
        epi.block_pt(v, cb * 128 + cbk, row, row < M);
      }
      tc::fence_before_sync();
    }
    epi.finish();
  } else if (warp < 16) {
    const int g = tid - kEpiThreads;
    union RawU { typename PProd::Raw a; uint4 w[4]; __device__ RawU() {} };
    unit_pipeline<RawU>(2 * nk,
        [&](int w, RawU& r) {
          const int k = w >> 1;
          if ((w & 1) == 0) wload_mn(Wb, Kp, k * 64, cb * 128, g, r.w);
          else pp.load(g, kProdThreads, m0, k, NoIdx{}, r.a);
        },
        [&](int w, const RawU& r) {
          const int k = w >> 1, s = k % nst;
          const uint32_t st = sS + (uint32_t)s * sbytes;
          if ((w & 1) == 0) {
            if (k >= nst) tc::mbar_wait(&bar.empty[s], (uint32_t)((k / nst - 1) & 1));
            wstore_mn(st, g, r.w);
          } else {
            pp.store(g, kProdThreads, m0, k, r.a, st + kPart, 64, -64 * k);
            tc::fence_proxy_async();
            mbar_arrive(&bar.full[s]);
          }
        });
  } else {   // warp 16: MMA issue, warp-uniform loop, one elected lane issues
    const uint32_t tmem = tc::uniform_u32(tmem_base);
    const uint32_t idesc = tc::make_idesc_bf16(128, 128, true, true);
    for (int k = 0; k < nk; ++k) {
      const int s = k % nst;
      tc::mbar_wait(&bar.full[s], (uint32_t)((k / nst) & 1));
      tc::fence_after_sync();
      const uint32_t sWp = sS + (uint32_t)s * sbytes, sPp = sWp + kPart;
      for (int q = 0; q < 4; ++q) {
        const uint64_t dw = tc::make_desc_sw128(sWp + (uint32_t)q * 2048, 8192, 1024);
        const uint64_t dp = tc::make_desc_sw128(sPp + (uint32_t)q * 2048, 8192, 1024);
        tc::mma_bf16_warp(tmem, PT ? dp : dw, PT ? dw : dp, idesc, k > 0 || q > 0);
      }
      tc::mma_commit_warp(&bar.empty[s]);
    }
    tc::mma_commit_warp(&bar.tmem_full);
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc<128>(tmem);
}

// ---------------------------------------------------------------------------------------------
// wgrad: CTA (cl block, q block, split) accumulates dW[128 x nq] over its tiles in TMEM and adds it
// to global memory at the end.  Stage = [P: 128 ch x 128 pts][Q: 128 ch x 128 pts (channel-major) or
// 128 pts x 2 blocks of 64 ch (point-major, layer 1)], four 64-channel work units per tile.
// ---------------------------------------------------------------------------------------------
template <class PProd, class QProd>
__global__ void __launch_bounds__(kThreads, 1)
tc5_wgrad_kernel(PProd pp, QProd qp, float* __restrict__ dW, int ldo, int cq_valid, int perm_d, int M, int tps, int nst) {
  PCOE_V5_PROLOGUE(nst)
  const int cl0 = blockIdx.x * 128, qb = blockIdx.y;
  const int ntiles = (M + kPts - 1) / kPts;
  const int t0 = blockIdx.z * tps, t1 = min(ntiles, t0 + tps), nt = t1 - t0;
  const uint32_t sS = smem0, sbytes = 4 * kPart;
  float* csm = reinterpret_cast<float*>(smem_gen + (size_t)nst * sbytes);
  pp.init(csm, tid, kThreads);
  qp.init(csm + pp.nconst(), tid, kThreads);
  int nq = 128, nqu = 2;                                     // columns of this q block, 64-channel units of Q
  if constexpr (!QProd::kChMajor) {
    const int nb = qp.nblocks();
    nqu = min(2, nb - 2 * qb);
    nq = 0;
    for (int u = 0; u < nqu; ++u) nq += qp.block_k(2 * qb + u);
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = tmem_base;
  const int eq = warp & 3, eh = (warp >> 2) & 1;
  const int upt = 2 + nqu;                                   // work units per tile

  if (warp < 8) {
    if (nt > 0) {
      tc::mbar_wait(&bar.tmem_full, 0u);
      tc::fence_after_sync();
      const int crow = cl0 + eq * 32 + lane;
#pragma unroll 1
      for (int cbk = eh * 32; cbk < nq; cbk += 64) {
        float v[32];
        tc::tmem_ld32(tmem + ((uint32_t)(eq * 32) << 16) + (uint32_t)cbk, v);   // may read past nq: unused columns
        float* dst = dW + (size_t)crow * ldo;
        const int cb = qb * 128 + cbk;
        if (perm_d < 0 && (ldo & 3) == 0 && cb + 32 <= cq_valid) {
#pragma unroll
          for (int q = 0; q < 8; ++q) red_add_v4(dst + cb + 4 * q, v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
        } else {
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            int c = cb + e;
            if (c >= cq_valid || cbk + e >= nq) continue;
            if (perm_d >= 0) c = c < perm_d ? c + 3 : c - perm_d;   // [feats | xyz] -> [xyz | feats]
            atomicAdd(dst + c, v[e]);
          }
        }
      }
      tc::fence_before_sync();
    }
  } else if (warp < 16) {
    const int g = tid - kEpiThreads;
    union RawU { typename PProd::Raw p; typename QProd::Raw q; __device__ RawU() {} };
    struct Cur { int i, u, m0; };   // tile ordinal, unit, first row: advanced incrementally (no division)
    auto adv = [&](Cur& c) { if (++c.u == upt) { c.u = 0; ++c.i; c.m0 += kPts; } };
    Cur cl{0, 0, t0 * kPts}, cst = cl;
    int ring_s = 0, ring_n = 0;
    unit_pipeline<RawU>(nt * upt,
        [&](int, RawU& r) {
          const int u = cl.u, m0 = cl.m0;
          if (u < 2) pp.load(g, kProdThreads, m0, (cl0 >> 6) + u, NoIdx{}, r.p);
          else if constexpr (QProd::kChMajor) qp.load(g, kProdThreads, m0, 2 * qb + (u - 2), NoIdx{}, r.q);
          else qp.load64(g, m0, 2 * qb + (u - 2), r.q);
          adv(cl);
        },
        [&](int, const RawU& r) {
          const int u = cst.u, m0 = cst.m0, s = ring_s;
          const uint32_t st = sS + (uint32_t)s * sbytes;
          if (u == 0 && ring_n > 0) tc::mbar_wait(&bar.empty[s], (uint32_t)((ring_n - 1) & 1));
          if (u < 2) pp.store(g, kProdThreads, m0, (cl0 >> 6) + u, r.p, st, 128, -cl0);
          else if constexpr (QProd::kChMajor) qp.store(g, kProdThreads, m0, 2 * qb + (u - 2), r.q, st + 2 * kPart, 128, -128 * qb);
          else qp.store64(g, 2 * qb + (u - 2), r.q, st + 2 * kPart + (uint32_t)(u - 2) * kPart);
          if (u == upt - 1) {
            tc::fence_proxy_async();
            mbar_arrive(&bar.full[s]);
            if (++ring_s == nst) { ring_s = 0; ++ring_n; }
          }
          adv(cst);
        });
  } else {   // warp 16: MMA issue, warp-uniform loop, one elected lane issues
    const uint32_t tmem = tc::uniform_u32(tmem_base);
    const uint32_t idesc = tc::make_idesc_bf16(128, nq, false, !QProd::kChMajor);
    for (int i = 0; i < nt; ++i) {
      const int s = i % nst;
      tc::mbar_wait(&bar.full[s], (uint32_t)((i / nst) & 1));
      tc::fence_after_sync();
      const uint32_t sP = sS + (uint32_t)s * sbytes, sQ = sP + 2 * kPart;
      for (int ks = 0; ks < 8; ++ks) {                          // 16 points per MMA
        const uint64_t ad = tc::make_desc_sw128(sP + (uint32_t)(ks >> 2) * kPart + (uint32_t)((ks & 3) * 32), 16, 1024);
        const uint64_t bd = QProd::kChMajor
            ? tc::make_desc_sw128(sQ + (uint32_t)(ks >> 2) * kPart + (uint32_t)((ks & 3) * 32), 16, 1024)
            : tc::make_desc_sw128(sQ + (uint32_t)ks * 2048, kPart, 1024);
        tc::mma_bf16_warp(tmem, ad, bd, idesc, i > 0 || ks > 0);
      }
      tc::mma_commit_warp(&bar.empty[s]);
    }
    if (nt > 0) tc::mma_commit_warp(&bar.tmem_full);
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc<128>(tmem);
}

}  // namespace v5
}  // namespace pcoe
