// sa_tc3.cuh — warp-specialised versions of the persistent tcgen05 kernels of sa_tc2.cuh (same
// producers, same register epilogues, same operand layouts).
//
// CTA = 12 warps:  warps 0-3  EPILOGUE  (warp w owns TMEM lanes 32w..32w+31 = rows 32w.. of the tile)
//                  warps 4-11 PRODUCERS (256 threads build the bf16 operand tiles; thread 128 also
//                                        issues the tcgen05.mma chain)
// Shared-memory operand tiles are single-buffered (the MMA chain of a tile takes well under a
// microsecond), the TMEM accumulator that the epilogue drains is double-buffered, so producers load
// and convert tile t+1 from HBM while the epilogue warps are still reducing / storing tile t:
//
//   producers : wait smem_free(t-1) | produce(t) | fence.proxy.async | bar.sync(producers)
//   thread 128: wait tmem_empty[b](t-2) | tcgen05.mma ... | commit -> smem_free, commit -> tmem_full[b]
//   epilogue  : wait tmem_full[b](t) | tcgen05.ld + epilogue | arrive tmem_empty[b]
#pragma once
#include "sa_tc2.cuh"

namespace pcoe {
namespace v3 {

constexpr int kEpiThreads = 128, kProdThreads = 256, kCtaThreads = kEpiThreads + kProdThreads;

__device__ __forceinline__ void mbar_arrive(uint64_t* mbar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc::smem_u32(mbar)) : "memory");
}
__device__ __forceinline__ void producers_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// all rows of the thread's channel unit are loaded before any arithmetic (ROWS <= 16 independent
// 16/32-byte loads in flight per thread)
template <class Prod, int UNITS>
__device__ __forceinline__ void produce_tile_u(const Prod& p, int ptid, int m0, uint8_t* tiles) {
  constexpr int RSTEP = kProdThreads / UNITS, ROWS = 128 / RSTEP, BATCH = ROWS < 8 ? ROWS : 8;
  const int j = ptid % UNITS, r0 = ptid / UNITS;
  const uint32_t dst = tc::smem_u32(tiles) + (uint32_t)(j >> 3) * (128 * 128);
#pragma unroll
  for (int b = 0; b < ROWS; b += BATCH) {
    float v[BATCH][8];
#pragma unroll
    for (int i = 0; i < BATCH; ++i) p.load8(m0 + r0 + (b + i) * RSTEP, v[i]);
#pragma unroll
    for (int i = 0; i < BATCH; ++i)
      tc::sts128(dst + tc::sw128_off(r0 + (b + i) * RSTEP, (j & 7) * 8), tc::pack8_bf16(v[i]));
  }
}

template <class Prod>
__device__ __forceinline__ void produce_tile(const Prod& p, int units, int ptid, int m0, uint8_t* tiles) {
  if (units == 8) produce_tile_u<Prod, 8>(p, ptid, m0, tiles);
  else if (units == 16) produce_tile_u<Prod, 16>(p, ptid, m0, tiles);
  else produce_tile_u<Prod, 32>(p, ptid, m0, tiles);
}

__device__ __forceinline__ void load_weights(const __nv_bfloat16* __restrict__ W, int ldw, int nrows, int kpad,
                                             uint8_t* tiles, int tid, int nthr) {
  const int upr = kpad / 8;
  for (int e = tid; e < nrows * upr; e += nthr) {
    const int n = e / upr, j = e % upr;
    const uint4 w = __ldg(reinterpret_cast<const uint4*>(W + (size_t)n * ldw + j * 8));
    tc::sts128(tc::smem_u32(tiles) + (uint32_t)(j >> 3) * (uint32_t)(nrows * 128) + tc::sw128_off(n, (j & 7) * 8), w);
  }
}

// epilogue warp w (0..3): rows 32w.. of the tile, all columns in blocks of 32
template <class Epi>
__device__ __forceinline__ void drain3(Epi& epi, uint32_t tmem, int ncols, int m0, int M) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = m0 + warp * 32 + lane;
  const bool valid = row < M;
  int blk = 0;
  for (int cb = 0; cb < ncols; cb += 32, ++blk) {
    float v[32];
    tc::tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)cb, v);
    epi.block(v, row, valid, cb, blk, lane);
  }
}

struct Barriers {
  uint64_t smem_free;
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
};

// ---------------------------------------------------------------------------------------------
// forward layer (see tc2_fwd_kernel for the operand layout).  TMEM: two accumulators of nacc
// columns each (nacc = N rounded up to 32).
// ---------------------------------------------------------------------------------------------
template <class Prod, class Epi, int TCOLS>
__global__ void __launch_bounds__(kCtaThreads, 1)
tc3_fwd_kernel(Prod prod, const __nv_bfloat16* __restrict__ Wb, int ldw, int wrows, Epi epi, int M, int ncols,
               int kpad, int kmma) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
  const int kt = kpad / 64;
  uint8_t* sA = smem;
  uint8_t* sW = smem + (size_t)kt * 128 * 128;
  __shared__ Barriers bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool is_epi = tid < kEpiThreads;
  const int ptid = tid - kEpiThreads;
  const int nmma = (ncols + 15) / 16 * 16;
  const int nacc = (ncols + 31) / 32 * 32;
  const int wr = min(wrows, (nmma + 7) / 8 * 8);

  if (warp == 0) tc::tmem_alloc<TCOLS>(&tmem_base);
  if (tid == 0) {
    tc::mbar_init(&bar.smem_free, 1);
    tc::mbar_init(&bar.tmem_full[0], 1);
    tc::mbar_init(&bar.tmem_full[1], 1);
    tc::mbar_init(&bar.tmem_empty[0], kEpiThreads);
    tc::mbar_init(&bar.tmem_empty[1], kEpiThreads);
  }
  load_weights(Wb, ldw, wr, kpad, sW, tid, kCtaThreads);
  if (!is_epi) prod.bind((ptid % (kpad / 8)) * 8);
  epi.init(reinterpret_cast<float*>(sW + (size_t)kt * wr * 128));
  tc::fence_proxy_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = tmem_base;
  const int ntiles = (M + 127) / 128;

  if (!is_epi) {
    const uint32_t idesc = tc::make_idesc_bf16(128, nmma, false, false);
    const uint32_t a0 = tc::smem_u32(sA), w0 = tc::smem_u32(sW);
    int i = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++i) {
      if (i > 0) tc::mbar_wait(&bar.smem_free, (uint32_t)((i - 1) & 1));   // MMAs of tile i-1 finished reading smem
      produce_tile(prod, kpad / 8, ptid, tile * 128, sA);
      tc::fence_proxy_async();
      producers_sync();
      if (ptid == 0) {
        const int b = i & 1, u = i >> 1;
        if (u > 0) tc::mbar_wait(&bar.tmem_empty[b], (uint32_t)((u - 1) & 1));   // epilogue drained tile i-2
        tc::fence_after_sync();
        for (int k = 0; k < kmma; k += 16) {
          const uint32_t off = (uint32_t)(k >> 6), sub = (uint32_t)((k & 63) * 2);
          tc::mma_bf16(tmem + (uint32_t)(b * nacc), tc::make_desc_sw128(a0 + off * (128 * 128) + sub, 16, 1024),
                       tc::make_desc_sw128(w0 + off * (uint32_t)(wr * 128) + sub, 16, 1024), idesc, k > 0);
        }
        tc::mma_commit(&bar.smem_free);
        tc::mma_commit(&bar.tmem_full[b]);
      }
    }
  } else {
    int i = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++i) {
      const int b = i & 1, u = i >> 1;
      tc::mbar_wait(&bar.tmem_full[b], (uint32_t)(u & 1));
      tc::fence_after_sync();
      drain3(epi, tmem + (uint32_t)(b * nacc), ncols, tile * 128, M);
      tc::fence_before_sync();
      mbar_arrive(&bar.tmem_empty[b]);
    }
    epi.finish(lane);
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc<TCOLS>(tmem);
}

// ---------------------------------------------------------------------------------------------
// fused backward of one layer (see tc2_bwd_kernel).  TMEM: [dW: mt*cbmma columns, rounded to 64]
// [dx buffer 0: dacc columns][dx buffer 1: dacc columns]
// ---------------------------------------------------------------------------------------------
template <class PProd, class QProd, class Epi, int TCOLS>
__global__ void __launch_bounds__(kCtaThreads, 1)
tc3_bwd_kernel(PProd pp, QProd qp, const __nv_bfloat16* __restrict__ WbT, int ldw, int wrows, Epi epi,
               float* __restrict__ dW, int ldo, int cb_valid, int perm_d, int M, int ca, int cbk, int cbmma,
               int cdn) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
  const int pt = ca / 64, qt = cbk / 64;
  uint8_t* sP = smem;
  uint8_t* sQ = sP + (size_t)pt * 128 * 128;
  uint8_t* sW = sQ + (size_t)qt * 128 * 128;
  __shared__ Barriers bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool is_epi = tid < kEpiThreads;
  const int ptid = tid - kEpiThreads;
  const int mt = (ca + 127) / 128;
  const int dnm = (cdn + 15) / 16 * 16, dacc = (cdn + 31) / 32 * 32;
  const int wr = cdn ? min(wrows, (dnm + 7) / 8 * 8) : 0;
  const uint32_t dx_col = (uint32_t)((mt * cbmma + 63) / 64 * 64);

  if (warp == 0) tc::tmem_alloc<TCOLS>(&tmem_base);
  if (tid == 0) {
    tc::mbar_init(&bar.smem_free, 1);
    tc::mbar_init(&bar.tmem_full[0], 1);
    tc::mbar_init(&bar.tmem_full[1], 1);
    tc::mbar_init(&bar.tmem_empty[0], kEpiThreads);
    tc::mbar_init(&bar.tmem_empty[1], kEpiThreads);
  }
  if (cdn) load_weights(WbT, ldw, wr, ca, sW, tid, kCtaThreads);
  if (!is_epi) {
    pp.bind((ptid % (ca / 8)) * 8);
    qp.bind((ptid % (cbk / 8)) * 8);
  }
  epi.init(reinterpret_cast<float*>(sW + (size_t)pt * wr * 128));
  tc::fence_proxy_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = tmem_base;
  const int ntiles = (M + 127) / 128;
  int my_tiles = 0;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) ++my_tiles;

  if (!is_epi) {
    const uint32_t idesc_w = tc::make_idesc_bf16(128, cbmma, true, true);
    const uint32_t idesc_d = tc::make_idesc_bf16(128, dnm ? dnm : 16, false, false);
    const uint32_t p0 = tc::smem_u32(sP), q0 = tc::smem_u32(sQ), w0 = tc::smem_u32(sW);
    int i = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++i) {
      if (i > 0) tc::mbar_wait(&bar.smem_free, (uint32_t)((i - 1) & 1));
      produce_tile(pp, ca / 8, ptid, tile * 128, sP);
      produce_tile(qp, cbk / 8, ptid, tile * 128, sQ);
      tc::fence_proxy_async();
      producers_sync();
      if (ptid == 0) {
        const int b = i & 1, u = i >> 1;
        if (u > 0) tc::mbar_wait(&bar.tmem_empty[b], (uint32_t)((u - 1) & 1));
        tc::fence_after_sync();
        for (int mi = 0; mi < mt; ++mi)
          for (int ks = 0; ks < 8; ++ks)
            tc::mma_bf16(tmem + (uint32_t)(mi * cbmma),
                         tc::make_desc_sw128(p0 + (uint32_t)(mi * 2) * (128 * 128) + ks * 2048, 128 * 128, 1024),
                         tc::make_desc_sw128(q0 + ks * 2048, 128 * 128, 1024), idesc_w, i > 0 || ks > 0);
        if (cdn)
          for (int k = 0; k < ca; k += 16) {
            const uint32_t off = (uint32_t)(k >> 6), sub = (uint32_t)((k & 63) * 2);
            tc::mma_bf16(tmem + dx_col + (uint32_t)(b * dacc), tc::make_desc_sw128(p0 + off * (128 * 128) + sub, 16, 1024),
                         tc::make_desc_sw128(w0 + off * (uint32_t)(wr * 128) + sub, 16, 1024), idesc_d, k > 0);
          }
        tc::mma_commit(&bar.smem_free);
        tc::mma_commit(&bar.tmem_full[b]);
      }
    }
  } else {
    int i = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++i) {
      const int b = i & 1, u = i >> 1;
      tc::mbar_wait(&bar.tmem_full[b], (uint32_t)(u & 1));
      tc::fence_after_sync();
      if (cdn) drain3(epi, tmem + dx_col + (uint32_t)(b * dacc), cdn, tile * 128, M);
      tc::fence_before_sync();
      mbar_arrive(&bar.tmem_empty[b]);
    }
    epi.finish(lane);
    // flush dW (complete once the last tile's commit has been observed above)
    if (my_tiles > 0) {
      tc::fence_after_sync();
      for (int mi = 0; mi < mt; ++mi) {
        const int crow = mi * 128 + warp * 32 + lane;
        for (int cb = 0; cb < cbmma; cb += 32) {
          float v[32];
          tc::tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(mi * cbmma + cb), v);
          if (crow < ca) {
#pragma unroll
            for (int k = 0; k < 32; ++k) {
              int c = cb + k;
              if (c >= cb_valid) continue;
              if (perm_d >= 0) c = c < perm_d ? c + 3 : c - perm_d;
              atomicAdd(dW + (size_t)crow * ldo + c, v[k]);
            }
          }
        }
      }
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc<TCOLS>(tmem);
}

}  // namespace v3
}  // namespace pcoe
