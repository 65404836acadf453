// sa_tc2.cuh — second-generation tcgen05 kernels for the set-abstraction MLP (bf16 mode), used
// whenever the layer fits on chip (K_group == 32, channel counts multiples of 64 up to 256 — i.e.
// SA1 and SA2 of every reference model); other shapes fall back to sa_tc.cuh.
//
//   tc2_fwd_kernel  persistent CTAs; the layer's weights stay resident in shared memory as
//                   SWIZZLE_128B K-major tiles; per 128-row tile: producers build the bf16 A tile,
//                   one thread issues the tcgen05.mma chain (M=128, N=Cout<=256), the epilogue runs
//                   out of registers straight from TMEM (tcgen05.ld 32x32b): bf16 row stores,
//                   batch statistics and max/min/arg over the 32 neighbours by butterfly
//                   transpose-reductions across the warp (a warp's 32 TMEM lanes are exactly one
//                   group of K=32 neighbours).  Channel sums live in registers for the whole
//                   kernel: one fp64 atomic per channel per warp at the very end.
//   tc2_bwd_kernel  fused weight- and data-gradient of one layer: the dy tile (BatchNorm backward
//                   applied on load) and the x_prev tile are produced once and consumed twice —
//                   as MN-major operands for dW += dy^T x (accumulated in TMEM over the CTA's whole
//                   row range) and dy as the K-major A operand of dx = dy W (W^T resident in smem);
//                   the dx epilogue applies the previous layer's ReLU mask, accumulates the
//                   BatchNorm-backward sums and stores dz_prev (or scatter-adds with 16-byte vector
//                   atomics for layer 1).
#pragma once
#include "sa_common.cuh"
#include "tc_common.cuh"

namespace pcoe {
namespace v2 {

constexpr int kThreads = 256;

// lane L returns sum over the warp's 32 lanes of v[L] (butterfly transpose-reduce, 31 shuffles)
__device__ __forceinline__ float colsum32(float (&v)[32], int lane) {
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
    const bool up = (lane & s) != 0;
#pragma unroll
    for (int i = 0; i < s; ++i) {
      const float keep = up ? v[i + s] : v[i];
      const float send = up ? v[i] : v[i + s];
      v[i] = keep + __shfl_xor_sync(0xFFFFFFFFu, send, s);
    }
  }
  return v[0];
}

// lane L returns (extreme value, lowest row index attaining it) of column L; MAX=true: maximum
template <bool MAX>
__device__ __forceinline__ void colext32(float (&v)[32], int lane, float* val, int* arg) {
  int ix[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) ix[i] = lane;
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
    const bool up = (lane & s) != 0;
#pragma unroll
    for (int i = 0; i < s; ++i) {
      const float kv = up ? v[i + s] : v[i];
      const float sv = up ? v[i] : v[i + s];
      const int ki = up ? ix[i + s] : ix[i];
      const int si = up ? ix[i] : ix[i + s];
      const float pv = __shfl_xor_sync(0xFFFFFFFFu, sv, s);
      const int pi = __shfl_xor_sync(0xFFFFFFFFu, si, s);
      const bool take = MAX ? (pv > kv || (pv == kv && pi < ki)) : (pv < kv || (pv == kv && pi < ki));
      v[i] = take ? pv : kv;
      ix[i] = take ? pi : ki;
    }
  }
  *val = v[0];
  *arg = ix[0];
}

__device__ __forceinline__ void store32_bf16(__nv_bfloat16* dst, const float (&v)[32]) {
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float t[8] = {v[8 * q], v[8 * q + 1], v[8 * q + 2], v[8 * q + 3], v[8 * q + 4], v[8 * q + 5], v[8 * q + 6], v[8 * q + 7]};
    reinterpret_cast<uint4*>(dst)[q] = tc::pack8_bf16(t);
  }
}

__device__ __forceinline__ void load32_bf16(const __nv_bfloat16* src, float (&v)[32]) {
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const uint4 a = __ldg(reinterpret_cast<const uint4*>(src) + q);
    const uint32_t w[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      v[8 * q + 2 * u] = __uint_as_float(w[u] << 16);
      v[8 * q + 2 * u + 1] = __uint_as_float(w[u] & 0xFFFF0000u);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Producers with per-thread constants.  bind(c0) fixes the 8 channels this thread produces for
// the whole kernel; load8(row, v) returns them for one row (zeros outside [0,M) x [0,C)).
// ---------------------------------------------------------------------------------------------

// layer-1 input in "features first" channel order: [feats(D) | xyz - centroid (3) | 0...]
struct Gather2 {
  const float* __restrict__ xyz;
  const float* __restrict__ new_xyz;
  const int32_t* __restrict__ nbr;
  const float* __restrict__ feats;
  int N, S, K, D, group_all, M;
  int c0;
  __device__ __forceinline__ void bind(int c) { c0 = c; }
  __device__ __forceinline__ void load8(int row, float (&v)[8]) const {
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = 0.f;
    if (c0 >= D + 3) return;                 // per-thread constant: this thread only writes padding
    const bool ok = row < M;
    const int rr = min(row, M - 1);
    const int g = rr >> 5;                   // K == 32 on this path
    int pt = rr;
    if (!group_all) {
      int i = __ldg(nbr + rr);
      i = min(max(i, 0), N - 1);
      pt = (g / S) * N + i;
    }
    if (c0 + 8 <= D) {   // D is a multiple of 8 here: aligned 32-byte feature read
      const float4 a = __ldg(reinterpret_cast<const float4*>(feats + (size_t)pt * D + c0));
      const float4 b = __ldg(reinterpret_cast<const float4*>(feats + (size_t)pt * D + c0) + 1);
      v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else {
#pragma unroll
      for (int u = 0; u < 3; ++u) {          // c0 == D: the three centred coordinates
        float x = __ldg(xyz + (size_t)pt * 3 + u);
        if (!group_all) x = __fsub_rn(x, __ldg(new_xyz + (size_t)g * 3 + u));
        v[u] = x;
      }
    }
    if (!ok) {
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = 0.f;
    }
  }
};

struct BnRelu2 {
  const __nv_bfloat16* __restrict__ y;
  const float* __restrict__ scale;
  const float* __restrict__ shift;
  int M, C;
  int c0;
  float sc[8], sh[8];
  __device__ __forceinline__ void bind(int c) {
    c0 = c;
#pragma unroll
    for (int u = 0; u < 8; ++u) { sc[u] = c + u < C ? scale[c + u] : 0.f; sh[u] = c + u < C ? shift[c + u] : 0.f; }
  }
  // branch-free (clamped address + select) so that the loads of a batch can be hoisted together
  __device__ __forceinline__ void load8(int row, float (&v)[8]) const {
    const bool ok = row < M && c0 < C;
    load_row8<__nv_bfloat16>(y, (size_t)min(row, M - 1) * C + min(c0, C - 8), v);
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = ok ? fmaxf(fmaf(v[u], sc[u], sh[u]), 0.f) : 0.f;
  }
};

struct Dy2 {   // a*dz + p*y + q, dense upstream dz
  const __nv_bfloat16* __restrict__ dz;
  const __nv_bfloat16* __restrict__ y;
  const float* __restrict__ a;
  const float* __restrict__ p;
  const float* __restrict__ q;
  int M, C;
  int c0;
  float ca[8], cp[8], cq[8];
  __device__ __forceinline__ void bind(int c) {
    c0 = c;
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const bool ok = c + u < C;
      ca[u] = ok ? a[c + u] : 0.f; cp[u] = ok ? p[c + u] : 0.f; cq[u] = ok ? q[c + u] : 0.f;
    }
  }
  __device__ __forceinline__ void load8(int row, float (&v)[8]) const {
    const bool ok = row < M && c0 < C;
    const size_t off = (size_t)min(row, M - 1) * C + min(c0, C - 8);
    float d[8], yy[8];
    load_row8<__nv_bfloat16>(dz, off, d);
    load_row8<__nv_bfloat16>(y, off, yy);
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = ok ? fmaf(ca[u], d[u], fmaf(cp[u], yy[u], cq[u])) : 0.f;
  }
};

struct DyLast2 {   // upstream = max-pool routing of gm[G,C] to the saved slot; K == 32
  const float* __restrict__ gm;
  const uint8_t* __restrict__ slot;
  const __nv_bfloat16* __restrict__ y;
  const float* __restrict__ a;
  const float* __restrict__ p;
  const float* __restrict__ q;
  int M, C;
  int c0;
  float ca[8], cp[8], cq[8];
  __device__ __forceinline__ void bind(int c) {
    c0 = c;
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const bool ok = c + u < C;
      ca[u] = ok ? a[c + u] : 0.f; cp[u] = ok ? p[c + u] : 0.f; cq[u] = ok ? q[c + u] : 0.f;
    }
  }
  __device__ __forceinline__ void load8(int row, float (&v)[8]) const {
    const bool ok = row < M && c0 < C;
    const int rr = min(row, M - 1), cc = min(c0, C - 8);
    const int g = rr >> 5, j = rr & 31;
    float yy[8], gg[8];
    load_row8<__nv_bfloat16>(y, (size_t)rr * C + cc, yy);
    load_row8<float>(gm, (size_t)g * C + cc, gg);
    const uint2 sl = __ldg(reinterpret_cast<const uint2*>(slot + (size_t)g * C + cc));
    const uint32_t sw[2] = {sl.x, sl.y};
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int s = (sw[u >> 2] >> (8 * (u & 3))) & 0xFF;
      v[u] = ok ? fmaf(ca[u], (s == j) ? gg[u] : 0.f, fmaf(cp[u], yy[u], cq[u])) : 0.f;
    }
  }
};

// Build one [128 rows x (8*units) channels] operand as `units/8` SWIZZLE_128B tiles.
// Thread t owns channel unit (t % units) for the whole kernel (constants bound once) and walks the
// rows t/units, t/units + 256/units, ...: global reads are contiguous across the warp, shared
// stores hit 4 distinct 128-byte rows per warp (conflict free).
// Rows are processed in batches of 4 with all loads of a batch issued before any arithmetic, so
// every thread keeps >= 4 independent 16/32-byte global loads in flight (memory-level parallelism).
template <class Prod, int UNITS>
__device__ __forceinline__ void produce_tile_u(const Prod& p, int m0, uint8_t* tiles) {
  constexpr int RSTEP = kThreads / UNITS, ROWS = 128 / RSTEP, BATCH = 4;
  const int j = threadIdx.x % UNITS, r0 = threadIdx.x / UNITS;
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
__device__ __forceinline__ void produce_tile(const Prod& p, int units, int m0, uint8_t* tiles) {
  if (units == 8) produce_tile_u<Prod, 8>(p, m0, tiles);
  else if (units == 16) produce_tile_u<Prod, 16>(p, m0, tiles);
  else produce_tile_u<Prod, 32>(p, m0, tiles);
}

// resident weight operand: rows [0,nrows) x kpad bf16 from a padded row-major bf16 matrix
__device__ __forceinline__ void load_weights(const __nv_bfloat16* __restrict__ W, int ldw, int nrows, int kpad,
                                             uint8_t* tiles) {
  const int upr = kpad / 8;   // 16-byte units per row
  for (int e = threadIdx.x; e < nrows * upr; e += kThreads) {
    const int n = e / upr, j = e % upr;
    const uint4 w = __ldg(reinterpret_cast<const uint4*>(W + (size_t)n * ldw + j * 8));
    tc::sts128(tc::smem_u32(tiles) + (uint32_t)(j >> 3) * (uint32_t)(nrows * 128) + tc::sw128_off(n, (j & 7) * 8), w);
  }
}

// ---------------------------------------------------------------------------------------------
// Register epilogues.  block(v, row, valid, cb, blk, lane): one 32-row x 32-column accumulator block
// (thread = row, v = 32 consecutive columns starting at cb); blk = index of the block inside this
// warp's column range (for the per-thread running sums).  finish(): flush the running sums.
// ---------------------------------------------------------------------------------------------
constexpr int kMaxBlk = 8;   // a warp covers at most 256 columns (4 epilogue warps in sa_tc3.cuh)

struct StoreStats2 {
  __nv_bfloat16* __restrict__ y;
  double* __restrict__ sums;   // [2,C] or nullptr (eval)
  int C;
  float s0[kMaxBlk], s1[kMaxBlk];
  int cbs[kMaxBlk];
  __device__ __forceinline__ void init(float*) {
#pragma unroll
    for (int b = 0; b < kMaxBlk; ++b) { s0[b] = s1[b] = 0.f; cbs[b] = -1; }
  }
  __device__ __forceinline__ void block(float (&v)[32], int row, bool valid, int cb, int blk, int lane) {
    if (valid) store32_bf16(y + (size_t)row * C + cb, v);
    if (sums) {
      float w[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) { v[i] = valid ? v[i] : 0.f; w[i] = v[i] * v[i]; }
      const float a = colsum32(v, lane), b = colsum32(w, lane);
#pragma unroll
      for (int k = 0; k < kMaxBlk; ++k)
        if (k == blk) { s0[k] += a; s1[k] += b; cbs[k] = cb; }
    }
  }
  __device__ __forceinline__ void finish(int lane) {
    if (!sums) return;
#pragma unroll
    for (int k = 0; k < kMaxBlk; ++k)
      if (cbs[k] >= 0) {
        atomicAdd(sums + cbs[k] + lane, (double)s0[k]);
        atomicAdd(sums + C + cbs[k] + lane, (double)s1[k]);
      }
  }
};

struct Group2 {   // last layer, K == 32: a warp's 32 rows are one group
  __nv_bfloat16* __restrict__ y;   // or nullptr (eval)
  double* __restrict__ sums;       // or nullptr
  float* __restrict__ ymax;
  float* __restrict__ ymin;
  uint8_t* __restrict__ amax;
  uint8_t* __restrict__ amin;
  int C;
  float s0[kMaxBlk], s1[kMaxBlk];
  int cbs[kMaxBlk];
  __device__ __forceinline__ void init(float*) {
#pragma unroll
    for (int b = 0; b < kMaxBlk; ++b) { s0[b] = s1[b] = 0.f; cbs[b] = -1; }
  }
  __device__ __forceinline__ void block(float (&v)[32], int row, bool valid, int cb, int blk, int lane) {
    if (y && valid) store32_bf16(y + (size_t)row * C + cb, v);
    float w[32];
    if (sums) {
#pragma unroll
      for (int i = 0; i < 32; ++i) w[i] = valid ? v[i] : 0.f;
      float w2[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) w2[i] = w[i] * w[i];
      const float a = colsum32(w, lane), b = colsum32(w2, lane);
#pragma unroll
      for (int k = 0; k < kMaxBlk; ++k)
        if (k == blk) { s0[k] += a; s1[k] += b; cbs[k] = cb; }
    }
    // groups never straddle M (M is a multiple of K = 32): the whole warp is valid or not
    float mx, mn;
    int ax, an;
#pragma unroll
    for (int i = 0; i < 32; ++i) w[i] = v[i];
    colext32<true>(w, lane, &mx, &ax);
    colext32<false>(v, lane, &mn, &an);
    if (valid) {
      const size_t o = (size_t)(row >> 5) * C + cb + lane;
      ymax[o] = mx; ymin[o] = mn; amax[o] = (uint8_t)ax; amin[o] = (uint8_t)an;
    }
  }
  __device__ __forceinline__ void finish(int lane) {
    if (!sums) return;
#pragma unroll
    for (int k = 0; k < kMaxBlk; ++k)
      if (cbs[k] >= 0) {
        atomicAdd(sums + cbs[k] + lane, (double)s0[k]);
        atomicAdd(sums + C + cbs[k] + lane, (double)s1[k]);
      }
  }
};

struct MaskStats2 {   // dz_prev = dx * [z_prev > 0]; sums of dz_prev and dz_prev * xhat_prev
  const __nv_bfloat16* __restrict__ yprev;
  const float* __restrict__ scale;
  const float* __restrict__ shift;
  const float* __restrict__ mean;
  const float* __restrict__ invstd;
  __nv_bfloat16* __restrict__ dz;
  double* __restrict__ sums;
  int C;
  float s0[kMaxBlk], s1[kMaxBlk];
  int cbs[kMaxBlk];
  const float* k_sc; const float* k_sh; const float* k_mu; const float* k_is;   // shared-memory copies
  __device__ __forceinline__ void init(float* scratch) {
#pragma unroll
    for (int b = 0; b < kMaxBlk; ++b) { s0[b] = s1[b] = 0.f; cbs[b] = -1; }
    for (int c = threadIdx.x; c < C; c += kThreads) {
      scratch[c] = scale[c]; scratch[C + c] = shift[c]; scratch[2 * C + c] = mean[c]; scratch[3 * C + c] = invstd[c];
    }
    k_sc = scratch; k_sh = scratch + C; k_mu = scratch + 2 * C; k_is = scratch + 3 * C;
  }
  __device__ __forceinline__ void block(float (&v)[32], int row, bool valid, int cb, int blk, int lane) {
    float yv[32], w[32];
    if (valid) load32_bf16(yprev + (size_t)row * C + cb, yv);
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const float yy = valid ? yv[i] : 0.f;
      const bool on = valid && fmaf(yy, k_sc[cb + i], k_sh[cb + i]) > 0.f;
      v[i] = on ? v[i] : 0.f;
      w[i] = v[i] * ((yy - k_mu[cb + i]) * k_is[cb + i]);
    }
    if (valid) store32_bf16(dz + (size_t)row * C + cb, v);
    const float a = colsum32(v, lane), b = colsum32(w, lane);
#pragma unroll
    for (int k = 0; k < kMaxBlk; ++k)
      if (k == blk) { s0[k] += a; s1[k] += b; cbs[k] = cb; }
  }
  __device__ __forceinline__ void finish(int lane) {
#pragma unroll
    for (int k = 0; k < kMaxBlk; ++k)
      if (cbs[k] >= 0) {
        atomicAdd(sums + cbs[k] + lane, (double)s0[k]);
        atomicAdd(sums + C + cbs[k] + lane, (double)s1[k]);
      }
  }
};

struct Scatter2 {   // layer-1 dgrad over the D feature columns: grad_feats[pt, cb..cb+31] += v
  float* __restrict__ grad_feats;
  const int32_t* __restrict__ nbr;
  int N, S, K, D, group_all;
  __device__ __forceinline__ void init(float*) {}
  __device__ __forceinline__ void block(float (&v)[32], int row, bool valid, int cb, int blk, int lane) {
    if (!valid || cb >= D) return;
    int pt;
    if (group_all) pt = row;
    else {
      int i = __ldg(nbr + row);
      i = min(max(i, 0), N - 1);
      pt = (row / K / S) * N + i;
    }
    float4* dst = reinterpret_cast<float4*>(grad_feats + (size_t)pt * D + cb);
#pragma unroll
    for (int q = 0; q < 8; ++q) atomicAdd(dst + q, make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]));
  }
  __device__ __forceinline__ void finish(int) {}
};

struct NoEpi2 {
  __device__ __forceinline__ void init(float*) {}
  __device__ __forceinline__ void block(float (&)[32], int, bool, int, int, int) {}
  __device__ __forceinline__ void finish(int) {}
};

// warp w reads TMEM lanes 32*(w%4).., columns [ (w/4)*ncols/2, (w/4+1)*ncols/2 ) in blocks of 32
template <class Epi>
__device__ __forceinline__ void drain_tile(Epi& epi, uint32_t tmem, int ncols, int m0, int M) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, q = warp & 3, half = warp >> 2;
  const int per = (ncols + 63) / 64 * 32;           // columns per half, multiple of 32
  const int row = m0 + q * 32 + lane;
  const bool valid = row < M;
  int blk = 0;
  for (int cb = half * per; cb < min(ncols, (half + 1) * per); cb += 32, ++blk) {
    float v[32];
    tc::tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)cb, v);
    epi.block(v, row, valid, cb, blk, lane);
  }
}

// ---------------------------------------------------------------------------------------------
// forward layer: y[M x N] = A'[M x kpad] * W^T, N = ncols (multiple of 32, <= 256)
// smem: [A: kpad/64 tiles of 16 KB][W: kpad/64 tiles of ncols_pad*128 B]
// ---------------------------------------------------------------------------------------------
template <class Prod, class Epi, int TCOLS>
__global__ void __launch_bounds__(kThreads)
tc2_fwd_kernel(Prod prod, const __nv_bfloat16* __restrict__ Wb, int ldw, int wrows, Epi epi, int M, int ncols,
               int kpad, int kmma /* K extent actually multiplied, multiple of 16 */) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);  // keeps the shared address space
  const int kt = kpad / 64;
  uint8_t* sA = smem;
  uint8_t* sW = smem + (size_t)kt * 128 * 128;
  __shared__ uint64_t mbar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nmma = (ncols + 15) / 16 * 16;        // UMMA N
  const int wr = min(wrows, (nmma + 7) / 8 * 8);  // weight rows staged (zero padded in global)

  if (warp == 0) tc::tmem_alloc<TCOLS>(&tmem_base);
  if (tid == 0) tc::mbar_init(&mbar, 1);
  load_weights(Wb, ldw, wr, kpad, sW);
  prod.bind((tid % (kpad / 8)) * 8);
  epi.init(reinterpret_cast<float*>(sW + (size_t)kt * wr * 128));
  tc::fence_proxy_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = tmem_base;
  const uint32_t idesc = tc::make_idesc_bf16(128, nmma, false, false);
  const uint32_t a0 = tc::smem_u32(sA), w0 = tc::smem_u32(sW);

  uint32_t phase = 0;
  const int ntiles = (M + 127) / 128;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int m0 = tile * 128;
    produce_tile(prod, kpad / 8, m0, sA);
    tc::fence_proxy_async();
    __syncthreads();
    if (tid == 0) {
      tc::fence_after_sync();
      for (int k = 0; k < kmma; k += 16) {
        const uint32_t off = (uint32_t)(k >> 6), sub = (uint32_t)((k & 63) * 2);
        tc::mma_bf16(tmem, tc::make_desc_sw128(a0 + off * (128 * 128) + sub, 16, 1024),
                     tc::make_desc_sw128(w0 + off * (uint32_t)(wr * 128) + sub, 16, 1024), idesc, k > 0);
      }
      tc::mma_commit(&mbar);
    }
    tc::mbar_wait(&mbar, phase);
    phase ^= 1;
    tc::fence_after_sync();
    drain_tile(epi, tmem, ncols, m0, M);
    tc::fence_before_sync();
    __syncthreads();
  }
  epi.finish(lane);
  if (warp == 0) tc::tmem_dealloc<TCOLS>(tmem);
}

// ---------------------------------------------------------------------------------------------
// fused backward of one layer.  P = dy tile [128 x ca] (ca = C_l), Q = x_prev tile [128 x cbk]
//   dW[ca x cbk] += P^T Q   (TMEM columns [0, (ca/128)*cbn) , accumulated over all tiles of the CTA)
//   dx[128 x cdn] = P W     (TMEM columns after that; W^T resident: rows = cdn outputs, K = ca)
// dW is written with fp32 atomics through `wmap` (column permutation of layer 1) at the end.
// smem: [P: ca/64 tiles][Q: cbk/64 tiles][W^T: ca/64 tiles of cdn_pad*128 B]
// ---------------------------------------------------------------------------------------------
template <class PProd, class QProd, class Epi, int TCOLS>
__global__ void __launch_bounds__(kThreads)
tc2_bwd_kernel(PProd pp, QProd qp, const __nv_bfloat16* __restrict__ WbT, int ldw, int wrows, Epi epi,
               float* __restrict__ dW, int ldo, int cb_valid, int perm_d /* >=0: layer-1 column permutation */,
               int M, int ca, int cbk, int cbmma, int cdn /* dgrad output columns, 0 = no dgrad */) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);  // keeps the shared address space
  const int pt = ca / 64, qt = cbk / 64;
  uint8_t* sP = smem;
  uint8_t* sQ = sP + (size_t)pt * 128 * 128;
  uint8_t* sW = sQ + (size_t)qt * 128 * 128;
  __shared__ uint64_t mbar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int mt = (ca + 127) / 128;                 // M tiles of dW
  const int dnm = (cdn + 15) / 16 * 16;            // UMMA N of the dgrad
  const int wr = cdn ? min(wrows, (dnm + 7) / 8 * 8) : 0;
  const uint32_t dx_col = (uint32_t)((mt * cbmma + 63) / 64 * 64);

  if (warp == 0) tc::tmem_alloc<TCOLS>(&tmem_base);
  if (tid == 0) tc::mbar_init(&mbar, 1);
  if (cdn) load_weights(WbT, ldw, wr, ca, sW);
  // one thread may own a P unit and a Q unit
  pp.bind((tid % (ca / 8)) * 8);
  qp.bind((tid % (cbk / 8)) * 8);
  epi.init(reinterpret_cast<float*>(sW + (size_t)pt * wr * 128));
  tc::fence_proxy_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = tmem_base;
  const uint32_t idesc_w = tc::make_idesc_bf16(128, cbmma, true, true);
  const uint32_t idesc_d = tc::make_idesc_bf16(128, dnm ? dnm : 16, false, false);
  const uint32_t p0 = tc::smem_u32(sP), q0 = tc::smem_u32(sQ), w0 = tc::smem_u32(sW);

  uint32_t phase = 0;
  bool acc = false;
  const int ntiles = (M + 127) / 128;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int m0 = tile * 128;
    produce_tile(pp, ca / 8, m0, sP);
    produce_tile(qp, cbk / 8, m0, sQ);
    tc::fence_proxy_async();
    __syncthreads();
    if (tid == 0) {
      tc::fence_after_sync();
      // dW += P^T Q : contraction over the 128 rows, 16 per MMA; M tile mi = channels [128*mi, 128*mi+128)
      for (int mi = 0; mi < mt; ++mi)
        for (int ks = 0; ks < 8; ++ks)
          tc::mma_bf16(tmem + (uint32_t)(mi * cbmma),
                       tc::make_desc_sw128(p0 + (uint32_t)(mi * 2) * (128 * 128) + ks * 2048, 128 * 128, 1024),
                       tc::make_desc_sw128(q0 + ks * 2048, 128 * 128, 1024), idesc_w, acc || ks > 0);
      // dx = P W : contraction over the ca channels
      if (cdn)
        for (int k = 0; k < ca; k += 16) {
          const uint32_t off = (uint32_t)(k >> 6), sub = (uint32_t)((k & 63) * 2);
          tc::mma_bf16(tmem + dx_col, tc::make_desc_sw128(p0 + off * (128 * 128) + sub, 16, 1024),
                       tc::make_desc_sw128(w0 + off * (uint32_t)(wr * 128) + sub, 16, 1024), idesc_d, k > 0);
        }
      tc::mma_commit(&mbar);
    }
    acc = true;
    tc::mbar_wait(&mbar, phase);
    phase ^= 1;
    tc::fence_after_sync();
    if (cdn) drain_tile(epi, tmem + dx_col, cdn, m0, M);
    tc::fence_before_sync();
    __syncthreads();
  }
  epi.finish(lane);
  // flush dW: TMEM lane = channel ca index within its M tile, columns = x_prev channel
  if (acc) {
    tc::fence_after_sync();
    const int q = warp & 3, half = warp >> 2;
    const int per = (cbmma + 63) / 64 * 32;
    for (int mi = 0; mi < mt; ++mi) {
      const int crow = mi * 128 + q * 32 + lane;
      for (int cb = half * per; cb < min(cbmma, (half + 1) * per); cb += 32) {
        float v[32];
        tc::tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(mi * cbmma + cb), v);
        if (crow < ca) {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            int c = cb + i;
            if (c >= cb_valid) continue;
            if (perm_d >= 0) c = c < perm_d ? c + 3 : c - perm_d;   // [feats | xyz] -> [xyz | feats]
            atomicAdd(dW + (size_t)crow * ldo + c, v[i]);
          }
        }
      }
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc<TCOLS>(tmem);
}

}  // namespace v2
}  // namespace pcoe
