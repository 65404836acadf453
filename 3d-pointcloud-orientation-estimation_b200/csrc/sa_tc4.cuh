// sa_tc4.cuh — tcgen05 kernels of the set-abstraction MLP with CHANNELS ON THE MMA-M AXIS (bf16 mode).
//
// Every GEMM of the layer is oriented so that a TMEM lane (= one epilogue thread) is one channel
// and the TMEM columns are points:
//     forward   Y^T[C_out x P]  = W[C_out x C_in]   * X^T[C_in x P]
//     dgrad     dX^T[C_in x P]  = W^T[C_in x C_out] * dY^T[C_out x P]
//     wgrad     dW[C_out x C_in] += dY^T[C_out x P] * X[P x C_in]          (contraction over points)
// so BatchNorm batch statistics, the BatchNorm-backward sums and the max / arg-max over a group of
// K = 32 neighbours (= one 32-column tcgen05.ld) are plain per-thread register reductions - no warp
// shuffles - and BatchNorm's scale/shift are per-thread constants.  Activations live in HBM
// channel-major: y^T [C][Mld] bf16 (Mld = rows rounded up to 128), so a thread's 32 points are 64
// contiguous bytes and a tile row is 256 contiguous bytes.
//
// Shared-memory operand images (all SWIZZLE_128B, 8-row / 1024-byte atoms, see tc_common.cuh):
//   activation tile, channel-major  [C rows][128 points] = 2 blocks (64 points = 128 B per row) of
//       C rows each.  Used K-major (contraction over points: wgrad A and B) and MN-major
//       (points on the N axis: forward / dgrad B operand).
//   activation tile, point-major    [128 points][C/64 blocks of 64 channels] (layer-1 gather of
//       fp32 row-major features).  Used K-major (forward B) and MN-major (wgrad B).
//   weight image [Rp = pad128(C_out) rows][Kp/64 blocks of 64 input channels], zero padded.
//       Used K-major (forward A) and MN-major (dgrad A: M = input channels, K = C_out rows;
//       layer-1 dgrad B: N = input channels).  One bf16 copy per layer serves every GEMM.
//
// CTA = 13 warps, persistent over 128-point tiles, one CTA per SM:
//   warps 0-3   epilogue   (warp w owns TMEM lanes 32w..32w+31)
//   warps 4-11  producers  (global -> BN/ReLU or BN-backward transform -> bf16 -> swizzled smem)
//   warp  12    MMA issue  (one elected thread)
// Pipelines: smem stage ring (full/empty mbarriers), double-buffered TMEM accumulators
// (tmem_full/tmem_empty); dW accumulates in TMEM over the CTA's whole tile range.
#pragma once
#include "sa_common.cuh"
#include "tc_common.cuh"

namespace pcoe {
namespace v4 {

constexpr int kPts = 128;
constexpr int kEpiThreads = 128, kProdThreads = 256, kThreads = kEpiThreads + kProdThreads + 32;
constexpr int kMaxStages = 3;

__device__ __forceinline__ void mbar_arrive(uint64_t* mbar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc::smem_u32(mbar)) : "memory");
}

// byte offset of the 16-byte chunk (8 points) `chunk` (0..15) of channel row c in a channel-major tile
__device__ __forceinline__ uint32_t cm_off(int crows, int c, int chunk) {
  return (uint32_t)((chunk >> 3) * crows * 128 + (c >> 3) * 1024 + (c & 7) * 128 + ((((chunk & 7) ^ (c & 7))) << 4));
}

__device__ __forceinline__ void unpack8(const uint4& a, float (&v)[8]) {
  const uint32_t w[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    v[2 * u] = __uint_as_float(w[u] << 16);
    v[2 * u + 1] = __uint_as_float(w[u] & 0xFFFF0000u);
  }
}

__device__ __forceinline__ void red_add_v4(float* dst, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// ---------------------------------------------------------------------------------------------
// Producers.  kChMajor: tile image (see header).  rows(): rows of the image (channel rows for a
// channel-major tile).  tile_bytes(): smem bytes of one stage.  nconst(): floats of per-channel
// constants cached in shared memory.  produce(ptid, m0, saddr): write the tile of points
// [m0, m0+128); every element of the K extent that the MMAs read must be written (or stay zero
// from the one-time clear of the stage buffers).
// ---------------------------------------------------------------------------------------------

// relu(scale * y + shift) of the previous layer's pre-activations y^T [C][Mld]
struct BnRelu4 {
  static constexpr bool kChMajor = true;
  const __nv_bfloat16* __restrict__ y;
  const float* __restrict__ scale;
  const float* __restrict__ shift;
  int M, Mld, C;
  const float* cs;
  __host__ __device__ __forceinline__ int rows() const { return C; }
  __host__ __device__ __forceinline__ int kext() const { return C; }
  __host__ __device__ __forceinline__ int nconst() const { return 2 * C; }
  __device__ __forceinline__ void init(float* csm, int tid, int nthr) {
    for (int c = tid; c < C; c += nthr) { csm[c] = scale[c]; csm[C + c] = shift[c]; }
    cs = csm;
  }
  __device__ __forceinline__ void produce(int ptid, int m0, uint32_t saddr) const {
    const int chunk = ptid & 15, r0 = ptid >> 4, crows = rows();
    const bool ok = m0 + chunk * 8 < M;
    const __nv_bfloat16* src = y + (size_t)m0 + chunk * 8;
#pragma unroll 4
    for (int c = r0; c < C; c += 16) {
      uint4 raw = make_uint4(0, 0, 0, 0);
      if (ok) raw = __ldg(reinterpret_cast<const uint4*>(src + (size_t)c * Mld));
      float v[8];
      unpack8(raw, v);
      const float sc = cs[c], sh = cs[C + c];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = ok ? fmaxf(fmaf(v[u], sc, sh), 0.f) : 0.f;
      tc::sts128(saddr + cm_off(crows, c, chunk), tc::pack8_bf16(v));
    }
  }
};

// dy^T = a*dz^T + p*y^T + q  (BatchNorm backward folded into per-channel constants), dense dz
struct Dy4 {
  static constexpr bool kChMajor = true;
  const __nv_bfloat16* __restrict__ dz;
  const __nv_bfloat16* __restrict__ y;
  const float* __restrict__ a;
  const float* __restrict__ p;
  const float* __restrict__ q;
  int M, Mld, C;
  const float* cs;
  __host__ __device__ __forceinline__ int rows() const { return C < 128 ? 128 : C; }
  __host__ __device__ __forceinline__ int kext() const { return C; }
  __host__ __device__ __forceinline__ int nconst() const { return 3 * C; }
  __device__ __forceinline__ void init(float* csm, int tid, int nthr) {
    for (int c = tid; c < C; c += nthr) { csm[c] = a[c]; csm[C + c] = p[c]; csm[2 * C + c] = q[c]; }
    cs = csm;
  }
  __device__ __forceinline__ void produce(int ptid, int m0, uint32_t saddr) const {
    const int chunk = ptid & 15, r0 = ptid >> 4, crows = rows();
    const bool ok = m0 + chunk * 8 < M;
    const size_t col = (size_t)m0 + chunk * 8;
#pragma unroll 4
    for (int c = r0; c < C; c += 16) {
      uint4 rd = make_uint4(0, 0, 0, 0), ry = rd;
      if (ok) {
        rd = __ldg(reinterpret_cast<const uint4*>(dz + (size_t)c * Mld + col));
        ry = __ldg(reinterpret_cast<const uint4*>(y + (size_t)c * Mld + col));
      }
      float d[8], yy[8], v[8];
      unpack8(rd, d);
      unpack8(ry, yy);
      const float ca = cs[c], cp = cs[C + c], cq = cs[2 * C + c];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = ok ? fmaf(ca, d[u], fmaf(cp, yy[u], cq)) : 0.f;
      tc::sts128(saddr + cm_off(crows, c, chunk), tc::pack8_bf16(v));
    }
  }
};

// last layer: the upstream gradient is the max-pool routing of gm[G,C] to the saved arg slot (K == 32)
struct DyLast4 {
  static constexpr bool kChMajor = true;
  const float* __restrict__ gm;       // [G,C]
  const uint8_t* __restrict__ slot;   // [G,C]
  const __nv_bfloat16* __restrict__ y;
  const float* __restrict__ a;
  const float* __restrict__ p;
  const float* __restrict__ q;
  int M, Mld, C;
  const float* cs;
  __host__ __device__ __forceinline__ int rows() const { return C < 128 ? 128 : C; }
  __host__ __device__ __forceinline__ int kext() const { return C; }
  __host__ __device__ __forceinline__ int nconst() const { return 3 * C; }
  __device__ __forceinline__ void init(float* csm, int tid, int nthr) {
    for (int c = tid; c < C; c += nthr) { csm[c] = a[c]; csm[C + c] = p[c]; csm[2 * C + c] = q[c]; }
    cs = csm;
  }
  __device__ __forceinline__ void produce(int ptid, int m0, uint32_t saddr) const {
    const int chunk = ptid & 15, r0 = ptid >> 4, crows = rows();
    const int m = m0 + chunk * 8;
    const bool ok = m < M;
    const int g = min(m, M - 1) >> 5, j0 = m & 31;
#pragma unroll 4
    for (int c = r0; c < C; c += 16) {
      uint4 ry = make_uint4(0, 0, 0, 0);
      float gv = 0.f;
      int sl = -1;
      if (ok) {
        ry = __ldg(reinterpret_cast<const uint4*>(y + (size_t)c * Mld + m));
        gv = __ldg(gm + (size_t)g * C + c);
        sl = (int)__ldg(slot + (size_t)g * C + c) - j0;
      }
      float yy[8], v[8];
      unpack8(ry, yy);
      const float ca = cs[c] * gv, cp = cs[C + c], cq = cs[2 * C + c];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = ok ? fmaf(cp, yy[u], cq) + (u == sl ? ca : 0.f) : 0.f;
      tc::sts128(saddr + cm_off(crows, c, chunk), tc::pack8_bf16(v));
    }
  }
};

// layer-1 input without features (SA1): [xyz[nbr] - centroid] as a channel-major tile of 16 rows
// (rows 3..15 stay zero from the one-time clear)
struct GatherXyz4 {
  static constexpr bool kChMajor = true;
  const float* __restrict__ xyz;
  const float* __restrict__ new_xyz;
  const int32_t* __restrict__ nbr;
  int N, S, group_all, M;
  __host__ __device__ __forceinline__ int rows() const { return 16; }
  __host__ __device__ __forceinline__ int kext() const { return 16; }
  __host__ __device__ __forceinline__ int nconst() const { return 0; }
  __device__ __forceinline__ void init(float*, int, int) {}
  __device__ __forceinline__ void produce(int ptid, int m0, uint32_t saddr) const {
    if (ptid >= kPts) return;
    const int row = m0 + ptid;
    float v[3] = {0.f, 0.f, 0.f};
    if (row < M) {
      const int g = row >> 5;
      int pt = row;
      if (!group_all) {
        int i = __ldg(nbr + row);
        i = min(max(i, 0), N - 1);
        pt = (g / S) * N + i;
      }
#pragma unroll
      for (int u = 0; u < 3; ++u) {
        float x = __ldg(xyz + (size_t)pt * 3 + u);
        if (!group_all) x = __fsub_rn(x, __ldg(new_xyz + (size_t)g * 3 + u));
        v[u] = x;
      }
    }
    const uint32_t base = saddr + (uint32_t)((ptid >> 6) * 16 * 128) + (uint32_t)((ptid & 7) * 2);
    const int ch = (ptid & 63) >> 3;
#pragma unroll
    for (int u = 0; u < 3; ++u) {
      const uint16_t h = __bfloat16_as_ushort(__float2bfloat16_rn(v[u]));
      asm volatile("st.shared.b16 [%0], %1;" ::"r"(base + (uint32_t)(u * 128 + ((ch ^ u) << 4))), "h"(h) : "memory");
    }
  }
};

// layer-1 input with features (SA2, SA3): point-major tile [128 points][kq channels],
// channel order [feats(D) | xyz - centroid (3) | zeros], D % 8 == 0
struct GatherFeat4 {
  static constexpr bool kChMajor = false;
  const float* __restrict__ xyz;
  const float* __restrict__ new_xyz;
  const int32_t* __restrict__ nbr;
  const float* __restrict__ feats;
  int N, S, D, group_all, M;
  __host__ __device__ __forceinline__ int kext() const { return (D + 3 + 15) / 16 * 16; }
  __host__ __device__ __forceinline__ int rows() const { return kPts; }
  __host__ __device__ __forceinline__ int nblocks() const { return (kext() + 63) / 64; }
  __host__ __device__ __forceinline__ int nconst() const { return 0; }
  __device__ __forceinline__ void init(float*, int, int) {}
  __device__ __forceinline__ void produce(int ptid, int m0, uint32_t saddr) const {
    const int upr = kext() / 8;   // 16-byte units per point row
    for (int e = ptid; e < kPts * upr; e += kProdThreads) {
      const int r = e / upr, j = e - r * upr, c0 = j * 8;
      const int row = m0 + r;
      float v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = 0.f;
      if (row < M && c0 < D + 3) {
        const int g = row >> 5;
        int pt = row;
        if (!group_all) {
          int i = __ldg(nbr + row);
          i = min(max(i, 0), N - 1);
          pt = (g / S) * N + i;
        }
        if (c0 < D) {
          const float4 a = __ldg(reinterpret_cast<const float4*>(feats + (size_t)pt * D + c0));
          const float4 b = __ldg(reinterpret_cast<const float4*>(feats + (size_t)pt * D + c0) + 1);
          v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
        } else {
#pragma unroll
          for (int u = 0; u < 3; ++u) {
            float x = __ldg(xyz + (size_t)pt * 3 + u);
            if (!group_all) x = __fsub_rn(x, __ldg(new_xyz + (size_t)g * 3 + u));
            v[u] = x;
          }
        }
      }
      tc::sts128(saddr + (uint32_t)(j >> 3) * (kPts * 128) + tc::sw128_off(r, (j & 7) * 8), tc::pack8_bf16(v));
    }
  }
};

// ---------------------------------------------------------------------------------------------
// Epilogues.  Thread = channel (TMEM lane), v = 32 consecutive points.
//   block(v, c, m, valid, mi): channel c (may be >= C: padding lane), points m..m+31, mi = M tile
//   finish(): flush per-thread running sums
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void store32_bf16(__nv_bfloat16* dst, const float (&v)[32]) {
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float t[8] = {v[8 * q], v[8 * q + 1], v[8 * q + 2], v[8 * q + 3], v[8 * q + 4], v[8 * q + 5], v[8 * q + 6], v[8 * q + 7]};
    reinterpret_cast<uint4*>(dst)[q] = tc::pack8_bf16(t);
  }
}

struct StoreStats4 {
  __nv_bfloat16* __restrict__ y;   // [C][Mld]
  double* __restrict__ sums;       // [2,C] or nullptr (eval)
  int C, Mld;
  float s0[2], s1[2];
  __host__ __device__ __forceinline__ int nconst() const { return 0; }
  __device__ __forceinline__ void init(float*, int, int) { s0[0] = s0[1] = s1[0] = s1[1] = 0.f; }
  __device__ __forceinline__ void block(float (&v)[32], int c, int m, bool valid, int mi) {
    if (c >= C || !valid) return;
    store32_bf16(y + (size_t)c * Mld + m, v);
    float a = 0.f, b = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) { a += v[i]; b = fmaf(v[i], v[i], b); }
    if (mi == 0) { s0[0] += a; s1[0] += b; } else { s0[1] += a; s1[1] += b; }
  }
  __device__ __forceinline__ void finish(int lane128) {
    if (!sums) return;
#pragma unroll
    for (int mi = 0; mi < 2; ++mi) {
      const int c = mi * 128 + lane128;
      if (c < C) { atomicAdd(sums + c, (double)s0[mi]); atomicAdd(sums + C + c, (double)s1[mi]); }
    }
  }
};

struct Group4 {   // last layer, K == 32: the 32 columns of a block are one group
  __nv_bfloat16* __restrict__ y;   // [C][Mld] or nullptr (eval)
  double* __restrict__ sums;       // or nullptr
  float* __restrict__ ymax;        // [G,C]
  float* __restrict__ ymin;
  uint8_t* __restrict__ amax;
  uint8_t* __restrict__ amin;
  int C, Mld;
  float s0[2], s1[2];
  __host__ __device__ __forceinline__ int nconst() const { return 0; }
  __device__ __forceinline__ void init(float*, int, int) { s0[0] = s0[1] = s1[0] = s1[1] = 0.f; }
  __device__ __forceinline__ void block(float (&v)[32], int c, int m, bool valid, int mi) {
    if (c >= C || !valid) return;
    if (y) store32_bf16(y + (size_t)c * Mld + m, v);
    float a = 0.f, b = 0.f, mx = v[0], mn = v[0];
    int ax = 0, an = 0;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      a += v[i];
      b = fmaf(v[i], v[i], b);
      if (v[i] > mx) { mx = v[i]; ax = i; }     // strict: first arg-max / arg-min
      if (v[i] < mn) { mn = v[i]; an = i; }
    }
    if (mi == 0) { s0[0] += a; s1[0] += b; } else { s0[1] += a; s1[1] += b; }
    const size_t o = (size_t)(m >> 5) * C + c;
    ymax[o] = mx; ymin[o] = mn; amax[o] = (uint8_t)ax; amin[o] = (uint8_t)an;
  }
  __device__ __forceinline__ void finish(int lane128) {
    if (!sums) return;
#pragma unroll
    for (int mi = 0; mi < 2; ++mi) {
      const int c = mi * 128 + lane128;
      if (c < C) { atomicAdd(sums + c, (double)s0[mi]); atomicAdd(sums + C + c, (double)s1[mi]); }
    }
  }
};

// dz_prev^T = dx^T * [z_prev > 0]; sums of dz_prev and dz_prev * xhat_prev per channel
struct MaskStats4 {
  const __nv_bfloat16* __restrict__ yprev;   // [C][Mld]
  const float* __restrict__ scale;
  const float* __restrict__ shift;
  const float* __restrict__ mean;
  const float* __restrict__ invstd;
  __nv_bfloat16* __restrict__ dz;            // [C][Mld]
  double* __restrict__ sums;
  int C, Mld;
  float s0[2], s1[2];
  float sc[2], sh[2], mu[2], is[2];
  __host__ __device__ __forceinline__ int nconst() const { return 0; }
  __device__ __forceinline__ void init(float*, int tid, int) {
    s0[0] = s0[1] = s1[0] = s1[1] = 0.f;
#pragma unroll
    for (int mi = 0; mi < 2; ++mi) {
      const int c = mi * 128 + (tid & 127);
      const bool ok = c < C;
      sc[mi] = ok ? scale[c] : 0.f; sh[mi] = ok ? shift[c] : 0.f; mu[mi] = ok ? mean[c] : 0.f; is[mi] = ok ? invstd[c] : 0.f;
    }
  }
  __device__ __forceinline__ void block(float (&v)[32], int c, int m, bool valid, int mi) {
    if (c >= C || !valid) return;
    const float ksc = mi ? sc[1] : sc[0], ksh = mi ? sh[1] : sh[0], kmu = mi ? mu[1] : mu[0], kis = mi ? is[1] : is[0];
    const __nv_bfloat16* src = yprev + (size_t)c * Mld + m;
    float a = 0.f, b = 0.f;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float yy[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(src) + q), yy);
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const bool on = fmaf(yy[u], ksc, ksh) > 0.f;
        const float d = on ? v[8 * q + u] : 0.f;
        v[8 * q + u] = d;
        a += d;
        b = fmaf(d, (yy[u] - kmu) * kis, b);
      }
    }
    store32_bf16(dz + (size_t)c * Mld + m, v);
    if (mi == 0) { s0[0] += a; s1[0] += b; } else { s0[1] += a; s1[1] += b; }
  }
  __device__ __forceinline__ void finish(int lane128) {
#pragma unroll
    for (int mi = 0; mi < 2; ++mi) {
      const int c = mi * 128 + lane128;
      if (c < C) { atomicAdd(sums + c, (double)s0[mi]); atomicAdd(sums + C + c, (double)s1[mi]); }
    }
  }
};

// layer-1 dgrad over the D feature columns, POINT-on-lane orientation: thread = point, v = 32 feature channels
struct Scatter4 {
  float* __restrict__ grad_feats;
  const int32_t* __restrict__ nbr;
  int N, S, D, group_all;
  __host__ __device__ __forceinline__ int nconst() const { return 0; }
  __device__ __forceinline__ void init(float*, int, int) {}
  // here c = first feature channel of the block, m = the thread's point row
  __device__ __forceinline__ void block_pt(float (&v)[32], int cb, int row, bool valid) {
    if (!valid || cb >= D) return;
    int pt = row;
    if (!group_all) {
      int i = __ldg(nbr + row);
      i = min(max(i, 0), N - 1);
      pt = ((row >> 5) / S) * N + i;
    }
    float* dst = grad_feats + (size_t)pt * D + cb;
#pragma unroll
    for (int q = 0; q < 8; ++q) red_add_v4(dst + 4 * q, v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
  }
  __device__ __forceinline__ void finish(int) {}
};

struct NoEpi4 {
  __host__ __device__ __forceinline__ int nconst() const { return 0; }
  __device__ __forceinline__ void init(float*, int, int) {}
  __device__ __forceinline__ void finish(int) {}
};

struct Barriers4 {
  uint64_t full[kMaxStages];
  uint64_t empty[kMaxStages];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
};

// weight image: bf16 [Rp][Kp] row-major (zero padded) -> [Rp rows][Kp/64 blocks], SWIZZLE_128B
__device__ __forceinline__ void load_wimage(const __nv_bfloat16* __restrict__ W, int Rp, int Kp, uint32_t saddr, int tid, int nthr) {
  const int upr = Kp / 8;
  for (int e = tid; e < Rp * upr; e += nthr) {
    const int n = e / upr, j = e - n * upr;
    const uint4 w = __ldg(reinterpret_cast<const uint4*>(W + (size_t)n * Kp + j * 8));
    tc::sts128(saddr + (uint32_t)(j >> 3) * (uint32_t)(Rp * 128) + tc::sw128_off(n, (j & 7) * 8), w);
  }
}

__device__ __forceinline__ void zero_smem(uint32_t saddr, uint32_t bytes, int tid, int nthr) {
  const uint4 z = make_uint4(0, 0, 0, 0);
  for (uint32_t o = (uint32_t)tid * 16; o < bytes; o += (uint32_t)nthr * 16) tc::sts128(saddr + o, z);
}

template <class Prod>
__device__ __forceinline__ uint32_t prod_tile_bytes(const Prod& p) {
  if (Prod::kChMajor) return (uint32_t)(2 * p.rows() * 128);
  else return (uint32_t)(((p.kext() + 63) / 64) * kPts * 128);
}

// B-operand descriptor of activation tile `saddr` for the k-th 16-channel step (forward / dgrad:
// contraction over channels, N = 128 points)
template <class Prod>
__device__ __forceinline__ uint64_t act_desc_chan_k(const Prod& p, uint32_t saddr, int k) {
  if (Prod::kChMajor) return tc::make_desc_sw128(saddr + (uint32_t)(k >> 4) * 2048, (uint32_t)p.rows() * 128, 1024);   // MN-major
  else return tc::make_desc_sw128(saddr + (uint32_t)(k >> 6) * (kPts * 128) + (uint32_t)((k & 63) * 2), 16, 1024);      // K-major
}

// ---------------------------------------------------------------------------------------------
// forward layer:  Y^T[C_out x 128] = W * X^T per tile;  TMEM: 2 buffers x mt x 128 columns
// smem: [W image][stages x tile][constants]
// ---------------------------------------------------------------------------------------------
template <class Prod, class Epi, int TCOLS>
__global__ void __launch_bounds__(kThreads, 1)
tc4_fwd_kernel(Prod prod, const __nv_bfloat16* __restrict__ Wb, int Rp, int Kp, Epi epi, int M, int nstages) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem0 = (tc::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem0 - tc::smem_u32(smem_raw));
  __shared__ Barriers4 bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int mt = Rp / 128;
  const uint32_t wbytes = (uint32_t)Rp * (uint32_t)Kp * 2u;
  const uint32_t tbytes = (prod_tile_bytes(prod) + 1023u) & ~1023u;
  const uint32_t sW = smem0, sT = smem0 + wbytes;
  float* csm = reinterpret_cast<float*>(smem_gen + wbytes + (size_t)nstages * tbytes);

  if (warp == 0) tc::tmem_alloc<TCOLS>(&tmem_base);
  if (tid == 0) {
    for (int s = 0; s < nstages; ++s) { tc::mbar_init(&bar.full[s], kProdThreads); tc::mbar_init(&bar.empty[s], 1); }
    for (int b = 0; b < 2; ++b) { tc::mbar_init(&bar.tmem_full[b], 1); tc::mbar_init(&bar.tmem_empty[b], kEpiThreads); }
  }
  load_wimage(Wb, Rp, Kp, sW, tid, kThreads);
  zero_smem(sT, (uint32_t)nstages * tbytes, tid, kThreads);
  prod.init(csm, tid, kThreads);
  epi.init(csm + prod.nconst(), tid, kThreads);
  tc::fence_proxy_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = tmem_base;
  const int ntiles = (M + kPts - 1) / kPts;

  if (warp < 4) {
    // ---- epilogue: thread = TMEM lane = channel (mi*128 + tid) ----
    int i = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++i) {
      const int b = i & 1, u = i >> 1, m0 = tile * kPts;
      tc::mbar_wait(&bar.tmem_full[b], (uint32_t)(u & 1));
      tc::fence_after_sync();
      for (int mi = 0; mi < mt; ++mi)
#pragma unroll 1
        for (int j = 0; j < 4; ++j) {
          float v[32];
          tc::tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(b * mt * kPts + mi * kPts + j * 32), v);
          epi.block(v, mi * 128 + tid, m0 + j * 32, m0 + j * 32 < M, mi);
        }
      tc::fence_before_sync();
      mbar_arrive(&bar.tmem_empty[b]);
    }
    epi.finish(tid);
  } else if (warp < 12) {
    // ---- producers ----
    const int ptid = tid - kEpiThreads;
    int i = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++i) {
      const int s = i % nstages, n = i / nstages;
      if (n > 0) tc::mbar_wait(&bar.empty[s], (uint32_t)((n - 1) & 1));
      prod.produce(ptid, tile * kPts, sT + (uint32_t)s * tbytes);
      tc::fence_proxy_async();
      mbar_arrive(&bar.full[s]);
    }
  } else if (tid == kEpiThreads + kProdThreads) {
    // ---- MMA issue ----
    const uint32_t idesc = tc::make_idesc_bf16(128, kPts, false, Prod::kChMajor);
    const int kext = prod.kext();
    int i = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++i) {
      const int s = i % nstages, n = i / nstages, b = i & 1, u = i >> 1;
      tc::mbar_wait(&bar.full[s], (uint32_t)(n & 1));
      if (u > 0) tc::mbar_wait(&bar.tmem_empty[b], (uint32_t)((u - 1) & 1));
      tc::fence_after_sync();
      const uint32_t st = sT + (uint32_t)s * tbytes;
      for (int mi = 0; mi < mt; ++mi)
        for (int k = 0; k < kext; k += 16)
          tc::mma_bf16(tmem + (uint32_t)(b * mt * kPts + mi * kPts),
                       tc::make_desc_sw128(sW + (uint32_t)(k >> 6) * (uint32_t)(Rp * 128) + (uint32_t)(mi * 128 * 128) + (uint32_t)((k & 63) * 2), 16, 1024),
                       act_desc_chan_k(prod, st, k), idesc, k > 0);
      tc::mma_commit(&bar.empty[s]);
      tc::mma_commit(&bar.tmem_full[b]);
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc<TCOLS>(tmem);
}

// ---------------------------------------------------------------------------------------------
// fused backward of one layer.  P = dy^T tile [C_l x 128] (channel-major), Q = x_prev tile.
//   dW[C_l x Cq] += P * Q^T            TMEM columns [0, mtl*nw), accumulated over the CTA's tiles
//   DGRAD == 1:  dx^T[C_prev x 128] = W^T * P    (thread = channel epilogue; 2 buffers of mtp*128 cols)
//   DGRAD == 2:  dx[128 x D]        = P^T * W     (thread = point epilogue: layer-1 scatter; 2 buffers of D cols)
// smem: [W image (DGRAD)][P tile][Q tile][constants]   (single stage)
// ---------------------------------------------------------------------------------------------
template <class PProd, class QProd, class Epi, int DGRAD, int TCOLS>
__global__ void __launch_bounds__(kThreads, 1)
tc4_bwd_kernel(PProd pp, QProd qp, const __nv_bfloat16* __restrict__ Wb, int Rp, int Kp, Epi epi,
               float* __restrict__ dW, int ldo, int cq_valid, int perm_d /* >=0: layer-1 [feats|xyz] column order */,
               int M, int cprev /* dgrad output channels */) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem0 = (tc::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem0 - tc::smem_u32(smem_raw));
  __shared__ Barriers4 bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int cl = pp.C, mtl = (cl + 127) / 128;
  const int nw = qp.kext();                       // dW columns per M tile (Cq rounded up to 16)
  const int mtp = (cprev + 127) / 128;
  const uint32_t wbytes = DGRAD ? (uint32_t)Rp * (uint32_t)Kp * 2u : 0u;
  const uint32_t pbytes = (prod_tile_bytes(pp) + 1023u) & ~1023u, qbytes = (prod_tile_bytes(qp) + 1023u) & ~1023u;
  const uint32_t sW = smem0, sP = smem0 + wbytes, sQ = sP + pbytes;
  float* csm = reinterpret_cast<float*>(smem_gen + wbytes + pbytes + qbytes);
  const uint32_t dx_col = (uint32_t)((mtl * nw + 31) / 32 * 32);
  const uint32_t dx_cols = DGRAD == 1 ? (uint32_t)(mtp * kPts) : (uint32_t)((cprev + 31) / 32 * 32);

  if (warp == 0) tc::tmem_alloc<TCOLS>(&tmem_base);
  if (tid == 0) {
    tc::mbar_init(&bar.full[0], kProdThreads); tc::mbar_init(&bar.empty[0], 1);
    for (int b = 0; b < 2; ++b) { tc::mbar_init(&bar.tmem_full[b], 1); tc::mbar_init(&bar.tmem_empty[b], kEpiThreads); }
  }
  if (DGRAD) load_wimage(Wb, Rp, Kp, sW, tid, kThreads);
  zero_smem(sP, pbytes + qbytes, tid, kThreads);
  pp.init(csm, tid, kThreads);
  qp.init(csm + pp.nconst(), tid, kThreads);
  epi.init(csm + pp.nconst() + qp.nconst(), tid, kThreads);
  tc::fence_proxy_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = tmem_base;
  const int ntiles = (M + kPts - 1) / kPts;
  int my_tiles = 0;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) ++my_tiles;

  if (warp < 4) {
    int i = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++i) {
      const int b = i & 1, u = i >> 1, m0 = tile * kPts;
      tc::mbar_wait(&bar.tmem_full[b], (uint32_t)(u & 1));
      tc::fence_after_sync();
      if constexpr (DGRAD == 1) {
        for (int mi = 0; mi < mtp; ++mi)
#pragma unroll 1
          for (int j = 0; j < 4; ++j) {
            float v[32];
            tc::tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + dx_col + (uint32_t)b * dx_cols + (uint32_t)(mi * kPts + j * 32), v);
            epi.block(v, mi * 128 + tid, m0 + j * 32, m0 + j * 32 < M, mi);
          }
      } else if constexpr (DGRAD == 2) {
#pragma unroll 1
        for (int cb = 0; cb < cprev; cb += 32) {
          float v[32];
          tc::tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + dx_col + (uint32_t)b * dx_cols + (uint32_t)cb, v);
          epi.block_pt(v, cb, m0 + tid, m0 + tid < M);
        }
      }
      tc::fence_before_sync();
      mbar_arrive(&bar.tmem_empty[b]);
    }
    epi.finish(tid);
    // flush dW (complete: the last tmem_full commit covers every earlier MMA): thread = row of dW
    if (my_tiles > 0) {
      tc::fence_after_sync();
      for (int mi = 0; mi < mtl; ++mi) {
        const int crow = mi * 128 + tid;
#pragma unroll 1
        for (int cb = 0; cb < nw; cb += 32) {
          float v[32];
          tc::tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(mi * nw + cb), v);   // may read past nw: unused columns
          if (crow < cl) {
            float* dst = dW + (size_t)crow * ldo;
            if (perm_d < 0 && (ldo & 3) == 0 && cb + 32 <= cq_valid) {
#pragma unroll
              for (int q = 0; q < 8; ++q) red_add_v4(dst + cb + 4 * q, v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
            } else {
#pragma unroll
              for (int k = 0; k < 32; ++k) {
                int c = cb + k;
                if (c >= cq_valid) continue;
                if (perm_d >= 0) c = c < perm_d ? c + 3 : c - perm_d;   // [feats | xyz] -> [xyz | feats]
                atomicAdd(dst + c, v[k]);
              }
            }
          }
        }
      }
    }
  } else if (warp < 12) {
    const int ptid = tid - kEpiThreads;
    int i = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++i) {
      if (i > 0) tc::mbar_wait(&bar.empty[0], (uint32_t)((i - 1) & 1));
      pp.produce(ptid, tile * kPts, sP);
      qp.produce(ptid, tile * kPts, sQ);
      tc::fence_proxy_async();
      mbar_arrive(&bar.full[0]);
    }
  } else if (tid == kEpiThreads + kProdThreads) {
    const uint32_t idesc_w = tc::make_idesc_bf16(128, nw, false, !QProd::kChMajor);
    const uint32_t idesc_d = DGRAD == 1 ? tc::make_idesc_bf16(128, kPts, true, true)
                                        : tc::make_idesc_bf16(128, (cprev + 15) / 16 * 16, true, true);
    const int prow = pp.rows();
    int i = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++i) {
      const int b = i & 1, u = i >> 1;
      tc::mbar_wait(&bar.full[0], (uint32_t)(i & 1));
      if (DGRAD && u > 0) tc::mbar_wait(&bar.tmem_empty[b], (uint32_t)((u - 1) & 1));
      tc::fence_after_sync();
      // dW += P Q^T: contraction over the 128 points, 16 per MMA
      for (int mi = 0; mi < mtl; ++mi)
        for (int ks = 0; ks < 8; ++ks) {
          const uint64_t ad = tc::make_desc_sw128(sP + (uint32_t)(ks >> 2) * (uint32_t)(prow * 128) + (uint32_t)(mi * 128 * 128) + (uint32_t)((ks & 3) * 32), 16, 1024);
          const uint64_t bd = QProd::kChMajor
              ? tc::make_desc_sw128(sQ + (uint32_t)(ks >> 2) * (uint32_t)(qp.rows() * 128) + (uint32_t)((ks & 3) * 32), 16, 1024)
              : tc::make_desc_sw128(sQ + (uint32_t)ks * 2048, kPts * 128, 1024);
          tc::mma_bf16(tmem + (uint32_t)(mi * nw), ad, bd, idesc_w, i > 0 || ks > 0);
        }
      if constexpr (DGRAD == 1) {
        for (int mj = 0; mj < mtp; ++mj)
          for (int k = 0; k < cl; k += 16)
            tc::mma_bf16(tmem + dx_col + (uint32_t)b * dx_cols + (uint32_t)(mj * kPts),
                         tc::make_desc_sw128(sW + (uint32_t)(2 * mj) * (uint32_t)(Rp * 128) + (uint32_t)(k >> 4) * 2048, (uint32_t)Rp * 128, 1024),
                         tc::make_desc_sw128(sP + (uint32_t)(k >> 4) * 2048, (uint32_t)prow * 128, 1024), idesc_d, k > 0);
      } else if constexpr (DGRAD == 2) {
        for (int k = 0; k < cl; k += 16)
          tc::mma_bf16(tmem + dx_col + (uint32_t)b * dx_cols,
                       tc::make_desc_sw128(sP + (uint32_t)(k >> 4) * 2048, (uint32_t)prow * 128, 1024),
                       tc::make_desc_sw128(sW + (uint32_t)(k >> 4) * 2048, (uint32_t)Rp * 128, 1024), idesc_d, k > 0);
      }
      tc::mma_commit(&bar.empty[0]);
      tc::mma_commit(&bar.tmem_full[b]);
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc<TCOLS>(tmem);
}

// fp32 [C_out][C_in] -> zero-padded bf16 [Rp][Kp]; perm_d >= 0: layer-1 column order [feats(perm_d) | xyz(3)]
struct ConvW4 { const float* W; __nv_bfloat16* dst; int cout, cin, Rp, Kp, perm_d; };
__global__ void convert_weights4_kernel(ConvW4 a, ConvW4 b, ConvW4 c) {
  const ConvW4* L[3] = {&a, &b, &c};
  const int n0 = a.Rp * a.Kp, n1 = b.Rp * b.Kp, n2 = c.Rp * c.Kp;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n0 + n1 + n2; e += gridDim.x * blockDim.x) {
    const int l = e < n0 ? 0 : (e < n0 + n1 ? 1 : 2);
    const ConvW4& w = *L[l];
    const int ee = e - (l == 0 ? 0 : (l == 1 ? n0 : n0 + n1));
    const int r = ee / w.Kp, k = ee - r * w.Kp;
    int src = k;
    if (w.perm_d >= 0) src = k < w.perm_d ? k + 3 : (k < w.perm_d + 3 ? k - w.perm_d : w.cin);
    float v = 0.f;
    if (r < w.cout && src < w.cin) v = w.W[(size_t)r * w.cin + src];
    w.dst[ee] = __float2bfloat16_rn(v);
  }
}

}  // namespace v4
}  // namespace pcoe
