// sa_common.cuh — building blocks of the set-abstraction MLP kernels.
//
// Rows.  The grouped tensor of the reference, (B,S,K,3+D) (models/pointnet_pp_8dir.py:31-37), is
// never built.  A "row" r = (b*S + s)*K + j is one neighbour of one centroid; row-major [M x C]
// matrices with M = B*S*K are the only activation layout.  A group (one centroid) is K consecutive
// rows, so the max over neighbours (:42) is a reduction over K consecutive rows of a 128-row tile.
//
// Producers build the A operand of a GEMM on the fly (gather+centre, BN+ReLU of the previous
// layer's pre-activations, or the BatchNorm-backward combination); epilogues consume the 128 x BN
// accumulator tile from shared memory (batch statistics, stores, max/min over the group, ReLU
// mask, scatter-add).  The same producers/epilogues feed the fp32 CUDA-core GEMM below and the
// tcgen05 GEMM (sa_tc.cu).
#pragma once
#include "common.cuh"
#include <cuda_bf16.h>
#include <math.h>

namespace pcoe {

template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// 8 consecutive elements of a row, vectorised when aligned.
template <typename T>
__device__ __forceinline__ void load_row8(const T* __restrict__ base, size_t off, float (&v)[8]);
template <>
__device__ __forceinline__ void load_row8<float>(const float* __restrict__ base, size_t off, float (&v)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(base + off));
  const float4 b = __ldg(reinterpret_cast<const float4*>(base + off) + 1);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
template <>
__device__ __forceinline__ void load_row8<__nv_bfloat16>(const __nv_bfloat16* __restrict__ base, size_t off, float (&v)[8]) {
  const uint4 a = __ldg(reinterpret_cast<const uint4*>(base + off));
  const uint32_t w[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    v[2 * u] = __uint_as_float(w[u] << 16);
    v[2 * u + 1] = __uint_as_float(w[u] & 0xFFFF0000u);
  }
}

// ---------------------------------------------------------------------------------------------
// A-operand producers.  load8(row, c0, v): channels c0..c0+7 of `row`, zeros outside [0,M)x[0,C).
// ---------------------------------------------------------------------------------------------

// Layer-1 input: [xyz[nbr] - centroid, feats[nbr]]   (pointnet_pp_8dir.py:23-37)
struct GatherProd {
  const float* __restrict__ xyz;      // [B,N,3]
  const float* __restrict__ new_xyz;  // [B,S,3]
  const int32_t* __restrict__ nbr;    // [B,S,K]
  const float* __restrict__ feats;    // [B,N,D] or nullptr
  int N, S, K, D, group_all, M, C;    // C = 3 + D

  __device__ __forceinline__ void load8(int row, int c0, float (&v)[8]) const {
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = 0.f;
    if (row >= M || c0 >= C) return;
    const int g = row / K;
    int pt;
    if (group_all) pt = row;  // K == N, S == 1: row = b*N + j
    else {
      int i = __ldg(nbr + row);
      i = min(max(i, 0), N - 1);
      pt = (g / S) * N + i;
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int c = c0 + u;
      if (c < 3) {
        float x = __ldg(xyz + (size_t)pt * 3 + c);
        if (!group_all) x = __fsub_rn(x, __ldg(new_xyz + (size_t)g * 3 + c));
        v[u] = x;
      } else if (c < C) {
        v[u] = __ldg(feats + (size_t)pt * D + (c - 3));
      }
    }
  }
};

// relu(y * scale + shift): the previous layer's BatchNorm + ReLU applied on load (:41)
template <typename TY>
struct BnReluProd {
  const TY* __restrict__ y;  // [M,C] pre-BN
  const float* __restrict__ scale;
  const float* __restrict__ shift;
  int M, C;

  __device__ __forceinline__ void load8(int row, int c0, float (&v)[8]) const {
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = 0.f;
    if (row >= M || c0 >= C) return;
    if ((C & 7) == 0) {
      load_row8<TY>(y, (size_t)row * C + c0, v);
      const float4 s0 = __ldg(reinterpret_cast<const float4*>(scale + c0)), s1 = __ldg(reinterpret_cast<const float4*>(scale + c0) + 1);
      const float4 h0 = __ldg(reinterpret_cast<const float4*>(shift + c0)), h1 = __ldg(reinterpret_cast<const float4*>(shift + c0) + 1);
      const float sc[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
      const float sh[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = fmaxf(fmaf(v[u], sc[u], sh[u]), 0.f);
    } else {
      for (int u = 0; u < 8 && c0 + u < C; ++u)
        v[u] = fmaxf(fmaf(to_f<TY>(y[(size_t)row * C + c0 + u]), __ldg(scale + c0 + u), __ldg(shift + c0 + u)), 0.f);
    }
  }
};

// BatchNorm backward folded into three per-channel constants:
//   dy = a*(dz - m1 - xhat*m2) = a*dz + p*y + q,  p = -a*invstd*m2,  q = -a*m1 - p*mean
// Dense upstream gradient dz[M,C] (already ReLU-masked).
template <typename TY>
struct DyProd {
  const TY* __restrict__ dz;
  const TY* __restrict__ y;
  const float* __restrict__ a;
  const float* __restrict__ p;
  const float* __restrict__ q;
  int M, C;

  __device__ __forceinline__ void load8(int row, int c0, float (&v)[8]) const {
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = 0.f;
    if (row >= M || c0 >= C) return;
    if ((C & 7) == 0) {
      float d[8], yy[8];
      load_row8<TY>(dz, (size_t)row * C + c0, d);
      load_row8<TY>(y, (size_t)row * C + c0, yy);
#pragma unroll
      for (int u = 0; u < 8; ++u)
        v[u] = fmaf(__ldg(a + c0 + u), d[u], fmaf(__ldg(p + c0 + u), yy[u], __ldg(q + c0 + u)));
    } else {
      for (int u = 0; u < 8 && c0 + u < C; ++u) {
        const int c = c0 + u;
        v[u] = fmaf(__ldg(a + c), to_f<TY>(dz[(size_t)row * C + c]),
                    fmaf(__ldg(p + c), to_f<TY>(y[(size_t)row * C + c]), __ldg(q + c)));
      }
    }
  }
};

// Same, for the last layer: the upstream gradient is the max-pool routing of gm[G,C]
// (gm = grad_out * [out > 0]) to the saved arg-max slot of every (group, channel).
template <typename TY>
struct DyLastProd {
  const float* __restrict__ gm;      // [G,C]
  const uint8_t* __restrict__ slot;  // [G,C]
  const TY* __restrict__ y;          // [M,C]
  const float* __restrict__ a;
  const float* __restrict__ p;
  const float* __restrict__ q;
  int M, C, K;

  __device__ __forceinline__ void load8(int row, int c0, float (&v)[8]) const {
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = 0.f;
    if (row >= M || c0 >= C) return;
    const int g = row / K, j = row - g * K;
    if ((C & 7) == 0) {
      float yy[8], gg[8];
      load_row8<TY>(y, (size_t)row * C + c0, yy);
      load_row8<float>(gm, (size_t)g * C + c0, gg);
      const uint2 sl = __ldg(reinterpret_cast<const uint2*>(slot + (size_t)g * C + c0));
      const uint32_t sw[2] = {sl.x, sl.y};
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int s = (sw[u >> 2] >> (8 * (u & 3))) & 0xFF;
        const float d = (s == j) ? gg[u] : 0.f;
        v[u] = fmaf(__ldg(a + c0 + u), d, fmaf(__ldg(p + c0 + u), yy[u], __ldg(q + c0 + u)));
      }
    } else {
      for (int u = 0; u < 8 && c0 + u < C; ++u) {
        const int c = c0 + u;
        const float d = (slot[(size_t)g * C + c] == j) ? gm[(size_t)g * C + c] : 0.f;
        v[u] = fmaf(__ldg(a + c), d, fmaf(__ldg(p + c), to_f<TY>(y[(size_t)row * C + c]), __ldg(q + c)));
      }
    }
  }
};

// ---------------------------------------------------------------------------------------------
// Epilogues.  run(Cs, ld, m0, n0, rows, cols): Cs is the accumulator tile in shared memory
// (row-major, leading dimension ld), rows/cols the valid extent; called by all 256 threads.
// ---------------------------------------------------------------------------------------------

__device__ __forceinline__ void tile_col_stats(const float* Cs, int ld, int rows, int cols, int bn,
                                               double* __restrict__ sums, int n0, int C) {
  // column sum and sum of squares over the tile rows; one fp64 atomic per column per tile
  const int tid = threadIdx.x, parts = 256 / bn, col = tid % bn, part = tid / bn;
  if (part >= parts || col >= cols) return;
  float s = 0.f, ss = 0.f;
  for (int r = part; r < rows; r += parts) {
    const float v = Cs[r * ld + col];
    s += v;
    ss = fmaf(v, v, ss);
  }
  atomicAdd(sums + n0 + col, (double)s);
  atomicAdd(sums + C + n0 + col, (double)ss);
}

// pre-BN store + batch statistics (layers 1 and 2)
template <typename TY>
struct StoreStatsEpi {
  TY* __restrict__ y;       // [M,C]
  double* __restrict__ sums;  // [2,C] or nullptr (eval)
  int C;
  template <int BN>
  __device__ __forceinline__ void run(const float* Cs, int ld, int m0, int n0, int rows, int cols) const {
    for (int e = threadIdx.x; e < rows * BN; e += 256) {
      const int r = e / BN, c = e % BN;
      if (c < cols) y[(size_t)(m0 + r) * C + n0 + c] = from_f<TY>(Cs[r * ld + c]);
    }
    if (sums) tile_col_stats(Cs, ld, rows, cols, BN, sums, n0, C);
  }
};

// last layer: statistics + max / min / first arg-max / first arg-min over each group of K rows.
// BatchNorm is a per-channel affine map a*y+b, so max_k relu(a*y_k+b) = relu(a*max_k y_k + b) for
// a >= 0 and relu(a*min_k y_k + b) for a < 0: the (M x C3) post-activation never exists.
template <typename TY>
struct GroupEpi {
  TY* __restrict__ y;          // [M,C] (train: needed by BN backward) or nullptr
  double* __restrict__ sums;   // or nullptr
  float* __restrict__ ymax;    // [G,C]
  float* __restrict__ ymin;
  uint8_t* __restrict__ amax;
  uint8_t* __restrict__ amin;
  int C, K;
  template <int BN>
  __device__ __forceinline__ void run(const float* Cs, int ld, int m0, int n0, int rows, int cols) const {
    if (y) {
      for (int e = threadIdx.x; e < rows * BN; e += 256) {
        const int r = e / BN, c = e % BN;
        if (c < cols) y[(size_t)(m0 + r) * C + n0 + c] = from_f<TY>(Cs[r * ld + c]);
      }
    }
    if (sums) tile_col_stats(Cs, ld, rows, cols, BN, sums, n0, C);
    const int groups = rows / K;
    for (int e = threadIdx.x; e < groups * BN; e += 256) {
      const int gl = e / BN, c = e % BN;
      if (c >= cols) continue;
      float mx = -INFINITY, mn = INFINITY;
      int ax = 0, an = 0;
      for (int j = 0; j < K; ++j) {
        const float v = Cs[(gl * K + j) * ld + c];
        if (v > mx) { mx = v; ax = j; }
        if (v < mn) { mn = v; an = j; }
      }
      const size_t o = (size_t)(m0 / K + gl) * C + n0 + c;
      ymax[o] = mx; ymin[o] = mn; amax[o] = (uint8_t)ax; amin[o] = (uint8_t)an;
    }
  }
};

// dgrad epilogue: dz_prev = dx * [z_prev > 0]; store; accumulate sum(dz) and sum(dz * xhat_prev)
template <typename TY>
struct MaskStatsEpi {
  const TY* __restrict__ yprev;  // [M,C]
  const float* __restrict__ scale;
  const float* __restrict__ shift;
  const float* __restrict__ mean;
  const float* __restrict__ invstd;
  TY* __restrict__ dz;           // [M,C]
  double* __restrict__ sums;     // [2,C]
  int C;
  template <int BN>
  __device__ __forceinline__ void run(float* Cs, int ld, int m0, int n0, int rows, int cols) const {
    const int tid = threadIdx.x, parts = 256 / BN, col = tid % BN, part = tid / BN;
    if (part < parts && col < cols) {
      const int c = n0 + col;
      const float sc = __ldg(scale + c), sh = __ldg(shift + c), mu = __ldg(mean + c), is = __ldg(invstd + c);
      float s0 = 0.f, s1 = 0.f;
      for (int r = part; r < rows; r += parts) {
        const float yv = to_f<TY>(yprev[(size_t)(m0 + r) * C + c]);
        const float g = (fmaf(yv, sc, sh) > 0.f) ? Cs[r * ld + col] : 0.f;
        dz[(size_t)(m0 + r) * C + c] = from_f<TY>(g);
        s0 += g;
        s1 = fmaf(g, (yv - mu) * is, s1);
      }
      atomicAdd(sums + c, (double)s0);
      atomicAdd(sums + C + c, (double)s1);
    }
  }
};

// layer-1 dgrad epilogue: scatter-add the feature part (columns >= 3) back through the neighbour
// index = autograd of index_points (index_put_ accumulate); xyz columns carry no gradient.
struct ScatterEpi {
  float* __restrict__ grad_feats;   // [B,N,D]
  const int32_t* __restrict__ nbr;  // [B,S,K]
  int N, S, K, D, group_all;
  template <int BN>
  __device__ __forceinline__ void run(const float* Cs, int ld, int m0, int n0, int rows, int cols) const {
    for (int e = threadIdx.x; e < rows * BN; e += 256) {
      const int r = e / BN, col = e % BN, c = n0 + col;
      if (col >= cols || c < 3) continue;
      const int row = m0 + r;
      int pt;
      if (group_all) pt = row;
      else {
        int i = __ldg(nbr + row);
        i = min(max(i, 0), N - 1);
        pt = (row / K / S) * N + i;
      }
      atomicAdd(grad_feats + (size_t)pt * D + (c - 3), Cs[r * ld + col]);
    }
  }
};

// ---------------------------------------------------------------------------------------------
// fp32 CUDA-core GEMM, C[M x Ncols] = A'[M x Kdim] * B^T, 128 x BN tile, 256 threads, 8 x TN
// register tile.  BT=false: B[n][k] = Bmat[n*ldb + k]  (forward, Bmat = W [Cout][Cin])
//                 BT=true : B[n][k] = Bmat[k*ldb + n]  (dgrad,   Bmat = W [Cout][Cin], n over Cin)
// ---------------------------------------------------------------------------------------------
constexpr int kBM = 128, kBK = 16;

template <int TN>
constexpr size_t gemm_nt_smem() {
  return sizeof(float) * (kBK * kBM + kBK * (16 * TN + 4) + kBM * (16 * TN + 1));
}

template <class AProd, class Epi, int TN, bool BT>
__global__ void __launch_bounds__(256)
gemm_nt_kernel(const AProd ap, const float* __restrict__ Bmat, int ldb, const Epi epi, int M,
               int Ncols, int Kdim) {
  constexpr int BN = 16 * TN, LDB = BN + 4, LDC = BN + 1;
  extern __shared__ __align__(16) float smem[];
  float* As = smem;                // [kBK][kBM]
  float* Bs = As + kBK * kBM;      // [kBK][LDB]
  float* Cs = Bs + kBK * LDB;      // [kBM][LDC]

  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.x * kBM, n0 = blockIdx.y * BN;
  const int rows = min(kBM, M - m0), cols = min(BN, Ncols - n0);

  float acc[8][TN];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  const int a_row = tid & 127, a_kh = (tid >> 7) * 8;
  for (int k0 = 0; k0 < Kdim; k0 += kBK) {
    {
      float v[8];
      ap.load8(m0 + a_row, k0 + a_kh, v);
#pragma unroll
      for (int u = 0; u < 8; ++u) As[(a_kh + u) * kBM + a_row] = v[u];
    }
    for (int e = tid; e < BN * kBK; e += 256) {
      int n, kk;
      float v = 0.f;
      if (BT) { n = e % BN; kk = e / BN; }
      else { n = e / kBK; kk = e % kBK; }
      if (n < cols && k0 + kk < Kdim)
        v = BT ? __ldg(Bmat + (size_t)(k0 + kk) * ldb + n0 + n) : __ldg(Bmat + (size_t)(n0 + n) * ldb + k0 + kk);
      Bs[kk * LDB + n] = v;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < kBK; ++kk) {
      float a[8], b[TN];
      const float4 a0 = *reinterpret_cast<const float4*>(As + kk * kBM + ty * 8);
      const float4 a1 = *reinterpret_cast<const float4*>(As + kk * kBM + ty * 8 + 4);
      a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
#pragma unroll
      for (int h = 0; h < TN / 4; ++h) {
        const float4 bv = *reinterpret_cast<const float4*>(Bs + kk * LDB + h * 64 + tx * 4);
        b[4 * h] = bv.x; b[4 * h + 1] = bv.y; b[4 * h + 2] = bv.z; b[4 * h + 3] = bv.w;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
  // thread (ty,tx) owns rows ty*8..+7 and columns {h*64 + tx*4 + 0..3}
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) Cs[(ty * 8 + i) * LDC + (j / 4) * 64 + tx * 4 + (j & 3)] = acc[i][j];
  __syncthreads();
  epi.template run<BN>(Cs, LDC, m0, n0, rows, cols);
}

// fp32 CUDA-core weight-gradient GEMM: out[Ca x Cb] += sum_rows P[row][ca] * Q[row][cb].
// 64 x 64 output tile per CTA, rows split over blockIdx.z, fp32 atomics at the end.
constexpr int kWgRows = 32;

template <class PProd, class QProd>
__global__ void __launch_bounds__(256)
gemm_tn_kernel(const PProd pp, const QProd qp, float* __restrict__ out, int ldo, int M, int Ca, int Cb,
               int rows_per_split) {
  __shared__ __align__(16) float Ps[kWgRows][64];
  __shared__ __align__(16) float Qs[kWgRows][64];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int ca0 = blockIdx.x * 64, cb0 = blockIdx.y * 64;
  const int r_begin = blockIdx.z * rows_per_split, r_end = min(M, r_begin + rows_per_split);
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int lr = tid >> 3, lc = (tid & 7) * 8;
  for (int r0 = r_begin; r0 < r_end; r0 += kWgRows) {
    float v[8];
    const int row = r0 + lr;
    if (row < r_end) pp.load8(row, ca0 + lc, v);
    else {
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = 0.f;
    }
    *reinterpret_cast<float4*>(&Ps[lr][lc]) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(&Ps[lr][lc + 4]) = make_float4(v[4], v[5], v[6], v[7]);
    if (row < r_end) qp.load8(row, cb0 + lc, v);
    else {
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = 0.f;
    }
    *reinterpret_cast<float4*>(&Qs[lr][lc]) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(&Qs[lr][lc + 4]) = make_float4(v[4], v[5], v[6], v[7]);
    __syncthreads();
#pragma unroll
    for (int r = 0; r < kWgRows; ++r) {
      const float4 pv = *reinterpret_cast<const float4*>(&Ps[r][ty * 4]);
      const float4 qv = *reinterpret_cast<const float4*>(&Qs[r][tx * 4]);
      const float pa[4] = {pv.x, pv.y, pv.z, pv.w}, qa[4] = {qv.x, qv.y, qv.z, qv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(pa[i], qa[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int ca = ca0 + ty * 4 + i, cb = cb0 + tx * 4 + j;
      if (ca < Ca && cb < Cb) atomicAdd(out + (size_t)ca * ldo + cb, acc[i][j]);
    }
}

}  // namespace pcoe
