// trunk.cu — the 1024 -> 512 -> 256 -> heads trunk of the mixture-of-von-Mises model on B rows (B = clouds).
//
// Replaces fc1 / ln1 / ReLU / dropout / fc2 / ln2 / ReLU / dropout / head_pi, head_mu, head_kappa of
// PointNetPPMvM (models/pointnet_pp_mvM.py:56-66,79-84,91-105) and their autograd: in PyTorch this is ~25
// forward and ~60 backward launches (cuBLAS fp32 GEMMs at M = 64 rows take ~10 us each, plus bias reductions,
// LayerNorm, dropout, gradient-accumulation adds) = 215 us of a 930 us training step.  Five building blocks, all
// fp32 on CUDA cores (the whole trunk is 1.3 MFLOP per cloud; the point is launch count and latency, not FLOPs):
//   linear_fwd        Y[B,N] = X[B,K] W^T + b          up to 3 (W, b, N) segments share X (the three heads)
//   linear_bwd_dx     dX[B,K] = dY[B,N] W
//   linear_bwd_dw     dW[N,K] (+)= dY^T X, db (+)= sum_b dY
//   ln_relu_drop_fwd  out = dropout(relu(LayerNorm(x))), saves mean / rstd / keep-mask
//   ln_relu_drop_bwd  dx, dgamma, dbeta
#include "common.cuh"
#include "mvm_head.cuh"

namespace pcoe {

constexpr int kMaxSeg = 3;
struct LinSeg {
  const float* W[kMaxSeg];
  const float* b[kMaxSeg];
  float* dW[kMaxSeg];
  float* db[kMaxSeg];
  int N[kMaxSeg];
  int nseg, Ntot;
};

__device__ __forceinline__ void seg_of(const LinSeg& s, int n, int& seg, int& local) {
  seg = 0; local = n;
#pragma unroll
  for (int i = 0; i < kMaxSeg - 1; ++i)
    if (seg == i && i + 1 < s.nseg && local >= s.N[i]) { local -= s.N[i]; seg = i + 1; }
}

// Outputs of a multi-segment linear are stored SEGMENT-MAJOR: [B x N0][B x N1][B x N2], so that each head's output is
// a dense [B, N_i] tensor (for one segment this is plain row-major).  n = global output index, (seg, ln) = seg_of(n).
__device__ __forceinline__ size_t seg_major(const LinSeg& s, int B, int b, int n, int seg, int ln) {
  return (size_t)B * (n - ln) + (size_t)b * s.N[seg] + ln;
}

// ---- one tiled fp32 GEMM for the three products of a linear layer on few rows ----
//   C[M x N] (+)= A[M x Kc] * Bm[Kc x N], CTA tile 64 x 32, 256 threads x (4 x 2) outputs, contraction in chunks of 32
//   staged through shared memory, optional split over the contraction (gridDim.z slices, atomics into a zeroed C).
//   MODE 0  forward  C = Y (segment-major), A = X[b][k],        Bm(k, n) = W_n[k]     (split: partials, see below)
//   MODE 1  dgrad    C = dX[b][k'],         A = dY (seg-major), Bm(n, k') = W_n[k']
//   MODE 2  wgrad    C = dW_n[k'],          A(n, b) = dY,       Bm(b, k') = X[b][k']
constexpr int kTM = 64, kTN = 32, kTK = 32;
template <int MODE>
__global__ void __launch_bounds__(256)
linear_gemm_kernel(const float* __restrict__ P0, const float* __restrict__ P1, float* __restrict__ Cout, int B, int K, LinSeg s,
                   int accumulate) {
  // dimensions of this product
  const int M = MODE == 2 ? s.Ntot : B, N = MODE == 0 ? s.Ntot : K, Kc = MODE == 0 ? K : (MODE == 1 ? s.Ntot : B);
  __shared__ float As[kTK][kTM + 4];
  __shared__ float Bs[kTK][kTN + 4];
  const int m0 = blockIdx.y * kTM, n0 = blockIdx.x * kTN;
  const int kslice = (Kc + gridDim.z - 1) / gridDim.z, kbeg = blockIdx.z * kslice, kend = min(Kc, kbeg + kslice);
  const int tm = (threadIdx.x >> 4) * 4, tn = (threadIdx.x & 15) * 2;   // 16 x 16 threads, 4 x 2 outputs each
  float acc[4][2] = {};

  auto loadA = [&](int m, int k) -> float {        // A(m, k), zero outside
    if (m >= M || k >= kend) return 0.f;
    if (MODE == 0) return __ldg(P0 + (size_t)m * K + k);                       // X[b][k]
    int sg, ln;
    seg_of(s, MODE == 1 ? k : m, sg, ln);
    return __ldg(P0 + seg_major(s, B, MODE == 1 ? m : k, MODE == 1 ? k : m, sg, ln));   // dY(b, n)
  };
  auto loadB = [&](int k, int n) -> float {        // Bm(k, n), zero outside
    if (n >= N || k >= kend) return 0.f;
    if (MODE == 2) return __ldg(P1 + (size_t)k * K + n);                       // X[b][k']
    int sg, ln;
    seg_of(s, MODE == 0 ? n : k, sg, ln);
    return __ldg(s.W[sg] + (size_t)ln * K + (MODE == 0 ? k : n));              // W_n[k] / W_n[k']
  };

  // chunk loop, register-prefetched: the 12 global loads of chunk i+1 are in flight while chunk i is multiplied.
  // A tile: 64 x 32 elements, 8 per thread (MODE 0/1: memory contiguous along k -> threads run along k; MODE 2:
  // contiguous along m).  B tile: 32 x 32, 4 per thread (MODE 0: contiguous along k; MODE 1/2: along n).
  float ra[8], rb[4];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int e = threadIdx.x + i * 256;
      const int kk = MODE == 2 ? e >> 6 : e & 31, mm = MODE == 2 ? e & 63 : e >> 5;
      ra[i] = loadA(m0 + mm, k0 + kk);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int e = threadIdx.x + i * 256;
      const int kk = MODE == 0 ? e & 31 : e >> 5, nn = MODE == 0 ? e >> 5 : e & 31;
      rb[i] = loadB(k0 + kk, n0 + nn);
    }
  };
  if (kbeg < kend) fetch(kbeg);
  for (int k0 = kbeg; k0 < kend; k0 += kTK) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int e = threadIdx.x + i * 256;
      As[MODE == 2 ? e >> 6 : e & 31][MODE == 2 ? e & 63 : e >> 5] = ra[i];
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int e = threadIdx.x + i * 256;
      Bs[MODE == 0 ? e & 31 : e >> 5][MODE == 0 ? e >> 5 : e & 31] = rb[i];
    }
    __syncthreads();
    if (k0 + kTK < kend) fetch(k0 + kTK);
#pragma unroll
    for (int kk = 0; kk < kTK; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(&As[kk][tm]);
      const float2 bb = *reinterpret_cast<const float2*>(&Bs[kk][tn]);
      acc[0][0] = fmaf(a.x, bb.x, acc[0][0]); acc[0][1] = fmaf(a.x, bb.y, acc[0][1]);
      acc[1][0] = fmaf(a.y, bb.x, acc[1][0]); acc[1][1] = fmaf(a.y, bb.y, acc[1][1]);
      acc[2][0] = fmaf(a.z, bb.x, acc[2][0]); acc[2][1] = fmaf(a.z, bb.y, acc[2][1]);
      acc[3][0] = fmaf(a.w, bb.x, acc[3][0]); acc[3][1] = fmaf(a.w, bb.y, acc[3][1]);
    }
    __syncthreads();
  }

  const bool split = gridDim.z > 1;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + tm + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int n = n0 + tn + j;
      if (n >= N) continue;
      float v = acc[i][j];
      float* dst;
      if (MODE == 0) {
        // forward is deterministic: with a split contraction every slice writes its own partial result
        // [slice][B x Ntot] (no bias); the consumer (ln_relu_drop_fwd) adds the slices in a fixed order
        int sg, ln;
        seg_of(s, n, sg, ln);
        if (!split && s.b[sg]) v += __ldg(s.b[sg] + ln);
        Cout[(size_t)blockIdx.z * B * s.Ntot + seg_major(s, B, m, n, sg, ln)] = v;
        continue;
      } else if (MODE == 1) {
        dst = Cout + (size_t)m * K + n;
      } else {
        int sg, ln;
        seg_of(s, m, sg, ln);
        dst = s.dW[sg] + (size_t)ln * K + n;
      }
      if (split) atomicAdd(dst, v);
      else *dst = (MODE == 2 && accumulate) ? *dst + v : v;
    }
  }
}

// Fast path of linear_gemm_kernel for ONE weight segment with K and N multiples of 4 and M a multiple of... nothing:
// every operand is a plain strided matrix, tiles are fetched with 16-byte loads through precomputed pointers
// (3 LDG.128 per thread and chunk instead of 12 scalar loads with per-element segment lookups).
//   Aop(m, k) = PA[m * sam + k * sak],  Bop(k, n) = PB[k * sbk + n * sbn],  C(m, n) = PC[m * ldc + n]
//   AK / BK: the operand is contiguous along k (transposed into the k-major shared tiles), else along m / n.
template <bool AK, bool BK>
__global__ void __launch_bounds__(256)
linear_gemm_fast_kernel(const float* __restrict__ PA, size_t sam, size_t sak, const float* __restrict__ PB, size_t sbk,
                        size_t sbn, float* __restrict__ PC, size_t ldc, size_t part_stride, int M, int N, int Kc,
                        const float* __restrict__ bias, int atomic_out, int accumulate) {
  __shared__ __align__(16) float As[kTK][kTM + 4];
  __shared__ __align__(16) float Bs[kTK][kTN + 4];
  const int m0 = blockIdx.y * kTM, n0 = blockIdx.x * kTN;
  const int kslice = ((Kc + gridDim.z - 1) / gridDim.z + kTK - 1) / kTK * kTK, kbeg = blockIdx.z * kslice;
  const int kend = min(Kc, kbeg + kslice);
  const int tm = (threadIdx.x >> 4) * 4, tn = (threadIdx.x & 15) * 2;
  float acc[4][2] = {};
  // per-thread fetch coordinates (fixed across chunks)
  int am[2], ak[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int e = threadIdx.x + i * 256;
    if (AK) { am[i] = e >> 3; ak[i] = (e & 7) * 4; } else { ak[i] = e >> 4; am[i] = (e & 15) * 4; }
  }
  const int bk = BK ? (threadIdx.x & 7) * 4 : threadIdx.x >> 3, bn = BK ? threadIdx.x >> 3 : (threadIdx.x & 7) * 4;
  float4 ra[2], rb;
  auto fetch = [&](int k0) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      ra[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      const int m = m0 + am[i], k = k0 + ak[i];
      if (AK) { if (m < M && k < kend) ra[i] = __ldg(reinterpret_cast<const float4*>(PA + (size_t)m * sam + k)); }
      else { if (m < M && k < kend) ra[i] = __ldg(reinterpret_cast<const float4*>(PA + (size_t)k * sak + m)); }
    }
    rb = make_float4(0.f, 0.f, 0.f, 0.f);
    const int k = k0 + bk, n = n0 + bn;
    if (BK) { if (n < N && k < kend) rb = __ldg(reinterpret_cast<const float4*>(PB + (size_t)n * sbn + k)); }
    else { if (n < N && k < kend) rb = __ldg(reinterpret_cast<const float4*>(PB + (size_t)k * sbk + n)); }
  };
  if (kbeg < kend) fetch(kbeg);
  for (int k0 = kbeg; k0 < kend; k0 += kTK) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      if (AK) { As[ak[i]][am[i]] = ra[i].x; As[ak[i] + 1][am[i]] = ra[i].y; As[ak[i] + 2][am[i]] = ra[i].z; As[ak[i] + 3][am[i]] = ra[i].w; }
      else *reinterpret_cast<float4*>(&As[ak[i]][am[i]]) = ra[i];
    }
    if (BK) { Bs[bk][bn] = rb.x; Bs[bk + 1][bn] = rb.y; Bs[bk + 2][bn] = rb.z; Bs[bk + 3][bn] = rb.w; }
    else *reinterpret_cast<float4*>(&Bs[bk][bn]) = rb;
    __syncthreads();
    if (k0 + kTK < kend) fetch(k0 + kTK);
#pragma unroll
    for (int kk = 0; kk < kTK; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(&As[kk][tm]);
      const float2 bb = *reinterpret_cast<const float2*>(&Bs[kk][tn]);
      acc[0][0] = fmaf(a.x, bb.x, acc[0][0]); acc[0][1] = fmaf(a.x, bb.y, acc[0][1]);
      acc[1][0] = fmaf(a.y, bb.x, acc[1][0]); acc[1][1] = fmaf(a.y, bb.y, acc[1][1]);
      acc[2][0] = fmaf(a.z, bb.x, acc[2][0]); acc[2][1] = fmaf(a.z, bb.y, acc[2][1]);
      acc[3][0] = fmaf(a.w, bb.x, acc[3][0]); acc[3][1] = fmaf(a.w, bb.y, acc[3][1]);
    }
    __syncthreads();
  }
  float* C = PC + (size_t)blockIdx.z * part_stride;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + tm + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int n = n0 + tn + j;
      if (n >= N) continue;
      float v = acc[i][j];
      if (bias) v += __ldg(bias + n);
      float* dst = C + (size_t)m * ldc + n;
      if (atomic_out) atomicAdd(dst, v);
      else *dst = accumulate ? *dst + v : v;
    }
  }
}

// bias gradient: db_n (+)= sum_b dY(b, n); one warp per output
__global__ void linear_bias_grad_kernel(const float* __restrict__ dY, int B, LinSeg s, int accumulate) {
  const int n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (n >= s.Ntot) return;
  int sg, ln;
  seg_of(s, n, sg, ln);
  if (!s.db[sg]) return;
  float t = 0.f;
  for (int b = lane; b < B; b += 32) t += __ldg(dY + seg_major(s, B, b, n, sg, ln));
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) t += __shfl_xor_sync(0xFFFFFFFFu, t, m);
  if (lane == 0) s.db[sg][ln] = accumulate ? s.db[sg][ln] + t : t;
}

// Philox4x32-10 (Salmon et al. 2011), counter (i, j, offset), key = seed; one 32-bit output
__device__ __forceinline__ uint32_t philox_u32(uint32_t i, uint32_t j, uint64_t seed, uint64_t offset) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
  uint4 ctr = make_uint4(i, j, (uint32_t)offset, (uint32_t)(offset >> 32));
  uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x, hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += W0; key.y += W1;
  }
  return ctr.x;
}

// block-wide sum of one value per thread (blockDim.x = 256), result broadcast to every thread
__device__ __forceinline__ float block_sum256(float v, float* red) {
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, m);
  __syncthreads();                       // protects `red` against the previous use
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) t += red[w];
  return t;
}

// x may arrive as `nparts` partial sums [part][B x N] of a split linear_fwd plus a bias vector: they are added in a
// fixed order (deterministic) and the sum is written to h (the pre-LayerNorm activation saved for backward).
// One CTA (256 threads) per row, up to kLnPer elements per thread in registers (N <= 256 * kLnPer).
constexpr int kLnPer = 4;
constexpr int kTailMaxOut = 64;
// Optional tail of the row kernels (the trunk's last stage): the up-to-three small linear heads that share the row
// (Ntot <= 64 outputs, written segment-major like linear_fwd) and, when K > 0, the mixture head transform of the
// row (N = {K, 2K, K}: pi | mu_raw | kappa_raw).  Everything after the last LayerNorm is per-row work: fusing it
// replaces a one-CTA GEMM launch (13 us) and the head-transform launch (7 us) by ~1 us inside this kernel.
struct HeadTail {
  LinSeg seg;          // W, b, N (forward) ; W, N (backward)
  float* raw;          // forward: [B x N0][B x N1][B x N2] linear outputs ; backward: the saved outputs (read)
  float* d_raw;        // backward: gradient w.r.t. raw (written, same layout)
  int K;               // > 0: mixture head transform
  float temp, kappa_max;
  int clamp_kappa;
  float *weight, *mu, *kappa;              // forward outputs [B,K]
  const float *g_w, *g_mu, *g_k;           // backward upstream gradients [B,K] (may be NULL)
};

template <bool TAIL>
__global__ void __launch_bounds__(256)
ln_relu_drop_fwd_kernel(const float* __restrict__ x, int nparts, const float* __restrict__ xbias, float* __restrict__ h,
                        const float* __restrict__ gamma, const float* __restrict__ beta,
                        int B, int N, float eps, float p, int train, uint64_t seed, const uint64_t* __restrict__ counter,
                        float* __restrict__ out, float* __restrict__ mean, float* __restrict__ rstd,
                        uint8_t* __restrict__ mask, HeadTail tail) {
  __shared__ float red[8];
  __shared__ float rowbuf[TAIL ? 256 * kLnPer : 1];
  __shared__ float rawbuf[TAIL ? kTailMaxOut : 1];
  const int row = blockIdx.x;
  float v[kLnPer];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < kLnPer; ++i) {
    const int n = threadIdx.x + i * 256;
    v[i] = 0.f;
    if (n < N) {
      if (h) {
        float pv[8];
#pragma unroll
        for (int z = 0; z < 8; ++z) pv[z] = z < nparts ? __ldg(x + ((size_t)z * B + row) * N + n) : 0.f;   // independent loads
        float t = xbias ? xbias[n] : 0.f;
#pragma unroll
        for (int z = 0; z < 8; ++z) t += pv[z];                  // fixed order: deterministic
        for (int z = 8; z < nparts; ++z) t += x[((size_t)z * B + row) * N + n];
        h[(size_t)row * N + n] = t;
        v[i] = t;
      } else {
        v[i] = x[(size_t)row * N + n];
      }
      s += v[i];
    }
  }
  const float mu = block_sum256(s, red) / N;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < kLnPer; ++i)
    if (threadIdx.x + i * 256 < N) { const float d = v[i] - mu; q = fmaf(d, d, q); }
  const float rs = rsqrtf(block_sum256(q, red) / N + eps);
  if (threadIdx.x == 0 && mean) { mean[row] = mu; rstd[row] = rs; }
  const bool drop = train && p > 0.f;
  const float keep_scale = drop ? 1.f / (1.f - p) : 1.f;
  const uint32_t thr = drop ? (uint32_t)fminf(p * 4294967296.f, 4294967295.f) : 0u;
  const uint64_t off = counter ? *counter : 0ull;
#pragma unroll
  for (int i = 0; i < kLnPer; ++i) {
    const int n = threadIdx.x + i * 256;
    if (n >= N) continue;
    const float y = fmaxf(fmaf((v[i] - mu) * rs, gamma[n], beta[n]), 0.f);
    bool keep = true;
    if (drop) keep = philox_u32((uint32_t)n, (uint32_t)row, seed, off) >= thr;
    if (mask) mask[(size_t)row * N + n] = keep ? 1 : 0;
    const float o = keep ? y * keep_scale : 0.f;
    out[(size_t)row * N + n] = o;
    if constexpr (TAIL) rowbuf[n] = o;
  }
  if constexpr (TAIL) {
    __syncthreads();
    const LinSeg& sg = tail.seg;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int o = warp; o < sg.Ntot; o += 8) {          // one warp per output: coalesced reads of the weight row
      int si, lo;
      seg_of(sg, o, si, lo);
      const float* w = sg.W[si] + (size_t)lo * N;
      float a0 = 0.f, a1 = 0.f;
      int n = lane;
      for (; n + 32 < N; n += 64) { a0 = fmaf(rowbuf[n], __ldg(w + n), a0); a1 = fmaf(rowbuf[n + 32], __ldg(w + n + 32), a1); }
      if (n < N) a0 = fmaf(rowbuf[n], __ldg(w + n), a0);
      float a = a0 + a1;
#pragma unroll
      for (int m = 16; m >= 1; m >>= 1) a += __shfl_xor_sync(0xFFFFFFFFu, a, m);
      if (lane == 0) {
        a += sg.b[si] ? sg.b[si][lo] : 0.f;
        tail.raw[seg_major(sg, B, row, o, si, lo)] = a;
        rawbuf[o] = a;
      }
    }
    if (tail.K > 0) {
      __syncthreads();
      if (threadIdx.x == 0) {
        const int K = tail.K;
        mvm_head_fwd_row(rawbuf, rawbuf + K, rawbuf + 3 * K, K, tail.temp, tail.kappa_max, tail.clamp_kappa,
                         tail.weight + (size_t)row * K, tail.mu + (size_t)row * K, tail.kappa + (size_t)row * K);
      }
    }
  }
}

// TAIL: dout is not read - it is the heads' data gradient of this row, derived here from the upstream gradients of
// the head transform:  d_raw = head_bwd(raw, g_w, g_mu, g_k)  (written for the heads' weight-gradient kernel),
// dout[n] = sum_o d_raw[o] * W_o[n].
template <bool TAIL>
__global__ void __launch_bounds__(256)
ln_relu_drop_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ x, const float* __restrict__ out,
                        const float* __restrict__ gamma, const float* __restrict__ mean, const float* __restrict__ rstd,
                        const uint8_t* __restrict__ mask, int B, int N, float p, int train, float* __restrict__ dx,
                        float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dxbias, HeadTail tail) {
  __shared__ float red[8];
  __shared__ float drawbuf[TAIL ? kTailMaxOut : 1];
  const int row = blockIdx.x;
  if constexpr (TAIL) {
    const LinSeg& sg = tail.seg;
    const int K = tail.K;
    if (threadIdx.x == 0) {      // K > 0 (checked on the host): raw = pi [B,K] | mu_raw [B,2K] | kappa_raw [B,K]
      float rw[4 * kHeadMaxK];
      const float* r0 = tail.raw;
      for (int k = 0; k < K; ++k) rw[k] = r0[(size_t)row * K + k];
      for (int k = 0; k < 2 * K; ++k) rw[K + k] = r0[(size_t)B * K + (size_t)row * 2 * K + k];
      for (int k = 0; k < K; ++k) rw[3 * K + k] = r0[(size_t)B * 3 * K + (size_t)row * K + k];
      mvm_head_bwd_row(rw, rw + K, rw + 3 * K, K, tail.temp, tail.kappa_max, tail.clamp_kappa,
                       tail.g_w ? tail.g_w + (size_t)row * K : nullptr, tail.g_mu ? tail.g_mu + (size_t)row * K : nullptr,
                       tail.g_k ? tail.g_k + (size_t)row * K : nullptr, drawbuf, drawbuf + K, drawbuf + 3 * K);
      float* d0 = tail.d_raw;
      for (int k = 0; k < K; ++k) d0[(size_t)row * K + k] = drawbuf[k];
      for (int k = 0; k < 2 * K; ++k) d0[(size_t)B * K + (size_t)row * 2 * K + k] = drawbuf[K + k];
      for (int k = 0; k < K; ++k) d0[(size_t)B * 3 * K + (size_t)row * K + k] = drawbuf[3 * K + k];
    }
    __syncthreads();
    (void)sg;
  }
  const float keep_scale = (train && p > 0.f) ? 1.f / (1.f - p) : 1.f;
  const float mu = mean[row], rs = rstd[row];
  const size_t o = (size_t)row * N;
  // dz = gradient at the LayerNorm output (after the dropout and ReLU masks); s1 = sum dz*gamma, s2 = sum dz*gamma*xhat
  float dzg[kLnPer], xh[kLnPer];
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int i = 0; i < kLnPer; ++i) {
    const int n = threadIdx.x + i * 256;
    dzg[i] = xh[i] = 0.f;
    if (n >= N) continue;
    const bool on = out[o + n] > 0.f;                     // ReLU and dropout both zero the output
    float dov;
    if constexpr (TAIL) {
      const LinSeg& sg = tail.seg;
      float a = 0.f;
      int oo = 0;
      for (int si = 0; si < sg.nseg; ++si)
        for (int lo = 0; lo < sg.N[si]; ++lo, ++oo) a = fmaf(drawbuf[oo], __ldg(sg.W[si] + (size_t)lo * N + n), a);
      dov = a;
    } else {
      dov = dout[o + n];
    }
    const float dz = on ? dov * keep_scale * (mask ? (float)mask[o + n] : 1.f) : 0.f;
    xh[i] = (x[o + n] - mu) * rs;
    dzg[i] = dz * gamma[n];
    s1 += dzg[i];
    s2 = fmaf(dzg[i], xh[i], s2);
    if (dz != 0.f) {
      atomicAdd(dgamma + n, dz * xh[i]);
      atomicAdd(dbeta + n, dz);
    }
  }
  const float m1 = block_sum256(s1, red) / N, m2 = block_sum256(s2, red) / N;
#pragma unroll
  for (int i = 0; i < kLnPer; ++i) {
    const int n = threadIdx.x + i * 256;
    if (n >= N) continue;
    const float g = rs * (dzg[i] - m1 - xh[i] * m2);
    dx[o + n] = g;
    if (dxbias) atomicAdd(dxbias + n, g);     // bias gradient of the linear layer that produced x (column sums of dx)
  }
}

static int make_seg(LinSeg& s, int nseg, const float* const* W, const float* const* b, float* const* dW, float* const* db,
                    const int* N) {
  if (nseg < 1 || nseg > kMaxSeg) return fail(PCOE_ERR_BAD_SHAPE, "linear: nseg=%d (1..%d)", nseg, kMaxSeg);
  s.nseg = nseg; s.Ntot = 0;
  for (int i = 0; i < kMaxSeg; ++i) {
    s.W[i] = i < nseg && W ? W[i] : nullptr; s.b[i] = i < nseg && b ? b[i] : nullptr;
    s.dW[i] = i < nseg && dW ? dW[i] : nullptr; s.db[i] = i < nseg && db ? db[i] : nullptr;
    s.N[i] = i < nseg ? N[i] : 0;
    if (i < nseg && N[i] <= 0) return fail(PCOE_ERR_BAD_SHAPE, "linear: N[%d]=%d", i, N[i]);
    s.Ntot += s.N[i];
  }
  return PCOE_OK;
}

}  // namespace pcoe

using namespace pcoe;

static int split_of(int Kc, int tiles) {   // slices of the contraction so that ~one wave of CTAs is in flight
  int z = kNumSMs / (tiles > 0 ? tiles : 1);
  const int maxz = ceil_div(Kc, 128);        // at least 128 contraction steps per slice
  z = z < 1 ? 1 : (z > maxz ? maxz : z);
  return z;
}

extern "C" int pcoe_linear_fwd(const float* x, int B, int K, int nseg, const float* const* W, const float* const* bias,
                               const int* N, float* y, int max_parts, int* nparts, void* stream) {
  if (B <= 0 || K <= 0) return fail(PCOE_ERR_BAD_SHAPE, "linear_fwd: B=%d K=%d", B, K);
  if (!x || !W || !N || !y) return fail(PCOE_ERR_NULL, "linear_fwd: NULL pointer");
  LinSeg s;
  PCOE_TRY(make_seg(s, nseg, W, bias, nullptr, nullptr, N));
  for (int i = 0; i < nseg; ++i) if (!W[i]) return fail(PCOE_ERR_NULL, "linear_fwd: W[%d] is NULL", i);
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid(ceil_div(s.Ntot, kTN), ceil_div(B, kTM), 1);
  int z = split_of(K, grid.x * grid.y);
  z = (max_parts < 1 || !nparts) ? 1 : (z > max_parts ? max_parts : z);
  grid.z = z;
  if (nparts) *nparts = z;
  LaunchScope ls("linear_fwd_kernel", st);
  if (nseg == 1 && (K & 3) == 0)      // Y[b][n] = sum_k X[b][k] W[n][k]: both operands contiguous along k
    linear_gemm_fast_kernel<true, true><<<grid, 256, 0, st>>>(x, (size_t)K, 1, W[0], 1, (size_t)K, y, (size_t)s.Ntot,
                                                               (size_t)B * s.Ntot, B, s.Ntot, K, z == 1 ? s.b[0] : nullptr, 0, 0);
  else
    linear_gemm_kernel<0><<<grid, 256, 0, st>>>(x, nullptr, y, B, K, s, 0);
  return ls.done();
}

extern "C" int pcoe_linear_bwd_dx(const float* dy, int B, int K, int nseg, const float* const* W, const int* N, float* dx,
                                  void* stream) {
  if (B <= 0 || K <= 0) return fail(PCOE_ERR_BAD_SHAPE, "linear_bwd_dx: B=%d K=%d", B, K);
  if (!dy || !W || !N || !dx) return fail(PCOE_ERR_NULL, "linear_bwd_dx: NULL pointer");
  LinSeg s;
  PCOE_TRY(make_seg(s, nseg, W, nullptr, nullptr, nullptr, N));
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid(ceil_div(K, kTN), ceil_div(B, kTM), 1);
  grid.z = split_of(s.Ntot, grid.x * grid.y);
  if (grid.z > 1) PCOE_CUDA(cudaMemsetAsync(dx, 0, sizeof(float) * (size_t)B * K, st));
  LaunchScope ls("linear_bwd_dx_kernel", st);
  if (nseg == 1 && (K & 3) == 0 && (s.Ntot & 3) == 0)   // dX[b][k'] = sum_n dY[b][n] W[n][k']: A along its k (= n), B along n (= k')
    linear_gemm_fast_kernel<true, false><<<grid, 256, 0, st>>>(dy, (size_t)s.Ntot, 1, W[0], (size_t)K, 1, dx, (size_t)K, 0, B, K,
                                                                s.Ntot, nullptr, grid.z > 1, 0);
  else
    linear_gemm_kernel<1><<<grid, 256, 0, st>>>(dy, nullptr, dx, B, K, s, 0);
  return ls.done();
}

extern "C" int pcoe_linear_bwd_dw(const float* dy, const float* x, int B, int K, int nseg, const int* N, float* const* dW,
                                  float* const* dbias, int accumulate, void* stream) {
  if (B <= 0 || K <= 0) return fail(PCOE_ERR_BAD_SHAPE, "linear_bwd_dw: B=%d K=%d", B, K);
  if (!dy || !x || !N || !dW) return fail(PCOE_ERR_NULL, "linear_bwd_dw: NULL pointer");
  LinSeg s;
  PCOE_TRY(make_seg(s, nseg, nullptr, nullptr, dW, dbias, N));
  for (int i = 0; i < nseg; ++i) if (!dW[i]) return fail(PCOE_ERR_NULL, "linear_bwd_dw: dW[%d] is NULL", i);
  cudaStream_t st = (cudaStream_t)stream;
  {
    LaunchScope ls("linear_bwd_dw_kernel", st);
    dim3 grid(ceil_div(K, kTN), ceil_div(s.Ntot, kTM), 1);   // contraction = the B rows: no split
    if (nseg == 1 && (K & 3) == 0 && (s.Ntot & 3) == 0)   // dW[n][k'] = sum_b dY[b][n] X[b][k']: A along m (= n), B along n (= k')
      linear_gemm_fast_kernel<false, false><<<grid, 256, 0, st>>>(dy, 1, (size_t)s.Ntot, x, (size_t)K, 1, dW[0], (size_t)K, 0, s.Ntot,
                                                                   K, B, nullptr, 0, accumulate);
    else
      linear_gemm_kernel<2><<<grid, 256, 0, st>>>(dy, x, nullptr, B, K, s, accumulate);
    PCOE_TRY(ls.done());
  }
  bool any = false;
  for (int i = 0; i < nseg; ++i) any = any || (dbias && dbias[i]);
  if (any) {
    LaunchScope ls("linear_bias_grad_kernel", st);
    linear_bias_grad_kernel<<<ceil_div(s.Ntot * 32, 128), 128, 0, st>>>(dy, B, s, accumulate);
    PCOE_TRY(ls.done());
  }
  return PCOE_OK;
}

extern "C" int pcoe_ln_relu_dropout_fwd(const float* x, int nparts, const float* xbias, float* h, const float* gamma,
                                        const float* beta, int B, int N, float eps, float p, int train, uint64_t seed,
                                        const uint64_t* counter_dev, float* out, float* mean, float* rstd, uint8_t* mask,
                                        void* stream) {
  if (B <= 0 || N <= 0 || !(p >= 0.f && p < 1.f)) return fail(PCOE_ERR_BAD_SHAPE, "ln_relu_dropout: B=%d N=%d p=%g", B, N, p);
  if (!x || !gamma || !beta || !out) return fail(PCOE_ERR_NULL, "ln_relu_dropout_fwd: NULL pointer");
  if ((nparts > 1 || xbias) && !h) return fail(PCOE_ERR_NULL, "ln_relu_dropout_fwd: partial sums / bias need the h output");
  cudaStream_t st = (cudaStream_t)stream;
  LaunchScope ls("ln_relu_drop_fwd_kernel", st);
  if (N > 256 * kLnPer) return fail(PCOE_ERR_UNSUPPORTED, "ln_relu_dropout: N=%d > %d", N, 256 * kLnPer);
  ln_relu_drop_fwd_kernel<false><<<B, 256, 0, st>>>(x, nparts < 1 ? 1 : nparts, xbias, h, gamma, beta, B, N, eps, p, train,
                                                          seed, counter_dev, out, mean, rstd, mask, HeadTail{});
  return ls.done();
}

static int check_tail(const char* what, int nseg, const int* Nout, int K) {
  if (nseg < 1 || nseg > kMaxSeg || !Nout) return fail(PCOE_ERR_BAD_SHAPE, "%s: nseg=%d", what, nseg);
  int tot = 0;
  for (int i = 0; i < nseg; ++i) tot += Nout[i];
  if (tot > kTailMaxOut) return fail(PCOE_ERR_UNSUPPORTED, "%s: %d head outputs > %d", what, tot, kTailMaxOut);
  if (K > 0 && (K > kHeadMaxK || nseg != 3 || Nout[0] != K || Nout[1] != 2 * K || Nout[2] != K))
    return fail(PCOE_ERR_BAD_SHAPE, "%s: the mixture head needs segments (K, 2K, K) with K <= %d", what, kHeadMaxK);
  return PCOE_OK;
}

extern "C" int pcoe_ln_relu_dropout_heads_fwd(const float* x, int nparts, const float* xbias, float* h, const float* gamma,
                                              const float* beta, int B, int N, float eps, float p, int train,
                                              uint64_t seed, const uint64_t* counter_dev, float* out, float* mean,
                                              float* rstd, uint8_t* mask, int nseg, const float* const* W,
                                              const float* const* bias, const int* Nout, float* raw, int K, float temp,
                                              float kappa_max, int clamp_kappa, float* weight, float* mu, float* kappa,
                                              void* stream) {
  if (B <= 0 || N <= 0 || !(p >= 0.f && p < 1.f)) return fail(PCOE_ERR_BAD_SHAPE, "ln_heads_fwd: B=%d N=%d p=%g", B, N, p);
  if (!x || !gamma || !beta || !out || !W || !raw) return fail(PCOE_ERR_NULL, "ln_heads_fwd: NULL pointer");
  if ((nparts > 1 || xbias) && !h) return fail(PCOE_ERR_NULL, "ln_heads_fwd: partial sums / bias need the h output");
  if (N > 256 * kLnPer) return fail(PCOE_ERR_UNSUPPORTED, "ln_heads_fwd: N=%d > %d", N, 256 * kLnPer);
  PCOE_TRY(check_tail("ln_heads_fwd", nseg, Nout, K));
  if (K > 0 && (!(temp > 0.f) || !weight || !mu || !kappa)) return fail(PCOE_ERR_NULL, "ln_heads_fwd: head outputs / temp");
  HeadTail t{};
  PCOE_TRY(make_seg(t.seg, nseg, W, bias, nullptr, nullptr, Nout));
  for (int i = 0; i < nseg; ++i) if (!W[i]) return fail(PCOE_ERR_NULL, "ln_heads_fwd: W[%d] is NULL", i);
  t.raw = raw; t.K = K; t.temp = temp; t.kappa_max = kappa_max; t.clamp_kappa = clamp_kappa;
  t.weight = weight; t.mu = mu; t.kappa = kappa;
  cudaStream_t st = (cudaStream_t)stream;
  LaunchScope ls("ln_heads_fwd_kernel", st);
  ln_relu_drop_fwd_kernel<true><<<B, 256, 0, st>>>(x, nparts < 1 ? 1 : nparts, xbias, h, gamma, beta, B, N, eps, p, train,
                                                   seed, counter_dev, out, mean, rstd, mask, t);
  return ls.done();
}

extern "C" int pcoe_ln_relu_dropout_bwd(const float* dout, const float* x, const float* out, const float* gamma,
                                        const float* mean, const float* rstd, const uint8_t* mask, int B, int N, float p,
                                        int train, float* dx, float* dgamma, float* dbeta, float* dxbias, void* stream) {
  if (B <= 0 || N <= 0) return fail(PCOE_ERR_BAD_SHAPE, "ln_relu_dropout_bwd: B=%d N=%d", B, N);
  if (!dout || !x || !out || !gamma || !mean || !rstd || !dx || !dgamma || !dbeta)
    return fail(PCOE_ERR_NULL, "ln_relu_dropout_bwd: NULL pointer");
  cudaStream_t st = (cudaStream_t)stream;
  LaunchScope ls("ln_relu_drop_bwd_kernel", st);
  if (N > 256 * kLnPer) return fail(PCOE_ERR_UNSUPPORTED, "ln_relu_dropout: N=%d > %d", N, 256 * kLnPer);
  ln_relu_drop_bwd_kernel<false><<<B, 256, 0, st>>>(dout, x, out, gamma, mean, rstd, mask, B, N, p, train, dx, dgamma,
                                                          dbeta, dxbias, HeadTail{});
  return ls.done();
}

extern "C" int pcoe_heads_ln_relu_dropout_bwd(const float* raw, int K, float temp, float kappa_max, int clamp_kappa,
                                              const float* g_w, const float* g_mu, const float* g_k, float* d_raw,
                                              int nseg, const float* const* W, const int* Nout, const float* x,
                                              const float* out, const float* gamma, const float* mean, const float* rstd,
                                              const uint8_t* mask, int B, int N, float p, int train, float* dx,
                                              float* dgamma, float* dbeta, float* dxbias, void* stream) {
  if (B <= 0 || N <= 0) return fail(PCOE_ERR_BAD_SHAPE, "heads_ln_bwd: B=%d N=%d", B, N);
  if (!raw || !d_raw || !W || !x || !out || !gamma || !mean || !rstd || !dx || !dgamma || !dbeta)
    return fail(PCOE_ERR_NULL, "heads_ln_bwd: NULL pointer");
  if (N > 256 * kLnPer) return fail(PCOE_ERR_UNSUPPORTED, "heads_ln_bwd: N=%d > %d", N, 256 * kLnPer);
  if (K <= 0 || !(temp > 0.f)) return fail(PCOE_ERR_BAD_SHAPE, "heads_ln_bwd: K=%d temp=%g", K, temp);
  PCOE_TRY(check_tail("heads_ln_bwd", nseg, Nout, K));
  HeadTail t{};
  PCOE_TRY(make_seg(t.seg, nseg, W, nullptr, nullptr, nullptr, Nout));
  for (int i = 0; i < nseg; ++i) if (!W[i]) return fail(PCOE_ERR_NULL, "heads_ln_bwd: W[%d] is NULL", i);
  t.raw = const_cast<float*>(raw); t.d_raw = d_raw; t.K = K; t.temp = temp; t.kappa_max = kappa_max;
  t.clamp_kappa = clamp_kappa; t.g_w = g_w; t.g_mu = g_mu; t.g_k = g_k;
  cudaStream_t st = (cudaStream_t)stream;
  LaunchScope ls("heads_ln_bwd_kernel", st);
  ln_relu_drop_bwd_kernel<true><<<B, 256, 0, st>>>(nullptr, x, out, gamma, mean, rstd, mask, B, N, p, train, dx, dgamma,
                                                   dbeta, dxbias, t);
  return ls.done();
}
