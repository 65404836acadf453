// sa_tc6.cuh — the fp32-ACCURATE tensor-core mode of the set-abstraction MLP ("bf16x3").
//
// The reference computes its 1x1 convolutions in fp32 (models/pointnet_pp_8dir.py:40-41).  Plain bf16 operands
// (sa_tc4/5.cuh) keep the loss within 1e-3 but move the max-pool / ReLU routing and with it the weight gradients by
// tens of percent.  This mode keeps the tcgen05 pipeline and recovers fp32-class accuracy by SPLITTING
// every operand into bf16 planes and issuing one MMA per retained plane product into the same fp32 TMEM accumulator:
//   backward GEMMs (dgrad, wgrad - linear in their operands once the routing is fixed): TWO planes,
//        x = hi + lo,  hi = bf16(x), lo = bf16(x - hi)  (16 significant bits)
//        A*B  ~=  A_lo*B_hi + A_hi*B_lo + A_hi*B_hi                      (3 MMAs; dropped lo*lo is 2^-16 relative)
//   forward GEMMs (their rounding noise decides the arg-max / ReLU routing, and a re-routed element moves the
//   gradients by O(1): measured, 16-bit forward operands leave the weight gradients 1-3e-2 from the fp64 oracle,
//   24-bit ones 4e-4 .. 4e-3): THREE planes, x = hi + mid + lo exactly (24 bits),
//        A*B  ~=  hi*hi + (hi*mid + mid*hi) + (mid*mid + hi*lo + lo*hi)  (6 MMAs; dropped terms are 2^-24 relative)
// Everything that is not an MMA operand stays fp32: activations live in HBM as fp32 (tile-blocked, point-quad
// interleaved: see act_off), BatchNorm / ReLU / BatchNorm-backward transforms run in fp32 in the producers BEFORE the
// split, batch statistics are fp64 sums of the fp32 accumulators.
//
// Orientation, roles and operand images are those of sa_tc5.cuh (channels on the TMEM lanes; 17 warps: 8 epilogue,
// 8 producers, 1 MMA issue; the contraction streamed through a ring of shared-memory stages in 64-channel - or, for the
// weight gradient, 64-point - parts of 16 KB per plane).  What is new besides the planes:
//   * forward / dgrad kernels are PERSISTENT: CTA x owns output-channel block x % ncb and walks over the 128-point
//     tiles x / ncb, x / ncb + gridDim.x / ncb, ... with double-buffered TMEM accumulators, so the prologue (TMEM
//     allocation, barriers, inline BatchNorm finalisation) is paid once per SM and the epilogue of tile i overlaps the
//     MMAs of tile i+1;
//   * when the weight slice of the CTA's channel block fits beside the stage ring (SA1, SA2) it is loaded ONCE and stays
//     resident (both planes); otherwise (SA3) it is streamed with the activations as in sa_tc5.cuh.
#pragma once
#include "sa_tc5.cuh"

namespace pcoe {
namespace v6 {

using namespace v4;
using v5::kPart;
using v5::unit_pipeline;

constexpr int kMaxStages6 = 4;

// ---- operand split -----------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t bf2_pack(float lo, float hi) {
  uint32_t d;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}
// (a, b) -> packed hi plane and packed lo plane
__device__ __forceinline__ void split2(float a, float b, uint32_t& h, uint32_t& l) {
  h = bf2_pack(a, b);
  l = bf2_pack(a - __uint_as_float(h << 16), b - __uint_as_float(h & 0xFFFF0000u));
}
__device__ __forceinline__ void split8(const float (&v)[8], uint4& h, uint4& l) {
  split2(v[0], v[1], h.x, l.x); split2(v[2], v[3], h.y, l.y);
  split2(v[4], v[5], h.z, l.z); split2(v[6], v[7], h.w, l.w);
}
// (a, b) -> hi, mid, lo planes: a = hi + mid + lo exactly (three 8-bit pieces of the 24-bit significand)
__device__ __forceinline__ void split3(float a, float b, uint32_t& h, uint32_t& m, uint32_t& l) {
  h = bf2_pack(a, b);
  const float ra = a - __uint_as_float(h << 16), rb = b - __uint_as_float(h & 0xFFFF0000u);
  m = bf2_pack(ra, rb);
  l = bf2_pack(ra - __uint_as_float(m << 16), rb - __uint_as_float(m & 0xFFFF0000u));
}
// 8 values -> NP planes, stored as 16-byte chunks at s0 + off + p * kPart (the planes of a part are consecutive)
template <int NP>
__device__ __forceinline__ void split_store8(const float (&v)[8], uint32_t s0, uint32_t off) {
  if constexpr (NP == 2) {
    uint4 h, l;
    split8(v, h, l);
    tc::sts128(s0 + off, h);
    tc::sts128(s0 + off + kPart, l);
  } else {
    uint4 h, m, l;
    split3(v[0], v[1], h.x, m.x, l.x); split3(v[2], v[3], h.y, m.y, l.y);
    split3(v[4], v[5], h.z, m.z, l.z); split3(v[6], v[7], h.w, m.w, l.w);
    tc::sts128(s0 + off, h);
    tc::sts128(s0 + off + kPart, m);
    tc::sts128(s0 + off + 2 * kPart, l);
  }
}

// fp32 [C_out][C_in] -> three zero-padded bf16 planes [3][Rp][Kp] (hi, mid, lo; the backward kernels read the first
// two); perm_d >= 0: layer-1 column order [feats(perm_d) | xyz(xyz_cols)] (xyz_cols = 3, or 0 for a feature-only input)
// dup_cols: image columns 64..127 repeat columns 0..63 (64-wide inputs: the dgrad's TMEM lanes 64..127 then carry the same
// input channels, two epilogue threads per channel - MaskStatsW6; the forward only reads columns 0..63)
struct ConvW6 { const float* W; __nv_bfloat16* planes; int cout, cin, Rp, Kp, perm_d, xyz_cols, dup_cols; };
__global__ void convert_weights6_kernel(ConvW6 a, ConvW6 b, ConvW6 c) {
  const ConvW6* L[3] = {&a, &b, &c};
  const int n0 = a.Rp * a.Kp / 8, n1 = b.Rp * b.Kp / 8, n2 = c.Rp * c.Kp / 8;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n0 + n1 + n2; e += gridDim.x * blockDim.x) {
    const int l = e < n0 ? 0 : (e < n0 + n1 ? 1 : 2);
    const ConvW6& w = *L[l];
    const int ee = e - (l == 0 ? 0 : (l == 1 ? n0 : n0 + n1));
    const int kp8 = w.Kp >> 3, r = ee / kp8, k0 = (ee - r * kp8) * 8;
    float v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int k = (w.dup_cols && k0 + u < 128) ? ((k0 + u) & 63) : k0 + u;
      int src = k;
      if (w.perm_d >= 0) src = k < w.perm_d ? k + w.xyz_cols : (k < w.perm_d + w.xyz_cols ? k - w.perm_d : w.cin);
      v[u] = (r < w.cout && src < w.cin) ? __ldg(w.W + (size_t)r * w.cin + src) : 0.f;
    }
    uint4 h, m, lo;
    split3(v[0], v[1], h.x, m.x, lo.x); split3(v[2], v[3], h.y, m.y, lo.y);
    split3(v[4], v[5], h.z, m.z, lo.z); split3(v[6], v[7], h.w, m.w, lo.w);
    const size_t pl = (size_t)w.Rp * w.Kp / 8;
    reinterpret_cast<uint4*>(w.planes)[ee] = h;
    reinterpret_cast<uint4*>(w.planes)[pl + ee] = m;
    reinterpret_cast<uint4*>(w.planes)[2 * pl + ee] = lo;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// fp32 activations in HBM: tile-blocked and POINT-QUAD INTERLEAVED - element (tile, channel c, point p) of an
// activation with C channels sits at float index
//        ((tile * 32 + (p >> 2)) * C + c) * 4 + (p & 3)
// i.e. inside a 128-point tile the 16-byte quads of four consecutive points are stored channel after channel.  Every
// access pattern of these kernels is then fully coalesced with NO shared-memory staging: an epilogue thread (= one
// channel, TMEM lane) holding 32 points writes eight float4, and for each of them the 32 lanes of the warp (32
// consecutive channels) cover 512 contiguous bytes; a producer warp (lane = channel) reads the same way.  (With the
// plain [tile][C][128] layout each warp-wide 16-byte store touched 32 different 128-byte lines: measured 4.3 us per
// 128-point tile in the forward kernels, the stores alone ~2 us.)
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ size_t act_off(int tile, int C, int c, int quad) { return (((size_t)tile * 32 + quad) * C + c) * 4; }
// The LAST layer's pre-activations y3 are stored as bf16 in the same interleaving with 8-point octets (16 bytes):
// element (tile, c, p) at bf16 index ((tile * 16 + (p >> 3)) * C + c) * 8 + (p & 7).  y3 never feeds a forward GEMM and
// takes no ReLU-mask decision (the routing is the saved arg-max slot): its only reader is the dense BatchNorm-backward
// term p * y3 + q of dy3, where an independent 2^-9 relative rounding per element averages out over the contraction
// (measured: weight gradients unchanged at the 1e-5 level, tests/test_sa_gpu.py) - half the bytes of the largest tensor.
__device__ __forceinline__ size_t act16_off(int tile, int C, int c, int octet) { return (((size_t)tile * 16 + octet) * C + c) * 8; }

// ---------------------------------------------------------------------------------------------------------------
// Channel-major producers.  One UNIT of the 256 producer threads covers R = (PTS == 128 ? 16 : 32) * kUR channel
// rows x PTS points: thread g owns row g % R and the kUR 8-point chunks g / R + i * (256 / R), i < kUR (lanes of a warp
// = consecutive channels: coalesced global loads, conflict-free swizzled shared-memory stores).
//   PTS = 128 (forward / dgrad operand: 64-channel x 128-point part)     PTS = 64 (weight-gradient operand: 128 x 64)
// load():  raw fp32 global loads only.   store(): fp32 transform -> split into NP bf16 planes -> swizzled 16-byte stores.
// `m0` = first point row of the unit, `crow` = first channel, `lrow` = image row of `crow`, `lrows` = image rows.
// ---------------------------------------------------------------------------------------------------------------
template <int PTS, int UR>
struct CM {
  static constexpr int kR = (PTS == 128 ? 16 : 32) * UR;      // rows per unit
  static constexpr int kCstep = kProdThreads / kR;            // chunk stride between a thread's items
};
__device__ __forceinline__ float4 ldg4_or0(const float* p, bool ok) {
  return ok ? __ldg(reinterpret_cast<const float4*>(p)) : make_float4(0.f, 0.f, 0.f, 0.f);
}

// ---- asynchronous raw staging (LDGSTS) ---------------------------------------------------------------------------
// The register pipeline keeps ONE unit of global loads in flight per producer thread while the previous one is
// transformed; measured, the producers then spend most of their time on the long scoreboard (2.4 us per 32 KB unit at
// 17 warps per SM).  The channel-major producers can instead copy their raw fp32 operands global -> shared with
// cp.async into a per-thread ring of kRawDepth units ([unit][item][thread] x 16 bytes: every thread reads back exactly
// what it copied, so no barrier is involved - cp.async.wait_group orders a thread's own copies), which keeps
// kRawDepth - 1 units (64 KB per SM at depth 3) in flight without holding registers.
constexpr int kRawDepth = 3;
constexpr uint32_t kRawItemBytes = kProdThreads * 16u;        // one 16-byte item of every producer thread = 4 KB
__device__ __forceinline__ void cp_async16(uint32_t saddr, const void* g, bool ok) {
  const int sz = ok ? 16 : 0;                                  // src-size 0: the 16 bytes are zero-filled
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(saddr), "l"(g), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async4(uint32_t saddr, const void* g, bool ok) {
  const int sz = ok ? 4 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(saddr), "l"(g), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ float4 lds128f(uint32_t saddr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t lds32(uint32_t saddr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(saddr) : "memory");
  return v;
}
// slot of (item, thread g) inside a raw unit
__device__ __forceinline__ uint32_t raw_slot(uint32_t unit_base, int item, int g) { return unit_base + (uint32_t)(item * kProdThreads + g) * 16u; }

// kRawDepth-deep pipeline over W units: issue(w, slot) enqueues the async copies of unit w, consume(w, slot) runs when
// they have landed
template <int kRawDepth = 3, class IssueF, class ConsumeF>
__device__ __forceinline__ void async_pipeline(int W, uint32_t raw0, uint32_t kRawUnit, IssueF issue, ConsumeF consume) {
  if (W <= 0) return;
#pragma unroll
  for (int w = 0; w < kRawDepth - 1; ++w) {
    if (w < W) issue(w, raw0 + (uint32_t)w * kRawUnit);
    cp_async_commit();
  }
  int ri = kRawDepth - 1, rc = 0;                              // ring slots of the next issue / the next consume
  for (int w = 0; w < W; ++w) {
    if (w + kRawDepth - 1 < W) issue(w + kRawDepth - 1, raw0 + (uint32_t)ri * kRawUnit);
    cp_async_commit();                                          // one group per iteration, empty or not
    cp_async_wait<kRawDepth - 1>();                             // all but the newest kRawDepth-1 groups: unit w has landed
    consume(w, raw0 + (uint32_t)rc * kRawUnit);
    if (++ri == kRawDepth) ri = 0;
    if (++rc == kRawDepth) rc = 0;
  }
}

// relu(scale * y + shift) of the previous layer's pre-activations
struct BnRelu6 {
  static constexpr bool kChMajor = true;
  static constexpr int kUR = 4;
  struct Raw { float4 a[kUR][2]; };
  const float* __restrict__ y;        // fp32 activation (layout above)
  const float* __restrict__ scale;
  const float* __restrict__ shift;
  int M, C;
  const float* cs;
  BnFin fin;                          // fin.sums != nullptr: train-mode forward, statistics finalised here
  __host__ __device__ __forceinline__ int nconst() const { return 2 * C; }
  __host__ __device__ __forceinline__ int nchunks() const { return C / 64; }
  __host__ __device__ __forceinline__ int chunk_k(int) const { return 64; }
  __device__ __forceinline__ void init(float* csm, int tid, int nthr) {
    if (fin.sums) {
      const bool w = first_block();
      for (int c = tid; c < C; c += nthr) fin.eval(c, C, w, csm[c], csm[C + c]);
    } else {
      for (int c = tid; c < C; c += nthr) { csm[c] = scale[c]; csm[C + c] = shift[c]; }
    }
    cs = csm;
  }
  template <int PTS>
  __device__ __forceinline__ void load(int g, int m0, int crow, Raw& r) const {
    using G = CM<PTS, kUR>;
    const int row = g % G::kR, c = crow + row, ch0 = g / G::kR;
    const bool rok = c < C;
#pragma unroll
    for (int i = 0; i < kUR; ++i) {
      const int m = m0 + (ch0 + i * G::kCstep) * 8;
      const float* src = y + act_off(m >> 7, C, rok ? c : 0, (m & 127) >> 2);
      const bool ok = rok && m < M;
      r.a[i][0] = ldg4_or0(src, ok);
      r.a[i][1] = ldg4_or0(src + (size_t)C * 4, ok);
    }
  }
  static constexpr bool kAsync = true;
  template <int PTS>
  __device__ __forceinline__ void load_async(int g, int m0, int crow, uint32_t ub) const {
    using G = CM<PTS, kUR>;
    const int row = g % G::kR, c = crow + row, ch0 = g / G::kR;
    const bool rok = c < C;
#pragma unroll
    for (int i = 0; i < kUR; ++i) {
      const int m = m0 + (ch0 + i * G::kCstep) * 8;
      const float* src = y + act_off(m >> 7, C, rok ? c : 0, (m & 127) >> 2);
      const bool ok = rok && m < M;
      cp_async16(raw_slot(ub, 2 * i, g), src, ok);
      cp_async16(raw_slot(ub, 2 * i + 1, g), src + (size_t)C * 4, ok);
    }
  }
  static constexpr int kRawItems = 8;
  static constexpr int kRawItems64 = kRawItems;
  template <int PTS>
  __device__ __forceinline__ void fetch(int g, int, int, uint32_t ub, Raw& r) const {
#pragma unroll
    for (int i = 0; i < kUR; ++i) { r.a[i][0] = lds128f(raw_slot(ub, 2 * i, g)); r.a[i][1] = lds128f(raw_slot(ub, 2 * i + 1, g)); }
  }
  template <int PTS, int NP>
  __device__ __forceinline__ void store(int g, int, int crow, int lrow, int lrows, const Raw& r, uint32_t s0) const {
    using G = CM<PTS, kUR>;
    const int row = g % G::kR, c = crow + row, ch0 = g / G::kR;
    if (c >= C) return;                           // rows beyond C stay zero from the one-time clear
    const float sc = cs[c], sh = cs[C + c];
#pragma unroll
    for (int i = 0; i < kUR; ++i) {
      const float x[8] = {r.a[i][0].x, r.a[i][0].y, r.a[i][0].z, r.a[i][0].w, r.a[i][1].x, r.a[i][1].y, r.a[i][1].z, r.a[i][1].w};
      float v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = fmaxf(fmaf(x[u], sc, sh), 0.f);
      split_store8<NP>(v, s0, cm_off(lrows, lrow + row, ch0 + i * G::kCstep));
    }
  }
};

// dy^T = a*dz^T + p*y^T + q (BatchNorm backward folded into per-channel constants), dense dz
struct Dy6 {
  static constexpr bool kChMajor = true;
  static constexpr int kUR = 2;
  struct Raw { float4 d[kUR][2], y[kUR][2]; };
  const float* __restrict__ dz;
  const float* __restrict__ y;
  const float* __restrict__ a;
  const float* __restrict__ p;
  const float* __restrict__ q;
  int M, C;
  const float* cs;
  BnBwdFin fin;
  __host__ __device__ __forceinline__ int nconst() const { return 3 * C; }
  __device__ __forceinline__ void init(float* csm, int tid, int nthr) {
    if (fin.sums) {
      const bool w = first_block();
      for (int c = tid; c < C; c += nthr) fin.eval(c, C, w, csm[c], csm[C + c], csm[2 * C + c]);
    } else {
      for (int c = tid; c < C; c += nthr) { csm[c] = a[c]; csm[C + c] = p[c]; csm[2 * C + c] = q[c]; }
    }
    cs = csm;
  }
  template <int PTS>
  __device__ __forceinline__ void load(int g, int m0, int crow, Raw& r) const {
    using G = CM<PTS, kUR>;
    const int row = g % G::kR, c = crow + row, ch0 = g / G::kR;
    const bool rok = c < C;
#pragma unroll
    for (int i = 0; i < kUR; ++i) {
      const int m = m0 + (ch0 + i * G::kCstep) * 8;
      const size_t off = act_off(m >> 7, C, rok ? c : 0, (m & 127) >> 2);
      const bool ok = rok && m < M;
      r.d[i][0] = ldg4_or0(dz + off, ok); r.d[i][1] = ldg4_or0(dz + off + (size_t)C * 4, ok);
      r.y[i][0] = ldg4_or0(y + off, ok);  r.y[i][1] = ldg4_or0(y + off + (size_t)C * 4, ok);
    }
  }
  static constexpr bool kAsync = true;
  template <int PTS>
  __device__ __forceinline__ void load_async(int g, int m0, int crow, uint32_t ub) const {
    using G = CM<PTS, kUR>;
    const int row = g % G::kR, c = crow + row, ch0 = g / G::kR;
    const bool rok = c < C;
#pragma unroll
    for (int i = 0; i < kUR; ++i) {
      const int m = m0 + (ch0 + i * G::kCstep) * 8;
      const size_t off = act_off(m >> 7, C, rok ? c : 0, (m & 127) >> 2);
      const bool ok = rok && m < M;
      cp_async16(raw_slot(ub, 4 * i, g), dz + off, ok);     cp_async16(raw_slot(ub, 4 * i + 1, g), dz + off + (size_t)C * 4, ok);
      cp_async16(raw_slot(ub, 4 * i + 2, g), y + off, ok);  cp_async16(raw_slot(ub, 4 * i + 3, g), y + off + (size_t)C * 4, ok);
    }
  }
  static constexpr int kRawItems = 8;
  template <int PTS>
  __device__ __forceinline__ void fetch(int g, int, int, uint32_t ub, Raw& r) const {
#pragma unroll
    for (int i = 0; i < kUR; ++i) {
      r.d[i][0] = lds128f(raw_slot(ub, 4 * i, g));     r.d[i][1] = lds128f(raw_slot(ub, 4 * i + 1, g));
      r.y[i][0] = lds128f(raw_slot(ub, 4 * i + 2, g)); r.y[i][1] = lds128f(raw_slot(ub, 4 * i + 3, g));
    }
  }
  template <int PTS, int NP>
  __device__ __forceinline__ void store(int g, int m0, int crow, int lrow, int lrows, const Raw& r, uint32_t s0) const {
    using G = CM<PTS, kUR>;
    const int row = g % G::kR, c = crow + row, ch0 = g / G::kR;
    if (c >= C) return;
    const float ca0 = cs[c], cp0 = cs[C + c], cq0 = cs[2 * C + c];
#pragma unroll
    for (int i = 0; i < kUR; ++i) {
      const int chunk = ch0 + i * G::kCstep;
      const float okf = m0 + chunk * 8 < M ? 1.f : 0.f;     // points >= M contribute 0 to dW
      const float ca = ca0 * okf, cp = cp0 * okf, cq = cq0 * okf;
      const float d[8] = {r.d[i][0].x, r.d[i][0].y, r.d[i][0].z, r.d[i][0].w, r.d[i][1].x, r.d[i][1].y, r.d[i][1].z, r.d[i][1].w};
      const float yy[8] = {r.y[i][0].x, r.y[i][0].y, r.y[i][0].z, r.y[i][0].w, r.y[i][1].x, r.y[i][1].y, r.y[i][1].z, r.y[i][1].w};
      float v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = fmaf(ca, d[u], fmaf(cp, yy[u], cq));
      split_store8<NP>(v, s0, cm_off(lrows, lrow + row, chunk));
    }
  }
};

// last layer: dy3 = p*y3 + q + a * (max-pool routing of gm[G,C] to the saved arg slot), groups of K == 32 points
struct DyLast6 {
  static constexpr bool kChMajor = true;
  static constexpr int kUR = 4;
  struct Raw { uint4 y[kUR]; float gv[kUR]; int sl[kUR]; };      // y: 8 bf16 per chunk
  const float* __restrict__ gm;       // [G,C]
  const uint8_t* __restrict__ slot;   // [G,C]
  const __nv_bfloat16* __restrict__ y;   // bf16 (act16_off)
  const float* __restrict__ a;
  const float* __restrict__ p;
  const float* __restrict__ q;
  int M, C;
  const float* cs;
  BnBwdFin fin;
  __host__ __device__ __forceinline__ int nconst() const { return 3 * C; }
  __device__ __forceinline__ void init(float* csm, int tid, int nthr) {
    if (fin.sums) {
      const bool w = first_block();
      for (int c = tid; c < C; c += nthr) fin.eval(c, C, w, csm[c], csm[C + c], csm[2 * C + c]);
    } else {
      for (int c = tid; c < C; c += nthr) { csm[c] = a[c]; csm[C + c] = p[c]; csm[2 * C + c] = q[c]; }
    }
    cs = csm;
  }
  template <int PTS>
  __device__ __forceinline__ void load(int g, int m0, int crow, Raw& r) const {
    using G = CM<PTS, kUR>;
    const int row = g % G::kR, c = crow + row, ch0 = g / G::kR;
    const bool rok = c < C;
#pragma unroll
    for (int i = 0; i < kUR; ++i) {
      const int m = m0 + (ch0 + i * G::kCstep) * 8;
      const bool ok = rok && m < M;
      const __nv_bfloat16* sy = y + act16_off(m >> 7, C, rok ? c : 0, (m & 127) >> 3);
      const size_t go = (size_t)((ok ? m : 0) >> 5) * C + (rok ? c : 0);
      r.y[i] = ok ? __ldg(reinterpret_cast<const uint4*>(sy)) : make_uint4(0u, 0u, 0u, 0u);
      r.gv[i] = ok ? __ldg(gm + go) : 0.f;
      r.sl[i] = ok ? (int)__ldg(slot + go) : -1;
    }
  }
  // raw unit: items 0..3 = y (one 16-byte octet per chunk), item 4 = [gm x 4], item 5 = [slot word x 4]
  static constexpr bool kAsync = true;
  template <int PTS>
  __device__ __forceinline__ void load_async(int g, int m0, int crow, uint32_t ub) const {
    using G = CM<PTS, kUR>;
    const int row = g % G::kR, c = crow + row, ch0 = g / G::kR;
    const bool rok = c < C;
#pragma unroll
    for (int i = 0; i < kUR; ++i) {
      const int m = m0 + (ch0 + i * G::kCstep) * 8;
      const bool ok = rok && m < M;
      const __nv_bfloat16* sy = y + act16_off(m >> 7, C, rok ? c : 0, (m & 127) >> 3);
      const size_t go = (size_t)((ok ? m : 0) >> 5) * C + (rok ? c : 0);
      cp_async16(raw_slot(ub, i, g), sy, ok);
      cp_async4(raw_slot(ub, 4, g) + 4u * i, gm + go, ok);
      cp_async4(raw_slot(ub, 5, g) + 4u * i, slot + (go & ~(size_t)3), ok);    // the aligned word that holds the byte
    }
  }
  static constexpr int kRawItems = 6;
  template <int PTS>
  __device__ __forceinline__ void fetch(int g, int m0, int crow, uint32_t ub, Raw& r) const {
    using G = CM<PTS, kUR>;
    const int row = g % G::kR, c = crow + row, ch0 = g / G::kR;
    const bool rok = c < C;
#pragma unroll
    for (int i = 0; i < kUR; ++i) {
      const int m = m0 + (ch0 + i * G::kCstep) * 8;
      const bool ok = rok && m < M;
      const float4 t = lds128f(raw_slot(ub, i, g));
      r.y[i] = make_uint4(__float_as_uint(t.x), __float_as_uint(t.y), __float_as_uint(t.z), __float_as_uint(t.w));
      r.gv[i] = __uint_as_float(lds32(raw_slot(ub, 4, g) + 4u * i));
      const uint32_t w = lds32(raw_slot(ub, 5, g) + 4u * i);
      r.sl[i] = ok ? (int)((w >> (8 * (c & 3))) & 255u) : -1;                  // C % 4 == 0: byte index = c & 3
    }
  }
  template <int PTS, int NP>
  __device__ __forceinline__ void store(int g, int m0, int crow, int lrow, int lrows, const Raw& r, uint32_t s0) const {
    using G = CM<PTS, kUR>;
    const int row = g % G::kR, c = crow + row, ch0 = g / G::kR;
    if (c >= C) return;
    const float ca0 = cs[c], cp0 = cs[C + c], cq0 = cs[2 * C + c];
#pragma unroll
    for (int i = 0; i < kUR; ++i) {
      const int chunk = ch0 + i * G::kCstep, m = m0 + chunk * 8;
      const float okf = m < M ? 1.f : 0.f;
      const float add = ca0 * r.gv[i], cp = cp0 * okf, cq = cq0 * okf;   // gv is 0 for points >= M
      const int sl = r.sl[i] - (m & 31);          // slot relative to this 8-point chunk
      const uint32_t yw[4] = {r.y[i].x, r.y[i].y, r.y[i].z, r.y[i].w};
      float yy[8];
#pragma unroll
      for (int u = 0; u < 4; ++u) { yy[2 * u] = __uint_as_float(yw[u] << 16); yy[2 * u + 1] = __uint_as_float(yw[u] & 0xFFFF0000u); }
      float v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = fmaf(cp, yy[u], cq) + (u == sl ? add : 0.f);
      split_store8<NP>(v, s0, cm_off(lrows, lrow + row, chunk));
    }
  }
};

// layer-1 input, point-major, streamed in 64-channel blocks: block kb < D/64 holds feats[:, 64kb .. +63], block D/64
// holds [xyz - centroid (3) | zeros (13)] (channel order [feats | xyz], as the layer-1 weight images).  D % 64 == 0,
// D == 0 (SA1) is the xyz block alone.  Thread g owns the 8-channel unit g & 7 of rows (g >> 3) + 32 i, i < PTS/32.
struct GatherFeat6 {
  static constexpr bool kChMajor = false;
  static constexpr bool kAsync = true;
  static constexpr int kRawItems = 8;
  static constexpr int kRawItems64 = 4;   // items of a 64-point unit (wgrad Q operand): two per 32 rows
  static constexpr int kUR = 4;       // unused (point-major): one unit per 64-channel block
  struct Raw { float4 a[4], b[4]; };
  GatherBase gb;
  const float* __restrict__ feats;
  int D;
  int no_xyz;                         // 1: feature-only input (pointwise MLP stacks): the xyz block does not exist
  __host__ __device__ __forceinline__ int nconst() const { return 0; }
  __host__ __device__ __forceinline__ int nchunks() const { return D / 64 + (no_xyz ? 0 : 1); }
  __host__ __device__ __forceinline__ int chunk_k(int kb) const { return kb < D / 64 ? 64 : 16; }
  __device__ __forceinline__ void init(float*, int, int) {}
  template <int PTS>
  __device__ __forceinline__ void load(int g, int m0, int kb, Raw& r) const {
    if (kb < D / 64) {
      const int j = g & 7;
#pragma unroll
      for (int i = 0; i < PTS / 32; ++i) {
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
      if (g < PTS && m0 + g < gb.M) gb.load_xyz_raw(m0 + g, gb.point_of(m0 + g), v, c);
      r.a[0] = make_float4(v[0], v[1], v[2], 0.f);
      r.b[0] = make_float4(c[0], c[1], c[2], 0.f);
    }
  }
  // cp.async path: the neighbour indices are read synchronously (L2-resident int32 rows, shared by the 8 threads of a
  // point), the gathered 512-byte feature rows / xyz triples are copied asynchronously.
  // feature block: items 2i, 2i+1 = the two 16-byte halves of row i's 8-channel unit; xyz block: item 0 = point, item 1 = centroid
  template <int PTS>
  __device__ __forceinline__ void load_async(int g, int m0, int kb, uint32_t ub) const {
    if (kb < D / 64) {
      const int j = g & 7;
#pragma unroll
      for (int i = 0; i < PTS / 32; ++i) {
        const int row = m0 + (g >> 3) + 32 * i;
        const bool ok = row < gb.M;
        const float* src = feats + (size_t)(ok ? gb.point_of(row) : 0) * D + kb * 64 + j * 8;
        cp_async16(raw_slot(ub, 2 * i, g), src, ok);
        cp_async16(raw_slot(ub, 2 * i + 1, g), src + 4, ok);
      }
    } else if (g < PTS) {
      const int row = m0 + g;
      const bool ok = row < gb.M;
      const int pt = ok ? gb.point_of(row) : 0;
      const float* px = gb.xyz + (size_t)pt * 3;
      const float* pc = gb.group_all ? px : gb.new_xyz + (size_t)((ok ? row : 0) >> 5) * 3;
#pragma unroll
      for (int u = 0; u < 3; ++u) {
        cp_async4(raw_slot(ub, 0, g) + 4u * u, px + u, ok);
        cp_async4(raw_slot(ub, 1, g) + 4u * u, pc + u, ok && !gb.group_all);
      }
    }
  }
  template <int PTS>
  __device__ __forceinline__ void fetch(int g, int, int kb, uint32_t ub, Raw& r) const {
    if (kb < D / 64) {
#pragma unroll
      for (int i = 0; i < PTS / 32; ++i) { r.a[i] = lds128f(raw_slot(ub, 2 * i, g)); r.b[i] = lds128f(raw_slot(ub, 2 * i + 1, g)); }
    } else if (g < PTS) {
      r.a[0] = lds128f(raw_slot(ub, 0, g));      // .w holds stale bytes: only x, y, z are used
      r.b[0] = lds128f(raw_slot(ub, 1, g));
      if (gb.group_all) r.b[0] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  template <int PTS, int NP>
  __device__ __forceinline__ void store(int g, int kb, const Raw& r, uint32_t s0) const {
    if (kb < D / 64) {
      const int j = g & 7;
#pragma unroll
      for (int i = 0; i < PTS / 32; ++i) {
        const float v[8] = {r.a[i].x, r.a[i].y, r.a[i].z, r.a[i].w, r.b[i].x, r.b[i].y, r.b[i].z, r.b[i].w};
        split_store8<NP>(v, s0, tc::sw128_off((g >> 3) + 32 * i, j * 8));
      }
    } else if (g < PTS) {
      const float v[8] = {gb.centred(r.a[0].x, r.b[0].x), gb.centred(r.a[0].y, r.b[0].y), gb.centred(r.a[0].z, r.b[0].z),
                          0.f, 0.f, 0.f, 0.f, 0.f};
      const float z[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      split_store8<NP>(v, s0, tc::sw128_off(g, 0));
      split_store8<NP>(z, s0, tc::sw128_off(g, 8));
    }
  }
};

// ---------------------------------------------------------------------------------------------------------------
// Epilogues: thread = one channel (TMEM lane), v = 32 consecutive points (block j of the tile), fp32 outputs written
// straight to the quad-interleaved activation: eight float4 per thread, each warp-wide store 512 contiguous bytes.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void store32_f32(float* y, int tile, int C, int c, int j, const float (&v)[32]) {
  float* dst = y + act_off(tile, C, c, j * 8);
#pragma unroll
  for (int q = 0; q < 8; ++q)
    *reinterpret_cast<float4*>(dst + (size_t)q * C * 4) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
}

struct StoreStats6 {
  static constexpr bool kPre = false;
  static constexpr bool kHalf = false;
  float* __restrict__ y;           // fp32 activation (quad-interleaved, see act_off)
  double* __restrict__ sums;       // [kRedCopies][2,C] or nullptr (eval)
  int C;
  int c;
  float s0, s1;
  __host__ __device__ __forceinline__ int nconst() const { return 0; }
  __device__ __forceinline__ void init(float*, int ch) { c = ch; s0 = s1 = 0.f; }
  __device__ __forceinline__ void block(float (&v)[32], int tile, int j, bool valid) {
    if (c >= C || !valid) return;
    store32_f32(y, tile, C, c, j, v);
    sum_sumsq32(v, s0, s1);
  }
  __device__ __forceinline__ void finish() {
    if (!sums || c >= C) return;
    double* dst = sums + (size_t)(blockIdx.x % kRedCopies) * 2 * C;
    atomicAdd(dst + c, (double)s0);
    atomicAdd(dst + C + c, (double)s1);
  }
};

struct Group6 {   // last layer, K == 32: the 32 columns of a block are one group
  static constexpr bool kPre = false;
  static constexpr bool kHalf = false;
  __nv_bfloat16* __restrict__ y;   // bf16 activation (act16_off), or nullptr (eval / inference)
  double* __restrict__ sums;
  float* __restrict__ ymax;        // [G,C]
  float* __restrict__ ymin;
  uint8_t* __restrict__ amax;
  uint8_t* __restrict__ amin;
  int C;
  const float* __restrict__ gamma; // the affine's sign is gamma's sign: each channel needs only one of (max, min)
  int c;
  float s0, s1;
  bool want_max;
  __host__ __device__ __forceinline__ int nconst() const { return 0; }
  __device__ __forceinline__ void init(float*, int ch) { c = ch; s0 = s1 = 0.f; want_max = c < C ? !signbit(gamma[c]) : true; }
  __device__ __forceinline__ void block(float (&v)[32], int tile, int j, bool valid) {
    if (c >= C || !valid) return;
    if (y) {
      __nv_bfloat16* dst = y + act16_off(tile, C, c, j * 4);
#pragma unroll
      for (int o = 0; o < 4; ++o) {
        const float t[8] = {v[8 * o], v[8 * o + 1], v[8 * o + 2], v[8 * o + 3], v[8 * o + 4], v[8 * o + 5], v[8 * o + 6], v[8 * o + 7]};
        *reinterpret_cast<uint4*>(dst + (size_t)o * C * 8) = tc::pack8_bf16(t);
      }
    }
    sum_sumsq32(v, s0, s1);
    float ext;
    int arg;
    const size_t o = (size_t)(tile * 4 + j) * C + c;
    if (want_max) { argext32<true>(v, ext, arg); ymax[o] = ext; amax[o] = (uint8_t)arg; }
    else { argext32<false>(v, ext, arg); ymin[o] = ext; amin[o] = (uint8_t)arg; }
  }
  __device__ __forceinline__ void finish() {
    if (!sums || c >= C) return;
    double* dst = sums + (size_t)(blockIdx.x % kRedCopies) * 2 * C;
    atomicAdd(dst + c, (double)s0);
    atomicAdd(dst + C + c, (double)s1);
  }
};

// dz_prev^T = dx^T * [z_prev > 0]; sums of dz_prev and dz_prev * xhat_prev per channel
struct MaskStats6 {
  static constexpr bool kPre = false;
  static constexpr bool kHalf = false;
  const float* __restrict__ yprev;   // tile-blocked fp32
  const float* __restrict__ scale;
  const float* __restrict__ shift;
  const float* __restrict__ mean;
  const float* __restrict__ invstd;
  float* __restrict__ dz;            // tile-blocked fp32
  double* __restrict__ sums;
  int C;
  int c;
  float s0, s1, sc, sh, is, nmi;
  __host__ __device__ __forceinline__ int nconst() const { return 0; }
  __device__ __forceinline__ void init(float*, int ch) {
    c = ch; s0 = s1 = 0.f;
    const bool ok = c < C;
    sc = ok ? scale[c] : 0.f; sh = ok ? shift[c] : 0.f; is = ok ? invstd[c] : 0.f; nmi = ok ? -mean[c] * is : 0.f;
  }
  __device__ __forceinline__ void block(float (&v)[32], int tile, int j, bool valid) {
    if (c >= C || !valid) return;
    const size_t o = act_off(tile, C, c, j * 8), qs = (size_t)C * 4;   // quad q of this block at o + q * qs
    float a[4] = {0.f, 0.f, 0.f, 0.f}, b[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int h = 0; h < 2; ++h) {                 // two halves of 16 points: bounds the registers held by the y loads
      float4 yv[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) yv[q] = __ldg(reinterpret_cast<const float4*>(yprev + o + (size_t)(4 * h + q) * qs));
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float yy[4] = {yv[q].x, yv[q].y, yv[q].z, yv[q].w};
        float d[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const bool on = fmaf(yy[u], sc, sh) > 0.f;
          d[u] = on ? v[16 * h + 4 * q + u] : 0.f;
          a[u] += d[u];
          b[u] = fmaf(d[u], fmaf(yy[u], is, nmi), b[u]);
        }
        *reinterpret_cast<float4*>(dz + o + (size_t)(4 * h + q) * qs) = make_float4(d[0], d[1], d[2], d[3]);
      }
    }
    s0 += (a[0] + a[1]) + (a[2] + a[3]);
    s1 += (b[0] + b[1]) + (b[2] + b[3]);
  }
  __device__ __forceinline__ void finish() {
    if (c >= C) return;
    double* dst = sums + (size_t)(blockIdx.x % kRedCopies) * 2 * C;
    atomicAdd(dst + c, (double)s0);
    atomicAdd(dst + C + c, (double)s1);
  }
};


// SA1 (no input features): layer 1 has no data gradient, only dW1 [C1 x 3] = dy1^T x0 with dy1 = a1 dz1 + p1 y1 + q1 and
// y1 = W1 x0, i.e.  dW1 = a1 (dz1^T x0) + p1 W1 (x0^T x0) + q1 (sum x0)^T.  The layer-2 dgrad epilogue holds dz1 in
// registers (one channel per thread), so it accumulates dz1^T x0 itself in fp32: dz1 is never written, the layer-1
// weight-gradient kernel (a full pass over dz1 and y1 for a 64 x 3 result) does not run.  pre() parks the 32 point
// offsets of the warp's blocks in shared memory before the wait for the accumulator; one warp per block also
// accumulates x0^T x0 and sum x0.  dw1_finalize6_kernel applies (a1, p1, q1) once the batch sums are complete.
struct MaskStatsW6 {
  static constexpr bool kPre = true;
  static constexpr bool kHalf = true;   // C == 64 only: lanes 64..127 repeat the channels (ConvW6::dup_cols), 16 points per thread
  int half;
  const float* __restrict__ W1;      // layer-1 weights fp32 [C][3]: y1 = W1 x0 is recomputed (3 FMAs) instead of read back -
                                     // the same value as the forward's 24-bit-operand MMA to an ulp; the conv bias is not
                                     // part of y1 in train mode (BatchNorm cancels it)
  const float* __restrict__ scale;
  const float* __restrict__ shift;
  const float* __restrict__ mean;
  const float* __restrict__ invstd;
  double* __restrict__ sums;
  int C;
  GatherBase gb;
  float* __restrict__ acc;           // [kRedCopies][C][4]: dz1^T x0 in columns 0..2
  float* __restrict__ g0;            // [kRedCopies][16]: xx xy xz yy yz zz sx sy sz
  int c, eq;
  float s0, s1, sc, sh, is, nmi, w0, w1, w2;
  float A0, A1, A2;
  float G[9];
  float4* xs;                        // shared memory: [warp][2 blocks][32 points] (x, y, z, 0)
  // software pipeline of the gather (index -> coordinates are dependent loads): while tile i is processed, the
  // coordinates of tile i+1 and the neighbour indices of tile i+2 are in flight
  float xv[2][3], cv[2][3];          // raw point / centroid of this lane's 2 points of the NEXT tile to be parked
  int pt_n[2];                       // xyz row of this lane's 2 points of the tile after that (-1: beyond M)
  __host__ __device__ __forceinline__ int nconst() const { return 8 * 2 * 32 * 4; }
  __device__ __forceinline__ void init(float* csm, int ch) {
    xs = reinterpret_cast<float4*>(csm) + (threadIdx.x >> 5) * 64;   // csm is 16-byte aligned
    c = ch & 63; half = (ch >> 6) & 1;
    eq = (threadIdx.x >> 5) & 3; s0 = s1 = 0.f; A0 = A1 = A2 = 0.f;
#pragma unroll
    for (int i = 0; i < 9; ++i) G[i] = 0.f;
    const bool ok = c < C;
    sc = ok ? scale[c] : 0.f; sh = ok ? shift[c] : 0.f; is = ok ? invstd[c] : 0.f; nmi = ok ? -mean[c] * is : 0.f;
    w0 = ok ? W1[c * 3] : 0.f; w1 = ok ? W1[c * 3 + 1] : 0.f; w2 = ok ? W1[c * 3 + 2] : 0.f;
  }
  __device__ __forceinline__ void load_idx(int tile, int eh, int M) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int it = 0; it < 2; ++it) {
      const int row = tile * kPts + (2 * eh + it) * 32 + lane;
      pt_n[it] = row < M ? gb.point_of(row) : -1;
    }
  }
  __device__ __forceinline__ void load_xyz(int tile, int eh) {        // uses pt_n (of this tile)
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int it = 0; it < 2; ++it) {
      const int row = tile * kPts + (2 * eh + it) * 32 + lane;
#pragma unroll
      for (int u = 0; u < 3; ++u) xv[it][u] = cv[it][u] = 0.f;
      if (pt_n[it] >= 0) gb.load_xyz_raw(row, pt_n[it], xv[it], cv[it]);
    }
  }
  // before the first tile: coordinates of tile t0 and indices of tile t1 requested
  __device__ __forceinline__ void prime(int t0, int t1, int eh, int M) {
    load_idx(t0, eh, M);
    load_xyz(t0, eh);
    load_idx(t1, eh, M);
  }
  // tile = the tile about to be processed; t1 / t2 = the next two tiles of this CTA (may lie beyond the last tile)
  __device__ __forceinline__ void pre(int tile, int t1, int t2, int eh, int M) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int it = 0; it < 2; ++it)   // the forward's x0, bit for bit; (0, 0, 0) beyond M
      xs[it * 32 + lane] = make_float4(gb.centred(xv[it][0], cv[it][0]), gb.centred(xv[it][1], cv[it][1]),
                                       gb.centred(xv[it][2], cv[it][2]), 0.f);
    load_xyz(t1, eh);                // pt_n holds t1's rows (requested one tile ago)
    load_idx(t2, eh, M);
    __syncwarp();
  }
  // v = the accumulator's columns j*32 + half*16 .. +15 (16 points of block j) of channel c
  __device__ __forceinline__ void block16(float (&v)[16], int tile, int j, bool valid) {
    const int it = j & 1;
    const float4* xp = xs + it * 32;
    if (valid) {
      float a[4] = {0.f, 0.f, 0.f, 0.f}, b[4] = {0.f, 0.f, 0.f, 0.f};
      float t0[4] = {0.f, 0.f, 0.f, 0.f}, t1[4] = {0.f, 0.f, 0.f, 0.f}, t2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float4 x0 = xp[half * 16 + i];                              // broadcast 16-byte load
        const float yy = fmaf(w2, x0.z, fmaf(w1, x0.y, w0 * x0.x));       // y1 = W1 x0
        const bool on = fmaf(yy, sc, sh) > 0.f;
        const float d = on ? v[i] : 0.f;
        a[i & 3] += d;
        b[i & 3] = fmaf(d, fmaf(yy, is, nmi), b[i & 3]);
        t0[i & 3] = fmaf(d, x0.x, t0[i & 3]); t1[i & 3] = fmaf(d, x0.y, t1[i & 3]); t2[i & 3] = fmaf(d, x0.z, t2[i & 3]);
      }
      s0 += (a[0] + a[1]) + (a[2] + a[3]);
      s1 += (b[0] + b[1]) + (b[2] + b[3]);
      A0 += (t0[0] + t0[1]) + (t0[2] + t0[3]); A1 += (t1[0] + t1[1]) + (t1[2] + t1[3]); A2 += (t2[0] + t2[1]) + (t2[2] + t2[3]);
      if (eq == 3) {   // one warp per block: x0^T x0 and sum x0 of this lane's point (zero beyond M)
        const float4 own = xp[threadIdx.x & 31];
        G[0] = fmaf(own.x, own.x, G[0]); G[1] = fmaf(own.x, own.y, G[1]); G[2] = fmaf(own.x, own.z, G[2]);
        G[3] = fmaf(own.y, own.y, G[3]); G[4] = fmaf(own.y, own.z, G[4]); G[5] = fmaf(own.z, own.z, G[5]);
        G[6] += own.x; G[7] += own.y; G[8] += own.z;
      }
    }
  }
  __device__ __forceinline__ void finish() {
    const int cp = blockIdx.x % kRedCopies;
    if (c < C) {
      double* dst = sums + (size_t)cp * 2 * C;
      atomicAdd(dst + c, (double)s0);
      atomicAdd(dst + C + c, (double)s1);
      float* ad = acc + ((size_t)cp * C + c) * 4;
      atomicAdd(ad, A0); atomicAdd(ad + 1, A1); atomicAdd(ad + 2, A2);
    }
    if (eq == 3) {
#pragma unroll
      for (int i = 0; i < 9; ++i) {
        float t = G[i];
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) t += __shfl_xor_sync(0xFFFFFFFFu, t, m);
        if ((threadIdx.x & 31) == 0) atomicAdd(g0 + cp * 16 + i, t);
      }
    }
  }
};

// dW1[c][k] (+)= a_c A[c][k] + p_c sum_j W1[c][j] G0[j][k] + q_c s[k]   (W1 fp32 [C][3]); thread (c, k = 0) also writes
// the layer's dgamma / dbeta through fin.eval
__global__ void dw1_finalize6_kernel(const float* __restrict__ acc, const float* __restrict__ g0, const float* __restrict__ W1,
                                     BnBwdFin fin, int C, float* __restrict__ dW, int accumulate) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= C * 3) return;
  const int r = e / 3, k = e - r * 3;
  float A = 0.f, G[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) G[i] = 0.f;
#pragma unroll
  for (int g = 0; g < kRedCopies; ++g) {
    A += acc[((size_t)g * C + r) * 4 + k];
#pragma unroll
    for (int i = 0; i < 9; ++i) G[i] += g0[g * 16 + i];
  }
  float ca, cp, cq;
  fin.eval(r, C, k == 0, ca, cp, cq);
  const float w0 = W1[r * 3], w1 = W1[r * 3 + 1], w2 = W1[r * 3 + 2];
  const float gk0 = k == 0 ? G[0] : (k == 1 ? G[1] : G[2]);     // G0 rows: (xx xy xz), (xy yy yz), (xz yz zz)
  const float gk1 = k == 0 ? G[1] : (k == 1 ? G[3] : G[4]);
  const float gk2 = k == 0 ? G[2] : (k == 1 ? G[4] : G[5]);
  const float s = ca * A + cp * (w0 * gk0 + w1 * gk1 + w2 * gk2) + cq * G[6 + k];
  dW[e] = accumulate ? dW[e] + s : s;
}

// ---- weight-slice loads, one plane per call (4 x 16 bytes per thread) ---------------------------------------------
// K-major slice [128 rows x 64 k] (forward A operand)
__device__ __forceinline__ void wload_k1(const __nv_bfloat16* __restrict__ W, int Kp, int row0, int k0, int g, uint4* w) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
    w[i] = __ldg(reinterpret_cast<const uint4*>(W + (size_t)(row0 + (g >> 3) + 32 * i) * Kp + k0 + (g & 7) * 8));
}
__device__ __forceinline__ void wstore_k1(uint32_t saddr, int g, const uint4* w) {
#pragma unroll
  for (int i = 0; i < 4; ++i) tc::sts128(saddr + tc::sw128_off((g >> 3) + 32 * i, (g & 7) * 8), w[i]);
}
// MN-major slice [64 k rows x 128 columns] = 2 blocks of [64 rows x 64 columns] (dgrad A operand)
__device__ __forceinline__ void wload_mn1(const __nv_bfloat16* __restrict__ W, int Kp, int row0, int col0, int g, uint4* w) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
    w[i] = __ldg(reinterpret_cast<const uint4*>(W + (size_t)(row0 + (g >> 4) + 16 * i) * Kp + col0 + (g & 15) * 8));
}
__device__ __forceinline__ void wstore_mn1(uint32_t saddr, int g, const uint4* w) {
  const int j = g & 15;
#pragma unroll
  for (int i = 0; i < 4; ++i)
    tc::sts128(saddr + (uint32_t)(j >> 3) * 8192u + tc::sw128_off((g >> 4) + 16 * i, (j & 7) * 8), w[i]);
}

// the same slices copied asynchronously into raw-ring items 0..3 of the thread (streamed weights on the cp.async path)
__device__ __forceinline__ void wload_k1_async(const __nv_bfloat16* __restrict__ W, int Kp, int row0, int k0, int g, uint32_t ub) {
#pragma unroll
  for (int i = 0; i < 4; ++i) cp_async16(raw_slot(ub, i, g), W + (size_t)(row0 + (g >> 3) + 32 * i) * Kp + k0 + (g & 7) * 8, true);
}
__device__ __forceinline__ void wload_mn1_async(const __nv_bfloat16* __restrict__ W, int Kp, int row0, int col0, int g, uint32_t ub) {
#pragma unroll
  for (int i = 0; i < 4; ++i) cp_async16(raw_slot(ub, i, g), W + (size_t)(row0 + (g >> 4) + 16 * i) * Kp + col0 + (g & 15) * 8, true);
}
__device__ __forceinline__ void wfetch(uint32_t ub, int g, uint4* w) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float4 v = lds128f(raw_slot(ub, i, g));
    w[i] = make_uint4(__float_as_uint(v.x), __float_as_uint(v.y), __float_as_uint(v.z), __float_as_uint(v.w));
  }
}

struct Barriers6 {
  uint64_t full[kMaxStages6];
  uint64_t empty[kMaxStages6];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
};

#define PCOE_V6_PROLOGUE(TCOLS)                                                                    \
  extern __shared__ uint8_t smem_raw[];                                                            \
  const uint32_t smem0 = (tc::smem_u32(smem_raw) + 1023u) & ~1023u;                                \
  uint8_t* smem_gen = smem_raw + (smem0 - tc::smem_u32(smem_raw));                                 \
  __shared__ Barriers6 bar;                                                                        \
  __shared__ uint32_t tmem_base;                                                                   \
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;                                   \
  if (warp == 0) tc::tmem_alloc<TCOLS>(&tmem_base);                                                \
  if (tid == 0) {                                                                                  \
    for (int s = 0; s < kMaxStages6; ++s) { tc::mbar_init(&bar.full[s], kProdThreads); tc::mbar_init(&bar.empty[s], 1); } \
    for (int b = 0; b < 2; ++b) { tc::mbar_init(&bar.tmem_full[b], 1); tc::mbar_init(&bar.tmem_empty[b], kEpiThreads); } \
  }

// The retained plane products of one 16-deep MMA step, small terms first.  A / B: descriptors of plane 0; the planes of
// an operand part are kPart bytes apart, i.e. +(kPart >> 4) in the descriptor's start-address field.
template <int NP>
__device__ __forceinline__ void mma_planes(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, bool acc) {
  constexpr uint64_t P = kPart >> 4;
  if constexpr (NP == 2) {
    tc::mma_bf16_warp(d, a + P, b, idesc, acc);
    tc::mma_bf16_warp(d, a, b + P, idesc, true);
    tc::mma_bf16_warp(d, a, b, idesc, true);
  } else {
    tc::mma_bf16_warp(d, a + 2 * P, b, idesc, acc);       // lo  * hi
    tc::mma_bf16_warp(d, a, b + 2 * P, idesc, true);      // hi  * lo
    tc::mma_bf16_warp(d, a + P, b + P, idesc, true);      // mid * mid
    tc::mma_bf16_warp(d, a + P, b, idesc, true);          // mid * hi
    tc::mma_bf16_warp(d, a, b + P, idesc, true);          // hi  * mid
    tc::mma_bf16_warp(d, a, b, idesc, true);              // hi  * hi
  }
}

// ---------------------------------------------------------------------------------------------------------------
// forward (NP planes): Y^T[128 ch x 128 pts] = W[128 x Cin] * X^T[Cin x 128] per (tile, channel block).
// smem: wres ? [W: nk x NP parts][ring: nst x (NP X parts)] : [ring: nst x (NP W parts | NP X parts)]
// Wp = plane 0 of the weight image, wps = elements per plane.
// ---------------------------------------------------------------------------------------------------------------
template <class Prod, class Epi, int NP, bool ASYNC = false>
__global__ void __launch_bounds__(kThreads, 1)
x3_fwd_kernel(Prod prod, const __nv_bfloat16* __restrict__ Wp, size_t wps, int Kp, Epi epi, int M, int ncb, int nst, int wres) {
  PCOE_V6_PROLOGUE(256)
  const int nk = prod.nchunks();
  constexpr uint32_t kOp = NP * kPart;                       // one operand chunk, all planes
  constexpr uint32_t kRawUnit = ASYNC ? (uint32_t)Prod::kRawItems * kRawItemBytes : 0u, kRawBytes = kRawDepth * kRawUnit;
  const uint32_t sWres = smem0, wres_bytes = wres ? (uint32_t)nk * kOp : 0u;
  const uint32_t sS = smem0 + wres_bytes, sbytes = wres ? kOp : 2u * kOp, xoff = wres ? 0u : kOp;
  const uint32_t sRaw = sS + (uint32_t)nst * sbytes;          // ASYNC: raw staging ring
  float* csm = reinterpret_cast<float*>(smem_gen + wres_bytes + (size_t)nst * sbytes + kRawBytes);
  const int cb = blockIdx.x % ncb, t0 = blockIdx.x / ncb, tstep = gridDim.x / ncb;
  const int ntiles = (M + kPts - 1) / kPts;
  const int my_items = t0 < ntiles ? (ntiles - t0 + tstep - 1) / tstep : 0;
  zero_smem(sS, (uint32_t)nst * sbytes, tid, kThreads);
  prod.init(csm, tid, kThreads);
  const int eq = warp & 3, eh = (warp >> 2) & 1;
  epi.init(csm + prod.nconst(), cb * 128 + eq * 32 + lane);
  if (wres && tid < kProdThreads) {
    for (int k = 0; k < nk; ++k)
#pragma unroll
      for (int p = 0; p < NP; ++p) {
        uint4 w[4];
        wload_k1(Wp + (size_t)p * wps, Kp, cb * 128, k * 64, tid, w);
        wstore_k1(sWres + (uint32_t)k * kOp + (uint32_t)p * kPart, tid, w);
      }
  }
  tc::fence_proxy_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = tmem_base;

  if (warp < 8) {
    int i = 0;
    for (int tile = t0; tile < ntiles; tile += tstep, ++i) {
      const int b = i & 1, u = i >> 1, m0 = tile * kPts;
      tc::mbar_wait(&bar.tmem_full[b], (uint32_t)(u & 1));
      tc::fence_after_sync();
#pragma unroll 1
      for (int j = eh * 2; j < eh * 2 + 2; ++j) {
        float v[32];
        tc::tmem_ld32(tmem + ((uint32_t)(eq * 32) << 16) + (uint32_t)(b * kPts + j * 32), v);
        epi.block(v, tile, j, m0 + j * 32 < M);
      }
      tc::fence_before_sync();
      mbar_arrive_relaxed(&bar.tmem_empty[b]);
    }
    epi.finish();
  } else if (warp < 16) {
    const int g = tid - kEpiThreads;
    constexpr int kXU = Prod::kChMajor ? 4 / Prod::kUR : 1;   // activation units per 64-channel chunk
    const int wu = wres ? 0 : NP, upc = kXU + wu;             // streamed weights: one unit per plane
    union RawU { typename Prod::Raw x; uint4 w[4]; __device__ RawU() {} };
    struct Cur { int tile, k, u; };
    auto adv = [&](Cur& c) { if (++c.u == upc) { c.u = 0; if (++c.k == nk) { c.k = 0; c.tile += tstep; } } };
    Cur cl{t0, 0, 0}, cst = cl;
    int ring_s = 0, ring_r = 0;
    auto store_unit = [&](const RawU& r) {
      const uint32_t st = sS + (uint32_t)ring_s * sbytes;
      if (cst.u == 0 && ring_r > 0) tc::mbar_wait(&bar.empty[ring_s], (uint32_t)((ring_r - 1) & 1));
      const int xu = cst.u - wu;
      if (xu < 0) wstore_k1(st + (uint32_t)cst.u * kPart, g, r.w);
      else if constexpr (Prod::kChMajor)
        prod.template store<128, NP>(g, cst.tile * kPts, cst.k * 64 + xu * 16 * Prod::kUR, xu * 16 * Prod::kUR, 64, r.x, st + xoff);
      else prod.template store<128, NP>(g, cst.k, r.x, st + xoff);
      if (cst.u == upc - 1) {
        tc::fence_proxy_async();
        mbar_arrive(&bar.full[ring_s]);
        if (++ring_s == nst) { ring_s = 0; ++ring_r; }
      }
      adv(cst);
    };
    if constexpr (ASYNC) {   // resident weights, channel-major activations: raw operands staged with cp.async
      async_pipeline(my_items * nk * upc, sRaw, kRawUnit,
          [&](int, uint32_t ub) {
            const int xu = cl.u - wu;
            if (xu < 0) wload_k1_async(Wp + (size_t)cl.u * wps, Kp, cb * 128, cl.k * 64, g, ub);
            else if constexpr (Prod::kChMajor) prod.template load_async<128>(g, cl.tile * kPts, cl.k * 64 + xu * 16 * Prod::kUR, ub);
            else prod.template load_async<128>(g, cl.tile * kPts, cl.k, ub);
            adv(cl);
          },
          [&](int, uint32_t ub) {
            RawU r;
            const int xu = cst.u - wu;
            if (xu < 0) wfetch(ub, g, r.w);
            else if constexpr (Prod::kChMajor) prod.template fetch<128>(g, cst.tile * kPts, cst.k * 64 + xu * 16 * Prod::kUR, ub, r.x);
            else prod.template fetch<128>(g, cst.tile * kPts, cst.k, ub, r.x);
            store_unit(r);
          });
    } else {
      unit_pipeline<RawU>(my_items * nk * upc,
          [&](int, RawU& r) {
            const int xu = cl.u - wu;
            if (xu < 0) wload_k1(Wp + (size_t)cl.u * wps, Kp, cb * 128, cl.k * 64, g, r.w);
            else if constexpr (Prod::kChMajor) prod.template load<128>(g, cl.tile * kPts, cl.k * 64 + xu * 16 * Prod::kUR, r.x);
            else prod.template load<128>(g, cl.tile * kPts, cl.k, r.x);
            adv(cl);
          },
          [&](int, const RawU& r) { store_unit(r); });
    }
  } else {   // warp 16: MMA issue, warp-uniform loop, one elected lane issues
    const uint32_t tm = tc::uniform_u32(tmem_base);
    const uint32_t idesc = tc::make_idesc_bf16(128, kPts, false, Prod::kChMajor);
    int ring_s = 0, ring_r = 0, i = 0;
    for (int tile = t0; tile < ntiles; tile += tstep, ++i) {
      const int b = i & 1, u = i >> 1;
      for (int k = 0; k < nk; ++k) {
        tc::mbar_wait(&bar.full[ring_s], (uint32_t)(ring_r & 1));
        if (k == 0 && u > 0) tc::mbar_wait(&bar.tmem_empty[b], (uint32_t)((u - 1) & 1));
        tc::fence_after_sync();
        const uint32_t st = sS + (uint32_t)ring_s * sbytes;
        const uint32_t sA = wres ? sWres + (uint32_t)k * kOp : st, sB = st + xoff;
        const int kk = prod.chunk_k(k);
        for (int q = 0; q < kk / 16; ++q) {
          const uint64_t ad = tc::make_desc_sw128(sA + (uint32_t)q * 32, 16, 1024);
          const uint64_t bd = Prod::kChMajor ? tc::make_desc_sw128(sB + (uint32_t)q * 2048, 8192, 1024)
                                             : tc::make_desc_sw128(sB + (uint32_t)q * 32, 16, 1024);
          mma_planes<NP>(tm + (uint32_t)(b * kPts), ad, bd, idesc, k > 0 || q > 0);
        }
        tc::mma_commit_warp(&bar.empty[ring_s]);
        if (++ring_s == nst) { ring_s = 0; ++ring_r; }
      }
      tc::mma_commit_warp(&bar.tmem_full[b]);
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc<256>(tmem);
}

// ---------------------------------------------------------------------------------------------------------------
// dgrad (2 planes): dX^T[128 in-ch x 128 pts] = W^T[128 x Cout] * dY^T[Cout x 128] per (tile, input-channel block);
// contraction over the layer's output channels in chunks of 64.  PT: operands swapped, D[128 points x 128 in-ch],
// point-on-lane epilogue (layer-1 scatter-add into grad_feats).
// ---------------------------------------------------------------------------------------------------------------
template <class PProd, class Epi, bool PT, bool ASYNC = false>
__global__ void __launch_bounds__(kThreads, 1)
x3_dgrad_kernel(PProd pp, const __nv_bfloat16* __restrict__ Wp, size_t wps, int Kp, Epi epi, int M, int ncb, int nst, int wres) {
  PCOE_V6_PROLOGUE(256)
  constexpr int NP = 2;
  constexpr uint32_t kOp = NP * kPart;
  constexpr uint32_t kRawUnit = ASYNC ? (uint32_t)PProd::kRawItems * kRawItemBytes : 0u, kRawBytes = kRawDepth * kRawUnit;
  const int nk = pp.C / 64;
  const uint32_t sWres = smem0, wres_bytes = wres ? (uint32_t)nk * kOp : 0u;
  const uint32_t sS = smem0 + wres_bytes, sbytes = wres ? kOp : 2u * kOp, xoff = wres ? 0u : kOp;
  const uint32_t sRaw = sS + (uint32_t)nst * sbytes;
  float* csm = reinterpret_cast<float*>(smem_gen + wres_bytes + (size_t)nst * sbytes + kRawBytes);
  const int cb = blockIdx.x % ncb, t0 = blockIdx.x / ncb, tstep = gridDim.x / ncb;
  const int ntiles = (M + kPts - 1) / kPts;
  const int my_items = t0 < ntiles ? (ntiles - t0 + tstep - 1) / tstep : 0;
  zero_smem(sS, (uint32_t)nst * sbytes, tid, kThreads);
  pp.init(csm, tid, kThreads);
  const int eq = warp & 3, eh = (warp >> 2) & 1;
  epi.init(csm + pp.nconst(), cb * 128 + eq * 32 + lane);
  if (wres && tid < kProdThreads) {
    for (int k = 0; k < nk; ++k)
#pragma unroll
      for (int p = 0; p < NP; ++p) {
        uint4 w[4];
        wload_mn1(Wp + (size_t)p * wps, Kp, k * 64, cb * 128, tid, w);
        wstore_mn1(sWres + (uint32_t)k * kOp + (uint32_t)p * kPart, tid, w);
      }
  }
  tc::fence_proxy_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = tmem_base;

  if (warp < 8) {
    int i = 0;
    if constexpr (!PT) { if constexpr (Epi::kPre) epi.prime(t0, t0 + tstep, eh, M); }
    for (int tile = t0; tile < ntiles; tile += tstep, ++i) {
      const int b = i & 1, u = i >> 1, m0 = tile * kPts;
      if constexpr (!PT) { if constexpr (Epi::kPre) epi.pre(tile, tile + tstep, tile + 2 * tstep, eh, M); }
      tc::mbar_wait(&bar.tmem_full[b], (uint32_t)(u & 1));
      tc::fence_after_sync();
      if constexpr (!PT) {
        if constexpr (Epi::kHalf) {   // two threads per channel: 16 of a block's 32 points each
#pragma unroll 1
          for (int j = eh * 2; j < eh * 2 + 2; ++j) {
            float v[16];
            tc::tmem_ld16(tmem + ((uint32_t)(eq * 32) << 16) + (uint32_t)(b * kPts + j * 32 + (eq >> 1) * 16), v);
            epi.block16(v, tile, j, m0 + j * 32 < M);
          }
        } else {
#pragma unroll 1
          for (int j = eh * 2; j < eh * 2 + 2; ++j) {
            float v[32];
            tc::tmem_ld32(tmem + ((uint32_t)(eq * 32) << 16) + (uint32_t)(b * kPts + j * 32), v);
            epi.block(v, tile, j, m0 + j * 32 < M);
          }
        }
      } else {
        const int row = m0 + eq * 32 + lane;
#pragma unroll 1
        for (int cbk = eh * 32; cbk < 128; cbk += 64) {
          float v[32];
          tc::tmem_ld32(tmem + ((uint32_t)(eq * 32) << 16) + (uint32_t)(b * kPts + cbk), v);
          epi.block_pt(v, cb * 128 + cbk, row, row < M);
        }
      }
      tc::fence_before_sync();
      mbar_arrive_relaxed(&bar.tmem_empty[b]);
    }
    epi.finish();
  } else if (warp < 16) {
    const int g = tid - kEpiThreads;
    constexpr int kXU = 4 / PProd::kUR;
    const int wu = wres ? 0 : NP, upc = kXU + wu;
    union RawU { typename PProd::Raw x; uint4 w[4]; __device__ RawU() {} };
    struct Cur { int tile, k, u; };
    auto adv = [&](Cur& c) { if (++c.u == upc) { c.u = 0; if (++c.k == nk) { c.k = 0; c.tile += tstep; } } };
    Cur cl{t0, 0, 0}, cst = cl;
    int ring_s = 0, ring_r = 0;
    auto store_unit = [&](const RawU& r) {
      const uint32_t st = sS + (uint32_t)ring_s * sbytes;
      if (cst.u == 0 && ring_r > 0) tc::mbar_wait(&bar.empty[ring_s], (uint32_t)((ring_r - 1) & 1));
      const int xu = cst.u - wu;
      if (xu < 0) wstore_mn1(st + (uint32_t)cst.u * kPart, g, r.w);
      else pp.template store<128, NP>(g, cst.tile * kPts, cst.k * 64 + xu * 16 * PProd::kUR, xu * 16 * PProd::kUR, 64, r.x, st + xoff);
      if (cst.u == upc - 1) {
        tc::fence_proxy_async();
        mbar_arrive(&bar.full[ring_s]);
        if (++ring_s == nst) { ring_s = 0; ++ring_r; }
      }
      adv(cst);
    };
    if constexpr (ASYNC) {
      async_pipeline(my_items * nk * upc, sRaw, kRawUnit,
          [&](int, uint32_t ub) {
            const int xu = cl.u - wu;
            if (xu < 0) wload_mn1_async(Wp + (size_t)cl.u * wps, Kp, cl.k * 64, cb * 128, g, ub);
            else pp.template load_async<128>(g, cl.tile * kPts, cl.k * 64 + xu * 16 * PProd::kUR, ub);
            adv(cl);
          },
          [&](int, uint32_t ub) {
            RawU r;
            const int xu = cst.u - wu;
            if (xu < 0) wfetch(ub, g, r.w);
            else pp.template fetch<128>(g, cst.tile * kPts, cst.k * 64 + xu * 16 * PProd::kUR, ub, r.x);
            store_unit(r);
          });
    } else {
      unit_pipeline<RawU>(my_items * nk * upc,
          [&](int, RawU& r) {
            const int xu = cl.u - wu;
            if (xu < 0) wload_mn1(Wp + (size_t)cl.u * wps, Kp, cl.k * 64, cb * 128, g, r.w);
            else pp.template load<128>(g, cl.tile * kPts, cl.k * 64 + xu * 16 * PProd::kUR, r.x);
            adv(cl);
          },
          [&](int, const RawU& r) { store_unit(r); });
    }
  } else {
    const uint32_t tm = tc::uniform_u32(tmem_base);
    const uint32_t idesc = tc::make_idesc_bf16(128, 128, true, true);
    int ring_s = 0, ring_r = 0, i = 0;
    for (int tile = t0; tile < ntiles; tile += tstep, ++i) {
      const int b = i & 1, u = i >> 1;
      for (int k = 0; k < nk; ++k) {
        tc::mbar_wait(&bar.full[ring_s], (uint32_t)(ring_r & 1));
        if (k == 0 && u > 0) tc::mbar_wait(&bar.tmem_empty[b], (uint32_t)((u - 1) & 1));
        tc::fence_after_sync();
        const uint32_t st = sS + (uint32_t)ring_s * sbytes;
        const uint32_t sW = wres ? sWres + (uint32_t)k * kOp : st, sP = st + xoff;
        for (int q = 0; q < 4; ++q) {
          const uint64_t wd = tc::make_desc_sw128(sW + (uint32_t)q * 2048, 8192, 1024);
          const uint64_t pd = tc::make_desc_sw128(sP + (uint32_t)q * 2048, 8192, 1024);
          if constexpr (PT) mma_planes<NP>(tm + (uint32_t)(b * kPts), pd, wd, idesc, k > 0 || q > 0);
          else mma_planes<NP>(tm + (uint32_t)(b * kPts), wd, pd, idesc, k > 0 || q > 0);
        }
        tc::mma_commit_warp(&bar.empty[ring_s]);
        if (++ring_s == nst) { ring_s = 0; ++ring_r; }
      }
      tc::mma_commit_warp(&bar.tmem_full[b]);
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc<256>(tmem);
}

// p[0..32) += v[0..32): scalar reductions up to the first 16-byte boundary, red.v4 from there, scalars for the rest
template <int A>
__device__ __forceinline__ void red_add_row32_at(float* p, const float (&v)[32]) {
#pragma unroll
  for (int e = 0; e < A; ++e) atomicAdd(p + e, v[e]);
  constexpr int NV = (32 - A) / 4;
#pragma unroll
  for (int q = 0; q < NV; ++q) red_add_v4(p + A + 4 * q, v[A + 4 * q], v[A + 4 * q + 1], v[A + 4 * q + 2], v[A + 4 * q + 3]);
#pragma unroll
  for (int e = A + 4 * NV; e < 32; ++e) atomicAdd(p + e, v[e]);
}
__device__ __forceinline__ void red_add_row32(float* p, const float (&v)[32]) {
  const uint32_t lead = ((16u - ((uint32_t)(uintptr_t)p & 15u)) & 15u) >> 2;   // floats before the boundary
  if (lead == 0) red_add_row32_at<0>(p, v);
  else if (lead == 1) red_add_row32_at<1>(p, v);
  else if (lead == 2) red_add_row32_at<2>(p, v);
  else red_add_row32_at<3>(p, v);
}

// ---------------------------------------------------------------------------------------------------------------
// wgrad (2 planes): CTA (channel block, q block, split) accumulates dW[128 x nq] over its tiles in TMEM and adds it to
// global memory at the end.  Stage = 64 points: [P hi | P lo | Q hi | Q lo], P = dy^T part [128 ch x 64 pts] (K-major),
// Q = x_prev part, channel-major [128 ch x 64 pts] (K-major) or point-major [64 pts x 2 blocks of 64 ch] (MN-major).
// ---------------------------------------------------------------------------------------------------------------
template <class PProd, class QProd, bool ASYNC = false>
__global__ void __launch_bounds__(kThreads, 1)
x3_wgrad_kernel(PProd pp, QProd qp, float* __restrict__ dW, int ldo, int cq_valid, int perm_d, int M, int tps, int nst, int depths) {
  PCOE_V6_PROLOGUE(128)
  constexpr int NP = 2;
  constexpr uint32_t kUnitP = (uint32_t)PProd::kRawItems * kRawItemBytes;
  // ASYNC: raw staging rings of the two producer groups, depths = dP * 16 + dQ (see the producer branch)
  constexpr uint32_t kUnitQ = (uint32_t)QProd::kRawItems64 * kRawItemBytes;
  const uint32_t kRawBytes = ASYNC ? (uint32_t)(depths >> 4) * kUnitP + (uint32_t)(depths & 15) * kUnitQ : 0u;
  if (ASYNC && tid == 0)
    for (int s = 0; s < kMaxStages6; ++s) tc::mbar_init(&bar.full[s], 2 * kProdThreads);   // both groups arrive
  const int cl0 = blockIdx.x * 128, qb = blockIdx.y;
  const int ntiles = (M + kPts - 1) / kPts;
  const int t0 = blockIdx.z * tps, t1 = min(ntiles, t0 + tps), nt = max(t1 - t0, 0);
  const uint32_t sS = smem0, sbytes = 4u * kPart;
  const uint32_t sRaw = sS + (uint32_t)nst * sbytes;
  float* csm = reinterpret_cast<float*>(smem_gen + (size_t)nst * sbytes + kRawBytes);
  zero_smem(sS, (uint32_t)nst * sbytes, tid, kThreads);
  pp.init(csm, tid, kThreads);
  qp.init(csm + pp.nconst(), tid, kThreads);
  int nq, nqu;                                               // columns of this q block; Q units per stage
  if constexpr (QProd::kChMajor) {
    nq = min(128, qp.C - qb * 128);
    nqu = 4 / QProd::kUR;
  } else {
    nqu = min(2, qp.nchunks() - 2 * qb);
    nq = 0;
    for (int u = 0; u < nqu; ++u) nq += qp.chunk_k(2 * qb + u);
  }
  tc::fence_proxy_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = tmem_base;
  const int eq = warp & 3, eh = (warp >> 2) & 1;
  constexpr int kPU = 4 / PProd::kUR;                        // P units per stage (128 rows / (32 * kUR))
  const int ups = kPU + nqu, nstage = 2 * nt;                // units per stage; 64-point stages

  if (warp < 16) {
    // Both warp groups (0-7 and 8-15) are producers: the accumulator is read only once, after the last stage, so the
    // "epilogue" warps would otherwise idle for the whole kernel.  Group gsel builds the stages h = gsel, gsel + 2, ...
    // (each stage still gets its kProdThreads arrivals from one group), which doubles the loads in flight per SM.
    const int g = tid & (kProdThreads - 1);
    union RawU { typename PProd::Raw p; typename QProd::Raw q; __device__ RawU() {} };
    struct Cur { int u, m0; };
    if constexpr (ASYNC) {
      // Both warp groups produce, each through its own cp.async ring: warps 8-15 build the P (dy) units, warps 0-7 the
      // Q (x_prev) units of every stage (the accumulator is read once, after the last stage, so warps 0-7 would otherwise
      // idle; ncu: the 8-warp version was bound by the producers' instruction throughput, not by latency).  A stage is
      // complete when both groups have arrived (full barrier = 2 x kProdThreads).  Ring depths: `depths` = dP * 16 + dQ.
      const bool isP = warp >= 8;
      const int dP = depths >> 4, dQ = depths & 15;
      const int nu = isP ? kPU : nqu;
      const uint32_t rawP = sRaw, rawQ = sRaw + (uint32_t)dP * kUnitP;
      struct Cur2 { int u, m0; };
      Cur2 cl{0, t0 * kPts}, cst = cl;
      auto adv = [&](Cur2& c) { if (++c.u == nu) { c.u = 0; c.m0 += 64; } };
      int ring_s = 0, ring_r = 0;
      auto issue = [&](int, uint32_t ub) {
        if (isP) pp.template load_async<64>(g, cl.m0, cl0 + cl.u * 32 * PProd::kUR, ub);
        else if constexpr (QProd::kChMajor) qp.template load_async<64>(g, cl.m0, qb * 128 + cl.u * 32 * QProd::kUR, ub);
        else qp.template load_async<64>(g, cl.m0, 2 * qb + cl.u, ub);
        adv(cl);
      };
      auto consume = [&](int, uint32_t ub) {
        RawU r;
        const uint32_t st = sS + (uint32_t)ring_s * sbytes;
        if (cst.u == 0 && ring_r > 0) tc::mbar_wait(&bar.empty[ring_s], (uint32_t)((ring_r - 1) & 1));
        if (isP) {
          pp.template fetch<64>(g, cst.m0, cl0 + cst.u * 32 * PProd::kUR, ub, r.p);
          pp.template store<64, NP>(g, cst.m0, cl0 + cst.u * 32 * PProd::kUR, cst.u * 32 * PProd::kUR, 128, r.p, st);
        } else {
          const int qu = cst.u;
          if constexpr (QProd::kChMajor) {
            qp.template fetch<64>(g, cst.m0, qb * 128 + qu * 32 * QProd::kUR, ub, r.q);
            qp.template store<64, NP>(g, cst.m0, qb * 128 + qu * 32 * QProd::kUR, qu * 32 * QProd::kUR, 128, r.q, st + 2 * kPart);
          } else {
            qp.template fetch<64>(g, cst.m0, 2 * qb + qu, ub, r.q);
            qp.template store<64, NP>(g, 2 * qb + qu, r.q, st + 2 * kPart + (uint32_t)qu * 8192u);
          }
        }
        if (cst.u == nu - 1) {
          tc::fence_proxy_async();
          mbar_arrive(&bar.full[ring_s]);
          if (++ring_s == nst) { ring_s = 0; ++ring_r; }
        }
        adv(cst);
      };
      const uint32_t unitB = isP ? kUnitP : kUnitQ;
      if ((isP ? dP : dQ) == 3) async_pipeline<3>(nstage * nu, isP ? rawP : rawQ, unitB, issue, consume);
      else async_pipeline<2>(nstage * nu, isP ? rawP : rawQ, unitB, issue, consume);
    } else {
    const int gsel = warp >> 3;
    auto adv = [&](Cur& c) { if (++c.u == ups) { c.u = 0; c.m0 += 128; } };
    Cur cl{0, t0 * kPts + gsel * 64}, cst = cl;
    int ring_s = gsel % nst, ring_r = gsel / nst;
    const int my_stages = (nstage - gsel + 1) / 2;
    unit_pipeline<RawU>(my_stages * ups,
        [&](int, RawU& r) {
          if (cl.u < kPU) pp.template load<64>(g, cl.m0, cl0 + cl.u * 32 * PProd::kUR, r.p);
          else if constexpr (QProd::kChMajor) qp.template load<64>(g, cl.m0, qb * 128 + (cl.u - kPU) * 32 * QProd::kUR, r.q);
          else qp.template load<64>(g, cl.m0, 2 * qb + (cl.u - kPU), r.q);
          adv(cl);
        },
        [&](int, const RawU& r) {
          const uint32_t st = sS + (uint32_t)ring_s * sbytes;
          if (cst.u == 0 && ring_r > 0) tc::mbar_wait(&bar.empty[ring_s], (uint32_t)((ring_r - 1) & 1));
          if (cst.u < kPU)
            pp.template store<64, NP>(g, cst.m0, cl0 + cst.u * 32 * PProd::kUR, cst.u * 32 * PProd::kUR, 128, r.p, st);
          else {
            const int qu = cst.u - kPU;
            if constexpr (QProd::kChMajor)
              qp.template store<64, NP>(g, cst.m0, qb * 128 + qu * 32 * QProd::kUR, qu * 32 * QProd::kUR, 128, r.q, st + 2 * kPart);
            else
              qp.template store<64, NP>(g, 2 * qb + qu, r.q, st + 2 * kPart + (uint32_t)qu * 8192u);
          }
          if (cst.u == ups - 1) {
            tc::fence_proxy_async();
            mbar_arrive(&bar.full[ring_s]);
            ring_s += 2;
            while (ring_s >= nst) { ring_s -= nst; ++ring_r; }
          }
          adv(cst);
        });
    }
    if (warp < 8 && nt > 0) {
      tc::mbar_wait(&bar.tmem_full[0], 0u);
      tc::fence_after_sync();
      const int crow = cl0 + eq * 32 + lane;
#pragma unroll 1
      for (int cbk = eh * 32; cbk < nq; cbk += 64) {
        float v[32];
        tc::tmem_ld32(tmem + ((uint32_t)(eq * 32) << 16) + (uint32_t)cbk, v);   // may read past nq: unused columns
        if (crow >= pp.C) continue;
        float* dst = dW + (size_t)crow * ldo;
        const int cb = qb * 128 + cbk;
        if (cb + 32 <= cq_valid && cbk + 32 <= nq && (perm_d < 0 || cb + 32 <= perm_d)) {
          // 32 consecutive destination columns (shifted by the 3 xyz columns under the layer-1 permutation): vector
          // reductions wherever the row's address allows - every CTA of the grid adds into the same dW, and the
          // number of L2 reduction operations is what the tail of this kernel costs
          red_add_row32(dst + cb + (perm_d >= 0 ? 3 : 0), v);
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
  } else {
    const uint32_t tm = tc::uniform_u32(tmem_base);
    const uint32_t idesc = tc::make_idesc_bf16(128, nq, false, !QProd::kChMajor);
    int ring_s = 0, ring_r = 0;
    for (int h = 0; h < nstage; ++h) {
      tc::mbar_wait(&bar.full[ring_s], (uint32_t)(ring_r & 1));
      tc::fence_after_sync();
      const uint32_t sP = sS + (uint32_t)ring_s * sbytes, sQ = sP + 2 * kPart;
      for (int ks = 0; ks < 4; ++ks) {                        // 16 points per MMA
        const uint64_t ad = tc::make_desc_sw128(sP + (uint32_t)ks * 32, 16, 1024);
        const uint64_t bd = QProd::kChMajor ? tc::make_desc_sw128(sQ + (uint32_t)ks * 32, 16, 1024)
                                            : tc::make_desc_sw128(sQ + (uint32_t)ks * 2048, 8192, 1024);
        mma_planes<NP>(tm, ad, bd, idesc, h > 0 || ks > 0);
      }
      tc::mma_commit_warp(&bar.empty[ring_s]);
      if (++ring_s == nst) { ring_s = 0; ++ring_r; }
    }
    if (nt > 0) tc::mma_commit_warp(&bar.tmem_full[0]);
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc<128>(tmem);
}

}  // namespace v6
}  // namespace pcoe
