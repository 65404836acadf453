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
// channel-major and TILE-BLOCKED: y^T as [tile][C][128 points] bf16 with the eight-point (16-byte)
// chunks of a row XOR-swizzled by the channel: element (t, c, p) sits at byte
//     (t*C + c)*256 + (((p>>3) ^ (c&15)) << 4) + (p&7)*2 .
// A tile is one contiguous C*256-byte block: the epilogue stages it in shared memory in exactly this
// image (the swizzle makes the one-row-per-thread 16-byte stores bank-conflict free) and ONE bulk
// async copy per tile writes it out; consumers read 256-byte rows with coalesced 16-byte loads.
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
// CTA = 17 warps, persistent over 128-point tiles, one CTA per SM:
//   warps 0-7   epilogue   (warp w owns TMEM lanes 32(w%4)..+31 = one channel per thread and half of
//                           the tile's 32-column blocks, (w/4))
//   warps 8-15  producers  (global -> BN/ReLU or BN-backward transform -> bf16 -> swizzled smem)
//   warp  16    MMA issue  (one elected thread)
// Pipelines: smem stage ring (full/empty mbarriers), double-buffered TMEM accumulators
// (tmem_full/tmem_empty); dW accumulates in TMEM over the CTA's whole tile range.
#pragma once
#include "sa_common.cuh"
#include "sa_layout.h"
#include "tc_common.cuh"

namespace pcoe {
namespace v4 {

#ifdef PCOE_TC4_TRACE
// debug build: low-overhead clock trace of CTA 0.  Each traced thread records (tag, index, clock64)
// into a local array and dumps it when it leaves the kernel.
__device__ long long g_trace[8192];
__device__ int g_trace_n;
struct Tracer {
  long long buf[3 * 80];
  int n = 0;
  __device__ __forceinline__ void rec(int tag, int idx) {
    if (blockIdx.x == 0 && n < 80) { buf[3 * n] = tag; buf[3 * n + 1] = idx; buf[3 * n + 2] = clock64(); ++n; }
  }
  __device__ void dump() {
    if (blockIdx.x != 0 || n == 0) return;
    const int k = atomicAdd(&g_trace_n, n);
    for (int i = 0; i < n && k + i < 2700; ++i)
      for (int j = 0; j < 3; ++j) g_trace[3 * (k + i) + j] = buf[3 * i + j];
  }
};
#define TC4_TRACER Tracer tracer_
#define TC4_TRACE(tag, idx) tracer_.rec((tag), (idx))
#define TC4_TRACE_DUMP() tracer_.dump()
#else
#define TC4_TRACER do {} while (0)
#define TC4_TRACE(tag, idx) do {} while (0)
#define TC4_TRACE_DUMP() do {} while (0)
#endif

constexpr int kPts = 128;
constexpr int kEpiThreads = 256, kProdThreads = 256, kThreads = kEpiThreads + kProdThreads + 32;
constexpr int kMaxStages = 3;

__device__ __forceinline__ void mbar_arrive(uint64_t* mbar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc::smem_u32(mbar)) : "memory");
}
// arrive without release semantics: the epilogue's TMEM reads are already complete (tcgen05.wait::ld)
// and its global stores need no ordering against the barrier - a releasing arrive would wait for
// every outstanding global store of the thread to be acknowledged by L2
__device__ __forceinline__ void mbar_arrive_relaxed(uint64_t* mbar) {
  asm volatile("mbarrier.arrive.relaxed.cta.shared::cta.b64 _, [%0];" ::"r"(tc::smem_u32(mbar)) : "memory");
}

// producers' `full` arrival (experiment switch: relaxed arrive after the proxy fence)
__device__ __forceinline__ void mbar_arrive_full(uint64_t* mbar) {
#ifdef PCOE_EXP_RELAXED_ARRIVE
  mbar_arrive_relaxed(mbar);
#else
  mbar_arrive(mbar);
#endif
}

// byte offset of the 16-byte chunk (8 points) `chunk` (0..15) of channel row c in a channel-major tile
__device__ __forceinline__ uint32_t cm_off(int crows, int c, int chunk) {
  return (uint32_t)((chunk >> 3) * crows * 128 + (c >> 3) * 1024 + (c & 7) * 128 + ((((chunk & 7) ^ (c & 7))) << 4));
}

// 16-byte chunk `chunk` (points 8*chunk..+7) of channel c in tile `tile` of a tile-blocked activation
__device__ __forceinline__ const uint4* tb_chunk(const __nv_bfloat16* y, int C, int tile, int c, int chunk) {
  return reinterpret_cast<const uint4*>(reinterpret_cast<const char*>(y) + ((size_t)tile * C + c) * 256 + ((chunk ^ (c & 15)) << 4));
}
__device__ __forceinline__ uint32_t tb_stage_off(int c, int chunk) { return (uint32_t)(c * 256 + ((chunk ^ (c & 15)) << 4)); }

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

// Train-mode BatchNorm finalisation evaluated INLINE by the kernel that consumes it (instead of a
// per-channel kernel between two layers): every CTA derives (scale, shift) of the channels it needs
// from the fp64 batch sums the previous kernel accumulated; block (0,0,0) also writes the saved
// statistics for backward and updates the running buffers (bn_finalize_kernel's arithmetic, sa.cu).
struct BnFin {
  const double* sums;            // [2,C] sum, sum of squares; nullptr = not fused (read scale/shift arrays)
  double count, inv_count;       // inv_count = 1 / count from the host: no fp64 division / square root on the device
                                 // (every CTA evaluates this in its prologue; DDIV + DSQRT cost ~3 us per kernel)
  const float* gamma;
  const float* beta;
  const float* bias;             // conv bias: only shifts running_mean (cancelled by the batch-mean subtraction)
  float* running_mean;           // may be nullptr
  float* running_var;
  float eps, momentum;
  float* scale; float* shift; float* mean; float* invstd;   // saved for backward (written by one block)
  __device__ __forceinline__ void eval(int c, int C, bool write, float& sc, float& sh) const {
    double t0 = 0.0, t1 = 0.0;
#pragma unroll
    for (int k = 0; k < kRedCopies; ++k) { t0 += sums[(size_t)k * 2 * C + c]; t1 += sums[(size_t)k * 2 * C + C + c]; }
    const double mu = t0 * inv_count;
    double var = t1 * inv_count - mu * mu;   // biased, as BatchNorm normalises
    var = var < 0.0 ? 0.0 : var;
    const float ve = (float)(var + (double)eps);
    float is = rsqrtf(ve);
    is = is * (1.5f - 0.5f * ve * is * is);  // one Newton step: fp32-exact to ~1 ulp
    sc = gamma[c] * is;
    sh = beta[c] - (float)mu * sc;
    if (write) {
      scale[c] = sc; shift[c] = sh; mean[c] = (float)mu; invstd[c] = is;
      if (running_mean) {
        const float b = bias ? bias[c] : 0.f;
        const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
        running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)(mu + (double)b);
        running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
      }
    }
  }
};
// BatchNorm-backward constants (a, p, q) + parameter gradients, inline (bn_bwd_consts_kernel's arithmetic)
struct BnBwdFin {
  const double* sums;            // [2,C] sum dz, sum dz*xhat; nullptr = not fused (read a/p/q arrays)
  double count, inv_count;
  const float* scale; const float* mean; const float* invstd;
  float* dgamma; float* dbeta; float* dbias;
  int accumulate, write;         // write: this launch owns the parameter-gradient outputs
  __device__ __forceinline__ void eval(int c, int C, bool first_block, float& a, float& p, float& q) const {
    double s0 = 0.0, s1 = 0.0;
#pragma unroll
    for (int k = 0; k < kRedCopies; ++k) { s0 += sums[(size_t)k * 2 * C + c]; s1 += sums[(size_t)k * 2 * C + C + c]; }
    const double m1 = s0 * inv_count, m2 = s1 * inv_count;
    const double av = scale[c];
    const double pv = -av * (double)invstd[c] * m2;
    a = (float)av; p = (float)pv; q = (float)(-av * m1 - pv * (double)mean[c]);
    if (write && first_block) {
      if (dgamma) dgamma[c] = (accumulate ? dgamma[c] : 0.f) + (float)s1;
      if (dbeta) dbeta[c] = (accumulate ? dbeta[c] : 0.f) + (float)s0;
      if (dbias && !accumulate) dbias[c] = 0.f;
    }
  }
};
__device__ __forceinline__ bool first_block() { return (blockIdx.x | blockIdx.y | blockIdx.z) == 0; }

// ---------------------------------------------------------------------------------------------
// Producers.  A producer group of G threads builds one operand tile per 128 points in `nbatches(G)`
// batches of kBatch 16-byte units per thread, as a three-stage software pipeline so that global
// loads of the next batch (possibly of the next tile) are in flight while the current batch is
// transformed and stored:
//     load_idx(g, G, m0, b, Idx&)        neighbour indices of the batch (gather producers only)
//     load(g, G, m0, b, Idx, Raw&)       raw global loads into registers, no dependent arithmetic
//     store(g, G, m0, b, Raw, saddr)     transform -> bf16 -> swizzled shared memory
// kChMajor: tile image (see header).  rows(): rows of the image.  kext(): channel extent the MMAs
// read (multiple of 16).  nconst(): floats of per-channel constants cached in shared memory.
// Elements inside the K extent that are never written stay zero from the one-time clear.
// ---------------------------------------------------------------------------------------------
constexpr int kBatch = 4;
struct NoIdx {};

// Backward operands (dy = a*dz + p*y + q, and x_prev = relu(scale*y + shift) when it is the wgrad Q operand) are
// built with packed bf16x2 FMAs: 8 instead of ~36 arithmetic instructions per 8-element chunk - the CUDA-core
// transform is what bounds the backward kernels.  The per-channel constants are rounded to bf16 and dy takes one
// extra bf16 rounding (relative 2^-9 each, the size of the operand rounding the MMA needs anyway); the FORWARD
// activations keep fp32 arithmetic.  Set to false for fp32 transforms everywhere.
constexpr bool kPackedBwd = true;

// channel-major mapping: thread g -> 16-byte chunk (8 points) g&15 of channel rows c0 + 16 i, i < kBatch, where
// c0 = (g>>4) + 64 b for batch b (G = kProdThreads = 256 threads cover 64 channel rows per batch).  Consecutive rows
// of a thread are 16 channels apart, so the swizzle term (c & 15) / (c & 7) is the same for all of them and every
// address is base + i * constant: HBM tile rows +4096 B, shared-memory operand rows +2048 B, constants +16 floats.
constexpr int kRowStep = kProdThreads >> 4;            // 16
constexpr int kGRowBytes = kRowStep * 256;             // 4096
constexpr uint32_t kSRowBytes = (kRowStep >> 3) * 1024;   // 2048
#define PCOE_CM_MAP                                                        \
  const int chunk = g & 15, c0 = (g >> 4) + b * (kBatch * kRowStep);       \
  const int m = m0 + chunk * 8;                                            \
  const bool ok = m < M;                                                   \
  (void)G; (void)ok;
// byte offset of (channel c0, chunk) inside tile (m0 >> 7) of a tile-blocked activation with C channels
#define PCOE_TB_OFF(C_) (((size_t)(m0 >> 7) * (C_) + c0) * 256 + ((chunk ^ (c0 & 15)) << 4))

// relu(scale * y + shift) of the previous layer's pre-activations y^T [C][Mld]
struct BnRelu4 {
  static constexpr bool kChMajor = true;
  using Idx = NoIdx;
  struct Raw { uint4 a[kBatch]; };
  const __nv_bfloat16* __restrict__ y;
  const float* __restrict__ scale;
  const float* __restrict__ shift;
  int M, Mld, C;
  const float* cs;
  BnFin fin;     // fin.sums != nullptr: forward of a train-mode layer, statistics finalised here
  int packed;    // 1: backward Q operand, packed bf16x2 arithmetic (kPackedBwd)
  int prows;     // rows ALLOCATED for the image (0 = C): the Gram forward appends a row of ones and pads to the MMA M
  __host__ __device__ __forceinline__ int rows() const { return prows ? prows : C; }
  __host__ __device__ __forceinline__ int kext() const { return C; }
  __host__ __device__ __forceinline__ int nconst() const { return 2 * C; }
  __device__ __forceinline__ int nbatches(int G) const { return (C / (G >> 4) + kBatch - 1) / kBatch; }
  __device__ __forceinline__ void init(float* csm, int tid, int nthr) {
    if (fin.sums) {
      const bool w = first_block();
      for (int c = tid; c < C; c += nthr) fin.eval(c, C, w, csm[c], csm[C + c]);
    } else {
      for (int c = tid; c < C; c += nthr) { csm[c] = scale[c]; csm[C + c] = shift[c]; }
    }
    cs = csm;
  }
  __device__ __forceinline__ void load_idx(int, int, int, int, Idx&) const {}
  __device__ __forceinline__ void load(int g, int G, int m0, int b, const Idx&, Raw& r) const {
    PCOE_CM_MAP
    const char* src = reinterpret_cast<const char*>(y) + PCOE_TB_OFF(C);
#pragma unroll
    for (int i = 0; i < kBatch; ++i)
      r.a[i] = ok ? __ldg(reinterpret_cast<const uint4*>(src + i * kGRowBytes)) : make_uint4(0, 0, 0, 0);
  }
  // lrows / rshift (v5 chunk streaming): the image has lrows rows and channel c goes to row c + rshift
  __device__ __forceinline__ void store(int g, int G, int m0, int b, const Raw& r, uint32_t saddr, int lrows = 0,
                                        int rshift = 0) const {
    PCOE_CM_MAP
    const uint32_t dst = saddr + cm_off(lrows ? lrows : rows(), c0 + rshift, chunk);
    const float* k0 = cs + c0;
    if (kPackedBwd && packed) {
#pragma unroll
      for (int i = 0; i < kBatch; ++i) {
        const uint32_t sc2 = tc::bf2_bcast(k0[i * kRowStep]), sh2 = tc::bf2_bcast(k0[C + i * kRowStep]);
        uint4 o;
        o.x = tc::bf2_fma_relu(r.a[i].x, sc2, sh2); o.y = tc::bf2_fma_relu(r.a[i].y, sc2, sh2);
        o.z = tc::bf2_fma_relu(r.a[i].z, sc2, sh2); o.w = tc::bf2_fma_relu(r.a[i].w, sc2, sh2);
        tc::sts128(dst + i * kSRowBytes, o);
      }
      return;
    }
#pragma unroll
    for (int i = 0; i < kBatch; ++i) {
      // branch-free on purpose: the four chunks of a batch are independent and the compiler interleaves them
      // (C is a multiple of the 64 channels a batch covers; columns of points >= M hold relu(shift): they are
      // never read back - forward / dgrad epilogues skip them and the wgrad dy operand zeroes them)
#ifdef PCOE_EXP_NOXFORM
      tc::sts128(dst + i * kSRowBytes, r.a[i]);
      continue;
#endif
      float v[8];
      unpack8(r.a[i], v);
      const float sc = k0[i * kRowStep], sh = k0[C + i * kRowStep];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = fmaf(v[u], sc, sh);
      tc::sts128(dst + i * kSRowBytes, tc::pack8_bf16_relu(v));
    }
  }
};

// dy^T = a*dz^T + p*y^T + q  (BatchNorm backward folded into per-channel constants), dense dz
struct Dy4 {
  static constexpr bool kChMajor = true;
  using Idx = NoIdx;
  struct Raw { uint4 d[kBatch], y[kBatch]; };
  const __nv_bfloat16* __restrict__ dz;
  const __nv_bfloat16* __restrict__ y;
  const float* __restrict__ a;
  const float* __restrict__ p;
  const float* __restrict__ q;
  int M, Mld, C;
  const float* cs;
  BnBwdFin fin;  // fin.sums != nullptr: BatchNorm-backward constants derived here from the batch sums
  __host__ __device__ __forceinline__ int rows() const { return C < 128 ? 128 : C; }   // also an M operand: 128 rows
  __host__ __device__ __forceinline__ int kext() const { return C; }
  __host__ __device__ __forceinline__ int nconst() const { return 3 * C; }
  __device__ __forceinline__ int nbatches(int G) const { return (C / (G >> 4) + kBatch - 1) / kBatch; }
  __device__ __forceinline__ void init(float* csm, int tid, int nthr) {
    if (fin.sums) {
      const bool w = first_block();
      for (int c = tid; c < C; c += nthr) fin.eval(c, C, w, csm[c], csm[C + c], csm[2 * C + c]);
    } else {
      for (int c = tid; c < C; c += nthr) { csm[c] = a[c]; csm[C + c] = p[c]; csm[2 * C + c] = q[c]; }
    }
    cs = csm;
  }
  __device__ __forceinline__ void load_idx(int, int, int, int, Idx&) const {}
  __device__ __forceinline__ void load(int g, int G, int m0, int b, const Idx&, Raw& r) const {
    PCOE_CM_MAP
    const size_t off = PCOE_TB_OFF(C);
    const char* sd = reinterpret_cast<const char*>(dz) + off;
    const char* sy = reinterpret_cast<const char*>(y) + off;
#pragma unroll
    for (int i = 0; i < kBatch; ++i) {
      r.d[i] = ok ? __ldg(reinterpret_cast<const uint4*>(sd + i * kGRowBytes)) : make_uint4(0, 0, 0, 0);
      r.y[i] = ok ? __ldg(reinterpret_cast<const uint4*>(sy + i * kGRowBytes)) : make_uint4(0, 0, 0, 0);
    }
  }
  __device__ __forceinline__ void store(int g, int G, int m0, int b, const Raw& r, uint32_t saddr, int lrows = 0,
                                        int rshift = 0) const {
    PCOE_CM_MAP
    const uint32_t dst = saddr + cm_off(lrows ? lrows : rows(), c0 + rshift, chunk);
    const float* k0 = cs + c0;
    const float okf = ok ? 1.f : 0.f;
#pragma unroll
    for (int i = 0; i < kBatch; ++i) {                 // branch-free: see BnRelu4::store
      const float ca = k0[i * kRowStep] * okf, cp = k0[C + i * kRowStep] * okf, cq = k0[2 * C + i * kRowStep] * okf;   // points >= M contribute 0 to dW
      if constexpr (kPackedBwd) {
        const uint32_t a2 = tc::bf2_bcast(ca), p2 = tc::bf2_bcast(cp), q2 = tc::bf2_bcast(cq);
        uint4 o;
        o.x = tc::bf2_fma(a2, r.d[i].x, tc::bf2_fma(p2, r.y[i].x, q2));
        o.y = tc::bf2_fma(a2, r.d[i].y, tc::bf2_fma(p2, r.y[i].y, q2));
        o.z = tc::bf2_fma(a2, r.d[i].z, tc::bf2_fma(p2, r.y[i].z, q2));
        o.w = tc::bf2_fma(a2, r.d[i].w, tc::bf2_fma(p2, r.y[i].w, q2));
        tc::sts128(dst + i * kSRowBytes, o);
      } else {
        float d[8], yy[8], v[8];
        unpack8(r.d[i], d);
        unpack8(r.y[i], yy);
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = fmaf(ca, d[u], fmaf(cp, yy[u], cq));
        tc::sts128(dst + i * kSRowBytes, tc::pack8_bf16(v));
      }
    }
  }
};

// last layer: the upstream gradient is the max-pool routing of gm[G,C] to the saved arg slot (K == 32)
struct DyLast4 {
  static constexpr bool kChMajor = true;
  using Idx = NoIdx;
  struct Raw { uint4 y[kBatch]; float gv[kBatch]; int sl[kBatch]; };
  const float* __restrict__ gm;       // [G,C]
  const uint8_t* __restrict__ slot;   // [G,C]
  const __nv_bfloat16* __restrict__ y;
  const float* __restrict__ a;
  const float* __restrict__ p;
  const float* __restrict__ q;
  int M, Mld, C;
  const float* cs;
  BnBwdFin fin;
  __host__ __device__ __forceinline__ int rows() const { return C < 128 ? 128 : C; }
  __host__ __device__ __forceinline__ int kext() const { return C; }
  __host__ __device__ __forceinline__ int nconst() const { return 3 * C; }
  __device__ __forceinline__ int nbatches(int G) const { return (C / (G >> 4) + kBatch - 1) / kBatch; }
  __device__ __forceinline__ void init(float* csm, int tid, int nthr) {
    if (fin.sums) {
      const bool w = first_block();
      for (int c = tid; c < C; c += nthr) fin.eval(c, C, w, csm[c], csm[C + c], csm[2 * C + c]);
    } else {
      for (int c = tid; c < C; c += nthr) { csm[c] = a[c]; csm[C + c] = p[c]; csm[2 * C + c] = q[c]; }
    }
    cs = csm;
  }
  __device__ __forceinline__ void load_idx(int, int, int, int, Idx&) const {}
  __device__ __forceinline__ void load(int g, int G, int m0, int b, const Idx&, Raw& r) const {
    PCOE_CM_MAP
    const int grp = min(m, M - 1) >> 5;
    const char* sy = reinterpret_cast<const char*>(y) + PCOE_TB_OFF(C);
    const float* sg = gm + (size_t)grp * C + c0;
    const uint8_t* ss = slot + (size_t)grp * C + c0;
#pragma unroll
    for (int i = 0; i < kBatch; ++i) {
      r.y[i] = ok ? __ldg(reinterpret_cast<const uint4*>(sy + i * kGRowBytes)) : make_uint4(0, 0, 0, 0);
      r.gv[i] = ok ? __ldg(sg + i * kRowStep) : 0.f;
      r.sl[i] = ok ? (int)__ldg(ss + i * kRowStep) : -1;   // raw: no arithmetic on loaded values in load()
    }
  }
  __device__ __forceinline__ void store(int g, int G, int m0, int b, const Raw& r, uint32_t saddr, int lrows = 0,
                                        int rshift = 0) const {
    PCOE_CM_MAP
    const uint32_t dst = saddr + cm_off(lrows ? lrows : rows(), c0 + rshift, chunk);
    const float* k0 = cs + c0;
    const float okf = ok ? 1.f : 0.f;
    const int j0 = m & 31;
#pragma unroll
    for (int i = 0; i < kBatch; ++i) {                 // branch-free: see BnRelu4::store
      const float ca = k0[i * kRowStep] * r.gv[i], cp = k0[C + i * kRowStep] * okf, cq = k0[2 * C + i * kRowStep] * okf;   // gv is 0 for points >= M
      const int sl = r.sl[i] - j0;           // slot relative to this 8-point chunk (-1 - j0 < 0 never matches)
      if constexpr (kPackedBwd) {
        const uint32_t p2 = tc::bf2_bcast(cp), q2 = tc::bf2_bcast(cq);
        tc::sts128(dst + i * kSRowBytes, make_uint4(tc::bf2_fma(p2, r.y[i].x, q2), tc::bf2_fma(p2, r.y[i].y, q2),
                                                    tc::bf2_fma(p2, r.y[i].z, q2), tc::bf2_fma(p2, r.y[i].w, q2)));
      } else {
        float yy[8], v[8];
        unpack8(r.y[i], yy);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          v[u] = fmaf(cp, yy[u], cq);
          if (u == sl) v[u] += ca;           // the max-pool gradient lands on one point of the group
        }
        tc::sts128(dst + i * kSRowBytes, tc::pack8_bf16(v));
      }
    }
    if constexpr (kPackedBwd) {
      // the max-pool gradient lands on ONE point of each 32-point group: a separate pass (the loop above stays
      // branch-free) adds it in fp32 to that bf16 element of the chunk this thread has just written
#pragma unroll
      for (int i = 0; i < kBatch; ++i) {
        const int sl = r.sl[i] - j0;
        if ((unsigned)sl < 8u) {
          const uint32_t addr = dst + i * kSRowBytes + (uint32_t)sl * 2u;
          uint16_t h;
          asm volatile("ld.shared.u16 %0, [%1];" : "=h"(h) : "r"(addr) : "memory");
          const float f = __uint_as_float((uint32_t)h << 16) + k0[i * kRowStep] * r.gv[i];
          h = __bfloat16_as_ushort(__float2bfloat16_rn(f));
          asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"(h) : "memory");
        }
      }
    }
  }
};

// last layer WITHOUT the saved pre-activations y3 (v4 train path).  BatchNorm backward of the last layer is
//     dy3 = a*g + p*y3 + q,     g = max-pool routing of gm (one non-zero per group and channel),  y3 = W3 x2
// and its dense part is linear in x2, so it never has to be materialised:
//     dx2 = W3^T dy3 = W3^T (a*g) + (W3^T diag(p) W3) x2 + W3^T q          (Gm image + per-channel constant r)
//     dW3 = dy3^T x2 = (a*g)^T x2 + diag(p) W3 (x2^T x2) + q (sum x2)^T     (Gram matrix from the forward kernel)
// This producer builds only the SPARSE operand a*g (gm arrives pre-multiplied by a = scale): a zero tile with one
// bf16 value per (group, channel).  y3 is neither written by the forward nor read here.
struct DySparse4 {
  static constexpr bool kChMajor = true;
  using Idx = NoIdx;
  // ONE unit per tile: thread g owns up to two HALF rows (64 points = 2 groups of one channel): half-row index
  // hr = g + 256 u < 2 C, half = hr / C, channel = hr % C.  It clears its 8 chunks (8 x 16-byte stores of zero, no
  // dependence on the loads) and then drops the two max-pool gradients in (same thread: program order, no race).
  // The dense formulation needed C / 64 units per tile, each with its own load -> wait -> store latency chain.
  struct Raw { float gv[4]; int sl[4]; };
  const float* __restrict__ gm;       // [G,C], already multiplied by a
  const uint8_t* __restrict__ slot;   // [G,C]
  int M, Mld, C;
  __host__ __device__ __forceinline__ int rows() const { return C < 128 ? 128 : C; }
  __host__ __device__ __forceinline__ int kext() const { return C; }
  __host__ __device__ __forceinline__ int nconst() const { return 0; }
  __device__ __forceinline__ int nbatches(int) const { return 1; }
  __device__ __forceinline__ void init(float*, int, int) {}
  __device__ __forceinline__ void load_idx(int, int, int, int, Idx&) const {}
  __device__ __forceinline__ void load(int g, int G, int m0, int, const Idx&, Raw& r) const {
    const int g0 = m0 >> 5, ng = (M + 31) >> 5;
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int hr = g + u * G;
      const bool own = hr < 2 * C;
      const int half = hr >= C ? 1 : 0, c = hr - half * C;
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int grp = g0 + 2 * half + j;
        const bool ok = own && grp < ng;
        const size_t o = (size_t)(ok ? grp : 0) * C + (own ? c : 0);
        r.gv[2 * u + j] = ok ? __ldg(gm + o) : 0.f;
        r.sl[2 * u + j] = ok ? (int)__ldg(slot + o) : -1;
      }
    }
  }
  __device__ __forceinline__ void store(int g, int G, int, int, const Raw& r, uint32_t saddr, int = 0, int = 0) const {
    const uint4 z = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int hr = g + u * G;
      if (hr >= 2 * C) continue;
      const int half = hr >= C ? 1 : 0, c = hr - half * C;
      const uint32_t base = saddr + (uint32_t)half * (uint32_t)(rows() * 128) + (uint32_t)((c >> 3) * 1024 + (c & 7) * 128);
#pragma unroll
      for (int k = 0; k < 8; ++k) tc::sts128(base + (uint32_t)(((k + c) & 7) << 4), z);   // rotated by the row: the lanes of a
                                                                                            // warp (rows 128 B apart) hit 8 bank groups, not one
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int sl = r.sl[2 * u + j];
        if (sl >= 0) {
          const int pos = j * 32 + sl;
          const uint16_t h = __bfloat16_as_ushort(__float2bfloat16_rn(r.gv[2 * u + j]));
          asm volatile("st.shared.u16 [%0], %1;" ::"r"(base + (uint32_t)((((pos >> 3) ^ (c & 7)) << 4) + (pos & 7) * 2)), "h"(h) : "memory");
        }
      }
    }
  }
};

// shared by the gather producers: flat point index of neighbour row `row`
struct GatherBase {
  const float* __restrict__ xyz;
  const float* __restrict__ new_xyz;
  const int32_t* __restrict__ nbr;
  int N, S, group_all, M;
  __device__ __forceinline__ int point_of(int row) const {   // row < M
    if (group_all) return row;
    int i = __ldg(nbr + row);
    i = min(max(i, 0), N - 1);
    return ((row >> 5) / S) * N + i;
  }
  // The same for row m0 + r of the 128-row tile starting at m0, with ONE integer division per call site instead of
  // one per row: (cloud0, rem0) = divmod(first group of the tile, S) is passed in, the <= 3 further groups of the
  // tile only step the cloud index.
  __device__ __forceinline__ void tile_origin(int m0, int& cloud0, int& rem0) const {
    const int g0 = m0 >> 5;
    cloud0 = group_all ? 0 : g0 / S;
    rem0 = g0 - cloud0 * S;
  }
  __device__ __forceinline__ int point_of_tile(int m0, int r, int cloud0, int rem0) const {   // m0 + r < M
    if (group_all) return m0 + r;
    int i = __ldg(nbr + m0 + r);
    i = min(max(i, 0), N - 1);
    int q = rem0 + (r >> 5), cl = cloud0;
    while (q >= S) { q -= S; ++cl; }
    return cl * N + i;
  }
  // raw loads only (the subtraction happens in the producers' store(): no arithmetic on loaded values
  // in load(), otherwise the load stage of the software pipeline stalls on its own loads)
  __device__ __forceinline__ void load_xyz_raw(int row, int pt, float (&x)[3], float (&c)[3]) const {
#pragma unroll
    for (int u = 0; u < 3; ++u) {
      x[u] = __ldg(xyz + (size_t)pt * 3 + u);
      c[u] = group_all ? 0.f : __ldg(new_xyz + (size_t)(row >> 5) * 3 + u);
    }
  }
  __device__ __forceinline__ float centred(float x, float c) const { return group_all ? x : __fsub_rn(x, c); }
};

// layer-1 input without features (SA1): [xyz[nbr] - centroid] as a channel-major tile of 16 rows
// (rows 3..15 stay zero from the one-time clear).  Thread g < 128 owns point m0 + g.
struct GatherXyz4 {
  static constexpr bool kChMajor = true;
  struct Idx { int pt; };
  struct Raw { float v[3], c[3]; };
  GatherBase gb;
  __host__ __device__ __forceinline__ int rows() const { return 16; }
  __host__ __device__ __forceinline__ int kext() const { return 16; }
  __host__ __device__ __forceinline__ int nconst() const { return 0; }
  __device__ __forceinline__ int nbatches(int) const { return 1; }
  __device__ __forceinline__ void init(float*, int, int) {}
  __device__ __forceinline__ void load_idx(int g, int, int m0, int, Idx& ix) const {
    ix.pt = -1;
    if (g < kPts && m0 + g < gb.M) ix.pt = gb.point_of(m0 + g);   // one row per thread: one division
  }
  __device__ __forceinline__ void load(int g, int, int m0, int, const Idx& ix, Raw& r) const {
    r.v[0] = r.v[1] = r.v[2] = r.c[0] = r.c[1] = r.c[2] = 0.f;
    if (ix.pt >= 0) gb.load_xyz_raw(m0 + g, ix.pt, r.v, r.c);
  }
  __device__ __forceinline__ void store(int g, int, int, int, const Raw& r, uint32_t saddr) const {
    if (g >= kPts) return;
    const uint32_t base = saddr + (uint32_t)((g >> 6) * 16 * 128) + (uint32_t)((g & 7) * 2);
    const int ch = (g & 63) >> 3;
#pragma unroll
    for (int u = 0; u < 3; ++u) {
      const uint16_t h = __bfloat16_as_ushort(__float2bfloat16_rn(gb.centred(r.v[u], r.c[u])));
      asm volatile("st.shared.b16 [%0], %1;" ::"r"(base + (uint32_t)(u * 128 + ((ch ^ u) << 4))), "h"(h) : "memory");
    }
  }
};

// layer-1 input with features (SA2, SA3): point-major tile [128 points][kext channels],
// channel order [feats(D) | xyz - centroid (3) | zeros], D in {32, 64, 128, 256}.
// Thread g owns feature unit (g % fu) (8 channels, 32 bytes of fp32) of rows (g / fu) + j*(G / fu);
// in batch 0 thread g < 128 additionally owns the xyz unit of row g.
struct GatherFeat4 {
  static constexpr bool kChMajor = false;
  struct Idx { int pt[kBatch]; int ptx; };
  struct Raw { float4 a[kBatch], b[kBatch]; float xv[3], xc[3]; };
  GatherBase gb;
  const float* __restrict__ feats;
  int D;
  __host__ __device__ __forceinline__ int kext() const { return (D + 3 + 15) / 16 * 16; }
  __host__ __device__ __forceinline__ int rows() const { return kPts; }
  __host__ __device__ __forceinline__ int nconst() const { return 0; }
  __device__ __forceinline__ int nbatches(int G) const { return (kPts / (G / (D >> 3)) + kBatch - 1) / kBatch; }   // once per kernel
  __device__ __forceinline__ void init(float*, int, int) {}
  __device__ __forceinline__ void load_idx(int g, int G, int m0, int b, Idx& ix) const {
    const int fu = D >> 3, fs = 31 - __clz(fu), rstep = G >> fs, r0 = g >> fs;   // D in {32, 64, 128}: fu is a power of two
    int cloud0, rem0;
    gb.tile_origin(m0, cloud0, rem0);
#pragma unroll
    for (int i = 0; i < kBatch; ++i) {
      const int r = r0 + (b * kBatch + i) * rstep;
      ix.pt[i] = (r < kPts && m0 + r < gb.M) ? gb.point_of_tile(m0, r, cloud0, rem0) : -1;
    }
    ix.ptx = (b == 0 && g < kPts && m0 + g < gb.M) ? gb.point_of_tile(m0, g, cloud0, rem0) : -1;
  }
  __device__ __forceinline__ void load(int g, int G, int m0, int b, const Idx& ix, Raw& r) const {
    const int fu = D >> 3, j = g & (fu - 1);
#pragma unroll
    for (int i = 0; i < kBatch; ++i) {
      r.a[i] = r.b[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (ix.pt[i] >= 0) {
        const float4* src = reinterpret_cast<const float4*>(feats + (size_t)ix.pt[i] * D + j * 8);
        r.a[i] = __ldg(src);
        r.b[i] = __ldg(src + 1);
      }
    }
    r.xv[0] = r.xv[1] = r.xv[2] = r.xc[0] = r.xc[1] = r.xc[2] = 0.f;
    if (b == 0 && ix.ptx >= 0) gb.load_xyz_raw(m0 + g, ix.ptx, r.xv, r.xc);
  }
  __device__ __forceinline__ void store(int g, int G, int, int b, const Raw& r, uint32_t saddr) const {
    const int fu = D >> 3, fs = 31 - __clz(fu), rstep = G >> fs, r0 = g >> fs, j = g & (fu - 1);
#pragma unroll
    for (int i = 0; i < kBatch; ++i) {
      const int row = r0 + (b * kBatch + i) * rstep;
      if (row >= kPts) continue;
      const float v[8] = {r.a[i].x, r.a[i].y, r.a[i].z, r.a[i].w, r.b[i].x, r.b[i].y, r.b[i].z, r.b[i].w};
      tc::sts128(saddr + (uint32_t)(j >> 3) * (kPts * 128) + tc::sw128_off(row, (j & 7) * 8), tc::pack8_bf16(v));
    }
    if (b == 0 && g < kPts) {
      const float v[8] = {gb.centred(r.xv[0], r.xc[0]), gb.centred(r.xv[1], r.xc[1]), gb.centred(r.xv[2], r.xc[2]),
                          0.f, 0.f, 0.f, 0.f, 0.f};
      tc::sts128(saddr + (uint32_t)(fu >> 3) * (kPts * 128) + tc::sw128_off(g, (fu & 7) * 8), tc::pack8_bf16(v));
    }
  }
};

// Global loads kept in flight per producer thread, in units (batches): HBM latency x bandwidth needs ~44 KB in
// flight per SM; one unit of 256 threads x kBatch x 16 B is 16 KB.  Two units ahead when the raw registers of a
// unit are small (<= 16 registers), one otherwise (register budget: 120 per thread at 544 threads).
template <class Raw>
struct LoadDepth { static constexpr int value = sizeof(Raw) <= 64 ? 2 : 1; };

// The producer pipeline of one group (G threads, group-local id g) over the CTA's tiles.
// arrive(): `full` arrival after the last batch of a tile; wait_empty(t): the tile's stage is free.
template <class Prod, class WaitEmpty, class StageAddr, class Arrive>
__device__ __forceinline__ void producer_pipeline(const Prod& prod, int g, int G, int ntiles, WaitEmpty wait_empty,
                                                  StageAddr stage_addr, Arrive arrive) {
  constexpr int D = LoadDepth<typename Prod::Raw>::value;
  int my_tiles = 0;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) ++my_tiles;
  const int nb = prod.nbatches(G), W = my_tiles * nb;
  if (W == 0) return;
  TC4_TRACER;
  // cursors (tile ordinal t, batch b, first row m0) of the index / load / store stages, advanced
  // incrementally: no integer division in the steady state
  struct Cur { int t, b, m0; };
  const int mstep = (int)gridDim.x * kPts;
  auto adv = [&](Cur& c) { if (++c.b == nb) { c.b = 0; ++c.t; c.m0 += mstep; } };
  Cur ci{0, 0, (int)blockIdx.x * kPts}, cl = ci, cst = ci;
  typename Prod::Idx ix;
  typename Prod::Raw r[D + 1];   // unit u lives in r[u % (D + 1)]; indices are compile-time after unrolling
  prod.load_idx(g, G, ci.m0, ci.b, ix); adv(ci);
#pragma unroll
  for (int u = 0; u < D; ++u) {
    if (u < W) { prod.load(g, G, cl.m0, cl.b, ix, r[u]); adv(cl); }
    if (u + 1 < W) { prod.load_idx(g, G, ci.m0, ci.b, ix); adv(ci); }
  }
  auto step = [&](int w, typename Prod::Raw& cur, typename Prod::Raw& nxt) {
    if (w + D < W) { prod.load(g, G, cl.m0, cl.b, ix, nxt); adv(cl); }
    if (w + D + 1 < W) { prod.load_idx(g, G, ci.m0, ci.b, ix); adv(ci); }
    if (g == 0) TC4_TRACE(12, w);
    if (cst.b == 0) wait_empty(cst.t);
    if (g == 0) TC4_TRACE(10, w);
    prod.store(g, G, cst.m0, cst.b, cur, stage_addr(cst.t));
    if (cst.b == nb - 1) { tc::fence_proxy_async(); arrive(cst.t); }
    if (g == 0) TC4_TRACE(11, w);
    adv(cst);
  };
  for (int w = 0; w < W; w += D + 1) {
#pragma unroll
    for (int u = 0; u <= D; ++u)
      if (w + u < W) step(w + u, r[u], r[(u + D) % (D + 1)]);
  }
  if (g == 0) TC4_TRACE_DUMP();
}

// ---------------------------------------------------------------------------------------------
// Epilogues.  Thread = one channel (TMEM lane) for the whole kernel; v = 32 consecutive points.
//   init(csm, c): c = this thread's channel (may be >= C: padding lane -> the thread does nothing)
//   block(v, m, valid): points m..m+31      finish(): flush the per-thread running sums
// Reductions are written as 4-way / tree reductions: one epilogue warp per scheduler has to cover
// its own ALU latency.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void store32_bf16(__nv_bfloat16* dst, const float (&v)[32]) {
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float t[8] = {v[8 * q], v[8 * q + 1], v[8 * q + 2], v[8 * q + 3], v[8 * q + 4], v[8 * q + 5], v[8 * q + 6], v[8 * q + 7]};
    reinterpret_cast<uint4*>(dst)[q] = tc::pack8_bf16(t);
  }
}

// Epilogues with kStage write their output tile into a shared-memory staging image (the tile-blocked
// HBM layout, see header); the kernel copies it out with one bulk async copy per tile.
__device__ __forceinline__ void stage32_bf16(uint32_t stg, int c, int j, const float (&v)[32]) {
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float t[8] = {v[8 * q], v[8 * q + 1], v[8 * q + 2], v[8 * q + 3], v[8 * q + 4], v[8 * q + 5], v[8 * q + 6], v[8 * q + 7]};
    tc::sts128(stg + tb_stage_off(c, j * 4 + q), tc::pack8_bf16(t));
  }
}

__device__ __forceinline__ void sum_sumsq32(const float (&v)[32], float& s, float& ss) {
  float a[4] = {0.f, 0.f, 0.f, 0.f}, b[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int i = 0; i < 32; i += 4) {
#pragma unroll
    for (int k = 0; k < 4; ++k) { a[k] += v[i + k]; b[k] = fmaf(v[i + k], v[i + k], b[k]); }
  }
  s += (a[0] + a[1]) + (a[2] + a[3]);
  ss += (b[0] + b[1]) + (b[2] + b[3]);
}

// first arg-max (MAX) / arg-min of 32 values by a pairwise tree (depth 5); ties keep the lower index
template <bool MAX>
__device__ __forceinline__ void argext32(const float (&v)[32], float& val, int& arg) {
  float mv[16];
  int ix[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const bool r = MAX ? v[2 * i + 1] > v[2 * i] : v[2 * i + 1] < v[2 * i];
    mv[i] = r ? v[2 * i + 1] : v[2 * i];
    ix[i] = r ? 2 * i + 1 : 2 * i;
  }
#pragma unroll
  for (int n = 8; n >= 1; n >>= 1) {
#pragma unroll
    for (int i = 0; i < n; ++i) {
      const bool r = MAX ? mv[2 * i + 1] > mv[2 * i] : mv[2 * i + 1] < mv[2 * i];
      const float nv = r ? mv[2 * i + 1] : mv[2 * i];
      const int ni = r ? ix[2 * i + 1] : ix[2 * i];
      mv[i] = nv;
      ix[i] = ni;
    }
  }
  val = mv[0];
  arg = ix[0];
}

// kHalf epilogues (C == 64, weight image rows / columns duplicated): TMEM lanes 64..127 repeat channels 0..63, so the two
// threads of a channel each take HALF of a 32-point block (16 columns): the per-tile epilogue chain - the stage that
// bounds these kernels - is half as long, at no extra MMA cost (the MMA is M = 128 either way).
struct StoreStats4 {
  __nv_bfloat16* __restrict__ y;   // tile-blocked [tile][C][128]
  double* __restrict__ sums;       // [2,C] or nullptr (eval)
  int C, Mld;
  int c, half;
  float s0, s1;
  static constexpr bool kStage = true;
  static constexpr bool kHalf = true;
  __device__ __forceinline__ void init_half(float*, int ch) { c = ch & 63; half = (ch >> 6) & 1; s0 = s1 = 0.f; }
  __device__ __forceinline__ void block16(float (&v)[16], int, int j, bool valid, uint32_t stg) {
    if (!valid) return;
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const float t[8] = {v[8 * q], v[8 * q + 1], v[8 * q + 2], v[8 * q + 3], v[8 * q + 4], v[8 * q + 5], v[8 * q + 6], v[8 * q + 7]};
      tc::sts128(stg + tb_stage_off(c, j * 4 + half * 2 + q), tc::pack8_bf16(t));
    }
    float a[4] = {0.f, 0.f, 0.f, 0.f}, b[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < 16; ++i) { a[i & 3] += v[i]; b[i & 3] = fmaf(v[i], v[i], b[i & 3]); }
    s0 += (a[0] + a[1]) + (a[2] + a[3]);
    s1 += (b[0] + b[1]) + (b[2] + b[3]);
  }
  __host__ __device__ __forceinline__ int nconst() const { return 0; }
  __host__ __device__ __forceinline__ int stage_bytes() const { return C * 256; }
  __device__ __forceinline__ char* tile_dst(int tile) const { return y ? reinterpret_cast<char*>(y) + (size_t)tile * C * 256 : nullptr; }
  __device__ __forceinline__ void init(float*, int ch) { c = ch; s0 = s1 = 0.f; }
  __device__ __forceinline__ void block(float (&v)[32], int, int j, bool valid, uint32_t stg) {
    if (c >= C || !valid) return;
#ifndef PCOE_EXP_NOSTAGE
    stage32_bf16(stg, c, j, v);
#endif
#ifndef PCOE_EXP_NOSUMS
    sum_sumsq32(v, s0, s1);
#endif
  }
  __device__ __forceinline__ void finish() {
    if (!sums || c >= C) return;
    double* dst = sums + (size_t)(blockIdx.x % kRedCopies) * 2 * C;
    atomicAdd(dst + c, (double)s0);
    atomicAdd(dst + C + c, (double)s1);
  }
};

struct Group4 {   // last layer, K == 32: the 32 columns of a block are one group
  __nv_bfloat16* __restrict__ y;   // tile-blocked, or nullptr (eval)
  double* __restrict__ sums;       // or nullptr
  float* __restrict__ ymax;        // [G,C]
  float* __restrict__ ymin;
  uint8_t* __restrict__ amax;
  uint8_t* __restrict__ amin;
  int C, Mld;
  const float* __restrict__ gamma;   // BatchNorm weight of this layer: the affine's sign is its sign, so each channel
                                     // needs only ONE of (max, arg-max) / (min, arg-min) - the other pair is not computed
  int c;
  float s0, s1;
  bool want_max;
  static constexpr bool kStage = true;
  static constexpr bool kHalf = false;
  __host__ __device__ __forceinline__ int nconst() const { return 0; }
  __host__ __device__ __forceinline__ int stage_bytes() const { return y ? C * 256 : 0; }
  __device__ __forceinline__ char* tile_dst(int tile) const { return y ? reinterpret_cast<char*>(y) + (size_t)tile * C * 256 : nullptr; }
  __device__ __forceinline__ void init(float*, int ch) { c = ch; s0 = s1 = 0.f; want_max = c < C ? !signbit(gamma[c]) : true; }
  __device__ __forceinline__ void block(float (&v)[32], int tile, int j, bool valid, uint32_t stg) {
    if (c >= C || !valid) return;
    if (y) stage32_bf16(stg, c, j, v);
#ifndef PCOE_EXP_NOSUMS
    sum_sumsq32(v, s0, s1);
#endif
    float mx, mn;
    int ax, an;
    const size_t o = (size_t)(tile * 4 + j) * C + c;
#ifndef PCOE_EXP_NOARG
    if (want_max) { argext32<true>(v, mx, ax); ymax[o] = mx; amax[o] = (uint8_t)ax; }
    else { argext32<false>(v, mn, an); ymin[o] = mn; amin[o] = (uint8_t)an; }
#else
    mx = v[0]; mn = v[1]; ax = 0; an = 1;
    ymax[o] = mx; ymin[o] = mn; amax[o] = (uint8_t)ax; amin[o] = (uint8_t)an;
#endif
  }
  __device__ __forceinline__ void finish() {
    if (!sums || c >= C) return;
    double* dst = sums + (size_t)(blockIdx.x % kRedCopies) * 2 * C;
    atomicAdd(dst + c, (double)s0);
    atomicAdd(dst + C + c, (double)s1);
  }
};

// dz_prev^T = dx^T * [z_prev > 0]; sums of dz_prev and dz_prev * xhat_prev per channel
struct MaskStats4 {
  const __nv_bfloat16* __restrict__ yprev;   // tile-blocked
  const float* __restrict__ scale;
  const float* __restrict__ shift;
  const float* __restrict__ mean;
  const float* __restrict__ invstd;
  __nv_bfloat16* __restrict__ dz;            // tile-blocked
  double* __restrict__ sums;
  int C, Mld;
  const float* __restrict__ addc;            // optional per-channel constant added to dx before the mask (DySparse4: W^T q)
  int c;
  float s0, s1, sc, sh, is, nmi, ac;   // nmi = -mean * invstd: xhat = y * invstd + nmi
  int half;
  static constexpr bool kStage = true;
  static constexpr bool kPre = false;
  static constexpr bool kHalf = true;
  __device__ __forceinline__ void init_half(float* csm, int ch) { init(csm, ch & 63); half = (ch >> 6) & 1; }
  __device__ __forceinline__ void block16(float (&v)[16], int tile, int j, bool valid, uint32_t stg) {
    if (!valid) return;
    uint4 raw[2];
#pragma unroll
    for (int q = 0; q < 2; ++q) raw[q] = __ldg(tb_chunk(yprev, C, tile, c, j * 4 + half * 2 + q));
    float a[4] = {0.f, 0.f, 0.f, 0.f}, b[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      float yy[8], t[8];
      unpack8(raw[q], yy);
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const bool on = fmaf(yy[u], sc, sh) > 0.f;
        const float d = on ? v[8 * q + u] + ac : 0.f;
        t[u] = d;
        a[u & 3] += d;
        b[u & 3] = fmaf(d, fmaf(yy[u], is, nmi), b[u & 3]);
      }
      tc::sts128(stg + tb_stage_off(c, j * 4 + half * 2 + q), tc::pack8_bf16(t));
    }
    s0 += (a[0] + a[1]) + (a[2] + a[3]);
    s1 += (b[0] + b[1]) + (b[2] + b[3]);
  }
  __host__ __device__ __forceinline__ int nconst() const { return 0; }
  __host__ __device__ __forceinline__ int stage_bytes() const { return C * 256; }
  __device__ __forceinline__ char* tile_dst(int tile) const { return reinterpret_cast<char*>(dz) + (size_t)tile * C * 256; }
  __device__ __forceinline__ void init(float*, int ch) {
    c = ch; s0 = s1 = 0.f;
    const bool ok = c < C;
    sc = ok ? scale[c] : 0.f; sh = ok ? shift[c] : 0.f; is = ok ? invstd[c] : 0.f; nmi = ok ? -mean[c] * is : 0.f;
    ac = (ok && addc) ? addc[c] : 0.f;
  }
  __device__ __forceinline__ void block(float (&v)[32], int tile, int j, bool valid, uint32_t stg) {
    if (c >= C || !valid) return;
    uint4 raw[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) raw[q] = __ldg(tb_chunk(yprev, C, tile, c, j * 4 + q));
    float a[4] = {0.f, 0.f, 0.f, 0.f}, b[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float yy[8];
      unpack8(raw[q], yy);
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const bool on = fmaf(yy[u], sc, sh) > 0.f;
        const float d = on ? v[8 * q + u] + ac : 0.f;
        v[8 * q + u] = d;
        a[u & 3] += d;
        b[u & 3] = fmaf(d, fmaf(yy[u], is, nmi), b[u & 3]);
      }
    }
    stage32_bf16(stg, c, j, v);
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

// SA1 (no input features, C1 <= 128): layer 1 has no data gradient, only dW1 [C1 x 3] = dy1^T x0 with
//     dy1 = a1*dz1 + p1*y1 + q1,  y1 = W1 x0   =>   dW1 = a1 .* (dz1^T x0) + p1 .* (W1 (x0^T x0)) + q1 (sum x0)^T .
// The layer-2 backward epilogue therefore accumulates A = dz1^T x0 itself (thread = channel: 3 floats; the 32 point
// offsets of a block are exchanged with shuffles), one warp per block adds x0^T x0 and sum x0, and dw_combine
// applies (a1, p1, q1) once the batch sums are complete.  dz1 is never stored and the layer-1 backward kernel
// (35 us: a full pass over dz1 and y1 for a 64 x 3 result) does not run.  x0 is rounded to bf16 exactly as the
// forward's MMA operand was.  The gather of x0 (nbr -> xyz, two dependent loads) is issued by pre() BEFORE the wait
// for the accumulator, so its latency hides behind the MMAs.
struct MaskStatsW1 {
  const __nv_bfloat16* __restrict__ yprev;   // tile-blocked y1
  const float* __restrict__ scale;
  const float* __restrict__ shift;
  const float* __restrict__ mean;
  const float* __restrict__ invstd;
  double* __restrict__ sums;
  int C, Mld;
  GatherBase gb;
  float* __restrict__ acc;                   // [kRedCopies][C][4]: A in columns 0..2
  float* __restrict__ g0;                    // [kRedCopies][16]: xx xy xz yy yz zz sx sy sz
  int c, eq;
  float s0, s1, sc, sh, is, nmi;
  float A0, A1, A2;
  float G[9];
  float4* xs;                                // shared memory: [warp][2 blocks][32 points] (x, y, z, 0), written by pre()
  static constexpr bool kStage = false;
  static constexpr bool kPre = true;
  static constexpr bool kHalf = true;
  int half;
  __host__ __device__ __forceinline__ int nconst() const { return 8 * 2 * 32 * 4; }
  __device__ __forceinline__ void init_half(float* csm, int ch) {   // C == 64: channel ch & 63, points half*16..+15 of a block
    init(csm, ch & 63);
    eq = (ch >> 5) & 3;
    half = (ch >> 6) & 1;
  }
  __device__ __forceinline__ void block16(float (&v)[16], int tile, int j, bool valid, uint32_t) {
    const int it = j & 1;
    const float4* xp = xs + it * 32;
    if (valid) {
      uint4 raw[2];
#pragma unroll
      for (int q = 0; q < 2; ++q) raw[q] = __ldg(tb_chunk(yprev, C, tile, c, j * 4 + half * 2 + q));
      float a[4] = {0.f, 0.f, 0.f, 0.f}, b[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        float yy[8];
        unpack8(raw[q], yy);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const bool on = fmaf(yy[u], sc, sh) > 0.f;
          const float d = on ? v[8 * q + u] : 0.f;
          v[8 * q + u] = d;
          a[u & 3] += d;
          b[u & 3] = fmaf(d, fmaf(yy[u], is, nmi), b[u & 3]);
        }
      }
      s0 += (a[0] + a[1]) + (a[2] + a[3]);
      s1 += (b[0] + b[1]) + (b[2] + b[3]);
      float t0[4] = {0.f, 0.f, 0.f, 0.f}, t1[4] = {0.f, 0.f, 0.f, 0.f}, t2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float4 o = xp[half * 16 + i];
        t0[i & 3] = fmaf(v[i], o.x, t0[i & 3]); t1[i & 3] = fmaf(v[i], o.y, t1[i & 3]); t2[i & 3] = fmaf(v[i], o.z, t2[i & 3]);
      }
      A0 += (t0[0] + t0[1]) + (t0[2] + t0[3]); A1 += (t1[0] + t1[1]) + (t1[2] + t1[3]); A2 += (t2[0] + t2[1]) + (t2[2] + t2[3]);
      if (eq == 3) {   // one warp per block: x0^T x0 and sum x0 of this lane's point
        const float4 own = xp[threadIdx.x & 31];
        G[0] = fmaf(own.x, own.x, G[0]); G[1] = fmaf(own.x, own.y, G[1]); G[2] = fmaf(own.x, own.z, G[2]);
        G[3] = fmaf(own.y, own.y, G[3]); G[4] = fmaf(own.y, own.z, G[4]); G[5] = fmaf(own.z, own.z, G[5]);
        G[6] += own.x; G[7] += own.y; G[8] += own.z;
      }
    }
  }
  __host__ __device__ __forceinline__ int stage_bytes() const { return 0; }
  __device__ __forceinline__ char* tile_dst(int) const { return nullptr; }
  __device__ __forceinline__ void init(float* csm, int ch) {
    xs = reinterpret_cast<float4*>(csm) + (threadIdx.x >> 5) * 64;   // csm is 16-byte aligned
    c = ch; eq = (ch >> 5) & 3; s0 = s1 = 0.f; A0 = A1 = A2 = 0.f;
#pragma unroll
    for (int i = 0; i < 9; ++i) G[i] = 0.f;
    const bool ok = c < C;
    sc = ok ? scale[c] : 0.f; sh = ok ? shift[c] : 0.f; is = ok ? invstd[c] : 0.f; nmi = ok ? -mean[c] * is : 0.f;
  }
  // Gather of the tile's point offsets, issued BEFORE the wait for the accumulator.  (A version that fetched the
  // neighbour indices two tiles and the coordinates one tile ahead, double-buffered, measured no faster: the cost of
  // this epilogue is its arithmetic, not the gather latency.)
  __device__ __forceinline__ void pre(int tile, int eh, int M) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int it = 0; it < 2; ++it) {
      const int row = tile * kPts + (2 * eh + it) * 32 + lane;
      float o[3] = {0.f, 0.f, 0.f};
      if (row < M) {
        float x[3], cc[3];
        gb.load_xyz_raw(row, gb.point_of(row), x, cc);
#pragma unroll
        for (int u = 0; u < 3; ++u) o[u] = __bfloat162float(__float2bfloat16_rn(gb.centred(x[u], cc[u])));
      }
      xs[it * 32 + lane] = make_float4(o[0], o[1], o[2], 0.f);   // read back by this warp only (block())
    }
    __syncwarp();
  }
  __device__ __forceinline__ void block(float (&v)[32], int tile, int j, bool valid, uint32_t) {
    const int it = j & 1;
    const bool act = c < C && valid;
    if (act) {
      uint4 raw[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) raw[q] = __ldg(tb_chunk(yprev, C, tile, c, j * 4 + q));
      float a[4] = {0.f, 0.f, 0.f, 0.f}, b[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float yy[8];
        unpack8(raw[q], yy);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const bool on = fmaf(yy[u], sc, sh) > 0.f;
          const float d = on ? v[8 * q + u] : 0.f;
          v[8 * q + u] = d;
          a[u & 3] += d;
          b[u & 3] = fmaf(d, fmaf(yy[u], is, nmi), b[u & 3]);
        }
      }
      s0 += (a[0] + a[1]) + (a[2] + a[3]);
      s1 += (b[0] + b[1]) + (b[2] + b[3]);
    }
    // A += dz1[c, point i] * x0[point i, :]: the 32 offsets come from shared memory (one broadcast 16-byte load each)
    const float4* xp = xs + it * 32;
    const float4 own = xp[threadIdx.x & 31];
    const float qx = own.x, qy = own.y, qz = own.z;
    if (act) {
      float t0[4] = {0.f, 0.f, 0.f, 0.f}, t1[4] = {0.f, 0.f, 0.f, 0.f}, t2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const float4 o = xp[i];
        t0[i & 3] = fmaf(v[i], o.x, t0[i & 3]); t1[i & 3] = fmaf(v[i], o.y, t1[i & 3]); t2[i & 3] = fmaf(v[i], o.z, t2[i & 3]);
      }
      A0 += (t0[0] + t0[1]) + (t0[2] + t0[3]); A1 += (t1[0] + t1[1]) + (t1[2] + t1[3]); A2 += (t2[0] + t2[1]) + (t2[2] + t2[3]);
    }
    if (eq == 3 && valid) {   // one warp per block: x0^T x0 and sum x0 of this lane's point (zero beyond M)
      G[0] = fmaf(qx, qx, G[0]); G[1] = fmaf(qx, qy, G[1]); G[2] = fmaf(qx, qz, G[2]);
      G[3] = fmaf(qy, qy, G[3]); G[4] = fmaf(qy, qz, G[4]); G[5] = fmaf(qz, qz, G[5]);
      G[6] += qx; G[7] += qy; G[8] += qz;
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

// layer-1 dgrad over the D feature columns, POINT-on-lane orientation: thread = point, v = 32 feature channels
struct Scatter4 {
  float* __restrict__ grad_feats;
  const int32_t* __restrict__ nbr;
  int N, S, D, group_all;
  static constexpr bool kStage = false;
  static constexpr bool kPre = false;
  static constexpr bool kHalf = false;
  __host__ __device__ __forceinline__ int nconst() const { return 0; }
  __host__ __device__ __forceinline__ int stage_bytes() const { return 0; }
  __device__ __forceinline__ char* tile_dst(int) const { return nullptr; }
  __device__ __forceinline__ void init(float*, int) {}
  // cb = first feature channel of the block, row = the thread's point row
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
  __device__ __forceinline__ void finish() {}
};

struct NoEpi4 {
  static constexpr bool kStage = false;
  static constexpr bool kPre = false;
  static constexpr bool kHalf = false;
  __host__ __device__ __forceinline__ int nconst() const { return 0; }
  __host__ __device__ __forceinline__ int stage_bytes() const { return 0; }
  __device__ __forceinline__ char* tile_dst(int) const { return nullptr; }
  __device__ __forceinline__ void init(float*, int) {}
  __device__ __forceinline__ void finish() {}
};

struct Barriers4 {
  uint64_t full[kMaxStages];
  uint64_t empty[kMaxStages];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
};

// weight image: bf16 [Rp][Kp] row-major (zero padded) -> [Rp rows][Kp/64 blocks], SWIZZLE_128B
// Split in two so that the kernel prologue can keep the global loads in flight while it does other work (clearing
// the operand stages, the per-channel constants with their own L2 round trips): wimage_issue() loads the first
// kWImgRegs 16-byte units of this thread into registers, wimage_store() stores them swizzled and loops over the rest.
constexpr int kWImgRegs = 8;
struct WImgRegs { uint4 w[kWImgRegs]; };
__device__ __forceinline__ void wimage_issue(const __nv_bfloat16* __restrict__ W, int Rp, int Kp, int tid, int nthr, WImgRegs& r) {
  const int total = Rp * (Kp / 8);            // the image is contiguous: unit e is the e-th uint4
  const uint4* src = reinterpret_cast<const uint4*>(W);
#pragma unroll
  for (int u = 0; u < kWImgRegs; ++u) {
    const int e = tid + u * nthr;
    r.w[u] = e < total ? __ldg(src + e) : make_uint4(0, 0, 0, 0);
  }
}
__device__ __forceinline__ void wimage_store(const __nv_bfloat16* __restrict__ W, int Rp, int Kp, uint32_t saddr, int tid, int nthr,
                                             const WImgRegs& r) {
  const int upr = Kp / 8, total = Rp * upr;
  const uint4* src = reinterpret_cast<const uint4*>(W);
  auto put = [&](int e, const uint4& w) {
    const int n = e / upr, j = e - n * upr;
    tc::sts128(saddr + (uint32_t)(j >> 3) * (uint32_t)(Rp * 128) + tc::sw128_off(n, (j & 7) * 8), w);
  };
#pragma unroll
  for (int u = 0; u < kWImgRegs; ++u) {
    const int e = tid + u * nthr;
    if (e < total) put(e, r.w[u]);
  }
  for (int e = tid + kWImgRegs * nthr; e < total; e += nthr) put(e, __ldg(src + e));
}
__device__ __forceinline__ void load_wimage(const __nv_bfloat16* __restrict__ W, int Rp, int Kp, uint32_t saddr, int tid, int nthr) {
  WImgRegs r;
  wimage_issue(W, Rp, Kp, tid, nthr, r);
  wimage_store(W, Rp, Kp, saddr, tid, nthr, r);
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

// Copy-out of a staged output tile, executed by all 256 epilogue threads after their last block():
// one elected thread issues the bulk async copy.  nstg == 2: the staging buffer alternates and the
// elected thread only waits (before the barrier) for the copy of the PREVIOUS tile to have read its
// buffer; nstg == 1: everybody waits for this tile's copy before the buffer is reused.
__device__ __forceinline__ void stage_copy_out(char* dst, uint32_t stg, int bytes, int nstg, bool elected) {
  if (elected) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  tc::fence_proxy_async();
  asm volatile("bar.sync 1, 256;" ::: "memory");
  if (elected) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(stg), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    if (nstg == 1) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  }
  if (nstg == 1) asm volatile("bar.sync 2, 256;" ::: "memory");
}

// ---------------------------------------------------------------------------------------------
// forward layer:  Y^T[C_out x 128] = W * X^T per tile;  TMEM: nbuf buffers x mt x 128 columns
// smem: [W image][stages x tile][constants]
// GRAM: additionally accumulate  Gram[C_in x (C_in+16)] += X^T_ext X_ext  over the CTA's tiles in TMEM columns
// [nbuf*mt*128, +C_in+16) (X_ext = the input tile with a row of ones appended as row C_in, so column C_in of the
// result is sum_points x): the statistics the last layer's backward needs instead of the saved y3 (DySparse4).
// Requires M % 128 == 0 and prod.rows() >= max(128, C_in + 16); flushed into gram[blockIdx.x % kRedCopies].
// ---------------------------------------------------------------------------------------------
template <class Prod, class Epi, int TCOLS, int GRAM, bool HALF>
__global__ void __launch_bounds__(kThreads, 1)
tc4_fwd_kernel(Prod prod, const __nv_bfloat16* __restrict__ Wb, int Rp, int Kp, Epi epi, int M, int nstages, int nstg,
               float* __restrict__ gram, int nbuf) {
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

  TC4_TRACER;
  if (tid == 0) TC4_TRACE(0, 0);
  if (warp == 0) tc::tmem_alloc<TCOLS>(&tmem_base);
  if (tid == 0) {
    for (int s = 0; s < nstages; ++s) { tc::mbar_init(&bar.full[s], kProdThreads); tc::mbar_init(&bar.empty[s], 1); }
    for (int b = 0; b < 2; ++b) { tc::mbar_init(&bar.tmem_full[b], 1); tc::mbar_init(&bar.tmem_empty[b], kEpiThreads); }
  }
  WImgRegs wreg;
  wimage_issue(Wb, Rp, Kp, tid, kThreads, wreg);            // in flight during the clear and the constants below
  zero_smem(sT, (uint32_t)nstages * tbytes, tid, kThreads);
  if constexpr (GRAM) {   // row kext() of every stage image = ones (never touched by the producers)
    __syncthreads();
    const int orow = prod.kext();
    for (int e = tid; e < nstages * 16; e += kThreads) {
      const uint32_t a = sT + (uint32_t)(e >> 4) * tbytes + (uint32_t)((e >> 3) & 1) * (uint32_t)(prod.rows() * 128) +
                         (uint32_t)((orow >> 3) * 1024 + (orow & 7) * 128 + (e & 7) * 16);
      tc::sts128(a, make_uint4(0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u));
    }
  }
  prod.init(csm, tid, kThreads);
  const uint32_t gcol = (uint32_t)(nbuf * mt * kPts);   // Gram accumulator columns
  // epilogue work split: warp w -> TMEM lane quadrant w&3, items [eh*mt*2, (eh+1)*mt*2) of the mt*4
  // (M tile, 32-column block) pairs of a tile; all of one thread's items share one M tile
  const int eq = warp & 3, eh = (warp >> 2) & 1, lane = tid & 31;
  const int emi = (eh * mt) >> 1;
  // staging tiles (tile-blocked output image) follow the constants
  const uint32_t stg0 = (sT + (uint32_t)nstages * tbytes + (uint32_t)(prod.nconst() + epi.nconst()) * 4u + 127u) & ~127u;
  const uint32_t stg_bytes = (uint32_t)epi.stage_bytes();
  if constexpr (HALF) epi.init_half(csm + prod.nconst(), emi * 128 + eq * 32 + lane);
  else epi.init(csm + prod.nconst(), emi * 128 + eq * 32 + lane);
  wimage_store(Wb, Rp, Kp, sW, tid, kThreads, wreg);
  tc::fence_proxy_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = tmem_base;
  const int ntiles = (M + kPts - 1) / kPts;
  if (tid == 0) TC4_TRACE(1, 0);

  if (warp < 8) {
    // ---- epilogue ----
    int i = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++i) {
      const int b = nbuf == 2 ? (i & 1) : 0, u = nbuf == 2 ? (i >> 1) : i, m0 = tile * kPts;
      tc::mbar_wait(&bar.tmem_full[b], (uint32_t)(u & 1));
      if (tid == 0) TC4_TRACE(30, i);
      tc::fence_after_sync();
      const uint32_t stg = stg0 + (uint32_t)(i % nstg) * stg_bytes;
      const int it0 = eh * mt * 2;
#pragma unroll 1
      for (int it = it0; it < it0 + mt * 2; ++it) {
        const int j = it & 3;
        if constexpr (HALF) {   // lanes 64..127 repeat channels 0..63: this thread takes 16 of the block's 32 points
          float v[16];
          tc::tmem_ld16(tmem + ((uint32_t)(eq * 32) << 16) + (uint32_t)(b * mt * kPts + emi * kPts + j * 32 + (eq >> 1) * 16), v);
          epi.block16(v, tile, j, m0 + j * 32 < M, stg);
        } else {
          float v[32];
          tc::tmem_ld32(tmem + ((uint32_t)(eq * 32) << 16) + (uint32_t)(b * mt * kPts + emi * kPts + j * 32), v);
          if (tid == 0) TC4_TRACE(33, it);
          epi.block(v, tile, j, m0 + j * 32 < M, stg);
          if (tid == 0) TC4_TRACE(34, it);
        }
      }
      tc::fence_before_sync();
      mbar_arrive_relaxed(&bar.tmem_empty[b]);
      if (tid == 0) TC4_TRACE(35, i);
      if (Epi::kStage && epi.tile_dst(0) != nullptr) stage_copy_out(epi.tile_dst(tile), stg, (int)stg_bytes, nstg, tid == 0);
      if (tid == 0) TC4_TRACE(31, i);
    }
    if (Epi::kStage && tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    epi.finish();
    if (tid == 0) TC4_TRACE(32, 0);
    if constexpr (GRAM) {
      // flush the Gram accumulator (complete: the last tmem_full commit covers every earlier MMA): thread = row
      if ((int)blockIdx.x < ntiles) {
        tc::fence_after_sync();
        const int cin = prod.kext(), gn = cin + 16, nblk = (gn + 31) / 32;
        float* base = gram + (size_t)(blockIdx.x % kRedCopies) * cin * gn;
        const int row = eq * 32 + lane;
#pragma unroll 1
        for (int k = eh; k < nblk; k += 2) {
          const int cb = ((k + (int)blockIdx.x) % nblk) * 32;
          float v[32];
          tc::tmem_ld32(tmem + ((uint32_t)(eq * 32) << 16) + gcol + (uint32_t)cb, v);   // may read past gn: unused columns
          if (row < cin) {
            float* dst = base + (size_t)row * gn + cb;
#pragma unroll
            for (int q = 0; q < 8; ++q)
              if (cb + 4 * q < gn) red_add_v4(dst + 4 * q, v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
          }
        }
      }
    }
  } else if (warp < 16) {
    // ---- producers ----
    int ring_s = 0, ring_n = 0;   // stage slot / round of the tile being stored (tiles are stored in order)
    producer_pipeline(prod, tid - kEpiThreads, kProdThreads, ntiles,
                      [&](int t) { if (t >= nstages) tc::mbar_wait(&bar.empty[ring_s], (uint32_t)((ring_n - 1) & 1)); },
                      [&](int) { return sT + (uint32_t)ring_s * tbytes; },
                      [&](int) { mbar_arrive_full(&bar.full[ring_s]); if (++ring_s == nstages) { ring_s = 0; ++ring_n; } });
  } else {   // warp 16: MMA issue, warp-uniform loop, one elected lane issues
    const uint32_t tmem = tc::uniform_u32(tmem_base);
    // ---- MMA issue ----
    const uint32_t idesc = tc::make_idesc_bf16(128, kPts, false, Prod::kChMajor);
    const int kext = prod.kext();
    const uint32_t idesc_g = tc::make_idesc_bf16(128, kext + 16, false, false);
    const uint32_t prow = (uint32_t)prod.rows();
    int i = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++i) {
      const int s = i % nstages, n = i / nstages, b = nbuf == 2 ? (i & 1) : 0, u = nbuf == 2 ? (i >> 1) : i;
      tc::mbar_wait(&bar.full[s], (uint32_t)(n & 1));
      if ((tid & 31) == 0) TC4_TRACE(20, i);
      if (u > 0) tc::mbar_wait(&bar.tmem_empty[b], (uint32_t)((u - 1) & 1));
      tc::fence_after_sync();
      if ((tid & 31) == 0) TC4_TRACE(21, i);
      const uint32_t st = sT + (uint32_t)s * tbytes;
      for (int mi = 0; mi < mt; ++mi)
        for (int k = 0; k < kext; k += 16)
          tc::mma_bf16_warp(tmem + (uint32_t)(b * mt * kPts + mi * kPts),
                       tc::make_desc_sw128(sW + (uint32_t)(k >> 6) * (uint32_t)(Rp * 128) + (uint32_t)(mi * 128 * 128) + (uint32_t)((k & 63) * 2), 16, 1024),
                       act_desc_chan_k(prod, st, k), idesc, k > 0);
      if constexpr (GRAM) {   // Gram += X_ext^T X_ext: contraction over the 128 points, A = rows 0..127, B = rows 0..kext+15
        for (int ks = 0; ks < 8; ++ks) {
          const uint64_t xd = tc::make_desc_sw128(st + (uint32_t)(ks >> 2) * (prow * 128u) + (uint32_t)((ks & 3) * 32), 16, 1024);
          tc::mma_bf16_warp(tmem + gcol, xd, xd, idesc_g, i > 0 || ks > 0);
        }
      }
      tc::mma_commit_warp(&bar.empty[s]);
      tc::mma_commit_warp(&bar.tmem_full[b]);
      if ((tid & 31) == 0) TC4_TRACE(22, i);
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (tid == 0) TC4_TRACE(2, 0);
  if (tid == 0 || tid == kEpiThreads + kProdThreads) TC4_TRACE_DUMP();
  if (warp == 0) tc::tmem_dealloc<TCOLS>(tmem);
}

// ---------------------------------------------------------------------------------------------
// fused backward of one layer.  P = dy^T tile [C_l x 128] (channel-major), Q = x_prev tile.
//   dW[C_l x Cq] += P * Q^T            TMEM columns [0, mtl*nw), accumulated over the CTA's tiles
//   DGRAD == 1:  dx^T[C_prev x 128] = W^T * P    (thread = channel epilogue; 2 buffers of mtp*128 cols)
//   DGRAD == 2:  dx[128 x D]        = P^T * W     (thread = point epilogue: layer-1 scatter; 2 buffers of D cols)
//   GM (DGRAD == 1, Q channel-major, C_prev <= 128):  dx^T += Gm * Q  with Gm a [128 x gk] bf16 image (gk = C_prev):
//                the dense part of the last layer's BatchNorm backward folded into a matrix (see DySparse4)
// smem: [W image (DGRAD)][Gm image (GM)][P tile][Q tile][constants]   (single stage)
// ---------------------------------------------------------------------------------------------
template <class PProd, class QProd, class Epi, int DGRAD, int TCOLS, int GM, bool HALF>
__global__ void __launch_bounds__(kThreads, 1)
tc4_bwd_kernel(PProd pp, QProd qp, const __nv_bfloat16* __restrict__ Wb, int Rp, int Kp, Epi epi,
               float* __restrict__ dW, int ldo, int cq_valid, int perm_d /* >=0: layer-1 [feats|xyz] column order */,
               int M, int cprev /* dgrad output channels */, int nstg, int npq /* P/Q operand stages: 1 or 2 */,
               const __nv_bfloat16* __restrict__ Gmb, int gk) {
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
  const uint32_t gbytes = GM ? 128u * (uint32_t)gk * 2u : 0u;
  const uint32_t pbytes = (prod_tile_bytes(pp) + 1023u) & ~1023u, qbytes = (prod_tile_bytes(qp) + 1023u) & ~1023u;
  const uint32_t sW = smem0, sG = smem0 + wbytes, sP0 = sG + gbytes, pqbytes = pbytes + qbytes;   // stage s: P at sP0 + s*pqbytes, Q behind it
  float* csm = reinterpret_cast<float*>(smem_gen + wbytes + gbytes + (size_t)npq * pqbytes);
  const uint32_t dx_col = (uint32_t)((mtl * nw + 31) / 32 * 32);
  const uint32_t dx_cols = DGRAD == 1 ? (uint32_t)(mtp * kPts) : (uint32_t)((cprev + 31) / 32 * 32);

  if (warp == 0) tc::tmem_alloc<TCOLS>(&tmem_base);
  if (tid == 0) {
    for (int s = 0; s < npq; ++s) { tc::mbar_init(&bar.full[s], kProdThreads); tc::mbar_init(&bar.empty[s], 1); }
    for (int b = 0; b < 2; ++b) { tc::mbar_init(&bar.tmem_full[b], 1); tc::mbar_init(&bar.tmem_empty[b], kEpiThreads); }
  }
  WImgRegs wreg, greg;
  if (DGRAD) wimage_issue(Wb, Rp, Kp, tid, kThreads, wreg);   // in flight during the clear and the constants below
  if (GM) wimage_issue(Gmb, 128, gk, tid, kThreads, greg);
  zero_smem(sP0, (uint32_t)npq * pqbytes, tid, kThreads);
  pp.init(csm, tid, kThreads);
  qp.init(csm + pp.nconst(), tid, kThreads);
  if (DGRAD) wimage_store(Wb, Rp, Kp, sW, tid, kThreads, wreg);
  if (GM) wimage_store(Gmb, 128, gk, sG, tid, kThreads, greg);
  // epilogue split as in the forward kernel (DGRAD == 1: mtp*4 items; DGRAD == 2: column blocks)
  const int eq = warp & 3, eh = (warp >> 2) & 1, lane = tid & 31;
  const int emi = (eh * mtp) >> 1;
  const uint32_t stg0 = (sP0 + (uint32_t)npq * pqbytes + (uint32_t)(pp.nconst() + qp.nconst() + epi.nconst()) * 4u + 127u) & ~127u;
  const uint32_t stg_bytes = (uint32_t)epi.stage_bytes();
  if constexpr (HALF) epi.init_half(csm + pp.nconst() + qp.nconst(), emi * 128 + eq * 32 + lane);
  else epi.init(csm + pp.nconst() + qp.nconst(), emi * 128 + eq * 32 + lane);
  tc::fence_proxy_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = tmem_base;
  const int ntiles = (M + kPts - 1) / kPts;
  int my_tiles = 0;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) ++my_tiles;

  TC4_TRACER;
  if (tid == 0) { TC4_TRACE(0, 0); TC4_TRACE(1, 0); }
  if (warp < 8) {
    int i = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++i) {
      const int b = i & 1, u = i >> 1, m0 = tile * kPts;
      if constexpr (Epi::kPre) epi.pre(tile, eh, M);   // loads whose latency hides behind the wait below
      tc::mbar_wait(&bar.tmem_full[b], (uint32_t)(u & 1));
      if (tid == 0) TC4_TRACE(30, i);
      tc::fence_after_sync();
      const uint32_t stg = stg0 + (uint32_t)(i % nstg) * stg_bytes;
      if constexpr (DGRAD == 1) {
        const int it0 = eh * mtp * 2;
#pragma unroll 1
        for (int it = it0; it < it0 + mtp * 2; ++it) {
          const int j = it & 3;
          if constexpr (HALF) {
            float v[16];
            tc::tmem_ld16(tmem + ((uint32_t)(eq * 32) << 16) + dx_col + (uint32_t)b * dx_cols + (uint32_t)(emi * kPts + j * 32 + (eq >> 1) * 16), v);
            epi.block16(v, tile, j, m0 + j * 32 < M, stg);
          } else {
            float v[32];
            tc::tmem_ld32(tmem + ((uint32_t)(eq * 32) << 16) + dx_col + (uint32_t)b * dx_cols + (uint32_t)(emi * kPts + j * 32), v);
            epi.block(v, tile, j, m0 + j * 32 < M, stg);
          }
        }
      } else if constexpr (DGRAD == 2) {
        const int row = m0 + eq * 32 + lane;
#pragma unroll 1
        for (int cb = eh * 32; cb < cprev; cb += 64) {
          float v[32];
          tc::tmem_ld32(tmem + ((uint32_t)(eq * 32) << 16) + dx_col + (uint32_t)b * dx_cols + (uint32_t)cb, v);
          epi.block_pt(v, cb, row, row < M);
        }
      }
      tc::fence_before_sync();
      mbar_arrive_relaxed(&bar.tmem_empty[b]);
      if (tid == 0) TC4_TRACE(35, i);
      if constexpr (DGRAD == 1 && Epi::kStage) stage_copy_out(epi.tile_dst(tile), stg, (int)stg_bytes, nstg, tid == 0);
      if (tid == 0) TC4_TRACE(31, i);
    }
    if (DGRAD == 1 && Epi::kStage && tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    epi.finish();
    if (tid == 0) TC4_TRACE(32, 0);
    // flush dW (complete: the last tmem_full commit covers every earlier MMA): thread = row of dW, into this
    // CTA's copy dWc[blockIdx.x % kRedCopies][C_l][ldo] (ldo = nw, a multiple of 16; padding columns hold the
    // zeros of the Q tile's unused channels).  pcoe_sa_backward adds the copies up (dw_combine_kernel).
#ifdef PCOE_EXP_NOFLUSH
    if (false) {
#else
    if (my_tiles > 0) {
#endif
      tc::fence_after_sync();
      if (tid == 0) TC4_TRACE(36, 0);
      const int nblk = (nw + 31) / 32;
      float* base = dW + (size_t)(blockIdx.x % kRedCopies) * cl * ldo;
      for (int mi = 0; mi < mtl; ++mi) {
        const int crow = mi * 128 + eq * 32 + lane;
#pragma unroll 1
        for (int k = eh; k < nblk; k += 2) {
          const int cb = ((k + (int)blockIdx.x) % nblk) * 32;
          float v[32];
          tc::tmem_ld32(tmem + ((uint32_t)(eq * 32) << 16) + (uint32_t)(mi * nw + cb), v);   // may read past nw: unused columns
          if (crow < cl) {
            float* dst = base + (size_t)crow * ldo + cb;
#pragma unroll
            for (int q = 0; q < 8; ++q)
              if (cb + 4 * q < nw) red_add_v4(dst + 4 * q, v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
          }
        }
      }
    }
  } else if (warp < 16) {
    // ---- producers: all 8 warps build P (dy^T) and then Q (x_prev) of a tile, batch by batch, as one
    // three-stage software pipeline (indices two units ahead, global loads one unit ahead); one shared stage ----
    const int g = tid - kEpiThreads;
    const int nbp = pp.nbatches(kProdThreads), nbq = qp.nbatches(kProdThreads), upt = nbp + nbq;
    const int W = my_tiles * upt;
    union RawU { typename PProd::Raw p; typename QProd::Raw q; __device__ RawU() {} };
    constexpr int D = LoadDepth<RawU>::value;
    typename QProd::Idx ix;   // the dy producers have no index stage
    struct Cur { int t, u, m0; };   // tile ordinal, unit within the tile, first row: advanced incrementally (no division)
    const int mstep = (int)gridDim.x * kPts;
    auto adv = [&](Cur& c) { if (++c.u == upt) { c.u = 0; ++c.t; c.m0 += mstep; } };
    Cur ci{0, 0, (int)blockIdx.x * kPts}, cl = ci, cst = ci;
    auto do_idx = [&]() {
      if (ci.u >= nbp) qp.load_idx(g, kProdThreads, ci.m0, ci.u - nbp, ix);
      adv(ci);
    };
    auto do_load = [&](RawU& r) {
      if (cl.u < nbp) pp.load(g, kProdThreads, cl.m0, cl.u, NoIdx{}, r.p);
      else qp.load(g, kProdThreads, cl.m0, cl.u - nbp, ix, r.q);
      adv(cl);
    };
    auto step = [&](int w, RawU& cur, RawU& nxt) {
      if (w + D < W) do_load(nxt);
      if (w + D + 1 < W) do_idx();
      if (g == 0) TC4_TRACE(12, w);
      const int s = cst.t & (npq - 1), n = npq == 2 ? cst.t >> 1 : cst.t;     // stage slot, round
      if (cst.u == 0 && n > 0) tc::mbar_wait(&bar.empty[s], (uint32_t)((n - 1) & 1));
      if (g == 0) TC4_TRACE(10, w);
      const uint32_t sP = sP0 + (uint32_t)s * pqbytes;
      if (cst.u < nbp) pp.store(g, kProdThreads, cst.m0, cst.u, cur.p, sP);
      else qp.store(g, kProdThreads, cst.m0, cst.u - nbp, cur.q, sP + pbytes);
      if (cst.u == upt - 1) { tc::fence_proxy_async(); mbar_arrive_full(&bar.full[s]); }
      if (g == 0) TC4_TRACE(11, w);
      adv(cst);
    };
    if (W > 0) {
      RawU r[D + 1];   // unit u lives in r[u % (D + 1)]
      do_idx();
#pragma unroll
      for (int u = 0; u < D; ++u) {
        if (u < W) do_load(r[u]);
        if (u + 1 < W) do_idx();
      }
      for (int w = 0; w < W; w += D + 1) {
#pragma unroll
        for (int u = 0; u <= D; ++u)
          if (w + u < W) step(w + u, r[u], r[(u + D) % (D + 1)]);
      }
    }
    if (g == 0) TC4_TRACE_DUMP();
  } else {   // warp 16: MMA issue, warp-uniform loop, one elected lane issues
    const uint32_t tmem = tc::uniform_u32(tmem_base);
    const uint32_t idesc_w = tc::make_idesc_bf16(128, nw, false, !QProd::kChMajor);
    const uint32_t idesc_d = DGRAD == 1 ? tc::make_idesc_bf16(128, kPts, true, true)
                                        : tc::make_idesc_bf16(128, (cprev + 15) / 16 * 16, true, true);
    const int prow = pp.rows();
    int i = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++i) {
      const int b = i & 1, u = i >> 1;
      const int s = i & (npq - 1), n = npq == 2 ? i >> 1 : i;
      const uint32_t sP = sP0 + (uint32_t)s * pqbytes, sQ = sP + pbytes;
      tc::mbar_wait(&bar.full[s], (uint32_t)(n & 1));
      if ((tid & 31) == 0) TC4_TRACE(20, i);
      if (DGRAD && u > 0) tc::mbar_wait(&bar.tmem_empty[b], (uint32_t)((u - 1) & 1));
      tc::fence_after_sync();
      if ((tid & 31) == 0) TC4_TRACE(21, i);
      // dW += P Q^T: contraction over the 128 points, 16 per MMA
      for (int mi = 0; mi < mtl; ++mi)
        for (int ks = 0; ks < 8; ++ks) {
          const uint64_t ad = tc::make_desc_sw128(sP + (uint32_t)(ks >> 2) * (uint32_t)(prow * 128) + (uint32_t)(mi * 128 * 128) + (uint32_t)((ks & 3) * 32), 16, 1024);
          const uint64_t bd = QProd::kChMajor
              ? tc::make_desc_sw128(sQ + (uint32_t)(ks >> 2) * (uint32_t)(qp.rows() * 128) + (uint32_t)((ks & 3) * 32), 16, 1024)
              : tc::make_desc_sw128(sQ + (uint32_t)ks * 2048, kPts * 128, 1024);
          tc::mma_bf16_warp(tmem + (uint32_t)(mi * nw), ad, bd, idesc_w, i > 0 || ks > 0);
        }
      if constexpr (DGRAD == 1) {
        for (int mj = 0; mj < mtp; ++mj)
          for (int k = 0; k < cl; k += 16)
            tc::mma_bf16_warp(tmem + dx_col + (uint32_t)b * dx_cols + (uint32_t)(mj * kPts),
                         tc::make_desc_sw128(sW + (uint32_t)(2 * mj) * (uint32_t)(Rp * 128) + (uint32_t)(k >> 4) * 2048, (uint32_t)Rp * 128, 1024),
                         tc::make_desc_sw128(sP + (uint32_t)(k >> 4) * 2048, (uint32_t)prow * 128, 1024), idesc_d, k > 0);
        if constexpr (GM) {   // dx^T += Gm * Q : A = Gm image K-major, B = the Q tile MN-major (N = points)
          const uint32_t idesc_g = tc::make_idesc_bf16(128, kPts, false, true);
          for (int k = 0; k < gk; k += 16)
            tc::mma_bf16_warp(tmem + dx_col + (uint32_t)b * dx_cols,
                         tc::make_desc_sw128(sG + (uint32_t)(k >> 6) * (128u * 128u) + (uint32_t)((k & 63) * 2), 16, 1024),
                         tc::make_desc_sw128(sQ + (uint32_t)(k >> 4) * 2048, (uint32_t)qp.rows() * 128, 1024), idesc_g, true);
        }
      } else if constexpr (DGRAD == 2) {
        for (int k = 0; k < cl; k += 16)
          tc::mma_bf16_warp(tmem + dx_col + (uint32_t)b * dx_cols,
                       tc::make_desc_sw128(sP + (uint32_t)(k >> 4) * 2048, (uint32_t)prow * 128, 1024),
                       tc::make_desc_sw128(sW + (uint32_t)(k >> 4) * 2048, (uint32_t)Rp * 128, 1024), idesc_d, k > 0);
      }
      tc::mma_commit_warp(&bar.empty[s]);
      tc::mma_commit_warp(&bar.tmem_full[b]);
      if ((tid & 31) == 0) TC4_TRACE(22, i);
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (tid == 0) TC4_TRACE(2, 0);
  if (tid == 0 || tid == kEpiThreads + kProdThreads) TC4_TRACE_DUMP();
  if (warp == 0) tc::tmem_dealloc<TCOLS>(tmem);
}

// dW_l (+)= sum over the kRedCopies copies the v4 backward kernels accumulated; layer 1 goes back from the
// [feats(perm_d) | xyz(3)] column order to the parameter's [xyz | feats]
struct DwComb { const float* copies; float* dW; int rows, cin, ld, perm_d; };
// last layer on the DySparse4 path: dW3 += E,  E[c][k] = p_c * sum_j W3[c][j] * Gram[j][k] + q_c * sum_points x2[k]
// (l3_prep_kernel), the dense part of the BatchNorm backward
struct DwL3 { const float* E; };
// layer 1 on the MaskStatsW1 path (SA1): dW1[c][k] = a_c * A[c][k] + p_c * sum_j W1[c][j] G0[j][k] + q_c * s[k]
struct DwL1 { const float* acc; const float* g0; const __nv_bfloat16* Wb; int kp; BnBwdFin fin; };
__global__ void dw_combine_kernel(DwComb a, DwComb b, DwComb c, int accumulate, DwL3 x3, DwL1 x1) {
  const DwComb* L[3] = {&a, &b, &c};
  const int n0 = a.rows * a.cin, n1 = b.rows * b.cin, n2 = c.rows * c.cin;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n0 + n1 + n2; e += gridDim.x * blockDim.x) {
    const int l = e < n0 ? 0 : (e < n0 + n1 ? 1 : 2);
    const DwComb& w = *L[l];
    const int ee = e - (l == 0 ? 0 : (l == 1 ? n0 : n0 + n1));
    const int r = ee / w.cin, k = ee - r * w.cin;
    int src = k;
    if (w.perm_d >= 0) src = k < 3 ? w.perm_d + k : k - 3;
    float s = 0.f;
    if (l == 0 && x1.acc) {          // cin == 3
      float A = 0.f, G[9];
#pragma unroll
      for (int i = 0; i < 9; ++i) G[i] = 0.f;
#pragma unroll
      for (int g = 0; g < kRedCopies; ++g) {
        A += x1.acc[((size_t)g * w.rows + r) * 4 + k];
#pragma unroll
        for (int i = 0; i < 9; ++i) G[i] += x1.g0[g * 16 + i];
      }
      float ca, cp, cq;
      x1.fin.eval(r, w.rows, k == 0, ca, cp, cq);     // the k == 0 thread of a channel also writes dgamma / dbeta
      const float w0 = __bfloat162float(x1.Wb[(size_t)r * x1.kp]), w1 = __bfloat162float(x1.Wb[(size_t)r * x1.kp + 1]),
                  w2 = __bfloat162float(x1.Wb[(size_t)r * x1.kp + 2]);
      // G0 rows: (xx xy xz), (xy yy yz), (xz yz zz)
      const float gk0 = k == 0 ? G[0] : (k == 1 ? G[1] : G[2]);
      const float gk1 = k == 0 ? G[1] : (k == 1 ? G[3] : G[4]);
      const float gk2 = k == 0 ? G[2] : (k == 1 ? G[4] : G[5]);
      s = ca * A + cp * (w0 * gk0 + w1 * gk1 + w2 * gk2) + cq * G[6 + k];
      w.dW[ee] = accumulate ? w.dW[ee] + s : s;
      continue;
    }
#pragma unroll
    for (int g = 0; g < kRedCopies; ++g) s += w.copies[((size_t)g * w.rows + r) * w.ld + src];
    if (l == 2 && x3.E) s += x3.E[ee];
    w.dW[ee] = accumulate ? w.dW[ee] + s : s;
  }
}

// Last layer on the DySparse4 path, after bwd_last_reduce (which also sums the forward's Gram copies into gsum):
// BatchNorm-backward constants (a, p, q) of layer 3 and its parameter gradients, and the three small matrices the
// rewritten backward needs, from the layer's bf16 weight image W [C3][Kp]:
//     Gm = W^T diag(p) W          bf16 image [128][C2] (rows >= C2 zero)      block k1 -> row k1   (W staged in smem)
//     r  = W^T q                  [C2]                                          block 0
//     E  = diag(p) W Gram + q s^T [C3][C2] fp32 (s = column C2 of gsum)         block k1 -> column k1, thread = row c:
//                                 Gram is symmetric, so the column is the contiguous row k1 of gsum; each thread
//                                 dots its own W row (16-byte global loads, all in flight at once) with it
// Grid = 128 blocks of 256 threads; dynamic smem = C3*Kp*2 + 8*C3 + 4*ld bytes.  This kernel sits on the backward's
// critical path: every global load is issued in large independent batches (a first version that looped over global
// memory took 70 us, staging with an unbatched copy loop 12 us).
__global__ void __launch_bounds__(256)
l3_prep_kernel(BnBwdFin fin, int C3, int C2, const __nv_bfloat16* __restrict__ Wb, int Kp,
               const float* __restrict__ gsum, int ld, float* __restrict__ a_out, float* __restrict__ p_out,
               float* __restrict__ q_out, __nv_bfloat16* __restrict__ gmimg, float* __restrict__ rvec,
               float* __restrict__ E) {
  extern __shared__ uint4 l3sm[];
  const __nv_bfloat16* sW = reinterpret_cast<const __nv_bfloat16*>(l3sm);
  const int nw4 = C3 * Kp / 8;
  float* sp = reinterpret_cast<float*>(l3sm + nw4);
  float* sq = sp + C3;
  float* sg = sq + C3;   // row k1 of gsum
  const int k1 = blockIdx.x, t = threadIdx.x, nthr = blockDim.x;
  {
    const uint4* src = reinterpret_cast<const uint4*>(Wb);
    int e = t;
    for (; e + 7 * nthr < nw4; e += 8 * nthr) {
      uint4 r[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) r[u] = __ldg(src + e + u * nthr);
#pragma unroll
      for (int u = 0; u < 8; ++u) l3sm[e + u * nthr] = r[u];
    }
    for (; e < nw4; e += nthr) l3sm[e] = __ldg(src + e);
  }
  if (k1 < C2 && t < ld) sg[t] = __ldg(gsum + (size_t)k1 * ld + t);
  for (int c = t; c < C3; c += nthr) {
    float a, p, q;
    fin.eval(c, C3, k1 == 0, a, p, q);
    sp[c] = p; sq[c] = q;
    if (k1 == 0) { a_out[c] = a; p_out[c] = p; q_out[c] = q; }
  }
  __syncthreads();
  if (t < C2) {
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    const int kr = C2 == 64 ? (k1 & 63) : k1;   // C2 == 64: rows 64..127 repeat 0..63 (half-block epilogues, kHalf)
    if (kr < C2) {
#pragma unroll 4
      for (int c = 0; c < C3; c += 4) {
#pragma unroll
        for (int u = 0; u < 4; ++u)
          acc[u] = fmaf(__bfloat162float(sW[(c + u) * Kp + kr]) * sp[c + u], __bfloat162float(sW[(c + u) * Kp + t]), acc[u]);
      }
    }
    gmimg[(size_t)k1 * C2 + t] = __float2bfloat16_rn((acc[0] + acc[1]) + (acc[2] + acc[3]));
    if (k1 == 0) {
      float r0 = 0.f, r1 = 0.f;
      for (int c = 0; c < C3; c += 2) {
        r0 = fmaf(__bfloat162float(sW[c * Kp + t]), sq[c], r0);
        r1 = fmaf(__bfloat162float(sW[(c + 1) * Kp + t]), sq[c + 1], r1);
      }
      rvec[t] = r0 + r1;
    }
  }
  if (k1 < C2) {
    for (int c = t; c < C3; c += nthr) {
      const uint4* wr = reinterpret_cast<const uint4*>(Wb + (size_t)c * Kp);
      float e0 = 0.f, e1 = 0.f;
      for (int j0 = 0; j0 < C2; j0 += 64) {   // C2 in {64, 128}
        uint4 w[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) w[u] = __ldg(wr + (j0 >> 3) + u);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          float v[8];
          unpack8(w[u], v);
#pragma unroll
          for (int x = 0; x < 8; x += 2) {
            e0 = fmaf(v[x], sg[j0 + u * 8 + x], e0);
            e1 = fmaf(v[x + 1], sg[j0 + u * 8 + x + 1], e1);
          }
        }
      }
      E[(size_t)c * C2 + k1] = fmaf(sp[c], e0 + e1, sq[c] * sg[C2]);
    }
  }
}

// fp32 [C_out][C_in] -> zero-padded bf16 [Rp][Kp]; perm_d >= 0: layer-1 column order [feats(perm_d) | xyz(3)]
// dup_rows / dup_cols (64-channel layers, "two half-block epilogue threads per channel", see kHalf below): image rows /
// columns 64..127 repeat rows / columns 0..63 instead of being zero, so TMEM lanes 64..127 carry the same channels
struct ConvW4 { const float* W; __nv_bfloat16* dst; int cout, cin, Rp, Kp, perm_d, dup_rows, dup_cols; };
// one thread = 8 consecutive image columns (Kp is a multiple of 128): one 16-byte store, one division
__global__ void convert_weights4_kernel(ConvW4 a, ConvW4 b, ConvW4 c) {
  const ConvW4* L[3] = {&a, &b, &c};
  const int n0 = a.Rp * a.Kp / 8, n1 = b.Rp * b.Kp / 8, n2 = c.Rp * c.Kp / 8;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n0 + n1 + n2; e += gridDim.x * blockDim.x) {
    const int l = e < n0 ? 0 : (e < n0 + n1 ? 1 : 2);
    const ConvW4& w = *L[l];
    const int ee = e - (l == 0 ? 0 : (l == 1 ? n0 : n0 + n1));
    const int kp8 = w.Kp >> 3, ri = ee / kp8, k0 = (ee - ri * kp8) * 8;
    const int r = (w.dup_rows && ri < 128) ? (ri & 63) : ri;
    float v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int k = (w.dup_cols && k0 + u < 128) ? ((k0 + u) & 63) : k0 + u;
      int src = k;
      if (w.perm_d >= 0) src = k < w.perm_d ? k + 3 : (k < w.perm_d + 3 ? k - w.perm_d : w.cin);
      v[u] = (r < w.cout && src < w.cin) ? __ldg(w.W + (size_t)r * w.cin + src) : 0.f;
    }
    reinterpret_cast<uint4*>(w.dst)[ee] = tc::pack8_bf16(v);
  }
}

}  // namespace v4
}  // namespace pcoe
