// sa.cu — pcoe_sa_forward / pcoe_sa_backward: the set-abstraction MLP
// (PointNetSetAbstraction.forward, models/pointnet_pp_8dir.py:21-43, and its autograd).
//
// Forward (train):   L1 gather->GEMM->y1,stats | bn1 | L2 relu(bn(y1))->GEMM->y2,stats | bn2 |
//                    L3 relu(bn(y2))->GEMM->y3,stats,group max/min | bn3 | out = relu(a*sel+b)
// Backward (train):  reduce(gm,sums3) | consts3 | wgrad3, dgrad3->dz2,sums2 | consts2 |
//                    wgrad2, dgrad2->dz1,sums1 | consts1 | wgrad1, dgrad1->scatter-add
// Conv biases are not added in train mode: BatchNorm subtracts the batch mean, which cancels a
// per-channel constant exactly; they only shift running_mean (and enter eval mode).
#include "sa_common.cuh"
#include "sa_tc.cuh"
#include "sa_tc4.cuh"
#include "sa_tc5.cuh"
#include "sa_tc6.cuh"
#include "pm_tma.cuh"
#include "sa_layout.h"

namespace pcoe {

// ---- small kernels ---------------------------------------------------------------------------

// batch statistics -> affine (scale, shift), saved (mean, invstd), running-stat update
__global__ void bn_finalize_kernel(const double* __restrict__ sums, double count, int C,
                                   const float* __restrict__ gamma, const float* __restrict__ beta,
                                   const float* __restrict__ bias, float* __restrict__ running_mean,
                                   float* __restrict__ running_var, float eps, float momentum, int train,
                                   float* __restrict__ scale, float* __restrict__ shift,
                                   float* __restrict__ mean_out, float* __restrict__ invstd_out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float b = bias ? bias[c] : 0.f;
  if (train) {
    double t0 = 0.0, t1 = 0.0;   // the v4 / v5 epilogues spread their atomics over kRedCopies copies
    for (int k = 0; k < kRedCopies; ++k) { t0 += sums[(size_t)k * 2 * C + c]; t1 += sums[(size_t)k * 2 * C + C + c]; }
    const double mean = t0 / count;
    double var = t1 / count - mean * mean;  // biased, as BatchNorm normalises
    var = var < 0.0 ? 0.0 : var;
    const float invstd = (float)(1.0 / sqrt(var + (double)eps));
    const float sc = gamma[c] * invstd;
    scale[c] = sc;
    shift[c] = beta[c] - (float)mean * sc;
    mean_out[c] = (float)mean;
    invstd_out[c] = invstd;
    if (running_mean) {
      const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
      running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)(mean + (double)b);
      running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
    }
  } else {
    const float invstd = 1.f / sqrtf(running_var[c] + eps);
    const float sc = gamma[c] * invstd;
    scale[c] = sc;
    shift[c] = beta[c] + (b - running_mean[c]) * sc;
    if (mean_out) { mean_out[c] = running_mean[c] - b; invstd_out[c] = invstd; }
  }
}

// out = relu(a * (a >= 0 ? ymax : ymin) + b), slot = the matching arg, ysel = the selected y
// fin.sums != nullptr (v4/v5 train mode): the last layer's batch statistics are finalised here, per block,
// into a shared-memory table (block 0 writes the saved statistics / running buffers).
__global__ void sa_out_finalize_kernel(const float* __restrict__ ymax, const float* __restrict__ ymin,
                                       const uint8_t* __restrict__ amax, const uint8_t* __restrict__ amin,
                                       const float* __restrict__ scale, const float* __restrict__ shift,
                                       size_t total, int C, float* __restrict__ out,
                                       uint8_t* __restrict__ slot, float* __restrict__ ysel, v4::BnFin fin,
                                       long long* nbt0, long long* nbt1, long long* nbt2) {
  extern __shared__ float fin_tab[];   // [2,C] when fused
  if (blockIdx.x == 0 && threadIdx.x < 3) {   // BatchNorm2d.num_batches_tracked += 1 (train mode; NULL otherwise)
    long long* n = threadIdx.x == 0 ? nbt0 : (threadIdx.x == 1 ? nbt1 : nbt2);
    if (n) *n += 1;
  }
  if (fin.sums) {
    const bool w = blockIdx.x == 0;
    for (int c = threadIdx.x; c < C; c += blockDim.x) fin.eval(c, C, w, fin_tab[c], fin_tab[C + c]);
    __syncthreads();
    scale = fin_tab;
    shift = fin_tab + C;
  }
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(e % C);
    const float a = scale[c];
    const bool up = !signbit(a);   // = the sign of gamma (a = gamma * invstd), the rule the v4 / v5 epilogue uses to pick max or min
    const float ys = up ? ymax[e] : ymin[e];
    out[e] = fmaxf(fmaf(ys, a, shift[c]), 0.f);
    if (slot) { slot[e] = up ? amax[e] : amin[e]; ysel[e] = ys; }
  }
}

// gm = grad_out * [out > 0];  sums3 = (sum gm, sum gm * xhat(ysel)) per channel
__global__ void bwd_last_reduce_kernel(const float* __restrict__ grad_out, const float* __restrict__ out,
                                       const float* __restrict__ ysel, const float* __restrict__ mean,
                                       const float* __restrict__ invstd, int G, int C, int groups_per_block,
                                       float* __restrict__ gm, double* __restrict__ sums,
                                       const float* __restrict__ prescale /* DySparse4 path: gm is stored times a = scale */,
                                       const float* __restrict__ gram, float* __restrict__ gsum, int gram_n) {
  // DySparse4 path: the forward's kRedCopies Gram copies are summed here (independent work, no extra launch)
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < gram_n; e += gridDim.x * blockDim.x) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < kRedCopies; ++k) s += __ldg(gram + (size_t)k * gram_n + e);
    gsum[e] = s;
  }
  const int g0 = blockIdx.x * groups_per_block, g1 = min(G, g0 + groups_per_block);
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float is = invstd[c], nmi = -mean[c] * is;
    const float ps = prescale ? prescale[c] : 1.f;
    float s0[4] = {0.f, 0.f, 0.f, 0.f}, s1[4] = {0.f, 0.f, 0.f, 0.f};
    int g = g0;
    for (; g + 4 <= g1; g += 4) {          // four independent groups per iteration: the twelve loads overlap
      float o[4], go[4], ys[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const size_t e = (size_t)(g + k) * C + c;
        o[k] = __ldg(out + e); go[k] = __ldg(grad_out + e); ys[k] = __ldg(ysel + e);
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float v = o[k] > 0.f ? go[k] : 0.f;
        gm[(size_t)(g + k) * C + c] = v * ps;
        s0[k] += v;
        s1[k] = fmaf(v, fmaf(ys[k], is, nmi), s1[k]);
      }
    }
    for (; g < g1; ++g) {
      const size_t e = (size_t)g * C + c;
      const float v = out[e] > 0.f ? grad_out[e] : 0.f;
      gm[e] = v * ps;
      s0[0] += v;
      s1[0] = fmaf(v, fmaf(ysel[e], is, nmi), s1[0]);
    }
    double* dst = sums + (size_t)(blockIdx.x % kRedCopies) * 2 * C;
    atomicAdd(dst + c, (double)((s0[0] + s0[1]) + (s0[2] + s0[3])));
    atomicAdd(dst + C + c, (double)((s1[0] + s1[1]) + (s1[2] + s1[3])));
  }
}

// (sum dz, sum dz*xhat) -> BN-backward constants (a,p,q) + parameter gradients
__global__ void bn_bwd_consts_kernel(const double* __restrict__ sums, double count, int C,
                                     const float* __restrict__ scale, const float* __restrict__ mean,
                                     const float* __restrict__ invstd, float* __restrict__ a,
                                     float* __restrict__ p, float* __restrict__ q,
                                     float* __restrict__ dgamma, float* __restrict__ dbeta,
                                     float* __restrict__ dbias, int accumulate) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double s0 = 0.0, s1 = 0.0;
  for (int k = 0; k < kRedCopies; ++k) { s0 += sums[(size_t)k * 2 * C + c]; s1 += sums[(size_t)k * 2 * C + C + c]; }
  const double m1 = s0 / count, m2 = s1 / count;
  const double av = scale[c];
  const double pv = -av * (double)invstd[c] * m2;
  a[c] = (float)av;
  p[c] = (float)pv;
  q[c] = (float)(-av * m1 - pv * (double)mean[c]);
  if (dgamma) dgamma[c] = (accumulate ? dgamma[c] : 0.f) + (float)s1;
  if (dbeta) dbeta[c] = (accumulate ? dbeta[c] : 0.f) + (float)s0;
  if (dbias && !accumulate) dbias[c] = 0.f;  // exactly cancelled by the batch-mean subtraction
}

// per-layer kernel names for the profile report: sa1 = no input features, sa3 = group-all, sa2 = the rest
enum { kF1, kF2, kF3, kBL3, kBL2, kBL1, kWG3, kDG3, kWG2, kDG2, kWG1, kDG1, kNumNames };
static const char* kname(const pcoe_sa_desc& d, int which) {
  static const char* t[3][kNumNames] = {
      {"sa1_fwd_l1", "sa1_fwd_l2", "sa1_fwd_l3", "sa1_bwd_l3", "sa1_bwd_l2", "sa1_bwd_l1", "sa1_bwd_wgrad3",
       "sa1_bwd_dgrad3", "sa1_bwd_wgrad2", "sa1_bwd_dgrad2", "sa1_bwd_wgrad1", "sa1_bwd_dgrad1"},
      {"sa2_fwd_l1", "sa2_fwd_l2", "sa2_fwd_l3", "sa2_bwd_l3", "sa2_bwd_l2", "sa2_bwd_l1", "sa2_bwd_wgrad3",
       "sa2_bwd_dgrad3", "sa2_bwd_wgrad2", "sa2_bwd_dgrad2", "sa2_bwd_wgrad1", "sa2_bwd_dgrad1"},
      {"sa3_fwd_l1", "sa3_fwd_l2", "sa3_fwd_l3", "sa3_bwd_l3", "sa3_bwd_l2", "sa3_bwd_l1", "sa3_bwd_wgrad3",
       "sa3_bwd_dgrad3", "sa3_bwd_wgrad2", "sa3_bwd_dgrad2", "sa3_bwd_wgrad1", "sa3_bwd_dgrad1"}};
  return t[d.group_all ? 2 : (d.D == 0 ? 0 : 1)][which];
}

// ---- launch helpers --------------------------------------------------------------------------

template <class AProd, class Epi, bool BT>
static int launch_nt(const AProd& ap, const float* Bmat, int ldb, const Epi& epi, int M, int Ncols,
                     int Kdim, cudaStream_t st, const char* what) {
  LaunchScope ls(what, st);
  if (Ncols > 64) {
    auto k = gemm_nt_kernel<AProd, Epi, 8, BT>;
    static bool attr = false;
    if (!attr) {
      PCOE_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gemm_nt_smem<8>()));
      attr = true;
    }
    dim3 grid(ceil_div(M, kBM), ceil_div(Ncols, 128));
    k<<<grid, 256, gemm_nt_smem<8>(), st>>>(ap, Bmat, ldb, epi, M, Ncols, Kdim);
  } else {
    auto k = gemm_nt_kernel<AProd, Epi, 4, BT>;
    static bool attr = false;
    if (!attr) {
      PCOE_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gemm_nt_smem<4>()));
      attr = true;
    }
    dim3 grid(ceil_div(M, kBM), 1);
    k<<<grid, 256, gemm_nt_smem<4>(), st>>>(ap, Bmat, ldb, epi, M, Ncols, Kdim);
  }
  return ls.done();
}

template <class PProd, class QProd>
static int launch_tn(const PProd& pp, const QProd& qp, float* out, int ldo, int M, int Ca, int Cb,
                     cudaStream_t st, const char* what) {
  const int ta = ceil_div(Ca, 64), tb = ceil_div(Cb, 64);
  int splits = ceil_div(kNumSMs * 4, ta * tb);
  const int max_splits = ceil_div(M, kWgRows * 4);
  splits = splits < 1 ? 1 : (splits > max_splits ? max_splits : splits);
  int rps = ceil_div(ceil_div(M, splits), kWgRows) * kWgRows;
  splits = ceil_div(M, rps);
  dim3 grid(ta, tb, splits);
  LaunchScope ls(what, st);
  gemm_tn_kernel<PProd, QProd><<<grid, 256, 0, st>>>(pp, qp, out, ldo, M, Ca, Cb, rps);
  return ls.done();
}

template <class AProd, class Epi>
static int launch_nt_tc(const AProd& ap, const __nv_bfloat16* Bw, int ldb, int brows, const Epi& epi, int M,
                        int Ncols, int Kdim, cudaStream_t st, const char* what) {
  auto k = tc_gemm_nt_kernel<AProd, Epi>;
  static bool attr = false;
  if (!attr) {
    PCOE_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcNtSmem));
    attr = true;
  }
  dim3 grid(ceil_div(M, kTcTile), ceil_div(Ncols, kTcN));
  LaunchScope ls(what, st);
  k<<<grid, 256, kTcNtSmem, st>>>(ap, Bw, ldb, brows, epi, M, Ncols, Kdim);
  return ls.done();
}

template <class PProd, class QProd>
static int launch_tn_tc(const PProd& pp, const QProd& qp, float* out, int ldo, int M, int Ca, int Cb,
                        cudaStream_t st, const char* what) {
  auto k = tc_gemm_tn_kernel<PProd, QProd>;
  static bool attr = false;
  if (!attr) {
    PCOE_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcTnSmem));
    attr = true;
  }
  const int ta = ceil_div(Ca, 128), tb = ceil_div(Cb, 128);
  int splits = ceil_div(kNumSMs * 2, ta * tb);
  const int max_splits = ceil_div(M, 128);
  splits = splits < 1 ? 1 : (splits > max_splits ? max_splits : splits);
  const int rps = ceil_div(ceil_div(M, splits), 128) * 128;
  splits = ceil_div(M, rps);
  dim3 grid(ta, tb, splits);
  LaunchScope ls(what, st);
  k<<<grid, 256, kTcTnSmem, st>>>(pp, qp, out, ldo, M, Ca, Cb, rps);
  return ls.done();
}

// bf16 copies of the three weight matrices (and their transposes for dgrad), tensor-core path only.
// On the v2 path layer 1 uses the "features first" channel order [feats(D) | xyz(3)].
static int convert_weights(const pcoe_sa_desc& d, const SaLayout& L, const pcoe_sa_params& P, char* base,
                           cudaStream_t st) {
  const int Cs[3] = {d.C1, d.C2, d.C3}, Kin[3] = {3 + d.D, d.C1, d.C2};
  for (int l = 0; l < 3; ++l) {
    const int total = L.wb_rows[l] * L.wb_k[l] + L.wbt_rows[l] * L.wbt_k[l];
    LaunchScope ls("convert_weights_kernel", st);
    convert_weights_kernel<<<min(ceil_div(total, 256), kNumSMs * 4), 256, 0, st>>>(
        P.W[l], Cs[l], Kin[l], (__nv_bfloat16*)(base + L.wb_off[l]), L.wb_rows[l], L.wb_k[l],
        (__nv_bfloat16*)(base + L.wbt_off[l]), L.wbt_rows[l], L.wbt_k[l], (L.v2 && l == 0) ? d.D : -1);
    PCOE_TRY(ls.done());
  }
  return PCOE_OK;
}

// ---- v4 (channel-on-lane, persistent, warp-specialised) launchers --------------------------------
template <class Prod>
static size_t tile_bytes4(const Prod& p) {   // host mirror of v4::prod_tile_bytes, rounded to 1 KB
  size_t b = Prod::kChMajor ? (size_t)2 * p.rows() * 128 : (size_t)((p.kext() + 63) / 64) * v4::kPts * 128;
  return align_up(b, 1024);
}
constexpr size_t kSmemBudget4 = 224 * 1024;

template <int GRAM = 0, class Prod, class Epi>
static int launch_fwd4(const Prod& prod, const __nv_bfloat16* Wb, int Rp, int Kp, const Epi& epi, int M,
                       cudaStream_t st, const char* what, float* gram = nullptr) {
  const size_t wbytes = (size_t)Rp * Kp * 2, tb = tile_bytes4(prod);
  const size_t cbytes = sizeof(float) * (size_t)(prod.nconst() + epi.nconst()) + 512, sb = (size_t)epi.stage_bytes();
  const size_t avail = kSmemBudget4 - 1024 - wbytes - cbytes;
  // prefer two operand stages, then a double-buffered output staging tile, then more operand stages
  int nstg = (sb && avail >= 2 * sb + 2 * tb) ? 2 : 1;
  if (avail < nstg * sb + tb) return fail(PCOE_ERR_UNSUPPORTED, "%s: layer does not fit shared memory", what);
  int stages = (int)((avail - nstg * sb) / tb);
  stages = stages > v4::kMaxStages ? v4::kMaxStages : stages;
  const size_t smem = 1024 + wbytes + (size_t)stages * tb + cbytes + nstg * sb;
  const int tiles = ceil_div(M, v4::kPts), grid = tiles < kNumSMs ? tiles : kNumSMs;   // one persistent CTA per SM
  // TMEM: double-buffered accumulators when they fit beside the Gram accumulator (kext + 16 columns)
  const int nbuf = (2 * Rp + (GRAM ? prod.kext() + 16 : 0) <= 512) ? 2 : 1;
  if constexpr (Epi::kHalf && GRAM == 0) {
    if (epi.C == 64 && Rp == 128) {   // two half-block epilogue threads per channel (weight image rows duplicated)
      auto kh = v4::tc4_fwd_kernel<Prod, Epi, 512, GRAM, true>;
      static bool attrh = false;
      if (!attrh) { PCOE_CUDA(cudaFuncSetAttribute(kh, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBudget4)); attrh = true; }
      LaunchScope ls(what, st);
      kh<<<grid, v4::kThreads, smem, st>>>(prod, Wb, Rp, Kp, epi, M, stages, nstg, gram, nbuf);
      return ls.done();
    }
  }
  auto k = v4::tc4_fwd_kernel<Prod, Epi, 512, GRAM, false>;
  static bool attr = false;
  if (!attr) { PCOE_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBudget4)); attr = true; }
  LaunchScope ls(what, st);
  k<<<grid, v4::kThreads, smem, st>>>(prod, Wb, Rp, Kp, epi, M, stages, nstg, gram, nbuf);
  return ls.done();
}

template <int DGRAD, int GM = 0, class PProd, class QProd, class Epi>
static int launch_bwd4(const PProd& pp, const QProd& qp, const __nv_bfloat16* Wb, int Rp, int Kp, const Epi& epi,
                       float* dW, int ldo, int cq_valid, int perm_d, int M, int cprev, cudaStream_t st,
                       const char* what, const __nv_bfloat16* Gmb = nullptr, int gk = 0) {
  const size_t wbytes = (DGRAD ? (size_t)Rp * Kp * 2 : 0) + (GM ? (size_t)2 * 128 * gk : 0);
  const size_t cbytes = sizeof(float) * (size_t)(pp.nconst() + qp.nconst() + epi.nconst()) + 512, sb = (size_t)epi.stage_bytes();
  const size_t pq = tile_bytes4(pp) + tile_bytes4(qp), base = 1024 + wbytes + pq + cbytes;
  if (base + sb > kSmemMax4) return fail(PCOE_ERR_UNSUPPORTED, "%s: layer does not fit shared memory", what);
  // A second P/Q operand stage (producers working on tile i+1 while the MMAs of tile i run) is implemented
  // (npq = 2) but measured SLOWER on B200 (sa1_bwd_l2 55.7 -> 63.5 us, sa2_bwd_l2 39 -> 47.5 us): the larger
  // carve-out shrinks L1 and the producers running further ahead evict the y tile that the MaskStats epilogue
  // re-reads.  The remaining shared memory goes to a second output staging tile instead.
  // With the sparse last-layer operand (GM) the producers are light and the single stage serialises them with the
  // MMAs of the previous tile; PCOE_BWD_NPQ2=1 switches the second stage on where it fits (experiment switch).
  static const bool want2 = [] { const char* e = getenv("PCOE_BWD_NPQ2"); return e && e[0] == '1'; }();
  const int npq = (GM && want2 && base + pq + sb <= kSmemMax4) ? 2 : 1;
  const int nstg = (sb && base + (npq - 1) * pq + 2 * sb <= (npq == 2 ? kSmemMax4 : kSmemBudget4)) ? 2 : 1;
  const size_t smem = base + (npq - 1) * pq + nstg * sb;
  const int tiles = ceil_div(M, v4::kPts), grid = tiles < kNumSMs ? tiles : kNumSMs;
  if constexpr (Epi::kHalf && DGRAD == 1) {
    if (cprev == 64) {   // two half-block epilogue threads per channel (weight image columns duplicated)
      auto kh = v4::tc4_bwd_kernel<PProd, QProd, Epi, DGRAD, 512, GM, true>;
      static bool attrh = false;
      if (!attrh) { PCOE_CUDA(cudaFuncSetAttribute(kh, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemMax4)); attrh = true; }
      LaunchScope ls(what, st);
      kh<<<grid, v4::kThreads, smem, st>>>(pp, qp, Wb, Rp, Kp, epi, dW, ldo, cq_valid, perm_d, M, cprev, nstg, npq, Gmb, gk);
      return ls.done();
    }
  }
  auto k = v4::tc4_bwd_kernel<PProd, QProd, Epi, DGRAD, 512, GM, false>;
  static bool attr = false;
  if (!attr) { PCOE_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemMax4)); attr = true; }
  LaunchScope ls(what, st);
  k<<<grid, v4::kThreads, smem, st>>>(pp, qp, Wb, Rp, Kp, epi, dW, ldo, cq_valid, perm_d, M, cprev, nstg, npq, Gmb, gk);
  return ls.done();
}

// ---- v5 (K-chunk-streamed, (tile x channel-block) grid) launchers ----------------------------------
constexpr size_t kSmemBudget5 = 224 * 1024;
static int stages5(int want, size_t stage_bytes, size_t cbytes) {
  int n = (int)((kSmemBudget5 - 1024 - cbytes) / stage_bytes);
  n = n > v5::kMaxStages5 ? v5::kMaxStages5 : n;
  return want < n ? want : n;
}

template <class Prod, class Epi>
static int launch_fwd5(const Prod& prod, const __nv_bfloat16* Wb, int Kp, const Epi& epi, int M, int Cout, int nk,
                       cudaStream_t st, const char* what) {
  const size_t cbytes = sizeof(float) * (size_t)(prod.nconst() + epi.nconst()) + 512;
  const int nst = stages5(nk, 32768, cbytes);
  if (nst < 1) return fail(PCOE_ERR_UNSUPPORTED, "%s: layer does not fit shared memory", what);
  const size_t smem = 1024 + (size_t)nst * 32768 + cbytes;
  auto k = v5::tc5_fwd_kernel<Prod, Epi>;
  static bool attr = false;
  if (!attr) { PCOE_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBudget5)); attr = true; }
  LaunchScope ls(what, st);
  k<<<dim3(ceil_div(M, v4::kPts), Cout / 128), v4::kThreads, smem, st>>>(prod, Wb, Kp, epi, M, nst);
  return ls.done();
}

template <bool PT, class PProd, class Epi>
static int launch_dgrad5(const PProd& pp, const __nv_bfloat16* Wb, int Kp, const Epi& epi, int M, int Cprev,
                         cudaStream_t st, const char* what) {
  const size_t cbytes = sizeof(float) * (size_t)(pp.nconst() + epi.nconst()) + 512;
  const int nst = stages5(pp.C / 64, 32768, cbytes);
  if (nst < 1) return fail(PCOE_ERR_UNSUPPORTED, "%s: layer does not fit shared memory", what);
  const size_t smem = 1024 + (size_t)nst * 32768 + cbytes;
  auto k = v5::tc5_dgrad_kernel<PProd, Epi, PT>;
  static bool attr = false;
  if (!attr) { PCOE_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBudget5)); attr = true; }
  LaunchScope ls(what, st);
  k<<<dim3(ceil_div(M, v4::kPts), Cprev / 128), v4::kThreads, smem, st>>>(pp, Wb, Kp, epi, M, nst);
  return ls.done();
}

template <class PProd, class QProd>
static int launch_wgrad5(const PProd& pp, const QProd& qp, float* dW, int ldo, int cq_valid, int perm_d, int M, int nqb,
                         cudaStream_t st, const char* what) {
  const size_t cbytes = sizeof(float) * (size_t)(pp.nconst() + qp.nconst()) + 512;
  const int ntiles = ceil_div(M, v4::kPts), items = (pp.C / 128) * nqb;
  int splits = kNumSMs / items;
  splits = splits < 1 ? 1 : (splits > ntiles ? ntiles : splits);
  const int tps = ceil_div(ntiles, splits);
  splits = ceil_div(ntiles, tps);
  const int nst = stages5(tps, 65536, cbytes);
  if (nst < 1) return fail(PCOE_ERR_UNSUPPORTED, "%s: layer does not fit shared memory", what);
  const size_t smem = 1024 + (size_t)nst * 65536 + cbytes;
  auto k = v5::tc5_wgrad_kernel<PProd, QProd>;
  static bool attr = false;
  if (!attr) { PCOE_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBudget5)); attr = true; }
  LaunchScope ls(what, st);
  k<<<dim3(pp.C / 128, nqb, splits), v4::kThreads, smem, st>>>(pp, qp, dW, ldo, cq_valid, perm_d, M, tps, nst);
  return ls.done();
}


// ---- v6 (bf16x3: split-operand, fp32-accurate) launchers ---------------------------------------------
constexpr size_t kSmemBudget6 = 224 * 1024;
static bool sa_x3_supported(const pcoe_sa_desc& d) {
  return d.K == 32 && d.C1 % 64 == 0 && d.C2 % 64 == 0 && d.C3 % 64 == 0 && d.D % 64 == 0;
}
struct Cfg6 { int nst, wres, grid; size_t smem; };
// persistent (tile x channel-block) kernels: CTAs per channel block, resident weights or streamed, ring depth
static Cfg6 cfg6(int nk, int ncb, int ntiles, size_t cbytes, int np, int min_ring = 2) {
  Cfg6 c{};
  const int per = kNumSMs / ncb > 0 ? kNumSMs / ncb : 1;
  const int ctas = ntiles < per ? ntiles : per;
  c.grid = ctas * ncb;
  const size_t op = (size_t)np * 16384;   // one operand chunk, all planes
  const size_t avail = kSmemBudget6 - 1024 - cbytes, wb = (size_t)nk * op;
  c.wres = (ntiles > ctas && wb + min_ring * op <= avail) ? 1 : 0;   // a CTA that sees one tile gains nothing from residency
  const size_t ring = c.wres ? avail - wb : avail, sb = c.wres ? op : 2 * op;
  int n = (int)(ring / sb);
  c.nst = n > v6::kMaxStages6 ? v6::kMaxStages6 : n;
  c.smem = 1024 + (c.wres ? wb : 0) + (size_t)c.nst * sb + cbytes;
  return c;
}

constexpr int kFwdPlanes6 = 3;   // forward operands keep all 24 significant bits (6 MMAs per step), backward 16 (3 MMAs)
// PCOE_X3_ASYNC=0 (debug / A-B timing): register pipeline everywhere
static bool x3_async_enabled() {
  static const bool on = [] { const char* e = getenv("PCOE_X3_ASYNC"); return !(e && e[0] == '0'); }();
  return on;
}

template <int NP = kFwdPlanes6, class Prod, class Epi>
static int launch_fwd6(const Prod& prod, const __nv_bfloat16* Wp, size_t wps, int Kp, const Epi& epi, int M,
                       int Cout, cudaStream_t st, const char* what) {
  const size_t cbytes = sizeof(float) * (size_t)(prod.nconst() + epi.nconst()) + 512;
  const int ncb = ceil_div(Cout, 128);
  if constexpr (Prod::kAsync) {
    // resident weights + room for the raw staging ring: the producers copy their fp32 operands with cp.async
    const size_t raw = (size_t)v6::kRawDepth * Prod::kRawItems * v6::kRawItemBytes;
    const Cfg6 ca = cfg6(prod.nchunks(), ncb, ceil_div(M, v4::kPts), cbytes + raw, NP, 1);
    // never trade weight residency for the staging ring (measured: SA2 forward 42 -> 58 us when its weights are
    // streamed): order of preference = ring + resident weights, resident weights alone, ring + streamed weights
    const bool res_alone = cfg6(prod.nchunks(), ncb, ceil_div(M, v4::kPts), cbytes, NP, 1).wres;
    if (x3_async_enabled() && ca.nst >= 1 && (ca.wres || !res_alone)) {
      auto k = v6::x3_fwd_kernel<Prod, Epi, NP, true>;
      static bool attr = false;
      if (!attr) { PCOE_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBudget6)); attr = true; }
      LaunchScope ls(what, st);
      k<<<ca.grid, v4::kThreads, ca.smem, st>>>(prod, Wp, wps, Kp, epi, M, ncb, ca.nst, ca.wres);
      return ls.done();
    }
  }
  const Cfg6 c = cfg6(prod.nchunks(), ncb, ceil_div(M, v4::kPts), cbytes, NP, 1);
  if (c.nst < 1) return fail(PCOE_ERR_UNSUPPORTED, "%s: layer does not fit shared memory", what);
  auto k = v6::x3_fwd_kernel<Prod, Epi, NP, false>;
  static bool attr = false;
  if (!attr) { PCOE_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBudget6)); attr = true; }
  LaunchScope ls(what, st);
  k<<<c.grid, v4::kThreads, c.smem, st>>>(prod, Wp, wps, Kp, epi, M, ncb, c.nst, c.wres);
  return ls.done();
}

template <bool PT, class PProd, class Epi>
static int launch_dgrad6(const PProd& pp, const __nv_bfloat16* Wp, size_t wps, int Kp, const Epi& epi, int M,
                         int Cprev, cudaStream_t st, const char* what) {
  const size_t cbytes = sizeof(float) * (size_t)(pp.nconst() + epi.nconst()) + 512;
  const int ncb = ceil_div(Cprev, 128);
  {
    const size_t raw = (size_t)v6::kRawDepth * PProd::kRawItems * v6::kRawItemBytes;
    const Cfg6 ca = cfg6(pp.C / 64, ncb, ceil_div(M, v4::kPts), cbytes + raw, 2, 1);
    // dgrad (two planes): the staging ring beats weight residency (SA2 dgrad3: 74 us resident + register pipeline,
    // 50 us streamed + cp.async ring)
    if (x3_async_enabled() && ca.nst >= 1) {
      auto k = v6::x3_dgrad_kernel<PProd, Epi, PT, true>;
      static bool attr = false;
      if (!attr) { PCOE_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBudget6)); attr = true; }
      LaunchScope ls(what, st);
      k<<<ca.grid, v4::kThreads, ca.smem, st>>>(pp, Wp, wps, Kp, epi, M, ncb, ca.nst, ca.wres);
      return ls.done();
    }
  }
  const Cfg6 c = cfg6(pp.C / 64, ncb, ceil_div(M, v4::kPts), cbytes, 2);
  if (c.nst < 1) return fail(PCOE_ERR_UNSUPPORTED, "%s: layer does not fit shared memory", what);
  auto k = v6::x3_dgrad_kernel<PProd, Epi, PT, false>;
  static bool attr = false;
  if (!attr) { PCOE_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBudget6)); attr = true; }
  LaunchScope ls(what, st);
  k<<<c.grid, v4::kThreads, c.smem, st>>>(pp, Wp, wps, Kp, epi, M, ncb, c.nst, c.wres);
  return ls.done();
}

template <class PProd, class QProd>
static int launch_wgrad6(const PProd& pp, const QProd& qp, float* dW, int ldo, int cq_valid, int perm_d, int M, int nqb,
                         cudaStream_t st, const char* what) {
  const size_t cbytes = sizeof(float) * (size_t)(pp.nconst() + qp.nconst()) + 512;
  const int ntiles = ceil_div(M, v4::kPts), clb = ceil_div(pp.C, 128), items = clb * nqb;
  int splits = kNumSMs / items;
  splits = splits < 1 ? 1 : (splits > ntiles ? ntiles : splits);
  const int tps = ceil_div(ntiles, splits);
  splits = ceil_div(ntiles, tps);
  if constexpr (PProd::kAsync && QProd::kAsync) {
    // two producer groups (P units / Q units), each on its own cp.async ring: deepest pair of rings that fits beside
    // one operand stage (a 64-point Q unit of the gather producer is half the size of its 128-point forward unit)
    const size_t up = (size_t)PProd::kRawItems * v6::kRawItemBytes, uq = (size_t)QProd::kRawItems64 * v6::kRawItemBytes;
    const size_t room = kSmemBudget6 - 1024 - cbytes - 65536;
    int dP = 0, dQ = 0;
    if (kSmemBudget6 > 1024 + cbytes + 65536) {
      if (3 * up + 3 * uq <= room) { dP = 3; dQ = 3; }
      else if (3 * up + 2 * uq <= room) { dP = 3; dQ = 2; }
      else if (2 * up + 3 * uq <= room) { dP = 2; dQ = 3; }
      else if (2 * up + 2 * uq <= room) { dP = 2; dQ = 2; }
    }
    if (x3_async_enabled() && dP) {
      const size_t raw = dP * up + dQ * uq;
      int na = (int)((kSmemBudget6 - 1024 - cbytes - raw) / 65536);
      na = na > v6::kMaxStages6 ? v6::kMaxStages6 : na;
      const size_t smem = 1024 + (size_t)na * 65536 + raw + cbytes;
      auto k = v6::x3_wgrad_kernel<PProd, QProd, true>;
      static bool attr = false;
      if (!attr) { PCOE_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBudget6)); attr = true; }
      LaunchScope ls(what, st);
      k<<<dim3(clb, nqb, splits), v4::kThreads, smem, st>>>(pp, qp, dW, ldo, cq_valid, perm_d, M, tps, na, dP * 16 + dQ);
      return ls.done();
    }
  }
  int nst = (int)((kSmemBudget6 - 1024 - cbytes) / 65536);
  nst = nst > v6::kMaxStages6 ? v6::kMaxStages6 : nst;
  if (nst < 1) return fail(PCOE_ERR_UNSUPPORTED, "%s: layer does not fit shared memory", what);
  const size_t smem = 1024 + (size_t)nst * 65536 + cbytes;
  auto k = v6::x3_wgrad_kernel<PProd, QProd, false>;
  static bool attr = false;
  if (!attr) { PCOE_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBudget6)); attr = true; }
  LaunchScope ls(what, st);
  k<<<dim3(clb, nqb, splits), v4::kThreads, smem, st>>>(pp, qp, dW, ldo, cq_valid, perm_d, M, tps, nst, 0);
  return ls.done();
}

static int convert_weights6(const pcoe_sa_desc& d, const SaLayout& L, const pcoe_sa_params& P, char* base, cudaStream_t st) {
  const int Cs[3] = {d.C1, d.C2, d.C3}, Kin[3] = {3 + d.D, d.C1, d.C2};
  v6::ConvW6 w[3];
  int total = 0;
  for (int l = 0; l < 3; ++l) {
    w[l] = v6::ConvW6{P.W[l], (__nv_bfloat16*)(base + L.wb_off[l]), Cs[l], Kin[l], L.w4_rp[l], L.w4_kp[l], l == 0 ? d.D : -1, 3,
                      (l >= 1 && Kin[l] == 64 && L.w4_kp[l] == 128) ? 1 : 0};
    total += L.w4_rp[l] * L.w4_kp[l];
  }
  LaunchScope ls("convert_weights_kernel", st);
  v6::convert_weights6_kernel<<<min(ceil_div(total / 8, 256), kNumSMs * 4), 256, 0, st>>>(w[0], w[1], w[2]);
  return ls.done();
}

static int convert_weights4(const pcoe_sa_desc& d, const SaLayout& L, const pcoe_sa_params& P, char* base,
                            cudaStream_t st) {
  const int Cs[3] = {d.C1, d.C2, d.C3}, Kin[3] = {3 + d.D, d.C1, d.C2};
  v4::ConvW4 w[3];
  int total = 0;
  for (int l = 0; l < 3; ++l) {
    // 64-channel layers of the v4 kernels: rows / columns 64..127 of the image repeat 0..63 (half-block epilogues)
    const int dup_rows = L.v2 && l < 2 && Cs[l] == 64 && L.w4_rp[l] == 128;
    const int dup_cols = L.v2 && l >= 1 && Kin[l] == 64 && L.w4_kp[l] == 128;
    w[l] = v4::ConvW4{P.W[l], (__nv_bfloat16*)(base + L.wb_off[l]), Cs[l], Kin[l], L.w4_rp[l], L.w4_kp[l], l == 0 ? d.D : -1,
                      dup_rows, dup_cols};
    total += L.w4_rp[l] * L.w4_kp[l];
  }
  LaunchScope ls("convert_weights_kernel", st);
  v4::convert_weights4_kernel<<<min(ceil_div(total / 8, 256), kNumSMs * 4), 256, 0, st>>>(w[0], w[1], w[2]);
  return ls.done();
}

static int validate(const pcoe_sa_desc* d) {
  if (!d) return fail(PCOE_ERR_NULL, "sa: desc is NULL");
  if (d->B <= 0 || d->N <= 0 || d->S <= 0 || d->K <= 0 || d->D < 0 || d->C1 <= 0 || d->C2 <= 0 || d->C3 <= 0)
    return fail(PCOE_ERR_BAD_SHAPE, "sa: B=%d N=%d S=%d K=%d D=%d C=(%d,%d,%d)", d->B, d->N, d->S, d->K,
                d->D, d->C1, d->C2, d->C3);
  if (d->group_all && (d->S != 1 || d->K != d->N))
    return fail(PCOE_ERR_BAD_SHAPE, "sa: group_all needs S=1 and K=N (S=%d K=%d N=%d)", d->S, d->K, d->N);
  if (d->K > 128 || (d->K & (d->K - 1)))
    return fail(PCOE_ERR_UNSUPPORTED, "sa: K=%d must be a power of two <= 128", d->K);
  if (d->train && (long long)d->B * d->S * d->K < 2)
    return fail(PCOE_ERR_BAD_SHAPE, "sa: train-mode BatchNorm needs more than 1 value per channel");
  if (d->precision != PCOE_PRECISION_FP32 && d->precision != PCOE_PRECISION_BF16 && d->precision != PCOE_PRECISION_BF16X3)
    return fail(PCOE_ERR_UNSUPPORTED, "sa: precision=%d", d->precision);
  if (d->precision == PCOE_PRECISION_BF16X3 && !sa_x3_supported(*d))
    return fail(PCOE_ERR_UNSUPPORTED, "sa: bf16x3 needs K == 32 and C1, C2, C3, D multiples of 64 (K=%d D=%d C=(%d,%d,%d))",
                d->K, d->D, d->C1, d->C2, d->C3);
  return PCOE_OK;
}

template <typename TY, bool TC>
static int sa_forward_impl(const pcoe_sa_desc& d, const float* xyz, const float* new_xyz, const int32_t* nbr,
                           const float* feats, const pcoe_sa_params& P, float* out, void* saved,
                           void* workspace, cudaStream_t st) {
  const SaLayout L = sa_layout(d);
  char* sv = (char*)saved;
  char* ws = (char*)workspace;
  const int M = L.M, G = L.G, Cin = 3 + d.D;
  const int Cs[3] = {d.C1, d.C2, d.C3};
  const bool train = d.train != 0;

  // pre-BN activations: saved for backward in train mode, transient otherwise
  TY* y[3];
  for (int l = 0; l < 3; ++l) y[l] = train ? (TY*)(sv + L.sv_y[l]) : (l < 2 ? (TY*)(ws + L.ws_y[l]) : nullptr);
  if (L.l3s) y[2] = nullptr;
  float *scale[3], *shift[3], *mean[3], *invstd[3];
  for (int l = 0; l < 3; ++l) {
    char* base = train ? sv + L.sv_stat[l] : ws + L.ws_stat[l];
    scale[l] = (float*)base; shift[l] = scale[l] + Cs[l]; mean[l] = shift[l] + Cs[l]; invstd[l] = mean[l] + Cs[l];
  }
  double* sums[3];
  for (int l = 0; l < 3; ++l) sums[l] = train ? (double*)(ws + L.ws_sums[l]) : nullptr;
  if (train) PCOE_CUDA(cudaMemsetAsync(ws + L.ws_sums[0], 0, L.ws_sums_bytes, st));
  float* ymax = (float*)(ws + L.ws_ymax);
  float* ymin = (float*)(ws + L.ws_ymin);
  uint8_t* amax = (uint8_t*)(ws + L.ws_amax);
  uint8_t* amin = (uint8_t*)(ws + L.ws_amin);

  auto finalize = [&](int l) -> int {
    LaunchScope ls("bn_finalize_kernel", st);
    bn_finalize_kernel<<<ceil_div(Cs[l], 128), 128, 0, st>>>(sums[l], (double)M, Cs[l], P.gamma[l], P.beta[l],
        P.bias[l], P.running_mean[l], P.running_var[l], d.eps, d.momentum, d.train, scale[l], shift[l],
        mean[l], invstd[l]);
    return ls.done();
  };
  if (!train) for (int l = 0; l < 3; ++l) PCOE_TRY(finalize(l));
  // v4 / v5 train mode: the statistics of layer l are finalised inline by the kernel that consumes them
  auto mkfin = [&](int l) {
    v4::BnFin f{};
    if (train) {
      f.sums = sums[l]; f.count = (double)M; f.inv_count = 1.0 / (double)M; f.gamma = P.gamma[l]; f.beta = P.beta[l]; f.bias = P.bias[l];
      f.running_mean = P.running_mean[l]; f.running_var = P.running_var[l]; f.eps = d.eps; f.momentum = d.momentum;
      f.scale = scale[l]; f.shift = shift[l]; f.mean = mean[l]; f.invstd = invstd[l];
    }
    return f;
  };
  v4::BnFin fin_out{};
  static const bool sep_fin = [] { const char* e = getenv("PCOE_SA_FINALIZE_KERNEL"); return e && e[0] == '1'; }();

  const int Kin[3] = {Cin, d.C1, d.C2};
  char* wbase = train ? sv : ws;   // bf16 weight copies live with the saved state in train mode
  bool use4 = false, use5 = false;
  if constexpr (TC) { use4 = L.v2; use5 = L.v5; }
  if (TC && !use4 && !use5) PCOE_TRY(convert_weights(d, L, P, wbase, st));
  auto nt = [&](const auto& ap, int l, const auto& epi, const char* what) -> int {
    using AP = std::decay_t<decltype(ap)>;
    using EP = std::decay_t<decltype(epi)>;
    if constexpr (TC)
      return launch_nt_tc<AP, EP>(ap, (const __nv_bfloat16*)(wbase + L.wb_off[l]), L.wb_k[l], L.wb_rows[l], epi, M,
                                  Cs[l], Kin[l], st, what);
    else
      return launch_nt<AP, EP, false>(ap, P.W[l], Kin[l], epi, M, Cs[l], Kin[l], st, what);
  };

  bool done = false;
  if constexpr (TC) {
    if (use4) {   // channel-on-lane persistent kernels, activations channel-major [C][Mld]
      PCOE_TRY(convert_weights4(d, L, P, wbase, st));
      auto wb = [&](int l) { return (const __nv_bfloat16*)(wbase + L.wb_off[l]); };
      const int Mld = L.Mld;
      v4::StoreStats4 e0{}; e0.y = y[0]; e0.sums = sums[0]; e0.C = d.C1; e0.Mld = Mld;
      if (d.D == 0) {
        v4::GatherXyz4 gp{v4::GatherBase{xyz, new_xyz, nbr, d.N, d.S, d.group_all, M}};
        PCOE_TRY(launch_fwd4(gp, wb(0), L.w4_rp[0], L.w4_kp[0], e0, M, st, kname(d, kF1)));
      } else {
        v4::GatherFeat4 gp{v4::GatherBase{xyz, new_xyz, nbr, d.N, d.S, d.group_all, M}, feats, d.D};
        PCOE_TRY(launch_fwd4(gp, wb(0), L.w4_rp[0], L.w4_kp[0], e0, M, st, kname(d, kF1)));
      }
      v4::BnRelu4 p1{}; p1.y = y[0]; p1.scale = scale[0]; p1.shift = shift[0]; p1.M = M; p1.Mld = Mld; p1.C = d.C1;
      p1.fin = mkfin(0);
      v4::StoreStats4 e1{}; e1.y = y[1]; e1.sums = sums[1]; e1.C = d.C2; e1.Mld = Mld;
      PCOE_TRY(launch_fwd4(p1, wb(1), L.w4_rp[1], L.w4_kp[1], e1, M, st, kname(d, kF2)));
      v4::BnRelu4 p2{}; p2.y = y[1]; p2.scale = scale[1]; p2.shift = shift[1]; p2.M = M; p2.Mld = Mld; p2.C = d.C2;
      p2.fin = mkfin(1);
      v4::Group4 e2{}; e2.y = y[2]; e2.sums = sums[2]; e2.ymax = ymax; e2.ymin = ymin; e2.amax = amax; e2.amin = amin;
      e2.C = d.C3; e2.Mld = Mld; e2.gamma = P.gamma[2];
      if (L.l3s) {   // y3 is not stored: the kernel accumulates the Gram matrix of its input instead (DySparse4 backward)
        float* gram = (float*)(sv + L.sv_gram);
        PCOE_CUDA(cudaMemsetAsync(gram, 0, L.sv_gram_bytes, st));
        e2.y = nullptr;
        p2.prows = L.gram_ld > 128 ? L.gram_ld : 128;
        PCOE_TRY(launch_fwd4<1>(p2, wb(2), L.w4_rp[2], L.w4_kp[2], e2, M, st, kname(d, kF3), gram));
      } else {
        PCOE_TRY(launch_fwd4(p2, wb(2), L.w4_rp[2], L.w4_kp[2], e2, M, st, kname(d, kF3)));
      }
      // layer 3's statistics are finalised by sa_out_finalize itself (per-block table; cheap since the inline
      // finalisation lost its fp64 division / square root), PCOE_SA_FINALIZE_KERNEL=1 brings the separate launch back
      if (train) { if (sep_fin) PCOE_TRY(finalize(2)); else fin_out = mkfin(2); }
      done = true;
    }
    if (use5) {   // wide layers: (tile x 128-channel block) grid, K streamed in 64-channel chunks
      PCOE_TRY(convert_weights4(d, L, P, wbase, st));
      auto wb = [&](int l) { return (const __nv_bfloat16*)(wbase + L.wb_off[l]); };
      const int Mld = L.Mld;
      v4::StoreStats4 e0{}; e0.y = y[0]; e0.sums = sums[0]; e0.C = d.C1; e0.Mld = Mld;
      v5::GatherFeat5 gp{v4::GatherBase{xyz, new_xyz, nbr, d.N, d.S, d.group_all, M}, feats, d.D};
      PCOE_TRY(launch_fwd5(gp, wb(0), L.w4_kp[0], e0, M, d.C1, gp.nblocks(), st, kname(d, kF1)));
      v4::BnRelu4 p1{}; p1.y = y[0]; p1.scale = scale[0]; p1.shift = shift[0]; p1.M = M; p1.Mld = Mld; p1.C = d.C1;
      p1.fin = mkfin(0);
      v4::StoreStats4 e1{}; e1.y = y[1]; e1.sums = sums[1]; e1.C = d.C2; e1.Mld = Mld;
      PCOE_TRY(launch_fwd5(p1, wb(1), L.w4_kp[1], e1, M, d.C2, d.C1 / 64, st, kname(d, kF2)));
      v4::BnRelu4 p2{}; p2.y = y[1]; p2.scale = scale[1]; p2.shift = shift[1]; p2.M = M; p2.Mld = Mld; p2.C = d.C2;
      p2.fin = mkfin(1);
      v4::Group4 e2{}; e2.y = y[2]; e2.sums = sums[2]; e2.ymax = ymax; e2.ymin = ymin; e2.amax = amax; e2.amin = amin;
      e2.C = d.C3; e2.Mld = Mld; e2.gamma = P.gamma[2];
      PCOE_TRY(launch_fwd5(p2, wb(2), L.w4_kp[2], e2, M, d.C3, d.C2 / 64, st, kname(d, kF3)));
      // layer 3's statistics are finalised by sa_out_finalize itself (per-block table; cheap since the inline
      // finalisation lost its fp64 division / square root), PCOE_SA_FINALIZE_KERNEL=1 brings the separate launch back
      if (train) { if (sep_fin) PCOE_TRY(finalize(2)); else fin_out = mkfin(2); }
      done = true;
    }
  }
  if constexpr (!TC) {
    if (L.v6) {   // bf16x3: split-operand tcgen05 kernels, fp32 activations (sa_tc6.cuh)
      PCOE_TRY(convert_weights6(d, L, P, wbase, st));
      auto wh = [&](int l) { return (const __nv_bfloat16*)(wbase + L.wb_off[l]); };
      auto wps = [&](int l) { return (size_t)L.w4_rp[l] * L.w4_kp[l]; };
      v6::GatherFeat6 gp{v4::GatherBase{xyz, new_xyz, nbr, d.N, d.S, d.group_all, M}, feats, d.D, 0};
      v6::StoreStats6 e0{}; e0.y = y[0]; e0.sums = sums[0]; e0.C = d.C1;
      PCOE_TRY(launch_fwd6(gp, wh(0), wps(0), L.w4_kp[0], e0, M, d.C1, st, kname(d, kF1)));
      v6::BnRelu6 p1{}; p1.y = y[0]; p1.scale = scale[0]; p1.shift = shift[0]; p1.M = M; p1.C = d.C1; p1.fin = mkfin(0);
      v6::StoreStats6 e1{}; e1.y = y[1]; e1.sums = sums[1]; e1.C = d.C2;
      PCOE_TRY(launch_fwd6(p1, wh(1), wps(1), L.w4_kp[1], e1, M, d.C2, st, kname(d, kF2)));
      v6::BnRelu6 p2{}; p2.y = y[1]; p2.scale = scale[1]; p2.shift = shift[1]; p2.M = M; p2.C = d.C2; p2.fin = mkfin(1);
      v6::Group6 e2{}; e2.y = (__nv_bfloat16*)y[2]; e2.sums = sums[2]; e2.ymax = ymax; e2.ymin = ymin; e2.amax = amax; e2.amin = amin;
      e2.C = d.C3; e2.gamma = P.gamma[2];
      PCOE_TRY(launch_fwd6(p2, wh(2), wps(2), L.w4_kp[2], e2, M, d.C3, st, kname(d, kF3)));
      if (train) { if (sep_fin) PCOE_TRY(finalize(2)); else fin_out = mkfin(2); }
      done = true;
    }
  }
  if (!done) {
    GatherProd gp{xyz, new_xyz, nbr, feats, d.N, d.S, d.K, d.D, d.group_all, M, Cin};
    PCOE_TRY(nt(gp, 0, StoreStatsEpi<TY>{y[0], sums[0], d.C1}, kname(d, kF1)));
    if (train) PCOE_TRY(finalize(0));
    BnReluProd<TY> p1{y[0], scale[0], shift[0], M, d.C1};
    PCOE_TRY(nt(p1, 1, StoreStatsEpi<TY>{y[1], sums[1], d.C2}, kname(d, kF2)));
    if (train) PCOE_TRY(finalize(1));
    BnReluProd<TY> p2{y[1], scale[1], shift[1], M, d.C2};
    GroupEpi<TY> ge{y[2], sums[2], ymax, ymin, amax, amin, d.C3, d.K};
    PCOE_TRY(nt(p2, 2, ge, kname(d, kF3)));
    if (train) PCOE_TRY(finalize(2));
  }

  const size_t total = (size_t)G * d.C3;
  int blocks = (int)((total + 255) / 256);
  blocks = blocks > kNumSMs * 8 ? kNumSMs * 8 : blocks;
  LaunchScope ls("sa_out_finalize_kernel", st);
  sa_out_finalize_kernel<<<blocks, 256, fin_out.sums ? sizeof(float) * 2 * d.C3 : 0, st>>>(
      ymax, ymin, amax, amin, scale[2], shift[2], total, d.C3, out, train ? (uint8_t*)(sv + L.sv_slot) : nullptr,
      train ? (float*)(sv + L.sv_ysel) : nullptr, fin_out, train ? P.num_batches_tracked[0] : nullptr,
      train ? P.num_batches_tracked[1] : nullptr, train ? P.num_batches_tracked[2] : nullptr);
  return ls.done();
}

template <typename TY, bool TC>
static int sa_backward_impl(const pcoe_sa_desc& d, const float* xyz, const float* new_xyz, const int32_t* nbr,
                            const float* feats, const pcoe_sa_params& P, const float* out, const float* grad_out,
                            const void* saved, float* grad_feats, const pcoe_sa_grads& Gr, void* workspace,
                            cudaStream_t st) {
  const SaLayout L = sa_layout(d);
  const char* sv = (const char*)saved;
  char* ws = (char*)workspace;
  const int M = L.M, G = L.G, Cin = 3 + d.D;
  const int Cs[3] = {d.C1, d.C2, d.C3};
  const int Kin[3] = {Cin, d.C1, d.C2};

  const TY* y[3];
  const float *scale[3], *shift[3], *mean[3], *invstd[3];
  for (int l = 0; l < 3; ++l) {
    y[l] = (const TY*)(sv + L.sv_y[l]);
    scale[l] = (const float*)(sv + L.sv_stat[l]); shift[l] = scale[l] + Cs[l];
    mean[l] = shift[l] + Cs[l]; invstd[l] = mean[l] + Cs[l];
  }
  const uint8_t* slot = (const uint8_t*)(sv + L.sv_slot);
  const float* ysel = (const float*)(sv + L.sv_ysel);

  double* bs[3];
  float *ca[3], *cp[3], *cq[3];
  for (int l = 0; l < 3; ++l) {
    bs[l] = (double*)(ws + L.wb_sums[l]);
    ca[l] = (float*)(ws + L.wb_consts[l]); cp[l] = ca[l] + Cs[l]; cq[l] = cp[l] + Cs[l];
  }
  float* gm = (float*)(ws + L.wb_gm);
  TY* dz[2] = {(TY*)(ws + L.wb_dz[0]), (TY*)(ws + L.wb_dz[1])};

  bool v4path = false;
  if constexpr (TC) v4path = L.v2;
  // the v4 kernels accumulate dW into copies (dw_combine_kernel writes / adds into dW at the end); the copies follow
  // the batch sums in the workspace: one memset node for both
  // bf16x3, SA1: dW1 comes out of the layer-2 dgrad epilogue (MaskStatsW6); PCOE_SA_BWD_L1_KERNEL=1 = the separate kernel
  bool w1_epi6 = false;
  if constexpr (!TC) {
    const char* keep_env = getenv("PCOE_SA_BWD_L1_KERNEL");
    w1_epi6 = L.v6 && d.D == 0 && !d.group_all && d.C1 == 64 && !(keep_env && keep_env[0] == '1');
  }
  PCOE_CUDA(cudaMemsetAsync(ws + L.wb_sums[0], 0, L.wb_sums_bytes + ((v4path || w1_epi6) ? L.wb_dwc_bytes : 0), st));
  if (v4path) {
  } else if (!Gr.accumulate) {
    for (int l = 0; l < 3; ++l)
      PCOE_CUDA(cudaMemsetAsync(Gr.dW[l], 0, sizeof(float) * (size_t)Cs[l] * Kin[l], st));
  }
  if (d.D > 0 && grad_feats)
    PCOE_CUDA(cudaMemsetAsync(grad_feats, 0, sizeof(float) * (size_t)d.B * d.N * d.D, st));

  auto consts = [&](int l) -> int {
    LaunchScope ls("bn_bwd_consts_kernel", st);
    bn_bwd_consts_kernel<<<ceil_div(Cs[l], 128), 128, 0, st>>>(bs[l], (double)M, Cs[l], scale[l], mean[l],
        invstd[l], ca[l], cp[l], cq[l], Gr.dgamma[l], Gr.dbeta[l], Gr.dbias[l], Gr.accumulate);
    return ls.done();
  };

  // v4 / v5: the BatchNorm-backward constants of layer l are derived inline by the kernels that consume them
  auto mkbfin = [&](int l, int write) {
    v4::BnBwdFin f{};
    f.sums = bs[l]; f.count = (double)M; f.inv_count = 1.0 / (double)M; f.scale = scale[l]; f.mean = mean[l]; f.invstd = invstd[l];
    f.dgamma = Gr.dgamma[l]; f.dbeta = Gr.dbeta[l]; f.dbias = Gr.dbias[l]; f.accumulate = Gr.accumulate; f.write = write;
    return f;
  };
  bool fused_consts = L.v6;
  if constexpr (TC) fused_consts = L.v2 || L.v5;
  bool l3s = false;
  if constexpr (TC) l3s = L.l3s;

  {
    // every block ends with 2*C3 fp64 atomics spread over kRedCopies copies: ~8 blocks per SM keep the per-thread
    // chain of dependent load batches short (it was 17 us at 2 blocks per SM) without long same-address queues
    const int gpb = ceil_div(G, kNumSMs * 8);
    LaunchScope ls("bwd_last_reduce_kernel", st);
    bwd_last_reduce_kernel<<<ceil_div(G, gpb), d.C3 >= 1024 ? 1024 : (d.C3 >= 256 ? 256 : (d.C3 >= 128 ? 128 : 64)), 0, st>>>(grad_out, out, ysel, mean[2], invstd[2], G, d.C3,
                                                            gpb, gm, bs[2], l3s ? scale[2] : nullptr,
                                                            l3s ? (const float*)(sv + L.sv_gram) : nullptr,
                                                            l3s ? (float*)(ws + L.wb_gsum) : nullptr,
                                                            l3s ? d.C2 * L.gram_ld : 0);
    PCOE_TRY(ls.done());
  }
  if (!fused_consts) PCOE_TRY(consts(2));

  // dgrad: dx_prev[M x Kin[l]] = dy_l[M x Cs[l]] * W_l ;  wgrad: dW_l[Cs[l] x Kin[l]] = dy_l^T x_prev
  auto dgrad = [&](const auto& ap, int l, const auto& epi, const char* what) -> int {
    using AP = std::decay_t<decltype(ap)>;
    using EP = std::decay_t<decltype(epi)>;
    if constexpr (TC)
      return launch_nt_tc<AP, EP>(ap, (const __nv_bfloat16*)(sv + L.wbt_off[l]), L.wbt_k[l], L.wbt_rows[l], epi, M,
                                  Kin[l], Cs[l], st, what);
    else
      return launch_nt<AP, EP, true>(ap, P.W[l], Kin[l], epi, M, Kin[l], Cs[l], st, what);
  };
  auto wgrad = [&](const auto& pp, const auto& qp, int l, const char* what) -> int {
    if constexpr (TC) return launch_tn_tc(pp, qp, Gr.dW[l], Kin[l], M, Cs[l], Kin[l], st, what);
    else return launch_tn(pp, qp, Gr.dW[l], Kin[l], M, Cs[l], Kin[l], st, what);
  };

  if constexpr (TC) {
    if (L.v2) {   // one fused wgrad+dgrad kernel per layer (sa_tc4.cuh)
      auto wb = [&](int l) { return (const __nv_bfloat16*)(sv + L.wb_off[l]); };
      auto dwc = [&](int l) { return (float*)(ws + L.wb_dwc[l]); };
      const int Mld = L.Mld;
      v4::BnRelu4 x2{}; x2.y = y[1]; x2.scale = scale[1]; x2.shift = shift[1]; x2.M = M; x2.Mld = Mld; x2.C = d.C2; x2.packed = 1;
      v4::MaskStats4 m2{}; m2.yprev = y[1]; m2.scale = scale[1]; m2.shift = shift[1]; m2.mean = mean[1]; m2.invstd = invstd[1];
      m2.dz = dz[1]; m2.sums = bs[1]; m2.C = d.C2; m2.Mld = Mld;
      v4::DwL3 x3{};
      if (l3s) {
        // y3 was never stored: constants, Gm = W3^T diag(p) W3, r = W3^T q and the summed Gram matrix first, then the
        // backward kernel with the sparse max-pool-routing operand (sa_tc4.cuh, DySparse4)
        __nv_bfloat16* gmimg = (__nv_bfloat16*)(ws + L.wb_gmimg);
        float* rvec = (float*)(ws + L.wb_rvec);
        float* gsum = (float*)(ws + L.wb_gsum);
        float* ext = (float*)(ws + L.wb_l3e);
        {
          const size_t sm = (size_t)2 * d.C3 * L.w4_kp[2] + sizeof(float) * (size_t)(2 * d.C3 + L.gram_ld);
          static bool attr = false;
          if (!attr) { PCOE_CUDA(cudaFuncSetAttribute(v4::l3_prep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024)); attr = true; }
          LaunchScope ls("l3_prep_kernel", st);
          v4::l3_prep_kernel<<<128, 256, sm, st>>>(mkbfin(2, 1), d.C3, d.C2, wb(2), L.w4_kp[2], gsum, L.gram_ld,
                                                   ca[2], cp[2], cq[2], gmimg, rvec, ext);
          PCOE_TRY(ls.done());
        }
        v4::DySparse4 ds{}; ds.gm = gm; ds.slot = slot; ds.M = M; ds.Mld = Mld; ds.C = d.C3;
        m2.addc = rvec;
        PCOE_TRY((launch_bwd4<1, 1>(ds, x2, wb(2), L.w4_rp[2], L.w4_kp[2], m2, dwc(2), L.dwc_ld[2], d.C2, -1, M, d.C2, st,
                                    kname(d, kBL3), gmimg, d.C2)));
        x3 = v4::DwL3{ext};
      } else {
        v4::DyLast4 dy3{}; dy3.gm = gm; dy3.slot = slot; dy3.y = y[2]; dy3.a = ca[2]; dy3.p = cp[2]; dy3.q = cq[2];
        dy3.M = M; dy3.Mld = Mld; dy3.C = d.C3; dy3.fin = mkbfin(2, 1);
        PCOE_TRY(launch_bwd4<1>(dy3, x2, wb(2), L.w4_rp[2], L.w4_kp[2], m2, dwc(2), L.dwc_ld[2], d.C2, -1, M, d.C2, st, kname(d, kBL3)));
      }
      v4::Dy4 dy2{}; dy2.dz = dz[1]; dy2.y = y[1]; dy2.a = ca[1]; dy2.p = cp[1]; dy2.q = cq[1]; dy2.M = M; dy2.Mld = Mld; dy2.C = d.C2;
      dy2.fin = mkbfin(1, 1);
      v4::BnRelu4 x1{}; x1.y = y[0]; x1.scale = scale[0]; x1.shift = shift[0]; x1.M = M; x1.Mld = Mld; x1.C = d.C1; x1.packed = 1;
      v4::MaskStats4 m1{}; m1.yprev = y[0]; m1.scale = scale[0]; m1.shift = shift[0]; m1.mean = mean[0]; m1.invstd = invstd[0];
      m1.dz = dz[0]; m1.sums = bs[0]; m1.C = d.C1; m1.Mld = Mld;
      // no input features (SA1): layer 1 only needs dW1 [C1 x 3] - accumulated by the layer-2 epilogue itself
      // (MaskStatsW1: dz1 is never stored, the layer-1 backward kernel does not run); PCOE_SA_BWD_L1_KERNEL=1 = old path
      const char* keep_env = getenv("PCOE_SA_BWD_L1_KERNEL");   // read per call: the A/B test flips it
      const bool keep_l1 = keep_env && keep_env[0] == '1';
      const bool w1_in_epi = d.D == 0 && d.C1 <= 128 && !d.group_all && !keep_l1;
      v4::DwL1 x1c{};
      if (w1_in_epi) {
        v4::MaskStatsW1 mw{}; mw.yprev = y[0]; mw.scale = scale[0]; mw.shift = shift[0]; mw.mean = mean[0]; mw.invstd = invstd[0];
        mw.sums = bs[0]; mw.C = d.C1; mw.Mld = Mld; mw.gb = v4::GatherBase{xyz, new_xyz, nbr, d.N, d.S, d.group_all, M};
        mw.acc = dwc(0); mw.g0 = (float*)(ws + L.wb_g0);
        PCOE_TRY(launch_bwd4<1>(dy2, x1, wb(1), L.w4_rp[1], L.w4_kp[1], mw, dwc(1), L.dwc_ld[1], d.C1, -1, M, d.C1, st, kname(d, kBL2)));
        x1c = v4::DwL1{dwc(0), (const float*)(ws + L.wb_g0), wb(0), L.w4_kp[0], mkbfin(0, 1)};
      } else {
        PCOE_TRY(launch_bwd4<1>(dy2, x1, wb(1), L.w4_rp[1], L.w4_kp[1], m1, dwc(1), L.dwc_ld[1], d.C1, -1, M, d.C1, st, kname(d, kBL2)));
      }
      v4::Dy4 dy1{}; dy1.dz = dz[0]; dy1.y = y[0]; dy1.a = ca[0]; dy1.p = cp[0]; dy1.q = cq[0]; dy1.M = M; dy1.Mld = Mld; dy1.C = d.C1;
      dy1.fin = mkbfin(0, 1);
      if (w1_in_epi) {
      } else if (d.D == 0) {
        v4::GatherXyz4 x0{v4::GatherBase{xyz, new_xyz, nbr, d.N, d.S, d.group_all, M}};
        PCOE_TRY(launch_bwd4<0>(dy1, x0, wb(0), L.w4_rp[0], L.w4_kp[0], v4::NoEpi4{}, dwc(0), L.dwc_ld[0], Cin, 0, M, 0, st, kname(d, kBL1)));
      } else {
        v4::GatherFeat4 x0{v4::GatherBase{xyz, new_xyz, nbr, d.N, d.S, d.group_all, M}, feats, d.D};
        if (grad_feats) {
          v4::Scatter4 se{grad_feats, nbr, d.N, d.S, d.D, d.group_all};
          PCOE_TRY(launch_bwd4<2>(dy1, x0, wb(0), L.w4_rp[0], L.w4_kp[0], se, dwc(0), L.dwc_ld[0], Cin, d.D, M, d.D, st, kname(d, kBL1)));
        } else {
          PCOE_TRY(launch_bwd4<0>(dy1, x0, wb(0), L.w4_rp[0], L.w4_kp[0], v4::NoEpi4{}, dwc(0), L.dwc_ld[0], Cin, d.D, M, 0, st, kname(d, kBL1)));
        }
      }
      {   // dW_l (+)= sum of the copies
        v4::DwComb cmb[3];
        int total = 0;
        for (int l = 0; l < 3; ++l) {
          cmb[l] = v4::DwComb{dwc(l), Gr.dW[l], Cs[l], Kin[l], L.dwc_ld[l], l == 0 ? d.D : -1};
          total += Cs[l] * Kin[l];
        }
        LaunchScope ls("dw_combine_kernel", st);
        v4::dw_combine_kernel<<<min(ceil_div(total, 256), kNumSMs * 4), 256, 0, st>>>(cmb[0], cmb[1], cmb[2], Gr.accumulate, x3, x1c);
        PCOE_TRY(ls.done());
      }
      return PCOE_OK;
    }
  }

  if constexpr (TC) {
    if (L.v5) {   // wide layers: separate streamed wgrad / dgrad kernels (sa_tc5.cuh)
      auto wb = [&](int l) { return (const __nv_bfloat16*)(sv + L.wb_off[l]); };
      const int Mld = L.Mld;
      v4::DyLast4 dy3{}; dy3.gm = gm; dy3.slot = slot; dy3.y = y[2]; dy3.a = ca[2]; dy3.p = cp[2]; dy3.q = cq[2];
      dy3.M = M; dy3.Mld = Mld; dy3.C = d.C3; dy3.fin = mkbfin(2, 1);   // wgrad owns the parameter-gradient outputs
      v4::BnRelu4 x2{}; x2.y = y[1]; x2.scale = scale[1]; x2.shift = shift[1]; x2.M = M; x2.Mld = Mld; x2.C = d.C2; x2.packed = 1;
      v4::MaskStats4 m2{}; m2.yprev = y[1]; m2.scale = scale[1]; m2.shift = shift[1]; m2.mean = mean[1]; m2.invstd = invstd[1];
      m2.dz = dz[1]; m2.sums = bs[1]; m2.C = d.C2; m2.Mld = Mld;
      // The weight-gradient kernels are off the critical path (only the optimizer consumes them) and none of these
      // grids fills the GPU: they run on the auxiliary stream, each forked after the dgrad that produces its dz.
      AuxStream* ax = aux_stream();
      AuxGuard aux_guard(ax != nullptr);   // released when this call returns (after the join has been enqueued)
      cudaStream_t sw = ax ? ax->s : st;
      auto fork = [&](int e) -> int {
        if (!ax) return PCOE_OK;
        PCOE_CUDA(cudaEventRecord(ax->ev[e], st));
        PCOE_CUDA(cudaStreamWaitEvent(sw, ax->ev[e], 0));
        return PCOE_OK;
      };
      PCOE_TRY(fork(0));
      PCOE_TRY(launch_wgrad5(dy3, x2, Gr.dW[2], d.C2, d.C2, -1, M, d.C2 / 128, sw, kname(d, kWG3)));
      dy3.fin.write = 0;
      PCOE_TRY(launch_dgrad5<false>(dy3, wb(2), L.w4_kp[2], m2, M, d.C2, st, kname(d, kDG3)));
      PCOE_TRY(fork(1));
      v4::Dy4 dy2{}; dy2.dz = dz[1]; dy2.y = y[1]; dy2.a = ca[1]; dy2.p = cp[1]; dy2.q = cq[1]; dy2.M = M; dy2.Mld = Mld; dy2.C = d.C2;
      dy2.fin = mkbfin(1, 1);
      v4::BnRelu4 x1{}; x1.y = y[0]; x1.scale = scale[0]; x1.shift = shift[0]; x1.M = M; x1.Mld = Mld; x1.C = d.C1; x1.packed = 1;
      v4::MaskStats4 m1{}; m1.yprev = y[0]; m1.scale = scale[0]; m1.shift = shift[0]; m1.mean = mean[0]; m1.invstd = invstd[0];
      m1.dz = dz[0]; m1.sums = bs[0]; m1.C = d.C1; m1.Mld = Mld;
      PCOE_TRY(launch_wgrad5(dy2, x1, Gr.dW[1], d.C1, d.C1, -1, M, d.C1 / 128, sw, kname(d, kWG2)));
      dy2.fin.write = 0;
      PCOE_TRY(launch_dgrad5<false>(dy2, wb(1), L.w4_kp[1], m1, M, d.C1, st, kname(d, kDG2)));
      PCOE_TRY(fork(2));
      v4::Dy4 dy1{}; dy1.dz = dz[0]; dy1.y = y[0]; dy1.a = ca[0]; dy1.p = cp[0]; dy1.q = cq[0]; dy1.M = M; dy1.Mld = Mld; dy1.C = d.C1;
      dy1.fin = mkbfin(0, 1);
      v5::GatherFeat5 x0{v4::GatherBase{xyz, new_xyz, nbr, d.N, d.S, d.group_all, M}, feats, d.D};
      PCOE_TRY(launch_wgrad5(dy1, x0, Gr.dW[0], Cin, Cin, d.D, M, ceil_div(x0.nblocks(), 2), sw, kname(d, kWG1)));
      dy1.fin.write = 0;
      if (grad_feats) {
        v4::Scatter4 se{grad_feats, nbr, d.N, d.S, d.D, d.group_all};
        PCOE_TRY(launch_dgrad5<true>(dy1, wb(0), L.w4_kp[0], se, M, d.D, st, kname(d, kDG1)));
      }
      if (ax) {   // join
        PCOE_CUDA(cudaEventRecord(ax->ev[3], sw));
        PCOE_CUDA(cudaStreamWaitEvent(st, ax->ev[3], 0));
      }
      return PCOE_OK;
    }
  }

  if constexpr (!TC) {
    if (L.v6) {   // bf16x3: streamed wgrad / persistent dgrad kernels with split operands (sa_tc6.cuh)
      auto wh = [&](int l) { return (const __nv_bfloat16*)(sv + L.wb_off[l]); };
      auto wps = [&](int l) { return (size_t)L.w4_rp[l] * L.w4_kp[l]; };
      v6::DyLast6 dy3{}; dy3.gm = gm; dy3.slot = slot; dy3.y = (const __nv_bfloat16*)y[2]; dy3.a = ca[2]; dy3.p = cp[2]; dy3.q = cq[2];
      dy3.M = M; dy3.C = d.C3; dy3.fin = mkbfin(2, 1);   // the wgrad launch owns the parameter-gradient outputs
      v6::BnRelu6 x2{}; x2.y = y[1]; x2.scale = scale[1]; x2.shift = shift[1]; x2.M = M; x2.C = d.C2;
      v6::MaskStats6 m2{}; m2.yprev = y[1]; m2.scale = scale[1]; m2.shift = shift[1]; m2.mean = mean[1]; m2.invstd = invstd[1];
      m2.dz = dz[1]; m2.sums = bs[1]; m2.C = d.C2;
      PCOE_TRY(launch_wgrad6(dy3, x2, Gr.dW[2], d.C2, d.C2, -1, M, ceil_div(d.C2, 128), st, kname(d, kWG3)));
      dy3.fin.write = 0;
      PCOE_TRY(launch_dgrad6<false>(dy3, wh(2), wps(2), L.w4_kp[2], m2, M, d.C2, st, kname(d, kDG3)));
      v6::Dy6 dy2{}; dy2.dz = dz[1]; dy2.y = y[1]; dy2.a = ca[1]; dy2.p = cp[1]; dy2.q = cq[1]; dy2.M = M; dy2.C = d.C2;
      dy2.fin = mkbfin(1, 1);
      v6::BnRelu6 x1{}; x1.y = y[0]; x1.scale = scale[0]; x1.shift = shift[0]; x1.M = M; x1.C = d.C1;
      v6::MaskStats6 m1{}; m1.yprev = y[0]; m1.scale = scale[0]; m1.shift = shift[0]; m1.mean = mean[0]; m1.invstd = invstd[0];
      m1.dz = dz[0]; m1.sums = bs[0]; m1.C = d.C1;
      PCOE_TRY(launch_wgrad6(dy2, x1, Gr.dW[1], d.C1, d.C1, -1, M, ceil_div(d.C1, 128), st, kname(d, kWG2)));
      dy2.fin.write = 0;
      if (w1_epi6) {
        v6::MaskStatsW6 mw{}; mw.W1 = P.W[0]; mw.scale = scale[0]; mw.shift = shift[0]; mw.mean = mean[0]; mw.invstd = invstd[0];
        mw.sums = bs[0]; mw.C = d.C1; mw.gb = v4::GatherBase{xyz, new_xyz, nbr, d.N, d.S, d.group_all, M};
        mw.acc = (float*)(ws + L.wb_dwc[0]); mw.g0 = (float*)(ws + L.wb_g0);
        PCOE_TRY(launch_dgrad6<false>(dy2, wh(1), wps(1), L.w4_kp[1], mw, M, d.C1, st, kname(d, kDG2)));
        LaunchScope ls("dw1_finalize_kernel", st);
        v6::dw1_finalize6_kernel<<<ceil_div(d.C1 * 3, 128), 128, 0, st>>>(mw.acc, mw.g0, P.W[0], mkbfin(0, 1), d.C1, Gr.dW[0],
                                                                          Gr.accumulate);
        return ls.done();
      }
      PCOE_TRY(launch_dgrad6<false>(dy2, wh(1), wps(1), L.w4_kp[1], m1, M, d.C1, st, kname(d, kDG2)));
      v6::Dy6 dy1{}; dy1.dz = dz[0]; dy1.y = y[0]; dy1.a = ca[0]; dy1.p = cp[0]; dy1.q = cq[0]; dy1.M = M; dy1.C = d.C1;
      dy1.fin = mkbfin(0, 1);
      v6::GatherFeat6 x0{v4::GatherBase{xyz, new_xyz, nbr, d.N, d.S, d.group_all, M}, feats, d.D, 0};
      PCOE_TRY(launch_wgrad6(dy1, x0, Gr.dW[0], Cin, Cin, d.D, M, ceil_div(x0.nchunks(), 2), st, kname(d, kWG1)));
      dy1.fin.write = 0;
      if (d.D > 0 && grad_feats) {
        v4::Scatter4 se{grad_feats, nbr, d.N, d.S, d.D, d.group_all};
        PCOE_TRY(launch_dgrad6<true>(dy1, wh(0), wps(0), L.w4_kp[0], se, M, d.D, st, kname(d, kDG1)));
      }
      return PCOE_OK;
    }
  }

  // layer 3
  DyLastProd<TY> dy3{gm, slot, y[2], ca[2], cp[2], cq[2], M, d.C3, d.K};
  BnReluProd<TY> x2{y[1], scale[1], shift[1], M, d.C2};
  PCOE_TRY(wgrad(dy3, x2, 2, kname(d, kWG3)));
  MaskStatsEpi<TY> me2{y[1], scale[1], shift[1], mean[1], invstd[1], dz[1], bs[1], d.C2};
  PCOE_TRY(dgrad(dy3, 2, me2, kname(d, kDG3)));
  PCOE_TRY(consts(1));

  // layer 2
  DyProd<TY> dy2{dz[1], y[1], ca[1], cp[1], cq[1], M, d.C2};
  BnReluProd<TY> x1{y[0], scale[0], shift[0], M, d.C1};
  PCOE_TRY(wgrad(dy2, x1, 1, kname(d, kWG2)));
  MaskStatsEpi<TY> me1{y[0], scale[0], shift[0], mean[0], invstd[0], dz[0], bs[0], d.C1};
  PCOE_TRY(dgrad(dy2, 1, me1, kname(d, kDG2)));
  PCOE_TRY(consts(0));

  // layer 1
  DyProd<TY> dy1{dz[0], y[0], ca[0], cp[0], cq[0], M, d.C1};
  GatherProd x0{xyz, new_xyz, nbr, feats, d.N, d.S, d.K, d.D, d.group_all, M, Cin};
  PCOE_TRY(wgrad(dy1, x0, 0, kname(d, kWG1)));
  if (d.D > 0 && grad_feats) {
    ScatterEpi se{grad_feats, nbr, d.N, d.S, d.K, d.D, d.group_all};
    PCOE_TRY(dgrad(dy1, 0, se, kname(d, kDG1)));
  }
  return PCOE_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// Pointwise MLP stack + max over each cloud's points (inference): the conv1d(k=1) + BatchNorm(eval) [+ ReLU] chains of
// the vanilla PointNet (models/pointnet.py: STN3d :22-28, STNkd :53-59, PointNetEncoder :93,102-105).  Runs the
// split-operand tcgen05 forward kernels (sa_tc6.cuh) over all M = clouds * points rows with the identity grouping
// (row = point, blocks of 32 consecutive points pooled by the last layer's epilogue), then reduces the blocks of a cloud.
// ---------------------------------------------------------------------------------------------------------------
// Inference has no gradient routing to protect (the max pool is continuous in the activations): two bf16 planes per operand
// (16 significant bits, 3 MMAs per product, 2^-16 relative per product) instead of the training forward's three planes.
constexpr int kPmPlanes = 2;
struct PmLayout {
  int nl, M, G, Cin0, Kp[3], Rp[3];
  size_t y[2], stat[3], ymax, ymin, amax, amin, wb[3], xin, total;   // y[l]: fp32 activation OR the two-plane operand image (same bytes)
};
static PmLayout pm_layout(const pcoe_pointmlp_desc& d) {
  PmLayout L{};
  L.nl = d.nlayers; L.M = d.M; L.G = d.M / 32;
  L.Cin0 = (d.use_xyz ? 3 : 0) + d.D;
  auto take = [](size_t& cur, size_t bytes) { size_t o = cur; cur = align_up(cur + bytes, 256); return o; };
  size_t s = 0;
  const size_t Mld = align_up((size_t)d.M, 128);
  for (int l = 0; l < d.nlayers; ++l) {
    const int kin = l == 0 ? L.Cin0 : d.C[l - 1];
    L.Rp[l] = (int)align_up(d.C[l], 128);
    L.Kp[l] = (int)align_up(l == 0 ? (d.D + (d.use_xyz ? 16 : 0)) : kin, 128);
    if (l + 1 < d.nlayers) L.y[l] = take(s, Mld * d.C[l] * sizeof(float));
    L.stat[l] = take(s, sizeof(float) * 4 * d.C[l]);
    L.wb[l] = take(s, (size_t)3 * 2 * L.Rp[l] * L.Kp[l]);
  }
  L.xin = d.D > 0 ? take(s, Mld * d.D * sizeof(float)) : 0;       // packed input image of a feature stack (TMA path)
  const int CL = d.C[d.nlayers - 1];
  L.ymax = take(s, sizeof(float) * (size_t)L.G * CL);
  L.ymin = take(s, sizeof(float) * (size_t)L.G * CL);
  L.amax = take(s, (size_t)L.G * CL);
  L.amin = take(s, (size_t)L.G * CL);
  L.total = s;
  return L;
}

// out[b,c] = max over the cloud's blocks of act(scale * (scale >= 0 ? ymax : ymin) + shift)
__global__ void pm_pool_kernel(const float* __restrict__ ymax, const float* __restrict__ ymin,
                               const float* __restrict__ scale, const float* __restrict__ shift, int blocks_per_cloud,
                               int C, int relu, float* __restrict__ out) {
  const int b = blockIdx.y;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float a = scale[c], sh = shift[c];
  const float* src = (!signbit(a) ? ymax : ymin) + (size_t)b * blocks_per_cloud * C + c;
  // affine is monotone per channel: pool the raw extreme first, one FMA at the end
  float ext = __ldg(src);
  if (!signbit(a)) { for (int g = 1; g < blocks_per_cloud; ++g) ext = fmaxf(ext, __ldg(src + (size_t)g * C)); }
  else { for (int g = 1; g < blocks_per_cloud; ++g) ext = fminf(ext, __ldg(src + (size_t)g * C)); }
  const float v = fmaf(ext, a, sh);
  out[(size_t)b * C + c] = relu ? fmaxf(v, 0.f) : v;
}

static int pm_validate(const pcoe_pointmlp_desc* d) {
  if (!d) return fail(PCOE_ERR_NULL, "pointmlp: desc is NULL");
  if (d->M <= 0 || d->rows_per_cloud <= 0 || d->M % d->rows_per_cloud != 0)
    return fail(PCOE_ERR_BAD_SHAPE, "pointmlp: M=%d rows_per_cloud=%d", d->M, d->rows_per_cloud);
  if (d->rows_per_cloud % 32 != 0)
    return fail(PCOE_ERR_UNSUPPORTED, "pointmlp: rows_per_cloud=%d must be a multiple of 32 (pad a cloud by repeating a point)", d->rows_per_cloud);
  if (d->nlayers < 2 || d->nlayers > 3) return fail(PCOE_ERR_UNSUPPORTED, "pointmlp: nlayers=%d (2 or 3)", d->nlayers);
  if (d->D < 0 || d->D % 64 != 0 || (d->D == 0 && !d->use_xyz))
    return fail(PCOE_ERR_UNSUPPORTED, "pointmlp: D=%d must be a multiple of 64 (or 0 with use_xyz)", d->D);
  for (int l = 0; l < d->nlayers; ++l)
    if (d->C[l] <= 0 || d->C[l] % 64 != 0) return fail(PCOE_ERR_UNSUPPORTED, "pointmlp: C[%d]=%d must be a multiple of 64", l, d->C[l]);
  return PCOE_OK;
}

}  // namespace pcoe

using namespace pcoe;

#ifdef PCOE_TC4_TRACE
// debug build only: copy out / reset the in-kernel clock trace of CTA 0 (tag, index, clock64 triples)
extern "C" int pcoe_debug_trace(long long* host_buf, int cap_triples, int reset) {
  int n = 0;
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(&n, v4::g_trace_n, sizeof(int));
  if (n > 2700) n = 2700;
  if (n > cap_triples) n = cap_triples;
  if (host_buf && n > 0) cudaMemcpyFromSymbol(host_buf, v4::g_trace, sizeof(long long) * 3 * n);
  if (reset) { int z = 0; cudaMemcpyToSymbol(v4::g_trace_n, &z, sizeof(int)); }
  return n;
}
#endif

extern "C" size_t pcoe_sa_saved_bytes(const pcoe_sa_desc* desc) {
  if (validate(desc) != PCOE_OK) return 0;
  return sa_layout(*desc).saved_bytes;
}

extern "C" size_t pcoe_sa_workspace_bytes(const pcoe_sa_desc* desc) {
  if (validate(desc) != PCOE_OK) return 0;
  return sa_layout(*desc).workspace_bytes;
}

extern "C" int pcoe_sa_forward(const pcoe_sa_desc* desc, const float* xyz, const float* new_xyz,
                               const int32_t* nbr, const float* feats, const pcoe_sa_params* params,
                               float* out, void* saved, size_t saved_bytes, void* workspace,
                               size_t workspace_bytes, void* stream) {
  PCOE_TRY(validate(desc));
  const pcoe_sa_desc& d = *desc;
  if (!xyz || !params || !out || !workspace) return fail(PCOE_ERR_NULL, "sa_forward: NULL pointer");
  if (!d.group_all && (!new_xyz || !nbr)) return fail(PCOE_ERR_NULL, "sa_forward: new_xyz/nbr is NULL");
  if (d.D > 0 && !feats) return fail(PCOE_ERR_NULL, "sa_forward: feats is NULL but D=%d", d.D);
  for (int l = 0; l < 3; ++l)
    if (!params->W[l] || !params->gamma[l] || !params->beta[l] || !params->running_mean[l] || !params->running_var[l])
      return fail(PCOE_ERR_NULL, "sa_forward: parameter pointer of layer %d is NULL", l + 1);
  const SaLayout L = sa_layout(d);
  if (d.train && (!saved || saved_bytes < L.saved_bytes))
    return fail(PCOE_ERR_WORKSPACE, "sa_forward: saved buffer %zu < %zu bytes", saved_bytes, L.saved_bytes);
  if (workspace_bytes < L.workspace_bytes)
    return fail(PCOE_ERR_WORKSPACE, "sa_forward: workspace %zu < %zu bytes", workspace_bytes, L.workspace_bytes);
  if (d.precision == PCOE_PRECISION_BF16)
    return sa_forward_impl<__nv_bfloat16, true>(d, xyz, new_xyz, nbr, feats, *params, out, saved, workspace, (cudaStream_t)stream);
  return sa_forward_impl<float, false>(d, xyz, new_xyz, nbr, feats, *params, out, saved, workspace, (cudaStream_t)stream);
}

extern "C" int pcoe_sa_backward(const pcoe_sa_desc* desc, const float* xyz, const float* new_xyz,
                                const int32_t* nbr, const float* feats, const pcoe_sa_params* params,
                                const float* out, const float* grad_out, const void* saved,
                                size_t saved_bytes, float* grad_feats, const pcoe_sa_grads* grads,
                                void* workspace, size_t workspace_bytes, void* stream) {
  PCOE_TRY(validate(desc));
  const pcoe_sa_desc& d = *desc;
  if (!d.train) return fail(PCOE_ERR_UNSUPPORTED, "sa_backward: only defined for a train-mode forward");
  if (!xyz || !params || !out || !grad_out || !saved || !grads || !workspace)
    return fail(PCOE_ERR_NULL, "sa_backward: NULL pointer");
  if (!d.group_all && (!new_xyz || !nbr)) return fail(PCOE_ERR_NULL, "sa_backward: new_xyz/nbr is NULL");
  if (d.D > 0 && !feats) return fail(PCOE_ERR_NULL, "sa_backward: feats is NULL but D=%d", d.D);
  for (int l = 0; l < 3; ++l)
    if (!params->W[l] || !grads->dW[l]) return fail(PCOE_ERR_NULL, "sa_backward: W/dW of layer %d is NULL", l + 1);
  const SaLayout L = sa_layout(d);
  if (saved_bytes < L.saved_bytes)
    return fail(PCOE_ERR_WORKSPACE, "sa_backward: saved buffer %zu < %zu bytes", saved_bytes, L.saved_bytes);
  if (workspace_bytes < L.workspace_bytes)
    return fail(PCOE_ERR_WORKSPACE, "sa_backward: workspace %zu < %zu bytes", workspace_bytes, L.workspace_bytes);
  if (d.precision == PCOE_PRECISION_BF16)
    return sa_backward_impl<__nv_bfloat16, true>(d, xyz, new_xyz, nbr, feats, *params, out, grad_out, saved, grad_feats,
                                           *grads, workspace, (cudaStream_t)stream);
  return sa_backward_impl<float, false>(d, xyz, new_xyz, nbr, feats, *params, out, grad_out, saved, grad_feats, *grads,
                                 workspace, (cudaStream_t)stream);
}

// ---- TMA-fed path (pm_tma.cuh): epilogue-produced operand planes, bulk async copies, no producer warps ---------------
template <class Epi, bool CHMAJOR>
static int launch_pm_layer(const uint8_t* ximg, int nk, const __nv_bfloat16* Wp, size_t wps, int Kp, const Epi& epi, int M,
                           int Cout, cudaStream_t st, const char* what) {
  const int ncb = ceil_div(Cout, 128), ntiles = ceil_div(M, v4::kPts);
  const int per = kNumSMs / ncb > 0 ? kNumSMs / ncb : 1;
  const int grid = (ntiles < per ? ntiles : per) * ncb;
  const size_t tb = (size_t)nk * pm::kOpB, fixed = 1024 + tb + epi.stage_bytes();
  if (fixed + tb > kSmemBudget6) return fail(PCOE_ERR_UNSUPPORTED, "%s: layer does not fit shared memory", what);
  int nst = (int)((kSmemBudget6 - fixed) / tb);
  nst = nst > pm::kPmStages ? pm::kPmStages : nst;
  auto k = pm::pm_layer_kernel<Epi, CHMAJOR>;
  static bool attr = false;
  if (!attr) { PCOE_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBudget6)); attr = true; }
  LaunchScope ls(what, st);
  k<<<grid, pm::kPmThreads, fixed + (size_t)nst * tb, st>>>(ximg, nk, Wp, wps, Kp, epi, M, ncb, nst);
  return ls.done();
}

static bool pm_tma_enabled() {
  static const bool on = [] { const char* e = getenv("PCOE_PM_TMA"); return !(e && e[0] == '0'); }();
  return on;
}

static int pointmlp_forward_tma(const pcoe_pointmlp_desc& d, const PmLayout& L, const float* xyz, const float* feats,
                                const pcoe_sa_params& P, float* const* scale, float* const* shift, char* ws, cudaStream_t st) {
  const int nl = d.nlayers, M = d.M, CL = d.C[nl - 1];
  const size_t ntiles = (size_t)ceil_div(M, v4::kPts);
  auto wh = [&](int l) { return (const __nv_bfloat16*)(ws + L.wb[l]); };
  auto wps = [&](int l) { return (size_t)L.Rp[l] * L.Kp[l]; };
  pm::PoolPm pool{(float*)(ws + L.ymax), (float*)(ws + L.ymin), P.gamma[nl - 1], CL, 0, true};
  const uint8_t* cur;            // operand image feeding the next MMA layer
  int cur_nk, first;             // its 64-channel chunks; index of the next MMA layer
  bool chmajor;
  if (d.use_xyz) {               // layer 0 on the CUDA cores (K = 3), writes layer 1's channel-major operand image
    uint8_t* img0 = (uint8_t*)(ws + L.y[0]);
    const size_t threads = ntiles * (size_t)(d.C[0] >> 6) * 1024;
    LaunchScope ls("pointmlp_first_xyz", st);
    pm::pm_first_xyz_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(xyz, M, P.W[0], scale[0], shift[0], d.C[0], 1, img0);
    PCOE_TRY(ls.done());
    cur = img0; cur_nk = d.C[0] >> 6; first = 1; chmajor = true;
  } else {                       // feature stack: pack the point-major rows into a K-major operand image
    uint8_t* xin = (uint8_t*)(ws + L.xin);
    const size_t total = ntiles * v4::kPts * (size_t)(d.D >> 3);
    LaunchScope ls("pointmlp_pack_rows", st);
    pm::pm_pack_rows_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(feats, M, d.D, xin);
    PCOE_TRY(ls.done());
    cur = xin; cur_nk = d.D >> 6; first = 0; chmajor = false;
  }
  for (int l = first; l < nl; ++l) {
    const bool last = l == nl - 1;
    if (!last) {
      pm::StorePlanesPm ep{(uint8_t*)(ws + L.y[l]), scale[l], shift[l], d.C[l], 1, 0, 0.f, 0.f};
      if (chmajor) PCOE_TRY((launch_pm_layer<pm::StorePlanesPm, true>(cur, cur_nk, wh(l), wps(l), L.Kp[l], ep, M, d.C[l], st, l == 1 ? "pointmlp_l2" : "pointmlp_l1")));
      else PCOE_TRY((launch_pm_layer<pm::StorePlanesPm, false>(cur, cur_nk, wh(l), wps(l), L.Kp[l], ep, M, d.C[l], st, l == 1 ? "pointmlp_l2" : "pointmlp_l1")));
      cur = (const uint8_t*)(ws + L.y[l]); cur_nk = d.C[l] >> 6; chmajor = true;
    } else {
      const char* nm = nl == 2 ? "pointmlp_l2_pool" : "pointmlp_l3_pool";
      if (chmajor) PCOE_TRY((launch_pm_layer<pm::PoolPm, true>(cur, cur_nk, wh(l), wps(l), L.Kp[l], pool, M, CL, st, nm)));
      else PCOE_TRY((launch_pm_layer<pm::PoolPm, false>(cur, cur_nk, wh(l), wps(l), L.Kp[l], pool, M, CL, st, nm)));
    }
  }
  return PCOE_OK;
}

extern "C" size_t pcoe_pointmlp_workspace_bytes(const pcoe_pointmlp_desc* desc) {
  if (pm_validate(desc) != PCOE_OK) return 0;
  return pm_layout(*desc).total;
}

extern "C" int pcoe_pointmlp_forward(const pcoe_pointmlp_desc* desc, const float* xyz, const float* feats,
                                     const pcoe_sa_params* params, float* out, void* workspace, size_t workspace_bytes,
                                     void* stream) {
  PCOE_TRY(pm_validate(desc));
  const pcoe_pointmlp_desc& d = *desc;
  if (!params || !out || !workspace) return fail(PCOE_ERR_NULL, "pointmlp: NULL pointer");
  if (d.use_xyz && !xyz) return fail(PCOE_ERR_NULL, "pointmlp: xyz is NULL");
  if (d.D > 0 && !feats) return fail(PCOE_ERR_NULL, "pointmlp: feats is NULL but D=%d", d.D);
  for (int l = 0; l < d.nlayers; ++l)
    if (!params->W[l] || !params->gamma[l] || !params->beta[l] || !params->running_mean[l] || !params->running_var[l])
      return fail(PCOE_ERR_NULL, "pointmlp: parameter pointer of layer %d is NULL", l + 1);
  const PmLayout L = pm_layout(d);
  if (workspace_bytes < L.total) return fail(PCOE_ERR_WORKSPACE, "pointmlp: workspace %zu < %zu bytes", workspace_bytes, L.total);
  cudaStream_t st = (cudaStream_t)stream;
  char* ws = (char*)workspace;
  const int nl = d.nlayers, M = d.M, CL = d.C[nl - 1];
  float *scale[3], *shift[3];
  for (int l = 0; l < nl; ++l) {
    scale[l] = (float*)(ws + L.stat[l]); shift[l] = scale[l] + d.C[l];
    LaunchScope ls("bn_finalize_kernel", st);
    bn_finalize_kernel<<<ceil_div(d.C[l], 128), 128, 0, st>>>(nullptr, 1.0, d.C[l], params->gamma[l], params->beta[l],
        params->bias[l], params->running_mean[l], params->running_var[l], d.eps, 0.f, 0, scale[l], shift[l], nullptr, nullptr);
    PCOE_TRY(ls.done());
  }
  {
    v6::ConvW6 w[3];
    int total = 0;
    for (int l = 0; l < 3; ++l) {
      const int ll = l < nl ? l : nl - 1;            // unused third slot of a 2-layer stack: converts layer 2 again (tiny)
      const int kin = ll == 0 ? L.Cin0 : d.C[ll - 1];
      w[l] = v6::ConvW6{params->W[ll], (__nv_bfloat16*)(ws + L.wb[ll]), d.C[ll], kin, L.Rp[ll], L.Kp[ll], ll == 0 ? d.D : -1,
                        d.use_xyz ? 3 : 0, 0};
      total += L.Rp[ll] * L.Kp[ll];
    }
    LaunchScope ls("convert_weights_kernel", st);
    v6::convert_weights6_kernel<<<min(ceil_div(total / 8, 256), kNumSMs * 4), 256, 0, st>>>(w[0], w[1], w[2]);
    PCOE_TRY(ls.done());
  }
  auto wh = [&](int l) { return (const __nv_bfloat16*)(ws + L.wb[l]); };
  auto wps = [&](int l) { return (size_t)L.Rp[l] * L.Kp[l]; };
  float* ymax = (float*)(ws + L.ymax);
  float* ymin = (float*)(ws + L.ymin);
  // default: epilogue-produced operand planes + bulk async copies (pm_tma.cuh); [xyz | feats] inputs and PCOE_PM_TMA=0
  // take the producer-warp kernels below
  if (pm_tma_enabled() && !(d.use_xyz && d.D > 0) && (!d.use_xyz || d.C[0] % 64 == 0)) {
    PCOE_TRY(pointmlp_forward_tma(d, L, xyz, feats, *params, scale, shift, ws, st));
    LaunchScope lsp("pm_pool_kernel", st);
    pm_pool_kernel<<<dim3(ceil_div(CL, 128), M / d.rows_per_cloud), 128, 0, st>>>(ymax, ymin, scale[nl - 1], shift[nl - 1],
                                                                                d.rows_per_cloud / 32, CL, d.relu_last, out);
    return lsp.done();
  }
  v6::Group6 eg{}; eg.y = nullptr; eg.sums = nullptr; eg.ymax = ymax; eg.ymin = ymin; eg.amax = (uint8_t*)(ws + L.amax);
  eg.amin = (uint8_t*)(ws + L.amin); eg.C = CL; eg.gamma = params->gamma[nl - 1];
  // identity grouping: group_all = 1 makes row r read point r with absolute coordinates
  v6::GatherFeat6 gp{v4::GatherBase{d.use_xyz ? xyz : feats, nullptr, nullptr, d.rows_per_cloud, 1, 1, M}, feats, d.D, d.use_xyz ? 0 : 1};
  float* y0 = (float*)(ws + L.y[0]);
  v6::StoreStats6 e0{}; e0.y = y0; e0.sums = nullptr; e0.C = d.C[0];
  PCOE_TRY(launch_fwd6<kPmPlanes>(gp, wh(0), wps(0), L.Kp[0], e0, M, d.C[0], st, "pointmlp_l1"));
  v6::BnRelu6 p1{}; p1.y = y0; p1.scale = scale[0]; p1.shift = shift[0]; p1.M = M; p1.C = d.C[0];
  if (nl == 2) {
    PCOE_TRY(launch_fwd6<kPmPlanes>(p1, wh(1), wps(1), L.Kp[1], eg, M, d.C[1], st, "pointmlp_l2_pool"));
  } else {
    float* y1 = (float*)(ws + L.y[1]);
    v6::StoreStats6 e1{}; e1.y = y1; e1.sums = nullptr; e1.C = d.C[1];
    PCOE_TRY(launch_fwd6<kPmPlanes>(p1, wh(1), wps(1), L.Kp[1], e1, M, d.C[1], st, "pointmlp_l2"));
    v6::BnRelu6 p2{}; p2.y = y1; p2.scale = scale[1]; p2.shift = shift[1]; p2.M = M; p2.C = d.C[1];
    PCOE_TRY(launch_fwd6<kPmPlanes>(p2, wh(2), wps(2), L.Kp[2], eg, M, d.C[2], st, "pointmlp_l3_pool"));
  }
  LaunchScope ls("pm_pool_kernel", st);
  pm_pool_kernel<<<dim3(ceil_div(CL, 128), M / d.rows_per_cloud), 128, 0, st>>>(ymax, ymin, scale[nl - 1], shift[nl - 1],
                                                                              d.rows_per_cloud / 32, CL, d.relu_last, out);
  return ls.done();
}

// y[m, :] = act(scale * (W x[m, :] + bias) ...) for a narrow input (Cin <= 8): the 3 -> 64 first layer of the PointNet
// encoder, point-major output for the feature-transform bmm.  One thread per (point, 4 output channels).
namespace pcoe {
__global__ void pointwise_linear_kernel(const float* __restrict__ x, const float* __restrict__ W, const float* __restrict__ scale,
                                        const float* __restrict__ shift, int M, int Cin, int Cout, int relu, float* __restrict__ y) {
  const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int q = Cout / 4;
  if (e >= (size_t)M * q) return;
  const int m = (int)(e / q), c0 = (int)(e % q) * 4;
  float xin[8];
  for (int k = 0; k < Cin; ++k) xin[k] = __ldg(x + (size_t)m * Cin + k);
  float r[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    float acc = 0.f;
    for (int k = 0; k < Cin; ++k) acc = fmaf(__ldg(W + (size_t)(c0 + u) * Cin + k), xin[k], acc);
    acc = fmaf(acc, scale[c0 + u], shift[c0 + u]);
    r[u] = relu ? fmaxf(acc, 0.f) : acc;
  }
  *reinterpret_cast<float4*>(y + (size_t)m * Cout + c0) = make_float4(r[0], r[1], r[2], r[3]);
}
}  // namespace pcoe

extern "C" int pcoe_pointwise_linear_f32(const float* x, int M, int Cin, const float* W, const float* scale, const float* shift,
                                         int Cout, int relu, float* y, void* stream) {
  if (M <= 0 || Cin <= 0 || Cout <= 0) return fail(PCOE_ERR_BAD_SHAPE, "pointwise_linear: M=%d Cin=%d Cout=%d", M, Cin, Cout);
  if (Cin > 8 || Cout % 4 != 0) return fail(PCOE_ERR_UNSUPPORTED, "pointwise_linear: Cin=%d (<= 8), Cout=%d (multiple of 4)", Cin, Cout);
  if (!x || !W || !scale || !shift || !y) return fail(PCOE_ERR_NULL, "pointwise_linear: NULL pointer");
  const size_t total = (size_t)M * (Cout / 4);
  LaunchScope ls("pointwise_linear_kernel", (cudaStream_t)stream);
  pointwise_linear_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, W, scale, shift, M, Cin, Cout, relu, y);
  return ls.done();
}
