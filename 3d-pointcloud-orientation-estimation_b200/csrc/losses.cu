// losses.cu — fused value+gradient kernels for the three training losses of the reference.
//   pcoe_vm_kl_fwd_bwd      train_single_peak_vonMises_KL.py:23-28 / train_multi_peaks_vonMises_KL.py:38-52
//   pcoe_mvm_match_fwd_bwd  train_multi_peaks_vonMises_KL.py:54-81 (+ scipy linear_sum_assignment)
//   pcoe_soft_ce_fwd_bwd    train_8dir_KL.py:60-68
// The tensors are tiny (B x <=16 values): one thread (one warp for the matched loss) per sample,
// arithmetic in fp64 so the result
// is at least as accurate as the reference's fp32 Cephes evaluation; the fp32 overflow behaviour
// of torch.special.i0/i1 (exp(x) = inf for x > log(FLT_MAX)) is reproduced explicitly.
#include "common.cuh"
#include "mvm_head.cuh"
#include <math.h>

namespace pcoe {

constexpr double kPi = 3.14159265358979323846;
constexpr double kF32ExpOverflow = 88.72283905206835;  // log(FLT_MAX)

// Exponentially scaled modified Bessel functions i0e(x)=exp(-|x|) I0(x), i1e(x)=exp(-|x|) I1(x).
// |x| <= 20: ascending power series (all terms positive); otherwise Hankel asymptotic expansion.
__device__ void bessel_i0e_i1e(double x, double* i0e, double* i1e) {
  const double ax = fabs(x);
  double r0, r1;
  if (ax <= 20.0) {
    const double q = 0.25 * ax * ax;
    double t0 = 1.0, s0 = 1.0, t1 = 1.0, s1 = 1.0;
    for (int k = 1; k < 200; ++k) {
      t0 *= q / ((double)k * k);
      t1 *= q / ((double)k * (k + 1));
      s0 += t0;
      s1 += t1;
      if (t0 < 1e-17 * s0) break;
    }
    const double e = exp(-ax);
    r0 = s0 * e;
    r1 = 0.5 * ax * s1 * e;
  } else {
    const double pref = 1.0 / sqrt(2.0 * kPi * ax);
    double t0 = 1.0, s0 = 1.0, t1 = 1.0, s1 = 1.0;
    for (int k = 1; k < 60; ++k) {
      const double o = (2.0 * k - 1.0) * (2.0 * k - 1.0);
      const double n0 = t0 * o / (8.0 * k * ax);            // mu = 0
      const double n1 = t1 * -(4.0 - o) / (8.0 * k * ax);   // mu = 4
      if (fabs(n0) > fabs(t0) || fabs(n0) < 1e-17) break;
      t0 = n0; t1 = n1;
      s0 += t0; s1 += t1;
    }
    r0 = pref * s0;
    r1 = pref * s1;
  }
  *i0e = r0;
  *i1e = x < 0 ? -r1 : r1;
}

struct VmTerm {   // per-kappa quantities
  double k;       // kappa actually used (after clamp)
  double logi0;   // log I0(k)   (inf when the fp32 reference overflows)
  double A;       // I1/I0       (NaN when the fp32 reference overflows)
  double dA;      // dA/dk
  bool pass;      // gradient passes the clamp
};

__device__ VmTerm vm_term(float kappa_f, bool clamp) {
  VmTerm t;
  float kf = kappa_f;
  t.pass = true;
  if (clamp) {
    const float lo = 1e-6f, hi = 500.0f;
    t.pass = (kf >= lo) && (kf <= hi);
    kf = fminf(fmaxf(kf, lo), hi);
  }
  t.k = (double)kf;
  double i0e, i1e;
  bessel_i0e_i1e(t.k, &i0e, &i1e);
  if (fabs(t.k) > kF32ExpOverflow) {  // torch fp32: i0 = i1 = inf
    t.logi0 = INFINITY;
    t.A = NAN;
    t.dA = NAN;
  } else {
    t.logi0 = fabs(t.k) + log(i0e);
    t.A = i1e / i0e;
    t.dA = (t.k != 0.0) ? 1.0 - t.A / t.k - t.A * t.A : 0.5;
  }
  return t;
}

__device__ __forceinline__ double wrap_pi(double d) {
  return d - 2.0 * kPi * floor((d + kPi) / (2.0 * kPi));  // (d + pi) % (2 pi) - pi, Python modulo
}

// KL of one (p, q) pair and its gradient w.r.t. (mu_p, kappa_p).
__device__ void vm_kl_pair(float mu_p, const VmTerm& p, float mu_q, const VmTerm& q, bool multi,
                           float kappa_p_raw, double* kl, double* dmu, double* dk) {
  double delta = (double)mu_p - (double)mu_q;
  if (multi) delta = wrap_pi(delta);
  const double c = cos(delta), s = sin(delta);
  if (multi) {
    // log(i0_q / i0_p) + A_p * (kappa_p - kappa_q * cos(delta))
    double lr;
    if (isinf(q.logi0) && isinf(p.logi0)) lr = NAN;          // log(inf/inf)
    else if (isinf(p.logi0)) lr = -INFINITY;                 // log(x/inf) = log 0
    else lr = q.logi0 - p.logi0;
    *kl = lr + p.A * (p.k - q.k * c);
    *dmu = p.A * q.k * s;
    *dk = p.pass ? p.dA * (p.k - q.k * c) : 0.0;
  } else {
    // log(i0_q) - log(i0_p) + kappa_p*a1 - kappa_q*a1*cos(delta), a1 = 0 for kappa_p <= 1e-6
    const bool small = kappa_p_raw <= 1e-6f;
    const double a1 = small ? 0.0 : p.A;
    *kl = q.logi0 - p.logi0 + p.k * a1 - q.k * a1 * c;
    *dmu = q.k * a1 * s;
    *dk = small ? -p.A : p.dA * (p.k - q.k * c);
  }
}

__global__ void vm_kl_kernel(const float* __restrict__ mu_p, const float* __restrict__ kappa_p,
                             const float* __restrict__ mu_q, const float* __restrict__ kappa_q,
                             int n, int variant, float* __restrict__ loss, float* __restrict__ dmu,
                             float* __restrict__ dkappa) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const bool multi = variant == PCOE_VM_MULTI;
  const VmTerm p = vm_term(kappa_p[i], multi), q = vm_term(kappa_q[i], multi);
  double kl, gm, gk;
  vm_kl_pair(mu_p[i], p, mu_q[i], q, multi, kappa_p[i], &kl, &gm, &gk);
  loss[i] = (float)kl;
  if (dmu) dmu[i] = (float)gm;
  if (dkappa) dkappa[i] = (float)gk;
}

__device__ __forceinline__ double shfl_d(double v, int src) {
  return __hiloint2double(__shfl_sync(0xFFFFFFFFu, __double2hiint(v), src), __shfl_sync(0xFFFFFFFFu, __double2loint(v), src));
}
__device__ __forceinline__ double shfl_xor_d(double v, int m) {
  return __hiloint2double(__shfl_xor_sync(0xFFFFFFFFu, __double2hiint(v), m), __shfl_xor_sync(0xFFFFFFFFu, __double2loint(v), m));
}

// One warp per sample: lane 4*i + j (< 16) evaluates the cost of matching predicted component i to
// ground-truth component j (two Bessel evaluations and one KL per lane instead of 8 + 16 per
// thread), the <= 256 assignment codes are scored 32 at a time with warp shuffles.
__global__ void mvm_match_kernel(const float* __restrict__ mu, const float* __restrict__ kappa,
                                 const float* __restrict__ w, const float* __restrict__ gt,
                                 int gt_stride, const int32_t* __restrict__ K_gt, int B, int Kmax,
                                 float* __restrict__ loss, float* __restrict__ dmu,
                                 float* __restrict__ dkappa, float* __restrict__ dw,
                                 int32_t* __restrict__ perm) {
  const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (b >= B) return;                       // whole warps only
  int K = K_gt[b];
  K = K > Kmax ? Kmax : K;
  if (K <= 0) {
    if (lane < Kmax) {
      if (dmu) dmu[b * Kmax + lane] = 0.f;
      if (dkappa) dkappa[b * Kmax + lane] = 0.f;
      if (dw) dw[b * Kmax + lane] = 0.f;
      if (perm) perm[b * Kmax + lane] = -1;
    }
    if (lane == 0) loss[b] = 0.f;
    return;
  }

  const int ci = (lane >> 2) & 3, cj = lane & 3;
  double kl = 0.0, ga = 0.0, gc = 0.0;
  if (lane < 16 && ci < K && cj < K) {
    const VmTerm tp = vm_term(kappa[b * Kmax + ci], true);
    const VmTerm tq = vm_term(gt[((size_t)b * Kmax + cj) * gt_stride + 1], true);
    vm_kl_pair(mu[b * Kmax + ci], tp, gt[((size_t)b * Kmax + cj) * gt_stride], tq, true, 0.f, &kl, &ga, &gc);
    const float klf = (float)kl;  // the reference stores the cost matrix in fp32 before nan_to_num
    if (isnan(klf) || isinf(klf)) {
      // nan_to_num(1e6): autograd multiplies the (non-finite) local derivatives by 0, which is
      // NaN wherever fp32 I0 overflowed - reproduced so that gradients match the reference.
      kl = 1e6;
      ga = isinf(tp.logi0) ? NAN : 0.0;
      gc = NAN;
    }
  }
  // minimum-cost assignment: brute force over the K! permutations (K <= 4).  Codes are base-4
  // digit strings, element 0 most significant (lexicographic order); the lowest code among the
  // minima wins, as in a sequential first-minimum scan.
  double best = INFINITY;
  int best_code = 0x7FFFFFFF;
  const int ncode = 1 << (2 * K);
  for (int base = 0; base < ncode; base += 32) {
    const int code = base + lane;
    int pj[4] = {0, 0, 0, 0}, used = 0;
    bool ok = code < ncode;
    for (int i = 0; i < K; ++i) {
      const int j = (code >> (2 * (K - 1 - i))) & 3;
      pj[i] = j;
      ok = ok && j < K && !((used >> j) & 1);
      used |= 1 << j;
    }
    double tot = 0.0;
    for (int i = 0; i < K; ++i) tot += shfl_d(kl, i * 4 + (ok ? pj[i] : 0));
    if (ok && tot < best) { best = tot; best_code = code; }
  }
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) {
    const double ob = shfl_xor_d(best, m);
    const int oc = __shfl_xor_sync(0xFFFFFFFFu, best_code, m);
    if (ob < best || (ob == best && oc < best_code)) { best = ob; best_code = oc; }
  }
  // lane i < K: predicted component i is matched to ground-truth component j
  const bool mine = lane < K;
  const int j = mine ? (best_code >> (2 * (K - 1 - lane))) & 3 : 0;
  const int src = mine ? lane * 4 + j : 0;
  const double cs = shfl_d(kl, src), gms = shfl_d(ga, src), gks = shfl_d(gc, src);
  const double wi = mine ? (double)w[b * Kmax + lane] : 0.0;
  double wsum = wi, num = mine ? wi * cs : 0.0;
#pragma unroll
  for (int m = 2; m >= 1; m >>= 1) { wsum += shfl_xor_d(wsum, m); num += shfl_xor_d(num, m); }   // lanes 0..3
  const double W = wsum + 1e-8;
  const double L = num / W;
  if (lane == 0) loss[b] = (float)L;
  if (lane < Kmax) {
    if (dmu) dmu[b * Kmax + lane] = mine ? (float)(wi / W * gms) : 0.f;
    if (dkappa) dkappa[b * Kmax + lane] = mine ? (float)(wi / W * gks) : 0.f;
    if (dw) dw[b * Kmax + lane] = mine ? (float)((cs - L) / W) : 0.f;
    if (perm) perm[b * Kmax + lane] = mine ? j : -1;
  }
}

__global__ void soft_ce_kernel(const float* __restrict__ logits, const float* __restrict__ p, int B,
                               int C, float* __restrict__ loss, float* __restrict__ dlogits) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const float* z = logits + (size_t)b * C;
  const float* pb = p + (size_t)b * C;
  double m = -INFINITY;
  for (int j = 0; j < C; ++j) m = fmax(m, (double)z[j]);
  double se = 0.0, ps = 0.0;
  for (int j = 0; j < C; ++j) { se += exp((double)z[j] - m); ps += (double)pb[j]; }
  const double lse = m + log(se);
  double L = 0.0;
  for (int j = 0; j < C; ++j) {
    L -= (double)pb[j] * ((double)z[j] - lse);
    if (dlogits) dlogits[(size_t)b * C + j] = (float)(exp((double)z[j] - lse) * ps - (double)pb[j]);
  }
  loss[b] = (float)L;
}

}  // namespace pcoe

using namespace pcoe;

extern "C" int pcoe_vm_kl_fwd_bwd(const float* mu_p, const float* kappa_p, const float* mu_q,
                                  const float* kappa_q, int n, int variant, float* loss,
                                  float* dmu, float* dkappa, void* stream) {
  if (n < 0) return fail(PCOE_ERR_BAD_SHAPE, "vm_kl: n=%d", n);
  if (variant != PCOE_VM_SINGLE && variant != PCOE_VM_MULTI)
    return fail(PCOE_ERR_UNSUPPORTED, "vm_kl: variant=%d", variant);
  if (n == 0) return PCOE_OK;
  if (!mu_p || !kappa_p || !mu_q || !kappa_q || !loss) return fail(PCOE_ERR_NULL, "vm_kl: NULL pointer");
  LaunchScope ls("vm_kl_kernel", (cudaStream_t)stream);
  vm_kl_kernel<<<ceil_div(n, 128), 128, 0, (cudaStream_t)stream>>>(mu_p, kappa_p, mu_q, kappa_q, n,
                                                                   variant, loss, dmu, dkappa);
  return ls.done();
}

extern "C" int pcoe_mvm_match_fwd_bwd(const float* mu, const float* kappa, const float* w,
                                      const float* gt, int gt_stride, const int32_t* K_gt, int B,
                                      int Kmax, float* loss, float* dmu, float* dkappa, float* dw,
                                      int32_t* perm, void* stream) {
  if (B < 0 || Kmax <= 0 || gt_stride < 2)
    return fail(PCOE_ERR_BAD_SHAPE, "mvm_match: B=%d Kmax=%d gt_stride=%d", B, Kmax, gt_stride);
  if (Kmax > 4) return fail(PCOE_ERR_UNSUPPORTED, "mvm_match: Kmax=%d > 4", Kmax);
  if (B == 0) return PCOE_OK;
  if (!mu || !kappa || !w || !gt || !K_gt || !loss) return fail(PCOE_ERR_NULL, "mvm_match: NULL pointer");
  LaunchScope ls("mvm_match_kernel", (cudaStream_t)stream);
  mvm_match_kernel<<<ceil_div(B, 4), 128, 0, (cudaStream_t)stream>>>(mu, kappa, w, gt, gt_stride, K_gt,
                                                                     B, Kmax, loss, dmu, dkappa, dw, perm);
  return ls.done();
}

// ---- mixture-of-von-Mises head transform (models/pointnet_pp_mvM.py:91-125), forward and backward ----
// One thread per sample.  The reference spells this as ~20 elementwise torch ops (and ~35 in backward);
// here: weight = softmax(pi / temp); (c, s) = normalize(mu_raw pair, eps 1e-4), fallback (1, 0) when the
// normalised vector is shorter than 1e-3, mu = atan2(s, c); kappa = min(softplus(kappa_raw) + 1e-6, kappa_max).
namespace pcoe {

__global__ void mvm_head_fwd_kernel(const float* __restrict__ pi, const float* __restrict__ mu_raw,
                                    const float* __restrict__ kappa_raw, int B, int K, float temp, float kappa_max,
                                    int clamp_kappa, float* __restrict__ weight, float* __restrict__ mu,
                                    float* __restrict__ kappa) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  mvm_head_fwd_row(pi + b * K, mu_raw + b * K * 2, kappa_raw + b * K, K, temp, kappa_max, clamp_kappa, weight + b * K,
                   mu + b * K, kappa + b * K);
}

__global__ void mvm_head_bwd_kernel(const float* __restrict__ pi, const float* __restrict__ mu_raw,
                                    const float* __restrict__ kappa_raw, int B, int K, float temp, float kappa_max,
                                    int clamp_kappa, const float* __restrict__ g_w, const float* __restrict__ g_mu,
                                    const float* __restrict__ g_k, float* __restrict__ d_pi,
                                    float* __restrict__ d_mu_raw, float* __restrict__ d_kappa_raw) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  mvm_head_bwd_row(pi + b * K, mu_raw + b * K * 2, kappa_raw + b * K, K, temp, kappa_max, clamp_kappa,
                   g_w ? g_w + b * K : nullptr, g_mu ? g_mu + b * K : nullptr, g_k ? g_k + b * K : nullptr, d_pi + b * K,
                   d_mu_raw + b * K * 2, d_kappa_raw + b * K);
}

}  // namespace pcoe

extern "C" int pcoe_mvm_head_fwd(const float* pi, const float* mu_raw, const float* kappa_raw, int B, int K,
                                 float temp, float kappa_max, int clamp_kappa, float* weight, float* mu, float* kappa,
                                 void* stream) {
  if (B < 0 || K <= 0 || !(temp > 0.f)) return fail(PCOE_ERR_BAD_SHAPE, "mvm_head: B=%d K=%d temp=%g", B, K, temp);
  if (K > kHeadMaxK) return fail(PCOE_ERR_UNSUPPORTED, "mvm_head: K=%d > %d", K, kHeadMaxK);
  if (B == 0) return PCOE_OK;
  if (!pi || !mu_raw || !kappa_raw || !weight || !mu || !kappa) return fail(PCOE_ERR_NULL, "mvm_head: NULL pointer");
  LaunchScope ls("mvm_head_fwd_kernel", (cudaStream_t)stream);
  mvm_head_fwd_kernel<<<ceil_div(B, 64), 64, 0, (cudaStream_t)stream>>>(pi, mu_raw, kappa_raw, B, K, temp, kappa_max,
                                                                      clamp_kappa, weight, mu, kappa);
  return ls.done();
}

extern "C" int pcoe_mvm_head_bwd(const float* pi, const float* mu_raw, const float* kappa_raw, int B, int K,
                                 float temp, float kappa_max, int clamp_kappa, const float* g_weight,
                                 const float* g_mu, const float* g_kappa, float* d_pi, float* d_mu_raw,
                                 float* d_kappa_raw, void* stream) {
  if (B < 0 || K <= 0 || !(temp > 0.f)) return fail(PCOE_ERR_BAD_SHAPE, "mvm_head: B=%d K=%d temp=%g", B, K, temp);
  if (K > kHeadMaxK) return fail(PCOE_ERR_UNSUPPORTED, "mvm_head: K=%d > %d", K, kHeadMaxK);
  if (B == 0) return PCOE_OK;
  if (!pi || !mu_raw || !kappa_raw || !d_pi || !d_mu_raw || !d_kappa_raw) return fail(PCOE_ERR_NULL, "mvm_head: NULL pointer");
  LaunchScope ls("mvm_head_bwd_kernel", (cudaStream_t)stream);
  mvm_head_bwd_kernel<<<ceil_div(B, 64), 64, 0, (cudaStream_t)stream>>>(pi, mu_raw, kappa_raw, B, K, temp, kappa_max,
                                                                      clamp_kappa, g_weight, g_mu, g_kappa, d_pi,
                                                                      d_mu_raw, d_kappa_raw);
  return ls.done();
}

extern "C" int pcoe_soft_ce_fwd_bwd(const float* logits, const float* p, int B, int C, float* loss,
                                    float* dlogits, void* stream) {
  if (B < 0 || C <= 0) return fail(PCOE_ERR_BAD_SHAPE, "soft_ce: B=%d C=%d", B, C);
  if (C > 64) return fail(PCOE_ERR_UNSUPPORTED, "soft_ce: C=%d > 64", C);
  if (B == 0) return PCOE_OK;
  if (!logits || !p || !loss) return fail(PCOE_ERR_NULL, "soft_ce: NULL pointer");
  LaunchScope ls("soft_ce_kernel", (cudaStream_t)stream);
  soft_ce_kernel<<<ceil_div(B, 128), 128, 0, (cudaStream_t)stream>>>(logits, p, B, C, loss, dlogits);
  return ls.done();
}
