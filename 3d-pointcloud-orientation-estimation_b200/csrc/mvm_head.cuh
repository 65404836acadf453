// mvm_head.cuh — the mixture-of-von-Mises head transform of ONE sample, shared by the stand-alone kernels
// (losses.cu: pcoe_mvm_head_fwd / _bwd) and the fused trunk tail (trunk.cu).
//   (pi, mu_raw, kappa_raw) -> weight = softmax(pi / temp), mu = atan2 of the eps-normalised 2-vector (with the
//   reference's where-fallback), kappa = softplus(kappa_raw) + 1e-6 clamped to kappa_max.
// Reference: models/pointnet_pp_mvM.py:91-125.
#pragma once
#include <math.h>

namespace pcoe {

constexpr int kHeadMaxK = 8;

__device__ __forceinline__ float softplus_f(float x) { return x > 20.f ? x : log1pf(expf(x)); }   // torch threshold 20

// pi [K], mu_raw [2K], kappa_raw [K] -> weight [K], mu [K], kappa [K]
__device__ __forceinline__ void mvm_head_fwd_row(const float* pi, const float* mu_raw, const float* kappa_raw, int K,
                                                 float temp, float kappa_max, int clamp_kappa, float* weight, float* mu,
                                                 float* kappa) {
  float z[kHeadMaxK], mx = -INFINITY, sum = 0.f;
  for (int k = 0; k < K; ++k) { z[k] = pi[k] / temp; mx = fmaxf(mx, z[k]); }
  for (int k = 0; k < K; ++k) { z[k] = expf(z[k] - mx); sum += z[k]; }
  for (int k = 0; k < K; ++k) {
    weight[k] = z[k] / sum;
    const float vx = mu_raw[k * 2], vy = mu_raw[k * 2 + 1];
    const float dn = fmaxf(sqrtf(vx * vx + vy * vy), 1e-4f);
    const float c = vx / dn, s = vy / dn;
    const bool masked = sqrtf(c * c + s * s) < 1e-3f;
    mu[k] = masked ? 0.f : atan2f(s, c);
    float kp = softplus_f(kappa_raw[k]) + 1e-6f;
    if (clamp_kappa) kp = fminf(kp, kappa_max);
    kappa[k] = kp;
  }
}

// gradients w.r.t. (pi, mu_raw, kappa_raw) from the upstream (g_w, g_mu, g_k) of one sample; NULL = no gradient
__device__ __forceinline__ void mvm_head_bwd_row(const float* pi, const float* mu_raw, const float* kappa_raw, int K,
                                                 float temp, float kappa_max, int clamp_kappa, const float* g_w,
                                                 const float* g_mu, const float* g_k, float* d_pi, float* d_mu_raw,
                                                 float* d_kappa_raw) {
  float z[kHeadMaxK], mx = -INFINITY, sum = 0.f, dot = 0.f;
  for (int k = 0; k < K; ++k) { z[k] = pi[k] / temp; mx = fmaxf(mx, z[k]); }
  for (int k = 0; k < K; ++k) { z[k] = expf(z[k] - mx); sum += z[k]; }
  for (int k = 0; k < K; ++k) { z[k] /= sum; dot += (g_w ? g_w[k] : 0.f) * z[k]; }
  for (int k = 0; k < K; ++k) {
    d_pi[k] = g_w ? z[k] * (g_w[k] - dot) / temp : 0.f;
    // mu: atan2 -> where-fallback -> normalize(eps)
    const float vx = mu_raw[k * 2], vy = mu_raw[k * 2 + 1];
    const float n = sqrtf(vx * vx + vy * vy), dn = fmaxf(n, 1e-4f);
    const float c = vx / dn, s = vy / dn, r2 = c * c + s * s;
    float gx = 0.f, gy = 0.f;
    if (g_mu && !(sqrtf(r2) < 1e-3f)) {
      const float g = g_mu[k];
      const float gc = -s / r2 * g, gs = c / r2 * g;           // d atan2(s, c)
      if (n > 1e-4f) {                                         // u = v / |v|
        const float proj = (vx * gc + vy * gs) / (n * n * n);
        gx = gc / n - vx * proj;
        gy = gs / n - vy * proj;
      } else {                                                 // u = v / eps (clamp_min passes no gradient to |v|)
        gx = gc / 1e-4f;
        gy = gs / 1e-4f;
      }
    }
    d_mu_raw[k * 2] = gx;
    d_mu_raw[k * 2 + 1] = gy;
    const float x = kappa_raw[k];
    float gk = g_k ? g_k[k] : 0.f;
    if (clamp_kappa && softplus_f(x) + 1e-6f > kappa_max) gk = 0.f;
    d_kappa_raw[k] = gk * (x > 20.f ? 1.f : 1.f / (1.f + expf(-x)));
  }
}

}  // namespace pcoe
