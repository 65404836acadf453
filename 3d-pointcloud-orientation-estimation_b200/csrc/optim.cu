// optim.cu — pcoe_adam_step: gradient-norm clipping + Adam over ONE flat fp32 parameter buffer.
//
// Replaces, per training step of the reference (train_multi_peaks_vonMises_KL.py:235-236,
// train_single_peak_vonMises_KL.py:90, train_8dir_KL.py:97):
//     torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm)    (mvM script only)
//     optimizer.step()            torch.optim.Adam(lr=1e-3), betas (0.9, 0.999), eps 1e-8
//     optimizer.zero_grad()       (optionally folded in: the gradient is cleared as it is consumed)
// which PyTorch runs as ~14 multi-tensor launches over ~60 small tensors.  Parameters, gradients
// and both moment buffers are flat arrays (pcoe.dp.FlatGradBuffer / pcoe.optim.FusedAdam), so the
// whole update is two launches that stream 7 x 4 bytes per parameter once:
//   1. grad_sumsq_kernel   per-block partial sums of g^2 (fp64), the last block to finish adds the
//                          partials in a fixed order (deterministic), writes ||g||_2 and advances
//                          the device-side step counter (CUDA-graph replays keep counting);
//   2. adam_update_kernel  clip coefficient min(1, max_norm / (||g|| + 1e-6)), moment updates,
//                          bias-corrected parameter update, float4 vectorised.
#include "common.cuh"

namespace pcoe {

constexpr int kOptBlocks = kNumSMs * 2, kOptThreads = 256;

struct AdamWs {              // layout of the workspace
  double partial[kOptBlocks];
  unsigned int ticket;
};

__global__ void __launch_bounds__(kOptThreads)
grad_sumsq_kernel(const float* __restrict__ g, size_t n, AdamWs* __restrict__ ws, float* __restrict__ norm_out,
                  long long* __restrict__ step) {
  __shared__ double red[kOptThreads / 32];
  __shared__ bool last;
  // fp32 partial sums of 4 squares per load, promoted to fp64 per load: ~22 bits of headroom over 1.5 M elements
  double acc = 0.0;
  const size_t n4 = n >> 2, stride = (size_t)gridDim.x * blockDim.x;
  const float4* g4 = reinterpret_cast<const float4*>(g);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 v = __ldg(g4 + i);
    acc += (double)((v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w));
  }
  if (blockIdx.x == 0)
    for (size_t i = (n4 << 2) + threadIdx.x; i < n; i += blockDim.x) acc += (double)g[i] * (double)g[i];
  auto block_sum = [&](double v) -> double {      // fixed tree: deterministic
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, m);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
    for (int w = 0; w < kOptThreads / 32; ++w) t += red[w];
    return t;
  };
  const double bs = block_sum(acc);
  if (threadIdx.x == 0) {
    ws->partial[blockIdx.x] = bs;
    __threadfence();
    last = atomicAdd(&ws->ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (last) {                                     // the last block adds the per-block partials, all threads helping
    __threadfence();
    double v = 0.0;
    for (unsigned b = threadIdx.x; b < gridDim.x; b += blockDim.x) v += ((volatile double*)ws->partial)[b];
    const double tot = block_sum(v);
    if (threadIdx.x == 0) {
      *norm_out = (float)sqrt(tot);
      *step += 1;
      ws->ticket = 0;   // self-cleaning: ready for the next launch / graph replay
    }
  }
}

struct AdamArgs {
  float lr, beta1, beta2, eps, weight_decay, max_norm, grad_scale;
  int zero_grad;
};

__device__ __forceinline__ float adam_one(float& p, float g, float& m, float& v, const AdamArgs& a, float coef,
                                          float step_size, float inv_sqrt_bc2) {
  g *= coef;
  const float clipped = g;
  if (a.weight_decay != 0.f) g = fmaf(a.weight_decay, p, g);
  m = fmaf(g - m, 1.f - a.beta1, m);                       // exp_avg.lerp_(grad, 1 - beta1)
  v = fmaf(g * g, 1.f - a.beta2, a.beta2 * v);             // exp_avg_sq.mul_(beta2).addcmul_(g, g, 1 - beta2)
  const float denom = sqrtf(v) * inv_sqrt_bc2 + a.eps;
  p -= step_size * (m / denom);
  return clipped;
}

__global__ void __launch_bounds__(kOptThreads)
adam_update_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, size_t n,
                   AdamArgs a, const float* __restrict__ norm, const long long* __restrict__ step) {
  const float total = *norm * a.grad_scale;                                  // norm of the scaled gradient
  float coef = a.grad_scale;
  if (a.max_norm > 0.f) coef *= fminf(a.max_norm / (total + 1e-6f), 1.f);    // clip_grad_norm_
  const double t = (double)*step;
  const float bc1 = (float)(1.0 - pow((double)a.beta1, t)), bc2 = (float)(1.0 - pow((double)a.beta2, t));
  const float step_size = a.lr / bc1, inv_sqrt_bc2 = 1.f / sqrtf(bc2);
  const size_t n4 = n >> 2, stride = (size_t)gridDim.x * blockDim.x;
  float4 *p4 = reinterpret_cast<float4*>(p), *g4 = reinterpret_cast<float4*>(g), *m4 = reinterpret_cast<float4*>(m),
         *v4 = reinterpret_cast<float4*>(v);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 pp = p4[i], gg = g4[i], mm = m4[i], vv = v4[i];
    gg.x = adam_one(pp.x, gg.x, mm.x, vv.x, a, coef, step_size, inv_sqrt_bc2);
    gg.y = adam_one(pp.y, gg.y, mm.y, vv.y, a, coef, step_size, inv_sqrt_bc2);
    gg.z = adam_one(pp.z, gg.z, mm.z, vv.z, a, coef, step_size, inv_sqrt_bc2);
    gg.w = adam_one(pp.w, gg.w, mm.w, vv.w, a, coef, step_size, inv_sqrt_bc2);
    p4[i] = pp; m4[i] = mm; v4[i] = vv;
    if (a.zero_grad) g4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    else if (a.max_norm > 0.f || a.grad_scale != 1.f) g4[i] = gg;     // the scaled / clipped gradient, as clip_grad_norm_ leaves it
  }
  if (blockIdx.x == 0)
    for (size_t i = (n4 << 2) + threadIdx.x; i < n; i += blockDim.x) {
      const float gi = adam_one(p[i], g[i], m[i], v[i], a, coef, step_size, inv_sqrt_bc2);
      g[i] = a.zero_grad ? 0.f : gi;
    }
}

}  // namespace pcoe

using namespace pcoe;

extern "C" size_t pcoe_adam_workspace_bytes(void) { return align_up(sizeof(AdamWs), 256); }

extern "C" int pcoe_adam_step(float* param, float* grad, float* exp_avg, float* exp_avg_sq, size_t n, float lr,
                              float beta1, float beta2, float eps, float weight_decay, float max_grad_norm,
                              float grad_scale, int zero_grad, int64_t* step_dev, float* grad_norm_dev, void* workspace, void* stream) {
  if (n == 0) return PCOE_OK;
  if (!param || !grad || !exp_avg || !exp_avg_sq || !step_dev || !grad_norm_dev || !workspace)
    return fail(PCOE_ERR_NULL, "adam_step: NULL pointer");
  if (((uintptr_t)param | (uintptr_t)grad | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq) & 15)
    return fail(PCOE_ERR_BAD_SHAPE, "adam_step: the flat buffers must be 16-byte aligned");
  if (!(lr >= 0.f) || !(beta1 >= 0.f && beta1 < 1.f) || !(beta2 >= 0.f && beta2 < 1.f) || !(eps >= 0.f))
    return fail(PCOE_ERR_BAD_SHAPE, "adam_step: lr=%g betas=(%g,%g) eps=%g", lr, beta1, beta2, eps);
  cudaStream_t st = (cudaStream_t)stream;
  {
    LaunchScope ls("grad_sumsq_kernel", st);
    grad_sumsq_kernel<<<kOptBlocks, kOptThreads, 0, st>>>(grad, n, (AdamWs*)workspace, grad_norm_dev,
                                                         (long long*)step_dev);
    PCOE_TRY(ls.done());
  }
  AdamArgs a{lr, beta1, beta2, eps, weight_decay, max_grad_norm, grad_scale, zero_grad};
  LaunchScope ls("adam_update_kernel", st);
  const size_t n4 = n >> 2;
  int blocks = (int)((n4 + kOptThreads - 1) / kOptThreads);
  blocks = blocks < 1 ? 1 : (blocks > kNumSMs * 8 ? kNumSMs * 8 : blocks);
  adam_update_kernel<<<blocks, kOptThreads, 0, st>>>(param, grad, exp_avg, exp_avg_sq, n, a, grad_norm_dev,
                                                    (const long long*)step_dev);
  return ls.done();
}
