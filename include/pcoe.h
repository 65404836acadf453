/*
 * pcoe.h — C ABI of libpcoe.so: the B200 (sm_100a) PointNet++ set-abstraction hot path of
 * 0xPabloxx/3d-pointcloud-orientation-estimation.
 *
 * The reference is pure Python/PyTorch and has no FFI; each entry point below names the reference
 * function (file:line under the reference root) whose arithmetic it replaces.  A maintainer binds
 * these with ctypes.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the parameter name ends in `_host`;
 *   - tensors are dense, row-major, with the shapes given in the comments;
 *   - indices are int32 on this side (the Python layer widens to int64 where the reference exposes them);
 *   - `stream` is a cudaStream_t passed as void*; every call only enqueues work on it (no device
 *     synchronisation, no allocation, CUDA-graph capturable).  pcoe_sa_backward may run independent kernels of
 *     one call on a library-owned auxiliary stream (one per device, created on first use): it is forked from and
 *     joined back into `stream` inside the call, under a host lock, so callers see plain stream semantics;
 *   - every call returns 0 on success or a negative pcoe_status; the message of the last failure
 *     on the calling thread is available from pcoe_last_error();
 *   - the library never keeps a caller pointer after the call returns.
 */
#ifndef PCOE_H_
#define PCOE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PCOE_VERSION 100 /* 0.1.0 */

typedef enum pcoe_status {
  PCOE_OK = 0,
  PCOE_ERR_BAD_SHAPE = -1,     /* non-positive / inconsistent dimensions                       */
  PCOE_ERR_UNSUPPORTED = -2,   /* a size or mode this build has no kernel for                  */
  PCOE_ERR_NULL = -3,          /* a required pointer is NULL                                   */
  PCOE_ERR_CUDA = -4,          /* a CUDA runtime call or launch failed                         */
  PCOE_ERR_WORKSPACE = -5      /* workspace / saved buffer smaller than *_bytes() reports       */
} pcoe_status;

int pcoe_version(void);
/* Thread-local, NUL-terminated description of the most recent failing call on this thread. */
const char* pcoe_last_error(void);
/* Number of kernels this library has enqueued since load (all threads); used by bench.py for
 * its `gpu_launches` claim. */
uint64_t pcoe_launch_count(void);

/* Per-kernel timing for bench.py's roofline line: when enabled every kernel this library enqueues
 * is bracketed by CUDA events on its stream.  pcoe_profile_report synchronises on those events and
 * writes "name,launches,total_ms\n" lines (NUL-terminated, truncated to cap) into the HOST buffer
 * `buf_host`, then clears the records.  Not for use during CUDA-graph capture. */
int pcoe_profile_enable(int on);
int pcoe_profile_report(char* buf_host, size_t cap);

/* --------------------------------------------------------------------------------------------
 * Sampling
 * ------------------------------------------------------------------------------------------ */

/* Farthest-point sampling.  Replaces farthest_point_sample(), PointNet++Demo.py:8-29.
 *   xyz       [B,N,3] f32
 *   start_idx [B] i32 first centroid of every cloud (the reference draws it with torch.randint,
 *             :20); NULL = index 0 for every cloud
 *   out_idx   [B,S] i32   indices in visiting order (out_idx[b,0] == start_idx[b])
 *   out_xyz   [B,S,3] f32 the sampled coordinates, or NULL
 * Bit-exact contract: distances are ((dx*dx)+(dy*dy))+(dz*dz) with individually rounded fp32
 * operations, running minimum starts at 1e10, arg-max ties resolve to the lowest index. */
int pcoe_fps_f32(const float* xyz, int B, int N, int S, const int32_t* start_idx,
                 int32_t* out_idx, float* out_xyz, void* stream);

/* Gather of sampled centroids.  Replaces index_points(xyz, fps_idx) with a (B,S) index,
 * models/base.py:4-14 (call site models/pointnet_pp_8dir.py:29).
 *   src [B,N,C] f32, idx [B,S] i32 -> out [B,S,C] f32.  Fails (BAD_SHAPE) only on shapes;
 *   an out-of-range index is clamped into [0,N) (the reference would raise IndexError). */
int pcoe_gather_points_f32(const float* src, int B, int N, int C, const int32_t* idx, int S,
                           float* out, void* stream);

/* HOST function (no GPU work, all pointers are host pointers): exact replay of the reference's subset draw
 *     torch.stack([torch.randperm(N)[:S] for _ in range(B)])          (models/pointnet_pp_8dir.py:28)
 * on torch's CPU generator.  `rng_state` = the bytes of torch.get_rng_state() (mt19937 state, legacy layout), advanced
 * in place exactly as B calls of torch.randperm(N) advance it - hand it back with torch.set_rng_state.  out_idx [B,S]
 * i32 on the host (typically pinned, uploaded as a static input of the captured training step).  Bit-identical to
 * torch (tests/test_host_rng_cpu.py), ~10x faster than the Python-level loop. */
int pcoe_host_randperm_subsets(uint8_t* rng_state, size_t state_bytes, int B, int N, int S, int32_t* out_idx);

/* Uniform random subset without replacement drawn on the device (partial Fisher-Yates, Philox-
 * style counter RNG keyed by (seed, call offset, cloud)).  Same distribution as the reference's
 * torch.randperm(N)[:S] (models/pointnet_pp_8dir.py:28; the on-device variant is
 * models/pointnet_pp_Fwd.py:44-47) but a different random stream: use the host-replayed
 * permutation (Python layer, sampler="randperm_host") when index parity with a seeded reference
 * run is required.   out_idx [B,S] i32.  `offset_dev` (device pointer, may be NULL) is added to
 * `offset` on the device, so that a step captured in a CUDA graph draws fresh subsets on every
 * replay (the caller increments the counter with an ordinary captured kernel). */
int pcoe_random_subset(int B, int N, int S, uint64_t seed, uint64_t offset,
                       const uint64_t* offset_dev, int32_t* out_idx, void* stream);
/* The same draw, and the selected points gathered in the same launch: out_xyz [B,S,3] = xyz[b, out_idx[b,s]]
 * (replaces the index_points call that follows the draw, models/pointnet_pp_8dir.py:28-29). */
int pcoe_random_subset_xyz(int B, int N, int S, uint64_t seed, uint64_t offset,
                           const uint64_t* offset_dev, int32_t* out_idx, const float* xyz,
                           float* out_xyz, void* stream);

/* --------------------------------------------------------------------------------------------
 * Grouping
 * ------------------------------------------------------------------------------------------ */

/* Pairwise squared distances.  Replaces square_distance(src, dst), models/base.py:20-27:
 *   out[b,i,j] = -2 <src[b,i], dst[b,j]> + |src[b,i]|^2 + |dst[b,j]|^2   (the reference's expanded form, fp32; like the
 *   reference's it can be slightly negative for coincident points).  src [B,N,C], dst [B,M,C] -> out [B,N,M] f32.
 * The hot path never materialises this matrix (pcoe_knn_f32 fuses it with the selection); this entry point exists for
 * callers of the helper itself. */
int pcoe_square_distance_f32(const float* src, const float* dst, int B, int N, int M, int C, float* out,
                             void* stream);

/* k nearest neighbours of every centroid.  Replaces query_ball_point(new_xyz, xyz, nsample) =
 * square_distance + topk(largest=False, sorted=False), models/base.py:20-35.
 *   xyz [B,N,3] f32, new_xyz [B,S,3] f32 -> out_idx [B,S,K] i32.
 * The reference's order inside a row is unspecified (sorted=False); this kernel writes ascending
 * point index; among points at exactly the K-th distance the lowest indices are kept.  Distances
 * are the direct sum of squared differences in fp32.  Requires 1 <= K <= min(N, 128). */
int pcoe_knn_f32(const float* xyz, const float* new_xyz, int B, int N, int S, int K,
                 int32_t* out_idx, void* stream);

/* Radius ball query.  Replaces query_ball_point(radius, nsample, xyz, new_xyz),
 * PointNet++Demo.py:49-70: the first `nsample` point indices (ascending index) whose squared
 * distance is NOT greater than (float)(radius*radius); short rows are padded with the row's
 * first hit; a row with no hit is filled with N (as the reference does).
 * Distances as in pcoe_fps_f32 (bit-exact contract).  out_idx [B,S,nsample] i32. */
int pcoe_ball_query_f32(const float* xyz, const float* new_xyz, int B, int N, int S, int nsample,
                        double radius, int32_t* out_idx, void* stream);

/* Multi-scale (MSG-style) radius grouping: pcoe_ball_query_f32 for `nscales` (1..4) radii in one pass over the cloud
 * (the cloud is staged in shared memory once, every squared distance is computed once and compared with every radius).
 * out_idx_host[r] -> [B,S,nsample_host[r]] i32 device buffers; row r is bit-identical to
 * pcoe_ball_query_f32(radius_host[r], nsample_host[r]).  Built from query_ball_point, PointNet++Demo.py:49-70, applied
 * per scale.  The three `_host` arrays are HOST arrays of length nscales. */
int pcoe_ball_query_multi_f32(const float* xyz, const float* new_xyz, int B, int N, int S, int nscales,
                              const double* radius_host, const int* nsample_host, int32_t* const* out_idx_host,
                              void* stream);

/* --------------------------------------------------------------------------------------------
 * Set abstraction: gather + centre + 3 x (1x1 conv -> BatchNorm -> ReLU) + max over neighbours
 * Replaces PointNetSetAbstraction.forward, models/pointnet_pp_8dir.py:21-43 (identical copies in
 * models/Pointnet_pp_xyz.py:22-44, pointnet_pp.py, Pointnet_pp_xyz_Schedmit.py, pointnet_pp_Fwd.py)
 * and its autograd.
 * ------------------------------------------------------------------------------------------ */

/* FP32:   CUDA-core fp32 GEMMs (exact-arithmetic parity mode).
 * BF16:   tcgen05, bf16 operands / bf16 stored activations, fp32 accumulate (throughput mode, stated tolerance).
 * BF16X3: tcgen05, every operand split into two bf16 planes (hi + lo), three MMAs per step, fp32 stored
 *         activations and fp32 transforms: fp32-class accuracy (2^-16 relative per product) on the tensor pipe. */
enum { PCOE_PRECISION_FP32 = 0, PCOE_PRECISION_BF16 = 1, PCOE_PRECISION_BF16X3 = 2 };

typedef struct pcoe_sa_desc {
  int32_t B;          /* clouds                                                               */
  int32_t N;          /* points per cloud entering the layer                                  */
  int32_t S;          /* centroids per cloud (1 when group_all)                               */
  int32_t K;          /* neighbours per centroid (N when group_all); power of two, 1..128     */
  int32_t D;          /* feature channels of `feats` (0 = none); conv input width is 3 + D     */
  int32_t C1, C2, C3; /* output widths of the three 1x1 convolutions                          */
  int32_t group_all;  /* 1: one group of all N points, absolute xyz (pointnet_pp_8dir.py:23-26) */
  int32_t train;      /* 1: batch statistics + running-stat update; 0: running statistics     */
  int32_t precision;  /* PCOE_PRECISION_FP32 | PCOE_PRECISION_BF16 | PCOE_PRECISION_BF16X3 */
  float eps;          /* BatchNorm eps (1e-5 in the reference)                                */
  float momentum;     /* BatchNorm momentum (0.1 in the reference)                            */
} pcoe_sa_desc;

/* Parameters of the three conv+BN stages, PyTorch layouts: W[l] is Conv2d.weight (Cout,Cin,1,1)
 * = row-major [Cout][Cin]; running_* are updated in place when desc.train != 0. */
typedef struct pcoe_sa_params {
  const float* W[3];
  const float* bias[3];
  const float* gamma[3];
  const float* beta[3];
  float* running_mean[3];
  float* running_var[3];
  long long* num_batches_tracked[3]; /* BatchNorm2d.num_batches_tracked (int64 scalars), incremented by one per
                                        train-mode forward; entries may be NULL */
} pcoe_sa_params;

/* Gradients written by pcoe_sa_backward, same layouts.  accumulate == 0: every array is
 * overwritten.  accumulate != 0: the gradient is ADDED to what the arrays hold (autograd's
 * `p.grad += g`): the Python layer points these at slices of the flat gradient buffer that the
 * data-parallel all-reduce and the fused optimizer work on, so no per-parameter add kernel runs. */
typedef struct pcoe_sa_grads {
  float* dW[3];
  float* dbias[3];
  float* dgamma[3];
  float* dbeta[3];
  int32_t accumulate;
} pcoe_sa_grads;

/* Bytes of the `saved` buffer (forward -> backward state: pre-BN activations, batch statistics,
 * arg-max slots) and of the transient `workspace`.  Both must be 256-byte aligned. */
size_t pcoe_sa_saved_bytes(const pcoe_sa_desc* desc);
size_t pcoe_sa_workspace_bytes(const pcoe_sa_desc* desc);

/*   xyz     [B,N,3] f32     coordinates entering the layer
 *   new_xyz [B,S,3] f32     centroids (ignored when group_all)
 *   nbr     [B,S,K] i32     neighbour indices into N (ignored when group_all)
 *   feats   [B,N,D] f32     or NULL when D == 0
 *   out     [B,S,C3] f32    max-pooled features = the reference's second return value
 *   saved / workspace       see above; `saved` may be NULL when desc.train == 0 */
int pcoe_sa_forward(const pcoe_sa_desc* desc, const float* xyz, const float* new_xyz,
                    const int32_t* nbr, const float* feats, const pcoe_sa_params* params,
                    float* out, void* saved, size_t saved_bytes, void* workspace,
                    size_t workspace_bytes, void* stream);

/*   grad_out   [B,S,C3] f32   gradient w.r.t. `out`
 *   out        [B,S,C3] f32   the forward result (ReLU mask)
 *   grad_feats [B,N,D] f32    overwritten; NULL when D == 0
 * Only defined for a forward that ran with desc.train != 0 (the reference never back-propagates
 * through eval-mode BatchNorm). */
int pcoe_sa_backward(const pcoe_sa_desc* desc, const float* xyz, const float* new_xyz,
                     const int32_t* nbr, const float* feats, const pcoe_sa_params* params,
                     const float* out, const float* grad_out, const void* saved,
                     size_t saved_bytes, float* grad_feats, const pcoe_sa_grads* grads,
                     void* workspace, size_t workspace_bytes, void* stream);

/* --------------------------------------------------------------------------------------------
 * Vanilla PointNet inference path (models/pointnet.py:6-129, eval mode)
 * ------------------------------------------------------------------------------------------ */

/* A pointwise MLP stack followed by the max over each cloud's points:
 *     h = x;  for l < nlayers: h = BN_eval_l(W_l h + b_l), ReLU after every layer except - when relu_last == 0 - the last;
 *     out[b, :] = max over the rows of cloud b of h
 * i.e. STN3d / STNkd's conv1..3 + bn1..3 + torch.max (models/pointnet.py:24-27 / :55-58) and PointNetEncoder's
 * conv1..3 + bn1..3 + torch.max (:93,102-104; bn3 without ReLU).  Same split-operand tcgen05 kernels as
 * PCOE_PRECISION_BF16X3 set abstraction with two bf16 planes per operand (16 significant bits, 3 MMAs per product, fp32
 * accumulate and fp32 activations: ~1e-5 relative on the outputs).
 *   M              rows = clouds * rows_per_cloud;  rows_per_cloud must be a multiple of 32 (pad a cloud by repeating
 *                  one of its points: the max does not change)
 *   use_xyz, D     first-layer input = [xyz (3, when use_xyz) | feats (D)], D a multiple of 64 (0 allowed with use_xyz);
 *                  W[0] is [C[0]][3*use_xyz + D] row-major in that column order
 *   nlayers        2 or 3;  C[l] multiples of 64
 *   params         W / bias / gamma / beta / running_mean / running_var of layers 0..nlayers-1 (bias entries may be NULL)
 *   xyz [M,3], feats [M,D] f32 point-major;  out [M / rows_per_cloud, C[nlayers-1]] f32 */
typedef struct pcoe_pointmlp_desc {
  int32_t M, rows_per_cloud, D, use_xyz, nlayers;
  int32_t C[3];
  int32_t relu_last;
  float eps;
} pcoe_pointmlp_desc;
size_t pcoe_pointmlp_workspace_bytes(const pcoe_pointmlp_desc* desc);
int pcoe_pointmlp_forward(const pcoe_pointmlp_desc* desc, const float* xyz, const float* feats,
                          const pcoe_sa_params* params, float* out, void* workspace, size_t workspace_bytes,
                          void* stream);

/* y[m,:] = act(scale * (W x[m,:]) + shift) for a narrow input: x [M,Cin] (Cin <= 8), W [Cout,Cin], per-channel
 * scale / shift (BatchNorm eval folded with the conv bias), Cout a multiple of 4, y [M,Cout] point-major.  The 3 -> 64
 * first layer of PointNetEncoder (models/pointnet.py:93) when its output feeds the feature transform. */
int pcoe_pointwise_linear_f32(const float* x, int M, int Cin, const float* W, const float* scale, const float* shift,
                              int Cout, int relu, float* y, void* stream);

/* --------------------------------------------------------------------------------------------
 * Device-side input pipeline
 * ------------------------------------------------------------------------------------------ */

/* Builds a batch from a dataset resident in HBM: cloud `cloud_ids[b]` (rows offsets[c] .. offsets[c+1] of `points`) is
 * resampled to `num` points.  Replaces sample_pts(arr, num) = arr[np.random.choice(len(arr), num, replace=len(arr) < num)]
 * (dataloader_multi_peak_vonMises.py:21-26, dataloader_8dir_sampled.py:13-15, dataloader_single_peak_vonMises.py:12-14):
 * n >= num -> uniform subset without replacement (ascending source order), n < num -> uniform draws with replacement;
 * same distribution as numpy's, another random stream (keyed by seed and base_draw + *counter_dev * B + b, so a CUDA-graph
 * replay that increments the counter draws fresh subsets).  Deterministic; oracle/data.py reproduces the indices.
 *   points [total,3] f32, offsets [nclouds+1] i64, cloud_ids [B] i32 (values in [0,nclouds)),
 *   out_xyz [B,num,3] f32, out_idx [B,num] i32 source row inside the cloud (may be NULL), counter_dev may be NULL. */
int pcoe_resample_clouds_f32(const float* points, const int64_t* offsets, int nclouds, const int32_t* cloud_ids,
                             int B, int num, uint64_t seed, uint64_t base_draw, const uint64_t* counter_dev,
                             float* out_xyz, int32_t* out_idx, void* stream);

/* --------------------------------------------------------------------------------------------
 * Losses (value and gradient in one launch; gradients are d loss[b] / d input[b,...])
 * ------------------------------------------------------------------------------------------ */

enum { PCOE_VM_SINGLE = 0, PCOE_VM_MULTI = 1 };

/* Elementwise KL(vM(mu_p,kappa_p) || vM(mu_q,kappa_q)).
 *   variant PCOE_VM_SINGLE: kl_von_mises, train_single_peak_vonMises_KL.py:23-28
 *           (no clamp, no wrap, A:=0 for kappa_p<=1e-6, fp32 overflow of I0 -> NaN/inf as there)
 *   variant PCOE_VM_MULTI : kl_von_mises, train_multi_peaks_vonMises_KL.py:38-52
 *           (kappa clamped to [1e-6,500], delta wrapped to [-pi,pi))
 *   all arrays [n] f32; dmu / dkappa may be NULL. */
int pcoe_vm_kl_fwd_bwd(const float* mu_p, const float* kappa_p, const float* mu_q,
                       const float* kappa_q, int n, int variant, float* loss, float* dmu,
                       float* dkappa, void* stream);

/* Matched mixture loss.  Replaces match_loss, train_multi_peaks_vonMises_KL.py:54-81 including
 * the host scipy.optimize.linear_sum_assignment (minimum over the <= 24 permutations).
 *   mu,kappa,w [B,Kmax] f32; gt [B,Kmax,gt_stride] f32 with (mu,kappa) in columns 0,1;
 *   K_gt [B] i32 (<=0 -> loss 0, no gradient); Kmax <= 4.
 *   loss [B]; dmu,dkappa,dw [B,Kmax] (may be NULL); perm [B,Kmax] i32 matched gt column per
 *   predicted component (-1 beyond K), may be NULL. */
int pcoe_mvm_match_fwd_bwd(const float* mu, const float* kappa, const float* w, const float* gt,
                           int gt_stride, const int32_t* K_gt, int B, int Kmax, float* loss,
                           float* dmu, float* dkappa, float* dw, int32_t* perm, void* stream);

/* Mixture-of-von-Mises head transform of PointNetPPMvM.forward, models/pointnet_pp_mvM.py:91-125:
 *   weight = softmax(pi / temp);  mu = atan2 of the eps-normalised (1e-4) mu_raw pair with the (1,0)
 *   fallback for vectors shorter than 1e-3 after normalisation;  kappa = softplus(kappa_raw) + 1e-6,
 *   clamped to kappa_max when clamp_kappa != 0.   pi, kappa_raw, outputs [B,K]; mu_raw [B,K,2]; K <= 8.
 * The backward call recomputes the forward from the raw head outputs; g_* may be NULL (= zero). */
int pcoe_mvm_head_fwd(const float* pi, const float* mu_raw, const float* kappa_raw, int B, int K,
                      float temp, float kappa_max, int clamp_kappa, float* weight, float* mu,
                      float* kappa, void* stream);
int pcoe_mvm_head_bwd(const float* pi, const float* mu_raw, const float* kappa_raw, int B, int K,
                      float temp, float kappa_max, int clamp_kappa, const float* g_weight,
                      const float* g_mu, const float* g_kappa, float* d_pi, float* d_mu_raw,
                      float* d_kappa_raw, void* stream);

/* --------------------------------------------------------------------------------------------
 * Trunk building blocks (fp32): the 1024 -> 512 -> 256 -> heads MLP of PointNetPPMvM on B rows,
 * models/pointnet_pp_mvM.py:56-66 (fc1, ln1, fc2, ln2, dropout, head_pi / head_mu / head_kappa) and
 * :79-84,91-105 (their use), plus autograd.  W[i] is nn.Linear.weight [N_i, K] row-major.
 * A call with nseg > 1 applies several linears to the same input (the three heads); its output (and the
 * matching dy) is stored segment-major: [B x N_0][B x N_1]... so that every head's block is dense.
 * ------------------------------------------------------------------------------------------ */
/* y = x W^T + bias.  x [B,K], bias[i] may be NULL.  Deterministic.  max_parts <= 1 (or nparts NULL): y [B, sum N]
 * is the final result.  max_parts > 1: the contraction may be split over *nparts <= max_parts slices for
 * parallelism; y then holds *nparts partial results [part][B, sum N] WITHOUT the bias, to be summed (in order)
 * by the consumer - pcoe_ln_relu_dropout_fwd does. */
int pcoe_linear_fwd(const float* x, int B, int K, int nseg, const float* const* W,
                    const float* const* bias, const int* N, float* y, int max_parts, int* nparts,
                    void* stream);
/* dx [B,K] = dy W (overwritten). */
int pcoe_linear_bwd_dx(const float* dy, int B, int K, int nseg, const float* const* W, const int* N,
                       float* dx, void* stream);
/* dW[i] [N_i,K] = dy_i^T x, dbias[i] [N_i] = column sums of dy_i (dbias[i] may be NULL);
 * accumulate != 0 adds to the arrays instead of overwriting them. */
int pcoe_linear_bwd_dw(const float* dy, const float* x, int B, int K, int nseg, const int* N,
                       float* const* dW, float* const* dbias, int accumulate, void* stream);
/* out = dropout_p(relu(LayerNorm_eps(x) * gamma + beta)), one row per cloud.  train == 0: no dropout.
 * The keep mask is drawn from Philox4x32-10 keyed by (seed, *counter_dev, row, column) - a different stream
 * than torch's dropout, same distribution; mean / rstd [B] and mask [B,N] u8 are saved for backward
 * (each may be NULL in inference).  x may be nparts partial sums [part][B,N] (+ xbias [N]): their ordered
 * sum is written to h [B,N] and used as the input (h may be NULL when nparts <= 1 and xbias is NULL). */
int pcoe_ln_relu_dropout_fwd(const float* x, int nparts, const float* xbias, float* h,
                             const float* gamma, const float* beta, int B, int N,
                             float eps, float p, int train, uint64_t seed,
                             const uint64_t* counter_dev, float* out, float* mean, float* rstd,
                             uint8_t* mask, void* stream);
/* dx overwritten; dgamma / dbeta are ADDED to (zero them first for a plain gradient); dxbias [N], optional, also
 * ADDED to: the column sums of dx = the bias gradient of the linear layer that produced x. */
int pcoe_ln_relu_dropout_bwd(const float* dout, const float* x, const float* out, const float* gamma,
                             const float* mean, const float* rstd, const uint8_t* mask, int B, int N,
                             float p, int train, float* dx, float* dgamma, float* dbeta, float* dxbias,
                             void* stream);

/* The last stage of the mixture model's trunk as ONE launch per direction (models/pointnet_pp_mvM.py:84,91-125):
 * pcoe_ln_relu_dropout_fwd followed, per row, by up to three small linear heads that share the row (nseg, W, bias,
 * Nout as in pcoe_linear_fwd; sum(Nout) <= 64; `raw` = their outputs, segment-major) and, when K > 0, by the head
 * transform of pcoe_mvm_head_fwd (segments must be pi (K) | mu_raw (2K) | kappa_raw (K)). */
int pcoe_ln_relu_dropout_heads_fwd(const float* x, int nparts, const float* xbias, float* h, const float* gamma,
                                   const float* beta, int B, int N, float eps, float p, int train, uint64_t seed,
                                   const uint64_t* counter_dev, float* out, float* mean, float* rstd,
                                   uint8_t* mask, int nseg, const float* const* W, const float* const* bias,
                                   const int* Nout, float* raw, int K, float temp, float kappa_max,
                                   int clamp_kappa, float* weight, float* mu, float* kappa, void* stream);
/* Its backward: d_raw = pcoe_mvm_head_bwd(raw; g_w, g_mu, g_k) (written, for the heads' pcoe_linear_bwd_dw),
 * the heads' data gradient and pcoe_ln_relu_dropout_bwd on it, per row. */
int pcoe_heads_ln_relu_dropout_bwd(const float* raw, int K, float temp, float kappa_max, int clamp_kappa,
                                   const float* g_w, const float* g_mu, const float* g_k, float* d_raw, int nseg,
                                   const float* const* W, const int* Nout, const float* x, const float* out,
                                   const float* gamma, const float* mean, const float* rstd, const uint8_t* mask,
                                   int B, int N, float p, int train, float* dx, float* dgamma, float* dbeta,
                                   float* dxbias, void* stream);

/* Soft-label cross entropy -(p * log_softmax(logits)).sum(1).  Replaces
 * kl_loss_per_sample_from_logits, train_8dir_KL.py:60-68.  logits,p [B,C] f32, C <= 64. */
int pcoe_soft_ce_fwd_bwd(const float* logits, const float* p, int B, int C, float* loss,
                         float* dlogits, void* stream);

/* --------------------------------------------------------------------------------------------
 * Optimizer step over flat buffers: gradient-norm clipping + Adam (+ zero_grad)
 * Replaces torch.nn.utils.clip_grad_norm_(params, max_norm) + torch.optim.Adam.step()
 * (+ optimizer.zero_grad()), train_multi_peaks_vonMises_KL.py:182,221,235-236,
 * train_single_peak_vonMises_KL.py:68,81,90, train_8dir_KL.py:72,93,97.
 *   param, grad, exp_avg, exp_avg_sq   [n] f32, 16-byte aligned flat buffers (all parameters
 *                                      concatenated; the drop-in modules' p.data / p.grad are views)
 *   grad_scale     the gradient is multiplied by this first (1/world_size after a SUM all-reduce; 1 otherwise)
 *   max_grad_norm  > 0: grad *= min(1, max_grad_norm / (||grad||_2 + 1e-6)) next; <= 0: no clipping
 *   zero_grad      != 0: grad is cleared after use; == 0: grad holds the clipped gradient on return
 *   step_dev       [1] i64 device step counter, incremented by this call (bias correction uses
 *                  the incremented value, as torch does); lives on the device so that a step captured
 *                  in a CUDA graph keeps counting on replay
 *   grad_norm_dev  [1] f32 out: ||grad||_2 of the UNSCALED buffer (times grad_scale = clip_grad_norm_'s return value)
 *   workspace      pcoe_adam_workspace_bytes() bytes, ZEROED ONCE by the caller before the first call
 *                  (the library leaves it zeroed)
 * Update rule (torch.optim.Adam, amsgrad=False, maximize=False; weight_decay is L2 as in Adam):
 *   m = m + (g-m)(1-b1); v = b2 v + (1-b2) g^2; p -= lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps) */
size_t pcoe_adam_workspace_bytes(void);
int pcoe_adam_step(float* param, float* grad, float* exp_avg, float* exp_avg_sq, size_t n, float lr,
                   float beta1, float beta2, float eps, float weight_decay, float max_grad_norm,
                   float grad_scale, int zero_grad, int64_t* step_dev, float* grad_norm_dev, void* workspace,
                   void* stream);

/* --------------------------------------------------------------------------------------------
 * Data-parallel gradient exchange over NVLink peer memory ("allreduce for the MLP gradients only")
 * ------------------------------------------------------------------------------------------ */

/* In-place sum-all-reduce of n floats at float offset `offset` of a SYMMETRIC buffer: bufs_host[p] is THIS process's
 * mapping of rank p's buffer, pads_host[p] of rank p's signal pad (>= 32 uint32, zeroed once before the first call;
 * torch.distributed._symmetric_memory provides both).  Two-shot with a pushed second shot: every rank announces its
 * gradients (flag A), rank r sums slice r over the ranks in rank order and stores the sum into every rank's buffer, then
 * the ranks exchange "pushes landed" flags (B); when the launch completes nobody touches this rank's buffer any more.
 * mc_buf: NVSwitch multicast mapping of the same buffer (multimem.ld_reduce / multimem.st do the sum and the broadcast
 * in the switch) or NULL for plain peer loads / stores.  Every rank of the node must make the same sequence of calls.
 * ctl_dev: 16 uint32 of this rank's device memory, zeroed once (the launch epoch lives there, so the call can be captured
 * in a CUDA graph and replayed).  offset, n multiples of 4; world <= 8.  The two `_host` arrays are HOST arrays of
 * `world` device pointers.  max_ctas <= 0: 64. */
int pcoe_peer_allreduce_f32(const void* const* bufs_host, const void* const* pads_host, const void* mc_buf, int rank,
                            int world, size_t offset, size_t n, uint32_t* ctl_dev, int max_ctas, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PCOE_H_ */
