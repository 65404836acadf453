// sa_tc.cuh — tcgen05 / TMEM versions of the two GEMM shapes of the set-abstraction MLP
// (precision = PCOE_PRECISION_BF16): bf16 operands built in shared memory by the same producers
// as the fp32 path (gather+centre, BN+ReLU on load, BatchNorm-backward combination), fp32
// accumulation in tensor memory, and the same epilogues (batch statistics, group max/min, ReLU mask,
// scatter-add) fed from the accumulator through a shared-memory staging tile.
//
//   tc_gemm_nt_kernel   C[128 rows x 128] = A'[rows x K] * W^T        forward layers and dgrad
//                       (A' and W K-major SWIZZLE_128B tiles, K in chunks of 64)
//   tc_gemm_tn_kernel   dW[128 x 128] += P[rows x 128]^T Q[rows x 128]  weight gradient
//                       (the SAME row-major tiles, consumed as MN-major operands: the contraction
//                        runs over the rows, 16 rows per tcgen05.mma, accumulating in TMEM over the
//                        whole row range of the CTA)
// One elected thread issues the MMAs; completion is tracked with tcgen05.commit -> mbarrier.
#pragma once
#include "sa_common.cuh"
#include "tc_common.cuh"

namespace pcoe {

constexpr int kTcTile = 128;                       // rows per tile = UMMA M
constexpr int kTcN = 128;                          // UMMA N
constexpr int kTcKC = 64;                          // K elements per operand tile (128 bytes)
constexpr int kTcTileBytes = kTcTile * 128;        // one [128 x 64] bf16 tile
constexpr int kTcStageLd = 65;
constexpr size_t kTcNtSmem = 1024 + 2 * kTcTileBytes + sizeof(float) * kTcTile * kTcStageLd;
constexpr size_t kTcTnSmem = 1024 + 4 * kTcTileBytes;

// bf16 weights for the tensor-core path: Wb[rows_pad][kpad] (zero padded), row-major
__global__ void convert_weights_kernel(const float* __restrict__ W, int Cout, int Cin,
                                       __nv_bfloat16* __restrict__ Wb, int rows_pad, int kpad,
                                       __nv_bfloat16* __restrict__ WbT, int rows_pad_t, int kpad_t,
                                       int perm_d /* >= 0: input channels reordered [feats(perm_d) | xyz(3)] */) {
  const int total = rows_pad * kpad, total_t = rows_pad_t * kpad_t;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total + total_t; e += gridDim.x * blockDim.x) {
    if (e < total) {
      const int r = e / kpad, c = e % kpad;
      int cs = c;   // source input channel
      if (perm_d >= 0) cs = c < perm_d ? c + 3 : c - perm_d;
      Wb[e] = __float2bfloat16_rn((r < Cout && c < Cin) ? W[(size_t)r * Cin + cs] : 0.f);
    } else {
      const int t = e - total, r = t / kpad_t, c = t % kpad_t;   // r over Cin, c over Cout
      int rs = r;
      if (perm_d >= 0) rs = r < perm_d ? r + 3 : r - perm_d;
      WbT[t] = __float2bfloat16_rn((r < Cin && c < Cout) ? W[(size_t)c * Cin + rs] : 0.f);
    }
  }
}

template <class AProd, class Epi>
__global__ void __launch_bounds__(256)
tc_gemm_nt_kernel(const AProd ap, const __nv_bfloat16* __restrict__ Bw, int ldb, int brows, const Epi epi,
                  int M, int Ncols, int Kdim) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);  // keeps the shared address space
  uint8_t* sA = smem;
  uint8_t* sB = smem + kTcTileBytes;
  float* Cs = reinterpret_cast<float*>(smem + 2 * kTcTileBytes);
  __shared__ uint64_t mbar;
  __shared__ uint32_t tmem_base;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.x * kTcTile, n0 = blockIdx.y * kTcN;
  const int rows = min(kTcTile, M - m0), cols = min(kTcN, Ncols - n0);

  if (warp == 0) tc::tmem_alloc<kTcN>(&tmem_base);
  if (tid == 0) tc::mbar_init(&mbar, 1);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = tmem_base;
  const uint32_t idesc = tc::make_idesc_bf16(kTcTile, kTcN, false, false);

  uint32_t phase = 0;
  const int nk = (Kdim + kTcKC - 1) / kTcKC;
  for (int kb = 0; kb < nk; ++kb) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {          // 128 rows x 8 sixteen-byte units per operand tile
      const int u = tid + i * 256, r = u & 127, j = u >> 7;
      float v[8];
      ap.load8(m0 + r, kb * kTcKC + j * 8, v);
      *reinterpret_cast<uint4*>(sA + tc::sw128_off(r, j * 8)) = tc::pack8_bf16(v);
      uint4 w = make_uint4(0u, 0u, 0u, 0u);
      if (n0 + r < brows) w = __ldg(reinterpret_cast<const uint4*>(Bw + (size_t)(n0 + r) * ldb + kb * kTcKC + j * 8));
      *reinterpret_cast<uint4*>(sB + tc::sw128_off(r, j * 8)) = w;
    }
    tc::fence_proxy_async();
    __syncthreads();
    if (tid == 0) {
      tc::fence_after_sync();
      const uint32_t a0 = tc::smem_u32(sA), b0 = tc::smem_u32(sB);
#pragma unroll
      for (int ks = 0; ks < kTcKC / 16; ++ks)
        tc::mma_bf16(tmem, tc::make_desc_sw128(a0 + ks * 32, 16, 1024), tc::make_desc_sw128(b0 + ks * 32, 16, 1024),
                     idesc, kb > 0 || ks > 0);
      tc::mma_commit(&mbar);
    }
    tc::mbar_wait(&mbar, phase);
    phase ^= 1;
  }
  tc::fence_after_sync();

  // epilogue: two halves of 64 accumulator columns through the fp32 staging tile
  const int q = warp & 3, cpart = warp >> 2;
  for (int half = 0; half < 2; ++half) {
    const int hc = min(64, cols - half * 64);
    if (hc <= 0) break;
    float v[32];
    tc::tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(half * 64 + cpart * 32), v);
#pragma unroll
    for (int i = 0; i < 32; ++i) Cs[(q * 32 + lane) * kTcStageLd + cpart * 32 + i] = v[i];
    __syncthreads();
    epi.template run<64>(Cs, kTcStageLd, m0, n0 + half * 64, rows, hc);
    __syncthreads();
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc<kTcN>(tmem);
}

template <class PProd, class QProd>
__global__ void __launch_bounds__(256)
tc_gemm_tn_kernel(const PProd pp, const QProd qp, float* __restrict__ out, int ldo, int M, int Ca, int Cb,
                  int rows_per_split) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);  // keeps the shared address space
  uint8_t* sP = smem;                        // 2 tiles [128 rows x 64 ch]
  uint8_t* sQ = smem + 2 * kTcTileBytes;     // 2 tiles
  __shared__ uint64_t mbar;
  __shared__ uint32_t tmem_base;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int ca0 = blockIdx.x * 128, cb0 = blockIdx.y * 128;
  const int r_begin = blockIdx.z * rows_per_split, r_end = min(M, r_begin + rows_per_split);
  if (r_begin >= r_end) return;              // uniform for the whole CTA

  if (warp == 0) tc::tmem_alloc<kTcN>(&tmem_base);
  if (tid == 0) tc::mbar_init(&mbar, 1);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = tmem_base;
  const uint32_t idesc = tc::make_idesc_bf16(128, 128, true, true);

  uint32_t phase = 0;
  bool acc = false;
  for (int r0 = r_begin; r0 < r_end; r0 += 128) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {            // 128 rows x 16 units (128 channels) per operand
      const int u = tid + i * 256, r = u & 127, j = u >> 7;
      const int row = r0 + r;
      float v[8];
      if (row < r_end) pp.load8(row, ca0 + j * 8, v);
      else {
#pragma unroll
        for (int t = 0; t < 8; ++t) v[t] = 0.f;
      }
      *reinterpret_cast<uint4*>(sP + (j >> 3) * kTcTileBytes + tc::sw128_off(r, (j & 7) * 8)) = tc::pack8_bf16(v);
      if (row < r_end) qp.load8(row, cb0 + j * 8, v);
      else {
#pragma unroll
        for (int t = 0; t < 8; ++t) v[t] = 0.f;
      }
      *reinterpret_cast<uint4*>(sQ + (j >> 3) * kTcTileBytes + tc::sw128_off(r, (j & 7) * 8)) = tc::pack8_bf16(v);
    }
    tc::fence_proxy_async();
    __syncthreads();
    if (tid == 0) {
      tc::fence_after_sync();
      const uint32_t p0 = tc::smem_u32(sP), q0 = tc::smem_u32(sQ);
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) {       // 16 contraction rows per MMA = two 1024-byte atoms
        tc::mma_bf16(tmem, tc::make_desc_sw128(p0 + ks * 2048, kTcTileBytes, 1024),
                     tc::make_desc_sw128(q0 + ks * 2048, kTcTileBytes, 1024), idesc, acc || ks > 0);
      }
      tc::mma_commit(&mbar);
    }
    acc = true;
    tc::mbar_wait(&mbar, phase);
    phase ^= 1;
  }
  tc::fence_after_sync();
  const int q = warp & 3, cpart = warp >> 2;
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    float v[32];
    tc::tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(cpart * 64 + c * 32), v);
    const int ca = ca0 + q * 32 + lane;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const int cb = cb0 + cpart * 64 + c * 32 + i;
      if (ca < Ca && cb < Cb) atomicAdd(out + (size_t)ca * ldo + cb, v[i]);
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc<kTcN>(tmem);
}

}  // namespace pcoe
