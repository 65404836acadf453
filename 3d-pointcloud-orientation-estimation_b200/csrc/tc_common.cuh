// tc_common.cuh — sm_100a tensor-core primitives used by the bf16 set-abstraction kernels:
// tcgen05.mma (kind::f16, bf16 operands from shared memory, fp32 accumulators in TMEM), TMEM
// allocation / loads, mbarrier completion, and the SWIZZLE_128B shared-memory operand layout.
//
// Operand tiles.  A tile is R rows x 64 bf16 (128 bytes per row), stored as 8-row / 1024-byte
// swizzle atoms: byte(row, col) = (row/8)*1024 + (row%8)*128 + (((col/8) ^ (row%8)) * 16) + (col%8)*2.
// The same physical tile serves two roles:
//   * K-major operand (contraction over the 64 columns): descriptor SBO = 1024 (next 8 rows),
//     K advance of 16 elements = +32 bytes on the start address;
//   * MN-major operand (contraction over the rows, MN = the 64 columns): SBO = 1024 (next 8
//     contraction rows), LBO = byte distance to the next 64-column tile, K advance of 16 rows = +2048 B.
// Tiles must be 1024-byte aligned (the swizzle XOR acts on absolute shared-memory address bits).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace pcoe {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// 16-byte store to a 32-bit shared-memory address (explicit st.shared: pointers rounded up to the
// 1024-byte tile alignment lose their address space and would otherwise compile to generic ST)
__device__ __forceinline__ void sts128(uint32_t saddr, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// byte offset of element (row, col) inside an [R x 64] bf16 SWIZZLE_128B tile
__device__ __host__ __forceinline__ uint32_t sw128_off(int row, int col) {
  return (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((((col >> 3) ^ (row & 7)) & 7) << 4) + (col & 7) * 2);
}

// ---- descriptors ---------------------------------------------------------------------------
// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start>>4 [0,14), LBO>>4 [16,30),
// SBO>>4 [32,46), version=1 [46,48), layout_type [61,64) (2 = SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// instruction descriptor (cute::UMMA::InstrDescriptor) for kind::f16, bf16 x bf16 -> fp32
__device__ __host__ __forceinline__ uint32_t make_idesc_bf16(int M, int N, bool a_mn_major, bool b_mn_major) {
  uint32_t d = 0;
  d |= 1u << 4;                         // c_format = F32
  d |= 1u << 7;                         // a_format = BF16
  d |= 1u << 10;                        // b_format = BF16
  d |= (a_mn_major ? 1u : 0u) << 15;
  d |= (b_mn_major ? 1u : 0u) << 16;
  d |= (uint32_t)(N >> 3) << 17;
  d |= (uint32_t)(M >> 4) << 24;
  return d;
}

// ---- TMEM ------------------------------------------------------------------------------------
template <int COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst) {  // one full warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "n"(COLS));
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {  // same warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS));
}

__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// make generic-proxy shared-memory writes visible to the async proxy (tensor core reads)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem]; issued by ONE thread
__device__ __forceinline__ void mma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, bool accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate)
      : "memory");
}
// Warp-wide variants: the WHOLE warp executes the call (convergent, warp-uniform operands) and one
// elected lane issues the instruction.  Keeping the issue loop warp-uniform lets the compiler hold
// descriptors, TMEM addresses and loop state in uniform registers; a loop run by a single thread of a
// divergent branch instead pays a vector->uniform register broadcast (R2UR + ELECT loop, ~100+ cycles)
// per tcgen05.mma.  elect.sync picks the same lane for the same member mask, so the commit below
// tracks the MMAs issued here.
__device__ __forceinline__ void mma_bf16_warp(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, bool accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\telect.sync _|q, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate)
      : "memory");
}
__device__ __forceinline__ void mma_commit_warp(uint64_t* mbar) {
  asm volatile(
      "{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}"
      ::"r"(smem_u32(mbar)) : "memory");
}
// broadcast of a value every lane already holds: tells the compiler it is warp-uniform
__device__ __forceinline__ uint32_t uniform_u32(uint32_t v) { return __shfl_sync(0xFFFFFFFFu, v, 0); }

// arrive on an mbarrier when all previously issued MMAs of this thread have completed
__device__ __forceinline__ void mma_commit(uint64_t* mbar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(mbar)) : "memory");
}

// 32 lanes x 32 consecutive fp32 columns: thread i of the warp gets TMEM lane (lane_base + i)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// 32 lanes x 16 consecutive fp32 columns
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// ---- mbarrier --------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* mbar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(mbar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* mbar, uint32_t parity) {
#if defined(PCOE_EXP_WAIT_HINT)
  // experiment: explicit suspend-time hint (ns) - the waiting warp leaves the issue slots to the working warps
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}"
      ::"r"(smem_u32(mbar)), "r"(parity), "r"((uint32_t)PCOE_EXP_WAIT_HINT)
      : "memory");
#elif defined(PCOE_EXP_WAIT_SLEEP)
  // experiment: back off between polls
  uint32_t done = 0;
  while (true) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(smem_u32(mbar)), "r"(parity) : "memory");
    if (done) break;
    __nanosleep(PCOE_EXP_WAIT_SLEEP);
  }
#else
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}"
      ::"r"(smem_u32(mbar)), "r"(parity)
      : "memory");
#endif
}

// pack 8 floats to 8 bf16 (round to nearest even) as one 16-byte value
__device__ __forceinline__ uint4 pack8_bf16(const float (&v)[8]) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
  __nv_bfloat162 c = __floats2bfloat162_rn(v[4], v[5]), d = __floats2bfloat162_rn(v[6], v[7]);
  uint4 r;
  r.x = *reinterpret_cast<uint32_t*>(&a); r.y = *reinterpret_cast<uint32_t*>(&b);
  r.z = *reinterpret_cast<uint32_t*>(&c); r.w = *reinterpret_cast<uint32_t*>(&d);
  return r;
}

// packed bf16x2 arithmetic (one instruction per two elements; every result is rounded to bf16 once)
__device__ __forceinline__ uint32_t bf2_bcast(float x) {   // (bf16(x), bf16(x))
  uint32_t d;
  asm("cvt.rn.bf16x2.f32 %0, %1, %1;" : "=r"(d) : "f"(x));
  return d;
}
__device__ __forceinline__ uint32_t bf2_fma(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t d;
  asm("fma.rn.bf16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}
__device__ __forceinline__ uint32_t bf2_fma_relu(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t d;
  asm("fma.rn.relu.bf16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}

// pack 8 floats to 8 bf16 with ReLU folded into the conversion (cvt.rn.relu: max(x, 0) then round;
// identical to rounding relu(x) because rounding preserves sign and zero)
__device__ __forceinline__ uint32_t pack2_bf16_relu(float lo, float hi) {
  uint32_t d;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}
__device__ __forceinline__ uint4 pack8_bf16_relu(const float (&v)[8]) {
  uint4 r;
  r.x = pack2_bf16_relu(v[0], v[1]); r.y = pack2_bf16_relu(v[2], v[3]);
  r.z = pack2_bf16_relu(v[4], v[5]); r.w = pack2_bf16_relu(v[6], v[7]);
  return r;
}

}  // namespace tc
}  // namespace pcoe
