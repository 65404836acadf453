// host_rng.cu — HOST-side exact replay of the reference's sampling RNG (no device code).
//
// The reference draws its "FPS" subsets with `torch.stack([torch.randperm(N)[:npoint] for _ in range(B)])` on torch's
// CPU generator (models/pointnet_pp_8dir.py:28), SA1 first, then SA2.  torch.randperm on the CPU is a forward
// Fisher-Yates shuffle driven by one 32-bit mt19937 draw per position (n - 1 draws; ATen randperm_cpu:
// `z = generator->random() % (n - i); swap(r[i], r[z + i])`), and position i is final after step i - so the first S
// entries need S swaps, but the generator must still advance by all n - 1 draws to stay in step with the reference.
// 128 Python-level randperm calls cost 1.9 ms per 64-cloud step (2.8 x the whole GPU step); this replay consumes the
// generator state blob of torch.get_rng_state() (CPUGeneratorImplStateLegacy layout: seed u64, left i32, seeded i32,
// next u64, state[624] u64, ...), produces bit-identical subsets ~10 x faster and hands the advanced state back
// (torch.set_rng_state), so everything torch draws afterwards is unchanged too.  Checked against torch.randperm
// itself in tests/test_host_rng_cpu.py.
#include "common.cuh"
#include <cstring>
#include <vector>

namespace {

constexpr int kN = 624, kM = 397;
constexpr uint32_t kMatrixA = 0x9908b0dfu, kUMask = 0x80000000u, kLMask = 0x7fffffffu;

struct Mt {
  uint32_t state[kN];
  int left;
  uint32_t next;
  static uint32_t twist(uint32_t u, uint32_t v) { return (((u & kUMask) | (v & kLMask)) >> 1) ^ ((v & 1u) ? kMatrixA : 0u); }
  void next_state() {
    uint32_t* p = state;
    left = kN;
    next = 0;
    for (int j = kN - kM + 1; --j; p++) *p = p[kM] ^ twist(p[0], p[1]);
    for (int j = kM; --j; p++) *p = p[kM - kN] ^ twist(p[0], p[1]);
    *p = p[kM - kN] ^ twist(p[0], state[0]);
  }
  uint32_t draw() {
    if (--left == 0) next_state();
    uint32_t y = state[next++];
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
  }
};

// offsets inside the legacy state blob
constexpr size_t kOffLeft = 8, kOffNext = 16, kOffState = 24, kLegacyBytes = kOffState + 8 * kN;

}  // namespace

extern "C" int pcoe_host_randperm_subsets(uint8_t* rng_state, size_t state_bytes, int B, int N, int S, int32_t* out_idx) {
  using namespace pcoe;
  if (!rng_state || !out_idx) return fail(PCOE_ERR_NULL, "host_randperm_subsets: NULL pointer");
  if (state_bytes < kLegacyBytes) return fail(PCOE_ERR_BAD_SHAPE, "host_randperm_subsets: state blob of %zu bytes (< %zu)", state_bytes, kLegacyBytes);
  if (B <= 0 || N <= 0 || S <= 0 || S > N) return fail(PCOE_ERR_BAD_SHAPE, "host_randperm_subsets: B=%d N=%d S=%d", B, N, S);
  Mt mt;
  int32_t left;
  uint64_t next;
  memcpy(&left, rng_state + kOffLeft, 4);
  memcpy(&next, rng_state + kOffNext, 8);
  mt.left = left;
  mt.next = (uint32_t)next;
  for (int i = 0; i < kN; ++i) {
    uint64_t v;
    memcpy(&v, rng_state + kOffState + 8 * (size_t)i, 8);
    mt.state[i] = (uint32_t)v;
  }
  std::vector<int32_t> r((size_t)N);
  for (int b = 0; b < B; ++b) {
    for (int i = 0; i < N; ++i) r[i] = i;
    int i = 0;
    for (; i < N - 1 && i < S; ++i) {          // positions 0 .. S-1 are final after S steps
      const uint32_t z = mt.draw() % (uint32_t)(N - i);
      const int32_t t = r[i]; r[i] = r[z + i]; r[z + i] = t;
    }
    for (; i < N - 1; ++i) (void)mt.draw();    // the remaining positions only advance the generator
    memcpy(out_idx + (size_t)b * S, r.data(), sizeof(int32_t) * (size_t)S);
  }
  left = mt.left;
  next = mt.next;
  memcpy(rng_state + kOffLeft, &left, 4);
  memcpy(rng_state + kOffNext, &next, 8);
  for (int i = 0; i < kN; ++i) {
    const uint64_t v = mt.state[i];
    memcpy(rng_state + kOffState + 8 * (size_t)i, &v, 8);
  }
  return PCOE_OK;
}
