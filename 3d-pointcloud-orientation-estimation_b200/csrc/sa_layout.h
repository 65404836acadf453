// sa_layout.h — byte offsets inside the `saved` and `workspace` buffers of pcoe_sa_forward /
// pcoe_sa_backward.  Everything is 256-byte aligned; forward and backward share one workspace
// (the regions overlap in time, not in use).
#pragma once
#include "common.cuh"
#include <cstdlib>

namespace pcoe {

// Grid-wide reductions (BatchNorm batch sums, weight gradients) are accumulated into kRedCopies
// interleaved copies - CTA k adds into copy k % kRedCopies - and the consumer adds the copies up.
// All CTAs of a persistent kernel flush at the same time; atomics that land in the same 32-byte L2
// sector are serialised, so with one copy the tail of every kernel was ~10 us of 148-deep atomic
// chains (measured: in-kernel clock trace, profiles/README.md).
constexpr int kRedCopies = 8;
// largest dynamic shared memory a v4 kernel may ask for (227 KB opt-in limit minus the kernels' static barriers)
constexpr size_t kSmemMax4 = 232448 - 256;

struct SaLayout {
  int M, G;        // rows = B*S*K, groups = B*S
  size_t esz;      // bytes per stored activation element (4 fp32, 2 bf16)
  // saved (train only)
  size_t sv_y[3], sv_stat[3], sv_slot, sv_ysel, saved_bytes;
  // workspace, forward
  size_t ws_sums[3], ws_sums_bytes, ws_ymax, ws_ymin, ws_amax, ws_amin, ws_y[2], ws_stat[3];
  // workspace, backward
  size_t wb_sums[3], wb_sums_bytes, wb_consts[3], wb_gm, wb_dz[2];
  // weight-gradient copies of the v4 kernels: [kRedCopies][C_l][dwc_ld[l]] fp32, layer-1 columns in [feats | xyz] order
  size_t wb_dwc[3], wb_dwc_bytes;
  size_t wb_g0;    // [kRedCopies][16] fp32: x0^T x0 and sum x0 of the MaskStatsW1 path (zeroed with the dW copies)
  int dwc_ld[3];
  // bf16 weight copies (tensor-core path): offsets relative to `saved` (train) or `workspace` (eval)
  size_t wb_off[3], wbt_off[3];
  int wb_rows[3], wb_k[3], wbt_rows[3], wbt_k[3];
  bool v2;         // bf16 mode and the layer fits the persistent channel-on-lane kernels (sa_tc4.cuh):
                   // activations are then stored channel-major [C][Mld], one weight image [Rp][Kp] per layer
  bool v5;         // bf16 mode, wide layers that do not fit v4 (SA3): K-chunk-streamed kernels (sa_tc5.cuh), same
                   // HBM layouts as v2
  bool v6;         // bf16x3 mode (sa_tc6.cuh): fp32 tile-blocked channel-major activations [tile][C][128], two bf16
                   // planes (hi, mid, lo) of every weight image
  int Mld;         // rows rounded up to 128 (v2 / v5)
  int w4_rp[3], w4_kp[3];
  // v2 train: the last layer's pre-activations y3 are NOT stored; the forward accumulates the Gram matrix of the
  // layer's input instead (sa_tc4.cuh, DySparse4).  sv_gram: [kRedCopies][C2][gram_ld] fp32 in `saved`;
  // backward workspace: Gm image bf16 [128][C2], r [C2], summed Gram [C2][gram_ld], dense weight-gradient term E [C3][C2].
  bool l3s;
  int gram_ld;
  size_t sv_gram, sv_gram_bytes, wb_gmimg, wb_rvec, wb_gsum, wb_l3e;
  size_t workspace_bytes;
};

inline SaLayout sa_layout(const pcoe_sa_desc& d) {
  SaLayout L;
  L.M = d.B * d.S * d.K;
  L.G = d.B * d.S;
  L.esz = d.precision == PCOE_PRECISION_BF16 ? 2 : 4;
  const int C[3] = {d.C1, d.C2, d.C3};
  auto take = [](size_t& cur, size_t bytes) { size_t o = cur; cur = align_up(cur + bytes, 256); return o; };

  const bool tc = d.precision == PCOE_PRECISION_BF16;
  const bool x3 = d.precision == PCOE_PRECISION_BF16X3;
  const int Kin[3] = {3 + d.D, d.C1, d.C2};
  auto chan_ok = [](int c) { return c == 64 || c == 128 || c == 256; };
  L.v2 = tc && d.K == 32 && chan_ok(d.C1) && chan_ok(d.C2) && chan_ok(d.C3) && (d.D == 0 || d.D == 32 || d.D == 64 || d.D == 128);
  L.v5 = tc && !L.v2 && d.K == 32 && d.C1 % 128 == 0 && d.C2 % 128 == 0 && d.C3 % 128 == 0 && d.D > 0 && d.D % 128 == 0;
  L.v6 = x3;                      // shape support is checked by sa_x3_supported() (sa.cu)
  const bool cm = L.v2 || L.v5 || L.v6;   // tile-blocked channel-major activations + one weight image per layer
  L.Mld = (int)align_up(L.M, 128);
  const size_t rows_ld = cm ? (size_t)L.Mld : (size_t)L.M;
  for (int l = 0; l < 3; ++l) {
    L.w4_rp[l] = (int)align_up(C[l], 128);
    L.w4_kp[l] = (int)align_up(l == 0 ? (Kin[0] + 15) / 16 * 16 : Kin[l], 128);
  }

  // sparse last-layer backward: needs whole tiles (tail columns would pollute the Gram matrix), C2 <= 128 (one Gm
  // M tile) and the backward kernel's shared memory: W3 image + Gm image + P tile + Q tile + constants + one staging tile
  L.gram_ld = d.C2 + 16;
  {
    const size_t need = 1024 + (size_t)2 * L.w4_rp[2] * L.w4_kp[2] + (size_t)2 * 128 * d.C2 +
                        (size_t)2 * (d.C3 < 128 ? 128 : d.C3) * 128 + (size_t)2 * d.C2 * 128 + (size_t)(8 * d.C2 + 512) +
                        (size_t)d.C2 * 256;
    // PCOE_SA_STORE_Y3=1 (debug / A-B tests): keep the older formulation that stores y3 and re-reads it in backward
    const char* keep = getenv("PCOE_SA_STORE_Y3");
    L.l3s = L.v2 && d.train && L.M % 128 == 0 && d.C2 <= 128 && need <= kSmemMax4 &&
            L.w4_rp[2] + L.gram_ld <= 512 &&   // forward TMEM: one accumulator buffer (pad128(C3) columns) + Gram
            !(keep && keep[0] == '1');
  }

  size_t s = 0;
  // v6 (bf16x3): y3 is stored as bf16 (sa_tc6.cuh, act16_off) - its only reader is the dense BatchNorm-backward term
  for (int l = 0; l < 3; ++l) L.sv_y[l] = (l == 2 && L.l3s) ? 0 : take(s, rows_ld * C[l] * ((l == 2 && L.v6) ? 2 : L.esz));
  L.sv_gram = L.sv_gram_bytes = 0;
  if (L.l3s) {
    L.sv_gram_bytes = sizeof(float) * (size_t)kRedCopies * d.C2 * L.gram_ld;
    L.sv_gram = take(s, L.sv_gram_bytes);
  }
  for (int l = 0; l < 3; ++l) L.sv_stat[l] = take(s, sizeof(float) * 4 * C[l]);
  L.sv_slot = take(s, (size_t)L.G * d.C3);
  L.sv_ysel = take(s, sizeof(float) * (size_t)L.G * d.C3);
  auto kpad = [](int k) { return k <= 64 ? 64 : k <= 128 ? 128 : k <= 256 ? 256 : (int)align_up(k, 64); };
  for (int l = 0; l < 3; ++l) {
    L.wb_rows[l] = (int)align_up(C[l], 128);   L.wb_k[l] = kpad(Kin[l]);
    L.wbt_rows[l] = (int)align_up(Kin[l], 128); L.wbt_k[l] = kpad(C[l]);
    L.wb_off[l] = L.wbt_off[l] = 0;
  }
  const size_t planes = x3 ? 3 : 1;   // bf16x3: hi, mid, lo planes (backward reads the first two)
  if ((tc || x3) && d.train)
    for (int l = 0; l < 3; ++l) {
      if (cm) { L.wb_off[l] = take(s, planes * 2 * L.w4_rp[l] * L.w4_kp[l]); continue; }
      L.wb_off[l] = take(s, (size_t)2 * L.wb_rows[l] * L.wb_k[l]);
      L.wbt_off[l] = take(s, (size_t)2 * L.wbt_rows[l] * L.wbt_k[l]);
    }
  L.saved_bytes = d.train ? s : 0;

  size_t f = 0;
  for (int l = 0; l < 3; ++l) L.ws_sums[l] = take(f, sizeof(double) * 2 * C[l] * kRedCopies);
  L.ws_sums_bytes = f - L.ws_sums[0];
  L.ws_ymax = take(f, sizeof(float) * (size_t)L.G * d.C3);
  L.ws_ymin = take(f, sizeof(float) * (size_t)L.G * d.C3);
  L.ws_amax = take(f, (size_t)L.G * d.C3);
  L.ws_amin = take(f, (size_t)L.G * d.C3);
  for (int l = 0; l < 3; ++l) L.ws_stat[l] = take(f, sizeof(float) * 4 * C[l]);
  L.ws_y[0] = L.ws_y[1] = 0;
  if (!d.train) {
    for (int l = 0; l < 2; ++l) L.ws_y[l] = take(f, rows_ld * C[l] * L.esz);
    if (tc || x3)
      for (int l = 0; l < 3; ++l) {
        if (cm) { L.wb_off[l] = take(f, planes * 2 * L.w4_rp[l] * L.w4_kp[l]); continue; }
        L.wb_off[l] = take(f, (size_t)2 * L.wb_rows[l] * L.wb_k[l]);
        L.wbt_off[l] = take(f, (size_t)2 * L.wbt_rows[l] * L.wbt_k[l]);
      }
  }

  size_t b = 0;
  for (int l = 0; l < 3; ++l) L.wb_sums[l] = take(b, sizeof(double) * 2 * C[l] * kRedCopies);
  L.wb_sums_bytes = b - L.wb_sums[0];
  for (int l = 0; l < 3; ++l) {
    L.dwc_ld[l] = l == 0 ? (Kin[0] + 15) / 16 * 16 : Kin[l];       // = the Q operand's channel extent (multiple of 16)
    // v6 (bf16x3), no input features: layer 1's dz1^T x0 accumulator of the MaskStatsW6 epilogue, [kRedCopies][C1][4]
    const bool w6 = L.v6 && l == 0 && d.D == 0 && !d.group_all && d.C1 == 64;
    L.wb_dwc[l] = L.v2 ? take(b, sizeof(float) * (size_t)kRedCopies * C[l] * L.dwc_ld[l])
                       : (w6 ? take(b, sizeof(float) * (size_t)kRedCopies * C[l] * 4) : 0);
  }
  const bool w6 = L.v6 && d.D == 0 && !d.group_all && d.C1 == 64;
  L.wb_g0 = (L.v2 || w6) ? take(b, sizeof(float) * 16 * kRedCopies) : 0;
  L.wb_dwc_bytes = L.v2 ? b - L.wb_dwc[0] : (w6 ? b - L.wb_dwc[0] : 0);
  for (int l = 0; l < 3; ++l) L.wb_consts[l] = take(b, sizeof(float) * 3 * C[l]);
  L.wb_gmimg = L.wb_rvec = L.wb_gsum = L.wb_l3e = 0;
  if (L.l3s) {
    L.wb_gmimg = take(b, (size_t)2 * 128 * d.C2);
    L.wb_rvec = take(b, sizeof(float) * d.C2);
    L.wb_gsum = take(b, sizeof(float) * (size_t)d.C2 * L.gram_ld);
    L.wb_l3e = take(b, sizeof(float) * (size_t)d.C3 * d.C2);
  }
  L.wb_gm = take(b, sizeof(float) * (size_t)L.G * d.C3);
  for (int l = 0; l < 2; ++l) L.wb_dz[l] = take(b, rows_ld * C[l] * L.esz);
  if (!d.train) b = 0;

  L.workspace_bytes = f > b ? f : b;
  return L;
}

}  // namespace pcoe
