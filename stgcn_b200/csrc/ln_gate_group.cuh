// ln_gate_group.cuh -- LayerNorm backward + gate backward of the producing temporal conv with ONE CTA per (b, t) group
// (bf16 storage mode; opt-in, STGCN_LN_GROUP=1; written at the end of round 1, not yet measured on the GPU).
//
// Why: after round 1 the step is bounded by the single dependency chain of the caller's stream.  On that chain the
// two-launch version spends a whole read-only pass (ln_bwd_sums_kernel, x and dy = 120 MB for block 0) before the kernel
// that produces dz can start, because that kernel maps a thread to a fixed (vertex, channel) chunk ACROSS groups (it
// accumulates the LayerNorm parameter gradients in registers) and therefore cannot form the per-group sums itself.
// Here a CTA owns one group: x and dy of the group are loaded once into registers (16-byte chunks, all loads in flight
// together, as in ln_fwd_cached_kernel), the two sums are block-reduced, and dH -> gate derivative -> dz follows from the
// registers.  The parameter gradients dw, db -- which nothing on the critical path consumes -- are left to the existing
// ln_param_grad_kernel on helper stream q (ops.cuh).  Critical path: one kernel moving x, dy, Q in and dz out.
#pragma once
#include "simt_kernels.cuh"

namespace stgcn {
namespace simt {

constexpr int kLnGroupThreads = 512;

// NCH: 16-byte chunks per thread (M <= 512 * 8 * NCH); QONLY: GLU q-only saved state (z = Q only, h = x)
template <int ACT, bool QONLY, int NCH>
__global__ void __launch_bounds__(kLnGroupThreads, NCH <= 4 ? 2 : 1) ln_gate_bwd_group_kernel(LnGateArgs<bf16> a) {
  a.seed = live_seed(a.seed);      // + the device-side step counter, if one is registered (common.cuh)
  constexpr bool gated = ACT == STGCN_ACT_GLU || ACT == STGCN_ACT_GTU;
  constexpr bool q_only = QONLY && ACT == STGCN_ACT_GLU;
  __shared__ float red[64];
  const long long g = blockIdx.x;
  const int tid = threadIdx.x, M = a.M, C = a.C, W = a.W;
  const bf16* xp = a.x + g * M;
  const bf16* dp = a.dy + g * M;
  uint4 xr[NCH], dr[NCH];
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    const int i = (c * kLnGroupThreads + tid) * 8;
    xr[c] = make_uint4(0, 0, 0, 0); dr[c] = make_uint4(0, 0, 0, 0);
    if (i < M) {
      xr[c] = *reinterpret_cast<const uint4*>(xp + i);
      dr[c] = *reinterpret_cast<const uint4*>(dp + i);
    }
  }
  const float mu = a.mean[g], rs = a.rstd[g];
  const bool drop = a.training && a.p > 0.f;
  const float keep_scale = drop ? 1.f / (1.f - a.p) : 1.f;
  // ---- the group's two sums (same expressions as ln_bwd_sums_kernel)
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    const int i = (c * kLnGroupThreads + tid) * 8;
    if (i < M) {
      float xv[8], dv[8], wv[8];
      unpack8(xr[c], xv); unpack8(dr[c], dv); load8(a.w + i, wv);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        float d = dv[e];
        if (drop) d = dropout_keep(a.seed, (uint64_t)(g * M + i + e), a.p) ? d * keep_scale : 0.f;
        const float gi = d * wv[e];
        s1 += gi; s2 += gi * (xv[e] - mu) * rs;
      }
    }
  }
  block_sum2(s1, s2, red);
  s1 /= (float)M; s2 /= (float)M;
  // ---- dH -> gate backward -> dz
  const long long b = g / a.T_out;
  const int t = (int)(g - b * a.T_out);
  const bf16* res_slab = a.xin + ((b * a.T_in + t + a.Kt - 1) * a.N) * a.Cin;
  const bf16* zg = a.z + g * a.N * (q_only ? C : W);
  bf16* dzg = a.dz + g * a.N * W;
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    const int i = (c * kLnGroupThreads + tid) * 8;
    if (i < M) {
      const int n = i / C, c0 = i - n * C;
      float xv[8], dv[8], wv[8], zp[8], zq[8], dh[8], du[8], dq[8];
      unpack8(xr[c], xv); unpack8(dr[c], dv); load8(a.w + i, wv);
      if (q_only) {
        load8(zg + n * C + c0, zq);
      } else {
        load8(zg + n * W + c0, zp);
        if (gated) load8(zg + n * W + C + c0, zq);
        if (a.explicit_res && c0 < a.Cin) {
          float res[8];
          load8(res_slab + n * a.Cin + c0, res);
#pragma unroll
          for (int e = 0; e < 8; ++e) zp[e] += res[e];
        }
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        float d = dv[e];
        if (drop) d = dropout_keep(a.seed, (uint64_t)(g * M + i + e), a.p) ? d * keep_scale : 0.f;
        const float xh = (xv[e] - mu) * rs;
        dh[e] = rs * (d * wv[e] - s1 - xh * s2);
      }
      if (q_only) {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float sg = sigmoid_tanh_(zq[e]);
          du[e] = dh[e] * sg;
          dq[e] = dh[e] * xv[e] * (1.f - sg);
        }
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) act_bwd<true>(ACT, zp[e], gated ? zq[e] : 0.f, dh[e], du[e], dq[e]);
      }
      store8(dzg + n * W + c0, du);
      if (gated) store8(dzg + n * W + C + c0, dq);
    }
  }
}

inline bool ln_gate_group_supported(const LnGateArgs<bf16>& a) {
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  return a.G > 0 && a.G < (1LL << 31) && a.M % 8 == 0 && a.M <= kLnGroupThreads * 8 * 8 && a.C % 8 == 0 && a.W % 8 == 0 &&
         a.M == a.N * a.C && (!a.explicit_res || a.Cin % 8 == 0) && al16(a.x) && al16(a.dy) && al16(a.w) && al16(a.z) &&
         al16(a.xin) && al16(a.dz);
}

// dz only; the caller launches ln_param_grad_kernel for dw / db (helper stream)
inline void launch_ln_gate_bwd_group(int act, const LnGateArgs<bf16>& a, cudaStream_t s) {
  const bool small = a.M <= kLnGroupThreads * 8 * 4;
  const unsigned grid = (unsigned)a.G;
#define STGCN_LNG(ACTV, Q)                                                                                          \
  do {                                                                                                                \
    if (small) STGCN_LAUNCH((ln_gate_bwd_group_kernel<ACTV, Q, 4>), grid, kLnGroupThreads, 0, s, a);                 \
    else STGCN_LAUNCH((ln_gate_bwd_group_kernel<ACTV, Q, 8>), grid, kLnGroupThreads, 0, s, a);                       \
  } while (0)
  switch (act) {
    case STGCN_ACT_GLU:
      if (a.q_only) STGCN_LNG(STGCN_ACT_GLU, true); else STGCN_LNG(STGCN_ACT_GLU, false);
      break;
    case STGCN_ACT_GTU: STGCN_LNG(STGCN_ACT_GTU, false); break;
    case STGCN_ACT_RELU: STGCN_LNG(STGCN_ACT_RELU, false); break;
    case STGCN_ACT_SILU: STGCN_LNG(STGCN_ACT_SILU, false); break;
    default: STGCN_LNG(STGCN_ACT_LINEAR, false); break;
  }
#undef STGCN_LNG
}

}  // namespace simt
}  // namespace stgcn
