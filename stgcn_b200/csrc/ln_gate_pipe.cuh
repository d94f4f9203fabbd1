// ln_gate_pipe.cuh -- LayerNorm backward + LayerNorm parameter gradients + gate backward of the producing temporal
// conv (layers.py:255-256 and :92-115, backward) as ONE persistent, bulk-copy-pipelined kernel (bf16 storage mode).
//
// The two-launch version (simt_kernels.cuh: ln_bwd_sums_kernel + ln_gate_bwd_kernel) reads x and dy twice and keeps
// only what 127 registers x 16 warps can hold in flight (long-scoreboard bound, 2.7 TB/s).  Here a CTA owns a contiguous
// range of (b, t) groups.  For every group its x (pre-LayerNorm activations), dy and the residual slab of the conv input
// arrive in shared memory by cp.async.bulk (1-D TMA), two groups ahead, signalled by an mbarrier; the group's two
// LayerNorm sums are reduced from shared memory (no second HBM read), then each thread walks its 8-channel chunks:
// dH from the group scalars -> through the gate derivative (P/Q pre-activations straight from global, prefetched two
// chunks ahead in registers) -> dz stored, while dw / db of its chunks accumulate in registers across the CTA's groups
// (one atomic per element per CTA at the end, as before).
#pragma once
#include "umma.cuh"
#include "simt_kernels.cuh"

namespace stgcn {
namespace simt {

constexpr int kLnPipeThreads = 512;
constexpr int kLnPipeChunks = 4;                                   // 8-element chunks per thread per group
constexpr int kLnPipeMaxM = kLnPipeThreads * 8 * kLnPipeChunks;    // 16384 elements per group

__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   umma::smem_u32(dst_smem)),
               "l"(src), "r"(bytes), "r"(umma::smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ uint4 ldg16(const bf16* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }

struct LnPipeGeom { uint32_t m_bytes, res_bytes, m_pad, stage_bytes; size_t smem; };
inline LnPipeGeom ln_pipe_geom(int M, int N, int Cin, int explicit_res) {
  LnPipeGeom g;
  g.m_bytes = (uint32_t)M * 2;
  g.res_bytes = explicit_res ? (uint32_t)N * Cin * 2 : 0;
  g.m_pad = (g.m_bytes + 127) & ~127u;
  g.stage_bytes = 2 * g.m_pad + ((g.res_bytes + 127) & ~127u);
  g.smem = 2 * (size_t)g.stage_bytes + 128;
  return g;
}

template <int ACT, bool QONLY>
__global__ void __launch_bounds__(kLnPipeThreads, 1) ln_gate_bwd_pipe_kernel(LnGateArgs<bf16> a, LnPipeGeom geo) {
  pdl_begin();
  constexpr bool gated = ACT == STGCN_ACT_GLU || ACT == STGCN_ACT_GTU;
  extern __shared__ __align__(128) uint8_t smraw[];
  uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smraw) + 127) & ~uintptr_t(127));
  __shared__ __align__(8) uint64_t full[2];
  __shared__ float red[64];
  const int tid = threadIdx.x;
  const long long g0 = (long long)blockIdx.x * a.groups_per_cta;
  const long long g1 = min(a.G, g0 + a.groups_per_cta);
  if (g0 >= g1) return;                                            // CTA-uniform
  if (tid == 0) {
    umma::mbar_init(&full[0], 1);
    umma::mbar_init(&full[1], 1);
    umma::fence_barrier_init();
  }
  __syncthreads();
  const int M = a.M, C = a.C, W = a.W;
  constexpr bool q_only = QONLY && ACT == STGCN_ACT_GLU;     // z = Q only, [rows, C]; h = x (the LayerNorm input)
  const int zrow = q_only ? C : W;                             // channels per row of the saved tensor
  auto issue = [&](long long g, int s) {                           // thread 0 only
    uint8_t* st = sm + (size_t)s * geo.stage_bytes;
    umma::mbar_arrive_expect_tx(&full[s], 2 * geo.m_bytes + geo.res_bytes);
    bulk_g2s(st, a.x + g * M, geo.m_bytes, &full[s]);
    bulk_g2s(st + geo.m_pad, a.dy + g * M, geo.m_bytes, &full[s]);
    if (geo.res_bytes) {
      const long long b = g / a.T_out;
      const int t = (int)(g - b * a.T_out);
      bulk_g2s(st + 2 * geo.m_pad, a.xin + ((b * a.T_in + t + a.Kt - 1) * a.N) * a.Cin, geo.res_bytes, &full[s]);
    }
  };
  if (tid == 0) {
    issue(g0, 0);
    if (g0 + 1 < g1) issue(g0 + 1, 1);
  }
  // this thread's chunks: element offset i, vertex n, first channel c0 (fixed for the whole kernel)
  int ci[kLnPipeChunks], cz[kLnPipeChunks], cdz[kLnPipeChunks], cres[kLnPipeChunks];   // cz / cdz: chunk offset inside the group's z / dz slab
#pragma unroll
  for (int k = 0; k < kLnPipeChunks; ++k) {
    const int i = (k * kLnPipeThreads + tid) * 8;
    ci[k] = i < M ? i : -1;
    const int n = i / C, c0 = i - n * C;
    cz[k] = q_only ? n * C + c0 : n * W + c0;
    cdz[k] = n * W + c0;
    cres[k] = (geo.res_bytes && c0 < a.Cin) ? n * a.Cin + c0 : -1;
  }
  float aw[kLnPipeChunks][8], ab[kLnPipeChunks][8];
#pragma unroll
  for (int k = 0; k < kLnPipeChunks; ++k)
#pragma unroll
    for (int e = 0; e < 8; ++e) { aw[k][e] = 0.f; ab[k][e] = 0.f; }
  const bool drop = a.training && a.p > 0.f;
  const float keep_scale = drop ? 1.f / (1.f - a.p) : 1.f;
  const float inv_m = 1.f / (float)M;

  uint32_t it = 0;
  for (long long g = g0; g < g1; ++g, ++it) {
    const int s = it & 1;
    const uint32_t ph = (it >> 1) & 1;
    const bf16* zg = a.z + g * a.N * zrow;
    bf16* dzg = a.dz + g * a.N * W;
    // z of the first two chunks is requested before anything else: the wait and the reduction hide its latency
    uint4 zq_p[2], zq_q[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      zq_p[k] = make_uint4(0, 0, 0, 0); zq_q[k] = make_uint4(0, 0, 0, 0);
      if (ci[k] >= 0) {
        zq_p[k] = ldg16(zg + cz[k]);                               // q_only: this IS the Q chunk
        if (gated && !q_only) zq_q[k] = ldg16(zg + cz[k] + C);
      }
    }
    const float mu = a.mean[g], rs = a.rstd[g];
    umma::mbar_wait(&full[s], ph);
    const uint8_t* st = sm + (size_t)s * geo.stage_bytes;
    const bf16* xs = reinterpret_cast<const bf16*>(st);
    const bf16* ds = reinterpret_cast<const bf16*>(st + geo.m_pad);
    const bf16* rsd = reinterpret_cast<const bf16*>(st + 2 * geo.m_pad);
    // ---- phase 1: the group's two sums (same expressions as ln_bwd_sums_kernel)
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int k = 0; k < kLnPipeChunks; ++k) {
      if (ci[k] >= 0) {
        float xv[8], dv[8], wv[8];
        unpack8(*reinterpret_cast<const uint4*>(xs + ci[k]), xv);
        unpack8(*reinterpret_cast<const uint4*>(ds + ci[k]), dv);
        load8(a.w + ci[k], wv);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          float d = dv[e];
          if (drop) d = dropout_keep(a.seed, (uint64_t)(g * M + ci[k] + e), a.p) ? d * keep_scale : 0.f;
          const float gi = d * wv[e];
          s1 += gi; s2 += gi * (xv[e] - mu) * rs;
        }
      }
    }
    block_sum2(s1, s2, red);
    s1 *= inv_m; s2 *= inv_m;
    // ---- phase 2: dH -> gate backward -> dz; parameter gradients accumulate
#pragma unroll
    for (int k = 0; k < kLnPipeChunks; ++k) {
      const uint4 zp_raw = zq_p[k & 1], zq_raw = zq_q[k & 1];
      if (k + 2 < kLnPipeChunks && ci[k + 2] >= 0) {               // refill the slot two chunks ahead
        zq_p[k & 1] = ldg16(zg + cz[k + 2]);
        if (gated && !q_only) zq_q[k & 1] = ldg16(zg + cz[k + 2] + C);
      }
      if (ci[k] >= 0) {
        float xv[8], dv[8], wv[8], zp[8], zq[8], dh[8], du[8], dq[8];
        unpack8(*reinterpret_cast<const uint4*>(xs + ci[k]), xv);
        unpack8(*reinterpret_cast<const uint4*>(ds + ci[k]), dv);
        load8(a.w + ci[k], wv);
        if (q_only) unpack8(zp_raw, zq); else unpack8(zp_raw, zp);
        if (gated && !q_only) unpack8(zq_raw, zq);
        if (cres[k] >= 0 && !q_only) {
          float res[8];
          unpack8(*reinterpret_cast<const uint4*>(rsd + cres[k]), res);
#pragma unroll
          for (int e = 0; e < 8; ++e) zp[e] += res[e];
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          float d = dv[e];
          if (drop) d = dropout_keep(a.seed, (uint64_t)(g * M + ci[k] + e), a.p) ? d * keep_scale : 0.f;
          const float xh = (xv[e] - mu) * rs;
          dh[e] = rs * (d * wv[e] - s1 - xh * s2);
          aw[k][e] += d * xh;
          ab[k][e] += d;
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
        store8(dzg + cdz[k], du);
        if (gated) store8(dzg + cdz[k] + C, dq);
      }
    }
    __syncthreads();                                               // every thread is done with stage s
    if (tid == 0 && g + 2 < g1) issue(g + 2, s);
  }
#pragma unroll
  for (int k = 0; k < kLnPipeChunks; ++k) {
    if (ci[k] >= 0) {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        if (a.dw) atomicAdd(a.dw + ci[k] + e, aw[k][e]);
        if (a.db) atomicAdd(a.db + ci[k] + e, ab[k][e]);
      }
    }
  }
}

inline bool ln_gate_pipe_supported(const LnGateArgs<bf16>& a) {
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  if (!(a.G > 0 && a.M % 8 == 0 && a.M <= kLnPipeMaxM && a.C % 8 == 0 && a.W % 8 == 0 && a.M == a.N * a.C)) return false;
  if (a.explicit_res && (a.Cin % 8 != 0 || ((long long)a.N * a.Cin * 2) % 16 != 0)) return false;
  if (!(al16(a.x) && al16(a.dy) && al16(a.w) && al16(a.z) && al16(a.xin) && al16(a.dz))) return false;
  return ln_pipe_geom(a.M, a.N, a.Cin, a.explicit_res && !a.q_only).smem <= 225 * 1024;
}

inline void launch_ln_gate_bwd_pipe(int act, LnGateArgs<bf16> a, int sms, cudaStream_t s) {
  const LnPipeGeom geo = ln_pipe_geom(a.M, a.N, a.Cin, a.explicit_res && !a.q_only);      // q-only: no residual slab needed
  a.groups_per_cta = ceil_div(a.G, sms);
  const int grid = ceil_div(a.G, a.groups_per_cta);
  auto go = [&](auto kern, const char* name) {
    STGCN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)geo.smem));
    STGCN_LAUNCH_NAMED(name, kern, grid, kLnPipeThreads, geo.smem, s, a, geo);
  };
  switch (act) {
    case STGCN_ACT_GLU:
      if (a.q_only) go(ln_gate_bwd_pipe_kernel<STGCN_ACT_GLU, true>, "ln_gate_bwd_pipe_kernel<GLU,q>");
      else go(ln_gate_bwd_pipe_kernel<STGCN_ACT_GLU, false>, "ln_gate_bwd_pipe_kernel<GLU>");
      break;
    case STGCN_ACT_GTU: go(ln_gate_bwd_pipe_kernel<STGCN_ACT_GTU, false>, "ln_gate_bwd_pipe_kernel<GTU>"); break;
    case STGCN_ACT_RELU: go(ln_gate_bwd_pipe_kernel<STGCN_ACT_RELU, false>, "ln_gate_bwd_pipe_kernel<RELU>"); break;
    case STGCN_ACT_SILU: go(ln_gate_bwd_pipe_kernel<STGCN_ACT_SILU, false>, "ln_gate_bwd_pipe_kernel<SILU>"); break;
    default: go(ln_gate_bwd_pipe_kernel<STGCN_ACT_LINEAR, false>, "ln_gate_bwd_pipe_kernel<LINEAR>"); break;
  }
}

}  // namespace simt
}  // namespace stgcn
