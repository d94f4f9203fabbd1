// umma_gso.cuh -- tcgen05 node contraction for the bf16 path (ChebGraphConv/GraphConv, layers.py:154-161,198):
//
//   out[g, h, c] = alpha * sum_i Lhat[h, i] * in[g, i, c] + beta * aux[g, h, c]      g = (b, t) group, c < C
//
// as the GEMM  D[h, (g, c)] = Lhat[h, :] . X[:, (g, c)]  with
//   A = Lhat (bf16, K-major, zero padded to a multiple of 64 columns), one 128-row tile resident in shared
//       memory per CTA (blockIdx.y = row tile), loaded once by TMA with 128B swizzle;
//   B = in viewed as [i][(g, c)]: MN-major, one swizzle-atom-wide chunk (C elements) per group, streamed through a
//       TMA ring in 64-row K blocks (rows past N are zero-filled by TMA, which also pads K);
//   D = fp32 in TMEM, 128 lanes x (Gb*C <= 256) columns, double buffered against the epilogue warps.
// Same warp roles as umma_tap.cuh.  Lhat is constant (no gradient: torch.from_numpy, main.py:103), so the
// backward pass is this kernel again on Lhat^T.
#pragma once
#include "umma_tap.cuh"

namespace stgcn {
namespace umma {

struct GsoParams {
  int N, C, Gb, nKB, S;
  long long G;
  int n_sets;
  uint32_t a_bytes, stage_bytes, b_swz, b_lbo, b_sbo, b_kadv;
  float alpha, beta;
  const bf16* aux;
  bf16* out;
};

__global__ void __launch_bounds__(kTapThreads, 1)
umma_gso_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, GsoParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* a_s = smem;                    // [nKB][128 rows][128 B]
  uint8_t* ring = smem + p.a_bytes;       // S stages of [Gb][64 rows][C*2 B]
  __shared__ __align__(8) uint64_t full[kMaxStages], empty[kMaxStages], afull, tfull[2], tempty[2];
  __shared__ uint32_t tmem_base_s;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h0 = blockIdx.y * 128;
  const int NC = p.Gb * p.C;              // accumulator width
  uint32_t ncols = 32;
  while ((int)ncols < 2 * NC) ncols <<= 1;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.S; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(&afull, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 4); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&tmem_base_s, ncols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (warp == 0) {
    if (lane == 0) {
      tma_prefetch_desc(&tmA);
      tma_prefetch_desc(&tmB);
      mbar_arrive_expect_tx(&afull, p.a_bytes);
      for (int kb = 0; kb < p.nKB; ++kb) tma_load_2d(a_s + (size_t)kb * 16384, &tmA, &afull, kb * 64, h0);
      RingPos rp{0, 0};                               // no integer division in the single-thread roles (umma.cuh)
      for (int set = blockIdx.x; set < p.n_sets; set += gridDim.x) {
        for (int kb = 0; kb < p.nKB; ++kb, rp.advance(p.S)) {
          const uint32_t s = rp.s, ph = rp.ph;
          mbar_wait(&empty[s], ph ^ 1);
          mbar_arrive_expect_tx(&full[s], p.stage_bytes);
          tma_load_3d(ring + (size_t)s * p.stage_bytes, &tmB, &full[s], 0, kb * 64, set * p.Gb);
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {      // one elected lane, known to the compiler as such (issue cost: see umma.cuh)
      const uint32_t idesc = make_idesc_bf16(128, NC, 0, 1);
      mbar_wait(&afull, 0);
      uint32_t acc_cnt = 0;
      RingPos rp{0, 0};
      for (int set = blockIdx.x; set < p.n_sets; set += gridDim.x, ++acc_cnt) {
        const uint32_t ab = acc_cnt & 1, aph = (acc_cnt >> 1) & 1;
        mbar_wait(&tempty[ab], aph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + ab * NC;
        for (int kb = 0; kb < p.nKB; ++kb, rp.advance(p.S)) {
          const uint32_t s = rp.s, ph = rp.ph;
          mbar_wait(&full[s], ph);
          tc_fence_after();
          const uint32_t a_base = smem_u32(a_s + (size_t)kb * 16384);
          const uint32_t b_base = smem_u32(ring + (size_t)s * p.stage_bytes);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t da = make_smem_desc(a_base + k * 32, 16, 1024, SWZ_128B);
            const uint64_t db = make_smem_desc(b_base + k * p.b_kadv, p.b_lbo, p.b_sbo, p.b_swz);
            mma_bf16_ss(d_tmem, da, db, idesc, (kb | k) != 0);
          }
          mma_commit(&empty[s]);
        }
        mma_commit(&tfull[ab]);
      }
    }
  } else {
    const int q = warp & 3;
    const int h = h0 + q * 32 + lane;
    const bool hvalid = h < p.N;
    uint32_t acc_cnt = 0;
    for (int set = blockIdx.x; set < p.n_sets; set += gridDim.x, ++acc_cnt) {
      const uint32_t ab = acc_cnt & 1, aph = (acc_cnt >> 1) & 1;
      mbar_wait(&tfull[ab], aph);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + ab * NC;
      for (int gl = 0; gl < p.Gb; ++gl) {
        const long long g = (long long)set * p.Gb + gl;
        const bool ok = hvalid && g < p.G;
        const long long base = (g * p.N + h) * p.C;
        for (int c0 = 0; c0 < p.C; c0 += 16) {
          uint32_t r[16];
          tmem_ld_32x32b_x16(t_addr + gl * p.C + c0, r);
          tmem_ld_wait();
          if (ok) {
            float v[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = p.alpha * __uint_as_float(r[i]);
            if (p.aux) {
              float av[16];
              load16_bf16(p.aux + base + c0, av);
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] += p.beta * av[i];
            }
            store16_bf16(p.out + base + c0, v);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[ab]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, ncols);
}

// ------------------------------------------------------------------------------------------------
// K-tiled variant for operators that do not fit shared memory (N > ~1500; BASELINE configs[4], N = 2048): the A tile
// [128 x 64] of every K block travels through the ring next to its B block instead of staying resident.  Lhat (8 MB
// in bf16 at N = 2048) stays in the 126 MB L2, so the re-reads per group set are L2 traffic.  Serves the shapes the
// resident-operator kernel cannot (STGCN_GSO_KTILED=1 forces it for every shape: test knob).  Same roles / accumulator
// handling as umma_gso_kernel.
// Next step for the roofline at N = 2048: cta_group::2 (M = 256) with TMA multicast of the B block across the pair.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kTapThreads, 1)
umma_gso_ktiled_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, GsoParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* ring = smem;                   // S stages of { A [128 rows][128 B] (16 KB) , B [Gb][64 rows][C*2 B] }
  __shared__ __align__(8) uint64_t full[kMaxStages], empty[kMaxStages], tfull[2], tempty[2];
  __shared__ uint32_t tmem_base_s;
  const uint32_t stage_total = 16384u + p.stage_bytes;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h0 = blockIdx.y * 128;
  const int NC = p.Gb * p.C;              // accumulator width
  uint32_t ncols = 32;
  while ((int)ncols < 2 * NC) ncols <<= 1;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.S; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 4); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&tmem_base_s, ncols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (warp == 0) {
    if (lane == 0) {
      tma_prefetch_desc(&tmA);
      tma_prefetch_desc(&tmB);
      RingPos rp{0, 0};                               // no integer division in the single-thread roles (umma.cuh)
      for (int set = blockIdx.x; set < p.n_sets; set += gridDim.x) {
        for (int kb = 0; kb < p.nKB; ++kb, rp.advance(p.S)) {
          const uint32_t s = rp.s, ph = rp.ph;
          mbar_wait(&empty[s], ph ^ 1);
          mbar_arrive_expect_tx(&full[s], stage_total);
          uint8_t* st = ring + (size_t)s * stage_total;
          tma_load_2d(st, &tmA, &full[s], kb * 64, h0);
          tma_load_3d(st + 16384, &tmB, &full[s], 0, kb * 64, set * p.Gb);
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      const uint32_t idesc = make_idesc_bf16(128, NC, 0, 1);
      uint32_t acc_cnt = 0;
      RingPos rp{0, 0};
      for (int set = blockIdx.x; set < p.n_sets; set += gridDim.x, ++acc_cnt) {
        const uint32_t ab = acc_cnt & 1, aph = (acc_cnt >> 1) & 1;
        mbar_wait(&tempty[ab], aph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + ab * NC;
        for (int kb = 0; kb < p.nKB; ++kb, rp.advance(p.S)) {
          const uint32_t s = rp.s, ph = rp.ph;
          mbar_wait(&full[s], ph);
          tc_fence_after();
          const uint32_t a_base = smem_u32(ring + (size_t)s * stage_total);
          const uint32_t b_base = a_base + 16384;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t da = make_smem_desc(a_base + k * 32, 16, 1024, SWZ_128B);
            const uint64_t db = make_smem_desc(b_base + k * p.b_kadv, p.b_lbo, p.b_sbo, p.b_swz);
            mma_bf16_ss(d_tmem, da, db, idesc, (kb | k) != 0);
          }
          mma_commit(&empty[s]);
        }
        mma_commit(&tfull[ab]);
      }
    }
  } else {
    const int q = warp & 3;
    const int h = h0 + q * 32 + lane;
    const bool hvalid = h < p.N;
    uint32_t acc_cnt = 0;
    for (int set = blockIdx.x; set < p.n_sets; set += gridDim.x, ++acc_cnt) {
      const uint32_t ab = acc_cnt & 1, aph = (acc_cnt >> 1) & 1;
      mbar_wait(&tfull[ab], aph);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + ab * NC;
      for (int gl = 0; gl < p.Gb; ++gl) {
        const long long g = (long long)set * p.Gb + gl;
        const bool ok = hvalid && g < p.G;
        const long long base = (g * p.N + h) * p.C;
        for (int c0 = 0; c0 < p.C; c0 += 16) {
          uint32_t r[16];
          tmem_ld_32x32b_x16(t_addr + gl * p.C + c0, r);
          tmem_ld_wait();
          if (ok) {
            float v[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = p.alpha * __uint_as_float(r[i]);
            if (p.aux) {
              float av[16];
              load16_bf16(p.aux + base + c0, av);
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] += p.beta * av[i];
            }
            store16_bf16(p.out + base + c0, v);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[ab]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, ncols);
}

// Lhat (fp32 [N,N]) -> bf16 [N][Kp], Kp = N rounded up to 64, zero padded; optionally transposed
__global__ void gso_prep_kernel(const float* M, bf16* out, int N, int Kp, int trans) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= N * Kp) return;
  int h = idx / Kp, i = idx - h * Kp;
  float v = 0.f;
  if (i < N) v = trans ? M[(long long)i * N + h] : M[(long long)h * N + i];
  out[idx] = __float2bfloat16_rn(v);
}

struct GsoPlan { bool ok; int Kp, nKB, Gb, S, nMT; uint32_t a_bytes, stage_bytes; size_t smem; };

inline GsoPlan plan_gso(int N, int C) {
  GsoPlan pl{};
  pl.ok = false;
  if (C != 16 && C != 32 && C != 64) return pl;
  pl.Kp = (N + 63) / 64 * 64;
  pl.nKB = pl.Kp / 64;
  pl.a_bytes = (uint32_t)pl.nKB * 16384;
  pl.Gb = 256 / C;
  pl.stage_bytes = (uint32_t)pl.Gb * 64 * C * 2;      // = 32 KB
  if (pl.a_bytes + 2 * (size_t)pl.stage_bytes > kSmemBudget) return pl;
  int S = (int)((kSmemBudget - pl.a_bytes) / pl.stage_bytes);
  pl.S = S > kMaxStages ? kMaxStages : S;
  pl.nMT = (N + 127) / 128;
  pl.smem = pl.a_bytes + (size_t)pl.S * pl.stage_bytes + 1024;
  pl.ok = true;
  return pl;
}
// test knob: route EVERY shape through the K-tiled kernel (the small-N graph-conv tests then exercise it); by default it
// serves only the shapes the resident-operator kernel cannot take
inline bool gso_ktiled_forced() {
  static const bool on = std::getenv("STGCN_GSO_KTILED") != nullptr;
  return on;
}
// K-tiled plan: stages of (16 KB A block + B block); nothing resident
inline GsoPlan plan_gso_ktiled(int N, int C) {
  GsoPlan pl{};
  pl.ok = false;
  if (C != 16 && C != 32 && C != 64) return pl;
  pl.Kp = (N + 63) / 64 * 64;
  pl.nKB = pl.Kp / 64;
  pl.a_bytes = 0;
  pl.Gb = 256 / C;
  pl.stage_bytes = (uint32_t)pl.Gb * 64 * C * 2;      // = 32 KB
  int S = (int)(kSmemBudget / (16384 + pl.stage_bytes));
  pl.S = S > kMaxStages ? kMaxStages : S;
  if (pl.S < 2) return pl;
  pl.nMT = (N + 127) / 128;
  pl.smem = (size_t)pl.S * (16384 + pl.stage_bytes) + 1024;
  pl.ok = true;
  return pl;
}
inline bool gso_supported(int N, int C, long long G) {
  if (G <= 0) return false;
  if (gso_ktiled_forced()) return plan_gso_ktiled(N, C).ok;
  return plan_gso(N, C).ok || plan_gso_ktiled(N, C).ok;
}
inline size_t gso_prep_elems(int N) { return (size_t)N * ((N + 63) / 64 * 64); }

// mbf: bf16 [N][Kp] prepared operator (gso_prep_kernel)
inline void launch_gso_umma(const bf16* mbf, const bf16* in, const bf16* aux, bf16* out, int N, int C, long long G,
                            float alpha, float beta, cudaStream_t stream) {
  const bool ktiled = gso_ktiled_forced() || !plan_gso(N, C).ok;
  GsoPlan pl = ktiled ? plan_gso_ktiled(N, C) : plan_gso(N, C);
  STGCN_CHECK(pl.ok, STGCN_E_UNSUPPORTED, "umma gso: unsupported shape");
  uint64_t ad[2] = {(uint64_t)pl.Kp, (uint64_t)N};
  uint64_t as[1] = {(uint64_t)pl.Kp * 2};
  uint32_t ab[2] = {64, 128};
  CUtensorMap tmA = make_tmap_bf16(mbf, 2, ad, as, ab, CU_TENSOR_MAP_SWIZZLE_128B);
  uint64_t bd[3] = {(uint64_t)C, (uint64_t)N, (uint64_t)G};
  uint64_t bs[2] = {(uint64_t)C * 2, (uint64_t)N * C * 2};
  uint32_t bb[3] = {(uint32_t)C, 64, (uint32_t)pl.Gb};
  const CUtensorMapSwizzle sw = C == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (C == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  CUtensorMap tmB = make_tmap_bf16(in, 3, bd, bs, bb, sw);
  GsoParams p{};
  p.N = N; p.C = C; p.Gb = pl.Gb; p.nKB = pl.nKB; p.S = pl.S; p.G = G;
  p.n_sets = (int)((G + pl.Gb - 1) / pl.Gb);
  p.a_bytes = pl.a_bytes; p.stage_bytes = pl.stage_bytes;
  p.b_swz = C == 64 ? SWZ_128B : (C == 32 ? SWZ_64B : SWZ_32B);
  p.b_lbo = 64u * C * 2;          // next group's chunk (64 rows of C*2 bytes)
  p.b_sbo = 8u * C * 2;           // next 8-row group along K
  p.b_kadv = 16u * C * 2;         // 16 K rows per MMA
  p.alpha = alpha; p.beta = beta; p.aux = aux; p.out = out;
  int per = sm_count() / pl.nMT;
  int gx = p.n_sets < per ? p.n_sets : per;
  if (gx < 1) gx = 1;
  if (ktiled) {
    STGCN_CUDA(cudaFuncSetAttribute(umma_gso_ktiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
    STGCN_LAUNCH(umma_gso_ktiled_kernel, dim3(gx, pl.nMT), kTapThreads, pl.smem, stream, tmA, tmB, p);
    return;
  }
  STGCN_CUDA(cudaFuncSetAttribute(umma_gso_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
  STGCN_LAUNCH(umma_gso_kernel, dim3(gx, pl.nMT), kTapThreads, pl.smem, stream, tmA, tmB, p);
}

}  // namespace umma
}  // namespace stgcn
