// umma_selftest.cuh -- a minimal tcgen05 GEMM used by tests/ to pin the operand layouts the
// production kernels rely on (descriptor fields, swizzle modes, TMA boxes) against a plain matmul.
//   C[M,N] (fp32) = A * B^T-like product with bf16 operands, fp32 accumulation in TMEM.
// Modes (operand storage in global memory, all bf16):
//   0: A [M,K] K-major, B [N,K] K-major, 128B swizzle, K blocks of 64
//   1: A [M,K] K-major, B [N,K] K-major, 32B swizzle,  K blocks of 16
//   2: A [K,M] MN-major, B [K,N] MN-major, 128B swizzle (64-element chunks along M/N), K blocks of 64
//   3: A [M,K] K-major 128B swizzle; B [G][K][16] MN-major 32B swizzle (N = 16*G), K blocks of 64
#pragma once
#include "umma.cuh"

namespace stgcn {
namespace umma {

struct SelftestArgs {
  float* C;
  int M, N, K, mode;
  uint32_t lbo_a, sbo_a, lbo_b, sbo_b;   // descriptor byte offsets (swept by the test for the MN-major modes)
};

__global__ void __launch_bounds__(128) umma_selftest_kernel(const __grid_constant__ CUtensorMap tmA,
                                                            const __grid_constant__ CUtensorMap tmB, SelftestArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* As = smem;                 // <= 16 KB
  uint8_t* Bs = smem + 16384;         // <= 32 KB
  __shared__ __align__(8) uint64_t bar_tma, bar_mma;
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.x * 128;
  uint32_t ncols = 32;
  while ((int)ncols < a.N) ncols <<= 1;

  if (warp == 0) tmem_alloc(&tmem_base_s, ncols);
  if (tid == 0) {
    mbar_init(&bar_tma, 1);
    mbar_init(&bar_mma, 1);
    fence_barrier_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (tid == 0) {
    const int KB = (a.mode == 1) ? 16 : 64;
    const int nkb = a.K / KB;
    const uint32_t idesc = make_idesc_bf16(128, a.N, a.mode == 2, a.mode == 2 || a.mode == 3);
    uint32_t phase = 0;
    for (int kb = 0; kb < nkb; ++kb) {
      uint32_t bytes;
      if (a.mode == 0) bytes = 128 * 128 + a.N * 128;
      else if (a.mode == 1) bytes = 128 * 32 + a.N * 32;
      else if (a.mode == 2) bytes = 2 * 64 * 128 + (a.N / 64) * 64 * 128;
      else bytes = 128 * 128 + (a.N / 16) * 64 * 32;
      mbar_arrive_expect_tx(&bar_tma, bytes);
      if (a.mode == 0) {
        tma_load_2d(As, &tmA, &bar_tma, kb * 64, m0);
        tma_load_2d(Bs, &tmB, &bar_tma, kb * 64, 0);
      } else if (a.mode == 1) {
        tma_load_2d(As, &tmA, &bar_tma, kb * 16, m0);
        tma_load_2d(Bs, &tmB, &bar_tma, kb * 16, 0);
      } else if (a.mode == 2) {
        for (int c = 0; c < 2; ++c) tma_load_2d(As + c * 8192, &tmA, &bar_tma, m0 + c * 64, kb * 64);
        for (int c = 0; c < a.N / 64; ++c) tma_load_2d(Bs + c * 8192, &tmB, &bar_tma, c * 64, kb * 64);
      } else {
        tma_load_2d(As, &tmA, &bar_tma, kb * 64, m0);
        tma_load_3d(Bs, &tmB, &bar_tma, 0, kb * 64, 0);      // box [16, 64, G]
      }
      mbar_wait(&bar_tma, phase);
      tc_fence_after();
      const int nk16 = KB / 16;
      for (int k = 0; k < nk16; ++k) {
        uint64_t da, db;
        if (a.mode == 0) {
          da = make_smem_desc(smem_u32(As) + k * 32, 16, 1024, SWZ_128B);
          db = make_smem_desc(smem_u32(Bs) + k * 32, 16, 1024, SWZ_128B);
        } else if (a.mode == 1) {
          da = make_smem_desc(smem_u32(As), 16, 256, SWZ_32B);
          db = make_smem_desc(smem_u32(Bs), 16, 256, SWZ_32B);
        } else if (a.mode == 2) {
          da = make_smem_desc(smem_u32(As) + k * 2048, a.lbo_a, a.sbo_a, SWZ_128B);
          db = make_smem_desc(smem_u32(Bs) + k * 2048, a.lbo_b, a.sbo_b, SWZ_128B);
        } else {
          da = make_smem_desc(smem_u32(As) + k * 32, 16, 1024, SWZ_128B);
          db = make_smem_desc(smem_u32(Bs) + k * 512, a.lbo_b, a.sbo_b, SWZ_32B);
        }
        mma_bf16_ss(tmem_base, da, db, idesc, (kb | k) != 0);
      }
      mma_commit(&bar_mma);
      mbar_wait(&bar_mma, phase);
      phase ^= 1;
    }
  }
  __syncthreads();
  tc_fence_after();
  // epilogue: warp w owns TMEM lanes [32w, 32w+32)
  const int row = m0 + warp * 32 + lane;
  for (int c0 = 0; c0 < a.N; c0 += 16) {
    uint32_t r[16];
    tmem_ld_32x32b_x16(tmem_base + ((uint32_t)(warp * 32) << 16) + c0, r);
    tmem_ld_wait();
    if (row < a.M) {
#pragma unroll
      for (int j = 0; j < 16; ++j)
        if (c0 + j < a.N) a.C[(size_t)row * a.N + c0 + j] = __uint_as_float(r[j]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, ncols);
}

inline void run_selftest(int mode, const void* A, const void* B, float* C, int M, int N, int K, uint32_t lbo_a,
                         uint32_t sbo_a, uint32_t lbo_b, uint32_t sbo_b, cudaStream_t s) {
  STGCN_CHECK(M % 128 == 0 && N % 16 == 0 && N >= 16 && N <= 256, STGCN_E_INVALID, "selftest: bad M/N");
  STGCN_CHECK(mode >= 0 && mode <= 3, STGCN_E_INVALID, "selftest: bad mode");
  STGCN_CHECK(K % (mode == 1 ? 16 : 64) == 0, STGCN_E_INVALID, "selftest: bad K");
  if (mode == 2) STGCN_CHECK(N % 64 == 0, STGCN_E_INVALID, "selftest mode 2: N must be a multiple of 64");
  CUtensorMap tmA, tmB;
  if (mode == 0 || mode == 1 || mode == 3) {
    uint64_t dims[2] = {(uint64_t)K, (uint64_t)M};
    uint64_t str[1] = {(uint64_t)K * 2};
    uint32_t box[2] = {(uint32_t)(mode == 1 ? 16 : 64), 128};
    tmA = make_tmap_bf16(A, 2, dims, str, box, mode == 1 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_128B);
  } else {
    uint64_t dims[2] = {(uint64_t)M, (uint64_t)K};
    uint64_t str[1] = {(uint64_t)M * 2};
    uint32_t box[2] = {64, 64};
    tmA = make_tmap_bf16(A, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
  }
  if (mode == 0 || mode == 1) {
    uint64_t dims[2] = {(uint64_t)K, (uint64_t)N};
    uint64_t str[1] = {(uint64_t)K * 2};
    uint32_t box[2] = {(uint32_t)(mode == 1 ? 16 : 64), (uint32_t)N};
    tmB = make_tmap_bf16(B, 2, dims, str, box, mode == 1 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_128B);
  } else if (mode == 2) {
    uint64_t dims[2] = {(uint64_t)N, (uint64_t)K};
    uint64_t str[1] = {(uint64_t)N * 2};
    uint32_t box[2] = {64, 64};
    tmB = make_tmap_bf16(B, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
  } else {
    int G = N / 16;
    uint64_t dims[3] = {16, (uint64_t)K, (uint64_t)G};
    uint64_t str[2] = {32, (uint64_t)K * 32};
    uint32_t box[3] = {16, 64, (uint32_t)G};
    tmB = make_tmap_bf16(B, 3, dims, str, box, CU_TENSOR_MAP_SWIZZLE_32B);
  }
  SelftestArgs a{C, M, N, K, mode, lbo_a, sbo_a, lbo_b, sbo_b};
  size_t smem = 16384 + 32768 + 1024;
  STGCN_CUDA(cudaFuncSetAttribute(umma_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  STGCN_LAUNCH(umma_selftest_kernel, M / 128, 128, smem, s, tmA, tmB, a);
}

}  // namespace umma
}  // namespace stgcn
