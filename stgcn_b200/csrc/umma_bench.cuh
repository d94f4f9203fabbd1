// umma_bench.cuh -- tcgen05.mma issue / completion cost microbenchmark (diagnostics: stgcn_umma_microbench).
// One CTA issues `n_mma` instructions of one shape / operand-layout configuration from one warp (warp-collective,
// uniform operands), round-robin over `n_chains` independent accumulators, and reports clock64 deltas:
//   out[0] = cycles to ISSUE all instructions, out[1] = cycles until the commit barrier fires, out[2] = n_mma.
// Operand memory is zero-filled shared memory (values do not matter for timing); descriptors advance by `kadv` per
// instruction inside a 16-step window like the production K loops do.  The table this produces is the performance model
// behind the kernel structure choices in DESIGN.md.
#pragma once
#include "umma.cuh"

namespace stgcn {
namespace umma {

struct MmaBenchCfg {
  int M, N, a_mn, b_mn, a_tmem;            // shape, operand majors, A from tensor memory
  uint32_t a_swz, a_lbo, a_sbo, a_kadv;    // A descriptor (shared-memory A)
  uint32_t b_swz, b_lbo, b_sbo, b_kadv;    // B descriptor
  int n_mma, n_chains, chain_cols;         // accumulator i % n_chains at column (i % n_chains) * chain_cols
  int style, n_warps;                      // 0: warp-collective (elect), 1: `if (lane == 0)` branch, 2: elect, descriptors
};                                         //    advanced by adds only; n_warps issuing warps, each with its own accumulators

__global__ void __launch_bounds__(192, 1) umma_bench_kernel(MmaBenchCfg c, unsigned long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ __align__(8) uint64_t done;
  __shared__ uint32_t tmem_base_s;
  const int warp = warp_idx_uniform(), lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 160 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) { mbar_init(&done, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = uniform_u32(tmem_base_s);
  if (warp >= 1 && warp <= c.n_warps) {
    const uint32_t idesc = make_idesc_bf16(c.M, c.N, c.a_mn, c.b_mn);
    const uint32_t a_u = smem_u32(smem), b_u = smem_u32(smem + 96 * 1024);
    const uint32_t acc0 = tb + 128 + (uint32_t)(warp - 1) * c.n_chains * c.chain_cols;   // after the (optional) A region
    for (int rep = 0; rep < 2; ++rep) {                  // first pass warms up, second is reported
      asm volatile("bar.sync 1, %0;" ::"r"(32 * c.n_warps));
      const long long t0 = clock64();
      if (c.style == 0) {
        for (int i = 0; i < c.n_mma; ++i) {
          const int ks = i & 15;
          const uint32_t d = acc0 + (uint32_t)(i % c.n_chains) * c.chain_cols;
          const uint64_t db = make_smem_desc(b_u + ks * c.b_kadv, c.b_lbo, c.b_sbo, c.b_swz);
          if (c.a_tmem) {
            mma_bf16_ts_w(d, tb + ks * 8, db, idesc, i >= c.n_chains);
          } else {
            const uint64_t da = make_smem_desc(a_u + ks * c.a_kadv, c.a_lbo, c.a_sbo, c.a_swz);
            mma_bf16_ss_w(d, da, db, idesc, i >= c.n_chains);
          }
        }
      } else if (c.style == 1) {
        if (lane == 0) {
          for (int i = 0; i < c.n_mma; ++i) {
            const int ks = i & 15;
            const uint32_t d = acc0 + (uint32_t)(i % c.n_chains) * c.chain_cols;
            const uint64_t db = make_smem_desc(b_u + ks * c.b_kadv, c.b_lbo, c.b_sbo, c.b_swz);
            const uint64_t da = make_smem_desc(a_u + ks * c.a_kadv, c.a_lbo, c.a_sbo, c.a_swz);
            mma_bf16_ss(d, da, db, idesc, i >= c.n_chains);
          }
        }
        __syncwarp();
      } else {
        // one elect for the whole batch, descriptors advanced by 64-bit adds, single accumulator chain per 16 steps
        const uint64_t da0 = make_smem_desc(a_u, c.a_lbo, c.a_sbo, c.a_swz), db0 = make_smem_desc(b_u, c.b_lbo, c.b_sbo, c.b_swz);
        const uint64_t da_step = c.a_kadv >> 4, db_step = c.b_kadv >> 4;
        if (elect_one()) {
          for (int i0 = 0; i0 < c.n_mma; i0 += 16) {
            uint64_t da = da0, db = db0;
#pragma unroll
            for (int ks = 0; ks < 16; ++ks) {
              mma_bf16_ss(acc0, da, db, idesc, (i0 | ks) != 0);
              da += da_step; db += db_step;
            }
          }
        }
        __syncwarp();
      }
      const long long t1 = clock64();
      if (warp == 1) {
        asm volatile("bar.sync 2, %0;" ::"r"(32 * c.n_warps));    // every issuer has issued
        mma_commit_w(&done);
        mbar_wait(&done, rep & 1);
      } else {
        if (elect_one()) {
          // each issuing warp must commit its own instructions: arrive on a private scratch barrier is not needed for
          // timing -- the reported completion time is warp 1's; the other warps only add load
        }
        asm volatile("bar.sync 2, %0;" ::"r"(32 * c.n_warps));
      }
      const long long t2 = clock64();
      if (rep == 1 && lane == 0 && warp == 1) { out[0] = (unsigned long long)(t1 - t0); out[1] = (unsigned long long)(t2 - t0); out[2] = c.n_mma; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tb, 512);
}

inline void run_mma_bench(const MmaBenchCfg& c, unsigned long long* out_dev, cudaStream_t stream) {
  STGCN_CHECK(c.n_chains >= 1 && c.n_warps >= 1 && c.n_warps <= 4 && 128 + c.n_warps * c.n_chains * c.chain_cols <= 512 &&
                  c.chain_cols >= c.N, STGCN_E_INVALID,
              "mma bench: accumulators do not fit TMEM");
  const size_t smem = 161 * 1024;
  STGCN_CUDA(cudaFuncSetAttribute(umma_bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  STGCN_LAUNCH(umma_bench_kernel, 1, 192, smem, stream, c, out_dev);
}

}  // namespace umma
}  // namespace stgcn
