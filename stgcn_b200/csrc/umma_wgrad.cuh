// umma_wgrad.cuh -- tcgen05 weight gradient of the temporal convolution (bf16 path):
//
//   dW_j[o, c] = sum_{b, t, n} dZ[(b, t, n), o] * X[(b, t + j, n), c]     j < Kt
//   db[o]      = sum_{b, t, n} dZ[(b, t, n), o]
//
// The contraction runs over rows (b,t,n), so both operands are MN-major in their natural channels-last
// storage: A = dZ tile [64 rows][128 o] (two 64-wide swizzle-128B chunks), B = X tile [64 rows][Cin].
// Work item = (sample b, 64-vertex chunk); for each output step t the CTA issues, per tap j,
// 4 x tcgen05.mma (M = 128 output channels, N = Cin, K = 16 rows) into the tap's own TMEM accumulator
// D_j, plus one N=16 MMA against a constant ones tile for the bias gradient.  The X time slices slide
// through a TMA ring exactly as in umma_tap.cuh (each slice is used by Kt taps), dZ slices through a second
// ring.  Accumulators live in TMEM for the CTA's whole life (split-K over work items); at the end the
// epilogue warps add them to the fp32 gradient buffer with coalesced atomics.
// Output layout = the SIMT wgrad kernel's: dwt[(j*Cin + c)*W + o], bias row at j = Kt.
#pragma once
#include "umma_tap.cuh"

namespace stgcn {
namespace umma {

struct WgradParams {
  int B, N, T_in, T_out, Kt, Cin, W;
  int Sx, Sz, n_items, n_chunks;
  // time split (as in umma_tap.cuh): item = (sample, vertex chunk, output steps [ts*t_chunk, +t_chunk)); the partial sums
  // all land in the same TMEM accumulators, so a cut only costs the Kt-1 re-loaded X slices
  int n_tsplit, t_chunk;
  uint32_t x_bytes, z_bytes, b_swz, b_sbo, b_kadv;
  float* dwt;
  int want_bias;
  // A operand (dZ) geometry: `a_real` chunks of `a_cw` channels are loaded by TMA; when they cover fewer than the 128
  // M rows of the MMA the remaining chunk slots of the stage stay zero (narrow dZ, e.g. 16 channels)
  int a_cw, a_real;
  uint32_t a_swz, a_lbo, a_sbo, a_kadv, a_chunk_bytes;
  int KR;                       // vertices (K rows) per tile: 128 (64 made the ~1.6 us per-step latency dominate)
  // plane mode: tap j reads batch coordinate b + j*plane_b at the SAME time step (stacked operands) instead of time t+j
  int plane_mode, plane_b;
};

__global__ void __launch_bounds__(kTapThreads, 1)
umma_wgrad_kernel(const __grid_constant__ CUtensorMap tmZ, const __grid_constant__ CUtensorMap tmX, WgradParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* ones = smem;                               // [64 rows][16] bf16 1.0 (2 KB, padded to 1 KB multiple)
  uint8_t* zring = smem + 2048;                       // Sz x [2 chunks][KR rows][128 B]
  uint8_t* xring = zring + (size_t)p.Sz * p.z_bytes;  // Sx x [KR rows][Cin*2 B]
  __shared__ __align__(8) uint64_t xfull[kMaxStages], xempty[kMaxStages], zfull[kMaxStages], zempty[kMaxStages], done;
  __shared__ uint32_t tmem_base_s;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int o0 = blockIdx.y * 128;
  const int ncol_used = p.Kt * p.Cin + 16;
  uint32_t ncols = 32;
  while ((int)ncols < ncol_used) ncols <<= 1;

  for (int i = threadIdx.x; i < 1024; i += blockDim.x) reinterpret_cast<__nv_bfloat16*>(ones)[i] = __float2bfloat16_rn(1.f);
  if (p.a_real * p.a_cw < 128) {         // narrow dZ: the unloaded chunk slots must read as zeros
    uint4* zr = reinterpret_cast<uint4*>(smem + 2048);
    const int n16 = (int)((size_t)p.Sz * p.z_bytes / 16);
    for (int i = threadIdx.x; i < n16; i += blockDim.x) zr[i] = make_uint4(0, 0, 0, 0);
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < p.Sx; ++s) { mbar_init(&xfull[s], 1); mbar_init(&xempty[s], 1); }
    for (int s = 0; s < p.Sz; ++s) { mbar_init(&zfull[s], 1); mbar_init(&zempty[s], 1); }
    mbar_init(&done, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&tmem_base_s, ncols);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  const int nxc = p.Cin > 64 ? p.Cin / 64 : 1;        // 64-wide chunks of the X tile
  const int xcw = p.Cin > 64 ? 64 : p.Cin;

  if (warp == 0) {
    if (lane == 0) {
      tma_prefetch_desc(&tmZ);
      tma_prefetch_desc(&tmX);
      RingPos rx{0, 0}, rz{0, 0};                   // ring positions advance incrementally: no integer division (umma.cuh)
      const uint32_t z_tx = (uint32_t)p.a_real * p.a_chunk_bytes;
      auto load_x = [&](int ti, int bi, int n0) {
        const uint32_t s = rx.s, ph = rx.ph;
        mbar_wait(&xempty[s], ph ^ 1);
        mbar_arrive_expect_tx(&xfull[s], p.x_bytes);
        uint8_t* dst = xring + (size_t)s * p.x_bytes;
        for (int c = 0; c < nxc; ++c) tma_load_4d(dst + (size_t)c * p.KR * xcw * 2, &tmX, &xfull[s], c * xcw, n0, ti, bi);
        rx.advance(p.Sx);
      };
      auto load_z = [&](int t_o, int b, int n0) {
        const uint32_t s = rz.s, ph = rz.ph;
        mbar_wait(&zempty[s], ph ^ 1);
        mbar_arrive_expect_tx(&zfull[s], z_tx);
        uint8_t* dst = zring + (size_t)s * p.z_bytes;
        for (int c = 0; c < p.a_real; ++c) tma_load_4d(dst + (size_t)c * p.a_chunk_bytes, &tmZ, &zfull[s], o0 + c * p.a_cw, n0, t_o, b);
        rz.advance(p.Sz);
      };
      for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
        const int ts = item % p.n_tsplit, rest = item / p.n_tsplit;
        const int b = rest / p.n_chunks, n0 = (rest % p.n_chunks) * p.KR;
        const int t_begin = ts * p.t_chunk, t_end = t_begin + p.t_chunk < p.T_out ? t_begin + p.t_chunk : p.T_out;
        if (p.plane_mode) {
          for (int t_o = t_begin; t_o < t_end; ++t_o) {
            for (int j = 0; j < p.Kt; ++j) load_x(t_o, b + j * p.plane_b, n0);
            load_z(t_o, b, n0);
          }
        } else {
          for (int ti = t_begin; ti < t_end + p.Kt - 1; ++ti) {
            load_x(ti, b, n0);
            const int t_o = ti - (p.Kt - 1);
            if (t_o >= t_begin && t_o < t_end) load_z(t_o, b, n0);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {      // one elected lane, known to the compiler as such (issue cost: see umma.cuh)
      const uint32_t idesc = make_idesc_bf16(128, p.Cin, 1, 1);
      const uint32_t idesc_b = make_idesc_bf16(128, 16, 1, 1);
      const uint32_t ones_a = smem_u32(ones);
      const uint64_t pa = make_smem_desc(0, p.a_lbo, p.a_sbo, p.a_swz);
      const uint64_t pb = make_smem_desc(0, (uint32_t)p.KR * xcw * 2, p.b_sbo, p.b_swz);
      const uint64_t pones = make_smem_desc(0, 2048, 256, SWZ_32B);
      const uint64_t a_step = p.a_kadv >> 4, b_step = p.b_kadv >> 4;
      uint32_t started = 0;
      RingPos rz{0, 0}, xwin{0, 0};                 // xwin: ring position of the first X slice of the current output step
      const uint32_t zring_s = smem_u32(zring), xring_s = smem_u32(xring);
      const int nk = p.KR / 16;
      for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
        const int ts = item % p.n_tsplit;
        const int t_begin = ts * p.t_chunk, t_end = t_begin + p.t_chunk < p.T_out ? t_begin + p.t_chunk : p.T_out;
        const int t_len = t_end - t_begin;
        int n_waited = 0;                           // X slices of this item already waited for (the window slides by one)
        for (int t_r = 0; t_r < t_len; ++t_r, rz.advance(p.Sz)) {     // t_r: output step relative to the item's first
          mbar_wait(&zfull[rz.s], rz.ph);
          tc_fence_after();
          const uint32_t a_base = zring_s + rz.s * p.z_bytes;
          RingPos px = xwin;
          for (int j = 0; j < p.Kt; ++j, px.advance(p.Sx)) {
            const int d = p.plane_mode ? t_r * p.Kt + j : t_r + j;
            if (d >= n_waited) {
              mbar_wait(&xfull[px.s], px.ph);
              tc_fence_after();
              n_waited = d + 1;
            }
            const uint32_t b_base = xring_s + px.s * p.x_bytes;
            const uint32_t acc = (started >> j) & 1;
            uint64_t da = desc_at(pa, a_base), db = desc_at(pb, b_base);
            for (int k = 0; k < nk; ++k) {
              mma_bf16_ss(tmem_base + j * p.Cin, da, db, idesc, acc | (k != 0));
              da += a_step; db += b_step;
            }
            started |= 1u << j;
          }
          if (p.want_bias) {
            const uint32_t acc = (started >> 31) & 1;
            uint64_t da = desc_at(pa, a_base);
            for (int k = 0; k < nk; ++k) {
              mma_bf16_ss(tmem_base + p.Kt * p.Cin, da, desc_at(pones, ones_a + (k & 3) * 512), idesc_b, acc | (k != 0));
              da += a_step;
            }
            started |= 1u << 31;
          }
          mma_commit(&zempty[rz.s]);
          if (p.plane_mode) {
            for (int j = 0; j < p.Kt; ++j, xwin.advance(p.Sx)) mma_commit(&xempty[xwin.s]);
          } else {
            mma_commit(&xempty[xwin.s]);
            xwin.advance(p.Sx);
            if (t_r == t_len - 1)
              for (int j = 0; j < p.Kt - 1; ++j, xwin.advance(p.Sx)) mma_commit(&xempty[xwin.s]);
          }
        }
      }
      mma_commit(&done);
    }
  } else {
    // epilogue: wait for all MMAs of this CTA, then add the accumulators to the global gradient
    const int q = warp & 3;
    const int o = o0 + q * 32 + lane;
    const bool any = blockIdx.x < p.n_items;
    if (any) {
      mbar_wait(&done, 0);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16);
      for (int j = 0; j < p.Kt; ++j)
        for (int c0 = 0; c0 < p.Cin; c0 += 16) {
          uint32_t r[16];
          tmem_ld_32x32b_x16(t_addr + j * p.Cin + c0, r);
          tmem_ld_wait();
          if (o < p.W) {
#pragma unroll
            for (int i = 0; i < 16; ++i)
              atomicAdd(p.dwt + ((size_t)(j * p.Cin + c0 + i)) * p.W + o, __uint_as_float(r[i]));
          }
        }
      if (p.want_bias) {
        uint32_t r[16];
        tmem_ld_32x32b_x16(t_addr + p.Kt * p.Cin, r);
        tmem_ld_wait();
        if (o < p.W) atomicAdd(p.dwt + (size_t)p.Kt * p.Cin * p.W + o, __uint_as_float(r[0]));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, ncols);
}

// ------------------------------------------------------------------------------------------------
// Flat-row variant for weight gradients WITHOUT a time shift (1x1 align conv, Chebyshev / GCN weight contraction):
//   dW_j[o, c] = sum_r dZ[r, o] * X_j[r, c],   r over all B*T*N rows, X_j = plane j of a stack (or the input itself).
// The rows are one flat axis, so a pipeline stage is 256 consecutive rows of dZ and of every plane (2-D TMA boxes):
// 4x fewer, 4x larger steps than the (sample, 64-vertex) items above, whose ~1.6 us per-step latency dominated.
// dZ is narrow here (W = 16/32/64 channels): the MMA's M = 128 rows are filled by aliasing the single W-wide chunk
// (descriptor leading-dim offset 0); accumulator lanes >= W hold duplicates and are never read.
// ------------------------------------------------------------------------------------------------
struct WgradFlatParams {
  long long rows, plane_rows;
  int Kt, Cin, W, S, n_tiles;
  uint32_t z_bytes, x_bytes, stage_bytes, a_swz, a_sbo, a_kadv, b_swz, b_sbo, b_kadv;
  float* dwt;
  int want_bias;
  // merge_taps: the Kt plane tiles of a stage sit x_bytes apart, which is exactly the leading-dimension stride of the
  // MN-major B descriptor, so ONE instruction with N = Kt*Cin covers all planes (accumulator columns [j*Cin, +Cin) are
  // contiguous too).  With Cin = 16 the per-tap N = 16 instructions cost the same ~45 cycles each and made the kernel
  // issue bound (64 instructions per 32 KB tile).
  int merge_taps;
};
constexpr int kFlatRows = 256;

__global__ void __launch_bounds__(kTapThreads, 1)
umma_wgrad_flat_kernel(const __grid_constant__ CUtensorMap tmZ, const __grid_constant__ CUtensorMap tmX, WgradFlatParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* ones = smem;                       // [64 rows][16] bf16 1.0
  uint8_t* ring = smem + 2048;                // S x { dZ [256][W*2 B] , Kt x X [256][Cin*2 B] }
  __shared__ __align__(8) uint64_t full[kMaxStages], empty[kMaxStages], done;
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t ncols = 32;
  while ((int)ncols < p.Kt * p.Cin + 16) ncols <<= 1;
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) reinterpret_cast<__nv_bfloat16*>(ones)[i] = __float2bfloat16_rn(1.f);
  if (threadIdx.x == 0) {
    for (int s = 0; s < p.S; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(&done, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&tmem_base_s, ncols);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  const int nxc = p.Cin > 64 ? p.Cin / 64 : 1, xcw = p.Cin > 64 ? 64 : p.Cin;

  if (warp == 0) {
    if (lane == 0) {
      tma_prefetch_desc(&tmZ);
      tma_prefetch_desc(&tmX);
      RingPos rp{0, 0};
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, rp.advance(p.S)) {
        const uint32_t s = rp.s, ph = rp.ph;
        const int r0 = tile * kFlatRows;
        mbar_wait(&empty[s], ph ^ 1);
        mbar_arrive_expect_tx(&full[s], p.stage_bytes);
        uint8_t* dst = ring + (size_t)s * p.stage_bytes;
        tma_load_2d(dst, &tmZ, &full[s], 0, r0);
        for (int j = 0; j < p.Kt; ++j)
          for (int c = 0; c < nxc; ++c)
            tma_load_2d(dst + p.z_bytes + (size_t)j * p.x_bytes + (size_t)c * kFlatRows * xcw * 2, &tmX, &full[s], c * xcw,
                        (int)(r0 + j * p.plane_rows));
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {      // one elected lane, known to the compiler as such (issue cost: see umma.cuh)
      const uint32_t idesc = make_idesc_bf16(128, p.Cin, 1, 1), idesc_b = make_idesc_bf16(128, 16, 1, 1);
      const uint32_t ones_a = smem_u32(ones);
      const uint64_t pa = make_smem_desc(0, 0, p.a_sbo, p.a_swz);                              // LBO 0: chunk aliased
      const uint64_t pb = make_smem_desc(0, (uint32_t)kFlatRows * xcw * 2, p.b_sbo, p.b_swz);
      const uint64_t pones = make_smem_desc(0, 2048, 256, SWZ_32B);
      const uint64_t a_step = p.a_kadv >> 4, b_step = p.b_kadv >> 4;
      uint32_t g = 0;
      RingPos rp{0, 0};
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++g, rp.advance(p.S)) {
        const uint32_t s = rp.s, ph = rp.ph;
        mbar_wait(&full[s], ph);
        tc_fence_after();
        const uint32_t a_base = smem_u32(ring + (size_t)s * p.stage_bytes);
        if (p.merge_taps) {
          const uint32_t idesc_m = make_idesc_bf16(128, p.Kt * p.Cin, 1, 1);
          uint64_t da = desc_at(pa, a_base), db = desc_at(pb, a_base + p.z_bytes);
#pragma unroll 4
          for (int k = 0; k < kFlatRows / 16; ++k) {
            mma_bf16_ss(tmem_base, da, db, idesc_m, (g | (uint32_t)k) != 0);
            da += a_step; db += b_step;
          }
        }
        for (int j = 0; j < (p.merge_taps ? 0 : p.Kt); ++j) {
          const uint32_t b_base = a_base + p.z_bytes + j * p.x_bytes;
          uint64_t da = desc_at(pa, a_base), db = desc_at(pb, b_base);
#pragma unroll 4
          for (int k = 0; k < kFlatRows / 16; ++k) {
            mma_bf16_ss(tmem_base + j * p.Cin, da, db, idesc, (g | (uint32_t)k) != 0);
            da += a_step; db += b_step;
          }
        }
        if (p.want_bias) {
          uint64_t da = desc_at(pa, a_base);
#pragma unroll 4
          for (int k = 0; k < kFlatRows / 16; ++k) {
            mma_bf16_ss(tmem_base + p.Kt * p.Cin, da, desc_at(pones, ones_a + (k & 3) * 512), idesc_b, (g | (uint32_t)k) != 0);
            da += a_step;
          }
        }
        mma_commit(&empty[s]);
      }
      mma_commit(&done);
    }
  } else {
    const int q = warp & 3;
    const int o = q * 32 + lane;
    if (blockIdx.x < p.n_tiles) {
      mbar_wait(&done, 0);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
      for (int c0 = 0; c0 < p.Kt * p.Cin; c0 += 8) {
        uint32_t r[8];
        tmem_ld_32x32b_x8(t_addr + c0, r);
        tmem_ld_wait();
        if (o < p.W) {
#pragma unroll
          for (int i = 0; i < 8; ++i) atomicAdd(p.dwt + (size_t)(c0 + i) * p.W + o, __uint_as_float(r[i]));
        }
      }
      if (p.want_bias) {
        uint32_t r[8];
        tmem_ld_32x32b_x8(t_addr + p.Kt * p.Cin, r);
        tmem_ld_wait();
        if (o < p.W) atomicAdd(p.dwt + (size_t)p.Kt * p.Cin * p.W + o, __uint_as_float(r[0]));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, ncols);
}

struct WgradFlatPlan { bool ok; int S; uint32_t z_bytes, x_bytes, stage_bytes; size_t smem; };
inline WgradFlatPlan plan_wgrad_flat(int Cin, int W, int Kt) {
  WgradFlatPlan pl{};
  pl.ok = false;
  if (W != 16 && W != 32 && W != 64) return pl;
  if (!(Cin == 16 || Cin == 32 || Cin == 64 || Cin == 128)) return pl;
  if (Kt < 1 || Kt * Cin + 16 > 512) return pl;
  pl.z_bytes = (uint32_t)kFlatRows * W * 2;
  pl.x_bytes = (uint32_t)kFlatRows * Cin * 2;
  pl.stage_bytes = pl.z_bytes + (uint32_t)Kt * pl.x_bytes;
  int S = (int)((kSmemBudget - 2048) / pl.stage_bytes);
  if (S > 6) S = 6;
  if (S < 2) return pl;
  pl.S = S;
  pl.smem = 2048 + (size_t)S * pl.stage_bytes + 1024;
  pl.ok = true;
  return pl;
}
inline bool wgrad_flat_supported(int Cin, int W, int Kt, long long rows) {
  return rows > 0 && rows * (long long)(Kt > 1 ? Kt : 1) < (1LL << 31) && plan_wgrad_flat(Cin, W, Kt).ok;
}
// x: Kt planes of [rows, Cin] (plane stride = plane_rows rows); dz: [rows, W]; dwt: fp32 [(Kt*Cin + 1), W], pre-zeroed
inline void launch_wgrad_flat(const bf16* x, const bf16* dz, float* dwt, long long rows, long long plane_rows, int Kt,
                              int Cin, int W, int want_bias, cudaStream_t stream) {
  WgradFlatPlan pl = plan_wgrad_flat(Cin, W, Kt);
  STGCN_CHECK(pl.ok, STGCN_E_UNSUPPORTED, "umma flat wgrad: unsupported shape");
  auto swz_of = [](int cw) { return cw == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (cw == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B); };
  auto dsw_of = [](int cw) { return cw == 64 ? SWZ_128B : (cw == 32 ? SWZ_64B : SWZ_32B); };
  uint64_t zd[2] = {(uint64_t)W, (uint64_t)rows};
  uint64_t zs[1] = {(uint64_t)W * 2};
  uint32_t zb[2] = {(uint32_t)W, (uint32_t)kFlatRows};
  CUtensorMap tmZ = make_tmap_bf16(dz, 2, zd, zs, zb, swz_of(W));
  const int xcw = Cin > 64 ? 64 : Cin;
  const long long xrows = plane_rows * (Kt - 1) + rows;
  uint64_t xd[2] = {(uint64_t)Cin, (uint64_t)xrows};
  uint64_t xs[1] = {(uint64_t)Cin * 2};
  uint32_t xb[2] = {(uint32_t)xcw, (uint32_t)kFlatRows};
  CUtensorMap tmX = make_tmap_bf16(x, 2, xd, xs, xb, swz_of(xcw));
  WgradFlatParams p{};
  p.rows = rows; p.plane_rows = plane_rows; p.Kt = Kt; p.Cin = Cin; p.W = W; p.S = pl.S;
  p.n_tiles = (int)((rows + kFlatRows - 1) / kFlatRows);
  p.z_bytes = pl.z_bytes; p.x_bytes = pl.x_bytes; p.stage_bytes = pl.stage_bytes;
  p.a_swz = dsw_of(W); p.a_sbo = 8u * W * 2; p.a_kadv = 16u * W * 2;
  p.b_swz = dsw_of(xcw); p.b_sbo = 8u * xcw * 2; p.b_kadv = 16u * xcw * 2;
  p.dwt = dwt; p.want_bias = want_bias;
  p.merge_taps = (Kt > 1 && Cin <= 64 && Kt * Cin <= 256 && (Kt * Cin) % 16 == 0) ? 1 : 0;
  int gx = p.n_tiles < sm_count() ? p.n_tiles : sm_count();
  STGCN_CUDA(cudaFuncSetAttribute(umma_wgrad_flat_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
  STGCN_LAUNCH(umma_wgrad_flat_kernel, gx, kTapThreads, pl.smem, stream, tmZ, tmX, p);
}

constexpr int kWgradKR = 128;
struct WgradPlan { bool ok; int Sx, Sz, nMT, a_cw, a_real; uint32_t x_bytes, z_bytes; size_t smem; };

// W: channels of dz (128-multiples: full tiles; 16/32/64: one narrow, zero-padded tile)
inline WgradPlan plan_wgrad(int Cin, int W, int Kt, int T_in, bool plane_mode) {
  WgradPlan pl{};
  pl.ok = false;
  if (W >= 128) {
    if (W % 128) return pl;
    pl.a_cw = 64; pl.a_real = 2; pl.nMT = W / 128;
  } else {
    if (W != 16 && W != 32 && W != 64) return pl;
    pl.a_cw = W; pl.a_real = 1; pl.nMT = 1;
  }
  if (!(Cin == 16 || Cin == 32 || Cin == 64 || Cin == 128)) return pl;
  if (Kt * Cin + 16 > 512 || Kt > 30) return pl;
  pl.x_bytes = (uint32_t)kWgradKR * Cin * 2;
  pl.z_bytes = (uint32_t)kWgradKR * 128 * 2;       // 128 channel slots (real chunks + zero padding for narrow dZ)
  int live = plane_mode ? Kt : (Kt < T_in ? Kt : T_in);
  // deepest rings that fit: X needs the `live` window + >= 1 slot in flight, dZ >= 2 slots
  for (int extra = 3; extra >= 1 && !pl.ok; --extra)
    for (int sz = 3; sz >= 2 && !pl.ok; --sz) {
      pl.Sx = live + extra > kMaxStages ? kMaxStages : live + extra;
      if (pl.Sx < live + 1) continue;
      pl.Sz = sz;
      pl.smem = 2048 + (size_t)pl.Sz * pl.z_bytes + (size_t)pl.Sx * pl.x_bytes + 1024;
      pl.ok = pl.smem <= kSmemBudget;
    }
  return pl;
}
inline bool wgrad_supported(int Cin, int W, int Kt, int T_in, int B, bool plane_mode = false) {
  return B > 0 && plan_wgrad(Cin, W, Kt, T_in, plane_mode).ok;
}

// Time mode : x [B, T_in, N, Cin], dz [B, T_out, N, W], T_out = T_in - Kt + 1; tap j pairs dz(t) with x(t + j).
// Plane mode: x [Kt*B, T, N, Cin] (Kt stacked planes), dz [B, T, N, W]; tap j pairs dz(b, t) with x(b + j*B, t).
// dwt: fp32 [(Kt*Cin + 1), W], pre-zeroed.
inline void launch_wgrad_umma(const bf16* x, const bf16* dz, float* dwt, int B, int N, int T_in, int Kt, int Cin, int W,
                              int want_bias, cudaStream_t stream, bool plane_mode = false) {
  WgradPlan pl = plan_wgrad(Cin, W, Kt, T_in, plane_mode);
  STGCN_CHECK(pl.ok, STGCN_E_UNSUPPORTED, "umma wgrad: unsupported shape");
  const int T_out = plane_mode ? T_in : T_in - Kt + 1;
  uint64_t zd[4] = {(uint64_t)W, (uint64_t)N, (uint64_t)T_out, (uint64_t)B};
  uint64_t zs[3] = {(uint64_t)W * 2, (uint64_t)N * W * 2, (uint64_t)T_out * N * W * 2};
  uint32_t zb[4] = {(uint32_t)pl.a_cw, (uint32_t)kWgradKR, 1, 1};
  const CUtensorMapSwizzle zsw = pl.a_cw == 64 ? CU_TENSOR_MAP_SWIZZLE_128B
                               : (pl.a_cw == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  CUtensorMap tmZ = make_tmap_bf16(dz, 4, zd, zs, zb, zsw);
  const int xcw = Cin > 64 ? 64 : Cin;
  const int xB = plane_mode ? Kt * B : B;
  uint64_t xd[4] = {(uint64_t)Cin, (uint64_t)N, (uint64_t)T_in, (uint64_t)xB};
  uint64_t xs[3] = {(uint64_t)Cin * 2, (uint64_t)N * Cin * 2, (uint64_t)T_in * N * Cin * 2};
  uint32_t xb[4] = {(uint32_t)xcw, (uint32_t)kWgradKR, 1, 1};
  const CUtensorMapSwizzle sw = xcw == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (xcw == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  CUtensorMap tmX = make_tmap_bf16(x, 4, xd, xs, xb, sw);
  WgradParams p{};
  p.B = B; p.N = N; p.T_in = T_in; p.T_out = T_out; p.Kt = Kt; p.Cin = Cin; p.W = W;
  p.KR = kWgradKR;
  p.Sx = pl.Sx; p.Sz = pl.Sz; p.n_chunks = (N + kWgradKR - 1) / kWgradKR;
  {   // time split: minimise rounds x bytes loaded per item (dZ slices + X slices incl. the Kt-1 re-loaded ones)
    const long long base_items = (long long)B * p.n_chunks;
    const int ctas = sm_count() / pl.nMT > 0 ? sm_count() / pl.nMT : 1;
    long long best = -1;
    p.n_tsplit = 1; p.t_chunk = T_out;
    for (int ns = 1; ns <= 4 && ns <= T_out; ++ns) {
      const int chunk = (T_out + ns - 1) / ns, ns_eff = (T_out + chunk - 1) / chunk;
      if (ns_eff != ns) continue;
      const long long items = base_items * ns, g = items < ctas ? items : ctas;
      const long long rounds = (items + g - 1) / g;
      const long long xs_n = plane_mode ? (long long)chunk * Kt : chunk + Kt - 1;
      const long long cost = rounds * (chunk * (long long)(W < 128 ? W : 128) + xs_n * Cin);
      if (best < 0 || cost < best) { best = cost; p.n_tsplit = ns; p.t_chunk = chunk; }
    }
  }
  p.n_items = B * p.n_chunks * p.n_tsplit;
  p.x_bytes = pl.x_bytes; p.z_bytes = pl.z_bytes;
  p.b_swz = xcw == 64 ? SWZ_128B : (xcw == 32 ? SWZ_64B : SWZ_32B);
  p.b_sbo = 8u * xcw * 2; p.b_kadv = 16u * xcw * 2;
  p.a_cw = pl.a_cw; p.a_real = pl.a_real;
  p.a_swz = pl.a_cw == 64 ? SWZ_128B : (pl.a_cw == 32 ? SWZ_64B : SWZ_32B);
  p.a_chunk_bytes = (uint32_t)kWgradKR * pl.a_cw * 2;          // one chunk: KR K-rows x a_cw channels
  p.a_lbo = p.a_chunk_bytes; p.a_sbo = 8u * pl.a_cw * 2; p.a_kadv = 16u * pl.a_cw * 2;
  p.plane_mode = plane_mode ? 1 : 0; p.plane_b = B;
  p.dwt = dwt; p.want_bias = want_bias;
  int per = sm_count() / pl.nMT;
  int gx = p.n_items < per ? p.n_items : per;
  if (gx < 1) gx = 1;
  STGCN_CUDA(cudaFuncSetAttribute(umma_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
  STGCN_LAUNCH(umma_wgrad_kernel, dim3(gx, pl.nMT), kTapThreads, pl.smem, stream, tmZ, tmX, p);
}

}  // namespace umma
}  // namespace stgcn
