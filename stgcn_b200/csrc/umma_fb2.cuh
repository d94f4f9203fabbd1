// umma_fb2.cuh -- backward of an ST block's SECOND gated temporal convolution (layers.py:87-115 applied at
// layers.py:254) fused with the backward of the LayerNorm that follows it (layers.py:255), on tcgen05 (bf16 mode).
//
// The chain it replaces moved the 128-channel pre-activation gradient dZ through HBM three times:
//     ln_gate_bwd_kernel       dY, H3, Q -> dZ                  (wrote 120 MB at B = 256 in block 0, 93 us)
//     umma_tap_kernel<LINEAR>  dZ -> dH2 (data gradient)        (read it, 42 us)
//     umma_wgrad_kernel        dZ, H2 -> dW, db                 (read it again, 45 us on the helper stream)
// Here dZ exists only as a shared-memory tile.  Per (sample, 128-vertex tile, output step t) -- "tile" below --
//   E1  (16 epilogue warps; thread = one vertex row x 16 of the 64 channels)
//         dH3 = rstd * (dY * gamma - s1 - xhat * s2)            LayerNorm backward; s1, s2 = the group sums of
//                                                               ln_bwd_sums_kernel, xhat = (H3 - mean) * rstd
//         dP = dH3 * sigma(Q),  dQ = dH3 * H3 * (1 - sigma(Q))  GLU backward from the q-only saved state
//         -> bf16 tile dZ_t [128 rows][128 channels] in shared memory (128B swizzle: K-major A operand of the data-gradient
//            MMAs and, the same bytes, MN-major A operand of the weight-gradient MMAs);
//         (the LayerNorm parameter gradients come from ln_bwd_sums_pg_kernel, the pass that forms s1, s2)
//   MMA data gradient:  X_{t+j}[128 rows x 16] += dZ_t . W_j^T, j = 0..2   (K = 128 channels; a ring of per-input-step
//         accumulators in tensor memory, X_tau complete after tile tau) + one identity instruction for the zero-padded
//         residual into X_{t+2}.  (A scatter form -- one N = 48 accumulator per tile, summed over three tiles by the
//         epilogue -- needs a third of the instructions but three readers per accumulator; its epilogue step took 1700
//         cycles, profiles/r02_ab_batch_h.md.)
//   MMA weight gradient:  G_j[128 channels x 16] += dZ_t^T . H2_{t+j},  G_b += dZ_t^T . 1    (K = 128 rows; accumulators
//         persistent in tensor memory over all tiles of the CTA, flushed once with fp32 atomics)
//   E2  dH2_tau = X_tau -> bf16, 32 bytes per row to HBM.
// HBM traffic per tile: dY, H3, Q rows in (3 x 16 KB), H2 slice in (4 KB), dH2 slice out (4 KB).
// Serves the default architecture's blocks: 16 -> 64 GLU channels, Kt = 3, q-only saved state, no dropout in training.
#pragma once
#include "umma_fb0.cuh"

namespace stgcn {
namespace umma {

constexpr int kFb2EpiWarps = 16;
constexpr int kFb2Threads = 64 + 32 * kFb2EpiWarps + 128; // warp 0 H2 producer, warp 1 MMA issuer, 16 E1 warps, 4 E2 warps
constexpr int kFb2HStages = 8;                           // H2 slices in flight (4 KB each)
constexpr int kFb2NZ = 3;                                // dZ tiles in shared memory
constexpr int kFb2NX = 6;                                // data-gradient accumulators X_tau (16 columns each) in tensor memory
constexpr int kFb2Kt = 3, kFb2Ci = 16, kFb2Co = 64, kFb2W = 128;

struct Fb2Params {
  const bf16* dy;          // [B, T2, N, 64] gradient w.r.t. the LayerNorm output
  const bf16* h3;          // [B, T2, N, 64] LayerNorm input = output of the gated conv (saved)
  const bf16* q;           // [B, T2, N, 64] gate half of the conv's pre-activation (saved, q-only state)
  const bf16* h2;          // [B, T1, N, 16] input of the conv
  const float* mean; const float* rstd;   // [B * T2]
  const float* sums;       // [n_parts][B * T2][2]: per column part, (sum dY gamma, sum dY gamma xhat) / M of the group
  int n_parts; long long part_stride;      // (ln_bwd_sums_pg_kernel; the parts are added here, in a fixed order)
  const float* gamma;      // [N, 64] LayerNorm weight
  const float* conv_w;     // [128][16][3] the conv weight in the reference layout (o, c, k)
  bf16* dh2;               // [B, T1, N, 16] out
  float* dwt;              // [(3 * 16 + 1)][128] pre-zeroed: dwt[(j * 16 + c) * 128 + o], then the bias row
  int B, T2, T1, N, nnt;   // nnt = number of 128-vertex tiles
  unsigned long long* dbg; // optional SM-cycle stamps of CTA 0 (diagnostics, -DSTGCN_TIMELINE)
};

// shared-memory map (offsets from the 1024-aligned base)
constexpr uint32_t kFb2HRing = 0;                                   // kFb2HStages x 4096
constexpr uint32_t kFb2Wd = kFb2HRing + kFb2HStages * 4096;         // 2 x [48 rows = (j, c)][64 ch] = 2 x 6144
constexpr uint32_t kFb2Wres = kFb2Wd + 2 * 6144;                    // [16 rows][16] identity for the residual: 512 -> 2048
constexpr uint32_t kFb2Ones = kFb2Wres + 2048;                      // [16 rows][16] ones: 512 -> 1024
constexpr uint32_t kFb2Dz = kFb2Ones + 1024;                        // kFb2NZ x 32768 (1024-aligned)
constexpr uint32_t kFb2Pf = kFb2Dz + kFb2NZ * 32768;                // E1 operand slots: [6 chunks][512 threads][16 B] = 49152
constexpr uint32_t kFb2Smem = kFb2Pf + 6 * 512 * 16 + 1024;
static_assert(kFb2Dz % 1024 == 0 && kFb2Wd % 1024 == 0, "swizzled operands need 1024-byte alignment");
constexpr uint32_t kFb2D2Col = kFb2NX * 16;                         // 96: weight-gradient accumulators (3 taps x 16 + bias 16)

__global__ void __launch_bounds__(kFb2Threads, 1) umma_fb2_kernel(Fb2Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ __align__(8) uint64_t h_full[kFb2HStages], h_empty[kFb2HStages], dz_full[kFb2NZ], dz_empty[kFb2NZ],
      x_full[kFb2NX], x_free[kFb2NX], done;
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool dbg_on = p.dbg != nullptr && blockIdx.x == 0 && lane == 0;
  (void)dbg_on;

  // ---- one-time staging of the constant operands
  // W for the data gradient: K-major B [n = (j, c)][k = o], two 64-channel blocks of [48 rows x 128 B], 128B swizzle
  for (int i = threadIdx.x; i < 48 * 16; i += blockDim.x) {
    const int n = i >> 4, ch = i & 15, j = n >> 4, c = n & 15, o0 = ch * 8;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = p.conv_w[(o0 + e) * (kFb2Ci * kFb2Kt) + c * kFb2Kt + j];
    *reinterpret_cast<uint4*>(smem + kFb2Wd + (ch >> 3) * 6144 + n * 128 + (((ch & 7) ^ (n & 7)) << 4)) = pack8_bf16(v);
  }
  // identity for the zero-padded residual (h3 = (P + pad(x_{t+2})) sigma(Q): dH2[t+2][c] += dP_t[c], c < 16):
  // K-major B [n = c][k = 16], 32-byte rows, 32B swizzle
  for (int i = threadIdx.x; i < 16 * 2; i += blockDim.x) {
    const int n = i >> 1, h = i & 1;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = (n == h * 8 + e) ? 1.f : 0.f;
    *reinterpret_cast<uint4*>(smem + kFb2Wres + n * 32 + ((h ^ ((n >> 2) & 1)) << 4)) = pack8_bf16(v);
  }
  for (int i = threadIdx.x; i < 32; i += blockDim.x)
    reinterpret_cast<uint4*>(smem + kFb2Ones)[i] = make_uint4(0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u);
  if (threadIdx.x == 0) {
    for (int s = 0; s < kFb2HStages; ++s) { mbar_init(&h_full[s], 1); mbar_init(&h_empty[s], 1); }
    for (int i = 0; i < kFb2NZ; ++i) { mbar_init(&dz_full[i], kFb2EpiWarps); mbar_init(&dz_empty[i], 1); }
    for (int i = 0; i < kFb2NX; ++i) { mbar_init(&x_full[i], 1); mbar_init(&x_free[i], 4); }          // one E2 group (4 warps) reads X_tau
    mbar_init(&done, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&tmem_base_s, 256);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  const uint32_t hfull_a = smem_u32(h_full), hempty_a = smem_u32(h_empty), zfull_a = smem_u32(dz_full), zempty_a = smem_u32(dz_empty),
                 xfull_a = smem_u32(x_full), xfree_a = smem_u32(x_free), done_a = smem_u32(&done);
  const uint32_t smem_s = smem_u32(smem);

  // items of this CTA: a fixed vertex tile, every bstep-th sample (so the LayerNorm parameter gradients of a thread's
  // (vertex, channel) positions accumulate in registers over the whole kernel)
  const int nt = (int)blockIdx.x % p.nnt, b0 = (int)blockIdx.x / p.nnt, bstep = (int)gridDim.x / p.nnt;
  const int n0 = nt * 128;
  const int T2 = p.T2, T1 = p.T1;

  if (warp == 0) {
    // =========================== producer: H2 slices by cp.async ================================
    RingPos rp{0, 0};
    int pending = -1;
    for (int b = b0; b < p.B; b += bstep) {
      for (int tau = 0; tau < T1; ++tau, rp.advance(kFb2HStages)) {
        mbar_wait_a(hempty_a + rp.s * 8, rp.ph ^ 1);
        uint8_t* dst = smem + kFb2HRing + rp.s * 4096;
        const bf16* src0 = p.h2 + ((long long)b * T1 + tau) * p.N * kFb2Ci;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const int qd = lane + 32 * c, row = qd >> 1, h = qd & 1;
          const bool ok = n0 + row < p.N;
          const bf16* src = src0 + (long long)(ok ? n0 + row : 0) * kFb2Ci + h * 8;
          cp_async16(dst + row * 32 + ((h ^ ((row >> 2) & 1)) << 4), src, ok ? 16u : 0u);
        }
        cp_async_commit();
        if (pending >= 0) {                                 // the previous slice has landed after this wait
          cp_async_wait<1>();
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) mbar_arrive_a(hfull_a + pending * 8);
        }
        pending = (int)rp.s;
      }
    }
    if (pending >= 0) {
      cp_async_wait<0>();
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive_a(hfull_a + pending * 8);
    }
  } else if (warp == 1) {
    // =========================== MMA issuer =============================
    const uint32_t u_smem = uniform_u32(smem_s), u_tmem = uniform_u32(tmem_base);
    if (elect_one()) {
      const uint32_t idesc_dg = make_idesc_bf16(128, 16, 0, 0);           // K-major x K-major, N = 16
      const uint32_t idesc_wg = make_idesc_bf16(128, 16, 1, 1);           // MN-major x MN-major (contraction over rows)
      const uint64_t pk128 = make_smem_desc(0, 16, 1024, SWZ_128B);       // K-major, 128-byte rows
      const uint64_t pk32 = make_smem_desc(0, 16, 256, SWZ_32B);          // K-major, 32-byte rows
      const uint64_t pdz = make_smem_desc(0, 16384, 1024, SWZ_128B);      // MN-major: 64-channel chunks 16 KB apart
      const uint64_t ph2 = make_smem_desc(0, 4096, 256, SWZ_32B);         // MN-major: [rows][16]
      // W_j = rows j*16 .. j*16+15 of the [48 x 64-channel] blocks (2 KB per tap and block: 2 x 1024-byte row groups)
      const uint64_t d_wd0 = desc_at(pk128, u_smem + kFb2Wd), d_wd1 = desc_at(pk128, u_smem + kFb2Wd + 6144);
      const uint64_t d_wres = desc_at(pk32, u_smem + kFb2Wres), d_ones = desc_at(ph2, u_smem + kFb2Ones);
      const uint32_t d2 = u_tmem + kFb2D2Col;
      uint32_t zk = 0, zph = 0;                             // dZ tile ring position of the current tile
      RingPos xw{0, 0};                                     // X ring position of input step tau = t (window start)
      uint32_t first = 1;
      RingPos hbase{0, 0};
      for (int b = b0; b < p.B; b += bstep) {
        RingPos hwin = hbase;
        int n_hw = 0;                                       // H2 slices of this item already waited for
        for (int t = 0; t < T2; ++t) {
          STGCN_CSTAMP(b == b0 && t >= 4 && t < 7, 160 + 0 + (t - 4) * 8);
          mbar_wait_a(zfull_a + zk * 8, zph);
          STGCN_CSTAMP(b == b0 && t >= 4 && t < 7, 160 + 1 + (t - 4) * 8);
          RingPos hp = hwin;
          uint32_t hs[kFb2Kt];
#pragma unroll
          for (int j = 0; j < kFb2Kt; ++j, hp.advance(kFb2HStages)) {
            if (t + j >= n_hw) { mbar_wait_a(hfull_a + hp.s * 8, hp.ph); n_hw = t + j + 1; }
            hs[j] = hp.s;
          }
          STGCN_CSTAMP(b == b0 && t >= 4 && t < 7, 160 + 2 + (t - 4) * 8);
          // accumulators X_t, X_{t+1}, X_{t+2}; the ones this tile writes FIRST (X_{t+2}; all three at t = 0) must have been
          // drained by the epilogue of their previous use
          RingPos xp = xw;
          uint32_t xs[kFb2Kt];
#pragma unroll
          for (int j = 0; j < kFb2Kt; ++j, xp.advance(kFb2NX)) {
            if (j == kFb2Kt - 1 || t == 0) mbar_wait_a(xfree_a + xp.s * 8, xp.ph ^ 1);
            xs[j] = xp.s;
          }
          tc_fence_after();
          STGCN_CSTAMP(b == b0 && t >= 4 && t < 7, 160 + 3 + (t - 4) * 8);
          const uint32_t dz_s = u_smem + kFb2Dz + zk * 32768;
          // ---- data gradient + residual
          {
            const uint64_t a0 = desc_at(pk128, dz_s), a1 = desc_at(pk128, dz_s + 16384);
#pragma unroll
            for (int j = 0; j < kFb2Kt; ++j) {
              const uint32_t xd = u_tmem + xs[j] * 16;
              const uint32_t fresh = (j == kFb2Kt - 1 || t == 0) ? 1u : 0u;
#pragma unroll
              for (int ks = 0; ks < 4; ++ks)
                mma_bf16_ss(xd, a0 + 2 * ks, d_wd0 + j * (2048 >> 4) + 2 * ks, idesc_dg, (fresh && ks == 0) ? 0u : 1u);
#pragma unroll
              for (int ks = 0; ks < 4; ++ks) mma_bf16_ss(xd, a1 + 2 * ks, d_wd1 + j * (2048 >> 4) + 2 * ks, idesc_dg, 1);
            }
            mma_bf16_ss(u_tmem + xs[kFb2Kt - 1] * 16, a0, d_wres, idesc_dg, 1);
            mma_commit_a(xfull_a + xs[0] * 8);              // X_t is complete
            if (t == T2 - 1) { mma_commit_a(xfull_a + xs[1] * 8); mma_commit_a(xfull_a + xs[2] * 8); }
          }
          STGCN_CSTAMP(b == b0 && t >= 4 && t < 7, 160 + 4 + (t - 4) * 8);
          // ---- weight and bias gradients
          {
            const uint64_t az = desc_at(pdz, dz_s);
#pragma unroll
            for (int j = 0; j < kFb2Kt; ++j) {
              const uint64_t bh = desc_at(ph2, u_smem + kFb2HRing + hs[j] * 4096);
#pragma unroll
              for (int ks = 0; ks < 8; ++ks)
                mma_bf16_ss(d2 + j * 16, az + ks * (2048 >> 4), bh + ks * (512 >> 4), idesc_wg, (first && ks == 0) ? 0u : 1u);
            }
#pragma unroll
            for (int ks = 0; ks < 8; ++ks)
              mma_bf16_ss(d2 + 48, az + ks * (2048 >> 4), d_ones, idesc_wg, (first && ks == 0) ? 0u : 1u);
            first = 0;
          }
          STGCN_CSTAMP(b == b0 && t >= 4 && t < 7, 160 + 5 + (t - 4) * 8);
          mma_commit_a(zempty_a + zk * 8);
          mma_commit_a(hempty_a + hwin.s * 8);
          hwin.advance(kFb2HStages);
          if (t == T2 - 1)
            for (int e = 0; e < kFb2Kt - 1; ++e, hwin.advance(kFb2HStages)) mma_commit_a(hempty_a + hwin.s * 8);
          if (++zk == kFb2NZ) { zk = 0; zph ^= 1; }
          xw.advance(kFb2NX);
        }
        xw.advance_by(kFb2Kt - 1, kFb2NX);                  // the item used T1 = T2 + 2 accumulators
        hbase.advance_by((uint32_t)T1, kFb2HStages);
      }
      mma_commit_a(done_a);
    }
  } else if (warp >= 2 + kFb2EpiWarps) {
    // =========================== E2 warps: dH2 of every input step, then the weight-gradient flush =================
    // (Four dedicated warps: with the step rotating over the E1 warp groups, one group was ~600 cycles late at every
    // tile's rendezvous -- tile period 4400 cycles against 3300 of E1 arithmetic, profiles/r02_ab_batch_h.md.)
    const int q = warp & 3;                               // TMEM lane quarter
    const int row = q * 32 + lane, n = n0 + row;
    const bool valid = n < p.N;
    const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
    uint32_t xi = 0, xph = 0;                             // X ring position of the current input step
    for (int b = b0; b < p.B; b += bstep) {
      for (int tau = 0; tau < T1; ++tau) {
        STGCN_CSTAMP(b == b0 && tau >= 3 && tau < 7 && q == 2, 160 + 52 + (tau - 3) * 2);
        mbar_wait_a(xfull_a + xi * 8, xph);
        tc_fence_after();
        uint32_t rr[16];
        tmem_ld_32x32b_x16(t_lane + xi * 16, rr);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_a(xfree_a + xi * 8);
        STGCN_CSTAMP(b == b0 && tau >= 3 && tau < 7 && q == 2, 160 + 53 + (tau - 3) * 2);
        if (valid) {
          float acc[16];
#pragma unroll
          for (int e = 0; e < 16; ++e) acc[e] = __uint_as_float(rr[e]);
          uint4* dst = reinterpret_cast<uint4*>(p.dh2 + (((long long)b * T1 + tau) * p.N + n) * kFb2Ci);
          dst[0] = pack8_bf16(acc); dst[1] = pack8_bf16(acc + 8);
        }
        if (++xi == kFb2NX) { xi = 0; xph ^= 1; }
      }
    }
    // ---- weight-gradient flush: G[o][(j, c)] -> dwt[(j * 16 + c) * 128 + o], bias row behind
    if (b0 < p.B) {
      mbar_wait_a(done_a, 0);
      tc_fence_after();
#pragma unroll
      for (int j = 0; j < kFb2Kt; ++j) {
        uint32_t rr[16];
        tmem_ld_32x32b_x16(t_lane + kFb2D2Col + j * 16, rr);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 16; ++c) atomicAdd(p.dwt + (j * 16 + c) * kFb2W + row, __uint_as_float(rr[c]));
      }
      uint32_t rb[8];
      tmem_ld_32x32b_x8(t_lane + kFb2D2Col + 48, rb);
      tmem_ld_wait();
      atomicAdd(p.dwt + 48 * kFb2W + row, __uint_as_float(rb[0]));
    }
  } else {
    // =========================== E1 warps ==========================
    // E1 mapping: COALESCED.  The dY / H3 / Q rows of a tile are 128 contiguous 128-byte rows, so warp w takes rows
    // w*8 .. w*8+7 and lane l the 16-byte chunk (l & 7) of rows w*8 + (l >> 3) and + 4: every LDG.128 of a warp reads 512
    // contiguous bytes.  (With the TMEM-style mapping -- one row per thread -- every warp-wide load touched 32 different
    // 128-byte lines: 16 loads x 32 L1 wavefronts x 16 warps ~ 8000 cycles per tile, 5 us per tile measured,
    // profiles/r02_ab_batch_h.md.)  E1 never reads tensor memory, so nothing ties it to the lane = row layout.
    const int ew = warp - 2, c8 = lane & 7;
    const int er[2] = {ew * 8 + (lane >> 3), ew * 8 + 4 + (lane >> 3)};       // tile rows of this thread
    const bool ev[2] = {n0 + er[0] < p.N, n0 + er[1] < p.N};
    float gw[2][8];                                       // LayerNorm weight of this thread's (vertex, 8 channels) x 2 rows
#pragma unroll
    for (int r2 = 0; r2 < 2; ++r2) {
      const float4* gp = reinterpret_cast<const float4*>(p.gamma + (long long)(ev[r2] ? n0 + er[r2] : 0) * kFb2Co + c8 * 8);
      const float4 g0 = gp[0], g1 = gp[1];
      gw[r2][0] = g0.x; gw[r2][1] = g0.y; gw[r2][2] = g0.z; gw[r2][3] = g0.w;
      gw[r2][4] = g1.x; gw[r2][5] = g1.y; gw[r2][6] = g1.z; gw[r2][7] = g1.w;
    }
    // staging offsets of this thread's two 16-byte chunks inside a [128 rows][128 B] 128B-swizzled sub-tile
    const uint32_t so[2] = {(uint32_t)er[0] * 128u + (((uint32_t)c8 ^ ((uint32_t)er[0] & 7u)) << 4),
                            (uint32_t)er[1] * 128u + (((uint32_t)c8 ^ ((uint32_t)er[1] & 7u)) << 4)};
    // The operands of the NEXT tile are requested while the current one is computed -- by cp.async into a private
    // shared-memory slot per thread, not into registers: with 24 prefetch registers live across the arithmetic the
    // compiler (80-96 registers per thread at this CTA size) sank the loads to the end of the tile and the first use
    // of the loaded values was the kernel's top stall (ncu source page, profiles/r02_ab_batch_h.md).
    const int et = (warp - 2) * 32 + lane;                // 0..511
    const uint32_t pf_s = smem_s + kFb2Pf + (uint32_t)et * 16u;      // chunk k of this thread at + k * 8192
    float sc[4];                                          // mean, rstd, s1, s2 of the next tile's (b, t) group
    auto fetch = [&](int b, int t) {
      sc[0] = sc[1] = sc[2] = sc[3] = 0.f;
      if (b < p.B) {
        const long long g = (long long)b * T2 + t;
        sc[0] = p.mean[g]; sc[1] = p.rstd[g];
        for (int k = 0; k < p.n_parts; ++k) { sc[2] += p.sums[k * p.part_stride + 2 * g]; sc[3] += p.sums[k * p.part_stride + 2 * g + 1]; }
#pragma unroll
        for (int r2 = 0; r2 < 2; ++r2) {
          const bool ok = ev[r2];
          const long long off = ok ? (g * p.N + n0 + er[r2]) * kFb2Co + c8 * 8 : 0;
          const uint32_t nb = ok ? 16u : 0u;              // rows past N: zero fill
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(pf_s + (uint32_t)(r2 * 3 + 0) * 8192u), "l"(p.dy + off), "r"(nb) : "memory");
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(pf_s + (uint32_t)(r2 * 3 + 1) * 8192u), "l"(p.h3 + off), "r"(nb) : "memory");
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(pf_s + (uint32_t)(r2 * 3 + 2) * 8192u), "l"(p.q + off), "r"(nb) : "memory");
        }
      }
      cp_async_commit();
    };
    auto take = [&](uint4 (&dv)[2], uint4 (&hv)[2], uint4 (&qv)[2]) {      // this thread's own slots: no cross-thread hazard
      cp_async_wait<0>();
#pragma unroll
      for (int r2 = 0; r2 < 2; ++r2) {
        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(dv[r2].x), "=r"(dv[r2].y), "=r"(dv[r2].z), "=r"(dv[r2].w) : "r"(pf_s + (uint32_t)(r2 * 3 + 0) * 8192u) : "memory");
        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(hv[r2].x), "=r"(hv[r2].y), "=r"(hv[r2].z), "=r"(hv[r2].w) : "r"(pf_s + (uint32_t)(r2 * 3 + 1) * 8192u) : "memory");
        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(qv[r2].x), "=r"(qv[r2].y), "=r"(qv[r2].z), "=r"(qv[r2].w) : "r"(pf_s + (uint32_t)(r2 * 3 + 2) * 8192u) : "memory");
      }
    };
    uint32_t gt = 0;                                      // tiles done by this CTA (all warps count alike)
    fetch(b0, 0);
    for (int b = b0; b < p.B; b += bstep) {
      for (int i = 0; i < T2; ++i) {
        {
          STGCN_CSTAMP(b == b0 && i >= 4 && i < 7 && warp == 2, 160 + 24 + (i - 4) * 8);
          STGCN_CSTAMP(b == b0 && i >= 4 && i < 7 && warp == 17, 160 + 48 + (i - 4));
          // ---------------- E1: dZ tile of output step i
          uint4 dv[2], hv[2], qv[2];
          take(dv, hv, qv);
          const float mu = sc[0], rs = sc[1], s1 = sc[2], s2 = sc[3];
          if (i + 1 < T2) fetch(b, i + 1); else fetch(b + bstep, 0);      // the slots are free again: values are in registers
          uint4 pu[2], pq[2];                                 // packed dP / dQ of this thread's 2 x 8 channels
          const float c1 = -mu * rs;                          // xhat = fma(h, rs, c1)
#pragma unroll
          for (int r2 = 0; r2 < 2; ++r2) {
            float df[8], hf[8], qf[8], du[8], dq[8];
            unpack8_bf16(dv[r2], df); unpack8_bf16(hv[r2], hf); unpack8_bf16(qv[r2], qf);
            const float hrs = ev[r2] ? 0.5f * rs : 0.f;       // rows past N produce a zero dZ row; the 0.5 of the sigmoid rides here
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              // sigma(q) = 0.5 + 0.5 th, th = tanh(q / 2):  dP = dH3 sigma = hu + hu th,  dQ = dH3 h (1 - sigma) = hu h - hu h th
              const float xh = fmaf(hf[e], rs, c1);
              const float hu = hrs * fmaf(-xh, s2, fmaf(df[e], gw[r2][e], -s1));     // dH3 / 2
              const float th = tanh_approx(0.5f * qf[e]);
              const float huh = hu * hf[e];
              du[e] = fmaf(hu, th, hu);
              dq[e] = fmaf(-huh, th, huh);
            }
            pu[r2] = pack8_bf16(du); pq[r2] = pack8_bf16(dq);
          }
          const uint32_t zk = gt % kFb2NZ, zu = gt / kFb2NZ;
          STGCN_CSTAMP(b == b0 && i >= 4 && i < 7 && warp == 2, 160 + 25 + (i - 4) * 8);
          mbar_wait_a(zempty_a + zk * 8, (zu & 1) ^ 1);
          STGCN_CSTAMP(b == b0 && i >= 4 && i < 7 && warp == 2, 160 + 26 + (i - 4) * 8);
          const uint32_t dzs = smem_s + kFb2Dz + zk * 32768;
#pragma unroll
          for (int r2 = 0; r2 < 2; ++r2) {
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dzs + so[r2]), "r"(pu[r2].x), "r"(pu[r2].y), "r"(pu[r2].z), "r"(pu[r2].w) : "memory");
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dzs + 16384u + so[r2]), "r"(pq[r2].x), "r"(pq[r2].y), "r"(pq[r2].z), "r"(pq[r2].w) : "memory");
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) mbar_arrive_a(zfull_a + zk * 8);
          STGCN_CSTAMP(b == b0 && i >= 4 && i < 7 && warp == 2, 160 + 27 + (i - 4) * 8);
          ++gt;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 256);
}

// shapes this kernel serves
inline bool fb2_supported(int c_in, int c_out, int Kt, int act, int T_in, long long rows_out, int N) {
  return c_in == kFb2Ci && c_out == kFb2Co && Kt == kFb2Kt && act == STGCN_ACT_GLU && T_in >= Kt && rows_out > 0 &&
         rows_out < (1LL << 31) && N >= 1 && (N + 127) / 128 <= 64;    // (no CUDA call here: the sizing pass also runs without a GPU)
}

inline void launch_fb2(const Fb2Params& p0, cudaStream_t stream) {
  Fb2Params p = p0;
  p.nnt = (p.N + 127) / 128;
  p.dbg = g_tap_dbg;
  int per = sm_count() / p.nnt;                            // CTAs per vertex tile
  if (per > p.B) per = p.B;
  if (per < 1) per = 1;
  const int grid = per * p.nnt;
  STGCN_CUDA(cudaFuncSetAttribute(umma_fb2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFb2Smem));
  STGCN_LAUNCH_NAMED("umma_fb2_kernel", umma_fb2_kernel, grid, kFb2Threads, kFb2Smem, stream, p);
}

}  // namespace umma
}  // namespace stgcn
