// umma_fb0.cuh -- backward of the FIRST temporal convolution of the model (Cin = 1: layers.py:87-105 on the raw
// (B,1,T,N) window) fused with the data gradient of the graph-conv layer's 1x1 align conv (layers.py:16,225) that
// feeds it, on tcgen05 (bf16 throughput mode).  One persistent kernel replaces
//     lowrank_expand_kernel   dH1[r, 0:64] = dX0[r, 0:16] . Wa            (wrote 75 MB, 68 us at B = 256)
//   + smallc1_gate_wgrad      dZ = GLU'(dH1; z recomputed from x), dW = dZ^T . x-window   (read them back, 110 us; the
//                             384 fp32 weight-gradient accumulators per thread made it issue bound at 10.6 % of HBM)
// and reads only the 16-channel gradient dX0 (19 MB) and the 3 MB input:
//
//   per tile of 128 flat rows r = (b, t, n):
//     producer warp            dX0 tile [128 x 16] by cp.async; window tile xw [128 r][16] = (x_t, x_t+1, x_t+2, x_t, x_t+1,
//                              x_t+2, 1, 1, x_res, 0 ...) built from the input taps (32-byte rows, 32B swizzle)
//     MMA 1   D1[128 r x 64]   = dX0 tile (K-major)  x  Wa^T [64 x 16] (resident)
//     MMA 0   D0[128 r x 128]  = xw (K-major A)  x  [w_hi | w_lo | b_hi, b_lo | res] [128 o x 16]: the pre-activations
//                              P | Q with bias and zero-padded residual, fp32-exact weights through the hi / lo split
//     epilogue (16 warps)      dH1, P, Q from tensor memory;  dU = dH1 * s,  dQ = dH1 * P * s (1 - s),  s = sigmoid(Q)
//                              -> bf16 dZ tile [128 r][128 o] in shared memory (MN-major A operand, 128B swizzle)
//     MMA 2   D2[128 o x 16]  += dZ^T . xw (the same tile, MN-major B)   (K = 128 rows; accumulates over ALL tiles of the
//                              CTA in TMEM)
//   end:      D2 columns 0..Kt-1 = dW taps, column 6 = bias gradient -> fp32 atomics into dwt[(k) * 128 + o].
//   (Up to GPU call AA the epilogue threads recomputed P, Q themselves from the taps -- 6 FMA + 2 LDS.128 per element,
//   40 % of their instructions: 72 us instead of 58, profiles/r02_ab_batch_h.md.)
//
// The block-0 input needs no data gradient (it is the model input), so dZ never reaches HBM.
// Serves the default architecture's first block: c_in = 1, 64 GLU channels, 16 graph-conv channels, Kt in {2, 3}.
//
// MODE 1 (FB_GATE) -- the same front end for the temporal convs of the LATER blocks (c_in >= 16, GLU with the q-only saved
// state): MMA 1 as above; the saved gate half Q and the layer output H1 arrive by TMA into 128B-swizzled tiles, the epilogue
// forms dU = dH1 * s, dQ = dH1 * H1 * (1 - s) and the dZ = (dU | dQ) tile leaves by TMA store for the data-gradient and
// weight-gradient kernels.  Replaces lowrank_expand_kernel + gate_vec_kernel (33 + 51 us at B = 256: dH1 written and read
// back at 64 channels).
#pragma once
#include "umma_tap.cuh"

namespace stgcn {
namespace umma {

constexpr int kFb0EpiWarps = 16;
constexpr int kFb0Threads = 64 + 32 * kFb0EpiWarps;      // warp 0 producer, warp 1 MMA issuer, 16 epilogue warps
constexpr int kFb0Stages = 8;                            // dX0 tiles in flight (4 KB each)

enum { FB_FIRST = 0, FB_GATE = 1 };
// FB_FIRST: the pre-activations P, Q of the tile are not recomputed by the epilogue threads (6 FMA + 2 LDS.128 per element
// were 40 % of their instructions: 72 -> 58 us): a producer-built window tile [x_t, x_t+1, x_t+2, x_t, x_t+1, x_t+2, 1, 1,
// x_res, 0 ...] times [w_hi | w_lo | b_hi, b_lo | res] is one more tcgen05.mma (K = 16, N = 128, fp32-exact weights
// through the hi / lo split) into tensor memory, and the same tile is the MN-major B operand of the weight-gradient MMA
// (taps in columns 0..2, bias in column 6).
struct Fb0Params {
  const bf16* dst0;        // [rows, 16] gradient w.r.t. the aligned (16-channel) graph-conv input
  const bf16* wa;          // [64][16] K-major: wa[j * 16 + o] = align_w[o][j]
  const bf16* x;           // [B, T_in, N] model input (Cin = 1)
  const float* wt;         // [Kt][128] forward-layout conv weights: wt[k * 128 + o] = conv_w[o][0][k]
  const float* bias;       // [128]
  float* dwt;              // [(Kt + 1)][128], pre-zeroed: taps then bias row
  long long rows;
  int n_tiles, Kt, T_out, T_in, N, explicit_res;
  // FB_GATE
  const bf16* q;           // [rows, 64] saved gate half of the pre-activation
  const bf16* h;           // [rows, 64] layer output
  bf16* dz;                // [rows, 128] out: (dU | dQ)
};

// shared-memory map (offsets from the 1024-aligned base)
constexpr uint32_t kFb0ARing = 0;                                   // kFb0Stages x 4096
constexpr uint32_t kFb0Wa = kFb0ARing + kFb0Stages * 4096;          // 2048
constexpr uint32_t kFb0Wpq = kFb0Wa + 2048;                         // (unused since the P / Q recompute moved to the tensor pipe)
constexpr int kFb0ND1 = 3;                                          // dH1 accumulators in TMEM (64 columns each); 3 x 64 + 16 fit a 256-column
                                                                    // allocation, which leaves room for a weight-gradient kernel of the helper stream on the same SM
constexpr int kFb0NZ = 3;                                           // dZ / x-window tile pairs in shared memory
constexpr uint32_t kFb0X3 = kFb0Wpq + 2048;                         // kFb0NZ x 4096
constexpr uint32_t kFb0Dz = kFb0X3 + kFb0NZ * 4096;                 // kFb0NZ x 32768 (1024-aligned: 36864 + 12288 = 49152)
constexpr uint32_t kFb0Smem = kFb0Dz + kFb0NZ * 32768 + 1024;
// FB_GATE only: the saved Q and H1 tiles [128 rows][64 ch] arrive by TMA (128B swizzle), two stages of 2 x 16 KB
constexpr int kFb0NQH = 2;
constexpr uint32_t kFb0QH = kFb0Dz + kFb0NZ * 32768;
constexpr uint32_t kFb0SmemGate = kFb0QH + kFb0NQH * 32768 + 1024;
// FB_FIRST (same allocation): the [128 o][16] weight tile and a ring of window tiles live where FB_GATE keeps Q / H1
constexpr int kFb0NXw = 4;
constexpr uint32_t kFb0Wtc = kFb0QH, kFb0Xw = kFb0QH + 4096;
constexpr uint32_t kFb0D0Col = 256;                                 // FB_FIRST: P | Q accumulators, 2 x 128 columns
constexpr uint32_t kFb0D2Col = kFb0ND1 * 64;                        // TMEM column of the weight-gradient accumulator
// (the first version double-buffered both: 2.4 us per 128-row tile against ~0.9 us of epilogue issue time -- every tile
// waited for the previous tile's MMA 2 and the next tile's MMA 1 in turn; profiles/r02_ab_batch_c.md)

template <int MODE>
__global__ void __launch_bounds__(kFb0Threads, 1)
umma_fb0_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmH,
                const __grid_constant__ CUtensorMap tmZ, Fb0Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ __align__(8) uint64_t a_full[kFb0Stages], a_empty[kFb0Stages], d1_full[kFb0ND1], d1_empty[kFb0ND1], dz_full[kFb0NZ], dz_empty[kFb0NZ], done,
      qh_full[kFb0NQH], qh_empty[kFb0NQH], xw_full[kFb0NXw], xw_empty[kFb0NXw];
  __shared__ uint32_t tmem_base_s;
  constexpr bool kFirst = MODE != FB_GATE, kTC = kFirst;
  constexpr uint32_t kTmemCols = kTC ? 512 : 256;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // ---- one-time staging: Wa^T (K-major, 32-byte rows, 32B swizzle) and the per-channel (w_0..w_2, bias) quads
  if (threadIdx.x < 128) {
    const int j = threadIdx.x >> 1, h = threadIdx.x & 1;
    const uint4 v = *reinterpret_cast<const uint4*>(p.wa + j * 16 + h * 8);
    *reinterpret_cast<uint4*>(smem + kFb0Wa + j * 32 + ((h ^ ((j >> 2) & 1)) << 4)) = v;
  } else if (kTC && threadIdx.x < 256) {
    const int o = threadIdx.x - 128;
    float v[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) v[k] = 0.f;
    for (int k = 0; k < 3; ++k) {
      const float w = k < p.Kt ? p.wt[k * 128 + o] : 0.f;
      const float hi = __bfloat162float(__float2bfloat16_rn(w));
      v[k] = hi; v[3 + k] = w - hi;
    }
    const float b = p.bias[o], bh = __bfloat162float(__float2bfloat16_rn(b));
    v[6] = bh; v[7] = b - bh;
    v[8] = (o == 0 && p.explicit_res) ? 1.f : 0.f;
    const int sw = (o >> 2) & 1;
    *reinterpret_cast<uint4*>(smem + kFb0Wtc + o * 32 + ((0 ^ sw) << 4)) = pack8_bf16(v);
    *reinterpret_cast<uint4*>(smem + kFb0Wtc + o * 32 + ((1 ^ sw) << 4)) = pack8_bf16(v + 8);
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < kFb0Stages; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
    for (int i = 0; i < kFb0ND1; ++i) { mbar_init(&d1_full[i], 1); mbar_init(&d1_empty[i], kFb0EpiWarps); }
    for (int i = 0; i < kFb0NZ; ++i) { mbar_init(&dz_full[i], kFb0EpiWarps); mbar_init(&dz_empty[i], 1); }
    mbar_init(&done, 1);
    for (int i = 0; i < kFb0NQH; ++i) { mbar_init(&qh_full[i], 1); mbar_init(&qh_empty[i], kFb0EpiWarps); }
    for (int i = 0; i < kFb0NXw; ++i) { mbar_init(&xw_full[i], 1); mbar_init(&xw_empty[i], 1); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&tmem_base_s, kTmemCols);    // D1: kFb0ND1 x 64 columns, D2: 16 columns (FB_FIRST: + 2 x 128)
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  const int n_my = p.n_tiles > (int)blockIdx.x ? (p.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

  if (warp == 0) {
    // =========================== producer: dX0 tiles by cp.async ================================
    int pending = -1;
    unsigned short xnx[4][3];                             // FB_FIRST: input taps of the next tile's rows lane + 32 c
    auto load_taps = [&](int i) {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        xnx[c][0] = xnx[c][1] = xnx[c][2] = 0;
        if (i >= n_my) continue;
        const long long r = ((long long)blockIdx.x + (long long)i * gridDim.x) * 128 + lane + 32 * c;
        if (r >= p.rows) continue;
        long long in0; int t_unused;
        simt::row_decode(r, p.T_out * p.N, p.N, (long long)p.T_in * p.N, in0, t_unused);
        const unsigned short* xs = reinterpret_cast<const unsigned short*>(p.x);
        xnx[c][0] = xs[in0];
        xnx[c][1] = xs[in0 + p.N];
        if (p.Kt > 2) xnx[c][2] = xs[in0 + 2LL * p.N];
      }
    };
    if (kTC) load_taps(0);
    for (int i = 0; i < n_my; ++i) {
      const long long r0 = ((long long)blockIdx.x + (long long)i * gridDim.x) * 128;
      const uint32_t s = i % kFb0Stages, ph = (i / kFb0Stages) & 1;
      if (MODE == FB_GATE && lane == 0) {
        // saved gate half Q and layer output H1 of this tile: two [128 rows x 64 ch] boxes, rows past the end read as zeros
        const uint32_t s2 = i % kFb0NQH, ph2 = (i / kFb0NQH) & 1;
        mbar_wait(&qh_empty[s2], ph2 ^ 1);
        mbar_arrive_expect_tx(&qh_full[s2], 32768);
        tma_load_2d(smem + kFb0QH + s2 * 32768, &tmQ, &qh_full[s2], 0, (int)r0);
        tma_load_2d(smem + kFb0QH + s2 * 32768 + 16384, &tmH, &qh_full[s2], 0, (int)r0);
      }
      if (kTC) {
        // window tile of this tile's rows; the taps of the NEXT tile are requested first (raw bits, no use behind the load)
        unsigned short xc[4][3];
#pragma unroll
        for (int c = 0; c < 4; ++c) { xc[c][0] = xnx[c][0]; xc[c][1] = xnx[c][1]; xc[c][2] = xnx[c][2]; }
        load_taps(i + 1);
        const uint32_t sx = i % kFb0NXw, phx = (i / kFb0NXw) & 1;
        mbar_wait(&xw_empty[sx], phx ^ 1);
        uint8_t* xd = smem + kFb0Xw + sx * 4096;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int row = lane + 32 * c;
          uint4 a = make_uint4(0, 0, 0, 0), b = make_uint4(0, 0, 0, 0);
          if (r0 + row < p.rows) {
            const uint32_t x0 = xc[c][0], x1 = xc[c][1], x2 = xc[c][2];
            a = make_uint4(x0 | (x1 << 16), x2 | (x0 << 16), x1 | (x2 << 16), 0x3F803F80u);
            b.x = p.explicit_res ? (p.Kt > 2 ? x2 : x1) : 0u;
          }
          const int sw = (row >> 2) & 1;
          *reinterpret_cast<uint4*>(xd + row * 32 + ((0 ^ sw) << 4)) = a;
          *reinterpret_cast<uint4*>(xd + row * 32 + ((1 ^ sw) << 4)) = b;
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&xw_full[sx]);
      }
      __syncwarp();
      mbar_wait(&a_empty[s], ph ^ 1);
      uint8_t* dst = smem + kFb0ARing + s * 4096;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const int q = lane + 32 * c, row = q >> 1, h = q & 1;
        const bool ok = r0 + row < p.rows;
        const bf16* src = p.dst0 + (ok ? (r0 + row) : 0) * 16 + h * 8;
        cp_async16(dst + row * 32 + ((h ^ ((row >> 2) & 1)) << 4), src, ok ? 16u : 0u);
      }
      cp_async_commit();
      if (pending >= 0) {                                 // the previous tile has landed after this wait
        cp_async_wait<1>();
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&a_full[pending]);
      }
      pending = (int)s;
    }
    if (pending >= 0) {
      cp_async_wait<0>();
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&a_full[pending]);
    }
  } else if (warp == 1) {
    // =========================== MMA issuer =============================
    if (elect_one()) {
      const uint32_t idesc1 = make_idesc_bf16(128, 64, 0, 0);            // K-major x K-major
      const uint32_t idesc2 = make_idesc_bf16(128, 16, 1, 1);            // MN-major x MN-major (contraction over rows)
      const uint64_t pk32 = make_smem_desc(0, 16, 256, SWZ_32B);         // K-major, 32-byte rows
      const uint64_t pdz = make_smem_desc(0, 16384, 1024, SWZ_128B);     // MN-major: 64-channel chunks 16 KB apart
      const uint64_t px3 = make_smem_desc(0, 4096, 256, SWZ_32B);        // MN-major: [rows][16]
      const uint32_t wa_s = smem_u32(smem + kFb0Wa);
      auto mma1 = [&](int i) {
        const uint32_t s = i % kFb0Stages, ph = (i / kFb0Stages) & 1, ab = i % kFb0ND1, aph = (i / kFb0ND1) & 1;
        mbar_wait(&a_full[s], ph);
        mbar_wait(&d1_empty[ab], aph ^ 1);
        if (kTC) {
          mbar_wait(&xw_full[i % kFb0NXw], (i / kFb0NXw) & 1);
          // P | Q slot i & 1 was last read by the epilogue of tile i - 2 (its d1_empty arrival comes after those loads)
          if (i >= 2) mbar_wait(&d1_empty[(i - 2) % kFb0ND1], ((i - 2) / kFb0ND1) & 1);
        }
        tc_fence_after();
        mma_bf16_ss(tmem_base + ab * 64, desc_at(pk32, smem_u32(smem + kFb0ARing + s * 4096)), desc_at(pk32, wa_s), idesc1, 0);
        if (kTC)
          mma_bf16_ss(tmem_base + kFb0D0Col + (i & 1) * 128, desc_at(pk32, smem_u32(smem + kFb0Xw + (i % kFb0NXw) * 4096)),
                      desc_at(pk32, smem_u32(smem + kFb0Wtc)), make_idesc_bf16(128, 128, 0, 0), 0);
        mma_commit(&d1_full[ab]);
        mma_commit(&a_empty[s]);
      };
      int next1 = 0;                                      // MMA 1 runs up to kFb0ND1 - 1 tiles ahead of MMA 2
      for (int i = 0; i < n_my; ++i) {
        for (; next1 < n_my && next1 < i + (kTC ? 2 : kFb0ND1); ++next1) mma1(next1);
        const uint32_t zb = i % kFb0NZ, zph = (i / kFb0NZ) & 1;
        if (!kFirst) {
          // FB_GATE: the finished dZ tile leaves through two TMA stores (P half | Q half); the buffer is free once read
          mbar_wait(&dz_full[zb], zph);
          const long long r0 = ((long long)blockIdx.x + (long long)i * gridDim.x) * 128;
          tma_store_2d(&tmZ, smem + kFb0Dz + zb * 32768, 0, (int)r0);
          tma_store_2d(&tmZ, smem + kFb0Dz + zb * 32768 + 16384, 64, (int)r0);
          tma_store_commit();
          tma_store_wait_read<0>();
          mbar_arrive(&dz_empty[zb]);
          continue;
        }
        mbar_wait(&dz_full[zb], zph);
        tc_fence_after();
        uint64_t da = desc_at(pdz, smem_u32(smem + kFb0Dz + zb * 32768)),
                 db = desc_at(px3, smem_u32(smem + kFb0Xw + (i % kFb0NXw) * 4096));
#pragma unroll
        for (int k = 0; k < 8; ++k) {                     // 16 rows per instruction
          mma_bf16_ss(tmem_base + kFb0D2Col, da, db, idesc2, (i != 0 || k != 0) ? 1u : 0u);
          da += 2048 >> 4; db += 512 >> 4;
        }
        mma_commit(&dz_empty[zb]);
        if (kTC) mma_commit(&xw_empty[i % kFb0NXw]);
      }
      if (!kFirst) tma_store_wait_all<0>();
      mma_commit(&done);
    }
  } else {
    // =========================== epilogue warps ==========================
    const int q = warp & 3, grp = (warp - 2) >> 2;        // TMEM lane quarter; 16-channel group
    const int row = q * 32 + lane, c0 = grp * 16;
    for (int i = 0; i < n_my; ++i) {
      const long long r = ((long long)blockIdx.x + (long long)i * gridDim.x) * 128 + row;
      const bool valid = r < p.rows;
      uint4 qv[2], hv[2];
      if (MODE == FB_GATE) {
        // Q / H1 chunks of this thread's row from the TMA-staged tiles (one row per thread straight from global memory
        // touched 32 different 128-byte lines per warp-wide load: ~4000 L1 wavefront cycles per tile with the dZ stores,
        // profiles/r02_ab_batch_h.md); the 128B swizzle makes the row-per-thread reads conflict free
        const uint32_t s2 = i % kFb0NQH, ph2 = (i / kFb0NQH) & 1;
        mbar_wait(&qh_full[s2], ph2);
        const uint8_t* qs_ = smem + kFb0QH + s2 * 32768 + row * 128;
        const int ch = c0 >> 3, sw = row & 7;
        qv[0] = *reinterpret_cast<const uint4*>(qs_ + ((ch ^ sw) << 4));
        qv[1] = *reinterpret_cast<const uint4*>(qs_ + (((ch + 1) ^ sw) << 4));
        hv[0] = *reinterpret_cast<const uint4*>(qs_ + 16384 + ((ch ^ sw) << 4));
        hv[1] = *reinterpret_cast<const uint4*>(qs_ + 16384 + (((ch + 1) ^ sw) << 4));
        __syncwarp();
        if (lane == 0) mbar_arrive(&qh_empty[s2]);
      }
      const uint32_t ab = i % kFb0ND1, aph = (i / kFb0ND1) & 1;
      mbar_wait(&d1_full[ab], aph);
      tc_fence_after();
      uint32_t rr[16], rp[16], rq[16];
      tmem_ld_32x32b_x16(tmem_base + ((uint32_t)(q * 32) << 16) + ab * 64 + c0, rr);
      if (kTC) {
        const uint32_t d0 = tmem_base + ((uint32_t)(q * 32) << 16) + kFb0D0Col + (i & 1) * 128;
        tmem_ld_32x32b_x16(d0 + c0, rp);
        tmem_ld_32x32b_x16(d0 + 64 + c0, rq);
      }
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&d1_empty[ab]);
      float du[16], dq[16];
      if (kTC) {
#pragma unroll
        for (int e = 0; e < 16; ++e) {                    // P (bias and residual included) and Q straight from tensor memory
          const float u = __uint_as_float(rp[e]);
          const float s = sigmoid_tanh_(__uint_as_float(rq[e]));
          const float dh = valid ? __uint_as_float(rr[e]) : 0.f;
          du[e] = dh * s;
          dq[e] = dh * u * s * (1.f - s);
        }
      } else {
        float qf[16], hf[16];
        unpack8_bf16(qv[0], qf); unpack8_bf16(qv[1], qf + 8);
        unpack8_bf16(hv[0], hf); unpack8_bf16(hv[1], hf + 8);
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          const float s = sigmoid_tanh_(qf[e]);
          const float dh = __uint_as_float(rr[e]);
          du[e] = dh * s;
          dq[e] = dh * hf[e] * (1.f - s);
        }
      }
      const uint32_t zb = i % kFb0NZ, zph = (i / kFb0NZ) & 1;
      mbar_wait(&dz_empty[zb], zph ^ 1);
      const uint32_t dzs = smem_u32(smem + kFb0Dz + zb * 32768);
      stage_store8_s(dzs, row, c0, pack8_bf16(du));
      stage_store8_s(dzs, row, c0 + 8, pack8_bf16(du + 8));
      stage_store8_s(dzs + 16384u, row, c0, pack8_bf16(dq));
      stage_store8_s(dzs + 16384u, row, c0 + 8, pack8_bf16(dq + 8));
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&dz_full[zb]);
    }
    // ---- flush: D2[o][k] -> dwt[k * 128 + o]
    if (kFirst && grp == 0 && n_my > 0) {
      mbar_wait(&done, 0);
      tc_fence_after();
      uint32_t rr[16];
      tmem_ld_32x32b_x16(tmem_base + ((uint32_t)(q * 32) << 16) + kFb0D2Col, rr);
      tmem_ld_wait();
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (k <= p.Kt) atomicAdd(p.dwt + k * 128 + row, __uint_as_float(rr[(kTC && k == p.Kt) ? 6 : k]));     // bias column 6
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, kTmemCols);
}

// shapes this kernel serves (all three datasets' first block in the default architecture)
inline bool fb0_supported(int c_in, int c1, int c2, int Kt, int act, long long rows) {
  return c_in == 1 && c1 == 64 && c2 == 16 && (Kt == 2 || Kt == 3) && act == STGCN_ACT_GLU && rows > 0 && rows < (1LL << 31);
}

inline void launch_fb0(const bf16* dst0, const bf16* wa, const bf16* x, const float* wt, const float* bias, float* dwt,
                       long long rows, int Kt, int T_out, int T_in, int N, int explicit_res, cudaStream_t stream) {
  Fb0Params p{};
  p.dst0 = dst0; p.wa = wa; p.x = x; p.wt = wt; p.bias = bias; p.dwt = dwt; p.rows = rows;
  p.n_tiles = (int)((rows + 127) / 128); p.Kt = Kt; p.T_out = T_out; p.T_in = T_in; p.N = N; p.explicit_res = explicit_res;
  const int grid = p.n_tiles < sm_count() ? p.n_tiles : sm_count();
  CUtensorMap none{};                                     // FB_FIRST touches no tensor map
  STGCN_CUDA(cudaFuncSetAttribute(umma_fb0_kernel<FB_FIRST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFb0SmemGate));
  STGCN_LAUNCH_NAMED("umma_fb0_kernel<FIRST>", umma_fb0_kernel<FB_FIRST>, grid, kFb0Threads, kFb0SmemGate, stream, none, none, none, p);
}

// later blocks: dZ = GLU'(dX0 . Wa; Q, H1) for a 64-channel GLU conv in front of a 64 -> 16 align conv (q-only saved state)
inline bool fb_gate_supported(int c1, int c2, int act, long long rows) {
  return c1 == 64 && c2 == 16 && act == STGCN_ACT_GLU && rows > 0 && rows < (1LL << 31);
}
inline void launch_fb_gate(const bf16* dst0, const bf16* wa, const bf16* q, const bf16* h, bf16* dz, long long rows,
                           cudaStream_t stream) {
  Fb0Params p{};
  p.dst0 = dst0; p.wa = wa; p.q = q; p.h = h; p.dz = dz; p.rows = rows; p.n_tiles = (int)((rows + 127) / 128);
  p.Kt = 2; p.T_out = 1; p.T_in = 1; p.N = 1;
  const int grid = p.n_tiles < sm_count() ? p.n_tiles : sm_count();
  // flat-row tensor maps: Q, H1 [rows][64] and dZ [rows][128]; boxes of 64 channels x 128 rows, 128B swizzle
  const uint64_t qd[2] = {64, (uint64_t)rows}, qs[1] = {128}, zd[2] = {128, (uint64_t)rows}, zs[1] = {256};
  const uint32_t box[2] = {64, 128};
  const CUtensorMap tmQ = make_tmap_bf16(q, 2, qd, qs, box, CU_TENSOR_MAP_SWIZZLE_128B);
  const CUtensorMap tmH = make_tmap_bf16(h, 2, qd, qs, box, CU_TENSOR_MAP_SWIZZLE_128B);
  const CUtensorMap tmZ = make_tmap_bf16(dz, 2, zd, zs, box, CU_TENSOR_MAP_SWIZZLE_128B);
  STGCN_CUDA(cudaFuncSetAttribute(umma_fb0_kernel<FB_GATE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFb0SmemGate));
  STGCN_LAUNCH_NAMED("umma_fb0_kernel<GATE>", umma_fb0_kernel<FB_GATE>, grid, kFb0Threads, kFb0SmemGate, stream, tmQ, tmH, tmZ, p);
}

}  // namespace umma
}  // namespace stgcn
