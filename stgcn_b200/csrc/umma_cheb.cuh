// umma_cheb.cuh -- the graph convolution of an ST block as ONE tcgen05 kernel per direction (bf16 path, C = 16):
//
//   forward  (layers.py:154-172 ChebGraphConv, :194-206 GraphConv, :229-231 residual, :253 ReLU)
//       x_1 = Lhat x_0 ;  x_k = 2 Lhat x_{k-1} - x_{k-2} ;  y = relu( sum_k x_k W_k + b + x_0 )
//   backward (autograd of the same lines; Lhat has no gradient, main.py:103)
//       dG = dy * [y > 0] ;  D_k = dG W_k^T ;  D_{k-1} += alpha_k Lhat^T D_k ;  D_{k-2} -= D_k ;  dx_0 = D_0 + dG
//
// The dense operator Lhat (bf16, every 128-row tile, 128B-swizzled K-major) is staged in shared memory ONCE per CTA
// by bulk copies and stays there; a work item is Gb consecutive (b, t) groups = Gb x [N vertices x 16 channels].
// Per item the Chebyshev terms never leave the SM: x_0 comes in by cp.async into a 32B-swizzled buffer that is at
// the same time the MN-major B operand of the node contraction (K = vertex) and the K-major A operand of the
// per-hop weight GEMM (K = channel); each hop is tcgen05.mma (M = 128 vertices, N = Gb*16, K = 16 per instruction)
// into TMEM; the epilogue warps apply the recurrence, write x_k back to shared memory as the next hop's operand and
// to HBM (saved for the weight gradients); the Ks weight GEMMs accumulate into a second TMEM region whose epilogue
// adds bias + residual, applies ReLU and stores y.  The backward kernel runs the adjoint recurrence the same way and
// accumulates Lhat^T D_k directly on top of dG W_{k-1}^T in TMEM.
//
// A CTA runs kChebSlots independent item pipelines ("slots") that share the resident operator: while one slot's
// epilogue warps drain TMEM, the other slot's MMAs run.  Warps: 0 = operator load + TMEM allocation; per slot: one MMA
// issuer, four epilogue warps (one vertex row per thread per tile), two cp.async producer warps.
#pragma once
#include "umma_gso.cuh"

namespace stgcn {
namespace umma {

constexpr int kChebC = 16;
constexpr int kChebSlots = 2;              // independent item pipelines per CTA (one's MMAs overlap the other's epilogue)
constexpr int kChebSlotWarps = 7;          // per slot: 1 MMA issuer + 4 epilogue + 2 cp.async producer warps
constexpr int kChebProducers = 2;
constexpr int kChebThreads = 32 * (1 + kChebSlots * kChebSlotWarps);
constexpr int kChebMaxMT = 4;
constexpr int kChebMaxDepth = 8;

__device__ __forceinline__ void bulk_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

struct ChebParams {
  int N, nMT, nK16, rows_pad, Gb, depth, tap_first, n_taps;
  int relu, residual;
  long long G, plane;                   // groups, elements per stack plane
  int n_items;
  uint32_t a_bytes, gs, buf_bytes;      // operator image bytes, per-group buffer stride, per-plane buffer bytes
  uint32_t mt_off[kChebMaxMT], mt_rows[kChebMaxMT];
  const uint8_t* a_img;                 // prepared operator image (cheb_prep_kernel)
  const float* w;                       // [n_taps][16][16] fp32 (c_in, c_out)
  const float* bias;                    // [16] or nullptr
  // forward: in = x_0 plane (= stack plane 0), stack = [depth][G][N][16] (planes 1.. written), out = y
  // backward: in = dy, in2 = y, out = dx_0, out2 = dG
  const bf16* in; const bf16* in2; bf16* stack; bf16* out; bf16* out2;
  unsigned long long* dbg;              // optional timeline stamps (diagnostics)
};

// 16 bf16 of one 32-byte row in a 32B-swizzled buffer (rows 32 B apart; the two 16-byte halves swap when bit 2 of
// the row index is set)
__device__ __forceinline__ void row_store(uint8_t* buf, int row, const uint4& lo, const uint4& hi) {
  const int sw = (row >> 2) & 1;
  uint8_t* r = buf + row * 32;
  *reinterpret_cast<uint4*>(r + (sw << 4)) = lo;
  *reinterpret_cast<uint4*>(r + ((sw ^ 1) << 4)) = hi;
}
__device__ __forceinline__ void row_load_raw(const uint8_t* buf, int row, uint4& lo, uint4& hi) {
  const int sw = (row >> 2) & 1;
  const uint8_t* r = buf + row * 32;
  lo = *reinterpret_cast<const uint4*>(r + (sw << 4));
  hi = *reinterpret_cast<const uint4*>(r + ((sw ^ 1) << 4));
}
__device__ __forceinline__ void row_load(const uint8_t* buf, int row, float* v) {
  const int sw = (row >> 2) & 1;
  const uint8_t* r = buf + row * 32;
  const uint4 lo = *reinterpret_cast<const uint4*>(r + (sw << 4));
  const uint4 hi = *reinterpret_cast<const uint4*>(r + ((sw ^ 1) << 4));
  unpack8_bf16(lo, v);
  unpack8_bf16(hi, v + 8);
}

// cp.async fill of one plane buffer with Gb groups of `src` ([G][N][16] bf16); rows >= N and groups >= G are zeroed
__device__ __forceinline__ void cheb_fill(uint8_t* buf, const bf16* src, long long g0, const ChebParams& p, int tid, int nthr) {
  const int per_group = p.rows_pad * 2;
  const int total = p.Gb * per_group;
  for (int q = tid; q < total; q += nthr) {
    const int g = q / per_group, rem = q - g * per_group, row = rem >> 1, h = rem & 1;
    const bool ok = row < p.N && g0 + g < p.G;
    const bf16* s = src + ((ok ? (g0 + g) * p.N + row : 0LL) * kChebC + h * 8);
    cp_async16(buf + (size_t)g * p.gs + row * 32 + ((h ^ ((row >> 2) & 1)) << 4), s, ok ? 16u : 0u);
  }
}

// node contraction: acc[mt] (+)= A(mt) . buf   for every row tile; A = resident operator, B = plane buffer (MN-major)
__device__ __forceinline__ void cheb_issue_hop(uint32_t a_s, uint32_t buf, uint32_t d_tmem, int NC, uint32_t accumulate,
                                               const ChebParams& p) {
  const uint32_t idesc = make_idesc_bf16(128, NC, 0, 1);
  for (int mt = 0; mt < p.nMT; ++mt) {
    const uint32_t a_mt = a_s + p.mt_off[mt], blk = p.mt_rows[mt] * 128u;
    for (int ks = 0; ks < p.nK16; ++ks) {
      const uint64_t da = make_smem_desc(a_mt + (ks >> 2) * blk + (ks & 3) * 32, 16, 1024, SWZ_128B);
      const uint64_t db = make_smem_desc(buf + ks * 512, p.gs, 256, SWZ_32B);
      mma_bf16_ss(d_tmem + mt * NC, da, db, idesc, (ks != 0) ? 1u : accumulate);
    }
  }
}
// channel contraction: acc[mt][g] = buf[g][mt rows] . Wimg   (M = 128 vertices, N = 16, K = 16)
__device__ __forceinline__ void cheb_issue_mix(uint32_t buf, uint32_t w_img, uint32_t d_tmem, int NC, uint32_t accumulate,
                                               const ChebParams& p) {
  const uint32_t idesc = make_idesc_bf16(128, kChebC, 0, 0);
  const uint64_t db = make_smem_desc(w_img, 16, 256, SWZ_32B);
  for (int mt = 0; mt < p.nMT; ++mt)
    for (int g = 0; g < p.Gb; ++g) {
      const uint64_t da = make_smem_desc(buf + g * p.gs + mt * 4096, 16, 256, SWZ_32B);
      mma_bf16_ss(d_tmem + mt * NC + g * kChebC, da, db, idesc, accumulate);
    }
}

// Timeline stamps (diagnostics, stgcn_debug_timeline): CTA 0 records globaltimer for its first 3 items per slot at
// dbg[slot*72 + item*24 + e]: e = 0 fill begin, 1 fill landed, 2+j MMA stage j issued, 8+2j / 9+2j epilogue stage j
// begin / end; dbg[150] = kernel start, dbg[151] = operator resident.
#define CHEB_STAMP(e) do { if (dbg_on && it < 3) p.dbg[slot * 72 + it * 24 + (e)] = gtime(); } while (0)

template <bool BWD, int GB>
__global__ void __launch_bounds__(kChebThreads, 1) umma_cheb_kernel(ChebParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* a_s = smem;                                   // operator image
  uint8_t* w_s = smem + p.a_bytes;                       // n_taps x [16][16] bf16, 32B-swizzled K-major
  uint8_t* bufs0 = w_s + ((p.n_taps * 512 + 1023) & ~1023);   // per slot: depth plane buffers of GB groups
  __shared__ __align__(8) uint64_t afull, in_full_[2], in_free_[2], acc_full_[2], xk_ready_[2], mix_free_[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(16) float bias_s[kChebC];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int NC = GB * kChebC;
  constexpr int SB = GB < 2 ? GB : 2;          // groups drained per TMEM round trip
  const int regions = BWD ? p.depth : 2;
  const int slot_cols = regions * p.nMT * NC;
  uint32_t ncols = 32;
  while ((int)ncols < kChebSlots * slot_cols) ncols <<= 1;
  if (p.dbg && blockIdx.x == 0 && threadIdx.x == 0) p.dbg[150] = gtime();

  // weight images: forward B[n = c_out][k = c_in] = w[c_in][c_out]; backward B[n = c_in][k = c_out] = w[c_in][c_out]
  for (int i = threadIdx.x; i < p.n_taps * 256; i += blockDim.x) {
    const int tap = i >> 8, r = (i >> 4) & 15, kk = i & 15;
    const float v = BWD ? p.w[tap * 256 + r * 16 + kk] : p.w[tap * 256 + kk * 16 + r];
    const uint32_t off = tap * 512 + r * 32 + ((((kk >> 3) ^ ((r >> 2) & 1))) << 4) + (kk & 7) * 2;
    *reinterpret_cast<bf16*>(w_s + off) = __float2bfloat16_rn(v);
  }
  if (threadIdx.x < kChebC) bias_s[threadIdx.x] = p.bias ? p.bias[threadIdx.x] : 0.f;
  if (threadIdx.x == 0) {
    mbar_init(&afull, 1);
    for (int s = 0; s < kChebSlots; ++s) {
      mbar_init(&in_full_[s], kChebProducers);
      mbar_init(&in_free_[s], 4);
      mbar_init(&acc_full_[s], 1);
      mbar_init(&xk_ready_[s], 4);
      mbar_init(&mix_free_[s], 4);
    }
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(&tmem_base_s, ncols);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(&afull, p.a_bytes);
      for (uint32_t off = 0; off < p.a_bytes; off += 16384) {
        const uint32_t n = p.a_bytes - off < 16384 ? p.a_bytes - off : 16384;
        bulk_load_1d(a_s + off, p.a_img + off, n, &afull);
      }
    }
  } else {
    // ---- slot-local roles: warp 1 + 7*slot = MMA issuer, +1..+4 = epilogue, +5..+6 = cp.async producers ----
    const int slot = (warp - 1) / kChebSlotWarps, role = (warp - 1) % kChebSlotWarps;
    uint8_t* bufs = bufs0 + (size_t)slot * p.depth * p.buf_bytes;
    const uint32_t tmem_base = tmem_base_s + slot * slot_cols;
    uint64_t* in_full = &in_full_[slot]; uint64_t* in_free = &in_free_[slot]; uint64_t* acc_full = &acc_full_[slot];
    uint64_t* xk_ready = &xk_ready_[slot]; uint64_t* mix_free = &mix_free_[slot];
    const int item0 = blockIdx.x * kChebSlots + slot, item_step = gridDim.x * kChebSlots;
    const bool dbg_on = p.dbg != nullptr && blockIdx.x == 0 && lane == 0;

    if (role >= 5) {
      // =========================== producers ===============================
      const int ptid = (role - 5) * 32 + lane;
      uint32_t it = 0;
      for (int item = item0; item < p.n_items; item += item_step, ++it) {
        const long long g0 = (long long)item * GB;
        mbar_wait(in_free, (it & 1) ^ 1);
        if (role == 5) CHEB_STAMP(0);
        cheb_fill(bufs, p.in, g0, p, ptid, kChebProducers * 32);
        if (BWD && p.relu) cheb_fill(bufs + p.buf_bytes, p.in2, g0, p, ptid, kChebProducers * 32);
        cp_async_commit();
        cp_async_wait<0>();
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(in_full);
        if (role == 5) CHEB_STAMP(1);
      }
    } else if (role == 0) {
      // =========================== MMA issuer ==============================
      if (lane == 0) {
        const uint32_t a_u = smem_u32(a_s), w_u = smem_u32(w_s), b_u = smem_u32(bufs);
        mbar_wait(&afull, 0);
        if (dbg_on && slot == 0) p.dbg[151] = gtime();
        uint32_t it = 0, n_xk = 0;
        for (int item = item0; item < p.n_items; item += item_step, ++it) {
          if (!BWD) {
            mbar_wait(in_full, it & 1);
            tc_fence_after();
            for (int k = 1; k < p.depth; ++k) {
              if (k >= 2) { mbar_wait(xk_ready, n_xk & 1); ++n_xk; tc_fence_after(); }
              cheb_issue_hop(a_u, b_u + (k - 1) * p.buf_bytes, tmem_base, NC, 0, p);
              mma_commit(acc_full);
              CHEB_STAMP(2 + k - 1);
            }
            mbar_wait(xk_ready, n_xk & 1); ++n_xk;
            mbar_wait(mix_free, (it & 1) ^ 1);
            tc_fence_after();
            for (int t = 0; t < p.n_taps; ++t)
              cheb_issue_mix(b_u + (p.tap_first + t) * p.buf_bytes, w_u + t * 512, tmem_base + p.nMT * NC, NC, t != 0, p);
            mma_commit(acc_full);
            CHEB_STAMP(2 + p.depth - 1);
          } else {
            mbar_wait(xk_ready, n_xk & 1); ++n_xk;         // dG in buffer 0
            mbar_wait(mix_free, (it & 1) ^ 1);             // previous item's accumulators drained
            tc_fence_after();
            for (int t = 0; t < p.n_taps; ++t)
              cheb_issue_mix(b_u, w_u + t * 512, tmem_base + (p.tap_first + t) * p.nMT * NC, NC, 0, p);
            mma_commit(acc_full);
            CHEB_STAMP(2);
            for (int k = p.depth - 1; k >= 1; --k) {
              mbar_wait(xk_ready, n_xk & 1); ++n_xk;       // P_k in buffer k
              tc_fence_after();
              const uint32_t has_product = (k - 1 >= p.tap_first && k - 1 < p.tap_first + p.n_taps) ? 1u : 0u;
              cheb_issue_hop(a_u, b_u + k * p.buf_bytes, tmem_base + (k - 1) * p.nMT * NC, NC, has_product, p);
              mma_commit(acc_full);
              CHEB_STAMP(2 + p.depth - k);
            }
          }
        }
      }
    } else {
      // =========================== epilogue warps ==========================
      const int q = warp & 3;
      const int r = q * 32 + lane;
      const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
      const bool stamp = role == 1;
      uint32_t it = 0, n_acc = 0;
      for (int item = item0; item < p.n_items; item += item_step, ++it) {
        const long long g0 = (long long)item * GB;
        if (!BWD) {
          for (int k = 1; k < p.depth; ++k) {
            const float alpha = k == 1 ? 1.f : 2.f;
            uint8_t* bk = bufs + (size_t)k * p.buf_bytes;
            const uint8_t* bm2 = bufs + (size_t)(k >= 2 ? k - 2 : 0) * p.buf_bytes;
            bf16* plane = p.stack + (size_t)k * p.plane;
            mbar_wait(acc_full, n_acc & 1); ++n_acc;
            tc_fence_after();
            if (stamp) CHEB_STAMP(8 + 2 * (k - 1));
            for (int mt = 0; mt < p.nMT; ++mt) {
              if (mt * 128 + q * 32 >= p.rows_pad) break;          // warp-uniform (tcgen05.ld is .sync.aligned)
              const int n = mt * 128 + r;
              const bool nvalid = n < p.N, inbuf = n < p.rows_pad;
#pragma unroll 1
              for (int gb = 0; gb < GB; gb += SB) {
              uint32_t a[SB][16];
              uint4 m2[SB][2];
#pragma unroll
              for (int j = 0; j < SB; ++j) tmem_ld_32x32b_x16(t_lane + mt * NC + (gb + j) * kChebC, a[j]);
#pragma unroll
              for (int j = 0; j < SB; ++j)
                if (k >= 2 && inbuf) row_load_raw(bm2 + (size_t)(gb + j) * p.gs, n, m2[j][0], m2[j][1]);
              tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < SB; ++j) {
                const int g = gb + j;
                float v[16], m[16];
                if (k >= 2) { unpack8_bf16(m2[j][0], m); unpack8_bf16(m2[j][1], m + 8); }
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                  v[i] = alpha * __uint_as_float(a[j][i]);
                  if (k >= 2) v[i] -= m[i];
                  if (!nvalid) v[i] = 0.f;
                }
                const uint4 lo = pack8_bf16(v), hi = pack8_bf16(v + 8);
                if (inbuf) row_store(bk + (size_t)g * p.gs, n, lo, hi);
                if (nvalid && g0 + g < p.G) {
                  uint4* dst = reinterpret_cast<uint4*>(plane + ((g0 + g) * p.N + n) * kChebC);
                  dst[0] = lo; dst[1] = hi;
                }
              }
              }
            }
            fence_proxy_async();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(xk_ready);
            if (stamp) CHEB_STAMP(9 + 2 * (k - 1));
          }
          // y = relu(acc + bias + x_0)
          mbar_wait(acc_full, n_acc & 1); ++n_acc;
          tc_fence_after();
          if (stamp) CHEB_STAMP(8 + 2 * (p.depth - 1));
          for (int mt = 0; mt < p.nMT; ++mt) {
            if (mt * 128 + q * 32 >= p.N) break;                   // warp-uniform
            const int n = mt * 128 + r;
            const bool nvalid = n < p.N;
#pragma unroll 1
            for (int gb = 0; gb < GB; gb += SB) {
            uint32_t a[SB][16];
            uint4 x0[SB][2];
#pragma unroll
            for (int j = 0; j < SB; ++j) tmem_ld_32x32b_x16(t_lane + (p.nMT + mt) * NC + (gb + j) * kChebC, a[j]);
#pragma unroll
            for (int j = 0; j < SB; ++j)
              if (p.residual && nvalid) row_load_raw(bufs + (size_t)(gb + j) * p.gs, n, x0[j][0], x0[j][1]);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < SB; ++j) {
              const int g = gb + j;
              float v[16], m[16];
              if (p.residual) { unpack8_bf16(x0[j][0], m); unpack8_bf16(x0[j][1], m + 8); }
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                v[i] = __uint_as_float(a[j][i]) + bias_s[i];
                if (p.residual) v[i] += m[i];
                if (p.relu) v[i] = fmaxf(v[i], 0.f);
              }
              if (nvalid && g0 + g < p.G) {
                uint4* dst = reinterpret_cast<uint4*>(p.out + ((g0 + g) * p.N + n) * kChebC);
                dst[0] = pack8_bf16(v); dst[1] = pack8_bf16(v + 8);
              }
            }
            }
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) { mbar_arrive(mix_free); mbar_arrive(in_free); }
          if (stamp) CHEB_STAMP(9 + 2 * (p.depth - 1));
        } else {
          // S0: dG = dy * [y > 0] in place in buffer 0 (+ HBM copy for the weight-gradient kernels)
          mbar_wait(in_full, it & 1);
          for (int mt = 0; mt < p.nMT; ++mt) {
            const int n = mt * 128 + r;
            if (n >= p.N) break;                         // padded rows were zero-filled by the producers
#pragma unroll
            for (int g = 0; g < GB; ++g) {
              if (g0 + g >= p.G) break;
              float dy[16], y[16];
              row_load(bufs + (size_t)g * p.gs, n, dy);
              if (p.relu) {
                row_load(bufs + p.buf_bytes + (size_t)g * p.gs, n, y);
#pragma unroll
                for (int i = 0; i < 16; ++i) dy[i] = y[i] > 0.f ? dy[i] : 0.f;
              }
              const uint4 lo = pack8_bf16(dy), hi = pack8_bf16(dy + 8);
              if (p.relu) row_store(bufs + (size_t)g * p.gs, n, lo, hi);
              uint4* dst = reinterpret_cast<uint4*>(p.out2 + ((g0 + g) * p.N + n) * kChebC);
              dst[0] = lo; dst[1] = hi;
            }
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) mbar_arrive(xk_ready);
          // E_k, k = depth-1 .. 1: P_k = alpha_k * (R_k - D_{k+2}) -> buffer k ;  E_0: dx_0 = R_0 - D_2 + dG
          for (int k = p.depth - 1; k >= 0; --k) {
            const float alpha = k >= 2 ? 2.f : 1.f;
            const bool sub = k + 2 <= p.depth - 1;
            const bool addg = k == 0 && p.residual;
            uint8_t* bk = bufs + (size_t)k * p.buf_bytes;
            const uint8_t* bp2 = bufs + (size_t)(sub ? k + 2 : 0) * p.buf_bytes;
            mbar_wait(acc_full, n_acc & 1); ++n_acc;
            tc_fence_after();
            if (stamp) CHEB_STAMP(8 + 2 * (p.depth - 1 - k));
            for (int mt = 0; mt < p.nMT; ++mt) {
              if (mt * 128 + q * 32 >= p.rows_pad) break;          // warp-uniform
              const int n = mt * 128 + r;
              const bool nvalid = n < p.N, inbuf = n < p.rows_pad;
#pragma unroll 1
              for (int gb = 0; gb < GB; gb += SB) {
              uint32_t a[SB][16];
              uint4 p2[SB][2], dg[SB][2];
#pragma unroll
              for (int j = 0; j < SB; ++j) tmem_ld_32x32b_x16(t_lane + (k * p.nMT + mt) * NC + (gb + j) * kChebC, a[j]);
#pragma unroll
              for (int j = 0; j < SB; ++j) {
                if (sub && inbuf) row_load_raw(bp2 + (size_t)(gb + j) * p.gs, n, p2[j][0], p2[j][1]);
                if (addg && inbuf) row_load_raw(bufs + (size_t)(gb + j) * p.gs, n, dg[j][0], dg[j][1]);
              }
              tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < SB; ++j) {
                const int g = gb + j;
                float v[16], m[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(a[j][i]);
                if (sub) {
                  unpack8_bf16(p2[j][0], m); unpack8_bf16(p2[j][1], m + 8);
#pragma unroll
                  for (int i = 0; i < 16; ++i) v[i] -= 0.5f * m[i];     // buffer k+2 holds 2 * D_{k+2}
                }
                if (addg) {
                  unpack8_bf16(dg[j][0], m); unpack8_bf16(dg[j][1], m + 8);
#pragma unroll
                  for (int i = 0; i < 16; ++i) v[i] += m[i];
                }
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = nvalid ? alpha * v[i] : 0.f;
                const uint4 lo = pack8_bf16(v), hi = pack8_bf16(v + 8);
                if (k > 0) {
                  if (inbuf) row_store(bk + (size_t)g * p.gs, n, lo, hi);
                } else if (nvalid && g0 + g < p.G) {
                  uint4* dst = reinterpret_cast<uint4*>(p.out + ((g0 + g) * p.N + n) * kChebC);
                  dst[0] = lo; dst[1] = hi;
                }
              }
              }
            }
            tc_fence_before();
            if (k > 0) {
              fence_proxy_async();
              __syncwarp();
              if (lane == 0) mbar_arrive(xk_ready);
            } else {
              __syncwarp();
              if (lane == 0) { mbar_arrive(mix_free); mbar_arrive(in_free); }
            }
            if (stamp) CHEB_STAMP(9 + 2 * (p.depth - 1 - k));
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base_s, ncols);
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
struct ChebPlan {
  bool ok; int nMT, nKB, nK16, rows_pad, Gb;
  uint32_t a_bytes, gs, buf_bytes, w_bytes, mt_off[kChebMaxMT], mt_rows[kChebMaxMT];
  size_t smem;
};

// depth = number of plane buffers (Ks for Chebyshev, 2 for GraphConv); bwd: TMEM holds `depth` regions instead of 2
inline ChebPlan plan_cheb(int N, int C, int depth, int n_taps, bool bwd) {
  ChebPlan pl{};
  pl.ok = false;
  if (C != kChebC || depth < 2 || depth > kChebMaxDepth || N < 1) return pl;
  pl.nMT = (N + 127) / 128;
  if (pl.nMT > kChebMaxMT) return pl;
  pl.nKB = (N + 63) / 64;
  pl.nK16 = (N + 15) / 16;
  pl.rows_pad = pl.nK16 * 16;
  uint32_t off = 0;
  for (int mt = 0; mt < pl.nMT; ++mt) {
    int rows = N - mt * 128;
    rows = rows > 128 ? 128 : (rows + 7) / 8 * 8;
    pl.mt_off[mt] = off; pl.mt_rows[mt] = (uint32_t)rows;
    off += (uint32_t)rows * 128u * pl.nKB;
  }
  pl.a_bytes = off;
  pl.gs = (uint32_t)pl.rows_pad * 32u;
  pl.w_bytes = (uint32_t)((n_taps * 512 + 1023) & ~1023);
  const int regions = bwd ? depth : 2;
  const size_t slack = 1024 + 4096;       // base alignment + operand over-read past the last buffer (never stored)
  for (int Gb = 4; Gb >= 1; Gb /= 2) {
    if (kChebSlots * regions * pl.nMT * Gb * kChebC > 512) continue;
    const size_t need = (size_t)pl.a_bytes + pl.w_bytes + (size_t)kChebSlots * depth * Gb * pl.gs + slack;
    if (need > kSmemBudget) continue;
    pl.Gb = Gb; pl.buf_bytes = (uint32_t)Gb * pl.gs; pl.smem = need;
    pl.ok = true;
    return pl;
  }
  return pl;
}
inline bool cheb_supported(int N, int C, int depth, int n_taps, long long G) {
  return G > 0 && plan_cheb(N, C, depth, n_taps, false).ok && plan_cheb(N, C, depth, n_taps, true).ok;
}
inline size_t cheb_image_bytes(int N) {      // upper bound of the operator image (plan independent)
  return (size_t)((N + 127) / 128) * 128 * 128 * ((N + 63) / 64);
}

// Lhat (fp32 [N,N]) -> shared-memory image: per row tile mt, per 64-column block kb, [rows][128 B] with the 128B
// swizzle (16-byte chunk index ^= row & 7); zero padded; optionally transposed
struct ChebImg { int N, nMT, nKB; uint32_t mt_off[kChebMaxMT], mt_rows[kChebMaxMT]; };
__global__ void cheb_prep_kernel(const float* M, uint8_t* img, ChebImg im, int trans, uint32_t total_elems) {
  uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total_elems) return;
  const uint32_t byte = idx * 2;
  int mt = 0;
  while (mt + 1 < im.nMT && byte >= im.mt_off[mt + 1]) ++mt;
  const uint32_t local = byte - im.mt_off[mt];
  const uint32_t blk = im.mt_rows[mt] * 128u;
  const int kb = (int)(local / blk);
  const uint32_t inb = local - (uint32_t)kb * blk;
  const int r = (int)(inb >> 7), chunk = (int)((inb >> 4) & 7), e = (int)((inb & 15) >> 1);
  const int kk = ((chunk ^ (r & 7)) << 3) + e;
  const int row = mt * 128 + r, col = kb * 64 + kk;
  float v = 0.f;
  if (row < im.N && col < im.N) v = trans ? M[(long long)col * im.N + row] : M[(long long)row * im.N + col];
  reinterpret_cast<bf16*>(img)[idx] = __float2bfloat16_rn(v);
}
inline void launch_cheb_prep(const float* M, void* img, int N, int trans, cudaStream_t stream) {
  ChebPlan pl = plan_cheb(N, kChebC, 2, 1, false);
  STGCN_CHECK(pl.nMT >= 1 && pl.nMT <= kChebMaxMT, STGCN_E_UNSUPPORTED, "cheb prep: too many row tiles");
  ChebImg im{};
  im.N = N; im.nMT = pl.nMT; im.nKB = pl.nKB;
  for (int i = 0; i < pl.nMT; ++i) { im.mt_off[i] = pl.mt_off[i]; im.mt_rows[i] = pl.mt_rows[i]; }
  const uint32_t total = pl.a_bytes / 2;
  STGCN_LAUNCH(cheb_prep_kernel, ceil_div(total, 256), 256, 0, stream, M, reinterpret_cast<uint8_t*>(img), im, trans, total);
}

struct ChebProblem {
  int N; long long G; int depth, tap_first, n_taps, relu, residual;
  const void* a_img; const float* w; const float* bias;
  const bf16* in; const bf16* in2; bf16* stack; bf16* out; bf16* out2;
};
inline void launch_cheb(const ChebProblem& q, bool bwd, cudaStream_t stream) {
  ChebPlan pl = plan_cheb(q.N, kChebC, q.depth, q.n_taps, bwd);
  STGCN_CHECK(pl.ok, STGCN_E_UNSUPPORTED, "umma cheb: unsupported shape");
  ChebParams p{};
  p.N = q.N; p.nMT = pl.nMT; p.nK16 = pl.nK16; p.rows_pad = pl.rows_pad; p.Gb = pl.Gb; p.depth = q.depth;
  p.tap_first = q.tap_first; p.n_taps = q.n_taps; p.relu = q.relu; p.residual = q.residual;
  p.G = q.G; p.plane = q.G * q.N * kChebC;
  p.n_items = (int)((q.G + pl.Gb - 1) / pl.Gb);
  p.a_bytes = pl.a_bytes; p.gs = pl.gs; p.buf_bytes = pl.buf_bytes;
  for (int i = 0; i < pl.nMT; ++i) { p.mt_off[i] = pl.mt_off[i]; p.mt_rows[i] = pl.mt_rows[i]; }
  p.a_img = reinterpret_cast<const uint8_t*>(q.a_img); p.w = q.w; p.bias = q.bias;
  p.in = q.in; p.in2 = q.in2; p.stack = q.stack; p.out = q.out; p.out2 = q.out2;
  p.dbg = g_tap_dbg;
  const int pairs = (p.n_items + kChebSlots - 1) / kChebSlots;
  int gx = pairs < sm_count() ? pairs : sm_count();
  if (gx < 1) gx = 1;
  auto go = [&](auto kern, const char* name) {
    STGCN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
    STGCN_LAUNCH_NAMED(name, kern, gx, kChebThreads, pl.smem, stream, p);
  };
  if (bwd) {
    const char* nm = "umma_cheb_kernel<bwd>";
    if (pl.Gb == 4) go(umma_cheb_kernel<true, 4>, nm); else if (pl.Gb == 2) go(umma_cheb_kernel<true, 2>, nm); else go(umma_cheb_kernel<true, 1>, nm);
  } else {
    const char* nm = "umma_cheb_kernel<fwd>";
    if (pl.Gb == 4) go(umma_cheb_kernel<false, 4>, nm); else if (pl.Gb == 2) go(umma_cheb_kernel<false, 2>, nm); else go(umma_cheb_kernel<false, 1>, nm);
  }
}

}  // namespace umma
}  // namespace stgcn
