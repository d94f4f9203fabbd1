// umma_x3.cuh -- the tensor-core PARITY mode (precision == STGCN_PREC_TF32X3): fp32 activations in HBM, every GEMM of
// the fp32 chain on tcgen05 with "3xTF32" operand splitting, fp32 accumulation in TMEM.
//
//   x = x_hi + x_lo,  x_hi = tf32(x) (11 significant bits),  x_lo = tf32(x - x_hi)         (22 bits in total)
//   A.B ~= A_lo.B_hi + A_hi.B_lo + A_hi.B_hi      (three tcgen05.mma kind::tf32 per product; A_lo.B_lo ~ 2^-22 dropped)
//
// Measured on the reference's golden vectors by operand-rounding emulation (oracle + autograd, all 14 cases): one-pass
// bf16 0.475 / one-pass tf32 0.96 / split bf16x2 5.3e-3 / split tf32x2 1.7e-5 worst per-tensor gradient rel-L2, so the
// north-star gate (1e-3) needs >= 22 operand bits; bf16 would need three pieces and six products for the same.
//
// One persistent warp-specialised kernel serves the three GEMM shapes of the fp32 chain (simt_kernels.cuh has their
// CUDA-core twins, which remain the fallback for shapes this kernel does not take):
//   X3_TAP    out[r, o]   = bias[o] + sum_{tap, c} in[row(r, tap), c] * wt[(tap, c), o]           (tapgemm_kernel)
//   X3_GSO    out[g, h,:] = alpha * sum_i L[h, i] x[g, i, :] + beta * aux[g, h, :]                 (gso_kernel)
//   X3_WGRAD  dwt[(tap, c) | bias][o] += sum_r in[row(r, tap), c] * dz[r, o]     (as D^T: M = o, N = (tap, c))  (wgrad_kernel)
// all as  D[128 x BN] += A[128 x 32] . B[BN x 32]^T  per K chunk of 32 fp32 (= one 128-byte swizzle row).
//
// Operands come straight from the fp32 tensors: 8 producer warps load them (coalesced along whichever axis is contiguous
// in memory), split every value and write the hi / lo tiles K-major with the 128-byte swizzle -- the operand layout the
// bf16 tap kernel uses for 64-channel inputs (validated there; only the instruction kind and element size differ).
// Warp roles: 0-3 epilogue (TMEM lane quarter = warp index), 4 MMA issuer + TMEM allocation, 5-12 producers (two groups
// of four warps on alternate pipeline stages).
#pragma once
#include "umma.cuh"
#include "simt_kernels.cuh"
#include "umma_tap.cuh"        // kSmemBudget, sm_count()

namespace stgcn {
namespace umma {

enum { X3_TAP = 0, X3_GSO = 1, X3_WGRAD = 2 };
constexpr int kX3EpiWarps = 4, kX3ProdWarps = 8;
constexpr int kX3ProdThreads = 32 * kX3ProdWarps;
constexpr int kX3GroupThreads = kX3ProdThreads / 2;      // producers work as two groups on alternate stages
constexpr int kX3Threads = 32 * (kX3EpiWarps + 1) + kX3ProdThreads;      // 416
constexpr int kX3KC = 32;               // K elements per stage
constexpr int kX3MaxStages = 4;
constexpr uint32_t kX3ATile = 128u * 128u;      // bytes of one [128 rows x 32 fp32] operand tile

// instruction descriptor, kind::tf32: D = f32 (bits [4,6) = 1), A and B = tf32 (format 2), both K-major
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_tf32_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

struct X3Params {
  int BN, S;                        // tile width (16..256, power of two), pipeline stages
  int m_tiles, n_tiles, k_splits;
  long long Ktot, k_per_split;      // contraction length; per-split share (multiple of 32)
  uint32_t stage_bytes;
  // X3_TAP / X3_WGRAD
  const float* in;                  // [*, Cin]
  const float* wt;                  // TAP: [ntaps*Cin, Co]
  const float* bias;                // TAP: [Co] or nullptr
  float* out;                       // TAP: [rows, ldo];  GSO: [G, N, C]
  long long rows;
  int Cin, Co, ntaps, ldo, accumulate;
  simt::RowMap map;
  // X3_GSO (x = in)
  const float* L;                   // [N, N]
  const float* aux;
  int trans, N, C;
  long long G;
  float alpha, beta;
  // X3_WGRAD
  const float* dz;                  // [rows, ldz]
  float* dst;                       // partial [k_splits][Mw][Co], or dwt itself (atomics) when k_splits == 1 without scratch
  int ldz, bias_row, Kw, Mw, atomic;
  int vec;                          // TAP / GSO epilogue: 16-byte stores are legal (alignment and strides)
};

// split 4 values and store the hi / lo 16-byte chunks of row `row`, chunk `chunk` of a 128B-swizzled K-major tile pair
__device__ __forceinline__ void x3_store_chunk(uint32_t hi_tile, uint32_t lo_tile, int row, int chunk, const float* v) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint32_t hb = (__float_as_uint(v[i]) + 0x1000u) & 0xFFFFE000u;      // round to 10 mantissa bits
    const float lo = v[i] - __uint_as_float(hb);                               // exact in fp32
    h[i] = hb;
    l[i] = (__float_as_uint(lo) + 0x1000u) & 0xFFFFE000u;
  }
  const uint32_t off = (uint32_t)row * 128u + (uint32_t)((chunk ^ (row & 7)) << 4);
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(hi_tile + off), "r"(h[0]), "r"(h[1]), "r"(h[2]), "r"(h[3]) : "memory");
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(lo_tile + off), "r"(l[0]), "r"(l[1]), "r"(l[2]), "r"(l[3]) : "memory");
}

// All index arithmetic of the pipeline is 32-bit: the host guarantees tiles, rows and K below 2^31, and 64-bit integer
// division is a ~150-instruction subroutine on the GPU -- with it in the producers' per-chunk path the first version of
// this kernel was instruction bound at 3 us per stage (profiles/r02_ab_batch_b.md).
struct X3Tile { int mt, nt, ks, kbeg, kend, nkc; };
__device__ __forceinline__ X3Tile x3_tile(const X3Params& p, unsigned tile) {
  X3Tile t;
  const unsigned rest = tile / (unsigned)p.k_splits;
  t.ks = (int)(tile - rest * (unsigned)p.k_splits);
  t.mt = (int)(rest / (unsigned)p.n_tiles);
  t.nt = (int)(rest - (unsigned)t.mt * (unsigned)p.n_tiles);
  t.kbeg = t.ks * (int)p.k_per_split;
  const long long ke = (long long)t.kbeg + p.k_per_split;
  t.kend = (int)(ke < p.Ktot ? ke : p.Ktot);
  t.nkc = (t.kend - t.kbeg + kX3KC - 1) / kX3KC;
  return t;
}

template <int MODE>
__global__ void __launch_bounds__(kX3Threads, 1) umma_x3_kernel(X3Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ __align__(8) uint64_t full[kX3MaxStages], empty[kX3MaxStages], tfull[2], tempty[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ long long dec_base[kX3MaxStages][kX3KC];      // X3_WGRAD: input row of tap 0 per K row of the stage (-1: past the end)
  __shared__ int dec_t[kX3MaxStages][kX3KC];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t ncols = 32;
  while ((int)ncols < 2 * p.BN) ncols <<= 1;
  if (threadIdx.x == 0) {
    for (int s = 0; s < p.S; ++s) { mbar_init(&full[s], kX3ProdWarps / 2); mbar_init(&empty[s], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], kX3EpiWarps); }
    fence_barrier_init();
  }
  if (warp == kX3EpiWarps) tmem_alloc(&tmem_base_s, ncols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  const unsigned n_tiles = (unsigned)p.m_tiles * (unsigned)p.n_tiles * (unsigned)p.k_splits;
  const uint32_t b_off = 2u * kX3ATile, b_tile = (uint32_t)p.BN * 128u;

  if (warp > kX3EpiWarps) {
    // =========================== producers ================================
    // Two groups of 4 warps take alternate stages, so two stages' global loads are always in flight (one group alone
    // exposed a full HBM round trip per stage: 3 us per 128x128x32 stage, profiles/r02_ab_batch_a.md).  Within a stage a
    // thread issues ALL its loads first, then splits and stores.
    const int tp = threadIdx.x - 32 * (kX3EpiWarps + 1);
    const int grp = tp >> 7, tq = tp & 127;
    const long long TN_out = (long long)p.map.T_out * p.map.N, TN_in = (long long)p.map.T_in * p.map.N;
    const long long tap_step = (long long)p.map.t_shift * p.map.N + p.map.tap_row_stride;
    const bool wide = p.BN == 256;                       // two B rows per thread
    const int NB = p.BN / 16;                            // B tasks (16-byte chunks) per thread and stage
    const int brow0 = wide ? tq : tq % p.BN, bchunk0 = wide ? 0 : tq / p.BN, bstep = wide ? 1 : kX3GroupThreads / p.BN;
    uint32_t g = 0;
    for (unsigned tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const X3Tile t = x3_tile(p, tile);
      const int m0 = t.mt * 128, n0 = t.nt * p.BN;
      // ---- per-tile row contexts
      long long a_base[8];
      int a_t[8];
      long long b_base[2] = {-1, -1};     // GSO: offset of column J in x
      int b_tap[2] = {-2, -2}, b_c[2] = {0, 0};            // WGRAD: (tap, c) of row mm; -1 = bias row; -2 = padding
      if (MODE == X3_TAP) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const long long r = (long long)m0 + (tq >> 3) + 16 * j;
          a_base[j] = 0;
          if (r < p.rows) simt::row_decode(r, (int)TN_out, p.map.N, TN_in, a_base[j], a_t[j]);
          else a_t[j] = -(1 << 24);
        }
      } else if (MODE == X3_GSO) {
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const long long J = (long long)n0 + brow0 + 128 * q;
          if ((q == 0 || wide) && J < p.G * p.C) { const long long gg = J / p.C; b_base[q] = gg * p.N * p.C + (J - gg * p.C); }
        }
      } else {
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const int mm = n0 + brow0 + 128 * q;
          if (q == 0 || wide) {
            if (mm < p.Kw) { b_tap[q] = mm / p.Cin; b_c[q] = mm - b_tap[q] * p.Cin; }
            else if (mm == p.Kw && p.bias_row) b_tap[q] = -1;
          }
        }
      }
      for (int kc = 0; kc < t.nkc; ++kc, ++g) {
        if ((int)(g & 1) != grp) continue;
        const uint32_t s = g % p.S, ph = (g / p.S) & 1;
        mbar_wait(&empty[s], ph ^ 1);
        const uint32_t st = smem_u32(smem + (size_t)s * p.stage_bytes);
        const uint32_t a_hi = st, a_lo = st + kX3ATile, b_hi = st + b_off, b_lo = st + b_off + b_tile;
        const int k0 = t.kbeg + kc * kX3KC;
        if (MODE == X3_WGRAD) {
          if (tq < kX3KC) {
            const int r = k0 + tq;
            long long base = -1; int tt = 0;
            if (r < t.kend) simt::row_decode(r, (int)TN_out, p.map.N, TN_in, base, tt);
            dec_base[s][tq] = base; dec_t[s][tq] = tt;
          }
          named_bar_sync(3 + grp, kX3GroupThreads);
        }
        const bool a_kcontig = MODE == X3_TAP || (MODE == X3_GSO && !p.trans);
        // ---- loads: A tile (8 chunks per thread)
        float va[8][4];
        const int ka = k0 + 4 * (tq & 7);                  // K-contiguous sources: this thread's chunk of every row
        int a_tap = 0, a_c = 0;
        if (MODE == X3_TAP) { a_tap = (int)((unsigned)ka / (unsigned)p.Cin); a_c = ka - a_tap * p.Cin; }
        const int Kend = (int)p.Ktot;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          va[j][0] = va[j][1] = va[j][2] = va[j][3] = 0.f;
          if (a_kcontig) {          // source contiguous along K: 8 lanes cover one row's 128 bytes
            const int row = (tq >> 3) + 16 * j;
            const int gk = ka;
            if (MODE == X3_TAP) {
              if (gk < Kend) {
                const int tap = a_tap, c = a_c;
                const int ti = a_t[j] + p.map.t_shift * tap;
                if (ti >= 0 && ti < p.map.T_in) {
                  const float4 q = __ldg(reinterpret_cast<const float4*>(p.in + (a_base[j] + tap * tap_step) * p.Cin + c));
                  va[j][0] = q.x; va[j][1] = q.y; va[j][2] = q.z; va[j][3] = q.w;
                }
              }
            } else {
              const int h = m0 + row;
              if (h < p.N) {
#pragma unroll
                for (int i = 0; i < 4; ++i)
                  if (gk + i < p.N) va[j][i] = __ldg(p.L + (long long)h * p.N + (gk + i));
              }
            }
          } else {                  // source contiguous along M: one lane per row, four K steps by four coalesced loads
            const int row = tq, chunk = j;
            const int gk = k0 + 4 * chunk;
            if (MODE == X3_GSO) {
              const int h = m0 + row;
              if (h < p.N) {
#pragma unroll
                for (int i = 0; i < 4; ++i)
                  if (gk + i < p.N) va[j][i] = __ldg(p.L + (long long)(gk + i) * p.N + h);
              }
            } else {                // X3_WGRAD: A(o, r) = dz[r, o]
              const int o = m0 + row;
              if (o < p.Co) {
#pragma unroll
                for (int i = 0; i < 4; ++i)
                  if (gk + i < t.kend) va[j][i] = __ldg(p.dz + (long long)(gk + i) * p.ldz + o);
              }
            }
          }
        }
        // ---- loads: B tile (always contiguous along N in memory), in halves of up to 8 chunks per thread
        auto load_b = [&](int u, float* v) {
          v[0] = v[1] = v[2] = v[3] = 0.f;
          const int q = wide ? (u & 1) : 0;
          const int chunk = wide ? (u >> 1) : bchunk0 + u * bstep;
          if (chunk >= 8) return;
          const int gk = k0 + 4 * chunk;
          if (MODE == X3_TAP) {
            const int o = n0 + brow0 + 128 * q;
            if (o < p.Co) {
#pragma unroll
              for (int i = 0; i < 4; ++i)
                if (gk + i < Kend) v[i] = __ldg(p.wt + (long long)(gk + i) * p.Co + o);
            }
          } else if (MODE == X3_GSO) {
            if (b_base[q] >= 0) {
#pragma unroll
              for (int i = 0; i < 4; ++i)
                if (gk + i < p.N) v[i] = __ldg(p.in + b_base[q] + (long long)(gk + i) * p.C);
            }
          } else {
            if (b_tap[q] >= -1) {
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const long long rb = dec_base[s][4 * chunk + i];
                if (rb >= 0) {
                  if (b_tap[q] < 0) v[i] = 1.f;
                  else {
                    const int ti = dec_t[s][4 * chunk + i] + p.map.t_shift * b_tap[q];
                    if (ti >= 0 && ti < p.map.T_in) v[i] = __ldg(p.in + (rb + b_tap[q] * tap_step) * p.Cin + b_c[q]);
                  }
                }
              }
            }
          }
        };
        auto store_b = [&](int u, const float* v) {
          const int q = wide ? (u & 1) : 0;
          const int chunk = wide ? (u >> 1) : bchunk0 + u * bstep;
          if (chunk < 8) x3_store_chunk(b_hi, b_lo, brow0 + 128 * q, chunk, v);
        };
        float vb[8][4];
#pragma unroll
        for (int u = 0; u < 8; ++u)
          if (u < NB) load_b(u, vb[u]);
        // ---- split + store
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (a_kcontig) x3_store_chunk(a_hi, a_lo, (tq >> 3) + 16 * j, tq & 7, va[j]);
          else x3_store_chunk(a_hi, a_lo, tq, j, va[j]);
        }
#pragma unroll
        for (int u = 0; u < 8; ++u)
          if (u < NB) store_b(u, vb[u]);
        if (NB > 8) {
#pragma unroll
          for (int u = 0; u < 8; ++u) load_b(8 + u, vb[u]);
#pragma unroll
          for (int u = 0; u < 8; ++u) store_b(8 + u, vb[u]);
        }
        fence_proxy_async();              // generic-proxy stores -> visible to the tensor core (async proxy)
        __syncwarp();
        if (lane == 0) mbar_arrive(&full[s]);
      }
    }
  } else if (warp == kX3EpiWarps) {
    // =========================== MMA issuer =============================
    if (elect_one()) {
      const uint32_t idesc = make_idesc_tf32(128, p.BN);
      const uint64_t dproto = make_smem_desc(0, 16, 1024, SWZ_128B);
      uint32_t g = 0, acc_cnt = 0;
      for (unsigned tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++acc_cnt) {
        const X3Tile t = x3_tile(p, tile);
        const uint32_t ab = acc_cnt & 1, aph = (acc_cnt >> 1) & 1;
        mbar_wait(&tempty[ab], aph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + ab * p.BN;
        uint32_t accumulate = 0;
        for (int kc = 0; kc < t.nkc; ++kc, ++g) {
          const uint32_t s = g % p.S, ph = (g / p.S) & 1;
          mbar_wait(&full[s], ph);
          tc_fence_after();
          const uint32_t st = smem_u32(smem + (size_t)s * p.stage_bytes);
          uint64_t ah = desc_at(dproto, st), al = desc_at(dproto, st + kX3ATile);
          uint64_t bh = desc_at(dproto, st + b_off), bl = desc_at(dproto, st + b_off + b_tile);
#pragma unroll
          for (int k = 0; k < kX3KC / 8; ++k) {         // K = 8 per kind::tf32 instruction = 32 bytes
            mma_tf32_ss(d_tmem, al, bh, idesc, accumulate);
            mma_tf32_ss(d_tmem, ah, bl, idesc, 1);
            mma_tf32_ss(d_tmem, ah, bh, idesc, 1);
            accumulate = 1;
            ah += 2; al += 2; bh += 2; bl += 2;
          }
          mma_commit(&empty[s]);
        }
        mma_commit(&tfull[ab]);
      }
    }
  } else {
    // =========================== epilogue warps ==========================
    const int row = warp * 32 + lane;           // TMEM lane = tile row
    uint32_t acc_cnt = 0;
    for (unsigned tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++acc_cnt) {
      const X3Tile t = x3_tile(p, tile);
      const int m0 = t.mt * 128, n0 = t.nt * p.BN;
      const uint32_t ab = acc_cnt & 1, aph = (acc_cnt >> 1) & 1;
      mbar_wait(&tfull[ab], aph);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + ((uint32_t)(warp * 32) << 16) + ab * p.BN;
#pragma unroll 1
      for (int c0 = 0; c0 < p.BN; c0 += 16) {
        uint32_t rr[16];
        tmem_ld_32x32b_x16(t_addr + c0, rr);
        tmem_ld_wait();
        float v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(rr[i]);
        if (MODE == X3_TAP) {
          const long long r = (long long)m0 + row;
          if (r < p.rows) {
            float* orow = p.out + r * p.ldo;
            if (p.vec && n0 + c0 + 16 <= p.Co) {          // 64 contiguous bytes per thread
#pragma unroll
              for (int i = 0; i < 16; i += 4) {
                const int o = n0 + c0 + i;
                float4 x = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                if (p.bias) { x.x += __ldg(p.bias + o); x.y += __ldg(p.bias + o + 1); x.z += __ldg(p.bias + o + 2); x.w += __ldg(p.bias + o + 3); }
                if (p.accumulate) { const float4 y = *reinterpret_cast<const float4*>(orow + o); x.x += y.x; x.y += y.y; x.z += y.z; x.w += y.w; }
                *reinterpret_cast<float4*>(orow + o) = x;
              }
            } else {
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                const int o = n0 + c0 + i;
                if (o < p.Co) {
                  float x = v[i] + (p.bias ? __ldg(p.bias + o) : 0.f);
                  if (p.accumulate) x += orow[o];
                  orow[o] = x;
                }
              }
            }
          }
        } else if (MODE == X3_GSO) {
          const int h = m0 + row;
          if (h < p.N) {
            const unsigned J0 = (unsigned)(n0 + c0);
            unsigned gg = J0 / (unsigned)p.C, cc = J0 - gg * (unsigned)p.C;
            const unsigned Jtot = (unsigned)(p.G * p.C);
            if (p.vec && J0 + 16 <= Jtot) {               // C % 4 == 0: a quad of columns never straddles a group
#pragma unroll
              for (int i = 0; i < 16; i += 4) {
                const long long idx = ((long long)gg * p.N + h) * p.C + cc;
                float4 x = make_float4(p.alpha * v[i], p.alpha * v[i + 1], p.alpha * v[i + 2], p.alpha * v[i + 3]);
                if (p.aux) {
                  const float4 y = __ldg(reinterpret_cast<const float4*>(p.aux + idx));
                  x.x += p.beta * y.x; x.y += p.beta * y.y; x.z += p.beta * y.z; x.w += p.beta * y.w;
                }
                *reinterpret_cast<float4*>(p.out + idx) = x;
                cc += 4;
                if (cc >= (unsigned)p.C) { cc = 0; ++gg; }
              }
            } else {
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                if (J0 + i < Jtot) {
                  const long long idx = ((long long)gg * p.N + h) * p.C + cc;
                  float x = p.alpha * v[i];
                  if (p.aux) x += p.beta * __ldg(p.aux + idx);
                  p.out[idx] = x;
                }
                if (++cc >= (unsigned)p.C) { cc = 0; ++gg; }
              }
            }
          }
        } else {
          const int o = m0 + row;
          if (o < p.Co) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const int mm = n0 + c0 + i;
              if (mm < p.Mw) {
                if (p.atomic) atomicAdd(p.dst + (long long)mm * p.Co + o, v[i]);
                else p.dst[((long long)t.ks * p.Mw + mm) * p.Co + o] = v[i];
              }
            }
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
  if (warp == kX3EpiWarps) tmem_dealloc(tmem_base, ncols);
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
inline int x3_pow2_width(long long n) {          // tile width: power of two in [16, 256]
  int bn = 16;
  while (bn < 256 && bn < n) bn <<= 1;
  return bn;
}
inline void x3_plan_tiles(X3Params& p, long long Mrows, long long Ncols) {
  p.BN = x3_pow2_width(Ncols);
  p.m_tiles = (int)((Mrows + 127) / 128);
  p.n_tiles = (int)((Ncols + p.BN - 1) / p.BN);
  p.stage_bytes = 2u * kX3ATile + 2u * (uint32_t)p.BN * 128u;
  int S = (int)((kSmemBudget - 1024) / p.stage_bytes);
  p.S = S > kX3MaxStages ? kX3MaxStages : S;
}
template <int MODE>
inline void x3_launch(const X3Params& p, const char* name, cudaStream_t stream) {
  const long long tiles = (long long)p.m_tiles * p.n_tiles * p.k_splits;
  STGCN_CHECK(tiles < (1LL << 31) && p.Ktot < (1LL << 31) && p.k_per_split < (1LL << 31), STGCN_E_UNSUPPORTED, "x3 GEMM: index range");
  const int grid = (int)(tiles < sm_count() ? tiles : sm_count());
  const size_t smem = (size_t)p.S * p.stage_bytes + 1024;
  STGCN_CUDA(cudaFuncSetAttribute(umma_x3_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  STGCN_LAUNCH_NAMED(name, umma_x3_kernel<MODE>, grid, kX3Threads, smem, stream, p);
}
inline bool x3_al16(const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; }

// out = tap GEMM (simt::launch_tapgemm's contract).  Returns false when the shape is not served.
inline bool x3_tapgemm(const simt::TapArgs<float, float>& a, cudaStream_t stream) {
  if (a.rows <= 0 || a.Co <= 0 || a.Cin % 4 != 0 || a.Cin < 4 || !x3_al16(a.in) || a.rows >= (1LL << 31)) return false;
  X3Params p{};
  x3_plan_tiles(p, a.rows, a.Co);
  p.k_splits = 1; p.Ktot = (long long)a.ntaps * a.Cin; p.k_per_split = (p.Ktot + kX3KC - 1) / kX3KC * kX3KC;
  p.in = a.in; p.wt = a.wt; p.bias = a.bias; p.out = a.out; p.rows = a.rows; p.Cin = a.Cin; p.Co = a.Co;
  p.ntaps = a.ntaps; p.ldo = a.ldo; p.accumulate = a.accumulate; p.map = a.map;
  p.vec = (a.ldo % 4 == 0 && x3_al16(a.out)) ? 1 : 0;
  x3_launch<X3_TAP>(p, "umma_x3_kernel<TAP>", stream);
  return true;
}
// node contraction (simt::launch_gso's contract)
inline bool x3_gso(const simt::GsoArgs<float>& a, cudaStream_t stream) {
  if (a.G <= 0 || a.N <= 0 || a.C <= 0 || a.G * a.C >= (1LL << 31)) return false;
  X3Params p{};
  x3_plan_tiles(p, a.N, a.G * a.C);
  p.k_splits = 1; p.Ktot = a.N; p.k_per_split = (p.Ktot + kX3KC - 1) / kX3KC * kX3KC;
  p.L = a.M; p.trans = a.trans; p.in = a.in; p.aux = a.aux; p.out = a.out; p.N = a.N; p.C = a.C; p.G = a.G;
  p.alpha = a.alpha; p.beta = a.beta;
  p.vec = (a.C % 4 == 0 && x3_al16(a.out) && (!a.aux || x3_al16(a.aux))) ? 1 : 0;
  x3_launch<X3_GSO>(p, "umma_x3_kernel<GSO>", stream);
  return true;
}
// weight gradient (simt::launch_wgrad's contract: dwt pre-zeroed, += the contraction over rows).  `partial` is the caller's
// scratch of simt::wgrad_partial_elems(rows, Mw, Co) floats (or nullptr: fp32 atomics into dwt).
inline bool x3_wgrad(const simt::WgradArgs<float>& a, cudaStream_t stream) {
  if (a.rows <= 0 || a.Co <= 0 || a.rows >= (1LL << 31)) return false;
  X3Params p{};
  p.Kw = a.ntaps * a.Cin; p.Mw = p.Kw + (a.bias_row ? 1 : 0);
  x3_plan_tiles(p, a.Co, p.Mw);
  const int base_tiles = p.m_tiles * p.n_tiles;
  long long want = (2LL * sm_count() + base_tiles - 1) / base_tiles;       // ~2 tiles per SM
  const long long cap = a.partial ? (long long)simt::plan_wgrad_simt(a.rows, p.Mw, a.Co).chunks : want;
  if (want > cap) want = cap;
  long long kps = (a.rows + want - 1) / want;
  if (kps < 1024) kps = 1024;                                              // at least 32 stages per tile
  kps = (kps + kX3KC - 1) / kX3KC * kX3KC;
  p.k_per_split = kps; p.Ktot = a.rows; p.k_splits = (int)((a.rows + kps - 1) / kps);
  p.in = a.in; p.dz = a.dz; p.rows = a.rows; p.Cin = a.Cin; p.Co = a.Co; p.ntaps = a.ntaps; p.ldz = a.ldz;
  p.bias_row = a.bias_row; p.map = a.map;
  p.atomic = a.partial ? 0 : 1;
  p.dst = a.partial ? a.partial : a.dwt;
  x3_launch<X3_WGRAD>(p, "umma_x3_kernel<WGRAD>", stream);
  if (a.partial) simt::launch_reduce_partials(a.partial, a.dwt, p.Mw * a.Co, p.k_splits, stream);
  return true;
}

}  // namespace umma
}  // namespace stgcn
