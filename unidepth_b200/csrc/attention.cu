// Fused attention forward for sm_100a:  O = softmax(Q K^T * scale) V   (no mask, no dropout)
//
// One CTA = one (batch, head, 128-query tile); two CTAs are resident per SM (TMEM 2 x 256 columns,
// <= 113 KB shared memory each).  Keys are processed in tiles of 128:
//   warp 0      : TMA producer (Q once; K_j / V_j tiles through two independent 2-stage rings)
//   warp 1      : TMEM allocator + MMA issuer.  S_j = Q K_j^T (M128 N128 K64) into TMEM; the softmax
//                 warps copy the whole score row to registers first thing and release the buffer, so
//                 S_{j+1} is computed while they work on tile j.  O += P_j V_j (M128 N64 K128, V is the
//                 MN-major B operand) accumulates in TMEM.
//   warps 2..5  : softmax, one query row per thread, 128 scores in registers: row max / sum in fp32
//                 with packed f32x2 math (optional FMA-pipe polynomial exp2, off by default), P_j as f16
//                 into 128B-swizzled shared memory.  O is rescaled lazily: only when a row maximum
//                 grows by more than 2^8 over the reference the probabilities are expressed against.
// Replaces F.scaled_dot_product_attention (reference: metadinov2/attention.py:58,
// layers/attention.py:136).
#include "common.h"
#include "ptx.cuh"

// every UDB_ATTN_POLY-th pair of probabilities uses the FMA-pipe polynomial exp2 (0 = never).
// Measured on B200 (8x16x1611^2): POLY 0: 162 us, 4: 174 us, 2: 180 us -- the kernel is issue /
// latency bound (2 softmax warps per SM sub-partition), not MUFU bound, so the extra FMA-pipe
// instructions cost more than the MUFU work they save; kept as a compile-time option.
#ifndef UDB_ATTN_POLY
#define UDB_ATTN_POLY 0
#endif

namespace udb {

constexpr int AT_BQ = 128;       // queries per CTA
#ifndef UDB_ATTN_SPEC
#define UDB_ATTN_SPEC 1   // speculative exp2 against the running reference (rare redo)
#endif
#ifdef UDB_ATTN_TIMING   // variant build only: per-phase clock64 accumulation in the softmax warps
__device__ unsigned long long g_attn_phase[8];
#define AT_TICK(k) do { const long long t_now = clock64(); t_acc[k] += t_now - t_prev; t_prev = t_now; } while (0)
#else
#define AT_TICK(k) do {} while (0)
#endif
#ifdef UDB_ATTN_TRACE   // variant build only: event timeline of CTA (1,0,0): softmax warp 2 + the MMA thread
__device__ long long g_attn_trace[32 * 16];
#define AT_EV(j, k) do { if (trace_on && (j) < 32) g_attn_trace[(j) * 16 + (k)] = clock64(); } while (0)
#else
#define AT_EV(j, k) do {} while (0)
#endif
#ifndef UDB_ATTN_BK
#define UDB_ATTN_BK 128
#endif
constexpr int AT_BK = UDB_ATTN_BK;   // keys per tile (128: 2 CTAs/SM; 64: 3 CTAs/SM)
constexpr int AT_CTAS_PER_SM = AT_BK == 128 ? 2 : 3;
constexpr int AT_KV_STAGES = 2;  // per ring
constexpr int AT_THREADS = 192;  // TMA warp, MMA warp, 4 softmax warps

struct AttnArgs {
  __half* out;
  int seq_q, seq_k, n_kv_tiles;
  int ldo, o_col0;
  int q_col0, k_col0, v_col0;
  float scale_log2;
};

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// exp2 of two values on the FMA / ALU pipes (no MUFU): x = n + f with n = round(x), f in [-0.5, 0.5];
// 2^f by a degree-3 minimax polynomial (max relative error 7.6e-5, far below the f16 rounding of P);
// 2^n is added into the exponent field.  Requires -126 < x < 126 (callers clamp the scores).
__device__ __forceinline__ void exp2_poly_pair(const uint64_t x2, float& e0, float& e1) {
  const uint64_t magic = pack2(12582912.f, 12582912.f);        // 1.5 * 2^23: x + magic rounds to integer
  const uint64_t t2 = add2(x2, magic);
  const uint64_t n2 = add2(t2, pack2(-12582912.f, -12582912.f));
  const uint64_t f2 = fma2(n2, pack2(-1.f, -1.f), x2);
  uint64_t q = fma2(pack2(0.05520550534129143f, 0.05520550534129143f), f2, pack2(0.24261397123336792f, 0.24261397123336792f));
  q = fma2(q, f2, pack2(0.6932547688484192f, 0.6932547688484192f));
  q = fma2(q, f2, pack2(0.9999276995658875f, 0.9999276995658875f));
  float p0, p1, t0, t1;
  unpack2(q, p0, p1);
  unpack2(t2, t0, t1);
  e0 = __uint_as_float(__float_as_uint(p0) + (__float_as_uint(t0) << 23));
  e1 = __uint_as_float(__float_as_uint(p1) + (__float_as_uint(t1) << 23));
}

// exp2(s*scale - m) for the 128 scores of one row; returns the row sum; P (f16) packed in place into
// sv[0..63].  MASK: only the first kv_left entries are valid keys (last tile).  The MUFU unit
// (16 ex2 / clk / SM) is the bottleneck of d=64 attention, so every second pair of elements is
// computed with the polynomial instead.
template <bool MASK>
__device__ __forceinline__ float softmax_row(uint32_t (&sv)[AT_BK], const float sc, const float m_used,
                                             const int kv_left) {
  const uint64_t sc2 = pack2(sc, sc), nm2 = pack2(-m_used, -m_used);
  const float s_floor = (m_used - 100.0f) / sc;      // scores below this give exp2(< -100) = 0 in f16 anyway
  uint64_t ps[4] = {0ull, 0ull, 0ull, 0ull};   // independent partial sums: no serial FADD chain behind the MUFUs
#pragma unroll
  for (int i = 0; i < AT_BK; i += 2) {
    float e0, e1;
    if (UDB_ATTN_POLY > 0 && ((i >> 1) % (UDB_ATTN_POLY > 0 ? UDB_ATTN_POLY : 1)) == (UDB_ATTN_POLY > 0 ? UDB_ATTN_POLY : 1) - 1) {
      const float s0 = fmaxf(__uint_as_float(sv[i]), s_floor), s1 = fmaxf(__uint_as_float(sv[i + 1]), s_floor);
      exp2_poly_pair(fma2(pack2(s0, s1), sc2, nm2), e0, e1);
    } else {
      float t0, t1;
      unpack2(fma2(pack2u(sv[i], sv[i + 1]), sc2, nm2), t0, t1);
      e0 = ex2(t0);
      e1 = ex2(t1);
    }
    if (MASK) {
      e0 = (i < kv_left) ? e0 : 0.f;
      e1 = (i + 1 < kv_left) ? e1 : 0.f;
    }
    ps[(i >> 1) & 3] = add2(ps[(i >> 1) & 3], pack2(e0, e1));
    sv[i >> 1] = pack_half2(e0, e1);   // pair i -> word i/2 (already consumed)
  }
  float ps0, ps1;
  unpack2(add2(add2(ps[0], ps[1]), add2(ps[2], ps[3])), ps0, ps1);
  return ps0 + ps1;
}

// Speculative variant: probabilities against the CURRENT reference m_used AND the tile's row maximum
// in one pass, so the max (ALU pipe, FMNMX3) overlaps the exp2 (MUFU pipe) instead of preceding it.
// The caller checks afterwards that the maximum did not outgrow the reference by more than the
// lazy-rescale threshold; if it did (rare) the results -- possibly overflowed -- are discarded and
// the tile is redone from the scores still held in TMEM.
template <bool MASK>
__device__ __forceinline__ float softmax_row_spec(uint32_t (&sv)[AT_BK], const float sc, const float m_used,
                                                  const int kv_left, float& mx_out) {
  const uint64_t sc2 = pack2(sc, sc), nm2 = pack2(-m_used, -m_used);
  uint64_t ps[4] = {0ull, 0ull, 0ull, 0ull};
  float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
  for (int i = 0; i < AT_BK; i += 2) {
    const float s0 = __uint_as_float(sv[i]), s1 = __uint_as_float(sv[i + 1]);
    float t0, t1;
    unpack2(fma2(pack2(s0, s1), sc2, nm2), t0, t1);
    float e0 = ex2(t0), e1 = ex2(t1);
    if (MASK) {
      e0 = (i < kv_left) ? e0 : 0.f;
      e1 = (i + 1 < kv_left) ? e1 : 0.f;
      m0 = (i < kv_left) ? fmaxf(m0, s0) : m0;
      m1 = (i + 1 < kv_left) ? fmaxf(m1, s1) : m1;
    } else if ((i >> 1) & 1) {
      m1 = max3(m1, s0, s1);
    } else {
      m0 = max3(m0, s0, s1);
    }
    ps[(i >> 1) & 3] = add2(ps[(i >> 1) & 3], pack2(e0, e1));
    sv[i >> 1] = pack_half2(e0, e1);
  }
  mx_out = fmaxf(m0, m1);
  float ps0, ps1;
  unpack2(add2(add2(ps[0], ps[1]), add2(ps[2], ps[3])), ps0, ps1);
  return ps0 + ps1;
}

template <bool MASK>
__device__ __forceinline__ float row_max(const uint32_t (&sv)[AT_BK], const int kv_left) {
  float m0 = -INFINITY, m1 = -INFINITY;    // two chains for ILP
  if (!MASK) {
#pragma unroll
    for (int i = 0; i < AT_BK; i += 4) {
      m0 = max3(m0, __uint_as_float(sv[i]), __uint_as_float(sv[i + 1]));
      m1 = max3(m1, __uint_as_float(sv[i + 2]), __uint_as_float(sv[i + 3]));
    }
  } else {
#pragma unroll
    for (int i = 0; i < AT_BK; ++i) m0 = (i < kv_left) ? fmaxf(m0, __uint_as_float(sv[i])) : m0;
  }
  return fmaxf(m0, m1);
}

template <int HD>
__global__ void __launch_bounds__(AT_THREADS, AT_CTAS_PER_SM)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const AttnArgs p) {
  static_assert(HD == 64, "head_dim 64 only");
  constexpr int kQBytes = AT_BQ * HD * 2;      // 16 KB
  constexpr int kKBytes = AT_BK * HD * 2;      // 16 KB
  constexpr int kPBytes = AT_BQ * AT_BK * 2;   // 32 KB (two 64-key sub-tiles of 16 KB)
  constexpr uint32_t kTmemCols = AT_BK == 128 ? 256 : 128;   // S: [0,AT_BK)  O: [AT_BK, AT_BK+64)
  constexpr int NS = AT_KV_STAGES;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + kQBytes;                  // NS stages
  uint8_t* sV = sK + NS * kKBytes;             // NS stages
  uint8_t* sP = sV + NS * kKBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + kPBytes);
  uint64_t* q_full = bars;
  uint64_t* k_full = bars + 1;             // [NS]
  uint64_t* v_full = k_full + NS;          // [NS]
  uint64_t* k_empty = v_full + NS;         // [NS]  K stage free once its QK MMA has completed
  uint64_t* v_empty = k_empty + NS;        // [NS]  V stage free once its PV MMA has completed
  uint64_t* s_full = v_empty + NS;         // S_j complete in TMEM
  uint64_t* s_free = s_full + 1;           // S_j copied to registers by all softmax warps
  uint64_t* p_full = s_free + 1;           // P_j written to smem
  uint64_t* p_free = p_full + 1;           // PV_j MMA done (one phase per tile)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(p_free + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * AT_BQ;
  const int head = blockIdx.y;
  const int b = blockIdx.z;
  const int n_tiles = p.n_kv_tiles;
  pdl_launch_dependents();

  if (threadIdx.x == 0) {
    if (smem_u32(smem) & 1023) __trap();
    prefetch_tmap(&tmQ);
    prefetch_tmap(&tmK);
    prefetch_tmap(&tmV);
    mbar_init(q_full, 1);
    for (int i = 0; i < NS; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&k_empty[i], 1);
      mbar_init(&v_empty[i], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(s_free, 4);
    mbar_init(p_full, 4);
    mbar_init(p_free, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<kTmemCols>(tmem_ptr);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr;
  const uint32_t tmem_O = tmem_base + AT_BK;
  pdl_wait();   // everything above overlapped the previous kernel's tail
#ifdef UDB_ATTN_TRACE
  const bool trace_on = blockIdx.x == 1 && blockIdx.y == 0 && blockIdx.z == 0 && (warp == 1 || warp == 2) && lane == 0;
#endif

  if (warp == 0) {
    const bool leader = elect_one();   // whole warp in the control flow, one lane issues (see the MMA warp)
    if (leader) {
      mbar_arrive_expect_tx(q_full, kQBytes);
      tma_load_3d(sQ, &tmQ, q_full, p.q_col0 + head * HD, q0, b);
    }
    __syncwarp();
    for (int j = 0; j < n_tiles; ++j) {
      const int st = j % NS;
      const uint32_t ph = ((j / NS) & 1) ^ 1;
      mbar_wait(&k_empty[st], ph);
      if (leader) {
        mbar_arrive_expect_tx(&k_full[st], kKBytes);
        tma_load_3d(sK + st * kKBytes, &tmK, &k_full[st], p.k_col0 + head * HD, j * AT_BK, b);
      }
      __syncwarp();
      mbar_wait(&v_empty[st], ph);
      if (leader) {
        mbar_arrive_expect_tx(&v_full[st], kKBytes);
        tma_load_3d(sV + st * kKBytes, &tmV, &v_full[st], p.v_col0 + head * HD, j * AT_BK, b);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // The whole warp runs the control flow so that descriptors and addresses stay in uniform registers
    // (under a lane-0 branch ptxas wraps every tcgen05.mma in an ELECT / R2UR waterfall loop that costs
    // ~90 cycles per instruction); one elected lane issues the MMAs and commits.
    const bool leader = elect_one();
    {
      constexpr uint32_t idesc_qk = umma_idesc_f16(AT_BQ, AT_BK, false, false);
      constexpr uint32_t idesc_pv = umma_idesc_f16(AT_BQ, HD, false, true);   // B = V is MN-major
      const uint64_t dq = umma_desc_sw128(smem_u32(sQ), 16, 1024);
      auto issue_qk = [&](int j) {   // S = Q K_j^T
        const int st = j % NS;
        mbar_wait(&k_full[st], (j / NS) & 1);
        tc_fence_after_sync();
        const uint64_t dk = umma_desc_sw128(smem_u32(sK + st * kKBytes), 16, 1024);
        if (leader) {
#pragma unroll
          for (int k = 0; k < HD / 16; ++k) umma_f16_ss(tmem_base, dq + 2 * k, dk + 2 * k, idesc_qk, k != 0);
          umma_commit(&k_empty[st]);
          umma_commit(s_full);
        }
        __syncwarp();
      };
      mbar_wait(q_full, 0);
      issue_qk(0);
      for (int j = 0; j < n_tiles; ++j) {
        const int st = j % NS;
        if (j + 1 < n_tiles) {
          mbar_wait(s_free, j & 1);                // score row of tile j is in registers: S may be overwritten
          AT_EV(j, 8);
          issue_qk(j + 1);
          AT_EV(j, 9);
        }
        mbar_wait(p_full, j & 1);                  // P_j written
        AT_EV(j, 10);
        mbar_wait(&v_full[st], (j / NS) & 1);
        tc_fence_after_sync();
        const uint64_t dv = umma_desc_sw128(smem_u32(sV + st * kKBytes), 1024, 1024);
        const uint64_t dp0 = umma_desc_sw128(smem_u32(sP), 16, 1024);
        if (leader) {
#pragma unroll
          for (int ks = 0; ks < AT_BK / 16; ++ks) {
            // A = P: sub-tile (ks/4) of 16 KB, 32 B per 16-key step inside the swizzle atom; B = V: 2 KB per step
            const uint64_t dp = dp0 + (uint64_t)((ks >> 2) * (AT_BQ * 128) >> 4) + 2 * (ks & 3);
            umma_f16_ss(tmem_O, dp, dv + (uint64_t)(ks * 2048 >> 4), idesc_pv, (j | ks) != 0);
          }
          umma_commit(&v_empty[st]);
          umma_commit(p_free);
        }
        __syncwarp();
        AT_EV(j, 11);
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax warps
    const int quad = warp & 3;                  // TMEM lane quadrant of this warp
    const int row = quad * 32 + lane;           // query row inside the tile == TMEM lane
    const uint32_t lane_addr = static_cast<uint32_t>(quad * 32) << 16;
    float m_used = -INFINITY, l_run = 0.f;
    const int sw = row & 7;
    const float sc = p.scale_log2;
    constexpr float kRescaleThreshold = 8.0f;   // log2 domain

#ifdef UDB_ATTN_TIMING
    long long t_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long t_prev = clock64();
#endif
    for (int j = 0; j < n_tiles; ++j) {
      mbar_wait(s_full, j & 1);
      tc_fence_after_sync();
      AT_TICK(0);
      AT_EV(j, 0);
      const int kv_left = p.seq_k - j * AT_BK;   // valid keys in this tile (>= 1)
      const bool full = kv_left >= AT_BK;
      uint32_t sv[AT_BK];
      auto load_scores = [&]() {
#pragma unroll
        for (int c = 0; c < AT_BK; c += 32)
          tmem_ld_32x32b_x32(tmem_base + lane_addr + c, *reinterpret_cast<uint32_t(*)[32]>(&sv[c]));
        tmem_ld_wait();
      };
      load_scores();
      if (!UDB_ATTN_SPEC) {   // scores are in registers: S may be overwritten by the next QK^T right away
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(s_free);
      }
      AT_TICK(1);
      AT_EV(j, 1);
      bool careful = (j == 0) || !UDB_ATTN_SPEC;   // first tile: no reference yet
      bool pv_done = false;
      if (!careful) {
        // common case: exp2 against the current reference and the row max in the same pass
        float mx;
        const float psum = full ? softmax_row_spec<false>(sv, sc, m_used, kv_left, mx)
                                : softmax_row_spec<true>(sv, sc, m_used, kv_left, mx);
        careful = __any_sync(0xffffffffu, mx * sc > m_used + kRescaleThreshold);
        if (careful) load_scores();      // rare: the speculative results are discarded
        else l_run += psum;
      }
      if (careful) {
        const float m_tile = (full ? row_max<false>(sv, kv_left) : row_max<true>(sv, kv_left)) * sc;
        const bool need = m_tile > m_used + kRescaleThreshold;   // always true on the first tile
        float alpha = 1.0f;
        if (need) {
          alpha = ex2(m_used - m_tile);      // 0 on the first tile
          m_used = m_tile;
        }
        if (j > 0 && __any_sync(0xffffffffu, need)) {
          // rescale this warp's 32 rows of O (rows that do not need it multiply by 1)
          mbar_wait(p_free, (j - 1) & 1);
          tc_fence_after_sync();
          pv_done = true;
          const uint64_t a2 = pack2(alpha, alpha);
#pragma unroll 1
          for (int c = 0; c < HD; c += 16) {
            uint32_t r[16];
            tmem_ld_32x32b_x16(tmem_O + lane_addr + c, r);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; i += 2) {
              float lo, hi;
              unpack2(mul2(pack2u(r[i], r[i + 1]), a2), lo, hi);
              r[i] = __float_as_uint(lo);
              r[i + 1] = __float_as_uint(hi);
            }
            tmem_st_32x32b_x16(tmem_O + lane_addr + c, r);
          }
          tmem_st_wait();
        }
        const float psum = full ? softmax_row<false>(sv, sc, m_used, kv_left) : softmax_row<true>(sv, sc, m_used, kv_left);
        l_run = fmaf(l_run, alpha, psum);
      }
      AT_TICK(2);
      AT_EV(j, 2);
      // the score buffer is released only now (the rare path re-reads it); S_{j+1} is then computed
      // while P_j is packed / stored and is ready when the next iteration starts
      if (UDB_ATTN_SPEC) {
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(s_free);
      }
      AT_TICK(3);
      AT_EV(j, 3);
      // the PV MMA of tile j-1 must have finished reading sP (the next completion of p_free needs
      // this thread's own arrival on p_full, so the parity is unambiguous)
      if (j > 0 && !pv_done) mbar_wait(p_free, (j - 1) & 1);
      AT_TICK(4);
      AT_EV(j, 4);
      uint8_t* p_row = sP + row * 128;
#pragma unroll
      for (int q = 0; q < AT_BK / 8; ++q)      // chunks of 8 halves (16 B); 64-key sub-tiles of 16 KB
        *reinterpret_cast<uint4*>(p_row + (q >> 3) * (AT_BQ * 128) + (((q & 7) ^ sw) << 4)) =
            make_uint4(sv[4 * q], sv[4 * q + 1], sv[4 * q + 2], sv[4 * q + 3]);
      fence_proxy_async_smem();
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
      AT_TICK(5);
      AT_EV(j, 5);
    }
    // epilogue: O / l
    mbar_wait(p_free, (n_tiles - 1) & 1);   // last PV done => all done
    tc_fence_after_sync();
    AT_TICK(6);
#ifdef UDB_ATTN_TIMING
    if (lane == 0) {
      for (int k = 0; k < 7; ++k) atomicAdd(&g_attn_phase[k], (unsigned long long)t_acc[k]);
      atomicAdd(&g_attn_phase[7], (unsigned long long)n_tiles);
    }
#endif
    const float inv = 1.0f / l_run;
    const int q = q0 + row;
    __half* op = p.out + ((long long)b * p.seq_q + (q < p.seq_q ? q : 0)) * p.ldo + p.o_col0 + head * HD;
#pragma unroll
    for (int c = 0; c < HD; c += 32) {
      uint32_t r[32];
      tmem_ld_32x32b_x32(tmem_O + lane_addr + c, r);
      tmem_ld_wait();
      if (q < p.seq_q) {
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
          *reinterpret_cast<uint4*>(op + c + i) = make_uint4(
              pack_half2(__uint_as_float(r[i]) * inv, __uint_as_float(r[i + 1]) * inv),
              pack_half2(__uint_as_float(r[i + 2]) * inv, __uint_as_float(r[i + 3]) * inv),
              pack_half2(__uint_as_float(r[i + 4]) * inv, __uint_as_float(r[i + 5]) * inv),
              pack_half2(__uint_as_float(r[i + 6]) * inv, __uint_as_float(r[i + 7]) * inv));
        }
      }
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc<kTmemCols>(tmem_base);
  }
}

}  // namespace udb

#ifdef UDB_ATTN_TRACE
extern "C" int udb_attn_trace_read(long long* out) {
  cudaMemcpyFromSymbol(out, udb::g_attn_trace, sizeof(long long) * 32 * 16);
  return 0;
}
#endif
#ifdef UDB_ATTN_TIMING
extern "C" int udb_attn_phase_read(unsigned long long* out, int reset) {
  cudaMemcpyFromSymbol(out, udb::g_attn_phase, sizeof(unsigned long long) * 8);
  if (reset) { unsigned long long z[8] = {}; cudaMemcpyToSymbol(udb::g_attn_phase, z, sizeof(z)); }
  return 0;
}
#endif

extern "C" int udb_attention_f16(const udb_attn_t* a, void* stream) {
  using namespace udb;
  if (a->head_dim != 64) { set_error("udb_attention_f16: head_dim %d unsupported (64 only)", a->head_dim); return 1; }
  if ((a->ldq | a->ldk | a->ldv | a->ldo) % 8) { set_error("udb_attention_f16: leading dims must be multiples of 8"); return 1; }
  constexpr int HD = 64;
  CUtensorMap tq, tk, tv;
  const uint32_t box_q[3] = {HD, AT_BQ, 1};
  const uint32_t box_kv[3] = {HD, AT_BK, 1};
  {
    const uint64_t dims[3] = {(uint64_t)a->ldq, (uint64_t)a->seq_q, (uint64_t)a->B};
    const uint64_t str[2] = {(uint64_t)a->ldq * 2, (uint64_t)a->seq_q * a->ldq * 2};
    if (make_tmap_f16(&tq, a->q, 3, dims, str, box_q, true)) return 1;
  }
  {
    const uint64_t dims[3] = {(uint64_t)a->ldk, (uint64_t)a->seq_k, (uint64_t)a->B};
    const uint64_t str[2] = {(uint64_t)a->ldk * 2, (uint64_t)a->seq_k * a->ldk * 2};
    if (make_tmap_f16(&tk, a->k, 3, dims, str, box_kv, true)) return 1;
  }
  {
    const uint64_t dims[3] = {(uint64_t)a->ldv, (uint64_t)a->seq_k, (uint64_t)a->B};
    const uint64_t str[2] = {(uint64_t)a->ldv * 2, (uint64_t)a->seq_k * a->ldv * 2};
    if (make_tmap_f16(&tv, a->v, 3, dims, str, box_kv, true)) return 1;
  }
  AttnArgs p{};
  p.out = reinterpret_cast<__half*>(a->out);
  p.seq_q = a->seq_q; p.seq_k = a->seq_k;
  p.n_kv_tiles = (a->seq_k + AT_BK - 1) / AT_BK;
  p.ldo = a->ldo; p.o_col0 = a->o_col0;
  p.q_col0 = a->q_col0; p.k_col0 = a->k_col0; p.v_col0 = a->v_col0;
  p.scale_log2 = a->scale * 1.4426950408889634f;
#ifndef UDB_ATTN_SMEM_PAD
#define UDB_ATTN_SMEM_PAD 0   // experiments: extra dynamic smem to lower the CTAs/SM
#endif
  constexpr int smem_bytes = 16384 + 2 * AT_KV_STAGES * (AT_BK * 128) + AT_BQ * AT_BK * 2 + 256 + UDB_ATTN_SMEM_PAD;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(attn_fwd_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    if (e != cudaSuccess) { set_error("attention: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return 1; }
    attr_set = true;
  }
  dim3 grid((a->seq_q + AT_BQ - 1) / AT_BQ, a->heads, a->B);
  cudaError_t e = launch_ex(attn_fwd_kernel<HD>, grid, dim3(AT_THREADS), smem_bytes, reinterpret_cast<cudaStream_t>(stream), 1,
                            tq, tk, tv, p);
  if (e != cudaSuccess) { set_error("attn_fwd_kernel launch: %s", cudaGetErrorString(e)); return 1; }
  return check_launch("attn_fwd_kernel");
}
