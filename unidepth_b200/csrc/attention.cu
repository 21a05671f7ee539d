// Fused attention forward for sm_100a:  O = softmax(Q K^T * scale) V   (no mask, no dropout), head dim 64.
// One kernel, attn_fwd2_kernel (design notes in front of it); a split-f16 fp32 variant for the parity mode at the end.
// Replaces F.scaled_dot_product_attention (reference: metadinov2/attention.py:58, layers/attention.py:136).
#include "common.h"
#include "ptx.cuh"

namespace udb {

constexpr int AT_BQ = 128;       // queries per CTA
constexpr int AT_BK = 128;       // keys per tile
constexpr int AT_HK = 64;        // keys per half tile (softmax / QK^T granularity)
constexpr int AT_CTAS_PER_SM = 2;

#ifdef UDB_ATTN_TRACE   // variant build only: event timeline of CTA (1,0,0): softmax warp 2 + the MMA warp
__device__ long long g_attn_trace[32 * 16];
#define AT_EV(j, k) do { if (trace_on && (j) < 32) g_attn_trace[(j) * 16 + (k)] = clock64(); } while (0)
#else
#define AT_EV(j, k) do {} while (0)
#endif

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

// The MUFU unit (16 ex2 / clk / SM) is the binding pipe of d = 64 attention, so every UDB_ATTN_POLY-th PAIR of scores takes
// its exp2 on the FMA / ALU pipes instead: x = k + f, k = round(x), 2^f by a degree-4 minimax polynomial on [-0.5, 0.5]
// (relative error 2.7e-6, two orders below the f16 rounding of P), 2^k added into the exponent field; inputs clamped to
// >= -100 (exp2 = 0 in f16 anyway).  Same-box A/B on the B200 (profiles/r02_attn_poly_ab.txt, round-1 kernel, degree 3):
// share 1/4 -> 4.12 ms of attention per step instead of 4.44 (1/3 and 1/5: 4.22; 1/2: 4.53, slower -- the FMA issue slots
// become the limit; 1/8: 4.30); on the present kernel 1/3 .. 1/6 are within noise of each other, none: +4 %
// (profiles/r02_attn_kernels_ab.txt).  -DUDB_ATTN_POLY=0 turns it off.
#ifndef UDB_ATTN_POLY
#define UDB_ATTN_POLY 4
#endif
__device__ __forceinline__ void exp2_poly_pair(const uint64_t x2, float& e0, float& e1) {
  float x0, x1;
  unpack2(x2, x0, x1);
  const uint64_t xc = pack2(fmaxf(x0, -100.f), fmaxf(x1, -100.f));
  const uint64_t magic = pack2(12582912.f, 12582912.f);        // 1.5 * 2^23: x + magic rounds to an integer in the mantissa
  const uint64_t t2 = add2(xc, magic);
  const uint64_t n2 = add2(t2, pack2(-12582912.f, -12582912.f));
  const uint64_t f2 = fma2(n2, pack2(-1.f, -1.f), xc);
  uint64_t q = fma2(pack2(0.009570102207362652f, 0.009570102207362652f), f2, pack2(0.05591785907745361f, 0.05591785907745361f));
  q = fma2(q, f2, pack2(0.240247443318367f, 0.240247443318367f));
  q = fma2(q, f2, pack2(0.6931217908859253f, 0.6931217908859253f));
  q = fma2(q, f2, pack2(0.9999992847442627f, 0.9999992847442627f));
  float p0, p1, t0, t1;
  unpack2(q, p0, p1);
  unpack2(t2, t0, t1);
  e0 = __uint_as_float(__float_as_uint(p0) + (__float_as_uint(t0) << 23));
  e1 = __uint_as_float(__float_as_uint(p1) + (__float_as_uint(t1) << 23));
}

// exp2(s*scale - m) for the 64 scores of one half row; returns their sum; P (f16 pairs) into pk[0..31].
// MASK: only the first kv_left entries are valid keys (last tile; kv_left may be <= 0).
template <bool MASK, int W>
__device__ __forceinline__ float softmax_half(const uint32_t (&sv)[W], uint32_t* pk, const float sc, const float m_used,
                                              const int kv_left) {
  const uint64_t sc2 = pack2(sc, sc), nm2 = pack2(-m_used, -m_used);
  uint64_t ps[4] = {0ull, 0ull, 0ull, 0ull};   // independent partial sums: no serial FADD chain behind the MUFUs
#pragma unroll
  for (int i = 0; i < W; i += 2) {
    float e0, e1;
    if (UDB_ATTN_POLY && ((i >> 1) % (UDB_ATTN_POLY ? UDB_ATTN_POLY : 1)) == (UDB_ATTN_POLY ? UDB_ATTN_POLY - 1 : 1)) {
      exp2_poly_pair(fma2(pack2u(sv[i], sv[i + 1]), sc2, nm2), e0, e1);
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
    pk[i >> 1] = pack_half2(e0, e1);
  }
  float ps0, ps1;
  unpack2(add2(add2(ps[0], ps[1]), add2(ps[2], ps[3])), ps0, ps1);
  return ps0 + ps1;
}

template <bool MASK, int W>
__device__ __forceinline__ float half_max(const uint32_t (&sv)[W], const int kv_left) {
  float m0 = -INFINITY, m1 = -INFINITY;    // two chains for ILP
  if (!MASK) {
#pragma unroll
    for (int i = 0; i < W; i += 4) {
      m0 = max3(m0, __uint_as_float(sv[i]), __uint_as_float(sv[i + 1]));
      m1 = max3(m1, __uint_as_float(sv[i + 2]), __uint_as_float(sv[i + 3]));
    }
  } else {
#pragma unroll
    for (int i = 0; i < W; ++i) m0 = (i < kv_left) ? fmaxf(m0, __uint_as_float(sv[i])) : m0;
  }
  return fmaxf(m0, m1);
}

// ---------------------------------------------------------------------------------------------
// One CTA = one (batch, head, 128-query tile); two CTAs per SM (TMEM 2 x 256 columns, 113 KB of shared memory each); keys in
// tiles of 128 = two 64-key halves.  Warps: 0 = TMA producer (Q once, K / V tiles through 3-stage rings), 1 = TMEM allocator
// and MMA issuer (whole warp in the control flow, one elected lane issues), 2..9 = two groups of four softmax warps, one
// query row per thread.  The round-1 kernel had ONE softmax group working through both halves with P staged in shared
// memory; it ran at 0.34 of the tensor peak with no unit above 53 % -- every warp waiting on a chain.  Here the two halves are
// INDEPENDENT online-softmax streams, each with its own group of four softmax warps,
// its own reference maximum / row sum and its own output accumulator in tensor memory; the streams are merged once, in the
// epilogue (O = (a0 O0 + a1 O1) / (a0 l0 + a1 l1), a_g = 2^(m_g - max m)).  What that buys:
//   * four softmax warps per sub-partition instead of two, each with half the dependent chain per tile -- the MUFU pipe
//     (the binding unit at head dim 64) finds a ready warp far more often;
//   * no cross-thread traffic in the main loop: a row's two threads never exchange a maximum;
//   * P never touches shared memory: the f16 probabilities overwrite the first 32 columns of their own score half with
//     tcgen05.st and the PV product takes them as the TMEM A operand (no STS / proxy fence / 32 KB buffer, and the MMA
//     runs at its 32-clk floor instead of the 53 clk the shared-memory A read costs at N = 64);
//   * registers: 32 scores at a time (speculation unit = 32 keys), 80 per thread (cap 96 for 2 x 320 threads per SM).
// TMEM (256 columns): S_g / P_g at [64g, 64g+64), O_g at [128+64g, 128+64g+64).  Because P_g aliases S_g the issuer puts
// PV(j,g) and QK(j+1,g) back to back (the tensor pipe executes one thread's MMAs in order), and the commit that publishes
// S_g(j+1) also tells the softmax group that O_g is quiescent.
// Measured on the B200 (tools/tmem_probe.cu, issue_probe.cu, attn_trace.py; profiles/r02_attn2_trace.txt, r02_*_probe.txt):
// tcgen05.ld.x32 + wait costs a lone warp ~120 clk (16 warps reach 356 B/clk per SM, so TMEM reads are latency, not
// bandwidth); MUFU.EX2 saturates at 15.4 / clk / SM; issuing a tcgen05.mma costs ~27 clk, but one thread's tcgen05.commit
// arrivals reach their mbarriers ~130 clk apart -- so the commit a softmax group waits for (s_full) is issued first in every
// event.  One tile takes a CTA ~2350 clk: ~1350 of softmax per stream and ~1000 until its next scores arrive (P published ->
// the issuer wakes -> 8 MMAs through the FIFO the SM's four streams share -> commit -> wake-up), which the other stream and
// the other CTA fill.  ONE issuer warp serving the streams in turn keeps them in anti-phase; a second issuer (one per stream)
// let them drift into phase and was 6 % slower.
// ---------------------------------------------------------------------------------------------
constexpr int AT2_THREADS = 320;   // TMA warp, MMA warp, 2 groups x 4 softmax warps
#ifndef UDB_AT2_STAGES
#define UDB_AT2_STAGES 3
#endif
constexpr int AT2_STAGES = UDB_AT2_STAGES;   // K / V rings (no P buffer: 16 + 3 x 32 KB)
constexpr int AT2_CK = 32;         // keys per speculation chunk

template <int HD>
__global__ void __launch_bounds__(AT2_THREADS, AT_CTAS_PER_SM)
attn_fwd2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                 const __grid_constant__ CUtensorMap tmV, const AttnArgs p) {
  static_assert(HD == 64, "head_dim 64 only");
  constexpr int kQBytes = AT_BQ * HD * 2;      // 16 KB
  constexpr int kKBytes = AT_BK * HD * 2;      // 16 KB
  constexpr uint32_t kTmemCols = 256;
  constexpr int NS = AT2_STAGES;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + kQBytes;                  // NS stages
  uint8_t* sV = sK + NS * kKBytes;             // NS stages
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + NS * kKBytes);
  uint64_t* q_full = bars;
  uint64_t* k_full = bars + 1;             // [NS]
  uint64_t* v_full = k_full + NS;          // [NS]
  uint64_t* k_empty = v_full + NS;         // [NS]  both QK halves of the tile have completed
  uint64_t* v_empty = k_empty + NS;        // [NS]  both PV halves of the tile have completed
  uint64_t* s_full = v_empty + NS;         // [2]   S_g(j) complete in TMEM (and PV(j-1,g) before it)
  uint64_t* p_full = s_full + 2;           // [2]   P_g(j) stored in TMEM by the group's four warps
  uint64_t* o_full = p_full + 2;           //       every MMA of the CTA has completed
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(o_full + 1);
  float2* stats = reinterpret_cast<float2*>(sQ);   // [2][128] {reference maximum, row sum} per group; Q is dead by then

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
    for (int g = 0; g < 2; ++g) {
      mbar_init(&s_full[g], 1);
      mbar_init(&p_full[g], 4);
    }
    mbar_init(o_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<kTmemCols>(tmem_ptr);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr;
  pdl_wait();
#ifdef UDB_ATTN_TRACE
  const bool trace_on = blockIdx.x == 1 && blockIdx.y == 0 && blockIdx.z == 0 && (warp == 1 || warp == 2 || warp == 6) && lane == 0;
#endif

  if (warp == 0) {
    const bool leader = elect_one();
    if (leader) {
      mbar_arrive_expect_tx(q_full, kQBytes);
      tma_load_3d(sQ, &tmQ, q_full, p.q_col0 + head * HD, q0, b);
    }
    __syncwarp();
    int st = 0;
    uint32_t ph = 1;
    for (int j = 0; j < n_tiles; ++j) {
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
      if (++st == NS) { st = 0; ph ^= 1; }
    }
  } else if (warp == 1) {
    // whole warp in the control flow (uniform registers for descriptors), one elected lane issues
    const bool leader = elect_one();
    constexpr uint32_t idesc_qk = umma_idesc_f16(AT_BQ, AT_HK, false, false);
    constexpr uint32_t idesc_pv = umma_idesc_f16(AT_BQ, HD, false, true);   // A = P from TMEM, B = V MN-major
    const uint64_t dq = umma_desc_sw128(smem_u32(sQ), 16, 1024);
    // tcgen05.commit arrivals of one thread reach their mbarriers one after the other, ~130 clk apart (tools/issue_probe.cu:
    // 4 MMAs + 1 commit: barrier seen after 287 clk, + 2 commits: 416, + 3 commits: 620), so in every event the commit a
    // softmax group is waiting for (s_full) goes FIRST and the stage releases (k_empty / v_empty, needed two tiles later)
    // after it; waits that are normally satisfied long before (k_full / v_full) sit in front of the P wait.
    const uint64_t dk0 = umma_desc_sw128(smem_u32(sK), 16, 1024);
    const uint64_t dv0 = umma_desc_sw128(smem_u32(sV), 1024, 1024);
    auto issue_qk = [&](int st, int g) {   // leader only: S_g = Q (K rows [64g, 64g+64))^T of the tile in stage st
      const uint64_t dk = dk0 + (uint64_t)(st * (kKBytes >> 4) + g * (AT_HK * 128 >> 4));
#pragma unroll
      for (int k = 0; k < HD / 16; ++k) umma_f16_ss(tmem_base + g * AT_HK, dq + 2 * k, dk + 2 * k, idesc_qk, k != 0);
      umma_commit(&s_full[g]);
    };
    mbar_wait(q_full, 0);
    mbar_wait(&k_full[0], 0);
    tc_fence_after_sync();
    if (leader) {
      issue_qk(0, 0);
      issue_qk(0, 1);
      umma_commit(&k_empty[0]);
    }
    __syncwarp();
    int st = 0;
    uint32_t ph = 0;
    for (int j = 0; j < n_tiles; ++j) {
      int st_n = st + 1;
      uint32_t ph_n = ph;
      if (st_n == NS) { st_n = 0; ph_n ^= 1; }
      const bool more = j + 1 < n_tiles;
      mbar_wait(&v_full[st], ph);
      if (more) mbar_wait(&k_full[st_n], ph_n);
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        mbar_wait(&p_full[g], j & 1);                // P_g(j) in TMEM, S_g(j) consumed
        tc_fence_after_sync();
        AT_EV(j, 8 + 2 * g);
        const uint64_t dv = dv0 + (uint64_t)(st * (kKBytes >> 4) + g * 4 * (2048 >> 4));
        if (leader) {
#pragma unroll
          for (int k = 0; k < AT_HK / 16; ++k)     // A: 8 columns (16 keys) per step; B: V rows [64g + 16k, +16), 2 KB per step
            umma_f16_ts(tmem_base + 2 * AT_HK + g * HD, tmem_base + g * AT_HK + 8 * k, dv + (uint64_t)(k * 2048 >> 4), idesc_pv,
                        (j | k) != 0);
          if (more) issue_qk(st_n, g);               // overwrites S_g / P_g: ordered behind the PV above; publishes S_g(j+1)
          if (g == 1) {
            umma_commit(&v_empty[st]);
            if (more) umma_commit(&k_empty[st_n]);
          }
        }
        __syncwarp();
        AT_EV(j, 9 + 2 * g);
      }
      st = st_n;
      ph = ph_n;
    }
    if (leader) umma_commit(o_full);
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ softmax warps
    const int g = (warp - 2) >> 2;              // key half this warp's group owns (warps 2..5 / 6..9)
    const int quad = warp & 3;                  // TMEM lane quadrant of this warp
    const int row = quad * 32 + lane;           // query row inside the tile == TMEM lane
    const uint32_t lane_addr = static_cast<uint32_t>(quad * 32) << 16;
    const uint32_t tS = tmem_base + lane_addr + g * AT_HK;
    const uint32_t tO = tmem_base + lane_addr + 2 * AT_HK + g * HD;
    float m_used = -INFINITY, l_run = 0.f;
    const float sc = p.scale_log2;
    constexpr float kRescaleThreshold = 8.0f;   // log2 domain

    // Scores are read 16 columns at a time into two register halves; while one half is exponentiated the other is in
    // flight (tcgen05.ld latency ~120 clk per load would otherwise be exposed four times per tile), and the first half of
    // the next chunk is requested as soon as the current one has been consumed.  Speculation check: the sum of the 32
    // probabilities of a chunk against the current reference -- if it stays <= 2^12 no single value can overflow the f16
    // range or lose precision, so the reference is kept (no max pass over the scores at all); otherwise (rare) the chunk
    // is redone from the scores still in TMEM with its true maximum, rescaling O / l / the stored first chunk if the
    // maximum outgrew the reference by more than 2^8.
    constexpr float kSpecLimit = 4096.0f;
    for (int j = 0; j < n_tiles; ++j) {
      const int kv_left = p.seq_k - j * AT_BK - g * AT_HK;   // valid keys in this group's half (may be <= 0 in the last tile)
      mbar_wait(&s_full[g], j & 1);
      tc_fence_after_sync();
      AT_EV(j, 4 * g);
      uint32_t sa[16], sb[16];
      tmem_ld_32x32b_x16(tS, sa);
      tmem_ld_wait();
#pragma unroll
      for (int c = 0; c < AT_HK / AT2_CK; ++c) {
        const int kvl = kv_left - c * AT2_CK;
        const bool full = kvl >= AT2_CK;
        uint32_t pk[AT2_CK / 2];
        bool careful = (j | c) == 0;             // very first chunk: no reference yet
        tmem_ld_32x32b_x16(tS + c * AT2_CK + 16, sb);                       // second half in flight
        if (!careful) {
          float psum = full ? softmax_half<false>(sa, pk, sc, m_used, kvl) : softmax_half<true>(sa, pk, sc, m_used, kvl);
          tmem_ld_wait();
          if (c == 0) tmem_ld_32x32b_x16(tS + AT2_CK, sa);                  // next chunk's first half in flight
          psum += full ? softmax_half<false>(sb, pk + 8, sc, m_used, kvl - 16) : softmax_half<true>(sb, pk + 8, sc, m_used, kvl - 16);
          careful = __any_sync(0xffffffffu, !(psum <= kSpecLimit));
          if (!careful) l_run += psum;
        }
        if (careful) {
          tmem_ld_wait();                         // rare: whatever is in flight lands first, then the chunk is read again
          tmem_ld_32x32b_x16(tS + c * AT2_CK, sa);
          tmem_ld_32x32b_x16(tS + c * AT2_CK + 16, sb);
          tmem_ld_wait();
          const float m_chunk = sc * (full ? fmaxf(half_max<false>(sa, kvl), half_max<false>(sb, kvl - 16))
                                           : fmaxf(half_max<true>(sa, kvl), half_max<true>(sb, kvl - 16)));
          const bool need = m_chunk > m_used + kRescaleThreshold;   // true on the very first chunk unless it holds no key
          float alpha = 1.0f;
          if (need) {
            alpha = ex2(m_used - m_chunk);       // 0 on the very first chunk
            m_used = m_chunk;
          }
          if ((j | c) != 0 && __any_sync(0xffffffffu, need)) {
            if (j > 0) {
              // this warp's 32 rows of O_g (quiescent: PV(j-1,g) completed before S_g(j) was published, PV(j,g) needs this
              // warp's arrival); rows that do not need it multiply by 1
              const uint64_t a2 = pack2(alpha, alpha);
#pragma unroll 1
              for (int cc = 0; cc < HD; cc += 16) {
                uint32_t r[16];
                tmem_ld_32x32b_x16(tO + cc, r);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 16; i += 2) {
                  float lo, hi;
                  unpack2(mul2(pack2u(r[i], r[i + 1]), a2), lo, hi);
                  r[i] = __float_as_uint(lo);
                  r[i + 1] = __float_as_uint(hi);
                }
                tmem_st_32x32b_x16(tO + cc, r);
              }
            }
            if (c == 1) {
              // the first chunk of this half was stored against the old reference
              tmem_st_wait();
              uint32_t r[16];
              tmem_ld_32x32b_x16(tS, r);
              tmem_ld_wait();
              const __half2 ah = __float2half2_rn(alpha);
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                __half2 v = __hmul2(*reinterpret_cast<__half2*>(&r[i]), ah);
                r[i] = *reinterpret_cast<uint32_t*>(&v);
              }
              tmem_st_32x32b_x16(tS, r);
            }
          }
          float psum = full ? softmax_half<false>(sa, pk, sc, m_used, kvl) : softmax_half<true>(sa, pk, sc, m_used, kvl);
          psum += full ? softmax_half<false>(sb, pk + 8, sc, m_used, kvl - 16) : softmax_half<true>(sb, pk + 8, sc, m_used, kvl - 16);
          l_run = fmaf(l_run, alpha, psum);
          if (c == 0) tmem_ld_32x32b_x16(tS + AT2_CK, sa);                  // the next chunk's first half again
        }
        tmem_st_32x32b_x16(tS + c * (AT2_CK / 2), pk);   // P chunk c: columns [16c, 16c+16) of the half (scores already consumed)
        AT_EV(j, 4 * g + 1 + c);
        if (c == 0) tmem_ld_wait();
      }
      tmem_st_wait();
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[g]);
      AT_EV(j, 4 * g + 3);
    }
    // epilogue: merge the two streams.  A group that never saw a valid key (l = 0) takes no part.
    mbar_wait(o_full, 0);
    tc_fence_after_sync();
    stats[g * AT_BQ + row] = make_float2(l_run > 0.f ? m_used : -INFINITY, l_run);
    asm volatile("bar.sync 1, 256;" ::: "memory");
    const float2 s0 = stats[row], s1 = stats[AT_BQ + row];
    const float m = fmaxf(s0.x, s1.x);
    const float a0 = ex2(s0.x - m), a1 = ex2(s1.x - m);
    const float inv = 1.0f / fmaf(s0.y, a0, s1.y * a1);
    const float w0 = a0 * inv, w1 = a1 * inv;
    const int q = q0 + row;
    __half* op = p.out + ((long long)b * p.seq_q + (q < p.seq_q ? q : 0)) * p.ldo + p.o_col0 + head * HD + g * 32;
    const uint32_t tO0 = tmem_base + lane_addr + 2 * AT_HK + g * 32;   // this thread's 32 output columns of O_0; O_1 is HD further
#pragma unroll
    for (int c = 0; c < 32; c += 16) {
      uint32_t r0[16], r1[16];
      tmem_ld_32x32b_x16(tO0 + c, r0);
      tmem_ld_32x32b_x16(tO0 + HD + c, r1);
      tmem_ld_wait();
      if (q < p.seq_q) {
        uint32_t o[8];
#pragma unroll
        for (int i = 0; i < 16; i += 2)
          o[i >> 1] = pack_half2(fmaf(__uint_as_float(r0[i]), w0, __uint_as_float(r1[i]) * w1),
                                 fmaf(__uint_as_float(r0[i + 1]), w0, __uint_as_float(r1[i + 1]) * w1));
        *reinterpret_cast<uint4*>(op + c) = make_uint4(o[0], o[1], o[2], o[3]);
        *reinterpret_cast<uint4*>(op + c + 8) = make_uint4(o[4], o[5], o[6], o[7]);
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

// ---------------------------------------------------------------------------------------------
// Split-f16 ("precise") attention: operands arrive as hi + lo f16 pairs, everything is computed in
// fp32 on the CUDA cores (exact expf, f32 products and sums) and the result leaves as a hi/lo pair.
// A parity / debugging mode (udb_attn_t.split): it exists to show that the default path's residual
// against the fp32 reference is operand rounding, not logic.  64 queries x 64 keys per step, 256 threads,
// thread (ty, tx) owns rows ty+16i and columns / head dims tx+16j.
// ---------------------------------------------------------------------------------------------
constexpr int SP_T = 64;          // tile edge
constexpr int SP_P = 65;          // smem pitch (floats): conflict-free row-strided reads
constexpr int SP_SMEM = (4 * SP_T * SP_P + 3 * SP_T) * 4;

__global__ void __launch_bounds__(256) attn_split_f32_kernel(const udb_attn_t a) {
  extern __shared__ float sp_smem[];
  float* Qs = sp_smem;
  float* Ks = Qs + SP_T * SP_P;
  float* Vs = Ks + SP_T * SP_P;
  float* Ps = Vs + SP_T * SP_P;
  float* m_s = Ps + SP_T * SP_P;
  float* l_s = m_s + SP_T;
  float* al_s = l_s + SP_T;
  const int t = threadIdx.x, ty = t >> 4, tx = t & 15;
  const int q0 = blockIdx.x * SP_T, h = blockIdx.y, b = blockIdx.z;
  const __half* qp = reinterpret_cast<const __half*>(a.q);
  const __half* kp = reinterpret_cast<const __half*>(a.k);
  const __half* vp = reinterpret_cast<const __half*>(a.v);
  auto ld = [](const __half* base, long long off, int lo_off) {
    return __half2float(base[off]) + (lo_off ? __half2float(base[off + lo_off]) : 0.f);
  };
  for (int i = t; i < SP_T * SP_T; i += 256) {
    const int r = i >> 6, d = i & 63;
    const int sq = q0 + r;
    Qs[r * SP_P + d] = sq < a.seq_q ? ld(qp, ((long long)b * a.seq_q + sq) * a.ldq + a.q_col0 + h * 64 + d, a.lo_off_q) * a.scale : 0.f;
  }
  if (t < SP_T) { m_s[t] = -INFINITY; l_s[t] = 0.f; }
  float O[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) O[i][j] = 0.f;
  const int n_kt = (a.seq_k + SP_T - 1) / SP_T;
  for (int kt = 0; kt < n_kt; ++kt) {
    __syncthreads();     // previous tile's Ks / Vs / Ps fully consumed (and Qs / m / l initialised)
    for (int i = t; i < SP_T * SP_T; i += 256) {
      const int r = i >> 6, d = i & 63;
      const int sk = kt * SP_T + r;
      const bool ok = sk < a.seq_k;
      const long long row = (long long)b * a.seq_k + sk;
      Ks[r * SP_P + d] = ok ? ld(kp, row * a.ldk + a.k_col0 + h * 64 + d, a.lo_off_k) : 0.f;
      Vs[r * SP_P + d] = ok ? ld(vp, row * a.ldv + a.v_col0 + h * 64 + d, a.lo_off_v) : 0.f;
    }
    __syncthreads();
    float S[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) S[i][j] = 0.f;
    for (int d = 0; d < 64; ++d) {
      float qv[4], kv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) qv[i] = Qs[(ty + 16 * i) * SP_P + d];
#pragma unroll
      for (int j = 0; j < 4; ++j) kv[j] = Ks[(tx + 16 * j) * SP_P + d];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) S[i][j] = fmaf(qv[i], kv[j], S[i][j]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int c = tx + 16 * j;
        Ps[(ty + 16 * i) * SP_P + c] = (kt * SP_T + c < a.seq_k) ? S[i][j] : -INFINITY;
      }
    __syncthreads();
    {   // online softmax: 4 threads per row, 16 columns each
      const int r = t >> 2, part = t & 3;
      float* pr = Ps + r * SP_P + part * 16;
      float mx = -INFINITY;
#pragma unroll
      for (int c = 0; c < 16; ++c) mx = fmaxf(mx, pr[c]);
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
      const float m_old = m_s[r];
      const float m_new = fmaxf(m_old, mx);
      float sum = 0.f;
#pragma unroll
      for (int c = 0; c < 16; ++c) {
        const float pv = expf(pr[c] - m_new);
        pr[c] = pv;
        sum += pv;
      }
      sum += __shfl_xor_sync(0xffffffffu, sum, 1);
      sum += __shfl_xor_sync(0xffffffffu, sum, 2);
      __syncwarp();
      if (part == 0) {
        const float al = expf(m_old - m_new);
        al_s[r] = al;
        m_s[r] = m_new;
        l_s[r] = l_s[r] * al + sum;
      }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float al = al_s[ty + 16 * i];
#pragma unroll
      for (int j = 0; j < 4; ++j) O[i][j] *= al;
    }
    for (int c = 0; c < SP_T; ++c) {
      float pv[4], vv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) pv[i] = Ps[(ty + 16 * i) * SP_P + c];
#pragma unroll
      for (int j = 0; j < 4; ++j) vv[j] = Vs[c * SP_P + tx + 16 * j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) O[i][j] = fmaf(pv[i], vv[j], O[i][j]);
    }
  }
  __half* op = reinterpret_cast<__half*>(a.out);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = ty + 16 * i;
    const int sq = q0 + r;
    if (sq >= a.seq_q) continue;
    const float inv = 1.0f / l_s[r];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float v = O[i][j] * inv;
      const long long off = ((long long)b * a.seq_q + sq) * a.ldo + a.o_col0 + h * 64 + tx + 16 * j;
      const __half hi = __float2half_rn(v);
      op[off] = hi;
      if (a.lo_off_o) op[off + a.lo_off_o] = __float2half_rn(v - __half2float(hi));
    }
  }
}

}  // namespace udb

#ifdef UDB_ATTN_TRACE
extern "C" int udb_attn_trace_read(long long* out) {
  cudaMemcpyFromSymbol(out, udb::g_attn_trace, sizeof(long long) * 32 * 16);
  return 0;
}
#endif
extern "C" int udb_attention_f16(const udb_attn_t* a, void* stream) {
  using namespace udb;
  if (a->head_dim != 64) { set_error("udb_attention_f16: head_dim %d unsupported (64 only)", a->head_dim); return 1; }
  if ((a->ldq | a->ldk | a->ldv | a->ldo) % 8) { set_error("udb_attention_f16: leading dims must be multiples of 8"); return 1; }
  if (a->split) {
    static std::atomic<uint64_t> sp_mask{0};
    if (first_on_device(sp_mask)) {
      cudaError_t e = cudaFuncSetAttribute(attn_split_f32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SP_SMEM);
      if (e != cudaSuccess) { set_error("attention(split): cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return 1; }
    }
    dim3 grid((a->seq_q + SP_T - 1) / SP_T, a->heads, a->B);
    attn_split_f32_kernel<<<grid, 256, SP_SMEM, reinterpret_cast<cudaStream_t>(stream)>>>(*a);
    return check_launch("attn_split_f32_kernel");
  }
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
  dim3 grid((a->seq_q + AT_BQ - 1) / AT_BQ, a->heads, a->B);
  note_work(4.0 * a->B * a->heads * (double)a->seq_q * a->seq_k * HD, 2.0 * a->B * a->heads * HD * (2.0 * a->seq_q + 2.0 * a->seq_k));
  constexpr int smem_bytes = 16384 + 2 * AT2_STAGES * (AT_BK * 128) + 256;
  static_assert(2 * (smem_bytes + 1024) <= 228 * 1024, "two CTAs per SM");
  static std::atomic<uint64_t> attr_mask{0};
  if (first_on_device(attr_mask)) {
    cudaError_t ea = cudaFuncSetAttribute(attn_fwd2_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    if (ea != cudaSuccess) { set_error("attention: cudaFuncSetAttribute: %s", cudaGetErrorString(ea)); return 1; }
  }
  cudaError_t e = launch_ex(attn_fwd2_kernel<HD>, grid, dim3(AT2_THREADS), smem_bytes, reinterpret_cast<cudaStream_t>(stream), 1,
                            tq, tk, tv, p);
  if (e != cudaSuccess) { set_error("attn_fwd_kernel launch: %s", cudaGetErrorString(e)); return 1; }
  return check_launch("attn_fwd_kernel");
}
