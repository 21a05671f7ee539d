// Fused attention forward for sm_100a:  O = softmax(Q K^T * scale) V   (no mask, no dropout)
//
// One CTA = one (batch, head, 128-query tile); two CTAs are resident per SM (TMEM 2 x 256 columns,
// <= 113 KB shared memory each) so one CTA's softmax overlaps the other's tensor-core work.
//   warp 0      : TMA producer (Q once; K_j / V_j tiles of 128 keys through a 2-stage ring)
//   warp 1      : TMEM allocator + MMA issuer (S = Q K_j^T : M128 N128 K64 ;  O_j = P_j V_j : M128 N64 K128)
//   warps 2..9  : softmax, two threads per query row (64 keys each): S from TMEM (tcgen05.ld), max /
//                 sum in fp32, P_j written as f16 into 128B-swizzled shared memory for the PV MMA.
//                 O stays in TMEM, accumulated by the PV MMAs, and is rescaled lazily.
// Replaces F.scaled_dot_product_attention (reference: metadinov2/attention.py:58,
// layers/attention.py:136).
#include "common.h"
#include "ptx.cuh"

namespace udb {

constexpr int AT_BQ = 128;   // queries per CTA
constexpr int AT_BK = 128;   // keys per tile
constexpr int AT_THREADS = 320;   // TMA warp, MMA warp, 8 softmax warps

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

template <int HD>
__global__ void __launch_bounds__(AT_THREADS, 2)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const AttnArgs p) {
  static_assert(HD == 64, "head_dim 64 only");
  constexpr int kQBytes = AT_BQ * HD * 2;      // 16 KB
  constexpr int kKBytes = AT_BK * HD * 2;      // 16 KB
  constexpr int kPBytes = AT_BQ * AT_BK * 2;   // 32 KB (two 64-key sub-tiles of 16 KB)
  constexpr uint32_t kTmemCols = 256;          // S: [0,128)  O: [128,192)
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + kQBytes;                  // 2 stages
  uint8_t* sV = sK + 2 * kKBytes;              // 2 stages
  uint8_t* sP = sV + 2 * kKBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + kPBytes);
  uint64_t* q_full = bars;
  uint64_t* k_full = bars + 1;   // [2]
  uint64_t* v_full = bars + 3;   // [2]
  uint64_t* k_empty = bars + 5;  // [2]  K stage free once QK_j has completed
  uint64_t* v_empty = bars + 12; // [2]  V stage free once PV_j has completed
  uint64_t* s_full = bars + 7;
  uint64_t* p_full = bars + 8;
  uint64_t* o_full = bars + 9;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 10);
  uint16_t* xch = reinterpret_cast<uint16_t*>(bars + 16);   // bars[0..13] + tmem_ptr at bars[10] used   // [2][128] partial row maxima (bf16, rounded up)

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * AT_BQ;
  const int head = blockIdx.y;
  const int b = blockIdx.z;
  const int n_tiles = p.n_kv_tiles;

  if (threadIdx.x == 0) {
    if (smem_u32(smem) & 1023) __trap();
    prefetch_tmap(&tmQ);
    prefetch_tmap(&tmK);
    prefetch_tmap(&tmV);
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&k_empty[i], 1);
      mbar_init(&v_empty[i], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(p_full, 8);
    mbar_init(o_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<kTmemCols>(tmem_ptr);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr;
  const uint32_t tmem_S = tmem_base;
  const uint32_t tmem_O = tmem_base + 128;

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, kQBytes);
      tma_load_3d(sQ, &tmQ, q_full, p.q_col0 + head * HD, q0, b);
      for (int j = 0; j < n_tiles; ++j) {
        const int st = j & 1;
        // separate K / V rings: K_{j+2} can be fetched as soon as QK_j is done (two tiles ahead of its
        // use) instead of after PV_j (one tile ahead) -- the K fetch latency was the critical path
        mbar_wait(&k_empty[st], ((j >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&k_full[st], kKBytes);
        tma_load_3d(sK + st * kKBytes, &tmK, &k_full[st], p.k_col0 + head * HD, j * AT_BK, b);
        mbar_wait(&v_empty[st], ((j >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&v_full[st], kKBytes);
        tma_load_3d(sV + st * kKBytes, &tmV, &v_full[st], p.v_col0 + head * HD, j * AT_BK, b);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_qk = umma_idesc_f16(AT_BQ, AT_BK, false, false);
      constexpr uint32_t idesc_pv = umma_idesc_f16(AT_BQ, HD, false, true);   // B = V is MN-major
      const uint64_t dq = umma_desc_sw128(smem_u32(sQ), 16, 1024);
      auto issue_qk = [&](int j) {
        const int st = j & 1;
        mbar_wait(&k_full[st], (j >> 1) & 1);
        tc_fence_after_sync();
        const uint64_t dk = umma_desc_sw128(smem_u32(sK + st * kKBytes), 16, 1024);
#pragma unroll
        for (int k = 0; k < HD / 16; ++k) umma_f16_ss(tmem_S, dq + 2 * k, dk + 2 * k, idesc_qk, k != 0);
        umma_commit(&k_empty[st]);
        umma_commit(s_full);
      };
      mbar_wait(q_full, 0);
      issue_qk(0);
      for (int j = 0; j < n_tiles; ++j) {
        const int st = j & 1;
        mbar_wait(p_full, j & 1);
        mbar_wait(&v_full[st], (j >> 1) & 1);
        tc_fence_after_sync();
        const uint64_t dv = umma_desc_sw128(smem_u32(sV + st * kKBytes), 1024, 1024);
#pragma unroll
        for (int ks = 0; ks < AT_BK / 16; ++ks) {
          // A = P: sub-tile (ks/4) of 16 KB, 32 B per 16-key step inside the swizzle atom
          const uint64_t dp = umma_desc_sw128(smem_u32(sP + (ks >> 2) * (kPBytes / 2)), 16, 1024) + 2 * (ks & 3);
          // B = V (MN-major): 16 key rows = 2 KB per step
          umma_f16_ss(tmem_O, dp, dv + (uint64_t)(ks * 2048 >> 4), idesc_pv, (j | ks) != 0);
        }
        umma_commit(&v_empty[st]);
        umma_commit(o_full);
        if (j + 1 < n_tiles) issue_qk(j + 1);
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax warps
    // Two threads per query row: warp (2 + quad') handles keys [0,64) of 32 rows, warp (6 + quad')
    // keys [64,128) of the same rows.  The partial row maxima are exchanged through shared memory
    // as bf16 values rounded UP (any common upper bound is a valid softmax reference point), so both
    // threads derive the identical reference maximum.  O stays in TMEM, accumulated by the PV MMAs;
    // it is rescaled (tcgen05.ld -> mul -> tcgen05.st) only when a row maximum grows by more than
    // 2^8 over the current reference ("lazy rescale"), so probabilities are bounded by 2^8 (fine
    // for f16) and the final O / l is exact because both use the same reference.
    const int quad = warp & 3;                  // TMEM lane quadrant of this warp
    const int half = (warp - 2) >> 2;           // which 64 keys of the tile
    const int row = quad * 32 + lane;           // query row inside the tile == TMEM lane
    const uint32_t lane_addr = static_cast<uint32_t>(quad * 32) << 16;
    float m_used = -INFINITY, l_run = 0.f;
    uint8_t* p_sub = sP + half * (kPBytes / 2) + row * 128;
    const int sw = row & 7;
    const float sc = p.scale_log2;
    constexpr float kRescaleThreshold = 8.0f;   // log2 domain
    auto pair_sync = [&]() { asm volatile("bar.sync %0, 64;" ::"r"(1 + quad) : "memory"); };

    for (int j = 0; j < n_tiles; ++j) {
      mbar_wait(s_full, j & 1);
      tc_fence_after_sync();
      const int kv_left = p.seq_k - j * AT_BK - half * 64;   // valid keys in this thread's 64 (may be <= 0)
      const bool full = kv_left >= 64;
      uint32_t sv[64];
      {
        uint32_t(&s0)[32] = *reinterpret_cast<uint32_t(*)[32]>(&sv[0]);
        uint32_t(&s1)[32] = *reinterpret_cast<uint32_t(*)[32]>(&sv[32]);
        tmem_ld_32x32b_x32(tmem_S + lane_addr + half * 64, s0);
        tmem_ld_32x32b_x32(tmem_S + lane_addr + half * 64 + 32, s1);
        tmem_ld_wait();
      }
      float mx = -INFINITY;
      if (full) {
#pragma unroll
        for (int i = 0; i < 64; i += 2) mx = max3(mx, __uint_as_float(sv[i]), __uint_as_float(sv[i + 1]));
      } else {
#pragma unroll
        for (int i = 0; i < 64; ++i) mx = (i < kv_left) ? fmaxf(mx, __uint_as_float(sv[i])) : mx;
      }
      // exchange partial maxima (scaled domain), rounded up to bf16
      float mp = mx * sc;
      uint32_t mb = 0xFF800000u;                               // -inf
      if (mp > -INFINITY) mb = (__float_as_uint(mp + fabsf(mp) * 0.0079f) & 0xFFFF0000u);
      xch[half * 128 + row] = static_cast<uint16_t>(mb >> 16);
      pair_sync();
      const uint32_t ob = static_cast<uint32_t>(xch[(half ^ 1) * 128 + row]) << 16;
      const float m_tile = fmaxf(__uint_as_float(mb), __uint_as_float(ob));
      const bool need = m_tile > m_used + kRescaleThreshold;   // always true on the first tile
      float alpha = 1.0f;
      if (need) {
        alpha = ex2(m_used - m_tile);      // 0 on the first tile
        m_used = m_tile;
      }
      bool o_waited = false;
      if (j > 0 && __any_sync(0xffffffffu, need)) {
        // rescale this warp's 32 rows x 32 columns of O (rows that do not need it multiply by 1)
        mbar_wait(o_full, (j - 1) & 1);
        tc_fence_after_sync();
        o_waited = true;
        const uint64_t a2 = pack2(alpha, alpha);
#pragma unroll 1
        for (int c = 0; c < 32; c += 16) {      // 16 columns at a time: the score row is still live
          uint32_t r[16];
          tmem_ld_32x32b_x16(tmem_O + lane_addr + half * 32 + c, r);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; i += 2) {
            float lo, hi;
            unpack2(mul2(pack2u(r[i], r[i + 1]), a2), lo, hi);
            r[i] = __float_as_uint(lo);
            r[i + 1] = __float_as_uint(hi);
          }
          tmem_st_32x32b_x16(tmem_O + lane_addr + half * 32 + c, r);
        }
        tmem_st_wait();
      }
      // p = exp2(s*scale - m_used); row sum; f16 P into swizzled smem
      const uint64_t sc2 = pack2(sc, sc), nm2 = pack2(-m_used, -m_used);
      uint64_t psum2 = 0ull;
#pragma unroll
      for (int i = 0; i < 64; i += 2) {
        float t0, t1;
        unpack2(fma2(pack2u(sv[i], sv[i + 1]), sc2, nm2), t0, t1);
        float e0 = ex2(t0), e1 = ex2(t1);
        if (!full) {
          e0 = (i < kv_left) ? e0 : 0.f;
          e1 = (i + 1 < kv_left) ? e1 : 0.f;
        }
        psum2 = add2(psum2, pack2(e0, e1));
        sv[i >> 1] = pack_half2(e0, e1);   // in place: pair i -> word i/2 (already consumed)
      }
      float ps0, ps1;
      unpack2(psum2, ps0, ps1);
      l_run = fmaf(l_run, alpha, ps0 + ps1);
      // the PV MMA of the previous tile must have finished reading sP before it is overwritten
      if (j > 0 && !o_waited) {
        mbar_wait(o_full, (j - 1) & 1);
        tc_fence_after_sync();
      }
#pragma unroll
      for (int q = 0; q < 8; ++q)              // 8 chunks of 8 halves (16 B) in this thread's sub-tile row
        *reinterpret_cast<uint4*>(p_sub + ((q ^ sw) << 4)) =
            make_uint4(sv[4 * q], sv[4 * q + 1], sv[4 * q + 2], sv[4 * q + 3]);
      fence_proxy_async_smem();
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
    }
    // epilogue: O / l ; the two threads of a row add their partial sums through the (now idle) K stage
    mbar_wait(o_full, (n_tiles - 1) & 1);
    tc_fence_after_sync();
    float* xl = reinterpret_cast<float*>(sK);
    xl[half * 128 + row] = l_run;
    pair_sync();
    const float inv = 1.0f / (l_run + xl[(half ^ 1) * 128 + row]);
    const int q = q0 + row;
    uint32_t r[32];
    tmem_ld_32x32b_x32(tmem_O + lane_addr + half * 32, r);
    tmem_ld_wait();
    if (q < p.seq_q) {
      __half* op = p.out + ((long long)b * p.seq_q + q) * p.ldo + p.o_col0 + head * HD + half * 32;
#pragma unroll
      for (int i = 0; i < 32; i += 8) {
        *reinterpret_cast<uint4*>(op + i) = make_uint4(
            pack_half2(__uint_as_float(r[i]) * inv, __uint_as_float(r[i + 1]) * inv),
            pack_half2(__uint_as_float(r[i + 2]) * inv, __uint_as_float(r[i + 3]) * inv),
            pack_half2(__uint_as_float(r[i + 4]) * inv, __uint_as_float(r[i + 5]) * inv),
            pack_half2(__uint_as_float(r[i + 6]) * inv, __uint_as_float(r[i + 7]) * inv));
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

extern "C" int udb_attention_f16(const udb_attn_t* a, void* stream) {
  using namespace udb;
  if (a->head_dim != 64) { set_error("udb_attention_f16: head_dim %d unsupported (64 only)", a->head_dim); return 1; }
  if ((a->ldq | a->ldk | a->ldv | a->ldo) % 8) { set_error("udb_attention_f16: leading dims must be multiples of 8"); return 1; }
  constexpr int HD = 64;
  CUtensorMap tq, tk, tv;
  const uint32_t box[3] = {HD, 128, 1};
  {
    const uint64_t dims[3] = {(uint64_t)a->ldq, (uint64_t)a->seq_q, (uint64_t)a->B};
    const uint64_t str[2] = {(uint64_t)a->ldq * 2, (uint64_t)a->seq_q * a->ldq * 2};
    if (make_tmap_f16(&tq, a->q, 3, dims, str, box, true)) return 1;
  }
  {
    const uint64_t dims[3] = {(uint64_t)a->ldk, (uint64_t)a->seq_k, (uint64_t)a->B};
    const uint64_t str[2] = {(uint64_t)a->ldk * 2, (uint64_t)a->seq_k * a->ldk * 2};
    if (make_tmap_f16(&tk, a->k, 3, dims, str, box, true)) return 1;
  }
  {
    const uint64_t dims[3] = {(uint64_t)a->ldv, (uint64_t)a->seq_k, (uint64_t)a->B};
    const uint64_t str[2] = {(uint64_t)a->ldv * 2, (uint64_t)a->seq_k * a->ldv * 2};
    if (make_tmap_f16(&tv, a->v, 3, dims, str, box, true)) return 1;
  }
  AttnArgs p{};
  p.out = reinterpret_cast<__half*>(a->out);
  p.seq_q = a->seq_q; p.seq_k = a->seq_k;
  p.n_kv_tiles = (a->seq_k + AT_BK - 1) / AT_BK;
  p.ldo = a->ldo; p.o_col0 = a->o_col0;
  p.q_col0 = a->q_col0; p.k_col0 = a->k_col0; p.v_col0 = a->v_col0;
  p.scale_log2 = a->scale * 1.4426950408889634f;
  constexpr int smem_bytes = 16384 + 2 * 16384 + 2 * 16384 + 32768 + 128 + 512;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(attn_fwd_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    if (e != cudaSuccess) { set_error("attention: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return 1; }
    attr_set = true;
  }
  dim3 grid((a->seq_q + AT_BQ - 1) / AT_BQ, a->heads, a->B);
  attn_fwd_kernel<HD><<<grid, AT_THREADS, smem_bytes, reinterpret_cast<cudaStream_t>(stream)>>>(tq, tk, tv, p);
  return check_launch("attn_fwd_kernel");
}
