// Persistent warp-specialised tcgen05 GEMM for sm_100a:   D = epilogue(A . W^T)
//
//   warp 0      : TMA producer  (A tile 128x64 f16, W tile BNx64 f16, 128B swizzle, mbarrier ring)
//   warp 1      : MMA issuer    (one elected thread, tcgen05.mma.cta_group::1.kind::f16, M=128,N=BN,K=16)
//   warp 2      : TMEM allocator (2 accumulator stages of BN fp32 columns)
//   warps 4..11 : epilogue      (tcgen05.ld 32x32b -> registers -> bias/act/gamma/residual -> global)
//
// The A operand is either a row-major matrix (2-D tensor map) or a 3x3 convolution window over an
// NHWC image (4-D tensor map, one (dy,dx,64-channel) slab per k-block; zero padding comes from TMA
// out-of-bounds fill), so linear layers, 1x1 / 3x3 convolutions and k=s transposed convolutions all
// run through this one kernel.  See include/udb.h (udb_gemm) for the reference call sites.
#include <stdlib.h>

#include "common.h"
#include "ptx.cuh"

namespace udb {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int kEpiWarps = 8;
constexpr int kTP = 36;   // pitch (floats) of the per-warp 32-row transpose tile: 16 B aligned rows, conflict-free 128-bit access
constexpr int kThreads = (4 + kEpiWarps) * 32;

struct GemmArgs {
  int M, N, K, num_kb;
  int tiles_m, tiles_n;
  int a_mode;
  int conv_B, conv_H, conv_W, conv_cpb /*64-ch blocks per tap*/, conv_off, conv_TH, conv_TW, conv_tx, conv_ty;
  const float* bias;
  const float* gamma;
  const void* resid;
  int resid_f32;
  void* out;
  int out_f32;
  __half* out2;
  int out2_leaky;
  int act;
  int store_mode;
  long long ldc, ldr;
  int rpg, gstride, roff;
  int resid_mod, resid_roff;
  int ct_k, ct_cout, ct_h, ct_w, ct_pad;
  int conv_coff;
  const float* head_w;
  float head_b, head_add;
  int a_wrap;      // split-f16 operands: k-blocks at or beyond this element offset re-read A from (k - a_wrap); 0 = off
  int out_split;   // f16 `out`: lo half stored out_split elements to the right; 0 = off
  // Fused LayerNorm (udb_gemm_t.ln_*): a PRODUCER writes, per output row and per (column tile, column half), the mean and
  // the centred sum of squares of the values it stores; a CONSUMER whose A operand is the un-normalised f16 copy of those
  // rows merges the partials and applies  v = rstd * (acc - mean * c1[n]) + bias[n]  (weights pre-multiplied by the
  // LayerNorm scale, c1 = their row sums, bias = W ln_bias + bias).
  float* stats_out;
  const float* ln_stats;
  const float* ln_c1;
  int ln_parts, ln_part_cols;
  float ln_eps;
};

template <int BN>
struct GemmCfg {
  static constexpr int kStages = BN >= 256 ? 3 : (BN >= 128 ? 5 : 7);
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBBytes = BN * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kTmemCols = (2 * BN) < 32 ? 32 : (2 * BN);
  // epilogue staging: per epilogue warp a 32x32 f32 transpose tile (XOR-swizzled, no padding) and
  // 2 x 32 row offsets, so that global loads/stores are row-contiguous (coalesced) per instruction
  static constexpr int kStagingBytes = kEpiWarps * (32 * kTP * 4 + 2 * 32 * 4);
  static constexpr int kSmemBytes = kStages * kStageBytes + kStagingBytes + 256 /*barriers*/;
};

// One accumulator tile (128 rows x BN columns, fp32 in TMEM at `t_acc`) -> global memory.
// `mt` indexes this CTA's 128-row block (matrix rows mt*128.. or spatial conv tile mt), `nt` the
// BN-wide column block.  Executed by the 8 epilogue warps; warp (quad, grp) owns TMEM lanes
// [32*quad, +32) and column half `grp`.
struct EpiWarp {
  float* T;              // [32][kTP] transpose tile
  uint32_t* roff_out;    // [32] element offsets of this warp's rows in out / out2
  uint32_t* roff_res;    // [32] element offsets in resid
  int quad, grp, lane, r_in_tile;
};

// LNF: compile the fused-LayerNorm producer / consumer code in (udb_gemm_t.ln_*); the default instantiation does not carry it.
template <int BN, bool LNF = false>
__device__ __forceinline__ void epilogue_tile(const GemmArgs& p, const EpiWarp& w, const uint32_t t_acc,
                                              const int mt, const int nt) {
  constexpr int kGroups = BN >= 64 ? 2 : 1;
  constexpr int kColsPerGrp = BN / kGroups;
  const int grp = w.grp, quad = w.quad, lane = w.lane, r_in_tile = w.r_in_tile;
  float* T = w.T;
  uint32_t* roff_out = w.roff_out;
  uint32_t* roff_res = w.roff_res;
  if (grp < kGroups) {
        // ---- per-row addressing
        bool valid;
        long long out_off = 0, res_off = 0;
        const int m = mt * BM + r_in_tile;
        if (p.store_mode == UDB_STORE_CONVTILE || p.store_mode == UDB_STORE_HEAD) {
          const int per_img = p.conv_tx * p.conv_ty;
          const int b = mt / per_img;
          const int r = mt % per_img;
          const int y = (r / p.conv_tx) * p.conv_TH + r_in_tile / p.conv_TW;
          const int x = (r % p.conv_tx) * p.conv_TW + r_in_tile % p.conv_TW;
          valid = (b < p.conv_B) && (y < p.conv_H) && (x < p.conv_W);
          out_off = (((long long)b * p.conv_H + y) * p.conv_W + x) * p.ldc;
          res_off = (((long long)b * p.conv_H + y) * p.conv_W + x) * p.ldr;
        } else if (p.store_mode == UDB_STORE_CONVT) {
          valid = m < p.M;
          const int hw = p.ct_h * p.ct_w;
          const int b = m / hw;
          const int r = m % hw;
          const int y = r / p.ct_w, x = r % p.ct_w;
          const long long W2 = (long long)p.ct_w * p.ct_k + 2 * p.ct_pad;
          const long long H2 = (long long)p.ct_h * p.ct_k + 2 * p.ct_pad;
          out_off = ((b * H2 + (long long)y * p.ct_k + p.ct_pad) * W2 + (long long)x * p.ct_k + p.ct_pad) * p.ct_cout;
          res_off = out_off;
        } else {
          valid = m < p.M;
          long long orow = m;
          if (p.rpg > 0) orow = (long long)(m / p.rpg) * p.gstride + (m % p.rpg) + p.roff;
          out_off = orow * p.ldc;
          res_off = (p.resid_mod > 0) ? ((long long)(m % p.resid_mod) + p.resid_roff) * p.ldr
                                      : orow * p.ldr;
        }
        const uint32_t vmask = __ballot_sync(0xffffffffu, valid);
        roff_out[lane] = static_cast<uint32_t>(out_off);
        roff_res[lane] = static_cast<uint32_t>(res_off);
        __syncwarp();
        const uint32_t t_row = t_acc + (static_cast<uint32_t>(quad * 32) << 16);
        // fused LayerNorm, consumer side: merge this row's partial statistics (equal counts: plain mean of the means,
        // M2 = sum M2_p + n_p * sum (mean_p - mean)^2)
        float ln_mean = 0.f, ln_rstd = 1.f;
        if (LNF && p.ln_stats) {
          const float2* sp = reinterpret_cast<const float2*>(p.ln_stats) + (long long)(valid ? m : 0) * p.ln_parts;
          float ms = 0.f, m2 = 0.f;
          for (int q = 0; q < p.ln_parts; ++q) ms += sp[q].x;
          ln_mean = ms / (float)p.ln_parts;
          for (int q = 0; q < p.ln_parts; ++q) {
            const float2 t = sp[q];
            m2 += t.y + (float)p.ln_part_cols * (t.x - ln_mean) * (t.x - ln_mean);
          }
          ln_rstd = rsqrtf(m2 / (float)(p.ln_parts * p.ln_part_cols) + p.ln_eps);
        }
        // fused LayerNorm, producer side: running (pivot-shifted) sums over this thread's columns of the tile
        float st_pivot = 0.f, st_s1 = 0.f, st_s2 = 0.f;
        bool st_first = true;
#pragma unroll 1
        for (int c = 0; c < kColsPerGrp; c += 32) {
          const int col = grp * kColsPerGrp + c;   // column inside the tile
          const int n0 = nt * BN + col;            // global column
          uint32_t r[32];
          tmem_ld_32x32b_x32(t_row + col, r);
          tmem_ld_wait();
          if (n0 >= p.N) continue;   // warp-uniform
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
          if (LNF && p.ln_stats) {
            const float4* cp = reinterpret_cast<const float4*>(p.ln_c1 + n0);
            const float mr = ln_mean * ln_rstd;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 c4 = __ldg(cp + j);
              v[4 * j] = fmaf(v[4 * j], ln_rstd, -mr * c4.x);
              v[4 * j + 1] = fmaf(v[4 * j + 1], ln_rstd, -mr * c4.y);
              v[4 * j + 2] = fmaf(v[4 * j + 2], ln_rstd, -mr * c4.z);
              v[4 * j + 3] = fmaf(v[4 * j + 3], ln_rstd, -mr * c4.w);
            }
          }
          if (p.bias) {
            const float4* bp = reinterpret_cast<const float4*>(p.bias + n0);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 b4 = __ldg(bp + j);
              v[4 * j] += b4.x; v[4 * j + 1] += b4.y; v[4 * j + 2] += b4.z; v[4 * j + 3] += b4.w;
            }
          }
          if (p.act == UDB_ACT_GELU) {
#pragma unroll
            for (int j = 0; j < 32; j += 2) gelu_erf_pair(v[j], v[j + 1]);
          } else if (p.act == UDB_ACT_LEAKY) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = leaky(v[j]);
          }
          if (p.store_mode == UDB_STORE_HEAD) {
            float acc = p.head_b;
#pragma unroll
            for (int j = 0; j < 32; ++j) acc = fmaf(v[j], __ldg(p.head_w + j), acc);
            acc = fminf(fmaxf(acc, -8.0f), 8.0f) + p.head_add;
            if (valid) reinterpret_cast<float*>(p.out)[out_off] = expf(acc);
            continue;
          }
          if (p.gamma) {
            const float4* gp = reinterpret_cast<const float4*>(p.gamma + n0);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 g4 = __ldg(gp + j);
              v[4 * j] *= g4.x; v[4 * j + 1] *= g4.y; v[4 * j + 2] *= g4.z; v[4 * j + 3] *= g4.w;
            }
          }
          long long coff = n0;
          if (p.store_mode == UDB_STORE_CONVT) {
            const int tap = n0 / p.ct_cout;
            const int co = n0 % p.ct_cout;
            const long long W2 = (long long)p.ct_w * p.ct_k + 2 * p.ct_pad;
            coff = ((long long)(tap / p.ct_k) * W2 + (tap % p.ct_k)) * p.ct_cout + co;
          }
          // ---- residual / stores through the per-warp transpose tile T[32][TP]: in registers a thread
          //      owns a row; in global memory 8 lanes x 16 B (f32) or 8 B (f16) cover one 32-column row
          //      segment and one instruction covers 4 rows, so every access is contiguous.  All smem
          //      traffic is 128-bit and conflict-free with the 36-float pitch.
          const uint32_t c32 = static_cast<uint32_t>(coff);
          const int l4 = (lane & 7) * 4, rq = lane >> 3;
          float* Trow = T + lane * kTP;
          auto stage_rows = [&](const float (&x)[32]) {      // thread == row  ->  T
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              *reinterpret_cast<float4*>(Trow + j) = make_float4(x[j], x[j + 1], x[j + 2], x[j + 3]);
          };
          auto for_rows = [&](auto&& body) {                 // body(rr, valid) for this lane's 8 rows
            if (vmask == 0xffffffffu) {
#pragma unroll
              for (int i = 0; i < 8; ++i) body(4 * i + rq, true);
            } else {
#pragma unroll
              for (int i = 0; i < 8; ++i) body(4 * i + rq, ((vmask >> (4 * i + rq)) & 1u) != 0);
            }
          };
          if (p.resid) {
            // all 8 row-segment loads of a lane are issued before the first use (latency-bound)
            if (p.resid_f32) {
              const float* rp = reinterpret_cast<const float*>(p.resid) + c32 + l4;
              float4 tmp[8];
              int k = 0;
              for_rows([&](int rr, bool ok) {
                tmp[k++] = ok ? *reinterpret_cast<const float4*>(rp + roff_res[rr]) : make_float4(0.f, 0.f, 0.f, 0.f);
              });
#pragma unroll
              for (int i = 0; i < 8; ++i) *reinterpret_cast<float4*>(T + (4 * i + rq) * kTP + l4) = tmp[i];
            } else {
              const __half* rp = reinterpret_cast<const __half*>(p.resid) + c32 + l4;
              uint2 tmp[8];
              int k = 0;
              for_rows([&](int rr, bool ok) {
                tmp[k++] = ok ? *reinterpret_cast<const uint2*>(rp + roff_res[rr]) : make_uint2(0u, 0u);
              });
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&tmp[i].x));
                const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&tmp[i].y));
                *reinterpret_cast<float4*>(T + (4 * i + rq) * kTP + l4) = make_float4(a.x, a.y, b.x, b.y);
              }
            }
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 t = *reinterpret_cast<const float4*>(Trow + j);
              v[j] += t.x; v[j + 1] += t.y; v[j + 2] += t.z; v[j + 3] += t.w;
            }
            __syncwarp();
          }
          if (LNF && p.stats_out) {
            if (st_first) { st_pivot = v[0]; st_first = false; }
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float dlt = v[j] - st_pivot;
              st_s1 += dlt;
              st_s2 = fmaf(dlt, dlt, st_s2);
            }
          }
          if (p.out) {
            stage_rows(v);
            __syncwarp();
            if (p.out_f32) {
              float* op = reinterpret_cast<float*>(p.out) + c32 + l4;
              for_rows([&](int rr, bool ok) {
                if (ok) *reinterpret_cast<float4*>(op + roff_out[rr]) = *reinterpret_cast<const float4*>(T + rr * kTP + l4);
              });
            } else {
              __half* op = reinterpret_cast<__half*>(p.out) + c32 + l4;
              if (p.out_split == 0) {
                for_rows([&](int rr, bool ok) {
                  const float4 t = *reinterpret_cast<const float4*>(T + rr * kTP + l4);
                  if (ok) *reinterpret_cast<uint2*>(op + roff_out[rr]) = make_uint2(pack_half2(t.x, t.y), pack_half2(t.z, t.w));
                });
              } else {   // split-f16 output: hi and lo = f16(v - f32(hi))
                for_rows([&](int rr, bool ok) {
                  const float4 t = *reinterpret_cast<const float4*>(T + rr * kTP + l4);
                  const uint2 hi = make_uint2(pack_half2(t.x, t.y), pack_half2(t.z, t.w));
                  const float2 h01 = __half22float2(*reinterpret_cast<const __half2*>(&hi.x));
                  const float2 h23 = __half22float2(*reinterpret_cast<const __half2*>(&hi.y));
                  if (ok) {
                    *reinterpret_cast<uint2*>(op + roff_out[rr]) = hi;
                    *reinterpret_cast<uint2*>(op + roff_out[rr] + p.out_split) =
                        make_uint2(pack_half2(t.x - h01.x, t.y - h01.y), pack_half2(t.z - h23.x, t.w - h23.y));
                  }
                });
              }
            }
            __syncwarp();
          }
          if (p.out2) {
            if (p.out2_leaky) {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = leaky(v[j]);
              stage_rows(v);
            } else if (!p.out) {
              stage_rows(v);
            }   // else: T still holds v from the `out` pass
            __syncwarp();
            __half* op = p.out2 + c32 + l4;
            for_rows([&](int rr, bool ok) {
              const float4 t = *reinterpret_cast<const float4*>(T + rr * kTP + l4);
              if (ok) *reinterpret_cast<uint2*>(op + roff_out[rr]) = make_uint2(pack_half2(t.x, t.y), pack_half2(t.z, t.w));
            });
            __syncwarp();
          }
        }
        if (LNF && p.stats_out && valid && !st_first) {
          const float n = (float)kColsPerGrp;
          const float mean_p = st_pivot + st_s1 / n;
          const float m2_p = fmaxf(st_s2 - st_s1 * st_s1 / n, 0.f);
          const long long orow = out_off / p.ldc;
          reinterpret_cast<float2*>(p.stats_out)[orow * (p.tiles_n * kGroups) + nt * kGroups + grp] = make_float2(mean_p, m2_p);
        }
      }
}

template <int BN>
__global__ void __launch_bounds__(kThreads, 1)
gemm_f16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const GemmArgs p) {
  using Cfg = GemmCfg<BN>;
  constexpr int kStages = Cfg::kStages;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;
  uint8_t* sB = smem + kStages * Cfg::kABytes;
  uint8_t* staging = smem + kStages * Cfg::kStageBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(staging + Cfg::kStagingBytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kStages;
  uint64_t* tmem_full = bars + 2 * kStages;
  uint64_t* tmem_empty = bars + 2 * kStages + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  pdl_launch_dependents();

  if (warp == 0 && lane == 0) {
    if (smem_u32(smem) & 1023) __trap();   // 128B-swizzle atoms need a 1024 B aligned base
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], kEpiWarps);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<Cfg::kTmemCols>(tmem_ptr);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr;
  pdl_wait();   // everything above overlapped the previous kernel's tail

  const int num_tiles = p.tiles_m * p.tiles_n;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    // (whole warp in the control flow, one elected lane issues: keeps addresses/descriptors in uniform
    //  registers -- under a lane-0 branch ptxas wraps each TMA / MMA in an ELECT+R2UR waterfall loop)
    const bool elected = elect_one();
    {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int mt = tile / p.tiles_n;
        const int nt = tile % p.tiles_n;
        int cb = 0, cy = 0, cx = 0;
        if (p.a_mode == UDB_A_CONV3X3) {
          const int per_img = p.conv_tx * p.conv_ty;
          cb = mt / per_img;
          const int r = mt % per_img;
          cy = (r / p.conv_tx) * p.conv_TH + p.conv_off;
          cx = (r % p.conv_tx) * p.conv_TW + p.conv_off;
        }
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (elected) {
            mbar_arrive_expect_tx(&full_bar[stage], Cfg::kStageBytes);
            if (p.a_mode == UDB_A_CONV3X3) {
              const int tap = kb / p.conv_cpb;
              const int c0 = p.conv_coff + (kb % p.conv_cpb) * BK;
              tma_load_4d(sA + stage * Cfg::kABytes, &tmA, &full_bar[stage], c0, cx + tap % 3,
                          cy + tap / 3, cb);
            } else {
              int ka = kb * BK;
              if (p.a_wrap && ka >= p.a_wrap) ka -= p.a_wrap;
              tma_load_2d(sA + stage * Cfg::kABytes, &tmA, &full_bar[stage], ka, mt * BM);
            }
            tma_load_2d(sB + stage * Cfg::kBBytes, &tmB, &full_bar[stage], kb * BK, nt * BN);
          }
          __syncwarp();
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    const bool elected = elect_one();
    {
      constexpr uint32_t idesc = umma_idesc_f16(BM, BN, false, false);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        mbar_wait(&tmem_empty[as], aphase ^ 1);
        tc_fence_after_sync();
        const uint32_t d_tmem = tmem_base + as * BN;
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after_sync();
          const uint64_t da = umma_desc_sw128(smem_u32(sA + stage * Cfg::kABytes), 16, 1024);
          const uint64_t db = umma_desc_sw128(smem_u32(sB + stage * Cfg::kBBytes), 16, 1024);
          if (elected) {
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) {
              // advance 16 elements (32 B) along K inside the 128 B swizzle atom: +2 in 16-byte units
              umma_f16_ss(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
            }
            if (kb == p.num_kb - 1) umma_commit(&tmem_full[as]);   // before the stage release: see gemm2_f16_kernel
            umma_commit(&empty_bar[stage]);
          }
          __syncwarp();
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue
    const int ew = warp - 4;
    const int quad = warp & 3;            // TMEM lane quadrant this warp may access
    const int grp = ew >> 2;              // column half
    const int r_in_tile = quad * 32 + lane;
    float* T = reinterpret_cast<float*>(staging) + ew * 32 * kTP;                   // [32][kTP]
    uint32_t* roff_out = reinterpret_cast<uint32_t*>(staging + kEpiWarps * 32 * kTP * 4) + ew * 64;
    uint32_t* roff_res = roff_out + 32;
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      const int mt = tile / p.tiles_n;
      const int nt = tile % p.tiles_n;
      mbar_wait(&tmem_full[as], aphase);
      tc_fence_after_sync();
      EpiWarp ew_ctx{T, roff_out, roff_res, quad, grp, lane, r_in_tile};
      epilogue_tile<BN>(p, ew_ctx, tmem_base + as * BN, mt, nt);
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[as]);
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after_sync();
    tmem_dealloc<Cfg::kTmemCols>(tmem_base);
  }
}

template <int BN>
static int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmArgs& a, cudaStream_t st) {
  using Cfg = GemmCfg<BN>;
  static std::atomic<uint64_t> attr_mask{0};
  if (first_on_device(attr_mask)) {
    cudaError_t e = cudaFuncSetAttribute(gemm_f16_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg::kSmemBytes);
    if (e != cudaSuccess) {
      set_error("gemm: cudaFuncSetAttribute(%d B smem): %s", Cfg::kSmemBytes, cudaGetErrorString(e));
      return 1;
    }
  }
  const int tiles = a.tiles_m * a.tiles_n;
  const int grid = tiles < num_sms() ? tiles : num_sms();
  cudaError_t e = launch_ex(gemm_f16_kernel<BN>, dim3(grid), dim3(kThreads), Cfg::kSmemBytes, st, 1, tmA, tmB, a);
  if (e != cudaSuccess) {
    set_error("gemm_f16_kernel launch: %s", cudaGetErrorString(e));
    return 1;
  }
  return check_launch("gemm_f16_kernel");
}


// ---------------------------------------------------------------------------------------------
// CTA-pair variant (tcgen05 cta_group::2): a cluster of two CTAs computes a 256 x BN tile.  Each CTA
// loads its own 128 rows of A and BN/2 rows of W per k-block (so per-SM operand traffic drops to
// 16 KB + BN*64 B per 64-deep step and the ring holds 6 stages instead of 4), the leader issues
// M=256 MMAs for both, each CTA drains its own 128 accumulator rows from its own TMEM.
// ---------------------------------------------------------------------------------------------
template <int BN>
struct Gemm2Cfg {
  static constexpr int kStages = BN >= 256 ? 5 : (BN >= 192 ? 6 : 7);
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBBytes = (BN / 2) * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kTmemCols = 2 * BN <= 256 ? 256 : 512;   // power of two (BN = 192: 384 columns used)
  static constexpr int kStagingBytes = kEpiWarps * (32 * kTP * 4 + 2 * 32 * 4);
  static constexpr int kSmemBytes = kStages * kStageBytes + kStagingBytes + 256;
};

template <int BN, bool LNF>
__global__ void __launch_bounds__(kThreads, 1)
gemm2_f16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const GemmArgs p) {
  using Cfg = Gemm2Cfg<BN>;
  constexpr int kStages = Cfg::kStages;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;
  uint8_t* sB = smem + kStages * Cfg::kABytes;
  uint8_t* staging = smem + kStages * Cfg::kStageBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(staging + Cfg::kStagingBytes);
  uint64_t* full_bar = bars;                    // leader's copies are the live ones (count 2)
  uint64_t* empty_bar = bars + kStages;         // per CTA, arrived by the multicast commit
  uint64_t* tmem_full = bars + 2 * kStages;     // per CTA, multicast commit
  uint64_t* tmem_empty = bars + 2 * kStages + 2;  // leader's copies: 2 x kEpiWarps arrivals
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int pair = blockIdx.x >> 1;
  const int num_pairs = gridDim.x >> 1;
  pdl_launch_dependents();

  if (warp == 0 && lane == 0) {
    if (smem_u32(smem) & 1023) __trap();
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], 2);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 2 * kEpiWarps);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc_2cta<Cfg::kTmemCols>(tmem_ptr);
  tc_fence_before_sync();
  cluster_sync_all();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr;
  pdl_wait();   // everything above overlapped the previous kernel's tail

  const int tiles_m2 = (p.tiles_m + 1) >> 1;       // 256-row (two 128-row blocks) tiles
  const int num_tiles = tiles_m2 * p.tiles_n;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (both CTAs)
    const bool elected = elect_one();
    {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs) {
        const int mt = 2 * (tile / p.tiles_n) + (int)rank;   // this CTA's 128-row block
        const int nt = tile % p.tiles_n;
        int cb = 0, cy = 0, cx = 0;
        if (p.a_mode == UDB_A_CONV3X3) {
          const int per_img = p.conv_tx * p.conv_ty;
          cb = mt / per_img;                                   // may be == conv_B: fully out of bounds -> zeros
          const int r = mt % per_img;
          cy = (r / p.conv_tx) * p.conv_TH + p.conv_off;
          cx = (r % p.conv_tx) * p.conv_TW + p.conv_off;
        }
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (elected) {
            if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * Cfg::kStageBytes);
            else mbar_arrive_remote(&full_bar[stage], 0);
            if (p.a_mode == UDB_A_CONV3X3) {
              const int tap = kb / p.conv_cpb;
              const int c0 = p.conv_coff + (kb % p.conv_cpb) * BK;
              tma_load_4d_2sm(sA + stage * Cfg::kABytes, &tmA, &full_bar[stage], c0, cx + tap % 3, cy + tap / 3, cb);
            } else {
              int ka = kb * BK;
              if (p.a_wrap && ka >= p.a_wrap) ka -= p.a_wrap;
              tma_load_2d_2sm(sA + stage * Cfg::kABytes, &tmA, &full_bar[stage], ka, mt * BM);
            }
            tma_load_2d_2sm(sB + stage * Cfg::kBBytes, &tmB, &full_bar[stage], kb * BK, nt * BN + (int)rank * (BN / 2));
          }
          __syncwarp();
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA only)
    if (leader) {
      const bool elected = elect_one();
      constexpr uint32_t idesc = umma_idesc_f16(2 * BM, BN, false, false);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs, ++it) {
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        mbar_wait(&tmem_empty[as], aphase ^ 1);
        tc_fence_after_sync();
        const uint32_t d_tmem = tmem_base + as * BN;
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after_sync();
          const uint64_t da = umma_desc_sw128(smem_u32(sA + stage * Cfg::kABytes), 16, 1024);
          const uint64_t db = umma_desc_sw128(smem_u32(sB + stage * Cfg::kBBytes), 16, 1024);
          if (elected) {
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              umma_f16_ss_2cta(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
            // one thread's commits reach their barriers one after the other (~130 clk apart, tools/issue_probe.cu): the
            // accumulator hand-over the epilogue is waiting for goes before the release of the last operand stage
            if (kb == p.num_kb - 1) umma_commit_2cta(&tmem_full[as]);
            umma_commit_2cta(&empty_bar[stage]);
          }
          __syncwarp();
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue (both CTAs)
    const int ew = warp - 4;
    EpiWarp ctx{reinterpret_cast<float*>(staging) + ew * 32 * kTP,
                reinterpret_cast<uint32_t*>(staging + kEpiWarps * 32 * kTP * 4) + ew * 64,
                reinterpret_cast<uint32_t*>(staging + kEpiWarps * 32 * kTP * 4) + ew * 64 + 32,
                warp & 3, ew >> 2, lane, (warp & 3) * 32 + lane};
    int it = 0;
    for (int tile = pair; tile < num_tiles; tile += num_pairs, ++it) {
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      const int mt = 2 * (tile / p.tiles_n) + (int)rank;
      const int nt = tile % p.tiles_n;
      mbar_wait(&tmem_full[as], aphase);
      tc_fence_after_sync();
      epilogue_tile<BN, LNF>(p, ctx, tmem_base + as * BN, mt, nt);
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) {
        if (leader) mbar_arrive(&tmem_empty[as]);
        else mbar_arrive_remote(&tmem_empty[as], 0);
      }
    }
  }

  tc_fence_before_sync();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after_sync();
    tmem_dealloc_2cta<Cfg::kTmemCols>(tmem_base);
  }
}

template <int BN, bool LNF>
static int launch_gemm2(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmArgs& a, cudaStream_t st) {
  using Cfg = Gemm2Cfg<BN>;
  static std::atomic<uint64_t> attr_mask{0};
  if (first_on_device(attr_mask)) {
    cudaError_t e = cudaFuncSetAttribute(gemm2_f16_kernel<BN, LNF>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg::kSmemBytes);
    if (e != cudaSuccess) {
      set_error("gemm2: cudaFuncSetAttribute(%d B smem): %s", Cfg::kSmemBytes, cudaGetErrorString(e));
      return 1;
    }
  }
  const int tiles = ((a.tiles_m + 1) / 2) * a.tiles_n;
  const int max_pairs = num_sms() / 2;
  const int pairs = tiles < max_pairs ? tiles : max_pairs;
  cudaError_t e = launch_ex(gemm2_f16_kernel<BN, LNF>, dim3(2 * pairs), dim3(kThreads), Cfg::kSmemBytes, st, 2, tmA, tmB, a);
  if (e != cudaSuccess) {
    set_error("gemm2_f16_kernel launch: %s", cudaGetErrorString(e));
    return 1;
  }
  return check_launch("gemm2_f16_kernel");
}

}  // namespace udb

extern "C" int udb_gemm_f16(const udb_gemm_t* g, void* stream) {
  using namespace udb;
  if (!g || !g->a || !g->w) { set_error("udb_gemm_f16: null operand"); return 1; }
  if (g->N % 32 != 0) { set_error("udb_gemm_f16: N=%d must be a multiple of 32", g->N); return 1; }
  GemmArgs a{};
  a.M = g->M; a.N = g->N; a.K = g->K;
  a.num_kb = (g->K + BK - 1) / BK;
  a.a_mode = g->a_mode;
  a.bias = g->bias; a.gamma = g->gamma; a.resid = g->resid; a.resid_f32 = g->resid_f32;
  a.out = g->out; a.out_f32 = g->out_f32; a.out2 = reinterpret_cast<__half*>(g->out2); a.out2_leaky = g->out2_leaky;
  a.act = g->act; a.store_mode = g->store_mode;
  a.ldc = g->ldc; a.ldr = g->ldr > 0 ? g->ldr : g->ldc;
  a.rpg = g->rows_per_group; a.gstride = g->group_stride; a.roff = g->row_offset;
  a.resid_mod = g->resid_mod; a.resid_roff = g->resid_row_offset;
  a.ct_k = g->ct_k; a.ct_cout = g->ct_cout; a.ct_h = g->ct_h; a.ct_w = g->ct_w; a.ct_pad = g->ct_pad;
  a.conv_coff = g->conv_coff;
  a.head_w = g->head_w; a.head_b = g->head_b; a.head_add = g->head_add;
  a.a_wrap = 0;
  a.out_split = g->out_split;
  if (g->a_split_k > 0) {
    if (g->a_mode != UDB_A_MATRIX || g->a_split_k % BK || g->K != 3 * g->a_split_k || g->lda < 2 * g->a_split_k) {
      set_error("udb_gemm_f16: split operands need a_mode MATRIX, K1 %% 64 == 0, K == 3*K1, lda >= 2*K1 (K1=%d K=%d lda=%d)",
                g->a_split_k, g->K, g->lda);
      return 1;
    }
    a.a_wrap = 2 * g->a_split_k;
  }
  a.stats_out = g->ln_stats_out; a.ln_stats = g->ln_stats_in; a.ln_c1 = g->ln_c1; a.ln_eps = g->ln_eps;
  a.ln_parts = g->ln_parts; a.ln_part_cols = g->ln_part_cols;
  if ((g->ln_stats_out || g->ln_stats_in) && g->store_mode != UDB_STORE_ROWS) {
    set_error("udb_gemm_f16: fused LayerNorm statistics need the ROWS store"); return 1;
  }
  if (g->ln_stats_in && (!g->ln_c1 || g->ln_parts <= 0 || g->ln_part_cols <= 0 || g->rows_per_group > 0)) {
    set_error("udb_gemm_f16: ln_stats_in needs ln_c1, ln_parts, ln_part_cols and an identity row map"); return 1;
  }
  if (g->out_split && (g->out_f32 || !g->out || g->store_mode != UDB_STORE_ROWS)) {
    set_error("udb_gemm_f16: out_split needs an f16 `out` with the ROWS store"); return 1;
  }

  // CTA-pair kernel (cta_group::2) for the wide tiles; UDB_GEMM_PAIR=0 forces the single-CTA kernel
  static const int pair_env = [] {
    const char* e = getenv("UDB_GEMM_PAIR");
    return e ? atoi(e) : 1;
  }();
  // UDB_GEMM_BN_CAP (experiments): upper bound on the tile width
  static const int bn_cap = [] {
    const char* e = getenv("UDB_GEMM_BN_CAP");
    return e ? atoi(e) : 256;
  }();
  // tile width: widest that divides the work sensibly
  int bn;
  if (g->store_mode == UDB_STORE_HEAD) {
    if (g->N != 32) { set_error("udb_gemm_f16: HEAD store needs N == 32"); return 1; }
    bn = 32;
  } else if (g->N % 256 == 0 && bn_cap >= 256) bn = 256;
  else if (g->N % 192 == 0 && pair_env != 0 && bn_cap >= 192) bn = 192;      // ConvNeXt widths 192 / 384: 3 (or 6) times fewer passes over A than 64 / 128
  else if (g->N % 128 == 0) bn = 128;
  else if (g->N % 64 == 0) bn = 64;
  else bn = 32;
  if (g->store_mode == UDB_STORE_CONVT && (g->ct_cout % 32) != 0) {
    set_error("udb_gemm_f16: CONVT needs Cout %% 32 == 0"); return 1;
  }
  // Experiment (UDB_GEMM_BN_AUTO=1, off by default).  Wave quantisation: the persistent grid has num_sms/2 CTA pairs and a
  // static round-robin tile list, so a launch takes ceil(tiles / pairs) tile times; when 128-wide tiles need fewer
  // column-units of time than 256-wide ones (the 12888 x 3072 qkv GEMM: 9 waves x 256 vs 17 waves x 128) take the narrower
  // tile.  Measured on the B200 (profiles/r02_gemm_bn_auto_ab.txt): ViT-L GEMM time 10.99 ms instead of 10.19 -- the
  // 128-wide tile's main loop (A re-streamed twice as often, half the MMA N) loses more than the shorter tail gains;
  // ConvNeXt-L: 13.41 vs 13.52 ms.
  static const int bn_auto = [] {
    const char* e = getenv("UDB_GEMM_BN_AUTO");
    return e ? atoi(e) : 0;
  }();
  if (bn_auto && bn == 256 && pair_env != 0 && g->store_mode != UDB_STORE_HEAD && !g->ln_stats_out) {
    const long long pairs = num_sms() / 2;
    const long long tm2 = ((g->a_mode == UDB_A_CONV3X3 ? 0 : (g->M + BM - 1) / BM) + 1) / 2;
    if (tm2 > 0) {
      const long long w256 = (tm2 * (g->N / 256) + pairs - 1) / pairs, w128 = (tm2 * (g->N / 128) + pairs - 1) / pairs;
      if (w128 * 128 * 100 < w256 * 256 * 97) bn = 128;
    }
  }
  a.tiles_n = (g->N + bn - 1) / bn;
  if (g->ln_stats_out) {
    const int groups = bn >= 64 ? 2 : 1;
    if (g->N % bn || g->ln_parts != a.tiles_n * groups || g->ln_part_cols != bn / groups) {
      set_error("udb_gemm_f16: ln_stats_out with N=%d runs as %d parts of %d columns (caller said %d x %d)", g->N, a.tiles_n * groups,
                bn / groups, g->ln_parts, g->ln_part_cols);
      return 1;
    }
  }
  const bool use_pair = pair_env != 0 && (bn == 256 || bn == 192 || bn == 128);

  CUtensorMap tmA, tmB;
  if (g->a_mode == UDB_A_CONV3X3) {
    if (g->conv_C % BK != 0 || g->K != 9 * g->conv_C) {
      set_error("udb_gemm_f16: conv3x3 needs C %% 64 == 0 and K == 9*C (C=%d K=%d)", g->conv_C, g->K);
      return 1;
    }
    if (g->store_mode != UDB_STORE_CONVTILE && g->store_mode != UDB_STORE_HEAD) {
      set_error("udb_gemm_f16: conv3x3 operand needs a CONVTILE or HEAD store"); return 1;
    }
    const int TH = g->conv_TH > 0 ? g->conv_TH : 8, TW = g->conv_TW > 0 ? g->conv_TW : 16;
    if (TH * TW != BM) { set_error("udb_gemm_f16: conv tile %dx%d != 128 pixels", TH, TW); return 1; }
    a.conv_B = g->conv_B; a.conv_H = g->conv_H; a.conv_W = g->conv_W; a.conv_cpb = g->conv_C / BK; a.conv_off = g->conv_off;
    a.conv_TH = TH; a.conv_TW = TW;
    a.conv_tx = (g->conv_W + TW - 1) / TW; a.conv_ty = (g->conv_H + TH - 1) / TH;
    a.tiles_m = g->conv_B * a.conv_tx * a.conv_ty;
    a.M = a.tiles_m * BM;
    const uint64_t cs = g->conv_cstride > 0 ? g->conv_cstride : g->conv_C;
    if ((uint64_t)g->conv_coff + g->conv_C > cs || (g->conv_coff % 8)) {
      set_error("udb_gemm_f16: bad conv channel slice (off %d, C %d, stride %d)", g->conv_coff, g->conv_C, (int)cs);
      return 1;
    }
    const uint64_t dims[4] = {cs, (uint64_t)g->conv_inW, (uint64_t)g->conv_inH, (uint64_t)g->conv_B};
    const uint64_t str[3] = {cs * 2, (uint64_t)g->conv_inW * cs * 2, (uint64_t)g->conv_inH * g->conv_inW * cs * 2};
    const uint32_t box[4] = {(uint32_t)BK, (uint32_t)TW, (uint32_t)TH, 1};
    if (make_tmap_f16(&tmA, g->a, 4, dims, str, box, true)) return 1;
  } else {
    a.tiles_m = (g->M + BM - 1) / BM;
    if (g->store_mode == UDB_STORE_CONVTILE || g->store_mode == UDB_STORE_HEAD) {
      set_error("udb_gemm_f16: tile store modes need a_mode == CONV3X3"); return 1;
    }
    const uint64_t dims[2] = {(uint64_t)(g->a_split_k > 0 ? 2 * g->a_split_k : g->K), (uint64_t)g->M};
    const uint64_t str[1] = {(uint64_t)g->lda * 2};
    const uint32_t box[2] = {(uint32_t)BK, (uint32_t)BM};
    if (make_tmap_f16(&tmA, g->a, 2, dims, str, box, true)) return 1;
  }
  {
    const uint64_t dims[2] = {(uint64_t)g->K, (uint64_t)g->N};
    const uint64_t str[1] = {(uint64_t)g->ldw * 2};
    const uint32_t box[2] = {(uint32_t)BK, (uint32_t)(use_pair ? bn / 2 : bn)};
    if (make_tmap_f16(&tmB, g->w, 2, dims, str, box, true)) return 1;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  note_work(2.0 * a.M * (double)g->N * g->K,
            2.0 * ((double)a.M * (g->a_mode == UDB_A_CONV3X3 ? g->conv_C : (g->a_split_k ? 2 * g->a_split_k : g->K)) + (double)g->N * g->K) +
                (double)a.M * g->N * ((g->out ? (g->out_f32 ? 4 : 2) : 0) + (g->out2 ? 2 : 0) + (g->resid ? (g->resid_f32 ? 4 : 2) : 0)));
  const bool lnf = g->ln_stats_out || g->ln_stats_in;
  if (lnf && !use_pair) { set_error("udb_gemm_f16: fused LayerNorm needs the CTA-pair kernel (N %% 128 == 0)"); return 1; }
  if (use_pair) {
    if (lnf) return bn == 256 ? launch_gemm2<256, true>(tmA, tmB, a, st) : (bn == 192 ? launch_gemm2<192, true>(tmA, tmB, a, st)
                                                                                      : launch_gemm2<128, true>(tmA, tmB, a, st));
    return bn == 256 ? launch_gemm2<256, false>(tmA, tmB, a, st) : (bn == 192 ? launch_gemm2<192, false>(tmA, tmB, a, st)
                                                                               : launch_gemm2<128, false>(tmA, tmB, a, st));
  }
  switch (bn) {
    case 256: return launch_gemm<256>(tmA, tmB, a, st);
    case 128: return launch_gemm<128>(tmA, tmB, a, st);
    case 64: return launch_gemm<64>(tmA, tmB, a, st);
    default: return launch_gemm<32>(tmA, tmB, a, st);
  }
}
