// 3x3 convolution with few output channels (Cout = 32 / 64) over a pre-padded NHWC f16 image, as an
// implicit GEMM whose A operand is loaded ONCE per tile: the (16+2) x 16-pixel halo of a 16 x 8-pixel
// output tile is brought into shared memory by one TMA box per 64-channel slab and all nine filter
// taps read it through shifted UMMA descriptors (start address + (dy*16 + dx) pixel rows, 8-pixel row
// groups 2 KB apart).  The generic implicit-GEMM path (gemm.cu, UDB_A_CONV3X3) re-fetches the tile
// for every tap and is L2-bandwidth bound when Cout is small (each 16 KB A tile feeds only 32-64
// output columns); here the per-tile L2 traffic drops from 9 x 16 KB to 36 KB per slab and the
// weights (<= 147 KB) stay resident in shared memory for the whole persistent kernel.
//
// Covers the decoder heads (reference unidepth/models/unidepthv2/decoder.py:200-229, 288-313):
//   to_*_lr  (128 -> 64, reflect)                       -> f16 NHWC out
//   to_*_hr  (64 -> 32, reflect) + LeakyReLU + 1x1 + clip + exp  -> f32 plane
#include <stdlib.h>

#include "common.h"
#include "ptx.cuh"

namespace udb {

constexpr int HC_TH = 16, HC_TW = 8;          // output tile (128 pixels = UMMA M)
constexpr int HC_HW = 16, HC_HH = HC_TH + 2;  // halo box: 16 x 18 pixels (10 of the 16 columns are used)
constexpr int HC_SLAB_BYTES = HC_HW * HC_HH * 128;   // 36864
constexpr int HC_THREADS = 256;

struct HaloArgs {
  int B, H, W;            // output size (input is [B, H+2, W+2, cstride])
  int slabs;              // C / 64
  int coff;               // first input channel
  int tiles_x, tiles_y, num_tiles;
  int a_stages;           // slab buffers in the ring
  int base_off_mode;      // 1: descriptor base_offset = dx ; 0: base_offset = 0
  const float* bias;      // [COUT]
  int act;                // UDB_ACT_*
  // f16 NHWC output (ldc elements per pixel) or fused head (f32 plane)
  __half* out;
  long long ldc;
  const float* head_w;
  float head_b, head_add;
  float* head_out;
};

template <int COUT>
__global__ void __launch_bounds__(HC_THREADS, 1)
conv3x3_halo_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW, const HaloArgs p) {
  constexpr int kWTile = COUT * 128;                 // one (tap, slab) weight tile
  extern __shared__ __align__(1024) uint8_t smem[];
  const int w_bytes = 9 * p.slabs * kWTile;
  uint8_t* sW = smem;
  uint8_t* sA = smem + ((w_bytes + 1023) & ~1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sA + p.a_stages * HC_SLAB_BYTES);
  uint64_t* w_full = bars;
  uint64_t* a_full = bars + 1;                       // [a_stages] (<= 4)
  uint64_t* a_empty = bars + 5;
  uint64_t* t_full = bars + 9;                       // [2]
  uint64_t* t_empty = bars + 11;                     // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 13);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_launch_dependents();
  if (threadIdx.x == 0) {
    if (smem_u32(smem) & 1023) __trap();
    prefetch_tmap(&tmX);
    prefetch_tmap(&tmW);
    mbar_init(w_full, 1);
    for (int i = 0; i < p.a_stages; ++i) {
      mbar_init(&a_full[i], 1);
      mbar_init(&a_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&t_full[i], 1);
      mbar_init(&t_empty[i], 4);
    }
    fence_barrier_init();
  }
  constexpr uint32_t kTmemCols = 2 * COUT < 32 ? 32 : 2 * COUT;
  if (warp == 2) tmem_alloc<kTmemCols>(tmem_ptr);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr;
  pdl_wait();

  const int per_img = p.tiles_x * p.tiles_y;
  if (warp == 0) {
    // whole warp in the control flow, one elected lane issues (operands stay in uniform registers;
    // a lane-0 branch makes ptxas wrap every TMA / MMA in an ELECT + R2UR waterfall loop)
    const bool elected = elect_one();
    {
      // weights: resident for the whole kernel
      if (elected) {
        mbar_arrive_expect_tx(w_full, w_bytes);
        for (int kb = 0; kb < 9 * p.slabs; ++kb) tma_load_2d(sW + kb * kWTile, &tmW, w_full, kb * 64, 0);
      }
      __syncwarp();
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        const int b = tile / per_img, r = tile % per_img;
        const int y0 = (r / p.tiles_x) * HC_TH, x0 = (r % p.tiles_x) * HC_TW;
        for (int s = 0; s < p.slabs; ++s) {
          mbar_wait(&a_empty[stage], phase ^ 1);
          if (elected) {
            mbar_arrive_expect_tx(&a_full[stage], HC_SLAB_BYTES);
            tma_load_4d(sA + stage * HC_SLAB_BYTES, &tmX, &a_full[stage], p.coff + s * 64, x0, y0, b);
          }
          __syncwarp();
          if (++stage == p.a_stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    const bool elected = elect_one();
    {
      constexpr uint32_t idesc = umma_idesc_f16(128, COUT, false, false);
      mbar_wait(w_full, 0);
      int stage = 0, it = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
        const int as = it & 1;
        mbar_wait(&t_empty[as], ((it >> 1) & 1) ^ 1);
        tc_fence_after_sync();
        const uint32_t d = tmem_base + as * COUT;
        for (int s = 0; s < p.slabs; ++s) {
          mbar_wait(&a_full[stage], phase);
          tc_fence_after_sync();
          const uint32_t a_base = smem_u32(sA + stage * HC_SLAB_BYTES);
#pragma unroll 1
          for (int tap = 0; tap < 9; ++tap) {
            const int dy = tap / 3, dx = tap % 3;
            // rows of the operand = pixels (y+dy, x+dx): 8-pixel groups, one per tile row, 16 pixel rows apart
            uint64_t da = umma_desc_sw128(a_base + (dy * HC_HW + dx) * 128, 16, HC_HW * 128);
            if (p.base_off_mode) da |= static_cast<uint64_t>(dx) << 49;   // start is dx rows into the 8-row swizzle pattern
            const uint64_t dw = umma_desc_sw128(smem_u32(sW + (tap * p.slabs + s) * kWTile), 16, 1024);
            if (elected) {
#pragma unroll
              for (int k = 0; k < 4; ++k) umma_f16_ss(d, da + 2 * k, dw + 2 * k, idesc, (s | tap | k) != 0 ? 1u : 0u);
            }
          }
          if (elected) {
            if (s == p.slabs - 1) umma_commit(&t_full[as]);   // before the slab release: one thread's commits land in turn
            umma_commit(&a_empty[stage]);
          }
          __syncwarp();
          if (++stage == p.a_stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp >= 4) {
    const int quad = warp & 3;
    const int r_in_tile = quad * 32 + lane;
    const int ty = r_in_tile / HC_TW, tx = r_in_tile % HC_TW;
    int it = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
      const int as = it & 1;
      const int b = tile / per_img, r = tile % per_img;
      const int y = (r / p.tiles_x) * HC_TH + ty, x = (r % p.tiles_x) * HC_TW + tx;
      const bool valid = y < p.H && x < p.W;
      mbar_wait(&t_full[as], (it >> 1) & 1);
      tc_fence_after_sync();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + as * COUT;
      float head_acc = p.head_b;
#pragma unroll
      for (int c = 0; c < COUT; c += 32) {
        uint32_t rr[32];
        tmem_ld_32x32b_x32(t_row + c, rr);
        tmem_ld_wait();
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(rr[j]) + __ldg(p.bias + c + j);
        if (p.act == UDB_ACT_LEAKY) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = leaky(v[j]);
        }
        if (p.head_out) {
#pragma unroll
          for (int j = 0; j < 32; ++j) head_acc = fmaf(v[j], __ldg(p.head_w + c + j), head_acc);
        } else if (valid) {
          uint4* op = reinterpret_cast<uint4*>(p.out + (((long long)b * p.H + y) * p.W + x) * p.ldc + c);
#pragma unroll
          for (int j = 0; j < 4; ++j)
            op[j] = make_uint4(pack_half2(v[8 * j], v[8 * j + 1]), pack_half2(v[8 * j + 2], v[8 * j + 3]),
                               pack_half2(v[8 * j + 4], v[8 * j + 5]), pack_half2(v[8 * j + 6], v[8 * j + 7]));
        }
      }
      if (p.head_out && valid)
        p.head_out[((long long)b * p.H + y) * p.W + x] = expf(fminf(fmaxf(head_acc, -8.0f), 8.0f) + p.head_add);
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&t_empty[as]);
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after_sync();
    tmem_dealloc<kTmemCols>(tmem_base);
  }
}

template <int COUT>
static int launch_halo(const CUtensorMap& tx, const CUtensorMap& tw, const HaloArgs& a, size_t smem, cudaStream_t st) {
  static std::atomic<size_t> set_for[64];
  int dev = 0;
  cudaGetDevice(&dev);
  if (smem > set_for[dev & 63].load()) {
    cudaError_t e = cudaFuncSetAttribute(conv3x3_halo_kernel<COUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("conv3x3_halo: cudaFuncSetAttribute(%zu): %s", smem, cudaGetErrorString(e)); return 1; }
    set_for[dev & 63].store(smem);
  }
  const int grid = a.num_tiles < num_sms() ? a.num_tiles : num_sms();
  note_work(2.0 * a.B * (double)a.H * a.W * COUT * 9 * 64 * a.slabs, 0.0);
  cudaError_t e = launch_ex(conv3x3_halo_kernel<COUT>, dim3(grid), dim3(HC_THREADS), smem, st, 1, tx, tw, a);
  if (e != cudaSuccess) { set_error("conv3x3_halo_kernel launch: %s", cudaGetErrorString(e)); return 1; }
  return check_launch("conv3x3_halo_kernel");
}

}  // namespace udb

extern "C" int udb_conv3x3_halo_f16(const udb_conv_halo_t* c, void* stream) {
  using namespace udb;
  if (c->C % 64 || (c->cout != 32 && c->cout != 64)) { set_error("udb_conv3x3_halo_f16: needs C %% 64 == 0 and Cout in {32, 64}"); return 1; }
  const int cs = c->cstride > 0 ? c->cstride : c->C;
  HaloArgs a{};
  a.B = c->B; a.H = c->H; a.W = c->W;
  a.slabs = c->C / 64; a.coff = c->coff;
  a.tiles_x = (c->W + HC_TW - 1) / HC_TW; a.tiles_y = (c->H + HC_TH - 1) / HC_TH;
  a.num_tiles = c->B * a.tiles_x * a.tiles_y;
  // Measured on B200: the UMMA 128B swizzle is a function of the absolute shared-memory address bits
  // (like TMA's), so a descriptor that starts dx pixel-rows into the 8-row pattern needs NO
  // base_offset; UDB_HALO_BASEOFF=1 (descriptor base_offset = dx) gives wrong results and is kept
  // only as the recorded experiment.
  static const int mode = [] { const char* e = getenv("UDB_HALO_BASEOFF"); return e ? atoi(e) : 0; }();
  a.base_off_mode = mode;
  a.bias = c->bias; a.act = c->act;
  a.out = reinterpret_cast<__half*>(c->out); a.ldc = c->ldc > 0 ? c->ldc : c->cout;
  a.head_w = c->head_w; a.head_b = c->head_b; a.head_add = c->head_add; a.head_out = c->head_out;
  if (a.head_out && c->cout != 32) { set_error("udb_conv3x3_halo_f16: fused head needs Cout == 32"); return 1; }
  const size_t w_bytes = (size_t)9 * a.slabs * c->cout * 128;
  const size_t w_al = (w_bytes + 1023) & ~size_t(1023);
  int stages = (int)((232448 - 256 - w_al) / HC_SLAB_BYTES);
  if (stages > 4) stages = 4;
  if (stages < 2) { set_error("udb_conv3x3_halo_f16: weights too large for shared memory"); return 1; }
  a.a_stages = stages;
  const size_t smem = w_al + (size_t)stages * HC_SLAB_BYTES + 256;
  CUtensorMap tx, tw;
  {
    const uint64_t dims[4] = {(uint64_t)cs, (uint64_t)c->W + 2, (uint64_t)c->H + 2, (uint64_t)c->B};
    const uint64_t str[3] = {(uint64_t)cs * 2, (uint64_t)(c->W + 2) * cs * 2, (uint64_t)(c->H + 2) * (c->W + 2) * cs * 2};
    const uint32_t box[4] = {64, (uint32_t)HC_HW, (uint32_t)HC_HH, 1};
    if (make_tmap_f16(&tx, c->x, 4, dims, str, box, true)) return 1;
  }
  {
    const uint64_t dims[2] = {(uint64_t)9 * c->C, (uint64_t)c->cout};
    const uint64_t str[1] = {(uint64_t)9 * c->C * 2};
    const uint32_t box[2] = {64, (uint32_t)c->cout};
    if (make_tmap_f16(&tw, c->w, 2, dims, str, box, true)) return 1;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  return c->cout == 32 ? launch_halo<32>(tx, tw, a, smem, st) : launch_halo<64>(tx, tw, a, smem, st);
}
