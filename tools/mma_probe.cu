// Micro-benchmark of tcgen05.mma issue / execution rates on one SM (not part of the product):
// one CTA, one elected thread issues R MMAs of a given shape back to back, commits, waits.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/bin/mma_probe tools/mma_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../unidepth_b200/csrc/ptx.cuh"
using namespace udb;

struct Res { long long issue, total; };

// mode 0: SS, B K-major; 1: SS, B MN-major; 2: TS (A in TMEM), B K-major; 3: TS, B MN-major
template <int N, int MODE>
__global__ void __launch_bounds__(128) probe(Res* out, int R) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = sm;                  // 128 x 64 halves, SW128 K-major: 16 KB
  uint8_t* sB = sm + 16384;          // up to 256 x 64 halves: 32 KB
  uint64_t* bar = reinterpret_cast<uint64_t*>(sm + 16384 + 32768);
  uint32_t* tptr = reinterpret_cast<uint32_t*>(bar + 1);
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += 128) reinterpret_cast<uint32_t*>(sm)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc<(N == 256 ? 512 : 256)>(tptr);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *tptr;
  if (warp == 1) {
    const bool elected = elect_one();
    constexpr bool b_mn = (MODE & 1) != 0;
    constexpr uint32_t idesc = umma_idesc_f16(128, N, false, b_mn);
    const uint64_t da = umma_desc_sw128(smem_u32(sA), 16, 1024);
    const uint64_t db = b_mn ? umma_desc_sw128(smem_u32(sB), 1024, 1024) : umma_desc_sw128(smem_u32(sB), 16, 1024);
    long long t0 = 0, t1 = 0;
    for (int pass = 0; pass < 2; ++pass) {   // pass 0 = warm-up
      __syncwarp();
      t0 = clock64();
      if (elected) {
        for (int r = 0; r < R; ++r) {
          const int k = r & 3;
          const uint64_t dbk = b_mn ? db + (uint64_t)(k * 2048 >> 4) : db + 2 * k;
          if (MODE >= 2) umma_f16_ts(tmem, tmem + N + 8 * k, dbk, idesc, r != 0);
          else umma_f16_ss(tmem, da + 2 * k, dbk, idesc, r != 0);
        }
        umma_commit(bar);
      }
      __syncwarp();
      t1 = clock64();
      mbar_wait(bar, pass & 1);
      tc_fence_after_sync();
    }
    const long long t2 = clock64();
    if (elected) { out[blockIdx.x].issue = t1 - t0; out[blockIdx.x].total = t2 - t0; }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc<(N == 256 ? 512 : 256)>(tmem);
}

template <int N, int MODE>
void run(const char* name, int grid, int R) {
  Res* d;
  cudaMalloc(&d, sizeof(Res) * grid);
  const int smem = 16384 + 32768 + 64 + 1024;
  cudaFuncSetAttribute(probe<N, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  probe<N, MODE><<<grid, 128, smem>>>(d, R);
  cudaError_t e = cudaDeviceSynchronize();
  Res h[512];
  cudaMemcpy(h, d, sizeof(Res) * grid, cudaMemcpyDeviceToHost);
  double iss = 0, tot = 0;
  for (int i = 0; i < grid; ++i) { iss += h[i].issue; tot += h[i].total; }
  printf("%-34s grid %3d R %3d: issue %7.1f clk/MMA  total %7.1f clk/MMA  (%s)\n", name, grid, R, iss / grid / R,
         tot / grid / R, cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  for (int R : {1, 2, 4, 6, 8, 16}) {   // issue cost of short bursts (the attention kernels issue 4-8 MMAs per event)
    run<32, 2>("TS 128x32x16  B K-major", 1, R);
    run<64, 3>("TS 128x64x16  B MN-major", 1, R);
    run<64, 0>("SS 128x64x16  B K-major", 1, R);
  }
  for (int grid : {1, 148, 296}) {
    for (int R : {8, 64}) {
      if (grid <= 148) run<256, 0>("SS 128x256x16 B K-major", grid, R);
      run<128, 0>("SS 128x128x16 B K-major", grid, R);
      run<64, 0>("SS 128x64x16  B K-major", grid, R);
      run<64, 1>("SS 128x64x16  B MN-major", grid, R);
      run<128, 1>("SS 128x128x16 B MN-major", grid, R);
      run<128, 2>("TS 128x128x16 B K-major", grid, R);
      run<64, 2>("TS 128x64x16  B K-major", grid, R);
      run<64, 3>("TS 128x64x16  B MN-major", grid, R);
    }
  }
  return 0;
}
