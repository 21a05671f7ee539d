// Micro-benchmark (not part of the product): where does the issuing thread of a tcgen05.mma burst spend its time?
// clock64() stamps inside the elected lane: before the first MMA, after each MMA, after each commit, after the warp re-converges.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o variants/issue_probe tools/issue_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../unidepth_b200/csrc/ptx.cuh"
using namespace udb;

template <int NMMA, int NCOMMIT>
__global__ void __launch_bounds__(128) probe(long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sB = sm;                  // 64 x 64 halves
  uint64_t* bar = reinterpret_cast<uint64_t*>(sm + 16384);   // [4]
  uint32_t* tptr = reinterpret_cast<uint32_t*>(bar + 4);
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 16384 / 4; i += 128) reinterpret_cast<uint32_t*>(sm)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { for (int i = 0; i < 4; ++i) mbar_init(&bar[i], 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc<256>(tptr);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *tptr;
  if (warp == 1) {
    const bool elected = elect_one();
    constexpr uint32_t idesc = umma_idesc_f16(128, 64, false, true);
    const uint64_t db = umma_desc_sw128(smem_u32(sB), 1024, 1024);
    long long ts[16];
    for (int pass = 0; pass < 3; ++pass) {   // passes 0, 1 = warm-up
      __syncwarp();
      const long long t0 = clock64();
      int n = 0;
      if (elected) {
#pragma unroll
        for (int r = 0; r < NMMA; ++r) {
          umma_f16_ts(tmem, tmem + 64 + 8 * (r & 3), db + (uint64_t)((r & 3) * 2048 >> 4), idesc, r != 0);
          ts[n++] = clock64();
        }
#pragma unroll
        for (int c = 0; c < NCOMMIT; ++c) {
          umma_commit(&bar[c]);
          ts[n++] = clock64();
        }
      }
      __syncwarp();
      const long long t1 = clock64();
      for (int c = 0; c < NCOMMIT; ++c) mbar_wait(&bar[c], pass & 1);
      tc_fence_after_sync();
      const long long t2 = clock64();
      if (elected && pass == 2) {
        for (int i = 0; i < NMMA + NCOMMIT; ++i) out[i] = ts[i] - t0;
        out[14] = t1 - t0;
        out[15] = t2 - t0;
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc<256>(tmem);
}

template <int NMMA, int NCOMMIT>
void run() {
  long long* d;
  cudaMalloc(&d, 16 * sizeof(long long));
  cudaMemset(d, 0, 16 * sizeof(long long));
  const int smem = 16384 + 64 + 1024;
  probe<NMMA, NCOMMIT><<<1, 128, smem>>>(d);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[16];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  printf("%d MMA (TS 128x64x16) + %d commit:", NMMA, NCOMMIT);
  for (int i = 0; i < NMMA; ++i) printf(" m%lld", h[i]);
  for (int i = NMMA; i < NMMA + NCOMMIT; ++i) printf(" c%lld", h[i]);
  printf("  | reconverged %lld, barriers seen %lld  (%s)\n", h[14], h[15], cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  run<1, 1>(); run<2, 1>(); run<4, 1>(); run<8, 1>(); run<4, 2>(); run<4, 3>(); run<0, 1>(); run<6, 2>();
  return 0;
}
