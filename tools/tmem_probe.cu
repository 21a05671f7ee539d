// Micro-benchmark (not part of the product): tcgen05.ld / MUFU.EX2 throughput on one SM as a function of the number of warps.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o variants/tmem_probe tools/tmem_probe.cu
// mode 0: W warps each issue R x (tcgen05.ld.32x32b.x32 + wait::ld) on their lane quadrant
// mode 1: W warps each issue R x 32 MUFU.EX2 (dependent only through a running sum)
// mode 2: both interleaved (32 scores loaded, 32 ex2) -- the attention kernel's inner pattern without the rest
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../unidepth_b200/csrc/ptx.cuh"
using namespace udb;

__device__ __forceinline__ float ex2f(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

template <int MODE>
__global__ void __launch_bounds__(512) probe(long long* out, float* sink, int R) {
  __shared__ uint32_t tptr;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc<512>(&tptr);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = tptr;
  const uint32_t addr = tmem + (static_cast<uint32_t>((warp & 3) * 32) << 16) + (warp >> 2) * 64;
  float acc = 0.f;
  long long t0 = 0;
  for (int pass = 0; pass < 2; ++pass) {
    __syncthreads();
    t0 = clock64();
    for (int r = 0; r < R; ++r) {
      uint32_t v[32];
      if (MODE != 1) {
        tmem_ld_32x32b_x32(addr + (r & 1) * 32, v);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(acc + i);
      }
      if (MODE != 0) {
#pragma unroll
        for (int i = 0; i < 32; ++i) acc += ex2f(__uint_as_float(v[i]) * 1e-30f);
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) acc += __uint_as_float(v[i]) * 1e-30f;
      }
    }
  }
  const long long t1 = clock64();
  __syncthreads();
  if ((threadIdx.x & 31) == 0) out[warp] = t1 - t0;
  sink[threadIdx.x] = acc;
  if (warp == 0) tmem_dealloc<512>(tmem);
}

int main() {
  long long* out; float* sink;
  cudaMallocManaged(&out, 16 * sizeof(long long));
  cudaMalloc(&sink, 512 * sizeof(float));
  const int R = 2000;
  for (int mode = 0; mode < 3; ++mode)
    for (int W : {1, 2, 4, 8, 16}) {
      if (mode == 0) probe<0><<<1, 32 * W>>>(out, sink, R);
      if (mode == 1) probe<1><<<1, 32 * W>>>(out, sink, R);
      if (mode == 2) probe<2><<<1, 32 * W>>>(out, sink, R);
      cudaError_t e = cudaDeviceSynchronize();
      long long mx = 0;
      for (int w = 0; w < W; ++w) mx = out[w] > mx ? out[w] : mx;
      const double per_iter = (double)mx / R;
      printf("mode %d warps %2d: %8.1f clk per iteration per warp  -> SM-wide %.1f B/clk of TMEM reads, %.2f MUFU/clk  (%s)\n", mode, W,
             per_iter, mode != 1 ? W * 4096.0 / per_iter : 0.0, mode != 0 ? W * 32 * 32.0 / per_iter : 0.0, cudaGetErrorString(e));
    }
  return 0;
}
