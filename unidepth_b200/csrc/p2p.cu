// Peer-memory plumbing for the multi-GPU output gather (unidepth_b200/parallel.py): buffers that
// every rank of the node can read directly over NVLink (CUDA IPC), a device-side barrier through flags in peer memory,
// and copy-engine pulls.  No SM does any data movement: the GEMM kernels are persistent with one CTA per SM, so a
// collective kernel that occupies even a few SMs stalls whole tile columns (measured in round 1), while DMA copies are free.
#include <string.h>

#include "common.h"

namespace udb {

// One block, `world` threads.  Thread p publishes this rank's arrival in peer p's flag array, then waits until peer p has
// published in ours.  Flags only grow (epoch numbers), so no reset is needed.  Bounded spin: a missing peer makes the
// kernel give up after ~4 s and raise *timeout_flag instead of hanging the GPU.
__global__ void p2p_barrier_kernel(unsigned int* const* peer_flags, volatile unsigned int* my_flags, int rank, int world, unsigned int epoch,
                                   int* timeout_flag) {
  const int p = threadIdx.x;
  if (p >= world) return;
  __threadfence_system();                       // everything this rank wrote before the barrier is visible to its peers
  if (p != rank) {
    volatile unsigned int* dst = peer_flags[p] + rank;
    *dst = epoch;
    __threadfence_system();
    const long long t0 = clock64();
    while (my_flags[p] < epoch) {
      if (clock64() - t0 > 8000000000LL) { *timeout_flag = 1; break; }
      __nanosleep(200);
    }
  }
  __threadfence_system();
}

}  // namespace udb

using namespace udb;

extern "C" {

int udb_p2p_alloc(size_t bytes, void** dev_ptr, void* handle64) {
  if (!dev_ptr || !handle64 || bytes == 0) { set_error("udb_p2p_alloc: bad argument"); return 1; }
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, bytes);
  if (e != cudaSuccess) { set_error("udb_p2p_alloc: cudaMalloc(%zu): %s", bytes, cudaGetErrorString(e)); return 1; }
  cudaMemset(p, 0, bytes);
  cudaIpcMemHandle_t h;
  e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) { cudaFree(p); set_error("udb_p2p_alloc: cudaIpcGetMemHandle: %s", cudaGetErrorString(e)); return 1; }
  memcpy(handle64, &h, 64);
  *dev_ptr = p;
  return 0;
}

int udb_p2p_open(const void* handle64, void** peer_ptr) {
  if (!handle64 || !peer_ptr) { set_error("udb_p2p_open: bad argument"); return 1; }
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  cudaError_t e = cudaIpcOpenMemHandle(peer_ptr, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) { set_error("udb_p2p_open: cudaIpcOpenMemHandle: %s", cudaGetErrorString(e)); return 1; }
  return 0;
}

int udb_p2p_close(void* peer_ptr) { return cudaIpcCloseMemHandle(peer_ptr) == cudaSuccess ? 0 : 1; }
int udb_p2p_free(void* dev_ptr) { return cudaFree(dev_ptr) == cudaSuccess ? 0 : 1; }

int udb_p2p_barrier(void* const* peer_flags_dev, void* my_flags, int32_t rank, int32_t world, uint32_t epoch, int32_t* timeout_flag_dev,
                    void* stream) {
  if (world < 1 || world > 32) { set_error("udb_p2p_barrier: world %d out of range", world); return 1; }
  p2p_barrier_kernel<<<1, 32, 0, reinterpret_cast<cudaStream_t>(stream)>>>(reinterpret_cast<unsigned int* const*>(peer_flags_dev),
                                                                           reinterpret_cast<volatile unsigned int*>(my_flags), rank, world, epoch,
                                                                           timeout_flag_dev);
  return check_launch("p2p_barrier_kernel");
}

int udb_p2p_copy(void* dst, const void* src, size_t bytes, void* stream) {
  cudaError_t e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, reinterpret_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) { set_error("udb_p2p_copy: %s", cudaGetErrorString(e)); return 1; }
  return 0;
}

}  // extern "C"
