// peer.cu -- peer-memory plumbing for the multi-GPU gather.
//
// The path shards by image with no exchange step; its only collective is the final gather of
// the per-image mask canvases to rank 0 (BASELINE.json north_star).  Besides the NCCL gather
// (sharding.gather_*), the expand kernels can write their output DIRECTLY into rank 0's HBM:
// rank 0 allocates the receive buffer with mrx_peer_alloc, exports a CUDA IPC handle, every
// other rank maps it with mrx_peer_open and passes the mapped address as d_canvas / d_packed
// to mrx_mask_expand / mrx_mask_expand_packed -- the kernels' bulk (TMA) stores then travel over
// NVLink / NVSwitch as they are produced, i.e. compute and gather are one kernel.  Completion is
// signalled on the device: mrx_peer_signal (stream-ordered, after the sender's kernel) bumps a
// flag in rank 0's memory with system-scope release; mrx_peer_wait spins on it on rank 0's
// stream.  No NCCL call, no host round trip inside the timed region.
//
// One process per GPU (torchrun); handles travel through torch.distributed as 64-byte tensors.
#include <string.h>

#include "common.cuh"

namespace mrx {

__global__ void peer_signal_kernel(unsigned int *flag, unsigned int value) {
  __threadfence_system();   // everything this stream wrote before is visible system-wide first
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flag), "r"(value) : "memory");
}

__global__ void peer_wait_kernel(const unsigned int *flags, int n_flags, unsigned int value) {
  const int i = threadIdx.x;
  if (i < n_flags) {
    unsigned int v;
    do {
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flags + i) : "memory");
      if (v < value) __nanosleep(200);
    } while (v < value);
  }
  __syncthreads();
  __threadfence_system();
}

}  // namespace mrx

using namespace mrx;

extern "C" int mrx_peer_alloc(unsigned long long bytes, void **d_ptr) {
  MRX_CHECK_ARG(d_ptr != nullptr && bytes > 0, "mrx_peer_alloc: bad arguments");
  MRX_CUDA(cudaMalloc(d_ptr, static_cast<size_t>(bytes)));
  return MRX_OK;
}

extern "C" int mrx_peer_free(void *d_ptr) {
  if (d_ptr != nullptr) MRX_CUDA(cudaFree(d_ptr));
  return MRX_OK;
}

extern "C" int mrx_peer_export(void *d_ptr, unsigned char *handle64) {
  MRX_CHECK_ARG(d_ptr != nullptr && handle64 != nullptr, "mrx_peer_export: null pointer");
  static_assert(sizeof(cudaIpcMemHandle_t) == MRX_PEER_HANDLE_BYTES, "IPC handle size");
  cudaIpcMemHandle_t h;
  MRX_CUDA(cudaIpcGetMemHandle(&h, d_ptr));
  memcpy(handle64, &h, sizeof(h));
  return MRX_OK;
}

extern "C" int mrx_peer_open(const unsigned char *handle64, void **d_ptr) {
  MRX_CHECK_ARG(d_ptr != nullptr && handle64 != nullptr, "mrx_peer_open: null pointer");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, sizeof(h));
  MRX_CUDA(cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return MRX_OK;
}

extern "C" int mrx_peer_close(void *d_ptr) {
  if (d_ptr != nullptr) MRX_CUDA(cudaIpcCloseMemHandle(d_ptr));
  return MRX_OK;
}

extern "C" int mrx_peer_signal(unsigned int *d_flag, unsigned int value, void *stream) {
  MRX_CHECK_ARG(d_flag != nullptr, "mrx_peer_signal: null pointer");
  peer_signal_kernel<<<1, 1, 0, static_cast<cudaStream_t>(stream)>>>(d_flag, value);
  MRX_LAUNCH_CHECK("peer_signal_kernel");
  return MRX_OK;
}

extern "C" int mrx_peer_wait(const unsigned int *d_flags, int n_flags, unsigned int value,
                             void *stream) {
  MRX_CHECK_ARG(d_flags != nullptr && n_flags >= 0 && n_flags <= 1024, "mrx_peer_wait: bad arguments");
  if (n_flags == 0) return MRX_OK;
  peer_wait_kernel<<<1, 1024, 0, static_cast<cudaStream_t>(stream)>>>(d_flags, n_flags, value);
  MRX_LAUNCH_CHECK("peer_wait_kernel");
  return MRX_OK;
}
