// hostcopy.cuh -- host <-> device copies of PAGEABLE message buffers for the single-frame host seams.
//
// The two ROS nodes hand the C ABI ordinary std::vector storage (sensor_msgs/Image.data, PointCloud2.data).  A plain
// cudaMemcpyAsync from / to pageable memory makes the driver stage the bytes through its own bounce buffers on ONE
// host thread: 6-19 GB/s measured on the B200 boxes, against 55 GB/s on the PCIe Gen5 link - and at 32 B per pixel the
// cloud copy, not the kernels, is what a frame costs (profiles/r01_single_frame.jsonl).
// HostStager does the staging itself: a few host threads copy the buffer chunk-wise into (out of) a pinned
// double-buffered window while the copy engine moves the previous window, so the two overlap and the host side scales
// with the thread count.  Buffers that already are pinned (mod_host_alloc, mod_host_register) skip the stager: one DMA.
#pragma once

#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace mod {

// Process-wide worker threads (started on first use).  n <= 0: the stager is off and the seams fall back to the
// driver's pageable path.  Default: MOD_B200_COPY_THREADS if set, else min(4, hardware threads / 2), at least 1.
int host_copy_threads();
void set_host_copy_threads(int n);

class HostStager {
 public:
  HostStager() = default;
  ~HostStager();
  HostStager(const HostStager&) = delete;
  HostStager& operator=(const HostStager&) = delete;

  // src (host, any memory) -> dst (device), ordered on `stream`.  Returns once every byte of src has been read (the
  // caller may release src); the last DMAs may still be in flight on the stream.
  cudaError_t h2d(void* d_dst, const void* h_src, size_t bytes, cudaStream_t stream);
  // src (device) -> dst (host, any memory), ordered after the work already on `stream`.  Returns once dst is filled.
  cudaError_t d2h(void* h_dst, const void* d_src, size_t bytes, cudaStream_t stream);

  // copies smaller than this go straight to cudaMemcpyAsync
  static constexpr size_t kMinBytes = 512u << 10;

 private:
  cudaError_t ensure(int threads);
  static bool is_pinned(const void* p);

  uint8_t* window_ = nullptr;     // 2 halves of threads_ * kChunk bytes, pinned
  cudaEvent_t done_[2] = {nullptr, nullptr};  // last DMA that used each half
  int threads_ = 0;
#ifndef MOD_COPY_CHUNK_KB
#define MOD_COPY_CHUNK_KB 512
#endif
  static constexpr size_t kChunk = static_cast<size_t>(MOD_COPY_CHUNK_KB) << 10;  // bytes one thread copies per wave
};

}  // namespace mod
