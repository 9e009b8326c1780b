// H2D bandwidth probe: plain pinned vs write-combined pinned host memory, chunked like mod_pipeline_sequence_host
#include <cstdio>
#include <cstring>
#include <cuda_runtime.h>
int main() {
  const size_t bytes = 700ull << 20;
  void *h_pin, *h_wc, *d;
  cudaMallocHost(&h_pin, bytes);
  cudaHostAlloc(&h_wc, bytes, cudaHostAllocWriteCombined);
  cudaMalloc(&d, bytes);
  memset(h_pin, 1, bytes);
  memset(h_wc, 1, bytes);
  cudaStream_t s;
  cudaStreamCreate(&s);
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  for (int mode = 0; mode < 2; mode++) {
    void* h = mode ? h_wc : h_pin;
    for (size_t chunk : {bytes, bytes / 8, bytes / 64}) {
      for (int rep = 0; rep < 3; rep++) {
        cudaEventRecord(a, s);
        for (size_t off = 0; off < bytes; off += chunk)
          cudaMemcpyAsync((char*)d + off, (char*)h + off, chunk, cudaMemcpyHostToDevice, s);
        cudaEventRecord(b, s);
        cudaEventSynchronize(b);
        float ms;
        cudaEventElapsedTime(&ms, a, b);
        if (rep == 2) printf("%s chunk %zu MiB: %.1f GB/s\n", mode ? "write-combined" : "pinned", chunk >> 20, bytes / ms / 1e6);
      }
    }
  }
  return 0;
}
