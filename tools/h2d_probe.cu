// H2D bandwidth from pinned memory as a function of copy size and number of streams.
#include <cuda_runtime.h>
#include <chrono>
#include <cstdio>
#include <vector>
int main() {
  const size_t total = 1228800000;
  char* h; char* d;
  cudaMallocHost(&h, total); cudaMalloc(&d, total);
  memset(h, 1, total);
  size_t sizes[] = {2400000, 4800000, 19200000, 307200000};
  int nstreams[] = {1, 2, 4};
  std::vector<cudaStream_t> st(4);
  for (auto& s : st) cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
  for (size_t sz : sizes) for (int ns : nstreams) {
    for (int rep = 0; rep < 2; rep++) {
      cudaDeviceSynchronize();
      auto t0 = std::chrono::steady_clock::now();
      size_t off = 0; int k = 0;
      while (off + sz <= total) { cudaMemcpyAsync(d + off, h + off, sz, cudaMemcpyHostToDevice, st[k % ns]); off += sz; k++; }
      cudaDeviceSynchronize();
      double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
      if (rep) printf("copy %9zu B x %4d on %d stream(s): %.2f ms, %.1f GB/s\n", sz, k, ns, ms, off / ms / 1e6);
    }
  }
  return 0;
}
