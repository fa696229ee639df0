// How fast can kernels store small granules into page-locked host memory (posted PCIe writes)? For each granule size and
// spacing: microseconds for `count` granules, granules per microsecond, payload GB/s. Build: nvcc -O2 -arch=sm_100a -o pcie_write_probe pcie_write_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

// granule g (lanes16 uint4 lanes wide) goes to byte offset g * pitch; thread = one 16-byte lane
__global__ void store_granules(uint4* host, long long count, int lanes16, long long pitch16, unsigned tag) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long g = i / lanes16; const int l = (int)(i % lanes16);
  if (g < count) host[g * pitch16 + l] = make_uint4(tag, (unsigned)g, l, 7u);
}

int main() {
  const size_t bytes = 512ull << 20;
  uint4* h; CK(cudaHostAlloc(&h, bytes, cudaHostAllocMapped | cudaHostAllocPortable));
  uint4* d; CK(cudaHostGetDevicePointer(&d, h, 0));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const long long counts[] = {10000, 40000, 160000};
  const int sizes[] = {16, 32, 64, 128, 256, 512};
  const int pitches[] = {0 /* adjacent */, 84 * 8 /* like 8 rows apart, unaligned multiple of 16 */, 4096};
  for (long long count : counts) for (int sz : sizes) for (int pitch : pitches) {
    const int lanes16 = sz / 16;
    long long pitch16 = pitch ? ((pitch + sz + 15) / 16 + 3) / 4 * 4 : lanes16;   // in uint4 units; spaced variants keep 64-byte alignment
    if ((size_t)(count * pitch16 * 16 + sz) > bytes) continue;
    const long long threads = count * lanes16;
    const int blocks = (int)((threads + 255) / 256);
    float best = 1e9f;
    for (int rep = 0; rep < 7; rep++) {
      CK(cudaEventRecord(e0));
      store_granules<<<blocks, 256>>>(d, count, lanes16, pitch16, rep);
      CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
      float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
      if (rep >= 2 && ms < best) best = ms;
    }
    printf("count %7lld size %4d B spacing %5lld B: %8.1f us  %7.1f granules/us  %6.2f GB/s payload\n", count, sz, pitch16 * 16, best * 1e3, count / (best * 1e3), count * (double)sz / (best * 1e-3) / 1e9);
  }
  return 0;
}
