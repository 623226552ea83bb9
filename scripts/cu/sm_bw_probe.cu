// How much HBM bandwidth can k SMs pull?  One CTA per SM (200 KB of dynamic shared memory forces that), 1024 threads,
// 16-byte streaming loads, 8 in flight per thread, over a 16 GB buffer.  nvcc -O3 -arch=sm_100a -o sm_bw_probe sm_bw_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
__global__ void __launch_bounds__(1024, 1) read_kernel(const uint4* __restrict__ p, size_t n16, unsigned long long* sink) {
  extern __shared__ uint8_t pad[];
  uint32_t acc = 0;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + 7 * stride < n16; i += 8 * stride) {
    uint4 v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u)
      asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v[u].x), "=r"(v[u].y), "=r"(v[u].z), "=r"(v[u].w) : "l"(p + i + u * stride));
#pragma unroll
    for (int u = 0; u < 8; ++u) acc ^= v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
  }
  if (acc == 0x12345678u) atomicAdd(sink, 1ull);
  if (threadIdx.x == 0 && pad[0] == 77) atomicAdd(sink, 1ull);
}
int main() {
  const size_t bytes = 16ull << 30;
  uint4* buf;
  unsigned long long* sink;
  cudaMalloc(&buf, bytes);
  cudaMalloc(&sink, 8);
  cudaMemset(buf, 1, bytes);
  cudaFuncSetAttribute(read_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  const int ks[] = {148, 120, 111, 100, 92, 84, 76, 68, 60, 48, 37};
  for (int k : ks) {
    read_kernel<<<k, 1024, 200 * 1024>>>(buf, bytes / 16, sink);  // warm-up
    cudaEventRecord(a);
    read_kernel<<<k, 1024, 200 * 1024>>>(buf, bytes / 16, sink);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    printf("%3d SMs (1 CTA of 1024 threads each, 8 x 16 B in flight per thread): %7.1f GB/s  = %5.1f GB/s per SM\n", k, bytes / ms * 1e-6,
           bytes / ms * 1e-6 / k);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
