// The same question for the TMA path: k CTAs (one per SM), one thread per CTA issues cp.async.bulk (32 KB each, 6 in
// flight = 192 KB per SM) global -> shared, nobody reads the data.  nvcc -O3 -gencode arch=compute_100a,code=sm_100a
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
constexpr int kStages = 6;
constexpr uint32_t kChunk = 32768;
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void __launch_bounds__(128, 1) tma_read_kernel(const uint8_t* __restrict__ p, size_t n_chunks, unsigned long long* sink) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar[kStages];
  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[i])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t phase[kStages] = {0};
    size_t issued = 0, done = 0;
    const size_t mine = (n_chunks - blockIdx.x + gridDim.x - 1) / gridDim.x;
    while (done < mine) {
      while (issued < mine && issued - done < (size_t)kStages) {
        const int st = (int)(issued % kStages);
        const uint8_t* src = p + (blockIdx.x + issued * gridDim.x) * (size_t)kChunk;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar[st])), "r"(kChunk) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         smem_u32(smem + (size_t)st * kChunk)),
                     "l"(src), "r"(kChunk), "r"(smem_u32(&bar[st]))
                     : "memory");
        ++issued;
      }
      const int st = (int)(done % kStages);
      uint32_t ok = 0;
      while (!ok) {
        asm volatile(
            "{\n\t.reg .pred q;\n\tmbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2;\n\tselp.u32 %0, 1, 0, q;\n\t}"
            : "=r"(ok)
            : "r"(smem_u32(&bar[st])), "r"(phase[st])
            : "memory");
      }
      phase[st] ^= 1u;
      ++done;
    }
    if (smem[5] == 77) atomicAdd(sink, 1ull);
  }
}
int main() {
  const size_t bytes = 16ull << 30;
  uint8_t* buf;
  unsigned long long* sink;
  cudaMalloc(&buf, bytes);
  cudaMalloc(&sink, 8);
  cudaMemset(buf, 1, bytes);
  const size_t smem = (size_t)kStages * kChunk;
  cudaFuncSetAttribute(tma_read_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  const int ks[] = {148, 111, 84, 60, 37};
  for (int k : ks) {
    tma_read_kernel<<<k, 128, smem>>>(buf, bytes / kChunk, sink);
    cudaEventRecord(a);
    tma_read_kernel<<<k, 128, smem>>>(buf, bytes / kChunk, sink);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    printf("%3d SMs (cp.async.bulk, 6 x 32 KB in flight per SM): %7.1f GB/s  = %5.1f GB/s per SM\n", k, bytes / ms * 1e-6, bytes / ms * 1e-6 / k);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
