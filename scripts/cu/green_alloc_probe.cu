// Does the existence of CUDA green contexts slow down fresh stream-ordered allocations?  (nvcc -o probe green_alloc_probe.cu -lcuda)
#include <cuda.h>
#include <cuda_runtime.h>
#include <chrono>
#include <cstdio>
static double now() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
static void probe(const char* tag, cudaStream_t s, size_t bytes) {
  cudaDeviceSynchronize();
  void* p = nullptr;
  double t0 = now();
  cudaMallocAsync(&p, bytes, s);
  cudaStreamSynchronize(s);
  double t1 = now();
  cudaMemsetAsync(p, 0, bytes, s);
  cudaStreamSynchronize(s);
  double t2 = now();
  cudaFreeAsync(p, s);
  cudaStreamSynchronize(s);
  cudaMemPool_t pool;
  cudaDeviceGetDefaultMemPool(&pool, 0);
  cudaMemPoolTrimTo(pool, 0);  // next probe allocates fresh memory again
  double t3 = now();
  printf("%-34s malloc %.2f ms  memset %.2f ms  free+trim %.2f ms\n", tag, t1 - t0, t2 - t1, t3 - t2);
}
int main() {
  cudaSetDevice(0);
  cudaFree(0);
  cudaStream_t s;
  cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
  const size_t GB = 1ull << 30;
  probe("fresh process, 1 GB", s, GB);
  probe("fresh process, 1 GB (again)", s, GB);
  probe("fresh process, 8 GB", s, 8 * GB);
  CUdevice dev;
  cuDeviceGet(&dev, 0);
  CUdevResource all, part, rest;
  unsigned int groups = 1;
  cuDeviceGetDevResource(dev, &all, CU_DEV_RESOURCE_TYPE_SM);
  CUresult r = cuDevSmResourceSplitByCount(&part, &groups, &all, &rest, 0, 64);
  CUdevResourceDesc d0, d1;
  cuDevResourceGenerateDesc(&d0, &part, 1);
  cuDevResourceGenerateDesc(&d1, &rest, 1);
  CUgreenCtx g0, g1;
  CUresult r0 = cuGreenCtxCreate(&g0, d0, dev, CU_GREEN_CTX_DEFAULT_STREAM);
  CUresult r1 = cuGreenCtxCreate(&g1, d1, dev, CU_GREEN_CTX_DEFAULT_STREAM);
  CUstream s0, s1;
  cuGreenCtxStreamCreate(&s0, g0, CU_STREAM_NON_BLOCKING, 0);
  cuGreenCtxStreamCreate(&s1, g1, CU_STREAM_NON_BLOCKING, 0);
  printf("split %d, create %d %d: %u + %u SMs\n", (int)r, (int)r0, (int)r1, part.sm.smCount, rest.sm.smCount);
  probe("2 green contexts alive, 1 GB", s, GB);
  probe("2 green contexts alive, 8 GB", s, 8 * GB);
  cuStreamDestroy(s0);
  cuStreamDestroy(s1);
  printf("destroy %d %d\n", (int)cuGreenCtxDestroy(g0), (int)cuGreenCtxDestroy(g1));
  probe("green contexts destroyed, 1 GB", s, GB);
  probe("green contexts destroyed, 8 GB", s, 8 * GB);
  return 0;
}
