// common.cuh -- shared device helpers for libvsm (sm_100a only).
//
// Arithmetic that decides voxel keys is written with explicit round-to-nearest
// intrinsics so that nvcc can neither contract nor reorder it: the reference
// computes the 4x4 transform in float64 (numpy matmul), divides by w in float64,
// rounds to float32, divides by the voxel size in float32 and floors
// (vggt_slam/submap.py:277-282, vggt_slam/map.py:232-234, 351).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/vsm.h"

namespace vsm {

// ---------------------------------------------------------------------------
// error plumbing
// ---------------------------------------------------------------------------
void set_error(const char* fmt, ...);
extern int64_t g_launches;  // kernels launched by this library (vsm_launch_count)

#define VSM_CUDA(expr)                                                                              \
  do {                                                                                              \
    cudaError_t _e = (expr);                                                                        \
    if (_e != cudaSuccess) {                                                                        \
      ::vsm::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e));       \
      return (_e == cudaErrorMemoryAllocation) ? VSM_E_NOMEM : VSM_E_CUDA;                          \
    }                                                                                               \
  } while (0)

#define VSM_TRY(expr)            \
  do {                           \
    int _s = (expr);             \
    if (_s != VSM_OK) return _s; \
  } while (0)

#define VSM_LAUNCHED()                                                                     \
  do {                                                                                     \
    ++::vsm::g_launches;                                                                   \
    cudaError_t _e = cudaGetLastError();                                                   \
    if (_e != cudaSuccess) {                                                               \
      ::vsm::set_error("%s:%d: kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(_e)); \
      return VSM_E_CUDA;                                                                   \
    }                                                                                      \
  } while (0)

static inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }
int sm_count();  // multiprocessors of the current device (cached per device; select.cu)
static inline int grid_for(int64_t n, int block, int max_blocks = 0) {
  if (max_blocks <= 0) max_blocks = sm_count() * 16;
  int64_t g = cdiv(n, block);
  if (g < 1) g = 1;
  if (g > max_blocks) g = max_blocks;
  return (int)g;
}
static inline uint64_t next_pow2(uint64_t x) {
  uint64_t p = 1;
  while (p < x) p <<= 1;
  return p;
}

// ---------------------------------------------------------------------------
// voxel keys
// ---------------------------------------------------------------------------
// A key packs three axis codes of 21 bits: code 0 stands for INT64_MIN (what the
// reference's float->int64 cast yields for NaN / Inf / out-of-range values),
// code c in [1, 2^21-1] stands for the coordinate c - 2^20.  Unsigned order of
// the packed key == lexicographic signed order of (ix,iy,iz) == np.unique(axis=0).
constexpr uint64_t kEmptyKey = 0xFFFFFFFFFFFFFFFFull;
constexpr int kAxisBias = 1 << 20;

struct HMat {
  double m[16];
};

__host__ __device__ __forceinline__ uint64_t mix64(uint64_t x) {
  x ^= x >> 33;
  x *= 0xff51afd7ed558ccdull;
  x ^= x >> 33;
  x *= 0xc4ceb9fe1a85ec53ull;
  x ^= x >> 33;
  return x;
}

#ifdef __CUDACC__
// (H @ [p;1]) / w in float64, as the FMA chain  fma(h3,1,fma(h2,z,fma(h1,y,h0*x))).
__device__ __forceinline__ void transform_f64(const HMat& Hm, float px, float py, float pz, double& ox, double& oy,
                                              double& oz) {
  const double x = (double)px, y = (double)py, z = (double)pz;
  double r[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    double a = __dmul_rn(Hm.m[4 * i + 0], x);
    a = __fma_rn(Hm.m[4 * i + 1], y, a);
    a = __fma_rn(Hm.m[4 * i + 2], z, a);
    a = __fma_rn(Hm.m[4 * i + 3], 1.0, a);
    r[i] = a;
  }
  ox = __ddiv_rn(r[0], r[3]);
  oy = __ddiv_rn(r[1], r[3]);
  oz = __ddiv_rn(r[2], r[3]);
}

__device__ __forceinline__ void transform_f32(const HMat& Hm, float px, float py, float pz, float& ox, float& oy,
                                              float& oz) {
  double x, y, z;
  transform_f64(Hm, px, py, pz, x, y, z);
  ox = __double2float_rn(x);
  oy = __double2float_rn(y);
  oz = __double2float_rn(z);
}

__device__ __forceinline__ bool finite3(float x, float y, float z) {
  return (fabsf(x) <= 3.402823466e38f) && (fabsf(y) <= 3.402823466e38f) && (fabsf(z) <= 3.402823466e38f);
}

// floor(p / cell) -> axis code.  range_err is set for finite coordinates we cannot pack.
__device__ __forceinline__ uint32_t axis_code(float p, float cell, bool& range_err) {
  const float q = floorf(__fdiv_rn(p, cell));
  if (fabsf(q) < 1048576.0f) return (uint32_t)((int)q + kAxisBias);
  // NaN, +-Inf and |q| >= 2^63 become INT64_MIN in the reference's cast
  if (!(fabsf(q) < 9223372036854775808.0f)) return 0u;
  range_err = true;
  return 0u;
}

__device__ __forceinline__ uint64_t pack_key(float x, float y, float z, float cell, bool& range_err) {
  const uint64_t cx = axis_code(x, cell, range_err);
  const uint64_t cy = axis_code(y, cell, range_err);
  const uint64_t cz = axis_code(z, cell, range_err);
  return (cx << 42) | (cy << 21) | cz;
}

__host__ __device__ __forceinline__ int64_t axis_from_code(uint32_t c) {
  return c == 0u ? (int64_t)0x8000000000000000ull : (int64_t)c - (int64_t)kAxisBias;
}
__host__ __device__ __forceinline__ void unpack_key(uint64_t k, int64_t& x, int64_t& y, int64_t& z) {
  x = axis_from_code((uint32_t)((k >> 42) & 0x1FFFFFu));
  y = axis_from_code((uint32_t)((k >> 21) & 0x1FFFFFu));
  z = axis_from_code((uint32_t)(k & 0x1FFFFFu));
}

// float32 <-> order-preserving uint32 (radix select / top-k keys); NaN sorts above +Inf
__device__ __forceinline__ uint32_t float_to_ordered(float f) {
  uint32_t u = __float_as_uint(f);
  if ((u & 0x7FFFFFFFu) > 0x7F800000u) return 0xFFFFFFFFu;  // NaN (either sign)
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ordered_to_float(uint32_t u) {
  if (u == 0xFFFFFFFFu) return __uint_as_float(0x7FC00000u);
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u);
}

// 128-bit streaming load (read once, keep out of L1)
__device__ __forceinline__ uint4 ld_stream_v4(const void* p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(p));
  return v;
}
// vector reduction into global memory, no return value (RED.E.ADD.F32x4 on sm_100a)
__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__device__ __forceinline__ int lane_id() { return (int)(threadIdx.x & 31u); }
#endif  // __CUDACC__

}  // namespace vsm
