// bracket.cuh -- the bbox filter's percentiles without a pass of their own (included by fuse.cu only).
//
// map.py:257-258 needs the 0.5 % / 99.5 % order statistics of ~5 M world points per axis.  The three-pass radix select
// (select.cu) reads all points three times.  Here:
//   (1) bracket_sample_kernel: one CTA per axis transforms a hashed sample of 2048 RAW points itself, sorts the
//       valid ones in shared memory and reads off, per percentile, a bracket [lo, hi] of sample order statistics
//       6 sigma either side of the wanted sample rank -- before the world-point kernel runs;
//   (2) world_points_kernel, which visits every point anyway, counts per bracket the valid points below it and
//       appends those inside it (about 2 % of the points) to a list (bracket_collect);
//   (3) bracket_resolve_kernel: one CTA per bracket radix-selects the wanted ranks among the collected values and
//       applies numpy's lerp.
// The answer is the exact order statistic whenever the wanted ranks fall inside the bracket.  A miss (probability
// ~1e-9 per call for a random sample), a list that overflows or a sample with too few valid points sets *miss: the
// fuse call is aborted before it touches the map and repeated with the radix select -- never an approximation.
#pragma once
#include "state.cuh"

namespace vsm {

constexpr int kBrSample = 2048;
constexpr int kBrLists = 6;  // [axis][lo pct, hi pct]

struct BracketState {
  float lo[kBrLists], hi[kBrLists];  // bracket bounds (inclusive); -inf / +inf when the sample rank was clipped
  uint32_t below[kBrLists];          // valid points < lo
  uint32_t cursor[kBrLists];         // points collected (can exceed the list capacity: overflow)
  uint32_t miss;
  uint32_t pad[3];
};
static_assert(sizeof(BracketState) <= 256, "bracket state fits its 256-byte slot");

struct BracketArgs {
  BracketState* bs;
  float* lists;  // [kBrLists][cap]
  uint32_t cap;
};

static inline uint32_t bracket_list_cap(int64_t n_items) { return (uint32_t)std::max<int64_t>(n_items / 8, 4096); }
static inline size_t bracket_scratch_bytes(int64_t n_items) {
  return 256 + (size_t)kBrLists * bracket_list_cap(n_items) * sizeof(float);
}

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
  x ^= x >> 16;
  x *= 0x7feb352du;
  x ^= x >> 15;
  x *= 0x846ca68bu;
  x ^= x >> 16;
  return x;
}

// Per-thread state of the collect step inside the world-point kernel.
struct BracketLocal {
  float lo[kBrLists], hi[kBrLists];
  uint32_t below[kBrLists];
};

__device__ __forceinline__ void bracket_load(BracketLocal& L, const BracketState* bs) {
#pragma unroll
  for (int t = 0; t < kBrLists; ++t) {
    L.lo[t] = bs->lo[t];
    L.hi[t] = bs->hi[t];
    L.below[t] = 0u;
  }
}

// all 32 lanes must call (ballots); `valid`: the point takes part in the percentiles
__device__ __forceinline__ void bracket_collect(BracketLocal& L, const BracketArgs& a, float x, float y, float z, bool valid) {
  const float v[3] = {x, y, z};
  const int lane = lane_id();
#pragma unroll
  for (int t = 0; t < kBrLists; ++t) {
    const float c = v[t >> 1];
    L.below[t] += (valid && c < L.lo[t]) ? 1u : 0u;
    const bool in = valid && c >= L.lo[t] && c <= L.hi[t];
    const unsigned mk = __ballot_sync(0xffffffffu, in);
    if (mk) {
      uint32_t base = 0;
      const int leader = __ffs(mk) - 1;
      if (lane == leader) base = atomicAdd(&a.bs->cursor[t], (uint32_t)__popc(mk));
      base = __shfl_sync(0xffffffffu, base, leader);
      if (in) {
        const uint32_t pos = base + (uint32_t)__popc(mk & ((1u << lane) - 1u));
        if (pos < a.cap) a.lists[(size_t)t * a.cap + pos] = c;
      }
    }
  }
}

// end of the kernel: one atomic per warp and bracket
__device__ __forceinline__ void bracket_flush(BracketLocal& L, const BracketArgs& a) {
#pragma unroll
  for (int t = 0; t < kBrLists; ++t) {
    uint32_t b = L.below[t];
    for (int o = 16; o > 0; o >>= 1) b += __shfl_xor_sync(0xffffffffu, b, o);
    if (lane_id() == 0 && b) atomicAdd(&a.bs->below[t], b);
  }
}

// grid = 6 CTAs of 1024 threads: exact ranks inside the collected list, then the numpy lerp
__global__ void __launch_bounds__(1024) bracket_resolve_kernel(BracketState* bs, const float* __restrict__ lists,
                                                                uint32_t list_cap, const unsigned long long* n_dev,
                                                                float q0, float q1, float* __restrict__ out,
                                                                uint32_t* miss_out) {
  __shared__ uint32_t hist[2][2048];
  __shared__ uint32_t s_prefix[2], s_rem[2];
  const int t = blockIdx.x;
  const float q = (t & 1) ? q1 : q0;
  const unsigned long long n = *n_dev;
  if (n == 0) {
    if (threadIdx.x == 0) out[t] = __uint_as_float(0x7FC00000u);
    return;
  }
  // ranks and lerp weight exactly as sel_plan_kernel
  const float nm1 = (float)(n - 1);
  const float vidx = __fmul_rn(nm1, q);
  unsigned long long rlo, rhi;
  if (vidx >= nm1) {
    rlo = rhi = n - 1;
  } else if (vidx < 0.f) {
    rlo = rhi = 0;
  } else {
    rlo = (unsigned long long)floorf(vidx);
    rhi = rlo + 1;
    if (rhi > n - 1) rhi = n - 1;
  }
  const float g = __fsub_rn(vidx, floorf(vidx));
  const unsigned long long below = bs->below[t];
  const uint32_t len = bs->cursor[t];
  if (bs->miss || len > list_cap || rlo < below || rhi - below >= (unsigned long long)len) {
    if (threadIdx.x == 0) {
      atomicOr(miss_out, 1u);
      out[t] = __uint_as_float(0x7FC00000u);
    }
    return;
  }
  const float* L = lists + (size_t)t * list_cap;
  // The collected values span a narrow range, so their high bits are all alike: select on (key - smallest key),
  // starting at the highest bit that differs -- two well-spread 11-bit passes instead of three crowded ones.
  __shared__ uint32_t s_min, s_max;
  if (threadIdx.x == 0) {
    s_min = 0xFFFFFFFFu;
    s_max = 0u;
  }
  if (threadIdx.x < 2) {
    s_prefix[threadIdx.x] = 0u;
    s_rem[threadIdx.x] = (uint32_t)((threadIdx.x == 0 ? rlo : rhi) - below);
  }
  __syncthreads();
  uint32_t kmin = 0xFFFFFFFFu, kmax = 0u;
  for (uint32_t i = threadIdx.x; i < len; i += blockDim.x) {
    const uint32_t o = float_to_ordered(L[i]);
    kmin = min(kmin, o);
    kmax = max(kmax, o);
  }
  for (int o = 16; o > 0; o >>= 1) {
    kmin = min(kmin, __shfl_xor_sync(0xffffffffu, kmin, o));
    kmax = max(kmax, __shfl_xor_sync(0xffffffffu, kmax, o));
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMin(&s_min, kmin);
    atomicMax(&s_max, kmax);
  }
  __syncthreads();
  const uint32_t base_key = s_min;
  int top = 32 - __clz(s_max - base_key);  // significant bits of (key - base_key); 0: all values equal
  while (top > 0) {
    const int nbit = top < 11 ? top : 11;
    const int shift = top - nbit;
    const int nb = 1 << nbit;
    for (int i = threadIdx.x; i < 2 * 2048; i += blockDim.x) (&hist[0][0])[i] = 0u;
    __syncthreads();
    const uint32_t p0 = s_prefix[0], p1 = s_prefix[1];
    for (uint32_t i = threadIdx.x; i < len; i += blockDim.x) {
      const uint32_t o = float_to_ordered(L[i]) - base_key;
      const uint32_t hi_bits = top >= 32 ? 0u : (o >> top);
      const uint32_t dig = (o >> shift) & (uint32_t)(nb - 1);
      if (hi_bits == p0) atomicAdd(&hist[0][dig], 1u);
      if (hi_bits == p1) atomicAdd(&hist[1][dig], 1u);
    }
    __syncthreads();
    if (threadIdx.x < 64) {
      // one warp per target scans its histogram (as sel_pick_kernel)
      const int k = threadIdx.x >> 5, lane = threadIdx.x & 31;
      const int per = (nb + 31) / 32;
      uint32_t mine = 0;
      for (int i = 0; i < per; ++i)
        if (lane * per + i < nb) mine += hist[k][lane * per + i];
      uint32_t incl = mine;
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t up = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += up;
      }
      const uint32_t excl = incl - mine, want = s_rem[k];
      if (want >= excl && want < incl) {
        uint32_t run = excl;
        int bin = lane * per;
        for (int i = 0; i < per; ++i) {
          const uint32_t cc = hist[k][lane * per + i];
          if (want < run + cc) {
            bin = lane * per + i;
            break;
          }
          run += cc;
        }
        s_prefix[k] = (s_prefix[k] << nbit) | (uint32_t)bin;
        s_rem[k] = want - run;
      }
    }
    __syncthreads();
    top = shift;
  }
  if (threadIdx.x == 0) {
    const float a = ordered_to_float(base_key + s_prefix[0]);
    const float b = ordered_to_float(base_key + s_prefix[1]);
    const float diff = __fsub_rn(b, a);
    float r = __fadd_rn(a, __fmul_rn(diff, g));
    if (g >= 0.5f) r = __fsub_rn(b, __fmul_rn(diff, __fsub_rn(1.0f, g)));
    out[t] = r;
  }
}


}  // namespace vsm
