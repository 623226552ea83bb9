// bracket.cuh -- the bbox filter's percentiles without a pass of their own (included by fuse.cu only).
//
// map.py:257-258 needs the 0.5 % / 99.5 % order statistics of ~5 M world points per axis.  The three-pass radix select
// (select.cu) reads all points three times.  Here:
//   (1) bracket_sample_kernel: one CTA per axis transforms a hashed sample of 8000 RAW points itself, sorts the
//       valid ones in shared memory and reads off, per percentile, a bracket [lo, hi] of sample order statistics
//       5 sigma either side of the wanted sample rank -- before the world-point kernel runs;
//   (2) world_points_kernel, which visits every point anyway, counts per bracket the valid points below it and
//       appends those inside it (under 1 % of the points per bracket) to a list (bracket_collect: staged per warp in
//       shared memory, one global atomic per ~48 values);
//   (3) bracket_resolve_kernel: one CTA per bracket radix-selects the wanted ranks among the collected values and
//       applies numpy's lerp.
// The answer is the exact order statistic whenever the wanted ranks fall inside the bracket.  A miss (probability
// ~1e-6 per call for a random sample), a list that overflows or a sample with too few valid points sets *miss: the
// fuse call is aborted before it touches the map and repeated with the radix select -- never an approximation.
#pragma once
#include "state.cuh"

namespace vsm {

constexpr int kBrSample = 8000;  // (keys + the select's histograms stay within 48 KB of static shared memory)
constexpr int kBrLists = 6;  // [axis][lo pct, hi pct]

struct BracketState {
  float lo[kBrLists], hi[kBrLists];  // bracket bounds (inclusive); -inf / +inf when the sample rank was clipped
  uint32_t below[kBrLists];          // valid points < lo
  uint32_t cursor[kBrLists];         // points collected (can exceed the list capacity: overflow)
  uint32_t miss;
  // deferred mode (the brackets classify the points inside the insert kernel, fuse.cu insert7d_kernel): valid points
  // ABOVE the high brackets are counted instead of those below them (which would be nearly all points), and the
  // pixels whose box test has to wait for the exact percentiles are listed
  uint32_t deferred;
  uint32_t above[kBrLists];
  uint32_t und_cursor;
};
static_assert(sizeof(BracketState) <= 256, "bracket state fits its 256-byte slot");

struct BracketArgs {
  BracketState* bs;
  float* lists;  // [kBrLists][cap]
  uint32_t cap;
  uint32_t* und;  // undecided pixels (deferred mode)
  uint32_t und_cap;
};

static inline uint32_t bracket_list_cap(int64_t n_items) { return (uint32_t)std::max<int64_t>(n_items / 8, 4096); }
static inline uint32_t bracket_und_cap(int64_t n_items) { return (uint32_t)std::max<int64_t>(n_items / 2, 4096); }
static inline size_t bracket_scratch_bytes(int64_t n_items) {
  return 256 + (size_t)kBrLists * bracket_list_cap(n_items) * sizeof(float) + (size_t)bracket_und_cap(n_items) * sizeof(uint32_t);
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

// Collect step inside the world-point kernel.  Every warp stages its hits per bracket in shared memory and appends
// them to the bracket's list ~48 at a time: one global atomic per flush.  (The first version reserved a list slot
// per warp-iteration and bracket with a global atomic on six hot counters: it more than doubled the kernel.)
constexpr int kBrStage = 64;   // staged values per (warp, bracket)
constexpr int kBrWarps = 8;    // warps per CTA of the world-point kernel

__device__ __forceinline__ void bracket_stage_flush(const BracketArgs& a, float* buf, uint32_t* cnt, int t) {
  const int lane = lane_id();
  __syncwarp();
  const uint32_t n = *cnt;
  if (n == 0u) return;
  uint32_t base = 0;
  if (lane == 0) base = atomicAdd(&a.bs->cursor[t], n);
  base = __shfl_sync(0xffffffffu, base, 0);
  for (uint32_t i = lane; i < n; i += 32u)
    if (base + i < a.cap) a.lists[(size_t)t * a.cap + base + i] = buf[i];
  __syncwarp();
  if (lane == 0) *cnt = 0u;
  __syncwarp();
}

// all 32 lanes must call.  P points per thread: v[j] = (x,y,z) of point j, valid[j]: it takes part in the percentiles.
// stage / stage_n: this warp's shared-memory staging area, [kBrLists][kBrStage] floats and [kBrLists] counters.
template <int P>
__device__ __forceinline__ void bracket_collect(BracketLocal& L, const BracketArgs& a, const float4* v, const bool* valid,
                                                float* stage, uint32_t* stage_n) {
  const int lane = lane_id();
  uint32_t hits = 0u;  // bit (t * P + j): point j lies inside bracket t
#pragma unroll
  for (int j = 0; j < P; ++j) {
    const float c3[3] = {v[j].x, v[j].y, v[j].z};
#pragma unroll
    for (int t = 0; t < kBrLists; ++t) {
      const float c = c3[t >> 1];
      L.below[t] += (valid[j] && c < L.lo[t]) ? 1u : 0u;
      hits |= (valid[j] && c >= L.lo[t] && c <= L.hi[t]) ? (1u << (t * P + j)) : 0u;
    }
  }
  if (!__any_sync(0xffffffffu, hits != 0u)) return;
#pragma unroll
  for (int t = 0; t < kBrLists; ++t) {
    const uint32_t mine = (hits >> (t * P)) & ((1u << P) - 1u);
    const uint32_t cnt = (uint32_t)__popc(mine);
    if (__ballot_sync(0xffffffffu, cnt != 0u) == 0u) continue;
    uint32_t incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t up = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += up;
    }
    const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
    float* buf = stage + t * kBrStage;
    if (total > (uint32_t)kBrStage / 2) {
      // many hits at once (a clipped bracket, a tiny submap): straight to the list
      uint32_t base = 0;
      if (lane == 0) base = atomicAdd(&a.bs->cursor[t], total);
      uint32_t pos = __shfl_sync(0xffffffffu, base, 0) + incl - cnt;
#pragma unroll
      for (int j = 0; j < P; ++j)
        if (mine & (1u << j)) {
          if (pos < a.cap) a.lists[(size_t)t * a.cap + pos] = (t >> 1) == 0 ? v[j].x : ((t >> 1) == 1 ? v[j].y : v[j].z);
          ++pos;
        }
      continue;
    }
    if (stage_n[t] + total > (uint32_t)kBrStage) bracket_stage_flush(a, buf, stage_n + t, t);
    uint32_t pos = stage_n[t] + incl - cnt;
#pragma unroll
    for (int j = 0; j < P; ++j)
      if (mine & (1u << j)) buf[pos++] = (t >> 1) == 0 ? v[j].x : ((t >> 1) == 1 ? v[j].y : v[j].z);
    __syncwarp();
    if (lane == 0) stage_n[t] += total;
    __syncwarp();
  }
}

// end of the kernel: staged values, then one atomic per warp and bracket for the counts below the brackets
__device__ __forceinline__ void bracket_flush(BracketLocal& L, const BracketArgs& a, float* stage, uint32_t* stage_n) {
#pragma unroll
  for (int t = 0; t < kBrLists; ++t) {
    bracket_stage_flush(a, stage + t * kBrStage, stage_n + t, t);
    uint32_t b = L.below[t];
    for (int o = 16; o > 0; o >>= 1) b += __shfl_xor_sync(0xffffffffu, b, o);
    if (lane_id() == 0 && b) atomicAdd(&a.bs->below[t], b);
  }
}

// Two order statistics (0-based ranks r0 <= r1 < n) of n float-ordered keys, by one CTA.  keys: shared or global
// memory; hist: 2 x 2048 words of shared memory; sm: 8 words of shared memory.  The keys of a bracket / a sample span a
// narrow range, so the select runs on (key - smallest key) from the highest bit that differs: usually two 11-bit
// passes.  All threads of the CTA must call; the results are valid in every thread.
__device__ __forceinline__ void cta_select2(const uint32_t* keys, uint32_t n, uint32_t r0, uint32_t r1, uint32_t* hist,
                                            uint32_t* sm, uint32_t& k0, uint32_t& k1) {
  uint32_t* s_prefix = sm;      // [2]
  uint32_t* s_rem = sm + 2;     // [2]
  uint32_t* s_minmax = sm + 4;  // [2]
  if (threadIdx.x == 0) {
    s_minmax[0] = 0xFFFFFFFFu;
    s_minmax[1] = 0u;
    s_prefix[0] = s_prefix[1] = 0u;
    s_rem[0] = r0;
    s_rem[1] = r1;
  }
  __syncthreads();
  uint32_t kmin = 0xFFFFFFFFu, kmax = 0u;
#pragma unroll 8
  for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
    const uint32_t o = keys[i];
    kmin = min(kmin, o);
    kmax = max(kmax, o);
  }
  for (int o = 16; o > 0; o >>= 1) {
    kmin = min(kmin, __shfl_xor_sync(0xffffffffu, kmin, o));
    kmax = max(kmax, __shfl_xor_sync(0xffffffffu, kmax, o));
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMin(&s_minmax[0], kmin);
    atomicMax(&s_minmax[1], kmax);
  }
  __syncthreads();
  const uint32_t base_key = s_minmax[0];
  int top = 32 - __clz(s_minmax[1] - base_key);  // significant bits of (key - base_key); 0: all keys equal
  while (top > 0) {
    const int nbit = top < 11 ? top : 11;
    const int shift = top - nbit;
    const int nb = 1 << nbit;
    for (int i = threadIdx.x; i < 2 * 2048; i += blockDim.x) hist[i] = 0u;
    __syncthreads();
    const uint32_t p0 = s_prefix[0], p1 = s_prefix[1];
#pragma unroll 8
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
      const uint32_t o = keys[i] - base_key;
      const uint32_t hi_bits = top >= 32 ? 0u : (o >> top);
      const uint32_t dig = (o >> shift) & (uint32_t)(nb - 1);
      if (hi_bits == p0) atomicAdd(&hist[dig], 1u);
      if (hi_bits == p1) atomicAdd(&hist[2048 + dig], 1u);
    }
    __syncthreads();
    if (threadIdx.x < 64) {
      // one warp per target scans its histogram
      const int k = threadIdx.x >> 5, lane = threadIdx.x & 31;
      const uint32_t* h = hist + 2048 * k;
      const int per = (nb + 31) / 32;
      uint32_t mine = 0;
      for (int i = 0; i < per; ++i)
        if (lane * per + i < nb) mine += h[lane * per + i];
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
          const uint32_t cc = h[lane * per + i];
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
  k0 = base_key + s_prefix[0];
  k1 = base_key + s_prefix[1];
  __syncthreads();
}

constexpr uint32_t kResolveStage = 40960;  // keys of a bracket staged in shared memory (160 KB); longer lists are read in place

// grid = 6 CTAs of 1024 threads (dynamic shared memory: kResolveStage words): exact ranks inside the collected list,
// then the numpy lerp
__global__ void __launch_bounds__(1024) bracket_resolve_kernel(BracketState* bs, const float* __restrict__ lists,
                                                                uint32_t list_cap, const unsigned long long* n_dev,
                                                                float q0, float q1, float* __restrict__ out,
                                                                uint32_t* miss_out, uint32_t und_cap) {
  extern __shared__ uint32_t s_keys[];
  __shared__ uint32_t hist[2 * 2048];
  __shared__ uint32_t sm[8];
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
  const uint32_t len = bs->cursor[t];
  // deferred mode counts the valid points above a high bracket: those below it are the rest
  const unsigned long long below =
      (bs->deferred && (t & 1)) ? n - (unsigned long long)bs->above[t] - (unsigned long long)len : bs->below[t];
  if (bs->miss || bs->und_cursor > und_cap || len > list_cap || (unsigned long long)bs->above[t] + len > n || rlo < below ||
      rhi - below >= (unsigned long long)len) {
    if (threadIdx.x == 0) {
      atomicOr(miss_out, 1u);
      out[t] = __uint_as_float(0x7FC00000u);
    }
    return;
  }
  const float* L = lists + (size_t)t * list_cap;
  const uint32_t* keys;
  if (len <= kResolveStage) {
#pragma unroll 8
    for (uint32_t i = threadIdx.x; i < len; i += blockDim.x) s_keys[i] = float_to_ordered(L[i]);
    keys = s_keys;
  } else {
    // a long list (a clipped bracket): ordered keys written back in place, read from L2
    uint32_t* Lk = reinterpret_cast<uint32_t*>(const_cast<float*>(L));
    for (uint32_t i = threadIdx.x; i < len; i += blockDim.x) Lk[i] = float_to_ordered(L[i]);
    keys = Lk;
  }
  __syncthreads();
  uint32_t k0, k1;
  cta_select2(keys, len, (uint32_t)(rlo - below), (uint32_t)(rhi - below), hist, sm, k0, k1);
  if (threadIdx.x == 0) {
    const float a = ordered_to_float(k0);
    const float b = ordered_to_float(k1);
    const float diff = __fsub_rn(b, a);
    float r = __fadd_rn(a, __fmul_rn(diff, g));
    if (g >= 0.5f) r = __fsub_rn(b, __fmul_rn(diff, __fsub_rn(1.0f, g)));
    out[t] = r;
  }
}

}  // namespace vsm
