// select.cuh -- exact order statistics on the device with numpy-2 semantics.
//
// np.percentile(x, q) (method 'linear') on float32 data computes the virtual
// index (n-1)*q in float32 and lerps the two neighbouring order statistics as
//   a + (b-a)*g,  or  b - (b-a)*(1-g)  when g >= 0.5       (all float32)
// (vggt_slam/submap.py:38 for the confidence threshold, vggt_slam/map.py:257-258
// for the bbox filter).  The order statistics come from a 3-pass (11+11+10 bit)
// radix select over an order-preserving integer image of the floats; several
// ranks and up to three columns are selected in the same passes.
#include <algorithm>

#include "state.cuh"

namespace vsm {

__device__ __forceinline__ bool sel_valid(const SelSrc& s, int64_t i) {
  if (s.flag_off < 0) return true;
  const uint32_t f = __float_as_uint(s.base[i * s.stride + s.flag_off]);
  return (f & s.flag_need) == s.flag_need;
}

// pass 0: per-column histogram of the top 11 bits (+ NaN count, + n)
__global__ void __launch_bounds__(256) sel_hist0_kernel(SelSrc src, uint32_t* __restrict__ hist /*[ncol][2048]*/,
                                                         SelectState* st, int count_n) {
  __shared__ uint32_t sh[3 * 2048];
  for (int i = threadIdx.x; i < src.ncol * 2048; i += blockDim.x) sh[i] = 0;
  __syncthreads();
  unsigned long long local_n = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < src.n_items;
       i += (int64_t)gridDim.x * blockDim.x) {
    if (!sel_valid(src, i)) continue;
    ++local_n;
    for (int c = 0; c < src.ncol; ++c) {
      const float v = src.base[i * src.stride + c];
      const uint32_t o = float_to_ordered(v);
      if (o == 0xFFFFFFFFu) atomicAdd(&st->nan_count[c], 1u);
      atomicAdd(&sh[c * 2048 + (o >> 21)], 1u);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < src.ncol * 2048; i += blockDim.x)
    if (sh[i]) atomicAdd(&hist[i], sh[i]);
  if (count_n) {
    // warp-reduce then one atomic per warp
    for (int o = 16; o > 0; o >>= 1) local_n += __shfl_xor_sync(0xffffffffu, local_n, o);
    if (lane_id() == 0 && local_n) atomicAdd(&st->n, local_n);
  }
}

// turn percentiles into ranks:  pcts are q/100 as float32 (host computes f32(q)/f32(100))
__global__ void sel_plan_kernel(SelectState* st, int ncol, int npct, float q0, float q1,
                                const unsigned long long* n_dev) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  if (n_dev) st->n = *n_dev;
  const unsigned long long n = st->n;
  st->n_targets = ncol * npct * 2;
  st->targets_per_col = npct * 2;
  for (int j = 0; j < npct; ++j) {
    const float q = j == 0 ? q0 : q1;
    unsigned long long lo = 0, hi = 0;
    float g = 0.f;
    if (n > 0) {
      const float nm1 = (float)(n - 1);
      const float vidx = __fmul_rn(nm1, q);
      if (vidx >= nm1) {
        lo = hi = n - 1;
      } else if (vidx < 0.f) {
        lo = hi = 0;
      } else {
        lo = (unsigned long long)floorf(vidx);
        hi = lo + 1;
        if (hi > n - 1) hi = n - 1;
      }
      g = __fsub_rn(vidx, floorf(vidx));
    }
    for (int c = 0; c < ncol; ++c) {
      const int t = (c * npct + j) * 2;
      st->rank[t] = lo;
      st->rank[t + 1] = hi;
      st->rem[t] = lo;
      st->rem[t + 1] = hi;
      st->prefix[t] = st->prefix[t + 1] = 0;
      st->gamma[t] = st->gamma[t + 1] = g;
    }
  }
}

// one warp per target: find the bin where the running count crosses rem[t]
__global__ void sel_pick_kernel(SelectState* st, const uint32_t* __restrict__ hist, int bins, int shift_bits,
                                int per_col_hist) {
  const int t = blockIdx.x;
  if (t >= st->n_targets) return;
  const int lane = lane_id();
  const uint32_t* h = hist + (size_t)(per_col_hist ? (t / st->targets_per_col) : t) * bins;
  const int per = bins / 32;
  unsigned long long mine = 0;
  for (int i = 0; i < per; ++i) mine += h[lane * per + i];
  unsigned long long incl = mine;
  for (int o = 1; o < 32; o <<= 1) {
    unsigned long long up = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += up;
  }
  const unsigned long long excl = incl - mine;
  const unsigned long long want = st->rem[t];
  const bool here = (want >= excl) && (want < incl);
  if (here) {
    unsigned long long run = excl;
    int bin = lane * per;
    for (int i = 0; i < per; ++i) {
      const unsigned long long c = h[lane * per + i];
      if (want < run + c) {
        bin = lane * per + i;
        break;
      }
      run += c;
    }
    st->prefix[t] = (st->prefix[t] << shift_bits) | (uint32_t)bin;
    st->rem[t] = want - run;
  }
}

// passes 1 and 2: histogram of the next bits of elements whose high bits equal a target's prefix
__global__ void __launch_bounds__(256) sel_histn_kernel(SelSrc src, uint32_t* __restrict__ hist, const SelectState* st,
                                                         int prefix_shift, int bins_shift, uint32_t bins_mask) {
  __shared__ uint32_t s_prefix[kSelMaxTargets];
  __shared__ int s_nt, s_tpc;
  if (threadIdx.x < kSelMaxTargets) s_prefix[threadIdx.x] = st->prefix[threadIdx.x];
  if (threadIdx.x == 0) {
    s_nt = st->n_targets;
    s_tpc = st->targets_per_col;
  }
  __syncthreads();
  const int tpc = s_tpc;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < src.n_items;
       i += (int64_t)gridDim.x * blockDim.x) {
    if (!sel_valid(src, i)) continue;
    for (int c = 0; c < src.ncol; ++c) {
      const uint32_t o = float_to_ordered(src.base[i * src.stride + c]);
      const uint32_t hi = o >> prefix_shift;
      for (int k = 0; k < tpc; ++k) {
        const int t = c * tpc + k;
        if (t < s_nt && hi == s_prefix[t]) atomicAdd(&hist[(size_t)t * (bins_mask + 1) + ((o >> bins_shift) & bins_mask)], 1u);
      }
    }
  }
}

// value[t] from the completed 32-bit prefix; then the numpy lerp per (column, percentile) pair into out[]
__global__ void sel_finish_kernel(SelectState* st, float* __restrict__ out, int n_out) {
  const int j = threadIdx.x;
  if (j >= n_out) return;
  const int t = 2 * j;
  const int col = t / st->targets_per_col;
  float r;
  if (st->n == 0 || st->nan_count[col] > 0) {
    r = __uint_as_float(0x7FC00000u);
  } else {
    const float a = ordered_to_float(st->prefix[t]);
    const float b = ordered_to_float(st->prefix[t + 1]);
    st->value[t] = a;
    st->value[t + 1] = b;
    const float g = st->gamma[t];
    const float diff = __fsub_rn(b, a);
    r = __fadd_rn(a, __fmul_rn(diff, g));
    if (g >= 0.5f) r = __fsub_rn(b, __fmul_rn(diff, __fsub_rn(1.0f, g)));
  }
  out[j] = r;
}


// ---------------------------------------------------------------------------
// bracket select: the same order statistics in ONE pass over the data
// ---------------------------------------------------------------------------
// The bbox filter needs the 0.5 % / 99.5 % order statistics of ~5 M points per axis.  Instead of three radix passes
// over everything:  (1) one CTA per column sorts a hashed sample of 2048 elements in shared memory and reads off,
// per percentile, a bracket [a, b] of sample order statistics 6 sigma either side of the wanted sample rank;
// (2) ONE pass counts, per bracket, the elements below a and collects the elements inside [a, b] (about 2 % of
// the data); (3) one CTA per bracket radix-selects the wanted ranks among the collected elements.  The result is
// the exact order statistic whenever the wanted ranks fall inside the bracket (a miss has probability ~1e-9 per
// call with a random sample; it, a collected list that overflows, or a sample that is too small set *miss and the
// caller repeats the call with the three-pass radix select) -- never an approximation.
constexpr int kBrSample = 2048;
constexpr int kBrLists = 6;  // [col][lo pct, hi pct]

struct BracketState {
  float lo[kBrLists], hi[kBrLists];  // bracket bounds (inclusive); -inf / +inf when the sample rank was clipped
  uint32_t below[kBrLists];          // valid elements < lo
  uint32_t cursor[kBrLists];         // elements collected (can exceed the list capacity: overflow)
  uint32_t miss;
  uint32_t pad[3];
};

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
  x ^= x >> 16;
  x *= 0x7feb352du;
  x ^= x >> 15;
  x *= 0x846ca68bu;
  x ^= x >> 16;
  return x;
}

// grid = ncol CTAs of 1024 threads
__global__ void __launch_bounds__(1024) bracket_sample_kernel(SelSrc src, BracketState* bs, float q0, float q1) {
  __shared__ float sv[kBrSample];
  __shared__ uint32_t s_n;
  const int c = blockIdx.x;
  if (threadIdx.x == 0) s_n = 0u;
  __syncthreads();
  const int64_t n_items = src.n_items;
  const int64_t m = n_items < kBrSample ? n_items : kBrSample;
  const int64_t step = m > 0 ? n_items / m : 1;
  for (int64_t j = threadIdx.x; j < m; j += blockDim.x) {
    const int64_t i = j * step + (int64_t)(hash32((uint32_t)j * 3u + 0x9E3779B9u) % (uint32_t)step);
    if (sel_valid(src, i)) sv[atomicAdd(&s_n, 1u)] = src.base[i * src.stride + c];
  }
  __syncthreads();
  const uint32_t mv = s_n;
  for (uint32_t j = mv + threadIdx.x; j < (uint32_t)kBrSample; j += blockDim.x) sv[j] = __int_as_float(0x7F800000);  // +inf pads
  __syncthreads();
  // bitonic sort, one compare-exchange per thread and step
  for (uint32_t k = 2; k <= (uint32_t)kBrSample; k <<= 1) {
    for (uint32_t j = k >> 1; j > 0; j >>= 1) {
      for (uint32_t t = threadIdx.x; t < (uint32_t)kBrSample / 2; t += blockDim.x) {
        const uint32_t i = 2 * t - (t & (j - 1));  // index with bit j clear
        const uint32_t l = i + j;
        const bool up = (i & k) == 0;
        const float a = sv[i], b = sv[l];
        if ((a > b) == up) {
          sv[i] = b;
          sv[l] = a;
        }
      }
      __syncthreads();
    }
  }
  if (threadIdx.x < 2) {
    const int t = c * 2 + threadIdx.x;
    const float q = threadIdx.x == 0 ? q0 : q1;
    float lo = -__int_as_float(0x7F800000), hi = __int_as_float(0x7F800000);
    if (mv < 64u) {
      atomicOr(&bs->miss, 1u);  // too few valid samples to bracket anything
    } else {
      const float r = q * (float)(mv - 1);
      const float w = 6.0f * sqrtf(fmaxf((float)mv * q * (1.0f - q), 1.0f)) + 2.0f;
      const float ra = floorf(r - w), rb = ceilf(r + w);
      if (ra >= 1.0f) lo = sv[(uint32_t)ra];
      if (rb <= (float)(mv - 2)) hi = sv[(uint32_t)rb];
    }
    bs->lo[t] = lo;
    bs->hi[t] = hi;
  }
}

// one pass: per bracket, count the valid elements below it and collect those inside it
__global__ void __launch_bounds__(256) bracket_collect_kernel(SelSrc src, BracketState* bs, float* __restrict__ lists,
                                                               uint32_t list_cap) {
  __shared__ float s_lo[kBrLists], s_hi[kBrLists];
  __shared__ uint32_t s_below[kBrLists];
  if (threadIdx.x < kBrLists) {
    s_lo[threadIdx.x] = bs->lo[threadIdx.x];
    s_hi[threadIdx.x] = bs->hi[threadIdx.x];
    s_below[threadIdx.x] = 0u;
  }
  __syncthreads();
  float lo[kBrLists], hi[kBrLists];
  uint32_t below[kBrLists];
#pragma unroll
  for (int t = 0; t < kBrLists; ++t) {
    lo[t] = s_lo[t];
    hi[t] = s_hi[t];
    below[t] = 0u;
  }
  const int lane = lane_id();
  const int64_t n_round = (src.n_items + 31) & ~(int64_t)31;
  const float4* p4 = reinterpret_cast<const float4*>(src.base);  // stride 4, flags in .w (the world-point layout)
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += (int64_t)gridDim.x * blockDim.x) {
    bool valid = false;
    float v[3] = {0.f, 0.f, 0.f};
    if (i < src.n_items) {
      const float4 p = p4[i];
      valid = (__float_as_uint(p.w) & src.flag_need) == src.flag_need;
      v[0] = p.x;
      v[1] = p.y;
      v[2] = p.z;
    }
#pragma unroll
    for (int t = 0; t < kBrLists; ++t) {
      const float x = v[t >> 1];
      below[t] += (valid && x < lo[t]) ? 1u : 0u;
      const bool in = valid && x >= lo[t] && x <= hi[t];
      const unsigned mk = __ballot_sync(0xffffffffu, in);
      if (mk) {
        uint32_t base = 0;
        const int leader = __ffs(mk) - 1;
        if (lane == leader) base = atomicAdd(&bs->cursor[t], (uint32_t)__popc(mk));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (in) {
          const uint32_t pos = base + (uint32_t)__popc(mk & ((1u << lane) - 1u));
          if (pos < list_cap) lists[(size_t)t * list_cap + pos] = x;
        }
      }
    }
  }
#pragma unroll
  for (int t = 0; t < kBrLists; ++t) {
    uint32_t b = below[t];
    for (int o = 16; o > 0; o >>= 1) b += __shfl_xor_sync(0xffffffffu, b, o);
    if (lane == 0 && b) atomicAdd(&s_below[t], b);
  }
  __syncthreads();
  if (threadIdx.x < kBrLists && s_below[threadIdx.x]) atomicAdd(&bs->below[threadIdx.x], s_below[threadIdx.x]);
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

size_t bracket_scratch_bytes(int64_t n_items) {
  const uint32_t cap = (uint32_t)std::max<int64_t>(n_items / 8, 4096);
  return 256 + (size_t)kBrLists * cap * sizeof(float);
}

// src must be the world-point layout (float4 per element, flags in .w), 3 columns, 2 percentiles.
// scratch: bracket_scratch_bytes(n_items) bytes.  out_dev[6] laid out [col][pct]; *miss_dev is OR-ed with 1 when the
// result is not valid and the caller must fall back to run_percentiles*.
int run_percentiles_bracket(void* scratch, const SelSrc& src, float q0, float q1, float* out_dev,
                            const unsigned long long* n_dev, uint32_t* miss_dev, cudaStream_t s) {
  if (src.ncol != 3 || src.stride != 4 || src.flag_off != 3) {
    set_error("run_percentiles_bracket: unsupported layout");
    return VSM_E_INVALID;
  }
  BracketState* bs = reinterpret_cast<BracketState*>(scratch);
  static_assert(sizeof(BracketState) <= 256, "bracket state fits its slot");
  float* lists = reinterpret_cast<float*>(reinterpret_cast<char*>(scratch) + 256);
  const uint32_t cap = (uint32_t)std::max<int64_t>(src.n_items / 8, 4096);
  VSM_CUDA(cudaMemsetAsync(bs, 0, sizeof(BracketState), s));
  bracket_sample_kernel<<<3, 1024, 0, s>>>(src, bs, q0, q1);
  VSM_LAUNCHED();
  bracket_collect_kernel<<<grid_for(src.n_items, 256, 148 * 8), 256, 0, s>>>(src, bs, lists, cap);
  VSM_LAUNCHED();
  bracket_resolve_kernel<<<kBrLists, 1024, 0, s>>>(bs, lists, cap, n_dev, q0, q1, out_dev, miss_dev);
  VSM_LAUNCHED();
  return VSM_OK;
}

constexpr size_t kSelHistWords = (size_t)3 * 2048 + (size_t)kSelMaxTargets * 2048 + (size_t)kSelMaxTargets * 1024;

// process-wide scratch per device (never freed): select state, histograms, 16 result floats
int select_scratch(SelectState** st, uint32_t** hist, float** out) {
  static void* bufs[64] = {nullptr};
  int dev = 0;
  VSM_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) {
    set_error("select_scratch: device ordinal %d out of range", dev);
    return VSM_E_INVALID;
  }
  const size_t st_bytes = (sizeof(SelectState) + 255) & ~(size_t)255;
  if (!bufs[dev]) VSM_CUDA(cudaMalloc(&bufs[dev], st_bytes + kSelHistWords * sizeof(uint32_t) + 256));
  char* b = (char*)bufs[dev];
  *st = (SelectState*)b;
  *hist = (uint32_t*)(b + st_bytes);
  *out = (float*)(b + st_bytes + kSelHistWords * sizeof(uint32_t));
  return VSM_OK;
}

int run_percentiles(SelectState* st, uint32_t* h0, const SelSrc& src, int npct, float q0, float q1, float* out_dev,
                    cudaStream_t s) {
  if (src.ncol < 1 || src.ncol > 3 || npct < 1 || npct > 2) {
    set_error("run_percentiles: unsupported shape");
    return VSM_E_INVALID;
  }
  const int nt = src.ncol * npct * 2;
  const size_t hist_words = kSelHistWords;
  uint32_t* h1 = h0 + 3 * 2048;
  uint32_t* h2 = h1 + kSelMaxTargets * 2048;
  VSM_CUDA(cudaMemsetAsync(st, 0, sizeof(SelectState), s));
  VSM_CUDA(cudaMemsetAsync(h0, 0, hist_words * sizeof(uint32_t), s));
  const int grid = grid_for(src.n_items, 256, 148 * 8);
  sel_hist0_kernel<<<grid, 256, 0, s>>>(src, h0, st, 1);
  VSM_LAUNCHED();
  sel_plan_kernel<<<1, 32, 0, s>>>(st, src.ncol, npct, q0, q1, nullptr);
  VSM_LAUNCHED();
  sel_pick_kernel<<<nt, 32, 0, s>>>(st, h0, 2048, 11, 1);
  VSM_LAUNCHED();
  sel_histn_kernel<<<grid, 256, 0, s>>>(src, h1, st, 21, 10, 2047u);
  VSM_LAUNCHED();
  sel_pick_kernel<<<nt, 32, 0, s>>>(st, h1, 2048, 11, 0);
  VSM_LAUNCHED();
  sel_histn_kernel<<<grid, 256, 0, s>>>(src, h2, st, 10, 0, 1023u);
  VSM_LAUNCHED();
  sel_pick_kernel<<<nt, 32, 0, s>>>(st, h2, 1024, 10, 0);
  VSM_LAUNCHED();
  sel_finish_kernel<<<1, 32, 0, s>>>(st, out_dev, src.ncol * npct);
  VSM_LAUNCHED();
  return VSM_OK;
}


int select_reset(SelectState* st, uint32_t* hist, cudaStream_t s) {
  VSM_CUDA(cudaMemsetAsync(st, 0, sizeof(SelectState), s));
  VSM_CUDA(cudaMemsetAsync(hist, 0, kSelHistWords * sizeof(uint32_t), s));
  return VSM_OK;
}

int run_percentiles_after_hist0(SelectState* st, uint32_t* h0, const SelSrc& src, int npct, float q0, float q1,
                                float* out_dev, const unsigned long long* n_dev, cudaStream_t s) {
  if (src.ncol < 1 || src.ncol > 3 || npct < 1 || npct > 2) {
    set_error("run_percentiles: unsupported shape");
    return VSM_E_INVALID;
  }
  const int nt = src.ncol * npct * 2;
  uint32_t* h1 = h0 + 3 * 2048;
  uint32_t* h2 = h1 + kSelMaxTargets * 2048;
  const int grid = grid_for(src.n_items, 256, 148 * 8);
  sel_plan_kernel<<<1, 32, 0, s>>>(st, src.ncol, npct, q0, q1, n_dev);
  VSM_LAUNCHED();
  sel_pick_kernel<<<nt, 32, 0, s>>>(st, h0, 2048, 11, 1);
  VSM_LAUNCHED();
  sel_histn_kernel<<<grid, 256, 0, s>>>(src, h1, st, 21, 10, 2047u);
  VSM_LAUNCHED();
  sel_pick_kernel<<<nt, 32, 0, s>>>(st, h1, 2048, 11, 0);
  VSM_LAUNCHED();
  sel_histn_kernel<<<grid, 256, 0, s>>>(src, h2, st, 10, 0, 1023u);
  VSM_LAUNCHED();
  sel_pick_kernel<<<nt, 32, 0, s>>>(st, h2, 1024, 10, 0);
  VSM_LAUNCHED();
  sel_finish_kernel<<<1, 32, 0, s>>>(st, out_dev, src.ncol * npct);
  VSM_LAUNCHED();
  return VSM_OK;
}

}  // namespace vsm
