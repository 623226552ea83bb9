// select.cuh -- exact order statistics on the device with numpy-2 semantics.
//
// np.percentile(x, q) (method 'linear') on float32 data computes the virtual
// index (n-1)*q in float32 and lerps the two neighbouring order statistics as
//   a + (b-a)*g,  or  b - (b-a)*(1-g)  when g >= 0.5       (all float32)
// (vggt_slam/submap.py:38 for the confidence threshold, vggt_slam/map.py:257-258
// for the bbox filter).  The order statistics come from a 3-pass (11+11+10 bit)
// radix select over an order-preserving integer image of the floats; several
// ranks and up to three columns are selected in the same passes.
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

// The same pass specialised for the world-point layout (float4 per element: x, y, z, flags; 3 columns x 4 targets):
// ONE 128-bit load per element instead of four scalar ones, the twelve prefixes in registers, and the matches of a
// warp grouped (match.any on target and digit) so that neighbouring pixels on the same wall add to a bin once.
__global__ void __launch_bounds__(256) sel_histn4_kernel(const float4* __restrict__ pts, int64_t n_items, uint32_t flag_need,
                                                          uint32_t* __restrict__ hist, const SelectState* st,
                                                          int prefix_shift, int bins_shift, uint32_t bins_mask) {
  uint32_t prefix[12];
#pragma unroll
  for (int t = 0; t < 12; ++t) prefix[t] = st->prefix[t];
  const int lane = lane_id();
  const int64_t n_round = (n_items + 31) & ~(int64_t)31;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += (int64_t)gridDim.x * blockDim.x) {
    float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < n_items) p = pts[i];
    const bool valid = (__float_as_uint(p.w) & flag_need) == flag_need && i < n_items;
    const float v[3] = {p.x, p.y, p.z};
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const uint32_t o = float_to_ordered(v[c]);
      const uint32_t hi = o >> prefix_shift;
      const uint32_t dig = (o >> bins_shift) & bins_mask;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int t = c * 4 + k;
        const bool hit = valid && hi == prefix[t];
        const unsigned any = __ballot_sync(0xffffffffu, hit);
        if (any == 0u) continue;  // warp-uniform
        const unsigned grp = __match_any_sync(0xffffffffu, hit ? dig : 0xFFFFFFFFu);
        if (hit && lane == __ffs(grp) - 1) atomicAdd(&hist[(size_t)t * (bins_mask + 1) + dig], (uint32_t)__popc(grp));
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
  const int grid = grid_for(src.n_items, 256, sm_count() * 8);
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
  // state and histograms sit in one block (select_scratch): one memset when they are contiguous
  const char* a = reinterpret_cast<const char*>(st);
  const char* b = reinterpret_cast<const char*>(hist);
  if (b > a && (size_t)(b - a) <= 4096) {
    VSM_CUDA(cudaMemsetAsync(st, 0, (size_t)(b - a) + kSelHistWords * sizeof(uint32_t), s));
    return VSM_OK;
  }
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
  const int grid = grid_for(src.n_items, 256, sm_count() * 8);
  sel_plan_kernel<<<1, 32, 0, s>>>(st, src.ncol, npct, q0, q1, n_dev);
  VSM_LAUNCHED();
  sel_pick_kernel<<<nt, 32, 0, s>>>(st, h0, 2048, 11, 1);
  VSM_LAUNCHED();
  const bool world_layout = src.ncol == 3 && npct == 2 && src.stride == 4 && src.flag_off == 3 &&
                            (reinterpret_cast<uintptr_t>(src.base) & 15u) == 0;
  const float4* p4 = reinterpret_cast<const float4*>(src.base);
  if (world_layout)
    sel_histn4_kernel<<<grid, 256, 0, s>>>(p4, src.n_items, src.flag_need, h1, st, 21, 10, 2047u);
  else
    sel_histn_kernel<<<grid, 256, 0, s>>>(src, h1, st, 21, 10, 2047u);
  VSM_LAUNCHED();
  sel_pick_kernel<<<nt, 32, 0, s>>>(st, h1, 2048, 11, 0);
  VSM_LAUNCHED();
  if (world_layout)
    sel_histn4_kernel<<<grid, 256, 0, s>>>(p4, src.n_items, src.flag_need, h2, st, 10, 0, 1023u);
  else
    sel_histn_kernel<<<grid, 256, 0, s>>>(src, h2, st, 10, 0, 1023u);
  VSM_LAUNCHED();
  sel_pick_kernel<<<nt, 32, 0, s>>>(st, h2, 1024, 10, 0);
  VSM_LAUNCHED();
  sel_finish_kernel<<<1, 32, 0, s>>>(st, out_dev, src.ncol * npct);
  VSM_LAUNCHED();
  return VSM_OK;
}

// ---------------------------------------------------------------------------
// fast path for the world-point layout: the one-block steps merged, one warp per target
// ---------------------------------------------------------------------------
// Lane l owns bins [l*PER, (l+1)*PER) (vector loads, all in flight at once); a warp scan finds the lane whose range
// holds the wanted rank, that lane walks its own bins.  Returns false when no bin holds the rank (empty input).
template <int PER>
__device__ __forceinline__ bool warp_pick(const uint32_t* __restrict__ h, unsigned long long want, uint32_t& bin,
                                          unsigned long long& rem) {
  const int lane = lane_id();
  uint32_t v[PER];
  const uint4* p = reinterpret_cast<const uint4*>(h) + lane * (PER / 4);
#pragma unroll
  for (int i = 0; i < PER / 4; ++i) {
    const uint4 q = p[i];
    v[4 * i] = q.x;
    v[4 * i + 1] = q.y;
    v[4 * i + 2] = q.z;
    v[4 * i + 3] = q.w;
  }
  unsigned long long mine = 0;
#pragma unroll
  for (int i = 0; i < PER; ++i) mine += v[i];
  unsigned long long incl = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned long long up = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += up;
  }
  const unsigned long long excl = incl - mine;
  const bool here = (want >= excl) && (want < incl);
  uint32_t b = 0;
  unsigned long long run = excl;
  if (here) {
    bool done = false;
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      if (!done && want < run + v[i]) {
        b = (uint32_t)(lane * PER + i);
        done = true;
      }
      if (!done) run += v[i];
    }
  }
  const unsigned m = __ballot_sync(0xffffffffu, here);
  if (m == 0u) return false;
  const int src = __ffs(m) - 1;
  bin = __shfl_sync(0xffffffffu, b, src);
  rem = want - __shfl_sync(0xffffffffu, run, src);
  return true;
}

__device__ __forceinline__ void sel_write_alias(SelectState* st, const uint32_t* s_pref, int tpc, int nt) {
  const int t = threadIdx.x;
  if (t < nt) {
    int a = t;
    for (int u = (t / tpc) * tpc; u < t; ++u)
      if (s_pref[u] == s_pref[t]) {
        a = u;
        break;
      }
    st->alias[t] = a;
  }
}

// sel_plan + the first pick (3 columns x 2 percentiles x 2 neighbouring ranks = 12 targets, one warp each)
__global__ void __launch_bounds__(384) sel_planpick0_kernel(SelectState* st, const uint32_t* __restrict__ h0, float q0, float q1,
                                                            const unsigned long long* n_dev) {
  __shared__ uint32_t s_pref[kSelMaxTargets];
  const int t = threadIdx.x >> 5, lane = lane_id();
  const unsigned long long n = n_dev ? *n_dev : st->n;
  const int c = t >> 2, j = (t >> 1) & 1, which = t & 1;
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
  const unsigned long long want = which ? hi : lo;
  uint32_t bin = 0;
  unsigned long long rem = want;
  const bool found = warp_pick<64>(h0 + (size_t)c * 2048, want, bin, rem);
  if (lane == 0) {
    st->rank[t] = want;
    st->gamma[t] = g;
    st->prefix[t] = found ? bin : 0u;
    st->rem[t] = found ? rem : want;
    s_pref[t] = found ? bin : 0u;
  }
  if (threadIdx.x == 0) {
    st->n = n;
    st->n_targets = 12;
    st->targets_per_col = 4;
  }
  __syncthreads();
  sel_write_alias(st, s_pref, 4, 12);
}

// passes 1 and 2 over the world points; targets whose prefix equals an earlier target's (the two neighbouring ranks
// of a percentile nearly always do) are skipped: their pick reads the earlier target's histogram
__global__ void __launch_bounds__(256) sel_histn4a_kernel(const float4* __restrict__ pts, int64_t n_items, uint32_t flag_need,
                                                           uint32_t* __restrict__ hist, const SelectState* st,
                                                           int prefix_shift, int bins_shift, uint32_t bins_mask) {
  uint32_t prefix[12];
  bool own[12];
#pragma unroll
  for (int t = 0; t < 12; ++t) {
    prefix[t] = st->prefix[t];
    own[t] = st->alias[t] == t;
  }
  const int lane = lane_id();
  const int64_t n_round = (n_items + 31) & ~(int64_t)31;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += (int64_t)gridDim.x * blockDim.x) {
    float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < n_items) p = pts[i];
    const bool valid = (__float_as_uint(p.w) & flag_need) == flag_need && i < n_items;
    const float v[3] = {p.x, p.y, p.z};
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const uint32_t o = float_to_ordered(v[c]);
      const uint32_t hi = o >> prefix_shift;
      const uint32_t dig = (o >> bins_shift) & bins_mask;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int t = c * 4 + k;
        if (!own[t]) continue;  // uniform
        const bool hit = valid && hi == prefix[t];
        const unsigned any = __ballot_sync(0xffffffffu, hit);
        if (any == 0u) continue;  // warp-uniform
        const unsigned grp = __match_any_sync(0xffffffffu, hit ? dig : 0xFFFFFFFFu);
        if (hit && lane == __ffs(grp) - 1) atomicAdd(&hist[(size_t)t * (bins_mask + 1) + dig], (uint32_t)__popc(grp));
      }
    }
  }
}

// pick of pass 1 / pass 2 for all 12 targets; FINISH: + sel_finish (values from the completed prefixes, numpy lerp)
template <int PER, bool FINISH>
__global__ void __launch_bounds__(384) sel_pickn_kernel(SelectState* st, const uint32_t* __restrict__ hist, int shift_bits,
                                                        float* __restrict__ out, int n_out) {
  __shared__ uint32_t s_pref[kSelMaxTargets];
  const int t = threadIdx.x >> 5, lane = lane_id();
  int a = 0;
  unsigned long long want = 0;
  uint32_t pre = 0;
  if (lane == 0) {
    a = st->alias[t];
    want = st->rem[t];
    pre = st->prefix[t];
  }
  a = __shfl_sync(0xffffffffu, a, 0);
  want = __shfl_sync(0xffffffffu, want, 0);
  uint32_t bin = 0;
  unsigned long long rem = want;
  const bool found = warp_pick<PER>(hist + (size_t)a * (32 * PER), want, bin, rem);
  if (lane == 0) {
    const uint32_t np = found ? ((pre << shift_bits) | bin) : pre;
    st->prefix[t] = np;
    st->rem[t] = found ? rem : want;
    s_pref[t] = np;
  }
  __syncthreads();
  if (!FINISH) {
    sel_write_alias(st, s_pref, 4, 12);
  } else {
    const int j = threadIdx.x;
    if (j < n_out) {
      const int tt = 2 * j;
      const int col = tt / 4;
      float r;
      if (st->n == 0 || st->nan_count[col] > 0) {
        r = __uint_as_float(0x7FC00000u);
      } else {
        const float va = ordered_to_float(s_pref[tt]);
        const float vb = ordered_to_float(s_pref[tt + 1]);
        st->value[tt] = va;
        st->value[tt + 1] = vb;
        const float g = st->gamma[tt];
        const float diff = __fsub_rn(vb, va);
        r = __fadd_rn(va, __fmul_rn(diff, g));
        if (g >= 0.5f) r = __fsub_rn(vb, __fmul_rn(diff, __fsub_rn(1.0f, g)));
      }
      out[j] = r;
    }
  }
}

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (!cached[dev]) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

int run_percentiles_world_fast(SelectState* st, uint32_t* h0, const float4* pw, int64_t n_items, uint32_t flag_need, float q0,
                               float q1, float* out_dev, const unsigned long long* n_dev, cudaStream_t s) {
  uint32_t* h1 = h0 + 3 * 2048;
  uint32_t* h2 = h1 + kSelMaxTargets * 2048;
  const int grid = grid_for(n_items, 256, sm_count() * 8);
  sel_planpick0_kernel<<<1, 384, 0, s>>>(st, h0, q0, q1, n_dev);
  VSM_LAUNCHED();
  sel_histn4a_kernel<<<grid, 256, 0, s>>>(pw, n_items, flag_need, h1, st, 21, 10, 2047u);
  VSM_LAUNCHED();
  sel_pickn_kernel<64, false><<<1, 384, 0, s>>>(st, h1, 11, nullptr, 0);
  VSM_LAUNCHED();
  sel_histn4a_kernel<<<grid, 256, 0, s>>>(pw, n_items, flag_need, h2, st, 10, 0, 1023u);
  VSM_LAUNCHED();
  sel_pickn_kernel<32, true><<<1, 384, 0, s>>>(st, h2, 10, out_dev, 6);
  VSM_LAUNCHED();
  return VSM_OK;
}

}  // namespace vsm
