// query.cu -- text query: scores = features . prompt, top-k voxels.
//
//   vsm_query     SemanticVoxelMap.query_with_embedding      vggt_slam/semantic_voxel.py:97-116
//
// The reference scores ONE prompt per call with a float32 matmul on the CPU and torch.topk; voxel features are not
// normalised (only the prompt is, by the caller: voxel_evaluators.py:66-68).  Here P prompts are scored in one
// pass over the feature sums: score[v,p] = (sum[v] . q[p]) / count[v]; the V x P score matrix is never
// materialised -- each CTA keeps a running top-k per prompt behind a threshold and a last kernel merges the
// per-CTA lists.  Engine 1 (this file) is exact fp32 FMA on the CUDA cores and is HBM-bound for up to ~8
// prompts; engine 2 (query_tc.cu) runs the contraction on the tcgen05 tensor cores.
#include <atomic>
#include "state.cuh"

namespace vsm {

constexpr int kQThreads = 256;
constexpr int kQWarps = kQThreads / 32;
constexpr int kQRowsPerWarp = 8;                        // rows a warp scores between two block barriers
constexpr int kQBatch = kQWarps * kQRowsPerWarp;        // rows per CTA batch
constexpr int kQMaxPB = 8;                              // prompts per pass

__device__ __forceinline__ unsigned long long cand_key(float score, uint32_t rank) {
  return ((unsigned long long)float_to_ordered(score) << 32) | (unsigned long long)(0xFFFFFFFFu - rank);
}

// bitonic sort, descending, n a power of two, all threads of the CTA
__device__ void cta_sort_desc(unsigned long long* buf, int n) {
  for (int k = 2; k <= n; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const unsigned long long a = buf[i], b = buf[ixj];
          const bool desc = (i & k) == 0;
          if (desc ? (a < b) : (a > b)) {
            buf[i] = b;
            buf[ixj] = a;
          }
        }
      }
      __syncthreads();
    }
  }
}

// keep the best k of cnt entries (cnt <= cap, cap a power of two); returns new count; sets thr to the k-th key
__device__ int cta_compact(unsigned long long* buf, int cnt, int cap, int k, unsigned long long* thr) {
  for (int i = cnt + threadIdx.x; i < cap; i += blockDim.x) buf[i] = 0ull;
  __syncthreads();
  cta_sort_desc(buf, cap);
  const int kept = min(cnt, k);
  if (threadIdx.x == 0 && cnt >= k) *thr = buf[k - 1];
  __syncthreads();
  return kept;
}

struct QueryArgs {
  const float* vsum;
  const uint32_t* vcount;
  const uint32_t* rank_of_id;
  uint32_t V;           // rows scored: logical row j is voxel id j * row_stride
  uint32_t row_stride;
  int d;
  int nvec;        // float4 per row
  const float* q;  // [P][d], this pass starts at prompt p0
  int p0;
  int pb;          // prompts in this pass
  int k;
  int cap;         // candidate buffer entries per prompt (power of two >= 2k, >= 2*kQBatch)
  int normalize;
  unsigned long long* cand;  // [P][gridDim.x][k]
  uint32_t* cand_cnt;        // [P][gridDim.x]
};

template <int NV, int PB>
__global__ void __launch_bounds__(kQThreads) query_exact_kernel(QueryArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* sq = reinterpret_cast<float*>(smem_raw);                                    // [PB][d]
  unsigned long long* buf = reinterpret_cast<unsigned long long*>(sq + PB * a.d);   // [PB][cap]
  __shared__ unsigned long long s_thr[PB];
  __shared__ int s_cnt[PB];

  const int lane = lane_id(), warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < PB * a.d; i += blockDim.x) {
    const int p = i / a.d;
    sq[i] = p < a.pb ? a.q[(size_t)(a.p0 + p) * a.d + (i % a.d)] : 0.f;
  }
  if (threadIdx.x < PB) {
    s_thr[threadIdx.x] = 0ull;
    s_cnt[threadIdx.x] = 0;
  }
  __syncthreads();

  const uint32_t n_batches = (a.V + kQBatch - 1) / kQBatch;
  for (uint32_t b = blockIdx.x; b < n_batches; b += gridDim.x) {
    const uint32_t row0 = b * kQBatch + warp * kQRowsPerWarp;
    constexpr int R = (NV <= 2) ? 4 : 2;  // rows in flight per warp
#pragma unroll 1
    for (int rr = 0; rr < kQRowsPerWarp; rr += R) {
      float4 x[R][NV];
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const uint32_t row = row0 + rr + r;
        const uint32_t id = row * a.row_stride;
#pragma unroll
        for (int j = 0; j < NV; ++j) {
          const int c = lane + 32 * j;
          if (row < a.V && c < a.nvec) {
            const uint4 u = ld_stream_v4(a.vsum + (size_t)id * a.d + 4 * c);
            x[r][j] = make_float4(__uint_as_float(u.x), __uint_as_float(u.y), __uint_as_float(u.z), __uint_as_float(u.w));
          } else {
            x[r][j] = make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
      }
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const uint32_t row = row0 + rr + r;
        if (row >= a.V) continue;  // warp-uniform
        const uint32_t id = row * a.row_stride;
        float dot[PB];
        float ss = 0.f;
#pragma unroll
        for (int p = 0; p < PB; ++p) dot[p] = 0.f;
#pragma unroll
        for (int j = 0; j < NV; ++j) {
          const int c = lane + 32 * j;
          if (c < a.nvec) {
            const float4 v = x[r][j];
            ss = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, fmaf(v.w, v.w, ss))));
#pragma unroll
            for (int p = 0; p < PB; ++p) {
              const float4 w = *reinterpret_cast<const float4*>(sq + p * a.d + 4 * c);
              dot[p] = fmaf(v.x, w.x, fmaf(v.y, w.y, fmaf(v.z, w.z, fmaf(v.w, w.w, dot[p]))));
            }
          }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          ss += __shfl_xor_sync(0xffffffffu, ss, o);
#pragma unroll
          for (int p = 0; p < PB; ++p) dot[p] += __shfl_xor_sync(0xffffffffu, dot[p], o);
        }
        if (lane == 0) {
          const float cnt = (float)a.vcount[id];
          const uint32_t rank = a.rank_of_id[id];
          float inv;
          if (a.normalize) {
            const float fn = __fdiv_rn(sqrtf(ss), cnt);
            inv = __fdiv_rn(1.0f, fmaxf(fn, 1e-12f));
          } else {
            inv = 1.0f;
          }
#pragma unroll
          for (int p = 0; p < PB; ++p) {
            if (p >= a.pb) break;
            const float sc = __fmul_rn(__fdiv_rn(dot[p], cnt), inv);
            const unsigned long long key = cand_key(sc, rank);
            if (key > s_thr[p]) {
              const int pos = atomicAdd(&s_cnt[p], 1);
              buf[(size_t)p * a.cap + pos] = key;  // cap >= count before the batch + kQBatch: never overflows
            }
          }
        }
      }
    }
    __syncthreads();
    // compact the prompts whose buffer could overflow in the next batch
    for (int p = 0; p < a.pb; ++p) {
      const int cnt = s_cnt[p];
      if (cnt + kQBatch > a.cap) {
        const int kept = cta_compact(buf + (size_t)p * a.cap, cnt, a.cap, a.k, &s_thr[p]);
        if (threadIdx.x == 0) s_cnt[p] = kept;
        __syncthreads();
      }
    }
  }
  __syncthreads();
  for (int p = 0; p < a.pb; ++p) {
    const int cnt = s_cnt[p];
    const int kept = cta_compact(buf + (size_t)p * a.cap, cnt, a.cap, a.k, &s_thr[p]);
    unsigned long long* dst = a.cand + ((size_t)(a.p0 + p) * gridDim.x + blockIdx.x) * a.k;
    for (int i = threadIdx.x; i < kept; i += blockDim.x) dst[i] = buf[(size_t)p * a.cap + i];
    if (threadIdx.x == 0) a.cand_cnt[(size_t)(a.p0 + p) * gridDim.x + blockIdx.x] = (uint32_t)kept;
    __syncthreads();
  }
}

// one CTA per prompt: merge the per-CTA candidate lists into the final top-k
__global__ void __launch_bounds__(kQThreads) query_merge_kernel(const unsigned long long* __restrict__ cand,
                                                                const uint32_t* __restrict__ cand_cnt, int n_lists, int k,
                                                                int cap, int64_t* __restrict__ idx,
                                                                float* __restrict__ score) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  unsigned long long* buf = reinterpret_cast<unsigned long long*>(smem_raw);  // [cap]
  __shared__ unsigned long long s_thr;
  const int p = blockIdx.x;
  int cnt = 0;
  for (int l = 0; l < n_lists; ++l) {
    const int n = (int)cand_cnt[(size_t)p * n_lists + l];
    const unsigned long long* src = cand + ((size_t)p * n_lists + l) * k;
    if (cnt + n > cap) {
      cnt = cta_compact(buf, cnt, cap, k, &s_thr);
      __syncthreads();
    }
    for (int i = threadIdx.x; i < n; i += blockDim.x) buf[cnt + i] = src[i];
    cnt += n;
    __syncthreads();
  }
  cnt = cta_compact(buf, cnt, cap, k, &s_thr);
  __syncthreads();
  for (int i = threadIdx.x; i < k; i += blockDim.x) {
    if (i < cnt) {
      const unsigned long long key = buf[i];
      idx[(size_t)p * k + i] = (int64_t)(0xFFFFFFFFu - (uint32_t)(key & 0xFFFFFFFFull));
      score[(size_t)p * k + i] = ordered_to_float((uint32_t)(key >> 32));
    } else {
      idx[(size_t)p * k + i] = -1;
      score[(size_t)p * k + i] = __uint_as_float(0x7FC00000u);
    }
  }
}

// one CTA per prompt: top-k of an arbitrary-length key list (the tensor-core engine's re-scored candidates)
__global__ void __launch_bounds__(kQThreads) query_select_kernel(const unsigned long long* __restrict__ keys,
                                                                 const uint32_t* __restrict__ counts, int list_cap, int k,
                                                                 int cap, int64_t* __restrict__ idx,
                                                                 float* __restrict__ score) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  unsigned long long* buf = reinterpret_cast<unsigned long long*>(smem_raw);  // [cap]
  __shared__ unsigned long long s_thr;
  const int p = blockIdx.x;
  const int n = min((int)counts[p], list_cap);
  const unsigned long long* src = keys + (size_t)p * list_cap;
  int cnt = 0;
  for (int base = 0; base < n;) {
    const int room = cap - cnt;
    const int take = min(room, n - base);
    for (int i = threadIdx.x; i < take; i += blockDim.x) buf[cnt + i] = src[base + i];
    cnt += take;
    base += take;
    __syncthreads();
    if (base < n) {
      cnt = cta_compact(buf, cnt, cap, k, &s_thr);
      __syncthreads();
    }
  }
  cnt = cta_compact(buf, cnt, cap, k, &s_thr);
  __syncthreads();
  for (int i = threadIdx.x; i < k; i += blockDim.x) {
    if (i < cnt) {
      const unsigned long long key = buf[i];
      idx[(size_t)p * k + i] = (int64_t)(0xFFFFFFFFu - (uint32_t)(key & 0xFFFFFFFFull));
      score[(size_t)p * k + i] = ordered_to_float((uint32_t)(key >> 32));
    } else {
      idx[(size_t)p * k + i] = -1;
      score[(size_t)p * k + i] = __uint_as_float(0x7FC00000u);
    }
  }
}

int query_select_from_keys(vsm_map* m, const unsigned long long* keys, const uint32_t* counts, int P, int list_cap, int k,
                           int64_t* idx_dev, float* score_dev, cudaStream_t s) {
  (void)m;
  int cap = 2048;
  while (cap < 2 * k) cap <<= 1;
  VSM_CUDA(cudaFuncSetAttribute(query_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, cap * 8));
  query_select_kernel<<<P, kQThreads, (size_t)cap * 8, s>>>(keys, counts, list_cap, k, cap, idx_dev, score_dev);
  VSM_LAUNCHED();
  return VSM_OK;
}

template <int NV>
static int launch_query_exact(const QueryArgs& a, int grid, size_t smem, cudaStream_t s) {
#define VSM_Q_CASE(PB)                                                                                          \
  case PB: {                                                                                                    \
    VSM_CUDA(cudaFuncSetAttribute(query_exact_kernel<NV, PB>, cudaFuncAttributeMaxDynamicSharedMemorySize,     \
                                  (int)smem));                                                                  \
    query_exact_kernel<NV, PB><<<grid, kQThreads, smem, s>>>(a);                                                \
    break;                                                                                                      \
  }
  const int pb_t = a.pb <= 1 ? 1 : a.pb <= 2 ? 2 : a.pb <= 4 ? 4 : 8;
  switch (pb_t) {
    VSM_Q_CASE(1)
    VSM_Q_CASE(2)
    VSM_Q_CASE(4)
    VSM_Q_CASE(8)
  }
#undef VSM_Q_CASE
  VSM_LAUNCHED();
  return VSM_OK;
}

int query_tc(vsm_map* m, const float* q_dev, int P, int k, int normalize, int64_t* idx_dev, float* score_dev,
             cudaStream_t s, bool bf16_shadow);  // query_tc.cu

int query_exact_rows(vsm_map* m, const float* q_dev, int P, int k, int normalize, uint32_t row_stride, uint32_t n_rows,
                     int64_t* idx_dev, float* score_dev, cudaStream_t s);

int query_exact(vsm_map* m, const float* q_dev, int P, int k, int normalize, int64_t* idx_dev, float* score_dev,
                cudaStream_t s) {
  return query_exact_rows(m, q_dev, P, k, normalize, 1u, (uint32_t)m->n_vox, idx_dev, score_dev, s);
}

// scores the voxels with ids 0, row_stride, 2*row_stride, ... (n_rows of them): the whole map for stride 1, a
// sample for the tensor-core engine's thresholds otherwise
int query_exact_rows(vsm_map* m, const float* q_dev, int P, int k, int normalize, uint32_t row_stride, uint32_t n_rows,
                     int64_t* idx_dev, float* score_dev, cudaStream_t s) {
  const uint32_t V = n_rows;
  const int d = m->d;
  int cap = 2 * kQBatch;
  while (cap < 2 * k) cap <<= 1;
  const int n_sm = sm_count();
  const int n_batches = (int)((V + kQBatch - 1) / kQBatch);
  const int grid = std::max(1, std::min(n_batches, n_sm * 2));
  VSM_TRY(m->q_cand.ensure((size_t)P * grid * k * 8, s));
  VSM_TRY(m->q_tmp.ensure((size_t)P * grid * 4, s));
  QueryArgs a;
  a.vsum = m->vsum.as<float>();
  a.vcount = m->vcount.as<uint32_t>();
  a.rank_of_id = m->rank_of_id.as<uint32_t>();
  a.V = V;
  a.row_stride = row_stride;
  a.d = d;
  a.nvec = d / 4;
  a.q = q_dev;
  a.k = k;
  a.cap = cap;
  a.normalize = normalize;
  a.cand = m->q_cand.as<unsigned long long>();
  a.cand_cnt = m->q_tmp.as<uint32_t>();
  const int nv = (a.nvec + 31) / 32;
  for (int p0 = 0; p0 < P; p0 += kQMaxPB) {
    a.p0 = p0;
    a.pb = std::min(kQMaxPB, P - p0);
    const int pb_t = a.pb <= 1 ? 1 : a.pb <= 2 ? 2 : a.pb <= 4 ? 4 : 8;
    const size_t smem = (size_t)pb_t * d * 4 + (size_t)pb_t * cap * 8;
    if (smem > 200 * 1024) {
      set_error("vsm_query: k=%d with d=%d needs %zu bytes of shared memory", k, d, smem);
      return VSM_E_INVALID;
    }
    if (nv <= 1)
      VSM_TRY(launch_query_exact<1>(a, grid, smem, s));
    else if (nv <= 2)
      VSM_TRY(launch_query_exact<2>(a, grid, smem, s));
    else if (nv <= 4)
      VSM_TRY(launch_query_exact<4>(a, grid, smem, s));
    else if (nv <= 8)
      VSM_TRY(launch_query_exact<8>(a, grid, smem, s));
    else if (nv <= 16)
      VSM_TRY(launch_query_exact<16>(a, grid, smem, s));
    else {
      set_error("vsm_query: d=%d too large", d);
      return VSM_E_INVALID;
    }
  }
  int mcap = 1024;
  while (mcap < 2 * k) mcap <<= 1;
  VSM_CUDA(cudaFuncSetAttribute(query_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, mcap * 8));
  query_merge_kernel<<<P, kQThreads, (size_t)mcap * 8, s>>>(a.cand, a.cand_cnt, grid, k, mcap, idx_dev, score_dev);
  VSM_LAUNCHED();
  return VSM_OK;
}

}  // namespace vsm

namespace vsm {
std::atomic<int> g_query_shadow{0};  // "query_shadow": engine 0 (auto) uses the bf16 shadow on large maps
}
using namespace vsm;

extern "C" int vsm_query(vsm_map* m, const float* q_dev, int32_t P, int32_t k, int normalize, int engine,
                         int64_t* idx_dev, float* score_dev, void* stream) {
  if (!m || !q_dev || !idx_dev || !score_dev) {
    set_error("vsm_query: null argument");
    return VSM_E_INVALID;
  }
  if (!m->finalized) {
    set_error("vsm_query: map is not finalised");
    return VSM_E_STATE;
  }
  if (P < 1 || P > VSM_MAX_PROMPTS || k < 1 || k > VSM_MAX_TOPK) {
    set_error("vsm_query: P=%d (1..%d) k=%d (1..%d)", P, VSM_MAX_PROMPTS, k, VSM_MAX_TOPK);
    return VSM_E_INVALID;
  }
  if ((int64_t)k > m->n_vox) {
    // torch.topk raises "selected index k out of range" (semantic_voxel.py:112)
    set_error("vsm_query: top_k=%d exceeds the %lld voxels of the map", k, (long long)m->n_vox);
    return VSM_E_INVALID;
  }
  VSM_CUDA(cudaSetDevice(m->device));
  cudaStream_t s = (cudaStream_t)stream;
  // engine 0 (auto): on maps of some size the tensor-core engine is faster for every P (measured on 10 M voxels:
  // 3.7 ms vs 3.9 ms at P=1, 3.8 ms vs 11 ms at P=8); small maps are latency-bound and take the simpler path
  // engine 3: the tensor-core passes read a bf16 shadow of the sums (half the bytes, twice the tensor rate, a wider
  // selection margin; +2 bytes per value of memory): the same answer, chosen explicitly or with "query_shadow" = 1
  if (engine == 3 || (engine == 0 && g_query_shadow.load() && m->n_vox >= 65536 && m->d % 64 == 0 && m->d <= 1024))
    return query_tc(m, q_dev, P, k, normalize, idx_dev, score_dev, s, true);
  if (engine == 2 || (engine == 0 && m->n_vox >= 65536 && m->d % 32 == 0 && m->d <= 1024))
    return query_tc(m, q_dev, P, k, normalize, idx_dev, score_dev, s, false);
  return query_exact(m, q_dev, P, k, normalize, idx_dev, score_dev, s);
}

extern "C" int vsm_query_stats(const vsm_map* m, int64_t* last_candidates_host, int64_t* fallbacks_host) {
  if (!m) {
    set_error("null map");
    return VSM_E_INVALID;
  }
  if (last_candidates_host) *last_candidates_host = m->tc_last_candidates;
  if (fallbacks_host) *fallbacks_host = m->tc_fallbacks;
  return VSM_OK;
}
