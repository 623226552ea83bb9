// fuse.cu -- the fusion hot path: world-frame transform, confidence / outlier
// filters, voxel keys, submap-local hash with warp-level key dedup, global hash
// merge, voxel-sorted point lists and the embedding accumulate kernel.
//
// Reference behaviour restated here:
//   Submap.get_semantic_voxel_in_world_frame   vggt_slam/submap.py:246-293
//   GraphMap.build_semantic_voxel_map          vggt_slam/map.py:196-291 (per-submap loop, three filters)
//                                              vggt_slam/map.py:351-362 (global voxelisation, numpy branch)
//
// One fuse call queues all of its kernels back to back on the caller's stream and synchronises once, at the
// end, to hand the counters back.  Sizes that are only known on the device (survivors, distinct voxels) are
// read by the kernels from the counter block; buffers are sized by upper bounds; a capacity check on the
// device aborts the call before anything global is modified if the map has to grow (the host then grows it
// and repeats the call).
//   world_points      px -> (x,y,z,flags) float4: conf mask, stride grid, f64 transform, finite test,
//                     + top-11-bit histograms of the three axes for the percentile select
//   [filters]         radix-select passes 1,2 -> bbox; coarse-cell hash count -> isolation filter
//   fine_insert       packed voxel key -> submap-local hash (count + frame mask), point -> slot
//   count_new_decide  exact capacity check; the last block sets the abort flag or reserves the log entries
//   compact_merge     one thread per DISTINCT voxel (V_sub, not N): dense local id, segment of the sorted list,
//                     insert into the global hash, counts, contributor log
//   scatter           counting sort of the points by local voxel -> packed (voxel id, pixel) entries
//   tables_cleanup    resets the touched slots of both submap-local tables
//   accumulate        warps walk 32-entry chunks of the sorted list, sum embedding rows in fp32 registers
//                     (FHADD.BF16) and flush with vector REDs at voxel boundaries
#include <algorithm>
#include <atomic>

#include <cuda.h>

#include "bracket.cuh"
#include "hash.cuh"

namespace vsm {

constexpr uint32_t kFuseForceRadix = 1u << 30;  // internal flag: this call must use the three-pass radix select
constexpr uint32_t kFuseForceBigTables = 1u << 28;  // internal flag: this call must use the full-size submap-local tables
std::atomic<int> g_small_tables{1};             // "small_tables": size the submap-local tables from earlier calls
extern std::atomic<int> g_query_shadow;          // query.cu
std::atomic<int> g_select_mode{0};              // 0 default (= 3), 1 radix, 2 bracket in the world kernel, 3 deferred box ("select_mode")
// preparation kernels (vsm_set_option "prep_variant"): bit 0 = a warp takes a 4x8 pixel patch instead of 32 pixels of a
// row; bit 1 = read the frame-mask word before the atomic OR; bit 2 = select with merged one-block steps; bit 3 = one-table
// preparation (coarse cells counted per distinct voxel, ordinals instead of a second round of atomics)
std::atomic<int> g_prep_variant{13};
std::atomic<int> g_range_policy{0};             // "coord_range_policy": 0 = fail the call (VSM_E_COORD_RANGE), 1 = drop + count
std::atomic<long long> g_select_misses{0};      // calls repeated because the bracket select could not answer
std::atomic<long long> g_capacity_retries{0};   // calls repeated after the map / contributor log had to grow
std::atomic<long long> g_table_retries{0};      // calls repeated because the table prefix sized from earlier calls was too small
std::atomic<long long> g_early_collects{0};     // queued calls collected by a later submit (full ring, log growth)
std::atomic<int> g_host_zero_copy{1};            // "host_zero_copy": pinned host embeddings are read in place by the accumulate kernel
std::atomic<int> g_acc_ctas_per_sm{2};          // resident CTAs per SM of the voxel-sorted accumulate kernel ("acc_ctas_per_sm")

constexpr uint32_t PF_SEL = 1u;     // conf >= thr, on the stride grid, frame < end_idx
constexpr uint32_t PF_FINITE = 2u;  // world point (and embedding row, if a mask was given) finite

// ---------------------------------------------------------------------------
// world points (+ pass 0 of the percentile select)
// ---------------------------------------------------------------------------
struct WorldArgs {
  const float* pts;
  const float* conf;
  const uint8_t* emb_ok;
  float4* pw;
  uint32_t n_px;
  uint32_t H, W, stride;
  float thr;
  uint32_t* hist0;  // [3][2048] top-11-bit histograms of the finite selected points, or nullptr
};

__device__ __forceinline__ float4 world_one(const WorldArgs& a, const HMat& Hm, uint32_t pix, float px, float py,
                                            float pz, float c, uint32_t& flags) {
  flags = 0u;
  float x = 0.f, y = 0.f, z = 0.f;
  bool on_grid = true;
  if (a.stride > 1) {
    const uint32_t w = pix % a.W;
    const uint32_t h = (pix / a.W) % a.H;
    on_grid = (w % a.stride == 0) && (h % a.stride == 0);
  }
  if (on_grid && c >= a.thr) {
    transform_f32(Hm, px, py, pz, x, y, z);
    flags = PF_SEL;
    if (finite3(x, y, z) && (a.emb_ok == nullptr || a.emb_ok[pix] != 0)) flags |= PF_FINITE;
  }
  return make_float4(x, y, z, __uint_as_float(flags));
}

// warp-aggregated shared-memory histogram update: neighbouring pixels share their high bits
__device__ __forceinline__ void hist0_add(uint32_t* sh, const float4& p, bool valid) {
  const float v[3] = {p.x, p.y, p.z};
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const uint32_t bin = valid ? (float_to_ordered(v[c]) >> 21) : 0xFFFFFFFFu;
    const unsigned grp = __match_any_sync(0xffffffffu, bin);
    if (valid && lane_id() == __ffs(grp) - 1) atomicAdd(&sh[c * 2048 + bin], (uint32_t)__popc(grp));
  }
}

// One CTA per bracket (axis x {low, high} percentile): transform a hashed sample of the RAW points, keep the valid
// ones' coordinate on this axis as ordered keys in shared memory, and read off the bracket [lo, hi] -- the sample order
// statistics 5 sigma either side of the wanted sample rank -- with a two-target radix select (bracket.cuh).
__global__ void __launch_bounds__(1024) bracket_sample_kernel(WorldArgs a, HMat Hm, BracketState* bs, float q0, float q1) {
  __shared__ uint32_t sv[kBrSample];
  __shared__ uint32_t hist[2 * 2048];
  __shared__ uint32_t sm[8];
  __shared__ uint32_t s_n;
  const int t = blockIdx.x, c = t >> 1;
  if (threadIdx.x == 0) s_n = 0u;
  __syncthreads();
  const uint32_t m = a.n_px < (uint32_t)kBrSample ? a.n_px : (uint32_t)kBrSample;
  const uint32_t step = m > 0 ? a.n_px / m : 1u;
#pragma unroll 4
  for (uint32_t j = threadIdx.x; j < m; j += blockDim.x) {
    const uint32_t pix = j * step + hash32(j * 3u + 0x9E3779B9u) % step;
    uint32_t f;
    const float4 w = world_one(a, Hm, pix, a.pts[3 * (size_t)pix], a.pts[3 * (size_t)pix + 1], a.pts[3 * (size_t)pix + 2],
                               a.conf[pix], f);
    if ((f & (PF_SEL | PF_FINITE)) == (PF_SEL | PF_FINITE))
      sv[atomicAdd(&s_n, 1u)] = float_to_ordered(c == 0 ? w.x : (c == 1 ? w.y : w.z));
  }
  __syncthreads();
  const uint32_t mv = s_n;
  float lo = -__int_as_float(0x7F800000), hi = __int_as_float(0x7F800000);
  if (mv < 64u) {
    if (threadIdx.x == 0) atomicOr(&bs->miss, 1u);  // too few valid samples to bracket anything
  } else {
    const float q = (t & 1) ? q1 : q0;
    const float r = q * (float)(mv - 1);
    const float w = 5.0f * sqrtf(fmaxf((float)mv * q * (1.0f - q), 1.0f)) + 2.0f;
    const float ra = floorf(r - w), rb = ceilf(r + w);
    const bool has_lo = ra >= 1.0f, has_hi = rb <= (float)(mv - 2);
    uint32_t k0, k1;
    cta_select2(sv, mv, has_lo ? (uint32_t)ra : 0u, has_hi ? (uint32_t)rb : mv - 1u, hist, sm, k0, k1);
    if (has_lo) lo = ordered_to_float(k0);
    if (has_hi) hi = ordered_to_float(k1);
  }
  if (threadIdx.x == 0) {
    bs->lo[t] = lo;
    bs->hi[t] = hi;
  }
}

// VEC4: 4 pixels per thread -- three 128-bit loads of xyz, one of conf, four 128-bit stores.
// MODE 1: + pass 0 of the radix select (histogram of the top 11 bits); MODE 2: + the bracket select's collect step.
template <bool VEC4, int MODE>
__global__ void __launch_bounds__(256) world_points_kernel(WorldArgs a, HMat Hm, FuseCounters* ctr, BracketArgs br) {
  constexpr bool HIST = MODE == 1;
  __shared__ uint32_t sh[HIST ? 3 * 2048 : 1];
  __shared__ float s_stage[MODE == 2 ? kBrWarps * kBrLists * kBrStage : 1];
  __shared__ uint32_t s_stage_n[MODE == 2 ? kBrWarps * kBrLists : 1];
  float* const my_stage = s_stage + (MODE == 2 ? (threadIdx.x >> 5) * kBrLists * kBrStage : 0);
  uint32_t* const my_stage_n = s_stage_n + (MODE == 2 ? (threadIdx.x >> 5) * kBrLists : 0);
  BracketLocal bl;
  if (MODE == 2) {
    bracket_load(bl, br.bs);
    if (lane_id() < kBrLists) my_stage_n[lane_id()] = 0u;
    __syncwarp();
  }
  if (HIST) {
    for (int i = threadIdx.x; i < 3 * 2048; i += blockDim.x) sh[i] = 0u;
    __syncthreads();
  }
  unsigned n_sel = 0, n_fin = 0;
  constexpr uint32_t PPT = VEC4 ? 4u : 1u;
  const uint32_t n_items = (a.n_px + PPT - 1) / PPT;
  const uint32_t n_round = (n_items + 31u) & ~31u;  // warp-uniform trip count (match.any inside)
  for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < n_round; t += gridDim.x * blockDim.x) {
    float4 o[PPT];
    uint32_t f[PPT];
#pragma unroll
    for (uint32_t j = 0; j < PPT; ++j) {
      o[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      f[j] = 0u;
    }
    if (t < n_items) {
      if (VEC4 && 4u * t + 3u < a.n_px) {
        const float4* p4 = reinterpret_cast<const float4*>(a.pts) + 3 * (size_t)t;
        const uint4 r0 = ld_stream_v4(p4), r1 = ld_stream_v4(p4 + 1), r2 = ld_stream_v4(p4 + 2);
        const uint4 rc = ld_stream_v4(reinterpret_cast<const float4*>(a.conf) + t);
        const float v[12] = {__uint_as_float(r0.x), __uint_as_float(r0.y), __uint_as_float(r0.z),
                             __uint_as_float(r0.w), __uint_as_float(r1.x), __uint_as_float(r1.y),
                             __uint_as_float(r1.z), __uint_as_float(r1.w), __uint_as_float(r2.x),
                             __uint_as_float(r2.y), __uint_as_float(r2.z), __uint_as_float(r2.w)};
        const float c[4] = {__uint_as_float(rc.x), __uint_as_float(rc.y), __uint_as_float(rc.z),
                            __uint_as_float(rc.w)};
#pragma unroll
        for (uint32_t j = 0; j < 4; ++j) {
          o[j] = world_one(a, Hm, 4 * t + j, v[3 * j], v[3 * j + 1], v[3 * j + 2], c[j], f[j]);
          a.pw[4 * (size_t)t + j] = o[j];
        }
      } else {
#pragma unroll
        for (uint32_t j = 0; j < PPT; ++j) {
          const uint32_t pix = PPT * t + j;
          if (pix < a.n_px) {
            o[j] = world_one(a, Hm, pix, a.pts[3 * (size_t)pix], a.pts[3 * (size_t)pix + 1],
                             a.pts[3 * (size_t)pix + 2], a.conf[pix], f[j]);
            a.pw[pix] = o[j];
          }
        }
      }
    }
#pragma unroll
    for (uint32_t j = 0; j < PPT; ++j) {
      n_sel += (f[j] & PF_SEL) ? 1u : 0u;
      n_fin += (f[j] & PF_FINITE) ? 1u : 0u;
      if (HIST) hist0_add(sh, o[j], (f[j] & PF_FINITE) != 0u);
    }
    if (MODE == 2) {
      bool valid[PPT];
#pragma unroll
      for (uint32_t j = 0; j < PPT; ++j) valid[j] = (f[j] & (PF_SEL | PF_FINITE)) == (PF_SEL | PF_FINITE);
      bracket_collect<(int)PPT>(bl, br, o, valid, my_stage, my_stage_n);
    }
  }
  if (MODE == 2) bracket_flush(bl, br, my_stage, my_stage_n);
  for (int o = 16; o > 0; o >>= 1) {
    n_sel += __shfl_xor_sync(0xffffffffu, n_sel, o);
    n_fin += __shfl_xor_sync(0xffffffffu, n_fin, o);
  }
  if (lane_id() == 0) {
    if (n_sel) atomicAdd(&ctr->n_conf, (unsigned long long)n_sel);
    if (n_fin) atomicAdd(&ctr->n_finite, (unsigned long long)n_fin);
  }
  if (HIST) {
    __syncthreads();
    for (int i = threadIdx.x; i < 3 * 2048; i += blockDim.x)
      if (sh[i]) atomicAdd(&a.hist0[i], sh[i]);
  }
}

// ---------------------------------------------------------------------------
// submap-local hash: insert with warp-level key dedup
// ---------------------------------------------------------------------------
// Neighbouring pixels mostly fall into the same voxel, so the lanes of a warp first group equal keys (match.any) and
// only each group's leader touches the table, adding the whole group's count with one atomic.  A warp takes a
// 4 x 8 PATCH of one frame rather than 32 pixels of one image row: a 5 cm voxel covers a few pixels in BOTH image
// directions, so a patch holds 2-3x fewer distinct keys than a row segment (fewer claims, fewer atomics); rows of a
// patch are 8 pixels = 128 contiguous bytes of world points, so the loads stay sector-exact.
constexpr int kPatchRows = 4, kPatchCols = 8;
// unsigned division by a runtime constant as multiply-high + shift (exact for all 32-bit numerators the kernels use:
// checked on the host for the largest one) -- the index arithmetic below runs once per warp-iteration and lane
struct FastDiv {
  uint32_t d, mul, shift;
};
static FastDiv make_fastdiv(uint32_t d) {
  FastDiv f{d, 0u, 0u};
  if (d <= 1u) return f;  // handled as a special case
  uint32_t l = 0;
  while ((1ull << l) < d) ++l;  // ceil(log2 d)
  f.shift = l - 1u;
  f.mul = (uint32_t)((((1ull << 32) * ((1ull << l) - d)) / d) + 1ull);  // round-up method (Granlund-Montgomery, n < 2^32)
  return f;
}
__device__ __forceinline__ uint32_t fdiv_u32(uint32_t n, const FastDiv& f) {
  if (f.d <= 1u) return n;
  const uint32_t t = __umulhi(n, f.mul);
  return (t + ((n - t) >> 1)) >> f.shift;
}

struct PixMap {
  uint32_t n_px, H, W, tiles_x, tiles_per_frame;
  uint32_t n_items;  // warp-iterations: patches (PATCH) or 32-pixel runs (linear)
  uint32_t px_per_frame;
  FastDiv div_tpf, div_tx, div_ppf;
};
static PixMap make_pixmap(int64_t n_px, int frames, int H, int W, bool patch) {
  PixMap pm{};
  pm.n_px = (uint32_t)n_px;
  pm.H = (uint32_t)H;
  pm.W = (uint32_t)W;
  pm.tiles_x = (uint32_t)((W + kPatchCols - 1) / kPatchCols);
  pm.tiles_per_frame = pm.tiles_x * (uint32_t)((H + kPatchRows - 1) / kPatchRows);
  pm.n_items = patch ? (uint32_t)frames * pm.tiles_per_frame : (uint32_t)((n_px + 31) / 32);
  pm.px_per_frame = (uint32_t)(H * W);
  pm.div_tpf = make_fastdiv(pm.tiles_per_frame);
  pm.div_tx = make_fastdiv(pm.tiles_x);
  pm.div_ppf = make_fastdiv(pm.px_per_frame);
  return pm;
}
// pixel of (warp-iteration, lane) and the index of its frame inside this call
template <bool PATCH>
__device__ __forceinline__ bool map_pixel(const PixMap& pm, uint32_t item, int lane, uint32_t& pix, uint32_t& frame) {
  if (PATCH) {
    const uint32_t f = fdiv_u32(item, pm.div_tpf);
    const uint32_t r = item - f * pm.tiles_per_frame;
    const uint32_t ty = fdiv_u32(r, pm.div_tx), tx = r - ty * pm.tiles_x;
    const uint32_t y = ty * kPatchRows + ((uint32_t)lane >> 3), x = tx * kPatchCols + ((uint32_t)lane & 7u);
    pix = (f * pm.H + y) * pm.W + x;
    frame = f;
    return y < pm.H && x < pm.W;
  }
  pix = item * 32u + (uint32_t)lane;
  frame = fdiv_u32(pix, pm.div_ppf);
  return pix < pm.n_px;
}

constexpr uint32_t kMaxLocalProbes = 4096u;
constexpr int kOptMaskProbe = 1;  // read the frame-mask word before the atomic OR (most ORs would set a bit that is set)
constexpr int kOptDropRange = 2;  // drop points whose finite voxel coordinate cannot be packed instead of failing the call

// finds or claims the slot of `key` and adds `add` to its count; is_new: this call claimed it (the caller gives it
// a place in the claim list)
__device__ __forceinline__ int table_claim(const LocalTable& t, unsigned long long key, uint32_t add, bool& is_new,
                                           FuseCounters* ctr) {
  uint32_t h = (uint32_t)mix64(key) & t.cap_mask;
  const uint32_t max_probes = min(t.cap_mask, kMaxLocalProbes);  // a table that fills up is reported, not searched end to end
  for (uint32_t probes = 0; probes <= max_probes; ++probes) {
    Slot* sl = t.slots + h;
    unsigned long long cur = sl->key;
    if (cur == kEmptyKey) {
      cur = atomicCAS(&sl->key, kEmptyKey, key);
      if (cur == kEmptyKey) {
        is_new = true;
        cur = key;
      }
    }
    if (cur == key) {
      atomicAdd(&sl->count, add);
      return (int)h;
    }
    h = (h + 1u) & t.cap_mask;
  }
  atomicAdd(&ctr->internal_err, 1u);
  return -1;
}

// all 32 lanes must call; inactive lanes pass active=false.  frame < 0: no frame mask.
__device__ __forceinline__ int warp_insert(const LocalTable& t, bool active, unsigned long long key, int frame, int opts,
                                           FuseCounters* ctr) {
  const int lane = lane_id();
  const unsigned long long k = active ? key : kEmptyKey;
  const unsigned grp = __match_any_sync(0xffffffffu, k);
  const int leader = __ffs(grp) - 1;
  int slot = -1;
  bool is_new = false;
  if (active && lane == leader) slot = table_claim(t, key, (uint32_t)__popc(grp), is_new, ctr);
  // the slots this warp claimed take consecutive places in the claim list: ONE atomic per warp on the shared counter
  // (every new voxel used to add 1 to the same address -- ~350 k serialised atomics per submap at 2 cm voxels)
  const unsigned newm = __ballot_sync(0xffffffffu, is_new);
  if (newm) {
    const int nl = __ffs(newm) - 1;
    uint32_t base = 0;
    if (lane == nl) {
      base = atomicAdd(t.n_occ, (uint32_t)__popc(newm));
      if (base + (uint32_t)__popc(newm) > t.limit) atomicExch(t.overflow, 1u);
    }
    base = __shfl_sync(0xffffffffu, base, nl);
    if (is_new) t.slot_list[base + (uint32_t)__popc(newm & ((1u << lane) - 1u))] = (uint32_t)slot;
  }
  slot = __shfl_sync(0xffffffffu, slot, leader);
  if (frame >= 0) {
    const int lf = __shfl_sync(0xffffffffu, frame, leader);
    if (active && slot >= 0 && (lane == leader || frame != lf)) {
      unsigned long long* mp = &t.slots[slot].mask[frame >> 6];
      const unsigned long long bit = 1ull << (frame & 63);
      if (!(opts & kOptMaskProbe) || (__ldcg(mp) & bit) == 0ull) atomicOr(mp, bit);
    }
  }
  return active ? slot : -1;
}

struct FilterArgs {
  const float4* pw;
  int32_t* pt_slot;
  uint32_t frame_base;
  float cell;  // coarse cell (bbox_coarse) or voxel size (fine)
  uint32_t min_pts;
  int opts;
};

// a finite coordinate that cannot be packed: fail the call (default) or drop the point and count it
__device__ __forceinline__ bool range_problem(bool rerr, int opts, FuseCounters* ctr) {
  if (!rerr) return false;
  atomicAdd((opts & kOptDropRange) ? &ctr->range_dropped : &ctr->range_err, 1u);
  return (opts & kOptDropRange) != 0;
}

// filter 2 (inclusive percentile box, map.py:257-263) + count per coarse cell (map.py:271-275)
template <bool PATCH>
__global__ void __launch_bounds__(256) bbox_coarse_kernel(FilterArgs a, PixMap pm, LocalTable ta, FuseCounters* ctr) {
  const float lx = ctr->bounds[0], hx = ctr->bounds[1], ly = ctr->bounds[2], hy = ctr->bounds[3], lz = ctr->bounds[4],
              hz = ctr->bounds[5];
  unsigned n_in = 0;
  const int lane = lane_id();
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
  for (uint32_t item = warp; item < pm.n_items; item += n_warps) {
    uint32_t pix, frame_unused;
    const bool inside = map_pixel<PATCH>(pm, item, lane, pix, frame_unused);
    bool act = false;
    unsigned long long key = kEmptyKey;
    if (inside) {
      const float4 p = a.pw[pix];
      const uint32_t f = __float_as_uint(p.w);
      if ((f & (PF_SEL | PF_FINITE)) == (PF_SEL | PF_FINITE)) {
        act = (p.x >= lx) && (p.x <= hx) && (p.y >= ly) && (p.y <= hy) && (p.z >= lz) && (p.z <= hz);
        if (act) {
          bool rerr = false;
          key = pack_key(p.x, p.y, p.z, a.cell, rerr);
          if (range_problem(rerr, a.opts, ctr)) act = false;
        }
      }
    }
    const int slot = warp_insert(ta, act, key, -1, a.opts, ctr);
    if (inside) a.pt_slot[pix] = slot;
    n_in += act ? 1u : 0u;
  }
  for (int o = 16; o > 0; o >>= 1) n_in += __shfl_xor_sync(0xffffffffu, n_in, o);
  if (lane == 0 && n_in) atomicAdd(&ctr->n_bbox, (unsigned long long)n_in);
}

// filter 3 (cells with >= min_pts points, map.py:276-280) + fine voxel keys (map.py:351 / submap.py:282)
template <bool FILTERS, bool PATCH>
__global__ void __launch_bounds__(256) fine_insert_kernel(FilterArgs a, PixMap pm, LocalTable ta, LocalTable tb,
                                                          FuseCounters* ctr) {
  unsigned n_in = 0;
  const int lane = lane_id();
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
  for (uint32_t item = warp; item < pm.n_items; item += n_warps) {
    uint32_t pix, fidx;
    const bool inside = map_pixel<PATCH>(pm, item, lane, pix, fidx);
    bool act = false;
    unsigned long long key = kEmptyKey;
    int frame = 0;
    if (inside) {
      const float4 p = a.pw[pix];
      if (FILTERS) {
        const int sa = a.pt_slot[pix];
        act = sa >= 0 && ta.slots[sa].count >= a.min_pts;
      } else {
        act = (__float_as_uint(p.w) & PF_SEL) != 0u;
      }
      if (act) {
        bool rerr = false;
        key = pack_key(p.x, p.y, p.z, a.cell, rerr);
        if (range_problem(rerr, a.opts, ctr)) act = false;
        frame = (int)(a.frame_base + fidx);
      }
    }
    const int slot = warp_insert(tb, act, key, frame, a.opts, ctr);
    if (inside) a.pt_slot[pix] = slot;
    n_in += act ? 1u : 0u;
  }
  for (int o = 16; o > 0; o >>= 1) n_in += __shfl_xor_sync(0xffffffffu, n_in, o);
  if (lane == 0 && n_in) atomicAdd(&ctr->n_fused, (unsigned long long)n_in);
}

// How many of this call's distinct voxels are not in the global map yet (read-only probe): the exact number the
// capacity check needs, so that a call is only stopped when the map really has to grow.  The LAST block to finish
// decides, on the device, whether the call may go on.  It may not if an error was flagged or if the map / contributor
// log cannot take this call's voxels: nothing global has been touched yet, so the host can grow the map and simply
// repeat the call.
__global__ void __launch_bounds__(256) count_new_decide_kernel(LocalTable tb, GlobalStore g, FuseCounters* ctr,
                                                               uint32_t* __restrict__ map_state /* [0] voxels, [1] log */,
                                                               uint32_t vcap, uint32_t log_cap, uint32_t entry_cap) {
  const uint32_t n_occ = ctr->n_occ_b;
  unsigned n_new = 0;
  for (uint32_t lid = blockIdx.x * blockDim.x + threadIdx.x; lid < n_occ; lid += gridDim.x * blockDim.x) {
    const unsigned long long key = tb.slots[tb.slot_list[lid]].key;
    uint64_t h = mix64(key) & g.gmask;
    bool found = false;
    for (uint64_t probes = 0; probes <= g.gmask; ++probes) {
      const unsigned long long cur = g.gkeys[h];
      if (cur == key) {
        found = true;
        break;
      }
      if (cur == kEmptyKey) break;
      h = (h + 1) & g.gmask;
    }
    n_new += found ? 0u : 1u;
  }
  for (int o = 16; o > 0; o >>= 1) n_new += __shfl_xor_sync(0xffffffffu, n_new, o);
  if (lane_id() == 0 && n_new) atomicAdd(&ctr->n_new, n_new);
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0 && atomicAdd(&ctr->ticket, 1u) == gridDim.x - 1) {
    __threadfence();
    const uint32_t total_new = atomicAdd(&ctr->n_new, 0u);
    const uint32_t log_n = map_state[1];
    const bool fits = ((unsigned long long)map_state[0] + total_new <= vcap) &&
                      ((unsigned long long)log_n + n_occ <= log_cap) &&
                      (ctr->n_finite <= (unsigned long long)entry_cap || ctr->n_fused <= (unsigned long long)entry_cap);
    if (!fits || ctr->range_err || ctr->internal_err || ctr->sel_miss || ctr->bad_index || ctr->tbl_overflow) {
      ctr->abort = 1u;
    } else {
      ctr->log_base = log_n;  // calls on one stream run one after the other: plain read-modify-write
      ctr->vox_base = map_state[0];
      map_state[1] = log_n + n_occ;
    }
  }
}

__device__ __forceinline__ void slot_reset(Slot* sl) {
  uint4* q = reinterpret_cast<uint4*>(sl);
  q[0] = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0u, 0u);  // key = kEmptyKey, count = 0, lid = 0
  q[1] = make_uint4(0u, 0u, 0u, 0u);                    // frame mask
}

// leaves both submap-local tables clean for the next call (only the touched slots are visited)
__global__ void __launch_bounds__(256) tables_cleanup_kernel(LocalTable ta, int has_ta, LocalTable tb) {
  const uint32_t nb = *tb.n_occ;
  for (uint32_t lid = blockIdx.x * blockDim.x + threadIdx.x; lid < nb; lid += gridDim.x * blockDim.x)
    slot_reset(tb.slots + tb.slot_list[lid]);
  if (has_ta) {
    const uint32_t na = *ta.n_occ;
    for (uint32_t lid = blockIdx.x * blockDim.x + threadIdx.x; lid < na; lid += gridDim.x * blockDim.x)
      slot_reset(ta.slots + ta.slot_list[lid]);
  }
}

__global__ void __launch_bounds__(256) table_init_kernel(Slot* slots, uint64_t n) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
    slot_reset(slots + i);
}

// ---------------------------------------------------------------------------
// global hash
// ---------------------------------------------------------------------------
__global__ void rehash_kernel(GlobalStore g, uint32_t n) {
  for (uint32_t id = blockIdx.x * blockDim.x + threadIdx.x; id < n; id += gridDim.x * blockDim.x) {
    const unsigned long long key = g.vkey[id];
    uint64_t h = mix64(key) & g.gmask;
    while (true) {
      const unsigned long long prev = atomicCAS(&g.gkeys[h], kEmptyKey, key);
      if (prev == kEmptyKey) {
        g.gids[h] = (int32_t)id;
        break;
      }
      h = (h + 1) & g.gmask;
    }
  }
}

// One thread per DISTINCT voxel of this call: dense local id (= place in the claim list), count, its segment of the
// sorted list (warp scan + one atomic per warp), then the global insert, the map's point count and the contributor
// log entry.  New voxels take consecutive ids per warp (one atomic on the map's voxel counter per warp).
__global__ void __launch_bounds__(256) compact_merge_kernel(LocalTable tb, GlobalStore g, FuseCounters* ctr,
                                                            uint32_t* __restrict__ lv_off, uint32_t* __restrict__ lv_cursor,
                                                            int32_t* __restrict__ lv_gid, int32_t* __restrict__ log_gid,
                                                            int32_t* __restrict__ log_sub,
                                                            unsigned long long* __restrict__ log_mask, int32_t submap_id) {
  if (ctr->abort) return;
  const uint32_t n_occ = ctr->n_occ_b;
  const uint32_t n_round = (n_occ + 31u) & ~31u;
  const size_t log_base = ctr->log_base;
  const int lane = lane_id();
  for (uint32_t lid = blockIdx.x * blockDim.x + threadIdx.x; lid < n_round; lid += gridDim.x * blockDim.x) {
    const bool on = lid < n_occ;
    uint32_t cnt = 0, slot = 0;
    unsigned long long key = kEmptyKey, m0 = 0ull, m1 = 0ull;
    if (on) {
      slot = tb.slot_list[lid];
      const Slot* sl = tb.slots + slot;
      key = sl->key;
      cnt = sl->count;
      m0 = sl->mask[0];
      m1 = sl->mask[1];
    }
    uint32_t incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t up = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += up;
    }
    uint32_t base = 0;
    if (lane == 31) base = atomicAdd(&ctr->seg_total, incl);
    base = __shfl_sync(0xffffffffu, base, 31);
    // global map: find, or claim the hash slot; claimed slots get their ids below
    int gid = -1;
    uint64_t hslot = 0;
    bool claimed = false;
    if (on) {
      uint64_t h = mix64(key) & g.gmask;
      bool done = false;
      for (uint64_t probes = 0; probes <= g.gmask && !done; ++probes) {
        unsigned long long cur = g.gkeys[h];
        if (cur == kEmptyKey) {
          cur = atomicCAS(&g.gkeys[h], kEmptyKey, key);
          if (cur == kEmptyKey) {
            claimed = true;
            hslot = h;
            done = true;
            break;
          }
        }
        if (cur == key) {
          // published by an earlier call (or, if two streams raced, by a claimer that is about to publish)
          while ((gid = reinterpret_cast<volatile int32_t*>(g.gids)[h]) == -1) {
          }
          done = true;
          break;
        }
        h = (h + 1) & g.gmask;
      }
      if (!done) atomicAdd(&ctr->internal_err, 1u);
    }
    const unsigned newm = __ballot_sync(0xffffffffu, claimed);
    if (newm) {
      const int nl = __ffs(newm) - 1;
      uint32_t id0 = 0;
      if (lane == nl) id0 = atomicAdd(g.n_vox, (uint32_t)__popc(newm));
      id0 = __shfl_sync(0xffffffffu, id0, nl);
      if (claimed) {
        const uint32_t id = id0 + (uint32_t)__popc(newm & ((1u << lane) - 1u));
        if (id >= g.vcap) {
          atomicAdd(&ctr->internal_err, 1u);
          reinterpret_cast<volatile int32_t*>(g.gids)[hslot] = -2;  // release waiters; callers treat < 0 as failure
        } else {
          g.vkey[id] = key;
          __threadfence();
          reinterpret_cast<volatile int32_t*>(g.gids)[hslot] = (int32_t)id;
          gid = (int)id;
        }
      }
    }
    if (on) {
      lv_off[lid] = base + incl - cnt;
      lv_cursor[lid] = 0u;
      tb.slots[slot].lid = lid;
      lv_gid[lid] = gid;
      if (gid >= 0) atomicAdd(&g.vcount[gid], cnt);
      log_gid[log_base + lid] = gid;
      log_sub[log_base + lid] = submap_id;
      log_mask[2 * (log_base + lid)] = m0;
      log_mask[2 * (log_base + lid) + 1] = m1;
    }
  }
}

// A sorted-list entry packs (voxel id << 32 | pixel).  Ids: >= 0 fused point; -2 = point that passed the
// confidence and finite tests but was dropped by the bbox / coarse filters: its embedding row is only checked
// for non-finite values (the reference removes such rows BEFORE the percentiles, map.py:247-258, so the
// optimistic pass must notice every one of them).
__device__ __forceinline__ unsigned long long make_entry(int gid, uint32_t pix) {
  return ((unsigned long long)(uint32_t)gid << 32) | (unsigned long long)pix;
}

// counting sort of the fused points by local voxel.  Lanes with the same voxel claim a block of consecutive
// positions with one atomic; check-only pixels are appended behind.
template <bool PATCH>
__global__ void __launch_bounds__(256) scatter_kernel(const int32_t* __restrict__ pt_slot, const float4* __restrict__ pw,
                                                      PixMap pm, LocalTable tb, const uint32_t* __restrict__ lv_off,
                                                      uint32_t* __restrict__ lv_cursor,
                                                      const int32_t* __restrict__ lv_gid, int mark_checks,
                                                      unsigned long long* __restrict__ entries,
                                                      int32_t* __restrict__ point_gid, FuseCounters* ctr) {
  if (ctr->abort) return;
  const uint32_t n_fused = (uint32_t)ctr->n_fused;
  const int lane = lane_id();
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
  for (uint32_t item = warp; item < pm.n_items; item += n_warps) {
    uint32_t pix, frame_unused;
    const bool inside = map_pixel<PATCH>(pm, item, lane, pix, frame_unused);
    const int slot = inside ? pt_slot[pix] : -1;
    const bool act = slot >= 0;
    const uint32_t lid = act ? tb.slots[slot].lid : 0xFFFFFFFFu;
    const unsigned grp = __match_any_sync(0xffffffffu, lid);
    const int leader = __ffs(grp) - 1;
    uint32_t base = 0;
    if (act && lane == leader) base = atomicAdd(&lv_cursor[lid], (uint32_t)__popc(grp));
    base = __shfl_sync(0xffffffffu, base, leader);
    int gid = -1;
    if (act) {
      const uint32_t pos = lv_off[lid] + base + (uint32_t)__popc(grp & ((1u << lane) - 1u));
      gid = lv_gid[lid];
      entries[pos] = make_entry(gid, pix);
    }
    if (mark_checks) {
      bool chk = false;
      if (!act && inside) {
        const uint32_t f = __float_as_uint(pw[pix].w);
        chk = (f & (PF_SEL | PF_FINITE)) == (PF_SEL | PF_FINITE);
      }
      const unsigned cm = __ballot_sync(0xffffffffu, chk);
      if (cm) {
        uint32_t cbase = 0;
        if (lane == __ffs(cm) - 1) cbase = atomicAdd(&ctr->n_check, (uint32_t)__popc(cm));
        cbase = __shfl_sync(0xffffffffu, cbase, __ffs(cm) - 1);
        if (chk) entries[n_fused + cbase + (uint32_t)__popc(cm & ((1u << lane) - 1u))] = make_entry(-2, pix);
      }
    }
    if (point_gid != nullptr && inside) point_gid[pix] = gid;
  }
}

// pixel -> voxel id without the sorted lists (pixel-order / host-streaming path); -2 marks check-only pixels
__global__ void __launch_bounds__(256) point_gid_kernel(const int32_t* __restrict__ pt_slot, const float4* __restrict__ pw,
                                                        uint32_t n_px, LocalTable tb, const int32_t* __restrict__ lv_gid,
                                                        int mark_checks, int32_t* __restrict__ point_gid,
                                                        const FuseCounters* ctr) {
  if (ctr->abort) return;
  for (uint32_t pix = blockIdx.x * blockDim.x + threadIdx.x; pix < n_px; pix += gridDim.x * blockDim.x) {
    const int slot = pt_slot[pix];
    int g = -1;
    if (slot >= 0) {
      g = lv_gid[tb.slots[slot].lid];
    } else if (mark_checks) {
      const uint32_t f = __float_as_uint(pw[pix].w);
      if ((f & (PF_SEL | PF_FINITE)) == (PF_SEL | PF_FINITE)) g = -2;
    }
    point_gid[pix] = g;
  }
}

#include "prep7.cuh"  // one-table preparation: insert7 / insert7d / insert7u, coarse7, keep7, compact_merge7, scatter7

// ---------------------------------------------------------------------------
// embedding accumulate
// ---------------------------------------------------------------------------
// A row is d embedding channels of one pixel (bf16: 2d bytes, f32: 4d bytes), read as 16-byte vectors:
// lane l of a warp owns vectors l, l+32, ... (VPL of them), i.e. a warp reads 512 contiguous bytes per
// vector index.  Each warp walks a chunk of 32 entries; rows are summed in fp32 registers and flushed with
// vector REDs whenever the voxel id changes.  U rows are in flight per warp.
template <bool BF16>
struct RowVec;
// f32 accumulator += one bf16 half of a 32-bit word: Blackwell's mixed-precision add (PTX add.f32.bf16,
// SASS FHADD.BF16) takes the half register directly, so widening costs no instruction of its own.
__device__ __forceinline__ float add_bf16_lo(float acc, uint32_t w) {
  float d;
  asm("{ .reg .b16 lo, hi; mov.b32 {lo, hi}, %2; add.f32.bf16 %0, lo, %1; }" : "=f"(d) : "f"(acc), "r"(w));
  return d;
}
__device__ __forceinline__ float add_bf16_hi(float acc, uint32_t w) {
  float d;
  asm("{ .reg .b16 lo, hi; mov.b32 {lo, hi}, %2; add.f32.bf16 %0, hi, %1; }" : "=f"(d) : "f"(acc), "r"(w));
  return d;
}
template <>
struct RowVec<true> {
  static constexpr int EPV = 8;  // elements per 16-byte vector
  __device__ static __forceinline__ void add(float* acc, const uint4& v) {
    acc[0] = add_bf16_lo(acc[0], v.x);
    acc[1] = add_bf16_hi(acc[1], v.x);
    acc[2] = add_bf16_lo(acc[2], v.y);
    acc[3] = add_bf16_hi(acc[3], v.y);
    acc[4] = add_bf16_lo(acc[4], v.z);
    acc[5] = add_bf16_hi(acc[5], v.z);
    acc[6] = add_bf16_lo(acc[6], v.w);
    acc[7] = add_bf16_hi(acc[7], v.w);
  }
  // acc += scale * v  (run-length path of indexed embeddings)
  __device__ static __forceinline__ void fma(float* acc, const uint4& v, float scale) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      acc[2 * i] = fmaf(__uint_as_float(w[i] << 16), scale, acc[2 * i]);
      acc[2 * i + 1] = fmaf(__uint_as_float(w[i] & 0xFFFF0000u), scale, acc[2 * i + 1]);
    }
  }
  // exponent all ones in either half
  __device__ static __forceinline__ bool nonfinite(const uint4& v) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    uint32_t bad = 0u;
#pragma unroll
    for (int i = 0; i < 4; ++i) bad |= ((w[i] & 0x7F807F80u) + 0x00800080u) & 0x80008000u;
    return bad != 0u;
  }
};
template <>
struct RowVec<false> {
  static constexpr int EPV = 4;
  __device__ static __forceinline__ void add(float* acc, const uint4& v) {
    acc[0] += __uint_as_float(v.x);
    acc[1] += __uint_as_float(v.y);
    acc[2] += __uint_as_float(v.z);
    acc[3] += __uint_as_float(v.w);
  }
  __device__ static __forceinline__ void fma(float* acc, const uint4& v, float scale) {
    acc[0] = fmaf(__uint_as_float(v.x), scale, acc[0]);
    acc[1] = fmaf(__uint_as_float(v.y), scale, acc[1]);
    acc[2] = fmaf(__uint_as_float(v.z), scale, acc[2]);
    acc[3] = fmaf(__uint_as_float(v.w), scale, acc[3]);
  }
  __device__ static __forceinline__ bool nonfinite(const uint4& v) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    bool bad = false;
#pragma unroll
    for (int i = 0; i < 4; ++i) bad |= (w[i] & 0x7F800000u) == 0x7F800000u;
    return bad;
  }
};

struct AccArgs {
  const int32_t* emb_index;    // indexed embeddings: pixel p reads row emb_index[p] of the table `emb` (else nullptr)
  const uint8_t* emb;          // row of pixel p starts at emb + (p - pix_base) * row_bytes
  int64_t pix_base;
  int64_t row_bytes;
  const unsigned long long* entries;  // SORTED mode: packed (voxel id << 32 | pixel), n_fused + n_check of them
  const int32_t* point_gid;    // PIXEL mode: voxel id per pixel (-1: skip); entries are pixels [pix_base, pix_base+n)
  int64_t n;                   // PIXEL mode: number of pixels
  float* vsum;
  int d;
  int nvec;                    // 16-byte vectors per row
  FuseCounters* ctr;
};

// CHECK (optimistic non-finite detection, filters on): a non-finite embedding value makes the fp32 sum of its
// voxel segment non-finite, so the accumulators are tested once per flush instead of testing every row; entries
// that are only to be checked (id -2: behind the fused entries in the sorted list, anywhere in pixel mode) are
// tested directly.  Any hit is counted in ctr->n_bad_emb and makes the fuse call fail with VSM_E_NONFINITE_EMB
// (the caller redoes the build with the exact pre-check).
// FULL: every lane owns VPL vectors (nvec == 32*VPL, e.g. d=512): no per-lane predicates in the hot loop.
// RLE (indexed embeddings, SORTED && FULL): neighbouring entries of a voxel mostly carry the same table row (pixels of
// one mask), so runs of equal (voxel, row) are found with one ballot per chunk and added as  length x row  -- one row
// read per run instead of one per point.
template <bool BF16, int VPL, bool SORTED, bool CHECK, bool FULL, bool RLE = false>
__global__ void __launch_bounds__(256, (VPL <= 2) ? 3 : 2) accumulate_kernel(AccArgs a) {
  constexpr int EPV = RowVec<BF16>::EPV;
  constexpr int U = (VPL <= 2) ? 4 : 2;  // rows in flight per warp
  if (a.ctr->abort) return;
  const int lane = lane_id();
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int64_t n_entries = SORTED ? (int64_t)a.ctr->n_fused : a.n;
  const int64_t n_chunks = (n_entries + 31) >> 5;
  const uint32_t rb = (uint32_t)a.row_bytes;
  const int nvec = a.nvec;
  const uint8_t* emb0 = a.emb - a.pix_base * a.row_bytes + (size_t)lane * 16;
  float* const vsum0 = a.vsum + (size_t)lane * EPV;
  const uint32_t d = (uint32_t)a.d;
  unsigned n_bad = 0;

  for (int64_t chunk = warp; chunk < n_chunks; chunk += n_warps) {
    const int64_t base = chunk << 5;
    const int cnt = (int)min((int64_t)32, n_entries - base);
    uint32_t my_pix = 0;
    int my_gid = -1;
    if (lane < cnt) {
      if (SORTED) {
        const unsigned long long e = a.entries[base + lane];
        my_pix = (uint32_t)e;
        my_gid = (int)(uint32_t)(e >> 32);
      } else {
        my_pix = (uint32_t)(a.pix_base + base + lane);
        my_gid = a.point_gid[my_pix];
      }
    }
    // pixel mode: -1 = not selected; -2 = selected but filtered out (its row is only checked for non-finite values)
    if (!SORTED && __ballot_sync(0xffffffffu, CHECK ? (my_gid != -1) : (my_gid >= 0)) == 0u) continue;
    // indexed embeddings: one gather per entry here, outside the row loop
    if (a.emb_index != nullptr && lane < cnt && my_gid != -1) my_pix = (uint32_t)a.emb_index[my_pix];

    float acc[VPL * EPV];
#pragma unroll
    for (int i = 0; i < VPL * EPV; ++i) acc[i] = 0.f;
    int cur = -1;

    auto flush = [&]() {
      if (cur >= 0) {
        bool ok = true;
        if (CHECK) {
          float t = 0.f;
#pragma unroll
          for (int i = 0; i < VPL * EPV; ++i) t = fmaf(acc[i], 0.f, t);  // NaN iff some accumulator is Inf/NaN
          ok = !__any_sync(0xffffffffu, t != t);
          if (!ok && lane == 0) ++n_bad;
        }
        if (ok) {
          float* dst = vsum0 + (size_t)cur * d;
#pragma unroll
          for (int v = 0; v < VPL; ++v) {
            if (FULL || lane + 32 * v < nvec) {
#pragma unroll
              for (int q = 0; q < EPV / 4; ++q)
                red_add_v4(dst + v * 32 * EPV + 4 * q, acc[v * EPV + 4 * q], acc[v * EPV + 4 * q + 1],
                           acc[v * EPV + 4 * q + 2], acc[v * EPV + 4 * q + 3]);
            }
          }
        }
      }
#pragma unroll
      for (int i = 0; i < VPL * EPV; ++i) acc[i] = 0.f;
    };

    int j = 0;
    if (RLE) {
      const int pg = __shfl_up_sync(0xffffffffu, my_gid, 1);
      const uint32_t pr = __shfl_up_sync(0xffffffffu, my_pix, 1);
      unsigned starts = __ballot_sync(0xffffffffu, lane < cnt && (lane == 0 || my_gid != pg || my_pix != pr));
      while (starts) {
        // up to U runs in flight
        uint4 rows[U][VPL];
        int gids[U];
        float len[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          gids[u] = -1;
          len[u] = 0.f;
          if (starts) {
            const int sidx = __ffs(starts) - 1;
            starts &= starts - 1;
            const int eidx = starts ? __ffs(starts) - 1 : cnt;
            len[u] = (float)(eidx - sidx);
            gids[u] = __shfl_sync(0xffffffffu, my_gid, sidx);
            const uint32_t rj = __shfl_sync(0xffffffffu, my_pix, sidx);
            const uint8_t* row = emb0 + (unsigned long long)rj * rb;
#pragma unroll
            for (int v = 0; v < VPL; ++v) rows[u][v] = ld_stream_v4(row + v * 512);
          }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          if (gids[u] < 0) continue;  // warp-uniform
          if (gids[u] != cur) {
            flush();
            cur = gids[u];
          }
#pragma unroll
          for (int v = 0; v < VPL; ++v) RowVec<BF16>::fma(acc + v * EPV, rows[u][v], len[u]);
        }
      }
      j = cnt;
    } else if (SORTED && FULL) {
      // hot loop: U rows in flight, every entry is a fused point, every lane owns VPL vectors
      for (; j + U <= cnt; j += U) {
        uint4 rows[U][VPL];
        int gids[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const uint32_t pj = __shfl_sync(0xffffffffu, my_pix, j + u);
          gids[u] = __shfl_sync(0xffffffffu, my_gid, j + u);
          const uint8_t* row = emb0 + (unsigned long long)pj * rb;
#pragma unroll
          for (int v = 0; v < VPL; ++v) rows[u][v] = ld_stream_v4(row + v * 512);
        }
        bool same = true;
#pragma unroll
        for (int u = 0; u < U; ++u) same &= (gids[u] == cur);
        if (same) {
          // common case (tens of points per voxel): straight-line adds, no boundary inside the batch
#pragma unroll
          for (int u = 0; u < U; ++u)
#pragma unroll
            for (int v = 0; v < VPL; ++v) RowVec<BF16>::add(acc + v * EPV, rows[u][v]);
        } else {
#pragma unroll
          for (int u = 0; u < U; ++u) {
            if (gids[u] != cur) {
              flush();
              cur = gids[u];
            }
#pragma unroll
            for (int v = 0; v < VPL; ++v) RowVec<BF16>::add(acc + v * EPV, rows[u][v]);
          }
        }
      }
    }
    // generic loop: tails, partial rows (small d), pixel mode with its skipped / check-only pixels
    for (; j < cnt; j += U) {
      uint4 rows[U][VPL];
      int gids[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int src = (j + u) & 31;
        const uint32_t pj = __shfl_sync(0xffffffffu, my_pix, src);
        int gj = __shfl_sync(0xffffffffu, my_gid, src);
        if (j + u >= cnt) gj = -1;
        gids[u] = gj;
        const bool want_row = CHECK ? (gj != -1) : (gj >= 0);
        const uint8_t* row = emb0 + (unsigned long long)pj * rb;
#pragma unroll
        for (int v = 0; v < VPL; ++v)
          rows[u][v] = (want_row && (FULL || lane + 32 * v < nvec)) ? ld_stream_v4(row + v * 512)
                                                                   : make_uint4(0u, 0u, 0u, 0u);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int gj = gids[u];
        if (gj == -1) continue;  // warp-uniform
        if (CHECK && gj == -2) {
          bool bad = false;
#pragma unroll
          for (int v = 0; v < VPL; ++v) bad |= RowVec<BF16>::nonfinite(rows[u][v]);
          if (__any_sync(0xffffffffu, bad) && lane == 0) ++n_bad;
          continue;
        }
        if (gj < 0) continue;
        if (gj != cur) {
          flush();
          cur = gj;
        }
#pragma unroll
        for (int v = 0; v < VPL; ++v) RowVec<BF16>::add(acc + v * EPV, rows[u][v]);
      }
    }
    flush();
  }

  if (SORTED && CHECK) {
    // check-only entries sit behind the fused ones: one warp per row, test the raw values
    const int64_t n_check = (int64_t)a.ctr->n_check;
    for (int64_t i = warp; i < n_check; i += n_warps) {
      uint32_t pj = (uint32_t)a.entries[n_entries + i];
      if (a.emb_index != nullptr) pj = (uint32_t)a.emb_index[pj];
      const uint8_t* row = emb0 + (unsigned long long)pj * rb;
      bool bad = false;
#pragma unroll
      for (int v = 0; v < VPL; ++v)
        if (FULL || lane + 32 * v < nvec) bad |= RowVec<BF16>::nonfinite(ld_stream_v4(row + v * 512));
      if (__any_sync(0xffffffffu, bad) && lane == 0) ++n_bad;
    }
  }
  if (CHECK && n_bad) atomicAdd(&a.ctr->n_bad_emb, (unsigned long long)n_bad);
}

// ---- the same accumulate with the rows staged through shared memory (cp.async) --------------------------------
// accumulate_kernel keeps U = 4 rows of a warp in REGISTERS while they are on their way from HBM: 96 KB in flight per
// SM at three CTAs, which needs every SM of the device to cover HBM's latency x bandwidth.  Here every lane copies its
// 16-byte pieces of the next DEPTH rows straight into a ring in shared memory (cp.async.cg, no register in between) and
// adds row r while rows r+1 .. r+DEPTH are in flight; a lane only ever reads the bytes it copied itself, so the ring
// needs no barrier, only cp.async.wait_group.  With ~190 KB in flight per SM the kernel saturates HBM on a PART of the
// SMs -- which is what lets the next call's preparation kernels run beside it on the rest (SM partitions, above).
// Voxel-sorted lists with full rows only (d a multiple of 32 vectors); check-only entries as in accumulate_kernel.
__device__ __forceinline__ void cp_async_16(uint32_t smem_addr, const void* gptr) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr), "l"(gptr) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ uint4 lds_v4(uint32_t smem_addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(smem_addr));
  return v;
}

template <bool BF16, int VPL, bool CHECK, int DEPTH>
__global__ void __launch_bounds__(256, 2) accumulate_ring_kernel(AccArgs a) {
  constexpr int EPV = RowVec<BF16>::EPV;
  constexpr uint32_t ROW = (uint32_t)VPL * 512u;  // bytes per row
  extern __shared__ __align__(16) uint8_t ring_smem[];
  if (a.ctr->abort) return;
  const int lane = lane_id();
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int64_t n_entries = (int64_t)a.ctr->n_fused;
  const int64_t n_chunks = (n_entries + 31) >> 5;
  const uint32_t rb = (uint32_t)a.row_bytes;
  const uint8_t* emb0 = a.emb - a.pix_base * a.row_bytes + (size_t)lane * 16;
  float* const vsum0 = a.vsum + (size_t)lane * EPV;
  const uint32_t d = (uint32_t)a.d;
  // this lane's 16-byte column of the warp's ring
  const uint32_t ring0 = (uint32_t)__cvta_generic_to_shared(ring_smem) + (threadIdx.x >> 5) * (uint32_t)DEPTH * ROW + (uint32_t)lane * 16u;
  unsigned n_bad = 0;

  float acc[VPL * EPV];
#pragma unroll
  for (int i = 0; i < VPL * EPV; ++i) acc[i] = 0.f;
  int cur = -1;
  auto flush = [&]() {
    if (cur >= 0) {
      bool ok = true;
      if (CHECK) {
        float t = 0.f;
#pragma unroll
        for (int i = 0; i < VPL * EPV; ++i) t = fmaf(acc[i], 0.f, t);  // NaN iff some accumulator is Inf/NaN
        ok = !__any_sync(0xffffffffu, t != t);
        if (!ok && lane == 0) ++n_bad;
      }
      if (ok) {
        float* dst = vsum0 + (size_t)cur * d;
#pragma unroll
        for (int v = 0; v < VPL; ++v)
#pragma unroll
          for (int q = 0; q < EPV / 4; ++q)
            red_add_v4(dst + v * 32 * EPV + 4 * q, acc[v * EPV + 4 * q], acc[v * EPV + 4 * q + 1], acc[v * EPV + 4 * q + 2],
                       acc[v * EPV + 4 * q + 3]);
      }
    }
#pragma unroll
    for (int i = 0; i < VPL * EPV; ++i) acc[i] = 0.f;
  };

  // entries of the chunk being added (C) and of the next one (N): the copies run up to DEPTH <= 32 rows ahead
  auto load_entries = [&](int64_t chunk, uint32_t& pix, int& gid) {
    pix = 0u;
    gid = -1;
    if (chunk < n_chunks) {
      const int64_t i = (chunk << 5) + lane;
      if (i < n_entries) {
        const unsigned long long e = a.entries[i];
        pix = (uint32_t)e;
        gid = (int)(uint32_t)(e >> 32);
        if (a.emb_index != nullptr) pix = (uint32_t)a.emb_index[pix];
      }
    }
  };
  uint32_t pixC, pixN;
  int gidC, gidN;
  load_entries(warp, pixC, gidC);
  load_entries(warp + n_warps, pixN, gidN);
  auto issue = [&](uint32_t pix, int gid, int src_lane, uint32_t slot) {
    const uint32_t pj = __shfl_sync(0xffffffffu, pix, src_lane);
    const int gj = __shfl_sync(0xffffffffu, gid, src_lane);
    if (gj >= 0) {
      const uint8_t* row = emb0 + (unsigned long long)pj * rb;
#pragma unroll
      for (int v = 0; v < VPL; ++v) cp_async_16(ring0 + slot * ROW + (uint32_t)v * 512u, row + v * 512);
    }
    cp_async_commit();  // one group per row, empty or not: wait_group counts rows
  };
  // prologue: the first DEPTH rows of the first chunk
#pragma unroll
  for (int r = 0; r < DEPTH; ++r) issue(pixC, gidC, r, (uint32_t)r);
  uint32_t slot = 0;
  for (int64_t chunk = warp; chunk < n_chunks; chunk += n_warps) {
#pragma unroll 4
    for (int j = 0; j < 32; ++j) {
      cp_async_wait<DEPTH - 1>();  // row (chunk, j) has landed
      const int gj = __shfl_sync(0xffffffffu, gidC, j);
      uint4 rows[VPL];
      if (gj >= 0) {
#pragma unroll
        for (int v = 0; v < VPL; ++v) rows[v] = lds_v4(ring0 + slot * ROW + (uint32_t)v * 512u);
      }
      // the slot is free again: row j + DEPTH goes into it
      const int jj = j + DEPTH;
      if (jj < 32)
        issue(pixC, gidC, jj, slot);
      else
        issue(pixN, gidN, jj - 32, slot);
      slot = slot + 1u == (uint32_t)DEPTH ? 0u : slot + 1u;
      if (gj >= 0) {
        if (gj != cur) {
          flush();
          cur = gj;
        }
#pragma unroll
        for (int v = 0; v < VPL; ++v) RowVec<BF16>::add(acc + v * EPV, rows[v]);
      }
    }
    pixC = pixN;
    gidC = gidN;
    load_entries(chunk + 2 * n_warps, pixN, gidN);
  }
  cp_async_wait<0>();
  flush();

  if (CHECK) {
    // check-only entries sit behind the fused ones: one warp per row, test the raw values
    const int64_t n_check = (int64_t)a.ctr->n_check;
    for (int64_t i = warp; i < n_check; i += n_warps) {
      uint32_t pj = (uint32_t)a.entries[n_entries + i];
      if (a.emb_index != nullptr) pj = (uint32_t)a.emb_index[pj];
      const uint8_t* row = emb0 + (unsigned long long)pj * rb;
      bool bad = false;
#pragma unroll
      for (int v = 0; v < VPL; ++v) bad |= RowVec<BF16>::nonfinite(ld_stream_v4(row + v * 512));
      if (__any_sync(0xffffffffu, bad) && lane == 0) ++n_bad;
    }
    if (n_bad) atomicAdd(&a.ctr->n_bad_emb, (unsigned long long)n_bad);
  }
}

// ---- the same accumulate with the rows fetched by TMA bulk copies ------------------------------------------------------
// scripts/cu/sm_bw_probe*.cu: through the load/store path an SM pulls at most ~57 GB/s from L2/HBM whatever is in
// flight (148 SMs: 6.6 TB/s; 84 SMs: 4.6 TB/s; 60 SMs: 3.4 TB/s) -- which is why accumulate_kernel needs every SM and why
// the cp.async ring changed nothing.  cp.async.bulk does not share that limit: 60 SMs pull 7.2 TB/s, 37 SMs 6.3 TB/s.
// Here lane 0 of every warp issues ONE bulk copy per row (1 KB, global -> the warp's ring in shared memory, completion on
// the slot's mbarrier), DEPTH rows ahead; the warp waits for a slot's barrier, reads its 16-byte pieces, re-issues the
// slot and adds.  A slot is written by the async proxy only after every lane has read it (__syncwarp before the
// re-issue); the mbarrier wait makes the bytes visible to the readers.
__device__ __forceinline__ void mbar_init_cta(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx_cta(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void mbar_wait_cta(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{\n\t.reg .pred q;\n\tmbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2;\n\tselp.u32 %0, 1, 0, q;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  }
}

template <bool BF16, int VPL, bool CHECK, int DEPTH>
__global__ void __launch_bounds__(256, 2) accumulate_tma_kernel(AccArgs a) {
  constexpr int EPV = RowVec<BF16>::EPV;
  constexpr uint32_t ROW = (uint32_t)VPL * 512u;  // bytes per row
  constexpr int U = 4;                            // rows taken out of the ring per step
  static_assert(DEPTH % U == 0 && DEPTH <= 32, "ring depth: a multiple of the step, one phase bit per slot");
  extern __shared__ __align__(128) uint8_t tma_smem[];
  const int lane = lane_id(), wib = (int)(threadIdx.x >> 5);
  const uint32_t ring0 = (uint32_t)__cvta_generic_to_shared(tma_smem) + (uint32_t)wib * (uint32_t)DEPTH * ROW;
  const uint32_t bar0 = (uint32_t)__cvta_generic_to_shared(tma_smem) + 8u * (uint32_t)DEPTH * ROW + (uint32_t)wib * (uint32_t)DEPTH * 8u;
  if (lane < DEPTH) mbar_init_cta(bar0 + (uint32_t)lane * 8u, 1u);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  if (a.ctr->abort) return;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int64_t n_entries = (int64_t)a.ctr->n_fused;
  const int64_t n_chunks = (n_entries + 31) >> 5;
  const uint32_t rb = (uint32_t)a.row_bytes;
  const uint8_t* emb_row0 = a.emb - a.pix_base * a.row_bytes;
  float* const vsum0 = a.vsum + (size_t)lane * EPV;
  const uint32_t d = (uint32_t)a.d;
  unsigned n_bad = 0;

  float acc[VPL * EPV];
#pragma unroll
  for (int i = 0; i < VPL * EPV; ++i) acc[i] = 0.f;
  int cur = -1;
  auto flush = [&]() {
    if (cur >= 0) {
      bool ok = true;
      if (CHECK) {
        float t = 0.f;
#pragma unroll
        for (int i = 0; i < VPL * EPV; ++i) t = fmaf(acc[i], 0.f, t);  // NaN iff some accumulator is Inf/NaN
        ok = !__any_sync(0xffffffffu, t != t);
        if (!ok && lane == 0) ++n_bad;
      }
      if (ok) {
        float* dst = vsum0 + (size_t)cur * d;
#pragma unroll
        for (int v = 0; v < VPL; ++v)
#pragma unroll
          for (int q = 0; q < EPV / 4; ++q)
            red_add_v4(dst + v * 32 * EPV + 4 * q, acc[v * EPV + 4 * q], acc[v * EPV + 4 * q + 1], acc[v * EPV + 4 * q + 2],
                       acc[v * EPV + 4 * q + 3]);
      }
    }
#pragma unroll
    for (int i = 0; i < VPL * EPV; ++i) acc[i] = 0.f;
  };
  auto load_entries = [&](int64_t chunk, uint32_t& pix, int& gid) {
    pix = 0u;
    gid = -1;
    if (chunk < n_chunks) {
      const int64_t i = (chunk << 5) + lane;
      if (i < n_entries) {
        const unsigned long long e = a.entries[i];
        pix = (uint32_t)e;
        gid = (int)(uint32_t)(e >> 32);
      }
    }
  };
  uint32_t pixC, pixN;
  int gidC, gidN;
  load_entries(warp, pixC, gidC);
  load_entries(warp + n_warps, pixN, gidN);
  // one bulk copy per row, issued by lane 0; rows of empty list positions (the last chunk's tail) are not fetched
  auto issue = [&](uint32_t pix, int gid, int src_lane, uint32_t slot) {
    const uint32_t pj = __shfl_sync(0xffffffffu, pix, src_lane);
    const int gj = __shfl_sync(0xffffffffu, gid, src_lane);
    if (gj >= 0 && lane == 0) {
      const uint32_t bar = bar0 + slot * 8u;
      mbar_expect_tx_cta(bar, ROW);
      bulk_g2s(ring0 + slot * ROW, emb_row0 + (unsigned long long)pj * rb, ROW, bar);
    }
  };
#pragma unroll
  for (int r = 0; r < DEPTH; ++r) issue(pixC, gidC, r, (uint32_t)r);
  uint32_t slot = 0, phases = 0u;
  for (int64_t chunk = warp; chunk < n_chunks; chunk += n_warps) {
#pragma unroll 1
    for (int j = 0; j < 32; j += U) {
      int g[U];
      uint4 rows[U][VPL];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        g[u] = __shfl_sync(0xffffffffu, gidC, j + u);
        if (g[u] >= 0) {  // warp-uniform
          const uint32_t su = slot + (uint32_t)u;
          mbar_wait_cta(bar0 + su * 8u, (phases >> su) & 1u);
          phases ^= 1u << su;
#pragma unroll
          for (int v = 0; v < VPL; ++v) rows[u][v] = lds_v4(ring0 + su * ROW + (uint32_t)v * 512u + (uint32_t)lane * 16u);
        }
      }
      __syncwarp();  // every lane has read its pieces: the slots may be overwritten
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int jj = j + u + DEPTH;
        if (jj < 32)
          issue(pixC, gidC, jj, slot + (uint32_t)u);
        else
          issue(pixN, gidN, jj - 32, slot + (uint32_t)u);
      }
      slot = slot + (uint32_t)U == (uint32_t)DEPTH ? 0u : slot + (uint32_t)U;
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (g[u] < 0) continue;
        if (g[u] != cur) {
          flush();
          cur = g[u];
        }
#pragma unroll
        for (int v = 0; v < VPL; ++v) RowVec<BF16>::add(acc + v * EPV, rows[u][v]);
      }
    }
    pixC = pixN;
    gidC = gidN;
    load_entries(chunk + 2 * n_warps, pixN, gidN);
  }
  flush();

  if (CHECK) {
    const uint8_t* emb0 = emb_row0 + (size_t)lane * 16;
    const int64_t n_check = (int64_t)a.ctr->n_check;
    for (int64_t i = warp; i < n_check; i += n_warps) {
      const uint32_t pj = (uint32_t)a.entries[n_entries + i];
      const uint8_t* row = emb0 + (unsigned long long)pj * rb;
      bool bad = false;
#pragma unroll
      for (int v = 0; v < VPL; ++v) bad |= RowVec<BF16>::nonfinite(ld_stream_v4(row + v * 512));
      if (__any_sync(0xffffffffu, bad) && lane == 0) ++n_bad;
    }
    if (n_bad) atomicAdd(&a.ctr->n_bad_emb, (unsigned long long)n_bad);
  }
}

// ---- both paths at once -------------------------------------------------------------------------------------------------
// Measured (scripts/green_ab.py): accumulate_tma_kernel alone is no faster -- a bulk copy costs the SM's TMA unit ~46
// cycles whatever its size, so 1 KB rows give ~43 GB/s per SM, below even the load/store path's ~57 GB/s.  But the two
// limits are different units of the SM: half of a CTA's warps fetch their rows with bulk copies, the other half with
// register loads, and the SM pulls the sum.  The chunks of the sorted list are split between the two kinds of warps in
// the ratio of their speeds (tma_share).
template <bool BF16, int VPL, bool CHECK, int DEPTH, int TMA_WARPS>
__global__ void __launch_bounds__(256, (VPL <= 2) ? 3 : 2) accumulate_mix_kernel(AccArgs a, float tma_share) {
  constexpr int EPV = RowVec<BF16>::EPV;
  constexpr uint32_t ROW = (uint32_t)VPL * 512u;
  constexpr int U = 4;
  static_assert(DEPTH % U == 0 && DEPTH <= 32 && TMA_WARPS >= 1 && TMA_WARPS < 8, "ring depth / warp roles");
  extern __shared__ __align__(128) uint8_t mix_smem[];
  const int lane = lane_id(), wib = (int)(threadIdx.x >> 5);
  const bool tma_warp = wib < TMA_WARPS;
  const uint32_t ring0 = (uint32_t)__cvta_generic_to_shared(mix_smem) + (uint32_t)wib * (uint32_t)DEPTH * ROW;
  const uint32_t bar0 = (uint32_t)__cvta_generic_to_shared(mix_smem) + (uint32_t)TMA_WARPS * (uint32_t)DEPTH * ROW +
                        (uint32_t)wib * (uint32_t)DEPTH * 8u;
  if (tma_warp && lane < DEPTH) mbar_init_cta(bar0 + (uint32_t)lane * 8u, 1u);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  if (a.ctr->abort) return;
  const int64_t n_entries = (int64_t)a.ctr->n_fused;
  const int64_t n_chunks = (n_entries + 31) >> 5;
  const int64_t split = min(n_chunks, (int64_t)((double)n_chunks * (double)tma_share));  // chunks [0, split): bulk copies
  // my index among the warps of my kind, their number, and my kind's range of chunks
  const int64_t kidx = tma_warp ? (int64_t)blockIdx.x * TMA_WARPS + wib : (int64_t)blockIdx.x * (8 - TMA_WARPS) + (wib - TMA_WARPS);
  const int64_t n_kind = (int64_t)gridDim.x * (tma_warp ? TMA_WARPS : 8 - TMA_WARPS);
  const int64_t lo = tma_warp ? 0 : split, hi = tma_warp ? split : n_chunks;
  const uint32_t rb = (uint32_t)a.row_bytes;
  const uint8_t* emb_row0 = a.emb - a.pix_base * a.row_bytes;
  const uint8_t* emb0 = emb_row0 + (size_t)lane * 16;
  float* const vsum0 = a.vsum + (size_t)lane * EPV;
  const uint32_t d = (uint32_t)a.d;
  unsigned n_bad = 0;

  float acc[VPL * EPV];
#pragma unroll
  for (int i = 0; i < VPL * EPV; ++i) acc[i] = 0.f;
  int cur = -1;
  auto flush = [&]() {
    if (cur >= 0) {
      bool ok = true;
      if (CHECK) {
        float t = 0.f;
#pragma unroll
        for (int i = 0; i < VPL * EPV; ++i) t = fmaf(acc[i], 0.f, t);
        ok = !__any_sync(0xffffffffu, t != t);
        if (!ok && lane == 0) ++n_bad;
      }
      if (ok) {
        float* dst = vsum0 + (size_t)cur * d;
#pragma unroll
        for (int v = 0; v < VPL; ++v)
#pragma unroll
          for (int q = 0; q < EPV / 4; ++q)
            red_add_v4(dst + v * 32 * EPV + 4 * q, acc[v * EPV + 4 * q], acc[v * EPV + 4 * q + 1], acc[v * EPV + 4 * q + 2],
                       acc[v * EPV + 4 * q + 3]);
      }
    }
#pragma unroll
    for (int i = 0; i < VPL * EPV; ++i) acc[i] = 0.f;
  };
  auto load_entries = [&](int64_t chunk, uint32_t& pix, int& gid) {
    pix = 0u;
    gid = -1;
    if (chunk < hi) {
      const int64_t i = (chunk << 5) + lane;
      if (i < n_entries) {
        const unsigned long long e = a.entries[i];
        pix = (uint32_t)e;
        gid = (int)(uint32_t)(e >> 32);
      }
    }
  };

  if (tma_warp) {
    uint32_t pixC, pixN;
    int gidC, gidN;
    load_entries(lo + kidx, pixC, gidC);
    load_entries(lo + kidx + n_kind, pixN, gidN);
    auto issue = [&](uint32_t pix, int gid, int src_lane, uint32_t slot) {
      const uint32_t pj = __shfl_sync(0xffffffffu, pix, src_lane);
      const int gj = __shfl_sync(0xffffffffu, gid, src_lane);
      if (gj >= 0 && lane == 0) {
        const uint32_t bar = bar0 + slot * 8u;
        mbar_expect_tx_cta(bar, ROW);
        bulk_g2s(ring0 + slot * ROW, emb_row0 + (unsigned long long)pj * rb, ROW, bar);
      }
    };
#pragma unroll
    for (int r = 0; r < DEPTH; ++r) issue(pixC, gidC, r, (uint32_t)r);
    uint32_t slot = 0, phases = 0u;
    for (int64_t chunk = lo + kidx; chunk < hi; chunk += n_kind) {
#pragma unroll 1
      for (int j = 0; j < 32; j += U) {
        int g[U];
        uint4 rows[U][VPL];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          g[u] = __shfl_sync(0xffffffffu, gidC, j + u);
          if (g[u] >= 0) {
            const uint32_t su = slot + (uint32_t)u;
            mbar_wait_cta(bar0 + su * 8u, (phases >> su) & 1u);
            phases ^= 1u << su;
#pragma unroll
            for (int v = 0; v < VPL; ++v) rows[u][v] = lds_v4(ring0 + su * ROW + (uint32_t)v * 512u + (uint32_t)lane * 16u);
          }
        }
        __syncwarp();
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int jj = j + u + DEPTH;
          if (jj < 32)
            issue(pixC, gidC, jj, slot + (uint32_t)u);
          else
            issue(pixN, gidN, jj - 32, slot + (uint32_t)u);
        }
        slot = slot + (uint32_t)U == (uint32_t)DEPTH ? 0u : slot + (uint32_t)U;
#pragma unroll
        for (int u = 0; u < U; ++u) {
          if (g[u] < 0) continue;
          if (g[u] != cur) {
            flush();
            cur = g[u];
          }
#pragma unroll
          for (int v = 0; v < VPL; ++v) RowVec<BF16>::add(acc + v * EPV, rows[u][v]);
        }
      }
      pixC = pixN;
      gidC = gidN;
      load_entries(chunk + 2 * n_kind, pixN, gidN);
    }
    flush();
  } else {
    // register loads, U rows in flight per warp (the hot loop of accumulate_kernel)
    for (int64_t chunk = lo + kidx; chunk < hi; chunk += n_kind) {
      uint32_t my_pix;
      int my_gid;
      load_entries(chunk, my_pix, my_gid);
#pragma unroll 1
      for (int j = 0; j < 32; j += U) {
        uint4 rows[U][VPL];
        int g[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const uint32_t pj = __shfl_sync(0xffffffffu, my_pix, j + u);
          g[u] = __shfl_sync(0xffffffffu, my_gid, j + u);
          const uint8_t* row = emb0 + (unsigned long long)pj * rb;
#pragma unroll
          for (int v = 0; v < VPL; ++v) rows[u][v] = g[u] >= 0 ? ld_stream_v4(row + v * 512) : make_uint4(0u, 0u, 0u, 0u);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          if (g[u] < 0) continue;
          if (g[u] != cur) {
            flush();
            cur = g[u];
          }
#pragma unroll
          for (int v = 0; v < VPL; ++v) RowVec<BF16>::add(acc + v * EPV, rows[u][v]);
        }
      }
      flush();
    }
  }

  if (CHECK) {
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t n_check = (int64_t)a.ctr->n_check;
    for (int64_t i = warp; i < n_check; i += n_warps) {
      const uint32_t pj = (uint32_t)a.entries[n_entries + i];
      const uint8_t* row = emb0 + (unsigned long long)pj * rb;
      bool bad = false;
#pragma unroll
      for (int v = 0; v < VPL; ++v) bad |= RowVec<BF16>::nonfinite(ld_stream_v4(row + v * 512));
      if (__any_sync(0xffffffffu, bad) && lane == 0) ++n_bad;
    }
    if (n_bad) atomicAdd(&a.ctr->n_bad_emb, (unsigned long long)n_bad);
  }
}

std::atomic<int> g_acc_mix{0};            // "acc_mix": per mille of the sorted list fetched with bulk copies (0: off)
std::atomic<int> g_acc_mix_ctas{3};       // "acc_mix_ctas": resident CTAs per SM of the mixed kernel

template <bool BF16, int VPL>
static int launch_accumulate_mix(const AccArgs& a, bool check, cudaStream_t s, int n_sm) {
  constexpr int DEPTH = VPL == 1 ? 24 : (VPL == 2 ? 12 : 4);
  constexpr int TW = 2;
  constexpr size_t smem = (size_t)TW * DEPTH * VPL * 512 + (size_t)TW * DEPTH * 8;
  static const cudaError_t opt0 = cudaFuncSetAttribute(accumulate_mix_kernel<BF16, VPL, false, DEPTH, TW>,
                                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  static const cudaError_t opt1 = cudaFuncSetAttribute(accumulate_mix_kernel<BF16, VPL, true, DEPTH, TW>,
                                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  VSM_CUDA(opt0);
  VSM_CUDA(opt1);
  const int grid = (n_sm > 0 ? n_sm : sm_count()) * std::max(1, std::min(3, g_acc_mix_ctas.load()));
  const float share = (float)g_acc_mix.load() * 1e-3f;
  if (check)
    accumulate_mix_kernel<BF16, VPL, true, DEPTH, TW><<<grid, 256, smem, s>>>(a, share);
  else
    accumulate_mix_kernel<BF16, VPL, false, DEPTH, TW><<<grid, 256, smem, s>>>(a, share);
  VSM_LAUNCHED();
  return VSM_OK;
}

// ---- bulk copies of RUNS of rows --------------------------------------------------------------------------------------
// A bulk copy costs ~46 cycles of the SM's TMA unit whatever its size: 1 KB rows cap an SM at ~43 GB/s (measured: the
// kernel above is no faster than register loads).  But consecutive entries of a voxel are often NEIGHBOURING PIXELS of an
// image row (the insert hands out ordinals in lane order of a 4x8 patch), whose embedding rows are contiguous in memory:
// a run of n such entries is ONE copy of n KB.  Rows are staged in groups of 8 entries (3 groups per warp in flight or
// being read, one mbarrier per group); inside a group, every run start issues its run.
template <bool BF16, int VPL, bool CHECK>
__global__ void __launch_bounds__(256, 1) accumulate_tmarun_kernel(AccArgs a) {
  constexpr int EPV = RowVec<BF16>::EPV;
  constexpr uint32_t ROW = (uint32_t)VPL * 512u;
  constexpr int G = 8, NB = 3, U = 4;
  extern __shared__ __align__(128) uint8_t run_smem[];
  const int lane = lane_id(), wib = (int)(threadIdx.x >> 5);
  const uint32_t ring0 = (uint32_t)__cvta_generic_to_shared(run_smem) + (uint32_t)wib * (uint32_t)(NB * G) * ROW;
  const uint32_t bar0 = (uint32_t)__cvta_generic_to_shared(run_smem) + 8u * (uint32_t)(NB * G) * ROW + (uint32_t)wib * (uint32_t)NB * 8u;
  if (lane < NB) mbar_init_cta(bar0 + (uint32_t)lane * 8u, 1u);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  if (a.ctr->abort) return;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int64_t n_entries = (int64_t)a.ctr->n_fused;
  const int64_t n_chunks = (n_entries + 31) >> 5;
  const uint32_t rb = (uint32_t)a.row_bytes;
  const uint8_t* emb_row0 = a.emb - a.pix_base * a.row_bytes;
  float* const vsum0 = a.vsum + (size_t)lane * EPV;
  const uint32_t d = (uint32_t)a.d;
  unsigned n_bad = 0;

  float acc[VPL * EPV];
#pragma unroll
  for (int i = 0; i < VPL * EPV; ++i) acc[i] = 0.f;
  int cur = -1;
  auto flush = [&]() {
    if (cur >= 0) {
      bool ok = true;
      if (CHECK) {
        float t = 0.f;
#pragma unroll
        for (int i = 0; i < VPL * EPV; ++i) t = fmaf(acc[i], 0.f, t);
        ok = !__any_sync(0xffffffffu, t != t);
        if (!ok && lane == 0) ++n_bad;
      }
      if (ok) {
        float* dst = vsum0 + (size_t)cur * d;
#pragma unroll
        for (int v = 0; v < VPL; ++v)
#pragma unroll
          for (int q = 0; q < EPV / 4; ++q)
            red_add_v4(dst + v * 32 * EPV + 4 * q, acc[v * EPV + 4 * q], acc[v * EPV + 4 * q + 1], acc[v * EPV + 4 * q + 2],
                       acc[v * EPV + 4 * q + 3]);
      }
    }
#pragma unroll
    for (int i = 0; i < VPL * EPV; ++i) acc[i] = 0.f;
  };
  auto load_entries = [&](int64_t chunk, uint32_t& pix, int& gid) {
    pix = 0u;
    gid = -1;
    if (chunk < n_chunks) {
      const int64_t i = (chunk << 5) + lane;
      if (i < n_entries) {
        const unsigned long long e = a.entries[i];
        pix = (uint32_t)e;
        gid = (int)(uint32_t)(e >> 32);
      }
    }
  };
  // group q (0..3) of the chunk whose entries the lanes hold, into buffer b: the valid entries' bytes are announced
  // on the buffer's barrier, then every run start copies its run.  Returns nothing; an empty group issues nothing.
  auto issue_group = [&](uint32_t pix, int gid, int q, uint32_t b) {
    const uint32_t ppix = __shfl_up_sync(0xffffffffu, pix, 1);
    const int pgid = __shfl_up_sync(0xffffffffu, gid, 1);
    const bool in_group = (lane >> 3) == q;
    const bool valid = in_group && gid >= 0;
    const bool start = valid && ((lane & 7) == 0 || pgid < 0 || pix != ppix + 1u);
    const unsigned vmask = __ballot_sync(0xffffffffu, valid);
    const unsigned smask = __ballot_sync(0xffffffffu, start);
    if (vmask == 0u) return;
    const uint32_t bar = bar0 + b * 8u;
    if (lane == (q << 3)) mbar_expect_tx_cta(bar, (uint32_t)__popc(vmask) * ROW);
    __syncwarp();
    if (start) {
      // my run ends before the next run start, the end of the valid entries, or the end of the group
      const unsigned after = (smask | ~vmask) & ~((2u << lane) - 1u) & (0xFFu << (q << 3));
      const int end = after ? __ffs(after) - 1 : ((q << 3) + 8);
      const uint32_t len = (uint32_t)(end - lane);
      bulk_g2s(ring0 + (b * (uint32_t)G + (uint32_t)(lane & 7)) * ROW, emb_row0 + (unsigned long long)pix * rb, len * ROW, bar);
    }
  };
  uint32_t pixC, pixN;
  int gidC, gidN;
  load_entries(warp, pixC, gidC);
  load_entries(warp + n_warps, pixN, gidN);
  // prologue: the first NB groups (all in the first chunk: NB <= 4)
#pragma unroll
  for (int g = 0; g < NB; ++g) issue_group(pixC, gidC, g, (uint32_t)g);
  uint32_t buf = 0, phases = 0u;
  for (int64_t chunk = warp; chunk < n_chunks; chunk += n_warps) {
#pragma unroll 1
    for (int q = 0; q < 4; ++q) {
      const unsigned vmask = __ballot_sync(0xffffffffu, (lane >> 3) == q && gidC >= 0);
      if (vmask) {
        mbar_wait_cta(bar0 + buf * 8u, (phases >> buf) & 1u);
        phases ^= 1u << buf;
      }
#pragma unroll
      for (int h = 0; h < G / U; ++h) {
        int g[U];
        uint4 rows[U][VPL];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          g[u] = __shfl_sync(0xffffffffu, gidC, (q << 3) + h * U + u);
          if (g[u] >= 0) {
#pragma unroll
            for (int v = 0; v < VPL; ++v)
              rows[u][v] = lds_v4(ring0 + (buf * (uint32_t)G + (uint32_t)(h * U + u)) * ROW + (uint32_t)v * 512u + (uint32_t)lane * 16u);
          }
        }
        if (h == G / U - 1) {
          // the group has been read: its buffer takes the group NB ahead (this chunk's, or the next chunk's)
          __syncwarp();
          const int qq = q + NB;
          if (qq < 4)
            issue_group(pixC, gidC, qq, buf);
          else
            issue_group(pixN, gidN, qq - 4, buf);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          if (g[u] < 0) continue;
          if (g[u] != cur) {
            flush();
            cur = g[u];
          }
#pragma unroll
          for (int v = 0; v < VPL; ++v) RowVec<BF16>::add(acc + v * EPV, rows[u][v]);
        }
      }
      buf = buf + 1u == (uint32_t)NB ? 0u : buf + 1u;
    }
    pixC = pixN;
    gidC = gidN;
    load_entries(chunk + 2 * n_warps, pixN, gidN);
  }
  flush();

  if (CHECK) {
    const uint8_t* emb0 = emb_row0 + (size_t)lane * 16;
    const int64_t n_check = (int64_t)a.ctr->n_check;
    for (int64_t i = warp; i < n_check; i += n_warps) {
      const uint32_t pj = (uint32_t)a.entries[n_entries + i];
      const uint8_t* row = emb0 + (unsigned long long)pj * rb;
      bool bad = false;
#pragma unroll
      for (int v = 0; v < VPL; ++v) bad |= RowVec<BF16>::nonfinite(ld_stream_v4(row + v * 512));
      if (__any_sync(0xffffffffu, bad) && lane == 0) ++n_bad;
    }
    if (n_bad) atomicAdd(&a.ctr->n_bad_emb, (unsigned long long)n_bad);
  }
}

std::atomic<int> g_acc_tmarun{0};  // "acc_tmarun": bulk copies of runs of neighbouring rows

template <bool BF16, int VPL>
static int launch_accumulate_tmarun(const AccArgs& a, bool check, cudaStream_t s, int n_sm) {
  constexpr size_t smem = (size_t)8 * 24 * VPL * 512 + (size_t)8 * 3 * 8;
  static const cudaError_t opt0 = cudaFuncSetAttribute(accumulate_tmarun_kernel<BF16, VPL, false>,
                                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  static const cudaError_t opt1 = cudaFuncSetAttribute(accumulate_tmarun_kernel<BF16, VPL, true>,
                                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  VSM_CUDA(opt0);
  VSM_CUDA(opt1);
  const int grid = n_sm > 0 ? n_sm : sm_count();
  if (check)
    accumulate_tmarun_kernel<BF16, VPL, true><<<grid, 256, smem, s>>>(a);
  else
    accumulate_tmarun_kernel<BF16, VPL, false><<<grid, 256, smem, s>>>(a);
  VSM_LAUNCHED();
  return VSM_OK;
}

std::atomic<int> g_acc_tma{0};  // "acc_tma": the TMA bulk-copy accumulate for voxel-sorted lists with full rows

template <bool BF16, int VPL>
static int launch_accumulate_tma(const AccArgs& a, bool check, cudaStream_t s, int n_sm) {
  constexpr int DEPTH = VPL == 1 ? 24 : (VPL == 2 ? 12 : (VPL == 4 ? 4 : 4));
  constexpr size_t smem = (size_t)8 * DEPTH * VPL * 512 + (size_t)8 * DEPTH * 8;
  static const cudaError_t opt0 = cudaFuncSetAttribute(accumulate_tma_kernel<BF16, VPL, false, DEPTH>,
                                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  static const cudaError_t opt1 = cudaFuncSetAttribute(accumulate_tma_kernel<BF16, VPL, true, DEPTH>,
                                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  VSM_CUDA(opt0);
  VSM_CUDA(opt1);
  const int grid = (n_sm > 0 ? n_sm : sm_count()) * 2;
  if (check)
    accumulate_tma_kernel<BF16, VPL, true, DEPTH><<<grid, 256, smem, s>>>(a);
  else
    accumulate_tma_kernel<BF16, VPL, false, DEPTH><<<grid, 256, smem, s>>>(a);
  VSM_LAUNCHED();
  return VSM_OK;
}

// Measured on B200 (config 2, 8 submaps): 1.08 ms per submap against 1.01 ms with accumulate_kernel on the whole device,
// 0.95 against 0.90 with 64 SMs set aside for the preparation kernels, and no better on 60 SMs -- twice the bytes in
// flight per SM do not buy bandwidth on fewer SMs here.  Kept as an option ("acc_ring"), off by default.
std::atomic<int> g_acc_ring{0};

template <bool BF16, int VPL>
static int launch_accumulate_ring(const AccArgs& a, bool check, cudaStream_t s, int n_sm) {
  // ring depth: ~96 KB of rows per CTA of 8 warps, two CTAs per SM
  constexpr int DEPTH = VPL == 1 ? 24 : (VPL == 2 ? 12 : (VPL == 4 ? 6 : 3));
  constexpr size_t smem = (size_t)8 * DEPTH * VPL * 512;
  static const cudaError_t opt0 = cudaFuncSetAttribute(accumulate_ring_kernel<BF16, VPL, false, DEPTH>,
                                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  static const cudaError_t opt1 = cudaFuncSetAttribute(accumulate_ring_kernel<BF16, VPL, true, DEPTH>,
                                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  VSM_CUDA(opt0);
  VSM_CUDA(opt1);
  const int grid = (n_sm > 0 ? n_sm : sm_count()) * 2;
  if (check)
    accumulate_ring_kernel<BF16, VPL, true, DEPTH><<<grid, 256, smem, s>>>(a);
  else
    accumulate_ring_kernel<BF16, VPL, false, DEPTH><<<grid, 256, smem, s>>>(a);
  VSM_LAUNCHED();
  return VSM_OK;
}

template <bool BF16, int VPL>
static int launch_accumulate_t(const AccArgs& a, bool sorted, bool check, cudaStream_t s, int n_sm) {
  const int block = 256;
  // The sorted list's length lives on the device: a fixed grid of resident CTAs strides over the chunks.  Two CTAs
  // per SM (16 warps x 4 rows of 1 KB in flight) already saturate HBM (measured: same time as 3 or 6 per SM).
  // n_sm: the SMs the stream may use (an SM partition), 0 = all.
  int grid = (n_sm > 0 ? n_sm : sm_count()) * std::max(1, std::min(8, g_acc_ctas_per_sm.load()));
  if (!sorted) {
    const int64_t n_chunks = (a.n + 31) >> 5;
    grid = (int)std::min<int64_t>(std::max<int64_t>(cdiv(n_chunks, block / 32), 1), (int64_t)sm_count() * 6);
  }
  const bool full = a.nvec == 32 * VPL;
  static const bool env_run = getenv("VSM_ACC_TMARUN") && getenv("VSM_ACC_TMARUN")[0] == '1';
  if (sorted && full && VPL <= 2 && a.emb_index == nullptr && (g_acc_tmarun.load() || env_run) &&
      (reinterpret_cast<uintptr_t>(a.emb) & 15u) == 0)
    return launch_accumulate_tmarun<BF16, VPL>(a, check, s, n_sm);
  static const int env_mix = getenv("VSM_ACC_MIX") ? atoi(getenv("VSM_ACC_MIX")) : 0;
  if (sorted && full && VPL <= 4 && a.emb_index == nullptr && (g_acc_mix.load() > 0 || env_mix > 0) &&
      (reinterpret_cast<uintptr_t>(a.emb) & 15u) == 0) {
    if (g_acc_mix.load() == 0) g_acc_mix = env_mix;
    return launch_accumulate_mix<BF16, VPL>(a, check, s, n_sm);
  }
  static const bool env_tma = getenv("VSM_ACC_TMA") && getenv("VSM_ACC_TMA")[0] == '1';  // (for running the test suite on it)
  if (sorted && full && VPL <= 4 && a.emb_index == nullptr && (g_acc_tma.load() || env_tma) && (reinterpret_cast<uintptr_t>(a.emb) & 15u) == 0)
    return launch_accumulate_tma<BF16, VPL>(a, check, s, n_sm);
  if (sorted && full && a.emb_index == nullptr && g_acc_ring.load()) return launch_accumulate_ring<BF16, VPL>(a, check, s, n_sm);
  if (sorted && full && a.emb_index != nullptr) {
    if (check)
      accumulate_kernel<BF16, VPL, true, true, true, true><<<grid, block, 0, s>>>(a);
    else
      accumulate_kernel<BF16, VPL, true, false, true, true><<<grid, block, 0, s>>>(a);
    VSM_LAUNCHED();
    return VSM_OK;
  }
#define VSM_ACC(SORTED_, CHECK_)                                                      \
  do {                                                                                \
    if (full)                                                                         \
      accumulate_kernel<BF16, VPL, SORTED_, CHECK_, true><<<grid, block, 0, s>>>(a);  \
    else                                                                              \
      accumulate_kernel<BF16, VPL, SORTED_, CHECK_, false><<<grid, block, 0, s>>>(a); \
  } while (0)
  if (sorted) {
    if (check)
      VSM_ACC(true, true);
    else
      VSM_ACC(true, false);
  } else {
    if (check)
      VSM_ACC(false, true);
    else
      VSM_ACC(false, false);
  }
#undef VSM_ACC
  VSM_LAUNCHED();
  return VSM_OK;
}

int launch_accumulate(const AccArgs& a, bool bf16, bool sorted, bool check, cudaStream_t s, int n_sm = 0) {
  if (!sorted && a.n <= 0) return VSM_OK;
  const int vpl = (a.nvec + 31) / 32;
  if (bf16) {
    if (vpl <= 1) return launch_accumulate_t<true, 1>(a, sorted, check, s, n_sm);
    if (vpl <= 2) return launch_accumulate_t<true, 2>(a, sorted, check, s, n_sm);
    if (vpl <= 4) return launch_accumulate_t<true, 4>(a, sorted, check, s, n_sm);
    if (vpl <= 8) return launch_accumulate_t<true, 8>(a, sorted, check, s, n_sm);
  } else {
    if (vpl <= 1) return launch_accumulate_t<false, 1>(a, sorted, check, s, n_sm);
    if (vpl <= 2) return launch_accumulate_t<false, 2>(a, sorted, check, s, n_sm);
    if (vpl <= 4) return launch_accumulate_t<false, 4>(a, sorted, check, s, n_sm);
    if (vpl <= 8) return launch_accumulate_t<false, 8>(a, sorted, check, s, n_sm);
  }
  set_error("embedding dimension %d too large (max %d)", a.d, bf16 ? 2048 : 1024);
  return VSM_E_INVALID;
}

// indexed embeddings: every pixel that can be selected must point into the table (checked before anything is fused)
__global__ void __launch_bounds__(256) index_check_kernel(const float* __restrict__ conf, const int32_t* __restrict__ emb_index,
                                                          uint32_t n_px, float thr, int32_t emb_rows, FuseCounters* ctr) {
  unsigned bad = 0;
  for (uint32_t pix = blockIdx.x * blockDim.x + threadIdx.x; pix < n_px; pix += gridDim.x * blockDim.x) {
    const int32_t r = emb_index[pix];
    if ((r < 0 || r >= emb_rows) && conf[pix] >= thr) ++bad;
  }
  for (int o = 16; o > 0; o >>= 1) bad += __shfl_xor_sync(0xffffffffu, bad, o);
  if (lane_id() == 0 && bad) atomicAdd(&ctr->bad_index, bad);
}

// uint8 row mask: 1 where the pixel passes conf/stride/end_idx and all d channels are finite (map.py:247)
template <bool BF16>
__global__ void __launch_bounds__(256) emb_row_mask_kernel(const float* __restrict__ conf, const uint8_t* __restrict__ emb,
                                                           const int32_t* __restrict__ emb_index, int32_t emb_rows,
                                                           int64_t row_bytes, int nvec, int64_t n_px, int H, int W,
                                                           int stride, float thr, uint8_t* __restrict__ out) {
  const int lane = lane_id();
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t pix = warp; pix < n_px; pix += n_warps) {
    bool on = conf[pix] >= thr;
    if (on && stride > 1) {
      const int w = (int)(pix % W), h = (int)((pix / W) % H);
      on = (w % stride == 0) && (h % stride == 0);
    }
    bool bad = false;
    if (on) {
      int64_t r = pix;
      if (emb_index) {
        r = emb_index[pix];
        if (r < 0 || r >= emb_rows) r = -1;  // reported by index_check_kernel; never dereferenced
      }
      if (r < 0) {
        bad = true;
      } else {
        const uint8_t* row = emb + r * row_bytes;
        for (int c = lane; c < nvec; c += 32) bad |= RowVec<BF16>::nonfinite(ld_stream_v4(row + (size_t)c * 16));
      }
    }
    bad = __any_sync(0xffffffffu, bad);
    if (lane == 0) out[pix] = (on && !bad) ? 1 : 0;
  }
}


int map_grow(vsm_map* m, int64_t need, cudaStream_t s) {
  if (need <= m->vcap && m->gcap >= (uint64_t)2 * (uint64_t)std::max<int64_t>(m->vcap, 1)) return VSM_OK;
  if (need >= (int64_t)1 << 31) {
    set_error("voxel capacity %lld exceeds 2^31", (long long)need);
    return VSM_E_INVALID;
  }
  int64_t new_cap = std::max<int64_t>(need, m->vcap + m->vcap / 2);
  new_cap = std::max<int64_t>(new_cap, 1024);
  const size_t d = (size_t)m->d;
  const size_t keep = (size_t)m->n_vox;
  VSM_CUDA(cudaStreamSynchronize(s));
  // dense arrays (new space zeroed)
  {
    DevBuf nk, nc, ns;
    VSM_TRY(nk.ensure((size_t)new_cap * 8, s));
    VSM_TRY(nc.ensure((size_t)new_cap * 4, s));
    VSM_TRY(ns.ensure((size_t)new_cap * d * 4, s));
    VSM_CUDA(cudaMemsetAsync(nc.p, 0, (size_t)new_cap * 4, s));
    VSM_CUDA(cudaMemsetAsync(ns.p, 0, (size_t)new_cap * d * 4, s));
    if (keep) {
      VSM_CUDA(cudaMemcpyAsync(nk.p, m->vkey.p, keep * 8, cudaMemcpyDeviceToDevice, s));
      VSM_CUDA(cudaMemcpyAsync(nc.p, m->vcount.p, keep * 4, cudaMemcpyDeviceToDevice, s));
      VSM_CUDA(cudaMemcpyAsync(ns.p, m->vsum.p, keep * d * 4, cudaMemcpyDeviceToDevice, s));
    }
    VSM_CUDA(cudaStreamSynchronize(s));
    m->vkey.release();
    m->vcount.release();
    m->vsum.release();
    m->vkey = nk;
    m->vcount = nc;
    m->vsum = ns;
  }
  m->vcap = new_cap;
  // hash table at load <= 0.5
  const uint64_t gcap = next_pow2((uint64_t)2 * (uint64_t)new_cap);
  if (gcap != m->gcap) {
    m->gkeys.release();
    m->gids.release();
    VSM_TRY(m->gkeys.ensure(gcap * 8, s));
    VSM_TRY(m->gids.ensure(gcap * 4, s));
    m->gcap = gcap;
    VSM_CUDA(cudaMemsetAsync(m->gkeys.p, 0xFF, gcap * 8, s));
    VSM_CUDA(cudaMemsetAsync(m->gids.p, 0xFF, gcap * 4, s));
    if (keep && !m->dense_loaded) {
      rehash_kernel<<<grid_for((int64_t)keep, 256), 256, 0, s>>>(global_store(m), (uint32_t)keep);
      VSM_LAUNCHED();
    }
  }
  return VSM_OK;
}

int log_grow(vsm_map* m, int64_t need, cudaStream_t s) {
  if (need <= m->log_cap) return VSM_OK;
  const int64_t new_cap = std::max<int64_t>(std::max<int64_t>(need, m->log_cap * 2), m->vcap);
  const size_t keep = (size_t)m->log_n;
  VSM_TRY(m->log_gid.ensure((size_t)new_cap * 4, s, keep * 4));
  VSM_TRY(m->log_fuse.ensure((size_t)new_cap * 4, s, keep * 4));
  VSM_TRY(m->log_mask.ensure((size_t)new_cap * 16, s, keep * 16));
  m->log_cap = new_cap;
  return VSM_OK;
}

static int ensure_local_table(DevBuf& slots, DevBuf& list, uint64_t& cap, uint64_t need_cap, cudaStream_t s) {
  if (cap >= need_cap) return VSM_OK;
  VSM_TRY(slots.ensure(need_cap * sizeof(Slot), s));
  VSM_TRY(list.ensure(need_cap / 2 * 4 + 256, s));
  table_init_kernel<<<grid_for((int64_t)need_cap, 256), 256, 0, s>>>(slots.as<Slot>(), need_cap);
  VSM_LAUNCHED();
  cap = need_cap;
  return VSM_OK;
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

static int validate_params(const vsm_map* m, const vsm_fuse_params* p) {
  if (!m || !p) {
    set_error("null map or params");
    return VSM_E_INVALID;
  }
  if (p->S <= 0 || p->H <= 0 || p->W <= 0 || p->end_idx < 0 || p->end_idx > p->S) {
    set_error("bad dims S=%d H=%d W=%d end_idx=%d", p->S, p->H, p->W, p->end_idx);
    return VSM_E_INVALID;
  }
  if (p->stride < 1) {
    set_error("stride must be >= 1");  // map.py:187-188
    return VSM_E_INVALID;
  }
  if (p->frame_base < 0 || p->end_idx + p->frame_base > VSM_MAX_FRAMES) {
    set_error("%d frames in one submap, max %d", p->end_idx + p->frame_base, VSM_MAX_FRAMES);
    return VSM_E_TOO_MANY_FRAMES;
  }
  if ((int64_t)p->end_idx * p->H * p->W >= ((int64_t)1 << 32)) {
    set_error("more than 2^32 pixels in one fuse call");
    return VSM_E_INVALID;
  }
  if (m->dense_loaded) {
    set_error("map was loaded from dense rows; fusing into it is not supported");
    return VSM_E_STATE;
  }
  if (p->emb_index_dev != nullptr && p->emb_rows < 1) {
    set_error("indexed embeddings need emb_rows >= 1 (got %d)", p->emb_rows);
    return VSM_E_INVALID;
  }
  return VSM_OK;
}


// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
// A fuse call is ENQUEUED (all kernels queued on the caller's stream, nothing read back) and later COLLECTED
// (one stream synchronisation for any number of queued calls; counters read, aborted calls repeated after the
// map has grown).  vsm_fuse_submap = enqueue + collect; vsm_fuse_submap_async / vsm_fuse_collect expose the
// two halves so that a whole build runs without the host ever waiting on the device between submaps.
struct HostEmb {
  const uint8_t* emb_host = nullptr;  // (S,H,W,d) rows on the host, map dtype
  // device-side address of the same rows when they live in pinned (mapped) host memory: the accumulate kernel then
  // reads the rows it needs straight over PCIe -- rows of pixels that were not selected never cross the bus
  const uint8_t* emb_mapped = nullptr;
};

static int ensure_stream_objects(vsm_map* m, size_t chunk_bytes) {
  if (!m->copy_stream) VSM_CUDA(cudaStreamCreateWithFlags(&m->copy_stream, cudaStreamNonBlocking));
  for (int b = 0; b < 2; ++b) {
    if (!m->ev_stage[b]) VSM_CUDA(cudaEventCreateWithFlags(&m->ev_stage[b], cudaEventDisableTiming));
    if (!m->ev_copy[b]) VSM_CUDA(cudaEventCreateWithFlags(&m->ev_copy[b], cudaEventDisableTiming));
    VSM_TRY(m->stage_emb[b].ensure(chunk_bytes, nullptr));
  }
  return VSM_OK;
}

static LocalTable table_view(DevBuf& slots, DevBuf& list, uint64_t cap, uint32_t* n_occ, uint32_t* overflow) {
  LocalTable t{};
  t.slots = slots.as<Slot>();
  t.slot_list = list.as<uint32_t>();
  t.n_occ = n_occ;
  t.cap_mask = (uint32_t)(cap - 1);
  t.limit = (uint32_t)(cap - cap / 4);
  t.overflow = overflow;
  return t;
}

static int fuse_collect_locked(vsm_map* m, cudaStream_t s, std::vector<vsm_fuse_stats>* out);

static int ensure_acc_stream(Workspace* ws) {
  if (ws->acc_stream) return VSM_OK;
  const char* e = getenv("VSM_OVERLAP");
  ws->overlap = e && e[0] == '1';
  VSM_CUDA(cudaStreamCreateWithFlags(&ws->acc_stream, cudaStreamNonBlocking));
  for (int b = 0; b < 2; ++b) {
    VSM_CUDA(cudaEventCreateWithFlags(&ws->ev_prep_done[b], cudaEventDisableTiming));
    VSM_CUDA(cudaEventCreateWithFlags(&ws->ev_acc_done[b], cudaEventDisableTiming));
  }
  return VSM_OK;
}

// ---- SM partitions (CUDA green contexts; driver entry points looked up at run time: libvsm does not link libcuda) ----
template <class F>
static bool driver_fn(const char* name, F* out) {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult qr;
  if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &qr) != cudaSuccess || !p || qr != cudaDriverEntryPointSuccess) {
    cudaGetLastError();
    set_error("green contexts: %s is not available from this driver", name);
    return false;
  }
  *out = reinterpret_cast<F>(p);
  return true;
}

static void green_deactivate(Workspace* ws) {
  for (int i = 0; i < 2; ++i) {
    ws->green_ctx[i] = nullptr;
    ws->green_stream[i] = nullptr;
    ws->green_sms[i] = 0;
  }
  ws->green_on = false;
}

// Splits the device's SMs into a preparation partition of >= prep_sms SMs (rounded up to the hardware's granularity,
// 8 on sm_100) and an accumulate partition of the rest; one stream in each.  Needs the workspace lock and an idle
// device.  Partitions are created once per requested size and kept (prep_sms = 0 only deactivates).
static int green_setup(Workspace* ws, int dev, int prep_sms) {
  green_deactivate(ws);
  if (prep_sms <= 0) return VSM_OK;
  for (const Workspace::GreenPart& gp : ws->green_parts)
    if (gp.requested == prep_sms) {
      for (int i = 0; i < 2; ++i) {
        ws->green_ctx[i] = gp.ctx[i];
        ws->green_stream[i] = gp.stream[i];
        ws->green_sms[i] = gp.sms[i];
      }
      ws->green_on = true;
      return VSM_OK;
    }
  CUresult (*dev_get)(CUdevice*, int) = nullptr;
  CUresult (*get_res)(CUdevice, CUdevResource*, CUdevResourceType) = nullptr;
  CUresult (*split)(CUdevResource*, unsigned int*, const CUdevResource*, CUdevResource*, unsigned int, unsigned int) = nullptr;
  CUresult (*gen_desc)(CUdevResourceDesc*, CUdevResource*, unsigned int) = nullptr;
  CUresult (*ctx_create)(CUgreenCtx*, CUdevResourceDesc, CUdevice, unsigned int) = nullptr;
  CUresult (*stream_create)(CUstream*, CUgreenCtx, unsigned int, int) = nullptr;
  if (!driver_fn("cuDeviceGet", &dev_get) || !driver_fn("cuDeviceGetDevResource", &get_res) ||
      !driver_fn("cuDevSmResourceSplitByCount", &split) || !driver_fn("cuDevResourceGenerateDesc", &gen_desc) ||
      !driver_fn("cuGreenCtxCreate", &ctx_create) || !driver_fn("cuGreenCtxStreamCreate", &stream_create))
    return VSM_E_CUDA;
  CUdevice cu_dev;
  CUdevResource all, part[2];
  unsigned int groups = 1;
  CUresult r = dev_get(&cu_dev, dev);
  if (r == CUDA_SUCCESS) r = get_res(cu_dev, &all, CU_DEV_RESOURCE_TYPE_SM);
  if (r == CUDA_SUCCESS) r = split(&part[0], &groups, &all, &part[1], 0u, (unsigned int)prep_sms);
  if (r != CUDA_SUCCESS || groups != 1 || part[1].sm.smCount == 0) {
    set_error("green contexts: cannot split %d SMs off the device (CUresult %d)", prep_sms, (int)r);
    return VSM_E_INVALID;
  }
  Workspace::GreenPart gp{};
  gp.requested = prep_sms;
  for (int i = 0; i < 2; ++i) {
    CUdevResourceDesc desc;
    CUgreenCtx ctx = nullptr;
    CUstream st = nullptr;
    r = gen_desc(&desc, &part[i], 1);
    if (r == CUDA_SUCCESS) r = ctx_create(&ctx, desc, cu_dev, CU_GREEN_CTX_DEFAULT_STREAM);
    if (r == CUDA_SUCCESS) r = stream_create(&st, ctx, CU_STREAM_NON_BLOCKING, 0);
    if (r != CUDA_SUCCESS) {
      set_error("green contexts: cannot create partition %d (CUresult %d)", i, (int)r);
      return VSM_E_CUDA;  // (a context created for i = 0 stays allocated: the failure is not expected to repeat)
    }
    gp.ctx[i] = ctx;
    gp.stream[i] = reinterpret_cast<cudaStream_t>(st);
    gp.sms[i] = (int)part[i].sm.smCount;
  }
  if (!ws->ev_fork) VSM_CUDA(cudaEventCreateWithFlags(&ws->ev_fork, cudaEventDisableTiming));
  if (!ws->ev_prep_join) VSM_CUDA(cudaEventCreateWithFlags(&ws->ev_prep_join, cudaEventDisableTiming));
  ws->green_parts.push_back(gp);
  for (int i = 0; i < 2; ++i) {
    ws->green_ctx[i] = gp.ctx[i];
    ws->green_stream[i] = gp.stream[i];
    ws->green_sms[i] = gp.sms[i];
  }
  ws->green_on = true;
  return VSM_OK;
}

// moves a fuse call's preparation kernels from the caller's stream to the preparation partition and back
struct PrepFork {
  Workspace* ws;
  cudaStream_t user;
  bool on = false;
  PrepFork(Workspace* w, cudaStream_t u) : ws(w), user(u) {}
  int fork(cudaStream_t* s) {
    VSM_CUDA(cudaEventRecord(ws->ev_fork, user));
    VSM_CUDA(cudaStreamWaitEvent(ws->green_stream[0], ws->ev_fork, 0));
    *s = ws->green_stream[0];
    on = true;
    return VSM_OK;
  }
  ~PrepFork() {
    // every return path: the caller's stream continues behind the preparation kernels queued so far
    if (on && (cudaEventRecord(ws->ev_prep_join, ws->green_stream[0]) != cudaSuccess ||
               cudaStreamWaitEvent(user, ws->ev_prep_join, 0) != cudaSuccess))
      cudaGetLastError();
  }
};

int join_accumulates(Workspace* ws, cudaStream_t s) {
  for (int b = 0; b < 2; ++b)
    if (ws->acc_used[b]) VSM_CUDA(cudaStreamWaitEvent(s, ws->ev_acc_done[b], 0));
  return VSM_OK;
}

// Queue one fuse call.  Needs the workspace lock.  `host` != null: embeddings are streamed from the host.
static int fuse_enqueue(vsm_map* m, PendingCall& call, const HostEmb* host, cudaStream_t s) {
  Workspace* ws = m->ws;
  const vsm_fuse_params* p = &call.p;
  const bool filters = (p->flags & VSM_FUSE_FILTERS) != 0;
  const bool keep_index = (p->flags & VSM_FUSE_KEEP_POINT_INDEX) != 0;
  const bool pixel_order = (p->flags & VSM_FUSE_PIXEL_ORDER) != 0 || host != nullptr;
  const bool bf16 = m->cfg.emb_dtype == VSM_BF16;
  const int64_t n_px = (int64_t)p->end_idx * p->H * p->W;
  const int64_t px_per_frame = (int64_t)p->H * p->W;
  const int64_t row_bytes = (int64_t)m->d * m->esize;
  const float* pts = call.pts;
  const float* conf = call.conf;
  const uint8_t* emb_dev = call.emb;
  const uint8_t* emb_ok = call.emb_ok;
  const int32_t* emb_index = p->emb_index_dev;

  VSM_TRY(m->ctr_ring.ensure(sizeof(FuseCounters) * kCallRing, s));
  FuseCounters* ctr = m->ctr_ring.as<FuseCounters>() + call.slot;
  VSM_CUDA(cudaMemsetAsync(ctr, 0, sizeof(FuseCounters), s));
  m->finalized = false;
  m->ck_built = false;
  call.profiled = m->profiling && host == nullptr;
  cudaEvent_t* ev = m->ev_ring[call.slot];
  if (call.profiled) {
    for (int i = 0; i < 3; ++i)
      if (!ev[i]) VSM_CUDA(cudaEventCreate(&ev[i]));
    VSM_CUDA(cudaEventRecord(ev[0], s));
  }
  if (n_px == 0) return VSM_OK;

  // ---- scratch, sized by upper bounds (the exact sizes only exist on the device) -----------------------
  // DevBuf::ensure synchronises the stream before it replaces a block, so queued calls never lose a buffer.
  const int64_t hs = cdiv(p->H, p->stride), ws_ = cdiv(p->W, p->stride);
  const uint64_t n_sel_max = (uint64_t)p->end_idx * hs * ws_;
  const uint64_t lcap = std::max<uint64_t>(next_pow2(2 * n_sel_max), 1024);
  VSM_TRY(ws->pw.ensure((size_t)n_px * 16, s));
  VSM_TRY(ws->pt_slot.ensure((size_t)n_px * 4, s));
  VSM_TRY(ensure_local_table(ws->tb_slots, ws->tb_list, ws->tb_cap, lcap, s));
  if (filters) VSM_TRY(ensure_local_table(ws->ta_slots, ws->ta_list, ws->ta_cap, lcap, s));
  VSM_TRY(ws->lv_off.ensure((size_t)n_sel_max * 4, s));
  VSM_TRY(ws->lv_cursor.ensure((size_t)n_sel_max * 4, s));
  VSM_TRY(ws->lv_gid.ensure((size_t)n_sel_max * 4, s));
  if (filters && (g_prep_variant.load() & 8)) {  // one-table preparation: (slot, ordinal) per pixel, kept list, irregular points
    VSM_TRY(ws->pt2.ensure((size_t)n_px * 8, s));
    VSM_TRY(ws->kept.ensure((size_t)n_sel_max * 4, s));
    VSM_TRY(ws->irr.ensure((size_t)(1u << 16) * 32, s));
  }
  if (filters && g_select_mode.load() != 1) VSM_TRY(ws->sel_bracket.ensure(bracket_scratch_bytes(n_px), s));
  VSM_TRY(ensure_acc_stream(ws));
  const int ab = ws->acc_parity;  // which sorted list this call fills
  if (!pixel_order) {
    // the accumulate kernel that last read this list (two calls ago) must be done before it is replaced or refilled
    if (ws->acc_used[ab]) VSM_CUDA(cudaStreamWaitEvent(s, ws->ev_acc_done[ab], 0));
    VSM_TRY(ws->sorted_pix[ab].ensure((size_t)n_sel_max * 8, s));  // packed entries
  }

  // SM partitions: from here on the call's preparation kernels go to the preparation partition's stream (the plain
  // device-resident, voxel-sorted call only: the optional paths allocate per-call buffers on the stream they run on)
  const int sel_mode = g_select_mode.load() == 0 ? 3 : g_select_mode.load();  // (3 needs the one-table preparation: else radix)
  PrepFork prep_fork(ws, s);
  const bool green = ws->overlap && ws->green_on && !pixel_order && !keep_index &&
                     !(filters && (p->flags & VSM_FUSE_EMB_PRECHECK) && emb_ok == nullptr);
  if (green) {
    if (filters) {  // the select scratch is allocated on first use: do that on the caller's stream
      SelectState* st0 = nullptr;
      uint32_t* h0 = nullptr;
      float* o0 = nullptr;
      VSM_TRY(select_scratch(&st0, &h0, &o0));
    }
    VSM_TRY(prep_fork.fork(&s));
  }

  // The tables are allocated for the worst case (every selected pixel a voxel of its own: 2^24 slots, 0.5 GB, at the
  // benchmark's shape) but a call only uses a prefix sized from the most voxels any call on this device has had: 4x
  // that, a power of two -- 16 MB at 5 cm, resident in L2 instead of probed at random in HBM.  A call that outgrows
  // the prefix is stopped by the device before it touches the map and repeated with the whole table.
  uint64_t cap_b = lcap, cap_a = lcap;  // (the allocation can be larger than this call's worst case: probe no more than that)
  if (g_small_tables.load() && ws->hint_n_occ > 0 && !(p->flags & kFuseForceBigTables)) {
    // the hint was taken at voxel size hint_vs: surfaces fill (hint_vs / vs)^2 as many voxels of size vs
    double scale = ws->hint_vs > 0.0 ? (ws->hint_vs / m->cfg.voxel_size) * (ws->hint_vs / m->cfg.voxel_size) : 1.0;
    scale = std::min(64.0, std::max(1.0 / 64.0, scale));
    const int64_t hint = std::max<int64_t>(m->last_n_occ, (int64_t)((double)ws->hint_n_occ * scale));
    const uint64_t want = next_pow2(4 * (uint64_t)std::max<int64_t>(hint, 4096));
    cap_b = std::min<uint64_t>(lcap, want);
    cap_a = std::min<uint64_t>(lcap, std::max<uint64_t>(want / 4, 1u << 14));
  }
  LocalTable tb = table_view(ws->tb_slots, ws->tb_list, cap_b, &ctr->n_occ_b, &ctr->tbl_overflow);
  LocalTable ta{};
  if (filters) ta = table_view(ws->ta_slots, ws->ta_list, cap_a, &ctr->n_occ_a, &ctr->tbl_overflow);

  if (emb_index != nullptr) {
    index_check_kernel<<<grid_for(n_px, 256), 256, 0, s>>>(conf, emb_index, (uint32_t)n_px, p->conf_threshold, p->emb_rows, ctr);
    VSM_LAUNCHED();
  }
  // optional exact finite-row filter on the embeddings (second read of the rows)
  if (filters && (p->flags & VSM_FUSE_EMB_PRECHECK) && emb_ok == nullptr && emb_dev != nullptr) {
    VSM_TRY(call.precheck_mask.ensure((size_t)n_px, s));
    const int nvec = (int)(row_bytes / 16);
    if (bf16)
      emb_row_mask_kernel<true><<<sm_count() * 8, 256, 0, s>>>(conf, emb_dev, emb_index, p->emb_rows, row_bytes, nvec, n_px, p->H, p->W,
                                                        p->stride, p->conf_threshold, call.precheck_mask.as<uint8_t>());
    else
      emb_row_mask_kernel<false><<<sm_count() * 8, 256, 0, s>>>(conf, emb_dev, emb_index, p->emb_rows, row_bytes, nvec, n_px, p->H, p->W,
                                                         p->stride, p->conf_threshold, call.precheck_mask.as<uint8_t>());
    VSM_LAUNCHED();
    emb_ok = call.precheck_mask.as<uint8_t>();
  }
  const bool check = filters && emb_ok == nullptr;  // optimistic non-finite detection on the embeddings

  // ---- world points (+ pass 0 of the radix select) ------------------------------------------------------
  // The bbox percentiles come from the three-pass radix select (select.cu).  The bracket select (bracket.cuh: sample,
  // collect inside the world-point kernel, exact resolve; repeated with the radix select when it cannot answer) gives
  // the same bounds and is kept as an option: on B200 it measured 6 % SLOWER per fuse call (its collect step more
  // than doubles the instruction-bound world-point kernel), see DESIGN.md 5.
  const int variant = g_prep_variant.load();
  // one-table preparation: an integer number of voxels per coarse cell (the reference's 3) is what lets the cell of
  // a voxel follow from its coordinates
  const double cf = p->coarse_factor;
  const bool v7 = filters && (variant & 8) != 0 && !(p->flags & kFuseForceTwoTables) && cf >= 1.0 && cf <= 1024.0 &&
                  cf == (double)(int64_t)cf;
  // select_mode 2: brackets collected by the world-point kernel; 3 (one-table preparation only): brackets applied and
  // collected by the insert kernel, no percentile pass at all
  const bool deferred = v7 && !(p->flags & kFuseForceRadix) && sel_mode == 3 && ws->sel_bracket.p != nullptr;
  const bool bracket = filters && !(p->flags & kFuseForceRadix) && (sel_mode == 2 || deferred) && ws->sel_bracket.p != nullptr;
  SelectState* sst = nullptr;
  uint32_t* hist = nullptr;
  float* sel_out = nullptr;
  if (filters && !bracket) {
    VSM_TRY(select_scratch(&sst, &hist, &sel_out));
    VSM_TRY(select_reset(sst, hist, s));
  }
  HMat Hm;
  for (int i = 0; i < 16; ++i) Hm.m[i] = p->H_world_map[i];
  WorldArgs wa;
  wa.pts = pts;
  wa.conf = conf;
  wa.emb_ok = emb_ok;
  wa.pw = ws->pw.as<float4>();
  wa.n_px = (uint32_t)n_px;
  wa.H = (uint32_t)p->H;
  wa.W = (uint32_t)p->W;
  wa.stride = (uint32_t)p->stride;
  wa.thr = p->conf_threshold;
  wa.hist0 = hist;
  const bool vec4 = aligned16(pts) && aligned16(conf) && n_px >= 4;
  const int wgrid = grid_for(vec4 ? cdiv(n_px, 4) : n_px, 256);
  const float q0 = (float)p->bbox_lo_pct / 100.0f;  // numpy: q / float32(100) in float32
  const float q1 = (float)p->bbox_hi_pct / 100.0f;
  BracketArgs br{};
  if (bracket) {
    static const cudaError_t smem_opt_in = cudaFuncSetAttribute(bracket_resolve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                                (int)(kResolveStage * sizeof(uint32_t)));
    VSM_CUDA(smem_opt_in);
    br.bs = reinterpret_cast<BracketState*>(ws->sel_bracket.p);
    br.lists = reinterpret_cast<float*>(reinterpret_cast<char*>(ws->sel_bracket.p) + 256);
    br.cap = bracket_list_cap(n_px);
    br.und = reinterpret_cast<uint32_t*>(br.lists + (size_t)kBrLists * br.cap);
    br.und_cap = bracket_und_cap(n_px);
    VSM_CUDA(cudaMemsetAsync(br.bs, 0, sizeof(BracketState), s));
    if (deferred) VSM_CUDA(cudaMemsetAsync(&br.bs->deferred, 1, 1, s));  // (little-endian: the word becomes 1)
    bracket_sample_kernel<<<kBrLists, 1024, 0, s>>>(wa, Hm, br.bs, q0, q1);
    VSM_LAUNCHED();
  }
  const int wmode = (!filters || deferred) ? 0 : (bracket ? 2 : 1);
#define VSM_WORLD(VEC4_)                                                            \
  do {                                                                              \
    if (wmode == 0)                                                                 \
      world_points_kernel<VEC4_, 0><<<wgrid, 256, 0, s>>>(wa, Hm, ctr, br);         \
    else if (wmode == 1)                                                            \
      world_points_kernel<VEC4_, 1><<<wgrid, 256, 0, s>>>(wa, Hm, ctr, br);         \
    else                                                                            \
      world_points_kernel<VEC4_, 2><<<wgrid, 256, 0, s>>>(wa, Hm, ctr, br);         \
  } while (0)
  if (vec4)
    VSM_WORLD(true);
  else
    VSM_WORLD(false);
#undef VSM_WORLD
  VSM_LAUNCHED();

  // ---- filters and the two submap-local tables -----------------------------------------------------------
  const bool patch = (variant & 1) != 0;
  static_assert(kIrrCap == (1u << 16) && sizeof(IrrEntry) == 32, "irregular list: 64 K entries of 32 bytes");
  FilterArgs fa;
  fa.pw = ws->pw.as<float4>();
  fa.pt_slot = ws->pt_slot.as<int32_t>();
  fa.frame_base = (uint32_t)p->frame_base;
  fa.min_pts = (uint32_t)std::max(p->coarse_min_points, 0);
  fa.opts = ((variant & 2) ? kOptMaskProbe : 0) | (g_range_policy.load() == 1 ? kOptDropRange : 0);
  const PixMap pm = make_pixmap(n_px, p->end_idx, p->H, p->W, patch);
  const int grid = grid_for((int64_t)pm.n_items * 32, 256);
  call.v7 = v7;
  Prep7Args a7{};
  if (v7) {
    a7.pw = ws->pw.as<float4>();
    a7.pt2 = ws->pt2.as<uint2>();
    a7.irr = ws->irr.as<IrrEntry>();
    a7.frame_base = (uint32_t)p->frame_base;
    a7.cell = m->vs_f;
    a7.cell_coarse = (float)(m->cfg.voxel_size * cf);
    a7.inv_coarse = 1.0f / a7.cell_coarse;
    a7.cm.div = make_fastdiv((uint32_t)cf);
    a7.cm.add = ((uint32_t)cf - 1u) << 20;
    a7.min_pts = fa.min_pts;
    a7.opts = fa.opts;
  }
  if (filters) {
    if (deferred) {
      // nothing here: the insert kernel applies and fills the brackets, the exact percentiles follow it
    } else if (bracket) {
      bracket_resolve_kernel<<<kBrLists, 1024, kResolveStage * sizeof(uint32_t), s>>>(br.bs, br.lists, br.cap, &ctr->n_finite, q0, q1, ctr->bounds,
                                                       &ctr->sel_miss, br.und_cap);
      VSM_LAUNCHED();
    } else if (variant & 4) {
      VSM_TRY(run_percentiles_world_fast(sst, hist, ws->pw.as<float4>(), n_px, PF_SEL | PF_FINITE, q0, q1, ctr->bounds,
                                         &ctr->n_finite, s));
    } else {
      SelSrc src;
      src.base = reinterpret_cast<const float*>(ws->pw.p);
      src.stride = 4;
      src.ncol = 3;
      src.flag_off = 3;
      src.flag_need = PF_SEL | PF_FINITE;
      src.n_items = n_px;
      VSM_TRY(run_percentiles_after_hist0(sst, hist, src, 2, q0, q1, ctr->bounds, &ctr->n_finite, s));
    }
    if (deferred) {
      if (patch)
        insert7d_kernel<true><<<grid, 256, 0, s>>>(a7, pm, tb, ctr, br);
      else
        insert7d_kernel<false><<<grid, 256, 0, s>>>(a7, pm, tb, ctr, br);
      VSM_LAUNCHED();
      bracket_resolve_kernel<<<kBrLists, 1024, kResolveStage * sizeof(uint32_t), s>>>(br.bs, br.lists, br.cap, &ctr->n_finite, q0, q1, ctr->bounds,
                                                       &ctr->sel_miss, br.und_cap);
      VSM_LAUNCHED();
      insert7u_kernel<<<grid_for((int64_t)br.und_cap, 256, sm_count() * 8), 256, 0, s>>>(a7, pm, tb, ctr, br);
      VSM_LAUNCHED();
    } else if (v7) {
      if (patch)
        insert7_kernel<true><<<grid, 256, 0, s>>>(a7, pm, tb, ctr);
      else
        insert7_kernel<false><<<grid, 256, 0, s>>>(a7, pm, tb, ctr);
      VSM_LAUNCHED();
    } else {
      fa.cell = (float)(m->cfg.voxel_size * p->coarse_factor);  // float(voxel_size) * 3.0, weak scalar -> float32
      if (patch)
        bbox_coarse_kernel<true><<<grid, 256, 0, s>>>(fa, pm, ta, ctr);
      else
        bbox_coarse_kernel<false><<<grid, 256, 0, s>>>(fa, pm, ta, ctr);
      VSM_LAUNCHED();
      fa.cell = m->vs_f;
      if (patch)
        fine_insert_kernel<true, true><<<grid, 256, 0, s>>>(fa, pm, ta, tb, ctr);
      else
        fine_insert_kernel<true, false><<<grid, 256, 0, s>>>(fa, pm, ta, tb, ctr);
      VSM_LAUNCHED();
    }
  } else {
    fa.cell = m->vs_f;
    if (patch)
      fine_insert_kernel<false, true><<<grid, 256, 0, s>>>(fa, pm, ta, tb, ctr);
    else
      fine_insert_kernel<false, false><<<grid, 256, 0, s>>>(fa, pm, ta, tb, ctr);
    VSM_LAUNCHED();
  }
  const int vgrid = grid_for((int64_t)std::min<uint64_t>(n_sel_max, (uint64_t)sm_count() * 8 * 256), 256, sm_count() * 8);
  if (v7) {
    coarse7_kernel<<<vgrid, 256, 0, s>>>(a7, tb, ta, ctr);
    VSM_LAUNCHED();
    keep7_kernel<<<vgrid, 256, 0, s>>>(tb, ta, global_store(m), ctr, ws->kept.as<uint32_t>(), a7.min_pts,
                                       m->d_n_vox.as<uint32_t>(), (uint32_t)m->vcap,
                                       (uint32_t)std::min<int64_t>(m->log_cap, 0xFFFFFFFFll),
                                       (uint32_t)std::min<uint64_t>(n_sel_max, 0xFFFFFFFFull));
    VSM_LAUNCHED();
    compact_merge7_kernel<<<vgrid, 256, 0, s>>>(tb, ws->kept.as<uint32_t>(), global_store(m), ctr, m->log_gid.as<int32_t>(),
                                                m->log_fuse.as<int32_t>(), m->log_mask.as<unsigned long long>(),
                                                p->submap_id);
    VSM_LAUNCHED();
  } else {
    count_new_decide_kernel<<<vgrid, 256, 0, s>>>(tb, global_store(m), ctr, m->d_n_vox.as<uint32_t>(), (uint32_t)m->vcap,
                                                  (uint32_t)std::min<int64_t>(m->log_cap, 0xFFFFFFFFll),
                                                  (uint32_t)std::min<uint64_t>(n_sel_max, 0xFFFFFFFFull));
    VSM_LAUNCHED();

    // ---- distinct voxels -> global map; points -> sorted entries ---------------------------------------------
    compact_merge_kernel<<<vgrid, 256, 0, s>>>(tb, global_store(m), ctr, ws->lv_off.as<uint32_t>(),
                                               ws->lv_cursor.as<uint32_t>(), ws->lv_gid.as<int32_t>(),
                                               m->log_gid.as<int32_t>(), m->log_fuse.as<int32_t>(),
                                               m->log_mask.as<unsigned long long>(), p->submap_id);
    VSM_LAUNCHED();
  }

  FuseRecord& rec = m->fuses[call.fuse_index];
  int32_t* point_gid = nullptr;
  if (keep_index) {
    VSM_TRY(rec.point_gid.ensure((size_t)p->S * px_per_frame * 4, s));
    point_gid = rec.point_gid.as<int32_t>();
    VSM_CUDA(cudaMemsetAsync(point_gid, 0xFF, (size_t)p->S * px_per_frame * 4, s));
  } else if (pixel_order) {
    VSM_TRY(ws->sorted_gid.ensure((size_t)n_px * 4, s, 0, 1.1));  // per-pixel voxel ids
    point_gid = ws->sorted_gid.as<int32_t>();
  }

  AccArgs aa{};
  aa.emb_index = emb_index;
  aa.row_bytes = row_bytes;
  aa.vsum = m->vsum.as<float>();
  aa.d = m->d;
  aa.nvec = (int)(row_bytes / 16);
  aa.ctr = ctr;
  const int grid7 = grid_for(n_px, 256);
  if (!pixel_order) {
    if (v7)
      scatter7_kernel<<<grid7, 256, 0, s>>>(ws->pt2.as<uint2>(), ws->pw.as<float4>(), (uint32_t)n_px, tb, check ? 1 : 0,
                                            ws->sorted_pix[ab].as<unsigned long long>(), point_gid, ctr);
    else if (patch)
      scatter_kernel<true><<<grid, 256, 0, s>>>(ws->pt_slot.as<int32_t>(), ws->pw.as<float4>(), pm, tb,
                                                ws->lv_off.as<uint32_t>(), ws->lv_cursor.as<uint32_t>(),
                                                ws->lv_gid.as<int32_t>(), check ? 1 : 0,
                                                ws->sorted_pix[ab].as<unsigned long long>(), point_gid, ctr);
    else
      scatter_kernel<false><<<grid, 256, 0, s>>>(ws->pt_slot.as<int32_t>(), ws->pw.as<float4>(), pm, tb,
                                                 ws->lv_off.as<uint32_t>(), ws->lv_cursor.as<uint32_t>(),
                                                 ws->lv_gid.as<int32_t>(), check ? 1 : 0,
                                                 ws->sorted_pix[ab].as<unsigned long long>(), point_gid, ctr);
    VSM_LAUNCHED();
    tables_cleanup_kernel<<<vgrid, 256, 0, s>>>(ta, filters ? 1 : 0, tb);
    VSM_LAUNCHED();
    aa.emb = emb_dev;
    aa.pix_base = 0;
    aa.entries = ws->sorted_pix[ab].as<unsigned long long>();
    // the accumulate kernel goes to the side stream: the caller's stream is free for the next call's preparation
    cudaStream_t sa = green ? ws->green_stream[1] : (ws->overlap ? ws->acc_stream : s);
    if (ws->overlap) {
      VSM_CUDA(cudaEventRecord(ws->ev_prep_done[ab], s));
      VSM_CUDA(cudaStreamWaitEvent(sa, ws->ev_prep_done[ab], 0));
    }
    if (call.profiled) VSM_CUDA(cudaEventRecord(ev[1], sa));
    VSM_TRY(launch_accumulate(aa, bf16, true, check, sa, green ? ws->green_sms[1] : 0));
    if (call.profiled) VSM_CUDA(cudaEventRecord(ev[2], sa));
    if (ws->overlap) {
      VSM_CUDA(cudaEventRecord(ws->ev_acc_done[ab], sa));
      ws->acc_used[ab] = true;
      ws->acc_parity ^= 1;
    }
  } else {
    if (v7)
      scatter7_kernel<<<grid7, 256, 0, s>>>(ws->pt2.as<uint2>(), ws->pw.as<float4>(), (uint32_t)n_px, tb, check ? 1 : 0,
                                            nullptr, point_gid, ctr);
    else
      point_gid_kernel<<<grid_for(n_px, 256), 256, 0, s>>>(ws->pt_slot.as<int32_t>(), ws->pw.as<float4>(), (uint32_t)n_px, tb,
                                                           ws->lv_gid.as<int32_t>(), check ? 1 : 0, point_gid, ctr);
    VSM_LAUNCHED();
    tables_cleanup_kernel<<<vgrid, 256, 0, s>>>(ta, filters ? 1 : 0, tb);
    VSM_LAUNCHED();
    aa.point_gid = point_gid;
    if (host == nullptr) {
      aa.emb = emb_dev;
      aa.pix_base = 0;
      aa.n = n_px;
      if (call.profiled) VSM_CUDA(cudaEventRecord(ev[1], s));
      VSM_TRY(launch_accumulate(aa, bf16, false, check, s));
      if (call.profiled) VSM_CUDA(cudaEventRecord(ev[2], s));
    } else if (host->emb_mapped != nullptr) {
      // zero-copy: one accumulate over all pixels, rows fetched from mapped host memory by the kernel itself
      aa.emb = host->emb_mapped;
      aa.pix_base = 0;
      aa.n = n_px;
      VSM_TRY(launch_accumulate(aa, bf16, false, check, s));
    } else {
      // stream the embeddings frame by frame through two device buffers
      const size_t chunk_bytes = (size_t)px_per_frame * row_bytes;
      VSM_TRY(ensure_stream_objects(m, chunk_bytes));
      for (int f = 0; f < p->end_idx; ++f) {
        const int b = f & 1;
        if (f >= 2) VSM_CUDA(cudaStreamWaitEvent(m->copy_stream, m->ev_stage[b], 0));
        VSM_CUDA(cudaMemcpyAsync(m->stage_emb[b].p, host->emb_host + (size_t)f * chunk_bytes, chunk_bytes,
                                 cudaMemcpyHostToDevice, m->copy_stream));
        VSM_CUDA(cudaEventRecord(m->ev_copy[b], m->copy_stream));
        VSM_CUDA(cudaStreamWaitEvent(s, m->ev_copy[b], 0));
        aa.emb = m->stage_emb[b].as<uint8_t>();
        aa.pix_base = (int64_t)f * px_per_frame;
        aa.n = px_per_frame;
        VSM_TRY(launch_accumulate(aa, bf16, false, check, s));
        VSM_CUDA(cudaEventRecord(m->ev_stage[b], s));
      }
    }
  }
  return VSM_OK;
}

// Make room before queueing a call.  The device aborts a call that does not fit (and the collect repeats it after
// growing), so this is only a heuristic that keeps aborts rare: voxel capacity is raised while nothing is queued;
// the contributor log (exactly one entry per call and voxel) is reserved many calls ahead, sized by the last
// call's voxel count on this map or, for a fresh map, on this device.
static int pregrow(vsm_map* m, cudaStream_t s) {
  const int64_t occ = std::max<int64_t>(m->last_n_occ, m->ws->hint_n_occ);
  const int64_t per_call = occ + occ / 4 + 1024;
  if (m->pending.empty() && m->n_vox + per_call > m->vcap) VSM_TRY(map_grow(m, m->n_vox + 2 * per_call, s));
  const int64_t in_flight = (int64_t)m->pending.size() + 1;
  if (m->log_n + in_flight * 2 * per_call > m->log_cap) {
    if (!m->pending.empty()) {
      ++g_early_collects;
      VSM_TRY(fuse_collect_locked(m, s, &m->stats_backlog));
    }
    VSM_TRY(log_grow(m, m->log_n + 32 * 2 * per_call, s));
  }
  return VSM_OK;
}

static int fuse_submit_locked(vsm_map* m, const float* pts, const float* conf, const uint8_t* emb, const uint8_t* emb_ok,
                              const HostEmb* host, const vsm_fuse_params* p, cudaStream_t s) {
  if ((int)m->pending.size() >= kCallRing) {
    ++g_early_collects;
    VSM_TRY(fuse_collect_locked(m, s, &m->stats_backlog));
  }
  VSM_TRY(pregrow(m, s));
  FuseRecord rec{};
  rec.submap_id = p->submap_id;
  rec.S = p->S;
  rec.H = p->H;
  rec.W = p->W;
  rec.end_idx = p->end_idx;
  rec.stride = p->stride;
  m->fuses.push_back(rec);
  PendingCall call{};
  call.p = *p;
  call.pts = pts;
  call.conf = conf;
  call.emb = emb;
  call.emb_ok = emb_ok;
  call.slot = (int)m->pending.size();
  call.fuse_index = (int)m->fuses.size() - 1;
  m->pending.push_back(call);
  const int st = fuse_enqueue(m, m->pending.back(), host, s);
  if (st != VSM_OK) {
    cudaStreamSynchronize(s);
    m->pending.back().precheck_mask.release();
    m->pending.pop_back();
    m->fuses.back().point_gid.release();
    m->fuses.pop_back();
  }
  return st;
}

// Synchronise once, read every queued call's counters, repeat the calls the device aborted for lack of room.
// Stats are appended to `out` in call order.  The first failing call's status is returned (after all calls
// have been accounted for); m->last_stats holds the stats of the last call.
static int fuse_collect_locked(vsm_map* m, cudaStream_t s, std::vector<vsm_fuse_stats>* out) {
  if (m->pending.empty()) return VSM_OK;
  const size_t n_calls = m->pending.size();
  VSM_TRY(join_accumulates(m->ws, s));
  std::vector<FuseCounters> hc(kCallRing);
  VSM_TRY(read_back(m, hc.data(), m->ctr_ring.p, sizeof(FuseCounters) * n_calls, s));
  std::vector<PendingCall> calls;
  calls.swap(m->pending);
  int first_error = VSM_OK;
  char first_msg[512] = "";
  // profiling: calls overlap (accumulate of call i runs beside the preparation of call i+1), so the time of the
  // batch is the span from the first call's start to the last call's end, not a sum over calls
  cudaEvent_t span_begin = nullptr, span_end = nullptr;
  bool any_abort = false;
  for (size_t k = 0; k < n_calls; ++k) any_abort |= hc[k].abort != 0;
  for (size_t k = 0; k < n_calls; ++k) {
    PendingCall& call = calls[k];
    FuseCounters c = hc[k];
    const bool filters = (call.p.flags & VSM_FUSE_FILTERS) != 0;
    int status = VSM_OK;
    for (int attempt = 0; c.abort && (!c.internal_err || c.tbl_overflow) && !c.range_err && !c.bad_index; ++attempt) {
      // nothing was modified: grow and run the call again, alone
      if (attempt >= 3) {
        set_error("internal: fuse call kept aborting after the map was grown");
        status = VSM_E_INTERNAL;
        break;
      }
      uint32_t state[2] = {0, 0};
      VSM_TRY(read_back(m, state, m->d_n_vox.p, sizeof(state), s));
      m->n_vox = state[0];
      m->log_n = state[1];
      if (c.sel_miss) {
        // the one-pass percentile select could not answer (its counts are meaningless): radix select this time
        call.p.flags |= kFuseForceRadix;
        ++g_select_misses;
      } else if (c.tbl_overflow) {
        call.p.flags |= kFuseForceBigTables;  // more voxels than the table prefix sized from earlier calls holds
        ++g_table_retries;
      } else if (c.irr_overflow) {
        call.p.flags |= kFuseForceTwoTables;  // more irregular points than the side list holds: two-table preparation
      } else {
        ++g_capacity_retries;
        VSM_TRY(map_grow(m, m->n_vox + (int64_t)c.n_new, s));
        VSM_TRY(log_grow(m, m->log_n + (int64_t)(call.v7 ? c.n_kept : c.n_occ_b), s));
      }
      call.slot = 0;
      m->fuses[call.fuse_index].point_gid.release();
      VSM_TRY(fuse_enqueue(m, call, nullptr, s));
      VSM_TRY(join_accumulates(m->ws, s));
      VSM_TRY(read_back(m, &c, m->ctr_ring.p, sizeof(FuseCounters), s));
    }
    call.precheck_mask.release();
    vsm_fuse_stats st{};
    st.n_conf = (int64_t)c.n_conf;
    st.n_finite = (int64_t)c.n_finite;
    st.n_bbox = filters ? (int64_t)c.n_bbox : (int64_t)c.n_conf;
    st.n_fused = (int64_t)c.n_fused;
    const uint32_t n_local = call.v7 ? c.n_kept : c.n_occ_b;  // local table entries merged into the map (log entries)
    st.n_submap_voxels = call.v7 ? c.n_distinct : c.n_occ_b;
    st.n_bad_emb_rows = (int64_t)c.n_bad_emb;
    st.n_range_dropped = (int64_t)c.range_dropped;
    for (int i = 0; i < 3; ++i) {
      st.bbox_lo[i] = filters ? c.bounds[2 * i] : __builtin_nanf("");
      st.bbox_hi[i] = filters ? c.bounds[2 * i + 1] : __builtin_nanf("");
    }
    if (status == VSM_OK) {
      if (c.internal_err) {
        set_error("internal: hash probe limit / overflow (%u)", c.internal_err);
        status = VSM_E_INTERNAL;
      } else if (c.range_err) {
        set_error("%u points have a finite voxel coordinate outside +-(2^20-1) cells", c.range_err);
        status = VSM_E_COORD_RANGE;
      } else if (c.bad_index) {
        set_error("%u confident pixels carry an embedding index outside the table of %d rows", c.bad_index,
                  call.p.emb_rows);
        status = VSM_E_INVALID;
      } else if (c.n_bad_emb) {
        set_error("%llu non-finite embedding rows / voxel sums met in the optimistic filter pass; clear the map "
                  "and fuse again with VSM_FUSE_EMB_PRECHECK", (unsigned long long)c.n_bad_emb);
        status = VSM_E_NONFINITE_EMB;
      }
    }
    if (status == VSM_OK || status == VSM_E_NONFINITE_EMB) {
      m->last_n_occ = n_local;
      if (m->ws->hint_vs != m->cfg.voxel_size) {  // the hint follows the voxel size of the calls
        m->ws->hint_n_occ = 0;
        m->ws->hint_vs = m->cfg.voxel_size;
      }
      m->ws->hint_n_occ = std::max<int64_t>(m->ws->hint_n_occ, n_local);  // largest call seen on this device
      m->fuses[call.fuse_index].n_fused = (int64_t)c.n_fused;
      if (call.profiled && c.n_fused > 0 && !hc[k].abort) {
        float t_acc = 0.f;
        cudaEvent_t* ev = m->ev_ring[call.slot];
        VSM_CUDA(cudaEventElapsedTime(&t_acc, ev[1], ev[2]));
        if (!span_begin) span_begin = ev[0];
        span_end = ev[2];
        static const bool trace = getenv("VSM_TRACE") && getenv("VSM_TRACE")[0] == '1';  // timeline of every profiled call
        if (trace) {
          float a0 = 0.f, a1 = 0.f, a2 = 0.f;
          cudaEventElapsedTime(&a0, span_begin, ev[0]);
          cudaEventElapsedTime(&a1, span_begin, ev[1]);
          cudaEventElapsedTime(&a2, span_begin, ev[2]);
          fprintf(stderr, "[vsm trace] call %zu: start %.3f  acc %.3f .. %.3f ms\n", k, a0, a1, a2);
        }
        m->prof.accumulate_ms += t_acc;
        m->prof.fuse_calls += 1;
        m->prof.accumulate_launches += 1;
        m->prof.accumulate_bytes += (int64_t)c.n_fused * m->d * m->esize + (int64_t)n_local * m->d * 4;
        m->prof.points_fused += (int64_t)c.n_fused;
      }
    }
    if (status != VSM_OK && first_error == VSM_OK) {
      first_error = status;
      snprintf(first_msg, sizeof(first_msg), "%s", vsm_last_error());
    }
    m->last_stats = st;
    if (out) out->push_back(st);
  }
  if (span_begin && span_end && !any_abort) {
    float t_all = 0.f;
    VSM_CUDA(cudaEventElapsedTime(&t_all, span_begin, span_end));
    m->prof.fuse_ms += t_all;
  }
  uint32_t state[2] = {0, 0};
  VSM_TRY(read_back(m, state, m->d_n_vox.p, sizeof(state), s));
  m->n_vox = state[0];
  m->log_n = state[1];
  m->last_stats.n_map_voxels = m->n_vox;
  if (out && !out->empty()) out->back().n_map_voxels = m->n_vox;
  if (first_error != VSM_OK) set_error("%s", first_msg);
  return first_error;
}

int fuse_collect_pending(vsm_map* m, cudaStream_t s) {
  if (m->pending.empty()) return VSM_OK;
  WsLease ws_lock(m->ws, s);
  VSM_TRY(ws_lock.status());
  return fuse_collect_locked(m, s, nullptr);
}

}  // namespace vsm

using namespace vsm;

extern "C" int vsm_fuse_submap_async(vsm_map* m, const float* pts_dev, const float* conf_dev, const void* emb_dev,
                                     const uint8_t* emb_ok_dev, const vsm_fuse_params* p, void* stream) {
  VSM_TRY(validate_params(m, p));
  if (!pts_dev || !conf_dev || !emb_dev) {
    set_error("null device pointer");
    return VSM_E_INVALID;
  }
  if (p->flags & VSM_FUSE_KEEP_POINT_INDEX) {
    // keeps a per-call array alive in the record: fine, but nothing to gain from queueing
  }
  VSM_CUDA(cudaSetDevice(m->device));
  WsLease ws_lock(m->ws, (cudaStream_t)stream);
  VSM_TRY(ws_lock.status());
  return fuse_submit_locked(m, pts_dev, conf_dev, (const uint8_t*)emb_dev, emb_ok_dev, nullptr, p,
                            (cudaStream_t)stream);
}

extern "C" int vsm_fuse_collect(vsm_map* m, vsm_fuse_stats* stats_host, int32_t max_stats, int32_t* n_stats_host,
                                void* stream) {
  if (!m) {
    set_error("null map");
    return VSM_E_INVALID;
  }
  VSM_CUDA(cudaSetDevice(m->device));
  WsLease ws_lock(m->ws, (cudaStream_t)stream);
  VSM_TRY(ws_lock.status());
  // calls that an earlier submit had to collect itself (full ring, contributor-log growth) come first
  std::vector<vsm_fuse_stats> out;
  out.swap(m->stats_backlog);
  const int st = fuse_collect_locked(m, (cudaStream_t)stream, &out);
  const int n = (int)std::min<size_t>(out.size(), (size_t)std::max(max_stats, 0));
  if (stats_host)
    for (int i = 0; i < n; ++i) stats_host[i] = out[i];
  if (n_stats_host) *n_stats_host = (int32_t)out.size();
  return st;
}

extern "C" int vsm_fuse_submap(vsm_map* m, const float* pts_dev, const float* conf_dev, const void* emb_dev,
                               const uint8_t* emb_ok_dev, const vsm_fuse_params* p, vsm_fuse_stats* stats_host,
                               void* stream) {
  VSM_TRY(validate_params(m, p));
  if (!pts_dev || !conf_dev || !emb_dev) {
    set_error("null device pointer");
    return VSM_E_INVALID;
  }
  VSM_CUDA(cudaSetDevice(m->device));
  cudaStream_t s = (cudaStream_t)stream;
  WsLease ws_lock(m->ws, s);
  VSM_TRY(ws_lock.status());
  VSM_TRY(fuse_collect_locked(m, s, &m->stats_backlog));
  VSM_TRY(fuse_submit_locked(m, pts_dev, conf_dev, (const uint8_t*)emb_dev, emb_ok_dev, nullptr, p, s));
  const int st = fuse_collect_locked(m, s, nullptr);
  if (stats_host) *stats_host = m->last_stats;
  return st;
}

extern "C" int vsm_fuse_submap_host(vsm_map* m, const float* pts_host, const float* conf_host, const void* emb_host,
                                    const vsm_fuse_params* p, vsm_fuse_stats* stats_host, void* stream) {
  VSM_TRY(validate_params(m, p));
  if (!pts_host || !conf_host || !emb_host) {
    set_error("null host pointer");
    return VSM_E_INVALID;
  }
  if (p->emb_index_dev != nullptr) {
    set_error("vsm_fuse_submap_host streams dense embeddings; copy an index image and its table to the device "
              "(a few MB) and use vsm_fuse_submap");
    return VSM_E_INVALID;
  }
  VSM_CUDA(cudaSetDevice(m->device));
  cudaStream_t s = (cudaStream_t)stream;
  WsLease ws_lock(m->ws, s);
  VSM_TRY(ws_lock.status());
  VSM_TRY(fuse_collect_locked(m, s, &m->stats_backlog));
  const size_t n_px = (size_t)p->end_idx * p->H * p->W;
  VSM_TRY(m->stage_pts.ensure(std::max<size_t>(n_px * 12, 16), s));
  VSM_TRY(m->stage_conf.ensure(std::max<size_t>(n_px * 4, 16), s));
  if (n_px) {
    VSM_CUDA(cudaMemcpyAsync(m->stage_pts.p, pts_host, n_px * 12, cudaMemcpyHostToDevice, s));
    VSM_CUDA(cudaMemcpyAsync(m->stage_conf.p, conf_host, n_px * 4, cudaMemcpyHostToDevice, s));
  }
  HostEmb he;
  he.emb_host = (const uint8_t*)emb_host;
  if (g_host_zero_copy.load()) {
    cudaPointerAttributes at{};
    if (cudaPointerGetAttributes(&at, emb_host) == cudaSuccess && at.type == cudaMemoryTypeHost && at.devicePointer != nullptr)
      he.emb_mapped = (const uint8_t*)at.devicePointer;
    else
      cudaGetLastError();  // pageable memory: staged copies
  }
  vsm_fuse_params q = *p;
  q.flags &= ~VSM_FUSE_EMB_PRECHECK;  // not available when streaming from the host
  // a call the device aborts for lack of room is repeated by the collect with device-resident inputs only, which
  // a host-streamed call does not have: make room up front instead (worst case: every selected pixel a new voxel
  // is far too pessimistic, so use the same guess as the device path and fall back to an explicit retry here)
  for (int attempt = 0; attempt < 3; ++attempt) {
    VSM_TRY(pregrow(m, s));
    FuseRecord rec{};
    rec.submap_id = q.submap_id;
    rec.S = q.S;
    rec.H = q.H;
    rec.W = q.W;
    rec.end_idx = q.end_idx;
    rec.stride = q.stride;
    m->fuses.push_back(rec);
    PendingCall call{};
    call.p = q;
    call.pts = m->stage_pts.as<float>();
    call.conf = m->stage_conf.as<float>();
    call.slot = 0;
    call.fuse_index = (int)m->fuses.size() - 1;
    VSM_TRY(fuse_enqueue(m, call, &he, s));
    FuseCounters c{};
    VSM_TRY(read_back(m, &c, m->ctr_ring.p, sizeof(FuseCounters), s));
    if (c.abort && (!c.internal_err || c.tbl_overflow) && !c.range_err) {
      m->fuses.back().point_gid.release();
      m->fuses.pop_back();
      if (c.sel_miss) {
        q.flags |= kFuseForceRadix;
        ++g_select_misses;
        --attempt;  // not a capacity retry
        continue;
      }
      if (c.tbl_overflow) {
        q.flags |= kFuseForceBigTables;
        ++g_table_retries;
        --attempt;
        continue;
      }
      if (c.irr_overflow) {
        q.flags |= kFuseForceTwoTables;
        --attempt;
        continue;
      }
      uint32_t state[2] = {0, 0};
      VSM_TRY(read_back(m, state, m->d_n_vox.p, sizeof(state), s));
      VSM_TRY(map_grow(m, (int64_t)state[0] + (int64_t)c.n_new, s));
      VSM_TRY(log_grow(m, (int64_t)state[1] + (int64_t)(call.v7 ? c.n_kept : c.n_occ_b), s));
      m->last_n_occ = call.v7 ? c.n_kept : c.n_occ_b;
      continue;
    }
    // account for it through the common path
    m->pending.push_back(call);
    // counters are already on the host; collect re-reads them (cheap) and fills stats / errors
    const int st = fuse_collect_locked(m, s, nullptr);
    if (stats_host) *stats_host = m->last_stats;
    return st;
  }
  set_error("internal: host-streamed fuse call kept aborting after the map was grown");
  return VSM_E_INTERNAL;
}

extern "C" int vsm_profile_enable(vsm_map* m, int on) {
  if (!m) {
    set_error("null map");
    return VSM_E_INVALID;
  }
  m->profiling = on != 0;
  m->prof = vsm_profile{};
  return VSM_OK;
}

extern "C" int vsm_profile_get(const vsm_map* m, vsm_profile* out_host) {
  if (!m || !out_host) {
    set_error("null argument");
    return VSM_E_INVALID;
  }
  *out_host = m->prof;
  return VSM_OK;
}

extern "C" int vsm_embedding_row_mask(const vsm_map* m, const float* conf_dev, const void* emb_dev,
                                      const vsm_fuse_params* p, uint8_t* out_mask_dev, void* stream) {
  VSM_TRY(validate_params(m, p));
  if (!conf_dev || !emb_dev || !out_mask_dev) {
    set_error("null device pointer");
    return VSM_E_INVALID;
  }
  VSM_CUDA(cudaSetDevice(m->device));
  cudaStream_t s = (cudaStream_t)stream;
  const int64_t n_px = (int64_t)p->end_idx * p->H * p->W;
  if (n_px == 0) return VSM_OK;
  const int64_t row_bytes = (int64_t)m->d * m->esize;
  const int nvec = (int)(row_bytes / 16);
  if (m->cfg.emb_dtype == VSM_BF16)
    emb_row_mask_kernel<true><<<sm_count() * 8, 256, 0, s>>>(conf_dev, (const uint8_t*)emb_dev, p->emb_index_dev, p->emb_rows, row_bytes, nvec, n_px, p->H,
                                                      p->W, p->stride, p->conf_threshold, out_mask_dev);
  else
    emb_row_mask_kernel<false><<<sm_count() * 8, 256, 0, s>>>(conf_dev, (const uint8_t*)emb_dev, p->emb_index_dev, p->emb_rows, row_bytes, nvec, n_px, p->H,
                                                       p->W, p->stride, p->conf_threshold, out_mask_dev);
  VSM_LAUNCHED();
  return VSM_OK;
}

extern "C" int vsm_set_option(const char* key, int64_t value) {
  if (!key) {
    set_error("vsm_set_option: null key");
    return VSM_E_INVALID;
  }
  if (!strcmp(key, "select_mode") && value >= 0 && value <= 3) {
    g_select_mode = (int)value;
    return VSM_OK;
  }
  if (!strcmp(key, "prep_variant") && value >= 0 && value <= 15) {
    g_prep_variant = (int)value;
    return VSM_OK;
  }
  if (!strcmp(key, "coord_range_policy") && (value == 0 || value == 1)) {
    g_range_policy = (int)value;
    return VSM_OK;
  }
  if (!strcmp(key, "query_shadow") && (value == 0 || value == 1)) {
    g_query_shadow = (int)value;
    return VSM_OK;
  }
  if (!strcmp(key, "acc_tmarun") && (value == 0 || value == 1)) {
    g_acc_tmarun = (int)value;
    return VSM_OK;
  }
  if (!strcmp(key, "acc_mix") && value >= 0 && value <= 1000) {
    g_acc_mix = (int)value;
    return VSM_OK;
  }
  if (!strcmp(key, "acc_mix_ctas") && value >= 1 && value <= 3) {
    g_acc_mix_ctas = (int)value;
    return VSM_OK;
  }
  if (!strcmp(key, "acc_tma") && (value == 0 || value == 1)) {
    g_acc_tma = (int)value;
    return VSM_OK;
  }
  if (!strcmp(key, "acc_ring") && (value == 0 || value == 1)) {
    g_acc_ring = (int)value;
    return VSM_OK;
  }
  if (!strcmp(key, "small_tables") && (value == 0 || value == 1)) {
    g_small_tables = (int)value;
    return VSM_OK;
  }
  if (!strcmp(key, "table_hint")) {  // forget (0) or set the voxel count the table prefix is sized from (tests)
    int dev = 0;
    VSM_CUDA(cudaGetDevice(&dev));
    Workspace* ws = workspace_for_device(dev);
    if (!ws || value < 0) {
      set_error("vsm_set_option: table_hint needs a device workspace and a value >= 0");
      return VSM_E_INVALID;
    }
    std::lock_guard<std::mutex> lock(ws->mu);
    ws->hint_n_occ = value;
    ws->hint_vs = 0.0;
    return VSM_OK;
  }
  if (!strcmp(key, "host_zero_copy") && (value == 0 || value == 1)) {
    g_host_zero_copy = (int)value;
    return VSM_OK;
  }
  if (!strcmp(key, "acc_ctas_per_sm") && value >= 1 && value <= 8) {
    g_acc_ctas_per_sm = (int)value;
    return VSM_OK;
  }
  if (!strcmp(key, "green_prep_sms") && value >= 0 && value <= 1024) {
    // value > 0: partition the SMs (>= value for the preparation kernels, the rest for the accumulate kernel) and turn
    // the overlap on; 0: back to one stream on the whole device
    int dev = 0;
    VSM_CUDA(cudaGetDevice(&dev));
    Workspace* ws = workspace_for_device(dev);
    if (!ws) {
      set_error("vsm_set_option: no workspace for device %d", dev);
      return VSM_E_INVALID;
    }
    std::lock_guard<std::mutex> lock(ws->mu);
    VSM_TRY(ensure_acc_stream(ws));
    VSM_CUDA(cudaDeviceSynchronize());
    VSM_TRY(green_setup(ws, dev, (int)value));
    ws->overlap = value != 0;
    return VSM_OK;
  }
  if (!strcmp(key, "overlap") && (value == 0 || value == 1)) {
    int dev = 0;
    VSM_CUDA(cudaGetDevice(&dev));
    Workspace* ws = workspace_for_device(dev);
    if (!ws) {
      set_error("vsm_set_option: no workspace for device %d", dev);
      return VSM_E_INVALID;
    }
    std::lock_guard<std::mutex> lock(ws->mu);
    VSM_TRY(ensure_acc_stream(ws));
    VSM_CUDA(cudaDeviceSynchronize());
    ws->overlap = value != 0;
    return VSM_OK;
  }
  set_error("vsm_set_option: unknown key or bad value: %s = %lld", key, (long long)value);
  return VSM_E_INVALID;
}

extern "C" int vsm_get_counter(const char* key, int64_t* out_host) {
  if (!key || !out_host) {
    set_error("vsm_get_counter: null argument");
    return VSM_E_INVALID;
  }
  if (!strcmp(key, "select_misses")) {
    *out_host = g_select_misses.load();
    return VSM_OK;
  }
  if (!strcmp(key, "table_retries")) {
    *out_host = g_table_retries.load();
    return VSM_OK;
  }
  if (!strcmp(key, "capacity_retries")) {
    *out_host = g_capacity_retries.load();
    return VSM_OK;
  }
  if (!strcmp(key, "early_collects")) {
    *out_host = g_early_collects.load();
    return VSM_OK;
  }
  set_error("vsm_get_counter: unknown key %s", key);
  return VSM_E_INVALID;
}
