// prep7.cuh -- the one-table preparation of a fuse call and the deferred percentile box (included by fuse.cu only, inside
// namespace vsm, after the pixel map, the submap-local table and the bracket helpers it builds on).
#pragma once

// ---------------------------------------------------------------------------
// one-table preparation (vsm_set_option("prep_variant") bit 3)
// ---------------------------------------------------------------------------
// The two-table preparation above visits every point twice: bbox_coarse_kernel counts the points per coarse cell
// (map.py:271-275), fine_insert_kernel looks the count up and inserts the survivors into the voxel table.  Here every
// bbox survivor is inserted into the voxel table at once and the coarse cells are counted per DISTINCT VOXEL
// afterwards (~10^5 voxels instead of ~5 x 10^6 points): a coarse cell is F x F x F voxels, so the cell of a voxel
// follows from its integer coordinates, floor(i / F).  The reference computes the cell of a POINT in float32,
// floor(p / float32(vs * F)), which differs from floor(floor(p / float32(vs)) / F) for the few points within ~1e-6
// cells of a cell boundary (float32(vs * F) != F * float32(vs), and the two quotients round differently).  Those
// IRREGULAR points (tens per submap) are found exactly -- a cheap conservative band test, then the reference's own
// arithmetic -- and take a side path: they are counted in their float32 cell, and the accepted ones enter the voxel
// table under a PSEUDO key (voxel key | bit 63) which the global merge maps back to the voxel itself.  Every regular
// point of a voxel lies in the voxel's derived cell, so the coarse-cell filter keeps or drops a table entry as a whole.
// The insert also hands every point its ordinal inside its voxel (the old value of the count it adds to), so the
// counting sort needs no second round of atomics: position = segment start + ordinal.
//   insert7        px: bbox test, voxel key, irregularity test, claim + count + frame mask -> (slot, ordinal)
//   coarse7        distinct voxel: derived cell -> coarse table (weighted by the voxel's count); irregular points:
//                  +1 in their float32 cell; last block: accepted irregular points claim their pseudo voxels
//   keep7          distinct voxel: cell count >= min_pts -> kept list, n_fused, exact capacity check; last block decides
//   compact_merge7 kept voxel: segment of the sorted list, global insert, counts, contributor log
//   scatter7       px: entries[segment start + ordinal] = (voxel id, pixel)
constexpr uint32_t kLidReject = 0xFFFFFFFFu;
constexpr unsigned long long kPseudoBit = 1ull << 63;
constexpr uint32_t kIrrCap = 1u << 16;
constexpr uint32_t kFuseForceTwoTables = 1u << 29;  // internal flag: this call must use the two-table preparation

struct alignas(32) IrrEntry {
  unsigned long long fine_key, coarse_key;
  uint32_t pix, frame, cslot, pad;
};

struct CoarseMap {  // coarse axis code from a fine axis code: floor((c + (F-1) 2^20) / F), c >= 1
  FastDiv div;
  uint32_t add;
};
__device__ __forceinline__ unsigned long long coarse_from_fine(unsigned long long key, const CoarseMap& cm) {
  unsigned long long out = 0ull;
#pragma unroll
  for (int ax = 0; ax < 3; ++ax) {
    const uint32_t c = (uint32_t)(key >> (21 * ax)) & 0x1FFFFFu;
    const uint32_t cc = c == 0u ? 0u : fdiv_u32(c + cm.add, cm.div);
    out |= (unsigned long long)cc << (21 * ax);
  }
  return out;
}

// table_claim that also returns the count before the add (the group's first ordinal inside its voxel)
__device__ __forceinline__ int table_claim_ord(const LocalTable& t, unsigned long long key, uint32_t add, bool& is_new,
                                               uint32_t& old, FuseCounters* ctr) {
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
      old = atomicAdd(&sl->count, add);
      return (int)h;
    }
    h = (h + 1u) & t.cap_mask;
  }
  atomicAdd(&ctr->internal_err, 1u);
  return -1;
}

__device__ __forceinline__ void claim_list_append(const LocalTable& t, bool is_new, int slot) {
  const int lane = lane_id();
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
}

// warp_insert + ordinals.  All 32 lanes must call.
__device__ __forceinline__ int warp_insert_ord(const LocalTable& t, bool active, unsigned long long key, int frame, int opts,
                                               FuseCounters* ctr, uint32_t& ord) {
  const int lane = lane_id();
  const unsigned long long k = active ? key : kEmptyKey;
  const unsigned grp = __match_any_sync(0xffffffffu, k);
  const int leader = __ffs(grp) - 1;
  int slot = -1;
  uint32_t old = 0;
  bool is_new = false;
  if (active && lane == leader) slot = table_claim_ord(t, key, (uint32_t)__popc(grp), is_new, old, ctr);
  claim_list_append(t, is_new, slot);
  slot = __shfl_sync(0xffffffffu, slot, leader);
  old = __shfl_sync(0xffffffffu, old, leader);
  ord = old + (uint32_t)__popc(grp & ((1u << lane) - 1u));
  const int lf = __shfl_sync(0xffffffffu, frame, leader);
  if (active && slot >= 0 && (lane == leader || frame != lf)) {
    unsigned long long* mp = &t.slots[slot].mask[frame >> 6];
    const unsigned long long bit = 1ull << (frame & 63);
    if (!(opts & kOptMaskProbe) || (__ldcg(mp) & bit) == 0ull) atomicOr(mp, bit);
  }
  return active ? slot : -1;
}

// insert with a weight per lane (the points of a voxel): lanes with equal keys add their weights with one atomic
__device__ __forceinline__ int warp_insert_weighted(const LocalTable& t, bool active, unsigned long long key, uint32_t w,
                                                    FuseCounters* ctr) {
  const int lane = lane_id();
  const unsigned long long k = active ? key : kEmptyKey;
  const unsigned grp = __match_any_sync(0xffffffffu, k);
  const int leader = __ffs(grp) - 1;
  const uint32_t sum = __reduce_add_sync(grp, active ? w : 0u);
  int slot = -1;
  bool is_new = false;
  if (active && lane == leader) slot = table_claim(t, key, sum, is_new, ctr);
  claim_list_append(t, is_new, slot);
  slot = __shfl_sync(0xffffffffu, slot, leader);
  return active ? slot : -1;
}

struct Prep7Args {
  const float4* pw;
  uint2* pt2;
  IrrEntry* irr;
  uint32_t frame_base;
  float cell;         // float32(voxel size)
  float cell_coarse;  // float32(voxel size * F), the reference's coarse cell
  float inv_coarse;   // 1 / cell_coarse (band test only)
  CoarseMap cm;
  uint32_t min_pts;
  int opts;
};

// Warp-collective (all 32 lanes call): lanes with cand = true hold a point that has passed the confidence, finite and
// percentile-box tests.  Voxel key, irregularity test, claim + count + frame mask; (slot, ordinal) of every lane with
// inside = true is written to pt2 (slot -1: not fused, or irregular -- coarse7 decides).
__device__ __forceinline__ void insert7_lanes(const Prep7Args& a, const LocalTable& tb, FuseCounters* ctr, bool inside,
                                              bool cand, const float4& p, uint32_t pix, int frame, unsigned& n_in) {
  const int lane = lane_id();
  bool act = false, irr = false, never = false;
  unsigned long long key = kEmptyKey, ckey = 0ull;
  if (cand) {
    bool rerr = false;
    key = pack_key(p.x, p.y, p.z, a.cell, rerr);
    if (rerr) {
      // a voxel coordinate that cannot be packed: the call fails, or (drop policy) the point is dropped -- but it
      // still counts in its coarse cell, as in the two-table preparation, through the irregular list
      if (range_problem(true, a.opts, ctr)) {
        bool r2 = false;
        ckey = pack_key(p.x, p.y, p.z, a.cell_coarse, r2);
        if (!r2) {
          n_in += 1u;
          irr = true;
          never = true;
        }
      }
    } else {
      n_in += 1u;  // passes the percentile box (map.py:259-263)
      act = true;
      // Irregular only if p / cell_coarse is within ~3e-7 (relative) of an integer: the three quotients involved
      // (this estimate, the reference's float32 quotient, the voxel coordinate / F) all lie that close to the real
      // p / cell_coarse.  The band is 1e-6: everything outside it is regular without looking further.
      const float tx = __fmul_rn(p.x, a.inv_coarse), ty = __fmul_rn(p.y, a.inv_coarse), tz = __fmul_rn(p.z, a.inv_coarse);
      const bool near_x = fabsf(tx - rintf(tx)) <= fmaf(fabsf(tx), 1e-6f, 1e-6f);
      const bool near_y = fabsf(ty - rintf(ty)) <= fmaf(fabsf(ty), 1e-6f, 1e-6f);
      const bool near_z = fabsf(tz - rintf(tz)) <= fmaf(fabsf(tz), 1e-6f, 1e-6f);
      if (near_x || near_y || near_z) {
        bool r2 = false;
        ckey = pack_key(p.x, p.y, p.z, a.cell_coarse, r2);  // the reference's arithmetic (map.py:271-274)
        irr = ckey != coarse_from_fine(key, a.cm);
        act = !irr;
      }
    }
  }
  uint32_t ord = 0;
  const int slot = warp_insert_ord(tb, act, key, frame, a.opts, ctr, ord);
  const unsigned im = __ballot_sync(0xffffffffu, irr);
  if (im) {
    const int il = __ffs(im) - 1;
    uint32_t base = 0;
    if (lane == il) base = atomicAdd(&ctr->n_irr, (uint32_t)__popc(im));
    base = __shfl_sync(0xffffffffu, base, il);
    if (irr) {
      const uint32_t idx = base + (uint32_t)__popc(im & ((1u << lane) - 1u));
      if (idx < kIrrCap) {
        IrrEntry e;
        e.fine_key = key;
        e.coarse_key = ckey;
        e.pix = pix;
        e.frame = never ? 0xFFFFFFFFu : (uint32_t)frame;  // 0xFFFFFFFF: counted in its cell, never fused
        e.cslot = 0u;
        e.pad = 0u;
        a.irr[idx] = e;
      } else {
        atomicExch(&ctr->irr_overflow, 1u);
      }
    }
  }
  if (inside) a.pt2[pix] = make_uint2((uint32_t)slot, ord);  // irregular points: -1 for now (coarse7 decides)
}

template <bool PATCH>
__global__ void __launch_bounds__(256) insert7_kernel(Prep7Args a, PixMap pm, LocalTable tb, FuseCounters* ctr) {
  const float lx = ctr->bounds[0], hx = ctr->bounds[1], ly = ctr->bounds[2], hy = ctr->bounds[3], lz = ctr->bounds[4],
              hz = ctr->bounds[5];
  unsigned n_in = 0;
  const int lane = lane_id();
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
  for (uint32_t item = warp; item < pm.n_items; item += n_warps) {
    uint32_t pix, fidx;
    const bool inside = map_pixel<PATCH>(pm, item, lane, pix, fidx);
    bool cand = false;
    float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
    if (inside) {
      p = a.pw[pix];
      const uint32_t f = __float_as_uint(p.w);
      cand = (f & (PF_SEL | PF_FINITE)) == (PF_SEL | PF_FINITE) && (p.x >= lx) && (p.x <= hx) && (p.y >= ly) &&
             (p.y <= hy) && (p.z >= lz) && (p.z <= hz);
    }
    insert7_lanes(a, tb, ctr, inside, cand, p, pix, (int)(a.frame_base + fidx), n_in);
  }
  for (int o = 16; o > 0; o >>= 1) n_in += __shfl_xor_sync(0xffffffffu, n_in, o);
  if (lane == 0 && n_in) atomicAdd(&ctr->n_bbox, (unsigned long long)n_in);
}

// ---- deferred percentile box ("select_mode" 3) ------------------------------------------------------------------
// The percentile box needs the 0.5 % / 99.5 % order statistics of all points of the submap -- two more passes over
// the points with the radix select.  Here a sorted sample (bracket_sample_kernel) gives, per axis, a bracket around
// each of the two percentiles, and the brackets alone already classify ~95 % of the points: below the low bracket or
// above the high one -> outside the box; between the two brackets on every axis -> inside, inserted right away.  The
// rest (inside some bracket, outside nowhere) is listed.  The same pass counts the points below / above the brackets
// and collects the coordinates inside them, so bracket_resolve_kernel can pick the exact order statistics (ranks
// relative to the counts) without reading the points again; insert7u_kernel then applies the exact box to the listed
// pixels.  A percentile that falls outside its bracket (probability ~1e-6) or a list that overflows sets sel_miss: the
// call stops before it touches the map and is repeated with the radix select.  Never an approximation.
constexpr int kStageLists = kBrLists + 1;  // + the undecided pixels

__device__ __forceinline__ void stage7_flush(const BracketArgs& br, uint32_t* buf, uint32_t* cnt, int t) {
  const int lane = lane_id();
  const uint32_t n = min(cnt[t], (uint32_t)kBrStage);
  uint32_t base = 0;
  if (lane == 0) base = atomicAdd(t < kBrLists ? &br.bs->cursor[t] : &br.bs->und_cursor, n);
  base = __shfl_sync(0xffffffffu, base, 0);
  for (uint32_t i = lane; i < n; i += 32u) {
    const uint32_t v = buf[t * kBrStage + i];
    if (t < kBrLists) {
      if (base + i < br.cap) br.lists[(size_t)t * br.cap + base + i] = __uint_as_float(v);
    } else if (base + i < br.und_cap) {
      br.und[base + i] = v;
    }
  }
  __syncwarp();
  if (lane == 0) cnt[t] = 0u;
  __syncwarp();
}

template <bool PATCH>
__global__ void __launch_bounds__(256) insert7d_kernel(Prep7Args a, PixMap pm, LocalTable tb, FuseCounters* ctr, BracketArgs br) {
  __shared__ uint32_t s_buf[kBrWarps * kStageLists * kBrStage];
  __shared__ uint32_t s_cnt[kBrWarps * kStageLists];
  uint32_t* const buf = s_buf + (threadIdx.x >> 5) * kStageLists * kBrStage;
  uint32_t* const cnt = s_cnt + (threadIdx.x >> 5) * kStageLists;
  const int lane = lane_id();
  if (lane < kStageLists) cnt[lane] = 0u;
  __syncwarp();
  float lo_l[3], hi_l[3], lo_h[3], hi_h[3];
#pragma unroll
  for (int ax = 0; ax < 3; ++ax) {
    lo_l[ax] = br.bs->lo[2 * ax];
    hi_l[ax] = br.bs->hi[2 * ax];
    lo_h[ax] = br.bs->lo[2 * ax + 1];
    hi_h[ax] = br.bs->hi[2 * ax + 1];
  }
  uint32_t below[3] = {0u, 0u, 0u}, above[3] = {0u, 0u, 0u};
  unsigned n_in = 0;
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
  for (uint32_t item = warp; item < pm.n_items; item += n_warps) {
    uint32_t pix, fidx;
    const bool inside = map_pixel<PATCH>(pm, item, lane, pix, fidx);
    bool cand = false;
    float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
    if (inside) {
      p = a.pw[pix];
      const uint32_t f = __float_as_uint(p.w);
      if ((f & (PF_SEL | PF_FINITE)) == (PF_SEL | PF_FINITE)) {
        const float c3[3] = {p.x, p.y, p.z};
        bool out = false, und = false;
#pragma unroll
        for (int ax = 0; ax < 3; ++ax) {
          const float c = c3[ax];
          if (c <= hi_l[ax]) {  // at or below the low bracket's upper end: ~1 % of the points
            if (c < lo_l[ax]) {
              ++below[ax];
              out = true;
            } else {
              und = true;
              buf[(2 * ax) * kBrStage + atomicAdd(&cnt[2 * ax], 1u)] = __float_as_uint(c);
            }
          }
          if (c >= lo_h[ax]) {
            if (c > hi_h[ax]) {
              ++above[ax];
              out = true;
            } else {
              und = true;
              buf[(2 * ax + 1) * kBrStage + atomicAdd(&cnt[2 * ax + 1], 1u)] = __float_as_uint(c);
            }
          }
        }
        cand = !out && !und;
        if (und && !out) buf[kBrLists * kBrStage + atomicAdd(&cnt[kBrLists], 1u)] = pix;  // box test waits for the percentiles
      }
    }
    // a list is flushed once it is half full: an iteration adds at most 32 values (one per lane) to each
    __syncwarp();
    unsigned full = __ballot_sync(0xffffffffu, lane < kStageLists && cnt[lane < kStageLists ? lane : 0] >= (uint32_t)kBrStage / 2);
    while (full) {
      const int t = __ffs(full) - 1;
      full &= full - 1;
      stage7_flush(br, buf, cnt, t);
    }
    insert7_lanes(a, tb, ctr, inside, cand, p, pix, (int)(a.frame_base + fidx), n_in);
  }
  __syncwarp();
  for (int t = 0; t < kStageLists; ++t)
    if (cnt[t]) stage7_flush(br, buf, cnt, t);
#pragma unroll
  for (int ax = 0; ax < 3; ++ax) {
    uint32_t b = below[ax], ab = above[ax];
    for (int o = 16; o > 0; o >>= 1) {
      b += __shfl_xor_sync(0xffffffffu, b, o);
      ab += __shfl_xor_sync(0xffffffffu, ab, o);
    }
    if (lane == 0) {
      if (b) atomicAdd(&br.bs->below[2 * ax], b);
      if (ab) atomicAdd(&br.bs->above[2 * ax + 1], ab);
    }
  }
  for (int o = 16; o > 0; o >>= 1) n_in += __shfl_xor_sync(0xffffffffu, n_in, o);
  if (lane == 0 && n_in) atomicAdd(&ctr->n_bbox, (unsigned long long)n_in);
}

// the listed pixels against the exact box (ctr->bounds, written by bracket_resolve_kernel; NaN after a miss: nothing passes)
__global__ void __launch_bounds__(256) insert7u_kernel(Prep7Args a, PixMap pm, LocalTable tb, FuseCounters* ctr, BracketArgs br) {
  const float lx = ctr->bounds[0], hx = ctr->bounds[1], ly = ctr->bounds[2], hy = ctr->bounds[3], lz = ctr->bounds[4],
              hz = ctr->bounds[5];
  const uint32_t n = min(br.bs->und_cursor, br.und_cap);
  const uint32_t n_round = (n + 31u) & ~31u;
  unsigned n_in = 0;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += gridDim.x * blockDim.x) {
    const bool inside = i < n;
    uint32_t pix = 0;
    float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
    bool cand = false;
    if (inside) {
      pix = br.und[i];
      p = a.pw[pix];
      cand = (p.x >= lx) && (p.x <= hx) && (p.y >= ly) && (p.y <= hy) && (p.z >= lz) && (p.z <= hz);
    }
    insert7_lanes(a, tb, ctr, inside, cand, p, pix, (int)(a.frame_base + fdiv_u32(pix, pm.div_ppf)), n_in);
  }
  for (int o = 16; o > 0; o >>= 1) n_in += __shfl_xor_sync(0xffffffffu, n_in, o);
  if (lane_id() == 0 && n_in) atomicAdd(&ctr->n_bbox, (unsigned long long)n_in);
}

__global__ void __launch_bounds__(256) coarse7_kernel(Prep7Args a, LocalTable tb, LocalTable ta, FuseCounters* ctr) {
  const uint32_t n_occ = ctr->n_occ_b;
  const uint32_t n_round = (n_occ + 31u) & ~31u;
  for (uint32_t lid = blockIdx.x * blockDim.x + threadIdx.x; lid < n_round; lid += gridDim.x * blockDim.x) {
    const bool on = lid < n_occ;
    uint32_t slot = 0, cnt = 0;
    unsigned long long ckey = kEmptyKey;
    if (on) {
      slot = tb.slot_list[lid];
      const Slot* sl = tb.slots + slot;
      ckey = coarse_from_fine(sl->key, a.cm);
      cnt = sl->count;
    }
    const int cs = warp_insert_weighted(ta, on, ckey, cnt, ctr);
    if (on) tb.slots[slot].lid = (uint32_t)cs;  // the voxel's coarse cell, until keep7 decides
  }
  const uint32_t n_irr = min(ctr->n_irr, kIrrCap);
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_irr; i += gridDim.x * blockDim.x) {
    bool nw = false;
    const int cs = table_claim(ta, a.irr[i].coarse_key, 1u, nw, ctr);
    if (nw) {
      const uint32_t at = atomicAdd(ta.n_occ, 1u);
      if (at + 1u > ta.limit) atomicExch(ta.overflow, 1u);
      ta.slot_list[at] = (uint32_t)cs;
    }
    a.irr[i].cslot = (uint32_t)cs;
  }
  if (n_irr == 0u) return;
  // the last block to finish sees every cell's final count: accepted irregular points enter the voxel table
  __shared__ bool s_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(&ctr->ticket2, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  for (uint32_t i = threadIdx.x; i < n_irr; i += blockDim.x) {
    const IrrEntry* e = a.irr + i;
    const uint32_t cs = __ldcg(&e->cslot);
    const uint32_t fr = __ldcg(&e->frame);
    if (fr == 0xFFFFFFFFu || (int)cs < 0 || __ldcg(&ta.slots[cs].count) < a.min_pts) continue;
    const unsigned long long pkey = __ldcg(&e->fine_key) | kPseudoBit;
    if (pkey == kEmptyKey) {
      atomicAdd(&ctr->internal_err, 1u);
      continue;
    }
    bool nw = false;
    uint32_t old = 0;
    const int s = table_claim_ord(tb, pkey, 1u, nw, old, ctr);
    if (s < 0) continue;
    if (nw) {
      const uint32_t at = atomicAdd(tb.n_occ, 1u);
      if (at + 1u > tb.limit) atomicExch(tb.overflow, 1u);
      tb.slot_list[at] = (uint32_t)s;
    }
    atomicOr(&tb.slots[s].mask[fr >> 6], 1ull << (fr & 63u));
    a.pt2[__ldcg(&e->pix)] = make_uint2((uint32_t)s, old);
  }
}

__global__ void __launch_bounds__(256) keep7_kernel(LocalTable tb, LocalTable ta, GlobalStore g, FuseCounters* ctr,
                                                    uint32_t* __restrict__ kept, uint32_t min_pts,
                                                    uint32_t* __restrict__ map_state, uint32_t vcap, uint32_t log_cap,
                                                    uint32_t entry_cap) {
  const uint32_t n_occ = ctr->n_occ_b;
  const uint32_t n_round = (n_occ + 31u) & ~31u;
  const int lane = lane_id();
  unsigned n_new = 0, n_dist = 0;
  unsigned long long n_fused = 0;
  for (uint32_t lid = blockIdx.x * blockDim.x + threadIdx.x; lid < n_round; lid += gridDim.x * blockDim.x) {
    const bool on = lid < n_occ;
    uint32_t slot = 0, cnt = 0;
    unsigned long long key = 0ull;
    bool keep = false;
    if (on) {
      slot = tb.slot_list[lid];
      const Slot* sl = tb.slots + slot;
      key = sl->key;
      cnt = sl->count;
      keep = (key & kPseudoBit) != 0ull || ta.slots[sl->lid].count >= min_pts;
    }
    const unsigned km = __ballot_sync(0xffffffffu, keep);
    if (km) {
      const int kl = __ffs(km) - 1;
      uint32_t base = 0;
      if (lane == kl) base = atomicAdd(&ctr->n_kept, (uint32_t)__popc(km));
      base = __shfl_sync(0xffffffffu, base, kl);
      if (keep) kept[base + (uint32_t)__popc(km & ((1u << lane) - 1u))] = slot;
    }
    if (keep) {
      n_fused += cnt;
      const unsigned long long real = key & ~kPseudoBit;
      bool found = false;
      uint64_t h = mix64(real) & g.gmask;
      for (uint64_t probes = 0; probes <= g.gmask; ++probes) {
        const unsigned long long cur = g.gkeys[h];
        if (cur == real) {
          found = true;
          break;
        }
        if (cur == kEmptyKey) break;
        h = (h + 1) & g.gmask;
      }
      n_new += found ? 0u : 1u;  // a pseudo voxel and its voxel both count: an upper bound, which is all the check needs
      bool distinct = true;
      if (key & kPseudoBit) {
        // n_submap_voxels counts voxels: not distinct if the voxel itself is in the table and kept
        uint32_t h2 = (uint32_t)mix64(real) & tb.cap_mask;
        for (uint32_t probes = 0; probes <= tb.cap_mask; ++probes) {
          const unsigned long long cur = tb.slots[h2].key;
          if (cur == kEmptyKey) break;
          if (cur == real) {
            const uint32_t l2 = __ldcg(&tb.slots[h2].lid);  // its cell, or kLidReject if this kernel has dropped it already
            distinct = l2 == kLidReject || ta.slots[l2].count < min_pts;
            break;
          }
          h2 = (h2 + 1u) & tb.cap_mask;
        }
      }
      n_dist += distinct ? 1u : 0u;
    } else if (on) {
      tb.slots[slot].lid = kLidReject;
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    n_new += __shfl_xor_sync(0xffffffffu, n_new, o);
    n_dist += __shfl_xor_sync(0xffffffffu, n_dist, o);
    n_fused += __shfl_xor_sync(0xffffffffu, n_fused, o);
  }
  if (lane == 0) {
    if (n_new) atomicAdd(&ctr->n_new, n_new);
    if (n_dist) atomicAdd(&ctr->n_distinct, n_dist);
    if (n_fused) atomicAdd(&ctr->n_fused, n_fused);
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0 && atomicAdd(&ctr->ticket, 1u) == gridDim.x - 1) {
    __threadfence();
    const uint32_t total_new = atomicAdd(&ctr->n_new, 0u);
    const uint32_t n_kept = atomicAdd(&ctr->n_kept, 0u);
    const unsigned long long fused = atomicAdd(&ctr->n_fused, 0ull);
    const uint32_t log_n = map_state[1];
    const bool fits = ((unsigned long long)map_state[0] + total_new <= vcap) &&
                      ((unsigned long long)log_n + n_kept <= log_cap) &&
                      (ctr->n_finite <= (unsigned long long)entry_cap || fused <= (unsigned long long)entry_cap);
    if (!fits || ctr->range_err || ctr->internal_err || ctr->sel_miss || ctr->bad_index || ctr->irr_overflow ||
        ctr->tbl_overflow) {
      ctr->abort = 1u;
    } else {
      ctr->log_base = log_n;
      ctr->vox_base = map_state[0];
      map_state[1] = log_n + n_kept;
    }
  }
}

// compact_merge_kernel over the kept list.  Pseudo voxels merge into their voxel (two local entries, one global id:
// the waits for a published id therefore come AFTER the warp's own publications).  Leaves (segment start, voxel id) in
// the slot's first mask word -- the frame masks have gone to the contributor log -- for scatter7.
__global__ void __launch_bounds__(256) compact_merge7_kernel(LocalTable tb, const uint32_t* __restrict__ kept, GlobalStore g,
                                                             FuseCounters* ctr, int32_t* __restrict__ log_gid,
                                                             int32_t* __restrict__ log_sub,
                                                             unsigned long long* __restrict__ log_mask, int32_t submap_id) {
  if (ctr->abort) return;
  const uint32_t n_kept = ctr->n_kept;
  const uint32_t n_round = (n_kept + 31u) & ~31u;
  const size_t log_base = ctr->log_base;
  const int lane = lane_id();
  for (uint32_t lid = blockIdx.x * blockDim.x + threadIdx.x; lid < n_round; lid += gridDim.x * blockDim.x) {
    const bool on = lid < n_kept;
    uint32_t cnt = 0, slot = 0;
    unsigned long long key = kEmptyKey, m0 = 0ull, m1 = 0ull;
    if (on) {
      slot = kept[lid];
      const Slot* sl = tb.slots + slot;
      key = sl->key & ~kPseudoBit;
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
    int gid = -1;
    uint64_t hslot = 0;
    bool claimed = false, found = false;
    if (on) {
      uint64_t h = mix64(key) & g.gmask;
      for (uint64_t probes = 0; probes <= g.gmask; ++probes) {
        unsigned long long cur = g.gkeys[h];
        if (cur == kEmptyKey) {
          cur = atomicCAS(&g.gkeys[h], kEmptyKey, key);
          if (cur == kEmptyKey) {
            claimed = true;
            hslot = h;
            break;
          }
        }
        if (cur == key) {
          found = true;
          hslot = h;
          break;
        }
        h = (h + 1) & g.gmask;
      }
      if (!claimed && !found) atomicAdd(&ctr->internal_err, 1u);
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
          reinterpret_cast<volatile int32_t*>(g.gids)[hslot] = -2;
        } else {
          g.vkey[id] = key;
          __threadfence();
          reinterpret_cast<volatile int32_t*>(g.gids)[hslot] = (int32_t)id;
          gid = (int)id;
        }
      }
    }
    if (found) {
      while ((gid = reinterpret_cast<volatile int32_t*>(g.gids)[hslot]) == -1) {
      }
    }
    if (on) {
      Slot* sl = tb.slots + slot;
      sl->lid = lid;
      sl->mask[0] = (unsigned long long)(base + incl - cnt) | ((unsigned long long)(uint32_t)gid << 32);
      if (gid >= 0) atomicAdd(&g.vcount[gid], cnt);
      log_gid[log_base + lid] = gid;
      log_sub[log_base + lid] = submap_id;
      log_mask[2 * (log_base + lid)] = m0;
      log_mask[2 * (log_base + lid) + 1] = m1;
    }
  }
}

// counting sort without a second round of atomics: entries[segment start + ordinal] = (voxel id, pixel); check-only
// pixels behind.  entries == nullptr (pixel-order / host paths): only the per-pixel voxel ids (-2 = check-only).
__global__ void __launch_bounds__(256) scatter7_kernel(const uint2* __restrict__ pt2, const float4* __restrict__ pw,
                                                       uint32_t n_px, LocalTable tb, int mark_checks,
                                                       unsigned long long* __restrict__ entries,
                                                       int32_t* __restrict__ point_gid, FuseCounters* ctr) {
  if (ctr->abort) return;
  const uint32_t n_fused = (uint32_t)ctr->n_fused;
  const int lane = lane_id();
  const uint32_t n_round = (n_px + 31u) & ~31u;
  for (uint32_t pix = blockIdx.x * blockDim.x + threadIdx.x; pix < n_round; pix += gridDim.x * blockDim.x) {
    const bool inside = pix < n_px;
    int gid = -1;
    if (inside) {
      const uint2 ps = pt2[pix];
      if ((int)ps.x >= 0) {
        const Slot* sl = tb.slots + ps.x;
        if (sl->lid != kLidReject) {
          const unsigned long long m0 = sl->mask[0];
          gid = (int)(uint32_t)(m0 >> 32);
          if (entries != nullptr && gid >= 0) {
            const uint32_t pos = (uint32_t)m0 + ps.y;  // segment start + ordinal: inside the list by construction
            if (pos < n_fused)
              entries[pos] = make_entry(gid, pix);
            else
              atomicAdd(&ctr->internal_err, 1u);
          }
        }
      }
    }
    if (mark_checks) {
      bool chk = false;
      if (inside && gid < 0) {
        const uint32_t f = __float_as_uint(pw[pix].w);
        chk = (f & (PF_SEL | PF_FINITE)) == (PF_SEL | PF_FINITE);
      }
      if (entries != nullptr) {
        const unsigned cm = __ballot_sync(0xffffffffu, chk);
        if (cm) {
          uint32_t cbase = 0;
          if (lane == __ffs(cm) - 1) cbase = atomicAdd(&ctr->n_check, (uint32_t)__popc(cm));
          cbase = __shfl_sync(0xffffffffu, cbase, __ffs(cm) - 1);
          if (chk) entries[n_fused + cbase + (uint32_t)__popc(cm & ((1u << lane) - 1u))] = make_entry(-2, pix);
        }
      } else if (chk) {
        gid = -2;
      }
    }
    if (point_gid != nullptr && inside) point_gid[pix] = gid;
  }
}

