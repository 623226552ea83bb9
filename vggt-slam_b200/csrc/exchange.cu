// exchange.cu -- multi-GPU voxel exchange (the reference is single-process and has no counterpart).
//
// Every rank fuses its own submaps into a local map.  A voxel is owned by rank owner(key) = hash(key) % world.
// pack   groups this map's voxels (key, count, fp32 sums) by owner into contiguous send segments;
// merge  inserts received partial voxels into the owner's map and adds counts and sums.
// The host side (vsm/dist.py) moves the segments with one NCCL all-to-all per array.
#include "hash.cuh"

namespace vsm {

__host__ __device__ __forceinline__ uint32_t owner_of(unsigned long long key, uint32_t world) {
  return (uint32_t)((mix64(key ^ 0x9E3779B97F4A7C15ull) >> 32) % world);
}

// There are only `world` counters: one atomic per element serialises on them (1 M contributor entries on 2 ranks took
// 0.6 ms per kernel).  The lanes of a warp that name the same owner are grouped (match.any) and their leader adds the
// group's size once; every lane gets the group's old value + its rank in the group.  All 32 lanes must call.
__device__ __forceinline__ uint32_t warp_owner_add(uint32_t* counters, uint32_t owner, bool active) {
  const int lane = (int)(threadIdx.x & 31u);
  const unsigned grp = __match_any_sync(0xffffffffu, active ? owner : 0xFFFFFFFFu);
  const int leader = __ffs(grp) - 1;
  uint32_t base = 0;
  if (active && lane == leader) base = atomicAdd(&counters[owner], (uint32_t)__popc(grp));
  base = __shfl_sync(0xffffffffu, base, leader);
  return base + (uint32_t)__popc(grp & ((1u << lane) - 1u));
}

__global__ void __launch_bounds__(256) owner_count_kernel(const unsigned long long* __restrict__ vkey, uint32_t n,
                                                          uint32_t world, uint32_t* __restrict__ owner_cnt) {
  const uint32_t n_round = (n + 31u) & ~31u;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += gridDim.x * blockDim.x) {
    const bool on = i < n;
    warp_owner_add(owner_cnt, on ? owner_of(vkey[i], world) : 0u, on);
  }
}

__global__ void __launch_bounds__(256) owner_place_kernel(const unsigned long long* __restrict__ vkey, uint32_t n,
                                                          uint32_t world, const uint32_t* __restrict__ owner_base,
                                                          uint32_t* __restrict__ owner_cursor, uint32_t* __restrict__ dest) {
  const uint32_t n_round = (n + 31u) & ~31u;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += gridDim.x * blockDim.x) {
    const bool on = i < n;
    const uint32_t o = on ? owner_of(vkey[i], world) : 0u;
    const uint32_t at = warp_owner_add(owner_cursor, o, on);
    if (on) dest[i] = owner_base[o] + at;
  }
}

// one warp per voxel: move key, count and the fp32 sum row to its send position
__global__ void __launch_bounds__(256) pack_rows_kernel(const unsigned long long* __restrict__ vkey,
                                                        const uint32_t* __restrict__ vcount, const float* __restrict__ vsum,
                                                        const uint32_t* __restrict__ dest, uint32_t n, int d,
                                                        unsigned long long* __restrict__ keys, uint32_t* __restrict__ counts,
                                                        float* __restrict__ sums) {
  const int lane = lane_id();
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t i = warp; i < n; i += n_warps) {
    const uint32_t t = dest[i];
    if (lane == 0) {
      keys[t] = vkey[i];
      counts[t] = vcount[i];
    }
    const float4* src = reinterpret_cast<const float4*>(vsum + (size_t)i * d);
    float4* dst = reinterpret_cast<float4*>(sums + (size_t)t * d);
    for (int c = lane; c < d / 4; c += 32) dst[c] = src[c];
  }
}

__global__ void __launch_bounds__(256) merge_keys_kernel(GlobalStore g, const unsigned long long* __restrict__ keys,
                                                         const uint32_t* __restrict__ counts, int64_t n,
                                                         int32_t* __restrict__ gid_out, uint32_t* err) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int gid = global_find_or_insert(g, keys[i], err);
    gid_out[i] = gid;
    if (gid >= 0) atomicAdd(&g.vcount[gid], counts[i]);
  }
}

__global__ void __launch_bounds__(256) merge_rows_kernel(const int32_t* __restrict__ gid, const float* __restrict__ sums,
                                                         int64_t n, int d, float* __restrict__ vsum) {
  const int lane = lane_id();
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t i = warp; i < n; i += n_warps) {
    const int g = gid[i];
    if (g < 0) continue;
    const float4* src = reinterpret_cast<const float4*>(sums + (size_t)i * d);
    float* dst = vsum + (size_t)g * d;
    for (int c = lane; c < d / 4; c += 32) {
      const float4 v = src[c];
      red_add_v4(dst + 4 * c, v.x, v.y, v.z, v.w);
    }
  }
}

// contributor log entries, keyed by voxel key instead of the local dense id
__global__ void __launch_bounds__(256) contrib_owner_count_kernel(const int32_t* __restrict__ log_gid,
                                                                  const unsigned long long* __restrict__ vkey, int64_t n,
                                                                  uint32_t world, uint32_t* __restrict__ owner_cnt) {
  const int64_t n_round = (n + 31) & ~(int64_t)31;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n_round; e += (int64_t)gridDim.x * blockDim.x) {
    const int g = e < n ? log_gid[e] : -1;
    warp_owner_add(owner_cnt, g >= 0 ? owner_of(vkey[g], world) : 0u, g >= 0);
  }
}
__global__ void __launch_bounds__(256) contrib_pack_kernel(const int32_t* __restrict__ log_gid, const int32_t* __restrict__ log_sub,
                                                           const unsigned long long* __restrict__ log_mask,
                                                           const unsigned long long* __restrict__ vkey, int64_t n,
                                                           uint32_t world, const uint32_t* __restrict__ owner_base,
                                                           uint32_t* __restrict__ owner_cursor,
                                                           unsigned long long* __restrict__ keys, int32_t* __restrict__ subs,
                                                           unsigned long long* __restrict__ masks) {
  const int64_t n_round = (n + 31) & ~(int64_t)31;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n_round; e += (int64_t)gridDim.x * blockDim.x) {
    const int g = e < n ? log_gid[e] : -1;
    const unsigned long long key = g >= 0 ? vkey[g] : 0ull;
    const uint32_t o = g >= 0 ? owner_of(key, world) : 0u;
    const uint32_t at = warp_owner_add(owner_cursor, o, g >= 0);
    if (g < 0) continue;
    const uint32_t t = owner_base[o] + at;
    keys[t] = key;
    subs[t] = log_sub[e];
    masks[2 * (size_t)t] = log_mask[2 * (size_t)e];
    masks[2 * (size_t)t + 1] = log_mask[2 * (size_t)e + 1];
  }
}
__global__ void __launch_bounds__(256) contrib_merge_kernel(GlobalStore g, const unsigned long long* __restrict__ keys,
                                                            const int32_t* __restrict__ subs,
                                                            const unsigned long long* __restrict__ masks, int64_t n,
                                                            int64_t log_base, int32_t* __restrict__ log_gid,
                                                            int32_t* __restrict__ log_sub,
                                                            unsigned long long* __restrict__ log_mask) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const unsigned long long key = keys[i];
    int gid = -1;
    uint64_t h = mix64(key) & g.gmask;
    for (uint64_t probes = 0; probes <= g.gmask; ++probes) {
      const unsigned long long cur = g.gkeys[h];
      if (cur == kEmptyKey) break;
      if (cur == key) {
        gid = g.gids[h];
        break;
      }
      h = (h + 1) & g.gmask;
    }
    log_gid[log_base + i] = gid;  // -1 entries are skipped by vsm_finalize
    log_sub[log_base + i] = subs[i];
    log_mask[2 * (size_t)(log_base + i)] = masks[2 * (size_t)i];
    log_mask[2 * (size_t)(log_base + i) + 1] = masks[2 * (size_t)i + 1];
  }
}

static int owner_layout(vsm_map* m, uint32_t world, uint32_t** cnt, uint32_t** base, uint32_t** cursor, cudaStream_t s) {
  VSM_TRY(m->q_tmp.ensure((size_t)3 * world * 4 + 64, s));
  *cnt = m->q_tmp.as<uint32_t>();
  *base = *cnt + world;
  *cursor = *base + world;
  VSM_CUDA(cudaMemsetAsync(*cnt, 0, (size_t)3 * world * 4, s));
  return VSM_OK;
}

static int finish_layout(vsm_map* m, uint32_t world, uint32_t* cnt, uint32_t* base, int64_t* owner_counts_host,
                         cudaStream_t s) {
  std::vector<uint32_t> h(world), b(world);
  VSM_TRY(read_back(m, h.data(), cnt, (size_t)world * 4, s));
  uint32_t run = 0;
  for (uint32_t o = 0; o < world; ++o) {
    b[o] = run;
    run += h[o];
    owner_counts_host[o] = h[o];
  }
  VSM_CUDA(cudaMemcpyAsync(base, b.data(), (size_t)world * 4, cudaMemcpyHostToDevice, s));
  VSM_CUDA(cudaStreamSynchronize(s));  // b goes out of scope
  return VSM_OK;
}

}  // namespace vsm

using namespace vsm;

extern "C" int vsm_partials_pack(vsm_map* m, int32_t world, uint64_t* keys_dev, uint32_t* counts_dev, float* sums_dev,
                                 int64_t* owner_counts_host, void* stream) {
  if (!m || world < 1 || world > 1024 || !owner_counts_host) {
    set_error("vsm_partials_pack: bad arguments");
    return VSM_E_INVALID;
  }
  if (m->dense_loaded) {
    set_error("vsm_partials_pack: dense-loaded maps have no keys");
    return VSM_E_STATE;
  }
  VSM_CUDA(cudaSetDevice(m->device));
  cudaStream_t s = (cudaStream_t)stream;
  VSM_TRY(fuse_collect_pending(m, s));
  const uint32_t V = (uint32_t)m->n_vox;
  for (int o = 0; o < world; ++o) owner_counts_host[o] = 0;
  if (V == 0) return VSM_OK;
  uint32_t *cnt, *base, *cursor;
  VSM_TRY(owner_layout(m, (uint32_t)world, &cnt, &base, &cursor, s));
  owner_count_kernel<<<grid_for(V, 256), 256, 0, s>>>(m->vkey.as<unsigned long long>(), V, (uint32_t)world, cnt);
  VSM_LAUNCHED();
  VSM_TRY(finish_layout(m, (uint32_t)world, cnt, base, owner_counts_host, s));
  if (!keys_dev) return VSM_OK;  // size query only
  if (!counts_dev || !sums_dev) {
    set_error("vsm_partials_pack: null output");
    return VSM_E_INVALID;
  }
  VSM_TRY(m->xch_tmp.ensure((size_t)V * 4, s));
  owner_place_kernel<<<grid_for(V, 256), 256, 0, s>>>(m->vkey.as<unsigned long long>(), V, (uint32_t)world, base, cursor,
                                                      m->xch_tmp.as<uint32_t>());
  VSM_LAUNCHED();
  pack_rows_kernel<<<grid_for((int64_t)V * 32, 256), 256, 0, s>>>(
      m->vkey.as<unsigned long long>(), m->vcount.as<uint32_t>(), m->vsum.as<float>(), m->xch_tmp.as<uint32_t>(), V, m->d,
      reinterpret_cast<unsigned long long*>(keys_dev), counts_dev, sums_dev);
  VSM_LAUNCHED();
  return VSM_OK;
}

extern "C" int vsm_partials_merge(vsm_map* m, const uint64_t* keys_dev, const uint32_t* counts_dev,
                                  const float* sums_dev, int64_t n, void* stream) {
  if (!m || n < 0 || (n > 0 && (!keys_dev || !counts_dev || !sums_dev))) {
    set_error("vsm_partials_merge: bad arguments");
    return VSM_E_INVALID;
  }
  if (m->dense_loaded) {
    set_error("vsm_partials_merge: dense-loaded maps cannot be merged into");
    return VSM_E_STATE;
  }
  if (n == 0) return VSM_OK;
  VSM_CUDA(cudaSetDevice(m->device));
  cudaStream_t s = (cudaStream_t)stream;
  VSM_TRY(fuse_collect_pending(m, s));
  m->finalized = false;
  m->ck_built = false;
  VSM_TRY(map_grow(m, m->n_vox + n, s));
  VSM_TRY(m->ctr.ensure(sizeof(FuseCounters), s));
  FuseCounters* ctr = m->ctr.as<FuseCounters>();
  VSM_CUDA(cudaMemsetAsync(ctr, 0, sizeof(FuseCounters), s));
  VSM_TRY(m->xch_tmp.ensure((size_t)n * 4, s));
  merge_keys_kernel<<<grid_for(n, 256), 256, 0, s>>>(global_store(m), reinterpret_cast<const unsigned long long*>(keys_dev),
                                                     counts_dev, n, m->xch_tmp.as<int32_t>(), &ctr->internal_err);
  VSM_LAUNCHED();
  merge_rows_kernel<<<grid_for(n * 32, 256), 256, 0, s>>>(m->xch_tmp.as<int32_t>(), sums_dev, n, m->d, m->vsum.as<float>());
  VSM_LAUNCHED();
  FuseCounters hc{};
  VSM_TRY(read_back(m, &hc, ctr, sizeof(FuseCounters), s));
  uint32_t n_vox_dev = 0;
  VSM_TRY(read_back(m, &n_vox_dev, m->d_n_vox.p, sizeof(uint32_t), s));
  m->n_vox = n_vox_dev;
  if (hc.internal_err) {
    set_error("internal: global hash overflow while merging (%u)", hc.internal_err);
    return VSM_E_INTERNAL;
  }
  return VSM_OK;
}

extern "C" int vsm_contrib_pack(vsm_map* m, int32_t world, uint64_t* keys_dev, int32_t* submap_ids_dev,
                                uint64_t* masks_dev, int64_t* owner_counts_host, void* stream) {
  if (!m || world < 1 || world > 1024 || !owner_counts_host) {
    set_error("vsm_contrib_pack: bad arguments");
    return VSM_E_INVALID;
  }
  VSM_CUDA(cudaSetDevice(m->device));
  cudaStream_t s = (cudaStream_t)stream;
  VSM_TRY(fuse_collect_pending(m, s));
  const int64_t M = m->log_n;
  for (int o = 0; o < world; ++o) owner_counts_host[o] = 0;
  if (M == 0) return VSM_OK;
  uint32_t *cnt, *base, *cursor;
  VSM_TRY(owner_layout(m, (uint32_t)world, &cnt, &base, &cursor, s));
  contrib_owner_count_kernel<<<grid_for(M, 256), 256, 0, s>>>(m->log_gid.as<int32_t>(), m->vkey.as<unsigned long long>(), M,
                                                              (uint32_t)world, cnt);
  VSM_LAUNCHED();
  VSM_TRY(finish_layout(m, (uint32_t)world, cnt, base, owner_counts_host, s));
  if (!keys_dev) return VSM_OK;
  if (!submap_ids_dev || !masks_dev) {
    set_error("vsm_contrib_pack: null output");
    return VSM_E_INVALID;
  }
  contrib_pack_kernel<<<grid_for(M, 256), 256, 0, s>>>(m->log_gid.as<int32_t>(), m->log_fuse.as<int32_t>(),
                                                       m->log_mask.as<unsigned long long>(),
                                                       m->vkey.as<unsigned long long>(), M, (uint32_t)world, base, cursor,
                                                       reinterpret_cast<unsigned long long*>(keys_dev), submap_ids_dev,
                                                       reinterpret_cast<unsigned long long*>(masks_dev));
  VSM_LAUNCHED();
  return VSM_OK;
}

extern "C" int vsm_contrib_merge(vsm_map* m, const uint64_t* keys_dev, const int32_t* submap_ids_dev,
                                 const uint64_t* masks_dev, int64_t n, void* stream) {
  if (!m || n < 0 || (n > 0 && (!keys_dev || !submap_ids_dev || !masks_dev))) {
    set_error("vsm_contrib_merge: bad arguments");
    return VSM_E_INVALID;
  }
  if (n == 0) return VSM_OK;
  VSM_CUDA(cudaSetDevice(m->device));
  cudaStream_t s = (cudaStream_t)stream;
  VSM_TRY(fuse_collect_pending(m, s));
  m->finalized = false;
  VSM_TRY(log_grow(m, m->log_n + n, s));
  contrib_merge_kernel<<<grid_for(n, 256), 256, 0, s>>>(global_store(m), reinterpret_cast<const unsigned long long*>(keys_dev),
                                                        submap_ids_dev, reinterpret_cast<const unsigned long long*>(masks_dev),
                                                        n, m->log_n, m->log_gid.as<int32_t>(), m->log_fuse.as<int32_t>(),
                                                        m->log_mask.as<unsigned long long>());
  VSM_LAUNCHED();
  m->log_n += n;
  const uint32_t log_n32 = (uint32_t)m->log_n;
  VSM_CUDA(cudaMemcpyAsync(m->d_n_vox.as<uint32_t>() + 1, &log_n32, 4, cudaMemcpyHostToDevice, s));
  VSM_CUDA(cudaStreamSynchronize(s));
  return VSM_OK;
}
