// finalize.cu -- voxel ordering, export, position lookup.
//
//   vsm_finalize          np.unique(axis=0) order of the voxel keys                  vggt_slam/map.py:352, submap.py:283
//   vsm_export_geometry   centres ((coords + 0.5) * vs), counts, reconstructed coords vggt_slam/map.py:362,
//                                                                                    semantic_voxel.py:62-66
//   vsm_export_features   feat_sum / counts                                          vggt_slam/map.py:360
//   vsm_lookup            get_index_at_position                                      vggt_slam/semantic_voxel.py:68-80
//   vsm_map_load_dense    load_from_directory                                        vggt_slam/semantic_voxel.py:150-165
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "state.cuh"

namespace vsm {

__global__ void iota_kernel(uint32_t* p, uint32_t n) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) p[i] = i;
}
__global__ void invert_perm_kernel(const uint32_t* __restrict__ id_of_rank, uint32_t* __restrict__ rank_of_id, uint32_t n) {
  for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x) rank_of_id[id_of_rank[r]] = r;
}
__global__ void log_rank_kernel(const int32_t* __restrict__ log_gid, const uint32_t* __restrict__ rank_of_id,
                                uint32_t* __restrict__ out_rank, uint32_t* __restrict__ per_rank, uint32_t n) {
  for (uint32_t e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
    const int g = log_gid[e];
    const uint32_t r = g >= 0 ? rank_of_id[g] : 0xFFFFFFFFu;
    out_rank[e] = r;
    if (g >= 0) atomicAdd(&per_rank[r], 1u);
  }
}
__global__ void csr_gather_kernel(const uint32_t* __restrict__ entry_of_pos, const int32_t* __restrict__ log_sub,
                                  const unsigned long long* __restrict__ log_mask, int32_t* __restrict__ csr_sub,
                                  unsigned long long* __restrict__ csr_mask, uint32_t n) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const uint32_t e = entry_of_pos[i];
    csr_sub[i] = log_sub[e];
    csr_mask[2 * (size_t)i] = log_mask[2 * (size_t)e];
    csr_mask[2 * (size_t)i + 1] = log_mask[2 * (size_t)e + 1];
  }
}

__device__ __forceinline__ int64_t recon_axis(float center, float vs) {
  // floor(center / vs - 0.5) in float32, then the reference's float -> int64 cast
  const float q = floorf(__fsub_rn(__fdiv_rn(center, vs), 0.5f));
  if (!(fabsf(q) < 9223372036854775808.0f)) return (int64_t)0x8000000000000000ull;
  return (int64_t)q;
}
__device__ __forceinline__ float center_axis(int64_t c, float vs) { return __fmul_rn(__fadd_rn((float)c, 0.5f), vs); }

__global__ void __launch_bounds__(256) export_geometry_kernel(const unsigned long long* __restrict__ sorted_keys,
                                                              const uint32_t* __restrict__ id_of_rank,
                                                              const uint32_t* __restrict__ vcount,
                                                              const float* __restrict__ dense_centers, float vs,
                                                              uint32_t n, int64_t* __restrict__ coords,
                                                              float* __restrict__ centers, int64_t* __restrict__ counts,
                                                              int64_t* __restrict__ recon) {
  for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x) {
    float c[3];
    int64_t k[3];
    if (dense_centers) {
      for (int a = 0; a < 3; ++a) {
        c[a] = dense_centers[3 * (size_t)r + a];
        k[a] = recon_axis(c[a], vs);
      }
    } else {
      unpack_key(sorted_keys[r], k[0], k[1], k[2]);
      for (int a = 0; a < 3; ++a) c[a] = center_axis(k[a], vs);
    }
    for (int a = 0; a < 3; ++a) {
      if (coords) coords[3 * (size_t)r + a] = k[a];
      if (centers) centers[3 * (size_t)r + a] = c[a];
      if (recon) recon[3 * (size_t)r + a] = recon_axis(c[a], vs);
    }
    if (counts) counts[r] = (int64_t)vcount[id_of_rank[r]];
  }
}

__global__ void __launch_bounds__(256) export_features_kernel(const float* __restrict__ vsum,
                                                              const uint32_t* __restrict__ vcount,
                                                              const uint32_t* __restrict__ id_of_rank, int64_t r0,
                                                              int64_t rows, int d, float* __restrict__ out) {
  const int64_t total = rows * (d / 4);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = i / (d / 4);
    const int c4 = (int)(i % (d / 4));
    const uint32_t id = id_of_rank[r0 + row];
    const float4 v = *reinterpret_cast<const float4*>(vsum + (size_t)id * d + 4 * c4);
    const double cnt = (double)vcount[id];
    float4 o;
    o.x = __double2float_rn(__ddiv_rn((double)v.x, cnt));
    o.y = __double2float_rn(__ddiv_rn((double)v.y, cnt));
    o.z = __double2float_rn(__ddiv_rn((double)v.z, cnt));
    o.w = __double2float_rn(__ddiv_rn((double)v.w, cnt));
    *reinterpret_cast<float4*>(out + (size_t)row * d + 4 * c4) = o;
  }
}

__global__ void widen_u32_kernel(const uint32_t* __restrict__ in, int64_t* __restrict__ out, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) out[i] = in[i];
}

__global__ void point_index_kernel(const int32_t* __restrict__ point_gid, const uint32_t* __restrict__ rank_of_id,
                                   int64_t n, int32_t* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int g = point_gid[i];
    out[i] = g >= 0 ? (int32_t)rank_of_id[g] : -1;
  }
}

// ---- position lookup -------------------------------------------------------
__device__ __forceinline__ bool pack_coords(int64_t x, int64_t y, int64_t z, unsigned long long& key) {
  const int64_t c[3] = {x, y, z};
  unsigned long long k = 0;
  for (int a = 0; a < 3; ++a) {
    unsigned long long code;
    if (c[a] == (int64_t)0x8000000000000000ull)
      code = 0;
    else if (c[a] > -(int64_t)kAxisBias && c[a] < (int64_t)kAxisBias)
      code = (unsigned long long)(c[a] + kAxisBias);
    else
      return false;
    k = (k << 21) | code;
  }
  key = k;
  return true;
}

// compat table: reconstructed coords -> highest sorted index (dict semantics of semantic_voxel.py:39-41)
__global__ void __launch_bounds__(256) compat_build_kernel(const unsigned long long* __restrict__ sorted_keys,
                                                           const float* __restrict__ dense_centers, float vs, uint32_t n,
                                                           unsigned long long* __restrict__ ck_keys,
                                                           uint32_t* __restrict__ ck_val, uint64_t mask) {
  for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x) {
    int64_t k[3];
    if (dense_centers) {
      for (int a = 0; a < 3; ++a) k[a] = recon_axis(dense_centers[3 * (size_t)r + a], vs);
    } else {
      unpack_key(sorted_keys[r], k[0], k[1], k[2]);
      for (int a = 0; a < 3; ++a) k[a] = recon_axis(center_axis(k[a], vs), vs);
    }
    unsigned long long key;
    if (!pack_coords(k[0], k[1], k[2], key)) continue;  // unreachable from any packable query position
    uint64_t h = mix64(key) & mask;
    while (true) {
      unsigned long long cur = ck_keys[h];
      if (cur == kEmptyKey) cur = atomicCAS(&ck_keys[h], kEmptyKey, key);
      if (cur == kEmptyKey || cur == key) {
        atomicMax(&ck_val[h], r + 1u);  // 0 = unset
        break;
      }
      h = (h + 1) & mask;
    }
  }
}

__global__ void __launch_bounds__(256) lookup_kernel(const float* __restrict__ pos, int64_t M, float vs,
                                                     const unsigned long long* __restrict__ keys,
                                                     const void* __restrict__ vals, uint64_t mask,
                                                     const uint32_t* __restrict__ rank_of_id, int compat,
                                                     int64_t* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < M; i += (int64_t)gridDim.x * blockDim.x) {
    bool rerr = false;
    const unsigned long long key = pack_key(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2], vs, rerr);
    int64_t res = -1;
    if (!rerr) {
      uint64_t h = mix64(key) & mask;
      for (uint64_t probes = 0; probes <= mask; ++probes) {
        const unsigned long long cur = keys[h];
        if (cur == kEmptyKey) break;
        if (cur == key) {
          if (compat) {
            res = (int64_t)reinterpret_cast<const uint32_t*>(vals)[h] - 1;
          } else {
            const int id = reinterpret_cast<const int32_t*>(vals)[h];
            res = id >= 0 ? (int64_t)rank_of_id[id] : -1;
          }
          break;
        }
        h = (h + 1) & mask;
      }
    }
    out[i] = res;
  }
}

__global__ void fill_u32_kernel(uint32_t* p, uint32_t v, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = v;
}

static int require_finalized(const vsm_map* m) {
  if (!m) {
    set_error("null map");
    return VSM_E_INVALID;
  }
  if (!m->finalized) {
    set_error("map is not finalised: call vsm_finalize after the last fuse / merge");
    return VSM_E_STATE;
  }
  return VSM_OK;
}

}  // namespace vsm

using namespace vsm;

extern "C" int vsm_finalize(vsm_map* m, void* stream) {
  if (!m) {
    set_error("null map");
    return VSM_E_INVALID;
  }
  VSM_CUDA(cudaSetDevice(m->device));
  cudaStream_t s = (cudaStream_t)stream;
  VSM_TRY(fuse_collect_pending(m, s));
  const uint32_t V = (uint32_t)m->n_vox;
  m->ck_built = false;
  m->norms_valid = false;
  VSM_TRY(m->id_of_rank.ensure(std::max<size_t>((size_t)V * 4, 16), s));
  VSM_TRY(m->rank_of_id.ensure(std::max<size_t>((size_t)V * 4, 16), s));
  VSM_TRY(m->sorted_keys.ensure(std::max<size_t>((size_t)V * 8, 16), s));
  if (V == 0) {
    m->csr_entries = 0;
    VSM_TRY(m->csr_off.ensure(16, s));
    VSM_CUDA(cudaMemsetAsync(m->csr_off.p, 0, 16, s));
    m->finalized = true;
    return VSM_OK;
  }
  if (m->dense_loaded) {
    iota_kernel<<<grid_for(V, 256), 256, 0, s>>>(m->id_of_rank.as<uint32_t>(), V);
    VSM_LAUNCHED();
    iota_kernel<<<grid_for(V, 256), 256, 0, s>>>(m->rank_of_id.as<uint32_t>(), V);
    VSM_LAUNCHED();
  } else {
    // rank_of_id doubles as the iota input of the sort
    iota_kernel<<<grid_for(V, 256), 256, 0, s>>>(m->rank_of_id.as<uint32_t>(), V);
    VSM_LAUNCHED();
    size_t tmp = 0;
    VSM_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp, m->vkey.as<unsigned long long>(),
                                             m->sorted_keys.as<unsigned long long>(), m->rank_of_id.as<uint32_t>(),
                                             m->id_of_rank.as<uint32_t>(), (int)V, 0, 63, s));
    VSM_TRY(m->cub_tmp.ensure(tmp, s));
    VSM_CUDA(cub::DeviceRadixSort::SortPairs(m->cub_tmp.p, tmp, m->vkey.as<unsigned long long>(),
                                             m->sorted_keys.as<unsigned long long>(), m->rank_of_id.as<uint32_t>(),
                                             m->id_of_rank.as<uint32_t>(), (int)V, 0, 63, s));
    ++g_launches;
    invert_perm_kernel<<<grid_for(V, 256), 256, 0, s>>>(m->id_of_rank.as<uint32_t>(), m->rank_of_id.as<uint32_t>(), V);
    VSM_LAUNCHED();
  }
  // contributor CSR: log entries ordered by sorted voxel index (stable: fuse order inside a voxel)
  const uint32_t M = (uint32_t)m->log_n;
  VSM_TRY(m->csr_off.ensure(((size_t)V + 1) * 4, s));
  VSM_CUDA(cudaMemsetAsync(m->csr_off.p, 0, ((size_t)V + 1) * 4, s));
  m->csr_entries = 0;
  if (M > 0) {
    DevBuf ranks, ranks_sorted, ent, ent_sorted, per_rank;
    VSM_TRY(ranks.ensure((size_t)M * 4, s));
    VSM_TRY(ranks_sorted.ensure((size_t)M * 4, s));
    VSM_TRY(ent.ensure((size_t)M * 4, s));
    VSM_TRY(ent_sorted.ensure((size_t)M * 4, s));
    VSM_TRY(per_rank.ensure(((size_t)V + 1) * 4, s));
    int status = VSM_OK;
    do {
      if (cudaMemsetAsync(per_rank.p, 0, ((size_t)V + 1) * 4, s) != cudaSuccess) {
        status = VSM_E_CUDA;
        break;
      }
      log_rank_kernel<<<grid_for(M, 256), 256, 0, s>>>(m->log_gid.as<int32_t>(), m->rank_of_id.as<uint32_t>(),
                                                       ranks.as<uint32_t>(), per_rank.as<uint32_t>(), M);
      iota_kernel<<<grid_for(M, 256), 256, 0, s>>>(ent.as<uint32_t>(), M);
      g_launches += 2;
      size_t tmp = 0;
      cub::DeviceRadixSort::SortPairs(nullptr, tmp, ranks.as<uint32_t>(), ranks_sorted.as<uint32_t>(),
                                      ent.as<uint32_t>(), ent_sorted.as<uint32_t>(), (int)M, 0, 32, s);
      size_t tmp2 = 0;
      cub::DeviceScan::ExclusiveSum(nullptr, tmp2, per_rank.as<uint32_t>(), m->csr_off.as<uint32_t>(), (int)V + 1, s);
      if ((status = m->cub_tmp.ensure(std::max(tmp, tmp2), s)) != VSM_OK) break;
      if (cub::DeviceRadixSort::SortPairs(m->cub_tmp.p, tmp, ranks.as<uint32_t>(), ranks_sorted.as<uint32_t>(),
                                          ent.as<uint32_t>(), ent_sorted.as<uint32_t>(), (int)M, 0, 32,
                                          s) != cudaSuccess ||
          cub::DeviceScan::ExclusiveSum(m->cub_tmp.p, tmp2, per_rank.as<uint32_t>(), m->csr_off.as<uint32_t>(),
                                        (int)V + 1, s) != cudaSuccess) {
        status = VSM_E_CUDA;
        break;
      }
      g_launches += 2;
      if ((status = m->csr_sub.ensure((size_t)M * 4, s)) != VSM_OK) break;
      if ((status = m->csr_mask.ensure((size_t)M * 16, s)) != VSM_OK) break;
      csr_gather_kernel<<<grid_for(M, 256), 256, 0, s>>>(ent_sorted.as<uint32_t>(), m->log_fuse.as<int32_t>(),
                                                         m->log_mask.as<unsigned long long>(), m->csr_sub.as<int32_t>(),
                                                         m->csr_mask.as<unsigned long long>(), M);
      ++g_launches;
      if (cudaStreamSynchronize(s) != cudaSuccess) status = VSM_E_CUDA;
    } while (0);
    ranks.release();
    ranks_sorted.release();
    ent.release();
    ent_sorted.release();
    per_rank.release();
    if (status != VSM_OK) {
      if (status == VSM_E_CUDA) set_error("vsm_finalize: %s", cudaGetErrorString(cudaGetLastError()));
      return status;
    }
    m->csr_entries = M;
  }
  VSM_CUDA(cudaStreamSynchronize(s));
  m->finalized = true;
  return VSM_OK;
}

extern "C" int vsm_export_geometry(const vsm_map* m, int64_t* coords_dev, float* centers_dev, int64_t* counts_dev,
                                   int64_t* recon_coords_dev, void* stream) {
  VSM_TRY(require_finalized(m));
  const uint32_t V = (uint32_t)m->n_vox;
  if (V == 0) return VSM_OK;
  VSM_CUDA(cudaSetDevice(m->device));
  export_geometry_kernel<<<grid_for(V, 256), 256, 0, (cudaStream_t)stream>>>(
      m->sorted_keys.as<unsigned long long>(), m->id_of_rank.as<uint32_t>(), m->vcount.as<uint32_t>(),
      m->dense_loaded ? m->dense_centers.as<float>() : nullptr, m->vs_f, V, coords_dev, centers_dev, counts_dev,
      recon_coords_dev);
  VSM_LAUNCHED();
  return VSM_OK;
}

extern "C" int vsm_export_features(const vsm_map* m, int64_t r0, int64_t r1, float* out_dev, void* stream) {
  VSM_TRY(require_finalized(m));
  if (r0 < 0 || r1 < r0 || r1 > m->n_vox || !out_dev) {
    set_error("vsm_export_features: bad row range [%lld,%lld) of %lld", (long long)r0, (long long)r1,
              (long long)m->n_vox);
    return VSM_E_INVALID;
  }
  if (r1 == r0) return VSM_OK;
  VSM_CUDA(cudaSetDevice(m->device));
  const int64_t total = (r1 - r0) * (m->d / 4);
  export_features_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(
      m->vsum.as<float>(), m->vcount.as<uint32_t>(), m->id_of_rank.as<uint32_t>(), r0, r1 - r0, m->d, out_dev);
  VSM_LAUNCHED();
  return VSM_OK;
}

extern "C" int vsm_num_contributor_entries(const vsm_map* m, int64_t* out_host) {
  VSM_TRY(require_finalized(m));
  if (!out_host) {
    set_error("null output");
    return VSM_E_INVALID;
  }
  *out_host = m->csr_entries;
  return VSM_OK;
}

extern "C" int vsm_export_contributors(const vsm_map* m, int64_t* offsets_dev, int32_t* submap_ids_dev,
                                       uint64_t* masks_dev, void* stream) {
  VSM_TRY(require_finalized(m));
  VSM_CUDA(cudaSetDevice(m->device));
  cudaStream_t s = (cudaStream_t)stream;
  const int64_t V = m->n_vox;
  if (offsets_dev) {
    widen_u32_kernel<<<grid_for(V + 1, 256), 256, 0, s>>>(m->csr_off.as<uint32_t>(), offsets_dev, V + 1);
    VSM_LAUNCHED();
  }
  if (m->csr_entries) {
    if (submap_ids_dev)
      VSM_CUDA(cudaMemcpyAsync(submap_ids_dev, m->csr_sub.p, (size_t)m->csr_entries * 4, cudaMemcpyDeviceToDevice, s));
    if (masks_dev)
      VSM_CUDA(cudaMemcpyAsync(masks_dev, m->csr_mask.p, (size_t)m->csr_entries * 16, cudaMemcpyDeviceToDevice, s));
  }
  return VSM_OK;
}

extern "C" int vsm_export_point_index(const vsm_map* m, int32_t fuse_index, int32_t* out_dev, int64_t n_pixels,
                                      void* stream) {
  VSM_TRY(require_finalized(m));
  if (fuse_index < 0 || fuse_index >= (int32_t)m->fuses.size() || !out_dev) {
    set_error("vsm_export_point_index: bad fuse index %d", fuse_index);
    return VSM_E_INVALID;
  }
  const FuseRecord& r = m->fuses[fuse_index];
  const int64_t n = (int64_t)r.S * r.H * r.W;
  if (r.point_gid.p == nullptr || n_pixels != n) {
    set_error("vsm_export_point_index: fuse call %d kept no point index (or size mismatch %lld vs %lld)", fuse_index,
              (long long)n_pixels, (long long)n);
    return VSM_E_STATE;
  }
  VSM_CUDA(cudaSetDevice(m->device));
  point_index_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(r.point_gid.as<int32_t>(),
                                                                         m->rank_of_id.as<uint32_t>(), n, out_dev);
  VSM_LAUNCHED();
  return VSM_OK;
}

extern "C" int vsm_export_packed_keys(const vsm_map* m, uint64_t* keys_dev, void* stream) {
  VSM_TRY(require_finalized(m));
  if (m->dense_loaded) {
    set_error("a dense-loaded map has no voxel keys");
    return VSM_E_STATE;
  }
  if (m->n_vox)
    VSM_CUDA(cudaMemcpyAsync(keys_dev, m->sorted_keys.p, (size_t)m->n_vox * 8, cudaMemcpyDeviceToDevice,
                             (cudaStream_t)stream));
  return VSM_OK;
}

// A loaded map (semantic_voxel.py:150-165) arrives in chunks so that a 50 M-voxel file is never resident twice:
// begin (sizes the map), rows (any number of row blocks, straight into the sum rows), end (= finalisation).
extern "C" int vsm_map_load_begin(vsm_map* m, int64_t V, void* stream) {
  if (!m || V < 0 || V >= ((int64_t)1 << 31)) {
    set_error("vsm_map_load_begin: bad arguments");
    return VSM_E_INVALID;
  }
  cudaStream_t s = (cudaStream_t)stream;
  VSM_TRY(vsm_map_clear(m, stream));
  VSM_TRY(map_grow(m, std::max<int64_t>(V, 1), s));
  m->dense_loaded = true;
  m->n_vox = V;
  if (V) {
    VSM_TRY(m->dense_centers.ensure((size_t)V * 12, s));
    fill_u32_kernel<<<grid_for(V, 256), 256, 0, s>>>(m->vcount.as<uint32_t>(), 1u, V);
    VSM_LAUNCHED();
    const uint32_t v32 = (uint32_t)V;
    VSM_CUDA(cudaMemcpyAsync(m->d_n_vox.p, &v32, 4, cudaMemcpyHostToDevice, s));
    VSM_CUDA(cudaStreamSynchronize(s));
  }
  return VSM_OK;
}

extern "C" int vsm_map_load_rows(vsm_map* m, int64_t r0, int64_t n_rows, const float* centers_dev, const float* features_dev,
                                 void* stream) {
  if (!m || !m->dense_loaded || m->finalized || r0 < 0 || n_rows < 0 || r0 + n_rows > m->n_vox ||
      (n_rows > 0 && (!centers_dev || !features_dev))) {
    set_error("vsm_map_load_rows: bad row block, or no vsm_map_load_begin before it");
    return m && (!m->dense_loaded || m->finalized) ? VSM_E_STATE : VSM_E_INVALID;
  }
  if (n_rows == 0) return VSM_OK;
  VSM_CUDA(cudaSetDevice(m->device));
  cudaStream_t s = (cudaStream_t)stream;
  VSM_CUDA(cudaMemcpyAsync(m->dense_centers.as<float>() + 3 * r0, centers_dev, (size_t)n_rows * 12,
                           cudaMemcpyDeviceToDevice, s));
  VSM_CUDA(cudaMemcpyAsync(m->vsum.as<float>() + (size_t)r0 * m->d, features_dev, (size_t)n_rows * m->d * 4,
                           cudaMemcpyDeviceToDevice, s));
  return VSM_OK;
}

extern "C" int vsm_map_load_dense(vsm_map* m, const float* centers_dev, const float* features_dev, int64_t V,
                                  void* stream) {
  if (!m || V < 0 || (V > 0 && (!centers_dev || !features_dev))) {
    set_error("vsm_map_load_dense: bad arguments");
    return VSM_E_INVALID;
  }
  VSM_TRY(vsm_map_load_begin(m, V, stream));
  VSM_TRY(vsm_map_load_rows(m, 0, V, centers_dev, features_dev, stream));
  return vsm_finalize(m, stream);
}

// ---- global voxel index of an owner shard ------------------------------------------------------------
// Every rank holds the sorted keys of ALL ranks (all-gathered, 8 bytes per voxel) and ranks its own keys among them:
// shards are disjoint, so the global index of a key is the number of smaller keys, one binary search per shard.
namespace vsm {
struct ShardSizes {
  long long n[64];
};
__global__ void __launch_bounds__(256) global_rank_kernel(const unsigned long long* __restrict__ mine, int64_t n_mine,
                                                          const unsigned long long* __restrict__ all, long long stride,
                                                          ShardSizes sz, int world, int64_t* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_mine; i += (int64_t)gridDim.x * blockDim.x) {
    const unsigned long long k = mine[i];
    long long rank = 0;
    for (int r = 0; r < world; ++r) {
      const unsigned long long* shard = all + (size_t)r * stride;
      long long lo = 0, hi = sz.n[r];
      while (lo < hi) {  // first position whose key is >= k
        const long long mid = (lo + hi) >> 1;
        if (shard[mid] < k)
          lo = mid + 1;
        else
          hi = mid;
      }
      rank += lo;
    }
    out[i] = rank;
  }
}
}  // namespace vsm

extern "C" int vsm_global_ranks(const uint64_t* my_keys_dev, int64_t n_mine, const uint64_t* all_keys_dev,
                                int64_t shard_stride, const int64_t* shard_sizes_host, int32_t world, int64_t* ranks_dev,
                                void* stream) {
  if (n_mine < 0 || world < 1 || world > 64 || !shard_sizes_host || shard_stride < 0 ||
      (n_mine > 0 && (!my_keys_dev || !all_keys_dev || !ranks_dev))) {
    set_error("vsm_global_ranks: bad arguments (world <= 64)");
    return VSM_E_INVALID;
  }
  if (n_mine == 0) return VSM_OK;
  ShardSizes sz{};
  for (int r = 0; r < world; ++r) {
    if (shard_sizes_host[r] < 0 || shard_sizes_host[r] > shard_stride) {
      set_error("vsm_global_ranks: shard %d holds %lld keys, stride %lld", r, (long long)shard_sizes_host[r], (long long)shard_stride);
      return VSM_E_INVALID;
    }
    sz.n[r] = shard_sizes_host[r];
  }
  global_rank_kernel<<<grid_for(n_mine, 256), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const unsigned long long*>(my_keys_dev), n_mine, reinterpret_cast<const unsigned long long*>(all_keys_dev),
      (long long)shard_stride, sz, world, ranks_dev);
  VSM_LAUNCHED();
  return VSM_OK;
}

extern "C" int vsm_lookup(vsm_map* m, const float* pos_dev, int64_t M, int64_t* idx_dev, int compat, void* stream) {
  VSM_TRY(require_finalized(m));
  if (!pos_dev || !idx_dev || M < 0) {
    set_error("vsm_lookup: bad arguments");
    return VSM_E_INVALID;
  }
  if (M == 0) return VSM_OK;
  VSM_CUDA(cudaSetDevice(m->device));
  cudaStream_t s = (cudaStream_t)stream;
  if (m->dense_loaded) compat = 1;  // a loaded map only has centres
  if (compat && !m->ck_built) {
    const uint32_t V = (uint32_t)m->n_vox;
    const uint64_t cap = next_pow2(std::max<uint64_t>(2 * (uint64_t)V, 1024));
    VSM_TRY(m->ck_keys.ensure(cap * 8, s));
    VSM_TRY(m->ck_val.ensure(cap * 4, s));
    m->ck_cap = cap;
    VSM_CUDA(cudaMemsetAsync(m->ck_keys.p, 0xFF, cap * 8, s));
    VSM_CUDA(cudaMemsetAsync(m->ck_val.p, 0, cap * 4, s));
    if (V) {
      compat_build_kernel<<<grid_for(V, 256), 256, 0, s>>>(m->sorted_keys.as<unsigned long long>(),
                                                           m->dense_loaded ? m->dense_centers.as<float>() : nullptr,
                                                           m->vs_f, V, m->ck_keys.as<unsigned long long>(),
                                                           m->ck_val.as<uint32_t>(), cap - 1);
      VSM_LAUNCHED();
    }
    m->ck_built = true;
  }
  if (compat)
    lookup_kernel<<<grid_for(M, 256), 256, 0, s>>>(pos_dev, M, m->vs_f, m->ck_keys.as<unsigned long long>(), m->ck_val.p,
                                                   m->ck_cap - 1, nullptr, 1, idx_dev);
  else
    lookup_kernel<<<grid_for(M, 256), 256, 0, s>>>(pos_dev, M, m->vs_f, m->gkeys.as<unsigned long long>(), m->gids.p,
                                                   m->gcap - 1, m->rank_of_id.as<uint32_t>(), 0, idx_dev);
  VSM_LAUNCHED();
  return VSM_OK;
}
