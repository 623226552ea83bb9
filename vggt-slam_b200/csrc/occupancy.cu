// occupancy.cu -- 2-D occupancy grid of a point cloud, the same unique + scatter-min/max pattern
// as the voxel map: an open-addressing hash of the (ix, iy) cells with atomic min / max of z, then one sort of the
// DISTINCT cells.
//
// Reference: get_occupancy.py:130-179 build_occupancy_from_pointcloud
//   pts = float32 points, rows with a non-finite value dropped, rows with z > ceiling_z dropped      (:145-148)
//   ix, iy = floor(x / voxel_size), floor(y / voxel_size) in float32 -> int64                       (:158-159)
//   cells = np.unique(axis=0) (lexicographic signed order); min / max of z per cell                 (:161-168)
//   blocked = (maxz - minz) > height_thresh in float32                                              (:170-172)
//   centre = ((ix + 0.5) * vs, (iy + 0.5) * vs, minz + vs * 0.5) in float32                         (:174-178)
#include <cub/device/device_radix_sort.cuh>

#include "state.cuh"

namespace vsm {

struct alignas(16) CellSlot {
  unsigned long long key;  // biased (ix, iy): unsigned order == lexicographic signed order; kEmptyKey when free
  uint32_t zmin, zmax;     // order-preserving images of the float32 heights
};

__global__ void __launch_bounds__(256) cell_init_kernel(CellSlot* t, uint64_t n) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    t[i].key = kEmptyKey;
    t[i].zmin = 0xFFFFFFFFu;
    t[i].zmax = 0u;
  }
}

struct OccCounters {
  unsigned long long n_kept;
  uint32_t n_cells, range_err;
};

__global__ void __launch_bounds__(256) cell_insert_kernel(const float* __restrict__ pts, int64_t n, float vs, float ceiling,
                                                          CellSlot* __restrict__ t, uint64_t mask,
                                                          unsigned long long* __restrict__ cell_list, OccCounters* ctr) {
  unsigned kept = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float x = pts[3 * i], y = pts[3 * i + 1], z = pts[3 * i + 2];
    if (!finite3(x, y, z) || !(z <= ceiling)) continue;
    ++kept;
    const float qx = floorf(__fdiv_rn(x, vs)), qy = floorf(__fdiv_rn(y, vs));
    if (!(fabsf(qx) < 2147483648.0f) || !(fabsf(qy) < 2147483648.0f)) {
      atomicAdd(&ctr->range_err, 1u);
      continue;
    }
    const unsigned long long key = ((unsigned long long)((uint32_t)(int)qx ^ 0x80000000u) << 32) |
                                   (unsigned long long)((uint32_t)(int)qy ^ 0x80000000u);
    const uint32_t zo = float_to_ordered(z);
    uint64_t h = mix64(key) & mask;
    while (true) {
      unsigned long long cur = t[h].key;
      if (cur == kEmptyKey) {
        cur = atomicCAS(&t[h].key, kEmptyKey, key);
        if (cur == kEmptyKey) {
          cell_list[atomicAdd(&ctr->n_cells, 1u)] = key;
          cur = key;
        }
      }
      if (cur == key) {
        atomicMin(&t[h].zmin, zo);
        atomicMax(&t[h].zmax, zo);
        break;
      }
      h = (h + 1) & mask;
    }
  }
  for (int o = 16; o > 0; o >>= 1) kept += __shfl_xor_sync(0xffffffffu, kept, o);
  if (lane_id() == 0 && kept) atomicAdd(&ctr->n_kept, (unsigned long long)kept);
}

__global__ void __launch_bounds__(256) cell_export_kernel(const unsigned long long* __restrict__ sorted, uint32_t n_cells,
                                                          const CellSlot* __restrict__ t, uint64_t mask, float vs,
                                                          float half_vs, float thresh, float* __restrict__ centers,
                                                          uint8_t* __restrict__ blocked, int64_t* __restrict__ keys,
                                                          float* __restrict__ minz) {
  for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < n_cells; r += gridDim.x * blockDim.x) {
    const unsigned long long key = sorted[r];
    uint64_t h = mix64(key) & mask;
    while (t[h].key != key) h = (h + 1) & mask;
    const float lo = ordered_to_float(t[h].zmin), hi = ordered_to_float(t[h].zmax);
    const int ix = (int)((uint32_t)(key >> 32) ^ 0x80000000u), iy = (int)((uint32_t)key ^ 0x80000000u);
    if (centers) {
      centers[3 * (size_t)r] = __fmul_rn(__fadd_rn((float)ix, 0.5f), vs);
      centers[3 * (size_t)r + 1] = __fmul_rn(__fadd_rn((float)iy, 0.5f), vs);
      centers[3 * (size_t)r + 2] = __fadd_rn(lo, half_vs);
    }
    if (blocked) blocked[r] = __fsub_rn(hi, lo) > thresh ? 1 : 0;
    if (keys) {
      keys[2 * (size_t)r] = ix;
      keys[2 * (size_t)r + 1] = iy;
    }
    if (minz) minz[r] = lo;
  }
}

}  // namespace vsm

using namespace vsm;

extern "C" int vsm_occupancy_build(const float* pts_dev, int64_t n, double voxel_size, double ceiling_z,
                                   double height_thresh, int64_t cap_cells, float* centers_dev, uint8_t* blocked_dev,
                                   int64_t* keys_dev, float* minz_dev, int64_t* n_cells_host, int64_t* n_kept_host,
                                   void* stream) {
  if (n < 0 || (n > 0 && !pts_dev) || !(voxel_size > 0.0) || cap_cells < 0 || !n_cells_host) {
    set_error("vsm_occupancy_build: bad arguments");
    return VSM_E_INVALID;
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    set_error("no CUDA device: libvsm has no CPU fallback");
    return VSM_E_CUDA;
  }
  *n_cells_host = 0;
  if (n_kept_host) *n_kept_host = 0;
  if (n == 0) return VSM_OK;
  if (n >= ((int64_t)1 << 31)) {
    set_error("vsm_occupancy_build: more than 2^31 points");
    return VSM_E_INVALID;
  }
  cudaStream_t s = (cudaStream_t)stream;
  const uint64_t cap = next_pow2(std::max<uint64_t>(2 * (uint64_t)n, 1024));
  DevBuf table, list, sorted, tmp, ctr;
  int status = VSM_OK;
  OccCounters hc{};
  do {
    if ((status = table.ensure(cap * sizeof(CellSlot), s)) != VSM_OK) break;
    if ((status = list.ensure((size_t)n * 8, s)) != VSM_OK) break;
    if ((status = sorted.ensure((size_t)n * 8, s)) != VSM_OK) break;
    if ((status = ctr.ensure(sizeof(OccCounters), s)) != VSM_OK) break;
    cudaMemsetAsync(ctr.p, 0, sizeof(OccCounters), s);
    cell_init_kernel<<<grid_for((int64_t)cap, 256), 256, 0, s>>>(table.as<CellSlot>(), cap);
    cell_insert_kernel<<<grid_for(n, 256), 256, 0, s>>>(pts_dev, n, (float)voxel_size, (float)ceiling_z, table.as<CellSlot>(),
                                                        cap - 1, list.as<unsigned long long>(), ctr.as<OccCounters>());
    g_launches += 2;
    if (cudaMemcpyAsync(&hc, ctr.p, sizeof(hc), cudaMemcpyDeviceToHost, s) != cudaSuccess ||
        cudaStreamSynchronize(s) != cudaSuccess) {
      status = VSM_E_CUDA;
      break;
    }
    if (hc.range_err) {
      set_error("vsm_occupancy_build: %u points fall into cells beyond +-2^31", hc.range_err);
      status = VSM_E_COORD_RANGE;
      break;
    }
    *n_cells_host = hc.n_cells;
    if (n_kept_host) *n_kept_host = (int64_t)hc.n_kept;
    if (hc.n_cells == 0) break;
    if ((int64_t)hc.n_cells > cap_cells) {
      if (centers_dev || blocked_dev || keys_dev || minz_dev) {
        set_error("vsm_occupancy_build: %u cells, room for %lld", hc.n_cells, (long long)cap_cells);
        status = VSM_E_NOMEM;
      }
      break;  // size query
    }
    size_t tmp_bytes = 0;
    cub::DeviceRadixSort::SortKeys(nullptr, tmp_bytes, list.as<unsigned long long>(), sorted.as<unsigned long long>(),
                                   (int)hc.n_cells, 0, 64, s);
    if ((status = tmp.ensure(tmp_bytes + 16, s)) != VSM_OK) break;
    if (cub::DeviceRadixSort::SortKeys(tmp.p, tmp_bytes, list.as<unsigned long long>(), sorted.as<unsigned long long>(),
                                       (int)hc.n_cells, 0, 64, s) != cudaSuccess) {
      status = VSM_E_CUDA;
      break;
    }
    // float(voxel_size) * 0.5 is a Python float product, rounded to float32 when it meets the float32 heights
    cell_export_kernel<<<grid_for(hc.n_cells, 256), 256, 0, s>>>(sorted.as<unsigned long long>(), hc.n_cells,
                                                                table.as<CellSlot>(), cap - 1, (float)voxel_size,
                                                                (float)(voxel_size * 0.5), (float)height_thresh,
                                                                centers_dev, blocked_dev, keys_dev, minz_dev);
    g_launches += 2;
    if (cudaStreamSynchronize(s) != cudaSuccess) status = VSM_E_CUDA;
  } while (0);
  if (status == VSM_E_CUDA) set_error("vsm_occupancy_build: %s", cudaGetErrorString(cudaGetLastError()));
  cudaStreamSynchronize(s);
  table.release();
  list.release();
  sorted.release();
  tmp.release();
  ctr.release();
  return status;
}
