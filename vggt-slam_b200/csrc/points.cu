// points.cu -- world-frame point extraction and the confidence threshold.
//
//   vsm_conf_threshold     Submap.add_all_points: np.percentile(conf, pct)           vggt_slam/submap.py:34-39
//   vsm_transform_points   (H @ [p;1]) / w                                           vggt_slam/submap.py:171-174
//   vsm_select_points      Submap.get_points_in_world_frame / get_points_colors      vggt_slam/submap.py:155-164,
//                                                                                    182-188, 217-219
#include <cub/device/device_scan.cuh>

#include "state.cuh"

namespace vsm {

__global__ void __launch_bounds__(256) transform_kernel(const float* __restrict__ pts, int64_t n, HMat Hm,
                                                        double* __restrict__ out64, float* __restrict__ out32) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    double x, y, z;
    transform_f64(Hm, pts[3 * i], pts[3 * i + 1], pts[3 * i + 2], x, y, z);
    if (out64) {
      out64[3 * i] = x;
      out64[3 * i + 1] = y;
      out64[3 * i + 2] = z;
    } else {
      out32[3 * i] = __double2float_rn(x);
      out32[3 * i + 1] = __double2float_rn(y);
      out32[3 * i + 2] = __double2float_rn(z);
    }
  }
}

// selection flags on the strided grid, in (s, h/stride, w/stride) order
__global__ void __launch_bounds__(256) select_flags_kernel(const float* __restrict__ conf, int S, int H, int W, int stride,
                                                           int hs, int ws, float thr, uint32_t* __restrict__ flags) {
  const int64_t n = (int64_t)S * hs * ws;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int w = (int)(i % ws), h = (int)((i / ws) % hs);
    const int64_t s = i / ((int64_t)ws * hs);
    const int64_t pix = (s * H + (int64_t)h * stride) * W + (int64_t)w * stride;
    flags[i] = conf[pix] >= thr ? 1u : 0u;
  }
}

__global__ void __launch_bounds__(256) select_gather_kernel(const float* __restrict__ pts, const uint8_t* __restrict__ colors,
                                                            int S, int H, int W, int stride, int hs, int ws,
                                                            const uint32_t* __restrict__ flags,
                                                            const uint32_t* __restrict__ offs, HMat Hm,
                                                            double* __restrict__ out_world,
                                                            uint8_t* __restrict__ out_colors) {
  const int64_t n = (int64_t)S * hs * ws;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    if (!flags[i]) continue;
    const int w = (int)(i % ws), h = (int)((i / ws) % hs);
    const int64_t s = i / ((int64_t)ws * hs);
    const int64_t pix = (s * H + (int64_t)h * stride) * W + (int64_t)w * stride;
    const int64_t o = offs[i];
    if (out_world) {
      double x, y, z;
      transform_f64(Hm, pts[3 * pix], pts[3 * pix + 1], pts[3 * pix + 2], x, y, z);
      out_world[3 * o] = x;
      out_world[3 * o + 1] = y;
      out_world[3 * o + 2] = z;
    }
    if (out_colors) {
      out_colors[3 * o] = colors[3 * pix];
      out_colors[3 * o + 1] = colors[3 * pix + 1];
      out_colors[3 * o + 2] = colors[3 * pix + 2];
    }
  }
}

}  // namespace vsm

using namespace vsm;

extern "C" int vsm_conf_threshold(const float* conf_dev, int64_t n, double percentile, float* out_host, void* stream) {
  if (!conf_dev || !out_host || n <= 0) {
    set_error("vsm_conf_threshold: null pointer or empty input");
    return VSM_E_INVALID;
  }
  cudaStream_t s = (cudaStream_t)stream;
  int dev = 0;
  VSM_CUDA(cudaGetDevice(&dev));
  Workspace* ws = workspace_for_device(dev);
  if (!ws) {
    set_error("vsm_conf_threshold: device ordinal %d not supported", dev);
    return VSM_E_INVALID;
  }
  WsLease lease(ws, s);  // the select scratch is shared with the fuse calls of this device
  VSM_TRY(lease.status());
  SelectState* st;
  uint32_t* hist;
  float* out_dev;
  VSM_TRY(select_scratch(&st, &hist, &out_dev));
  SelSrc src;
  src.base = conf_dev;
  src.stride = 1;
  src.ncol = 1;
  src.flag_off = -1;
  src.flag_need = 0;
  src.n_items = n;
  const float q = (float)percentile / 100.0f;
  VSM_TRY(run_percentiles(st, hist, src, 1, q, q, out_dev, s));
  VSM_CUDA(cudaMemcpyAsync(out_host, out_dev, sizeof(float), cudaMemcpyDeviceToHost, s));
  VSM_CUDA(cudaStreamSynchronize(s));
  return VSM_OK;
}

extern "C" int vsm_transform_points(const float* pts_dev, int64_t n, const double* H_host16, void* out_dev, int out_f64,
                                    void* stream) {
  if (!pts_dev || !H_host16 || !out_dev || n < 0) {
    set_error("vsm_transform_points: null pointer");
    return VSM_E_INVALID;
  }
  if (n == 0) return VSM_OK;
  HMat Hm;
  for (int i = 0; i < 16; ++i) Hm.m[i] = H_host16[i];
  transform_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(pts_dev, n, Hm, out_f64 ? (double*)out_dev : nullptr,
                                                                       out_f64 ? nullptr : (float*)out_dev);
  VSM_LAUNCHED();
  return VSM_OK;
}

extern "C" int vsm_select_points(const float* pts_dev, const float* conf_dev, const uint8_t* colors_dev, int32_t S,
                                 int32_t H, int32_t W, int32_t stride, float conf_threshold, const double* H_host16,
                                 double* out_world_dev, uint8_t* out_colors_dev, int64_t* n_selected_host,
                                 void* stream) {
  if (!conf_dev || !n_selected_host || S < 0 || H <= 0 || W <= 0 || stride < 1) {
    set_error("vsm_select_points: bad arguments");
    return VSM_E_INVALID;
  }
  if ((out_world_dev && (!pts_dev || !H_host16)) || (out_colors_dev && !colors_dev)) {
    set_error("vsm_select_points: output requested without its input");
    return VSM_E_INVALID;
  }
  cudaStream_t s = (cudaStream_t)stream;
  const int hs = (H + stride - 1) / stride, ws = (W + stride - 1) / stride;
  const int64_t n = (int64_t)S * hs * ws;
  *n_selected_host = 0;
  if (n == 0) return VSM_OK;
  if (n >= ((int64_t)1 << 31)) {
    set_error("vsm_select_points: too many pixels");
    return VSM_E_INVALID;
  }
  DevBuf b_flags, b_offs, b_tmp;
  VSM_TRY(b_flags.ensure((size_t)n * 4, s));
  VSM_TRY(b_offs.ensure(((size_t)n + 1) * 4, s));
  uint32_t* flags = b_flags.as<uint32_t>();
  uint32_t* offs = b_offs.as<uint32_t>();
  size_t tmp_bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, flags, offs, (int)n, s);
  VSM_TRY(b_tmp.ensure(tmp_bytes + 16, s));
  void* tmp = b_tmp.p;
  int status = VSM_OK;
  do {
    select_flags_kernel<<<grid_for(n, 256), 256, 0, s>>>(conf_dev, S, H, W, stride, hs, ws, conf_threshold, flags);
    ++g_launches;
    if (cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, flags, offs, (int)n, s) != cudaSuccess) {
      status = VSM_E_CUDA;
      break;
    }
    ++g_launches;
    HMat Hm{};
    if (H_host16)
      for (int i = 0; i < 16; ++i) Hm.m[i] = H_host16[i];
    if (out_world_dev || out_colors_dev) {
      select_gather_kernel<<<grid_for(n, 256), 256, 0, s>>>(pts_dev, colors_dev, S, H, W, stride, hs, ws, flags, offs, Hm,
                                                            out_world_dev, out_colors_dev);
      ++g_launches;
    }
    uint32_t last_off = 0, last_flag = 0;
    if (cudaMemcpyAsync(&last_off, offs + (n - 1), 4, cudaMemcpyDeviceToHost, s) != cudaSuccess ||
        cudaMemcpyAsync(&last_flag, flags + (n - 1), 4, cudaMemcpyDeviceToHost, s) != cudaSuccess ||
        cudaStreamSynchronize(s) != cudaSuccess) {
      status = VSM_E_CUDA;
      break;
    }
    *n_selected_host = (int64_t)last_off + last_flag;
  } while (0);
  if (status != VSM_OK) set_error("vsm_select_points: %s", cudaGetErrorString(cudaGetLastError()));
  cudaStreamSynchronize(s);
  b_flags.release();
  b_offs.release();
  b_tmp.release();
  return status;
}
