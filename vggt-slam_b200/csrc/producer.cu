// producer.cu -- the producer hand-off: what Solver.add_points does to the VGGT predictions before
// Submap.add_all_points, on the device, so that the point maps never make the round trip through host numpy
// (vggt_slam/solver.py:478-480 copies every prediction to the host; :249-263, 337-340 prepare and store them).
//
//   vsm_unproject_depth   depth map -> point map (solver.py:254-256).  The reference calls
//                         vggt.utils.geometry.unproject_depth_map_to_point_map, a third-party dependency that is not
//                         vendored in the reference repository (requirements.txt installs facebookresearch/vggt from
//                         git, unpinned); its published algorithm is restated here and in oracle/producer_oracle.py:
//                           x_cam = (u - cu) * depth / fu,  y_cam = (v - cv) * depth / fv,  z_cam = depth   (float64,
//                           stored as float32);  world = cam @ R_c2w^T + t_c2w  (float64),  [R_c2w | t_c2w] the
//                           closed-form inverse of the 3x4 extrinsic.  PARITY UNPINNED for this entry point.
//   vsm_images_to_colors  (images.transpose(0,2,3,1) * 255).astype(uint8)  (solver.py:260)
//   vsm_scale_points      world_points *= scale_factor  (solver.py:301, Sim(3) mode)
#include "state.cuh"

namespace vsm {

__global__ void __launch_bounds__(256) unproject_kernel(const float* __restrict__ depth, const double* __restrict__ c2w,
                                                        const float* __restrict__ K, int S, int H, int W,
                                                        float* __restrict__ out32, double* __restrict__ out64) {
  const int64_t n = (int64_t)S * H * W;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int u = (int)(i % W), v = (int)((i / W) % H);
    const int s = (int)(i / ((int64_t)W * H));
    const float* k = K + 9 * s;
    const double* m = c2w + 12 * s;  // row-major 3x4: R_c2w | t_c2w
    const double dz = (double)depth[i];
    // numpy: int64 grid minus a float32 scalar is float64; times the float32 depth, over the float32 focal length
    const double xc64 = __ddiv_rn(__dmul_rn(__dsub_rn((double)u, (double)k[2]), dz), (double)k[0]);
    const double yc64 = __ddiv_rn(__dmul_rn(__dsub_rn((double)v, (double)k[5]), dz), (double)k[4]);
    const double xc = (double)__double2float_rn(xc64), yc = (double)__double2float_rn(yc64), zc = dz;  // .astype(float32)
    double w[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      double acc = __dmul_rn(xc, m[4 * a]);
      acc = __fma_rn(yc, m[4 * a + 1], acc);
      acc = __fma_rn(zc, m[4 * a + 2], acc);
      w[a] = __dadd_rn(acc, m[4 * a + 3]);
    }
    if (out64) {
      out64[3 * i] = w[0];
      out64[3 * i + 1] = w[1];
      out64[3 * i + 2] = w[2];
    } else {
      out32[3 * i] = __double2float_rn(w[0]);
      out32[3 * i + 1] = __double2float_rn(w[1]);
      out32[3 * i + 2] = __double2float_rn(w[2]);
    }
  }
}

__global__ void __launch_bounds__(256) images_to_colors_kernel(const float* __restrict__ img, int S, int H, int W,
                                                               uint8_t* __restrict__ out) {
  const int64_t n = (int64_t)S * H * W;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t s = i / ((int64_t)H * W), hw = i % ((int64_t)H * W);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float v = __fmul_rn(img[(s * 3 + c) * (int64_t)H * W + hw], 255.0f);
      out[3 * i + c] = (uint8_t)(int)v;  // C cast of the float, as ndarray.astype(uint8) does for in-range values
    }
  }
}

__global__ void __launch_bounds__(256) scale_points_kernel(float* __restrict__ p, int64_t n, double scale) {
  // numpy: float32 array *= np.float64 scalar (np.mean's result) -> the product is formed in float64 and stored float32
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    p[i] = __double2float_rn(__dmul_rn((double)p[i], scale));
}

static int need_device() {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    set_error("no CUDA device: libvsm has no CPU fallback");
    return VSM_E_CUDA;
  }
  return VSM_OK;
}

}  // namespace vsm

using namespace vsm;

extern "C" int vsm_unproject_depth(const float* depth_dev, const double* cam_to_world_dev, const float* intrinsic_dev,
                                   int32_t S, int32_t H, int32_t W, void* out_dev, int out_f64, void* stream) {
  if (!depth_dev || !cam_to_world_dev || !intrinsic_dev || !out_dev || S < 0 || H <= 0 || W <= 0) {
    set_error("vsm_unproject_depth: bad arguments");
    return VSM_E_INVALID;
  }
  VSM_TRY(need_device());
  const int64_t n = (int64_t)S * H * W;
  if (n == 0) return VSM_OK;
  unproject_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(depth_dev, cam_to_world_dev, intrinsic_dev, S, H, W,
                                                                       out_f64 ? nullptr : (float*)out_dev,
                                                                       out_f64 ? (double*)out_dev : nullptr);
  VSM_LAUNCHED();
  return VSM_OK;
}

extern "C" int vsm_images_to_colors(const float* images_dev, int32_t S, int32_t H, int32_t W, uint8_t* colors_dev,
                                    void* stream) {
  if (!images_dev || !colors_dev || S < 0 || H <= 0 || W <= 0) {
    set_error("vsm_images_to_colors: bad arguments");
    return VSM_E_INVALID;
  }
  VSM_TRY(need_device());
  const int64_t n = (int64_t)S * H * W;
  if (n == 0) return VSM_OK;
  images_to_colors_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(images_dev, S, H, W, colors_dev);
  VSM_LAUNCHED();
  return VSM_OK;
}

extern "C" int vsm_scale_points(float* pts_dev, int64_t n_floats, double scale, void* stream) {
  if (!pts_dev || n_floats < 0) {
    set_error("vsm_scale_points: bad arguments");
    return VSM_E_INVALID;
  }
  VSM_TRY(need_device());
  if (n_floats == 0) return VSM_OK;
  scale_points_kernel<<<grid_for(n_floats, 256), 256, 0, (cudaStream_t)stream>>>(pts_dev, n_floats, scale);
  VSM_LAUNCHED();
  return VSM_OK;
}
