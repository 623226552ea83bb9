// ransac.cu -- scoring of the projective RANSAC hypotheses.
//
// Reference: vggt_slam/h_solve.py
//   apply_homography_batch   :16-41    X_trans = bmm(H_batch, [X;1]^T) in float32, divide by the w row
//   ransac_projective        :150-160  errors = ||X2_pred - X2||_2, inlier_counts = (errors < threshold).sum(1),
//                                      best = argmax(inlier_counts)
// The reference materialises three (B, N, 3) float32 tensors (B = 300 hypotheses, N = 152 k points of a 518x294 frame:
// 1.6 GB of traffic); here a CTA keeps 256 point pairs in registers, walks all hypotheses from shared memory and only
// the B counters leave the chip.  The minimal-sample estimation (null space of a 15x16 system per hypothesis,
// :43-93) stays on the host, where the reference runs it too (numpy / scipy).
#include "state.cuh"

namespace vsm {

constexpr int kRansacHypsPerPass = 512;  // 32 KB of shared memory

// float32 arithmetic of the reference, one operation after the other: the dot products as an FMA chain over k
// (what a float32 GEMM with K = 4 does), IEEE division, sum of squares, IEEE square root.
__global__ void __launch_bounds__(256) ransac_score_kernel(const float* __restrict__ Hs, const float* __restrict__ X1,
                                                           const float* __restrict__ X2, int64_t N, int B, float thr,
                                                           int32_t* __restrict__ counts) {
  __shared__ float sH[kRansacHypsPerPass * 16];
  __shared__ int sCnt[kRansacHypsPerPass];
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool on = i < N;
  float x = 0.f, y = 0.f, z = 0.f, tx = 0.f, ty = 0.f, tz = 0.f;
  if (on) {
    x = X1[3 * i];
    y = X1[3 * i + 1];
    z = X1[3 * i + 2];
    tx = X2[3 * i];
    ty = X2[3 * i + 1];
    tz = X2[3 * i + 2];
  }
  for (int b0 = 0; b0 < B; b0 += kRansacHypsPerPass) {
    const int nb = min(kRansacHypsPerPass, B - b0);
    __syncthreads();
    for (int k = threadIdx.x; k < nb * 16; k += blockDim.x) sH[k] = Hs[(size_t)b0 * 16 + k];
    for (int k = threadIdx.x; k < nb; k += blockDim.x) sCnt[k] = 0;
    __syncthreads();
    for (int b = 0; b < nb; ++b) {
      const float* h = sH + b * 16;
      float r[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        float acc = __fmul_rn(h[4 * a], x);
        acc = __fmaf_rn(h[4 * a + 1], y, acc);
        acc = __fmaf_rn(h[4 * a + 2], z, acc);
        acc = __fmaf_rn(h[4 * a + 3], 1.0f, acc);
        r[a] = acc;
      }
      const float dx = __fsub_rn(__fdiv_rn(r[0], r[3]), tx);
      const float dy = __fsub_rn(__fdiv_rn(r[1], r[3]), ty);
      const float dz = __fsub_rn(__fdiv_rn(r[2], r[3]), tz);
      float s = __fmul_rn(dx, dx);
      s = __fmaf_rn(dy, dy, s);
      s = __fmaf_rn(dz, dz, s);
      const bool inl = on && (__fsqrt_rn(s) < thr);  // NaN (w = 0) compares false, as in the reference
      const unsigned m = __ballot_sync(0xffffffffu, inl);
      if (lane_id() == 0 && m) atomicAdd(&sCnt[b], __popc(m));
    }
    __syncthreads();
    for (int k = threadIdx.x; k < nb; k += blockDim.x)
      if (sCnt[k]) atomicAdd(&counts[b0 + k], sCnt[k]);
  }
}

// torch.argmax: the first index holding the maximum
__global__ void __launch_bounds__(256) argmax_first_kernel(const int32_t* __restrict__ counts, int B, int32_t* __restrict__ out) {
  __shared__ long long best[256];
  long long mine = -1;
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    const long long key = ((long long)counts[b] << 32) | (long long)(0x7FFFFFFF - b);  // larger count, then smaller index
    if (key > mine) mine = key;
  }
  best[threadIdx.x] = mine;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o && best[threadIdx.x + o] > best[threadIdx.x]) best[threadIdx.x] = best[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    out[0] = 0x7FFFFFFF - (int32_t)(best[0] & 0xFFFFFFFFll);
    out[1] = (int32_t)(best[0] >> 32);
  }
}

}  // namespace vsm

using namespace vsm;

extern "C" int vsm_ransac_score(const float* H_dev, const float* X1_dev, const float* X2_dev, int64_t N, int32_t B,
                                float threshold, int32_t* counts_dev, int32_t* best_dev, int32_t* best_idx_host,
                                int32_t* best_count_host, void* stream) {
  if (!H_dev || !X1_dev || !X2_dev || !counts_dev || !best_dev || N < 0 || B < 1) {
    set_error("vsm_ransac_score: bad arguments (B >= 1, non-null device pointers)");
    return VSM_E_INVALID;
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    set_error("no CUDA device: libvsm has no CPU fallback");
    return VSM_E_CUDA;
  }
  cudaStream_t s = (cudaStream_t)stream;
  VSM_CUDA(cudaMemsetAsync(counts_dev, 0, (size_t)B * sizeof(int32_t), s));
  if (N > 0) {
    ransac_score_kernel<<<(unsigned)cdiv(N, 256), 256, 0, s>>>(H_dev, X1_dev, X2_dev, N, B, threshold, counts_dev);
    VSM_LAUNCHED();
  }
  argmax_first_kernel<<<1, 256, 0, s>>>(counts_dev, B, best_dev);
  VSM_LAUNCHED();
  if (best_idx_host || best_count_host) {
    int32_t host[2] = {0, 0};
    VSM_CUDA(cudaMemcpyAsync(host, best_dev, sizeof(host), cudaMemcpyDeviceToHost, s));
    VSM_CUDA(cudaStreamSynchronize(s));
    if (best_idx_host) *best_idx_host = host[0];
    if (best_count_host) *best_count_host = host[1];
  }
  return VSM_OK;
}
