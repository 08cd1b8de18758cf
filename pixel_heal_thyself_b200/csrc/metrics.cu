// Validation metrics of the reference's validation loop (pht/models/base_trainer.py:549-571) on the GPU:
//   tensor2img   (pht/models/afgsa/util.py:77-119)   log-space NCHW fp32 -> tone-mapped uint8 NHWC
//   PSNR / SSIM  (pht/models/afgsa/metric.py:9-73)    on the uint8 images (integer MSE; 11x11 gaussian SSIM, fp64)
//   MRSE         (metric.py:76-94)                    0.5 * mean((a-b)^2 / (b^2 + 0.01)) on the linear radiance
// All reductions are two-stage with a fixed summation order (deterministic).
#include "common.cuh"

namespace pht {

// out[b][y][x][c] = uint8(clip(clip(v^(1/2.2), 0, 1) * 255, 0, 255)),  v = post_spec ? exp(x) - 1 : x  (NaN -> 0)
__global__ void tonemap_u8_kernel(const float* __restrict__ x, uint8_t* __restrict__ out, int B, int C, int H, int W, int post_spec) {
  const long long total = (long long)B * C * H * W;
  const float inv_gamma = (float)(1.0 / 2.2);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int px = (int)(i % W);
    long long r = i / W;
    const int py = (int)(r % H); r /= H;
    const int c = (int)(r % C), b = (int)(r / C);
    float v = x[i];
    if (post_spec) v = expf(v) - 1.0f;
    float t = powf(v, inv_gamma);
    t = fminf(fmaxf(t, 0.f), 1.f) * 255.0f;          // (fmaxf / fminf drop a NaN operand: NaN -> 0, like the uint8 cast)
    t = fminf(fmaxf(t, 0.f), 255.f);
    out[(((long long)b * H + py) * W + px) * C + c] = (uint8_t)t;   // truncation, like numpy's astype(uint8)
  }
}

// per-block partial sums of (a - b)^2 over one image: grid (blocks, B)
__global__ void sqdiff_u8_kernel(const uint8_t* __restrict__ a, const uint8_t* __restrict__ b, long long n,
                                 unsigned long long* __restrict__ part) {
  const uint8_t* pa = a + (long long)blockIdx.y * n;
  const uint8_t* pb = b + (long long)blockIdx.y * n;
  unsigned long long s = 0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int d = (int)pa[i] - (int)pb[i];
    s += (unsigned long long)(d * d);
  }
  __shared__ unsigned long long ws[8];
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long t = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += ws[w];
    part[(long long)blockIdx.y * gridDim.x + blockIdx.x] = t;
  }
}

// SSIM map sum of one uint8 HWC image pair: every valid pixel (5 <= y < H-5, 5 <= x < W-5) and channel, 11x11 gaussian
// window (sigma 1.5, normalised like cv2.getGaussianKernel), fp64.  grid (blocks, B); per-block partial sums.
__global__ void ssim_u8_kernel(const uint8_t* __restrict__ a, const uint8_t* __restrict__ b, int H, int W, int C,
                               double* __restrict__ part) {
  __shared__ double g[11];
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < 11; ++i) { g[i] = exp(-((i - 5) * (i - 5)) / (2.0 * 1.5 * 1.5)); s += g[i]; }
    for (int i = 0; i < 11; ++i) g[i] /= s;
  }
  __syncthreads();
  const uint8_t* pa = a + (long long)blockIdx.y * H * W * C;
  const uint8_t* pb = b + (long long)blockIdx.y * H * W * C;
  const int vh = H - 10, vw = W - 10;
  const long long n = (long long)vh * vw * C;
  const double c1 = (0.01 * 255) * (0.01 * 255), c2 = (0.03 * 255) * (0.03 * 255);
  double acc = 0.0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    long long r = i / C;
    const int x = (int)(r % vw) + 5, y = (int)(r / vw) + 5;
    double m1 = 0, m2 = 0, s11 = 0, s22 = 0, s12 = 0;
    for (int dy = -5; dy <= 5; ++dy) {
      const double gy = g[dy + 5];
      const long long row = ((long long)(y + dy) * W + (x - 5)) * C + c;
      for (int dx = 0; dx < 11; ++dx) {
        const double w = gy * g[dx];
        const double u = (double)pa[row + (long long)dx * C], v = (double)pb[row + (long long)dx * C];
        m1 += w * u; m2 += w * v; s11 += w * u * u; s22 += w * v * v; s12 += w * u * v;
      }
    }
    const double m11 = m1 * m1, m22 = m2 * m2, m12 = m1 * m2;
    acc += ((2 * m12 + c1) * (2 * (s12 - m12) + c2)) / ((m11 + m22 + c1) * ((s11 - m11) + (s22 - m22) + c2));
  }
  __shared__ double ws[8];
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += ws[w];
    part[(long long)blockIdx.y * gridDim.x + blockIdx.x] = t;
  }
}

// per-block partial sums of (a - b)^2 / (b^2 + 0.01), a = a_is_log ? exp(a) - 1 : a   (fp32 terms, fp64 accumulation)
__global__ void mrse_kernel(const float* __restrict__ a, const float* __restrict__ b, long long n, int a_is_log,
                            double* __restrict__ part) {
  const float* pa = a + (long long)blockIdx.y * n;
  const float* pb = b + (long long)blockIdx.y * n;
  double acc = 0.0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float u = pa[i];
    if (a_is_log) u = expf(u) - 1.0f;
    const float v = pb[i], d = u - v;
    acc += (double)((d * d) / (v * v + 1.0e-2f));
  }
  __shared__ double ws[8];
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += ws[w];
    part[(long long)blockIdx.y * gridDim.x + blockIdx.x] = t;
  }
}

// out[b] = sum of the image's block partials in block order
template <typename T>
__global__ void sum_partials_kernel(const T* __restrict__ part, int nblk, T* __restrict__ out) {
  if (threadIdx.x == 0) {
    T t = 0;
    for (int i = 0; i < nblk; ++i) t += part[(long long)blockIdx.x * nblk + i];
    out[blockIdx.x] = t;
  }
}

constexpr int METRIC_BLOCKS = 592;

}  // namespace pht

extern "C" {

using namespace pht;

int pht_tonemap_u8(const float* x_nchw, uint8_t* out_nhwc, int32_t B, int32_t C, int32_t H, int32_t W, int32_t post_spec,
                   void* stream) {
  PHT_CHECK_ARG(x_nchw && out_nhwc && B > 0 && C > 0 && H > 0 && W > 0, "tonemap_u8: bad args");
  const long long n = (long long)B * C * H * W;
  int grid = (int)((n + 255) / 256);
  if (grid > 148 * 16) grid = 148 * 16;
  tonemap_u8_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x_nchw, out_nhwc, B, C, H, W, post_spec);
  count_launch(CNT_OTHER);
  PHT_LAUNCH_CHECK();
  return PHT_OK;
}

size_t pht_image_metrics_ws_bytes(int32_t B) { return (size_t)(B > 0 ? B : 0) * METRIC_BLOCKS * 8 * 2 + 256; }

int pht_image_metrics_u8(const uint8_t* a, const uint8_t* b, int32_t B, int32_t H, int32_t W, int32_t C, uint64_t* sqdiff,
                         double* ssim_sum, void* workspace, size_t workspace_bytes, void* stream) {
  PHT_CHECK_ARG(a && b && sqdiff && ssim_sum && B > 0 && H > 10 && W > 10 && C > 0, "image_metrics_u8: bad args (H, W must exceed 10)");
  PHT_CHECK_ARG(workspace && workspace_bytes >= pht_image_metrics_ws_bytes(B) && ((uintptr_t)workspace & 7) == 0,
                "image_metrics_u8: workspace too small / misaligned");
  cudaStream_t st = (cudaStream_t)stream;
  unsigned long long* p1 = (unsigned long long*)workspace;
  double* p2 = (double*)(p1 + (size_t)B * METRIC_BLOCKS);
  dim3 grid(METRIC_BLOCKS, B);
  sqdiff_u8_kernel<<<grid, 256, 0, st>>>(a, b, (long long)H * W * C, p1);
  sum_partials_kernel<unsigned long long><<<B, 32, 0, st>>>(p1, METRIC_BLOCKS, (unsigned long long*)sqdiff);
  ssim_u8_kernel<<<grid, 256, 0, st>>>(a, b, H, W, C, p2);
  sum_partials_kernel<double><<<B, 32, 0, st>>>(p2, METRIC_BLOCKS, ssim_sum);
  count_launch(CNT_OTHER, 4);
  PHT_LAUNCH_CHECK();
  return PHT_OK;
}

int pht_mrse(const float* a, const float* b, int32_t B, int64_t n_per_image, int32_t a_is_log, double* out_sum, void* workspace,
             size_t workspace_bytes, void* stream) {
  PHT_CHECK_ARG(a && b && out_sum && B > 0 && n_per_image > 0, "mrse: bad args");
  PHT_CHECK_ARG(workspace && workspace_bytes >= pht_image_metrics_ws_bytes(B) && ((uintptr_t)workspace & 7) == 0,
                "mrse: workspace too small / misaligned");
  cudaStream_t st = (cudaStream_t)stream;
  double* p = (double*)workspace;
  dim3 grid(METRIC_BLOCKS, B);
  mrse_kernel<<<grid, 256, 0, st>>>(a, b, (long long)n_per_image, a_is_log, p);
  sum_partials_kernel<double><<<B, 32, 0, st>>>(p, METRIC_BLOCKS, out_sum);
  count_launch(CNT_OTHER, 2);
  PHT_LAUNCH_CHECK();
  return PHT_OK;
}

}  // extern "C"
