// Importance map of the patch sampler (reference: pht/models/afgsa/preprocessing.py:119-168, 293-300):
//   imp = (V_rel(noisy) + V(normal)) / max(.),  V(.) = min(max_c(var_c)^(1/2.2), 1) / max(.),
//   var = max(E[x^2] - E[x]^2, 0) (/ max(E[x]^2, 1e-4) for the relative variant), E = P x P box filter
// E is scipy.ndimage.uniform_filter(size=(P, P, 1)) (scipy 1.15.2, third party): separable, axis 0 then axis 1, window
// [i - P/2, i - P/2 + P), 'reflect' borders (d c b a | a b c d | d c b a), double accumulation, float32 results after
// every pass.  The kernels keep exactly those rounding points (double sliding sums -> float32), so the map agrees with
// the reference to the last bit except where a double sum lands on a float32 rounding boundary.
// The frames are the raw HBM-resident "EXR" frames; preprocess_data's cleaning (nan_to_num, radiance clipped at 0,
// preprocessing.py:97-103) is applied on the fly.
#include <float.h>
#include <math.h>

#include "common.cuh"

namespace pht {

constexpr int IMP_SEG = 32;   // outputs per thread along the filtered axis
constexpr int IMP_CH = 12;    // 6 channels (noisy rgb, normal xyz) x {x, x^2}

__device__ __forceinline__ float nan_to_num(float v) {
  if (isnan(v)) return 0.f;
  if (isinf(v)) return v > 0.f ? FLT_MAX : -FLT_MAX;
  return v;
}
__device__ __forceinline__ int reflect_idx(int i, int n) {
  if (i < 0) i = -i - 1;
  if (i >= n) i = 2 * n - 1 - i;
  return i;
}

// pass 1 (axis 0 = y): thread = (image, y-segment, x, channel c6); both the value and its float32 square are filtered
__global__ void imp_box_y_kernel(const float* __restrict__ noisy, const float* __restrict__ aux, int n_img, int H, int W, int P,
                                 float* __restrict__ t1) {
  const int nseg = (H + IMP_SEG - 1) / IMP_SEG;
  const long long total = (long long)n_img * nseg * W * 6;
  const int h = P / 2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c6 = (int)(i % 6);
    long long r = i / 6;
    const int x = (int)(r % W); r /= W;
    const int seg = (int)(r % nseg);
    const int img = (int)(r / nseg);
    auto in = [&](int y) -> float {
      const long long p = ((long long)img * H + reflect_idx(y, H)) * W + x;
      if (c6 < 3) return fmaxf(nan_to_num(noisy[p * 3 + c6]), 0.f);
      return nan_to_num(aux[p * 7 + (c6 - 3)]);
    };
    const int y0 = seg * IMP_SEG, y1 = min(y0 + IMP_SEG, H);
    double s = 0.0, s2 = 0.0;
    for (int j = y0 - h; j < y0 - h + P; ++j) {
      const float v = in(j);
      s += (double)v;
      s2 += (double)(v * v);      // buffer ** 2 is a float32 array in the reference
    }
    for (int y = y0; y < y1; ++y) {
      float* o = t1 + (((long long)img * H + y) * W + x) * IMP_CH;
      o[c6] = (float)(s / (double)P);
      o[6 + c6] = (float)(s2 / (double)P);
      const float a = in(y - h + P), b = in(y - h);
      s += (double)a - (double)b;
      s2 += (double)(a * a) - (double)(b * b);
    }
  }
}

// pass 2 (axis 1 = x) on the float32 results of pass 1: thread = (image, y, x-segment, channel of 12)
__global__ void imp_box_x_kernel(const float* __restrict__ t1, int n_img, int H, int W, int P, float* __restrict__ t2) {
  const int nseg = (W + IMP_SEG - 1) / IMP_SEG;
  const long long total = (long long)n_img * H * nseg * IMP_CH;
  const int h = P / 2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % IMP_CH);
    long long r = i / IMP_CH;
    const int seg = (int)(r % nseg); r /= nseg;
    const long long row = r;   // img * H + y
    const float* src = t1 + row * W * IMP_CH + c;
    const int x0 = seg * IMP_SEG, x1 = min(x0 + IMP_SEG, W);
    double s = 0.0;
    for (int j = x0 - h; j < x0 - h + P; ++j) s += (double)src[(long long)reflect_idx(j, W) * IMP_CH];
    for (int x = x0; x < x1; ++x) {
      t2[(row * W + x) * IMP_CH + c] = (float)(s / (double)P);
      s += (double)src[(long long)reflect_idx(x - h + P, W) * IMP_CH] - (double)src[(long long)reflect_idx(x - h, W) * IMP_CH];
    }
  }
}

__device__ __forceinline__ void atomic_max_nonneg(float* addr, float v) { atomicMax((int*)addr, __float_as_int(v)); }

// per pixel: the two gamma-corrected variance maps (before their max-normalisation) + their per-image maxima
__global__ void imp_variance_kernel(const float* __restrict__ t2, int n_img, int H, int W, float* __restrict__ v2,
                                    float* __restrict__ vmax) {
  const long long npx = (long long)n_img * H * W;
  const float expo = (float)(1.0 / 2.2);
  for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < npx; p += (long long)gridDim.x * blockDim.x) {
    const int img = (int)(p / ((long long)H * W));
    const float* m = t2 + p * IMP_CH;
    float best[2] = {0.f, 0.f};
#pragma unroll
    for (int c = 0; c < 6; ++c) {
      const float mean = m[c], sq = m[6 + c];
      const float mean2 = __fmul_rn(mean, mean);
      float var = fmaxf(__fsub_rn(sq, mean2), 0.f);
      if (c < 3) var = __fdiv_rn(var, fmaxf(mean2, 1e-4f));      // relative variance for the radiance
      best[c / 3] = fmaxf(best[c / 3], var);
    }
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const float g = fminf((float)pow((double)best[k], (double)expo), 1.0f);   // float32 ** float32(1/2.2), correctly rounded
      v2[p * 2 + k] = g;
      atomic_max_nonneg(vmax + img * 2 + k, g);
    }
  }
}

__global__ void imp_combine_kernel(const float* __restrict__ v2, const float* __restrict__ vmax, int n_img, int H, int W,
                                   float* __restrict__ imp, float* __restrict__ imax) {
  const long long npx = (long long)n_img * H * W;
  for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < npx; p += (long long)gridDim.x * blockDim.x) {
    const int img = (int)(p / ((long long)H * W));
    const float a = __fdiv_rn(v2[p * 2], fmaxf(vmax[img * 2], 1e-4f));
    const float b = __fdiv_rn(v2[p * 2 + 1], fmaxf(vmax[img * 2 + 1], 1e-4f));
    const float r = __fadd_rn(a, b);   // temp * 1.0 + temp * 1.0
    imp[p] = r;
    atomic_max_nonneg(imax + img, r);
  }
}

__global__ void imp_normalise_kernel(float* __restrict__ imp, const float* __restrict__ imax, int n_img, int H, int W) {
  const long long npx = (long long)n_img * H * W;
  for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < npx; p += (long long)gridDim.x * blockDim.x)
    imp[p] = __fdiv_rn(imp[p], imax[p / ((long long)H * W)]);
}

static int imp_grid(long long items) {
  long long b = (items + 255) / 256;
  if (b > 148 * 32) b = 148 * 32;
  return (int)(b < 1 ? 1 : b);
}

}  // namespace pht

using namespace pht;

extern "C" size_t pht_importance_map_ws_bytes(int32_t n_img, int32_t Hf, int32_t Wf) {
  const size_t npx = (size_t)n_img * Hf * Wf;
  return (2 * npx * IMP_CH + 2 * npx) * sizeof(float) + (size_t)n_img * 3 * sizeof(float) + 256;
}

extern "C" int pht_importance_map(const float* noisy_f, const float* aux_f, int32_t n_img, int32_t Hf, int32_t Wf, int32_t P,
                                  float* imp, void* workspace, size_t workspace_bytes, void* stream) {
  PHT_CHECK_ARG(noisy_f && aux_f && imp && workspace && n_img > 0 && Hf > 0 && Wf > 0, "importance_map: bad args");
  PHT_CHECK_ARG(P >= 1 && P <= Hf && P <= Wf, "importance_map: the filter window must fit the frame (single reflection)");
  PHT_CHECK_ARG(workspace_bytes >= pht_importance_map_ws_bytes(n_img, Hf, Wf) && ((uintptr_t)workspace & 15) == 0,
                "importance_map: workspace too small or misaligned");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t npx = (size_t)n_img * Hf * Wf;
  float* t1 = (float*)workspace;
  float* t2 = t1 + npx * IMP_CH;
  float* v2 = t2 + npx * IMP_CH;
  float* vmax = v2 + 2 * npx;      // [n_img][2]
  float* imax = vmax + 2 * n_img;  // [n_img]
  PHT_CUDA(cudaMemsetAsync(vmax, 0, (size_t)n_img * 3 * sizeof(float), st));
  const int nseg_y = (Hf + IMP_SEG - 1) / IMP_SEG, nseg_x = (Wf + IMP_SEG - 1) / IMP_SEG;
  imp_box_y_kernel<<<imp_grid((long long)n_img * nseg_y * Wf * 6), 256, 0, st>>>(noisy_f, aux_f, n_img, Hf, Wf, P, t1);
  imp_box_x_kernel<<<imp_grid((long long)n_img * Hf * nseg_x * IMP_CH), 256, 0, st>>>(t1, n_img, Hf, Wf, P, t2);
  imp_variance_kernel<<<imp_grid((long long)npx), 256, 0, st>>>(t2, n_img, Hf, Wf, v2, vmax);
  imp_combine_kernel<<<imp_grid((long long)npx), 256, 0, st>>>(v2, vmax, n_img, Hf, Wf, imp, imax);
  imp_normalise_kernel<<<imp_grid((long long)npx), 256, 0, st>>>(imp, imax, n_img, Hf, Wf);
  count_launch(CNT_OTHER, 5);
  PHT_LAUNCH_CHECK();
  return PHT_OK;
}
