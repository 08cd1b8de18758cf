// Optional MS-SSIM + L1 image loss of the reference trainer, fused forward + backward.
//
//   reference   SSIMLoss.forward, pht/models/losses.py:248-263: scale = max(channel-max of the TARGET, 1) per pixel,
//               kornia.losses.MS_SSIMLoss(reduction="mean")(input / scale, target / scale); used with weight 0.1 at
//               pht/models/base_trainer.py:450-452.
//   arithmetic  kornia 0.8.0's MS_SSIMLoss (not vendored with the reference, not installed here: PARITY UNPINNED, see
//               oracle/msssim_oracle.py, which restates the same published algorithm and is this kernel's checker):
//               five 33 x 33 normalised gaussians (sigma 0.5, 1, 2, 4, 8; zero padding 16) applied by ONE grouped
//               convolution whose 15 outputs pair output c with input channel c / 5 and sigma c / 3, i.e. the seven
//               distinct (channel, sigma) pairs below with multiplicities; luminance term from the last pair, cubed;
//               product of all 15 contrast-structure terms; gaussian(sigma 8)-weighted L1; alpha 0.025, x 200.
//
// The gaussians are separable (outer products), so every window is two 33-tap passes.  Everything is fp32 and
// HBM / L2-bound stencil work on 3-channel images (8 x 128 x 128 pixels at prod): clarity over speed.
//   prep     x = out / scale, y = gt / scale; source planes x, y, x^2, y^2, xy per channel and |x - y|
//   blur     38 (source plane, sigma) jobs, horizontal then vertical
//   combine  per pixel: lc, cs of the 7 pairs -> loss pixel (+ block partial sums, fixed-order final sum) and the
//            24 gradient planes d/d mu_x, d/d E[x^2], d/d E[xy] per pair, d/d blurred-L1 per channel
//   blur     the 24 gradient planes (a zero-padded symmetric correlation is its own transpose)
//   grad     d loss / d out = (sum over the channel's pairs: B(dmu) + 2 x B(dExx) + y B(dExy)) + sign(x - y) B(dl1), / scale
#include <math.h>

#include "common.cuh"

namespace pht {

constexpr int MS_TAPS = 33, MS_PAD = 16, MS_NSIG = 5, MS_NPAIR = 7;
constexpr int MS_SRC = 18;                      // 5 per channel (x, y, xx, yy, xy) + 3 |x - y|
constexpr int MS_FWD = 5 * MS_NPAIR + 3;        // 38 blurred planes
constexpr int MS_BWD = 3 * MS_NPAIR + 3;        // 24 gradient planes
constexpr int MS_PLANES = 1 + MS_SRC + 2 * MS_FWD + MS_BWD;   // + 1 / scale
__constant__ float c_gauss[MS_NSIG][MS_TAPS];
__constant__ int c_pair_ch[MS_NPAIR] = {0, 0, 1, 1, 1, 2, 2};
__constant__ int c_pair_sig[MS_NPAIR] = {0, 1, 1, 2, 3, 3, 4};
__constant__ int c_pair_mult[MS_NPAIR] = {3, 2, 1, 3, 1, 2, 3};
constexpr float MS_C1 = 0.01f * 0.01f, MS_C2 = 0.03f * 0.03f, MS_ALPHA = 0.025f, MS_COMP = 200.0f;

struct MsJob {
  int src, dst, sig;   // plane indices (relative to the workspace base), sigma index
};
__constant__ MsJob c_fwd_h[MS_FWD], c_fwd_v[MS_FWD], c_bwd_h[MS_BWD], c_bwd_v[MS_BWD];

// workspace plane map
constexpr int PL_INV = 0, PL_SRC = 1, PL_TMP = PL_SRC + MS_SRC, PL_OUT = PL_TMP + MS_FWD, PL_GIN = PL_OUT + MS_FWD;

__global__ void msssim_prep_kernel(const float* __restrict__ out, const float* __restrict__ gt, float* __restrict__ ws, int B,
                                   int HW) {
  const long long N = (long long)B * HW;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / HW), p = (int)(i % HW);
    const float* o = out + (long long)b * 3 * HW + p;
    const float* g = gt + (long long)b * 3 * HW + p;
    const float g0 = g[0], g1 = g[HW], g2 = g[2 * (long long)HW];
    const float scale = fmaxf(fmaxf(fmaxf(g0, g1), g2), 1.0f);   // losses.py:259-262
    ws[PL_INV * N + i] = scale;
    const float gv[3] = {g0, g1, g2};
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float x = o[c * (long long)HW] / scale, y = gv[c] / scale;
      float* s = ws + (PL_SRC + 5 * c) * N + i;
      s[0] = x; s[N] = y; s[2 * N] = x * x; s[3 * N] = y * y; s[4 * N] = x * y;
      ws[(PL_SRC + 15 + c) * N + i] = fabsf(x - y);
    }
  }
}

// one separable pass of every job: blockIdx.y = job; zero padding
template <bool VERT>
__global__ void msssim_blur_kernel(float* __restrict__ ws, const MsJob* __restrict__ jobs, int B, int H, int W) {
  const MsJob j = jobs[blockIdx.y];
  const long long N = (long long)B * H * W;
  const float* src = ws + j.src * N;
  float* dst = ws + j.dst * N;
  const float* w = c_gauss[j.sig];
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % W);
    const int y = (int)((i / W) % H);
    float a = 0.f;
    if (VERT) {
#pragma unroll
      for (int k = 0; k < MS_TAPS; ++k) {
        const int yy = y + k - MS_PAD;
        if ((unsigned)yy < (unsigned)H) a = fmaf(w[k], src[i + (long long)(k - MS_PAD) * W], a);
      }
    } else {
#pragma unroll
      for (int k = 0; k < MS_TAPS; ++k) {
        const int xx = x + k - MS_PAD;
        if ((unsigned)xx < (unsigned)W) a = fmaf(w[k], src[i + (k - MS_PAD)], a);
      }
    }
    dst[i] = a;
  }
}

__global__ void __launch_bounds__(256) msssim_combine_kernel(float* __restrict__ ws, long long N, float gscale,
                                                              float* __restrict__ partials) {
  float acc = 0.f;
  const float g_t = -MS_COMP * MS_ALPHA * gscale / (float)N;          // d mean-loss / d T(p)
  const float g_l1 = MS_COMP * (1.0f - MS_ALPHA) * gscale / (float)N / 3.0f;   // d mean-loss / d blurred |x - y|_c (p)
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x) {
    float cs[MS_NPAIR], b2[MS_NPAIR], mux[MS_NPAIR], muy[MS_NPAIR];
    float lc6 = 1.f, b1_6 = 1.f;
#pragma unroll
    for (int p = 0; p < MS_NPAIR; ++p) {
      const float* o = ws + (PL_OUT + 5 * p) * N + i;
      const float mx = o[0], my = o[N], exx = o[2 * N], eyy = o[3 * N], exy = o[4 * N];
      const float sxy = exy - mx * my, sx2 = exx - mx * mx, sy2 = eyy - my * my;
      b2[p] = sx2 + sy2 + MS_C2;
      cs[p] = (2.f * sxy + MS_C2) / b2[p];
      mux[p] = mx; muy[p] = my;
      if (p == MS_NPAIR - 1) {
        b1_6 = mx * mx + my * my + MS_C1;
        lc6 = (2.f * mx * my + MS_C1) / b1_6;
      }
    }
    const float lm = lc6 * lc6 * lc6;
    float t = lm;
#pragma unroll
    for (int p = 0; p < MS_NPAIR; ++p)
      for (int m = 0; m < c_pair_mult[p]; ++m) t *= cs[p];
    const float* l = ws + (PL_OUT + 5 * MS_NPAIR) * N + i;
    const float gl1 = (l[0] + l[N] + l[2 * N]) * (1.0f / 3.0f);
    acc += MS_COMP * (MS_ALPHA * (1.f - t) + (1.f - MS_ALPHA) * gl1);
    // ---- gradient planes ----
#pragma unroll
    for (int p = 0; p < MS_NPAIR; ++p) {
      float others = lm;                       // T with one factor cs[p] removed (no division: cs may be ~0)
#pragma unroll
      for (int q = 0; q < MS_NPAIR; ++q)
        for (int m = 0; m < c_pair_mult[q] - (q == p ? 1 : 0); ++m) others *= cs[q];
      const float g_cs = g_t * (float)c_pair_mult[p] * others;
      float d_mu = g_cs * (2.f * mux[p] * cs[p] - 2.f * muy[p]) / b2[p];
      if (p == MS_NPAIR - 1) {
        float pics = 1.f;
#pragma unroll
        for (int q = 0; q < MS_NPAIR; ++q)
          for (int m = 0; m < c_pair_mult[q]; ++m) pics *= cs[q];
        const float g_lc = g_t * 3.f * lc6 * lc6 * pics;
        d_mu += g_lc * (2.f * muy[p] - 2.f * mux[p] * lc6) / b1_6;
      }
      float* gi = ws + (PL_GIN + 3 * p) * N + i;
      gi[0] = d_mu;
      gi[N] = -g_cs * cs[p] / b2[p];           // d / d E[x^2]
      gi[2 * N] = g_cs * 2.f / b2[p];          // d / d E[xy]
    }
    float* gl = ws + (PL_GIN + 3 * MS_NPAIR) * N + i;
    gl[0] = g_l1; gl[N] = g_l1; gl[2 * N] = g_l1;
  }
  __shared__ float red[8];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int w = 0; w < 8; ++w) s += red[w];
    partials[blockIdx.x] = s;
  }
}

__global__ void msssim_finish_kernel(const float* __restrict__ partials, int n, long long N, float* __restrict__ loss) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < n; ++i) s += (double)partials[i];   // fixed order: deterministic
    loss[0] = (float)(s / (double)N);
  }
}

__global__ void msssim_grad_kernel(const float* __restrict__ ws, float* __restrict__ grad, int B, int HW) {
  const long long N = (long long)B * HW;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / HW), p = (int)(i % HW);
    const float inv = 1.0f / ws[PL_INV * N + i];
    float g[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int q = 0; q < MS_NPAIR; ++q) {
      const int c = c_pair_ch[q];
      const float* bo = ws + (PL_OUT + 3 * q) * N + i;      // blurred gradient planes of pair q
      const float x = ws[(PL_SRC + 5 * c) * N + i], y = ws[(PL_SRC + 5 * c + 1) * N + i];
      g[c] += bo[0] + 2.f * x * bo[N] + y * bo[2 * N];
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float x = ws[(PL_SRC + 5 * c) * N + i], y = ws[(PL_SRC + 5 * c + 1) * N + i];
      const float d = x - y;
      g[c] += (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f)) * ws[(PL_OUT + 3 * MS_NPAIR + c) * N + i];
      grad[((long long)b * 3 + c) * HW + p] = g[c] * inv;
    }
  }
}

static int msssim_init_tables() {
  static PerDeviceOnce once;
  const int d = cur_device() & (PHT_MAX_DEVICES - 1);
  if (once.done[d]) return PHT_OK;
  const float sig[MS_NSIG] = {0.5f, 1.0f, 2.0f, 4.0f, 8.0f};
  float g[MS_NSIG][MS_TAPS];
  for (int s = 0; s < MS_NSIG; ++s) {          // float arithmetic like torch: exp(-(c^2) / (2 sigma^2)), / sum
    float sum = 0.f;
    for (int k = 0; k < MS_TAPS; ++k) {
      const float c = (float)(k - MS_TAPS / 2);
      g[s][k] = expf(-(c * c) / (2.f * sig[s] * sig[s]));
      sum += g[s][k];
    }
    for (int k = 0; k < MS_TAPS; ++k) g[s][k] /= sum;
  }
  PHT_CUDA(cudaMemcpyToSymbol(c_gauss, g, sizeof(g)));
  const int pch[MS_NPAIR] = {0, 0, 1, 1, 1, 2, 2}, psig[MS_NPAIR] = {0, 1, 1, 2, 3, 3, 4};
  MsJob fh[MS_FWD], fv[MS_FWD], bh[MS_BWD], bv[MS_BWD];
  for (int p = 0; p < MS_NPAIR; ++p)
    for (int m = 0; m < 5; ++m) {
      fh[5 * p + m] = {PL_SRC + 5 * pch[p] + m, PL_TMP + 5 * p + m, psig[p]};
      fv[5 * p + m] = {PL_TMP + 5 * p + m, PL_OUT + 5 * p + m, psig[p]};
    }
  for (int c = 0; c < 3; ++c) {
    fh[5 * MS_NPAIR + c] = {PL_SRC + 15 + c, PL_TMP + 5 * MS_NPAIR + c, MS_NSIG - 1};
    fv[5 * MS_NPAIR + c] = {PL_TMP + 5 * MS_NPAIR + c, PL_OUT + 5 * MS_NPAIR + c, MS_NSIG - 1};
  }
  for (int p = 0; p < MS_NPAIR; ++p)
    for (int m = 0; m < 3; ++m) {
      bh[3 * p + m] = {PL_GIN + 3 * p + m, PL_TMP + 3 * p + m, psig[p]};
      bv[3 * p + m] = {PL_TMP + 3 * p + m, PL_OUT + 3 * p + m, psig[p]};
    }
  for (int c = 0; c < 3; ++c) {
    bh[3 * MS_NPAIR + c] = {PL_GIN + 3 * MS_NPAIR + c, PL_TMP + 3 * MS_NPAIR + c, MS_NSIG - 1};
    bv[3 * MS_NPAIR + c] = {PL_TMP + 3 * MS_NPAIR + c, PL_OUT + 3 * MS_NPAIR + c, MS_NSIG - 1};
  }
  PHT_CUDA(cudaMemcpyToSymbol(c_fwd_h, fh, sizeof(fh)));
  PHT_CUDA(cudaMemcpyToSymbol(c_fwd_v, fv, sizeof(fv)));
  PHT_CUDA(cudaMemcpyToSymbol(c_bwd_h, bh, sizeof(bh)));
  PHT_CUDA(cudaMemcpyToSymbol(c_bwd_v, bv, sizeof(bv)));
  once.done[d] = 1;
  return PHT_OK;
}

constexpr int MS_COMBINE_BLOCKS = 592;

}  // namespace pht

using namespace pht;

extern "C" {

size_t pht_msssim_ws_bytes(int32_t B, int32_t H, int32_t W) {
  return ((size_t)MS_PLANES * B * H * W + MS_COMBINE_BLOCKS) * sizeof(float) + 256;
}

int pht_msssim_loss(const float* out_nchw, const float* gt_nchw, int32_t B, int32_t H, int32_t W, float grad_scale, float* loss,
                    float* grad, void* workspace, size_t workspace_bytes, void* stream) {
  PHT_CHECK_ARG(out_nchw && gt_nchw && loss && B > 0 && H > 0 && W > 0, "msssim_loss: bad args");
  PHT_CHECK_ARG(workspace && ((uintptr_t)workspace & 15) == 0 && workspace_bytes >= pht_msssim_ws_bytes(B, H, W),
                "msssim_loss: workspace too small (pht_msssim_ws_bytes)");
  int rc = msssim_init_tables();
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  float* ws = (float*)workspace;
  const long long N = (long long)B * H * W;
  float* partials = ws + (size_t)MS_PLANES * N;
  int gx = (int)((N + 255) / 256);
  if (gx > sm_count() * 8) gx = sm_count() * 8;
  const MsJob *fh, *fv, *bh, *bv;
  PHT_CUDA(cudaGetSymbolAddress((void**)&fh, c_fwd_h));
  PHT_CUDA(cudaGetSymbolAddress((void**)&fv, c_fwd_v));
  PHT_CUDA(cudaGetSymbolAddress((void**)&bh, c_bwd_h));
  PHT_CUDA(cudaGetSymbolAddress((void**)&bv, c_bwd_v));
  msssim_prep_kernel<<<gx, 256, 0, st>>>(out_nchw, gt_nchw, ws, B, H * W);
  msssim_blur_kernel<false><<<dim3(gx, MS_FWD), 256, 0, st>>>(ws, fh, B, H, W);
  msssim_blur_kernel<true><<<dim3(gx, MS_FWD), 256, 0, st>>>(ws, fv, B, H, W);
  const int cb = gx < MS_COMBINE_BLOCKS ? gx : MS_COMBINE_BLOCKS;
  msssim_combine_kernel<<<cb, 256, 0, st>>>(ws, N, grad_scale, partials);
  msssim_finish_kernel<<<1, 32, 0, st>>>(partials, cb, N, loss);
  int launches = 5;
  if (grad) {
    msssim_blur_kernel<false><<<dim3(gx, MS_BWD), 256, 0, st>>>(ws, bh, B, H, W);
    msssim_blur_kernel<true><<<dim3(gx, MS_BWD), 256, 0, st>>>(ws, bv, B, H, W);
    msssim_grad_kernel<<<gx, 256, 0, st>>>(ws, grad, B, H * W);
    launches += 3;
  }
  count_launch(CNT_OTHER, launches);
  PHT_LAUNCH_CHECK();
  return PHT_OK;
}

}  // extern "C"
