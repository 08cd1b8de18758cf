// BatchNorm2d (training mode) + LeakyReLU of the WGAN-GP critic, with first- AND second-order backward, for sm_100a.
//
//   reference   DiscriminatorVGG's conv_block(norm_type="batch", act_type="leakyrelu") layers
//               (pht/models/afgsa/model.py:264-344, norm :52-61 -> nn.BatchNorm2d(affine=True), act :64-83 ->
//               nn.LeakyReLU(0.2)), differentiated twice by GradientPenaltyLoss (pht/models/losses.py:12-57: the gradient
//               of the critic w.r.t. its input is itself part of the loss).  Under stock PyTorch the second-order pass
//               through native_batch_norm_backward unrolls into hundreds of small element-wise / reduction launches:
//               ~45 % of the critic step's GPU time at prod.
//
// Per channel c over m = B*H*W values (fp32, channels-last / NHWC):
//   forward        mu = mean(x), var = mean((x - mu)^2), r = 1 / sqrt(var + eps), xh = (x - mu) r,
//                  y = gamma xh + beta, z = y > 0 ? y : slope y;  running stats updated like nn.BatchNorm2d
//   backward       gy = gz s(y)  (s = 1 or slope);  g_beta = sum gy,  g_gamma = sum gy xh,
//                  gx = gamma r (gy - mean(gy) - xh mean(gy xh))
//   double bwd     cotangent h of gx (the gradient-penalty pass), with a = mean(gy), b = mean(gy xh),
//                  Q = sum h gy - sum h sum gy / m - sum h xh sum gy xh / m:
//                    h_gz    = s(y) gamma r (h - mean(h) - xh mean(h xh))          (the map gy -> gx is self-adjoint)
//                    h_gamma = r Q
//                    h_x     = -gamma r^2 [ Q xh / m + w - mean(w) - xh mean(w xh) ],   w = b h + mean(h xh) gy
//                  (mu and r are functions of x: their dependence is folded into h_x; s(y) is piecewise constant.)
// Every pass is a per-channel column-sum kernel (fp64 accumulation, block partials + last-block finalisation in a fixed
// order: deterministic) followed by one vectorised element-wise kernel.
#include <math.h>

#include "common.cuh"

namespace pht {

constexpr int BN_MAXC = 1024;
constexpr int BN_THREADS = 256;
constexpr int BN_MAX_BLOCKS = 512;

struct BnP {
  const float* x;      // [m][C]
  const float* g;      // gz (modes 1, 2)
  const float* h;      // cotangent of gx (mode 2)
  const float* gamma;
  const float* beta;
  const float* stat;   // [2][C] mean, rstd (modes 1, 2)
  long long m;
  int C;
  float slope, eps, momentum;
  float* out_stat;     // mode 0: [2][C] mean, rstd
  float* run_mean;     // mode 0 (may be null)
  float* run_var;
  float* sums;         // modes 1, 2: [NS][C] finished sums (fp32)
  const float* pre_bias;  // mode 0 (may be null): per-channel constant the producer did NOT add to x (see pht_bn_act_fwd)
  float* g_beta;       // mode 1 (may be null): copies of sums[0] / sums[1] = d beta / d gamma
  float* g_gamma;
  double* partials;    // [blocks][NS][C]
  unsigned* ticket;
};

template <int MODE> struct BnNS { static constexpr int value = MODE == 2 ? 5 : (MODE == 3 ? 1 : 2); };   // MODE 3: plain column sum of x

// column sums: thread (tx, ty) = (float4 column group, row lane); a block strides over the rows
template <int MODE>
__global__ void __launch_bounds__(BN_THREADS) bn_colsum_kernel(BnP P) {
  constexpr int NS = BnNS<MODE>::value;
  const int C4 = P.C >> 2, RP = BN_THREADS / C4;
  const int tx = threadIdx.x % C4, ty = threadIdx.x / C4;
  double acc[NS][4];
#pragma unroll
  for (int s = 0; s < NS; ++s)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[s][j] = 0.0;
  float mu[4], r[4], ga[4], be[4];
  if (MODE == 1 || MODE == 2) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = tx * 4 + j;
      mu[j] = P.stat[c]; r[j] = P.stat[P.C + c]; ga[j] = P.gamma[c]; be[j] = P.beta[c];
    }
  }
  // U rows per thread in flight per trip (the kernel is latency-bound otherwise: a block only covers RP rows per pass)
  constexpr int U = 4;
  const long long stride = (long long)gridDim.x * RP;
  for (long long row0 = (long long)blockIdx.x * RP + ty; row0 < P.m; row0 += stride * U) {
    float4 xv[U], gv[U], hv[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long row = row0 + u * stride;
      if (row < P.m) {
        xv[u] = *reinterpret_cast<const float4*>(P.x + row * P.C + tx * 4);
        if (MODE != 0 && MODE != 3) gv[u] = *reinterpret_cast<const float4*>(P.g + row * P.C + tx * 4);
        if (MODE == 2) hv[u] = *reinterpret_cast<const float4*>(P.h + row * P.C + tx * 4);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (row0 + u * stride >= P.m) break;
      const float xs[4] = {xv[u].x, xv[u].y, xv[u].z, xv[u].w};
      if (MODE == 0) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          acc[0][j] += (double)xs[j];
          acc[1][j] += (double)xs[j] * (double)xs[j];
        }
      } else if (MODE == 3) {
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[0][j] += (double)xs[j];
      } else {
        const float gs[4] = {gv[u].x, gv[u].y, gv[u].z, gv[u].w};
        float hs[4] = {0.f, 0.f, 0.f, 0.f};
        if (MODE == 2) { hs[0] = hv[u].x; hs[1] = hv[u].y; hs[2] = hv[u].z; hs[3] = hv[u].w; }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float xh = (xs[j] - mu[j]) * r[j];
          const float y = fmaf(ga[j], xh, be[j]);
          const float gy = gs[j] * (y > 0.f ? 1.f : P.slope);
          if (MODE == 1) {
            acc[0][j] += (double)gy;
            acc[1][j] += (double)gy * (double)xh;
          } else {
            acc[0][j] += (double)hs[j];
            acc[1][j] += (double)hs[j] * (double)xh;
            acc[2][j] += (double)gy;
            acc[3][j] += (double)gy * (double)xh;
            acc[4][j] += (double)hs[j] * (double)gy;
          }
        }
      }
    }
  }
  // reduce over the row lanes through shared memory (fixed order)
  __shared__ double red[BN_THREADS * 4];
  __shared__ bool last;
  double* mine = P.partials + (size_t)blockIdx.x * NS * P.C;
  for (int s = 0; s < NS; ++s) {
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 4; ++j) red[threadIdx.x * 4 + j] = acc[s][j];
    __syncthreads();
    if (ty == 0) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        double t = 0.0;
        for (int k = 0; k < RP; ++k) t += red[(k * C4 + tx) * 4 + j];
        mine[s * P.C + tx * 4 + j] = t;
      }
    }
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) last = atomicAdd(P.ticket, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!last) return;
  __threadfence();
  // the last block adds the block partials in block order and finalises
  for (int c = threadIdx.x; c < P.C; c += BN_THREADS) {
    double t[NS];
#pragma unroll
    for (int s = 0; s < NS; ++s) t[s] = 0.0;
    // (L2 loads, several in flight: a volatile read per add would serialise ~300 L2 round trips)
#pragma unroll 8
    for (unsigned b = 0; b < gridDim.x; ++b)
#pragma unroll
      for (int s = 0; s < NS; ++s) t[s] += __ldcg(P.partials + ((size_t)b * NS + s) * P.C + c);
    if (MODE == 0) {
      const double mean = t[0] / (double)P.m;
      double var = t[1] / (double)P.m - mean * mean;
      if (var < 0.0) var = 0.0;
      P.out_stat[c] = (float)mean;
      P.out_stat[P.C + c] = (float)(1.0 / sqrt(var + (double)P.eps));
      if (P.run_mean) {   // nn.BatchNorm2d: running = (1 - momentum) running + momentum batch (unbiased variance)
        const double unb = P.m > 1 ? var * (double)P.m / (double)(P.m - 1) : var;
        const double shift = P.pre_bias ? (double)P.pre_bias[c] : 0.0;
        P.run_mean[c] = (float)((1.0 - P.momentum) * (double)P.run_mean[c] + P.momentum * (mean + shift));
        P.run_var[c] = (float)((1.0 - P.momentum) * (double)P.run_var[c] + P.momentum * unb);
      }
    } else if (MODE == 3) {
      P.out_stat[c] = (float)t[0];
    } else {
#pragma unroll
      for (int s = 0; s < NS; ++s) P.sums[s * P.C + c] = (float)t[s];
      if (MODE == 1) {
        if (P.g_beta) P.g_beta[c] = (float)t[0];
        if (P.g_gamma) P.g_gamma[c] = (float)t[1];
      }
    }
  }
  if (threadIdx.x == 0) *P.ticket = 0u;
}

__global__ void bn_act_fwd_kernel(const float* __restrict__ x, const float* __restrict__ stat, const float* __restrict__ gamma,
                                  const float* __restrict__ beta, float* __restrict__ z, long long n4, int C, float slope) {
  const int C4 = C >> 2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C4) * 4;
    const float4 xv = reinterpret_cast<const float4*>(x)[i];
    const float xs[4] = {xv.x, xv.y, xv.z, xv.w};
    float o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float y = fmaf(gamma[c + j], (xs[j] - stat[c + j]) * stat[C + c + j], beta[c + j]);
      o[j] = y > 0.f ? y : y * slope;
    }
    reinterpret_cast<float4*>(z)[i] = make_float4(o[0], o[1], o[2], o[3]);
  }
}

// gx = gamma r (gy - S_g / m - xh S_gx / m)
__global__ void bn_act_bwd_kernel(const float* __restrict__ x, const float* __restrict__ gz, const float* __restrict__ stat,
                                  const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ sums,
                                  float* __restrict__ gx, long long n4, int C, float slope, float inv_m) {
  const int C4 = C >> 2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C4) * 4;
    const float4 xv = reinterpret_cast<const float4*>(x)[i], gv = reinterpret_cast<const float4*>(gz)[i];
    const float xs[4] = {xv.x, xv.y, xv.z, xv.w}, gs[4] = {gv.x, gv.y, gv.z, gv.w};
    float o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float r = stat[C + c + j], xh = (xs[j] - stat[c + j]) * r;
      const float y = fmaf(gamma[c + j], xh, beta[c + j]);
      const float gy = gs[j] * (y > 0.f ? 1.f : slope);
      o[j] = gamma[c + j] * r * (gy - sums[c + j] * inv_m - xh * sums[C + c + j] * inv_m);
    }
    reinterpret_cast<float4*>(gx)[i] = make_float4(o[0], o[1], o[2], o[3]);
  }
}

// h_gz and h_x from the five sums {S_h, S_hx, S_g, S_gx, S_hg}
__global__ void bn_act_bwd_bwd_kernel(const float* __restrict__ x, const float* __restrict__ gz, const float* __restrict__ h,
                                      const float* __restrict__ stat, const float* __restrict__ gamma, const float* __restrict__ beta,
                                      const float* __restrict__ sums, float* __restrict__ h_gz, float* __restrict__ h_x, long long n4,
                                      int C, float slope, float inv_m) {
  const int C4 = C >> 2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C4) * 4;
    const float4 xv = reinterpret_cast<const float4*>(x)[i], gv = reinterpret_cast<const float4*>(gz)[i],
                 hv = reinterpret_cast<const float4*>(h)[i];
    const float xs[4] = {xv.x, xv.y, xv.z, xv.w}, gs[4] = {gv.x, gv.y, gv.z, gv.w}, hs[4] = {hv.x, hv.y, hv.z, hv.w};
    float og[4], ox[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int cc = c + j;
      const float r = stat[C + cc], xh = (xs[j] - stat[cc]) * r, ga = gamma[cc];
      const float y = fmaf(ga, xh, beta[cc]);
      const float s = y > 0.f ? 1.f : slope;
      const float gy = gs[j] * s;
      const float Sh = sums[cc], Shx = sums[C + cc], Sg = sums[2 * C + cc], Sgx = sums[3 * C + cc], Shg = sums[4 * C + cc];
      og[j] = s * ga * r * (hs[j] - Sh * inv_m - xh * Shx * inv_m);
      const float b = Sgx * inv_m, mhx = Shx * inv_m;
      const float Q = Shg - Sh * Sg * inv_m - Shx * Sgx * inv_m;
      const float w = b * hs[j] + mhx * gy;
      const float mean_w = b * Sh * inv_m + mhx * Sg * inv_m;
      const float mean_wx = 2.f * b * mhx;
      ox[j] = -ga * r * r * (Q * xh * inv_m + w - mean_w - xh * mean_wx);
    }
    reinterpret_cast<float4*>(h_gz)[i] = make_float4(og[0], og[1], og[2], og[3]);
    reinterpret_cast<float4*>(h_x)[i] = make_float4(ox[0], ox[1], ox[2], ox[3]);
  }
}

// h_gamma = r Q
__global__ void bn_hgamma_kernel(const float* __restrict__ stat, const float* __restrict__ sums, float* __restrict__ h_gamma, int C,
                                 float inv_m) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) {
    const float Sh = sums[c], Shx = sums[C + c], Sg = sums[2 * C + c], Sgx = sums[3 * C + c], Shg = sums[4 * C + c];
    h_gamma[c] = stat[C + c] * (Shg - Sh * Sg * inv_m - Shx * Sgx * inv_m);
  }
}

static int bn_check(long long m, int C, const void* ws, size_t ws_bytes) {
  PHT_CHECK_ARG(m > 0 && C >= 4 && C % 4 == 0 && C <= BN_MAXC && BN_THREADS % (C / 4) == 0,
                "bn_act: C must be 4 x a power of two, <= 1024");
  PHT_CHECK_ARG(ws && ((uintptr_t)ws & 15) == 0 && ws_bytes >= (size_t)BN_MAX_BLOCKS * 5 * C * sizeof(double) + 5 * C * sizeof(float) + 64,
                "bn_act: workspace too small (pht_bn_act_ws_bytes)");
  return PHT_OK;
}
static int bn_blocks(long long m, int C) {
  const int rp = BN_THREADS / (C / 4);
  long long b = (m + rp * 8 - 1) / (rp * 8);
  const long long cap = sm_count();     // one block per SM: fewer partials for the last block to add up
  if (b > cap) b = cap;
  if (b > BN_MAX_BLOCKS) b = BN_MAX_BLOCKS;
  return b < 1 ? 1 : (int)b;
}
static int ew_blocks(long long n4) {
  long long b = (n4 + 255) / 256;
  const long long cap = 8ll * sm_count();
  return (int)(b > cap ? cap : (b < 1 ? 1 : b));
}

}  // namespace pht

using namespace pht;

extern "C" {

size_t pht_bn_act_ws_bytes(int32_t C) { return (size_t)BN_MAX_BLOCKS * 5 * C * sizeof(double) + 5 * C * sizeof(float) + 64; }

// workspace layout: [BN_MAX_BLOCKS][5][C] double partials | [5][C] float sums | ticket
#define BN_WS(ws, C)                                                                  \
  double* partials = (double*)(ws);                                                   \
  float* sums = (float*)(partials + (size_t)BN_MAX_BLOCKS * 5 * (C));                 \
  unsigned* ticket = (unsigned*)(sums + 5 * (C))

int pht_bn_act_fwd(const float* x, const float* gamma, const float* beta, const float* pre_bias, float* run_mean, float* run_var,
                   float* stat, float* z, int64_t m, int32_t C, float eps, float momentum, float slope, void* workspace,
                   size_t workspace_bytes, void* stream) {
  PHT_CHECK_ARG(x && gamma && beta && stat && z, "bn_act_fwd: null arg");
  int rc = bn_check(m, C, workspace, workspace_bytes);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  BN_WS(workspace, C);
  BnP P = {};
  P.x = x; P.gamma = gamma; P.beta = beta; P.m = m; P.C = C; P.slope = slope; P.eps = eps; P.momentum = momentum;
  P.out_stat = stat; P.run_mean = run_mean; P.run_var = run_var; P.sums = sums; P.partials = partials; P.ticket = ticket;
  P.pre_bias = pre_bias;
  bn_colsum_kernel<0><<<bn_blocks(m, C), BN_THREADS, 0, st>>>(P);
  const long long n4 = m * C / 4;
  bn_act_fwd_kernel<<<ew_blocks(n4), 256, 0, st>>>(x, stat, gamma, beta, z, n4, C, slope);
  count_launch(CNT_OTHER, 2);
  PHT_LAUNCH_CHECK();
  return PHT_OK;
}

// out[c] = sum over the m rows of x[.][c] (the bias gradient of the critic's convolutions), same deterministic reduction
int pht_colsum_f32(const float* x, float* out, int64_t m, int32_t C, void* workspace, size_t workspace_bytes, void* stream) {
  PHT_CHECK_ARG(x && out, "colsum_f32: null arg");
  int rc = bn_check(m, C, workspace, workspace_bytes);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  BN_WS(workspace, C);
  (void)sums;
  BnP P = {};
  P.x = x; P.m = m; P.C = C; P.out_stat = out; P.partials = partials; P.ticket = ticket;
  bn_colsum_kernel<3><<<bn_blocks(m, C), BN_THREADS, 0, st>>>(P);
  count_launch(CNT_OTHER);
  PHT_LAUNCH_CHECK();
  return PHT_OK;
}

int pht_bn_act_bwd(const float* x, const float* gz, const float* gamma, const float* beta, const float* stat, float* gx,
                   float* g_gamma, float* g_beta, int64_t m, int32_t C, float slope, void* workspace, size_t workspace_bytes,
                   void* stream) {
  PHT_CHECK_ARG(x && gz && gamma && beta && stat && gx, "bn_act_bwd: null arg");
  int rc = bn_check(m, C, workspace, workspace_bytes);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  BN_WS(workspace, C);
  BnP P = {};
  P.x = x; P.g = gz; P.gamma = gamma; P.beta = beta; P.stat = stat; P.m = m; P.C = C; P.slope = slope;
  P.sums = sums; P.partials = partials; P.ticket = ticket;
  P.g_beta = g_beta; P.g_gamma = g_gamma;
  bn_colsum_kernel<1><<<bn_blocks(m, C), BN_THREADS, 0, st>>>(P);
  const long long n4 = m * C / 4;
  bn_act_bwd_kernel<<<ew_blocks(n4), 256, 0, st>>>(x, gz, stat, gamma, beta, sums, gx, n4, C, slope, 1.0f / (float)m);
  count_launch(CNT_OTHER, 2);
  PHT_LAUNCH_CHECK();
  return PHT_OK;
}

int pht_bn_act_bwd_bwd(const float* x, const float* gz, const float* h, const float* gamma, const float* beta, const float* stat,
                       float* h_gz, float* h_x, float* h_gamma, int64_t m, int32_t C, float slope, void* workspace,
                       size_t workspace_bytes, void* stream) {
  PHT_CHECK_ARG(x && gz && h && gamma && beta && stat && h_gz && h_x, "bn_act_bwd_bwd: null arg");
  int rc = bn_check(m, C, workspace, workspace_bytes);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  BN_WS(workspace, C);
  BnP P = {};
  P.x = x; P.g = gz; P.h = h; P.gamma = gamma; P.beta = beta; P.stat = stat; P.m = m; P.C = C; P.slope = slope;
  P.sums = sums; P.partials = partials; P.ticket = ticket;
  bn_colsum_kernel<2><<<bn_blocks(m, C), BN_THREADS, 0, st>>>(P);
  const long long n4 = m * C / 4;
  const float inv_m = 1.0f / (float)m;
  bn_act_bwd_bwd_kernel<<<ew_blocks(n4), 256, 0, st>>>(x, gz, h, stat, gamma, beta, sums, h_gz, h_x, n4, C, slope, inv_m);
  if (h_gamma) bn_hgamma_kernel<<<(C + 127) / 128, 128, 0, st>>>(stat, sums, h_gamma, C, inv_m);
  count_launch(CNT_OTHER, 2 + (h_gamma ? 1 : 0));
  PHT_LAUNCH_CHECK();
  return PHT_OK;
}

}  // extern "C"
