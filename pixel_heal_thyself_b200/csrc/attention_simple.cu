// CUDA-core block-local attention (forward + recompute backward), fp32 math.
// Parity path for PHT_F32 and rounding model for the tensor-core path.
// One CTA per (query block, head).  Semantics: pht_b200.h / model.py:474-516.
#include "common.cuh"

namespace pht {

int attn_fwd_tc(const pht_attn_args* a, cudaStream_t st, bool* handled);       // attention_tc.cu
int attn_bwd_tc(const pht_attn_bwd_args* a, cudaStream_t st, bool* handled);   // attention_tc.cu
int attn_bwd_zero_tc(const pht_attn_bwd_args* a, cudaStream_t st, bool* handled);   // attention_tc.cu
size_t attn_bwd_tc_ws_bytes(const pht_attn_args& f);                            // attention_tc.cu

struct AttnP {
  int B, H, W, heads, d, block, halo, win, nq, nk, nby, nbx;
  View q, k, v, resid, out, d_out, dq;
  const float* rel_h;
  const float* rel_w;
  float* lse;
  float* dk_acc;
  float* dv_acc;
  float* rel_part;  // [nblk*heads][2*win*d/2]
};

// smem layout helpers (strides padded by 1 float to dodge bank conflicts)
struct Smem {
  float *Q, *K, *V, *S, *dO, *aux;
};
__device__ __forceinline__ Smem carve(float* base, const AttnP& P, bool bwd) {
  Smem s;
  const int ds = P.d + 1, ss = P.nk + 1;
  s.Q = base;
  s.K = s.Q + P.nq * ds;
  s.V = s.K + P.nk * ds;
  s.S = s.V + P.nk * ds;
  s.dO = s.S + P.nq * ss;
  s.aux = s.dO + (bwd ? P.nq * ds : 0);
  return s;
}
static size_t smem_bytes(const AttnP& P, bool bwd) {
  size_t ds = P.d + 1, ss = P.nk + 1;
  size_t n = P.nq * ds + 2 * P.nk * ds + P.nq * ss + (bwd ? P.nq * ds : 0) + 2 * P.nq;
  return n * sizeof(float);
}

template <typename T>
__device__ void load_tiles(const AttnP& P, const Smem& s, int b, int by, int bx, int h) {
  const int ds = P.d + 1;
  for (int e = threadIdx.x; e < P.nq * P.d; e += blockDim.x) {
    int q = e / P.d, j = e % P.d;
    int y = by * P.block + q / P.block, x = bx * P.block + q % P.block;
    s.Q[q * ds + j] = view_ld<T>(P.q, b, y, x, h * P.d + j);
  }
  for (int e = threadIdx.x; e < P.nk * P.d; e += blockDim.x) {
    int kk = e / P.d, j = e % P.d;
    int r = kk / P.win, c = kk % P.win;
    int y = by * P.block - P.halo + r, x = bx * P.block - P.halo + c;
    bool in = (unsigned)y < (unsigned)P.H && (unsigned)x < (unsigned)P.W;
    float kv = in ? view_ld<T>(P.k, b, y, x, h * P.d + j) : 0.f;
    float vv = in ? view_ld<T>(P.v, b, y, x, h * P.d + j) : 0.f;
    const int hd = P.d / 2;
    kv += j < hd ? P.rel_h[r * hd + j] : P.rel_w[c * hd + (j - hd)];
    s.K[kk * ds + j] = kv;
    s.V[kk * ds + j] = vv;
  }
}

__device__ void compute_scores(const AttnP& P, const Smem& s) {
  const int ds = P.d + 1, ss = P.nk + 1;
  for (int e = threadIdx.x; e < P.nq * P.nk; e += blockDim.x) {
    int q = e / P.nk, kk = e % P.nk;
    float acc = 0.f;
    for (int j = 0; j < P.d; ++j) acc += s.Q[q * ds + j] * s.K[kk * ds + j];
    s.S[q * ss + kk] = acc;
  }
}

template <typename T>
__global__ void __launch_bounds__(256) attn_fwd_simple_kernel(AttnP P) {
  extern __shared__ float smem[];
  Smem s = carve(smem, P, false);
  const int blk = blockIdx.x, h = blockIdx.y;
  const int bx = blk % P.nbx, by = (blk / P.nbx) % P.nby, b = blk / (P.nbx * P.nby);
  const int ds = P.d + 1, ss = P.nk + 1;
  load_tiles<T>(P, s, b, by, bx, h);
  __syncthreads();
  compute_scores(P, s);
  __syncthreads();
  float* inv_sum = s.aux;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int q = warp; q < P.nq; q += nw) {
    float m = -INFINITY;
    for (int kk = lane; kk < P.nk; kk += 32) m = fmaxf(m, s.S[q * ss + kk]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float sum = 0.f;
    for (int kk = lane; kk < P.nk; kk += 32) {
      float e = expf(s.S[q * ss + kk] - m);
      s.S[q * ss + kk] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    if (lane == 0) {
      inv_sum[q] = 1.f / sum;
      int y = by * P.block + q / P.block, x = bx * P.block + q % P.block;
      if (P.lse) P.lse[(((long long)b * P.H + y) * P.W + x) * P.heads + h] = m + logf(sum);
    }
  }
  __syncthreads();
  for (int e = threadIdx.x; e < P.nq * P.d; e += blockDim.x) {
    int q = e / P.d, j = e % P.d;
    float acc = 0.f;
    for (int kk = 0; kk < P.nk; ++kk) acc += s.S[q * ss + kk] * s.V[kk * ds + j];
    acc *= inv_sum[q];
    int y = by * P.block + q / P.block, x = bx * P.block + q % P.block;
    if (P.resid.ptr) acc += view_ld<T>(P.resid, b, y, x, h * P.d + j);
    ((T*)P.out.ptr)[view_off(P.out, b, y, x) + h * P.d + j] = from_f<T>(acc);
  }
}

template <typename T>
__global__ void __launch_bounds__(256) attn_bwd_simple_kernel(AttnP P) {
  extern __shared__ float smem[];
  Smem s = carve(smem, P, true);
  const int blk = blockIdx.x, h = blockIdx.y;
  const int bx = blk % P.nbx, by = (blk / P.nbx) % P.nby, b = blk / (P.nbx * P.nby);
  const int ds = P.d + 1, ss = P.nk + 1;
  load_tiles<T>(P, s, b, by, bx, h);
  for (int e = threadIdx.x; e < P.nq * P.d; e += blockDim.x) {
    int q = e / P.d, j = e % P.d;
    int y = by * P.block + q / P.block, x = bx * P.block + q % P.block;
    s.dO[q * ds + j] = view_ld<T>(P.d_out, b, y, x, h * P.d + j);
  }
  float* lse_s = s.aux;
  for (int q = threadIdx.x; q < P.nq; q += blockDim.x) {
    int y = by * P.block + q / P.block, x = bx * P.block + q % P.block;
    lse_s[q] = P.lse[(((long long)b * P.H + y) * P.W + x) * P.heads + h];
  }
  __syncthreads();
  compute_scores(P, s);
  __syncthreads();
  // P = exp(S - lse)
  for (int e = threadIdx.x; e < P.nq * P.nk; e += blockDim.x) {
    int q = e / P.nk, kk = e % P.nk;
    s.S[q * ss + kk] = expf(s.S[q * ss + kk] - lse_s[q]);
  }
  __syncthreads();
  // dV[kk][j] = sum_q P[q][kk] dO[q][j]
  for (int e = threadIdx.x; e < P.nk * P.d; e += blockDim.x) {
    int kk = e / P.d, j = e % P.d;
    int r = kk / P.win, c = kk % P.win;
    int y = by * P.block - P.halo + r, x = bx * P.block - P.halo + c;
    if ((unsigned)y >= (unsigned)P.H || (unsigned)x >= (unsigned)P.W) continue;
    float acc = 0.f;
    for (int q = 0; q < P.nq; ++q) acc += s.S[q * ss + kk] * s.dO[q * ds + j];
    atomicAdd(P.dv_acc + (((long long)b * P.H + y) * P.W + x) * (P.heads * P.d) + h * P.d + j, acc);
  }
  __syncthreads();
  // dS = P * (dP - delta), row by row (warp per row; dP kept in registers)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int q = warp; q < P.nq; q += nw) {
    float dp[8];
    float part = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      int kk = lane + 32 * i;
      dp[i] = 0.f;
      if (kk < P.nk) {
        float acc = 0.f;
        for (int j = 0; j < P.d; ++j) acc += s.dO[q * ds + j] * s.V[kk * ds + j];
        dp[i] = acc;
        part += s.S[q * ss + kk] * acc;
      }
    }
    float delta = warp_sum(part);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      int kk = lane + 32 * i;
      if (kk < P.nk) s.S[q * ss + kk] *= (dp[i] - delta);
    }
  }
  __syncthreads();
  // dQ[q][j] = sum_kk dS[q][kk] K'[kk][j]
  for (int e = threadIdx.x; e < P.nq * P.d; e += blockDim.x) {
    int q = e / P.d, j = e % P.d;
    float acc = 0.f;
    for (int kk = 0; kk < P.nk; ++kk) acc += s.S[q * ss + kk] * s.K[kk * ds + j];
    int y = by * P.block + q / P.block, x = bx * P.block + q % P.block;
    ((T*)P.dq.ptr)[view_off(P.dq, b, y, x) + h * P.d + j] = from_f<T>(acc);
  }
  // dK'[kk][j] = sum_q dS[q][kk] Q[q][j]  -> global atomics (in-image keys) + smem copy for rel grads
  float* dK = s.V;  // V is dead after the dS pass; all warps passed the barrier above
  for (int e = threadIdx.x; e < P.nk * P.d; e += blockDim.x) {
    int kk = e / P.d, j = e % P.d;
    float acc = 0.f;
    for (int q = 0; q < P.nq; ++q) acc += s.S[q * ss + kk] * s.Q[q * ds + j];
    dK[kk * ds + j] = acc;
    int r = kk / P.win, c = kk % P.win;
    int y = by * P.block - P.halo + r, x = bx * P.block - P.halo + c;
    if ((unsigned)y < (unsigned)P.H && (unsigned)x < (unsigned)P.W)
      atomicAdd(P.dk_acc + (((long long)b * P.H + y) * P.W + x) * (P.heads * P.d) + h * P.d + j, acc);
  }
  __syncthreads();
  const int hd = P.d / 2;
  float* part = P.rel_part + ((long long)blk * P.heads + h) * (2 * P.win * hd);
  for (int e = threadIdx.x; e < 2 * P.win * hd; e += blockDim.x) {
    int which = e / (P.win * hd), rc = (e / hd) % P.win, j = e % hd;
    float acc = 0.f;
    if (which == 0) for (int c = 0; c < P.win; ++c) acc += dK[(rc * P.win + c) * ds + j];
    else for (int r = 0; r < P.win; ++r) acc += dK[(r * P.win + rc) * ds + hd + j];
    part[e] = acc;
  }
}

__global__ void rel_reduce_kernel(const float* __restrict__ part, int nparts, int n, int half, float* __restrict__ d_rel_h,
                                  float* __restrict__ d_rel_w) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float s = 0.f;
  for (int p = 0; p < nparts; ++p) s += part[(long long)p * n + i];
  if (i < half) d_rel_h[i] = s;
  else d_rel_w[i - half] = s;
}

static int fill_params(const pht_attn_args* a, AttnP* P) {
  PHT_CHECK_ARG(a->heads > 0 && a->head_dim > 0 && a->head_dim % 2 == 0 && a->head_dim <= 64, "attn: head_dim must be even and <= 64");
  PHT_CHECK_ARG(a->block > 0 && a->halo >= 0, "attn: bad block/halo");
  PHT_CHECK_ARG(a->H % a->block == 0 && a->W % a->block == 0,
                "attn: feature map dimensions must be divisible by the block size");  // model.py:469-471
  PHT_CHECK_ARG(a->q.ptr && a->k.ptr && a->v.ptr && a->rel_h && a->rel_w, "attn: null input");
  P->B = a->B; P->H = a->H; P->W = a->W; P->heads = a->heads; P->d = a->head_dim; P->block = a->block; P->halo = a->halo;
  P->win = a->block + 2 * a->halo;
  P->nq = a->block * a->block;
  P->nk = P->win * P->win;
  PHT_CHECK_ARG(P->nk <= 256, "attn: window too large for the simple kernel");
  P->nby = a->H / a->block; P->nbx = a->W / a->block;
  P->q = make_view(a->q); P->k = make_view(a->k); P->v = make_view(a->v);
  P->resid = a->resid.ptr ? make_view(a->resid) : null_view();
  P->out = a->out.ptr ? make_view(a->out) : null_view();
  P->rel_h = a->rel_h; P->rel_w = a->rel_w; P->lse = a->lse;
  P->d_out = null_view(); P->dq = null_view();
  P->dk_acc = P->dv_acc = P->rel_part = nullptr;
  return PHT_OK;
}

int attn_fwd_simple(const pht_attn_args* a, cudaStream_t st) {
  AttnP P;
  int rc = fill_params(a, &P);
  if (rc) return rc;
  PHT_CHECK_ARG(P.out.ptr, "attn_fwd: null out");
  size_t sm = smem_bytes(P, false);
  dim3 grid(P.B * P.nby * P.nbx, P.heads);
  if (a->dtype == PHT_F32) {
    PHT_CUDA(cudaFuncSetAttribute(attn_fwd_simple_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    attn_fwd_simple_kernel<float><<<grid, 256, sm, st>>>(P);
  } else {
    PHT_CUDA(cudaFuncSetAttribute(attn_fwd_simple_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    attn_fwd_simple_kernel<bf16><<<grid, 256, sm, st>>>(P);
  }
  count_launch(CNT_ATTN_SIMPLE);
  PHT_LAUNCH_CHECK();
  return PHT_OK;
}

// fp32 accumulator [npx][C] -> strided NHWC view in the activation dtype
template <typename T>
__global__ void acc_to_view_kernel(const float* __restrict__ acc, View out, int B, int H, int W, int C) {
  long long total = (long long)B * H * W * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % C);
    long long p = i / C;
    int x = (int)(p % W);
    long long r = p / W;
    int y = (int)(r % H);
    int b = (int)(r / H);
    ((T*)out.ptr)[view_off(out, b, y, x) + c] = from_f<T>(acc[i]);
  }
}

size_t attn_bwd_simple_ws_bytes(const pht_attn_args& f) {
  size_t npx = (size_t)f.B * f.H * f.W, C = (size_t)f.heads * f.head_dim;
  size_t nblk = (size_t)f.B * (f.H / f.block) * (f.W / f.block);
  size_t win = f.block + 2 * f.halo;
  return 2 * npx * C * sizeof(float) + nblk * f.heads * (2 * win * (f.head_dim / 2)) * sizeof(float);
}

int attn_bwd_simple(const pht_attn_bwd_args* a, cudaStream_t st) {
  AttnP P;
  int rc = fill_params(&a->fwd, &P);
  if (rc) return rc;
  PHT_CHECK_ARG(a->d_out.ptr && a->dq.ptr && a->dk.ptr && a->dv.ptr && a->d_rel_h && a->d_rel_w && a->fwd.lse, "attn_bwd: null arg");
  PHT_CHECK_ARG(a->workspace && a->workspace_bytes >= attn_bwd_simple_ws_bytes(a->fwd), "attn_bwd: workspace too small");
  P.d_out = make_view(a->d_out); P.dq = make_view(a->dq);
  const size_t npx = (size_t)P.B * P.H * P.W, C = (size_t)P.heads * P.d;
  P.dk_acc = (float*)a->workspace;
  P.dv_acc = P.dk_acc + npx * C;
  P.rel_part = P.dv_acc + npx * C;
  PHT_CUDA(cudaMemsetAsync(P.dk_acc, 0, 2 * npx * C * sizeof(float), st));
  size_t sm = smem_bytes(P, true);
  dim3 grid(P.B * P.nby * P.nbx, P.heads);
  int cgrid = (int)((npx * C + 255) / 256);
  if (cgrid > 148 * 16) cgrid = 148 * 16;
  if (a->fwd.dtype == PHT_F32) {
    PHT_CUDA(cudaFuncSetAttribute(attn_bwd_simple_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    attn_bwd_simple_kernel<float><<<grid, 256, sm, st>>>(P);
    acc_to_view_kernel<float><<<cgrid, 256, 0, st>>>(P.dk_acc, make_view(a->dk), P.B, P.H, P.W, (int)C);
    acc_to_view_kernel<float><<<cgrid, 256, 0, st>>>(P.dv_acc, make_view(a->dv), P.B, P.H, P.W, (int)C);
  } else {
    PHT_CUDA(cudaFuncSetAttribute(attn_bwd_simple_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    attn_bwd_simple_kernel<bf16><<<grid, 256, sm, st>>>(P);
    acc_to_view_kernel<bf16><<<cgrid, 256, 0, st>>>(P.dk_acc, make_view(a->dk), P.B, P.H, P.W, (int)C);
    acc_to_view_kernel<bf16><<<cgrid, 256, 0, st>>>(P.dv_acc, make_view(a->dv), P.B, P.H, P.W, (int)C);
  }
  PHT_LAUNCH_CHECK();
  int n = 2 * P.win * (P.d / 2);
  rel_reduce_kernel<<<ceil_div(n, 128), 128, 0, st>>>(P.rel_part, grid.x * grid.y, n, n / 2, a->d_rel_h, a->d_rel_w);
  count_launch(CNT_ATTN_SIMPLE);
  count_launch(CNT_OTHER, 3);
  PHT_LAUNCH_CHECK();
  return PHT_OK;
}

}  // namespace pht

using namespace pht;

extern "C" {

int pht_attn_fwd(const pht_attn_args* a, void* stream) {
  PHT_CHECK_ARG(a != nullptr, "attn_fwd: null args");
  PHT_CHECK_ARG(a->dtype == PHT_F32 || a->dtype == PHT_BF16, "attn_fwd: bad dtype");
  cudaStream_t st = (cudaStream_t)stream;
  if (a->dtype == PHT_BF16 && !force_simple()) {
    bool handled = false;
    int rc = attn_fwd_tc(a, st, &handled);
    if (rc) return rc;
    if (handled) return PHT_OK;
    if (!bf16_fallback_allowed() && a->H % a->block == 0 && a->W % a->block == 0) {   // (else: the reference's assertion below)
      set_error("attn_fwd: bf16 launch (heads=%d head_dim=%d block=%d halo=%d) is not eligible for the tcgen05 kernel and the "
                "CUDA-core fallback is disabled (option bf16_fallback)", a->heads, a->head_dim, a->block, a->halo);
      return PHT_ERR_UNSUPPORTED;
    }
  }
  if (a->ring) {
    set_error("attn_fwd: ring (frame of a padded output) needs the bf16 tensor-core path");
    return PHT_ERR_UNSUPPORTED;
  }
  return attn_fwd_simple(a, st);
}

size_t pht_attn_bwd_workspace_bytes(const pht_attn_bwd_args* a) {
  if (!a || a->fwd.block <= 0 || a->fwd.heads <= 0) return 0;
  size_t s = attn_bwd_simple_ws_bytes(a->fwd);
  size_t t = a->fwd.dtype == PHT_BF16 ? attn_bwd_tc_ws_bytes(a->fwd) : 0;
  return s > t ? s : t;
}

int pht_attn_bwd_zero(const pht_attn_bwd_args* a, void* stream) {
  PHT_CHECK_ARG(a != nullptr, "attn_bwd_zero: null args");
  if (a->fwd.dtype == PHT_BF16 && !force_simple()) {
    bool handled = false;
    int rc = attn_bwd_zero_tc(a, (cudaStream_t)stream, &handled);
    if (rc) return rc;
  }
  return PHT_OK;   // (the CUDA-core path overwrites its outputs: nothing to prepare)
}

int pht_attn_bwd(const pht_attn_bwd_args* a, void* stream) {
  PHT_CHECK_ARG(a != nullptr, "attn_bwd: null args");
  PHT_CHECK_ARG(a->fwd.dtype == PHT_F32 || a->fwd.dtype == PHT_BF16, "attn_bwd: bad dtype");
  cudaStream_t st = (cudaStream_t)stream;
  if (a->fwd.dtype == PHT_BF16 && !force_simple()) {
    bool handled = false;
    int rc = attn_bwd_tc(a, st, &handled);
    if (rc) return rc;
    if (handled) return PHT_OK;
    if (!bf16_fallback_allowed()) {
      set_error("attn_bwd: bf16 launch (heads=%d head_dim=%d block=%d halo=%d, workspace %zu B) is not eligible for the tcgen05 "
                "kernel and the CUDA-core fallback is disabled (option bf16_fallback)", a->fwd.heads, a->fwd.head_dim,
                a->fwd.block, a->fwd.halo, a->workspace_bytes);
      return PHT_ERR_UNSUPPORTED;
    }
  }
  return attn_bwd_simple(a, st);
}

}  // extern "C"
