// CUDA-core implicit-GEMM convolution (forward / data-grad) and weight-grad.
//
// This is the fp32 PARITY path (PHT_F32: fp32 storage, fp32 FMA, used to show
// 1e-5 agreement with the reference) and the bit-level model of the tcgen05
// path for PHT_BF16 (same rounding points: bf16 operands, fp32 accumulation).
// Production bf16 shapes are routed to igemm_tc.cu by the dispatchers at the
// bottom of this file.
#include "common.cuh"

namespace pht {

int conv_gemm_tc(const pht_conv_gemm_args* a, cudaStream_t st, bool* handled);  // igemm_tc.cu
int wgrad_tc(const pht_wgrad_args* a, cudaStream_t st, bool* handled, pht_wgrad_reduce_job* defer);  // wgrad_tc.cu
int wgrad_reduce_batched(const pht_wgrad_reduce_job* jobs, int n, void* table_dev, size_t table_bytes, int upload, cudaStream_t st);
size_t wgrad_tc_workspace_bytes(const pht_wgrad_args* a);

struct GemmP {
  int B, Ho, Wo, N, ks, n_src, Ktot;
  unsigned flags;
  View src[3];
  int koff[4];
  const void* w;
  const float* bias;
  const float* slope;
  const float* mslope;
  View resid, mask, out1, out2;
};

template <typename T> __device__ __forceinline__ void ld4(const T* p, float* o);
template <> __device__ __forceinline__ void ld4<float>(const float* p, float* o) {
  float4 v = *reinterpret_cast<const float4*>(p);
  o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
}
template <> __device__ __forceinline__ void ld4<bf16>(const bf16* p, float* o) {
  uint2 v = *reinterpret_cast<const uint2*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
  float2 a = __bfloat1622float2(h[0]), b = __bfloat1622float2(h[1]);
  o[0] = a.x; o[1] = a.y; o[2] = b.x; o[3] = b.y;
}
template <typename T> __device__ __forceinline__ void st4(T* p, const float* o);
template <> __device__ __forceinline__ void st4<float>(float* p, const float* o) {
  *reinterpret_cast<float4*>(p) = make_float4(o[0], o[1], o[2], o[3]);
}
template <> __device__ __forceinline__ void st4<bf16>(bf16* p, const float* o) {
  uint2 v;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&v);
  h[0] = __floats2bfloat162_rn(o[0], o[1]);
  h[1] = __floats2bfloat162_rn(o[2], o[3]);
  *reinterpret_cast<uint2*>(p) = v;
}

constexpr int BM = 64, BN = 64, BK = 16;

template <typename T>
__global__ void __launch_bounds__(256) conv_gemm_simple_kernel(GemmP P) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const long long npx = (long long)P.B * P.Ho * P.Wo;
  const long long p0 = (long long)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  // loader role: pixel lm / weight row ln, 4 consecutive k at kq
  const int lm = tid >> 2, kq = (tid & 3) * 4;
  long long lp = p0 + lm;
  bool lvalid = lp < npx;
  int lb = 0, ly = 0, lx = 0;
  if (lvalid) {
    lx = (int)(lp % P.Wo);
    long long r = lp / P.Wo;
    ly = (int)(r % P.Ho);
    lb = (int)(r / P.Ho);
  }
  const int ln = n0 + lm;
  const bool nvalid = ln < P.N;
  const int ty = tid >> 4, tx = tid & 15;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int T_ = P.ks * P.ks, half = P.ks / 2;
  const T* wbase = (const T*)P.w;
  for (int t = 0; t < T_; ++t) {
    const int dy = t / P.ks - half, dx = t % P.ks - half;
    for (int k0 = 0; k0 < P.Ktot; k0 += BK) {
      const int kg = k0 + kq;
      // A tile
      float av[4] = {0.f, 0.f, 0.f, 0.f};
      if (lvalid) {
        int s = (kg >= P.koff[1]) + (kg >= P.koff[2]);
        const View& v = P.src[s];
        int yy = ly + dy + v.oy, xx = lx + dx + v.ox;
        if (view_inb(v, yy, xx)) ld4<T>((const T*)v.ptr + view_off(v, lb, yy, xx) + (kg - P.koff[s]), av);
      }
      float bv[4] = {0.f, 0.f, 0.f, 0.f};
      if (nvalid) ld4<T>(wbase + ((long long)t * P.N + ln) * P.Ktot + kg, bv);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        As[kq + j][lm] = av[j];
        Bs[kq + j][lm] = bv[j];
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < BK; ++k) {
        float a[4], b[4];
        *reinterpret_cast<float4*>(a) = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
        *reinterpret_cast<float4*>(b) = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] += a[i] * b[j];
      }
      __syncthreads();
    }
  }
  // epilogue
  const int n = n0 + tx * 4;
  if (n >= P.N) return;
  float bias[4] = {0.f, 0.f, 0.f, 0.f}, slope[4] = {1.f, 1.f, 1.f, 1.f}, msl[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    if (P.bias) bias[j] = P.bias[n + j];
    if (P.slope) slope[j] = P.slope[n + j];
    if (P.mslope) msl[j] = P.mslope[n + j];
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    long long p = p0 + ty * 4 + i;
    if (p >= npx) continue;
    int x = (int)(p % P.Wo);
    long long r = p / P.Wo;
    int y = (int)(r % P.Ho);
    int b = (int)(r / P.Ho);
    float v[4], rs[4] = {0.f, 0.f, 0.f, 0.f};
    if (P.flags & (PHT_EPI_RESID_PRE | PHT_EPI_RESID_POST))
      ld4<T>((const T*)P.resid.ptr + view_off(P.resid, b, y + P.resid.oy, x + P.resid.ox) + n, rs);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float t = acc[i][j] + bias[j];
      if (P.flags & PHT_EPI_RESID_PRE) t += rs[j];
      if (P.slope) t = t > 0.f ? t : t * slope[j];
      v[j] = t;
    }
    if (P.out1.ptr) st4<T>((T*)P.out1.ptr + view_off(P.out1, b, y + P.out1.oy, x + P.out1.ox) + n, v);
    if (P.out2.ptr) {
      if (P.flags & PHT_EPI_RESID_POST) {
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] += rs[j];
      }
      if (P.flags & PHT_EPI_MASK) {
        float mk[4];
        ld4<T>((const T*)P.mask.ptr + view_off(P.mask, b, y + P.mask.oy, x + P.mask.ox) + n, mk);
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] *= (mk[j] > 0.f ? 1.f : msl[j]);
      }
      st4<T>((T*)P.out2.ptr + view_off(P.out2, b, y + P.out2.oy, x + P.out2.ox) + n, v);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// weight gradient: dw[t][n][k] = sum_p dy[p][n] * src(p + tap_t)[k]
// grid: x = n-tile * k-tile, y = tap, z = pixel split; fp32 atomics into dw
// ---------------------------------------------------------------------------------------------
struct WgradP {
  int B, Ho, Wo, N, ks, n_src, Ktot, chunk;
  View dy;
  View src[3];
  int koff[4];
  float* dw;
};

template <typename T>
__global__ void __launch_bounds__(256) wgrad_simple_kernel(WgradP P) {
  __shared__ float As[BK][BM + 4];  // [pixel][n]
  __shared__ float Bs[BK][BN + 4];  // [pixel][k]
  const int tid = threadIdx.x;
  const int ktiles = (P.Ktot + BN - 1) / BN;
  const int n0 = (blockIdx.x / ktiles) * BM, k0 = (blockIdx.x % ktiles) * BN;
  const int t = blockIdx.y;
  const int half = P.ks / 2, dy_ = t / P.ks - half, dx_ = t % P.ks - half;
  const long long npx = (long long)P.B * P.Ho * P.Wo;
  const long long pbeg = (long long)blockIdx.z * P.chunk;
  const long long pend = pbeg + P.chunk < npx ? pbeg + P.chunk : npx;
  const int lpix = tid >> 4, c4 = (tid & 15) * 4;
  const int ty = tid >> 4, tx = tid & 15;
  // source of this thread's k group
  const int kg = k0 + c4;
  const bool kvalid = kg < P.Ktot;
  const int s = kvalid ? (kg >= P.koff[1]) + (kg >= P.koff[2]) : 0;
  const View sv = P.src[s];
  const int kc = kg - P.koff[s];
  const bool nvalid = n0 + c4 < P.N;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (long long pb = pbeg; pb < pend; pb += BK) {
    long long p = pb + lpix;
    float av[4] = {0.f, 0.f, 0.f, 0.f}, bv[4] = {0.f, 0.f, 0.f, 0.f};
    if (p < pend) {
      int x = (int)(p % P.Wo);
      long long r = p / P.Wo;
      int y = (int)(r % P.Ho);
      int b = (int)(r / P.Ho);
      if (nvalid) ld4<T>((const T*)P.dy.ptr + view_off(P.dy, b, y + P.dy.oy, x + P.dy.ox) + n0 + c4, av);
      int yy = y + dy_ + sv.oy, xx = x + dx_ + sv.ox;
      if (kvalid && view_inb(sv, yy, xx)) ld4<T>((const T*)sv.ptr + view_off(sv, b, yy, xx) + kc, bv);
    }
    *reinterpret_cast<float4*>(&As[lpix][c4]) = make_float4(av[0], av[1], av[2], av[3]);
    *reinterpret_cast<float4*>(&Bs[lpix][c4]) = make_float4(bv[0], bv[1], bv[2], bv[3]);
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[4], b[4];
      *reinterpret_cast<float4*>(a) = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      *reinterpret_cast<float4*>(b) = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] += a[i] * b[j];
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int n = n0 + ty * 4 + i;
    if (n >= P.N) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int k = k0 + tx * 4 + j;
      if (k < P.Ktot) atomicAdd(P.dw + ((long long)t * P.N + n) * P.Ktot + k, acc[i][j]);
    }
  }
}

// column sums of dy (bias gradient), HBM-bound: every pixel row of N channels is read once with 4-element
// vector loads (tpp = N/4 consecutive threads cover one pixel -> fully coalesced), `rows` pixels per block
// iteration, fp32 partials reduced through shared memory, one atomicAdd per channel per block.
template <typename T>
__global__ void colsum_kernel(View dy, int B, int Ho, int Wo, int N, int tpp, int rows, int chunk, float* __restrict__ out) {
  extern __shared__ float sm[];  // [rows][N]
  const int col = (threadIdx.x % tpp) * 4, prow = threadIdx.x / tpp;
  const long long npx = (long long)B * Ho * Wo;
  const long long pbeg = (long long)blockIdx.x * chunk, pend = pbeg + chunk < npx ? pbeg + chunk : npx;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  if (prow < rows)
    for (long long p = pbeg + prow; p < pend; p += rows) {
      int x = (int)(p % Wo);
      long long q = p / Wo;
      int y = (int)(q % Ho);
      int b = (int)(q / Ho);
      float v[4];
      ld4<T>((const T*)dy.ptr + view_off(dy, b, y + dy.oy, x + dy.ox) + col, v);
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[j] += v[j];
    }
  if (prow < rows) *reinterpret_cast<float4*>(&sm[prow * N + col]) = make_float4(acc[0], acc[1], acc[2], acc[3]);
  __syncthreads();
  for (int c = threadIdx.x; c < N; c += blockDim.x) {
    float t = 0.f;
    for (int i = 0; i < rows; ++i) t += sm[i * N + c];
    atomicAdd(out + c, t);
  }
}

// bf16 fast path of the column sum for any strided NHWC view: 16-byte loads, 4 loads in flight per thread, each
// thread owns one fixed 8-channel group (tpp = N/8 threads cover a pixel), blocks stride over image rows (no per-pixel
// division), block partials through shared memory, one atomicAdd per channel per block.
__global__ void __launch_bounds__(256) colsum_bf16_rows_kernel(const bf16* __restrict__ dy, long long sb, long long sy,
                                                               long long sx, int Ho, int Wo, int nrows, int N, int tpp,
                                                               int rows, float* __restrict__ out) {
  extern __shared__ float sm[];  // [rows][N]
  const int cg = threadIdx.x % tpp, prow = threadIdx.x / tpp;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  if (prow < rows) {
    for (int r = blockIdx.x; r < nrows; r += gridDim.x) {
      const int b = r / Ho, y = r - b * Ho;
      const bf16* base = dy + b * sb + y * sy + cg * 8;
      int x = prow;
      for (; x + 3 * rows < Wo; x += 4 * rows) {
        uint4 u[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) u[k] = __ldg(reinterpret_cast<const uint4*>(base + (long long)(x + k * rows) * sx));
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u[k]);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float2 f = __bfloat1622float2(h[j]);
            acc[2 * j] += f.x;
            acc[2 * j + 1] += f.y;
          }
        }
      }
      for (; x < Wo; x += rows) {
        uint4 u = __ldg(reinterpret_cast<const uint4*>(base + (long long)x * sx));
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float2 f = __bfloat1622float2(h[j]);
          acc[2 * j] += f.x;
          acc[2 * j + 1] += f.y;
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) sm[prow * N + cg * 8 + j] = acc[j];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < N; c += blockDim.x) {
    float t = 0.f;
    for (int i = 0; i < rows; ++i) t += sm[i * N + c];
    atomicAdd(out + c, t);
  }
}

static int check_view(const pht_view& v, int dtype, const char* what) {
  PHT_CHECK_ARG(v.ptr != nullptr, "%s: null view", what);
  PHT_CHECK_ARG(v.dtype == dtype, "%s: dtype mismatch", what);
  int ea = dtype == PHT_F32 ? 4 : 4;  // 4-element vector access
  PHT_CHECK_ARG(v.C % 4 == 0 && v.sx % ea == 0 && v.sy % ea == 0 && v.sb % ea == 0, "%s: view not 4-element aligned", what);
  size_t es = dtype == PHT_F32 ? 4 : 2;
  PHT_CHECK_ARG(((uintptr_t)v.ptr % (4 * es)) == 0, "%s: pointer misaligned", what);
  return PHT_OK;
}

int conv_gemm_simple(const pht_conv_gemm_args* a, cudaStream_t st) {
  GemmP P;
  P.B = a->B; P.Ho = a->Ho; P.Wo = a->Wo; P.N = a->N; P.ks = a->ksize; P.n_src = a->n_src; P.flags = a->flags;
  int k = 0;
  for (int s = 0; s < 3; ++s) {
    P.koff[s] = k;
    if (s < a->n_src) {
      int rc = check_view(a->src[s], a->dtype, "conv_gemm src");
      if (rc) return rc;
      P.src[s] = make_view(a->src[s]);
      k += a->src[s].C;
    } else {
      P.src[s] = null_view();
      P.koff[s] = 1 << 30;
    }
  }
  P.koff[3] = 1 << 30;
  P.Ktot = k;
  PHT_CHECK_ARG(P.Ktot % BK == 0, "conv_gemm: Ktot %% 16 != 0");
  PHT_CHECK_ARG(P.N % 4 == 0, "conv_gemm: N %% 4 != 0");
  P.w = a->w; P.bias = a->bias; P.slope = a->slope; P.mslope = a->mslope;
  P.resid = null_view(); P.mask = null_view(); P.out1 = null_view(); P.out2 = null_view();
  if (a->flags & (PHT_EPI_RESID_PRE | PHT_EPI_RESID_POST)) {
    int rc = check_view(a->resid, a->dtype, "conv_gemm resid");
    if (rc) return rc;
    P.resid = make_view(a->resid);
  }
  if (a->flags & PHT_EPI_MASK) {
    int rc = check_view(a->mask, a->dtype, "conv_gemm mask");
    if (rc) return rc;
    P.mask = make_view(a->mask);
  }
  if (a->out1.ptr) {
    int rc = check_view(a->out1, a->dtype, "conv_gemm out1");
    if (rc) return rc;
    P.out1 = make_view(a->out1);
  }
  if (a->out2.ptr) {
    int rc = check_view(a->out2, a->dtype, "conv_gemm out2");
    if (rc) return rc;
    P.out2 = make_view(a->out2);
  }
  long long npx = (long long)a->B * a->Ho * a->Wo;
  dim3 grid(ceil_div(npx, BM), ceil_div(a->N, BN));
  if (a->dtype == PHT_F32) conv_gemm_simple_kernel<float><<<grid, 256, 0, st>>>(P);
  else conv_gemm_simple_kernel<bf16><<<grid, 256, 0, st>>>(P);
  count_launch(CNT_GEMM_SIMPLE);
  PHT_LAUNCH_CHECK();
  return PHT_OK;
}

int wgrad_simple(const pht_wgrad_args* a, cudaStream_t st) {
  WgradP P;
  P.B = a->B; P.Ho = a->Ho; P.Wo = a->Wo; P.N = a->N; P.ks = a->ksize; P.n_src = a->n_src;
  int rc = check_view(a->dy, a->dtype, "wgrad dy");
  if (rc) return rc;
  P.dy = make_view(a->dy);
  int k = 0;
  for (int s = 0; s < 3; ++s) {
    P.koff[s] = k;
    if (s < a->n_src) {
      rc = check_view(a->src[s], a->dtype, "wgrad src");
      if (rc) return rc;
      P.src[s] = make_view(a->src[s]);
      k += a->src[s].C;
    } else {
      P.src[s] = null_view();
      P.koff[s] = 1 << 30;
    }
  }
  P.koff[3] = 1 << 30;
  P.Ktot = k;
  P.dw = a->dw;
  PHT_CHECK_ARG(P.N % 4 == 0 && P.Ktot % 4 == 0, "wgrad: N, Ktot must be multiples of 4");
  const int T_ = a->ksize * a->ksize;
  long long npx = (long long)a->B * a->Ho * a->Wo;
  int tiles = ceil_div(a->N, BM) * ceil_div(P.Ktot, BN) * T_;
  const int sms = sm_count();
  int splits = (4 * sms + tiles - 1) / tiles;
  long long maxsplits = (npx + 255) / 256;
  if (splits > maxsplits) splits = (int)maxsplits;
  if (splits < 1) splits = 1;
  P.chunk = (int)((npx + splits - 1) / splits);
  P.chunk = ((P.chunk + BK - 1) / BK) * BK;
  splits = (int)((npx + P.chunk - 1) / P.chunk);
  PHT_CUDA(cudaMemsetAsync(a->dw, 0, (size_t)T_ * a->N * P.Ktot * sizeof(float), st));
  dim3 grid(ceil_div(a->N, BM) * ceil_div(P.Ktot, BN), T_, splits);
  if (a->dtype == PHT_F32) wgrad_simple_kernel<float><<<grid, 256, 0, st>>>(P);
  else wgrad_simple_kernel<bf16><<<grid, 256, 0, st>>>(P);
  count_launch(CNT_WGRAD_SIMPLE);
  PHT_LAUNCH_CHECK();
  return PHT_OK;
}

int colsum(const pht_view& dy, int dtype, int B, int Ho, int Wo, int N, float* out, cudaStream_t st) {
  long long npx = (long long)B * Ho * Wo;
  PHT_CHECK_ARG(N % 4 == 0 && N <= 4096, "colsum: N must be a multiple of 4 and <= 4096");
  const int tpp = N / 4;                           // threads per pixel row
  int rows = tpp >= 256 ? 1 : 256 / tpp;           // pixel rows per block iteration
  int threads = tpp * rows;
  if (threads > 1024) { rows = 1; threads = tpp; }
  const int sms = sm_count();
  int splits = (int)((npx + 16 * rows - 1) / (16 * rows));
  if (splits > 4 * sms) splits = 4 * sms;
  if (splits < 1) splits = 1;
  int chunk = (int)((npx + splits - 1) / splits);
  splits = (int)((npx + chunk - 1) / chunk);
  PHT_CUDA(cudaMemsetAsync(out, 0, (size_t)N * sizeof(float), st));
  if (dtype == PHT_BF16 && N % 8 == 0 && N / 8 <= 256 && dy.sx % 8 == 0 && dy.sy % 8 == 0 && dy.sb % 8 == 0 &&
      ((uintptr_t)dy.ptr & 15) == 0) {
    const int tpp8 = N / 8, rows8 = 256 / tpp8;
    const int nrows = B * Ho;
    int blocks = nrows < 8 * sms ? nrows : 8 * sms;
    const bf16* base = (const bf16*)dy.ptr + (long long)dy.oy * dy.sy + (long long)dy.ox * dy.sx;
    colsum_bf16_rows_kernel<<<blocks, 256, (size_t)rows8 * N * sizeof(float), st>>>(base, dy.sb, dy.sy, dy.sx, Ho, Wo, nrows, N,
                                                                                     tpp8, rows8, out);
    count_launch(CNT_OTHER);
    PHT_LAUNCH_CHECK();
    return PHT_OK;
  }
  View v = make_view(dy);
  size_t smem = (size_t)rows * N * sizeof(float);
  if (dtype == PHT_F32) colsum_kernel<float><<<splits, threads, smem, st>>>(v, B, Ho, Wo, N, tpp, rows, chunk, out);
  else colsum_kernel<bf16><<<splits, threads, smem, st>>>(v, B, Ho, Wo, N, tpp, rows, chunk, out);
  count_launch(CNT_OTHER);
  PHT_LAUNCH_CHECK();
  return PHT_OK;
}

}  // namespace pht

using namespace pht;

extern "C" {

int pht_conv_gemm(const pht_conv_gemm_args* a, void* stream) {
  PHT_CHECK_ARG(a != nullptr, "conv_gemm: null args");
  PHT_CHECK_ARG(a->dtype == PHT_F32 || a->dtype == PHT_BF16, "conv_gemm: bad dtype");
  PHT_CHECK_ARG(a->ksize == 1 || a->ksize == 3 || a->ksize == 5, "conv_gemm: ksize must be 1, 3 or 5");
  PHT_CHECK_ARG(a->n_src >= 1 && a->n_src <= 3, "conv_gemm: n_src must be 1..3");
  PHT_CHECK_ARG(a->B > 0 && a->Ho > 0 && a->Wo > 0 && a->N > 0 && a->w, "conv_gemm: bad dims");
  PHT_CHECK_ARG(a->out1.ptr || a->out2.ptr, "conv_gemm: no output");
  cudaStream_t st = (cudaStream_t)stream;
  if (a->dtype == PHT_BF16 && !force_simple()) {
    bool handled = false;
    int rc = conv_gemm_tc(a, st, &handled);
    if (rc) return rc;
    if (handled) return PHT_OK;
  }
  if (a->flags & PHT_EPI_PADFOLD) {
    set_error("conv_gemm: PHT_EPI_PADFOLD needs the bf16 tensor-core path (ksize 3, H and W multiples of 8, TMA-able views)");
    return PHT_ERR_UNSUPPORTED;
  }
  if (a->flags & (PHT_EPI_RING1 | PHT_EPI_RING2)) {
    set_error("conv_gemm: PHT_EPI_RING1 / RING2 need the bf16 tensor-core path (bf16 outputs, H, W >= 4, no PHT_EPI_PADFOLD)");
    return PHT_ERR_UNSUPPORTED;
  }
  if (a->dtype == PHT_BF16 && !force_simple() && !bf16_fallback_allowed()) {
    set_error("conv_gemm: bf16 launch (N=%d ksize=%d n_src=%d C0=%d) is not eligible for the tcgen05 kernel (channels must be "
              "multiples of 64, views 16-byte aligned) and the CUDA-core fallback is disabled (option bf16_fallback)",
              a->N, a->ksize, a->n_src, a->src[0].C);
    return PHT_ERR_UNSUPPORTED;
  }
  return conv_gemm_simple(a, st);
}

size_t pht_wgrad_workspace_bytes(const pht_wgrad_args* a) {
  size_t tc = (a && a->dtype == PHT_BF16) ? wgrad_tc_workspace_bytes(a) : 0;
  return tc > 256 ? tc : 256;
}

int pht_wgrad(const pht_wgrad_args* a, void* stream) {
  PHT_CHECK_ARG(a != nullptr && a->dw, "wgrad: null args");
  PHT_CHECK_ARG(a->dtype == PHT_F32 || a->dtype == PHT_BF16, "wgrad: bad dtype");
  PHT_CHECK_ARG(a->ksize == 1 || a->ksize == 3 || a->ksize == 5, "wgrad: ksize must be 1, 3 or 5");
  PHT_CHECK_ARG(a->n_src >= 1 && a->n_src <= 3, "wgrad: n_src must be 1..3");
  cudaStream_t st = (cudaStream_t)stream;
  if (a->dtype == PHT_BF16 && !force_simple()) {
    bool handled = false;
    int rc = wgrad_tc(a, st, &handled, nullptr);   // also produces dbias (fused column sums)
    if (rc) return rc;
    if (handled) return PHT_OK;
    if (!bf16_fallback_allowed()) {
      set_error("wgrad: bf16 launch (N=%d ksize=%d n_src=%d C0=%d, workspace %zu B) is not eligible for the tcgen05 kernel and "
                "the CUDA-core fallback is disabled (option bf16_fallback)", a->N, a->ksize, a->n_src, a->src[0].C,
                a->workspace_bytes);
      return PHT_ERR_UNSUPPORTED;
    }
  }
  if (a->dbias) {
    int rc = colsum(a->dy, a->dtype, a->B, a->Ho, a->Wo, a->N, a->dbias, st);
    if (rc) return rc;
  }
  return wgrad_simple(a, st);
}

int pht_wgrad_partial(const pht_wgrad_args* a, pht_wgrad_reduce_job* job, void* stream) {
  PHT_CHECK_ARG(a != nullptr && a->dw && job, "wgrad_partial: null args");
  PHT_CHECK_ARG(a->ksize == 1 || a->ksize == 3 || a->ksize == 5, "wgrad_partial: ksize must be 1, 3 or 5");
  PHT_CHECK_ARG(a->n_src >= 1 && a->n_src <= 3, "wgrad_partial: n_src must be 1..3");
  if (a->dtype == PHT_BF16 && !force_simple()) {
    bool handled = false;
    int rc = wgrad_tc(a, (cudaStream_t)stream, &handled, job);
    if (rc) return rc;
    if (handled) return PHT_OK;
  }
  set_error("wgrad_partial: shape not taken by the split tensor-core kernel");
  return PHT_ERR_UNSUPPORTED;
}

int pht_wgrad_reduce_batched(const pht_wgrad_reduce_job* jobs, int32_t n, void* table_dev, size_t table_bytes, int32_t upload,
                             void* stream) {
  return wgrad_reduce_batched(jobs, n, table_dev, table_bytes, upload, (cudaStream_t)stream);
}

}  // extern "C"
