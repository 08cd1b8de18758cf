// Bandwidth-bound kernels of the AFGSA hot path: padding border fill / fold,
// encoder im2col, decoder tail, L1 loss, preprocessing, Adam, weight packing.
// All are vectorised (16-byte accesses along the channel dim), coalesced and
// use warp-shuffle reductions; grids are sized in multiples of the SM count
// where a grid-stride loop is used.
#include <vector>

#include "common.cuh"

namespace pht {

static inline int num_sms() { return sm_count(); }

template <typename T> struct Vec;
template <> struct Vec<float> {
  static constexpr int N = 4;
  __device__ static void ld(const float* p, float* o) {
    float4 v = *reinterpret_cast<const float4*>(p);
    o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
  }
  __device__ static void st(float* p, const float* o) { *reinterpret_cast<float4*>(p) = make_float4(o[0], o[1], o[2], o[3]); }
};
template <> struct Vec<bf16> {
  static constexpr int N = 8;
  __device__ static void ld(const bf16* p, float* o) {
    uint4 v = *reinterpret_cast<const uint4*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float2 f = __bfloat1622float2(h[i]);
      o[2 * i] = f.x; o[2 * i + 1] = f.y;
    }
  }
  __device__ static void st(bf16* p, const float* o) {
    uint4 v;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(o[2 * i], o[2 * i + 1]);
    *reinterpret_cast<uint4*>(p) = v;
  }
};

// ---------------------------------------------------------------------------------------------
// border fill: buf [B][H+2][W+2][C]
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void border_fill_kernel(T* buf, int B, int H, int W, int C, int mode) {
  constexpr int V = Vec<T>::N;
  const int Wp = W + 2, Hp = H + 2;
  const int per_img = 2 * Wp + 2 * H;  // top row, bottom row, left col, right col (without corners)
  const int cv = C / V;
  long long total = (long long)B * per_img * cv;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % cv) * V;
    long long r = i / cv;
    int e = (int)(r % per_img);
    int b = (int)(r / per_img);
    int qy, qx;
    if (e < Wp) { qy = 0; qx = e; }
    else if (e < 2 * Wp) { qy = Hp - 1; qx = e - Wp; }
    else if (e < 2 * Wp + H) { qy = e - 2 * Wp + 1; qx = 0; }
    else { qy = e - 2 * Wp - H + 1; qx = Wp - 1; }
    int sy = pad_src_index(qy, H, mode) + 1, sx = pad_src_index(qx, W, mode) + 1;
    const T* src = buf + (((long long)b * Hp + sy) * Wp + sx) * C + c;
    T* dst = buf + (((long long)b * Hp + qy) * Wp + qx) * C + c;
    *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(src);
  }
}

// ---------------------------------------------------------------------------------------------
// pad fold (backward of the border fill) + residual + activation-derivative mask
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void pad_fold_kernel(const T* gpad, int B, int H, int W, int C, int mode, View resid, View mask,
                                const float* __restrict__ mslope, View out1, View out2) {
  constexpr int V = Vec<T>::N;
  const int Wp = W + 2, Hp = H + 2;
  const int cv = C / V;
  long long total = (long long)B * H * W * cv;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % cv) * V;
    long long r = i / cv;
    int x = (int)(r % W); r /= W;
    int y = (int)(r % H);
    int b = (int)(r / H);
    int qys[3], qxs[3], ny = 0, nx = 0;
    qys[ny++] = y + 1;
    if (pad_src_index(0, H, mode) == y) qys[ny++] = 0;
    if (pad_src_index(Hp - 1, H, mode) == y) qys[ny++] = Hp - 1;
    qxs[nx++] = x + 1;
    if (pad_src_index(0, W, mode) == x) qxs[nx++] = 0;
    if (pad_src_index(Wp - 1, W, mode) == x) qxs[nx++] = Wp - 1;
    float acc[V];
#pragma unroll
    for (int j = 0; j < V; ++j) acc[j] = 0.f;
    for (int a = 0; a < ny; ++a)
      for (int d = 0; d < nx; ++d) {
        float t[V];
        Vec<T>::ld(gpad + (((long long)b * Hp + qys[a]) * Wp + qxs[d]) * C + c, t);
#pragma unroll
        for (int j = 0; j < V; ++j) acc[j] += t[j];
      }
    if (resid.ptr) {
      float t[V];
      Vec<T>::ld((const T*)resid.ptr + view_off(resid, b, y, x) + c, t);
#pragma unroll
      for (int j = 0; j < V; ++j) acc[j] += t[j];
    }
    if (out1.ptr) Vec<T>::st((T*)out1.ptr + view_off(out1, b, y, x) + c, acc);
    if (out2.ptr) {
      if (mask.ptr) {
        float t[V];
        Vec<T>::ld((const T*)mask.ptr + view_off(mask, b, y, x) + c, t);
#pragma unroll
        for (int j = 0; j < V; ++j) acc[j] *= (t[j] > 0.f ? 1.f : (mslope ? mslope[c + j] : 0.f));
      }
      Vec<T>::st((T*)out2.ptr + view_off(out2, b, y, x) + c, acc);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// im2col 5x5 for the tiny-channel encoders
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void im2col5_kernel(const float* __restrict__ x, T* col, int B, int Cin, int H, int W, int Kpad, int mode) {
  long long total = (long long)B * H * W * Kpad;
  const int kreal = 25 * Cin;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int k = (int)(i % Kpad);
    long long p = i / Kpad;
    float v = 0.f;
    if (k < kreal) {
      int ci = k % Cin, t = k / Cin;
      int ky = t / 5, kx = t % 5;
      int px = (int)(p % W);
      long long r = p / W;
      int py = (int)(r % H);
      int b = (int)(r / H);
      int sy = pad_index(py + ky - 2, H, mode), sx = pad_index(px + kx - 2, W, mode);
      v = x[(((long long)b * Cin + ci) * H + sy) * W + sx];
    }
    col[i] = from_f<T>(v);
  }
}

// vectorised variant (Kpad % 8 == 0): one thread produces 8 consecutive k of one pixel and stores them with one
// (bf16) or two (fp32) 16-byte writes; a warp writes 512 contiguous bytes.  The tiny NCHW input stays in L1/L2.
template <typename T>
__global__ void im2col5_vec8_kernel(const float* __restrict__ x, T* col, int B, int Cin, int H, int W, int Kpad, int mode) {
  const int kv = Kpad / 8;
  long long total = (long long)B * H * W * kv;
  const int kreal = 25 * Cin;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int k0 = (int)(i % kv) * 8;
    long long p = i / kv;
    const int px = (int)(p % W);
    long long r = p / W;
    const int py = (int)(r % H);
    const int b = (int)(r / H);
    float v[8];
    int ci = k0 % Cin, t = k0 / Cin;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      v[j] = 0.f;
      if (k0 + j < kreal) {
        const int ky = t / 5, kx = t - ky * 5;
        const int sy = pad_index(py + ky - 2, H, mode), sx = pad_index(px + kx - 2, W, mode);
        v[j] = __ldg(x + (((long long)b * Cin + ci) * H + sy) * W + sx);
      }
      if (++ci == Cin) { ci = 0; ++t; }
    }
    T* dst = col + p * Kpad + k0;
    if (sizeof(T) == 2) {
      Vec<bf16>::st((bf16*)dst, v);
    } else {
      Vec<float>::st((float*)dst, v);
      Vec<float>::st((float*)dst + 4, v + 4);
    }
  }
}

// tiled variant: a block stages a (4+4) x (32+4) x Cin input tile (padding mode applied while loading) and a k ->
// tile-offset table in shared memory; every thread then emits 16-byte groups of 8 consecutive k with no integer
// division and no global reads in the inner loop.
constexpr int I2C_TH = 4, I2C_TW = 32;
template <typename T>
__global__ void __launch_bounds__(256) im2col5_tile_kernel(const float* __restrict__ x, T* col, int B, int Cin, int H, int W,
                                                           int Kpad, int mode) {
  extern __shared__ float i2c_sm[];
  float* tile = i2c_sm;                                                  // [(TH+4)][(TW+4)][Cin]
  int* koff = reinterpret_cast<int*>(tile + (I2C_TH + 4) * (I2C_TW + 4) * Cin);   // [Kpad]
  const int tiles_x = (W + I2C_TW - 1) / I2C_TW, tiles_y = (H + I2C_TH - 1) / I2C_TH;
  const int b = blockIdx.x / (tiles_x * tiles_y), tr = blockIdx.x % (tiles_x * tiles_y);
  const int y0 = (tr / tiles_x) * I2C_TH, x0 = (tr % tiles_x) * I2C_TW;
  const int TWp = I2C_TW + 4, n_in = (I2C_TH + 4) * TWp * Cin;
  for (int i = threadIdx.x; i < n_in; i += blockDim.x) {
    const int ci = i % Cin, q = i / Cin;
    const int ty = q / TWp, tx = q - ty * TWp;
    const int sy = pad_index(min(y0 + ty - 2, H + 1), H, mode), sx = pad_index(min(x0 + tx - 2, W + 1), W, mode);
    tile[i] = x[(((long long)b * Cin + ci) * H + sy) * W + sx];
  }
  const int kreal = 25 * Cin;
  for (int k = threadIdx.x; k < Kpad; k += blockDim.x) {
    int off = -1;
    if (k < kreal) {
      const int ci = k % Cin, t = k / Cin;
      off = ((t / 5) * TWp + (t % 5)) * Cin + ci;
    }
    koff[k] = off;
  }
  __syncthreads();
  const int kv = Kpad / 8, items = I2C_TH * I2C_TW * kv;
  for (int it = threadIdx.x; it < items; it += blockDim.x) {
    const int g = it % kv, p = it / kv;
    const int py = p / I2C_TW, px = p - py * I2C_TW;
    if (y0 + py >= H || x0 + px >= W) continue;
    const float* base = tile + (py * TWp + px) * Cin;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int o = koff[g * 8 + j];
      v[j] = o >= 0 ? base[o] : 0.f;
    }
    T* dst = col + (((long long)b * H + y0 + py) * W + x0 + px) * Kpad + g * 8;
    if (sizeof(T) == 2) {
      Vec<bf16>::st((bf16*)dst, v);
    } else {
      Vec<float>::st((float*)dst, v);
      Vec<float>::st((float*)dst + 4, v + 4);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// decoder tail (256 -> 3, zero padding) forward / data-grad / weight-grad
// ---------------------------------------------------------------------------------------------
// one warp per pixel; lanes split the channels; weights [3][9][C] staged in smem
template <typename T>
__global__ void dec_tail_fwd_kernel(View h, const float* __restrict__ w, const float* __restrict__ bias,
                                    const float* __restrict__ x, float* __restrict__ out, int B, int H, int W) {
  extern __shared__ float sw[];  // [27][C]
  const int C = h.C;
  for (int i = threadIdx.x; i < 27 * C; i += blockDim.x) sw[i] = w[i];
  __syncthreads();
  constexpr int V = Vec<T>::N;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  long long npx = (long long)B * H * W;
  for (long long p = (long long)blockIdx.x * wpb + warp; p < npx; p += (long long)gridDim.x * wpb) {
    int px = (int)(p % W);
    long long r = p / W;
    int py = (int)(r % H);
    int b = (int)(r / H);
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
    for (int t = 0; t < 9; ++t) {
      int yy = py + t / 3 - 1 + h.oy, xx = px + t % 3 - 1 + h.ox;
      if (!view_inb(h, yy, xx)) continue;
      const T* hp = (const T*)h.ptr + view_off(h, b, yy, xx);
      for (int c = lane * V; c < C; c += 32 * V) {
        float hv[V];
        Vec<T>::ld(hp + c, hv);
#pragma unroll
        for (int j = 0; j < V; ++j) {
          a0 += hv[j] * sw[(0 * 9 + t) * C + c + j];
          a1 += hv[j] * sw[(1 * 9 + t) * C + c + j];
          a2 += hv[j] * sw[(2 * 9 + t) * C + c + j];
        }
      }
    }
    a0 = warp_sum(a0); a1 = warp_sum(a1); a2 = warp_sum(a2);
    if (lane < 3) {
      float a = lane == 0 ? a0 : (lane == 1 ? a1 : a2);
      long long o = (((long long)b * 3 + lane) * H + py) * W + px;
      out[o] = a + bias[lane] + x[o];
    }
  }
}

template <typename T>
__global__ void dec_tail_bwd_data_kernel(const float* __restrict__ dout, const float* __restrict__ w, View h,
                                         View dh, int B, int H, int W) {
  extern __shared__ float sw[];  // [27][C]
  const int C = h.C;
  for (int i = threadIdx.x; i < 27 * C; i += blockDim.x) sw[i] = w[i];
  __syncthreads();
  constexpr int V = Vec<T>::N;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  long long npx = (long long)B * H * W;
  for (long long p = (long long)blockIdx.x * wpb + warp; p < npx; p += (long long)gridDim.x * wpb) {
    int px = (int)(p % W);
    long long r = p / W;
    int py = (int)(r % H);
    int b = (int)(r / H);
    // dh[p,c] = sum_t sum_co dout(p - tap_t)[co] * w[co][t][c]
    float d[27];
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      int yy = py - (t / 3 - 1), xx = px - (t % 3 - 1);
      bool in = (unsigned)yy < (unsigned)H && (unsigned)xx < (unsigned)W;
#pragma unroll
      for (int co = 0; co < 3; ++co) d[co * 9 + t] = in ? dout[(((long long)b * 3 + co) * H + yy) * W + xx] : 0.f;
    }
    for (int c = lane * V; c < C; c += 32 * V) {
      float acc[V], hv[V];
#pragma unroll
      for (int j = 0; j < V; ++j) acc[j] = 0.f;
#pragma unroll
      for (int q = 0; q < 27; ++q) {
#pragma unroll
        for (int j = 0; j < V; ++j) acc[j] += d[q] * sw[q * C + c + j];
      }
      Vec<T>::ld((const T*)h.ptr + view_off(h, b, py + h.oy, px + h.ox) + c, hv);
#pragma unroll
      for (int j = 0; j < V; ++j) acc[j] = hv[j] > 0.f ? acc[j] : 0.f;
      Vec<T>::st((T*)dh.ptr + view_off(dh, b, py + dh.oy, px + dh.ox) + c, acc);
    }
  }
}

// thread per channel, block per pixel chunk; partial [nblk][27][C] then reduce
template <typename T>
__global__ void dec_tail_bwd_weight_partial(const float* __restrict__ dout, View h, float* __restrict__ part,
                                            float* __restrict__ bpart, int B, int H, int W, int chunk) {
  const int C = h.C;
  long long npx = (long long)B * H * W;
  long long p0 = (long long)blockIdx.x * chunk, p1 = p0 + chunk < npx ? p0 + chunk : npx;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float acc[27];
#pragma unroll
    for (int q = 0; q < 27; ++q) acc[q] = 0.f;
    for (long long p = p0; p < p1; ++p) {
      int px = (int)(p % W);
      long long r = p / W;
      int py = (int)(r % H);
      int b = (int)(r / H);
      float d0 = dout[(((long long)b * 3 + 0) * H + py) * W + px];
      float d1 = dout[(((long long)b * 3 + 1) * H + py) * W + px];
      float d2 = dout[(((long long)b * 3 + 2) * H + py) * W + px];
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        int yy = py + t / 3 - 1 + h.oy, xx = px + t % 3 - 1 + h.ox;
        float hv = view_inb(h, yy, xx) ? to_f<T>(((const T*)h.ptr)[view_off(h, b, yy, xx) + c]) : 0.f;
        acc[0 * 9 + t] += d0 * hv;
        acc[1 * 9 + t] += d1 * hv;
        acc[2 * 9 + t] += d2 * hv;
      }
    }
#pragma unroll
    for (int q = 0; q < 27; ++q) part[((long long)blockIdx.x * 27 + q) * C + c] = acc[q];
  }
  if (threadIdx.x < 3) {
    float s = 0.f;
    for (long long p = p0; p < p1; ++p) {
      int px = (int)(p % W);
      long long r = p / W;
      int py = (int)(r % H);
      int b = (int)(r / H);
      s += dout[(((long long)b * 3 + threadIdx.x) * H + py) * W + px];
    }
    bpart[blockIdx.x * 3 + threadIdx.x] = s;
  }
}

__global__ void reduce_partials_kernel(const float* __restrict__ part, float* __restrict__ out, int nblk, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float s = 0.f;
  for (int b = 0; b < nblk; ++b) s += part[(long long)b * n + i];
  out[i] = s;
}

// ---------------------------------------------------------------------------------------------
// L1 loss fused fwd + bwd; deterministic two-level reduction (ticket pattern)
// ---------------------------------------------------------------------------------------------
#define L1_MAX_BLOCKS 2048
// scratch = [L1_MAX_BLOCKS partials | ticket], private to the launching (device, stream): see stream_scratch()
__global__ void l1_kernel(const float* __restrict__ a, const float* __restrict__ b, long long n, float gscale,
                          float* __restrict__ loss, float* __restrict__ grad, float* __restrict__ g_l1_partials) {
  unsigned int* g_l1_ticket = reinterpret_cast<unsigned int*>(g_l1_partials + L1_MAX_BLOCKS);
  float s = 0.f;
  const float gs = gscale / (float)n;
  long long n4 = n >> 2;
  const float4* a4 = reinterpret_cast<const float4*>(a);
  const float4* b4 = reinterpret_cast<const float4*>(b);
  float4* g4 = reinterpret_cast<float4*>(grad);
  auto one = [&](const float4& x, const float4& y, long long i) {
    float d0 = x.x - y.x, d1 = x.y - y.y, d2 = x.z - y.z, d3 = x.w - y.w;
    s += fabsf(d0) + fabsf(d1) + fabsf(d2) + fabsf(d3);
    if (grad) {
      float4 g;
      g.x = d0 > 0.f ? gs : (d0 < 0.f ? -gs : 0.f);
      g.y = d1 > 0.f ? gs : (d1 < 0.f ? -gs : 0.f);
      g.z = d2 > 0.f ? gs : (d2 < 0.f ? -gs : 0.f);
      g.w = d3 > 0.f ? gs : (d3 < 0.f ? -gs : 0.f);
      g4[i] = g;
    }
  };
  const long long stride = (long long)gridDim.x * blockDim.x;
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  for (; i + 3 * stride < n4; i += 4 * stride) {   // 8 independent 16-byte loads in flight per thread
    float4 x0 = a4[i], x1 = a4[i + stride], x2 = a4[i + 2 * stride], x3 = a4[i + 3 * stride];
    float4 y0 = b4[i], y1 = b4[i + stride], y2 = b4[i + 2 * stride], y3 = b4[i + 3 * stride];
    one(x0, y0, i); one(x1, y1, i + stride); one(x2, y2, i + 2 * stride); one(x3, y3, i + 3 * stride);
  }
  for (; i < n4; i += stride) one(a4[i], b4[i], i);
  for (long long i = (n4 << 2) + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float d = a[i] - b[i];
    s += fabsf(d);
    if (grad) grad[i] = d > 0.f ? gs : (d < 0.f ? -gs : 0.f);
  }
  __shared__ float ws[32];
  __shared__ bool last;
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = threadIdx.x < (blockDim.x >> 5) ? ws[threadIdx.x] : 0.f;
    t = warp_sum(t);
    if (threadIdx.x == 0) {
      g_l1_partials[blockIdx.x] = t;
      __threadfence();
      unsigned int tk = atomicAdd(g_l1_ticket, 1u);
      last = (tk == gridDim.x - 1);
    }
  }
  __syncthreads();
  if (last && threadIdx.x < 32) {
    __threadfence();
    double t = 0.0;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += 32) t += (double)((volatile float*)g_l1_partials)[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (threadIdx.x == 0) {
      loss[0] = (float)(t / (double)n);
      *g_l1_ticket = 0;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// preprocessing (+ optional crop)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float prep_normal(float v) {
  // np.nan_to_num: nan -> 0, +-inf -> +-FLT_MAX; then (v+1)/2 clamped to [0,1]
  if (isnan(v)) v = 0.f;
  else if (isinf(v)) v = v > 0.f ? 3.4028234664e38f : -3.4028234664e38f;
  v = (v + 1.0f) * 0.5f;
  return fmaxf(fminf(v, 1.0f), 0.0f);
}

// frames NHWC [Hf][Wf][C] (or patches [n][PH][PW][C] when centres == NULL) -> NCHW [n][C][PH][PW].
// ONE launch for noisy + gt + aux, one thread per output pixel.  A warp reads 32 consecutive pixels =
// 384 / 896 contiguous bytes per tensor (every fetched sector fully used) and writes one coalesced 128-byte row per
// channel plane.
__global__ void __launch_bounds__(256) crop_preprocess_px_kernel(const float* __restrict__ noisy, const float* __restrict__ gt,
                                                                 const float* __restrict__ aux, float* __restrict__ noisy_o,
                                                                 float* __restrict__ gt_o, float* __restrict__ aux_o, int Hf,
                                                                 int Wf, const int* __restrict__ centres,
                                                                 const int* __restrict__ img_idx, int n, int P, int PW) {
  const long long pp = (long long)P * PW, total = (long long)n * pp;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / pp);
    const int r = (int)(i - (long long)b * pp);
    const int y = r / PW, x = r - y * PW;
    long long sp;   // source pixel index
    if (centres) {
      const int cx = centres[2 * b], cy = centres[2 * b + 1];
      const long long f = img_idx ? (long long)img_idx[b] * Hf * Wf : 0;
      sp = f + (long long)(cy - P / 2 + y) * Wf + (cx - PW / 2 + x);
    } else {
      sp = i;
    }
    const float* s3 = noisy + sp * 3;
    float* d3 = noisy_o + (long long)b * 3 * pp + r;
#pragma unroll
    for (int c = 0; c < 3; ++c) __stcs(d3 + c * pp, logf(__ldcs(s3 + c) + 1.0f));
    if (gt) {
      const float* g3 = gt + sp * 3;
      float* e3 = gt_o + (long long)b * 3 * pp + r;
#pragma unroll
      for (int c = 0; c < 3; ++c) __stcs(e3 + c * pp, logf(__ldcs(g3 + c) + 1.0f));
    }
    const float* s7 = aux + sp * 7;
    float* d7 = aux_o + (long long)b * 7 * pp + r;
#pragma unroll
    for (int c = 0; c < 7; ++c) {
      const float v = __ldcs(s7 + c);
      __stcs(d7 + c * pp, c < 3 ? prep_normal(v) : v);
    }
  }
}

// 4 consecutive pixels of a row per thread (PW % 4 == 0): 12 / 28 contiguous input floats (16-byte loads when the
// source pixel index is a multiple of 4, which always holds without a crop), one 16-byte store per channel plane.
template <int n>
__device__ __forceinline__ void ld_run(const float* p, float* o, bool aligned) {
  if (aligned) {
#pragma unroll
    for (int i = 0; i < n; i += 4) {
      const float4 v = __ldcs(reinterpret_cast<const float4*>(p + i));
      o[i] = v.x; o[i + 1] = v.y; o[i + 2] = v.z; o[i + 3] = v.w;
    }
  } else {
#pragma unroll
    for (int i = 0; i < n; ++i) o[i] = __ldcs(p + i);
  }
}
__global__ void __launch_bounds__(256) crop_preprocess_px4_kernel(const float* __restrict__ noisy, const float* __restrict__ gt,
                                                                  const float* __restrict__ aux, float* __restrict__ noisy_o,
                                                                  float* __restrict__ gt_o, float* __restrict__ aux_o, int Hf,
                                                                  int Wf, const int* __restrict__ centres,
                                                                  const int* __restrict__ img_idx, int n, int P, int PW) {
  const long long pp = (long long)P * PW, q4 = pp >> 2, total = (long long)n * q4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / q4);
    const int r = (int)(i - (long long)b * q4) << 2;      // first of 4 pixels inside the patch
    const int y = r / PW, x = r - y * PW;
    long long sp;
    if (centres) {
      const int cx = centres[2 * b], cy = centres[2 * b + 1];
      const long long f = img_idx ? (long long)img_idx[b] * Hf * Wf : 0;
      sp = f + (long long)(cy - P / 2 + y) * Wf + (cx - PW / 2 + x);
    } else {
      sp = (long long)b * pp + r;
    }
    const bool al = (sp & 3) == 0;
    float v[28];
    ld_run<12>(noisy + sp * 3, v, al);
    float* d3 = noisy_o + (long long)b * 3 * pp + r;
#pragma unroll
    for (int c = 0; c < 3; ++c)
      __stcs(reinterpret_cast<float4*>(d3 + c * pp), make_float4(logf(v[c] + 1.0f), logf(v[3 + c] + 1.0f), logf(v[6 + c] + 1.0f),
                                                                logf(v[9 + c] + 1.0f)));
    if (gt) {
      ld_run<12>(gt + sp * 3, v, al);
      float* e3 = gt_o + (long long)b * 3 * pp + r;
#pragma unroll
      for (int c = 0; c < 3; ++c)
        __stcs(reinterpret_cast<float4*>(e3 + c * pp), make_float4(logf(v[c] + 1.0f), logf(v[3 + c] + 1.0f), logf(v[6 + c] + 1.0f),
                                                                  logf(v[9 + c] + 1.0f)));
    }
    ld_run<28>(aux + sp * 7, v, al);
    float* d7 = aux_o + (long long)b * 7 * pp + r;
#pragma unroll
    for (int c = 0; c < 7; ++c) {
      float4 o = make_float4(v[c], v[7 + c], v[14 + c], v[21 + c]);
      if (c < 3) o = make_float4(prep_normal(o.x), prep_normal(o.y), prep_normal(o.z), prep_normal(o.w));
      __stcs(reinterpret_cast<float4*>(d7 + c * pp), o);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Adam (flat arena)
// ---------------------------------------------------------------------------------------------
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, long long n, float step_size, float beta1, float beta2, float eps,
                            float inv_sqrt_bc2, float gscale) {
  long long n4 = n >> 2;
  float4* p4 = reinterpret_cast<float4*>(p);
  const float4* g4 = reinterpret_cast<const float4*>(g);
  float4* m4 = reinterpret_cast<float4*>(m);
  float4* v4 = reinterpret_cast<float4*>(v);
  auto upd = [&](float& pp, float gg, float& mm, float& vv) {
    gg *= gscale;
    mm = mm + (1.0f - beta1) * (gg - mm);            // torch: exp_avg.lerp_(grad, 1 - beta1)
    vv = vv * beta2 + (1.0f - beta2) * gg * gg;      // exp_avg_sq.mul_(beta2).addcmul_(g, g, 1 - beta2)
    float denom = sqrtf(vv) * inv_sqrt_bc2 + eps;    // (sqrt(v) / sqrt(bc2)) + eps
    pp = pp - step_size * (mm / denom);              // p.addcdiv_(m, denom, value=-lr/bc1)
  };
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 pp = p4[i], gg = g4[i], mm = m4[i], vv = v4[i];
    upd(pp.x, gg.x, mm.x, vv.x); upd(pp.y, gg.y, mm.y, vv.y); upd(pp.z, gg.z, mm.z, vv.z); upd(pp.w, gg.w, mm.w, vv.w);
    p4[i] = pp; m4[i] = mm; v4[i] = vv;
  }
  for (long long i = (n4 << 2) + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    upd(p[i], g[i], m[i], v[i]);
}

// Device-resident hyper-parameters (CUDA-graph friendly: nothing step-dependent is a kernel argument).
// hyper = {lr, step (int32 bits), step_size, 1/sqrt(bc2)}: the prelude bumps the step count and derives the two
// bias-correction factors in fp64 exactly like pht_adam does on the host.
__global__ void adam_hyper_kernel(float* __restrict__ hyper, float beta1, float beta2) {
  const int step = __float_as_int(hyper[1]) + 1;
  hyper[1] = __int_as_float(step);
  const double bc1 = 1.0 - pow((double)beta1, (double)step);
  const double bc2 = 1.0 - pow((double)beta2, (double)step);
  hyper[2] = (float)((double)hyper[0] / bc1);
  hyper[3] = (float)(1.0 / sqrt(bc2));
}
__global__ void adam_dev_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                float* __restrict__ v, long long n, const float* __restrict__ hyper, float beta1, float beta2,
                                float eps, float gscale) {
  const float step_size = hyper[2], inv_sqrt_bc2 = hyper[3];
  long long n4 = n >> 2;
  float4* p4 = reinterpret_cast<float4*>(p);
  const float4* g4 = reinterpret_cast<const float4*>(g);
  float4* m4 = reinterpret_cast<float4*>(m);
  float4* v4 = reinterpret_cast<float4*>(v);
  auto upd = [&](float& pp, float gg, float& mm, float& vv) {
    gg *= gscale;
    mm = mm + (1.0f - beta1) * (gg - mm);
    vv = vv * beta2 + (1.0f - beta2) * gg * gg;
    float denom = sqrtf(vv) * inv_sqrt_bc2 + eps;
    pp = pp - step_size * (mm / denom);
  };
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 pp = p4[i], gg = g4[i], mm = m4[i], vv = v4[i];
    upd(pp.x, gg.x, mm.x, vv.x); upd(pp.y, gg.y, mm.y, vv.y); upd(pp.z, gg.z, mm.z, vv.z); upd(pp.w, gg.w, mm.w, vv.w);
    p4[i] = pp; m4[i] = mm; v4[i] = vv;
  }
  for (long long i = (n4 << 2) + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    upd(p[i], g[i], m[i], v[i]);
}

// ---------------------------------------------------------------------------------------------
// weight pack / unpack
// ---------------------------------------------------------------------------------------------
struct PackP {
  int O, I, ks, Ntot, Ktot, n_off, k_off, transpose, grid, i_begin, i_count;
  float scale;
};
__device__ __forceinline__ long long packed_index(const PackP& a, int o, int i, int ky, int kx) {
  if (a.grid > 0) {
    int e = (a.grid - a.ks) / 2;
    int k = a.k_off + ((ky + e) * a.grid + (kx + e)) * a.i_count + i;
    return (long long)(a.n_off + o) * a.Ktot + k;
  }
  int t = ky * a.ks + kx, T = a.ks * a.ks;
  if (a.transpose == 2) return (long long)(a.n_off + i) * a.Ktot + a.k_off + t * a.O + o;  // taps x outputs folded into K
  if (!a.transpose) return ((long long)t * a.Ntot + a.n_off + o) * a.Ktot + a.k_off + i;
  return ((long long)(T - 1 - t) * a.Ntot + a.n_off + i) * a.Ktot + a.k_off + o;
}
template <typename T>
__global__ void pack_weight_kernel(const float* __restrict__ w, T* dst, PackP a) {
  long long total = (long long)a.O * a.I * a.ks * a.ks;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    int kx = (int)(idx % a.ks);
    long long r = idx / a.ks;
    int ky = (int)(r % a.ks); r /= a.ks;
    int i = (int)(r % a.I);
    int o = (int)(r / a.I);
    if (i < a.i_begin || i >= a.i_begin + a.i_count) continue;
    dst[packed_index(a, o, i - a.i_begin, ky, kx)] = from_f<T>(w[idx] * a.scale);
  }
}
__global__ void unpack_wgrad_kernel(float* __restrict__ wg, const float* __restrict__ src, PackP a) {
  long long total = (long long)a.O * a.I * a.ks * a.ks;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    int kx = (int)(idx % a.ks);
    long long r = idx / a.ks;
    int ky = (int)(r % a.ks); r /= a.ks;
    int i = (int)(r % a.I);
    int o = (int)(r / a.I);
    if (i < a.i_begin || i >= a.i_begin + a.i_count) continue;
    wg[idx] = src[packed_index(a, o, i - a.i_begin, ky, kx)] * a.scale;
  }
}

// all weight (re)packs of a step in ONE launch: blockIdx.y selects the descriptor (table lives in device memory)
struct PackJob {
  const float* w;
  void* dst;
  PackP p;
  int dtype;
  int pad_;
};
// One thread per (o, i) pair, looping over the taps: the thread re-reads its own contiguous OIHW taps (L1 hits after the
// first) and the warp's stores are contiguous in the packed layout for every tap (the fastest thread index follows the
// packed layout's fastest index: i for the forward [tap][N][K] pack, o for the transposed / tap-folded packs).
__global__ void pack_weights_batched_kernel(const PackJob* __restrict__ jobs) {
  const PackJob j = jobs[blockIdx.y];
  const PackP a = j.p;
  const int T = a.ks * a.ks;
  const long long pairs = (long long)a.O * a.i_count;
  const bool o_fast = a.transpose != 0 && a.grid <= 0;
  for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < pairs; p += (long long)gridDim.x * blockDim.x) {
    int o, i;
    if (o_fast) { o = (int)(p % a.O); i = (int)(p / a.O); }
    else { i = (int)(p % a.i_count); o = (int)(p / a.i_count); }
    const float* src = j.w + ((long long)o * a.I + a.i_begin + i) * T;
    for (int t = 0; t < T; ++t) {
      const int ky = t / a.ks, kx = t - ky * a.ks;
      const float v = __ldg(src + t) * a.scale;
      const long long di = packed_index(a, o, i, ky, kx);
      if (j.dtype == PHT_F32) ((float*)j.dst)[di] = v;
      else ((bf16*)j.dst)[di] = __float2bfloat16_rn(v);
    }
  }
}

// Tiled version (the default): one CTA per 32 (o) x 32 (i) x taps tile of one job.  The tile's OIHW rows are read as
// contiguous runs (a row o holds ni * taps consecutive floats) into shared memory; the packed layout is then written
// tap by tap with the lanes of a warp on the layout's fastest index (i for the forward [tap][N][K] pack, o for the
// transposed / tap-folded packs): both sides are coalesced, and no thread walks 9 strided addresses.  PackJob.pad_
// holds the job's first tile index (ascending), the kernel finds its job by binary search.
constexpr int PK_TILE = 32;
__global__ void __launch_bounds__(256) pack_weights_tiled_kernel(const PackJob* __restrict__ jobs, int njobs) {
  extern __shared__ float pk_tile[];
  int lo = 0, hi = njobs - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (jobs[mid].pad_ <= (int)blockIdx.x) lo = mid;
    else hi = mid - 1;
  }
  const PackJob j = jobs[lo];
  const PackP a = j.p;
  const int T = a.ks * a.ks;
  const int tiles_i = (a.i_count + PK_TILE - 1) / PK_TILE;
  const int tl = (int)blockIdx.x - j.pad_;
  const int o0 = (tl / tiles_i) * PK_TILE, i0 = (tl % tiles_i) * PK_TILE;
  const int no = min(PK_TILE, a.O - o0), ni = min(PK_TILE, a.i_count - i0);
  const int rowlen = ni * T, pitch = min(PK_TILE, a.i_count) * T + 1;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int r = warp; r < no; r += 8) {
    const float* src = j.w + ((long long)(o0 + r) * a.I + a.i_begin + i0) * T;
    for (int c = lane; c < rowlen; c += 32) pk_tile[r * pitch + c] = __ldg(src + c) * a.scale;
  }
  __syncthreads();
  const bool o_fast = a.transpose != 0 && a.grid <= 0;
  const int nf = o_fast ? no : ni, ns = o_fast ? ni : no;
  for (int r = warp; r < T * ns; r += 8) {
    const int t = r / ns, sl = r - t * ns;
    if (lane < nf) {
      const int ol = o_fast ? lane : sl, il = o_fast ? sl : lane;
      const float v = pk_tile[ol * pitch + il * T + t];
      const int ky = t / a.ks, kx = t - ky * a.ks;
      const long long di = packed_index(a, o0 + ol, i0 + il, ky, kx);
      if (j.dtype == PHT_F32) ((float*)j.dst)[di] = v;
      else ((bf16*)j.dst)[di] = __float2bfloat16_rn(v);
    }
  }
}

__global__ void unpack_wgrads_batched_kernel(const PackJob* __restrict__ jobs) {
  const PackJob j = jobs[blockIdx.y];
  const PackP a = j.p;
  long long total = (long long)a.O * a.I * a.ks * a.ks;
  float* wg = const_cast<float*>(j.w);
  const float* src = (const float*)j.dst;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    int kx = (int)(idx % a.ks);
    long long r = idx / a.ks;
    int ky = (int)(r % a.ks); r /= a.ks;
    int i = (int)(r % a.I);
    int o = (int)(r / a.I);
    if (i < a.i_begin || i >= a.i_begin + a.i_count) continue;
    wg[idx] = src[packed_index(a, o, i - a.i_begin, ky, kx)] * a.scale;
  }
}

// decoder tail on the GEMM path: out = y[..., 0:3] + bias + x  (y = fp32 [B,H,W,64] conv result, NHWC -> NCHW)
__global__ void tail_finish_kernel(const float* __restrict__ y, int ldy, const float* __restrict__ bias,
                                   const float* __restrict__ x, float* __restrict__ out, int B, int H, int W) {
  long long total = (long long)B * 3 * H * W;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int px = (int)(i % W);
    long long r = i / W;
    int py = (int)(r % H); r /= H;
    int co = (int)(r % 3);
    int b = (int)(r / 3);
    out[i] = y[(((long long)b * H + py) * W + px) * ldy + co] + bias[co] + x[i];
  }
}
// a[p][t*3+co] = dout[p - tap_t][co] (zero outside the image), columns 27..63 zero (the bias gradient is tail_dbias_kernel)
__global__ void tail_im2col_bwd_kernel(const float* __restrict__ dout, bf16* __restrict__ a, int B, int H, int W) {
  const long long npx = (long long)B * H * W;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < npx * 8; i += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(i & 7);          // 8-column group of the 64-wide row
    const long long p = i >> 3;
    const int px = (int)(p % W);
    long long r = p / W;
    const int py = (int)(r % H), b = (int)(r / H);
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int col = g * 8 + j;
      v[j] = 0.f;
      if (col < 27) {
        const int t = col / 3, co = col % 3;
        const int yy = py - (t / 3 - 1), xx = px - (t % 3 - 1);
        if ((unsigned)yy < (unsigned)H && (unsigned)xx < (unsigned)W) v[j] = dout[(((long long)b * 3 + co) * H + yy) * W + xx];
      }
    }
    Vec<bf16>::st(a + p * 64 + g * 8, v);
  }
}
// dbias[co] = sum over the batch of dout[b][co][:, :].  grid (TAIL_DB_BLOCKS, 3): every block sums a fixed slice, the
// last block to finish (ticket) adds the block partials in block order -> deterministic regardless of scheduling.
constexpr int TAIL_DB_BLOCKS = 48;
// scratch = [3 * TAIL_DB_BLOCKS partials | 3 tickets], private to the launching (device, stream)
__global__ void __launch_bounds__(256) tail_dbias_kernel(const float* __restrict__ dout, float* __restrict__ dbias, int B, int HW,
                                                         float* __restrict__ g_tail_db_part) {
  unsigned int* g_tail_db_ticket = reinterpret_cast<unsigned int*>(g_tail_db_part + 3 * TAIL_DB_BLOCKS);
  const int co = blockIdx.y;
  float s = 0.f;
  for (int b = 0; b < B; ++b) {
    const float* src = dout + ((long long)b * 3 + co) * HW;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += gridDim.x * blockDim.x) s += src[i];
  }
  __shared__ float ws[8];
  __shared__ bool last;
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += ws[w];
    g_tail_db_part[co * TAIL_DB_BLOCKS + blockIdx.x] = t;
    __threadfence();
    last = atomicAdd(&g_tail_db_ticket[co], 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (last && threadIdx.x == 0) {
    __threadfence();
    float t = 0.f;
    for (int i = 0; i < (int)gridDim.x; ++i) t += ((volatile float*)g_tail_db_part)[co * TAIL_DB_BLOCKS + i];
    dbias[co] = t;
    g_tail_db_ticket[co] = 0;
  }
}

template <typename S, typename D>
__global__ void cast_kernel(const S* __restrict__ s, long long sld, D* __restrict__ d, long long dld, long long rows,
                            long long cols) {
  long long n = rows * cols;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    long long r = i / cols, c = i % cols;
    d[r * dld + c] = from_f<D>(to_f<S>(s[r * sld + c]));
  }
}

static int grid_for(long long work_items, int threads) {
  long long blocks = (work_items + threads - 1) / threads;
  long long cap = (long long)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

// ---------------------------------------------------------------------------------------------
// FiLM modulation (reference: pht/models/afgsa/film.py:36-45, spatial gamma / beta):
//   forward   out = gb[..., :C] * x + gb[..., C:]
//   backward  dgb[..., :C] = dout * x,  dgb[..., C:] = dout,  dx = dx_in + gb[..., :C] * dout
// one thread per (pixel, vector of channels); HBM-bound
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void film_fwd_kernel(View gb, View x, View out, int B, int H, int W, int C) {
  constexpr int V = Vec<T>::N;
  const int cv = C / V;
  const long long total = (long long)B * H * W * cv;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % cv) * V;
    long long p = i / cv;
    const int px = (int)(p % W);
    p /= W;
    const int py = (int)(p % H), b = (int)(p / H);
    float g[V], be[V], xv[V], o[V];
    Vec<T>::ld((const T*)gb.ptr + view_off(gb, b, py + gb.oy, px + gb.ox) + c, g);
    Vec<T>::ld((const T*)gb.ptr + view_off(gb, b, py + gb.oy, px + gb.ox) + C + c, be);
    Vec<T>::ld((const T*)x.ptr + view_off(x, b, py + x.oy, px + x.ox) + c, xv);
#pragma unroll
    for (int j = 0; j < V; ++j) o[j] = fmaf(g[j], xv[j], be[j]);
    Vec<T>::st((T*)out.ptr + view_off(out, b, py + out.oy, px + out.ox) + c, o);
  }
}
template <typename T>
__global__ void film_bwd_kernel(View gb, View x, View dout, View dgb, View dx_in, View dx, int B, int H, int W, int C) {
  constexpr int V = Vec<T>::N;
  const int cv = C / V;
  const long long total = (long long)B * H * W * cv;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % cv) * V;
    long long p = i / cv;
    const int px = (int)(p % W);
    p /= W;
    const int py = (int)(p % H), b = (int)(p / H);
    float g[V], xv[V], d[V], t[V];
    Vec<T>::ld((const T*)gb.ptr + view_off(gb, b, py + gb.oy, px + gb.ox) + c, g);
    Vec<T>::ld((const T*)x.ptr + view_off(x, b, py + x.oy, px + x.ox) + c, xv);
    Vec<T>::ld((const T*)dout.ptr + view_off(dout, b, py + dout.oy, px + dout.ox) + c, d);
#pragma unroll
    for (int j = 0; j < V; ++j) t[j] = d[j] * xv[j];
    T* dg = (T*)dgb.ptr + view_off(dgb, b, py + dgb.oy, px + dgb.ox);
    Vec<T>::st(dg + c, t);
    Vec<T>::st(dg + C + c, d);
    if (dx.ptr) {
#pragma unroll
      for (int j = 0; j < V; ++j) t[j] = 0.f;
      if (dx_in.ptr) Vec<T>::ld((const T*)dx_in.ptr + view_off(dx_in, b, py + dx_in.oy, px + dx_in.ox) + c, t);
#pragma unroll
      for (int j = 0; j < V; ++j) t[j] = fmaf(g[j], d[j], t[j]);
      Vec<T>::st((T*)dx.ptr + view_off(dx, b, py + dx.oy, px + dx.ox) + c, t);
    }
  }
}

}  // namespace pht

using namespace pht;

extern "C" {

static bool film_view_ok(const pht_view* v, int dtype, int C, int vec) {
  return v && v->ptr && v->dtype == dtype && v->C >= C && v->sx % vec == 0 && v->sy % vec == 0 && v->sb % vec == 0 &&
         ((uintptr_t)v->ptr & 15) == 0;
}
int pht_film_fwd(const pht_view* gb, const pht_view* x, const pht_view* out, int32_t B, int32_t C, void* stream) {
  PHT_CHECK_ARG(gb && x && out && B > 0 && C > 0, "film_fwd: bad args");
  const int dt = x->dtype, vec = dt == PHT_F32 ? 4 : 8;
  PHT_CHECK_ARG(C % vec == 0 && film_view_ok(gb, dt, 2 * C, vec) && film_view_ok(x, dt, C, vec) && film_view_ok(out, dt, C, vec),
                "film_fwd: views must share the dtype, be 16-byte aligned and hold 2C / C / C channels");
  const int H = x->H, W = x->W;
  const long long items = (long long)B * H * W * (C / vec);
  cudaStream_t st = (cudaStream_t)stream;
  if (dt == PHT_F32) film_fwd_kernel<float><<<grid_for(items, 256), 256, 0, st>>>(make_view(gb), make_view(x), make_view(out), B, H, W, C);
  else film_fwd_kernel<bf16><<<grid_for(items, 256), 256, 0, st>>>(make_view(gb), make_view(x), make_view(out), B, H, W, C);
  count_launch(CNT_OTHER);
  PHT_LAUNCH_CHECK();
  return PHT_OK;
}
int pht_film_bwd(const pht_view* gb, const pht_view* x, const pht_view* dout, const pht_view* dgb, const pht_view* dx_in,
                 const pht_view* dx, int32_t B, int32_t C, void* stream) {
  PHT_CHECK_ARG(gb && x && dout && dgb && B > 0 && C > 0, "film_bwd: bad args");
  const int dt = x->dtype, vec = dt == PHT_F32 ? 4 : 8;
  PHT_CHECK_ARG(C % vec == 0 && film_view_ok(gb, dt, 2 * C, vec) && film_view_ok(x, dt, C, vec) && film_view_ok(dout, dt, C, vec) &&
                    film_view_ok(dgb, dt, 2 * C, vec), "film_bwd: bad views");
  PHT_CHECK_ARG(!(dx && dx->ptr) || film_view_ok(dx, dt, C, vec), "film_bwd: bad dx view");
  PHT_CHECK_ARG(!(dx_in && dx_in->ptr) || film_view_ok(dx_in, dt, C, vec), "film_bwd: bad dx_in view");
  const int H = x->H, W = x->W;
  const long long items = (long long)B * H * W * (C / vec);
  cudaStream_t st = (cudaStream_t)stream;
  if (dt == PHT_F32)
    film_bwd_kernel<float><<<grid_for(items, 256), 256, 0, st>>>(make_view(gb), make_view(x), make_view(dout), make_view(dgb),
                                                                 make_view(dx_in), make_view(dx), B, H, W, C);
  else
    film_bwd_kernel<bf16><<<grid_for(items, 256), 256, 0, st>>>(make_view(gb), make_view(x), make_view(dout), make_view(dgb),
                                                                make_view(dx_in), make_view(dx), B, H, W, C);
  count_launch(CNT_OTHER);
  PHT_LAUNCH_CHECK();
  return PHT_OK;
}

int pht_border_fill(void* buf, int32_t dtype, int32_t B, int32_t H, int32_t W, int32_t C, int32_t mode, void* stream) {
  PHT_CHECK_ARG(buf && B > 0 && H > 0 && W > 0, "border_fill: bad args");
  PHT_CHECK_ARG(mode == PHT_PAD_REPLICATE || (mode == PHT_PAD_REFLECT && H > 1 && W > 1), "border_fill: bad mode");
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == PHT_F32) {
    PHT_CHECK_ARG(C % 4 == 0, "border_fill: C %% 4 != 0");
    long long items = (long long)B * (2 * (W + 2) + 2 * H) * (C / 4);
    border_fill_kernel<float><<<grid_for(items, 256), 256, 0, st>>>((float*)buf, B, H, W, C, mode);
  } else {
    PHT_CHECK_ARG(C % 8 == 0, "border_fill: C %% 8 != 0");
    long long items = (long long)B * (2 * (W + 2) + 2 * H) * (C / 8);
    border_fill_kernel<bf16><<<grid_for(items, 256), 256, 0, st>>>((bf16*)buf, B, H, W, C, mode);
  }
  count_launch(CNT_OTHER);
  PHT_LAUNCH_CHECK();
  return PHT_OK;
}

int pht_pad_fold(const void* gpad, int32_t dtype, int32_t B, int32_t H, int32_t W, int32_t C, int32_t mode,
                 const pht_view* resid, const pht_view* mask, const float* mslope, const pht_view* out1,
                 const pht_view* out2, void* stream) {
  PHT_CHECK_ARG(gpad && B > 0 && H > 0 && W > 0, "pad_fold: bad args");
  PHT_CHECK_ARG((out1 && out1->ptr) || (out2 && out2->ptr), "pad_fold: no output");
  cudaStream_t st = (cudaStream_t)stream;
  View r = make_view(resid), m = make_view(mask), o1 = make_view(out1), o2 = make_view(out2);
  if (dtype == PHT_F32) {
    PHT_CHECK_ARG(C % 4 == 0, "pad_fold: C %% 4 != 0");
    long long items = (long long)B * H * W * (C / 4);
    pad_fold_kernel<float><<<grid_for(items, 256), 256, 0, st>>>((const float*)gpad, B, H, W, C, mode, r, m, mslope, o1, o2);
  } else {
    PHT_CHECK_ARG(C % 8 == 0, "pad_fold: C %% 8 != 0");
    long long items = (long long)B * H * W * (C / 8);
    pad_fold_kernel<bf16><<<grid_for(items, 256), 256, 0, st>>>((const bf16*)gpad, B, H, W, C, mode, r, m, mslope, o1, o2);
  }
  count_launch(CNT_OTHER);
  PHT_LAUNCH_CHECK();
  return PHT_OK;
}

int pht_im2col5(const float* x, void* col, int32_t dtype, int32_t B, int32_t Cin, int32_t H, int32_t W, int32_t Kpad,
                int32_t mode, void* stream) {
  PHT_CHECK_ARG(x && col && Kpad >= 25 * Cin, "im2col5: bad args");
  PHT_CHECK_ARG(mode == PHT_PAD_REPLICATE || (H > 2 && W > 2), "im2col5: reflect needs H,W > 2");
  cudaStream_t st = (cudaStream_t)stream;
  long long items = (long long)B * H * W * Kpad;
  if (Kpad % 8 == 0 && ((uintptr_t)col & 15) == 0 && Cin <= 16) {
    const int tiles = B * ((H + I2C_TH - 1) / I2C_TH) * ((W + I2C_TW - 1) / I2C_TW);
    const size_t smem = (size_t)(I2C_TH + 4) * (I2C_TW + 4) * Cin * sizeof(float) + (size_t)Kpad * sizeof(int);
    if (dtype == PHT_F32) im2col5_tile_kernel<float><<<tiles, 256, smem, st>>>(x, (float*)col, B, Cin, H, W, Kpad, mode);
    else im2col5_tile_kernel<bf16><<<tiles, 256, smem, st>>>(x, (bf16*)col, B, Cin, H, W, Kpad, mode);
    count_launch(CNT_OTHER);
    PHT_LAUNCH_CHECK();
    return PHT_OK;
  }
  if (Kpad % 8 == 0 && ((uintptr_t)col & 15) == 0) {
    if (dtype == PHT_F32) im2col5_vec8_kernel<float><<<grid_for(items / 8, 256), 256, 0, st>>>(x, (float*)col, B, Cin, H, W, Kpad, mode);
    else im2col5_vec8_kernel<bf16><<<grid_for(items / 8, 256), 256, 0, st>>>(x, (bf16*)col, B, Cin, H, W, Kpad, mode);
    count_launch(CNT_OTHER);
    PHT_LAUNCH_CHECK();
    return PHT_OK;
  }
  if (dtype == PHT_F32) im2col5_kernel<float><<<grid_for(items, 256), 256, 0, st>>>(x, (float*)col, B, Cin, H, W, Kpad, mode);
  else im2col5_kernel<bf16><<<grid_for(items, 256), 256, 0, st>>>(x, (bf16*)col, B, Cin, H, W, Kpad, mode);
  count_launch(CNT_OTHER);
  PHT_LAUNCH_CHECK();
  return PHT_OK;
}

int pht_dec_tail_fwd(const pht_view* h, const float* w, const float* bias, const float* x, float* out, int32_t B,
                     int32_t H, int32_t W, void* stream) {
  PHT_CHECK_ARG(h && h->ptr && w && bias && x && out, "dec_tail_fwd: null arg");
  View hv = make_view(h);
  size_t smem = (size_t)27 * hv.C * sizeof(float);
  cudaStream_t st = (cudaStream_t)stream;
  int grid = grid_for((long long)B * H * W * 32, 256);
  if (hv.dtype == PHT_F32) {
    PHT_CHECK_ARG(hv.C % 4 == 0, "dec_tail: C %% 4");
    PHT_CUDA(cudaFuncSetAttribute(dec_tail_fwd_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dec_tail_fwd_kernel<float><<<grid, 256, smem, st>>>(hv, w, bias, x, out, B, H, W);
  } else {
    PHT_CHECK_ARG(hv.C % 8 == 0, "dec_tail: C %% 8");
    PHT_CUDA(cudaFuncSetAttribute(dec_tail_fwd_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dec_tail_fwd_kernel<bf16><<<grid, 256, smem, st>>>(hv, w, bias, x, out, B, H, W);
  }
  count_launch(CNT_OTHER);
  PHT_LAUNCH_CHECK();
  return PHT_OK;
}

int pht_dec_tail_bwd_data(const float* dout, const float* w, const pht_view* h, const pht_view* dh, int32_t B,
                          int32_t H, int32_t W, void* stream) {
  PHT_CHECK_ARG(dout && w && h && h->ptr && dh && dh->ptr, "dec_tail_bwd_data: null arg");
  View hv = make_view(h), dv = make_view(dh);
  PHT_CHECK_ARG(hv.dtype == dv.dtype && hv.C == dv.C, "dec_tail_bwd_data: view mismatch");
  size_t smem = (size_t)27 * hv.C * sizeof(float);
  cudaStream_t st = (cudaStream_t)stream;
  int grid = grid_for((long long)B * H * W * 32, 256);
  if (hv.dtype == PHT_F32) {
    PHT_CUDA(cudaFuncSetAttribute(dec_tail_bwd_data_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dec_tail_bwd_data_kernel<float><<<grid, 256, smem, st>>>(dout, w, hv, dv, B, H, W);
  } else {
    PHT_CUDA(cudaFuncSetAttribute(dec_tail_bwd_data_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dec_tail_bwd_data_kernel<bf16><<<grid, 256, smem, st>>>(dout, w, hv, dv, B, H, W);
  }
  count_launch(CNT_OTHER);
  PHT_LAUNCH_CHECK();
  return PHT_OK;
}

static int dec_tail_blocks(long long npx, int* chunk) {
  int nb = num_sms() * 4;
  if (npx < nb * 32) nb = (int)((npx + 31) / 32);
  if (nb < 1) nb = 1;
  *chunk = (int)((npx + nb - 1) / nb);
  return (int)((npx + *chunk - 1) / *chunk);
}
size_t pht_dec_tail_ws_bytes(int32_t B, int32_t H, int32_t W, int32_t C) {
  int chunk;
  int nb = dec_tail_blocks((long long)B * H * W, &chunk);
  return (size_t)nb * (27 * (size_t)C + 3) * sizeof(float);
}
int pht_dec_tail_bwd_weight(const float* dout, const pht_view* h, float* dw, float* dbias, void* workspace,
                            size_t workspace_bytes, int32_t B, int32_t H, int32_t W, void* stream) {
  PHT_CHECK_ARG(dout && h && h->ptr && dw && dbias && workspace, "dec_tail_bwd_weight: null arg");
  View hv = make_view(h);
  int chunk;
  int nb = dec_tail_blocks((long long)B * H * W, &chunk);
  PHT_CHECK_ARG(workspace_bytes >= pht_dec_tail_ws_bytes(B, H, W, hv.C), "dec_tail_bwd_weight: workspace too small");
  float* part = (float*)workspace;
  float* bpart = part + (size_t)nb * 27 * hv.C;
  cudaStream_t st = (cudaStream_t)stream;
  if (hv.dtype == PHT_F32) dec_tail_bwd_weight_partial<float><<<nb, 256, 0, st>>>(dout, hv, part, bpart, B, H, W, chunk);
  else dec_tail_bwd_weight_partial<bf16><<<nb, 256, 0, st>>>(dout, hv, part, bpart, B, H, W, chunk);
  PHT_LAUNCH_CHECK();
  int n = 27 * hv.C;
  reduce_partials_kernel<<<ceil_div(n, 256), 256, 0, st>>>(part, dw, nb, n);
  reduce_partials_kernel<<<1, 32, 0, st>>>(bpart, dbias, nb, 3);
  count_launch(CNT_OTHER, 3);
  PHT_LAUNCH_CHECK();
  return PHT_OK;
}

int pht_l1_loss(const float* a, const float* b, int64_t n, float grad_scale, float* loss, float* grad, void* stream) {
  PHT_CHECK_ARG(a && b && loss && n > 0, "l1_loss: bad args");
  PHT_CHECK_ARG((((uintptr_t)a | (uintptr_t)b | (uintptr_t)grad) & 15) == 0, "l1_loss: pointers must be 16B aligned");
  int grid = grid_for((n + 3) / 4, 256);
  const int one_wave = num_sms() * 8;   // 8 resident 256-thread blocks per SM: exactly one wave, grid-stride inside
  if (grid > one_wave) grid = one_wave;
  if (grid > L1_MAX_BLOCKS) grid = L1_MAX_BLOCKS;
  float* scratch = (float*)stream_scratch((cudaStream_t)stream, 0, (L1_MAX_BLOCKS + 4) * sizeof(float));
  if (!scratch) return PHT_ERR_CUDA;
  l1_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(a, b, (long long)n, grad_scale, loss, grad, scratch);
  count_launch(CNT_OTHER);
  PHT_LAUNCH_CHECK();
  return PHT_OK;
}

int pht_preprocess(const float* noisy, const float* gt, const float* aux, float* noisy_o, float* gt_o, float* aux_o,
                   int32_t B, int32_t H, int32_t W, void* stream) {
  PHT_CHECK_ARG(noisy && aux && noisy_o && aux_o && B > 0 && H > 0 && W > 0, "preprocess: bad args");
  cudaStream_t st = (cudaStream_t)stream;
  PHT_CHECK_ARG(!gt || gt_o, "preprocess: gt without gt output");
  const bool al16 = ((((uintptr_t)noisy | (uintptr_t)(gt ? gt : noisy) | (uintptr_t)aux | (uintptr_t)noisy_o | (uintptr_t)(gt_o ? gt_o : noisy_o) |
                       (uintptr_t)aux_o) & 15) == 0);
  if (W % 4 == 0 && ((long long)H * W) % 4 == 0 && al16)
    crop_preprocess_px4_kernel<<<grid_for((long long)B * H * W / 4, 256), 256, 0, st>>>(noisy, gt, aux, noisy_o, gt_o, aux_o, 0, 0,
                                                                                        nullptr, nullptr, B, H, W);
  else
    crop_preprocess_px_kernel<<<grid_for((long long)B * H * W, 256), 256, 0, st>>>(noisy, gt, aux, noisy_o, gt_o, aux_o, 0, 0,
                                                                                   nullptr, nullptr, B, H, W);
  count_launch(CNT_OTHER);
  PHT_LAUNCH_CHECK();
  return PHT_OK;
}

int pht_crop_preprocess(const float* noisy_f, const float* gt_f, const float* aux_f, int32_t Hf, int32_t Wf,
                        const int32_t* centres, const int32_t* img_idx, int32_t n, int32_t P, float* noisy_o,
                        float* gt_o, float* aux_o, void* stream) {
  PHT_CHECK_ARG(noisy_f && aux_f && centres && noisy_o && aux_o && n > 0 && P > 0, "crop_preprocess: bad args");
  cudaStream_t st = (cudaStream_t)stream;
  PHT_CHECK_ARG(!gt_f || gt_o, "crop_preprocess: gt without gt output");
  const bool al16 = ((((uintptr_t)noisy_f | (uintptr_t)(gt_f ? gt_f : noisy_f) | (uintptr_t)aux_f | (uintptr_t)noisy_o |
                       (uintptr_t)(gt_o ? gt_o : noisy_o) | (uintptr_t)aux_o) & 15) == 0);
  if (P % 4 == 0 && al16)
    crop_preprocess_px4_kernel<<<grid_for((long long)n * P * P / 4, 256), 256, 0, st>>>(noisy_f, gt_f, aux_f, noisy_o, gt_o, aux_o,
                                                                                        Hf, Wf, centres, img_idx, n, P, P);
  else
    crop_preprocess_px_kernel<<<grid_for((long long)n * P * P, 256), 256, 0, st>>>(noisy_f, gt_f, aux_f, noisy_o, gt_o, aux_o, Hf,
                                                                                   Wf, centres, img_idx, n, P, P);
  count_launch(CNT_OTHER);
  PHT_LAUNCH_CHECK();
  return PHT_OK;
}

int pht_adam(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2, float eps,
             int32_t step, float grad_scale, void* stream) {
  PHT_CHECK_ARG(p && g && m && v && n > 0 && step >= 1, "adam: bad args");
  PHT_CHECK_ARG((((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) & 15) == 0, "adam: pointers must be 16B aligned");
  double bc1 = 1.0 - pow((double)beta1, (double)step);
  double bc2 = 1.0 - pow((double)beta2, (double)step);
  float step_size = (float)((double)lr / bc1);
  float inv_sqrt_bc2 = (float)(1.0 / sqrt(bc2));
  adam_kernel<<<grid_for((n + 3) / 4, 256), 256, 0, (cudaStream_t)stream>>>(p, g, m, v, (long long)n, step_size, beta1,
                                                                          beta2, eps, inv_sqrt_bc2, grad_scale);
  count_launch(CNT_OTHER);
  PHT_LAUNCH_CHECK();
  return PHT_OK;
}

int pht_adam_dev(float* p, const float* g, float* m, float* v, int64_t n, float beta1, float beta2, float eps, float grad_scale,
                 float* hyper, void* stream) {
  PHT_CHECK_ARG(p && g && m && v && hyper && n > 0, "adam_dev: bad args");
  PHT_CHECK_ARG((((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v | (uintptr_t)hyper) & 15) == 0,
                "adam_dev: pointers must be 16B aligned");
  adam_hyper_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(hyper, beta1, beta2);
  adam_dev_kernel<<<grid_for((n + 3) / 4, 256), 256, 0, (cudaStream_t)stream>>>(p, g, m, v, (long long)n, hyper, beta1, beta2, eps,
                                                                              grad_scale);
  count_launch(CNT_OTHER, 2);
  PHT_LAUNCH_CHECK();
  return PHT_OK;
}

static int check_pack(const pht_pack_args* a, PackP* p) {
  PHT_CHECK_ARG(a && a->w && a->packed, "pack: null arg");
  PHT_CHECK_ARG(a->ksize == 1 || a->ksize == 3 || a->ksize == 5, "pack: ksize");
  p->O = a->O; p->I = a->I; p->ks = a->ksize; p->Ntot = a->Ntot; p->Ktot = a->Ktot; p->n_off = a->n_off;
  p->k_off = a->k_off; p->transpose = a->transpose; p->grid = a->grid; p->scale = a->scale;
  p->i_begin = a->i_begin; p->i_count = a->i_count > 0 ? a->i_count : a->I - a->i_begin;
  PHT_CHECK_ARG(p->i_begin >= 0 && p->i_count > 0 && p->i_begin + p->i_count <= a->I, "pack: bad input-channel slice");
  const int Ic = p->i_count;
  if (a->grid > 0) {
    PHT_CHECK_ARG(a->grid >= a->ksize && !a->transpose, "pack: bad embed");
    PHT_CHECK_ARG(a->k_off + a->grid * a->grid * Ic <= a->Ktot && a->n_off + a->O <= a->Ntot, "pack: embed out of range");
  } else if (a->transpose == 2) {
    PHT_CHECK_ARG(a->n_off + Ic <= a->Ntot && a->k_off + a->ksize * a->ksize * a->O <= a->Ktot, "pack: out of range (T2)");
  } else if (!a->transpose) {
    PHT_CHECK_ARG(a->n_off + a->O <= a->Ntot && a->k_off + Ic <= a->Ktot, "pack: out of range");
  } else {
    PHT_CHECK_ARG(a->n_off + Ic <= a->Ntot && a->k_off + a->O <= a->Ktot, "pack: out of range (T)");
  }
  return PHT_OK;
}
int pht_pack_weight(const pht_pack_args* a, void* stream) {
  PackP p;
  int rc = check_pack(a, &p);
  if (rc) return rc;
  long long total = (long long)p.O * p.I * p.ks * p.ks;
  if (a->dtype == PHT_F32) pack_weight_kernel<float><<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(a->w, (float*)a->packed, p);
  else pack_weight_kernel<bf16><<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(a->w, (bf16*)a->packed, p);
  count_launch(CNT_OTHER);
  PHT_LAUNCH_CHECK();
  return PHT_OK;
}
size_t pht_pack_table_bytes(int32_t n) { return (size_t)(n > 0 ? n : 0) * sizeof(PackJob); }

int pht_pack_weights_batched(const pht_pack_args* jobs, int32_t n, void* table_dev, size_t table_bytes, int32_t upload,
                             void* stream) {
  PHT_CHECK_ARG(jobs && n > 0 && table_dev && table_bytes >= pht_pack_table_bytes(n), "pack_batched: bad args");
  cudaStream_t st = (cudaStream_t)stream;
  long long max_total = 0, tiles = 0;
  int max_pitch = 1;
  static thread_local std::vector<PackJob> host;
  host.resize(n);
  for (int i = 0; i < n; ++i) {
    PackP p;
    int rc = check_pack(&jobs[i], &p);
    if (rc) return rc;
    PHT_CHECK_ARG(jobs[i].dtype == PHT_F32 || jobs[i].dtype == PHT_BF16, "pack_batched: bad dtype");
    host[i].w = jobs[i].w; host[i].dst = jobs[i].packed; host[i].p = p; host[i].dtype = jobs[i].dtype;
    host[i].pad_ = (int)tiles;                       // first tile of this job (tiled kernel)
    tiles += (long long)((p.O + PK_TILE - 1) / PK_TILE) * ((p.i_count + PK_TILE - 1) / PK_TILE);
    const int pitch = (p.i_count < PK_TILE ? p.i_count : PK_TILE) * p.ks * p.ks + 1;
    if (pitch > max_pitch) max_pitch = pitch;
    long long total = (long long)p.O * p.i_count;   // one thread per (o, i) pair
    if (total > max_total) max_total = total;
  }
  if (upload) PHT_CUDA(cudaMemcpyAsync(table_dev, host.data(), (size_t)n * sizeof(PackJob), cudaMemcpyHostToDevice, st));
  const size_t smem = (size_t)PK_TILE * max_pitch * sizeof(float);
  if (smem <= 48 * 1024 && tiles < (1ll << 30)) {
    pack_weights_tiled_kernel<<<(unsigned)tiles, 256, smem, st>>>((const PackJob*)table_dev, n);
    count_launch(CNT_OTHER);
    PHT_LAUNCH_CHECK();
    return PHT_OK;
  }
  int gx = (int)((max_total + 255) / 256);
  if (gx > 592) gx = 592;
  dim3 grid(gx, n);
  pack_weights_batched_kernel<<<grid, 256, 0, st>>>((const PackJob*)table_dev);
  count_launch(CNT_OTHER);
  PHT_LAUNCH_CHECK();
  return PHT_OK;
}

int pht_unpack_wgrads_batched(const pht_pack_args* jobs, int32_t n, void* table_dev, size_t table_bytes, int32_t upload,
                              void* stream) {
  PHT_CHECK_ARG(jobs && n > 0 && table_dev && table_bytes >= pht_pack_table_bytes(n), "unpack_batched: bad args");
  cudaStream_t st = (cudaStream_t)stream;
  long long max_total = 0, tiles = 0;
  int max_pitch = 1;
  static thread_local std::vector<PackJob> host;
  host.resize(n);
  for (int i = 0; i < n; ++i) {
    PackP p;
    int rc = check_pack(&jobs[i], &p);
    if (rc) return rc;
    host[i].w = jobs[i].w; host[i].dst = jobs[i].packed; host[i].p = p; host[i].dtype = PHT_F32;
    host[i].pad_ = (int)tiles;                       // first tile of this job (tiled kernel)
    tiles += (long long)((p.O + PK_TILE - 1) / PK_TILE) * ((p.i_count + PK_TILE - 1) / PK_TILE);
    const int pitch = (p.i_count < PK_TILE ? p.i_count : PK_TILE) * p.ks * p.ks + 1;
    if (pitch > max_pitch) max_pitch = pitch;
    long long total = (long long)p.O * p.I * p.ks * p.ks;
    if (total > max_total) max_total = total;
  }
  if (upload) PHT_CUDA(cudaMemcpyAsync(table_dev, host.data(), (size_t)n * sizeof(PackJob), cudaMemcpyHostToDevice, st));
  (void)tiles; (void)max_pitch;   // (a tiled unpack like the pack kernel measured 3x slower: 29 vs 10 us per launch)
  int gx = (int)((max_total + 255) / 256);
  if (gx > 592) gx = 592;
  dim3 grid(gx, n);
  unpack_wgrads_batched_kernel<<<grid, 256, 0, st>>>((const PackJob*)table_dev);
  count_launch(CNT_OTHER);
  PHT_LAUNCH_CHECK();
  return PHT_OK;
}

// Decoder tail as a 1x1 GEMM + gather: y[p][t * 3 + co] = sum_c in(p)[c] w[co][c][t] (ONE pass over the 256-channel input
// instead of nine shifted ones), then out(p)[co] = bias[co] + x(p)[co] + sum_t y[p + tap_t][t * 3 + co] (zero padding).
// One thread per pixel: 27 floats from the 9 neighbouring rows of y (L2-resident), three coalesced NCHW stores.
__global__ void tail_gather_kernel(const float* __restrict__ y, int ldy, const float* __restrict__ bias, const float* __restrict__ x,
                                   float* __restrict__ out, int B, int H, int W) {
  const long long total = (long long)B * H * W;
  const float b0 = bias[0], b1 = bias[1], b2 = bias[2];
  for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < total; p += (long long)gridDim.x * blockDim.x) {
    const int px = (int)(p % W);
    const long long r = p / W;
    const int py = (int)(r % H), b = (int)(r / H);
    float a0 = b0, a1 = b1, a2 = b2;
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const int yy = py + t / 3 - 1, xx = px + t % 3 - 1;
      if ((unsigned)yy < (unsigned)H && (unsigned)xx < (unsigned)W) {
        const float* q = y + (((long long)b * H + yy) * W + xx) * ldy + t * 3;
        a0 += q[0]; a1 += q[1]; a2 += q[2];
      }
    }
    const long long o = ((long long)b * 3 * H + py) * W + px, hw = (long long)H * W;
    out[o] = a0 + x[o];
    out[o + hw] = a1 + x[o + hw];
    out[o + 2 * hw] = a2 + x[o + 2 * hw];
  }
}

int pht_tail_gather(const float* y, int32_t ldy, const float* bias, const float* x_nchw, float* out_nchw, int32_t B, int32_t H,
                    int32_t W, void* stream) {
  PHT_CHECK_ARG(y && bias && x_nchw && out_nchw && ldy >= 27, "tail_gather: bad args");
  long long n = (long long)B * H * W;
  tail_gather_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(y, ldy, bias, x_nchw, out_nchw, B, H, W);
  count_launch(CNT_OTHER);
  PHT_LAUNCH_CHECK();
  return PHT_OK;
}

int pht_tail_finish(const float* y, int32_t ldy, const float* bias, const float* x_nchw, float* out_nchw, int32_t B, int32_t H,
                    int32_t W, void* stream) {
  PHT_CHECK_ARG(y && bias && x_nchw && out_nchw && ldy >= 3, "tail_finish: bad args");
  long long n = (long long)B * 3 * H * W;
  tail_finish_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(y, ldy, bias, x_nchw, out_nchw, B, H, W);
  count_launch(CNT_OTHER);
  PHT_LAUNCH_CHECK();
  return PHT_OK;
}

int pht_tail_im2col_bwd(const float* dout_nchw, void* a_bf16, float* dbias, int32_t B, int32_t H, int32_t W, void* stream) {
  PHT_CHECK_ARG(dout_nchw && a_bf16 && dbias, "tail_im2col_bwd: bad args");
  cudaStream_t st = (cudaStream_t)stream;
  long long n = (long long)B * H * W * 8;
  tail_im2col_bwd_kernel<<<grid_for(n, 256), 256, 0, st>>>(dout_nchw, (bf16*)a_bf16, B, H, W);
  float* scratch = (float*)stream_scratch(st, 1, (3 * TAIL_DB_BLOCKS + 4) * sizeof(float));
  if (!scratch) return PHT_ERR_CUDA;
  tail_dbias_kernel<<<dim3(TAIL_DB_BLOCKS, 3), 256, 0, st>>>(dout_nchw, dbias, B, H * W, scratch);
  count_launch(CNT_OTHER, 2);
  PHT_LAUNCH_CHECK();
  return PHT_OK;
}

int pht_unpack_wgrad(const pht_pack_args* a, void* stream) {
  PackP p;
  int rc = check_pack(a, &p);
  if (rc) return rc;
  long long total = (long long)p.O * p.I * p.ks * p.ks;
  unpack_wgrad_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>((float*)a->w, (const float*)a->packed, p);
  count_launch(CNT_OTHER);
  PHT_LAUNCH_CHECK();
  return PHT_OK;
}

int pht_cast2d(const void* src, int32_t sd, int64_t sld, void* dst, int32_t dd, int64_t dld, int64_t rows, int64_t cols,
               void* stream) {
  PHT_CHECK_ARG(src && dst && rows > 0 && cols > 0 && sld >= cols && dld >= cols, "cast: bad args");
  cudaStream_t st = (cudaStream_t)stream;
  int grid = grid_for(rows * cols, 256);
  if (sd == PHT_F32 && dd == PHT_BF16) cast_kernel<float, bf16><<<grid, 256, 0, st>>>((const float*)src, sld, (bf16*)dst, dld, rows, cols);
  else if (sd == PHT_BF16 && dd == PHT_F32) cast_kernel<bf16, float><<<grid, 256, 0, st>>>((const bf16*)src, sld, (float*)dst, dld, rows, cols);
  else if (sd == PHT_F32 && dd == PHT_F32) cast_kernel<float, float><<<grid, 256, 0, st>>>((const float*)src, sld, (float*)dst, dld, rows, cols);
  else if (sd == PHT_BF16 && dd == PHT_BF16) cast_kernel<bf16, bf16><<<grid, 256, 0, st>>>((const bf16*)src, sld, (bf16*)dst, dld, rows, cols);
  else PHT_CHECK_ARG(false, "cast: bad dtype");
  count_launch(CNT_OTHER);
  PHT_LAUNCH_CHECK();
  return PHT_OK;
}
int pht_cast(const void* src, int32_t sd, void* dst, int32_t dd, int64_t n, void* stream) {
  return pht_cast2d(src, sd, n, dst, dd, n, 1, n, stream);
}

}  // extern "C"
