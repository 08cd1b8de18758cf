// tcgen05 / TMEM / TMA implicit-GEMM convolution for sm_100a (bf16 operands, fp32 accumulation).
//
//   acc[p, n] = sum_taps sum_sources sum_k  src_s(p + tap)[k] * w[tap][n][koff_s + k]
//
// One persistent CTA per SM, warp-specialised:
//   warp 0      TMA producer: per (tap, source, 64-channel chunk) one 4-D box load of the activation tile
//               (8 rows x 16 cols x 64 ch of the NHWC view, out-of-image pixels zero-filled by TMA = zero padding,
//               dgrad's implicit border, ragged tiles) and one 3-D box load of the weight slice [BN][64];
//               both land 128B-swizzled, K-major, in a NUM_STAGES-deep mbarrier ring.
//   warp 1      MMA issuer: one elected thread issues 4 x tcgen05.mma (M=128, N=BN, K=16) per stage into one of two
//               TMEM accumulator buffers (2 x BN fp32 columns), tcgen05.commit releases the smem stage / publishes the
//               accumulator.
//   warps 2..5  epilogue: tcgen05.ld the accumulator (each thread owns one output pixel = one TMEM lane), fused
//               bias / residual / (leaky)ReLU / activation-derivative mask, bf16 pack into a 128B-swizzled staging
//               tile [128 px][64 ch] and one TMA tensor store per 64-channel chunk (full-line writes, ragged tiles
//               clipped by the tensor map).  Runs concurrently with the MMAs of the next tile (double-buffered TMEM).
#include "tc_common.cuh"

namespace pht {

using namespace tc;

constexpr int TILE_H = 8, TILE_W = 16, TILE_M = TILE_H * TILE_W;  // 128 output pixels per tile
constexpr int BK = 64;                                            // bf16 elements per 128-byte swizzle row
constexpr int TC_THREADS = 192;
constexpr int MAX_VEC_N = 1024;
constexpr int STG_BYTES = TILE_M * 128;  // epilogue staging tile: 128 pixels x 64 bf16

struct TcGemmP {
  int B, Ho, Wo, N, ks, n_src;
  unsigned flags;
  int srcC[3], srcOy[3], srcOx[3], koff[3];
  int tiles_x, tiles_y, n_tiles, num_tiles;
  const float* bias;
  const float* slope;
  const float* mslope;
  View resid, mask, out1, out2;
  int out1_f32;  // out1 is fp32 (direct stores; used by the 64-wide decoder-tail GEMM only)
};

template <int BN> struct TcCfg {
  static constexpr int A_BYTES = TILE_M * BK * 2;                  // 16 KB
  static constexpr int B_BYTES = BN * BK * 2;                      // 32 KB @ BN=256
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = BN == 256 ? 4 : 6;
  static constexpr int TMEM_COLS = 2 * BN < 32 ? 32 : 2 * BN;      // power of two for BN in {64,128,256}
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + STG_BYTES + 3 * MAX_VEC_N * 4 + 256 + 1024;
};

__device__ __forceinline__ void unpack8(const uint4& u, float* f) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 t = __bfloat1622float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 pack8(const float* f) {
  uint4 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  return u;
}

template <int BN>
__global__ void __launch_bounds__(TC_THREADS, 1)
conv_gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                    const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmW,
                    const __grid_constant__ CUtensorMap tmO1, const __grid_constant__ CUtensorMap tmO2, const TcGemmP P) {
  using Cfg = TcCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B needs 1024-byte aligned stage bases
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* stage_base = smem;
  uint8_t* stg = smem + Cfg::STAGES * Cfg::STAGE_BYTES;   // 1024-aligned (stage sizes are multiples of 1024)
  float* s_bias = reinterpret_cast<float*>(stg + STG_BYTES);
  float* s_slope = s_bias + MAX_VEC_N;
  float* s_mslope = s_slope + MAX_VEC_N;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_mslope + MAX_VEC_N);
  uint64_t* full_bar = bars;                      // [STAGES]
  uint64_t* empty_bar = bars + Cfg::STAGES;       // [STAGES]
  uint64_t* tfull_bar = bars + 2 * Cfg::STAGES;   // [2]
  uint64_t* tempty_bar = tfull_bar + 2;           // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  for (int i = threadIdx.x; i < P.N; i += blockDim.x) {
    s_bias[i] = P.bias ? P.bias[i] : 0.f;
    s_slope[i] = P.slope ? P.slope[i] : 1.f;
    s_mslope[i] = P.mslope ? P.mslope[i] : 0.f;
  }
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA0);
    if (P.n_src > 1) prefetch_tmap(&tmA1);
    if (P.n_src > 2) prefetch_tmap(&tmA2);
    prefetch_tmap(&tmW);
    for (int s = 0; s < Cfg::STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull_bar[a], 1);
      mbar_init(&tempty_bar[a], 128);
    }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int T = P.ks * P.ks, half = P.ks / 2;

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < P.num_tiles; tile += gridDim.x) {
        const int nt = tile % P.n_tiles, mt = tile / P.n_tiles;
        const int tx = mt % P.tiles_x, ty = (mt / P.tiles_x) % P.tiles_y, b = mt / (P.tiles_x * P.tiles_y);
        const int x0 = tx * TILE_W, y0 = ty * TILE_H, n0 = nt * BN;
        for (int t = 0; t < T; ++t) {
          const int dy = t / P.ks - half, dx = t % P.ks - half;
          for (int s = 0; s < P.n_src; ++s) {
            const CUtensorMap* tm = s == 0 ? &tmA0 : (s == 1 ? &tmA1 : &tmA2);
            for (int kc = 0; kc < P.srcC[s]; kc += BK) {
              mbar_wait(&empty_bar[stage], phase ^ 1);
              uint8_t* a_dst = stage_base + stage * Cfg::STAGE_BYTES;
              uint8_t* b_dst = a_dst + Cfg::A_BYTES;
              mbar_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
              tma_load_4d(a_dst, tm, &full_bar[stage], kc, x0 + dx + P.srcOx[s], y0 + dy + P.srcOy[s], b);
              tma_load_3d(b_dst, &tmW, &full_bar[stage], P.koff[s] + kc, n0, t);
              if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(TILE_M, BN, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int kiters = 0;
      for (int s = 0; s < P.n_src; ++s) kiters += P.srcC[s] / BK;
      kiters *= T;
      int it = 0;
      for (int tile = blockIdx.x; tile < P.num_tiles; tile += gridDim.x, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int ki = 0; ki < kiters; ++ki) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(stage_base + stage * Cfg::STAGE_BYTES);
          const uint64_t adesc = umma_desc_k_sw128(a_addr);
          const uint64_t bdesc = umma_desc_k_sw128(a_addr + Cfg::A_BYTES);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // advance 16 bf16 = 32 bytes along K inside the 128B swizzle row: +2 in the (>>4) address field
            umma_bf16(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (ki | k) ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);  // frees this smem stage when the MMAs above have read it
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tfull_bar[acc]);      // accumulator complete -> epilogue
      }
    }
  } else {
    // ================================ epilogue (warps 2..5) ================================
    const int quad = warp & 3;             // TMEM lane quadrant this warp may access
    const int row = quad * 32 + lane;      // accumulator row == pixel index inside the tile
    const int py = row / TILE_W, px = row % TILE_W;
    const bool issuer = (warp == 2 && lane == 0);
    uint8_t* srow = stg + row * 128;
    auto stage_and_store = [&](const float* v, const CUtensorMap* tm, int c_glob, int x0, int y0, int b) {
      // staging tile free? (the previous TMA store has finished READING it)
      if (issuer) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      asm volatile("bar.sync 1, 128;" ::: "memory");
#pragma unroll
      for (int g = 0; g < 8; ++g) *reinterpret_cast<uint4*>(srow + ((g ^ (row & 7)) * 16)) = pack8(v + g * 8);
      fence_proxy_async();
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (issuer) {
        asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                     ::"l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(stg)), "r"(c_glob), "r"(x0), "r"(y0), "r"(b)
                     : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
    };
    int it = 0;
    for (int tile = blockIdx.x; tile < P.num_tiles; tile += gridDim.x, ++it) {
      const int nt = tile % P.n_tiles, mt = tile / P.n_tiles;
      const int tx = mt % P.tiles_x, ty = (mt / P.tiles_x) % P.tiles_y, b = mt / (P.tiles_x * P.tiles_y);
      const int x0 = tx * TILE_W, y0 = ty * TILE_H;
      const int x = x0 + px, y = y0 + py, n0 = nt * BN;
      const bool valid = x < P.Wo && y < P.Ho;
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + acc * BN + ((uint32_t)(quad * 32) << 16);
      const bf16* rp = nullptr;
      const bf16* mp = nullptr;
      if (valid) {
        if (P.flags & (PHT_EPI_RESID_PRE | PHT_EPI_RESID_POST))
          rp = (const bf16*)P.resid.ptr + view_off(P.resid, b, y + P.resid.oy, x + P.resid.ox) + n0;
        if (P.flags & PHT_EPI_MASK) mp = (const bf16*)P.mask.ptr + view_off(P.mask, b, y + P.mask.oy, x + P.mask.ox) + n0;
      }
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 64) {
        float v[64];
        {
          uint32_t r[32];
          tmem_ld32(t_addr + c0, r);       // warp-collective: executed by all lanes, valid or not
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
          tmem_ld32(t_addr + c0 + 32, r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) v[32 + j] = __uint_as_float(r[j]);
        }
        if (c0 + 64 >= BN) {               // accumulator fully in registers: release it to the MMA warp
          tc_fence_before();
          mbar_arrive(&tempty_bar[acc]);
        }
#pragma unroll
        for (int j = 0; j < 64; ++j) v[j] += s_bias[n0 + c0 + j];
        if (rp && (P.flags & PHT_EPI_RESID_PRE)) {
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            float rs[8];
            unpack8(*reinterpret_cast<const uint4*>(rp + c0 + g * 8), rs);
#pragma unroll
            for (int j = 0; j < 8; ++j) v[g * 8 + j] += rs[j];
          }
        }
        if (P.slope) {
#pragma unroll
          for (int j = 0; j < 64; ++j) v[j] = v[j] > 0.f ? v[j] : v[j] * s_slope[n0 + c0 + j];
        }
        if (P.out1.ptr) {
          if (P.out1_f32) {
            if (valid) {
              float* o = (float*)P.out1.ptr + view_off(P.out1, b, y + P.out1.oy, x + P.out1.ox) + n0 + c0;
#pragma unroll
              for (int j = 0; j < 64; j += 4) *reinterpret_cast<float4*>(o + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            }
          } else {
            stage_and_store(v, &tmO1, n0 + c0, x0 + P.out1.ox, y0 + P.out1.oy, b);
          }
        }
        if (P.out2.ptr) {
          if (rp && (P.flags & PHT_EPI_RESID_POST)) {
#pragma unroll
            for (int g = 0; g < 8; ++g) {
              float rs[8];
              unpack8(*reinterpret_cast<const uint4*>(rp + c0 + g * 8), rs);
#pragma unroll
              for (int j = 0; j < 8; ++j) v[g * 8 + j] += rs[j];
            }
          }
          if (mp) {
#pragma unroll
            for (int g = 0; g < 8; ++g) {
              float mk[8];
              unpack8(*reinterpret_cast<const uint4*>(mp + c0 + g * 8), mk);
#pragma unroll
              for (int j = 0; j < 8; ++j) v[g * 8 + j] *= (mk[j] > 0.f ? 1.f : s_mslope[n0 + c0 + g * 8 + j]);
            }
          }
          stage_and_store(v, &tmO2, n0 + c0, x0 + P.out2.ox, y0 + P.out2.oy, b);
        }
      }
    }
    if (issuer) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // all tensor stores complete before exit
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------ host
namespace tc {

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

int make_tmap_bf16(CUtensorMap* out, void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled is not available from the driver");
    return PHT_ERR_CUDA;
  }
  cuuint64_t d[5], st[5];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) {
    d[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
    if (i > 0) st[i - 1] = strides_bytes[i - 1];
  }
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, base, d, st, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (rank %d dims %llu %llu %llu strides %llu %llu)", (int)r, rank,
              (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)(rank > 2 ? dims[2] : 0),
              (unsigned long long)strides_bytes[0], (unsigned long long)(rank > 2 ? strides_bytes[1] : 0));
    return PHT_ERR_CUDA;
  }
  return PHT_OK;
}

}  // namespace tc

static bool view_tma_ok(const pht_view& v) {
  if (!v.ptr || v.dtype != PHT_BF16) return false;
  if (v.C % BK != 0) return false;
  if (((uintptr_t)v.ptr & 15) != 0) return false;
  // TMA global strides: multiples of 16 bytes
  if ((v.sx * 2) % 16 || (v.sy * 2) % 16 || (v.sb * 2) % 16) return false;
  if (v.sx <= 0 || v.sy <= 0 || v.sb <= 0) return false;
  return true;
}
static bool view_vec8_ok(const pht_view& v) {
  return v.ptr && v.dtype == PHT_BF16 && ((uintptr_t)v.ptr & 15) == 0 && v.sx % 8 == 0 && v.sy % 8 == 0 && v.sb % 8 == 0;
}

static int make_src_tmap(CUtensorMap* tm, const pht_view& v, int B) {
  uint64_t dims[4] = {(uint64_t)v.C, (uint64_t)v.W, (uint64_t)v.H, (uint64_t)B};
  uint64_t strides[3] = {(uint64_t)v.sx * 2, (uint64_t)v.sy * 2, (uint64_t)v.sb * 2};
  uint32_t box[4] = {BK, TILE_W, TILE_H, 1};
  return make_tmap_bf16(tm, v.ptr, 4, dims, strides, box);
}

template <int BN>
static int launch_tc(const pht_conv_gemm_args* a, const TcGemmP& P, const CUtensorMap* tmA, const CUtensorMap& tmW,
                     const CUtensorMap* tmO, cudaStream_t st) {
  using Cfg = TcCfg<BN>;
  static bool attr_set = false;
  if (!attr_set) {
    PHT_CUDA(cudaFuncSetAttribute(conv_gemm_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_set = true;
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int grid = P.num_tiles < sms ? P.num_tiles : sms;
  conv_gemm_tc_kernel<BN><<<grid, TC_THREADS, Cfg::SMEM_BYTES, st>>>(tmA[0], tmA[1], tmA[2], tmW, tmO[0], tmO[1], P);
  PHT_LAUNCH_CHECK();
  return PHT_OK;
}

int conv_gemm_tc(const pht_conv_gemm_args* a, cudaStream_t st, bool* handled) {
  *handled = false;
  if (a->dtype != PHT_BF16 || a->N > MAX_VEC_N) return PHT_OK;
  int BN = a->N % 256 == 0 ? 256 : (a->N % 128 == 0 ? 128 : (a->N % 64 == 0 ? 64 : 0));
  if (!BN) return PHT_OK;
  int ktot = 0;
  for (int s = 0; s < a->n_src; ++s) {
    if (!view_tma_ok(a->src[s])) return PHT_OK;
    ktot += a->src[s].C;
  }
  if ((a->flags & (PHT_EPI_RESID_PRE | PHT_EPI_RESID_POST)) && !view_vec8_ok(a->resid)) return PHT_OK;
  if ((a->flags & PHT_EPI_MASK) && !view_vec8_ok(a->mask)) return PHT_OK;
  const bool out1_f32 = a->out1.ptr && a->out1.dtype == PHT_F32;
  if (out1_f32) {
    if (((uintptr_t)a->out1.ptr & 15) || a->out1.sx % 4 || a->out1.sy % 4 || a->out1.sb % 4) return PHT_OK;
  } else if (a->out1.ptr && (!view_tma_ok(a->out1) || a->out1.C < a->N)) {
    return PHT_OK;
  }
  if (a->out2.ptr && (!view_tma_ok(a->out2) || a->out2.C < a->N)) return PHT_OK;
  if (((uintptr_t)a->w & 15) != 0) return PHT_OK;
  if (!get_encode_fn()) return PHT_OK;

  TcGemmP P;
  P.B = a->B; P.Ho = a->Ho; P.Wo = a->Wo; P.N = a->N; P.ks = a->ksize; P.n_src = a->n_src; P.flags = a->flags;
  CUtensorMap tmA[3], tmW;
  int k = 0;
  for (int s = 0; s < 3; ++s) {
    if (s < a->n_src) {
      P.srcC[s] = a->src[s].C; P.srcOy[s] = a->src[s].oy; P.srcOx[s] = a->src[s].ox; P.koff[s] = k;
      k += a->src[s].C;
      int rc = make_src_tmap(&tmA[s], a->src[s], a->B);
      if (rc) return rc;
    } else {
      P.srcC[s] = 0; P.srcOy[s] = P.srcOx[s] = 0; P.koff[s] = 0;
      tmA[s] = tmA[0];
    }
  }
  const int T = a->ksize * a->ksize;
  {
    uint64_t dims[3] = {(uint64_t)ktot, (uint64_t)a->N, (uint64_t)T};
    uint64_t strides[2] = {(uint64_t)ktot * 2, (uint64_t)ktot * a->N * 2};
    uint32_t box[3] = {BK, (uint32_t)BN, 1};
    int rc = make_tmap_bf16(&tmW, const_cast<void*>(a->w), 3, dims, strides, box);
    if (rc) return rc;
  }
  P.tiles_x = ceil_div(a->Wo, TILE_W);
  P.tiles_y = ceil_div(a->Ho, TILE_H);
  P.n_tiles = a->N / BN;
  P.num_tiles = a->B * P.tiles_x * P.tiles_y * P.n_tiles;
  P.bias = a->bias; P.slope = a->slope; P.mslope = a->mslope;
  P.resid = (a->flags & (PHT_EPI_RESID_PRE | PHT_EPI_RESID_POST)) ? make_view(a->resid) : null_view();
  P.mask = (a->flags & PHT_EPI_MASK) ? make_view(a->mask) : null_view();
  P.out1 = a->out1.ptr ? make_view(a->out1) : null_view();
  P.out2 = a->out2.ptr ? make_view(a->out2) : null_view();
  P.out1_f32 = out1_f32 ? 1 : 0;
  CUtensorMap tmO[2];
  tmO[0] = tmO[1] = tmW;
  if (a->out1.ptr && !out1_f32) {
    int rc1 = make_src_tmap(&tmO[0], a->out1, a->B);
    if (rc1) return rc1;
  }
  if (a->out2.ptr) {
    int rc2 = make_src_tmap(&tmO[1], a->out2, a->B);
    if (rc2) return rc2;
  }
  int rc;
  if (BN == 256) rc = launch_tc<256>(a, P, tmA, tmW, tmO, st);
  else if (BN == 128) rc = launch_tc<128>(a, P, tmA, tmW, tmO, st);
  else rc = launch_tc<64>(a, P, tmA, tmW, tmO, st);
  if (rc) return rc;
  count_launch(CNT_GEMM_TC);
  *handled = true;
  return PHT_OK;
}

}  // namespace pht
