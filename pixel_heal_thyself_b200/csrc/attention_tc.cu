// tcgen05 block-local attention for sm_100a (bf16 operands, fp32 softmax / accumulation).
// Fixed geometry of the AFGSA layer: head_dim 64, block 8x8 (64 queries), halo 3 (14x14 = 196 keys).
//
// Forward, per 8x8 query block and per PAIR of heads (two M=64 tcgen05 tiles interleaved in the two 16-lane
// halves of every TMEM sub-partition, so all 128 lanes / softmax threads are busy):
//   TMA      Q box [8x8 px x 64 ch] and K / V boxes [14x14 px x 64 ch] per head; window pixels outside the image
//            are zero-filled by TMA (== the reference's zero-padded, un-masked keys).
//   S = Q K'^T   one tcgen05.mma chain M=64, N=240, K=64 per head.  The key tile has 240 rows: 196 keys, 12 zero
//            rows, then 32 "relative position rows" [rel_h[r] | 0] and [0 | rel_w[c]], so columns 208..239 of S
//            hold q_h.rel_h[r] and q_w.rel_w[c]:  S'[q, (r,c)] = S[q, key] + S[q, 208+r] + S[q, 224+c]
//            == q.(k + rel) of the reference, without ever materialising K + rel.
//   softmax  128 threads, one (head, query) row each, two passes over TMEM (max, then exp2 / sum), P -> bf16 into a
//            128B-swizzled K-major smem tile.
//   O = P V  tcgen05.mma M=64, N=64, K=208 with V consumed MN-major straight from its TMA box; O overwrites
//            the first 64 columns of the (already consumed) S region.
//   epilogue O / sum + residual -> bf16 NHWC, log-sum-exp saved for the backward.
// Two TMEM regions (one per head pair) ping-pong so S(it+1) is computed while softmax(it) runs.
#include "tc_common.cuh"

namespace pht {

using namespace tc;

int attn_bwd_tc(const pht_attn_bwd_args* a, cudaStream_t st, bool* handled) {  // backward: CUDA-core kernel for now
  *handled = false;
  return PHT_OK;
}

constexpr int AT_THREADS = 192;
constexpr int AT_NK = 196, AT_NKP = 208, AT_NS = 240;          // keys, keys padded to 16, S columns incl. rel rows
constexpr int AT_Q_BYTES = 64 * 128;                           // 8 KB per head
constexpr int AT_K_BYTES = AT_NS * 128;                        // 30720
constexpr int AT_V_BYTES = AT_NKP * 128;                       // 26624
constexpr int AT_P_BYTES = 4 * 64 * 128;                       // 4 K-tiles of 64 keys
constexpr int AT_KV_BOX_BYTES = AT_NK * 128;                   // 25088 written by one TMA box
constexpr int AT_SMEM = 2 * (AT_Q_BYTES + AT_K_BYTES + AT_V_BYTES + AT_P_BYTES) + 256 + 1024;

struct AtP {
  int B, H, W, nbx, nby, nblocks;
  View resid, out;
  const float* rel_h;
  const float* rel_w;
  float* lse;
};

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(AT_THREADS, 1)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                   const __grid_constant__ CUtensorMap tmV, const AtP P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* Qs = smem;                                   // [2][64 x 128B]
  uint8_t* Ks = Qs + 2 * AT_Q_BYTES;                    // [2][240 x 128B]
  uint8_t* Vs = Ks + 2 * AT_K_BYTES;                    // [2][208 x 128B]
  uint8_t* Ps = Vs + 2 * AT_V_BYTES;                    // [2][4][64 x 128B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(Ps + 2 * AT_P_BYTES);
  uint64_t* qk_full = bars + 0;
  uint64_t* qk_empty = bars + 1;
  uint64_t* v_full = bars + 2;
  uint64_t* pv_done = bars + 3;
  uint64_t* p_full = bars + 4;
  uint64_t* s_full = bars + 5;     // [2]
  uint64_t* tmem_free = bars + 7;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // one-time smem constants: zero pad rows of K / V, relative-position rows of K (swizzled like a TMA box)
  for (int i = threadIdx.x; i < 2 * (AT_NS - AT_NK) * 8; i += blockDim.x) {  // K rows 196..239, 16B chunks
    const int h = i / ((AT_NS - AT_NK) * 8), rem = i % ((AT_NS - AT_NK) * 8);
    const int R = AT_NK + rem / 8, ch = rem % 8;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = 0.f;
    if (R >= AT_NKP) {
      const int rr = R - AT_NKP;                 // 0..31
      if (rr < 14 && ch < 4) {                   // [rel_h[r] | 0]
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = P.rel_h[rr * 32 + ch * 8 + j];
      } else if (rr >= 16 && rr < 30 && ch >= 4) {  // [0 | rel_w[c]]
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = P.rel_w[(rr - 16) * 32 + (ch - 4) * 8 + j];
      }
    }
    uint4 u;
    __nv_bfloat162* hh = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
    for (int j = 0; j < 4; ++j) hh[j] = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
    *reinterpret_cast<uint4*>(Ks + h * AT_K_BYTES + R * 128 + ((ch ^ (R & 7)) * 16)) = u;
  }
  for (int i = threadIdx.x; i < 2 * (AT_NKP - AT_NK) * 8; i += blockDim.x) {  // V rows 196..207
    const int h = i / ((AT_NKP - AT_NK) * 8), rem = i % ((AT_NKP - AT_NK) * 8);
    *reinterpret_cast<uint4*>(Vs + h * AT_V_BYTES + (AT_NK + rem / 8) * 128 + (rem % 8) * 16) = make_uint4(0, 0, 0, 0);
  }
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmQ);
    prefetch_tmap(&tmK);
    prefetch_tmap(&tmV);
    mbar_init(qk_full, 1);
    mbar_init(qk_empty, 1);
    mbar_init(v_full, 1);
    mbar_init(pv_done, 1);
    mbar_init(p_full, 128);
    for (int r = 0; r < 2; ++r) {
      mbar_init(&s_full[r], 1);
      mbar_init(&tmem_free[r], 128);
    }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int my_blocks = (P.nblocks - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int n_it = 2 * my_blocks;  // (block, head pair) iterations

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (lane == 0) {
      for (int it = 0; it < n_it; ++it) {
        const int blk = blockIdx.x + (it >> 1) * gridDim.x, pair = it & 1;
        const int bx = blk % P.nbx, by = (blk / P.nbx) % P.nby, b = blk / (P.nbx * P.nby);
        const uint32_t ph = it & 1;
        mbar_wait(qk_empty, ph ^ 1);
        mbar_expect_tx(qk_full, 2 * (AT_Q_BYTES + AT_KV_BOX_BYTES));
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int c0 = (pair * 2 + h) * 64;
          tma_load_4d(Qs + h * AT_Q_BYTES, &tmQ, qk_full, c0, bx * 8, by * 8, b);
          tma_load_4d(Ks + h * AT_K_BYTES, &tmK, qk_full, c0, bx * 8 - 3, by * 8 - 3, b);
        }
        mbar_wait(pv_done, ph ^ 1);  // V (and P) of the previous iteration consumed
        mbar_expect_tx(v_full, 2 * AT_KV_BOX_BYTES);
#pragma unroll
        for (int h = 0; h < 2; ++h)
          tma_load_4d(Vs + h * AT_V_BYTES, &tmV, v_full, (pair * 2 + h) * 64, bx * 8 - 3, by * 8 - 3, b);
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    if (lane == 0) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(64, AT_NS, 0, 0);
      constexpr uint32_t idesc_pv = umma_idesc_bf16(64, 64, 0, 1);  // B = V is MN-major
      auto issue_s = [&](int it) {
        const int r = it & 1;
        mbar_wait(qk_full, it & 1);
        mbar_wait(&tmem_free[r], ((it >> 1) & 1) ^ 1);
        tc_fence_after();
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const uint64_t qd = umma_desc_k_sw128(smem_u32(Qs + h * AT_Q_BYTES));
          const uint64_t kd = umma_desc_k_sw128(smem_u32(Ks + h * AT_K_BYTES));
          const uint32_t d = tmem_base + r * 256 + ((uint32_t)(h * 16) << 16);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16(d, qd + 2 * k, kd + 2 * k, idesc_s, k ? 1u : 0u);
        }
        umma_commit(qk_empty);
        umma_commit(&s_full[r]);
      };
      if (n_it > 0) issue_s(0);
      for (int it = 0; it < n_it; ++it) {
        if (it + 1 < n_it) issue_s(it + 1);
        const int r = it & 1;
        mbar_wait(p_full, it & 1);
        mbar_wait(v_full, it & 1);
        tc_fence_after();
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const uint32_t d = tmem_base + r * 256 + ((uint32_t)(h * 16) << 16);
          const uint32_t p_addr = smem_u32(Ps + h * AT_P_BYTES);
          const uint32_t v_addr = smem_u32(Vs + h * AT_V_BYTES);
#pragma unroll
          for (int kk = 0; kk < AT_NKP / 16; ++kk) {
            const uint64_t pd = umma_desc_k_sw128(p_addr + (kk >> 2) * 8192) + 2 * (kk & 3);
            const uint64_t vd = umma_desc_mn_sw128(v_addr + kk * 2048, 8192, 1024);
            umma_bf16(d, pd, vd, idesc_pv, kk ? 1u : 0u);
          }
        }
        umma_commit(pv_done);
      }
    }
  } else {
    // ================================ softmax + epilogue (warps 2..5) ================================
    const int quad = warp & 3;
    const int hp = lane >> 4;                    // head inside the pair (TMEM lane half)
    const int q = quad * 16 + (lane & 15);       // query row
    const int qy = q >> 3, qx = q & 7;
    const float LOG2E = 1.4426950408889634f;
    float prev_m = 0.f, prev_sum = 1.f;
    int prev_blk = 0, prev_pair = 0;

    auto epilogue = [&](int it, float m, float sum, int blk, int pair) {
      const int r = it & 1;
      mbar_wait(pv_done, it & 1);
      tc_fence_after();
      const int bx = blk % P.nbx, by = (blk / P.nbx) % P.nby, b = blk / (P.nbx * P.nby);
      const int y = by * 8 + qy, x = bx * 8 + qx, head = pair * 2 + hp;
      const float inv = 1.f / sum;
      const uint32_t t_addr = tmem_base + r * 256 + ((uint32_t)(quad * 32) << 16);
      bf16* op = (bf16*)P.out.ptr + view_off(P.out, b, y, x) + head * 64;
      const bf16* rp = P.resid.ptr ? (const bf16*)P.resid.ptr + view_off(P.resid, b, y, x) + head * 64 : nullptr;
#pragma unroll
      for (int c0 = 0; c0 < 64; c0 += 32) {
        uint32_t o[32];
        tmem_ld32(t_addr + c0, o);
        tmem_ld_wait();
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          float v[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(o[g * 8 + j]) * inv;
          if (rp) {
            uint4 ru = *reinterpret_cast<const uint4*>(rp + c0 + g * 8);
            const __nv_bfloat162* rh = reinterpret_cast<const __nv_bfloat162*>(&ru);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              float2 f = __bfloat1622float2(rh[j]);
              v[2 * j] += f.x;
              v[2 * j + 1] += f.y;
            }
          }
          uint4 u;
          __nv_bfloat162* uh = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
          for (int j = 0; j < 4; ++j) uh[j] = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
          *reinterpret_cast<uint4*>(op + c0 + g * 8) = u;
        }
      }
      if (P.lse) P.lse[(((long long)b * P.H + y) * P.W + x) * 4 + head] = m + logf(sum);
      tc_fence_before();
      mbar_arrive(&tmem_free[r]);
    };

    for (int it = 0; it < n_it; ++it) {
      const int blk = blockIdx.x + (it >> 1) * gridDim.x, pair = it & 1;
      const int r = it & 1;
      mbar_wait(&s_full[r], (it >> 1) & 1);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + r * 256 + ((uint32_t)(quad * 32) << 16);
      uint32_t rr[32];
      tmem_ld32(t_addr + AT_NKP, rr);  // [0..13] = q_h.rel_h[r], [16..29] = q_w.rel_w[c]
      tmem_ld_wait();
      // pass 1: row maximum
      float m = -INFINITY;
#pragma unroll
      for (int c0 = 0; c0 < 224; c0 += 32) {
        uint32_t s[32];
        tmem_ld32(t_addr + c0, s);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int key = c0 + j;
          if (key < AT_NK) {
            const float v = __uint_as_float(s[j]) + __uint_as_float(rr[key / 14]) + __uint_as_float(rr[16 + key % 14]);
            m = fmaxf(m, v);
          }
        }
      }
      // the previous pair's O is final by now: write it out and free its TMEM region for S(it+1)
      if (it > 0) epilogue(it - 1, prev_m, prev_sum, prev_blk, prev_pair);
      // pass 2: p = exp(s - m), row sum, bf16 P tile (K-major, 128B swizzle)
      const float m2 = m * LOG2E;
      float sum = 0.f;
      uint8_t* prow = Ps + hp * AT_P_BYTES + q * 128;
#pragma unroll
      for (int c0 = 0; c0 < 224; c0 += 32) {
        uint32_t s[32];
        tmem_ld32(t_addr + c0, s);
        tmem_ld_wait();
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int key0 = c0 + g * 8;
          if (key0 < AT_NKP) {
            float p[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const int key = key0 + j;
              if (key < AT_NK) {
                const float v = __uint_as_float(s[g * 8 + j]) + __uint_as_float(rr[key / 14]) + __uint_as_float(rr[16 + key % 14]);
                p[j] = ex2(fmaf(v, LOG2E, -m2));
                sum += p[j];
              } else {
                p[j] = 0.f;
              }
            }
            uint4 u;
            __nv_bfloat162* uh = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
            for (int j = 0; j < 4; ++j) uh[j] = __floats2bfloat162_rn(p[2 * j], p[2 * j + 1]);
            const int tile = key0 >> 6, ch = (key0 & 63) >> 3;
            *reinterpret_cast<uint4*>(prow + tile * 8192 + ((ch ^ (q & 7)) * 16)) = u;
          }
        }
      }
      fence_proxy_async();   // generic-proxy smem writes -> visible to the tensor core (async proxy)
      tc_fence_before();     // all TMEM reads of S done before the MMA warp may overwrite the region with O
      mbar_arrive(p_full);
      prev_m = m; prev_sum = sum; prev_blk = blk; prev_pair = pair;
    }
    if (n_it > 0) epilogue(n_it - 1, prev_m, prev_sum, prev_blk, prev_pair);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

static bool at_view_ok(const pht_view& v, int C) {
  if (!v.ptr || v.dtype != PHT_BF16 || v.C != C) return false;
  if (((uintptr_t)v.ptr & 15) != 0) return false;
  if ((v.sx * 2) % 16 || (v.sy * 2) % 16 || (v.sb * 2) % 16) return false;
  return v.sx > 0 && v.sy > 0 && v.sb > 0;
}

static int at_tmap(CUtensorMap* tm, const pht_view& v, int B, int box_w, int box_h) {
  uint64_t dims[4] = {(uint64_t)v.C, (uint64_t)v.W, (uint64_t)v.H, (uint64_t)B};
  uint64_t strides[3] = {(uint64_t)v.sx * 2, (uint64_t)v.sy * 2, (uint64_t)v.sb * 2};
  uint32_t box[4] = {64, (uint32_t)box_w, (uint32_t)box_h, 1};
  return make_tmap_bf16(tm, v.ptr, 4, dims, strides, box);
}

int attn_fwd_tc(const pht_attn_args* a, cudaStream_t st, bool* handled) {
  *handled = false;
  if (a->dtype != PHT_BF16 || a->heads != 4 || a->head_dim != 64 || a->block != 8 || a->halo != 3) return PHT_OK;
  if (a->H % 8 || a->W % 8) return PHT_OK;  // the CUDA-core entry reports the reference's assertion
  if (!at_view_ok(a->q, 256) || !at_view_ok(a->k, 256) || !at_view_ok(a->v, 256) || !at_view_ok(a->out, 256)) return PHT_OK;
  if (a->resid.ptr && !at_view_ok(a->resid, 256)) return PHT_OK;
  if (!get_encode_fn()) return PHT_OK;
  CUtensorMap tmQ, tmK, tmV;
  int rc = at_tmap(&tmQ, a->q, a->B, 8, 8);
  if (rc) return rc;
  rc = at_tmap(&tmK, a->k, a->B, 14, 14);
  if (rc) return rc;
  rc = at_tmap(&tmV, a->v, a->B, 14, 14);
  if (rc) return rc;
  AtP P;
  P.B = a->B; P.H = a->H; P.W = a->W; P.nbx = a->W / 8; P.nby = a->H / 8; P.nblocks = a->B * P.nbx * P.nby;
  P.resid = a->resid.ptr ? make_view(a->resid) : null_view();
  P.out = make_view(a->out);
  P.rel_h = a->rel_h; P.rel_w = a->rel_w; P.lse = a->lse;
  static bool attr = false;
  if (!attr) {
    PHT_CUDA(cudaFuncSetAttribute(attn_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM));
    attr = true;
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int grid = P.nblocks < sms ? P.nblocks : sms;
  attn_fwd_tc_kernel<<<grid, AT_THREADS, AT_SMEM, st>>>(tmQ, tmK, tmV, P);
  PHT_LAUNCH_CHECK();
  count_launch(CNT_ATTN_TC);
  *handled = true;
  return PHT_OK;
}

}  // namespace pht
