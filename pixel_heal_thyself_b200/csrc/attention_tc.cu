// tcgen05 block-local attention for sm_100a (bf16 operands, fp32 softmax / accumulation).
// Fixed geometry of the AFGSA layer: head_dim 64, block 8x8 (64 queries), halo 3 (14x14 = 196 keys).
//
// Forward, per 8x8 query block and per PAIR of heads (two M=64 tcgen05 tiles interleaved in the two 16-lane
// halves of every TMEM sub-partition, so all 128 lanes / softmax threads are busy):
//   TMA      Q box [8x8 px x 64 ch] and K / V boxes [14x14 px x 64 ch] per head; window pixels outside the image
//            are zero-filled by TMA (== the reference's zero-padded, un-masked keys).
//   S = Q K^T + Q REL^T   two accumulating tcgen05.mma chains M=64, N=200, K=64 per head.  REL is a constant smem tile
//            [key (r,c)][rel_h[r] | rel_w[c]], so S == q.(k + rel) of the reference (model.py:496-501) without ever
//            materialising K + rel and without any per-element bias arithmetic in the softmax.
//   softmax  256 threads: two warps per TMEM sub-partition share every (head, query) row (columns [0,104) / [104,200)),
//            two passes over TMEM (max, then exp2 / sum) in 32-column chunks with the next chunk's tcgen05.ld in
//            flight; partial maxima / sums are exchanged through smem; P -> bf16 into a 128B-swizzled K-major tile.
//   O = P V  tcgen05.mma M=64, N=64, K=208 with V consumed MN-major straight from its TMA box, into its own TMEM
//            columns.
//   epilogue O / sum + residual (TMA-loaded tile, updated in place) -> bf16 tile -> TMA store; log-sum-exp saved.
// Two S regions ping-pong: S runs two iterations ahead of P.V on the tensor pipe (S(it+2) is issued right behind P.V(it)).
// Warps: 0 TMA producer, 1 MMA issuer, 2..5 and 7..10 softmax/epilogue, 6 output store + residual prefetch.
#include <atomic>
#include <type_traits>

#include "tc_common.cuh"

namespace pht {

using namespace tc;

void walk_dir_set(const void* p, int dir);   // igemm_tc.cu: tile-walk direction bookkeeping (serpentine order)

constexpr int AF_THREADS = 352;                                // forward kernel (11 warps)
constexpr int AT_NK = 196, AT_NKP = 208, AT_NS = 200;          // keys, keys padded to 16 (P / V), S columns = K-tile rows
constexpr int AT_Q_BYTES = 64 * 128;                           // 8 KB per head
constexpr int AT_K_BYTES = AT_NS * 128;                        // 25600 (25 swizzle atoms)
constexpr int AT_V_BYTES = AT_NKP * 128;                       // 26624
constexpr int AT_P_BYTES = 4 * 64 * 128;                       // 4 K-tiles of 64 keys
constexpr int AT_KV_BOX_BYTES = AT_NK * 128;                   // 25088 written by one TMA box
constexpr int AT_RO_BYTES = 2 * AT_Q_BYTES;                    // residual-in == output staging: 2 heads x [64 px][64 ch]
constexpr int AT_COL_O = 2 * AT_NS;                               // TMEM: S(even) [0,200), S(odd) [200,400), O [400,464)
constexpr int AT_XCH_BYTES = 2 * 2 * 128 * 4;                      // [max | sum][part][128 rows] fp32 partials of the two warp groups
constexpr int AT_SMEM = 2 * (AT_Q_BYTES + AT_K_BYTES + AT_V_BYTES + AT_P_BYTES) + AT_K_BYTES + AT_RO_BYTES + AT_XCH_BYTES + 256 + 1024;
static_assert(AT_SMEM <= 232448, "attn_fwd_tc: shared memory budget");

constexpr int AB_TRACE_ITERS = 48, AB_TRACE_EVENTS = 12;
__device__ long long g_attn_bwd_trace[AB_TRACE_ITERS * AB_TRACE_EVENTS];   // clock64 stamps of CTA 0 (diagnostics)
static int g_attn_trace_on = 0;
void set_attn_trace(int v) { g_attn_trace_on = v; }
int read_attn_trace(long long* host, int n) {
  if (n > AB_TRACE_ITERS * AB_TRACE_EVENTS) n = AB_TRACE_ITERS * AB_TRACE_EVENTS;
  PHT_CUDA(cudaDeviceSynchronize());
  PHT_CUDA(cudaMemcpyFromSymbol(host, g_attn_bwd_trace, (size_t)n * sizeof(long long)));
  return n;
}

struct AtP {
  int B, H, W, nbx, nby, nblocks, trace;
  int has_resid, residOy, residOx, outOy, outOx;
  const float* rel_h;
  const float* rel_w;
  float* lse;
  View out;   // frame stores (ring)
  int ring;   // 1 / 2: `out` is the interior of a padded buffer whose replicate / reflect frame is written too
};

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  uint32_t u;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(u) : "f"(hi), "f"(lo));
  return u;
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

__global__ void __launch_bounds__(AF_THREADS, 1)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                   const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmR,
                   const __grid_constant__ CUtensorMap tmO, const AtP P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // keeps the .shared address space
  uint8_t* Qs = smem;                                   // [2][64 x 128B]
  uint8_t* Ks = Qs + 2 * AT_Q_BYTES;                    // [2][200 x 128B]
  uint8_t* RELs = Ks + 2 * AT_K_BYTES;                  // [200 x 128B]  row (r,c) = [rel_h[r] | rel_w[c]]
  uint8_t* Vs = RELs + AT_K_BYTES;                      // [2][208 x 128B]
  uint8_t* Ps = Vs + 2 * AT_V_BYTES;                    // [2][4][64 x 128B]
  uint8_t* ROs = Ps + 2 * AT_P_BYTES;                   // [2][64 x 128B] residual tile in, output tile out (in place)
  float* xmax = reinterpret_cast<float*>(ROs + AT_RO_BYTES);   // [2 parts][128 rows]
  float* xsum = xmax + 256;                                     // [2 parts][128 rows]
  uint64_t* bars = reinterpret_cast<uint64_t*>(xsum + 256);
  uint64_t* qk_full = bars + 0;
  uint64_t* qk_empty = bars + 1;
  uint64_t* v_full = bars + 2;
  uint64_t* pv_done = bars + 3;
  uint64_t* p_full = bars + 4;
  uint64_t* s_full = bars + 5;     // [2]
  uint64_t* o_free = bars + 7;     // O read out of TMEM (256 arrivals)
  uint64_t* ro_in = bars + 9;      // residual tile of the iteration has landed / the staging tile is free
  uint64_t* ro_out = bars + 10;    // output tile complete (256 arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 11);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // one-time smem constants that need no global memory: zero pad rows of K / V, zero P tile
  for (int i = threadIdx.x; i < 2 * (AT_NS - AT_NK) * 8; i += blockDim.x) {   // K rows 196..199
    const int h = i / ((AT_NS - AT_NK) * 8), rem = i % ((AT_NS - AT_NK) * 8);
    *reinterpret_cast<uint4*>(Ks + h * AT_K_BYTES + (AT_NK + rem / 8) * 128 + (rem % 8) * 16) = make_uint4(0, 0, 0, 0);
  }
  for (int i = threadIdx.x; i < 2 * (AT_NKP - AT_NK) * 8; i += blockDim.x) {  // V rows 196..207
    const int h = i / ((AT_NKP - AT_NK) * 8), rem = i % ((AT_NKP - AT_NK) * 8);
    *reinterpret_cast<uint4*>(Vs + h * AT_V_BYTES + (AT_NK + rem / 8) * 128 + (rem % 8) * 16) = make_uint4(0, 0, 0, 0);
  }
  for (int i = threadIdx.x; i < 2 * AT_P_BYTES / 16; i += blockDim.x)         // P columns 196..207 stay zero
    reinterpret_cast<uint4*>(Ps)[i] = make_uint4(0, 0, 0, 0);
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmQ);
    prefetch_tmap(&tmK);
    prefetch_tmap(&tmV);
    mbar_init(qk_full, 1);
    mbar_init(qk_empty, 1);
    mbar_init(v_full, 1);
    mbar_init(pv_done, 1);
    mbar_init(p_full, 256);
    mbar_init(ro_in, 1);
    mbar_init(ro_out, 256);
    mbar_init(o_free, 256);
    for (int r = 0; r < 2; ++r) {
      mbar_init(&s_full[r], 1);
    }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  griddep_wait();   // (programmatic dependent launch: the preamble above overlapped the previous kernel's tail)
  // the relative-position tile (swizzled like a TMA box), built from the rel_h / rel_w parameters
  for (int i = threadIdx.x; i < AT_NS * 8; i += blockDim.x) {   // REL rows 0..199, 16-byte chunks
    const int R = i >> 3, ch = i & 7;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = 0.f;
    if (R < AT_NK) {
      const int wr = R / 14, wc = R - wr * 14;
      const float* src = ch < 4 ? P.rel_h + wr * 32 + ch * 8 : P.rel_w + wc * 32 + (ch - 4) * 8;
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = src[j];
    }
    uint4 u;
    __nv_bfloat162* hh = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
    for (int j = 0; j < 4; ++j) hh[j] = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
    *reinterpret_cast<uint4*>(RELs + R * 128 + ((ch ^ (R & 7)) * 16)) = u;
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  griddep_launch();
  const uint32_t tmem_base = *tmem_slot;

  const int my_blocks = (P.nblocks - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int n_it = 2 * my_blocks;  // (block, head pair) iterations
  const bool tracing = P.trace && blockIdx.x == 0 && lane == 0 && (warp == 1 || warp == 2);
  auto stamp = [&](int it, int ev) {
    if (tracing && it < AB_TRACE_ITERS) g_attn_bwd_trace[it * AB_TRACE_EVENTS + ev] = clock64();
  };

  if (warp == 0) {
    // ================================ TMA producer ================================
    // Two independent streams polled by one thread: Q/K(it) as soon as S(it-1) has consumed the tiles, V(it) as soon
    // as P.V(it-1) has -- neither load may wait behind the other (S runs two iterations ahead of P.V).
    if (lane == 0) {
      int qk_it = 0, v_it = 0;
      while (qk_it < n_it || v_it < n_it) {
        if (qk_it < n_it && mbar_try_wait(qk_empty, (qk_it & 1) ^ 1)) {
          const int blk = blockIdx.x + (qk_it >> 1) * gridDim.x, pair = qk_it & 1;
          const int bx = blk % P.nbx, by = (blk / P.nbx) % P.nby, b = blk / (P.nbx * P.nby);
          mbar_expect_tx(qk_full, 2 * (AT_Q_BYTES + AT_KV_BOX_BYTES));
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int c0 = (pair * 2 + h) * 64;
            tma_load_4d(Qs + h * AT_Q_BYTES, &tmQ, qk_full, c0, bx * 8, by * 8, b);
            tma_load_4d(Ks + h * AT_K_BYTES, &tmK, qk_full, c0, bx * 8 - 3, by * 8 - 3, b);
          }
          ++qk_it;
        }
        if (v_it < n_it && mbar_try_wait(pv_done, (v_it & 1) ^ 1)) {   // V (and P) of the previous iteration consumed
          const int blk = blockIdx.x + (v_it >> 1) * gridDim.x, pair = v_it & 1;
          const int bx = blk % P.nbx, by = (blk / P.nbx) % P.nby, b = blk / (P.nbx * P.nby);
          mbar_expect_tx(v_full, 2 * AT_KV_BOX_BYTES);
#pragma unroll
          for (int h = 0; h < 2; ++h)
            tma_load_4d(Vs + h * AT_V_BYTES, &tmV, v_full, (pair * 2 + h) * 64, bx * 8 - 3, by * 8 - 3, b);
          ++v_it;
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
#ifdef PHT_ATTN_FWD_LANE0_ISSUE
    constexpr bool LANE0 = true;      // A/B switch: single-lane issue region (per-MMA broadcast loop in SASS)
#else
    constexpr bool LANE0 = false;     // all 32 lanes run the loop (warp-converged); one elected lane issues each instruction
#endif
    auto mma = [&](uint32_t d, uint64_t ad, uint64_t bd, uint32_t id, uint32_t acc) {
      if constexpr (LANE0) umma_bf16(d, ad, bd, id, acc);
      else umma_bf16_elect(d, ad, bd, id, acc);
    };
    auto commit = [&](uint64_t* bar) {
      if constexpr (LANE0) umma_commit(bar);
      else umma_commit_elect(bar);
    };
    if (!LANE0 || lane == 0) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(64, AT_NS, 0, 0);
      constexpr uint32_t idesc_pv = umma_idesc_bf16(64, 64, 0, 1);  // B = V is MN-major
      const uint64_t reld = umma_desc_k_sw128(smem_u32(RELs));
      auto issue_s = [&](int it) {
        const int r = it & 1;
        mbar_wait(qk_full, it & 1);   // (the S region r is free: the caller has seen p_full of iteration it-2)
        tc_fence_after();
        stamp(it, 0);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const uint64_t qd = umma_desc_k_sw128(smem_u32(Qs + h * AT_Q_BYTES));
          const uint64_t kd = umma_desc_k_sw128(smem_u32(Ks + h * AT_K_BYTES));
          const uint32_t d = tmem_base + r * AT_NS + ((uint32_t)(h * 16) << 16);
#pragma unroll
          for (int k = 0; k < 4; ++k) mma(d, qd + 2 * k, kd + 2 * k, idesc_s, k ? 1u : 0u);
#pragma unroll
          for (int k = 0; k < 4; ++k) mma(d, qd + 2 * k, reld + 2 * k, idesc_s, 1u);   // += q . [rel_h | rel_w]
        }
        commit(qk_empty);
        commit(&s_full[r]);
      };
      // S runs two iterations ahead of P.V: S(it+2) is issued right behind P.V(it), so the tensor pipe never has a
      // P.V waiting behind an S that itself waits for the epilogue.
      if (n_it > 0) issue_s(0);
      if (n_it > 1) issue_s(1);
      for (int it = 0; it < n_it; ++it) {
        mbar_wait(p_full, it & 1);               // P(it) written; all reads of S(it) done -> its region is free
        mbar_wait(v_full, it & 1);
        if (it > 0) mbar_wait(o_free, (it - 1) & 1);   // O(it-1) read out by the epilogue
        tc_fence_after();
        stamp(it, 1);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const uint32_t d = tmem_base + AT_COL_O + ((uint32_t)(h * 16) << 16);
          const uint32_t p_addr = smem_u32(Ps + h * AT_P_BYTES);
          const uint32_t v_addr = smem_u32(Vs + h * AT_V_BYTES);
#pragma unroll
          for (int kk = 0; kk < AT_NKP / 16; ++kk) {
            const uint64_t pd = umma_desc_k_sw128(p_addr + (kk >> 2) * 8192) + 2 * (kk & 3);
            const uint64_t vd = umma_desc_mn_sw128(v_addr + kk * 2048, 8192, 1024);
            mma(d, pd, vd, idesc_pv, kk ? 1u : 0u);
          }
        }
        commit(pv_done);
        if (it + 2 < n_it) issue_s(it + 2);
      }
    }
  } else if (warp == 6) {
    // ================================ output store + residual prefetch ================================
    if (lane == 0) {
      auto load_resid = [&](int it) {
        if (P.has_resid) {
          const int blk = blockIdx.x + (it >> 1) * gridDim.x, pair = it & 1;
          const int bx = blk % P.nbx, by = (blk / P.nbx) % P.nby, b = blk / (P.nbx * P.nby);
          mbar_expect_tx(ro_in, AT_RO_BYTES);
#pragma unroll
          for (int h = 0; h < 2; ++h)
            tma_load_4d(ROs + h * AT_Q_BYTES, &tmR, ro_in, (pair * 2 + h) * 64, bx * 8 + P.residOx, by * 8 + P.residOy, b);
        } else {
          mbar_arrive(ro_in);   // no residual: the staging tile is simply free
        }
      };
      if (n_it > 0) load_resid(0);
      for (int it = 0; it < n_it; ++it) {
        const int blk = blockIdx.x + (it >> 1) * gridDim.x, pair = it & 1;
        const int bx = blk % P.nbx, by = (blk / P.nbx) % P.nby, b = blk / (P.nbx * P.nby);
        mbar_wait(ro_out, it & 1);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                       ::"l"(reinterpret_cast<uint64_t>(&tmO)), "r"(smem_u32(ROs + h * AT_Q_BYTES)), "r"((pair * 2 + h) * 64),
                         "r"(bx * 8 + P.outOx), "r"(by * 8 + P.outOy), "r"(b)
                       : "memory");
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // the tile may be overwritten again
        if (it + 1 < n_it) load_resid(it + 1);
      }
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // all output stores complete before exit
    }
  } else {
    // ================================ softmax + epilogue (warps 2..5, 7..10) ================================
    const int quad = warp & 3;                   // TMEM sub-partition of this warp
    const int part = warp >= 7 ? 1 : 0;          // which of the two warps sharing the sub-partition
    const int hp = lane >> 4;                    // head inside the pair (TMEM lane half)
    const int q = quad * 16 + (lane & 15);       // query row
    const int xrow = hp * 64 + q;                // row in the exchange arrays
    const int qy = q >> 3, qx = q & 7;
    const int qsw = q & 7;
    const float LOG2E = 1.4426950408889634f;
    const uint32_t p_row = smem_u32(Ps + hp * AT_P_BYTES + q * 128);
    float prev_m = 0.f;
    int prev_blk = 0, prev_pair = 0;

    // O / sum + residual -> output tile (this warp's 32 of the 64 channels); needs the other group's partial sum.
    // Everything that does not depend on O (coordinates, 1 / sum, lse, the residual cells) is done BEFORE waiting for
    // P.V: the wait -> TMEM load -> scale -> store tail is on the kernel's critical path.
    auto epilogue = [&](int it, float m, int blk, int pair) {
      const float sum = xsum[xrow] + xsum[128 + xrow];
      const float inv = 1.f / sum;
      const int bx = blk % P.nbx, by = (blk / P.nbx) % P.nby, b = blk / (P.nbx * P.nby);
      if (P.lse && part == 0)
        P.lse[(((long long)b * P.H + by * 8 + qy) * P.W + bx * 8 + qx) * 4 + pair * 2 + hp] = m + logf(sum);
      // frame of the padded output buffer: copies of the edge pixel (replicate) or of the pixel next to it (reflect),
      // stored straight from the registers of the thread that owns it
      bf16 *d0 = nullptr, *d1 = nullptr, *d2 = nullptr;
#ifndef PHT_NO_RING
      if (P.ring) {
        const int x = bx * 8 + qx, y = by * 8 + qy, e0 = P.ring == 2 ? 1 : 0;
        const int tx = x == e0 ? -1 : (x == P.W - 1 - e0 ? P.W : -2);
        const int ty = y == e0 ? -1 : (y == P.H - 1 - e0 ? P.H : -2);
        bf16* base = (bf16*)P.out.ptr + (pair * 2 + hp) * 64 + part * 32;
        if (tx != -2) d0 = base + view_off(P.out, b, y + P.out.oy, tx + P.out.ox);
        if (ty != -2) d1 = base + view_off(P.out, b, ty + P.out.oy, x + P.out.ox);
        if (tx != -2 && ty != -2) d2 = base + view_off(P.out, b, ty + P.out.oy, tx + P.out.ox);
      }
#endif
      uint8_t* row = ROs + hp * AT_Q_BYTES + q * 128;
      uint4 ru[4];
      mbar_wait(ro_in, it & 1);                  // residual tile landed (or staging tile free)
#pragma unroll
      for (int g = 0; g < 4; ++g)
        ru[g] = P.has_resid ? *reinterpret_cast<const uint4*>(row + (((part * 4 + g) ^ qsw) * 16)) : make_uint4(0, 0, 0, 0);
      mbar_wait(pv_done, it & 1);
      tc_fence_after();
      stamp(it, 6);
      uint32_t o[32];
      tmem_ld32(tmem_base + AT_COL_O + ((uint32_t)(quad * 32) << 16) + part * 32, o);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(o_free);
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const uint32_t rw[4] = {ru[g].x, ru[g].y, ru[g].z, ru[g].w};
        uint32_t w[4];
#pragma unroll
        for (int j = 0; j < 4; ++j)
          w[j] = pack_bf16x2(fmaf(__uint_as_float(o[g * 8 + 2 * j]), inv, __uint_as_float(rw[j] << 16)),
                             fmaf(__uint_as_float(o[g * 8 + 2 * j + 1]), inv, __uint_as_float(rw[j] & 0xffff0000u)));
        const uint4 q4 = make_uint4(w[0], w[1], w[2], w[3]);
        *reinterpret_cast<uint4*>(row + (((part * 4 + g) ^ qsw) * 16)) = q4;
        if (d0) *reinterpret_cast<uint4*>(d0 + g * 8) = q4;
        if (d1) *reinterpret_cast<uint4*>(d1 + g * 8) = q4;
        if (d2) *reinterpret_cast<uint4*>(d2 + g * 8) = q4;
      }
      fence_proxy_async();
      mbar_arrive(ro_out);
    };

    // one softmax iteration for the columns of warp group PART (compile-time): 13 / 12 groups of 8 columns
    auto softmax_it = [&](auto part_c, int it, uint32_t t_addr) {
      constexpr int PART = decltype(part_c)::value;
      constexpr int C0 = PART * 104;                       // first S column of this group
      constexpr int NG = PART ? 12 : 13;                   // 8-column groups (the last group of PART 1 holds keys 192..195)
      constexpr int NCH = (NG + 3) / 4;                    // 32-column TMEM loads (PART 0's last one is 8 columns wide)
      uint32_t a[32], bq[32];
      // ---- pass 1: row maximum of the own columns ----
      float m = -INFINITY;
      tmem_ld32(t_addr + C0, a);
      tmem_ld_wait();
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        uint32_t* cur = (c & 1) ? bq : a;
        uint32_t* nxt = (c & 1) ? a : bq;
        if (c + 1 < NCH) {
          if (PART == 0 && c + 1 == NCH - 1) tmem_ld16(t_addr + C0 + (c + 1) * 32, nxt);   // 8 valid columns
          else tmem_ld32(t_addr + C0 + (c + 1) * 32, nxt);
        }
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
          const int col = C0 + c * 32 + j;
          if (col < (PART ? AT_NK : 104)) m = fmaxf(m, fmaxf(__uint_as_float(cur[j]), __uint_as_float(cur[j + 1])));
        }
        if (c + 1 < NCH) tmem_ld_wait();
      }
      xmax[PART * 128 + xrow] = m;
      asm volatile("bar.sync 1, 256;" ::: "memory");       // both groups' maxima (and the previous sums) are visible
      m = fmaxf(xmax[xrow], xmax[128 + xrow]);
      stamp(it, 3);
      // the previous pair's O is final by now: write it out and free its TMEM region for S(it+1)
      if (it > 0) epilogue(it - 1, prev_m, prev_blk, prev_pair);
      stamp(it, 4);
      // ---- pass 2: p = exp(s - m), partial row sum, bf16 P tile (K-major, 128B swizzle) ----
      const float m2 = m * LOG2E;
      float s0 = 0.f, s1 = 0.f;
      tmem_ld32(t_addr + C0, a);
      tmem_ld_wait();
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        uint32_t* cur = (c & 1) ? bq : a;
        uint32_t* nxt = (c & 1) ? a : bq;
        if (c + 1 < NCH) {
          if (PART == 0 && c + 1 == NCH - 1) tmem_ld16(t_addr + C0 + (c + 1) * 32, nxt);
          else tmem_ld32(t_addr + C0 + (c + 1) * 32, nxt);
        }
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int gi = c * 4 + g;
          if (gi < NG) {
            float p[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const int key = C0 + gi * 8 + j;
              p[j] = key < AT_NK ? ex2(fmaf(__uint_as_float(cur[g * 8 + j]), LOG2E, -m2)) : 0.f;
            }
            s0 += (p[0] + p[1]) + (p[2] + p[3]);
            s1 += (p[4] + p[5]) + (p[6] + p[7]);
            const int key0 = C0 + gi * 8;
            st_shared_v4(p_row + (key0 >> 6) * 8192 + ((((key0 & 63) >> 3) ^ qsw) * 16), pack_bf16x2(p[0], p[1]),
                         pack_bf16x2(p[2], p[3]), pack_bf16x2(p[4], p[5]), pack_bf16x2(p[6], p[7]));
          }
        }
        if (c + 1 < NCH) tmem_ld_wait();
      }
      fence_proxy_async();   // generic-proxy smem writes -> visible to the tensor core (async proxy)
      tc_fence_before();     // all TMEM reads of S done before the MMA warp may overwrite the region with O
      mbar_arrive(p_full);
      stamp(it, 5);
      asm volatile("bar.sync 2, 256;" ::: "memory");       // every thread has consumed the previous partial sums
      xsum[PART * 128 + xrow] = s0 + s1;
      return m;
    };

    for (int it = 0; it < n_it; ++it) {
      const int blk = blockIdx.x + (it >> 1) * gridDim.x, pair = it & 1;
      const int r = it & 1;
      mbar_wait(&s_full[r], (it >> 1) & 1);
      tc_fence_after();
      stamp(it, 2);
      const uint32_t t_addr = tmem_base + r * AT_NS + ((uint32_t)(quad * 32) << 16);
      float m;
      if (part == 0) m = softmax_it(std::integral_constant<int, 0>{}, it, t_addr);
      else m = softmax_it(std::integral_constant<int, 1>{}, it, t_addr);
      prev_m = m; prev_blk = blk; prev_pair = pair;
    }
    if (n_it > 0) {
      asm volatile("bar.sync 1, 256;" ::: "memory");       // the last partial sums are visible
      epilogue(n_it - 1, prev_m, prev_blk, prev_pair);
    }
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
  if (a->ring < 0 || a->ring > 2 || (a->ring && (a->H < 4 || a->W < 4))) return PHT_OK;
  if (!get_encode_fn()) return PHT_OK;
  CUtensorMap tmQ, tmK, tmV, tmR, tmO;
  int rc = at_tmap(&tmQ, a->q, a->B, 8, 8);
  if (rc) return rc;
  rc = at_tmap(&tmK, a->k, a->B, 14, 14);
  if (rc) return rc;
  rc = at_tmap(&tmV, a->v, a->B, 14, 14);
  if (rc) return rc;
  rc = at_tmap(&tmO, a->out, a->B, 8, 8);
  if (rc) return rc;
  tmR = tmO;
  if (a->resid.ptr) {
    rc = at_tmap(&tmR, a->resid, a->B, 8, 8);
    if (rc) return rc;
  }
  AtP P;
  P.B = a->B; P.H = a->H; P.W = a->W; P.nbx = a->W / 8; P.nby = a->H / 8; P.nblocks = a->B * P.nbx * P.nby;
  P.has_resid = a->resid.ptr ? 1 : 0;
  P.residOy = a->resid.oy; P.residOx = a->resid.ox; P.outOy = a->out.oy; P.outOx = a->out.ox;
  P.rel_h = a->rel_h; P.rel_w = a->rel_w; P.lse = a->lse;
  P.out = make_view(a->out); P.ring = a->ring;
  P.trace = g_attn_trace_on == 2;
  PHT_SMEM_ATTR_ONCE(attn_fwd_tc_kernel, AT_SMEM);
  const int sms = sm_count();
  int grid = P.nblocks < sms ? P.nblocks : sms;
  PHT_CUDA(launch_pdl(attn_fwd_tc_kernel, dim3(grid), dim3(AF_THREADS), AT_SMEM, st, tmQ, tmK, tmV, tmR, tmO, P));
  PHT_LAUNCH_CHECK();
  walk_dir_set(a->out.ptr, 0);   // written first-to-last
  count_launch(CNT_ATTN_TC);
  *handled = true;
  return PHT_OK;
}


// =================================================================================================
// Backward (recompute), one (block, head) per iteration, everything on tcgen05:
//   S = Q K^T + Q REL^T, dP = dO V^T    (keys 0..103 in TMEM lanes +0, keys 104..207 in lanes +16: both 16-lane
//                                         halves of every sub-partition hold one query row)
//   P = exp(S - lse), delta = sum P.dP, dS = P.(dP - delta)   -> bf16 P / dS tiles in smem (128B swizzle)
//   dQ = dS K + dS REL                   (== dS.(K + rel) exactly, REL = the constant [key][rel_h[r] | rel_w[c]] tile)
//   dV = P^T dO, dK = dS^T Q             (A operands MN-major: the [query x key] tiles are read transposed)
//   dREL = sum over the CTA's iterations of dK, accumulated in fp32 registers by the read-out threads (its
//          window-row / window-column sums are d rel_h / d rel_w)
// Warps: 0 TMA producer, 1 MMA issuer, 2..9 softmax + read-out.  Two warps share each TMEM sub-partition: they split
// the S / dP columns of a query row between them (partial deltas exchanged through smem) and the dV / dK key tiles.
// Pipeline: the operands of iteration i+1 are prefetched (Q, dO, K double-buffered; V reloaded as soon as dP(i) is
// done), S/dP(i+1) is issued right behind dQ/dV/dK(i) on the tensor pipe, and the dV/dK read-out of iteration i
// runs while S/dP(i+1) is being computed.  TMEM: S [0,104) dP [104,208) (dQ re-uses [0,64)), dV [208,336), dK [336,464).
// dK / dV: every read-out warp transposes its 32 key rows x 32 channels through a small smem tile so that each memory
// instruction covers eight complete 64-byte row segments, and then either
//   P.direct = 1 (default)  ADDS them straight into the final NHWC gradient with vector reductions
//                (red.global.add.noftz.v4.bf16x2; the read-modify-write happens in L2, window pixels outside the image are
//                skipped; the outputs are zeroed beforehand, attn_bwd_zero_tc).  No scratch, no second pass.  The <= 4
//                overlapping windows of a key pixel are added in arrival order, each add rounding to bf16: the result
//                is NOT bit-reproducible from run to run (it differs by the rounding order of at most four terms);
//   P.direct = 0 ("attn_bwd_direct" = 0, the bit-reproducible mode) stores them window-major in a bf16 scratch and a fold
//                kernel sums the <= 4 windows of every pixel in a fixed order in fp32 (round 1's scheme).
// Tried and dropped (measured, see profiles/README.md): the direct adds in a FIXED order (per-block flags in global
// memory, priority = (iteration, parity colour), polled / published by helper warps): the GPU-scope fence + flag
// round trip costs ~4 us per (head, tensor) against a 3.7 us iteration -- 566 us per layer; and one TMA reduce
// (cp.reduce.async.bulk.tensor .add on a bf16 tensor map) per window tile, which faults ("illegal instruction") here.
// =================================================================================================
#ifndef PHT_AB_DK_EARLY
#define PHT_AB_DK_EARLY 0
#endif
constexpr bool AB_DK_EARLY = PHT_AB_DK_EARLY != 0;   // A/B: dK rows read out right behind the dV rows (under S/dP(i+1)) instead of behind the next softmax
constexpr int AB_THREADS = 320;
constexpr int AB_HALF = 104;                                    // keys per lane half (2 x 104 = 208 >= 196)
constexpr int AB_ROWS = 2 * AB_HALF;                            // K / V / REL tile rows
constexpr int AB_Q_BYTES = 8192, AB_K_BYTES = AB_ROWS * 128;
constexpr int AB_P_BYTES = 4 * 8192, AB_DS_BYTES = 4 * 8192;
constexpr int AB_STAGE_BYTES = 2 * AB_Q_BYTES + AB_K_BYTES;     // Q, dO, K
constexpr int AB_STG_BYTES = 32 * 64;                            // per-warp read-out staging tile: 32 key rows x 32 ch bf16
constexpr int AB_SMEM = 2 * AB_STAGE_BYTES + 2 * AB_K_BYTES + AB_P_BYTES + AB_DS_BYTES + 8 * AB_STG_BYTES + 1024 + 1024;
static_assert(AB_SMEM <= 232448, "attn_bwd_tc: shared memory budget");
constexpr int AB_COL_DP = AB_HALF, AB_COL_DQ = 0, AB_COL_DV = 2 * AB_HALF, AB_COL_DK = 2 * AB_HALF + 128;
static_assert(AB_COL_DK + 128 <= 512, "attn_bwd_tc: TMEM budget");
constexpr int AB_PART0 = 56;                                    // S/dP columns of a lane half handled by warp part 0 (part 1: 48)
constexpr int AB_REL_PART = AT_NK * 64;  // floats per CTA partial: dREL [196 keys][64]

struct AbP {
  int B, H, W, nbx, nby, nblocks, trace;
  int direct;         // 1: vector reductions straight into dk / dv; 0: window-major scratch + fold kernel
  View dq, dk, dv;
  const float* rel_h;
  const float* rel_w;
  const float* lse;
  bf16* dk_scratch;   // [nblocks][4][196][64] (direct == 0)
  bf16* dv_scratch;
  float* rel_part;    // [gridDim.x][196][64]
};

__device__ __forceinline__ void st_row32_bf16(bf16* dst, const uint32_t* r) {
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    uint4 u;
    __nv_bfloat162* uh = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
    for (int j = 0; j < 4; ++j) uh[j] = __floats2bfloat162_rn(__uint_as_float(r[g * 8 + 2 * j]), __uint_as_float(r[g * 8 + 2 * j + 1]));
    *reinterpret_cast<uint4*>(dst + g * 8) = u;
  }
}

__global__ void __launch_bounds__(AB_THREADS, 1)
attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                   const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmDO, const AbP P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // keeps the .shared address space
  uint8_t* St = smem;                                   // [2 stages][Q 8K | dO 8K | K 26K]
  uint8_t* RELs = St + 2 * AB_STAGE_BYTES;
  uint8_t* Vs = RELs + AB_K_BYTES;
  uint8_t* Ps = Vs + AB_K_BYTES;
  uint8_t* dSs = Ps + AB_P_BYTES;
  uint8_t* STG = dSs + AB_DS_BYTES;                      // [8 warps][32 rows x 64 B]
  float* dpart = reinterpret_cast<float*>(STG + 8 * AB_STG_BYTES);   // [2 parts][64 queries] partial deltas
  uint64_t* bars = reinterpret_cast<uint64_t*>(dpart + 128);
  uint64_t* qk_full = bars + 0;    // [2]
  uint64_t* qk_empty = bars + 2;   // [2]
  uint64_t* v_full = bars + 4;
  uint64_t* v_empty = bars + 5;
  uint64_t* sdp_full = bars + 6;
  uint64_t* ds_full = bars + 7;
  uint64_t* dq_full = bars + 8;
  uint64_t* dq_free = bars + 9;
  uint64_t* out_full = bars + 10;
  uint64_t* dvk_free = bars + 11;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // ---- one-time smem constants that need no global memory ----------------------------------------------------
  for (int i = threadIdx.x; i < 3 * (AB_ROWS - AT_NK) * 8; i += blockDim.x) {   // K (both stages) / V rows 196..207
    const int which = i / ((AB_ROWS - AT_NK) * 8), rem = i % ((AB_ROWS - AT_NK) * 8);
    uint8_t* tile = which < 2 ? St + which * AB_STAGE_BYTES + 2 * AB_Q_BYTES : Vs;
    *reinterpret_cast<uint4*>(tile + (AT_NK + rem / 8) * 128 + (rem % 8) * 16) = make_uint4(0, 0, 0, 0);
  }
  for (int i = threadIdx.x; i < (AB_P_BYTES + AB_DS_BYTES) / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(Ps)[i] = make_uint4(0, 0, 0, 0);
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmQ);
    prefetch_tmap(&tmK);
    prefetch_tmap(&tmV);
    prefetch_tmap(&tmDO);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&qk_full[s], 1);
      mbar_init(&qk_empty[s], 1);
    }
    mbar_init(v_full, 1);
    mbar_init(v_empty, 1);
    mbar_init(sdp_full, 1);
    mbar_init(ds_full, 256);
    mbar_init(dq_full, 1);
    mbar_init(dq_free, 256);
    mbar_init(out_full, 1);
    mbar_init(dvk_free, 256);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  griddep_wait();   // (programmatic dependent launch: the preamble above overlapped the previous kernel's tail)
  // the relative-position tile, built from the rel_h / rel_w parameters
  for (int i = threadIdx.x; i < AB_ROWS * 8; i += blockDim.x) {   // REL tile rows 0..207
    const int R = i >> 3, ch = i & 7;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = 0.f;
    if (R < AT_NK) {
      const int wr = R / 14, wc = R - wr * 14;
      const float* src = ch < 4 ? P.rel_h + wr * 32 + ch * 8 : P.rel_w + wc * 32 + (ch - 4) * 8;
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = src[j];
    }
    uint4 u;
    __nv_bfloat162* hh = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
    for (int j = 0; j < 4; ++j) hh[j] = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
    *reinterpret_cast<uint4*>(RELs + R * 128 + ((ch ^ (R & 7)) * 16)) = u;
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  griddep_launch();
  const uint32_t tmem_base = *tmem_slot;
  const int my_blocks = (P.nblocks - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int n_it = 4 * my_blocks;  // (block, head) iterations
  const bool tracing = P.trace && blockIdx.x == 0 && lane == 0 && (warp == 1 || warp == 2);
  auto stamp = [&](int it, int ev) {
    if (tracing && it < AB_TRACE_ITERS) g_attn_bwd_trace[it * AB_TRACE_EVENTS + ev] = clock64();
  };

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (lane == 0) {
      for (int it = 0; it < n_it; ++it) {
        const int blk = blockIdx.x + (it >> 2) * gridDim.x, head = it & 3;
        const int bx = blk % P.nbx, by = (blk / P.nbx) % P.nby, b = blk / (P.nbx * P.nby);
        const int s = it & 1;
        uint8_t* st = St + s * AB_STAGE_BYTES;
        mbar_wait(&qk_empty[s], ((it >> 1) & 1) ^ 1);   // dQ/dV/dK MMAs of iteration it-2 have read this stage
        mbar_expect_tx(&qk_full[s], 2 * AB_Q_BYTES + AT_KV_BOX_BYTES);
        tma_load_4d(st, &tmQ, &qk_full[s], head * 64, bx * 8, by * 8, b);
        tma_load_4d(st + AB_Q_BYTES, &tmDO, &qk_full[s], head * 64, bx * 8, by * 8, b);
        tma_load_4d(st + 2 * AB_Q_BYTES, &tmK, &qk_full[s], head * 64, bx * 8 - 3, by * 8 - 3, b);
        mbar_wait(v_empty, (it & 1) ^ 1);               // dP of iteration it-1 complete
        mbar_expect_tx(v_full, AT_KV_BOX_BYTES);
        tma_load_4d(Vs, &tmV, v_full, head * 64, bx * 8 - 3, by * 8 - 3, b);
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    {   // all 32 lanes run the loop (warp-converged); one elected lane issues each tcgen05 instruction
      constexpr uint32_t id_s = umma_idesc_bf16(64, AB_HALF, 0, 0);
      constexpr uint32_t id_dvk = umma_idesc_bf16(128, 64, 1, 1);
      constexpr uint32_t id_dq = umma_idesc_bf16(64, 64, 0, 1);
      const uint32_t rel_a = smem_u32(RELs), v_a = smem_u32(Vs), p_a = smem_u32(Ps), ds_a = smem_u32(dSs);
      for (int it = 0; it < n_it; ++it) {
        const int s = it & 1;
        const uint32_t q_a = smem_u32(St + s * AB_STAGE_BYTES), do_a = q_a + AB_Q_BYTES, k_a = q_a + 2 * AB_Q_BYTES;
        mbar_wait(&qk_full[s], (it >> 1) & 1);
        mbar_wait(v_full, it & 1);
        if (it > 0) mbar_wait(dq_free, (it - 1) & 1);   // dQ(it-1) read out of the columns S(it) is about to overwrite
        tc_fence_after();
        stamp(it, 0);
        const uint64_t qd = umma_desc_k_sw128(q_a), dod = umma_desc_k_sw128(do_a);
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {  // lane half hf handles keys [104*hf, 104*hf + 104)
          const uint32_t d = tmem_base + ((uint32_t)(hf * 16) << 16);
          const uint64_t kd = umma_desc_k_sw128(k_a + hf * AB_HALF * 128);
          const uint64_t rd = umma_desc_k_sw128(rel_a + hf * AB_HALF * 128);
          const uint64_t vd = umma_desc_k_sw128(v_a + hf * AB_HALF * 128);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16_elect(d + AB_COL_DP, dod + 2 * k, vd + 2 * k, id_s, k ? 1u : 0u);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16_elect(d, qd + 2 * k, kd + 2 * k, id_s, k ? 1u : 0u);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16_elect(d, qd + 2 * k, rd + 2 * k, id_s, 1u);          // += q . [rel_h | rel_w]
        }
        umma_commit_elect(v_empty);
        umma_commit_elect(sdp_full);
        if (tracing) {                 // diagnostics: when did S/dP(it) actually complete?
          mbar_wait(sdp_full, it & 1);
          stamp(it, 7);
        }
        mbar_wait(ds_full, it & 1);
        tc_fence_after();
        stamp(it, 1);
#pragma unroll
        for (int kk = 0; kk < 13; ++kk) {   // dQ first (keys 0..207; 196..207 are zero columns of dS)
          const uint64_t dsa = umma_desc_k_sw128(ds_a + (kk >> 2) * 8192) + 2 * (kk & 3);
          const uint64_t kb = umma_desc_mn_sw128(k_a + kk * 2048, 8192, 1024);
          const uint64_t rb = umma_desc_mn_sw128(rel_a + kk * 2048, 8192, 1024);
          umma_bf16_elect(tmem_base + AB_COL_DQ, dsa, kb, id_dq, kk ? 1u : 0u);
          umma_bf16_elect(tmem_base + AB_COL_DQ, dsa, rb, id_dq, 1u);
        }
        umma_commit_elect(dq_full);
        if (tracing) {
          mbar_wait(dq_full, it & 1);
          stamp(it, 8);
        }
        if (it > 0) {
          mbar_wait(dvk_free, (it - 1) & 1);            // dV/dK(it-1) read out
          tc_fence_after();
        }
#pragma unroll
        for (int t2 = 0; t2 < 2; ++t2) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t pa = umma_desc_mn_sw128(p_a + t2 * 16384 + k * 2048, 8192, 1024);
            const uint64_t dsa = umma_desc_mn_sw128(ds_a + t2 * 16384 + k * 2048, 8192, 1024);
            const uint64_t dob = umma_desc_mn_sw128(do_a + k * 2048, 8192, 1024);
            const uint64_t qb = umma_desc_mn_sw128(q_a + k * 2048, 8192, 1024);
            umma_bf16_elect(tmem_base + AB_COL_DV + t2 * 64, pa, dob, id_dvk, k ? 1u : 0u);
            umma_bf16_elect(tmem_base + AB_COL_DK + t2 * 64, dsa, qb, id_dvk, k ? 1u : 0u);
          }
        }
        umma_commit_elect(out_full);
        umma_commit_elect(&qk_empty[s]);
        if (tracing) {
          mbar_wait(out_full, it & 1);
          stamp(it, 9);
        }
      }
    }
  } else {
    // ================================ softmax / dS / read-out (warps 2..9) ================================
    const int sp = warp & 3;                     // TMEM sub-partition of this warp
    const int part = (warp - 2) >> 2;            // which of the two warps sharing the sub-partition
    const int hf = lane >> 4;                    // key half handled by this thread
    const int q = sp * 16 + (lane & 15);         // query row
    const int qy = q >> 3, qx = q & 7;
    const int qsw = q & 7;
    const float LOG2E = 1.4426950408889634f;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(sp * 32) << 16);
    const int col0 = part * AB_PART0;            // first S/dP column of this thread (7 / 6 groups of 8 columns)
    const uint32_t p_row = smem_u32(Ps + q * 128), ds_row = smem_u32(dSs + q * 128);
    uint32_t goff[7];                            // byte offset of the 16-byte cell of column group g inside a P / dS row
#pragma unroll
    for (int g = 0; g < 7; ++g) {
      const int key0 = hf * AB_HALF + col0 + g * 8;
      goff[g] = (uint32_t)((key0 >> 6) * 8192 + ((((key0 & 63) >> 3) ^ qsw) * 16));
    }
    const int row0 = part * 128 + sp * 32;            // first key row of dV / dK read out by this warp
    const bool rows_valid = row0 < AT_NK;             // warp-uniform (the last warp's rows are all padding)
    uint8_t* stage = STG + (warp - 2) * AB_STG_BYTES; // [32 rows][64 B] bf16, 64B-swizzled
    const uint32_t stage_row = smem_u32(stage) + lane * 64;
    const int ssw = (lane >> 1) & 3;
    float racc[64];                                   // dREL row accumulator (== sum of this thread's dK rows)
#pragma unroll
    for (int j = 0; j < 64; ++j) racc[j] = 0.f;
    auto lse_of = [&](int it) -> float {
      const int blk = blockIdx.x + (it >> 2) * gridDim.x, head = it & 3;
      const int bx = blk % P.nbx, by = (blk / P.nbx) % P.nby, b = blk / (P.nbx * P.nby);
      return P.lse[(((long long)b * P.H + by * 8 + qy) * P.W + bx * 8 + qx) * 4 + head];
    };
    float lse_next = n_it > 0 ? lse_of(0) : 0.f;   // raw value: consumed (scaled) one iteration later, so the load never stalls

    // P = exp(S - lse) -> smem (and packed registers), partial delta; then dS = P (dP - delta) -> smem.
    // PART is a compile-time copy of `part` so that group counts and the padding mask fold away.
    auto softmax_ds = [&](auto part_c, float l2) {
      constexpr int PART = decltype(part_c)::value;
      constexpr int NG = PART ? (AB_HALF - AB_PART0) / 8 : AB_PART0 / 8;   // 6 / 7 groups of 8 columns
      constexpr int NC = (NG + 1) / 2;                                     // 16-column TMEM loads
      constexpr int C0 = PART * AB_PART0;
      uint32_t ppk[NG * 4];                      // this thread's probabilities, bf16x2
      uint32_t s[2][16], dp[2][16];
      float d0 = 0.f, d1 = 0.f;
      tmem_ld16(lane_addr + C0, s[0]);
      tmem_ld16(lane_addr + AB_COL_DP + C0, dp[0]);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < NC; ++i) {
        if (i + 1 < NC) {                        // next 16 columns in flight while these are processed
          tmem_ld16(lane_addr + C0 + (i + 1) * 16, s[(i + 1) & 1]);
          tmem_ld16(lane_addr + AB_COL_DP + C0 + (i + 1) * 16, dp[(i + 1) & 1]);
        }
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          constexpr int dummy = 0;
          (void)dummy;
          const int gi = i * 2 + g;
          if (gi < NG) {
            float p[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              p[j] = ex2(fmaf(__uint_as_float(s[i & 1][g * 8 + j]), LOG2E, -l2));
              // keys 196..207 (second lane half, columns 92..103) are padding
              if (PART == 1 && C0 + gi * 8 + j >= AT_NK - AB_HALF) p[j] = hf ? 0.f : p[j];
            }
#pragma unroll
            for (int j = 0; j < 8; j += 2) {
              d0 = fmaf(p[j], __uint_as_float(dp[i & 1][g * 8 + j]), d0);
              d1 = fmaf(p[j + 1], __uint_as_float(dp[i & 1][g * 8 + j + 1]), d1);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) ppk[gi * 4 + j] = pack_bf16x2(p[2 * j], p[2 * j + 1]);
            st_shared_v4(p_row + goff[gi], ppk[gi * 4], ppk[gi * 4 + 1], ppk[gi * 4 + 2], ppk[gi * 4 + 3]);
          }
        }
        if (i + 1 < NC) tmem_ld_wait();
      }
      float delta = d0 + d1;
      delta += __shfl_xor_sync(0xffffffffu, delta, 16);
      if (hf == 0) dpart[PART * 64 + q] = delta;
      tmem_ld16(lane_addr + AB_COL_DP + C0, dp[0]);   // re-read dP while waiting for the other warp's partial
      asm volatile("bar.sync 1, 256;" ::: "memory");
      delta = dpart[q] + dpart[64 + q];
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < NC; ++i) {
        if (i + 1 < NC) tmem_ld16(lane_addr + AB_COL_DP + C0 + (i + 1) * 16, dp[(i + 1) & 1]);
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          const int gi = i * 2 + g;
          if (gi < NG) {
            uint32_t u[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint32_t pw = ppk[gi * 4 + j];
              const float plo = __uint_as_float(pw << 16), phi = __uint_as_float(pw & 0xffff0000u);
              u[j] = pack_bf16x2(plo * (__uint_as_float(dp[i & 1][g * 8 + 2 * j]) - delta),
                                 phi * (__uint_as_float(dp[i & 1][g * 8 + 2 * j + 1]) - delta));
            }
            st_shared_v4(ds_row + goff[gi], u[0], u[1], u[2], u[3]);
          }
        }
        if (i + 1 < NC) tmem_ld_wait();
      }
    };

    // (block, head) -> coordinates; the block index only changes every 4th iteration
    int cbx = 0, cby = 0, cb = 0;
    auto coords = [&](int it) {
      const int blk = blockIdx.x + (it >> 2) * gridDim.x;
      cbx = blk % P.nbx; cby = (blk / P.nbx) % P.nby; cb = blk / (P.nbx * P.nby);
    };
    // read-out helper: one 32-column chunk of this warp's dV / dK rows (thread = key row) is transposed through a
    // swizzled 2 KB smem tile so that every reduction instruction covers eight complete 64-byte row segments of the
    // NHWC gradient (a row-per-thread access would touch 32 different lines per instruction)
    auto stage_red = [&](const uint32_t* r, const View& o, bf16* scratch, int zrow, int bx, int by, int b, int ch) {
      if (rows_valid) {
        __syncwarp();                            // the previous chunk's reads of the tile are done
#pragma unroll
        for (int g = 0; g < 4; ++g)
          st_shared_v4(stage_row + ((g ^ ssw) * 16), pack_bf16x2(__uint_as_float(r[g * 8]), __uint_as_float(r[g * 8 + 1])),
                       pack_bf16x2(__uint_as_float(r[g * 8 + 2]), __uint_as_float(r[g * 8 + 3])),
                       pack_bf16x2(__uint_as_float(r[g * 8 + 4]), __uint_as_float(r[g * 8 + 5])),
                       pack_bf16x2(__uint_as_float(r[g * 8 + 6]), __uint_as_float(r[g * 8 + 7])));
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int rr = i * 8 + (lane >> 2);    // tile row handled by this lane in pass i
          const int key = row0 + rr, wy = key / 14, wx = key - wy * 14;
          const int y = by * 8 - 3 + wy, x = bx * 8 - 3 + wx;
          uint4 v;
          asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                       : "r"(smem_u32(stage) + rr * 64 + (((lane & 3) ^ ((rr >> 1) & 3)) << 4)));
          if (!P.direct) {   // window-major scratch [block * 4 + head][key][64]
            if (key < AT_NK)
              *reinterpret_cast<uint4*>(scratch + ((long long)zrow * AT_NK + key) * 64 + (ch & 63) + (lane & 3) * 8) = v;
          } else if (key < AT_NK && (unsigned)y < (unsigned)P.H && (unsigned)x < (unsigned)P.W) {
            bf16* dst = (bf16*)o.ptr + view_off(o, b, y + o.oy, x + o.ox) + ch + (lane & 3) * 8;
            asm volatile("red.global.add.noftz.v4.bf16x2 [%0], {%1, %2, %3, %4};"
                         ::"l"(dst), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
          }
        }
      }
    };
    // dK rows of a finished iteration (+ dREL accumulation); releases the dV / dK columns to the MMA warp
    auto readout_dk = [&](int zrow, int head, int bx, int by, int b, int it_dbg) {
      uint32_t a[32], c[32];
      tmem_ld32(lane_addr + AB_COL_DK + part * 64, a);
      tmem_ld_wait();
      stamp(it_dbg, 11);
      tmem_ld32(lane_addr + AB_COL_DK + part * 64 + 32, c);
#pragma unroll
      for (int j = 0; j < 32; ++j) racc[j] += __uint_as_float(a[j]);
      stage_red(a, P.dk, P.dk_scratch, zrow, bx, by, b, head * 64);
      stamp(it_dbg, 6);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(dvk_free);
#pragma unroll
      for (int j = 0; j < 32; ++j) racc[32 + j] += __uint_as_float(c[j]);
      stage_red(c, P.dk, P.dk_scratch, zrow, bx, by, b, head * 64 + 32);
    };

    int pz = 0, phead = 0, pbx = 0, pby = 0, pb = 0;   // the iteration whose dK rows are still to be read out
    if (n_it > 0) coords(0);
    for (int it = 0; it < n_it; ++it) {
      const int blk = blockIdx.x + (it >> 2) * gridDim.x, head = it & 3;
      const int bx = cbx, by = cby, b = cb;
      const float l2 = lse_next * LOG2E;
      mbar_wait(sdp_full, it & 1);
      tc_fence_after();
      stamp(it, 2);
      if (part == 0) softmax_ds(std::integral_constant<int, 0>{}, l2);
      else softmax_ds(std::integral_constant<int, 1>{}, l2);
      fence_proxy_async();
      tc_fence_before();
      mbar_arrive(ds_full);
      stamp(it, 3);
      // While the tensor core computes dQ(it): the deferred half of the previous iteration's read-out (dK rows), the
      // next iteration's coordinates and lse.  dV/dK(it) are issued behind dQ(it) and wait for dvk_free.
      if (!AB_DK_EARLY && it > 0) readout_dk(pz, phead, pbx, pby, pb, it);
      stamp(it, 10);
      if (it + 1 < n_it) {
        if (((it + 1) & 3) == 0) coords(it + 1);
        lse_next = P.lse[(((long long)cb * P.H + cby * 8 + qy) * P.W + cbx * 8 + qx) * 4 + ((it + 1) & 3)];
      }
      // ---- read-out: dQ (32 channels per warp part) ----
      mbar_wait(dq_full, it & 1);
      tc_fence_after();
      stamp(it, 4);
      {
        uint32_t a[32];
        tmem_ld32(lane_addr + AB_COL_DQ + part * 32, a);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(dq_free);
        if (hf == 0) st_row32_bf16((bf16*)P.dq.ptr + view_off(P.dq, b, by * 8 + qy, bx * 8 + qx) + head * 64 + part * 32, a);
      }
      // ---- read-out: dV rows -> 64B-swizzled smem tile -> vector reductions into the NHWC gradient (the dK rows
      //      follow after the next softmax) ----
      mbar_wait(out_full, it & 1);
      tc_fence_after();
      stamp(it, 5);
      {
        uint32_t a[32], c[32];
        tmem_ld32(lane_addr + AB_COL_DV + part * 64, a);
        tmem_ld_wait();
        tmem_ld32(lane_addr + AB_COL_DV + part * 64 + 32, c);
        stage_red(a, P.dv, P.dv_scratch, blk * 4 + head, bx, by, b, head * 64);
        tmem_ld_wait();
        stage_red(c, P.dv, P.dv_scratch, blk * 4 + head, bx, by, b, head * 64 + 32);
      }
      pz = blk * 4 + head; phead = head; pbx = bx; pby = by; pb = b;
      if (AB_DK_EARLY) readout_dk(pz, phead, pbx, pby, pb, it + 1);
    }
    if (!AB_DK_EARLY && n_it > 0) readout_dk(pz, phead, pbx, pby, pb, AB_TRACE_ITERS);
    // relative-position gradient partial of this CTA
    const int rkey = row0 + lane;
    if (rkey < AT_NK) {
      float4* dst = reinterpret_cast<float4*>(P.rel_part + (long long)blockIdx.x * AB_REL_PART + rkey * 64);
#pragma unroll
      for (int j = 0; j < 16; ++j) dst[j] = make_float4(racc[4 * j], racc[4 * j + 1], racc[4 * j + 2], racc[4 * j + 3]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// zero the dK / dV outputs (any strided NHWC view, 16 bytes per thread): the direct mode accumulates into them
__global__ void attn_bwd_zero_kernel(View dk, View dv, int B, int H, int W, int C) {
  const int cv = C / 8;
  const long long per = (long long)B * H * W * cv, total = 2 * per;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const View& v = i < per ? dk : dv;
    long long r = i < per ? i : i - per;
    const int c = (int)(r % cv) * 8;
    r /= cv;
    const int x = (int)(r % W);
    r /= W;
    const int y = (int)(r % H), b = (int)(r / H);
    *reinterpret_cast<uint4*>((bf16*)v.ptr + view_off(v, b, y + v.oy, x + v.ox) + c) = make_uint4(0, 0, 0, 0);
  }
}

// sum the (up to 4) window-major contributions of every key pixel; one thread per (pixel, head, 8 channels)
__global__ void attn_bwd_fold_kernel(const bf16* __restrict__ dks, const bf16* __restrict__ dvs, View dk, View dv, int B,
                                     int H, int W, int nby, int nbx) {
  long long total = (long long)B * H * W * 32;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int cg = (int)(i & 7), head = (int)((i >> 3) & 3);
    long long p = i >> 5;
    const int x = (int)(p % W);
    p /= W;
    const int y = (int)(p % H), b = (int)(p / H);
    float ak[8], av[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) ak[j] = av[j] = 0.f;
    const int by0 = y >> 3, bx0 = x >> 3;
    for (int by = by0 - 1; by <= by0 + 1; ++by) {
      const int wy = y - (by * 8 - 3);
      if (by < 0 || by >= nby || wy < 0 || wy >= 14) continue;
      for (int bx = bx0 - 1; bx <= bx0 + 1; ++bx) {
        const int wx = x - (bx * 8 - 3);
        if (bx < 0 || bx >= nbx || wx < 0 || wx >= 14) continue;
        const long long off = (((((long long)b * nby + by) * nbx + bx) * 4 + head) * AT_NK + wy * 14 + wx) * 64 + cg * 8;
        const uint4 uk = *reinterpret_cast<const uint4*>(dks + off);
        const uint4 uv = *reinterpret_cast<const uint4*>(dvs + off);
        const __nv_bfloat162* hk = reinterpret_cast<const __nv_bfloat162*>(&uk);
        const __nv_bfloat162* hv = reinterpret_cast<const __nv_bfloat162*>(&uv);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 fk = __bfloat1622float2(hk[j]), fv = __bfloat1622float2(hv[j]);
          ak[2 * j] += fk.x; ak[2 * j + 1] += fk.y;
          av[2 * j] += fv.x; av[2 * j + 1] += fv.y;
        }
      }
    }
    uint4 ok, ov;
    __nv_bfloat162* hk = reinterpret_cast<__nv_bfloat162*>(&ok);
    __nv_bfloat162* hv = reinterpret_cast<__nv_bfloat162*>(&ov);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      hk[j] = __floats2bfloat162_rn(ak[2 * j], ak[2 * j + 1]);
      hv[j] = __floats2bfloat162_rn(av[2 * j], av[2 * j + 1]);
    }
    *reinterpret_cast<uint4*>((bf16*)dk.ptr + view_off(dk, b, y, x) + head * 64 + cg * 8) = ok;
    *reinterpret_cast<uint4*>((bf16*)dv.ptr + view_off(dv, b, y, x) + head * 64 + cg * 8) = ov;
  }
}

// d rel_h[r][j] = sum_parts sum_c dREL[(r,c)][j] (j < 32);  d rel_w[c][j] = sum_parts sum_r dREL[(r,c)][32 + j].
// Two small launches, fixed summation order (deterministic): (1) one block per key sums the CTA partials -- every
// (part, key) row is one coalesced 256-byte read, four part-groups per block combined through shared memory; (2) one
// thread per output adds its 14 keys.
__global__ void __launch_bounds__(256) attn_bwd_rel_keysum_kernel(const float* __restrict__ part, int nparts, float* __restrict__ keysum) {
  const int key = blockIdx.x, ch = threadIdx.x & 63, grp = threadIdx.x >> 6;
  float s = 0.f;
  for (int p = grp; p < nparts; p += 4) s += part[(long long)p * AB_REL_PART + key * 64 + ch];
  __shared__ float red[256];
  red[threadIdx.x] = s;
  __syncthreads();
  if (grp == 0) keysum[key * 64 + ch] = (red[ch] + red[64 + ch]) + (red[128 + ch] + red[192 + ch]);
}
__global__ void attn_bwd_rel_reduce_kernel(const float* __restrict__ keysum, float* __restrict__ d_rel_h, float* __restrict__ d_rel_w) {
  const int o = blockIdx.x * blockDim.x + threadIdx.x;     // 0..447: rel_h[r][j], 448..895: rel_w[c][j]
  if (o >= 896) return;
  const bool is_w = o >= 448;
  const int rc = (is_w ? o - 448 : o) / 32, j = (is_w ? o - 448 : o) % 32;
  float t = 0.f;
  for (int u = 0; u < 14; ++u) {
    const int key = is_w ? u * 14 + rc : rc * 14 + u;
    t += keysum[key * 64 + (is_w ? 32 + j : j)];
  }
  if (is_w) d_rel_w[o - 448] = t;
  else d_rel_h[o] = t;
}

static int at_grid(int nblocks) {
  const int sms = sm_count();
  return nblocks < sms ? nblocks : sms;
}
static std::atomic<int> g_attn_bwd_direct{1};
void set_attn_bwd_direct(int v) { g_attn_bwd_direct.store(v ? 1 : 0, std::memory_order_relaxed); }

// workspace: [grid][196][64] fp32 relative-position partials | (scratch mode) window-major bf16 dK, dV
size_t attn_bwd_tc_ws_bytes(const pht_attn_args& f) {
  if (f.heads != 4 || f.head_dim != 64 || f.block != 8 || f.halo != 3) return 0;
  size_t nblk = (size_t)f.B * (f.H / 8) * (f.W / 8);
  size_t scratch = g_attn_bwd_direct.load(std::memory_order_relaxed) ? 0 : 2 * nblk * 4 * AT_NK * 64 * sizeof(bf16);
  return (size_t)at_grid((int)nblk) * AB_REL_PART * sizeof(float) + scratch + 256;
}

static bool attn_bwd_tc_eligible(const pht_attn_bwd_args* a) {
  const pht_attn_args& f = a->fwd;
  if (f.dtype != PHT_BF16 || f.heads != 4 || f.head_dim != 64 || f.block != 8 || f.halo != 3) return false;
  if (f.H % 8 || f.W % 8 || !f.lse) return false;
  if (!at_view_ok(f.q, 256) || !at_view_ok(f.k, 256) || !at_view_ok(f.v, 256) || !at_view_ok(a->d_out, 256)) return false;
  if (!at_view_ok(a->dq, 256) || !at_view_ok(a->dk, 256) || !at_view_ok(a->dv, 256)) return false;
  if (!a->workspace || ((uintptr_t)a->workspace & 255) || a->workspace_bytes < attn_bwd_tc_ws_bytes(f)) return false;
  return get_encode_fn() != nullptr;
}

// dk / dv := 0: in the direct mode the backward kernel accumulates into them (nothing to do in the scratch mode)
int attn_bwd_zero_tc(const pht_attn_bwd_args* a, cudaStream_t st, bool* handled) {
  *handled = false;
  if (!attn_bwd_tc_eligible(a)) return PHT_OK;
  *handled = true;
  if (!g_attn_bwd_direct.load(std::memory_order_relaxed)) return PHT_OK;
  const pht_attn_args& f = a->fwd;
  const long long items = 2ll * f.B * f.H * f.W * 32;
  int grid = (int)((items + 255) / 256);
  if (grid > sm_count() * 16) grid = sm_count() * 16;
  attn_bwd_zero_kernel<<<grid, 256, 0, st>>>(make_view(a->dk), make_view(a->dv), f.B, f.H, f.W, 256);
  PHT_LAUNCH_CHECK();
  count_launch(CNT_OTHER);
  return PHT_OK;
}

int attn_bwd_tc(const pht_attn_bwd_args* a, cudaStream_t st, bool* handled) {
  *handled = false;
  const pht_attn_args& f = a->fwd;
  if (!attn_bwd_tc_eligible(a)) return PHT_OK;
  if (!a->prezeroed) {
    bool z = false;
    int rc = attn_bwd_zero_tc(a, st, &z);
    if (rc) return rc;
  }
  CUtensorMap tmQ, tmK, tmV, tmDO;
  int rc = at_tmap(&tmQ, f.q, f.B, 8, 8);
  if (rc) return rc;
  rc = at_tmap(&tmK, f.k, f.B, 14, 14);
  if (rc) return rc;
  rc = at_tmap(&tmV, f.v, f.B, 14, 14);
  if (rc) return rc;
  rc = at_tmap(&tmDO, a->d_out, f.B, 8, 8);
  if (rc) return rc;
  AbP P;
  P.B = f.B; P.H = f.H; P.W = f.W; P.nbx = f.W / 8; P.nby = f.H / 8; P.nblocks = f.B * P.nbx * P.nby;
  P.dq = make_view(a->dq);
  P.rel_h = f.rel_h; P.rel_w = f.rel_w; P.lse = f.lse;
  P.trace = g_attn_trace_on == 1;
  P.direct = g_attn_bwd_direct.load(std::memory_order_relaxed);
  P.dk = make_view(a->dk); P.dv = make_view(a->dv);
  const int grid = at_grid(P.nblocks);
  P.rel_part = (float*)a->workspace;
  const size_t scratch = (size_t)P.nblocks * 4 * AT_NK * 64;
  P.dk_scratch = (bf16*)(P.rel_part + (size_t)grid * AB_REL_PART);
  P.dv_scratch = P.dk_scratch + scratch;
  PHT_SMEM_ATTR_ONCE(attn_bwd_tc_kernel, AB_SMEM);
  PHT_CUDA(launch_pdl(attn_bwd_tc_kernel, dim3(grid), dim3(AB_THREADS), AB_SMEM, st, tmQ, tmK, tmV, tmDO, P));
  PHT_LAUNCH_CHECK();
  if (!P.direct) {
    long long items = (long long)f.B * f.H * f.W * 32;
    int fgrid = (int)((items + 255) / 256);
    if (fgrid > sm_count() * 16) fgrid = sm_count() * 16;
    attn_bwd_fold_kernel<<<fgrid, 256, 0, st>>>(P.dk_scratch, P.dv_scratch, make_view(a->dk), make_view(a->dv), f.B, f.H, f.W,
                                               P.nby, P.nbx);
    count_launch(CNT_OTHER);
  }
  float* keysum = (float*)stream_scratch(st, 2, (size_t)AB_REL_PART * sizeof(float));
  if (!keysum) return PHT_ERR_CUDA;
  attn_bwd_rel_keysum_kernel<<<AT_NK, 256, 0, st>>>(P.rel_part, grid, keysum);
  attn_bwd_rel_reduce_kernel<<<7, 128, 0, st>>>(keysum, a->d_rel_h, a->d_rel_w);
  PHT_LAUNCH_CHECK();
  walk_dir_set(a->dq.ptr, 0);
  walk_dir_set(a->dk.ptr, 0);
  walk_dir_set(a->dv.ptr, 0);
  count_launch(CNT_ATTN_TC);
  count_launch(CNT_OTHER, 2);
  *handled = true;
  return PHT_OK;
}

}  // namespace pht
