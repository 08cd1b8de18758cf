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
//   warps 2..5  epilogue (the "wide" config adds warps 7..10: two warps per TMEM quadrant take alternate 64-channel
//               chunks, each group with its own staging tile, epilogue-input slot and named barrier): tcgen05.ld the accumulator (each thread owns one output pixel = one TMEM lane), fused
//               bias / residual / (leaky)ReLU / activation-derivative mask, bf16 pack into a 128B-swizzled staging
//               tile [128 px][64 ch] and one TMA tensor store per 64-channel chunk (full-line writes, ragged tiles
//               clipped by the tensor map).  Runs concurrently with the MMAs of the next tile (double-buffered TMEM).
#include <atomic>

#include "tc_common.cuh"

namespace pht {

using namespace tc;

constexpr int TILE_H = 8, TILE_W = 16, TILE_M = TILE_H * TILE_W;  // 128 output pixels per tile
constexpr int BK = 64;                                            // bf16 elements per 128-byte swizzle row
constexpr int TC_THREADS = 224;                                   // 7 warps: TMA, MMA, 4 x epilogue, aux TMA (+ 4 x epilogue, MODE 1)
constexpr int MAX_VEC_N = 768;                                     // largest N (bias / slope vectors)
constexpr int STG_BYTES = TILE_M * 128;  // one epilogue tile: 128 pixels x 64 bf16, 128B-swizzled rows

struct TcGemmP {
  int B, Ho, Wo, N, ks, n_src;
  unsigned flags;
  int srcC[3], srcOy[3], srcOx[3], koff[3];
  int tiles_x, tiles_y, n_tiles, num_tiles;
  const float* bias;
  const float* slope;
  const float* mslope;
  View out1;          // used when out1_f32 and for the frame stores (ring)
  View out2;          // frame stores (ring)
  int ring;           // bit 0 / 1: out1 / out2 is the interior of a padded buffer whose 1-pixel frame is written too; bit 2: reflect
  int has_out1, has_out2, has_resid, has_mask;
  int out1_f32;       // out1 is fp32 (direct stores; used by the 64-wide decoder-tail GEMM only)
  int num_pairs;      // CTA-pair kernels: work units = (pair of pixel tiles, output tile)
  int rev;            // walk the tiles last-to-first
  int trace;          // diagnostics: CTA 0 records clock64 stamps of its pipeline events per tile
  int strip_rows;     // > 0: the last 2 output rows are covered by 2 x 64-pixel strip tiles (pad-fold domain, see conv_gemm_tc)
  int n_strip, reg_tiles_y;
  int residOy, residOx, maskOy, maskOx, out1Oy, out1Ox, out2Oy, out2Ox;
};

// MODE 0 "deep":     4-stage operand ring, one output staging tile, no epilogue inputs: K-heavy 3x3 convolutions.
// MODE 1 "wide":     3-stage ring, TWO epilogue warp groups (their short K loop makes the 1x1 GEMMs epilogue-bound), each
//                    with its own output staging tile and TMA-prefetched epilogue-input slot (residual / mask).
// MODE 2 "deep+aux": 4-stage ring, one staging tile, one epilogue-input slot: 3x3 convolutions with a fused
//                    residual / mask / second output / padding fold (their long K loop hides the serial epilogue);
//                    slope vectors are read through the L1 instead of shared memory to fit in 227 KB.
// CG = 2 (BN = 256 only): the kernel runs as CTA pairs (2-CTA clusters, one TPC).  A pair computes two 128-pixel tiles
//                    against the SAME 256 output channels with one M = 256 tcgen05.mma.cta_group::2 per K step: each
//                    CTA loads its own activation tile and only HALF of the weight tile, so a stage is 32 KB instead of
//                    48 KB, the rings are 6 / 4 deep instead of 4 / 3 and the weights cross L2 -> SM once per pair.
// BKS = 32 (BN = 256, CG = 1, MODE 0 / 2): the operand ring is cut into 32-deep half stages (24 KB, 64-byte swizzle
//                    rows), 8 of them.  The K loop of the 3x3 convolutions is bound by the refill latency of the ring
//                    (measured: tile period independent of the number of CTAs, 586 clk per 64-deep step with 4 stages,
//                    703 with 3): with half stages 7 x 24 KB instead of 3 x 48 KB are in flight behind the one being
//                    consumed.
template <int BN, int MODE, int CG = 1, int BKS_ = BK> struct TcCfg {
  static constexpr int BKS = BKS_;                                 // K depth of one ring stage
  static constexpr int A_BYTES = TILE_M * BKS * 2;                 // 16 KB
  static constexpr int B_BYTES = BN / CG * BKS * 2;                // 32 KB @ BN=256 (16 KB per CTA of a pair)
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = BKS == 32 ? 8 : (BN == 256 ? (CG == 2 ? (MODE == 1 ? 4 : 6) : (MODE == 1 ? 3 : 4)) : (BN == 128 ? 4 : 6));
  static constexpr int NSTG = MODE == 1 ? 2 : 1;
  static constexpr int PARTS = MODE == 1 ? 2 : 1;                  // epilogue warp groups (4 warps = 128 rows each)
  static constexpr int THREADS = PARTS == 2 ? TC_THREADS + 128 : TC_THREADS;
  static constexpr int AUX_SLOTS = MODE == 1 ? 2 : (MODE == 2 ? 1 : 0);
  static constexpr bool VEC_SMEM = MODE != 2;                      // slope / mslope staged in shared memory
  static constexpr int VEC_N = VEC_SMEM ? MAX_VEC_N : 256;         // largest N this config accepts
  static constexpr int VEC_BYTES = (VEC_SMEM ? 3 : 1) * VEC_N * 4;
  static constexpr int TMEM_COLS = 2 * BN < 32 ? 32 : 2 * BN;      // power of two for BN in {64,128,256}
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + (NSTG + AUX_SLOTS) * STG_BYTES + VEC_BYTES + 256 + 1024;
};
static_assert(TcCfg<256, 0>::SMEM_BYTES <= 232448 && TcCfg<256, 1>::SMEM_BYTES <= 232448 && TcCfg<256, 2>::SMEM_BYTES <= 232448,
              "conv_gemm_tc: shared memory budget");
static_assert(TcCfg<256, 0, 1, 32>::SMEM_BYTES <= 232448 && TcCfg<256, 2, 1, 32>::SMEM_BYTES <= 232448, "conv_gemm_tc (half stages)");
static_assert(TcCfg<256, 0, 2>::SMEM_BYTES <= 232448 && TcCfg<256, 1, 2>::SMEM_BYTES <= 232448 &&
              TcCfg<256, 2, 2>::SMEM_BYTES <= 232448, "conv_gemm_tc (CTA pairs): shared memory budget");

constexpr int CG_TRACE_TILES = 32, CG_TRACE_EVENTS = 8;
constexpr int CG_TRACE_CTAS = 160;   // after the tile events: (start ns, end ns, SM id, entry ns) of every CTA
constexpr int CG_TRACE_CTA_F = 4;
__device__ long long g_conv_trace[CG_TRACE_TILES * CG_TRACE_EVENTS + CG_TRACE_CTA_F * CG_TRACE_CTAS];
static int g_conv_trace_on = 0;
void set_conv_trace(int v) { g_conv_trace_on = v; }
int read_conv_trace(long long* host, int n) {
  if (n > CG_TRACE_TILES * CG_TRACE_EVENTS + CG_TRACE_CTA_F * CG_TRACE_CTAS) n = CG_TRACE_TILES * CG_TRACE_EVENTS + CG_TRACE_CTA_F * CG_TRACE_CTAS;
  PHT_CUDA(cudaDeviceSynchronize());
  PHT_CUDA(cudaMemcpyFromSymbol(host, g_conv_trace, (size_t)n * sizeof(long long)));
  return n;
}

// tile index -> image, tile origin, shape.  Regular tiles are 8 x 16 pixels; with P.strip_rows the domain's last two
// rows are covered by 2 x 64 strips (a 130-row padded domain needs 16 tile rows + 3 strips per image instead of 17 x 9
// tiles: 147 instead of 153 tiles per image, i.e. 8 instead of 9 waves of 148 CTAs at B = 8)
struct TileXY {
  int b, x0, y0, nt, strip;
};
// CG == 2: `tile` is the pair's work unit and `rank` the CTA's rank in the pair; the pair takes pixel tiles 2m and 2m+1
// of ONE output tile (they share the weights).  A pixel tile past the end decodes to b == P.B: its loads are zero-filled
// and its stores clipped by the TMA unit.
template <int CG>
__device__ __forceinline__ TileXY decode_tile(const TcGemmP& P, int tile, int rank) {
  TileXY t;
  int mt;
  if (CG == 2) {
    const int u = P.rev ? P.num_pairs - 1 - tile : tile;
    t.nt = u % P.n_tiles;
    mt = (u / P.n_tiles) * 2 + rank;
  } else {
    const int te = P.rev ? P.num_tiles - 1 - tile : tile;   // serpentine launch order (see conv_gemm_tc)
    t.nt = te % P.n_tiles;
    mt = te / P.n_tiles;
  }
  const int reg = P.tiles_x * P.reg_tiles_y, per_img = reg + P.n_strip;
  t.b = mt / per_img;
  const int r = mt - t.b * per_img;
  if (r < reg) {
    const int ty = r / P.tiles_x;
    t.strip = 0; t.x0 = (r - ty * P.tiles_x) * TILE_W; t.y0 = ty * TILE_H;
  } else {
    t.strip = 1; t.x0 = (r - reg) * 64; t.y0 = P.Ho - 2;
  }
  return t;
}

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
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* tm, const void* src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}

template <int BN, int MODE, int CG, int BKS>
__global__ void __launch_bounds__((TcCfg<BN, MODE, CG, BKS>::THREADS), 1)
conv_gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                    const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmW,
                    const __grid_constant__ CUtensorMap tmO1, const __grid_constant__ CUtensorMap tmO2,
                    const __grid_constant__ CUtensorMap tmR, const __grid_constant__ CUtensorMap tmM,
                    const __grid_constant__ CUtensorMap tmA0s, const __grid_constant__ CUtensorMap tmO1s,
                    const __grid_constant__ CUtensorMap tmO2s, const __grid_constant__ CUtensorMap tmRs,
                    const __grid_constant__ CUtensorMap tmMs, const TcGemmP P) {
  using Cfg = TcCfg<BN, MODE, CG, BKS>;
  constexpr int NCHUNK = BN / 64;
  constexpr bool AUX = Cfg::AUX_SLOTS > 0;
  constexpr int PARTS = Cfg::PARTS;
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B needs 1024-byte aligned stage bases
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // keeps the .shared address space
  uint8_t* stage_base = smem;
  uint8_t* stg = smem + Cfg::STAGES * Cfg::STAGE_BYTES;   // 1024-aligned (stage sizes are multiples of 1024)
  uint8_t* aux = stg + Cfg::NSTG * STG_BYTES;
  float* s_bias = reinterpret_cast<float*>(aux + Cfg::AUX_SLOTS * STG_BYTES);
  float* s_slope = s_bias + MAX_VEC_N;      // only when Cfg::VEC_SMEM
  float* s_mslope = s_slope + MAX_VEC_N;
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(s_bias) + Cfg::VEC_BYTES);
  uint64_t* full_bar = bars;                      // [STAGES]
  uint64_t* empty_bar = bars + Cfg::STAGES;       // [STAGES]
  uint64_t* tfull_bar = bars + 2 * Cfg::STAGES;   // [2]
  uint64_t* tempty_bar = tfull_bar + 2;           // [2]
  uint64_t* afull_bar = tempty_bar + 2;           // [2] epilogue-input slot filled (TMA tx)
  uint64_t* aempty_bar = afull_bar + 2;           // [2] epilogue-input slot consumed (128 arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(aempty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // work units: tiles, or (CTA pairs) pairs of pixel tiles
  const int cta_rank = CG == 2 ? (int)cluster_ctarank() : 0;
  const int unit0 = CG == 2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int unit_step = CG == 2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int num_units = CG == 2 ? P.num_pairs : P.num_tiles;

  long long t_entry = 0;
  if (P.trace && threadIdx.x == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_entry));
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA0);
    if (P.n_src > 1) prefetch_tmap(&tmA1);
    if (P.n_src > 2) prefetch_tmap(&tmA2);
    prefetch_tmap(&tmW);
    for (int s = 0; s < Cfg::STAGES; ++s) {
      mbar_init(&full_bar[s], 1);       // pairs: only the leader's barrier is used; it is armed for the bytes of both CTAs
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull_bar[a], 1);
      // pairs: the leader's barrier also collects one arrival per epilogue WARP of the peer
      mbar_init(&tempty_bar[a], CG == 2 ? 128 * PARTS + 4 * PARTS : 128 * PARTS);
      mbar_init(&afull_bar[a], 1);
      mbar_init(&aempty_bar[a], 128);
    }
    mbar_fence_init();
  }
  if (warp == 1) {
    if (CG == 2) tmem_alloc_2sm(tmem_slot, Cfg::TMEM_COLS);
    else tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  }
  tc_fence_before();
  if (CG == 2) cluster_sync_all();   // the peer's barriers are initialised before anything signals them
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // Everything above ran while the previous kernel of the stream was still finishing (no global memory touched).  The
  // loads start the moment it has completed; the bias / slope vectors are staged by the epilogue warps on their own
  // (they have nothing to do until the first accumulator is complete), off the producer's critical path.
  griddep_wait();
  griddep_launch();
  if (P.trace && threadIdx.x == 0 && blockIdx.x < CG_TRACE_CTAS) {
    long long t;
    uint32_t sm;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
    long long* rec = g_conv_trace + CG_TRACE_TILES * CG_TRACE_EVENTS + CG_TRACE_CTA_F * blockIdx.x;
    rec[0] = t; rec[2] = sm; rec[3] = t_entry;
  }

  const int T = P.ks * P.ks, half = P.ks / 2;
  const int n_kinds = P.has_resid + P.has_mask;   // epilogue-input tiles per 64-channel chunk
  const bool tracing = P.trace && blockIdx.x == 0 && lane == 0;
  auto stamp = [&](int it, int ev) {
    if (tracing && it < CG_TRACE_TILES) g_conv_trace[it * CG_TRACE_EVENTS + ev] = clock64();
  };

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int pit = 0;
      const uint32_t full_leader = CG == 2 ? mapa_u32(&full_bar[0], 0) : 0u;
      for (int tile = unit0; tile < num_units; tile += unit_step, ++pit) {
        const TileXY tl = decode_tile<CG>(P, tile, cta_rank);
        const int b = tl.b, x0 = tl.x0, y0 = tl.y0, n0 = tl.nt * BN;
        bool first = true;
        for (int t = 0; t < T; ++t) {
          const int dy = t / P.ks - half, dx = t % P.ks - half;
          for (int s = 0; s < P.n_src; ++s) {
            const CUtensorMap* tm = tl.strip ? &tmA0s : (s == 0 ? &tmA0 : (s == 1 ? &tmA1 : &tmA2));
            for (int kc = 0; kc < P.srcC[s]; kc += BKS) {
              mbar_wait(&empty_bar[stage], phase ^ 1);
              if (first) { stamp(pit, 0); first = false; }
              uint8_t* a_dst = stage_base + stage * Cfg::STAGE_BYTES;
              uint8_t* b_dst = a_dst + Cfg::A_BYTES;
              if (CG == 2) {
                // this CTA's pixel tile and its half of the weight tile; both signal the LEADER's barrier, which the
                // leader arms for the bytes of both CTAs (the peer's bytes may land first: the count just goes negative)
                const uint32_t fb = full_leader + stage * 8;
                if (cta_rank == 0) mbar_expect_tx(&full_bar[stage], 2 * Cfg::STAGE_BYTES);
                tma_load_4d_2sm(a_dst, tm, fb, kc, x0 + dx + P.srcOx[s], y0 + dy + P.srcOy[s], b);
                tma_load_3d_2sm(b_dst, &tmW, fb, P.koff[s] + kc, n0 + cta_rank * (BN / 2), t);
              } else {
                mbar_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
                tma_load_4d(a_dst, tm, &full_bar[stage], kc, x0 + dx + P.srcOx[s], y0 + dy + P.srcOy[s], b);
                tma_load_3d(b_dst, &tmW, &full_bar[stage], P.koff[s] + kc, n0, t);
              }
              if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
            }
          }
        }
        stamp(pit, 1);
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    if (CG == 1 || cta_rank == 0) {   // all 32 lanes run the loop (warp-converged); one elected lane issues each tcgen05 instruction
      constexpr uint32_t idesc = umma_idesc_bf16(TILE_M * CG, BN, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int kiters = 0;
      for (int s = 0; s < P.n_src; ++s) kiters += P.srcC[s] / BKS;
      kiters *= T;
      int it = 0;
      for (int tile = unit0; tile < num_units; tile += unit_step, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int ki = 0; ki < kiters; ++ki) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          if (ki == 0) stamp(it, 2);
          const uint32_t a_addr = smem_u32(stage_base + stage * Cfg::STAGE_BYTES);
          const uint64_t adesc = BKS == 32 ? umma_desc_k_sw64(a_addr) : umma_desc_k_sw128(a_addr);
          const uint64_t bdesc = BKS == 32 ? umma_desc_k_sw64(a_addr + Cfg::A_BYTES) : umma_desc_k_sw128(a_addr + Cfg::A_BYTES);
#pragma unroll
          for (int k = 0; k < BKS / 16; ++k) {
            // advance 16 bf16 = 32 bytes along K inside the swizzle row: +2 in the (>>4) address field
            if (CG == 2) umma_bf16_elect_2sm(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (ki | k) ? 1u : 0u);
            else umma_bf16_elect(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (ki | k) ? 1u : 0u);
          }
          // frees this smem stage (in both CTAs of a pair) when the MMAs above have read it
          if (CG == 2) umma_commit_elect_2sm(&empty_bar[stage]);
          else umma_commit_elect(&empty_bar[stage]);
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
        }
        // accumulator complete -> epilogue (of both CTAs)
        if (CG == 2) umma_commit_elect_2sm(&tfull_bar[acc]);
        else umma_commit_elect(&tfull_bar[acc]);
        stamp(it, 3);
      }
    }
  } else if (warp == 6) {
    // ================================ epilogue-input TMA producer ================================
    // streams the residual / mask tiles in exactly the order the epilogue consumes them:
    // for tile: for 64-channel chunk: [resid], [mask]
    if (AUX && lane == 0 && n_kinds > 0) {
      constexpr int NS = AUX ? Cfg::AUX_SLOTS : 1;
      const int fo = (P.flags & PHT_EPI_PADFOLD) ? 1 : 0;   // residual / mask live on the interior domain
      prefetch_tmap(&tmR);
      prefetch_tmap(&tmM);
      int j = 0;
      int fills[2] = {0, 0};   // PARTS == 2: slot p is a depth-1 FIFO feeding epilogue group p (chunks c with c % 2 == p)
      for (int tile = unit0; tile < num_units; tile += unit_step) {
        const TileXY tl = decode_tile<CG>(P, tile, cta_rank);
        const int b = tl.b, x0 = tl.x0, y0 = tl.y0, n0 = tl.nt * BN;
        const CUtensorMap* tmRr = tl.strip ? &tmRs : &tmR;
        const CUtensorMap* tmMm = tl.strip ? &tmMs : &tmM;
        for (int c = 0; c < NCHUNK; ++c) {
          for (int kind = P.has_resid ? 0 : 1; kind < (P.has_mask ? 2 : 1); ++kind, ++j) {
            const int slot = PARTS == 2 ? (c & 1) : j % NS;
            const int fill = PARTS == 2 ? fills[slot]++ : j / NS;
            mbar_wait(&aempty_bar[slot], (fill & 1) ^ 1);
            mbar_expect_tx(&afull_bar[slot], STG_BYTES);
            if (kind == 0) tma_load_4d(aux + slot * STG_BYTES, tmRr, &afull_bar[slot], n0 + c * 64, x0 - fo + P.residOx, y0 - fo + P.residOy, b);
            else tma_load_4d(aux + slot * STG_BYTES, tmMm, &afull_bar[slot], n0 + c * 64, x0 - fo + P.maskOx, y0 - fo + P.maskOy, b);
          }
        }
      }
    }
  } else {
    // ================================ epilogue (warps 2..5 [, 7..10]) ================================
    const int quad = warp & 3;             // TMEM lane quadrant this warp may access
    const int part = (PARTS == 2 && warp >= 7) ? 1 : 0;   // epilogue group: takes the 64-channel chunks c with c % PARTS == part
    const int row = quad * 32 + lane;      // accumulator row == pixel index inside the tile
    const bool issuer = ((warp == 2 || warp == 7) && lane == 0);
    {   // stage the per-channel vectors (all epilogue threads; nobody else reads them)
      const int eid = (warp >= 7 ? 128 + (warp - 7) * 32 : (warp - 2) * 32) + lane;
      for (int i = eid; i < P.N; i += 128 * PARTS) {
        s_bias[i] = P.bias ? P.bias[i] : 0.f;
        if (Cfg::VEC_SMEM) {
          s_slope[i] = P.slope ? P.slope[i] : 1.f;
          s_mslope[i] = P.mslope ? P.mslope[i] : 0.f;
        }
      }
      asm volatile("bar.sync 3, %0;" ::"n"(128 * PARTS) : "memory");
    }
    auto group_sync = [&]() {
      if (part == 0) asm volatile("bar.sync 1, 128;" ::: "memory");
      else asm volatile("bar.sync 2, 128;" ::: "memory");
    };
    const int rsw = row & 7;
    int j_aux = 0;                         // epilogue-input tiles consumed so far
    auto stage_and_store = [&](const float* v, const CUtensorMap* tm, int c_glob, int x0, int y0, int b) {
      uint8_t* buf = stg + part * STG_BYTES;   // one staging tile per epilogue group
      // wait until the group's previous TMA store has finished READING the tile
      if (issuer) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      group_sync();
      uint8_t* srow = buf + row * 128;
#pragma unroll
      for (int g = 0; g < 8; ++g) *reinterpret_cast<uint4*>(srow + ((g ^ rsw) * 16)) = pack8(v + g * 8);
      fence_proxy_async();
      group_sync();
      if (issuer) tma_store_4d(tm, buf, c_glob, x0, y0, b);
    };
    // v[64] (op)= the epilogue-input tile that is next in the stream; releases its slot
    auto aux_apply = [&](float* v, int c_abs, bool is_mask) {
      constexpr int NS = AUX ? Cfg::AUX_SLOTS : 1;
      const int slot = PARTS == 2 ? part : j_aux % NS;
      mbar_wait(&afull_bar[slot], (PARTS == 2 ? j_aux : j_aux / NS) & 1);
      const uint8_t* arow = aux + slot * STG_BYTES + row * 128;
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        float t[8];
        unpack8(*reinterpret_cast<const uint4*>(arow + ((g ^ rsw) * 16)), t);
        if (is_mask) {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            v[g * 8 + j] *= (t[j] > 0.f ? 1.f : (Cfg::VEC_SMEM ? s_mslope[c_abs + g * 8 + j]
                                                               : (P.mslope ? __ldg(P.mslope + c_abs + g * 8 + j) : 0.f)));
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) v[g * 8 + j] += t[j];
        }
      }
      mbar_arrive(&aempty_bar[slot]);
      ++j_aux;
    };
    // frame (padding ring) of a padded output buffer: every frame pixel is a copy of one output pixel -- the edge pixel
    // (replicate) or the pixel next to it (reflect) -- stored straight from the registers of the thread that owns it
    auto ring_store = [&](const float* v, const View& o, int c_abs, int b, int y, int x) {
      const int e0 = (P.ring & 4) ? 1 : 0;
      const int tx = x == e0 ? -1 : (x == P.Wo - 1 - e0 ? P.Wo : -2);
      const int ty = y == e0 ? -1 : (y == P.Ho - 1 - e0 ? P.Ho : -2);
      if (tx == -2 && ty == -2) return;
      bf16* base = (bf16*)o.ptr + c_abs;
      bf16* d0 = tx != -2 ? base + view_off(o, b, y + o.oy, tx + o.ox) : nullptr;
      bf16* d1 = ty != -2 ? base + view_off(o, b, ty + o.oy, x + o.ox) : nullptr;
      bf16* d2 = (tx != -2 && ty != -2) ? base + view_off(o, b, ty + o.oy, tx + o.ox) : nullptr;
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        const uint4 q = pack8(v + g * 8);
        if (d0) *reinterpret_cast<uint4*>(d0 + g * 8) = q;
        if (d1) *reinterpret_cast<uint4*>(d1 + g * 8) = q;
        if (d2) *reinterpret_cast<uint4*>(d2 + g * 8) = q;
      }
    };
    int it = 0;
    const uint32_t tempty_leader = CG == 2 ? mapa_u32(&tempty_bar[0], 0) : 0u;
    auto release_acc = [&](int acc) {
      tc_fence_before();
      if (CG == 2 && cta_rank != 0) {
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster_relaxed(tempty_leader + acc * 8);
      } else {
        mbar_arrive(&tempty_bar[acc]);
      }
    };
    for (int tile = unit0; tile < num_units; tile += unit_step, ++it) {
      const TileXY tl = decode_tile<CG>(P, tile, cta_rank);
      const int b = tl.b, x0 = tl.x0, y0 = tl.y0;
      const int tw = tl.strip ? 64 : TILE_W;               // tile width in pixels (rows of the tile are tw pixels apart)
      const int py = tl.strip ? (row >> 6) : row / TILE_W, px = tl.strip ? (row & 63) : row % TILE_W;
      const CUtensorMap* tmOut1 = tl.strip ? &tmO1s : &tmO1;
      const CUtensorMap* tmOut2 = tl.strip ? &tmO2s : &tmO2;
      const int x = x0 + px, y = y0 + py, n0 = tl.nt * BN;
      const bool valid = x < P.Wo && y < P.Ho && b < P.B;
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      if (warp == 2 || warp == 7) stamp(it, warp == 2 ? 4 : 6);
      const uint32_t t_addr = tmem_base + acc * BN + ((uint32_t)(quad * 32) << 16);
      if (part * 64 >= BN) release_acc(acc);   // this group has no chunk of the tile
#pragma unroll 1
      for (int c0 = part * 64; c0 < BN; c0 += 64 * PARTS) {
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
        if (c0 + 64 * PARTS >= BN) release_acc(acc);   // this group's share of the accumulator is in registers: release it
#pragma unroll
        for (int j = 0; j < 64; ++j) v[j] += s_bias[n0 + c0 + j];
        if (MODE != 0 && (P.flags & PHT_EPI_PADFOLD)) {
          // backward of replicate padding, fused: the tile lives on the padded domain; every border pixel's value is
          // added to the interior pixel it was replicated from (always inside the same tile, see the host check).
          // Runs per epilogue group on the group's own staging tile.
          uint8_t* ftile = stg + part * STG_BYTES;
          if (issuer) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          group_sync();
          uint8_t* srow = ftile + row * 128;
#pragma unroll
          for (int g = 0; g < 8; ++g) *reinterpret_cast<uint4*>(srow + ((g ^ rsw) * 16)) = pack8(v + g * 8);
          group_sync();
          const int ey = y == 1 ? -1 : (y == P.Ho - 2 ? 1 : 0), ex = x == 1 ? -1 : (x == P.Wo - 2 ? 1 : 0);
          const bool interior = y >= 1 && y <= P.Ho - 2 && x >= 1 && x <= P.Wo - 2;
          if (interior && (ey | ex)) {
            auto add_row = [&](int r2) {
              const uint8_t* nrow = ftile + r2 * 128;
#pragma unroll
              for (int g = 0; g < 8; ++g) {
                float t[8];
                unpack8(*reinterpret_cast<const uint4*>(nrow + ((g ^ (r2 & 7)) * 16)), t);
#pragma unroll
                for (int j = 0; j < 8; ++j) v[g * 8 + j] += t[j];
              }
            };
            if (ey) add_row(row + ey * tw);
            if (ex) add_row(row + ex);
            if (ey && ex) add_row(row + ey * tw + ex);
          }
        }
        if (AUX && (P.flags & PHT_EPI_RESID_PRE)) aux_apply(v, n0 + c0, false);
        if (P.slope) {
#pragma unroll
          for (int j = 0; j < 64; ++j)
            v[j] = v[j] > 0.f ? v[j] : v[j] * (Cfg::VEC_SMEM ? s_slope[n0 + c0 + j] : __ldg(P.slope + n0 + c0 + j));
        }
        if (P.has_out1) {
          if (P.out1_f32) {
            if (valid) {
              float* o = (float*)P.out1.ptr + view_off(P.out1, b, y + P.out1.oy, x + P.out1.ox) + n0 + c0;
#pragma unroll
              for (int j = 0; j < 64; j += 4) *reinterpret_cast<float4*>(o + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            }
          } else {
            stage_and_store(v, tmOut1, n0 + c0, x0 + P.out1Ox, y0 + P.out1Oy, b);
#ifndef PHT_NO_RING   /* (A/B builds: tools/ab_ring_build.sh) */
            if ((P.ring & 1) && valid) ring_store(v, P.out1, n0 + c0, b, y, x);
#endif
          }
        }
        if (P.has_out2) {   // (the host clears RESID_POST / MASK when there is no out2)
          if (AUX && (P.flags & PHT_EPI_RESID_POST)) aux_apply(v, n0 + c0, false);
          if (AUX && (P.flags & PHT_EPI_MASK)) aux_apply(v, n0 + c0, true);
          stage_and_store(v, tmOut2, n0 + c0, x0 + P.out2Ox, y0 + P.out2Oy, b);
#ifndef PHT_NO_RING
          if ((P.ring & 2) && valid) ring_store(v, P.out2, n0 + c0, b, y, x);
#endif
        }
      }
      if (warp == 2 || warp == 7) stamp(it, warp == 2 ? 5 : 7);
    }
    // the staging tiles may be released once the tensor stores have READ them; their global writes are complete (and
    // visible to a dependent launch) when the grid is
    if (issuer) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  }

  tc_fence_before();
  if (CG == 2) cluster_sync_all();   // the leader's MMAs read the peer's shared memory and write its TMEM until here
  else __syncthreads();
  if (P.trace && threadIdx.x == 0 && blockIdx.x < CG_TRACE_CTAS) {
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    g_conv_trace[CG_TRACE_TILES * CG_TRACE_EVENTS + CG_TRACE_CTA_F * blockIdx.x + 1] = t;
  }
  if (warp == 1) {
    tc_fence_after();
    if (CG == 2) tmem_dealloc_2sm(tmem_base, Cfg::TMEM_COLS);
    else tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
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
                   const uint32_t* box, int swizzle_bytes) {
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
                  swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
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

// bk = channels per box: 64 (128-byte swizzle rows) or 32 (64-byte swizzle rows, half-stage operand ring)
static int make_src_tmap(CUtensorMap* tm, const pht_view& v, int B, bool strip = false, int bk = BK) {
  uint64_t dims[4] = {(uint64_t)v.C, (uint64_t)v.W, (uint64_t)v.H, (uint64_t)B};
  uint64_t strides[3] = {(uint64_t)v.sx * 2, (uint64_t)v.sy * 2, (uint64_t)v.sb * 2};
  uint32_t box[4] = {(uint32_t)bk, strip ? 64u : (uint32_t)TILE_W, strip ? 2u : (uint32_t)TILE_H, 1};
  return make_tmap_bf16(tm, v.ptr, 4, dims, strides, box, bk * 2);
}

struct TcMaps {
  CUtensorMap A[3], W, O[2], R, M;
  CUtensorMap As, Os[2], Rs, Ms;   // 2 x 64-pixel strip boxes (pad-fold domain only)
};

static std::atomic<int> g_grid_cap{0};    // diagnostics: at most this many CTAs per launch (0 = one per SM)
void set_conv_grid_cap(int v) { g_grid_cap.store(v, std::memory_order_relaxed); }
static inline int conv_sms() {
  const int cap = g_grid_cap.load(std::memory_order_relaxed), sms = sm_count();
  return cap > 0 && cap < sms ? cap : sms;
}
template <int BN, int MODE, int BKS = BK>
static int launch_tc(const TcGemmP& P, const TcMaps& m, cudaStream_t st) {
  using Cfg = TcCfg<BN, MODE, 1, BKS>;
  PHT_SMEM_ATTR_ONCE((conv_gemm_tc_kernel<BN, MODE, 1, BKS>), Cfg::SMEM_BYTES);
  const int sms = conv_sms();
  int grid = P.num_tiles < sms ? P.num_tiles : sms;
  PHT_CUDA(launch_pdl(conv_gemm_tc_kernel<BN, MODE, 1, BKS>, dim3(grid), dim3(Cfg::THREADS), Cfg::SMEM_BYTES, st, m.A[0], m.A[1], m.A[2], m.W,
                      m.O[0], m.O[1], m.R, m.M, m.As, m.Os[0], m.Os[1], m.Rs, m.Ms, P));
  PHT_LAUNCH_CHECK();
  return PHT_OK;
}
// CTA-pair variant (BN = 256; m.W must have been encoded with a 128-row box)
template <int MODE>
static int launch_tc_pairs(const TcGemmP& P, const TcMaps& m, cudaStream_t st) {
  using Cfg = TcCfg<256, MODE, 2>;
  PHT_SMEM_ATTR_ONCE((conv_gemm_tc_kernel<256, MODE, 2, BK>), Cfg::SMEM_BYTES);
  const int pairs_max = conv_sms() / 2;
  const int grid = 2 * (P.num_pairs < pairs_max ? P.num_pairs : pairs_max);
  PHT_CUDA(launch_pdl_pairs(conv_gemm_tc_kernel<256, MODE, 2, BK>, dim3(grid), dim3(Cfg::THREADS), Cfg::SMEM_BYTES, st, m.A[0], m.A[1],
                            m.A[2], m.W, m.O[0], m.O[1], m.R, m.M, m.As, m.Os[0], m.Os[1], m.Rs, m.Ms, P));
  PHT_LAUNCH_CHECK();
  return PHT_OK;
}

static std::atomic<int> g_tc_cfg{0};  // 0 = auto, 1 = force "deep" where legal, 2 = force "wide"
static std::atomic<int> g_serpentine{1};
// direction (0 = first-to-last tile, 1 = last-to-first) in which a buffer was most recently written; small direct-mapped
// table keyed by the base pointer (host-side bookkeeping only; a wrong entry costs L2 hits, never correctness)
static struct { const void* ptr; int dir; } g_walk[256];
static inline unsigned walk_slot(const void* p) { return (unsigned)(((uintptr_t)p >> 8) * 2654435761u) >> 24; }
int walk_dir_get(const void* p) {
  const auto& e = g_walk[walk_slot(p)];
  return e.ptr == p ? e.dir : 0;
}
void walk_dir_set(const void* p, int dir) {
  auto& e = g_walk[walk_slot(p)];
  e.ptr = p; e.dir = dir;
}
static std::atomic<int> g_strips{1};
static std::atomic<int> g_cta_pairs{0};   // BN = 256 GEMMs as CTA pairs (tcgen05 cta_group::2)
static std::atomic<int> g_half_ring{0};   // deep configs (3x3 convolutions) on the 8 x 24 KB half-stage operand ring
void set_half_ring(int v) { g_half_ring.store(v, std::memory_order_relaxed); }
void set_cta_pairs(int v) { g_cta_pairs.store(v, std::memory_order_relaxed); }
void set_strips(int v) { g_strips.store(v, std::memory_order_relaxed); }
void set_serpentine(int v) { g_serpentine.store(v, std::memory_order_relaxed); }
void set_tc_cfg(int v) { g_tc_cfg.store(v, std::memory_order_relaxed); }

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
  unsigned flags = a->flags;
  if (!a->out2.ptr) flags &= ~(PHT_EPI_RESID_POST | PHT_EPI_MASK);   // they only shape out2
  const bool has_resid = (flags & (PHT_EPI_RESID_PRE | PHT_EPI_RESID_POST)) != 0, has_mask = (flags & PHT_EPI_MASK) != 0;
  if (has_resid && (!view_tma_ok(a->resid) || a->resid.C < a->N)) return PHT_OK;
  if (has_mask && (!view_tma_ok(a->mask) || a->mask.C < a->N)) return PHT_OK;
  const bool padfold = (flags & PHT_EPI_PADFOLD) != 0;
  if (padfold && (a->ksize != 3 || (a->Ho - 2) % TILE_H || (a->Wo - 2) % 8 || a->Ho < 10 || a->Wo < 10 ||
                  (a->out1.ptr && a->out1.dtype != PHT_BF16)))
    return PHT_OK;
  const unsigned ring_flags = flags & (PHT_EPI_RING1 | PHT_EPI_RING2);
  if (ring_flags && (padfold || a->Ho < 4 || a->Wo < 4 || ((flags & PHT_EPI_RING1) && (!a->out1.ptr || a->out1.dtype != PHT_BF16)) ||
                     ((flags & PHT_EPI_RING2) && !a->out2.ptr)))
    return PHT_OK;
  const bool out1_f32 = a->out1.ptr && a->out1.dtype == PHT_F32;
  if (out1_f32) {
    if (((uintptr_t)a->out1.ptr & 15) || a->out1.sx % 4 || a->out1.sy % 4 || a->out1.sb % 4) return PHT_OK;
  } else if (a->out1.ptr && (!view_tma_ok(a->out1) || a->out1.C < a->N)) {
    return PHT_OK;
  }
  if (a->out2.ptr && (!view_tma_ok(a->out2) || a->out2.C < a->N)) return PHT_OK;
  if (((uintptr_t)a->w & 15) != 0) return PHT_OK;
  if (!get_encode_fn()) return PHT_OK;

  // 1x1 GEMMs are HBM-bound: wide epilogue.  3x3: deep ring; with a fused epilogue (inputs, second output, fold) the
  // deep+aux variant.
  const int force = g_tc_cfg.load(std::memory_order_relaxed);
  const bool fused = has_resid || has_mask || (a->out1.ptr && a->out2.ptr) || padfold;
  int mode = a->ksize == 1 ? 1 : (fused ? 2 : 0);
  if (mode == 2 && a->N > TcCfg<256, 2>::VEC_N) mode = 1;   // (the slope vectors do not fit beside the deep ring)
  if (force == 2 && !padfold) mode = 1;
  if (force == 4 && a->ksize == 3 && fused) mode = 1;   // A/B: fused 3x3 (pad-fold included) on the two-group "wide" config
  if (force == 1 && !fused) mode = 0;
  if (force == 3 && a->ksize == 1 && a->N <= TcCfg<256, 2>::VEC_N) mode = fused ? 2 : 0;
  const bool pairs = BN == 256 && g_cta_pairs.load(std::memory_order_relaxed) != 0;
  // latency-bound K loops (the deep configs) on the half-stage ring
  const bool half = BN == 256 && !pairs && mode != 1 && g_half_ring.load(std::memory_order_relaxed) != 0;
  const int bk = half ? 32 : BK;

  TcGemmP P;
  P.B = a->B; P.Ho = a->Ho; P.Wo = a->Wo; P.N = a->N; P.ks = a->ksize; P.n_src = a->n_src; P.flags = flags;
  TcMaps m;
  int k = 0;
  for (int s = 0; s < 3; ++s) {
    if (s < a->n_src) {
      P.srcC[s] = a->src[s].C; P.srcOy[s] = a->src[s].oy; P.srcOx[s] = a->src[s].ox; P.koff[s] = k;
      k += a->src[s].C;
      int rc = make_src_tmap(&m.A[s], a->src[s], a->B, false, bk);
      if (rc) return rc;
    } else {
      P.srcC[s] = 0; P.srcOy[s] = P.srcOx[s] = 0; P.koff[s] = 0;
      m.A[s] = m.A[0];
    }
  }
  const int T = a->ksize * a->ksize;
  {
    uint64_t dims[3] = {(uint64_t)ktot, (uint64_t)a->N, (uint64_t)T};
    uint64_t strides[2] = {(uint64_t)ktot * 2, (uint64_t)ktot * a->N * 2};
    uint32_t box[3] = {(uint32_t)bk, (uint32_t)(pairs ? BN / 2 : BN), 1};
    int rc = make_tmap_bf16(&m.W, const_cast<void*>(a->w), 3, dims, strides, box, bk * 2);
    if (rc) return rc;
  }
  // Serpentine tile order: consecutive launches walk their tiles in opposite directions, so a kernel starts on the
  // part of its input that the previous kernel wrote LAST and that is still resident in the 126 MB L2
  // (the activations are 67-134 MB each; walking them in the same direction every time always misses).
  // The direction every buffer was last written in is remembered (walk_dir_*); a consumer walks opposite to its
  // producer.  Buffers written by other kernels (attention, elementwise) count as first-to-last.
  P.rev = g_serpentine.load(std::memory_order_relaxed) ? (walk_dir_get(a->src[0].ptr) ^ 1) : 0;
  if (a->out1.ptr) walk_dir_set(a->out1.ptr, P.rev);
  if (a->out2.ptr) walk_dir_set(a->out2.ptr, P.rev);
  P.tiles_x = ceil_div(a->Wo, TILE_W);
  P.tiles_y = ceil_div(a->Ho, TILE_H);
  P.n_tiles = a->N / BN;
  // pad-fold domain (H+2 rows, (H+2) % 8 == 2): the last two rows as 2 x 64 strips instead of a ninth/seventeenth
  // row of 8 x 16 tiles that would be 3/4 empty
  const bool strips = padfold && a->n_src == 1 && g_strips.load(std::memory_order_relaxed) != 0;
  P.trace = g_conv_trace_on;
  P.strip_rows = strips ? 2 : 0;
  P.reg_tiles_y = strips ? (a->Ho - 2) / TILE_H : P.tiles_y;
  P.n_strip = strips ? ceil_div(a->Wo, 64) : 0;
  P.num_tiles = a->B * (P.tiles_x * P.reg_tiles_y + P.n_strip) * P.n_tiles;
  P.num_pairs = ceil_div(P.num_tiles / P.n_tiles, 2) * P.n_tiles;
  P.bias = a->bias; P.slope = a->slope; P.mslope = a->mslope;
  P.out1 = a->out1.ptr ? make_view(a->out1) : null_view();
  P.out2 = a->out2.ptr ? make_view(a->out2) : null_view();
  P.ring = ((flags & PHT_EPI_RING1) ? 1 : 0) | ((flags & PHT_EPI_RING2) ? 2 : 0) | ((flags & PHT_EPI_RING_REFLECT) ? 4 : 0);
  P.has_out1 = a->out1.ptr ? 1 : 0; P.has_out2 = a->out2.ptr ? 1 : 0;
  P.has_resid = has_resid ? 1 : 0; P.has_mask = has_mask ? 1 : 0;
  P.out1_f32 = out1_f32 ? 1 : 0;
  P.residOy = a->resid.oy; P.residOx = a->resid.ox; P.maskOy = a->mask.oy; P.maskOx = a->mask.ox;
  P.out1Oy = a->out1.oy; P.out1Ox = a->out1.ox; P.out2Oy = a->out2.oy; P.out2Ox = a->out2.ox;
  m.O[0] = m.O[1] = m.R = m.M = m.W;
  int rc = PHT_OK;
  if (a->out1.ptr && !out1_f32) rc = make_src_tmap(&m.O[0], a->out1, a->B);
  if (!rc && a->out2.ptr) rc = make_src_tmap(&m.O[1], a->out2, a->B);
  if (!rc && has_resid) rc = make_src_tmap(&m.R, a->resid, a->B);
  if (!rc && has_mask) rc = make_src_tmap(&m.M, a->mask, a->B);
  m.As = m.Os[0] = m.Os[1] = m.Rs = m.Ms = m.W;
  if (!rc && strips) {
    rc = make_src_tmap(&m.As, a->src[0], a->B, true, bk);
    if (!rc && a->out1.ptr && !out1_f32) rc = make_src_tmap(&m.Os[0], a->out1, a->B, true);
    if (!rc && a->out2.ptr) rc = make_src_tmap(&m.Os[1], a->out2, a->B, true);
    if (!rc && has_resid) rc = make_src_tmap(&m.Rs, a->resid, a->B, true);
    if (!rc && has_mask) rc = make_src_tmap(&m.Ms, a->mask, a->B, true);
  }
  if (rc) return rc;
  if (half) rc = mode == 0 ? launch_tc<256, 0, 32>(P, m, st) : launch_tc<256, 2, 32>(P, m, st);
  else if (pairs) rc = mode == 0 ? launch_tc_pairs<0>(P, m, st) : (mode == 1 ? launch_tc_pairs<1>(P, m, st) : launch_tc_pairs<2>(P, m, st));
  else if (BN == 256) rc = mode == 0 ? launch_tc<256, 0>(P, m, st) : (mode == 1 ? launch_tc<256, 1>(P, m, st) : launch_tc<256, 2>(P, m, st));
  else if (BN == 128) rc = mode == 0 ? launch_tc<128, 0>(P, m, st) : (mode == 1 ? launch_tc<128, 1>(P, m, st) : launch_tc<128, 2>(P, m, st));
  else rc = mode == 0 ? launch_tc<64, 0>(P, m, st) : (mode == 1 ? launch_tc<64, 1>(P, m, st) : launch_tc<64, 2>(P, m, st));
  if (rc) return rc;
  count_launch(CNT_GEMM_TC);
  *handled = true;
  return PHT_OK;
}

}  // namespace pht
