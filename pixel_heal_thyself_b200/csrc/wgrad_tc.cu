// tcgen05 weight-gradient GEMM for sm_100a:
//   dw[t][n][koff_s + k] = sum_p dy[p, n] * src_s(p + tap_t)[k]
// i.e. a GEMM with M = dy channels, N = source channels, K = pixels.  Both operands are "MN-major" (the channel
// dim is contiguous in the NHWC buffers, the pixel dim is the K row), which tcgen05 consumes directly from the
// 128B-swizzled TMA boxes [64 ch x (4 x 16) pixels] -- no transposes are ever materialised.
// The bias gradient (column sums of dy) is fused: in the jobs of the centre tap / first k-tile the four otherwise idle
// epilogue warps sum the dy tiles straight out of the TMA stages while the tensor core consumes them.
//
// One CTA = one job (tap t, n-tile of 128*MH channels, k-tile of <= 256 source channels, pixel split): it streams its
// pixel range through a 3/4-stage TMA ring, accumulates in TMEM (MH accumulators of 128 x 256 fp32) and writes the
// fp32 partial tile to the split workspace; a second tiny kernel reduces the splits in a fixed order (deterministic).
#include <atomic>

#include "tc_common.cuh"

namespace pht {

using namespace tc;

constexpr int WG_PH = 4, WG_PW = 16, WG_PIX = WG_PH * WG_PW;  // 64 pixels (= K) per stage
constexpr int WG_BOX_BYTES = WG_PIX * 128;                    // [64 px][64 ch] bf16 = 8 KB
constexpr int WG_THREADS = 192;
constexpr int WG_MAX_KT = 8;

struct WgKTile {
  int src, c0, wk, koff;  // source index, channel offset inside the source, width (multiple of 64, <= 256), offset in Ktot
};

struct WgP {
  int B, Ho, Wo, N, ks, Ktot;
  int n_ktiles, n_ntiles, splits, T;
  int ptiles_x, ptiles_y, ptiles_total, ptiles_per_split;
  int dyOy, dyOx;
  int srcOy[3], srcOx[3];
  WgKTile kt[WG_MAX_KT];
  float* ws;       // [splits][T][N][Ktot]
  float* bias_ws;  // [splits][T][N] partial column sums of dy, or nullptr
};

template <int MH> struct WgCfg {
  static constexpr int A_BYTES = MH * 2 * WG_BOX_BYTES;   // MH*128 dy channels
  static constexpr int B_BYTES = 4 * WG_BOX_BYTES;        // up to 256 source channels
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;   // 48 KB (MH=1) / 64 KB (MH=2)
  static constexpr int STAGES = MH == 1 ? 4 : 3;
  static constexpr int TMEM_COLS = MH == 1 ? 256 : 512;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 256 + 1024;
};

template <int MH>
__global__ void __launch_bounds__(WG_THREADS, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmDy, const __grid_constant__ CUtensorMap tmS0,
                const __grid_constant__ CUtensorMap tmS1, const __grid_constant__ CUtensorMap tmS2, const WgP P) {
  using Cfg = WgCfg<MH>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // keeps the .shared address space
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + Cfg::STAGES;
  uint64_t* done_bar = bars + 2 * Cfg::STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done_bar + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // job decode
  int job = blockIdx.x;
  const int kti = job % P.n_ktiles; job /= P.n_ktiles;
  const int nti = job % P.n_ntiles; job /= P.n_ntiles;
  const int t = job % P.T;
  const int split = job / P.T;
  const WgKTile kt = P.kt[kti];
  const int n0 = nti * 128 * MH;
  const int half = P.ks / 2, dy = t / P.ks - half, dx = t % P.ks - half;
  const int pt_beg = split * P.ptiles_per_split;
  const int pt_end = min(pt_beg + P.ptiles_per_split, P.ptiles_total);
  const int npt = pt_end - pt_beg;
  // this job also produces a bias-gradient partial of its (split, n-tile): first k-tile only
  // (the T tap-jobs of a split stream the same dy tiles: job t sums every T-th tile, so the work is spread evenly)
  const bool do_bias = P.bias_ws != nullptr && kti == 0;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmDy);
    prefetch_tmap(&tmS0);
    for (int s = 0; s < Cfg::STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], do_bias ? 5 : 1);   // tensor core (commit) + the 4 column-sum warps
    }
    mbar_init(done_bar, 1);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  griddep_wait();   // (programmatic dependent launch: the preamble above overlapped the previous kernel's tail)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  griddep_launch();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t stage_tx = (uint32_t)(Cfg::A_BYTES + (kt.wk / 64) * WG_BOX_BYTES);

  if (warp == 0) {
    if (lane == 0) {
      const CUtensorMap* tmS = kt.src == 0 ? &tmS0 : (kt.src == 1 ? &tmS1 : &tmS2);
      const int soy = P.srcOy[kt.src], sox = P.srcOx[kt.src];
      int stage = 0;
      uint32_t phase = 0;
      for (int pt = pt_beg; pt < pt_end; ++pt) {
        const int tx = pt % P.ptiles_x, ty = (pt / P.ptiles_x) % P.ptiles_y, b = pt / (P.ptiles_x * P.ptiles_y);
        const int x0 = tx * WG_PW, y0 = ty * WG_PH;
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* a_dst = smem + stage * Cfg::STAGE_BYTES;
        uint8_t* b_dst = a_dst + Cfg::A_BYTES;
        mbar_expect_tx(&full_bar[stage], stage_tx);
#pragma unroll
        for (int j = 0; j < 2 * MH; ++j)
          tma_load_4d(a_dst + j * WG_BOX_BYTES, &tmDy, &full_bar[stage], n0 + j * 64, x0 + P.dyOx, y0 + P.dyOy, b);
        for (int j = 0; j < kt.wk / 64; ++j)
          tma_load_4d(b_dst + j * WG_BOX_BYTES, tmS, &full_bar[stage], kt.c0 + j * 64, x0 + dx + sox, y0 + dy + soy, b);
        if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(128, kt.wk, 1, 1);  // A and B are MN-major
      int stage = 0;
      uint32_t phase = 0;
      for (int i = 0; i < npt; ++i) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smem + stage * Cfg::STAGE_BYTES);
        const uint32_t b_addr = a_addr + Cfg::A_BYTES;
#pragma unroll
        for (int k = 0; k < WG_PIX / 16; ++k) {
          // K = 16 pixels = 16 rows of 128 B; 64-channel groups are WG_BOX_BYTES apart (LBO), 8-row groups 1024 B (SBO)
          const uint64_t bdesc = umma_desc_mn_sw128(b_addr + k * 2048, WG_BOX_BYTES, 1024);
#pragma unroll
          for (int h = 0; h < MH; ++h) {
            const uint64_t adesc = umma_desc_mn_sw128(a_addr + h * 2 * WG_BOX_BYTES + k * 2048, WG_BOX_BYTES, 1024);
            umma_bf16(tmem_base + h * 256, adesc, bdesc, idesc, (i | k) ? 1u : 0u);
          }
        }
        umma_commit(&empty_bar[stage]);
        if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
      }
      umma_commit(done_bar);
    }
  } else {
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    if (do_bias) {
      // column sums of the dy tiles: warp w of the 4 owns channel box (w, w+4, ..) of 64 channels, lane = channel pair;
      // a warp reads one whole 128-byte pixel row per instruction (conflict-free under the 128B swizzle)
      const int ew = warp - 2;
      float bs[MH][2];
#pragma unroll
      for (int h = 0; h < MH; ++h) bs[h][0] = bs[h][1] = 0.f;
      int stage = 0;
      uint32_t phase = 0;
      for (int i = 0; i < npt; ++i) {
        mbar_wait(&full_bar[stage], phase);
        const uint32_t a_base = smem_u32(smem + stage * Cfg::STAGE_BYTES);
        const bool mine = (pt_beg + i) % P.T == t;
#pragma unroll
        for (int h = 0; h < MH; ++h) {
          const uint32_t box = a_base + (h * 4 + ew) * WG_BOX_BYTES + ((lane & 3) << 2);   // 4 warps x MH boxes
          if (mine && h * 4 + ew < 2 * MH) {
            float e0 = 0.f, e1 = 0.f, o0 = 0.f, o1 = 0.f;   // two independent add chains per channel
#pragma unroll
            for (int r = 0; r < WG_PIX; r += 2) {
              uint32_t u, w;
              asm volatile("ld.shared.u32 %0, [%1];" : "=r"(u) : "r"(box + r * 128 + (((lane >> 2) ^ (r & 7)) << 4)));
              asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w) : "r"(box + (r + 1) * 128 + (((lane >> 2) ^ ((r + 1) & 7)) << 4)));
              e0 += __uint_as_float(u << 16);
              e1 += __uint_as_float(u & 0xffff0000u);
              o0 += __uint_as_float(w << 16);
              o1 += __uint_as_float(w & 0xffff0000u);
            }
            bs[h][0] += e0 + o0;
            bs[h][1] += e1 + o1;
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[stage]);
        if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
      }
#pragma unroll
      for (int h = 0; h < MH; ++h) {
        const int bidx = h * 4 + ew;
        if (bidx < 2 * MH) {
          float* dst = P.bias_ws + ((size_t)split * P.T + t) * P.N + n0 + bidx * 64 + lane * 2;
          dst[0] = bs[h][0];
          dst[1] = bs[h][1];
        }
      }
    }
    if (npt > 0) {
      mbar_wait(done_bar, 0);
      tc_fence_after();
    }
    // Read-out: every thread owns one accumulator row (= output channel n), but a row-per-thread store would touch 32
    // different 128-byte lines per instruction.  Each warp transposes its [32 rows][32 columns] fp32 chunk through a
    // swizzled 4 KB tile of the (now idle) operand ring and writes whole 128-byte row segments: 4 lines per instruction.
    asm volatile("bar.sync 1, 128;" ::: "memory");   // the column-sum warps are done reading the operand ring
    const uint32_t xbuf = smem_u32(smem + (warp - 2) * 4096);
    const int rsw = lane & 7;
#pragma unroll 1
    for (int h = 0; h < MH; ++h) {
      float* dst0 = P.ws + (((size_t)split * P.T + t) * P.N + n0 + h * 128 + quad * 32) * P.Ktot + kt.koff;
      const uint32_t t_addr = tmem_base + h * 256 + ((uint32_t)(quad * 32) << 16);
      uint32_t r[2][32];
      if (npt > 0) tmem_ld32(t_addr, r[0]);
#pragma unroll 1
      for (int c0 = 0; c0 < kt.wk; c0 += 64) {
#pragma unroll
        for (int u = 0; u < 2; ++u) {          // two 32-column chunks per trip (compile-time register buffers)
          const int c = c0 + u * 32;
          if (npt > 0) {
            tmem_ld_wait();
            if (c + 32 < kt.wk) tmem_ld32(t_addr + c + 32, r[u ^ 1]);   // next chunk in flight
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) r[u][j] = 0u;
          }
#pragma unroll
          for (int j = 0; j < 8; ++j)
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(xbuf + lane * 128 + ((j ^ rsw) << 4)),
                         "r"(r[u][4 * j]), "r"(r[u][4 * j + 1]), "r"(r[u][4 * j + 2]), "r"(r[u][4 * j + 3]) : "memory");
          __syncwarp();
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int rr = i * 4 + (lane >> 3);
            uint4 v;
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                         : "r"(xbuf + rr * 128 + (((lane & 7) ^ (rr & 7)) << 4)));
            *reinterpret_cast<uint4*>(dst0 + (size_t)rr * P.Ktot + c + (lane & 7) * 4) = v;
          }
          __syncwarp();
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// fixed-order (deterministic) sum of the split partials; the LAST block reduces the bias partials [splits][N]
__global__ void wgrad_reduce_splits_kernel(const float4* __restrict__ ws, float4* __restrict__ dw, long long n4, int splits,
                                           const float* __restrict__ bias_ws, int bias_rows, float* __restrict__ dbias, int N) {
  if (bias_ws != nullptr && blockIdx.x == gridDim.x - 1) {
    for (int n = threadIdx.x; n < N; n += blockDim.x) {
      float a = 0.f;
      for (int s = 0; s < bias_rows; ++s) a += bias_ws[(long long)s * N + n];
      dbias[n] = a;
    }
    return;
  }
  const long long nblk = bias_ws != nullptr ? gridDim.x - 1 : gridDim.x;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += nblk * blockDim.x) {
    float4 a = ws[i];
    for (int s = 1; s < splits; ++s) {
      float4 b = ws[(long long)s * n4 + i];
      a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    }
    dw[i] = a;
  }
}

static bool wg_view_ok(const pht_view& v) {
  if (!v.ptr || v.dtype != PHT_BF16) return false;
  if (v.C % 64 != 0) return false;
  if (((uintptr_t)v.ptr & 15) != 0) return false;
  if ((v.sx * 2) % 16 || (v.sy * 2) % 16 || (v.sb * 2) % 16) return false;
  return v.sx > 0 && v.sy > 0 && v.sb > 0;
}

static std::atomic<int> g_wgrad_split_div{2};   // measured: 668.7 -> 675.6 patches/s at prod (4: 642.8)
void set_wgrad_split_div(int v) { g_wgrad_split_div.store(v < 1 ? 1 : v, std::memory_order_relaxed); }

struct WgPlan {
  bool ok;
  int MH, n_ntiles, n_ktiles, splits, ptiles_x, ptiles_y, ptiles_total, per_split, T, Ktot;
  WgKTile kt[WG_MAX_KT];
};

static WgPlan wg_plan(const pht_wgrad_args* a, size_t ws_limit_bytes) {
  WgPlan p;
  p.ok = false;
  if (a->dtype != PHT_BF16 || a->N % 128 != 0) return p;
  if (!wg_view_ok(a->dy) || a->dy.C < a->N) return p;
  int k = 0, nk = 0;
  for (int s = 0; s < a->n_src; ++s) {
    if (!wg_view_ok(a->src[s])) return p;
    for (int c0 = 0; c0 < a->src[s].C; c0 += 256) {
      if (nk >= WG_MAX_KT) return p;
      int wk = a->src[s].C - c0 < 256 ? a->src[s].C - c0 : 256;
      p.kt[nk].src = s; p.kt[nk].c0 = c0; p.kt[nk].wk = wk; p.kt[nk].koff = k + c0;
      ++nk;
    }
    k += a->src[s].C;
  }
  p.Ktot = k;
  if (p.Ktot % 4 != 0) return p;
  p.n_ktiles = nk;
  p.MH = a->N % 256 == 0 ? 2 : 1;
  p.n_ntiles = a->N / (128 * p.MH);
  p.T = a->ksize * a->ksize;
  p.ptiles_x = ceil_div(a->Wo, WG_PW);
  p.ptiles_y = ceil_div(a->Ho, WG_PH);
  p.ptiles_total = a->B * p.ptiles_x * p.ptiles_y;
  int base_jobs = p.T * p.n_ntiles * p.n_ktiles;
  int sms = sm_count();
  // one CTA per SM (192 KB of smem): never spill a few jobs into a second wave
  int splits = base_jobs >= sms ? 1 : sms / base_jobs;
  // HBM-bound 1x1 jobs: fewer, longer CTAs move the same bytes but write (and later reduce) fewer fp32 partials
  const int div1 = g_wgrad_split_div.load(std::memory_order_relaxed);
  if (a->ksize == 1 && div1 > 1 && splits >= 2 * div1) splits /= div1;
  if (splits > p.ptiles_total) splits = p.ptiles_total;
  size_t per_split_bytes = ((size_t)p.T * a->N * p.Ktot + (size_t)p.T * a->N) * sizeof(float);   // + T bias partial rows
  if (ws_limit_bytes > 0) {
    size_t max_splits = ws_limit_bytes / per_split_bytes;
    if (max_splits < 1) return p;
    if ((size_t)splits > max_splits) splits = (int)max_splits;
  }
  if (splits < 1) splits = 1;
  p.per_split = (p.ptiles_total + splits - 1) / splits;
  p.splits = (p.ptiles_total + p.per_split - 1) / p.per_split;
  p.ok = get_encode_fn() != nullptr;
  return p;
}

size_t wgrad_tc_workspace_bytes(const pht_wgrad_args* a) {
  WgPlan p = wg_plan(a, 0);
  if (!p.ok) return 0;
  return (size_t)p.splits * ((size_t)p.T * a->N * p.Ktot + (size_t)p.T * a->N) * sizeof(float);
}

static int wg_tmap(CUtensorMap* tm, const pht_view& v, int B) {
  uint64_t dims[4] = {(uint64_t)v.C, (uint64_t)v.W, (uint64_t)v.H, (uint64_t)B};
  uint64_t strides[3] = {(uint64_t)v.sx * 2, (uint64_t)v.sy * 2, (uint64_t)v.sb * 2};
  uint32_t box[4] = {64, WG_PW, WG_PH, 1};
  return make_tmap_bf16(tm, v.ptr, 4, dims, strides, box);
}

// every pending reduction of a bucket in one launch: blockIdx.y = job, blockIdx.x strides over the job's float4s
// (the last x-block of a job reduces its bias partials); fixed summation order => deterministic
__global__ void wgrad_reduce_batched_kernel(const pht_wgrad_reduce_job* __restrict__ jobs) {
  const pht_wgrad_reduce_job j = jobs[blockIdx.y];
  if (j.bias_partials != nullptr && blockIdx.x == gridDim.x - 1) {
    for (int n = threadIdx.x; n < j.N; n += blockDim.x) {
      float a = 0.f;
      for (int s = 0; s < j.bias_rows; ++s) a += j.bias_partials[(long long)s * j.N + n];
      j.dbias[n] = a;
    }
    return;
  }
  const long long n4 = j.elems >> 2;
  const float4* ws = reinterpret_cast<const float4*>(j.partials);
  float4* dw = reinterpret_cast<float4*>(j.dw);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)(gridDim.x - 1) * blockDim.x) {
    float4 a = ws[i];
    for (int s = 1; s < j.splits; ++s) {
      float4 b = ws[(long long)s * n4 + i];
      a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    }
    dw[i] = a;
  }
}

int wgrad_reduce_batched(const pht_wgrad_reduce_job* jobs, int n, void* table_dev, size_t table_bytes, int upload, cudaStream_t st) {
  PHT_CHECK_ARG(jobs && n > 0 && table_dev && table_bytes >= (size_t)n * sizeof(pht_wgrad_reduce_job), "wgrad_reduce_batched: bad args");
  for (int i = 0; i < n; ++i)
    PHT_CHECK_ARG(jobs[i].partials && jobs[i].dw && jobs[i].elems % 4 == 0 && jobs[i].splits >= 1 &&
                  !((uintptr_t)jobs[i].partials & 15) && !((uintptr_t)jobs[i].dw & 15), "wgrad_reduce_batched: bad job");
  if (upload) PHT_CUDA(cudaMemcpyAsync(table_dev, jobs, (size_t)n * sizeof(pht_wgrad_reduce_job), cudaMemcpyHostToDevice, st));
  int gx = 4 * sm_count() / n;                        // ~4 CTAs per SM over all jobs
  if (gx < 16) gx = 16;
  dim3 grid(gx + 1, n);                                 // + each job's bias block
  wgrad_reduce_batched_kernel<<<grid, 256, 0, st>>>((const pht_wgrad_reduce_job*)table_dev);
  PHT_LAUNCH_CHECK();
  count_launch(CNT_OTHER);
  return PHT_OK;
}

int wgrad_tc(const pht_wgrad_args* a, cudaStream_t st, bool* handled, pht_wgrad_reduce_job* defer) {
  *handled = false;
  if (!a->workspace || ((uintptr_t)a->workspace & 15) || ((uintptr_t)a->dw & 15)) return PHT_OK;
  WgPlan p = wg_plan(a, a->workspace_bytes);
  if (!p.ok) return PHT_OK;
  WgP P;
  P.B = a->B; P.Ho = a->Ho; P.Wo = a->Wo; P.N = a->N; P.ks = a->ksize; P.Ktot = p.Ktot;
  P.n_ktiles = p.n_ktiles; P.n_ntiles = p.n_ntiles; P.splits = p.splits; P.T = p.T;
  P.ptiles_x = p.ptiles_x; P.ptiles_y = p.ptiles_y; P.ptiles_total = p.ptiles_total; P.ptiles_per_split = p.per_split;
  P.dyOy = a->dy.oy; P.dyOx = a->dy.ox;
  for (int i = 0; i < WG_MAX_KT; ++i) P.kt[i] = p.kt[i < p.n_ktiles ? i : 0];
  P.ws = (float*)a->workspace;
  P.bias_ws = a->dbias ? P.ws + (size_t)p.splits * p.T * a->N * p.Ktot : nullptr;
  CUtensorMap tmDy, tmS[3];
  int rc = wg_tmap(&tmDy, a->dy, a->B);
  if (rc) return rc;
  for (int s = 0; s < 3; ++s) {
    if (s < a->n_src) {
      P.srcOy[s] = a->src[s].oy; P.srcOx[s] = a->src[s].ox;
      rc = wg_tmap(&tmS[s], a->src[s], a->B);
      if (rc) return rc;
    } else {
      P.srcOy[s] = P.srcOx[s] = 0;
      tmS[s] = tmS[0];
    }
  }
  const int jobs = p.splits * p.T * p.n_ntiles * p.n_ktiles;
  if (p.MH == 2) {
    PHT_SMEM_ATTR_ONCE(wgrad_tc_kernel<2>, WgCfg<2>::SMEM_BYTES);
    PHT_CUDA(launch_pdl(wgrad_tc_kernel<2>, dim3(jobs), dim3(WG_THREADS), WgCfg<2>::SMEM_BYTES, st, tmDy, tmS[0], tmS[1], tmS[2], P));
  } else {
    PHT_SMEM_ATTR_ONCE(wgrad_tc_kernel<1>, WgCfg<1>::SMEM_BYTES);
    PHT_CUDA(launch_pdl(wgrad_tc_kernel<1>, dim3(jobs), dim3(WG_THREADS), WgCfg<1>::SMEM_BYTES, st, tmDy, tmS[0], tmS[1], tmS[2], P));
  }
  PHT_LAUNCH_CHECK();
  if (defer) {   // leave the partials in the workspace; the caller batches the reduction
    defer->partials = P.ws; defer->dw = a->dw; defer->bias_partials = P.bias_ws; defer->dbias = a->dbias;
    defer->elems = (int64_t)p.T * a->N * p.Ktot; defer->splits = p.splits; defer->bias_rows = p.splits * p.T; defer->N = a->N;
    defer->pad_ = 0;
    count_launch(CNT_WGRAD_TC);
    *handled = true;
    return PHT_OK;
  }
  long long n4 = (long long)p.T * a->N * p.Ktot / 4;
  int grid = (int)((n4 + 255) / 256);
  int cap = sm_count() * 8;
  if (grid > cap) grid = cap;
  wgrad_reduce_splits_kernel<<<grid + (P.bias_ws ? 1 : 0), 256, 0, st>>>((const float4*)a->workspace, (float4*)a->dw, n4, p.splits,
                                                                       P.bias_ws, p.splits * p.T, a->dbias, a->N);
  PHT_LAUNCH_CHECK();
  count_launch(CNT_WGRAD_TC);
  count_launch(CNT_OTHER);
  *handled = true;
  return PHT_OK;
}

}  // namespace pht
