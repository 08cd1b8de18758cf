// tcgen05.mma issue-rate microbenchmark (sm_100a): clocks per MMA for a list of (M, N, a_major, b_major) shapes,
// operands in shared memory (SS mode, bf16, K = 16 per instruction).  Used to size the attention kernels' tile shapes.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I pixel_heal_thyself_b200/csrc -o tools/_bin/mma_rate tools/mma_rate.cu
#include <cstdio>
#include <cstdlib>
#include "tc_common.cuh"

using namespace pht::tc;

struct Shape { int M, N, amn, bmn, nacc; };

__global__ void __launch_bounds__(160, 1) rate_kernel(const Shape* shapes, int nshape, int reps, long long* out, int ld_warps, int ld_cols, long long* ld_stats) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  __shared__ volatile int stop_flag;
  if (threadIdx.x == 0) stop_flag = 0;
  for (int i = threadIdx.x; i < 160 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  if (threadIdx.x < 32) tmem_alloc(&slot, 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  if (threadIdx.x == 0) {
    uint32_t ph = 0;
    for (int s = 0; s < nshape; ++s) {
      const Shape sh = shapes[s];
      const uint32_t idesc = umma_idesc_bf16(sh.M, sh.N, sh.amn, sh.bmn);
      const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 64 * 1024);
      for (int pass = 0; pass < 2; ++pass) {   // pass 0 warms up
        const long long t0 = clock64();
        uint64_t ad[4], bd[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          ad[k] = sh.amn ? umma_desc_mn_sw128(a0 + k * 2048, 8192, 1024) : umma_desc_k_sw128(a0) + 2 * k;
          bd[k] = sh.bmn ? umma_desc_mn_sw128(b0 + k * 2048, 8192, 1024) : umma_desc_k_sw128(b0) + 2 * k;
        }
        const uint32_t cstep = sh.nacc == 4 ? 128u : (sh.nacc == 2 ? 256u : 0u);
#pragma unroll 1
        for (int r = 0; r < reps; r += 8) {   // 8 MMAs per trip, no per-MMA integer work
#pragma unroll
          for (int u = 0; u < 8; ++u)
            umma_bf16(tm + (uint32_t)(u & 3) * cstep % 512u, ad[u & 3], bd[u & 3], idesc, 1u);
        }
        umma_commit(&bar);
        mbar_wait(&bar, ph);
        ph ^= 1;
        const long long t1 = clock64();
        if (pass) out[blockIdx.x * nshape + s] = t1 - t0;
      }
    }
  }
  if (threadIdx.x == 0) stop_flag = 1;
  // warps 1..ld_warps hammer tcgen05.ld on their TMEM sub-partition while thread 0 issues the MMAs
  if (threadIdx.x >= 32 && (int)(threadIdx.x >> 5) <= ld_warps) {
    const uint32_t la = tm + ((uint32_t)(((threadIdx.x >> 5) & 3) * 32) << 16) + ld_cols;
    float acc = 0.f;
    long long n = 0;
    const long long t0 = clock64();
    while (!stop_flag) {
      uint32_t r[32];
      tmem_ld32(la, r);
      tmem_ld_wait();
      acc += __uint_as_float(r[0] ^ r[31]);
      ++n;
    }
    const long long t1 = clock64();
    if ((threadIdx.x & 31) == 0 && blockIdx.x == 0) { ld_stats[(threadIdx.x >> 5) * 2] = n; ld_stats[(threadIdx.x >> 5) * 2 + 1] = t1 - t0; }
    if (acc == 123.f) out[0] = 1;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tm, 512); }
}

// CTA-pair rate: a 2-CTA cluster, the leader issues tcgen05.mma.cta_group::2 (M = 256 over both CTAs), operands are
// whatever the (zeroed) shared memory of the two CTAs holds.
__global__ void __launch_bounds__(160, 1) rate2_kernel(int M, int N, int reps, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < 160 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  if (threadIdx.x < 32) tmem_alloc_2sm(&slot, 512);
  fence_proxy_async();
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tm = slot;
  const uint32_t rank = cluster_ctarank();
  if (threadIdx.x < 32) {
    uint32_t ph = 0;
    const uint32_t idesc = umma_idesc_bf16(M, N, 0, 0);
    const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 64 * 1024);
    for (int pass = 0; pass < 2; ++pass) {
      const long long t0 = clock64();
      if (rank == 0) {
#pragma unroll 1
        for (int r = 0; r < reps; r += 4) {
#pragma unroll
          for (int u = 0; u < 4; ++u) umma_bf16_elect_2sm(tm, umma_desc_k_sw128(a0) + 2 * u, umma_desc_k_sw128(b0) + 2 * u, idesc, 1u);
        }
        umma_commit_elect_2sm(&bar);
      }
      mbar_wait(&bar, ph);
      ph ^= 1;
      const long long t1 = clock64();
      if (pass && threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
    }
  }
  tc_fence_before();
  cluster_sync_all();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc_2sm(tm, 512); }
}

static void run_pairs(long long* dout, int grid, int reps) {
  cudaFuncSetAttribute(rate2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  printf("---- CTA pairs: tcgen05.mma.cta_group::2, K-major bf16 operands ----\n%5s %5s %12s %18s\n", "M", "N", "clk/MMA", "MAC/clk/SM");
  const int shapes[][2] = {{256, 256}, {256, 128}, {256, 64}, {128, 256}, {128, 128}};
  for (auto& sh : shapes) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(160); cfg.dynamicSmemBytes = 200 * 1024;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, rate2_kernel, sh[0], sh[1], reps, dout);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return; }
    long long* h = (long long*)malloc(sizeof(long long) * grid);
    cudaMemcpy(h, dout, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (int b = 0; b < grid; ++b) mx = h[b] > mx ? h[b] : mx;
    const double c = (double)mx / reps;
    printf("%5d %5d %12.1f %18.0f\n", sh[0], sh[1], c, (double)sh[0] / 2 * sh[1] * 16 / c);
    free(h);
  }
}

int main(int argc, char** argv) {
  const Shape hs[] = {
      {128, 256, 0, 0, 1}, {128, 256, 0, 0, 2}, {128, 128, 0, 0, 1}, {128, 128, 0, 0, 2}, {128, 64, 0, 0, 1}, {128, 64, 0, 0, 2},
      {128, 64, 0, 0, 4},  {128, 64, 1, 1, 4},  {128, 32, 0, 0, 4},  {128, 16, 0, 0, 4},  {64, 256, 0, 0, 1}, {64, 256, 0, 0, 2},
      {64, 200, 0, 0, 2},  {64, 112, 0, 0, 2},  {64, 64, 0, 0, 1},   {64, 64, 0, 0, 2},   {64, 64, 0, 0, 4},  {64, 64, 0, 1, 4},
      {64, 128, 0, 0, 2},  {128, 112, 0, 0, 2}, {128, 72, 0, 1, 4},  {64, 32, 0, 0, 4},   {64, 16, 0, 0, 4},  {128, 96, 0, 0, 4},
  };
  const int n = sizeof(hs) / sizeof(hs[0]);
  const int reps = 256;
  Shape* ds;
  long long* dout;
  const int grid = 148;
  cudaMalloc(&ds, sizeof(hs));
  cudaMalloc(&dout, sizeof(long long) * n * grid);
  cudaMemcpy(ds, hs, sizeof(hs), cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  long long* dstats;
  cudaMalloc(&dstats, sizeof(long long) * 16);
  run_pairs(dout, grid, reps);
  if (argc > 1) return 0;   // any argument: the CTA-pair table only
  for (int ldw : {0, 4, 104}) {
    const int g = grid;
    const int ld_warps = ldw % 100, ld_cols = ldw >= 100 ? 0 : 384;
    printf("---- %d warps looping tcgen05.ld.x32 (+wait) at TMEM column %d (MMAs accumulate into columns [0,256)) ----\n", ld_warps, ld_cols);
    printf("%5s %5s %4s %4s %4s %12s %14s %16s\n", "M", "N", "aMN", "bMN", "nacc", "clk/MMA", "MAC/clk/SM", "clk per ld+wait");
    for (int s = 0; s < n; ++s) {
      cudaMemset(dstats, 0, sizeof(long long) * 16);
      rate_kernel<<<g, 160, 200 * 1024>>>(ds + s, 1, reps, dout, ld_warps, ld_cols, dstats);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
      long long* h = (long long*)malloc(sizeof(long long) * g);
      cudaMemcpy(h, dout, sizeof(long long) * g, cudaMemcpyDeviceToHost);
      long long st[16];
      cudaMemcpy(st, dstats, sizeof(st), cudaMemcpyDeviceToHost);
      long long mx = 0;
      for (int b = 0; b < g; ++b) mx = h[b] > mx ? h[b] : mx;
      const double c = (double)mx / reps;
      printf("%5d %5d %4d %4d %4d %12.1f %14.0f %16.1f\n", hs[s].M, hs[s].N, hs[s].amn, hs[s].bmn, hs[s].nacc, c,
             (double)hs[s].M * hs[s].N * 16 / c, st[2] ? (double)st[3] / st[2] : 0.0);
      free(h);
    }
  }
  return 0;
}
