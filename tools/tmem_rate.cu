// tcgen05.ld throughput / latency microbenchmark (sm_100a).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I pixel_heal_thyself_b200/csrc -o tools/_bin/tmem_rate tools/tmem_rate.cu
#include <cstdio>
#include <cstdlib>
#include "tc_common.cuh"

using namespace pht::tc;

// mode 0: x32 loads, wait after every load (latency); 1: x32, wait after 4 loads; 2: x16, wait after every; 3: x16 wait after 4
template <int MODE>
__global__ void __launch_bounds__(512, 1) tmem_kernel(int reps, long long* out, float* sink) {
  __shared__ uint32_t slot;
  if (threadIdx.x < 32) tmem_alloc(&slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  const int warp = threadIdx.x >> 5;
  const uint32_t lane_addr = tm + ((uint32_t)((warp & 3) * 32) << 16);
  float acc = 0.f;
  __syncthreads();
  const long long t0 = clock64();
  for (int r = 0; r < reps; ++r) {
    if (MODE == 0 || MODE == 1) {
      uint32_t a[32], b[32], c[32], d[32];
      tmem_ld32(lane_addr + 0, a);
      if (MODE == 0) tmem_ld_wait();
      tmem_ld32(lane_addr + 32, b);
      if (MODE == 0) tmem_ld_wait();
      tmem_ld32(lane_addr + 64, c);
      if (MODE == 0) tmem_ld_wait();
      tmem_ld32(lane_addr + 96, d);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) acc += __uint_as_float(a[j] ^ b[j] ^ c[j] ^ d[j]);
    } else {
      uint32_t a[16], b[16], c[16], d[16];
      tmem_ld16(lane_addr + 0, a);
      if (MODE == 2) tmem_ld_wait();
      tmem_ld16(lane_addr + 16, b);
      if (MODE == 2) tmem_ld_wait();
      tmem_ld16(lane_addr + 32, c);
      if (MODE == 2) tmem_ld_wait();
      tmem_ld16(lane_addr + 48, d);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j) acc += __uint_as_float(a[j] ^ b[j] ^ c[j] ^ d[j]);
    }
  }
  const long long t1 = clock64();
  if (threadIdx.x % 32 == 0) out[blockIdx.x * 16 + warp] = t1 - t0;
  if (acc == 12345.f) sink[0] = acc;
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tm, 512); }
}

template <int MODE>
void run(int nwarps, const char* name, long long* dout, float* sink) {
  const int reps = 1000;
  tmem_kernel<MODE><<<1, nwarps * 32>>>(reps, dout, sink);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); exit(1); }
  long long h[16];
  cudaMemcpy(h, dout, sizeof(h), cudaMemcpyDeviceToHost);
  long long mx = 0;
  for (int w = 0; w < nwarps; ++w) mx = h[w] > mx ? h[w] : mx;
  const double per_ld = (double)mx / (reps * 4);
  const double bytes = (MODE < 2 ? 4096.0 : 2048.0);
  printf("%-34s warps=%2d  clk/ld(per warp)=%7.1f  SM B/clk=%7.1f  per-subpartition B/clk=%7.1f\n", name, nwarps, per_ld,
         bytes * nwarps / per_ld, bytes * nwarps / per_ld / (nwarps < 4 ? nwarps : 4));
}

int main() {
  long long* dout;
  float* sink;
  cudaMalloc(&dout, sizeof(long long) * 16 * 4);
  cudaMalloc(&sink, 4);
  for (int nw : {1, 2, 4, 8, 16}) {
    run<0>(nw, "x32 (4 KB/warp), wait each", dout, sink);
    run<1>(nw, "x32 (4 KB/warp), wait per 4", dout, sink);
    run<2>(nw, "x16 (2 KB/warp), wait each", dout, sink);
    run<3>(nw, "x16 (2 KB/warp), wait per 4", dout, sink);
  }
  return 0;
}
