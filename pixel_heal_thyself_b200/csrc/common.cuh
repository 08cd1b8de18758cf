// Shared helpers for libpht_b200 (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/pht_b200.h"

namespace pht {

void set_error(const char* fmt, ...);
void count_launch(int slot, uint64_t n = 1);
bool force_simple();
// bf16 launches that are not taken by a tcgen05 kernel fail (PHT_ERR_UNSUPPORTED) unless the "bf16_fallback" option
// allows the CUDA-core kernels: the production dtype must never drop to a 20x slower path silently
bool bf16_fallback_allowed();
int cur_device();        // cudaGetDevice
int sm_count();          // multiprocessor count of the current device (cached per device ordinal)
// Small library-owned device scratch private to (device, stream, slot): ticketed reductions of launches that overlap
// on different streams must not share partials.  Zero-initialised once; allocated at the first use (never under
// stream capture: warm-up runs come first).  Returns nullptr (and sets the error) on failure.
void* stream_scratch(cudaStream_t st, int slot, size_t bytes);

constexpr int PHT_MAX_DEVICES = 64;
struct PerDeviceOnce {   // "done once per device" flags (cudaFuncSetAttribute is per device, not per process)
  unsigned char done[PHT_MAX_DEVICES] = {};
};

enum CounterSlot { CNT_GEMM_TC = 0, CNT_GEMM_SIMPLE = 1, CNT_WGRAD_TC = 2, CNT_WGRAD_SIMPLE = 3, CNT_ATTN_TC = 4,
                   CNT_ATTN_SIMPLE = 5, CNT_OTHER = 6 };

#define PHT_CHECK_ARG(cond, ...)                 \
  do {                                           \
    if (!(cond)) {                               \
      pht::set_error(__VA_ARGS__);               \
      return PHT_ERR_INVALID;                    \
    }                                            \
  } while (0)

#define PHT_CUDA(call)                                                                        \
  do {                                                                                        \
    cudaError_t e__ = (call);                                                                 \
    if (e__ != cudaSuccess) {                                                                 \
      pht::set_error("%s:%d CUDA error %s (%s)", __FILE__, __LINE__, cudaGetErrorName(e__),   \
                     cudaGetErrorString(e__));                                                \
      return PHT_ERR_CUDA;                                                                    \
    }                                                                                         \
  } while (0)

#define PHT_LAUNCH_CHECK() PHT_CUDA(cudaGetLastError())

// opt the kernel in to `bytes` of dynamic shared memory, once per device
#define PHT_SMEM_ATTR_ONCE(kernel, bytes)                                                                     \
  do {                                                                                                        \
    static pht::PerDeviceOnce once__;                                                                         \
    const int d__ = pht::cur_device() & (pht::PHT_MAX_DEVICES - 1);                                           \
    if (!once__.done[d__]) {                                                                                  \
      PHT_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes)));      \
      once__.done[d__] = 1;                                                                                   \
    }                                                                                                         \
  } while (0)

typedef __nv_bfloat16 bf16;

// Device-side copy of pht_view with typed access.
struct View {
  char* ptr;
  int H, W, C, oy, ox, dtype;
  long long sb, sy, sx;
};

static inline View make_view(const pht_view& v) {
  View r;
  r.ptr = (char*)v.ptr;
  r.H = v.H; r.W = v.W; r.C = v.C; r.oy = v.oy; r.ox = v.ox; r.dtype = v.dtype;
  r.sb = v.sb; r.sy = v.sy; r.sx = v.sx;
  return r;
}
static inline View null_view() {
  View r;
  r.ptr = nullptr; r.H = r.W = r.C = r.oy = r.ox = 0; r.dtype = 0; r.sb = r.sy = r.sx = 0;
  return r;
}
static inline View make_view(const pht_view* v) { return v ? make_view(*v) : null_view(); }

template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<bf16>(bf16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f<bf16>(float v) { return __float2bfloat16_rn(v); }

// element offset of pixel (b, y, x) in a view, no bounds handling
__device__ __forceinline__ long long view_off(const View& v, int b, int y, int x) {
  return (long long)b * v.sb + (long long)y * v.sy + (long long)x * v.sx;
}
__device__ __forceinline__ bool view_inb(const View& v, int y, int x) {
  return (unsigned)y < (unsigned)v.H && (unsigned)x < (unsigned)v.W;
}

template <typename T>
__device__ __forceinline__ float view_ld(const View& v, int b, int y, int x, int c) {
  return to_f<T>(((const T*)v.ptr)[view_off(v, b, y, x) + c]);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

// border index maps shared by forward fill and backward fold.
// padded coordinate q in [0, n+1] -> interior coordinate in [0, n-1]
__host__ __device__ __forceinline__ int pad_src_index(int q, int n, int mode) {
  int i = q - 1;
  if (mode == PHT_PAD_REPLICATE) return i < 0 ? 0 : (i >= n ? n - 1 : i);
  return i < 0 ? -i : (i >= n ? 2 * n - 2 - i : i);  // reflect
}
// clamp / mirror of an arbitrary out-of-range coordinate (|overshoot| <= 2)
__host__ __device__ __forceinline__ int pad_index(int i, int n, int mode) {
  if (mode == PHT_PAD_REPLICATE) return i < 0 ? 0 : (i >= n ? n - 1 : i);
  return i < 0 ? -i : (i >= n ? 2 * n - 2 - i : i);
}

}  // namespace pht
