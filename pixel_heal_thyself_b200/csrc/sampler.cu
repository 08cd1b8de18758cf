// Patch-index sampler: Poisson-disk dart throwing, bit-exact with the
// reference's sample_patches_dart_throwing(shape, P, n, random.Random(seed))
// (pht/models/afgsa/preprocessing.py:171-213).  The RNG is a device
// re-implementation of CPython's MT19937 front end:
//   seed(int) -> init_by_array(32-bit limbs), randint(a,b) -> a + _randbelow(b-a+1),
//   _randbelow(n): k = bit_length(n); r = getrandbits(k) until r < n;
//   getrandbits(k<=32) = genrand_uint32() >> (32-k).
// One warp per image (the accept test against the accepted set is the parallel
// part: lanes stride over the accepted points, ballot for early reject); the
// 624-word state lives in shared memory and is regenerated cooperatively.
#include <math.h>

#include "common.cuh"

namespace pht {

constexpr int MT_N = 624, MT_M = 397;

struct Mt {
  uint32_t* mt;  // smem [624]
  int idx;
};

__device__ void mt_seed(Mt& g, unsigned long long seed, int lane) {
  // init_genrand(19650218) + init_by_array(key): serial, lane 0 only
  if (lane == 0) {
    uint32_t* mt = g.mt;
    mt[0] = 19650218u;
    for (int i = 1; i < MT_N; ++i) mt[i] = 1812433253u * (mt[i - 1] ^ (mt[i - 1] >> 30)) + (uint32_t)i;
    uint32_t key[2] = {(uint32_t)(seed & 0xffffffffull), (uint32_t)(seed >> 32)};
    int klen = key[1] ? 2 : 1;
    int i = 1, j = 0;
    int k = MT_N > klen ? MT_N : klen;
    for (; k; --k) {
      mt[i] = (mt[i] ^ ((mt[i - 1] ^ (mt[i - 1] >> 30)) * 1664525u)) + key[j] + (uint32_t)j;
      ++i; ++j;
      if (i >= MT_N) { mt[0] = mt[MT_N - 1]; i = 1; }
      if (j >= klen) j = 0;
    }
    for (k = MT_N - 1; k; --k) {
      mt[i] = (mt[i] ^ ((mt[i - 1] ^ (mt[i - 1] >> 30)) * 1566083941u)) - (uint32_t)i;
      ++i;
      if (i >= MT_N) { mt[0] = mt[MT_N - 1]; i = 1; }
    }
    mt[0] = 0x80000000u;
  }
  g.idx = MT_N;
  __syncwarp();
}

__device__ __forceinline__ uint32_t mt_mix(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t y = (a & 0x80000000u) | (b & 0x7fffffffu);
  return c ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
}

// cooperative regeneration: three dependency-free phases, read-all / sync / write-all
__device__ void mt_twist(Mt& g, int lane) {
  uint32_t* mt = g.mt;
  uint32_t v[8];
  // phase 1: k in [0, 227): old mt[k], old mt[k+1], old mt[k+397]
  for (int i = 0, k = lane; i < 8; ++i, k += 32) v[i] = k < 227 ? mt_mix(mt[k], mt[k + 1], mt[k + MT_M]) : 0u;
  __syncwarp();
  for (int i = 0, k = lane; i < 8; ++i, k += 32) if (k < 227) mt[k] = v[i];
  __syncwarp();
  // phase 2: k in [227, 454): old mt[k], old mt[k+1], NEW mt[k-227]
  for (int i = 0, k = 227 + lane; i < 8; ++i, k += 32) v[i] = k < 454 ? mt_mix(mt[k], mt[k + 1], mt[k - 227]) : 0u;
  __syncwarp();
  for (int i = 0, k = 227 + lane; i < 8; ++i, k += 32) if (k < 454) mt[k] = v[i];
  __syncwarp();
  // phase 3: k in [454, 624): old mt[k], old mt[k+1] (new mt[0] for k = 623), NEW mt[k-227]
  for (int i = 0, k = 454 + lane; i < 8; ++i, k += 32) v[i] = k < MT_N ? mt_mix(mt[k], mt[(k + 1) % MT_N], mt[k - 227]) : 0u;
  __syncwarp();
  for (int i = 0, k = 454 + lane; i < 8; ++i, k += 32) if (k < MT_N) mt[k] = v[i];
  __syncwarp();
  g.idx = 0;
}

// warp-uniform: every lane executes this with identical state
__device__ __forceinline__ uint32_t mt_u32(Mt& g, int lane) {
  if (g.idx >= MT_N) mt_twist(g, lane);
  uint32_t y = g.mt[g.idx++];
  y ^= y >> 11;
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  y ^= y >> 18;
  return y;
}

__device__ __forceinline__ int mt_randbelow(Mt& g, int n, int lane) {
  int k = 32 - __clz(n);  // n.bit_length()
  uint32_t r = mt_u32(g, lane) >> (32 - k);
  while (r >= (uint32_t)n) r = mt_u32(g, lane) >> (32 - k);
  return (int)r;
}

// dart throwing with the warp-uniform RNG; px/py in shared memory.  Returns false if the radius shrank > 1e5 times.
__device__ bool dart_throw(Mt& g, int lane, int Hf, int Wf, int P, int n, int max_iter, int* px, int* py) {
  // preprocessing.py:187-192 in IEEE double, same operation order
  double radius = sqrt(((double)((long long)Hf * Wf) / (double)n) / 3.141592653589793);
  double min_sq = (2.0 * radius) * (2.0 * radius);
  const int x_span = Wf - P - 1 + 1, y_span = Hf - P - 1 + 1;  // randint(0, max) -> randbelow(max + 1)
  int shrinks = 0;
  for (int idx = 0; idx < n; ++idx) {
    bool done = false;
    while (!done) {
      for (int it = 0; it < max_iter; ++it) {
        int x = mt_randbelow(g, x_span, lane);
        int y = mt_randbelow(g, y_span, lane);
        bool reject = false;
        for (int base = 0; base < idx; base += 32) {
          int i = base + lane;
          bool bad = false;
          if (i < idx) {
            long long dx = px[i] - x, dy = py[i] - y;
            double d2 = (double)(dx * dx + dy * dy);
            bad = !(d2 > min_sq);
          }
          if (__any_sync(0xffffffffu, bad)) { reject = true; break; }
        }
        if (!reject) {
          if (lane == 0) { px[idx] = x; py[idx] = y; }
          __syncwarp();
          done = true;
          break;
        }
      }
      if (!done) {
        radius *= 0.96;
        min_sq = (2.0 * radius) * (2.0 * radius);
        if (++shrinks > 100000) return false;
      }
    }
  }
  return true;
}

__global__ void __launch_bounds__(32) sample_patches_kernel(const long long* __restrict__ seeds, int Hf, int Wf, int P,
                                                            int n, int max_iter, int* __restrict__ out) {
  extern __shared__ uint32_t smem_u[];
  uint32_t* mt = smem_u;
  int* px = (int*)(mt + MT_N);
  int* py = px + n;
  const int lane = threadIdx.x;
  const int img = blockIdx.x;
  Mt g;
  g.mt = mt;
  long long s = seeds[img];
  mt_seed(g, (unsigned long long)(s < 0 ? -s : s), lane);
  const bool ok = dart_throw(g, lane, Hf, Wf, P, n, max_iter, px, py);
  for (int i = lane; i < n; i += 32) {
    out[((long long)img * n + i) * 2 + 0] = ok ? px[i] : -1;
    out[((long long)img * n + i) * 2 + 1] = ok ? py[i] : -1;
  }
}

// random.random(): 53-bit double from two draws (CPython _random.c: (a >> 5, b >> 6))
__device__ __forceinline__ double mt_random(Mt& g, int lane) {
  const uint32_t a = mt_u32(g, lane) >> 5, b = mt_u32(g, lane) >> 6;
  return ((double)a * 67108864.0 + (double)b) * (1.0 / 9007199254740992.0);
}

// importance_sampling (preprocessing.py:284-322): dart throwing, then prune_patches (:259-281) on the SAME RNG stream:
// regions of 4P x 4P pixels in serpentine order (:223-238), a patch belongs to the first region whose INCLUSIVE bounds
// contain its centre (:241-256); error diffusion in numpy-float32 scalar arithmetic (NEP 50), rng.random() rounded to
// float32 for the comparison.  Everything is warp-uniform (the RNG state is shared), the loop is serial like the reference.
__global__ void __launch_bounds__(32) importance_sample_kernel(const long long* __restrict__ seeds, int Hf, int Wf, int P,
                                                               int n, int max_iter, const float* __restrict__ imp,
                                                               int* __restrict__ out, int* __restrict__ counts) {
  extern __shared__ uint32_t smem_u[];
  uint32_t* mt = smem_u;
  int* px = (int*)(mt + MT_N);
  int* py = px + n;
  float* iv = (float*)(py + n);
  int* state = (int*)(iv + n);    // 0 = not yet visited, 1 = visited
  const int lane = threadIdx.x;
  const int img = blockIdx.x;
  Mt g;
  g.mt = mt;
  long long s = seeds[img];
  mt_seed(g, (unsigned long long)(s < 0 ? -s : s), lane);
  const bool ok = dart_throw(g, lane, Hf, Wf, P, n, max_iter, px, py);
  int* o = out + (long long)img * n * 2;
  for (int i = lane; i < 2 * n; i += 32) o[i] = -1;
  if (!ok) {
    if (lane == 0) counts[img] = -1;
    return;
  }
  const int pad = P / 2;
  for (int i = lane; i < n; i += 32) {
    px[i] += pad;                 // centres (importance_sampling: patches + pad)
    py[i] += pad;
    iv[i] = imp[((long long)img * Hf + py[i]) * Wf + px[i]];
    state[i] = 0;
  }
  __syncwarp();
  const int step = 4 * P;
  int count = 0;
  float error = 0.f;
  for (int ry = 0, row = 0; ry < Hf; ry += step, ++row) {
    const int ncol = (Wf + step - 1) / step;
    for (int k = 0; k < ncol; ++k) {
      const int rx = ((row & 1) ? (ncol - 1 - k) : k) * step;
      for (int i = 0; i < n; ++i) {
        if (state[i]) continue;
        const int x = px[i], y = py[i];
        if (!(rx <= x && x <= rx + step && ry <= y && y <= ry + step)) continue;
        __syncwarp();
        if (lane == 0) state[i] = 1;
        const float v = iv[i];
        const float u = (float)mt_random(g, lane);
        if (__fsub_rn(v, error) > u) {
          if (lane == 0) { o[count * 2] = x; o[count * 2 + 1] = y; }
          ++count;
          error = __fadd_rn(error, __fsub_rn(1.0f, v));
        } else {
          error = __fadd_rn(error, __fsub_rn(0.0f, v));
        }
        __syncwarp();
      }
    }
  }
  if (lane == 0) counts[img] = count;
}

}  // namespace pht

using namespace pht;

extern "C" int pht_sample_patches(const int64_t* seeds, int32_t n_img, int32_t Hf, int32_t Wf, int32_t P, int32_t n,
                                  int32_t max_iter, int32_t* out, void* stream) {
  PHT_CHECK_ARG(seeds && out && n_img > 0 && n > 0 && max_iter > 0, "sample_patches: bad args");
  PHT_CHECK_ARG(Wf - P - 1 >= 0 && Hf - P - 1 >= 0, "sample_patches: patch larger than frame");
  PHT_CHECK_ARG(n <= 8192, "sample_patches: at most 8192 patches per image");
  size_t smem = MT_N * sizeof(uint32_t) + 2 * (size_t)n * sizeof(int);
  PHT_CUDA(cudaFuncSetAttribute(sample_patches_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  sample_patches_kernel<<<n_img, 32, smem, (cudaStream_t)stream>>>((const long long*)seeds, Hf, Wf, P, n, max_iter, out);
  count_launch(CNT_OTHER);
  PHT_LAUNCH_CHECK();
  return PHT_OK;
}

extern "C" int pht_importance_sample(const int64_t* seeds, int32_t n_img, int32_t Hf, int32_t Wf, int32_t P, int32_t n,
                                     int32_t max_iter, const float* imp, int32_t* out_centres, int32_t* out_counts,
                                     void* stream) {
  PHT_CHECK_ARG(seeds && imp && out_centres && out_counts && n_img > 0 && n > 0 && max_iter > 0, "importance_sample: bad args");
  PHT_CHECK_ARG(Wf - P - 1 >= 0 && Hf - P - 1 >= 0, "importance_sample: patch larger than frame");
  PHT_CHECK_ARG(n <= 8192, "importance_sample: at most 8192 patches per image");
  size_t smem = MT_N * sizeof(uint32_t) + 4 * (size_t)n * sizeof(int);
  PHT_CUDA(cudaFuncSetAttribute(importance_sample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  importance_sample_kernel<<<n_img, 32, smem, (cudaStream_t)stream>>>((const long long*)seeds, Hf, Wf, P, n, max_iter, imp,
                                                                      out_centres, out_counts);
  count_launch(CNT_OTHER);
  PHT_LAUNCH_CHECK();
  return PHT_OK;
}
