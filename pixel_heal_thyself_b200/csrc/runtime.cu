// Error reporting, launch counters, ABI introspection.
#include <atomic>
#include <map>
#include <mutex>
#include <stdarg.h>
#include <string.h>
#include <tuple>

#include "common.cuh"

namespace pht {

static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_counters[8];
static std::atomic<int> g_force_simple{0};
static std::atomic<int> g_pdl{1};
static std::atomic<int> g_bf16_fallback{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int slot, uint64_t n) { g_counters[slot & 7].fetch_add(n, std::memory_order_relaxed); }
void set_tc_cfg(int v);  // igemm_tc.cu
void set_wgrad_split_div(int v);  // wgrad_tc.cu
void set_serpentine(int v);  // igemm_tc.cu
void set_strips(int v);  // igemm_tc.cu
void set_cta_pairs(int v);  // igemm_tc.cu
void set_conv_grid_cap(int v);  // igemm_tc.cu
void set_half_ring(int v);  // igemm_tc.cu
void set_conv_trace(int v);  // igemm_tc.cu
int read_conv_trace(long long* host, int n);  // igemm_tc.cu
void set_attn_trace(int v);  // attention_tc.cu
void set_attn_bwd_direct(int v);  // attention_tc.cu
int read_attn_trace(long long* host, int n);  // attention_tc.cu
bool force_simple() { return g_force_simple.load(std::memory_order_relaxed) != 0; }
bool bf16_fallback_allowed() { return g_bf16_fallback.load(std::memory_order_relaxed) != 0; }

int cur_device() {
  int d = 0;
  cudaGetDevice(&d);
  return d;
}
int sm_count() {
  static std::atomic<int> cache[PHT_MAX_DEVICES];
  const int d = cur_device() & (PHT_MAX_DEVICES - 1);
  int n = cache[d].load(std::memory_order_relaxed);
  if (!n) {
    n = 148;
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, d);
    cache[d].store(n, std::memory_order_relaxed);
  }
  return n;
}

void* stream_scratch(cudaStream_t st, int slot, size_t bytes) {
  static std::mutex mu;
  static std::map<std::tuple<int, cudaStream_t, int>, std::pair<void*, size_t>> pool;
  std::lock_guard<std::mutex> lk(mu);
  auto key = std::make_tuple(cur_device(), st, slot);
  auto it = pool.find(key);
  if (it != pool.end() && it->second.second >= bytes) return it->second.first;
  void* p = nullptr;
  if (cudaMalloc(&p, bytes) != cudaSuccess || cudaMemset(p, 0, bytes) != cudaSuccess) {
    cudaGetLastError();
    set_error("stream_scratch: cannot allocate %zu bytes (first use of a launch that needs library scratch must not "
              "happen under stream capture)", bytes);
    return nullptr;
  }
  pool[key] = std::make_pair(p, bytes);   // (a smaller earlier block of the same key is leaked: sizes are constants)
  return p;
}
namespace tc { bool pdl_enabled() { return g_pdl.load(std::memory_order_relaxed) != 0; } }

}  // namespace pht

extern "C" {

int pht_abi_version(void) { return PHT_ABI_VERSION; }
const char* pht_last_error(void) { return pht::g_err; }
void pht_get_counters(uint64_t* c) {
  for (int i = 0; i < 8; ++i) c[i] = pht::g_counters[i].load(std::memory_order_relaxed);
}
void pht_add_counters(const uint64_t* c) {
  for (int i = 0; i < 8; ++i) pht::g_counters[i].fetch_add(c[i], std::memory_order_relaxed);
}
void pht_reset_counters(void) {
  for (int i = 0; i < 8; ++i) pht::g_counters[i].store(0, std::memory_order_relaxed);
}
int pht_set_option(const char* name, int value) {
  if (name && !strcmp(name, "tc_cfg")) { pht::set_tc_cfg(value); return PHT_OK; }
  if (name && !strcmp(name, "wgrad_split_div")) { pht::set_wgrad_split_div(value); return PHT_OK; }
  if (name && !strcmp(name, "conv_trace")) { pht::set_conv_trace(value); return PHT_OK; }
  if (name && !strcmp(name, "strips")) { pht::set_strips(value); return PHT_OK; }
  if (name && !strcmp(name, "cta_pairs")) { pht::set_cta_pairs(value); return PHT_OK; }
  if (name && !strcmp(name, "conv_grid_cap")) { pht::set_conv_grid_cap(value); return PHT_OK; }
  if (name && !strcmp(name, "half_ring")) { pht::set_half_ring(value); return PHT_OK; }
  if (name && !strcmp(name, "serpentine")) { pht::set_serpentine(value); return PHT_OK; }
  if (name && !strcmp(name, "pdl")) { pht::g_pdl.store(value ? 1 : 0, std::memory_order_relaxed); return PHT_OK; }
  if (name && !strcmp(name, "attn_trace")) { pht::set_attn_trace(value); return PHT_OK; }
  if (name && !strcmp(name, "attn_bwd_direct")) { pht::set_attn_bwd_direct(value); return PHT_OK; }
  if (name && !strcmp(name, "bf16_fallback")) { pht::g_bf16_fallback.store(value ? 1 : 0, std::memory_order_relaxed); return PHT_OK; }
  pht::set_error("pht_set_option: unknown option");
  return PHT_ERR_INVALID;
}
int pht_conv_gemm_trace(int64_t* host, int32_t n) { return pht::read_conv_trace((long long*)host, n); }
int pht_attn_bwd_trace(int64_t* host, int32_t n) { return pht::read_attn_trace((long long*)host, n); }
void pht_set_force_simple(int on) { pht::g_force_simple.store(on ? 1 : 0, std::memory_order_relaxed); }

}  // extern "C"
