// Error reporting + device probe for libmelissa_b200.
#include <stdarg.h>

#include "common.cuh"

#include <atomic>
static thread_local char g_err[512] = "";
static std::atomic<unsigned long long> g_launches{0};
void mls_count_launch(int n) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }
extern "C" unsigned long long mls_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

void mls_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" int mls_version(void) { return MLS_VERSION; }
extern "C" const char* mls_last_error(void) { return g_err; }

extern "C" int mls_device_info(int32_t* sm_count, int32_t* cc_major, int32_t* cc_minor) {
  int dev = 0;
  MLS_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  MLS_CUDA(cudaGetDeviceProperties(&prop, dev));
  if (sm_count) *sm_count = prop.multiProcessorCount;
  if (cc_major) *cc_major = prop.major;
  if (cc_minor) *cc_minor = prop.minor;
  return MLS_OK;
}

// ---- tuning options (process wide).  "fused_conv": 1 = run the GATv2 convolutions of L-DGN with the fused
// projection+attention kernel (conv_fused.cu; experimental, slower than the two-kernel path in round 1).
// Initial values come from the environment (MLS_FUSED_CONV).
#include <stdlib.h>
#include <string.h>
static int g_fused_conv = -1;
// "attn_mma": 1 (default) = in discrete-feature mode run conv1's attention through the pair-logit table and
// the tensor-core aggregation kernel (attn_table.cu); 0 = gather kernel only.  Environment: MLS_ATTN_MMA.
static int g_attn_mma = -1;
extern "C" int mls_get_option(const char* key) {
  if (key && !strcmp(key, "attn_mma")) {
    if (g_attn_mma < 0) { const char* e = getenv("MLS_ATTN_MMA"); g_attn_mma = e ? atoi(e) : 1; }
    return g_attn_mma;
  }
  if (key && !strcmp(key, "fused_conv")) {
    if (g_fused_conv < 0) { const char* e = getenv("MLS_FUSED_CONV"); g_fused_conv = e ? atoi(e) : 0; }
    return g_fused_conv;
  }
  return -1;
}
extern "C" int mls_set_option(const char* key, int value) {
  if (key && !strcmp(key, "fused_conv")) { g_fused_conv = value ? 1 : 0; return MLS_OK; }
  if (key && !strcmp(key, "attn_mma")) { g_attn_mma = value ? 1 : 0; return MLS_OK; }
  mls_set_error("unknown option %s", key ? key : "(null)");
  return MLS_ERR_INVALID;
}
