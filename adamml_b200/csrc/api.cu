// Error reporting + version entry points of the C-ABI (include/adamml_b200.h).
#include "common.cuh"
#include <stdarg.h>
#include <string.h>

static thread_local char g_err[512] = "";

void adamml_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static unsigned long long g_launches = 0;

int adamml_check_launch(const char* what) {
  __atomic_fetch_add(&g_launches, 1ULL, __ATOMIC_RELAXED);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    adamml_set_error("%s: CUDA launch failed: %s", what, cudaGetErrorString(e));
    return ADAMML_ERR_CUDA;
  }
  return ADAMML_OK;
}

static thread_local const int* g_live_n = nullptr;
static thread_local int g_live_cap = 0;

LiveLimit adamml_live_limit(long long capacity_items) {
  if (!g_live_n || g_live_cap <= 0 || capacity_items % g_live_cap) return LiveLimit{nullptr, 0};
  return LiveLimit{g_live_n, (int)(capacity_items / g_live_cap)};
}

extern "C" {

// see include/adamml_b200.h: device-side work limit of the calling thread's following forward launches
int adamml_set_live_clips(const int* live_clips, int clip_capacity) {
  g_live_n = live_clips;
  g_live_cap = live_clips ? clip_capacity : 0;
  return ADAMML_OK;
}

const char* adamml_last_error(void) { return g_err; }

int adamml_abi_version(void) { return 1; }

// Number of kernel launches issued through this library by the calling process (all
// threads); bench.py reports the per-step delta as "gpu_launches".
unsigned long long adamml_launch_count(void) { return g_launches; }

}  // extern "C"
