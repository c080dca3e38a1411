// Library-wide state: last-error string, launch counter, device properties, version.
#include "common.cuh"
#include <stdarg.h>
#include <atomic>

static thread_local char g_err[1024] = "";
static std::atomic<long long> g_launches{0};

void dlb_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int dlb_check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    dlb_set_error("%s: launch failed: %s", what, cudaGetErrorString(e));
    return (int)e;
  }
  return DLB_OK;
}

void dlb_count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

static std::atomic<int> g_sm_budget{0};

int dlb_num_sms() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess ||
        sms <= 0)
      sms = 148;
  }
  const int budget = g_sm_budget.load(std::memory_order_relaxed);
  return (budget > 0 && budget < sms) ? budget : sms;
}

// Persistent kernels (GEMMs, attention) size their grids to dlb_num_sms(). While gradient buckets are being all-reduced
// behind backward, the data-parallel reducer lowers the budget by the collective's CTA count so that the NCCL kernel gets
// SMs of its own instead of time-slicing with a statically scheduled persistent CTA (which makes that CTA the straggler of
// every GEMM launched meanwhile). 0 restores all SMs. Rounded down to an even count (CTA-pair kernels). Returns the old budget.
DLB_EXPORT int dlb_set_sm_budget(int sms) {
  if (sms < 0) sms = 0;
  if (sms > 0 && sms < 16) sms = 16;
  return g_sm_budget.exchange(sms & ~1, std::memory_order_relaxed);
}

DLB_EXPORT const char* dlb_last_error(void) { return g_err; }
DLB_EXPORT int dlb_version(void) { return 100; }  // 0.1.0
DLB_EXPORT long long dlb_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
DLB_EXPORT void dlb_reset_launch_count(void) { g_launches.store(0, std::memory_order_relaxed); }

// Returns 0 when the current device is a compute-capability 10.x part this library was built for.
DLB_EXPORT int dlb_device_check(void) {
  int dev = 0, major = 0, minor = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) { dlb_set_error("cudaGetDevice: %s", cudaGetErrorString(e)); return (int)e; }
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  if (major != 10) {
    dlb_set_error("libdiffulab_b200 is built for sm_100a only; device %d is sm_%d%d", dev, major, minor);
    return DLB_ERR_UNSUPPORTED;
  }
  return DLB_OK;
}
