// Error plumbing, launch accounting and device gate for libwspc.
#include "common.cuh"
#include <atomic>

namespace wspc {

static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
  return WSPC_ERR_CUDA;
}

void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

// The library is compiled for sm_100a only; refuse anything else loudly
// instead of letting a launch fail with "no kernel image".
int check_arch() {
  static thread_local int cached_dev = -1;
  static thread_local int cached_rc = WSPC_OK;
  int dev = -1;
  WSPC_CUDA(cudaGetDevice(&dev));
  if (dev == cached_dev) return cached_rc;
  int major = 0, minor = 0;
  WSPC_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  WSPC_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
  cached_dev = dev;
  if (major != 10 || minor != 0) {
    set_error("device %d is sm_%d%d; libwspc is built for sm_100a only (no fallback)", dev, major, minor);
    cached_rc = WSPC_ERR_ARCH;
  } else {
    cached_rc = WSPC_OK;
  }
  return cached_rc;
}

}  // namespace wspc

extern "C" int wspc_version(void) { return 100; }
extern "C" const char* wspc_last_error(void) { return wspc::g_err; }
extern "C" uint64_t wspc_launch_count(void) { return wspc::g_launches.load(std::memory_order_relaxed); }
