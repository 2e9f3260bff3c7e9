// Library-level plumbing of libatq_sm100: error text, device selection, device checks.
#include <stdarg.h>
#include <atomic>
#include "common.cuh"

namespace atq {

static thread_local char g_err[512] = "";
static int g_sm_count[64] = {0};

static std::atomic<unsigned long long> g_launches{0};
void note_launch(int n) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }
unsigned long long launches() { return g_launches.load(std::memory_order_relaxed); }

char* last_error_buf() { return g_err; }

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int DeviceGuard::set(int device) {
  // The caller (PyTorch) owns the thread's current device and changes it behind this library's back (device guards,
  // torch.cuda.set_device, the autograd thread), so nothing is cached: ask, switch only when it differs, and put the
  // caller's device back when the entry point returns.
  if (device < 0 || device >= 64) {
    set_error("bad device index %d", device);
    return ATQ_EINVAL;
  }
  int cur = -1;
  cudaError_t e = cudaGetDevice(&cur);
  if (e == cudaSuccess && cur == device) return ATQ_OK;
  e = cudaSetDevice(device);
  if (e != cudaSuccess) {
    set_error("cudaSetDevice(%d) failed: %s", device, cudaGetErrorString(e));
    return ATQ_ECUDA;
  }
  prev = cur;
  return ATQ_OK;
}
DeviceGuard::~DeviceGuard() {
  if (prev >= 0) cudaSetDevice(prev);
}

int sm_count(int device) {
  if (device < 0 || device >= 64) return kNumSMsB200;
  if (g_sm_count[device] == 0) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || v <= 0) v = kNumSMsB200;
    g_sm_count[device] = v;
  }
  return g_sm_count[device];
}

}  // namespace atq

extern "C" {

int atq_abi_version(void) { return 6; }

const char* atq_last_error_string(void) { return atq::last_error_buf(); }

int atq_device_check(int device) {
  int major = 0, minor = 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device) != cudaSuccess ||
      cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, device) != cudaSuccess) {
    atq::set_error("atq_device_check: cannot query device %d", device);
    return ATQ_ECUDA;
  }
  if (major != 10) {
    atq::set_error("atq_device_check: device %d is sm_%d%d; this library is sm_100a only", device, major, minor);
    return ATQ_EARCH;
  }
  return ATQ_OK;
}

int atq_num_sms(int device) { return atq::sm_count(device); }

uint64_t atq_kernel_launch_count(void) { return (uint64_t)atq::launches(); }

}  // extern "C"
