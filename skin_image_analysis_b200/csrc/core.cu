// Library-wide state and the small informational entry points of the C ABI.
#include "sia_host.cuh"
#include "sia_ptx.cuh"

namespace sia {

static volatile unsigned int* g_wd_host = nullptr;

volatile unsigned int* watchdog_host_word() { return g_wd_host; }

int ensure_watchdog() {
  if (g_wd_host != nullptr) return 0;
  unsigned int* h = nullptr;
  cudaError_t e = cudaHostAlloc(reinterpret_cast<void**>(&h), sizeof(unsigned int), cudaHostAllocMapped);
  if (e != cudaSuccess) return (int)e;
  *h = 0;
  unsigned int* d = nullptr;
  e = cudaHostGetDevicePointer(reinterpret_cast<void**>(&d), h, 0);
  if (e != cudaSuccess) return (int)e;
  e = cudaMemcpyToSymbol(g_watchdog_word, &d, sizeof(d));
  if (e != cudaSuccess) return (int)e;
  g_wd_host = h;
  return 0;
}

}  // namespace sia

extern "C" int sia_debug_set_trace(long long* device_buffer_or_null) {
#ifndef SIA_INSTRUMENT
  if (device_buffer_or_null != nullptr) return SIA_E_UNSUPPORTED;   // needs a -DSIA_INSTRUMENT build
#endif
  cudaError_t e = cudaMemcpyToSymbol(sia::g_trace, &device_buffer_or_null, sizeof(device_buffer_or_null));
  return e == cudaSuccess ? 0 : (int)e;
}

extern "C" int sia_debug_set_stats(unsigned long long* device_buffer_or_null) {
#ifndef SIA_INSTRUMENT
  if (device_buffer_or_null != nullptr) return SIA_E_UNSUPPORTED;   // needs a -DSIA_INSTRUMENT build
#endif
  cudaError_t e = cudaMemcpyToSymbol(sia::g_stats, &device_buffer_or_null, sizeof(device_buffer_or_null));
  return e == cudaSuccess ? 0 : (int)e;
}

extern "C" {

int sia_version(void) { return SIA_VERSION; }

const char* sia_error_string(int code) {
  if (code == 0) return "ok";
  if (code > 0) return cudaGetErrorString((cudaError_t)code);
  switch (code) {
    case SIA_E_INVALID: return "sia: invalid argument";
    case SIA_E_UNSUPPORTED: return "sia: unsupported shape";
    case SIA_E_DRIVER: return "sia: cuTensorMapEncodeTiled unavailable or failed";
    case SIA_E_WATCHDOG: return "sia: kernel watchdog fired (mbarrier wait timed out)";
    default: return "sia: unknown error";
  }
}

int sia_device_info(int* sm_count_host, int* cc_major_host, int* cc_minor_host) {
  int dev = 0;
  SIA_CUDA_OK(cudaGetDevice(&dev));
  if (sm_count_host) SIA_CUDA_OK(cudaDeviceGetAttribute(sm_count_host, cudaDevAttrMultiProcessorCount, dev));
  if (cc_major_host) SIA_CUDA_OK(cudaDeviceGetAttribute(cc_major_host, cudaDevAttrComputeCapabilityMajor, dev));
  if (cc_minor_host) SIA_CUDA_OK(cudaDeviceGetAttribute(cc_minor_host, cudaDevAttrComputeCapabilityMinor, dev));
  return 0;
}

unsigned int sia_debug_watchdog(int reset) {
  volatile unsigned int* w = sia::watchdog_host_word();
  if (w == nullptr) return 0;
  const unsigned int v = *w;
  if (reset) *w = 0;
  return v;
}

}  // extern "C"
