// Library-wide state and the small informational entry points of the C ABI.
#include <stdlib.h>

#include "sia_host.cuh"
#include "sia_ptx.cuh"
#include "../../include/sia_b200_debug.h"

namespace sia {

static volatile unsigned int* g_wd_host = nullptr;

volatile unsigned int* watchdog_host_word() { return g_wd_host; }

int ensure_watchdog() {
  // one pinned host word per process; the device-side pointer to it is a per-DEVICE symbol
  static bool uploaded[kMaxDevices] = {false};
  const int dev = current_device();
  if (g_wd_host != nullptr && uploaded[dev]) return 0;
  if (g_wd_host == nullptr) {
    unsigned int* h = nullptr;
    cudaError_t e = cudaHostAlloc(reinterpret_cast<void**>(&h), sizeof(unsigned int),
                                  cudaHostAllocMapped | cudaHostAllocPortable);
    if (e != cudaSuccess) return (int)e;
    *h = 0;
    g_wd_host = h;
  }
  unsigned int* d = nullptr;
  cudaError_t e = cudaHostGetDevicePointer(reinterpret_cast<void**>(&d), const_cast<unsigned int*>(g_wd_host), 0);
  if (e != cudaSuccess) return (int)e;
  e = cudaMemcpyToSymbol(g_watchdog_word, &d, sizeof(d));
  if (e != cudaSuccess) return (int)e;
  uploaded[dev] = true;
  return 0;
}

// [B,h,w,C] bf16 -> [B,hp,wp,C] bf16: copies the valid hv x wv corner, zero everywhere else (16-byte vectors).
__global__ void pad_nhwc_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, int h, int w, int hp, int wp,
                                int hv, int wv, int c8, long long total) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int cc = (int)(i % c8);
    long long r = i / c8;
    const int x = (int)(r % wp);
    r /= wp;
    const int y = (int)(r % hp);
    const long long b = r / hp;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (y < hv && x < wv) v = in[((b * h + y) * w + x) * c8 + cc];
    out[i] = v;
  }
}

// [B,3,H,W] u8 (what a GPU JPEG decoder emits) -> [B,H,W,3] u8 (the decode-buffer layout of the transform kernels).
// Thread = 4 pixels of one row: three 32-bit plane loads, three 32-bit stores of 12 interleaved bytes.
__global__ void chw_to_hwc_u8_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int h, int w,
                                     long long quads_total) {
  const int wq = w >> 2;
  const size_t plane = (size_t)h * w;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < quads_total;
       i += (long long)gridDim.x * blockDim.x) {
    const int xq = (int)(i % wq);
    const long long row = i / wq;                       // b * h + y
    const long long b = row / h;
    const size_t in = (size_t)b * 3 * plane + (size_t)(row - b * h) * w + (size_t)xq * 4;
    const uint32_t r = *reinterpret_cast<const uint32_t*>(src + in);
    const uint32_t g = *reinterpret_cast<const uint32_t*>(src + in + plane);
    const uint32_t bl = *reinterpret_cast<const uint32_t*>(src + in + 2 * plane);
    uint32_t* out = reinterpret_cast<uint32_t*>(dst + ((size_t)row * w + (size_t)xq * 4) * 3);
    const uint32_t rg0 = __byte_perm(r, g, 0x5140), rg1 = __byte_perm(r, g, 0x7362);     // r0 g0 r1 g1 | r2 g2 r3 g3
    out[0] = __byte_perm(rg0, bl, 0x2410);                                   // r0 g0 b0 r1
    out[1] = __byte_perm(__byte_perm(rg0, bl, 0x0053), rg1, 0x5410);         // g1 b1 r2 g2
    out[2] = __byte_perm(rg1, bl, 0x7326);                                   // b2 r3 g3 b3
  }
}

static int g_pdl = -1;      // -1: not decided yet (reads SIA_PDL on first use)

bool pdl_enabled() {
  if (g_pdl < 0) {
    const char* env = getenv("SIA_PDL");
    g_pdl = (env != nullptr && env[0] == '0') ? 0 : 1;
  }
  return g_pdl != 0;
}
void set_pdl_enabled(bool on) { g_pdl = on ? 1 : 0; }

}  // namespace sia

#ifndef SIA_DEBUG_LIB
extern "C" int sia_debug_set_programmatic_launch(int on) {
  sia::set_pdl_enabled(on != 0);
  return 0;
}

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

int sia_pad_nhwc_bf16(const void* in_nhwc, int batch, int h, int w, int channels, int valid_h, int valid_w,
                      void* out_nhwc, int out_h, int out_w, void* stream) {
  using namespace sia;
  SIA_REQUIRE(in_nhwc && out_nhwc && batch >= 1 && h >= 1 && w >= 1 && out_h >= 1 && out_w >= 1);
  SIA_REQUIRE(valid_h >= 0 && valid_w >= 0 && valid_h <= h && valid_w <= w && valid_h <= out_h && valid_w <= out_w);
  SIA_REQUIRE(channels >= 8 && channels % 8 == 0 && aligned(in_nhwc, 16) && aligned(out_nhwc, 16));
  const int c8 = channels / 8;
  const long long total = (long long)batch * out_h * out_w * c8;
  const int blocks = (int)((total + 255) / 256 < 4096 ? (total + 255) / 256 : 4096);
  pad_nhwc_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint4*>(in_nhwc), static_cast<uint4*>(out_nhwc), h, w, out_h, out_w, valid_h, valid_w, c8,
      total);
  return launch_status();
}

int sia_chw_u8_to_hwc_u8(const uint8_t* src_chw, int batch, int h, int w, uint8_t* dst_hwc, void* stream) {
  using namespace sia;
  SIA_REQUIRE(src_chw && dst_hwc && batch >= 1 && h >= 1 && w >= 4);
  if (w % 4 != 0) return SIA_E_UNSUPPORTED;
  SIA_REQUIRE(aligned(src_chw, 4) && aligned(dst_hwc, 4));
  const long long quads = (long long)batch * h * (w / 4);
  const int blocks = (int)((quads + 255) / 256 < 8192 ? (quads + 255) / 256 : 8192);
  chw_to_hwc_u8_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(src_chw, dst_hwc, h, w, quads);
  return launch_status();
}

unsigned int sia_watchdog_status(int reset) {
  volatile unsigned int* w = sia::watchdog_host_word();
  if (w == nullptr) return 0;
  const unsigned int v = *w;
  if (reset) *w = 0;
  return v;
}

}  // extern "C"
#endif  // SIA_DEBUG_LIB
