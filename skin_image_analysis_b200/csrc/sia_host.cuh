// Host-side helpers shared by the C-ABI entry points: error plumbing and TMA tensor-map encoding
// (cuTensorMapEncodeTiled resolved at run time through the CUDA runtime, so the library has no
// link-time dependency on libcuda and loads on a box without a driver).
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/sia_b200.h"

namespace sia {

#define SIA_CUDA_OK(expr)                      \
  do {                                         \
    cudaError_t _e = (expr);                   \
    if (_e != cudaSuccess) return (int)_e;     \
  } while (0)

#define SIA_REQUIRE(cond) \
  do {                    \
    if (!(cond)) return SIA_E_INVALID; \
  } while (0)

// Sets up the pinned watchdog word on first use (per process); every launcher calls it.
int ensure_watchdog();
volatile unsigned int* watchdog_host_word();

inline int launch_status() {
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : (int)e;
}

inline bool aligned(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }

constexpr int kMaxDevices = 64;     // per-device caches below are indexed by the CUDA device ordinal

inline int current_device() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return 0;
  return dev;
}

inline int sm_count() {
  static int cached[kMaxDevices] = {0};
  const int dev = current_device();
  if (cached[dev] == 0) {
    if (cudaDeviceGetAttribute(&cached[dev], cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) cached[dev] = 148;
  }
  return cached[dev];
}

// Raises the dynamic shared-memory limit of a kernel once PER DEVICE (function attributes are per device): keeps
// cudaFuncSetAttribute out of CUDA-graph capture after the first, un-captured, warm-up launch on that device.
// `configured` is the call site's own static table, one slot per device ordinal.
struct SmemSlots { int bytes[kMaxDevices]; };
template <typename Kernel>
inline int ensure_dynamic_smem(Kernel kern, int bytes, SmemSlots* configured) {
  int& have = configured->bytes[current_device()];
  if (have < bytes) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e != cudaSuccess) return (int)e;
    have = bytes;
  }
  return 0;
}

// Whether the hot-path kernels are launched with programmatic dependent launch (default on; SIA_PDL=0 in the
// environment or sia_debug_set_programmatic_launch(0) switches it off for A/B timing).
bool pdl_enabled();
void set_pdl_enabled(bool on);

// Launch with (pdl = true) or without the programmatic-stream-serialization attribute.  A kernel launched with it MUST
// execute pdl_wait() before it reads anything an earlier kernel wrote or writes anything an earlier kernel reads.
template <typename... KArgs, typename... Args>
inline int launch_kernel(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl,
                         Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (pdl && pdl_enabled()) ? 1 : 0;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
  if (e != cudaSuccess) return (int)e;
  return launch_status();
}

// rank <= 5.  dims / box in elements (innermost first); strides in BYTES for dims 1..rank-1.
int encode_tmap_bf16(CUtensorMap* map, const void* base, int rank, const uint64_t* dims,
                     const uint64_t* strides_bytes, const uint32_t* box, CUtensorMapSwizzle swizzle);

// Same for any element type (dims / box in elements of that type).
int encode_tmap(CUtensorMap* map, CUtensorMapDataType dtype, const void* base, int rank, const uint64_t* dims,
                const uint64_t* strides_bytes, const uint32_t* box, CUtensorMapSwizzle swizzle);

}  // namespace sia
