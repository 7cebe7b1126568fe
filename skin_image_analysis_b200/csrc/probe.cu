// Bring-up harness: issue a caller-defined list of tcgen05.mma over a caller-defined shared-memory
// image and hand back the accumulator.  This is how the descriptor conventions used by conv1.cu /
// conv3x3.cu / linear.cu (swizzle atoms, LBO/SBO meaning, K-advance inside an atom, overlapping
// windows in the no-swizzle layout) are pinned on real hardware -- see tests/test_umma_probe.py.
#include "sia_host.cuh"
#include "sia_ptx.cuh"

namespace sia {

constexpr int PROBE_MAX_MMA = 64;

struct ProbeParams {
  uint64_t a_desc[PROBE_MAX_MMA];
  uint64_t b_desc[PROBE_MAX_MMA];
  int n_mma;
  int n;
  int image_bytes;
  int repeat;
};

__global__ void __launch_bounds__(128, 1)
umma_probe_kernel(const uint8_t* __restrict__ image, const __grid_constant__ ProbeParams p, float* __restrict__ out,
                  long long* __restrict__ cycles) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ uint64_t done_bar;
  __shared__ uint32_t tmem_slot;

  // 1024-byte aligned image base so that relative alignment == absolute alignment
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t base_addr = smem_u32(base);

  for (int i = threadIdx.x * 16; i < p.image_bytes; i += blockDim.x * 16) {
    *reinterpret_cast<uint4*>(base + i) = *reinterpret_cast<const uint4*>(image + i);
  }
  fence_proxy_async_smem();

  const uint32_t ncols = p.n <= 32 ? 32 : p.n <= 64 ? 64 : p.n <= 128 ? 128 : 256;
  if (threadIdx.x < 32) {
    tmem_alloc(&tmem_slot, ncols);
  }
  if (threadIdx.x == 32) {
    mbar_init(&done_bar, 1);
    fence_mbar_init();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = tmem_slot;

  const uint32_t idesc = make_idesc_bf16(128, p.n);
  if (threadIdx.x == 0) {
    long long t0 = clock64();
    for (int rep = 0; rep < p.repeat; ++rep) {
      for (int i = 0; i < p.n_mma; ++i) {
        // start-address field is relative: add the image base (16-byte units, 14 bits)
        uint64_t a = p.a_desc[i];
        uint64_t b = p.b_desc[i];
        a = (a & ~uint64_t(0x3fff)) | (((a & 0x3fff) + (base_addr >> 4)) & 0x3fff);
        b = (b & ~uint64_t(0x3fff)) | (((b & 0x3fff) + (base_addr >> 4)) & 0x3fff);
        umma_bf16_ss(tmem, a, b, idesc, i > 0 ? 1u : 0u);
      }
    }
    umma_commit(&done_bar);
    mbar_wait(&done_bar, 0, 1);
    long long t1 = clock64();
    if (cycles) *cycles = t1 - t0;
  }
  __syncthreads();
  mbar_wait(&done_bar, 0, 2);
  tc_fence_after_sync();

  const uint32_t warp = threadIdx.x >> 5;
  const uint32_t row = threadIdx.x;  // TMEM lane == accumulator row
  for (int c0 = 0; c0 < p.n; c0 += 32) {
    uint32_t v[32];
    tmem_ld32(tmem + ((warp * 32u) << 16) + (uint32_t)c0, v);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      if (c0 + j < p.n) out[(size_t)row * p.n + c0 + j] = __uint_as_float(v[j]);
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (threadIdx.x < 32) tmem_free(tmem, ncols);
}

}  // namespace sia

extern "C" int sia_debug_umma_probe(const void* smem_image, int image_bytes, const uint64_t* a_desc_host,
                                    const uint64_t* b_desc_host, int n_mma, int n, float* out_128xn, int repeat,
                                    long long* cycles_host, void* stream) {
  using namespace sia;
  SIA_REQUIRE(smem_image && a_desc_host && b_desc_host && out_128xn);
  SIA_REQUIRE(n_mma >= 1 && n_mma <= PROBE_MAX_MMA && n >= 16 && n <= 256 && n % 16 == 0);
  SIA_REQUIRE(image_bytes > 0 && image_bytes % 16 == 0 && image_bytes <= 200 * 1024);
  SIA_REQUIRE(aligned(smem_image, 16) && repeat >= 1);
  if (int wrc = ensure_watchdog()) return wrc;
  ProbeParams p;
  for (int i = 0; i < n_mma; ++i) {
    p.a_desc[i] = a_desc_host[i];
    p.b_desc[i] = b_desc_host[i];
  }
  p.n_mma = n_mma;
  p.n = n;
  p.image_bytes = image_bytes;
  p.repeat = repeat;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  long long* d_cycles = nullptr;
  if (cycles_host) SIA_CUDA_OK(cudaMalloc(&d_cycles, sizeof(long long)));
  const int smem = image_bytes + 1024;
  static int configured = 0;
  if (int rc2 = ensure_dynamic_smem(umma_probe_kernel, smem, &configured)) return rc2;
  umma_probe_kernel<<<1, 128, smem, st>>>(static_cast<const uint8_t*>(smem_image), p, out_128xn, d_cycles);
  int rc = launch_status();
  if (rc == 0 && cycles_host) {
    cudaError_t e = cudaStreamSynchronize(st);
    if (e == cudaSuccess) e = cudaMemcpy(cycles_host, d_cycles, sizeof(long long), cudaMemcpyDeviceToHost);
    rc = (int)e;
  }
  if (d_cycles) cudaFree(d_cycles);
  return rc;
}

// ----------------------------------------------------------------------------------------------
// TMA bring-up: load one box of a bf16 tensor with the given swizzle / coordinates (negative and
// out-of-range coordinates included) and return the shared-memory bytes exactly as TMA wrote them.
// ----------------------------------------------------------------------------------------------
namespace sia {

struct TmaProbeParams {
  int rank;
  int coords[5];
  uint32_t box_bytes;
};

__global__ void __launch_bounds__(128, 1)
tma_probe_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ TmaProbeParams p,
                 uint8_t* __restrict__ out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ uint64_t bar;
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  for (uint32_t i = threadIdx.x * 16; i < p.box_bytes; i += blockDim.x * 16) {
    *reinterpret_cast<uint4*>(base + i) = make_uint4(0xdeadbeefu, 0xdeadbeefu, 0xdeadbeefu, 0xdeadbeefu);
  }
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  fence_proxy_async_smem();
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(&bar, p.box_bytes);
    if (p.rank == 2) tma_load_2d(base, &tmap, &bar, p.coords[0], p.coords[1]);
    if (p.rank == 3) tma_load_3d(base, &tmap, &bar, p.coords[0], p.coords[1], p.coords[2]);
    if (p.rank == 4) tma_load_4d(base, &tmap, &bar, p.coords[0], p.coords[1], p.coords[2], p.coords[3]);
  }
  mbar_wait(&bar, 0, 5);
  for (uint32_t i = threadIdx.x * 16; i < p.box_bytes; i += blockDim.x * 16) {
    *reinterpret_cast<uint4*>(out + i) = *reinterpret_cast<const uint4*>(base + i);
  }
}

}  // namespace sia

extern "C" int sia_debug_tma_probe(const void* base, int rank, const uint64_t* dims_host,
                                   const uint64_t* strides_bytes_host, const uint32_t* box_host, int swizzle_bytes,
                                   const int* coords_host, void* out, void* stream) {
  using namespace sia;
  SIA_REQUIRE(base && dims_host && strides_bytes_host && box_host && coords_host && out);
  SIA_REQUIRE(rank >= 2 && rank <= 4 && aligned(base, 16) && aligned(out, 16));
  if (int wrc = ensure_watchdog()) return wrc;
  CUtensorMapSwizzle sw = swizzle_bytes == 128  ? CU_TENSOR_MAP_SWIZZLE_128B
                          : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                          : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B
                                                : CU_TENSOR_MAP_SWIZZLE_NONE;
  CUtensorMap tmap;
  int rc = encode_tmap_bf16(&tmap, base, rank, dims_host, strides_bytes_host, box_host, sw);
  if (rc != 0) return rc;
  TmaProbeParams p;
  p.rank = rank;
  uint64_t bytes = 2;
  for (int i = 0; i < rank; ++i) {
    p.coords[i] = coords_host[i];
    bytes *= box_host[i];
  }
  SIA_REQUIRE(bytes % 16 == 0 && bytes <= 200 * 1024);
  p.box_bytes = (uint32_t)bytes;
  const int smem = (int)bytes + 1024;
  static int configured = 0;
  if (int rc2 = ensure_dynamic_smem(tma_probe_kernel, smem, &configured)) return rc2;
  tma_probe_kernel<<<1, 128, smem, static_cast<cudaStream_t>(stream)>>>(tmap, p, static_cast<uint8_t*>(out));
  return launch_status();
}
