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
  int kind;          // 0 = kind::f16 (bf16 inputs, fp32 accumulate), 1 = kind::i8 (s32 accumulate)
  int switch_every;  // > 0 (timing only): after this many MMAs move to the next of two accumulators, restarting it
  int commit_each;   // != 0 (timing only): tcgen05.commit to a scratch barrier at every accumulator switch
  uint32_t idesc;    // 0 = the default bf16 K-major descriptor for (128, n)
};

__global__ void __launch_bounds__(128, 1)
umma_probe_kernel(const uint8_t* __restrict__ image, const __grid_constant__ ProbeParams p, float* __restrict__ out,
                  long long* __restrict__ cycles) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ uint64_t done_bar;
  __shared__ uint64_t scratch_bar;
  __shared__ uint32_t tmem_slot;

  // 1024-byte aligned image base so that relative alignment == absolute alignment
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t base_addr = smem_u32(base);

  for (int i = threadIdx.x * 16; i < p.image_bytes; i += blockDim.x * 16) {
    *reinterpret_cast<uint4*>(base + i) = *reinterpret_cast<const uint4*>(image + i);
  }
  fence_proxy_async_smem();

  const uint32_t ncols = p.switch_every > 0 ? 512 : p.n <= 32 ? 32 : p.n <= 64 ? 64 : p.n <= 128 ? 128 : 256;
  if (threadIdx.x < 32) {
    tmem_alloc(&tmem_slot, ncols);
  }
  if (threadIdx.x == 32) {
    mbar_init(&done_bar, 1);
    mbar_init(&scratch_bar, 1 << 20);
    fence_mbar_init();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = bcast0(tmem_slot);

  const uint32_t idesc = p.idesc != 0u ? p.idesc : make_idesc_bf16(128, p.n);
  if (threadIdx.x < 32) {
    // warp-uniform control flow (descriptors come from the constant bank); one elected lane issues
    const uint32_t base16 = base_addr >> 4;
    long long t0 = 0;
    if (elect_one()) {
      t0 = clock64();
      if (p.switch_every > 0) {              // timing experiment: rotate over two accumulators every k MMAs
        // n_mma is a multiple of switch_every; chains alternate between the two accumulators
        uint32_t acc = 0;
        for (int rep = 0; rep < p.repeat; ++rep) {
          for (int i0 = 0; i0 < p.n_mma; i0 += p.switch_every) {
            for (int i = i0; i < i0 + p.switch_every; ++i) {
              umma_bf16_ss(tmem + acc, p.a_desc[i] + base16, p.b_desc[i] + base16, idesc, i > i0 ? 1u : 0u);
            }
            if (p.commit_each) umma_commit(&scratch_bar);
            acc ^= 256u;
          }
        }
      } else
      for (int rep = 0; rep < p.repeat; ++rep) {
        for (int i = 0; i < p.n_mma; ++i) {
          // start-address field is relative to the image base (no carry: the image is < 256 KB)
          if (p.kind == 1) {
            umma_i8_ss(tmem, p.a_desc[i] + base16, p.b_desc[i] + base16, idesc, i > 0 ? 1u : 0u);
          } else {
            umma_bf16_ss(tmem, p.a_desc[i] + base16, p.b_desc[i] + base16, idesc, i > 0 ? 1u : 0u);
          }
        }
      }
      umma_commit(&done_bar);
    }
    __syncwarp();
    mbar_wait(&done_bar, 0, 1);
    if (threadIdx.x == 0 && cycles) *cycles = clock64() - t0;
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

static int g_probe_switch_every = 0, g_probe_commit_each = 0;
extern "C" int sia_debug_umma_probe_switch(int switch_every, int commit_each) {
  g_probe_switch_every = switch_every;
  g_probe_commit_each = commit_each;
  return 0;
}

extern "C" int sia_debug_umma_probe(const void* smem_image, int image_bytes, const uint64_t* a_desc_host,
                                    const uint64_t* b_desc_host, int n_mma, int n, float* out_128xn, int repeat,
                                    long long* cycles_host, void* stream) {
  return sia_debug_umma_probe_ex(smem_image, image_bytes, a_desc_host, b_desc_host, n_mma, n, 0, 0u, out_128xn, repeat,
                                 cycles_host, stream);
}

extern "C" int sia_debug_umma_probe_ex(const void* smem_image, int image_bytes, const uint64_t* a_desc_host,
                                       const uint64_t* b_desc_host, int n_mma, int n, int kind, uint32_t idesc,
                                       void* out_128xn_raw, int repeat, long long* cycles_host, void* stream) {
  using namespace sia;
  float* out_128xn = static_cast<float*>(out_128xn_raw);
  SIA_REQUIRE(kind == 0 || kind == 1);
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
  p.kind = kind;
  p.idesc = idesc;
  p.switch_every = g_probe_switch_every;
  p.commit_each = g_probe_commit_each;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  long long* d_cycles = nullptr;
  if (cycles_host) SIA_CUDA_OK(cudaMalloc(&d_cycles, sizeof(long long)));
  const int smem = image_bytes + 1024;
  static SmemSlots configured = {};
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
  int repeat;          // > 1: throughput mode -- `repeat` loads in flight on one barrier, 4 rotating buffers
  int step_dim;        // coordinate advanced by step per repeat (so successive boxes differ)
  int step;
};

__global__ void __launch_bounds__(128, 1)
tma_probe_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ TmaProbeParams p,
                 uint8_t* __restrict__ out, long long* __restrict__ cycles) {
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
    const uint32_t stride = (p.box_bytes + 1023u) & ~1023u;
    const long long t0 = clock64();
    mbar_arrive_expect_tx(&bar, p.box_bytes * (uint32_t)p.repeat);
    for (int i = 0; i < p.repeat; ++i) {
      int c[5] = {p.coords[0], p.coords[1], p.coords[2], p.coords[3], p.coords[4]};
      c[p.step_dim] += i * p.step;
      uint8_t* dst = base + (p.repeat > 1 ? (uint32_t)(i & 3) * stride : 0u);
      if (p.rank == 2) tma_load_2d(dst, &tmap, &bar, c[0], c[1]);
      if (p.rank == 3) tma_load_3d(dst, &tmap, &bar, c[0], c[1], c[2]);
      if (p.rank == 4) tma_load_4d(dst, &tmap, &bar, c[0], c[1], c[2], c[3]);
    }
    mbar_wait(&bar, 0, 5);
    if (cycles) *cycles = clock64() - t0;
  }
  mbar_wait(&bar, 0, 5);
  for (uint32_t i = threadIdx.x * 16; i < p.box_bytes; i += blockDim.x * 16) {
    *reinterpret_cast<uint4*>(out + i) = *reinterpret_cast<const uint4*>(base + i);
  }
}

}  // namespace sia

extern "C" int sia_debug_tma_probe(const void* base, int rank, const uint64_t* dims_host,
                                   const uint64_t* strides_bytes_host, const uint32_t* box_host, int swizzle_bytes,
                                   const int* coords_host, void* out, int repeat, int step_dim, int step,
                                   long long* cycles_host, void* stream) {
  using namespace sia;
  SIA_REQUIRE(base && dims_host && strides_bytes_host && box_host && coords_host && out);
  SIA_REQUIRE(rank >= 2 && rank <= 4 && aligned(base, 16) && aligned(out, 16) && repeat >= 1);
  SIA_REQUIRE(step_dim >= 0 && step_dim < rank);
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
  p.repeat = repeat;
  p.step_dim = step_dim;
  p.step = step;
  uint64_t bytes = 2;
  for (int i = 0; i < 5; ++i) p.coords[i] = 0;
  for (int i = 0; i < rank; ++i) {
    p.coords[i] = coords_host[i];
    bytes *= box_host[i];
  }
  SIA_REQUIRE(bytes % 16 == 0 && bytes <= 48 * 1024 && bytes * (uint64_t)repeat < (1u << 20));
  p.box_bytes = (uint32_t)bytes;
  const int smem = 4 * (((int)bytes + 1023) & ~1023) + 1024;
  static SmemSlots configured = {};
  if (int rc2 = ensure_dynamic_smem(tma_probe_kernel, smem, &configured)) return rc2;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  long long* d_cycles = nullptr;
  if (cycles_host) SIA_CUDA_OK(cudaMalloc(&d_cycles, sizeof(long long)));
  tma_probe_kernel<<<1, 128, smem, st>>>(tmap, p, static_cast<uint8_t*>(out), d_cycles);
  rc = launch_status();
  if (rc == 0 && cycles_host) {
    cudaError_t e = cudaStreamSynchronize(st);
    if (e == cudaSuccess) e = cudaMemcpy(cycles_host, d_cycles, sizeof(long long), cudaMemcpyDeviceToHost);
    rc = (int)e;
  }
  if (d_cycles) cudaFree(d_cycles);
  return rc;
}

// ----------------------------------------------------------------------------------------------
// ALU issue-rate probe: lane-operations per SM clock for the instruction kinds the preprocess kernel
// can be built from (measured, because the B200 rates of I2F / IDP / packed-fp32 are not documented).
// ----------------------------------------------------------------------------------------------
namespace sia {

constexpr int RATE_KINDS = 8;
constexpr int RATE_ITERS = 512;
constexpr int RATE_CHAINS = 8;

template <int KIND>
__device__ __forceinline__ void rate_body(uint32_t (&r)[RATE_CHAINS], uint32_t s0, uint32_t s1) {
#pragma unroll
  for (int c = 0; c < RATE_CHAINS; ++c) {
    if constexpr (KIND == 0) {  // FFMA
      r[c] = __float_as_uint(fmaf(__uint_as_float(r[c]), __uint_as_float(s0), __uint_as_float(s1)));
    } else if constexpr (KIND == 1) {  // PRMT
      r[c] = __byte_perm(r[c], s0, 0x7650 + (c & 3));
    } else if constexpr (KIND == 2) {  // I2F from a byte
      r[c] = __float_as_uint((float)((r[c] >> 8) & 0xffu)) + s0;
    } else if constexpr (KIND == 3) {  // dp4a
      r[c] = __dp4a(r[c], s0, r[c]);
    } else if constexpr (KIND == 4) {  // dp2a
      r[c] = __dp2a_lo(s0, r[c], r[c]);
    } else if constexpr (KIND == 5) {  // IMAD
      r[c] = r[c] * s0 + s1;
    } else if constexpr (KIND == 6) {  // funnel shift
      r[c] = __funnelshift_r(r[c], s0, s1);
    }
  }
}

template <int KIND>
__device__ __forceinline__ long long rate_run(uint32_t seed, uint32_t* sink) {
  uint32_t r[RATE_CHAINS];
#pragma unroll
  for (int c = 0; c < RATE_CHAINS; ++c) r[c] = seed + c * 0x01010101u;
  const uint32_t s0 = seed | 0x3f000001u, s1 = (seed & 7u) + 1u;
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < RATE_ITERS; ++i) rate_body<KIND>(r, s0, s1);
  const long long t1 = clock64();
  uint32_t x = 0;
#pragma unroll
  for (int c = 0; c < RATE_CHAINS; ++c) x ^= r[c];
  if (x == 0x12345678u) *sink = x;
  __syncthreads();
  return t1 - t0;
}

__global__ void __launch_bounds__(1024, 1) alu_rate_kernel(long long* cycles, uint32_t seed, uint32_t* sink) {
  long long c[RATE_KINDS];
  c[0] = rate_run<0>(seed, sink);
  c[1] = rate_run<1>(seed, sink);
  c[2] = rate_run<2>(seed, sink);
  c[3] = rate_run<3>(seed, sink);
  c[4] = rate_run<4>(seed, sink);
  c[5] = rate_run<5>(seed, sink);
  c[6] = rate_run<6>(seed, sink);
  // packed fp32: 2 FMAs per lane-instruction
  {
    float2 r[RATE_CHAINS];
#pragma unroll
    for (int k = 0; k < RATE_CHAINS; ++k) r[k] = make_float2(1.0f + k, 2.0f + k);
    const float2 a = make_float2(1.0000001f, 0.9999999f), b = make_float2(1e-7f, -1e-7f);
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < RATE_ITERS; ++i) {
#pragma unroll
      for (int k = 0; k < RATE_CHAINS; ++k) r[k] = __ffma2_rn(r[k], a, b);
    }
    c[7] = clock64() - t0;
    float x = 0.f;
#pragma unroll
    for (int k = 0; k < RATE_CHAINS; ++k) x += r[k].x + r[k].y;
    if (x == 1234.5f) *sink = 1;
  }
  if (threadIdx.x == 0) {
    for (int k = 0; k < RATE_KINDS; ++k) cycles[k] = c[k];
  }
}

}  // namespace sia

// out_host[k] = lane-operations per SM clock for k = FFMA, PRMT, I2F.U8(+IADD), DP4A, DP2A, IMAD, SHF, FFMA2
extern "C" int sia_debug_alu_rates(double* out_host, int n) {
  using namespace sia;
  SIA_REQUIRE(out_host && n >= RATE_KINDS);
  long long* d = nullptr;
  uint32_t* sink = nullptr;
  SIA_CUDA_OK(cudaMalloc(&d, RATE_KINDS * sizeof(long long)));
  SIA_CUDA_OK(cudaMalloc(&sink, sizeof(uint32_t)));
  alu_rate_kernel<<<1, 1024>>>(d, 12345u, sink);
  long long h[RATE_KINDS];
  cudaError_t e = cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  cudaFree(d);
  cudaFree(sink);
  if (e != cudaSuccess) return (int)e;
  const double ops = 1024.0 * RATE_ITERS * RATE_CHAINS;
  for (int k = 0; k < RATE_KINDS; ++k) out_host[k] = ops / (double)h[k];
  return 0;
}

// ----------------------------------------------------------------------------------------------
// TMEM read-rate probe: bytes per SM clock of back-to-back tcgen05.ld for 1 / 4 / 8 warps and the
// x32 / x8 / x1 shapes (what bounds an epilogue that streams accumulator columns).
// ----------------------------------------------------------------------------------------------
namespace sia {

__global__ void __launch_bounds__(256, 1) tmem_rate_kernel(long long* cycles, int n_warps, int shape, uint32_t* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc(&slot, 512);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t base = slot + ((uint32_t)(32 * (warp & 3)) << 16);
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  if (warp < n_warps) {
    for (int it = 0; it < 64; ++it) {
      if (shape == 32) {
        uint32_t v[32];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          tmem_ld32(base + ((it * 8 + k) * 32) % 480, v);
          acc += v[0];
        }
      } else if (shape == 8) {
        uint32_t v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                       : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                       : "r"(base + ((it * 8 + k) * 9) % 500)
                       : "memory");
          acc += v[0];
        }
      } else {
        uint32_t v;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(base + ((it * 8 + k) * 9) % 500)
                       : "memory");
          acc += v;
        }
      }
    }
    tmem_ld_wait();
  }
  const long long t1 = clock64();
  if (acc == 0x12345u) *sink = acc;
  __syncthreads();
  if (threadIdx.x == 0) *cycles = t1 - t0;
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_free(slot, 512);
}

}  // namespace sia

// out_host[i] = bytes per SM clock for (warps, shape) in {(1,32), (4,32), (8,32), (4,8), (8,8), (4,1)}
extern "C" int sia_debug_tmem_ld_rates(double* out_host, int n) {
  using namespace sia;
  SIA_REQUIRE(out_host && n >= 6);
  long long* d = nullptr;
  uint32_t* sink = nullptr;
  SIA_CUDA_OK(cudaMalloc(&d, sizeof(long long)));
  SIA_CUDA_OK(cudaMalloc(&sink, sizeof(uint32_t)));
  const int warps[6] = {1, 4, 8, 4, 8, 4};
  const int shapes[6] = {32, 32, 32, 8, 8, 1};
  int rc = 0;
  for (int i = 0; i < 6 && rc == 0; ++i) {
    for (int rep = 0; rep < 2; ++rep) tmem_rate_kernel<<<1, 256>>>(d, warps[i], shapes[i], sink);
    long long h = 0;
    cudaError_t e = cudaMemcpy(&h, d, sizeof(h), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { rc = (int)e; break; }
    out_host[i] = (double)warps[i] * 64 * 8 * shapes[i] * 32 * 4 / (double)h;
  }
  cudaFree(d);
  cudaFree(sink);
  return rc;
}


// ------------------------------------------------------------------------------------------------------------
// Bring-up of tcgen05.mma with the A operand in TENSOR MEMORY (the form a second product on an accumulator needs:
// DESIGN.md section 5, "second tensor-core product").  Thread m writes row m of A -- a_cols 32-bit words, i.e.
// 2 * a_cols 16-bit elements -- to TMEM columns [0, a_cols) of lane m with tcgen05.st; one thread then issues
// n_mma UMMAs  D[128 x n] (+)= A[:, 16 i .. 16 i + 15] * B_i^T  with A addressed as TMEM column a_col_step * i and
// B_i through the caller's shared-memory descriptors; D (fp32, TMEM columns 256 ..) is read back.
namespace sia {

__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

struct TsProbeParams {
  uint64_t b_desc[32];
  int n_mma, n, a_cols, a_col_step, image_bytes;
  uint32_t idesc;
};

__global__ void __launch_bounds__(128, 1)
umma_ts_probe_kernel(const uint8_t* __restrict__ image, const uint32_t* __restrict__ a_words,
                     const __grid_constant__ TsProbeParams p, float* __restrict__ out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ uint64_t done_bar;
  __shared__ uint32_t tmem_slot;
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  for (int i = threadIdx.x; i < p.image_bytes / 4; i += blockDim.x)
    reinterpret_cast<uint32_t*>(smem)[i] = reinterpret_cast<const uint32_t*>(image)[i];
  fence_proxy_async_smem();
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc(&tmem_slot, 512);
  if (threadIdx.x == 32) {
    mbar_init(&done_bar, 1);
    fence_mbar_init();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = tmem_slot;
  const uint32_t lanes = tmem + ((uint32_t)(32 * warp) << 16);
  // A rows -> TMEM columns [0, a_cols)
  for (int c = 0; c < p.a_cols; c += 8) {
    uint32_t v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = a_words[(size_t)threadIdx.x * p.a_cols + c + j];
    tmem_st8(lanes + c, v);
  }
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  if (warp == 0) {
    if (elect_one()) {
      const uint64_t base16 = (uint64_t)((smem_u32(smem) >> 4) & 0x3FFF);
      const uint32_t idesc = p.idesc ? p.idesc : make_idesc_bf16(128, p.n);
      for (int i = 0; i < p.n_mma; ++i)
        umma_f16_ts(tmem + 256, tmem + (uint32_t)(p.a_col_step * i), p.b_desc[i] + base16, idesc, i > 0 ? 1u : 0u);
      umma_commit(&done_bar);
    }
    __syncwarp();
  }
  mbar_wait(&done_bar, 0, 60);
  tc_fence_after_sync();
  for (int c = 0; c < p.n; c += 8) {
    uint32_t v[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(lanes + 256 + c)
                 : "memory");
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 8; ++j) out[(size_t)threadIdx.x * p.n + c + j] = __uint_as_float(v[j]);
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_free(tmem, 512);
}

}  // namespace sia

extern "C" int sia_debug_umma_ts_probe(const void* smem_image, int image_bytes, const void* a_words, int a_cols,
                                       int a_col_step, const uint64_t* b_desc_host, int n_mma, int n, uint32_t idesc,
                                       float* out_128xn, void* stream) {
  using namespace sia;
  SIA_REQUIRE(smem_image && a_words && b_desc_host && out_128xn);
  SIA_REQUIRE(image_bytes > 0 && image_bytes % 16 == 0 && image_bytes <= 200 * 1024);
  SIA_REQUIRE(n_mma >= 1 && n_mma <= 32 && n >= 8 && n <= 256 && n % 8 == 0 && a_cols >= 8 && a_cols <= 256 && a_cols % 8 == 0);
  if (int wrc = ensure_watchdog()) return wrc;
  TsProbeParams p;
  for (int i = 0; i < n_mma; ++i) p.b_desc[i] = b_desc_host[i];
  p.n_mma = n_mma; p.n = n; p.a_cols = a_cols; p.a_col_step = a_col_step; p.image_bytes = image_bytes; p.idesc = idesc;
  const int smem = image_bytes + 1024;
  static SmemSlots configured = {};
  if (int rc = ensure_dynamic_smem(umma_ts_probe_kernel, smem, &configured)) return rc;
  umma_ts_probe_kernel<<<1, 128, smem, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint8_t*>(smem_image), static_cast<const uint32_t*>(a_words), p, out_128xn);
  return launch_status();
}
