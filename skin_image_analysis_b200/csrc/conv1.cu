// K4 (first block): Conv2d(3, 32, 7, stride 1, 'same') + bias + ReLU + MaxPool2d(2,2) on tcgen05.
// Replaces tone_bias_model.py:83-92 (layers.0) / :169-172 (conv1).
//
// Cin = 3 makes the textbook implicit GEMM hopeless (K = 147, N = 32).  Two tricks instead:
//
// 1. Input is NHWC4 bf16 (8 bytes / pixel, written directly by the preprocess kernel).  One GEMM
//    row computes TWO horizontally adjacent output pixels, so N = 2*32 = 64 and, per filter row r,
//    K = the 8-pixel input window both pixels need = 8 px * 4 ch = 32.  The weight operand becomes
//    a small block-Toeplitz matrix B[p*32+co][r*32 + xw*4 + c] = W[co][c][r][xw-p] (zero outside).
// 2. In the NO-SWIZZLE K-major UMMA layout a core matrix is 8 rows x 16 bytes with rows 16 bytes
//    apart -- exactly the distance between the windows of neighbouring pixel pairs.  So with
//    LBO = 16 B (next 2 pixels of the window) and SBO = the smem row pitch (next image row) the
//    tensor core reads the overlapping windows straight from the raw image patch:
//    NO im2col expansion exists anywhere, the A tile for 256 output pixels is 22 x 24 pixels = 4 KB,
//    fetched by one TMA box (zero fill outside the image = 'same' padding).
//
// M-tile = 16 rows x 8 pixel pairs (16 x 16 outputs); per tile 7 filter rows x 2 UMMAs of K = 16.
// The 2x2 max-pool is x: in-thread (the two pixel halves of the accumulator row), y: lane ^ 8.
#include "sia_host.cuh"
#include "sia_ptx.cuh"

namespace sia {

constexpr int C1_TILE = 16;                  // output tile is 16 x 16
constexpr int C1_WIN_PX = C1_TILE + 8;       // 24 input pixels per smem row (3 left, 5 right)
constexpr int C1_ROWB = C1_WIN_PX * 8;       // 192 bytes
constexpr int C1_ROWS = C1_TILE + 6;         // 22 input rows
constexpr int C1_STAGE_BYTES = C1_ROWS * C1_ROWB;        // 4224
constexpr int C1_STAGE_STRIDE = 4352;                    // 17 * 256
constexpr int C1_NSTAGE = 8;
constexpr int C1_N = 64;
constexpr int C1_K = 224;                    // 7 rows * 32
constexpr int C1_B_BYTES = C1_N * C1_K * 2;  // 28672
constexpr int C1_B_SBO = (C1_K / 8) * 128;   // 3584: next 8 columns-of-N group
constexpr int C1_THREADS = 256;

__global__ void __launch_bounds__(C1_THREADS, 1)
conv1_kernel(const __grid_constant__ CUtensorMap tmap_in, const uint8_t* __restrict__ w_packed,
             const float* __restrict__ bias, __nv_bfloat16* __restrict__ out, int H, int W, int tiles_y, int tiles_x,
             int total_tiles) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_b = smem;                                   // 28672
  uint8_t* smem_a = smem + C1_B_BYTES;                      // NSTAGE * STRIDE
  float* smem_bias = reinterpret_cast<float*>(smem_a + C1_NSTAGE * C1_STAGE_STRIDE);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_bias + 32);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + C1_NSTAGE;
  uint64_t* tfull_bar = bars + 2 * C1_NSTAGE;
  uint64_t* tempty_bar = bars + 2 * C1_NSTAGE + 2;
  uint64_t* wload_bar = bars + 2 * C1_NSTAGE + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * C1_NSTAGE + 5);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < C1_NSTAGE; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 4);
    }
    mbar_init(wload_bar, 1);
    fence_mbar_init();
    tma_prefetch_desc(&tmap_in);
  }
  if (warp == 2) tmem_alloc(tmem_slot, 2 * C1_N);
  if (threadIdx.x < 32) smem_bias[threadIdx.x] = bias[threadIdx.x];
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = bcast0(*tmem_slot);

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(wload_bar, C1_B_BYTES);
      bulk_load_1d(smem_b, w_packed, C1_B_BYTES / 2, wload_bar);
      bulk_load_1d(smem_b + C1_B_BYTES / 2, w_packed + C1_B_BYTES / 2, C1_B_BYTES / 2, wload_bar);
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int tx = tile % tiles_x;
        const int ty = (tile / tiles_x) % tiles_y;
        const int n = tile / (tiles_x * tiles_y);
        mbar_wait(&empty_bar[stage], phase ^ 1, 30);
        mbar_arrive_expect_tx(&full_bar[stage], C1_STAGE_BYTES);
        // innermost coordinate is in bf16 elements (4 per pixel) and must be 16-byte aligned for TMA:
        // image pixel x sits in column x+1 of the padded row, so the window start x0-3 is column x0-2
        tma_load_3d(smem_a + stage * C1_STAGE_STRIDE, &tmap_in, &full_bar[stage], (tx * C1_TILE - 2) * 4,
                    ty * C1_TILE - 3, n);
        if (++stage == C1_NSTAGE) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // whole warp runs the uniform control flow; one elected lane issues UMMAs + commits
    constexpr uint32_t idesc = make_idesc_bf16(128, C1_N);
    // A: rows = pixel pairs 16 B apart, K-adjacent core matrix = next 2 pixels (LBO 16 B), next 8 rows =
    //    next image row (SBO = row pitch).  B: canonical no-swizzle, core matrices contiguous along K.
    constexpr uint32_t a_hi = desc_hi(C1_ROWB, SW_NONE);
    constexpr uint32_t b_hi = desc_hi(C1_B_SBO, SW_NONE);
    const uint32_t a_lo0 = desc_lo(smem_u32(smem_a), 16);
    const uint32_t b_lo0 = desc_lo(smem_u32(smem_b), 128);
    mbar_wait(wload_bar, 0, 31);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      mbar_wait(&tempty_bar[acc], acc_phase ^ 1, 32);
      mbar_wait(&full_bar[stage], phase, 33);
      tc_fence_after_sync();
      if (elect_one()) {
        const uint32_t d_tmem = tmem_base + acc * C1_N;
        const uint32_t a_lo = a_lo0 + stage * (C1_STAGE_STRIDE >> 4);
#pragma unroll
        for (int r = 0; r < 7; ++r) {
#pragma unroll
          for (int kk = 0; kk < 2; ++kk) {
            umma_bf16_ss_w(d_tmem, a_lo + r * (C1_ROWB >> 4) + kk * 2, a_hi, b_lo0 + (r * 4 + kk * 2) * 8, b_hi, idesc,
                           (r | kk) ? 1u : 0u);
          }
        }
        umma_commit(&empty_bar[stage]);
        umma_commit(&tfull_bar[acc]);
      }
      __syncwarp();
      if (++stage == C1_NSTAGE) { stage = 0; phase ^= 1; }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else if (warp >= 4) {
    const int e = warp - 4;
    const int Ho = H >> 1, Wo = W >> 1;
    const int ly = lane >> 3;
    const int xp = lane & 7;
    const bool odd_y = (lane >> 3) & 1;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int tx = tile % tiles_x;
      const int ty = (tile / tiles_x) % tiles_y;
      const int n = tile / (tiles_x * tiles_y);
      const int py = (ty * C1_TILE + 4 * e + ly) >> 1;
      const int px = tx * (C1_TILE / 2) + xp;
      __nv_bfloat16* opix = out + (((size_t)n * Ho + py) * Wo + px) * 32;
      mbar_wait(&tfull_bar[acc], acc_phase, 34);
      tc_fence_after_sync();
      const uint32_t t_addr = tmem_base + ((uint32_t)(32 * e) << 16) + acc * C1_N;
      uint32_t v0[32], v1[32];
      tmem_ld32(t_addr, v0);        // left pixel of the pair, 32 channels
      tmem_ld32(t_addr + 32, v1);   // right pixel
      tmem_ld_wait();
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);   // accumulator is in registers now
      uint32_t pk[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float a = fmaxf(fmaxf(__uint_as_float(v0[2 * j]), __uint_as_float(v1[2 * j])) + smem_bias[2 * j], 0.f);
        const float b =
            fmaxf(fmaxf(__uint_as_float(v0[2 * j + 1]), __uint_as_float(v1[2 * j + 1])) + smem_bias[2 * j + 1], 0.f);
        pk[j] = pack_bf16x2(a, b);
      }
      uint32_t h8[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const uint32_t keep = odd_y ? pk[8 + j] : pk[j];
        const uint32_t send = odd_y ? pk[j] : pk[8 + j];
        h8[j] = max_bf16x2(keep, __shfl_xor_sync(0xffffffffu, send, 8));
      }
      if (py < Ho && px < Wo) {
        uint4* d = reinterpret_cast<uint4*>(opix + (odd_y ? 16 : 0));
        d[0] = make_uint4(h8[0], h8[1], h8[2], h8[3]);
        d[1] = make_uint4(h8[4], h8[5], h8[6], h8[7]);
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 2) tmem_free(tmem_base, 2 * C1_N);
}

// [32][3][7][7] fp32 -> B[n = p*32+co][k = r*32 + xw*4 + c] bf16 in no-swizzle core-matrix order:
// byte offset = (n/8)*3584 + (k/8)*128 + (n%8)*16 + (k%8)*2.
__global__ void pack_conv1_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ dst) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < C1_N * C1_K; i += gridDim.x * blockDim.x) {
    const int n = i / C1_K, k = i % C1_K;
    const int p = n / 32, co = n % 32;
    const int r = k / 32, xw = (k % 32) / 4, c = k % 4;
    const int t = xw - p;
    float v = 0.f;
    if (c < 3 && t >= 0 && t < 7) v = w[((co * 3 + c) * 7 + r) * 7 + t];
    const int off = (n / 8) * (C1_B_SBO / 2) + (k / 8) * 64 + (n % 8) * 8 + (k % 8);
    dst[off] = __float2bfloat16_rn(v);
  }
}

}  // namespace sia

extern "C" size_t sia_pack_conv7x7_c3_bytes(void) { return sia::C1_B_BYTES; }

extern "C" int sia_pack_conv7x7_c3(const float* w_oihw, void* packed, void* stream) {
  using namespace sia;
  SIA_REQUIRE(w_oihw && packed && aligned(packed, 16));
  pack_conv1_kernel<<<(C1_N * C1_K + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      w_oihw, static_cast<__nv_bfloat16*>(packed));
  return launch_status();
}

extern "C" int sia_conv7x7_c3_relu_pool2(const void* in_nhwc4, int batch, int h, int w, const void* w_packed,
                                         const float* bias, void* out_nhwc, void* stream) {
  using namespace sia;
  SIA_REQUIRE(in_nhwc4 && w_packed && bias && out_nhwc && batch >= 1 && h >= 16 && w >= 16);
  SIA_REQUIRE(aligned(in_nhwc4, 16) && aligned(w_packed, 16) && aligned(out_nhwc, 16));
  if (h % C1_TILE != 0 || w % C1_TILE != 0) return SIA_E_UNSUPPORTED;
  if (int wrc = ensure_watchdog()) return wrc;
  CUtensorMap tmap;
  // padded NHWC4 rows: (w + 8) pixels of 8 bytes, image pixel x in column x + 1
  const uint64_t wp = (uint64_t)w + SIA_NHWC4_PAD;
  const uint64_t dims[3] = {wp * 4, (uint64_t)h, (uint64_t)batch};
  const uint64_t strides[2] = {wp * 8, (uint64_t)h * wp * 8};
  const uint32_t box[3] = {C1_WIN_PX * 4, C1_ROWS, 1};
  int rc = encode_tmap_bf16(&tmap, in_nhwc4, 3, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE);
  if (rc != 0) return rc;
  const int tiles_y = h / C1_TILE, tiles_x = w / C1_TILE;
  const int total = tiles_y * tiles_x * batch;
  const int smem = 1024 + C1_B_BYTES + C1_NSTAGE * C1_STAGE_STRIDE + 32 * 4 + (2 * C1_NSTAGE + 6) * 8;
  static int configured = 0;
  if (int rc2 = ensure_dynamic_smem(conv1_kernel, smem, &configured)) return rc2;
  const int grid = total < sm_count() ? total : sm_count();
  conv1_kernel<<<grid, C1_THREADS, smem, static_cast<cudaStream_t>(stream)>>>(
      tmap, static_cast<const uint8_t*>(w_packed), bias, static_cast<__nv_bfloat16*>(out_nhwc), h, w, tiles_y,
      tiles_x, total);
  return launch_status();
}
