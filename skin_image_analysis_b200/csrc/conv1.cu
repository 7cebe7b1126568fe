// K4 (first block): Conv2d(3, 32, 7, stride 1, 'same') + bias + ReLU + MaxPool2d(2,2) on tcgen05.
// Replaces tone_bias_model.py:83-92 (layers.0) / :169-172 (conv1).
//
// Cin = 3 makes the textbook implicit GEMM hopeless (K = 147, N = 32).  Two tricks instead:
//
// 1. One GEMM row computes a 2 x 2 block of output pixels (= one max-pool window), so N = 4*32 = 128
//    and K = the 8 x 8-pixel input window the block needs = 8 rows * (8 px * 4 ch) = 256.  The weight
//    operand becomes a block-Toeplitz matrix
//        B[(dy*2+dx)*32 + co][r'*32 + xw*4 + c] = W[co][c][r'-dy][xw-dx]        (zero outside 0..6)
//    57 % of the issued MACs are useful -- far better than padding K = 147 / N = 32 to MMA shapes --
//    and the max-pool becomes a max over four column groups of the SAME accumulator row: no shuffles.
// 2. Input is padded NHWC4 bf16 (8 bytes / pixel, written directly by the preprocess kernel).  In the
//    NO-SWIZZLE K-major UMMA layout a core matrix is 8 rows x 16 bytes with rows 16 bytes apart --
//    exactly the distance between the windows of neighbouring pixel PAIRS.  With LBO = 16 B (next two
//    pixels of the window) and SBO = two image rows (next row pair) the tensor core reads the
//    overlapping windows straight from the raw image patch: NO im2col expansion exists anywhere; the
//    A tile for 512 output pixels is 38 x 24 pixels = 7.1 KB, fetched by one TMA box (zero fill
//    outside the image = 'same' padding).
//
// M-tile = 16 row pairs x 8 pixel pairs (32 x 16 outputs -> 16 x 8 pooled); 8 window rows x 2 UMMAs of
// K = 16 per tile, N = 128.  Tensor-core bound (operand fetch: 8 KB per UMMA = 128 B/clk).
#include "sia_host.cuh"
#include "sia_ptx.cuh"

namespace sia {

constexpr int C1_TILE_X = 16;                // output columns per tile (8 pairs)
constexpr int C1_TILE_Y = 32;                // output rows per tile (16 pairs)
constexpr int C1_WIN_PX = C1_TILE_X + 8;     // 24 input pixels per smem row (3 left, 5 right)
constexpr int C1_ROWB = C1_WIN_PX * 8;       // 192 bytes
constexpr int C1_ROWS = C1_TILE_Y + 6;       // 38 input rows
constexpr int C1_STAGE_BYTES = C1_ROWS * C1_ROWB;        // 7296
constexpr int C1_STAGE_STRIDE = 7424;                    // 29 * 256
#ifndef SIA_C1_NSTAGE
#define SIA_C1_NSTAGE 4
#endif
constexpr int C1_NSTAGE = SIA_C1_NSTAGE;     // input-patch stages in flight (7.3 KB each)
constexpr int C1_N = 128;
constexpr int C1_K = 256;                    // 8 window rows * 32
constexpr int C1_B_BYTES = C1_N * C1_K * 2;  // 65536
constexpr int C1_B_SBO = (C1_K / 8) * 128;   // 4096: next group of 8 B rows
#ifndef SIA_C1_NACC
#define SIA_C1_NACC 4
#endif
#ifndef SIA_C1_EPI_GROUPS
#define SIA_C1_EPI_GROUPS 2
#endif
#ifndef SIA_C1_BIAS_UMMA
#define SIA_C1_BIAS_UMMA 1
#endif
constexpr int C1_NACC = SIA_C1_NACC;         // TMEM accumulator ring (NACC x 128 columns)
constexpr int C1_EPI_GROUPS = SIA_C1_EPI_GROUPS;
constexpr int C1_THREADS = 128 + 128 * C1_EPI_GROUPS;  // warps 0-3 TMA / MMA / TMEM alloc / idle, then epilogue groups
constexpr bool C1_BIAS_UMMA = SIA_C1_BIAS_UMMA != 0;
#ifndef SIA_C1_MMA_WARPS
#define SIA_C1_MMA_WARPS 3
#endif
constexpr int C1_MMA_WARPS = SIA_C1_MMA_WARPS;  // warps issuing UMMAs, tiles dealt round-robin: warps 1 .. C1_MMA_WARPS (<= 3)
constexpr int C1_BIAS_BYTES = C1_N * 32;

struct TileWalker1 {
  int tx, ty, n, dtx, dty, dn, tiles_x, tiles_y;
  __device__ TileWalker1(int first, int step, int tiles_x_, int tiles_y_) : tiles_x(tiles_x_), tiles_y(tiles_y_) {
    tx = first % tiles_x;
    ty = (first / tiles_x) % tiles_y;
    n = first / (tiles_x * tiles_y);
    dtx = step % tiles_x;
    dty = (step / tiles_x) % tiles_y;
    dn = step / (tiles_x * tiles_y);
  }
  __device__ __forceinline__ void next() {
    tx += dtx;
    ty += dty;
    n += dn;
    if (tx >= tiles_x) { tx -= tiles_x; ++ty; }
    if (ty >= tiles_y) { ty -= tiles_y; ++n; }
  }
};

// 32 lanes x 16 consecutive fp32 columns
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}

__global__ void __launch_bounds__(C1_THREADS, 1)
conv1_kernel(const __grid_constant__ CUtensorMap tmap_in, const uint8_t* __restrict__ w_packed,
             const float* __restrict__ bias, __nv_bfloat16* __restrict__ out, int H, int W, int tiles_y, int tiles_x,
             int total_tiles, int c_stride) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* smem_b = smem;                                   // 65536
  uint8_t* smem_a = smem + C1_B_BYTES;                      // NSTAGE * STRIDE
  uint8_t* smem_ones = smem_a + C1_NSTAGE * C1_STAGE_STRIDE;
  uint8_t* smem_biasop = smem_ones + ONES_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_biasop + C1_BIAS_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + C1_NSTAGE;
  uint64_t* tfull_bar = bars + 2 * C1_NSTAGE;
  uint64_t* tempty_bar = bars + 2 * C1_NSTAGE + C1_NACC;
  uint64_t* wload_bar = bars + 2 * C1_NSTAGE + 2 * C1_NACC;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * C1_NSTAGE + 2 * C1_NACC + 1);
  volatile uint32_t* issued = tmem_slot + 1;            // [C1_MMA_WARPS] tiles issued per issuing warp (issue_gate)

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < C1_NSTAGE; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < C1_NACC; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 4);
    }
    mbar_init(wload_bar, 1);
    for (int i = 0; i < C1_MMA_WARPS; ++i) issued[i] = 0;
    fence_mbar_init();
    tma_prefetch_desc(&tmap_in);
  }
  if (warp == 2) tmem_alloc(tmem_slot, C1_NACC * C1_N);
  // the bias enters through the tensor core (see conv3x3.cu): row n = (dy*2+dx)*32 + co -> bias[co]
  if (C1_BIAS_UMMA) {
    fill_ones_operand(smem_ones, threadIdx.x, blockDim.x);
    fill_bias_operand(smem_biasop, bias, C1_N, 32, threadIdx.x, blockDim.x);
  } else if (threadIdx.x < 32) {
    reinterpret_cast<float*>(smem_biasop)[threadIdx.x] = bias[threadIdx.x];
  }
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = bcast0(*tmem_slot);
  const float* smem_bias = reinterpret_cast<const float*>(smem_biasop);
  // Programmatic dependent launch: everything above (and the resident weights below) overlaps the previous kernel's
  // tail; nothing that kernel wrote is read, and nothing it reads is written, before pdl_wait().
  pdl_launch_dependents();
  if (!(warp == 0 && lane == 0)) pdl_wait();

#ifdef SIA_C1_RAW_ISSUE
  // timing experiment: only the MMA warp runs, issuing every tile's UMMAs back to back without any barrier traffic
  if (warp == 1) {
    constexpr uint32_t idesc = make_idesc_bf16(128, C1_N);
    constexpr uint32_t a_hi = desc_hi(2 * C1_ROWB, SW_NONE);
    constexpr uint32_t b_hi = desc_hi(C1_B_SBO, SW_NONE);
    const uint32_t a_lo0 = desc_lo(smem_u32(smem_a), 16);
    const uint32_t b_lo0 = desc_lo(smem_u32(smem_b), 128);
    int stage = 0, acc = 0;
    uint32_t sphase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
#if SIA_C1_RAW_ISSUE == 3 || SIA_C1_RAW_ISSUE == 4 || SIA_C1_RAW_ISSUE == 6
      // a barrier that is always already complete: the cost of the wait (+ fence) themselves
      if (lane < 4) mbar_arrive(&tempty_bar[0]);
      mbar_wait(&tempty_bar[0], sphase, 38);
#if SIA_C1_RAW_ISSUE == 4
      if (lane == 0) mbar_arrive(&full_bar[0]);
      mbar_wait(&full_bar[0], sphase, 38);
#endif
      sphase ^= 1;
#endif
#if SIA_C1_RAW_ISSUE == 3 || SIA_C1_RAW_ISSUE == 4 || SIA_C1_RAW_ISSUE == 5
      tc_fence_after_sync();
#endif
      if (elect_one()) {
        const uint32_t d_tmem = tmem_base + acc * C1_N;
        const uint32_t a_lo = a_lo0 + stage * (C1_STAGE_STRIDE >> 4);
#pragma unroll
        for (int r = 0; r < 8; ++r) {
#pragma unroll
          for (int kk = 0; kk < 2; ++kk) {
            umma_bf16_ss_w(d_tmem, a_lo + r * (C1_ROWB >> 4) + kk * 2, a_hi, b_lo0 + (r * 4 + kk * 2) * 8, b_hi, idesc,
                           (r | kk) ? 1u : 0u);
          }
        }
#if SIA_C1_RAW_ISSUE >= 2
        umma_commit(&empty_bar[stage]);
        umma_commit(&tfull_bar[acc]);
#endif
      }
      __syncwarp();
      if (++stage == C1_NSTAGE) stage = 0;
      if (++acc == C1_NACC) acc = 0;
    }
    if (elect_one()) umma_commit(wload_bar);
    __syncwarp();
    mbar_wait(wload_bar, 0, 39);
  }
  if (warp < 64) goto c1_done;
#endif
  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(wload_bar, C1_B_BYTES);
      for (int off = 0; off < C1_B_BYTES; off += 16384) bulk_load_1d(smem_b + off, w_packed + off, 16384, wload_bar);
      pdl_wait();                       // the input tiles below are the previous kernel's output
      int stage = 0;
      uint32_t phase = 0;
      TileWalker1 t(blockIdx.x, gridDim.x, tiles_x, tiles_y);
      RoleTimer wait_stage;
      int lt = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, t.next(), ++lt) {
        wait_stage.begin();
        mbar_wait(&empty_bar[stage], phase ^ 1, 30);
        wait_stage.end();
        trace(lt, 0);
#ifdef SIA_C1_NO_TMA
        mbar_arrive(&full_bar[stage]);                    // timing experiment: operands are whatever is in smem
#else
        mbar_arrive_expect_tx(&full_bar[stage], C1_STAGE_BYTES);
        // innermost coordinate is in bf16 elements (4 per pixel) and must be 16-byte aligned for TMA:
        // image pixel x sits in column x+1 of the padded row, so the window start x0-3 is column x0-2
        tma_load_3d(smem_a + stage * C1_STAGE_STRIDE, &tmap_in, &full_bar[stage], (t.tx * C1_TILE_X - 2) * 4,
                    t.ty * C1_TILE_Y - 3, t.n);
#endif
        trace(lt, 1);
        if (++stage == C1_NSTAGE) { stage = 0; phase ^= 1; }
      }
      // Drain: the tcgen05.commit arrivals on the empty barriers of the last stages are asynchronous and nobody else
      // waits for them; the CTA must not exit (and hand its shared memory to the next kernel's CTA, which under
      // programmatic dependent launch is already queued) while one is in flight.
      for (int i = 0; i < C1_NSTAGE; ++i) {
        mbar_wait(&empty_bar[stage], phase ^ 1, 35);
        if (++stage == C1_NSTAGE) { stage = 0; phase ^= 1; }
      }
      wait_stage.store(0);
    }
  } else if (warp >= 1 && warp <= C1_MMA_WARPS) {
    // whole warp runs the uniform control flow; one elected lane issues UMMAs + commits
    constexpr uint32_t idesc = make_idesc_bf16(128, C1_N);
    // A: rows = pixel pairs 16 B apart, K-adjacent core matrix = next 2 pixels (LBO 16 B), next 8 rows =
    //    next output-row pair = two image rows down (SBO).  B: canonical no-swizzle, contiguous along K.
    constexpr uint32_t a_hi = desc_hi(2 * C1_ROWB, SW_NONE);
    constexpr uint32_t b_hi = desc_hi(C1_B_SBO, SW_NONE);
    constexpr uint32_t c_hi = desc_hi(256, SW_NONE);
    const uint32_t a_lo0 = desc_lo(smem_u32(smem_a), 16);
    const uint32_t b_lo0 = desc_lo(smem_u32(smem_b), 128);
    const uint32_t ones_lo = desc_lo(smem_u32(smem_ones), 128);
    const uint32_t bias_lo = desc_lo(smem_u32(smem_biasop), 128);
    mbar_wait(wload_bar, 0, 31);
    // Several issuing warps (1 .. C1_MMA_WARPS) take the tiles round-robin: a wait on an mbarrier costs the issuing thread ~130 clocks
    // even when the phase completed long ago (measured), and the tensor pipe's queue is too shallow to hide two of
    // them per 17-instruction tile; with two issuers one warp's waits overlap the other's instruction stream.
    // Tiles use disjoint stages / accumulators and every barrier is per tile, so no ordering between the two is
    // needed (tcgen05.commit tracks the MMAs of the committing thread).
    const int which = warp - 1;
    RoleTimer wait_acc, wait_ops, loop;
    loop.begin();
    for (int lt = which; blockIdx.x + (long long)lt * gridDim.x < total_tiles; lt += C1_MMA_WARPS) {
      const int stage = lt % C1_NSTAGE, acc = lt % C1_NACC;
      const uint32_t phase = (uint32_t)(lt / C1_NSTAGE) & 1u, acc_phase = (uint32_t)(lt / C1_NACC) & 1u;
      issue_gate(issued, lt, C1_NSTAGE, C1_MMA_WARPS, 36);
      if (C1_NACC != C1_NSTAGE) issue_gate(issued, lt, C1_NACC, C1_MMA_WARPS, 36);
      wait_acc.begin();
      mbar_wait(&tempty_bar[acc], acc_phase ^ 1, 32);
      wait_acc.end();
      if (lane == 0) trace(lt, 2);
      wait_ops.begin();
      mbar_wait(&full_bar[stage], phase, 33);
      wait_ops.end();
      if (lane == 0) trace(lt, 3);
      tc_fence_after_sync();
      if (elect_one()) {
        const uint32_t d_tmem = tmem_base + acc * C1_N;
        const uint32_t a_lo = a_lo0 + stage * (C1_STAGE_STRIDE >> 4);
        if (C1_BIAS_UMMA) umma_bf16_ss_w(d_tmem, ones_lo, c_hi, bias_lo, c_hi, idesc, 0u);     // D = bias
#ifndef SIA_C1_ROWS
#define SIA_C1_ROWS 8
#endif
#pragma unroll
        for (int r = 0; r < SIA_C1_ROWS; ++r) {       // (timing experiments may issue fewer window rows)
#pragma unroll
          for (int kk = 0; kk < 2; ++kk) {
            umma_bf16_ss_w(d_tmem, a_lo + r * (C1_ROWB >> 4) + kk * 2, a_hi, b_lo0 + (r * 4 + kk * 2) * 8, b_hi, idesc,
                           (C1_BIAS_UMMA || (r | kk)) ? 1u : 0u);
          }
        }
        umma_commit(&empty_bar[stage]);
        umma_commit(&tfull_bar[acc]);
      }
      __syncwarp();
      if (lane == 0) issue_done(issued, lt, C1_MMA_WARPS);
      if (lane == 0) trace(lt, 4);
    }
    loop.end();
    if (lane == 0 && warp == 1) { wait_acc.store(1); wait_ops.store(2); loop.store(3); }
  } else if (warp >= 4) {
    // epilogue: thread = one GEMM row = one pooled output pixel, 32 channels.  Two groups of four warps;
    // group g owns the tiles with local index j = g, g+2, ...
    const int group = (warp - 4) >> 2;
    const int e = (warp - 4) & 3;
    const int Ho = H >> 1, Wo = W >> 1;
    const int yp = 4 * e + (lane >> 3);
    const int xp = lane & 7;
    TileWalker1 t(blockIdx.x + group * gridDim.x, C1_EPI_GROUPS * gridDim.x, tiles_x, tiles_y);
    RoleTimer wait_full, eloop;
    unsigned long long ntiles = 0;
    eloop.begin();
    int j = group;
    for (int tile = blockIdx.x + group * gridDim.x; tile < total_tiles;
         tile += C1_EPI_GROUPS * gridDim.x, t.next(), j += C1_EPI_GROUPS) {
      ++ntiles;
      const int acc = j % C1_NACC;
      const uint32_t acc_phase = (j / C1_NACC) & 1;
      const int py = t.ty * (C1_TILE_Y / 2) + yp;
      const int px = t.tx * (C1_TILE_X / 2) + xp;
      const bool in_range = py < Ho && px < Wo;
      uint4* opix = reinterpret_cast<uint4*>(out + (((size_t)t.n * Ho + py) * Wo + px) * c_stride);
      wait_full.begin();
      mbar_wait(&tfull_bar[acc], acc_phase, 34);
      wait_full.end();
      if (e == 0 && lane == 0) trace(j, 5);
      tc_fence_after_sync();
      const uint32_t t_addr = tmem_base + ((uint32_t)(32 * e) << 16) + acc * C1_N;

      // max over the 2x2 window (4 column groups; the bias is already in the accumulator), ReLU, bf16;
      // 16 channels -> 2 x 16-byte stores
      auto finish_half = [&](const uint32_t (&q0)[16], const uint32_t (&q1)[16], const uint32_t (&q2)[16],
                             const uint32_t (&q3)[16], int half) {
        uint32_t pk[8];
#pragma unroll
        for (int c2 = 0; c2 < 8; ++c2) {
          const int c = 2 * c2;
          const float a = fmaxf(fmaxf(__uint_as_float(q0[c]), __uint_as_float(q1[c])),
                                fmaxf(__uint_as_float(q2[c]), __uint_as_float(q3[c])));
          const float b = fmaxf(fmaxf(__uint_as_float(q0[c + 1]), __uint_as_float(q1[c + 1])),
                                fmaxf(__uint_as_float(q2[c + 1]), __uint_as_float(q3[c + 1])));
          if (C1_BIAS_UMMA) {
            pk[c2] = max_bf16x2(pack_bf16x2(a, b), 0u);
          } else {
            pk[c2] = pack_bf16x2(fmaxf(a + smem_bias[16 * half + c], 0.f), fmaxf(b + smem_bias[16 * half + c + 1], 0.f));
          }
        }
        if (in_range) {
          opix[2 * half] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          opix[2 * half + 1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
        }
      };

#ifdef SIA_C1_NO_EPI
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);     // timing experiment: the accumulator is not read
      continue;
#endif
      uint32_t a0[16], a1[16], a2[16], a3[16], b0[16], b1[16], b2[16], b3[16];
      tmem_ld16(t_addr + 0, a0);
      tmem_ld16(t_addr + 32, a1);
      tmem_ld16(t_addr + 64, a2);
      tmem_ld16(t_addr + 96, a3);
      tmem_ld_wait();
      tmem_ld16(t_addr + 16, b0);
      tmem_ld16(t_addr + 48, b1);
      tmem_ld16(t_addr + 80, b2);
      tmem_ld16(t_addr + 112, b3);
      finish_half(a0, a1, a2, a3, 0);
      tmem_ld_wait();
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);   // the whole accumulator is in registers now
      if (e == 0 && lane == 0) trace(j, 6);
      finish_half(b0, b1, b2, b3, 1);
      if (e == 0 && lane == 0) trace(j, 7);
    }
    eloop.end();
    if (warp == 4 && lane == 0) {
      wait_full.store(4);
      eloop.store(5);
      stats_store(6, ntiles);
    }
  }

#ifdef SIA_C1_RAW_ISSUE
c1_done:
#endif
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 2) tmem_free(tmem_base, C1_NACC * C1_N);
}

// [32][3][7][7] fp32 -> B[n = (dy*2+dx)*32 + co][k = r'*32 + xw*4 + c] = W[co][c][r'-dy][xw-dx], bf16, in
// no-swizzle core-matrix order: byte offset = (n/8)*4096 + (k/8)*128 + (n%8)*16 + (k%8)*2.
__global__ void pack_conv1_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ dst) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < C1_N * C1_K; i += gridDim.x * blockDim.x) {
    const int n = i / C1_K, k = i % C1_K;
    const int q = n / 32, co = n % 32;
    const int dy = q >> 1, dx = q & 1;
    const int rr = k / 32, xw = (k % 32) / 4, c = k % 4;
    const int r = rr - dy, t = xw - dx;
    float v = 0.f;
    if (c < 3 && r >= 0 && r < 7 && t >= 0 && t < 7) v = w[((co * 3 + c) * 7 + r) * 7 + t];
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

extern "C" int sia_conv7x7_c3_relu_pool2_strided(const void* in_nhwc4, int batch, int h, int w, const void* w_packed,
                                                 const float* bias, void* out_nhwc, int c_stride, int c_offset,
                                                 void* stream) {
  using namespace sia;
  SIA_REQUIRE(in_nhwc4 && w_packed && bias && out_nhwc && batch >= 1 && h >= 2 && w >= 16);
  SIA_REQUIRE(aligned(in_nhwc4, 16) && aligned(w_packed, 16) && aligned(out_nhwc, 16));
  SIA_REQUIRE(c_stride >= 32 && c_stride % 8 == 0 && c_offset >= 0 && c_offset % 8 == 0 && c_offset + 32 <= c_stride);
  if (h % 2 != 0 || w % C1_TILE_X != 0) return SIA_E_UNSUPPORTED;
  if (int wrc = ensure_watchdog()) return wrc;
  CUtensorMap tmap;
  // padded NHWC4 rows: (w + 8) pixels of 8 bytes, image pixel x in column x + 1
  const uint64_t wp = (uint64_t)w + SIA_NHWC4_PAD;
  const uint64_t dims[3] = {wp * 4, (uint64_t)h, (uint64_t)batch};
  const uint64_t strides[2] = {wp * 8, (uint64_t)h * wp * 8};
  const uint32_t box[3] = {C1_WIN_PX * 4, C1_ROWS, 1};
  int rc = encode_tmap_bf16(&tmap, in_nhwc4, 3, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE);
  if (rc != 0) return rc;
  const int tiles_y = (h + C1_TILE_Y - 1) / C1_TILE_Y, tiles_x = w / C1_TILE_X;
  const int total = tiles_y * tiles_x * batch;
  const int smem = 1024 + C1_B_BYTES + C1_NSTAGE * C1_STAGE_STRIDE + ONES_BYTES + C1_BIAS_BYTES +
                   (2 * C1_NSTAGE + 2 * C1_NACC + 4) * 8;
  static SmemSlots configured = {};
  if (int rc2 = ensure_dynamic_smem(conv1_kernel, smem, &configured)) return rc2;
  const int grid = total < sm_count() ? total : sm_count();
  return launch_kernel(conv1_kernel, dim3(grid), dim3(C1_THREADS), smem, static_cast<cudaStream_t>(stream), true,
      tmap, static_cast<const uint8_t*>(w_packed), bias, static_cast<__nv_bfloat16*>(out_nhwc) + c_offset, h, w, tiles_y,
      tiles_x, total, c_stride);
  return launch_status();
}

extern "C" int sia_conv7x7_c3_relu_pool2(const void* in_nhwc4, int batch, int h, int w, const void* w_packed,
                                         const float* bias, void* out_nhwc, void* stream) {
  return sia_conv7x7_c3_relu_pool2_strided(in_nhwc4, batch, h, w, w_packed, bias, out_nhwc, 32, 0, stream);
}
