// K4 (3x3 blocks): Conv2d(cin, cout, 3, stride 1, 'same') + bias + ReLU + MaxPool2d(2,2) as an
// implicit GEMM on the 5th-gen tensor cores.  Replaces the ATen conv / relu / max_pool2d calls of
// tone_bias_model.py:83-92 (layers.3, layers.6) and :174-184 (conv2..conv4).
//
//   D[128 pixels, cout] = sum over 9 taps (r,s), cin:  X[pixel + (r-1, s-1), cin] * W[cout, cin, r, s]
//
// * M-tile  = 16 rows x 8 columns of output pixels (row m = y*8 + x); N = cout; K = 9*cin.
// * A operand: ONE halo copy of the input patch per tile -- TMA box [<=64 ch, 10 x, 18 y] (zero fill
//   outside the image = 'same' padding) lands as 180 shared-memory rows of one pixel each in the
//   swizzled K-major layout.  Tap (r,s) is the same buffer read from row r*10+s with an 8-row-group
//   stride of one halo row (SBO = 10 rows): the swizzle XOR is a function of the absolute smem address
//   (pinned on hardware by tests/test_umma_probe.py), so shifted starts and non-1024-byte group
//   strides address exactly what TMA wrote.  Every input byte is fetched once per tile and feeds all
//   9 taps; a ring of such buffers decouples TMA from the MMA issue.
// * B operand: all 9 taps of the packed weights stay resident in shared memory (loaded once per
//   persistent CTA with bulk copies).
// * Accumulators: two 128 x cout fp32 tiles in TMEM, so the epilogue of tile i overlaps the MMAs
//   of tile i+1.
// * Epilogue (4 warps): tcgen05.ld -> +bias -> ReLU -> bf16 -> 2x2 max-pool across lanes with a
//   shuffle reduce-scatter (x-neighbour = lane^1, y-neighbour = lane^8) -> 16-byte NHWC stores.
//   The un-pooled activation never leaves the SM.
//
// Tensor-core bound; algorithmic FLOPs = 2 * H*W * 9*cin * cout per image.
#include <stdlib.h>

#include "sia_host.cuh"
#include "sia_ptx.cuh"

namespace sia {

constexpr int CV_TILE_Y = 16;
constexpr int CV_TILE_X = 8;
constexpr int CV_HALO_Y = CV_TILE_Y + 2;
constexpr int CV_HALO_X = CV_TILE_X + 2;
constexpr int CV_HALO_ROWS = CV_HALO_Y * CV_HALO_X;   // 180 pixels
constexpr int CV_NACC = 4;       // TMEM accumulator ring
#ifndef SIA_CV_MMA_WARPS
#define SIA_CV_MMA_WARPS 2
#endif
constexpr int CV_MMA_WARPS = SIA_CV_MMA_WARPS;   // warps issuing UMMAs (1 .. 3), tiles dealt round-robin
#ifndef SIA_CP_MMA_WARPS
#define SIA_CP_MMA_WARPS 3
#endif
constexpr int CP_MMA_WARPS = SIA_CP_MMA_WARPS;   // same for the pixel-pair kernel (measured: 3 beats 2 there, 2 beats 3 above)
constexpr int CV_THREADS = 384;  // warp0 TMA, warp1 MMA, warp2 TMEM alloc, warp3 idle,
                                 // warps 4-7 epilogue group 0 (even tiles), warps 8-11 group 1 (odd tiles)
constexpr int ONES_BYTES = 4096;                      // [128 rows][16 k] bf16, no-swizzle core matrices

template <int CIN, int COUT>
struct ConvCfg {
  static constexpr int CK = CIN < 64 ? CIN : 64;       // channels per smem row
  static constexpr int NCHUNK = CIN / CK;
  static constexpr int ROWB = CK * 2;                   // bytes per smem row (one pixel)
  static constexpr uint32_t SWZ = ROWB == 128 ? SW_128B : SW_64B;
  static constexpr int GROUP_STRIDE = CV_HALO_X * ROWB;                               // SBO: next tile row
  static constexpr int CHUNK_BYTES = CV_HALO_ROWS * ROWB;                              // one TMA box
  static constexpr int CHUNK_STRIDE = (CHUNK_BYTES + 1023) / 1024 * 1024;
  static constexpr int STAGE_STRIDE = NCHUNK * CHUNK_STRIDE;
  static constexpr int STAGE_TX_BYTES = NCHUNK * CHUNK_BYTES;
  static constexpr int B_TAP_BYTES = COUT * ROWB;       // one (tap, chunk) weight block
  static constexpr int B_BYTES = 9 * NCHUNK * B_TAP_BYTES;
  static constexpr int BIAS_BYTES = COUT * 32;          // [COUT rows][16 k] bf16: k0 = hi(bias), k1 = lo(bias)
  static constexpr int TMEM_COLS = CV_NACC * COUT;     // resident-weight kernel: ring of CV_NACC accumulators
};

// Decomposes the persistent-CTA tile stride once, so that walking tiles needs no integer division.
struct TileWalker {
  int tx, ty, n, dtx, dty, dn, tiles_x, tiles_y;
  __device__ TileWalker(int first, int step, int tiles_x_, int tiles_y_) : tiles_x(tiles_x_), tiles_y(tiles_y_) {
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

// The bias enters through the tensor core: the first UMMA of every tile multiplies a constant "ones"
// operand (columns 0,1 = 1) with [hi(bias), lo(bias)] (bias split into two bf16 so that hi+lo carries 16
// mantissa bits) and INITIALISES the accumulator with it.  The epilogue is left with ReLU + pool + pack.
__device__ __forceinline__ void fill_ones_operand(uint8_t* ones, int tid, int nthreads) {
  // element (r,k) at (r/8)*256 + (k/8)*128 + (r%8)*16 + (k%8)*2 ; k in {0,1} -> 1.0
  for (int i = tid; i < ONES_BYTES / 4; i += nthreads) {
    const int byte = i * 4;
    const int k = ((byte >> 7) & 1) * 8 + ((byte & 15) >> 1);
    reinterpret_cast<uint32_t*>(ones)[i] = (k == 0) ? 0x3F803F80u : 0u;   // bf16 1.0 in k = 0 and k = 1
  }
}
__device__ __forceinline__ void fill_bias_operand(uint8_t* dst, const float* bias, int n_rows, int bias_mod, int tid,
                                                  int nthreads) {
  for (int r = tid; r < n_rows; r += nthreads) {
    const float b = bias[r % bias_mod];
    const __nv_bfloat16 hi = __float2bfloat16_rn(b);
    const __nv_bfloat16 lo = __float2bfloat16_rn(b - __bfloat162float(hi));
    uint32_t* row = reinterpret_cast<uint32_t*>(dst + (r / 8) * 256 + (r % 8) * 16);
    row[0] = (uint32_t)__bfloat16_as_ushort(hi) | ((uint32_t)__bfloat16_as_ushort(lo) << 16);
    row[1] = row[2] = row[3] = 0u;
    uint32_t* row_k8 = reinterpret_cast<uint32_t*>(dst + (r / 8) * 256 + 128 + (r % 8) * 16);
    row_k8[0] = row_k8[1] = row_k8[2] = row_k8[3] = 0u;
  }
}

template <int CIN, int COUT, int NSTAGE>
__global__ void __launch_bounds__(CV_THREADS, 1)
conv3x3_kernel(const __grid_constant__ CUtensorMap tmap_in, const uint8_t* __restrict__ w_packed,
               const float* __restrict__ bias, __nv_bfloat16* __restrict__ out, int H, int W, int tiles_y,
               int tiles_x, int total_tiles) {
  using C = ConvCfg<CIN, COUT>;
  static_assert(C::TMEM_COLS <= 512, "accumulator ring does not fit TMEM");
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // keep the shared address space visible to the compiler (offset arithmetic, no integer round trip)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* smem_b = smem;                               // B_BYTES (multiple of 1024)
  uint8_t* smem_a = smem + C::B_BYTES;                  // NSTAGE * STAGE_STRIDE
  uint8_t* smem_ones = smem_a + NSTAGE * C::STAGE_STRIDE;
  uint8_t* smem_biasop = smem_ones + ONES_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_biasop + C::BIAS_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + NSTAGE;
  uint64_t* tfull_bar = bars + 2 * NSTAGE;
  uint64_t* tempty_bar = bars + 2 * NSTAGE + CV_NACC;
  uint64_t* wload_bar = bars + 2 * NSTAGE + 2 * CV_NACC;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NSTAGE + 2 * CV_NACC + 1);
  volatile uint32_t* issued = tmem_slot + 1;            // [CV_MMA_WARPS] tiles issued per issuing warp (issue_gate)

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < NSTAGE; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < CV_NACC; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 4);
    }
    mbar_init(wload_bar, 1);
    for (int i = 0; i < CV_MMA_WARPS; ++i) issued[i] = 0;
    fence_mbar_init();
    tma_prefetch_desc(&tmap_in);
  }
  if (warp == 2) tmem_alloc(tmem_slot, C::TMEM_COLS);
  fill_ones_operand(smem_ones, threadIdx.x, blockDim.x);
  fill_bias_operand(smem_biasop, bias, COUT, COUT, threadIdx.x, blockDim.x);
  fence_proxy_async_smem();      // generic-proxy writes above -> visible to the UMMA operand reads
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = bcast0(*tmem_slot);
  // programmatic dependent launch (see sia_ptx.cuh): the prologue above and the resident weights overlap the previous
  // kernel's tail; the producer lane waits after it has issued the weight copies
  pdl_launch_dependents();
  if (!(warp == 0 && lane == 0)) pdl_wait();

  if (warp == 0) {
    // ================================ TMA producer ==========================================
    if (lane == 0) {
      // resident weights: one expect_tx, a few bulk copies
      mbar_arrive_expect_tx(wload_bar, C::B_BYTES);
      constexpr int PIECE = 16384;
      for (int off = 0; off < C::B_BYTES; off += PIECE) {
        const int n = C::B_BYTES - off < PIECE ? C::B_BYTES - off : PIECE;
        bulk_load_1d(smem_b + off, w_packed + off, n, wload_bar);
      }
      pdl_wait();                       // the input tiles below are the previous kernel's output
      int stage = 0;
      uint32_t phase = 0;
      TileWalker t(blockIdx.x, gridDim.x, tiles_x, tiles_y);
      RoleTimer wait_stage;
      int lt = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, t.next(), ++lt) {
        wait_stage.begin();
        mbar_wait(&empty_bar[stage], phase ^ 1, 20);
        wait_stage.end();
        trace(lt, 0);
        mbar_arrive_expect_tx(&full_bar[stage], C::STAGE_TX_BYTES);
#pragma unroll
        for (int kc = 0; kc < C::NCHUNK; ++kc) {
          tma_load_4d(smem_a + stage * C::STAGE_STRIDE + kc * C::CHUNK_STRIDE, &tmap_in, &full_bar[stage],
                      kc * C::CK, t.tx * CV_TILE_X - 1, t.ty * CV_TILE_Y - 1, t.n);
        }
        trace(lt, 1);
        if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
      }
      // Drain: the tcgen05.commit arrivals on the empty barriers of the last stages are asynchronous and nobody else
      // waits for them; the CTA must not exit (and hand its shared memory to the next kernel's CTA, which under
      // programmatic dependent launch is already queued) while one is in flight.
      for (int i = 0; i < NSTAGE; ++i) {
        mbar_wait(&empty_bar[stage], phase ^ 1, 60);
        if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
      }
      wait_stage.store(0);
    }
  } else if (warp >= 1 && warp <= CV_MMA_WARPS) {
    // ================================ MMA issuer ============================================
    // The whole warp runs the (warp-uniform) control flow so descriptors stay in uniform registers;
    // one elected lane issues the UMMAs and their commits.
    constexpr uint32_t idesc = make_idesc_bf16(128, COUT);
    constexpr uint32_t a_hi = desc_hi(C::GROUP_STRIDE, C::SWZ);   // 8-row groups one halo row apart
    constexpr uint32_t b_hi = desc_hi(8 * C::ROWB, C::SWZ);       // dense rows
    constexpr uint32_t c_hi = desc_hi(256, SW_NONE);              // ones / bias operands
    const uint32_t a_lo0 = desc_lo(smem_u32(smem_a), 0);
    const uint32_t b_lo0 = desc_lo(smem_u32(smem_b), 0);
    const uint32_t ones_lo = desc_lo(smem_u32(smem_ones), 128);
    const uint32_t bias_lo = desc_lo(smem_u32(smem_biasop), 128);
    mbar_wait(wload_bar, 0, 21);
    // CV_MMA_WARPS issuing warps (1 ..) take the tiles round-robin: an mbarrier wait costs the issuing thread ~130
    // clocks even when the phase completed long ago, and the tensor pipe's queue does not hide two of them per
    // tile; one warp's waits now overlap another's instruction stream (see conv1.cu).  Tiles use disjoint stages /
    // accumulators and every barrier is per tile, so the issuers need no ordering among themselves.
    RoleTimer wait_acc, wait_ops, loop;
    loop.begin();
    for (int lt = warp - 1; blockIdx.x + (long long)lt * gridDim.x < total_tiles; lt += CV_MMA_WARPS) {
      const int stage = lt % NSTAGE, acc = lt % CV_NACC;
      const uint32_t phase = (uint32_t)(lt / NSTAGE) & 1u, acc_phase = (uint32_t)(lt / CV_NACC) & 1u;
      issue_gate(issued, lt, NSTAGE, CV_MMA_WARPS, 61);       // sia_ptx.cuh: parity waits need the previous use issued
      if (CV_NACC != NSTAGE) issue_gate(issued, lt, CV_NACC, CV_MMA_WARPS, 61);
      wait_acc.begin();
      mbar_wait(&tempty_bar[acc], acc_phase ^ 1, 22);
      wait_acc.end();
      if (lane == 0) trace(lt, 2);
      wait_ops.begin();
      mbar_wait(&full_bar[stage], phase, 23);
      wait_ops.end();
      if (lane == 0) trace(lt, 3);
      tc_fence_after_sync();
      if (elect_one()) {
        const uint32_t d_tmem = tmem_base + acc * COUT;
        const uint32_t a_stage = a_lo0 + stage * (C::STAGE_STRIDE >> 4);
        umma_bf16_ss_w(d_tmem, ones_lo, c_hi, bias_lo, c_hi, idesc, 0u);     // D = bias
#pragma unroll
        for (int kc = 0; kc < C::NCHUNK; ++kc) {
#pragma unroll
          for (int r = 0; r < 3; ++r) {
#pragma unroll
            for (int s = 0; s < 3; ++s) {
              const uint32_t a_lo = a_stage + ((kc * C::CHUNK_STRIDE + (r * CV_HALO_X + s) * C::ROWB) >> 4);
              const uint32_t b_lo = b_lo0 + (((r * 3 + s) * C::NCHUNK + kc) * C::B_TAP_BYTES >> 4);
#pragma unroll
              for (int kk = 0; kk < C::CK / 16; ++kk) {
                umma_bf16_ss_w(d_tmem, a_lo + kk * 2, a_hi, b_lo + kk * 2, b_hi, idesc, 1u);
              }
            }
          }
        }
        umma_commit(&empty_bar[stage]);   // halo buffer reusable once these MMAs have read it
        umma_commit(&tfull_bar[acc]);     // accumulator complete -> epilogue
      }
      __syncwarp();
      if (lane == 0) issue_done(issued, lt, CV_MMA_WARPS);
      if (lane == 0) trace(lt, 4);
    }
    loop.end();
    if (lane == 0 && warp == 1) { wait_acc.store(1); wait_ops.store(2); loop.store(3); }
  } else if (warp >= 4) {
    // ================================ epilogue ==============================================
    // two groups of four warps; group g owns the tiles with local index j = g, g+2, g+4, ...
    const int group = (warp - 4) >> 2;
    const int e = (warp - 4) & 3;             // TMEM lanes 32e .. 32e+31
    const int Ho = H >> 1, Wo = W >> 1;
    const int ly = lane >> 3;                 // 0..3  (tile row 4e + ly)
    const int lx = lane & 7;                  // tile column
    const bool odd_x = lane & 1;
    const bool odd_y = (lane >> 3) & 1;
    TileWalker t(blockIdx.x + group * gridDim.x, 2 * gridDim.x, tiles_x, tiles_y);
    RoleTimer wait_full, eloop;
    unsigned long long ntiles = 0;
    eloop.begin();
    int j = group;
    for (int tile = blockIdx.x + group * gridDim.x; tile < total_tiles; tile += 2 * gridDim.x, t.next(), j += 2) {
      ++ntiles;
      const int acc = j % CV_NACC;
      const uint32_t acc_phase = (j / CV_NACC) & 1;
      const int py = ((t.ty * CV_TILE_Y + 4 * e + ly) >> 1);
      const int px = ((t.tx * CV_TILE_X + lx) >> 1);
      const bool in_range = py < Ho && px < Wo;
      __nv_bfloat16* orow = out + (((size_t)t.n * Ho + py) * Wo + px) * COUT;
      wait_full.begin();
      mbar_wait(&tfull_bar[acc], acc_phase, 24);
      wait_full.end();
      if (e == 0 && lane == 0) trace(j, 5);
      tc_fence_after_sync();
      const uint32_t t_addr = tmem_base + ((uint32_t)(32 * e) << 16) + acc * COUT;

      // bf16 + 2x2 max-pool (shuffle reduce-scatter) + ReLU + 16-byte store of one 32-channel chunk
      // (rounding, max and ReLU commute, so the order is free; the bias is already in the accumulator)
      auto finish_chunk = [&](const uint32_t (&v)[32], int cb) {
        uint32_t pk[16];
#pragma unroll
        for (int q = 0; q < 16; ++q) pk[q] = pack_bf16x2(__uint_as_float(v[2 * q]), __uint_as_float(v[2 * q + 1]));
        uint32_t h8[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {     // x partner (lane^1): keep 8 of 16 registers, send the other 8
          const uint32_t keep = odd_x ? pk[8 + q] : pk[q];
          const uint32_t send = odd_x ? pk[q] : pk[8 + q];
          h8[q] = max_bf16x2(keep, __shfl_xor_sync(0xffffffffu, send, 1));
        }
        uint32_t q4[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {     // y partner (lane^8): keep 4 of 8, then ReLU
          const uint32_t keep = odd_y ? h8[4 + q] : h8[q];
          const uint32_t send = odd_y ? h8[q] : h8[4 + q];
          q4[q] = max_bf16x2(max_bf16x2(keep, __shfl_xor_sync(0xffffffffu, send, 8)), 0u);
        }
        if (in_range) {
          const int ch = cb + (odd_x ? 16 : 0) + (odd_y ? 8 : 0);
          *reinterpret_cast<uint4*>(orow + ch) = make_uint4(q4[0], q4[1], q4[2], q4[3]);
        }
      };

      // two register buffers: the TMEM load of chunk c+1 is in flight while chunk c is finished
      uint32_t va[32], vb[32];
      tmem_ld32(t_addr, va);
#pragma unroll
      for (int cb = 0; cb < COUT; cb += 64) {
        tmem_ld_wait();
        tmem_ld32(t_addr + cb + 32, vb);
        finish_chunk(va, cb);
        tmem_ld_wait();
        if (cb + 64 < COUT) {
          tmem_ld32(t_addr + cb + 64, va);
        } else {
          // every column of this accumulator is in registers: hand it back to the MMA warp
          tc_fence_before_sync();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tempty_bar[acc]);
          if (e == 0 && lane == 0) trace(j, 6);
        }
        finish_chunk(vb, cb + 32);
      }
      if (e == 0 && lane == 0) trace(j, 7);
    }
    eloop.end();
    if (warp == 4 && lane == 0) {
      wait_full.store(4);
      eloop.store(5);
      stats_store(6, ntiles);
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 2) tmem_free(tmem_base, C::TMEM_COLS);
}


// ------------------------------- streamed-weight variant --------------------------------------
// Any 64-channel-chunked input (cin_pad = 64 * nchunk, nchunk <= 4) and COUT in {64, 128, 192, 256}: the weights
// (up to 9 * 256 * 256 bf16 = 1.1 MB) do not have to fit in shared memory, they stream from L2 through a ring of
// (tap, 64-channel chunk) blocks of COUT x 128 bytes while TILES output tiles stay resident (halo buffers in
// shared memory, 128 x COUT fp32 accumulators in TMEM).  With TILES = 2 every weight block is used by both
// tiles, which halves the L2 traffic per FLOP.  Epilogue group t (four warps) drains accumulator t; the MMA warp
// starts the next pass once all are drained.  Used for Conv2d(128, 256) of SkinCancerModel and, with zero-padded
// channels, for the arbitrary widths of tone_bias_optuna.define_isic_model.
constexpr int CS_NB = 3;            // weight-block ring
constexpr int CS_CHUNK_BYTES = CV_HALO_ROWS * 128;                         // one 64-channel halo chunk (23040)
constexpr int CS_CHUNK_STRIDE = (CS_CHUNK_BYTES + 1023) / 1024 * 1024;

template <int COUT, int TILES>
__global__ void __launch_bounds__(CV_THREADS, 1)
conv3x3_stream_kernel(const __grid_constant__ CUtensorMap tmap_in, const uint8_t* __restrict__ w_packed,
                      const float* __restrict__ bias, __nv_bfloat16* __restrict__ out, int H, int W, int tiles_y,
                      int tiles_x, int total_tiles, int nchunk) {
  static_assert(TILES * COUT <= 512, "the accumulators must fit TMEM");
  static_assert(COUT % 64 == 0 && COUT <= 256 && (TILES == 1 || TILES == 2), "unsupported shape");
  constexpr int B_BLOCK = COUT * 128;                  // one (tap, chunk) weight block
  constexpr int BIAS_BYTES = COUT * 32;
  constexpr uint32_t TMEM_COLS = TILES * COUT <= 64 ? 64 : TILES * COUT <= 128 ? 128 : TILES * COUT <= 256 ? 256 : 512;
  const int nblk = 9 * nchunk;                         // weight blocks per pass
  const int stage_stride = nchunk * CS_CHUNK_STRIDE;   // one halo buffer (all chunks)
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* smem_b = smem;                               // CS_NB * B_BLOCK
  uint8_t* smem_a = smem + CS_NB * B_BLOCK;             // TILES * stage_stride
  uint8_t* smem_ones = smem_a + TILES * stage_stride;
  uint8_t* smem_biasop = smem_ones + ONES_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_biasop + BIAS_BYTES);
  uint64_t* b_full = bars;                 // [CS_NB]
  uint64_t* b_empty = bars + CS_NB;        // [CS_NB]
  uint64_t* a_full = bars + 2 * CS_NB;     // the halos of the pass landed
  uint64_t* a_empty = a_full + 1;          // the MMAs of the pass have read them
  uint64_t* tfull_bar = a_full + 2;        // accumulators complete
  uint64_t* tempty_bar = a_full + 3;       // [2] drained by epilogue group t
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a_full + 5);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_pass = (total_tiles + TILES - 1) / TILES;

  if (threadIdx.x == 0) {
    for (int i = 0; i < CS_NB; ++i) {
      mbar_init(&b_full[i], 1);
      mbar_init(&b_empty[i], 1);
    }
    mbar_init(a_full, 1);
    mbar_init(a_empty, 1);
    mbar_init(tfull_bar, 1);
    mbar_init(&tempty_bar[0], 4);
    mbar_init(&tempty_bar[1], 4);
    fence_mbar_init();
    tma_prefetch_desc(&tmap_in);
  }
  if (warp == 2) tmem_alloc(tmem_slot, TMEM_COLS);
  fill_ones_operand(smem_ones, threadIdx.x, blockDim.x);
  fill_bias_operand(smem_biasop, bias, COUT, COUT, threadIdx.x, blockDim.x);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = bcast0(*tmem_slot);
  pdl_launch_dependents();          // programmatic dependent launch: the prologue above overlaps the previous kernel
  pdl_wait();
  const int tiles_per_img = tiles_x * tiles_y;

  if (warp == 0) {
    // ================================ producer ==============================================
    if (lane == 0) {
      int bs = 0;
      uint32_t bphase = 0, pphase = 0;
      for (int pass = blockIdx.x; pass < n_pass; pass += gridDim.x) {
        mbar_wait(a_empty, pphase ^ 1, 25);
        const int n_tiles = (TILES * pass + TILES <= total_tiles) ? TILES : total_tiles - TILES * pass;
        mbar_arrive_expect_tx(a_full, n_tiles * nchunk * CS_CHUNK_BYTES);
        for (int t = 0; t < n_tiles; ++t) {
          const int tile = TILES * pass + t;
          const int n = tile / tiles_per_img, rem = tile % tiles_per_img;
          const int ty = rem / tiles_x, tx = rem % tiles_x;
          for (int kc = 0; kc < nchunk; ++kc) {
            tma_load_4d(smem_a + t * stage_stride + kc * CS_CHUNK_STRIDE, &tmap_in, a_full, kc * 64,
                        tx * CV_TILE_X - 1, ty * CV_TILE_Y - 1, n);
          }
        }
        pphase ^= 1;
        for (int blk = 0; blk < nblk; ++blk) {
          mbar_wait(&b_empty[bs], bphase ^ 1, 26);
          mbar_arrive_expect_tx(&b_full[bs], B_BLOCK);
          constexpr int PIECE = 8192;
#pragma unroll
          for (int off = 0; off < B_BLOCK; off += PIECE) {
            bulk_load_1d(smem_b + bs * B_BLOCK + off, w_packed + (size_t)blk * B_BLOCK + off, PIECE, &b_full[bs]);
          }
          if (++bs == CS_NB) { bs = 0; bphase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ============================================
    constexpr uint32_t idesc = make_idesc_bf16(128, COUT);
    constexpr uint32_t a_hi = desc_hi(CV_HALO_X * 128, SW_128B);
    constexpr uint32_t b_hi = desc_hi(8 * 128, SW_128B);
    constexpr uint32_t c_hi = desc_hi(256, SW_NONE);
    const uint32_t a_lo0 = desc_lo(smem_u32(smem_a), 0);
    const uint32_t b_lo0 = desc_lo(smem_u32(smem_b), 0);
    const uint32_t ones_lo = desc_lo(smem_u32(smem_ones), 128);
    const uint32_t bias_lo = desc_lo(smem_u32(smem_biasop), 128);
    int bs = 0;
    uint32_t bphase = 0, pphase = 0;
    for (int pass = blockIdx.x; pass < n_pass; pass += gridDim.x) {
      const int n_tiles = (TILES * pass + TILES <= total_tiles) ? TILES : total_tiles - TILES * pass;
      mbar_wait(&tempty_bar[0], pphase ^ 1, 27);
      mbar_wait(&tempty_bar[1], pphase ^ 1, 27);
      mbar_wait(a_full, pphase, 28);
      tc_fence_after_sync();
      if (elect_one()) {
        for (int t = 0; t < n_tiles; ++t)
          umma_bf16_ss_w(tmem_base + t * COUT, ones_lo, c_hi, bias_lo, c_hi, idesc, 0u);     // D = bias
      }
      __syncwarp();
      int tap = 0, kc = 0;                 // block order of the packed weights: (tap, chunk), chunk fastest
      for (int blk = 0; blk < nblk; ++blk) {
        const int r = tap / 3, s = tap % 3;
        mbar_wait(&b_full[bs], bphase, 29);
        tc_fence_after_sync();
        if (elect_one()) {
          const uint32_t b_lo = b_lo0 + bs * (B_BLOCK >> 4);
          for (int t = 0; t < n_tiles; ++t) {
            const uint32_t a_lo = a_lo0 + ((t * stage_stride + kc * CS_CHUNK_STRIDE + (r * CV_HALO_X + s) * 128) >> 4);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              umma_bf16_ss_w(tmem_base + t * COUT, a_lo + kk * 2, a_hi, b_lo + kk * 2, b_hi, idesc, 1u);
            }
          }
          umma_commit(&b_empty[bs]);
          if (blk == nblk - 1) {
            umma_commit(a_empty);
            umma_commit(tfull_bar);
          }
        }
        __syncwarp();
        if (++bs == CS_NB) { bs = 0; bphase ^= 1; }
        if (++kc == nchunk) { kc = 0; ++tap; }
      }
      pphase ^= 1;
    }
  } else if (warp >= 4) {
    // ================================ epilogue ==============================================
    const int group = (warp - 4) >> 2;        // tile of the pass
    const int e = (warp - 4) & 3;             // TMEM lanes 32e .. 32e+31
    const int Ho = H >> 1, Wo = W >> 1;
    const int ly = lane >> 3;
    const int lx = lane & 7;
    const bool odd_x = lane & 1;
    const bool odd_y = (lane >> 3) & 1;
    uint32_t pphase = 0;
    for (int pass = blockIdx.x; pass < n_pass; pass += gridDim.x) {
      const int tile = TILES * pass + group;
      const bool have = group < TILES && tile < total_tiles;
      const int n = tile / tiles_per_img, rem = tile % tiles_per_img;
      const int ty = rem / tiles_x, tx = rem % tiles_x;
      const int py = ((ty * CV_TILE_Y + 4 * e + ly) >> 1);
      const int px = ((tx * CV_TILE_X + lx) >> 1);
      const bool in_range = have && py < Ho && px < Wo;
      __nv_bfloat16* orow = out + (((size_t)n * Ho + py) * Wo + px) * COUT;
      mbar_wait(tfull_bar, pphase, 24);
      pphase ^= 1;
      tc_fence_after_sync();
      const uint32_t t_addr = tmem_base + ((uint32_t)(32 * e) << 16) + (group < TILES ? group : 0) * COUT;

      auto finish_chunk = [&](const uint32_t (&v)[32], int cb) {
        uint32_t pk[16];
#pragma unroll
        for (int q = 0; q < 16; ++q) pk[q] = pack_bf16x2(__uint_as_float(v[2 * q]), __uint_as_float(v[2 * q + 1]));
        uint32_t h8[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const uint32_t keep = odd_x ? pk[8 + q] : pk[q];
          const uint32_t send = odd_x ? pk[q] : pk[8 + q];
          h8[q] = max_bf16x2(keep, __shfl_xor_sync(0xffffffffu, send, 1));
        }
        uint32_t q4[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const uint32_t keep = odd_y ? h8[4 + q] : h8[q];
          const uint32_t send = odd_y ? h8[q] : h8[4 + q];
          q4[q] = max_bf16x2(max_bf16x2(keep, __shfl_xor_sync(0xffffffffu, send, 8)), 0u);
        }
        if (in_range) {
          const int ch = cb + (odd_x ? 16 : 0) + (odd_y ? 8 : 0);
          *reinterpret_cast<uint4*>(orow + ch) = make_uint4(q4[0], q4[1], q4[2], q4[3]);
        }
      };

      if (have) {
        uint32_t va[32], vb[32];
        tmem_ld32(t_addr, va);
#pragma unroll
        for (int cb = 0; cb < COUT; cb += 64) {
          tmem_ld_wait();
          tmem_ld32(t_addr + cb + 32, vb);
          finish_chunk(va, cb);
          tmem_ld_wait();
          if (cb + 64 < COUT) tmem_ld32(t_addr + cb + 64, va);
          finish_chunk(vb, cb + 32);
        }
      }
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[group]);
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 2) tmem_free(tmem_base, TMEM_COLS);
}

// ------------------------------- pixel-pair variant (cin = 32) --------------------------------
// With cout = 64 the plain kernel issues N = 64 UMMAs, and a kind::f16 UMMA costs 64 clocks for any N <= 128
// (measured, tests/test_umma_probe.py / tools/probe_i8_timing.py): half of the tensor pipe idles.  Here one GEMM
// row computes TWO horizontally adjacent output pixels (x, x+1), as conv1.cu does for 2 x 2:
//     N = 2 * 64 (dx, co),   K = 3 rows * 4 pixels * 32 ch = 384,   B[(dx, co), (r, xw, c)] = W[co, c, r, xw - dx]
// 75 % of the issued MACs are useful and every UMMA runs at N = 128: 25 UMMAs per 256 output pixels instead of
// 2 * 19.  Two 32-channel pixels are exactly one 128-byte shared-memory row, so the input is fetched as pixel
// PAIRS (TMA box [64 elem, 10 pairs, 18 rows], 128B swizzle) and the 4-pixel window of a row pair is the 256
// contiguous bytes starting 64 bytes into the pair to its left -- again only descriptor start offsets, no im2col.
// The 2 x 2 max-pool is a max over the two column halves of one accumulator row (x) and one shuffle (y).
constexpr int CP_TILE_Y = 16, CP_TILE_XP = 8;                 // 16 rows x 8 pixel pairs = 256 output pixels
constexpr int CP_HALO_Y = CP_TILE_Y + 2, CP_HALO_XP = CP_TILE_XP + 2;
constexpr int CP_ROWB = 128;                                   // one pixel pair
constexpr int CP_STAGE_BYTES = CP_HALO_Y * CP_HALO_XP * CP_ROWB;           // 23040
constexpr int CP_STAGE_STRIDE = (CP_STAGE_BYTES + 1023) / 1024 * 1024;
constexpr int CP_N = 128, CP_K = 384;
constexpr int CP_B_CHUNK = CP_N * 128;                          // 64 k-elements of all 128 rows
constexpr int CP_B_BYTES = (CP_K / 64) * CP_B_CHUNK;            // 98304
constexpr int CP_NSTAGE = 4;
constexpr int CP_BIAS_BYTES = CP_N * 32;

__global__ void __launch_bounds__(CV_THREADS, 1)
conv3x3_pair_kernel(const __grid_constant__ CUtensorMap tmap_in, const uint8_t* __restrict__ w_packed,
                    const float* __restrict__ bias, __nv_bfloat16* __restrict__ out, int H, int W, int tiles_y,
                    int tiles_x, int total_tiles) {
  constexpr int COUT = 64;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* smem_b = smem;                               // CP_B_BYTES
  uint8_t* smem_a = smem + CP_B_BYTES;                  // CP_NSTAGE * CP_STAGE_STRIDE
  uint8_t* smem_ones = smem_a + CP_NSTAGE * CP_STAGE_STRIDE;
  uint8_t* smem_biasop = smem_ones + ONES_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_biasop + CP_BIAS_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + CP_NSTAGE;
  uint64_t* tfull_bar = bars + 2 * CP_NSTAGE;
  uint64_t* tempty_bar = bars + 2 * CP_NSTAGE + CV_NACC;
  uint64_t* wload_bar = bars + 2 * CP_NSTAGE + 2 * CV_NACC;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * CP_NSTAGE + 2 * CV_NACC + 1);
  volatile uint32_t* issued = tmem_slot + 1;            // [CP_MMA_WARPS] tiles issued per issuing warp (issue_gate)

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < CP_NSTAGE; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < CV_NACC; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 4);
    }
    mbar_init(wload_bar, 1);
    for (int i = 0; i < CP_MMA_WARPS; ++i) issued[i] = 0;
    fence_mbar_init();
    tma_prefetch_desc(&tmap_in);
  }
  if (warp == 2) tmem_alloc(tmem_slot, CV_NACC * CP_N);
  fill_ones_operand(smem_ones, threadIdx.x, blockDim.x);
  fill_bias_operand(smem_biasop, bias, CP_N, COUT, threadIdx.x, blockDim.x);    // row (dx, co) -> bias[co]
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = bcast0(*tmem_slot);
  // programmatic dependent launch (see sia_ptx.cuh): the prologue above and the resident weights overlap the previous
  // kernel's tail; the producer lane waits after it has issued the weight copies
  pdl_launch_dependents();
  if (!(warp == 0 && lane == 0)) pdl_wait();

  if (warp == 0) {
    // ================================ TMA producer ==========================================
    if (lane == 0) {
      mbar_arrive_expect_tx(wload_bar, CP_B_BYTES);
      for (int off = 0; off < CP_B_BYTES; off += 16384) bulk_load_1d(smem_b + off, w_packed + off, 16384, wload_bar);
      pdl_wait();                       // the input tiles below are the previous kernel's output
      int stage = 0;
      uint32_t phase = 0;
      TileWalker t(blockIdx.x, gridDim.x, tiles_x, tiles_y);
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, t.next()) {
        mbar_wait(&empty_bar[stage], phase ^ 1, 50);
        mbar_arrive_expect_tx(&full_bar[stage], CP_STAGE_BYTES);
        // coordinates: (element in the pair, pixel pair, row, image); the halo starts one pair / one row early
        tma_load_4d(smem_a + stage * CP_STAGE_STRIDE, &tmap_in, &full_bar[stage], 0, t.tx * CP_TILE_XP - 1,
                    t.ty * CP_TILE_Y - 1, t.n);
        if (++stage == CP_NSTAGE) { stage = 0; phase ^= 1; }
      }
      // Drain: the tcgen05.commit arrivals on the empty barriers of the last stages are asynchronous and nobody else
      // waits for them; the CTA must not exit (and hand its shared memory to the next kernel's CTA, which under
      // programmatic dependent launch is already queued) while one is in flight.
      for (int i = 0; i < CP_NSTAGE; ++i) {
        mbar_wait(&empty_bar[stage], phase ^ 1, 62);
        if (++stage == CP_NSTAGE) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp >= 1 && warp <= CP_MMA_WARPS) {
    // ================================ MMA issuer ============================================
    constexpr uint32_t idesc = make_idesc_bf16(128, CP_N);
    constexpr uint32_t a_hi = desc_hi(CP_HALO_XP * CP_ROWB, SW_128B);   // 8-row groups = tile rows, one halo row apart
    constexpr uint32_t b_hi = desc_hi(8 * 128, SW_128B);
    constexpr uint32_t c_hi = desc_hi(256, SW_NONE);
    const uint32_t a_lo0 = desc_lo(smem_u32(smem_a), 0);
    const uint32_t b_lo0 = desc_lo(smem_u32(smem_b), 0);
    const uint32_t ones_lo = desc_lo(smem_u32(smem_ones), 128);
    const uint32_t bias_lo = desc_lo(smem_u32(smem_biasop), 128);
    mbar_wait(wload_bar, 0, 51);
    // CP_MMA_WARPS issuing warps take the tiles round-robin (see conv3x3_kernel)
    for (int lt = warp - 1; blockIdx.x + (long long)lt * gridDim.x < total_tiles; lt += CP_MMA_WARPS) {
      const int stage = lt % CP_NSTAGE, acc = lt % CV_NACC;
      const uint32_t phase = (uint32_t)(lt / CP_NSTAGE) & 1u, acc_phase = (uint32_t)(lt / CV_NACC) & 1u;
      issue_gate(issued, lt, CP_NSTAGE, CP_MMA_WARPS, 63);    // sia_ptx.cuh: parity waits need the previous use issued
      if (CV_NACC != CP_NSTAGE) issue_gate(issued, lt, CV_NACC, CP_MMA_WARPS, 63);
      mbar_wait(&tempty_bar[acc], acc_phase ^ 1, 52);
      mbar_wait(&full_bar[stage], phase, 53);
      tc_fence_after_sync();
      if (elect_one()) {
        const uint32_t d_tmem = tmem_base + acc * CP_N;
        const uint32_t a_stage = a_lo0 + stage * (CP_STAGE_STRIDE >> 4);
        umma_bf16_ss_w(d_tmem, ones_lo, c_hi, bias_lo, c_hi, idesc, 0u);     // D = bias
#pragma unroll
        for (int r = 0; r < 3; ++r) {
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) {
            // window of output pair (row y, pair xp): halo row y + r, starting 64 bytes into halo pair xp
            const uint32_t a_lo = a_stage + ((r * CP_HALO_XP * CP_ROWB + 64 + kk * 32) >> 4);
            const uint32_t b_lo = b_lo0 + (((r * 2 + kk / 4) * CP_B_CHUNK + (kk % 4) * 32) >> 4);
            umma_bf16_ss_w(d_tmem, a_lo, a_hi, b_lo, b_hi, idesc, 1u);
          }
        }
        umma_commit(&empty_bar[stage]);
        umma_commit(&tfull_bar[acc]);
      }
      __syncwarp();
      if (lane == 0) issue_done(issued, lt, CP_MMA_WARPS);
    }
  } else if (warp >= 4) {
    // ================================ epilogue ==============================================
    const int group = (warp - 4) >> 2;
    const int e = (warp - 4) & 3;             // TMEM lanes 32e .. 32e+31 = tile rows 4e .. 4e+3
    const int Ho = H >> 1, Wo = W >> 1;
    const int ly = lane >> 3;
    const int xp = lane & 7;                  // pixel pair = pooled output column inside the tile
    const bool odd_y = (lane >> 3) & 1;
    TileWalker t(blockIdx.x + group * gridDim.x, 2 * gridDim.x, tiles_x, tiles_y);
    int j = group;
    for (int tile = blockIdx.x + group * gridDim.x; tile < total_tiles; tile += 2 * gridDim.x, t.next(), j += 2) {
      const int acc = j % CV_NACC;
      const uint32_t acc_phase = (j / CV_NACC) & 1;
      const int py = (t.ty * CP_TILE_Y + 4 * e + ly) >> 1;
      const int px = t.tx * CP_TILE_XP + xp;
      const bool in_range = py < Ho && px < Wo;
      __nv_bfloat16* opix = out + (((size_t)t.n * Ho + py) * Wo + px) * COUT;
      mbar_wait(&tfull_bar[acc], acc_phase, 54);
      tc_fence_after_sync();
      const uint32_t t_addr = tmem_base + ((uint32_t)(32 * e) << 16) + acc * CP_N;

      // 32 channels: max over dx (columns cb and 64 + cb), bf16, y partner (lane ^ 8) reduce-scatter, ReLU
      auto finish = [&](const uint32_t (&v0)[32], const uint32_t (&v1)[32], int cb) {
        uint32_t pk[16];
#pragma unroll
        for (int q = 0; q < 16; ++q) {
          pk[q] = pack_bf16x2(fmaxf(__uint_as_float(v0[2 * q]), __uint_as_float(v1[2 * q])),
                              fmaxf(__uint_as_float(v0[2 * q + 1]), __uint_as_float(v1[2 * q + 1])));
        }
        uint32_t h8[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const uint32_t keep = odd_y ? pk[8 + q] : pk[q];
          const uint32_t send = odd_y ? pk[q] : pk[8 + q];
          h8[q] = max_bf16x2(max_bf16x2(keep, __shfl_xor_sync(0xffffffffu, send, 8)), 0u);
        }
        if (in_range) {
          uint4* dst = reinterpret_cast<uint4*>(opix + cb + (odd_y ? 16 : 0));
          dst[0] = make_uint4(h8[0], h8[1], h8[2], h8[3]);
          dst[1] = make_uint4(h8[4], h8[5], h8[6], h8[7]);
        }
      };

      uint32_t a0[32], a1[32], b0[32], b1[32];
      tmem_ld32(t_addr, a0);
      tmem_ld32(t_addr + 64, a1);
      tmem_ld_wait();
      tmem_ld32(t_addr + 32, b0);
      tmem_ld32(t_addr + 96, b1);
      finish(a0, a1, 0);
      tmem_ld_wait();
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);     // the whole accumulator is in registers
      finish(b0, b1, 32);
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 2) tmem_free(tmem_base, CV_NACC * CP_N);
}

// [64][32][3][3] fp32 -> B[n = dx*64 + co][k = r*128 + xw*32 + c] = W[co][c][r][xw - dx] (0 outside 0..2), bf16, as six
// 128-row x 128-byte blocks (64 k each) with the 16-byte units of every row XOR-swizzled by the row index.
__global__ void pack_conv3x3_pair_kernel(const float* __restrict__ w, int cin, int cout,
                                         __nv_bfloat16* __restrict__ dst) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < CP_N * CP_K; i += gridDim.x * blockDim.x) {
    const int n = i / CP_K, k = i % CP_K;
    const int dx = n / 64, co = n % 64;
    const int r = k / 128, xw = (k % 128) / 32, c = k % 32;
    const int s = xw - dx;
    float v = 0.f;
    if (s >= 0 && s < 3 && co < cout && c < cin) v = w[((co * cin + c) * 3 + r) * 3 + s];
    const int chunk = k / 64, kq = k % 64;
    const int unit = (kq / 8) ^ (n & 7);
    dst[(size_t)chunk * (CP_B_CHUNK / 2) + n * 64 + unit * 8 + (kq & 7)] = __float2bfloat16_rn(v);
  }
}

static int launch_conv3x3_pair(const void* in, int batch, int h, int w, const void* w_packed, const float* bias,
                               void* out, cudaStream_t st) {
  if (int wrc = ensure_watchdog()) return wrc;
  CUtensorMap tmap;
  // the NHWC input seen as pixel pairs: [64 elements, w/2 pairs, h, batch]
  const uint64_t dims[4] = {64, (uint64_t)w / 2, (uint64_t)h, (uint64_t)batch};
  const uint64_t strides[3] = {128, (uint64_t)w * 64, (uint64_t)h * w * 64};
  const uint32_t box[4] = {64, CP_HALO_XP, CP_HALO_Y, 1};
  int rc = encode_tmap_bf16(&tmap, in, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc != 0) return rc;
  const int tiles_y = (h + CP_TILE_Y - 1) / CP_TILE_Y;
  const int tiles_x = (w / 2 + CP_TILE_XP - 1) / CP_TILE_XP;
  const int total = tiles_y * tiles_x * batch;
  const int smem = 1024 + CP_B_BYTES + CP_NSTAGE * CP_STAGE_STRIDE + ONES_BYTES + CP_BIAS_BYTES +
                   (2 * CP_NSTAGE + 2 * CV_NACC + 4) * 8;
  static SmemSlots configured = {};
  if (int rc2 = ensure_dynamic_smem(conv3x3_pair_kernel, smem, &configured)) return rc2;
  const int grid = total < sm_count() ? total : sm_count();
  return launch_kernel(conv3x3_pair_kernel, dim3(grid), dim3(CV_THREADS), smem, st, true, tmap, static_cast<const uint8_t*>(w_packed), bias,
                                                      static_cast<__nv_bfloat16*>(out), h, w, tiles_y, tiles_x, total);
  return launch_status();
}

// ------------------------------- weight packing ----------------------------------------------
// [cout][cin][3][3] fp32 -> for tap (r,s), chunk kc: [cout rows][CK channels] bf16, K-major, with the
// 16-byte units of each row XOR-swizzled by the row index exactly as TMA / UMMA swizzle modes do
// (128B rows: unit ^= row%8 ; 64B rows: unit ^= (row/2)%4).
// cin_pad / cout_pad: channel counts of the (zero-padded) activation buffers; weights outside [cout][cin] are zero.
__global__ void pack_conv3x3_kernel(const float* __restrict__ w, int cin, int cout, int cin_pad, int cout_pad,
                                    __nv_bfloat16* __restrict__ dst) {
  const int ck = cin_pad < 64 ? cin_pad : 64;
  const int nchunk = cin_pad / ck;
  const int total = 9 * cin_pad * cout_pad;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int c_in_chunk = i % ck;
    const int row = (i / ck) % cout_pad;
    const int kc = (i / (ck * cout_pad)) % nchunk;
    const int tap = i / (ck * cout_pad * nchunk);
    const int r = tap / 3, s = tap % 3;
    const int c = kc * ck + c_in_chunk;
    const float v = (row < cout && c < cin) ? w[(((size_t)row * cin + c) * 3 + r) * 3 + s] : 0.f;
    const int unit = c_in_chunk / 8;
    const int swz = ck == 64 ? (unit ^ (row & 7)) : (unit ^ ((row >> 1) & 3));
    const size_t block = ((size_t)tap * nchunk + kc) * cout_pad * ck;
    dst[block + (size_t)row * ck + swz * 8 + (c_in_chunk & 7)] = __float2bfloat16_rn(v);
  }
}

template <int CIN, int COUT, int NSTAGE>
static int launch_conv3x3(const void* in, int batch, int h, int w, const void* w_packed, const float* bias, void* out,
                          cudaStream_t st) {
  using C = ConvCfg<CIN, COUT>;
  if (int wrc = ensure_watchdog()) return wrc;
  CUtensorMap tmap;
  const uint64_t dims[4] = {(uint64_t)CIN, (uint64_t)w, (uint64_t)h, (uint64_t)batch};
  const uint64_t strides[3] = {(uint64_t)CIN * 2, (uint64_t)w * CIN * 2, (uint64_t)h * w * CIN * 2};
  const uint32_t box[4] = {(uint32_t)C::CK, CV_HALO_X, CV_HALO_Y, 1};
  int rc = encode_tmap_bf16(&tmap, in, 4, dims, strides, box,
                            C::ROWB == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B);
  if (rc != 0) return rc;
  const int tiles_y = (h + CV_TILE_Y - 1) / CV_TILE_Y;
  const int tiles_x = (w + CV_TILE_X - 1) / CV_TILE_X;      // TMA zero-fills past the image: 'same' padding
  const int total = tiles_y * tiles_x * batch;
  const int smem = 1024 + C::B_BYTES + NSTAGE * C::STAGE_STRIDE + ONES_BYTES + C::BIAS_BYTES +
                   (2 * NSTAGE + 2 * CV_NACC + 4) * 8;
  auto kern = conv3x3_kernel<CIN, COUT, NSTAGE>;
  static SmemSlots configured = {};
  if (int rc2 = ensure_dynamic_smem(kern, smem, &configured)) return rc2;
  const int grid = total < sm_count() ? total : sm_count();
  return launch_kernel(kern, dim3(grid), dim3(CV_THREADS), smem, st, true, tmap, static_cast<const uint8_t*>(w_packed), bias,
                                       static_cast<__nv_bfloat16*>(out), h, w, tiles_y, tiles_x, total);
  return launch_status();
}


template <int COUT, int TILES>
static int launch_conv3x3_stream_t(const CUtensorMap& tmap, int nchunk, int batch, int h, int w, const void* w_packed,
                                   const float* bias, void* out, cudaStream_t st) {
  const int tiles_y = (h + CV_TILE_Y - 1) / CV_TILE_Y;
  const int tiles_x = (w + CV_TILE_X - 1) / CV_TILE_X;
  const int total = tiles_y * tiles_x * batch;
  const int passes = (total + TILES - 1) / TILES;
  const int smem = 1024 + CS_NB * COUT * 128 + TILES * nchunk * CS_CHUNK_STRIDE + ONES_BYTES + COUT * 32 +
                   (2 * CS_NB + 6) * 8;
  auto kern = conv3x3_stream_kernel<COUT, TILES>;
  static SmemSlots configured = {};
  if (int rc2 = ensure_dynamic_smem(kern, smem, &configured)) return rc2;
  const int grid = passes < sm_count() ? passes : sm_count();
  return launch_kernel(kern, dim3(grid), dim3(CV_THREADS), smem, st, true, tmap, static_cast<const uint8_t*>(w_packed), bias,
                                       static_cast<__nv_bfloat16*>(out), h, w, tiles_y, tiles_x, total, nchunk);
  return launch_status();
}

// cin_pad: channels of the NHWC input buffer (multiple of 64, <= 256); cout_pad in {64, 128, 192, 256}
static int launch_conv3x3_stream(const void* in, int batch, int h, int w, int cin_pad, int cout_pad,
                                 const void* w_packed, const float* bias, void* out, cudaStream_t st) {
  if (cin_pad % 64 != 0 || cin_pad < 64 || cin_pad > 256 || cout_pad % 64 != 0 || cout_pad < 64 || cout_pad > 256)
    return SIA_E_UNSUPPORTED;
  if (int wrc = ensure_watchdog()) return wrc;
  const int nchunk = cin_pad / 64;
  CUtensorMap tmap;
  const uint64_t dims[4] = {(uint64_t)cin_pad, (uint64_t)w, (uint64_t)h, (uint64_t)batch};
  const uint64_t strides[3] = {(uint64_t)cin_pad * 2, (uint64_t)w * cin_pad * 2, (uint64_t)h * w * cin_pad * 2};
  const uint32_t box[4] = {64, CV_HALO_X, CV_HALO_Y, 1};
  int rc = encode_tmap_bf16(&tmap, in, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc != 0) return rc;
  // two resident tiles whenever both halo sets and the weight ring fit in 227 KB of shared memory
  const bool two = 1024 + CS_NB * cout_pad * 128 + 2 * nchunk * CS_CHUNK_STRIDE + ONES_BYTES + cout_pad * 32 + 128 <=
                   227 * 1024;
  switch (cout_pad) {
    case 64:
      return two ? launch_conv3x3_stream_t<64, 2>(tmap, nchunk, batch, h, w, w_packed, bias, out, st)
                 : launch_conv3x3_stream_t<64, 1>(tmap, nchunk, batch, h, w, w_packed, bias, out, st);
    case 128:
      return two ? launch_conv3x3_stream_t<128, 2>(tmap, nchunk, batch, h, w, w_packed, bias, out, st)
                 : launch_conv3x3_stream_t<128, 1>(tmap, nchunk, batch, h, w, w_packed, bias, out, st);
    case 192:
      return two ? launch_conv3x3_stream_t<192, 2>(tmap, nchunk, batch, h, w, w_packed, bias, out, st)
                 : launch_conv3x3_stream_t<192, 1>(tmap, nchunk, batch, h, w, w_packed, bias, out, st);
    default:
      return two ? launch_conv3x3_stream_t<256, 2>(tmap, nchunk, batch, h, w, w_packed, bias, out, st)
                 : launch_conv3x3_stream_t<256, 1>(tmap, nchunk, batch, h, w, w_packed, bias, out, st);
  }
}

}  // namespace sia

// (32, 64) uses the pixel-pair operand (conv3x3_pair_kernel); SIA_CONV2_PLAIN=1 in the environment keeps the plain
// 9-tap operand and kernel for A/B comparisons.
static bool use_pair_variant(int cin, int cout) {
  static int plain = -1;
  if (plain < 0) {
    const char* e = getenv("SIA_CONV2_PLAIN");
    plain = (e != nullptr && e[0] == '1') ? 1 : 0;
  }
  return cin == 32 && cout == 64 && plain == 0;
}

extern "C" size_t sia_pack_conv3x3_bytes(int cin, int cout) {
  return use_pair_variant(cin, cout) ? (size_t)sia::CP_B_BYTES : (size_t)9 * cin * cout * 2;
}

extern "C" int sia_pack_conv3x3_padded(const float* w_oihw, int cin, int cout, int cin_pad, int cout_pad, void* packed,
                                       void* stream) {
  using namespace sia;
  SIA_REQUIRE(w_oihw && packed && cin >= 1 && cout >= 1 && cin_pad >= cin && cout_pad >= cout);
  if (!((cin_pad == 32) || (cin_pad % 64 == 0)) || cout_pad % 8 != 0) return SIA_E_UNSUPPORTED;
  if (use_pair_variant(cin_pad, cout_pad)) {
    pack_conv3x3_pair_kernel<<<(CP_N * CP_K + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        w_oihw, cin, cout, static_cast<__nv_bfloat16*>(packed));
    return launch_status();
  }
  const int total = 9 * cin_pad * cout_pad;
  pack_conv3x3_kernel<<<(total + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      w_oihw, cin, cout, cin_pad, cout_pad, static_cast<__nv_bfloat16*>(packed));
  return launch_status();
}

extern "C" int sia_pack_conv3x3(const float* w_oihw, int cin, int cout, void* packed, void* stream) {
  return sia_pack_conv3x3_padded(w_oihw, cin, cout, cin, cout, packed, stream);
}

extern "C" int sia_conv3x3_relu_pool2(const void* in_nhwc, int batch, int h, int w, int cin, int cout,
                                      const void* w_packed, const float* bias, void* out_nhwc, void* stream) {
  using namespace sia;
  SIA_REQUIRE(in_nhwc && w_packed && bias && out_nhwc && batch >= 1 && h >= 2 && w >= 2);
  SIA_REQUIRE(aligned(in_nhwc, 16) && aligned(w_packed, 16) && aligned(out_nhwc, 16));
  if (h % 2 != 0 || w % 2 != 0) return SIA_E_UNSUPPORTED;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (use_pair_variant(cin, cout)) return launch_conv3x3_pair(in_nhwc, batch, h, w, w_packed, bias, out_nhwc, st);
  if (cin == 32 && cout == 64) return launch_conv3x3<32, 64, 6>(in_nhwc, batch, h, w, w_packed, bias, out_nhwc, st);
  if (cin == 64 && cout == 128) return launch_conv3x3<64, 128, 3>(in_nhwc, batch, h, w, w_packed, bias, out_nhwc, st);
  // everything else with 64-channel-chunked buffers: the streamed-weight kernel (cin, cout = buffer widths)
  return launch_conv3x3_stream(in_nhwc, batch, h, w, cin, cout, w_packed, bias, out_nhwc, st);
}
