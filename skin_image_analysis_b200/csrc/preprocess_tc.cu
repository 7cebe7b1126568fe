// K1-K3 on the tensor cores: the fast path of sia_preprocess_u8hwc for the padded NHWC4 bf16 layout.
//
// Same operator as preprocess.cu (out = Wy * img * Wx^T, the reference's float32(u8)/255 +
// skimage.transform.resize + HWC->CHW, tone_bias_dataset.py:335, :425, :470), split differently:
//
//   vertical pass   V[i, k] = sum_r Wy[i, r] * S[r, k]        i = output row, r = source row,
//                                                            k = byte of the image row (3*x + c)
//       = ONE GEMM per (image, tile of <=128 output rows): tcgen05.mma kind::f16, fp16 x fp16 -> fp32.
//         A = the tile's Wy rows (fp16, K-major, resident in shared memory, K = 256 source rows),
//         B = the image itself: TMA brings raw 128-row x 160-byte windows of the u8 rows into a shared-memory
//             ring (the decode buffer is viewed as double rows of 2*row_bytes so that the row pitch is a
//             multiple of 16 bytes; a row whose window starts 8 bytes off 16-byte alignment is fetched 8 bytes
//             early and read back shifted), converter warps widen the bytes to fp16 (PRMT + HSUB2,
//             exact) and store them as the MN-major operand (the byte index k is the contiguous one,
//             so the HWC decode buffer needs no transposition),
//         D = 128 lanes (output rows) x 128 columns (bytes 120b .. 120b+127 of the row) fp32 in TMEM; four of
//             them: two images are in flight per CTA (one per horizontal-pass warp group), double buffered.
//   horizontal pass  out[i, j, c] = sum_x Wx[j, x] * V[i, 3x + c]
//       = the epilogue: thread = TMEM lane = output row.  It executes a static schedule of "items" built
//         on the host (resize_weights.build_tc_tables): item n loads 9 accumulator columns (<= 3 source
//         pixels), adds them into 4 accumulator slots with the item's 3 x 4 weights, then emits one output
//         pixel from slot n % 4 (scale, bias, bf16, 8 bytes of NHWC4) and clears it.  Four items are
//         unrolled, so every slot index is a compile-time constant: straight-line FFMA code.  Two images
//         are in flight per CTA, one per warp group and TMEM accumulator, because a single warp per
//         scheduler cannot hide the TMEM-load latency.
//
// Why: the CUDA-core kernel is issue-bound (~150 instructions per source row x output column, most of
// them unpacking interleaved bytes).  Here the 6-tap vertical contraction -- the one with the larger
// reduction -- costs no issue slots at all, and the horizontal pass touches every V value once.
// HBM traffic is unchanged: each source byte is read once, each output byte written once.
//
// Precision: fp16 weights (11 significant bits; row sums renormalised per lane), exact products, fp32
// accumulation => <= 1e-4 of full scale; used for the bf16 layout only (tests: <= 1 bf16 ulp).
#include <cuda_fp16.h>

#include "sia_host.cuh"
#include "sia_ptx.cuh"

namespace sia {

constexpr int TC_KWIN = 256;                     // source rows per tile window (K of the GEMM)
constexpr int TC_KSTAGE = 128;                   // source rows per pipeline stage
constexpr int TC_NQ = TC_KWIN / TC_KSTAGE;       // stages per column block
constexpr int TC_STRIDE = 120;                   // bytes of the image row between column blocks (40 pixels)
constexpr int TC_COLS = 128;                     // accumulator columns per block (8 bytes of overlap: one item)
constexpr int TC_UNITS = TC_COLS / 8;            // 16-byte fp16 units along N per stage row
constexpr int TC_B_LBO = 128;                    // next group of 8 source rows
constexpr int TC_B_SBO = (TC_KSTAGE / 8) * 128 + 16;   // next 8 bytes of the row (+16: bank spread for the converter)
constexpr int TC_STAGE_BYTES = ((TC_UNITS * TC_B_SBO + 127) / 128) * 128;
constexpr int TC_NSTAGE = 2;                     // fp16 operand stages (converter -> MMA)
constexpr int TC_NRAW = 3;                       // raw u8 stages (TMA -> converter); fewer when the item table is large
constexpr int TC_RAW_ROWB = TC_COLS + 32;        // bytes per raw row: a row may have to be fetched up to 8 bytes early
constexpr int TC_RAW_HALF = (TC_KSTAGE / 2) * TC_RAW_ROWB;   // even-row box, then odd-row box
constexpr int TC_RAW_BYTES = 2 * TC_RAW_HALF;
constexpr int TC_A_BYTES = 128 * TC_KWIN * 2;    // 65536
constexpr int TC_A_LBO = 128, TC_A_SBO = (TC_KWIN / 8) * 128;
constexpr int TC_THREADS = 448;                  // warps 0-3 converters, 4-7 / 8-11 horizontal-pass groups 0 / 1,
constexpr int TC_WARP_MMA = 12;                  // 12 MMA issuer (+ TMEM alloc), 13 TMA producer
constexpr int TC_WARP_TMA = 13;                  // (16 warps x 128 registers is what one SM sub-partition quartet holds)
constexpr int TC_ITEM_BYTES = 64;                // int4 {column, emit, block, vector column} + 3 x float4 weights

struct TcParams {
  const uint8_t* a_packed;      // [n_tiles][TC_A_BYTES]
  const float* lane_scale;      // [n_tiles][128]
  const int32_t* tile_row0;     // [n_tiles]
  const uint4* items;           // [n_items][4]
  uint2* dst;                   // [batch][out_h][out_w + 8] NHWC4 bf16
  int batch, src_h, src_w, out_h, out_w;
  int n_tiles, tile_rows, n_blocks, last_block_cols, n_items;
  int odd_shift;                // row_bytes % 16: how far the start of an odd row is from 16-byte alignment
  int n_raw;                    // raw stages in use (<= TC_NRAW)
  int pads_in_schedule;         // the vector-store groups of the schedule also write the zero pad columns
  float scale[3], bias[3];
};

// kind::f16, fp16 x fp16 -> fp32, A K-major, B MN-major
__host__ __device__ constexpr uint32_t tc_idesc(uint32_t n) {
  return (1u << 4) | (1u << 16) | ((n >> 3) << 17) | ((128u >> 4) << 24);
}

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t* v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_ld1(uint32_t taddr, uint32_t& v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld16p(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}

// Work of one CTA: its tile (blockIdx.x % n_tiles) of the images img0, img0 + img_step, ...  Two images are in
// flight at a time -- image 2p + g of the CTA's sequence belongs to horizontal-pass warp group g and to TMEM
// accumulator g -- and every role walks the same order: for pair p, for block b, for g in {0, 1}.
__global__ void __launch_bounds__(TC_THREADS, 1)
preprocess_tc_kernel(const __grid_constant__ CUtensorMap tmap_src, const __grid_constant__ TcParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* smem_a = smem;                                        // TC_A_BYTES
  uint8_t* smem_raw_ring = smem + TC_A_BYTES;                    // n_raw * TC_RAW_BYTES
  uint8_t* smem_b = smem_raw_ring + p.n_raw * TC_RAW_BYTES;      // TC_NSTAGE * TC_STAGE_BYTES
  uint4* items_s = reinterpret_cast<uint4*>(smem_b + TC_NSTAGE * TC_STAGE_BYTES);   // [n_items][4]
  uint64_t* bars = reinterpret_cast<uint64_t*>(items_s + 4 * p.n_items);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + TC_NSTAGE;
  uint64_t* tfull_bar = bars + 2 * TC_NSTAGE;          // [4]: accumulator 2g + b of warp group g
  uint64_t* tempty_bar = bars + 2 * TC_NSTAGE + 4;     // [4]
  uint64_t* a_bar = bars + 2 * TC_NSTAGE + 8;
  uint64_t* raw_full = bars + 2 * TC_NSTAGE + 9;
  uint64_t* raw_empty = raw_full + TC_NRAW;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(raw_empty + TC_NRAW);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int tile = blockIdx.x % p.n_tiles;
  const int img0 = blockIdx.x / p.n_tiles;
  const int img_step = gridDim.x / p.n_tiles;
  const int n_img = img0 < p.batch ? (p.batch - img0 + img_step - 1) / img_step : 0;   // images of this CTA
  const int row_bytes = p.src_w * 3;
  const int row0 = p.tile_row0[tile];

  if (threadIdx.x == 0) {
    for (int i = 0; i < TC_NSTAGE; ++i) {
      mbar_init(&full_bar[i], 4);      // one arrival per converter warp
      mbar_init(&empty_bar[i], 1);     // tcgen05.commit
    }
    for (int i = 0; i < 4; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 4);    // one arrival per warp of the group
    }
    for (int i = 0; i < TC_NRAW; ++i) {
      mbar_init(&raw_full[i], 1);      // TMA transaction bytes
      mbar_init(&raw_empty[i], 4);     // one arrival per converter warp
    }
    mbar_init(a_bar, 1);
    fence_mbar_init();
    tma_prefetch_desc(&tmap_src);
  }
  if (warp == TC_WARP_MMA) tmem_alloc(tmem_slot, 512);
  for (int i = threadIdx.x; i < 4 * p.n_items; i += blockDim.x) items_s[i] = p.items[i];
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = bcast0(*tmem_slot);

#ifdef SIA_TC_EPI_ONLY
  if (warp >= 4 && warp < 12) {} else { goto tc_done; }
#endif
  if (warp == TC_WARP_TMA) {
    // ================================ TMA producer ==========================================
    if (lane == 0) {
      mbar_arrive_expect_tx(a_bar, TC_A_BYTES);
      for (int off = 0; off < TC_A_BYTES; off += 16384)
        bulk_load_1d(smem_a + off, p.a_packed + (size_t)tile * TC_A_BYTES + off, 16384, a_bar);
      int rs = 0;
      uint32_t rphase = 0;
      for (int k0 = 0; k0 < n_img; k0 += 2) {
        for (int blk = 0; blk < p.n_blocks; ++blk) {
          for (int g = 0; g < 2 && k0 + g < n_img; ++g) {
            const int img = img0 + (k0 + g) * img_step;
            const int seq = ((k0 >> 1) * p.n_blocks + blk) * 2 + g;
            // TMA needs 16-byte aligned starts: a row whose window starts 8 bytes off is fetched 8 bytes early
            const int col = blk * TC_STRIDE;
            const int mis_even = col & 15, mis_odd = (col + p.odd_shift) & 15;
            for (int q = 0; q < TC_NQ; ++q) {
              const int r0 = row0 + q * TC_KSTAGE;           // first source row of the stage
              const int even0 = (r0 + 1) >> 1;               // double row of the first even / odd row
              const int odd0 = r0 >> 1;
              mbar_wait(&raw_empty[rs], rphase ^ 1, 45);
              if (q == 0) trace(seq, 0);
              mbar_arrive_expect_tx(&raw_full[rs], TC_RAW_BYTES);
              uint8_t* dst = smem_raw_ring + rs * TC_RAW_BYTES;
              // innermost coordinate in 16-bit elements (the tensor map views the bytes as u16 pairs)
              tma_load_3d(dst, &tmap_src, &raw_full[rs], (col - mis_even) >> 1, even0, img);
              tma_load_3d(dst + TC_RAW_HALF, &tmap_src, &raw_full[rs], (row_bytes + col - mis_odd) >> 1, odd0, img);
              if (q == TC_NQ - 1) trace(seq, 1);
              if (++rs == p.n_raw) { rs = 0; rphase ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp < 4) {
    // ================================ converters ============================================
    // Per 128-row stage warp w converts rows 8*it + 2*w + (lane >> 4), it = 0..15; lanes 0-15 / 16-31 own the 16
    // 8-byte pieces of an even / odd row of the pair.
    const int half = lane >> 4, unit = lane & 15;
    const int rr0 = 2 * warp + half;                           // this lane's row inside the stage for it = 0
    const uint32_t parity = (uint32_t)(row0 + rr0) & 1u;       // absolute parity of this lane's rows
    const uint32_t ld_lane = parity * TC_RAW_HALF + (uint32_t)(rr0 >> 1) * TC_RAW_ROWB + (uint32_t)unit * 8u;
    const uint32_t st_lane = (uint32_t)unit * TC_B_SBO + (uint32_t)(rr0 & 7) * 16u;
    const __half2 k1024 = __half2half2(__ushort_as_half((unsigned short)0x6400));
    int stage = 0, rs = 0;
    uint32_t phase = 0, rphase = 0;
    int it0 = 0;
    for (int k0 = 0; k0 < n_img; k0 += 2) {
      for (int blk = 0; blk < p.n_blocks; ++blk) {
        const int col = blk * TC_STRIDE;
        const uint32_t mis = (uint32_t)((col + (parity ? p.odd_shift : 0)) & 15);   // this row was fetched `mis` bytes early
        for (int g = 0; g < 2 && k0 + g < n_img; ++g) {
          for (int q = 0; q < TC_NQ; ++q, ++it0) {
            mbar_wait(&raw_full[rs], rphase, 46);
            if (warp == 0 && lane == 0 && q == 0) trace(it0 / TC_NQ, 2);
            uint2 v[16];
#ifdef SIA_TC_NOCONV
            if (it0 < 0)                             // timing experiment: barriers only, operands are whatever is in smem
#endif
            {
              const uint32_t src = smem_u32(smem_raw_ring) + rs * TC_RAW_BYTES + ld_lane + mis;
#pragma unroll
              for (int it = 0; it < 16; ++it) {              // row 8*it + rr0 -> index 4*it + (rr0 >> 1) of its parity box
                asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];"
                             : "=r"(v[it].x), "=r"(v[it].y)
                             : "r"(src + (uint32_t)(4 * it) * TC_RAW_ROWB));
              }
            }
            mbar_wait(&empty_bar[stage], phase ^ 1, 40);
#ifdef SIA_TC_NOCONV
            if (it0 < 0)
#endif
            {
              const uint32_t base = smem_u32(smem_b) + stage * TC_STAGE_BYTES + st_lane;
#pragma unroll
              for (int it = 0; it < 16; ++it) {              // row 8*it + rr0: core-matrix group `it`, row rr0 of it
                // u8 -> fp16, exact: byte b becomes the half 0x6400 | b = 1024 + b, then subtract 1024
                uint32_t h[4];
                h[0] = __byte_perm(v[it].x, 0x64646464u, 0x4140);
                h[1] = __byte_perm(v[it].x, 0x64646464u, 0x4342);
                h[2] = __byte_perm(v[it].y, 0x64646464u, 0x4140);
                h[3] = __byte_perm(v[it].y, 0x64646464u, 0x4342);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  __half2 x = __hsub2(*reinterpret_cast<__half2*>(&h[k]), k1024);
                  h[k] = *reinterpret_cast<uint32_t*>(&x);
                }
                const uint32_t dst = base + (uint32_t)it * TC_B_LBO;
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(h[0]), "r"(h[1]), "r"(h[2]),
                             "r"(h[3])
                             : "memory");
              }
            }
            fence_proxy_async_smem();      // generic-proxy stores -> visible to the UMMA operand reads
            __syncwarp();
            if (lane == 0) {
              mbar_arrive(&full_bar[stage]);
              mbar_arrive(&raw_empty[rs]);
              if (warp == 0 && q == TC_NQ - 1) trace(it0 / TC_NQ, 3);
            }
            if (++stage == TC_NSTAGE) { stage = 0; phase ^= 1; }
            if (++rs == p.n_raw) { rs = 0; rphase ^= 1; }
          }
        }
      }
    }
  } else if (warp == TC_WARP_MMA) {
    // ================================ MMA issuer ============================================
    constexpr uint32_t a_hi = desc_hi(TC_A_SBO, SW_NONE);
    constexpr uint32_t b_hi = desc_hi(TC_B_SBO, SW_NONE);
    const uint32_t a_lo0 = desc_lo(smem_u32(smem_a), TC_A_LBO);
    const uint32_t b_lo0 = desc_lo(smem_u32(smem_b), TC_B_LBO);
    mbar_wait(a_bar, 0, 41);
    int stage = 0;
    uint32_t phase = 0;
    uint32_t acc_cnt[2] = {0u, 0u};                 // blocks issued so far for warp group g: buffer = cnt & 1
    for (int k0 = 0; k0 < n_img; k0 += 2) {
      for (int blk = 0; blk < p.n_blocks; ++blk) {
        const uint32_t idesc = tc_idesc((uint32_t)(blk == p.n_blocks - 1 ? p.last_block_cols : TC_COLS));
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          if (k0 + g < n_img) {
            const int seq = ((k0 >> 1) * p.n_blocks + blk) * 2 + g;
            const int buf = 2 * g + (int)(acc_cnt[g] & 1u);
#ifndef SIA_TC_FREE
            mbar_wait(&tempty_bar[buf], ((acc_cnt[g] >> 1) & 1u) ^ 1u, 42);
#endif
            ++acc_cnt[g];
            tc_fence_after_sync();
            if (lane == 0) trace(seq, 4);
            const uint32_t d_tmem = tmem_base + buf * TC_COLS;
            for (int q = 0; q < TC_NQ; ++q) {
              mbar_wait(&full_bar[stage], phase, 43);
              tc_fence_after_sync();
              if (elect_one()) {
                const uint32_t b_stage = b_lo0 + stage * (TC_STAGE_BYTES >> 4);
#pragma unroll
                for (int kk = 0; kk < TC_KSTAGE / 16; ++kk) {
                  // K = 16 source rows per instruction: two 8-row core-matrix groups of either operand
                  umma_bf16_ss_w(d_tmem, a_lo0 + (q * (TC_KSTAGE / 16) + kk) * ((2 * TC_A_LBO) >> 4), a_hi,
                                 b_stage + kk * ((2 * TC_B_LBO) >> 4), b_hi, idesc, (q | kk) ? 1u : 0u);
                }
                umma_commit(&empty_bar[stage]);
                if (q == TC_NQ - 1) umma_commit(&tfull_bar[buf]);
              }
              __syncwarp();
              if (++stage == TC_NSTAGE) { stage = 0; phase ^= 1; }
            }
            if (lane == 0) trace(seq, 5);
          }
        }
      }
    }
  } else if (warp < 12) {
    // ================================ horizontal pass =======================================
    const int g = (warp - 4) >> 2;                 // warp group = accumulator = image parity in the CTA's sequence
    const int e = warp & 3;                        // TMEM lanes 32e .. 32e+31
    const int l = 32 * e + lane;                   // output row inside the tile
    const int i = tile * p.tile_rows + l;
    const bool row_ok = l < p.tile_rows && i < p.out_h;
    const float ls = p.lane_scale[tile * 128 + l];
    const float sc0 = p.scale[0] * ls, sc1 = p.scale[1] * ls, sc2 = p.scale[2] * ls;
    const float bi0 = p.bias[0], bi1 = p.bias[1], bi2 = p.bias[2];
    const int pitch = p.out_w + SIA_NHWC4_PAD;
    const uint32_t t_lanes = tmem_base + ((uint32_t)(32 * e) << 16) + 2 * g * TC_COLS;
    uint32_t acc_cnt = 0;                          // blocks acquired so far: accumulator 2g + ((cnt - 1) & 1) is current
    uint32_t t_acc = t_lanes;
    for (int k = g; k < n_img; k += 2) {
      const int img = img0 + k * img_step;
      uint2* orow = p.dst + ((size_t)img * p.out_h + (row_ok ? i : 0)) * pitch;
      if (row_ok && !p.pads_in_schedule) {         // zero pad columns of the NHWC4 row
        orow[0] = make_uint2(0u, 0u);
#pragma unroll
        for (int c = 1; c < SIA_NHWC4_PAD; ++c) orow[p.out_w + c] = make_uint2(0u, 0u);
      }
      float acc[4][3];
#pragma unroll
      for (int s = 0; s < 4; ++s) acc[s][0] = acc[s][1] = acc[s][2] = 0.f;
      int cur_block = -1;

      // makes `block` the accumulator being read: every block is waited for and handed back in order
      auto acquire = [&](int block) {
        while (cur_block < block) {
          if (cur_block >= 0) {
            tmem_ld_wait();                        // loads of the block being handed back may still be in flight
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty_bar[2 * g + ((acc_cnt - 1) & 1u)]);
            if (e == 0 && lane == 0) trace(((k >> 1) * p.n_blocks + cur_block) * 2 + g, 7);
          }
          const uint32_t b = acc_cnt & 1u;
#ifndef SIA_TC_FREE
          mbar_wait(&tfull_bar[2 * g + b], (acc_cnt >> 1) & 1u, 44);
#endif
          ++acc_cnt;
          t_acc = t_lanes + b * TC_COLS;
          tc_fence_after_sync();
          ++cur_block;
          if (e == 0 && lane == 0) trace(((k >> 1) * p.n_blocks + cur_block) * 2 + g, 6);
        }
      };

      // Items are processed in groups of 4 (n_items is a multiple of 8, padded with no-op items).  tcgen05.wait::ld
      // waits for EVERY outstanding load, so loads are pipelined a whole group ahead: the 4 x 9 accumulator columns
      // of group m+1 are requested right after the wait that delivers group m and land while group m is computed
      // (~36 FFMA + one predicated 8-byte store per item).
      uint32_t va[4][9], vb[4][9];
      auto load_group = [&](int n0, uint32_t (&v)[4][9]) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const uint4 info = items_s[4 * (n0 + u)];
          if ((int)info.z != cur_block) acquire((int)info.z);
          tmem_ld8(t_acc + info.x, v[u]);
          tmem_ld1(t_acc + info.x + 8, v[u][8]);
        }
      };
      auto compute_group = [&](int n0, const uint32_t (&v)[4][9]) {
        float4 w_next[3];
        int emit_next;
        const uint4 i0 = items_s[4 * n0];
        const int vec_col = (int)i0.w;               // >= 0: the group's 4 pixels are padded columns vec_col .. +3
        emit_next = (int)i0.y;
#pragma unroll
        for (int q = 0; q < 3; ++q) w_next[q] = *reinterpret_cast<const float4*>(&items_s[4 * n0 + 1 + q]);
        uint32_t st[4][2];                           // the group's pixels: one full 32-byte sector per lane
        int js[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float4 w0 = w_next[0], w1 = w_next[1], w2 = w_next[2];
          const int j = emit_next;
          js[u] = j;
          if (u < 3) {                               // next item's table entry in flight during this item's FFMAs
            emit_next = (int)items_s[4 * (n0 + u + 1)].y;
#pragma unroll
            for (int q = 0; q < 3; ++q) w_next[q] = *reinterpret_cast<const float4*>(&items_s[4 * (n0 + u + 1) + 1 + q]);
          }
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            const float f0 = __uint_as_float(v[u][c]), f1 = __uint_as_float(v[u][3 + c]),
                        f2 = __uint_as_float(v[u][6 + c]);
            acc[0][c] = fmaf(w2.x, f2, fmaf(w1.x, f1, fmaf(w0.x, f0, acc[0][c])));
            acc[1][c] = fmaf(w2.y, f2, fmaf(w1.y, f1, fmaf(w0.y, f0, acc[1][c])));
            acc[2][c] = fmaf(w2.z, f2, fmaf(w1.z, f1, fmaf(w0.z, f0, acc[2][c])));
            acc[3][c] = fmaf(w2.w, f2, fmaf(w1.w, f1, fmaf(w0.w, f0, acc[3][c])));
          }
          // slot u is complete: one NHWC4 pixel (zero for an item that emits nothing: a pad column), slot restarts
          const uint32_t ox = pack_bf16x2(fmaf(acc[u][0], sc0, bi0), fmaf(acc[u][1], sc1, bi1));
          const uint32_t oy = pack_bf16x2(fmaf(acc[u][2], sc2, bi2), 0.f);
          st[u][0] = j >= 0 ? ox : 0u;
          st[u][1] = j >= 0 ? oy : 0u;
          acc[u][0] = acc[u][1] = acc[u][2] = 0.f;
        }
#ifdef SIA_TC_NOSTORE
        const bool store_ok = row_ok && vec_col == 123456;
#else
        const bool store_ok = row_ok;
#endif
        if (vec_col >= 0) {                          // uniform: 2 x 16-byte stores, 32-byte aligned
          if (store_ok) {
            st_global_256(orow + vec_col, make_uint4(st[0][0], st[0][1], st[1][0], st[1][1]),
                          make_uint4(st[2][0], st[2][1], st[3][0], st[3][1]));
          }
        } else {
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            if (store_ok && js[u] >= 0) orow[js[u] + 1] = make_uint2(st[u][0], st[u][1]);
          }
        }
      };

#ifdef SIA_TC_NOEPI
      for (int bb = 0; bb < p.n_blocks; ++bb) acquire(bb);
      if (false)
#endif
      load_group(0, va);
#ifdef SIA_TC_NOEPI
      if (false)
#endif
      for (int n0 = 0; n0 < p.n_items; n0 += 8) {
        tmem_ld_wait();                              // group n0 has landed
        load_group(n0 + 4, vb);
        compute_group(n0, va);
        tmem_ld_wait();                              // group n0 + 4 has landed
        if (n0 + 8 < p.n_items) load_group(n0 + 8, va);
        compute_group(n0 + 4, vb);
      }
      // hand the last accumulator(s) of this image back
      acquire(p.n_blocks - 1);
      tmem_ld_wait();
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[2 * g + ((acc_cnt - 1) & 1u)]);
    }
  }

#ifdef SIA_TC_EPI_ONLY
tc_done:
#endif
  tc_fence_before_sync();
  __syncthreads();
  if (warp == TC_WARP_MMA) tmem_free(tmem_base, 512);
}

}  // namespace sia

extern "C" int sia_preprocess_tc_u8hwc(const uint8_t* src, int batch, int src_h, int src_w, const void* a_packed,
                                       const float* lane_scale, const int32_t* tile_row0, int n_tiles, int tile_rows,
                                       const void* items, int n_items, int n_blocks, int last_block_cols,
                                       int pads_in_schedule, int out_h, int out_w, const float* out_scale_host, const float* out_bias_host,
                                       void* dst_nhwc4, void* stream) {
  using namespace sia;
  SIA_REQUIRE(src && a_packed && lane_scale && tile_row0 && items && dst_nhwc4 && out_scale_host && out_bias_host);
  SIA_REQUIRE(batch >= 1 && src_h >= 2 && src_w >= 8 && out_h >= 1 && out_w >= 1 && n_tiles >= 1);
  SIA_REQUIRE(n_items >= 8 && n_items % 8 == 0);
  SIA_REQUIRE(tile_rows >= 1 && tile_rows <= 128 && n_tiles * tile_rows >= out_h && n_blocks >= 1);
  SIA_REQUIRE(last_block_cols >= 16 && last_block_cols <= TC_COLS && last_block_cols % 16 == 0);
  SIA_REQUIRE(aligned(a_packed, 16) && aligned(items, 16) && aligned(dst_nhwc4, 32));   // 256-bit stores
  const int row_bytes = src_w * 3;
  // 8-byte pieces; rows paired for the TMA view; every image 16-byte aligned
  if (row_bytes % 8 != 0 || src_h % 2 != 0 || !aligned(src, 16) || ((uint64_t)src_h * row_bytes) % 16 != 0)
    return SIA_E_UNSUPPORTED;
  if ((n_blocks - 1) * TC_STRIDE >= row_bytes || (n_blocks + 1) * TC_STRIDE < row_bytes) return SIA_E_INVALID;
  if (n_tiles > sm_count()) return SIA_E_UNSUPPORTED;
  if (int wrc = ensure_watchdog()) return wrc;

  TcParams p;
  p.a_packed = static_cast<const uint8_t*>(a_packed);
  p.lane_scale = lane_scale;
  p.tile_row0 = tile_row0;
  p.items = static_cast<const uint4*>(items);
  p.dst = static_cast<uint2*>(dst_nhwc4);
  p.batch = batch; p.src_h = src_h; p.src_w = src_w; p.out_h = out_h; p.out_w = out_w;
  p.n_tiles = n_tiles; p.tile_rows = tile_rows; p.n_blocks = n_blocks; p.last_block_cols = last_block_cols;
  p.n_items = n_items;
  p.pads_in_schedule = pads_in_schedule ? 1 : 0;
  p.odd_shift = row_bytes % 16;
  for (int c = 0; c < 3; ++c) { p.scale[c] = out_scale_host[c]; p.bias[c] = out_bias_host[c]; }

  // the decode buffers as [batch][src_h / 2] double rows of row_bytes u16 elements (a pitch TMA accepts)
  CUtensorMap tmap;
  const uint64_t dims[3] = {(uint64_t)row_bytes, (uint64_t)src_h / 2, (uint64_t)batch};
  const uint64_t strides[2] = {(uint64_t)2 * row_bytes, (uint64_t)src_h * row_bytes};
  const uint32_t box[3] = {TC_RAW_ROWB / 2, TC_KSTAGE / 2, 1};
  int trc = encode_tmap(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT16, src, 3, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE);
  if (trc != 0) return trc;
  int smem = 0;
  for (p.n_raw = TC_NRAW; p.n_raw >= 2; --p.n_raw) {
    smem = 1024 + TC_A_BYTES + p.n_raw * TC_RAW_BYTES + TC_NSTAGE * TC_STAGE_BYTES + n_items * TC_ITEM_BYTES +
           (2 * TC_NSTAGE + 2 * TC_NRAW + 10) * 8;
    if (smem <= 227 * 1024) break;
  }
  if (p.n_raw < 2) return SIA_E_UNSUPPORTED;
  static SmemSlots configured = {};
  if (int rc2 = ensure_dynamic_smem(preprocess_tc_kernel, smem, &configured)) return rc2;
  int grid = (sm_count() / n_tiles) * n_tiles;
  if (grid > batch * n_tiles) grid = batch * n_tiles;
  preprocess_tc_kernel<<<grid, TC_THREADS, smem, static_cast<cudaStream_t>(stream)>>>(tmap, p);
  return launch_status();
}
