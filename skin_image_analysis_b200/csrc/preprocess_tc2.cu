// K1-K3, two tensor-core products (impl="tensor_core2"): the kernel of preprocess_tc.cu with the HORIZONTAL pass
// moved onto the tensor cores as well.
//
//   vertical    V[i, k]        = sum_r Wy[i, r] * S[r, k]            exactly as in preprocess_tc.cu (TMA -> byte -> fp16
//                                                                   converters -> tcgen05.mma, fp32 accumulator in TMEM)
//   horizontal  out[i, 3j + c] = sum_k V[i, k] * Wx_b[3j + c, k]     per 120-byte column block b: a SECOND tcgen05.mma whose
//               A operand is V itself, read from tensor memory: the warp group that owns the accumulator loads its 128
//               fp32 columns, applies the lane scale, packs them to fp16 pairs and stores them back over the first 64
//               columns of the same accumulator (tcgen05.st; layout pinned by tests/test_umma_probe.py::
//               test_exploratory_a_operand_in_tensor_memory); B = the block's 64 x 128 fp16 slice of Wx, streamed by a
//               bulk copy into a 2-deep ring; D2 = 64 fp32 columns written over the second half of the accumulator.
//               Output pixels whose taps straddle two blocks are finished by a 12-register carry (slot layout of
//               resize_weights.build_tc2_tables: 4 "in" + 13 "full" + 4 "out" pixel slots x 3 channels per block).
//
// Against the one-product kernel this replaces ~1 000 FFMA / TMEM-load / table-lookup instructions per block and
// thread by ~250 and takes the item table out of shared memory; the price is V rounded to fp16 (<= 2.5e-4 of full
// scale in total, <= 1 bf16 ulp, ~1.8 % of the bf16 outputs round the other way compared with the fp64 oracle).
//
// Two MMA-issuing warps walk the same sequence of items (image pair, block, image of the pair): warp 12 issues the
// vertical products, warp 14 the horizontal ones as soon as the fp16 V of an item is back in tensor memory.  Each
// warp group alternates: outputs of block b - 1 (as soon as its second product completes), then the repack of block b.
#include <cuda_fp16.h>

#include <type_traits>

#include "sia_host.cuh"
#include "sia_ptx.cuh"

namespace sia {

constexpr int T2_N = 64;                          // columns of the second product (21 pixel slots x 3 + 1)
constexpr int T2_SLOTS_IN = 4, T2_SLOTS_FULL = 13, T2_SLOTS_OUT = 4;
constexpr int T2_SLOTS = T2_SLOTS_IN + T2_SLOTS_FULL + T2_SLOTS_OUT;
constexpr int T2_EMIT = T2_SLOTS_IN + T2_SLOTS_FULL;             // slots that can complete an output pixel
constexpr int T2_B2_BYTES = T2_N * TC_COLS * 2;   // 16384: one block's slice of Wx (fp16, K-major core matrices)
constexpr int T2_NB2 = 2;                         // ring of Wx slices
constexpr int T2_THREADS = TC_THREADS + 32;       // one more warp than the one-product kernel:
constexpr int T2_WARP_MMA2 = 14;                  // the issuer of the horizontal products
constexpr int T2_B2_LBO = 128, T2_B2_SBO = (TC_COLS / 8) * 128;

struct Tc2Params {
  TcParams t;                   // vertical side: identical to the one-product kernel (items unused)
  const uint8_t* b2;            // [n_blocks][T2_B2_BYTES]
  const int32_t* block_meta;    // [n_blocks][4]: s_lo, s_hi (slots the block completes), j_lo (their first column), 0
  const float* slot_scale;      // [n_blocks][T2_SLOTS]
};

// kind::f16, fp16 x fp16 -> fp32, A from tensor memory, B K-major
__host__ __device__ constexpr uint32_t tc2_idesc() {
  return (1u << 4) | (((uint32_t)T2_N >> 3) << 17) | ((128u >> 4) << 24);
}

__device__ __forceinline__ void umma_f16_ts_w(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_lo, uint32_t b_hi,
                                              uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 db;\n\t"
      "mov.b64 db, {%2, %3};\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
        "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]),
        "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]),
        "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t pack_f16x2(float lo, float hi) {
  const __half2 h = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&h);
}

__global__ void __launch_bounds__(T2_THREADS, 1)
preprocess_tc2_kernel(const __grid_constant__ CUtensorMap tmap_src, const __grid_constant__ Tc2Params pp) {
  const TcParams& p = pp.t;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* smem_a = smem;                                        // TC_A_BYTES
  uint8_t* smem_raw_ring = smem + TC_A_BYTES;                    // n_raw * TC_RAW_BYTES
  uint8_t* smem_b = smem_raw_ring + p.n_raw * TC_RAW_BYTES;      // TC_NSTAGE * TC_STAGE_BYTES
  uint8_t* smem_b2 = smem_b + TC_NSTAGE * TC_STAGE_BYTES;        // T2_NB2 * T2_B2_BYTES
  int4* block_meta_s = reinterpret_cast<int4*>(smem_b2 + T2_NB2 * T2_B2_BYTES);        // [n_blocks]
  float* slot_scale_s = reinterpret_cast<float*>(block_meta_s + p.n_blocks);             // [n_blocks][T2_SLOTS]
  uint64_t* bars = reinterpret_cast<uint64_t*>(
      (reinterpret_cast<uintptr_t>(slot_scale_s + p.n_blocks * T2_SLOTS) + 7) & ~uintptr_t(7));
  uint64_t* full_bar = bars;                           // [TC_NSTAGE]
  uint64_t* empty_bar = bars + TC_NSTAGE;              // [TC_NSTAGE]
  uint64_t* tfull_bar = bars + 2 * TC_NSTAGE;          // [4] vertical product of accumulator 2g + b complete
  uint64_t* tempty_bar = tfull_bar + 4;                // [4] accumulator drained
  uint64_t* vready_bar = tfull_bar + 8;                // [4] fp16 V stored back
  uint64_t* d2full_bar = tfull_bar + 12;               // [4] horizontal product complete
  uint64_t* a_bar = tfull_bar + 16;
  uint64_t* raw_full = tfull_bar + 17;                 // [TC_NRAW]
  uint64_t* raw_empty = raw_full + TC_NRAW;            // [TC_NRAW]
  uint64_t* b2_full = raw_empty + TC_NRAW;             // [T2_NB2]
  uint64_t* b2_empty = b2_full + T2_NB2;               // [T2_NB2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(b2_empty + T2_NB2);

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
      mbar_init(&vready_bar[i], 4);
      mbar_init(&d2full_bar[i], 1);
    }
    for (int i = 0; i < TC_NRAW; ++i) {
      mbar_init(&raw_full[i], 1);      // TMA transaction bytes
      mbar_init(&raw_empty[i], 4);     // one arrival per converter warp
    }
    for (int i = 0; i < T2_NB2; ++i) {
      mbar_init(&b2_full[i], 1);
      mbar_init(&b2_empty[i], 1);
    }
    mbar_init(a_bar, 1);
    fence_mbar_init();
    tma_prefetch_desc(&tmap_src);
  }
  if (warp == TC_WARP_MMA) tmem_alloc(tmem_slot, 512);
  for (int i = threadIdx.x; i < p.n_blocks * T2_SLOTS; i += blockDim.x) slot_scale_s[i] = pp.slot_scale[i];
  for (int i = threadIdx.x; i < p.n_blocks; i += blockDim.x) block_meta_s[i] = reinterpret_cast<const int4*>(pp.block_meta)[i];
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = bcast0(*tmem_slot);

  if (warp == TC_WARP_TMA) {
    // ================================ TMA producer ==========================================
    if (lane == 0) {
      mbar_arrive_expect_tx(a_bar, TC_A_BYTES);
      for (int off = 0; off < TC_A_BYTES; off += 16384)
        bulk_load_1d(smem_a + off, p.a_packed + (size_t)tile * TC_A_BYTES + off, 16384, a_bar);
      int rs = 0;
      uint32_t rphase = 0;
      int b2_loads = 0;                                  // Wx slices requested so far: slot = count % T2_NB2
      for (int k0 = 0; k0 < n_img; k0 += 2) {
        for (int blk = 0; blk < p.n_blocks; ++blk) {
          for (int g = 0; g < 2 && k0 + g < n_img; ++g) {
            const int img = img0 + (k0 + g) * img_step;
            // TMA needs 16-byte aligned starts: a row whose window starts 8 bytes off is fetched 8 bytes early
            const int col = blk * TC_STRIDE;
            const int mis_even = col & 15, mis_odd = (col + p.odd_shift) & 15;
            for (int q = 0; q < TC_NQ; ++q) {
              const int r0 = row0 + q * TC_KSTAGE;           // first source row of the stage
              const int even0 = (r0 + 1) >> 1;               // double row of the first even / odd row
              const int odd0 = r0 >> 1;
              mbar_wait(&raw_empty[rs], rphase ^ 1, 45);
              mbar_arrive_expect_tx(&raw_full[rs], TC_RAW_BYTES);
              uint8_t* dst = smem_raw_ring + rs * TC_RAW_BYTES;
              // innermost coordinate in 16-bit elements (the tensor map views the bytes as u16 pairs)
              tma_load_3d(dst, &tmap_src, &raw_full[rs], (col - mis_even) >> 1, even0, img);
              tma_load_3d(dst + TC_RAW_HALF, &tmap_src, &raw_full[rs], (row_bytes + col - mis_odd) >> 1, odd0, img);
              if (++rs == p.n_raw) { rs = 0; rphase ^= 1; }
            }
          }
          // the block's slice of Wx, AFTER the raw stages of the block (it is first needed two sequence items later,
          // and its ring slot is released while the previous block's items run: the raw prefetch never waits on it)
          {
            const int slot = b2_loads % T2_NB2;
            mbar_wait(&b2_empty[slot], (((uint32_t)(b2_loads / T2_NB2)) & 1u) ^ 1u, 47);
            mbar_arrive_expect_tx(&b2_full[slot], T2_B2_BYTES);
            bulk_load_1d(smem_b2 + slot * T2_B2_BYTES, pp.b2 + (size_t)blk * T2_B2_BYTES, T2_B2_BYTES, &b2_full[slot]);
            ++b2_loads;
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
            }
            if (++stage == TC_NSTAGE) { stage = 0; phase ^= 1; }
            if (++rs == p.n_raw) { rs = 0; rphase ^= 1; }
          }
        }
      }
    }
  } else if (warp == TC_WARP_MMA) {
    // ================================ MMA issuer: vertical products ==========================
    constexpr uint32_t a_hi = desc_hi(TC_A_SBO, SW_NONE);
    constexpr uint32_t b_hi = desc_hi(TC_B_SBO, SW_NONE);
    const uint32_t a_lo0 = desc_lo(smem_u32(smem_a), TC_A_LBO);
    const uint32_t b_lo0 = desc_lo(smem_u32(smem_b), TC_B_LBO);
    mbar_wait(a_bar, 0, 41);
    int stage = 0;
    uint32_t phase = 0;
    uint32_t acc_cnt[2] = {0u, 0u};                 // blocks issued so far for warp group g: buffer = 2g + (cnt & 1)
    for (int k0 = 0; k0 < n_img; k0 += 2) {
      for (int blk = 0; blk < p.n_blocks; ++blk) {
        const uint32_t idesc = tc_idesc((uint32_t)(blk == p.n_blocks - 1 ? p.last_block_cols : TC_COLS));
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          if (k0 + g < n_img) {
            const int buf = 2 * g + (int)(acc_cnt[g] & 1u);
            mbar_wait(&tempty_bar[buf], ((acc_cnt[g] >> 1) & 1u) ^ 1u, 42);
            ++acc_cnt[g];
            tc_fence_after_sync();
            const uint32_t d_tmem = tmem_base + buf * TC_COLS;
            for (int q = 0; q < TC_NQ; ++q) {
              mbar_wait(&full_bar[stage], phase, 43);
              tc_fence_after_sync();
              if (elect_one()) {
                const uint32_t b_stage = b_lo0 + stage * (TC_STAGE_BYTES >> 4);
#pragma unroll
                for (int kk = 0; kk < TC_KSTAGE / 16; ++kk) {
                  umma_bf16_ss_w(d_tmem, a_lo0 + (q * (TC_KSTAGE / 16) + kk) * ((2 * TC_A_LBO) >> 4), a_hi,
                                 b_stage + kk * ((2 * TC_B_LBO) >> 4), b_hi, idesc, (q | kk) ? 1u : 0u);
                }
                umma_commit(&empty_bar[stage]);
                if (q == TC_NQ - 1) umma_commit(&tfull_bar[buf]);
              }
              __syncwarp();
              if (++stage == TC_NSTAGE) { stage = 0; phase ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == T2_WARP_MMA2) {
    // ================================ MMA issuer: horizontal products ========================
    // A second issuing warp (an mbarrier wait costs the issuing thread ~130 clocks; see conv1.cu): it walks the same
    // sequence, purely dependency driven -- fp16 V stored back (vready) and the block's Wx slice landed (b2_full).
    constexpr uint32_t b2_hi = desc_hi(T2_B2_SBO, SW_NONE);
    constexpr uint32_t idesc2 = tc2_idesc();
    const uint32_t b2_lo0 = desc_lo(smem_u32(smem_b2), T2_B2_LBO);
    uint32_t acc_cnt[2] = {0u, 0u};
    for (int k0 = 0; k0 < n_img; k0 += 2) {
      for (int blk = 0; blk < p.n_blocks; ++blk) {
        const int load = (k0 >> 1) * p.n_blocks + blk;
        const int slot = load % T2_NB2;
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          if (k0 + g < n_img) {
            const int buf = 2 * g + (int)(acc_cnt[g] & 1u);
            mbar_wait(&vready_bar[buf], (acc_cnt[g] >> 1) & 1u, 48);
            ++acc_cnt[g];
            mbar_wait(&b2_full[slot], ((uint32_t)(load / T2_NB2)) & 1u, 49);
            tc_fence_after_sync();
            if (elect_one()) {
              const uint32_t acc = tmem_base + buf * TC_COLS;
              const uint32_t b2_lo = b2_lo0 + slot * (T2_B2_BYTES >> 4);
#pragma unroll
              for (int kk = 0; kk < TC_COLS / 16; ++kk) {
                // K = 16 bytes of the block per instruction: 8 packed-pair columns of A, two core matrices of B
                umma_f16_ts_w(acc + 64, acc + 8 * kk, b2_lo + kk * ((2 * T2_B2_LBO) >> 4), b2_hi, idesc2, kk ? 1u : 0u);
              }
              umma_commit(&d2full_bar[buf]);
              if (g == 1 || k0 + 1 >= n_img) umma_commit(&b2_empty[slot]);      // last use of the slice
            }
            __syncwarp();
          }
        }
      }
    }
  } else if (warp < 12) {
    // ================================ repack + output =======================================
    const int g = (warp - 4) >> 2;                 // warp group = image parity in the CTA's sequence
    const int e = warp & 3;                        // TMEM lanes 32e .. 32e+31
    const int l = 32 * e + lane;                   // output row inside the tile
    const int i = tile * p.tile_rows + l;
    const bool row_ok = l < p.tile_rows && i < p.out_h;
    const float ls = p.lane_scale[tile * 128 + l];
    const float sc0 = p.scale[0], sc1 = p.scale[1], sc2 = p.scale[2];
    const float bi0 = p.bias[0], bi1 = p.bias[1], bi2 = p.bias[2];
    const int pitch = p.out_w + SIA_NHWC4_PAD;
    const uint32_t t_lanes = tmem_base + ((uint32_t)(32 * e) << 16) + 2 * g * TC_COLS;
    uint32_t acc_cnt = 0;                          // blocks of this group so far: accumulator 2g + (cnt & 1)
    float carry[3 * T2_SLOTS_OUT];
#pragma unroll
    for (int q = 0; q < 3 * T2_SLOTS_OUT; ++q) carry[q] = 0.f;
    uint2 dfr[4];                                  // completed pixels of an unfinished 4-column group (see output())
#pragma unroll
    for (int u = 0; u < 4; ++u) dfr[u] = make_uint2(0u, 0u);
    int prev_blk = -1, prev_img = 0;               // block whose outputs are still to be written
    uint32_t prev_cnt = 0;

    auto output = [&]() {
      const uint32_t b = prev_cnt & 1u;
      mbar_wait(&d2full_bar[2 * g + b], (prev_cnt >> 1) & 1u, 50);
      tc_fence_after_sync();
      uint32_t d[64];
      {
        uint32_t lo[32], hi[32];
        tmem_ld32(t_lanes + b * TC_COLS + 64, lo);
        tmem_ld32(t_lanes + b * TC_COLS + 96, hi);
        tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 32; ++q) { d[q] = lo[q]; d[32 + q] = hi[q]; }
      }
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[2 * g + b]);          // the accumulator is in registers
      if (prev_blk == 0) {                         // a new image row: no partial sums, no deferred pixels
#pragma unroll
        for (int q = 0; q < 3 * T2_SLOTS_OUT; ++q) carry[q] = 0.f;
#pragma unroll
        for (int u = 0; u < 4; ++u) dfr[u] = make_uint2(0u, 0u);
      }
#ifdef SIA_T2_SAMEADDR
      uint2* orow = p.dst + ((size_t)(prev_img & 1) * p.out_h + (row_ok ? i : 0)) * pitch;   // timing experiment
#else
      uint2* orow = p.dst + ((size_t)prev_img * p.out_h + (row_ok ? i : 0)) * pitch;
#endif
      // The block completes the pixel slots [s_lo, s_hi) = output columns j_lo, j_lo + 1, ...: the pixel of slot s
      // lives in padded column c0 + s.  Only whole groups of four padded columns starting at a multiple of four are
      // ever written -- one 32-byte sector per lane in one 256-bit store.  The pixels of a group the block leaves
      // unfinished are deferred in registers (dfr) and head the first group of the next block; a row starts with a
      // zero deferred pixel (pad column 0) and its last block fills its last group and the rest of the pad columns
      // with zeros, so no partial-sector write exists at all.
      const int4 meta = block_meta_s[prev_blk];
      const int s_lo = meta.x, s_hi = meta.y;
      const int c0 = meta.z - s_lo + 1;
      const bool last_blk = prev_blk == p.n_blocks - 1;
      const float* scl = slot_scale_s + prev_blk * T2_SLOTS;
      float new_carry[3 * T2_SLOTS_OUT];
#pragma unroll
      for (int q = 0; q < 3 * T2_SLOTS_OUT; ++q) new_carry[q] = __uint_as_float(d[3 * T2_EMIT + q]);
      uint2 px[T2_SLOTS];
#pragma unroll
      for (int s = 0; s < T2_SLOTS; ++s) {
        float v0 = __uint_as_float(d[3 * s]), v1 = __uint_as_float(d[3 * s + 1]), v2 = __uint_as_float(d[3 * s + 2]);
        if (s < T2_SLOTS_IN) { v0 += carry[3 * s]; v1 += carry[3 * s + 1]; v2 += carry[3 * s + 2]; }
        const float cs = scl[s];
        px[s] = make_uint2(pack_bf16x2(fmaf(v0, sc0 * cs, bi0), fmaf(v1, sc1 * cs, bi1)),
                           pack_bf16x2(fmaf(v2, sc2 * cs, bi2), 0.f));
      }
#ifdef SIA_T2_NOSTORE
      const bool st_ok = row_ok && s_lo == 123456;
#else
      const bool st_ok = row_ok;
#endif
      auto emit = [&](auto a_tag) {
        constexpr int A = decltype(a_tag)::value;          // c0 & 3
        constexpr int S0 = (4 - A) & 3;                     // first slot >= 0 whose column is a multiple of four
#pragma unroll
        for (int sg = S0 - 4; sg < T2_SLOTS; sg += 4) {     // groups of four columns (the first may begin before slot 0)
          if (sg + 4 <= s_lo || sg >= s_hi) continue;       // uniform: nothing of this block in the group
          uint2 v[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int sl = sg + u;                          // compile-time
            const uint2 own = (sl >= 0 && sl < T2_SLOTS) ? px[sl >= 0 && sl < T2_SLOTS ? sl : 0] : make_uint2(0u, 0u);
            v[u] = sl < s_lo ? dfr[u] : own;                // before s_lo: what the previous block deferred
          }
          if (sg + 4 <= s_hi || last_blk) {
#pragma unroll
            for (int u = 0; u < 4; ++u)
              if (sg + u >= s_hi) v[u] = make_uint2(0u, 0u);                          // row end: pad columns
            if (st_ok) st_global_256(orow + c0 + sg, make_uint4(v[0].x, v[0].y, v[1].x, v[1].y),
                                     make_uint4(v[2].x, v[2].y, v[3].x, v[3].y));
          } else {
#pragma unroll
            for (int u = 0; u < 4; ++u) dfr[u] = v[u];      // unfinished group: finished by the next block
          }
        }
      };
      switch (c0 & 3) {
        case 0: emit(std::integral_constant<int, 0>{}); break;
        case 1: emit(std::integral_constant<int, 1>{}); break;
        case 2: emit(std::integral_constant<int, 2>{}); break;
        default: emit(std::integral_constant<int, 3>{}); break;
      }
      if (last_blk && st_ok) {                              // remaining pad columns of the row
        for (int c = (c0 + s_hi + 3) & ~3; c < pitch; c += 4)
          st_global_256(orow + c, make_uint4(0u, 0u, 0u, 0u), make_uint4(0u, 0u, 0u, 0u));
      }
#pragma unroll
      for (int q = 0; q < 3 * T2_SLOTS_OUT; ++q) carry[q] = new_carry[q];
    };


    for (int k = g; k < n_img; k += 2) {
      const int img = img0 + k * img_step;
      for (int blk = 0; blk < p.n_blocks; ++blk) {
        // ---- repack block blk FIRST (its horizontal product then runs while the previous block's outputs are
        //      written), then the outputs of the previous block: fp32 V * lane scale -> fp16 pairs over the first 64 columns of the accumulator
        const uint32_t b = acc_cnt & 1u;
        mbar_wait(&tfull_bar[2 * g + b], (acc_cnt >> 1) & 1u, 44);
        tc_fence_after_sync();
        const uint32_t t_acc = t_lanes + b * TC_COLS;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint32_t lo[32], hi[32], pk[32];
          tmem_ld32(t_acc + 64 * h, lo);
          tmem_ld32(t_acc + 64 * h + 32, hi);
          tmem_ld_wait();
#pragma unroll
          for (int q = 0; q < 16; ++q) {
            pk[q] = pack_f16x2(__uint_as_float(lo[2 * q]) * ls, __uint_as_float(lo[2 * q + 1]) * ls);
            pk[16 + q] = pack_f16x2(__uint_as_float(hi[2 * q]) * ls, __uint_as_float(hi[2 * q + 1]) * ls);
          }
          tmem_st32(t_acc + 32 * h, pk);           // columns 0..63 have been read: overwriting 0..31 / 32..63 is safe
        }
        tmem_st_wait();
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&vready_bar[2 * g + b]);
        if (prev_blk >= 0) output();
        prev_blk = blk; prev_img = img; prev_cnt = acc_cnt;
        ++acc_cnt;
      }
    }
    if (prev_blk >= 0) output();
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == TC_WARP_MMA) tmem_free(tmem_base, 512);
}

}  // namespace sia

extern "C" int sia_preprocess_tc2_u8hwc(const uint8_t* src, int batch, int src_h, int src_w, const void* a_packed,
                                        const float* lane_scale, const int32_t* tile_row0, int n_tiles, int tile_rows,
                                        const void* b2, const int32_t* block_meta, const float* slot_scale, int n_blocks,
                                        int last_block_cols, int out_h, int out_w, const float* out_scale_host,
                                        const float* out_bias_host, void* dst_nhwc4, void* stream) {
  using namespace sia;
  SIA_REQUIRE(src && a_packed && lane_scale && tile_row0 && b2 && block_meta && slot_scale && dst_nhwc4 &&
              out_scale_host && out_bias_host);
  SIA_REQUIRE(batch >= 1 && src_h >= 2 && src_w >= 8 && out_h >= 1 && out_w >= 1 && n_tiles >= 1);
  SIA_REQUIRE(tile_rows >= 1 && tile_rows <= 128 && n_tiles * tile_rows >= out_h && n_blocks >= 1);
  SIA_REQUIRE(last_block_cols >= 16 && last_block_cols <= TC_COLS && last_block_cols % 16 == 0);
  SIA_REQUIRE(aligned(a_packed, 16) && aligned(b2, 16) && aligned(block_meta, 16) && aligned(dst_nhwc4, 32));
  const int row_bytes = src_w * 3;
  if (row_bytes % 8 != 0 || src_h % 2 != 0 || !aligned(src, 16) || ((uint64_t)src_h * row_bytes) % 16 != 0)
    return SIA_E_UNSUPPORTED;
  if ((n_blocks - 1) * TC_STRIDE >= row_bytes || (n_blocks + 1) * TC_STRIDE < row_bytes) return SIA_E_INVALID;
  if (n_tiles > sm_count()) return SIA_E_UNSUPPORTED;
  if (int wrc = ensure_watchdog()) return wrc;

  Tc2Params pp;
  TcParams& p = pp.t;
  p.a_packed = static_cast<const uint8_t*>(a_packed);
  p.lane_scale = lane_scale;
  p.tile_row0 = tile_row0;
  p.items = nullptr;
  p.dst = static_cast<uint2*>(dst_nhwc4);
  p.batch = batch; p.src_h = src_h; p.src_w = src_w; p.out_h = out_h; p.out_w = out_w;
  p.n_tiles = n_tiles; p.tile_rows = tile_rows; p.n_blocks = n_blocks; p.last_block_cols = last_block_cols;
  p.n_items = 0;
  p.pads_in_schedule = 0;
  p.odd_shift = row_bytes % 16;
  for (int c = 0; c < 3; ++c) { p.scale[c] = out_scale_host[c]; p.bias[c] = out_bias_host[c]; }
  pp.b2 = static_cast<const uint8_t*>(b2);
  pp.block_meta = block_meta;
  pp.slot_scale = slot_scale;

  CUtensorMap tmap;
  const uint64_t dims[3] = {(uint64_t)row_bytes, (uint64_t)src_h / 2, (uint64_t)batch};
  const uint64_t strides[2] = {(uint64_t)2 * row_bytes, (uint64_t)src_h * row_bytes};
  const uint32_t box[3] = {TC_RAW_ROWB / 2, TC_KSTAGE / 2, 1};
  int trc = encode_tmap(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT16, src, 3, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE);
  if (trc != 0) return trc;
  int smem = 0;
  for (p.n_raw = TC_NRAW; p.n_raw >= 2; --p.n_raw) {
    smem = 1024 + TC_A_BYTES + p.n_raw * TC_RAW_BYTES + TC_NSTAGE * TC_STAGE_BYTES + T2_NB2 * T2_B2_BYTES +
           n_blocks * (T2_SLOTS * 4 + 16) + 8 + (2 * TC_NSTAGE + 2 * TC_NRAW + 2 * T2_NB2 + 18) * 8;
    if (smem <= 227 * 1024) break;
  }
  if (p.n_raw < 2) return SIA_E_UNSUPPORTED;
  static SmemSlots configured = {};
  if (int rc2 = ensure_dynamic_smem(preprocess_tc2_kernel, smem, &configured)) return rc2;
  int grid = (sm_count() / n_tiles) * n_tiles;
  if (grid > batch * n_tiles) grid = batch * n_tiles;
  preprocess_tc2_kernel<<<grid, T2_THREADS, smem, static_cast<cudaStream_t>(stream)>>>(tmap, pp);
  return launch_status();
}
