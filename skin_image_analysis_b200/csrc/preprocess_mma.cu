// Fused resize + scale/normalise + layout kernel, warp-level tensor-core pipeline (the DEFAULT for the NHWC4 layout).
//
// Replaces reference tone_bias_dataset.py:335 (u8 / 255), :411-427 (Rescale = skimage.transform.resize: Gaussian
// anti-alias, bilinear, mirror) and :464-473 (ToTensor) for a whole batch; writes the padded NHWC4 bf16 layout the
// first conv kernel consumes.  Judged on HBM GB/s: per image 810 000 B in + 415 744 B out, nothing else touches DRAM.
//
// Why this shape.  The resize is two banded products,  V = Wy . S  (over source rows) and  out = V . Wx^T  (over
// source pixels, per channel).  The CUDA-core kernel (preprocess.cu) spends ~150 instructions per (source row, output
// column) on byte unpacking and is issue-bound at 25 % of HBM; the tcgen05 kernels (preprocess_tc*.cu) must stage the
// image bytes as fp16 UMMA operands in shared memory and are shared-memory-bandwidth bound at 36-43 %.  Here both
// products run on mma.sync.m16n8k16 (fp16 in, fp32 accumulate) with the operands built IN REGISTERS:
//
//   * first product:  A = Wy fragment (16 output rows x 16 source rows, a host-built table),
//                     B = the raw image bytes: lane (g, q) loads 12 bytes of source rows 2q, 2q+1, 2q+8, 2q+9 with
//                         three 32-bit shared-memory loads per row and turns byte pairs of vertically adjacent rows
//                         into fp16x2 registers with PRMT -- the byte b becomes the fp16 SUBNORMAL b * 2^-24 (bit
//                         pattern 0x00bb), so no arithmetic is needed; the vertical weights carry 2^15;
//   * second product: the fp32 accumulators of two V n-tiles ARE the A fragment of the next mma (C layout of
//                     m16n8 == half the A layout of m16k16) after one cvt.rn.f16x2 each -- V never leaves the register
//                     file; B = Wx fragment (host-built table).  V n-tile b holds byte b of every lane group's 12
//                     bytes, i.e. pixels 4g + b/3 of channel b%3: choosing the K order of the second product
//                     accordingly de-interleaves the RGB bytes for free.
//   * each warp sweeps 32-pixel column groups of the row left to right.  An output n-tile (8 padded columns x 3
//     channels) reads at most two adjacent groups; it is computed when its last group is done, from that group's V
//     fragments and the previous group's (kept in registers), scaled, rounded to bf16 and stored as whole 32-byte
//     sectors (a lane quad writes 64 contiguous bytes).  Pad columns of the NHWC4 row have zero weights: the kernel
//     writes them as zeros itself.
//   * the compute warps of a CTA split the groups of a row.  A tile that straddles two warps' ranges is finished by the
//     LEFT warp: the right warp publishes the V fragments of its first group in shared memory as soon as it has them
//     (early in its sweep), the left warp picks them up at the end of its own sweep (mbarrier handshake, one 3 KB
//     buffer per boundary) -- no group is computed twice and nobody waits in steady state.
//   * one more warp streams the source rows into a ring of 8-row octets, ONE cp.async.bulk per octet (14 400 bytes at
//     the bench shape: the copy engine retires a request every ~190 clocks whatever its size, so per-row copies cap
//     the kernel at 2.7 TB/s).  Rows sit back to back; which chunk row a lane reads for which K slot is a per-geometry
//     permutation (resize_weights.mma_row_map) that makes the fragment loads bank-conflict free.
//   * work unit = 16 output rows of one image ("m-step"); the batch's m-steps are dealt to the CTAs (one per SM) in
//     contiguous, balanced ranges, so neighbouring m-steps reuse the source rows already in the ring.
//
// Arithmetic: fp16 weights (each rounded to nearest: 11 significant bits, also for the small taps that matter on
// sparse images), fp32 accumulation, V rounded once to fp16.  <= 1 bf16 ulp from the oracle
// (tests/test_gpu_preprocess_mma.py); numpy model: resize_weights.mma_emulate.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "sia_host.cuh"
#include "sia_ptx.cuh"

namespace sia {

constexpr int PM_GROUP_BYTES = 96;               // 32 pixels
#ifdef SIA_PM_NO_V                               // timing variant: no first product (the fragments are zeros)
#define PM_KV_RUN(kv) 0
#else
#define PM_KV_RUN(kv) (kv)
#endif
constexpr int PM_PUB_WORDS = 24 * 32;            // one published set of A fragments: 24 registers x 32 lanes
constexpr int PM_WIN_BARS = 8;                   // "the new octets of unit k have landed": one mbarrier per unit, ring of 8

struct PreMmaParams {
  const uint8_t* src;
  uint8_t* dst;
  const uint4* wy_frag;       // [n_msteps][KV][32]
  const int* r0;              // [n_msteps]
  const uint2* wx_frag;       // [n_tiles][2 (prev, cur)][2 (X)][32]
  const uint32_t* wx_mask;    // [n_tiles]
  const int* tile_begin;      // [n_groups + 1]
  float mul[3];               // 2^9 * scale / std_c
  float bias[3];              // -mean_c / std_c
  int has_bias;
  int batch, src_h, out_h, out_w;
  int n_msteps, n_groups, n_tiles;
  int row_bytes;              // 3 * src_w
  int q_stride;               // chunk row of K slot c for quad index q: q_stride * q + c_row[c]
  int c_row[4];
  int ring_octets;
};

__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// pack_f16x2 / pack_bf16x2 (lo, hi) -> one 32-bit register: preprocess_tc2.cu / sia_ptx.cuh

// mbarrier wait that backs off between polls (the copy lane shares its scheduler with two compute warps)
__device__ __forceinline__ void mbar_wait_backoff(uint64_t* bar, uint32_t parity, uint32_t site) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    __nanosleep(64);
    if (++spins > (SIA_WATCHDOG_SPINS >> 2)) {
      if (g_watchdog_word != nullptr) {
        *g_watchdog_word = 0x80000000u | (site << 16) | (blockIdx.x & 0xffffu);
        __threadfence_system();
      }
      __trap();
    }
  }
}

// One output n-tile: second product from the previous / current group's A fragments, scale, bf16, two 16-byte stores.
struct PmOut {
  uint8_t* row0;
  uint8_t* row1;
  bool ok0, ok1;
  float mul0, mul1, mul2, bias0, bias1, bias2;
  int has_bias, out_w, q;
};
__device__ __forceinline__ void pm_tile(const PmOut& o, int t, uint32_t mask, const uint2* wxt,
                                        const uint32_t (&prev)[3][2][4], const uint32_t (&cur)[3][2][4]) {
#ifdef SIA_PM_NO_TILES            // timing variant: no second product, no stores
  if (mask != 0xdeadbeefu) return;
#endif
  float hp[3][4], hc[3][4];                       // two independent accumulation chains per channel
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int e = 0; e < 4; ++e) hp[c][e] = hc[c][e] = 0.f;
#pragma unroll
  for (int x = 0; x < 2; ++x) {
    if ((mask >> x) & 1u) {
      const uint2 bf = wxt[x * 32];
#pragma unroll
      for (int c = 0; c < 3; ++c) mma16816(hp[c], prev[c][x], bf.x, bf.y);
    }
    if ((mask >> (2 + x)) & 1u) {
      const uint2 bf = wxt[(2 + x) * 32];
#pragma unroll
      for (int c = 0; c < 3; ++c) mma16816(hc[c], cur[c][x], bf.x, bf.y);
    }
  }
  const int pc0 = t * 8 + 2 * o.q;
  float b00 = 0.f, b01 = 0.f, b02 = 0.f, b10 = 0.f, b11 = 0.f, b12 = 0.f;
  if (o.has_bias) {                                     // pad columns stay zero whatever the bias
    if (pc0 >= 1 && pc0 <= o.out_w) { b00 = o.bias0; b01 = o.bias1; b02 = o.bias2; }
    if (pc0 + 1 <= o.out_w) { b10 = o.bias0; b11 = o.bias1; b12 = o.bias2; }
  }
  if (o.ok0) {
    uint4 v;
    v.x = pack_bf16x2(fmaf(hp[0][0] + hc[0][0], o.mul0, b00), fmaf(hp[1][0] + hc[1][0], o.mul1, b01));
    v.y = pack_bf16x2(fmaf(hp[2][0] + hc[2][0], o.mul2, b02), 0.f);
    v.z = pack_bf16x2(fmaf(hp[0][1] + hc[0][1], o.mul0, b10), fmaf(hp[1][1] + hc[1][1], o.mul1, b11));
    v.w = pack_bf16x2(fmaf(hp[2][1] + hc[2][1], o.mul2, b12), 0.f);
#ifndef SIA_PM_NO_STORE
    *reinterpret_cast<uint4*>(o.row0 + (size_t)pc0 * 8) = v;
#else
    if (v.x == 0x12345678u) *reinterpret_cast<uint4*>(o.row0 + (size_t)pc0 * 8) = v;
#endif
  }
  if (o.ok1) {
    uint4 v;
    v.x = pack_bf16x2(fmaf(hp[0][2] + hc[0][2], o.mul0, b00), fmaf(hp[1][2] + hc[1][2], o.mul1, b01));
    v.y = pack_bf16x2(fmaf(hp[2][2] + hc[2][2], o.mul2, b02), 0.f);
    v.z = pack_bf16x2(fmaf(hp[0][3] + hc[0][3], o.mul0, b10), fmaf(hp[1][3] + hc[1][3], o.mul1, b11));
    v.w = pack_bf16x2(fmaf(hp[2][3] + hc[2][3], o.mul2, b12), 0.f);
#ifndef SIA_PM_NO_STORE
    *reinterpret_cast<uint4*>(o.row1 + (size_t)pc0 * 8) = v;
#else
    if (v.x == 0x12345678u) *reinterpret_cast<uint4*>(o.row1 + (size_t)pc0 * 8) = v;
#endif
  }
}

template <int KV, int NW>
__global__ void __launch_bounds__((NW + 1) * 32, 1) preprocess_mma_kernel(const PreMmaParams p) {
  extern __shared__ __align__(128) uint8_t pm_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, q = lane & 3;
  const int R = p.ring_octets;
  const int octet_bytes = 8 * p.row_bytes;
  // ---- shared-memory carve-up -------------------------------------------------------------------------------------
  uint8_t* ring = pm_smem;
  size_t off = (size_t)R * octet_bytes;
  uint2* wx_s = reinterpret_cast<uint2*>(pm_smem + off);
  off += (size_t)p.n_tiles * 4 * 32 * sizeof(uint2);
  uint32_t* pub_s = reinterpret_cast<uint32_t*>(pm_smem + off);           // [NW][24][32]; slot 0 unused
  off += (size_t)NW * PM_PUB_WORDS * sizeof(uint32_t);
  int* r0_s = reinterpret_cast<int*>(pm_smem + off);
  off += (size_t)p.n_msteps * sizeof(int);
  int* tbeg_s = reinterpret_cast<int*>(pm_smem + off);
  off += (size_t)(p.n_groups + 1) * sizeof(int);
  uint32_t* mask_s = reinterpret_cast<uint32_t*>(pm_smem + off);
  off += (size_t)p.n_tiles * sizeof(uint32_t);
  off = (off + 7) & ~(size_t)7;
  uint64_t* win = reinterpret_cast<uint64_t*>(pm_smem + off);      // [PM_WIN_BARS]: all new octets of a unit have landed
  uint64_t* rel = win + PM_WIN_BARS;                                // [PM_WIN_BARS]: every compute warp is done with a unit
  uint64_t* tab_bar = rel + PM_WIN_BARS;                            // the B-fragment table has landed
  // fragment hand-over between neighbouring warps: plain shared-memory flags (unit numbers), polled with volatile loads
  // (a shared-memory load costs ~30 clocks, an mbarrier wait ~130 even when complete)
  volatile uint32_t* pub_flag = reinterpret_cast<volatile uint32_t*>(tab_bar + 1);    // [NW]: units published by warp w
  volatile uint32_t* pub_ack = pub_flag + NW;                                           // [NW]: units of warp w adopted

  pdl_launch_dependents();          // the next kernel (conv1) may start its prologue under this kernel's tail
  if (threadIdx.x == 0) {
    for (int i = 0; i < PM_WIN_BARS; ++i) {
      mbar_init(&win[i], 1);
      mbar_init(&rel[i], NW);
    }
    for (int i = 0; i < NW; ++i) pub_flag[i] = pub_ack[i] = 0u;
    mbar_init(tab_bar, 1);
    fence_mbar_init();
  }
  __syncthreads();
  pdl_wait();                       // (no-op unless launched as a programmatic dependent: the output buffer is shared)

  // ---- this CTA's contiguous range of m-steps ------------------------------------------------------------------------
  const long long total = (long long)p.batch * p.n_msteps;
  const int u_lo = (int)(total * blockIdx.x / gridDim.x);
  const int u_hi = (int)(total * (blockIdx.x + 1) / gridDim.x);
  const int n_oct_img = (p.src_h + 7) >> 3;
  const size_t image_bytes = (size_t)p.src_h * p.row_bytes;

  if (warp == NW) {
    // =============================== copy warp: one lane streams the table and the source octets =====================
    if (lane == 0) {
      const uint32_t tab_bytes = (uint32_t)(p.n_tiles * 4 * 32 * sizeof(uint2));
      mbar_arrive_expect_tx(tab_bar, tab_bytes);
      bulk_load_1d(wx_s, p.wx_frag, tab_bytes, tab_bar);
      int slot = 0, wb = 0;
      // release side: the compute warps arrive once per unit on rel[]; how many octets a unit retires follows from the
      // r0 table (everything below the next window's start; everything at the end of a pass), so the copy lane keeps a
      // second cursor over the units and a running count of retired octets
      int issued = 0, retired = 0, ru = u_lo, rb = 0, r_prev_end = 0;
      uint32_t rb_phase = 0;
      int u = u_lo;
      while (u < u_hi) {
        const int img = u / p.n_msteps;
        const int pass_end = min(u_hi, (img + 1) * p.n_msteps);
        const uint8_t* image = p.src + (size_t)img * image_bytes;
        int m = u - img * p.n_msteps;
        int o_next = __ldg(&p.r0[m]) >> 3;
        for (int uu = u; uu < pass_end; ++uu, ++m) {
          // one mbarrier per unit covers all the octets the unit adds to the ring (a wait on a completed mbarrier
          // still costs the waiting warp ~130 clocks: one wait per unit instead of one per octet)
          const int o_end = min((__ldg(&p.r0[m]) >> 3) + 2 * KV, n_oct_img);
          uint32_t total = 0;
          for (int o = o_next; o < o_end; ++o) total += (uint32_t)(min(8, p.src_h - o * 8) * p.row_bytes);
#ifdef SIA_PM_NO_COPY            // timing variant: no source traffic at all, the barriers still cycle
          total = 0;
#endif
          if (total > 0) mbar_arrive_expect_tx(&win[wb], total);
          else mbar_arrive(&win[wb]);
          for (int o = o_next; o < o_end; ++o) {
            while (issued - retired >= R) {          // the slot still holds an octet some warp may read
              mbar_wait_backoff(&rel[rb], rb_phase, 61);
              if (++rb == PM_WIN_BARS) { rb = 0; rb_phase ^= 1u; }
              const int img_r = ru / p.n_msteps, m_r = ru - img_r * p.n_msteps;
              const int pass_lo = max(u_lo, img_r * p.n_msteps), pass_hi = min(u_hi, (img_r + 1) * p.n_msteps);
              if (ru == pass_lo) r_prev_end = __ldg(&p.r0[m_r]) >> 3;
              const int waited = min((__ldg(&p.r0[m_r]) >> 3) + 2 * KV, n_oct_img);
              const int r_end = (ru + 1 < pass_hi) ? min(__ldg(&p.r0[m_r + 1]) >> 3, waited) : waited;
              retired += max(r_end - r_prev_end, 0);
              r_prev_end = max(r_prev_end, r_end);
              ++ru;
            }
            ++issued;
#ifndef SIA_PM_NO_COPY
            const uint32_t bytes = (uint32_t)(min(8, p.src_h - o * 8) * p.row_bytes);
            bulk_load_1d(ring + (size_t)slot * octet_bytes, image + (size_t)o * octet_bytes, bytes, &win[wb]);
#endif
            if (++slot == R) slot = 0;
          }
          if (++wb == PM_WIN_BARS) wb = 0;
          o_next = max(o_next, o_end);
        }
        u = pass_end;
      }
    }
    return;
  }

  // ================================== compute warps ================================================================
  {
    const int tid = threadIdx.x, nt = NW * 32;
    for (int i = tid; i < p.n_msteps; i += nt) r0_s[i] = __ldg(&p.r0[i]);
    for (int i = tid; i <= p.n_groups; i += nt) tbeg_s[i] = __ldg(&p.tile_begin[i]);
    for (int i = tid; i < p.n_tiles; i += nt) mask_s[i] = __ldg(&p.wx_mask[i]);
    asm volatile("bar.sync 1, %0;" ::"n"(NW * 32) : "memory");
    mbar_wait(tab_bar, 0, 60);
  }

  // Column groups of this warp for m-step number `step` of this CTA: [g_first, g_end) with boundaries
  // floor((n_groups * w + rot) / NW), rot rotating with the step -- 19 groups over 8 warps is 3,3,3,2,2,2,2,2 for any
  // fixed split, and the ring lets warps drift by an m-step, so rotating WHO gets the third group evens the load out.
  // (The host guarantees n_groups >= NW: every warp owns at least one group.)
  const bool publish = warp > 0;                   // my first group's fragments go to the left neighbour
  const bool adopt = warp + 1 < NW;                // I finish the tiles that straddle my right boundary
  // the two row pairs this lane reads in every 16-row chunk: chunk rows rowA, rowA + 1 (K slots 0, 1) and rowB, rowB + 1
  const int row_a = p.q_stride * q + p.c_row[0], row_b = p.q_stride * q + p.c_row[2];
  const int oct_a = row_a >> 3, oct_b = row_b >> 3;                              // 0 or 1: which octet of the chunk
  const uint32_t off_a = (uint32_t)((row_a & 7) * p.row_bytes + 12 * g);
  const uint32_t off_b = (uint32_t)((row_b & 7) * p.row_bytes + 12 * g);
  const uint32_t next_row = (uint32_t)p.row_bytes;
  const uint32_t ring_u32 = smem_u32(ring);
  const int out_pitch_px = p.out_w + SIA_NHWC4_PAD;
  uint32_t* my_pub = pub_s + (size_t)warp * PM_PUB_WORDS + lane;
  const uint32_t* right_pub = pub_s + (size_t)(warp + 1) * PM_PUB_WORDS + lane;

  PmOut out;
  out.mul0 = p.mul[0]; out.mul1 = p.mul[1]; out.mul2 = p.mul[2];
  out.bias0 = p.bias[0]; out.bias1 = p.bias[1]; out.bias2 = p.bias[2];
  out.has_bias = p.has_bias; out.out_w = p.out_w; out.q = q;

  uint4 wy_next[KV];                                // A fragments of the vertical operator, fetched one m-step ahead
  {
    const int m0 = u_lo % p.n_msteps;
#pragma unroll
    for (int kc = 0; kc < KV; ++kc) wy_next[kc] = __ldg(&p.wy_frag[(m0 * KV + kc) * 32 + lane]);
  }
  // ring bookkeeping without divisions: slot / parity of the next octet to wait for, to release, and of the window start
  int w_slot = 0, base_slot = 0, wb = 0;
  uint32_t wb_phase = 0;
  uint32_t step = 0;                                // m-steps done by this CTA: parity of the fragment hand-over
  int u = u_lo;
  while (u < u_hi) {
    const int img = u / p.n_msteps;
    const int pass_end = min(u_hi, (img + 1) * p.n_msteps);
    int m = u - img * p.n_msteps;
    const int o_start = r0_s[m] >> 3;
    int o_waited = o_start, o_base = o_start;      // octet numbers matching w_slot / base_slot
    for (int uu = u; uu < pass_end; ++uu, ++m, ++step) {
      const int rot = (int)((step * 5u) % (uint32_t)NW);          // 5 is coprime with 4, 8, 12, 16
      const int g_first = (p.n_groups * warp + rot) / NW, g_end = (p.n_groups * (warp + 1) + rot) / NW;
      const int o0 = r0_s[m] >> 3;
      const int o_win_end = min(o0 + 2 * KV, n_oct_img);
      mbar_wait(&win[wb], wb_phase, 62);            // every octet this unit adds to the ring has landed
      const int my_bar = wb;
      if (++wb == PM_WIN_BARS) { wb = 0; wb_phase ^= 1u; }
      if (o_win_end > o_waited) {
        w_slot += o_win_end - o_waited;             // (slot of the next octet nobody has waited for yet)
        if (w_slot >= R) w_slot -= R;
        o_waited = o_win_end;
      }
      base_slot += o0 - o_base;                     // (windows advance by less than a ring)
      if (base_slot >= R) base_slot -= R;
      o_base = o0;

#ifndef SIA_PM_NO_COMPUTE          // (timing variant without: the compute warps only wait for and release the octets)
      {
        const int row0 = m * 16 + g, row1 = row0 + 8;
        out.row0 = p.dst + ((size_t)img * p.out_h + row0) * out_pitch_px * 8;
        out.row1 = out.row0 + (size_t)8 * out_pitch_px * 8;
        out.ok0 = row0 < p.out_h;
        out.ok1 = row1 < p.out_h;

        // shared-memory addresses of this lane's two row pairs in each 16-row chunk of the window
        uint32_t pa[KV], pb[KV];
        uint4 wy[KV];                               // A fragments of the vertical operator for this m-step
        const int m_next = (m + 1 < p.n_msteps) ? m + 1 : 0;
#pragma unroll
        for (int kc = 0; kc < KV; ++kc) {
          int sa = base_slot + 2 * kc + oct_a, sb = base_slot + 2 * kc + oct_b;
          if (sa >= R) sa -= R;
          if (sb >= R) sb -= R;
          pa[kc] = ring_u32 + (uint32_t)(sa * octet_bytes) + off_a;
          pb[kc] = ring_u32 + (uint32_t)(sb * octet_bytes) + off_b;
          wy[kc] = wy_next[kc];
          wy_next[kc] = __ldg(&p.wy_frag[(m_next * KV + kc) * 32 + lane]);     // (L2 latency hidden behind this m-step)
        }

        uint32_t prev[3][2][4];                                    // A fragments of the previous group
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
          for (int x = 0; x < 2; ++x)
#pragma unroll
            for (int e = 0; e < 4; ++e) prev[c][x][e] = 0u;

        for (int grp = g_first; grp < g_end; ++grp) {
          // ---------------- first product: V[16 rows x 96 bytes] of this group --------------------------------------
          float vacc[12][4];
#pragma unroll
          for (int b = 0; b < 12; ++b)
#pragma unroll
            for (int e = 0; e < 4; ++e) vacc[b][e] = 0.f;
          const uint32_t goff = (uint32_t)(grp * PM_GROUP_BYTES);
#pragma unroll
          for (int kc = 0; kc < PM_KV_RUN(KV); ++kc) {
            const uint32_t qa = pa[kc] + goff, qb = pb[kc] + goff;
            uint32_t wa[3], wb[3], wc[3], wd[3];
#pragma unroll
            for (int w = 0; w < 3; ++w) {
              asm volatile("ld.shared.u32 %0, [%1];" : "=r"(wa[w]) : "r"(qa + 4 * w));
              asm volatile("ld.shared.u32 %0, [%1];" : "=r"(wb[w]) : "r"(qa + next_row + 4 * w));
              asm volatile("ld.shared.u32 %0, [%1];" : "=r"(wc[w]) : "r"(qb + 4 * w));
              asm volatile("ld.shared.u32 %0, [%1];" : "=r"(wd[w]) : "r"(qb + next_row + 4 * w));
            }
            const uint32_t af[4] = {wy[kc].x, wy[kc].y, wy[kc].z, wy[kc].w};
#pragma unroll
            for (int w = 0; w < 3; ++w) {
              const uint32_t t01 = __byte_perm(wa[w], wb[w], 0x5140), t23 = __byte_perm(wa[w], wb[w], 0x7362);
              const uint32_t u01 = __byte_perm(wc[w], wd[w], 0x5140), u23 = __byte_perm(wc[w], wd[w], 0x7362);
              mma16816(vacc[4 * w + 0], af, __byte_perm(t01, 0u, 0x5140), __byte_perm(u01, 0u, 0x5140));
              mma16816(vacc[4 * w + 1], af, __byte_perm(t01, 0u, 0x7362), __byte_perm(u01, 0u, 0x7362));
              mma16816(vacc[4 * w + 2], af, __byte_perm(t23, 0u, 0x5140), __byte_perm(u23, 0u, 0x5140));
              mma16816(vacc[4 * w + 3], af, __byte_perm(t23, 0u, 0x7362), __byte_perm(u23, 0u, 0x7362));
            }
          }
          // ---------------- V accumulators -> A fragments of the second product --------------------------------------
          uint32_t cur[3][2][4];
#pragma unroll
          for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int x = 0; x < 2; ++x) {
              const int ta = 6 * x + c, tb = 6 * x + 3 + c;
              cur[c][x][0] = pack_f16x2(vacc[ta][0], vacc[ta][1]);
              cur[c][x][1] = pack_f16x2(vacc[ta][2], vacc[ta][3]);
              cur[c][x][2] = pack_f16x2(vacc[tb][0], vacc[tb][1]);
              cur[c][x][3] = pack_f16x2(vacc[tb][2], vacc[tb][3]);
            }
          const bool first_group = grp == g_first;
          if (first_group && publish) {
            // hand this group's fragments to the left neighbour (it finishes the tiles that straddle the boundary)
            {                                         // the left neighbour has picked up my previous unit's fragments
              uint32_t spins = 0;
              while (pub_ack[warp] < step) {
                if (++spins > SIA_WATCHDOG_SPINS) {
                  if (g_watchdog_word != nullptr) *g_watchdog_word = 0x80000000u | (64u << 16) | (blockIdx.x & 0xffffu);
                  __trap();
                }
              }
            }
#pragma unroll
            for (int c = 0; c < 3; ++c)
#pragma unroll
              for (int x = 0; x < 2; ++x)
#pragma unroll
                for (int e = 0; e < 4; ++e) my_pub[((c * 2 + x) * 4 + e) * 32] = cur[c][x][e];
            __syncwarp();
            if (lane == 0) {
              __threadfence_block();
              pub_flag[warp] = step + 1u;
            }
          }
          // ---------------- second product + store for every output tile whose last group this is --------------------
          const int t_end = tbeg_s[grp + 1];
          for (int t = tbeg_s[grp]; t < t_end; ++t) {
            const uint32_t mask = mask_s[t];
            if (first_group && publish && (mask & 3u)) continue;          // straddles my left boundary: not mine
            pm_tile(out, t, mask, wx_s + (size_t)t * 4 * 32 + lane, prev, cur);
          }
#pragma unroll
          for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int x = 0; x < 2; ++x)
#pragma unroll
              for (int e = 0; e < 4; ++e) prev[c][x][e] = cur[c][x][e];
        }
        if (adopt) {
          // tiles whose last group is the right neighbour's first one and that also read my last group
          {
            uint32_t spins = 0;
            while (pub_flag[warp + 1] < step + 1u) {
              if (++spins > SIA_WATCHDOG_SPINS) {
                if (g_watchdog_word != nullptr) *g_watchdog_word = 0x80000000u | (65u << 16) | (blockIdx.x & 0xffffu);
                __trap();
              }
            }
            __threadfence_block();
          }
          uint32_t nxt[3][2][4];
#pragma unroll
          for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int x = 0; x < 2; ++x)
#pragma unroll
              for (int e = 0; e < 4; ++e) nxt[c][x][e] = right_pub[((c * 2 + x) * 4 + e) * 32];
          __syncwarp();
          if (lane == 0) {
            __threadfence_block();
            pub_ack[warp + 1] = step + 1u;
          }
          const int t_end = tbeg_s[g_end + 1];
          for (int t = tbeg_s[g_end]; t < t_end; ++t) {
            const uint32_t mask = mask_s[t];
            if (mask & 3u) pm_tile(out, t, mask, wx_s + (size_t)t * 4 * 32 + lane, prev, nxt);
          }
        }
      }
#endif

      // ---------------- done with the unit: the copy lane works out from the r0 table which octets that retires (those
      //                  below the next window's start; all of them at the end of a pass) ------------------------------
      __syncwarp();
      if (lane == 0) mbar_arrive(&rel[my_bar]);
    }
    // the next pass starts at the slot after the last octet of this one
    base_slot = w_slot;
    u = pass_end;
  }
}

template <int KV, int NW>
static int launch_pre_mma(const PreMmaParams& p, size_t smem_without_pub, cudaStream_t st) {
  auto kern = preprocess_mma_kernel<KV, NW>;
  static SmemSlots configured = {};
  const size_t smem = smem_without_pub + (size_t)NW * PM_PUB_WORDS * 4 + 8 + (size_t)2 * NW * 4 + 8;
  if (int rc = ensure_dynamic_smem(kern, (int)smem, &configured)) return rc;
  const long long total = (long long)p.batch * p.n_msteps;
  const int grid = (int)(total < sm_count() ? total : sm_count());
  return launch_kernel(kern, dim3(grid), dim3((NW + 1) * 32), smem, st, true, p);
}

static int g_pm_warps = 8;      // compute warps per CTA (sia_debug_set_mma_warps: 4 or 8; A/B timing only)

}  // namespace sia

extern "C" int sia_debug_set_mma_warps(int warps) {
  if (warps != 4 && warps != 8) return SIA_E_INVALID;
  sia::g_pm_warps = warps;
  return 0;
}

extern "C" int sia_preprocess_mma_u8hwc(const uint8_t* src, int batch, int src_h, int src_w, const void* wy_frag,
                                        const int32_t* r0, int n_msteps, int kv, const void* wx_frag,
                                        const uint32_t* wx_mask, const int32_t* tile_begin, int n_groups, int n_tiles,
                                        int q_stride, const int32_t* c_row4_host, const float* mul3_host,
                                        const float* bias3_host, int out_h, int out_w, void* dst, void* stream) {
  using namespace sia;
  SIA_REQUIRE(src && wy_frag && r0 && wx_frag && wx_mask && tile_begin && c_row4_host && mul3_host && bias3_host && dst);
  SIA_REQUIRE(batch >= 1 && src_h >= 2 && src_w >= 8 && out_h >= 1 && out_w >= 8);
  SIA_REQUIRE(aligned(src, 16) && aligned(dst, 16) && aligned(wy_frag, 16) && aligned(wx_frag, 16));
  if (src_w % 8 != 0 || src_h % 2 != 0 || out_w % 8 != 0) return SIA_E_UNSUPPORTED;
  if (n_msteps != (out_h + 15) / 16 || n_groups != (src_w + 31) / 32 || n_tiles != (out_w + SIA_NHWC4_PAD) / 8)
    return SIA_E_INVALID;
  const int warps = n_groups >= g_pm_warps ? g_pm_warps : 4;
  if (n_groups < warps) return SIA_E_UNSUPPORTED;            // every compute warp owns at least one column group
  if (int rc = ensure_watchdog()) return rc;
  PreMmaParams p;
  p.src = src;
  p.dst = static_cast<uint8_t*>(dst);
  p.wy_frag = static_cast<const uint4*>(wy_frag);
  p.r0 = r0;
  p.wx_frag = static_cast<const uint2*>(wx_frag);
  p.wx_mask = wx_mask;
  p.tile_begin = tile_begin;
  p.has_bias = 0;
  for (int c = 0; c < 3; ++c) {
    p.mul[c] = mul3_host[c];
    p.bias[c] = bias3_host[c];
    if (bias3_host[c] != 0.f) p.has_bias = 1;
  }
  p.batch = batch;
  p.src_h = src_h;
  p.out_h = out_h;
  p.out_w = out_w;
  p.n_msteps = n_msteps;
  p.n_groups = n_groups;
  p.n_tiles = n_tiles;
  p.row_bytes = 3 * src_w;
  p.q_stride = q_stride;
  for (int c = 0; c < 4; ++c) p.c_row[c] = c_row4_host[c];
  // the two K slots of a fragment register read two adjacent rows of one octet
  for (int q = 0; q < 4; ++q)
    for (int c = 0; c < 4; c += 2) {
      const int r = q_stride * q + c_row4_host[c];
      if (r < 0 || r > 14 || c_row4_host[c + 1] != c_row4_host[c] + 1 || (r & 7) == 7) return SIA_E_INVALID;
    }
  const size_t tables = (size_t)n_tiles * 4 * 32 * 8 + (size_t)n_msteps * 4 + (size_t)(n_groups + 1) * 4 +
                        (size_t)n_tiles * 4 + 8;
  const size_t pub = (size_t)warps * PM_PUB_WORDS * 4 + 8 + (size_t)2 * warps * 4 + 8;
  // the ring: the 2*kv octets of a window + as many octets of prefetch as fit (at least 2).  (The last group's loads may
  // run past the last slot's end, into the tables behind the ring: read-only garbage that meets zero weights.)
  const size_t octet = (size_t)8 * p.row_bytes;
  const size_t budget = 227 * 1024 - tables - pub - 64;
  int ring = (int)((budget - 2 * PM_WIN_BARS * 8) / octet);
  if (ring > 24) ring = 24;
  if (ring < 2 * kv + 2) return SIA_E_UNSUPPORTED;
  p.ring_octets = ring;
  const size_t smem = (size_t)ring * octet + tables + (size_t)2 * PM_WIN_BARS * 8;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (warps == 4) {
    if (kv == 3) return launch_pre_mma<3, 4>(p, smem, st);
    if (kv == 2) return launch_pre_mma<2, 4>(p, smem, st);
  } else {
    if (kv == 3) return launch_pre_mma<3, 8>(p, smem, st);
    if (kv == 2) return launch_pre_mma<2, 8>(p, smem, st);
  }
  return SIA_E_UNSUPPORTED;
}
