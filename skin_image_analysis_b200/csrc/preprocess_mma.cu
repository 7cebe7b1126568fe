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
//   * the compute warps of a CTA split the groups of a row; each also runs the first product of the last group of
//     its left neighbour (halo) so that every n-tile is finished by exactly one warp without any exchange.
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
  uint4* wy_s = reinterpret_cast<uint4*>(pm_smem + off);
  off += (size_t)p.n_msteps * KV * 32 * sizeof(uint4);
  uint2* wx_s = reinterpret_cast<uint2*>(pm_smem + off);
  off += (size_t)p.n_tiles * 4 * 32 * sizeof(uint2);
  int* r0_s = reinterpret_cast<int*>(pm_smem + off);
  off += (size_t)p.n_msteps * sizeof(int);
  int* tbeg_s = reinterpret_cast<int*>(pm_smem + off);
  off += (size_t)(p.n_groups + 1) * sizeof(int);
  uint32_t* mask_s = reinterpret_cast<uint32_t*>(pm_smem + off);
  off += (size_t)p.n_tiles * sizeof(uint32_t);
  off = (off + 7) & ~(size_t)7;
  uint64_t* full = reinterpret_cast<uint64_t*>(pm_smem + off);
  uint64_t* empty = full + R;

  if (threadIdx.x == 0) {
    for (int i = 0; i < R; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], NW);
    }
    fence_mbar_init();
  }
  __syncthreads();

  // ---- this CTA's contiguous range of m-steps ------------------------------------------------------------------------
  const long long total = (long long)p.batch * p.n_msteps;
  const int u_lo = (int)(total * blockIdx.x / gridDim.x);
  const int u_hi = (int)(total * (blockIdx.x + 1) / gridDim.x);
  const int n_oct_img = (p.src_h + 7) >> 3;
  const size_t image_bytes = (size_t)p.src_h * p.row_bytes;

  if (warp == NW) {
    // =============================== copy warp: one lane streams the source octets =================================
    if (lane == 0) {
      int seq = 0;
      int u = u_lo;
      while (u < u_hi) {
        const int img = u / p.n_msteps;
        const int pass_end = min(u_hi, (img + 1) * p.n_msteps);
        const uint8_t* image = p.src + (size_t)img * image_bytes;
        int o_next = __ldg(&p.r0[u % p.n_msteps]) >> 3;
        for (int uu = u; uu < pass_end; ++uu) {
          const int o_end = min((__ldg(&p.r0[uu % p.n_msteps]) >> 3) + 2 * KV, n_oct_img);
          for (int o = o_next; o < o_end; ++o, ++seq) {
            const int slot = seq % R;
            if (seq >= R) mbar_wait(&empty[slot], ((seq / R) & 1) ^ 1, 61);
            const uint32_t bytes = (uint32_t)(min(8, p.src_h - o * 8) * p.row_bytes);
#ifdef SIA_PM_NO_COPY            // timing variant: no source traffic at all, the barriers still cycle
            mbar_arrive(&full[slot]);
#else
            mbar_arrive_expect_tx(&full[slot], bytes);
            bulk_load_1d(ring + (size_t)slot * octet_bytes, image + (size_t)o * octet_bytes, bytes, &full[slot]);
#endif
          }
          o_next = max(o_next, o_end);
        }
        u = pass_end;
      }
    }
    return;
  }

  // ================================== compute warps ================================================================
  // tables -> shared memory (the copy warp is already fetching the first octets)
  {
    const int tid = threadIdx.x, nt = NW * 32;
    for (int i = tid; i < p.n_msteps * KV * 32; i += nt) wy_s[i] = __ldg(&p.wy_frag[i]);
    for (int i = tid; i < p.n_tiles * 4 * 32; i += nt) wx_s[i] = __ldg(&p.wx_frag[i]);
    for (int i = tid; i < p.n_msteps; i += nt) r0_s[i] = __ldg(&p.r0[i]);
    for (int i = tid; i <= p.n_groups; i += nt) tbeg_s[i] = __ldg(&p.tile_begin[i]);
    for (int i = tid; i < p.n_tiles; i += nt) mask_s[i] = __ldg(&p.wx_mask[i]);
    asm volatile("bar.sync 1, %0;" ::"n"(NW * 32) : "memory");
  }

  // column groups of this warp: [g_first, g_end) are its own, one halo group (first product only) before them
  const int g_first = (p.n_groups * warp + NW - 1) / NW;
  const int g_end = (p.n_groups * (warp + 1) + NW - 1) / NW;
  const int g_start = max(g_first - 1, 0);
  // the two row pairs this lane reads in every 16-row chunk: chunk rows rowA, rowA + 1 (K slots 0, 1) and rowB, rowB + 1
  const int row_a = p.q_stride * q + p.c_row[0], row_b = p.q_stride * q + p.c_row[2];
  const int oct_a = row_a >> 3, oct_b = row_b >> 3;                              // 0 or 1: which octet of the chunk
  const uint32_t off_a = (uint32_t)((row_a & 7) * p.row_bytes + 12 * g);
  const uint32_t off_b = (uint32_t)((row_b & 7) * p.row_bytes + 12 * g);
  const uint32_t next_row = (uint32_t)p.row_bytes;
  const uint32_t ring_u32 = smem_u32(ring);
  const int out_pitch_px = p.out_w + SIA_NHWC4_PAD;
  const float mul0 = p.mul[0], mul1 = p.mul[1], mul2 = p.mul[2];

  int seq_base = 0;       // sequence number of the first octet of the current pass
  int u = u_lo;
  while (u < u_hi) {
    const int img = u / p.n_msteps;
    const int pass_end = min(u_hi, (img + 1) * p.n_msteps);
    const int o_start = r0_s[u % p.n_msteps] >> 3;
    int o_waited = o_start, o_released = o_start, o_issued_end = o_start;
    for (int uu = u; uu < pass_end; ++uu) {
      const int m = uu % p.n_msteps;
      const int o0 = r0_s[m] >> 3;
      const int o_win_end = min(o0 + 2 * KV, n_oct_img);
      o_issued_end = max(o_issued_end, o_win_end);
      for (; o_waited < o_win_end; ++o_waited) {
        const int s = seq_base + (o_waited - o_start);
        mbar_wait(&full[s % R], (s / R) & 1, 62);
      }

#ifdef SIA_PM_NO_COMPUTE          // timing variant: the compute warps only wait for and release the octets
      if (false) {
#else
      if (g_end > g_first) {
#endif
        const int row0 = m * 16 + g, row1 = row0 + 8;
        uint8_t* out_row0 = p.dst + ((size_t)img * p.out_h + row0) * out_pitch_px * 8;
        uint8_t* out_row1 = out_row0 + (size_t)8 * out_pitch_px * 8;
        const bool ok0 = row0 < p.out_h, ok1 = row1 < p.out_h;

        // shared-memory addresses of this lane's two row pairs in each 16-row chunk of the window
        uint32_t pa[KV], pb[KV];
#pragma unroll
        for (int kc = 0; kc < KV; ++kc) {
          const int sa = seq_base + (o0 + 2 * kc + oct_a - o_start), sb = seq_base + (o0 + 2 * kc + oct_b - o_start);
          pa[kc] = ring_u32 + (uint32_t)((sa % R) * octet_bytes) + off_a;
          pb[kc] = ring_u32 + (uint32_t)((sb % R) * octet_bytes) + off_b;
        }

        uint32_t prev[3][2][4];                                    // A fragments of the previous group
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
          for (int x = 0; x < 2; ++x)
#pragma unroll
            for (int e = 0; e < 4; ++e) prev[c][x][e] = 0u;

        for (int grp = g_start; grp < g_end; ++grp) {
          // ---------------- first product: V[16 rows x 96 bytes] of this group --------------------------------------
          float vacc[12][4];
#pragma unroll
          for (int b = 0; b < 12; ++b)
#pragma unroll
            for (int e = 0; e < 4; ++e) vacc[b][e] = 0.f;
          const uint32_t goff = (uint32_t)(grp * PM_GROUP_BYTES);
#pragma unroll
          for (int kc = 0; kc < KV; ++kc) {
            const uint32_t qa = pa[kc] + goff, qb = pb[kc] + goff;
            uint32_t wa[3], wb[3], wc[3], wd[3];
#pragma unroll
            for (int w = 0; w < 3; ++w) {
              asm volatile("ld.shared.u32 %0, [%1];" : "=r"(wa[w]) : "r"(qa + 4 * w));
              asm volatile("ld.shared.u32 %0, [%1];" : "=r"(wb[w]) : "r"(qa + next_row + 4 * w));
              asm volatile("ld.shared.u32 %0, [%1];" : "=r"(wc[w]) : "r"(qb + 4 * w));
              asm volatile("ld.shared.u32 %0, [%1];" : "=r"(wd[w]) : "r"(qb + next_row + 4 * w));
            }
            const uint4 af4 = wy_s[(m * KV + kc) * 32 + lane];
            const uint32_t af[4] = {af4.x, af4.y, af4.z, af4.w};
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
          // ---------------- second product + store for every output tile whose last group this is --------------------
          if (grp >= g_first) {
            const int t_end = tbeg_s[grp + 1];
            for (int t = tbeg_s[grp]; t < t_end; ++t) {
              float hacc[3][4];
#pragma unroll
              for (int c = 0; c < 3; ++c)
#pragma unroll
                for (int e = 0; e < 4; ++e) hacc[c][e] = 0.f;
              const uint32_t mask = mask_s[t];
              const uint2* wxt = wx_s + (size_t)t * 4 * 32 + lane;
#pragma unroll
              for (int x = 0; x < 2; ++x) {
                if ((mask >> x) & 1u) {
                  const uint2 bf = wxt[x * 32];
#pragma unroll
                  for (int c = 0; c < 3; ++c) mma16816(hacc[c], prev[c][x], bf.x, bf.y);
                }
                if ((mask >> (2 + x)) & 1u) {
                  const uint2 bf = wxt[(2 + x) * 32];
#pragma unroll
                  for (int c = 0; c < 3; ++c) mma16816(hacc[c], cur[c][x], bf.x, bf.y);
                }
              }
              const int pc0 = t * 8 + 2 * q;
              float b00 = 0.f, b01 = 0.f, b02 = 0.f, b10 = 0.f, b11 = 0.f, b12 = 0.f;
              if (p.has_bias) {                                     // pad columns stay zero whatever the bias
                if (pc0 >= 1 && pc0 <= p.out_w) { b00 = p.bias[0]; b01 = p.bias[1]; b02 = p.bias[2]; }
                if (pc0 + 1 <= p.out_w) { b10 = p.bias[0]; b11 = p.bias[1]; b12 = p.bias[2]; }
              }
              if (ok0) {
                uint4 v;
                v.x = pack_bf16x2(fmaf(hacc[0][0], mul0, b00), fmaf(hacc[1][0], mul1, b01));
                v.y = pack_bf16x2(fmaf(hacc[2][0], mul2, b02), 0.f);
                v.z = pack_bf16x2(fmaf(hacc[0][1], mul0, b10), fmaf(hacc[1][1], mul1, b11));
                v.w = pack_bf16x2(fmaf(hacc[2][1], mul2, b12), 0.f);
#ifndef SIA_PM_NO_STORE
                *reinterpret_cast<uint4*>(out_row0 + (size_t)pc0 * 8) = v;
#else
                if (v.x == 0x12345678u) *reinterpret_cast<uint4*>(out_row0 + (size_t)pc0 * 8) = v;
#endif
              }
              if (ok1) {
                uint4 v;
                v.x = pack_bf16x2(fmaf(hacc[0][2], mul0, b00), fmaf(hacc[1][2], mul1, b01));
                v.y = pack_bf16x2(fmaf(hacc[2][2], mul2, b02), 0.f);
                v.z = pack_bf16x2(fmaf(hacc[0][3], mul0, b10), fmaf(hacc[1][3], mul1, b11));
                v.w = pack_bf16x2(fmaf(hacc[2][3], mul2, b12), 0.f);
#ifndef SIA_PM_NO_STORE
                *reinterpret_cast<uint4*>(out_row1 + (size_t)pc0 * 8) = v;
#else
                if (v.x == 0x12345678u) *reinterpret_cast<uint4*>(out_row1 + (size_t)pc0 * 8) = v;
#endif
              }
            }
          }
#pragma unroll
          for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int x = 0; x < 2; ++x)
#pragma unroll
              for (int e = 0; e < 4; ++e) prev[c][x][e] = cur[c][x][e];
        }
      }

      // ---------------- release the octets no later m-step of this pass reads ------------------------------------------
      const int o_rel_end = (uu + 1 < pass_end) ? min(r0_s[(uu + 1) % p.n_msteps] >> 3, n_oct_img) : o_issued_end;
      for (; o_waited < o_rel_end; ++o_waited) {            // (only if windows ever leave a gap)
        const int s = seq_base + (o_waited - o_start);
        mbar_wait(&full[s % R], (s / R) & 1, 63);
      }
      __syncwarp();
      if (lane == 0) {
        for (int o = o_released; o < o_rel_end; ++o) {
          const int s = seq_base + (o - o_start);
          mbar_arrive(&empty[s % R]);
        }
      }
      o_released = max(o_released, o_rel_end);
    }
    seq_base += o_issued_end - o_start;
    u = pass_end;
  }
}

template <int KV, int NW>
static int launch_pre_mma(const PreMmaParams& p, size_t smem, cudaStream_t st) {
  auto kern = preprocess_mma_kernel<KV, NW>;
  static SmemSlots configured = {};
  if (int rc = ensure_dynamic_smem(kern, (int)smem, &configured)) return rc;
  const long long total = (long long)p.batch * p.n_msteps;
  const int grid = (int)(total < sm_count() ? total : sm_count());
  kern<<<grid, (NW + 1) * 32, smem, st>>>(p);
  return launch_status();
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
  SIA_REQUIRE(aligned(src, 16) && aligned(dst, 16) && aligned(wy_frag, 16) && aligned(wx_frag, 8));
  if (src_w % 8 != 0 || src_h % 2 != 0 || out_w % 8 != 0) return SIA_E_UNSUPPORTED;
  if (n_msteps != (out_h + 15) / 16 || n_groups != (src_w + 31) / 32 || n_tiles != (out_w + SIA_NHWC4_PAD) / 8)
    return SIA_E_INVALID;
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
  const size_t tables = (size_t)n_msteps * kv * 32 * 16 + (size_t)n_tiles * 4 * 32 * 8 + (size_t)n_msteps * 4 +
                        (size_t)(n_groups + 1) * 4 + (size_t)n_tiles * 4 + 8;
  // the ring: the 2*kv octets of a window + as many octets of prefetch as fit (at least 2)
  const size_t octet = (size_t)8 * p.row_bytes;
  const size_t budget = 227 * 1024 - tables - 64;     // (the last group's loads may run past the last slot's end, into
                                                       // the tables behind the ring: read-only garbage, zero weights)
  int ring = (int)(budget / octet);
  if (ring > 24) ring = 24;
  if (ring < 2 * kv + 2) return SIA_E_UNSUPPORTED;
  p.ring_octets = ring;
  const size_t smem = (size_t)ring * octet + tables + (size_t)2 * ring * 8;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (g_pm_warps == 4) {
    if (kv == 3) return launch_pre_mma<3, 4>(p, smem, st);
    if (kv == 2) return launch_pre_mma<2, 4>(p, smem, st);
  } else {
    if (kv == 3) return launch_pre_mma<3, 8>(p, smem, st);
    if (kv == 2) return launch_pre_mma<2, 8>(p, smem, st);
  }
  return SIA_E_UNSUPPORTED;
}
