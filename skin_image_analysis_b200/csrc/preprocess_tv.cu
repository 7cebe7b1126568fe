// the ToneClassifier test transform (notebooks/ToneClassifier/CNNTrialDataset.py:71-76)
//
//     v2.Resize((224, 224))                      uint8, bilinear, antialias=True
//     v2.ToDtype(torch.float32, scale=True)      u8 * (1/255)
//     v2.Normalize(mean, std)                    (x - mean[c]) / std[c]
//
// for a whole batch of decoded u8 HWC images resident in HBM.  torchvision's uint8 resize is ATen's
// Pillow-style FIXED-POINT separable resampler (int16 taps, horizontal pass -> uint8 -> vertical pass ->
// uint8), so this is integer work and the kernel is bit-exact, not "within a tolerance":
//
//     h[r][j][c]   = clamp_u8((sum_t xw[j][t] * src[r][xmin[j] + t][c] + 2^(xp-1)) >> xp)
//     out[i][j][c] = lut[c][ clamp_u8((sum_t yw[i][t] * h[ymin[i] + t][j][c] + 2^(yp-1)) >> yp) ]
//
// The taps come from the host (resize_weights.build_tv_tables, same formulas as ATen's
// _compute_index_ranges_int16_weights); lut[c][b] is the float32 value ToDtype + Normalize give to byte b,
// computed on the host with the same float32 operations, so the float tensor is bit-identical too.
//
// One CTA = (image, tile of tile_rows output rows): the source rows the tile needs are copied once into shared
// memory with 16-byte loads (HBM reads every source byte once per tile, halo rows twice), the horizontal pass
// writes its uint8 rows into a second shared buffer, the vertical pass reads them and stores the output layout
// with coalesced stores.  HBM-bound by design: src_h*src_w*3 bytes in, out_h*out_w*3*sizeof(out) bytes out.
#include <cuda_bf16.h>

#include "sia_host.cuh"
#include "sia_ptx.cuh"

namespace sia {

constexpr int TV_THREADS = 512;

struct TvParams {
  const uint8_t* src;
  const int32_t* x_min;   // [out_w]
  const int16_t* x_w;     // [x_taps][out_w]  (tap-major: lanes of a warp read consecutive shorts)
  const int32_t* y_min;   // [out_h]
  const int16_t* y_w;     // [out_h][y_taps]
  const float* lut;       // [3][256]
  void* dst;
  int batch, src_h, src_w, out_h, out_w;
  int x_taps, y_taps, x_prec, y_prec;
  int tile_rows, max_rows;          // output rows per CTA; capacity (source rows) of the shared window
  int planar;                       // 0: src is [B,H,W,3] (HWC); 1: src is [B,3,H,W] (planar CHW)
};

__host__ __device__ inline size_t tv_rows_bytes(int max_rows, int src_w, int planar) {
  return planar ? 3 * (((size_t)max_rows * src_w + 31) & ~(size_t)15) : (((size_t)max_rows * src_w * 3 + 32 + 15) & ~(size_t)15);
}

__device__ __forceinline__ int clamp_u8(int v) { return min(max(v, 0), 255); }

// Copies len bytes from global g into shared memory at dst16 (16-byte aligned) + (g & 15): the same misalignment
// on both sides, so the body moves as aligned 16-byte vectors; returns where byte 0 landed.  All threads call it.
__device__ __forceinline__ const uint8_t* stage_span(uint8_t* dst16, const uint8_t* g, uint32_t len) {
  const uint32_t mis = (uint32_t)(reinterpret_cast<uintptr_t>(g) & 15u);
  uint8_t* s = dst16 + mis;
  const uint32_t head = min(len, (16u - mis) & 15u);
  const uint32_t nvec = (len - head) >> 4;
  const uint32_t tail0 = head + (nvec << 4);
  for (uint32_t b = threadIdx.x; b < head; b += TV_THREADS) s[b] = g[b];
  const uint4* gv = reinterpret_cast<const uint4*>(g + head);
  uint4* sv = reinterpret_cast<uint4*>(s + head);
  for (uint32_t v = threadIdx.x; v < nvec; v += TV_THREADS) sv[v] = __ldg(gv + v);
  for (uint32_t b = tail0 + threadIdx.x; b < len; b += TV_THREADS) s[b] = g[b];
  return s;
}

// XT / YT: compile-time tap counts (fully unrolled passes, horizontal taps held in registers) or 0 = run-time loops.
template <int LAYOUT, int XT, int YT>
__global__ void __launch_bounds__(TV_THREADS)
preprocess_tv_kernel(const __grid_constant__ TvParams p) {
  extern __shared__ __align__(128) uint8_t smem_tv[];
  const int row_bytes = p.src_w * 3;
  const int hrow = p.out_w * 3;
  const int hpitch = (hrow + 3) & ~3;
  uint8_t* s_rows = smem_tv;                                                  // max_rows * row_bytes + 32
  uint8_t* s_h = s_rows + tv_rows_bytes(p.max_rows, p.src_w, p.planar);                   // max_rows * hpitch
  float* s_lut = reinterpret_cast<float*>(s_h + (((size_t)p.max_rows * hpitch + 15) & ~(size_t)15));
  int32_t* s_xmin = reinterpret_cast<int32_t*>(s_lut + 768);
  int16_t* s_xw = reinterpret_cast<int16_t*>(s_xmin + p.out_w);

  const int n = blockIdx.y;
  const int i0 = blockIdx.x * p.tile_rows;
  const int i1 = min(i0 + p.tile_rows, p.out_h);
  const int r0 = p.y_min[i0];
  const int rows = min(p.y_min[i1 - 1] + p.y_taps - r0, p.max_rows);
  const uint8_t* s_plane[3];

  // ---- stage the source window: rows [r0, r0 + rows) are one contiguous span of the decode buffer (HWC) or one
  //      span per colour plane (planar CHW input, what torchvision.io.read_image / decode_jpeg produce) ---------
  int row_stride, px_stride;           // s_plane[c] + r * row_stride + x * px_stride = channel c of pixel (r0 + r, x)
  if (p.planar) {
    const uint32_t cap = ((uint32_t)p.max_rows * p.src_w + 31u) & ~15u;
    for (int c = 0; c < 3; ++c)          // each plane keeps its own misalignment
      s_plane[c] = stage_span(s_rows + c * cap, p.src + (((size_t)n * 3 + c) * p.src_h + r0) * p.src_w,
                              (uint32_t)rows * p.src_w);
    row_stride = p.src_w; px_stride = 1;
  } else {
    const uint8_t* s0 = stage_span(s_rows, p.src + ((size_t)n * p.src_h + r0) * row_bytes, (uint32_t)rows * row_bytes);
    s_plane[0] = s0; s_plane[1] = s0 + 1; s_plane[2] = s0 + 2;
    row_stride = row_bytes; px_stride = 3;
  }
  for (int k = threadIdx.x; k < 768; k += TV_THREADS) s_lut[k] = p.lut[k];
  for (int k = threadIdx.x; k < p.out_w; k += TV_THREADS) s_xmin[k] = p.x_min[k];
  for (int k = threadIdx.x; k < p.out_w * p.x_taps; k += TV_THREADS) s_xw[k] = p.x_w[k];
  __syncthreads();

  // ---- horizontal pass: every staged source row -> out_w uint8 pixels -------------------------------------
  if constexpr (XT > 0) {
    // thread = (output column, row group): the column's taps and window offset stay in registers for all rows
    constexpr int COLS = 256, GROUPS = TV_THREADS / COLS;
    const int half = 1 << (p.x_prec - 1);
    const int grp = threadIdx.x / COLS;
    for (int j = threadIdx.x % COLS; j < p.out_w; j += COLS) {
      int w[XT];
#pragma unroll
      for (int t = 0; t < XT; ++t) w[t] = s_xw[t * p.out_w + j];
      const size_t off0 = (size_t)px_stride * s_xmin[j];
      for (int r = grp; r < rows; r += GROUPS) {
        const size_t off = (size_t)r * row_stride + off0;
        const uint8_t *p0 = s_plane[0] + off, *p1 = s_plane[1] + off, *p2 = s_plane[2] + off;
        int v0[XT], v1[XT], v2[XT];
#pragma unroll
        for (int t = 0; t < XT; ++t) { v0[t] = p0[px_stride * t]; v1[t] = p1[px_stride * t]; v2[t] = p2[px_stride * t]; }
        int a0 = half, a1 = half, a2 = half;
#pragma unroll
        for (int t = 0; t < XT; ++t) { a0 += w[t] * v0[t]; a1 += w[t] * v1[t]; a2 += w[t] * v2[t]; }
        uint8_t* h = s_h + (size_t)r * hpitch + 3 * j;
        h[0] = (uint8_t)clamp_u8(a0 >> p.x_prec);
        h[1] = (uint8_t)clamp_u8(a1 >> p.x_prec);
        h[2] = (uint8_t)clamp_u8(a2 >> p.x_prec);
      }
    }
  } else {
    const int half = 1 << (p.x_prec - 1);
    const int total = rows * p.out_w;
    for (int idx = threadIdx.x; idx < total; idx += TV_THREADS) {
      const int r = idx / p.out_w, j = idx - r * p.out_w;
      const size_t off = (size_t)r * row_stride + (size_t)px_stride * s_xmin[j];
      const uint8_t *p0 = s_plane[0] + off, *p1 = s_plane[1] + off, *p2 = s_plane[2] + off;
      int a0 = half, a1 = half, a2 = half;
      for (int t = 0; t < p.x_taps; ++t) {
        const int w = s_xw[t * p.out_w + j];
        a0 += w * p0[px_stride * t];
        a1 += w * p1[px_stride * t];
        a2 += w * p2[px_stride * t];
      }
      uint8_t* h = s_h + (size_t)r * hpitch + 3 * j;
      h[0] = (uint8_t)clamp_u8(a0 >> p.x_prec);
      h[1] = (uint8_t)clamp_u8(a1 >> p.x_prec);
      h[2] = (uint8_t)clamp_u8(a2 >> p.x_prec);
    }
  }
  __syncthreads();

  // ---- vertical pass + ToDtype/Normalize (table) + output layout -------------------------------------------
  {
    const int half = 1 << (p.y_prec - 1);
    const int total = (i1 - i0) * p.out_w;
    for (int idx = threadIdx.x; idx < total; idx += TV_THREADS) {
      const int di = idx / p.out_w, j = idx - di * p.out_w;
      const int i = i0 + di;
      const uint8_t* h = s_h + (size_t)(p.y_min[i] - r0) * hpitch + 3 * j;
      const int16_t* yw = p.y_w + (size_t)i * p.y_taps;
      int a0 = half, a1 = half, a2 = half;
      if constexpr (YT > 0) {
        int w[YT], v0[YT], v1[YT], v2[YT];
#pragma unroll
        for (int t = 0; t < YT; ++t) {
          w[t] = __ldg(yw + t);
          v0[t] = h[(size_t)t * hpitch]; v1[t] = h[(size_t)t * hpitch + 1]; v2[t] = h[(size_t)t * hpitch + 2];
        }
#pragma unroll
        for (int t = 0; t < YT; ++t) { a0 += w[t] * v0[t]; a1 += w[t] * v1[t]; a2 += w[t] * v2[t]; }
      } else {
        for (int t = 0; t < p.y_taps; ++t) {
          const int w = __ldg(yw + t);
          a0 += w * h[(size_t)t * hpitch];
          a1 += w * h[(size_t)t * hpitch + 1];
          a2 += w * h[(size_t)t * hpitch + 2];
        }
      }
      const float f0 = s_lut[clamp_u8(a0 >> p.y_prec)];
      const float f1 = s_lut[256 + clamp_u8(a1 >> p.y_prec)];
      const float f2 = s_lut[512 + clamp_u8(a2 >> p.y_prec)];
      if constexpr (LAYOUT == SIA_LAYOUT_NCHW_F32) {
        float* d = static_cast<float*>(p.dst) + (((size_t)n * 3) * p.out_h + i) * p.out_w + j;
        const size_t plane = (size_t)p.out_h * p.out_w;
        d[0] = f0;
        d[plane] = f1;
        d[2 * plane] = f2;
      } else if constexpr (LAYOUT == SIA_LAYOUT_NCHW_BF16) {
        __nv_bfloat16* d = static_cast<__nv_bfloat16*>(p.dst) + (((size_t)n * 3) * p.out_h + i) * p.out_w + j;
        const size_t plane = (size_t)p.out_h * p.out_w;
        d[0] = __float2bfloat16_rn(f0);
        d[plane] = __float2bfloat16_rn(f1);
        d[2 * plane] = __float2bfloat16_rn(f2);
      } else {
        const int pitch = p.out_w + SIA_NHWC4_PAD;
        uint2* d = static_cast<uint2*>(p.dst) + ((size_t)n * p.out_h + i) * pitch;
        d[j + 1] = make_uint2(pack_bf16x2(f0, f1), pack_bf16x2(f2, 0.f));
        if (j == 0) d[0] = make_uint2(0u, 0u);
        if (j == p.out_w - 1) {
#pragma unroll
          for (int c = 1; c < SIA_NHWC4_PAD; ++c) d[p.out_w + c] = make_uint2(0u, 0u);
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// Fast instance for the case the transform is used for (interleaved HWC source whose rows are a multiple of 4
// bytes, 7 x 7 taps = down-sampling by 2..3 on both axes, out_w % 4 == 0): same staging, same integer results,
// but both passes read shared memory as 32-bit words and multiply-accumulate two taps per instruction with
// IDP.2A (dp2a: s16 tap pair x u8 byte pair -> s32), the byte pairs gathered by PRMT with compile-time selectors.
//   horizontal: thread = output column; its 21-byte window [3*xmin, 3*xmin+21) is 6 aligned words, realigned by 5
//               funnel shifts; tap pairs (0,1), (2,3), (4,5) + tap 6 -> 8 PRMT + 12 IDP.2A per (row, column).
//   vertical:   thread = 4 output pixels = 3 words of each of the 7 horizontal-pass rows; adjacent rows are paired
//               byte-wise by PRMT -> 24 PRMT + 48 IDP.2A for 12 output bytes, then table + 32 bytes of output.
__device__ __forceinline__ int dp2a_lo_s16u8(uint32_t w2, uint32_t b4, int c) {
  int d;
  asm("dp2a.lo.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(w2), "r"(b4), "r"(c));
  return d;
}
__device__ __forceinline__ int dp2a_hi_s16u8(uint32_t w2, uint32_t b4, int c) {
  int d;
  asm("dp2a.hi.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(w2), "r"(b4), "r"(c));
  return d;
}
__device__ __forceinline__ uint32_t pack_s16x2(int lo, int hi) {
  return ((uint32_t)lo & 0xFFFFu) | ((uint32_t)hi << 16);
}

template <int LAYOUT>
__global__ void __launch_bounds__(TV_THREADS)
preprocess_tv_fast_kernel(const __grid_constant__ TvParams p) {
  extern __shared__ __align__(128) uint8_t smem_tv[];
  const int row_bytes = p.src_w * 3;                 // % 4 == 0 (checked by the launcher)
  const int hpitch = (p.out_w * 3 + 3) & ~3;
  uint8_t* s_rows = smem_tv;
  uint8_t* s_h = s_rows + tv_rows_bytes(p.max_rows, p.src_w, 0);
  float* s_lut = reinterpret_cast<float*>(s_h + (((size_t)p.max_rows * hpitch + 15) & ~(size_t)15));
  int32_t* s_xmin = reinterpret_cast<int32_t*>(s_lut + 768);
  int16_t* s_xw = reinterpret_cast<int16_t*>(s_xmin + p.out_w);

  const int n = blockIdx.y;
  const int i0 = blockIdx.x * p.tile_rows;
  const int i1 = min(i0 + p.tile_rows, p.out_h);
  const int r0 = p.y_min[i0];
  const int rows = min(p.y_min[i1 - 1] + 7 - r0, p.max_rows);

  const uint8_t* s0 = stage_span(s_rows, p.src + ((size_t)n * p.src_h + r0) * row_bytes, (uint32_t)rows * row_bytes);
  for (int k = threadIdx.x; k < 768; k += TV_THREADS) s_lut[k] = p.lut[k];
  for (int k = threadIdx.x; k < p.out_w; k += TV_THREADS) s_xmin[k] = p.x_min[k];
  for (int k = threadIdx.x; k < p.out_w * 7; k += TV_THREADS) s_xw[k] = p.x_w[k];
  __syncthreads();

  // ---- horizontal pass ------------------------------------------------------------------------------------
  {
    constexpr int COLS = 256, GROUPS = TV_THREADS / COLS;
    const int half = 1 << (p.x_prec - 1);
    const int grp = threadIdx.x / COLS;
    for (int j = threadIdx.x % COLS; j < p.out_w; j += COLS) {
      const uint32_t w01 = pack_s16x2(s_xw[j], s_xw[p.out_w + j]);
      const uint32_t w23 = pack_s16x2(s_xw[2 * p.out_w + j], s_xw[3 * p.out_w + j]);
      const uint32_t w45 = pack_s16x2(s_xw[4 * p.out_w + j], s_xw[5 * p.out_w + j]);
      const uint32_t w6 = pack_s16x2(s_xw[6 * p.out_w + j], 0);
      const uint32_t a0 = smem_u32(s0) + 3u * (uint32_t)s_xmin[j];
      const uint32_t base = a0 & ~3u, sh = (a0 & 3u) * 8u;      // rows are a multiple of 4 bytes: same shift in every row
      uint8_t* hcol = s_h + 3 * j;
      for (int r = grp; r < rows; r += GROUPS) {
        const uint32_t a = base + (uint32_t)r * row_bytes;
        uint32_t q[6];
#pragma unroll
        for (int k = 0; k < 6; ++k) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(q[k]) : "r"(a + 4u * k));
        uint32_t w[6];                                           // the window byte-aligned: w[k] = bytes 4k .. 4k+3
#pragma unroll
        for (int k = 0; k < 5; ++k) w[k] = __funnelshift_r(q[k], q[k + 1], sh);
        w[5] = q[5] >> sh;
        uint32_t t = __byte_perm(w[0], w[1], 0x4130);           // R0 R1 G0 G1
        int c0 = dp2a_lo_s16u8(w01, t, half), c1 = dp2a_hi_s16u8(w01, t, half);
        int c2 = dp2a_lo_s16u8(w01, __byte_perm(w[0], w[1], 0x0052), half);          // B0 B1
        t = __byte_perm(w[1], w[2], 0x6352);                    // R2 R3 G2 G3
        c0 = dp2a_lo_s16u8(w23, t, c0); c1 = dp2a_hi_s16u8(w23, t, c1);
        c2 = dp2a_lo_s16u8(w23, __byte_perm(w[1], w[2], 0x0074), c2);
        t = __byte_perm(w[3], w[4], 0x4130);                    // R4 R5 G4 G5
        c0 = dp2a_lo_s16u8(w45, t, c0); c1 = dp2a_hi_s16u8(w45, t, c1);
        c2 = dp2a_lo_s16u8(w45, __byte_perm(w[3], w[4], 0x0052), c2);
        t = __byte_perm(w[4], w[5], 0x3322);                    // R6 R6 G6 G6 (second tap of the pair has weight 0)
        c0 = dp2a_lo_s16u8(w6, t, c0); c1 = dp2a_hi_s16u8(w6, t, c1);
        c2 = dp2a_lo_s16u8(w6, __byte_perm(w[4], w[5], 0x0044), c2);
        uint8_t* h = hcol + (size_t)r * hpitch;
        h[0] = (uint8_t)clamp_u8(c0 >> p.x_prec);
        h[1] = (uint8_t)clamp_u8(c1 >> p.x_prec);
        h[2] = (uint8_t)clamp_u8(c2 >> p.x_prec);
      }
    }
  }
  __syncthreads();

  // ---- vertical pass: 4 pixels (12 bytes of the horizontal-pass rows) per thread ---------------------------
  {
    const int half = 1 << (p.y_prec - 1);
    const int quads = p.out_w >> 2;
    const int total = (i1 - i0) * quads;
    for (int idx = threadIdx.x; idx < total; idx += TV_THREADS) {
      const int di = idx / quads, qd = idx - di * quads;
      const int i = i0 + di;
      const int16_t* yw = p.y_w + (size_t)i * 7;
      const uint32_t w01 = pack_s16x2(__ldg(yw), __ldg(yw + 1)), w23 = pack_s16x2(__ldg(yw + 2), __ldg(yw + 3));
      const uint32_t w45 = pack_s16x2(__ldg(yw + 4), __ldg(yw + 5)), w6 = pack_s16x2(__ldg(yw + 6), 0);
      const uint32_t a = smem_u32(s_h) + (uint32_t)(p.y_min[i] - r0) * hpitch + 12u * qd;
      int acc[12];
#pragma unroll
      for (int m = 0; m < 3; ++m) {
        uint32_t h[7];
#pragma unroll
        for (int t = 0; t < 7; ++t) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(h[t]) : "r"(a + (uint32_t)t * hpitch + 4u * m));
        uint32_t lo = __byte_perm(h[0], h[1], 0x5140), hi = __byte_perm(h[0], h[1], 0x7362);
        int b0 = dp2a_lo_s16u8(w01, lo, half), b1 = dp2a_hi_s16u8(w01, lo, half);
        int b2 = dp2a_lo_s16u8(w01, hi, half), b3 = dp2a_hi_s16u8(w01, hi, half);
        lo = __byte_perm(h[2], h[3], 0x5140); hi = __byte_perm(h[2], h[3], 0x7362);
        b0 = dp2a_lo_s16u8(w23, lo, b0); b1 = dp2a_hi_s16u8(w23, lo, b1);
        b2 = dp2a_lo_s16u8(w23, hi, b2); b3 = dp2a_hi_s16u8(w23, hi, b3);
        lo = __byte_perm(h[4], h[5], 0x5140); hi = __byte_perm(h[4], h[5], 0x7362);
        b0 = dp2a_lo_s16u8(w45, lo, b0); b1 = dp2a_hi_s16u8(w45, lo, b1);
        b2 = dp2a_lo_s16u8(w45, hi, b2); b3 = dp2a_hi_s16u8(w45, hi, b3);
        lo = __byte_perm(h[6], h[6], 0x1100); hi = __byte_perm(h[6], h[6], 0x3322);
        b0 = dp2a_lo_s16u8(w6, lo, b0); b1 = dp2a_hi_s16u8(w6, lo, b1);
        b2 = dp2a_lo_s16u8(w6, hi, b2); b3 = dp2a_hi_s16u8(w6, hi, b3);
        acc[4 * m] = b0; acc[4 * m + 1] = b1; acc[4 * m + 2] = b2; acc[4 * m + 3] = b3;
      }
      float f[12];                                               // byte b of the quad = channel b % 3 of pixel b / 3
#pragma unroll
      for (int b = 0; b < 12; ++b) f[b] = s_lut[256 * (b % 3) + clamp_u8(acc[b] >> p.y_prec)];
      const int j0 = 4 * qd;
      if constexpr (LAYOUT == SIA_LAYOUT_NCHW_F32) {
        float* d = static_cast<float*>(p.dst) + (((size_t)n * 3) * p.out_h + i) * p.out_w + j0;
        const size_t plane = (size_t)p.out_h * p.out_w;
#pragma unroll
        for (int c = 0; c < 3; ++c)
          *reinterpret_cast<float4*>(d + c * plane) = make_float4(f[c], f[3 + c], f[6 + c], f[9 + c]);
      } else if constexpr (LAYOUT == SIA_LAYOUT_NCHW_BF16) {
        __nv_bfloat16* d = static_cast<__nv_bfloat16*>(p.dst) + (((size_t)n * 3) * p.out_h + i) * p.out_w + j0;
        const size_t plane = (size_t)p.out_h * p.out_w;
#pragma unroll
        for (int c = 0; c < 3; ++c)
          *reinterpret_cast<uint2*>(d + c * plane) = make_uint2(pack_bf16x2(f[c], f[3 + c]), pack_bf16x2(f[6 + c], f[9 + c]));
      } else {
        const int pitch = p.out_w + SIA_NHWC4_PAD;
        uint2* d = static_cast<uint2*>(p.dst) + ((size_t)n * p.out_h + i) * pitch;
#pragma unroll
        for (int u = 0; u < 4; ++u)
          d[j0 + 1 + u] = make_uint2(pack_bf16x2(f[3 * u], f[3 * u + 1]), pack_bf16x2(f[3 * u + 2], 0.f));
        if (qd == 0) d[0] = make_uint2(0u, 0u);
        if (qd == quads - 1) {
#pragma unroll
          for (int c = 1; c < SIA_NHWC4_PAD; ++c) d[p.out_w + c] = make_uint2(0u, 0u);
        }
      }
    }
  }
}

}  // namespace sia

// Debug / A-B: != 0 routes every call through the byte-wise kernel even where the IDP.2A instance applies.
static int g_tv_force_generic = 0;
extern "C" int sia_debug_tv_force_generic(int on) {
  g_tv_force_generic = on;
  return 0;
}

extern "C" int sia_preprocess_tv_u8hwc(const uint8_t* src, int batch, int src_h, int src_w, const int32_t* x_min,
                                       const int16_t* x_w_tapmajor, int x_taps, int x_prec, const int32_t* y_min,
                                       const int16_t* y_w, int y_taps, int y_prec, const float* lut_3x256, int out_h,
                                       int out_w, int tile_rows, int max_window_rows, int layout, int planar_chw,
                                       void* dst, void* stream) {
  using namespace sia;
  SIA_REQUIRE(src && x_min && x_w_tapmajor && y_min && y_w && lut_3x256 && dst);
  SIA_REQUIRE(batch >= 1 && src_h >= 1 && src_w >= 1 && out_h >= 1 && out_w >= 1);
  SIA_REQUIRE(x_taps >= 1 && y_taps >= 1 && x_prec >= 1 && x_prec < 23 && y_prec >= 1 && y_prec < 23);
  SIA_REQUIRE(tile_rows >= 1 && max_window_rows >= y_taps && max_window_rows <= src_h);
  SIA_REQUIRE(layout == SIA_LAYOUT_NCHW_F32 || layout == SIA_LAYOUT_NCHW_BF16 || layout == SIA_LAYOUT_NHWC4_BF16);
  if (batch > 65535) return SIA_E_UNSUPPORTED;
  const size_t hpitch = ((size_t)out_w * 3 + 3) & ~(size_t)3;
  const size_t smem = tv_rows_bytes(max_window_rows, src_w, planar_chw) +
                      (((size_t)max_window_rows * hpitch + 15) & ~(size_t)15) + 768 * 4 + (size_t)out_w * 4 +
                      (size_t)out_w * x_taps * 2 + 16;
  if (smem > 227 * 1024) return SIA_E_UNSUPPORTED;

  TvParams p;
  p.src = src; p.x_min = x_min; p.x_w = x_w_tapmajor; p.y_min = y_min; p.y_w = y_w; p.lut = lut_3x256; p.dst = dst;
  p.batch = batch; p.src_h = src_h; p.src_w = src_w; p.out_h = out_h; p.out_w = out_w;
  p.x_taps = x_taps; p.y_taps = y_taps; p.x_prec = x_prec; p.y_prec = y_prec;
  p.tile_rows = tile_rows; p.max_rows = max_window_rows; p.planar = planar_chw ? 1 : 0;
  const dim3 grid((out_h + tile_rows - 1) / tile_rows, batch);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // 7 x 7 taps = down-sampling by 2..3 on both axes (600x450 -> 224x224, the case the transform is used for)
  const bool unrolled = x_taps == 7 && y_taps == 7;
  const bool fast = unrolled && !planar_chw && (src_w * 3) % 4 == 0 && out_w % 4 == 0 && !g_tv_force_generic;
  static SmemSlots configured[9] = {};
#define SIA_TV_LAUNCH_FAST(LAYOUT, SLOT)                                                                     \
  do {                                                                                                       \
    if (int rc = ensure_dynamic_smem(preprocess_tv_fast_kernel<LAYOUT>, (int)smem, &configured[SLOT])) return rc; \
    preprocess_tv_fast_kernel<LAYOUT><<<grid, TV_THREADS, smem, st>>>(p);                                     \
  } while (0)
  if (fast) {
    if (layout == SIA_LAYOUT_NCHW_F32) SIA_TV_LAUNCH_FAST(SIA_LAYOUT_NCHW_F32, 6);
    else if (layout == SIA_LAYOUT_NCHW_BF16) SIA_TV_LAUNCH_FAST(SIA_LAYOUT_NCHW_BF16, 7);
    else SIA_TV_LAUNCH_FAST(SIA_LAYOUT_NHWC4_BF16, 8);
    return launch_status();
  }
#undef SIA_TV_LAUNCH_FAST
#define SIA_TV_LAUNCH(LAYOUT, XT, YT, SLOT)                                                                  \
  do {                                                                                                       \
    if (int rc = ensure_dynamic_smem(preprocess_tv_kernel<LAYOUT, XT, YT>, (int)smem, &configured[SLOT])) return rc; \
    preprocess_tv_kernel<LAYOUT, XT, YT><<<grid, TV_THREADS, smem, st>>>(p);                                  \
  } while (0)
  if (layout == SIA_LAYOUT_NCHW_F32) {
    if (unrolled) SIA_TV_LAUNCH(SIA_LAYOUT_NCHW_F32, 7, 7, 0); else SIA_TV_LAUNCH(SIA_LAYOUT_NCHW_F32, 0, 0, 1);
  } else if (layout == SIA_LAYOUT_NCHW_BF16) {
    if (unrolled) SIA_TV_LAUNCH(SIA_LAYOUT_NCHW_BF16, 7, 7, 2); else SIA_TV_LAUNCH(SIA_LAYOUT_NCHW_BF16, 0, 0, 3);
  } else {
    if (unrolled) SIA_TV_LAUNCH(SIA_LAYOUT_NHWC4_BF16, 7, 7, 4); else SIA_TV_LAUNCH(SIA_LAYOUT_NHWC4_BF16, 0, 0, 5);
  }
#undef SIA_TV_LAUNCH
  return launch_status();
}
