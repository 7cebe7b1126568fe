// K5 / K6: the fully connected part of the network.
//
//   sia_linear_splitk : Flatten + Linear(K -> N) as a split-K tcgen05 GEMM           tone_bias_model.py:100,111
//   sia_head_tail     : split-K reduce + bias + ReLU, Linear(512,256)+ReLU, Linear(256,2),
//                       LogSoftmax, argmax (+ optional fused confusion counts)        tone_bias_model.py:111-129,
//                                                                                    tone_bias_test.py:199
//
// fc1 is skinny (M = batch, N = 512, K = 100352): only (M/128)*(N/128) output tiles exist, so K is
// split across CTAs until the grid fills the GPU; each CTA streams its K-slice of A (activations,
// NHWC-flattened) and W (columns pre-permuted CHW->HWC at load) through a TMA/mbarrier ring in the
// canonical 128B-swizzled K-major layout and leaves an fp32 partial tile.  The partials are summed
// in a fixed order by the tail kernel, so the result is deterministic.
#include <math.h>

#include "sia_host.cuh"
#include "sia_ptx.cuh"

namespace sia {

constexpr int LN_BM = 128, LN_BN = 128, LN_BK = 64;
constexpr int LN_STAGE_BYTES = (LN_BM + LN_BN) * LN_BK * 2;  // 32 KB
#ifndef SIA_LN_NSTAGE
#define SIA_LN_NSTAGE 7
#endif
constexpr int LN_NSTAGE = SIA_LN_NSTAGE;
constexpr int LN_OUT_PITCH = LN_BN * 4 + 16;   // epilogue staging: rows 4 words apart modulo the 32 banks
constexpr int LN_THREADS = 192;  // warp0 TMA, warp1 MMA (+TMEM alloc), warps2-5 epilogue

// W_TILED: the weights are stored tile by tile, each 128 x 64 tile as the 16 KB shared-memory image the MMA reads
// (128-byte rows, 16-byte chunks XOR-swizzled by the row; sia_retile_linear_w), so a stage's weight tile is ONE
// contiguous bulk copy instead of 128 strided 128-byte rows.
template <bool W_TILED>
__global__ void __launch_bounds__(LN_THREADS, 1)
linear_splitk_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w,
                     const uint8_t* __restrict__ w_tiles, float* __restrict__ partial, int M, int N, int kblocks_total,
                     int splits) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + LN_NSTAGE * LN_STAGE_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + LN_NSTAGE;
  uint64_t* done_bar = bars + 2 * LN_NSTAGE;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * LN_NSTAGE + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m_blk = blockIdx.x, n_blk = blockIdx.y, split = blockIdx.z;
  // contiguous, near-equal K slices
  const int kb_lo = (int)((long long)kblocks_total * split / splits);
  const int kb_hi = (int)((long long)kblocks_total * (split + 1) / splits);

  if (threadIdx.x == 0) {
    for (int i = 0; i < LN_NSTAGE; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    mbar_init(done_bar, 1);
    fence_mbar_init();
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_w);
  }
  if (warp == 1) tmem_alloc(tmem_slot, LN_BN);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = bcast0(*tmem_slot);
  pdl_launch_dependents();          // programmatic dependent launch: the prologue above overlaps the previous kernel

  if (warp == 0) {
    if (lane == 0) {
      auto load_w = [&](int stage, int kb) {
        uint8_t* sw = smem + stage * LN_STAGE_BYTES + LN_BM * LN_BK * 2;
        if constexpr (W_TILED) {
          bulk_load_1d(sw, w_tiles + ((size_t)n_blk * kblocks_total + kb) * (LN_BN * LN_BK * 2), LN_BN * LN_BK * 2,
                       &full_bar[stage]);
        } else {
          tma_load_2d(sw, &tmap_w, &full_bar[stage], kb * LN_BK, n_blk * LN_BN);
        }
      };
      // the weights are not the predecessor's output: the weight tiles of the first stages are requested before
      // griddepcontrol.wait, the activations after it
      const int pre = min(LN_NSTAGE, kb_hi - kb_lo);
      for (int i = 0; i < pre; ++i) {
        mbar_arrive_expect_tx(&full_bar[i], LN_STAGE_BYTES);
        load_w(i, kb_lo + i);
      }
      pdl_wait();
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = kb_lo; kb < kb_hi; ++kb) {
        if (kb - kb_lo >= pre) {
          mbar_wait(&empty_bar[stage], phase ^ 1, 40);
          mbar_arrive_expect_tx(&full_bar[stage], LN_STAGE_BYTES);
          load_w(stage, kb);
        }
        tma_load_2d(smem + stage * LN_STAGE_BYTES, &tmap_a, &full_bar[stage], kb * LN_BK, m_blk * LN_BM);
        if (++stage == LN_NSTAGE) { stage = 0; phase ^= 1; }
      }
      // Drain: the last stages' tcgen05.commit arrivals on the empty barriers are asynchronous and nobody else waits
      // for them; they must have landed before the CTA exits and its shared memory goes to the next kernel's CTA.
      for (int i = 0; i < LN_NSTAGE; ++i) {
        mbar_wait(&empty_bar[stage], phase ^ 1, 43);
        if (++stage == LN_NSTAGE) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    pdl_wait();
    constexpr uint32_t idesc = make_idesc_bf16(LN_BM, LN_BN);
    constexpr uint32_t hi = desc_hi(1024, SW_128B);
    const uint32_t lo0 = desc_lo(smem_u32(smem), 0);
    int stage = 0;
    uint32_t phase = 0;
    for (int kb = kb_lo; kb < kb_hi; ++kb) {
      mbar_wait(&full_bar[stage], phase, 41);
      tc_fence_after_sync();
      if (elect_one()) {
        const uint32_t a_lo = lo0 + stage * (LN_STAGE_BYTES >> 4);
        const uint32_t b_lo = a_lo + ((LN_BM * LN_BK * 2) >> 4);
#pragma unroll
        for (int kk = 0; kk < LN_BK / 16; ++kk) {
          umma_bf16_ss_w(tmem_base, a_lo + kk * 2, hi, b_lo + kk * 2, hi, idesc, (kb > kb_lo || kk > 0) ? 1u : 0u);
        }
        umma_commit(&empty_bar[stage]);
        if (kb == kb_hi - 1) umma_commit(done_bar);
      }
      __syncwarp();
      if (++stage == LN_NSTAGE) { stage = 0; phase ^= 1; }
    }
  } else {
    // epilogue: warp w may only touch TMEM lanes 32*(w%4) ..  Each lane owns one output row (512 bytes of this tile):
    // it stages the row in the pipeline memory (idle once done_bar has completed) and sends it with ONE bulk copy,
    // instead of 16-byte pieces scattered over 32 rows per store instruction.
    const int e = warp & 3;
    const int row = m_blk * LN_BM + e * 32 + lane;
    float* mine = reinterpret_cast<float*>(smem + (e * 32 + lane) * LN_OUT_PITCH);
    pdl_wait();
    mbar_wait(done_bar, 0, 42);
    tc_fence_after_sync();
#pragma unroll 1
    for (int cb = 0; cb < LN_BN; cb += 32) {
      uint32_t v[32];
      tmem_ld32(tmem_base + ((uint32_t)(32 * e) << 16) + cb, v);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        *reinterpret_cast<float4*>(mine + cb + j) = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                                                                __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
      }
    }
    fence_proxy_async_smem();        // the generic-proxy stores above -> visible to the bulk copy engine
    if (row < M) bulk_store_1d(partial + ((size_t)split * M + row) * N + n_blk * LN_BN, mine, LN_BN * 4);
    bulk_commit_group();
    bulk_wait_group0();
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_free(tmem_base, LN_BN);
}

// ------------------------------------- tail --------------------------------------------------
constexpr int TAIL_THREADS = 512;
constexpr int TAIL_IMGS = 4;     // images per CTA (amortises the W2 read)
constexpr int TAIL_MAX_N1 = 512;
constexpr int TAIL_MAX_N2 = 256;
constexpr int TAIL_KPARTS = 8;   // fc2: K is split over 8 thread groups (64 threads x 4 outputs each)

__global__ void __launch_bounds__(TAIL_THREADS, 1)
head_tail_kernel(const float* __restrict__ partial, int splits, int M, int n1, int n2, const float* __restrict__ b1,
                 const float* __restrict__ w2t, const float* __restrict__ b2, const float* __restrict__ w3,
                 const float* __restrict__ b3, float* __restrict__ logp, uint8_t* __restrict__ pred,
                 const uint8_t* __restrict__ label, const uint8_t* __restrict__ groups, int groups_stride, int n_attr,
                 int n_groups, unsigned long long* __restrict__ counts) {
  static_assert(TAIL_IMGS == 4, "h1 is stored as one float4 per k (x, y, z, w = the CTA's four images)");
  __shared__ float4 h1v[TAIL_MAX_N1];                // [k] -> the four images: one LDS.128 per weight in fc2
  float* h1 = reinterpret_cast<float*>(h1v);         // h1[k * 4 + img]
  __shared__ __align__(16) float h2p[TAIL_KPARTS][TAIL_IMGS][TAIL_MAX_N2];   // partial sums of fc2 over the K parts
  __shared__ float z[TAIL_IMGS][2];
  const int m0 = blockIdx.x * TAIL_IMGS;
  // Programmatic dependent launch: this kernel may be resident before the split-K GEMM has finished; its partial sums
  // (and the labels / group ids another stream operation wrote) are read with ld.global.cg, i.e. from L2, never
  // through an L1 line left over from the previous batch.
  pdl_launch_dependents();
  pdl_wait();

  // h1 = relu(b1 + sum over splits, in split order): the loads of one element are independent
  for (int i = threadIdx.x; i < TAIL_IMGS * n1; i += blockDim.x) {
    const int img = i / n1, k = i % n1;
    float s = 0.f;
    if (m0 + img < M) {
      const float* src = partial + (size_t)(m0 + img) * n1 + k;
      const size_t step = (size_t)M * n1;
      // all loads of a batch are issued before the first add (one memory round trip per 16 splits); the adds stay
      // in split order, so the sum is deterministic
      int sp = 0;
      for (; sp + 16 <= splits; sp += 16) {
        float v[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) v[u] = __ldcg(src + (size_t)(sp + u) * step);
#pragma unroll
        for (int u = 0; u < 16; ++u) s += v[u];
      }
      for (; sp + 4 <= splits; sp += 4) {
        const float a = __ldcg(src + (size_t)sp * step), b = __ldcg(src + (size_t)(sp + 1) * step);
        const float c = __ldcg(src + (size_t)(sp + 2) * step), d = __ldcg(src + (size_t)(sp + 3) * step);
        s = (((s + a) + b) + c) + d;
      }
      for (; sp < splits; ++sp) s += __ldcg(src + (size_t)sp * step);
      s = fmaxf(s + b1[k], 0.f);
    }
    h1[k * TAIL_IMGS + img] = s;
  }
  __syncthreads();

  // h2 = relu(W2 h1 + b2); w2t is [n1][n2].  Fast path (n2 % 4 == 0, n2 <= 256): thread (g, q) owns the four outputs
  // 4g .. 4g+3 over the q-th eighth of K for all four images -- one 16-byte weight load per k (a warp reads 512
  // contiguous bytes) and one LDS.128 of h1 per k, all loads of an eighth in flight at once (the loop is L2-latency
  // bound).  The eight partial sums are added in a fixed order: deterministic.
  if ((n2 & 3) == 0 && n2 <= 4 * (TAIL_THREADS / TAIL_KPARTS) && (reinterpret_cast<uintptr_t>(w2t) & 15) == 0) {
    const int q = threadIdx.x / (TAIL_THREADS / TAIL_KPARTS);           // K part
    const int g = threadIdx.x % (TAIL_THREADS / TAIL_KPARTS);           // output quad
    const int k_lo = (int)((long long)n1 * q / TAIL_KPARTS), k_hi = (int)((long long)n1 * (q + 1) / TAIL_KPARTS);
    if (4 * g < n2) {
      float a[TAIL_IMGS][4];
#pragma unroll
      for (int img = 0; img < TAIL_IMGS; ++img) a[img][0] = a[img][1] = a[img][2] = a[img][3] = 0.f;
      const float4* wq = reinterpret_cast<const float4*>(w2t) + g;
      const int n2q = n2 >> 2;
      int k = k_lo;
      for (; k + 16 <= k_hi; k += 16) {
        float4 w[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) w[u] = __ldg(wq + (size_t)(k + u) * n2q);
#pragma unroll
        for (int u = 0; u < 16; ++u) {
          const float4 hv = h1v[k + u];
          const float hs[TAIL_IMGS] = {hv.x, hv.y, hv.z, hv.w};
#pragma unroll
          for (int img = 0; img < TAIL_IMGS; ++img) {
            a[img][0] = fmaf(w[u].x, hs[img], a[img][0]);
            a[img][1] = fmaf(w[u].y, hs[img], a[img][1]);
            a[img][2] = fmaf(w[u].z, hs[img], a[img][2]);
            a[img][3] = fmaf(w[u].w, hs[img], a[img][3]);
          }
        }
      }
      for (; k < k_hi; ++k) {
        const float4 w = __ldg(wq + (size_t)k * n2q);
        const float4 hv = h1v[k];
        const float hs[TAIL_IMGS] = {hv.x, hv.y, hv.z, hv.w};
#pragma unroll
        for (int img = 0; img < TAIL_IMGS; ++img) {
          a[img][0] = fmaf(w.x, hs[img], a[img][0]);
          a[img][1] = fmaf(w.y, hs[img], a[img][1]);
          a[img][2] = fmaf(w.z, hs[img], a[img][2]);
          a[img][3] = fmaf(w.w, hs[img], a[img][3]);
        }
      }
#pragma unroll
      for (int img = 0; img < TAIL_IMGS; ++img)
        *reinterpret_cast<float4*>(&h2p[q][img][4 * g]) = make_float4(a[img][0], a[img][1], a[img][2], a[img][3]);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < TAIL_IMGS * n2; i += blockDim.x) {
      const int img = i / n2, j = i % n2;
      float s2 = h2p[0][img][j];
#pragma unroll
      for (int part = 1; part < TAIL_KPARTS; ++part) s2 += h2p[part][img][j];
      h2p[0][img][j] = fmaxf(s2 + b2[j], 0.f);
    }
    __syncthreads();
  } else {
    // generic path: thread (j, half) owns output j over half of K for all images
    const int half = threadIdx.x / (TAIL_THREADS / 2);
    const int j0 = threadIdx.x % (TAIL_THREADS / 2);
    const int k_lo = half * (n1 / 2), k_hi = half == 0 ? n1 / 2 : n1;
    for (int j = j0; j < n2; j += TAIL_THREADS / 2) {
      float a[TAIL_IMGS];
#pragma unroll
      for (int img = 0; img < TAIL_IMGS; ++img) a[img] = 0.f;
      for (int k = k_lo; k < k_hi; ++k) {
        const float w = __ldg(w2t + (size_t)k * n2 + j);
        const float4 hv = h1v[k];
        a[0] = fmaf(w, hv.x, a[0]);
        a[1] = fmaf(w, hv.y, a[1]);
        a[2] = fmaf(w, hv.z, a[2]);
        a[3] = fmaf(w, hv.w, a[3]);
      }
#pragma unroll
      for (int img = 0; img < TAIL_IMGS; ++img) h2p[half][img][j] = a[img];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < TAIL_IMGS * n2; i += blockDim.x) {
      const int img = i / n2, j = i % n2;
      h2p[0][img][j] = fmaxf(h2p[0][img][j] + h2p[1][img][j] + b2[j], 0.f);
    }
    __syncthreads();
  }

  // z = W3 h2 + b3: one warp per (image, class)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp < TAIL_IMGS * 2) {
    const int img = warp >> 1, cls = warp & 1;
    float s = 0.f;
    for (int j = lane; j < n2; j += 32) s = fmaf(w3[cls * n2 + j], h2p[0][img][j], s);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) z[img][cls] = s + b3[cls];
  }
  __syncthreads();

  if (threadIdx.x < TAIL_IMGS && m0 + threadIdx.x < M) {
    const int img = threadIdx.x, m = m0 + img;
    const float z0 = z[img][0], z1 = z[img][1];
    const float mx = fmaxf(z0, z1);
    const float lse = mx + logf(expf(z0 - mx) + expf(z1 - mx));
    logp[2 * m] = z0 - lse;
    logp[2 * m + 1] = z1 - lse;
    const int pr = (z1 > z0) ? 1 : 0;   // torch.max: first maximal index on ties
    pred[m] = (uint8_t)pr;
    if (counts != nullptr) {
      const int cell = ((__ldcg(label + m) != 0) ? 2 : 0) | pr;
      for (int a = 0; a < n_attr; ++a) {
        const int g = __ldcg(groups + (size_t)a * groups_stride + m);
        if (g < n_groups) atomicAdd(&counts[(a * n_groups + g) * 4 + cell], 1ull);
      }
    }
  }
}

// fc1 weight [n][c*hw] (column = c*hw + p) -> bf16 [n][hw*c] (column = p*c_total + c)
__global__ void pack_linear_kernel(const float* __restrict__ w, int n, int c, int hw, __nv_bfloat16* __restrict__ dst) {
  const size_t total = (size_t)n * c * hw;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int ch = (int)(i % c);
    const size_t rest = i / c;
    const int p = (int)(rest % hw);
    const size_t row = rest / hw;
    dst[i] = __float2bfloat16_rn(w[(row * c + ch) * hw + p]);
  }
}

}  // namespace sia

namespace sia {
// fc1 weight [n][c*hw] -> bf16 [n_pad][hw*c_pad] (column = p*c_pad + ch), zero outside [n][c]: the activations are
// NHWC with c_pad (zero-padded) channels and the split-K GEMM wants n_pad % 128 == 0
__global__ void pack_linear_padded_kernel(const float* __restrict__ w, int n, int c, int hw, int n_pad, int c_pad,
                                          __nv_bfloat16* __restrict__ dst) {
  const size_t total = (size_t)n_pad * c_pad * hw;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int ch = (int)(i % c_pad);
    const size_t rest = i / c_pad;
    const int p = (int)(rest % hw);
    const size_t row = rest / hw;
    const float v = (row < (size_t)n && ch < c) ? w[(row * c + ch) * hw + p] : 0.f;
    dst[i] = __float2bfloat16_rn(v);
  }
}
}  // namespace sia

extern "C" int sia_pack_linear_chw_to_hwc_padded(const float* w, int n, int c, int hw, int n_pad, int c_pad,
                                                 void* packed_bf16, void* stream) {
  using namespace sia;
  SIA_REQUIRE(w && packed_bf16 && n >= 1 && c >= 1 && hw >= 1 && n_pad >= n && c_pad >= c);
  pack_linear_padded_kernel<<<sm_count() * 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      w, n, c, hw, n_pad, c_pad, static_cast<__nv_bfloat16*>(packed_bf16));
  return launch_status();
}

extern "C" int sia_pack_linear_chw_to_hwc(const float* w, int n, int c, int hw, void* packed_bf16, void* stream) {
  using namespace sia;
  SIA_REQUIRE(w && packed_bf16 && n >= 1 && c >= 1 && hw >= 1);
  pack_linear_kernel<<<sm_count() * 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      w, n, c, hw, static_cast<__nv_bfloat16*>(packed_bf16));
  return launch_status();
}

namespace sia {
// [n][k] bf16 (k contiguous) -> tiles [n/128][k/64][128 rows][8 chunks of 16 bytes], chunk c of row r at position
// c ^ (r & 7): the K-major SWIZZLE_128B image of the tile
__global__ void retile_linear_kernel(const uint4* __restrict__ src, int n, int k, uint4* __restrict__ dst) {
  const int kb_total = k / LN_BK;
  const size_t chunks = (size_t)n * k / 8;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < chunks; i += (size_t)gridDim.x * blockDim.x) {
    const int p = (int)(i & 7);
    const int r = (int)((i >> 3) & (LN_BN - 1));
    const size_t tile = i >> 10;
    const int kb = (int)(tile % kb_total);
    const size_t nb = tile / kb_total;
    const int c = p ^ (r & 7);
    dst[i] = src[((nb * LN_BN + r) * (size_t)k + (size_t)kb * LN_BK) / 8 + c];
  }
}

static int linear_splitk_launch(const void* a_bf16, const void* w_bf16, bool tiled, int m, int n, int k, int splits,
                                float* partial, void* stream) {
  SIA_REQUIRE(a_bf16 && w_bf16 && partial && m >= 1 && n >= 1 && k >= 1 && splits >= 1);
  SIA_REQUIRE(aligned(a_bf16, 16) && aligned(w_bf16, 16) && aligned(partial, 16));
  if (n % LN_BN != 0 || k % LN_BK != 0 || splits > k / LN_BK || splits > 65535) return SIA_E_UNSUPPORTED;
  if (int wrc = ensure_watchdog()) return wrc;
  CUtensorMap ta, tw;
  {
    const uint64_t dims[2] = {(uint64_t)k, (uint64_t)m};
    const uint64_t strides[1] = {(uint64_t)k * 2};
    const uint32_t box[2] = {LN_BK, LN_BM};
    int rc = encode_tmap_bf16(&ta, a_bf16, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc != 0) return rc;
  }
  if (!tiled) {
    const uint64_t dims[2] = {(uint64_t)k, (uint64_t)n};
    const uint64_t strides[1] = {(uint64_t)k * 2};
    const uint32_t box[2] = {LN_BK, LN_BN};
    int rc = encode_tmap_bf16(&tw, w_bf16, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc != 0) return rc;
  } else {
    tw = ta;      // not read by the tiled kernel
  }
  const int smem = 1024 + LN_NSTAGE * LN_STAGE_BYTES + (2 * LN_NSTAGE + 2) * 8;
  dim3 grid((m + LN_BM - 1) / LN_BM, n / LN_BN, splits);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (tiled) {
    static SmemSlots configured = {};
    if (int rc2 = ensure_dynamic_smem(linear_splitk_kernel<true>, smem, &configured)) return rc2;
    return launch_kernel(linear_splitk_kernel<true>, grid, dim3(LN_THREADS), smem, st, true, ta, tw,
                         static_cast<const uint8_t*>(w_bf16), partial, m, n, k / LN_BK, splits);
  }
  static SmemSlots configured = {};
  if (int rc2 = ensure_dynamic_smem(linear_splitk_kernel<false>, smem, &configured)) return rc2;
  return launch_kernel(linear_splitk_kernel<false>, grid, dim3(LN_THREADS), smem, st, true, ta, tw,
                       static_cast<const uint8_t*>(nullptr), partial, m, n, k / LN_BK, splits);
}
}  // namespace sia

extern "C" int sia_linear_splitk(const void* a_bf16, const void* w_bf16, int m, int n, int k, int splits,
                                 float* partial, void* stream) {
  return sia::linear_splitk_launch(a_bf16, w_bf16, false, m, n, k, splits, partial, stream);
}

extern "C" int sia_linear_splitk_tiled(const void* a_bf16, const void* w_tiles_bf16, int m, int n, int k, int splits,
                                       float* partial, void* stream) {
  return sia::linear_splitk_launch(a_bf16, w_tiles_bf16, true, m, n, k, splits, partial, stream);
}

extern "C" int sia_retile_linear_w(const void* w_bf16, int n, int k, void* w_tiles_bf16, void* stream) {
  using namespace sia;
  SIA_REQUIRE(w_bf16 && w_tiles_bf16 && w_bf16 != w_tiles_bf16 && n >= 1 && k >= 1);
  SIA_REQUIRE(aligned(w_bf16, 16) && aligned(w_tiles_bf16, 16));
  if (n % LN_BN != 0 || k % LN_BK != 0) return SIA_E_UNSUPPORTED;
  retile_linear_kernel<<<sm_count() * 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint4*>(w_bf16), n, k, static_cast<uint4*>(w_tiles_bf16));
  return launch_status();
}

extern "C" int sia_head_tail(const float* partial, int splits, int m, int n1, int n2, const float* b1,
                             const float* w2t, const float* b2, const float* w3, const float* b3, float* logp,
                             uint8_t* pred, const uint8_t* label, const uint8_t* groups, int groups_stride, int n_attr,
                             int n_groups, long long* counts, void* stream) {
  using namespace sia;
  SIA_REQUIRE(partial && b1 && w2t && b2 && w3 && b3 && logp && pred && splits >= 1 && m >= 1);
  if (n1 < 2 || n1 > TAIL_MAX_N1 || n1 % 2 != 0 || n2 < 1 || n2 > TAIL_MAX_N2) return SIA_E_UNSUPPORTED;
  if (counts != nullptr) {
    SIA_REQUIRE(label && groups && n_attr >= 1 && n_groups >= 1 && groups_stride >= m);
  }
  // the reference's head (512 -> 256 -> 2) runs on eight-CTA clusters (tail_cluster.cu); any other head here
  int crc = 0;
  if (launch_head_tail_cluster(&crc, partial, splits, m, n1, n2, b1, w2t, b2, w3, b3, logp, pred, label, groups,
                               groups_stride, n_attr, n_groups, reinterpret_cast<unsigned long long*>(counts),
                               static_cast<cudaStream_t>(stream)))
    return crc;
  return launch_kernel(head_tail_kernel, dim3((m + TAIL_IMGS - 1) / TAIL_IMGS), dim3(TAIL_THREADS), 0, static_cast<cudaStream_t>(stream), true,
      partial, splits, m, n1, n2, b1, w2t, b2, w3, b3, logp, pred, label, groups, groups_stride, n_attr, n_groups,
      reinterpret_cast<unsigned long long*>(counts));
  return launch_status();
}

// ------------------------------------- tail, any depth -----------------------------------------
// tone_bias_optuna.define_isic_model (:123-173) builds 2-5 hidden Linear layers of 16-256 units: the same tail as
// head_tail_kernel for a chain  h1 = relu(sum_s partial + b1);  h_{l+1} = relu(W_l h_l + b_l) ...;  z = W_L h_L + b_L
// (2 classes); log-softmax; argmax; optional confusion counts.  fp32 throughout; four images per CTA.
namespace sia {
constexpr int CHAIN_MAX_LAYERS = 6;
struct ChainParams {
  const float* wt[CHAIN_MAX_LAYERS];   // [n_in][n_out] (transposed Linear weights)
  const float* b[CHAIN_MAX_LAYERS];
  int n_in[CHAIN_MAX_LAYERS], n_out[CHAIN_MAX_LAYERS];
  int n_layers;                        // layers after fc1; the last one has n_out == 2
};

__global__ void __launch_bounds__(256)
tail_chain_kernel(const float* __restrict__ partial, int splits, int M, int n1, int n1_stride,
                  const float* __restrict__ b1, const __grid_constant__ ChainParams cp, float* __restrict__ logp,
                  uint8_t* __restrict__ pred, const uint8_t* __restrict__ label, const uint8_t* __restrict__ groups,
                  int groups_stride, int n_attr, int n_groups, unsigned long long* __restrict__ counts) {
  __shared__ float h[2][TAIL_IMGS][TAIL_MAX_N1];
  const int m0 = blockIdx.x * TAIL_IMGS;
  for (int i = threadIdx.x; i < TAIL_IMGS * n1; i += blockDim.x) {
    const int img = i / n1, k = i % n1;
    float s = 0.f;
    if (m0 + img < M) {
      const float* src = partial + (size_t)(m0 + img) * n1_stride + k;
      const size_t step = (size_t)M * n1_stride;
      // sixteen loads in flight per round trip; the adds stay in split order: deterministic
      for (int sp0 = 0; sp0 < splits; sp0 += 16) {
        float v[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) {
          if (sp0 + u < splits) v[u] = __ldg(src + (size_t)(sp0 + u) * step);
        }
#pragma unroll
        for (int u = 0; u < 16; ++u) {
          if (sp0 + u < splits) s += v[u];
        }
      }
      s = fmaxf(s + b1[k], 0.f);
    }
    h[0][img][k] = s;
  }
  __syncthreads();
  int cur = 0;
  for (int l = 0; l < cp.n_layers; ++l) {
    const int n_in = cp.n_in[l], n_out = cp.n_out[l];
    const bool last = l == cp.n_layers - 1;
    for (int j = threadIdx.x; j < n_out; j += blockDim.x) {
      float a[TAIL_IMGS];
#pragma unroll
      for (int img = 0; img < TAIL_IMGS; ++img) a[img] = 0.f;
      for (int k = 0; k < n_in; ++k) {
        const float w = __ldg(cp.wt[l] + (size_t)k * n_out + j);
#pragma unroll
        for (int img = 0; img < TAIL_IMGS; ++img) a[img] = fmaf(w, h[cur][img][k], a[img]);
      }
      const float bj = cp.b[l][j];
#pragma unroll
      for (int img = 0; img < TAIL_IMGS; ++img) h[cur ^ 1][img][j] = last ? a[img] + bj : fmaxf(a[img] + bj, 0.f);
    }
    __syncthreads();
    cur ^= 1;
  }
  if (threadIdx.x < TAIL_IMGS && m0 + threadIdx.x < M) {
    const int img = threadIdx.x, m = m0 + img;
    const float z0 = h[cur][img][0], z1 = h[cur][img][1];
    const float mx = fmaxf(z0, z1);
    const float lse = mx + logf(expf(z0 - mx) + expf(z1 - mx));
    logp[2 * m] = z0 - lse;
    logp[2 * m + 1] = z1 - lse;
    const int pr = (z1 > z0) ? 1 : 0;   // torch.max: first maximal index on ties
    pred[m] = (uint8_t)pr;
    if (counts != nullptr) {
      const int cell = ((label[m] != 0) ? 2 : 0) | pr;
      for (int a = 0; a < n_attr; ++a) {
        const int g = groups[(size_t)a * groups_stride + m];
        if (g < n_groups) atomicAdd(&counts[(a * n_groups + g) * 4 + cell], 1ull);
      }
    }
  }
}
}  // namespace sia

extern "C" int sia_head_tail_chain(const float* partial, int splits, int m, int n1, int n1_stride, const float* b1,
                                   int n_layers, const float* const* wt_host, const float* const* b_host,
                                   const int* n_out_host, float* logp, uint8_t* pred, const uint8_t* label,
                                   const uint8_t* groups, int groups_stride, int n_attr, int n_groups,
                                   long long* counts, void* stream) {
  using namespace sia;
  SIA_REQUIRE(partial && b1 && wt_host && b_host && n_out_host && logp && pred && splits >= 1 && m >= 1);
  if (n1 < 1 || n1 > TAIL_MAX_N1 || n1_stride < n1 || n_layers < 1 || n_layers > CHAIN_MAX_LAYERS) return SIA_E_UNSUPPORTED;
  if (counts != nullptr) {
    SIA_REQUIRE(label && groups && n_attr >= 1 && n_groups >= 1 && groups_stride >= m);
  }
  ChainParams cp;
  int n_in = n1;
  for (int l = 0; l < n_layers; ++l) {
    SIA_REQUIRE(wt_host[l] && b_host[l]);
    if (n_out_host[l] < 1 || n_out_host[l] > TAIL_MAX_N1) return SIA_E_UNSUPPORTED;
    cp.wt[l] = wt_host[l];
    cp.b[l] = b_host[l];
    cp.n_in[l] = n_in;
    cp.n_out[l] = n_out_host[l];
    n_in = n_out_host[l];
  }
  if (n_in != 2) return SIA_E_UNSUPPORTED;            // the fused tail handles exactly two classes
  cp.n_layers = n_layers;
  tail_chain_kernel<<<(m + TAIL_IMGS - 1) / TAIL_IMGS, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      partial, splits, m, n1, n1_stride, b1, cp, logp, pred, label, groups, groups_stride, n_attr, n_groups,
      reinterpret_cast<unsigned long long*>(counts));
  return launch_status();
}
