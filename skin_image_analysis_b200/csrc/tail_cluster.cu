// K6 on thread-block clusters: the same computation as head_tail_kernel (linear.cu) for the reference's head
// dimensions, Linear(512,256) after the split-K sums and Linear(256,2) after that                tone_bias_model.py:111-129,
//                                                                                               tone_bias_test.py:199
//
// head_tail_kernel gives four images to a CTA, and every CTA reads all of W2 (512 KB fp32) out of L2: its time is the
// per-SM ingest of that matrix plus a chain of L2 round trips (profiles/r02_ncu_full_summary.txt: 22.8 us).  Here a
// cluster of eight CTAs owns 8J images (J = 1..4, chosen by the host so that ALL clusters are resident at once: a
// B200 holds 15 clusters of eight such CTAs, so batch 256 runs as 11 clusters of 24 images) and each CTA owns ONE
// EIGHTH of every weight matrix:
//
//   before griddepcontrol.wait   CTA r copies W2[:, 32r .. 32r+32) (64 KB) into shared memory with cp.async -- weights
//                                are not produced by the split-K GEMM, so this runs under the GEMM's tail wave;
//   phase 1                      CTA r sums the split-K partials of k in [64r, 64r+64) for the cluster's images (all
//                                loads of a thread in flight at once, adds in split order), adds b1, ReLU, and sends the
//                                slice into the h1 buffer of ALL eight CTAs with st.async: each 16-byte remote store
//                                completes its bytes on an mbarrier of the receiving CTA, which simply waits for the
//                                whole h1 (no fence, no cluster-wide barrier on the data path);
//   phase 2                      h2[:, 32r .. 32r+32) = relu(W2 slice . h1 + b2) entirely from shared memory
//                                (thread = J images x 4 outputs x one eighth of K; the eight K parts are added in a
//                                fixed order);
//   phase 3                      the CTA's share of z = W3 h2 (its 32 columns) goes to CTA 0 of the cluster the same
//                                way; CTA 0 adds the eight shares in rank order, b3, log-softmax, argmax, counts.
//
// Every sum has a fixed order: the result is deterministic (and equal between GPUs), as the sharded evaluation
// requires.  fp32 throughout.
#include <math.h>

#include "sia_host.cuh"
#include "sia_ptx.cuh"

namespace sia {

constexpr int TCL_CLUSTER = 8;
constexpr int TCL_MAX_J = 4;                      // images per cluster = 8 J <= 32
constexpr int TCL_THREADS = 512;
constexpr int TCL_N1 = 512, TCL_N2 = 256;
constexpr int TCL_KS = TCL_N1 / TCL_CLUSTER;      // 64 columns of h1 reduced by one CTA
constexpr int TCL_NO = TCL_N2 / TCL_CLUSTER;      // 32 outputs of fc2 owned by one CTA
constexpr int TCL_KPARTS = 8;
constexpr int TCL_H1_STRIDE = TCL_N1 + 4;         // rows 4 words apart modulo the 32 banks: the 8 image rows a warp
                                                  // reads with one LDS.128 each fall into disjoint bank groups
constexpr int TCL_H2_STRIDE = TCL_NO + 4;         // same trick for the 16-byte stores of the fc2 partial sums
constexpr int TCL_LOAD_BATCH = 18;                // split-K slices a thread keeps in flight (16-byte loads)

struct TclSmem {
  float w2s[TCL_N1][TCL_NO];                              // 64 KB   W2 slice, [k][local output]
  float h1s[8 * TCL_MAX_J][TCL_H1_STRIDE];                // 65 KB   relu(fc1) of the cluster's images, all 512 columns
  float h2p[TCL_KPARTS][8 * TCL_MAX_J][TCL_H2_STRIDE];    // 36 KB   fc2 partial sums
  float zp[TCL_CLUSTER][8 * TCL_MAX_J][2];                // (CTA 0) the eight shares of the logits
  uint64_t h1_bar, z_bar;                                 // transaction barriers the remote stores complete on
};

template <int J>
__global__ void __cluster_dims__(TCL_CLUSTER, 1, 1) __launch_bounds__(TCL_THREADS, 1)
head_tail_cluster_kernel(const float* __restrict__ partial, int splits, int M, const float* __restrict__ b1,
                         const float* __restrict__ w2t, const float* __restrict__ b2, const float* __restrict__ w3,
                         const float* __restrict__ b3, float* __restrict__ logp, uint8_t* __restrict__ pred,
                         const uint8_t* __restrict__ label, const uint8_t* __restrict__ groups, int groups_stride,
                         int n_attr, int n_groups, unsigned long long* __restrict__ counts) {
  constexpr int IMGS = 8 * J;
  extern __shared__ __align__(16) uint8_t tcl_smem_raw[];
  TclSmem& s = *reinterpret_cast<TclSmem*>(tcl_smem_raw);
  const int tid = threadIdx.x;
  const uint32_t rank = cluster_ctarank();
  const int m0 = (blockIdx.x / TCL_CLUSTER) * IMGS;

  pdl_launch_dependents();
  if (tid == 0) {
    mbar_init(&s.h1_bar, 1);
    mbar_init(&s.z_bar, 1);
    fence_mbar_init();
    mbar_arrive_expect_tx(&s.h1_bar, IMGS * TCL_N1 * 4);                      // the whole h1, from all eight CTAs
    if (rank == 0) mbar_arrive_expect_tx(&s.z_bar, TCL_CLUSTER * IMGS * 8);   // eight shares of two logits per image
  }
  __syncwarp();
  cluster_arrive_release();        // matched by the wait ahead of the first remote store: every CTA of the cluster
                                   // runs and its barriers exist (nothing is outstanding yet: the fence is cheap)

  // W2 slice: 512 rows of 128 bytes.  A warp copies four whole rows per pass (conflict-free, coalesced).
  {
    const int c = tid & 7;
#pragma unroll
    for (int pass = 0; pass < TCL_N1 / (TCL_THREADS / 8); ++pass) {
      const int k = pass * (TCL_THREADS / 8) + (tid >> 3);
      cp_async_16(&s.w2s[k][4 * c], w2t + (size_t)k * TCL_N2 + TCL_NO * rank + 4 * c);
    }
  }
  // phase 1 roles: thread = (image, column quad) of the CTA's k slice
  const int img1 = tid >> 4;
  const int kk = TCL_KS * (int)rank + 4 * (tid & 15);
  const float4 bias1 = __ldg(reinterpret_cast<const float4*>(b1 + kk));
  // phase 2 / 3 constants
  const float bias2 = __ldg(b2 + TCL_NO * rank + (tid & 31));
  const float w3a = __ldg(w3 + TCL_NO * rank + (tid & 31));
  const float w3b = __ldg(w3 + TCL_N2 + TCL_NO * rank + (tid & 31));

  pdl_wait();                      // the split-K partials (and labels / groups) are the predecessor's output

  float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
  if (img1 < IMGS && m0 + img1 < M) {
    const float* src = partial + (size_t)(m0 + img1) * TCL_N1 + kk;
    const size_t step = (size_t)M * TCL_N1;
    for (int sp0 = 0; sp0 < splits; sp0 += TCL_LOAD_BATCH) {
      float4 v[TCL_LOAD_BATCH];
#pragma unroll
      for (int u = 0; u < TCL_LOAD_BATCH; ++u) {
        if (sp0 + u < splits) v[u] = __ldcg(reinterpret_cast<const float4*>(src + (size_t)(sp0 + u) * step));
      }
#pragma unroll
      for (int u = 0; u < TCL_LOAD_BATCH; ++u) {
        if (sp0 + u < splits) {    // split order: deterministic
          sum.x += v[u].x;
          sum.y += v[u].y;
          sum.z += v[u].z;
          sum.w += v[u].w;
        }
      }
    }
    sum.x = fmaxf(sum.x + bias1.x, 0.f);
    sum.y = fmaxf(sum.y + bias1.y, 0.f);
    sum.z = fmaxf(sum.z + bias1.z, 0.f);
    sum.w = fmaxf(sum.w + bias1.w, 0.f);
  }
  cluster_wait_acquire();
  if (img1 < IMGS) {
    const uint32_t local = smem_u32(&s.h1s[img1][kk]);
    const uint32_t bar = smem_u32(&s.h1_bar);
#pragma unroll
    for (uint32_t d = 0; d < TCL_CLUSTER; ++d)
      st_async_v4(mapa_shared(local, d), mapa_shared(bar, d), sum.x, sum.y, sum.z, sum.w);
  }
  cp_async_wait_all();
  __syncthreads();                 // every thread's part of the W2 slice has landed
  mbar_wait_cluster(&s.h1_bar, 0, 45);

  // phase 2: thread (kp, og, ig) = outputs 4og..4og+3 of images ig, ig+8, .. over k in [64 kp, 64 kp + 64)
  {
    const int kp = tid >> 6, og = (tid >> 3) & 7, ig = tid & 7;
    float a[J][4];
#pragma unroll
    for (int j = 0; j < J; ++j) a[j][0] = a[j][1] = a[j][2] = a[j][3] = 0.f;
    const float* hp = &s.h1s[ig][kp * (TCL_N1 / TCL_KPARTS)];
    const float* wp = &s.w2s[kp * (TCL_N1 / TCL_KPARTS)][4 * og];
#pragma unroll 2
    for (int k = 0; k < TCL_N1 / TCL_KPARTS; k += 4) {
      float hs[J][4];
#pragma unroll
      for (int j = 0; j < J; ++j) {
        const float4 h = *reinterpret_cast<const float4*>(hp + (size_t)j * 8 * TCL_H1_STRIDE + k);
        hs[j][0] = h.x; hs[j][1] = h.y; hs[j][2] = h.z; hs[j][3] = h.w;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float4 w = *reinterpret_cast<const float4*>(wp + (size_t)(k + u) * TCL_NO);
#pragma unroll
        for (int j = 0; j < J; ++j) {
          a[j][0] = fmaf(w.x, hs[j][u], a[j][0]);
          a[j][1] = fmaf(w.y, hs[j][u], a[j][1]);
          a[j][2] = fmaf(w.z, hs[j][u], a[j][2]);
          a[j][3] = fmaf(w.w, hs[j][u], a[j][3]);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < J; ++j)
      *reinterpret_cast<float4*>(&s.h2p[kp][ig + 8 * j][4 * og]) = make_float4(a[j][0], a[j][1], a[j][2], a[j][3]);
  }
  __syncthreads();

  // h2 of the slice (warp = one image, lane = local output) and its share of the two logits, sent to CTA 0
  {
    const int j = tid & 31;
    const uint32_t zbar0 = mapa_shared(smem_u32(&s.z_bar), 0);
#pragma unroll
    for (int img = tid >> 5; img < IMGS; img += TCL_THREADS / 32) {
      float h2 = s.h2p[0][img][j];
#pragma unroll
      for (int part = 1; part < TCL_KPARTS; ++part) h2 += s.h2p[part][img][j];
      h2 = fmaxf(h2 + bias2, 0.f);
      float z0 = w3a * h2, z1 = w3b * h2;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        z0 += __shfl_xor_sync(0xffffffffu, z0, o);
        z1 += __shfl_xor_sync(0xffffffffu, z1, o);
      }
      if (j == 0) st_async_v2(mapa_shared(smem_u32(&s.zp[rank][img][0]), 0), zbar0, z0, z1);
    }
  }
  if (rank != 0) return;           // nothing is sent to a CTA other than 0 after its h1 barrier completed

  mbar_wait_cluster(&s.z_bar, 0, 46);
  if (tid < IMGS && m0 + tid < M) {
    const int m = m0 + tid;
    float z0 = s.zp[0][tid][0], z1 = s.zp[0][tid][1];
#pragma unroll
    for (int r = 1; r < TCL_CLUSTER; ++r) {
      z0 += s.zp[r][tid][0];
      z1 += s.zp[r][tid][1];
    }
    z0 += __ldg(b3);
    z1 += __ldg(b3 + 1);
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

static int g_tail_impl = -1;     // 0 = clusters where the dimensions allow, 1 = always head_tail_kernel (A/B);
                                 // -1 = not decided yet (SIA_TAIL_IMPL=1 in the environment selects 1)
static int tail_impl() {
  if (g_tail_impl < 0) {
    const char* env = getenv("SIA_TAIL_IMPL");
    g_tail_impl = (env != nullptr && env[0] == '1') ? 1 : 0;
  }
  return g_tail_impl;
}

template <int J>
static int launch_tail_cluster_j(int clusters, cudaStream_t st, const float* partial, int splits, int m,
                                 const float* b1, const float* w2t, const float* b2, const float* w3, const float* b3,
                                 float* logp, uint8_t* pred, const uint8_t* label, const uint8_t* groups,
                                 int groups_stride, int n_attr, int n_groups, unsigned long long* counts) {
  static SmemSlots configured = {};
  if (int rc = ensure_dynamic_smem(head_tail_cluster_kernel<J>, (int)sizeof(TclSmem), &configured)) return rc;
  return launch_kernel(head_tail_cluster_kernel<J>, dim3(clusters * TCL_CLUSTER), dim3(TCL_THREADS), sizeof(TclSmem), st,
                       true, partial, splits, m, b1, w2t, b2, w3, b3, logp, pred, label, groups, groups_stride, n_attr,
                       n_groups, counts);
}

// Clusters of eight of these CTAs the device holds at once (one CTA per SM by shared memory; the CTAs of a cluster
// share a GPC).  Asked once per device; -1 = none.
static int max_resident_tail_clusters() {
  static int cached[kMaxDevices] = {};
  int& have = cached[current_device()];
  if (have == 0) {
    static SmemSlots configured = {};
    if (ensure_dynamic_smem(head_tail_cluster_kernel<1>, (int)sizeof(TclSmem), &configured) != 0) return -1;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(TCL_CLUSTER * 64);
    cfg.blockDim = dim3(TCL_THREADS);
    cfg.dynamicSmemBytes = sizeof(TclSmem);
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, head_tail_cluster_kernel<1>, &cfg) != cudaSuccess || n < 1) {
      cudaGetLastError();
      n = -1;                       // no room for a cluster of eight (a partitioned GPU): head_tail_kernel runs instead
    }
    have = n;
  }
  return have;
}

// true (and *rc set) when the cluster kernel took the call
inline bool launch_head_tail_cluster(int* rc, const float* partial, int splits, int m, int n1, int n2, const float* b1,
                                     const float* w2t, const float* b2, const float* w3, const float* b3, float* logp,
                                     uint8_t* pred, const uint8_t* label, const uint8_t* groups, int groups_stride,
                                     int n_attr, int n_groups, unsigned long long* counts, cudaStream_t st) {
  if (tail_impl() == 1 || n1 != TCL_N1 || n2 != TCL_N2) return false;
  if (!aligned(partial, 16) || !aligned(w2t, 16) || !aligned(b1, 16)) return false;
  if ((*rc = ensure_watchdog()) != 0) return true;
  // fewest images per cluster (8 J) with which every cluster is resident at once; beyond 32 images per resident
  // cluster the grid simply takes several waves
  const int resident = max_resident_tail_clusters();
  if (resident < 1) return false;
  int j = 1;
  while (j < TCL_MAX_J && (m + 8 * j - 1) / (8 * j) > resident) ++j;
  const int clusters = (m + 8 * j - 1) / (8 * j);
#define SIA_TAIL_J(JJ)                                                                                              \
  case JJ:                                                                                                          \
    *rc = launch_tail_cluster_j<JJ>(clusters, st, partial, splits, m, b1, w2t, b2, w3, b3, logp, pred, label,       \
                                    groups, groups_stride, n_attr, n_groups, counts);                               \
    break;
  switch (j) {
    SIA_TAIL_J(1)
    SIA_TAIL_J(2)
    SIA_TAIL_J(3)
    SIA_TAIL_J(4)
  }
#undef SIA_TAIL_J
  return true;
}

}  // namespace sia

extern "C" int sia_debug_set_tail_impl(int impl) {
  if (impl != 0 && impl != 1) return SIA_E_INVALID;
  sia::g_tail_impl = impl;
  return 0;
}
