// sm_100a building blocks shared by every kernel in this library: mbarrier, TMA (tiled tensor
// loads and 1-D bulk copies), tcgen05 (TMEM alloc, UMMA issue/commit, TMEM loads) and the
// shared-memory / instruction descriptor encoders.  Inline PTX only -- no CUTLASS.
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace sia {

// ------------------------------------------------------------------------------------------
// Watchdog: a kernel that waits too long on an mbarrier records where and traps, so a protocol
// bug surfaces as a launch error instead of a hung GPU.
// ------------------------------------------------------------------------------------------
// The library is built as ONE translation unit (libsia_unity.cu), so this is the single copy.
// It points at a word of pinned, device-mapped HOST memory: a trap kills the CUDA context, the
// host word survives and sia_watchdog_status() can still say which wait timed out.
static __device__ volatile unsigned int* g_watchdog_word = nullptr;

// Optional per-CTA role timing (sia_debug_set_stats): 8 counters per CTA, SM clock cycles.
//   0 producer: waiting for a free stage      1 MMA: waiting for a free accumulator
//   2 MMA: waiting for operands (TMA)         3 MMA: whole loop
//   4 epilogue warp 4: waiting for tfull      5 epilogue warp 4: whole loop      6 tiles done by this CTA
static __device__ unsigned long long* g_stats = nullptr;

// Optional event trace of CTA 0 (sia_debug_set_trace): clock64 stamps, 8 events x TRACE_TILES tiles.
//   0 producer: stage free   1 producer: TMA issued   2 MMA: accumulator free   3 MMA: operands landed
//   4 MMA: tile issued       5 epilogue: accumulator complete   6 epilogue: accumulator drained   7 epilogue: stored
constexpr int TRACE_TILES = 64;
static __device__ long long* g_trace = nullptr;
// Both instruments cost a global load of their switch per call site -- enough to slow the conv kernels by
// 40-70 % even when switched off -- so they only exist in builds with -DSIA_INSTRUMENT (tools/build_variant.sh).
#ifdef SIA_INSTRUMENT
__device__ __forceinline__ void trace(int local_tile, int event) {
  if (g_trace != nullptr && blockIdx.x == 0 && local_tile < TRACE_TILES) g_trace[local_tile * 8 + event] = clock64();
}

struct RoleTimer {
  unsigned long long acc = 0;
  long long t0 = 0;
  __device__ __forceinline__ void begin() { if (g_stats) t0 = clock64(); }
  __device__ __forceinline__ void end() { if (g_stats) acc += (unsigned long long)(clock64() - t0); }
  __device__ __forceinline__ void store(int slot) { if (g_stats) g_stats[blockIdx.x * 8 + slot] = acc; }
};
__device__ __forceinline__ void stats_store(int slot, unsigned long long v) {
  if (g_stats) g_stats[blockIdx.x * 8 + slot] = v;
}
#else
__device__ __forceinline__ void trace(int, int) {}
struct RoleTimer {
  __device__ __forceinline__ void begin() {}
  __device__ __forceinline__ void end() {}
  __device__ __forceinline__ void store(int) {}
};
__device__ __forceinline__ void stats_store(int, unsigned long long) {}
#endif

#ifndef SIA_WATCHDOG_SPINS
#define SIA_WATCHDOG_SPINS (1u << 24)
#endif

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ uint32_t lane_id() {
  uint32_t l;
  asm volatile("mov.u32 %0, %%laneid;" : "=r"(l));
  return l;
}

// Broadcast from lane 0: tells the compiler the value is warp-uniform (it can then live in a uniform
// register, which is what UTCHMMA / UTMALDG operands need -- otherwise ptxas emits a serialising loop).
__device__ __forceinline__ uint32_t bcast0(uint32_t v) { return __shfl_sync(0xffffffffu, v, 0); }

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// ---------------------------------- mbarrier -------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Blocks until the phase with the given parity completes.  `site` tags the wait for the watchdog.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, uint32_t site) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > SIA_WATCHDOG_SPINS) {
      if (g_watchdog_word != nullptr) {
        *g_watchdog_word = 0x80000000u | (site << 16) | (blockIdx.x & 0xffffu);
        __threadfence_system();
      }
      __trap();
    }
  }
}

// ---------------------------------- several MMA-issuing warps ---------------------------------
// The conv kernels deal their tiles round-robin to W issuing warps.  A warp that handles only every W-th tile sees
// only every W-th use of a ring slot's barrier, and an mbarrier wait is by PARITY: if that warp runs ahead of the owner
// of the use in between, a phase completed two uses ago satisfies its wait (ABA) -- it issues on stale operands, the
// arrival counts drift apart and the CTA deadlocks.  It took a delayed issuer to show (six back-to-back launches of
// conv1 under programmatic dependent launch; a spinning NCCL kernel sharing the SMs): the watchdog fired in
// whichever role starved first.  issue_gate() closes it: before waiting for tile lt, the issuer waits until tile
// lt - ring has been ISSUED by its owner -- who, by induction, saw the previous phase of the same slot complete -- so
// the slot's barrier is exactly one phase away and the parity is unambiguous.  `issued[w]` = tiles issued by warp w.
__device__ __forceinline__ void issue_gate(volatile uint32_t* issued, int lt, int ring, int n_warps, uint32_t site) {
  const int g = lt - ring;
  if (g < 0 || (g % n_warps) == (lt % n_warps)) return;      // no earlier use, or this warp's own (program order)
  const uint32_t need = (uint32_t)(g / n_warps) + 1u;
  uint32_t spins = 0;
  while (issued[g % n_warps] < need) {
    if (++spins > 8u * SIA_WATCHDOG_SPINS) {
      if (g_watchdog_word != nullptr) {
        *g_watchdog_word = 0x80000000u | (site << 16) | (blockIdx.x & 0xffffu);
        __threadfence_system();
      }
      __trap();
    }
  }
}
// After the tile's MMAs and commits have been issued (and a __syncwarp), by one lane of the issuing warp.  No fence:
// the progress word and the barriers live in the same CTA's shared memory, each side touches them in program order
// (owner: barrier waits, then this store; reader: the load above, then its barrier waits) and a barrier's phase
// only moves forward.
__device__ __forceinline__ void issue_done(volatile uint32_t* issued, int lt, int n_warps) {
  issued[lt % n_warps] = (uint32_t)(lt / n_warps) + 1u;
}

// ---------------------------------- programmatic dependent launch ---------------------------
// A kernel launched with the programmatic-stream-serialization attribute may start while its predecessor in the
// stream is still running: everything before pdl_wait() (barrier init, TMEM allocation, tensor-map prefetch, loads of
// weights that no kernel writes) overlaps the predecessor's tail; pdl_wait() returns once the predecessor grid has
// completed and its memory is visible.  pdl_launch_dependents() lets the NEXT kernel's CTAs be scheduled as soon as
// every CTA of this grid has executed it (or exited).  Both are no-ops for a normally launched kernel.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ---------------------------------- fences ---------------------------------------------------
__device__ __forceinline__ void fence_proxy_async_smem() {
  // generic-proxy writes to shared memory -> visible to the async proxy (TMA / UMMA operand reads)
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before_sync() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after_sync() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

// ---------------------------------- TMA ------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                            int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                            int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2),
      "r"(c3)
      : "memory");
}
// 1-D bulk copy global -> shared.  dst, src 16-byte aligned; bytes a multiple of 16.
__device__ __forceinline__ void bulk_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

// 1-D bulk copy shared -> global (bulk async-group).  Both 16-byte aligned; bytes a multiple of 16.  The issuing
// thread commits the group and waits: `read` = the shared-memory source may be overwritten; `all` = complete.
__device__ __forceinline__ void bulk_store_1d(void* gdst, const void* ssrc, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
               ::"l"(reinterpret_cast<uint64_t>(gdst)), "r"(smem_u32(ssrc)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_group_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_group0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// ---------------------------------- TMEM -----------------------------------------------------
// ncols: power of two in [32, 512].  Must be executed by one full warp; the same warp frees.
__device__ __forceinline__ void tmem_alloc(uint32_t* slot_in_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot_in_smem)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_free(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// D[tmem] (+)= A[smem desc] * B[smem desc]^T ; one thread issues for the CTA.
__device__ __forceinline__ void umma_bf16_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Same, with the descriptors as (lo, hi) words: lo = (addr>>4) | (LBO>>4)<<16 advances by plain
// 32-bit adds (byte offset >> 4), hi = (SBO>>4) | version | layout is a compile-time constant.
__device__ __forceinline__ void umma_bf16_ss_w(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo,
                                               uint32_t b_hi, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
      ::"r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Integer variant: u8/s8 x u8/s8 -> s32 (K = 32 per instruction).
__device__ __forceinline__ void umma_i8_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_i8_ss_w(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo,
                                             uint32_t b_hi, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], da, db, %5, p;\n\t}"
      ::"r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
// All previously issued UMMAs of this thread arrive on `bar` when they complete.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// 32 lanes x 32 consecutive fp32 columns: thread t of the warp gets lane (taddr.lane + t).
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---------------------------------- descriptors ----------------------------------------------
enum : uint32_t { SW_NONE = 0, SW_128B = 2, SW_64B = 4, SW_32B = 6 };

// Shared-memory matrix descriptor (sm_100 "version 1").  Byte quantities; 16-byte granular.
//   K-major, no swizzle : 8x16B core matrices; LBO = step between core matrices along K,
//                         SBO = step between 8-row groups along M/N.
//   K-major, swizzled   : rows of 32/64/128 B; SBO = step between 8-row groups; LBO unused.
__host__ __device__ constexpr uint64_t make_smem_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout,
                                                      uint32_t base_offset = 0) {
  return (uint64_t)((addr >> 4) & 0x3fffu) | ((uint64_t)((lbo >> 4) & 0x3fffu) << 16) |
         ((uint64_t)((sbo >> 4) & 0x3fffu) << 32) | ((uint64_t)1 << 46) | ((uint64_t)(base_offset & 7u) << 49) |
         ((uint64_t)(layout & 7u) << 61);
}

__host__ __device__ constexpr uint32_t desc_lo(uint32_t addr, uint32_t lbo) {
  return ((addr >> 4) & 0x3fffu) | (((lbo >> 4) & 0x3fffu) << 16);
}
__host__ __device__ constexpr uint32_t desc_hi(uint32_t sbo, uint32_t layout) {
  return ((sbo >> 4) & 0x3fffu) | (1u << 14) | ((layout & 7u) << 29);
}

// Instruction descriptor, kind::f16: bf16 x bf16 -> fp32, both operands K-major.
__host__ __device__ constexpr uint32_t make_idesc_bf16(uint32_t m, uint32_t n) {
  return (1u << 4)            // D format  = f32
         | (1u << 7)          // A format  = bf16
         | (1u << 10)         // B format  = bf16
         | ((n >> 3) << 17)   // N / 8
         | ((m >> 4) << 24);  // M / 16
}

// Instruction descriptor, kind::i8: 8-bit integers -> s32.  *_signed: 0 = u8, 1 = s8;  *_mn_major: 0 = K-major
// operand, 1 = MN-major operand (the M/N index is the contiguous one in shared memory).
__host__ __device__ constexpr uint32_t make_idesc_i8(uint32_t m, uint32_t n, uint32_t a_signed, uint32_t b_signed,
                                                     uint32_t a_mn_major, uint32_t b_mn_major) {
  return (2u << 4)               // D format = s32
         | (a_signed << 7) | (b_signed << 10) | (a_mn_major << 15) | (b_mn_major << 16) | ((n >> 3) << 17) |
         ((m >> 4) << 24);
}

// ---------------------------------- clusters / distributed shared memory ---------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
// Address of the same shared-memory location in CTA `rank` of this cluster (shared::cluster window).
__device__ __forceinline__ uint32_t mapa_shared(uint32_t cta_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(cta_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_cluster_v2(uint32_t cluster_addr, float a, float b) {
  asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(cluster_addr), "f"(a), "f"(b) : "memory");
}
// Asynchronous remote store: 16 / 8 bytes into another CTA's shared memory, completing `bytes` on an mbarrier of THAT
// CTA (both addresses in the shared::cluster window of the same CTA).  The receiver waits on its own barrier with
// mbar_wait_cluster: no fence, no cluster-wide barrier on the data path.
__device__ __forceinline__ void st_async_v4(uint32_t cluster_addr, uint32_t cluster_bar, float a, float b, float c,
                                            float d) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.f32 [%0], {%1, %2, %3, %4}, [%5];"
               ::"r"(cluster_addr), "f"(a), "f"(b), "f"(c), "f"(d), "r"(cluster_bar)
               : "memory");
}
__device__ __forceinline__ void st_async_v2(uint32_t cluster_addr, uint32_t cluster_bar, float a, float b) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.f32 [%0], {%1, %2}, [%3];"
               ::"r"(cluster_addr), "f"(a), "f"(b), "r"(cluster_bar)
               : "memory");
}
// mbar_wait for a barrier completed by other CTAs of the cluster (acquire at cluster scope).
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity, uint32_t site) {
  uint32_t spins = 0;
  for (;;) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (ok) break;
    if (++spins > SIA_WATCHDOG_SPINS) {
      if (g_watchdog_word != nullptr) {
        *g_watchdog_word = 0x80000000u | (site << 16) | (blockIdx.x & 0xffffu);
        __threadfence_system();
      }
      __trap();
    }
  }
}
// The two halves of barrier.cluster: every thread of every CTA of the cluster arrives, then waits.  The release /
// acquire pair orders the distributed-shared-memory stores before the arrive ahead of the loads after the wait.
__device__ __forceinline__ void cluster_arrive_release() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
}
__device__ __forceinline__ void cluster_wait_acquire() {
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// ---------------------------------- cp.async (LDGSTS) ----------------------------------------
__device__ __forceinline__ void cp_async_16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

// ---------------------------------- small math helpers ---------------------------------------
// One 256-bit global store (sm_100: STG.256): a whole 32-byte sector per lane in ONE request, instead of two 16-byte
// halves that reach L2 as partial-sector writes.  p must be 32-byte aligned.
__device__ __forceinline__ void st_global_256(void* p, uint4 a, uint4 b) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"l"(p), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w)
               : "memory");
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ uint32_t max_bf16x2(uint32_t a, uint32_t b) {
  __nv_bfloat162 x = *reinterpret_cast<__nv_bfloat162*>(&a);
  __nv_bfloat162 y = *reinterpret_cast<__nv_bfloat162*>(&b);
  __nv_bfloat162 r = __hmax2(x, y);
  return *reinterpret_cast<uint32_t*>(&r);
}

}  // namespace sia
