/* sia_b200_debug.h -- bring-up probes and instrumentation switches; NOT part of the product boundary.
 *
 * Two groups:
 *   (1) switches exported by the product library libsia_b200.so itself, because they toggle instrumentation or
 *       A/B paths inside the product kernels: sia_debug_tv_force_generic, sia_debug_set_trace, sia_debug_set_stats,
 *       sia_debug_set_mma_warps, sia_debug_set_programmatic_launch, sia_debug_set_tail_impl;
 *   (2) hardware probes (tcgen05 / TMA / TMEM / ALU micro-benchmarks used by tests/test_umma_probe.py to pin the
 *       descriptor conventions the kernels rely on), built into a SEPARATE library libsia_b200_debug.so
 *       (csrc/libsia_debug_unity.cu): the product library carries none of them.
 * A reference-side binding never needs this header.
 */
#ifndef SIA_B200_DEBUG_H
#define SIA_B200_DEBUG_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- (1) exported by libsia_b200.so ------------------------------------------------------------------ */

/* Debug / A-B: on != 0 makes sia_preprocess_tv_u8hwc use its byte-wise kernel even where the IDP.2A instance
 * applies (the two must agree bit for bit; tests compare them). */
int sia_debug_tv_force_generic(int on);

/* Debug: event trace of CTA 0 of the conv kernels: 64 tiles x 8 int64 clock64 stamps (producer stage-free /
 * TMA-issued, MMA accumulator-free / operands-landed / tile-issued, epilogue accumulator-complete / drained /
 * stored); NULL switches it off (default). */
int sia_debug_set_trace(long long* device_buffer_or_null);

/* Debug: per-CTA role timing of the conv kernels.  device_buffer holds 8 uint64 per CTA (SM-clock cycles:
 * producer wait-for-stage, MMA wait-for-accumulator, MMA wait-for-operands, MMA loop, epilogue
 * wait-for-accumulator, epilogue loop, tiles); NULL switches the instrumentation off (default). */
int sia_debug_set_stats(unsigned long long* device_buffer_or_null);

/* Debug / A-B (timing): programmatic dependent launch of the hot-path kernels (default on; SIA_PDL=0 in the
 * environment switches it off as well).  With it, every kernel of the step lets the next one start its prologue
 * (barrier init, tensor-memory allocation, resident weights) under its own tail wave. */
int sia_debug_set_programmatic_launch(int on);

/* Debug / A-B (timing): compute warps per CTA of sia_preprocess_mma_u8hwc: 4 or 8 (default 8). */
int sia_debug_set_mma_warps(int warps);

/* Debug / A-B (timing, parity): which kernel sia_head_tail launches for the 512 -> 256 -> 2 head: 0 = the
 * eight-CTA cluster kernel (default), 1 = the one-CTA-per-four-images kernel every other head uses. */
int sia_debug_set_tail_impl(int impl);

/* ---- (2) exported by libsia_b200_debug.so ------------------------------------------------------------ */

/* ------------------------------------------------------------------------------------------
 * Debug / bring-up: run a list of tcgen05.mma (kind::f16, bf16 inputs) over a caller-supplied
 * shared-memory image and return the 128 x n fp32 accumulator.  Descriptor start addresses are
 * relative to the (1024-byte aligned) image base.  Used by tests/test_umma_probe.py to pin the
 * descriptor conventions the conv kernels rely on; cycles_host (optional) gets the SM-clock
 * cycles of `repeat` back-to-back issues of the list.
 * ------------------------------------------------------------------------------------------ */
int sia_debug_umma_probe(const void* smem_image, int image_bytes, const uint64_t* a_desc_host,
                         const uint64_t* b_desc_host, int n_mma, int n, float* out_128xn, int repeat,
                         long long* cycles_host, void* stream);

/* Same, with the MMA kind (0 = kind::f16 with bf16 inputs, 1 = kind::i8) and the 32-bit instruction
 * descriptor given by the caller (0 = default bf16 K-major); the accumulator comes back as raw 32-bit words
 * (fp32 for kind 0, int32 for kind 1). */
int sia_debug_umma_probe_ex(const void* smem_image, int image_bytes, const uint64_t* a_desc_host,
                            const uint64_t* b_desc_host, int n_mma, int n, int kind, uint32_t idesc,
                            void* out_128xn_raw, int repeat, long long* cycles_host, void* stream);

/* Debug (timing only): the following probes move to the other of two accumulators every switch_every MMAs
 * (0 = off), optionally with a tcgen05.commit at every switch: the cost of short accumulation chains. */
int sia_debug_umma_probe_switch(int switch_every, int commit_each);

/* Debug / bring-up: tcgen05.mma with the A operand in tensor memory.  a_words [128][a_cols] uint32 (row m = the
 * 32-bit words thread m stores to TMEM columns 0 .. a_cols-1 of lane m); UMMA i reads A at TMEM column a_col_step * i
 * and B through b_desc_host[i] (relative to the shared-memory image, as for sia_debug_umma_probe); the 128 x n fp32
 * accumulator is returned.  idesc 0 = bf16 K-major (128, n). */
int sia_debug_umma_ts_probe(const void* smem_image, int image_bytes, const void* a_words, int a_cols, int a_col_step,
                            const uint64_t* b_desc_host, int n_mma, int n, uint32_t idesc, float* out_128xn,
                            void* stream);

/* Debug / bring-up: one TMA tiled load of a bf16 tensor (rank 2..4; dims / box in elements, innermost
 * first; strides in bytes for dims 1..rank-1; swizzle_bytes in {0,32,64,128}) at the given coordinates;
 * `out` receives the box bytes exactly as they landed in shared memory.  repeat > 1 issues that many
 * loads back to back (coordinate step_dim advanced by step each time) and reports the SM-clock cycles
 * until all have landed in cycles_host: the TMA engine's throughput for that box shape. */
int sia_debug_tma_probe(const void* base, int rank, const uint64_t* dims_host, const uint64_t* strides_bytes_host,
                        const uint32_t* box_host, int swizzle_bytes, const int* coords_host, void* out,
                        int repeat, int step_dim, int step, long long* cycles_host, void* stream);

/* Debug / bring-up: lane-operations per SM clock (one resident CTA of 1024 threads) for
 * FFMA, PRMT, I2F.U8(+IADD), DP4A, DP2A, IMAD, SHF, FFMA2 -- out_host needs room for 8 doubles. */
int sia_debug_alu_rates(double* out_host, int n);

/* Debug / bring-up: TMEM read rate in bytes per SM clock of back-to-back tcgen05.ld.32x32b for
 * (warps, columns per load) = (1,32) (4,32) (8,32) (4,8) (8,8) (4,1) -- out_host needs room for 6 doubles. */
int sia_debug_tmem_ld_rates(double* out_host, int n);

#ifdef __cplusplus
}
#endif
#endif /* SIA_B200_DEBUG_H */
