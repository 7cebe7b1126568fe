/*
 * sia_b200.h -- C ABI of the B200-native batched-evaluation path of skin-image-analysis.
 *
 * One shared library (libsia_b200.so, built from skin_image_analysis_b200/csrc with nvcc for
 * sm_100a).  Plain pointers, sizes and a cudaStream_t (passed as void*): no torch / C++ types cross
 * this boundary and no C++ exception leaves it.  Every pointer is a DEVICE pointer unless the
 * parameter name ends in `_host`.  Kernels never allocate; the caller owns every buffer.
 *
 * Each entry point replaces one library call site of the reference's evaluation path (the
 * reference is pure Python, so "the FFI it would bind" is the ctypes stub shown in INTEGRATION.md):
 *
 *   sia_preprocess_u8hwc        np.float32(img)/255            src/tone_bias_dataset.py:335
 *                               skimage.transform.resize       src/tone_bias_dataset.py:425
 *                               image.transpose((2,0,1))       src/tone_bias_dataset.py:470
 *   sia_preprocess_tc_u8hwc,    the same three call sites for the padded NHWC4 bf16 layout, with the vertical pass
 *   sia_preprocess_tc2_u8hwc    (tc) or both passes (tc2) of the resize on the tensor cores
 *   sia_preprocess_tv_u8hwc     v2.Resize + v2.ToDtype + v2.Normalize     notebooks/ToneClassifier/CNNTrialDataset.py:71-76
 *   sia_conv7x7_c3_relu_pool2   Conv2d(3,32,7,'same')+ReLU+MaxPool2d   src/tone_bias_model.py:83-92, :169-172
 *   sia_conv3x3_relu_pool2      Conv2d(C,2C,3,'same')+ReLU+MaxPool2d   src/tone_bias_model.py:83-92, :174-184
 *   sia_linear_splitk           Flatten + Linear(100352,512)           src/tone_bias_model.py:100,111
 *   sia_head_tail               ReLU, Linear(512,256)+ReLU, Linear(256,2), LogSoftmax, torch.max(.,1)
 *                                                                      src/tone_bias_model.py:111-129,
 *                                                                      src/tone_bias_test.py:199
 *   sia_confusion_counts        the per-instance Python loops of predict_with_instance /
 *                               confusion_matrix / filter / values_counts
 *                                                                      src/tone_bias_test.py:207-234, 240-289
 *
 * Return value: 0 on success; > 0 a cudaError_t; < 0 one of SIA_E_*.
 */
#ifndef SIA_B200_H
#define SIA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SIA_VERSION 100

#define SIA_E_INVALID (-1)     /* bad argument (null pointer, non-positive size, misaligned pointer)  */
#define SIA_E_UNSUPPORTED (-2) /* shape outside what the kernels were built for                       */
#define SIA_E_DRIVER (-3)      /* could not resolve / call cuTensorMapEncodeTiled                     */
#define SIA_E_WATCHDOG (-4)    /* a kernel's mbarrier watchdog fired (see sia_watchdog_status)          */

/* Output layouts of sia_preprocess_u8hwc */
#define SIA_LAYOUT_NCHW_F32 0   /* [B,3,S,S] float32  -- what ToTensor + default collate produce        */
#define SIA_LAYOUT_NCHW_BF16 1  /* [B,3,S,S] bfloat16                                                    */
#define SIA_LAYOUT_NHWC4_BF16 2 /* [B,S,S+8,4] bfloat16 -- the layout conv7x7_c3 consumes: channel 3 = 0,  */
                                /* image pixel x in column x+1, columns 0 and S+1..S+7 zero (the 8-byte  */
                                /* shift makes every TMA window start 16-byte aligned)                   */
#define SIA_NHWC4_PAD 8         /* extra pixels per row of the NHWC4 layout                              */

int sia_version(void);
const char* sia_error_string(int code);
/* Fills SM count and compute capability of the current device. */
int sia_device_info(int* sm_count_host, int* cc_major_host, int* cc_minor_host);
/* Returns the watchdog word of the current device (0 = never fired); reset != 0 clears it. */
unsigned int sia_watchdog_status(int reset);

/* ------------------------------------------------------------------------------------------
 * K1-K3  fused  u8 HWC -> (x/255) -> anti-aliased bilinear resize -> (v-mean)/std -> layout.
 *
 * The resize is the separable operator  out = Wy * img * Wx^T  where Wy/Wx are the banded
 * matrices of (ndimage.zoom order 1, grid_mode, 'mirror') o (ndimage.gaussian_filter 'mirror'),
 * i.e. skimage.transform.resize with the reference's defaults.  The caller passes the bands:
 *   x_off[out_w]            first source column of output column j's window
 *   x_w  [out_w * x_taps]   its x_taps weights (zero padded); x_taps is 8 or 16
 *   x_wq [out_w * 4]        optional (may be NULL): the same 8 weights as packed pairs of 15-bit fixed
 *                           point (w*2^15, renormalised to sum to 2^15); enables the integer-dot-product
 *                           horizontal pass for the bf16 layouts (error <= 0.03 bf16 ulp)
 *   row_w   [src_h * 4]     per SOURCE row r: weight of row r in each of the 4 accumulator slots
 *   row_emit[src_h * 4]     per SOURCE row r: output row completed in that slot after r, or -1
 *   y_first_last[out_h * 2] first and last source row that contribute to each output row
 * (skin_image_analysis_b200/resize_weights.py builds them; 1/255 is folded into row_w).
 * out = acc * out_scale[c] + out_bias[c]  (1/std and -mean/std; identity for the reference path).
 * rows_per_cta output rows are produced per thread block.
 * ------------------------------------------------------------------------------------------ */
int sia_preprocess_u8hwc(const uint8_t* src, int batch, int src_h, int src_w, const int32_t* x_off,
                         const float* x_w, const uint32_t* x_wq, int x_taps, const float* row_w,
                         const int32_t* row_emit,
                         const int32_t* y_first_last, int out_h, int out_w, const float* out_scale_host,
                         const float* out_bias_host, int layout, int rows_per_cta, void* dst, void* stream);

/* Same operator, tensor-core formulation, SIA_LAYOUT_NHWC4_BF16 output only (csrc/preprocess_tc.cu): the
 * vertical pass is a tcgen05 GEMM (fp16 copies of the Wy rows x the image bytes widened to fp16, fp32
 * accumulate in TMEM), the horizontal pass streams the accumulator columns.  Tables come from
 * resize_weights.build_tc_tables:
 *   a_packed  [n_tiles][65536]   fp16 Wy rows of each tile of tile_rows output rows, in UMMA K-major order
 *   lane_scale[n_tiles][128]     sum(w) / sum(fp16(w)) of each output row
 *   tile_row0 [n_tiles]          first source row of each tile's 256-row window
 *   items     [n_items][16]      the horizontal schedule, 64 bytes per item: int32 {accumulator column,
 *                                output column to emit or -1, column block, source pixels consumed}, then
 *                                3 pixels x 4 slots of fp32 weights
 *                                (word 3 of the first item of every group of 4: >= 0 = the group's four pixels
 *                                are the padded columns word3 .. word3+3, written as one 32-byte sector)
 *   n_blocks, last_block_cols    column blocks of 240 bytes per image row; MMA N of the last one
 *   pads_in_schedule             != 0: those groups also cover the zero pad columns of every row
 * Needs (3 * src_w) % 8 == 0, an even src_h and src 16-byte aligned (the images are fetched by TMA as double
 * rows); SIA_E_UNSUPPORTED otherwise. */
int sia_preprocess_tc_u8hwc(const uint8_t* src, int batch, int src_h, int src_w, const void* a_packed,
                            const float* lane_scale, const int32_t* tile_row0, int n_tiles, int tile_rows,
                            const void* items, int n_items, int n_blocks, int last_block_cols, int pads_in_schedule,
                            int out_h, int out_w, const float* out_scale_host, const float* out_bias_host, void* dst_nhwc4, void* stream);

/* Same operator with BOTH passes on the tensor cores (csrc/preprocess_tc2.cu): the horizontal pass is a second
 * tcgen05.mma per 120-byte column block whose A operand is the (fp16-repacked) accumulator of the vertical product, read
 * from tensor memory.  Vertical tables as for sia_preprocess_tc_u8hwc; horizontal tables from
 * resize_weights.build_tc2_tables:
 *   b2        [n_blocks][16384]   the block's 64 x 128 fp16 slice of Wx in UMMA K-major core-matrix order
 *                                 (row = 3 * pixel slot + channel, column = byte inside the block)
 *   block_meta[n_blocks][4]       int32 {s_lo, s_hi, j_lo, 0}: the block completes pixel slots [s_lo, s_hi) = output
 *                                 columns j_lo, j_lo + 1, ... (slots 17..20 are carried into slots 0..3 of the next block)
 *   slot_scale[n_blocks][21]      sum(w) / sum(fp16(w)) of that output column
 * Same geometry restrictions as the one-product kernel plus <= 4 + 13 + 4 output columns per block
 * (the builder raises otherwise).  V is rounded to fp16: <= 2.5e-4 of full scale, <= 1 bf16 ulp. */
int sia_preprocess_tc2_u8hwc(const uint8_t* src, int batch, int src_h, int src_w, const void* a_packed,
                             const float* lane_scale, const int32_t* tile_row0, int n_tiles, int tile_rows,
                             const void* b2, const int32_t* block_meta, const float* slot_scale, int n_blocks,
                             int last_block_cols, int out_h, int out_w, const float* out_scale_host,
                             const float* out_bias_host, void* dst_nhwc4, void* stream);

/* Warp-MMA variant of the same transform -- the DEFAULT for the padded NHWC4 bf16 layout (csrc/preprocess_mma.cu):
 * both banded products of the resize run on mma.sync.m16n8k16 with the operands built in registers from the raw image
 * bytes (the accumulators of the vertical product are the A fragments of the horizontal one), source rows streamed by
 * one cp.async.bulk per 8-row octet into a shared-memory ring, whole-sector stores; pad columns are written as zeros.
 * Replaces tone_bias_dataset.py:335, :411-427, :464-473 like sia_preprocess_u8hwc.  Tables:
 * resize_weights.build_mma_tables (wy_frag [n_msteps][kv][32] x 16 B A fragments; r0 [n_msteps] window start rows,
 * multiples of 8; wx_frag [n_tiles][2][2][32] x 8 B B fragments; wx_mask [n_tiles]; tile_begin [n_groups + 1]);
 * q_stride / c_row4_host [4]: which row of a 16-row chunk the K slots read (resize_weights.MMA_ROW_MAPS);
 * mul3_host [3] = 2^9 * scale / std_c, bias3_host [3] = -mean_c / std_c (host memory).  kv in {2, 3}.  Needs
 * src_w % 8 == 0, even src_h, out_w % 8 == 0, 16-byte aligned src / dst; SIA_E_UNSUPPORTED otherwise. */
int sia_preprocess_mma_u8hwc(const uint8_t* src, int batch, int src_h, int src_w, const void* wy_frag,
                             const int32_t* r0, int n_msteps, int kv, const void* wx_frag, const uint32_t* wx_mask,
                             const int32_t* tile_begin, int n_groups, int n_tiles, int q_stride,
                             const int32_t* c_row4_host, const float* mul3_host, const float* bias3_host, int out_h,
                             int out_w, void* dst, void* stream);

/* SURVEY 8(f) row 2 -- the ToneClassifier test transform (notebooks/ToneClassifier/CNNTrialDataset.py:71-76:
 * v2.Resize((224,224)) on uint8 [bilinear, antialias] -> v2.ToDtype(float32, scale=True) -> v2.Normalize(mean, std))
 * for a batch of u8 HWC decode buffers.  torchvision's uint8 resize is ATen's fixed-point separable resampler
 * (aten/src/ATen/native/cpu/UpSampleKernel.cpp): horizontal pass -> uint8 -> vertical pass -> uint8, so the
 * kernel is bit-exact.  Tables (resize_weights.build_tv_tables), all device pointers:
 *   x_min[out_w], x_w_tapmajor[x_taps][out_w] (int16, scaled by 2^x_prec)     horizontal taps
 *   y_min[out_h], y_w[out_h][y_taps]          (int16, scaled by 2^y_prec)     vertical taps
 *   lut_3x256                                 float32 value of byte b in channel c after ToDtype + Normalize
 * x_min[j] + x_taps <= src_w and y_min[i] + y_taps <= src_h for every j, i (the builder shifts windows that
 * would run off the edge and zero-pads their weights).  One CTA produces tile_rows output rows of one image from
 * at most max_window_rows source rows staged in shared memory (SIA_E_UNSUPPORTED if that exceeds 227 KB).
 * layout: SIA_LAYOUT_* as for sia_preprocess_u8hwc.  planar_chw != 0: src is [B,3,H,W] uint8 (what
 * torchvision.io.read_image / decode_jpeg produce, CNNTrialDataset.py:93) instead of [B,H,W,3]. */
int sia_preprocess_tv_u8hwc(const uint8_t* src, int batch, int src_h, int src_w, const int32_t* x_min,
                            const int16_t* x_w_tapmajor, int x_taps, int x_prec, const int32_t* y_min,
                            const int16_t* y_w, int y_taps, int y_prec, const float* lut_3x256, int out_h, int out_w,
                            int tile_rows, int max_window_rows, int layout, int planar_chw, void* dst, void* stream);

/* Model boundary: NCHW fp32 [B,3,h,w] (the tensor the reference DataLoader feeds to model(images),
 * src/tone_bias_test.py:190-196) -> padded NHWC4 bf16 [B,h,w+8,4] (SIA_LAYOUT_NHWC4_BF16). */
int sia_nchw_f32_to_nhwc4_bf16(const float* src, int batch, int h, int w, void* dst, void* stream);

/* SURVEY 8(f) row 4, the decode side: a GPU JPEG decoder (nvJPEG, e.g. torchvision.io.decode_jpeg(device="cuda"))
 * emits planar [B,3,H,W] uint8; the transform kernels read the interleaved decode buffer [B,H,W,3] that
 * skimage.io.imread produces on the reference path (src/tone_bias_dataset.py:326).  w % 4 == 0. */
int sia_chw_u8_to_hwc_u8(const uint8_t* src_chw, int batch, int h, int w, uint8_t* dst_hwc, void* stream);

/* Floor pooling on odd sizes (tone_bias_optuna.define_isic_model with n_conv_layers >= 4: 14 -> 7 -> 3 -> 1,
 * src/tone_bias_optuna.py:143-152; nn.MaxPool2d drops the last row / column): the conv kernels take even sizes, so
 * an odd-sized activation [B,h,w,C] bf16 is copied into an even-sized buffer [B,out_h,out_w,C] whose cells outside
 * the valid valid_h x valid_w corner are zero -- exactly the zero padding the next 'same' convolution would see. */
int sia_pad_nhwc_bf16(const void* in_nhwc, int batch, int h, int w, int channels, int valid_h, int valid_w,
                      void* out_nhwc, int out_h, int out_w, void* stream);

/* ------------------------------------------------------------------------------------------
 * Weight packing (one-off, at model load).  Inputs are the reference's state_dict tensors
 * (fp32, OIHW conv weights / [out,in] linear weights).
 * ------------------------------------------------------------------------------------------ */
/* conv1 [32,3,7,7] -> the 128 x 256 bf16 "2x2 output pixels per GEMM row" operand (65536 bytes). */
int sia_pack_conv7x7_c3(const float* w_oihw, void* packed, void* stream);
size_t sia_pack_conv7x7_c3_bytes(void);
/* conv [cout,cin,3,3] -> 9 taps x (cin/64 or 1) chunks of [cout][<=64] bf16, pre-swizzled. */
int sia_pack_conv3x3(const float* w_oihw, int cin, int cout, void* packed, void* stream);
size_t sia_pack_conv3x3_bytes(int cin, int cout);
/* fc1 [n, c*h*w] (CHW-flattened columns) -> bf16 [n, h*w*c] (HWC-flattened columns). */
int sia_pack_linear_chw_to_hwc(const float* w, int n, int c, int hw, void* packed_bf16, void* stream);
/* Same with zero padding to [n_pad, h*w*c_pad]: channel-padded activations, n_pad % 128 == 0 for sia_linear_splitk. */
int sia_pack_linear_chw_to_hwc_padded(const float* w, int n, int c, int hw, int n_pad, int c_pad, void* packed_bf16,
                                      void* stream);
/* conv [cout,cin,3,3] zero-padded to the buffer widths cin_pad (32 or a multiple of 64) x cout_pad (multiple of 8):
 * what the arbitrary layer widths of tone_bias_optuna.define_isic_model (src/tone_bias_optuna.py:123-173) need. */
int sia_pack_conv3x3_padded(const float* w_oihw, int cin, int cout, int cin_pad, int cout_pad, void* packed,
                            void* stream);

/* ------------------------------------------------------------------------------------------
 * K4  convolution blocks: conv + bias + ReLU + 2x2/2 max-pool, bf16 in / fp32 accumulate / bf16 out.
 *   in  : padded NHWC4 bf16 [B,h,w+8,4] (conv7x7_c3)  or NHWC bf16 [B,h,w,cin] (conv3x3)
 *   out : NHWC bf16 [B,h/2,w/2,cout]
 * h and w must be even; conv7x7_c3 needs w%16==0 (conv3x3 tiles past the right / bottom edge are masked).
 * (cin,cout) = channel counts of the buffers: (32,64), (64,128) use resident-weight kernels; any other pair with
 * cin % 64 == 0, cin <= 256 and cout in {64,128,192,256} uses the streamed-weight kernel.
 * ------------------------------------------------------------------------------------------ */
int sia_conv7x7_c3_relu_pool2(const void* in_nhwc4, int batch, int h, int w, const void* w_packed,
                              const float* bias, void* out_nhwc, void* stream);
int sia_conv3x3_relu_pool2(const void* in_nhwc, int batch, int h, int w, int cin, int cout, const void* w_packed,
                           const float* bias, void* out_nhwc, void* stream);
/* conv7x7_c3 writing its 32 output channels at channel offset c_offset of an NHWC buffer with c_stride channels per
 * pixel: first layers wider than 32 channels run as several launches over 32-channel slices of the weights. */
int sia_conv7x7_c3_relu_pool2_strided(const void* in_nhwc4, int batch, int h, int w, const void* w_packed,
                                      const float* bias, void* out_nhwc, int c_stride, int c_offset, void* stream);

/* ------------------------------------------------------------------------------------------
 * K5  split-K linear:  partial[s][m][n] = sum_{k in slice s} a[m][k] * w[n][k]   (fp32)
 *   a : bf16 [m,k] row-major (the NHWC-flattened last conv output), w : bf16 [n,k] row-major.
 *   n % 128 == 0, k % 64 == 0, 1 <= splits <= k/64.
 * ------------------------------------------------------------------------------------------ */
int sia_linear_splitk(const void* a_bf16, const void* w_bf16, int m, int n, int k, int splits, float* partial,
                      void* stream);
/* The same product with the weights stored tile by tile (sia_retile_linear_w, once at model load): every 128 x 64
 * weight tile is one contiguous 16 KB block in HBM, already in the shared-memory order the tensor cores read, so a
 * pipeline stage fetches it with ONE bulk copy instead of 128 strided 128-byte rows.  Bit-identical results. */
int sia_linear_splitk_tiled(const void* a_bf16, const void* w_tiles_bf16, int m, int n, int k, int splits,
                            float* partial, void* stream);
/* w : bf16 [n,k] row-major  ->  w_tiles : bf16, n*k elements, tiles [n/128][k/64] of 128 rows x 128 bytes with the
 * 16-byte chunks of row r XOR-swizzled by (r & 7).  n % 128 == 0, k % 64 == 0; not in place. */
int sia_retile_linear_w(const void* w_bf16, int n, int k, void* w_tiles_bf16, void* stream);

/* ------------------------------------------------------------------------------------------
 * K6 (+K7)  tail:  h1 = relu(sum_s partial + b1); h2 = relu(W2 h1 + b2); z = W3 h2 + b3;
 *           logp = log_softmax(z); pred = argmax (first maximum on ties).
 *   w2t : fp32 [n1, n2] (TRANSPOSED Linear weight), w3 : fp32 [2, n2].   n1 <= 512, n2 <= 256.
 *   If counts != NULL the per-group confusion counts of this batch are accumulated as by
 *   sia_confusion_counts (label / groups must then be non-NULL).
 * ------------------------------------------------------------------------------------------ */
int sia_head_tail(const float* partial, int splits, int m, int n1, int n2, const float* b1, const float* w2t,
                  const float* b2, const float* w3, const float* b3, float* logp, uint8_t* pred,
                  const uint8_t* label, const uint8_t* groups, int groups_stride, int n_attr, int n_groups,
                  long long* counts, void* stream);

/* The same tail for a chain of n_layers Linear layers after fc1 (tone_bias_optuna.define_isic_model builds 2-5
 * hidden layers): wt_host / b_host / n_out_host are HOST arrays of n_layers device pointers / widths, weights
 * transposed to [n_in][n_out]; ReLU after every layer but the last, whose width must be 2.  partial is
 * [splits][m][n1_stride] (n1 <= n1_stride: the split-K GEMM pads N to a multiple of 128). */
int sia_head_tail_chain(const float* partial, int splits, int m, int n1, int n1_stride, const float* b1, int n_layers,
                        const float* const* wt_host, const float* const* b_host, const int* n_out_host, float* logp,
                        uint8_t* pred, const uint8_t* label, const uint8_t* groups, int groups_stride, int n_attr,
                        int n_groups, long long* counts, void* stream);

/* ------------------------------------------------------------------------------------------
 * K7  counts[a][g][label][pred] += 1 for every instance i and attribute a with
 *     g = groups[a*groups_stride + i] < n_groups  (any other value: the instance is in no group of
 *     that attribute -- the reference's filter() semantics).  pred/label in {0,1}; 1 = 'malignant'.
 *     counts: int64 [n_attr][n_groups][2][2], accumulated (caller zeroes).  n_attr*n_groups <= 256.
 * ------------------------------------------------------------------------------------------ */
int sia_confusion_counts(const uint8_t* pred, const uint8_t* label, const uint8_t* groups, long long n,
                         long long groups_stride, int n_attr, int n_groups, long long* counts, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SIA_B200_H */
