// TORCH_LIBRARY registration of the hot-path operators: a thin C++ shim over the C ABI of include/sia_b200.h
// (libsia_b200.so).  Every op TORCH_CHECKs device / dtype / contiguity / shape, takes the current CUDA stream of the
// tensor's device and calls the same extern "C" entry point the ctypes binding (skin_image_analysis_b200/_lib.py) calls
// -- so `torch.ops.sia_b200.*` and the ctypes path are the same kernels, bit for bit.  There is no CPU implementation:
// a CPU tensor fails the TORCH_CHECK.
//
// Reference operators replaced (file:line in /root/reference/src): Conv2d+ReLU+MaxPool2d blocks tone_bias_model.py:83-92,
// :169-184; Flatten+Linear :100-115; Linear/LogSoftmax + torch.max :126-129, tone_bias_test.py:199; the per-instance
// counting loops tone_bias_test.py:207-289; the Rescale / ToTensor transform tone_bias_dataset.py:335, :411-427, :464-473.
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <torch/library.h>
#include <torch/types.h>

#include <tuple>
#include <vector>

#include "sia_b200.h"

namespace {

void check_cuda(const at::Tensor& t, at::ScalarType dtype, const char* name) {
  TORCH_CHECK(t.defined(), name, ": undefined tensor");
  TORCH_CHECK(t.is_cuda(), name, ": expected a CUDA tensor (the sm_100a path has no CPU fallback)");
  TORCH_CHECK(t.scalar_type() == dtype, name, ": expected ", dtype, ", got ", t.scalar_type());
  TORCH_CHECK(t.is_contiguous(), name, ": expected a contiguous tensor");
}

void check_rc(int rc, const char* what) {
  TORCH_CHECK(rc == 0, what, " failed: ", sia_error_string(rc), " (code ", rc, ")");
}

void* stream_of(const at::Tensor& t) { return at::cuda::getCurrentCUDAStream(t.get_device()).stream(); }

at::Tensor nchw_f32_to_nhwc4(const at::Tensor& x) {
  check_cuda(x, at::kFloat, "x");
  TORCH_CHECK(x.dim() == 4 && x.size(1) == 3, "x must be [B,3,H,W]");
  c10::cuda::CUDAGuard guard(x.device());
  auto out = at::empty({x.size(0), x.size(2), x.size(3) + SIA_NHWC4_PAD, 4}, x.options().dtype(at::kBFloat16));
  check_rc(sia_nchw_f32_to_nhwc4_bf16(x.data_ptr<float>(), (int)x.size(0), (int)x.size(2), (int)x.size(3),
                                       out.data_ptr(), stream_of(x)), "sia_nchw_f32_to_nhwc4_bf16");
  return out;
}

at::Tensor conv7x7_c3_relu_pool2(const at::Tensor& x4, const at::Tensor& w_packed, const at::Tensor& bias) {
  check_cuda(x4, at::kBFloat16, "x4");
  check_cuda(w_packed, at::kByte, "w_packed");
  check_cuda(bias, at::kFloat, "bias");
  TORCH_CHECK(x4.dim() == 4 && x4.size(3) == 4 && x4.size(2) > SIA_NHWC4_PAD, "x4 must be padded NHWC4 [B,H,W+8,4]");
  TORCH_CHECK((size_t)w_packed.numel() == sia_pack_conv7x7_c3_bytes() && bias.numel() == 32,
              "w_packed / bias are not a packed 32-channel 7x7 block (sia_pack_conv7x7_c3)");
  const int64_t b = x4.size(0), h = x4.size(1), w = x4.size(2) - SIA_NHWC4_PAD;
  c10::cuda::CUDAGuard guard(x4.device());
  auto out = at::empty({b, h / 2, w / 2, 32}, x4.options());
  check_rc(sia_conv7x7_c3_relu_pool2(x4.data_ptr(), (int)b, (int)h, (int)w, w_packed.data_ptr(),
                                     bias.data_ptr<float>(), out.data_ptr(), stream_of(x4)),
           "sia_conv7x7_c3_relu_pool2");
  return out;
}

at::Tensor conv3x3_relu_pool2(const at::Tensor& x, const at::Tensor& w_packed, const at::Tensor& bias, int64_t cout) {
  check_cuda(x, at::kBFloat16, "x");
  check_cuda(w_packed, at::kByte, "w_packed");
  check_cuda(bias, at::kFloat, "bias");
  TORCH_CHECK(x.dim() == 4, "x must be NHWC [B,H,W,C]");
  TORCH_CHECK(bias.numel() == cout, "bias must have cout elements");
  const int64_t b = x.size(0), h = x.size(1), w = x.size(2), cin = x.size(3);
  c10::cuda::CUDAGuard guard(x.device());
  auto out = at::empty({b, h / 2, w / 2, cout}, x.options());
  check_rc(sia_conv3x3_relu_pool2(x.data_ptr(), (int)b, (int)h, (int)w, (int)cin, (int)cout, w_packed.data_ptr(),
                                  bias.data_ptr<float>(), out.data_ptr(), stream_of(x)),
           "sia_conv3x3_relu_pool2");
  return out;
}

at::Tensor linear_splitk(const at::Tensor& a, const at::Tensor& w, int64_t splits) {
  check_cuda(a, at::kBFloat16, "a");
  check_cuda(w, at::kBFloat16, "w");
  TORCH_CHECK(a.dim() == 2 && w.dim() == 2 && a.size(1) == w.size(1), "a [M,K] and w [N,K] must share K");
  TORCH_CHECK(splits >= 1, "splits must be positive");
  c10::cuda::CUDAGuard guard(a.device());
  auto out = at::empty({splits, a.size(0), w.size(0)}, a.options().dtype(at::kFloat));
  check_rc(sia_linear_splitk(a.data_ptr(), w.data_ptr(), (int)a.size(0), (int)w.size(0), (int)a.size(1), (int)splits,
                             out.data_ptr<float>(), stream_of(a)), "sia_linear_splitk");
  return out;
}

std::tuple<at::Tensor, at::Tensor> head_tail(const at::Tensor& partial, const at::Tensor& b1, const at::Tensor& w2t,
                                             const at::Tensor& b2, const at::Tensor& w3, const at::Tensor& b3) {
  check_cuda(partial, at::kFloat, "partial");
  for (const at::Tensor* t : {&b1, &w2t, &b2, &w3, &b3}) check_cuda(*t, at::kFloat, "tail parameter");
  TORCH_CHECK(partial.dim() == 3, "partial must be [splits, M, n1]");
  const int64_t splits = partial.size(0), m = partial.size(1), n1 = partial.size(2);
  TORCH_CHECK(w2t.dim() == 2 && w2t.size(0) == n1 && b1.numel() == n1, "w2t must be [n1, n2], b1 [n1]");
  const int64_t n2 = w2t.size(1);
  TORCH_CHECK(b2.numel() == n2 && w3.numel() == 2 * n2 && b3.numel() == 2, "the classifier has two classes");
  c10::cuda::CUDAGuard guard(partial.device());
  auto logp = at::empty({m, 2}, partial.options());
  auto pred = at::empty({m}, partial.options().dtype(at::kByte));
  check_rc(sia_head_tail(partial.data_ptr<float>(), (int)splits, (int)m, (int)n1, (int)n2, b1.data_ptr<float>(),
                         w2t.data_ptr<float>(), b2.data_ptr<float>(), w3.data_ptr<float>(), b3.data_ptr<float>(),
                         logp.data_ptr<float>(), pred.data_ptr<uint8_t>(), nullptr, nullptr, 0, 0, 0, nullptr,
                         stream_of(partial)), "sia_head_tail");
  return {logp, pred};
}

at::Tensor confusion_counts(const at::Tensor& pred, const at::Tensor& label, const at::Tensor& groups,
                            int64_t n_groups) {
  check_cuda(pred, at::kByte, "pred");
  check_cuda(label, at::kByte, "label");
  check_cuda(groups, at::kByte, "groups");
  TORCH_CHECK(pred.dim() == 1 && label.sizes() == pred.sizes() && groups.dim() == 2 && groups.size(1) == pred.size(0),
              "expected pred [N], label [N], groups [A,N]");
  TORCH_CHECK(n_groups >= 1 && n_groups <= 255, "n_groups must be in 1..255");
  c10::cuda::CUDAGuard guard(pred.device());
  auto counts = at::zeros({groups.size(0), n_groups, 2, 2}, pred.options().dtype(at::kLong));
  check_rc(sia_confusion_counts(pred.data_ptr<uint8_t>(), label.data_ptr<uint8_t>(), groups.data_ptr<uint8_t>(),
                                pred.size(0), pred.size(0), (int)groups.size(0), (int)n_groups,
                                reinterpret_cast<long long*>(counts.data_ptr<int64_t>()), stream_of(pred)),
           "sia_confusion_counts");
  return counts;
}

// Fused resize + scale / normalise + layout (the warp-MMA kernel); the tables come from
// skin_image_analysis_b200.resize_weights.build_mma_tables (host arithmetic on a few hundred numbers).
at::Tensor preprocess_mma(const at::Tensor& src, const at::Tensor& wy_frag, const at::Tensor& r0,
                          const at::Tensor& wx_frag, const at::Tensor& wx_mask, const at::Tensor& tile_begin,
                          int64_t kv, int64_t q_stride, at::IntArrayRef c_row, at::ArrayRef<double> mul,
                          at::ArrayRef<double> bias, int64_t out_h, int64_t out_w) {
  check_cuda(src, at::kByte, "src");
  check_cuda(wy_frag, at::kInt, "wy_frag");
  check_cuda(r0, at::kInt, "r0");
  check_cuda(wx_frag, at::kInt, "wx_frag");
  check_cuda(wx_mask, at::kInt, "wx_mask");
  check_cuda(tile_begin, at::kInt, "tile_begin");
  TORCH_CHECK(src.dim() == 4 && src.size(3) == 3, "src must be [B,H,W,3] uint8");
  TORCH_CHECK(c_row.size() == 4 && mul.size() == 3 && bias.size() == 3, "c_row has 4 entries, mul / bias 3");
  const int n_msteps = (int)r0.numel(), n_groups = (int)tile_begin.numel() - 1, n_tiles = (int)wx_mask.numel();
  int32_t c_row_h[4];
  float mul_h[3], bias_h[3];
  for (int i = 0; i < 4; ++i) c_row_h[i] = (int32_t)c_row[i];
  for (int i = 0; i < 3; ++i) {
    mul_h[i] = (float)mul[i];
    bias_h[i] = (float)bias[i];
  }
  c10::cuda::CUDAGuard guard(src.device());
  auto out = at::empty({src.size(0), out_h, out_w + SIA_NHWC4_PAD, 4}, src.options().dtype(at::kBFloat16));
  check_rc(sia_preprocess_mma_u8hwc(src.data_ptr<uint8_t>(), (int)src.size(0), (int)src.size(1), (int)src.size(2),
                                    wy_frag.data_ptr(), r0.data_ptr<int32_t>(), n_msteps, (int)kv, wx_frag.data_ptr(),
                                    reinterpret_cast<const uint32_t*>(wx_mask.data_ptr<int32_t>()),
                                    tile_begin.data_ptr<int32_t>(), n_groups, n_tiles, (int)q_stride, c_row_h, mul_h,
                                    bias_h, (int)out_h, (int)out_w, out.data_ptr(), stream_of(src)),
           "sia_preprocess_mma_u8hwc");
  return out;
}

int64_t version() { return sia_version(); }

}  // namespace

TORCH_LIBRARY(sia_b200, m) {
  m.def("version() -> int");
  m.def("nchw_f32_to_nhwc4(Tensor x) -> Tensor");
  m.def("conv7x7_c3_relu_pool2(Tensor x4, Tensor w_packed, Tensor bias) -> Tensor");
  m.def("conv3x3_relu_pool2(Tensor x, Tensor w_packed, Tensor bias, int cout) -> Tensor");
  m.def("linear_splitk(Tensor a, Tensor w, int splits) -> Tensor");
  m.def("head_tail(Tensor partial, Tensor b1, Tensor w2t, Tensor b2, Tensor w3, Tensor b3) -> (Tensor, Tensor)");
  m.def("confusion_counts(Tensor pred, Tensor label, Tensor groups, int n_groups) -> Tensor");
  m.def("preprocess_mma(Tensor src, Tensor wy_frag, Tensor r0, Tensor wx_frag, Tensor wx_mask, Tensor tile_begin, "
        "int kv, int q_stride, int[] c_row, float[] mul, float[] bias, int out_h, int out_w) -> Tensor");
}

// CompositeExplicitAutograd: one implementation for every backend key -- the TORCH_CHECKs inside refuse anything that is
// not a CUDA tensor, so there is no silent CPU path.
TORCH_LIBRARY_IMPL(sia_b200, CompositeExplicitAutograd, m) {
  m.impl("version", &version);
  m.impl("nchw_f32_to_nhwc4", &nchw_f32_to_nhwc4);
  m.impl("conv7x7_c3_relu_pool2", &conv7x7_c3_relu_pool2);
  m.impl("conv3x3_relu_pool2", &conv3x3_relu_pool2);
  m.impl("linear_splitk", &linear_splitk);
  m.impl("head_tail", &head_tail);
  m.impl("confusion_counts", &confusion_counts);
  m.impl("preprocess_mma", &preprocess_mma);
}
