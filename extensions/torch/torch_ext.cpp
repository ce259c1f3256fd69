// torch_ext.cpp — Python module `torch_ext` with the reference's entry point
//   flash_solve(Q, K, V, d_model, num_heads, kernel="fa_tc_int8_b") -> Tensor
// (surface of extensions/torch/torch_ext.cpp:11-57 in the reference), implemented on the B200
// library's stream-ordered C-ABI.  PyTorch appears only here: tensors in, raw pointers and torch's
// current CUDA stream out.  Differences from the reference wrapper: `kernel` really selects the
// variant, a leading batch dimension is accepted, the work is enqueued on the caller's stream.
#include <ATen/cuda/CUDAContext.h>
#include <torch/extension.h>

#include <array>
#include <string>

#include "../../include/launchers.h"

namespace {

struct Problem {
  int batch, rows, width, heads;
};

// Validates the three operands the way the reference does (CUDA, float32) and derives the shape.
Problem describe(const std::array<const at::Tensor*, 3>& qkv, int64_t d_model, int64_t num_heads) {
  static const char* const names[3] = {"Q", "K", "V"};
  for (const at::Tensor* t : qkv) TORCH_CHECK(t->is_cuda(), "Inputs must be CUDA tensors");
  // float32 like the reference (torch_ext.cpp:15-17); float16 / bfloat16 are accepted as an extension
  for (int i = 0; i < 3; ++i) {
    const auto st = qkv[i]->scalar_type();
    TORCH_CHECK(st == at::kFloat || st == at::kHalf || st == at::kBFloat16, names[i], " must be float32");
    TORCH_CHECK(st == qkv[0]->scalar_type(), "Q, K, V must have the same dtype");
  }
  for (int i = 1; i < 3; ++i)
    TORCH_CHECK(qkv[i]->sizes() == qkv[0]->sizes(), "Q, K, V must have the same shape");
  const at::Tensor& q = *qkv[0];
  TORCH_CHECK(d_model > 0 && q.numel() % d_model == 0, "Q.numel() must be divisible by d_model");
  const int64_t batch = q.dim() == 3 ? q.size(0) : 1;  // [B, N, d_model] or the reference's [N, d_model]
  return {(int)batch, (int)(q.numel() / d_model / batch), (int)d_model, (int)num_heads};
}

at::Tensor flash_solve(const at::Tensor& Q, const at::Tensor& K, const at::Tensor& V,
                       int64_t d_model, int64_t num_heads, const std::string& kernel) {
  const Problem p = describe({&Q, &K, &V}, d_model, num_heads);
  const int variant = qmha_kernel_from_name(kernel.c_str());
  TORCH_CHECK(variant >= 0, "unknown kernel '", kernel, "'");
  const at::Tensor q = Q.contiguous(), k = K.contiguous(), v = V.contiguous();
  at::Tensor result = at::empty_like(q);
  const int dtype = q.scalar_type() == at::kFloat ? QMHA_DTYPE_F32 : (q.scalar_type() == at::kHalf ? QMHA_DTYPE_F16 : QMHA_DTYPE_BF16);
  qmha_args a;
  qmha_args_init(&a);   // kernel, granularity (-1 = what solve() uses for the shape) and stream are per-call
  a.Q = q.data_ptr(); a.K = k.data_ptr(); a.V = v.data_ptr(); a.O = result.data_ptr();
  a.B = p.batch; a.N = p.rows; a.d_model = p.width; a.h = p.heads;
  a.kernel = variant;
  a.in_dtype = a.out_dtype = dtype;
  a.stream = at::cuda::getCurrentCUDAStream().stream();
  // a pipeline failure recorded by an EARLIER launch on this device makes this call fail (checked without a sync)
  TORCH_CHECK(qmha_forward_ex(&a) == 0, qmha_last_error());
  return result;
}

}  // namespace

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  namespace py = pybind11;
  m.def("flash_solve", &flash_solve, py::arg("Q"), py::arg("K"), py::arg("V"), py::arg("d_model"),
        py::arg("num_heads"), py::arg("kernel") = "fa_tc_int8_b",
        "Multi-head attention forward on a B200 (tcgen05 kernels behind the `solve` C-ABI).\n"
        "Q, K, V: float32 CUDA tensors, [N, d_model] or [B, N, d_model]; d_model / num_heads <= 128.\n"
        "kernel: any reference kernel name — the INT8 family (fa_tc_int8_a/b) or the FP16 family\n"
        "(fa, unfused, fa_tc_v1a ... fa_tc_v2b) — or the native names fa_b200_int8 / fa_b200_f16.");
}
