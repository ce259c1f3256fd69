// torch_ext.cpp — same Python surface as the reference's extensions/torch/torch_ext.cpp:11-57
// (module `torch_ext`, flash_solve(Q, K, V, d_model, num_heads, kernel="fa_tc_int8_b")), bound
// to the B200 library's stream-ordered C entry point.  PyTorch appears only here: tensors in,
// raw pointers + the current CUDA stream out.
#include <ATen/cuda/CUDAContext.h>
#include <torch/extension.h>

#include <string>

#include "../../include/launchers.h"

using torch::Tensor;
namespace py = pybind11;

Tensor flash_solve(const Tensor& Q, const Tensor& K, const Tensor& V, int64_t d_model,
                   int64_t num_heads, const std::string& kernel = "fa_tc_int8_b") {
  TORCH_CHECK(Q.is_cuda() && K.is_cuda() && V.is_cuda(), "Inputs must be CUDA tensors");
  TORCH_CHECK(Q.dtype() == torch::kFloat32, "Q must be float32");
  TORCH_CHECK(K.dtype() == torch::kFloat32, "K must be float32");
  TORCH_CHECK(V.dtype() == torch::kFloat32, "V must be float32");
  TORCH_CHECK(Q.sizes() == K.sizes() && Q.sizes() == V.sizes(), "Q, K, V must have the same shape");
  auto Qc = Q.contiguous(), Kc = K.contiguous(), Vc = V.contiguous();
  const int64_t elems = Qc.numel();
  TORCH_CHECK(elems % d_model == 0, "Q.numel() must be divisible by d_model");
  // [N, d_model] like the reference, or [B, N, d_model]
  const int64_t B = Qc.dim() == 3 ? Qc.size(0) : 1;
  const int64_t N = elems / d_model / B;
  auto out = torch::empty_like(Qc);
  const int kid = qmha_kernel_from_name(kernel.c_str());
  TORCH_CHECK(kid >= 0, "unknown kernel '", kernel, "'");
  const int rc = qmha_forward(Qc.data_ptr<float>(), Kc.data_ptr<float>(), Vc.data_ptr<float>(),
                              out.data_ptr<float>(), (int)B, (int)N, (int)d_model, (int)num_heads, kid,
                              qmha_default_granularity((int)d_model, (int)num_heads),
                              at::cuda::getCurrentCUDAStream().stream());
  TORCH_CHECK(rc == 0, qmha_last_error());
  return out;
}

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.def("flash_solve", &flash_solve,
        "Fused multi-head attention forward on B200 (tcgen05).\n\n"
        "Args:\n  Q, K, V: float32 CUDA tensors [N, d_model] or [B, N, d_model]\n"
        "  d_model: model dimension\n  num_heads: attention heads (d_model/num_heads <= 128)\n"
        "  kernel: 'fa_tc_int8_b' (INT8, default), 'fa_tc_v2a' (FP16) or any reference kernel name",
        py::arg("Q"), py::arg("K"), py::arg("V"), py::arg("d_model"), py::arg("num_heads"),
        py::arg("kernel") = "fa_tc_int8_b");
}
