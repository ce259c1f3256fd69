// jax_ext.cpp — Python module `jax_ext`: the raw-pointer ABI of the reference's
// extensions/jax/jax_ext.cpp:12-36.  The caller (JAX via DLPack/CuPy, or anything else that can
// produce device addresses) passes four integers and the shape; nothing here depends on JAX, CuPy
// or PyTorch, so the module builds and is tested with plain pybind11.
#include <pybind11/pybind11.h>

#include <cstdint>
#include <stdexcept>
#include <string>

#include "../../include/launchers.h"

namespace {

template <typename T>
T* device_ptr(std::uintptr_t address) {
  return reinterpret_cast<T*>(address);
}

// Synchronous on return, like the reference's `solve`; errors become Python exceptions instead of
// being dropped (the reference ignores CUDA errors inside launch(), include/launchers.h:27-71).
void run(std::uintptr_t q, std::uintptr_t k, std::uintptr_t v, std::uintptr_t out, int rows,
         int width, int heads, const std::string& variant) {
  // The kernel is chosen per call (no process-wide state is touched, so concurrent callers with different
  // kernels cannot race); gran -1 = the granularity solve() uses for this shape.
  const int kernel = qmha_kernel_from_name(variant.c_str());
  if (kernel < 0) throw std::invalid_argument("unknown kernel '" + variant + "'");
  if (qmha_forward(device_ptr<const float>(q), device_ptr<const float>(k), device_ptr<const float>(v),
                   device_ptr<float>(out), 1, rows, width, heads, kernel, -1, nullptr) != 0)
    throw std::runtime_error(qmha_last_error());
  if (qmha_synchronize(nullptr) != 0) throw std::runtime_error(qmha_last_error());   // complete on return, like solve()
}

}  // namespace

PYBIND11_MODULE(jax_ext, m) {
  namespace py = pybind11;
  m.doc() = "B200 quantised multi-head attention: `solve` on raw device addresses";
  m.def("flash_solve", &run, py::arg("q_ptr"), py::arg("k_ptr"), py::arg("v_ptr"), py::arg("out_ptr"),
        py::arg("N"), py::arg("d_model"), py::arg("num_heads"), py::arg("kernel") = "fa_tc_int8_b",
        "flash_solve(q_ptr, k_ptr, v_ptr, out_ptr, N, d_model, num_heads, kernel): attention forward on\n"
        "fp32 device buffers [N, d_model] given as integer addresses; returns when the result is complete.");
}
