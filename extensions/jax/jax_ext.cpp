// jax_ext.cpp — same pointer ABI as the reference's extensions/jax/jax_ext.cpp:12-36: device
// addresses as integers, forwarded to the C entry point.  No JAX/CuPy needed to build or test it.
#include <pybind11/pybind11.h>

#include <cstdint>
#include <stdexcept>
#include <string>

#include "../../include/launchers.h"

namespace py = pybind11;

void flash_solve(unsigned long long q_ptr, unsigned long long k_ptr, unsigned long long v_ptr,
                 unsigned long long out_ptr, int N, int d_model, int num_heads,
                 const std::string& kernel = "fa_tc_int8_b") {
  if (qmha_set_kernel(kernel.c_str()) != 0) throw std::invalid_argument(qmha_last_error());
  solve(reinterpret_cast<const float*>(q_ptr), reinterpret_cast<const float*>(k_ptr),
        reinterpret_cast<const float*>(v_ptr), reinterpret_cast<float*>(out_ptr), N, d_model,
        num_heads);  // synchronous on return, like the reference
  if (*qmha_last_error()) throw std::runtime_error(qmha_last_error());
}

PYBIND11_MODULE(jax_ext, m) {
  m.doc() = "Pointer-based wrapper of the B200 quantised-MHA `solve` (JAX / DLPack / CuPy callers)";
  m.def("flash_solve", &flash_solve, "Call `solve` with device pointers (uint64 addresses)",
        py::arg("q_ptr"), py::arg("k_ptr"), py::arg("v_ptr"), py::arg("out_ptr"), py::arg("N"),
        py::arg("d_model"), py::arg("num_heads"), py::arg("kernel") = "fa_tc_int8_b");
}
