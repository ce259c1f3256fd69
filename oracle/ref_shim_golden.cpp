// ref_shim_golden.cpp — C entry points onto the UNMODIFIED reference tests/generate_golden.cpp.
//
// TEST INFRASTRUCTURE ONLY (see qmha_oracle.cpp header).  The reference source is compiled from
// where it lies (-I$(REF) on the command line, oracle/Makefile); nothing is copied into this
// repository.  generate_golden.cpp is a program with static helpers and a main(); including it
// into this translation unit with main renamed makes cpu_mha / quantize_int8 / apply_rope_cpu
// callable so the restatement in qmha_oracle.cpp can be validated against the real thing.
#include <sstream>
#define main qmha_ref_generate_golden_main
#include "tests/generate_golden.cpp"
#undef main

extern "C" {

// tests/generate_golden.cpp:53-92
void ref_cpu_mha(const float* Q, const float* K, const float* V, float* O, int N, int d_model,
                 int h) {
  std::vector<float> q(Q, Q + (size_t)N * d_model), k(K, K + (size_t)N * d_model),
      v(V, V + (size_t)N * d_model);
  std::vector<float> o = cpu_mha(q, k, v, N, d_model, h, nullptr, nullptr);
  std::copy(o.begin(), o.end(), O);
}

// tests/generate_golden.cpp:94-101
void ref_quantize_int8(const float* src, long long n, float scale, float zero_point,
                       signed char* dst) {
  std::vector<float> s(src, src + n);
  std::vector<int8_t> d;
  quantize_int8(s, d, scale, zero_point);
  std::copy(d.begin(), d.end(), dst);
}

// tests/generate_golden.cpp:38-51
void ref_apply_rope_row(float* row, int pos, int d) { apply_rope_cpu(row, pos, d); }

// The whole program (writes tests/golden/<case>/ under the current directory, ~1.5 GB).
int ref_generate_golden_main(void) { return qmha_ref_generate_golden_main(); }

}  // extern "C"
