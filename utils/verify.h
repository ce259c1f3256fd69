// utils/verify.h — the driver's self-check (drivers/main.cu --check).  Same pass/fail rule as the
// reference's verify_results (utils/verify.cu:153-173): |a-b| <= max(eps, rel*|b|), non-finite
// values fail — and the same ".cache/ref_N%d_d%d.bin" file format (utils/verify.cu:106-151).
// The expected values are computed in float64 for a SAMPLE of query rows (plain attention, no
// RoPE: the GPU kernels never applied it, SURVEY.md finding 2); the check never feeds results
// back into the product path.
#pragma once
#include <string>
#include <vector>

namespace qmha_driver {

struct CheckReport {
  bool pass = true;
  size_t first_bad = 0;
  double worst_abs = 0.0;
  size_t checked = 0;
};

// Expected output rows `rows` of one [N, d_model] problem with h heads, float64 accumulation.
void expected_rows(const std::vector<float>& q, const std::vector<float>& k,
                   const std::vector<float>& v, int N, int d_model, int h,
                   const std::vector<int>& rows, std::vector<double>& expect);

// RoPE of the reference's CPU check (utils/verify.cu:9-23), applied in place to every head row of a
// [N, d_model] matrix (position = row index); used by the driver's --rope check.
void apply_rope_host(std::vector<float>& x, int N, int d_model, int h, float base = 10000.0f);

CheckReport compare_rows(const std::vector<float>& out, const std::vector<double>& expect,
                         const std::vector<int>& rows, int d_model, float eps, float rel);

std::string ref_cache_path(int N, int d_model);
bool write_ref_cache(const std::vector<float>& data, const std::string& path, int N, int d_model);
bool read_ref_cache(std::vector<float>& data, const std::string& path, int N, int d_model);

}  // namespace qmha_driver
