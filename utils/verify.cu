// utils/verify.cu — see verify.h.  Host-only.
#include "verify.h"

#include <cmath>
#include <cstdio>
#include <fstream>

namespace qmha_driver {

void expected_rows(const std::vector<float>& q, const std::vector<float>& k,
                   const std::vector<float>& v, int N, int d_model, int h,
                   const std::vector<int>& rows, std::vector<double>& expect) {
  const int d = d_model / h;
  const double inv_sqrt_d = 1.0 / std::sqrt((double)d);
  expect.assign(rows.size() * (size_t)d_model, 0.0);
  std::vector<double> w(N);
  for (size_t ri = 0; ri < rows.size(); ++ri) {
    const int i = rows[ri];
    for (int head = 0; head < h; ++head) {
      const int c0 = head * d;
      double top = -INFINITY;
      for (int j = 0; j < N; ++j) {
        double dot = 0.0;
        for (int c = 0; c < d; ++c)
          dot += (double)q[(size_t)i * d_model + c0 + c] * (double)k[(size_t)j * d_model + c0 + c];
        w[j] = dot * inv_sqrt_d;
        if (w[j] > top) top = w[j];
      }
      double denom = 0.0;
      for (int j = 0; j < N; ++j) {
        w[j] = std::exp(w[j] - top);
        denom += w[j];
      }
      double* e = &expect[ri * (size_t)d_model + c0];
      for (int j = 0; j < N; ++j) {
        const double p = w[j] / denom;
        for (int c = 0; c < d; ++c) e[c] += p * (double)v[(size_t)j * d_model + c0 + c];
      }
    }
  }
}

void apply_rope_host(std::vector<float>& x, int N, int d_model, int h, float base) {
  const int d = d_model / h;
  for (int i = 0; i < N; ++i)
    for (int head = 0; head < h; ++head) {
      float* row = &x[(size_t)i * d_model + (size_t)head * d];
      for (int k = 0; k < d / 2; ++k) {
        const float theta = std::pow(base, -static_cast<float>(2 * k) / d);
        const float angle = i * theta;
        const float sin_a = std::sin(angle), cos_a = std::cos(angle);
        const float a = row[k], b = row[k + d / 2];
        row[k] = a * cos_a - b * sin_a;
        row[k + d / 2] = a * sin_a + b * cos_a;
      }
    }
}

CheckReport compare_rows(const std::vector<float>& out, const std::vector<double>& expect,
                         const std::vector<int>& rows, int d_model, float eps, float rel) {
  CheckReport r;
  for (size_t ri = 0; ri < rows.size(); ++ri)
    for (int c = 0; c < d_model; ++c) {
      const size_t idx = (size_t)rows[ri] * d_model + c;
      const double a = out[idx], b = expect[ri * (size_t)d_model + c];
      ++r.checked;
      const double tol = std::fmax((double)eps, (double)rel * std::fabs(b));
      const double diff = std::fabs(a - b);
      const bool bad = !std::isfinite(a) || !std::isfinite(b) || diff > tol;
      if (std::isfinite(diff) && diff > r.worst_abs) r.worst_abs = diff;
      if (bad && r.pass) {
        r.pass = false;
        r.first_bad = idx;
        std::fprintf(stderr, "Mismatch at index: %zu: got=%g ref=%g tol=%g\n", idx, a, b, tol);
      }
    }
  return r;
}

std::string ref_cache_path(int N, int d_model) {
  char buf[128];
  std::snprintf(buf, sizeof buf, ".cache/ref_N%d_d%d.bin", N, d_model);
  return buf;
}

bool write_ref_cache(const std::vector<float>& data, const std::string& path, int N, int d_model) {
  std::ofstream f(path, std::ios::binary);
  if (!f) return false;
  const int hdr[2] = {N, d_model};
  f.write(reinterpret_cast<const char*>(hdr), sizeof hdr);
  f.write(reinterpret_cast<const char*>(data.data()), (std::streamsize)(data.size() * sizeof(float)));
  return (bool)f;
}

bool read_ref_cache(std::vector<float>& data, const std::string& path, int N, int d_model) {
  std::ifstream f(path, std::ios::binary);
  if (!f) return false;
  int hdr[2] = {0, 0};
  f.read(reinterpret_cast<char*>(hdr), sizeof hdr);
  if (!f || hdr[0] != N || hdr[1] != d_model) return false;
  data.resize((size_t)N * d_model);
  f.read(reinterpret_cast<char*>(data.data()), (std::streamsize)(data.size() * sizeof(float)));
  return (bool)f;
}

}  // namespace qmha_driver
