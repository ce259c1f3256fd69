// ref_shim_verify.cpp — C entry points onto the UNMODIFIED reference utils/verify.cu and
// inputs/data.cu (host-only code despite the .cu extension).  TEST INFRASTRUCTURE ONLY; sources
// are compiled from where they lie (-I$(REF)), nothing is copied.
#include <cstdio>
#include "utils/verify.cu"
#include "inputs/data.cu"

extern "C" {

// utils/verify.cu:25-104 — the reference's CPU path *with RoPE*, single-threaded.
void ref_cpu_reference(const float* Q, const float* K, const float* V, float* O, int N,
                       int d_model, int h) {
  std::vector<float> q(Q, Q + (size_t)N * d_model), k(K, K + (size_t)N * d_model),
      v(V, V + (size_t)N * d_model), o;
  // cpu_reference prints a progress line per 512 rows; silence it for library use.
  FILE* saved = stdout;
  FILE* devnull = fopen("/dev/null", "w");
  if (devnull) stdout = devnull;
  cpu_reference(q, k, v, o, N, d_model, h);
  if (devnull) { stdout = saved; fclose(devnull); }
  std::copy(o.begin(), o.end(), O);
}

// utils/verify.cu:153-173
int ref_verify_results(const float* out, const float* ref, long long n, float eps, float rel) {
  std::vector<float> a(out, out + n), b(ref, ref + n);
  return verify_results(a, b, eps, rel) ? 1 : 0;
}

// inputs/data.cu:9-30
void ref_initialize_host_data(float* Q, float* K, float* V, int N, int d_model, int use_random) {
  std::vector<float> q, k, v;
  initialize_host_data(q, k, v, N, d_model, use_random != 0);
  std::copy(q.begin(), q.end(), Q);
  std::copy(k.begin(), k.end(), K);
  std::copy(v.begin(), v.end(), V);
}

// utils/verify.cu:106-151 and inputs/data.cu:54-109 (cache file formats)
int ref_save_reference(const float* data, const char* path, int N, int d_model) {
  std::vector<float> d(data, data + (size_t)N * d_model);
  return save_reference(d, path, N, d_model) ? 1 : 0;
}
int ref_load_reference(float* data, const char* path, int N, int d_model) {
  std::vector<float> d;
  if (!load_reference(d, path, N, d_model)) return 0;
  std::copy(d.begin(), d.end(), data);
  return 1;
}

// inputs/data.cu:54-109 (".cache/input_random_N%d_d%d.bin")
int ref_save_inputs(const float* Q, const float* K, const float* V, const char* path, int N, int d_model) {
  const size_t n = (size_t)N * d_model;
  std::vector<float> q(Q, Q + n), k(K, K + n), v(V, V + n);
  FILE* saved = stdout;
  FILE* devnull = fopen("/dev/null", "w");
  if (devnull) stdout = devnull;
  const bool ok = save_inputs(q, k, v, path, N, d_model);
  if (devnull) { stdout = saved; fclose(devnull); }
  return ok ? 1 : 0;
}
int ref_load_inputs(float* Q, float* K, float* V, const char* path, int N, int d_model) {
  std::vector<float> q, k, v;
  FILE* saved = stdout;
  FILE* devnull = fopen("/dev/null", "w");
  if (devnull) stdout = devnull;
  const bool ok = load_inputs(q, k, v, path, N, d_model);
  if (devnull) { stdout = saved; fclose(devnull); }
  if (!ok) return 0;
  std::copy(q.begin(), q.end(), Q);
  std::copy(k.begin(), k.end(), K);
  std::copy(v.begin(), v.end(), V);
  return 1;
}

}  // extern "C"
