// inputs/data.h — synthetic Q/K/V for the profile_* driver and the on-disk input cache.
// Same observable behaviour as the reference's inputs/data.{h,cu}: constant 1.0 inputs for the
// known-answer check, mt19937(42) U[0,1) inputs drawn Q,K,V interleaved for profiling
// (inputs/data.cu:15-22), and the ".cache/input_random_N%d_d%d.bin" file format
// ({int N; int d_model;} header + Q + K + V as fp32, inputs/data.cu:54-109).
#pragma once
#include <string>
#include <vector>

namespace qmha_driver {

struct HostQKV {
  int rows = 0;     // B * N
  int d_model = 0;
  std::vector<float> q, k, v;
  size_t elems() const { return (size_t)rows * d_model; }
};

enum class Fill { Ones, Uniform01 };

void fill_inputs(HostQKV& t, int rows, int d_model, Fill how);
std::string input_cache_path(int N, int d_model);
bool write_input_cache(const HostQKV& t, const std::string& path, int N);
bool read_input_cache(HostQKV& t, const std::string& path, int N, int d_model);

struct DeviceQKV {
  float *q = nullptr, *k = nullptr, *v = nullptr, *out = nullptr;
  size_t bytes = 0;
  void upload(const HostQKV& t);
  void release();
};

}  // namespace qmha_driver
