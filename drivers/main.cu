// drivers/main.cu — bin/profile_<KERNEL>: the reference's profiling driver (drivers/main.cu:38-157)
// rebuilt around the run-time-shaped C-ABI.
//
// Flags kept from the reference (drivers/main.cu:45-58), same defaults (warmup 2, runs 3, check
// on) and same exit code 1 on a failed check (:97-99):
//     --kernel=K | -k K   --warmup=N   --runs=M   --check=0|1   --no-check   --check   --random
//     --help
// New optional flags: --N= --d_model= --h= --B= (defaults: include/config.h), --json.
// The reference takes no timing in code (numbers came from ncu); this driver brackets the
// profiled region with cudaProfilerStart/Stop like the reference AND times it with CUDA events.
#include <cuda_profiler_api.h>
#include <cuda_runtime.h>
#include <sys/stat.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../include/config.h"
#include "../include/launchers.h"
#include "../inputs/data.h"
#include "../tools/check_cuda.h"
#include "../utils/verify.h"

using namespace qmha_driver;

#ifndef QMHA_DEFAULT_KERNEL
#define QMHA_DEFAULT_KERNEL "fa_tc_int8_b"
#endif

static bool starts_with(const char* s, const char* p) { return std::strncmp(s, p, std::strlen(p)) == 0; }

int main(int argc, char** argv) {
  std::string kernel = QMHA_DEFAULT_KERNEL;
  int warmup = 2, runs = 3;
  bool use_random = false, do_check = true, json = false, rope = false;
  int pN = N, pD = d_model, pH = h, pB = 1;

  for (int i = 1; i < argc; ++i) {
    const char* a = argv[i];
    if (starts_with(a, "--kernel=")) kernel = a + 9;
    else if (!std::strcmp(a, "-k") && i + 1 < argc) kernel = argv[++i];
    else if (starts_with(a, "--warmup=")) warmup = std::atoi(a + 9);
    else if (starts_with(a, "--runs=")) runs = std::atoi(a + 7);
    else if (starts_with(a, "--check=")) do_check = std::atoi(a + 8) != 0;
    else if (!std::strcmp(a, "--no-check")) do_check = false;
    else if (!std::strcmp(a, "--check")) do_check = true;
    else if (!std::strcmp(a, "--random")) use_random = true;
    else if (starts_with(a, "--N=")) pN = std::atoi(a + 4);
    else if (starts_with(a, "--d_model=")) pD = std::atoi(a + 10);
    else if (starts_with(a, "--h=")) pH = std::atoi(a + 4);
    else if (starts_with(a, "--B=")) pB = std::atoi(a + 4);
    else if (!std::strcmp(a, "--json")) json = true;
    else if (!std::strcmp(a, "--rope")) rope = true;  // fused RoPE on Q,K (utils/verify.cu:56-69 semantics)
    else if (!std::strcmp(a, "--help")) {
      std::printf("Usage: %s [--kernel=KERNEL] [--warmup=N] [--runs=M] [--check=0|1] [--no-check] [--random]\n"
                  "          [--N=rows] [--d_model=cols] [--h=heads] [--B=batch] [--json] [--rope]\n", argv[0]);
      std::printf("  KERNEL options: fa_tc_int8_b fa_tc_int8_a (INT8 tcgen05 path); fa_tc_v2a fa_tc_v1a fa unfused ...\n"
                  "                  (FP16 tcgen05 path); native names fa_b200_int8, fa_b200_f16\n");
      return 0;
    }
  }
  if (qmha_set_kernel(kernel.c_str()) != 0) {
    std::fprintf(stderr, "%s\n", qmha_last_error());
    return 2;
  }
  const int kid = qmha_kernel_from_name(kernel.c_str());
  if (rope && qmha_set_rope(1, 10000.0f) != 0) {
    std::fprintf(stderr, "%s\n", qmha_last_error());
    return 2;
  }
  mkdir(".cache", 0755);
  const size_t per_batch = (size_t)pN * pD;

  if (do_check) {
    // Known-answer test on constant inputs (drivers/main.cu:73-101) through solve(), plus — new —
    // a float64 spot check of sampled query rows on the U[0,1) profiling inputs, which unlike the
    // all-ones case depends on the attention weights being right.
    std::printf("Initializing host data (constant values for correctness check)...\n");
    HostQKV ones;
    fill_inputs(ones, pN, pD, Fill::Ones);
    DeviceQKV dev;
    dev.upload(ones);
    std::printf("Running correctness check \n");
    solve(dev.q, dev.k, dev.v, dev.out, pN, pD, pH);
    CHECK_CUDA(cudaDeviceSynchronize());
    std::vector<float> got(per_batch);
    CHECK_CUDA(cudaMemcpy(got.data(), dev.out, dev.bytes, cudaMemcpyDeviceToHost));
    if (std::getenv("QMHA_DRIVER_CORRUPT")) got[got.size() / 2] += 1.0f;   // test hook: force a failed check
    // Reference of the constant-input check, cached like the reference's driver does (drivers/main.cu:85-94,
    // ".cache/ref_N%d_d%d.bin"): a file written by the reference's own binary is used as is; otherwise the exact
    // answer (softmax of equal scores times a constant V = 1) is written in the same format.
    std::vector<float> ref;
    const std::string ref_path = ref_cache_path(pN, pD);
    if (read_ref_cache(ref, ref_path, pN, pD)) {
      std::printf("Loaded CPU reference from %s\n", ref_path.c_str());
    } else {
      ref.assign(per_batch, 1.0f);
      if (write_ref_cache(ref, ref_path, pN, pD)) std::printf("Saved CPU reference to %s\n", ref_path.c_str());
    }
    bool ok = *qmha_last_error() == 0;
    for (size_t i = 0; ok && i < got.size(); ++i)   // verify_results(h_output, ref_output, 1e-3f, 1e-3f), utils/verify.cu:153-173
      if (!std::isfinite(got[i]) || std::fabs(got[i] - ref[i]) > std::fmax(1e-3f, 1e-3f * std::fabs(ref[i]))) {
        std::fprintf(stderr, "Mismatch at index: %zu: got=%g ref=%g tol=0.001\n", i, got[i], ref[i]);
        ok = false;
      }
    if (ok) {
      HostQKV rnd;
      fill_inputs(rnd, pN, pD, Fill::Uniform01);
      dev.upload(rnd);
      solve(dev.q, dev.k, dev.v, dev.out, pN, pD, pH);
      CHECK_CUDA(cudaDeviceSynchronize());
      CHECK_CUDA(cudaMemcpy(got.data(), dev.out, dev.bytes, cudaMemcpyDeviceToHost));
      std::vector<int> rows;
      for (int r = 0; r < pN; r += std::max(1, pN / 7)) rows.push_back(r);
      rows.push_back(pN - 1);
      std::vector<double> expect;
      if (rope) {  // the reference's CPU check rotates Q and K (utils/verify.cu:56-69)
        apply_rope_host(rnd.q, pN, pD, pH);
        apply_rope_host(rnd.k, pN, pD, pH);
      }
      expected_rows(rnd.q, rnd.k, rnd.v, pN, pD, pH, rows, expect);
      const float eps = kid == QMHA_KERNEL_INT8 ? 2e-2f : 2e-3f;
      CheckReport rep = compare_rows(got, expect, rows, pD, eps, 1e-3f);
      std::printf("Sampled-row check: %zu values, worst |diff| %.3e (tol %.0e)\n", rep.checked, rep.worst_abs, (double)eps);
      ok = rep.pass && *qmha_last_error() == 0;
    }
    dev.release();
    if (!ok) {
      std::fprintf(stderr, "Correctness check FAILED. Aborting.\n");
      return 1;
    }
    std::printf("Correctness check PASSED.\n");
  } else {
    std::printf("Skipping correctness check and CPU reference load/compute (--no-check / --check=0).\n");
  }

  // Profiling inputs: cached U[0,1) data in the reference's file format (B=1), or regenerated.
  HostQKV in;
  const std::string cache = input_cache_path(pN, pD);
  if (pB != 1 || use_random || !read_input_cache(in, cache, pN, pD)) {
    std::printf("Generating random input data for profiling...\n");
    fill_inputs(in, pB * pN, pD, Fill::Uniform01);
    if (pB == 1 && write_input_cache(in, cache, pN)) std::printf("Saved input matrices to %s\n", cache.c_str());
  } else {
    std::printf("Loaded input matrices from %s\n", cache.c_str());
  }
  DeviceQKV dev;
  dev.upload(in);

  auto run_once = [&]() {
    if (pB == 1) {
      solve(dev.q, dev.k, dev.v, dev.out, pN, pD, pH);
    } else {
      if (qmha_forward(dev.q, dev.k, dev.v, dev.out, pB, pN, pD, pH, kid, -1, nullptr) != 0) {
        std::fprintf(stderr, "%s\n", qmha_last_error());
        std::exit(1);
      }
    }
    CHECK_CUDA(cudaDeviceSynchronize());
  };

  CHECK_CUDA(cudaProfilerStart());
  std::printf("Running %d warmup iterations...\n", warmup);
  for (int i = 0; i < warmup; ++i) run_once();
  std::printf("Running %d profiling iterations...\n", runs);
  std::vector<float> ms(std::max(runs, 0));
  cudaEvent_t e0, e1;
  CHECK_CUDA(cudaEventCreate(&e0));
  CHECK_CUDA(cudaEventCreate(&e1));
  for (int r = 0; r < runs; ++r) {
    CHECK_CUDA(cudaEventRecord(e0));
    run_once();
    CHECK_CUDA(cudaEventRecord(e1));
    CHECK_CUDA(cudaEventSynchronize(e1));
    CHECK_CUDA(cudaEventElapsedTime(&ms[r], e0, e1));
  }
  CHECK_CUDA(cudaProfilerStop());

  std::vector<float> host_out((size_t)pB * per_batch);
  CHECK_CUDA(cudaMemcpy(host_out.data(), dev.out, dev.bytes, cudaMemcpyDeviceToHost));
  dev.release();
  if (runs > 0) {
    std::sort(ms.begin(), ms.end());
    const double flops = 4.0 * pB * pH * (double)pN * pN * (pD / pH);
    const double med = ms[runs / 2], best = ms[0];
    std::printf("solve(): min %.3f ms, median %.3f ms  ->  %.1f TFLOP/s (median, quantise + attention)\n", best, med,
                flops / med / 1e9);
    if (json)
      std::printf("{\"kernel\": \"%s\", \"B\": %d, \"N\": %d, \"d_model\": %d, \"h\": %d, \"ms_min\": %.4f, "
                  "\"ms_median\": %.4f, \"tflops_median\": %.2f}\n", kernel.c_str(), pB, pN, pD, pH, best, med,
                  flops / med / 1e9);
  }
  std::printf("Profiling complete.\n");
  qmha_shutdown();
  return 0;
}
