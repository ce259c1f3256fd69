// api.cu — the C-ABI of include/qmha.h: solve() and the extended entry points, per-device
// workspaces, error reporting.  Host code only; kernels live in prepare.cu and attn_fwd.cu.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/qmha.h"
#include "attn_fwd.cuh"
#include "prepare.cuh"

#ifndef QMHA_DEFAULT_KERNEL
#define QMHA_DEFAULT_KERNEL "fa_tc_int8_b"
#endif
#ifndef QMHA_DEFAULT_ATTN_VARIANT
#define QMHA_DEFAULT_ATTN_VARIANT 0
#endif

namespace {

thread_local std::string g_err;
std::atomic<long long> g_launches{0};
std::mutex g_mu;
int g_default_kernel = -2;  // -2 = not resolved yet
// Fused RoPE on Q and K inside the quantise / convert pass (qmha_set_rope; QMHA_ROPE=1 in the
// environment turns it on for processes that only call solve(), e.g. bin/profile_* --rope).
int g_rope = -1;            // -1 = not resolved yet (environment), 0 = off, 1 = on
float g_rope_base = 10000.0f;

int fail(const std::string& msg) {
  g_err = msg;
  return 1;
}
int fail_cuda(const char* what, cudaError_t e) {
  return fail(std::string(what) + ": " + cudaGetErrorString(e));
}

struct Workspace {
  void* Qp = nullptr;
  void* Kp = nullptr;
  void* Vt = nullptr;
  float* scales = nullptr;
  unsigned* amax = nullptr;
  float* aux = nullptr;    // block mode: [units][n_pad/32][2]
  float* vmax = nullptr;   // block mode: [units]
  int* error_flag = nullptr;
  unsigned long long* cycles = nullptr;  // {sum of CTA residency clocks, CTAs}: qmha_debug_cycles
  size_t qk_bytes = 0, vt_bytes = 0, scale_elems = 0;
  float2* rope_tab = nullptr;  // {cos, sin}[rope_n][rope_d/2] for rope_base
  int rope_n = 0, rope_d = 0;
  float rope_base = 0.f;
  // The workspace is shared by every call on the device.  `call_mu` serialises the host-side enqueue
  // sequences (and any growth) of concurrent callers; `last_use` is recorded behind the last enqueued
  // work that touches the workspace, and a call on another stream waits for it on the device first
  // (WorkspaceUse), so asynchronous calls on different streams cannot overwrite each other's operands.
  std::mutex call_mu;
  cudaEvent_t last_use = nullptr;
  cudaStream_t last_stream = nullptr;
  bool in_flight = false;
};
std::map<int, Workspace> g_ws;  // per device

// Holds a workspace for one enqueue sequence on stream `s` (see Workspace::call_mu / last_use).
struct WorkspaceUse {
  Workspace* w = nullptr;
  cudaStream_t s = nullptr;
  std::unique_lock<std::mutex> lk;
  bool armed = false;
  // takes over the lock acquired by get_workspace()
  void begin(Workspace* w_, std::unique_lock<std::mutex>&& lk_, cudaStream_t s_) {
    w = w_; s = s_; lk = std::move(lk_);
    if (w->in_flight && w->last_stream != s) cudaStreamWaitEvent(s, w->last_use, 0);
    armed = true;
  }
  ~WorkspaceUse() {
    if (!armed) return;
    if (!w->last_use && cudaEventCreateWithFlags(&w->last_use, cudaEventDisableTiming) != cudaSuccess) return;
    if (cudaEventRecord(w->last_use, s) == cudaSuccess) { w->last_stream = s; w->in_flight = true; }
  }
};

int round_up(int x, int m) { return (x + m - 1) / m * m; }

int pad_head_dim(int d) { return d <= 32 ? 32 : (d <= 64 ? 64 : (d <= 128 ? 128 : -1)); }

int check_shape(int B, int N, int d_model, int h, int* d_out, int* n_pad, int* d_pad) {
  if (B < 1 || N < 1 || d_model < 1 || h < 1) return fail("B, N, d_model and h must be positive");
  if (d_model % h != 0) return fail("d_model must be divisible by h (config.h:27)");
  const int d = d_model / h;
  const int dp = pad_head_dim(d);
  if (dp < 0) return fail("per-head dimension d = d_model/h must be <= 128");
  *d_out = d;
  *n_pad = round_up(N, 256);
  *d_pad = dp;
  return 0;
}

int require_device() {
  int dev = -1;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) { fail_cuda("no CUDA device (this library has no CPU fallback)", e); return -1; }
  int major = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (major != 10) {
    fail("device is not sm_100 (B200); this library only contains sm_100a code");
    return -1;
  }
  return dev;
}

// Grows (never shrinks) the calling device's workspace.  cudaMalloc only on growth, so steady
// state has no allocation in the timed path (the reference mallocs per head per call,
// launchers.h:27-39 and fa_tc_int8_b.cu:589-597).
// Returns with the workspace's call lock held (`call_lock`), which also covers the growth below.
int get_workspace(int dev, size_t qk_bytes, size_t vt_bytes, size_t scale_elems, Workspace** out,
                  std::unique_lock<std::mutex>* call_lock) {
  Workspace* wp;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    wp = &g_ws[dev];   // std::map nodes never move
  }
  Workspace& w = *wp;
  *call_lock = std::unique_lock<std::mutex>(w.call_mu);
  cudaError_t e;
  if (!w.error_flag) {
    if ((e = cudaMalloc(&w.error_flag, sizeof(int))) != cudaSuccess) return fail_cuda("cudaMalloc", e);
    cudaMemset(w.error_flag, 0, sizeof(int));
  }
  if (!w.cycles && getenv("QMHA_CYCLES")) {
    if ((e = cudaMalloc(&w.cycles, 2 * sizeof(unsigned long long))) != cudaSuccess) return fail_cuda("cudaMalloc", e);
    cudaMemset(w.cycles, 0, 2 * sizeof(unsigned long long));
  }
  if (qk_bytes > w.qk_bytes) {
    cudaDeviceSynchronize();
    cudaFree(w.Qp); cudaFree(w.Kp);
    w.Qp = w.Kp = nullptr; w.qk_bytes = 0;
    if ((e = cudaMalloc(&w.Qp, qk_bytes)) != cudaSuccess) return fail_cuda("cudaMalloc(Qp)", e);
    if ((e = cudaMalloc(&w.Kp, qk_bytes)) != cudaSuccess) return fail_cuda("cudaMalloc(Kp)", e);
    w.qk_bytes = qk_bytes;
  }
  if (vt_bytes > w.vt_bytes) {
    cudaDeviceSynchronize();
    cudaFree(w.Vt);
    w.Vt = nullptr; w.vt_bytes = 0;
    if ((e = cudaMalloc(&w.Vt, vt_bytes)) != cudaSuccess) return fail_cuda("cudaMalloc(Vt)", e);
    w.vt_bytes = vt_bytes;
  }
  if (scale_elems > w.scale_elems) {
    cudaDeviceSynchronize();
    cudaFree(w.scales); cudaFree(w.amax); cudaFree(w.aux); cudaFree(w.vmax);
    w.scales = nullptr; w.amax = nullptr; w.aux = nullptr; w.vmax = nullptr; w.scale_elems = 0;
    if ((e = cudaMalloc(&w.scales, scale_elems * sizeof(float))) != cudaSuccess) return fail_cuda("cudaMalloc(scales)", e);
    if ((e = cudaMalloc(&w.amax, scale_elems * sizeof(unsigned))) != cudaSuccess) return fail_cuda("cudaMalloc(amax)", e);
    if ((e = cudaMalloc(&w.aux, scale_elems * sizeof(float))) != cudaSuccess) return fail_cuda("cudaMalloc(aux)", e);
    if ((e = cudaMalloc(&w.vmax, scale_elems * sizeof(float))) != cudaSuccess) return fail_cuda("cudaMalloc(vmax)", e);
    w.scale_elems = scale_elems;
  }
  *out = &w;
  return 0;
}

int check_aligned16(const void* p, const char* name) {
  if ((reinterpret_cast<uintptr_t>(p) & 15) != 0)
    return fail(std::string(name) + " must be 16-byte aligned");
  return 0;
}

// Reads and clears the device-side error flag (set when an mbarrier wait timed out).
int check_error_flag(Workspace* w) {
  int flag = 0;
  cudaError_t e = cudaMemcpy(&flag, w->error_flag, sizeof(int), cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) return fail_cuda("reading kernel error flag", e);
  if (flag != 0) {
    cudaMemset(w->error_flag, 0, sizeof(int));
    return fail("attention kernel pipeline stalled at wait site " + std::to_string(flag));
  }
  return 0;
}

int resolve_default_kernel() {
  if (g_default_kernel == -2) {
    const char* env = getenv("QMHA_KERNEL");
    int k = qmha_kernel_from_name(env && *env ? env : QMHA_DEFAULT_KERNEL);
    g_default_kernel = k < 0 ? QMHA_KERNEL_INT8 : k;
  }
  return g_default_kernel;
}

size_t scale_count(int units, int n_pad, int gran) {
  return gran == QMHA_GRAN_BLOCK ? (size_t)3 * units * (n_pad / 32) : (size_t)3 * units;
}

bool rope_enabled() {
  if (g_rope < 0) {
    const char* env = getenv("QMHA_ROPE");
    g_rope = (env && *env && *env != '0') ? 1 : 0;
  }
  return g_rope == 1;
}

// {cos, sin} table of the reference's RoPE (utils/verify.cu:9-23: theta = powf(base, -2k/d),
// angle = pos * theta, sinf / cosf), computed with the host's libm — the same functions the CPU
// reference calls — so the fused rotation reproduces its fp32 values bit for bit.  Cached per
// device for the largest N seen with this (d, base).
int get_rope_table(int dev, int N, int d, float base, const float2** out) {
  std::lock_guard<std::mutex> lk(g_mu);
  Workspace& w = g_ws[dev];
  if (w.rope_tab && w.rope_d == d && w.rope_base == base && w.rope_n >= N) { *out = w.rope_tab; return 0; }
  const int half = d / 2;
  std::vector<float2> tab((size_t)N * half);
  std::vector<float> theta(half);
  for (int k = 0; k < half; ++k) theta[k] = powf(base, -static_cast<float>(2 * k) / d);
  for (int pos = 0; pos < N; ++pos)
    for (int k = 0; k < half; ++k) {
      const float angle = pos * theta[k];
      tab[(size_t)pos * half + k] = make_float2(cosf(angle), sinf(angle));
    }
  cudaDeviceSynchronize();
  cudaFree(w.rope_tab);
  w.rope_tab = nullptr; w.rope_n = 0;
  cudaError_t e = cudaMalloc(&w.rope_tab, tab.size() * sizeof(float2));
  if (e != cudaSuccess) return fail_cuda("cudaMalloc(rope table)", e);
  e = cudaMemcpy(w.rope_tab, tab.data(), tab.size() * sizeof(float2), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) return fail_cuda("uploading the rope table", e);
  w.rope_n = N; w.rope_d = d; w.rope_base = base;
  *out = w.rope_tab;
  return 0;
}

int prepare_impl(const float* Q, const float* K, const float* V, int B, int N, int d_model, int h,
                 int kernel, int gran, void* Qp, void* Kp, void* Vt, float* scales, unsigned* amax,
                 cudaStream_t stream) {
  int d, n_pad, d_pad;
  if (check_shape(B, N, d_model, h, &d, &n_pad, &d_pad)) return 1;
  if (check_aligned16(Q, "Q") || check_aligned16(K, "K") || check_aligned16(V, "V")) return 1;
  qmha::PrepareArgs a;
  a.Q = Q; a.K = K; a.V = V; a.scales = scales; a.Qp = Qp; a.Kp = Kp; a.Vt = Vt;
  a.B = B; a.N = N; a.H = h; a.d = d; a.n_pad = n_pad; a.d_pad = d_pad;
  a.int8 = kernel == QMHA_KERNEL_INT8;
  a.stream = stream;
  if (rope_enabled()) {
    if ((d & 7) != 0) return fail("fused RoPE needs a head dimension that is a multiple of 8");
    static const bool two_pass = getenv("QMHA_TWO_PASS_QUANT") != nullptr;
    if (a.int8 && !(gran == QMHA_GRAN_BLOCK || (gran == QMHA_GRAN_HEAD && !two_pass)))
      return fail("fused RoPE is implemented for QMHA_GRAN_BLOCK / QMHA_GRAN_HEAD (INT8) and for the F16 kernel");
    int dev = -1;
    cudaGetDevice(&dev);
    if (get_rope_table(dev, N, d, g_rope_base, &a.rope)) return 1;
  }
  cudaError_t e;
  if (a.int8) {
    if (gran == QMHA_GRAN_BLOCK) {
      // the reference's granularity: one scale per 32-row block, single pass over the inputs
      if ((d & 3) != 0) return fail("QMHA_GRAN_BLOCK needs a head dimension that is a multiple of 4");
      if ((e = qmha::launch_block_quantize(a)) != cudaSuccess) return fail_cuda("block quantise launch", e);
      g_launches += 1;
      return 0;
    }
    if (gran != QMHA_GRAN_TENSOR && gran != QMHA_GRAN_HEAD) return fail("unknown scale granularity");
    // Per-(batch, head) scales: one cluster kernel reads the inputs from HBM once.  Per-tensor
    // scales need a global maximum first and keep the two-pass path (as does an odd head dim).
    static const bool two_pass_env = getenv("QMHA_TWO_PASS_QUANT") != nullptr;
    if (gran == QMHA_GRAN_HEAD && (d & 3) == 0 && !two_pass_env) {
      if ((e = qmha::launch_fused_quantize(a)) != cudaSuccess) return fail_cuda("fused quantise launch", e);
      g_launches += 1;
      return 0;
    }
    if ((e = qmha::launch_absmax_and_scales(a, amax, gran == QMHA_GRAN_TENSOR)) != cudaSuccess)
      return fail_cuda("absmax launch", e);
    g_launches += 2;
  }
  if ((e = qmha::launch_prepare(a)) != cudaSuccess) return fail_cuda("prepare launch", e);
  g_launches += 1;
  return 0;
}

// Kernel variant k: exp2 of every k-th score pair on the FMA-pipe polynomial (0 = all MUFU).
// QMHA_ATTN_VARIANT overrides the built-in default (tuning / A-B measurements).
int attention_variant(int kernel) {
  const char* env = getenv("QMHA_ATTN_VARIANT");
  if (env && *env) return atoi(env);
  (void)kernel;
  return QMHA_DEFAULT_ATTN_VARIANT;
}

// scales: [3][B*h] (per-head / per-tensor) or, for gran == QMHA_GRAN_BLOCK, [3][B*h][n_pad/32];
// aux / vmax: scratch of the same element count used only in block mode.
int attention_impl(const void* Qp, const void* Kp, const void* Vt, const float* scales, float* O,
                   int B, int N, int d_model, int h, int kernel, int* error_flag,
                   cudaStream_t stream, long long* trace = nullptr, int variant = -1,
                   int gran = QMHA_GRAN_HEAD, float* aux = nullptr, float* vmax = nullptr,
                   unsigned long long* cycles = nullptr) {
  int d, n_pad, d_pad;
  if (check_shape(B, N, d_model, h, &d, &n_pad, &d_pad)) return 1;
  if (check_aligned16(O, "output")) return 1;
  qmha::AttnLaunch a;
  a.Qp = Qp; a.Kp = Kp; a.Vt = Vt; a.scales = scales; a.O = O; a.error_flag = error_flag;
  a.B = B; a.N = N; a.H = h; a.d = d; a.n_pad = n_pad; a.d_pad = d_pad;
  a.int8 = kernel == QMHA_KERNEL_INT8;
  a.stream = stream;
  a.trace = trace;
  a.cycles = cycles;
  a.variant = variant >= 0 ? variant : attention_variant(kernel);
  if (a.int8 && gran == QMHA_GRAN_BLOCK) {
    if (!aux || !vmax) return fail("internal: block mode needs scratch");
    const int units = B * h, nblk = n_pad / 32;
    cudaError_t e = qmha::launch_block_aux(scales + (size_t)2 * units * nblk, aux, vmax, units, nblk, stream);
    if (e != cudaSuccess) return fail_cuda("block aux launch", e);
    g_launches += 1;
    a.blk_scales = scales;
    a.blk_aux = aux;
    a.blk_vmax = vmax;
  }
  std::string err;
  if (!qmha::launch_attention(a, &err)) return fail(err);
  g_launches += 1;
  return 0;
}

}  // namespace

extern "C" {

const char* qmha_last_error(void) { return g_err.c_str(); }
const char* qmha_version(void) { return "quantizedmha_b200 0.1 (sm_100a)"; }
int64_t qmha_launch_count(void) { return g_launches.load(); }

int qmha_kernel_from_name(const char* name) {
  if (!name) return -1;
  const std::string n(name);
  if (n == "int8" || n == "fa_b200_int8" || n == "fa_tc_int8_a" || n == "fa_tc_int8_b" ||
      n == "fa_int8")
    return QMHA_KERNEL_INT8;
  if (n == "f16" || n == "fp16" || n == "fa_b200_f16" || n == "fa" || n == "unfused" ||
      n == "fa_tc_v1a" || n == "fa_tc_v1b" || n == "fa_tc_v2" || n == "fa_tc_v2a" ||
      n == "fa_tc_v2b" || n == "fa_tc" || n == "fa_warps")
    return QMHA_KERNEL_F16;
  return -1;
}

int qmha_set_kernel(const char* name) {
  int k = qmha_kernel_from_name(name);
  if (k < 0) return fail(std::string("unknown kernel name: ") + (name ? name : "(null)"));
  g_default_kernel = k;
  g_err.clear();
  return 0;
}

int qmha_set_rope(int enable, float base) {
  if (enable && !(base > 1.0f)) return fail("rope base must be > 1");
  g_rope = enable ? 1 : 0;
  if (enable) g_rope_base = base;
  g_err.clear();
  return 0;
}
int qmha_get_rope(void) { return rope_enabled() ? 1 : 0; }

int qmha_default_granularity(int d_model, int h) {
  const char* env = getenv("QMHA_SCALES");
  if (env && !strcmp(env, "head")) return QMHA_GRAN_HEAD;
  if (env && !strcmp(env, "tensor")) return QMHA_GRAN_TENSOR;
  if (h > 0 && d_model % h == 0 && ((d_model / h) & 3) == 0) return QMHA_GRAN_BLOCK;
  return QMHA_GRAN_HEAD;
}

const char* qmha_get_kernel(void) { return resolve_default_kernel() == QMHA_KERNEL_INT8 ? "int8" : "f16"; }

int qmha_workspace_dims(int N, int d_model, int h, int* n_pad, int* d_pad) {
  int d, np, dp;
  if (check_shape(1, N, d_model, h, &d, &np, &dp)) return 1;
  if (n_pad) *n_pad = np;
  if (d_pad) *d_pad = dp;
  g_err.clear();
  return 0;
}

int qmha_quantize_qkv(const float* Q, const float* K, const float* V, int B, int N, int d_model,
                      int h, int gran, int8_t* Qp, int8_t* Kp, uint16_t* Vt, float* scales,
                      void* stream) {
  const int dev = require_device();
  if (dev < 0) return 1;
  if (gran != QMHA_GRAN_TENSOR && gran != QMHA_GRAN_HEAD && gran != QMHA_GRAN_BLOCK) return fail("unknown scale granularity");
  Workspace* w;
  std::unique_lock<std::mutex> call_lock;
  if (get_workspace(dev, 0, 0, (size_t)3 * B * h, &w, &call_lock)) return 1;
  WorkspaceUse use;
  use.begin(w, std::move(call_lock), (cudaStream_t)stream);
  if (prepare_impl(Q, K, V, B, N, d_model, h, QMHA_KERNEL_INT8, gran, Qp, Kp, Vt, scales, w->amax,
                   (cudaStream_t)stream))
    return 1;
  g_err.clear();
  return 0;
}

int qmha_convert_qkv_f16(const float* Q, const float* K, const float* V, int B, int N, int d_model,
                         int h, uint16_t* Qp, uint16_t* Kp, uint16_t* Vt, void* stream) {
  if (require_device() < 0) return 1;
  if (prepare_impl(Q, K, V, B, N, d_model, h, QMHA_KERNEL_F16, QMHA_GRAN_HEAD, Qp, Kp, Vt, nullptr,
                   nullptr, (cudaStream_t)stream))
    return 1;
  g_err.clear();
  return 0;
}

int qmha_quantize_blocks(const float* X, int B, int N, int d_model, int h, int block_rows,
                         int8_t* q, float* scales, void* stream) {
  if (require_device() < 0) return 1;
  int d, n_pad, d_pad;
  if (B < 1 || N < 1 || h < 1 || d_model % h != 0 || block_rows < 1)
    return fail("invalid shape for qmha_quantize_blocks");
  (void)n_pad; (void)d_pad;
  d = d_model / h;
  cudaError_t e = qmha::launch_quantize_blocks(X, B, N, h, d, block_rows, q, scales,
                                               (cudaStream_t)stream);
  if (e != cudaSuccess) return fail_cuda("quantize_blocks launch", e);
  g_launches += 1;
  g_err.clear();
  return 0;
}

int qmha_quantize_static(const float* X, int64_t n, float scale, float zero_point, int8_t* q,
                         void* stream) {
  if (require_device() < 0) return 1;
  if (n < 0 || !(scale > 0.f)) return fail("invalid arguments for qmha_quantize_static");
  cudaError_t e = qmha::launch_quantize_static(X, n, scale, zero_point, q, (cudaStream_t)stream);
  if (e != cudaSuccess) return fail_cuda("quantize_static launch", e);
  g_launches += 1;
  g_err.clear();
  return 0;
}

int qmha_attention_prepared(const void* Qp, const void* Kp, const uint16_t* Vt, const float* scales,
                            float* O, int B, int N, int d_model, int h, int kernel, int gran,
                            void* stream) {
  const int dev = require_device();
  if (dev < 0) return 1;
  if (kernel != QMHA_KERNEL_INT8 && kernel != QMHA_KERNEL_F16) return fail("unknown kernel id");
  int d, n_pad, d_pad;
  if (check_shape(B, N, d_model, h, &d, &n_pad, &d_pad)) return 1;
  Workspace* w;
  std::unique_lock<std::mutex> call_lock;
  if (get_workspace(dev, 0, 0, scale_count(B * h, n_pad, gran), &w, &call_lock)) return 1;
  WorkspaceUse use;
  use.begin(w, std::move(call_lock), (cudaStream_t)stream);
  if (attention_impl(Qp, Kp, Vt, scales, O, B, N, d_model, h, kernel, w->error_flag,
                     (cudaStream_t)stream, nullptr, -1, gran, w->aux, w->vmax, w->cycles))
    return 1;
  g_err.clear();
  return 0;
}

int qmha_forward(const float* Q, const float* K, const float* V, float* O, int B, int N,
                 int d_model, int h, int kernel, int gran, void* stream) {
  const int dev = require_device();
  if (dev < 0) return 1;
  if (kernel != QMHA_KERNEL_INT8 && kernel != QMHA_KERNEL_F16) return fail("unknown kernel id");
  int d, n_pad, d_pad;
  if (check_shape(B, N, d_model, h, &d, &n_pad, &d_pad)) return 1;
  const size_t units = (size_t)B * h;
  const size_t elt = kernel == QMHA_KERNEL_INT8 ? 1 : 2;
  Workspace* w;
  std::unique_lock<std::mutex> call_lock;
  if (get_workspace(dev, units * n_pad * d_pad * elt, units * n_pad * d_pad * 2,
                    scale_count((int)units, n_pad, gran), &w, &call_lock))
    return 1;
  cudaStream_t s = (cudaStream_t)stream;
  WorkspaceUse use;   // orders this call behind earlier work on other streams that uses the workspace
  use.begin(w, std::move(call_lock), s);
  if (prepare_impl(Q, K, V, B, N, d_model, h, kernel, gran, w->Qp, w->Kp, w->Vt, w->scales, w->amax, s))
    return 1;
  if (attention_impl(w->Qp, w->Kp, w->Vt, w->scales, O, B, N, d_model, h, kernel, w->error_flag, s,
                     nullptr, -1, gran, w->aux, w->vmax))
    return 1;
  g_err.clear();
  return 0;
}

// Debug: runs the traced INT8 d=128 kernel once (synchronously) and copies the timeline of CTA
// in the middle of the grid to host_trace[9][ceil(N/64)][4] (clock64 stamps: softmax warps 0-7, MMA
// warp) followed by 8 phase stamps of the CTA (entry, setup, first scores, last P, O final, stores, exit).
int qmha_debug_attention_trace(const void* Qp, const void* Kp, const uint16_t* Vt,
                               const float* scales, float* O, int B, int N, int d_model, int h,
                               int variant, long long* host_trace) {
  const int dev = require_device();
  if (dev < 0) return 1;
  Workspace* w;
  std::unique_lock<std::mutex> call_lock;
  if (get_workspace(dev, 0, 0, 0, &w, &call_lock)) return 1;
  WorkspaceUse use;
  use.begin(w, std::move(call_lock), nullptr);
  const size_t n = (size_t)9 * ((N + 63) / 64) * 4 + 16;
  long long* dtrace = nullptr;
  cudaError_t e = cudaMalloc(&dtrace, n * sizeof(long long));
  if (e != cudaSuccess) return fail_cuda("cudaMalloc(trace)", e);
  cudaMemset(dtrace, 0, n * sizeof(long long));
  int rc = attention_impl(Qp, Kp, Vt, scales, O, B, N, d_model, h, QMHA_KERNEL_INT8, w->error_flag,
                          nullptr, dtrace, variant);
  if (rc == 0) {
    e = cudaDeviceSynchronize();
    if (e != cudaSuccess) rc = fail_cuda("trace run", e);
    else cudaMemcpy(host_trace, dtrace, n * sizeof(long long), cudaMemcpyDeviceToHost);
  }
  cudaFree(dtrace);
  if (rc == 0) g_err.clear();
  return rc;
}

// Development aid (QMHA_CYCLES=1 in the environment): SM clocks summed over the CTAs of every
// qmha_attention_prepared() launch since the last reset, and the number of CTAs.  Synchronises.
int qmha_debug_cycles(unsigned long long* out2, int reset) {
  const int dev = require_device();
  if (dev < 0) return 1;
  Workspace* w;
  std::unique_lock<std::mutex> call_lock;
  if (get_workspace(dev, 0, 0, 0, &w, &call_lock)) return 1;
  if (!w->cycles) return fail("set QMHA_CYCLES=1 before the first call");
  cudaError_t e = cudaMemcpy(out2, w->cycles, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) return fail_cuda("reading the cycle counters", e);
  if (reset) cudaMemset(w->cycles, 0, 2 * sizeof(unsigned long long));
  g_err.clear();
  return 0;
}

// Checks the asynchronous failure flag of the current device after the caller synchronised.
int qmha_check_async_error(void) {
  const int dev = require_device();
  if (dev < 0) return 1;
  Workspace* w;
  std::unique_lock<std::mutex> call_lock;
  if (get_workspace(dev, 0, 0, 0, &w, &call_lock)) return 1;
  if (check_error_flag(w)) return 1;
  g_err.clear();
  return 0;
}

void solve(const float* Q, const float* K, const float* V, float* output, int N, int d_model,
           int h) {
  // Synchronous on return like the reference (launchers.h:64); errors are recorded, reported on
  // stderr and queryable with qmha_last_error() — solve() itself stays void.
  const int kernel = resolve_default_kernel();
  // Scales at the reference's own granularity (one per 32-row tile, fa_tc_int8_b.cu:484-518) when the
  // head dimension allows the vectorised single-pass quantiser, per (batch, head) otherwise.
  int rc = qmha_forward(Q, K, V, output, 1, N, d_model, h, kernel, qmha_default_granularity(d_model, h), nullptr);
  if (rc == 0) {
    cudaError_t e = cudaStreamSynchronize(nullptr);
    if (e != cudaSuccess) rc = fail_cuda("solve", e);
    else rc = qmha_check_async_error();
  }
  if (rc != 0) fprintf(stderr, "qmha solve() failed: %s\n", g_err.c_str());
}

int qmha_forward_host(const float* Q, const float* K, const float* V, float* O, int B, int N,
                      int d_model, int h, int kernel, int gran) {
  const int dev = require_device();
  if (dev < 0) return 1;
  int d, n_pad, d_pad;
  if (check_shape(B, N, d_model, h, &d, &n_pad, &d_pad)) return 1;
  if (kernel == QMHA_KERNEL_INT8 && gran == QMHA_GRAN_TENSOR)
    return fail("qmha_forward_host pipelines over batches; use QMHA_GRAN_HEAD scales");
  // Pipeline over batch entries: H2D(b+1) overlaps compute(b) overlaps D2H(b-1).  Two slots of
  // device staging; each slot has its own stream so copies and kernels of different slots overlap.
  struct Slot { float *q = nullptr, *k = nullptr, *v = nullptr, *o = nullptr; cudaStream_t s = nullptr; };
  static std::map<int, std::pair<size_t, std::vector<Slot>>> staging;  // per device
  const size_t slab = (size_t)N * d_model;  // elements per batch entry
  cudaError_t e;
  std::vector<Slot>* slots;
  // Each slot needs its own operand workspace region: the slot streams run against disjoint halves.
  // The workspace's call lock is held for the whole (synchronous) call: it also guards the staging slots.
  const size_t units1 = (size_t)h;
  const size_t elt = kernel == QMHA_KERNEL_INT8 ? 1 : 2;
  const size_t qk1 = units1 * n_pad * d_pad * elt, vt1 = units1 * n_pad * d_pad * 2;
  const size_t sc1 = scale_count((int)units1, n_pad, gran);
  Workspace* w;
  std::unique_lock<std::mutex> call_lock;
  if (get_workspace(dev, 2 * qk1, 2 * vt1, 2 * sc1, &w, &call_lock)) return 1;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    auto& st = staging[dev];
    if (st.second.empty()) st.second.resize(2);
    if (st.first < slab) {
      cudaDeviceSynchronize();
      for (auto& sl : st.second) {
        cudaFree(sl.q); cudaFree(sl.k); cudaFree(sl.v); cudaFree(sl.o);
        sl.q = sl.k = sl.v = sl.o = nullptr;
        if ((e = cudaMalloc(&sl.q, slab * 4)) != cudaSuccess || (e = cudaMalloc(&sl.k, slab * 4)) != cudaSuccess ||
            (e = cudaMalloc(&sl.v, slab * 4)) != cudaSuccess || (e = cudaMalloc(&sl.o, slab * 4)) != cudaSuccess)
          return fail_cuda("cudaMalloc(staging)", e);
        if (!sl.s && (e = cudaStreamCreateWithFlags(&sl.s, cudaStreamNonBlocking)) != cudaSuccess)
          return fail_cuda("cudaStreamCreate", e);
      }
      st.first = slab;
    }
    slots = &st.second;
  }
  // earlier asynchronous calls (qmha_forward on a caller stream) may still be using the workspace
  if (w->in_flight)
    for (auto& sl : *slots) cudaStreamWaitEvent(sl.s, w->last_use, 0);
  for (int b = 0; b < B; ++b) {
    Slot& sl = (*slots)[b & 1];
    const size_t off = (size_t)b * slab;
    const int half = b & 1;
    // stream order on sl.s serialises reuse of this slot (b-2's D2H precedes b's H2D).
    if ((e = cudaMemcpyAsync(sl.q, Q + off, slab * 4, cudaMemcpyHostToDevice, sl.s)) != cudaSuccess ||
        (e = cudaMemcpyAsync(sl.k, K + off, slab * 4, cudaMemcpyHostToDevice, sl.s)) != cudaSuccess ||
        (e = cudaMemcpyAsync(sl.v, V + off, slab * 4, cudaMemcpyHostToDevice, sl.s)) != cudaSuccess)
      return fail_cuda("H2D copy", e);
    void* Qp = (char*)w->Qp + half * qk1;
    void* Kp = (char*)w->Kp + half * qk1;
    void* Vt = (char*)w->Vt + half * vt1;
    float* sc = w->scales + half * sc1;
    unsigned* am = w->amax + half * sc1;
    if (prepare_impl(sl.q, sl.k, sl.v, 1, N, d_model, h, kernel, gran, Qp, Kp, Vt, sc, am, sl.s)) return 1;
    if (attention_impl(Qp, Kp, Vt, sc, sl.o, 1, N, d_model, h, kernel, w->error_flag, sl.s, nullptr, -1,
                       gran, w->aux + half * sc1, w->vmax + half * sc1))
      return 1;
    if ((e = cudaMemcpyAsync(O + off, sl.o, slab * 4, cudaMemcpyDeviceToHost, sl.s)) != cudaSuccess)
      return fail_cuda("D2H copy", e);
  }
  for (auto& sl : *slots)
    if ((e = cudaStreamSynchronize(sl.s)) != cudaSuccess) return fail_cuda("forward_host sync", e);
  w->in_flight = false;   // everything that used the workspace, this call's and earlier work, has completed
  if (check_error_flag(w)) return 1;
  g_err.clear();
  return 0;
}

void qmha_shutdown(void) {
  std::lock_guard<std::mutex> lk(g_mu);
  int cur = 0;
  cudaGetDevice(&cur);
  for (auto& kv : g_ws) {
    cudaSetDevice(kv.first);
    Workspace& w = kv.second;
    cudaFree(w.Qp); cudaFree(w.Kp); cudaFree(w.Vt); cudaFree(w.scales); cudaFree(w.amax); cudaFree(w.aux); cudaFree(w.vmax);
    cudaFree(w.error_flag); cudaFree(w.rope_tab); cudaFree(w.cycles);
    if (w.last_use) cudaEventDestroy(w.last_use);
  }
  g_ws.clear();
  cudaSetDevice(cur);
}

}  // extern "C"
