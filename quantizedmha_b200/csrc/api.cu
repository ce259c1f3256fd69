// api.cu — the C-ABI of include/qmha.h: solve() and the extended entry points, per-device
// workspaces, error reporting.  Host code only; kernels live in prepare.cu and attn_fwd.cu.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include <cuda.h>
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
std::atomic<unsigned> g_launch_seq{0};
std::mutex g_mu;
// Process-wide defaults (only what solve() and the *_default paths use; every extended entry can override
// them per call through qmha_args).  Atomics: they are read by concurrent callers.
std::atomic<int> g_default_kernel{-2};  // -2 = not resolved yet
// Fused RoPE on Q and K inside the quantise / convert pass (qmha_set_rope; QMHA_ROPE=1 in the
// environment turns it on for processes that only call solve(), e.g. bin/profile_* --rope).
std::atomic<int> g_rope{-1};            // -1 = not resolved yet (environment), 0 = off, 1 = on
std::atomic<float> g_rope_base{10000.0f};

int fail(const std::string& msg) {
  g_err = msg;
  return 1;
}
int fail_cuda(const char* what, cudaError_t e) {
  return fail(std::string(what) + ": " + cudaGetErrorString(e));
}

struct RopeTable {
  float2* dev = nullptr;
  int n = 0;
};

// Staging of qmha_forward_host: one slot = device copies of one (batch entry, head group) chunk of Q, K, V
// and O plus the stream that carries its H2D copy, kernels and D2H copy.
struct HostSlot {
  float *q = nullptr, *k = nullptr, *v = nullptr, *o = nullptr;
  cudaStream_t s = nullptr;
};
constexpr int kHostSlots = 3;

struct Workspace {
  void* Qp = nullptr;
  void* Kp = nullptr;
  void* Vt = nullptr;
  float* scales = nullptr;
  unsigned* amax = nullptr;
  float* aux = nullptr;    // block mode: [units][n_pad/32][2]
  float* vmax = nullptr;   // block mode: [units]
  int* error_flag = nullptr;        // device: (launch id << 12) | wait site of the last stalled launch
  int* error_host = nullptr;        // mapped host word: wait site, written by the failing CTA
  int* error_host_dev = nullptr;    // its device address
  unsigned long long* cycles = nullptr;  // {sum of CTA residency clocks, CTAs}: qmha_debug_cycles
  size_t qk_bytes = 0, vt_bytes = 0, scale_elems = 0;
  // {cos, sin}[n][d/2] tables keyed by (d, base bits); grown tables replace their predecessor only in the
  // map — the old allocation stays alive until qmha_shutdown (a launch already enqueued may still read it).
  std::map<std::pair<int, uint32_t>, RopeTable> rope;
  std::vector<float2*> rope_retired;
  HostSlot host_slots[kHostSlots];
  size_t host_slot_elems = 0;
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
    // host-visible copy of the failure record: checked (without synchronising) at the start of every call
    if (cudaHostAlloc(&w.error_host, sizeof(int), cudaHostAllocMapped) == cudaSuccess) {
      *w.error_host = 0;
      if (cudaHostGetDevicePointer(&w.error_host_dev, w.error_host, 0) != cudaSuccess) w.error_host_dev = nullptr;
    } else {
      w.error_host = nullptr;
      cudaGetLastError();
    }
  }
  if (!w.cycles && getenv("QMHA_CYCLES")) {
    if ((e = cudaMalloc(&w.cycles, qmha::kCycleWords * sizeof(unsigned long long))) != cudaSuccess) return fail_cuda("cudaMalloc", e);
    cudaMemset(w.cycles, 0, qmha::kCycleWords * sizeof(unsigned long long));
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
    if ((e = cudaMalloc(&w.amax, (2 * scale_elems + 64) * sizeof(unsigned))) != cudaSuccess) return fail_cuda("cudaMalloc(amax)", e);
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

std::string stall_message(int site) {
  return "attention kernel pipeline stalled at wait site " + std::to_string(site) +
         ": the output of that launch is incomplete";
}

// A launch that stalled leaves its record in the mapped host word: every entry point looks at it first
// (no synchronisation), fails THAT call and clears the record — a stall cannot pass unnoticed just
// because the caller never asks qmha_check_async_error().  Later launches are not affected by the
// record (the kernel compares launch ids), so the call after the failing one runs normally.
int fail_on_recorded_stall(Workspace* w) {
  if (!w->error_host) return 0;
  const int site = *(volatile int*)w->error_host;
  if (site == 0) return 0;
  *(volatile int*)w->error_host = 0;
  cudaMemsetAsync(w->error_flag, 0, sizeof(int), nullptr);
  return fail("an earlier launch on this device failed: " + stall_message(site));
}

// Reads and clears the device-side error record after the caller synchronised.
int check_error_flag(Workspace* w) {
  int flag = 0;
  cudaError_t e = cudaMemcpy(&flag, w->error_flag, sizeof(int), cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) return fail_cuda("reading kernel error flag", e);
  if (flag != 0) {
    cudaMemset(w->error_flag, 0, sizeof(int));
    if (w->error_host) *(volatile int*)w->error_host = 0;
    return fail(stall_message(flag & 0xFFF));
  }
  return 0;
}

unsigned next_launch_id() {
  unsigned id;
  do { id = (g_launch_seq.fetch_add(1) + 1) & 0xFFFFFu; } while (id == 0);
  return id;
}

int resolve_default_kernel() {
  int k = g_default_kernel.load();
  if (k == -2) {
    const char* env = getenv("QMHA_KERNEL");
    k = qmha_kernel_from_name(env && *env ? env : QMHA_DEFAULT_KERNEL);
    if (k < 0) k = QMHA_KERNEL_INT8;
    g_default_kernel.store(k);
  }
  return k;
}

size_t scale_count(int units, int n_pad, int gran) {
  return gran == QMHA_GRAN_BLOCK ? (size_t)3 * units * (n_pad / 32) : (size_t)3 * units;
}

bool rope_default() {
  int r = g_rope.load();
  if (r < 0) {
    const char* env = getenv("QMHA_ROPE");
    r = (env && *env && *env != '0') ? 1 : 0;
    g_rope.store(r);
  }
  return r == 1;
}

// What solve() uses: the reference's own granularity (one scale per 32-row block) whenever the
// vectorised single-pass quantiser applies and the per-block table of one unit fits in shared memory
// behind the tiles (N <= ~68 k at d = 128); per (batch, head) otherwise.
int default_granularity(int N, int d_model, int h) {
  const char* env = getenv("QMHA_SCALES");
  if (env && !strcmp(env, "head")) return QMHA_GRAN_HEAD;
  if (env && !strcmp(env, "tensor")) return QMHA_GRAN_TENSOR;
  if (h > 0 && d_model % h == 0 && ((d_model / h) & 3) == 0) {
    const int dp = pad_head_dim(d_model / h);
    if (dp > 0 && (N <= 0 || round_up(N, 256) <= qmha::attention_max_block_keys(dp))) return QMHA_GRAN_BLOCK;
  }
  return QMHA_GRAN_HEAD;
}

// {cos, sin} table of the reference's RoPE (utils/verify.cu:9-23: theta = powf(base, -2k/d),
// angle = pos * theta, sinf / cosf), computed with the host's libm — the same functions the CPU
// reference calls — so the fused rotation reproduces its fp32 values bit for bit.  One table per
// (device, d, base), grown to the largest N seen.  The caller holds the workspace's call lock.
int get_rope_table(Workspace* w, int N, int d, float base, const float2** out) {
  uint32_t bits;
  memcpy(&bits, &base, sizeof(bits));
  RopeTable& t = w->rope[std::make_pair(d, bits)];
  if (t.dev && t.n >= N) { *out = t.dev; return 0; }
  const int half = d / 2;
  std::vector<float2> tab((size_t)N * half);
  std::vector<float> theta(half);
  for (int k = 0; k < half; ++k) theta[k] = powf(base, -static_cast<float>(2 * k) / d);
  for (int pos = 0; pos < N; ++pos)
    for (int k = 0; k < half; ++k) {
      const float angle = pos * theta[k];
      tab[(size_t)pos * half + k] = make_float2(cosf(angle), sinf(angle));
    }
  float2* dev = nullptr;
  cudaError_t e = cudaMalloc(&dev, tab.size() * sizeof(float2));
  if (e != cudaSuccess) return fail_cuda("cudaMalloc(rope table)", e);
  e = cudaMemcpy(dev, tab.data(), tab.size() * sizeof(float2), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) { cudaFree(dev); return fail_cuda("uploading the rope table", e); }
  if (t.dev) w->rope_retired.push_back(t.dev);   // may still be read by an enqueued launch
  t.dev = dev; t.n = N;
  *out = dev;
  return 0;
}

bool dtype_ok(int dt) { return dt == QMHA_DTYPE_F32 || dt == QMHA_DTYPE_F16 || dt == QMHA_DTYPE_BF16; }
bool kernel_ok(int k) { return k == QMHA_KERNEL_INT8 || k == QMHA_KERNEL_F16 || k == QMHA_KERNEL_BF16 || k == QMHA_KERNEL_INT8_PV8; }
bool is_int8(int k) { return k == QMHA_KERNEL_INT8 || k == QMHA_KERNEL_INT8_PV8; }

struct RopeOpt {
  bool on = false;
  float base = 10000.0f;
};
RopeOpt rope_from(int rope, float base) {   // -1 = process default
  RopeOpt r;
  if (rope < 0) { r.on = rope_default(); r.base = g_rope_base.load(); }
  else { r.on = rope != 0; r.base = base > 1.0f ? base : 10000.0f; }
  return r;
}

// The caller holds the call lock of `w` (the rope tables live in the workspace).
int prepare_impl(Workspace* w, const void* Q, const void* K, const void* V, int in_dtype, int B, int N,
                 int d_model, int h, int kernel, int gran, const RopeOpt& rope, void* Qp, void* Kp, void* Vt,
                 float* scales, unsigned* amax, cudaStream_t stream, long long in_ld = 0, long long in_bs = 0) {
  int d, n_pad, d_pad;
  if (check_shape(B, N, d_model, h, &d, &n_pad, &d_pad)) return 1;
  if (!dtype_ok(in_dtype)) return fail("unknown input dtype");
  if (check_aligned16(Q, "Q") || check_aligned16(K, "K") || check_aligned16(V, "V")) return 1;
  if (in_ld != 0 || in_bs != 0) {   // Q, K, V are slabs of larger tensors (same pitches for all three)
    const long long isz = in_dtype == QMHA_DTYPE_F32 ? 4 : 2;
    const long long ld = in_ld > 0 ? in_ld : d_model, bs = in_bs > 0 ? in_bs : (long long)N * ld;
    if (in_ld < 0 || in_bs < 0 || ld < d_model || bs < (long long)N * ld || ld > 0x7fffffffLL)
      return fail("input strides: in_row_stride >= d_model and in_batch_stride >= N * in_row_stride expected");
    if ((ld * isz) % 16 != 0 || (bs * isz) % 16 != 0) return fail("input row / batch strides must be multiples of 16 bytes");
  }
  if (in_dtype != QMHA_DTYPE_F32 && (d & 3) != 0)
    return fail("16-bit inputs need a head dimension that is a multiple of 4");
  qmha::PrepareArgs a;
  a.Q = Q; a.K = K; a.V = V; a.in_dtype = in_dtype; a.scales = scales; a.Qp = Qp; a.Kp = Kp; a.Vt = Vt;
  a.B = B; a.N = N; a.H = h; a.d = d; a.n_pad = n_pad; a.d_pad = d_pad;
  a.int8 = is_int8(kernel);
  a.bf16 = kernel == QMHA_KERNEL_BF16;
  a.v8 = kernel == QMHA_KERNEL_INT8_PV8;
  a.stream = stream;
  a.in_ld = in_ld; a.in_bs = in_bs;
  if (rope.on) {
    if ((d & 7) != 0) return fail("fused RoPE needs a head dimension that is a multiple of 8");
    if (get_rope_table(w, N, d, rope.base, &a.rope)) return 1;
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
    if (gran == QMHA_GRAN_HEAD && (d & 3) == 0 && !two_pass_env && !a.v8) {   // (INT8 P.V: block kernel or two-pass path)
      // cluster kernel by default; QMHA_STREAM_QUANT=1: the persistent-grid variant (same results, same speed)
      if (getenv("QMHA_STREAM_QUANT") == nullptr) {
        if ((e = qmha::launch_fused_quantize(a)) != cudaSuccess) return fail_cuda("fused quantise launch", e);
      } else {
        if ((e = qmha::launch_stream_quantize(a, amax)) != cudaSuccess) return fail_cuda("stream quantise launch", e);
      }
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

// Kernel variant k: exp2 of every k-th score pair on the FMA-pipe polynomial (0 = all MUFU; other values
// exist only in -DQMHA_BUILD_POLY builds).  QMHA_ATTN_VARIANT overrides the built-in default.
int attention_variant() {
  static const int v = [] {
    const char* env = getenv("QMHA_ATTN_VARIANT");
    return env && *env ? atoi(env) : QMHA_DEFAULT_ATTN_VARIANT;
  }();
  return v;
}

// scales: [3][B*h] (per-head / per-tensor) or, for gran == QMHA_GRAN_BLOCK, [3][B*h][n_pad/32];
// aux / vmax: scratch of the same element count used only in block mode.
// Where the output goes besides the dense [B, N, d_model] default: strides of a slab inside a larger tensor and
// replicas on peers (qmha_args.o_row_stride / o_batch_stride / peer_O).
struct OutPlace {
  long long ld = 0, bs = 0;
  int n_peers = 0;
  void* peers[QMHA_MAX_PEERS] = {};
};

int attention_impl(Workspace* w, const void* Qp, const void* Kp, const void* Vt, const float* scales, void* O,
                   int out_dtype, int B, int N, int d_model, int h, int kernel, cudaStream_t stream,
                   long long* trace = nullptr, int variant = -1, int gran = QMHA_GRAN_HEAD,
                   float* aux = nullptr, float* vmax = nullptr, unsigned long long* cycles = nullptr,
                   const OutPlace* place = nullptr) {
  int d, n_pad, d_pad;
  if (check_shape(B, N, d_model, h, &d, &n_pad, &d_pad)) return 1;
  if (!dtype_ok(out_dtype)) return fail("unknown output dtype");
  if (check_aligned16(O, "output")) return 1;
  qmha::AttnLaunch a;
  a.Qp = Qp; a.Kp = Kp; a.Vt = Vt; a.scales = scales; a.O = O; a.out_dtype = out_dtype;
  a.error_flag = w->error_flag; a.error_host = w->error_host_dev; a.launch_id = next_launch_id();
  a.B = B; a.N = N; a.H = h; a.d = d; a.n_pad = n_pad; a.d_pad = d_pad;
  a.int8 = is_int8(kernel);
  a.bf16 = kernel == QMHA_KERNEL_BF16;
  a.pv8 = kernel == QMHA_KERNEL_INT8_PV8;
  a.stream = stream;
  a.trace = trace;
  a.cycles = cycles;
  a.variant = variant >= 0 ? variant : attention_variant();
  if (place) {
    if (place->ld < 0 || place->bs < 0) return fail("negative output stride");
    if (place->ld > 0 && place->ld < d_model) return fail("o_row_stride is smaller than d_model");
    if (place->bs > 0 && place->bs < (long long)N * (place->ld > 0 ? place->ld : d_model))
      return fail("o_batch_stride is smaller than N rows");
    if (place->n_peers < 0 || place->n_peers > QMHA_MAX_PEERS) return fail("n_peers must be 0 .. QMHA_MAX_PEERS");
    a.o_ld = place->ld; a.o_bs = place->bs; a.n_peers = place->n_peers;
    for (int i = 0; i < place->n_peers; ++i) {
      if (!place->peers[i]) return fail("null peer output pointer");
      if (check_aligned16(place->peers[i], "peer output")) return 1;
      a.peer_O[i] = place->peers[i];
    }
  }
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

// Resolves the "-1 = default" fields of a qmha_args copy and validates it.
int resolve_args(qmha_args* a) {
  if (a->kernel < 0) a->kernel = resolve_default_kernel();
  if (!kernel_ok(a->kernel)) return fail("unknown kernel id");
  if (!dtype_ok(a->in_dtype) || !dtype_ok(a->out_dtype)) return fail("unknown dtype (QMHA_DTYPE_F32 / F16 / BF16)");
  if (a->gran < 0) a->gran = default_granularity(a->N, a->d_model, a->h);
  if (a->gran != QMHA_GRAN_TENSOR && a->gran != QMHA_GRAN_HEAD && a->gran != QMHA_GRAN_BLOCK)
    return fail("unknown scale granularity");
  return 0;
}

int forward_device(const qmha_args& a) {
  const int dev = require_device();
  if (dev < 0) return 1;
  int d, n_pad, d_pad;
  if (check_shape(a.B, a.N, a.d_model, a.h, &d, &n_pad, &d_pad)) return 1;
  const size_t units = (size_t)a.B * a.h;
  const size_t elt = is_int8(a.kernel) ? 1 : 2;
  Workspace* w;
  std::unique_lock<std::mutex> call_lock;
  if (get_workspace(dev, units * n_pad * d_pad * elt, units * n_pad * d_pad * 2,
                    scale_count((int)units, n_pad, a.gran), &w, &call_lock))
    return 1;
  if (fail_on_recorded_stall(w)) return 1;
  cudaStream_t s = (cudaStream_t)a.stream;
  WorkspaceUse use;   // orders this call behind earlier work on other streams that uses the workspace
  use.begin(w, std::move(call_lock), s);
  if (prepare_impl(w, a.Q, a.K, a.V, a.in_dtype, a.B, a.N, a.d_model, a.h, a.kernel, a.gran,
                   rope_from(a.rope, a.rope_base), w->Qp, w->Kp, w->Vt, w->scales, w->amax, s, a.in_row_stride,
                   a.in_batch_stride))
    return 1;
  OutPlace place;
  place.ld = a.o_row_stride; place.bs = a.o_batch_stride; place.n_peers = a.n_peers;
  for (int i = 0; i < QMHA_MAX_PEERS && i < a.n_peers; ++i) place.peers[i] = a.peer_O[i];
  if (attention_impl(w, w->Qp, w->Kp, w->Vt, w->scales, a.O, a.out_dtype, a.B, a.N, a.d_model, a.h, a.kernel, s,
                     nullptr, a.variant, a.gran, w->aux, w->vmax, nullptr, &place))
    return 1;
  g_err.clear();
  return 0;
}

qmha_args make_args(const void* Q, const void* K, const void* V, void* O, int B, int N, int d_model, int h,
                    int kernel, int gran, void* stream) {
  qmha_args a;
  memset(&a, 0, sizeof(a));
  a.struct_size = sizeof(a);
  a.Q = Q; a.K = K; a.V = V; a.O = O;
  a.B = B; a.N = N; a.d_model = d_model; a.h = h;
  a.kernel = kernel; a.gran = gran;
  a.in_dtype = a.out_dtype = QMHA_DTYPE_F32;
  a.rope = -1; a.rope_base = 0.f; a.variant = -1;
  a.stream = stream;
  a.device = -1;
  return a;
}

}  // namespace

extern "C" {

const char* qmha_last_error(void) { return g_err.c_str(); }
const char* qmha_version(void) { return "quantizedmha_b200 0.2 (sm_100a)"; }
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
  if (n == "bf16" || n == "fa_b200_bf16") return QMHA_KERNEL_BF16;
  if (n == "int8_pv8" || n == "fa_b200_int8_pv8") return QMHA_KERNEL_INT8_PV8;
  return -1;
}

int qmha_set_kernel(const char* name) {
  int k = qmha_kernel_from_name(name);
  if (k < 0) return fail(std::string("unknown kernel name: ") + (name ? name : "(null)"));
  g_default_kernel.store(k);
  g_err.clear();
  return 0;
}

int qmha_set_rope(int enable, float base) {
  if (enable && !(base > 1.0f)) return fail("rope base must be > 1");
  if (enable) g_rope_base.store(base);
  g_rope.store(enable ? 1 : 0);
  g_err.clear();
  return 0;
}
int qmha_get_rope(void) { return rope_default() ? 1 : 0; }

int qmha_default_granularity(int d_model, int h) { return default_granularity(0, d_model, h); }
int qmha_granularity_for(int N, int d_model, int h) { return default_granularity(N, d_model, h); }

const char* qmha_get_kernel(void) {
  const int k = resolve_default_kernel();
  return k == QMHA_KERNEL_INT8 ? "int8" : (k == QMHA_KERNEL_BF16 ? "bf16" : (k == QMHA_KERNEL_INT8_PV8 ? "int8_pv8" : "f16"));
}

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
  return qmha_quantize_qkv_ex(Q, K, V, QMHA_DTYPE_F32, B, N, d_model, h, gran, -1, 0.f, Qp, Kp, Vt, scales, stream);
}

int qmha_quantize_qkv_ex(const void* Q, const void* K, const void* V, int in_dtype, int B, int N, int d_model,
                         int h, int gran, int rope, float rope_base, int8_t* Qp, int8_t* Kp, uint16_t* Vt,
                         float* scales, void* stream) {
  return qmha_quantize_qkv_k(Q, K, V, in_dtype, B, N, d_model, h, QMHA_KERNEL_INT8, gran, rope, rope_base, Qp, Kp, Vt, scales, stream);
}

int qmha_quantize_qkv_k(const void* Q, const void* K, const void* V, int in_dtype, int B, int N, int d_model,
                        int h, int kernel, int gran, int rope, float rope_base, int8_t* Qp, int8_t* Kp, void* Vt,
                        float* scales, void* stream) {
  if (!is_int8(kernel)) return fail("qmha_quantize_qkv_k: kernel must be QMHA_KERNEL_INT8 or QMHA_KERNEL_INT8_PV8");
  const int dev = require_device();
  if (dev < 0) return 1;
  if (gran != QMHA_GRAN_TENSOR && gran != QMHA_GRAN_HEAD && gran != QMHA_GRAN_BLOCK) return fail("unknown scale granularity");
  Workspace* w;
  std::unique_lock<std::mutex> call_lock;
  if (get_workspace(dev, 0, 0, (size_t)3 * B * h, &w, &call_lock)) return 1;
  WorkspaceUse use;
  use.begin(w, std::move(call_lock), (cudaStream_t)stream);
  if (prepare_impl(w, Q, K, V, in_dtype, B, N, d_model, h, kernel, gran, rope_from(rope, rope_base),
                   Qp, Kp, Vt, scales, w->amax, (cudaStream_t)stream))
    return 1;
  g_err.clear();
  return 0;
}

int qmha_convert_qkv_16(const void* Q, const void* K, const void* V, int in_dtype, int B, int N, int d_model,
                        int h, int kernel, int rope, float rope_base, uint16_t* Qp, uint16_t* Kp, uint16_t* Vt,
                        void* stream) {
  const int dev = require_device();
  if (dev < 0) return 1;
  if (kernel != QMHA_KERNEL_F16 && kernel != QMHA_KERNEL_BF16) return fail("qmha_convert_qkv_16: kernel must be F16 or BF16");
  Workspace* w;   // the call lock guards the rope table cache
  std::unique_lock<std::mutex> call_lock;
  if (get_workspace(dev, 0, 0, 0, &w, &call_lock)) return 1;
  if (prepare_impl(w, Q, K, V, in_dtype, B, N, d_model, h, kernel, QMHA_GRAN_HEAD, rope_from(rope, rope_base),
                   Qp, Kp, Vt, nullptr, nullptr, (cudaStream_t)stream))
    return 1;
  g_err.clear();
  return 0;
}

int qmha_convert_qkv_f16(const float* Q, const float* K, const float* V, int B, int N, int d_model,
                         int h, uint16_t* Qp, uint16_t* Kp, uint16_t* Vt, void* stream) {
  return qmha_convert_qkv_16(Q, K, V, QMHA_DTYPE_F32, B, N, d_model, h, QMHA_KERNEL_F16, -1, 0.f, Qp, Kp, Vt, stream);
}

int qmha_quantize_blocks(const float* X, int B, int N, int d_model, int h, int block_rows,
                         int8_t* q, float* scales, void* stream) {
  if (require_device() < 0) return 1;
  if (B < 1 || N < 1 || h < 1 || d_model % h != 0 || block_rows < 1)
    return fail("invalid shape for qmha_quantize_blocks");
  const int d = d_model / h;
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
  return qmha_attention_prepared_ex(Qp, Kp, Vt, scales, O, QMHA_DTYPE_F32, B, N, d_model, h, kernel, gran, stream);
}

int qmha_attention_prepared_ex(const void* Qp, const void* Kp, const uint16_t* Vt, const float* scales,
                               void* O, int out_dtype, int B, int N, int d_model, int h, int kernel, int gran,
                               void* stream) {
  const int dev = require_device();
  if (dev < 0) return 1;
  if (!kernel_ok(kernel)) return fail("unknown kernel id");
  int d, n_pad, d_pad;
  if (check_shape(B, N, d_model, h, &d, &n_pad, &d_pad)) return 1;
  Workspace* w;
  std::unique_lock<std::mutex> call_lock;
  if (get_workspace(dev, 0, 0, scale_count(B * h, n_pad, gran), &w, &call_lock)) return 1;
  if (fail_on_recorded_stall(w)) return 1;
  WorkspaceUse use;
  use.begin(w, std::move(call_lock), (cudaStream_t)stream);
  if (attention_impl(w, Qp, Kp, Vt, scales, O, out_dtype, B, N, d_model, h, kernel, (cudaStream_t)stream,
                     nullptr, -1, gran, w->aux, w->vmax, w->cycles))
    return 1;
  g_err.clear();
  return 0;
}

int qmha_forward(const float* Q, const float* K, const float* V, float* O, int B, int N,
                 int d_model, int h, int kernel, int gran, void* stream) {
  qmha_args a = make_args(Q, K, V, O, B, N, d_model, h, kernel, gran, stream);
  if (!kernel_ok(kernel)) return fail("unknown kernel id");
  if (resolve_args(&a)) return 1;
  return forward_device(a);
}

int qmha_forward_ex(const qmha_args* args) {
  if (!args) return fail("qmha_forward_ex: null argument block");
  if (args->struct_size != sizeof(qmha_args))
    return fail("qmha_forward_ex: struct_size does not match this library's qmha_args (header / library mismatch)");
  qmha_args a = *args;
  if (resolve_args(&a)) return 1;
  if (a.device < 0) return forward_device(a);
  // explicit device ordinal: run the call with that device current and put the caller's device back
  int cur = 0, count = 0;
  if (cudaGetDevice(&cur) != cudaSuccess || cudaGetDeviceCount(&count) != cudaSuccess) {
    cudaGetLastError();
    return fail("no CUDA device (this library has no CPU fallback)");
  }
  if (a.device >= count) return fail("qmha_args.device: no such device");
  if (a.device == cur) return forward_device(a);
  if (cudaSetDevice(a.device) != cudaSuccess) { cudaGetLastError(); return fail("cudaSetDevice(qmha_args.device) failed"); }
  const int rc = forward_device(a);
  cudaSetDevice(cur);
  return rc;
}

void qmha_args_init(qmha_args* a) {
  if (!a) return;
  *a = make_args(nullptr, nullptr, nullptr, nullptr, 1, 0, 0, 0, -1, -1, nullptr);
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
  int rc = attention_impl(w, Qp, Kp, Vt, scales, O, QMHA_DTYPE_F32, B, N, d_model, h, QMHA_KERNEL_INT8,
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
  if (reset) cudaMemset(w->cycles, 0, qmha::kCycleWords * sizeof(unsigned long long));
  g_err.clear();
  return 0;
}

// Same counters per SM: out[2*s] = clocks between the first CTA start and the last CTA end on SM s since the last
// reset (0 if the SM ran nothing), out[2*s+1] unused; n_sms <= 192.  Meaningful for ONE launch between resets.
int qmha_debug_sm_spans(unsigned long long* out, int n_sms, int reset) {
  const int dev = require_device();
  if (dev < 0) return 1;
  Workspace* w;
  std::unique_lock<std::mutex> call_lock;
  if (get_workspace(dev, 0, 0, 0, &w, &call_lock)) return 1;
  if (!w->cycles) return fail("set QMHA_CYCLES=1 before the first call");
  if (n_sms < 0 || n_sms > qmha::kCycleSms) return fail("n_sms out of range");
  std::vector<unsigned long long> h(qmha::kCycleWords);
  cudaError_t e = cudaMemcpy(h.data(), w->cycles, qmha::kCycleWords * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) return fail_cuda("reading the cycle counters", e);
  for (int s = 0; s < n_sms; ++s) {
    const unsigned long long ns = h[2 + 2 * s], en = h[3 + 2 * s];
    out[2 * s] = en ? en - ~ns : 0ull;
    out[2 * s + 1] = 0ull;
  }
  if (reset) cudaMemset(w->cycles, 0, qmha::kCycleWords * sizeof(unsigned long long));
  g_err.clear();
  return 0;
}

// Test hook: plants a failure record exactly as a stalled CTA would (device flag of a launch id that is
// never used + the mapped host word), so the host-side reporting can be exercised without a hung kernel.
int qmha_debug_inject_stall(int site) {
  const int dev = require_device();
  if (dev < 0) return 1;
  Workspace* w;
  std::unique_lock<std::mutex> call_lock;
  if (get_workspace(dev, 0, 0, 0, &w, &call_lock)) return 1;
  const int flag = (int)((0xFFFFFu << 12) | (unsigned)(site & 0xFFF));
  cudaMemcpy(w->error_flag, &flag, sizeof(int), cudaMemcpyHostToDevice);
  if (w->error_host) *(volatile int*)w->error_host = site & 0xFFF;
  g_err.clear();
  return 0;
}

// Checks the asynchronous failure record of the current device after the caller synchronised.
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

// Waits for `stream` and reports a pipeline failure of the launches that ran on it.
int qmha_synchronize(void* stream) {
  if (require_device() < 0) return 1;
  cudaError_t e = cudaStreamSynchronize((cudaStream_t)stream);
  if (e != cudaSuccess) return fail_cuda("qmha_synchronize", e);
  return qmha_check_async_error();
}

void solve(const float* Q, const float* K, const float* V, float* output, int N, int d_model,
           int h) {
  // Synchronous on return like the reference (launchers.h:64); errors are recorded, reported on
  // stderr and queryable with qmha_last_error() — solve() itself stays void.
  // Scales at the reference's own granularity (one per 32-row tile, fa_tc_int8_b.cu:484-518) when the
  // head dimension allows the vectorised single-pass quantiser and the sequence fits the per-block table,
  // per (batch, head) otherwise.
  int rc = qmha_forward(Q, K, V, output, 1, N, d_model, h, resolve_default_kernel(), -1, nullptr);
  if (rc == 0) rc = qmha_synchronize(nullptr);
  if (rc != 0) fprintf(stderr, "qmha solve() failed: %s\n", g_err.c_str());
}

// Host buffers in, host buffer out.  The work is cut into (batch entry, head group) chunks — the reference's
// own call shape is B = 1, so batch entries alone would leave nothing to overlap — and chunk c+1's H2D copy
// overlaps chunk c's kernels and chunk c-1's D2H copy: kHostSlots staging slots, each with its own stream;
// a head group is a strided 2-D region of the [N, d_model] matrices (cudaMemcpy2DAsync), compact on the device.
int qmha_forward_host(const float* Q, const float* K, const float* V, float* O, int B, int N,
                      int d_model, int h, int kernel, int gran) {
  return qmha_forward_host_ex(Q, K, V, O, B, N, d_model, h, kernel, gran, QMHA_DTYPE_F32, QMHA_DTYPE_F32);
}

int qmha_forward_host_ex(const void* Qv, const void* Kv, const void* Vv, void* Ov, int B, int N, int d_model, int h,
                         int kernel, int gran, int in_dtype, int out_dtype) {
  const int dev = require_device();
  if (dev < 0) return 1;
  if (!dtype_ok(in_dtype) || !dtype_ok(out_dtype)) return fail("unknown dtype (QMHA_DTYPE_F32 / F16 / BF16)");
  const size_t isz = in_dtype == QMHA_DTYPE_F32 ? 4 : 2, osz = out_dtype == QMHA_DTYPE_F32 ? 4 : 2;
  const char* Q = static_cast<const char*>(Qv);
  const char* K = static_cast<const char*>(Kv);
  const char* V = static_cast<const char*>(Vv);
  char* O = static_cast<char*>(Ov);
  int d, n_pad, d_pad;
  if (check_shape(B, N, d_model, h, &d, &n_pad, &d_pad)) return 1;
  if (!kernel_ok(kernel)) return fail("unknown kernel id");
  if (gran < 0) gran = default_granularity(N, d_model, h);
  if (is_int8(kernel) && gran == QMHA_GRAN_TENSOR)
    return fail("qmha_forward_host pipelines over (batch, head group) chunks; use QMHA_GRAN_HEAD or QMHA_GRAN_BLOCK scales");
  // Head groups: enough chunks to overlap copies with compute (>= 4 per call when the heads allow it), rows of
  // at least 512 bytes per 2-D copy line, at most ~64 MB of input per tensor and chunk.
  int hg = h;
  {
    const size_t head_bytes = (size_t)N * d * isz;
    const int min_heads = std::max(1, (int)((512 + d * isz - 1) / (d * isz)));
    int want_chunks = std::max(1, (4 + B - 1) / B);                       // per batch entry
    int by_chunks = std::max(1, h / want_chunks);
    int by_bytes = std::max(1, (int)((size_t)64 * 1024 * 1024 / std::max<size_t>(head_bytes, 1)));
    hg = std::max(min_heads, std::min(by_chunks, by_bytes));
    hg = std::min(hg, h);
    while (h % hg != 0) --hg;   // equal groups keep one operand layout per slot
    const char* env = getenv("QMHA_HOST_HEAD_GROUP");
    if (env && atoi(env) > 0 && h % atoi(env) == 0) hg = atoi(env);
  }
  const int groups = h / hg;
  const int dmg = hg * d;                         // d_model of one chunk on the device
  const size_t chunk = (size_t)N * dmg;           // elements per tensor and chunk
  const size_t elt = is_int8(kernel) ? 1 : 2;
  const size_t qk1 = (size_t)hg * n_pad * d_pad * elt, vt1 = (size_t)hg * n_pad * d_pad * 2;
  const size_t sc1 = scale_count(hg, n_pad, gran);
  Workspace* w;
  std::unique_lock<std::mutex> call_lock;   // held for the whole (synchronous) call: it also guards the slots
  // (+32 per slot: the per-head quantiser's queue scratch is 16 + 6 * units words per slot, see amax below)
  if (get_workspace(dev, kHostSlots * qk1, kHostSlots * vt1, kHostSlots * (sc1 + 32), &w, &call_lock)) return 1;
  if (fail_on_recorded_stall(w)) return 1;
  cudaError_t e;
  if (w->host_slot_elems < chunk) {
    cudaDeviceSynchronize();
    w->host_slot_elems = 0;
    for (HostSlot& sl : w->host_slots) {
      cudaFree(sl.q); cudaFree(sl.k); cudaFree(sl.v); cudaFree(sl.o);
      sl.q = sl.k = sl.v = sl.o = nullptr;
    }
    for (HostSlot& sl : w->host_slots) {
      if ((e = cudaMalloc(&sl.q, chunk * 4)) != cudaSuccess || (e = cudaMalloc(&sl.k, chunk * 4)) != cudaSuccess ||
          (e = cudaMalloc(&sl.v, chunk * 4)) != cudaSuccess || (e = cudaMalloc(&sl.o, chunk * 4)) != cudaSuccess)
        return fail_cuda("cudaMalloc(staging)", e);   // host_slot_elems stays 0: the next call starts over
      if (!sl.s && (e = cudaStreamCreateWithFlags(&sl.s, cudaStreamNonBlocking)) != cudaSuccess)
        return fail_cuda("cudaStreamCreate", e);
    }
    w->host_slot_elems = chunk;
  }
  // earlier asynchronous calls (qmha_forward on a caller stream) may still be using the workspace
  if (w->in_flight)
    for (HostSlot& sl : w->host_slots) cudaStreamWaitEvent(sl.s, w->last_use, 0);
  const RopeOpt rope = rope_from(-1, 0.f);
  const size_t pitch_in = (size_t)d_model * isz, width_in = (size_t)dmg * isz;
  const size_t pitch_out = (size_t)d_model * osz, width_out = (size_t)dmg * osz;
  int c = 0;
  for (int b = 0; b < B; ++b) {
    for (int g = 0; g < groups; ++g, ++c) {
      const int si = c % kHostSlots;
      HostSlot& sl = w->host_slots[si];
      const size_t off = (size_t)b * N * d_model + (size_t)g * dmg;   // elements
      // stream order on sl.s serialises the reuse of this slot (chunk c-kHostSlots' D2H precedes c's H2D)
      if ((e = cudaMemcpy2DAsync(sl.q, width_in, Q + off * isz, pitch_in, width_in, N, cudaMemcpyHostToDevice, sl.s)) != cudaSuccess ||
          (e = cudaMemcpy2DAsync(sl.k, width_in, K + off * isz, pitch_in, width_in, N, cudaMemcpyHostToDevice, sl.s)) != cudaSuccess ||
          (e = cudaMemcpy2DAsync(sl.v, width_in, V + off * isz, pitch_in, width_in, N, cudaMemcpyHostToDevice, sl.s)) != cudaSuccess)
        return fail_cuda("H2D copy", e);
      void* Qp = (char*)w->Qp + si * qk1;
      void* Kp = (char*)w->Kp + si * qk1;
      void* Vt = (char*)w->Vt + si * vt1;
      float* sc = w->scales + si * sc1;
      unsigned* am = w->amax + si * (2 * sc1 + 64);
      if (prepare_impl(w, sl.q, sl.k, sl.v, in_dtype, 1, N, dmg, hg, kernel, gran, rope, Qp, Kp, Vt, sc, am, sl.s))
        return 1;
      if (attention_impl(w, Qp, Kp, Vt, sc, sl.o, out_dtype, 1, N, dmg, hg, kernel, sl.s, nullptr, -1, gran,
                         w->aux + si * sc1, w->vmax + si * sc1))
        return 1;
      if ((e = cudaMemcpy2DAsync(O + off * osz, pitch_out, sl.o, width_out, width_out, N, cudaMemcpyDeviceToHost, sl.s)) != cudaSuccess)
        return fail_cuda("D2H copy", e);
    }
  }
  for (HostSlot& sl : w->host_slots)
    if ((e = cudaStreamSynchronize(sl.s)) != cudaSuccess) return fail_cuda("forward_host sync", e);
  w->in_flight = false;   // everything that used the workspace, this call's and earlier work, has completed
  if (check_error_flag(w)) return 1;
  g_err.clear();
  return 0;
}

// ---- peer memory: CUDA IPC mappings of other ranks' output tensors (qmha_args.peer_O) ----------------------
namespace {
struct IpcMapping { int dev; void* base; size_t size; int refs; };
std::map<std::string, IpcMapping> g_ipc;   // key: device number + the 64 handle bytes
typedef CUresult (*GetAddressRangeFn)(CUdeviceptr*, size_t*, CUdeviceptr);
GetAddressRangeFn get_address_range_fn() {
  static GetAddressRangeFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<GetAddressRangeFn>(p);
  }
  return fn;
}
int close_ipc_locked() {
  int rc = 0, cur = 0;
  cudaGetDevice(&cur);
  for (auto& kv : g_ipc) {
    cudaSetDevice(kv.second.dev);
    if (cudaIpcCloseMemHandle(kv.second.base) != cudaSuccess) rc = 1;
  }
  g_ipc.clear();
  cudaSetDevice(cur);
  cudaGetLastError();
  return rc;
}
}  // namespace

int qmha_ipc_export(const void* dev_ptr, unsigned char handle[64], int64_t* offset) {
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
  if (!dev_ptr || !handle || !offset) return fail("qmha_ipc_export: null argument");
  if (require_device() < 0) return 1;
  GetAddressRangeFn range = get_address_range_fn();
  if (!range) return fail("cuMemGetAddressRange entry point not available");
  CUdeviceptr base = 0;
  size_t size = 0;
  if (range(&base, &size, (CUdeviceptr)(uintptr_t)dev_ptr) != CUDA_SUCCESS)
    return fail("qmha_ipc_export: not a device allocation");
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, reinterpret_cast<void*>((uintptr_t)base));
  if (e != cudaSuccess) cudaGetLastError();
  if (e != cudaSuccess) return fail_cuda("cudaIpcGetMemHandle (allocations from expandable segments / cudaMallocAsync pools cannot be exported)", e);
  memcpy(handle, &h, 64);
  *offset = (int64_t)((uintptr_t)dev_ptr - (uintptr_t)base);
  g_err.clear();
  return 0;
}

int qmha_ipc_open(const unsigned char handle[64], int64_t offset, void** dev_ptr) {
  if (!handle || !dev_ptr || offset < 0) return fail("qmha_ipc_open: bad argument");
  const int dev = require_device();
  if (dev < 0) return 1;
  std::lock_guard<std::mutex> lk(g_mu);
  std::string key(reinterpret_cast<const char*>(handle), 64);
  key.push_back((char)dev);
  auto it = g_ipc.find(key);
  if (it == g_ipc.end()) {
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    void* base = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) { cudaGetLastError(); return fail_cuda("cudaIpcOpenMemHandle", e); }
    size_t size = 0;
    CUdeviceptr b0 = 0;
    GetAddressRangeFn range = get_address_range_fn();
    if (range) range(&b0, &size, (CUdeviceptr)(uintptr_t)base);
    it = g_ipc.emplace(key, IpcMapping{dev, base, size, 0}).first;
  }
  it->second.refs += 1;
  *dev_ptr = static_cast<char*>(it->second.base) + offset;
  g_err.clear();
  return 0;
}

int qmha_ipc_close(void* dev_ptr) {
  std::lock_guard<std::mutex> lk(g_mu);
  for (auto it = g_ipc.begin(); it != g_ipc.end(); ++it) {
    IpcMapping& m = it->second;
    const char* b = static_cast<const char*>(m.base);
    const char* p = static_cast<const char*>(dev_ptr);
    if (p < b || (m.size ? p >= b + m.size : p != b)) continue;
    if (--m.refs > 0) return 0;
    int cur = 0;
    cudaGetDevice(&cur);
    cudaSetDevice(m.dev);
    const cudaError_t e = cudaIpcCloseMemHandle(m.base);
    cudaSetDevice(cur);
    g_ipc.erase(it);
    if (e != cudaSuccess) { cudaGetLastError(); return fail_cuda("cudaIpcCloseMemHandle", e); }
    return 0;
  }
  return fail("qmha_ipc_close: not a pointer returned by qmha_ipc_open");
}

int qmha_ipc_close_all(void) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (close_ipc_locked()) return fail("cudaIpcCloseMemHandle failed");
  return 0;
}

int qmha_enable_peer_access(int dev, int peer) {
  if (require_device() < 0) return 1;
  if (dev == peer) return 0;
  int can = 0;
  cudaError_t e = cudaDeviceCanAccessPeer(&can, dev, peer);
  if (e != cudaSuccess) return fail_cuda("cudaDeviceCanAccessPeer", e);
  if (!can) return fail("device " + std::to_string(dev) + " cannot access device " + std::to_string(peer) + " directly");
  int cur = 0;
  cudaGetDevice(&cur);
  cudaSetDevice(dev);
  e = cudaDeviceEnablePeerAccess(peer, 0);
  cudaSetDevice(cur);
  if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); e = cudaSuccess; }
  if (e != cudaSuccess) return fail_cuda("cudaDeviceEnablePeerAccess", e);
  g_err.clear();
  return 0;
}

void qmha_shutdown(void) {
  std::lock_guard<std::mutex> lk(g_mu);
  close_ipc_locked();
  int cur = 0;
  cudaGetDevice(&cur);
  for (auto& kv : g_ws) {
    cudaSetDevice(kv.first);
    Workspace& w = kv.second;
    cudaDeviceSynchronize();
    cudaFree(w.Qp); cudaFree(w.Kp); cudaFree(w.Vt); cudaFree(w.scales); cudaFree(w.amax); cudaFree(w.aux); cudaFree(w.vmax);
    cudaFree(w.error_flag); cudaFree(w.cycles);
    if (w.error_host) cudaFreeHost(w.error_host);
    for (auto& t : w.rope) cudaFree(t.second.dev);
    for (float2* p : w.rope_retired) cudaFree(p);
    for (HostSlot& sl : w.host_slots) {
      cudaFree(sl.q); cudaFree(sl.k); cudaFree(sl.v); cudaFree(sl.o);
      if (sl.s) cudaStreamDestroy(sl.s);
    }
    if (w.last_use) cudaEventDestroy(w.last_use);
  }
  g_ws.clear();
  cudaSetDevice(cur);
}

}  // extern "C"
