// attn_fwd.cuh — host-visible interface of the tcgen05 attention kernel (attn_fwd.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>

namespace qmha {

constexpr int kCycleSms = 192;                     // per-SM slots of the development cycle counters
constexpr int kCycleWords = 2 + 2 * kCycleSms;     // {sum, CTAs, per SM: ~earliest start, latest end}
constexpr int kMaxPeers = 7;                       // further destinations of the output (qmha_args.peer_O)

struct AttnParams {
  void* O;              // [B, N, H*d] fp32 / fp16 / bf16 (out_dtype)
  const float* scales;  // [3, units] (INT8 variant) or nullptr
  int* error_flag;      // device int: 0 = ok, otherwise (launch id << 12) | wait site that timed out
  int* error_host;      // mapped host copy of the failing wait site (nullptr = none)
  unsigned launch_id;   // 20-bit id of this launch (never 0)
  int out_dtype;        // 0 = fp32, 1 = fp16, 2 = bf16
  const float* blk_scales;  // block mode: [3][units][n_pad/32] raw scales (Q, K, V), else nullptr
  const float* blk_aux;     // block mode: [units][n_pad/32][2] = {log2 r, 1/r}, r = sV_block/sV_max
  const float* blk_vmax;    // block mode: [units] largest V block scale
  long long* trace;     // debug timeline buffer [3 roles][n_half_steps][4] (traced build only)
  int B, N, H, d;
  int n_pad;            // padded sequence length of the prepared operands (multiple of 256)
  int units;            // B*H
  int n_kv_tiles;       // ceil(N / 128): K/V tiles staged by TMA
  int n_half_steps;     // ceil(N / 64): 64-key softmax/MMA half-steps
  float scale_log2;     // log2(e) / sqrt(d)
  int debug_no_mma;     // measurement aid: skip every tcgen05.mma (results are garbage)
  int tma_store;        // epilogue writes the output with TMA tensor stores (needs d % 32 == 0)
  int one;              // 1 (a value the compiler cannot fold; see i2f_magic)
  unsigned long long* cycles;  // development aid: {sum of CTA residency clocks, CTA count} or nullptr
  long long o_ld, o_bs; // distance between consecutive rows / batch entries of O in elements (dense: H*d, N*H*d)
  int n_qblocks;        // 256-row query blocks per unit
  int n_items;          // work items (unit, query block) of the launch: n_qblocks * units (persistent kernel)
  unsigned stagger_ns;  // replicated output: odd CTAs of the first wave start this much later (0 = off) ...
  unsigned stagger_ctas;   // ... "first wave" = linear CTA index below this
  int n_peers;          // further destinations that receive the same bytes (replicas on NVLink peers)
  void* peer_O[kMaxPeers];
};

struct AttnLaunch {
  const void* Qp;       // [units, n_pad, d_pad] int8 or fp16
  const void* Kp;       // [units, n_pad, d_pad]
  const void* Vt;       // [units, d_pad, n_pad] fp16
  const float* scales;
  void* O;
  int out_dtype = 0;    // 0 = fp32, 1 = fp16, 2 = bf16
  int* error_flag;
  int* error_host = nullptr;
  unsigned launch_id = 1;
  const float* blk_scales = nullptr;  // non-null selects the per-32-row-block scale kernel (INT8)
  const float* blk_aux = nullptr;
  const float* blk_vmax = nullptr;
  long long* trace = nullptr;  // device buffer; non-null selects the traced instantiation
  unsigned long long* cycles = nullptr;  // development aid (qmha_debug_cycles)
  int variant = 0;             // k > 0: exp2 of every k-th score pair on the FMA-pipe polynomial
  int B, N, H, d, n_pad, d_pad;
  bool int8;
  bool bf16 = false;    // 16-bit kernel with bf16 operands (Qp / Kp / Vt hold bf16)
  bool pv8 = false;     // INT8 kernel with INT8 P.V: Vt holds int8 codes [units, d_pad, n_pad], P goes to the MMA as 8-bit codes
  cudaStream_t stream;
  long long o_ld = 0, o_bs = 0;   // 0 = dense
  int n_peers = 0;
  void* peer_O[kMaxPeers] = {};
  bool units_y_limit_exceeded() const { return (long long)B * H > 65535; }
};

// Longest sequence whose per-32-key-block scale table still fits in shared memory behind the tiles of the
// INT8 kernel with padded head dimension d_pad (block-scale mode); longer sequences need per-head scales.
int attention_max_block_keys(int d_pad);

// Enqueues the kernel; returns false and fills *err on a host-side failure.
bool launch_attention(const AttnLaunch& a, std::string* err);

}  // namespace qmha
