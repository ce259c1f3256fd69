// prepare.cuh — host-visible interface of the operand-preparation kernels (prepare.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace qmha {

struct PrepareArgs {
  const void* Q;   // [B, N, H*d] of in_dtype, device
  const void* K;
  const void* V;
  int in_dtype = 0;  // 0 = fp32 (the reference's API), 1 = fp16, 2 = bf16
  float* scales;   // [3, B*H] (INT8) — written by launch_absmax_and_scales, read by launch_prepare
  void* Qp;        // [B*H, n_pad, d_pad] int8 / fp16
  void* Kp;
  void* Vt;        // [B*H, d_pad, n_pad] fp16
  int B, N, H, d, n_pad, d_pad;
  bool int8;
  bool bf16 = false;  // 16-bit kernels: operands converted to bf16 instead of fp16
  bool v8 = false;    // INT8 P.V mode: Vt receives int8 codes [B*H, d_pad, n_pad] instead of codes stored as fp16
  cudaStream_t stream;
  // fused RoPE (utils/verify.cu:9-23 applied to Q and K rows before absmax / quantisation):
  // table of {cos, sin}(pos * base^(-2k/d)) as float2 [N][d/2], built on the host; nullptr = off
  const float2* rope = nullptr;
  // Q, K, V may be (batch, head-range) slabs of larger tensors: distance between consecutive rows / batch entries in
  // elements (0 = dense: H*d and N*H*d); the same pitches for all three
  long long in_ld = 0, in_bs = 0;
  int ld() const { return in_ld > 0 ? (int)in_ld : H * d; }
  size_t bs() const { return in_bs > 0 ? (size_t)in_bs : (size_t)N * (size_t)ld(); }
};

// absmax pass + scale finalisation (3 launches incl. the memset node).
cudaError_t launch_absmax_and_scales(const PrepareArgs& a, unsigned* amax_bits, int per_tensor);
// absmax + quantise + re-layout in ONE launch (cluster kernel; INT8, per-(batch,head) scales, d%4==0).
cudaError_t launch_fused_quantize(const PrepareArgs& a);
// the same on a persistent grid without clusters (default for per-head scales); ctl: 16 + 6*B*H words of scratch
cudaError_t launch_stream_quantize(const PrepareArgs& a, unsigned* ctl);
// single-pass quantise with the reference's per-32-row-block scales; scales = [3][B*H][n_pad/32]
cudaError_t launch_block_quantize(const PrepareArgs& a);
// per unit max V scale + per block {log2 r, 1/r}, r = sV_block / sV_max (attention, block mode)
cudaError_t launch_block_aux(const float* scales_v, float* aux, float* vmax, int units, int nblk,
                             cudaStream_t stream);
// quantise / convert + re-layout (1 launch).
cudaError_t launch_prepare(const PrepareArgs& a);
cudaError_t launch_quantize_blocks(const float* X, int B, int N, int H, int d, int block_rows,
                                   int8_t* q, float* scales, cudaStream_t stream);
cudaError_t launch_quantize_static(const float* X, long long n, float scale, float zp, int8_t* q,
                                   cudaStream_t stream);

}  // namespace qmha
