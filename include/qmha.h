/* qmha.h — C-ABI of the B200-native quantised multi-head-attention forward.
 *
 * This is the drop-in boundary for the reference's hot path.  Every entry point takes plain
 * pointers and sizes (no torch / C++ types) so it can be bound from ctypes, pybind, cgo, JNI …
 * Citations are file:line in the reference tree (MattJBorowski1991/QuantizedMHA).
 *
 *   solve()                 == include/launchers.h:9-10 (defined once per mha_kernels/<k>.cu,
 *                              e.g. fa_tc_int8_b.cu:600-609); same symbol, same signature,
 *                              same "complete on return" behaviour (launchers.h:64).
 *   qmha_forward()          replaces launchers.h:16-72 launch<KernelFn>() + the per-head
 *                              extract/kernel/concat loop; adds batch, stream, variant.
 *   qmha_quantize_*()       expose kernel (a), the replacement of fp32_to_int8sram
 *                              (fa_tc_int8_b.cu:33-152), for bit-exact checks.
 *   qmha_attention_prepared() exposes kernel (b)/(c), the replacement of fa_kernel
 *                              (fa_tc_int8_b.cu:408-579 / fa_tc_v2a.cu:274-496).
 *
 * Error model: solve() stays void and never throws (the reference ignores CUDA errors inside
 * launch(), launchers.h:27-71); every other entry returns 0 on success, non-zero on failure,
 * and qmha_last_error() returns a message for the calling thread.  There is NO CPU fallback:
 * without an sm_100 device every compute entry fails with an error.
 */
#ifndef QMHA_H
#define QMHA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Kernel variants (selected per call; solve() uses the build/runtime default, see below). */
#define QMHA_KERNEL_INT8 0 /* Q·K^T tcgen05 kind::i8, P·V kind::f16; replaces fa_tc_int8_a/b   */
#define QMHA_KERNEL_F16 1  /* Q·K^T and P·V tcgen05 kind::f16; replaces fa_tc_v1a..v2b, fa, unfused */
#define QMHA_KERNEL_BF16 2 /* same pipeline with bf16 operands (Q, K, P, V rounded to bf16), fp32 accumulation */
#define QMHA_KERNEL_INT8_PV8 3 /* INT8 kernel with the reference's P.V semantics (fa_tc_int8_b.cu:359-371): P quantised to
                                  8-bit codes (static scale) and multiplied with the int8 V codes on the INT8 pipe, int32
                                  accumulation.  Opt-in: coarser than fp16 P for long-tailed rows and not faster here
                                  (DESIGN.md §4.2.1).  Names: "int8_pv8", "fa_b200_int8_pv8". */

/* Element types of Q, K, V and of the output in the extended entries (the reference's API is fp32 only). */
#define QMHA_DTYPE_F32 0
#define QMHA_DTYPE_F16 1
#define QMHA_DTYPE_BF16 2

/* Granularity of the dynamic INT8 scales (symmetric, zero-point free, fa_tc_int8_b.cu:104). */
#define QMHA_GRAN_TENSOR 0 /* one scale per tensor                                   */
#define QMHA_GRAN_HEAD 1   /* one scale per (batch, head) slab [N, d]  (default)     */
#define QMHA_GRAN_BLOCK 2  /* one scale per (batch, head, 32-row block): the reference's own
                              granularity (Br x d / Bc x d tiles); scales are [3, B*h, n_pad/32] */

/* ---- the reference's own entry point -------------------------------------------------------
 * include/launchers.h:9-10.  Q,K,V,output: DEVICE pointers, fp32, contiguous row-major
 * [N, d_model]; head j occupies columns [j*d, (j+1)*d), d = d_model / h.  Synchronous on return.
 * Unlike the reference, N, d_model and h are honoured at run time (d <= 128, d % 4 == 0).
 * The variant is the library default: QMHA_DEFAULT_KERNEL at build time (Makefile KERNEL=),
 * overridable with qmha_set_kernel() or the QMHA_KERNEL environment variable; the reference's
 * kernel names are accepted (fa_tc_int8_a/b -> INT8; fa, unfused, fa_tc_v1a..v2b -> F16). */
void solve(const float* Q, const float* K, const float* V, float* output, int N, int d_model,
           int h);

/* ---- extended, stream-ordered entry --------------------------------------------------------
 * Q,K,V,O: DEVICE pointers, fp32, contiguous [B, N, d_model].  `stream` is a cudaStream_t
 * (NULL = legacy default stream).  Asynchronous: returns after enqueueing.  Scratch (int8 /
 * fp16 operands, scales) lives in a per-device workspace grown on demand and reused.
 * Threading / streams: entry points may be called from several host threads and on different streams
 * of one device; calls that use the workspace are serialised on the host for the duration of their
 * enqueue and ordered on the device behind the previous user of the workspace (an event wait when the
 * stream differs), so concurrent callers get correct results but do not overlap on one GPU.
 * qmha_shutdown() must not run concurrently with other calls. */
int qmha_forward(const float* Q, const float* K, const float* V, float* O, int B, int N,
                 int d_model, int h, int kernel, int gran, void* stream);

/* Everything a call can choose, per call (nothing here reads or writes process-wide state; the fields
 * marked "-1 = default" fall back to qmha_set_kernel / qmha_set_rope / the environment).  Initialise with
 * qmha_args_init(), which also sets struct_size (checked by the library: header / library mismatch fails).
 * SURVEY.md §8(b) "Data" row: batch, in/out dtype, mode, granularity, stream in one extended entry. */
typedef struct qmha_args {
  size_t struct_size;       /* sizeof(qmha_args)                                                        */
  const void* Q;            /* DEVICE pointers, contiguous [B, N, d_model] of in_dtype, 16-byte aligned  */
  const void* K;
  const void* V;
  void* O;                  /* [B, N, d_model] of out_dtype                                             */
  int B, N, d_model, h;
  int kernel;               /* QMHA_KERNEL_*; -1 = library default (what solve() uses)                   */
  int gran;                 /* QMHA_GRAN_*;   -1 = default for the shape (qmha_granularity_for)          */
  int in_dtype, out_dtype;  /* QMHA_DTYPE_*: 16-bit callers skip the fp32 round trip (the quantise pass
                               then reads 2 instead of 4 bytes per element)                             */
  int rope;                 /* fused RoPE on Q and K: 1 = on, 0 = off, -1 = process default              */
  float rope_base;          /* used when rope == 1; <= 1 means 10000                                     */
  int variant;              /* -1 (tuning aid: exp2 share on the FMA pipe, only in QMHA_BUILD_POLY builds)*/
  void* stream;             /* cudaStream_t; NULL = legacy default stream                                */
  /* Output placement (the caller-side concat / gather folded into the kernel's epilogue).
   * O may be a (batch, head-range) SLAB of a larger tensor: o_row_stride / o_batch_stride are the distances, in
   * elements of out_dtype, between consecutive rows / batch entries of O (0 = dense: d_model and N * d_model).
   * peer_O[0 .. n_peers): further destinations that receive the SAME bytes with the same strides — replicas of the
   * output on other GPUs (device pointers mapped through qmha_ipc_open / peer access; NVLink stores issued by the
   * epilogue's TMA, tile by tile, while the rest of the grid still computes) or on the same device.  The writes are
   * complete when the launch is; ordering against the peers' own streams is the caller's (see sharding.py). */
  int64_t o_row_stride, o_batch_stride;
  int n_peers;              /* 0 .. QMHA_MAX_PEERS                                                       */
  void* peer_O[7];
  /* Input placement: Q, K, V may likewise be slabs of larger tensors (the same pitches for all three, in elements of
   * in_dtype, multiples of 16 bytes; 0 = dense).  The quantise / convert pass reads them in place — a caller that shards
   * heads hands over its columns of the full [B, N, H*d] tensors without a copy ("inputs / outputs for a
   * unit are the strided slices [b, :, h, :]"). */
  int64_t in_row_stride, in_batch_stride;
  int device;               /* device ordinal that owns the pointers and the stream; -1 = the calling thread's current
                               device (one workspace per device; the caller's current device is restored on return)  */
} qmha_args;
#define QMHA_MAX_PEERS 7
void qmha_args_init(qmha_args* a);
int qmha_forward_ex(const qmha_args* a);   /* asynchronous, stream-ordered like qmha_forward             */

/* Same computation from HOST buffers (pageable or pinned): H2D, compute and D2H are pipelined
 * over (batch entry, head group) chunks on internal streams — also for B = 1, the reference's own call
 * shape — synchronous on return.  gran -1 = default for the shape.  This is what bench.py's `e2e` times. */
int qmha_forward_host(const float* Q, const float* K, const float* V, float* O, int B, int N,
                      int d_model, int h, int kernel, int gran);
/* The same with fp16 / bf16 host buffers (QMHA_DTYPE_*): a 16-bit caller moves half the bytes over PCIe. */
int qmha_forward_host_ex(const void* Q, const void* K, const void* V, void* O, int B, int N, int d_model, int h,
                         int kernel, int gran, int in_dtype, int out_dtype);

/* ---- operand preparation (kernel (a)) ------------------------------------------------------
 * Internal operand layout ("prepared" tensors), u = b*h + head:
 *   Qp, Kp : [B*h, n_pad, d_pad]   int8 (INT8 variant) or fp16 (F16 variant), zero padded
 *   Vt     : [B*h, d_pad, n_pad]   fp16, TRANSPOSED (keys contiguous); for INT8 the values are
 *                                   the int8 codes stored exactly in fp16
 *   scales : [3, B*h] fp32 (Q, K, V) — for QMHA_GRAN_TENSOR every entry of a row is equal;
 *            [3, B*h, n_pad/32] for QMHA_GRAN_BLOCK
 * n_pad = N rounded up to 256, d_pad = 32/64/128 >= d. */
int qmha_workspace_dims(int N, int d_model, int h, int* n_pad, int* d_pad);

/* Dynamic absmax quantisation, kernel spec (fa_tc_int8_b.cu:104-106,136-140):
 * sc = max(absmax/127, 1e-8), q = clamp(rint(v * (1.0f/sc)), -128, 127). */
int qmha_quantize_qkv(const float* Q, const float* K, const float* V, int B, int N, int d_model,
                      int h, int gran, int8_t* Qp, int8_t* Kp, uint16_t* Vt, float* scales,
                      void* stream);

/* Same with fp32 / fp16 / bf16 inputs and a per-call RoPE choice (rope: 1 / 0 / -1 = process default). */
int qmha_quantize_qkv_ex(const void* Q, const void* K, const void* V, int in_dtype, int B, int N, int d_model,
                         int h, int gran, int rope, float rope_base, int8_t* Qp, int8_t* Kp, uint16_t* Vt,
                         float* scales, void* stream);

/* Same, operands for `kernel` = QMHA_KERNEL_INT8 (V codes stored as fp16) or QMHA_KERNEL_INT8_PV8 (Vt = int8 codes
 * [B*h, d_pad, n_pad]). */
int qmha_quantize_qkv_k(const void* Q, const void* K, const void* V, int in_dtype, int B, int N, int d_model,
                        int h, int kernel, int gran, int rope, float rope_base, int8_t* Qp, int8_t* Kp, void* Vt,
                        float* scales, void* stream);

/* fp32 -> fp16 operand conversion for the F16 variant (fa_tc_v1a.cu:267,321,348 convert on
 * load; here it is one HBM-bound pre-pass).  Qp/Kp are fp16 stored as uint16_t. */
int qmha_convert_qkv_f16(const float* Q, const float* K, const float* V, int B, int N,
                         int d_model, int h, uint16_t* Qp, uint16_t* Kp, uint16_t* Vt,
                         void* stream);
/* Operand conversion for the 16-bit kernels: kernel = QMHA_KERNEL_F16 (fp16 operands) or QMHA_KERNEL_BF16
 * (bf16 operands), inputs of in_dtype. */
int qmha_convert_qkv_16(const void* Q, const void* K, const void* V, int in_dtype, int B, int N, int d_model,
                        int h, int kernel, int rope, float rope_base, uint16_t* Qp, uint16_t* Kp, uint16_t* Vt,
                        void* stream);

/* Reference-granularity quantisation of ONE tensor in the INPUT layout [B, N, d_model]:
 * one scale per (batch, head, block of block_rows rows) exactly like fp32_to_int8sram on a
 * Br x d tile.  q has the input layout; scales is [B*h*ceil(N/block_rows)]. */
int qmha_quantize_blocks(const float* X, int B, int N, int d_model, int h, int block_rows,
                         int8_t* q, float* scales, void* stream);

/* Static-scale quantisation, golden spec (tests/generate_golden.cpp:94-101):
 * q = clamp((int)round(x/scale + zero_point), -128, 127), round half away from zero. */
int qmha_quantize_static(const float* X, int64_t n, float scale, float zero_point, int8_t* q,
                         void* stream);

/* ---- attention on prepared operands (kernel (b)/(c)) --------------------------------------- */
int qmha_attention_prepared(const void* Qp, const void* Kp, const uint16_t* Vt,
                            const float* scales, float* O, int B, int N, int d_model, int h,
                            int kernel, int gran, void* stream);

int qmha_attention_prepared_ex(const void* Qp, const void* Kp, const uint16_t* Vt,
                               const float* scales, void* O, int out_dtype, int B, int N, int d_model,
                               int h, int kernel, int gran, void* stream);

/* qmha_forward*()/qmha_attention_prepared*() are asynchronous; after synchronising the stream,
 * this reports (and clears) a device-side pipeline failure recorded by the kernel (a bounded mbarrier
 * wait that gave up).  Independently of this call, a failing launch also writes its record to a mapped
 * host word that EVERY entry point checks first: the call after a stalled launch fails with that record
 * (once) instead of silently going on, and later launches are not affected by it. */
int qmha_check_async_error(void);
/* cudaStreamSynchronize(stream) + qmha_check_async_error() in one call (for bindings without CUDA headers). */
int qmha_synchronize(void* stream);
/* Test hook: plants a failure record exactly as a stalled CTA would. */
int qmha_debug_inject_stall(int site);

/* Debug/tuning aid: runs the instrumented INT8 d=128 attention kernel once (synchronous) and
 * returns the clock64 timeline of CTA (0,0): host_trace[9][ceil(N/64)][4] (softmax warps
 * 0-7, MMA issuer of tile 0).  variant k: polynomial exp2 on every k-th pair (0 = none). */
int qmha_debug_attention_trace(const void* Qp, const void* Kp, const uint16_t* Vt,
                               const float* scales, float* O, int B, int N, int d_model, int h,
                               int variant, long long* host_trace);

/* Development aid: with QMHA_CYCLES=1 in the environment, every qmha_attention_prepared() launch adds
 * the SM clocks each CTA was resident to a device counter; out2 = {sum of clocks, CTAs}. */
int qmha_debug_cycles(unsigned long long* out2, int reset);
/* Per SM s < n_sms (<= 192): out[2*s] = SM clocks between the first CTA start and the last CTA end on that SM since
 * the last reset (one launch between resets), out[2*s+1] = 0. */
int qmha_debug_sm_spans(unsigned long long* out, int n_sms, int reset);

/* ---- fused RoPE ------------------------------------------------------------
 * The reference's CPU check rotates Q and K (utils/verify.cu:56-69) but no GPU kernel ever calls
 * the device helper apply_rope (utils/utils.cu:50-65).  With qmha_set_rope(1, base) every entry
 * point (solve, qmha_forward*, qmha_quantize_qkv, qmha_convert_qkv_f16) rotates the Q and K rows
 * inside the HBM-bound quantise / convert pass, before the block maxima are taken: position =
 * row index within the batch entry, pairs (k, k + d/2), theta = powf(base, -2k/d) exactly as
 * verify.cu:9-23 / generate_golden.cpp:38-51.  The rotated values are bit-identical to the CPU
 * restatement (cos/sin come from a host-built table).  Needs d % 8 == 0; every scale granularity and the
 * 16-bit kernels are covered.  qmha_set_rope is the process-wide DEFAULT (QMHA_ROPE=1 in the environment enables it at
 * start); qmha_args.rope / the *_ex entries choose per call. */
int qmha_set_rope(int enable, float base);
int qmha_get_rope(void);

/* ---- peer memory (for qmha_args.peer_O) -------------------------------------------------------
 * One process per GPU: the owner of a device allocation exports it, the other ranks open it and get a device
 * pointer that is valid in THEIR process (CUDA IPC over NVLink / NVSwitch peer mappings).
 *   qmha_ipc_export: dev_ptr may point anywhere inside a cudaMalloc'ed allocation (a torch tensor's data_ptr()
 *                    works when the allocator does not use expandable segments); handle = 64 opaque bytes,
 *                    *offset = distance of dev_ptr from the allocation's base.
 *   qmha_ipc_open:   maps the allocation on the CURRENT device of the calling process and returns base + offset;
 *                    mappings are cached per handle and reference-counted: qmha_ipc_close(ptr) drops one reference
 *                    (the mapping goes with the last one), qmha_ipc_close_all() / qmha_shutdown() release everything.
 * Several devices in ONE process need no handles: qmha_enable_peer_access(dev, peer) once per ordered pair. */
int qmha_ipc_export(const void* dev_ptr, unsigned char handle[64], int64_t* offset);
int qmha_ipc_open(const unsigned char handle[64], int64_t offset, void** dev_ptr);
int qmha_ipc_close(void* dev_ptr);
int qmha_ipc_close_all(void);
int qmha_enable_peer_access(int dev, int peer);

/* ---- housekeeping -------------------------------------------------------------------------- */
const char* qmha_last_error(void);          /* "" when the last call on this thread succeeded */
int qmha_set_kernel(const char* name);      /* default variant used by solve(); 0 = ok        */
const char* qmha_get_kernel(void);          /* "int8", "f16" or "bf16"                        */
int qmha_default_granularity(int d_model, int h); /* QMHA_GRAN_BLOCK when d%4==0 (QMHA_SCALES=head|tensor
                                                     overrides), else HEAD                            */
int qmha_granularity_for(int N, int d_model, int h); /* what solve() uses: the same, but HEAD when the per-block
                                                     scale table of one unit no longer fits in shared
                                                     memory behind the tiles (N > ~68 k at d = 128)  */
int qmha_kernel_from_name(const char* name); /* QMHA_KERNEL_* or -1                           */
int64_t qmha_launch_count(void);            /* kernels launched by this library so far        */
void qmha_shutdown(void);                   /* frees every per-device workspace               */
const char* qmha_version(void);

#ifdef __cplusplus
}
#endif
#endif /* QMHA_H */
