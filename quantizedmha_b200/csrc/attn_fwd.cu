// attn_fwd.cu — warp-specialised FlashAttention-2 forward for sm_100a (B200).
//
// Replaces the reference's fa_kernel family: mha_kernels/fa_tc_int8_b.cu:408-579 (INT8, WMMA
// IMMA) and fa_tc_v1a.cu:222-413 / fa_tc_v2a.cu:274-496 (FP16, WMMA HMMA).  Same math
// (softmax(Q·K^T/sqrt(d))·V per head, FP32 statistics and output) — different machine:
//
//   * one CTA owns TWO 128-row query tiles of one (batch, head) unit and walks the KV sequence
//     in 128-key tiles that are processed as two pipelined 64-key HALF-STEPS;
//   * warps 8-11 form a service warpgroup (register budget handed to the softmax warps with
//     setmaxnreg); warp 9 is the TMA producer: Q once, then K and V^T tiles through two
//     mbarrier rings;
//   * warp 8 (one elected lane) issues every tcgen05.mma:  S_t[b] = Q_t·K_half^T  (kind::i8,
//     int32 accumulators in TMEM — kind::f16 for the FP16 variant) into a DOUBLE-BUFFERED score
//     tile, and  O_t += P_t[b]·V_half  (kind::f16, A operand = P read straight from TMEM,
//     B = V^T tile in shared memory).  P_t(i) is written over the score buffer of half-step i+1
//     (already copied to registers by then), so S_t for half-step i+3 can be issued right behind
//     P·V of half-step i: scores are ready more than a full step before the softmax warps ask
//     for them and the tensor-core round trip is off the softmax critical path;
//   * warps 0-3 / 4-7 are the softmax warpgroups of tile 0 / tile 1: one thread per query row,
//     S read with tcgen05.ld, dequant scale folded into the exponent FMA (packed FFMA2), exp2
//     on the MUFU with an optional share on the FMA pipe (Cody-Waite + cubic polynomial), P
//     written back over S with tcgen05.st, running row sum in FP32, lazy rescale of O in TMEM
//     only when the row max grows by more than 2^4;
//   * epilogue: O·(sV/l) from TMEM to global memory in the reference's [N, h·d] layout.
//
// TMEM plan (512 columns): S0[0] = [0,64)  S0[1] = [64,128)  S1[0] = [128,192)  S1[1] = [192,256)
//                          O0 = [256,384)  O1 = [384,512);  S_t(j) lives in S_t[j&1], P_t(i) in the
//                          first 32 columns of S_t[(i+1)&1].
#include <cstdlib>
#include <cstring>
#include <type_traits>
#include "attn_fwd.cuh"
#include "sm100_ptx.cuh"

namespace qmha {

using namespace ptx;

namespace {

constexpr int kBM = 128;        // query rows per tile == UMMA M
constexpr int kBN = 128;        // keys per KV tile in shared memory (one TMA stage)
constexpr int kHN = 64;         // keys per half-step == UMMA N of Q·K^T, UMMA K extent of P·V
// Service warps 8..11 (one per SM sub-partition).  tcgen05.mma issue visibly slows the softmax
// warps that share the issuing warp's sub-partition, so the issue work is split: with
// kMmaSplit == 4 every service warp issues: warp 8+2t+p handles the half-steps of parity p of query
// tile t, and the TMA production is folded into the even-parity warps (warp 8: Q and the K ring,
// warp 10: the V^T ring), one tile per iteration, so every sub-partition carries a quarter of the
// issue work (builds with -DQMHA_MMA_SPLIT=4, passes every GPU test; measured equal to the 2-way split
// for INT8 — 1397 vs 1404 clk per half-step — so the simpler layout is the default).  With
// kMmaSplit == 2, warp 8 issues everything for query tile 0 and warp 11 for tile 1 (warps 9 / 10
// are the K / V TMA producers).
#ifndef QMHA_MMA_SPLIT
#define QMHA_MMA_SPLIT 2
#endif
constexpr int kMmaSplit = QMHA_MMA_SPLIT;
// tcgen05.commit is expensive for the issuing warp (traced: ~100 clk each, MMAs in flight or not).  Two ways to
// commit less, each 0 = off, 1 = FP16/BF16 kernels only, 2 = every kernel:
//   QMHA_SOFT_RING: the K / V^T ring stages are handed back by one softmax thread of each tile when the scores
//     that prove the last read of the stage has retired arrive (S(2j+1) -> K tile j is free; S(2j+4) is issued
//     behind P.V(2j+1) by the same thread -> V tile j is free) instead of by two commits per tile per MMA warp;
//   QMHA_LAZY_PV: "P.V(i) retired" is read off s_full(i+3); only the last three half-steps keep a pv_done commit.
// Same-box A/B at the headline shape (clocks per CTA): FP16 219.6 k -> 198.3 k with both (-10 %: that kernel waits
// on the tensor side); INT8 184.6 k -> 193.1 k with both (+4.5 %: the softmax warps set the pace there and pay for
// the extra arrive), so INT8 keeps committing the ring barriers.  QMHA_LAZY_PV alone is neutral to -1 % for INT8 (184.5 k
// against 186.4 k with the bias MMA in place) and is on for every kernel.
#ifndef QMHA_SOFT_RING
#define QMHA_SOFT_RING 1
#endif
#ifndef QMHA_LAZY_PV
#define QMHA_LAZY_PV 2
#endif
static_assert((QMHA_SOFT_RING == 0 && QMHA_LAZY_PV == 0) || kMmaSplit == 2, "written for the two-warp MMA issue");
// QMHA_BIAS_MMA (INT8 kernels): the tensor pipe delivers the scores as FLOAT BIT PATTERNS.  One extra kind::f16
// instruction per score tile writes the constant 12582912.0f = 0x4B400000 into the accumulator (A = 128x16 tile of 768.0,
// B = 64x16 tile of 1024.0: 16 * 768 * 1024 = 1.5 * 2^23 exactly; every element equal, so the swizzle does not matter) and
// the four kind::i8 instructions ACCUMULATE their int32 dot products onto that bit pattern: bits(S) = 0x4B400000 + s is the
// float 12582912 + s, which the exponent FMA consumes as it is.  The per-element integer add of the int->float trick — the one
// instruction whose removal moved the kernel in the knock-out runs (5.85 -> 5.28 ms) — disappears from the softmax warps; the
// tensor side pays one more feed-bound instruction (~48 clk) per tile and half-step.
#ifndef QMHA_BIAS_MMA
#define QMHA_BIAS_MMA 1
#endif
constexpr bool kBiasMma = QMHA_BIAS_MMA != 0;
// Code-generation experiments on the persistent kernel (run with QMHA_PERSIST_GRID = number of items, i.e. one item per
// CTA): bit 0 = the softmax warps stop after their first item, bit 1 = the MMA warps, bit 2 = the producers, bit 3 = the
// epilogue stages through the K + V^T rings like the one-CTA-per-item kernel.
#ifndef QMHA_PERSIST_SINGLE
#define QMHA_PERSIST_SINGLE 0
#endif
// Persistent INT8 d = 128 kernel (see attn_fwd_kernel, kPersist): 0 = off unless QMHA_PERSIST=1, 1 = on unless QMHA_PERSIST=0.
// Built, bit-identical, and measured SLOWER than one CTA per item at C4 (187.9 k against 181.0 k clk per item): the gaps
// between CTAs, the setup and the first-scores latency do disappear (1.6 k clk per item against the same code run with one
// item per CTA), but ptxas schedules the main loop of the nested form 4.8 % slower (189.5 k clk with one item per CTA; same
// instruction counts for the arithmetic, more integer / predicate work around it).  Opt-in.
#ifndef QMHA_DEFAULT_PERSIST
#define QMHA_DEFAULT_PERSIST 0
#endif
constexpr int kBiasTileBytes = 4096 + 2048;   // A: 128 rows x 32 B, B: 64 rows x 32 B (fp16, SWIZZLE_32B layout)
constexpr int kAllocWarp = 8;
constexpr int kTmaWarp = 9;
constexpr int kTmaWarpV = 10;
__host__ __device__ constexpr bool is_mma_warp(int warp) {
  return kMmaSplit == 4 ? warp >= 8 : (warp == 8 || warp == 11);
}
constexpr int kThreads = 384;            // 2 softmax warpgroups + 1 service warpgroup (MMA, TMA, 2 idle)
#ifndef QMHA_REGS_SOFTMAX
#define QMHA_REGS_SOFTMAX 208
#endif
#ifndef QMHA_REGS_SERVICE
#define QMHA_REGS_SERVICE 72
#endif
constexpr int kRegsSoftmax = QMHA_REGS_SOFTMAX;        // setmaxnreg budgets: 2*128*208 + 128*72 = 62464 <= 65536
constexpr int kRegsService = QMHA_REGS_SERVICE;
static_assert(2 * 128 * kRegsSoftmax + 128 * kRegsService <= 65536 && kRegsSoftmax % 8 == 0 && kRegsService % 8 == 0, "register budgets");
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kColS0 = 0, kColS1 = 128, kColO0 = 256, kColO1 = 384;  // S_t[b] at kColS_t + 64*b
constexpr float kRescaleThreshold = 4.0f;  // log2 units: P <= 2^4, well inside fp16
constexpr int kMagicI2F = 0x4B400000;      // float(1.5 * 2^23): int -> float by bit tricks
constexpr float kMagicF = 12582912.0f;

template <bool kInt8, int kD, bool kPv8 = false>
struct Cfg {
  static_assert(!kPv8 || kInt8, "INT8 P.V is a mode of the INT8 kernel");
  static constexpr int kEltQK = kInt8 ? 1 : 2;
  static constexpr int kRowBytesQK = kD * kEltQK;                       // bytes per Q/K row
  static constexpr int kAtomQK = kRowBytesQK < 128 ? kRowBytesQK : 128;  // swizzle span
  static constexpr int kSubQK = kRowBytesQK / kAtomQK;                  // 128B sub-tiles per row
  static constexpr int kTileBytesQK = kBM * kRowBytesQK;
  static constexpr int kSubBytesQK = kBM * kAtomQK;
  static constexpr int kStepsQK = kRowBytesQK / 32;                     // UMMA K steps (32 B each)
  // V^T tile: kD rows x 128 keys.  fp16 (default): two 64-key sub-tiles of kD x 128 B; INT8 P.V: one sub-tile,
  // 128 keys = 128 B per row, the second half-step starts 64 B into the row (inside the swizzle atom).
  static constexpr int kTileBytesV = kD * kBN * (kPv8 ? 1 : 2);
  static constexpr int kSubTilesV = kPv8 ? 1 : 2;                       // TMA boxes per tile
  static constexpr int kSubBytesV = kD * 128;                           // one box
  static constexpr int kHalfOffV = kPv8 ? 64 : kD * 128;                // byte offset of keys 64..127 for the descriptor
  static constexpr int kStepsPV = kPv8 ? kHN / 32 : kHN / 16;           // UMMA K steps per half-step
  static constexpr int kHalfBytesQK = kHN * kAtomQK;                    // byte offset of K rows 64.. in a sub-tile
  // K tiles are consumed 1.5 tiles ahead of V tiles (S runs three half-steps ahead of P·V), so the
  // K ring is one stage deeper than the V ring; each ring has its own producer warp.
  static constexpr int kBudget = 224 * 1024 - 2 * kTileBytesQK;
  static constexpr int kStagesRaw = (kBudget - kTileBytesQK) / (kTileBytesQK + kTileBytesV);
  static constexpr int kStagesV = kStagesRaw > 3 ? 3 : kStagesRaw;
#ifdef QMHA_EXP_K3
  static constexpr int kStagesK = kStagesV;
#else
  static constexpr int kStagesK = kStagesV + 1;
#endif
  static_assert(kStagesV >= 2, "need at least double buffering");
  static constexpr bool kBias = kInt8 && kBiasMma;
  static constexpr int kSmemTiles = 2 * kTileBytesQK + kStagesK * kTileBytesQK + kStagesV * kTileBytesV +
                                    (kBias ? kBiasTileBytes : 0);
  static constexpr uint32_t kIdescBias = make_idesc(kAccF32, kFmtF16, kFmtF16, kBM, kHN);
  static constexpr int kSmemBytes = kSmemTiles + 1024 /*align slack*/ + 512 /*barriers*/;
  static constexpr uint32_t kIdescQK =
      kInt8 ? make_idesc(kAccS32, kFmtS8, kFmtS8, kBM, kHN)
            : make_idesc(kAccF32, kFmtF16, kFmtF16, kBM, kHN);
  static constexpr uint32_t kIdescPV = make_idesc(kAccF32, kFmtF16, kFmtF16, kBM, kD);
  // BF16 anchor: same kind::f16 instructions with the bf16 operand format (operands, P and V^T are bf16)
  // INT8 P.V (QMHA_KERNEL_INT8_PV8): P as unsigned 8-bit codes from TMEM, V^T as int8, int32 accumulators
  static constexpr uint32_t kIdescPV8 = make_idesc(kAccS32, kFmtU8, kFmtS8, kBM, kD);
  static constexpr uint32_t kIdescQKbf = make_idesc(kAccF32, kFmtBF16, kFmtBF16, kBM, kHN);
  static constexpr uint32_t kIdescPVbf = make_idesc(kAccF32, kFmtBF16, kFmtBF16, kBM, kD);
};

// Output tensor maps of the peer destinations (qmha_args.peer_O): the epilogue repeats every TMA tensor store of the
// staged tile once per peer, so a replica of the output lands in the other GPUs' memory over NVLink while the rest of
// the grid still computes (the "gather" of a sharded forward costs no extra pass and no extra kernel).
struct PeerMaps {
  CUtensorMap m[kMaxPeers];
};

struct Barriers {
  uint64_t q_full;
  uint64_t k_full[4], k_empty[4];
  uint64_t v_full[4], v_empty[4];
  uint64_t s_full[2][2], p_full[2][2], pv_done[2][2];  // [tile][buffer / step parity]
  uint64_t o_final[2];                              // one-shot: last P·V of the tile retired
  uint64_t s0_read[2];                              // one-shot: S_t(0) copied to registers
  uint64_t qk_done;                                 // once per item: every Q·K^T of the item retired (K ring, Q tile free)
  uint64_t epi_done;                                // persistent kernel, once per item: the output stores have read their staging tiles (V^T ring free)
  uint32_t tmem_base;
  uint32_t pad;
};
static_assert(sizeof(Barriers) <= 512, "barrier block too large");

// Bounded mbarrier wait.  A healthy wait is microseconds; after ~1e9 cycles (or as soon as any
// CTA has raised the global error flag) the wait gives up, records the site and lets the CTA
// drain so the kernel always terminates and the host can report the failure.
// Failure record of one launch.  flag = (launch id << 12) | wait site: a failure only drains the CTAs of the
// launch that raised it (an earlier version used a bare site number, which made every LATER launch return at
// kernel entry with its output unwritten); `host` is a mapped host word the library checks at the start of
// every call without synchronising.
struct ErrCtx {
  int* flag;
  int* host;
  unsigned id;
  __device__ __forceinline__ bool raised() const {
    return ((unsigned)*((volatile int*)flag) >> 12) == id;
  }
  __device__ __forceinline__ void raise(int site) const {
    atomicExch(flag, (int)((id << 12) | (unsigned)site));
    if (host) { *((volatile int*)host) = site; __threadfence_system(); }
  }
};

__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity, const ErrCtx& err_flag, int site,
                                          bool& dead) {
  if (dead) return false;
  if (mbar_try_wait(bar, parity)) return true;
  const long long t0 = clock64();
  uint32_t spins = 0;
#ifndef QMHA_WAIT_HINT_NS
#define QMHA_WAIT_HINT_NS 20000
#endif
  while (!mbar_try_wait_hint(bar, parity, QMHA_WAIT_HINT_NS)) {
    if ((++spins & 0xFu) == 0) {
      if (clock64() - t0 > 1000000000LL || err_flag.raised()) {
        if (!err_flag.raised()) err_flag.raise(site);
        dead = true;
        return false;
      }
    }
  }
  return true;
}

// ------------------------------------------------------------------------------------------------
// One softmax step for one thread (= one query row) over a 64-key half-step.
//   s[64]   : raw S row of the half-step from TMEM (int32 for INT8, fp32 bits for FP16)
//   c       : logits-to-log2 factor (sQ*sK*log2e/sqrt(d) or log2e/sqrt(d))
//   m_used  : the (lazily updated) reference max in log2 units
// Produces p[32] (packed fp16x2 P row) and adds the row sum into the lsum accumulators.
template <bool kInt8, bool kMasked>
__device__ __forceinline__ float tile_row_max(uint32_t (&s)[kHN], float c, int n_valid) {
  if constexpr (kInt8) {
    if constexpr (kMasked) {
#pragma unroll
      for (int i = 0; i < kHN; ++i)
        if (i >= n_valid) s[i] = (uint32_t)((kBiasMma ? kMagicI2F : 0) - (1 << 22));
    }
    int m0 = max((int)s[0], (int)s[1]), m1 = max((int)s[2], (int)s[3]);
    int m2 = max((int)s[4], (int)s[5]), m3 = max((int)s[6], (int)s[7]);
#pragma unroll
    for (int i = 8; i < kHN; i += 8) {
      m0 = max(max(m0, (int)s[i + 0]), (int)s[i + 1]);
      m1 = max(max(m1, (int)s[i + 2]), (int)s[i + 3]);
      m2 = max(max(m2, (int)s[i + 4]), (int)s[i + 5]);
      m3 = max(max(m3, (int)s[i + 6]), (int)s[i + 7]);
    }
    return (float)(max(max(m0, m1), max(m2, m3)) - (kBiasMma ? kMagicI2F : 0)) * c;   // (biased scores stay monotone)
  } else {
    if constexpr (kMasked) {
#pragma unroll
      for (int i = 0; i < kHN; ++i)
        if (i >= n_valid) s[i] = 0xFF800000u;  // -inf
    }
    float m0 = fmaxf(__uint_as_float(s[0]), __uint_as_float(s[1]));
    float m1 = fmaxf(__uint_as_float(s[2]), __uint_as_float(s[3]));
    float m2 = fmaxf(__uint_as_float(s[4]), __uint_as_float(s[5]));
    float m3 = fmaxf(__uint_as_float(s[6]), __uint_as_float(s[7]));
#pragma unroll
    for (int i = 8; i < kHN; i += 8) {
      m0 = fmaxf(fmaxf(m0, __uint_as_float(s[i + 0])), __uint_as_float(s[i + 1]));
      m1 = fmaxf(fmaxf(m1, __uint_as_float(s[i + 2])), __uint_as_float(s[i + 3]));
      m2 = fmaxf(fmaxf(m2, __uint_as_float(s[i + 4])), __uint_as_float(s[i + 5]));
      m3 = fmaxf(fmaxf(m3, __uint_as_float(s[i + 6])), __uint_as_float(s[i + 7]));
    }
    return fmaxf(fmaxf(m0, m1), fmaxf(m2, m3)) * c;
  }
}

// Packed fp32x2 helpers (FFMA2 / FADD2 / FMUL2: two lanes of fp32 per instruction on sm_100).
__device__ __forceinline__ uint64_t pack2(float a, float b) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void unpack2(uint64_t v, float& a, float& b) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ uint64_t fadd2(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ uint64_t fadd2_rm(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("add.rm.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ uint64_t fsub2(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ uint64_t fmul2(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}

// 2^x for a pair on the FMA pipe instead of the MUFU (the SFU is the scarcest unit of this
// kernel: 16 exp2/clk/SM).  Cody-Waite split x = n + f with n = floor(x) taken by a round-down
// add of 1.5*2^23, 2^f by a degree-3 minimax polynomial on [0,1) (max rel. error 9e-5, far
// below the fp16 rounding of P), 2^n by adding n to the exponent field.
__device__ __forceinline__ void exp2_poly_pair(float x0, float x1, float& e0, float& e1) {
  x0 = fmaxf(x0, -126.0f);
  x1 = fmaxf(x1, -126.0f);
  const uint64_t x = pack2(x0, x1);
  const uint64_t magic = pack2(kMagicF, kMagicF);
  const uint64_t xr = fadd2_rm(x, magic);           // mantissa low bits = floor(x)
  const uint64_t f = fsub2(x, fsub2(xr, magic));    // f in [0,1)
  uint64_t pl = ffma2(pack2(0.077119089663028717f, 0.077119089663028717f), f,
                      pack2(0.227564394474029541f, 0.227564394474029541f));
  pl = ffma2(pl, f, pack2(0.695146143436431885f, 0.695146143436431885f));
  pl = ffma2(pl, f, pack2(1.0f, 1.0f));
  float p0, p1, r0, r1;
  unpack2(pl, p0, p1);
  unpack2(xr, r0, r1);
  e0 = __int_as_float(__float_as_int(p0) + (__float_as_int(r0) << 23));
  e1 = __int_as_float(__float_as_int(p1) + (__float_as_int(r1) << 23));
}

// int32 score -> float(12582912 + s) bit pattern.  QMHA_I2F selects the instruction: 0 = integer add
// (ALU pipe), 1 = mad.lo by a run-time 1 (IMAD, FMA pipe), 2 = alternate (even columns add, odd IMAD).
#ifndef QMHA_I2F
#define QMHA_I2F 0
#endif
__device__ __forceinline__ float i2f_magic(uint32_t s, int one, int idx) {
  if (kBiasMma) return __uint_as_float(s);           // the tensor pipe already delivered 0x4B400000 + s
  if (QMHA_I2F == 3) return __uint_as_float(s);   // timing experiment only: no conversion at all (wrong results)
  if (QMHA_I2F == 5) return __int2float_rn((int)s);   // I2FP: a real conversion, no magic bias in the FMA
  if (QMHA_I2F == 1 || (QMHA_I2F == 2 && (idx & 1))) {
    int r;
    asm("mad.lo.s32 %0, %1, %2, 0x4B400000;" : "=r"(r) : "r"((int)s), "r"(one));
    return __int_as_float(r);
  }
  return __int_as_float((int)s + kMagicI2F);
}
// Pacing (QMHA_PACE_DEPTH > 0): ptxas schedules the 64 independent exp chains of a step as long bursts
// (all conversions, then all MUFU.EX2, then all packs), so the two softmax warps of a sub-partition queue
// on the MUFU together and then do their integer work together while the MUFU idles.  A data dependency
// that costs one LOP3 per group forces an interleaved order instead: the conversions of group g take a
// (run-time) zero derived from the packed P of group g - depth, so at most `depth` groups of MUFU work
// can be pulled ahead of the integer work of the groups before them.
#ifndef QMHA_PACE_DEPTH
#define QMHA_PACE_DEPTH 0
#endif
// Timing experiments only (results are wrong): bit 0 drops the fp16 pack, bit 1 the row sum, bit 2 the row
// max, bit 3 the MUFU itself — to see which unit the loop is really waiting for.
#ifndef QMHA_KO
#define QMHA_KO 0
#endif
#ifndef QMHA_PACE_GROUP
#define QMHA_PACE_GROUP 4   // pairs per group
#endif
__device__ __forceinline__ uint32_t pace_token(const uint32_t (&p)[kHN / 2], int i, uint32_t zero) {
  constexpr int G = QMHA_PACE_GROUP, D = QMHA_PACE_DEPTH;
  const int g = i / G;
  if (D == 0 || g < D) return 0u;
  return p[(g - D) * G + G - 1] & zero;
}
__device__ __forceinline__ void i2f_pair(uint32_t s0, uint32_t s1, int one, int i, float& f0, float& f1,
                                         uint32_t tok = 0u) {
  if (kBiasMma) { f0 = __uint_as_float(s0); f1 = __uint_as_float(s1); return; }
  if (QMHA_I2F == 4) {  // timing experiment only: one 64-bit add per pair (the carry makes s1 off by one when s0 < 0)
    uint64_t v;
    asm("mov.b64 %0, {%1, %2};" : "=l"(v) : "r"(s0), "r"(s1));
    v += 0x4B4000004B400000ull;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(f0), "=f"(f1) : "l"(v));
    return;
  }
  if (QMHA_PACE_DEPTH > 0) {
    f0 = __int_as_float((int)s0 + (int)(kMagicI2F + tok));
    f1 = __int_as_float((int)s1 + (int)(kMagicI2F + tok));
    return;
  }
  f0 = i2f_magic(s0, one, 2 * i);
  f1 = i2f_magic(s1, one, 2 * i + 1);
}

// INT8 P.V: P goes to the tensor pipe as unsigned 8-bit codes rn(127.5 * p), p <= 2 (the lazy-rescale threshold is one
// log2 unit in that mode).  log2(127.5) rides in the exponent bias; adding 1.5 * 2^23 leaves rn(e) in the low mantissa
// byte (no float->int conversion, which would run on the MUFU's pipe); one PRMT packs a pair, one more a quad.
constexpr float kLog2P8 = 6.994353436858858f;
__device__ __forceinline__ uint32_t pack_u8x2(float e0, float e1) {
  float m0, m1;
  unpack2(fadd2(pack2(e0, e1), pack2(kMagicF, kMagicF)), m0, m1);
  return __byte_perm(__float_as_uint(m0), __float_as_uint(m1), 0x0040);   // byte 0 = code(e0), byte 1 = code(e1)
}

// kPolyEvery: every kPolyEvery-th pair of the row takes the polynomial path (0 = all on MUFU).
template <bool kInt8, bool kMasked, int kPolyEvery, int kBegin = 0, int kEnd = kHN / 2, bool kBf16 = false, bool kPv8 = false>
__device__ __forceinline__ void tile_row_exp(const uint32_t (&s)[kHN], uint32_t (&p)[kHN / 2],
                                             float c, float m_used, int n_valid,
                                             uint64_t (&lsum)[2], int one) {
  // x = s*c - m_used.  INT8: s is an int32 with |s| < 2^22, so bits(s + 0x4B400000) is the
  // float 12582912 + s exactly and one FMA does int->float, scale and max subtraction.
  const float bias = ((kInt8 && QMHA_I2F != 5) ? -fmaf(kMagicF, c, m_used) : -m_used) + (kPv8 ? kLog2P8 : 0.f);
  const uint64_t c2 = pack2(c, c), bias2 = pack2(bias, bias);
#pragma unroll
  for (int i = kBegin; i < kEnd; ++i) {
    float f0, f1;
    if constexpr (kInt8) {
      i2f_pair(s[2 * i], s[2 * i + 1], one, i, f0, f1, pace_token(p, i, (uint32_t)(one ^ 1)));
    } else {
      f0 = __uint_as_float(s[2 * i]);
      f1 = __uint_as_float(s[2 * i + 1]);
    }
    float x0, x1, e0, e1;
    unpack2(ffma2(pack2(f0, f1), c2, bias2), x0, x1);
    if (kPolyEvery > 0 && (i % (kPolyEvery > 0 ? kPolyEvery : 1)) == (kPolyEvery > 0 ? kPolyEvery : 1) - 1) {
      exp2_poly_pair(x0, x1, e0, e1);
    } else {
      e0 = (QMHA_KO & 8) ? x0 : ex2_approx(x0);
      e1 = (QMHA_KO & 8) ? x1 : ex2_approx(x1);
    }
    if constexpr (kMasked) {
      if (2 * i >= n_valid) e0 = 0.f;
      if (2 * i + 1 >= n_valid) e1 = 0.f;
    }
    if (!(QMHA_KO & 2)) lsum[i & 1] = fadd2(lsum[i & 1], pack2(e0, e1));
    p[i] = (QMHA_KO & 1) ? __float_as_uint(e0) : (kPv8 ? pack_u8x2(e0, e1) : (kBf16 ? pack_bf16x2(e0, e1) : pack_f16x2(e0, e1)));
  }
}

// ---- per-32-key-block scale mode (the reference's granularity, INT8 only) -----------------------
// A 64-key half-step spans two K/V blocks: columns [0,32) use (c0, b0), columns [32,64) (c1, b1)
// where c_g = sQ(row block)·sK(block g)·log2e/sqrt(d) and b_g = -(2^23·1.5·c_g + m) + log2 r_g folds the
// V block scale ratio r_g = sV_g / sV_max into the exponent; the un-scaled row sum is recovered per
// block as (sum of P')/r_g.  All of it is compile-time column selection: no per-element cost.
struct StepConsts {
  float c0, c1, lr0, lr1, ir0, ir1;
};
template <int B_, int E_>
struct Range {
  static constexpr int kB = B_, kE = E_;
};

template <bool kMasked>
__device__ __forceinline__ float tile_row_max_blk(uint32_t (&s)[kHN], float c0, float c1, int n_valid) {
  if constexpr (kMasked) {
#pragma unroll
    for (int i = 0; i < kHN; ++i)
      if (i >= n_valid) s[i] = (uint32_t)((kBiasMma ? kMagicI2F : 0) - (1 << 22));
  }
  int lo0 = max((int)s[0], (int)s[1]), lo1 = max((int)s[2], (int)s[3]);
  int hi0 = max((int)s[32], (int)s[33]), hi1 = max((int)s[34], (int)s[35]);
#pragma unroll
  for (int i = 4; i < 32; i += 4) {
    lo0 = max(max(lo0, (int)s[i + 0]), (int)s[i + 1]);
    lo1 = max(max(lo1, (int)s[i + 2]), (int)s[i + 3]);
    hi0 = max(max(hi0, (int)s[32 + i + 0]), (int)s[32 + i + 1]);
    hi1 = max(max(hi1, (int)s[32 + i + 2]), (int)s[32 + i + 3]);
  }
  // A block that holds no real key contributes nothing: its scale is the quantiser's 1e-8 floor, so a
  // scaled sentinel would read as ~0 and could lift the row max above every real (negative) logit.
  constexpr int kOff = kBiasMma ? kMagicI2F : 0;
  if (kMasked && n_valid <= kHN / 2) return (float)(max(lo0, lo1) - kOff) * c0;
  return fmaxf((float)(max(lo0, lo1) - kOff) * c0, (float)(max(hi0, hi1) - kOff) * c1);
}

template <bool kMasked, int kPolyEvery, int kBegin = 0, int kEnd = kHN / 2, bool kPv8 = false>
__device__ __forceinline__ void tile_row_exp_blk(const uint32_t (&s)[kHN], uint32_t (&p)[kHN / 2],
                                                 const StepConsts& k, float m_used, int n_valid,
                                                 uint64_t (&ls)[2], int one) {
  const float b0 = (QMHA_I2F == 5 ? k.lr0 - m_used : k.lr0 - fmaf(kMagicF, k.c0, m_used)) + (kPv8 ? kLog2P8 : 0.f);
  const float b1 = (QMHA_I2F == 5 ? k.lr1 - m_used : k.lr1 - fmaf(kMagicF, k.c1, m_used)) + (kPv8 ? kLog2P8 : 0.f);
  const uint64_t c2[2] = {pack2(k.c0, k.c0), pack2(k.c1, k.c1)};
  const uint64_t bias2[2] = {pack2(b0, b0), pack2(b1, b1)};
#pragma unroll
  for (int i = kBegin; i < kEnd; ++i) {
    constexpr int kHalf = kHN / 4;  // pairs per 32-key block
    const int g = i >= kHalf ? 1 : 0;
    float f0, f1;
    i2f_pair(s[2 * i], s[2 * i + 1], one, i, f0, f1, pace_token(p, i, (uint32_t)(one ^ 1)));
    float x0, x1, e0, e1;
    unpack2(ffma2(pack2(f0, f1), c2[g], bias2[g]), x0, x1);
    if (kPolyEvery > 0 && (i % (kPolyEvery > 0 ? kPolyEvery : 1)) == (kPolyEvery > 0 ? kPolyEvery : 1) - 1) {
      exp2_poly_pair(x0, x1, e0, e1);
    } else {
      e0 = (QMHA_KO & 8) ? x0 : ex2_approx(x0);
      e1 = (QMHA_KO & 8) ? x1 : ex2_approx(x1);
    }
    if constexpr (kMasked) {
      if (2 * i >= n_valid) e0 = 0.f;
      if (2 * i + 1 >= n_valid) e1 = 0.f;
    }
    if (!(QMHA_KO & 2)) ls[g] = fadd2(ls[g], pack2(e0, e1));
    p[i] = (QMHA_KO & 1) ? __float_as_uint(e0) : (kPv8 ? pack_u8x2(e0, e1) : pack_f16x2(e0, e1));
  }
}

// kPersist: one CTA per SM walks the work items (unit, 256-row query block) item0, item0 + gridDim.x, ... instead of one
// CTA per item.  The mbarriers, the TMEM allocation and the constant tiles live for the whole kernel; every role loops
// over the items on its own, so the producers prefetch Q and the K tiles of the next item and the MMA warps issue its first
// scores while the softmax warps still write the previous output (which is staged in the V^T ring only: the K ring and the
// Q tile are free by then).  Barrier phases continue across items: the rings run on the global tile number J0 + j, the
// once-per-item barriers on the item number k; N is a multiple of 256 (four half-steps), so s_full / p_full and the score
// buffers start every item exactly as they start the first one.
template <bool kInt8, int kD, int kPolyEvery, bool kBlk, bool kTrace, int kFa = 6, int kFb = 25, bool kBf16 = false,
          bool kPv8 = false, bool kPersist = false>
__global__ void __launch_bounds__(kThreads, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                const __grid_constant__ CUtensorMap tm_v, const __grid_constant__ CUtensorMap tm_o,
                const __grid_constant__ PeerMaps tm_peers, AttnParams prm) {
  using C = Cfg<kInt8, kD, kPv8>;
  // lazy rescale: the reference max is raised when a row max exceeds it by more than this many log2 units.  fp16 P
  // holds 2^4 with full relative precision; 8-bit P codes are rn(127.5 p), so p must stay <= 2.
  constexpr float kThr = kPv8 ? 1.0f : kRescaleThreshold;
  constexpr bool kSoftRing = QMHA_SOFT_RING == 2 || (QMHA_SOFT_RING == 1 && !kInt8);
  constexpr bool kLazyPv = QMHA_LAZY_PV == 2 || (QMHA_LAZY_PV == 1 && !kInt8);
  static_assert(!kBf16 || !kInt8, "bf16 is a variant of the 16-bit kernel");
  static_assert(!kPersist || (kMmaSplit == 2 && kLazyPv && !kSoftRing && !kTrace), "persistent kernel: two-warp MMA issue, lazy pv_done, committed rings");
  constexpr uint32_t kIdQK = kBf16 ? C::kIdescQKbf : C::kIdescQK;
  constexpr uint32_t kIdPV = kBf16 ? C::kIdescPVbf : C::kIdescPV;
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B operands need 1024-byte aligned tiles.
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + 2 * C::kTileBytesQK;
  uint8_t* sV = sK + C::kStagesK * C::kTileBytesQK;
  uint8_t* sBias = sV + C::kStagesV * C::kTileBytesV;   // constant operand tiles of the bias MMA (INT8 kernels)
  Barriers* bars = reinterpret_cast<Barriers*>(sBias + (C::kBias ? kBiasTileBytes : 0));
  float4* blk_tab0 = reinterpret_cast<float4*>(reinterpret_cast<uint8_t*>(bars) + 512);  // block mode only (persistent: two tables)

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // work items of this CTA: item0, item0 + item_stride, ... (one item without kPersist); every role keeps its own
  // copy of (unit, q_base) because the roles run up to one item apart
  const int nqb = kPersist ? prm.n_qblocks : 1;
  const int n_items = kPersist ? prm.n_items : 1;
  const int item0 = kPersist ? (int)blockIdx.x : 0;
  const int item_stride = kPersist ? (int)gridDim.x : 1;
  int unit = kPersist ? item0 / nqb : (int)blockIdx.y;                        // b * H + head
  int q_base = (kPersist ? item0 % nqb : (int)blockIdx.x) * (2 * kBM);        // first query row of the item
  auto set_item = [&](int it) {
    if constexpr (kPersist) { unit = it / nqb; q_base = (it % nqb) * (2 * kBM); }
  };
  const int n_tiles = prm.n_kv_tiles;
  const ErrCtx err_flag{prm.error_flag, prm.error_host, prm.launch_id};
  bool dead = false;

  if (err_flag.raised()) return;  // an earlier CTA of this launch already failed: drain the grid
  // Replicated output (peer_O): the CTAs of a wave finish together, so their NVLink stores would all drain at the same
  // time while nothing computes, and then everything computes while the links idle.  Every other CTA of the FIRST wave
  // starts late by about the drain time of half a wave: from then on one half of the SMs computes while the other half's
  // stores drain (see launch_cfg for the estimate).
  if (prm.stagger_ns != 0u) {
    const unsigned lin = kPersist ? blockIdx.x : blockIdx.y * gridDim.x + blockIdx.x;
    if (lin < prm.stagger_ctas && (lin & 1u)) {
      unsigned long long t0, t1;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
      do {
        __nanosleep(2000);
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      } while (t1 - t0 < (unsigned long long)prm.stagger_ns);
    }
  }
  const long long t_entry = (prm.cycles != nullptr && threadIdx.x == 0) ? clock64() : 0;
  // traced build: a CTA from the middle of the run (steady state, warm caches); phase stamps of its
  // warp 0 go behind the per-step stamps: entry, setup done, first scores, last P, O final, stores, exit
  const bool traced_cta = kTrace && blockIdx.x == gridDim.x / 2 && blockIdx.y == gridDim.y / 2;
  long long* phase = kTrace ? prm.trace + (size_t)9 * prm.n_half_steps * 4 : nullptr;
  if (traced_cta && threadIdx.x == 0) phase[0] = clock64();

  if (warp == kTmaWarp && lane == 0) {
    prefetch_tmap(&tm_q);
    prefetch_tmap(&tm_k);
    prefetch_tmap(&tm_v);
    prefetch_tmap(&tm_o);
    mbar_init(&bars->q_full, 1);
    for (int i = 0; i < 4; ++i) {
      mbar_init(&bars->k_full[i], 1);
      mbar_init(&bars->k_empty[i], kMmaSplit);
      mbar_init(&bars->v_full[i], 1);
      mbar_init(&bars->v_empty[i], kMmaSplit);
    }
    for (int t = 0; t < 2; ++t) {
      for (int b = 0; b < 2; ++b) {
        mbar_init(&bars->s_full[t][b], 1);
        mbar_init(&bars->p_full[t][b], 128);
        mbar_init(&bars->pv_done[t][b], 1);
      }
      mbar_init(&bars->o_final[t], 1);
      mbar_init(&bars->s0_read[t], 128);
    }
    mbar_init(&bars->qk_done, kMmaSplit);
    mbar_init(&bars->epi_done, 8);
    fence_mbar_init();
    if constexpr (kMmaSplit == 2) {
      // Q and the first K tiles are requested right here, before the CTA-wide barrier: the TMEM
      // allocation and the barrier itself then overlap the ~1.5 k clk of TMA latency.
      mbar_arrive_expect_tx(&bars->q_full, 2 * C::kTileBytesQK);
#pragma unroll
      for (int t = 0; t < 2; ++t)
#pragma unroll
        for (int sub = 0; sub < C::kSubQK; ++sub)
          tma_load_2d(sQ + t * C::kTileBytesQK + sub * C::kSubBytesQK, &tm_q, &bars->q_full,
                      sub * (C::kAtomQK / C::kEltQK), unit * prm.n_pad + q_base + t * kBM);
      for (int j = 0; j < C::kStagesK && j < n_tiles; ++j) {
        mbar_arrive_expect_tx(&bars->k_full[j], C::kTileBytesQK);
#pragma unroll
        for (int sub = 0; sub < C::kSubQK; ++sub)
          tma_load_2d(sK + j * C::kTileBytesQK + sub * C::kSubBytesQK, &tm_k, &bars->k_full[j],
                      sub * (C::kAtomQK / C::kEltQK), unit * prm.n_pad + j * kBN);
      }
    }
  }
  if constexpr (C::kBias) {   // 384 threads x 16 B = 6 KB: A tile all 768.0 (0x6200), B tile all 1024.0 (0x6400)
    const uint32_t w = threadIdx.x < 256 ? 0x62006200u : 0x64006400u;
    reinterpret_cast<uint4*>(sBias)[threadIdx.x] = make_uint4(w, w, w, w);
    fence_proxy_async_smem();   // generic-proxy writes -> visible to the tensor pipe's (async proxy) operand reads
  }
  if (warp == kAllocWarp) {
    tmem_alloc(&bars->tmem_base, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;

  if (warp >= 8) {
   // setmaxnreg sits inside the role branch (which never re-joins the softmax code before the
   // final barrier) so ptxas allocates registers per role.
   asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegsService));
   if (kMmaSplit == 2 && warp == kTmaWarp) {
    // ======================================================================== TMA producer: Q, K ring
    if (lane == 0) {  // Q and K tiles 0 .. kStagesK-1 of the first item were requested during the CTA setup
      int k = 0;
      for (int it = item0; it < n_items && !dead && !((QMHA_PERSIST_SINGLE & 4) && k > 0); it += item_stride, ++k) {
        if (kPersist && k > 0) {
          set_item(it);
          // the Q tiles are dead once every Q.K^T of the previous item has retired
          mbar_wait(&bars->qk_done, (uint32_t)(k - 1) & 1, err_flag, 103, dead);
          mbar_arrive_expect_tx(&bars->q_full, 2 * C::kTileBytesQK);
#pragma unroll
          for (int t = 0; t < 2; ++t)
#pragma unroll
            for (int sub = 0; sub < C::kSubQK; ++sub)
              tma_load_2d(sQ + t * C::kTileBytesQK + sub * C::kSubBytesQK, &tm_q, &bars->q_full,
                          sub * (C::kAtomQK / C::kEltQK), unit * prm.n_pad + q_base + t * kBM);
        }
        const int k_row0 = unit * prm.n_pad;
        const int J0 = k * n_tiles;
        for (int j = (k == 0 ? C::kStagesK : 0); j < n_tiles; ++j) {
          const int J = J0 + j;
          const int st = J % C::kStagesK;
          const uint32_t ph = (uint32_t)(J / C::kStagesK);
          if (J >= C::kStagesK) mbar_wait(&bars->k_empty[st], (ph - 1) & 1, err_flag, 101, dead);
          if (QMHA_KO & 32) { mbar_arrive(&bars->k_full[st]); continue; }   // timing experiment: no K traffic after the first ring fill
          mbar_arrive_expect_tx(&bars->k_full[st], C::kTileBytesQK);
#pragma unroll
          for (int sub = 0; sub < C::kSubQK; ++sub)
            tma_load_2d(sK + st * C::kTileBytesQK + sub * C::kSubBytesQK, &tm_k, &bars->k_full[st],
                        sub * (C::kAtomQK / C::kEltQK), k_row0 + j * kBN);
        }
      }
    }
   } else if (kMmaSplit == 2 && warp == kTmaWarpV) {
    // ======================================================================== TMA producer: V^T ring
    if (lane == 0) {
      int k = 0;
      for (int it = item0; it < n_items && !dead && !((QMHA_PERSIST_SINGLE & 4) && k > 0); it += item_stride, ++k) {
        if (kPersist && k > 0) {
          set_item(it);
          // the previous item's output is staged in this ring: wait until its TMA stores have read the tiles
          mbar_wait(&bars->epi_done, (uint32_t)(k - 1) & 1, err_flag, 104, dead);
        }
        const int v_row = unit * kD;
        const int J0 = k * n_tiles;
        for (int j = 0; j < n_tiles; ++j) {
          const int J = J0 + j;
          const int st = J % C::kStagesV;
          const uint32_t ph = (uint32_t)(J / C::kStagesV);
          if (J >= C::kStagesV) mbar_wait(&bars->v_empty[st], (ph - 1) & 1, err_flag, 102, dead);
          if ((QMHA_KO & 32) && J >= C::kStagesV) { mbar_arrive(&bars->v_full[st]); continue; }   // same for V^T
          mbar_arrive_expect_tx(&bars->v_full[st], C::kTileBytesV);
#pragma unroll
          for (int sub = 0; sub < C::kSubTilesV; ++sub)
            tma_load_2d(sV + st * C::kTileBytesV + sub * C::kSubBytesV, &tm_v, &bars->v_full[st],
                        j * kBN + sub * 64, v_row);
        }
      }
    }
   } else if (is_mma_warp(warp)) {
    // ======================================================================== MMA issuer
    // The whole warp runs this code with uniform control flow (so descriptors live in uniform
    // registers); only the tcgen05.mma / tcgen05.commit instructions are issued by one elected
    // lane.  Issuing from inside a divergent `if (lane == 0)` costs ~100 clk per MMA.
    {
      const uint32_t sQ_a = smem_u32(sQ), sK_a = smem_u32(sK), sV_a = smem_u32(sV);
      const bool leader = elect_one() != 0;
      const bool do_mma = leader && !prm.debug_no_mma;
      // Base descriptors are built once; per MMA only the 14-bit address field moves (one add).
      const uint64_t bias_a = make_smem_desc(smem_u32(sBias), 32), bias_b = make_smem_desc(smem_u32(sBias) + 4096, 32);
      const uint64_t q_desc0 = make_smem_desc(sQ_a, C::kAtomQK);
      const uint64_t k_desc0 = make_smem_desc(sK_a, C::kAtomQK);
      const uint64_t v_desc0 = make_smem_desc(sV_a, 128);
      // S_t[buf] = Q_t · K(stage st, key half `half`)^T
      auto issue_qk = [&](int t, int buf, int st, int half) {
        const uint32_t d_tmem = tmem_base + (t ? kColS1 : kColS0) + buf * kHN;
        const uint64_t a0 = advance_smem_desc(q_desc0, (uint32_t)t * C::kTileBytesQK);
        const uint64_t b0 = advance_smem_desc(k_desc0, (uint32_t)st * C::kTileBytesQK + (uint32_t)half * C::kHalfBytesQK);
        if constexpr (C::kBias) {   // accumulator := 0x4B400000 everywhere; the int8 dot products accumulate onto it
          if (do_mma) mma_f16_ss(d_tmem, bias_a, bias_b, C::kIdescBias, 0u);
        }
#pragma unroll
        for (int ks = 0; ks < C::kStepsQK; ++ks) {
          constexpr int kDummy = 0; (void)kDummy;
          const uint32_t off = (uint32_t)((ks * 32) / C::kAtomQK) * C::kSubBytesQK +
                               (uint32_t)((ks * 32) % C::kAtomQK);
          const uint64_t a = advance_smem_desc(a0, off);
          const uint64_t b = advance_smem_desc(b0, off);
          if (do_mma) {
            if constexpr (kInt8) mma_i8_ss(d_tmem, a, b, kIdQK, (C::kBias || ks > 0) ? 1u : 0u);
            else mma_f16_ss(d_tmem, a, b, kIdQK, ks > 0);
          }
        }
      };
      // O_t (+)= P_t · V(stage st, key half `half`);  P_t sits in the first 32 columns of S_t[pbuf]
      auto issue_pv = [&](int t, int pbuf, int st, int half, bool accumulate) {
        const uint32_t d_tmem = tmem_base + (t ? kColO1 : kColO0);
        const uint32_t p_tmem = tmem_base + (t ? kColS1 : kColS0) + pbuf * kHN;
        const uint64_t b0 = advance_smem_desc(v_desc0, (uint32_t)st * C::kTileBytesV + (uint32_t)half * C::kHalfOffV);
#pragma unroll
        for (int ks = 0; ks < C::kStepsPV; ++ks) {
          const uint64_t b = advance_smem_desc(b0, (uint32_t)ks * 32);
          if (do_mma) {
            if constexpr (kPv8) mma_i8_ts(d_tmem, p_tmem + ks * 8, b, C::kIdescPV8, (accumulate || ks > 0) ? 1u : 0u);
            else mma_f16_ts(d_tmem, p_tmem + ks * 8, b, kIdPV, (accumulate || ks > 0) ? 1u : 0u);
          }
        }
      };

      auto commit = [&](uint64_t* bar) { if (leader) mma_commit(bar); };
      const int n_half = prm.n_half_steps;
      // This warp issues, for query tile `mt`, every half-step i with owns(i).  Half-step i means:
      // O += P(i)·V(i), then S(i+3) into the score buffer P(i) just left; the scores of half-steps 0..2
      // belong to the owners of "steps" -3..-1.
      constexpr int kStride = kMmaSplit == 2 ? 1 : 2;
      const int mt = kMmaSplit == 2 ? (warp == 8 ? 0 : 1) : ((warp - 8) >> 1);
      const int first = kMmaSplit == 2 ? 0 : ((warp - 8) & 1);
      auto owns = [&](int step) { return kStride == 1 || (step & 1) == first; };

      // ---- folded TMA production (kMmaSplit == 4): warp 8 = Q + K ring, warp 10 = V^T ring
      const bool k_prod = kMmaSplit == 4 && warp == 8, v_prod = kMmaSplit == 4 && warp == 10;
      auto load_k = [&](int j) {   // K tile j -> stage j % kStagesK (the stage must be free)
        const int st = j % C::kStagesK;
        if (leader) {
          mbar_arrive_expect_tx(&bars->k_full[st], C::kTileBytesQK);
#pragma unroll
          for (int sub = 0; sub < C::kSubQK; ++sub)
            tma_load_2d(sK + st * C::kTileBytesQK + sub * C::kSubBytesQK, &tm_k, &bars->k_full[st],
                        sub * (C::kAtomQK / C::kEltQK), unit * prm.n_pad + j * kBN);
        }
      };
      auto load_v = [&](int j) {
        const int st = j % C::kStagesV;
        if (leader) {
          mbar_arrive_expect_tx(&bars->v_full[st], C::kTileBytesV);
#pragma unroll
          for (int sub = 0; sub < C::kSubTilesV; ++sub)
            tma_load_2d(sV + st * C::kTileBytesV + sub * C::kSubBytesV, &tm_v, &bars->v_full[st],
                        j * kBN + sub * 64, unit * kD);
        }
      };
      if (k_prod) {
        if (leader) {
          mbar_arrive_expect_tx(&bars->q_full, 2 * C::kTileBytesQK);
#pragma unroll
          for (int t = 0; t < 2; ++t)
#pragma unroll
            for (int sub = 0; sub < C::kSubQK; ++sub)
              tma_load_2d(sQ + t * C::kTileBytesQK + sub * C::kSubBytesQK, &tm_q, &bars->q_full,
                          sub * (C::kAtomQK / C::kEltQK), unit * prm.n_pad + q_base + t * kBM);
        }
        for (int j = 0; j < C::kStagesK && j < n_tiles; ++j) load_k(j);
      }
      if (v_prod)
        for (int j = 0; j < C::kStagesV && j < n_tiles; ++j) load_v(j);
      // At the top of the iteration that handles tile j (even parity, j >= 1): tile j-1 was last read one
      // or two half-steps ago by every issuing warp; once its stage is free, tile j-1+stages goes in.
      auto refill = [&](int j) {
        if (j < 1) return;
        if (k_prod && j - 1 + C::kStagesK < n_tiles) {
          const int x = j - 1;
          mbar_wait(&bars->k_empty[x % C::kStagesK], (uint32_t)(x / C::kStagesK) & 1, err_flag, 101, dead);
          load_k(x + C::kStagesK);
        }
        if (v_prod && j - 1 + C::kStagesV < n_tiles) {
          const int x = j - 1;
          mbar_wait(&bars->v_empty[x % C::kStagesV], (uint32_t)(x / C::kStagesV) & 1, err_flag, 102, dead);
          load_v(x + C::kStagesV);
        }
      };

      // Items of this CTA, k = 0, 1, ...: G = global half-step of the item's step 0 (score / P buffers and the phases of
      // s_full / p_full follow G + i), J0 = global K / V tile of its tile 0 (ring stages and their phases follow J0 + j).
      int k = 0;
      for (int it = item0; it < n_items && !dead && !((QMHA_PERSIST_SINGLE & 2) && k > 0); it += item_stride, ++k) {
      constexpr int G = 0;                        // (persistent: n_half % 4 == 0, every item starts in an even phase on buffer 0)
      const int J0 = kPersist ? k * n_tiles : 0;
      mbar_wait(&bars->q_full, (uint32_t)k & 1, err_flag, 201, dead);
      tc_fence_after();
      __syncwarp();
      auto qk_step = [&](int in, int wait_site, long long* stamp = nullptr) {   // S_mt(in): wait for its K tile, issue, signal, release
        const int jn = J0 + (in >> 1), halfn = in & 1, stn = jn % C::kStagesK;
        mbar_wait(&bars->k_full[stn], (uint32_t)(jn / C::kStagesK) & 1, err_flag, wait_site, dead);
        tc_fence_after();
        if (kTrace && stamp) *stamp = clock64();
        issue_qk(mt, (G + in) & 1, stn, halfn);
        commit(&bars->s_full[mt][(G + in) & 1]);
        // last read of this K tile by this warp: one arrival per MMA warp frees the stage
        if (!kSoftRing && (kStride == 2 || halfn == 1 || in == n_half - 1)) commit(&bars->k_empty[stn]);
      };
      for (int in = 0; in < 3 && in < n_half; ++in) {
        if (!owns(in - 3)) continue;
        if (in == 2) {  // buffer 0 is reusable once the softmax warps hold S(0) in registers
          mbar_wait(&bars->s0_read[mt], (uint32_t)k & 1, err_flag, 208, dead);
          tc_fence_after();
        }
        qk_step(in, 202);
      }
      if (first + 3 >= n_half) commit(&bars->qk_done);  // this warp issues no further Q·K^T
      for (int i = first; i < n_half; i += kStride) {
        const int j = J0 + (i >> 1), half = i & 1;
        const int st = j % C::kStagesV;
        long long* trm = kTrace ? prm.trace + (size_t)8 * n_half * 4 + (size_t)i * 4 : nullptr;
        const bool tracer = traced_cta && lane == 0 && mt == 0;
        if constexpr (kMmaSplit == 4) refill(j);   // only warps 8 / 10 do anything here (even i)
        mbar_wait(&bars->v_full[st], (uint32_t)(j / C::kStagesV) & 1, err_flag, 203, dead);
        mbar_wait(&bars->p_full[mt][(G + i) & 1], (uint32_t)((G + i) >> 1) & 1, err_flag, 204, dead);
        if (kStride == 2 && i > 0)  // the other warp's P·V(i-1) must have retired before O is touched again
          mbar_wait(&bars->pv_done[mt][(i - 1) & 1], (uint32_t)((i - 1) >> 1) & 1, err_flag, 209, dead);
        tc_fence_after();
        if (tracer) trm[0] = clock64();
        issue_pv(mt, (G + i + 1) & 1, st, half, i > 0);
        if (!kLazyPv || i + 3 >= n_half) commit(&bars->pv_done[mt][half]);   // no Q.K^T behind the last three
        if (i == n_half - 1) commit(&bars->o_final[mt]);
        if (!kSoftRing && (kStride == 2 || half == 1 || i == n_half - 1)) commit(&bars->v_empty[st]);
        if (tracer) trm[1] = clock64();
        if (i + 3 < n_half) {
          qk_step(i + 3, 205, tracer ? trm + 2 : nullptr);
          if (i + 3 + kStride >= n_half) commit(&bars->qk_done);  // that was this warp's last Q·K^T
        }
        if (tracer) trm[3] = clock64();
      }
      }
    }
   }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRegsSoftmax));
    // ======================================================================== softmax warpgroups
    const int t = warp >> 2;                       // query tile handled by this warpgroup
    const int row_in_tile = (warp & 3) * 32 + lane;
    const uint32_t lane_addr = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t tS = tmem_base + lane_addr + (t ? kColS1 : kColS0);
    const uint32_t tO = tmem_base + lane_addr + (t ? kColO1 : kColO0);
    // measurement aid (QMHA_DEBUG_NO_MMA): P goes to the unused O columns so the score buffers keep sane data
    const uint32_t tP = prm.debug_no_mma ? tO : tS;

    static_assert(!kBlk || kInt8, "block scales only exist for the INT8 variant");
    // Persistent kernel: the host only selects it when n_half % 4 == 0 (N a multiple of 256), so every item starts on
    // score buffer 0 with the s_full / p_full barriers in an even phase — the softmax warps need no global step counter
    // (G = 0) and keep a single extra live value across the loop, the item number k.
    float pre_c = 1.f, pre_os = 1.f;   // persistent block mode: the next item's row-block scale and V scale, fetched early
    for (int k = 0; k < ((QMHA_PERSIST_SINGLE & 1) ? 1 : 0x7fffffff); ++k) {
    if (item0 + k * item_stride >= n_items) break;
    if (kPersist && k > 0) set_item(item0 + k * item_stride);
    constexpr int G = 0, J0 = 0;
    float4* blk_tab = blk_tab0 + ((kPersist && (k & 1)) ? prm.n_pad / 32 : 0);   // persistent: two tables, alternating
    float c = prm.scale_log2;  // log2(e) / sqrt(d)
    float out_scale = 1.0f;
    const int nblk = prm.n_pad / 32;
    // fills `tab` with the block constants of `unit_x` and returns {sQ of this row's block in item (unit_x, qb_x), sV_max}
    auto fetch_item_consts = [&](int unit_x, int qb_x, float4* tab, float& c_x, float& os_x) {
      const int qblk = (qb_x + t * kBM + row_in_tile) >> 5;
      c_x = __ldg(prm.blk_scales + (size_t)unit_x * nblk + qblk);
      os_x = __ldg(prm.blk_vmax + unit_x);
      const float* gk = prm.blk_scales + ((size_t)prm.units + unit_x) * nblk;
      const float2* ga = reinterpret_cast<const float2*>(prm.blk_aux) + (size_t)unit_x * nblk;
      for (int g = threadIdx.x; g < nblk; g += 256) {
        const float2 a = __ldg(ga + g);
        tab[g] = make_float4(__ldg(gk + g), a.x, a.y, 0.f);
      }
    };
    if constexpr (kBlk) {
      // c = log2e/sqrt(d) * sQ of this row's 32-row block; the K/V block factors come per step.  The per-32-key-block
      // constants {sK, log2 r, 1/r, -} of the unit are staged in shared memory (256 softmax threads), so the per-step
      // lookups are two broadcast LDS.128.  Persistent kernel: items after the first find all of it fetched by the
      // previous item, under its wait for the last P.V.
      if (kPersist && k > 0) {
        c *= pre_c;
        out_scale = pre_os;
      } else {
        float cq;
        fetch_item_consts(unit, q_base, blk_tab, cq, out_scale);
        c *= cq;
      }
    } else if constexpr (kInt8) {
      const float sq = prm.scales[unit];
      const float sk = prm.scales[prm.units + unit];
      out_scale = prm.scales[2 * prm.units + unit];
      c = sq * sk * c;
    }

    const int one = prm.one;   // run-time 1 (keeps the IMAD form of the int->float add from being folded)
    float m_used = -INFINITY;
    uint64_t lsum[2] = {0ull, 0ull};  // four fp32 partial row sums (packed pairs); block mode: l_acc
    float l_acc = 0.f;
    const int n_half = prm.n_half_steps;
    // trace layout: [8 softmax warps + MMA][n_half][4]; softmax: step start, before / after the
    // s_full wait of the prefetch, P(i) published
    const bool tracer = traced_cta && lane == 0;
    long long* tr = kTrace ? prm.trace + (size_t)warp * n_half * 4 : nullptr;
    if (tracer && warp == 0) phase[1] = clock64();

    if constexpr (kBlk && !kPersist) asm volatile("bar.sync 1, 256;" ::: "memory");
    if constexpr (kPersist) {
      // one rendezvous of the 256 softmax threads per item: it orders the table fill (the other table may still be in use
      // by warps that are finishing the previous item) and makes the decision to drain after a failure uniform
      uint32_t any_dead;
      asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.u32 q, %1, 0;\n\tbar.red.or.pred p, 1, 256, q;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(any_dead) : "r"((uint32_t)dead) : "memory");
      if (any_dead) break;
    }

    // The softmax loop is software-pipelined over half-steps.  While the exponentials of step i
    // run, (1) P(i-1) — finished at the end of the previous step — is stored to TMEM and handed
    // to the MMA warp (kP pairs into the step, so the MUFU queue never drains across the step
    // boundary), (2) the scores of step i+1 (ready long ago: S runs three half-steps ahead) are
    // fetched from TMEM, their row max is taken and the warp votes whether step i+1 must raise
    // the reference max.  Two register arrays for S (sA, sB) and for P (pA, pB) alternate.
    auto load_consts = [&](int i) {
      StepConsts k{c, c, 0.f, 0.f, 1.f, 1.f};
      if constexpr (kBlk) {
        const float4 a = blk_tab[2 * i], b = blk_tab[2 * i + 1];
        k.c0 = c * a.x; k.lr0 = a.y; k.ir0 = a.z;
        k.c1 = c * b.x; k.lr1 = b.y; k.ir1 = b.z;
      }
      return k;
    };
    auto row_max = [&](uint32_t (&sx)[kHN], const StepConsts& k, bool masked, int n_valid) {
      if constexpr (kBlk) {
        return masked ? tile_row_max_blk<true>(sx, k.c0, k.c1, n_valid) : tile_row_max_blk<false>(sx, k.c0, k.c1, kHN);
      } else {
        return masked ? tile_row_max<kInt8, true>(sx, c, n_valid) : tile_row_max<kInt8, false>(sx, c, kHN);
      }
    };
    // block mode: add this step's per-block sums of P' = p*r back as p
    auto fold_sums = [&](const uint64_t (&ls)[2], const StepConsts& k) {
      float a0, a1, b0, b1;
      unpack2(ls[0], a0, a1);
      unpack2(ls[1], b0, b1);
      l_acc = fmaf(a0 + a1, k.ir0, fmaf(b0 + b1, k.ir1, l_acc));
    };
    // exponentials of pairs [kBegin, kEnd) of an unmasked step
    auto exps = [&](auto range, const uint32_t (&sx)[kHN], uint32_t (&p)[kHN / 2], const StepConsts& k,
                    uint64_t (&ls)[2]) {
      constexpr int kB = decltype(range)::kB, kE = decltype(range)::kE;
      if (QMHA_KO & 16) {   // timing experiment: no softmax arithmetic at all (tensor side alone)
#pragma unroll
        for (int q = kB; q < kE; ++q) p[q] = sx[2 * q];
        return;
      }
      if constexpr (kBlk) tile_row_exp_blk<false, kPolyEvery, kB, kE, kPv8>(sx, p, k, m_used, kHN, ls, one);
      else tile_row_exp<kInt8, false, kPolyEvery, kB, kE, kBf16, kPv8>(sx, p, c, m_used, kHN, lsum, one);
    };

    // wait for S_t(i) and start its TMEM->register load (completion: tmem_wait_ld).  `probed` is the
    // result of an earlier non-blocking probe of the same barrier phase (keeps the barrier-unit round
    // trip off the critical path when the scores are already there, which is the normal case).
    // kSoftRing: the scores of half-step r have arrived (s_full observed) -> hand ring stages back.
    const bool releaser = kSoftRing && (threadIdx.x & 127) == 0;
    auto on_scores = [&](int r) {
      if (!kSoftRing || !releaser) return;
      if (r & 1) mbar_arrive(&bars->k_empty[(J0 + (r >> 1)) % C::kStagesK]);              // K tile r/2: both halves read
      else if (r >= 4) mbar_arrive(&bars->v_empty[(J0 + ((r - 4) >> 1)) % C::kStagesV]);   // P.V(r-3) retired: V tile (r-4)/2
    };
    // "P.V(j) has retired": s_full(j+3) when that Q.K^T exists, the step's own pv_done commit otherwise
    const int pv_c0 = n_half > 3 ? n_half - 3 : 0;   // first half-step with a pv_done commit (kLazyPv)
    // pv_done[b] completes once per committed step of local parity b: the last three half-steps of every item, i.e.
    // twice per item for the parity of step n_half - 1 and once for the other
    auto pv_base = [&](int par) { return kPersist ? k * ((par == ((n_half - 1) & 1)) ? 2 : 1) : 0; };
    auto wait_pv = [&](int j, int site) {
      if (kLazyPv && j + 3 < n_half)
        mbar_wait(&bars->s_full[t][(G + j + 3) & 1], (uint32_t)((G + j + 3) >> 1) & 1, err_flag, site, dead);
      else
        mbar_wait(&bars->pv_done[t][j & 1], (uint32_t)(pv_base(j & 1) + ((kLazyPv ? j - pv_c0 : j) >> 1)) & 1, err_flag, site, dead);
      tc_fence_after();
    };
    auto probe = [&](int i) {
      return mbar_test_wait(&bars->s_full[t][(G + i) & 1], (uint32_t)((G + i) >> 1) & 1) != 0;
    };
    auto fetch = [&](int i, uint32_t (&dst)[kHN], bool probed = false) {
      const int buf = (G + i) & 1;
      if (!probed) mbar_wait(&bars->s_full[t][buf], (uint32_t)((G + i) >> 1) & 1, err_flag, 301 + t, dead);
      tc_fence_after();
      on_scores(i);
      tmem_ld32(tS + buf * kHN, &dst[0]);
      tmem_ld32(tS + buf * kHN + 32, &dst[32]);
    };
    // P_t(i) -> TMEM, over the score buffer of step i+1 (its scores are in registers by now),
    // then tell the MMA warp.
    // fp16 / bf16 P: 32 columns; 8-bit P: the pairs' 16-bit halves are merged into 16 columns first
    auto store_p = [&](uint32_t taddr, const uint32_t (&p)[kHN / 2]) {
      if constexpr (kPv8) {
        uint32_t q8[kHN / 4];
#pragma unroll
        for (int j = 0; j < kHN / 4; ++j) q8[j] = __byte_perm(p[2 * j], p[2 * j + 1], 0x5410);
        tmem_st16(taddr, q8);
      } else {
        tmem_st32(taddr, &p[0]);
      }
    };
    auto publish = [&](int i, const uint32_t (&p)[kHN / 2]) {
      store_p(tP + ((G + i + 1) & 1) * kHN, p);
      tmem_wait_st();
      tc_fence_before();
      mbar_arrive(&bars->p_full[t][(G + i) & 1]);
      if (tracer) tr[i * 4 + 3] = clock64();
    };
    // Raise the reference max for step i (and, for i > 0, rescale l and the O rows in TMEM).
    // Called only when the warp voted for it: the old max is kept unless the new row max exceeds
    // it by more than 2^kRescaleThreshold.  P(i-1) must have been published before.
    auto raise_max = [&](int i, float mt) {
      const bool need = mt > m_used + kThr;
      const float m_new = need ? mt : m_used;
      if (i > 0) {
        const float alpha = need ? ex2_approx(m_used - m_new) : 1.0f;
        const uint64_t alpha2 = pack2(alpha, alpha);
        lsum[0] = fmul2(lsum[0], alpha2);
        lsum[1] = fmul2(lsum[1], alpha2);
        l_acc *= alpha;
        // P·V(i-1) retired?  (one barrier per step parity: P·V(i-3) is known to be complete, so
        // the phase cannot alias)
        wait_pv(i - 1, 311 + t);
#pragma unroll
        for (int ch = 0; ch < kD / 32; ++ch) {
          uint32_t o[32];
          tmem_ld32(tO + ch * 32, o);
          tmem_wait_ld();
#pragma unroll
          for (int q = 0; q < 32; ++q)   // INT8 P.V keeps O as int32: rescale through a float round trip (rare path)
            o[q] = kPv8 ? (uint32_t)__float2int_rn((float)(int)o[q] * alpha) : __float_as_uint(__uint_as_float(o[q]) * alpha);
          tmem_st32(tO + ch * 32, o);
        }
        tmem_wait_st();
      }
      m_used = m_new;
    };
    auto vote_raise = [&](float mt) { return __any_sync(0xffffffffu, mt > m_used + kThr) != 0; };

    // One pipelined step i < n_half-1 (always unmasked): exponentials of `cur` into `p`; publishes
    // `p_prev` = P(i-1) (if i > 0); if kPrefetch, fetches S(i+1) (unmasked, i.e. i+1 < n_half-1)
    // into `nxt` with row max and vote.
    using Yes = std::true_type;
    using No = std::false_type;
    // One pipelined step i < n_half-1 (always unmasked): exponentials of `cur` into `p`; publishes
    // `p_prev` = P(i-1) (if i > 0); if kPrefetch, fetches S(i+1) (unmasked, i.e. i+1 < n_half-1)
    // into `nxt` with row max and vote.
    // (A separate straight-line "fast path" selected by a vote at the top of the step was measured
    // slower: the probe -> vote -> branch chain is exposed latency.)
    auto pipe_step = [&](auto prefetch, int i, uint32_t (&cur)[kHN], float mt_cur, bool raise_cur,
                         const StepConsts& kc, uint32_t (&p)[kHN / 2], const uint32_t (&p_prev)[kHN / 2],
                         uint32_t (&nxt)[kHN], float& mt_nxt, bool& raise_nxt, StepConsts& kn) {
      constexpr bool kPrefetch = decltype(prefetch)::value;
      if (tracer) tr[i * 4 + 0] = clock64();
      bool published = i == 0;
      if (raise_cur) {  // rare after the first steps
        if (!published) publish(i - 1, p_prev);
        published = true;
        raise_max(i, mt_cur);
      }
      uint64_t ls[2] = {0ull, 0ull};
      bool s_ready = false;
      if constexpr (kPrefetch) s_ready = probe(i + 1);
      // Split points (in exp2 pairs) chosen so that this warp never sits on a busy unit: the P store
      // is issued at kFa and only waited for / signalled three pairs later; the two halves of the score
      // fetch (TMEM reads run at 64 B/clk: 64 clk per 32-column load) are issued six pairs apart.
      constexpr int kA1 = kFa + 3, kA2 = kFa + 9;
      static_assert(kA2 < kFb && kFb < kHN / 2, "split points out of order");
      exps(Range<0, kFa>{}, cur, p, kc, ls);
      if (!published) store_p(tP + ((G + i) & 1) * kHN, p_prev);  // P(i-1) over the S(i) buffer
      exps(Range<kFa, kA1>{}, cur, p, kc, ls);
      if (!published) {
        tmem_wait_st();
        tc_fence_before();
        mbar_arrive(&bars->p_full[t][(G + i - 1) & 1]);
        if (tracer) tr[(i - 1) * 4 + 3] = clock64();
      }
      if constexpr (kPrefetch) {
        if (tracer) tr[i * 4 + 1] = clock64();
        const int buf = (G + i + 1) & 1;
        if (!s_ready) mbar_wait(&bars->s_full[t][buf], (uint32_t)((G + i + 1) >> 1) & 1, err_flag, 301 + t, dead);
        tc_fence_after();
        on_scores(i + 1);
        tmem_ld32(tS + buf * kHN, &nxt[0]);
        if (tracer) tr[i * 4 + 2] = clock64();
        kn = load_consts(i + 1);
        exps(Range<kA1, kA2>{}, cur, p, kc, ls);
        tmem_ld32(tS + buf * kHN + 32, &nxt[32]);
        exps(Range<kA2, kFb>{}, cur, p, kc, ls);
        tmem_wait_ld();
        mt_nxt = (QMHA_KO & 4) ? __uint_as_float(nxt[0] ^ nxt[63]) * 1e-30f : row_max(nxt, kn, false, kHN);
        raise_nxt = vote_raise(mt_nxt);
        exps(Range<kFb, kHN / 2>{}, cur, p, kc, ls);
      } else {
        exps(Range<kA1, kHN / 2>{}, cur, p, kc, ls);
      }
      if constexpr (kBlk) fold_sums(ls, kc);
    };
    // last step (i == n_half-1): may cover fewer than 64 existing keys; not pipelined
    auto last_step = [&](int i, uint32_t (&cur)[kHN], uint32_t (&p)[kHN / 2], const uint32_t (&p_prev)[kHN / 2]) {
      if (tracer) tr[i * 4 + 0] = clock64();
      fetch(i, cur);
      const StepConsts kl = load_consts(i);
      const int n_valid = prm.N - i * kHN;  // >= 1
      const bool masked = n_valid < kHN;
      tmem_wait_ld();
      if (i > 0) publish(i - 1, p_prev);  // P(i-1) goes over the S(i) buffer: only after the load
      const float mt = row_max(cur, kl, masked, n_valid);
      if (vote_raise(mt)) raise_max(i, mt);
      if constexpr (kBlk) {
        uint64_t ls[2] = {0ull, 0ull};
        if (masked) tile_row_exp_blk<true, kPolyEvery, 0, kHN / 2, kPv8>(cur, p, kl, m_used, n_valid, ls, one);
        else tile_row_exp_blk<false, kPolyEvery, 0, kHN / 2, kPv8>(cur, p, kl, m_used, kHN, ls, one);
        fold_sums(ls, kl);
      } else {
        if (masked) tile_row_exp<kInt8, true, kPolyEvery, 0, kHN / 2, kBf16, kPv8>(cur, p, c, m_used, n_valid, lsum, one);
        else tile_row_exp<kInt8, false, kPolyEvery, 0, kHN / 2, kBf16, kPv8>(cur, p, c, m_used, kHN, lsum, one);
      }
      if (tracer) tr[i * 4 + 2] = clock64();
      // P(i) goes over P(i-2).  In the steady state the fetch of S(i+1) proves that P·V(i-2) has
      // retired; there is no S(i+1) here, so ask the tensor pipe directly.
      if (i >= 2) wait_pv(i - 2, 331 + t);
      publish(i, p);
    };

    uint32_t sA[kHN], sB[kHN], pA[kHN / 2], pB[kHN / 2];
    if (prm.debug_no_mma) {
#pragma unroll
      for (int q = 0; q < kHN; ++q) sA[q] = (uint32_t)((int)((threadIdx.x * 37 + q * 101) % 4001) - 2000 + (C::kBias ? kMagicI2F : 0));
      tmem_st32(tS, &sA[0]); tmem_st32(tS + 32, &sA[32]); tmem_st32(tS + 64, &sA[0]); tmem_st32(tS + 96, &sA[32]);
      tmem_wait_st();
    }
    float mtA = 0.f, mtB = 0.f;
    bool raiseA = true, raiseB = true;  // the first step always installs its row max
    StepConsts kA = load_consts(0), kB = kA;
    int i = 0;
    if (n_half >= 2) {
      // steps 0 .. n_half-2 are full (unmasked) by construction; only the last one can be ragged
      fetch(0, sA);
      tmem_wait_ld();
      tc_fence_before();
      mbar_arrive(&bars->s0_read[t]);  // buffer 0 is free for S(2)
      if (tracer && warp == 0) phase[2] = clock64();
      mtA = row_max(sA, kA, false, kHN);
      bool in_a = true;
      while (i + 2 < n_half) {
        pipe_step(Yes{}, i, sA, mtA, raiseA, kA, pA, pB, sB, mtB, raiseB, kB);
        ++i;
        if (!(i + 2 < n_half)) { in_a = false; break; }
        pipe_step(Yes{}, i, sB, mtB, raiseB, kB, pB, pA, sA, mtA, raiseA, kA);
        ++i;
      }
      // step n_half-2: nothing to prefetch (the last step fetches for itself)
      if (in_a) {
        pipe_step(No{}, i, sA, mtA, raiseA, kA, pA, pB, sB, mtB, raiseB, kB);
        ++i;
        last_step(i, sB, pB, pA);
      } else {
        pipe_step(No{}, i, sB, mtB, raiseB, kB, pB, pA, sA, mtA, raiseA, kA);
        ++i;
        last_step(i, sA, pA, pB);
      }
    } else {
      last_step(0, sA, pA, pB);
    }

    if (tracer && warp == 0) phase[3] = clock64();
    // ---------------------------------------------------------------- epilogue: O * sV / l
    // Everything that does not need O is computed before the wait (the divisions alone are ~300 clk).
    auto of = [](uint32_t x) { return kPv8 ? (float)(int)x : __uint_as_float(x); };   // O: int32 with INT8 P.V
    float l, la, lb, lc, ld;
    unpack2(lsum[0], la, lb);
    unpack2(lsum[1], lc, ld);
    l = kBlk ? l_acc : (la + lb) + (lc + ld);
    const float inv = (l > 0.f) ? out_scale / l : 0.f;  // fa_tc_int8_b.cu:549-553 guard
    if constexpr (kPersist) {   // (unit, q_base) are recomputed here instead of being kept live across the main loop
      int kk = k;
      asm volatile("" : "+r"(kk));
      if constexpr (kBlk) {     // the next item's constants go into the other table (its last readers finished an item ago)
        const int itn = item0 + (kk + 1) * item_stride;
        if (itn < n_items)
          fetch_item_consts(itn / nqb, (itn % nqb) * (2 * kBM), blk_tab0 + ((kk & 1) ? 0 : prm.n_pad / 32), pre_c, pre_os);
      }
      set_item(item0 + kk * item_stride);
    }
    const int row = q_base + t * kBM + row_in_tile;
    const int b = unit / prm.H, head = unit % prm.H;
    const int osz = prm.out_dtype == 0 ? 4 : 2;   // output element size
    auto pack16 = [&](float lo, float hi) { return prm.out_dtype == 2 ? pack_bf16x2(lo, hi) : pack_f16x2(lo, hi); };
    // O may be a slab of a larger tensor: row / batch strides come from the caller (dense: H*d, N*H*d)
    const size_t out_off = ((size_t)b * (size_t)prm.o_bs + (size_t)row * (size_t)prm.o_ld + (size_t)head * prm.d) * osz;
    // destination 0 is O itself, 1 .. n_peers the replicas (same strides, same bytes)
    auto dest = [&](int pe) { return reinterpret_cast<char*>(pe == 0 ? prm.O : prm.peer_O[pe - 1]); };
    const bool row_ok = row < prm.N;
    const bool vec_ok = (prm.d & 3) == 0;
    constexpr bool kStaged = C::kTileBytesQK >= 16384;
    constexpr bool kTmaStore = C::kStagesK * C::kTileBytesQK >= 65536;
    // (a parity wait on pv_done could alias here: the barrier may be two phases behind)
    mbar_wait(&bars->o_final[t], (uint32_t)k & 1, err_flag, 321 + t, dead);
    dead = __any_sync(0xffffffffu, dead);
    tc_fence_after();
    if (tracer && warp == 0) phase[4] = clock64();
    // Each thread owns one output row, so direct stores touch 32 different rows per instruction
    // (16 B each): measured 9.5 k clk per CTA, 5 % of its life.  Best case: the whole K ring is free
    // (every Q·K^T of both tiles has retired), it is at least 64 KB and d is a multiple of 32: each warp
    // stages 32 rows x 32 columns per 128B-swizzled 4 KB tile of the ring (two tiles per warp) and one
    // lane hands them to the TMA as tensor stores; rows beyond N are clipped by the tensor map
    // ([B][N][H*d], box 1 x 32 x 32).  All TMEM loads are issued up front, one proxy fence per two tiles.
    if (kTmaStore && prm.tma_store) {
      constexpr bool kStageV = kPersist && !(QMHA_PERSIST_SINGLE & 8);   // stage through the V^T ring only
      if constexpr (!kStageV) mbar_wait(&bars->qk_done, (uint32_t)k & 1, err_flag, 341 + t, dead);
      // The TMA needs ~1.2 k clk to read a staged tile, so re-using a staging tile costs that much.
      // When the K and V rings together give every warp 16 KB, all (up to four) tiles of a warp are
      // staged at once; the V ring is only free once the other query tile has finished as well.
      // Persistent kernel: only the V^T ring is used (the K ring already receives the next item's tiles).
      constexpr int kRingBytes = C::kStagesK * C::kTileBytesQK + C::kStagesV * C::kTileBytesV;
      constexpr int kChunks = kD / 32;
      constexpr int kBufsV = (C::kStagesV * C::kTileBytesV) / (8 * 4096);
      static_assert(!kPersist || kBufsV >= 2, "persistent kernel: the V^T ring must give every softmax warp two staging tiles");
      constexpr int kBufs = kStageV ? (kBufsV < kChunks ? kBufsV : kChunks)
                                    : ((kRingBytes >= 8 * 4 * 4096 && kChunks > 2) ? 4 : 2);
      if constexpr (kBufs == 4 || kStageV) {
        mbar_wait(&bars->o_final[t ^ 1], (uint32_t)k & 1, err_flag, 343 + t, dead);
        tc_fence_after();
      }
      float* stage = reinterpret_cast<float*>(kStageV ? sV : sK) + warp * (kBufs * 1024);   // kBufs x (32 rows x 32 floats)
      const int row0 = q_base + t * kBM + (warp & 3) * 32;
      // Rolled on purpose: this code runs once per CTA, straight out of a cold instruction cache, and
      // its fetch — not its execution — is what it costs.
#pragma unroll 1
      for (int ch = 0; ch < kChunks; ++ch) {
        if (ch >= kBufs) {
          if (lane == 0) bulk_wait_group_read<kBufs - 1>();  // the store that used this tile has read it
          __syncwarp();
        }
        float* buf = stage + (ch % kBufs) * 1024;
        uint32_t o[32];
        tmem_ld32(tO + ch * 32, o);
        tmem_wait_ld();
        if (prm.out_dtype == 0) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 v = make_float4(of(o[4 * j]) * inv, of(o[4 * j + 1]) * inv,
                                         of(o[4 * j + 2]) * inv, of(o[4 * j + 3]) * inv);
            *reinterpret_cast<float4*>(buf + lane * 32 + ((j ^ (lane & 7)) << 2)) = v;
          }
        } else {   // 16-bit output: 32 rows x 64 B, SWIZZLE_64B (16-byte chunk index ^= (row >> 1) & 3)
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint4 v;
            v.x = pack16(of(o[8 * j]) * inv, of(o[8 * j + 1]) * inv);
            v.y = pack16(of(o[8 * j + 2]) * inv, of(o[8 * j + 3]) * inv);
            v.z = pack16(of(o[8 * j + 4]) * inv, of(o[8 * j + 5]) * inv);
            v.w = pack16(of(o[8 * j + 6]) * inv, of(o[8 * j + 7]) * inv);
            *reinterpret_cast<uint4*>(reinterpret_cast<char*>(buf) + lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4)) = v;
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0 && row0 < prm.N && ch * 32 < prm.d) {
          tma_store_3d(&tm_o, buf, head * prm.d + ch * 32, row0, b);
#pragma unroll 1
          for (int pe = 0; pe < prm.n_peers; ++pe)   // replicas: the same staged tile, once per peer (NVLink)
            tma_store_3d(&tm_peers.m[pe], buf, head * prm.d + ch * 32, row0, b);
          bulk_commit_group();
        }
      }
      if (tracer && warp == 0) phase[11] = clock64();
      if (lane == 0) {
        bulk_wait_group_read<0>();
        if constexpr (kPersist) mbar_arrive(&bars->epi_done);   // this warp's staging tiles may be overwritten (V^T producer)
      }
      __syncwarp();
    } else if (kStaged && vec_ok && prm.out_dtype == 0) {
      // Same staging through the (dead) Q tile of this warpgroup, written out with ordinary 16-byte
      // stores: 8 lanes cover one 128-byte row segment, 4 rows per instruction.
      float* stage = reinterpret_cast<float*>(sQ + t * C::kTileBytesQK) + (warp & 3) * 1024;  // 32 x 32 floats
      const int r_sub = lane >> 3, c4 = lane & 7;
      const int row0 = q_base + t * kBM + (warp & 3) * 32;          // first row of this warp
      const size_t off0 = (size_t)b * (size_t)prm.o_bs + (size_t)row0 * (size_t)prm.o_ld + (size_t)head * prm.d;
      constexpr int kChunks = kD / 32;
      uint32_t o[kChunks][32];
#pragma unroll
      for (int ch = 0; ch < kChunks; ++ch) tmem_ld32(tO + ch * 32, o[ch]);
      tmem_wait_ld();
#pragma unroll
      for (int ch = 0; ch < kChunks; ++ch) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 v = make_float4(of(o[ch][4 * j]) * inv, of(o[ch][4 * j + 1]) * inv,
                                       of(o[ch][4 * j + 2]) * inv, of(o[ch][4 * j + 3]) * inv);
          *reinterpret_cast<float4*>(stage + lane * 32 + ((j ^ (lane & 7)) << 2)) = v;
        }
        __syncwarp();
        const int col = ch * 32 + c4 * 4;
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int r = it * 4 + r_sub;
          const float4 v = *reinterpret_cast<const float4*>(stage + r * 32 + ((c4 ^ (r & 7)) << 2));
          if (row0 + r < prm.N && col < prm.d) {
#pragma unroll 1
            for (int pe = 0; pe <= prm.n_peers; ++pe)
              *reinterpret_cast<float4*>(reinterpret_cast<float*>(dest(pe)) + off0 + (size_t)r * (size_t)prm.o_ld + col) = v;
          }
        }
        __syncwarp();
      }
    } else {
#pragma unroll
      for (int ch = 0; ch < kD / 32; ++ch) {
        uint32_t o[32];
        tmem_ld32(tO + ch * 32, o);
        tmem_wait_ld();
#pragma unroll 1
       for (int pe = 0; pe <= prm.n_peers; ++pe) {
        char* out_b = dest(pe) + out_off;
        float* out = reinterpret_cast<float*>(out_b);
        if (row_ok && prm.out_dtype != 0) {   // 16-bit output without TMA: pairs when d is even, scalars otherwise
          uint16_t* o16 = reinterpret_cast<uint16_t*>(out_b);
          if ((prm.d & 1) == 0) {
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              const int col = ch * 32 + i;
              if (col < prm.d)
                *reinterpret_cast<uint32_t*>(o16 + col) = pack16(of(o[i]) * inv, of(o[i + 1]) * inv);
            }
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const int col = ch * 32 + i;
              if (col < prm.d) o16[col] = (uint16_t)(pack16(of(o[i]) * inv, 0.f) & 0xFFFFu);
            }
          }
        } else if (row_ok) {
          if (vec_ok) {
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              const int col = ch * 32 + i;
              if (col < prm.d) {
                float4 v = make_float4(of(o[i]) * inv, of(o[i + 1]) * inv,
                                       of(o[i + 2]) * inv, of(o[i + 3]) * inv);
                *reinterpret_cast<float4*>(out + col) = v;
              }
            }
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const int col = ch * 32 + i;
              if (col < prm.d) out[col] = of(o[i]) * inv;
            }
          }
        }
       }
      }
    }
    }   // items
  }

  // ---------------------------------------------------------------------------- teardown
  if (traced_cta && threadIdx.x == 0) phase[5] = clock64();
  tc_fence_before();
  __syncthreads();
  if (traced_cta && threadIdx.x == 0) phase[6] = clock64();
  if (warp == kAllocWarp) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
  if (prm.cycles != nullptr && threadIdx.x == 0) {   // development aid: SM clocks this CTA was resident
    const long long t_exit = clock64();
    atomicAdd(prm.cycles, (unsigned long long)(t_exit - t_entry));
    atomicAdd(prm.cycles + 1, kPersist ? (unsigned long long)((n_items - item0 + item_stride - 1) / item_stride) : 1ull);   // work items
    // per SM: {~(earliest CTA start), latest CTA end} -> span of the launch on that SM against the sum of residencies
    uint32_t smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    if (smid < (uint32_t)kCycleSms) {
      atomicMax(prm.cycles + 2 + 2 * smid, ~(unsigned long long)t_entry);
      atomicMax(prm.cycles + 3 + 2 * smid, (unsigned long long)t_exit);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) ==
            cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// 2D row-major tensor [rows, cols] of `elt`-byte elements, box = [box_rows, box_cols].
bool make_map_2d(CUtensorMap* m, const void* base, int elt, uint64_t rows, uint64_t cols,
                 uint32_t box_rows, uint32_t box_cols, std::string* err) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) { *err = "cuTensorMapEncodeTiled entry point not available"; return false; }
  const cuuint64_t gdim[2] = {cols, rows};
  const cuuint64_t gstride[1] = {cols * (uint64_t)elt};
  const cuuint32_t box[2] = {box_cols, box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const uint32_t span = box_cols * elt;
  const CUtensorMapSwizzle sw = span == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                              : span == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                           : CU_TENSOR_MAP_SWIZZLE_32B;
  CUresult r = fn(m, elt == 1 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2,
                  const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    *err = "cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r);
    return false;
  }
  return true;
}

// 3D output tensor [d2][d1][d0] (d0 contiguous), box = [1][box1][box0]; out_dtype 0 = fp32 (SWIZZLE_128B,
// box0 * 4 == 128), 1 = fp16 / 2 = bf16 (SWIZZLE_64B, box0 * 2 == 64).
bool make_map_3d_out(CUtensorMap* m, const void* base, int out_dtype, uint64_t d0, uint64_t d1, uint64_t d2,
                     uint64_t ld1, uint64_t ld2, uint32_t box0, uint32_t box1, std::string* err) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) { *err = "cuTensorMapEncodeTiled entry point not available"; return false; }
  const uint64_t elt = out_dtype == 0 ? 4 : 2;
  const cuuint64_t gdim[3] = {d0, d1, d2};
  const cuuint64_t gstride[2] = {ld1 * elt, ld2 * elt};   // row / batch pitch (dense: d0, d0 * d1 elements)
  const cuuint32_t box[3] = {box0, box1, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  const CUtensorMapDataType dt = out_dtype == 0 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                               : out_dtype == 1 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  CUresult r = fn(m, dt, 3, const_cast<void*>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, out_dtype == 0 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                  CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    *err = "cuTensorMapEncodeTiled(output) failed with CUresult " + std::to_string((int)r);
    return false;
  }
  return true;
}

// Can this launch run on the persistent kernel?  It needs the TMA-store epilogue (staging in the V^T ring), at least four
// half-steps per item and room for the second block-scale table.
template <bool kInt8, int kD, bool kPv8>
bool persistent_ok(const AttnLaunch& a) {
  using C = Cfg<kInt8, kD, kPv8>;
  const char* e = getenv("QMHA_PERSIST");   // read per launch (tests switch it inside one process)
  const int mode = e && *e ? atoi(e) : QMHA_DEFAULT_PERSIST;
  if (!mode || a.trace) return false;
  const uint64_t o_ld = a.o_ld > 0 ? (uint64_t)a.o_ld : (uint64_t)a.H * a.d, o_elt = a.out_dtype == 0 ? 4 : 2;
  const uint64_t o_bs = a.o_bs > 0 ? (uint64_t)a.o_bs : (uint64_t)a.N * o_ld;
  if ((a.d % 32) != 0 || (o_ld * o_elt) % 16 != 0 || (o_bs * o_elt) % 16 != 0 || getenv("QMHA_NO_TMA_STORE") != nullptr) return false;
  if ((a.N + kHN - 1) / kHN < 4 || (a.N % 256) != 0) return false;   // every item must start on score buffer 0, phase 0
  const size_t smem = (size_t)C::kSmemBytes + (a.blk_scales ? (size_t)(a.n_pad / 32) * 32 : 0);
  return smem <= 227 * 1024;
}

template <bool kInt8, int kD, int kPolyEvery, bool kBlk, bool kTrace, int kFa = 6, int kFb = 25, bool kBf16 = false,
          bool kPv8 = false, bool kPersist = false>
bool launch_cfg(const AttnLaunch& a, std::string* err) {
  using C = Cfg<kInt8, kD, kPv8>;
  const uint64_t units = (uint64_t)a.B * a.H;
  CUtensorMap tq, tk, tv;
  if (!make_map_2d(&tq, a.Qp, C::kEltQK, units * a.n_pad, kD, kBM, C::kAtomQK / C::kEltQK, err) ||
      !make_map_2d(&tk, a.Kp, C::kEltQK, units * a.n_pad, kD, kBN, C::kAtomQK / C::kEltQK, err) ||
      !make_map_2d(&tv, a.Vt, kPv8 ? 1 : 2, units * kD, a.n_pad, kD, kPv8 ? 128 : 64, err))
    return false;
  // output tensor map for the TMA-store epilogue (only when a 32-column chunk never straddles a head)
  CUtensorMap to;
  memset(&to, 0, sizeof(to));
  PeerMaps peers;
  memset(&peers, 0, sizeof(peers));
  const uint64_t o_ld = a.o_ld > 0 ? (uint64_t)a.o_ld : (uint64_t)a.H * a.d;
  const uint64_t o_bs = a.o_bs > 0 ? (uint64_t)a.o_bs : (uint64_t)a.N * o_ld;
  const uint64_t o_elt = a.out_dtype == 0 ? 4 : 2;
  if (a.n_peers < 0 || a.n_peers > kMaxPeers) { *err = "n_peers out of range"; return false; }
  // tensor maps need 16-byte multiples for the pitches (and the plain paths store 16 bytes at a time)
  const bool pitch_ok = (o_ld * o_elt) % 16 == 0 && (o_bs * o_elt) % 16 == 0;
  if ((a.o_ld > 0 || a.o_bs > 0) && !pitch_ok) { *err = "output row / batch strides must be multiples of 16 bytes"; return false; }
  const bool tma_store = (a.d % 32) == 0 && pitch_ok && getenv("QMHA_NO_TMA_STORE") == nullptr;
  if (tma_store) {
    if (!make_map_3d_out(&to, a.O, a.out_dtype, (uint64_t)a.H * a.d, (uint64_t)a.N, (uint64_t)a.B, o_ld, o_bs, 32, 32, err))
      return false;
    for (int pe = 0; pe < a.n_peers; ++pe)
      if (!make_map_3d_out(&peers.m[pe], a.peer_O[pe], a.out_dtype, (uint64_t)a.H * a.d, (uint64_t)a.N, (uint64_t)a.B,
                           o_ld, o_bs, 32, 32, err))
        return false;
  }
  auto kern = attn_fwd_kernel<kInt8, kD, kPolyEvery, kBlk, kTrace, kFa, kFb, kBf16, kPv8, kPersist>;
  // block mode keeps one float4 of constants per 32-key block of the unit behind the barriers (persistent: two tables)
  const size_t smem_bytes = (size_t)C::kSmemBytes + (kBlk ? (size_t)(a.n_pad / 32) * 16 * (kPersist ? 2 : 1) : 0);
  if (smem_bytes > 227 * 1024) {
    *err = "sequence too long for the per-block scale table in shared memory (use head granularity)";
    return false;
  }
  {  // per device (context) attribute; cheap enough to set on every launch
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem_bytes);
    if (e != cudaSuccess) { *err = std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(e); return false; }
  }
  AttnParams p;
  p.O = a.O;
  p.scales = a.scales;
  p.error_flag = a.error_flag;
  p.error_host = a.error_host;
  p.launch_id = a.launch_id;
  p.out_dtype = a.out_dtype;
  p.trace = a.trace;
  p.blk_scales = a.blk_scales;
  p.blk_aux = a.blk_aux;
  p.blk_vmax = a.blk_vmax;
  p.B = a.B; p.N = a.N; p.H = a.H; p.d = a.d;
  p.n_pad = a.n_pad;
  p.units = (int)units;
  p.n_kv_tiles = (a.N + kBN - 1) / kBN;
  p.n_half_steps = (a.N + kHN - 1) / kHN;
  p.scale_log2 = 1.4426950408889634f / sqrtf((float)a.d);
  p.debug_no_mma = getenv("QMHA_DEBUG_NO_MMA") != nullptr;
  p.tma_store = tma_store ? 1 : 0;
  p.one = 1;
  p.cycles = a.cycles;
  p.o_ld = (long long)o_ld;
  p.o_bs = (long long)o_bs;
  p.n_peers = a.n_peers;
  for (int pe = 0; pe < kMaxPeers; ++pe) p.peer_O[pe] = pe < a.n_peers ? a.peer_O[pe] : nullptr;
  p.stagger_ns = 0u;
  p.stagger_ctas = 0u;
  if (a.n_peers > 0) {
    // drain time of half a wave of CTAs at ~750 GB/s of NVLink egress (one CTA sends 256 rows x d x n_peers elements);
    // never more than half of a CTA's own running time (~1400 clk per 64-key half-step at ~1.9 GHz)
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const double bytes_cta = 2.0 * kBM * a.d * (double)o_elt * a.n_peers;
    double ns = bytes_cta * (sms / 2) / 750.0;
    const double cta_ns = (double)((a.N + kHN - 1) / kHN) * 1400.0 / 1.9;
    if (ns > 0.9 * cta_ns) ns = 0.9 * cta_ns;
    // only worth its delay when the launch runs for several waves (a short launch would just finish later)
    const long long ctas = (long long)((a.N + 2 * kBM - 1) / (2 * kBM)) * (long long)units;
    if (ctas < 3LL * sms) ns = 0.0;
    const char* e = getenv("QMHA_PEER_STAGGER_NS");      // tuning aid: 0 switches the stagger off
    if (e && *e) ns = atof(e);
    p.stagger_ns = (unsigned)ns;
    p.stagger_ctas = (unsigned)sms;
  }
  dim3 grid((a.N + 2 * kBM - 1) / (2 * kBM), (unsigned)units, 1);
  p.n_qblocks = (int)grid.x;
  p.n_items = (int)(grid.x * units);
  if constexpr (kPersist) {
    static int n_sm[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && n_sm[dev] == 0) cudaDeviceGetAttribute(&n_sm[dev], cudaDevAttrMultiProcessorCount, dev);
    const char* ge = getenv("QMHA_PERSIST_GRID");   // tuning aid: CTAs of the persistent grid (default: one per SM)
    const int grid_env = ge && *ge ? atoi(ge) : 0;
    const int sms = grid_env > 0 ? grid_env : ((dev >= 0 && dev < 64 && n_sm[dev] > 0) ? n_sm[dev] : 148);
    grid = dim3((unsigned)(p.n_items < sms ? p.n_items : sms), 1, 1);
  }
  kern<<<grid, kThreads, smem_bytes, a.stream>>>(tq, tk, tv, to, peers, p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { *err = std::string("attention launch: ") + cudaGetErrorString(e); return false; }
  return true;
}

}  // namespace

int attention_max_block_keys(int d_pad) {
  const int tiles = d_pad == 32 ? Cfg<true, 32>::kSmemBytes : (d_pad == 64 ? Cfg<true, 64>::kSmemBytes : Cfg<true, 128>::kSmemBytes);
  return (227 * 1024 - tiles) / 16 * 32;
}

// split points of the softmax step (in exp2 pairs): P store after QMHA_FA pairs, row max of the next scores after QMHA_FB
#ifndef QMHA_FA
#define QMHA_FA 6
#endif
#ifndef QMHA_FB
#define QMHA_FB 25
#endif

bool launch_attention(const AttnLaunch& a, std::string* err) {
  if (a.units_y_limit_exceeded()) { *err = "B*h exceeds the CUDA grid.y limit (65535)"; return false; }
  // a.variant = k: exp2 of every k-th score pair goes to the FMA-pipe polynomial (0 = all MUFU)
  const int poly = a.variant;
  const bool blk = a.blk_scales != nullptr;
  if (blk && !a.int8) { *err = "block scales require the INT8 variant"; return false; }
  if (a.int8 && a.bf16) { *err = "bf16 selects the 16-bit kernel"; return false; }
  if (a.pv8 && !a.int8) { *err = "INT8 P.V is a mode of the INT8 kernel"; return false; }
#ifndef QMHA_ONLY_D128
  if (a.trace) {
    if (!(a.int8 && a.d_pad == 128)) { *err = "tracing is only built for the INT8 d=128 kernel"; return false; }
    return blk ? launch_cfg<true, 128, 0, true, true>(a, err) : launch_cfg<true, 128, 0, false, true>(a, err);
  }
#endif
#if defined(QMHA_ONLY_D128)   // quick experiment builds (tools/build_variant.sh): d = 128, all-MUFU exponentials only
#define QMHA_DISPATCH(INT8, BLK, D, BF)                                   \
  if (D == 128 && poly == 0) return launch_cfg<INT8, 128, 0, BLK, false, QMHA_FA, QMHA_FB, BF>(a, err);
#define QMHA_DISPATCH_PV8(BLK, D)                                         \
  if (D == 128 && poly == 0) return launch_cfg<true, 128, 0, BLK, false, QMHA_FA, QMHA_FB, false, true>(a, err);
#elif defined(QMHA_BUILD_POLY)  // experiment builds with the FMA-pipe exp2 share (measured slower)
#define QMHA_DISPATCH(INT8, BLK, D, BF)                                   \
  switch (poly) {                                                     \
    case 0: return launch_cfg<INT8, D, 0, BLK, false, QMHA_FA, QMHA_FB, BF>(a, err);        \
    case 4: return launch_cfg<INT8, D, 4, BLK, false, QMHA_FA, QMHA_FB, BF>(a, err);        \
    case 8: return launch_cfg<INT8, D, 8, BLK, false, QMHA_FA, QMHA_FB, BF>(a, err);        \
  }
#else
#define QMHA_DISPATCH(INT8, BLK, D, BF)                                   \
  if (poly == 0) return launch_cfg<INT8, D, 0, BLK, false, QMHA_FA, QMHA_FB, BF>(a, err);
#endif
#ifndef QMHA_DISPATCH_PV8
#define QMHA_DISPATCH_PV8(BLK, D)                                         \
  if (poly == 0) return launch_cfg<true, D, 0, BLK, false, QMHA_FA, QMHA_FB, false, true>(a, err);
#endif
  if (a.int8 && a.pv8 && blk) {
    switch (a.d_pad) {
      case 32: QMHA_DISPATCH_PV8(true, 32) break;
      case 64: QMHA_DISPATCH_PV8(true, 64) break;
      case 128: QMHA_DISPATCH_PV8(true, 128) break;
    }
  } else if (a.int8 && a.pv8) {
    switch (a.d_pad) {
      case 32: QMHA_DISPATCH_PV8(false, 32) break;
      case 64: QMHA_DISPATCH_PV8(false, 64) break;
      case 128: QMHA_DISPATCH_PV8(false, 128) break;
    }
  } else if (a.int8 && blk) {
    if (a.d_pad == 128 && poly == 0 && persistent_ok<true, 128, false>(a))
      return launch_cfg<true, 128, 0, true, false, QMHA_FA, QMHA_FB, false, false, true>(a, err);
    switch (a.d_pad) {
      case 32: QMHA_DISPATCH(true, true, 32, false) break;
      case 64: QMHA_DISPATCH(true, true, 64, false) break;
      case 128: QMHA_DISPATCH(true, true, 128, false) break;
    }
  } else if (a.int8) {
    if (a.d_pad == 128 && poly == 0 && persistent_ok<true, 128, false>(a))
      return launch_cfg<true, 128, 0, false, false, QMHA_FA, QMHA_FB, false, false, true>(a, err);
    switch (a.d_pad) {
      case 32: QMHA_DISPATCH(true, false, 32, false) break;
      case 64: QMHA_DISPATCH(true, false, 64, false) break;
      case 128: QMHA_DISPATCH(true, false, 128, false) break;
    }
  } else if (a.bf16) {
    switch (a.d_pad) {
      case 32: QMHA_DISPATCH(false, false, 32, true) break;
      case 64: QMHA_DISPATCH(false, false, 64, true) break;
      case 128: QMHA_DISPATCH(false, false, 128, true) break;
    }
  } else {
    switch (a.d_pad) {
      case 32: QMHA_DISPATCH(false, false, 32, false) break;
      case 64: QMHA_DISPATCH(false, false, 64, false) break;
      case 128: QMHA_DISPATCH(false, false, 128, false) break;
    }
  }
#undef QMHA_DISPATCH
#undef QMHA_DISPATCH_PV8
  *err = "unsupported padded head dimension or kernel variant";
  return false;
}

}  // namespace qmha
