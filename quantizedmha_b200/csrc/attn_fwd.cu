// attn_fwd.cu — warp-specialised FlashAttention-2 forward for sm_100a (B200).
//
// Replaces the reference's fa_kernel family: mha_kernels/fa_tc_int8_b.cu:408-579 (INT8, WMMA
// IMMA) and fa_tc_v1a.cu:222-413 / fa_tc_v2a.cu:274-496 (FP16, WMMA HMMA).  Same math
// (softmax(Q·K^T/sqrt(d))·V per head, FP32 statistics and output) — different machine:
//
//   * one CTA owns TWO 128-row query tiles of one (batch, head) unit and walks the KV sequence
//     in 128-key tiles;
//   * warps 8-11 form a service warpgroup (register budget handed to the softmax warps with
//     setmaxnreg); warp 9 is the TMA producer: Q once, then K and V^T tiles through two mbarrier rings;
//   * warp 8 (one elected lane) issues every tcgen05.mma:  S_t = Q_t·K_j^T  (kind::i8, int32
//     accumulators in TMEM — or kind::f16 for the FP16 variant) and  O_t += P_t·V_j
//     (kind::f16, A operand = P read straight from TMEM, B = V^T tile in shared memory);
//   * warps 0-3 / 4-7 are the softmax warpgroups of tile 0 / tile 1: one thread per query row,
//     S read with tcgen05.ld, dequant scale folded into the exponent FMA, exp2 on packed fp16
//     (INT8 variant) or fp32 (FP16 variant), P written back over S with tcgen05.st, running
//     row sum in FP32, lazy rescale of O in TMEM only when the row max grows by more than 2^4;
//   * epilogue: O·(sV/l) from TMEM to global memory in the reference's [N, h·d] layout.
//
// TMEM plan (512 columns): S0|P0 = [0,128)  S1|P1 = [128,256)  O0 = [256,384)  O1 = [384,512).
#include "attn_fwd.cuh"
#include "sm100_ptx.cuh"

namespace qmha {

using namespace ptx;

namespace {

constexpr int kBM = 128;        // query rows per tile == UMMA M
constexpr int kBN = 128;        // keys per KV tile   == UMMA N of Q·K^T
constexpr int kMmaWarp = 8;
constexpr int kTmaWarp = 9;
constexpr int kThreads = 384;            // 2 softmax warpgroups + 1 service warpgroup (MMA, TMA, 2 idle)
constexpr int kRegsSoftmax = 208;        // setmaxnreg budgets: 2*128*208 + 128*72 = 62464 <= 65536
constexpr int kRegsService = 72;
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kColS0 = 0, kColS1 = 128, kColO0 = 256, kColO1 = 384;
constexpr float kRescaleThreshold = 4.0f;  // log2 units: P <= 2^4, well inside fp16
constexpr int kMagicI2F = 0x4B400000;      // float(1.5 * 2^23): int -> float by bit tricks
constexpr float kMagicF = 12582912.0f;

template <bool kInt8, int kD>
struct Cfg {
  static constexpr int kEltQK = kInt8 ? 1 : 2;
  static constexpr int kRowBytesQK = kD * kEltQK;                       // bytes per Q/K row
  static constexpr int kAtomQK = kRowBytesQK < 128 ? kRowBytesQK : 128;  // swizzle span
  static constexpr int kSubQK = kRowBytesQK / kAtomQK;                  // 128B sub-tiles per row
  static constexpr int kTileBytesQK = kBM * kRowBytesQK;
  static constexpr int kSubBytesQK = kBM * kAtomQK;
  static constexpr int kStepsQK = kRowBytesQK / 32;                     // UMMA K steps (32 B each)
  static constexpr int kTileBytesV = kD * kBN * 2;                      // V^T tile: kD x 128 fp16
  static constexpr int kSubBytesV = kD * 128;                           // 64 keys x kD rows
  static constexpr int kStepsPV = kBN / 16;
  static constexpr int kBudget = 200 * 1024 - 2 * kTileBytesQK;
  static constexpr int kStagesRaw = kBudget / (kTileBytesQK + kTileBytesV);
  static constexpr int kStages = kStagesRaw > 4 ? 4 : kStagesRaw;       // K and V ring depth
  static_assert(kStages >= 2, "need at least double buffering");
  static constexpr int kSmemTiles = 2 * kTileBytesQK + kStages * (kTileBytesQK + kTileBytesV);
  static constexpr int kSmemBytes = kSmemTiles + 1024 /*align slack*/ + 256 /*barriers*/;
  static constexpr uint32_t kIdescQK =
      kInt8 ? make_idesc(kAccS32, kFmtS8, kFmtS8, kBM, kBN)
            : make_idesc(kAccF32, kFmtF16, kFmtF16, kBM, kBN);
  static constexpr uint32_t kIdescPV = make_idesc(kAccF32, kFmtF16, kFmtF16, kBM, kD);
};

struct Barriers {
  uint64_t q_full;
  uint64_t k_full[4], k_empty[4];
  uint64_t v_full[4], v_empty[4];
  uint64_t s_full[2], p_full[2], pv_done[2];
  uint32_t tmem_base;
  uint32_t pad;
};
static_assert(sizeof(Barriers) <= 256, "barrier block too large");

// Bounded mbarrier wait.  A healthy wait is microseconds; after ~1e9 cycles (or as soon as any
// CTA has raised the global error flag) the wait gives up, records the site and lets the CTA
// drain so the kernel always terminates and the host can report the failure.
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity, int* err_flag, int site,
                                          bool& dead) {
  if (dead) return false;
  if (mbar_try_wait(bar, parity)) return true;
  const long long t0 = clock64();
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 0x3FFu) == 0) {
      if (clock64() - t0 > 1000000000LL || *((volatile int*)err_flag) != 0) {
        atomicCAS(err_flag, 0, site);
        dead = true;
        return false;
      }
    }
  }
  return true;
}

// ------------------------------------------------------------------------------------------------
// One softmax step for one thread (= one query row) over a 128-key tile.
//   s[128]  : raw S row from TMEM (int32 for INT8, fp32 bits for FP16)
//   c       : logits-to-log2 factor (sQ*sK*log2e/sqrt(d) or log2e/sqrt(d))
//   m_used  : the (lazily updated) reference max in log2 units
// Produces p[64] (packed fp16x2 P row) and adds the row sum into l0/l1.
template <bool kInt8, bool kMasked>
__device__ __forceinline__ float tile_row_max(uint32_t (&s)[128], float c, int n_valid) {
  if constexpr (kInt8) {
    if constexpr (kMasked) {
#pragma unroll
      for (int i = 0; i < 128; ++i)
        if (i >= n_valid) s[i] = (uint32_t)(-(1 << 22));
    }
    int m0 = max((int)s[0], (int)s[1]), m1 = max((int)s[2], (int)s[3]);
    int m2 = max((int)s[4], (int)s[5]), m3 = max((int)s[6], (int)s[7]);
#pragma unroll
    for (int i = 8; i < 128; i += 8) {
      m0 = max(max(m0, (int)s[i + 0]), (int)s[i + 1]);
      m1 = max(max(m1, (int)s[i + 2]), (int)s[i + 3]);
      m2 = max(max(m2, (int)s[i + 4]), (int)s[i + 5]);
      m3 = max(max(m3, (int)s[i + 6]), (int)s[i + 7]);
    }
    return (float)max(max(m0, m1), max(m2, m3)) * c;
  } else {
    if constexpr (kMasked) {
#pragma unroll
      for (int i = 0; i < 128; ++i)
        if (i >= n_valid) s[i] = 0xFF800000u;  // -inf
    }
    float m0 = fmaxf(__uint_as_float(s[0]), __uint_as_float(s[1]));
    float m1 = fmaxf(__uint_as_float(s[2]), __uint_as_float(s[3]));
    float m2 = fmaxf(__uint_as_float(s[4]), __uint_as_float(s[5]));
    float m3 = fmaxf(__uint_as_float(s[6]), __uint_as_float(s[7]));
#pragma unroll
    for (int i = 8; i < 128; i += 8) {
      m0 = fmaxf(fmaxf(m0, __uint_as_float(s[i + 0])), __uint_as_float(s[i + 1]));
      m1 = fmaxf(fmaxf(m1, __uint_as_float(s[i + 2])), __uint_as_float(s[i + 3]));
      m2 = fmaxf(fmaxf(m2, __uint_as_float(s[i + 4])), __uint_as_float(s[i + 5]));
      m3 = fmaxf(fmaxf(m3, __uint_as_float(s[i + 6])), __uint_as_float(s[i + 7]));
    }
    return fmaxf(fmaxf(m0, m1), fmaxf(m2, m3)) * c;
  }
}

template <bool kInt8, bool kFastExp, bool kMasked>
__device__ __forceinline__ void tile_row_exp(const uint32_t (&s)[128], uint32_t (&p)[64], float c,
                                             float m_used, int n_valid, float& l0, float& l1) {
  // x = s*c - m_used.  INT8: s is an int32 with |s| < 2^22, so bits(s + 0x4B400000) is the
  // float 12582912 + s exactly and one FMA does int->float, scale and max subtraction.
  const float bias = kInt8 ? -fmaf(kMagicF, c, m_used) : -m_used;
#pragma unroll
  for (int i = 0; i < 64; ++i) {
    float x0, x1;
    if constexpr (kInt8) {
      x0 = fmaf(__int_as_float((int)s[2 * i] + kMagicI2F), c, bias);
      x1 = fmaf(__int_as_float((int)s[2 * i + 1] + kMagicI2F), c, bias);
    } else {
      x0 = fmaf(__uint_as_float(s[2 * i]), c, bias);
      x1 = fmaf(__uint_as_float(s[2 * i + 1]), c, bias);
    }
    if constexpr (kFastExp) {
      uint32_t e = ex2_f16x2(pack_f16x2(x0, x1));
      if constexpr (kMasked) {
        if (2 * i >= n_valid) e &= 0xFFFF0000u;
        if (2 * i + 1 >= n_valid) e &= 0x0000FFFFu;
      }
      add_f16x2_to_f32(l0, l1, e);
      p[i] = e;
    } else {
      float e0 = ex2_approx(x0), e1 = ex2_approx(x1);
      if constexpr (kMasked) {
        if (2 * i >= n_valid) e0 = 0.f;
        if (2 * i + 1 >= n_valid) e1 = 0.f;
      }
      l0 += e0;
      l1 += e1;
      p[i] = pack_f16x2(e0, e1);
    }
  }
}

template <bool kInt8, int kD, bool kFastExp>
__global__ void __launch_bounds__(kThreads, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                const __grid_constant__ CUtensorMap tm_v, AttnParams prm) {
  using C = Cfg<kInt8, kD>;
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B operands need 1024-byte aligned tiles.
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + 2 * C::kTileBytesQK;
  uint8_t* sV = sK + C::kStages * C::kTileBytesQK;
  Barriers* bars = reinterpret_cast<Barriers*>(sV + C::kStages * C::kTileBytesV);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int unit = blockIdx.y;                 // b * H + head
  const int q_base = blockIdx.x * (2 * kBM);   // first query row of this CTA
  const int n_tiles = prm.n_kv_tiles;
  int* err_flag = prm.error_flag;
  bool dead = false;

  if (*((volatile int*)err_flag) != 0) return;  // an earlier CTA already failed: drain the grid

  if (warp == kTmaWarp && lane == 0) {
    prefetch_tmap(&tm_q);
    prefetch_tmap(&tm_k);
    prefetch_tmap(&tm_v);
    mbar_init(&bars->q_full, 1);
    for (int i = 0; i < C::kStages; ++i) {
      mbar_init(&bars->k_full[i], 1);
      mbar_init(&bars->k_empty[i], 1);
      mbar_init(&bars->v_full[i], 1);
      mbar_init(&bars->v_empty[i], 1);
    }
    for (int t = 0; t < 2; ++t) {
      mbar_init(&bars->s_full[t], 1);
      mbar_init(&bars->p_full[t], 128);
      mbar_init(&bars->pv_done[t], 1);
    }
    fence_mbar_init();
  }
  if (warp == kMmaWarp) {
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
   if (warp == kTmaWarp) {
    // ======================================================================== TMA producer
    if (lane == 0) {
      const int q_row = unit * prm.n_pad + q_base;
      mbar_arrive_expect_tx(&bars->q_full, 2 * C::kTileBytesQK);
#pragma unroll
      for (int t = 0; t < 2; ++t)
#pragma unroll
        for (int sub = 0; sub < C::kSubQK; ++sub)
          tma_load_2d(sQ + t * C::kTileBytesQK + sub * C::kSubBytesQK, &tm_q, &bars->q_full,
                      sub * (C::kAtomQK / C::kEltQK), q_row + t * kBM);
      const int k_row0 = unit * prm.n_pad;
      const int v_row = unit * kD;
      for (int j = 0; j < n_tiles; ++j) {
        const int st = j % C::kStages;
        const uint32_t ph = (uint32_t)(j / C::kStages);
        if (j >= C::kStages) mbar_wait(&bars->k_empty[st], (ph - 1) & 1, err_flag, 101, dead);
        mbar_arrive_expect_tx(&bars->k_full[st], C::kTileBytesQK);
#pragma unroll
        for (int sub = 0; sub < C::kSubQK; ++sub)
          tma_load_2d(sK + st * C::kTileBytesQK + sub * C::kSubBytesQK, &tm_k, &bars->k_full[st],
                      sub * (C::kAtomQK / C::kEltQK), k_row0 + j * kBN);
        if (j >= C::kStages) mbar_wait(&bars->v_empty[st], (ph - 1) & 1, err_flag, 102, dead);
        mbar_arrive_expect_tx(&bars->v_full[st], C::kTileBytesV);
#pragma unroll
        for (int sub = 0; sub < 2; ++sub)
          tma_load_2d(sV + st * C::kTileBytesV + sub * C::kSubBytesV, &tm_v, &bars->v_full[st],
                      j * kBN + sub * 64, v_row);
      }
    }
   } else if (warp == kMmaWarp) {
    // ======================================================================== MMA issuer
    if (lane == 0) {
      const uint32_t sQ_a = smem_u32(sQ), sK_a = smem_u32(sK), sV_a = smem_u32(sV);
      auto issue_qk = [&](int t, int st) {
        const uint32_t d_tmem = tmem_base + (t ? kColS1 : kColS0);
#pragma unroll
        for (int ks = 0; ks < C::kStepsQK; ++ks) {
          const uint32_t off = (uint32_t)((ks * 32) / C::kAtomQK) * C::kSubBytesQK +
                               (uint32_t)((ks * 32) % C::kAtomQK);
          const uint64_t a = make_smem_desc(sQ_a + t * C::kTileBytesQK + off, C::kAtomQK);
          const uint64_t b = make_smem_desc(sK_a + st * C::kTileBytesQK + off, C::kAtomQK);
          if constexpr (kInt8) mma_i8_ss(d_tmem, a, b, C::kIdescQK, ks > 0);
          else mma_f16_ss(d_tmem, a, b, C::kIdescQK, ks > 0);
        }
      };
      auto issue_pv = [&](int t, int st, bool accumulate) {
        const uint32_t d_tmem = tmem_base + (t ? kColO1 : kColO0);
        const uint32_t p_tmem = tmem_base + (t ? kColS1 : kColS0);
#pragma unroll
        for (int ks = 0; ks < C::kStepsPV; ++ks) {
          const uint32_t off = (uint32_t)(ks / 4) * C::kSubBytesV + (uint32_t)(ks % 4) * 32;
          const uint64_t b = make_smem_desc(sV_a + st * C::kTileBytesV + off, 128);
          mma_f16_ts(d_tmem, p_tmem + ks * 8, b, C::kIdescPV, (accumulate || ks > 0) ? 1u : 0u);
        }
      };

      mbar_wait(&bars->q_full, 0, err_flag, 201, dead);
      mbar_wait(&bars->k_full[0], 0, err_flag, 202, dead);
      tc_fence_after();
      issue_qk(0, 0);
      mma_commit(&bars->s_full[0]);
      issue_qk(1, 0);
      mma_commit(&bars->s_full[1]);
      mma_commit(&bars->k_empty[0]);

      for (int j = 0; j < n_tiles; ++j) {
        const int st = j % C::kStages;
        const uint32_t ph = (uint32_t)(j / C::kStages) & 1;
        const int jn = j + 1;
        const int stn = jn % C::kStages;
        const uint32_t phn = (uint32_t)(jn / C::kStages) & 1;
        const bool more = jn < n_tiles;

        mbar_wait(&bars->v_full[st], ph, err_flag, 203, dead);
        mbar_wait(&bars->p_full[0], j & 1, err_flag, 204, dead);
        tc_fence_after();
        issue_pv(0, st, j > 0);
        mma_commit(&bars->pv_done[0]);
        if (more) {
          mbar_wait(&bars->k_full[stn], phn, err_flag, 205, dead);
          tc_fence_after();
          issue_qk(0, stn);
          mma_commit(&bars->s_full[0]);
        }
        mbar_wait(&bars->p_full[1], j & 1, err_flag, 206, dead);
        tc_fence_after();
        issue_pv(1, st, j > 0);
        mma_commit(&bars->pv_done[1]);
        mma_commit(&bars->v_empty[st]);
        if (more) {
          issue_qk(1, stn);
          mma_commit(&bars->s_full[1]);
          mma_commit(&bars->k_empty[stn]);
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

    float c = prm.scale_log2;  // log2(e) / sqrt(d)
    float out_scale = 1.0f;
    if constexpr (kInt8) {
      const float sq = prm.scales[unit];
      const float sk = prm.scales[prm.units + unit];
      out_scale = prm.scales[2 * prm.units + unit];
      c = sq * sk * c;
    }

    float m_used = -INFINITY;
    float l0 = 0.f, l1 = 0.f;

    for (int j = 0; j < n_tiles; ++j) {
      mbar_wait(&bars->s_full[t], j & 1, err_flag, 301 + t, dead);
      dead = __any_sync(0xffffffffu, dead);
      tc_fence_after();

      uint32_t s[128];
#pragma unroll
      for (int i = 0; i < 4; ++i) tmem_ld32(tS + i * 32, &s[i * 32]);
      tmem_wait_ld();

      const int n_valid = prm.N - j * kBN;  // keys of this tile that exist (>=1)
      const bool masked = n_valid < kBN;
      float mt = masked ? tile_row_max<kInt8, true>(s, c, n_valid)
                        : tile_row_max<kInt8, false>(s, c, n_valid);

      // Lazy max: keep the old reference max unless the new one is more than 2^kRescaleThreshold
      // larger.  The decision is made warp-uniform because tcgen05.ld/st are warp collectives.
      const bool need = mt > m_used + kRescaleThreshold;
      if (__any_sync(0xffffffffu, need)) {
        const float m_new = need ? mt : m_used;
        if (j > 0) {
          const float alpha = need ? ex2_approx(m_used - m_new) : 1.0f;
          l0 *= alpha;
          l1 *= alpha;
          mbar_wait(&bars->pv_done[t], (j - 1) & 1, err_flag, 311 + t, dead);
          dead = __any_sync(0xffffffffu, dead);
          tc_fence_after();
#pragma unroll
          for (int ch = 0; ch < kD / 32; ++ch) {
            uint32_t o[32];
            tmem_ld32(tO + ch * 32, o);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st32(tO + ch * 32, o);
          }
        }
        m_used = m_new;
      }

      uint32_t p[64];
      if (masked) tile_row_exp<kInt8, kFastExp, true>(s, p, c, m_used, n_valid, l0, l1);
      else tile_row_exp<kInt8, kFastExp, false>(s, p, c, m_used, n_valid, l0, l1);

      tmem_st32(tS, &p[0]);
      tmem_st32(tS + 32, &p[32]);
      tmem_wait_st();
      tc_fence_before();
      mbar_arrive(&bars->p_full[t]);
    }

    // ---------------------------------------------------------------- epilogue: O * sV / l
    mbar_wait(&bars->pv_done[t], (n_tiles - 1) & 1, err_flag, 321 + t, dead);
    dead = __any_sync(0xffffffffu, dead);
    tc_fence_after();
    const float l = l0 + l1;
    const float inv = (l > 0.f) ? out_scale / l : 0.f;  // fa_tc_int8_b.cu:549-553 guard
    const int row = q_base + t * kBM + row_in_tile;
    const int b = unit / prm.H, head = unit % prm.H;
    float* out = prm.O + ((size_t)b * prm.N + row) * ((size_t)prm.H * prm.d) + (size_t)head * prm.d;
    const bool row_ok = row < prm.N;
    const bool vec_ok = (prm.d & 3) == 0;
#pragma unroll
    for (int ch = 0; ch < kD / 32; ++ch) {
      uint32_t o[32];
      tmem_ld32(tO + ch * 32, o);
      tmem_wait_ld();
      if (row_ok) {
        if (vec_ok) {
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            const int col = ch * 32 + i;
            if (col < prm.d) {
              float4 v = make_float4(__uint_as_float(o[i]) * inv, __uint_as_float(o[i + 1]) * inv,
                                     __uint_as_float(o[i + 2]) * inv, __uint_as_float(o[i + 3]) * inv);
              *reinterpret_cast<float4*>(out + col) = v;
            }
          }
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const int col = ch * 32 + i;
            if (col < prm.d) out[col] = __uint_as_float(o[i]) * inv;
          }
        }
      }
    }
  }

  // ---------------------------------------------------------------------------- teardown
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
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

template <bool kInt8, int kD, bool kFastExp>
bool launch_cfg(const AttnLaunch& a, std::string* err) {
  using C = Cfg<kInt8, kD>;
  const uint64_t units = (uint64_t)a.B * a.H;
  CUtensorMap tq, tk, tv;
  if (!make_map_2d(&tq, a.Qp, C::kEltQK, units * a.n_pad, kD, kBM, C::kAtomQK / C::kEltQK, err) ||
      !make_map_2d(&tk, a.Kp, C::kEltQK, units * a.n_pad, kD, kBN, C::kAtomQK / C::kEltQK, err) ||
      !make_map_2d(&tv, a.Vt, 2, units * kD, a.n_pad, kD, 64, err))
    return false;
  auto kern = attn_fwd_kernel<kInt8, kD, kFastExp>;
  {  // per device (context) attribute; cheap enough to set on every launch
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         C::kSmemBytes);
    if (e != cudaSuccess) { *err = std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(e); return false; }
  }
  AttnParams p;
  p.O = a.O;
  p.scales = a.scales;
  p.error_flag = a.error_flag;
  p.B = a.B; p.N = a.N; p.H = a.H; p.d = a.d;
  p.n_pad = a.n_pad;
  p.units = (int)units;
  p.n_kv_tiles = (a.N + kBN - 1) / kBN;
  p.scale_log2 = 1.4426950408889634f / sqrtf((float)a.d);
  dim3 grid((a.N + 2 * kBM - 1) / (2 * kBM), (unsigned)units, 1);
  kern<<<grid, kThreads, C::kSmemBytes, a.stream>>>(tq, tk, tv, p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { *err = std::string("attention launch: ") + cudaGetErrorString(e); return false; }
  return true;
}

}  // namespace

bool launch_attention(const AttnLaunch& a, std::string* err) {
  if (a.units_y_limit_exceeded()) { *err = "B*h exceeds the CUDA grid.y limit (65535)"; return false; }
  if (a.int8) {
    switch (a.d_pad) {
      case 32: return launch_cfg<true, 32, true>(a, err);
      case 64: return launch_cfg<true, 64, true>(a, err);
      case 128: return launch_cfg<true, 128, true>(a, err);
    }
  } else {
    switch (a.d_pad) {
      case 32: return launch_cfg<false, 32, false>(a, err);
      case 64: return launch_cfg<false, 64, false>(a, err);
      case 128: return launch_cfg<false, 128, false>(a, err);
    }
  }
  *err = "unsupported padded head dimension";
  return false;
}

}  // namespace qmha
