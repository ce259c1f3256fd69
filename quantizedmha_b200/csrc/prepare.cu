// prepare.cu — HBM-bound operand preparation (kernel (a) of the north star).
//
// Replaces fp32_to_int8sram (mha_kernels/fa_tc_int8_b.cu:33-152), which the reference runs four
// times per KV step inside the attention loop (Q re-quantised every step, K/V re-quantised by
// every query block), and extract_mat/concat_mat (utils/utils.cu:6-22).  Here every element of
// Q, K, V is read from HBM, quantised ONCE and written in the layout the tcgen05 kernel's TMA
// descriptors want:
//      Qp, Kp : [B*h, n_pad, d_pad]  int8 (or fp16 for the FP16 variant)
//      Vt     : [B*h, d_pad, n_pad]  fp16, transposed so P·V's B operand is K-major
// Arithmetic is the reference's, bit for bit (fa_tc_int8_b.cu:104-106,136-140):
//      sc = fmaxf(absmax / 127.0f, 1e-8f);  inv = 1.0f / sc;
//      q  = clamp(__float2int_rn(v * inv), -128, 127)
// with one scale per tensor / per (batch, head) / per 32-row block.
//
// Kernels:  absmax_kernel  -> per-(batch,head) |x| maxima (128-bit loads, shuffle + smem reduce)
//           finalize_scales_kernel -> scales from maxima (optionally reduced to per-tensor)
//           prepare_kernel -> quantise/convert + re-layout (+ transpose of V through smem)
//           quantize_blocks_kernel / quantize_static_kernel -> reference-granularity and
//           golden-spec (generate_golden.cpp:94-101) quantisers in the input layout.
#include <cooperative_groups.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "prepare.cuh"

namespace qmha {

namespace {

__device__ __forceinline__ float4 ldg_f4(const float* p) {
  return __ldg(reinterpret_cast<const float4*>(p));
}

// Input element types of the extended entry (qmha_forward_ex: fp32 like the reference, or fp16 / bf16 so that
// 16-bit callers skip the fp32 round trip: 2 instead of 4 bytes read per element).  ld4 = four consecutive
// elements as fp32 (16-byte / 8-byte read-only load), ld1 = one element.
template <typename T> struct In;
template <> struct In<float> {
  static __device__ __forceinline__ float4 ld4(const float* p) { return ldg_f4(p); }
  static __device__ __forceinline__ float ld1(const float* p) { return __ldg(p); }
};
template <> struct In<__half> {
  static __device__ __forceinline__ float4 ld4(const __half* p) {
    const uint2 u = __ldg(reinterpret_cast<const uint2*>(p));
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&u.x));
    const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
    return make_float4(a.x, a.y, b.x, b.y);
  }
  static __device__ __forceinline__ float ld1(const __half* p) { return __half2float(__ldg(p)); }
};
template <> struct In<__nv_bfloat16> {
  static __device__ __forceinline__ float4 ld4(const __nv_bfloat16* p) {
    const uint2 u = __ldg(reinterpret_cast<const uint2*>(p));
    // bf16 -> fp32 is a 16-bit shift
    return make_float4(__uint_as_float(u.x << 16), __uint_as_float(u.x & 0xFFFF0000u),
                       __uint_as_float(u.y << 16), __uint_as_float(u.y & 0xFFFF0000u));
  }
  static __device__ __forceinline__ float ld1(const __nv_bfloat16* p) {
    return __uint_as_float((uint32_t)__ldg(reinterpret_cast<const unsigned short*>(p)) << 16);
  }
};
// 16-bit operand formats of the attention kernel: kOut 1 = fp16, 2 = bf16 (0 = int8 codes; V codes are fp16)
template <int kOut>
__device__ __forceinline__ uint16_t cvt16(float x) {
  if constexpr (kOut == 2) return __bfloat16_as_ushort(__float2bfloat16_rn(x));
  else return __half_as_ushort(__float2half_rn(x));
}

__device__ __forceinline__ float absmax4(float m, float4 v) {
  return fmaxf(fmaxf(fmaxf(m, fabsf(v.x)), fmaxf(fabsf(v.y), fabsf(v.z))), fabsf(v.w));
}

// fa_tc_int8_b.cu:136-140
__device__ __forceinline__ int quant1(float v, float inv_sc) {
  int r = __float2int_rn(v * inv_sc);
  return r < -128 ? -128 : (r > 127 ? 127 : r);
}


// Fused RoPE (next row of SURVEY §8f; reference: utils/verify.cu:9-23 == generate_golden.cpp:38-51
// == the dead device helper utils/utils.cu:50-65).  Element k of a head row pairs with k + d/2:
//   row[k]       = x*cos - y*sin          row[k + d/2] = x*sin + y*cos
// cos/sin come from a host-built table (same libm as the CPU reference, so the rotated fp32 values
// — and therefore the INT8 codes and scales — stay bit-identical to the CPU restatement); products
// and sums are rounded separately (__fmul_rn / __fadd_rn: no FMA contraction, as in the scalar
// CPU loop).  `x` holds columns [4*vec, 4*vec+4) of row n; the partner columns live in another lane
// of the same row group (rows are spread over kVecPerRow consecutive lanes), fetched by shuffle.
// Requires d % 8 == 0.  All 32 lanes must call this.
__device__ __forceinline__ float4 rope_rotate(float4 x, int vec, int n, int N, int d,
                                              const float2* __restrict__ tab) {
  const int half_vecs = d >> 3;
  const bool valid = vec * 4 < d;
  const bool lo = vec < half_vecs;
  const int pvec = lo ? vec + half_vecs : vec - half_vecs;
  const int lane = threadIdx.x & 31;
  const int src = valid ? lane - vec + pvec : lane;
  float4 y;
  y.x = __shfl_sync(0xffffffffu, x.x, src);
  y.y = __shfl_sync(0xffffffffu, x.y, src);
  y.z = __shfl_sync(0xffffffffu, x.z, src);
  y.w = __shfl_sync(0xffffffffu, x.w, src);
  if (!valid || n >= N) return x;
  const int kk = (lo ? vec : pvec) * 4;
  const float4* t = reinterpret_cast<const float4*>(tab + (size_t)n * (d >> 1) + kk);
  const float4 t0 = __ldg(t), t1 = __ldg(t + 1);  // {cos0, sin0, cos1, sin1}, {cos2, sin2, cos3, sin3}
  float4 r;
  if (lo) {  // x = row[k], y = row[k + d/2]
    r.x = __fsub_rn(__fmul_rn(x.x, t0.x), __fmul_rn(y.x, t0.y));
    r.y = __fsub_rn(__fmul_rn(x.y, t0.z), __fmul_rn(y.y, t0.w));
    r.z = __fsub_rn(__fmul_rn(x.z, t1.x), __fmul_rn(y.z, t1.y));
    r.w = __fsub_rn(__fmul_rn(x.w, t1.z), __fmul_rn(y.w, t1.w));
  } else {   // y = row[k] (partner), x = row[k + d/2] (mine)
    r.x = __fadd_rn(__fmul_rn(y.x, t0.y), __fmul_rn(x.x, t0.x));
    r.y = __fadd_rn(__fmul_rn(y.y, t0.w), __fmul_rn(x.y, t0.z));
    r.z = __fadd_rn(__fmul_rn(y.z, t1.y), __fmul_rn(x.z, t1.x));
    r.w = __fadd_rn(__fmul_rn(y.w, t1.w), __fmul_rn(x.w, t1.z));
  }
  return r;
}

// ------------------------------------------------------------------------------------------------
// absmax over [rows r0..r1) of batch b for every head; grid = (row_chunks, B, 3 tensors).
// amax_bits[z][b*H + head] accumulates max |x| as the bit pattern of a non-negative float
// (monotone under unsigned compare, so atomicMax works).
constexpr int kAbsmaxThreads = 256;
constexpr int kAbsmaxRows = 32;

template <typename TIn>
__global__ void __launch_bounds__(kAbsmaxThreads)
absmax_kernel(const TIn* __restrict__ Q, const TIn* __restrict__ K,
              const TIn* __restrict__ V, unsigned* __restrict__ amax_bits, int N, int H, int d, int d_pad,
              const float2* __restrict__ rope, int ld, size_t bs) {
  extern __shared__ unsigned s_amax[];  // [H]
  const int z = blockIdx.z, b = blockIdx.y;
  const TIn* X = z == 0 ? Q : (z == 1 ? K : V);
  const int d_model = H * d;   // logical row width; ld / bs = row / batch pitch of the input in elements
  const int r0 = blockIdx.x * kAbsmaxRows;
  const int r1 = min(N, r0 + kAbsmaxRows);
  for (int i = threadIdx.x; i < H; i += blockDim.x) s_amax[i] = 0u;
  __syncthreads();
  const TIn* base = X + (size_t)b * bs;
  if (rope != nullptr && z < 2) {
    // Fused RoPE (per-tensor / two-pass path): the maxima must be those of the ROTATED rows.  Slots are laid
    // out per head over the padded head dimension (d_pad/4 lanes per head row, a divisor of 32), so the
    // partner element of the rotation sits in the same warp; the trip count is warp-uniform.
    const int vpr = d_pad >> 2;
    const int slots = H * vpr, slots_up = (slots + 31) & ~31;
    for (int sidx = threadIdx.x; sidx < slots_up; sidx += blockDim.x) {
      const bool live = sidx < slots;
      const int head = live ? sidx / vpr : 0;
      const int vec = live ? sidx % vpr : vpr;          // vpr * 4 >= d: treated as padding by rope_rotate
      const bool col_ok = live && vec * 4 < d;
      const TIn* col = base + (size_t)head * d + vec * 4;
      float m = 0.f;
      for (int r = r0; r < r1; ++r) {
        float4 x = col_ok ? In<TIn>::ld4(col + (size_t)r * ld) : make_float4(0.f, 0.f, 0.f, 0.f);
        x = rope_rotate(x, vec, r, N, d, rope);
        m = absmax4(m, x);
      }
      if (col_ok) atomicMax(&s_amax[head], __float_as_uint(m));
    }
  } else if ((d & 3) == 0) {
    const int vecs = d_model >> 2;
    for (int v = threadIdx.x; v < vecs; v += blockDim.x) {
      const int head = (v << 2) / d;
      const TIn* col = base + (size_t)(v << 2);
      float m = 0.f;
      int r = r0;
      for (; r + 8 <= r1; r += 8) {
        float4 x[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) x[u] = In<TIn>::ld4(col + (size_t)(r + u) * ld);
#pragma unroll
        for (int u = 0; u < 8; ++u) m = absmax4(m, x[u]);
      }
      for (; r < r1; ++r) m = absmax4(m, In<TIn>::ld4(col + (size_t)r * ld));
      atomicMax(&s_amax[head], __float_as_uint(m));
    }
  } else {
    for (int c = threadIdx.x; c < d_model; c += blockDim.x) {
      float m = 0.f;
      for (int r = r0; r < r1; ++r) m = fmaxf(m, fabsf(In<TIn>::ld1(base + (size_t)r * ld + c)));
      atomicMax(&s_amax[c / d], __float_as_uint(m));
    }
  }
  __syncthreads();
  unsigned* out = amax_bits + ((size_t)z * gridDim.y + b) * H;
  for (int i = threadIdx.x; i < H; i += blockDim.x)
    if (s_amax[i]) atomicMax(&out[i], s_amax[i]);
}

// scales[z][u] = fmaxf(amax/127, 1e-8)  (fa_tc_int8_b.cu:104); per-tensor = max over all u first.
__global__ void finalize_scales_kernel(const unsigned* __restrict__ amax_bits,
                                       float* __restrict__ scales, int units, int per_tensor) {
  const int z = blockIdx.x;
  __shared__ unsigned s_max;
  if (threadIdx.x == 0) s_max = 0u;
  __syncthreads();
  if (per_tensor) {
    unsigned m = 0u;
    for (int u = threadIdx.x; u < units; u += blockDim.x) m = max(m, amax_bits[(size_t)z * units + u]);
    atomicMax(&s_max, m);
    __syncthreads();
  }
  for (int u = threadIdx.x; u < units; u += blockDim.x) {
    const float a = __uint_as_float(per_tensor ? s_max : amax_bits[(size_t)z * units + u]);
    scales[(size_t)z * units + u] = fmaxf(a / 127.0f, 1e-8f);
  }
}

// ------------------------------------------------------------------------------------------------
// prepare_kernel: one CTA = one 128-row tile of one (batch, head) of one tensor.
//   grid = (n_pad/128, B*H, 3);  z = 0:Q 1:K 2:V
constexpr int kPrepThreads = 256;
constexpr int kPrepRows = 128;

template <int kOut, int kD, typename TIn>
__global__ void __launch_bounds__(kPrepThreads)   // (a six-CTA register budget was measured: no gain)
prepare_kernel(const TIn* __restrict__ Q, const TIn* __restrict__ K,
               const TIn* __restrict__ V, const float* __restrict__ scales, void* __restrict__ Qp,
               void* __restrict__ Kp, uint16_t* __restrict__ Vt, int N, int H, int d, int n_pad,
               const float2* __restrict__ rope, int ld, size_t bs) {
  constexpr bool kInt8 = kOut == 0 || kOut == 3;   // 3 = int8 codes with an int8 V^T (INT8 P.V mode)
  const int z = blockIdx.z, unit = blockIdx.y;
  const int b = unit / H, head = unit % H;
  const int n0 = blockIdx.x * kPrepRows;
  const TIn* X = z == 0 ? Q : (z == 1 ? K : V);
  const int d_model = ld;      // row pitch of the input in elements (dense: H * d); bs = batch pitch
  const TIn* src = X + (size_t)b * bs + (size_t)head * d;
  float inv_sc = 1.0f;
  if constexpr (kInt8) inv_sc = 1.0f / scales[(size_t)z * gridDim.y + unit];  // fa_tc_int8_b.cu:106

  constexpr int kVecPerRow = kD / 4;                   // float4 slots per padded row
  constexpr int kRowsPerIter = kPrepThreads / kVecPerRow;
  const int vec = threadIdx.x % kVecPerRow;
  const int rsub = threadIdx.x / kVecPerRow;
  const bool vec_ok = (d & 3) == 0;

  auto load4 = [&](int n, int c, float (&x)[4]) {
    x[0] = x[1] = x[2] = x[3] = 0.f;
    if (n < N) {
      const TIn* p = src + (size_t)n * d_model + c;
      if (vec_ok) {
        if (c < d) { float4 v = In<TIn>::ld4(p); x[0] = v.x; x[1] = v.y; x[2] = v.z; x[3] = v.w; }
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e) if (c + e < d) x[e] = In<TIn>::ld1(p + e);
      }
    }
  };

  if (z < 2) {
    void* dst = z == 0 ? Qp : Kp;
#pragma unroll 8
    for (int r = rsub; r < kPrepRows; r += kRowsPerIter) {
      const int n = n0 + r;
      float x[4];
      load4(n, vec * 4, x);
      if (rope) {  // uniform branch; r covers whole warps, so all lanes take part in the shuffles
        const float4 rr = rope_rotate(make_float4(x[0], x[1], x[2], x[3]), vec, n, N, d, rope);
        x[0] = rr.x; x[1] = rr.y; x[2] = rr.z; x[3] = rr.w;
      }
      const size_t o = ((size_t)unit * n_pad + n) * kD + vec * 4;
      if constexpr (kInt8) {
        const int q0 = quant1(x[0], inv_sc), q1 = quant1(x[1], inv_sc);
        const int q2 = quant1(x[2], inv_sc), q3 = quant1(x[3], inv_sc);
        const uint32_t pk = (uint32_t)(q0 & 0xFF) | ((uint32_t)(q1 & 0xFF) << 8) |
                            ((uint32_t)(q2 & 0xFF) << 16) | ((uint32_t)(q3 & 0xFF) << 24);
        *reinterpret_cast<uint32_t*>(reinterpret_cast<int8_t*>(dst) + o) = pk;
      } else {
        uint2 pk;
        pk.x = (uint32_t)cvt16<kOut>(x[0]) | ((uint32_t)cvt16<kOut>(x[1]) << 16);
        pk.y = (uint32_t)cvt16<kOut>(x[2]) | ((uint32_t)cvt16<kOut>(x[3]) << 16);
        *reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(dst) + o) = pk;
      }
    }
  } else if constexpr (kOut == 3) {
    // V for the INT8 P.V mode: int8 codes, transposed (keys contiguous): Vt8[unit][dd][n]
    constexpr int kStride8 = kD + 4;   // bytes
    __shared__ __align__(4) int8_t tile8[kPrepRows * kStride8];
    int8_t* Vt8 = reinterpret_cast<int8_t*>(Vt);
    for (int r = rsub; r < kPrepRows; r += kRowsPerIter) {
      float x[4];
      load4(n0 + r, vec * 4, x);
      const int q0 = quant1(x[0], inv_sc), q1 = quant1(x[1], inv_sc), q2 = quant1(x[2], inv_sc), q3 = quant1(x[3], inv_sc);
      *reinterpret_cast<uint32_t*>(tile8 + r * kStride8 + vec * 4) =
          (uint32_t)(q0 & 0xFF) | ((uint32_t)(q1 & 0xFF) << 8) | ((uint32_t)(q2 & 0xFF) << 16) | ((uint32_t)(q3 & 0xFF) << 24);
    }
    __syncthreads();
    constexpr int kQuads = kPrepRows / 4;   // threads per d-row: each emits 4 consecutive keys
    const int kq = threadIdx.x % kQuads;
    for (int dd = threadIdx.x / kQuads; dd < kD; dd += kPrepThreads / kQuads) {
      const uint32_t o4 = (uint32_t)(uint8_t)tile8[(4 * kq + 0) * kStride8 + dd] | ((uint32_t)(uint8_t)tile8[(4 * kq + 1) * kStride8 + dd] << 8) |
                          ((uint32_t)(uint8_t)tile8[(4 * kq + 2) * kStride8 + dd] << 16) | ((uint32_t)(uint8_t)tile8[(4 * kq + 3) * kStride8 + dd] << 24);
      *reinterpret_cast<uint32_t*>(Vt8 + ((size_t)unit * kD + dd) * n_pad + n0 + 4 * kq) = o4;
    }
  } else {
    // V: quantise/convert into a shared tile, then write it transposed (keys contiguous).
    constexpr int kStride = kD + 2;  // halves; odd word stride spreads the transposed reads
    __shared__ uint16_t tile[kPrepRows * kStride];
    // all loads of the thread are issued before the first conversion (one load in flight per thread
    // leaves HBM idle: this third of the grid ran at ~60 % of the rest)
    constexpr int kIters = kPrepRows / kRowsPerIter;   // 4 / 8 / 16 rows per thread
    constexpr int kBatch = kIters < 8 ? kIters : 8;
#pragma unroll
    for (int r0 = 0; r0 < kIters; r0 += kBatch) {
      float x[kBatch][4];
#pragma unroll
      for (int k = 0; k < kBatch; ++k) load4(n0 + rsub + (r0 + k) * kRowsPerIter, vec * 4, x[k]);
#pragma unroll
      for (int k = 0; k < kBatch; ++k) {
        const int r = rsub + (r0 + k) * kRowsPerIter;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float y = x[k][e];
          if constexpr (kInt8) y = (float)quant1(x[k][e], inv_sc);  // int8 code, exact in fp16
          tile[r * kStride + vec * 4 + e] = cvt16<kOut>(y);
        }
      }
    }
    __syncthreads();
    // each thread emits 2 consecutive keys (4 bytes) of one d-row; 64 threads cover 128 keys.
    const int kp = threadIdx.x & 63;
    for (int dd = threadIdx.x >> 6; dd < kD; dd += kPrepThreads / 64) {
      const uint32_t o2 = (uint32_t)tile[(2 * kp) * kStride + dd] | ((uint32_t)tile[(2 * kp + 1) * kStride + dd] << 16);
      *reinterpret_cast<uint32_t*>(Vt + ((size_t)unit * kD + dd) * n_pad + n0 + 2 * kp) = o2;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// block_quantize_kernel: the reference's own granularity — one scale per (batch, head, 32-row
// block), exactly fp32_to_int8sram on a Br x d / Bc x d tile (fa_tc_int8_b.cu:484,496,518) — in
// ONE pass: a CTA keeps its kBlkRows-row tile (kBlkRows/32 blocks) in registers, reduces the block maxima,
// then quantises from registers.  HBM traffic = algorithmic (read fp32 once, write codes once).
//   grid = (n_pad/kBlkRows, B*H, 3);  scales[z][unit][n_pad/32]
// kBlkRows = rows per CTA (a multiple of 32): 64 keeps the tile at 8 float4 per thread.  Occupancy decides this
// kernel (measured at B8/H32/N8192/d128, same box, ms per pass): 128 rows x 2 CTAs/SM 0.83, 64 x 3 (78 registers)
// 0.81, 64 x 4 (64 registers) 0.73, 64 x 5 (48) 0.75, 64 x 6 (40 registers, 16 bytes of spill) 0.65 = the HBM
// copy peak, 64 x 8 (32 registers, 132 bytes of spill) 0.88, 32 x 8 0.70.  The RoPE instantiation keeps the
// three-CTA budget (it needs the registers for the rotation).
#ifndef QMHA_BLKQ_ROWS
#define QMHA_BLKQ_ROWS 64
#endif
constexpr int kBlkRows = QMHA_BLKQ_ROWS;
#ifndef QMHA_BLKQ_CTAS
#define QMHA_BLKQ_CTAS 6
#endif
template <int kD, bool kRope, typename TIn, bool kV8 = false>
__global__ void __launch_bounds__(kPrepThreads, kRope ? 3 : QMHA_BLKQ_CTAS)
block_quantize_kernel(const TIn* __restrict__ Q, const TIn* __restrict__ K,
                      const TIn* __restrict__ V, float* __restrict__ scales,
                      int8_t* __restrict__ Qp, int8_t* __restrict__ Kp, __half* __restrict__ Vt,
                      int N, int H, int d, int n_pad, const float2* __restrict__ rope, int ld, size_t bs) {
  const int z = blockIdx.z, unit = blockIdx.y;  // (heads-fastest CTA order was measured: no gain)
  const int b = unit / H, head = unit % H;
  const int n0 = blockIdx.x * kBlkRows;
  const TIn* X = z == 0 ? Q : (z == 1 ? K : V);
  const int d_model = ld;      // row pitch of the input in elements (dense: H * d); bs = batch pitch
  const TIn* src = X + (size_t)b * bs + (size_t)head * d;

  constexpr int kVecPerRow = kD / 4;
  constexpr int kRowsPerIter = kPrepThreads / kVecPerRow;  // 8 / 16 / 32 rows per pass
  constexpr int kLoads = kBlkRows / kRowsPerIter;           // 8 / 4 / 2 float4 per thread
  constexpr int kBlocks = kBlkRows / 32;                    // 32-row scale blocks per CTA
  constexpr int kPerBlock = 32 / kRowsPerIter;              // loads of one thread per 32-row block
  const int vec = threadIdx.x % kVecPerRow;
  const int rsub = threadIdx.x / kVecPerRow;
  const bool col_ok = vec * 4 < d;                          // host guarantees d % 4 == 0

  float4 x[kLoads];
#pragma unroll
  for (int k = 0; k < kLoads; ++k) {
    const int n = n0 + rsub + k * kRowsPerIter;
    x[k] = (n < N && col_ok) ? In<TIn>::ld4(src + (size_t)n * d_model + vec * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  if constexpr (kRope) {
    if (z < 2) {  // rotate Q and K rows before the block maxima are taken (uniform branch)
#pragma unroll
      for (int k = 0; k < kLoads; ++k)
        x[k] = rope_rotate(x[k], vec, n0 + rsub + k * kRowsPerIter, N, d, rope);
    }
  }
  // block maxima: thread -> warp (shuffle) -> CTA (shared memory)
  __shared__ float s_max[kPrepThreads / 32][kBlocks];
  float m[kBlocks];
#pragma unroll
  for (int blk = 0; blk < kBlocks; ++blk) {
    float v = 0.f;
#pragma unroll
    for (int k = 0; k < kPerBlock; ++k) v = absmax4(v, x[blk * kPerBlock + k]);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, off));
    m[blk] = v;
  }
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int blk = 0; blk < kBlocks; ++blk) s_max[threadIdx.x >> 5][blk] = m[blk];
  }
  __syncthreads();
  float inv_sc[kBlocks];
#pragma unroll
  for (int blk = 0; blk < kBlocks; ++blk) {
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < kPrepThreads / 32; ++w) v = fmaxf(v, s_max[w][blk]);
    const float sc = fmaxf(v / 127.0f, 1e-8f);  // fa_tc_int8_b.cu:104
    inv_sc[blk] = 1.0f / sc;                    // fa_tc_int8_b.cu:106
    if (threadIdx.x == 0)
      scales[((size_t)z * gridDim.y + unit) * (n_pad / 32) + (n0 / 32) + blk] = sc;
  }

  if (z < 2) {
    int8_t* dst = z == 0 ? Qp : Kp;
#pragma unroll
    for (int k = 0; k < kLoads; ++k) {
      const int n = n0 + rsub + k * kRowsPerIter;
      const float inv = inv_sc[k / kPerBlock];
      const int q0 = quant1(x[k].x, inv), q1 = quant1(x[k].y, inv);
      const int q2 = quant1(x[k].z, inv), q3 = quant1(x[k].w, inv);
      const uint32_t pk = (uint32_t)(q0 & 0xFF) | ((uint32_t)(q1 & 0xFF) << 8) |
                          ((uint32_t)(q2 & 0xFF) << 16) | ((uint32_t)(q3 & 0xFF) << 24);
      *reinterpret_cast<uint32_t*>(dst + ((size_t)unit * n_pad + n) * kD + vec * 4) = pk;
    }
  } else if constexpr (kV8) {
    // INT8 P.V mode: V^T as int8 codes, Vt8[unit][dd][n]
    constexpr int kStride8 = kD + 4;   // bytes
    __shared__ __align__(4) int8_t tile8[kBlkRows * kStride8];
    int8_t* Vt8 = reinterpret_cast<int8_t*>(Vt);
#pragma unroll
    for (int k = 0; k < kLoads; ++k) {
      const int r = rsub + k * kRowsPerIter;
      const float inv = inv_sc[k / kPerBlock];
      const int q0 = quant1(x[k].x, inv), q1 = quant1(x[k].y, inv), q2 = quant1(x[k].z, inv), q3 = quant1(x[k].w, inv);
      *reinterpret_cast<uint32_t*>(tile8 + r * kStride8 + vec * 4) =
          (uint32_t)(q0 & 0xFF) | ((uint32_t)(q1 & 0xFF) << 8) | ((uint32_t)(q2 & 0xFF) << 16) | ((uint32_t)(q3 & 0xFF) << 24);
    }
    __syncthreads();
    constexpr int kQuads = kBlkRows / 4;
    const int kq = threadIdx.x % kQuads;
    for (int dd = threadIdx.x / kQuads; dd < kD; dd += kPrepThreads / kQuads) {
      const uint32_t o4 = (uint32_t)(uint8_t)tile8[(4 * kq + 0) * kStride8 + dd] | ((uint32_t)(uint8_t)tile8[(4 * kq + 1) * kStride8 + dd] << 8) |
                          ((uint32_t)(uint8_t)tile8[(4 * kq + 2) * kStride8 + dd] << 16) | ((uint32_t)(uint8_t)tile8[(4 * kq + 3) * kStride8 + dd] << 24);
      *reinterpret_cast<uint32_t*>(Vt8 + ((size_t)unit * kD + dd) * n_pad + n0 + 4 * kq) = o4;
    }
  } else {
    constexpr int kStride = kD + 2;
    __shared__ __half tile[kBlkRows * kStride];
#pragma unroll
    for (int k = 0; k < kLoads; ++k) {
      const int r = rsub + k * kRowsPerIter;
      const float inv = inv_sc[k / kPerBlock];
      tile[r * kStride + vec * 4 + 0] = __float2half_rn((float)quant1(x[k].x, inv));
      tile[r * kStride + vec * 4 + 1] = __float2half_rn((float)quant1(x[k].y, inv));
      tile[r * kStride + vec * 4 + 2] = __float2half_rn((float)quant1(x[k].z, inv));
      tile[r * kStride + vec * 4 + 3] = __float2half_rn((float)quant1(x[k].w, inv));
    }
    __syncthreads();
    constexpr int kPairs = kBlkRows / 2;  // threads per d-row: each emits 2 consecutive keys
    const int kp = threadIdx.x % kPairs;
    for (int dd = threadIdx.x / kPairs; dd < kD; dd += kPrepThreads / kPairs) {
      __half2 o2 = __halves2half2(tile[(2 * kp) * kStride + dd], tile[(2 * kp + 1) * kStride + dd]);
      *reinterpret_cast<__half2*>(Vt + ((size_t)unit * kD + dd) * n_pad + n0 + 2 * kp) = o2;
    }
  }
}

template <int kD, typename TIn>
cudaError_t launch_block_cfg(const PrepareArgs& a) {
  dim3 grid(a.n_pad / kBlkRows, a.B * a.H, 3);
  // the RoPE variant is a separate instantiation: the plain one stays inside its register budget
  auto kern = a.v8 ? (a.rope ? block_quantize_kernel<kD, true, TIn, true> : block_quantize_kernel<kD, false, TIn, true>)
                   : (a.rope ? block_quantize_kernel<kD, true, TIn> : block_quantize_kernel<kD, false, TIn>);
  kern<<<grid, kPrepThreads, 0, a.stream>>>(
      reinterpret_cast<const TIn*>(a.Q), reinterpret_cast<const TIn*>(a.K), reinterpret_cast<const TIn*>(a.V), a.scales, reinterpret_cast<int8_t*>(a.Qp), reinterpret_cast<int8_t*>(a.Kp),
      reinterpret_cast<__half*>(a.Vt), a.N, a.H, a.d, a.n_pad, a.rope, a.ld(), a.bs());
  return cudaGetLastError();
}

// V-side helper of the attention kernel in block mode: per unit the largest V scale and, per
// 32-key block, log2(r) and 1/r with r = sV_block / sV_max (see attn_fwd.cu).
//   aux[unit][nblk][2];  vmax[unit]
__global__ void block_aux_kernel(const float* __restrict__ scales_v, float* __restrict__ aux,
                                 float* __restrict__ vmax, int nblk) {
  const int unit = blockIdx.x;
  const float* sv = scales_v + (size_t)unit * nblk;
  __shared__ float s_red[32];
  float m = 0.f;
  for (int i = threadIdx.x; i < nblk; i += blockDim.x) m = fmaxf(m, sv[i]);
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off));
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = m;
  __syncthreads();
  m = 0.f;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) m = fmaxf(m, s_red[w]);
  if (threadIdx.x == 0) vmax[unit] = m;
  for (int i = threadIdx.x; i < nblk; i += blockDim.x) {
    const float r = sv[i] / m;
    aux[((size_t)unit * nblk + i) * 2 + 0] = log2f(r);
    aux[((size_t)unit * nblk + i) * 2 + 1] = 1.0f / r;
  }
}

// ------------------------------------------------------------------------------------------------
// Reference-granularity quantiser in the input layout: one warp per (batch, head, row block).
__global__ void quantize_blocks_kernel(const float* __restrict__ X, int8_t* __restrict__ q,
                                       float* __restrict__ scales, int B, int N, int H, int d,
                                       int block_rows, int nblk) {
  const int warps_per_cta = blockDim.x >> 5;
  const long long blk_id = (long long)blockIdx.x * warps_per_cta + (threadIdx.x >> 5);
  const long long total = (long long)B * H * nblk;
  if (blk_id >= total) return;
  const int lane = threadIdx.x & 31;
  const int blk = (int)(blk_id % nblk);
  const int head = (int)((blk_id / nblk) % H);
  const int b = (int)(blk_id / ((long long)nblk * H));
  const int d_model = H * d;
  const int r0 = blk * block_rows, rows = min(block_rows, N - r0);
  const size_t base = ((size_t)b * N + r0) * d_model + (size_t)head * d;
  const int elems = rows * d;
  float m = 0.f;
  for (int i = lane; i < elems; i += 32)
    m = fmaxf(m, fabsf(X[base + (size_t)(i / d) * d_model + (i % d)]));
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off));
  const float sc = fmaxf(m / 127.0f, 1e-8f);
  const float inv = 1.0f / sc;
  if (lane == 0) scales[blk_id] = sc;
  for (int i = lane; i < elems; i += 32) {
    const size_t o = base + (size_t)(i / d) * d_model + (i % d);
    q[o] = (int8_t)quant1(X[o], inv);
  }
}

// Golden spec (generate_golden.cpp:94-101): round half away from zero, true division.
__global__ void quantize_static_kernel(const float* __restrict__ X, int8_t* __restrict__ q,
                                       long long n, float scale, float zp) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    int v = (int)roundf(X[i] / scale + zp);
    v = v > 127 ? 127 : (v < -128 ? -128 : v);
    q[i] = (int8_t)v;
  }
}

// ------------------------------------------------------------------------------------------------
// fused_quantize_kernel: absmax + quantise + re-layout of one (batch, head, tensor) slab by ONE
// thread-block cluster, so the inputs cross HBM once.
//   phase 1: every CTA of the cluster reduces max|x| over its share of the slab's rows,
//            the cluster combines the CTA maxima through distributed shared memory;
//   phase 2: every CTA re-reads its rows (now L2 resident: one CTA per SM keeps the live slabs
//            of all clusters at ~64 MB of the 126 MB L2) and writes the prepared operands.
// grid = (kClusterSize, B*H, 3), cluster = (kClusterSize,1,1), 1024 threads, 1 CTA/SM (the dynamic
// shared-memory request is sized to forbid a second CTA).
constexpr int kClusterSize = 8;
constexpr int kFusedThreads = 1024;
constexpr int kFusedSmemBytes = 120 * 1024;

template <int kD, typename TIn>
__global__ void __cluster_dims__(kClusterSize, 1, 1) __launch_bounds__(kFusedThreads, 1)
fused_quantize_kernel(const TIn* __restrict__ Q, const TIn* __restrict__ K,
                      const TIn* __restrict__ V, float* __restrict__ scales,
                      int8_t* __restrict__ Qp, int8_t* __restrict__ Kp, __half* __restrict__ Vt,
                      int N, int H, int d, int n_pad, const float2* __restrict__ rope, int ld, size_t bs) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  extern __shared__ __align__(16) uint8_t fused_smem[];
  __shared__ float s_warp_max[kFusedThreads / 32];
  __shared__ float s_cta_max;

  const int z = blockIdx.z, unit = blockIdx.y;
  const int b = unit / H, head = unit % H;
  const int rank = (int)cluster.block_rank();
  const TIn* X = z == 0 ? Q : (z == 1 ? K : V);
  const int d_model = ld;      // row pitch of the input in elements (dense: H * d); bs = batch pitch
  const TIn* src = X + (size_t)b * bs + (size_t)head * d;

  // rows of this CTA: a multiple of 128 so V tiles never straddle CTAs
  const int rows_per_cta = ((n_pad / 128 + kClusterSize - 1) / kClusterSize) * 128;
  const int r_begin = rank * rows_per_cta;
  const int r_end = min(n_pad, r_begin + rows_per_cta);

  constexpr int kVecPerRow = kD / 4;                      // float4 slots per padded row
  constexpr int kRowsPerPass = kFusedThreads / kVecPerRow;
  const int vec = threadIdx.x % kVecPerRow;
  const int rsub = threadIdx.x / kVecPerRow;
  const bool col_ok = vec * 4 < d;                        // (d % 4 == 0 is required by the host)

  // ---- phase 1: absmax over my rows
  float m = 0.f;
  const bool rotate = rope != nullptr && z < 2;   // fused RoPE on Q and K (uniform for the CTA)
  if (rotate) {
    // warp-uniform loop (the partner elements of the rotation come from other lanes by shuffle):
    // every lane walks the same row slots and loads zeros where it has nothing
    for (int r0 = r_begin; r0 < r_end; r0 += kRowsPerPass) {
      const int r = r0 + rsub;
      float4 x = (r < N && col_ok) ? In<TIn>::ld4(src + (size_t)r * d_model + vec * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
      x = rope_rotate(x, vec, r, N, d, rope);
      m = absmax4(m, x);
    }
  } else if (col_ok) {
    const TIn* col = src + vec * 4;
    int r = r_begin + rsub;
    const int r_stop = min(r_end, N);
    for (; r + 7 * kRowsPerPass < r_stop; r += 8 * kRowsPerPass) {
      float4 x[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) x[u] = In<TIn>::ld4(col + (size_t)(r + u * kRowsPerPass) * d_model);
#pragma unroll
      for (int u = 0; u < 8; ++u) m = absmax4(m, x[u]);
    }
    for (; r < r_stop; r += kRowsPerPass) m = absmax4(m, In<TIn>::ld4(col + (size_t)r * d_model));
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off));
  if ((threadIdx.x & 31) == 0) s_warp_max[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x < 32) {
    float w = s_warp_max[threadIdx.x];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) w = fmaxf(w, __shfl_xor_sync(0xffffffffu, w, off));
    if (threadIdx.x == 0) s_cta_max = w;
  }
  cluster.sync();
  float amax = 0.f;
#pragma unroll
  for (int rk = 0; rk < kClusterSize; ++rk) amax = fmaxf(amax, *cluster.map_shared_rank(&s_cta_max, rk));
  cluster.sync();  // nobody may leave (or reuse s_cta_max) while peers still read it

  const float sc = fmaxf(amax / 127.0f, 1e-8f);  // fa_tc_int8_b.cu:104
  const float inv_sc = 1.0f / sc;                // fa_tc_int8_b.cu:106
  if (rank == 0 && threadIdx.x == 0) scales[(size_t)z * gridDim.y + unit] = sc;

  // ---- phase 2: quantise + re-layout my rows
  auto load4 = [&](int n, float (&x)[4]) {
    x[0] = x[1] = x[2] = x[3] = 0.f;
    if (n < N && col_ok) {
      float4 v = In<TIn>::ld4(src + (size_t)n * d_model + vec * 4);
      x[0] = v.x; x[1] = v.y; x[2] = v.z; x[3] = v.w;
    }
  };
  if (z < 2) {
    int8_t* dst = z == 0 ? Qp : Kp;
    // 8 independent 16-byte loads in flight per thread (the re-read mostly hits L2)
    for (int n = r_begin + rsub; n < r_end; n += 8 * kRowsPerPass) {
      float x[8][4];
#pragma unroll
      for (int u = 0; u < 8; ++u) load4(n + u * kRowsPerPass, x[u]);
      if (rotate) {  // same rotation as in phase 1, bit for bit
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const float4 rr = rope_rotate(make_float4(x[u][0], x[u][1], x[u][2], x[u][3]), vec,
                                        n + u * kRowsPerPass, N, d, rope);
          x[u][0] = rr.x; x[u][1] = rr.y; x[u][2] = rr.z; x[u][3] = rr.w;
        }
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int nn = n + u * kRowsPerPass;
        if (nn < r_end) {
          const int q0 = quant1(x[u][0], inv_sc), q1 = quant1(x[u][1], inv_sc);
          const int q2 = quant1(x[u][2], inv_sc), q3 = quant1(x[u][3], inv_sc);
          const uint32_t pk = (uint32_t)(q0 & 0xFF) | ((uint32_t)(q1 & 0xFF) << 8) |
                              ((uint32_t)(q2 & 0xFF) << 16) | ((uint32_t)(q3 & 0xFF) << 24);
          *reinterpret_cast<uint32_t*>(dst + ((size_t)unit * n_pad + nn) * kD + vec * 4) = pk;
        }
      }
    }
  } else {
    constexpr int kStride = kD + 2;  // halves; odd word stride spreads the transposed reads
    constexpr int kLoads = 128 / kRowsPerPass;            // float4 loads per thread per 128-row tile
    __half* tiles = reinterpret_cast<__half*>(fused_smem);  // 2 x [128][kStride], double buffered
    float x[kLoads][4];
    auto load_tile = [&](int n0) {
#pragma unroll
      for (int u = 0; u < kLoads; ++u) load4(n0 + rsub + u * kRowsPerPass, x[u]);
    };
    if (r_begin < r_end) load_tile(r_begin);
    int buf = 0;
    for (int n0 = r_begin; n0 < r_end; n0 += 128, buf ^= 1) {
      __half* tile = tiles + buf * (128 * kStride);
#pragma unroll
      for (int u = 0; u < kLoads; ++u) {
        const int r = rsub + u * kRowsPerPass;
#pragma unroll
        for (int e = 0; e < 4; ++e)
          tile[r * kStride + vec * 4 + e] = __float2half_rn((float)quant1(x[u][e], inv_sc));
      }
      if (n0 + 128 < r_end) load_tile(n0 + 128);  // next tile's loads fly during the transposed writes
      __syncthreads();
      const int kp = threadIdx.x & 63;  // 2 consecutive keys per thread, 64 threads per d-row
      for (int dd = threadIdx.x >> 6; dd < kD; dd += kFusedThreads / 64) {
        __half2 o2 = __halves2half2(tile[(2 * kp) * kStride + dd], tile[(2 * kp + 1) * kStride + dd]);
        *reinterpret_cast<__half2*>(Vt + ((size_t)unit * kD + dd) * n_pad + n0 + 2 * kp) = o2;
      }
      // (the other buffer is written next; this one is rewritten two tiles later, after a sync)
    }
  }
}

// ------------------------------------------------------------------------------------------------
// stream_quantize_kernel: per-(batch, head) scales in ONE launch without clusters (QMHA_STREAM_QUANT=1 selects it;
// the cluster kernel above stays the default).  The cluster kernel keeps only ~15 slabs in flight and alternates
// between an HBM phase and an L2 phase per cluster (1.17-1.24 ms at the headline shape = 53-56 % of the HBM roofline).
// Here a persistent grid pulls items off one queue ordered as
//     round r:  absmax of the T row tiles of slab r,  then  quantise + re-layout of the tiles of slab r - kLag
// so HBM reads (absmax items), L2 re-reads and HBM writes (quantise items) of different slabs overlap all the time,
// and the re-read of a slab follows its first read by ~kLag slabs of traffic (a few MB: L2 hits — ncu: 3.22 GB of
// DRAM reads = one read of the inputs).  Measured at the headline shape, same box (tools/quant_ab.py): 512 threads x
// 2 CTAs/SM, 128-row items 1.34 ms; 64-row items 1.96 ms; 256 threads x 4 CTAs/SM, 64-row items, lag 3 (default)
// 1.175 ms; x 6 CTAs/SM 1.24 ms; x 3 CTAs/SM 1.42 ms; cluster kernel 1.165 ms; block kernel (one pass, no second
// read) 0.63 ms.  Prefetching the queue index one item ahead and decoding items / computing the scale in one thread
// (ncu: the per-thread divisions were a quarter of the instructions, barrier stalls 39 %) changed nothing (1.19-1.21 ms).
// I.e. the second read is not an HBM problem any more (DRAM runs at 3.6 TB/s), yet every implementation of this
// granularity lands on the same ~1.17 ms: 24*E bytes of loads = 5.5 TB/s into the SMs, which is also what the block
// kernel's 12*E in 0.63 ms and the two-pass path's 24*E in 1.3 ms amount to — the load path into the SMs (half of the
// L2 hits cross the die boundary), not HBM, is the common limit, and a per-head scale needs every element twice.
// A quantise item waits for the tile counter of its slab; every absmax item of that slab was taken off the queue
// earlier by a CTA that is running, so the wait always ends.
//   ctl[0] = queue head;  amax[s], done[s] per slab s = z * units + unit (zeroed by the launcher)
#ifndef QMHA_STREAM_ROWS
#define QMHA_STREAM_ROWS 64
#endif
#ifndef QMHA_STREAM_LAG
#define QMHA_STREAM_LAG 3
#endif
#ifndef QMHA_STREAM_THREADS
#define QMHA_STREAM_THREADS 256
#endif
#ifndef QMHA_STREAM_CTAS
#define QMHA_STREAM_CTAS 4
#endif
constexpr int kStreamThreads = QMHA_STREAM_THREADS;
constexpr int kStreamCtas = QMHA_STREAM_CTAS;
constexpr int kStreamRows = QMHA_STREAM_ROWS;
constexpr int kStreamLag = QMHA_STREAM_LAG;

template <int kD, typename TIn>
__global__ void __launch_bounds__(kStreamThreads, kStreamCtas)
stream_quantize_kernel(const TIn* __restrict__ Q, const TIn* __restrict__ K, const TIn* __restrict__ V,
                       float* __restrict__ scales, int8_t* __restrict__ Qp, int8_t* __restrict__ Kp,
                       __half* __restrict__ Vt, unsigned* __restrict__ ctl, unsigned* __restrict__ amax,
                       unsigned* __restrict__ done, int N, int H, int d, int n_pad, int units,
                       const float2* __restrict__ rope, int ld, size_t bs) {
  constexpr int kVecPerRow = kD / 4;
  constexpr int kRowsPerIter = kStreamThreads / kVecPerRow;   // 16 / 32 / 64 rows per pass
  constexpr int kLoads = kStreamRows / kRowsPerIter;           // 8 / 4 / 2 float4 per thread
  constexpr int kStride = kD + 2;
  __shared__ __half tile[kStreamRows * kStride];
  __shared__ float s_warp_max[kStreamThreads / 32];
  __shared__ int s_dec[2][6];
  const int T = n_pad / kStreamRows;
  const int slabs = 3 * units;
  const unsigned total = (unsigned)(slabs + kStreamLag) * 2u * (unsigned)T;
  const int vec = threadIdx.x % kVecPerRow;
  const int rsub = threadIdx.x / kVecPerRow;
  const bool col_ok = vec * 4 < d;
  const int d_model = ld;      // row pitch of the input in elements (dense: H * d); bs = batch pitch
  // Item decode (four divisions by run-time values) and the scale arithmetic are done by ONE thread and broadcast
  // through shared memory: done per thread they were a quarter of the kernel's instructions (ncu).
  auto decode = [&](unsigned item, int* o) {   // o = {slab (-1: nothing to do), tile, quantise?, z, b, head}
    o[0] = -1;
    if (item >= total) { o[0] = -2; return; }
    const int round = (int)(item / (2u * T)), j = (int)(item % (2u * T));
    const int quant = j >= T;
    const int slab = quant ? round - kStreamLag : round;
    if (slab < 0 || slab >= slabs) return;
    const int unit = slab % units;
    o[0] = slab; o[1] = quant ? j - T : j; o[2] = quant; o[3] = slab / units; o[4] = unit / H; o[5] = unit % H;
  };
  if (threadIdx.x == 0) decode(atomicAdd(&ctl[0], 1u), s_dec[0]);
  __syncthreads();
  for (int par = 0;; par ^= 1) {
    const int slab = s_dec[par][0];
    if (slab == -2) break;
    if (slab >= 0) {
    const int t = s_dec[par][1];
    const bool quant = s_dec[par][2] != 0;
    const int z = s_dec[par][3], b = s_dec[par][4], head = s_dec[par][5];
    const int unit = slab - z * units;
    const TIn* X = z == 0 ? Q : (z == 1 ? K : V);
    const TIn* src = X + (size_t)b * bs + (size_t)head * d;
    const int n0 = t * kStreamRows;
    float4 x[kLoads];
#pragma unroll
    for (int k = 0; k < kLoads; ++k) {
      const int n = n0 + rsub + k * kRowsPerIter;
      x[k] = (n < N && col_ok) ? In<TIn>::ld4(src + (size_t)n * d_model + vec * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (rope != nullptr && z < 2) {
#pragma unroll
      for (int k = 0; k < kLoads; ++k) x[k] = rope_rotate(x[k], vec, n0 + rsub + k * kRowsPerIter, N, d, rope);
    }
    // next item: the queue round trip (~1 us) and the decode hide behind this item's loads
    if (threadIdx.x == 0) decode(atomicAdd(&ctl[0], 1u), s_dec[par ^ 1]);
    if (!quant) {
      float m = 0.f;
#pragma unroll
      for (int k = 0; k < kLoads; ++k) m = absmax4(m, x[k]);
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off));
      if ((threadIdx.x & 31) == 0) s_warp_max[threadIdx.x >> 5] = m;
      __syncthreads();
      if (threadIdx.x < 32) {
        float w = threadIdx.x < kStreamThreads / 32 ? s_warp_max[threadIdx.x] : 0.f;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) w = fmaxf(w, __shfl_xor_sync(0xffffffffu, w, off));
        if (threadIdx.x == 0) {
          atomicMax(&amax[slab], __float_as_uint(w));
          __threadfence();
          atomicAdd(&done[slab], 1u);
        }
      }
    } else {
    // quantise item: the slab maximum is final once all T absmax items of the slab have reported
    if (threadIdx.x == 0) {
      while (*((volatile unsigned*)&done[slab]) < (unsigned)T) __nanosleep(64);
      __threadfence();
      const float sc1 = fmaxf(__uint_as_float(*((volatile unsigned*)&amax[slab])) / 127.0f, 1e-8f);  // fa_tc_int8_b.cu:104
      s_warp_max[0] = sc1;
      s_warp_max[1] = 1.0f / sc1;                             // fa_tc_int8_b.cu:106
      if (t == 0) scales[slab] = sc1;
    }
    __syncthreads();
    const float inv = s_warp_max[1];
    if (z < 2) {
      int8_t* dst = z == 0 ? Qp : Kp;
#pragma unroll
      for (int k = 0; k < kLoads; ++k) {
        const int n = n0 + rsub + k * kRowsPerIter;
        const int q0 = quant1(x[k].x, inv), q1 = quant1(x[k].y, inv);
        const int q2 = quant1(x[k].z, inv), q3 = quant1(x[k].w, inv);
        const uint32_t pk = (uint32_t)(q0 & 0xFF) | ((uint32_t)(q1 & 0xFF) << 8) |
                            ((uint32_t)(q2 & 0xFF) << 16) | ((uint32_t)(q3 & 0xFF) << 24);
        *reinterpret_cast<uint32_t*>(dst + ((size_t)unit * n_pad + n) * kD + vec * 4) = pk;
      }
    } else {
#pragma unroll
      for (int k = 0; k < kLoads; ++k) {
        const int r = rsub + k * kRowsPerIter;
        tile[r * kStride + vec * 4 + 0] = __float2half_rn((float)quant1(x[k].x, inv));
        tile[r * kStride + vec * 4 + 1] = __float2half_rn((float)quant1(x[k].y, inv));
        tile[r * kStride + vec * 4 + 2] = __float2half_rn((float)quant1(x[k].z, inv));
        tile[r * kStride + vec * 4 + 3] = __float2half_rn((float)quant1(x[k].w, inv));
      }
      __syncthreads();
      constexpr int kPairs = kStreamRows / 2;   // threads per d-row: each emits 2 consecutive keys
      const int kp = threadIdx.x % kPairs;
      for (int dd = threadIdx.x / kPairs; dd < kD; dd += kStreamThreads / kPairs) {
        __half2 o2 = __halves2half2(tile[(2 * kp) * kStride + dd], tile[(2 * kp + 1) * kStride + dd]);
        *reinterpret_cast<__half2*>(Vt + ((size_t)unit * kD + dd) * n_pad + n0 + 2 * kp) = o2;
      }
    }
    }   // quantise item
    } else {   // nothing to do for this queue slot (first / last kStreamLag rounds): just fetch the next one
      if (threadIdx.x == 0) decode(atomicAdd(&ctl[0], 1u), s_dec[par ^ 1]);
    }
    __syncthreads();   // item done: the next item is visible; tile / s_warp_max may be reused
  }
}

template <int kD, typename TIn>
cudaError_t launch_stream_cfg(const PrepareArgs& a, unsigned* ctl) {
  const int units = a.B * a.H;
  // ctl layout: [0] queue head, [16 .. 16 + 3u) slab maxima, [16 + 3u .. 16 + 6u) tile counters
  cudaError_t e = cudaMemsetAsync(ctl, 0, sizeof(unsigned) * (16 + 6 * (size_t)units), a.stream);
  if (e != cudaSuccess) return e;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  stream_quantize_kernel<kD, TIn><<<kStreamCtas * sms, kStreamThreads, 0, a.stream>>>(
      reinterpret_cast<const TIn*>(a.Q), reinterpret_cast<const TIn*>(a.K), reinterpret_cast<const TIn*>(a.V), a.scales,
      reinterpret_cast<int8_t*>(a.Qp), reinterpret_cast<int8_t*>(a.Kp), reinterpret_cast<__half*>(a.Vt), ctl, ctl + 16,
      ctl + 16 + 3 * (size_t)units, a.N, a.H, a.d, a.n_pad, units, a.rope, a.ld(), a.bs());
  return cudaGetLastError();
}

template <int kD, typename TIn>
cudaError_t launch_fused_cfg(const PrepareArgs& a) {
  auto kern = fused_quantize_kernel<kD, TIn>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kFusedSmemBytes);
  if (e != cudaSuccess) return e;
  dim3 grid(kClusterSize, a.B * a.H, 3);
  kern<<<grid, kFusedThreads, kFusedSmemBytes, a.stream>>>(
      reinterpret_cast<const TIn*>(a.Q), reinterpret_cast<const TIn*>(a.K), reinterpret_cast<const TIn*>(a.V), a.scales, reinterpret_cast<int8_t*>(a.Qp), reinterpret_cast<int8_t*>(a.Kp),
      reinterpret_cast<__half*>(a.Vt), a.N, a.H, a.d, a.n_pad, a.rope, a.ld(), a.bs());
  return cudaGetLastError();
}

template <int kOut, int kD, typename TIn>
cudaError_t launch_prepare_cfg(const PrepareArgs& a) {
  dim3 grid(a.n_pad / kPrepRows, a.B * a.H, 3);
  prepare_kernel<kOut, kD, TIn><<<grid, kPrepThreads, 0, a.stream>>>(
      reinterpret_cast<const TIn*>(a.Q), reinterpret_cast<const TIn*>(a.K), reinterpret_cast<const TIn*>(a.V),
      a.scales, a.Qp, a.Kp, reinterpret_cast<uint16_t*>(a.Vt), a.N, a.H, a.d, a.n_pad, a.rope, a.ld(), a.bs());
  return cudaGetLastError();
}

// in_dtype (0 fp32, 1 fp16, 2 bf16) -> element type
#define QMHA_BY_DTYPE(a, CALL)                                    \
  switch ((a).in_dtype) {                                         \
    case 0: { using TIn = float; return CALL; }                   \
    case 1: { using TIn = __half; return CALL; }                  \
    case 2: { using TIn = __nv_bfloat16; return CALL; }           \
  }                                                               \
  return cudaErrorInvalidValue;
template <int kD> cudaError_t launch_fused_d(const PrepareArgs& a) { QMHA_BY_DTYPE(a, (launch_fused_cfg<kD, TIn>(a))) }
template <int kD> cudaError_t launch_stream_d(const PrepareArgs& a, unsigned* ctl) { QMHA_BY_DTYPE(a, (launch_stream_cfg<kD, TIn>(a, ctl))) }
template <int kD> cudaError_t launch_block_d(const PrepareArgs& a) { QMHA_BY_DTYPE(a, (launch_block_cfg<kD, TIn>(a))) }
template <int kOut, int kD> cudaError_t launch_prepare_d(const PrepareArgs& a) { QMHA_BY_DTYPE(a, (launch_prepare_cfg<kOut, kD, TIn>(a))) }
template <typename TIn>
cudaError_t launch_absmax_t(const PrepareArgs& a, unsigned* amax_bits) {
  dim3 grid((a.N + kAbsmaxRows - 1) / kAbsmaxRows, a.B, 3);
  absmax_kernel<TIn><<<grid, kAbsmaxThreads, sizeof(unsigned) * a.H, a.stream>>>(
      reinterpret_cast<const TIn*>(a.Q), reinterpret_cast<const TIn*>(a.K), reinterpret_cast<const TIn*>(a.V), amax_bits,
      a.N, a.H, a.d, a.d_pad, a.rope, a.ld(), a.bs());
  return cudaGetLastError();
}
cudaError_t launch_absmax_d(const PrepareArgs& a, unsigned* amax_bits) { QMHA_BY_DTYPE(a, (launch_absmax_t<TIn>(a, amax_bits))) }

}  // namespace

cudaError_t launch_absmax_and_scales(const PrepareArgs& a, unsigned* amax_bits, int per_tensor) {
  const int units = a.B * a.H;
  cudaError_t e = cudaMemsetAsync(amax_bits, 0, sizeof(unsigned) * 3 * units, a.stream);
  if (e != cudaSuccess) return e;
  e = launch_absmax_d(a, amax_bits);
  if (e != cudaSuccess) return e;
  finalize_scales_kernel<<<3, 256, 0, a.stream>>>(amax_bits, a.scales, units, per_tensor);
  return cudaGetLastError();
}

// Single-pass INT8 preparation with per-(batch, head) scales (requires d % 4 == 0).
cudaError_t launch_fused_quantize(const PrepareArgs& a) {
  switch (a.d_pad) {
    case 32: return launch_fused_d<32>(a);
    case 64: return launch_fused_d<64>(a);
    case 128: return launch_fused_d<128>(a);
  }
  return cudaErrorInvalidValue;
}

// Single-launch INT8 preparation with per-(batch, head) scales on a persistent grid (requires d % 4 == 0);
// ctl = 16 + 6 * B * H unsigned words of scratch.
cudaError_t launch_stream_quantize(const PrepareArgs& a, unsigned* ctl) {
  switch (a.d_pad) {
    case 32: return launch_stream_d<32>(a, ctl);
    case 64: return launch_stream_d<64>(a, ctl);
    case 128: return launch_stream_d<128>(a, ctl);
  }
  return cudaErrorInvalidValue;
}

// Single-pass INT8 preparation with per-(batch, head, 32-row block) scales (requires d % 4 == 0).
cudaError_t launch_block_quantize(const PrepareArgs& a) {
  switch (a.d_pad) {
    case 32: return launch_block_d<32>(a);
    case 64: return launch_block_d<64>(a);
    case 128: return launch_block_d<128>(a);
  }
  return cudaErrorInvalidValue;
}

cudaError_t launch_block_aux(const float* scales_v, float* aux, float* vmax, int units, int nblk,
                             cudaStream_t stream) {
  block_aux_kernel<<<units, 256, 0, stream>>>(scales_v, aux, vmax, nblk);
  return cudaGetLastError();
}

cudaError_t launch_prepare(const PrepareArgs& a) {
  if (a.int8 && a.v8) {
    switch (a.d_pad) {
      case 32: return launch_prepare_d<3, 32>(a);
      case 64: return launch_prepare_d<3, 64>(a);
      case 128: return launch_prepare_d<3, 128>(a);
    }
  } else if (a.int8) {
    switch (a.d_pad) {
      case 32: return launch_prepare_d<0, 32>(a);
      case 64: return launch_prepare_d<0, 64>(a);
      case 128: return launch_prepare_d<0, 128>(a);
    }
  } else if (a.bf16) {
    switch (a.d_pad) {
      case 32: return launch_prepare_d<2, 32>(a);
      case 64: return launch_prepare_d<2, 64>(a);
      case 128: return launch_prepare_d<2, 128>(a);
    }
  } else {
    switch (a.d_pad) {
      case 32: return launch_prepare_d<1, 32>(a);
      case 64: return launch_prepare_d<1, 64>(a);
      case 128: return launch_prepare_d<1, 128>(a);
    }
  }
  return cudaErrorInvalidValue;
}

cudaError_t launch_quantize_blocks(const float* X, int B, int N, int H, int d, int block_rows,
                                   int8_t* q, float* scales, cudaStream_t stream) {
  const int nblk = (N + block_rows - 1) / block_rows;
  const long long total = (long long)B * H * nblk;
  const int warps = 8;
  const unsigned grid = (unsigned)((total + warps - 1) / warps);
  quantize_blocks_kernel<<<grid, warps * 32, 0, stream>>>(X, q, scales, B, N, H, d, block_rows, nblk);
  return cudaGetLastError();
}

cudaError_t launch_quantize_static(const float* X, long long n, float scale, float zp, int8_t* q,
                                   cudaStream_t stream) {
  const int threads = 256;
  long long blocks = (n + threads - 1) / threads;
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (blocks < 1) blocks = 1;
  quantize_static_kernel<<<(unsigned)blocks, threads, 0, stream>>>(X, q, n, scale, zp);
  return cudaGetLastError();
}

}  // namespace qmha
