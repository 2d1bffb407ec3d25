// K1: dense cost-volume KL distillation loss, forward + backward, batched over image pairs.
//
// Replaces the body of calculate_cost_loss (src/finetune_timm_mast3r.py:504-540 "mast3r" variant,
// src/finetune_timm_vggt.py:488-533 "vggt" variant) together with get_masked_patch_cost
// (utils/functions.py:402-422) and kl_divergence_map (utils/losses.py:5-15).
//
// Math (per pair; a = normalize(f1), b = normalize(f2), z_ij = a_i . b_j in [-1, 1]):
//   t~12_ij = max(t12_ij / max(R_i, eps), eps) on kept rows (mask1_i), R_i = sum_j t12_ij; same for 21.
//   KL_12 = (1/N) sum_i m1_i [ A_i - D_i + T_i log L_i ] + (#masked rows) * c_var
//     A_i = sum_j t~ log t~,  T_i = sum_j t~,  D_i = sum_j t~_ij z_ij,  L_i = sum_j exp(z_ij)
//     c_var = N eps log(N eps) / N for "mast3r" (masked rows become uniform 1/N), 0 for "vggt".
//   loss = (KL_12 + KL_21) / 2,  KL_21 the same on columns of z with teacher21 / mask2.
//   dz_ij = (1/2N) [ (r_i + c_j) exp(z_ij) - W_ij ],  r_i = m1_i T_i / L_i, c_j = m2_j T'_j / L'_j,
//           W_ij = m1_i t~12_ij + m2_j t~21_ji
//   df1_i = (sum_j dz_ij b_j - a_i rowdot_i) / |f1_i|,  rowdot_i = sum_j dz_ij z_ij   (normalize backward)
//   df2_j = (sum_i dz_ij a_i - b_j coldot_j) / |f2_j|.
// z needs no running max (bounded logits), so row and column softmax statistics come from the same exp.
//
// Pipeline per group of pairs (forward + backward):
//   1 kl_prep            SIMT   normalise rows -> a, b (bf16), 1/|f|
//   2 kl_teacher_stats   SIMT   R, T, A per teacher row (first read of the teacher)
//   3 tc_gemm<EpiKLStats>       z tiles on tcgen05 -> exp / row + col sums in registers; z kept as fp16
//   4 kl_finalize_stats  SIMT   r, c and the T log L terms
//   5 kl_dz              SIMT   builds the W tile from the teacher (second and last read), dz (bf16), rowdot /
//                               coldot and the D = sum W z term of the loss
//   6 tc_gemm<EpiGradOut> x2    df1 = dz b, df2 = dz^T a with the normalisation backward fused; a, b and dz are read
//                               in place through MN-major operand descriptors (no a^T / b^T / dz^T copies)
// Forward-only calls have no step 5: kl_build_w materialises W^T and the pass-1 epilogue accumulates D from it.
// The fp32 N x N cost volume never exists in memory; only fp16 z / bf16 dz are staged for the gradient GEMMs (see
// DESIGN.md for the TMEM budget argument against a single kernel).
#include <type_traits>

#include "../../include/gd3.h"
#include "common.cuh"
#include "tc_gemm.cuh"

namespace gd3 {
namespace {

constexpr float LOG2E = 1.4426950408889634f;
// Every loss term of a pair is accumulated with double atomics.  Thousands of them on ONE address serialise in
// L2 (~35 ns each), so each pair owns kLossSlots accumulators and kl_write_loss adds them up.
constexpr int kLossSlots = 64;
__device__ __forceinline__ void loss_add(double* loss_acc, int g, unsigned slot, double v) {
  atomicAdd(loss_acc + (int64_t)g * kLossSlots + (slot & (kLossSlots - 1)), v);
}

// ------------------------------------------------------------------------------------------
// 1. feature preparation
// ------------------------------------------------------------------------------------------
template <class T>
__device__ __forceinline__ float ld_feat(const T* p);
template <>
__device__ __forceinline__ float ld_feat<float>(const float* p) { return __ldg(p); }
template <>
__device__ __forceinline__ float ld_feat<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }

// grid: (ceil(N/32), G, 2 images); block 256.  Writes a (G, N, ldc) bf16 and inv (G, N); any element strides.
template <class T>
__global__ void __launch_bounds__(256)
    kl_prep_features(const T* __restrict__ f1, const T* __restrict__ f2, int64_t s1P, int64_t s1N, int64_t s1C,
                     int64_t s2P, int64_t s2N, int64_t s2C, int pair0, int N, int C, int ldc, int ldn,
                     __nv_bfloat16* __restrict__ a, __nv_bfloat16* __restrict__ b, float* __restrict__ inv1,
                     float* __restrict__ inv2) {
  __shared__ float s_inv[32];
  __shared__ float tile[32][33];
  const int img = blockIdx.z, g = blockIdx.y, n0 = blockIdx.x * 32;
  const T* f = (img == 0 ? f1 : f2) + (int64_t)(pair0 + g) * (img == 0 ? s1P : s2P);
  const int64_t sN = img == 0 ? s1N : s2N, sC = img == 0 ? s1C : s2C;
  __nv_bfloat16* o = (img == 0 ? a : b) + (int64_t)g * N * ldc;
  float* inv = (img == 0 ? inv1 : inv2) + (int64_t)g * N;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const bool row_major = sC <= sN;   // pick the coalesced direction of the source
  if (row_major) {
    for (int r = w; r < 32; r += 8) {
      const int n = n0 + r;
      float ss = 0.f;
      if (n < N)
        for (int c = lane; c < C; c += 32) { const float v = ld_feat(f + n * sN + c * sC); ss = fmaf(v, v, ss); }
      ss = warp_sum(ss);
      if (lane == 0) s_inv[r] = 1.f / fmaxf(sqrtf(ss), 1e-12f);
    }
  } else {
    // channel-major view (the MASt3R path hands over N-stride 1, C-stride N): lanes run along n
    float ss = 0.f;
    const int n = n0 + lane;
    if (n < N)
      for (int c = w; c < C; c += 8) { const float v = ld_feat(f + n * sN + c * sC); ss = fmaf(v, v, ss); }
    tile[w][lane] = ss;
    __syncthreads();
    if (w == 0) {
      float t = 0.f;
      for (int k = 0; k < 8; ++k) t += tile[k][lane];
      s_inv[lane] = 1.f / fmaxf(sqrtf(t), 1e-12f);
    }
  }
  __syncthreads();
  if (threadIdx.x < 32 && n0 + threadIdx.x < N) inv[n0 + threadIdx.x] = s_inv[threadIdx.x];
  for (int c0 = 0; c0 < C; c0 += 32) {
    __syncthreads();
    // load a 32 (n) x 32 (c) tile in the source's coalesced direction
    for (int r = w; r < 32; r += 8) {
      const int n = row_major ? n0 + r : n0 + lane;
      const int c = row_major ? c0 + lane : c0 + r;
      float v = 0.f;
      if (n < N && c < C) v = ld_feat(f + n * sN + c * sC) * s_inv[n - n0];
      if (row_major) tile[r][lane] = v; else tile[lane][r] = v;
    }
    __syncthreads();
    for (int r = w; r < 32; r += 8) {
      // a[n][c]: lanes along c
      if (n0 + r < N && c0 + lane < C) o[(int64_t)(n0 + r) * ldc + c0 + lane] = __float2bfloat16(tile[r][lane]);
    }
  }
}

// Fast path of step 1 for channel-contiguous features (sC == 1, C % 8 == 0, 16-byte aligned rows): 128-bit accesses.
template <class T>
__device__ __forceinline__ void ld8(const T* p, float (&v)[8]);
template <>
__device__ __forceinline__ void ld8<float>(const float* p, float (&v)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
template <>
__device__ __forceinline__ void ld8<__nv_bfloat16>(const __nv_bfloat16* p, float (&v)[8]) {
  const uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    v[2 * i] = bf16_bits_to_float(w[i] & 0xFFFFu);
    v[2 * i + 1] = bf16_bits_to_float(w[i] >> 16);
  }
}

// One warp per token row, 8 rows per CTA; a row of up to 256 * NIT channels is read once into registers
// (128-bit loads), normalised and written back as bf16 (128-bit stores).  Longer rows are read a second time.
template <class T, int NIT>
__global__ void __launch_bounds__(256)
    kl_prep_fast(const T* __restrict__ f1, const T* __restrict__ f2, int64_t s1P, int64_t s1N, int64_t s2P, int64_t s2N,
                 int pair0, int N, int C, int ldc, __nv_bfloat16* __restrict__ a, __nv_bfloat16* __restrict__ b,
                 float* __restrict__ inv1, float* __restrict__ inv2) {
  const int img = blockIdx.z, g = blockIdx.y;
  const int lane = threadIdx.x & 31;
  const int n = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (n >= N) return;
  const T* row = (img == 0 ? f1 : f2) + (int64_t)(pair0 + g) * (img == 0 ? s1P : s2P) + (int64_t)n * (img == 0 ? s1N : s2N);
  __nv_bfloat16* o = (img == 0 ? a : b) + ((int64_t)g * N + n) * ldc;
  float v[NIT][8];
  float ss = 0.f;
#pragma unroll
  for (int it = 0; it < NIT; ++it) {
    const int c = lane * 8 + 256 * it;
    if (c < C) {
      ld8(row + c, v[it]);
#pragma unroll
      for (int i = 0; i < 8; ++i) ss = fmaf(v[it][i], v[it][i], ss);
    }
  }
  for (int c = lane * 8 + 256 * NIT; c < C; c += 256) {       // rows longer than the register budget
    float t[8];
    ld8(row + c, t);
#pragma unroll
    for (int i = 0; i < 8; ++i) ss = fmaf(t[i], t[i], ss);
  }
  ss = warp_sum(ss);
  const float iv = 1.f / fmaxf(sqrtf(ss), 1e-12f);
  if (lane == 0) (img == 0 ? inv1 : inv2)[(int64_t)g * N + n] = iv;
  auto store8 = [&](int c, const float (&x)[8]) {
    *reinterpret_cast<uint4*>(o + c) = make_uint4(pack_bf16x2(x[0] * iv, x[1] * iv), pack_bf16x2(x[2] * iv, x[3] * iv),
                                                  pack_bf16x2(x[4] * iv, x[5] * iv), pack_bf16x2(x[6] * iv, x[7] * iv));
  };
#pragma unroll
  for (int it = 0; it < NIT; ++it) {
    const int c = lane * 8 + 256 * it;
    if (c < C) store8(c, v[it]);
  }
  for (int c = lane * 8 + 256 * NIT; c < C; c += 256) {
    float t[8];
    ld8(row + c, t);
    store8(c, t);
  }
}

// ------------------------------------------------------------------------------------------
// 2. teacher row statistics: one warp per (direction, pair, row)
// ------------------------------------------------------------------------------------------
// Teacher element types.  fp32 is the reference's; fp16 is the packed form the teacher-side producers emit
// (gd3_teacher_pack: values times a power-of-two `scale` so that small probabilities stay normal numbers, plus the
// row statistics): it halves the bytes of the largest input of the step.  A common scale cancels in the row
// normalisation c / max(sum c, eps); only the clamp of the row sum has to know it (eps * scale).
__device__ __forceinline__ float t_load(const float* p) { return __ldg(p); }
__device__ __forceinline__ float t_load(const __half* p) { return __half2float(__ldg(p)); }
__device__ __forceinline__ float4 t_load4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 t_load4(const __half* p) {
  const uint2 u = __ldg(reinterpret_cast<const uint2*>(p));
  const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&u.x));
  const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
  return make_float4(a.x, a.y, b.x, b.y);
}

template <class TT>
__global__ void __launch_bounds__(256)
    kl_teacher_stats(const TT* __restrict__ t12, const TT* __restrict__ t21, int64_t t_pair_stride,
                     int64_t t_row_stride, const uint8_t* __restrict__ m1, const uint8_t* __restrict__ m2, int pair0,
                     int G, int N, float eps, float eps_sum, float masked_row_const, float* __restrict__ invR /*(2,G,N)*/,
                     float* __restrict__ epsm /*(2,G,N)*/, float* __restrict__ Tsum /*(2,G,N)*/,
                     double* __restrict__ loss_acc /*(G)*/) {
  const int lane = threadIdx.x & 31;
  const int64_t wid = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (wid >= (int64_t)2 * G * N) return;
  const int dir = (int)(wid / ((int64_t)G * N));
  const int rem = (int)(wid - (int64_t)dir * G * N);
  const int g = rem / N, i = rem - g * N;
  const uint8_t keep = (dir == 0 ? m1 : m2)[(int64_t)(pair0 + g) * N + i];
  const int64_t o = ((int64_t)dir * G + g) * N + i;
  const double scale = 0.5 / (double)N;
  if (!keep) {
    if (lane == 0) {
      invR[o] = 0.f;
      epsm[o] = 0.f;
      Tsum[o] = 0.f;
      if (masked_row_const != 0.f) loss_add(loss_acc, g, i, scale * (double)masked_row_const);
    }
    return;
  }
  const TT* row = (dir == 0 ? t12 : t21) + (int64_t)(pair0 + g) * t_pair_stride + (int64_t)i * t_row_stride;
  float R = 0.f;
  for (int j = lane; j < N; j += 32) R += t_load(row + j);
  R = warp_sum(R);
  const float ir = 1.f / fmaxf(R, eps_sum);
  float T = 0.f, A = 0.f;
  for (int j = lane; j < N; j += 32) {
    const float tt = fmaxf(t_load(row + j) * ir, eps);
    T += tt;
    A = fmaf(tt, __logf(tt), A);
  }
  T = warp_sum(T);
  A = warp_sum(A);
  if (lane == 0) {
    invR[o] = ir;
    epsm[o] = eps;
    Tsum[o] = T;
    loss_add(loss_acc, g, i, scale * (double)A);
  }
}

// Single-read variant for N <= 128 * NV: the whole row sits in registers (4 NV floats per lane) and all loads of a
// row are in flight at once.  VEC: N % 4 == 0 and 16-byte aligned rows (128-bit loads); otherwise scalar loads
// (ragged N such as 37^2, whose rows are not 16-byte aligned).
template <int NV, bool VEC, class TT>
__global__ void __launch_bounds__(256)
    kl_teacher_stats_vec(const TT* __restrict__ t12, const TT* __restrict__ t21, int64_t t_pair_stride,
                         int64_t t_row_stride, const uint8_t* __restrict__ m1, const uint8_t* __restrict__ m2, int pair0,
                         int G, int N, float eps, float eps_sum, float masked_row_const, float* __restrict__ invR,
                         float* __restrict__ epsm, float* __restrict__ Tsum, double* __restrict__ loss_acc) {
  const int lane = threadIdx.x & 31;
  const int64_t wid = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (wid >= (int64_t)2 * G * N) return;
  const int dir = (int)(wid / ((int64_t)G * N));
  const int rem = (int)(wid - (int64_t)dir * G * N);
  const int g = rem / N, i = rem - g * N;
  const uint8_t keep = (dir == 0 ? m1 : m2)[(int64_t)(pair0 + g) * N + i];
  const int64_t o = ((int64_t)dir * G + g) * N + i;
  const double scale = 0.5 / (double)N;
  if (!keep) {
    if (lane == 0) {
      invR[o] = 0.f;
      epsm[o] = 0.f;
      Tsum[o] = 0.f;
      if (masked_row_const != 0.f) loss_add(loss_acc, g, i, scale * (double)masked_row_const);
    }
    return;
  }
  const TT* rowp = (dir == 0 ? t12 : t21) + (int64_t)(pair0 + g) * t_pair_stride + (int64_t)i * t_row_stride;
  float v[4 * NV];       // VEC: element 4 * (lane + 32 k) + q at v[4 k + q]; scalar: element lane + 32 k at v[k]
  if (VEC) {
    // rows are 4-element aligned and (for ragged N) padded to a multiple of 4: the last vector may reach into the
    // padding, whose elements are masked below
    const int nv = (N + 3) >> 2;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int j = lane + 32 * k;
      float4 t4 = (j < nv) ? t_load4(rowp + 4 * j) : make_float4(0.f, 0.f, 0.f, 0.f);
      if (4 * j + 1 >= N) t4.y = 0.f;
      if (4 * j + 2 >= N) t4.z = 0.f;
      if (4 * j + 3 >= N) t4.w = 0.f;
      v[4 * k] = t4.x; v[4 * k + 1] = t4.y; v[4 * k + 2] = t4.z; v[4 * k + 3] = t4.w;
    }
  } else {
#pragma unroll
    for (int k = 0; k < 4 * NV; ++k) {
      const int j = lane + 32 * k;
      v[k] = (j < N) ? t_load(rowp + j) : 0.f;
    }
  }
  float R = 0.f;
#pragma unroll
  for (int k = 0; k < NV; ++k) R += (v[4 * k] + v[4 * k + 1]) + (v[4 * k + 2] + v[4 * k + 3]);
  R = warp_sum(R);
  const float ir = 1.f / fmaxf(R, eps_sum);
  float T = 0.f, A = 0.f;
#pragma unroll
  for (int k = 0; k < 4 * NV; ++k) {
    const bool in_row = VEC ? (4 * (lane + 32 * (k >> 2)) + (k & 3) < N) : (lane + 32 * k < N);
    if (in_row) {
      const float tt = fmaxf(v[k] * ir, eps);
      T += tt;
      A = fmaf(tt, __log2f(tt), A);      // in bits; converted to nats once per row below
    }
  }
  A *= 0.69314718055994531f;
  T = warp_sum(T);
  A = warp_sum(A);
  if (lane == 0) {
    invR[o] = ir;
    epsm[o] = eps;
    Tsum[o] = T;
    loss_add(loss_acc, g, i, scale * (double)A);
  }
}

// Row statistics supplied by the teacher-side producer (gd3_teacher_pack): stats (P, 3, N) = [row sum R of the
// unscaled teacher | T = sum_j t~ | A = sum_j t~ ln t~].  One thread per (direction, pair, row); replaces the pass over
// the volume above.
__global__ void kl_stats_from_input(const float* __restrict__ s12, const float* __restrict__ s21,
                                    const uint8_t* __restrict__ m1, const uint8_t* __restrict__ m2, int pair0, int G, int N,
                                    float eps, float scale, float masked_row_const, float* __restrict__ invR,
                                    float* __restrict__ epsm, float* __restrict__ Tsum, double* __restrict__ loss_acc) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)2 * G * N) return;
  const int dir = (int)(idx / ((int64_t)G * N));
  const int rem = (int)(idx - (int64_t)dir * G * N);
  const int g = rem / N, i = rem - g * N;
  const uint8_t keep = (dir == 0 ? m1 : m2)[(int64_t)(pair0 + g) * N + i];
  const double lscale = 0.5 / (double)N;
  if (!keep) {
    invR[idx] = 0.f;
    epsm[idx] = 0.f;
    Tsum[idx] = 0.f;
    if (masked_row_const != 0.f) loss_add(loss_acc, g, i, lscale * (double)masked_row_const);
    return;
  }
  const float* st = (dir == 0 ? s12 : s21) + (int64_t)(pair0 + g) * 3 * N + i;
  invR[idx] = 1.f / (fmaxf(st[0], eps) * scale);      // applied to the stored (scaled) values
  epsm[idx] = eps;
  Tsum[idx] = st[N];
  loss_add(loss_acc, g, i, lscale * (double)st[2 * N]);
}

// Producer-side packing of a teacher volume (fp32 -> fp16 * scale + row statistics), one warp per row, the row read
// once (kept in registers for N <= 2048, re-read from L1 / L2 otherwise).
template <int NV>
__global__ void __launch_bounds__(256)
    teacher_pack_rows(const float* __restrict__ t, int64_t pair_stride, int64_t row_stride, int64_t rows_total, int N,
                      float eps, float scale, __half* __restrict__ out, int64_t out_pair_stride, int64_t out_row_stride,
                      float* __restrict__ stats) {
  const int lane = threadIdx.x & 31;
  const int64_t wid = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (wid >= rows_total) return;
  const int64_t p = wid / N;
  const int i = (int)(wid - p * N);
  const float* row = t + p * pair_stride + (int64_t)i * row_stride;
  __half* orow = out + p * out_pair_stride + (int64_t)i * out_row_stride;
  float v[NV > 0 ? 4 * NV : 1];
  float R = 0.f;
  if (NV > 0) {
#pragma unroll
    for (int k = 0; k < 4 * NV; ++k) {
      const int j = lane + 32 * k;
      v[k] = (j < N) ? __ldg(row + j) : 0.f;
      R += v[k];
    }
  } else {
    for (int j = lane; j < N; j += 32) R += __ldg(row + j);
  }
  R = warp_sum(R);
  const float ir = 1.f / fmaxf(R, eps);
  float T = 0.f, A = 0.f;
  if (NV > 0) {
#pragma unroll
    for (int k = 0; k < 4 * NV; ++k) {
      const int j = lane + 32 * k;
      if (j < N) {
        const float tt = fmaxf(v[k] * ir, eps);
        T += tt;
        A = fmaf(tt, __log2f(tt), A);
        orow[j] = __float2half_rn(v[k] * scale);
      } else if (j < out_row_stride) {
        orow[j] = __float2half_rn(0.f);      // row padding (rows are padded to 16 bytes for vector loads)
      }
    }
  } else {
    for (int j = lane; j < N; j += 32) {
      const float c = __ldg(row + j);
      const float tt = fmaxf(c * ir, eps);
      T += tt;
      A = fmaf(tt, __log2f(tt), A);
      orow[j] = __float2half_rn(c * scale);
    }
    for (int j = N + lane; j < out_row_stride; j += 32) orow[j] = __float2half_rn(0.f);
  }
  T = warp_sum(T);
  A = warp_sum(A) * 0.69314718055994531f;
  if (lane == 0) {
    float* st = stats + p * 3 * N + i;
    st[0] = R;
    st[N] = T;
    st[2 * N] = A;
  }
}

// (step 3, W^T[g][j][i] = m1_i t~12_ij + m2_j t~21_ji, is kl_build_w_fast below)

// ------------------------------------------------------------------------------------------
// 4. pass-1 epilogue: softmax statistics + D term straight from the TMEM accumulator
// ------------------------------------------------------------------------------------------
// WITH_D: also accumulate the D term  -1/(2N) sum_ij W_ij z_ij  (forward-only calls).  When the backward runs, kl_dz
// reads W and z anyway and takes the term over, so this epilogue issues no global loads at all.
template <bool WITH_D>
struct EpiKLStats {
  // the fp16 copy of z leaves through TMA stores (one 32 x 32 box per warp and chunk, see EpiGradOutTMA)
  static constexpr int kScratchBytes = tc::kMaxEpiWarps * tc::kStoreSlabBytes;
  struct Params {
    alignas(64) CUtensorMap tm_z;   // (ldz, N, g) fp16 view of the z staging buffer (valid when Z != nullptr)
    int N;
    const float* WT;   // (G, N, ldw)
    int ldw;
    float* Lrow;       // (G, N)  sum_j exp(z_ij)
    float* Lcol;       // (G, N)  sum_i exp(z_ij)
    double* loss_acc;  // (G)
    __half* Z;         // (G, N, ldz) fp16 staging of z for the backward, or nullptr (forward only)
    int ldz;
  };
  struct Pre {};
  __device__ static void pre(const Params&, const tc::EpiCtx&, Pre&) {}
  __device__ static void run(const Params& p, const tc::EpiCtx& cx, const Pre&) {
    const int i = cx.m0 + cx.row;
    const bool row_ok = i < p.N;
    const float* wt = p.WT + (int64_t)cx.b * p.N * p.ldw + i;
    float rowsum = 0.f, dsum = 0.f;
    for (int c = cx.col_begin; c < cx.col_end; c += 32) {
      const int j0 = cx.n0 + c;
      if (j0 >= p.N) break;                       // warp-uniform: the rest of the tile is padding
      float v[32];
      tc::tmem_ld32(cx.tmem + c, v);
      if (p.Z && cx.m0 + (cx.row & ~31) < p.N) {          // warp-uniform: some row of this warp lies inside the tensor
        uint8_t* slab = cx.scratch + cx.epi_warp * tc::kStoreSlabBytes;
        if (cx.lane == 0) tc::tma_store_wait_read();        // the previous box has been read out of the slab
        __syncwarp();
#pragma unroll
        for (int q = 0; q < 4; ++q)
          *reinterpret_cast<uint4*>(slab + tc::store_slab_offset(cx.lane, q)) =
              make_uint4(pack_f16x2(v[8 * q], v[8 * q + 1]), pack_f16x2(v[8 * q + 2], v[8 * q + 3]),
                         pack_f16x2(v[8 * q + 4], v[8 * q + 5]), pack_f16x2(v[8 * q + 6], v[8 * q + 7]));
        tc::fence_proxy_async_smem();
        __syncwarp();
        if (cx.lane == 0) {
          tc::tma_store_3d(&p.tm_z, slab, j0, cx.m0 + (cx.row & ~31), cx.b);     // columns >= ldz / rows >= N are clipped
          tc::tma_store_commit();
        }
      }
      float e[32];
#pragma unroll
      for (int q = 0; q < 32; ++q) {
        const bool ok = row_ok && (j0 + q < p.N);
        const float ex = ok ? exp2f(v[q] * LOG2E) : 0.f;
        if (WITH_D) {
          const float w = ok ? __ldg(wt + (int64_t)(j0 + q) * p.ldw) : 0.f;
          dsum = fmaf(w, v[q], dsum);
        }
        rowsum += ex;
        e[q] = ex;
      }
      // transpose-reduce: lane q ends with sum over the warp's 32 rows of column j0 + q
#pragma unroll
      for (int off = 16; off >= 1; off >>= 1) {
        const bool up = (cx.lane & off) != 0;
#pragma unroll
        for (int k = 0; k < off; ++k) {
          const float send = up ? e[k] : e[k + off];
          const float keep = up ? e[k + off] : e[k];
          e[k] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
      }
      if (j0 + cx.lane < p.N) atomicAdd(p.Lcol + (int64_t)cx.b * p.N + j0 + cx.lane, e[0]);
    }
    if (row_ok) atomicAdd(p.Lrow + (int64_t)cx.b * p.N + i, rowsum);
    if (p.Z) {
      if (cx.lane == 0) tc::tma_store_wait_read();
      __syncwarp();
    }
    if (WITH_D) {
      dsum = warp_sum(dsum);
      if (cx.lane == 0 && dsum != 0.f) loss_add(p.loss_acc, cx.b, (unsigned)i >> 5, -(0.5 / (double)p.N) * (double)dsum);
    }
  }
};

// ------------------------------------------------------------------------------------------
// 5. r_i, c_j and the T log L terms.  One thread per (direction, g, i).
// ------------------------------------------------------------------------------------------
__global__ void kl_finalize_stats(int G, int N, const float* __restrict__ invR, const float* __restrict__ Tsum,
                                  const float* __restrict__ Lrow, const float* __restrict__ Lcol,
                                  float* __restrict__ rc /*(2,G,N): r then c*/, double* __restrict__ loss_acc) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  double contrib = 0.0;
  int g = 0;
  if (idx < (int64_t)2 * G * N) {
    const int dir = (int)(idx / ((int64_t)G * N));
    const int rem = (int)(idx - (int64_t)dir * G * N);
    g = rem / N;
    const float T = Tsum[idx];
    const bool keep = invR[idx] != 0.f;
    const float L = (dir == 0 ? Lrow : Lcol)[rem];
    rc[idx] = keep ? T / L : 0.f;
    if (keep) contrib = (0.5 / (double)N) * (double)T * (double)logf(L);
  }
  // threads of a warp may straddle two pairs: reduce only when uniform, else add individually
  const unsigned full = 0xffffffffu;
  const int g0 = __shfl_sync(full, g, 0);
  const bool uniform = __all_sync(full, g == g0);
  if (uniform) {
    contrib = warp_sum(contrib);
    if ((threadIdx.x & 31) == 0 && contrib != 0.0) loss_add(loss_acc, g0, (unsigned)(idx >> 5), contrib);
  } else if (contrib != 0.0) {
    loss_add(loss_acc, g, (unsigned)idx, contrib);
  }
}

// ------------------------------------------------------------------------------------------
// Steps 3 and 6: 64 x 64 tiles, 128-bit global accesses wherever rows are 16-byte aligned (always for the padded
// workspace buffers), padded shared-memory transposes.
//   3. W^T[g][j][i] = m1_i t~12_ij + m2_j t~21_ji
//   6. dz = s ((r_i + c_j) exp(z) - W), written as dz (row i) and dz^T (row j) in bf16, + row / col dots + sum W z
// ------------------------------------------------------------------------------------------
// grid (N/64 i-tiles, N/64 j-tiles, G), block 256: thread = (row tr of a 16-row pass, float4 column tc)
// W^T tile [j][i] (64 x 64, row stride 65) of pair g into shared memory: t~12 read with lanes along j (its contiguous
// direction) and transposed on the way in, t~21 read with lanes along i and added.  256 threads; ends with a barrier.
// VEC: teacher rows are 16-byte aligned (128-bit loads); otherwise four scalar loads per thread (ragged N such as
// 37^2: still 256 contiguous bytes per 16 threads).
template <class TT>
struct TeacherView {
  const TT *t12, *t21;              // already offset to the pair
  int64_t row_stride;
  const float *ir12, *ir21;         // 1 / row sum (0 for masked rows), already offset to the pair
  const float *e12, *e21;           // eps of kept rows (0 for masked rows)
};
// Four teacher values of one thread.  Aligned rows (VEC): columns col .. col + 3 as one 128-bit load.  Ragged rows
// (N % 4 != 0, e.g. 37^2): columns col, col + 16, col + 32, col + 48, so that the 16 lanes of a row read 64 contiguous
// bytes per instruction instead of 16-byte-strided words.
template <bool VEC, class TT>
__device__ __forceinline__ float4 teacher_ld4(const TT* src, int col, int N) {
  if (VEC) return t_load4(src);
  float4 v;
  v.x = t_load(src);
  v.y = (col + 16 < N) ? t_load(src + 16) : 0.f;
  v.z = (col + 32 < N) ? t_load(src + 32) : 0.f;
  v.w = (col + 48 < N) ? t_load(src + 48) : 0.f;
  return v;
}
// fp16 teacher, aligned rows: 8 halves (one 128-bit load) per thread and pass, two passes of 32 rows -- the same bytes
// per load instruction as the fp32 path, so the tile stays latency-hidden at half the traffic.
__device__ __forceinline__ void h8_to_float(const uint4& u, float (&f)[8]) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float2 p = __half22float2(*reinterpret_cast<const __half2*>(&w[k]));
    f[2 * k] = p.x;
    f[2 * k + 1] = p.y;
  }
}
__device__ __forceinline__ void load_w_tile_h8(float (*ws)[65], const TeacherView<__half>& tv, int N, int i0, int j0) {
  const int tr = threadIdx.x >> 3, tc = (threadIdx.x & 7) * 8;
  uint4 v12[2], v21[2];
  float ir12[2], ee12[2], ir21[2], ee21[2];
#pragma unroll
  for (int ps = 0; ps < 2; ++ps) {
    const int r = ps * 32 + tr;
    const int i = i0 + r, j = j0 + tc;
    const bool ok12 = i < N && j < N;
    v12[ps] = ok12 ? __ldg(reinterpret_cast<const uint4*>(tv.t12 + (int64_t)i * tv.row_stride + j)) : make_uint4(0u, 0u, 0u, 0u);
    ir12[ps] = ok12 ? tv.ir12[i] : 0.f;
    ee12[ps] = ok12 ? tv.e12[i] : 0.f;
    const int jj = j0 + r, ii = i0 + tc;
    const bool ok21 = jj < N && ii < N;
    v21[ps] = ok21 ? __ldg(reinterpret_cast<const uint4*>(tv.t21 + (int64_t)jj * tv.row_stride + ii)) : make_uint4(0u, 0u, 0u, 0u);
    ir21[ps] = ok21 ? tv.ir21[jj] : 0.f;
    ee21[ps] = ok21 ? tv.e21[jj] : 0.f;
  }
#pragma unroll
  for (int ps = 0; ps < 2; ++ps) {
    const int r = ps * 32 + tr;
    float f[8];
    h8_to_float(v12[ps], f);
#pragma unroll
    for (int q = 0; q < 8; ++q) ws[tc + q][r] = fmaxf(f[q] * ir12[ps], ee12[ps]);
  }
  __syncthreads();
#pragma unroll
  for (int ps = 0; ps < 2; ++ps) {
    const int r = ps * 32 + tr;
    float f[8];
    h8_to_float(v21[ps], f);
#pragma unroll
    for (int q = 0; q < 8; ++q) ws[r][tc + q] += fmaxf(f[q] * ir21[ps], ee21[ps]);
  }
  __syncthreads();
}
template <bool VEC, class TT>
__device__ __forceinline__ void load_w_tile(float (*ws)[65], const TeacherView<TT>& tv, int N, int i0, int j0) {
  if constexpr (VEC && std::is_same<TT, __half>::value) {
    load_w_tile_h8(ws, tv, N, i0, j0);
    return;
  }
  const int tr = threadIdx.x >> 4;
  // tile column of a thread's element q: tc + q * cs
  const int tc = VEC ? (threadIdx.x & 15) * 4 : (threadIdx.x & 15);
  constexpr int cs = VEC ? 1 : 16;
  // all eight teacher loads of a thread (and their row factors) are issued before the first use
  float4 v12[4], v21[4];
  float ir12[4], ee12[4], ir21[4], ee21[4];
#pragma unroll
  for (int ps = 0; ps < 4; ++ps) {
    const int r = ps * 16 + tr;
    const int i = i0 + r, j = j0 + tc;
    const bool ok12 = i < N && j < N;
    v12[ps] = ok12 ? teacher_ld4<VEC, TT>(tv.t12 + (int64_t)i * tv.row_stride + j, j, N) : make_float4(0.f, 0.f, 0.f, 0.f);
    ir12[ps] = ok12 ? tv.ir12[i] : 0.f;
    ee12[ps] = ok12 ? tv.e12[i] : 0.f;
    const int jj = j0 + r, ii = i0 + tc;
    const bool ok21 = jj < N && ii < N;
    v21[ps] = ok21 ? teacher_ld4<VEC, TT>(tv.t21 + (int64_t)jj * tv.row_stride + ii, ii, N) : make_float4(0.f, 0.f, 0.f, 0.f);
    ir21[ps] = ok21 ? tv.ir21[jj] : 0.f;
    ee21[ps] = ok21 ? tv.e21[jj] : 0.f;
  }
#pragma unroll
  for (int ps = 0; ps < 4; ++ps) {
    const int r = ps * 16 + tr;
    const float ir = ir12[ps], ee = ee12[ps];
    ws[tc][r] = fmaxf(v12[ps].x * ir, ee);
    ws[tc + cs][r] = fmaxf(v12[ps].y * ir, ee);
    ws[tc + 2 * cs][r] = fmaxf(v12[ps].z * ir, ee);
    ws[tc + 3 * cs][r] = fmaxf(v12[ps].w * ir, ee);
  }
  __syncthreads();
#pragma unroll
  for (int ps = 0; ps < 4; ++ps) {
    const int r = ps * 16 + tr;
    const float ir = ir21[ps], ee = ee21[ps];
    ws[r][tc] += fmaxf(v21[ps].x * ir, ee);
    ws[r][tc + cs] += fmaxf(v21[ps].y * ir, ee);
    ws[r][tc + 2 * cs] += fmaxf(v21[ps].z * ir, ee);
    ws[r][tc + 3 * cs] += fmaxf(v21[ps].w * ir, ee);
  }
  __syncthreads();
}
template <class TT>
__device__ __forceinline__ TeacherView<TT> teacher_view(const TT* t12, const TT* t21, int64_t t_pair_stride,
                                                        int64_t t_row_stride, int pair, int g, int G, int N,
                                                        const float* invR, const float* epsm) {
  TeacherView<TT> tv;
  tv.t12 = t12 + (int64_t)pair * t_pair_stride;
  tv.t21 = t21 + (int64_t)pair * t_pair_stride;
  tv.row_stride = t_row_stride;
  tv.ir12 = invR + (int64_t)g * N;
  tv.ir21 = invR + ((int64_t)G + g) * N;
  tv.e12 = epsm + (int64_t)g * N;
  tv.e21 = epsm + ((int64_t)G + g) * N;
  return tv;
}

// Step 3 as a kernel of its own is only needed by forward-only calls (the pass-1 epilogue then reads W^T for the D
// term); with a backward, kl_dz_fast builds the tile itself and W^T is never written.
// grid (N/64 i-tiles, N/64 j-tiles, G), block 256.  W^T rows are padded to a multiple of 4 floats: 128-bit stores.
template <bool VEC, class TT>
__global__ void __launch_bounds__(256)
    kl_build_w_fast(const TT* __restrict__ t12, const TT* __restrict__ t21, int64_t t_pair_stride,
                    int64_t t_row_stride, int pair0, int G, int N, const float* __restrict__ invR,
                    const float* __restrict__ epsm, float* __restrict__ WT, int ldw) {
  __shared__ float ws[64][65];
  const int g = blockIdx.z, i0 = blockIdx.x * 64, j0 = blockIdx.y * 64;
  load_w_tile<VEC, TT>(ws, teacher_view<TT>(t12, t21, t_pair_stride, t_row_stride, pair0 + g, g, G, N, invR, epsm), N, i0, j0);
  const int tr = threadIdx.x >> 4, tc = (threadIdx.x & 15) * 4;
#pragma unroll
  for (int ps = 0; ps < 4; ++ps) {
    const int r = ps * 16 + tr, j = j0 + r, i = i0 + tc;
    if (j < N && i < N)
      *reinterpret_cast<float4*>(WT + ((int64_t)g * N + j) * ldw + i) =
          make_float4(ws[r][tc], ws[r][tc + 1], ws[r][tc + 2], ws[r][tc + 3]);
  }
}

// grid (N/64 j-tiles, N/64 i-tiles, G), block 256.  The W^T tile comes straight from the teacher volumes (no W^T
// buffer in the backward path).
template <bool VEC, class TT>
__global__ void __launch_bounds__(256)
    kl_dz_fast(int G, int N, float grad_scale, const __half* __restrict__ Z, int ldz, const TT* __restrict__ t12,
               const TT* __restrict__ t21, int64_t t_pair_stride, int64_t t_row_stride, int pair0,
               const float* __restrict__ invR, const float* __restrict__ epsm, const float* __restrict__ rc,
               __nv_bfloat16* __restrict__ dZ, int ldd, float* __restrict__ rowdot, float* __restrict__ coldot,
               double* __restrict__ loss_acc) {
  __shared__ float ws[64][65];      // W^T tile [j][i]
  __shared__ float cdot[32][65];
  __shared__ float red[32];
  float wz = 0.f;                   // sum W z of this tile (the D term of the loss)
  const int g = blockIdx.z, i0 = blockIdx.y * 64, j0 = blockIdx.x * 64;
  // thread = (row tr8 of a 32-row pass, 8 consecutive columns at jc); its two z vectors are requested before the W tile
  // is built so that their latency hides behind it
  const int tr8 = threadIdx.x >> 3, jc = (threadIdx.x & 7) * 8;
  uint4 zpre[2];
#pragma unroll
  for (int ps = 0; ps < 2; ++ps) {
    const int i = i0 + ps * 32 + tr8;
    zpre[ps] = (i < N && j0 + jc < N) ? __ldg(reinterpret_cast<const uint4*>(Z + ((int64_t)g * N + i) * ldz + j0 + jc))
                                      : make_uint4(0u, 0u, 0u, 0u);
  }
  load_w_tile<VEC, TT>(ws, teacher_view<TT>(t12, t21, t_pair_stride, t_row_stride, pair0 + g, g, G, N, invR, epsm), N, i0, j0);
  const float s = grad_scale * 0.5f / (float)N;
  const float* rr = rc + (int64_t)g * N;
  const float* cc = rc + ((int64_t)G + g) * N;
  float cj[8], cdp[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    cj[q] = (j0 + jc + q < N) ? cc[j0 + jc + q] : 0.f;
    cdp[q] = 0.f;
  }
#pragma unroll
  for (int ps = 0; ps < 2; ++ps) {
    const int r = ps * 32 + tr8, i = i0 + r;
    float rd = 0.f;
    float d[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) d[q] = 0.f;
    if (i < N && j0 + jc < N) {
      const uint4 zz = zpre[ps];
      const uint32_t zw[4] = {zz.x, zz.y, zz.z, zz.w};
      const float ri = rr[i];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const __half hv = __ushort_as_half((unsigned short)((q & 1) ? (zw[q >> 1] >> 16) : (zw[q >> 1] & 0xFFFFu)));
        const float z = (j0 + jc + q < N) ? __half2float(hv) : 0.f;
        const float wv = ws[jc + q][r];
        const float v = (j0 + jc + q < N) ? s * ((ri + cj[q]) * exp2f(z * LOG2E) - wv) : 0.f;
        wz = fmaf(wv, z, wz);
        d[q] = v;
        rd = fmaf(v, z, rd);
        cdp[q] = fmaf(v, z, cdp[q]);
      }
      *reinterpret_cast<uint4*>(dZ + ((int64_t)g * N + i) * ldd + j0 + jc) =
          make_uint4(pack_bf16x2(d[0], d[1]), pack_bf16x2(d[2], d[3]), pack_bf16x2(d[4], d[5]), pack_bf16x2(d[6], d[7]));
    }
    // row dot: the 8 threads of a row are consecutive lanes
    rd += __shfl_xor_sync(0xffffffffu, rd, 1);
    rd += __shfl_xor_sync(0xffffffffu, rd, 2);
    rd += __shfl_xor_sync(0xffffffffu, rd, 4);
    if ((threadIdx.x & 7) == 0 && i < N) atomicAdd(rowdot + (int64_t)g * N + i, rd);
  }
#pragma unroll
  for (int q = 0; q < 8; ++q) cdot[tr8][jc + q] = cdp[q];
  __syncthreads();
  if (threadIdx.x < 64) {
    float t = 0.f;
#pragma unroll 8
    for (int k = 0; k < 32; ++k) t += cdot[k][threadIdx.x];
    if (j0 + threadIdx.x < N) atomicAdd(coldot + (int64_t)g * N + j0 + threadIdx.x, t);
  }
  wz = block_sum(wz, red);
  if (threadIdx.x == 0 && wz != 0.f) loss_add(loss_acc, g, blockIdx.x + blockIdx.y * 7u, -(0.5 / (double)N) * (double)wz);
}

// ------------------------------------------------------------------------------------------
// 7. gradient GEMM epilogue: df = (acc - x * dot) * inv_norm, x = normalised feature (bf16)
// ------------------------------------------------------------------------------------------
template <class OutT>
struct EpiGradOut {
  static constexpr bool kOut16 = sizeof(OutT) == 2;
  static constexpr int kWarpBytes = kOut16 ? tc::kWarpTileBytes16 : tc::kWarpTileBytes;
  static constexpr int kScratchBytes = tc::kMaxEpiWarps * kWarpBytes;
  struct Params {
    int N, C;
    const __nv_bfloat16* X;   // (G, N, ldc) normalised features of the image being differentiated
    int ldc;
    const float* dot;         // (G, N)
    const float* inv;         // (G, N)
    OutT* out;                // (P, N, C) contiguous, already offset to this group's first pair
  };
  // x (the normalised feature row of this thread) does not depend on the accumulator: the loads for the
  // first 32 columns are issued before the tile's MMAs are awaited, the following ones one chunk ahead.
  // df goes out through the per-warp shared-memory transposition so that stores cover whole row segments.
  struct Pre {
    uint4 x[4];
    float dot, inv;
  };
  __device__ static __forceinline__ void load_x(const Params& p, const tc::EpiCtx& cx, int i, int c0, uint4 (&x)[4]) {
    const __nv_bfloat16* xr = p.X + ((int64_t)cx.b * p.N + i) * p.ldc + c0;
#pragma unroll
    for (int q = 0; q < 4; ++q) x[q] = __ldg(reinterpret_cast<const uint4*>(xr + 8 * q));
  }
  __device__ static void pre(const Params& p, const tc::EpiCtx& cx, Pre& pr) {
    const int i = cx.m0 + cx.row;
    pr.dot = 0.f;
    pr.inv = 0.f;
#pragma unroll
    for (int q = 0; q < 4; ++q) pr.x[q] = make_uint4(0, 0, 0, 0);
    if (i < p.N) {
      pr.dot = p.dot[(int64_t)cx.b * p.N + i];
      pr.inv = p.inv[(int64_t)cx.b * p.N + i];
      const int c0 = cx.n0 + cx.col_begin;
      if ((p.C % 8 == 0) && c0 + 32 <= p.C) load_x(p, cx, i, c0, pr.x);
    }
  }
  __device__ static void run(const Params& p, const tc::EpiCtx& cx, const Pre& pr) {
    uint8_t* t = cx.scratch + cx.epi_warp * kWarpBytes;
    const int i = cx.m0 + cx.row;
    const bool row_ok = i < p.N;
    const int m_warp = cx.m0 + (cx.row & ~31);
    const int rows = p.N - m_warp;
    const float dot = pr.dot, inv = pr.inv;
    const __nv_bfloat16* x = p.X + ((int64_t)cx.b * p.N + i) * p.ldc;
    OutT* oslab = p.out + ((int64_t)cx.b * p.N + m_warp) * p.C;
    const bool vec = (p.C % 8 == 0);
    uint4 xn[4] = {pr.x[0], pr.x[1], pr.x[2], pr.x[3]};
    for (int c = cx.col_begin; c < cx.col_end; c += 32) {
      const int c0 = cx.n0 + c;
      if (c0 >= p.C) break;
      uint4 xc[4] = {xn[0], xn[1], xn[2], xn[3]};
      if (row_ok && vec && c + 32 < cx.col_end && c0 + 64 <= p.C) load_x(p, cx, i, c0 + 32, xn);   // next chunk
      float v[32];
      tc::tmem_ld32(cx.tmem + c, v);
      if (rows <= 0) continue;
      if (vec && c0 + 32 <= p.C) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const uint32_t xs[4] = {xc[q].x, xc[q].y, xc[q].z, xc[q].w};
#pragma unroll
          for (int h = 0; h < 4; ++h) {
            v[8 * q + 2 * h] = (v[8 * q + 2 * h] - bf16_bits_to_float(xs[h] & 0xFFFFu) * dot) * inv;
            v[8 * q + 2 * h + 1] = (v[8 * q + 2 * h + 1] - bf16_bits_to_float(xs[h] >> 16) * dot) * inv;
          }
        }
      } else {
#pragma unroll
        for (int q = 0; q < 32; ++q) {
          const float xv = (row_ok && c0 + q < p.C) ? __bfloat162float(x[c0 + q]) : 0.f;
          v[q] = (v[q] - xv * dot) * inv;
        }
      }
      if constexpr (kOut16)
        tc::warp_store_rows_bf16(reinterpret_cast<uint32_t*>(t), v, oslab + c0, p.C, rows, p.C - c0, cx.lane);
      else
        tc::warp_store_rows<OutT>(reinterpret_cast<float*>(t), v, oslab + c0, p.C, rows, p.C - c0, cx.lane);
    }
  }
};

// The same epilogue for bf16 gradients with 16-byte aligned rows, leaving through TMA: a warp packs its 32 x 32 chunk
// into a 64-byte-swizzled shared slab (4 STS.128 per thread, conflict-free) and one elected lane issues a TMA store of
// the box; rows / columns outside the tensor are clipped by the tensor map.  Replaces the shared-memory transposition
// + 16 two-row global stores per chunk of warp_store_rows_bf16: the gradient GEMM at cfg2 is bound by its epilogue
// (MMA 3.3 us per tile at peak, epilogue ~8), so the store path is what its time consists of.
struct EpiGradOutTMA {
  static constexpr int kScratchBytes = tc::kMaxEpiWarps * tc::kStoreSlabBytes;      // 8 x 2 KB, each slab 1024-aligned
  struct Params {
    alignas(64) CUtensorMap tm_out;   // (C, N, g) bf16 view of this group's gradient tensor
    int N, C;
    const __nv_bfloat16* X;   // (G, N, ldc) normalised features of the image being differentiated
    int ldc;
    const float* dot;         // (G, N)
    const float* inv;         // (G, N)
  };
  struct Pre {
    uint4 x[4];
    float dot, inv;
  };
  __device__ static __forceinline__ void load_x(const Params& p, const tc::EpiCtx& cx, int i, int c0, uint4 (&x)[4]) {
    const __nv_bfloat16* xr = p.X + ((int64_t)cx.b * p.N + i) * p.ldc + c0;
#pragma unroll
    for (int q = 0; q < 4; ++q) x[q] = __ldg(reinterpret_cast<const uint4*>(xr + 8 * q));
  }
  __device__ static void pre(const Params& p, const tc::EpiCtx& cx, Pre& pr) {
    const int i = cx.m0 + cx.row;
    pr.dot = 0.f;
    pr.inv = 0.f;
#pragma unroll
    for (int q = 0; q < 4; ++q) pr.x[q] = make_uint4(0, 0, 0, 0);
    if (i < p.N) {
      pr.dot = p.dot[(int64_t)cx.b * p.N + i];
      pr.inv = p.inv[(int64_t)cx.b * p.N + i];
      const int c0 = cx.n0 + cx.col_begin;
      if (c0 + 32 <= p.C) load_x(p, cx, i, c0, pr.x);
    }
  }
  __device__ static void run(const Params& p, const tc::EpiCtx& cx, const Pre& pr) {
    uint8_t* slab = cx.scratch + cx.epi_warp * tc::kStoreSlabBytes;
    const int i = cx.m0 + cx.row;
    const bool row_ok = i < p.N;
    const int m_warp = cx.m0 + (cx.row & ~31);
    const float dot = pr.dot, inv = pr.inv;
    const __nv_bfloat16* x = p.X + ((int64_t)cx.b * p.N + i) * p.ldc;
    uint4 xn[4] = {pr.x[0], pr.x[1], pr.x[2], pr.x[3]};
    for (int c = cx.col_begin; c < cx.col_end; c += 32) {
      const int c0 = cx.n0 + c;
      if (c0 >= p.C) break;
      uint4 xc[4] = {xn[0], xn[1], xn[2], xn[3]};
      const bool full = c0 + 32 <= p.C;          // (C % 8 == 0: a partial chunk still has whole 16-byte pieces)
      if (row_ok && c + 32 < cx.col_end && c0 + 64 <= p.C) load_x(p, cx, i, c0 + 32, xn);   // next chunk
      float v[32];
      tc::tmem_ld32(cx.tmem + c, v);
      if (m_warp >= p.N) continue;               // warp-uniform: the whole 32-row slab lies outside the tensor
      if (!full) {
#pragma unroll
        for (int q = 0; q < 4; ++q)
          xc[q] = (row_ok && c0 + 8 * q < p.C) ? __ldg(reinterpret_cast<const uint4*>(x + c0 + 8 * q)) : make_uint4(0, 0, 0, 0);
      }
      // the previous TMA store must have finished reading the slab before it is overwritten
      if (cx.lane == 0) tc::tma_store_wait_read();
      __syncwarp();
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const uint32_t xs[4] = {xc[q].x, xc[q].y, xc[q].z, xc[q].w};
        uint32_t o[4];
#pragma unroll
        for (int h = 0; h < 4; ++h) {
          const float lo = (v[8 * q + 2 * h] - bf16_bits_to_float(xs[h] & 0xFFFFu) * dot) * inv;
          const float hi = (v[8 * q + 2 * h + 1] - bf16_bits_to_float(xs[h] >> 16) * dot) * inv;
          o[h] = pack_bf16x2(lo, hi);
        }
        *reinterpret_cast<uint4*>(slab + tc::store_slab_offset(cx.lane, q)) = make_uint4(o[0], o[1], o[2], o[3]);
      }
      tc::fence_proxy_async_smem();
      __syncwarp();
      if (cx.lane == 0) {
        tc::tma_store_3d(&p.tm_out, slab, c0, m_warp, cx.b);
        tc::tma_store_commit();
      }
    }
    if (cx.lane == 0) tc::tma_store_wait_read();     // the slab is free again (and stays valid until it has been read)
    __syncwarp();
  }
};

__global__ void kl_write_loss(const double* __restrict__ acc, float* __restrict__ loss, int G) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g < G) {
    double t = 0.0;
#pragma unroll 8
    for (int k = 0; k < kLossSlots; ++k) t += acc[(int64_t)g * kLossSlots + k];
    loss[g] = (float)t;
  }
}

struct KLWorkspace {
  __nv_bfloat16 *a, *b, *dZ;
  __half* Z;
  float *inv1, *inv2, *invR, *epsm, *Tsum, *WT, *Lrow, *Lcol, *rc, *rowdot, *coldot;
  double* loss_acc;
  // the four zero-initialised accumulators are contiguous: [Lrow | Lcol | rowdot | coldot]
  size_t total;
  int ldc, ldn, ldw;
};

KLWorkspace carve_kl(void* base, int64_t G, int64_t N, int64_t C, bool backward) {
  KLWorkspace w{};
  Carver c(base);
  w.ldc = (int)round_up<int64_t>(C, 8);
  w.ldn = (int)round_up<int64_t>(N, 8);
  w.ldw = (int)round_up<int64_t>(N, 4);
  w.a = c.take<__nv_bfloat16>(G * N * w.ldc);
  w.b = c.take<__nv_bfloat16>(G * N * w.ldc);
  w.inv1 = c.take<float>(G * N);
  w.inv2 = c.take<float>(G * N);
  w.invR = c.take<float>(2 * G * N);
  w.epsm = c.take<float>(2 * G * N);
  w.Tsum = c.take<float>(2 * G * N);
  w.WT = c.take<float>(backward ? 0 : G * N * w.ldw);      // only the forward-only pass-1 epilogue reads W^T
  w.Lrow = c.take<float>(4 * G * N);
  w.Lcol = w.Lrow + G * N;
  w.rowdot = w.Lrow + 2 * G * N;
  w.coldot = w.Lrow + 3 * G * N;
  w.rc = c.take<float>(2 * G * N);
  w.loss_acc = c.take<double>(G * kLossSlots);
  w.Z = c.take<__half>(backward ? G * N * w.ldn : 0);
  w.dZ = c.take<__nv_bfloat16>(backward ? G * N * w.ldn : 0);
  w.total = c.total();
  return w;
}

int64_t auto_group(int64_t P, int64_t N, int64_t C) {
  // Measured on B200 (tools/probe_kl.py): launches that cover many pairs beat small L2-resident groups,
  // because every kernel of the pipeline then runs several full waves.  So a group is as large as a
  // ~3 GB workspace allows (per pair: z fp16 + dz bf16, or W^T fp32 for forward-only calls, + a, b bf16).
  const double per_pair = N * (double)N * 4 + 2.0 * N * C * 2;
  int64_t g = (int64_t)(3.0e9 / per_pair);
  if (g < 1) g = 1;
  return g > P ? P : g;
}

template <class TT>
void launch_dz(bool vec_ok, dim3 grid, cudaStream_t stream, int g, int N, float grad_scale, const KLWorkspace& w,
               const TT* t12, const TT* t21, int64_t t_pair_stride, int64_t t_row_stride, int p0) {
  if (vec_ok)
    kl_dz_fast<true, TT><<<grid, 256, 0, stream>>>(g, N, grad_scale, w.Z, w.ldn, t12, t21, t_pair_stride, t_row_stride, p0,
                                                   w.invR, w.epsm, w.rc, w.dZ, w.ldn, w.rowdot, w.coldot, w.loss_acc);
  else
    kl_dz_fast<false, TT><<<grid, 256, 0, stream>>>(g, N, grad_scale, w.Z, w.ldn, t12, t21, t_pair_stride, t_row_stride, p0,
                                                    w.invR, w.epsm, w.rc, w.dZ, w.ldn, w.rowdot, w.coldot, w.loss_acc);
}

// teacher row statistics (computed here, or taken from the producer) and, for forward-only calls, the W^T tile
template <class TT>
int teacher_stage_t(const TT* t12, const TT* t21, int64_t t_pair_stride, int64_t t_row_stride, const float* tstats12,
                    const float* tstats21, float scale, const uint8_t* m1, const uint8_t* m2, int p0, int g, int N,
                    float eps, float masked_const, bool backward, const KLWorkspace& w, bool* vec_ok,
                    cudaStream_t stream) {
  const int64_t warps = 2 * (int64_t)g * N;
  // vector width of the teacher loads: 4 fp32 or 8 fp16 elements (16 bytes either way)
  constexpr int VW = 16 / (int)sizeof(TT);
  const bool rows_aligned = t_row_stride % VW == 0 && t_pair_stride % VW == 0 &&
                            reinterpret_cast<uintptr_t>(t12) % 16 == 0 && reinterpret_cast<uintptr_t>(t21) % 16 == 0;
  if (tstats12) {
    GD3_PROF("kl_stats_from_input", stream);
    kl_stats_from_input<<<(unsigned)ceil_div<int64_t>(warps, 256), 256, 0, stream>>>(
        tstats12, tstats21, m1, m2, p0, g, N, eps, scale, masked_const, w.invR, w.epsm, w.Tsum, w.loss_acc);
  } else {
    GD3_PROF("kl_teacher_stats", stream);
    const unsigned blocks = (unsigned)ceil_div<int64_t>(warps, 8);
    const float eps_sum = eps * scale;
#define GD3_TSTATS(NV, VEC)                                                                                            \
  kl_teacher_stats_vec<NV, VEC, TT><<<blocks, 256, 0, stream>>>(t12, t21, t_pair_stride, t_row_stride, m1, m2, p0, g, N, \
                                                                eps, eps_sum, masked_const, w.invR, w.epsm, w.Tsum,      \
                                                                w.loss_acc)
    // (the scalar-load instantiation of the register-resident kernel measured slower than the two-pass kernel on
    // unaligned rows -- 345 vs 185 us at N = 37^2 -- so ragged N keeps the two-pass kernel)
    const bool v4 = rows_aligned && (N % 4 == 0 || t_row_stride >= round_up<int64_t>(N, 4));     // (the statistics kernel loads 4 elements per lane)
    if (v4 && N <= 512) GD3_TSTATS(4, true);
    else if (v4 && N <= 1024) GD3_TSTATS(8, true);
    else if (v4 && N <= 2048) GD3_TSTATS(16, true);
    else
      kl_teacher_stats<TT><<<blocks, 256, 0, stream>>>(t12, t21, t_pair_stride, t_row_stride, m1, m2, p0, g, N, eps,
                                                       eps_sum, masked_const, w.invR, w.epsm, w.Tsum, w.loss_acc);
#undef GD3_TSTATS
  }
  GD3_CHECK_LAUNCH();
  // 128-bit (fp32) / 64-bit (fp16) teacher loads: rows aligned, and either N % 4 == 0 or rows padded so that the last
  // vector of a row stays inside it (gd3_teacher_pack pads to a multiple of 8; elements beyond N are masked at use)
  *vec_ok = rows_aligned && (N % VW == 0 || t_row_stride >= round_up<int64_t>(N, VW));
  if (!backward) {
    // forward only: the pass-1 epilogue needs W^T for the D term
    dim3 grid((unsigned)ceil_div<int64_t>(N, 64), (unsigned)ceil_div<int64_t>(N, 64), (unsigned)g);
    GD3_PROF("kl_build_w_fast", stream);
    if (*vec_ok)
      kl_build_w_fast<true, TT><<<grid, 256, 0, stream>>>(t12, t21, t_pair_stride, t_row_stride, p0, g, N, w.invR, w.epsm,
                                                          w.WT, w.ldw);
    else
      kl_build_w_fast<false, TT><<<grid, 256, 0, stream>>>(t12, t21, t_pair_stride, t_row_stride, p0, g, N, w.invR, w.epsm,
                                                           w.WT, w.ldw);
    GD3_CHECK_LAUNCH();
  }
  return GD3_OK;
}
int teacher_stage(bool f16, const void* t12, const void* t21, int64_t t_pair_stride, int64_t t_row_stride,
                  const float* tstats12, const float* tstats21, float scale, const uint8_t* m1, const uint8_t* m2, int p0,
                  int g, int N, float eps, float masked_const, bool backward, const KLWorkspace& w, bool* vec_ok,
                  cudaStream_t stream) {
  if (f16)
    return teacher_stage_t(static_cast<const __half*>(t12), static_cast<const __half*>(t21), t_pair_stride, t_row_stride,
                           tstats12, tstats21, scale, m1, m2, p0, g, N, eps, masked_const, backward, w, vec_ok, stream);
  return teacher_stage_t(static_cast<const float*>(t12), static_cast<const float*>(t21), t_pair_stride, t_row_stride,
                         tstats12, tstats21, scale, m1, m2, p0, g, N, eps, masked_const, backward, w, vec_ok, stream);
}

}  // namespace
}  // namespace gd3

using namespace gd3;

extern "C" {

int64_t gd3_cost_kl_group_size(int64_t P, int64_t N, int64_t C, int64_t pairs_per_group) {
  if (P <= 0) return 0;
  if (pairs_per_group > 0) return pairs_per_group > P ? P : pairs_per_group;
  return auto_group(P, N, C);
}

size_t gd3_cost_kl_workspace(int64_t P, int64_t N, int64_t C, int64_t pairs_per_group, int with_backward) {
  if (P <= 0 || N <= 0 || C <= 0) return 0;
  const int64_t G = gd3_cost_kl_group_size(P, N, C, pairs_per_group);
  return carve_kl(nullptr, G, N, C, with_backward != 0).total;
}

int gd3_cost_kl(const void* f1, const void* f2, int dtype, int64_t P, int64_t N, int64_t C, int64_t s1P, int64_t s1N,
                int64_t s1C, int64_t s2P, int64_t s2N, int64_t s2C, const void* t12, const void* t21,
                int teacher_dtype, float teacher_scale, const float* tstats12, const float* tstats21,
                int64_t t_pair_stride, int64_t t_row_stride, const uint8_t* m1, const uint8_t* m2, int variant,
                float eps, float grad_scale, float* loss, void* grad_f1, void* grad_f2, int64_t pairs_per_group,
                void* workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (P == 0) return GD3_OK;
  GD3_REQUIRE(P > 0 && N > 0 && C > 0, "gd3_cost_kl: bad sizes P=%lld N=%lld C=%lld", (long long)P, (long long)N,
              (long long)C);
  GD3_REQUIRE(f1 && f2 && t12 && t21 && m1 && m2 && loss, "gd3_cost_kl: null pointer");
  GD3_REQUIRE(dtype == GD3_DTYPE_F32 || dtype == GD3_DTYPE_BF16, "gd3_cost_kl: bad dtype %d", dtype);
  GD3_REQUIRE(variant == GD3_VARIANT_MAST3R || variant == GD3_VARIANT_VGGT, "gd3_cost_kl: unknown variant %d",
              variant);
  GD3_REQUIRE((grad_f1 == nullptr) == (grad_f2 == nullptr), "gd3_cost_kl: pass both gradients or neither");
  GD3_REQUIRE(eps > 0.f, "gd3_cost_kl: eps must be positive");
  GD3_REQUIRE(teacher_dtype == GD3_DTYPE_F32 || teacher_dtype == GD3_DTYPE_F16,
              "gd3_cost_kl: teacher volumes must be fp32 or fp16 (gd3_teacher_pack), got dtype %d", teacher_dtype);
  GD3_REQUIRE(teacher_scale > 0.f, "gd3_cost_kl: teacher_scale must be positive");
  GD3_REQUIRE((tstats12 == nullptr) == (tstats21 == nullptr), "gd3_cost_kl: pass both teacher statistics or neither");
  const bool t_f16 = teacher_dtype == GD3_DTYPE_F16;
  const bool backward = grad_f1 != nullptr;
  const int64_t G = gd3_cost_kl_group_size(P, N, C, pairs_per_group);
  GD3_REQUIRE(G <= 65535 && ceil_div<int64_t>(N, 32) <= 65535, "gd3_cost_kl: problem too large for one launch grid");
  KLWorkspace w = carve_kl(workspace, G, N, C, backward);
  if (!workspace || workspace_bytes < w.total) {
    set_error("gd3_cost_kl: workspace too small (%zu < %zu)", workspace_bytes, w.total);
    return GD3_ERR_WORKSPACE;
  }
  GD3_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "gd3_cost_kl: workspace must be 256-byte aligned");

  // masked rows: "mast3r" softmaxes a zeroed logit row (uniform 1/N) against a teacher row of eps
  const double ne = (double)N * (double)eps;
  const float masked_const = variant == GD3_VARIANT_MAST3R ? (float)(ne * log(ne)) : 0.f;

  // K-major maps for z = a b^T; for the gradient GEMMs the SAME buffers are read MN-major (tc_gemm.cuh):
  //   df1 = dz b      A = dz  (K-major: rows i, contiguous j)   B = b  as [k = j][mn = c]
  //   df2 = dz^T a    A = dz  as [k = i][mn = j]               B = a  as [k = i][mn = c]
  // so neither a^T / b^T nor dz^T is ever written.
  CUtensorMap tm_a, tm_b, tm_dz, tm_dz_mn, tm_a_mn, tm_b_mn;
  int rc;
  if ((rc = tc::make_tmap_bf16(&tm_a, w.a, C, N, G, w.ldc, N * (int64_t)w.ldc, tc::BM))) return rc;
  if ((rc = tc::make_tmap_bf16(&tm_b, w.b, C, N, G, w.ldc, N * (int64_t)w.ldc, 256))) return rc;
  if (backward) {
    if ((rc = tc::make_tmap_bf16(&tm_dz, w.dZ, N, N, G, w.ldn, N * (int64_t)w.ldn, tc::BM))) return rc;
    if ((rc = tc::make_tmap_bf16(&tm_dz_mn, w.dZ, N, N, G, w.ldn, N * (int64_t)w.ldn, 64))) return rc;
    if ((rc = tc::make_tmap_bf16(&tm_a_mn, w.a, C, N, G, w.ldc, N * (int64_t)w.ldc, 64))) return rc;
    if ((rc = tc::make_tmap_bf16(&tm_b_mn, w.b, C, N, G, w.ldc, N * (int64_t)w.ldc, 64))) return rc;
  }

  for (int64_t p0 = 0; p0 < P; p0 += G) {
    const int g = (int)((P - p0) < G ? (P - p0) : G);
    GD3_CHECK_CUDA(cudaMemsetAsync(w.Lrow, 0, sizeof(float) * 4 * G * N, stream));
    GD3_CHECK_CUDA(cudaMemsetAsync(w.loss_acc, 0, sizeof(double) * G * kLossSlots, stream));
    bool vec_ok = false;     // teacher rows allow 128-bit loads
    {
      const int esz = dtype == GD3_DTYPE_F32 ? 4 : 2;
      const bool fast = s1C == 1 && s2C == 1 && C % 8 == 0 && (s1N * esz) % 16 == 0 && (s2N * esz) % 16 == 0 &&
                        (s1P * esz) % 16 == 0 && (s2P * esz) % 16 == 0 &&
                        reinterpret_cast<uintptr_t>(f1) % 16 == 0 && reinterpret_cast<uintptr_t>(f2) % 16 == 0;
      if (fast) {
        dim3 grid((unsigned)ceil_div<int64_t>(N, 8), (unsigned)g, 2);
        GD3_PROF("kl_prep_fast", stream);
#define GD3_KL_PREP(T, NIT)                                                                                         \
  kl_prep_fast<T, NIT><<<grid, 256, 0, stream>>>(static_cast<const T*>(f1), static_cast<const T*>(f2), s1P, s1N, s2P, \
                                                 s2N, (int)p0, (int)N, (int)C, w.ldc, w.a, w.b, w.inv1, w.inv2)
        if (dtype == GD3_DTYPE_F32) {
          if (C <= 512) GD3_KL_PREP(float, 2);
          else GD3_KL_PREP(float, 4);
        } else {
          if (C <= 512) GD3_KL_PREP(__nv_bfloat16, 2);
          else GD3_KL_PREP(__nv_bfloat16, 4);
        }
#undef GD3_KL_PREP
      } else {
        // generic strides (e.g. the MASt3R path's channel-major view)
        dim3 grid((unsigned)ceil_div<int64_t>(N, 32), (unsigned)g, 2);
        GD3_PROF("kl_prep_features", stream);
        if (dtype == GD3_DTYPE_F32)
          kl_prep_features<float><<<grid, 256, 0, stream>>>(
              static_cast<const float*>(f1), static_cast<const float*>(f2), s1P, s1N, s1C, s2P, s2N, s2C, (int)p0,
              (int)N, (int)C, w.ldc, w.ldn, w.a, w.b, w.inv1, w.inv2);
        else
          kl_prep_features<__nv_bfloat16><<<grid, 256, 0, stream>>>(
              static_cast<const __nv_bfloat16*>(f1), static_cast<const __nv_bfloat16*>(f2), s1P, s1N, s1C, s2P, s2N,
              s2C, (int)p0, (int)N, (int)C, w.ldc, w.ldn, w.a, w.b, w.inv1, w.inv2);
      }
      GD3_CHECK_LAUNCH();
    }
    if ((rc = teacher_stage(t_f16, t12, t21, t_pair_stride, t_row_stride, tstats12, tstats21, teacher_scale, m1, m2, (int)p0,
                            g, (int)N, eps, masked_const, backward, w, &vec_ok, stream)))
      return rc;
    {
      tc::GemmShape s{(int)N, (int)N, (int)C, g};
      if (backward) {
        // the D term is taken over by kl_dz, which reads W and z anyway
        EpiKLStats<false>::Params ep{};
        if ((rc = tc::make_tmap_store16(&ep.tm_z, w.Z, w.ldn, N, g, w.ldn, N * (int64_t)w.ldn, true))) return rc;
        ep.N = (int)N; ep.WT = w.WT; ep.ldw = w.ldw; ep.Lrow = w.Lrow; ep.Lcol = w.Lcol; ep.loss_acc = w.loss_acc;
        ep.Z = w.Z; ep.ldz = w.ldn;
        rc = tc::launch_gemm<256, 8, EpiKLStats<false>>("kl_pass1_gemm", tm_a, tm_b, s, ep, stream);
      } else {
        EpiKLStats<true>::Params ep{};
        ep.N = (int)N; ep.WT = w.WT; ep.ldw = w.ldw; ep.Lrow = w.Lrow; ep.Lcol = w.Lcol; ep.loss_acc = w.loss_acc;
        ep.Z = nullptr; ep.ldz = w.ldn;
        rc = tc::launch_gemm<256, 8, EpiKLStats<true>>("kl_pass1_gemm", tm_a, tm_b, s, ep, stream);
      }
      if (rc) return rc;
    }
    {
      const int64_t n = 2 * (int64_t)g * N;
      {
        GD3_PROF("kl_finalize_stats", stream);
        kl_finalize_stats<<<(unsigned)ceil_div<int64_t>(n, 256), 256, 0, stream>>>(g, (int)N, w.invR, w.Tsum, w.Lrow,
                                                                               w.Lcol, w.rc, w.loss_acc);
      }
      GD3_CHECK_LAUNCH();
    }
    if (!backward) {
      GD3_PROF("kl_write_loss", stream);
      kl_write_loss<<<ceil_div(g, 64), 64, 0, stream>>>(w.loss_acc, loss + p0, g);
      GD3_CHECK_LAUNCH();
    }
    if (backward) {
      dim3 grid((unsigned)ceil_div<int64_t>(N, 64), (unsigned)ceil_div<int64_t>(N, 64), (unsigned)g);
      {
        // workspace rows are padded to a multiple of 8 elements, so the 128-bit kernel also serves ragged N (elements in
        // the padding are masked on read and never consumed: the TMA extents of the gradient GEMMs stop at N)
        GD3_PROF("kl_dz_fast", stream);
        if (t_f16)
          launch_dz(vec_ok, grid, stream, g, (int)N, grad_scale, w, static_cast<const __half*>(t12),
                    static_cast<const __half*>(t21), t_pair_stride, t_row_stride, (int)p0);
        else
          launch_dz(vec_ok, grid, stream, g, (int)N, grad_scale, w, static_cast<const float*>(t12),
                    static_cast<const float*>(t21), t_pair_stride, t_row_stride, (int)p0);
      }
      GD3_CHECK_LAUNCH();
      {
        GD3_PROF("kl_write_loss", stream);
        kl_write_loss<<<ceil_div(g, 64), 64, 0, stream>>>(w.loss_acc, loss + p0, g);
      }
      GD3_CHECK_LAUNCH();
      tc::GemmShape s{(int)N, (int)C, (int)N, g};
      const int bn = tc::pick_tile_n(s);
      auto run_grad = [&](auto tag) -> int {
        using E = EpiGradOut<decltype(tag)>;
        using OutT = decltype(tag);
        typename E::Params e1{(int)N, (int)C, w.a, w.ldc, w.rowdot, w.inv1, static_cast<OutT*>(grad_f1) + p0 * N * C};
        typename E::Params e2{(int)N, (int)C, w.b, w.ldc, w.coldot, w.inv2, static_cast<OutT*>(grad_f2) + p0 * N * C};
        int r;
#define GD3_KL_GRAD(BN)                                                                                             \
  do {                                                                                                              \
    if ((r = tc::launch_gemm<BN, 8, E, false, true>("kl_grad_gemm", tm_dz, tm_b_mn, s, e1, stream))) return r;      \
    return tc::launch_gemm<BN, 8, E, true, true>("kl_grad_gemm", tm_dz_mn, tm_a_mn, s, e2, stream);                 \
  } while (0)
        if (bn == 256) GD3_KL_GRAD(256);
        if (bn == 192) GD3_KL_GRAD(192);
        GD3_KL_GRAD(128);
#undef GD3_KL_GRAD
      };
      // bf16 gradients with 16-byte aligned rows leave through TMA stores (EpiGradOutTMA)
      const bool tma_out = dtype == GD3_DTYPE_BF16 && C % 8 == 0 && reinterpret_cast<uintptr_t>(grad_f1) % 16 == 0 &&
                           reinterpret_cast<uintptr_t>(grad_f2) % 16 == 0;
      if (tma_out) {
        using E = EpiGradOutTMA;
        E::Params e1{}, e2{};
        __nv_bfloat16* o1 = static_cast<__nv_bfloat16*>(grad_f1) + p0 * N * C;
        __nv_bfloat16* o2 = static_cast<__nv_bfloat16*>(grad_f2) + p0 * N * C;
        if ((rc = tc::make_tmap_store16(&e1.tm_out, o1, C, N, g, C, N * C, false))) return rc;
        if ((rc = tc::make_tmap_store16(&e2.tm_out, o2, C, N, g, C, N * C, false))) return rc;
        e1.N = e2.N = (int)N;
        e1.C = e2.C = (int)C;
        e1.X = w.a; e2.X = w.b;
        e1.ldc = e2.ldc = w.ldc;
        e1.dot = w.rowdot; e2.dot = w.coldot;
        e1.inv = w.inv1; e2.inv = w.inv2;
#define GD3_KL_GRAD_TMA(BN)                                                                                        \
  do {                                                                                                             \
    if ((rc = tc::launch_gemm<BN, 8, E, false, true>("kl_grad_gemm", tm_dz, tm_b_mn, s, e1, stream))) return rc;  \
    if ((rc = tc::launch_gemm<BN, 8, E, true, true>("kl_grad_gemm", tm_dz_mn, tm_a_mn, s, e2, stream))) return rc; \
  } while (0)
        if (bn == 256) GD3_KL_GRAD_TMA(256);
        else if (bn == 192) GD3_KL_GRAD_TMA(192);
        else GD3_KL_GRAD_TMA(128);
#undef GD3_KL_GRAD_TMA
      } else {
        if (dtype == GD3_DTYPE_F32) rc = run_grad(float{});
        else rc = run_grad(__nv_bfloat16{});
        if (rc) return rc;
      }
    }
  }
  return GD3_OK;
}

int gd3_teacher_pack(const float* t, int64_t P, int64_t N, int64_t t_pair_stride, int64_t t_row_stride, float eps,
                     float scale, void* out_f16, int64_t out_row_stride, float* stats, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (P == 0 || N == 0) return GD3_OK;
  GD3_REQUIRE(P > 0 && N > 0, "gd3_teacher_pack: bad sizes P=%lld N=%lld", (long long)P, (long long)N);
  GD3_REQUIRE(t && out_f16 && stats, "gd3_teacher_pack: null pointer");
  GD3_REQUIRE(eps > 0.f && scale > 0.f, "gd3_teacher_pack: eps and scale must be positive");
  GD3_REQUIRE(out_row_stride >= N && out_row_stride < N + 32, "gd3_teacher_pack: output row stride %lld must be in [N, N + 32)",
              (long long)out_row_stride);
  const int64_t rows = P * N;
  const unsigned blocks = (unsigned)ceil_div<int64_t>(rows, 8);
  __half* out = static_cast<__half*>(out_f16);
  {
    GD3_PROF("teacher_pack_rows", stream);
#define GD3_TPACK(NV) \
  teacher_pack_rows<NV><<<blocks, 256, 0, stream>>>(t, t_pair_stride, t_row_stride, rows, (int)N, eps, scale, out, N * out_row_stride, out_row_stride, stats)
    if (N <= 512) GD3_TPACK(4);
    else if (N <= 1024) GD3_TPACK(8);
    else if (N <= 2048) GD3_TPACK(16);
    else GD3_TPACK(0);
#undef GD3_TPACK
  }
  GD3_CHECK_LAUNCH();
  return GD3_OK;
}

}  // extern "C"
