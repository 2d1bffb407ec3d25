// K4: relative-depth losses on the depth-difference head, forward + backward, batched over keypoint sets.
//
// Replaces pairwise_logistic_ranking_loss (utils/losses.py:18-41), intra_depth_loss (utils/losses.py:44-69),
// the head DepthAwareFeatureFusion.fusion_layer (+tanh) (utils/model.py:100-105,122-127) evaluated on all
// K^2 feature differences, and the cross-view L1 term of calculate_depth_loss
// (src/finetune_timm_mast3r.py:489-494).
//
// The reference materialises the (K, K, D) difference tensor and runs the MLP on K^2 rows.  The first
// Linear is linear, so  W1 (f_b - f_a) + b1 = u_b - u_a + b1  with u = f W1^T (K x H, H = hidden = 128):
// the D-wide GEMM is hoisted out of the pair loop and runs once on the tensor cores (split-bf16, ~fp32
// accurate); only LayerNorm / GELU / w2 / tanh / loss are evaluated per pair, on chip.
//
// For an ordered pair (a -> b):  h = u_b - u_a + b1,  s = [tanh](w2 . GELU(LN(h)) + b2),  D = d_b - d_a
//   logistic (ranking): valid = |D| > thr,          l = log(1 + exp(-sign(D) s))
//   hinge (intra_depth): valid = |tanh D| > thr,    l = relu(margin - sign(D) s)      [reference pair (i,j) = (b,a)]
//   loss = mean over valid pairs (0 without gradient if none)
// L1 term: pairs keypoint k of set 2p+1 (a) with keypoint k of set 2p (b): mean_k |s - tanh(d_b - d_a)|.
//
// Layout of the pair kernel: 16 lanes own one pair, each lane 8 of the 128 hidden units, so LayerNorm / dot
// reductions are 4 xor-shuffles and a warp works on two pairs at a time.  A CTA owns a 128 x 128 tile of
// (a, b): every warp keeps the gradient of its b rows in registers and accumulates the a side into a shared
// tile; the a index is staggered per warp so no two warps touch the same row in the same step.
#include <type_traits>

#include "../../include/gd3.h"
#include "common.cuh"
#include "tc_gemm.cuh"
#include "split_bf16.cuh"

namespace gd3 {
namespace {

constexpr int H = 128;           // hidden width of fusion_layer (utils/model.py:88)
constexpr int HPL = 8;           // hidden units per lane
// CTA tile of the pair kernel: TILE_A rows on the shared (a) side, WARPS * b_per_warp rows on the register (b)
// side.  Two 8-warp CTAs per SM (independent barrier domains) hide latency better than one 16-warp CTA.
constexpr int TILE_A = 64;
constexpr int WARPS = 8;
constexpr int B_PER_WARP_MAX = 16;                   // b rows a warp walks per CTA: 16, 8, 4 or 2, chosen per problem (rank_b_per_warp)
constexpr int CTAS_PER_SM = 2;
#ifndef GD3_RANK_UNROLL
#define GD3_RANK_UNROLL 4
#endif
constexpr int kRankUnroll = GD3_RANK_UNROLL;      // steps of the a-tile walk per loop iteration (build knob for experiments)
constexpr int kBarrierPeriod = 4;                 // CTA barrier every this many steps of the walk (see rank_pairs)
constexpr int SLOTS = TILE_A / 2;                    // ring of row pairs a warp walks through
constexpr int SPACING = SLOTS / WARPS;               // ring distance between consecutive warps
static_assert(SLOTS % WARPS == 0 && SPACING >= 2, "stagger needs at least 2 ring slots between warps");

// sum over the 16 lanes of a half warp (xor offsets < 16 never cross the halves)
__device__ __forceinline__ float half_sum(float v, unsigned mask) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(mask, v, o);
  return v;
}
// hidden index of this lane's i-th element: two float4 groups, 64 apart -> conflict-free LDS.128
__device__ __forceinline__ int hidx(int l16, int i) { return (i < 4) ? 4 * l16 + i : 64 + 4 * l16 + (i - 4); }

// ---- packed fp32 arithmetic ------------------------------------------------------------------------------
// sm_100 has two-wide fp32 instructions (FFMA2 / FMUL2 / FADD2 on 64-bit register pairs).  The pair loop is
// bound by the instruction issue rate, and every elementwise step acts on 8 independent hidden units per
// lane, so they are processed as 4 float2 values: half the issue slots for the same IEEE results.
using F2 = float2;
constexpr int HP = HPL / 2;
__device__ __forceinline__ F2 fma2(F2 a, F2 b, F2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ F2 mul2(F2 a, F2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ F2 add2(F2 a, F2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ F2 bc(float x) { return make_float2(x, x); }

// Per-lane slice of the head parameters, pre-scaled so that the pair loop works on ys = c * y with
// c = sqrt(log2(e) / 2): exp(-y^2 / 2) is then a single ex2(-ys * ys), and all other constants fold.
constexpr float kC = 0.84932180028801904f;          // sqrt(0.5 * log2(e))
constexpr float kErfP = 0.47047f * 0.70710678118654752f / kC;     // A&S 7.1.25 p applied to |ys|
constexpr float kPdf = 0.3989422804014327f / kC;    // y * pdf(y) = ys * e * kPdf
struct HeadConst {
  F2 gs[HP], bs[HP];   // gamma * c, beta * c
  F2 w2c[HP];          // w2 / c    (w2 . GELU(y) = sum w2c * (ys * Phi))
  F2 w2g[HP];          // w2 * gamma (d xhat / d out)
  float b2;
  __device__ __forceinline__ void load(const float* gamma, const float* beta, const float* w2, const float* b2p,
                                       int l16);
};

struct PairOut {
  float s;           // head output
  float rstd;
  F2 xh[HP];         // normalised pre-activation
  F2 g[HP];          // c * GELU(y)
  F2 gp[HP];         // GELU'(y)
  float m1, m2;      // mean_h(q), mean_h(q * xh), q = w2 * gamma * gp
};

__device__ __forceinline__ float fast_rcp(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float fast_ex2(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
// tanh(x) = 1 - 2 / (1 + e^{2x}); saturates correctly when e^{2x} over/underflows
__device__ __forceinline__ float fast_tanh(float x) {
  return 1.f - 2.f * fast_rcp(1.f + fast_ex2(x * (2.f * 1.4426950408889634f)));
}

// Head on one pre-activation difference hc (already mean-free over h).  erf by Abramowitz-Stegun 7.1.25
// (|err| <= 2.5e-5), sharing exp(-y^2/2) between erf and the Gaussian density needed by GELU'.
// Both half warps of a warp always execute this together (an invalid pair is masked later), so the
// xor-shuffles with offsets < 16 use the full mask and stay inside each half.
// HOIST: o.rstd was filled by the caller (read from the rank_rstd_gemm output) and the sum of squares is skipped.
template <bool GRAD, bool HOIST = false>
__device__ __forceinline__ void head_eval(const F2 (&hc)[HP], const HeadConst& hcst, float ln_eps, int use_tanh,
                                          PairOut& o) {
  if (!HOIST) {
    F2 ss2 = bc(0.f);
#pragma unroll
    for (int i = 0; i < HP; ++i) ss2 = fma2(hc[i], hc[i], ss2);
    const float ss = half_sum(ss2.x + ss2.y, 0xffffffffu);
    o.rstd = rsqrtf(fmaf(ss, 1.f / H, ln_eps));
  }
  const F2 r2 = bc(o.rstd);
  F2 acc2 = bc(0.f), m1_2 = bc(0.f), m2_2 = bc(0.f);
#pragma unroll
  for (int i = 0; i < HP; ++i) {
    const F2 xh = mul2(hc[i], r2);
    const F2 ys = fma2(xh, hcst.gs[i], hcst.bs[i]);               // c * y
    const F2 ti = fma2(bc(kErfP), make_float2(fabsf(ys.x), fabsf(ys.y)), bc(1.f));
    const F2 t = make_float2(fast_rcp(ti.x), fast_rcp(ti.y));
    // -(a1 t + a2 t^2 + a3 t^3), Abramowitz-Stegun 7.1.25 (|erf error| <= 2.5e-5, far inside the 1e-3 loss
    // bar); negated so that erf_abs = 1 + np * e is a single fma
    F2 np = fma2(t, bc(-0.7478556f), bc(0.0958798f));
    np = fma2(t, np, bc(-0.3480242f));
    np = mul2(np, t);
    const F2 sq = mul2(ys, ys);
    const F2 e = make_float2(fast_ex2(-sq.x), fast_ex2(-sq.y));   // exp(-y^2 / 2)
    const F2 ea = fma2(np, e, bc(1.f));                            // erf(|y| / sqrt 2)
    const F2 phi = fma2(bc(0.5f), make_float2(copysignf(ea.x, ys.x), copysignf(ea.y, ys.y)), bc(0.5f));
    const F2 g = mul2(ys, phi);                                    // c * GELU(y)
    acc2 = fma2(hcst.w2c[i], g, acc2);
    o.xh[i] = xh;
    o.g[i] = g;
    if (GRAD) {
      const F2 gp = fma2(mul2(ys, e), bc(kPdf), phi);              // Phi(y) + y * pdf(y)
      const F2 q = mul2(hcst.w2g[i], gp);
      o.gp[i] = gp;
      m1_2 = add2(m1_2, q);
      m2_2 = fma2(q, xh, m2_2);
    }
  }
  float acc = acc2.x + acc2.y, m1 = m1_2.x + m1_2.y, m2 = m2_2.x + m2_2.y;
  // the three reductions are independent: one shuffle phase with three interleaved chains
#pragma unroll
  for (int off = 8; off > 0; off >>= 1) {
    acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if (GRAD) {
      m1 += __shfl_xor_sync(0xffffffffu, m1, off);
      m2 += __shfl_xor_sync(0xffffffffu, m2, off);
    }
  }
  acc += hcst.b2;
  o.s = use_tanh ? fast_tanh(acc) : acc;
  if (GRAD) {
    o.m1 = m1 * (1.f / H);
    o.m2 = m2 * (1.f / H);
  }
}

// element j (0..7) of a lane <-> hidden unit hidx(l16, j); pair i holds elements (2i, 2i+1)
__device__ __forceinline__ void HeadConst::load(const float* gamma, const float* beta, const float* w2,
                                                const float* b2p, int l16) {
#pragma unroll
  for (int i = 0; i < HP; ++i) {
    const int h0 = hidx(l16, 2 * i), h1 = hidx(l16, 2 * i + 1);
    gs[i] = make_float2(gamma[h0] * kC, gamma[h1] * kC);
    bs[i] = make_float2(beta[h0] * kC, beta[h1] * kC);
    w2c[i] = make_float2(w2[h0] * (1.f / kC), w2[h1] * (1.f / kC));
    w2g[i] = make_float2(w2[h0] * gamma[h0], w2[h1] * gamma[h1]);
  }
  b2 = b2p[0];
}

struct RankParams {
  const float* u;        // (S, K, H) fp32: f W1^T
  const float* depth;    // (S, K)
  const float* b1;       // (H)
  const float* gamma;
  const float* beta;
  const float* w2;
  const float* b2;
  const float* rstd;     // (S, K, K) [set][b][a]: LayerNorm 1 / sigma of every ordered pair (rank_rstd_gemm)
  const float* inv_count;  // (S)   1 / #valid pairs (joint or per set), 0 if none
  const float* w_rank;     // (S) or nullptr
  int K, S;
  int b_per_warp;        // b rows per warp and CTA (the b tile is WARPS * b_per_warp rows)
  int mode;              // 0 logistic, 1 hinge
  int use_tanh;
  float thr, margin, ln_eps;
  double* loss_sum;      // (S) sum of pair losses (unnormalised)
  float* dub_part;       // (S, TA, K, H) partial gradient of u on the b side, per a tile
  float* dua_part;       // (S, TB, K, H) partial (positive) gradient flowing to -u_a, per b tile
  float* gparam;         // packed parameter gradients: [W1 (H*D) | b1 | gamma | beta | w2 | b2]
  int64_t gparam_off;    // offset of b1 inside gparam (= H * D)
};

// ---- pair kernel: elementwise part of one ordered pair (a -> b), 8 hidden units per lane as 4 float2 ----------
// Differences to head_eval (kept for the L1 kernel): the Gaussian density is folded into the exponent
// (e' = pdf-scaled exp(-y^2/2) = ex2(-ys^2 + log2 kPdf), the erf polynomial is pre-divided by kPdf, so GELU' is one
// fma), and the LayerNorm-backward mean term m1 is not formed at all: sum_pairs alpha m1 = mean_h(sum_pairs alpha q),
// so the accumulated gradient rows are centred once in rank_reduce_du instead of once per pair.
// 22 packed fp32 ops per float2 forward + backward (was 26).  The pair loop is bound by issue slots as much as by the
// FMA pipe (a packed op holds the pipe 2 cycles but takes one slot), so integer-pipe tricks that trade one packed op
// for several LOP3 / SEL were measured in SASS and rejected.
constexpr float kLog2Pdf = -1.0901312512086083f;          // log2(kPdf)
constexpr float kN1 = -2.f * 0.37046028286393695f;       // -a1 / kPdf   (A&S 7.1.25: a1, a2, a3)
constexpr float kN2 = 2.f * 0.10206088492966209f;
constexpr float kN3 = -2.f * 0.7960676214969513f;

struct PairFwd {
  F2 xh[HP];   // normalised pre-activation
  F2 g[HP];    // c * GELU(y)
  F2 gp[HP];   // GELU'(y)
  float acc;   // partial (this lane) of w2 . GELU
  float m2;    // partial of sum_h w2 gamma GELU' xh
};

template <bool GRAD>
__device__ __forceinline__ void pair_forward(const F2 (&hcv)[HP], float rstd, const HeadConst& hc, PairFwd& o) {
  const F2 r2 = bc(rstd);
  F2 acc2 = bc(0.f), m2_2 = bc(0.f);
#pragma unroll
  for (int i = 0; i < HP; ++i) {
    const F2 xh = mul2(hcv[i], r2);
    const F2 ys = fma2(xh, hc.gs[i], hc.bs[i]);                   // c * y
    const F2 ti = fma2(bc(kErfP), make_float2(fabsf(ys.x), fabsf(ys.y)), bc(1.f));
    // (one reciprocal per float2 via 1 / a = b / (a b) was measured: -8 MUFU, +12 FMA-pipe cycles per step, 2 % slower)
    const F2 t = make_float2(fast_rcp(ti.x), fast_rcp(ti.y));
    F2 np = fma2(t, bc(kN3), bc(kN2));
    np = fma2(t, np, bc(kN1));
    np = mul2(np, t);                                             // -(a1 t + a2 t^2 + a3 t^3) / kPdf
    const F2 sq = fma2(ys, ys, bc(-kLog2Pdf));
    const F2 e = make_float2(fast_ex2(-sq.x), fast_ex2(-sq.y));   // kPdf exp(-y^2 / 2)
    const F2 ea = fma2(np, e, bc(1.f));                           // erf(|y| / sqrt 2)
    const F2 phi = fma2(bc(0.5f), make_float2(copysignf(ea.x, ys.x), copysignf(ea.y, ys.y)), bc(0.5f));
    const F2 g = mul2(ys, phi);                                   // c * GELU(y)
    acc2 = fma2(hc.w2c[i], g, acc2);
    o.xh[i] = xh;
    o.g[i] = g;
    if (GRAD) {
      const F2 gp = fma2(ys, e, phi);                             // Phi(y) + y pdf(y)
      o.gp[i] = gp;
      m2_2 = fma2(hc.w2g[i], mul2(gp, xh), m2_2);
    }
  }
  o.acc = acc2.x + acc2.y;
  o.m2 = m2_2.x + m2_2.y;
}

// dynamic smem: va[TILE_A][H] | dua[TILE_A][H] | da[TILE_A] | red[3*H + 1]
//
// Synchronisation of the shared a-side gradient tile `dua`.  Warp w walks the ring of 32 row-pair slots starting
// SPACING (= 4) slots after warp w - 1, so two warps touch the same slot only when their step counters differ by a
// multiple of SPACING.  A CTA barrier every kBarrierPeriod = SPACING steps bounds the drift between any two warps to
// SPACING - 1 steps, which keeps the read-modify-writes race-free; the steps between two barriers are unrolled.
// Measured alternatives at cfg2 (tools/time_rank.py): barrier every 2 steps 2.98 ms, every 4 steps unrolled 2.90 ms, a
// barrier-free release / acquire flag ring between neighbouring warps 3.05 ms (the flag traffic costs more than the
// barrier it removes), no synchronisation at all (wrong results, lower bound) 2.81 ms.
// The pair loop addresses shared memory through 32-bit shared-window addresses that are made opaque to the compiler
// once (opaque()): derived from threadIdx they would be rematerialised inside the loop (S2R + shifts, ~25 cycles of
// exposed latency each) whenever registers get tight.
__device__ __forceinline__ float4 lds128(uint32_t saddr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr) : "memory");
  return v;
}
__device__ __forceinline__ void sts128(uint32_t saddr, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float lds32(uint32_t saddr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(saddr) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t opaque(uint32_t x) {
  asm volatile("" : "+r"(x));
  return x;
}

// MODE 0 logistic / 1 hinge and the head's tanh are template parameters: the per-pair scalar chain is a third of the
// issue slots of a step, uniform branches on kernel parameters inside it are not free.
template <bool GRAD, int MODE, bool TANH>
__global__ void __launch_bounds__(WARPS * 32, CTAS_PER_SM) rank_pairs(RankParams p) {
  extern __shared__ __align__(16) float smem[];
  float* va = smem;                       // centred u_a rows of the a tile
  float* dua = va + TILE_A * H;           // accumulated d/d(u_a) (sign applied at reduction)
  float* da = dua + TILE_A * H;           // depths of the a tile (NaN outside the set: such a pair is never valid)
  float* red = da + TILE_A;                 // cross-warp reduction of parameter gradients
  const int set = blockIdx.z, ta = blockIdx.x, tb = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, l16 = lane & 15, half = lane >> 4;
  const int K = p.K;
  const float* U = p.u + (int64_t)set * K * H;
  const float* Dp = p.depth + (int64_t)set * K;

  // ---- load and centre the a tile; it is stored NEGATED so that hc = vb + va is a plain packed add ----
  for (int r = warp; r < TILE_A; r += WARPS) {
    const int a = ta * TILE_A + r;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (a < K) v = *reinterpret_cast<const float4*>(U + (int64_t)a * H + 4 * lane);
    float m = (v.x + v.y) + (v.z + v.w);
    m = warp_sum(m) * (1.f / H);
    *reinterpret_cast<float4*>(va + r * H + 4 * lane) = make_float4(m - v.x, m - v.y, m - v.z, m - v.w);
    if (GRAD) *reinterpret_cast<float4*>(dua + r * H + 4 * lane) = make_float4(0.f, 0.f, 0.f, 0.f);
    if (lane == 0) da[r] = (a < K) ? Dp[a] : __int_as_float(0x7fc00000);
  }
  HeadConst hc;
  float b1_mean;
  {
    float bsum = 0.f;
    for (int h = lane; h < H; h += 32) bsum += p.b1[h];
    b1_mean = warp_sum(bsum) * (1.f / H);
    hc.load(p.gamma, p.beta, p.w2, p.b2, l16);
  }
  const float inv_cnt = p.inv_count[set] * (p.w_rank ? p.w_rank[set] : 1.f);
  __syncthreads();

  F2 dgam[HP], dbet[HP], dw2[HP];
  float db2 = 0.f;
#pragma unroll
  for (int i = 0; i < HP; ++i) dgam[i] = dbet[i] = dw2[i] = bc(0.f);
  float loss_local = 0.f;
  // shared-window addresses of this lane's 16-byte column of a ring slot (rows 2 * slot + half), opaque to the compiler
  constexpr uint32_t kSlotBytes = 2 * H * 4, kRingBytes = SLOTS * kSlotBytes, kDuaOff = TILE_A * H * 4;
  const uint32_t va_lane = opaque((uint32_t)__cvta_generic_to_shared(va) + (half * H + 4 * l16) * 4);
  const uint32_t da_lane = opaque((uint32_t)__cvta_generic_to_shared(da) + half * 4);
  // ring position of this warp (byte offset of its current slot): starts SPACING * warp slots into the ring and comes
  // back to the same slot after every full walk of SLOTS steps, so it simply persists from one b row to the next
  uint32_t soff = opaque((uint32_t)(SPACING * warp) * kSlotBytes);
  const int a_tile0 = ta * TILE_A;

  // rstd of one b row against the 64 rows of the a tile: the 16 lanes of a half hold ring slots l16 and l16 + 16
  // (a = 2 * slot + half), i.e. two fully coalesced 128-byte loads per warp and b row
  static_assert(SLOTS == 32, "rstd registers cover 2 x 16 ring slots");
  auto load_rstd = [&](int b_, float& x0, float& x1) {
    const float* rrow = p.rstd + ((int64_t)set * K + b_) * K + ta * TILE_A + half;
    const int a0 = ta * TILE_A + 2 * l16 + half, a1 = a0 + 32;
    x0 = (b_ < K && a0 < K) ? __ldg(rrow + 2 * l16) : 1.f;
    x1 = (b_ < K && a1 < K) ? __ldg(rrow + 2 * l16 + 32) : 1.f;
  };
  float rs0_next, rs1_next;
  // b rows are dealt round-robin to the warps (row bi * WARPS + warp of the tile), so that a partial last tile
  // (K = 300: 44 of 128 rows) still keeps all 8 warps busy and the CTA stops as soon as the rows run out
  const int tile_b = WARPS * p.b_per_warp;
  load_rstd(tb * tile_b + warp, rs0_next, rs1_next);
  // depth of row a = 2 * slot + half of the current slot (slot * 8 bytes into `da`); always loaded one step ahead
  float da_cur = lds32(da_lane + (soff >> 7));
  for (int bi = 0; bi < p.b_per_warp; ++bi) {
    if (tb * tile_b + bi * WARPS >= K) break;       // CTA-uniform: no row left for any warp
    const int b = tb * tile_b + bi * WARPS + warp;
    const bool b_ok = b < K;     // warp-uniform
    F2 vb[HP], dub[HP];
    float d_b = __int_as_float(0x7fc00000);     // NaN: no pair of a row outside the set is valid
    {
      float4 u0 = make_float4(0.f, 0.f, 0.f, 0.f), u1 = u0;
      if (b_ok) {
        u0 = *reinterpret_cast<const float4*>(U + (int64_t)b * H + 4 * l16);
        u1 = *reinterpret_cast<const float4*>(U + (int64_t)b * H + 64 + 4 * l16);
        d_b = Dp[b];
      }
      const float4 c0 = __ldg(reinterpret_cast<const float4*>(p.b1 + 4 * l16));          // b1 (L1-resident)
      const float4 c1 = __ldg(reinterpret_cast<const float4*>(p.b1 + 64 + 4 * l16));
      float m = ((u0.x + u0.y) + (u0.z + u0.w)) + ((u1.x + u1.y) + (u1.z + u1.w));
      m = half_sum(m, 0xffffffffu) * (1.f / H);
      const F2 nm = bc(-(m + b1_mean));      // w_b = (u_b - mean u_b) + (b1 - mean b1)
      vb[0] = add2(add2(make_float2(u0.x, u0.y), nm), make_float2(c0.x, c0.y));
      vb[1] = add2(add2(make_float2(u0.z, u0.w), nm), make_float2(c0.z, c0.w));
      vb[2] = add2(add2(make_float2(u1.x, u1.y), nm), make_float2(c1.x, c1.y));
      vb[3] = add2(add2(make_float2(u1.z, u1.w), nm), make_float2(c1.z, c1.w));
#pragma unroll
      for (int i = 0; i < HP; ++i) dub[i] = bc(0.f);
    }
    // 1 / sigma of this b row's pairs: taken from the registers loaded during the previous row, and the next
    // row's values are requested now
    const float rs0 = rs0_next, rs1 = rs1_next;
    // does any pair of this b row carry the "recompute directly" flag of the Gram epilogue?  (warp-uniform, rare)
    const bool row_flagged = __any_sync(0xffffffffu, rs0 < 0.f || rs1 < 0.f);
    if (bi + 1 < p.b_per_warp) load_rstd(b + WARPS, rs0_next, rs1_next);
    // The walk over the a tile exists twice: the common one trusts the Gram rstd, the rare one (a flagged pair in
    // this b row) re-derives 1 / sigma from the pair itself where the flag is set.
    auto walk_a_tile = [&](auto check_tag) {
    constexpr bool CHECK = decltype(check_tag)::value;
#pragma unroll kRankUnroll
    for (int t = 0; t < SLOTS; ++t) {
      // staggered a index: at any step the half-warps of the CTA work on different rows
      const uint32_t slot = soff >> 10;
      const uint32_t va_addr = va_lane + soff;
      const uint32_t soff_next = (soff + kSlotBytes) & (kRingBytes - 1);
      const float dd = d_b - da_cur;       // NaN when a or b lies outside the set
      da_cur = lds32(da_lane + (soff_next >> 7));     // next step's depth: its latency hides behind this step
      float rs = __shfl_sync(0xffffffffu, (slot < 16) ? rs0 : rs1, slot & 15, 16);     // within each 16-lane half
      const bool valid = (MODE == 0) ? (fabsf(dd) > p.thr) : (fabsf(tanhf(dd)) > p.thr);
      // both halves run the same instruction stream (98 % of the pairs are valid); an invalid pair is masked out
      // of every accumulation below.  A slot is skipped only when it lies outside the set for both halves
      // (warp-uniform test on the indices, no vote on loaded data).
      if (b_ok && a_tile0 + 2 * (int)slot < K) {
        const float4 a0 = lds128(va_addr);
        const float4 a1 = lds128(va_addr + 256);
        F2 hcv[HP];
        hcv[0] = add2(vb[0], make_float2(a0.x, a0.y));
        hcv[1] = add2(vb[1], make_float2(a0.z, a0.w));
        hcv[2] = add2(vb[2], make_float2(a1.x, a1.y));
        hcv[3] = add2(vb[3], make_float2(a1.z, a1.w));
        if (CHECK) {
          // flagged by the Gram epilogue (near-duplicate rows): sum of squares of this pair's h_c directly
          F2 ss2 = bc(0.f);
#pragma unroll
          for (int i = 0; i < HP; ++i) ss2 = fma2(hcv[i], hcv[i], ss2);
          const float ss = half_sum(ss2.x + ss2.y, 0xffffffffu);
          if (rs < 0.f) rs = rsqrtf(fmaf(ss, 1.f / H, p.ln_eps));
        }
        PairFwd o;
        pair_forward<GRAD>(hcv, rs, hc, o);
        // the two reductions over the 16 lanes of the pair ride in one float2: packed adds, two shuffles per level
        F2 am = make_float2(o.acc, o.m2);
        // (a fixed-point REDUX over the 16 lanes instead of the shuffle butterfly was measured: 3.47 vs 2.91 ms)
#pragma unroll
        for (int off = 8; off > 0; off >>= 1) {
          F2 other;
          other.x = __shfl_xor_sync(0xffffffffu, am.x, off);
          other.y = GRAD ? __shfl_xor_sync(0xffffffffu, am.y, off) : 0.f;
          am = add2(am, other);
        }
        const float acc = am.x + hc.b2;
        const float sc = TANH ? fast_tanh(acc) : acc;
        // sign(dd) s without a multiply: flip the sign bit of s where dd < 0 (a valid pair has dd != 0)
        const float ssc = __uint_as_float(__float_as_uint(sc) ^ (__float_as_uint(dd) & 0x80000000u));
        float l, dl;      // pair loss and d l / d (sign(dd) s)
        if (MODE == 0) {
          const float ex = fast_ex2(ssc * -1.4426950408889634f);
          l = __logf(1.f + ex);
          dl = -ex * fast_rcp(1.f + ex);
        } else {
          const float m = p.margin - ssc;
          l = fmaxf(m, 0.f);
          dl = (m > 0.f) ? -1.f : 0.f;
        }
        // every lane of the pair adds the same value; the CTA reduction divides by 16
        loss_local += valid ? l : 0.f;
        if (GRAD) {
          // d total / d (w2.g + b2); zero for a masked pair, which zeroes every contribution below
          const float dsg = __uint_as_float(__float_as_uint(dl) ^ (__float_as_uint(dd) & 0x80000000u));   // d l / d s
          const float dout = valid ? dsg * (TANH ? fmaf(-sc, sc, 1.f) : 1.f) * inv_cnt : 0.f;
          db2 += dout;
          // d h = alpha (q - m1 - m2 xh), q = w2 gamma GELU'; the m1 part is the centring done by rank_reduce_du
          const F2 dout2 = bc(dout), alpha2 = bc(dout * rs), nm2 = bc(am.y * (-1.f / H));
          F2 tq[HP];
#pragma unroll
          for (int i = 0; i < HP; ++i) {
            // parameter sums are kept unscaled: d w2 = dw2 / c, d beta = w2 * dbet, d gamma = w2 * dgam
            dw2[i] = fma2(dout2, o.g[i], dw2[i]);
            dbet[i] = fma2(dout2, o.gp[i], dbet[i]);
            dgam[i] = fma2(dout2, mul2(o.gp[i], o.xh[i]), dgam[i]);
            tq[i] = fma2(nm2, o.xh[i], mul2(hc.w2g[i], o.gp[i]));
            dub[i] = fma2(alpha2, tq[i], dub[i]);
          }
          const float4 c0 = lds128(va_addr + kDuaOff), c1 = lds128(va_addr + kDuaOff + 256);
          const F2 s0 = fma2(alpha2, tq[0], make_float2(c0.x, c0.y)), s1 = fma2(alpha2, tq[1], make_float2(c0.z, c0.w));
          const F2 s2 = fma2(alpha2, tq[2], make_float2(c1.x, c1.y)), s3 = fma2(alpha2, tq[3], make_float2(c1.z, c1.w));
          sts128(va_addr + kDuaOff, make_float4(s0.x, s0.y, s1.x, s1.y));
          sts128(va_addr + kDuaOff + 256, make_float4(s2.x, s2.y, s3.x, s3.y));
        }
      }
      soff = soff_next;
      // every warp of the CTA runs the same number of steps (the row break above is CTA-uniform), so the barrier is safe
      static_assert(kBarrierPeriod <= SPACING && SLOTS % kBarrierPeriod == 0 && (kBarrierPeriod & (kBarrierPeriod - 1)) == 0,
                    "barrier period");
      if (GRAD && (t & (kBarrierPeriod - 1)) == kBarrierPeriod - 1) __syncthreads();
    }
    };
    if (row_flagged) walk_a_tile(std::true_type{});
    else walk_a_tile(std::false_type{});
    if (GRAD) {
      // combine the two half warps and store this b row's partial (over the a tile) gradient
#pragma unroll
      for (int i = 0; i < HP; ++i) {
        dub[i].x += __shfl_xor_sync(0xffffffffu, dub[i].x, 16);
        dub[i].y += __shfl_xor_sync(0xffffffffu, dub[i].y, 16);
      }
      if (b_ok && half == 0) {
        float* dst = p.dub_part + (((int64_t)set * gridDim.x + ta) * K + b) * H;
        *reinterpret_cast<float4*>(dst + 4 * l16) = make_float4(dub[0].x, dub[0].y, dub[1].x, dub[1].y);
        *reinterpret_cast<float4*>(dst + 64 + 4 * l16) = make_float4(dub[2].x, dub[2].y, dub[3].x, dub[3].y);
      }
    }
  }
  // ---- CTA-level reductions ----
  loss_local = warp_sum(loss_local) * (1.f / 16.f);      // all 16 lanes of a pair carried its loss
  if (lane == 0 && loss_local != 0.f) atomicAdd(p.loss_sum + set, (double)loss_local);
  if (GRAD) {
    __syncthreads();
    for (int e = threadIdx.x; e < TILE_A * H; e += blockDim.x) {
      const int r = e / H, a = ta * TILE_A + r;
      if (a < K) p.dua_part[(((int64_t)set * gridDim.y + tb) * K + a) * H + (e - r * H)] = dua[e];
    }
    for (int e = threadIdx.x; e < 3 * H + 1; e += blockDim.x) red[e] = 0.f;
    __syncthreads();
#pragma unroll
    for (int j = 0; j < HPL; ++j) {
      const int i = j >> 1;
      const bool hi = j & 1;
      const float w2h = (hi ? hc.w2c[i].y : hc.w2c[i].x) * kC;
      const float vg = hi ? dgam[i].y : dgam[i].x, vb_ = hi ? dbet[i].y : dbet[i].x, vw = hi ? dw2[i].y : dw2[i].x;
      const float g2 = (vg + __shfl_xor_sync(0xffffffffu, vg, 16)) * w2h;
      const float b2v = (vb_ + __shfl_xor_sync(0xffffffffu, vb_, 16)) * w2h;
      const float w2v = (vw + __shfl_xor_sync(0xffffffffu, vw, 16)) * (1.f / kC);
      if (half == 0) {
        const int h = hidx(l16, j);
        atomicAdd(red + h, g2);
        atomicAdd(red + H + h, b2v);
        atomicAdd(red + 2 * H + h, w2v);
      }
    }
    db2 = warp_sum(db2) * (1.f / 16.f);
    if (lane == 0) atomicAdd(red + 3 * H, db2);
    __syncthreads();
    float* gp = p.gparam + p.gparam_off;   // [b1 | gamma | beta | w2 | b2]
    for (int e = threadIdx.x; e < 3 * H + 1; e += blockDim.x)
      if (red[e] != 0.f) atomicAdd(gp + H + e, red[e]);
  }
}

// ------------------------------------------------------------------------------------------
// LayerNorm statistics of all K^2 pairs, hoisted onto the tensor cores.
//   h_c(a -> b) = w_b - v_a,  v = u - mean_h(u),  w = v + (b1 - mean(b1))        (mean-free over h)
//   sum_h h_c^2 = |w_b|^2 + |v_a|^2 - 2 w_b . v_a
// The Gram term is one split-bf16 GEMM per set ([hi|hi|lo] x [hi|lo|hi] panels, ~16 mantissa bits on the
// products, the same accuracy class as u itself); its epilogue writes rstd = rsqrt(ss / H + eps) for every pair.
// This removes the sum of squares, its 4 shuffle rounds and the rsqrt from the head of every pair step.
// ------------------------------------------------------------------------------------------
// mean keypoint feature of every centring group (one set, or the two sets of an image pair when the cross-view L1
// term couples them).  Every loss term depends on feature DIFFERENCES inside a group only, and sum_k d loss / d u_k = 0
// over a group, so u = (f - mu) W1^T gives the same losses and gradients as u = f W1^T -- while the bf16 split of
// the GEMM operands (and the Gram matrix of 4.3) resolves the deviations instead of the component all keypoints of
// an image share (real ViT tokens: the common component is 10-1000x the differences).
// grid (groups, ceil(D / 64)), block 1024 = 16 row slices x 64 channels
__global__ void __launch_bounds__(1024) rank_group_mean(const float* __restrict__ f, int rows, int D, float* __restrict__ mu) {
  __shared__ float part[16][64];
  const int g = blockIdx.x, cl = threadIdx.x & 63, c = blockIdx.y * 64 + cl, slice = threadIdx.x >> 6;
  float s0 = 0.f, s1 = 0.f;
  if (c < D) {
    const float* col = f + (int64_t)g * rows * D + c;
    int r = slice;
    for (; r + 16 < rows; r += 32) {      // two independent chains
      s0 += __ldg(col + (int64_t)r * D);
      s1 += __ldg(col + (int64_t)(r + 16) * D);
    }
    if (r < rows) s0 += __ldg(col + (int64_t)r * D);
  }
  part[slice][cl] = s0 + s1;
  __syncthreads();
  if (slice == 0 && c < D) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 16; ++k) t += part[k][cl];
    mu[(int64_t)g * D + c] = t / (float)rows;
  }
}

// one warp per row of u: centre, add the centred bias on the b side, split into bf16 panels, squared norms
__global__ void __launch_bounds__(256) rank_gram_prep(const float* __restrict__ u, const float* __restrict__ b1, int64_t R,
                                                      __nv_bfloat16* __restrict__ Wb3, __nv_bfloat16* __restrict__ Va3,
                                                      float* __restrict__ nb, float* __restrict__ na) {
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= R) return;
  const float4 bv = make_float4(b1[4 * lane], b1[4 * lane + 1], b1[4 * lane + 2], b1[4 * lane + 3]);   // no alignment assumed
  const float bm = warp_sum((bv.x + bv.y) + (bv.z + bv.w)) * (1.f / H);
  const float4 uv = *reinterpret_cast<const float4*>(u + r * H + 4 * lane);
  const float m = warp_sum((uv.x + uv.y) + (uv.z + uv.w)) * (1.f / H);
  const float v[4] = {uv.x - m, uv.y - m, uv.z - m, uv.w - m};
  const float w[4] = {v[0] + (bv.x - bm), v[1] + (bv.y - bm), v[2] + (bv.z - bm), v[3] + (bv.w - bm)};
  float sv = 0.f, sw = 0.f;
  uint16_t vh[4], vl[4], wh[4], wl[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    sv = fmaf(v[i], v[i], sv);
    sw = fmaf(w[i], w[i], sw);
    split_detail::hi_lo(v[i], vh[i], vl[i]);
    split_detail::hi_lo(w[i], wh[i], wl[i]);
  }
  sv = warp_sum(sv);
  sw = warp_sum(sw);
  auto pk = [](const uint16_t (&x)[4]) {
    return make_uint2((uint32_t)x[0] | ((uint32_t)x[1] << 16), (uint32_t)x[2] | ((uint32_t)x[3] << 16));
  };
  uint2* wo = reinterpret_cast<uint2*>(Wb3 + r * 3 * H) + lane;     // [hi | hi | lo]
  wo[0] = pk(wh);
  wo[H / 4] = pk(wh);
  wo[2 * (H / 4)] = pk(wl);
  uint2* vo = reinterpret_cast<uint2*>(Va3 + r * 3 * H) + lane;     // [hi | lo | hi]
  vo[0] = pk(vh);
  vo[H / 4] = pk(vl);
  vo[2 * (H / 4)] = pk(vh);
  if (lane == 0) {
    nb[r] = sw;
    na[r] = sv;
  }
}

// rstd[set][b][a] = rsqrt(max(|w_b|^2 + |v_a|^2 - 2 acc, 0) / H + eps), or -1 where the Gram form is too inaccurate
constexpr float kGramMinRatio = 1.f / 64.f;      // ss / (|w|^2 + |v|^2) below which the pair is recomputed directly
struct EpiRstd {
  static constexpr int kScratchBytes = tc::kMaxEpiWarps * tc::kWarpTileBytes;
  struct Params {
    float* out;          // (S, K, K)
    int K;
    const float* nb;     // (S, K)
    const float* na;     // (S, K)
    float eps;
    int use_tma = 0;     // K % 4 == 0 and an aligned buffer: rows leave through TMA stores (tc::EpiStoreF32 has the details)
    alignas(64) CUtensorMap tm_out = {};
  };
  struct Pre {
    float nb;
  };
  __device__ static void pre(const Params& p, const tc::EpiCtx& cx, Pre& pr) {
    const int b = cx.m0 + cx.row;
    pr.nb = (b < p.K) ? p.nb[(int64_t)cx.b * p.K + b] : 0.f;
  }
  __device__ static void run(const Params& p, const tc::EpiCtx& cx, const Pre& pr) {
    float* t = reinterpret_cast<float*>(cx.scratch) + cx.epi_warp * tc::kWarpTileFloats;
    const int m_warp = cx.m0 + (cx.row & ~31);
    const int rows = p.K - m_warp;
    float* oslab = p.out + ((int64_t)cx.b * p.K + m_warp) * p.K;
    const float* na = p.na + (int64_t)cx.b * p.K;
    for (int c = cx.col_begin; c < cx.col_end; c += 32) {
      const int n = cx.n0 + c;
      if (n >= p.K) break;
      // |v_a|^2 of this chunk's 32 columns: one coalesced load, handed out by shuffle
      const float na_l = (n + cx.lane < p.K) ? __ldg(na + n + cx.lane) : 0.f;
      float v[32];
      tc::tmem_ld32(cx.tmem + c, v);
      if (rows <= 0) continue;
#pragma unroll
      for (int q = 0; q < 32; ++q) {
        const float nsum = pr.nb + __shfl_sync(0xffffffffu, na_l, q);
        const float ss = fmaxf(fmaf(-2.f, v[q], nsum), 0.f);
        // the difference of nearly identical rows cancels in |w|^2 + |v|^2 - 2 w.v (split-bf16 products carry ~2^-16 of
        // |w| |v|): flag such pairs (negative value) and let the pair kernel sum the squares directly
        v[q] = (ss < kGramMinRatio * nsum) ? -1.f : rsqrtf(fmaf(ss, 1.f / H, p.eps));
      }
      if (p.use_tma) {
        uint8_t* slab = cx.scratch + cx.epi_warp * tc::kStoreSlab32Bytes;
        if (cx.lane == 0) tc::tma_store_wait_read();
        __syncwarp();
#pragma unroll
        for (int q = 0; q < 8; ++q)
          *reinterpret_cast<float4*>(slab + tc::store_slab32_offset(cx.lane, q)) =
              make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
        tc::fence_proxy_async_smem();
        __syncwarp();
        if (cx.lane == 0) {
          tc::tma_store_3d(&p.tm_out, slab, n, m_warp, cx.b);
          tc::tma_store_commit();
        }
      } else {
        tc::warp_store_rows<float>(t, v, oslab + n, p.K, rows, p.K - n, cx.lane);
      }
    }
    if (p.use_tma) {
      if (cx.lane == 0) tc::tma_store_wait_read();
      __syncwarp();
    }
  }
};

// ------------------------------------------------------------------------------------------
// number of valid pairs per set -> inv_count (per set, or shared when joint_mean)
// ------------------------------------------------------------------------------------------
// grid (ceil(K / 256), S), block 256: thread = one a, loop over all b with the set's depths in shared memory
__global__ void __launch_bounds__(256) rank_count(const float* __restrict__ depth, int K, int mode, float thr,
                                                  int* __restrict__ count) {
  extern __shared__ float sdep[];     // K depths of this set
  __shared__ int red[32];
  const int set = blockIdx.y;
  const float* d = depth + (int64_t)set * K;
  for (int k = threadIdx.x; k < K; k += blockDim.x) sdep[k] = d[k];
  __syncthreads();
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  int c = 0;
  if (a < K) {
    const float da = sdep[a];
    if (mode == 0) {
      for (int b = 0; b < K; ++b) c += (fabsf(sdep[b] - da) > thr) ? 1 : 0;
    } else {
      for (int b = 0; b < K; ++b) c += (fabsf(tanhf(sdep[b] - da)) > thr) ? 1 : 0;
    }
  }
  c = block_sum(c, red);
  if (threadIdx.x == 0 && c) atomicAdd(count + set, c);
}
__global__ void rank_inv_count(const int* __restrict__ count, int S, int joint, float* __restrict__ inv) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S) return;
  long long n = count[s];
  if (joint) {
    n = 0;
    for (int k = 0; k < S; ++k) n += count[k];
  }
  inv[s] = n > 0 ? (float)(1.0 / (double)n) : 0.f;
}

// ------------------------------------------------------------------------------------------
// cross-view L1: pair p couples keypoint k of set 2p+1 (a) with keypoint k of set 2p (b).
// One half warp per keypoint.  Adds its u-gradient straight into the dub/dua partial slot 0.
// ------------------------------------------------------------------------------------------
template <bool GRAD>
__global__ void __launch_bounds__(256) rank_l1(RankParams p, const float* __restrict__ w_l1, double* __restrict__ l1_sum,
                                              float* __restrict__ du_extra /* (S, K, H) zero-initialised */) {
  __shared__ float red[3 * H + 1];
  const int pair = blockIdx.y;
  const int lane = threadIdx.x & 31, l16 = lane & 15, half = lane >> 4;
  const int k = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 2 + half;
  const int K = p.K;
  const int sb = 2 * pair, sa = 2 * pair + 1;
  HeadConst hc;
  hc.load(p.gamma, p.beta, p.w2, p.b2, l16);
  float dgam[HPL], dbet[HPL], dw2[HPL], db2 = 0.f, loss_local = 0.f;
#pragma unroll
  for (int i = 0; i < HPL; ++i) dgam[i] = dbet[i] = dw2[i] = 0.f;
  const bool ok = k < K;
  float hs[HPL];
  float m = 0.f;
#pragma unroll
  for (int i = 0; i < HPL; ++i) {
    const int h = hidx(l16, i);
    hs[i] = ok ? p.u[((int64_t)sb * K + k) * H + h] - p.u[((int64_t)sa * K + k) * H + h] + p.b1[h] : 0.f;
    m += hs[i];
  }
  m = half_sum(m, 0xffffffffu) * (1.f / H);
  F2 hcv[HP];
#pragma unroll
  for (int i = 0; i < HP; ++i) hcv[i] = make_float2(hs[2 * i] - m, hs[2 * i + 1] - m);
  PairOut o;
  head_eval<GRAD>(hcv, hc, p.ln_eps, p.use_tanh, o);
  if (ok) {
    const float tgt = tanhf(p.depth[(int64_t)sb * K + k] - p.depth[(int64_t)sa * K + k]);
    const float diff = o.s - tgt;
    if (l16 == 0) loss_local = fabsf(diff);
    if (GRAD) {
      const float sg = (diff > 0.f) ? 1.f : ((diff < 0.f) ? -1.f : 0.f);
      const float dout = sg * (p.use_tanh ? (1.f - o.s * o.s) : 1.f) * w_l1[pair] / (float)K;
      db2 = (l16 == 0) ? dout : 0.f;
      const float coef = dout * o.rstd;
#pragma unroll
      for (int j = 0; j < HPL; ++j) {
        const int i = j >> 1;
        const bool hi = j & 1;
        const int h = hidx(l16, j);
        const float gp = hi ? o.gp[i].y : o.gp[i].x, gg = hi ? o.g[i].y : o.g[i].x, xh = hi ? o.xh[i].y : o.xh[i].x;
        const float w2h = (hi ? hc.w2c[i].y : hc.w2c[i].x) * kC, w2g = hi ? hc.w2g[i].y : hc.w2g[i].x;
        const float pg = w2h * gp;
        dw2[j] = dout * gg * (1.f / kC);
        dbet[j] = dout * pg;
        dgam[j] = dout * pg * xh;
        const float dh = coef * (w2g * gp - o.m1 - xh * o.m2);
        du_extra[((int64_t)sb * K + k) * H + h] = dh;
        du_extra[((int64_t)sa * K + k) * H + h] = -dh;
      }
    }
  }
  loss_local = warp_sum(loss_local);
  if (lane == 0 && loss_local != 0.f) atomicAdd(l1_sum + pair, (double)loss_local);
  if (GRAD) {
    for (int e = threadIdx.x; e < 3 * H + 1; e += blockDim.x) red[e] = 0.f;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < HPL; ++i) {
      const int h = hidx(l16, i);
      atomicAdd(red + h, dgam[i]);
      atomicAdd(red + H + h, dbet[i]);
      atomicAdd(red + 2 * H + h, dw2[i]);
    }
    db2 = warp_sum(db2);
    if (lane == 0) atomicAdd(red + 3 * H, db2);
    __syncthreads();
    float* gp = p.gparam + p.gparam_off;
    for (int e = threadIdx.x; e < 3 * H + 1; e += blockDim.x)
      if (red[e] != 0.f) atomicAdd(gp + H + e, red[e]);
  }
}

// ------------------------------------------------------------------------------------------
// du = sum_ta dub_part - sum_tb dua_part (+ L1 part); emits the bf16 hi / lo panels of du and the b1 gradient
// grid (ceil(K/32), S), block 256 (8 warps x 32 lanes: lane -> 4 h, warp -> rows)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
    rank_reduce_du(const float* __restrict__ dub_part, const float* __restrict__ dua_part,
                   const float* __restrict__ du_extra, int S, int K, int TA, int TB,
                   __nv_bfloat16* __restrict__ du2 /* (S K, 2 H): [hi | lo] */, float* __restrict__ gb1) {
  __shared__ float colsum[8][H];
  const int set = blockIdx.y, k0 = blockIdx.x * 32;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  float bsum[4] = {0.f, 0.f, 0.f, 0.f};
  for (int r = w; r < 32; r += 8) {
    const int k = k0 + r;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (k < K) {
      // rank_pairs accumulates alpha (q - m2 xh) per pair; the LayerNorm-backward term -alpha m1 = -alpha mean_h(q)
      // summed over pairs is the mean over h of the accumulated row (mean_h(xh) = 0), removed here once per row
#pragma unroll 4
      for (int t = 0; t < TA; ++t) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(dub_part + (((int64_t)set * TA + t) * K + k) * H + 4 * lane));
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
      {
        const float mb = warp_sum((acc.x + acc.y) + (acc.z + acc.w)) * (1.f / H);
        acc.x -= mb; acc.y -= mb; acc.z -= mb; acc.w -= mb;
      }
      bsum[0] += acc.x; bsum[1] += acc.y; bsum[2] += acc.z; bsum[3] += acc.w;   // d b1 = sum over pairs of dh
      float4 aa = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
      for (int t = 0; t < TB; ++t) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(dua_part + (((int64_t)set * TB + t) * K + k) * H + 4 * lane));
        aa.x += v.x; aa.y += v.y; aa.z += v.z; aa.w += v.w;
      }
      {
        const float ma = warp_sum((aa.x + aa.y) + (aa.z + aa.w)) * (1.f / H);
        acc.x -= aa.x - ma; acc.y -= aa.y - ma; acc.z -= aa.z - ma; acc.w -= aa.w - ma;
      }
      if (du_extra) {
        const float4 v = *reinterpret_cast<const float4*>(du_extra + ((int64_t)set * K + k) * H + 4 * lane);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        if ((set & 1) == 0) { bsum[0] += v.x; bsum[1] += v.y; bsum[2] += v.z; bsum[3] += v.w; }   // b side of the L1 pair
      }
      // bf16 hi / lo split of du, row-major: the d feats GEMM reads the hi panel K-major, the d W1 GEMM reads both
      // panels MN-major (tc_gemm.cuh), so no transposed copy is written
      const float v[4] = {acc.x, acc.y, acc.z, acc.w};
      uint16_t hi[4], lo[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) split_detail::hi_lo(v[i], hi[i], lo[i]);
      __nv_bfloat16* row = du2 + ((int64_t)set * K + k) * 2 * H + 4 * lane;
      *reinterpret_cast<uint2*>(row) = make_uint2((uint32_t)hi[0] | ((uint32_t)hi[1] << 16), (uint32_t)hi[2] | ((uint32_t)hi[3] << 16));
      *reinterpret_cast<uint2*>(row + H) = make_uint2((uint32_t)lo[0] | ((uint32_t)lo[1] << 16), (uint32_t)lo[2] | ((uint32_t)lo[3] << 16));
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) colsum[w][4 * lane + i] = bsum[i];
  __syncthreads();
  if (threadIdx.x < H) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += colsum[i][threadIdx.x];
    if (t != 0.f) atomicAdd(gb1 + threadIdx.x, t);
  }
}

// ------------------------------------------------------------------------------------------
// operand preparation for the GEMMs
// ------------------------------------------------------------------------------------------
// epilogue: atomically accumulate acc into a single fp32 matrix shared by all batches (split-K).  The 32 x 32 chunk
// goes through the per-warp shared tile so that every atomic instruction covers whole row segments: 8 lanes x
// red.global.add.v4.f32 per row (4 rows per instruction) when the rows are 16-byte aligned, else one row of scalar
// atomics per instruction.
struct EpiAtomicAddF32 {
  static constexpr int kScratchBytes = tc::kMaxEpiWarps * tc::kWarpTileBytes;
  struct Params {
    float* C;
    int M, N;
    int64_t ldc;
  };
  struct Pre {};
  __device__ static void pre(const Params&, const tc::EpiCtx&, Pre&) {}
  __device__ static void run(const Params& p, const tc::EpiCtx& cx, const Pre&) {
    float* t = reinterpret_cast<float*>(cx.scratch) + cx.epi_warp * tc::kWarpTileFloats;
    const int lane = cx.lane;
    const int m_warp = cx.m0 + (cx.row & ~31);
    const int rows = p.M - m_warp;
    const bool vec = (p.ldc % 4 == 0) && (reinterpret_cast<uintptr_t>(p.C) % 16 == 0);
    for (int c = cx.col_begin; c < cx.col_end; c += 32) {
      const int n = cx.n0 + c;
      if (n >= p.N) break;
      float v[32];
      tc::tmem_ld32(cx.tmem + c, v);
      if (rows <= 0) continue;
#pragma unroll
      for (int q = 0; q < 32; ++q) t[lane * 33 + q] = v[q];
      __syncwarp();
      float* cslab = p.C + (int64_t)m_warp * p.ldc + n;
      if (vec && n + 32 <= p.N) {
        const int sub = lane >> 3, c4 = 4 * (lane & 7);      // 4 rows per instruction, 8 lanes x float4 per row
#pragma unroll
        for (int r = 0; r < 32; r += 4) {
          const int rr = r + sub;
          const float4 x = make_float4(t[rr * 33 + c4], t[rr * 33 + c4 + 1], t[rr * 33 + c4 + 2], t[rr * 33 + c4 + 3]);
          if (rr < rows) atomicAdd(reinterpret_cast<float4*>(cslab + (int64_t)rr * p.ldc + c4), x);
        }
      } else {
#pragma unroll 8
        for (int r = 0; r < 32; ++r) {
          const float x = t[r * 33 + lane];
          if (r < rows && n + lane < p.N) atomicAdd(cslab + (int64_t)r * p.ldc + lane, x);
        }
      }
      __syncwarp();
    }
  }
};

__global__ void rank_finalize(const double* __restrict__ loss_sum, const float* __restrict__ inv_count,
                              const double* __restrict__ l1_sum, int S, int K, float* __restrict__ loss_rank,
                              float* __restrict__ loss_l1) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < S) loss_rank[s] = (float)(loss_sum[s] * (double)inv_count[s]);
  if (loss_l1 && s < S / 2) loss_l1[s] = (float)(l1_sum[s] / (double)K);
}

struct RankWorkspace {
  __nv_bfloat16 *F3, *W3, *du2;
  __nv_bfloat16 *Wb3, *Va3;
  float *u, *inv_count, *dub_part, *dua_part, *du_extra, *nb, *na, *rstd, *mu;
  double *loss_sum, *l1_sum;
  int* count;
  size_t total;
  int ldd, TA, TB, groups, bpw;
  int gs;        // sets per split-K group of the d W1 contraction
};

// b rows per warp of a rank_pairs CTA.  A CTA is the unit of scheduling, two fit on an SM: the kernel takes
// ceil(CTAs / (2 SMs)) waves of (b_per_warp + 1) row walks (the +1 stands for loading the a tile and flushing its
// gradient), so few sets (strong scaling: 8 pairs per GPU) or a ragged K want smaller b tiles, while large problems keep
// 16 rows per warp, which writes the fewest a-side partial tiles.  Deterministic in (S, K): the workspace depends on it.
int rank_b_per_warp(int64_t S, int64_t K) {
  const int64_t slots = 2 * (int64_t)num_sms();
  const int64_t ta = ceil_div<int64_t>(K, TILE_A);
  int best = B_PER_WARP_MAX;
  double best_cost = 1e30;
  for (int bpw = B_PER_WARP_MAX; bpw >= 2; bpw /= 2) {
    const int64_t ctas = ta * ceil_div<int64_t>(K, (int64_t)WARPS * bpw) * S;
    const double cost = (double)ceil_div<int64_t>(ctas, slots) * (bpw + 1);
    if (cost < 0.97 * best_cost) {      // a smaller tile has to win by 3 %: it multiplies the partial-tile traffic
      best_cost = cost;
      best = bpw;
    }
  }
  return best;
}

RankWorkspace carve_rank(void* base, int64_t S, int64_t K, int64_t D, bool backward, bool l1) {
  RankWorkspace w{};
  Carver c(base);
  const int64_t R = S * K;
  w.ldd = (int)round_up<int64_t>(D, 8);
  // d W1 sums over all S K keypoints: split-K into at most 48 groups of sets (one GEMM batch entry each), so that
  // (groups x D / BN) tiles fill the SMs while the fp32 atomics of the epilogue are issued once per group
  w.groups = (int)(S < 48 ? S : 48);
  w.gs = (int)ceil_div<int64_t>(S, w.groups);
  w.groups = (int)ceil_div<int64_t>(S, w.gs);
  w.TA = (int)ceil_div<int64_t>(K, TILE_A);
  w.bpw = rank_b_per_warp(S, K);
  w.TB = (int)ceil_div<int64_t>(K, (int64_t)WARPS * w.bpw);
  w.F3 = c.take<__nv_bfloat16>(R * 3 * w.ldd);
  w.W3 = c.take<__nv_bfloat16>((int64_t)H * 3 * w.ldd);
  w.u = c.take<float>(R * H);
  w.mu = c.take<float>(S * D);
  w.Wb3 = c.take<__nv_bfloat16>(R * 3 * H);
  w.Va3 = c.take<__nv_bfloat16>(R * 3 * H);
  w.nb = c.take<float>(R);
  w.na = c.take<float>(R);
  w.rstd = c.take<float>(R * K);
  w.count = c.take<int>(S);
  w.inv_count = c.take<float>(S);
  w.loss_sum = c.take<double>(S);
  w.l1_sum = c.take<double>(S);
  if (backward) {
    w.du2 = c.take<__nv_bfloat16>(R * 2 * H);
    w.dub_part = c.take<float>(S * w.TA * K * H);
    w.dua_part = c.take<float>(S * w.TB * K * H);
    if (l1) w.du_extra = c.take<float>(R * H);
  }
  w.total = c.total();
  return w;
}

}  // namespace
}  // namespace gd3

using namespace gd3;

extern "C" {

size_t gd3_depth_head_loss_workspace(int64_t S, int64_t K, int64_t D, int with_backward, int with_l1) {
  if (S <= 0 || K <= 0 || D <= 0) return 0;
  return carve_rank(nullptr, S, K, D, with_backward != 0, with_l1 != 0).total;
}

int gd3_depth_head_loss(const float* feats, const float* depths, int64_t S, int64_t K, int64_t D, int64_t hidden,
                        const float* W1, const float* b1, const float* gamma, const float* beta, const float* w2,
                        const float* b2, int use_tanh, float ln_eps, int mode, float thr, float margin,
                        int joint_mean, const float* w_rank, const float* w_l1, float* loss_rank, float* loss_l1,
                        float* grad_feats, float* grad_params, void* workspace, size_t workspace_bytes,
                        void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (S == 0) return GD3_OK;
  GD3_REQUIRE(S > 0 && K >= 0 && D > 0, "gd3_depth_head_loss: bad sizes S=%lld K=%lld D=%lld", (long long)S,
              (long long)K, (long long)D);
  GD3_REQUIRE(hidden == H, "gd3_depth_head_loss: hidden width %lld not supported (fusion_layer uses %d)",
              (long long)hidden, H);
  GD3_REQUIRE(mode == 0 || mode == 1, "gd3_depth_head_loss: mode must be 0 (logistic) or 1 (hinge)");
  GD3_REQUIRE(loss_rank, "gd3_depth_head_loss: null loss output");
  GD3_REQUIRE((grad_feats == nullptr) == (grad_params == nullptr),
              "gd3_depth_head_loss: pass both gradient buffers or neither");
  const bool l1 = w_l1 != nullptr;
  GD3_REQUIRE(!l1 || (S % 2 == 0 && loss_l1), "gd3_depth_head_loss: the L1 term needs an even number of sets and loss_l1");
  GD3_REQUIRE(S <= 65535, "gd3_depth_head_loss: at most 65535 sets per call");
  GD3_REQUIRE(K <= 12288, "gd3_depth_head_loss: at most 12288 keypoints per set (K^2 pairs are evaluated)");
  const bool backward = grad_feats != nullptr;
  GD3_CHECK_CUDA(cudaMemsetAsync(loss_rank, 0, sizeof(float) * S, stream));
  if (l1) GD3_CHECK_CUDA(cudaMemsetAsync(loss_l1, 0, sizeof(float) * (S / 2), stream));
  const int64_t nparam = (int64_t)H * D + 4 * H + 1;
  if (backward) {
    GD3_CHECK_CUDA(cudaMemsetAsync(grad_params, 0, sizeof(float) * nparam, stream));
    // grad_feats is written in full by the d feats GEMM below (K > 0); it only needs clearing for K == 0,
    // where it is empty anyway
  }
  if (K == 0) return GD3_OK;
  GD3_REQUIRE(feats && depths && W1 && b1 && gamma && beta && w2 && b2, "gd3_depth_head_loss: null input");
  RankWorkspace w = carve_rank(workspace, S, K, D, backward, l1);
  if (!workspace || workspace_bytes < w.total) {
    set_error("gd3_depth_head_loss: workspace too small (%zu < %zu)", workspace_bytes, w.total);
    return GD3_ERR_WORKSPACE;
  }
  const int64_t R = S * K;
  int rc;
  // ---- u = f W1^T on the tensor cores (split bf16, 3 K-concatenated panels) ----
  {
    // operands of u = f W1^T; the backward GEMMs read the same panels MN-major (no transposed copies)
    // centring group: the two sets of an image pair when the L1 term couples them, else one set (rank_group_mean)
    const int group_rows = (int)((l1 ? 2 : 1) * K);
    {
      dim3 grid((unsigned)(R / group_rows), (unsigned)ceil_div<int64_t>(D, 64));
      GD3_PROF("rank_group_mean", stream);
      rank_group_mean<<<grid, 1024, 0, stream>>>(feats, group_rows, (int)D, w.mu);
    }
    GD3_CHECK_LAUNCH();
    if ((rc = launch_split3("split3_feats", feats, R, (int)D, w.ldd, 2, w.F3, stream, w.mu, group_rows)))
      return rc;
    if ((rc = launch_split3("split3_w1", W1, H, (int)D, w.ldd, 1, w.W3, stream))) return rc;
  }
  {
    CUtensorMap ta, tb;
    if ((rc = tc::make_tmap_bf16(&ta, w.F3, 3 * (int64_t)w.ldd, R, 1, 3 * (int64_t)w.ldd, 0, tc::BM))) return rc;
    if ((rc = tc::make_tmap_bf16(&tb, w.W3, 3 * (int64_t)w.ldd, H, 1, 3 * (int64_t)w.ldd, 0, 128))) return rc;
    tc::EpiStoreF32::Params ep{w.u, (int)R, H, H, 0, 1.0f, nullptr};
    if ((rc = tc::enable_tma_store(ep, 1))) return rc;
    tc::GemmShape s{(int)R, H, 3 * w.ldd, 1};
    if ((rc = tc::launch_gemm<128, 8, tc::EpiStoreF32>("rank_u_gemm", ta, tb, s, ep, stream))) return rc;
  }
  // ---- LayerNorm 1 / sigma of every ordered pair (Gram matrix of the centred rows, per set) ----
  {
    {
      GD3_PROF("rank_gram_prep", stream);
      rank_gram_prep<<<(unsigned)ceil_div<int64_t>(R, 8), 256, 0, stream>>>(w.u, b1, R, w.Wb3, w.Va3, w.nb, w.na);
    }
    GD3_CHECK_LAUNCH();
    tc::GemmShape s{(int)K, (int)K, 3 * H, (int)S};
    const int bn = tc::pick_tile_n(s);
    CUtensorMap ta, tb;
    if ((rc = tc::make_tmap_bf16(&ta, w.Wb3, 3 * H, K, S, 3 * H, K * 3 * (int64_t)H, tc::BM))) return rc;
    if ((rc = tc::make_tmap_bf16(&tb, w.Va3, 3 * H, K, S, 3 * H, K * 3 * (int64_t)H, bn))) return rc;
    EpiRstd::Params ep{w.rstd, (int)K, w.nb, w.na, ln_eps};
    if (K % 4 == 0 && reinterpret_cast<uintptr_t>(w.rstd) % 16 == 0) {
      if ((rc = tc::make_tmap_store32(&ep.tm_out, w.rstd, K, K, S, K, K * K))) return rc;
      ep.use_tma = 1;
    }
    if (bn == 256) rc = tc::launch_gemm<256, 8, EpiRstd>("rank_rstd_gemm", ta, tb, s, ep, stream);
    else if (bn == 192) rc = tc::launch_gemm<192, 8, EpiRstd>("rank_rstd_gemm", ta, tb, s, ep, stream);
    else rc = tc::launch_gemm<128, 8, EpiRstd>("rank_rstd_gemm", ta, tb, s, ep, stream);
    if (rc) return rc;
  }
  // ---- valid-pair counts ----
  GD3_CHECK_CUDA(cudaMemsetAsync(w.count, 0, sizeof(int) * S, stream));
  GD3_CHECK_CUDA(cudaMemsetAsync(w.loss_sum, 0, sizeof(double) * S, stream));
  GD3_CHECK_CUDA(cudaMemsetAsync(w.l1_sum, 0, sizeof(double) * S, stream));
  {
    dim3 grid((unsigned)ceil_div<int64_t>(K, 256), (unsigned)S);
    {
      GD3_PROF("rank_count", stream);
      rank_count<<<grid, 256, sizeof(float) * K, stream>>>(depths, (int)K, mode, thr, w.count);
    }
    GD3_CHECK_LAUNCH();
    {
      GD3_PROF("rank_inv_count", stream);
      rank_inv_count<<<(unsigned)ceil_div<int64_t>(S, 128), 128, 0, stream>>>(w.count, (int)S, joint_mean, w.inv_count);
    }
    GD3_CHECK_LAUNCH();
  }
  RankParams rp{};
  rp.u = w.u;
  rp.depth = depths;
  rp.b1 = b1;
  rp.gamma = gamma;
  rp.beta = beta;
  rp.w2 = w2;
  rp.b2 = b2;
  rp.rstd = w.rstd;
  rp.inv_count = w.inv_count;
  rp.w_rank = w_rank;
  rp.K = (int)K;
  rp.S = (int)S;
  rp.b_per_warp = w.bpw;
  rp.mode = mode;
  rp.use_tanh = use_tanh;
  rp.thr = thr;
  rp.margin = margin;
  rp.ln_eps = ln_eps;
  rp.loss_sum = w.loss_sum;
  rp.dub_part = w.dub_part;
  rp.dua_part = w.dua_part;
  rp.gparam = grad_params;
  rp.gparam_off = (int64_t)H * D;
  {
    const size_t smem = sizeof(float) * (2 * TILE_A * H + TILE_A + 3 * H + 4);
    dim3 grid((unsigned)w.TA, (unsigned)w.TB, (unsigned)S);
#define GD3_RANK_LAUNCH(G, M, T)                                            \
  do {                                                                     \
    static SmemOptIn opt;                                                  \
    GD3_CHECK_CUDA(opt.ensure(rank_pairs<G, M, T>, smem));                 \
    GD3_PROF("rank_pairs", stream);                                        \
    rank_pairs<G, M, T><<<grid, WARPS * 32, smem, stream>>>(rp);           \
  } while (0)
#define GD3_RANK_LAUNCH_T(G, M)                 \
  do {                                          \
    if (use_tanh) GD3_RANK_LAUNCH(G, M, true);  \
    else GD3_RANK_LAUNCH(G, M, false);          \
  } while (0)
    if (backward) {
      if (mode == 0) GD3_RANK_LAUNCH_T(true, 0);
      else GD3_RANK_LAUNCH_T(true, 1);
    } else {
      if (mode == 0) GD3_RANK_LAUNCH_T(false, 0);
      else GD3_RANK_LAUNCH_T(false, 1);
    }
#undef GD3_RANK_LAUNCH_T
#undef GD3_RANK_LAUNCH
    GD3_CHECK_LAUNCH();
  }
  if (l1) {
    if (backward) GD3_CHECK_CUDA(cudaMemsetAsync(w.du_extra, 0, sizeof(float) * R * H, stream));
    dim3 grid((unsigned)ceil_div<int64_t>(K, 16), (unsigned)(S / 2));
    if (backward)
      {
        GD3_PROF("rank_l1", stream);
        rank_l1<true><<<grid, 256, 0, stream>>>(rp, w_l1, w.l1_sum, w.du_extra);
      }
    else
      {
        GD3_PROF("rank_l1", stream);
        rank_l1<false><<<grid, 256, 0, stream>>>(rp, w_l1, w.l1_sum, nullptr);
      }
    GD3_CHECK_LAUNCH();
  }
  {
    GD3_PROF("rank_finalize", stream);
    rank_finalize<<<(unsigned)ceil_div<int64_t>(S, 128), 128, 0, stream>>>(w.loss_sum, w.inv_count, w.l1_sum, (int)S,
                                                                        (int)K, loss_rank, l1 ? loss_l1 : nullptr);
  }
  GD3_CHECK_LAUNCH();
  if (!backward) return GD3_OK;
  // ---- gradients: du, then d feats = du W1 and d W1 = du^T f on the tensor cores ----
  {
    dim3 grid((unsigned)ceil_div<int64_t>(K, 32), (unsigned)S);
    {
      GD3_PROF("rank_reduce_du", stream);
      rank_reduce_du<<<grid, 256, 0, stream>>>(w.dub_part, w.dua_part, l1 ? w.du_extra : nullptr, (int)S, (int)K, w.TA,
                                             w.TB, w.du2, grad_params + (int64_t)H * D);
    }
    GD3_CHECK_LAUNCH();
  }
  {
    // d feats = du W1: A = du hi panel (K-major, row stride 2 H), B = W1 hi panel of W3 read MN-major ([k = h][mn = d])
    CUtensorMap t_du, t_w1;
    if ((rc = tc::make_tmap_bf16(&t_du, w.du2, H, R, 1, 2 * H, 0, tc::BM))) return rc;
    if ((rc = tc::make_tmap_bf16(&t_w1, w.W3, D, H, 1, 3 * (int64_t)w.ldd, 0, 64))) return rc;
    tc::EpiStoreF32::Params e1{grad_feats, (int)R, (int)D, D, 0, 1.0f, nullptr};
    if ((rc = tc::enable_tma_store(e1, 1))) return rc;
    tc::GemmShape s1{(int)R, (int)D, H, 1};
    if ((rc = tc::launch_gemm<256, 8, tc::EpiStoreF32, false, true>("rank_df_gemm", t_du, t_w1, s1, e1, stream))) return rc;
    // d W1 (H x D) = sum over sets of du_s^T f_s with the 3-term bf16 split  hi^T hi + lo^T hi + hi^T lo: a grouped
    // contraction over (term, set of the group, 64-keypoint block), both operands read MN-major from their row-major
    // panels (du2 = [hi | lo], F3 = [hi | hi | lo]); one GEMM batch entry per group, fp32 atomic accumulation
    CUtensorMap t_dum, t_fm;
    if ((rc = tc::make_tmap_bf16(&t_dum, w.du2, 2 * H, K, S, 2 * H, K * 2 * (int64_t)H, 64))) return rc;
    if ((rc = tc::make_tmap_bf16(&t_fm, w.F3, 3 * (int64_t)w.ldd, K, S, 3 * (int64_t)w.ldd, K * 3 * (int64_t)w.ldd, 64)))
      return rc;
    tc::GroupedK gk;
    gk.panels = 3;
    gk.sets_per_group = w.gs;
    gk.row_blocks = (int)ceil_div<int64_t>(K, tc::BK);
    gk.a_off[0] = 0; gk.a_off[1] = H; gk.a_off[2] = 0;
    gk.b_off[0] = 0; gk.b_off[1] = 0; gk.b_off[2] = 2 * w.ldd;
    EpiAtomicAddF32::Params e2{grad_params, H, (int)D, D};
    tc::GemmShape s2{H, (int)D, gk.panels * gk.sets_per_group * gk.row_blocks * tc::BK, w.groups};
    if ((rc = tc::launch_gemm<256, 8, EpiAtomicAddF32, true, true, true>("rank_dw1_gemm", t_dum, t_fm, s2, e2, stream, 0, gk)))
      return rc;
  }
  return GD3_OK;
}

}  // extern "C"
