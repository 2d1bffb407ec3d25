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
// Layout of the pair kernel (rank_pairs): a warp owns one b row and evaluates 4 pairs per step, every lane 4 of the
// 128 hidden units of each; the per-pair sums are 32-lane shuffle reductions arranged so that one scalar chain per
// lane serves 4 pairs, and the loop is software-pipelined (forward of step t + 1 under the chain of step t).  A CTA
// owns a 64 x (4 b_per_warp) tile of (a, b): every warp keeps the gradient of its b row in registers and accumulates
// the a side into a shared tile; the a index is staggered per warp so no two warps touch the same row in the same step.
#include <algorithm>
#include <cstdlib>
#include <map>
#include <mutex>
#include <utility>
#include <vector>
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
// side.  Two 4-warp CTAs per SM: the pipelined kernel holds two steps of forward state and wants ~224 registers.
constexpr int TILE_A = 64;
constexpr int WARPS = 4;
constexpr int B_PER_WARP_MAX = 32;                   // b rows a warp walks per CTA: 2 .. 32, chosen per problem (rank_b_per_warp)
constexpr int CTAS_PER_SM = 2;

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
  float* du;             // (S, K, H), zero on entry: d u_b rows and -(d u_a) tiles are added with vector reductions
  float* gparam;         // packed parameter gradients: [W1 (H*D) | b1 | gamma | beta | w2 | b2]
  int64_t gparam_off;    // offset of b1 inside gparam (= H * D)
};

// ---- pair kernel: elementwise part of one ordered pair (a -> b) --------------------------------------------------
// Differences to head_eval (kept for the L1 kernel): the Gaussian density is folded into the exponent
// (e' = pdf-scaled exp(-y^2/2) = ex2(-ys^2 + log2 kPdf), the erf polynomial is pre-divided by kPdf, so GELU' is one
// fma), and the LayerNorm-backward mean term m1 is not formed at all: sum_pairs alpha m1 = mean_h(sum_pairs alpha q),
// so the accumulated gradient rows are centred once in rank_reduce_du instead of once per pair.
// 22 packed fp32 ops per float2 forward + backward.
constexpr float kLog2Pdf = -1.0901312512086083f;          // log2(kPdf)
constexpr float kN1 = -2.f * 0.37046028286393695f;       // -a1 / kPdf   (A&S 7.1.25: a1, a2, a3)
constexpr float kN2 = 2.f * 0.10206088492966209f;
constexpr float kN3 = -2.f * 0.7960676214969513f;

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

// ---- pair kernel: whole warp per pair, four a rows per lane and step, software-pipelined ---------------------------
// Measured history at cfg2 (64 sets x 512 keypoints; tools/time_rank.py), all with identical results:
//   16 lanes per pair, 8 hidden units per lane, 1 pair per lane, 16 warps x 128 registers          2.91 ms  (round 2 start)
//   same, 2 pairs per lane and one shared scalar chain ("quad step"), 12 warps x 168 registers      2.64 ms
//   32 lanes per pair, 4 hidden units per lane, 4 pairs per lane (this layout), 12 warps x 168      2.63 ms
//   this layout, software-pipelined, 8 warps x 224 registers                                       2.44 ms
//   the same with the d u flush below                                                             2.39 ms
//   software-pipelined with 2 pairs per lane and step (half the stage state), 12 warps x 168      2.60 ms
// The register budget decides the layout.  Per lane, the state that lives across the scalar chain is 3 values
// (xh, GELU, GELU') per (pair, hidden unit) = 12 x pairs-per-warp-step registers whatever the layout, while the
// per-hidden-unit constants and accumulators (gamma, beta, w2 (x2), d gamma, d beta, d w2, v_b, d u_b: 9 values) scale
// with the hidden units a lane owns.  So a lane owns 4 hidden units (36 persistent registers instead of 72), a pair is
// spread over all 32 lanes, and a step covers a QUAD slot (4 rows of the a tile, one b row): 8 independent packed
// chains per lane in the forward.  Lane l names the rows of the slot X_j = row (j ^ p), p = bits 4..3 of l, so the
// first two levels of the 32-lane reduction are "keep X0 (X1), send X2 (X3)" and "keep X0, send X1" for every lane (no
// selects); three butterfly levels finish all four pairs at once: 12 shuffles for 4 pairs.  Each lane then runs the
// scalar chain (tanh, logistic / hinge, validity, loss) for ITS X0 pair only -- one chain per 4 pairs -- and fetches
// (d out, m2) of X1..X3 from lanes l ^ 8, l ^ 16, l ^ 24.
// Timing experiments with parts removed (results wrong, 12 warps): no barrier -2 %, no a-side read-modify-write -5 %, no
// parameter sums (-17 % of the FMA work) -5 %, no MUFU in the forward -2 %, forward arithmetic removed altogether
// 1.82 ms, and forward + barrier + RMW + parameter sums removed still 1.42 ms (1.22 ms with 16 warps): the skeleton --
// operand loads, 22 shuffles, the scalar chain and the d u accumulation -- is a dependent chain of ~1200 cycles per
// step, which is why hiding it (pipelining) pays and removing arithmetic does not.
constexpr int HPW = 4;                               // hidden units per lane
constexpr int NPW = HPW / 2;                         // ... as float2
constexpr int QSLOTS = TILE_A / 4;                   // ring of quad slots
constexpr int SPACING_W = QSLOTS / WARPS;            // ring distance between consecutive warps
static_assert(QSLOTS % WARPS == 0 && SPACING_W >= 2, "stagger needs at least 2 quad slots between warps");
static_assert((QSLOTS & (QSLOTS - 1)) == 0, "the ring offset wraps with a mask");

struct HeadConstW {
  F2 gs[NPW], bs[NPW], w2c[NPW], w2g[NPW];      // as HeadConst, hidden units 4 lane .. 4 lane + 3
  float b2;
  __device__ __forceinline__ void load(const float* gamma, const float* beta, const float* w2, const float* b2p, int lane) {
#pragma unroll
    for (int i = 0; i < NPW; ++i) {
      const int h0 = 4 * lane + 2 * i, h1 = h0 + 1;
      gs[i] = make_float2(gamma[h0] * kC, gamma[h1] * kC);
      bs[i] = make_float2(beta[h0] * kC, beta[h1] * kC);
      w2c[i] = make_float2(w2[h0] * (1.f / kC), w2[h1] * (1.f / kC));
      w2g[i] = make_float2(w2[h0] * gamma[h0], w2[h1] * gamma[h1]);
    }
    b2 = b2p[0];
  }
};
struct PairFwdW {
  F2 xh[NPW], g[NPW], gp[NPW];
  float acc, m2;
};
// pair_forward for NPW float2 per lane (same arithmetic, same constants)
template <bool GRAD>
__device__ __forceinline__ void pair_forward_w(const F2 (&hcv)[NPW], float rstd, const HeadConstW& hc, PairFwdW& o) {
  const F2 r2 = bc(rstd);
  F2 acc2 = bc(0.f), m2_2 = bc(0.f);
#pragma unroll
  for (int i = 0; i < NPW; ++i) {
    const F2 xh = mul2(hcv[i], r2);
    const F2 ys = fma2(xh, hc.gs[i], hc.bs[i]);
    const F2 ti = fma2(bc(kErfP), make_float2(fabsf(ys.x), fabsf(ys.y)), bc(1.f));
    // (one reciprocal per float2 via 1 / a = b / (a b) was measured: -8 MUFU, +12 FMA-pipe cycles per step, 2 % slower)
    const F2 t = make_float2(fast_rcp(ti.x), fast_rcp(ti.y));
    F2 np = fma2(t, bc(kN3), bc(kN2));
    np = fma2(t, np, bc(kN1));
    np = mul2(np, t);
    const F2 sq = fma2(ys, ys, bc(-kLog2Pdf));
    const F2 e = make_float2(fast_ex2(-sq.x), fast_ex2(-sq.y));   // kPdf exp(-y^2 / 2)
    const F2 ea = fma2(np, e, bc(1.f));
    const F2 phi = fma2(bc(0.5f), make_float2(copysignf(ea.x, ys.x), copysignf(ea.y, ys.y)), bc(0.5f));
    const F2 g = mul2(ys, phi);
    acc2 = fma2(hc.w2c[i], g, acc2);
    o.xh[i] = xh;
    o.g[i] = g;
    if (GRAD) {
      const F2 gp = fma2(ys, e, phi);
      o.gp[i] = gp;
      m2_2 = fma2(hc.w2g[i], mul2(gp, xh), m2_2);
    }
  }
  o.acc = acc2.x + acc2.y;
  o.m2 = m2_2.x + m2_2.y;
}

// ---- software pipeline: the reduction + scalar chain of slot t overlap the forward of slot t + 1 ----
// Per step a warp has ~280 instructions of independent packed arithmetic (forward, backward) and a serial section of
// ~55 instructions (5 shuffle levels, 4 dependent MUFU levels: ~330 cycles of latency).  With 3-4 warps per scheduler
// the serial sections are not hidden (the unpipelined kernel: FMA pipe 56 % busy).  Here the loop body is
//     R(t) | F(t + 1) | B(t)
// in ONE basic block: the forward of the next quad slot has no dependence on the chain of the current one, so the
// scheduler fills the chain's latency with it.  Two stages of forward state (2 x 54 registers) are live, so the kernel
// runs 8 warps per SM with up to 255 registers.  Control flow inside the walk depends on blockIdx / loop counters only
// (a b row beyond the set is clamped and made invalid through a NaN depth; a quad slot beyond the set is evaluated on
// zero rows with NaN depths): the compiler emits no divergence checks (BRA.DIV) around the shuffles, which would split
// the block.  Near-duplicate pairs (negative rstd from the Gram epilogue) are repaired by rank_fix_rstd beforehand.
constexpr int kTripP = 2;      // slots per trip of rank_pairs (= its barrier period; 4 measured equal)
static_assert(kTripP % 2 == 0 && kTripP <= SPACING_W && QSLOTS % kTripP == 0, "trip length");
struct StageW {
  PairFwdW o[4];
  float rs[4];
  float dd;
  uint32_t ad0;
};
struct ScalW {
  float d[4], al[4], nn[4];
};

template <bool GRAD, int MODE, bool TANH>
__global__ void __launch_bounds__(WARPS * 32, CTAS_PER_SM) rank_pairs(RankParams p) {
  extern __shared__ __align__(16) float smem_raw[];
  const uint32_t s_raw = (uint32_t)__cvta_generic_to_shared(smem_raw);
  float* va = smem_raw + (((2048u - (s_raw & 2047u)) & 2047u) >> 2);
  float* dua = va + TILE_A * H;
  float* da = dua + TILE_A * H;
  float* red = da + TILE_A;
  const int set = blockIdx.z, ta = blockIdx.x, tb = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pq = (lane >> 3) & 3;
  const int K = p.K;
  const float* U = p.u + (int64_t)set * K * H;
  const float* Dp = p.depth + (int64_t)set * K;

  for (int r = warp; r < TILE_A; r += WARPS) {
    const int a = ta * TILE_A + r;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (a < K) v = *reinterpret_cast<const float4*>(U + (int64_t)a * H + 4 * lane);
    float m = (v.x + v.y) + (v.z + v.w);
    m = warp_sum(m) * (1.f / H);
    *reinterpret_cast<float4*>(va + r * H + 4 * lane) = make_float4(m - v.x, m - v.y, m - v.z, m - v.w);   // negated
    if (GRAD) *reinterpret_cast<float4*>(dua + r * H + 4 * lane) = make_float4(0.f, 0.f, 0.f, 0.f);
    if (lane == 0) da[r] = (a < K) ? Dp[a] : __int_as_float(0x7fc00000);
  }
  HeadConstW hc;
  float4 b1v;
  {
    b1v = make_float4(p.b1[4 * lane], p.b1[4 * lane + 1], p.b1[4 * lane + 2], p.b1[4 * lane + 3]);
    const float b1_mean = warp_sum((b1v.x + b1v.y) + (b1v.z + b1v.w)) * (1.f / H);
    b1v.x -= b1_mean; b1v.y -= b1_mean; b1v.z -= b1_mean; b1v.w -= b1_mean;
    hc.load(p.gamma, p.beta, p.w2, p.b2, lane);
  }
  const float inv_cnt = p.inv_count[set] * (p.w_rank ? p.w_rank[set] : 1.f);
  __syncthreads();

  F2 dgam[NPW], dbet[NPW], dw2[NPW];
  float db2 = 0.f;
#pragma unroll
  for (int i = 0; i < NPW; ++i) dgam[i] = dbet[i] = dw2[i] = bc(0.f);
  float loss_local = 0.f;
  constexpr uint32_t kRowBytes = H * 4, kQslotBytes = 4 * kRowBytes, kRingBytes = QSLOTS * kQslotBytes, kDuaOff = TILE_A * H * 4;
  const uint32_t va_x0 = opaque((uint32_t)__cvta_generic_to_shared(va) + pq * kRowBytes + lane * 16);
  const uint32_t da_x0 = opaque((uint32_t)__cvta_generic_to_shared(da) + pq * 4);
  uint32_t soff = opaque((uint32_t)(SPACING_W * warp) * kQslotBytes);
  const int a_tile0 = ta * TILE_A;

  auto load_rstd = [&](int b_, float& x0, float& x1) {
    const float* rrow = p.rstd + ((int64_t)set * K + b_) * K + a_tile0;
    x0 = (b_ < K && a_tile0 + lane < K) ? __ldg(rrow + lane) : 1.f;
    x1 = (b_ < K && a_tile0 + lane + 32 < K) ? __ldg(rrow + lane + 32) : 1.f;
  };
  float rs0_next, rs1_next;
  const int tile_b = WARPS * p.b_per_warp;
  load_rstd(tb * tile_b + warp, rs0_next, rs1_next);
  for (int bi = 0; bi < p.b_per_warp; ++bi) {
    if (tb * tile_b + bi * WARPS >= K) break;       // CTA-uniform
    const int b = tb * tile_b + bi * WARPS + warp;
    const bool b_ok = b < K;
    const int b_ld = b_ok ? b : K - 1;              // a row beyond the set: evaluated on a valid row, every pair invalid
    F2 vb[NPW], dub[NPW];
    float d_b;
    {
      const float4 u0 = *reinterpret_cast<const float4*>(U + (int64_t)b_ld * H + 4 * lane);
      d_b = b_ok ? Dp[b_ld] : __int_as_float(0x7fc00000);
      const float m = warp_sum((u0.x + u0.y) + (u0.z + u0.w)) * (1.f / H);
      vb[0] = make_float2((u0.x - m) + b1v.x, (u0.y - m) + b1v.y);
      vb[1] = make_float2((u0.z - m) + b1v.z, (u0.w - m) + b1v.w);
      dub[0] = dub[1] = bc(0.f);
    }
    const float rs0 = rs0_next, rs1 = rs1_next;
    if (bi + 1 < p.b_per_warp) load_rstd(b + WARPS, rs0_next, rs1_next);

    // F: forward of the quad slot at ring offset so
    auto fwd = [&](StageW& st, uint32_t so) {
      const uint32_t row0 = (so >> 9) + pq;
      st.dd = d_b - lds32(da_x0 + (so >> 7));
      const float rsel = (so < 32 * kRowBytes) ? rs0 : rs1;
#pragma unroll
      for (int j = 0; j < 4; ++j) st.rs[j] = __shfl_sync(0xffffffffu, rsel, row0 ^ j);
      st.ad0 = va_x0 + so;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 x = lds128(st.ad0 ^ (j * kRowBytes));
        F2 h[NPW];
        h[0] = add2(vb[0], make_float2(x.x, x.y));
        h[1] = add2(vb[1], make_float2(x.z, x.w));
        pair_forward_w<GRAD>(h, st.rs[j], hc, st.o[j]);
      }
    };
    // R: 32-lane sums, scalar chain of this lane's X0 pair, exchange of (d out, m2)
    auto chain = [&](const StageW& st, ScalW& sc_) {
      F2 am0 = make_float2(st.o[0].acc, st.o[0].m2), am1 = make_float2(st.o[1].acc, st.o[1].m2);
      {
        F2 r0, r1;
        r0.x = __shfl_xor_sync(0xffffffffu, st.o[2].acc, 16);
        r1.x = __shfl_xor_sync(0xffffffffu, st.o[3].acc, 16);
        r0.y = GRAD ? __shfl_xor_sync(0xffffffffu, st.o[2].m2, 16) : 0.f;
        r1.y = GRAD ? __shfl_xor_sync(0xffffffffu, st.o[3].m2, 16) : 0.f;
        am0 = add2(am0, r0);
        am1 = add2(am1, r1);
        r0.x = __shfl_xor_sync(0xffffffffu, am1.x, 8);
        r0.y = GRAD ? __shfl_xor_sync(0xffffffffu, am1.y, 8) : 0.f;
        am0 = add2(am0, r0);
      }
#pragma unroll
      for (int off = 4; off > 0; off >>= 1) {
        F2 other;
        other.x = __shfl_xor_sync(0xffffffffu, am0.x, off);
        other.y = GRAD ? __shfl_xor_sync(0xffffffffu, am0.y, off) : 0.f;
        am0 = add2(am0, other);
      }
      const float dd = st.dd;
      const bool valid = (MODE == 0) ? (fabsf(dd) > p.thr) : (fabsf(tanhf(dd)) > p.thr);
      const float acc = am0.x + hc.b2;
      const float sc = TANH ? fast_tanh(acc) : acc;
      const float ssc = __uint_as_float(__float_as_uint(sc) ^ (__float_as_uint(dd) & 0x80000000u));
      float l, dl;
      if (MODE == 0) {
        const float ex = fast_ex2(ssc * -1.4426950408889634f);
        l = __logf(1.f + ex);
        dl = -ex * fast_rcp(1.f + ex);
      } else {
        const float m = p.margin - ssc;
        l = fmaxf(m, 0.f);
        dl = (m > 0.f) ? -1.f : 0.f;
      }
      loss_local += valid ? l : 0.f;
      if (GRAD) {
        const float dsg = __uint_as_float(__float_as_uint(dl) ^ (__float_as_uint(dd) & 0x80000000u));
        const float dout0 = valid ? dsg * (TANH ? fmaf(-sc, sc, 1.f) : 1.f) * inv_cnt : 0.f;
        db2 += dout0;
        sc_.d[0] = dout0;
        sc_.al[0] = dout0 * st.rs[0];
        sc_.nn[0] = am0.y * (-1.f / H);
#pragma unroll
        for (int j = 1; j < 4; ++j) {
          const float dj = __shfl_xor_sync(0xffffffffu, dout0, 8 * j);
          const float mj = __shfl_xor_sync(0xffffffffu, am0.y, 8 * j);
          sc_.d[j] = dj;
          sc_.al[j] = dj * st.rs[j];
          sc_.nn[j] = mj * (-1.f / H);
        }
      }
    };
    // B: parameter sums, d u_b, read-modify-write of the a-side rows
    auto bwd = [&](const StageW& st, const ScalW& sc_) {
      F2 tq[4][NPW];
#pragma unroll
      for (int i = 0; i < NPW; ++i) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const F2 dj = bc(sc_.d[j]);
          dw2[i] = fma2(dj, st.o[j].g[i], dw2[i]);
          dbet[i] = fma2(dj, st.o[j].gp[i], dbet[i]);
          dgam[i] = fma2(dj, mul2(st.o[j].gp[i], st.o[j].xh[i]), dgam[i]);
          tq[j][i] = fma2(bc(sc_.nn[j]), st.o[j].xh[i], mul2(hc.w2g[i], st.o[j].gp[i]));
          dub[i] = fma2(bc(sc_.al[j]), tq[j][i], dub[i]);
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint32_t ad = (st.ad0 ^ (j * kRowBytes)) + kDuaOff;
        const float4 c = lds128(ad);
        const F2 s0 = fma2(bc(sc_.al[j]), tq[j][0], make_float2(c.x, c.y));
        const F2 s1 = fma2(bc(sc_.al[j]), tq[j][1], make_float2(c.z, c.w));
        sts128(ad, make_float4(s0.x, s0.y, s1.x, s1.y));
      }
    };

    // One trip = kTripP slots between two CTA barriers (kTripP <= ring spacing of the warps).  Inside a trip nothing
    // separates the stages, so the chain of slot t + 1 (which needs the forward of t + 1 only) also overlaps the
    // backward of slot t.  The last trip of a row has no next slot to forward: it is a second copy of the body.
    StageW sA, sB;
    fwd(sA, soff);
    soff = (soff + kQslotBytes) & (kRingBytes - 1);
    auto trip = [&](auto last_tag) {
      constexpr bool LAST = decltype(last_tag)::value;
#pragma unroll
      for (int k = 0; k < kTripP; k += 2) {
        {
          ScalW sc_;
          chain(sA, sc_);
          fwd(sB, soff);
          soff = (soff + kQslotBytes) & (kRingBytes - 1);
          if (GRAD) bwd(sA, sc_);
        }
        {
          ScalW sc_;
          chain(sB, sc_);
          if (!(LAST && k + 2 == kTripP)) {
            fwd(sA, soff);
            soff = (soff + kQslotBytes) & (kRingBytes - 1);
          }
          if (GRAD) bwd(sB, sc_);
        }
      }
      if (GRAD) __syncthreads();
    };
#pragma unroll 1
    for (int tt = 0; tt < QSLOTS / kTripP - 1; ++tt) trip(std::false_type{});
    trip(std::true_type{});
    if (GRAD && b_ok) {
      // this b row's gradient over the a tile goes straight into du (red.global.add.v4.f32: one 512-byte row per warp)
      atomicAdd(reinterpret_cast<float4*>(p.du + ((int64_t)set * K + b) * H + 4 * lane),
                make_float4(dub[0].x, dub[0].y, dub[1].x, dub[1].y));
    }
  }
  // ---- CTA-level reductions ----
  loss_local = warp_sum(loss_local) * (1.f / 8.f);
  if (lane == 0 && loss_local != 0.f) atomicAdd(p.loss_sum + set, (double)loss_local);
  if (GRAD) {
    __syncthreads();
    // the a tile's accumulated gradient flows to -u_a
    for (int e = threadIdx.x; e < TILE_A * (H / 4); e += blockDim.x) {
      const int r = e / (H / 4), a = ta * TILE_A + r;
      if (a < K) {
        const float4 v = *reinterpret_cast<const float4*>(dua + 4 * e);
        atomicAdd(reinterpret_cast<float4*>(p.du + ((int64_t)set * K + a) * H + 4 * (e - r * (H / 4))),
                  make_float4(-v.x, -v.y, -v.z, -v.w));
      }
    }
    // red: [b1 | gamma | beta | w2 | b2], the layout of the parameter gradient behind W1.
    // d b1 = sum over pairs of the centred d h = column sums of the a tile (the same pairs as the b rows), centred over
    // h: sum_pairs alpha m1 is the h-mean of the accumulated rows (rank_reduce_du), and centring is linear
    static_assert(WARPS * 32 == H, "one thread per hidden unit for the b1 column sums");
    {
      float c = 0.f;
#pragma unroll 8
      for (int r = 0; r < TILE_A; ++r) c += dua[r * H + threadIdx.x];
      const float cm = warp_sum(c);
      if (lane == 0) red[4 * H + 1 + warp] = cm;
      __syncthreads();
      const float mean = ((red[4 * H + 1] + red[4 * H + 2]) + (red[4 * H + 3] + red[4 * H + 4])) * (1.f / H);
      __syncthreads();      // every thread has read the warp sums before `red` is rewritten
      red[threadIdx.x] = c - mean;
    }
    for (int e = H + threadIdx.x; e < 4 * H + 1; e += blockDim.x) red[e] = 0.f;
    __syncthreads();
#pragma unroll
    for (int j = 0; j < HPW; ++j) {
      const int i = j >> 1;
      const bool hi = j & 1;
      const float w2h = (hi ? hc.w2c[i].y : hc.w2c[i].x) * kC;
      const int h = 4 * lane + j;
      atomicAdd(red + H + h, (hi ? dgam[i].y : dgam[i].x) * w2h);
      atomicAdd(red + 2 * H + h, (hi ? dbet[i].y : dbet[i].x) * w2h);
      atomicAdd(red + 3 * H + h, (hi ? dw2[i].y : dw2[i].x) * (1.f / kC));
    }
    db2 = warp_sum(db2) * (1.f / 8.f);
    if (lane == 0) atomicAdd(red + 4 * H, db2);
    __syncthreads();
    float* gp = p.gparam + p.gparam_off;
    for (int e = threadIdx.x; e < 4 * H + 1; e += blockDim.x)
      if (red[e] != 0.f) atomicAdd(gp + e, red[e]);
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

// the same for D % 4 == 0 and 16-byte aligned rows: a thread owns 4 channels (one 128-bit load per row) and keeps 4 row
// loads in flight -- 37 -> ~20 us at cfg2 (100 MB).  grid (groups, ceil(D / 64)), block 256 = 16 row slices x 16 quads
__global__ void __launch_bounds__(256) rank_group_mean_v4(const float* __restrict__ f, int rows, int D, float* __restrict__ mu) {
  __shared__ float4 part[16][16];
  const int g = blockIdx.x, ql = threadIdx.x & 15, c = blockIdx.y * 64 + 4 * ql, slice = threadIdx.x >> 4;
  float4 s[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) s[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (c < D) {
    const float* col = f + (int64_t)g * rows * D + c;
    int r = slice;
    for (; r + 48 < rows; r += 64) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(col + (int64_t)(r + 16 * i) * D));
        s[i].x += v.x; s[i].y += v.y; s[i].z += v.z; s[i].w += v.w;
      }
    }
    for (; r < rows; r += 16) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(col + (int64_t)r * D));
      s[0].x += v.x; s[0].y += v.y; s[0].z += v.z; s[0].w += v.w;
    }
  }
  part[slice][ql] = make_float4((s[0].x + s[1].x) + (s[2].x + s[3].x), (s[0].y + s[1].y) + (s[2].y + s[3].y),
                                (s[0].z + s[1].z) + (s[2].z + s[3].z), (s[0].w + s[1].w) + (s[2].w + s[3].w));
  __syncthreads();
  if (slice == 0 && c < D) {
    float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const float4 v = part[k][ql];
      t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w;
    }
    const float inv = 1.f / (float)rows;
    *reinterpret_cast<float4*>(mu + (int64_t)g * D + c) = make_float4(t.x * inv, t.y * inv, t.z * inv, t.w * inv);
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
    int* rowflag;        // (S, K), zero on entry: number of flagged pairs of every b row (rank_fix_rstd repairs those rows)
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
    int nbad = 0;      // flagged pairs of this thread's row (a column past K reads as zero and is never flagged)
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
        const bool bad = ss < kGramMinRatio * nsum;
        nbad += bad ? 1 : 0;
        v[q] = bad ? -1.f : rsqrtf(fmaf(ss, 1.f / H, p.eps));
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
    if (nbad && cx.m0 + cx.row < p.K) atomicAdd(p.rowflag + (int64_t)cx.b * p.K + cx.m0 + cx.row, nbad);
  }
};

// Repairs the pairs the Gram epilogue flagged (rstd < 0: near-duplicate rows, where |w|^2 + |v|^2 - 2 w.v cancels): the
// sum of squares of h_c = w_b - v_a is taken directly from the fp32 rows.  One warp per b row.  The diagonal pair
// (a == b: h_c is the centred b1 alone) is flagged in nearly every row but never valid (D = 0; rank_pairs multiplies its
// rstd by a zero weight), so a row whose only flag is its diagonal returns at once: the kernel costs a launch and one
// pass over the row flags when nothing else is flagged (the usual case).  This keeps every branch out of the pair loop.
__global__ void __launch_bounds__(256) rank_fix_rstd(const float* __restrict__ u, const float* __restrict__ b1,
                                                     const int* __restrict__ rowflag, int64_t R, int K, float eps,
                                                     float* __restrict__ rstd) {
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= R) return;
  const int64_t set = r / K;
  {
    const int nflag = rowflag[r];
    if (nflag == 0) return;      // warp-uniform
    if (nflag == 1 && rstd[r * K + (r - set * K)] < 0.f) return;      // only the diagonal
  }
  float4 w = *reinterpret_cast<const float4*>(u + r * H + 4 * lane);
  {
    const float4 bv = make_float4(b1[4 * lane], b1[4 * lane + 1], b1[4 * lane + 2], b1[4 * lane + 3]);
    const float m = warp_sum((w.x + w.y) + (w.z + w.w)) * (1.f / H);
    const float bm = warp_sum((bv.x + bv.y) + (bv.z + bv.w)) * (1.f / H);
    w = make_float4((w.x - m) + (bv.x - bm), (w.y - m) + (bv.y - bm), (w.z - m) + (bv.z - bm), (w.w - m) + (bv.w - bm));
  }
  float* row = rstd + r * K;
  for (int a0 = 0; a0 < K; a0 += 32) {
    float v = (a0 + lane < K) ? row[a0 + lane] : 1.f;
    unsigned mask = __ballot_sync(0xffffffffu, v < 0.f);
    if (!mask) continue;
    for (unsigned mk = mask; mk; mk &= mk - 1) {
      const int bit = __ffs(mk) - 1;
      const float4 ua = *reinterpret_cast<const float4*>(u + (set * K + a0 + bit) * H + 4 * lane);
      const float ma = warp_sum((ua.x + ua.y) + (ua.z + ua.w)) * (1.f / H);
      const float hx = w.x - (ua.x - ma), hy = w.y - (ua.y - ma), hz = w.z - (ua.z - ma), hw = w.w - (ua.w - ma);
      const float ss = warp_sum(fmaf(hx, hx, fmaf(hy, hy, fmaf(hz, hz, hw * hw))));
      if (lane == bit) v = rsqrtf(fmaf(ss, 1.f / H, eps));
    }
    if ((mask >> lane) & 1u) row[a0 + lane] = v;
  }
}

// ------------------------------------------------------------------------------------------
// number of valid pairs per set -> inv_count (per set, or shared when joint_mean)
// ------------------------------------------------------------------------------------------
// grid (ceil(K / 256), S, splits of the b range), block 256: thread = one a, loop over its slice of b with the set's
// depths in shared memory.  The last CTA to finish (ticket counter count[S]) turns the counts into 1 / count.
__global__ void __launch_bounds__(256) rank_count(const float* __restrict__ depth, int K, int S, int mode, float thr,
                                                  int joint, int* __restrict__ count /* S + 1, zero on entry */,
                                                  float* __restrict__ inv) {
  extern __shared__ float sdep[];     // K depths of this set
  __shared__ int red[32];
  __shared__ int s_last;
  const int set = blockIdx.y;
  const float* d = depth + (int64_t)set * K;
  for (int k = threadIdx.x; k < K; k += blockDim.x) sdep[k] = d[k];
  __syncthreads();
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  const int chunk = (K + gridDim.z - 1) / gridDim.z, b0 = blockIdx.z * chunk, b1 = min(K, b0 + chunk);
  int c = 0;
  if (a < K) {
    const float da = sdep[a];
    if (mode == 0) {
      for (int b = b0; b < b1; ++b) c += (fabsf(sdep[b] - da) > thr) ? 1 : 0;
    } else {
      for (int b = b0; b < b1; ++b) c += (fabsf(tanhf(sdep[b] - da)) > thr) ? 1 : 0;
    }
  }
  c = block_sum(c, red);
  if (threadIdx.x == 0) {
    if (c) atomicAdd(count + set, c);
    __threadfence();
    const int ticket = atomicAdd(count + S, 1);
    s_last = ticket == (int)(gridDim.x * gridDim.y * gridDim.z) - 1;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  long long total = 0;
  if (joint)
    for (int k = 0; k < S; ++k) total += __ldcg(count + k);
  for (int s = threadIdx.x; s < S; s += blockDim.x) {
    const long long n = joint ? total : (long long)__ldcg(count + s);
    inv[s] = n > 0 ? (float)(1.0 / (double)n) : 0.f;
  }
}

// ------------------------------------------------------------------------------------------
// cross-view L1: pair p couples keypoint k of set 2p+1 (a) with keypoint k of set 2p (b).
// One half warp per keypoint.  Adds its u-gradient straight into the dub/dua partial slot 0.
// ------------------------------------------------------------------------------------------
template <bool GRAD>
__global__ void __launch_bounds__(256) rank_l1(RankParams p, const float* __restrict__ w_l1, double* __restrict__ l1_sum,
                                              float* __restrict__ du_extra /* (S, K, H) zero-initialised */) {
  __shared__ float part[16][3 * H + 1];      // per half warp: [gamma | beta | w2 | b2] partial sums
  const int pair = blockIdx.y;
  const int lane = threadIdx.x & 31, l16 = lane & 15, half = lane >> 4;
  const int K = p.K;
  const int sb = 2 * pair, sa = 2 * pair + 1;
  HeadConst hc;
  hc.load(p.gamma, p.beta, p.w2, p.b2, l16);
  float dgam[HPL], dbet[HPL], dw2[HPL], db2 = 0.f, loss_local = 0.f;
#pragma unroll
  for (int i = 0; i < HPL; ++i) dgam[i] = dbet[i] = dw2[i] = 0.f;
  // a CTA walks keypoints blockIdx.x * 16 + [0, 16), + 16 gridDim.x, ...: few CTAs per pair keep the number of global
  // parameter-gradient atomics (385 per CTA, all CTAs on the same addresses) small
  for (int k0 = blockIdx.x * 16; k0 < K; k0 += 16 * gridDim.x) {      // CTA-uniform bounds: both halves of a warp stay together
  const int k = k0 + (threadIdx.x >> 5) * 2 + half;
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
    if (l16 == 0) loss_local += fabsf(diff);
    if (GRAD) {
      const float sg = (diff > 0.f) ? 1.f : ((diff < 0.f) ? -1.f : 0.f);
      const float dout = sg * (p.use_tanh ? (1.f - o.s * o.s) : 1.f) * w_l1[pair] / (float)K;
      db2 += (l16 == 0) ? dout : 0.f;
      const float coef = dout * o.rstd;
#pragma unroll
      for (int j = 0; j < HPL; ++j) {
        const int i = j >> 1;
        const bool hi = j & 1;
        const int h = hidx(l16, j);
        const float gp = hi ? o.gp[i].y : o.gp[i].x, gg = hi ? o.g[i].y : o.g[i].x, xh = hi ? o.xh[i].y : o.xh[i].x;
        const float w2h = (hi ? hc.w2c[i].y : hc.w2c[i].x) * kC, w2g = hi ? hc.w2g[i].y : hc.w2g[i].x;
        const float pg = w2h * gp;
        dw2[j] += dout * gg * (1.f / kC);
        dbet[j] += dout * pg;
        dgam[j] += dout * pg * xh;
        const float dh = coef * (w2g * gp - o.m1 - xh * o.m2);
        du_extra[((int64_t)sb * K + k) * H + h] = dh;
        du_extra[((int64_t)sa * K + k) * H + h] = -dh;
      }
    }
  }
  }
  loss_local = warp_sum(loss_local);
  if (lane == 0 && loss_local != 0.f) atomicAdd(l1_sum + pair, (double)loss_local);
  if (GRAD) {
    // parameter sums of the CTA without atomics: every half warp (16 of them, each covering all 128 hidden units) writes
    // its 3 x 128 partial sums to shared memory, then one thread per entry adds the 16 rows.  (Shared float atomics are
    // compare-and-swap loops; 16 half warps hitting the same 384 addresses made this tail most of the kernel.)
    const int hw = threadIdx.x >> 4;      // 0..15
#pragma unroll
    for (int i = 0; i < HPL; ++i) {
      const int h = hidx(l16, i);
      part[hw][h] = dgam[i];
      part[hw][H + h] = dbet[i];
      part[hw][2 * H + h] = dw2[i];
    }
    db2 = warp_sum(db2);
    if (lane == 0) part[hw][3 * H] = db2;      // one entry per warp (its even half-warp row); the odd rows hold 0
    else if (l16 == 0) part[hw][3 * H] = 0.f;
    __syncthreads();
    float* gp = p.gparam + p.gparam_off;
    for (int e = threadIdx.x; e < 3 * H + 1; e += blockDim.x) {
      float t = 0.f;
#pragma unroll
      for (int k = 0; k < 16; ++k) t += part[k][e];
      if (t != 0.f) atomicAdd(gp + H + e, t);
    }
  }
}

// ------------------------------------------------------------------------------------------
// du = accumulated pair gradient (rank_pairs) + L1 part, every row centred over h; emits the bf16 hi / lo panels of du
// and the L1 term's share of the b1 gradient (the pair term's share comes from rank_pairs itself).
// grid (ceil(K/32), S), block 256 (8 warps x 32 lanes: lane -> 4 h, warp -> rows)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
    rank_reduce_du(const float* __restrict__ du, const float* __restrict__ du_extra, int S, int K,
                   __nv_bfloat16* __restrict__ du2 /* (S K, H) bf16 */, float* __restrict__ gb1) {
  __shared__ float colsum[8][H];
  const int set = blockIdx.y, k0 = blockIdx.x * 32;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  float bsum[4] = {0.f, 0.f, 0.f, 0.f};
  for (int r = w; r < 32; r += 8) {
    const int k = k0 + r;
    if (k < K) {
      // rank_pairs accumulates alpha (q - m2 xh) per pair; the LayerNorm-backward term -alpha m1 = -alpha mean_h(q)
      // summed over pairs is the mean over h of the accumulated row (mean_h(xh) = 0), removed here once per row
      float4 acc = __ldg(reinterpret_cast<const float4*>(du + ((int64_t)set * K + k) * H + 4 * lane));
      const float mb = warp_sum((acc.x + acc.y) + (acc.z + acc.w)) * (1.f / H);
      acc.x -= mb; acc.y -= mb; acc.z -= mb; acc.w -= mb;
      if (du_extra) {
        const float4 v = *reinterpret_cast<const float4*>(du_extra + ((int64_t)set * K + k) * H + 4 * lane);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        if ((set & 1) == 0) { bsum[0] += v.x; bsum[1] += v.y; bsum[2] += v.z; bsum[3] += v.w; }   // b side of the L1 pair
      }
      // du in bf16, row-major: the d feats GEMM reads it K-major, the d W1 GEMM MN-major (tc_gemm.cuh), so no
      // transposed copy is written
      __nv_bfloat16* row = du2 + ((int64_t)set * K + k) * H + 4 * lane;
      *reinterpret_cast<uint2*>(row) = make_uint2(pack_bf16x2(acc.x, acc.y), pack_bf16x2(acc.z, acc.w));
    }
  }
  if (!du_extra) return;      // CTA-uniform
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
  float *u, *inv_count, *du, *du_extra, *nb, *na, *rstd, *mu;
  double *loss_sum, *l1_sum;
  int* count;
  int* rowflag;
  uint8_t* zero_a;
  size_t zero_a_bytes, zero_b_bytes;
  size_t total;
  int ldd, TA, TB, groups, bpw;
  int gs;        // sets per split-K group of the d W1 contraction
};

// b rows per warp of a rank_pairs CTA.  A CTA is the unit of scheduling, two fit on an SM, and a CTA costs
// (b rows it walks per warp + 1) row walks (the +1 stands for loading the a tile and flushing its gradient).  Large
// problems keep 16 rows per warp, which issues the fewest a-side reductions; few sets (strong scaling: 8 pairs per GPU)
// or a ragged K leave the last wave of such CTAs mostly empty, so every b_per_warp in [2, 32] is tried on a model of
// the launch -- CTAs handed to the 2 x SMs slots in launch order (a tile fastest, then b tile, then set), the ragged
// last b tile at its real length -- and a different tile has to beat 16 rows by 5 % (the model overstates small
// gains: at cfg2 it prefers 32 rows by 3 %, which measures the same, 2.412 against 2.413 ms).  cfg4 with 8 pairs per GPU: 8 rows per warp were
// 3 waves x 9 walks, 7 rows are 2.97 waves x 8 (283 -> 263 us).  Deterministic in (S, K): the grid depends on it.
double rank_launch_cost(int64_t S, int64_t K, int bpw, int64_t slots) {
  const int64_t ta = ceil_div<int64_t>(K, TILE_A), tile_b = (int64_t)WARPS * bpw, tb = ceil_div<int64_t>(K, tile_b);
  const int64_t last_rows = K - (tb - 1) * tile_b;
  const int cost_full = bpw + 1, cost_last = (int)ceil_div<int64_t>(last_rows, WARPS) + 1;
  const int64_t ctas = ta * tb * S;
  const double work = (double)S * ta * ((tb - 1) * cost_full + cost_last);
  if (ctas > 16 * slots) return work / slots;        // many waves: the tail does not matter
  // list scheduling in launch order with small integer costs: free_at[t] = slots that become free at time t
  // (O(CTAs + makespan); K changes from step to step in training, so a new (S, K) must cost microseconds, not ms)
  const int64_t horizon = (ceil_div<int64_t>(ctas, slots) + 2) * cost_full + 2;
  std::vector<int> free_at((size_t)horizon, 0);
  free_at[0] = (int)slots;
  int64_t issued = 0, x = 0, y = 0;
  int end = 0;
  for (int t = 0; issued < ctas && t < horizon - cost_full - 1; ++t)
    for (int nfree = free_at[t]; nfree > 0 && issued < ctas; --nfree) {
      const int done = t + ((y == tb - 1) ? cost_last : cost_full);
      ++free_at[done];
      end = std::max(end, done);
      ++issued;
      if (++x == ta) {
        x = 0;
        if (++y == tb) y = 0;
      }
    }
  return (double)end;
}
int rank_b_per_warp(int64_t S, int64_t K) {
  static std::mutex mu;
  static std::map<std::pair<int64_t, int64_t>, int> cache;
  std::lock_guard<std::mutex> lock(mu);
  auto it = cache.find({S, K});
  if (it != cache.end()) return it->second;
  const int64_t slots = CTAS_PER_SM * (int64_t)num_sms();
  int best = 16;
  double best_cost = 0.95 * rank_launch_cost(S, K, 16, slots);      // another tile has to win by 5 %
  for (int bpw = 2; bpw <= B_PER_WARP_MAX; ++bpw) {      // ties go to the smaller tile (the measured choice at cfg4 / 8 pairs)
    if (bpw == 16) continue;
    const double cost = rank_launch_cost(S, K, bpw, slots);
    if (cost < best_cost) {
      best_cost = cost;
      best = bpw;
    }
  }
  cache[{S, K}] = best;
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
  w.inv_count = c.take<float>(S);
  // zero-initialised accumulators, contiguous: one memset each for [loss_sum | l1_sum | count + ticket | rowflag] and
  // [du | du_extra]
  w.loss_sum = c.take<double>(S);
  w.l1_sum = c.take<double>(S);
  w.count = c.take<int>(S + 1);
  w.rowflag = c.take<int>(R);
  w.zero_a = reinterpret_cast<uint8_t*>(w.loss_sum);
  w.zero_a_bytes = (size_t)(reinterpret_cast<uint8_t*>(w.rowflag + R) - w.zero_a);
  if (backward) {
    w.du2 = c.take<__nv_bfloat16>(R * H);
    w.du = c.take<float>(R * H);
    w.zero_b_bytes = sizeof(float) * R * H;
    if (l1) {
      w.du_extra = c.take<float>(R * H);
      w.zero_b_bytes = (size_t)(reinterpret_cast<uint8_t*>(w.du_extra + R * H) - reinterpret_cast<uint8_t*>(w.du));
    }
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
  GD3_REQUIRE(thr >= 0.f, "gd3_depth_head_loss: thr must be >= 0 (a pair of equal depths is never valid)");
  GD3_REQUIRE(loss_rank, "gd3_depth_head_loss: null loss output");
  GD3_REQUIRE((grad_feats == nullptr) == (grad_params == nullptr),
              "gd3_depth_head_loss: pass both gradient buffers or neither");
  const bool l1 = w_l1 != nullptr;
  GD3_REQUIRE(!l1 || (S % 2 == 0 && loss_l1), "gd3_depth_head_loss: the L1 term needs an even number of sets and loss_l1");
  GD3_REQUIRE(S <= 65535, "gd3_depth_head_loss: at most 65535 sets per call");
  GD3_REQUIRE(K <= 12288, "gd3_depth_head_loss: at most 12288 keypoints per set (K^2 pairs are evaluated)");
  const bool backward = grad_feats != nullptr;
  if (K == 0) {      // every other case writes all the losses (rank_finalize)
    GD3_CHECK_CUDA(cudaMemsetAsync(loss_rank, 0, sizeof(float) * S, stream));
    if (l1) GD3_CHECK_CUDA(cudaMemsetAsync(loss_l1, 0, sizeof(float) * (S / 2), stream));
  }
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
  // every zero-initialised accumulator of the call in two memsets
  GD3_CHECK_CUDA(cudaMemsetAsync(w.zero_a, 0, w.zero_a_bytes, stream));
  if (backward) GD3_CHECK_CUDA(cudaMemsetAsync(w.du, 0, w.zero_b_bytes, stream));
  // ---- u = f W1^T on the tensor cores (split bf16, 3 K-concatenated panels) ----
  {
    // operands of u = f W1^T; the backward GEMMs read the same panels MN-major (no transposed copies)
    // centring group: the two sets of an image pair when the L1 term couples them, else one set (rank_group_mean)
    const int group_rows = (int)((l1 ? 2 : 1) * K);
    {
      dim3 grid((unsigned)(R / group_rows), (unsigned)ceil_div<int64_t>(D, 64));
      GD3_PROF("rank_group_mean", stream);
      if (D % 4 == 0 && reinterpret_cast<uintptr_t>(feats) % 16 == 0 && reinterpret_cast<uintptr_t>(w.mu) % 16 == 0)
        rank_group_mean_v4<<<grid, 256, 0, stream>>>(feats, group_rows, (int)D, w.mu);
      else
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
    EpiRstd::Params ep{w.rstd, (int)K, w.nb, w.na, ln_eps, w.rowflag};
    if (K % 4 == 0 && reinterpret_cast<uintptr_t>(w.rstd) % 16 == 0) {
      if ((rc = tc::make_tmap_store32(&ep.tm_out, w.rstd, K, K, S, K, K * K))) return rc;
      ep.use_tma = 1;
    }
    if (bn == 256) rc = tc::launch_gemm<256, 8, EpiRstd>("rank_rstd_gemm", ta, tb, s, ep, stream);
    else if (bn == 192) rc = tc::launch_gemm<192, 8, EpiRstd>("rank_rstd_gemm", ta, tb, s, ep, stream);
    else rc = tc::launch_gemm<128, 8, EpiRstd>("rank_rstd_gemm", ta, tb, s, ep, stream);
    if (rc) return rc;
    {
      GD3_PROF("rank_fix_rstd", stream);
      rank_fix_rstd<<<(unsigned)ceil_div<int64_t>(R, 8), 256, 0, stream>>>(w.u, b1, w.rowflag, R, (int)K, ln_eps, w.rstd);
    }
    GD3_CHECK_LAUNCH();
  }
  // ---- valid-pair counts ----
  {
    // ~2 CTAs per SM: the b range of a set is split when there are few sets
    const int64_t xy = ceil_div<int64_t>(K, 256) * S;
    int64_t split = ceil_div<int64_t>(2 * (int64_t)num_sms(), xy);
    split = std::max<int64_t>(1, std::min<int64_t>(split, ceil_div<int64_t>(K, 64)));
    dim3 grid((unsigned)ceil_div<int64_t>(K, 256), (unsigned)S, (unsigned)split);
    {
      GD3_PROF("rank_count", stream);
      rank_count<<<grid, 256, sizeof(float) * K, stream>>>(depths, (int)K, (int)S, mode, thr, joint_mean, w.count,
                                                          w.inv_count);
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
  rp.du = w.du;
  rp.gparam = grad_params;
  rp.gparam_off = (int64_t)H * D;
  // the cross-view L1 term first: it needs u only, and ahead of the pair kernel it overlaps the other branches of a step
  // instead of sitting in the tail behind rank_pairs
  if (l1) {
    // about one wave of CTAs (2 per SM): a CTA walks several blocks of 16 keypoints
    const int64_t gx_max = ceil_div<int64_t>(2 * (int64_t)num_sms(), S / 2);
    const int64_t gx = std::min<int64_t>(ceil_div<int64_t>(K, 16), std::max<int64_t>(gx_max, 1));
    dim3 grid((unsigned)gx, (unsigned)(S / 2));
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
    const size_t smem = sizeof(float) * (2 * TILE_A * H + TILE_A + 4 * H + 8) + 2048;      // + alignment slack of rank_pairs_w
    dim3 grid((unsigned)w.TA, (unsigned)w.TB, (unsigned)S);
#define GD3_RANK_LAUNCH(G, M, T)                                            \
  do {                                                                     \
    static SmemOptIn opt;                                                  \
    GD3_CHECK_CUDA(opt.ensure(rank_pairs<G, M, T>, smem));            \
    GD3_PROF("rank_pairs", stream);                                        \
    rank_pairs<G, M, T><<<grid, WARPS * 32, smem, stream>>>(rp);      \
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
      rank_reduce_du<<<grid, 256, 0, stream>>>(w.du, l1 ? w.du_extra : nullptr, (int)S, (int)K, w.du2,
                                             grad_params + (int64_t)H * D);
    }
    GD3_CHECK_LAUNCH();
  }
  {
    // d feats = du W1: A = du (K-major), B = W1 hi panel of W3 read MN-major ([k = h][mn = d])
    CUtensorMap t_du, t_w1;
    if ((rc = tc::make_tmap_bf16(&t_du, w.du2, H, R, 1, H, 0, tc::BM))) return rc;
    if ((rc = tc::make_tmap_bf16(&t_w1, w.W3, D, H, 1, 3 * (int64_t)w.ldd, 0, 64))) return rc;
    tc::EpiStoreF32::Params e1{grad_feats, (int)R, (int)D, D, 0, 1.0f, nullptr};
    if ((rc = tc::enable_tma_store(e1, 1))) return rc;
    tc::GemmShape s1{(int)R, (int)D, H, 1};
    if ((rc = tc::launch_gemm<256, 8, tc::EpiStoreF32, false, true>("rank_df_gemm", t_du, t_w1, s1, e1, stream))) return rc;
    // d W1 (H x D) = sum over sets of du_s^T f_s: a grouped contraction over (set of the group, 64-keypoint block), both
    // operands read MN-major from their row-major buffers (du2, the hi panel of F3); one GEMM batch entry per group, fp32
    // atomic accumulation.  Plain bf16 operands: the sum runs over all S K keypoints with incoherent terms, so the 2^-9
    // operand rounding averages out (the three-term split of round 1 changed |d W1| by 3e-5 and cost 44 instead of 25 us).
    CUtensorMap t_dum, t_fm;
    if ((rc = tc::make_tmap_bf16(&t_dum, w.du2, H, K, S, H, K * (int64_t)H, 64))) return rc;
    if ((rc = tc::make_tmap_bf16(&t_fm, w.F3, 3 * (int64_t)w.ldd, K, S, 3 * (int64_t)w.ldd, K * 3 * (int64_t)w.ldd, 64)))
      return rc;
    tc::GroupedK gk;
    gk.panels = 1;
    gk.sets_per_group = w.gs;
    gk.row_blocks = (int)ceil_div<int64_t>(K, tc::BK);
    EpiAtomicAddF32::Params e2{grad_params, H, (int)D, D};
    tc::GemmShape s2{H, (int)D, gk.panels * gk.sets_per_group * gk.row_blocks * tc::BK, w.groups};
    if ((rc = tc::launch_gemm<256, 8, EpiAtomicAddF32, true, true, true>("rank_dw1_gemm", t_dum, t_fm, s2, e2, stream, 0, gk)))
      return rc;
  }
  return GD3_OK;
}

}  // extern "C"
