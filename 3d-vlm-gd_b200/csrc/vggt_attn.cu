// VGGT teacher cost volumes, fused per global-attention block (VGGT half).
//
// Replaces, for one block, the return_attn branch of Attention.custom_scaled_dot_product_attention
// (vggt/layers/attention.py:73-84): scores = (q * scale)[view 1 patches] . k[view 2 patches]^T per head (and the other
// direction), softmax(scores / temperature) over the keys -- and folds in what the callers do with these maps afterwards:
// the mean over the collected blocks (vggt/models/aggregator.py:259-260,273) and the mean over heads
// (src/finetune_timm_vggt.py:390-392).  The per-block (2 B, heads, n, n) probability tensors, their stack over 24
// blocks (2.6 GB at n = 925) and the two means are never built; one block costs one score GEMM per direction and one
// pass over the scores.
//
//   1 tc_gemm<Store> x2   S12[b, h] = q1 k2^T,  S21[b, h] = q2 k1^T     (bf16 in, fp32 out, K = head_dim)
//   2 va_softmax_heads    one warp per (row, pair, direction): loop over the heads, row softmax in registers, head sum,
//                         attn[b, row, :] (+)= weight / heads * sum
//
// The reference runs the teacher under bf16 autocast (src/finetune_timm_vggt.py:359): the score matmul returns bf16 and
// "scores / temperature" is a bf16 op again, the softmax then runs in fp32.  round_bf16 = 1 reproduces those two
// roundings on the fp32 accumulators.
#include "../../include/gd3.h"
#include "common.cuh"
#include "tc_gemm.cuh"

#include <cmath>

namespace gd3 {
namespace {

struct VAWorkspace {
  float* S;          // (2, B * heads, n, lds) scores of both directions
  int lds;
  size_t total;
};

VAWorkspace carve_va(void* base, int64_t B, int64_t heads, int64_t n) {
  Carver c(base);
  VAWorkspace w;
  w.lds = (int)round_up<int64_t>(n, 4);
  w.S = c.take<float>((size_t)(2 * B * heads * n * w.lds));
  w.total = c.total();
  return w;
}

__device__ __forceinline__ float round_to_bf16(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

// grid (ceil(n / 8), B, 2), block 256 (8 warps = 8 rows).  NIT = ceil(n / 32) register columns per lane.
// DIV: the temperature is no power of two, so "scores / temperature" stays a true division like the reference's
// (bf16 / float -> bf16); otherwise tparam = 1 / temperature and the multiplication is exact.
template <int NIT, bool DIV>
__global__ void __launch_bounds__(256) va_softmax_heads(const float* __restrict__ S, int lds, int B, int heads, int n,
                                                        float tparam, int round_bf16, float w_out, int accumulate,
                                                        float* __restrict__ attn12, float* __restrict__ attn21) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int row = blockIdx.x * 8 + warp, b = blockIdx.y, dir = blockIdx.z;
  if (row >= n) return;
  float acc[NIT];
#pragma unroll
  for (int t = 0; t < NIT; ++t) acc[t] = 0.f;
  const float* base = S + (((int64_t)dir * B + b) * heads * n + row) * lds;
  for (int h = 0; h < heads; ++h) {
    const float* srow = base + (int64_t)h * n * lds;
    float x[NIT];
    float m = -INFINITY;
#pragma unroll
    for (int t = 0; t < NIT; ++t) {
      const int j = lane + 32 * t;
      float v = (j < n) ? __ldg(srow + j) : -INFINITY;
      if (round_bf16) v = round_to_bf16(v);
      v = DIV ? __fdiv_rn(v, tparam) : v * tparam;
      if (round_bf16) v = round_to_bf16(v);
      x[t] = v;
      m = fmaxf(m, v);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off));
    float sum = 0.f;
#pragma unroll
    for (int t = 0; t < NIT; ++t) {
      x[t] = exp2f((x[t] - m) * 1.4426950408889634f);     // -inf padding -> 0
      sum += x[t];
    }
    sum = warp_sum(sum);
    const float inv = 1.f / sum;
#pragma unroll
    for (int t = 0; t < NIT; ++t) acc[t] = fmaf(x[t], inv, acc[t]);
  }
  float* out = (dir == 0 ? attn12 : attn21) + ((int64_t)b * n + row) * n;
#pragma unroll
  for (int t = 0; t < NIT; ++t) {
    const int j = lane + 32 * t;
    if (j < n) out[j] = accumulate ? fmaf(acc[t], w_out, out[j]) : acc[t] * w_out;
  }
}

bool is_pow2(float t) {
  int e;
  return std::frexp(t, &e) == 0.5f;
}

template <int NIT>
void launch_softmax(const VAWorkspace& w, int64_t B, int64_t heads, int64_t n, float temperature, int round_bf16,
                    float w_out, int accumulate, float* attn12, float* attn21, cudaStream_t stream) {
  dim3 grid((unsigned)ceil_div<int64_t>(n, 8), (unsigned)B, 2);
  GD3_PROF("va_softmax_heads", stream);
  if (is_pow2(temperature))
    va_softmax_heads<NIT, false><<<grid, 256, 0, stream>>>(w.S, w.lds, (int)B, (int)heads, (int)n, 1.f / temperature,
                                                           round_bf16, w_out, accumulate, attn12, attn21);
  else
    va_softmax_heads<NIT, true><<<grid, 256, 0, stream>>>(w.S, w.lds, (int)B, (int)heads, (int)n, temperature, round_bf16,
                                                          w_out, accumulate, attn12, attn21);
}

}  // namespace
}  // namespace gd3

using namespace gd3;

extern "C" {

size_t gd3_vggt_attn_workspace(int64_t B, int64_t heads, int64_t n) {
  if (B <= 0 || heads <= 0 || n <= 0) return 256;
  return carve_va(nullptr, B, heads, n).total;
}

int gd3_vggt_attn_accumulate(const void* q_scaled, const void* k, int64_t B, int64_t heads, int64_t n_tokens,
                             int64_t head_dim, int64_t skip, float temperature, int round_bf16, float weight,
                             int accumulate, float* attn12, float* attn21, void* workspace, size_t workspace_bytes,
                             void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  GD3_REQUIRE(B >= 0 && heads > 0 && n_tokens > 0 && head_dim > 0 && skip >= 0, "gd3_vggt_attn_accumulate: bad sizes");
  GD3_REQUIRE(n_tokens % 2 == 0, "gd3_vggt_attn_accumulate: the token axis must hold two views of equal length");
  GD3_REQUIRE(head_dim % 8 == 0, "gd3_vggt_attn_accumulate: head_dim must be a multiple of 8 (16-byte rows)");
  GD3_REQUIRE(temperature > 0.f, "gd3_vggt_attn_accumulate: temperature must be positive");
  const int64_t half = n_tokens / 2, n = half - skip;
  GD3_REQUIRE(n > 0, "gd3_vggt_attn_accumulate: no patch tokens left after skipping %lld", (long long)skip);
  GD3_REQUIRE(n <= 2048, "gd3_vggt_attn_accumulate: at most 2048 patch tokens per view (rows are kept in registers)");
  GD3_REQUIRE(B * heads <= 65535 && B <= 65535, "gd3_vggt_attn_accumulate: batch too large");
  if (B == 0) return GD3_OK;
  GD3_REQUIRE(q_scaled && k && attn12 && attn21, "gd3_vggt_attn_accumulate: null input or output");
  VAWorkspace w = carve_va(workspace, B, heads, n);
  if (!workspace || workspace_bytes < w.total) {
    set_error("gd3_vggt_attn_accumulate: workspace too small (%zu < %zu)", workspace_bytes, w.total);
    return GD3_ERR_WORKSPACE;
  }
  const __nv_bfloat16* Q = static_cast<const __nv_bfloat16*>(q_scaled);
  const __nv_bfloat16* Kk = static_cast<const __nv_bfloat16*>(k);
  const int64_t bstride = n_tokens * head_dim;
  int rc;
  for (int dir = 0; dir < 2; ++dir) {
    // direction 0: queries of view 1 (rows skip .. half) against keys of view 2 (rows half + skip ..); 1: the reverse
    const __nv_bfloat16* qa = Q + (dir == 0 ? skip : half + skip) * head_dim;
    const __nv_bfloat16* kb = Kk + (dir == 0 ? half + skip : skip) * head_dim;
    CUtensorMap ta, tb;
    if ((rc = tc::make_tmap_bf16(&ta, qa, head_dim, n, B * heads, head_dim, bstride, tc::BM))) return rc;
    if ((rc = tc::make_tmap_bf16(&tb, kb, head_dim, n, B * heads, head_dim, bstride, 128))) return rc;
    float* Sd = w.S + (int64_t)dir * B * heads * n * w.lds;
    tc::EpiStoreF32::Params ep{Sd, (int)n, (int)n, w.lds, n * (int64_t)w.lds, 1.0f, nullptr};
    tc::GemmShape s{(int)n, (int)n, (int)head_dim, (int)(B * heads)};
    if ((rc = tc::launch_gemm<128, 8, tc::EpiStoreF32>("va_score_gemm", ta, tb, s, ep, stream))) return rc;
  }
  const float w_out = weight / (float)heads;
  if (n <= 1024)
    launch_softmax<32>(w, B, heads, n, temperature, round_bf16, w_out, accumulate, attn12, attn21, stream);
  else
    launch_softmax<64>(w, B, heads, n, temperature, round_bf16, w_out, accumulate, attn12, attn21, stream);
  GD3_CHECK_LAUNCH();
  return GD3_OK;
}

}  // extern "C"
