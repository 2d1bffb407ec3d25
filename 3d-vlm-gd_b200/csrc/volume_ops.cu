// The two volume-level helpers of the reference's cost-volume loss, for callers that already hold (B, N, N2) volumes
// (the "minimal drop-in" of INTEGRATION.md section 2; the fused gd3_cost_kl never builds a volume):
//   get_masked_patch_cost  utils/functions.py:402-422   mask rows / columns, then row-normalise or softmax(x / T) in fp32
//   kl_divergence_map      utils/losses.py:5-15         mean over rows of sum_j t~ log(t~ / s~), t~ = clamp_min(t, eps)
// Both are HBM-bound streaming kernels with a row reduction: one warp per row, 128-bit loads, every element of a
// volume is read once per pass (forward: one pass for kl_divergence_map, two for the row-normalisation / softmax whose
// second pass hits L1 / L2 -- a row is a few KB).  Forward and backward of kl_divergence_map are ONE kernel (the
// gradients do not depend on the reduction); the masked-cost backward is a second kernel of the same shape.
#include "../../include/gd3.h"
#include "common.cuh"

namespace gd3 {
namespace {

constexpr int VO_WARPS = 8;      // rows per CTA

// torch.clamp_min semantics: NaN stays NaN (fmaxf would drop it)
__device__ __forceinline__ float clamp_min_t(float x, float lo) { return x < lo ? lo : x; }

// ------------------------------------------------------------------------------------------
// kl_divergence_map: row_kl[r] = sum_j t~ log(t~ / s~);  d/ds = -(t~ / s~) [s >= eps] / R;  d/dt = (log(t~ / s~) + 1) [t >= eps] / R
// ------------------------------------------------------------------------------------------
template <bool VEC>
__global__ void __launch_bounds__(VO_WARPS * 32)
    kl_map_kernel(const float* __restrict__ T, const float* __restrict__ S, int64_t rows, int n, float eps, float inv_rows,
                  float* __restrict__ row_kl, float* __restrict__ gS, float* __restrict__ gT) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * VO_WARPS + warp;
  if (r >= rows) return;
  const float* t = T + r * n;
  const float* s = S + r * n;
  float acc = 0.f;
  auto one = [&](float tv, float sv, float& gs, float& gt) {
    const float tc = clamp_min_t(tv, eps), sc = clamp_min_t(sv, eps);
    const float q = tc / sc;
    const float l = logf(q);
    acc = fmaf(tc, l, acc);
    gs = (sv >= eps) ? -q * inv_rows : 0.f;
    gt = (tv >= eps) ? (l + 1.f) * inv_rows : 0.f;
  };
  if (VEC) {
    for (int j = 4 * lane; j < n; j += 128) {
      const float4 tv = *reinterpret_cast<const float4*>(t + j);
      const float4 sv = *reinterpret_cast<const float4*>(s + j);
      float4 a, b;
      one(tv.x, sv.x, a.x, b.x);
      one(tv.y, sv.y, a.y, b.y);
      one(tv.z, sv.z, a.z, b.z);
      one(tv.w, sv.w, a.w, b.w);
      if (gS) *reinterpret_cast<float4*>(gS + r * n + j) = a;
      if (gT) *reinterpret_cast<float4*>(gT + r * n + j) = b;
    }
  } else {
    for (int j = lane; j < n; j += 32) {
      float a, b;
      one(t[j], s[j], a, b);
      if (gS) gS[r * n + j] = a;
      if (gT) gT[r * n + j] = b;
    }
  }
  acc = warp_sum(acc);
  if (lane == 0) row_kl[r] = acc;
}

// deterministic final sum (fixed order, double): loss = sum(row_kl) / R
__global__ void __launch_bounds__(1024) kl_map_reduce(const float* __restrict__ row_kl, int64_t rows, float* __restrict__ loss) {
  __shared__ double red[32];
  double acc = 0.0;
  for (int64_t i = threadIdx.x; i < rows; i += 1024) acc += (double)row_kl[i];
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) loss[0] = (float)(acc / (double)rows);
}

// ------------------------------------------------------------------------------------------
// get_masked_patch_cost.  xm_j = keep(i, j) ? x_j : 0;  mode 0: y = xm / max(sum xm, eps);  mode 1: y = softmax(xm / T)
// A masked row is all zeros: 0 after the row-normalisation, uniform 1 / N2 after the softmax -- as in the reference.
// ------------------------------------------------------------------------------------------
template <bool VEC>
__global__ void __launch_bounds__(VO_WARPS * 32)
    masked_cost_fwd(const float* __restrict__ X, int64_t rows, int hw, int n, const uint8_t* __restrict__ m1,
                    const uint8_t* __restrict__ m2, int use_softmax, float eps, float temp, float* __restrict__ Y,
                    float* __restrict__ row_sum) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * VO_WARPS + warp;
  if (r >= rows) return;
  const bool keep_row = m1[r % hw] != 0;
  const float* x = X + r * n;
  float* y = Y + r * n;
  auto masked = [&](int j, float v) { return (keep_row && (!m2 || m2[j])) ? v : 0.f; };
  auto logit = [&](int j, float v) { return temp != 1.f ? masked(j, v) / temp : masked(j, v); };      // IEEE division, like masked_cost / temperature
  // pass 1: row sum (mode 0) or row maximum (mode 1)
  float red = use_softmax ? -INFINITY : 0.f;
  if (VEC) {
    for (int j = 4 * lane; j < n; j += 128) {
      const float4 v = *reinterpret_cast<const float4*>(x + j);
      if (use_softmax)
        red = fmaxf(red, fmaxf(fmaxf(logit(j, v.x), logit(j + 1, v.y)), fmaxf(logit(j + 2, v.z), logit(j + 3, v.w))));
      else
        red += (masked(j, v.x) + masked(j + 1, v.y)) + (masked(j + 2, v.z) + masked(j + 3, v.w));
    }
  } else {
    for (int j = lane; j < n; j += 32) {
      if (use_softmax) red = fmaxf(red, logit(j, x[j]));
      else red += masked(j, x[j]);
    }
  }
  float scale, mx = 0.f;      // reciprocal of the row's divisor: sum of exponentials, or max(row sum, eps)
  if (use_softmax) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) red = fmaxf(red, __shfl_xor_sync(0xffffffffu, red, o));
    mx = red;
    float se = 0.f;
    if (VEC) {
      for (int j = 4 * lane; j < n; j += 128) {
        const float4 v = *reinterpret_cast<const float4*>(x + j);
        se += (expf(logit(j, v.x) - mx) + expf(logit(j + 1, v.y) - mx)) +
              (expf(logit(j + 2, v.z) - mx) + expf(logit(j + 3, v.w) - mx));
      }
    } else {
      for (int j = lane; j < n; j += 32) se += expf(logit(j, x[j]) - mx);
    }
    se = warp_sum(se);
    scale = 1.f / se;
    if (lane == 0 && row_sum) row_sum[r] = se;
  } else {
    red = warp_sum(red);
    scale = 1.f / clamp_min_t(red, eps);
    if (lane == 0 && row_sum) row_sum[r] = red;
  }
  // pass 2: write (the row is a few KB: L1 / L2 hits)
  if (VEC) {
    for (int j = 4 * lane; j < n; j += 128) {
      const float4 v = *reinterpret_cast<const float4*>(x + j);
      float4 o;
      if (use_softmax) {
        o.x = expf(logit(j, v.x) - mx) * scale;
        o.y = expf(logit(j + 1, v.y) - mx) * scale;
        o.z = expf(logit(j + 2, v.z) - mx) * scale;
        o.w = expf(logit(j + 3, v.w) - mx) * scale;
      } else {
        o.x = masked(j, v.x) * scale;
        o.y = masked(j + 1, v.y) * scale;
        o.z = masked(j + 2, v.z) * scale;
        o.w = masked(j + 3, v.w) * scale;
      }
      *reinterpret_cast<float4*>(y + j) = o;
    }
  } else {
    for (int j = lane; j < n; j += 32) {
      y[j] = use_softmax ? expf(logit(j, x[j]) - mx) * scale : masked(j, x[j]) * scale;
    }
  }
}

// The same for rows of at most 128 NV elements with 16-byte aligned rows: the row lives in registers, so the volume is
// read once and every exponential is evaluated once (the streaming kernel above reads a row three times and
// evaluates exp twice: 116 us against ~50 us per 134 MB volume).
template <int NV>
__global__ void __launch_bounds__(VO_WARPS * 32)
    masked_cost_fwd_reg(const float* __restrict__ X, int64_t rows, int hw, int n, const uint8_t* __restrict__ m1,
                        const uint8_t* __restrict__ m2, int use_softmax, float eps, float temp, float* __restrict__ Y,
                        float* __restrict__ row_sum) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * VO_WARPS + warp;
  if (r >= rows) return;
  const bool keep_row = m1[r % hw] != 0;
  const float* x = X + r * n;
  float* y = Y + r * n;
  const float pad = use_softmax ? -INFINITY : 0.f;      // beyond the row: exp(-inf) = 0, sum + 0
  auto val = [&](int j, float v) {
    const float m = (keep_row && (!m2 || m2[j])) ? v : 0.f;
    return (use_softmax && temp != 1.f) ? m / temp : m;   // IEEE division, like masked_cost / temperature (x / 1 = x)
  };
  float4 v[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int j = 4 * (lane + 32 * i);
    if (j < n) {
      const float4 q = *reinterpret_cast<const float4*>(x + j);
      v[i] = make_float4(val(j, q.x), val(j + 1, q.y), val(j + 2, q.z), val(j + 3, q.w));
    } else {
      v[i] = make_float4(pad, pad, pad, pad);
    }
  }
  float div;
  if (use_softmax) {
    float mx = -INFINITY;
#pragma unroll
    for (int i = 0; i < NV; ++i) mx = fmaxf(mx, fmaxf(fmaxf(v[i].x, v[i].y), fmaxf(v[i].z, v[i].w)));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float se = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      v[i] = make_float4(expf(v[i].x - mx), expf(v[i].y - mx), expf(v[i].z - mx), expf(v[i].w - mx));
      se += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
    div = warp_sum(se);
    if (lane == 0 && row_sum) row_sum[r] = div;
  } else {
    float sm = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) sm += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    sm = warp_sum(sm);
    if (lane == 0 && row_sum) row_sum[r] = sm;
    div = clamp_min_t(sm, eps);
  }
  // one reciprocal per row and a multiply per element (<= 1.5 ulp against the reference's division; the IEEE division per
  // element made the softmax launch compute-bound: 2 divisions + exp per element)
  const float inv = 1.f / div;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int j = 4 * (lane + 32 * i);
    if (j < n) *reinterpret_cast<float4*>(y + j) = make_float4(v[i].x * inv, v[i].y * inv, v[i].z * inv, v[i].w * inv);
  }
}

// backward.  dot = sum_j dy_j y_j.
//   softmax:        d xm_j = y_j (dy_j - dot) / T
//   row-normalise:  d xm_j = (dy_j - [sum >= eps] dot) / max(sum, eps)      (xm_j / r = y_j; clamp_min passes the gradient at sum >= eps)
// and d x_j = keep(i, j) ? d xm_j : 0 (masked entries were overwritten with 0).
template <bool VEC>
__global__ void __launch_bounds__(VO_WARPS * 32)
    masked_cost_bwd(const float* __restrict__ dY, const float* __restrict__ Y, const float* __restrict__ row_sum, int64_t rows,
                    int hw, int n, const uint8_t* __restrict__ m1, const uint8_t* __restrict__ m2, int use_softmax, float eps,
                    float temp, float* __restrict__ dX) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * VO_WARPS + warp;
  if (r >= rows) return;
  const bool keep_row = m1[r % hw] != 0;
  const float* dy = dY + r * n;
  const float* y = Y + r * n;
  float* dx = dX + r * n;
  if (!keep_row) {      // every entry of the row was overwritten: no gradient reaches the volume
    if (VEC) for (int j = 4 * lane; j < n; j += 128) *reinterpret_cast<float4*>(dx + j) = make_float4(0.f, 0.f, 0.f, 0.f);
    else for (int j = lane; j < n; j += 32) dx[j] = 0.f;
    return;
  }
  float dot = 0.f;
  if (VEC) {
    for (int j = 4 * lane; j < n; j += 128) {
      const float4 a = *reinterpret_cast<const float4*>(dy + j);
      const float4 b = *reinterpret_cast<const float4*>(y + j);
      dot += (a.x * b.x + a.y * b.y) + (a.z * b.z + a.w * b.w);
    }
  } else {
    for (int j = lane; j < n; j += 32) dot = fmaf(dy[j], y[j], dot);
  }
  dot = warp_sum(dot);
  const float s = row_sum[r];
  const float k_dot = use_softmax ? dot : ((s >= eps) ? dot : 0.f);
  const float k_out = use_softmax ? temp : clamp_min_t(s, eps);      // divided by, like autograd's grad / T and grad / row_sum
  auto one = [&](int j, float g, float yv) {
    const float v = use_softmax ? yv * (g - k_dot) / k_out : (g - k_dot) / k_out;
    return (!m2 || m2[j]) ? v : 0.f;
  };
  if (VEC) {
    for (int j = 4 * lane; j < n; j += 128) {
      const float4 a = *reinterpret_cast<const float4*>(dy + j);
      const float4 b = *reinterpret_cast<const float4*>(y + j);
      *reinterpret_cast<float4*>(dx + j) = make_float4(one(j, a.x, b.x), one(j + 1, a.y, b.y), one(j + 2, a.z, b.z), one(j + 3, a.w, b.w));
    }
  } else {
    for (int j = lane; j < n; j += 32) dx[j] = one(j, dy[j], y[j]);
  }
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace
}  // namespace gd3

using namespace gd3;

extern "C" {

size_t gd3_kl_divergence_map_workspace(int64_t rows) {
  Carver c(nullptr);
  c.take<float>((size_t)(rows > 0 ? rows : 1));
  return c.total();
}

int gd3_kl_divergence_map(const float* teacher, const float* student, int64_t rows, int64_t n, float eps, float* loss,
                          float* grad_student, float* grad_teacher, void* workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  GD3_REQUIRE(rows > 0 && n > 0 && n < (1ll << 31), "gd3_kl_divergence_map: bad sizes (%lld x %lld)", (long long)rows, (long long)n);
  GD3_REQUIRE(teacher && student && loss && workspace, "gd3_kl_divergence_map: null pointer");
  GD3_REQUIRE(workspace_bytes >= gd3_kl_divergence_map_workspace(rows), "gd3_kl_divergence_map: workspace too small");
  Carver c(workspace);
  float* row_kl = c.take<float>((size_t)rows);
  const bool vec = n % 4 == 0 && aligned16(teacher) && aligned16(student) && (!grad_student || aligned16(grad_student)) &&
                   (!grad_teacher || aligned16(grad_teacher));
  const unsigned grid = (unsigned)ceil_div<int64_t>(rows, VO_WARPS);
  {
    GD3_PROF("kl_map", stream);
    if (vec)
      kl_map_kernel<true><<<grid, VO_WARPS * 32, 0, stream>>>(teacher, student, rows, (int)n, eps, 1.f / (float)rows, row_kl,
                                                              grad_student, grad_teacher);
    else
      kl_map_kernel<false><<<grid, VO_WARPS * 32, 0, stream>>>(teacher, student, rows, (int)n, eps, 1.f / (float)rows, row_kl,
                                                               grad_student, grad_teacher);
  }
  GD3_CHECK_LAUNCH();
  {
    GD3_PROF("kl_map_reduce", stream);
    kl_map_reduce<<<1, 1024, 0, stream>>>(row_kl, rows, loss);
  }
  GD3_CHECK_LAUNCH();
  return GD3_OK;
}

int gd3_masked_patch_cost(const float* cost, int64_t B, int64_t hw, int64_t hw2, const uint8_t* mask1, const uint8_t* mask2,
                          int use_softmax, float eps, float temperature, float* out, float* row_sum, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  GD3_REQUIRE(B >= 0 && hw > 0 && hw2 > 0 && hw2 < (1ll << 31), "gd3_masked_patch_cost: bad sizes");
  if (B == 0) return GD3_OK;
  GD3_REQUIRE(cost && mask1 && out, "gd3_masked_patch_cost: null pointer");
  GD3_REQUIRE(!use_softmax || temperature != 0.f, "gd3_masked_patch_cost: zero temperature");
  const int64_t rows = B * hw;
  const bool vec = hw2 % 4 == 0 && aligned16(cost) && aligned16(out);
  const unsigned grid = (unsigned)ceil_div<int64_t>(rows, VO_WARPS);
  GD3_PROF("masked_cost_fwd", stream);
  if (vec && hw2 <= 1024)
    masked_cost_fwd_reg<8><<<grid, VO_WARPS * 32, 0, stream>>>(cost, rows, (int)hw, (int)hw2, mask1, mask2, use_softmax, eps,
                                                               temperature, out, row_sum);
  else if (vec && hw2 <= 2048)
    masked_cost_fwd_reg<16><<<grid, VO_WARPS * 32, 0, stream>>>(cost, rows, (int)hw, (int)hw2, mask1, mask2, use_softmax, eps,
                                                                temperature, out, row_sum);
  else if (vec)
    masked_cost_fwd<true><<<grid, VO_WARPS * 32, 0, stream>>>(cost, rows, (int)hw, (int)hw2, mask1, mask2, use_softmax, eps,
                                                              temperature, out, row_sum);
  else
    masked_cost_fwd<false><<<grid, VO_WARPS * 32, 0, stream>>>(cost, rows, (int)hw, (int)hw2, mask1, mask2, use_softmax, eps,
                                                               temperature, out, row_sum);
  GD3_CHECK_LAUNCH();
  return GD3_OK;
}

int gd3_masked_patch_cost_backward(const float* grad_out, const float* out, const float* row_sum, int64_t B, int64_t hw,
                                   int64_t hw2, const uint8_t* mask1, const uint8_t* mask2, int use_softmax, float eps,
                                   float temperature, float* grad_cost, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  GD3_REQUIRE(B >= 0 && hw > 0 && hw2 > 0 && hw2 < (1ll << 31), "gd3_masked_patch_cost_backward: bad sizes");
  if (B == 0) return GD3_OK;
  GD3_REQUIRE(grad_out && out && row_sum && mask1 && grad_cost, "gd3_masked_patch_cost_backward: null pointer");
  const int64_t rows = B * hw;
  const bool vec = hw2 % 4 == 0 && aligned16(grad_out) && aligned16(out) && aligned16(grad_cost);
  const unsigned grid = (unsigned)ceil_div<int64_t>(rows, VO_WARPS);
  GD3_PROF("masked_cost_bwd", stream);
  if (vec)
    masked_cost_bwd<true><<<grid, VO_WARPS * 32, 0, stream>>>(grad_out, out, row_sum, rows, (int)hw, (int)hw2, mask1, mask2,
                                                              use_softmax, eps, temperature, grad_cost);
  else
    masked_cost_bwd<false><<<grid, VO_WARPS * 32, 0, stream>>>(grad_out, out, row_sum, rows, (int)hw, (int)hw2, mask1, mask2,
                                                               use_softmax, eps, temperature, grad_cost);
  GD3_CHECK_LAUNCH();
  return GD3_OK;
}

}  // extern "C"
