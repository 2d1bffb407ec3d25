// Operand preparation shared by the Smooth-AP / InfoNCE and depth-head pipelines: split an fp32 matrix into bf16
// hi / lo parts laid out for the tcgen05 GEMMs.
//
//   X3  (R x 3 ldd)  row-major panels [hi | hi | lo] (lo_panel = 2, A side) or [hi | lo | hi] (lo_panel = 1, B side):
//                    one K-concatenated GEMM then yields hi*hi + hi*lo + lo*hi (~16 mantissa bits)
//
// The backward GEMMs read the same row-major panels through MN-major operand descriptors (tc_gemm.cuh), so no
// transposed copy is produced here.  An optional per-row-group mean is subtracted before the split, so that the bf16
// panels resolve the deviations and not a large component shared by the whole group.
#pragma once

#include "common.cuh"

namespace gd3 {

namespace split_detail {

__device__ __forceinline__ void hi_lo(float v, uint16_t& hi, uint16_t& lo) {
  const __nv_bfloat16 h = __float2bfloat16(v);
  const __nv_bfloat16 l = __float2bfloat16(v - __bfloat162float(h));
  hi = *reinterpret_cast<const uint16_t*>(&h);
  lo = *reinterpret_cast<const uint16_t*>(&l);
}

// generic kernel: one element per thread, columns [D, ldd) of every panel are zero
static __global__ void __launch_bounds__(256)
    split_scalar(const float* __restrict__ x, int64_t R, int D, int ldd, int lo_panel, __nv_bfloat16* __restrict__ X3,
                 const float* __restrict__ mu, int mu_rows) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= R * ldd) return;
  const int64_t row = e / ldd;
  const int c = (int)(e - row * ldd);
  float v = 0.f;
  if (c < D) {
    v = __ldg(x + row * D + c);
    if (mu) v -= __ldg(mu + (row / mu_rows) * D + c);
  }
  uint16_t hi, lo;
  hi_lo(v, hi, lo);
  uint16_t* o = reinterpret_cast<uint16_t*>(X3) + row * 3 * ldd + c;
  o[0] = hi;
  o[(3 - lo_panel) * ldd] = hi;
  o[lo_panel * ldd] = lo;
}

// fast kernel (D % 8 == 0, 16-byte aligned rows): one thread per 8 consecutive channels, 128-bit loads and stores
static __global__ void __launch_bounds__(256)
    split_fast(const float* __restrict__ x, int64_t R, int D, int lo_panel, __nv_bfloat16* __restrict__ X3,
               const float* __restrict__ mu, int mu_rows) {
  const int d8 = D >> 3;
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= R * d8) return;
  const int64_t row = e / d8;
  const int c = (int)(e - row * d8) * 8;
  const float4 a = __ldg(reinterpret_cast<const float4*>(x + row * D + c));
  const float4 b = __ldg(reinterpret_cast<const float4*>(x + row * D + c) + 1);
  float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
  if (mu) {
    const float* m = mu + (row / mu_rows) * D + c;
    const float4 ma = __ldg(reinterpret_cast<const float4*>(m)), mb = __ldg(reinterpret_cast<const float4*>(m) + 1);
    v[0] -= ma.x; v[1] -= ma.y; v[2] -= ma.z; v[3] -= ma.w; v[4] -= mb.x; v[5] -= mb.y; v[6] -= mb.z; v[7] -= mb.w;
  }
  uint16_t h[8], l[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) hi_lo(v[i], h[i], l[i]);
  uint32_t ph[4], pl[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    ph[i] = (uint32_t)h[2 * i] | ((uint32_t)h[2 * i + 1] << 16);
    pl[i] = (uint32_t)l[2 * i] | ((uint32_t)l[2 * i + 1] << 16);
  }
  const uint4 vh = make_uint4(ph[0], ph[1], ph[2], ph[3]), vl = make_uint4(pl[0], pl[1], pl[2], pl[3]);
  __nv_bfloat16* o = X3 + row * 3 * D + c;
  *reinterpret_cast<uint4*>(o) = vh;
  *reinterpret_cast<uint4*>(o + (3 - lo_panel) * D) = vh;
  *reinterpret_cast<uint4*>(o + lo_panel * D) = vl;
}

}  // namespace split_detail

// mu (optional): (R / mu_rows, D) fp32 row-group means, subtracted before the split (x - mu[row / mu_rows])
inline int launch_split3(const char* name, const float* x, int64_t R, int D, int ldd, int lo_panel, __nv_bfloat16* X3,
                         cudaStream_t stream, const float* mu = nullptr, int mu_rows = 1) {
  if (R <= 0) return GD3_OK;
  const bool fast = D % 8 == 0 && ldd == D && reinterpret_cast<uintptr_t>(x) % 16 == 0 &&
                    reinterpret_cast<uintptr_t>(X3) % 16 == 0 && reinterpret_cast<uintptr_t>(mu) % 16 == 0;
  {
    GD3_PROF(name, stream);
    if (fast)
      split_detail::split_fast<<<(unsigned)ceil_div<int64_t>(R * (D / 8), 256), 256, 0, stream>>>(x, R, D, lo_panel, X3, mu,
                                                                                              mu_rows);
    else
      split_detail::split_scalar<<<(unsigned)ceil_div<int64_t>(R * ldd, 256), 256, 0, stream>>>(x, R, D, ldd, lo_panel, X3,
                                                                                            mu, mu_rows);
  }
  GD3_CHECK_LAUNCH();
  return GD3_OK;
}

}  // namespace gd3
