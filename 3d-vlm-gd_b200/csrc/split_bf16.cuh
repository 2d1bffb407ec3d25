// Operand preparation shared by the Smooth-AP and depth-head pipelines: split an fp32 matrix into bf16
// hi / lo parts and lay them out for the tcgen05 GEMMs.
//
//   X3  (R x 3 ldd)  row-major panels [hi | hi | lo] (lo_panel = 2, A side) or [hi | lo | hi] (lo_panel = 1,
//                    B side): one K-concatenated GEMM then yields hi*hi + hi*lo + lo*hi (~16 mantissa bits)
//   XT               the transposed copy (channels x rows), either hi only or as three panels, addressed
//                    through XtLayout so the same kernel serves per-pair (Smooth-AP) and grouped (d W1) layouts
//
// Fast path: D % 8 == 0, K % 8 == 0, 16-byte aligned rows: 64 x 64 tiles, 128-bit global accesses, padded
// shared-memory transpose.  Everything else goes through the scalar kernel.
#pragma once

#include "common.cuh"

namespace gd3 {

// column of (row -> set = row / K, k = row % K) in the transposed buffer, per precision panel
struct XtLayout {
  int panels;          // 0: no transposed output, 1: hi only, 3: [hi | lo | hi]
  int K;               // rows per set
  int sets_per_group;  // gs
  int ldk;             // padded rows per set
  int64_t group_len;   // gl = gs * ldk (stride between panels)
  int64_t ld;          // elements between consecutive channels (rows of XT)
  int64_t set_stride;  // panels == 1: extra offset per set (channels * ld of one pair); 0 for the grouped layout
  __host__ __device__ int64_t off(int64_t c, int set, int k, int panel) const {
    if (panels == 1) return (int64_t)set * set_stride + c * ld + k;
    return c * ld + ((int64_t)(set / sets_per_group) * 3 + panel) * group_len + (int64_t)(set % sets_per_group) * ldk + k;
  }
};

namespace split_detail {

__device__ __forceinline__ void hi_lo(float v, uint16_t& hi, uint16_t& lo) {
  const __nv_bfloat16 h = __float2bfloat16(v);
  const __nv_bfloat16 l = __float2bfloat16(v - __bfloat162float(h));
  hi = *reinterpret_cast<const uint16_t*>(&h);
  lo = *reinterpret_cast<const uint16_t*>(&l);
}

// generic scalar kernel: 32 x 32 tiles
static __global__ void __launch_bounds__(256)
    split_scalar(const float* __restrict__ x, int64_t R, int D, int ldd, int lo_panel, __nv_bfloat16* __restrict__ X3,
                 __nv_bfloat16* __restrict__ XT, XtLayout xl, const float* __restrict__ mu, int mu_rows) {
  __shared__ float tile[32][33];
  const int64_t r0 = (int64_t)blockIdx.x * 32;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int c0 = 0; c0 < ldd; c0 += 32) {
    __syncthreads();
    for (int r = w; r < 32; r += 8) {
      const int64_t row = r0 + r;
      const int c = c0 + lane;
      float v = (row < R && c < D) ? __ldg(x + row * D + c) : 0.f;
      if (mu && row < R && c < D) v -= __ldg(mu + (row / mu_rows) * D + c);
      tile[r][lane] = v;
      if (row < R && c < ldd) {
        uint16_t hi, lo;
        hi_lo(v, hi, lo);
        uint16_t* o = reinterpret_cast<uint16_t*>(X3) + row * 3 * ldd + c;
        o[0] = hi;
        o[(3 - lo_panel) * ldd] = hi;
        o[lo_panel * ldd] = lo;
      }
    }
    __syncthreads();
    if (XT && xl.panels)
      for (int r = w; r < 32; r += 8) {
        const int c = c0 + r;
        const int64_t row = r0 + lane;
        if (c < D && row < R) {
          uint16_t hi, lo;
          hi_lo(tile[lane][r], hi, lo);
          const int set = (int)(row / xl.K), k = (int)(row % xl.K);
          uint16_t* o = reinterpret_cast<uint16_t*>(XT);
          o[xl.off(c, set, k, 0)] = hi;
          if (xl.panels == 3) {
            o[xl.off(c, set, k, 1)] = lo;
            o[xl.off(c, set, k, 2)] = hi;
          }
        }
      }
  }
}

// fast kernel: 64 rows x 64 channels per iteration, 256 threads, thread = (row tr, 16 channels at tc).
// With a transposed output a CTA's 64 rows belong to ONE set (tiles_per_set CTAs per set), so that an aligned group
// of 8 rows never straddles two sets whatever K is; the rows past the end of the set are zeros (padding up to ldk).
static __global__ void __launch_bounds__(256)
    split_fast(const float* __restrict__ x, int64_t R, int D, int ldd, int lo_panel, __nv_bfloat16* __restrict__ X3,
               __nv_bfloat16* __restrict__ XT, XtLayout xl, int tiles_per_set, const float* __restrict__ mu,
               int mu_rows) {
  constexpr int TS = 66;
  __shared__ uint16_t t_hi[64 * TS];
  __shared__ uint16_t t_lo[64 * TS];
  const bool want_t = XT != nullptr && xl.panels != 0;
  const int set = want_t ? blockIdx.x / tiles_per_set : 0;
  const int k0 = want_t ? (blockIdx.x - set * tiles_per_set) * 64 : 0;
  const int64_t r0 = want_t ? (int64_t)set * xl.K + k0 : (int64_t)blockIdx.x * 64;
  const int64_t r_end = want_t ? (int64_t)set * xl.K + xl.K : R;       // rows of this CTA stop at the end of its set
  const int tr = threadIdx.x >> 2, tc = (threadIdx.x & 3) * 16;
  for (int c0 = 0; c0 < D; c0 += 64) {
    {
      const int64_t row = r0 + tr;
      const bool row_ok = row < r_end && row < R;
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const int c = c0 + tc + 8 * hh;
        float v[8];
        if (row_ok && c < D) {
          const float4 a = __ldg(reinterpret_cast<const float4*>(x + row * D + c));
          const float4 b = __ldg(reinterpret_cast<const float4*>(x + row * D + c) + 1);
          v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
          if (mu) {
            const float* m = mu + (row / mu_rows) * D + c;
            const float4 ma = __ldg(reinterpret_cast<const float4*>(m)), mb = __ldg(reinterpret_cast<const float4*>(m) + 1);
            v[0] -= ma.x; v[1] -= ma.y; v[2] -= ma.z; v[3] -= ma.w; v[4] -= mb.x; v[5] -= mb.y; v[6] -= mb.z; v[7] -= mb.w;
          }
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) v[i] = 0.f;
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
        if (row_ok && c < D) {
          __nv_bfloat16* o = X3 + row * 3 * ldd + c;
          const uint4 vh = make_uint4(ph[0], ph[1], ph[2], ph[3]), vl = make_uint4(pl[0], pl[1], pl[2], pl[3]);
          *reinterpret_cast<uint4*>(o) = vh;
          *reinterpret_cast<uint4*>(o + (3 - lo_panel) * ldd) = vh;
          *reinterpret_cast<uint4*>(o + lo_panel * ldd) = vl;
        }
        if (want_t) {
          uint32_t* th = reinterpret_cast<uint32_t*>(t_hi + tr * TS + tc + 8 * hh);
          uint32_t* tl = reinterpret_cast<uint32_t*>(t_lo + tr * TS + tc + 8 * hh);
#pragma unroll
          for (int i = 0; i < 4; ++i) { th[i] = ph[i]; tl[i] = pl[i]; }
        }
      }
    }
    if (!want_t) continue;
    __syncthreads();
    {
      const int c = c0 + tr;            // this thread writes channel c, rows r0 + tc + [0, 16)
      if (c < D) {
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          const int nl = tc + 8 * hh;
          if (k0 + nl < xl.K) {         // groups of 8 are aligned inside the set; rows past K in the group are zeros
            uint32_t ph[4], pl[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              ph[i] = (uint32_t)t_hi[(nl + 2 * i) * TS + tr] | ((uint32_t)t_hi[(nl + 2 * i + 1) * TS + tr] << 16);
              pl[i] = (uint32_t)t_lo[(nl + 2 * i) * TS + tr] | ((uint32_t)t_lo[(nl + 2 * i + 1) * TS + tr] << 16);
            }
            const int k = k0 + nl;
            const uint4 vh = make_uint4(ph[0], ph[1], ph[2], ph[3]), vl = make_uint4(pl[0], pl[1], pl[2], pl[3]);
            *reinterpret_cast<uint4*>(XT + xl.off(c, set, k, 0)) = vh;
            if (xl.panels == 3) {
              *reinterpret_cast<uint4*>(XT + xl.off(c, set, k, 1)) = vl;
              *reinterpret_cast<uint4*>(XT + xl.off(c, set, k, 2)) = vh;
            }
          }
        }
      }
    }
    __syncthreads();
  }
}

}  // namespace split_detail

// rows beyond R inside the last aligned group of 8 are written as zeros by the fast path only when they exist
// in the tile; callers that rely on zero padding clear the buffers themselves.
// mu (optional): (R / mu_rows, D) fp32 row-group means that are subtracted before the split (x - mu[row / mu_rows]),
// so that the bf16 panels resolve the deviations and not a large component shared by the whole group.
inline int launch_split3(const char* name, const float* x, int64_t R, int D, int ldd, int lo_panel,
                         __nv_bfloat16* X3, __nv_bfloat16* XT, const XtLayout& xl, cudaStream_t stream,
                         const float* mu = nullptr, int mu_rows = 1) {
  if (R <= 0) return GD3_OK;
  const bool want_t = XT != nullptr && xl.panels != 0;
  const bool t_ok = !want_t || (xl.ld % 8 == 0 && xl.ldk % 8 == 0 && xl.ldk >= round_up(xl.K, 8) && xl.group_len % 8 == 0 &&
                                xl.set_stride % 8 == 0 && R % xl.K == 0 && reinterpret_cast<uintptr_t>(XT) % 16 == 0);
  const bool fast = D % 8 == 0 && ldd == D && reinterpret_cast<uintptr_t>(x) % 16 == 0 &&
                    reinterpret_cast<uintptr_t>(X3) % 16 == 0 && reinterpret_cast<uintptr_t>(mu) % 16 == 0 && t_ok;
  {
    GD3_PROF(name, stream);
    if (fast) {
      const int tps = want_t ? ceil_div(xl.K, 64) : 1;
      const int64_t blocks = want_t ? (R / xl.K) * tps : ceil_div<int64_t>(R, 64);
      split_detail::split_fast<<<(unsigned)blocks, 256, 0, stream>>>(x, R, D, ldd, lo_panel, X3, XT, xl, tps, mu, mu_rows);
    }
    else
      split_detail::split_scalar<<<(unsigned)ceil_div<int64_t>(R, 32), 256, 0, stream>>>(x, R, D, ldd, lo_panel, X3, XT,
                                                                                         xl, mu, mu_rows);
  }
  GD3_CHECK_LAUNCH();
  return GD3_OK;
}

}  // namespace gd3
