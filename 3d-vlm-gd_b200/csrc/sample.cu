// K3: bilinear sampling of patch-token feature maps at pixel keypoints (+ optional channel L2
// normalisation and mean over several layers), forward and backward.
//
// Replaces interpolate_features (utils/functions.py:55-76) and the glue around it in
// get_intermediate_feature / get_feature (src/finetune_timm_mast3r.py:271-277, 307-313):
//   pixel -> grid coords (a*pts + b, patch centres on [-1, 1]) -> grid_sample(align_corners=True,
//   padding_mode='border') -> optional F.normalize over channels.
// The arithmetic follows PyTorch's grid_sampler (unnormalise ((g+1)/2)*(size-1), clip to
// [0, size-1], 4 taps with (1-frac) weights) with the same fp32 operation order; FMA contraction is
// suppressed on the coordinate path so tap selection is identical.
// Generic strides let the same kernel read the reference's NCHW maps or token-major ViT outputs.
#include "../../include/gd3.h"
#include "common.cuh"

namespace gd3 {
namespace {

struct SampleGeom {
  float aw, ah, bw, bh;   // pixel -> normalised grid
  int ph, pw;             // feature map height / width in patches
};

struct Taps {
  int i00, i01, i10, i11;     // token indices (y*pw + x) of nw, ne, sw, se
  float w00, w01, w10, w11;
};

__device__ __forceinline__ Taps make_taps(const SampleGeom& g, float x, float y) {
  // keypoints = a * pts + b   (two roundings, like the reference's tensor ops)
  const float gx = __fadd_rn(__fmul_rn(g.aw, x), g.bw);
  const float gy = __fadd_rn(__fmul_rn(g.ah, y), g.bh);
  // grid_sampler_unnormalize, align_corners=True, then clip (padding_mode='border')
  float ix = __fmul_rn(__fdiv_rn(__fadd_rn(gx, 1.f), 2.f), (float)(g.pw - 1));
  float iy = __fmul_rn(__fdiv_rn(__fadd_rn(gy, 1.f), 2.f), (float)(g.ph - 1));
  ix = fminf((float)(g.pw - 1), fmaxf(ix, 0.f));
  iy = fminf((float)(g.ph - 1), fmaxf(iy, 0.f));
  const float fx = floorf(ix), fy = floorf(iy);
  const int x0 = (int)fx, y0 = (int)fy, x1 = x0 + 1, y1 = y0 + 1;
  const float wx1 = __fsub_rn(ix, fx), wy1 = __fsub_rn(iy, fy);
  const float wx0 = __fsub_rn((float)x1, ix), wy0 = __fsub_rn((float)y1, iy);
  Taps t;
  const bool xin = x1 < g.pw, yin = y1 < g.ph;   // out-of-range neighbours are skipped (weight is 0 anyway)
  t.w00 = __fmul_rn(wx0, wy0);
  t.w01 = xin ? __fmul_rn(wx1, wy0) : 0.f;
  t.w10 = yin ? __fmul_rn(wx0, wy1) : 0.f;
  t.w11 = (xin && yin) ? __fmul_rn(wx1, wy1) : 0.f;
  const int xc = xin ? x1 : x0, yc = yin ? y1 : y0;
  t.i00 = y0 * g.pw + x0;
  t.i01 = y0 * g.pw + xc;
  t.i10 = yc * g.pw + x0;
  t.i11 = yc * g.pw + xc;
  return t;
}

template <class T>
__device__ __forceinline__ float ldf(const T* p);
template <>
__device__ __forceinline__ float ldf<float>(const float* p) { return __ldg(p); }
template <>
__device__ __forceinline__ float ldf<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }

// one block per (pair, keypoint); threads stride over channels
template <class T>
__global__ void __launch_bounds__(256)
    sample_fwd_kernel(const T* __restrict__ tok, int L, int64_t sL, int64_t sP, int64_t sN, int64_t sC,
                      const float* __restrict__ kp, int K, int C, SampleGeom g, int normalize,
                      float* __restrict__ out, int64_t oP, int64_t oK, int64_t oC, float* __restrict__ inv_norm) {
  __shared__ float red[32];
  const int k = blockIdx.x, p = blockIdx.y;
  const float x = kp[((int64_t)p * K + k) * 2 + 0], y = kp[((int64_t)p * K + k) * 2 + 1];
  const Taps t = make_taps(g, x, y);
  const float invL = 1.f / (float)L;
  float ss = 0.f;
  float* o = out + p * oP + k * oK;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float acc = 0.f;
    for (int l = 0; l < L; ++l) {
      const T* base = tok + l * sL + p * sP + c * sC;
      // same accumulation order as grid_sample: nw, ne, sw, se
      float v = ldf(base + t.i00 * sN) * t.w00;
      v += ldf(base + t.i01 * sN) * t.w01;
      v += ldf(base + t.i10 * sN) * t.w10;
      v += ldf(base + t.i11 * sN) * t.w11;
      acc += v;
    }
    acc = (L > 1) ? acc * invL : acc;
    o[c * oC] = acc;
    ss += acc * acc;
  }
  if (normalize) {
    ss = block_sum(ss, red);
    const float inv = 1.f / fmaxf(sqrtf(ss), 1e-12f);   // F.normalize eps
    if (threadIdx.x == 0 && inv_norm) inv_norm[(int64_t)p * K + k] = inv;
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) o[c * oC] *= inv;   // each thread rescales what it wrote
  }
}

// grad_tokens (fp32, same strides as tokens, caller-zeroed) += scatter of grad_out through the taps
__global__ void __launch_bounds__(256)
    sample_bwd_kernel(const float* __restrict__ gout, int64_t gP, int64_t gK, int64_t gC,
                      const float* __restrict__ out, int64_t oP, int64_t oK, int64_t oC,
                      const float* __restrict__ inv_norm, const float* __restrict__ kp, int K, int C, SampleGeom g,
                      int normalize, int L, float* __restrict__ gtok, int64_t sL, int64_t sP, int64_t sN,
                      int64_t sC, const float* __restrict__ gextra, int64_t eP, int64_t eK, int64_t eC) {
  __shared__ float red[32];
  const int k = blockIdx.x, p = blockIdx.y;
  const float x = kp[((int64_t)p * K + k) * 2 + 0], y = kp[((int64_t)p * K + k) * 2 + 1];
  const Taps t = make_taps(g, x, y);
  const float* go = gout + p * gP + k * gK;
  float dot = 0.f, inv = 1.f;
  if (normalize) {
    const float* o = out + p * oP + k * oK;
    for (int c = threadIdx.x; c < C; c += blockDim.x) dot += o[c * oC] * go[c * gC];
    dot = block_sum(dot, red);
    inv = inv_norm[(int64_t)p * K + k];
  }
  const float invL = 1.f / (float)L;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float gv = go[c * gC];
    if (normalize) gv = (gv - out[p * oP + k * oK + c * oC] * dot) * inv;   // d/dx of x / max(|x|, eps)
    if (gextra) gv += gextra[p * eP + k * eK + c * eC];   // gradient w.r.t. the un-normalised sample
    gv *= invL;
    for (int l = 0; l < L; ++l) {
      float* base = gtok + l * sL + p * sP + c * sC;
      atomicAdd(base + t.i00 * sN, gv * t.w00);
      if (t.w01 != 0.f) atomicAdd(base + t.i01 * sN, gv * t.w01);
      if (t.w10 != 0.f) atomicAdd(base + t.i10 * sN, gv * t.w10);
      if (t.w11 != 0.f) atomicAdd(base + t.i11 * sN, gv * t.w11);
    }
  }
}

// ------------------------------------------------------------------------------------------
// Fast path: channel-contiguous tokens (sC == 1, C % 4 == 0, 16-byte aligned rows) and channel-contiguous
// output.  One warp per keypoint, 8 keypoints per CTA; each lane owns float4 channel groups, keeps the sampled
// row in registers for the normalisation (C <= 4096) and issues 128-bit loads / stores only.
// ------------------------------------------------------------------------------------------
constexpr int FAST_MAX_IT = 32;    // float4 groups per lane: C <= 32 lanes * 4 * 32 = 4096

// raw 4-channel load (8 bytes of bf16 or 16 bytes of fp32) and its conversion, so that many loads can be in flight
// before the first conversion
template <class T> struct Raw4;
template <> struct Raw4<float> {
  using type = float4;
  static __device__ __forceinline__ type ld(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
  static __device__ __forceinline__ float4 cvt(const type& v) { return v; }
};
template <> struct Raw4<__nv_bfloat16> {
  using type = uint2;
  static __device__ __forceinline__ type ld(const __nv_bfloat16* p) { return __ldg(reinterpret_cast<const uint2*>(p)); }
  static __device__ __forceinline__ float4 cvt(const type& v) {
    return make_float4(bf16_bits_to_float(v.x & 0xFFFFu), bf16_bits_to_float(v.x >> 16),
                       bf16_bits_to_float(v.y & 0xFFFFu), bf16_bits_to_float(v.y >> 16));
  }
};

// The 4 taps of GRP channel groups are requested together before the first conversion (the loop over the layers has a
// runtime trip count, so the compiler cannot hoist them itself): 150 -> 107 us per step at cfg2.
template <class T, int NIT>
__global__ void __launch_bounds__(256)
    sample_fwd_fast(const T* __restrict__ tok, int L, int64_t sL, int64_t sP, int64_t sN, const float* __restrict__ kp,
                    int K, int C, SampleGeom g, int normalize, float* __restrict__ out, int64_t oP, int64_t oK,
                    float* __restrict__ inv_norm) {
  using R = Raw4<T>;
  constexpr int GRP = (sizeof(typename R::type) == 8) ? 4 : 2;      // iterations whose 4 taps are loaded together
  const int lane = threadIdx.x & 31;
  const int k = blockIdx.x * 8 + (threadIdx.x >> 5), p = blockIdx.y;
  if (k >= K) return;
  const float x = kp[((int64_t)p * K + k) * 2 + 0], y = kp[((int64_t)p * K + k) * 2 + 1];
  const Taps t = make_taps(g, x, y);
  const float invL = 1.f / (float)L;
  float4 acc[NIT];
#pragma unroll
  for (int it = 0; it < NIT; ++it) acc[it] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int l = 0; l < L; ++l) {
    const T* lbase = tok + l * sL + p * sP;
#pragma unroll
    for (int g0 = 0; g0 < NIT; g0 += GRP) {
      typename R::type raw[GRP][4];
#pragma unroll
      for (int q = 0; q < GRP; ++q) {
        const int c = ((g0 + q) * 32 + lane) * 4;
        if (g0 + q < NIT && c < C) {
          const T* base = lbase + c;
          raw[q][0] = R::ld(base + t.i00 * sN);
          raw[q][1] = R::ld(base + t.i01 * sN);
          raw[q][2] = R::ld(base + t.i10 * sN);
          raw[q][3] = R::ld(base + t.i11 * sN);
        }
      }
#pragma unroll
      for (int q = 0; q < GRP; ++q) {
        const int c = ((g0 + q) * 32 + lane) * 4;
        if (g0 + q < NIT && c < C) {
          const float4 v00 = R::cvt(raw[q][0]), v01 = R::cvt(raw[q][1]), v10 = R::cvt(raw[q][2]), v11 = R::cvt(raw[q][3]);
          // same accumulation order as grid_sample: nw, ne, sw, se
          float4 v;
          v.x = v00.x * t.w00; v.x += v01.x * t.w01; v.x += v10.x * t.w10; v.x += v11.x * t.w11;
          v.y = v00.y * t.w00; v.y += v01.y * t.w01; v.y += v10.y * t.w10; v.y += v11.y * t.w11;
          v.z = v00.z * t.w00; v.z += v01.z * t.w01; v.z += v10.z * t.w10; v.z += v11.z * t.w11;
          v.w = v00.w * t.w00; v.w += v01.w * t.w01; v.w += v10.w * t.w10; v.w += v11.w * t.w11;
          float4& a = acc[g0 + q];
          a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
        }
      }
    }
  }
  float ss = 0.f;
#pragma unroll
  for (int it = 0; it < NIT; ++it) {
    float4& a = acc[it];
    if (L > 1) { a.x *= invL; a.y *= invL; a.z *= invL; a.w *= invL; }
    ss += a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w;
  }
  float inv = 1.f;
  if (normalize) {
    ss = warp_sum(ss);
    inv = 1.f / fmaxf(sqrtf(ss), 1e-12f);
    if (lane == 0) inv_norm[(int64_t)p * K + k] = inv;
  }
  float* o = out + p * oP + k * oK;
#pragma unroll
  for (int it = 0; it < NIT; ++it) {
    const int c = (it * 32 + lane) * 4;
    if (c < C)
      *reinterpret_cast<float4*>(o + c) = make_float4(acc[it].x * inv, acc[it].y * inv, acc[it].z * inv, acc[it].w * inv);
  }
}

// backward fast path: gradient rows and token-gradient rows are channel-contiguous fp32
__global__ void __launch_bounds__(256)
    sample_bwd_fast(const float* __restrict__ gout, int64_t gP, int64_t gK, const float* __restrict__ out, int64_t oP,
                    int64_t oK, const float* __restrict__ inv_norm, const float* __restrict__ kp, int K, int C,
                    SampleGeom g, int normalize, int L, float* __restrict__ gtok, int64_t sL, int64_t sP, int64_t sN,
                    const float* __restrict__ gextra, int64_t eP, int64_t eK) {
  const int lane = threadIdx.x & 31;
  const int k = blockIdx.x * 8 + (threadIdx.x >> 5), p = blockIdx.y;
  if (k >= K) return;
  const float x = kp[((int64_t)p * K + k) * 2 + 0], y = kp[((int64_t)p * K + k) * 2 + 1];
  const Taps t = make_taps(g, x, y);
  const float* go = gout + p * gP + k * gK;
  const float* o = out + p * oP + k * oK;
  float dot = 0.f, inv = 1.f;
  if (normalize) {
    for (int c = lane * 4; c < C; c += 128) {
      const float4 a = *reinterpret_cast<const float4*>(o + c), b = *reinterpret_cast<const float4*>(go + c);
      dot += a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w;
    }
    dot = warp_sum(dot);
    inv = inv_norm[(int64_t)p * K + k];
  }
  const float invL = 1.f / (float)L;
  const float* ge = gextra ? gextra + p * eP + k * eK : nullptr;
  for (int c = lane * 4; c < C; c += 128) {
    float4 gv = *reinterpret_cast<const float4*>(go + c);
    if (normalize) {
      const float4 a = *reinterpret_cast<const float4*>(o + c);
      gv.x = (gv.x - a.x * dot) * inv; gv.y = (gv.y - a.y * dot) * inv;
      gv.z = (gv.z - a.z * dot) * inv; gv.w = (gv.w - a.w * dot) * inv;
    }
    if (ge) {   // gradient w.r.t. the un-normalised sample of the same keypoint (e.g. the depth features)
      const float4 e = *reinterpret_cast<const float4*>(ge + c);
      gv.x += e.x; gv.y += e.y; gv.z += e.z; gv.w += e.w;
    }
    gv.x *= invL; gv.y *= invL; gv.z *= invL; gv.w *= invL;
    for (int l = 0; l < L; ++l) {
      float* base = gtok + l * sL + p * sP + c;
      atomicAdd(reinterpret_cast<float4*>(base + t.i00 * sN), make_float4(gv.x * t.w00, gv.y * t.w00, gv.z * t.w00, gv.w * t.w00));
      if (t.w01 != 0.f)
        atomicAdd(reinterpret_cast<float4*>(base + t.i01 * sN), make_float4(gv.x * t.w01, gv.y * t.w01, gv.z * t.w01, gv.w * t.w01));
      if (t.w10 != 0.f)
        atomicAdd(reinterpret_cast<float4*>(base + t.i10 * sN), make_float4(gv.x * t.w10, gv.y * t.w10, gv.z * t.w10, gv.w * t.w10));
      if (t.w11 != 0.f)
        atomicAdd(reinterpret_cast<float4*>(base + t.i11 * sN), make_float4(gv.x * t.w11, gv.y * t.w11, gv.z * t.w11, gv.w * t.w11));
    }
  }
}

template <class T>
bool launch_fwd_fast(const T* tok, int L, int64_t sL, int64_t sP, int64_t sN, const float* kp, int K, int P, int C,
                     const SampleGeom& g, int normalize, float* out, int64_t oP, int64_t oK, float* inv_norm,
                     cudaStream_t stream) {
  dim3 grid((unsigned)ceil_div(K, 8), (unsigned)P);
  const int nit = ceil_div(C, 128);
#define GD3_FWD_FAST(N)                                                                                           \
  sample_fwd_fast<T, N><<<grid, 256, 0, stream>>>(tok, L, sL, sP, sN, kp, K, C, g, normalize, out, oP, oK, inv_norm)
  GD3_PROF("sample_fwd_fast", stream);
  if (nit <= 2) GD3_FWD_FAST(2);
  else if (nit <= 4) GD3_FWD_FAST(4);
  else if (nit <= 6) GD3_FWD_FAST(6);
  else if (nit <= 8) GD3_FWD_FAST(8);
  else if (nit <= 16) GD3_FWD_FAST(16);
  else if (nit <= FAST_MAX_IT) GD3_FWD_FAST(32);
  else return false;
#undef GD3_FWD_FAST
  return true;
}

int make_geom(int64_t ph, int64_t pw, int64_t h, int64_t w, int patch, int stride, SampleGeom* g) {
  GD3_REQUIRE(ph >= 2 && pw >= 2 && patch >= 1 && stride >= 1, "sample_tokens: feature map must be at least 2 x 2 (got %lld x %lld)",
              (long long)ph, (long long)pw);
  GD3_REQUIRE(h - patch >= stride && w - patch >= stride, "sample_tokens: image %lld x %lld too small for patch %d / stride %d",
              (long long)h, (long long)w, patch, stride);
  // utils/functions.py:56-63, same operation order in double, then rounded to fp32 like torch.tensor(...).float()
  const double half = patch / 2.0;
  const double last_h = (double)(((h - patch) / stride) * stride) + half;
  const double last_w = (double)(((w - patch) / stride) * stride) + half;
  const double ah = 2.0 / (last_h - half), aw = 2.0 / (last_w - half);
  const double bh = 1.0 - last_h * 2.0 / (last_h - half), bw = 1.0 - last_w * 2.0 / (last_w - half);
  g->aw = (float)aw;
  g->ah = (float)ah;
  g->bw = (float)bw;
  g->bh = (float)bh;
  g->ph = (int)ph;
  g->pw = (int)pw;
  return GD3_OK;
}

}  // namespace
}  // namespace gd3

using namespace gd3;

extern "C" {

int gd3_sample_tokens_fwd(const void* tokens, int dtype, int64_t L, int64_t P, int64_t C, int64_t ph, int64_t pw,
                          int64_t h, int64_t w, int64_t sL, int64_t sP, int64_t sN, int64_t sC, const float* kp,
                          int64_t K, int patch, int stride, int normalize, float* out, int64_t oP, int64_t oK, int64_t oC,
                          float* inv_norm, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (P == 0 || K == 0 || C == 0) return GD3_OK;
  GD3_REQUIRE(tokens && kp && out, "gd3_sample_tokens_fwd: null pointer");
  GD3_REQUIRE(L >= 1 && P > 0 && K > 0 && C > 0, "gd3_sample_tokens_fwd: bad sizes");
  GD3_REQUIRE(dtype == GD3_DTYPE_F32 || dtype == GD3_DTYPE_BF16, "gd3_sample_tokens_fwd: bad dtype %d", dtype);
  GD3_REQUIRE(!normalize || inv_norm, "gd3_sample_tokens_fwd: normalize needs inv_norm");
  GD3_REQUIRE(P <= 65535, "gd3_sample_tokens_fwd: more than 65535 pairs per call");
  SampleGeom g;
  int rc = make_geom(ph, pw, h, w, patch, stride, &g);
  if (rc) return rc;
  {
    // fast path: channel-contiguous source and destination, rows 16-byte aligned
    const int esz = dtype == GD3_DTYPE_F32 ? 4 : 2;
    const bool aligned = sC == 1 && oC == 1 && C % 4 == 0 && C <= 128 * FAST_MAX_IT &&
                         (reinterpret_cast<uintptr_t>(tokens) % 16 == 0) && (sN * esz) % (4 * esz) == 0 &&
                         (sN % 4 == 0) && (sP % 4 == 0) && (sL % 4 == 0) &&
                         (reinterpret_cast<uintptr_t>(out) % 16 == 0) && oK % 4 == 0 && oP % 4 == 0;
    if (aligned) {
      bool ok = dtype == GD3_DTYPE_F32
                    ? launch_fwd_fast(static_cast<const float*>(tokens), (int)L, sL, sP, sN, kp, (int)K, (int)P, (int)C, g,
                                      normalize, out, oP, oK, inv_norm, stream)
                    : launch_fwd_fast(static_cast<const __nv_bfloat16*>(tokens), (int)L, sL, sP, sN, kp, (int)K, (int)P,
                                      (int)C, g, normalize, out, oP, oK, inv_norm, stream);
      if (ok) {
        GD3_CHECK_LAUNCH();
        return GD3_OK;
      }
    }
  }
  dim3 grid((unsigned)K, (unsigned)P);
  const int threads = C >= 256 ? 256 : (C >= 128 ? 128 : 64);
  if (dtype == GD3_DTYPE_F32)
    {
      GD3_PROF("sample_fwd_kernel", stream);
      sample_fwd_kernel<float><<<grid, threads, 0, stream>>>(static_cast<const float*>(tokens), (int)L, sL, sP, sN, sC,
                                                          kp, (int)K, (int)C, g, normalize, out, oP, oK, oC,
                                                          inv_norm);
    }
  else
    {
      GD3_PROF("sample_fwd_kernel", stream);
      sample_fwd_kernel<__nv_bfloat16><<<grid, threads, 0, stream>>>(static_cast<const __nv_bfloat16*>(tokens), (int)L,
                                                                  sL, sP, sN, sC, kp, (int)K, (int)C, g, normalize,
                                                                  out, oP, oK, oC, inv_norm);
    }
  GD3_CHECK_LAUNCH();
  return GD3_OK;
}

int gd3_sample_tokens_bwd(const float* grad_out, int64_t gP, int64_t gK, int64_t gC, const float* out, int64_t oP,
                          int64_t oK, int64_t oC, const float* inv_norm, const float* kp, int64_t L, int64_t P,
                          int64_t K, int64_t C, int64_t ph, int64_t pw, int64_t h, int64_t w, int patch, int stride,
                          int normalize, float* grad_tokens, int64_t sL, int64_t sP, int64_t sN, int64_t sC,
                          const float* grad_extra, int64_t eP, int64_t eK, int64_t eC, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (P == 0 || K == 0 || C == 0) return GD3_OK;
  GD3_REQUIRE(grad_out && kp && grad_tokens, "gd3_sample_tokens_bwd: null pointer");
  GD3_REQUIRE(!normalize || (out && inv_norm), "gd3_sample_tokens_bwd: normalize needs out and inv_norm");
  GD3_REQUIRE(P <= 65535, "gd3_sample_tokens_bwd: more than 65535 pairs per call");
  SampleGeom g;
  int rc = make_geom(ph, pw, h, w, patch, stride, &g);
  if (rc) return rc;
  {
    const bool aligned = sC == 1 && gC == 1 && (!normalize || oC == 1) && C % 4 == 0 && sN % 4 == 0 && sP % 4 == 0 &&
                         sL % 4 == 0 && gK % 4 == 0 && gP % 4 == 0 && (!normalize || (oK % 4 == 0 && oP % 4 == 0)) &&
                         reinterpret_cast<uintptr_t>(grad_tokens) % 16 == 0 &&
                         reinterpret_cast<uintptr_t>(grad_out) % 16 == 0 &&
                         (!normalize || reinterpret_cast<uintptr_t>(out) % 16 == 0) &&
                         (!grad_extra || (eC == 1 && eK % 4 == 0 && eP % 4 == 0 &&
                                          reinterpret_cast<uintptr_t>(grad_extra) % 16 == 0));
    if (aligned) {
      dim3 fgrid((unsigned)ceil_div<int64_t>(K, 8), (unsigned)P);
      {
        GD3_PROF("sample_bwd_fast", stream);
        sample_bwd_fast<<<fgrid, 256, 0, stream>>>(grad_out, gP, gK, out, oP, oK, inv_norm, kp, (int)K, (int)C, g,
                                                  normalize, (int)L, grad_tokens, sL, sP, sN, grad_extra, eP, eK);
      }
      GD3_CHECK_LAUNCH();
      return GD3_OK;
    }
  }
  dim3 grid((unsigned)K, (unsigned)P);
  const int threads = C >= 256 ? 256 : (C >= 128 ? 128 : 64);
  {
    GD3_PROF("sample_bwd_kernel", stream);
    sample_bwd_kernel<<<grid, threads, 0, stream>>>(grad_out, gP, gK, gC, out, oP, oK, oC, inv_norm, kp, (int)K, (int)C,
                                                  g, normalize, (int)L, grad_tokens, sL, sP, sN, sC, grad_extra, eP, eK, eC);
  }
  GD3_CHECK_LAUNCH();
  return GD3_OK;
}

}  // extern "C"
