// Keypoint -> best-matching pixel of the other image ("keypoint-vs-all-pixels similarity"), SURVEY 8f-3.
//
// Replaces the inline block of the semantic-transfer evaluation (src/evaluate_timm.py:532-547): the reference
// bilinearly upsamples the (1, C, ph, pw) patch descriptors of image 2 to ((S - p) // s * s + 1)^2
// (align_corners=True), edge-pads them to the image size S x S (1.26 GB of fp32 at C = 768, S = 640), takes the dot
// product of every pixel with each of the K normalised keypoint descriptors of image 1 and arg-maxes over the pixels.
//
// Bilinear upsampling and edge padding are linear in the descriptors, so
//     sim[k, pixel] = upsample(pad)( S_low[k, :] )[pixel],   S_low[k, patch] = <kp_desc[k], desc2[:, patch]>,
// i.e. one small K x (ph pw) x C contraction followed by a per-keypoint 4-tap interpolation + arg-max over the
// S^2 pixels, which never materialises the upsampled descriptor map (100x fewer FLOPs, ~10^4 x less memory).
// The interpolation weights follow PyTorch's upsample_bilinear2d (align_corners=True): src = dst * (in - 1) /
// (out - 1), i0 = floor(src), lambda = src - i0, value = h0 (w0 v00 + w1 v01) + h1 (w0 v10 + w1 v11).
// arg-max ties resolve to the lowest pixel index (torch.argmax).
#include "../../include/gd3.h"
#include "common.cuh"

namespace gd3 {
namespace {

constexpr int KB = 16;         // keypoints per CTA of the low-resolution contraction

__device__ __forceinline__ unsigned long long ea_key(float score, uint32_t idx) {
  score = score + 0.0f;        // -0 -> +0
  uint32_t u = __float_as_uint(score);
  u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
  return (static_cast<unsigned long long>(u) << 32) | static_cast<unsigned long long>(0xFFFFFFFFu - idx);
}

// S_low[k][p] = sum_c kd[k][c] * desc2[c][p]        grid (ceil(P / 128), ceil(K / KB)), block 128
// kd: element (k, c) at kd[k * kd_sk + c * kd_sc]  (the reference hands over (1, C, K): kd_sk = 1, kd_sc = K)
__global__ void __launch_bounds__(128)
    ea_lowres_sim(const float* __restrict__ kd, int64_t kd_sk, int64_t kd_sc, const float* __restrict__ desc2, int K, int C,
                  int P, float* __restrict__ S) {
  extern __shared__ __align__(16) float s_kd[];     // [C][KB]
  const int k0 = blockIdx.y * KB;
  for (int e = threadIdx.x; e < C * KB; e += blockDim.x) {
    const int c = e / KB, j = e - c * KB;
    s_kd[e] = (k0 + j < K) ? kd[(int64_t)(k0 + j) * kd_sk + (int64_t)c * kd_sc] : 0.f;
  }
  __syncthreads();
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  float acc[KB];
#pragma unroll
  for (int j = 0; j < KB; ++j) acc[j] = 0.f;
  for (int c = 0; c < C; ++c) {
    const float v = __ldg(desc2 + (int64_t)c * P + p);
    const float4* q = reinterpret_cast<const float4*>(s_kd + c * KB);
#pragma unroll
    for (int j = 0; j < KB; j += 4) {
      const float4 q4 = q[j >> 2];
      acc[j] = fmaf(q4.x, v, acc[j]);
      acc[j + 1] = fmaf(q4.y, v, acc[j + 1]);
      acc[j + 2] = fmaf(q4.z, v, acc[j + 2]);
      acc[j + 3] = fmaf(q4.w, v, acc[j + 3]);
    }
  }
#pragma unroll
  for (int j = 0; j < KB; ++j)
    if (k0 + j < K) S[(int64_t)(k0 + j) * P + p] = acc[j];
}

// grid (chunks, K), block 256; dynamic smem: S_low[k] (ph * pw) | per-coordinate tables i0[img] (int), lambda[img]
__global__ void __launch_bounds__(256)
    ea_upsample_argmax(const float* __restrict__ S, int ph, int pw, int img, int ds, int pad, unsigned long long* keys) {
  extern __shared__ __align__(16) float ea_smem[];
  float* s_low = ea_smem;
  float* lam_y = s_low + ph * pw;
  float* lam_x = lam_y + img;
  int* i0_y = reinterpret_cast<int*>(lam_x + img);
  int* i0_x = i0_y + img;
  __shared__ unsigned long long red[8];
  const int k = blockIdx.y;
  for (int e = threadIdx.x; e < ph * pw; e += blockDim.x) s_low[e] = S[(int64_t)k * ph * pw + e];
  // source coordinates of every padded pixel row / column
  const float scale_h = ds > 1 ? (float)(ph - 1) / (float)(ds - 1) : 0.f;
  const float scale_w = ds > 1 ? (float)(pw - 1) / (float)(ds - 1) : 0.f;
  for (int e = threadIdx.x; e < img; e += blockDim.x) {
    int o = e - pad;
    o = o < 0 ? 0 : (o > ds - 1 ? ds - 1 : o);          // edge padding replicates the border pixel
    const float sy = scale_h * (float)o, sx = scale_w * (float)o;
    const int y0 = (int)sy, x0 = (int)sx;
    i0_y[e] = y0;
    lam_y[e] = sy - (float)y0;
    i0_x[e] = x0;
    lam_x[e] = sx - (float)x0;
  }
  __syncthreads();
  const int64_t npix = (int64_t)img * img;
  const int64_t per = (npix + gridDim.x - 1) / gridDim.x;
  const int64_t begin = (int64_t)blockIdx.x * per;
  const int64_t end = begin + per < npix ? begin + per : npix;
  float bestv = 0.f;
  uint32_t besti = 0xFFFFFFFFu;
  // a thread visits pixels in increasing index order: strict '>' keeps the lowest index
  for (int64_t i = begin + threadIdx.x; i < end; i += blockDim.x) {
    const int y = (int)(i / img), x = (int)(i - (int64_t)y * img);
    const int y0 = i0_y[y], x0 = i0_x[x];
    const int yp = (y0 < ph - 1) ? 1 : 0, xp = (x0 < pw - 1) ? 1 : 0;
    const float h1 = lam_y[y], h0 = 1.f - h1, w1 = lam_x[x], w0 = 1.f - w1;
    const float* r0 = s_low + y0 * pw + x0;
    const float* r1 = r0 + yp * pw;
    const float v = h0 * (w0 * r0[0] + w1 * r0[xp]) + h1 * (w0 * r1[0] + w1 * r1[xp]);
    if (besti == 0xFFFFFFFFu || v > bestv) {
      bestv = v;
      besti = (uint32_t)i;
    }
  }
  unsigned long long key = (besti != 0xFFFFFFFFu) ? ea_key(bestv, besti) : 0ull;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long other = __shfl_xor_sync(0xffffffffu, key, o);
    key = other > key ? other : key;
  }
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = key;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) key = red[w] > key ? red[w] : key;
    if (key != 0ull) atomicMax(&keys[k], key);
  }
}

__global__ void ea_unpack(const unsigned long long* __restrict__ keys, int K, int64_t* __restrict__ idx,
                          float* __restrict__ val) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  const unsigned long long key = keys[k];
  idx[k] = key ? (int64_t)(0xFFFFFFFFu - (uint32_t)(key & 0xFFFFFFFFull)) : -1;
  if (val) {
    uint32_t u = (uint32_t)(key >> 32);
    u = (u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u;
    val[k] = key ? __uint_as_float(u) : __int_as_float(0x7fc00000);
  }
}

struct EaWorkspace {
  float* S;
  unsigned long long* keys;
  size_t total;
};
EaWorkspace carve_ea(void* base, int64_t K, int64_t P) {
  EaWorkspace w{};
  Carver c(base);
  w.S = c.take<float>(K * P);
  w.keys = c.take<unsigned long long>(K);
  w.total = c.total();
  return w;
}

}  // namespace
}  // namespace gd3

using namespace gd3;

extern "C" {

size_t gd3_semantic_argmax_workspace(int64_t K, int64_t ph, int64_t pw) {
  if (K <= 0 || ph <= 0 || pw <= 0) return 0;
  return carve_ea(nullptr, K, ph * pw).total;
}

int gd3_semantic_argmax(const float* kp_desc, int64_t kd_stride_k, int64_t kd_stride_c, const float* desc2, int64_t K,
                        int64_t C, int64_t ph, int64_t pw, int64_t img_size, int64_t patch_size, int64_t stride,
                        int64_t* nn_idx, float* nn_val, void* workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (K == 0) return GD3_OK;
  GD3_REQUIRE(K > 0 && C > 0 && ph > 0 && pw > 0 && img_size > 0 && patch_size > 0 && stride > 0,
              "gd3_semantic_argmax: bad sizes K=%lld C=%lld ph=%lld pw=%lld img=%lld patch=%lld stride=%lld", (long long)K,
              (long long)C, (long long)ph, (long long)pw, (long long)img_size, (long long)patch_size, (long long)stride);
  GD3_REQUIRE(kp_desc && desc2 && nn_idx, "gd3_semantic_argmax: null argument");
  GD3_REQUIRE(img_size * img_size < (1ll << 31), "gd3_semantic_argmax: image too large");
  // size of the upsampled map and the padding, src/evaluate_timm.py:532-539
  const int64_t ds = ((img_size - patch_size) / stride) * stride + 1;
  const int64_t pad = patch_size / 2;
  GD3_REQUIRE(ds >= 1 && img_size - ds - pad >= 0, "gd3_semantic_argmax: inconsistent image / patch sizes");
  EaWorkspace w = carve_ea(workspace, K, ph * pw);
  if (!workspace || workspace_bytes < w.total) {
    set_error("gd3_semantic_argmax: workspace too small (%zu < %zu)", workspace_bytes, w.total);
    return GD3_ERR_WORKSPACE;
  }
  const int P = (int)(ph * pw);
  {
    const size_t smem = sizeof(float) * C * KB;
    GD3_REQUIRE(smem <= 200 * 1024, "gd3_semantic_argmax: C=%lld too large", (long long)C);
    static SmemOptIn opt;
    GD3_CHECK_CUDA(opt.ensure(ea_lowres_sim, smem));
    dim3 grid((unsigned)ceil_div(P, 128), (unsigned)ceil_div<int64_t>(K, KB));
    GD3_PROF("ea_lowres_sim", stream);
    ea_lowres_sim<<<grid, 128, smem, stream>>>(kp_desc, kd_stride_k, kd_stride_c, desc2, (int)K, (int)C, P, w.S);
  }
  GD3_CHECK_LAUNCH();
  GD3_CHECK_CUDA(cudaMemsetAsync(w.keys, 0, sizeof(unsigned long long) * K, stream));
  {
    const size_t smem = sizeof(float) * (P + 4 * img_size);
    GD3_REQUIRE(smem <= 200 * 1024, "gd3_semantic_argmax: patch grid too large");
    static SmemOptIn opt;
    GD3_CHECK_CUDA(opt.ensure(ea_upsample_argmax, smem));
    int chunks = (int)ceil_div<int64_t>(4 * num_sms(), K);
    chunks = chunks < 1 ? 1 : chunks;
    dim3 grid((unsigned)chunks, (unsigned)K);
    GD3_PROF("ea_upsample_argmax", stream);
    ea_upsample_argmax<<<grid, 256, smem, stream>>>(w.S, (int)ph, (int)pw, (int)img_size, (int)ds, (int)pad, w.keys);
  }
  GD3_CHECK_LAUNCH();
  {
    GD3_PROF("ea_unpack", stream);
    ea_unpack<<<(unsigned)ceil_div<int64_t>(K, 128), 128, 0, stream>>>(w.keys, (int)K, nn_idx, nn_val);
  }
  GD3_CHECK_LAUNCH();
  return GD3_OK;
}

}  // extern "C"
