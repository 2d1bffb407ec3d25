// Keypoint-side inputs of the losses, batched over pairs:
//   patch mask   get_patch_mask_from_kp_tensor, utils/functions.py:375-399: patch (y // p) * (W // p) + (x // p) of
//                every in-image keypoint is marked; out-of-image keypoints are dropped.
//   keypoint depth   extract_kp_depth, utils/functions.py:348-372: mean of the replicate-padded window x window
//                neighbourhood of the depth map at flat index (y * W + x).long() (window sums in row-major order, then / n).
// One thread per keypoint; the reference does this per pair with a dozen small torch ops.
#include "../../include/gd3.h"
#include "common.cuh"

namespace gd3 {
namespace {

__global__ void kp_prepare_kernel(const float* __restrict__ kp, int P, int K, int H, int W, int patch, int window,
                                  const float* __restrict__ depth, int64_t depth_pair_stride, uint8_t* __restrict__ mask,
                                  float* __restrict__ kp_depth) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (int64_t)P * K) return;
  const int p = (int)(e / K);
  const float x = kp[2 * e], y = kp[2 * e + 1];
  if (mask) {
    // comparisons on the float coordinates, truncation afterwards, like the reference's `.long() // patch_size`
    if (x >= 0.f && x < (float)W && y >= 0.f && y < (float)H) {
      const int pw = W / patch, ph = H / patch;
      const int xi = (int)x / patch, yi = (int)y / patch;
      if (xi < pw && yi < ph) mask[(int64_t)p * ph * pw + yi * pw + xi] = 1;     // (image sizes that are no multiple of the patch)
    }
  }
  if (kp_depth) {
    const float* d = depth + (int64_t)p * depth_pair_stride;
    // the reference gathers at the flat index (y * W + x).long() computed in fp32 (utils/functions.py:366-369), which
    // differs from truncating x and y separately for fractional keypoints; its gather raises for an index outside
    // the map, here such a keypoint gets NaN (no host round trip inside the step)
    const float fidx = __fadd_rn(__fmul_rn(y, (float)W), x);
    const int64_t idx = (int64_t)fidx;       // truncation toward zero like .long()
    if (!(fidx > -1.f) || idx >= (int64_t)H * W) {
      kp_depth[e] = __int_as_float(0x7fc00000);
      return;
    }
    const int yi = (int)(idx / W), xi = (int)(idx - (int64_t)yi * W), r = window / 2;
    float acc = 0.f;
    for (int dy = -r; dy <= r; ++dy) {
      const int yy = min(max(yi + dy, 0), H - 1);
      for (int dx = -r; dx <= r; ++dx) {
        const int xx = min(max(xi + dx, 0), W - 1);
        acc += __ldg(d + (int64_t)yy * W + xx);
      }
    }
    kp_depth[e] = acc / (float)(window * window);
  }
}

}  // namespace
}  // namespace gd3

using namespace gd3;

extern "C" {

int gd3_kp_prepare(const float* kp, int64_t P, int64_t K, int64_t H, int64_t W, int patch, int window, const float* depth,
                   int64_t depth_pair_stride, uint8_t* mask, float* kp_depth, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  GD3_REQUIRE(P >= 0 && K >= 0 && H > 0 && W > 0 && patch > 0, "gd3_kp_prepare: bad sizes");
  GD3_REQUIRE(!kp_depth || (depth && window >= 1 && window % 2 == 1), "gd3_kp_prepare: depth needs a map and an odd window");
  if (mask && P > 0) GD3_CHECK_CUDA(cudaMemsetAsync(mask, 0, (size_t)(P * (H / patch) * (W / patch)), stream));
  if (P == 0 || K == 0 || (!mask && !kp_depth)) return GD3_OK;
  GD3_REQUIRE(kp, "gd3_kp_prepare: null keypoints");
  {
    GD3_PROF("kp_prepare", stream);
    kp_prepare_kernel<<<(unsigned)ceil_div<int64_t>(P * K, 256), 256, 0, stream>>>(kp, (int)P, (int)K, (int)H, (int)W, patch,
                                                                                 window, depth, depth_pair_stride, mask,
                                                                                 kp_depth);
  }
  GD3_CHECK_LAUNCH();
  return GD3_OK;
}

}  // extern "C"
