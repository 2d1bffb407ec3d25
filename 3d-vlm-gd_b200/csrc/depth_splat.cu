// Point map -> depth image: point_cloud_to_depth, utils/functions.py:218-260.
// Every camera-frame point with z > 0 lands on pixel (rint(x / z * fx + cx), rint(y / z * fy + cy)); the depth of a pixel
// is the mean z of the points that land on it, 0 where none does.  The reference does this with unique + two bincounts;
// here one pass accumulates (sum, count) per pixel with atomics and a second pass divides.  The sum is kept in double so
// that the result does not depend on the order in which the atomics arrive.
#include "../../include/gd3.h"
#include "common.cuh"

namespace gd3 {
namespace {

struct SplatWs {
  double* sum;
  int* hits;
  size_t bytes;
};

SplatWs carve_splat(void* base, int64_t B, int64_t h, int64_t w) {
  Carver c(base);
  SplatWs s;
  s.sum = c.take<double>((size_t)(B * h * w));
  s.hits = c.take<int>((size_t)(B * h * w));
  s.bytes = c.total();
  return s;
}

__global__ void splat_points(const float* __restrict__ pts, int64_t M, const float* __restrict__ intr, int64_t intr_stride,
                             int w, int h, double* __restrict__ sum, int* __restrict__ hits) {
  const int b = blockIdx.y;
  const float* Kb = intr + (int64_t)b * intr_stride;
  const float fx = __ldg(Kb + 0), cx = __ldg(Kb + 2), fy = __ldg(Kb + 4), cy = __ldg(Kb + 5);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < M; i += (int64_t)gridDim.x * blockDim.x) {
    const float* p = pts + ((int64_t)b * M + i) * 3;
    const float x = __ldg(p), y = __ldg(p + 1), z = __ldg(p + 2);
    if (!(z > 0.f)) continue;
    // divide, multiply, add: each rounded on its own like the reference's three torch ops (no FMA contraction), so the
    // pixel a point lands on is bit-identical; rintf = round half to even = torch.round
    const float u = rintf(__fadd_rn(__fmul_rn(__fdiv_rn(x, z), fx), cx));
    const float v = rintf(__fadd_rn(__fmul_rn(__fdiv_rn(y, z), fy), cy));
    if (!(u >= 0.f && u < (float)w && v >= 0.f && v < (float)h)) continue;
    const int64_t pix = (int64_t)b * h * w + (int64_t)v * w + (int64_t)u;
    atomicAdd(sum + pix, (double)z);
    atomicAdd(hits + pix, 1);
  }
}

__global__ void splat_resolve(const double* __restrict__ sum, const int* __restrict__ hits, int64_t n,
                              float* __restrict__ depth) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int c = hits[i];
  depth[i] = c > 0 ? __fdiv_rn((float)sum[i], (float)c) : 0.f;
}

}  // namespace
}  // namespace gd3

using namespace gd3;

extern "C" {

size_t gd3_point_cloud_to_depth_workspace(int64_t B, int64_t h, int64_t w) {
  if (B <= 0 || h <= 0 || w <= 0) return 256;
  return carve_splat(nullptr, B, h, w).bytes;
}

int gd3_point_cloud_to_depth(const float* points, int64_t B, int64_t M, const float* intrinsics, int64_t intr_stride,
                             int64_t w, int64_t h, float* depth, void* workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  GD3_REQUIRE(B >= 0 && M >= 0 && w > 0 && h > 0, "gd3_point_cloud_to_depth: bad sizes");
  GD3_REQUIRE(B <= 65535 && h * w < (1LL << 31), "gd3_point_cloud_to_depth: batch or image too large");
  if (B == 0) return GD3_OK;
  GD3_REQUIRE(depth && workspace, "gd3_point_cloud_to_depth: null output or workspace");
  SplatWs s = carve_splat(workspace, B, h, w);
  GD3_REQUIRE(workspace_bytes >= s.bytes, "gd3_point_cloud_to_depth: workspace too small (%zu < %zu)", workspace_bytes,
              s.bytes);
  GD3_CHECK_CUDA(cudaMemsetAsync(workspace, 0, s.bytes, stream));
  if (M > 0) {
    GD3_REQUIRE(points && intrinsics, "gd3_point_cloud_to_depth: null input");
    const int64_t want = ceil_div<int64_t>(M, 256);
    const int64_t cap = ceil_div<int64_t>(8 * (int64_t)num_sms(), B);
    dim3 grid((unsigned)(want < cap ? want : (cap < 1 ? 1 : cap)), (unsigned)B);
    GD3_PROF("splat_points", stream);
    splat_points<<<grid, 256, 0, stream>>>(points, M, intrinsics, intr_stride, (int)w, (int)h, s.sum, s.hits);
  }
  GD3_CHECK_LAUNCH();
  {
    const int64_t n = B * h * w;
    GD3_PROF("splat_resolve", stream);
    splat_resolve<<<(unsigned)ceil_div<int64_t>(n, 256), 256, 0, stream>>>(s.sum, s.hits, n, depth);
  }
  GD3_CHECK_LAUNCH();
  return GD3_OK;
}

}  // extern "C"
