// Teacher cost-volume post-processing of the MASt3R teacher, fused.
//
// Replaces dust3r/dust3r/model.py:346-366: for every decoder layer l the reference takes the pre-softmax
// cross-attention logits of both branches (dust3r/croco/models/blocks.py:163-164), tgt_l and src_l of shape
// (B, heads, N, N), and computes
//     t = mean_h tgt_l,  s = mean_h src_l,  sym = (t + s^T) / 2,  P_l = softmax(sym / temperature, dim=-1),
//     P_l[:, :, 0] = min(P_l)   (global minimum of the layer's tensor, :353-354),
// then tgt_attn_map = mean_l P_l (:363) -- about ten full passes over 2 L heads N^2 floats (1.2 GB per pair at
// N = 1024, L = heads = 12) plus as many temporaries.  Here every logit is read exactly once.
//   tv_layer_rows   grid (row blocks of 32 or 16, L, B): head sums of the tgt rows (coalesced) and of the transposed src
//                   column strip (128-byte segments) meet in a shared 32 x N tile, row softmax, P_l written once,
//                   layer minimum by an ordered-uint atomicMin.
//   tv_mean_layers  out = mean_l P_l with column 0 taken from the layer minima.
// mode GD3_TV_HEAD_MEAN (reciprocity off, :355-359): P_l = mean_h tgt_l without softmax; same column-0 rule.
// mode GD3_TV_PLAIN_MEAN: mean over layers and heads only -- the VGGT teacher's aggregation of its per-block
// attention maps (vggt/models/aggregator.py:273 followed by src/finetune_timm_vggt.py:390-392).
#include "../../include/gd3.h"
#include "common.cuh"

namespace gd3 {
namespace {

constexpr int TV_MAX_LAYERS = 32;

struct TvPtrs {
  const float* tgt[TV_MAX_LAYERS];
  const float* src[TV_MAX_LAYERS];
};

__device__ __forceinline__ uint32_t tv_orderable(float v) {
  const uint32_t u = __float_as_uint(v + 0.0f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float tv_from_orderable(uint32_t u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u);
}

constexpr int TV_THREADS = 512;

// sum over heads of one float4 position, four heads in flight at a time
__device__ __forceinline__ float4 tv_head_sum4(const float* __restrict__ base, int64_t head, int H) {
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
  int h = 0;
  for (; h + 4 <= H; h += 4) {
    const float4 v0 = __ldg(reinterpret_cast<const float4*>(base + (h + 0) * head));
    const float4 v1 = __ldg(reinterpret_cast<const float4*>(base + (h + 1) * head));
    const float4 v2 = __ldg(reinterpret_cast<const float4*>(base + (h + 2) * head));
    const float4 v3 = __ldg(reinterpret_cast<const float4*>(base + (h + 3) * head));
    a.x += v0.x; a.y += v0.y; a.z += v0.z; a.w += v0.w;
    a.x += v1.x; a.y += v1.y; a.z += v1.z; a.w += v1.w;
    a.x += v2.x; a.y += v2.y; a.z += v2.z; a.w += v2.w;
    a.x += v3.x; a.y += v3.y; a.z += v3.z; a.w += v3.w;
  }
  for (; h < H; ++h) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(base + h * head));
    a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
  }
  return a;
}
__device__ __forceinline__ float tv_head_sum1(const float* __restrict__ base, int64_t head, int H) {
  float a = 0.f;
  for (int h = 0; h < H; ++h) a += __ldg(base + h * head);     // same head order as the vector path
  return a;
}

// dynamic smem: tile[TV_ROWS][ldt], ldt odd.  Heads are summed in increasing order h = 0 .. H-1 for every element.
// TV_ROWS = 32 while two CTAs still fit on an SM (N <= 900), else 16 (three CTAs per SM at N = 1024).
template <int TV_ROWS>
__global__ void __launch_bounds__(TV_THREADS)
    tv_layer_rows(TvPtrs ptrs, int B, int H, int N, int ldt, int reciprocity, float inv_temp, float* __restrict__ P,
                  uint32_t* __restrict__ layer_min) {
  extern __shared__ float tile[];
  const int i0 = blockIdx.x * TV_ROWS, l = blockIdx.y, b = blockIdx.z;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  constexpr int NW = TV_THREADS / 32;
  const int64_t head = (int64_t)N * N;
  const float* T = ptrs.tgt[l] + (int64_t)b * H * head;
  const bool vec = (N % 4 == 0) && (reinterpret_cast<uintptr_t>(T) % 16 == 0) &&
                   (!reciprocity || reinterpret_cast<uintptr_t>(ptrs.src[l]) % 16 == 0);
  const int rows = min(TV_ROWS, N - i0);
  // ---- head sums of the tgt rows: consecutive threads along j ----
  if (vec) {
    const int n4 = N >> 2;
    for (int e = threadIdx.x; e < rows * n4; e += TV_THREADS) {
      const int r = e / n4, j = (e - r * n4) * 4;
      const float4 a = tv_head_sum4(T + (int64_t)(i0 + r) * N + j, head, H);
      float* t = tile + r * ldt + j;
      t[0] = a.x; t[1] = a.y; t[2] = a.z; t[3] = a.w;
    }
  } else {
    for (int e = threadIdx.x; e < rows * N; e += TV_THREADS) {
      const int r = e / N, j = e - r * N;
      tile[r * ldt + j] = tv_head_sum1(T + (int64_t)(i0 + r) * N + j, head, H);
    }
  }
  __syncthreads();
  if (reciprocity) {
    // ---- + head sums of src^T: for every j the segment src[h][j][i0 .. i0 + 31] ----
    const float* S = ptrs.src[l] + (int64_t)b * H * head;
    if (vec && rows == TV_ROWS) {
      for (int e = threadIdx.x; e < N * (TV_ROWS / 4); e += TV_THREADS) {
        const int j = e / (TV_ROWS / 4), q = (e - j * (TV_ROWS / 4)) * 4;
        const float4 a = tv_head_sum4(S + (int64_t)j * N + i0 + q, head, H);
        tile[(q + 0) * ldt + j] += a.x;
        tile[(q + 1) * ldt + j] += a.y;
        tile[(q + 2) * ldt + j] += a.z;
        tile[(q + 3) * ldt + j] += a.w;
      }
    } else {
      for (int e = threadIdx.x; e < N * TV_ROWS; e += TV_THREADS) {
        const int j = e / TV_ROWS, r = e - j * TV_ROWS;
        if (r < rows) tile[r * ldt + j] += tv_head_sum1(S + (int64_t)j * N + i0 + r, head, H);
      }
    }
    __syncthreads();
  }
  // ---- per row: mean over heads (and branches), temperature softmax, store, minimum ----
  const float mean_scale = reciprocity ? 0.5f / (float)H : 1.f / (float)H;
  uint32_t vmin = 0xFFFFFFFFu;
  for (int r = w; r < rows; r += NW) {
    const int i = i0 + r;
    float* row = tile + r * ldt;
    float* out = P + (((int64_t)l * B + b) * N + i) * N;
    if (reciprocity) {
      float m = -INFINITY;
      for (int j = lane; j < N; j += 32) {
        const float x = row[j] * mean_scale * inv_temp;
        row[j] = x;
        m = fmaxf(m, x);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
      float s = 0.f;
      for (int j = lane; j < N; j += 32) {
        const float e = expf(row[j] - m);
        row[j] = e;
        s += e;
      }
      s = warp_sum(s);
      for (int j = lane; j < N; j += 32) {
        const float p = row[j] / s;
        out[j] = p;
        vmin = min(vmin, tv_orderable(p));
      }
    } else {
      for (int j = lane; j < N; j += 32) {
        const float p = row[j] * mean_scale;
        out[j] = p;
        vmin = min(vmin, tv_orderable(p));
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) vmin = min(vmin, __shfl_xor_sync(0xffffffffu, vmin, o));
  if (lane == 0 && vmin != 0xFFFFFFFFu) atomicMin(layer_min + l, vmin);
}

__global__ void tv_mean_layers(const float* __restrict__ P, const uint32_t* __restrict__ layer_min, int L, int64_t BNN, int N,
                               int col0_rule, float* __restrict__ out) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= BNN) return;
  const bool col0 = col0_rule && (e % N) == 0;
  float a = 0.f;
  for (int l = 0; l < L; ++l) a += col0 ? tv_from_orderable(layer_min[l]) : P[(int64_t)l * BNN + e];
  out[e] = a / (float)L;
}

struct TvWorkspace {
  float* P;
  uint32_t* layer_min;
  size_t total;
};
TvWorkspace carve_tv(void* base, int64_t L, int64_t B, int64_t N) {
  TvWorkspace w{};
  Carver c(base);
  w.P = c.take<float>(L * B * N * N);
  w.layer_min = c.take<uint32_t>(L);
  w.total = c.total();
  return w;
}

}  // namespace
}  // namespace gd3

using namespace gd3;

extern "C" {

size_t gd3_teacher_volume_workspace(int64_t L, int64_t B, int64_t N) {
  if (L <= 0 || B <= 0 || N <= 0) return 0;
  return carve_tv(nullptr, L, B, N).total;
}

int gd3_teacher_volume(const float* const* tgt_layers, const float* const* src_layers, int64_t L, int64_t B, int64_t H,
                       int64_t N, float temperature, int mode, float* out, void* workspace,
                       size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  GD3_REQUIRE(mode == GD3_TV_RECIPROCAL || mode == GD3_TV_HEAD_MEAN || mode == GD3_TV_PLAIN_MEAN,
              "gd3_teacher_volume: unknown mode %d", mode);
  const int reciprocity = mode == GD3_TV_RECIPROCAL;
  GD3_REQUIRE(L > 0 && L <= TV_MAX_LAYERS && B > 0 && H > 0 && N > 0,
              "gd3_teacher_volume: bad sizes L=%lld B=%lld H=%lld N=%lld (at most %d layers)", (long long)L, (long long)B,
              (long long)H, (long long)N, TV_MAX_LAYERS);
  GD3_REQUIRE(B <= 65535 && L <= 65535, "gd3_teacher_volume: batch too large");
  GD3_REQUIRE(tgt_layers && out && (!reciprocity || src_layers), "gd3_teacher_volume: null argument");
  GD3_REQUIRE(!reciprocity || temperature > 0.f, "gd3_teacher_volume: temperature must be positive");
  TvWorkspace w = carve_tv(workspace, L, B, N);
  if (!workspace || workspace_bytes < w.total) {
    set_error("gd3_teacher_volume: workspace too small (%zu < %zu)", workspace_bytes, w.total);
    return GD3_ERR_WORKSPACE;
  }
  TvPtrs ptrs{};
  for (int l = 0; l < L; ++l) {
    GD3_REQUIRE(tgt_layers[l] && (!reciprocity || src_layers[l]), "gd3_teacher_volume: null layer %d", l);
    ptrs.tgt[l] = tgt_layers[l];
    ptrs.src[l] = reciprocity ? src_layers[l] : nullptr;
  }
  GD3_CHECK_CUDA(cudaMemsetAsync(w.layer_min, 0xFF, sizeof(uint32_t) * L, stream));
  const int ldt = (int)(N | 1);
  const int rows = (sizeof(float) * 32 * ldt <= 113 * 1024) ? 32 : 16;
  const size_t smem = sizeof(float) * rows * ldt;
  GD3_REQUIRE(smem <= 227 * 1024, "gd3_teacher_volume: N=%lld too large for the shared row tile", (long long)N);
  {
    static SmemOptIn opt32, opt16;
    if (rows == 32)
      GD3_CHECK_CUDA(opt32.ensure(tv_layer_rows<32>, smem));
    else
      GD3_CHECK_CUDA(opt16.ensure(tv_layer_rows<16>, smem));
    dim3 grid((unsigned)ceil_div<int64_t>(N, rows), (unsigned)L, (unsigned)B);
    GD3_PROF("tv_layer_rows", stream);
    if (rows == 32)
      tv_layer_rows<32><<<grid, TV_THREADS, smem, stream>>>(ptrs, (int)B, (int)H, (int)N, ldt, reciprocity,
                                                            1.f / temperature, w.P, w.layer_min);
    else
      tv_layer_rows<16><<<grid, TV_THREADS, smem, stream>>>(ptrs, (int)B, (int)H, (int)N, ldt, reciprocity,
                                                            1.f / temperature, w.P, w.layer_min);
  }
  GD3_CHECK_LAUNCH();
  {
    const int64_t BNN = B * N * N;
    GD3_PROF("tv_mean_layers", stream);
    tv_mean_layers<<<(unsigned)ceil_div<int64_t>(BNN, 256), 256, 0, stream>>>(w.P, w.layer_min, (int)L, BNN, (int)N,
                                                                              mode != GD3_TV_PLAIN_MEAN, out);
  }
  GD3_CHECK_LAUNCH();
  return GD3_OK;
}

}  // extern "C"
