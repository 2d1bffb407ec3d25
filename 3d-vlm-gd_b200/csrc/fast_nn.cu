// K5: brute-force nearest neighbour (dot / l2) with deterministic lowest-index ties.
//
// Replaces the blocked matmul + max loop of mast3r/fast_nn.py:16-70 (bruteforce_reciprocal_nns) by one
// fused kernel: the |A| x |B| score matrix is never written anywhere.  Scores are accumulated in fp32
// with a fixed k-order FFMA chain (so the result does not depend on the tiling), packed together with
// the column index into a 64-bit key and reduced with max():
//     key = orderable(score) << 32 | (0xFFFFFFFF - index)   ->   max score, lowest index on ties,
// which is exactly torch.max / torch.min semantics (and the strict '<' block merge at :60-61).
// The column direction (nn_B) is the same kernel with A and B swapped; a*b is commutative in fp32 so
// both directions see bit-identical scores.
#include "../../include/gd3.h"
#include "common.cuh"

namespace gd3 {
namespace {

constexpr int TILE = 128;      // rows of Q and rows of DB per tile
constexpr int KC = 32;         // k-chunk held in shared memory
constexpr int THREADS = 256;   // 16 x 16 threads, 8 x 8 scores each

__device__ __forceinline__ unsigned long long pack_key(float score, uint32_t idx) {
  score = score + 0.0f;        // -0 -> +0 so that signed zeros tie like torch.max
  uint32_t u = __float_as_uint(score);
  u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
  return (static_cast<unsigned long long>(u) << 32) | static_cast<unsigned long long>(0xFFFFFFFFu - idx);
}

// Q: (nq, D) queries, DB: (ndb, D).  Each CTA: one 128-row query tile x `tiles_per_cta` DB tiles.
// MODE 0: score = q.d ; MODE 1: score = -sqrt(max(|q|^2 + |d|^2 - 2 q.d, 0))
template <int MODE>
__global__ void __launch_bounds__(THREADS, 2)
    nn_rows_kernel(const float* __restrict__ Q, int nq, const float* __restrict__ DB, int ndb, int D,
                   int tiles_per_cta, unsigned long long* __restrict__ keys) {
  __shared__ __align__(16) float sq[KC][TILE];
  __shared__ __align__(16) float sd[KC][TILE];
  __shared__ float nq2[TILE], nd2[TILE];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int q0 = blockIdx.x * TILE;
  const int ndb_tiles = ceil_div(ndb, TILE);
  const int t_begin = blockIdx.y * tiles_per_cta;
  const int t_end = min(t_begin + tiles_per_cta, ndb_tiles);

  unsigned long long best[8];
#pragma unroll
  for (int r = 0; r < 8; ++r) best[r] = 0ull;

  if (MODE == 1) {
    // squared norms of this CTA's query rows (sequential k order, one thread per row)
    if (threadIdx.x < TILE) {
      const int q = q0 + threadIdx.x;
      float s = 0.f;
      if (q < nq)
        for (int k = 0; k < D; ++k) { const float v = Q[(size_t)q * D + k]; s = fmaf(v, v, s); }
      nq2[threadIdx.x] = s;
    }
  }

  for (int t = t_begin; t < t_end; ++t) {
    const int d0 = t * TILE;
    float acc[8][8];
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int c = 0; c < 8; ++c) acc[r][c] = 0.f;

    if (MODE == 1) {
      __syncthreads();
      if (threadIdx.x < TILE) {
        const int d = d0 + threadIdx.x;
        float s = 0.f;
        if (d < ndb)
          for (int k = 0; k < D; ++k) { const float v = DB[(size_t)d * D + k]; s = fmaf(v, v, s); }
        nd2[threadIdx.x] = s;
      }
    }

    for (int k0 = 0; k0 < D; k0 += KC) {
      const int kc = min(KC, D - k0);
      __syncthreads();
      // coalesced global reads of the (row, k) panel, transposed into smem [k][row]
      for (int e = threadIdx.x; e < TILE * kc; e += THREADS) {
        const int row = e / kc, k = e - row * kc;
        const int q = q0 + row, d = d0 + row;
        sq[k][row] = (q < nq) ? Q[(size_t)q * D + k0 + k] : 0.f;
        sd[k][row] = (d < ndb) ? DB[(size_t)d * D + k0 + k] : 0.f;
      }
      __syncthreads();
      for (int k = 0; k < kc; ++k) {
        const float4 a0 = *reinterpret_cast<const float4*>(&sq[k][ty * 8]);
        const float4 a1 = *reinterpret_cast<const float4*>(&sq[k][ty * 8 + 4]);
        const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        float b[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) b[c] = sd[k][c * 16 + tx];
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
          for (int c = 0; c < 8; ++c) acc[r][c] = fmaf(a[r], b[c], acc[r][c]);
      }
    }
    // fold this tile into the running per-row best; columns visited in increasing index order
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const int d = d0 + c * 16 + tx;
      if (d < ndb) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          float s = acc[r][c];
          if (MODE == 1) {
            const float d2 = (nq2[ty * 8 + r] + nd2[c * 16 + tx]) - 2.0f * s;
            s = -sqrtf(fmaxf(d2, 0.f));
          }
          const unsigned long long key = pack_key(s, (uint32_t)d);
          best[r] = key > best[r] ? key : best[r];
        }
      }
    }
  }
  // reduce across the 16 threads that share a row group (same ty -> a half warp), then one atomic per row
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    unsigned long long k = best[r];
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {
      const unsigned long long other = __shfl_xor_sync(0xffffffffu, k, o);
      k = other > k ? other : k;
    }
    const int q = q0 + ty * 8 + r;
    if (tx == 0 && q < nq && k != 0ull) atomicMax(&keys[q], k);
  }
}

__global__ void nn_unpack_kernel(const unsigned long long* __restrict__ keys, int64_t* __restrict__ out, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const unsigned long long k = keys[i];
    out[i] = (k == 0ull) ? int64_t(-1) : int64_t(0xFFFFFFFFu - (uint32_t)(k & 0xFFFFFFFFull));
  }
}

int nn_rows(const float* Q, int64_t nq, const float* DB, int64_t ndb, int64_t D, int dist,
            unsigned long long* keys, int64_t* out, cudaStream_t stream) {
  GD3_CHECK_CUDA(cudaMemsetAsync(keys, 0, sizeof(unsigned long long) * nq, stream));
  const int q_tiles = (int)ceil_div<int64_t>(nq, TILE);
  const int db_tiles = (int)ceil_div<int64_t>(ndb, TILE);
  // aim at ~4 CTAs per SM over the whole grid so that short query sets still fill the chip
  int y = (int)ceil_div<int64_t>(4 * num_sms(), q_tiles);
  y = y < 1 ? 1 : (y > db_tiles ? db_tiles : y);
  const int tiles_per_cta = ceil_div(db_tiles, y);
  y = ceil_div(db_tiles, tiles_per_cta);
  dim3 grid(q_tiles, y);
  if (dist == 0)
    {
      GD3_PROF("nn_rows_kernel", stream);
      nn_rows_kernel<0><<<grid, THREADS, 0, stream>>>(Q, (int)nq, DB, (int)ndb, (int)D, tiles_per_cta, keys);
    }
  else
    {
      GD3_PROF("nn_rows_kernel", stream);
      nn_rows_kernel<1><<<grid, THREADS, 0, stream>>>(Q, (int)nq, DB, (int)ndb, (int)D, tiles_per_cta, keys);
    }
  GD3_CHECK_LAUNCH();
  {
    GD3_PROF("nn_unpack_kernel", stream);
    nn_unpack_kernel<<<(int)ceil_div<int64_t>(nq, 256), 256, 0, stream>>>(keys, out, (int)nq);
  }
  GD3_CHECK_LAUNCH();
  return GD3_OK;
}

}  // namespace
}  // namespace gd3

using namespace gd3;

extern "C" {

size_t gd3_reciprocal_nn_workspace(int64_t nA, int64_t nB) {
  Carver c(nullptr);
  c.take<unsigned long long>(nA > 0 ? nA : 0);
  c.take<unsigned long long>(nB > 0 ? nB : 0);
  return c.total();
}

int gd3_reciprocal_nn(const float* A, int64_t nA, const float* B, int64_t nB, int64_t dim, int dist, int64_t* nn_A,
                      int64_t* nn_B, void* workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  GD3_REQUIRE(dist == GD3_DIST_DOT || dist == GD3_DIST_L2, "Unknown dist=%d", dist);
  GD3_REQUIRE(nA >= 0 && nB >= 0 && dim > 0, "gd3_reciprocal_nn: bad sizes nA=%lld nB=%lld dim=%lld",
              (long long)nA, (long long)nB, (long long)dim);
  GD3_REQUIRE(nA < (1ll << 31) && nB < (1ll << 31), "gd3_reciprocal_nn: more than 2^31 rows");
  if (nA == 0 || nB == 0) {
    // nothing to match against: indices are -1 like the reference's initial fill
    if (nn_A && nA) GD3_CHECK_CUDA(cudaMemsetAsync(nn_A, 0xFF, sizeof(int64_t) * nA, stream));
    if (nn_B && nB) GD3_CHECK_CUDA(cudaMemsetAsync(nn_B, 0xFF, sizeof(int64_t) * nB, stream));
    return GD3_OK;
  }
  GD3_REQUIRE(A && B, "gd3_reciprocal_nn: null input");
  GD3_REQUIRE(nn_A || nn_B, "gd3_reciprocal_nn: no output requested");
  if (workspace_bytes < gd3_reciprocal_nn_workspace(nA, nB) || !workspace) {
    set_error("gd3_reciprocal_nn: workspace too small (%zu < %zu)", workspace_bytes,
              gd3_reciprocal_nn_workspace(nA, nB));
    return GD3_ERR_WORKSPACE;
  }
  Carver c(workspace);
  unsigned long long* kA = c.take<unsigned long long>(nA);
  unsigned long long* kB = c.take<unsigned long long>(nB);
  int rc = GD3_OK;
  if (nn_A && (rc = nn_rows(A, nA, B, nB, dim, dist, kA, nn_A, stream))) return rc;
  if (nn_B && (rc = nn_rows(B, nB, A, nA, dim, dist, kB, nn_B, stream))) return rc;
  return GD3_OK;
}

}  // extern "C"
