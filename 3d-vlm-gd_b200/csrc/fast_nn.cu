// K5: brute-force nearest neighbour (dot / l2) with deterministic lowest-index ties.
//
// Replaces the blocked matmul + max loop of mast3r/fast_nn.py:16-70 (bruteforce_reciprocal_nns) by one
// fused kernel: the |A| x |B| score matrix is never written anywhere.  Scores are accumulated in fp32
// with a fixed k-order FFMA chain (so the result does not depend on the tiling), packed together with
// the column index into a 64-bit key and reduced with max():
//     key = orderable(score) << 32 | (0xFFFFFFFF - index)   ->   max score, lowest index on ties,
// which is exactly torch.max / torch.min semantics (and the strict '<' block merge at :60-61).
// The column direction (nn_B) is reduced from the very same accumulators in the same pass, so both directions
// see bit-identical scores.  The query tile stays resident in shared memory and DB tiles are double-buffered
// with cp.async.
#include "../../include/gd3.h"
#include "common.cuh"

namespace gd3 {
namespace {

constexpr int TILE = 128;      // rows of Q and rows of DB per tile
constexpr int KC = 24;         // k-chunk held in shared memory (MASt3R descriptors are 24-d: one chunk)
constexpr int THREADS = 256;   // 16 x 16 threads, 8 x 8 scores each

__device__ __forceinline__ unsigned long long pack_key(float score, uint32_t idx) {
  score = score + 0.0f;        // -0 -> +0 so that signed zeros tie like torch.max
  uint32_t u = __float_as_uint(score);
  u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
  return (static_cast<unsigned long long>(u) << 32) | static_cast<unsigned long long>(0xFFFFFFFFu - idx);
}
__device__ __forceinline__ unsigned long long kmax(unsigned long long a, unsigned long long b) { return a > b ? a : b; }

__device__ __forceinline__ void cp_async4(float* smem_dst, const float* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// (row, k) panel of X starting at row0 -> smem [k][row]; rows past n are zero
__device__ __forceinline__ void load_panel_async(float (*dst)[TILE], const float* __restrict__ X, int n, int row0, int D,
                                                 int k0, int kc) {
  for (int e = threadIdx.x; e < TILE * kc; e += THREADS) {
    const int row = e / kc, k = e - row * kc;
    const int g = row0 + row;
    if (g < n) cp_async4(&dst[k][row], X + (size_t)g * D + k0 + k);
    else dst[k][row] = 0.f;
  }
}

// Q: (nq, D) queries, DB: (ndb, D).  Each CTA: one 128-row query tile x `tiles_per_cta` DB tiles.
// MODE 0: score = q.d ; MODE 1: score = -sqrt(max(|q|^2 + |d|^2 - 2 q.d, 0))
// BOTH: also reduce every tile over its rows -> arg-best query for each DB row (the nn_B direction), from the
// very same accumulators, so both directions see bit-identical scores in a single pass.
template <int MODE, bool BOTH>
__global__ void __launch_bounds__(THREADS, 2)
    nn_tile_kernel(const float* __restrict__ Q, int nq, const float* __restrict__ DB, int ndb, int D, int tiles_per_cta,
                   unsigned long long* __restrict__ keysQ, unsigned long long* __restrict__ keysDB) {
  __shared__ __align__(16) float sq[KC][TILE];
  __shared__ __align__(16) float sd[2][KC][TILE];
  __shared__ float nq2[TILE], nd2[TILE];
  __shared__ unsigned long long colbest[BOTH ? 8 : 1][TILE];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int q0 = blockIdx.x * TILE;
  const int ndb_tiles = ceil_div(ndb, TILE);
  const int t_begin = blockIdx.y * tiles_per_cta;
  const int t_end = min(t_begin + tiles_per_cta, ndb_tiles);
  const bool one_chunk = D <= KC;      // whole descriptor in one chunk: Q stays resident, DB tiles are double-buffered

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
  if (one_chunk && t_begin < t_end) {
    load_panel_async(sq, Q, nq, q0, D, 0, D);
    load_panel_async(sd[0], DB, ndb, t_begin * TILE, D, 0, D);
    cp_async_commit();
  }

  for (int t = t_begin; t < t_end; ++t) {
    const int d0 = t * TILE;
    const int buf = one_chunk ? ((t - t_begin) & 1) : 0;
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
      if (one_chunk) {
        // prefetch the next DB tile into the other buffer, then wait for the current one
        if (t + 1 < t_end) load_panel_async(sd[buf ^ 1], DB, ndb, (t + 1) * TILE, D, 0, D);
        cp_async_commit();
        cp_async_wait<1>();
        __syncthreads();
      } else {
        __syncthreads();
        load_panel_async(sq, Q, nq, q0, D, k0, kc);
        load_panel_async(sd[0], DB, ndb, d0, D, k0, kc);
        cp_async_commit();
        cp_async_wait<0>();
        __syncthreads();
      }
      for (int k = 0; k < kc; ++k) {
        const float4 a0 = *reinterpret_cast<const float4*>(&sq[k][ty * 8]);
        const float4 a1 = *reinterpret_cast<const float4*>(&sq[k][ty * 8 + 4]);
        const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        float b[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) b[c] = sd[buf][k][c * 16 + tx];
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
          for (int c = 0; c < 8; ++c) acc[r][c] = fmaf(a[r], b[c], acc[r][c]);   // fixed k order: tiling-independent
      }
    }
    // fold this tile into the running per-row best (columns in increasing index order) and, for BOTH,
    // into the per-column best of this tile
    unsigned long long cb[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const int d = d0 + c * 16 + tx;
      cb[c] = 0ull;
      if (d < ndb) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          float s = acc[r][c];
          if (MODE == 1) {
            const float d2 = (nq2[ty * 8 + r] + nd2[c * 16 + tx]) - 2.0f * s;
            s = -sqrtf(fmaxf(d2, 0.f));
          }
          best[r] = kmax(best[r], pack_key(s, (uint32_t)d));
          if (BOTH) {
            const int q = q0 + ty * 8 + r;
            if (q < nq) cb[c] = kmax(cb[c], pack_key(s, (uint32_t)q));
          }
        }
      }
    }
    if (BOTH) {
      // the two ty groups of a warp first, then the 8 warps through shared memory, then one atomic per DB row
#pragma unroll
      for (int c = 0; c < 8; ++c) cb[c] = kmax(cb[c], __shfl_xor_sync(0xffffffffu, cb[c], 16));
      const int warp = threadIdx.x >> 5;
      if ((threadIdx.x & 16) == 0) {
#pragma unroll
        for (int c = 0; c < 8; ++c) colbest[warp][c * 16 + tx] = cb[c];
      }
      __syncthreads();
      if (threadIdx.x < TILE) {
        unsigned long long k = colbest[0][threadIdx.x];
#pragma unroll
        for (int w2 = 1; w2 < 8; ++w2) k = kmax(k, colbest[w2][threadIdx.x]);
        const int d = d0 + threadIdx.x;
        if (d < ndb && k != 0ull) atomicMax(&keysDB[d], k);
      }
    }
    if (one_chunk) __syncthreads();   // everyone is done with sd[buf] before the next prefetch overwrites it
  }
  cp_async_wait<0>();
  // reduce across the 16 threads that share a row group (same ty -> a half warp), then one atomic per row
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    unsigned long long k = best[r];
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) k = kmax(k, __shfl_xor_sync(0xffffffffu, k, o));
    const int q = q0 + ty * 8 + r;
    if (tx == 0 && q < nq && k != 0ull) atomicMax(&keysQ[q], k);
  }
}

__global__ void nn_unpack_kernel(const unsigned long long* __restrict__ keys, int64_t* __restrict__ out, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const unsigned long long k = keys[i];
    out[i] = (k == 0ull) ? int64_t(-1) : int64_t(0xFFFFFFFFu - (uint32_t)(k & 0xFFFFFFFFull));
  }
}

// arg-best DB row for every Q row (outQ) and, when keysDB / outDB are given, arg-best Q row for every DB row
int nn_search(const float* Q, int64_t nq, const float* DB, int64_t ndb, int64_t D, int dist, unsigned long long* keysQ,
              int64_t* outQ, unsigned long long* keysDB, int64_t* outDB, cudaStream_t stream) {
  const bool both = outDB != nullptr;
  GD3_CHECK_CUDA(cudaMemsetAsync(keysQ, 0, sizeof(unsigned long long) * nq, stream));
  if (both) GD3_CHECK_CUDA(cudaMemsetAsync(keysDB, 0, sizeof(unsigned long long) * ndb, stream));
  const int q_tiles = (int)ceil_div<int64_t>(nq, TILE);
  const int db_tiles = (int)ceil_div<int64_t>(ndb, TILE);
  // aim at ~4 CTAs per SM over the whole grid so that short query sets still fill the chip
  int y = (int)ceil_div<int64_t>(4 * num_sms(), q_tiles);
  y = y < 1 ? 1 : (y > db_tiles ? db_tiles : y);
  const int tiles_per_cta = ceil_div(db_tiles, y);
  y = ceil_div(db_tiles, tiles_per_cta);
  dim3 grid(q_tiles, y);
  {
    GD3_PROF("nn_tile_kernel", stream);
    if (dist == 0 && both)
      nn_tile_kernel<0, true><<<grid, THREADS, 0, stream>>>(Q, (int)nq, DB, (int)ndb, (int)D, tiles_per_cta, keysQ, keysDB);
    else if (dist == 0)
      nn_tile_kernel<0, false><<<grid, THREADS, 0, stream>>>(Q, (int)nq, DB, (int)ndb, (int)D, tiles_per_cta, keysQ, keysDB);
    else if (both)
      nn_tile_kernel<1, true><<<grid, THREADS, 0, stream>>>(Q, (int)nq, DB, (int)ndb, (int)D, tiles_per_cta, keysQ, keysDB);
    else
      nn_tile_kernel<1, false><<<grid, THREADS, 0, stream>>>(Q, (int)nq, DB, (int)ndb, (int)D, tiles_per_cta, keysQ, keysDB);
  }
  GD3_CHECK_LAUNCH();
  {
    GD3_PROF("nn_unpack_kernel", stream);
    nn_unpack_kernel<<<(int)ceil_div<int64_t>(nq, 256), 256, 0, stream>>>(keysQ, outQ, (int)nq);
  }
  GD3_CHECK_LAUNCH();
  if (both) {
    GD3_PROF("nn_unpack_kernel", stream);
    nn_unpack_kernel<<<(int)ceil_div<int64_t>(ndb, 256), 256, 0, stream>>>(keysDB, outDB, (int)ndb);
    GD3_CHECK_LAUNCH();
  }
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
  if (nn_A) return nn_search(A, nA, B, nB, dim, dist, kA, nn_A, nn_B ? kB : nullptr, nn_B, stream);
  return nn_search(B, nB, A, nA, dim, dist, kB, nn_B, nullptr, nullptr, stream);
}

}  // extern "C"
