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

#ifndef NN_THREADS
#define NN_THREADS 256
#endif
constexpr int TILE = 128;      // rows of DB per tile
constexpr int KC = 24;         // k-chunk held in shared memory (MASt3R descriptors are 24-d: one chunk)
constexpr int THREADS = NN_THREADS;   // (THREADS / 16) x 16 threads, 8 x 8 scores each
constexpr int TQ = THREADS / 2;       // rows of Q per CTA
constexpr int NWARP = THREADS / 32;
static_assert(THREADS >= TILE && THREADS % 32 == 0, "the per-tile column stages use one thread per DB row");
// Shared-memory panels (no [k][row] transposition: that made every 4-byte cp.async a 32-way bank conflict):
//   DB panel  sd[row][k]               row-major, row stride KSD = 28 floats: a DB tile arrives as 16-byte cp.async copies
//                                      (3 per thread and tile for 24-d descriptors)
//   Q  panel  sq[row / 2][2 k + row % 2]   two adjacent query rows interleaved, so that one LDS.128 yields the (k, k+1)
//                                      values of a ROW PAIR as two register pairs -> FFMA2 on two scores at once with
//                                      the DB value as the broadcast scalar operand; pair-row stride KSQ = 52 floats.
//                                      The query tile is resident for the whole CTA, so the 4-byte interleaving copies
//                                      are paid once per CTA, not once per tile (round 1 interleaved the DB panel: the
//                                      index arithmetic of its loader cost as many instructions as the FFMA work).
constexpr int KSD = KC + 4;
constexpr int KSQ = 2 * KC + 4;

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
__device__ __forceinline__ void cp_async16(float* smem_dst, const float* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// DB panel: rows [row0, row0 + TILE) x columns [k0, k0 + kc) -> sd[row][k]; rows past n and columns up to the next
// multiple of 4 are zero.  16-byte copies when the rows are 16-byte aligned (vec); KC4 > 0 fixes the number of
// 16-byte pieces per row at compile time (no integer division in the per-tile path).
template <int KC4>
__device__ __forceinline__ void load_db_vec(float (*dst)[KSD], const float* __restrict__ X, int n, int row0, int D, int k0,
                                            int kc4_rt) {
  const int kc4 = KC4 > 0 ? KC4 : kc4_rt;
  for (int e = threadIdx.x; e < TILE * kc4; e += THREADS) {
    const int row = e / kc4, j = e - row * kc4;
    const int g = row0 + row;
    if (g < n) cp_async16(&dst[row][4 * j], X + (size_t)g * D + k0 + 4 * j);
    else *reinterpret_cast<float4*>(&dst[row][4 * j]) = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}
__device__ __forceinline__ void load_db_async(float (*dst)[KSD], const float* __restrict__ X, int n, int row0, int D, int k0,
                                              int kc, bool vec) {
  const int kc4 = (kc + 3) >> 2;
  if (vec) {
    if (kc4 == KC / 4) load_db_vec<KC / 4>(dst, X, n, row0, D, k0, kc4);
    else load_db_vec<0>(dst, X, n, row0, D, k0, kc4);
  } else {
    for (int e = threadIdx.x; e < TILE * 4 * kc4; e += THREADS) {
      const int row = e / (4 * kc4), k = e - row * (4 * kc4);
      const int g = row0 + row;
      if (g < n && k < kc) cp_async4(&dst[row][k], X + (size_t)g * D + k0 + k);
      else dst[row][k] = 0.f;
    }
  }
}
// Q panel, interleaved by row pairs: element (row, k) -> sq[row / 2][2 k + row % 2]
__device__ __forceinline__ void load_q_async(float (*dst)[KSQ], const float* __restrict__ X, int n, int row0, int D,
                                             int k0, int kc, bool vec) {
  const int kp = ((kc + 3) >> 2) << 2;
  if (vec && kp == kc) {
    // 16-byte loads of both rows of a pair, interleaved in registers (synchronous: once per CTA for 24-d descriptors)
    const int kc4 = kc >> 2;
    for (int u = threadIdx.x; u < (TQ / 2) * kc4; u += THREADS) {
      const int p = u / kc4, j = u - p * kc4;
      const int g = row0 + 2 * p;
      float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0;
      if (g < n) v0 = __ldg(reinterpret_cast<const float4*>(X + (size_t)g * D + k0 + 4 * j));
      if (g + 1 < n) v1 = __ldg(reinterpret_cast<const float4*>(X + (size_t)(g + 1) * D + k0 + 4 * j));
      *reinterpret_cast<float4*>(&dst[p][8 * j]) = make_float4(v0.x, v1.x, v0.y, v1.y);
      *reinterpret_cast<float4*>(&dst[p][8 * j + 4]) = make_float4(v0.z, v1.z, v0.w, v1.w);
    }
    return;
  }
  for (int e = threadIdx.x; e < TQ * kp; e += THREADS) {
    const int row = e / kp, k = e - row * kp;
    const int g = row0 + row;
    float* d = &dst[row >> 1][2 * k + (row & 1)];
    if (g < n && k < kc) cp_async4(d, X + (size_t)g * D + k0 + k);
    else *d = 0.f;
  }
}

// dynamic shared memory: sq | sd0 | sd1 | nq2 | nd2 | colbest[NWARP][TILE] (BOTH only)
inline size_t nn_smem_bytes(bool both) {
  return sizeof(float) * ((TQ / 2) * KSQ + 2 * TILE * KSD + TQ + TILE) +
         (both ? sizeof(unsigned long long) * NWARP * TILE : 0);
}

// Q: (nq, D) queries, DB: (ndb, D).  Each CTA: one TQ-row query tile x `tiles_per_cta` DB tiles of 128 rows.
// MODE 0: score = q.d ; MODE 1: score = -sqrt(max(|q|^2 + |d|^2 - 2 q.d, 0))
// BOTH: also reduce every tile over its rows -> arg-best query for each DB row (the nn_B direction), from the
// very same accumulators, so both directions see bit-identical scores in a single pass.
// Thread (ty, tx) owns the four query row pairs ty * 8 + 2 rp + {0, 1} and the eight DB columns c * 16 + tx.
template <int MODE, bool BOTH>
__global__ void __launch_bounds__(THREADS, 512 / THREADS)
    nn_tile_kernel(const float* __restrict__ Q, int nq, const float* __restrict__ DB, int ndb, int D, int tiles_per_cta,
                   unsigned long long* __restrict__ keysQ, unsigned long long* __restrict__ keysDB,
                   const int* __restrict__ nq_dev = nullptr, int nq_min = 0) {
  extern __shared__ __align__(16) unsigned char nn_smem[];
  if (nq_dev) {
    // device-resident loops (gd3_fast_reciprocal_nn): the number of live queries is only known on the device; the
    // grid is sized for the maximum and surplus CTAs leave.  Below nq_min the streaming kernel takes the query.
    nq = *nq_dev;
    if (nq < nq_min || (int)(blockIdx.x * TQ) >= nq) return;
  }
  float (*sq)[KSQ] = reinterpret_cast<float (*)[KSQ]>(nn_smem);
  float (*sd0)[KSD] = reinterpret_cast<float (*)[KSD]>(sq + TQ / 2);
  float (*sd1)[KSD] = sd0 + TILE;
  float* nq2 = reinterpret_cast<float*>(sd1 + TILE);
  float* nd2 = nq2 + TQ;
  unsigned long long (*colbest)[TILE] = reinterpret_cast<unsigned long long (*)[TILE]>(nd2 + TILE);   // [NWARP][TILE] if BOTH
  // lane -> (ty, tx): the 16 lanes an LDS.64 serves together hold 8 different tx (and both ty of the warp), so the
  // DB reads, 8 rows of stride 28 floats x 2 floats, fall into disjoint banks; tx and tx + 8 would collide
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int tx = (lane & 7) | ((lane >> 4) << 3);
  const int ty = 2 * warp + ((lane >> 3) & 1);
  const int q0 = blockIdx.x * TQ;
  const int ndb_tiles = ceil_div(ndb, TILE);
  const int t_begin = blockIdx.y * tiles_per_cta;
  const int t_end = min(t_begin + tiles_per_cta, ndb_tiles);
  const bool one_chunk = D <= KC;      // whole descriptor in one chunk: Q stays resident, DB tiles are double-buffered
  const bool vec = (D % 4 == 0) && (reinterpret_cast<uintptr_t>(DB) % 16 == 0);
  const bool vecq = (D % 4 == 0) && (reinterpret_cast<uintptr_t>(Q) % 16 == 0);

  // running per-row best as (value, index): a thread visits its columns in increasing index order, so a strict
  // '>' keeps the lowest index on ties; keys are only packed for the cross-thread reductions
  float bestv[8];
  uint32_t besti[8];
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    bestv[r] = -INFINITY;
    besti[r] = 0xFFFFFFFFu;      // "no candidate yet"
  }

  if (MODE == 1) {
    // squared norms of this CTA's query rows (sequential k order, one thread per row)
    if (threadIdx.x < TQ) {
      const int q = q0 + threadIdx.x;
      float s = 0.f;
      if (q < nq)
        for (int k = 0; k < D; ++k) { const float v = Q[(size_t)q * D + k]; s = fmaf(v, v, s); }
      nq2[threadIdx.x] = s;
    }
  }
  if (one_chunk && t_begin < t_end) {
    load_q_async(sq, Q, nq, q0, D, 0, D, vecq);
    load_db_async(sd0, DB, ndb, t_begin * TILE, D, 0, D, vec);
    cp_async_commit();
  }

  for (int t = t_begin; t < t_end; ++t) {
    const int d0 = t * TILE;
    const int buf = one_chunk ? ((t - t_begin) & 1) : 0;
    float2 acc[4][8];            // [query row pair][column]
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 8; ++c) acc[r][c] = make_float2(0.f, 0.f);

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
        if (t + 1 < t_end) load_db_async(buf ? sd0 : sd1, DB, ndb, (t + 1) * TILE, D, 0, D, vec);
        cp_async_commit();
        cp_async_wait<1>();
        __syncthreads();
      } else {
        __syncthreads();
        load_q_async(sq, Q, nq, q0, D, k0, kc, vecq);
        load_db_async(sd0, DB, ndb, d0, D, k0, kc, vec);
        cp_async_commit();
        cp_async_wait<0>();
        __syncthreads();
      }
      float (*sdb)[KSD] = buf ? sd1 : sd0;
      // 2 k at a time (columns past kc are zero: fma(0, 0, acc) leaves acc unchanged).  Per score the FFMA chain runs
      // over k in increasing order, so the result does not depend on the tiling; FFMA2 works on the two scores of
      // a row pair with the DB value broadcast.
      auto step = [&](int k) {
        float4 a[4];             // (r0 k, r1 k, r0 k+1, r1 k+1)
#pragma unroll
        for (int r = 0; r < 4; ++r) a[r] = *reinterpret_cast<const float4*>(&sq[ty * 4 + r][2 * k]);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const float2 b = *reinterpret_cast<const float2*>(&sdb[c * 16 + tx][k]);
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            float2 v = acc[r][c];
            v = __ffma2_rn(make_float2(b.x, b.x), make_float2(a[r].x, a[r].y), v);
            v = __ffma2_rn(make_float2(b.y, b.y), make_float2(a[r].z, a[r].w), v);
            acc[r][c] = v;
          }
        }
      };
      for (int k = 0; k < kc; k += 2) step(k);
    }
    // ---- fold this tile into the running per-row best and, for BOTH, into the per-column best of this tile ----
    // Two stages per row / column: the maximum VALUE first (one FMNMX per score), then the lowest index that
    // attains it (compare + select per score, under ONE branch for the rows: the running best rarely improves).
    // This replaced a (value, index) update per score, which cost twice the FFMA work of a 24-d descriptor.
    float sc[8][8];              // [query row][column c]; column inside the tile: c * 16 + tx
    const bool partial = (d0 + TILE > ndb) || (q0 + TQ > nq);      // CTA-uniform
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        float v = (r & 1) ? acc[r >> 1][c].y : acc[r >> 1][c].x;
        if (MODE == 1) {
          const float d2 = (nq2[ty * 8 + r] + nd2[c * 16 + tx]) - 2.0f * v;
          v = -sqrtf(fmaxf(d2, 0.f));
        }
        sc[r][c] = v;
      }
    if (partial) {
      // rows / columns outside the problem never win: -inf (and they are never reported, see the guards below)
#pragma unroll
      for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const bool ok = (d0 + c * 16 + tx < ndb) && (q0 + ty * 8 + r < nq);
          sc[r][c] = ok ? sc[r][c] : -INFINITY;
        }
    }
    if (t == t_begin) {
      // first candidate of every row: with it in place a strict '>' implements "first maximum" even for -inf scores
      const int d = d0 + tx;
#pragma unroll
      for (int r = 0; r < 8; ++r)
        if (d < ndb) {
          bestv[r] = sc[r][0];
          besti[r] = (uint32_t)d;
        }
    }
    float rm[8];
    bool improved = false;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      float m = sc[r][0];
#pragma unroll
      for (int c = 1; c < 8; ++c) m = fmaxf(m, sc[r][c]);
      rm[r] = m;
      improved |= m > bestv[r];
    }
    if (improved) {
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const float m = rm[r];
        int ci = 7;
#pragma unroll
        for (int c = 6; c >= 0; --c) ci = (sc[r][c] == m) ? c : ci;
        const bool up = m > bestv[r];
        besti[r] = up ? (uint32_t)(d0 + ci * 16 + tx) : besti[r];
        bestv[r] = up ? m : bestv[r];
      }
    }
    if (BOTH) {
      unsigned long long cb[8];
      const bool rows_ok = q0 + ty * 8 < nq;       // valid rows are a prefix of the thread's 8 rows
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        float m = sc[0][c];
#pragma unroll
        for (int r = 1; r < 8; ++r) m = fmaxf(m, sc[r][c]);
        int ri = 0;
#pragma unroll
        for (int r = 7; r >= 1; --r) ri = (sc[r][c] == m) ? r : ri;
        ri = (sc[0][c] == m) ? 0 : ri;
        const bool col_ok = d0 + c * 16 + tx < ndb;
        cb[c] = (rows_ok && col_ok) ? pack_key(m, (uint32_t)(q0 + ty * 8 + ri)) : 0ull;
      }
      // the two ty groups of a warp first (lane bit 3), then the 8 warps through shared memory, then one atomic per DB row
#pragma unroll
      for (int c = 0; c < 8; ++c) cb[c] = kmax(cb[c], __shfl_xor_sync(0xffffffffu, cb[c], 8));
      if ((lane & 8) == 0) {
#pragma unroll
        for (int c = 0; c < 8; ++c) colbest[warp][c * 16 + tx] = cb[c];
      }
      __syncthreads();           // also: every warp is done with this DB buffer before the next prefetch overwrites it
      if (threadIdx.x < TILE) {
        unsigned long long k = colbest[0][threadIdx.x];
#pragma unroll
        for (int w2 = 1; w2 < NWARP; ++w2) k = kmax(k, colbest[w2][threadIdx.x]);
        const int d = d0 + threadIdx.x;
        if (d < ndb && k != 0ull) atomicMax(&keysDB[d], k);
      }
      // colbest is rewritten only after the barrier that follows the next tile's cp.async wait
    } else if (one_chunk) {
      __syncthreads();           // everyone is done with this DB buffer before the next prefetch overwrites it
    }
  }
  cp_async_wait<0>();
  // reduce across the 16 threads that share a row group (same ty: lane bits 4, 2, 1, 0), then one atomic per row
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    unsigned long long k = (besti[r] != 0xFFFFFFFFu) ? pack_key(bestv[r], besti[r]) : 0ull;
    k = kmax(k, __shfl_xor_sync(0xffffffffu, k, 16));
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) k = kmax(k, __shfl_xor_sync(0xffffffffu, k, o));
    const int q = q0 + ty * 8 + r;
    if (tx == 0 && q < nq && k != 0ull) atomicMax(&keysQ[q], k);
  }
}

// ------------------------------------------------------------------------------------------
// Few queries against a large DB (the late ping-pong iterations of fast_reciprocal_NNs have a handful of live
// seeds against 196,608 points): one DB row per thread, all queries in shared memory (broadcast reads), per query
// a warp-wide arg-max by two REDUX (max of the orderable score, then min index among the lanes that attain it).
// The FFMA chain per score runs over k in increasing order exactly like the tile kernel: bit-identical scores.
// ------------------------------------------------------------------------------------------
constexpr int STREAM_MAX_D = 128;
__device__ __forceinline__ uint32_t orderable(float score) {
  score = score + 0.0f;
  const uint32_t u = __float_as_uint(score);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
template <int MODE, int NQ>
__global__ void __launch_bounds__(256)
    nn_stream_kernel(const float* __restrict__ Q, int nq, const int* __restrict__ nq_dev, int nq_min,
                     const float* __restrict__ DB, int ndb, int D, unsigned long long* __restrict__ keysQ) {
  extern __shared__ __align__(16) float sQ[];     // [D][NQ], zero padded, then |q|^2 [NQ]
  if (nq_dev) nq = *nq_dev;
  if (nq < nq_min || nq <= 0 || nq > NQ) return;
  float* qn2 = sQ + D * NQ;
  for (int e = threadIdx.x; e < D * NQ; e += blockDim.x) {
    const int k = e / NQ, j = e - k * NQ;
    sQ[e] = (j < nq) ? Q[(size_t)j * D + k] : 0.f;
  }
  if (MODE == 1 && threadIdx.x < NQ) {
    float s2 = 0.f;
    if (threadIdx.x < nq)
      for (int k = 0; k < D; ++k) { const float v = Q[(size_t)threadIdx.x * D + k]; s2 = fmaf(v, v, s2); }
    qn2[threadIdx.x] = s2;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const bool vec = (D % 4 == 0) && (reinterpret_cast<uintptr_t>(DB) % 16 == 0);
  constexpr int OWN = (NQ + 31) / 32;
  unsigned long long best[OWN];
#pragma unroll
  for (int o = 0; o < OWN; ++o) best[o] = 0ull;
  for (int64_t base = (int64_t)blockIdx.x * blockDim.x; base < ndb; base += (int64_t)gridDim.x * blockDim.x) {
    const int64_t d = base + threadIdx.x;
    const bool ok = d < ndb;
    const float* row = DB + (size_t)(ok ? d : 0) * D;
    float acc[NQ];
#pragma unroll
    for (int j = 0; j < NQ; ++j) acc[j] = 0.f;
    float dn2 = 0.f;
    for (int k = 0; k < D; k += 4) {
      float dv[4];
      if (vec) {
        const float4 t4 = __ldg(reinterpret_cast<const float4*>(row + k));
        dv[0] = t4.x; dv[1] = t4.y; dv[2] = t4.z; dv[3] = t4.w;
      } else {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) dv[kk] = (k + kk < D) ? __ldg(row + k + kk) : 0.f;
      }
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        if (k + kk < D) {
          if (MODE == 1) dn2 = fmaf(dv[kk], dv[kk], dn2);
          const float4* qk = reinterpret_cast<const float4*>(sQ + (k + kk) * NQ);
#pragma unroll
          for (int j = 0; j < NQ; j += 4) {
            const float4 q4 = qk[j >> 2];
            acc[j] = fmaf(q4.x, dv[kk], acc[j]);
            acc[j + 1] = fmaf(q4.y, dv[kk], acc[j + 1]);
            acc[j + 2] = fmaf(q4.z, dv[kk], acc[j + 2]);
            acc[j + 3] = fmaf(q4.w, dv[kk], acc[j + 3]);
          }
        }
      }
    }
#pragma unroll
    for (int j = 0; j < NQ; ++j) {
      if (j < nq) {                     // uniform
        float sc = acc[j];
        if (MODE == 1) sc = -sqrtf(fmaxf((qn2[j] + dn2) - 2.0f * sc, 0.f));
        const uint32_t u = ok ? orderable(sc) : 0u;
        const uint32_t umax = __reduce_max_sync(0xffffffffu, u);
        const uint32_t cand = (ok && u == umax) ? (uint32_t)d : 0xFFFFFFFFu;
        const uint32_t imin = __reduce_min_sync(0xffffffffu, cand);
        if (imin != 0xFFFFFFFFu && lane == (j & 31)) {
          const unsigned long long key = ((unsigned long long)umax << 32) | (unsigned long long)(0xFFFFFFFFu - imin);
          best[j >> 5] = kmax(best[j >> 5], key);
        }
      }
    }
  }
#pragma unroll
  for (int o = 0; o < OWN; ++o) {
    const int j = lane + 32 * o;
    if (j < nq && best[o] != 0ull) atomicMax(&keysQ[j], best[o]);
  }
}

// ------------------------------------------------------------------------------------------
// Device-resident ping-pong of fast_reciprocal_NNs (mast3r/fast_nn.py:147-170): the seed state never leaves the
// GPU and no iteration synchronises with the host.  frn_step (one block) applies the result of the previous query,
// optionally rolls the "old" arrays, compacts the live seeds (in increasing seed order) and gathers their
// descriptors into a dense query buffer for the next nearest-neighbour search.
// ------------------------------------------------------------------------------------------
struct FrnState {
  int32_t *xy1, *xy2, *old1, *old2;
  uint8_t* notyet;
  int32_t* list;              // live seeds of the current query
  int* n_live;
  unsigned long long* keys;   // (n_seeds) result keys of the current query
  float* qbuf;                // (n_seeds, D) gathered query descriptors
  int n_seeds, D;
};

__global__ void frn_init(FrnState st, const int32_t* __restrict__ seeds) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < st.n_seeds) {
    st.xy1[s] = seeds[s];
    st.old1[s] = seeds[s];
    st.xy2[s] = -1;
    st.old2[s] = -1;
    st.notyet[s] = 1;
  }
  if (s == 0) *st.n_live = 0;
}

// apply: 0 nothing, 1 previous query answered xy2 (its "old" array is old2), 2 previous query answered xy1
// gather_from: 0 no new query (final call), 1 next query uses pts[xy1] (pts = pts1), 2 next query uses pts[xy2]
__global__ void __launch_bounds__(1024)
    frn_step(FrnState st, int apply, int roll, int gather_from, const float* __restrict__ pts, uint8_t* converged) {
  __shared__ int warp_tot[32];
  __shared__ int s_base;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int n = st.n_seeds;
  if (apply) {
    int32_t* dst = (apply == 1) ? st.xy2 : st.xy1;
    const int32_t* old = (apply == 1) ? st.old2 : st.old1;
    const int n_prev = *st.n_live;
    for (int i = tid; i < n_prev; i += blockDim.x) {
      const int s = st.list[i];
      const unsigned long long k = st.keys[i];
      const int32_t nn = (k == 0ull) ? -1 : (int32_t)(0xFFFFFFFFu - (uint32_t)(k & 0xFFFFFFFFull));
      dst[s] = nn;
      if (old[s] == nn) st.notyet[s] = 0;     // converged: the neighbour did not change (fast_nn.py:157,164)
    }
    __syncthreads();
  }
  if (roll) {
    for (int s = tid; s < n; s += blockDim.x) {
      st.old1[s] = st.xy1[s];
      st.old2[s] = st.xy2[s];
    }
    __syncthreads();
  }
  if (converged)
    for (int s = tid; s < n; s += blockDim.x) converged[s] = st.notyet[s] ? 0 : 1;
  if (!gather_from) return;
  // ordered compaction of the live seeds
  if (tid == 0) s_base = 0;
  __syncthreads();
  for (int s0 = 0; s0 < n; s0 += blockDim.x) {
    const int s = s0 + tid;
    const int live = (s < n && st.notyet[s]) ? 1 : 0;
    const unsigned bal = __ballot_sync(0xffffffffu, live);
    if (lane == 0) warp_tot[wid] = __popc(bal);
    __syncthreads();
    int before = s_base;
    for (int w = 0; w < wid; ++w) before += warp_tot[w];
    if (live) st.list[before + __popc(bal & ((1u << lane) - 1u))] = s;
    __syncthreads();
    if (tid == 0) {
      int tot = 0;
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += warp_tot[w];
      s_base += tot;
    }
    __syncthreads();
  }
  const int n_live = s_base;
  if (tid == 0) *st.n_live = n_live;
  const int32_t* src = (gather_from == 1) ? st.xy1 : st.xy2;
  for (int e = tid; e < n_live * st.D; e += blockDim.x) {
    const int i = e / st.D, k = e - i * st.D;
    st.qbuf[e] = pts[(size_t)src[st.list[i]] * st.D + k];
  }
  for (int i = tid; i < n; i += blockDim.x) st.keys[i] = 0ull;
}

// keys -> indices for both directions in one launch (outB may be null)
__global__ void nn_unpack_kernel(const unsigned long long* __restrict__ keysA, int64_t* __restrict__ outA, int nA,
                                 const unsigned long long* __restrict__ keysB, int64_t* __restrict__ outB, int nB) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned long long* keys = keysA;
  int64_t* out = outA;
  if (i >= nA) {
    i -= nA;
    keys = keysB;
    out = outB;
    if (!out || i >= nB) return;
  }
  const unsigned long long k = keys[i];
  out[i] = (k == 0ull) ? int64_t(-1) : int64_t(0xFFFFFFFFu - (uint32_t)(k & 0xFFFFFFFFull));
}

constexpr int STREAM_SMALL = 16, STREAM_BIG = 64;      // query-count classes of the streaming kernel

template <int NQ>
int launch_stream(const float* Q, int nq, const int* nq_dev, int nq_min, const float* DB, int64_t ndb, int64_t D, int dist,
                  unsigned long long* keysQ, cudaStream_t stream) {
  const size_t smem = sizeof(float) * ((size_t)D * NQ + NQ);
  int blocks = (int)ceil_div<int64_t>(ndb, 256);
  const int cap = 8 * num_sms();
  blocks = blocks > cap ? cap : blocks;
  GD3_PROF("nn_stream_kernel", stream);
  if (dist == 0) nn_stream_kernel<0, NQ><<<blocks, 256, smem, stream>>>(Q, nq, nq_dev, nq_min, DB, (int)ndb, (int)D, keysQ);
  else nn_stream_kernel<1, NQ><<<blocks, 256, smem, stream>>>(Q, nq, nq_dev, nq_min, DB, (int)ndb, (int)D, keysQ);
  GD3_CHECK_LAUNCH();
  return GD3_OK;
}
inline bool stream_ok(int64_t D) { return D <= STREAM_MAX_D; }

// arg-best DB row for every Q row (outQ) and, when keysDB / outDB are given, arg-best Q row for every DB row
int nn_search(const float* Q, int64_t nq, const float* DB, int64_t ndb, int64_t D, int dist, unsigned long long* keysQ,
              int64_t* outQ, unsigned long long* keysDB, int64_t* outDB, cudaStream_t stream) {
  const bool both = outDB != nullptr;
  GD3_CHECK_CUDA(cudaMemsetAsync(keysQ, 0, sizeof(unsigned long long) * nq, stream));
  if (!both && nq <= STREAM_BIG && stream_ok(D) && ndb >= 4096) {
    // a handful of queries against a large DB: streaming kernel (see nn_stream_kernel)
    int rc = (nq <= STREAM_SMALL)
                 ? launch_stream<STREAM_SMALL>(Q, (int)nq, nullptr, 1, DB, ndb, D, dist, keysQ, stream)
                 : launch_stream<STREAM_BIG>(Q, (int)nq, nullptr, 1, DB, ndb, D, dist, keysQ, stream);
    if (rc) return rc;
    GD3_PROF("nn_unpack_kernel", stream);
    nn_unpack_kernel<<<(int)ceil_div<int64_t>(nq, 256), 256, 0, stream>>>(keysQ, outQ, (int)nq, nullptr, nullptr, 0);
    GD3_CHECK_LAUNCH();
    return GD3_OK;
  }
  if (both) GD3_CHECK_CUDA(cudaMemsetAsync(keysDB, 0, sizeof(unsigned long long) * ndb, stream));
  const int q_tiles = (int)ceil_div<int64_t>(nq, TQ);
  const int db_tiles = (int)ceil_div<int64_t>(ndb, TILE);
  // aim at ~4 CTAs of 256 threads per SM over the whole grid so that short query sets still fill the chip
  int y = (int)ceil_div<int64_t>(4 * (256 / THREADS) * num_sms(), q_tiles);
  y = y < 1 ? 1 : (y > db_tiles ? db_tiles : y);
  const int tiles_per_cta = ceil_div(db_tiles, y);
  y = ceil_div(db_tiles, tiles_per_cta);
  dim3 grid(q_tiles, y);
  const size_t smem = nn_smem_bytes(both);
  {
    static SmemOptIn opt[4];
    GD3_CHECK_CUDA(opt[0].ensure(nn_tile_kernel<0, true>, nn_smem_bytes(true)));
    GD3_CHECK_CUDA(opt[1].ensure(nn_tile_kernel<1, true>, nn_smem_bytes(true)));
    GD3_CHECK_CUDA(opt[2].ensure(nn_tile_kernel<0, false>, nn_smem_bytes(false)));
    GD3_CHECK_CUDA(opt[3].ensure(nn_tile_kernel<1, false>, nn_smem_bytes(false)));
  }
  {
    GD3_PROF("nn_tile_kernel", stream);
    if (dist == 0 && both)
      nn_tile_kernel<0, true><<<grid, THREADS, smem, stream>>>(Q, (int)nq, DB, (int)ndb, (int)D, tiles_per_cta, keysQ, keysDB);
    else if (dist == 0)
      nn_tile_kernel<0, false><<<grid, THREADS, smem, stream>>>(Q, (int)nq, DB, (int)ndb, (int)D, tiles_per_cta, keysQ, keysDB);
    else if (both)
      nn_tile_kernel<1, true><<<grid, THREADS, smem, stream>>>(Q, (int)nq, DB, (int)ndb, (int)D, tiles_per_cta, keysQ, keysDB);
    else
      nn_tile_kernel<1, false><<<grid, THREADS, smem, stream>>>(Q, (int)nq, DB, (int)ndb, (int)D, tiles_per_cta, keysQ, keysDB);
  }
  GD3_CHECK_LAUNCH();
  {
    GD3_PROF("nn_unpack_kernel", stream);
    const int64_t n = nq + (both ? ndb : 0);
    nn_unpack_kernel<<<(int)ceil_div<int64_t>(n, 256), 256, 0, stream>>>(keysQ, outQ, (int)nq, keysDB, outDB,
                                                                       both ? (int)ndb : 0);
  }
  GD3_CHECK_LAUNCH();
  return GD3_OK;
}

// arg-best DB row for each of the first *nq_dev rows of Q (at most nq_max); keys must be zero on entry.  Every
// kernel class is launched and the ones whose query-count range does not contain *nq_dev return at once.
// The host only knows bounds nq_lo <= *nq_dev <= nq_max; classes outside the bounds are not launched at all.
int nn_query_dev(const float* Q, int nq_lo, int nq_max, const int* nq_dev, const float* DB, int64_t ndb, int64_t D,
                 int dist, unsigned long long* keys, cudaStream_t stream) {
  int rc;
  int tile_min = 1;
  if (stream_ok(D)) {
    if (nq_lo <= STREAM_SMALL)
      if ((rc = launch_stream<STREAM_SMALL>(Q, 0, nq_dev, 1, DB, ndb, D, dist, keys, stream))) return rc;
    tile_min = STREAM_SMALL + 1;
    if (nq_max > STREAM_SMALL) {
      if (nq_lo <= STREAM_BIG)
        if ((rc = launch_stream<STREAM_BIG>(Q, 0, nq_dev, STREAM_SMALL + 1, DB, ndb, D, dist, keys, stream))) return rc;
      tile_min = STREAM_BIG + 1;
    }
  }
  if (nq_max >= tile_min) {
    const int q_tiles = ceil_div(nq_max, TQ);
    const int db_tiles = (int)ceil_div<int64_t>(ndb, TILE);
    int y = (int)ceil_div<int64_t>(4 * (256 / THREADS) * num_sms(), q_tiles);
    y = y < 1 ? 1 : (y > db_tiles ? db_tiles : y);
    const int tiles_per_cta = ceil_div(db_tiles, y);
    y = ceil_div(db_tiles, tiles_per_cta);
    dim3 grid(q_tiles, y);
    const size_t smem = nn_smem_bytes(false);
    static SmemOptIn opt[2];
    GD3_CHECK_CUDA(opt[0].ensure(nn_tile_kernel<0, false>, smem));
    GD3_CHECK_CUDA(opt[1].ensure(nn_tile_kernel<1, false>, smem));
    GD3_PROF("nn_tile_kernel", stream);
    if (dist == 0)
      nn_tile_kernel<0, false><<<grid, THREADS, smem, stream>>>(Q, 0, DB, (int)ndb, (int)D, tiles_per_cta, keys, nullptr,
                                                              nq_dev, tile_min);
    else
      nn_tile_kernel<1, false><<<grid, THREADS, smem, stream>>>(Q, 0, DB, (int)ndb, (int)D, tiles_per_cta, keys, nullptr,
                                                              nq_dev, tile_min);
    GD3_CHECK_LAUNCH();
  }
  return GD3_OK;
}

struct FrnWorkspace {
  FrnState st;
  size_t total;
};
FrnWorkspace carve_frn(void* base, int64_t n_seeds, int64_t D) {
  FrnWorkspace w{};
  Carver c(base);
  w.st.old1 = c.take<int32_t>(n_seeds);
  w.st.old2 = c.take<int32_t>(n_seeds);
  w.st.notyet = c.take<uint8_t>(n_seeds);
  w.st.list = c.take<int32_t>(n_seeds);
  w.st.n_live = c.take<int>(1);
  w.st.keys = c.take<unsigned long long>(n_seeds);
  w.st.qbuf = c.take<float>(n_seeds * D);
  w.st.n_seeds = (int)n_seeds;
  w.st.D = (int)D;
  w.total = c.total();
  return w;
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

size_t gd3_fast_reciprocal_nn_workspace(int64_t n_seeds, int64_t dim) {
  if (n_seeds <= 0 || dim <= 0) return 0;
  return carve_frn(nullptr, n_seeds, dim).total;
}

int gd3_fast_reciprocal_nn(const float* pts1, int64_t n1, const float* pts2, int64_t n2, int64_t dim, int dist,
                           const int32_t* seeds, int64_t n_seeds, int max_iter, int host_poll, int32_t* xy1,
                           int32_t* xy2, uint8_t* converged, void* workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  GD3_REQUIRE(dist == GD3_DIST_DOT || dist == GD3_DIST_L2, "Unknown dist=%d", dist);
  GD3_REQUIRE(n1 > 0 && n2 > 0 && dim > 0 && n_seeds >= 0 && max_iter >= 1,
              "gd3_fast_reciprocal_nn: bad sizes n1=%lld n2=%lld dim=%lld seeds=%lld max_iter=%d", (long long)n1,
              (long long)n2, (long long)dim, (long long)n_seeds, max_iter);
  GD3_REQUIRE(n1 < (1ll << 31) && n2 < (1ll << 31) && n_seeds < (1ll << 31), "gd3_fast_reciprocal_nn: more than 2^31 rows");
  if (n_seeds == 0) return GD3_OK;
  GD3_REQUIRE(pts1 && pts2 && seeds && xy1 && xy2 && converged, "gd3_fast_reciprocal_nn: null argument");
  FrnWorkspace w = carve_frn(workspace, n_seeds, dim);
  if (!workspace || workspace_bytes < w.total) {
    set_error("gd3_fast_reciprocal_nn: workspace too small (%zu < %zu)", workspace_bytes, w.total);
    return GD3_ERR_WORKSPACE;
  }
  FrnState st = w.st;
  st.xy1 = xy1;
  st.xy2 = xy2;
  int rc;
  {
    GD3_PROF("frn_init", stream);
    frn_init<<<(unsigned)ceil_div<int64_t>(n_seeds, 256), 256, 0, stream>>>(st, seeds);
  }
  GD3_CHECK_LAUNCH();
  {
    GD3_PROF("frn_step", stream);
    frn_step<<<1, 1024, 0, stream>>>(st, 0, 0, 1, pts1, nullptr);      // live list + queries of the first 1 -> 2 search
  }
  GD3_CHECK_LAUNCH();
  int n_upper = (int)n_seeds;      // upper bound of the live seeds (refreshed by the host poll)
  for (int it = 0; it < max_iter; ++it) {
    const bool last = it + 1 == max_iter;
    // 1 -> 2: query tree2 with pts1[xy1[live]]
    // (in the first round every seed is live in both searches: old_xy2 is -1, so the 1 -> 2 answer retires nobody)
    const int n_lower = it == 0 ? n_upper : 0;
    if ((rc = nn_query_dev(st.qbuf, n_lower, n_upper, st.n_live, pts2, n2, dim, dist, st.keys, stream))) return rc;
    {
      GD3_PROF("frn_step", stream);
      frn_step<<<1, 1024, 0, stream>>>(st, 1, 0, 2, pts2, nullptr);
    }
    GD3_CHECK_LAUNCH();
    // 2 -> 1: query tree1 with pts2[xy2[live]]
    if ((rc = nn_query_dev(st.qbuf, n_lower, n_upper, st.n_live, pts1, n1, dim, dist, st.keys, stream))) return rc;
    // apply, and unless this was the last round roll the "old" arrays (fast_nn.py:166-170) and prepare the next search
    {
      GD3_PROF("frn_step", stream);
      frn_step<<<1, 1024, 0, stream>>>(st, 2, last ? 0 : 1, last ? 0 : 1, pts1, last ? converged : nullptr);
    }
    GD3_CHECK_LAUNCH();
    if (!last && host_poll && it >= 1) {
      // a seed can converge in round 0 already (its nearest neighbour maps straight back to it), but never all of
      // them on real data, so the poll is skipped for the first round only to save a synchronisation; from then on
      // ask the device how many seeds are still live
      // (4 bytes + one stream synchronisation per round) instead of launching rounds that have nothing to do
      int h_live = 0;
      GD3_CHECK_CUDA(cudaMemcpyAsync(&h_live, st.n_live, sizeof(int), cudaMemcpyDeviceToHost, stream));
      GD3_CHECK_CUDA(cudaStreamSynchronize(stream));
      if (h_live == 0) {
        GD3_PROF("frn_step", stream);
        frn_step<<<1, 1024, 0, stream>>>(st, 0, 0, 0, nullptr, converged);
        GD3_CHECK_LAUNCH();
        break;
      }
      n_upper = h_live;
    }
  }
  return GD3_OK;
}

}  // extern "C"
