// 2-CTA (cta_group::2) variant of the tcgen05 GEMM core -- EXPERIMENTAL, not used by any loss pipeline.
//
// Built and parity-tested through gd3_debug_gemm_bf16(tile_n < 0) only (tests/test_gpu_gemm.py).  On the plain
// 32 x 1024 x 768 x 1024 product it is 9 % faster than the 1-CTA kernel (56.6 vs 62.3 us); behind the KL gradient
// epilogue it measured slower (68.7 vs 65.3 us), so tc_gemm.cuh's 1-CTA kernel serves the product and this
// file is kept out of it.  K-major operands only.
#pragma once

#include "tc_gemm.cuh"

namespace gd3 {
namespace tc {

// ------------------------------------------------------------------ 2-CTA (cta_group::2) variant
// A cluster of two CTAs (one TPC) computes a 256 x BN tile: CTA r owns rows [m0 + 128 r, +128) of A and of
// the accumulator (its own TMEM), and loads only half of the B tile (rows n0 + r BN/2 ...); the leader's
// tcgen05.mma.cta_group::2 reads both halves, so B costs half the L2 -> SM traffic and half the shared
// memory per CTA (more pipeline stages).  Protocol (the usual one for 2-SM UMMA):
//   full[s]   lives in the leader; both CTAs' TMA loads complete_tx on it (peer bit of the mbarrier address
//             cleared), the leader's producer arms it with the bytes of both CTAs
//   empty[s], tmem_full[a]   exist in both CTAs; the leader's MMA thread signals them with a multicast commit
//   tmem_empty[a]            lives in the leader; the epilogue warps of both CTAs arrive on it remotely
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;   // clears the CTA-rank bit of a shared::cluster address -> leader CTA
__device__ __forceinline__ void tma_load_3d_2sm(void* smem_dst, const CUtensorMap* tmap, uint64_t* bar, int c0, int c1,
                                                int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0),
      "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_commit_2sm(uint64_t* bar) {   // arrives on `bar` in BOTH CTAs of the pair
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(smem_u32(bar)), "h"(static_cast<uint16_t>(3))
      : "memory");
}
__device__ __forceinline__ void umma_bf16_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {   // arrive on the leader CTA's copy of `bar`
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, 0;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}"
      ::"r"(smem_u32(bar))
      : "memory");
}

__host__ __device__ constexpr int stage_bytes_2sm(int BN) { return (BM + BN / 2) * BK * 2; }
__host__ __device__ constexpr int num_stages_2sm(int BN, int epi_scratch = 0) {
  const int fit = (kSmemBudget - 1024 - 256 - epi_scratch) / stage_bytes_2sm(BN);
  return fit > 8 ? 8 : fit;
}
__host__ __device__ constexpr int smem_bytes_2sm(int BN, int epi_scratch) {
  return 1024 + num_stages_2sm(BN, epi_scratch) * stage_bytes_2sm(BN) + epi_scratch + 256;
}

template <int BN, int EPI_WARPS, class Epi>
__global__ void __launch_bounds__((PRODUCER_WARPS + EPI_WARPS) * 32, 1)
    tc_gemm2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int tiles_m,
                    int tiles_n, int batch, int k_blocks, typename Epi::Params ep) {
  static_assert(BN % 32 == 0 && BN >= 64 && BN <= 256, "BN");
  static_assert(EPI_WARPS == 4 || EPI_WARPS == 8, "EPI_WARPS");
  constexpr int STAGES = num_stages_2sm(BN, Epi::kScratchBytes);
  constexpr int A_BYTES = BM * BK * 2;
  constexpr int STAGE_BYTES = stage_bytes_2sm(BN);
  constexpr uint32_t TMEM_COLS = (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;

  extern __shared__ uint8_t smem_raw[];
  // offset arithmetic on the __shared__ symbol (not a uintptr_t round trip) keeps the address space visible to the
  // compiler, so epilogue scratch accesses compile to LDS / STS instead of generic LD / ST
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* ring = smem;
  uint8_t* scratch = ring + STAGES * STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(scratch + Epi::kScratchBytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + STAGES;
  uint64_t* tmem_full = bars + 2 * STAGES;
  uint64_t* tmem_empty = bars + 2 * STAGES + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
  const int total_tiles = tiles_m * tiles_n * batch;     // tiles_m counts 256-row tiles

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], 2 * EPI_WARPS);
    }
    fence_barrier_init();
  }
  cluster_sync_all();                       // barriers of both CTAs are initialised before any remote use
  if (warp == 1) tmem_alloc_2sm(tmem_slot, TMEM_COLS);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs; each loads its A rows and its half of B) =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = cluster_id; t < total_tiles; t += num_clusters) {
        const int b = t / (tiles_m * tiles_n);
        const int r = t - b * (tiles_m * tiles_n);
        const int m0 = (r / tiles_n) * (2 * BM) + (int)rank * BM;
        const int n0 = (r % tiles_n) * BN + (int)rank * (BN / 2);
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = ring + stage * STAGE_BYTES;
          uint8_t* sb = sa + A_BYTES;
          if (leader) mbar_expect_tx(&full_bar[stage], 2 * STAGE_BYTES);
          tma_load_3d_2sm(sa, &tmA, &full_bar[stage], kb * BK, m0, b);
          tma_load_3d_2sm(sb, &tmB, &full_bar[stage], kb * BK, n0, b);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (leader && lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(2 * BM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int t = cluster_id; t < total_tiles; t += num_clusters) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BN;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(ring + stage * STAGE_BYTES);
          const uint64_t da = make_smem_desc_k128(sa);
          const uint64_t db = make_smem_desc_k128(sa + A_BYTES);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k)
            umma_bf16_2sm(tmem_d, da + static_cast<uint64_t>(2 * k), db + static_cast<uint64_t>(2 * k), idesc,
                          (kb > 0 || k > 0) ? 1u : 0u);
          tc_commit_2sm(&empty_bar[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        tc_commit_2sm(&tmem_full[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue warps (both CTAs, own 128 rows) =====================
    const int ew = warp - PRODUCER_WARPS;
    const int quad = warp & 3;
    constexpr int PARTS = EPI_WARPS / 4;
    const int part = (PARTS == 1) ? 0 : (ew >> 2);
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int t = cluster_id; t < total_tiles; t += num_clusters) {
      const int b = t / (tiles_m * tiles_n);
      const int r = t - b * (tiles_m * tiles_n);
      EpiCtx cx;
      cx.b = b;
      cx.m0 = (r / tiles_n) * (2 * BM) + (int)rank * BM;
      cx.n0 = (r % tiles_n) * BN;
      cx.row = quad * 32 + lane;
      cx.col_begin = part * (BN / PARTS);
      cx.col_end = cx.col_begin + BN / PARTS;
      cx.lane = lane;
      cx.epi_warp = ew;
      cx.scratch = scratch;
      cx.tmem = tmem_base + acc * BN + (static_cast<uint32_t>(quad * 32) << 16);
      typename Epi::Pre pre;
      Epi::pre(ep, cx, pre);
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      Epi::run(ep, cx, pre);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader(&tmem_empty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  cluster_sync_all();                       // nobody leaves while the peer may still signal / read its smem
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, TMEM_COLS);
  }
}

// 2-CTA launch: tmB must have been built with box_rows = BN / 2 and tmA with box_rows = 128.
template <int BN, int EPI_WARPS, class Epi>
int launch_gemm_2sm(const char* name, const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmShape& s,
                    const typename Epi::Params& ep, cudaStream_t stream) {
  if (s.M <= 0 || s.N <= 0 || s.batch <= 0) return GD3_OK;
  GD3_REQUIRE(s.K > 0, "tc_gemm: K must be positive");
  auto kern = tc_gemm2_kernel<BN, EPI_WARPS, Epi>;
  constexpr int SMEM = smem_bytes_2sm(BN, Epi::kScratchBytes);
  static_assert(SMEM <= 227 * 1024, "tc_gemm2 shared memory budget");
  static SmemOptIn opt;
  GD3_CHECK_CUDA(opt.ensure(kern, SMEM));
  const int tiles_m = ceil_div(s.M, 2 * BM), tiles_n = ceil_div(s.N, BN);
  const long long total = 1LL * tiles_m * tiles_n * s.batch;
  int clusters = num_sms() / 2;
  if (total < clusters) clusters = static_cast<int>(total);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * clusters);
  cfg.blockDim = dim3((PRODUCER_WARPS + EPI_WARPS) * 32);
  cfg.dynamicSmemBytes = SMEM;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  {
    GD3_PROF(name, stream);
    GD3_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, tmA, tmB, tiles_m, tiles_n, s.batch, ceil_div(s.K, BK), ep));
  }
  GD3_CHECK_LAUNCH();
  return GD3_OK;
}

}  // namespace tc
}  // namespace gd3
