// Persistent, warp-specialised tcgen05 GEMM for sm_100a with a pluggable epilogue.
//
//   D[b][m][n] = sum_k A[b][m][k] * B[b][n][k]        (both operands K-major bf16, fp32 accumulate)
//
// One CTA per SM, static round-robin over (batch, m-tile, n-tile).  Roles:
//   warp 0      : TMA producer  (cp.async.bulk.tensor.3d -> 128B-swizzled smem ring, mbarrier tx)
//   warp 1      : MMA issuer    (one thread issues tcgen05.mma.cta_group::1.kind::f16, 128 x BN x 16)
//                 + owner of the TMEM allocation (2 accumulator buffers of BN columns)
//   warps 2..   : epilogue      (tcgen05.ld 32x32b -> registers -> Epi functor); the accumulator is
//                 double-buffered in TMEM so the epilogue of tile t overlaps the MMAs of tile t+1.
//                 Epi::pre() runs before the accumulator is awaited (prefetch of epilogue operands),
//                 Epi::run() after.
// The accumulator never goes to HBM unless the epilogue functor writes it.
//
// Descriptor bit layouts follow the PTX ISA "tcgen05 matrix/instruction descriptor" tables (the
// CuTe headers cute/arch/mma_sm100_desc.hpp document the same fields).
#pragma once

#include "common.cuh"

namespace gd3 {
namespace tc {

constexpr int BM = 128;          // UMMA M (rows of A per tile) == TMEM lanes
constexpr int BK = 64;           // K elements per stage = one 128-byte swizzle atom of bf16
constexpr int UMMA_K = 16;       // K per tcgen05.mma for 16-bit inputs
constexpr int PRODUCER_WARPS = 2;  // warp 0 = TMA, warp 1 = MMA

__host__ __device__ constexpr int stage_bytes(int BN) { return (BM + BN) * BK * 2; }
constexpr int kSmemBudget = 227 * 1024;
// as many stages as fit next to the epilogue scratch (at most 8): 1024 B alignment slack + ring + scratch + barriers
__host__ __device__ constexpr int num_stages(int BN, int epi_scratch = 0) {
  const int fit = (kSmemBudget - 1024 - 256 - epi_scratch) / stage_bytes(BN);
#ifdef GD3_MAX_STAGES
  return fit > GD3_MAX_STAGES ? GD3_MAX_STAGES : fit;      // experiment knob: pipeline-depth sensitivity
#else
  return fit > 8 ? 8 : fit;
#endif
}
__host__ __device__ constexpr int smem_bytes(int BN, int epi_scratch) {
  return 1024 + num_stages(BN, epi_scratch) * stage_bytes(BN) + epi_scratch + 256;
}

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug traps (-> CUDA error on the host) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 26)) __trap();
  }
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* tmap, uint64_t* bar, int c0,
                                            int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(c1),
      "r"(c2)
      : "memory");
}

// TMA store of a shared-memory box (written by the generic proxy: fence_proxy_async_smem + a warp sync come first)
// into a 3-D tensor map; bulk-group completion.  wait_read: the shared-memory source may be overwritten again.
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* tmap, const void* smem_src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(tmap)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, 128 x N x 16, bf16 inputs, fp32 accumulate
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// 32 lanes x 32 consecutive fp32 columns: thread t of the warp gets lane (base_lane + t)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major, 128B-swizzled shared-memory matrix descriptor (PTX ISA "shared memory descriptor"):
//   [0,14)  start address >> 4      [16,30) leading byte offset >> 4 (unused for swizzled K-major, 1)
//   [32,46) stride byte offset >> 4 (8 rows x 128 B = 1024)      [46,48) version = 1 (Blackwell)
//   [61,64) layout type: 2 = SWIZZLE_128B
__device__ __forceinline__ uint64_t make_smem_desc_k128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
// K-major, 64B-swizzled operand (32 bf16 per row, boxes from make_tmap_bf16_k32): 8 rows x 64 B = 512 B between row groups
__device__ __forceinline__ uint64_t make_smem_desc_k64(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(512 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(4) << 61;
  return d;
}
// MN-major, 128B-swizzled operand: the tile sits in shared memory as [k rows][64 MN elements] blocks (128-byte rows,
// exactly how TMA delivers a box of a matrix whose MN dimension is the contiguous one), one 8 KB block per 64 MN
// elements.  Canonical layout ((8,n),(8,k)) : ((1,LBO),(8,SBO)) in 16-byte units: LBO = distance between 64-element MN
// blocks (64 k rows x 128 B = 8192), SBO = distance between groups of 8 k rows (1024).  One MMA consumes 16 k rows:
// the start address advances by 2048 bytes per UMMA_K step.
__device__ __forceinline__ uint64_t make_smem_desc_mn128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(8192 >> 4) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
// kind::f16 instruction descriptor: fp32 D, bf16 A/B, dense; bit 15 / 16 = A / B is MN-major
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N, bool a_mn = false, bool b_mn = false) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (a_mn ? (1u << 15) : 0u) | (b_mn ? (1u << 16) : 0u) |
         (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(M >> 4) << 24);
}

// ------------------------------------------------------------------ grouped contraction (split-K over sets)
// For C[g] = sum over the sets s of group g and over `panels` precision panels of A_s^T B_s, where A_s / B_s are
// row-major (rows = contraction index, i.e. both operands MN-major).  The k-blocks of a tile enumerate
// (panel, set in group, 64-row block of the set); the tensor maps are (inner = all panels side by side, rows of one
// set, sets), so rows past the end of a set and sets past the last one read as zeros.
struct GroupedK {
  int panels = 0;            // 0 = plain contraction
  int sets_per_group = 1;
  int row_blocks = 1;        // ceil(rows per set / 64)
  int a_off[3] = {0, 0, 0};  // inner-coordinate offset of the A panel used by term p
  int b_off[3] = {0, 0, 0};
};

// ------------------------------------------------------------------ epilogue context
struct EpiCtx {
  int b, m0, n0;        // batch index and tile origin
  uint32_t tmem;        // TMEM address of this tile's accumulator, lane field already set to the warp's quadrant
  int row;              // this thread's row inside the tile (0..127) == TMEM lane
  int col_begin;        // this warp's column range inside the tile [col_begin, col_end)
  int col_end;
  int lane;             // lane id
  int epi_warp;         // 0..EPI_WARPS-1
  uint8_t* scratch;     // per-CTA epilogue scratch in smem (Epi::kScratchBytes)
};

// ------------------------------------------------------------------ the kernel
// A_MN / B_MN: the operand's MN dimension is the contiguous one in global memory (its tensor map has inner extent =
// MN size, rows = K extent, box 64 x 64); no transposed copy of the operand is needed.
template <int BN, int EPI_WARPS, class Epi, bool A_MN = false, bool B_MN = false, bool GROUPED = false>
__global__ void __launch_bounds__((PRODUCER_WARPS + EPI_WARPS) * 32, 1)
    tc_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int tiles_m,
                   int tiles_n, int batch, int k_blocks, const __grid_constant__ typename Epi::Params ep, GroupedK gk) {
  static_assert(BN % 32 == 0 && BN >= 32 && BN <= 256, "BN");
  static_assert(!GROUPED || (A_MN && B_MN), "the grouped contraction reads both operands MN-major");
  static_assert(!B_MN || BN % 64 == 0, "MN-major B needs BN to be a multiple of 64");
  static_assert(EPI_WARPS == 4 || EPI_WARPS == 8, "EPI_WARPS");
  constexpr int STAGES = num_stages(BN, Epi::kScratchBytes);
  constexpr int A_BYTES = BM * BK * 2;
  constexpr int STAGE_BYTES = stage_bytes(BN);
  constexpr uint32_t TMEM_COLS = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;

  extern __shared__ uint8_t smem_raw[];
  // offset arithmetic on the __shared__ symbol (not a uintptr_t round trip) keeps the address space visible to the
  // compiler, so epilogue scratch accesses compile to LDS / STS instead of generic LD / ST
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* ring = smem;
  uint8_t* scratch = ring + STAGES * STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(scratch + Epi::kScratchBytes);
  uint64_t* full_bar = bars;                  // [STAGES]
  uint64_t* empty_bar = bars + STAGES;        // [STAGES]
  uint64_t* tmem_full = bars + 2 * STAGES;    // [2]
  uint64_t* tmem_empty = bars + 2 * STAGES + 2;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total_tiles = tiles_m * tiles_n * batch;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], EPI_WARPS);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const int b = t / (tiles_m * tiles_n);
        const int r = t - b * (tiles_m * tiles_n);
        const int m0 = (r / tiles_n) * BM;
        const int n0 = (r % tiles_n) * BN;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = ring + stage * STAGE_BYTES;
          uint8_t* sb = sa + A_BYTES;
          mbar_expect_tx(&full_bar[stage], STAGE_BYTES);
          if (GROUPED) {
            const int per_panel = gk.sets_per_group * gk.row_blocks;
            const int pnl = kb / per_panel, rem = kb - pnl * per_panel;
            const int set = b * gk.sets_per_group + rem / gk.row_blocks, row = (rem % gk.row_blocks) * BK;
#pragma unroll
            for (int c = 0; c < BM / 64; ++c)
              tma_load_3d(sa + c * 8192, &tmA, &full_bar[stage], gk.a_off[pnl] + m0 + c * 64, row, set);
#pragma unroll
            for (int c = 0; c < BN / 64; ++c)
              tma_load_3d(sb + c * 8192, &tmB, &full_bar[stage], gk.b_off[pnl] + n0 + c * 64, row, set);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
            continue;
          }
          if (A_MN) {
#pragma unroll
            for (int c = 0; c < BM / 64; ++c) tma_load_3d(sa + c * 8192, &tmA, &full_bar[stage], m0 + c * 64, kb * BK, b);
          } else {
            tma_load_3d(sa, &tmA, &full_bar[stage], kb * BK, m0, b);
          }
          if (B_MN) {
#pragma unroll
            for (int c = 0; c < BN / 64; ++c) tma_load_3d(sb + c * 8192, &tmB, &full_bar[stage], n0 + c * 64, kb * BK, b);
          } else {
            tma_load_3d(sb, &tmB, &full_bar[stage], kb * BK, n0, b);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(BM, BN, A_MN, B_MN);
      // per UMMA_K step: K-major operands advance 32 bytes inside the 128-byte swizzle atom, MN-major operands advance
      // 16 k rows = 2048 bytes (descriptor addresses are in 16-byte units)
      constexpr uint64_t STEP_A = A_MN ? (2048 >> 4) : 2, STEP_B = B_MN ? (2048 >> 4) : 2;
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BN;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(ring + stage * STAGE_BYTES);
          const uint64_t da = A_MN ? make_smem_desc_mn128(sa) : make_smem_desc_k128(sa);
          const uint64_t db = B_MN ? make_smem_desc_mn128(sa + A_BYTES) : make_smem_desc_k128(sa + A_BYTES);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            umma_bf16(tmem_d, da + STEP_A * k, db + STEP_B * k, idesc, (kb > 0 || k > 0) ? 1u : 0u);
          }
          tc_commit(&empty_bar[stage]);     // frees the smem slot once these MMAs retire
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        tc_commit(&tmem_full[acc]);         // accumulator complete -> epilogue
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue warps =====================
    const int ew = warp - PRODUCER_WARPS;
    const int quad = warp & 3;                       // TMEM lane quadrant this warp may access
    constexpr int PARTS = EPI_WARPS / 4;             // warps sharing a quadrant split the columns
    const int part = (PARTS == 1) ? 0 : (ew >> 2);   // ew 0..3 -> part 0, 4..7 -> part 1 (quad = warp & 3 differs per ew)
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      const int b = t / (tiles_m * tiles_n);
      const int r = t - b * (tiles_m * tiles_n);
      EpiCtx cx;
      cx.b = b;
      cx.m0 = (r / tiles_n) * BM;
      cx.n0 = (r % tiles_n) * BN;
      cx.row = quad * 32 + lane;
      cx.col_begin = part * (BN / PARTS);
      cx.col_end = cx.col_begin + BN / PARTS;
      cx.lane = lane;
      cx.epi_warp = ew;
      cx.scratch = scratch;
      cx.tmem = tmem_base + acc * BN + (static_cast<uint32_t>(quad * 32) << 16);
      typename Epi::Pre pre;
      Epi::pre(ep, cx, pre);            // global loads that do not depend on the accumulator start here
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      Epi::run(ep, cx, pre);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}


// ------------------------------------------------------------------ host side
// 3-D tensor map over a batched K-major bf16 matrix: dims (K, rows, batch), box (64, box_rows, 1),
// 128-byte swizzle, out-of-bounds elements read as zero (this is what pads ragged N / K).
int make_tmap_bf16(CUtensorMap* out, const void* base, int64_t k_extent, int64_t rows, int64_t batch,
                   int64_t row_stride_elems, int64_t batch_stride_elems, int box_rows);
int make_tmap_bf16_k32(CUtensorMap* out, const void* base, int64_t k_extent, int64_t rows, int64_t batch,
                       int64_t row_stride_elems, int64_t batch_stride_elems, int box_rows);
// Tensor map for epilogue TMA STORES of 16-bit outputs: dims (cols, rows, batch), box (32 cols = 64 bytes, 32 rows, 1),
// 64-byte swizzle (the staging slab of a warp is [32 rows][64 B], 16-byte chunk index xor ((row >> 1) & 3)); elements
// outside the tensor are clipped, which is what handles ragged M / N.  fp16 = true for __half outputs, else bf16.
int make_tmap_store16(CUtensorMap* out, const void* base, int64_t cols, int64_t rows, int64_t batch,
                      int64_t row_stride_elems, int64_t batch_stride_elems, bool fp16);
// byte offset of 16-byte chunk `c16` (0..3) of row `r` (0..31) inside such a slab
__host__ __device__ constexpr int store_slab_offset(int r, int c16) { return r * 64 + ((c16 ^ ((r >> 1) & 3)) << 4); }
constexpr int kStoreSlabBytes = 32 * 64;          // one 32 x 32 chunk of 16-bit values; slabs are 1024-byte aligned

struct GemmShape {
  int M, N, K, batch;
};

// Tile width with the least wave-quantisation loss: a persistent launch runs ceil(tiles / SMs) rounds whose cost is
// proportional to the tile width.  Candidates 256 / 192 / 128 (wider tiles reuse A better, so they win ties).
inline int pick_tile_n(const GemmShape& s, int ctas_per_tile = 1) {
  const int cand[3] = {256, 192, 128};
  int best = 256;
  double best_cost = 1e30;
  for (int bn : cand) {
    const long long tiles = 1LL * ceil_div(s.M, BM * ctas_per_tile) * ceil_div(s.N, bn) * s.batch;
    const double rounds = (double)ceil_div<long long>(tiles, num_sms() / ctas_per_tile);
    const double cost = rounds * bn * (1.0 + 0.02 * (256 - bn) / 64.0);
    if (cost < best_cost - 1e-9) {
      best_cost = cost;
      best = bn;
    }
  }
  return best;
}

// For an MN-major operand build its tensor map with make_tmap_bf16(&tm, base, MN extent, K extent, batch, row stride,
// batch stride, 64): the inner (contiguous) dimension is MN, the rows run along K.
template <int BN, int EPI_WARPS, class Epi, bool A_MN = false, bool B_MN = false, bool GROUPED = false>
int launch_gemm(const char* name, const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmShape& s,
                const typename Epi::Params& ep, cudaStream_t stream, int max_ctas = 0, const GroupedK& gk = GroupedK{}) {
  if (s.M <= 0 || s.N <= 0 || s.batch <= 0) return GD3_OK;
  GD3_REQUIRE(s.K > 0, "tc_gemm: K must be positive");
  GD3_REQUIRE(GROUPED == (gk.panels > 0), "tc_gemm: grouped-K description does not match the kernel variant");
  auto kern = tc_gemm_kernel<BN, EPI_WARPS, Epi, A_MN, B_MN, GROUPED>;
  constexpr int SMEM = smem_bytes(BN, Epi::kScratchBytes);
  static_assert(SMEM <= 227 * 1024, "tc_gemm shared memory budget");
  static SmemOptIn opt;   // per instantiation
  GD3_CHECK_CUDA(opt.ensure(kern, SMEM));
  const int tiles_m = ceil_div(s.M, BM), tiles_n = ceil_div(s.N, BN);
  const long long total = 1LL * tiles_m * tiles_n * s.batch;
  int grid = num_sms();
  if (max_ctas > 0 && max_ctas < grid) grid = max_ctas;
  if (total < grid) grid = static_cast<int>(total);
  {
    GD3_PROF(name, stream);
    const int k_blocks = GROUPED ? gk.panels * gk.sets_per_group * gk.row_blocks : ceil_div(s.K, BK);
    kern<<<grid, (PRODUCER_WARPS + EPI_WARPS) * 32, SMEM, stream>>>(tmA, tmB, tiles_m, tiles_n, s.batch, k_blocks, ep, gk);
  }
  GD3_CHECK_LAUNCH();
  return GD3_OK;
}



// ------------------------------------------------------------------ coalesced epilogue I/O
// After tcgen05.ld a thread owns one ROW of the tile (32 consecutive columns per chunk), so a direct global
// access touches 32 different rows per warp instruction.  These helpers bounce a 32 x 32 chunk through a padded
// per-warp shared-memory tile so that every global instruction covers whole contiguous row segments instead.
constexpr int kWarpTileFloats = 32 * 33;
constexpr int kWarpTileBytes = kWarpTileFloats * 4;
constexpr int kMaxEpiWarps = 8;

// thread `lane` holds row `lane` of the chunk in v[32]; writes rows [0, rows) x cols [0, cols) to dst (row stride ld)
template <class OutT>
__device__ __forceinline__ void warp_store_rows(float* t, const float (&v)[32], OutT* dst, int64_t ld, int rows,
                                                int cols, int lane) {
#pragma unroll
  for (int q = 0; q < 32; ++q) t[lane * 33 + q] = v[q];
  __syncwarp();
  if constexpr (sizeof(OutT) == 4) {
    // one row per instruction: 32 lanes x 4 B = one 128-byte line
    const bool c_ok = lane < cols;
#pragma unroll
    for (int r = 0; r < 32; ++r) {
      const float x = t[r * 33 + lane];
      if (r < rows && c_ok) dst[r * ld + lane] = x;
    }
  } else {
    // two rows per instruction: 16 lanes x (2 x bf16) = 64 contiguous bytes per row
    const int half = lane >> 4, c2 = 2 * (lane & 15);
    const bool pair_ok = (ld % 2 == 0) && ((reinterpret_cast<uintptr_t>(dst) & 3) == 0);
#pragma unroll
    for (int r = 0; r < 32; r += 2) {
      const int rr = r + half;
      const float x0 = t[rr * 33 + c2], x1 = t[rr * 33 + c2 + 1];
      const bool r_ok = rr < rows;
      if (pair_ok) {
        if (r_ok && c2 + 1 < cols) *reinterpret_cast<uint32_t*>(dst + rr * ld + c2) = pack_bf16x2(x0, x1);
        else if (r_ok && c2 < cols) dst[rr * ld + c2] = __float2bfloat16(x0);
      } else {
        if (r_ok && c2 < cols) dst[rr * ld + c2] = __float2bfloat16(x0);
        if (r_ok && c2 + 1 < cols) dst[rr * ld + c2 + 1] = __float2bfloat16(x1);
      }
    }
  }
  __syncwarp();
}

// bf16 output staged as packed pairs: a 32 x 17 word tile per warp (2176 B instead of 4224 B, which buys the KL
// gradient GEMM a fifth pipeline stage at BN = 192)
constexpr int kWarpTileWords16 = 32 * 17;
constexpr int kWarpTileBytes16 = kWarpTileWords16 * 4;
__device__ __forceinline__ void warp_store_rows_bf16(uint32_t* t, const float (&v)[32], __nv_bfloat16* dst, int64_t ld,
                                                     int rows, int cols, int lane) {
#pragma unroll
  for (int q = 0; q < 16; ++q) t[lane * 17 + q] = pack_bf16x2(v[2 * q], v[2 * q + 1]);
  __syncwarp();
  const int half = lane >> 4, l = lane & 15, c2 = 2 * l;
  const bool pair_ok = (ld % 2 == 0) && ((reinterpret_cast<uintptr_t>(dst) & 3) == 0);
#pragma unroll
  for (int r = 0; r < 32; r += 2) {
    const int rr = r + half;
    const uint32_t w = t[rr * 17 + l];
    const bool r_ok = rr < rows;
    __nv_bfloat16* o = dst + rr * ld + c2;
    if (pair_ok && c2 + 1 < cols) {
      if (r_ok) *reinterpret_cast<uint32_t*>(o) = w;
    } else {
      if (r_ok && c2 < cols) o[0] = __ushort_as_bfloat16((unsigned short)(w & 0xFFFFu));
      if (r_ok && c2 + 1 < cols) o[1] = __ushort_as_bfloat16((unsigned short)(w >> 16));
    }
  }
  __syncwarp();
}

// issue the coalesced loads of a 32 x 32 bf16 chunk (two rows per instruction); raw[i] holds row 2i + half
__device__ __forceinline__ void warp_load_rows_issue(const __nv_bfloat16* src, int64_t ld, int rows, int cols, int lane,
                                                     uint32_t (&raw)[16]) {
  const int half = lane >> 4, c2 = 2 * (lane & 15);
  const bool pair_ok = (ld % 2 == 0) && ((reinterpret_cast<uintptr_t>(src) & 3) == 0);
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int rr = 2 * i + half;
    uint32_t w = 0;
    if (rr < rows) {
      if (pair_ok && c2 + 1 < cols) {
        w = __ldg(reinterpret_cast<const uint32_t*>(src + rr * ld + c2));
      } else {
        const uint32_t lo = (c2 < cols) ? (uint32_t)__bfloat16_as_ushort(src[rr * ld + c2]) : 0u;
        const uint32_t hi = (c2 + 1 < cols) ? (uint32_t)__bfloat16_as_ushort(src[rr * ld + c2 + 1]) : 0u;
        w = lo | (hi << 16);
      }
    }
    raw[i] = w;
  }
}
// scatter the loaded chunk into the warp tile and read back this thread's row as floats
__device__ __forceinline__ void warp_load_rows_finish(float* t, const uint32_t (&raw)[16], int lane, float (&x)[32]) {
  const int half = lane >> 4, c2 = 2 * (lane & 15);
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int rr = 2 * i + half;
    t[rr * 33 + c2] = bf16_bits_to_float(raw[i] & 0xFFFFu);
    t[rr * 33 + c2 + 1] = bf16_bits_to_float(raw[i] >> 16);
  }
  __syncwarp();
#pragma unroll
  for (int q = 0; q < 32; ++q) x[q] = t[lane * 33 + q];
  __syncwarp();
}

// ------------------------------------------------------------------ a plain epilogue: store fp32
// C[b][m][n] = alpha * acc   (row-major, leading dimension ldc, batch stride in elements)
// Tensor map for TMA stores of fp32 outputs: dims (cols, rows, batch), box (32 cols = 128 bytes, 32 rows, 1), 128-byte
// swizzle (staging slab [32 rows][128 B], 16-byte chunk index xor (row & 7)).
int make_tmap_store32(CUtensorMap* out, const void* base, int64_t cols, int64_t rows, int64_t batch,
                      int64_t row_stride_elems, int64_t batch_stride_elems);
__host__ __device__ constexpr int store_slab32_offset(int r, int c16) { return r * 128 + ((c16 ^ (r & 7)) << 4); }
constexpr int kStoreSlab32Bytes = 32 * 128;

struct EpiStoreF32 {
  static constexpr int kScratchBytes = kMaxEpiWarps * kWarpTileBytes;
  static_assert(kWarpTileBytes >= kStoreSlab32Bytes, "the TMA slab reuses the transposition scratch");
  struct Params {
    float* C;
    int M, N;
    int64_t ldc, batch_stride;
    float alpha;
    const float* batch_scale;   // optional per-batch factor read from device memory (nullptr = 1)
    int batch_scale_stride0 = 0;   // 1: every batch entry uses batch_scale[0]
    int use_tma = 0;            // set by enable_tma_store(): rows are 16-byte aligned and the map below is valid
    alignas(64) CUtensorMap tm_out = {};
  };
  struct Pre {};
  __device__ static void pre(const Params&, const EpiCtx&, Pre&) {}
  __device__ static void run(const Params& p, const EpiCtx& cx, const Pre&) {
    const float alpha = p.batch_scale ? p.alpha * __ldg(p.batch_scale + (p.batch_scale_stride0 ? 0 : cx.b)) : p.alpha;
    const int m_warp = cx.m0 + (cx.row & ~31);            // first row of this warp's 32-row slab
    const int rows = p.M - m_warp;                         // valid rows in the slab (may be <= 0)
    if (p.use_tma) {
      // fp32 tile out through TMA: 8 conflict-free STS.128 per thread into a 128-byte-swizzled slab, one box store per chunk
      uint8_t* slab = cx.scratch + cx.epi_warp * kStoreSlab32Bytes;
      for (int c = cx.col_begin; c < cx.col_end; c += 32) {
        const int n = cx.n0 + c;
        if (n >= p.N) break;
        float v[32];
        tmem_ld32(cx.tmem + c, v);
        if (rows <= 0) continue;
        if (cx.lane == 0) tma_store_wait_read();
        __syncwarp();
#pragma unroll
        for (int q = 0; q < 8; ++q)
          *reinterpret_cast<float4*>(slab + store_slab32_offset(cx.lane, q)) =
              make_float4(v[4 * q] * alpha, v[4 * q + 1] * alpha, v[4 * q + 2] * alpha, v[4 * q + 3] * alpha);
        fence_proxy_async_smem();
        __syncwarp();
        if (cx.lane == 0) {
          tma_store_3d(&p.tm_out, slab, n, m_warp, cx.b);
          tma_store_commit();
        }
      }
      if (cx.lane == 0) tma_store_wait_read();
      __syncwarp();
      return;
    }
    float* t = reinterpret_cast<float*>(cx.scratch) + cx.epi_warp * kWarpTileFloats;
    float* cslab = p.C + cx.b * p.batch_stride + static_cast<int64_t>(m_warp) * p.ldc;
    for (int c = cx.col_begin; c < cx.col_end; c += 32) {
      const int n = cx.n0 + c;
      if (n >= p.N) break;
      float v[32];
      tmem_ld32(cx.tmem + c, v);       // warp-collective: every lane participates
      if (rows <= 0) continue;
#pragma unroll
      for (int q = 0; q < 32; ++q) v[q] *= alpha;
      warp_store_rows<float>(t, v, cslab + n, p.ldc, rows, p.N - n, cx.lane);
    }
  }
};

// Switch an EpiStoreF32 output to TMA stores when its layout allows it (16-byte aligned base and strides); otherwise the
// parameters are left as they are and the shared-memory transposition path runs.  Returns GD3_OK or an error code.
inline int enable_tma_store(EpiStoreF32::Params& p, int batch) {
  const bool ok = (reinterpret_cast<uintptr_t>(p.C) % 16 == 0) && p.ldc % 4 == 0 && (batch <= 1 || p.batch_stride % 4 == 0) &&
                  p.N >= 1 && p.M >= 1;
  if (!ok) return GD3_OK;
  const int rc = make_tmap_store32(&p.tm_out, p.C, p.N, p.M, batch, p.ldc, batch > 1 ? p.batch_stride : p.ldc * p.M);
  if (rc) return rc;
  p.use_tma = 1;
  return GD3_OK;
}

}  // namespace tc
}  // namespace gd3
