/* lib3dgd -- C ABI of the B200-native geometric-distillation hot path.
 *
 * This is the drop-in boundary for the hot path of kaist-cvml/3d-vlm-gd:
 * every entry point below replaces one reference interface, cited as file:line relative to the
 * reference checkout.  The reference is Python, so the binding a maintainer adds is a ctypes stub
 * (INTEGRATION.md); `3d-vlm-gd_b200/gd3/_lib.py` is exactly that stub.
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / C++ types.  All data pointers are DEVICE pointers
 *     unless the parameter name ends in `_host`.
 *   - the caller owns every buffer including `workspace` (size from the matching *_workspace()
 *     function, 256-byte aligned); the library allocates nothing.
 *   - work is enqueued on `stream` (a cudaStream_t passed as void*); no internal streams and no
 *     host synchronisation.
 *   - return value: 0 on success, negative error code otherwise (GD3_ERR_*), message available
 *     from gd3_last_error() (thread-local).
 *   - there is no CPU fallback: without a CUDA device every compute entry point fails.
 */
#ifndef GD3_H_
#define GD3_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GD3_VERSION 100

/* error codes */
#define GD3_OK 0
#define GD3_ERR_INVALID (-1)
#define GD3_ERR_CUDA (-2)
#define GD3_ERR_WORKSPACE (-3)
#define GD3_ERR_UNSUPPORTED (-4)

/* element types of feature tensors */
#define GD3_DTYPE_F32 0
#define GD3_DTYPE_BF16 1

/* distances of the reciprocal-NN matcher (mast3r/fast_nn.py:26-37) */
#define GD3_DIST_DOT 0
#define GD3_DIST_L2 1

/* loss variants: which fine-tune script's body is reproduced */
#define GD3_VARIANT_MAST3R 0 /* src/finetune_timm_mast3r.py */
#define GD3_VARIANT_VGGT 1   /* src/finetune_timm_vggt.py   */
#define GD3_VARIANT_ME 2     /* src/finetune_timm_me.py     */

int gd3_version(void);
const char* gd3_last_error(void);

/* ------------------------------------------------------------------------------------------
 * Reciprocal nearest neighbours.  Replaces bruteforce_reciprocal_nns (mast3r/fast_nn.py:16-70):
 * nn_A[i] = argbest_j score(A_i, B_j), nn_B[j] = argbest_i score(A_i, B_j), lowest index on ties.
 * A: (nA, dim) fp32 row-major, B: (nB, dim) fp32 row-major, outputs int64 (either may be NULL to
 * skip that direction -- cdistMatcher.query, :78-84, only uses nn_A).  The reference's block_size
 * argument only bounds its temporary and has no effect on the result, so it does not appear here.
 * ------------------------------------------------------------------------------------------ */
size_t gd3_reciprocal_nn_workspace(int64_t nA, int64_t nB);
int gd3_reciprocal_nn(const float* A, int64_t nA, const float* B, int64_t nB, int64_t dim, int dist,
                      int64_t* nn_A, int64_t* nn_B, void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * Debug / self-test: C[b] = A[b] * B[b]^T through the tcgen05 GEMM used by all fused losses.
 * A: (batch, M, lda) bf16, B: (batch, N, ldb) bf16, C: (batch, M, ldc) fp32.  Not a reference
 * interface; used by tests to validate the TMA / UMMA descriptor plumbing in isolation.
 * ------------------------------------------------------------------------------------------ */
int gd3_debug_gemm_bf16(const void* A, const void* B, float* C, int64_t M, int64_t N, int64_t K, int64_t batch,
                        int64_t lda, int64_t ldb, int64_t ldc, int tile_n, void* stream);

#ifdef __cplusplus
}
#endif

#endif /* GD3_H_ */
