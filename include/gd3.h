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
#define GD3_DTYPE_F16 2 /* packed teacher volumes only (gd3_teacher_pack) */

/* distances of the reciprocal-NN matcher (mast3r/fast_nn.py:26-37) */
#define GD3_DIST_DOT 0
#define GD3_DIST_L2 1

/* loss variants: which fine-tune script's body is reproduced */
#define GD3_VARIANT_MAST3R 0 /* src/finetune_timm_mast3r.py */
#define GD3_VARIANT_VGGT 1   /* src/finetune_timm_vggt.py   */
#define GD3_VARIANT_ME 2     /* src/finetune_timm_me.py, one mean per pair */
#define GD3_VARIANT_ME_JOINT 3 /* src/finetune_timm_me.py:202-217 with B > 1: ONE mean over the positives of all pairs (Smooth-AP only) */

int gd3_version(void);
const char* gd3_last_error(void);

/* Instrumentation (not reference interfaces): number of kernels this library has launched in the
 * process, and optional per-kernel CUDA-event timing on the launching stream.  gd3_profile_read
 * synchronises the device and writes "<kernel> <launches> <total_ms>\n" lines, then resets. */
long long gd3_launch_count(void);
void gd3_profile_enable(int on);
size_t gd3_profile_read(char* buf, size_t buf_bytes);

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
 * Device-resident ping-pong of fast_reciprocal_NNs (mast3r/fast_nn.py:109-188, the loop at :147-170
 * for ret_basin=False, pixel_tol=0): starting from `seeds` (unique, sorted flat indices into pts1,
 * :130-131), alternately replace every live seed's partner by its nearest neighbour in the other
 * image (1 -> 2, then 2 -> 1), retiring a seed as soon as a partner repeats (:157,164), for at
 * most max_iter rounds (10 for grid seeds, 1 for explicit seeds, :121-128).  No host
 * synchronisation inside; the reference copies indices to the host after every 8192-block.
 *   pts1 (n1, dim), pts2 (n2, dim) fp32 descriptors; dist as for gd3_reciprocal_nn
 *   xy1, xy2  (n_seeds) int32 out: final flat indices (:182 keeps the converged ones)
 *   converged (n_seeds) uint8 out: ~notyet of the reference
 * host_poll = 0: never synchronises (all max_iter rounds are enqueued; rounds without live seeds are
 *   cheap no-ops), usable under stream capture.  host_poll = 1: from the second round on, reads the
 *   number of live seeds back after each round (4 bytes + a stream synchronisation) and stops early,
 *   like the reference's `while notyet.any()`.
 * The unique + sort of merge_corres (:87-106) stays with the caller.
 * ------------------------------------------------------------------------------------------ */
size_t gd3_fast_reciprocal_nn_workspace(int64_t n_seeds, int64_t dim);
int gd3_fast_reciprocal_nn(const float* pts1, int64_t n1, const float* pts2, int64_t n2, int64_t dim, int dist,
                           const int32_t* seeds, int64_t n_seeds, int max_iter, int host_poll, int32_t* xy1,
                           int32_t* xy2, uint8_t* converged, void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * Dense cost-volume KL loss, forward + backward, batched over P image pairs.
 * Replaces the body of calculate_cost_loss (src/finetune_timm_mast3r.py:504-540 for
 * GD3_VARIANT_MAST3R, src/finetune_timm_vggt.py:488-533 for GD3_VARIANT_VGGT) including
 * F.normalize + bmm (:524-528), get_masked_patch_cost (utils/functions.py:402-422) and
 * kl_divergence_map (utils/losses.py:5-15).
 *   f1, f2   (P, N, C) student patch features, dtype f32 or bf16, arbitrary element strides
 *            (sP, sN, sC) -- the MASt3R path hands over a channel-major view (sN = 1, sC = N)
 *   t12, t21 (P, N, N) teacher volumes, rows of t12 = patches of view 1, rows of t21 = patches of view 2;
 *            t_row_stride / t_pair_stride in elements.  teacher_dtype GD3_DTYPE_F32 (the reference's
 *            tensors, teacher_scale 1) or GD3_DTYPE_F16: the packed form of gd3_teacher_pack, values
 *            times teacher_scale -- half the bytes of the largest input of the step.
 *   tstats12/21 (P, 3, N) fp32 row statistics [sum_j c | sum_j t~ | sum_j t~ ln t~] from
 *            gd3_teacher_pack, or NULL, NULL: computed here with one more pass over the volumes
 *   m1, m2   (P, N) uint8 patch masks (mask_patch_1 of each direction; mask_patch_2 is always
 *            None in the reference's callers)
 *   loss     (P) fp32, one value per pair = (KL_12 + KL_21) / 2
 *   grad_f1/2 (P, N, C) contiguous, same dtype as the features: d loss[p] / d f (NULL, NULL =
 *            forward only)
 * grad_scale: factor folded into the gradients (e.g. loss weight / P); loss values are unscaled.
 * pairs_per_group: how many pairs share one pass over the workspace (0 = choose so that a group's
 * working set stays L2-resident).
 * ------------------------------------------------------------------------------------------ */
int64_t gd3_cost_kl_group_size(int64_t P, int64_t N, int64_t C, int64_t pairs_per_group);
size_t gd3_cost_kl_workspace(int64_t P, int64_t N, int64_t C, int64_t pairs_per_group, int with_backward);
int gd3_cost_kl(const void* f1, const void* f2, int dtype, int64_t P, int64_t N, int64_t C, int64_t s1P, int64_t s1N,
                int64_t s1C, int64_t s2P, int64_t s2N, int64_t s2C, const void* t12, const void* t21,
                int teacher_dtype, float teacher_scale, const float* tstats12, const float* tstats21,
                int64_t t_pair_stride, int64_t t_row_stride, const uint8_t* m1, const uint8_t* m2, int variant,
                float eps, float grad_scale, float* loss, void* grad_f1, void* grad_f2, int64_t pairs_per_group,
                void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * Teacher-side packing of a cost volume for gd3_cost_kl: the last step of a teacher producer
 * (dust3r/dust3r/model.py:346-366 -> res2['tgt_attn_map'], vggt/models/aggregator.py:259-273 followed by
 * src/finetune_timm_vggt.py:390-392).  Every row is read once: out = fp16(c * scale) and the statistics
 * get_masked_patch_cost / kl_divergence_map derive from the row (utils/functions.py:419-420,
 * utils/losses.py:6-9): R = sum_j c, T = sum_j t~, A = sum_j t~ ln t~ with t~ = max(c / max(R, eps), eps).
 *   t      (P, N, N) fp32, pair / row strides in elements
 *   out    (P, N, out_row_stride) fp16, out_row_stride >= N elements (a multiple of 8 keeps rows 16-byte aligned for
 *          gd3_cost_kl's vector loads when N is ragged, e.g. 37^2; the padding is zero-filled)
 *   stats  (P, 3, N) fp32 [R | T | A]
 * scale: a power of two (1024 keeps probabilities down to 6e-8 normal fp16 numbers).
 * ------------------------------------------------------------------------------------------ */
int gd3_teacher_pack(const float* t, int64_t P, int64_t N, int64_t t_pair_stride, int64_t t_row_stride, float eps,
                     float scale, void* out_f16, int64_t out_row_stride, float* stats, void* stream);

/* ------------------------------------------------------------------------------------------
 * Smooth-AP sparse-correspondence loss, forward + backward, batched over P pairs.
 * Replaces the loss bodies of calculate_matching_loss (src/finetune_timm_mast3r.py:557-589,
 * src/finetune_timm_vggt.py:543-574) and of src/finetune_timm_me.py:196-217, including torch.cdist,
 * torch.bmm and the clamped sigmoid of utils/functions.py:24-33.
 *   d1, d2      (P, K, C) fp32 contiguous, L2-normalised keypoint descriptors
 *   pts3d_1/2   (P, K, 3) fp32 3-D points of the keypoints
 *   temp 0.01, thr_neg 0.1 (thres3d_neg), thr_pos 5e-3 (thresh3d_pos, GD3_VARIANT_ME only)
 *   loss        (P) fp32;  grad_d1/2 (P, K, C) fp32 contiguous (times grad_scale) or NULL, NULL = forward only
 *   GD3_VARIANT_ME: loss[p] = mean over pair p's positives (NaN without positives, like a B = 1 reference call);
 *   GD3_VARIANT_ME_JOINT: loss[p] = pair p's share of the single batch-wide mean (sum_p loss[p] = reference loss)
 * ------------------------------------------------------------------------------------------ */
size_t gd3_smooth_ap_workspace(int64_t P, int64_t K, int64_t C, int with_backward);
int gd3_smooth_ap(const float* d1, const float* d2, const float* pts3d_1, const float* pts3d_2, int64_t P, int64_t K,
                  int64_t C, int variant, float temp, float thr_neg, float thr_pos, float grad_scale, float* loss,
                  float* grad_d1, float* grad_d2, void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * InfoNCE softmax-CE correspondence loss (upstream MASt3R criterion, mast3r/losses.py:237-272 with
 * get_similarities :202-209 and the 'mean' reduction of MatchingCriterion.forward :217-231).  No src/ script of
 * the reference calls it; it is provided because the task statement words the correspondence loss as
 * "softmax-CE" (SURVEY.md section 8, row a3b).  Shares the split-bf16 similarity GEMM with gd3_smooth_ap.
 *   d1, d2  (P, K, C) fp32 contiguous descriptors; valid (P, K) uint8 or NULL (all valid)
 *   mode    0 'all', 1 'proper', 2 'dual';  temperature 0.07, eps 1e-8 in the reference
 *   loss_mean (1): mean over all valid rows of the batch;  row_loss (P, K): per-row values (0 where invalid)
 *   grad_d1/2 (P, K, C) fp32: gradient of grad_scale * loss_mean (both NULL = forward only)
 * ------------------------------------------------------------------------------------------ */
size_t gd3_infonce_workspace(int64_t P, int64_t K, int64_t C, int with_backward);
int gd3_infonce(const float* d1, const float* d2, const uint8_t* valid, int64_t P, int64_t K, int64_t C, int mode,
                float temperature, float eps, float grad_scale, float* loss_mean, float* row_loss, float* grad_d1,
                float* grad_d2, void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * Relative-depth losses on the depth-difference head, forward + backward, batched over S keypoint
 * sets (one set = the keypoints of one image).  Replaces pairwise_logistic_ranking_loss
 * (utils/losses.py:18-41, mode 0), intra_depth_loss (utils/losses.py:44-69, mode 1), the head
 * DepthAwareFeatureFusion.fusion_layer (+ tanh) (utils/model.py:100-105,122-127) applied to all K^2
 * feature differences, and the cross-view L1 term of calculate_depth_loss
 * (src/finetune_timm_mast3r.py:489-494).
 *   feats   (S, K, D) fp32 contiguous keypoint features;  depths (S, K) fp32
 *   head    W1 (hidden, D), b1, gamma, beta, w2 (hidden), b2 (1), hidden = 128; use_tanh; ln_eps
 *   thr     depth threshold, >= 0 (0.05 in the callers; a pair of equal depths is never valid);  margin: hinge base margin (mode 1)
 *   joint_mean != 0: one mean over the valid pairs of all sets (the reference's B > 1 semantics)
 *   w_rank  (S) weights with which each set's loss enters the differentiated total (NULL = 1)
 *   w_l1    (S/2) weights of the L1 term coupling set 2p (view 1) with set 2p+1 (view 2); NULL = no L1
 *   loss_rank (S), loss_l1 (S/2): unweighted per-set / per-pair losses
 *   grad_feats (S, K, D) and grad_params, packed [W1 (hidden*D) | b1 | gamma | beta | w2 | b2]:
 *           gradient of  sum_s w_rank[s] loss_rank[s] + sum_p w_l1[p] loss_l1[p]   (both NULL = forward only)
 * ------------------------------------------------------------------------------------------ */
size_t gd3_depth_head_loss_workspace(int64_t S, int64_t K, int64_t D, int with_backward, int with_l1);
int gd3_depth_head_loss(const float* feats, const float* depths, int64_t S, int64_t K, int64_t D, int64_t hidden,
                        const float* W1, const float* b1, const float* gamma, const float* beta, const float* w2,
                        const float* b2, int use_tanh, float ln_eps, int mode, float thr, float margin,
                        int joint_mean, const float* w_rank, const float* w_l1, float* loss_rank, float* loss_l1,
                        float* grad_feats, float* grad_params, void* workspace, size_t workspace_bytes,
                        void* stream);

/* ------------------------------------------------------------------------------------------
 * Bilinear sampling of patch-token maps at pixel keypoints.  Replaces interpolate_features
 * (utils/functions.py:55-76) and the glue of get_intermediate_feature / get_feature
 * (src/finetune_timm_mast3r.py:271-277, 307-313): optional mean over L layers (sampling is linear)
 * and optional channel L2 normalisation (F.normalize, eps 1e-12).
 *   tokens  element (l, p, n, c) at tokens[l*sL + p*sP + n*sN + c*sC], n = y*pw + x; f32 or bf16
 *   ph, pw  feature-map size in patches; h, w the image size in pixels the reference passes
 *   kp      (P, K, 2) fp32 pixel (x, y)
 *   out     element (p, k, c) at out[p*oP + k*oK + c*oC], fp32
 *   inv_norm (P, K) fp32, written when normalize != 0 (needed by the backward)
 * The backward accumulates into grad_tokens (fp32, same strides as tokens, caller-zeroed).  grad_extra
 * (optional, element (p, k, c) at p*eP + k*eK + c*eC) is a second gradient w.r.t. the UN-normalised sample of
 * the same keypoints; it is added after the normalisation backward so both share one scatter.
 * ------------------------------------------------------------------------------------------ */
int gd3_sample_tokens_fwd(const void* tokens, int dtype, int64_t L, int64_t P, int64_t C, int64_t ph, int64_t pw,
                          int64_t h, int64_t w, int64_t sL, int64_t sP, int64_t sN, int64_t sC, const float* kp,
                          int64_t K, int patch, int stride, int normalize, float* out, int64_t oP, int64_t oK, int64_t oC,
                          float* inv_norm, void* stream);
int gd3_sample_tokens_bwd(const float* grad_out, int64_t gP, int64_t gK, int64_t gC, const float* out, int64_t oP,
                          int64_t oK, int64_t oC, const float* inv_norm, const float* kp, int64_t L, int64_t P,
                          int64_t K, int64_t C, int64_t ph, int64_t pw, int64_t h, int64_t w, int patch, int stride,
                          int normalize, float* grad_tokens, int64_t sL, int64_t sP, int64_t sN, int64_t sC,
                          const float* grad_extra, int64_t eP, int64_t eK, int64_t eC, void* stream);

/* ------------------------------------------------------------------------------------------
 * Keypoint -> best-matching pixel of the other image.  Replaces the inline block
 * src/evaluate_timm.py:532-547: F.interpolate(img2_desc, (ds, ds), bilinear, align_corners=True)
 * with ds = (img - patch) // stride * stride + 1, VF.pad(..., padding_mode='edge') to img x img,
 * einsum('nfk,nif->nki') against the K keypoint descriptors and argmax over the img^2 pixels --
 * without materialising the upsampled (1, C, img, img) map (the interpolation is applied to the
 * K x (ph pw) low-resolution similarity instead; it is linear).
 *   kp_desc  element (k, c) at kp_desc[k * kd_stride_k + c * kd_stride_c] fp32 (the reference's
 *            interpolate_features output (1, C, K): stride_k = 1, stride_c = K)
 *   desc2    (C, ph, pw) fp32 contiguous patch descriptors of image 2
 *   nn_idx   (K) int64: y * img + x of the arg-max pixel, lowest index on ties (:543-545)
 *   nn_val   (K) fp32 or NULL: the similarity attained
 * ------------------------------------------------------------------------------------------ */
size_t gd3_semantic_argmax_workspace(int64_t K, int64_t ph, int64_t pw);
int gd3_semantic_argmax(const float* kp_desc, int64_t kd_stride_k, int64_t kd_stride_c, const float* desc2, int64_t K,
                        int64_t C, int64_t ph, int64_t pw, int64_t img_size, int64_t patch_size, int64_t stride,
                        int64_t* nn_idx, float* nn_val, void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * Keypoint-side inputs of the losses, batched over P pairs (one launch instead of a dozen torch
 * ops per pair).  Replaces get_patch_mask_from_kp_tensor (utils/functions.py:375-399) and
 * extract_kp_depth (utils/functions.py:348-372).
 *   kp        (P, K, 2) fp32 pixel keypoints (x, y)
 *   mask      (P, (H / patch) * (W / patch)) uint8 out, or NULL: patches containing >= 1 in-image keypoint
 *   depth     fp32 depth maps (H, W) per pair (depth_pair_stride elements apart; 0 = one shared map)
 *   kp_depth  (P, K) fp32 out, or NULL: mean of the replicate-padded window x window neighbourhood
 *             at the truncated keypoint
 * ------------------------------------------------------------------------------------------ */
int gd3_kp_prepare(const float* kp, int64_t P, int64_t K, int64_t H, int64_t W, int patch, int window, const float* depth,
                   int64_t depth_pair_stride, uint8_t* mask, float* kp_depth, void* stream);

/* ------------------------------------------------------------------------------------------
 * Volume-level helpers of the cost-volume loss for callers that already hold (B, hw, hw2) fp32
 * volumes (the minimal drop-in of INTEGRATION.md section 2; gd3_cost_kl never builds a volume).
 *
 * gd3_masked_patch_cost replaces get_masked_patch_cost (utils/functions.py:402-422):
 *   xm = cost with the rows of mask1 == 0 (and, when mask2 is given, the columns of mask2 == 0)
 *   overwritten by 0;  use_softmax == 0: out = xm / clamp_min(sum_j xm, eps);
 *   use_softmax != 0: out = softmax(xm / temperature) in fp32 (a masked row becomes uniform).
 *   mask1 (hw) / mask2 (hw2 or NULL): bool bytes shared by the B volumes.  row_sum (B * hw) out:
 *   the row's sum (or sum of exponentials), which the backward needs; may be NULL without one.
 * gd3_masked_patch_cost_backward: grad_cost from grad_out and the saved out / row_sum
 *   (softmax: y (dy - <dy, y>) / T;  row-normalisation: (dy - [sum >= eps] <dy, y>) / max(sum, eps);
 *   0 at the overwritten entries).
 *
 * gd3_kl_divergence_map replaces kl_divergence_map (utils/losses.py:5-15) on (rows, n) volumes:
 *   loss[0] = mean over the rows of sum_j t~ log(t~ / s~), t~ = clamp_min(teacher, eps), s~ likewise;
 *   grad_student / grad_teacher (rows, n; either may be NULL) = d loss / d student, d loss / d teacher
 *   from the same pass: -(t~ / s~) [s >= eps] / rows and (log(t~ / s~) + 1) [t >= eps] / rows.
 *   The row sums go through the workspace and are added in a fixed order (deterministic).
 * ------------------------------------------------------------------------------------------ */
int gd3_masked_patch_cost(const float* cost, int64_t B, int64_t hw, int64_t hw2, const uint8_t* mask1, const uint8_t* mask2,
                          int use_softmax, float eps, float temperature, float* out, float* row_sum, void* stream);
int gd3_masked_patch_cost_backward(const float* grad_out, const float* out, const float* row_sum, int64_t B, int64_t hw,
                                   int64_t hw2, const uint8_t* mask1, const uint8_t* mask2, int use_softmax, float eps,
                                   float temperature, float* grad_cost, void* stream);
size_t gd3_kl_divergence_map_workspace(int64_t rows);
int gd3_kl_divergence_map(const float* teacher, const float* student, int64_t rows, int64_t n, float eps, float* loss,
                          float* grad_student, float* grad_teacher, void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * VGGT teacher cost volumes, one call per global-attention block (VGGT half).
 * Replaces the return_attn branch of vggt/layers/attention.py:73-84 (cross-view scores of the
 * patch tokens, softmax(scores / temperature) per head, both directions) together with what the
 * callers do with the maps: mean over the collected blocks (vggt/models/aggregator.py:259-273) and
 * mean over heads (src/finetune_timm_vggt.py:390-392).
 *   q_scaled, k  (B, heads, n_tokens, head_dim) bf16 contiguous; q already multiplied by the
 *                attention scale (the reference's "q = q * self.scale")
 *   n_tokens     both views concatenated; each view's first `skip` tokens (camera / register
 *                tokens, 5 in the reference) are dropped: n = n_tokens / 2 - skip <= 2048
 *   round_bf16   1: round the scores and scores / temperature to bf16 as the bf16-autocast
 *                teacher does (src/finetune_timm_vggt.py:359); 0: keep fp32
 *   attn12/21    (B, n, n) fp32: view-1 queries over view-2 keys / the reverse.
 *                accumulate = 0: attn = weight * head-mean; 1: attn += weight * head-mean
 *                (weight = 1 / number of collected blocks gives the reference's cost_1, cost_2)
 * ------------------------------------------------------------------------------------------ */
size_t gd3_vggt_attn_workspace(int64_t B, int64_t heads, int64_t n);
int gd3_vggt_attn_accumulate(const void* q_scaled, const void* k, int64_t B, int64_t heads, int64_t n_tokens,
                             int64_t head_dim, int64_t skip, float temperature, int round_bf16, float weight,
                             int accumulate, float* attn12, float* attn21, void* workspace, size_t workspace_bytes,
                             void* stream);

/* ------------------------------------------------------------------------------------------
 * Point map -> depth image, batched.  Replaces point_cloud_to_depth
 * (utils/functions.py:218-260; called per image at src/finetune_timm_mast3r.py:627-633): points
 * with z > 0 are projected with the pinhole intrinsics, rounded half-to-even to a pixel, dropped
 * when outside the image, and every pixel gets the mean z of the points that landed on it (0 if
 * none).
 *   points      (B, M, 3) fp32 camera-frame points
 *   intrinsics  (3, 3) fp32 row-major DEVICE matrices, intr_stride elements apart (0 = one shared)
 *   depth       (B, h, w) fp32 out
 * ------------------------------------------------------------------------------------------ */
size_t gd3_point_cloud_to_depth_workspace(int64_t B, int64_t h, int64_t w);
int gd3_point_cloud_to_depth(const float* points, int64_t B, int64_t M, const float* intrinsics, int64_t intr_stride,
                             int64_t w, int64_t h, float* depth, void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * MASt3R teacher cost-volume post-processing, fused.  Replaces
 * dust3r/dust3r/model.py:346-366: per decoder layer, head-mean of both branches' pre-softmax
 * cross-attention logits (dust3r/croco/models/blocks.py:163-164), symmetrisation with the
 * transposed other branch, softmax(./temperature), column 0 := global minimum of the layer
 * (:353-354), then the mean over layers (:363).
 *   mode  GD3_TV_RECIPROCAL  the above (self.reciprocity = True, the configuration the fine-tuning uses)
 *         GD3_TV_HEAD_MEAN   :355-359: head-mean, column-0 rule, layer mean (src_layers unused)
 *         GD3_TV_PLAIN_MEAN  mean over layers and heads only: the VGGT teacher's aggregation of its
 *                            per-block maps, vggt/models/aggregator.py:273 + src/finetune_timm_vggt.py:390-392
 *   tgt_layers, src_layers  HOST arrays of L DEVICE pointers, each (B, H, N, N) fp32 contiguous
 *   out                     (B, N, N) fp32 = tgt_attn_map
 * Every logit is read once; workspace holds the per-layer maps (L B N^2 fp32).
 * ------------------------------------------------------------------------------------------ */
#define GD3_TV_HEAD_MEAN 0
#define GD3_TV_RECIPROCAL 1
#define GD3_TV_PLAIN_MEAN 2
size_t gd3_teacher_volume_workspace(int64_t L, int64_t B, int64_t N);
int gd3_teacher_volume(const float* const* tgt_layers, const float* const* src_layers, int64_t L, int64_t B, int64_t H,
                       int64_t N, float temperature, int mode, float* out, void* workspace,
                       size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * Debug / self-test: C[b] = A[b] * B[b]^T through the tcgen05 GEMM used by all fused losses.
 * A: (batch, M, lda) bf16, B: (batch, N, ldb) bf16, C: (batch, M, ldc) fp32.  Not a reference
 * interface; used by tests to validate the TMA / UMMA descriptor plumbing in isolation.
 * ------------------------------------------------------------------------------------------ */
int gd3_debug_gemm_bf16(const void* A, const void* B, float* C, int64_t M, int64_t N, int64_t K, int64_t batch,
                        int64_t lda, int64_t ldb, int64_t ldc, int tile_n, void* stream);
/* The same with MN-major operands (no transposed copies): a_mn -> A is (batch, K, M) with M contiguous,
 * b_mn -> B is (batch, K, N) with N contiguous; C (batch, M, N) fp32 contiguous. */
int gd3_debug_gemm_bf16_mn(const void* A, const void* B, float* C, int64_t M, int64_t N, int64_t K, int64_t batch,
                           int a_mn, int b_mn, int tile_n, void* stream);

#ifdef __cplusplus
}
#endif

#endif /* GD3_H_ */
