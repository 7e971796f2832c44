/* libmanipose_sm100.so — C ABI of the B200-native (sm_100a) ManiPose lifting hot path.
 *
 * The reference (cedricrommel/manipose) is pure PyTorch with no FFI; the entry points below are what
 * a ctypes binding for its hot path binds.  Each one cites the reference
 * function (path:line under the reference checkout) whose work it replaces.
 *
 * Conventions
 *   - plain pointers + sizes, no torch types; every pointer is a DEVICE pointer unless said otherwise;
 *   - all tensors are contiguous, row-major, with the reference's dimension order;
 *   - every call is asynchronous on `stream` (a cudaStream_t), allocates nothing, never synchronises,
 *     and is capturable in a CUDA graph;
 *   - returns MP_OK (0) or a negative MP_E* code; mp_last_error() gives the message (thread-local);
 *   - no CPU fallback and no other-architecture fallback: a non-sm_100 device is MP_EDEVICE.
 */
#ifndef MANIPOSE_SM100_H_
#define MANIPOSE_SM100_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MP_ABI_VERSION 1

enum {
  MP_OK = 0,
  MP_EINVAL = -1,       /* bad shape / null pointer / bad flag (reference: assert / ValueError)     */
  MP_EUNSUPPORTED = -2, /* valid in the reference but not built here (e.g. another skeleton tree)  */
  MP_EDEVICE = -3,      /* current device is not sm_100                                            */
  MP_ELAUNCH = -4,      /* cudaGetLastError() after a launch                                       */
  MP_EALIGN = -5,       /* pointer not 16-byte aligned where the kernel needs it                   */
  MP_EWORKSPACE = -6    /* workspace too small                                                     */
};

typedef void* mp_stream_t; /* cudaStream_t */

int mp_abi_version(void);
const char* mp_last_error(void);
/* 0 if the current CUDA device is sm_100 (B200), MP_EDEVICE otherwise. */
int mp_device_check(void);
/* Persistent kernels size their grids from min(SMs of the device, `sms`); 0 = no limit (the default). Two callers that drive two
 * streams (two micro-batches in flight) give each the half of the device it can fill, so that a tensor-bound launch of one
 * runs beside an HBM-bound launch of the other. Process-wide, read at launch time. Returns the previous limit (>= 0). */
int mp_set_sm_limit(int sms);

/* Skeleton tables (hpe/mh_so3_hpe/data/skeleton.py:7-172, t_pose_operators from
 * hpe/mh_so3_hpe/data/h36m_lifting.py:40-57).  The kernels are specialised at compile time for the
 * 17-joint H36M / MPI-INF-3DHP tree (both datasets use it: h36m_lifting.py:649-660 ==
 * dataset_3dhp.py:132-138); this call VALIDATES the caller's tables against it and returns
 * MP_EUNSUPPORTED for any other tree.  parents[num_joints], operators[num_joints*3] (row 0 ignored). */
int mp_set_skeleton(int num_joints, const int32_t* host_parents, const float* host_t_pose_operators);

/* ---- manifold decoder -----------------------------------------------------
 * PoseDecoder.forward (hpe/mh_so3_hpe/architectures/pose_decoder.py:32-55) =
 * compute_rotation_matrix_from_ortho6d (utils/rotation_tools.py:35-57) + build_t_pose_from_bone_lengths
 * (pose_decoder.py:98-120) + forward_kinematics (utils/forward_kinematics.py:6-48), fused with the
 * softmax over hypotheses of RMCLRotMixSTE.forward (rmcl_manifold_mix_ste.py:262).
 *   rot6d   [n_clips*n_hyp*n_frames, 17, rot_rep_dim] fp32, rot_rep_dim 6 (Gram-Schmidt) or 4 (R_theta R_phi,
 *           rotation_tools.py:60-116)   (the reference's "(B H L) J D" flattening)
 *   bone_len[n_clips, 16] fp32 (signed)             pose n uses clip n / (n_hyp*n_frames)
 *   root    [n_poses, 3] or NULL (= zeros, what the reference passes)
 *   logits  [n_clips, n_hyp, n_frames] or NULL; scores (same shape) = softmax over n_hyp
 *   poses   [n_poses, 17, 3] fp32 */
#define MP_DEC_EXACT 0 /* one correctly-rounded IEEE fp32 operation per reference operation, same order:
                          bit-identical to oracle.pose_decoder_ieee, ~1e-7 from torch CPU (default)  */
#define MP_DEC_FAST 1  /* rsqrt + FMA contraction: <= 1e-6 relative difference, fewer instructions   */
int mp_decoder_fwd(const float* rot6d, const float* bone_len, const float* root, const float* logits,
                   float* poses, float* scores, int64_t n_clips, int64_t n_hyp, int64_t n_frames,
                   int rot_rep_dim, int flags, mp_stream_t stream);
/* Backward of the above w.r.t. rot6d, bone_len (summed over a clip's poses in a fixed order: run-to-run identical; grad_bone_len
 * [n_clips, 16] is overwritten) and root (grad_root may be NULL).  Recomputes rotations from rot6d.
 * workspace >= mp_decoder_bwd_workspace_bytes(n_clips, n_hyp, n_frames): the per-tile partial rows of the bone-length sums. */
size_t mp_decoder_bwd_workspace_bytes(int64_t n_clips, int64_t n_hyp, int64_t n_frames);
int mp_decoder_bwd(const float* rot6d, const float* bone_len, const float* grad_poses,
                   float* grad_rot6d, float* grad_bone_len, float* grad_root, int64_t n_clips,
                   int64_t n_hyp, int64_t n_frames, int rot_rep_dim, void* workspace,
                   size_t workspace_bytes, mp_stream_t stream);
/* softmax over n_hyp alone (scores_logits.softmax(dim=1), rmcl_manifold_mix_ste.py:262); [n_clips, n_hyp, n_frames]. */
int mp_softmax_hyp_fwd(const float* logits, float* scores, int64_t n_clips, int64_t n_hyp, int64_t n_frames,
                       mp_stream_t stream);
/* softmax over n_hyp backward: grad_logits = s * (g - sum_k s g); all [n_clips, n_hyp, n_frames]. */
int mp_softmax_hyp_bwd(const float* scores, const float* grad_scores, float* grad_logits,
                       int64_t n_clips, int64_t n_hyp, int64_t n_frames, mp_stream_t stream);

/* ---- losses and hypothesis metrics ---------------------------------
 * terms layout (fp32[MP_LOSS_NTERMS]) written by mp_loss_fwd: */
enum {
  MP_TERM_WTA = 0,    /* wta_l2_loss_and_activate_head(...)[0].mean()  (losses.py:126-138)          */
  MP_TERM_BCE = 1,    /* F.binary_cross_entropy(scores, one_hot(winner)) (losses.py:165-168)        */
  MP_TERM_VEL = 2,    /* mean_velocity_error(axis=2) (losses.py:75-101)                             */
  MP_TERM_SMOOTH = 3, /* smoothness_regularization(axis=2) (regularizations.py:160-174)             */
  MP_TERM_TOTAL = 4,  /* wta + beta*bce + vel_w*vel + smooth_w*smooth (main_h36m_lifting.py:101-209) */
  MP_LOSS_NTERMS = 8
};
/* _l2_loss_per_hyp + torch.min(dim=1) (losses.py:104-138): hyp [B,K,T,17,3], y [B,T,17,3],
 * joint_weights[17] or NULL (= ones) -> wta_val [B,T] fp32, wta_idx [B,T] int64 (lowest k on ties),
 * per_hyp [B,K,T] or NULL.  Same operation order as torch CPU (FMA 3-norm, 8-lane sum of 17, /17): values agree
 * to 1 ulp (torch's CPU sqrt is not correctly rounded), winner indices are identical on identical inputs. */
int mp_wta_fwd(const float* hyp, const float* y, const float* joint_weights, int squared,
               float* wta_val, int64_t* wta_idx, float* per_hyp, int64_t B, int64_t K, int64_t T,
               mp_stream_t stream);
/* All four training-loss terms in one pass (make_loss closures, hpe/main_h36m_lifting.py:129-169).
 * scores [B,K,T] (or [B,K,T,1]); workspace >= mp_loss_workspace_bytes(B,K,T). */
size_t mp_loss_workspace_bytes(int64_t B, int64_t K, int64_t T);
int mp_loss_fwd(const float* hyp, const float* scores, const float* y, const float* joint_weights,
                int squared, float beta, float vel_w, float smooth_w, float* terms, float* wta_val,
                int64_t* wta_idx, int64_t B, int64_t K, int64_t T, void* workspace,
                size_t workspace_bytes, mp_stream_t stream);
/* Reverse mode of mp_loss_fwd w.r.t. hyp and scores (wta_idx from mp_loss_fwd; beta/vel_w/smooth_w the same).
 *   grad_terms  [MP_LOSS_NTERMS] device fp32: upstream gradient w.r.t. terms[WTA, BCE, VEL, SMOOTH, TOTAL]
 *   grad_wta_val[B,T] or NULL: additional upstream gradient w.r.t. the per-frame wta_val output
 *   grad_scores [B,K,T] or NULL */
int mp_loss_bwd(const float* hyp, const float* scores, const float* y, const float* joint_weights,
                const int64_t* wta_idx, int squared, float beta, float vel_w, float smooth_w,
                const float* grad_terms, const float* grad_wta_val, float* grad_hyp, float* grad_scores,
                int64_t B, int64_t K, int64_t T, mp_stream_t stream);

/* RMCLManifoldMixSTE.aggregate (rmcl_manifold_mix_ste.py:141-185) + poses_from_hyp_idx (:121-139). */
enum { MP_AGG_WEIGHTED_AVE = 0, MP_AGG_BEST_SCORE = 1, MP_AGG_ORACLE = 2 };
/* hyp [B,K,T,17,3]; scores [B,K,T] (modes 0,1); y [B,T,17,3] (mode 2).  out_pose [B,T,17,3];
 * out_val [B,T] (mode 2: unweighted per-frame MPJPE of the winner) or NULL; out_idx [B,T] int64 or NULL. */
int mp_aggregate(const float* hyp, const float* scores, const float* y, int mode, float* out_pose,
                 float* out_val, int64_t* out_idx, int64_t B, int64_t K, int64_t T, mp_stream_t stream);
/* mpjpe_error (mean_joint_errors.py:31-36): sum over n_points of ||gt - pred||_2 -> out[0] (fp32 sum),
 * out[1] (mean).  workspace >= mp_mpjpe_workspace_bytes(n_points). */
size_t mp_mpjpe_workspace_bytes(int64_t n_points);
int mp_mpjpe(const float* pred, const float* gt, int64_t n_points, float* out, void* workspace,
             size_t workspace_bytes, mp_stream_t stream);

/* ---- MixSTE backbone building blocks ------------------------------------------------------
 * Token order is always [clip, frame, token] (one layout, no transposes: the reference's rearranges,
 * mix_ste.py:131,144,167,171,184, become strided reads inside the attention kernel).
 * The residual stream x is fp32; everything that feeds a tensor-core contraction (normalised activations, q/k/v,
 * attention output, MLP hidden, weights) is 16-bit with fp32 accumulation.  `dtype` selects the 16-bit format: */
enum { MP_DTYPE_BF16 = 0, /* BASELINE config 3: "bf16 backbone" */
       MP_DTYPE_FP16 = 1  /* same rate and bytes, 3 more mantissa bits (saturating conversions) */ };

/* nn.Linear on tcgen05 (call sites mix_ste.py:209-222 fc1/fc2, :257,:280 qkv/proj):
 *   MP_EPI_BIAS:     Y[M,N] (16-bit) = A[M,K] W[N,K]^T + bias[N]
 *   MP_EPI_GELU:     Y[M,N] (16-bit) = GELU_erf(A W^T + bias)                       (nn.GELU, mix_ste.py:200)
 *   MP_EPI_RESIDUAL: Y[M,N] (fp32)   = resid[M,N] (fp32) + A W^T + bias             (Block.forward, mix_ste.py:352-358)
 *   MP_EPI_BIAS_F32: Y[M,N] (fp32)   = A W^T + bias                                  (the folded K-head projection: fp32 out)
 *   MP_EPI_ACCUMULATE: Y[M,N] (fp32) += A W^T  (bias ignored, may be NULL): the contraction is split over the SMs and the partial
 *                    tiles are added with TMA reduce stores (fp32 adds in L2, order not fixed) — weight gradients, few tiles, long K
 * A, W 16-bit dense row-major (row strides K); bias fp32; K % 64 == 0, N % 128 == 0; Y may alias resid. */
enum { MP_EPI_BIAS = 0, MP_EPI_GELU = 1, MP_EPI_RESIDUAL = 2, MP_EPI_ACCUMULATE = 3, MP_EPI_BIAS_F32 = 4 };
int mp_linear(const void* A, const void* W, const float* bias, const float* resid, void* Y, int64_t M, int64_t N,
              int64_t K, int epilogue, int dtype, mp_stream_t stream);

/* fc1 of the training forward (Mlp.forward, mix_ste.py:209-222): U[M,N] (16-bit) = A W^T + bias — the pre-activation the GELU backward
 * needs — and G[M,N] (16-bit) = GELU_erf(A W^T + bias), both from one accumulator (the GELU is taken on the fp32 value, like
 * mp_linear(MP_EPI_GELU)).  N % 256 == 0, K % 64 == 0. */
int mp_linear_gelu2(const void* A, const void* W, const float* bias, void* U, void* G, int64_t M, int64_t N, int64_t K, int dtype,
                    mp_stream_t stream);

/* Residual Linear with the LayerNorms that follow it fused into the epilogue (N = 512 = one whole row per CTA pair):
 *   x = resid + s * (A W^T + bias)                 (Block.forward residual adds, mix_ste.py:352-358; s = row_scale[row], the per-sample
 *                                                   DropPath factor of training, mix_ste.py:334-336 — NULL: s = 1)
 *   if x_pre (fp32, needs post_gamma, own buffer): x_pre = x     (the pre-post-norm value: the training tape keeps both)
 *   if post_gamma: x = LN(x; post_*) (+ pos_embed[(row / pos_div) % pos_mod])   (Spatial_norm / Temporal_norm, :143,149,154,166,170)
 *   x_out (fp32, may alias resid) = x
 *   if ln_gamma: h_out (16-bit) = LN(x; ln_*)      (norm2 of this block / norm1 of the next, :353,356)
 * Same results as mp_linear(MP_EPI_RESIDUAL) followed by mp_layernorm, without the extra passes over the residual stream. */
int mp_linear_ln(const void* A, const void* W, const float* bias, const float* resid, float* x_out, void* h_out,
                 const float* post_gamma, const float* post_beta, float post_eps, const float* pos_embed, int64_t pos_div,
                 int64_t pos_mod, const float* ln_gamma, const float* ln_beta, float ln_eps, const float* row_scale, float* x_pre,
                 int64_t M, int64_t N, int64_t K, int dtype, mp_stream_t stream);

/* The whole MLP branch of a C = 512 block in one launch (Mlp.forward + residual add, mix_ste.py:216-222,356-358, + the LayerNorms
 * of mp_linear_ln): x = resid + fc2(GELU_erf(fc1(h_in))) with the 1024-wide hidden activation kept on chip (TMEM / shared memory),
 * then the same epilogue as mp_linear_ln.  h_in [M,512] 16-bit, W1 [1024,512], W2 [512,1024] 16-bit, b1 [1024], b2 [512] fp32.
 * Same results as mp_linear(MP_EPI_GELU) followed by mp_linear_ln (the hidden activation is rounded to 16 bits in both).
 * h_out may alias h_in (a tile's rows are resident on chip before any of them is written).  C = 512 and hidden = 1024 only. */
int mp_mlp_ln(const void* h_in, const void* W1, const float* b1, const void* W2, const float* b2, const float* resid, float* x_out,
              void* h_out, const float* post_gamma, const float* post_beta, float post_eps, const float* pos_embed, int64_t pos_div,
              int64_t pos_mod, const float* ln_gamma, const float* ln_beta, float ln_eps, int64_t M, int64_t C, int64_t hidden, int dtype,
              mp_stream_t stream);

/* LayerNorm family (fp32 statistics over C in {128, 512}; one warp per token).
 *   x_in  [n_tokens, C] fp32
 *   if post_gamma != NULL: x = LN(x_in; post_gamma, post_beta, post_eps)       (shared Spatial_norm /
 *        Temporal_norm, mix_ste.py:143,154,166,170) (+ pos_embed[(token / pos_div) % pos_mod, C] fp32 if
 *        pos_embed != NULL: the Temporal_pos_embed add of mix_ste.py:149); written to x_out (fp32, may alias x_in)
 *   if ln_gamma != NULL: h_out = LN(x; ln_gamma, ln_beta, ln_eps) (next block's norm1 / this block's norm2), 16-bit */
int mp_layernorm(const float* x_in, float* x_out, void* h_out, const float* post_gamma, const float* post_beta,
                 float post_eps, const float* pos_embed, int64_t pos_div, int64_t pos_mod, const float* ln_gamma,
                 const float* ln_beta, float ln_eps, int64_t n_tokens, int C, int dtype, mp_stream_t stream);

/* Joint embedding + spatial pos-embed + first norm1 (MixSTE.STE_forward, mix_ste.py:128-138):
 *   x[tok, c] = W[c,0:2] . in[tok,0:2] + b[c] + spos[tok % 17, c]  -> x_out fp32, h_out = LN(x; ln_*, eps) 16-bit.
 *   in [n_tokens, 2] fp32, C = 512. */
int mp_embed_joints(const float* in2d, const float* W, const float* b, const float* spos, const float* ln_gamma,
                    const float* ln_beta, float ln_eps, float* x_out, void* h_out, int64_t n_tokens, int n_joints,
                    int C, int dtype, mp_stream_t stream);
/* joints_to_segments_proj + pos-embed + norm1 (BonesMixSTE.forward, manifold_mix_ste.py:139-148):
 *   in [n_frames, 34] fp32, W [16*128, 34], b [2048], spos [16,128] -> x_out fp32 / h_out 16-bit [n_frames*16,128]. */
int mp_embed_segments(const float* in2d, const float* W, const float* b, const float* spos, const float* ln_gamma,
                      const float* ln_beta, float ln_eps, float* x_out, void* h_out, int64_t n_frames,
                      int in_features, int n_segments, int C, int dtype, mp_stream_t stream);

/* Multi-head softmax attention (Attention.forward, mix_ste.py:255-282), fp32 softmax.  head_dim 64 (the C = 512 rotation
 * backbone) runs on tcgen05 with TMEM accumulators (attn_spatial_tc_kernel, attn_temporal_tc2_kernel: S = Q K^T and O = P V as
 * tcgen05.mma, operands by TMA); head_dim 16 (the C = 128 bone-length backbone, 1.7 % of the flops) on the mma.sync kernels.
 *   qkv [n_clips*n_frames*n_tok, 3*C] 16-bit, columns [q|k|v] x heads x head_dim (mix_ste.py:257-261)
 *   out [n_clips*n_frames*n_tok, C] 16-bit;  head_dim in {64, 16}; scale = head_dim^-0.5
 *   MP_ATTN_SPATIAL: sequences = the n_tok tokens of one frame;
 *   MP_ATTN_TEMPORAL: sequences = the n_frames frames of one (clip, token) track (n_frames <= 256). */
enum { MP_ATTN_SPATIAL = 0, MP_ATTN_TEMPORAL = 1 };
int mp_attention(const void* qkv, void* out, int64_t n_clips, int64_t n_frames, int n_tok, int C, int n_heads,
                 int mode, int dtype, mp_stream_t stream);

/* K hypothesis heads (RMCLRotMixSTE.forward tail + MCLHead, rmcl_manifold_mix_ste.py:251-298), all fp32:
 *   x [n_frames_total*17, 512] fp32 = output of the last temporal block BEFORE Temporal_norm;
 *   applies Temporal_norm (post_*; both NULL: x is already normalised, MCLHead.forward on its own), then per head k: LN(eps 1e-5; hg/hb [K,512]) -> Linear(512 -> out_dim
 *   (+1 if with_score); hw [K, out_dim+1, 512], hbias [K, out_dim+1]) -> rot [B,K,T,17,out_dim] fp32 and,
 *   if with_score, logits[B,K,T] = score_w[K,17] . score_emb + score_b[K].
 *   with_score = 0 is MixSTE.head of the single-hypothesis model (mix_ste.py:123-126, K = 1). */
int mp_heads_fwd(const float* x, const float* post_gamma, const float* post_beta, float post_eps,
                 const float* hg, const float* hb, const float* hw, const float* hbias,
                 const float* score_w, const float* score_b, float* rot, float* logits,
                 int64_t n_clips, int64_t n_frames, int n_hyp, int out_dim, int with_score,
                 mp_stream_t stream);
/* The same K heads on the tensor cores (the C = 512 model): xhat [n_frames_total*17, 512] 16-bit is the input ALREADY through
 * Temporal_norm and the affine-free LayerNorm(eps 1e-5) the K heads share (mp_linear_ln writes it as its h output), wf16 [n_pad, 512]
 * 16-bit / bf [n_pad] fp32 the folded parameters (row k*(D+1)+o: gamma_k (.) W_k[o], W_k[o] . beta_k + b_k[o]; rows past K*(D+1) zero):
 * one tcgen05 GEMM with fp32 output into `workspace` (room for [tokens, n_pad] fp32; only the K*(D+1) useful columns are written, as a dense
 * [tokens, ld] matrix), then the scatter into rot / the score dot product. */
int mp_heads_fwd16(const void* xhat16, const void* wf16, const float* bf, const float* score_w, const float* score_b, float* rot,
                   float* logits, float* workspace, size_t workspace_bytes, int64_t n_clips, int64_t n_frames, int n_hyp, int out_dim,
                   int with_score, int n_pad, int dtype, mp_stream_t stream);
/* Bone-length head (MixSTE.head + mean over time, mix_ste.py:123-126,187; manifold_mix_ste.py:150-154):
 *   x [n_clips*n_frames*16, 128] fp32 before Temporal_norm -> bone_len [n_clips,16] fp32.
 *   workspace >= n_clips*n_frames*16*4 bytes. */
int mp_bones_head(const float* x, const float* post_gamma, const float* post_beta, float post_eps,
                  const float* hg, const float* hb, const float* hw, const float* hbias,
                  float* bone_len, int64_t n_clips, int64_t n_frames, int n_segments, int C,
                  void* workspace, size_t workspace_bytes, mp_stream_t stream);

/* fp32 -> 16-bit (weight shadows refreshed by the host wrapper after optimizer.step()). */
int mp_cast_f32_to_16(const float* src, void* dst, int64_t n, int dtype, mp_stream_t stream);

/* Flip test-time-augmentation epilogue (hpe/eval_utils.py:83-142, augmentations/functional.py:7-28): hyp [2B,K,T,17,3] / scores [2B,K,T]
 * where clips [B, 2B) are the forward of the horizontally flipped input; out [B,T,17,3] = (aggregate(hyp[b]) + unflip(aggregate(hyp[B+b]))) / 2
 * for MP_AGG_WEIGHTED_AVE or MP_AGG_BEST_SCORE, unflip = negate x and swap left / right joints.  One launch instead of two aggregations, a
 * flip and an average. */
int mp_aggregate_tta(const float* hyp, const float* scores, int mode, float* out_pose, int64_t B, int64_t K, int64_t T, mp_stream_t stream);

/* Pose-consistency metrics (hpe/mh_so3_hpe/metrics/utils.py:4-20 measure_bones_length; regularizations.py:8-48
 * segments_time_consistency = MPSCE; regularizations.py:103-140 sagittal_symmetry = MPSSE) over poses [n_clips, n_frames, 17, 3]:
 *   seg_mean, seg_var [n_clips,16]: mean and unbiased variance over time of every bone length (bone b = joint b+1 -> its parent);
 *   sym_abs, sym_sq [n_clips,6]:   mean over time of |len[left] - len[right]| and of its square, pairs (bones_left[i], bones_right[i]);
 *   bone_len [n_clips,16,n_frames] or NULL: the bone lengths themselves (the reference's layout).
 * workspace >= mp_pose_consistency_workspace_bytes(n_clips, n_frames). */
size_t mp_pose_consistency_workspace_bytes(int64_t n_clips, int64_t n_frames);
int mp_pose_consistency(const float* poses, int64_t n_clips, int64_t n_frames, float* seg_mean, float* seg_var, float* sym_abs, float* sym_sq,
                        float* bone_len, void* workspace, size_t workspace_bytes, mp_stream_t stream);

/* P-MPJPE, "Protocol #2" (hpe/mh_so3_hpe/metrics/mean_joint_errors.py:144-189): per frame, align pred [n_frames,17,3] to gt with the
 * optimal similarity transform (Procrustes, reflections excluded), then out[0] = sum, out[1] = mean of the joint distances.
 * workspace >= mp_p_mpjpe_workspace_bytes(n_frames). */
size_t mp_p_mpjpe_workspace_bytes(int64_t n_frames);
int mp_p_mpjpe(const float* pred, const float* gt, int64_t n_frames, float* out, void* workspace, size_t workspace_bytes, mp_stream_t stream);

/* 3DPCK and AUC (hpe/mh_so3_hpe/metrics/pck.py:77-198, alignment 'none', no mask) over n_points 3-D points: out[0] = PCK at `threshold`
 * (percent), out[1] = AUC over the thresholds linspace(0, 150, 31) (percent).  Exact integer counts behind both. */
size_t mp_pck_auc_workspace_bytes(void);
int mp_pck_auc(const float* pred, const float* gt, int64_t n_points, float threshold, float* out, void* workspace, size_t workspace_bytes,
               mp_stream_t stream);

/* Clip windowing on the device (PoseSequenceGenerator.__getitem__, hpe/mh_so3_hpe/data/generators.py:106-219): frames2d [N, n_joints, in_chans] /
 * frames3d [N, n_joints, 3] hold all sequences back to back; table (device, int64[n_windows][3]) = {first frame of the window's sequence,
 * sequence length, start frame inside the sequence} (fixed starts :93-104 or the host-sampled random starts :121-127); out2d [n_windows,
 * n_frames, n_joints, in_chans], out3d [n_windows, n_frames, n_joints, 3]; frames past the end of the sequence replicate its last frame
 * (:132-146).  mask (nullable, float [n_windows, n_frames, n_joints]) = the occlusion pattern multiplied into the 2-D input (:166-215); noise
 * (nullable, double [n_windows, n_frames, n_joints, in_chans]) = miss_type "noisy" (:206-210), added in fp64 and rounded once like the reference;
 * flip (nullable, uint8 [n_windows]) with joint_perm (int32 [n_joints], left <-> right) = the PoseFlip transform of the training loader
 * (augmentations/transforms.py:8-31), applied to both outputs before noise and mask like the reference (:150-151). */
int mp_gather_windows(const float* frames2d, const float* frames3d, const int64_t* table, const float* mask, const double* noise,
                      const unsigned char* flip, const int* joint_perm, float* out2d, float* out3d, int64_t n_windows, int64_t n_frames,
                      int n_joints, int in_chans, mp_stream_t stream);

/* ---- backward (training) entry points -------------------------------------------------------------------------------
 * The reference differentiates Block / Attention / Mlp / LayerNorm with torch autograd (mix_ste.py:194-368) and steps
 * torch.optim.Adam (main_h36m_lifting.py:755-761).  Dense contractions of the backward pass reuse mp_linear:
 *   dgrad  dX[M,K] = dY[M,N] W[N,K]      -> mp_linear(A = dY, W = transposed 16-bit shadow [K,N])
 *   wgrad  dW[N,K] += dY^T X             -> mp_wgrad: dY [tokens,N] and X [tokens,K] are read IN PLACE as MN-major tcgen05 operands
 *                                           (rows = tokens = the contraction index), the token contraction is split over the SMs and
 *                                           the partial tiles are added into dW with TMA reduce stores; no transposed copies exist.
 * The bias gradient (a column sum of dY) is mp_colsum16; mp_transpose16 only builds the [K,N] weight shadows of the folded heads. */

/* LayerNorm backward: dx = LN'(x; gamma, eps)(dy) [+ dres]; dgamma += sum dy*xhat, dbeta += sum dy (fp32 atomics; both NULL to skip).
 * dy is fp32 (dy_is_16bit = 0) or `dtype` 16-bit; gamma NULL = no affine; dx may alias dres.  dx16 (may be NULL): also writes
 * 16-bit(rowscale[token] * dx) — the operand of the next backward GEMM (rowscale NULL = 1) — and, with dx16_colsum [C] (may be NULL),
 * adds its column sums there (the bias gradient of the Linear whose output gradient dx16 is).  C in {512, 128}. */
int mp_layernorm_bwd(const float* x, const float* gamma, float eps, const void* dy, int dy_is_16bit, const float* dres, float* dx,
                     float* dgamma, float* dbeta, void* dx16, const float* rowscale, float* dx16_colsum, int64_t n_tokens, int C,
                     int dtype, mp_stream_t stream);
/* exact-erf GELU on a 16-bit pre-activation (training keeps the pre-activation): a = gelu(u); du = da * gelu'(u).  n % 8 == 0. */
int mp_gelu_fwd(const void* u, void* a, int64_t n, int dtype, mp_stream_t stream);
int mp_gelu_bwd(const void* u, const void* da, void* du, int64_t n, int dtype, mp_stream_t stream);
/* the same over an [M, C] matrix, and colsum[C] (fp32) += column sums of du (the fc1 bias gradient) in the same pass.  C % 8 == 0. */
int mp_gelu_bwd_colsum(const void* u, const void* da, void* du, float* colsum, int64_t M, int64_t C, int dtype, mp_stream_t stream);
/* Attention backward (Attention.forward, mix_ste.py:257-275): qkv [tokens,3C], dout [tokens,C] -> dqkv [tokens,3C], all 16-bit, same
 * token layout and modes as mp_attention; P is recomputed, nothing of size L x L is stored.  head_dim 64 with spatial sequences of <= 32
 * tokens or temporal tracks of <= 128 frames: tcgen05 kernel over block-diagonal 128-row tiles (S and dP in tensor memory; o is not read,
 * D_i = sum_j P_ij dP_ij).  Otherwise (head_dim 16, L <= 256): mma.sync kernel, one CTA per (sequence, head), reads o.
 * dqkv_colsum [3C] (may be NULL): += column sums of dqkv (the qkv bias gradient). */
int mp_attention_bwd(const void* qkv, const void* o, const void* dout, void* dqkv, float* dqkv_colsum, int64_t n_clips, int64_t n_frames,
                     int n_tok, int C, int n_heads, int mode, int dtype, mp_stream_t stream);
/* Weight gradient dW[n_out, k_in] (fp32) += dY[tokens, n_out]^T X[tokens, k_in] (16-bit), both operands read in place as MN-major
 * UMMA operands (no transposed copies); the token contraction is split over the SMs, partial tiles added with TMA reduce stores.
 * n_out % 128 == 0, k_in % 128 == 0. */
int mp_wgrad(const void* dY, const void* X, float* dW, int64_t n_tokens, int64_t n_out, int64_t k_in, int dtype, mp_stream_t stream);
/* colsum[C] (fp32) += column sums of a 16-bit [M, C] matrix (bias gradients); C % 8 == 0, src 16-byte aligned. */
int mp_colsum16(const void* src, float* colsum, int64_t M, int64_t C, int dtype, mp_stream_t stream);
/* Training path of the K hypothesis heads (MCLHead, rmcl_manifold_mix_ste.py:267-298).  `params` / `grads`: device tables int64 [6][K]
 * of pointers to the heads' fp32 parameters / gradient buffers, rows = {norm.weight [C], norm.bias [C], prediction_head.weight [D+1, C],
 * prediction_head.bias [D+1], score_head.weight [17], score_head.bias [1]}.
 *   mp_heads_fold      the folded Linear of mp_heads_fwd16: wf16 [n_pad, C] (and its transpose wt16 [C, n_pad], may be NULL), bf [n_pad],
 *                      plus the stacked score weights score_w [K, 17] / score_b [K]
 *   mp_heads_bwd_pack  dY [tokens, n_pad] 16-bit from d_rot [B,K,T,17,D] and d_logits [B,K,T] (y: the workspace of the mp_heads_fwd16 call of the
 *                      same shape, i.e. its fp32 GEMM output in the layout that call left it in); dbf [n_pad] += column sums of dY; score_head gradients are accumulated through `grads`
 *   mp_heads_unfold    dWf [n_pad, C] (= mp_wgrad(dY, xhat)) and dbf -> the heads' norm / prediction_head gradients, accumulated */
int mp_heads_fold(const int64_t* params, int n_hyp, int out_dim, int C, int n_pad, void* wf16, void* wt16, float* bf, float* score_w,
                  float* score_b, int dtype, mp_stream_t stream);
int mp_heads_bwd_pack(const float* d_rot, const float* d_logits, const float* y, const float* score_w, void* dy16, float* dbf,
                      const int64_t* grads, int64_t n_clips, int64_t n_frames, int n_hyp, int out_dim, int n_pad, int dtype,
                      mp_stream_t stream);
int mp_heads_unfold(const int64_t* params, const int64_t* grads, const float* dwf, const float* dbf, int n_hyp, int out_dim, int C,
                    mp_stream_t stream);
/* Refresh the 16-bit shadows of n_weights GEMM weights in ONE launch (after an optimizer step).  table (device, int64[n_weights][5]) =
 * {fp32 source pointer, 16-bit shadow [rows, cols] pointer, transposed shadow [cols, rows] pointer or 0, rows, cols}; rows, cols % 64 == 0;
 * max_tiles = max over weights of rows * cols / 4096. */
int mp_refresh_shadows(const int64_t* table, int n_weights, int max_tiles, int dtype, mp_stream_t stream);
/* dst[C,Mpad] = src[M,C]^T (zero padded), colsum[C] += column sums of src (NULL to skip).  C % 64 == 0, Mpad % 64 == 0. */
int mp_transpose16(const void* src, void* dst, float* colsum, int64_t M, int64_t C, int64_t Mpad, int dtype, mp_stream_t stream);
/* out[(row / div) % mod, :] += x[row, :]  (gradients of Spatial_pos_embed / Temporal_pos_embed, mix_ste.py:137,149). */
int mp_group_rowsum(const float* x, float* out, int64_t n_rows, int C, int64_t div, int64_t mod, mp_stream_t stream);
/* dW[n_out,n_in] += dy^T in, db[n_out] += column sums of dy; fp32, n_in in {2,3,34,51} (Spatial_patch_to_embedding, joints_to_segments_proj). */
int mp_small_wgrad(const float* dy, const float* in, float* dW, float* db, int64_t n_rows, int n_out, int n_in, mp_stream_t stream);
/* Stochastic depth (timm DropPath on both residual branches, mix_ste.py:334-336,352-358): out = x + s[token] * y (y 16-bit branch
 * output, s = keep mask / keep_prob of the token's sample), and the backward operand out16 = 16-bit(s[token] * g) (s NULL: plain cast). */
int mp_residual_rowscale(const float* x, const void* y, const float* s, float* out, int64_t n_tokens, int C, int dtype, mp_stream_t stream);
int mp_cast_rowscale(const float* g, const float* s, void* out, int64_t n_tokens, int C, int dtype, mp_stream_t stream);
/* torch.optim.Adam step (L2-style weight_decay, bias correction, step counted from 1) over one flat fp32 buffer; grad is read as
 * grad * grad_scale (1 / world_size after a SUM all-reduce).  step_dev (device int64, may be NULL): when given, the step count is read
 * from it (value before this step) and incremented afterwards, so a captured CUDA graph of the step stays correct on replay.
 * lr_dev (device float, may be NULL): when given, the learning rate is read from it instead of `lr`, so an lr scheduler
 * (CosineAnnealingLR / ReduceLROnPlateau, main_h36m_lifting.py:763-771) keeps working across replays of a captured step. */
int mp_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr, float beta1, float beta2, float eps,
                 float weight_decay, int64_t step, int64_t* step_dev, const float* lr_dev, float grad_scale, mp_stream_t stream);

/* Per-point / joint-wise / coordinate-wise errors of the evaluation drivers (mean_joint_errors.py:31-141: mpjpe_error(no_agg),
 * mse_error, jointwise_error, jointwise_mse, coordwise_error; segments_len_err over the bone lengths of mp_pose_consistency).
 *   mode MP_ERR_L2 / MP_ERR_SQ: pred, gt hold n_elem 3-D points; e_i = ||gt_i - pred_i||_2 (or its square)
 *   mode MP_ERR_ABS / MP_ERR_DIFF: pred, gt hold n_elem scalars;   e_i = |gt_i - pred_i| (or the signed difference)
 *   per_elem (may be NULL): e [n_elem] (the reference's mode "no_agg");
 *   col_out  (may be NULL): [cols] with col_out[c] = scale * sum of e_i over i = c (mod cols) -- cols = 17 joints, 3 coordinates, or 1
 *   for a plain sum / mean; deterministic (fp64 partials combined in a fixed order).  workspace: mp_point_errors_workspace_bytes. */
enum { MP_ERR_L2 = 0, MP_ERR_SQ = 1, MP_ERR_ABS = 2, MP_ERR_DIFF = 3 };
size_t mp_point_errors_workspace_bytes(int64_t n_elem, int cols);
int mp_point_errors(const float* pred, const float* gt, int64_t n_elem, int cols, int mode, float scale, float* per_elem, float* col_out,
                    void* workspace, size_t workspace_bytes, mp_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* MANIPOSE_SM100_H_ */
