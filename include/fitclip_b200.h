/* libfitclip_b200 -- C ABI of the B200-native FitCLIP evaluation hot path.
 *
 * The reference (bryant1410/fitclip) is pure Python: its boundary for this path is the `VideoTextEncoder` plugin ABC
 * (aligner/encoder/video_encoder.py:14-52, aligner/encoder/video_text_encoder.py:15-31), not an FFI.  This header is
 * what the Python drop-in (`fitclip_b200.encoder.B200ClipVideoTextEncoder`, bound with ctypes -- see INTEGRATION.md)
 * calls underneath; each entry point cites the reference call it replaces.
 *
 * Conventions
 *   - every function returns 0 on success or a negative fc_status; fc_last_error() gives the text (thread-local).
 *     Nothing throws or aborts across the boundary.
 *   - every pointer is CALLER-OWNED CUDA DEVICE memory unless it says "host"; no ownership is transferred.
 *   - work is enqueued on the caller's `stream` (a cudaStream_t passed as void*) and is asynchronous.
 *   - no hidden allocation after fc_model_create(); a handle is bound to the device current at creation,
 *     is not thread-safe, and needs compute capability 10.x (FC_ERR_ARCH otherwise -- there is no CPU fallback).
 *   - matrices are row-major; "ld" is the row stride in elements.
 */
#ifndef FITCLIP_B200_H_
#define FITCLIP_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FC_API __attribute__((visibility("default")))

typedef enum fc_status {
  FC_STATUS_OK = 0,
  FC_STATUS_INVALID = -1, /* bad argument: null pointer, shape, alignment */
  FC_STATUS_CUDA = -2,    /* a CUDA call failed (message has the CUDA error string) */
  FC_STATUS_ARCH = -3,    /* device is not sm_100 */
  FC_STATUS_STATE = -4,   /* e.g. encode before every parameter was loaded */
  FC_STATUS_NOMEM = -5
} fc_status;

typedef enum fc_dtype { FC_F32 = 0, FC_BF16 = 1, FC_F16 = 2 } fc_dtype;

/* GEMM epilogues exposed for kernel-level parity tests (fc_gemm_bf16). */
typedef enum fc_epilogue {
  FC_EPI_BIAS = 0,       /* C(bf16) = A.B^T + bias                                  (K3)  */
  FC_EPI_BIAS_QGELU = 1, /* C(bf16) = quickgelu(A.B^T + bias), x*sigmoid(1.702x)    (K6)  */
  FC_EPI_BIAS_RESID = 2, /* C(bf16) = resid + A.B^T + bias                          (K5)  */
  FC_EPI_F32 = 4         /* C(fp32) = alpha * A.B^T                                 (K11/K12) */
} fc_epilogue;

/* Geometry of clip.model.CLIP(**kwargs): config/encoder/clip_from_scratch_vit_b_16.yaml:5-16. */
typedef struct fc_config {
  int32_t embed_dim;
  int32_t image_resolution;
  int32_t vision_layers;
  int32_t vision_width;
  int32_t vision_patch_size;
  int32_t context_length;
  int32_t vocab_size;
  int32_t transformer_width;
  int32_t transformer_heads;
  int32_t transformer_layers;
  int32_t max_frames_per_pass; /* max frames per internal pass (workspace is sized for it; passes are equal-sized); 0 = 512 */
  int32_t max_texts_per_pass;  /* max captions per internal pass; 0 = 1024 */
  /* Vision tower architecture.  FC_TOWER_OPENAI (0): clip.model.VisionTransformer -- ln_pre, QuickGELU, LayerNorm eps 1e-5.
   * FC_TOWER_TIMM (1): timm's VisionTransformer as the SLIP-layout models build it (aligner/encoder/slip.py:585-627,
   * `timm.create_model('vit_*_patch16_224', num_classes=0)`): no ln_pre, exact (erf) GELU, LayerNorm eps 1e-6; the
   * caller passes its weights under the OpenAI names (the patch-embedding bias folded into the positional rows 1..,
   * `norm` as `ln_post`, `image_projection` as `visual.proj`) and `visual.ln_pre.*` is neither expected nor accepted. */
  int32_t vision_tower;
  /* Width of the image tower's attention (q / k / v rows of in_proj, columns of out_proj): heads * 64.  0 = vision_width.
   * A tower whose heads are narrower than 64 (the SLIP ViT-S/16: 384 wide, 12 heads of 32, slip.py:566-569) is run by giving
   * every head a 64-wide slot: the caller passes in_proj_weight (3 * attn_width, vision_width) / in_proj_bias with the extra
   * rows zero and the q rows scaled by sqrt(64 / head_dim) (the kernels' softmax scale is 1/8), and out_proj.weight
   * (vision_width, attn_width) with the extra columns zero -- the zero dimensions add nothing to q.k and carry no value. */
  int32_t vision_attn_width;
} fc_config;
enum { FC_TOWER_OPENAI = 0, FC_TOWER_TIMM = 1 };

typedef struct fc_model fc_model;

FC_API int fc_version(void);
/* Copies the last error message of the calling thread into buf (NUL-terminated); returns its full length. */
FC_API size_t fc_last_error(char* buf, size_t cap);
/* Number of kernel launches issued by this library in this process so far (bench.py's gpu_launches). */
FC_API int64_t fc_launch_count(void);

/* Optional per-launch timing with CUDA events on the launching stream (bench.py's roofline numbers). Records are
 * aggregated by (kind, tag, n, k): kind 0 = tcgen05 GEMM (tag = epilogue), 1 = attention (tag = causal),
 * 2 = LayerNorm, 3 = other memory-bound kernels.  fc_profile_stop synchronises the device and returns the record count. */
typedef struct fc_profile_record {
  int32_t kind, tag;
  int64_t n, k;
  int64_t launches;
  double ms;    /* summed launch durations */
  double flops; /* summed algorithmic FLOPs */
  double bytes; /* summed algorithmic HBM bytes */
  double rows;  /* summed M */
} fc_profile_record;
FC_API int fc_profile_start(int32_t max_records);
FC_API int fc_profile_stop(fc_profile_record* out, int32_t cap);

/* ---- model lifetime and weights --------------------------------------------------------------------------------
 * Replaces clip.load / build_model as used by load_clip_model (aligner/encoder/clip_video_text_encoder.py:22-61):
 * the Python side owns the fp32 nn.Parameters (OpenAI state-dict names); the library keeps its own bf16 copies of
 * the GEMM weights and fp32 copies of biases / LayerNorm / embeddings / projections. */
FC_API int fc_model_create(const fc_config* cfg, fc_model** out);
FC_API int fc_model_destroy(fc_model* m);
/* `name` is an OpenAI-CLIP parameter name ("visual.conv1.weight", "transformer.resblocks.3.attn.in_proj_weight",
 * "token_embedding.weight", ...; "logit_scale" is accepted and ignored, clip_video_text_encoder.py:75-77).
 * `data` is a contiguous fp32 device tensor of `numel` elements. */
FC_API int fc_model_set_param(fc_model* m, const char* name, const float* data, int64_t numel, void* stream);
/* 1 when every parameter has been set, else 0 (and the first missing name in fc_last_error). */
FC_API int fc_model_ready(fc_model* m);
FC_API int64_t fc_model_workspace_bytes(const fc_model* m);

/* ---- encoders ---------------------------------------------------------------------------------------------------
 * ClipVideoTextEncoder.encode_video (clip_video_text_encoder.py:80-89): frames (videos*frames_per_video, 3, R, R) of
 * `dtype`, CLIP-normalised pixels -> CLIP.encode_image -> x/||x|| per FRAME -> mean over the frames of each video
 * (not re-normalised).  out_video: fp32 (videos, embed_dim).  out_frames (optional, may be NULL): the un-normalised
 * per-frame image features, fp32 (videos*frames_per_video, embed_dim). */
FC_API int fc_encode_video(fc_model* m, const void* frames, int dtype, int64_t videos, int32_t frames_per_video,
                           float* out_video, float* out_frames, void* stream);
/* ClipVideoTextEncoder.encode_text (clip_video_text_encoder.py:92-94): int32 token ids (texts, context_length) ->
 * CLIP.encode_text (EOT = first argmax of the ids) -> x/||x||.  out_text: fp32 (texts, embed_dim).
 * Out-of-vocabulary ids are reported as FC_STATUS_INVALID by fc_model_check() after the stream has run. */
FC_API int fc_encode_text(fc_model* m, const int32_t* ids, int64_t texts, float* out_text, void* stream);
/* Synchronises `stream` and reports deferred device-side input errors (bad token ids). */
FC_API int fc_model_check(fc_model* m, void* stream);

/* ---- evaluation pre-processing (the step before the path; the reference runs it on CPU DataLoader workers) ------
 * ClipVideoTextEncoder.get_eval_transform (clip_video_text_encoder.py:124-133): uint8 frames (n, H, W, 3) ->
 * x/255 -> resize, shorter side = `size` (align_corners = false, no antialias) -> centre crop `size` x `size` ->
 * (x - mean[c]) / std[c] -> (n, 3, size, size) of out_dtype (FC_F32 or FC_BF16), which is what fc_encode_video takes.
 * interpolation: FC_INTERP_BICUBIC (A = -0.75; CLIP's transform) or FC_INTERP_BILINEAR (Resize's default, which
 * SlipVideoTextEncoder.get_eval_transform keeps, slip_video_text_encoder.py:78-87).  mean / std: 3 HOST floats each.
 * n <= 65535 per call. */
enum { FC_INTERP_BICUBIC = 0, FC_INTERP_BILINEAR = 1 };
FC_API int fc_preprocess_frames(const uint8_t* frames, int64_t n, int32_t H, int32_t W, int32_t size,
                                const float* mean, const float* std, void* out, int out_dtype, int32_t interpolation,
                                void* stream);

/* Same transform, written as the bf16 patch matrix of the ViT patch embedding instead of an NCHW image: pixel (c, y, x)
 * of frame f -> row f*G*G + (y/patch)*G + x/patch, column c*patch*patch + (y%patch)*patch + x%patch, G = size/patch,
 * row stride ldp elements (>= 3*patch*patch; pad columns are NOT written) -- conv1.weight.reshape(width, -1)'s column
 * order, i.e. the A operand of the patch-embedding GEMM (VisionTransformer.forward's conv1 as a GEMM). */
FC_API int fc_preprocess_to_patches(const uint8_t* frames, int64_t n, int32_t H, int32_t W, int32_t size, int32_t patch,
                                    const float* mean, const float* std, void* patches, int64_t ldp,
                                    int32_t interpolation, void* stream);
/* The two steps above in one call, for raw decoded frames: uint8 (videos*frames_per_video, H, W, 3) -> eval transform
 * (same arithmetic as fc_preprocess_frames) written STRAIGHT into the bf16 patch matrix of the patch-embedding GEMM
 * (the normalised NCHW frame never exists in memory) -> fc_encode_video's path.  Replaces the CPU DataLoader transform
 * of clip_video_text_encoder.py:124-133 + encode_video (:80-89); 4x fewer host->device bytes than fp32 frames. */
FC_API int fc_encode_video_uint8(fc_model* m, const uint8_t* frames, int64_t videos, int32_t frames_per_video, int32_t H,
                                 int32_t W, const float* mean, const float* std, int32_t interpolation, float* out_video,
                                 float* out_frames, void* stream);

/* ---- pooling / WiSE ---------------------------------------------------------------------------------------------
 * out[b] = scale * mean_t( x[b*T+t] / ||x[b*T+t]||_2 ), fp32 (clip_video_text_encoder.py:85-89; T = 1 is the text
 * normalisation of :94).  out_bf16 may be NULL. */
FC_API int fc_pool_normalize(const float* x, float* out, void* out_bf16, int64_t rows_out, int32_t T, int32_t D,
                             float scale, void* stream);
/* wise_state_dict (aligner/wise.py:10-16): out = (1 - w) * p1 + w * p2 elementwise over n fp32 values, evaluated as
 * two rounded products and a rounded sum (bit-exact with torch).  out may alias p1 or p2; out_bf16 may be NULL. */
FC_API int fc_wise_lerp(const float* p1, const float* p2, float* out, void* out_bf16, int64_t n, double w,
                        void* stream);

/* ---- similarity + rank (rows = texts / queries, columns = videos / classes) ------------------------------------
 * TextVideoRetrievalLightningModule._validate_dataset (aligner/text_video_retrieval.py:67-83) without materialising
 * the Nt x Nv matrix.  `terms` = 1: operands rounded to bf16; 3: split-bf16 (hi.hi + hi.lo + lo.hi ~ fp32 product).
 * The column slab may be a shard: `col_offset` is the global index of local column 0, `target[i]` the GLOBAL target
 * column of row i.  Protocol: prepare -> target_scores (tscore zeroed by caller; all-reduce(sum) across shards)
 *                                   -> count (counts zeroed by caller; all-reduce(sum) across shards) = 0-based ranks.
 * Tie rule: rank_i = #{j: s_ij > s_it} + #{j < t: s_ij == s_it}. */
FC_API int64_t fc_sim_workspace_bytes(int64_t nt, int64_t nv, int32_t dim, int32_t terms);
FC_API int fc_sim_prepare(const float* text_emb, const float* video_emb, int64_t nt, int64_t nv, int32_t dim,
                          int32_t terms, void* workspace, void* stream);
FC_API int fc_sim_target_scores(const void* workspace, int64_t nt, int64_t nv, int32_t dim, int32_t terms,
                                const int32_t* target, int32_t col_offset, float* tscore, void* stream);
FC_API int fc_sim_count(const void* workspace, int64_t nt, int64_t nv, int32_t dim, int32_t terms,
                        const int32_t* target, int32_t col_offset, const float* tscore, int32_t* counts, void* stream);
/* Materialise S = alpha * text . video^T as fp32 (nt, ld) -- `scores = encoded_texts @ encoded_videos.T`
 * (text_video_retrieval.py:74) and the scaled per-batch scores of :49-50 / video_text_module.py:62-63. */
FC_API int fc_sim_scores(const void* workspace, int64_t nt, int64_t nv, int32_t dim, int32_t terms, float alpha,
                         float* scores, int64_t ld, void* stream);

/* Rank.update on a materialised matrix (aligner/metrics.py:16-19): ranks[i] (int64, 0-based), same tie rule. */
FC_API int fc_rank_from_scores(const float* scores, int64_t ld, int64_t rows, int64_t cols, const int32_t* target,
                               int64_t* ranks, void* stream);
FC_API int fc_counts_to_ranks(const int32_t* counts, int64_t* ranks, int64_t n, void* stream);
/* Recall(top_k=1|5|10) as hits/n (fp32 x3), MedianRank = lower median + 1 (int64; aligner/metrics.py:33-36),
 * MeanRank = mean + 1 (fp32, may be NULL; :27-30).  num_candidates = number of columns ranked against. */
FC_API int fc_metrics_from_ranks(const int64_t* ranks, int64_t n, int64_t num_candidates, float* recall_1_5_10,
                                 int64_t* median_rank, float* mean_rank, void* stream);
/* Per-row top-k (k <= 16), value descending then index ascending. */
FC_API int fc_topk_rows(const float* scores, int64_t ld, int64_t rows, int64_t cols, int32_t k, float* values,
                        int32_t* indices, void* stream);

/* ---- per-batch losses, forward only (aligner/loss.py:13-39); scores are (B, ld) fp32; workspace >= 2*B floats ---- */
FC_API int fc_nce_loss(const float* scores, int64_t ld, int32_t B, float* workspace, float* out, void* stream);
/* TeacherStudentNCELoss(reduction="batchmean") (aligner/teacher_student.py:73). */
FC_API int fc_ts_nce_loss(const float* scores, const float* teacher_scores, int64_t ld, int32_t B, float* workspace,
                          float* out, void* stream);

/* ---- training step: TeacherStudentLightningModule.training_step / _dataset_step_end
 * (aligner/teacher_student.py:99-183) + torch.optim.AdamW (config/trainer.yaml:22-24).  The reference differentiates
 * with torch.autograd; these are the explicit gradient kernels the Python trainer (fitclip_b200/training.py) chains.
 * All activations are bf16 row-major (tokens, width); gradients of parameters are fp32 and ACCUMULATE (+=). -------- */
/* C(fp32, M x N, zeroed or holding a running sum) += alpha * A[M,K] . B[N,K]^T, the K range cut into `k_splits` work
 * items per output tile (0 = enough to fill the SMs): the weight-gradient GEMM dW = dY^T . X, whose K is the token count. */
FC_API int fc_gemm_bf16_splitk(const void* A, int64_t lda, const void* B, int64_t ldb, float* C, int64_t ldc,
                               float alpha, int32_t M, int32_t N, int32_t K, int32_t k_splits, void* stream);
/* The tcgen05 GEMM with operands read in place as the TRANSPOSE of a row-major matrix ("MN-major"): a_mn: A is stored
 * (K, M) with row stride lda; b_mn: B is stored (K, N) with row stride ldb.  Built combinations:
 *   FC_EPI_BIAS / FC_EPI_BIAS_RESID / FC_EPI_F32 with b_mn   dgrad  dX = dY . W     (B = W as stored, (N_w, K_w))
 *   epilogue 11 with b_mn: C = (dY . W) o quickgelu'(resid)  the same dgrad fused with QuickGELU's backward (resid = the
 *                                                            pre-activation kept by the forward; bias unused)
 *   epilogue 9 (split-K, C fp32 +=) with a_mn and b_mn       wgrad  dW = dY^T . X   (A = dY, B = X as stored)
 * so the backward pass needs no transposed copies of weights or activations. */
FC_API int fc_gemm_bf16_layout(int epilogue, int a_mn, int b_mn, const void* A, int64_t lda, const void* B,
                               int64_t ldb, void* C, int64_t ldc, const float* bias, const void* resid, int64_t ldr,
                               float alpha, int32_t M, int32_t N, int32_t K, int32_t k_splits, void* stream);
/* colsum (cols) fp32 += column sums of the bf16 matrix x (rows, ld): the bias gradient of a Linear. */
FC_API int fc_colsum_bf16(const void* x, int64_t ld, int64_t rows, int32_t cols, float* colsum, void* stream);
/* out (cols, ld_out) = transpose of the kept rows of in (rows, ld_in): rows come in groups of `group_len` whose first
 * `group_skip` are dropped (0, 0 = keep all); columns [kept_rows, ld_out) are zero-filled.  colsum (optional, fp32
 * (cols)) += column sums of the kept rows -- the bias gradient of a Linear whose output gradient is `in`. */
FC_API int fc_transpose_bf16(const void* in, int64_t ld_in, void* out, int64_t ld_out, int64_t kept_rows, int32_t cols,
                             int32_t group_len, int32_t group_skip, float* colsum, void* stream);
/* LayerNorm backward (slip.py:350-356): dx = [add +] dLN(x, dy); dgamma += , dbeta += .  dx may alias add. */
FC_API int fc_layernorm_bwd_bf16(const void* x, const void* dy, const float* gamma, const void* add, void* dx,
                                 float* dgamma, float* dbeta, int64_t rows, int32_t D, float eps, void* stream);
/* QuickGELU (slip.py:359-361) forward g = u sigmoid(1.702 u) and backward du = dg * dg/du; du may alias dg.  g_out
 * (optional) also receives quickgelu(u) in the same pass (the c_proj weight gradient reads it). */
FC_API int fc_quickgelu_bf16(const void* u, void* g, int64_t n, void* stream);
FC_API int fc_quickgelu_bwd_bf16(const void* u, const void* dg, void* du, void* g_out, int64_t n, void* stream);
/* Backward of fc_attention_bf16: qkv / dqkv (seqs*L, 3*heads*64), out / dout (seqs*L, heads*64); L <= 432. */
FC_API int fc_attention_bwd_bf16(const void* qkv, const void* out, const void* dout, void* dqkv, int64_t seqs,
                                 int32_t L, int32_t heads, int32_t causal, void* stream);
/* nce_loss (teacher == NULL, rows == cols; aligner/loss.py:13-26) or TeacherStudentNCELoss("batchmean") (loss.py:29-39,
 * teacher_student.py:73) of (rows, ld) fp32 scores -- rows = videos, cols = texts; they differ when the unlabelled texts
 * were replaced by a prompt list (teacher_student.py:104-120): the row direction is then averaged over `rows`, the
 * column direction over `cols`.  value (optional) and gscale * d loss / d scores (optional).
 * lse: 2 * (rows + cols) floats of workspace. */
FC_API int fc_loss_fwd_bwd(const float* scores, const float* teacher, int64_t ld, int32_t rows, int32_t cols, float* lse,
                           float gscale, float* loss_out, float* dscores, int64_t ldd, void* stream);
/* C = alpha * op(A) . op(B) in fp32 (score-matrix gradients dV = dS . T, dT = dS^T . V); trans_x: operand stored transposed. */
FC_API int fc_sgemm_f32(int32_t trans_a, int32_t trans_b, int32_t M, int32_t N, int32_t K, float alpha, const float* A,
                        int64_t lda, const float* B, int64_t ldb, float* C, int64_t ldc, void* stream);
/* Backward of fc_pool_normalize: x fp32 (rows_out*T, D), dout fp32 (rows_out, D) -> dx bf16 (rows_out*T, D). */
FC_API int fc_pool_normalize_bwd(const float* x, const float* dout, void* dx_bf16, int64_t rows_out, int32_t T,
                                 int32_t D, float scale, void* stream);
/* Row of each sequence that feeds the head: EOT = first argmax of ids (slip.py:478), or row 0 (class token) when ids is
 * NULL.  scatter = 0: rows[s] = x[s, eot];  1: x[s, eot] = rows[s] (x zeroed by the caller). */
FC_API int fc_seq_rows(void* x, const int32_t* ids, void* rows, int64_t seqs, int32_t L, int32_t W, int32_t scatter,
                       void* stream);
/* out (L, W) fp32 += sum over sequences of dx (seqs, L, W): positional-embedding (and class-embedding) gradient. */
FC_API int fc_seq_sum(const void* dx, float* out, int64_t seqs, int32_t L, int32_t W, void* stream);
/* dtok[ids[t]] += dx[t]: token_embedding gradient. */
FC_API int fc_token_scatter_add(const int32_t* ids, const void* dx, float* dtok, int64_t tokens, int32_t W,
                                int32_t vocab, void* stream);
/* One AdamW step over a flat fp32 buffer (decoupled weight decay, bias-corrected; step counts from 1); p_bf16
 * (optional) receives the rounded parameters. */
FC_API int fc_adamw_step(float* p, const float* g, float* m, float* v, void* p_bf16, int64_t n, float lr, float beta1,
                         float beta2, float eps, float weight_decay, int32_t step, void* stream);
FC_API int fc_f32_to_bf16(const float* in, void* out, int64_t n, void* stream);
/* Training-forward front ends: conv1 patch embedding + class token + positional embedding (patches: bf16 scratch
 * (F*G*G, 3*P*P), kept for the weight gradient), and token + positional embedding. */
FC_API int fc_patch_embed(const void* frames, int dtype, const void* conv_w_bf16, const float* cls, const float* pos,
                          void* patches, void* x, int64_t F, int32_t R, int32_t P, int32_t W, void* stream);
FC_API int fc_text_embed(const int32_t* ids, const float* tok, const float* pos, void* x, int64_t C, int32_t L,
                         int32_t W, int32_t vocab, int32_t* err_flag, void* stream);

/* ---- kernel-level entry points (parity tests, microbenchmarks) -------------------------------------------------- */
/* C[M,N] = epilogue(A[M,K] . B[N,K]^T): bf16 operands, K contiguous, fp32 accumulation in TMEM (tcgen05.mma). */
FC_API int fc_gemm_bf16(int epilogue, const void* A, int64_t lda, const void* B, int64_t ldb, void* C, int64_t ldc,
                        const float* bias, const void* resid, int64_t ldr, float alpha, int32_t M, int32_t N,
                        int32_t K, void* stream);
FC_API int fc_layernorm_bf16(const void* x, void* y, const float* gamma, const float* beta, int64_t rows, int32_t D,
                             float eps, void* stream);
/* qkv: bf16 (seqs*L, 3*heads*64) rows [q|k|v]; out: bf16 (seqs*L, heads*64). */
FC_API int fc_attention_bf16(const void* qkv, void* out, int64_t seqs, int32_t L, int32_t heads, int32_t causal,
                             void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FITCLIP_B200_H_ */
