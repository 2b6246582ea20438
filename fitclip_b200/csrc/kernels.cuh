// Internal (C++) interface between the C-ABI layer (api.cu) and the kernel translation units.
#pragma once
#include <limits.h>

#include "common.cuh"

namespace fc {

enum DType : int { FC_DTYPE_F32 = 0, FC_DTYPE_BF16 = 1, FC_DTYPE_F16 = 2 };

// elementwise.cu
// stats_out (optional): per-row (sum, sumsq) of the OUTPUT, layout [rows, D/64, 2]
int layernorm_bf16(const bf16* x, int64_t ldx, bf16* y, int64_t ldy, const float* gamma, const float* beta,
                   int64_t rows, int D, float eps, float* stats_out, cudaStream_t s);
// Row statistics alone: stats_out[rows, D/64, 2] = per-row (sum, sum of squares) of x (everything in part 0) -- what a
// folded-LayerNorm GEMM needs when no LayerNorm kernel ran in front of it (timm's VisionTransformer has no ln_pre).
int row_stats_bf16(const bf16* x, int64_t ldx, int64_t rows, int D, float* stats_out, cudaStream_t s);
int fold_ln_weights(const float* W, const float* gamma, const float* beta, const float* bias, bf16* Wf, float* colsum,
                    float* bias_f, int N, int K, cudaStream_t s);
int im2col_patches(const void* frames, int dtype, bf16* patches, int64_t F, int R, int P, cudaStream_t s);
int cls_rows(bf16* x, const float* cls, const float* pos, int64_t F, int L, int D, cudaStream_t s);
int text_embed(const int32_t* ids, const float* tok, const float* pos, bf16* x, int64_t C, int L, int D, int vocab,
               int* err_flag, float* stats_out, cudaStream_t s);
int head_project(const bf16* x, const int32_t* ids, const float* gamma, const float* beta, const float* proj,
                 float* out, int64_t seqs, int L, int W, int E, float eps, cudaStream_t s);
int pool_normalize(const float* x, float* out, bf16* out_bf16, int64_t B, int T, int D, float scale, cudaStream_t s);
int wise_lerp(const float* p1, const float* p2, float* out, bf16* out_bf16, int64_t n, double w, cudaStream_t s);
int f32_to_bf16(const float* in, bf16* out, int64_t n, cudaStream_t s);
// (rows, cols) fp32 -> (rows, ld >= cols) bf16, padding columns zeroed
int f32_to_bf16_padded(const float* in, bf16* out, int64_t rows, int cols, int ld, cudaStream_t s);
int split_bf16(const float* in, bf16* out, int64_t rows, int D, int mode, int terms, cudaStream_t s);

// preprocess.cu : uint8 (F,H,W,3) -> [x/255 -> bicubic resize (shorter side = size) -> centre crop -> normalise] -> (F,3,size,size)
// interpolation: 0 = bicubic (A = -0.75), 1 = bilinear
int preprocess_frames(const uint8_t* frames, int64_t F, int H, int W, int size, const float* mean, const float* stdv,
                      void* out, int out_dtype, int interpolation, cudaStream_t s);
// uint8 frames -> the bf16 patch matrix [F * (size/patch)^2, ldp] the patch-embedding GEMM reads (no NCHW intermediate)
int preprocess_to_patches(const uint8_t* frames, int64_t F, int H, int W, int size, int patch, const float* mean,
                          const float* stdv, bf16* patches, int ldp, int interpolation, cudaStream_t s);

// attention.cu : out[s*L + l, h*64 + d] = softmax(q k^T / 8 [+ causal mask]) v, qkv rows are [q | k | v] of width 3*D
int attention_bf16(const bf16* qkv, bf16* out, int64_t seqs, int L, int heads, int causal, cudaStream_t s);
// attention_tc.cu : tcgen05 path for the un-masked 193..208-token case; *handled = 1 when it took the call
int attention_bf16_tc(const bf16* qkv, bf16* out, int64_t seqs, int L, int heads, int causal, cudaStream_t s,
                      int* handled);
// tcgen05 kernel for un-masked sequences of 209..768 tokens (key-block loop, online softmax): attention_tc_long.cu
int attention_bf16_tc_long(const bf16* qkv, bf16* out, int64_t seqs, int L, int heads, int causal, cudaStream_t s,
                           int* handled);

// attention_bwd_tc.cu : tcgen05 attention backward for L <= 208; *handled = 1 when it took the call
int attention_bwd_bf16_tc(const bf16* qkv, const bf16* O, const bf16* dO, bf16* dqkv, int64_t seqs, int L, int heads,
                          int causal, cudaStream_t s, int* handled);

// rank.cu
int rank_from_scores(const float* S, int64_t ld, int64_t rows, int64_t cols, const int32_t* target, int64_t* ranks,
                     cudaStream_t s);
int metrics_from_ranks(const int64_t* ranks, int64_t n, int64_t num_candidates, float* out_recall3,
                       int64_t* out_median, float* out_mean, cudaStream_t s);
int counts_to_ranks(const int32_t* counts, int64_t* ranks, int64_t n, cudaStream_t s);
int topk_rows(const float* S, int64_t ld, int64_t rows, int64_t cols, int k, float* out_val, int32_t* out_idx,
              cudaStream_t s);
int nce_loss(const float* S, int64_t ld, int B, float* workspace, float* out, cudaStream_t s);
int ts_nce_loss(const float* S, const float* Tt, int64_t ld, int B, float* workspace, float* out, cudaStream_t s);

}  // namespace fc
