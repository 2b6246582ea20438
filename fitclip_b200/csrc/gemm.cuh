// Host-visible interface of the tcgen05 GEMM (gemm.cu).
#pragma once
#include "common.cuh"

namespace fc {

enum Epilogue : int {
  EPI_BIAS = 0,        // C(bf16) = acc + bias
  EPI_BIAS_QGELU = 1,  // C(bf16) = quickgelu(acc + bias)             x * sigmoid(1.702 x)
  EPI_BIAS_RESID = 2,  // C(bf16) = resid + acc + bias                (C may alias resid)
  EPI_PATCH = 3,       // patch-embed: token-row scatter + positional embedding, C(bf16)
  EPI_F32 = 4,         // C(fp32) = alpha * acc                       (materialised similarity / scores)
  EPI_TARGET = 5,      // tscore_out[row] = acc[row, target[row] - col_offset]
  EPI_COUNT = 6,       // counts[row] += #{col: acc > ts[row]} + #{col: acc == ts[row] and gcol < target[row]}
  EPI_LN_BIAS = 7,        // C(bf16) = rstd_m * (acc - mean_m * colsum_n) + bias_n   == LayerNorm(A) . W^T + b with the
  EPI_LN_BIAS_QGELU = 8,  //   LayerNorm affine folded into B / colsum / bias (see fold_ln_weights); 8 adds QuickGELU
  EPI_F32_SPLITK = 9,     // C(fp32) += alpha * acc over `k_splits` K ranges (atomic adds; the caller zeroes C): weight
                          //   gradients, whose few output tiles would otherwise leave most SM pairs idle
  EPI_LN_BIAS_GELU = 10,  // EPI_LN_BIAS + exact GELU 0.5 x (1 + erf(x / sqrt 2)): the timm vision tower of the SLIP layout
  EPI_QGELU_BWD = 11,     // C(bf16) = acc * quickgelu'(resid): the dgrad of c_proj fused with QuickGELU's backward (resid = the
                          //   pre-activation u the forward kept); no bias
  EPI_NUM = 12,
};

struct GemmParams {
  int M = 0, N = 0, K = 0;
  void* C = nullptr;          // bf16 or fp32 [M, ldc]
  int64_t ldc = 0;
  const float* bias = nullptr;   // [N]
  const bf16* resid = nullptr;   // [M, ldr]
  int64_t ldr = 0;
  float alpha = 1.f;
  int k_splits = 1;              // EPI_F32_SPLITK
  // operand storage: false = K-major ([rows][K contiguous], the default), true = MN-major ([K rows][M or N contiguous],
  // i.e. the operand is the transpose of a row-major matrix and is read in place).  lda / ldb stay the row strides.
  bool a_mn = false, b_mn = false;
  // EPI_LN_*: per-row partial (sum, sumsq) of A's rows, [M, ln_parts, 2] fp32; colsum[n] = sum_k B[n,k]
  const float* ln_stats = nullptr;
  int ln_parts = 0;
  const float* colsum = nullptr;
  float ln_eps = 1e-5f;
  // EPI_BIAS_RESID: optional per-row partial (sum, sumsq) of the OUTPUT rows, [M, N/64, 2] fp32 (feeds the next EPI_LN_*)
  float* stats_out = nullptr;
  // EPI_PATCH
  const float* pos = nullptr;    // [patches_per_frame + 1, N]
  int patches_per_frame = 0;
  // EPI_TARGET / EPI_COUNT
  const int32_t* target = nullptr;       // [M] global column index of each row's target
  const float* target_score = nullptr;   // [M]
  float* tscore_out = nullptr;           // [M]
  int32_t* counts = nullptr;             // [M]
  int col_offset = 0;                    // global index of local column 0
};

// C[M,N] = epilogue(A[M,K] * B[N,K]^T); A and B are bf16, K contiguous (row strides lda / ldb in elements,
// multiples of 8; base pointers 16-byte aligned).  Asynchronous on `stream`.
int gemm_bf16_tn(int epilogue, const bf16* A, int64_t lda, const bf16* B, int64_t ldb, const GemmParams& p,
                 cudaStream_t stream);

}  // namespace fc
