// Rank / metric kernels (K13, K15) on a materialised score matrix.
//   rank_from_scores   : Rank.update, aligner/metrics.py:16-19, as a count instead of an N log N argsort:
//                        rank_i = #{j: s_ij > s_it} + #{j < t: s_ij == s_it}   (stable-descending-sort tie rule)
//   metrics_from_ranks : Recall@1/5/10 (torchmetrics micro top-k, aligner/text_video_retrieval.py:21),
//                        MedianRank = lower median + 1 (aligner/metrics.py:33-36), MeanRank (:27-30)
//   topk_rows          : per-row top-k (value desc, index asc) for predict-style outputs / distributed merges
//   nce_loss, ts_nce_loss : aligner/loss.py:13-39 forward (rows + columns)
// All of these stream the matrix once: 4 bytes per score, HBM-bound.
#include <float.h>

#include "kernels.cuh"

namespace fc {

namespace {

__device__ __forceinline__ int block_sum_int(int v, int* red) {
  v = warp_sum_int(v);
  const int w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[w] = v;
  __syncthreads();
  int t = 0;
  for (int i = 0; i < nw; ++i) t += red[i];
  return t;
}
__device__ __forceinline__ float block_sum_f(float v, float* red) {
  v = warp_sum(v);
  const int w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[w] = v;
  __syncthreads();
  float t = 0.f;
  for (int i = 0; i < nw; ++i) t += red[i];
  return t;
}
__device__ __forceinline__ float block_max_f(float v, float* red) {
  v = warp_max(v);
  const int w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[w] = v;
  __syncthreads();
  float t = -INFINITY;
  for (int i = 0; i < nw; ++i) t = fmaxf(t, red[i]);
  return t;
}

__global__ void __launch_bounds__(256) rank_from_scores_kernel(const float* __restrict__ S, int64_t ld, int64_t cols,
                                                               const int32_t* __restrict__ target,
                                                               int64_t* __restrict__ ranks) {
  __shared__ int red[8];
  const int64_t row = blockIdx.x;
  const float* r = S + row * ld;
  const int t = target[row];
  const float ts = r[t];
  int cnt = 0;
  const bool vec = ((ld & 3) == 0) && ((reinterpret_cast<uintptr_t>(S) & 15) == 0);
  if (vec) {
    const int64_t nv = cols >> 2;
    for (int64_t i = threadIdx.x; i < nv; i += blockDim.x) {
      const uint4 u = ld_nc_v4(reinterpret_cast<const uint4*>(r) + i);
      const float v[4] = {__uint_as_float(u.x), __uint_as_float(u.y), __uint_as_float(u.z), __uint_as_float(u.w)};
      const int j0 = static_cast<int>(i << 2);
#pragma unroll
      for (int e = 0; e < 4; ++e) cnt += (v[e] > ts || (v[e] == ts && j0 + e < t)) ? 1 : 0;
    }
    for (int64_t j = (nv << 2) + threadIdx.x; j < cols; j += blockDim.x)
      cnt += (r[j] > ts || (r[j] == ts && j < t)) ? 1 : 0;
  } else {
    for (int64_t j = threadIdx.x; j < cols; j += blockDim.x) cnt += (r[j] > ts || (r[j] == ts && j < t)) ? 1 : 0;
  }
  cnt = block_sum_int(cnt, red);
  if (threadIdx.x == 0) ranks[row] = cnt;
}

__global__ void counts_to_ranks_kernel(const int32_t* __restrict__ counts, int64_t* __restrict__ ranks, int64_t n) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) ranks[i] = counts[i];
}

// single CTA: recall@{1,5,10}, lower median (+1), mean (+1)
__global__ void __launch_bounds__(1024) metrics_from_ranks_kernel(const int64_t* __restrict__ ranks, int64_t n,
                                                                  int64_t num_candidates,
                                                                  float* __restrict__ out_recall,
                                                                  int64_t* __restrict__ out_median,
                                                                  float* __restrict__ out_mean) {
  __shared__ int red[32];
  __shared__ double dred[32];
  int h1 = 0, h5 = 0, h10 = 0;
  double sum = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
    const int64_t r = ranks[i];
    h1 += r < 1;
    h5 += r < 5;
    h10 += r < 10;
    sum += static_cast<double>(r);
  }
  h1 = block_sum_int(h1, red);
  h5 = block_sum_int(h5, red);
  h10 = block_sum_int(h10, red);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) dred[threadIdx.x >> 5] = sum;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < (blockDim.x >> 5); ++i) t += dred[i];
    out_recall[0] = static_cast<float>(h1) / static_cast<float>(n);
    out_recall[1] = static_cast<float>(h5) / static_cast<float>(n);
    out_recall[2] = static_cast<float>(h10) / static_cast<float>(n);
    if (out_mean) *out_mean = static_cast<float>(t / static_cast<double>(n)) + 1.f;
  }
  // lower median = value at sorted position (n-1)/2: the smallest v with #{rank <= v} >= (n-1)/2 + 1
  const int64_t need = (n - 1) / 2 + 1;
  int64_t lo = 0, hi = num_candidates > 0 ? num_candidates - 1 : 0;
  while (lo < hi) {
    const int64_t mid = lo + ((hi - lo) >> 1);
    int c = 0;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) c += ranks[i] <= mid;
    c = block_sum_int(c, red);
    if (c >= need) hi = mid; else lo = mid + 1;
  }
  if (threadIdx.x == 0) *out_median = lo + 1;
}

// ---- per-row top-k: thread-local sorted lists (one pass over the row), then k rounds of block argmax
constexpr int TOPK_MAX = 16;
__device__ __forceinline__ bool better(float v, int i, float bv, int bi) { return v > bv || (v == bv && i < bi); }

__global__ void __launch_bounds__(256) topk_rows_kernel(const float* __restrict__ S, int64_t ld, int64_t cols, int k,
                                                        float* __restrict__ out_val, int32_t* __restrict__ out_idx) {
  __shared__ float sv[8];
  __shared__ int si[8];
  __shared__ int s_win;
  const int64_t row = blockIdx.x;
  const float* r = S + row * ld;
  float lv[TOPK_MAX];
  int li[TOPK_MAX];
#pragma unroll
  for (int e = 0; e < TOPK_MAX; ++e) {
    lv[e] = -INFINITY;
    li[e] = INT_MAX;
  }
  for (int64_t j = threadIdx.x; j < cols; j += blockDim.x) {
    float v = r[j];
    int idx = static_cast<int>(j);
    if (better(v, idx, lv[TOPK_MAX - 1], li[TOPK_MAX - 1])) {
#pragma unroll
      for (int e = 0; e < TOPK_MAX; ++e) {  // insertion by bubbling the displaced element down
        if (better(v, idx, lv[e], li[e])) {
          const float tv = lv[e];
          const int ti = li[e];
          lv[e] = v;
          li[e] = idx;
          v = tv;
          idx = ti;
        }
      }
    }
  }
  for (int round = 0; round < k; ++round) {
    float bv = lv[0];
    int bi = li[0];
    int who = threadIdx.x;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      const int ow = __shfl_xor_sync(0xffffffffu, who, o);
      if (better(ov, oi, bv, bi)) {
        bv = ov;
        bi = oi;
        who = ow;
      }
    }
    __shared__ int sw_[8];
    if ((threadIdx.x & 31) == 0) {
      sv[threadIdx.x >> 5] = bv;
      si[threadIdx.x >> 5] = bi;
      sw_[threadIdx.x >> 5] = who;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      float fv = sv[0];
      int fi = si[0], fw = sw_[0];
      for (int w = 1; w < (blockDim.x >> 5); ++w)
        if (better(sv[w], si[w], fv, fi)) {
          fv = sv[w];
          fi = si[w];
          fw = sw_[w];
        }
      out_val[row * k + round] = fv;
      out_idx[row * k + round] = fi == INT_MAX ? -1 : fi;
      s_win = fw;
    }
    __syncthreads();
    if (threadIdx.x == s_win) {  // pop the winner's head
#pragma unroll
      for (int e = 0; e < TOPK_MAX - 1; ++e) {
        lv[e] = lv[e + 1];
        li[e] = li[e + 1];
      }
      lv[TOPK_MAX - 1] = -INFINITY;
      li[TOPK_MAX - 1] = INT_MAX;
    }
    __syncthreads();
  }
}

// ---- losses. CTA i < B handles row i, CTA B + i handles column i; terms[] are reduced in fixed order by one CTA.
__device__ __forceinline__ float lse_line(const float* p, int64_t stride, int B, float* red) {
  float mx = -INFINITY;
  for (int j = threadIdx.x; j < B; j += blockDim.x) mx = fmaxf(mx, p[j * stride]);
  mx = block_max_f(mx, red);
  float se = 0.f;
  for (int j = threadIdx.x; j < B; j += blockDim.x) se += expf(p[j * stride] - mx);
  se = block_sum_f(se, red);
  return logf(se) + mx;
}

__global__ void __launch_bounds__(256) nce_terms_kernel(const float* __restrict__ S, int64_t ld, int B,
                                                        float* __restrict__ terms) {
  __shared__ float red[8];
  const int i = blockIdx.x % B;
  const bool is_col = blockIdx.x >= B;
  const float* p = is_col ? S + i : S + static_cast<int64_t>(i) * ld;
  const int64_t stride = is_col ? ld : 1;
  const float lse = lse_line(p, stride, B, red);
  if (threadIdx.x == 0) terms[blockIdx.x] = lse - S[static_cast<int64_t>(i) * ld + i];
}

__global__ void __launch_bounds__(256) ts_nce_terms_kernel(const float* __restrict__ S, const float* __restrict__ T,
                                                           int64_t ld, int B, float* __restrict__ terms) {
  __shared__ float red[8];
  const int i = blockIdx.x % B;
  const bool is_col = blockIdx.x >= B;
  const int64_t off = is_col ? i : static_cast<int64_t>(i) * ld;
  const int64_t stride = is_col ? ld : 1;
  const float lse_s = lse_line(S + off, stride, B, red);
  const float lse_t = lse_line(T + off, stride, B, red);
  float kl = 0.f;
  for (int j = threadIdx.x; j < B; j += blockDim.x) {
    const float lt = T[off + j * stride] - lse_t;
    const float ls = S[off + j * stride] - lse_s;
    const float pt = expf(lt);
    kl += pt > 0.f ? pt * (lt - ls) : 0.f;
  }
  kl = block_sum_f(kl, red);
  if (threadIdx.x == 0) terms[blockIdx.x] = kl;
}

// out = (sum terms[0:B] + sum terms[B:2B]) / B  -- "mean" for nce (per direction), "batchmean" for the KL form
__global__ void __launch_bounds__(256) reduce_terms_kernel(const float* __restrict__ terms, int B,
                                                           float* __restrict__ out) {
  __shared__ double red[256];
  double a = 0.0;
  for (int j = threadIdx.x; j < 2 * B; j += blockDim.x) a += static_cast<double>(terms[j]);
  red[threadIdx.x] = a;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) *out = static_cast<float>(red[0] / static_cast<double>(B));
}

}  // namespace

int rank_from_scores(const float* S, int64_t ld, int64_t rows, int64_t cols, const int32_t* target, int64_t* ranks,
                     cudaStream_t s) {
  FC_REQUIRE(S && target && ranks, "rank_from_scores: null pointer");
  FC_REQUIRE(cols >= 1 && ld >= cols && cols < (int64_t(1) << 31), "rank_from_scores: bad shape");
  if (rows == 0) return FC_OK;
  rank_from_scores_kernel<<<static_cast<unsigned>(rows), 256, 0, s>>>(S, ld, cols, target, ranks);
  FC_CHECK_LAUNCH();
  return FC_OK;
}

int counts_to_ranks(const int32_t* counts, int64_t* ranks, int64_t n, cudaStream_t s) {
  if (n == 0) return FC_OK;
  counts_to_ranks_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, s>>>(counts, ranks, n);
  FC_CHECK_LAUNCH();
  return FC_OK;
}

int metrics_from_ranks(const int64_t* ranks, int64_t n, int64_t num_candidates, float* out_recall3,
                       int64_t* out_median, float* out_mean, cudaStream_t s) {
  FC_REQUIRE(ranks && out_recall3 && out_median, "metrics_from_ranks: null pointer");
  FC_REQUIRE(n >= 1, "metrics_from_ranks: needs at least one rank (torch.median of an empty tensor raises)");
  metrics_from_ranks_kernel<<<1, 1024, 0, s>>>(ranks, n, num_candidates, out_recall3, out_median, out_mean);
  FC_CHECK_LAUNCH();
  return FC_OK;
}

int topk_rows(const float* S, int64_t ld, int64_t rows, int64_t cols, int k, float* out_val, int32_t* out_idx,
              cudaStream_t s) {
  FC_REQUIRE(S && out_val && out_idx, "topk_rows: null pointer");
  FC_REQUIRE(k >= 1 && k <= TOPK_MAX, "topk_rows: k=%d must be in 1..%d", k, TOPK_MAX);
  if (rows == 0) return FC_OK;
  topk_rows_kernel<<<static_cast<unsigned>(rows), 256, 0, s>>>(S, ld, cols, k, out_val, out_idx);
  FC_CHECK_LAUNCH();
  return FC_OK;
}

int nce_loss(const float* S, int64_t ld, int B, float* workspace, float* out, cudaStream_t s) {
  FC_REQUIRE(S && workspace && out && B >= 1, "nce_loss: bad arguments");
  nce_terms_kernel<<<2 * B, 256, 0, s>>>(S, ld, B, workspace);
  FC_CHECK_LAUNCH();
  reduce_terms_kernel<<<1, 256, 0, s>>>(workspace, B, out);
  FC_CHECK_LAUNCH();
  return FC_OK;
}

int ts_nce_loss(const float* S, const float* Tt, int64_t ld, int B, float* workspace, float* out, cudaStream_t s) {
  FC_REQUIRE(S && Tt && workspace && out && B >= 1, "ts_nce_loss: bad arguments");
  ts_nce_terms_kernel<<<2 * B, 256, 0, s>>>(S, Tt, ld, B, workspace);
  FC_CHECK_LAUNCH();
  reduce_terms_kernel<<<1, 256, 0, s>>>(workspace, B, out);
  FC_CHECK_LAUNCH();
  return FC_OK;
}

}  // namespace fc
