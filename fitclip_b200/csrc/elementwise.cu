// Memory-bound kernels of the hot path: 128-bit vectorised, coalesced HBM access with warp-shuffle reductions.
//   K2  layernorm_bf16           (ln_pre / ln_1 / ln_2; reference twin aligner/encoder/slip.py:350-356)
//   K1a im2col_patches + cls row (feeds the patch-embed GEMM; [3P] VisionTransformer.forward)
//   K9  text_embed               (token_embedding(text) + positional_embedding; slip.py:469-470)
//   K7/K10 head                  (ln_post(x[:,0]) @ proj  /  ln_final(x)[eot] @ text_projection; slip.py:475-478)
//   K8  pool_normalize           (x / ||x|| per frame, mean over frames; clip_video_text_encoder.py:85-89,93-94)
//   K14 wise_lerp                ((1-w) p1 + w p2, no FMA contraction; aligner/wise.py:16)
#include "kernels.cuh"

namespace fc {

namespace {

// ------------------------------------------------------------------------------------------- LayerNorm (K2)
// One warp per row; the row lives in registers (D <= 1024): one HBM read, one HBM write. fp32 statistics, eps inside
// the sqrt, two-pass (mean, then centred variance) like ATen's CPU kernel within rounding.
// Warps walk rows with a grid stride; gamma / beta come from shared memory.  Rows reach a warp through a three-slot
// shared-memory ring filled by cp.async (L2 -> shared, no registers): while row r is normalised, rows r + 1 and r + 2 are
// in flight.  (Round 1: one row per warp, 0.71 of the copy peak; early round 2: two rows prefetched into registers, which
// cost an SM a block: 0.71 / 0.79 at 512 / 2000 frames.)
constexpr int LN_MAX_CHUNKS = 4;  // 4 x 32 lanes x 8 bf16 = 1024
constexpr int LN_WARPS = 8;
constexpr int LN_SLOTS = 3;
constexpr int ln_ring_bytes(int ch) { return LN_WARPS * LN_SLOTS * ch * 32 * 16; }

__device__ __forceinline__ void ln_cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst))),
               "l"(gsrc)
               : "memory");
}

template <int CH>
__global__ void __launch_bounds__(LN_WARPS * 32) layernorm_bf16_kernel(const bf16* x, bf16* y,  // may alias (ln_pre runs in place)
                                                                     const float* __restrict__ gamma,
                                                                     const float* __restrict__ beta, int64_t rows, int D,
                                                                     int64_t ldx, int64_t ldy, float eps,
                                                                     float* __restrict__ stats_out) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int chunks = D >> 3;
  const float inv_d = 1.f / static_cast<float>(D);
  extern __shared__ __align__(16) uint4 ln_dyn[];  // [warp][slot][16-byte chunk]
  uint4(*ring)[LN_SLOTS][CH * 32] = reinterpret_cast<uint4(*)[LN_SLOTS][CH * 32]>(ln_dyn);
  __shared__ __align__(16) float s_gamma[CH * 256], s_beta[CH * 256];
  for (int k = threadIdx.x; k < CH * 256; k += LN_WARPS * 32) {
    s_gamma[k] = k < D ? gamma[k] : 0.f;
    s_beta[k] = k < D ? beta[k] : 0.f;
  }
  __syncthreads();
  const int64_t nwarps = static_cast<int64_t>(gridDim.x) * LN_WARPS;
  int64_t row = static_cast<int64_t>(blockIdx.x) * LN_WARPS + warp;
  // every lane copies, and later reads, only its own chunks; rows ahead are only read (y may alias x: this warp is the one
  // that writes them, later)
  auto issue = [&](int64_t r, int slot) {
    if (r < rows) {
      const uint4* xr = reinterpret_cast<const uint4*>(x + r * ldx);
#pragma unroll
      for (int i = 0; i < CH; ++i)
        if (lane + 32 * i < chunks) ln_cp_async16(&ring[warp][slot][lane + 32 * i], xr + lane + 32 * i);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");  // an (empty) group per call keeps the wait arithmetic uniform
  };
  issue(row, 0);
  issue(row + nwarps, 1);
  for (int slot = 0; row < rows; row += nwarps, slot = slot == LN_SLOTS - 1 ? 0 : slot + 1) {
    issue(row + 2 * nwarps, slot >= 1 ? slot - 1 : LN_SLOTS - 1);  // the slot freed one iteration ago
    asm volatile("cp.async.wait_group 2;" ::: "memory");           // all but the two newest groups: this row has landed
    float v[CH][8];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < CH; ++i) {
      if (lane + 32 * i < chunks) {
        const uint4 q4 = ring[warp][slot][lane + 32 * i];
        const uint32_t w[4] = {q4.x, q4.y, q4.z, q4.w};
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const float2 f = unpack_bf16x2(w[t]);
          v[i][2 * t] = f.x;
          v[i][2 * t + 1] = f.y;
          sum += f.x + f.y;
        }
      }
    }
    const float mean = warp_sum(sum) * inv_d;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < CH; ++i) {
      if (lane + 32 * i < chunks) {
#pragma unroll
        for (int t = 0; t < 8; ++t) {
          const float d = v[i][t] - mean;
          sq += d * d;
        }
      }
    }
    const float rstd = rsqrtf(warp_sum(sq) * inv_d + eps);
    uint4* yr = reinterpret_cast<uint4*>(y + row * ldy);
    float o1 = 0.f, o2 = 0.f;  // sum / sum of squares of the OUTPUT row (consumed by a folded-LayerNorm GEMM)
#pragma unroll
    for (int i = 0; i < CH; ++i) {
      const int c = lane + 32 * i;
      if (c < chunks) {
        const float4 g0 = *reinterpret_cast<const float4*>(s_gamma + c * 8), g1 = *reinterpret_cast<const float4*>(s_gamma + c * 8 + 4);
        const float4 b0 = *reinterpret_cast<const float4*>(s_beta + c * 8), b1 = *reinterpret_cast<const float4*>(s_beta + c * 8 + 4);
        const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
        const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
        float r[8];
#pragma unroll
        for (int t = 0; t < 8; ++t) {
          r[t] = (v[i][t] - mean) * rstd * gg[t] + bb[t];
          o1 += r[t];
          o2 = fmaf(r[t], r[t], o2);
        }
        uint4 o;
        o.x = pack_bf16x2(r[0], r[1]);
        o.y = pack_bf16x2(r[2], r[3]);
        o.z = pack_bf16x2(r[4], r[5]);
        o.w = pack_bf16x2(r[6], r[7]);
        st_na_v4(yr + c, o);
      }
    }
    if (stats_out != nullptr) {  // layout [rows, D/64, 2]: everything in part 0, zeros elsewhere
      o1 = warp_sum(o1);
      o2 = warp_sum(o2);
      const int parts = D >> 6;
      float2* so = reinterpret_cast<float2*>(stats_out) + row * parts;
      for (int pi = lane; pi < parts; pi += 32) so[pi] = pi == 0 ? make_float2(o1, o2) : make_float2(0.f, 0.f);
    }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
}

// Row statistics of x alone, in the [rows, D/64, 2] layout the folded-LayerNorm GEMM epilogue reads.  One warp per row.
__global__ void __launch_bounds__(256) row_stats_kernel(const bf16* __restrict__ x, int64_t ldx, int64_t rows, int D,
                                                        float* __restrict__ stats_out) {
  const int lane = threadIdx.x & 31;
  const int chunks = D >> 3, parts = D >> 6;
  const int64_t nwarps = static_cast<int64_t>(gridDim.x) * 8;
  for (int64_t row = static_cast<int64_t>(blockIdx.x) * 8 + (threadIdx.x >> 5); row < rows; row += nwarps) {
    const uint4* xr = reinterpret_cast<const uint4*>(x + row * ldx);
    float o1 = 0.f, o2 = 0.f;
    for (int c = lane; c < chunks; c += 32) {
      const uint4 q4 = ld_stream_v4(xr + c);
      const uint32_t w[4] = {q4.x, q4.y, q4.z, q4.w};
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const float2 f = unpack_bf16x2(w[t]);
        o1 += f.x + f.y;
        o2 = fmaf(f.x, f.x, fmaf(f.y, f.y, o2));
      }
    }
    o1 = warp_sum(o1);
    o2 = warp_sum(o2);
    float2* so = reinterpret_cast<float2*>(stats_out) + row * parts;
    for (int pi = lane; pi < parts; pi += 32) so[pi] = pi == 0 ? make_float2(o1, o2) : make_float2(0.f, 0.f);
  }
}

// ------------------------------------------------------------------------------------------- im2col (K1a)
// frames (F,3,R,R) -> patches (F*G*G, 3*P*P) bf16, column order (c, ky, kx) = conv1.weight.reshape(width, -1).
// One thread moves 8 consecutive kx: a 32-byte (fp32) read and a 16-byte write; consecutive threads walk the image row.
template <typename T>
__device__ __forceinline__ void load8(const T* p, float (&v)[8]);
template <>
__device__ __forceinline__ void load8<float>(const float* p, float (&v)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p));
  const float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
template <>
__device__ __forceinline__ void load8<bf16>(const bf16* p, float (&v)[8]) {
  const uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const float2 f = unpack_bf16x2(w[t]);
    v[2 * t] = f.x;
    v[2 * t + 1] = f.y;
  }
}
template <>
__device__ __forceinline__ void load8<__half>(const __half* p, float (&v)[8]) {
  const uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
  const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const float2 f = __half22float2(h[t]);
    v[2 * t] = f.x;
    v[2 * t + 1] = f.y;
  }
}

template <typename T>
__global__ void __launch_bounds__(256) im2col_kernel(const T* __restrict__ frames, bf16* __restrict__ patches,
                                                     int64_t total_vec, int R, int P) {
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total_vec) return;
  const int vec_per_row = R >> 3;
  const int xv = static_cast<int>(idx % vec_per_row);
  int64_t t = idx / vec_per_row;
  const int y = static_cast<int>(t % R);
  t /= R;
  const int c = static_cast<int>(t % 3);
  const int64_t f = t / 3;
  float v[8];
  load8<T>(frames + ((f * 3 + c) * R + y) * static_cast<int64_t>(R) + xv * 8, v);
  const int G = R / P;
  const int x = xv * 8;
  const int py = y / P, ky = y - py * P, px = x / P, kx = x - px * P;
  const int64_t row = (f * G + py) * G + px;
  const int64_t col = (static_cast<int64_t>(c) * P + ky) * P + kx;
  uint4 o;
  o.x = pack_bf16x2(v[0], v[1]);
  o.y = pack_bf16x2(v[2], v[3]);
  o.z = pack_bf16x2(v[4], v[5]);
  o.w = pack_bf16x2(v[6], v[7]);
  *reinterpret_cast<uint4*>(patches + row * (3 * P * P) + col) = o;
}

// Patch sizes that are not multiples of 8 (ViT-L/14): one thread per patch-matrix element, columns padded with zeros up
// to `ldp` (a multiple of 8, what the GEMM's TMA loads need).
template <typename T>
__global__ void __launch_bounds__(256) im2col_generic_kernel(const T* __restrict__ frames, bf16* __restrict__ patches,
                                                             int64_t total, int R, int P, int ldp) {
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int col = static_cast<int>(idx % ldp);
  const int64_t row = idx / ldp;
  float v = 0.f;
  if (col < 3 * P * P) {
    const int G = R / P;
    const int px = static_cast<int>(row % G);
    const int py = static_cast<int>((row / G) % G);
    const int64_t f = row / (static_cast<int64_t>(G) * G);
    const int kx = col % P, ky = (col / P) % P, c = col / (P * P);
    v = static_cast<float>(frames[((f * 3 + c) * R + py * P + ky) * static_cast<int64_t>(R) + px * P + kx]);
  }
  patches[idx] = __float2bfloat16(v);
}

// class-token row: x[f*L + 0, :] = class_embedding + positional_embedding[0]
__global__ void cls_row_kernel(bf16* __restrict__ x, const float* __restrict__ cls, const float* __restrict__ pos,
                               int64_t frames, int L, int D) {
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= frames * D) return;
  const int64_t f = idx / D;
  const int d = static_cast<int>(idx - f * D);
  x[f * L * D + d] = __float2bfloat16(cls[d] + pos[d]);
}

// ------------------------------------------------------------------------------------------- text embed (K9)
__global__ void __launch_bounds__(256) text_embed_kernel(const int32_t* __restrict__ ids,
                                                         const float* __restrict__ tok, const float* __restrict__ pos,
                                                         bf16* __restrict__ x, int64_t tokens, int L, int D,
                                                         int vocab, int* __restrict__ err_flag,
                                                         float* __restrict__ stats_out) {
  const int vec_per_row = D >> 3;
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= tokens * vec_per_row) return;  // whole 8-thread groups drop out together (vec_per_row % 8 == 0)
  const int64_t row = idx / vec_per_row;
  const int dv = static_cast<int>(idx - row * vec_per_row) * 8;
  int id = ids[row];
  if (id < 0 || id >= vocab) {  // torch would raise an index error; flag it and clamp so we never read out of bounds
    atomicExch(err_flag, 1);
    id = 0;
  }
  const int l = static_cast<int>(row % L);
  float a[8], b[8];
  load8<float>(tok + static_cast<int64_t>(id) * D + dv, a);
  load8<float>(pos + static_cast<int64_t>(l) * D + dv, b);
  uint4 o;
  o.x = pack_bf16x2(a[0] + b[0], a[1] + b[1]);
  o.y = pack_bf16x2(a[2] + b[2], a[3] + b[3]);
  o.z = pack_bf16x2(a[4] + b[4], a[5] + b[5]);
  o.w = pack_bf16x2(a[6] + b[6], a[7] + b[7]);
  *reinterpret_cast<uint4*>(x + row * D + dv) = o;
  if (stats_out != nullptr) {
    // 8 consecutive threads cover one 64-column part of a row (D % 64 == 0 keeps them in one warp and one row)
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      const float r = a[t] + b[t];
      s1 += r;
      s2 = fmaf(r, r, s2);
    }
#pragma unroll
    for (int o2 = 4; o2 > 0; o2 >>= 1) {
      s1 += __shfl_xor_sync(0xffffffffu, s1, o2);
      s2 += __shfl_xor_sync(0xffffffffu, s2, o2);
    }
    if ((threadIdx.x & 7) == 0)
      reinterpret_cast<float2*>(stats_out)[row * (D >> 6) + (dv >> 6)] = make_float2(s1, s2);
  }
}

// Fold a LayerNorm (gamma, beta) into the Linear that consumes it:  LN(x) W^T + b  =  rstd (x W'^T - mean colsum) + b'
//   W'[n,k] = bf16(gamma[k] W[n,k]),  colsum[n] = sum_k float(W'[n,k]),  b'[n] = b[n] + sum_k beta[k] W[n,k].
// One warp per output feature n.
__global__ void __launch_bounds__(256) fold_ln_kernel(const float* __restrict__ W, const float* __restrict__ gamma,
                                                      const float* __restrict__ beta, const float* __restrict__ bias,
                                                      bf16* __restrict__ Wf, float* __restrict__ colsum,
                                                      float* __restrict__ bias_f, int N, int K) {
  const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (n >= N) return;
  const int lane = threadIdx.x & 31;
  float cs = 0.f, bs = 0.f;
  for (int k = lane; k < K; k += 32) {
    const float w = W[static_cast<int64_t>(n) * K + k];
    const bf16 wf = __float2bfloat16(gamma[k] * w);
    Wf[static_cast<int64_t>(n) * K + k] = wf;
    cs += __bfloat162float(wf);
    bs = fmaf(beta[k], w, bs);
  }
  cs = warp_sum(cs);
  bs = warp_sum(bs);
  if (lane == 0) {
    colsum[n] = cs;
    bias_f[n] = bias[n] + bs;
  }
}

// ------------------------------------------------------------------------------------------- heads (K7 / K10)
// A CTA takes HEAD_SEQS sequences x HEAD_COLS output columns.  Phase 1: one WARP per sequence picks the pooled token
// row (class token, or the EOT position = first argmax of the ids) and LayerNorms it in fp32 into shared memory.
// Phase 2: the K dimension of the fp32 projection (W, E) is split over HEAD_KG thread groups so that 16 warps per CTA
// hide the L2 latency of the weight loads (the kernel is latency-, not bandwidth-bound); every weight feeds HEAD_SEQS
// FMAs.  Phase 3: the HEAD_KG partial sums are added in a fixed order and the un-normalised embedding is written.
constexpr int HEAD_SEQS = 8;
constexpr int HEAD_COLS = 128;  // output columns per CTA (grid.y tiles the embedding dimension)
constexpr int HEAD_KG = 4;      // K groups
constexpr int HEAD_THREADS = HEAD_COLS * HEAD_KG;
static_assert(HEAD_SEQS <= HEAD_THREADS / 32, "one warp per sequence in phase 1");

__global__ void __launch_bounds__(HEAD_THREADS) head_kernel(const bf16* __restrict__ x, const int32_t* __restrict__ ids,
                                                            const float* __restrict__ gamma, const float* __restrict__ beta,
                                                            const float* __restrict__ proj, float* __restrict__ out,
                                                            int64_t seqs, int L, int W, int E, float eps) {
  extern __shared__ float sh[];  // HEAD_SEQS x W normalised values, then HEAD_KG x HEAD_SEQS x HEAD_COLS partial sums
  float* part = sh + HEAD_SEQS * W;
  const int64_t seq0 = static_cast<int64_t>(blockIdx.x) * HEAD_SEQS;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp < HEAD_SEQS) {
    float* y = sh + warp * W;
    const int64_t seq = seq0 + warp;
    if (seq >= seqs) {  // warp-uniform
      for (int k = lane; k < W; k += 32) y[k] = 0.f;
    } else {
      int pos = 0;
      if (ids != nullptr) {  // first index of the maximum id (torch.argmax semantics)
        int best = INT_MIN, best_i = 0;
        for (int l = lane; l < L; l += 32) {
          const int v = ids[seq * L + l];
          if (v > best) {
            best = v;
            best_i = l;
          }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const int ob = __shfl_xor_sync(0xffffffffu, best, o);
          const int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
          if (ob > best || (ob == best && oi < best_i)) {
            best = ob;
            best_i = oi;
          }
        }
        pos = best_i;
      }
      const bf16* row = x + (seq * L + pos) * static_cast<int64_t>(W);
      float sum = 0.f;
      for (int k = lane; k < W; k += 32) {
        const float v = __bfloat162float(row[k]);
        y[k] = v;
        sum += v;
      }
      const float mean = warp_sum(sum) / static_cast<float>(W);
      float sq = 0.f;
      for (int k = lane; k < W; k += 32) {
        const float d = y[k] - mean;
        sq += d * d;
      }
      const float rstd = rsqrtf(warp_sum(sq) / static_cast<float>(W) + eps);
      for (int k = lane; k < W; k += 32) y[k] = (y[k] - mean) * rstd * gamma[k] + beta[k];
    }
  }
  __syncthreads();
  const int col = threadIdx.x % HEAD_COLS, kg = threadIdx.x / HEAD_COLS;
  const int n = blockIdx.y * HEAD_COLS + col;
  float acc[HEAD_SEQS];
#pragma unroll
  for (int sq = 0; sq < HEAD_SEQS; ++sq) acc[sq] = 0.f;
  if (n < E) {
    const int kq = W / HEAD_KG;  // W is a multiple of 64: a whole number of 4-wide steps per group
    const int k0 = kg * kq;
    // four k per step: 16 weight loads in flight (4-way unroll), one 128-bit shared-memory read per sequence
#pragma unroll 4
    for (int k = k0; k < k0 + kq; k += 4) {
      const float w0 = __ldg(proj + static_cast<int64_t>(k) * E + n);
      const float w1 = __ldg(proj + static_cast<int64_t>(k + 1) * E + n);
      const float w2 = __ldg(proj + static_cast<int64_t>(k + 2) * E + n);
      const float w3 = __ldg(proj + static_cast<int64_t>(k + 3) * E + n);
#pragma unroll
      for (int sq = 0; sq < HEAD_SEQS; ++sq) {
        const float4 yv = *reinterpret_cast<const float4*>(sh + sq * W + k);
        acc[sq] = fmaf(yv.w, w3, fmaf(yv.z, w2, fmaf(yv.y, w1, fmaf(yv.x, w0, acc[sq]))));
      }
    }
  }
#pragma unroll
  for (int sq = 0; sq < HEAD_SEQS; ++sq) part[(kg * HEAD_SEQS + sq) * HEAD_COLS + col] = acc[sq];
  __syncthreads();
  for (int i = threadIdx.x; i < HEAD_SEQS * HEAD_COLS; i += HEAD_THREADS) {
    const int sq = i / HEAD_COLS, c = i % HEAD_COLS;
    const int nn = blockIdx.y * HEAD_COLS + c;
    if (seq0 + sq < seqs && nn < E) {
      float t = 0.f;
#pragma unroll
      for (int g = 0; g < HEAD_KG; ++g) t += part[(g * HEAD_SEQS + sq) * HEAD_COLS + c];
      out[(seq0 + sq) * E + nn] = t;
    }
  }
}

// ------------------------------------------------------------------------------------------- pool + normalise (K8)
// out[b,:] = scale * mean_t( x[b*T+t,:] / ||x[b*T+t,:]||_2 ).  One CTA per output row, one warp per frame slot.
__global__ void __launch_bounds__(256) pool_normalize_kernel(const float* __restrict__ x, float* __restrict__ out,
                                                             bf16* __restrict__ out_bf16, int T, int D, float scale) {
  extern __shared__ float sh[];  // T inverse norms
  const int64_t b = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int t = warp; t < T; t += nw) {
    const float* r = x + (b * T + t) * static_cast<int64_t>(D);
    float ss = 0.f;
    for (int d = lane; d < D; d += 32) ss += r[d] * r[d];
    ss = warp_sum(ss);
    if (lane == 0) sh[t] = sqrtf(ss);
  }
  __syncthreads();
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    float acc = 0.f;
    for (int t = 0; t < T; ++t) acc += x[(b * T + t) * static_cast<int64_t>(D) + d] / sh[t];
    const float v = (T == 1 ? acc : acc / static_cast<float>(T)) * scale;
    out[b * D + d] = v;
    if (out_bf16) out_bf16[b * D + d] = __float2bfloat16(v);
  }
}

// Fast path for D = NV * 128 <= 1024 (CLIP: 512, 768): one WARP per output row, the frame lives in registers, every
// input element is read from HBM exactly once with 128-bit streaming loads; the next frame is already in flight while
// the current one is reduced.  Same arithmetic as above (x / ||x|| per element, frames summed in order).
template <int NV>
__global__ void __launch_bounds__(256) pool_normalize_warp_kernel(const float* __restrict__ x, float* __restrict__ out,
                                                                  bf16* __restrict__ out_bf16, int64_t B, int T,
                                                                  float scale) {
  constexpr int D = NV * 128;
  const int64_t b = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= B) return;
  const int lane = threadIdx.x & 31;
  const uint4* r = reinterpret_cast<const uint4*>(x + b * T * static_cast<int64_t>(D)) + lane;
  float acc[NV][4];
#pragma unroll
  for (int i = 0; i < NV; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
  uint4 cur[NV], nxt[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) cur[i] = ld_nc_v4(r + 32 * i);
  for (int t = 0; t < T; ++t) {
    if (t + 1 < T) {
#pragma unroll
      for (int i = 0; i < NV; ++i) nxt[i] = ld_nc_v4(r + static_cast<int64_t>(t + 1) * (D / 4) + 32 * i);
    }
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const float v0 = __uint_as_float(cur[i].x), v1 = __uint_as_float(cur[i].y), v2 = __uint_as_float(cur[i].z),
                  v3 = __uint_as_float(cur[i].w);
      ss += v0 * v0 + v1 * v1 + v2 * v2 + v3 * v3;
    }
    const float nrm = sqrtf(warp_sum(ss));
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      acc[i][0] += __uint_as_float(cur[i].x) / nrm;
      acc[i][1] += __uint_as_float(cur[i].y) / nrm;
      acc[i][2] += __uint_as_float(cur[i].z) / nrm;
      acc[i][3] += __uint_as_float(cur[i].w) / nrm;
      cur[i] = nxt[i];
    }
  }
  const float tf = static_cast<float>(T);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    float4 o;
    o.x = (T == 1 ? acc[i][0] : acc[i][0] / tf) * scale;
    o.y = (T == 1 ? acc[i][1] : acc[i][1] / tf) * scale;
    o.z = (T == 1 ? acc[i][2] : acc[i][2] / tf) * scale;
    o.w = (T == 1 ? acc[i][3] : acc[i][3] / tf) * scale;
    reinterpret_cast<float4*>(out + b * D)[lane + 32 * i] = o;
    if (out_bf16) {
      uint2 pk;
      pk.x = pack_bf16x2(o.x, o.y);
      pk.y = pack_bf16x2(o.z, o.w);
      reinterpret_cast<uint2*>(out_bf16 + b * D)[lane + 32 * i] = pk;
    }
  }
}

// ------------------------------------------------------------------------------------------- WiSE lerp (K14)
// out = c1 * p1 + c2 * p2 with two rounded products and a rounded add -- exactly what torch evaluates for
// `(1 - w) * p1 + w * p2` (aligner/wise.py:16); __fmul_rn/__fadd_rn forbid FMA contraction.
__global__ void __launch_bounds__(256) wise_lerp_kernel(const float* __restrict__ p1, const float* __restrict__ p2,
                                                        float* __restrict__ out, bf16* __restrict__ out_bf16,
                                                        int64_t n, float c1, float c2) {
  const int64_t nvec = n >> 2;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < nvec; i += stride) {
    const uint4 ua = ld_nc_v4(reinterpret_cast<const uint4*>(p1) + i);
    const uint4 ub = ld_nc_v4(reinterpret_cast<const uint4*>(p2) + i);
    float4 r;
    r.x = __fadd_rn(__fmul_rn(c1, __uint_as_float(ua.x)), __fmul_rn(c2, __uint_as_float(ub.x)));
    r.y = __fadd_rn(__fmul_rn(c1, __uint_as_float(ua.y)), __fmul_rn(c2, __uint_as_float(ub.y)));
    r.z = __fadd_rn(__fmul_rn(c1, __uint_as_float(ua.z)), __fmul_rn(c2, __uint_as_float(ub.z)));
    r.w = __fadd_rn(__fmul_rn(c1, __uint_as_float(ua.w)), __fmul_rn(c2, __uint_as_float(ub.w)));
    uint4 o;
    o.x = __float_as_uint(r.x); o.y = __float_as_uint(r.y); o.z = __float_as_uint(r.z); o.w = __float_as_uint(r.w);
    st_na_v4(reinterpret_cast<uint4*>(out) + i, o);
    if (out_bf16) {
      uint2 h;
      h.x = pack_bf16x2(r.x, r.y);
      h.y = pack_bf16x2(r.z, r.w);
      reinterpret_cast<uint2*>(out_bf16)[i] = h;
    }
  }
  // tail (n % 4 elements)
  const int64_t tail0 = nvec << 2;
  const int64_t i = tail0 + static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) {
    const float r = __fadd_rn(__fmul_rn(c1, p1[i]), __fmul_rn(c2, p2[i]));
    out[i] = r;
    if (out_bf16) out_bf16[i] = __float2bfloat16(r);
  }
}

// fp32 -> bf16 weight conversion at load time
// (rows, cols) fp32, contiguous -> (rows, ld) bf16 with zero-filled padding columns (ld == cols: plain conversion)
__global__ void __launch_bounds__(256) f32_to_bf16_kernel(const float* __restrict__ in, bf16* __restrict__ out,
                                                          int64_t rows, int cols, int ld) {
  const int64_t n = rows * ld;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int64_t r = i / ld;
    const int c = static_cast<int>(i - r * ld);
    out[i] = c < cols ? __float2bfloat16(in[r * cols + c]) : __float2bfloat16(0.f);
  }
}

// split an fp32 matrix (rows, D) into bf16 hi / lo parts laid out for the 3-term similarity GEMM:
//   mode 0 (A side): [hi | hi | lo]     mode 1 (B side): [hi | lo | hi]      -> (rows, 3D)
// so that A'.B'^T = hi.hi + hi.lo + lo.hi  ~  fp32 dot product to ~2^-16.   terms == 1 writes [hi] only.
__global__ void __launch_bounds__(256) split_bf16_kernel(const float* __restrict__ in, bf16* __restrict__ out,
                                                         int64_t rows, int D, int mode, int terms) {
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= rows * D) return;
  const int64_t r = idx / D;
  const int d = static_cast<int>(idx - r * D);
  const float v = in[idx];
  const bf16 hi = __float2bfloat16(v);
  if (terms == 1) {
    out[idx] = hi;
    return;
  }
  const bf16 lo = __float2bfloat16(v - __bfloat162float(hi));
  bf16* o = out + r * 3 * D;
  o[d] = hi;
  o[D + d] = mode == 0 ? hi : lo;
  o[2 * D + d] = mode == 0 ? lo : hi;
}

inline int grid_for(int64_t n, int block, int cap_mult = 32) {
  int64_t g = (n + block - 1) / block;
  const int64_t cap = static_cast<int64_t>(num_sms()) * cap_mult;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return static_cast<int>(g);
}

}  // namespace

int row_stats_bf16(const bf16* x, int64_t ldx, int64_t rows, int D, float* stats_out, cudaStream_t s) {
  FC_REQUIRE(x && stats_out && rows > 0 && D % 64 == 0 && ldx % 8 == 0, "row_stats: bad arguments (D=%d)", D);
  const int64_t blocks = (rows + 7) / 8;
  const int grid = static_cast<int>(blocks < 16 * num_sms() ? blocks : 16 * num_sms());
  note_launch();
  row_stats_kernel<<<grid, 256, 0, s>>>(x, ldx, rows, D, stats_out);
  FC_CUDA(cudaGetLastError());
  return FC_OK;
}

int layernorm_bf16(const bf16* x, int64_t ldx, bf16* y, int64_t ldy, const float* gamma, const float* beta,
                   int64_t rows, int D, float eps, float* stats_out, cudaStream_t s) {
  FC_REQUIRE(stats_out == nullptr || D % 64 == 0, "layernorm: stats_out needs D %% 64 == 0");
  FC_REQUIRE(x && y && gamma && beta, "layernorm: null pointer");
  FC_REQUIRE(D % 8 == 0 && D <= 8 * 32 * LN_MAX_CHUNKS && ldx % 8 == 0 && ldy % 8 == 0,
             "layernorm: D=%d must be a multiple of 8 and <= 1024", D);
  if (rows == 0) return FC_OK;
  ProfScope prof(s, PROF_LAYERNORM, 0, rows, D, 0, 0.0, 4.0 * rows * D);
  // persistent-style grid: up to 4 blocks per SM (register use of the CH = 3 instantiation allows 2-3), rows by grid stride
  const int64_t want = (rows + LN_WARPS - 1) / LN_WARPS;
  const unsigned grid = static_cast<unsigned>(std::min<int64_t>(want, int64_t(4) * num_sms()));
  const int ch = (D + 255) / 256;
  static bool configured = false;
  if (!configured) {  // D = 1024: 48 KB of ring + 8 KB of gamma / beta exceed the default 48 KB limit
    FC_CUDA(cudaFuncSetAttribute(layernorm_bf16_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, ln_ring_bytes(4)));
    configured = true;
  }
  if (ch == 1) layernorm_bf16_kernel<1><<<grid, LN_WARPS * 32, ln_ring_bytes(1), s>>>(x, y, gamma, beta, rows, D, ldx, ldy, eps, stats_out);
  else if (ch == 2) layernorm_bf16_kernel<2><<<grid, LN_WARPS * 32, ln_ring_bytes(2), s>>>(x, y, gamma, beta, rows, D, ldx, ldy, eps, stats_out);
  else if (ch == 3) layernorm_bf16_kernel<3><<<grid, LN_WARPS * 32, ln_ring_bytes(3), s>>>(x, y, gamma, beta, rows, D, ldx, ldy, eps, stats_out);
  else layernorm_bf16_kernel<4><<<grid, LN_WARPS * 32, ln_ring_bytes(4), s>>>(x, y, gamma, beta, rows, D, ldx, ldy, eps, stats_out);
  FC_CHECK_LAUNCH();
  return FC_OK;
}

int im2col_patches(const void* frames, int dtype, bf16* patches, int64_t F, int R, int P, cudaStream_t s) {
  FC_REQUIRE(frames && patches, "im2col: null pointer");
  FC_REQUIRE(R % P == 0, "im2col: resolution %d is not a multiple of the patch size %d", R, P);
  if (F == 0) return FC_OK;
  if (P % 8 != 0) {
    const int ldp = (3 * P * P + 7) / 8 * 8;
    const int G = R / P;
    const int64_t total = F * G * G * ldp;
    ProfScope prof(s, PROF_OTHER, 0, F, R, P, 0.0, static_cast<double>(F) * 3 * R * R * ((dtype == FC_DTYPE_F32 ? 4 : 2) + 2));
    const unsigned grid = static_cast<unsigned>((total + 255) / 256);
    if (dtype == FC_DTYPE_F32)
      im2col_generic_kernel<float><<<grid, 256, 0, s>>>(static_cast<const float*>(frames), patches, total, R, P, ldp);
    else if (dtype == FC_DTYPE_BF16)
      im2col_generic_kernel<bf16><<<grid, 256, 0, s>>>(static_cast<const bf16*>(frames), patches, total, R, P, ldp);
    else if (dtype == FC_DTYPE_F16)
      im2col_generic_kernel<__half><<<grid, 256, 0, s>>>(static_cast<const __half*>(frames), patches, total, R, P, ldp);
    else
      FC_REQUIRE(false, "im2col: unsupported frame dtype %d", dtype);
    FC_CHECK_LAUNCH();
    return FC_OK;
  }
  ProfScope prof(s, PROF_OTHER, 0, F, R, P, 0.0, static_cast<double>(F) * 3 * R * R * ((dtype == FC_DTYPE_F32 ? 4 : 2) + 2));
  const int64_t total = F * 3 * R * (R / 8);
  const unsigned grid = static_cast<unsigned>((total + 255) / 256);
  if (dtype == FC_DTYPE_F32)
    im2col_kernel<float><<<grid, 256, 0, s>>>(static_cast<const float*>(frames), patches, total, R, P);
  else if (dtype == FC_DTYPE_BF16)
    im2col_kernel<bf16><<<grid, 256, 0, s>>>(static_cast<const bf16*>(frames), patches, total, R, P);
  else if (dtype == FC_DTYPE_F16)
    im2col_kernel<__half><<<grid, 256, 0, s>>>(static_cast<const __half*>(frames), patches, total, R, P);
  else
    FC_REQUIRE(false, "im2col: unsupported frame dtype %d", dtype);
  FC_CHECK_LAUNCH();
  return FC_OK;
}

int cls_rows(bf16* x, const float* cls, const float* pos, int64_t F, int L, int D, cudaStream_t s) {
  if (F == 0) return FC_OK;
  cls_row_kernel<<<static_cast<unsigned>((F * D + 255) / 256), 256, 0, s>>>(x, cls, pos, F, L, D);
  FC_CHECK_LAUNCH();
  return FC_OK;
}

int text_embed(const int32_t* ids, const float* tok, const float* pos, bf16* x, int64_t C, int L, int D, int vocab,
               int* err_flag, float* stats_out, cudaStream_t s) {
  FC_REQUIRE(D % 64 == 0, "text_embed: width must be a multiple of 64");
  if (C == 0) return FC_OK;
  ProfScope prof(s, PROF_OTHER, 2, C, L, D, 0.0, static_cast<double>(C) * L * D * (4 + 2));
  const int64_t total = C * L * (D / 8);
  text_embed_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, s>>>(ids, tok, pos, x, C * L, L, D, vocab,
                                                                              err_flag, stats_out);
  FC_CHECK_LAUNCH();
  return FC_OK;
}

int head_project(const bf16* x, const int32_t* ids, const float* gamma, const float* beta, const float* proj,
                 float* out, int64_t seqs, int L, int W, int E, float eps, cudaStream_t s) {
  if (seqs == 0) return FC_OK;
  ProfScope prof(s, PROF_OTHER, 3, seqs, W, E, 2.0 * seqs * W * E, 4.0 * seqs * W * E);
  const size_t smem = (HEAD_SEQS * W + HEAD_KG * HEAD_SEQS * HEAD_COLS) * sizeof(float);
  const dim3 grid(static_cast<unsigned>((seqs + HEAD_SEQS - 1) / HEAD_SEQS), (E + HEAD_COLS - 1) / HEAD_COLS);
  head_kernel<<<grid, HEAD_THREADS, smem, s>>>(x, ids, gamma, beta, proj, out, seqs, L, W, E, eps);
  FC_CHECK_LAUNCH();
  return FC_OK;
}

int pool_normalize(const float* x, float* out, bf16* out_bf16, int64_t B, int T, int D, float scale, cudaStream_t s) {
  FC_REQUIRE(x && out, "pool_normalize: null pointer");
  FC_REQUIRE(T >= 1 && T <= 4096 && D >= 1, "pool_normalize: bad T=%d D=%d", T, D);
  if (B == 0) return FC_OK;
  ProfScope prof(s, PROF_OTHER, 4, B, T, D, 0.0, 4.0 * B * D * (2.0 * T + 1));
  const bool aligned = (reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0 &&
                       (out_bf16 == nullptr || (reinterpret_cast<uintptr_t>(out_bf16) & 7) == 0);
  const unsigned wgrid = static_cast<unsigned>((B + 7) / 8);
  if (aligned && D % 128 == 0 && D <= 1024) {
    switch (D / 128) {
      case 1: pool_normalize_warp_kernel<1><<<wgrid, 256, 0, s>>>(x, out, out_bf16, B, T, scale); break;
      case 2: pool_normalize_warp_kernel<2><<<wgrid, 256, 0, s>>>(x, out, out_bf16, B, T, scale); break;
      case 3: pool_normalize_warp_kernel<3><<<wgrid, 256, 0, s>>>(x, out, out_bf16, B, T, scale); break;
      case 4: pool_normalize_warp_kernel<4><<<wgrid, 256, 0, s>>>(x, out, out_bf16, B, T, scale); break;
      case 5: pool_normalize_warp_kernel<5><<<wgrid, 256, 0, s>>>(x, out, out_bf16, B, T, scale); break;
      case 6: pool_normalize_warp_kernel<6><<<wgrid, 256, 0, s>>>(x, out, out_bf16, B, T, scale); break;
      case 7: pool_normalize_warp_kernel<7><<<wgrid, 256, 0, s>>>(x, out, out_bf16, B, T, scale); break;
      default: pool_normalize_warp_kernel<8><<<wgrid, 256, 0, s>>>(x, out, out_bf16, B, T, scale); break;
    }
  } else {
    pool_normalize_kernel<<<static_cast<unsigned>(B), 256, T * sizeof(float), s>>>(x, out, out_bf16, T, D, scale);
  }
  FC_CHECK_LAUNCH();
  return FC_OK;
}

int wise_lerp(const float* p1, const float* p2, float* out, bf16* out_bf16, int64_t n, double w, cudaStream_t s) {
  FC_REQUIRE(p1 && p2 && out, "wise_lerp: null pointer");
  FC_REQUIRE(((reinterpret_cast<uintptr_t>(p1) | reinterpret_cast<uintptr_t>(p2) | reinterpret_cast<uintptr_t>(out)) &
              15) == 0,
             "wise_lerp: pointers must be 16-byte aligned");
  if (n == 0) return FC_OK;
  // python evaluates (1 - w) in double; torch then multiplies the fp32 tensor by the scalar cast to fp32
  const float c1 = static_cast<float>(1.0 - w), c2 = static_cast<float>(w);
  ProfScope prof(s, PROF_OTHER, 5, n, 0, 0, 3.0 * n, 12.0 * n + (out_bf16 ? 2.0 * n : 0.0));
  wise_lerp_kernel<<<grid_for((n + 3) / 4, 256, 8), 256, 0, s>>>(p1, p2, out, out_bf16, n, c1, c2);
  FC_CHECK_LAUNCH();
  return FC_OK;
}

int fold_ln_weights(const float* W, const float* gamma, const float* beta, const float* bias, bf16* Wf, float* colsum,
                    float* bias_f, int N, int K, cudaStream_t s) {
  fold_ln_kernel<<<(N + 7) / 8, 256, 0, s>>>(W, gamma, beta, bias, Wf, colsum, bias_f, N, K);
  FC_CHECK_LAUNCH();
  return FC_OK;
}

int f32_to_bf16(const float* in, bf16* out, int64_t n, cudaStream_t s) {
  if (n == 0) return FC_OK;
  f32_to_bf16_kernel<<<grid_for(n, 256), 256, 0, s>>>(in, out, 1, static_cast<int>(n), static_cast<int>(n));
  FC_CHECK_LAUNCH();
  return FC_OK;
}

int f32_to_bf16_padded(const float* in, bf16* out, int64_t rows, int cols, int ld, cudaStream_t s) {
  if (rows == 0) return FC_OK;
  f32_to_bf16_kernel<<<grid_for(rows * ld, 256), 256, 0, s>>>(in, out, rows, cols, ld);
  FC_CHECK_LAUNCH();
  return FC_OK;
}

int split_bf16(const float* in, bf16* out, int64_t rows, int D, int mode, int terms, cudaStream_t s) {
  if (rows == 0) return FC_OK;
  split_bf16_kernel<<<static_cast<unsigned>((rows * D + 255) / 256), 256, 0, s>>>(in, out, rows, D, mode, terms);
  FC_CHECK_LAUNCH();
  return FC_OK;
}

}  // namespace fc
