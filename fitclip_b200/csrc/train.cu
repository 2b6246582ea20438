// Training-step kernels: the backward of the student towers and of the teacher-student losses,
// and the optimizer.  Reference: TeacherStudentLightningModule.training_step / _dataset_step_end
// (aligner/teacher_student.py:99-183), the losses (aligner/loss.py:13-39), torch.optim.AdamW (config/trainer.yaml:22-24).
// The reference gets all of this from torch.autograd; here every gradient is an explicit kernel:
//
//   dense contractions  dgrad  dX = dY . W          -> the tcgen05 GEMM (gemm.cu) on a transposed bf16 copy of W
//                       wgrad  dW = dY^T . X        -> the tcgen05 GEMM's split-K epilogue (EPI_F32_SPLITK) on
//                                                      transposed activations (transpose_bf16 below, which also
//                                                      produces the bias gradient = column sums of dY on the way)
//   attention backward  L <= 208: tcgen05 kernel in attention_bwd_tc.cu; longer sequences: mma.sync.m16n8k16 here, Q/K/V/dO
//                       of one (sequence, head) resident in shared memory, two phases (query-tile owners produce dQ,
//                       key-tile owners produce dK/dV) so no atomics are needed
//   memory-bound        LayerNorm backward (+ residual add), QuickGELU forward/backward, pool+normalise backward,
//                       embedding scatter/sums, loss gradients, fused AdamW over ONE flat parameter buffer.
#include <math.h>
#include <stdlib.h>

#include "../../include/fitclip_b200.h"
#include "gemm.cuh"
#include "kernels.cuh"
#include "ptx.cuh"

namespace fc {

namespace {

// ------------------------------------------------------------------------------------------- transpose (+ column sums)
// out[c, r'] = in[row(r'), c] for the KEPT input rows r' = 0..kept-1.  Input rows come in groups of `group_len` rows
// whose first `group_skip` rows are dropped (the class-token row of every frame when the patch rows are wanted);
// group_len = 0: every row is kept.  Columns [kept, ld_out) of `out` are zero-filled (the GEMM's K must be a multiple
// of 8).  colsum (optional) += per-column sums of the kept rows (the bias gradient), fp32 atomics.
constexpr int TR_TILES = 4;  // 64-row tiles per CTA

__device__ __forceinline__ int64_t kept_to_row(int64_t r, int group_len, int group_skip) {
  if (group_len <= 0) return r;
  const int per = group_len - group_skip;
  const int64_t q = r / per;
  return q * group_len + group_skip + (r - q * per);
}

__global__ void __launch_bounds__(256) transpose_bf16_kernel(const bf16* __restrict__ in, int64_t ld_in,
                                                             bf16* __restrict__ out, int64_t ld_out, int64_t kept,
                                                             int cols, int group_len, int group_skip,
                                                             float* __restrict__ colsum) {
  __shared__ uint32_t tile[64][33];
  __shared__ float cs[8][64];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c0 = blockIdx.x * 64;
  const bool col_ok = c0 + 2 * tx < cols;
  float s0 = 0.f, s1 = 0.f;
  for (int t = 0; t < TR_TILES; ++t) {
    const int64_t r0 = (static_cast<int64_t>(blockIdx.y) * TR_TILES + t) * 64;
    if (r0 >= ld_out) break;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = ty + 8 * i;
      uint32_t w = 0u;
      if (r0 + r < kept && col_ok) {
        w = *reinterpret_cast<const uint32_t*>(in + kept_to_row(r0 + r, group_len, group_skip) * ld_in + c0 + 2 * tx);
        const float2 f = unpack_bf16x2(w);
        s0 += f.x;
        s1 += f.y;
      }
      tile[r][tx] = w;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = ty + 8 * i;
      if (c0 + c < cols && r0 + 2 * tx < ld_out) {
        const uint32_t a = tile[2 * tx][c >> 1], b = tile[2 * tx + 1][c >> 1];
        const uint32_t lo = (c & 1) ? (a >> 16) : (a & 0xffffu);
        const uint32_t hi = (c & 1) ? (b >> 16) : (b & 0xffffu);
        *reinterpret_cast<uint32_t*>(out + static_cast<int64_t>(c0 + c) * ld_out + r0 + 2 * tx) = lo | (hi << 16);
      }
    }
    __syncthreads();
  }
  if (colsum) {
    cs[ty][2 * tx] = s0;
    cs[ty][2 * tx + 1] = s1;
    __syncthreads();
    if (threadIdx.x < 64 && c0 + threadIdx.x < cols) {
      float a = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) a += cs[i][threadIdx.x];
      atomicAdd(colsum + c0 + threadIdx.x, a);
    }
  }
}

// colsum[c] += sum_r x[r, c]: the bias gradient of a Linear from its output gradient (one read of dY, fp32 atomics).
// A CTA owns 256 columns x a slab of rows: a warp reads 512 contiguous bytes of one row per step (16 bytes per lane) and
// keeps four rows in flight; one shared-memory reduction over the 8 warps, then 256 atomics per CTA.
__global__ void __launch_bounds__(256) colsum_bf16_kernel(const bf16* __restrict__ x, int64_t ld, int64_t rows, int cols,
                                                          int64_t rows_per_block, float* __restrict__ colsum) {
  __shared__ float cs[8][256];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 256 + 8 * tx;
  const int64_t r0 = blockIdx.y * rows_per_block;
  const int64_t r1 = r0 + rows_per_block < rows ? r0 + rows_per_block : rows;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  auto add = [&](const uint4& u) {
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const float2 f = unpack_bf16x2(w[t]);
      acc[2 * t] += f.x;
      acc[2 * t + 1] += f.y;
    }
  };
  if (c < cols) {  // cols % 8 == 0: a lane's 8 columns are all inside or all outside
    const bf16* base = x + c;
    int64_t r = r0 + ty;
    for (; r + 24 < r1; r += 32) {
      const uint4 a = ld_nc_v4(base + r * ld), b = ld_nc_v4(base + (r + 8) * ld), d = ld_nc_v4(base + (r + 16) * ld),
                  e = ld_nc_v4(base + (r + 24) * ld);
      add(a); add(b); add(d); add(e);
    }
    for (; r < r1; r += 8) add(ld_nc_v4(base + r * ld));
  }
#pragma unroll
  for (int t = 0; t < 8; ++t) cs[ty][8 * tx + t] = acc[t];
  __syncthreads();
  const int cc = blockIdx.x * 256 + threadIdx.x;
  if (cc < cols) {
    float a = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) a += cs[i][threadIdx.x];
    atomicAdd(colsum + cc, a);
  }
}

// ------------------------------------------------------------------------------------------- LayerNorm backward
// y = (x - mean) * rstd * gamma + beta per row (reference twin aligner/encoder/slip.py:350-356, fp32 statistics).
//   dx = rstd * (g - mean_k(g) - xhat * mean_k(g * xhat)) [+ add],  g = dy * gamma
//   dgamma += sum_rows dy * xhat,  dbeta += sum_rows dy
// One warp per row (same register layout as layernorm_bf16_kernel: lane owns 16-byte chunks lane + 32 i), warps walk
// rows with a grid stride and keep their dgamma/dbeta partials in registers; one block reduction + atomics at the end.
constexpr int LNB_WARPS = 4;

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst))),
               "l"(gsrc)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Rows reach the warp through a two-slot shared-memory ring filled by cp.async (L2 -> shared, no registers): while row r is
// processed, rows r + 1 and r + 2 (x, dy and the residual gradient: 4.6 KB each at D = 768) are in flight.  Round 1 / early
// round 2 prefetched ONE row of x / dy into registers (and none of `add`: the registers cost a block per SM): 0.55 of the
// copy peak with ~1 us of HBM latency exposed per row.  Packed fp32x2 arithmetic (FFMA2 / FADD2 / FMUL2) halves the issue
// slots of the four passes over the row.
constexpr int ln_bwd_ring_bytes(int chunks) { return LNB_WARPS * 2 * 3 * chunks * 32 * 16; }

template <int LNB_CHUNKS>
__global__ void __launch_bounds__(LNB_WARPS * 32, LNB_CHUNKS == 4 ? 2 : 3) ln_bwd_kernel(const bf16* __restrict__ x, const bf16* dy,
                                                     const float* __restrict__ gamma, const bf16* add, bf16* dx,
                                                     float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                     int64_t rows, int D, float eps) {
  extern __shared__ __align__(16) uint4 lnb_dyn[];  // ln_bwd_ring_bytes(LNB_CHUNKS): 48 KB at D = 1024, above the static limit
  uint4(*ring)[2][3][LNB_CHUNKS * 32] = reinterpret_cast<uint4(*)[2][3][LNB_CHUNKS * 32]>(lnb_dyn);  // [warp][slot][x | dy | add][chunk]
  __shared__ float sgamma[LNB_CHUNKS * 256];
  static_assert(2 * 3 * 32 * 16 >= 256 * 4, "the block reduction re-uses the ring");
  float(*red)[LNB_CHUNKS * 256] = reinterpret_cast<float(*)[LNB_CHUNKS * 256]>(lnb_dyn);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int chunks = D >> 3;
  const float inv_d = 1.f / static_cast<float>(D);
  for (int k = threadIdx.x; k < LNB_CHUNKS * 256; k += LNB_WARPS * 32) sgamma[k] = k < D ? gamma[k] : 0.f;
  __syncthreads();
  uint64_t ag[LNB_CHUNKS][4], ab[LNB_CHUNKS][4];
#pragma unroll
  for (int i = 0; i < LNB_CHUNKS; ++i)
#pragma unroll
    for (int t = 0; t < 4; ++t) ag[i][t] = ab[i][t] = pack_f32x2(0.f, 0.f);
  // every lane copies, and later reads, only its own 16-byte chunks: no cross-lane hazard on the ring.  dx may alias dy or
  // add (training.py backward_video): rows ahead are only READ here, and this warp is the one that later writes them.
  auto issue = [&](int64_t r, int slot) {
    if (r < rows) {
#pragma unroll
      for (int i = 0; i < LNB_CHUNKS; ++i) {
        const int c = lane + 32 * i;
        if (c < chunks) {
          cp_async16(&ring[warp][slot][0][c], reinterpret_cast<const uint4*>(x + r * D) + c);
          cp_async16(&ring[warp][slot][1][c], reinterpret_cast<const uint4*>(dy + r * D) + c);
          if (add) cp_async16(&ring[warp][slot][2][c], reinterpret_cast<const uint4*>(add + r * D) + c);
        }
      }
    }
    cp_async_commit();  // an (empty) group per call keeps the wait_group arithmetic uniform
  };
  // bf16 pair -> two fp32 in one 64-bit register pair: low element = bits << 16, high element = bits & 0xffff0000
  auto widen = [](uint32_t w) { return (static_cast<uint64_t>(w & 0xffff0000u) << 32) | static_cast<uint64_t>(w << 16); };
  const int64_t nwarps = static_cast<int64_t>(gridDim.x) * LNB_WARPS;
  int64_t row = static_cast<int64_t>(blockIdx.x) * LNB_WARPS + warp;
  issue(row, 0);
  issue(row + nwarps, 1);
  for (int slot = 0; row < rows; row += nwarps, slot ^= 1) {
    cp_async_wait<1>();  // everything but the newest group (row + nwarps) has landed
    uint64_t xv[LNB_CHUNKS][4], dv[LNB_CHUNKS][4];
    uint4 av[LNB_CHUNKS];
    uint64_t sum2 = pack_f32x2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < LNB_CHUNKS; ++i) {
      const int c = lane + 32 * i;
      if (c < chunks) {
        const uint4 qx = ring[warp][slot][0][c], qd = ring[warp][slot][1][c];
        if (add) av[i] = ring[warp][slot][2][c];
        const uint32_t uw[4] = {qx.x, qx.y, qx.z, qx.w}, dw[4] = {qd.x, qd.y, qd.z, qd.w};
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          xv[i][t] = widen(uw[t]);
          dv[i][t] = widen(dw[t]);
          sum2 = add_f32x2(sum2, xv[i][t]);
        }
      }
    }
    issue(row + 2 * nwarps, slot);  // the slot just emptied: in flight during this row's and the next row's work
    float s0, s1;
    unpack_f32x2(sum2, s0, s1);
    const float mean = warp_sum(s0 + s1) * inv_d;
    const uint64_t nmean2 = pack_f32x2(-mean, -mean);
    uint64_t sq2 = pack_f32x2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < LNB_CHUNKS; ++i)
      if (lane + 32 * i < chunks) {
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const uint64_t d = add_f32x2(xv[i][t], nmean2);
          sq2 = fma_f32x2(d, d, sq2);
        }
      }
    unpack_f32x2(sq2, s0, s1);
    const float rstd = rsqrtf(warp_sum(s0 + s1) * inv_d + eps);
    const uint64_t rstd2 = pack_f32x2(rstd, rstd), shift2 = pack_f32x2(-mean * rstd, -mean * rstd);
    uint64_t c1_2 = pack_f32x2(0.f, 0.f), c2_2 = c1_2;
#pragma unroll
    for (int i = 0; i < LNB_CHUNKS; ++i) {
      const int c = lane + 32 * i;
      if (c < chunks) {
        const float4 g0 = *reinterpret_cast<const float4*>(sgamma + c * 8);
        const float4 g1 = *reinterpret_cast<const float4*>(sgamma + c * 8 + 4);
        const uint64_t gm[4] = {pack_f32x2(g0.x, g0.y), pack_f32x2(g0.z, g0.w), pack_f32x2(g1.x, g1.y), pack_f32x2(g1.z, g1.w)};
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const uint64_t xh = fma_f32x2(xv[i][t], rstd2, shift2);  // (x - mean) * rstd
          ag[i][t] = fma_f32x2(dv[i][t], xh, ag[i][t]);
          ab[i][t] = add_f32x2(ab[i][t], dv[i][t]);
          const uint64_t g = mul_f32x2(dv[i][t], gm[t]);
          xv[i][t] = xh;
          dv[i][t] = g;  // from here on dv holds dy * gamma
          c1_2 = add_f32x2(c1_2, g);
          c2_2 = fma_f32x2(g, xh, c2_2);
        }
      }
    }
    unpack_f32x2(c1_2, s0, s1);
    const float c1 = warp_sum(s0 + s1) * inv_d;
    unpack_f32x2(c2_2, s0, s1);
    const float c2 = warp_sum(s0 + s1) * inv_d;
    // dx = rstd * (g - c1 - xh * c2) [+ add]
    const uint64_t nc2r = pack_f32x2(-c2 * rstd, -c2 * rstd), nc1r = pack_f32x2(-c1 * rstd, -c1 * rstd);
#pragma unroll
    for (int i = 0; i < LNB_CHUNKS; ++i) {
      const int c = lane + 32 * i;
      if (c < chunks) {
        const uint32_t aw[4] = {av[i].x, av[i].y, av[i].z, av[i].w};
        uint32_t ow[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          uint64_t o = fma_f32x2(dv[i][t], rstd2, fma_f32x2(xv[i][t], nc2r, nc1r));
          if (add) o = add_f32x2(o, widen(aw[t]));
          float o0, o1;
          unpack_f32x2(o, o0, o1);
          ow[t] = pack_bf16x2(o0, o1);
        }
        *(reinterpret_cast<uint4*>(dx + row * D) + c) = make_uint4(ow[0], ow[1], ow[2], ow[3]);
      }
    }
  }
  cp_async_wait<0>();
  // block reduction of the per-warp partial sums through the (now idle) ring, gamma then beta
#pragma unroll 1
  for (int pass = 0; pass < 2; ++pass) {
    __syncthreads();
#pragma unroll
    for (int i = 0; i < LNB_CHUNKS; ++i) {
      const int c = lane + 32 * i;
      if (c < chunks) {
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          float lo, hi;
          unpack_f32x2(pass ? ab[i][t] : ag[i][t], lo, hi);
          red[warp][c * 8 + 2 * t] = lo;
          red[warp][c * 8 + 2 * t + 1] = hi;
        }
      }
    }
    __syncthreads();
    float* dst = pass ? dbeta : dgamma;
    for (int k = threadIdx.x; k < D; k += LNB_WARPS * 32) {
      float a = 0.f;
#pragma unroll
      for (int w = 0; w < LNB_WARPS; ++w) a += red[w][k];
      atomicAdd(dst + k, a);
    }
  }
}

// ------------------------------------------------------------------------------------------- QuickGELU
// g = u * sigmoid(1.702 u) (aligner/encoder/slip.py:359-361);  du = dg * s * (1 + 1.702 u (1 - s)),  s = sigmoid(1.702 u)
// sigmoid(y) = 0.5 + 0.5 tanh(y / 2) with tanh.approx (2^-11, inside the bf16 output rounding; the same form as the
// GEMM's fused QuickGELU epilogue): ONE MUFU op per element.  With exp + divide (two MUFU ops) these kernels were
// co-limited by the MUFU pipe (16 / clk / SM ~ 6.4 TB/s worth of elements): 4.1 TB/s forward.
__device__ __forceinline__ float tanh_mufu(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float sigmoid1702(float u) { return fmaf(0.5f, tanh_mufu(0.851f * u), 0.5f); }
__global__ void __launch_bounds__(256) quickgelu_kernel(const bf16* __restrict__ u, bf16* __restrict__ g, int64_t n8) {
  auto apply = [](const uint4& a) {
    const uint32_t w[4] = {a.x, a.y, a.z, a.w};
    uint32_t o[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const float2 f = unpack_bf16x2(w[t]);
      o[t] = pack_bf16x2(f.x * sigmoid1702(f.x), f.y * sigmoid1702(f.y));
    }
    return make_uint4(o[0], o[1], o[2], o[3]);
  };
  // a CTA walks 16 KiB spans (1024 x 16 bytes): four independent loads in flight per thread, all in one DRAM
  // neighbourhood (a grid-wide stride between a thread's loads cost 25 % of the bandwidth)
  const uint4* src = reinterpret_cast<const uint4*>(u);
  uint4* dst = reinterpret_cast<uint4*>(g);
  for (int64_t base = static_cast<int64_t>(blockIdx.x) * 1024; base < n8; base += static_cast<int64_t>(gridDim.x) * 1024) {
    const int64_t i = base + threadIdx.x;
    if (base + 1024 <= n8) {
      const uint4 a = ld_nc_v4(src + i), b = ld_nc_v4(src + i + 256), c = ld_nc_v4(src + i + 512), d = ld_nc_v4(src + i + 768);
      st_na_v4(dst + i, apply(a));
      st_na_v4(dst + i + 256, apply(b));
      st_na_v4(dst + i + 512, apply(c));
      st_na_v4(dst + i + 768, apply(d));
    } else {
      for (int64_t j = i; j < n8; j += 256) st_na_v4(dst + j, apply(ld_nc_v4(src + j)));
    }
  }
}

// g_out (optional): also writes g = quickgelu(u), which the weight gradient of the following Linear reads -- one pass
// over u instead of a separate recomputation.
__global__ void __launch_bounds__(256) quickgelu_bwd_kernel(const bf16* __restrict__ u, const bf16* dg, bf16* du,
                                                            bf16* __restrict__ g_out, int64_t n8) {
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x; i < n8; i += static_cast<int64_t>(gridDim.x) * 256) {
    const uint4 a = ld_nc_v4(reinterpret_cast<const uint4*>(u) + i);
    const uint4 b = *(reinterpret_cast<const uint4*>(dg) + i);
    const uint32_t w[4] = {a.x, a.y, a.z, a.w}, v[4] = {b.x, b.y, b.z, b.w};
    uint32_t o[4], go[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const float2 f = unpack_bf16x2(w[t]), d = unpack_bf16x2(v[t]);
      const float s0 = sigmoid1702(f.x), s1 = sigmoid1702(f.y);
      o[t] = pack_bf16x2(d.x * s0 * (1.f + 1.702f * f.x * (1.f - s0)), d.y * s1 * (1.f + 1.702f * f.y * (1.f - s1)));
      go[t] = pack_bf16x2(f.x * s0, f.y * s1);
    }
    *(reinterpret_cast<uint4*>(du) + i) = make_uint4(o[0], o[1], o[2], o[3]);
    if (g_out) st_na_v4(reinterpret_cast<uint4*>(g_out) + i, make_uint4(go[0], go[1], go[2], go[3]));
  }
}

// ------------------------------------------------------------------------------------------- attention backward
// qkv rows [q | k | v] (width 3 D), O = softmax(q k^T / 8 [+ causal mask]) v, head dim 64 (nn.MultiheadAttention inside
// ResidualAttentionBlock, aligner/encoder/slip.py:368,378-380).  With P = softmax(S), delta_i = sum_d dO_id O_id:
//   dV = P^T dO,  dP = dO V^T,  dS = P o (dP - delta),  dQ = dS K / 8,  dK = dS^T Q / 8.
// One CTA per (sequence, head); Q, K, V, dO (LP x 64 bf16 each, 128-byte rows, XOR-8 chunk swizzle) stay in shared
// memory.  Phase A: a warp owns 16 query rows: dQ and the row log-sum-exp in ONE sweep over the keys (online softmax
// rescaling of the un-normalised dQ).  Phase B: a warp owns
// 16 keys and walks the queries with the transposed products, producing dK and dV in registers.  P and dS are rounded
// to bf16 for the second MMA of each product, fp32 accumulation throughout.
constexpr int HD = 64;

__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
      "{%0, %1, %2, %3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ uint32_t sw(int row, int chunk) { return row * 128 + ((chunk ^ (row & 7)) << 4); }
__device__ __forceinline__ float ex2(float x) {  // one MUFU op; ex2(-inf) = 0
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// 16 x 64 A-operand fragments of rows row0.. of a swizzled tile
__device__ __forceinline__ void load_a_frags(uint32_t (&f)[4][4], uint32_t tile, int row0, int lane) {
#pragma unroll
  for (int ks = 0; ks < 4; ++ks)
    ldmatrix_x4(f[ks], tile + sw(row0 + (lane & 7) + (((lane >> 3) & 1) << 3), ks * 2 + (lane >> 4)));
}
// acc[NT][4] (16 x 8 NT) = A(16 x 64) . T[n0 .. n0 + 8 NT, 0..64]^T : the tile's rows are the product's columns
template <int NT>
__device__ __forceinline__ void mm_rows_as_cols(float (&acc)[NT][4], const uint32_t (&af)[4][4], uint32_t tile, int n0,
                                                int lane) {
#pragma unroll
  for (int j = 0; j < NT; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
    for (int jp = 0; jp < NT / 2; ++jp) {
      uint32_t b[4];
      const int krow = n0 + jp * 16 + (lane & 7) + ((lane >> 4) << 3);
      ldmatrix_x4(b, tile + sw(krow, ks * 2 + ((lane >> 3) & 1)));
      mma_bf16(acc[2 * jp], af[ks], b[0], b[1]);
      mma_bf16(acc[2 * jp + 1], af[ks], b[2], b[3]);
    }
  }
}
// o[8][4] (16 x 64) += bf16(P)(16 x 8 NT) . T[k0 .. k0 + 8 NT, 0..64] : the tile's rows are the reduction index
template <int NT>
__device__ __forceinline__ void mm_rows_as_k(float (&o)[8][4], const float (&p)[NT][4], uint32_t tile, int k0,
                                             int lane) {
#pragma unroll
  for (int kk = 0; kk < NT / 2; ++kk) {
    uint32_t a[4];
    a[0] = pack_bf16x2(p[2 * kk][0], p[2 * kk][1]);
    a[1] = pack_bf16x2(p[2 * kk][2], p[2 * kk][3]);
    a[2] = pack_bf16x2(p[2 * kk + 1][0], p[2 * kk + 1][1]);
    a[3] = pack_bf16x2(p[2 * kk + 1][2], p[2 * kk + 1][3]);
#pragma unroll
    for (int dn = 0; dn < 4; ++dn) {
      uint32_t b[4];
      const int vrow = k0 + kk * 16 + (lane & 7) + (((lane >> 3) & 1) << 3);
      ldmatrix_x4_trans(b, tile + sw(vrow, dn * 2 + (lane >> 4)));
      mma_bf16(o[2 * dn], a, b[0], b[1]);
      mma_bf16(o[2 * dn + 1], a, b[2], b[3]);
    }
  }
}

// Phase A, one block of keys for the 16 query rows of a warp: S = Q K^T, running row max m and sum l with the usual
// online-softmax rescaling, and dQ~ += (P~ o (dP - delta)) K with P~ = exp2(S c - m c) un-normalised: dQ~ is linear in
// P~, so it is rescaled with l whenever m moves and divided by the final l once, like O in the forward kernel.
template <int NT, bool MASK>
__device__ __forceinline__ void bwd_q_block(const uint32_t (&qf)[4][4], const uint32_t (&dof)[4][4], uint32_t sK,
                                            uint32_t sV, int key0, int L, bool causal, int qrow0, float scale_log2,
                                            const float (&delta)[2], float (&m)[2], float (&l)[2], float (&dq)[8][4],
                                            int lane) {
  float s[NT][4], dp[NT][4];
  mm_rows_as_cols<NT>(s, qf, sK, key0, lane);
  mm_rows_as_cols<NT>(dp, dof, sV, key0, lane);
  const int g = lane >> 2, tq = lane & 3;
  float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
  for (int j = 0; j < NT; ++j)
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      if (MASK) {  // interior blocks (all keys valid and visible to all 16 rows) skip the index arithmetic
        const int col = key0 + j * 8 + tq * 2 + (e & 1), row = qrow0 + g + ((e >> 1) << 3);
        if (col >= L || (causal && col > row)) s[j][e] = -INFINITY;
      }
      mx[e >> 1] = fmaxf(mx[e >> 1], s[j][e]);
    }
  float corr[2], mref[2];
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
    mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
    const float mnew = fmaxf(m[r], mx[r]);
    corr[r] = ex2((m[r] - mnew) * scale_log2);  // m = -inf on the first block -> 0
    m[r] = mnew;
    mref[r] = mnew * scale_log2;
    l[r] *= corr[r];
  }
#pragma unroll
  for (int dn = 0; dn < 8; ++dn) {
    dq[dn][0] *= corr[0];
    dq[dn][1] *= corr[0];
    dq[dn][2] *= corr[1];
    dq[dn][3] *= corr[1];
  }
#pragma unroll
  for (int j = 0; j < NT; ++j)
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float pv = ex2(fmaf(s[j][e], scale_log2, -mref[e >> 1]));  // masked entries: ex2(-inf) = 0
      l[e >> 1] += pv;
      s[j][e] = pv * (dp[j][e] - delta[e >> 1]);
    }
  mm_rows_as_k<NT>(dq, s, sK, key0, lane);
}

// Phase B: for the 16 keys of this warp, one block of queries: dV += P^T dO, dK += dS^T Q.
template <int NT, bool MASK>
__device__ __forceinline__ void bwd_dkv_block(const uint32_t (&kf)[4][4], const uint32_t (&vf)[4][4], uint32_t sQ,
                                              uint32_t sdO, const float* sLse, const float* sDelta, int q0, int L,
                                              bool causal, int krow0, float scale_log2, float (&dk)[8][4],
                                              float (&dv)[8][4], int lane) {
  float st[NT][4], dpt[NT][4];
  mm_rows_as_cols<NT>(st, kf, sQ, q0, lane);    // S^T[key, query]
  mm_rows_as_cols<NT>(dpt, vf, sdO, q0, lane);  // dP^T[key, query]
  const int g = lane >> 2, tq = lane & 3;
#pragma unroll
  for (int j = 0; j < NT; ++j) {
    // this thread's two query columns of the n-tile: log-sum-exp and delta as one 8-byte load each
    const int qc0 = q0 + j * 8 + tq * 2;
    const float2 lse = *reinterpret_cast<const float2*>(sLse + qc0);
    const float2 dlt = *reinterpret_cast<const float2*>(sDelta + qc0);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float pv = ex2(fmaf(st[j][e], scale_log2, -((e & 1) ? lse.y : lse.x)));
      if (MASK) {  // interior blocks (all queries valid and allowed to see all 16 keys) skip the index arithmetic
        const int qcol = qc0 + (e & 1), key = krow0 + g + ((e >> 1) << 3);
        if (qcol >= L || key >= L || (causal && key > qcol)) pv = 0.f;
      }
      st[j][e] = pv;
      dpt[j][e] = pv * (dpt[j][e] - ((e & 1) ? dlt.y : dlt.x));
    }
  }
  mm_rows_as_k<NT>(dv, st, sdO, q0, lane);
  mm_rows_as_k<NT>(dk, dpt, sQ, q0, lane);
}

template <int NWARPS, int MINB>
__global__ void __launch_bounds__(NWARPS * 32, MINB) attention_bwd_kernel(const bf16* __restrict__ qkv,
                                                                    const bf16* __restrict__ O,
                                                                    const bf16* __restrict__ dO,
                                                                    bf16* __restrict__ dqkv, int L, int LP, int D,
                                                                    int causal, float scale, float scale_log2) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int h = blockIdx.x;
  const int64_t seq = blockIdx.y;
  const uint32_t sQ = static_cast<uint32_t>(__cvta_generic_to_shared(smem));
  const uint32_t sK = sQ + LP * 128, sV = sK + LP * 128, sdO = sV + LP * 128;
  float* sLse = reinterpret_cast<float*>(smem + 4 * LP * 128);
  float* sDelta = sLse + LP;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, tq = lane & 3;

  const bf16* base = qkv + seq * L * static_cast<int64_t>(3 * D) + h * HD;
  const bf16* obase = O + seq * L * static_cast<int64_t>(D) + h * HD;
  const bf16* dobase = dO + seq * L * static_cast<int64_t>(D) + h * HD;
  for (int i = threadIdx.x; i < 4 * LP * 8; i += NWARPS * 32) {
    const int part = i / (LP * 8);
    const int rem = i - part * (LP * 8);
    const int row = rem >> 3, chunk = rem & 7;
    const bool valid = row < L;
    const int64_t r = valid ? row : 0;
    const bf16* src = part < 3 ? base + r * (3 * D) + part * D + chunk * 8 : dobase + r * D + chunk * 8;
    cp_async16(sQ + part * (LP * 128) + sw(row, chunk), src, valid);
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  // delta_i = sum_d dO_id O_id, straight from global memory while the tiles land
  for (int row = warp; row < LP; row += NWARPS) {
    float a = 0.f;
    if (row < L) {
      const float2 o = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(obase + static_cast<int64_t>(row) * D + 2 * lane));
      const float2 d = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(dobase + static_cast<int64_t>(row) * D + 2 * lane));
      a = o.x * d.x + o.y * d.y;
    }
    a = warp_sum(a);
    if (lane == 0) sDelta[row] = a;
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();

  const bool cz = causal != 0;
  // ---------------- phase A: query tiles -> log-sum-exp, dQ
  for (int tile = warp; tile < LP / 16; tile += NWARPS) {
    const int qrow0 = tile * 16;
    uint32_t qf[4][4], dof[4][4];
    load_a_frags(qf, sQ, qrow0, lane);
    load_a_frags(dof, sdO, qrow0, lane);
    float m[2] = {-INFINITY, -INFINITY}, l[2] = {0.f, 0.f};
    const float delta[2] = {sDelta[qrow0 + g], sDelta[qrow0 + g + 8]};
    float dq[8][4];
#pragma unroll
    for (int dn = 0; dn < 8; ++dn) dq[dn][0] = dq[dn][1] = dq[dn][2] = dq[dn][3] = 0.f;
    for (int key0 = 0; key0 < LP; key0 += 32) {
      if (cz && key0 > qrow0 + 15) break;
      if (key0 + 32 <= L && (!cz || key0 + 31 <= qrow0))
        bwd_q_block<4, false>(qf, dof, sK, sV, key0, L, cz, qrow0, scale_log2, delta, m, l, dq, lane);
      else if (key0 + 32 <= LP)
        bwd_q_block<4, true>(qf, dof, sK, sV, key0, L, cz, qrow0, scale_log2, delta, m, l, dq, lane);
      else
        bwd_q_block<2, true>(qf, dof, sK, sV, key0, L, cz, qrow0, scale_log2, delta, m, l, dq, lane);
    }
    float inv[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      l[r] += __shfl_xor_sync(0xffffffffu, l[r], 1);
      l[r] += __shfl_xor_sync(0xffffffffu, l[r], 2);
      inv[r] = scale / l[r];
      if (tq == 0) sLse[qrow0 + g + 8 * r] = m[r] * scale_log2 + log2f(l[r]);
    }
    const int r0 = qrow0 + g, r1 = r0 + 8;
    bf16* o0 = dqkv + (seq * L + r0) * static_cast<int64_t>(3 * D) + h * HD + tq * 2;
    bf16* o1 = dqkv + (seq * L + r1) * static_cast<int64_t>(3 * D) + h * HD + tq * 2;
#pragma unroll
    for (int dn = 0; dn < 8; ++dn) {
      if (r0 < L) *reinterpret_cast<uint32_t*>(o0 + dn * 8) = pack_bf16x2(dq[dn][0] * inv[0], dq[dn][1] * inv[0]);
      if (r1 < L) *reinterpret_cast<uint32_t*>(o1 + dn * 8) = pack_bf16x2(dq[dn][2] * inv[1], dq[dn][3] * inv[1]);
    }
  }
  __syncthreads();  // every row's log-sum-exp is in shared memory

  // ---------------- phase B: key tiles -> dK, dV
  for (int tile = warp; tile < LP / 16; tile += NWARPS) {
    const int krow0 = tile * 16;
    uint32_t kf[4][4], vf[4][4];
    load_a_frags(kf, sK, krow0, lane);
    load_a_frags(vf, sV, krow0, lane);
    float dk[8][4], dv[8][4];
#pragma unroll
    for (int dn = 0; dn < 8; ++dn) {
      dk[dn][0] = dk[dn][1] = dk[dn][2] = dk[dn][3] = 0.f;
      dv[dn][0] = dv[dn][1] = dv[dn][2] = dv[dn][3] = 0.f;
    }
    for (int q0 = 0; q0 < LP; q0 += 32) {
      if (cz && q0 + 31 < krow0) continue;  // every query of the block precedes every key of the tile
      if (q0 + 32 <= L && krow0 + 16 <= L && (!cz || krow0 + 15 <= q0))
        bwd_dkv_block<4, false>(kf, vf, sQ, sdO, sLse, sDelta, q0, L, cz, krow0, scale_log2, dk, dv, lane);
      else if (q0 + 32 <= LP)
        bwd_dkv_block<4, true>(kf, vf, sQ, sdO, sLse, sDelta, q0, L, cz, krow0, scale_log2, dk, dv, lane);
      else
        bwd_dkv_block<2, true>(kf, vf, sQ, sdO, sLse, sDelta, q0, L, cz, krow0, scale_log2, dk, dv, lane);
    }
    const int r0 = krow0 + g, r1 = r0 + 8;
    bf16* k0p = dqkv + (seq * L + r0) * static_cast<int64_t>(3 * D) + D + h * HD + tq * 2;
    bf16* k1p = dqkv + (seq * L + r1) * static_cast<int64_t>(3 * D) + D + h * HD + tq * 2;
#pragma unroll
    for (int dn = 0; dn < 8; ++dn) {
      if (r0 < L) {
        *reinterpret_cast<uint32_t*>(k0p + dn * 8) = pack_bf16x2(dk[dn][0] * scale, dk[dn][1] * scale);
        *reinterpret_cast<uint32_t*>(k0p + D + dn * 8) = pack_bf16x2(dv[dn][0], dv[dn][1]);
      }
      if (r1 < L) {
        *reinterpret_cast<uint32_t*>(k1p + dn * 8) = pack_bf16x2(dk[dn][2] * scale, dk[dn][3] * scale);
        *reinterpret_cast<uint32_t*>(k1p + D + dn * 8) = pack_bf16x2(dv[dn][2], dv[dn][3]);
      }
    }
  }
}

template <int NWARPS, int MINB>
int launch_attention_bwd(const bf16* qkv, const bf16* O, const bf16* dO, bf16* dqkv, int64_t seqs, int L, int heads,
                         int causal, cudaStream_t s) {
  const int LP = (L + 15) / 16 * 16;
  const int smem = 4 * LP * 128 + 2 * LP * 4;
  static int configured = 0;
  if (configured < smem) {
    FC_CUDA(cudaFuncSetAttribute(attention_bwd_kernel<NWARPS, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = smem;
  }
  const float scale = 0.125f, scale_log2 = 0.125f * 1.4426950408889634f;
  const int D = heads * HD;
  for (int64_t s0 = 0; s0 < seqs; s0 += 65535) {
    const int64_t n = seqs - s0 < 65535 ? seqs - s0 : 65535;
    dim3 grid(heads, static_cast<unsigned>(n));
    attention_bwd_kernel<NWARPS, MINB><<<grid, NWARPS * 32, smem, s>>>(qkv + s0 * L * static_cast<int64_t>(3 * D),
                                                                  O + s0 * L * static_cast<int64_t>(D),
                                                                  dO + s0 * L * static_cast<int64_t>(D),
                                                                  dqkv + s0 * L * static_cast<int64_t>(3 * D), L, LP, D,
                                                                  causal, scale, scale_log2);
    FC_CHECK_LAUNCH();
  }
  return FC_OK;
}

// ------------------------------------------------------------------------------------------- losses: value + gradient
// nce_loss (aligner/loss.py:13-26, mean reduction):  loss = mean_i(lse_row_i - S_ii) + mean_j(lse_col_j - S_jj)
//   dS_ij = gscale / B * (exp(S_ij - lse_row_i) + exp(S_ij - lse_col_j) - 2 [i == j])
// teacher_student_nce_loss (aligner/loss.py:29-39, "batchmean"): KL(softmax(T) || softmax(S)) over rows plus columns
//   dS_ij = gscale / B * (softmax_row(S) - softmax_row(T) + softmax_col(S) - softmax_col(T))_ij
// lse[0:B] rows of S, [B:2B] columns of S, [2B:3B] rows of T, [3B:4B] columns of T.
__device__ float block_max_f(float v, float* red) {
  v = warp_max(v);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = red[0];
  for (int i = 1; i < (blockDim.x >> 5); ++i) r = fmaxf(r, red[i]);
  __syncthreads();
  return r;
}
__device__ float block_sum_f(float v, float* red) {
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = 0.f;
  for (int i = 0; i < (blockDim.x >> 5); ++i) r += red[i];
  __syncthreads();
  return r;
}

// S / T are (R, C): R video rows, C text columns (R == C unless the unlabelled texts were replaced by prompts,
// teacher_student.py:104-120).  lse layout: rows(S) [R] | cols(S) [C] | rows(T) [R] | cols(T) [C].
__global__ void __launch_bounds__(256) lse_lines_kernel(const float* __restrict__ S, const float* __restrict__ T,
                                                        int64_t ld, int R, int C, float* __restrict__ lse) {
  __shared__ float red[8];
  const int b = blockIdx.x;
  const int half = b / (R + C), in_half = b - half * (R + C);  // half 0: S, half 1: T
  const bool is_col = in_half >= R;
  const int i = is_col ? in_half - R : in_half, n = is_col ? R : C;
  const float* M = half == 0 ? S : T;
  const int64_t off = is_col ? i : static_cast<int64_t>(i) * ld, stride = is_col ? ld : 1;
  float mx = -INFINITY;
  for (int j = threadIdx.x; j < n; j += 256) mx = fmaxf(mx, M[off + j * stride]);
  mx = block_max_f(mx, red);
  float sum = 0.f;
  for (int j = threadIdx.x; j < n; j += 256) sum += expf(M[off + j * stride] - mx);
  sum = block_sum_f(sum, red);
  if (threadIdx.x == 0) lse[b] = mx + logf(sum);
}

__global__ void __launch_bounds__(256) nce_grad_kernel(const float* __restrict__ S, int64_t ld, int B,
                                                       const float* __restrict__ lse, float gscale,
                                                       float* __restrict__ dS, int64_t ldd) {
  const int i = blockIdx.x;
  const float lr = lse[i], k = gscale / static_cast<float>(B);
  for (int j = threadIdx.x; j < B; j += 256) {
    const float s = S[static_cast<int64_t>(i) * ld + j];
    dS[static_cast<int64_t>(i) * ldd + j] = k * (expf(s - lr) + expf(s - lse[B + j]) - (i == j ? 2.f : 0.f));
  }
}

// "batchmean" divides the row direction by R and the column direction (the loss of scores.T, loss.py:36-39) by C
__global__ void __launch_bounds__(256) ts_nce_grad_kernel(const float* __restrict__ S, const float* __restrict__ T,
                                                          int64_t ld, int R, int C, const float* __restrict__ lse,
                                                          float gscale, float* __restrict__ dS, int64_t ldd) {
  const int i = blockIdx.x;
  const float kr = gscale / static_cast<float>(R), kc = gscale / static_cast<float>(C);
  const float* lse_sc = lse + R;
  const float* lse_tr = lse + R + C;
  const float* lse_tc = lse + 2 * R + C;
  for (int j = threadIdx.x; j < C; j += 256) {
    const float s = S[static_cast<int64_t>(i) * ld + j], t = T[static_cast<int64_t>(i) * ld + j];
    dS[static_cast<int64_t>(i) * ldd + j] =
        kr * (expf(s - lse[i]) - expf(t - lse_tr[i])) + kc * (expf(s - lse_sc[j]) - expf(t - lse_tc[j]));
  }
}

// value from the log-sum-exps: terms per line, then one block sums them (double accumulation like reduce_terms_kernel)
__global__ void __launch_bounds__(256) loss_value_kernel(const float* __restrict__ S, const float* __restrict__ T,
                                                         int64_t ld, int R, int C, const float* __restrict__ lse,
                                                         float* __restrict__ out) {
  __shared__ double red[256];
  double a = 0.0;
  if (!T) {  // R == C
    for (int i = threadIdx.x; i < R; i += 256)
      a += static_cast<double>(lse[i] + lse[R + i] - 2.f * S[static_cast<int64_t>(i) * ld + i]) / R;
  } else {
    // sum_ij p_row(T)_ij (log p_row(T)_ij - log p_row(S)_ij) / R + the same over columns / C
    const float* lse_sc = lse + R;
    const float* lse_tr = lse + R + C;
    const float* lse_tc = lse + 2 * R + C;
    double ar = 0.0, ac = 0.0;
    for (int64_t e = threadIdx.x; e < static_cast<int64_t>(R) * C; e += 256) {
      const int i = static_cast<int>(e / C), j = static_cast<int>(e - static_cast<int64_t>(i) * C);
      const float s = S[static_cast<int64_t>(i) * ld + j], t = T[static_cast<int64_t>(i) * ld + j];
      const float ltr = t - lse_tr[i], lsr = s - lse[i], ltc = t - lse_tc[j], lsc = s - lse_sc[j];
      const float pr = expf(ltr), pc = expf(ltc);
      ar += static_cast<double>(pr > 0.f ? pr * (ltr - lsr) : 0.f);
      ac += static_cast<double>(pc > 0.f ? pc * (ltc - lsc) : 0.f);
    }
    a = ar / R + ac / C;
  }
  red[threadIdx.x] = a;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) *out = static_cast<float>(red[0]);
}

// ------------------------------------------------------------------------------------------- small fp32 GEMM
// C[M,N] = alpha * op(A) . op(B)  (score-matrix gradients: B x B x 512).  64 x 64 tile, 4 x 4 outputs per thread.
__global__ void __launch_bounds__(256) sgemm_kernel(int ta, int tb, int M, int N, int K, float alpha,
                                                    const float* __restrict__ A, int64_t lda,
                                                    const float* __restrict__ Bm, int64_t ldb, float* __restrict__ C,
                                                    int64_t ldc) {
  __shared__ float As[16][65], Bs[16][65];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += 16) {
    for (int i = threadIdx.x; i < 1024; i += 256) {
      int kk, mm;
      if (ta) { mm = i & 63; kk = i >> 6; } else { kk = i & 15; mm = i >> 4; }
      const int m = m0 + mm, k = k0 + kk;
      As[kk][mm] = (m < M && k < K) ? (ta ? A[static_cast<int64_t>(k) * lda + m] : A[static_cast<int64_t>(m) * lda + k]) : 0.f;
      int kb, nn;
      if (tb) { kb = i & 15; nn = i >> 4; } else { nn = i & 63; kb = i >> 6; }
      const int n = n0 + nn, k2 = k0 + kb;
      Bs[kb][nn] = (n < N && k2 < K) ? (tb ? Bm[static_cast<int64_t>(n) * ldb + k2] : Bm[static_cast<int64_t>(k2) * ldb + n]) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        a[i] = As[kk][ty * 4 + i];
        b[i] = Bs[kk][tx * 4 + i];
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int m = m0 + ty * 4 + i, n = n0 + tx * 4 + j;
      if (m < M && n < N) C[static_cast<int64_t>(m) * ldc + n] = alpha * acc[i][j];
    }
}

// ------------------------------------------------------------------------------------------- pool + normalise backward
// out[b] = scale * mean_t x[b,t] / ||x[b,t]||  (clip_video_text_encoder.py:85-89)  =>
// dx[b,t] = scale / T * (dout[b] - xhat (xhat . dout[b])) / ||x[b,t]||.  One warp per input row; bf16 output.
__global__ void __launch_bounds__(256) pool_normalize_bwd_kernel(const float* __restrict__ x,
                                                                 const float* __restrict__ dout, bf16* __restrict__ dx,
                                                                 int64_t rows, int T, int D, float scale) {
  const int64_t row = static_cast<int64_t>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const float* xr = x + row * D;
  const float* dr = dout + (row / T) * D;
  float nn = 0.f, dot = 0.f;
  for (int k = lane; k < D; k += 32) {
    const float v = xr[k];
    nn += v * v;
    dot += v * dr[k];
  }
  nn = warp_sum(nn);
  dot = warp_sum(dot);
  const float inv = rsqrtf(nn), k0 = scale / static_cast<float>(T) * inv;
  const float proj = dot * inv * inv;  // (xhat . dout) / ||x||
  for (int k = lane; k < D; k += 32) dx[row * D + k] = __float2bfloat16(k0 * (dr[k] - xr[k] * proj));
}

// ------------------------------------------------------------------------------------------- sequence rows / embeddings
// EOT position of a caption = first argmax of its ids (slip.py:478); row 0 (class token) when ids == nullptr.
__device__ __forceinline__ int eot_index(const int32_t* ids, int64_t s, int L) {
  if (!ids) return 0;
  int best = 0, bv = ids[s * L];
  for (int l = 1; l < L; ++l) {
    const int v = ids[s * L + l];
    if (v > bv) {
      bv = v;
      best = l;
    }
  }
  return best;
}
// gather: out[s] = x[s, eot(s)];  scatter: dx[s, eot(s)] = src[s] (dx zeroed by the caller)
__global__ void __launch_bounds__(128) seq_row_kernel(bf16* x, const int32_t* __restrict__ ids, bf16* rows, int L, int W,
                                                      int scatter) {
  const int64_t s = blockIdx.x;
  __shared__ int pos;
  if (threadIdx.x == 0) pos = eot_index(ids, s, L);
  __syncthreads();
  bf16* xr = x + (s * L + pos) * static_cast<int64_t>(W);
  bf16* rr = rows + s * static_cast<int64_t>(W);
  for (int k = threadIdx.x; k < W / 8; k += 128) {
    if (scatter) reinterpret_cast<uint4*>(xr)[k] = reinterpret_cast<const uint4*>(rr)[k];
    else reinterpret_cast<uint4*>(rr)[k] = reinterpret_cast<const uint4*>(xr)[k];
  }
}

// out[l, :] += sum_s dx[s, l, :]   (positional-embedding gradient; row 0 of the image tower is also d class_embedding)
__global__ void __launch_bounds__(256) seq_sum_kernel(const bf16* __restrict__ dx, float* __restrict__ out, int64_t S,
                                                      int L, int W, int64_t s_per_block) {
  const int l = blockIdx.x;
  const int c = blockIdx.y * 512 + 2 * threadIdx.x;
  if (c >= W) return;
  const int64_t s0 = blockIdx.z * s_per_block, s1 = s0 + s_per_block < S ? s0 + s_per_block : S;
  float a0 = 0.f, a1 = 0.f;
  for (int64_t s = s0; s < s1; ++s) {
    const float2 f = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(dx + (s * L + l) * static_cast<int64_t>(W) + c));
    a0 += f.x;
    a1 += f.y;
  }
  atomicAdd(out + static_cast<int64_t>(l) * W + c, a0);
  atomicAdd(out + static_cast<int64_t>(l) * W + c + 1, a1);
}

// dtok[ids[s,l], :] += dx[s, l, :]  (token_embedding gradient; exact zeros -- rows behind the EOT token of a causal
// sequence -- are skipped, so the padding id does not serialise the atomics)
__global__ void __launch_bounds__(256) token_scatter_kernel(const int32_t* __restrict__ ids, const bf16* __restrict__ dx,
                                                            float* __restrict__ dtok, int64_t tokens, int W, int vocab) {
  const int64_t t = static_cast<int64_t>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (t >= tokens) return;
  const int id = ids[t];
  if (id < 0 || id >= vocab) return;
  const int lane = threadIdx.x & 31;
  for (int c = 2 * lane; c < W; c += 64) {
    const float2 f = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(dx + t * W + c));
    if (f.x != 0.f) atomicAdd(dtok + static_cast<int64_t>(id) * W + c, f.x);
    if (f.y != 0.f) atomicAdd(dtok + static_cast<int64_t>(id) * W + c + 1, f.y);
  }
}

// ------------------------------------------------------------------------------------------- AdamW
// torch.optim.AdamW (config/trainer.yaml:22-24: lr 3e-6, betas (0.9, 0.999), eps 1e-8, weight_decay 0.01), one launch
// over the flat fp32 parameter buffer; also refreshes the bf16 copy the GEMMs read.
// Four parameters per thread and iteration: 16-byte streaming loads of p / g / m / v (nothing is re-used: keep it out of
// L1), 16-byte stores, one 8-byte store of the bf16 mirror -- the scalar version (4-byte accesses, one element in flight per
// thread) ran at 0.56 of the copy peak.
__device__ __forceinline__ float adamw_one(float& p, float g, float& m, float& v, float lr, float beta1, float beta2, float eps,
                                           float decay, float step_size, float bc2_sqrt) {
  p *= decay;
  m = beta1 * m + (1.f - beta1) * g;
  v = beta2 * v + (1.f - beta2) * g * g;
  p -= step_size * (m / (sqrtf(v) / bc2_sqrt + eps));
  return p;
}

__global__ void __launch_bounds__(256) adamw_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                    float* __restrict__ m, float* __restrict__ v,
                                                    bf16* __restrict__ p_bf16, int64_t n, float lr, float beta1,
                                                    float beta2, float eps, float wd, float bc1, float bc2_sqrt) {
  const float decay = 1.f - lr * wd, step_size = lr / bc1;
  const int64_t n4 = n >> 2;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * 256;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x; i < n4; i += stride) {
    uint4 pu = ld_stream_v4(reinterpret_cast<const uint4*>(p) + i);
    const uint4 gu = ld_nc_v4(reinterpret_cast<const uint4*>(g) + i);
    uint4 mu = ld_stream_v4(reinterpret_cast<const uint4*>(m) + i);
    uint4 vu = ld_stream_v4(reinterpret_cast<const uint4*>(v) + i);
    float pf[4] = {__uint_as_float(pu.x), __uint_as_float(pu.y), __uint_as_float(pu.z), __uint_as_float(pu.w)};
    const float gf[4] = {__uint_as_float(gu.x), __uint_as_float(gu.y), __uint_as_float(gu.z), __uint_as_float(gu.w)};
    float mf[4] = {__uint_as_float(mu.x), __uint_as_float(mu.y), __uint_as_float(mu.z), __uint_as_float(mu.w)};
    float vf[4] = {__uint_as_float(vu.x), __uint_as_float(vu.y), __uint_as_float(vu.z), __uint_as_float(vu.w)};
#pragma unroll
    for (int t = 0; t < 4; ++t) adamw_one(pf[t], gf[t], mf[t], vf[t], lr, beta1, beta2, eps, decay, step_size, bc2_sqrt);
    st_na_v4(reinterpret_cast<uint4*>(p) + i, make_uint4(__float_as_uint(pf[0]), __float_as_uint(pf[1]), __float_as_uint(pf[2]), __float_as_uint(pf[3])));
    st_na_v4(reinterpret_cast<uint4*>(m) + i, make_uint4(__float_as_uint(mf[0]), __float_as_uint(mf[1]), __float_as_uint(mf[2]), __float_as_uint(mf[3])));
    st_na_v4(reinterpret_cast<uint4*>(v) + i, make_uint4(__float_as_uint(vf[0]), __float_as_uint(vf[1]), __float_as_uint(vf[2]), __float_as_uint(vf[3])));
    if (p_bf16)
      *reinterpret_cast<uint2*>(p_bf16 + 4 * i) = make_uint2(pack_bf16x2(pf[0], pf[1]), pack_bf16x2(pf[2], pf[3]));
  }
  // tail (n not a multiple of 4) and unaligned callers never happen for the trainer's flat buffers; kept for the generic op
  for (int64_t i = 4 * n4 + static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x; i < n; i += stride) {
    float pi = p[i], mi = m[i], vi = v[i];
    adamw_one(pi, g[i], mi, vi, lr, beta1, beta2, eps, decay, step_size, bc2_sqrt);
    p[i] = pi;
    m[i] = mi;
    v[i] = vi;
    if (p_bf16) p_bf16[i] = __float2bfloat16(pi);
  }
}

// Split count of a weight-gradient GEMM: (K range, tile) work items are dealt round-robin to the SM pairs, so the
// launch takes ceil(items / pairs) waves; pick the split whose last wave is fullest (2..4 waves, >= 8 K blocks per item).
// E.g. out_proj of ViT-B/16 (9 tiles, 74 pairs): 17 splits = 153 items = 3 waves at 69 %, 16 splits = 144 items = 2 waves at 97 %.
int pick_k_splits(int M, int N, int K) {
  const int num_k = (K + 63) / 64;
  const int64_t tiles = static_cast<int64_t>((((M + 127) / 128) + 1) / 2) * ((N + 255) / 256);
  const int pairs = num_sms() / 2;
  int best = 1;
  double best_score = -1.0;
  for (int sp = 1; sp <= num_k && sp <= 256; ++sp) {
    if (sp > 1 && num_k / sp < 8) break;
    const int64_t items = tiles * sp;
    const int64_t waves = (items + pairs - 1) / pairs;
    if (waves > 4 && sp > 1) break;
    double score = static_cast<double>(items) / static_cast<double>(waves * pairs);
    if (waves < 2) score *= 0.9;  // one wave leaves the pipeline prologue / accumulator drain of every item exposed
    if (score > best_score + 1e-9) {
      best_score = score;
      best = sp;
    }
  }
  return best;
}

int grid_for(int64_t work_items, int per_block) {
  const int64_t blocks = (work_items + per_block - 1) / per_block;
  const int64_t cap = static_cast<int64_t>(num_sms()) * 8;
  return static_cast<int>(blocks < 1 ? 1 : (blocks < cap ? blocks : cap));
}

}  // namespace

}  // namespace fc

using namespace fc;

// =================================================================================================== C ABI
extern "C" {

int fc_gemm_bf16_splitk(const void* A, int64_t lda, const void* B, int64_t ldb, float* C, int64_t ldc, float alpha,
                        int32_t M, int32_t N, int32_t K, int32_t k_splits, void* stream) {
  FC_REQUIRE(C, "fc_gemm_bf16_splitk: null C");
  GemmParams p;
  p.M = M; p.N = N; p.K = K; p.C = C; p.ldc = ldc; p.alpha = alpha;
  const int num_k = (K + 63) / 64;
  if (k_splits <= 0) k_splits = pick_k_splits(M, N, K);
  p.k_splits = k_splits > num_k ? num_k : k_splits;
  return gemm_bf16_tn(EPI_F32_SPLITK, static_cast<const bf16*>(A), lda, static_cast<const bf16*>(B), ldb, p,
                      static_cast<cudaStream_t>(stream));
}

int fc_gemm_bf16_layout(int epilogue, int a_mn, int b_mn, const void* A, int64_t lda, const void* B, int64_t ldb, void* C,
                        int64_t ldc, const float* bias, const void* resid, int64_t ldr, float alpha, int32_t M,
                        int32_t N, int32_t K, int32_t k_splits, void* stream) {
  GemmParams p;
  p.M = M; p.N = N; p.K = K; p.C = C; p.ldc = ldc; p.bias = bias;
  p.resid = static_cast<const bf16*>(resid); p.ldr = ldr; p.alpha = alpha;
  p.a_mn = a_mn != 0; p.b_mn = b_mn != 0;
  if (epilogue == EPI_F32_SPLITK) {
    const int num_k = (K + 63) / 64;
    if (k_splits <= 0) k_splits = pick_k_splits(M, N, K);
    p.k_splits = k_splits > num_k ? num_k : k_splits;
  } else {
    FC_REQUIRE(epilogue == EPI_BIAS || epilogue == EPI_BIAS_RESID || epilogue == EPI_F32 || epilogue == EPI_QGELU_BWD,
               "fc_gemm_bf16_layout: epilogue %d is not exposed", epilogue);
  }
  return gemm_bf16_tn(epilogue, static_cast<const bf16*>(A), lda, static_cast<const bf16*>(B), ldb, p,
                      static_cast<cudaStream_t>(stream));
}

int fc_colsum_bf16(const void* x, int64_t ld, int64_t rows, int32_t cols, float* colsum, void* stream) {
  FC_REQUIRE(x && colsum && cols > 0 && cols % 8 == 0 && ld % 8 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0,
             "fc_colsum_bf16: cols and ld must be multiples of 8, x 16-byte aligned");
  if (rows == 0) return FC_OK;
  const int col_blocks = (cols + 255) / 256;
  int64_t slabs = (4 * static_cast<int64_t>(num_sms()) + col_blocks - 1) / col_blocks;  // ~4 CTAs per SM
  const int64_t max_slabs = (rows + 63) / 64;
  if (slabs > max_slabs) slabs = max_slabs;
  if (slabs > 65535) slabs = 65535;
  const int64_t per = (rows + slabs - 1) / slabs;
  dim3 grid(col_blocks, static_cast<unsigned>((rows + per - 1) / per));
  ProfScope prof(static_cast<cudaStream_t>(stream), PROF_OTHER, 11, rows, cols, 0, 0.0, 2.0 * rows * cols);
  colsum_bf16_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const bf16*>(x), ld, rows, cols,
                                                                          per, colsum);
  FC_CHECK_LAUNCH();
  return FC_OK;
}

int fc_transpose_bf16(const void* in, int64_t ld_in, void* out, int64_t ld_out, int64_t kept_rows, int32_t cols,
                      int32_t group_len, int32_t group_skip, float* colsum, void* stream) {
  FC_REQUIRE(in && out, "fc_transpose_bf16: null pointer");
  FC_REQUIRE(kept_rows >= 0 && cols > 0 && cols % 2 == 0 && ld_in % 2 == 0 && ld_out % 2 == 0 && ld_out >= kept_rows,
             "fc_transpose_bf16: bad shape rows=%lld cols=%d ld_in=%lld ld_out=%lld", static_cast<long long>(kept_rows),
             cols, static_cast<long long>(ld_in), static_cast<long long>(ld_out));
  FC_REQUIRE(group_len == 0 || (group_skip >= 0 && group_skip < group_len), "fc_transpose_bf16: bad row groups");
  FC_REQUIRE((reinterpret_cast<uintptr_t>(in) & 3) == 0 && (reinterpret_cast<uintptr_t>(out) & 3) == 0,
             "fc_transpose_bf16: pointers must be 4-byte aligned");
  if (ld_out == 0) return FC_OK;
  const int64_t row_blocks = (ld_out + 64 * TR_TILES - 1) / (64 * TR_TILES);
  FC_REQUIRE(row_blocks <= 65535, "fc_transpose_bf16: too many rows");
  dim3 grid((cols + 63) / 64, static_cast<unsigned>(row_blocks));
  ProfScope prof(static_cast<cudaStream_t>(stream), PROF_OTHER, 10, kept_rows, cols, 0, 0.0, 4.0 * kept_rows * cols);
  transpose_bf16_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const bf16*>(in), ld_in, static_cast<bf16*>(out), ld_out, kept_rows, cols, group_len, group_skip,
      colsum);
  FC_CHECK_LAUNCH();
  return FC_OK;
}

int fc_layernorm_bwd_bf16(const void* x, const void* dy, const float* gamma, const void* add, void* dx, float* dgamma,
                          float* dbeta, int64_t rows, int32_t D, float eps, void* stream) {
  FC_REQUIRE(x && dy && gamma && dx && dgamma && dbeta, "fc_layernorm_bwd_bf16: null pointer");
  FC_REQUIRE(D % 8 == 0 && D >= 8 && D <= 1024, "fc_layernorm_bwd_bf16: D=%d must be a multiple of 8, <= 1024", D);
  if (rows == 0) return FC_OK;
  const int64_t want = (rows + LNB_WARPS - 1) / LNB_WARPS;
  const int blocks = static_cast<int>(want < 6 * num_sms() ? want : 6 * num_sms());
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  ProfScope prof(s, PROF_OTHER, 12, rows, D, 0, 0.0, (add ? 8.0 : 6.0) * rows * D);
  const bf16 *xb = static_cast<const bf16*>(x), *dyb = static_cast<const bf16*>(dy), *ab = static_cast<const bf16*>(add);
  bf16* dxb = static_cast<bf16*>(dx);
  static bool configured = false;
  if (!configured) {  // D = 1024: 48 KB of ring + the static gamma copy exceed the default 48 KB limit
    FC_CUDA(cudaFuncSetAttribute(ln_bwd_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, ln_bwd_ring_bytes(4)));
    configured = true;
  }
  if (D <= 256) ln_bwd_kernel<1><<<blocks, LNB_WARPS * 32, ln_bwd_ring_bytes(1), s>>>(xb, dyb, gamma, ab, dxb, dgamma, dbeta, rows, D, eps);
  else if (D <= 512) ln_bwd_kernel<2><<<blocks, LNB_WARPS * 32, ln_bwd_ring_bytes(2), s>>>(xb, dyb, gamma, ab, dxb, dgamma, dbeta, rows, D, eps);
  else if (D <= 768) ln_bwd_kernel<3><<<blocks, LNB_WARPS * 32, ln_bwd_ring_bytes(3), s>>>(xb, dyb, gamma, ab, dxb, dgamma, dbeta, rows, D, eps);
  else ln_bwd_kernel<4><<<blocks, LNB_WARPS * 32, ln_bwd_ring_bytes(4), s>>>(xb, dyb, gamma, ab, dxb, dgamma, dbeta, rows, D, eps);
  FC_CHECK_LAUNCH();
  return FC_OK;
}

int fc_quickgelu_bf16(const void* u, void* g, int64_t n, void* stream) {
  FC_REQUIRE(u && g && n % 8 == 0, "fc_quickgelu_bf16: null pointer or n %% 8 != 0");
  if (n == 0) return FC_OK;
  ProfScope prof(static_cast<cudaStream_t>(stream), PROF_OTHER, 13, n, 1, 0, 0.0, 4.0 * n);
  quickgelu_kernel<<<grid_for(n / 8, 1024), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const bf16*>(u), static_cast<bf16*>(g), n / 8);
  FC_CHECK_LAUNCH();
  return FC_OK;
}

int fc_quickgelu_bwd_bf16(const void* u, const void* dg, void* du, void* g_out, int64_t n, void* stream) {
  FC_REQUIRE(u && dg && du && n % 8 == 0, "fc_quickgelu_bwd_bf16: null pointer or n %% 8 != 0");
  if (n == 0) return FC_OK;
  ProfScope prof(static_cast<cudaStream_t>(stream), PROF_OTHER, 14, n, 1, 0, 0.0, (g_out ? 8.0 : 6.0) * n);
  quickgelu_bwd_kernel<<<grid_for(n / 8, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const bf16*>(u), static_cast<const bf16*>(dg), static_cast<bf16*>(du), static_cast<bf16*>(g_out), n / 8);
  FC_CHECK_LAUNCH();
  return FC_OK;
}

int fc_attention_bwd_bf16(const void* qkv, const void* out, const void* dout, void* dqkv, int64_t seqs, int32_t L,
                          int32_t heads, int32_t causal, void* stream) {
  FC_REQUIRE(qkv && out && dout && dqkv, "fc_attention_bwd_bf16: null pointer");
  FC_REQUIRE(L >= 1 && L <= 432 && heads >= 1, "fc_attention_bwd_bf16: sequence length %d unsupported (1..432)", L);
  if (seqs == 0) return FC_OK;
  const bf16* q = static_cast<const bf16*>(qkv);
  const bf16* o = static_cast<const bf16*>(out);
  const bf16* d = static_cast<const bf16*>(dout);
  bf16* dq = static_cast<bf16*>(dqkv);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int tiles = (L + 15) / 16;
  ProfScope prof(s, PROF_ATTENTION, 2 + causal, seqs, L, heads, 10.0 * L * L * HD * heads * static_cast<double>(seqs),
                 static_cast<double>(seqs) * L * heads * HD * 2.0 * 8.0);
  {
    int handled = 0;
    const int rc = attention_bwd_bf16_tc(q, o, d, dq, seqs, L, heads, causal, s, &handled);
    if (rc || handled) return rc;
  }
  // warps per CTA x CTAs per SM (register cap): FC_ATTN_BWD_CFG = 10 * warps + min_blocks overrides (diagnostics)
  static int cfg_override = -1;
  if (cfg_override < 0) {
    const char* e = getenv("FC_ATTN_BWD_CFG");
    cfg_override = e ? atoi(e) : 0;
  }
  int cfg = cfg_override;
  if (!cfg) {
    // measured on B200 (tools/attention_bwd_bench.py, 2048 x 197 x 12 heads): 6 warps x 2 CTAs/SM 6.14 ms, 5 x 2 6.41,
    // 4 x 2 7.05, 7 x 2 (128 registers, spills) 7.29, 7 x 1 8.20, 5 x 1 10.7; the 77-token causal case agrees
    if (tiles <= 4) cfg = 42;
    else if (tiles == 5) cfg = 52;
    else cfg = 62;
  }
  switch (cfg) {
    case 41: return launch_attention_bwd<4, 1>(q, o, d, dq, seqs, L, heads, causal, s);
    case 42: return launch_attention_bwd<4, 2>(q, o, d, dq, seqs, L, heads, causal, s);
    case 51: return launch_attention_bwd<5, 1>(q, o, d, dq, seqs, L, heads, causal, s);
    case 52: return launch_attention_bwd<5, 2>(q, o, d, dq, seqs, L, heads, causal, s);
    case 61: return launch_attention_bwd<6, 1>(q, o, d, dq, seqs, L, heads, causal, s);
    case 62: return launch_attention_bwd<6, 2>(q, o, d, dq, seqs, L, heads, causal, s);
    case 72: return launch_attention_bwd<7, 2>(q, o, d, dq, seqs, L, heads, causal, s);
    case 81: return launch_attention_bwd<8, 1>(q, o, d, dq, seqs, L, heads, causal, s);
    default: return launch_attention_bwd<7, 1>(q, o, d, dq, seqs, L, heads, causal, s);
  }
}

// lse: 2 (rows + cols) floats of workspace.  teacher == NULL: nce_loss (rows == cols);  else
// TeacherStudentNCELoss("batchmean") of (rows, cols) scores.  loss_out (optional) = the loss value, dscores (optional,
// (rows, ldd)) = gscale * d loss / d scores.
int fc_loss_fwd_bwd(const float* scores, const float* teacher, int64_t ld, int32_t rows, int32_t cols, float* lse,
                    float gscale, float* loss_out, float* dscores, int64_t ldd, void* stream) {
  FC_REQUIRE(scores && lse && rows >= 1 && cols >= 1 && ld >= cols, "fc_loss_fwd_bwd: bad arguments");
  FC_REQUIRE(teacher || rows == cols, "fc_loss_fwd_bwd: nce_loss needs a square score matrix (%d x %d)", rows, cols);
  FC_REQUIRE(!dscores || ldd >= cols, "fc_loss_fwd_bwd: bad gradient pitch");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  lse_lines_kernel<<<(teacher ? 2 : 1) * (rows + cols), 256, 0, s>>>(scores, teacher, ld, rows, cols, lse);
  FC_CHECK_LAUNCH();
  if (loss_out) {
    loss_value_kernel<<<1, 256, 0, s>>>(scores, teacher, ld, rows, cols, lse, loss_out);
    FC_CHECK_LAUNCH();
  }
  if (dscores) {
    if (teacher) ts_nce_grad_kernel<<<rows, 256, 0, s>>>(scores, teacher, ld, rows, cols, lse, gscale, dscores, ldd);
    else nce_grad_kernel<<<rows, 256, 0, s>>>(scores, ld, rows, lse, gscale, dscores, ldd);
    FC_CHECK_LAUNCH();
  }
  return FC_OK;
}

int fc_sgemm_f32(int32_t trans_a, int32_t trans_b, int32_t M, int32_t N, int32_t K, float alpha, const float* A,
                 int64_t lda, const float* B, int64_t ldb, float* C, int64_t ldc, void* stream) {
  FC_REQUIRE(A && B && C && M > 0 && N > 0 && K > 0, "fc_sgemm_f32: bad arguments");
  dim3 grid((N + 63) / 64, (M + 63) / 64);
  sgemm_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(trans_a, trans_b, M, N, K, alpha, A, lda, B, ldb, C,
                                                                    ldc);
  FC_CHECK_LAUNCH();
  return FC_OK;
}

int fc_pool_normalize_bwd(const float* x, const float* dout, void* dx_bf16, int64_t rows_out, int32_t T, int32_t D,
                          float scale, void* stream) {
  FC_REQUIRE(x && dout && dx_bf16 && T >= 1 && D >= 1, "fc_pool_normalize_bwd: bad arguments");
  const int64_t rows = rows_out * T;
  if (rows == 0) return FC_OK;
  pool_normalize_bwd_kernel<<<static_cast<unsigned>((rows + 7) / 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      x, dout, static_cast<bf16*>(dx_bf16), rows, T, D, scale);
  FC_CHECK_LAUNCH();
  return FC_OK;
}

int fc_seq_rows(void* x, const int32_t* ids, void* rows, int64_t seqs, int32_t L, int32_t W, int32_t scatter,
                void* stream) {
  FC_REQUIRE(x && rows && L >= 1 && W % 8 == 0, "fc_seq_rows: bad arguments");
  if (seqs == 0) return FC_OK;
  seq_row_kernel<<<static_cast<unsigned>(seqs), 128, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<bf16*>(x), ids, static_cast<bf16*>(rows), L, W, scatter);
  FC_CHECK_LAUNCH();
  return FC_OK;
}

int fc_seq_sum(const void* dx, float* out, int64_t seqs, int32_t L, int32_t W, void* stream) {
  FC_REQUIRE(dx && out && L >= 1 && W % 2 == 0, "fc_seq_sum: bad arguments");
  if (seqs == 0) return FC_OK;
  const int64_t zb = seqs < 16 ? seqs : 16;
  const int64_t per = (seqs + zb - 1) / zb;
  dim3 grid(L, (W + 511) / 512, static_cast<unsigned>((seqs + per - 1) / per));
  seq_sum_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const bf16*>(dx), out, seqs, L, W, per);
  FC_CHECK_LAUNCH();
  return FC_OK;
}

int fc_token_scatter_add(const int32_t* ids, const void* dx, float* dtok, int64_t tokens, int32_t W, int32_t vocab,
                         void* stream) {
  FC_REQUIRE(ids && dx && dtok && W % 2 == 0, "fc_token_scatter_add: bad arguments");
  if (tokens == 0) return FC_OK;
  token_scatter_kernel<<<static_cast<unsigned>((tokens + 7) / 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      ids, static_cast<const bf16*>(dx), dtok, tokens, W, vocab);
  FC_CHECK_LAUNCH();
  return FC_OK;
}

int fc_adamw_step(float* p, const float* g, float* m, float* v, void* p_bf16, int64_t n, float lr, float beta1,
                  float beta2, float eps, float weight_decay, int32_t step, void* stream) {
  FC_REQUIRE(p && g && m && v && step >= 1, "fc_adamw_step: bad arguments");
  if (n == 0) return FC_OK;
  const float bc1 = static_cast<float>(1.0 - pow(static_cast<double>(beta1), step));
  const float bc2s = static_cast<float>(sqrt(1.0 - pow(static_cast<double>(beta2), step)));
  ProfScope prof(static_cast<cudaStream_t>(stream), PROF_OTHER, 15, n, 1, 0, 0.0, (p_bf16 ? 30.0 : 28.0) * n);
  FC_REQUIRE(((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
               reinterpret_cast<uintptr_t>(v)) & 15) == 0 && (reinterpret_cast<uintptr_t>(p_bf16) & 7) == 0,
             "fc_adamw_step: buffers must be 16-byte aligned (the bf16 mirror 8-byte)");
  adamw_kernel<<<grid_for(n, 4096), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      p, g, m, v, static_cast<bf16*>(p_bf16), n, lr, beta1, beta2, eps, weight_decay, bc1, bc2s);
  FC_CHECK_LAUNCH();
  return FC_OK;
}

int fc_f32_to_bf16(const float* in, void* out, int64_t n, void* stream) {
  FC_REQUIRE(in && out, "fc_f32_to_bf16: null pointer");
  return f32_to_bf16(in, static_cast<bf16*>(out), n, static_cast<cudaStream_t>(stream));
}

// Training-forward front ends of the two towers (the evaluation path runs them inside fc_encode_*).
// patches: bf16 (F * G * G, 3 P P) scratch that the patch-embed weight gradient reuses; x: bf16 (F * (G G + 1), W).
int fc_patch_embed(const void* frames, int dtype, const void* conv_w_bf16, const float* cls, const float* pos,
                   void* patches, void* x, int64_t F, int32_t R, int32_t P, int32_t W, void* stream) {
  FC_REQUIRE(frames && conv_w_bf16 && cls && pos && patches && x, "fc_patch_embed: null pointer");
  FC_REQUIRE(R % P == 0 && (3 * P * P) % 8 == 0, "fc_patch_embed: 3 P P must be a multiple of 8 on the training path");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int G = R / P, pd = 3 * P * P;
  int rc;
  if ((rc = im2col_patches(frames, dtype, static_cast<bf16*>(patches), F, R, P, s))) return rc;
  GemmParams p;
  p.M = static_cast<int>(F * G * G); p.N = W; p.K = pd; p.C = x; p.ldc = W; p.pos = pos; p.patches_per_frame = G * G;
  if ((rc = gemm_bf16_tn(EPI_PATCH, static_cast<const bf16*>(patches), pd, static_cast<const bf16*>(conv_w_bf16), pd, p, s)))
    return rc;
  return cls_rows(static_cast<bf16*>(x), cls, pos, F, G * G + 1, W, s);
}

int fc_text_embed(const int32_t* ids, const float* tok, const float* pos, void* x, int64_t C, int32_t L, int32_t W,
                  int32_t vocab, int32_t* err_flag, void* stream) {
  FC_REQUIRE(ids && tok && pos && x && err_flag, "fc_text_embed: null pointer");
  return text_embed(ids, tok, pos, static_cast<bf16*>(x), C, L, W, vocab, err_flag, nullptr,
                    static_cast<cudaStream_t>(stream));
}

}  // extern "C"
