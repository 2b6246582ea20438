// Fused multi-head attention for the two CLIP sequence shapes (K4): 197 image tokens (no mask) and 77 text tokens
// (causal), head dim 64.  Reference: nn.MultiheadAttention inside ResidualAttentionBlock
// (twin aligner/encoder/slip.py:368,378-380; additive -inf upper-triangular mask :454-460).
//
// v1: one CTA per (sequence, head); the whole K and V of that head stay resident in shared memory (<= 208 keys), each
// warp owns 16-query row tiles, S = QK^T and O = PV run on mma.sync.m16n8k16 (bf16 in, fp32 acc) with an online
// softmax over at most two key blocks, P never leaves registers.  Shared-memory rows are 128 bytes with an XOR-8
// swizzle of the 16-byte chunks so that every ldmatrix is bank-conflict free.
#include "kernels.cuh"

namespace fc {

namespace {

constexpr int HD = 64;  // head dim

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
  const int sz = valid ? 16 : 0;  // src-size 0 => 16 bytes of zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
// byte offset of (row, 16-byte chunk) inside a [rows][128 B] swizzled tile
__device__ __forceinline__ uint32_t sw(int row, int chunk) { return row * 128 + ((chunk ^ (row & 7)) << 4); }

// One key block of NT n-tiles (8 keys each) starting at key0: S = QK^T, online softmax update, O += P V.
template <int NT>
__device__ __forceinline__ void key_block(const uint32_t (&qf)[4][4], uint32_t sK, uint32_t sV, int key0, int L,
                                          bool causal, int qrow0, float scale_log2, float (&o)[8][4], float (&m)[2],
                                          float (&l)[2]) {
  const int lane = threadIdx.x & 31;
  float s[NT][4];
#pragma unroll
  for (int j = 0; j < NT; ++j) s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;

  // ---- S = Q K^T
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
    for (int jp = 0; jp < NT / 2; ++jp) {
      uint32_t b[4];
      const int krow = key0 + jp * 16 + (lane & 7) + ((lane >> 4) << 3);
      ldmatrix_x4(b, sK + sw(krow, ks * 2 + ((lane >> 3) & 1)));
      mma_bf16(s[2 * jp], qf[ks], b[0], b[1]);
      mma_bf16(s[2 * jp + 1], qf[ks], b[2], b[3]);
    }
  }

  // ---- mask + row max
  const int g = lane >> 2, tq = lane & 3;
  float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
  for (int j = 0; j < NT; ++j) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int col = key0 + j * 8 + tq * 2 + (e & 1);
      const int row = qrow0 + g + ((e >> 1) << 3);
      const bool dead = col >= L || (causal && col > row);
      if (dead) s[j][e] = -INFINITY;
      mx[e >> 1] = fmaxf(mx[e >> 1], s[j][e]);
    }
  }
  float corr[2], mref[2];
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
    mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
    const float mnew = fmaxf(m[r], mx[r]);
    corr[r] = exp2f((m[r] - mnew) * scale_log2);  // m = -inf on the first block -> 0
    m[r] = mnew;
    mref[r] = mnew * scale_log2;
    l[r] *= corr[r];
  }
#pragma unroll
  for (int dn = 0; dn < 8; ++dn) {
    o[dn][0] *= corr[0];
    o[dn][1] *= corr[0];
    o[dn][2] *= corr[1];
    o[dn][3] *= corr[1];
  }
  // ---- P = exp2(S*scale - m), row sums (fp32), then O += P V with P rounded to bf16
#pragma unroll
  for (int j = 0; j < NT; ++j) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float pv = exp2f(fmaf(s[j][e], scale_log2, -mref[e >> 1]));
      s[j][e] = pv;
      l[e >> 1] += pv;
    }
  }
#pragma unroll
  for (int kk = 0; kk < NT / 2; ++kk) {
    uint32_t a[4];
    a[0] = pack_bf16x2(s[2 * kk][0], s[2 * kk][1]);
    a[1] = pack_bf16x2(s[2 * kk][2], s[2 * kk][3]);
    a[2] = pack_bf16x2(s[2 * kk + 1][0], s[2 * kk + 1][1]);
    a[3] = pack_bf16x2(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
    for (int dn = 0; dn < 4; ++dn) {
      uint32_t b[4];
      const int vrow = key0 + kk * 16 + (lane & 7) + (((lane >> 3) & 1) << 3);
      ldmatrix_x4_trans(b, sV + sw(vrow, dn * 2 + (lane >> 4)));
      mma_bf16(o[2 * dn], a, b[0], b[1]);
      mma_bf16(o[2 * dn + 1], a, b[2], b[3]);
    }
  }
}

// LP = L padded to a multiple of 16; keys are processed as block 0 (NT0 n-tiles) then block 1 (NT1 n-tiles, may be 0).
template <int LP, int NT0, int NT1, int NWARPS>
__global__ void __launch_bounds__(NWARPS * 32) attention_kernel(const bf16* __restrict__ qkv, bf16* __restrict__ out,
                                                                int L, int D, int causal, float scale_log2) {
  static_assert((NT0 + NT1) * 8 == LP, "key blocks must cover the padded length");
  extern __shared__ __align__(128) uint8_t smem[];
  const int h = blockIdx.x;
  const int64_t seq = blockIdx.y;
  const uint32_t sQ = static_cast<uint32_t>(__cvta_generic_to_shared(smem));
  const uint32_t sK = sQ + LP * 128;
  const uint32_t sV = sK + LP * 128;

  // ---- stage Q, K, V of this (sequence, head): 16-byte cp.async, zero fill for padded rows
  const bf16* base = qkv + seq * L * static_cast<int64_t>(3 * D) + h * HD;
  for (int i = threadIdx.x; i < 3 * LP * 8; i += NWARPS * 32) {
    const int part = i / (LP * 8);
    const int rem = i - part * (LP * 8);
    const int row = rem >> 3, chunk = rem & 7;
    const bool valid = row < L;
    const bf16* src = base + static_cast<int64_t>(valid ? row : 0) * (3 * D) + part * D + chunk * 8;
    cp_async16(sQ + part * (LP * 128) + sw(row, chunk), src, valid);
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, tq = lane & 3;
  for (int tile = warp; tile < LP / 16; tile += NWARPS) {
    const int qrow0 = tile * 16;
    uint32_t qf[4][4];
#pragma unroll
    for (int ks = 0; ks < 4; ++ks)
      ldmatrix_x4(qf[ks], sQ + sw(qrow0 + (lane & 7) + (((lane >> 3) & 1) << 3), ks * 2 + (lane >> 4)));
    float o[8][4];
#pragma unroll
    for (int dn = 0; dn < 8; ++dn) o[dn][0] = o[dn][1] = o[dn][2] = o[dn][3] = 0.f;
    float m[2] = {-INFINITY, -INFINITY}, l[2] = {0.f, 0.f};

    key_block<NT0>(qf, sK, sV, 0, L, causal != 0, qrow0, scale_log2, o, m, l);
    if (NT1 > 0) {
      // causal: the second block only matters for query tiles that reach into it (warp-uniform)
      if (!causal || qrow0 + 15 >= NT0 * 8)
        key_block<(NT1 > 0 ? NT1 : 2)>(qf, sK, sV, NT0 * 8, L, causal != 0, qrow0, scale_log2, o, m, l);
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      l[r] += __shfl_xor_sync(0xffffffffu, l[r], 1);
      l[r] += __shfl_xor_sync(0xffffffffu, l[r], 2);
    }
    const float inv0 = 1.f / l[0], inv1 = 1.f / l[1];
    const int r0 = qrow0 + g, r1 = r0 + 8;
    bf16* o0 = out + (seq * L + r0) * static_cast<int64_t>(D) + h * HD + tq * 2;
    bf16* o1 = out + (seq * L + r1) * static_cast<int64_t>(D) + h * HD + tq * 2;
#pragma unroll
    for (int dn = 0; dn < 8; ++dn) {
      if (r0 < L) *reinterpret_cast<uint32_t*>(o0 + dn * 8) = pack_bf16x2(o[dn][0] * inv0, o[dn][1] * inv0);
      if (r1 < L) *reinterpret_cast<uint32_t*>(o1 + dn * 8) = pack_bf16x2(o[dn][2] * inv1, o[dn][3] * inv1);
    }
  }
}

// Long sequences (209..768 tokens: ViT-L/14 has 257, ViT-L/14@336 577 -- SURVEY.md 8f row f4): the K and V of one
// (sequence, head) still fit shared memory, the queries do not, so a CTA takes a chunk of QCH = NWARPS * 16 query rows
// and walks the keys in blocks of 64 with the same online-softmax block routine.
template <int NWARPS>
__global__ void __launch_bounds__(NWARPS * 32) attention_long_kernel(const bf16* __restrict__ qkv, bf16* __restrict__ out,
                                                                     int L, int LPK, int D, int causal,
                                                                     float scale_log2) {
  constexpr int QCH = NWARPS * 16;
  extern __shared__ __align__(128) uint8_t smem[];
  const int h = blockIdx.x;
  const int64_t seq = blockIdx.y;
  const int q0 = blockIdx.z * QCH;
  const uint32_t sQ = static_cast<uint32_t>(__cvta_generic_to_shared(smem));
  const uint32_t sK = sQ + QCH * 128;
  const uint32_t sV = sK + LPK * 128;
  // causal: keys beyond the last query row of this chunk are never needed
  const int kend = causal ? min(LPK, ((q0 + QCH + 63) / 64) * 64) : LPK;

  const bf16* base = qkv + seq * L * static_cast<int64_t>(3 * D) + h * HD;
  for (int i = threadIdx.x; i < QCH * 8; i += NWARPS * 32) {
    const int row = i >> 3, chunk = i & 7;
    const bool valid = q0 + row < L;
    cp_async16(sQ + sw(row, chunk), base + static_cast<int64_t>(valid ? q0 + row : 0) * (3 * D) + chunk * 8, valid);
  }
  for (int i = threadIdx.x; i < 2 * kend * 8; i += NWARPS * 32) {
    const int part = i / (kend * 8);
    const int rem = i - part * (kend * 8);
    const int row = rem >> 3, chunk = rem & 7;
    const bool valid = row < L;
    const bf16* src = base + static_cast<int64_t>(valid ? row : 0) * (3 * D) + (part + 1) * D + chunk * 8;
    cp_async16((part ? sV : sK) + sw(row, chunk), src, valid);
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, tq = lane & 3;
  const int qrow0 = q0 + warp * 16;  // global query row of this warp's tile
  if (qrow0 >= L) return;
  uint32_t qf[4][4];
#pragma unroll
  for (int ks = 0; ks < 4; ++ks)
    ldmatrix_x4(qf[ks], sQ + sw(warp * 16 + (lane & 7) + (((lane >> 3) & 1) << 3), ks * 2 + (lane >> 4)));
  float o[8][4];
#pragma unroll
  for (int dn = 0; dn < 8; ++dn) o[dn][0] = o[dn][1] = o[dn][2] = o[dn][3] = 0.f;
  float m[2] = {-INFINITY, -INFINITY}, l[2] = {0.f, 0.f};
  for (int key0 = 0; key0 < kend; key0 += 64) {
    if (causal && key0 > qrow0 + 15) break;  // warp-uniform: the whole block lies above the diagonal
    key_block<8>(qf, sK, sV, key0, L, causal != 0, qrow0, scale_log2, o, m, l);
  }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    l[r] += __shfl_xor_sync(0xffffffffu, l[r], 1);
    l[r] += __shfl_xor_sync(0xffffffffu, l[r], 2);
  }
  const float inv0 = 1.f / l[0], inv1 = 1.f / l[1];
  const int r0 = qrow0 + g, r1 = r0 + 8;
  bf16* o0 = out + (seq * L + r0) * static_cast<int64_t>(D) + h * HD + tq * 2;
  bf16* o1 = out + (seq * L + r1) * static_cast<int64_t>(D) + h * HD + tq * 2;
#pragma unroll
  for (int dn = 0; dn < 8; ++dn) {
    if (r0 < L) *reinterpret_cast<uint32_t*>(o0 + dn * 8) = pack_bf16x2(o[dn][0] * inv0, o[dn][1] * inv0);
    if (r1 < L) *reinterpret_cast<uint32_t*>(o1 + dn * 8) = pack_bf16x2(o[dn][2] * inv1, o[dn][3] * inv1);
  }
}

int launch_long(const bf16* qkv, bf16* out, int64_t seqs, int L, int heads, int causal, cudaStream_t s) {
  constexpr int NWARPS = 8, QCH = NWARPS * 16;
  const int LPK = (L + 63) / 64 * 64;
  const int smem = QCH * 128 + 2 * LPK * 128;
  static int configured = 0;
  if (configured < smem) {
    FC_CUDA(cudaFuncSetAttribute(attention_long_kernel<NWARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = smem;
  }
  const float scale_log2 = 0.125f * 1.4426950408889634f;
  for (int64_t s0 = 0; s0 < seqs; s0 += 65535) {
    const int64_t n = seqs - s0 < 65535 ? seqs - s0 : 65535;
    dim3 grid(heads, static_cast<unsigned>(n), (L + QCH - 1) / QCH);
    attention_long_kernel<NWARPS><<<grid, NWARPS * 32, smem, s>>>(
        qkv + s0 * L * static_cast<int64_t>(3 * heads * HD), out + s0 * L * static_cast<int64_t>(heads * HD), L, LPK,
        heads * HD, causal, scale_log2);
    FC_CHECK_LAUNCH();
  }
  return FC_OK;
}

template <int LP, int NT0, int NT1, int NWARPS>
int launch(const bf16* qkv, bf16* out, int64_t seqs, int L, int heads, int causal, cudaStream_t s) {
  constexpr int smem = 3 * LP * 128;
  static bool configured = false;
  if (!configured) {
    FC_CUDA(cudaFuncSetAttribute(attention_kernel<LP, NT0, NT1, NWARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 smem));
    configured = true;
  }
  const float scale_log2 = 0.125f * 1.4426950408889634f;  // 1/sqrt(64) * log2(e)
  // gridDim.y is limited to 65535 sequences per launch
  for (int64_t s0 = 0; s0 < seqs; s0 += 65535) {
    const int64_t n = seqs - s0 < 65535 ? seqs - s0 : 65535;
    dim3 grid(heads, static_cast<unsigned>(n));
    attention_kernel<LP, NT0, NT1, NWARPS><<<grid, NWARPS * 32, smem, s>>>(
        qkv + s0 * L * static_cast<int64_t>(3 * heads * HD), out + s0 * L * static_cast<int64_t>(heads * HD), L,
        heads * HD, causal, scale_log2);
    FC_CHECK_LAUNCH();
  }
  return FC_OK;
}

}  // namespace

int attention_bf16(const bf16* qkv, bf16* out, int64_t seqs, int L, int heads, int causal, cudaStream_t s) {
  FC_REQUIRE(qkv && out, "attention: null pointer");
  FC_REQUIRE(L >= 1 && L <= 768, "attention: sequence length %d unsupported (1..768)", L);
  if (seqs == 0) return FC_OK;
  // dense count (QK^T + PV = 4 L^2 64 per head), the figure SURVEY.md 8d uses for both towers
  ProfScope prof(s, PROF_ATTENTION, causal, seqs, L, heads, 4.0 * L * L * HD * heads * static_cast<double>(seqs),
                 static_cast<double>(seqs) * L * heads * HD * 2.0 * 4.0);
  {
    int handled = 0;
    int rc = attention_bf16_tc(qkv, out, seqs, L, heads, causal, s, &handled);
    if (rc || handled) return rc;
    rc = attention_bf16_tc_long(qkv, out, seqs, L, heads, causal, s, &handled);
    if (rc || handled) return rc;
  }
  if (L > 208) return launch_long(qkv, out, seqs, L, heads, causal, s);  // ViT-L/14: 257, ViT-L/14@336: 577
  if (L <= 16) return launch<16, 2, 0, 1>(qkv, out, seqs, L, heads, causal, s);
  if (L <= 32) return launch<32, 4, 0, 2>(qkv, out, seqs, L, heads, causal, s);
  if (L <= 48) return launch<48, 6, 0, 3>(qkv, out, seqs, L, heads, causal, s);
  if (L <= 64) return launch<64, 8, 0, 4>(qkv, out, seqs, L, heads, causal, s);
  if (L <= 80) return launch<80, 10, 0, 5>(qkv, out, seqs, L, heads, causal, s);     // CLIP text: 77
  if (L <= 112) return launch<112, 14, 0, 7>(qkv, out, seqs, L, heads, causal, s);
  if (L <= 160) return launch<160, 10, 10, 5>(qkv, out, seqs, L, heads, causal, s);
  return launch<208, 14, 12, 7>(qkv, out, seqs, L, heads, causal, s);                // CLIP ViT-B/16 image: 197
}

}  // namespace fc
