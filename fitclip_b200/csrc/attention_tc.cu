// tcgen05 fused attention for the CLIP image sequence (197 tokens, no mask, head dim 64) -- K4 of SURVEY.md 2.2;
// reference: nn.MultiheadAttention(need_weights=False) in ResidualAttentionBlock (twin aligner/encoder/slip.py:378-380).
//
// One work item = (sequence, head, 128-row query tile); persistent CTAs (2 per SM, 256 TMEM columns each) loop over
// items.  Per item:
//   TMA      Q tile [128 x 64], K [KP x 64], V [KP x 64] through 3-D tensor maps (dims: column, token-in-sequence,
//            sequence) so tokens beyond the sequence are ZERO-FILLED on load and CLIPPED on store -- a tile never
//            touches its neighbour sequence.
//   MMA      S = Q K^T      tcgen05.mma M=128 N=KP K=16 x4, both operands K-major (128B swizzle), S -> TMEM cols [0,KP)
//   softmax  4 warps, one query row per thread: two passes over the row in TMEM (max; exp2 + sum), P rounded to bf16
//            and written BACK INTO TMEM over the dead S columns (tcgen05.st) -- P never touches shared memory
//   MMA      O = P V        tcgen05.mma M=128 N=64 K=16 x KP/16, A = P from TMEM, B = V MN-major from smem -> TMEM
//   epilogue O / rowsum -> bf16 -> per-warp 32x64 staging tile -> 3-D TMA store
// The TMA/MMA thread prefetches the next item's Q,K as soon as S is done and V as soon as O is done, so loads overlap the
// softmax; the second co-resident CTA overlaps its MMAs with this CTA's MUFU-bound softmax.
#include <stdlib.h>
#include <string.h>

#include "kernels.cuh"
#include "ptx.cuh"

namespace fc {

namespace {

constexpr int HD = 64;
constexpr int QT = 128;                    // query rows per tile
constexpr int Q_BYTES = QT * 128;          // 16 KiB
constexpr int STG_BYTES = 4 * 32 * 128;    // 4 warps x (32 rows x 128 B)
constexpr int O_COL = 128;                 // O accumulator columns [128, 192): inside the (dead) S region
constexpr uint32_t TMEM_COLS_ATT = 256;
constexpr int ATT_THREADS = 160;           // warps 0-3 softmax/epilogue, warp 4 TMA + MMA + TMEM alloc

template <int KP>
struct AttSmem {
  static constexpr int KV_BYTES = KP * 128;
  static constexpr int OFF_Q = 0;
  static constexpr int OFF_K = Q_BYTES;
  static constexpr int OFF_V = OFF_K + ((KV_BYTES + 1023) / 1024) * 1024;
  static constexpr int OFF_STG = OFF_V + ((KV_BYTES + 1023) / 1024) * 1024;
  static constexpr int OFF_BAR = OFF_STG + STG_BYTES;
  static constexpr int BYTES = OFF_BAR + 64;
};

struct AttItem {
  int seq, head, tile;
};

template <int KP>
__global__ void __launch_bounds__(ATT_THREADS, 2)
attention_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                    const __grid_constant__ CUtensorMap tmO, int L, int heads, int tiles, int num_items,
                    float scale_log2) {
  using S = AttSmem<KP>;
  static_assert(KP % 16 == 0 && KP >= 16 && KP <= 256, "padded key count");
  static_assert(KP / 2 <= O_COL && O_COL + HD <= 256, "P / O column ranges must not overlap");
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* kq_full = reinterpret_cast<uint64_t*>(smem + S::OFF_BAR);
  uint64_t* v_full = kq_full + 1;
  uint64_t* s_full = kq_full + 2;
  uint64_t* p_full = kq_full + 3;
  uint64_t* o_full = kq_full + 4;
  uint64_t* o_empty = kq_full + 5;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(kq_full + 6);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int D = heads * HD;
  const bool flip_ok = (gridDim.x & 1) == 0 && tiles == 2;  // alternate heavy/light tiles between iterations

  auto decode = [&](int item, int it) {
    AttItem w;
    int t = item % tiles;
    const int sh = item / tiles;
    if (flip_ok) t ^= (it & 1);
    w.tile = t;
    w.head = sh % heads;
    w.seq = sh / heads;
    return w;
  };

  griddep_launch_dependents();
  if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0) __trap();
  if (threadIdx.x == 128) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmKV);
    tma_prefetch_desc(&tmO);
    mbar_init(kq_full, 1);
    mbar_init(v_full, 1);
    mbar_init(s_full, 1);
    mbar_init(p_full, 128);
    mbar_init(o_full, 1);
    mbar_init(o_empty, 128);
    fence_barrier_init();
  }
  if (warp == 4) tmem_alloc<TMEM_COLS_ATT>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  griddep_wait();

  if (warp == 4) {
    // ===================== TMA + MMA thread =====================
    if (lane == 0) {
      constexpr uint32_t idesc_s = umma_idesc_bf16_f32(QT, KP);
      constexpr uint32_t idesc_pv = umma_idesc_bf16_f32_bmn(QT, HD);
      const uint32_t q_addr = smem_u32(smem + S::OFF_Q);
      const uint32_t k_addr = smem_u32(smem + S::OFF_K);
      const uint32_t v_addr = smem_u32(smem + S::OFF_V);
      auto load_qk = [&](const AttItem& w) {
        mbar_expect_tx(kq_full, Q_BYTES + S::KV_BYTES);
        tma_load_3d(smem + S::OFF_Q, &tmQ, kq_full, w.head * HD, w.tile * QT, w.seq);
        tma_load_3d(smem + S::OFF_K, &tmKV, kq_full, D + w.head * HD, 0, w.seq);
      };
      auto load_v = [&](const AttItem& w) {
        mbar_expect_tx(v_full, S::KV_BYTES);
        tma_load_3d(smem + S::OFF_V, &tmKV, v_full, 2 * D + w.head * HD, 0, w.seq);
      };
      int it = 0;
      if (static_cast<int>(blockIdx.x) < num_items) {
        const AttItem w0 = decode(blockIdx.x, 0);
        load_qk(w0);
        load_v(w0);
      }
      for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++it) {
        const uint32_t ph = it & 1;
        const int next = item + gridDim.x;
        // S(it) overwrites TMEM columns the epilogue of item it-1 may still be reading
        if (it > 0) mbar_wait(o_empty, (it - 1) & 1);
        mbar_wait(kq_full, ph);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < HD / 16; ++k)
          umma_bf16_ss(tmem_base, umma_desc_k_sw128(q_addr + k * 32), umma_desc_k_sw128(k_addr + k * 32), idesc_s,
                       k != 0);
        umma_commit(s_full);
        mbar_wait(s_full, ph);  // Q and K tiles are free again
        if (next < num_items) load_qk(decode(next, it + 1));
        mbar_wait(v_full, ph);
        mbar_wait(p_full, ph);  // all 128 rows of P are in TMEM
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < KP / 16; ++k)
          umma_bf16_ts(tmem_base + O_COL, tmem_base + k * 8, umma_desc_mn_sw128(v_addr + k * 2048), idesc_pv, k != 0);
        umma_commit(o_full);
        mbar_wait(o_full, ph);  // V tile is free again
        if (next < num_items) load_v(decode(next, it + 1));
      }
    }
  } else {
    // ===================== softmax + epilogue warps (one query row per thread) =====================
    const uint32_t trow = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
    uint8_t* stg_ptr = smem + S::OFF_STG + warp * (32 * 128);
    const uint32_t stg_row = smem_u32(stg_ptr) + lane * 128;
    const int sw = lane & 7;
    constexpr int NFULL = KP / 32;      // x32 chunks
    constexpr bool REM = (KP % 32) != 0;  // one trailing x16 chunk
    int it = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++it) {
      const uint32_t ph = it & 1;
      const AttItem w = decode(item, it);
      const int row0 = w.tile * QT + warp * 32;
      const bool active = row0 < L;  // warp-uniform: warps whose 32 rows all lie beyond the sequence only sync
      mbar_wait(s_full, ph);
      tc_fence_after();
      float l = 0.f;
      if (active) {
        // ---- pass 1: row maximum
        float m = -INFINITY;
        {
          uint32_t r[2][32];
          uint32_t r16[16];
          tmem_ld_32x32b_x32(trow, r[0]);
#pragma unroll
          for (int j = 0; j < NFULL; ++j) {
            tmem_ld_wait_fence(r[j & 1]);
            if (j + 1 < NFULL) tmem_ld_32x32b_x32(trow + (j + 1) * 32, r[(j + 1) & 1]);
            else if (REM) tmem_ld_32x32b_x16(trow + NFULL * 32, r16);
            if ((j + 1) * 32 <= L) {
#pragma unroll
              for (int c = 0; c < 32; ++c) m = fmaxf(m, __uint_as_float(r[j & 1][c]));
            } else {
#pragma unroll
              for (int c = 0; c < 32; ++c)
                if (j * 32 + c < L) m = fmaxf(m, __uint_as_float(r[j & 1][c]));
            }
          }
          if (REM) {
            tmem_ld_wait_fence16(r16);
#pragma unroll
            for (int c = 0; c < 16; ++c)
              if (NFULL * 32 + c < L) m = fmaxf(m, __uint_as_float(r16[c]));
          }
        }
        // ---- pass 2: P = exp2((s - m) * scale * log2e), row sum in fp32, P -> bf16 -> TMEM (over the S columns)
        const float mc = m * scale_log2;
        {
          uint32_t r[2][32];
          uint32_t r16[16];
          tmem_ld_32x32b_x32(trow, r[0]);
#pragma unroll
          for (int j = 0; j < NFULL; ++j) {
            tmem_ld_wait_fence(r[j & 1]);
            if (j + 1 < NFULL) tmem_ld_32x32b_x32(trow + (j + 1) * 32, r[(j + 1) & 1]);
            else if (REM) tmem_ld_32x32b_x16(trow + NFULL * 32, r16);
            uint32_t pk[16];
            const bool full = (j + 1) * 32 <= L;
#pragma unroll
            for (int c = 0; c < 16; ++c) {
              float p0 = ex2_approx(fmaf(__uint_as_float(r[j & 1][2 * c]), scale_log2, -mc));
              float p1 = ex2_approx(fmaf(__uint_as_float(r[j & 1][2 * c + 1]), scale_log2, -mc));
              if (!full) {
                if (j * 32 + 2 * c >= L) p0 = 0.f;
                if (j * 32 + 2 * c + 1 >= L) p1 = 0.f;
              }
              l += p0 + p1;
              pk[c] = pack_bf16x2(p0, p1);
            }
            tmem_st_32x32b_x16(trow + j * 16, pk);
          }
          if (REM) {
            tmem_ld_wait_fence16(r16);
            uint32_t pk[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              float p0 = ex2_approx(fmaf(__uint_as_float(r16[2 * c]), scale_log2, -mc));
              float p1 = ex2_approx(fmaf(__uint_as_float(r16[2 * c + 1]), scale_log2, -mc));
              if (NFULL * 32 + 2 * c >= L) p0 = 0.f;
              if (NFULL * 32 + 2 * c + 1 >= L) p1 = 0.f;
              l += p0 + p1;
              pk[c] = pack_bf16x2(p0, p1);
            }
            tmem_st_32x32b_x8(trow + NFULL * 16, pk);
          }
          tmem_st_wait();
        }
      }
      tc_fence_before();
      mbar_arrive(p_full);

      // ---- epilogue: O / l -> bf16 -> staging -> TMA store
      mbar_wait(o_full, ph);
      tc_fence_after();
      uint32_t o0[32], o1[32];
      if (active) {
        tmem_ld_32x32b_x32(trow + O_COL, o0);
        tmem_ld_32x32b_x32(trow + O_COL + 32, o1);
        tmem_ld_wait_fence(o0);
        tmem_ld_wait_fence(o1);
      }
      tc_fence_before();
      mbar_arrive(o_empty);
      if (active) {
        const float inv = 1.f / l;
        if (lane == 0) bulk_wait_group_read<0>();  // this warp's previous store has finished reading its staging tile
        __syncwarp();
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const uint32_t(&rr)[32] = c < 4 ? o0 : o1;
          const int o = (c & 3) * 8;
          uint4 u;
          u.x = pack_bf16x2(__uint_as_float(rr[o + 0]) * inv, __uint_as_float(rr[o + 1]) * inv);
          u.y = pack_bf16x2(__uint_as_float(rr[o + 2]) * inv, __uint_as_float(rr[o + 3]) * inv);
          u.z = pack_bf16x2(__uint_as_float(rr[o + 4]) * inv, __uint_as_float(rr[o + 5]) * inv);
          u.w = pack_bf16x2(__uint_as_float(rr[o + 6]) * inv, __uint_as_float(rr[o + 7]) * inv);
          st_shared_v4(stg_row + ((c ^ sw) << 4), u);
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_3d(&tmO, stg_ptr, w.head * HD, row0, w.seq);  // rows >= L are clipped by the tensor map
          bulk_commit_group();
        }
      }
    }
    if (lane == 0) bulk_wait_group<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc<TMEM_COLS_ATT>(tmem_base);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
      return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(sym);
  }
  return fn;
}

// bf16 tensor viewed as [seqs][L][cols] (cols contiguous); box = 64 columns x box_rows tokens x 1 sequence, SW128.
int make_tmap_3d(CUtensorMap* tm, const bf16* base, int64_t cols, int64_t L, int64_t seqs, int box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled entry point not available");
    return FC_ERR_CUDA;
  }
  cuuint64_t dims[3] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(L), static_cast<cuuint64_t>(seqs)};
  cuuint64_t strides[2] = {static_cast<cuuint64_t>(cols) * 2, static_cast<cuuint64_t>(cols) * 2 * L};
  cuuint32_t box[3] = {64, static_cast<cuuint32_t>(box_rows), 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<bf16*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(3d) failed (CUresult %d) cols=%lld L=%lld seqs=%lld", static_cast<int>(r),
              static_cast<long long>(cols), static_cast<long long>(L), static_cast<long long>(seqs));
    return FC_ERR_CUDA;
  }
  return FC_OK;
}

template <int KP>
int launch_tc(const bf16* qkv, bf16* out, int64_t seqs, int L, int heads, cudaStream_t s) {
  using S = AttSmem<KP>;
  static bool configured = false;
  if (!configured) {
    FC_CUDA(cudaFuncSetAttribute(attention_tc_kernel<KP>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::BYTES));
    configured = true;
  }
  const int D = heads * HD;
  CUtensorMap tq, tkv, to;
  int rc;
  if ((rc = make_tmap_3d(&tq, qkv, 3 * D, L, seqs, QT))) return rc;
  if ((rc = make_tmap_3d(&tkv, qkv, 3 * D, L, seqs, KP))) return rc;
  if ((rc = make_tmap_3d(&to, out, D, L, seqs, 32))) return rc;
  const int tiles = (L + QT - 1) / QT;
  const int64_t items64 = seqs * heads * tiles;
  FC_REQUIRE(items64 < (int64_t(1) << 31), "attention: too many work items");
  const int items = static_cast<int>(items64);
  int grid = 2 * num_sms();
  if (grid > items) grid = items;
  if (tiles == 2 && (grid & 1)) grid -= 1;
  const float scale_log2 = 0.125f * 1.4426950408889634f;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(ATT_THREADS);
  cfg.dynamicSmemBytes = S::BYTES;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  note_launch();
  FC_CUDA(cudaLaunchKernelEx(&cfg, attention_tc_kernel<KP>, tq, tkv, to, L, heads, tiles, items, scale_log2));
  return FC_OK;
}

}  // namespace

// tcgen05 path: un-masked sequences of 193..208 tokens (the ViT-B/16 image sequence, 197). Returns 1 if it handled the call.
int attention_bf16_tc(const bf16* qkv, bf16* out, int64_t seqs, int L, int heads, int causal, cudaStream_t s,
                      int* handled) {
  *handled = 0;
  static int disabled = -1;
  if (disabled < 0) {
    const char* e = getenv("FC_ATTENTION");
    disabled = (e && strcmp(e, "mma") == 0) ? 1 : 0;
  }
  if (disabled || causal || L <= 192 || L > 208) return FC_OK;
  FC_REQUIRE((reinterpret_cast<uintptr_t>(qkv) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
             "attention: buffers must be 16-byte aligned");
  *handled = 1;
  return launch_tc<208>(qkv, out, seqs, L, heads, s);
}

}  // namespace fc
