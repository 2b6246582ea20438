// tcgen05 fused attention for un-masked sequences of 209..768 tokens, head dim 64: the image sequences of the larger
// CLIP geometries the reference ships configs for (ViT-L/14: 257 tokens, config/encoder/clip_vit_l_14.yaml;
// ViT-L/14@336: 577 tokens, clip_vit_l_14_336px.yaml).  Reference arithmetic: nn.MultiheadAttention(need_weights=False)
// in ResidualAttentionBlock (twin aligner/encoder/slip.py:378-380).
//
// A row of S no longer fits the 256 TMEM columns a CTA gets (two CTAs per SM), so keys are walked in blocks of KB = 192
// (128 for 209..256 tokens, where 192 saves no block) with an ONLINE softmax; written out for KB = 128:
//   work item   (sequence, head, 128-row query tile); persistent CTAs, 2 per SM
//   TMA         Q tile once per item; K / V blocks [KB x 64], each with its own full / empty barriers (double-buffered for
//               KB = 128, single for 192), 3-D tensor maps: rows beyond the sequence are zero-filled on load and clipped
//               on store
//   MMA         S_j = Q K_j^T   (M = 128, N = 128, fp32, TMEM cols [0, 128))
//   softmax     one query row per thread: m' = max(m, rowmax(S_j)), P_j = exp2((S_j - m') c) as bf16 over the dead S
//               columns, l = l a + rowsum(P_j) with a = exp2((m - m') c); if any row of the warp moved its maximum,
//               the O accumulator (TMEM cols [128, 192)) is rescaled by a in place (tcgen05.ld / st)
//   MMA         O += P_j V_j    (A = P from TMEM, B = V MN-major from smem), issued right before S_{j+1}: the tensor pipe
//               runs its instructions in order, so S_{j+1} may overwrite P_j and its completion (s_full) tells the
//               softmax threads that O holds every block up to j
//   epilogue    O / l -> bf16 -> per-warp staging tile -> TMA store
// The chain S_j -> softmax -> P_j.V_j -> S_{j+1} is serial inside a CTA; the second co-resident CTA fills the gaps.
#include <stdlib.h>
#include <string.h>

#include "kernels.cuh"
#include "ptx.cuh"
#include "tmap.cuh"

namespace fc {

namespace {

constexpr int HD = 64;
constexpr int QT = 128;   // query rows per tile
constexpr int Q_BYTES = 128 * 128;     // a [128 x 64] bf16 tile, 128-byte rows
constexpr int LONG_THREADS = 160;      // warps 0-3 softmax / epilogue, warp 4 TMA + MMA + TMEM alloc
#ifndef ATT_LONG_POLY
#define ATT_LONG_POLY 4  // of every 16 column pairs of a x32 chunk, this many take exp2 on the FMA pipe (ex2_poly_x2)
#endif

// KB keys per block: S in TMEM columns [0, KB), P (bf16) over [0, KB / 2), O in [KB, KB + 64).
//   KB = 128: 256 TMEM columns (192 used), K / V double-buffered, 2 CTAs per SM
//   KB = 192: all 256 columns, K / V single-buffered (80 KB of shared memory keeps 2 CTAs per SM): 257 tokens are 2 key
//             blocks instead of 3, 577 are 4 instead of 5 -- the per-item chain S -> softmax -> P.V is that much shorter
template <int KB>
struct LongCfg {
  static_assert(KB == 128 || KB == 192, "key block: 128 or 192");
  static constexpr int STAGES = KB == 128 ? 2 : 1;
  static constexpr int KV_BYTES = KB * 128;               // a [KB x 64] bf16 tile
  static constexpr int O_COL = KB;
  static constexpr uint32_t TMEM_COLS = 256;
  static constexpr int CTAS_PER_SM = 2;
  static constexpr int OFF_Q = 0;
  static constexpr int OFF_K = Q_BYTES;
  static constexpr int OFF_V = OFF_K + STAGES * KV_BYTES;
  static constexpr int OFF_STG = OFF_V + STAGES * KV_BYTES;  // 4 warps x (32 rows x 128 B)
  static constexpr int OFF_BAR = OFF_STG + Q_BYTES;
  static constexpr int BYTES = OFF_BAR + 128;
};

// Softmax of one key block for one query row (thread): returns the block's contribution to the row sum and updates the
// running maximum; P goes to TMEM as bf16 over the S columns already read.  FULL: all 128 keys of the block exist (every
// block but the last): fully unrolled, no masks.  !FULL: `valid` < 128 keys exist; S has (valid + 15) & ~15 columns, only
// the x32 chunks that hold a valid key are touched, stale columns are masked.
template <int KB, bool FULL>
__device__ __forceinline__ void softmax_block(uint32_t trow, int valid, bool first, float scale_log2, float& m_run,
                                              float& l, float& alpha, bool& moved) {
  const int nch = FULL ? KB / 32 : (valid + 31) >> 5;
  // ---- pass 1: block maximum
  float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
  uint32_t r[2][32];
  tmem_ld_32x32b_x32(trow, r[0]);
  if (KB > 32 && (FULL || nch > 1)) tmem_ld_32x32b_x32(trow + 32, r[1]);
#pragma unroll
  for (int c4 = 0; c4 < KB / 32; ++c4) {
    if (!FULL && c4 >= nch) break;
    tmem_ld_wait_fence(r[c4 & 1]);
    const uint32_t(&rc)[32] = r[c4 & 1];
    if (FULL || (c4 + 1) * 32 <= valid) {
#pragma unroll
      for (int c = 0; c < 32; c += 8) {
        m0 = max3(m0, __uint_as_float(rc[c]), __uint_as_float(rc[c + 1]));
        m1 = max3(m1, __uint_as_float(rc[c + 2]), __uint_as_float(rc[c + 3]));
        m2 = max3(m2, __uint_as_float(rc[c + 4]), __uint_as_float(rc[c + 5]));
        m3 = max3(m3, __uint_as_float(rc[c + 6]), __uint_as_float(rc[c + 7]));
      }
    } else {
#pragma unroll
      for (int c = 0; c < 32; ++c)
        if (c4 * 32 + c < valid) m0 = fmaxf(m0, __uint_as_float(rc[c]));
    }
    if (c4 + 2 < nch) tmem_ld_32x32b_x32(trow + (c4 + 2) * 32, r[c4 & 1]);
  }
  const float m_new = fmaxf(fmaxf(max3(m0, m1, m2), m3), m_run);
  // rescale factor of everything accumulated so far; the first block has nothing to rescale (m_run = -inf)
  alpha = first ? 0.f : ex2_approx((m_run - m_new) * scale_log2);
  moved = !first && m_new > m_run;
  m_run = m_new;
  // ---- pass 2: P = exp2((s - m) c) -> bf16 -> TMEM
  tmem_ld_32x32b_x32(trow, r[0]);
  const float mc = m_new * scale_log2;
  const uint64_t sc2 = pack_f32x2(scale_log2, scale_log2), nmc2 = pack_f32x2(-mc, -mc);
  uint64_t l2a = pack_f32x2(0.f, 0.f), l2b = l2a;
#pragma unroll
  for (int c4 = 0; c4 < KB / 32; ++c4) {
    if (!FULL && c4 >= nch) break;
    tmem_ld_wait_fence(r[c4 & 1]);
    if (c4 + 1 < nch) tmem_ld_32x32b_x32(trow + (c4 + 1) * 32, r[(c4 + 1) & 1]);
    uint32_t pk[16];
    const bool full = FULL || (c4 + 1) * 32 <= valid;
#pragma unroll
    for (int c = 0; c < 16; ++c) {
      const uint64_t x2 =
          fma_f32x2(pack_f32x2(__uint_as_float(r[c4 & 1][2 * c]), __uint_as_float(r[c4 & 1][2 * c + 1])), sc2, nmc2);
      float p0, p1;
      if (ATT_LONG_POLY > 0 && (c * ATT_LONG_POLY) % 16 < ATT_LONG_POLY) {
        ex2_poly_x2(x2, p0, p1);
      } else {
        float x0, x1;
        unpack_f32x2(x2, x0, x1);
        p0 = ex2_approx(x0);
        p1 = ex2_approx(x1);
      }
      if (!full) {
        if (c4 * 32 + 2 * c >= valid) p0 = 0.f;
        if (c4 * 32 + 2 * c + 1 >= valid) p1 = 0.f;
      }
      if (c & 1) l2b = add_f32x2(l2b, pack_f32x2(p0, p1));
      else l2a = add_f32x2(l2a, pack_f32x2(p0, p1));
      pk[c] = pack_bf16x2(p0, p1);
    }
    tmem_st_32x32b_x16(trow + c4 * 16, pk);
  }
  float la, lb;
  unpack_f32x2(add_f32x2(l2a, l2b), la, lb);
  l = l * alpha + la + lb;
}

template <int KB>
__global__ void __launch_bounds__(LONG_THREADS, LongCfg<KB>::CTAS_PER_SM)
attention_tc_long_kernel(const __grid_constant__ CUtensorMap tmQK, const __grid_constant__ CUtensorMap tmKV,
                         const __grid_constant__ CUtensorMap tmO, int L, int heads, int tiles, int nkb, int num_items,
                         float scale_log2) {
  using S = LongCfg<KB>;
  constexpr int O_COL = S::O_COL;
  constexpr uint32_t TMEM_COLS = S::TMEM_COLS;
  constexpr int TILE_BYTES = S::KV_BYTES;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* q_full = reinterpret_cast<uint64_t*>(smem + S::OFF_BAR);
  uint64_t* k_full = q_full + 1;    // [2]
  uint64_t* k_empty = q_full + 3;   // [2]
  uint64_t* v_full = q_full + 5;    // [2]
  uint64_t* v_empty = q_full + 7;   // [2]
  uint64_t* s_full = q_full + 9;
  uint64_t* p_full = q_full + 10;
  uint64_t* o_full = q_full + 11;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(q_full + 12);
  constexpr int ST = S::STAGES;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int D = heads * HD;
  const int my_items = static_cast<int>(blockIdx.x) < num_items
                           ? (num_items - 1 - static_cast<int>(blockIdx.x)) / static_cast<int>(gridDim.x) + 1
                           : 0;

  griddep_launch_dependents();
  if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0) __trap();
  if (threadIdx.x == 128) {
    tma_prefetch_desc(&tmQK);
    tma_prefetch_desc(&tmKV);
    tma_prefetch_desc(&tmO);
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(k_full + i, 1);
      mbar_init(k_empty + i, 1);
      mbar_init(v_full + i, 1);
      mbar_init(v_empty + i, 1);
    }
    mbar_init(s_full, 1);
    mbar_init(p_full, 128);
    mbar_init(o_full, 1);
    fence_barrier_init();
  }
  if (warp == 4) tmem_alloc<TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  griddep_wait();

  if (warp == 4) {
    // ===================== TMA + MMA thread =====================
    if (lane == 0 && my_items > 0) {
      constexpr uint32_t idesc_pv = umma_idesc_bf16_f32_bmn(QT, HD);
      // key columns block j really needs: the keys that exist, rounded up to the MMA's N granularity (16).  The last
      // block of a 257-token sequence holds ONE key: S, softmax and P.V then cost 16 columns instead of 128.
      auto ncols_of = [&](int j) { return (min(KB, L - j * KB) + 15) & ~15; };
      const uint32_t q_addr = smem_u32(smem + S::OFF_Q);
      const int total_blocks = my_items * nkb;
      // block g of this CTA belongs to item blockIdx.x + (g / nkb) * gridDim.x, key block g % nkb
      auto coords = [&](int g, int& t, int& head, int& seq, int& j) {
        const int item = static_cast<int>(blockIdx.x) + (g / nkb) * static_cast<int>(gridDim.x);
        j = g % nkb;
        t = item % tiles;
        const int sh = item / tiles;
        head = sh % heads;
        seq = sh / heads;
      };
      // K and V have their own full / empty barriers: K_g is free once S_g has completed, V_g once P_g.V_g has -- with one
      // stage (KB = 192) K_{g+1} streams in under softmax(g) and V_{g+1} under S_{g+1} + softmax(g+1)
      auto load_k = [&](int g) {
        int t, head, seq, j;
        coords(g, t, head, seq, j);
        const int st = g % ST;
        mbar_wait(k_empty + st, ((g / ST) & 1) ^ 1);  // first use of a stage: passes at once
        mbar_expect_tx(k_full + st, TILE_BYTES);
        tma_load_3d(smem + S::OFF_K + st * TILE_BYTES, &tmKV, k_full + st, D + head * HD, j * KB, seq);
      };
      auto load_v = [&](int g) {
        int t, head, seq, j;
        coords(g, t, head, seq, j);
        const int st = g % ST;
        mbar_wait(v_empty + st, ((g / ST) & 1) ^ 1);
        mbar_expect_tx(v_full + st, TILE_BYTES);
        tma_load_3d(smem + S::OFF_V + st * TILE_BYTES, &tmKV, v_full + st, 2 * D + head * HD, j * KB, seq);
      };
      auto issue_pv = [&](int g, bool accumulate) {  // O (+)= P_g . V_g
        const int st = g % ST;
        mbar_wait(v_full + st, (g / ST) & 1);
        tc_fence_after();
        const uint32_t v_addr = smem_u32(smem + S::OFF_V + st * TILE_BYTES);
        const int nk = ncols_of(g % nkb) / 16;
        if (nk == KB / 16) {
#pragma unroll
          for (int k = 0; k < KB / 16; ++k)
            umma_bf16_ts(tmem_base + O_COL, tmem_base + k * 8, umma_desc_mn_sw128(v_addr + k * 2048), idesc_pv,
                         accumulate || k != 0);
        } else {
          for (int k = 0; k < nk; ++k)
            umma_bf16_ts(tmem_base + O_COL, tmem_base + k * 8, umma_desc_mn_sw128(v_addr + k * 2048), idesc_pv,
                         accumulate || k != 0);
        }
        umma_commit(v_empty + st);  // V_g may be overwritten once these MMAs are done
      };
      load_k(0);
      load_v(0);
      for (int g = 0; g < total_blocks; ++g) {
        int t, head, seq, j;
        coords(g, t, head, seq, j);
        const int st = g % ST;
        if (j == 0) {
          // Q of the previous item is free: its last S MMA completed before the softmax threads arrived on p_full, and
          // that arrival was waited for below before this point was reached
          mbar_expect_tx(q_full, Q_BYTES);
          tma_load_3d(smem + S::OFF_Q, &tmQK, q_full, head * HD, t * QT, seq);
        }
        mbar_wait(k_full + st, (g / ST) & 1);
        if (j == 0) mbar_wait(q_full, (g / nkb) & 1);
        if (j > 0) {
          mbar_wait(p_full, (g - 1) & 1);  // P_{g-1} is in TMEM and O has been rescaled
          tc_fence_after();
          issue_pv(g - 1, j - 1 > 0);
        }
        tc_fence_after();
        const uint32_t k_addr = smem_u32(smem + S::OFF_K + st * TILE_BYTES);
        const uint32_t idesc_s = umma_idesc_bf16_f32(QT, ncols_of(j));
#pragma unroll
        for (int k = 0; k < HD / 16; ++k)
          umma_bf16_ss(tmem_base, umma_desc_k_sw128(q_addr + k * 32), umma_desc_k_sw128(k_addr + k * 32), idesc_s, k != 0);
        umma_commit(s_full);
        umma_commit(k_empty + st);  // K_g may be overwritten once S_g is done
        // Loads for later blocks.  Each wait on an empty barrier below is for MMAs that have already been ISSUED (P.V of
        // block g - 1 above or in the previous iteration's tail, S of block g just now), never for one still to come.
        if (ST == 1) {
          if (g > 0) load_v(g);                     // the V stage was released by P_{g-1}.V_{g-1}
          if (g + 1 < total_blocks) load_k(g + 1);  // the K stage is released by S_g
        } else if (g + 1 < total_blocks) {
          load_k(g + 1);  // stage (g + 1) % 2 was used by block g - 1
          load_v(g + 1);
        }
        if (j == nkb - 1) {  // last key block of the item
          mbar_wait(p_full, g & 1);
          tc_fence_after();
          issue_pv(g, nkb > 1);
          umma_commit(o_full);
        }
      }
    }
  } else {
    // ===================== softmax + epilogue warps (one query row per thread) =====================
    const uint32_t trow = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
    uint8_t* stg_ptr = smem + S::OFF_STG + warp * (32 * 128);
    const uint32_t stg_row = smem_u32(stg_ptr) + lane * 128;
    const int sw = lane & 7;
    int g = 0;
    for (int n = 0; n < my_items; ++n) {
      const int item = static_cast<int>(blockIdx.x) + n * static_cast<int>(gridDim.x);
      const int t = item % tiles;
      const int sh = item / tiles;
      const int head = sh % heads, seq = sh / heads;
      const int row0 = t * QT + warp * 32;
      const bool active = row0 < L;  // warp-uniform
      float m_run = -INFINITY, l = 0.f;
      for (int j = 0; j < nkb; ++j, ++g) {
        mbar_wait(s_full, g & 1);
        tc_fence_after();
        if (active) {
          const int valid = min(KB, L - j * KB);  // keys of this block that exist (>= 1)
          float alpha;
          bool moved;
          if (valid == KB) softmax_block<KB, true>(trow, valid, j == 0, scale_log2, m_run, l, alpha, moved);
          else softmax_block<KB, false>(trow, valid, j == 0, scale_log2, m_run, l, alpha, moved);
          // ---- O <- a O where a row's maximum moved (s_full of this block implies P_{j-1}.V_{j-1} has completed)
          if (__any_sync(0xffffffffu, moved)) {
            const uint64_t a2 = pack_f32x2(alpha, alpha);
#pragma unroll
            for (int hh = 0; hh < 4; ++hh) {
              uint32_t o[16];
              tmem_ld_32x32b_x16(trow + O_COL + 16 * hh, o);
              tmem_ld_wait_fence16(o);
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                float a, b;
                unpack_f32x2(mul_f32x2(pack_f32x2(__uint_as_float(o[2 * e]), __uint_as_float(o[2 * e + 1])), a2), a, b);
                o[2 * e] = __float_as_uint(a);
                o[2 * e + 1] = __float_as_uint(b);
              }
              tmem_st_32x32b_x16(trow + O_COL + 16 * hh, o);
            }
          }
          tmem_st_wait();
        }
        tc_fence_before();
        mbar_arrive(p_full);
      }
      // ---- epilogue: O / l -> bf16 -> staging -> TMA store
      mbar_wait(o_full, n & 1);
      tc_fence_after();
      if (active) {
        const float inv = 1.f / l;
        const uint64_t inv2 = pack_f32x2(inv, inv);
        if (lane == 0) bulk_wait_group_read<0>();
        __syncwarp();
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          uint32_t o[32];
          tmem_ld_32x32b_x32(trow + O_COL + 32 * hh, o);
          tmem_ld_wait_fence(o);
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            float v[8];
#pragma unroll
            for (int e = 0; e < 4; ++e)
              unpack_f32x2(mul_f32x2(pack_f32x2(__uint_as_float(o[8 * c + 2 * e]), __uint_as_float(o[8 * c + 2 * e + 1])), inv2),
                           v[2 * e], v[2 * e + 1]);
            st_shared_v4(stg_row + (((4 * hh + c) ^ sw) << 4),
                         make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]),
                                    pack_bf16x2(v[6], v[7])));
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_3d(&tmO, stg_ptr, head * HD, row0, seq);  // rows >= L are clipped by the tensor map
          bulk_commit_group();
        }
      }
      // the next item's first P.V (accumulate = 0) overwrites O: it is issued only after every softmax thread has arrived on
      // p_full for that item's first block, i.e. after this read-out
    }
    if (lane == 0) bulk_wait_group<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

// bf16 tensor viewed as [seqs][L][cols] (cols contiguous); box = 64 columns x box_rows tokens x 1 sequence, SW128.
int make_tmap_3d(CUtensorMap* tm, const bf16* base, int64_t cols, int64_t L, int64_t seqs, int box_rows) {
  const uint64_t dims[3] = {static_cast<uint64_t>(cols), static_cast<uint64_t>(L), static_cast<uint64_t>(seqs)};
  const uint64_t strides[2] = {static_cast<uint64_t>(cols) * 2, static_cast<uint64_t>(cols) * 2 * L};
  const uint32_t box[3] = {64, static_cast<uint32_t>(box_rows), 1};
  return tmap_bf16_sw128(tm, base, 3, dims, strides, box);
}

template <int KB>
int launch_long(const bf16* qkv, bf16* out, int64_t seqs, int L, int heads, cudaStream_t s) {
  using Cfg = LongCfg<KB>;
  static bool configured = false;
  if (!configured) {
    FC_CUDA(cudaFuncSetAttribute(attention_tc_long_kernel<KB>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::BYTES));
    configured = true;
  }
  const int D = heads * HD;
  CUtensorMap tqk, tkv, to;
  int rc;
  if ((rc = make_tmap_3d(&tqk, qkv, 3 * D, L, seqs, 128))) return rc;  // Q tiles: 128 rows x 64 columns
  if ((rc = make_tmap_3d(&tkv, qkv, 3 * D, L, seqs, KB))) return rc;   // K / V blocks: KB rows x 64 columns
  if ((rc = make_tmap_3d(&to, out, D, L, seqs, 32))) return rc;
  const int tiles = (L + QT - 1) / QT, nkb = (L + KB - 1) / KB;
  const int64_t items64 = seqs * heads * tiles;
  FC_REQUIRE(items64 * nkb < (int64_t(1) << 30), "attention: too many work items");
  const int items = static_cast<int>(items64);
  int grid = Cfg::CTAS_PER_SM * num_sms();
  if (grid > items) grid = items;
  const float scale_log2 = 0.125f * 1.4426950408889634f;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(LONG_THREADS);
  cfg.dynamicSmemBytes = Cfg::BYTES;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  note_launch();
  FC_CUDA(cudaLaunchKernelEx(&cfg, attention_tc_long_kernel<KB>, tqk, tkv, to, L, heads, tiles, nkb, items, scale_log2));
  return FC_OK;
}

}  // namespace

// Un-masked sequences of 209..768 tokens. *handled = 1 when it took the call.
int attention_bf16_tc_long(const bf16* qkv, bf16* out, int64_t seqs, int L, int heads, int causal, cudaStream_t s,
                           int* handled) {
  *handled = 0;
  static int disabled = -1, forced = 0;
  if (disabled < 0) {
    // diagnostics: FC_ATTENTION=mma forces the mma.sync kernels, FC_ATTENTION=long this kernel for every un-masked length
    const char* e = getenv("FC_ATTENTION");
    disabled = (e && strcmp(e, "mma") == 0) ? 1 : 0;
    forced = (e && strcmp(e, "long") == 0) ? 1 : 0;
  }
  if (disabled || causal || (L <= 208 && !forced) || L > 768) return FC_OK;
  FC_REQUIRE((reinterpret_cast<uintptr_t>(qkv) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
             "attention: buffers must be 16-byte aligned");
  *handled = 1;
  // 192-key blocks wherever they save a block (every length above 256); FC_ATT_LONG_KB=128|192 overrides for measurements
  static int kb_env = -1;
  if (kb_env < 0) {
    const char* e = getenv("FC_ATT_LONG_KB");
    kb_env = e ? atoi(e) : 0;
  }
  const int kb = kb_env == 128 || kb_env == 192 ? kb_env : ((L + 191) / 192 < (L + 127) / 128 ? 192 : 128);
  return kb == 192 ? launch_long<192>(qkv, out, seqs, L, heads, s) : launch_long<128>(qkv, out, seqs, L, heads, s);
}

}  // namespace fc
