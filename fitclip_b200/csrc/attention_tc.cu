// tcgen05 fused attention for the CLIP image sequence (197 tokens, no mask) and text sequence (77 tokens, causal),
// head dim 64 -- K4 of SURVEY.md 2.2;
// reference: nn.MultiheadAttention(need_weights=False) in ResidualAttentionBlock (twin aligner/encoder/slip.py:378-380).
//
// One work item = (sequence, head, 128-row query tile); persistent CTAs (2 per SM, 256 TMEM columns each) loop over
// items.  Per item:
//   TMA      Q tile [128 x 64], K [KP x 64], V [KP x 64] through 3-D tensor maps (dims: column, token-in-sequence,
//            sequence) so tokens beyond the sequence are ZERO-FILLED on load and CLIPPED on store -- a tile never
//            touches its neighbour sequence.
//   MMA      S = Q K^T      tcgen05.mma M=128 K=16 x4, both operands K-major (128B swizzle), in TWO column groups:
//            "main" keys [0, KP-16) -> TMEM cols [0, KP-16) and "tail" keys [KP-16, KP) -> cols [KP-16, KP)
//   softmax  4 warps, one query row per thread: two passes over the row in TMEM (max; exp2 + sum), P rounded to bf16
//            and written BACK INTO TMEM over the dead S columns (tcgen05.st) -- P never touches shared memory
//   MMA      O = P V        tcgen05.mma M=128 N=64 K=16 x KP/16, A = P from TMEM, B = V MN-major from smem -> TMEM
//            cols [KP-16, KP+48): over the (dead) tail columns and the rest of the 256-column allocation
//   epilogue O / rowsum -> bf16 -> per-warp 32x64 staging tile -> 3-D TMA store (ATT_DIRECT_STORE=1: plain 16-byte
//            global stores instead -- measured 30 % slower)
// Pipelining inside a CTA: the main part of S(i+1) is issued right behind P(i).V (tensor-pipe order guarantees P(i) has
// been consumed before it is overwritten) and does not touch the O(i) columns, so it runs while the softmax warps read
// O(i) and store it; only the 16-column tail of S(i+1) has to wait for O(i) to be drained.  The TMA/MMA thread
// prefetches the next item's Q,K as soon as S is done and V as soon as O is done; the second co-resident CTA overlaps
// what is left of the MMA latency with this CTA's MUFU-bound softmax.
#include <stdlib.h>
#include <string.h>

#include "kernels.cuh"
#include "ptx.cuh"
#include "tmap.cuh"

namespace fc {

namespace {

constexpr int HD = 64;
constexpr int QT = 128;                    // query rows per tile
constexpr int Q_BYTES = QT * 128;          // 16 KiB
constexpr int STG_BYTES = 4 * 32 * 128;    // 4 warps x (32 rows x 128 B) output staging (TMA-store epilogue)
// tuning switches (kept as macros so variants can be built side by side: make VARIANT=x EXTRA=-DATT_...=v)
#ifndef ATT_DIRECT_STORE
#define ATT_DIRECT_STORE 0  // 1: epilogue writes O rows straight to global memory instead of staging + TMA store
#endif
#ifndef ATT_PHALF
#define ATT_PHALF 2  // early hand-overs of P to the tensor core while the softmax still works on the row: 0 = none,
                     // 1 = after 96 keys, 2 = after 64 and after 128 keys
#endif
#ifndef ATT_KNOCKOUT
#define ATT_KNOCKOUT 0  // timing experiments only (results are WRONG): 1 = no exp2, 2 = no row-max pass, 4 = no P store,
                       // 8 = no pass 2 at all, 16 = no epilogue
#endif
constexpr int ATT_THREADS = 160;           // warps 0-3 softmax/epilogue, warp 4 TMA + MMA + TMEM alloc
// ---- second-generation kernel (attention_tc2_kernel): softmax warpgroup + EPILOGUE warpgroup + TMA/MMA warp
constexpr int ATT2_THREADS = 288;          // warps 0-3 softmax, warps 4-7 epilogue (same lane quadrants), warp 8 TMA + MMA
#ifndef ATT2_POLY
#define ATT2_POLY 0  // of every 16 column pairs of a x32 chunk, this many take exp2 on the FMA pipe (degree-3 polynomial,
                     // Cody-Waite split, packed fp32x2) instead of the MUFU: 0 = none, 4 = 25 %, 8 = 50 %
#endif
#ifndef ATT2_REGS_SOFTMAX
// setmaxnreg targets (multiples of 8).  Registers move inside a CTA's own launch allocation, so
// 128 * SOFTMAX + 128 * EPILOGUE + 32 * (launch count, kept by the TMA/MMA warp) must not exceed 288 * (registers per
// thread at launch) -- checked on the host
// against cudaFuncGetAttributes before the first launch (an unsatisfiable setmaxnreg.inc would spin forever).
#define ATT2_REGS_SOFTMAX 120
#define ATT2_REGS_EPILOGUE 72
#endif

template <int KP>
struct AttSmem {
  static constexpr int KV_BYTES = KP * 128;
  static constexpr int OFF_Q = 0;
  static constexpr int OFF_K = Q_BYTES;
  static constexpr int OFF_V = OFF_K + ((KV_BYTES + 1023) / 1024) * 1024;
  static constexpr int OFF_STG = OFF_V + ((KV_BYTES + 1023) / 1024) * 1024;
  static constexpr int OFF_BAR = OFF_STG + (ATT_DIRECT_STORE ? 0 : STG_BYTES);
  static constexpr int BYTES = OFF_BAR + 128;
};

#ifdef FC_GEMM_TIMING
// Diagnostics build only (make TIMING=1): cycles of softmax warp 0, summed over CTAs and items.
//   [0] whole loop  [1] wait S main  [2] pass 1  [3] wait S tail  [4] pass 2  [5] wait O  [6] epilogue  [7] items
__device__ unsigned long long g_att_timing[8];
#define FC_T(...) __VA_ARGS__
#else
#define FC_T(...)
#endif

// Work-item cursor: item -> (sequence, head, query tile) without a division per item (the grid stride is decomposed once).
struct ItemCursor {
  int t, head, seq;
  int dt, dhead, dseq, tiles, heads;
  __device__ __forceinline__ void init(int item, int step, int tiles_, int heads_) {
    tiles = tiles_;
    heads = heads_;
    t = item % tiles;
    const int sh = item / tiles;
    head = sh % heads;
    seq = sh / heads;
    dt = step % tiles;
    const int dsh = step / tiles;
    dhead = dsh % heads;
    dseq = dsh / heads;
  }
  __device__ __forceinline__ void advance() {
    t += dt;
    int c = t >= tiles ? 1 : 0;
    t -= c * tiles;
    head += dhead + c;
    c = head >= heads ? 1 : 0;
    head -= c * heads;
    seq += dseq + c;
  }
};

// KP: keys padded to whole x32 chunks + one 16-key tail (208 for the 197-token image sequence, 80 for the 77-token
// text sequence).  CAUSAL: query row i attends to keys 0..i (aligner/encoder/slip.py:454-460).  NPH: early hand-overs
// of P (see ATT_PHALF).
template <int KP, bool CAUSAL, int NPH>
__global__ void __launch_bounds__(ATT_THREADS, 2)
attention_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                    const __grid_constant__ CUtensorMap tmO, bf16* __restrict__ out, int L, int heads, int tiles,
                    int num_items, float scale_log2) {
  using S = AttSmem<KP>;
  static_assert(KP % 32 == 16 && KP >= 48 && KP <= 208, "padded key count: whole x32 chunks plus one 16-key tail");
  constexpr int KMAIN = KP - 16;  // keys / S columns of the main group
  constexpr int O_COL = KMAIN;    // O accumulator columns [KP-16, KP+48): the tail of S and the columns behind it
  constexpr int KHALF = NPH == 2 ? 64 : 96;  // keys per early hand-over (a whole number of x32 chunks)
  constexpr uint32_t TMEM_COLS = O_COL + HD <= 128 ? 128 : 256;
  static_assert(KP / 2 <= O_COL && O_COL + HD <= 256, "P / O column ranges must not overlap");
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* kq_full = reinterpret_cast<uint64_t*>(smem + S::OFF_BAR);
  uint64_t* v_full = kq_full + 1;
  uint64_t* s_full = kq_full + 2;
  uint64_t* p_full = kq_full + 3;
  uint64_t* o_full = kq_full + 4;
  uint64_t* o_empty = kq_full + 5;
  uint64_t* t_full = kq_full + 6;  // tail columns of S
  uint64_t* p_half = kq_full + 7;  // [2] P of the first KHALF / 2 * KHALF keys is in TMEM
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(kq_full + 9);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int D = heads * HD;
  const bool flip_ok = (gridDim.x & 1) == 0 && tiles == 2;  // alternate heavy/light tiles between iterations

  griddep_launch_dependents();
  if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0) __trap();
  if (threadIdx.x == 128) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmKV);
    if (!ATT_DIRECT_STORE) tma_prefetch_desc(&tmO);
    mbar_init(kq_full, 1);
    mbar_init(v_full, 1);
    mbar_init(s_full, 1);
    mbar_init(p_full, 128);
    mbar_init(o_full, 1);
    mbar_init(o_empty, 128);
    mbar_init(t_full, 1);
    mbar_init(p_half, 128);
    mbar_init(p_half + 1, 128);
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
    if (lane == 0) {
      constexpr uint32_t idesc_s = umma_idesc_bf16_f32(QT, KMAIN);
      constexpr uint32_t idesc_t = umma_idesc_bf16_f32(QT, 16);
      constexpr uint32_t idesc_pv = umma_idesc_bf16_f32_bmn(QT, HD);
      const uint32_t q_addr = smem_u32(smem + S::OFF_Q);
      const uint32_t k_addr = smem_u32(smem + S::OFF_K);
      const uint32_t v_addr = smem_u32(smem + S::OFF_V);
      // `n`: ordinal of the item within this CTA (decides the heavy/light flip)
      auto load_qk = [&](const ItemCursor& c, int n) {
        const int tile = flip_ok ? (c.t ^ (n & 1)) : c.t;
        mbar_expect_tx(kq_full, Q_BYTES + S::KV_BYTES);
        tma_load_3d(smem + S::OFF_Q, &tmQ, kq_full, c.head * HD, tile * QT, c.seq);
        tma_load_3d(smem + S::OFF_K, &tmKV, kq_full, D + c.head * HD, 0, c.seq);
      };
      auto load_v = [&](const ItemCursor& c) {
        mbar_expect_tx(v_full, S::KV_BYTES);
        tma_load_3d(smem + S::OFF_V, &tmKV, v_full, 2 * D + c.head * HD, 0, c.seq);
      };
      auto issue_s_main = [&]() {
#pragma unroll
        for (int k = 0; k < HD / 16; ++k)
          umma_bf16_ss(tmem_base, umma_desc_k_sw128(q_addr + k * 32), umma_desc_k_sw128(k_addr + k * 32), idesc_s,
                       k != 0);
        umma_commit(s_full);
      };
      auto issue_s_tail = [&]() {  // keys [KMAIN, KP): K rows from byte offset KMAIN * 128 (a whole number of 8-row atoms)
#pragma unroll
        for (int k = 0; k < HD / 16; ++k)
          umma_bf16_ss(tmem_base + KMAIN, umma_desc_k_sw128(q_addr + k * 32),
                       umma_desc_k_sw128(k_addr + KMAIN * 128 + k * 32), idesc_t, k != 0);
        umma_commit(t_full);
      };
      ItemCursor cv, cqk;  // the items whose V / whose Q,K are loaded next
      cv.init(blockIdx.x, gridDim.x, tiles, heads);
      cqk = cv;
      if (static_cast<int>(blockIdx.x) < num_items) {
        load_qk(cqk, 0);
        load_v(cv);
        cqk.advance();
        cv.advance();
        mbar_wait(kq_full, 0);
        tc_fence_after();
        issue_s_main();
        issue_s_tail();
        mbar_wait(t_full, 0);  // Q and K tiles are free again
        if (static_cast<int>(blockIdx.x + gridDim.x) < num_items) load_qk(cqk, 1);
        cqk.advance();
      }
      int it = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++it) {
        const uint32_t ph = it & 1;
        const int next = item + gridDim.x, next2 = next + gridDim.x;
        const bool has_next = next < num_items;
        mbar_wait(v_full, ph);
#pragma unroll
        for (int part = 0; part < NPH; ++part) {
          mbar_wait(p_half + part, ph);  // P of the next KHALF keys is in TMEM: O += P.V while the softmax goes on
          tc_fence_after();
#pragma unroll
          for (int k = part * (KHALF / 16); k < (part + 1) * (KHALF / 16); ++k)
            umma_bf16_ts(tmem_base + O_COL, tmem_base + k * 8, umma_desc_mn_sw128(v_addr + k * 2048), idesc_pv, k != 0);
        }
        mbar_wait(p_full, ph);  // all of P is in TMEM
        tc_fence_after();
#pragma unroll
        for (int k = NPH * (KHALF / 16); k < KP / 16; ++k)
          umma_bf16_ts(tmem_base + O_COL, tmem_base + k * 8, umma_desc_mn_sw128(v_addr + k * 2048), idesc_pv, k != 0);
        umma_commit(o_full);
        if (has_next) {
          // main part of S(it+1): queued right behind P.V (which consumes P before these MMAs overwrite it); it leaves
          // the O columns alone, so it overlaps the softmax warps' read-out of O(it)
          mbar_wait(kq_full, ph ^ 1);
          tc_fence_after();
          issue_s_main();
        }
        mbar_wait(o_full, ph);  // V tile is free again
        if (has_next) {
          load_v(cv);
          cv.advance();
          mbar_wait(o_empty, ph);  // O(it) has been read out: the tail columns may be overwritten
          tc_fence_after();
          issue_s_tail();
          mbar_wait(t_full, ph ^ 1);  // Q and K tiles are free again
          if (next2 < num_items) load_qk(cqk, it + 2);
          cqk.advance();
        }
      }
    }
  } else {
    // ===================== softmax + epilogue warps (one query row per thread) =====================
    const uint32_t trow = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
    constexpr int NFULL = KP / 32;  // x32 chunks (main columns); the 16-column tail follows
    static_assert(NPH == 0 || (KHALF % 32 == 0 && NPH * (KHALF / 32) < NFULL), "early hand-overs must end on chunk boundaries");
    ItemCursor cur;
    cur.init(blockIdx.x, gridDim.x, tiles, heads);
    int it = 0;
    FC_T(long long tq[7] = {0, 0, 0, 0, 0, 0, 0}; long long n_it = 0; const long long t_begin = clock64();)
    for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++it, cur.advance()) {
      const uint32_t ph = it & 1;
      const int tile = flip_ok ? (cur.t ^ (it & 1)) : cur.t;
      const int row0 = tile * QT + warp * 32;
      const bool active = row0 < L;  // warp-uniform: warps whose 32 rows all lie beyond the sequence only sync
      // keys this thread's query row may attend to: [0, lim).  Without a mask lim = L for every row (warp-uniform, the
      // branches below stay uniform); causal rows see keys 0..row.
      const int lim = CAUSAL ? min(L, row0 + lane + 1) : L;
      FC_T(long long t0 = clock64(); long long t1;)
      mbar_wait(s_full, ph);
      FC_T(t1 = clock64(); tq[1] += t1 - t0; t0 = t1; ++n_it;)
      tc_fence_after();
      float l = 0.f;
      float m = -INFINITY;
      if (active && !(ATT_KNOCKOUT & 2)) {
        // ---- pass 1 (main columns): row maximum; four independent FMNMX3 chains, two columns per instruction.  The
        // loop is TMEM-load latency bound, so two x32 chunks are kept in flight (a third costs spills under the 168-register cap).
        float m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
        constexpr int DEPTH = 2;
        uint32_t r[DEPTH][32];
#pragma unroll
        for (int j = 0; j < DEPTH && j < NFULL; ++j) tmem_ld_32x32b_x32(trow + j * 32, r[j]);
#pragma unroll
        for (int j = 0; j < NFULL; ++j) {
          // loads complete in issue order; waiting for all outstanding ones is exact for the oldest
          tmem_ld_wait_fence(r[j % DEPTH]);
          const uint32_t(&rc)[32] = r[j % DEPTH];
          if ((j + 1) * 32 <= lim) {
#pragma unroll
            for (int c = 0; c < 32; c += 8) {
              m = max3(m, __uint_as_float(rc[c]), __uint_as_float(rc[c + 1]));
              m1 = max3(m1, __uint_as_float(rc[c + 2]), __uint_as_float(rc[c + 3]));
              m2 = max3(m2, __uint_as_float(rc[c + 4]), __uint_as_float(rc[c + 5]));
              m3 = max3(m3, __uint_as_float(rc[c + 6]), __uint_as_float(rc[c + 7]));
            }
          } else {
#pragma unroll
            for (int c = 0; c < 32; ++c)
              if (j * 32 + c < lim) m = fmaxf(m, __uint_as_float(rc[c]));
          }
          if (j + DEPTH < NFULL) tmem_ld_32x32b_x32(trow + (j + DEPTH) * 32, r[j % DEPTH]);
        }
        m = fmaxf(max3(m, m1, m2), m3);
      }
      FC_T(t1 = clock64(); tq[2] += t1 - t0; t0 = t1;)
      mbar_wait(t_full, ph);  // the 16 tail columns arrive a little later (they share TMEM columns with the previous O)
      FC_T(t1 = clock64(); tq[3] += t1 - t0; t0 = t1;)
      tc_fence_after();
      if (active && !(ATT_KNOCKOUT & 8)) {
        // ---- pass 2: P = exp2((s - m) * scale * log2e), row sum in fp32, P -> bf16 -> TMEM (over the S columns).
        // The tail goes FIRST and its P waits in registers: the early P.V below writes O over the tail's S columns.
        uint32_t r16[16];
        tmem_ld_32x32b_x16(trow + NFULL * 32, r16);
        uint32_t r[2][32];
        tmem_ld_32x32b_x32(trow, r[0]);
        tmem_ld_wait_fence16(r16);
#pragma unroll
        for (int c = 0; c < 16; ++c)
          if (NFULL * 32 + c < lim) m = fmaxf(m, __uint_as_float(r16[c]));
        const float mc = m * scale_log2;
        const uint64_t sc2 = pack_f32x2(scale_log2, scale_log2), nmc2 = pack_f32x2(-mc, -mc);
        uint64_t l2a = pack_f32x2(0.f, 0.f), l2b = l2a;  // packed partial row sums (FADD2), two chains
        // Causal kernel only: lanes whose row lies beyond the sequence or whose keys in a chunk are all masked skip the
        // arithmetic of that chunk (the MUFU works through a warp's ACTIVE lanes): 65.5 -> 63.9 us on the text
        // sequence.  The same branches cost the un-masked 197-token kernel 16 % (73.1 -> 84.8 us: the chunk bodies no
        // longer interleave with the TMEM loads), so there the conditions are compile-time true.
        const bool row_valid = !CAUSAL || row0 + lane < L;
        uint32_t pk_tail[8];
        if (row_valid && (!CAUSAL || NFULL * 32 < lim)) {
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            float x0, x1;
            unpack_f32x2(fma_f32x2(pack_f32x2(__uint_as_float(r16[2 * c]), __uint_as_float(r16[2 * c + 1])), sc2, nmc2),
                         x0, x1);
            float p0 = (ATT_KNOCKOUT & 1) ? x0 : ex2_approx(x0);
            float p1 = (ATT_KNOCKOUT & 1) ? x1 : ex2_approx(x1);
            if (NFULL * 32 + 2 * c >= lim) p0 = 0.f;
            if (NFULL * 32 + 2 * c + 1 >= lim) p1 = 0.f;
            l2b = add_f32x2(l2b, pack_f32x2(p0, p1));
            pk_tail[c] = pack_bf16x2(p0, p1);
          }
        } else {
#pragma unroll
          for (int c = 0; c < 8; ++c) pk_tail[c] = 0u;
        }
#pragma unroll
        for (int j = 0; j < NFULL; ++j) {
          tmem_ld_wait_fence(r[j & 1]);
          if (j + 1 < NFULL) tmem_ld_32x32b_x32(trow + (j + 1) * 32, r[(j + 1) & 1]);
          uint32_t pk[16];
          const bool full = (j + 1) * 32 <= lim;
          if (row_valid && (!CAUSAL || j * 32 < lim)) {
#pragma unroll
            for (int c = 0; c < 16; ++c) {
              float x0, x1;
              unpack_f32x2(fma_f32x2(pack_f32x2(__uint_as_float(r[j & 1][2 * c]), __uint_as_float(r[j & 1][2 * c + 1])),
                                     sc2, nmc2),
                           x0, x1);
              float p0 = (ATT_KNOCKOUT & 1) ? x0 : ex2_approx(x0);
              float p1 = (ATT_KNOCKOUT & 1) ? x1 : ex2_approx(x1);
              if (!full) {
                if (j * 32 + 2 * c >= lim) p0 = 0.f;
                if (j * 32 + 2 * c + 1 >= lim) p1 = 0.f;
              }
              if (c & 1) l2b = add_f32x2(l2b, pack_f32x2(p0, p1));
              else l2a = add_f32x2(l2a, pack_f32x2(p0, p1));
              pk[c] = pack_bf16x2(p0, p1);
            }
          } else {
#pragma unroll
            for (int c = 0; c < 16; ++c) pk[c] = 0u;
          }
          if (!(ATT_KNOCKOUT & 4)) tmem_st_32x32b_x16(trow + j * 16, pk);
          else asm volatile("" ::"r"(pk[0]), "r"(pk[5]), "r"(pk[11]), "r"(pk[15]));
          if (NPH > 0 && (j + 1) % (KHALF / 32) == 0 && (j + 1) / (KHALF / 32) <= NPH) {
            tmem_st_wait();  // KHALF more keys of P are complete: let the tensor core start on them
            tc_fence_before();
            mbar_arrive(p_half + (j + 1) / (KHALF / 32) - 1);
          }
        }
        tmem_st_32x32b_x8(trow + NFULL * 16, pk_tail);
        tmem_st_wait();
        float la, lb;
        unpack_f32x2(add_f32x2(l2a, l2b), la, lb);
        l = la + lb;
      } else {
        l = 1.f;
#pragma unroll
        for (int part = 0; part < NPH; ++part) mbar_arrive(p_half + part);
      }
      tc_fence_before();
      mbar_arrive(p_full);
      FC_T(t1 = clock64(); tq[4] += t1 - t0; t0 = t1;)

      // ---- epilogue: O / l -> bf16 -> global (every thread owns one 128-byte row segment: full lines, no staging)
      mbar_wait(o_full, ph);
      FC_T(t1 = clock64(); tq[5] += t1 - t0; t0 = t1;)
      tc_fence_after();
      uint32_t o0[32], o1[32];
      if (active && !(ATT_KNOCKOUT & 16)) {
        tmem_ld_32x32b_x32(trow + O_COL, o0);
        tmem_ld_32x32b_x32(trow + O_COL + 32, o1);
        tmem_ld_wait_fence(o0);
        tmem_ld_wait_fence(o1);
      }
      tc_fence_before();
      mbar_arrive(o_empty);
      const int row = row0 + lane;
      if ((ATT_DIRECT_STORE ? (active && row < L) : active) && !(ATT_KNOCKOUT & 16)) {
        const float inv = 1.f / l;
        const uint64_t inv2 = pack_f32x2(inv, inv);
        uint4* dst = reinterpret_cast<uint4*>(out + (static_cast<int64_t>(cur.seq) * L + row) * D + cur.head * HD);
        uint8_t* stg_ptr = smem + S::OFF_STG + warp * (32 * 128);
        const uint32_t stg_row = smem_u32(stg_ptr) + lane * 128;
        const int sw = lane & 7;
        if (!ATT_DIRECT_STORE) {
          if (lane == 0) bulk_wait_group_read<0>();  // this warp's previous store has finished reading its staging tile
          __syncwarp();
        }
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const uint32_t(&rr)[32] = c < 4 ? o0 : o1;
          const int o = (c & 3) * 8;
          float v[8];
#pragma unroll
          for (int e = 0; e < 4; ++e)
            unpack_f32x2(mul_f32x2(pack_f32x2(__uint_as_float(rr[o + 2 * e]), __uint_as_float(rr[o + 2 * e + 1])), inv2),
                         v[2 * e], v[2 * e + 1]);
          uint4 u;
          u.x = pack_bf16x2(v[0], v[1]);
          u.y = pack_bf16x2(v[2], v[3]);
          u.z = pack_bf16x2(v[4], v[5]);
          u.w = pack_bf16x2(v[6], v[7]);
          if (ATT_DIRECT_STORE) dst[c] = u;
          else st_shared_v4(stg_row + ((c ^ sw) << 4), u);
        }
        if (!ATT_DIRECT_STORE) {
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_3d(&tmO, stg_ptr, cur.head * HD, row0, cur.seq);  // rows >= L are clipped by the tensor map
            bulk_commit_group();
          }
        }
      }
      FC_T(t1 = clock64(); tq[6] += t1 - t0;)
    }
    FC_T(if (threadIdx.x == 0) {
      atomicAdd(&g_att_timing[0], static_cast<unsigned long long>(clock64() - t_begin));
      for (int i = 1; i < 7; ++i) atomicAdd(&g_att_timing[i], static_cast<unsigned long long>(tq[i]));
      atomicAdd(&g_att_timing[7], static_cast<unsigned long long>(n_it));
    })
    if (!ATT_DIRECT_STORE && lane == 0) bulk_wait_group<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}


template <int N>
__device__ __forceinline__ void setmaxnreg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N>
__device__ __forceinline__ void setmaxnreg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }

template <int KP>
struct Att2Smem {
  using B = AttSmem<KP>;
  static constexpr int OFF_L = B::OFF_BAR + 128;  // float[2][128]: row sums handed from the softmax to the epilogue warps
  static constexpr int BYTES = OFF_L + 2 * 128 * 4;
};

// Second-generation kernel for the un-masked image sequence: the work split of attention_tc_kernel plus a dedicated
// EPILOGUE warpgroup.  In the first kernel a softmax warp spends a third of every item waiting for P.V and reading out /
// scaling / storing O (1570 of 6100 cycles) with only two warps per scheduler to hide it; here warps 4-7 (the same TMEM
// lane quadrants as warps 0-3) do that, the softmax warps go straight from the last P chunk of item i to S(i+1), and
// the scheduler has four warps to pick from.  Registers are re-balanced with setmaxnreg (softmax 152, epilogue 88,
// TMA/MMA 40).  A fraction of the exponentials runs on the FMA pipe (ex2_poly_x2): the MUFU pipe is the busiest unit.
template <int KP, int NPH>
__global__ void __launch_bounds__(ATT2_THREADS, 2)
attention_tc2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                     const __grid_constant__ CUtensorMap tmO, int L, int heads, int tiles, int num_items,
                     float scale_log2) {
  using S = AttSmem<KP>;
  using S2 = Att2Smem<KP>;
  static_assert(KP % 32 == 16 && KP >= 48 && KP <= 208, "padded key count: whole x32 chunks plus one 16-key tail");
  static_assert(!ATT_DIRECT_STORE, "the epilogue warpgroup stores through the staging tile");
  constexpr int KMAIN = KP - 16;
  constexpr int O_COL = KMAIN;
  constexpr int KHALF = NPH == 2 ? 64 : 96;
  constexpr uint32_t TMEM_COLS = O_COL + HD <= 128 ? 128 : 256;
  static_assert(KP / 2 <= O_COL && O_COL + HD <= 256, "P / O column ranges must not overlap");
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* kq_full = reinterpret_cast<uint64_t*>(smem + S::OFF_BAR);
  uint64_t* v_full = kq_full + 1;
  uint64_t* s_full = kq_full + 2;
  uint64_t* p_full = kq_full + 3;
  uint64_t* o_full = kq_full + 4;
  uint64_t* o_empty = kq_full + 5;
  uint64_t* t_full = kq_full + 6;
  uint64_t* p_half = kq_full + 7;   // [2]
  uint64_t* l_full = kq_full + 10;  // row sums of item it are in l_buf[it & 1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(kq_full + 9);
  float* l_buf = reinterpret_cast<float*>(smem + S2::OFF_L);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int D = heads * HD;
  const bool flip_ok = (gridDim.x & 1) == 0 && tiles == 2;

  griddep_launch_dependents();
  if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0) __trap();
  if (threadIdx.x == 256) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmKV);
    tma_prefetch_desc(&tmO);
    mbar_init(kq_full, 1);
    mbar_init(v_full, 1);
    mbar_init(s_full, 1);
    mbar_init(p_full, 128);
    mbar_init(o_full, 1);
    mbar_init(o_empty, 128);
    mbar_init(t_full, 1);
    mbar_init(p_half, 128);
    mbar_init(p_half + 1, 128);
    mbar_init(l_full, 128);
    fence_barrier_init();
  }
  if (warp == 8) tmem_alloc<TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  griddep_wait();

  if (warp == 8) {
    // ===================== TMA + MMA thread (identical protocol to attention_tc_kernel) =====================
    // (keeps its launch-time registers: setmaxnreg is a warpgroup-wide instruction and this warp is alone in its group)
    if (lane == 0) {
      constexpr uint32_t idesc_s = umma_idesc_bf16_f32(QT, KMAIN);
      constexpr uint32_t idesc_t = umma_idesc_bf16_f32(QT, 16);
      constexpr uint32_t idesc_pv = umma_idesc_bf16_f32_bmn(QT, HD);
      const uint32_t q_addr = smem_u32(smem + S::OFF_Q);
      const uint32_t k_addr = smem_u32(smem + S::OFF_K);
      const uint32_t v_addr = smem_u32(smem + S::OFF_V);
      auto load_qk = [&](const ItemCursor& c, int n) {
        const int tile = flip_ok ? (c.t ^ (n & 1)) : c.t;
        mbar_expect_tx(kq_full, Q_BYTES + S::KV_BYTES);
        tma_load_3d(smem + S::OFF_Q, &tmQ, kq_full, c.head * HD, tile * QT, c.seq);
        tma_load_3d(smem + S::OFF_K, &tmKV, kq_full, D + c.head * HD, 0, c.seq);
      };
      auto load_v = [&](const ItemCursor& c) {
        mbar_expect_tx(v_full, S::KV_BYTES);
        tma_load_3d(smem + S::OFF_V, &tmKV, v_full, 2 * D + c.head * HD, 0, c.seq);
      };
      auto issue_s_main = [&]() {
#pragma unroll
        for (int k = 0; k < HD / 16; ++k)
          umma_bf16_ss(tmem_base, umma_desc_k_sw128(q_addr + k * 32), umma_desc_k_sw128(k_addr + k * 32), idesc_s,
                       k != 0);
        umma_commit(s_full);
      };
      auto issue_s_tail = [&]() {
#pragma unroll
        for (int k = 0; k < HD / 16; ++k)
          umma_bf16_ss(tmem_base + KMAIN, umma_desc_k_sw128(q_addr + k * 32),
                       umma_desc_k_sw128(k_addr + KMAIN * 128 + k * 32), idesc_t, k != 0);
        umma_commit(t_full);
      };
      ItemCursor cv, cqk;
      cv.init(blockIdx.x, gridDim.x, tiles, heads);
      cqk = cv;
      if (static_cast<int>(blockIdx.x) < num_items) {
        load_qk(cqk, 0);
        load_v(cv);
        cqk.advance();
        cv.advance();
        mbar_wait(kq_full, 0);
        tc_fence_after();
        issue_s_main();
        issue_s_tail();
        mbar_wait(t_full, 0);
        if (static_cast<int>(blockIdx.x + gridDim.x) < num_items) load_qk(cqk, 1);
        cqk.advance();
      }
      int it = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++it) {
        const uint32_t ph = it & 1;
        const int next = item + gridDim.x, next2 = next + gridDim.x;
        const bool has_next = next < num_items;
        mbar_wait(v_full, ph);
#pragma unroll
        for (int part = 0; part < NPH; ++part) {
          mbar_wait(p_half + part, ph);
          tc_fence_after();
#pragma unroll
          for (int k = part * (KHALF / 16); k < (part + 1) * (KHALF / 16); ++k)
            umma_bf16_ts(tmem_base + O_COL, tmem_base + k * 8, umma_desc_mn_sw128(v_addr + k * 2048), idesc_pv, k != 0);
        }
        mbar_wait(p_full, ph);
        tc_fence_after();
#pragma unroll
        for (int k = NPH * (KHALF / 16); k < KP / 16; ++k)
          umma_bf16_ts(tmem_base + O_COL, tmem_base + k * 8, umma_desc_mn_sw128(v_addr + k * 2048), idesc_pv, k != 0);
        umma_commit(o_full);
        if (has_next) {
          mbar_wait(kq_full, ph ^ 1);
          tc_fence_after();
          issue_s_main();
        }
        mbar_wait(o_full, ph);  // V tile is free again
        if (has_next) {
          load_v(cv);
          cv.advance();
          mbar_wait(o_empty, ph);  // the epilogue warps have read O(it): the tail columns may be overwritten
          tc_fence_after();
          issue_s_tail();
          mbar_wait(t_full, ph ^ 1);
          if (next2 < num_items) load_qk(cqk, it + 2);
          cqk.advance();
        }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue warps: O / l -> bf16 -> staging tile -> TMA store =====================
    setmaxnreg_dec<ATT2_REGS_EPILOGUE>();
    const int q = warp - 4;  // TMEM lane quadrant = warp % 4
    const uint32_t trow = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    uint8_t* stg_ptr = smem + S::OFF_STG + q * (32 * 128);
    const uint32_t stg_row = smem_u32(stg_ptr) + lane * 128;
    const int sw = lane & 7;
    ItemCursor cur;
    cur.init(blockIdx.x, gridDim.x, tiles, heads);
    int it = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++it, cur.advance()) {
      const uint32_t ph = it & 1;
      const int tile = flip_ok ? (cur.t ^ (it & 1)) : cur.t;
      const int row0 = tile * QT + q * 32;
      const bool active = row0 < L;
      mbar_wait(l_full, ph);  // arrived before p_full, hence before P.V was even issued: returns at once; orders the l read
      mbar_wait(o_full, ph);
      tc_fence_after();
      if (active) {
        const float inv = 1.f / l_buf[(it & 1) * 128 + q * 32 + lane];
        const uint64_t inv2 = pack_f32x2(inv, inv);
        // two halves of 32 columns, each scaled and packed to bf16 as soon as it arrives (32 + 16 live registers)
        uint32_t pk[2][16];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint32_t o[32];
          tmem_ld_32x32b_x32(trow + O_COL + 32 * h, o);
          tmem_ld_wait_fence(o);
          if (h == 1) {
            tc_fence_before();
            mbar_arrive(o_empty);  // O(it) is in registers: the MMA thread may overwrite the tail columns
          }
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            float a, b;
            unpack_f32x2(mul_f32x2(pack_f32x2(__uint_as_float(o[2 * e]), __uint_as_float(o[2 * e + 1])), inv2), a, b);
            pk[h][e] = pack_bf16x2(a, b);
          }
        }
        if (lane == 0) bulk_wait_group_read<0>();  // this warp's previous store has finished reading its staging tile
        __syncwarp();
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const uint32_t(&pp)[16] = pk[c >> 2];
          const int o = (c & 3) * 4;
          st_shared_v4(stg_row + ((c ^ sw) << 4), make_uint4(pp[o], pp[o + 1], pp[o + 2], pp[o + 3]));
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_3d(&tmO, stg_ptr, cur.head * HD, row0, cur.seq);  // rows >= L are clipped by the tensor map
          bulk_commit_group();
        }
      } else {
        tc_fence_before();
        mbar_arrive(o_empty);
      }
    }
    if (lane == 0) bulk_wait_group<0>();
  } else {
    // ===================== softmax warps (one query row per thread): S -> P, row sums =====================
    setmaxnreg_inc<ATT2_REGS_SOFTMAX>();
    const uint32_t trow = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
    constexpr int NFULL = KP / 32;
    static_assert(NPH == 0 || (KHALF % 32 == 0 && NPH * (KHALF / 32) < NFULL), "early hand-overs must end on chunk boundaries");
    ItemCursor cur;
    cur.init(blockIdx.x, gridDim.x, tiles, heads);
    int it = 0;
    FC_T(long long tq[7] = {0, 0, 0, 0, 0, 0, 0}; long long n_it = 0; const long long t_begin = clock64();)
    for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++it, cur.advance()) {
      const uint32_t ph = it & 1;
      const int tile = flip_ok ? (cur.t ^ (it & 1)) : cur.t;
      const int row0 = tile * QT + warp * 32;
      const bool active = row0 < L;
      const int lim = L;  // un-masked: every row attends to keys [0, L)
      FC_T(long long t0 = clock64(); long long t1;)
      mbar_wait(s_full, ph);
      FC_T(t1 = clock64(); tq[1] += t1 - t0; t0 = t1; ++n_it;)
      tc_fence_after();
      float l = 1.f;
      float m = -INFINITY;
      if (active) {
        float m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
        constexpr int DEPTH = 2;
        uint32_t r[DEPTH][32];
#pragma unroll
        for (int j = 0; j < DEPTH && j < NFULL; ++j) tmem_ld_32x32b_x32(trow + j * 32, r[j]);
#pragma unroll
        for (int j = 0; j < NFULL; ++j) {
          tmem_ld_wait_fence(r[j % DEPTH]);
          const uint32_t(&rc)[32] = r[j % DEPTH];
          if ((j + 1) * 32 <= lim) {
#pragma unroll
            for (int c = 0; c < 32; c += 8) {
              m = max3(m, __uint_as_float(rc[c]), __uint_as_float(rc[c + 1]));
              m1 = max3(m1, __uint_as_float(rc[c + 2]), __uint_as_float(rc[c + 3]));
              m2 = max3(m2, __uint_as_float(rc[c + 4]), __uint_as_float(rc[c + 5]));
              m3 = max3(m3, __uint_as_float(rc[c + 6]), __uint_as_float(rc[c + 7]));
            }
          } else {
#pragma unroll
            for (int c = 0; c < 32; ++c)
              if (j * 32 + c < lim) m = fmaxf(m, __uint_as_float(rc[c]));
          }
          if (j + DEPTH < NFULL) tmem_ld_32x32b_x32(trow + (j + DEPTH) * 32, r[j % DEPTH]);
        }
        m = fmaxf(max3(m, m1, m2), m3);
      }
      FC_T(t1 = clock64(); tq[2] += t1 - t0; t0 = t1;)
      mbar_wait(t_full, ph);
      FC_T(t1 = clock64(); tq[3] += t1 - t0; t0 = t1;)
      tc_fence_after();
      if (active) {
        uint32_t r16[16];
        tmem_ld_32x32b_x16(trow + NFULL * 32, r16);
        uint32_t r[2][32];
        tmem_ld_32x32b_x32(trow, r[0]);
        tmem_ld_wait_fence16(r16);
#pragma unroll
        for (int c = 0; c < 16; ++c)
          if (NFULL * 32 + c < lim) m = fmaxf(m, __uint_as_float(r16[c]));
        const float mc = m * scale_log2;
        const uint64_t sc2 = pack_f32x2(scale_log2, scale_log2), nmc2 = pack_f32x2(-mc, -mc);
        uint64_t l2a = pack_f32x2(0.f, 0.f), l2b = l2a;
        uint32_t pk_tail[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          float p0 = 0.f, p1 = 0.f;
          if (NFULL * 32 + 2 * c < lim) {  // compile-time per c only when lim is; L is a runtime value: a uniform branch
            float x0, x1;
            unpack_f32x2(fma_f32x2(pack_f32x2(__uint_as_float(r16[2 * c]), __uint_as_float(r16[2 * c + 1])), sc2, nmc2),
                         x0, x1);
            p0 = ex2_approx(x0);
            p1 = NFULL * 32 + 2 * c + 1 < lim ? ex2_approx(x1) : 0.f;
          }
          l2b = add_f32x2(l2b, pack_f32x2(p0, p1));
          pk_tail[c] = pack_bf16x2(p0, p1);
        }
#pragma unroll
        for (int j = 0; j < NFULL; ++j) {
          tmem_ld_wait_fence(r[j & 1]);
          if (j + 1 < NFULL) tmem_ld_32x32b_x32(trow + (j + 1) * 32, r[(j + 1) & 1]);
          uint32_t pk[16];
          const bool full = (j + 1) * 32 <= lim;
#pragma unroll
          for (int c = 0; c < 16; ++c) {
            const uint64_t x2 = fma_f32x2(pack_f32x2(__uint_as_float(r[j & 1][2 * c]), __uint_as_float(r[j & 1][2 * c + 1])),
                                          sc2, nmc2);
            float p0, p1;
            // ATT2_POLY of every 16 pairs take the FMA-pipe exp2, spread evenly so both pipes stay fed
            if (ATT2_POLY > 0 && (c * ATT2_POLY) % 16 < ATT2_POLY) {
              ex2_poly_x2(x2, p0, p1);
            } else {
              float x0, x1;
              unpack_f32x2(x2, x0, x1);
              p0 = ex2_approx(x0);
              p1 = ex2_approx(x1);
            }
            if (!full) {
              if (j * 32 + 2 * c >= lim) p0 = 0.f;
              if (j * 32 + 2 * c + 1 >= lim) p1 = 0.f;
            }
            if (c & 1) l2b = add_f32x2(l2b, pack_f32x2(p0, p1));
            else l2a = add_f32x2(l2a, pack_f32x2(p0, p1));
            pk[c] = pack_bf16x2(p0, p1);
          }
          tmem_st_32x32b_x16(trow + j * 16, pk);
          if (NPH > 0 && (j + 1) % (KHALF / 32) == 0 && (j + 1) / (KHALF / 32) <= NPH) {
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(p_half + (j + 1) / (KHALF / 32) - 1);
          }
        }
        tmem_st_32x32b_x8(trow + NFULL * 16, pk_tail);
        tmem_st_wait();
        float la, lb;
        unpack_f32x2(add_f32x2(l2a, l2b), la, lb);
        l = la + lb;
      } else {
#pragma unroll
        for (int part = 0; part < NPH; ++part) mbar_arrive(p_half + part);
      }
      l_buf[(it & 1) * 128 + threadIdx.x] = l;
      mbar_arrive(l_full);  // release: the epilogue thread of this row reads l after its acquire-wait on l_full
      tc_fence_before();
      mbar_arrive(p_full);
      FC_T(t1 = clock64(); tq[4] += t1 - t0; t0 = t1;)
    }
    FC_T(if (threadIdx.x == 0) {
      atomicAdd(&g_att_timing[0], static_cast<unsigned long long>(clock64() - t_begin));
      for (int i = 1; i < 7; ++i) atomicAdd(&g_att_timing[i], static_cast<unsigned long long>(tq[i]));
      atomicAdd(&g_att_timing[7], static_cast<unsigned long long>(n_it));
    })
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
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

template <int KP, bool CAUSAL, int NPH>
int launch_tc(const bf16* qkv, bf16* out, int64_t seqs, int L, int heads, cudaStream_t s) {
  using S = AttSmem<KP>;
  static bool configured = false;
  if (!configured) {
    FC_CUDA(cudaFuncSetAttribute(attention_tc_kernel<KP, CAUSAL, NPH>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::BYTES));
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
  FC_CUDA(cudaLaunchKernelEx(&cfg, attention_tc_kernel<KP, CAUSAL, NPH>, tq, tkv, to, out, L, heads, tiles, items, scale_log2));
  return FC_OK;
}

template <int KP, int NPH>
int launch_tc2(const bf16* qkv, bf16* out, int64_t seqs, int L, int heads, cudaStream_t s) {
  using S2 = Att2Smem<KP>;
  static bool configured = false;
  if (!configured) {
    cudaFuncAttributes fa;
    FC_CUDA(cudaFuncGetAttributes(&fa, attention_tc2_kernel<KP, NPH>));
    FC_REQUIRE(128 * ATT2_REGS_SOFTMAX + 128 * ATT2_REGS_EPILOGUE + 32 * fa.numRegs <= ATT2_THREADS * fa.numRegs &&
                   ATT2_REGS_EPILOGUE <= fa.numRegs && ATT2_REGS_SOFTMAX >= fa.numRegs,
               "attention_tc2: setmaxnreg plan (%d/%d) does not fit the %d registers per thread the kernel launches with",
               ATT2_REGS_SOFTMAX, ATT2_REGS_EPILOGUE, fa.numRegs);
    FC_CUDA(cudaFuncSetAttribute(attention_tc2_kernel<KP, NPH>, cudaFuncAttributeMaxDynamicSharedMemorySize, S2::BYTES));
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
  cfg.blockDim = dim3(ATT2_THREADS);
  cfg.dynamicSmemBytes = S2::BYTES;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  note_launch();
  FC_CUDA(cudaLaunchKernelEx(&cfg, attention_tc2_kernel<KP, NPH>, tq, tkv, to, L, heads, tiles, items, scale_log2));
  return FC_OK;
}

}  // namespace

#ifdef FC_GEMM_TIMING
extern "C" __attribute__((visibility("default"))) int fc_debug_att_timing(unsigned long long* out, int reset) {
  if (cudaDeviceSynchronize() != cudaSuccess) return -1;
  if (out && cudaMemcpyFromSymbol(out, g_att_timing, sizeof(g_att_timing)) != cudaSuccess) return -1;
  if (reset) {
    unsigned long long z[8] = {};
    if (cudaMemcpyToSymbol(g_att_timing, z, sizeof(z)) != cudaSuccess) return -1;
  }
  return 0;
}
#endif

// tcgen05 path for every sequence of up to 208 tokens, masked or not (ViT-B/16 image 197, ViT-B/32 image 50, CLIP text 77
// causal, ...). *handled = 1 when it took the call.
int attention_bf16_tc(const bf16* qkv, bf16* out, int64_t seqs, int L, int heads, int causal, cudaStream_t s,
                      int* handled) {
  *handled = 0;
  // diagnostics: FC_ATTENTION=mma forces the mma.sync kernels, =tc1 the first-generation tcgen05 kernel for the image
  // sequence, =long the key-block kernel (attention_tc_long.cu) for every un-masked length
  static int mode = -1;  // 0 default, 1 tc1, 2 mma, 3 long
  if (mode < 0) {
    const char* e = getenv("FC_ATTENTION");
    mode = !e ? 0 : strcmp(e, "tc1") == 0 ? 1 : strcmp(e, "mma") == 0 ? 2 : strcmp(e, "long") == 0 ? 3 : 0;
  }
  const bool gen1 = mode == 1;
  if (mode == 2 || (mode == 3 && !causal)) return FC_OK;
  if (L > 208) return FC_OK;  // longer un-masked sequences: attention_tc_long.cu; longer causal ones: mma.sync kernels
  FC_REQUIRE((reinterpret_cast<uintptr_t>(qkv) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
             "attention: buffers must be 16-byte aligned");
  *handled = 1;
  // padded key counts: 80 (CLIP text 77 causal; ViT-B/32 image 50), 144, 208 (ViT-B/16 image 197).  The kernels mask
  // keys >= L generically; attention_tc2_kernel relies on L > KP - 16 (un-masked main chunks).
  if (!causal) {
    if (L > 192 && !gen1) return launch_tc2<208, ATT_PHALF>(qkv, out, seqs, L, heads, s);
    if (L > 144) return launch_tc<208, false, ATT_PHALF>(qkv, out, seqs, L, heads, s);
    if (L > 80) return launch_tc<144, false, 0>(qkv, out, seqs, L, heads, s);
    return launch_tc<80, false, 0>(qkv, out, seqs, L, heads, s);
  }
  if (L > 144) return launch_tc<208, true, 0>(qkv, out, seqs, L, heads, s);
  if (L > 80) return launch_tc<144, true, 0>(qkv, out, seqs, L, heads, s);
  return launch_tc<80, true, 0>(qkv, out, seqs, L, heads, s);
}

}  // namespace fc
