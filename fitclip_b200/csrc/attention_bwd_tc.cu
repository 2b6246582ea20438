// tcgen05 attention BACKWARD for sequences of up to 208 tokens, head dim 64 (the 197-token image and the 77-token
// causal text sequence of CLIP ViT-B/16) -- training step, SURVEY.md 8f row f3.  Same math as attention_bwd_kernel
// (train.cu): with P = softmax(q k^T / 8 [+ causal mask]), delta_i = sum_d dO_id O_id:
//     dV = P^T dO,  dP = dO V^T,  dS = P o (dP - delta),  dQ = dS K / 8,  dK = dS^T Q / 8.
//
// One CTA per (sequence, head); Q, K, V, dO of the head are TMA-loaded once (3-D tensor maps: rows beyond the sequence
// are zero-filled) as K-major 128-byte-swizzled tiles of R = 128 or 256 rows.  All seven products run on the tensor
// core with accumulators in TMEM; one query (phase A) or key (phase B) row per softmax thread, so row statistics need
// no cross-thread reduction.
//   phase A, per 128-query tile:  S = Q_t K^T -> cols [0, NP);  dP = dO_t V^T -> cols [NP, 2 NP)      (NP = L rounded to 16)
//        softmax threads: row max, then P~ = exp2(S c - m c), the row sum l and dS~ = P~ o (dP - delta) as bf16 written
//        back over the dP columns; dQ = dS~ K (A = dS~ from TMEM, B = K read MN-major) -> cols [2 NP, 2 NP + 64)
//        -> x scale / l -> global   (short sequences, NP < 64: the 64-column accumulators start at NP + 64)
//        the row's log-sum-exp and delta go to shared memory for phase B
//   phase B, per 128-key tile:    S^T = K_t Q^T -> [0, NP);  dP^T = V_t dO^T -> [NP, 2 NP)
//        softmax threads: P^T = exp2(S^T c - lse_q) and dS^T = P^T o (dP^T - delta_q), bf16
//        dV = P^T dO (B = dO MN-major) -> [2 NP, +64);  dK = dS^T Q (B = Q MN-major) -> [NP, NP + 64) -> global
// bf16 results are written over the 16-column fp32 chunk they were computed from (chunk j -> columns 16 j .. 16 j + 15),
// so no thread overwrites input another thread still has to read, and each UMMA_K step takes its A operand from there.
// Inside a CTA a stage is MMA batch -> softmax -> MMA batch -> read-out; only the S / dP products of the stage behind a
// phase-A stage overlap that stage's read-out (four mbarriers: S/dP ready, operands ready, accumulators ready, read).
// Measured and not kept: splitting a phase-B stage into the two column parts of the softmax threads (part 1's products
// under part 0's softmax, part 0's dV / dK steps under part 1's softmax) -- 3.94 vs 3.86 ms per layer: the narrower
// MMAs (N = 112 / 96 instead of 208) cost what the overlap gains.
#include <stdlib.h>
#include <string.h>

#include "kernels.cuh"
#include "ptx.cuh"
#include "tmap.cuh"

namespace fc {

namespace {

constexpr int HD = 64;
constexpr int QT = 128;
// threads: NPARTS softmax threads per row (4 NPARTS warps), then one warp for TMA + MMA issue + TMEM allocation
constexpr int bwd_threads(int nparts) { return nparts * 128 + 32; }

// bf16 tensor viewed as [seqs][L][cols] (cols contiguous); box = 64 columns x box_rows tokens x 1 sequence, SW128.
int make_tmap_3d(CUtensorMap* tm, const bf16* base, int64_t cols, int64_t L, int64_t seqs, int box_rows) {
  const uint64_t dims[3] = {static_cast<uint64_t>(cols), static_cast<uint64_t>(L), static_cast<uint64_t>(seqs)};
  const uint64_t strides[2] = {static_cast<uint64_t>(cols) * 2, static_cast<uint64_t>(cols) * 2 * L};
  const uint32_t box[3] = {64, static_cast<uint32_t>(box_rows), 1};
  return tmap_bf16_sw128(tm, base, 3, dims, strides, box);
}

__device__ __forceinline__ void store_row32(bf16* dst, const uint32_t (&a)[16], const uint32_t (&b)[16], float k) {
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const uint32_t(&r)[16] = c < 2 ? a : b;
    const int o = (c & 1) * 8;
    uint4 u;
    u.x = pack_bf16x2(__uint_as_float(r[o + 0]) * k, __uint_as_float(r[o + 1]) * k);
    u.y = pack_bf16x2(__uint_as_float(r[o + 2]) * k, __uint_as_float(r[o + 3]) * k);
    u.z = pack_bf16x2(__uint_as_float(r[o + 4]) * k, __uint_as_float(r[o + 5]) * k);
    u.w = pack_bf16x2(__uint_as_float(r[o + 6]) * k, __uint_as_float(r[o + 7]) * k);
    reinterpret_cast<uint4*>(dst)[c] = u;
  }
}

// TMEM_COLS: 512 for the image sequence (S 208 + dP 208 + 64: one CTA per SM), 256 when 2 NP + 64 <= 256 (the text
// sequence: two co-resident CTAs hide each other's load and read-out).  BWD_NPARTS softmax threads per row: measured on
// 2048 x 197 x 12 heads, 1 -> 4.53 ms, 2 -> 3.86 ms, 3 -> 4.07 ms per layer (128-register cap, 5/4/4 chunk split, a third
// partner in the row-statistics exchange): a stage is a serial chain MMA -> softmax -> MMA -> read-out of about equal
// parts, so more softmax threads stop paying after two.
constexpr int BWD_NPARTS = 2;
template <bool CAUSAL, uint32_t TMEM_COLS>
__global__ void __launch_bounds__(bwd_threads(BWD_NPARTS), TMEM_COLS == 256 ? 2 : 1)
attention_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmDO,
                        const bf16* __restrict__ O, const bf16* __restrict__ dO, bf16* __restrict__ dqkv, int L, int NP,
                        int tiles, int heads, float scale, float scale_log2) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int R = tiles * QT;  // rows per resident tile (128 or 256)
  const int tile_bytes = R * 128;
  uint8_t* sQ = smem;
  uint8_t* sK = smem + tile_bytes;
  uint8_t* sV = smem + 2 * tile_bytes;
  uint8_t* sdO = smem + 3 * tile_bytes;
  float* sLse = reinterpret_cast<float*>(smem + 4 * tile_bytes);
  float* sDelta = sLse + 256;
  constexpr int NPARTS = BWD_NPARTS;  // softmax threads per row
  constexpr int SOFT_THREADS = NPARTS * 128;
  constexpr int T_WARP = NPARTS * 4;
  float* sRed = sDelta + 256;  // [2][NPARTS][128]: row max / row sum partials of the threads of a row
  uint64_t* bar_load = reinterpret_cast<uint64_t*>(sRed + 2 * 3 * 128);
  uint64_t* bar_sd = bar_load + 1;   // MMA -> softmax: S / dP (or S^T / dP^T) of a stage are in TMEM
  uint64_t* bar_sm = bar_load + 2;   // softmax -> MMA: the bf16 operands of the stage are in TMEM
  uint64_t* bar_acc = bar_load + 3;  // MMA -> read-out: the accumulators of the stage are complete
  uint64_t* bar_rd = bar_load + 4;   // read-out -> MMA: the accumulators have been read
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_load + 5);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int head = blockIdx.x;
  const int seq = blockIdx.y;
  const int D = heads * HD;
  const int nk16 = NP / 16;                                     // 16-wide column chunks / UMMA_K steps over NP
  const int DK_COL = NP;  // dK accumulator: over the dP^T columns, which are dead once the softmax threads are through
  const int ACC_COL = 2 * NP > NP + HD ? 2 * NP : NP + HD;  // dQ / dV accumulator: behind dP (and behind dK for short sequences)

  if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0) __trap();
  if (threadIdx.x == SOFT_THREADS) {
    tma_prefetch_desc(&tmQKV);
    tma_prefetch_desc(&tmDO);
    mbar_init(bar_load, 1);
    mbar_init(bar_sd, 1);
    mbar_init(bar_sm, SOFT_THREADS);
    mbar_init(bar_acc, 1);
    mbar_init(bar_rd, SOFT_THREADS);
    fence_barrier_init();
  }
  if (warp == T_WARP) tmem_alloc<TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == T_WARP) {
    // ===================== TMA + MMA thread =====================
    if (lane == 0) {
      mbar_expect_tx(bar_load, 4 * tile_bytes);
      tma_load_3d(sQ, &tmQKV, bar_load, head * HD, 0, seq);
      tma_load_3d(sK, &tmQKV, bar_load, D + head * HD, 0, seq);
      tma_load_3d(sV, &tmQKV, bar_load, 2 * D + head * HD, 0, seq);
      tma_load_3d(sdO, &tmDO, bar_load, head * HD, 0, seq);
      const uint32_t q_addr = smem_u32(sQ), k_addr = smem_u32(sK), v_addr = smem_u32(sV), do_addr = smem_u32(sdO);
      const uint32_t idesc_s = umma_idesc_bf16_f32(QT, NP);
      const uint32_t idesc_o = umma_idesc_bf16_f32_bmn(QT, HD);
      mbar_wait(bar_load, 0);
      tc_fence_after();
      // stages 0 .. tiles-1: phase A (query tiles), stages tiles .. 2 tiles-1: phase B (key tiles)
      const int stages = 2 * tiles;
      auto issue_sd = [&](int st) {
        const bool pa = st < tiles;
        const uint32_t row_off = (pa ? st : st - tiles) * QT * 128;
        const uint32_t a0 = (pa ? q_addr : k_addr) + row_off, b0 = pa ? k_addr : q_addr;
        const uint32_t a1 = (pa ? do_addr : v_addr) + row_off, b1 = pa ? v_addr : do_addr;
#pragma unroll
        for (int k = 0; k < HD / 16; ++k)
          umma_bf16_ss(tmem_base, umma_desc_k_sw128(a0 + k * 32), umma_desc_k_sw128(b0 + k * 32), idesc_s, k != 0);
#pragma unroll
        for (int k = 0; k < HD / 16; ++k)
          umma_bf16_ss(tmem_base + NP, umma_desc_k_sw128(a1 + k * 32), umma_desc_k_sw128(b1 + k * 32), idesc_s, k != 0);
        umma_commit(bar_sd);
      };
      uint32_t p_sm = 0, p_rd = 0;
      issue_sd(0);
      for (int st = 0; st < stages; ++st) {
        mbar_wait(bar_sm, p_sm);  // the stage's bf16 operands are in TMEM
        p_sm ^= 1;
        tc_fence_after();
        if (st < tiles) {  // dQ = dS~ K
          for (int k = 0; k < nk16; ++k)
            umma_bf16_ts(tmem_base + ACC_COL, tmem_base + NP + k * 16, umma_desc_mn_sw128(k_addr + k * 2048), idesc_o,
                         k != 0);
        } else {  // dV = P^T dO, dK = dS^T Q
          for (int k = 0; k < nk16; ++k)
            umma_bf16_ts(tmem_base + ACC_COL, tmem_base + k * 16, umma_desc_mn_sw128(do_addr + k * 2048), idesc_o,
                         k != 0);
          for (int k = 0; k < nk16; ++k)
            umma_bf16_ts(tmem_base + DK_COL, tmem_base + k * 16 + 8, umma_desc_mn_sw128(q_addr + k * 2048), idesc_o,
                         k != 0);
        }
        umma_commit(bar_acc);
        // The next stage's S / dP products are queued right behind a phase-A stage: the tensor pipe runs in order, so
        // they overwrite dS~ only after dQ consumed it, and they do not touch the dQ columns, so they overlap the
        // read-out.  Behind a phase-B stage they would land on the dK accumulator (it sits in the dP^T columns).
        const bool early = st < tiles && st + 1 < stages;
        if (early) issue_sd(st + 1);
        mbar_wait(bar_rd, p_rd);  // the accumulators have been read out
        p_rd ^= 1;
        tc_fence_after();
        if (!early && st + 1 < stages) issue_sd(st + 1);
      }
    }
  } else {
    // ===================== softmax / read-out warps: NPARTS threads per row =====================
    // warps 4 p .. 4 p + 3 take part p of a row's column chunks (TMEM lane quarter = warp % 4): with one thread per row
    // only four warps per SM work through the exponentials, one per scheduler, with nothing to hide their latencies
    // behind (4.53 ms per layer; two per row 3.99); the row maximum and sum are combined through shared memory.
    const int quarter = warp & 3, part = warp >> 2;
    const uint32_t trow = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    const int r = quarter * 32 + lane;
    const int64_t tok0 = static_cast<int64_t>(seq) * L;
    const int nh = (nk16 + NPARTS - 1) / NPARTS;         // chunks [p nh, (p + 1) nh) belong to part p
    const int own_lo = min(part * nh, nk16), own_hi = min(own_lo + nh, nk16);
    const uint32_t zero[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
    uint32_t pm = 0;
    uint32_t s0[16], s1[16], d0[16], d1[16];
    // ---------- phase A
    for (int t = 0; t < tiles; ++t) {
      const int row = t * QT + r;
      const bool valid = row < L;
      // delta_row = dO_row . O_row from global memory while the first MMAs run
      float delta = 0.f;
      if (valid) {
        const uint4* po = reinterpret_cast<const uint4*>(O + (tok0 + row) * D + head * HD);
        const uint4* pd = reinterpret_cast<const uint4*>(dO + (tok0 + row) * D + head * HD);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const uint4 a = po[c], b = pd[c];
          const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 x = unpack_bf16x2(aw[e]), y = unpack_bf16x2(bw[e]);
            delta = fmaf(x.x, y.x, fmaf(x.y, y.y, delta));
          }
        }
      }
      const int lim = !valid ? 0 : (CAUSAL ? min(L, row + 1) : L);  // keys [0, lim) are visible to this row
      // tcgen05.ld / .st are warp-collective: loop bounds are warp-uniform (wlim = the largest lim of the warp), the
      // per-row limit only masks the arithmetic
      const int wrow0 = t * QT + quarter * 32;
      const int wlim = wrow0 >= L ? 0 : (CAUSAL ? min(L, wrow0 + 32) : L);
      const int lo = own_lo, hi = min(own_hi, (wlim + 15) / 16);  // this thread's live chunks [lo, hi)
      mbar_wait(bar_sd, pm);
      tc_fence_after();
      float m = -INFINITY, l = 0.f;
      // TMEM loads are software-pipelined: chunk j + 1 is requested before chunk j is consumed (a load + wait per chunk
      // serialises on the TMEM latency: 5.03 ms per layer for the first version of this kernel)
      {  // pass 1: row maximum
        // chunks whose 16 keys are valid and visible to all rows of the warp skip the per-element limit (rows beyond
        // the sequence then hold garbage statistics, which only ever reach their own, never stored, dQ row)
        auto body = [&](int j, const uint32_t(&b)[16]) {
          if (j * 16 + 16 <= L && (!CAUSAL || j * 16 + 15 <= wrow0)) {
            float m1 = -INFINITY;
#pragma unroll
            for (int c = 0; c < 16; c += 2) {
              m = fmaxf(m, __uint_as_float(b[c]));
              m1 = fmaxf(m1, __uint_as_float(b[c + 1]));
            }
            m = fmaxf(m, m1);
          } else {
#pragma unroll
            for (int c = 0; c < 16; ++c)
              if (j * 16 + c < lim) m = fmaxf(m, __uint_as_float(b[c]));
          }
        };
        if (lo < hi) tmem_ld_32x32b_x16(trow + lo * 16, s0);
        for (int j = lo; j < hi; j += 2) {
          tmem_ld_wait_fence16(s0);
          if (j + 1 < hi) tmem_ld_32x32b_x16(trow + (j + 1) * 16, s1);
          body(j, s0);
          if (j + 1 < hi) {
            tmem_ld_wait_fence16(s1);
            if (j + 2 < hi) tmem_ld_32x32b_x16(trow + (j + 2) * 16, s0);
            body(j + 1, s1);
          }
        }
      }
      sRed[part * 128 + r] = m;
      named_bar_sync(2, SOFT_THREADS);
#pragma unroll
      for (int p = 0; p < NPARTS; ++p) m = fmaxf(m, sRed[p * 128 + r]);
      const float mc = valid ? m * scale_log2 : 0.f;
      {  // pass 2: P~ = exp2(S c - m c) (un-normalised), row sum, dS~ = P~ o (dP - delta) -> bf16 over the dP columns;
         // dQ is linear in P~, so the 1 / l goes into the read-out of dQ
        auto body = [&](int j, const uint32_t(&bs)[16], const uint32_t(&bd)[16]) {
          uint32_t pk[8];
          const bool full = j * 16 + 16 <= L && (!CAUSAL || j * 16 + 15 <= wrow0);  // warp-uniform
          float l1 = 0.f;
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            float v[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int col = j * 16 + 2 * c + e;
              float p = ex2_approx(fmaf(__uint_as_float(bs[2 * c + e]), scale_log2, -mc));
              if (!full && col >= lim) p = 0.f;
              if (e) l1 += p; else l += p;
              v[e] = p * (__uint_as_float(bd[2 * c + e]) - delta);
            }
            pk[c] = pack_bf16x2(v[0], v[1]);
          }
          l += l1;
          tmem_st_32x32b_x8(trow + NP + j * 16, pk);  // over the first half of the chunk's own (consumed) dP columns
        };
        if (lo < hi) {
          tmem_ld_32x32b_x16(trow + lo * 16, s0);
          tmem_ld_32x32b_x16(trow + NP + lo * 16, d0);
        }
        for (int j = lo; j < hi; j += 2) {
          tmem_ld_wait_fence16(s0);
          tmem_ld_wait_fence16(d0);
          if (j + 1 < hi) {
            tmem_ld_32x32b_x16(trow + (j + 1) * 16, s1);
            tmem_ld_32x32b_x16(trow + NP + (j + 1) * 16, d1);
          }
          body(j, s0, d0);
          if (j + 1 < hi) {
            tmem_ld_wait_fence16(s1);
            tmem_ld_wait_fence16(d1);
            if (j + 2 < hi) {
              tmem_ld_32x32b_x16(trow + (j + 2) * 16, s0);
              tmem_ld_32x32b_x16(trow + NP + (j + 2) * 16, d0);
            }
            body(j + 1, s1, d1);
          }
        }
        for (int j = max(hi, own_lo); j < own_hi; ++j) tmem_st_32x32b_x8(trow + NP + j * 16, zero);  // keys beyond the limit
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(bar_sm);
      sRed[384 + part * 128 + r] = l;
      named_bar_sync(2, SOFT_THREADS);
      l = 0.f;
#pragma unroll
      for (int p = 0; p < NPARTS; ++p) l += sRed[384 + p * 128 + r];
      if (part == 0) {
        sLse[row] = valid ? fmaf(m, scale_log2, log2f(l)) : 0.f;
        sDelta[row] = delta;
      }
      // dQ read-out: 32 of the 64 columns per thread (parts 0 and 1)
      mbar_wait(bar_acc, pm);
      pm ^= 1;
      tc_fence_after();
      if (wrow0 < L && part < 2) {  // warp-uniform
        tmem_ld_32x32b_x16(trow + ACC_COL + part * 32, s0);
        tmem_ld_32x32b_x16(trow + ACC_COL + part * 32 + 16, s1);
        tmem_ld_wait_fence16(s0);
        tmem_ld_wait_fence16(s1);
        if (valid) store_row32(dqkv + (tok0 + row) * (3 * D) + head * HD + part * 32, s0, s1, scale / l);
      }
      tc_fence_before();
      mbar_arrive(bar_rd);
    }
    named_bar_sync(1, SOFT_THREADS);  // every row's log-sum-exp / delta is in shared memory
    // ---------- phase B
    for (int t = 0; t < tiles; ++t) {
      const int key = t * QT + r;
      const bool kvalid = key < L;
      const int wkey0 = t * QT + quarter * 32;  // smallest key of the warp: the warp-uniform causal bound
      mbar_wait(bar_sd, pm);
      tc_fence_after();
      {
        // warp-uniform live range of this thread's query chunks (causal: queries before the warp's smallest key are masked)
        const int jhi = wkey0 < L ? min(own_hi, (L + 15) / 16) : own_lo;
        const int jlo = min(jhi, max(own_lo, CAUSAL ? wkey0 / 16 : 0));
        auto body = [&](int j, const uint32_t(&bs)[16], const uint32_t(&bd)[16]) {
          uint32_t pp[8], pd[8];
          // warp-uniform: all 16 queries valid and visible to all keys of the warp -> no per-element mask (key rows
          // beyond the sequence then hold garbage, which only reaches their own, never stored, dK / dV rows)
          const bool full = j * 16 + 16 <= L && (!CAUSAL || j * 16 >= wkey0 + 31);
          const float4* lse4 = reinterpret_cast<const float4*>(sLse + j * 16);
          const float4* dl4 = reinterpret_cast<const float4*>(sDelta + j * 16);
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const float4 ls = lse4[g], dl = dl4[g];
            const float lsv[4] = {ls.x, ls.y, ls.z, ls.w}, dlv[4] = {dl.x, dl.y, dl.z, dl.w};
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              float pv[2], dv[2];
#pragma unroll
              for (int e = 0; e < 2; ++e) {
                const int i = 4 * g + 2 * h + e, q = j * 16 + i;
                float p = ex2_approx(fmaf(__uint_as_float(bs[i]), scale_log2, -lsv[2 * h + e]));
                if (!full && (!kvalid || q >= L || (CAUSAL && key > q))) p = 0.f;
                pv[e] = p;
                dv[e] = p * (__uint_as_float(bd[i]) - dlv[2 * h + e]);
              }
              pp[2 * g + h] = pack_bf16x2(pv[0], pv[1]);
              pd[2 * g + h] = pack_bf16x2(dv[0], dv[1]);
            }
          }
          tmem_st_32x32b_x8(trow + j * 16, pp);      // both results go over the chunk's own (consumed) S^T columns:
          tmem_st_32x32b_x8(trow + j * 16 + 8, pd);  // no thread ever overwrites another thread's unread input
        };
        if (jlo < jhi) {
          tmem_ld_32x32b_x16(trow + jlo * 16, s0);
          tmem_ld_32x32b_x16(trow + NP + jlo * 16, d0);
        }
        for (int j = jlo; j < jhi; j += 2) {
          tmem_ld_wait_fence16(s0);
          tmem_ld_wait_fence16(d0);
          if (j + 1 < jhi) {
            tmem_ld_32x32b_x16(trow + (j + 1) * 16, s1);
            tmem_ld_32x32b_x16(trow + NP + (j + 1) * 16, d1);
          }
          body(j, s0, d0);
          if (j + 1 < jhi) {
            tmem_ld_wait_fence16(s1);
            tmem_ld_wait_fence16(d1);
            if (j + 2 < jhi) {
              tmem_ld_32x32b_x16(trow + (j + 2) * 16, s0);
              tmem_ld_32x32b_x16(trow + NP + (j + 2) * 16, d0);
            }
            body(j + 1, s1, d1);
          }
        }
        for (int j = own_lo; j < jlo; ++j) {
          tmem_st_32x32b_x8(trow + j * 16, zero);
          tmem_st_32x32b_x8(trow + j * 16 + 8, zero);
        }
        for (int j = jhi; j < own_hi; ++j) {
          tmem_st_32x32b_x8(trow + j * 16, zero);
          tmem_st_32x32b_x8(trow + j * 16 + 8, zero);
        }
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(bar_sm);
      mbar_wait(bar_acc, pm);
      pm ^= 1;
      tc_fence_after();
      if (wkey0 < L && part < 2) {  // warp-uniform
        bf16* dst = dqkv + (tok0 + key) * (3 * D) + D + head * HD + part * 32;
        tmem_ld_32x32b_x16(trow + DK_COL + part * 32, s0);
        tmem_ld_32x32b_x16(trow + DK_COL + part * 32 + 16, s1);
        tmem_ld_32x32b_x16(trow + ACC_COL + part * 32, d0);
        tmem_ld_32x32b_x16(trow + ACC_COL + part * 32 + 16, d1);
        tmem_ld_wait_fence16(s0);
        tmem_ld_wait_fence16(s1);
        tmem_ld_wait_fence16(d0);
        tmem_ld_wait_fence16(d1);
        if (kvalid) {
          store_row32(dst, s0, s1, scale);
          store_row32(dst + D, d0, d1, 1.f);
        }
      }
      tc_fence_before();
      mbar_arrive(bar_rd);
    }
  }


  tc_fence_before();
  __syncthreads();
  if (warp == T_WARP) {
    tc_fence_after();
    tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

template <bool CAUSAL, uint32_t TMEM_COLS>
int launch_bwd_tc(const bf16* qkv, const bf16* O, const bf16* dO, bf16* dqkv, int64_t seqs, int L, int heads,
                  cudaStream_t s) {
  const int tiles = (L + QT - 1) / QT;
  const int R = tiles * QT;
  const int smem = 4 * R * 128 + 2 * 256 * 4 + 2 * 3 * 128 * 4 + 64;
  static int configured = 0;
  if (configured < smem) {
    FC_CUDA(cudaFuncSetAttribute(attention_bwd_tc_kernel<CAUSAL, TMEM_COLS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = smem;
  }
  const int D = heads * HD;
  const int NP = (L + 15) / 16 * 16;
  CUtensorMap tqkv, tdo;
  int rc;
  if ((rc = make_tmap_3d(&tqkv, qkv, 3 * D, L, seqs, R))) return rc;
  if ((rc = make_tmap_3d(&tdo, dO, D, L, seqs, R))) return rc;
  const float scale = 0.125f, scale_log2 = 0.125f * 1.4426950408889634f;
  for (int64_t s0 = 0; s0 < seqs; s0 += 65535) {
    FC_REQUIRE(s0 == 0, "attention backward (tcgen05): more than 65535 sequences per call are not supported");
    dim3 grid(heads, static_cast<unsigned>(seqs));
    attention_bwd_tc_kernel<CAUSAL, TMEM_COLS><<<grid, bwd_threads(BWD_NPARTS), smem, s>>>(tqkv, tdo, O, dO, dqkv, L, NP, tiles, heads, scale,
                                                                    scale_log2);
    FC_CHECK_LAUNCH();
  }
  return FC_OK;
}

}  // namespace

// *handled = 1 when the tcgen05 kernel took the call (L <= 208, 16-byte aligned buffers, <= 65535 sequences).
// FC_ATTENTION_BWD=mma forces the mma.sync kernel of train.cu (diagnostics / A-B timing).
int attention_bwd_bf16_tc(const bf16* qkv, const bf16* O, const bf16* dO, bf16* dqkv, int64_t seqs, int L, int heads,
                          int causal, cudaStream_t s, int* handled) {
  *handled = 0;
  static int mode = -1;
  if (mode < 0) {
    const char* e = getenv("FC_ATTENTION_BWD");
    mode = (e && strcmp(e, "mma") == 0) ? 0 : 1;
  }
  if (!mode || L > 208 || L < 1 || seqs > 65535) return FC_OK;
  if ((reinterpret_cast<uintptr_t>(qkv) & 15) || (reinterpret_cast<uintptr_t>(O) & 15) ||
      (reinterpret_cast<uintptr_t>(dO) & 15) || (reinterpret_cast<uintptr_t>(dqkv) & 15))
    return FC_OK;
  *handled = 1;
  const int NP = (L + 15) / 16 * 16;
  const bool small = (2 * NP > NP + HD ? 2 * NP : NP + HD) + HD <= 256;  // last accumulator column fits 256 columns
  if (causal)
    return small ? launch_bwd_tc<true, 256>(qkv, O, dO, dqkv, seqs, L, heads, s)
                 : launch_bwd_tc<true, 512>(qkv, O, dO, dqkv, seqs, L, heads, s);
  return small ? launch_bwd_tc<false, 256>(qkv, O, dO, dqkv, seqs, L, heads, s)
               : launch_bwd_tc<false, 512>(qkv, O, dO, dqkv, seqs, L, heads, s);
}

}  // namespace fc
