// Persistent, warp-specialised tcgen05 GEMM for sm_100a:  C[M,N] = epilogue(A[M,K] * B[N,K]^T), bf16 in, fp32 acc.
//
//   warp 0    TMA producer   : cp.async.bulk.tensor (128B-swizzled 128x64 A tile + 128x64 half B tile per stage)
//   warp 1    MMA issuer     : (leader CTA) one thread issues tcgen05.mma.cta_group::2 256x256x16, accumulators in TMEM
//   warp 2    TMEM allocator : 512 columns = 2 accumulator stages of 256 fp32 columns
//   warp 3    idle
//   warps 4-11 epilogue      : two groups of 4 warps (one warp per TMEM lane quarter)
//
// CTA PAIRS (thread-block cluster of 2, tcgen05 cta_group::2): a pair owns a 256x256 output tile.  Each CTA TMA-loads
// its own 128 rows of A and HALF of the B tile (128 of 256 rows) into its own shared memory -- 32 KiB per K block
// instead of 48 -- and the pair's LEADER issues one tcgen05.mma.cta_group::2 (M = 256) per K step that reads both
// CTAs' operands and accumulates each CTA's 128 rows into that CTA's TMEM.  The main loop is bound by the operand bytes
// delivered into an SM (measured: DESIGN.md 3.1), so the bytes per unit of MMA work are what counts.
// Three pipelines: smem full/empty (TMA<->MMA, 5 stages of 32 KiB; both CTAs' loads complete on the leader's full
// barrier, the leader's multicast tcgen05.commit releases the stage in both CTAs), TMEM full/empty (MMA<->epilogue,
// 2 accumulator stages; both epilogues release onto the leader's barrier with a relaxed cluster-scope arrive), and a
// static persistent tile scheduler over pairs with N fastest, so the pairs of one wave share A tiles in L2 while B
// stays L2-resident.
//
// Epilogues
//   staged (bias / bias+QuickGELU / bias+residual / folded-LayerNorm variants, bf16 out): each warp group owns the
//     64-column sub-tiles {g, g+2} of the 128x256 tile and two 16 KiB staging tiles.  tcgen05.ld -> registers ->
//     packed fp32x2 arithmetic (folded LayerNorm, bias, QuickGELU, residual) -> bf16 -> 128B-swizzled 128x64 staging
//     tile -> ONE TMA store per sub-tile (full 128-byte lines, rows beyond M clipped by the hardware).  The residual
//     sub-tiles of the NEXT tile are TMA-loaded into the staging tiles while the current tile finishes and are updated
//     in place, so the epilogue issues no per-thread global loads or stores at all.
//   patch-embed (+ positional embedding, rows shifted by one class-token row per frame): staged the same way, copied
//     out by the warp group with full 128-byte lines per row (the row shift rules out a TMA store).
//   direct (fp32 scores, target-score extraction, rank counting): registers -> global, with the TMEM load of the next
//     32-column chunk in flight while the current one is consumed.
//
// This one kernel serves every dense contraction of the hot path (reference call sites in SURVEY.md 2.2):
//   K1 patch-embed, K3 QKV, K5 out-proj(+residual), K6 MLP fc1(+QuickGELU)/fc2(+residual), K11/K12 similarity,
//   K13 rank counting fused into the similarity epilogue (the similarity matrix is never materialised).
#include <stdlib.h>

#include "gemm.cuh"
#include "ptx.cuh"
#include "tmap.cuh"

namespace fc {

namespace {

#ifndef GEMM_STAGES
#define GEMM_STAGES 5  // 5 x 32 KiB operand stages + 4 x 16 KiB staging tiles fill the 227 KiB of shared memory
#endif
constexpr int BM = 128, BN = 256, BK = 64, STAGES = GEMM_STAGES, UMMA_K = 16;
constexpr int A_BYTES = BM * BK * 2;        // this CTA's 128 rows of the pair's 256-row A tile
constexpr int B_BYTES = (BN / 2) * BK * 2;  // this CTA's half (128 of 256 rows) of the B tile
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int EPI_THREADS = 256;               // 8 epilogue warps = 2 groups of 128
constexpr int SUB_N = 64;                      // staged sub-tile width (64 bf16 = one 128-byte swizzle row)
constexpr int STAGING_BYTES = BM * SUB_N * 2;  // 16 KiB; every warp group owns TWO (one per sub-tile of a tile)
constexpr int OFF_STAGING = STAGES * STAGE_BYTES;
constexpr int OFF_BIAS = OFF_STAGING + 4 * STAGING_BYTES;
constexpr int OFF_BARS = OFF_BIAS + 2 * BN * 4;  // bias[256] + colsum[256]
constexpr int SMEM_BYTES = OFF_BARS + 256;  // 2*STAGES + 8 mbarriers + the TMEM slot
static_assert((2 * STAGES + 8) * 8 + 4 <= 256, "barrier block overflows its reservation");
constexpr int NUM_THREADS = 128 + EPI_THREADS;
constexpr uint32_t TMEM_COLS = 512;
static_assert(SMEM_BYTES <= 232448, "exceeds the 227 KiB per-CTA shared memory limit");

#ifdef FC_GEMM_TIMING
// Diagnostics build only (make TIMING=1): cycle counters summed over all CTAs.
//   [0] MMA thread: whole loop  [1] MMA: waiting for a drained accumulator  [2] MMA: waiting for TMA bytes
//   [3] producer: waiting for a free stage  [4] epilogue thread 0: whole loop  [5] epilogue: waiting for accumulators
//   [6] tiles  [7] MMA threads
__device__ unsigned long long g_gemm_timing[8];
#define FC_T(...) __VA_ARGS__
#else
#define FC_T(...)
#endif

__device__ __forceinline__ float quick_gelu(float v) {
  // x * sigmoid(1.702 x)  (reference twin: aligner/encoder/slip.py:359-361), written with one MUFU op:
  // sigmoid(y) = 0.5 + 0.5 tanh(y/2); tanh.approx is ~2^-11 accurate, far inside the bf16 output rounding (2^-9).
  const float h = 0.5f * v;
  return fmaf(h, tanh_approx(0.851f * v), h);
}

// MAJ: bit 0 = A is MN-major (stored [K rows][M contiguous]), bit 1 = B is MN-major (stored [K rows][N contiguous]).
// An MN-major 128 x 64 operand tile is fetched as two 64 (MN) x 64 (K) boxes of 8 KiB: K rows of 128 bytes, 128-byte
// swizzle, LBO = 8 KiB between the two MN halves, SBO = 1 KiB between groups of 8 K rows, +2 KiB per UMMA_K step.  The
// backward GEMMs use it to read dY / X / W exactly as the forward pass stored them (no transposed copies).
template <int EPI, int MAJ = 0>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_bf16_tn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmR,
                    const GemmParams p) {
  constexpr bool kLn = (EPI == EPI_LN_BIAS || EPI == EPI_LN_BIAS_QGELU || EPI == EPI_LN_BIAS_GELU);
  constexpr bool kErfGelu = (EPI == EPI_LN_BIAS_GELU);
  constexpr bool kGelu = (EPI == EPI_BIAS_QGELU || EPI == EPI_LN_BIAS_QGELU);
  constexpr bool kResid = (EPI == EPI_BIAS_RESID || EPI == EPI_QGELU_BWD);  // a second bf16 input tile, TMA-loaded into staging
  constexpr bool kStaged = (EPI == EPI_BIAS || EPI == EPI_BIAS_QGELU || kResid || kLn);
  constexpr bool kAmn = (MAJ & 1) != 0, kBmn = (MAJ & 2) != 0;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + OFF_BARS);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint64_t* resid_bar = tmem_empty + 2;  // [2][2]: per epilogue warp group, per staging tile
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(resid_bar + 4);
  float* sbias = reinterpret_cast<float*>(smem + OFF_BIAS);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_m_tiles = (p.M + BM - 1) / BM;
  const int num_n_tiles = (p.N + BN - 1) / BN;
  const int num_k = (p.K + BK - 1) / BK;
  // CTA pairs (cluster of 2): the pair works on two vertically adjacent 128-row tiles of the SAME 256-column block, so
  // each CTA fetches only half of the B tile and multicasts it into both CTAs' shared memory.
  const uint32_t cta_rank = cluster_ctarank();
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
  // pair-tiles; tile -> (m pair, n block), N fastest.  EPI_F32_SPLITK: `k_splits` work items per pair-tile (item ->
  // (k split, m pair, n block)), each accumulating its K range into a zero-initialised fp32 C with red.global.add.
  constexpr bool kSplit = (EPI == EPI_F32_SPLITK);
  const int num_m_pairs = (num_m_tiles + 1) / 2;
  const int num_tiles = num_m_pairs * num_n_tiles * (kSplit ? p.k_splits : 1);
  auto m_pair_of = [&](int t) { return kSplit ? (t / num_n_tiles) % num_m_pairs : t / num_n_tiles; };
  auto k_begin = [&](int t) {
    return kSplit ? static_cast<int>(static_cast<int64_t>(t / (num_n_tiles * num_m_pairs)) * num_k / p.k_splits) : 0;
  };
  auto k_end = [&](int t) {
    return kSplit ? static_cast<int>(static_cast<int64_t>(t / (num_n_tiles * num_m_pairs) + 1) * num_k / p.k_splits)
                  : num_k;
  };

  griddep_launch_dependents();  // PDL: the next kernel's prologue may overlap this kernel's tail
  if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0) __trap();  // swizzled tiles need a 1024-byte aligned base
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (kStaged) tma_prefetch_desc(&tmC);
    if (kResid) tma_prefetch_desc(&tmR);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);   // (leader's is the one in use) armed with the bytes of BOTH CTAs' loads
      mbar_init(&empty_bar[i], 1);  // released by the leader's multicast tcgen05.commit
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 2 * EPI_THREADS);  // (leader's) both CTAs' epilogue threads arrive
      mbar_init(&resid_bar[2 * i], 1);
      mbar_init(&resid_bar[2 * i + 1], 1);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc_2cta<TMEM_COLS>(tmem_slot);  // same warp in both CTAs, same address in both
  tc_fence_before();
  cluster_sync_all();  // the peer's barriers must exist before anything is multicast into this CTA
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  griddep_wait();  // PDL: everything above overlapped the previous kernel; from here on we touch its outputs

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      FC_T(long long w_stage = 0;)
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
        const int m_blk = 2 * m_pair_of(tile) + cta_rank, n_blk = tile % num_n_tiles;
        const int kb1 = k_end(tile);
        for (int kb = k_begin(tile); kb < kb1; ++kb) {
          FC_T(const long long tw = clock64();)
          mbar_wait(&empty_bar[stage], phase ^ 1);
          FC_T(w_stage += clock64() - tw;)
          uint8_t* sA = smem + stage * STAGE_BYTES;
          uint8_t* sB = sA + A_BYTES;
          // both CTAs' bytes are credited to the LEADER's full barrier, which alone gates the pair's MMAs
          if (cta_rank == 0) mbar_expect_tx(&full_bar[stage], 2 * STAGE_BYTES);
          const uint32_t leader_full = mapa_shared(smem_u32(&full_bar[stage]), 0);
          if (kAmn) {
            tma_load_2d_2cta(sA, &tmA, leader_full, m_blk * BM, kb * BK);
            tma_load_2d_2cta(sA + A_BYTES / 2, &tmA, leader_full, m_blk * BM + 64, kb * BK);
          } else {
            tma_load_2d_2cta(sA, &tmA, leader_full, kb * BK, m_blk * BM);
          }
          if (kBmn) {
            tma_load_2d_2cta(sB, &tmB, leader_full, n_blk * BN + cta_rank * (BN / 2), kb * BK);
            tma_load_2d_2cta(sB + B_BYTES / 2, &tmB, leader_full, n_blk * BN + cta_rank * (BN / 2) + 64, kb * BK);
          } else {
            tma_load_2d_2cta(sB, &tmB, leader_full, kb * BK, n_blk * BN + cta_rank * (BN / 2));
          }
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
      FC_T(atomicAdd(&g_gemm_timing[3], static_cast<unsigned long long>(w_stage));)
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0 && cta_rank == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16_f32(2 * BM, BN) | (kAmn ? (1u << 15) : 0u) | (kBmn ? (1u << 16) : 0u);  // M = 256 across the CTA pair
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      FC_T(const long long t_begin = clock64(); long long w_acc = 0, w_tma = 0; long long n_tiles = 0;)
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
        FC_T(long long tw = clock64();)
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);  // epilogue has drained this accumulator stage
        FC_T(w_acc += clock64() - tw; ++n_tiles;)
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        const int kb0 = k_begin(tile), kb1 = k_end(tile);
        for (int kb = kb0; kb < kb1; ++kb) {
          FC_T(tw = clock64();)
          mbar_wait(&full_bar[stage], phase);  // TMA bytes have landed
          FC_T(w_tma += clock64() - tw;)
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + stage * STAGE_BYTES);
          const uint32_t b_addr = a_addr + A_BYTES;
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            // advance along K inside the 128-byte swizzle atom: +32 bytes per UMMA_K
            const uint64_t a_desc = kAmn ? umma_desc_mn_sw128_lbo(a_addr + k * UMMA_K * 128, A_BYTES / 2)
                                         : umma_desc_k_sw128(a_addr + k * UMMA_K * 2);
            const uint64_t b_desc = kBmn ? umma_desc_mn_sw128_lbo(b_addr + k * UMMA_K * 128, B_BYTES / 2)
                                         : umma_desc_k_sw128(b_addr + k * UMMA_K * 2);
            umma_bf16_ss_2cta(d_tmem, a_desc, b_desc, idesc, ((kb - kb0) | k) != 0);
          }
          umma_commit_2cta_multicast(&empty_bar[stage], 0x3);  // frees the slot in BOTH CTAs when these MMAs retire
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit_2cta_multicast(&tmem_full[acc], 0x3);  // accumulators ready for both CTAs' epilogues
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
      FC_T(atomicAdd(&g_gemm_timing[0], static_cast<unsigned long long>(clock64() - t_begin));
           atomicAdd(&g_gemm_timing[1], static_cast<unsigned long long>(w_acc));
           atomicAdd(&g_gemm_timing[2], static_cast<unsigned long long>(w_tma));
           atomicAdd(&g_gemm_timing[6], static_cast<unsigned long long>(n_tiles));
           atomicAdd(&g_gemm_timing[7], 1ull);)
    }
  } else if (warp >= 4) {
    // ===================== epilogue (8 warps) =====================
    // Warp w may only read TMEM lanes [32*(w%4), +32); thread -> one output row of the tile.
    const int q = warp & 3;
    const int grp = (warp - 4) >> 2;  // warp group 0 / 1
    const int row_in_tile = q * 32 + lane;
    const int etid = threadIdx.x - 128;  // 0..255
    int acc = 0;
    uint32_t acc_phase = 0;
    uint32_t resid_phase = 0;  // bit si: parity of staging tile si's residual barrier
    bool prev_two = true;  // the previous tile issued a store from BOTH staging tiles of this group (false: ragged N)
    const bool issuer = (etid & 127) == 0;  // one thread per warp group drives its TMA traffic
    // accumulator release goes to the LEADER's tmem_empty barriers (the leader's MMA thread waits for both epilogues)
    const uint32_t leader_tmem_empty[2] = {mapa_shared(smem_u32(&tmem_empty[0]), 0),
                                           mapa_shared(smem_u32(&tmem_empty[1]), 0)};
    // Two staging tiles per warp group (sub-tile si of a tile -> tile si): the TMA store of one drains while the
    // other is being filled, and BOTH residual sub-tiles of the next tile are fetched while this one is finished.
    uint8_t* const stg_base = smem + OFF_STAGING + grp * 2 * STAGING_BYTES;
    const int sw = row_in_tile & 7;
    // (issuer) TMA-load the residual sub-tiles of pair-tile t into this group's staging tiles.  `drained`: the caller
    // has made sure no earlier TMA store still reads them.
    auto prefetch_resid = [&](int t, bool two_pending) {
      if (!kResid || t >= num_tiles) return;
      const int mb = 2 * (t / num_n_tiles) + cta_rank, nb = t % num_n_tiles;
#pragma unroll
      for (int si = 0; si < 2; ++si) {
        const int c0 = nb * BN + (grp + 2 * si) * SUB_N;
        if (c0 >= p.N) break;
        // stores were committed in the order tile 0, tile 1: tile 0 is free once at most one group is pending
        if (si == 0 && two_pending) bulk_wait_group_read<1>(); else bulk_wait_group_read<0>();
        mbar_expect_tx(&resid_bar[2 * grp + si], STAGING_BYTES);
        tma_load_2d(stg_base + si * STAGING_BYTES, &tmR, &resid_bar[2 * grp + si], c0, mb * BM);
      }
    };
    if (issuer) prefetch_resid(cluster_id, false);

    // Values the next tile's epilogue needs from global memory (bias / column sums of its 256 columns, the LayerNorm
    // partial sums of this thread's row) are fetched one tile ahead, so their latency hides behind the current tile.
    float nxt_bias = 0.f, nxt_cs = 0.f, nxt_s1 = 0.f, nxt_s2 = 0.f;
    float4 nxt_st[8];  // raw partial sums of the next tile's row: summed only when that tile starts, so the loads
                       // stay in flight behind the current tile's work instead of stalling here
    const bool st_vec = kLn && (p.ln_parts & 1) == 0 && p.ln_parts <= 16;
    auto prefetch_tile = [&](int t) {
      if (!kStaged || t >= num_tiles) return;
      const int mb = 2 * (t / num_n_tiles) + cta_rank, nb = t % num_n_tiles;
      const int n = nb * BN + etid;
      nxt_bias = (EPI != EPI_QGELU_BWD && n < p.N) ? __ldg(p.bias + n) : 0.f;
      if (kLn) {
        nxt_cs = n < p.N ? __ldg(p.colsum + n) : 0.f;
        nxt_s1 = 0.f;
        nxt_s2 = 0.f;
        const int r = mb * BM + row_in_tile;
        if (st_vec) {
          const float4* ps = reinterpret_cast<const float4*>(p.ln_stats + static_cast<int64_t>(r < p.M ? r : 0) * p.ln_parts * 2);
#pragma unroll
          for (int i = 0; i < 8; ++i)
            nxt_st[i] = (r < p.M && 2 * i < p.ln_parts) ? __ldg(ps + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        } else if (r < p.M) {
          const float2* ps = reinterpret_cast<const float2*>(p.ln_stats) + static_cast<int64_t>(r) * p.ln_parts;
          for (int i = 0; i < p.ln_parts; ++i) {
            const float2 v = __ldg(ps + i);
            nxt_s1 += v.x;
            nxt_s2 += v.y;
          }
        }
      }
    };
    prefetch_tile(cluster_id);
    FC_T(const long long e_begin = clock64(); long long w_full = 0;)

    for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
      const int m_blk = 2 * m_pair_of(tile) + cta_rank, n_blk = tile % num_n_tiles;
      const int row = m_blk * BM + row_in_tile;
      const int n0 = n_blk * BN;
      const bool row_ok = row < p.M;

      if (kStaged) {
        // stage this tile's 256 bias (and folded-LayerNorm column-sum) values once
        named_bar_sync(1, EPI_THREADS);  // everyone is done with the previous tile's values
        sbias[etid] = nxt_bias;
        if (kLn) sbias[BN + etid] = nxt_cs;
        named_bar_sync(1, EPI_THREADS);
        // folded LayerNorm: this row's mean / rstd from the producer's partial sums
        float ln_rstd = 1.f, ln_shift = 0.f;  // out = rstd * acc + shift * colsum + bias,  shift = -rstd * mean
        if (kLn) {
          if (st_vec) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              nxt_s1 += nxt_st[i].x + nxt_st[i].z;
              nxt_s2 += nxt_st[i].y + nxt_st[i].w;
            }
          }
          const float inv_k = 1.f / static_cast<float>(p.K);
          const float mean = nxt_s1 * inv_k;
          const float var = fmaxf(nxt_s2 * inv_k - mean * mean, 0.f);
          ln_rstd = rsqrtf(var + p.ln_eps);
          ln_shift = -ln_rstd * mean;
        }
        prefetch_tile(tile + num_clusters);
        FC_T(const long long tw = clock64();)
        mbar_wait(&tmem_full[acc], acc_phase);
        FC_T(w_full += clock64() - tw;)
        tc_fence_after();
        bool released = false;
#pragma unroll 1
        for (int si = 0; si < 2; ++si) {
          const int sub = grp + 2 * si;  // 64-column sub-tile of this warp group
          const int col0 = n0 + sub * SUB_N;
          if (col0 >= p.N) break;  // uniform across the group
          uint8_t* const stg_ptr = stg_base + si * STAGING_BYTES;
          const uint32_t stg_row = smem_u32(stg_ptr) + row_in_tile * 128;
          // (a) staging tile si is free once the store issued from it one tile ago has finished reading it (the
          //     residual path learns that -- and that the residual has landed -- from resid_bar below)
          if (!kResid) {
            if (issuer) {
              if (prev_two) bulk_wait_group_read<1>(); else bulk_wait_group_read<0>();
            }
            named_bar_sync(2 + grp, 128);
          }
          // (b) accumulators: 64 fp32 columns of this thread's row
          uint32_t r0[32], r1[32];
          const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN + sub * SUB_N;
          tmem_ld_32x32b_x32(t_addr, r0);
          tmem_ld_32x32b_x32(t_addr + 32, r1);
          tmem_ld_wait_fence(r0);
          tmem_ld_wait_fence(r1);
          if (si == 1 || col0 + 2 * SUB_N >= p.N) {  // last TMEM read of this tile by this thread
            tc_fence_before();
            mbar_arrive_cluster(leader_tmem_empty[acc]);
            released = true;
          }
          if (kResid) {
            mbar_wait(&resid_bar[2 * grp + si], (resid_phase >> si) & 1u);
            resid_phase ^= 1u << si;
          }
          const float* bias_s = sbias + sub * SUB_N;
          // Packed fp32x2 arithmetic (FFMA2 / FADD2 / FMUL2): the epilogue's CUDA-core work is a visible share of the
          // energy of the K = 768 GEMMs, and two columns per instruction halve its issue slots.
          uint64_t st1_2 = pack_f32x2(0.f, 0.f), st2_2 = st1_2;
          const uint64_t rstd2 = pack_f32x2(ln_rstd, ln_rstd), shift2 = pack_f32x2(ln_shift, ln_shift);
          const uint64_t half2 = pack_f32x2(0.5f, 0.5f), k2 = pack_f32x2(0.851f, 0.851f);
#pragma unroll
          for (int c = 0; c < 8; ++c) {  // 8 chunks of 8 columns = 16 bytes of bf16 each
            const uint32_t(&rr)[32] = c < 4 ? r0 : r1;
            const int o = (c & 3) * 8;
            const float4 b0 = *reinterpret_cast<const float4*>(bias_s + c * 8);
            const float4 b1 = *reinterpret_cast<const float4*>(bias_s + c * 8 + 4);
            const uint64_t bias2[4] = {pack_f32x2(b0.x, b0.y), pack_f32x2(b0.z, b0.w), pack_f32x2(b1.x, b1.y),
                                       pack_f32x2(b1.z, b1.w)};
            uint64_t v2[4];
            if (kLn) {
              const float4 c0 = *reinterpret_cast<const float4*>(bias_s + BN + c * 8);
              const float4 c1 = *reinterpret_cast<const float4*>(bias_s + BN + c * 8 + 4);
              const uint64_t cs2[4] = {pack_f32x2(c0.x, c0.y), pack_f32x2(c0.z, c0.w), pack_f32x2(c1.x, c1.y),
                                       pack_f32x2(c1.z, c1.w)};
#pragma unroll
              for (int e = 0; e < 4; ++e)  // rstd * acc + (shift * colsum + bias)
                v2[e] = fma_f32x2(rstd2, pack_f32x2(__uint_as_float(rr[o + 2 * e]), __uint_as_float(rr[o + 2 * e + 1])),
                                  fma_f32x2(shift2, cs2[e], bias2[e]));
            } else {
#pragma unroll
              for (int e = 0; e < 4; ++e)
                v2[e] = add_f32x2(pack_f32x2(__uint_as_float(rr[o + 2 * e]), __uint_as_float(rr[o + 2 * e + 1])), bias2[e]);
            }
            if (kGelu) {
              // x * sigmoid(1.702 x) = h + h * tanh(0.851 x), h = x / 2  (see quick_gelu)
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const uint64_t h2 = mul_f32x2(v2[e], half2);
                float a0, a1;
                unpack_f32x2(mul_f32x2(v2[e], k2), a0, a1);
                v2[e] = fma_f32x2(h2, pack_f32x2(tanh_approx(a0), tanh_approx(a1)), h2);
              }
            }
            if (kErfGelu) {
              // nn.GELU() of timm's Mlp: x Phi(x) = max(x, 0) - |x| erfc(|x| / sqrt 2) / 2, with erfc from Abramowitz & Stegun
              // 7.1.26 (|error| <= 1.5e-7, relative accuracy kept in the negative tail): t = 1 / (1 + p |u|),
              // erfc(|u|) = (a1 t + ... + a5 t^5) exp(-u^2).  Two MUFU ops (rcp, ex2) and packed Horner steps per element
              // pair; CUDA's erff in this place made the epilogue longer than the K = 768 main loop (+16 % on the step).
              const uint64_t ch2 = pack_f32x2(-0.72134752044448170f, -0.72134752044448170f);  // -log2(e) / 2
              const uint64_t a5 = pack_f32x2(0.5307027145f, 0.5307027145f), a4 = pack_f32x2(-0.7265760135f, -0.7265760135f),
                             a3 = pack_f32x2(0.7107068705f, 0.7107068705f), a2 = pack_f32x2(-0.142248368f, -0.142248368f),
                             a1 = pack_f32x2(0.127414796f, 0.127414796f);  // the series' coefficients, halved
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                float x0, x1, s0, s1;
                unpack_f32x2(v2[e], x0, x1);
                unpack_f32x2(mul_f32x2(mul_f32x2(v2[e], v2[e]), ch2), s0, s1);
                const float ax0 = fabsf(x0), ax1 = fabsf(x1);
                const uint64_t t2 = pack_f32x2(rcp_approx(fmaf(0.23164189f, ax0, 1.f)), rcp_approx(fmaf(0.23164189f, ax1, 1.f)));
                uint64_t q2 = fma_f32x2(a5, t2, a4);
                q2 = fma_f32x2(q2, t2, a3);
                q2 = fma_f32x2(q2, t2, a2);
                q2 = fma_f32x2(q2, t2, a1);
                q2 = mul_f32x2(mul_f32x2(q2, t2), pack_f32x2(ex2_approx(s0), ex2_approx(s1)));  // erfc(|u|) / 2
                v2[e] = fma_f32x2(pack_f32x2(-ax0, -ax1), q2, pack_f32x2(fmaxf(x0, 0.f), fmaxf(x1, 0.f)));
              }
            }
            const uint32_t addr = stg_row + ((c ^ sw) << 4);
            if (EPI == EPI_BIAS_RESID) {
              const uint4 u = ld_shared_v4(addr);
              const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float2 f = unpack_bf16x2(w[e]);
                v2[e] = add_f32x2(v2[e], pack_f32x2(f.x, f.y));
              }
            }
            if (EPI == EPI_QGELU_BWD) {
              // d/du [u s(u)] = s + 1.702 u s (1 - s) with s = sigmoid(1.702 u) = 0.5 + 0.5 tanh(0.851 u); the staging tile
              // holds u (the same formula as quickgelu_bwd_kernel, applied to the fp32 accumulator instead of a bf16 dg)
              const uint4 u = ld_shared_v4(addr);
              const uint32_t w[4] = {u.x, u.y, u.z, u.w};
              const uint64_t one2 = pack_f32x2(1.f, 1.f), k1702 = pack_f32x2(1.702f, 1.702f);
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float2 f = unpack_bf16x2(w[e]);
                const uint64_t u2 = pack_f32x2(f.x, f.y);
                const uint64_t s2 = fma_f32x2(half2, pack_f32x2(tanh_approx(0.851f * f.x), tanh_approx(0.851f * f.y)), half2);
                // s * (1 + 1.702 u (1 - s))
                const uint64_t oms2 = fma_f32x2(s2, pack_f32x2(-1.f, -1.f), one2);
                const uint64_t d2 = mul_f32x2(s2, fma_f32x2(mul_f32x2(k1702, u2), oms2, one2));
                v2[e] = mul_f32x2(v2[e], d2);
              }
            }
            float v[8];
#pragma unroll
            for (int e = 0; e < 4; ++e) unpack_f32x2(v2[e], v[2 * e], v[2 * e + 1]);
            uint4 u;
            u.x = pack_bf16x2(v[0], v[1]);
            u.y = pack_bf16x2(v[2], v[3]);
            u.z = pack_bf16x2(v[4], v[5]);
            u.w = pack_bf16x2(v[6], v[7]);
            st_shared_v4(addr, u);
            if (EPI == EPI_BIAS_RESID) {
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                st1_2 = add_f32x2(st1_2, v2[e]);
                st2_2 = fma_f32x2(v2[e], v2[e], st2_2);
              }
            }
          }
          float st1 = 0.f, st2 = 0.f;
          if (EPI == EPI_BIAS_RESID) {
            float a0, a1, q0, q1;
            unpack_f32x2(st1_2, a0, a1);
            unpack_f32x2(st2_2, q0, q1);
            st1 = a0 + a1;
            st2 = q0 + q1;
          }
          if (EPI == EPI_BIAS_RESID && p.stats_out != nullptr && row_ok)  // row statistics for the next folded LayerNorm
            *reinterpret_cast<float2*>(p.stats_out + (static_cast<int64_t>(row) * (p.N / SUB_N) + (col0 / SUB_N)) * 2) =
                make_float2(st1, st2);
          // (c) hand the finished sub-tile to the TMA engine
          fence_proxy_async_smem();
          named_bar_sync(2 + grp, 128);
          if (issuer) {
            tma_store_2d(&tmC, stg_ptr, col0, m_blk * BM);
            bulk_commit_group();
          }
        }
        if (!released) {  // group had no sub-tile inside N (ragged N): it still owes the accumulator release
          tc_fence_before();
          mbar_arrive_cluster(leader_tmem_empty[acc]);
        }
        // fetch both residual sub-tiles of the next tile now: their latency hides behind the accumulator hand-over
        prev_two = n0 + (grp + 2) * SUB_N < p.N;
        if (issuer) prefetch_resid(tile + num_clusters, prev_two);
      } else if (EPI == EPI_PATCH && (p.N % SUB_N) == 0) {
        // ---------------- patch-embed epilogue: + positional embedding, rows shifted by one class-token row per frame.
        // The shift rules out a TMA store (a 128-row tile straddles frames), so the bf16 sub-tile is staged in shared
        // memory and copied out by the warp group with full 128-byte lines per row (8 threads x 16 bytes).
        mbar_wait(&tmem_full[acc], acc_phase);
        tc_fence_after();
        const float* pos_row = nullptr;
        if (row_ok) {
          const int f = row / p.patches_per_frame, pp = row - f * p.patches_per_frame;
          pos_row = p.pos + static_cast<int64_t>(pp + 1) * p.N;
        }
        const int gt = etid & 127;  // thread within the warp group
        bool released = false;
#pragma unroll 1
        for (int si = 0; si < 2; ++si) {
          const int sub = grp + 2 * si;
          const int col0 = n0 + sub * SUB_N;
          if (col0 >= p.N) break;  // uniform across the group
          uint8_t* const stg_ptr = stg_base + si * STAGING_BYTES;
          const uint32_t stg_row = smem_u32(stg_ptr) + row_in_tile * 128;
          uint32_t r0[32], r1[32];
          const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN + sub * SUB_N;
          tmem_ld_32x32b_x32(t_addr, r0);
          tmem_ld_32x32b_x32(t_addr + 32, r1);
          tmem_ld_wait_fence(r0);
          tmem_ld_wait_fence(r1);
          if (si == 1 || col0 + 2 * SUB_N >= p.N) {
            tc_fence_before();
            mbar_arrive_cluster(leader_tmem_empty[acc]);
            released = true;
          }
          if (row_ok) {
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              const uint32_t(&rr)[32] = c < 4 ? r0 : r1;
              const int o = (c & 3) * 8;
              const float4 a0 = __ldg(reinterpret_cast<const float4*>(pos_row + col0 + c * 8));
              const float4 a1 = __ldg(reinterpret_cast<const float4*>(pos_row + col0 + c * 8 + 4));
              uint4 u;
              u.x = pack_bf16x2(__uint_as_float(rr[o + 0]) + a0.x, __uint_as_float(rr[o + 1]) + a0.y);
              u.y = pack_bf16x2(__uint_as_float(rr[o + 2]) + a0.z, __uint_as_float(rr[o + 3]) + a0.w);
              u.z = pack_bf16x2(__uint_as_float(rr[o + 4]) + a1.x, __uint_as_float(rr[o + 5]) + a1.y);
              u.w = pack_bf16x2(__uint_as_float(rr[o + 6]) + a1.z, __uint_as_float(rr[o + 7]) + a1.w);
              st_shared_v4(stg_row + ((c ^ sw) << 4), u);
            }
          }
          named_bar_sync(2 + grp, 128);  // the sub-tile is complete in shared memory
          {
            const int chunk = gt & 7;
            int r = gt >> 3;                 // rows r, r + 16, ... of the tile
            int grow = m_blk * BM + r;       // global patch row
            int f = grow / p.patches_per_frame;
            int rem = grow - f * p.patches_per_frame;
            bf16* const cbase = static_cast<bf16*>(p.C) + col0 + chunk * 8;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              if (grow < p.M) {
                const uint4 u = ld_shared_v4(smem_u32(stg_ptr) + r * 128 + ((chunk ^ (r & 7)) << 4));
                *reinterpret_cast<uint4*>(cbase + (static_cast<int64_t>(grow) + f + 1) * p.ldc) = u;
              }
              r += 16;
              grow += 16;
              rem += 16;
              while (rem >= p.patches_per_frame) {
                rem -= p.patches_per_frame;
                ++f;
              }
            }
          }
          named_bar_sync(2 + grp, 128);  // the staging tile may be overwritten
        }
        if (!released) {
          tc_fence_before();
          mbar_arrive_cluster(leader_tmem_empty[acc]);
        }
      } else {
        // ---------------- direct epilogues ----------------
        mbar_wait(&tmem_full[acc], acc_phase);
        tc_fence_after();
        const int half = grp;
        const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN + half * (BN / 2);
        int64_t out_row = row;
        const float* pos_row = nullptr;
        int tgt = 0;
        float tsc = 0.f;
        int cnt = 0;
        if (EPI == EPI_PATCH && row_ok) {
          const int f = row / p.patches_per_frame, pp = row - f * p.patches_per_frame;
          out_row = static_cast<int64_t>(row) + f + 1;  // skip one class-token row per frame
          pos_row = p.pos + static_cast<int64_t>(pp + 1) * p.N;
        }
        if ((EPI == EPI_TARGET || EPI == EPI_COUNT) && row_ok) {
          tgt = p.target[row];
          if (EPI == EPI_COUNT) tsc = p.target_score[row];
        }
        const int cbase = n0 + half * (BN / 2);
        constexpr int NCH = BN / 2 / 32;  // 4 chunks of 32 columns per thread
        uint32_t r[2][32];
        tmem_ld_32x32b_x32(t_addr, r[0]);
#pragma unroll
        for (int i = 0; i < NCH; ++i) {
          uint32_t(&rc)[32] = r[i & 1];
          tmem_ld_wait_fence(rc);
          if (i + 1 < NCH) tmem_ld_32x32b_x32(t_addr + (i + 1) * 32, r[(i + 1) & 1]);
          const int col0 = cbase + i * 32;
          if (!row_ok || col0 >= p.N) continue;

          if (EPI == EPI_PATCH) {
            // N is a multiple of 32 on this path (checked on the host): full, 16-byte aligned chunks.
            float v[32];
            const float4* add4 = reinterpret_cast<const float4*>(pos_row + col0);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 b = __ldg(add4 + j);
              v[4 * j + 0] = __uint_as_float(rc[4 * j + 0]) + b.x;
              v[4 * j + 1] = __uint_as_float(rc[4 * j + 1]) + b.y;
              v[4 * j + 2] = __uint_as_float(rc[4 * j + 2]) + b.z;
              v[4 * j + 3] = __uint_as_float(rc[4 * j + 3]) + b.w;
            }
            uint4* out4 = reinterpret_cast<uint4*>(static_cast<bf16*>(p.C) + out_row * p.ldc + col0);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint4 u;
              u.x = pack_bf16x2(v[8 * j + 0], v[8 * j + 1]);
              u.y = pack_bf16x2(v[8 * j + 2], v[8 * j + 3]);
              u.z = pack_bf16x2(v[8 * j + 4], v[8 * j + 5]);
              u.w = pack_bf16x2(v[8 * j + 6], v[8 * j + 7]);
              out4[j] = u;
            }
          } else if (EPI == EPI_F32_SPLITK) {
            float* out = static_cast<float*>(p.C) + static_cast<int64_t>(row) * p.ldc + col0;
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (col0 + j < p.N) atomicAdd(out + j, p.alpha * __uint_as_float(rc[j]));
          } else if (EPI == EPI_F32) {
            float* out = static_cast<float*>(p.C) + static_cast<int64_t>(row) * p.ldc + col0;
            if (col0 + 32 <= p.N && (p.ldc & 3) == 0) {
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                float4 o;
                o.x = p.alpha * __uint_as_float(rc[4 * j + 0]);
                o.y = p.alpha * __uint_as_float(rc[4 * j + 1]);
                o.z = p.alpha * __uint_as_float(rc[4 * j + 2]);
                o.w = p.alpha * __uint_as_float(rc[4 * j + 3]);
                reinterpret_cast<float4*>(out)[j] = o;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (col0 + j < p.N) out[j] = p.alpha * __uint_as_float(rc[j]);
            }
          } else if (EPI == EPI_TARGET) {
            const int local = tgt - p.col_offset - col0;
            if (local >= 0 && local < 32 && col0 + local < p.N) {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (j == local) p.tscore_out[row] = __uint_as_float(rc[j]);
            }
          } else if (EPI == EPI_COUNT) {
            const int gcol0 = p.col_offset + col0;
            const int nvalid = min(32, p.N - col0);
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float s = __uint_as_float(rc[j]);
              const bool hit = (s > tsc) || (s == tsc && (gcol0 + j) < tgt);
              cnt += (j < nvalid && hit) ? 1 : 0;
            }
          }
        }
        if (EPI == EPI_COUNT && row_ok && cnt) atomicAdd(p.counts + row, cnt);
        tc_fence_before();
        mbar_arrive_cluster(leader_tmem_empty[acc]);
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    if (kStaged && issuer) bulk_wait_group<0>();  // smem must outlive the last TMA store's reads
    FC_T(if (etid == 0 && cta_rank == 0) {
      atomicAdd(&g_gemm_timing[4], static_cast<unsigned long long>(clock64() - e_begin));
      atomicAdd(&g_gemm_timing[5], static_cast<unsigned long long>(w_full));
    })
  }

  __syncwarp();
  tc_fence_before();
  cluster_sync_all();  // neither CTA may exit while its peer can still multicast into it or signal its barriers
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_2cta<TMEM_COLS>(tmem_base);
  }
}

// ---- host side: tensor maps (memoised per (pointer, shape): tmap.cuh) + launch ----
// rows x cols bf16 matrix, cols contiguous, row stride ld elements; box = box_rows x 64 elements, 128B swizzle.
int make_tmap(CUtensorMap* tm, const bf16* base, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
  const uint64_t dims[2] = {static_cast<uint64_t>(cols), static_cast<uint64_t>(rows)};
  const uint64_t strides[1] = {static_cast<uint64_t>(ld) * sizeof(bf16)};
  const uint32_t box[2] = {BK, static_cast<uint32_t>(box_rows)};
  return tmap_bf16_sw128(tm, base, 2, dims, strides, box);
}

// MN-major operand: stored [k_rows][mn_cols] with mn contiguous; box = 64 (MN) x 64 (K rows), 128B swizzle.
int make_tmap_mn(CUtensorMap* tm, const bf16* base, int64_t k_rows, int64_t mn_cols, int64_t ld) {
  const uint64_t dims[2] = {static_cast<uint64_t>(mn_cols), static_cast<uint64_t>(k_rows)};
  const uint64_t strides[1] = {static_cast<uint64_t>(ld) * sizeof(bf16)};
  const uint32_t box[2] = {64, BK};
  return tmap_bf16_sw128(tm, base, 2, dims, strides, box);
}

template <int EPI, int MAJ = 0>
int launch(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, const CUtensorMap& tr,
           const GemmParams& p, cudaStream_t stream) {
  static bool configured = false;
  if (!configured) {
    FC_CUDA(cudaFuncSetAttribute(gemm_bf16_tn_kernel<EPI, MAJ>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    configured = true;
  }
  const int pair_tiles = (((p.M + BM - 1) / BM + 1) / 2) * ((p.N + BN - 1) / BN) * (EPI == EPI_F32_SPLITK ? p.k_splits : 1);
  static int max_clusters = 0;
  if (!max_clusters) {
    max_clusters = num_sms() / 2;
    if (const char* e = getenv("FC_GEMM_MAX_CLUSTERS")) {  // diagnostics: run the persistent grid on fewer SM pairs
      const int v = atoi(e);
      if (v > 0 && v < max_clusters) max_clusters = v;
    }
  }
  const int clusters = pair_tiles < max_clusters ? pair_tiles : max_clusters;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * clusters);
  cfg.blockDim = dim3(NUM_THREADS);
  cfg.dynamicSmemBytes = SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 2;
  note_launch();
  FC_CUDA(cudaLaunchKernelEx(&cfg, gemm_bf16_tn_kernel<EPI, MAJ>, ta, tb, tc, tr, p));
  return FC_OK;
}

}  // namespace

#ifdef FC_GEMM_TIMING
extern "C" __attribute__((visibility("default"))) int fc_debug_gemm_timing(unsigned long long* out, int reset) {
  if (cudaDeviceSynchronize() != cudaSuccess) return -1;
  if (out && cudaMemcpyFromSymbol(out, g_gemm_timing, sizeof(g_gemm_timing)) != cudaSuccess) return -1;
  if (reset) {
    unsigned long long z[8] = {};
    if (cudaMemcpyToSymbol(g_gemm_timing, z, sizeof(z)) != cudaSuccess) return -1;
  }
  return 0;
}
#endif

int gemm_bf16_tn(int epilogue, const bf16* A, int64_t lda, const bf16* B, int64_t ldb, const GemmParams& p,
                 cudaStream_t stream) {
  FC_REQUIRE(A && B, "gemm: null operand");
  FC_REQUIRE(p.M > 0 && p.N > 0 && p.K > 0, "gemm: empty problem M=%d N=%d K=%d", p.M, p.N, p.K);
  FC_REQUIRE(lda % 8 == 0 && ldb % 8 == 0 && ((p.a_mn && p.b_mn) || p.K % 8 == 0),
             "gemm: lda/ldb (and K of a K-major operand) must be multiples of 8 (K=%d)", p.K);
  FC_REQUIRE((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(B) & 15) == 0,
             "gemm: operands must be 16-byte aligned");
  const bool ln = epilogue == EPI_LN_BIAS || epilogue == EPI_LN_BIAS_QGELU || epilogue == EPI_LN_BIAS_GELU;
  const bool resid_in = epilogue == EPI_BIAS_RESID || epilogue == EPI_QGELU_BWD;
  const bool staged = epilogue == EPI_BIAS || epilogue == EPI_BIAS_QGELU || resid_in || ln;
  if (ln) {
    FC_REQUIRE(p.C && p.N % 32 == 0 && p.ldc % 8 == 0 && (reinterpret_cast<uintptr_t>(p.C) & 15) == 0 && p.bias,
               "gemm: folded-LayerNorm epilogue needs bf16 C (N %% 32 == 0, ldc %% 8 == 0) and a bias");
    FC_REQUIRE(p.ln_stats && p.colsum && p.ln_parts > 0 && (reinterpret_cast<uintptr_t>(p.ln_stats) & 15) == 0,
               "gemm: folded-LayerNorm epilogue needs row statistics and column sums");
  } else if (epilogue <= EPI_PATCH) {
    FC_REQUIRE(p.C && p.N % 32 == 0 && p.ldc % 8 == 0, "gemm: bf16 epilogues need N %% 32 == 0 and ldc %% 8 == 0");
    FC_REQUIRE((reinterpret_cast<uintptr_t>(p.C) & 15) == 0, "gemm: C must be 16-byte aligned");
    if (epilogue == EPI_PATCH)
      FC_REQUIRE(p.pos && p.patches_per_frame > 0, "gemm: patch epilogue needs pos and patches_per_frame");
    else
      FC_REQUIRE(p.bias, "gemm: bias epilogue needs a bias vector");
    if (epilogue == EPI_BIAS_RESID) {
      FC_REQUIRE(p.resid && p.ldr % 8 == 0 && (reinterpret_cast<uintptr_t>(p.resid) & 15) == 0,
                 "gemm: residual must be 16-byte aligned with ldr %% 8 == 0");
      FC_REQUIRE(p.stats_out == nullptr || p.N % SUB_N == 0, "gemm: stats_out needs N %% 64 == 0");
    }
  } else if (epilogue == EPI_QGELU_BWD) {
    FC_REQUIRE(p.C && p.N % 32 == 0 && p.ldc % 8 == 0 && (reinterpret_cast<uintptr_t>(p.C) & 15) == 0,
               "gemm: bf16 epilogues need N %% 32 == 0, ldc %% 8 == 0 and a 16-byte aligned C");
    FC_REQUIRE(p.resid && p.ldr % 8 == 0 && (reinterpret_cast<uintptr_t>(p.resid) & 15) == 0,
               "gemm: the pre-activation input must be 16-byte aligned with ldr %% 8 == 0");
  } else if (epilogue == EPI_F32) {
    FC_REQUIRE(p.C, "gemm: null C");
  } else if (epilogue == EPI_F32_SPLITK) {
    FC_REQUIRE(p.C, "gemm: null C");
    FC_REQUIRE(p.k_splits >= 1 && p.k_splits <= (p.K + BK - 1) / BK, "gemm: k_splits=%d outside 1..%d", p.k_splits,
               (p.K + BK - 1) / BK);
  } else if (epilogue == EPI_TARGET) {
    FC_REQUIRE(p.target && p.tscore_out, "gemm: target epilogue needs target and tscore_out");
  } else if (epilogue == EPI_COUNT) {
    FC_REQUIRE(p.target && p.target_score && p.counts, "gemm: count epilogue needs target, target_score, counts");
  } else {
    FC_REQUIRE(false, "gemm: unknown epilogue %d", epilogue);
  }
  const double mn = static_cast<double>(p.M) * p.N;
  ProfScope prof(stream, PROF_GEMM, epilogue, p.M, p.N, p.K, 2.0 * mn * p.K,
                 2.0 * (static_cast<double>(p.M) + p.N) * p.K +
                     (epilogue <= EPI_PATCH || ln || resid_in ? 2.0 * mn : 0.0) * (resid_in ? 2.0 : 1.0) +
                     (epilogue == EPI_F32 || epilogue == EPI_F32_SPLITK ? 4.0 * mn : 0.0));
  CUtensorMap ta, tb, tc, tr;
  const int maj = (p.a_mn ? 1 : 0) | (p.b_mn ? 2 : 0);
  FC_REQUIRE(maj == 0 || (maj == 2 && (epilogue == EPI_BIAS || resid_in || epilogue == EPI_F32)) ||
                 (maj == 3 && epilogue == EPI_F32_SPLITK),
             "gemm: operand layout a_mn=%d b_mn=%d is not built for epilogue %d", p.a_mn, p.b_mn, epilogue);
  int rc = p.a_mn ? make_tmap_mn(&ta, A, p.K, p.M, lda) : make_tmap(&ta, A, p.M, p.K, lda, BM);
  if (rc) return rc;
  // each CTA of a pair loads half of the B tile
  rc = p.b_mn ? make_tmap_mn(&tb, B, p.K, p.N, ldb) : make_tmap(&tb, B, p.N, p.K, ldb, BN / 2);
  if (rc) return rc;
  tc = ta;
  tr = ta;
  if (staged) {
    rc = make_tmap(&tc, static_cast<const bf16*>(p.C), p.M, p.N, p.ldc, BM);
    if (rc) return rc;
    tr = tc;
    if (resid_in) {
      rc = make_tmap(&tr, p.resid, p.M, p.N, p.ldr, BM);
      if (rc) return rc;
    }
  }
  if (maj == 2) {
    if (epilogue == EPI_BIAS) return launch<EPI_BIAS, 2>(ta, tb, tc, tr, p, stream);
    if (epilogue == EPI_BIAS_RESID) return launch<EPI_BIAS_RESID, 2>(ta, tb, tc, tr, p, stream);
    if (epilogue == EPI_QGELU_BWD) return launch<EPI_QGELU_BWD, 2>(ta, tb, tc, tr, p, stream);
    return launch<EPI_F32, 2>(ta, tb, tc, tr, p, stream);
  }
  if (maj == 3) return launch<EPI_F32_SPLITK, 3>(ta, tb, tc, tr, p, stream);
  switch (epilogue) {
    case EPI_BIAS: return launch<EPI_BIAS>(ta, tb, tc, tr, p, stream);
    case EPI_BIAS_QGELU: return launch<EPI_BIAS_QGELU>(ta, tb, tc, tr, p, stream);
    case EPI_BIAS_RESID: return launch<EPI_BIAS_RESID>(ta, tb, tc, tr, p, stream);
    case EPI_PATCH: return launch<EPI_PATCH>(ta, tb, tc, tr, p, stream);
    case EPI_F32: return launch<EPI_F32>(ta, tb, tc, tr, p, stream);
    case EPI_F32_SPLITK: return launch<EPI_F32_SPLITK>(ta, tb, tc, tr, p, stream);
    case EPI_TARGET: return launch<EPI_TARGET>(ta, tb, tc, tr, p, stream);
    case EPI_LN_BIAS: return launch<EPI_LN_BIAS>(ta, tb, tc, tr, p, stream);
    case EPI_LN_BIAS_QGELU: return launch<EPI_LN_BIAS_QGELU>(ta, tb, tc, tr, p, stream);
    case EPI_LN_BIAS_GELU: return launch<EPI_LN_BIAS_GELU>(ta, tb, tc, tr, p, stream);
    case EPI_QGELU_BWD: return launch<EPI_QGELU_BWD>(ta, tb, tc, tr, p, stream);
    default: return launch<EPI_COUNT>(ta, tb, tc, tr, p, stream);
  }
}

}  // namespace fc
