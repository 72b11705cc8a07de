// In-batch cross-entropy backward, v3 (d <= 128): 16 epilogue warps, the softmax statistics ride in the score MMA.
//
// Math and scheduling as in ce.cu / ce_bwd2.cu:  acc[128, d] = sum_j E_j Y_j,  E_j = f(X Y_j^T), run once for dU
// (X = U, Y = V) and once for dV (X = V, Y = U); reference semantics src/two_tower_base_retrieval.py:287,301,310-312
// under autograd (train/train.py:124).  What v2's timeline showed (profiles/r02_ce_bwd_v2_timeline.txt): the tensor
// pipe needs 1024 cycles per 128x128 tile and the MUFU pipe 1024, but the 8 epilogue warps took 1450 (dU) / 1750 (dV)
// per tile plus 220 / 650 cycles of hand-shakes - two warps per scheduler cannot hide their own latencies, and the
// dV pass fetched (g, lse) of every column through shared memory with two named barriers per tile.  Changes:
//
//   * S' = X Y^T + b_user is produced by the tensor core itself: one extra K = 16 step whose user-side operand holds
//     the bf16 split (hi, mid, lo) of  b = -lse + ln|g|  and whose item-side operand holds ones.  The epilogue is
//     E = exp2(S' log2e + log2|g_scale|) for both passes: no per-row / per-column statistics, no multiply by g.
//     Signs: the sign of g_scale and (dU) of the row's g are XORed onto the packed bf16 pairs; (dV) a per-user sign
//     bitmask is applied to a 32-column chunk only when it is non-zero (never, with the clamped label weights).
//   * Two score buffers, released as soon as their values sit in registers, and two separate E buffers (bf16 pairs, the
//     TMEM A operand of  acc += E Y).  Earlier drafts kept E in place in the score buffer (two or three buffers): then
//     S(t + NB) has to wait for the COMPLETION of acc += E(t) Y(t), which the in-order tensor queue places right behind
//     S(t + NB - 1) - the pipe ran S, E Y, ~350 idle cycles of completion -> barrier -> wake-up -> issue, per tile
//     (profiles/r02_ce_bwd_v3_timelines.md).  With E elsewhere the score MMAs only wait for a tmem load.
//     TMEM (d = 128): 2 x 96 S + 2 x 48 E + 128 acc + 64 X + 8 bias step = 488 columns, hence 128 x 96 score tiles.
//   * BN / 32 column groups x 4 lane quarters of epilogue warps, 32 columns per thread and tile; the next segment's X
//     tile is staged into TMEM before the finished accumulator is drained.
//
// TMEM columns: S buffers [0, NB*BN) | E buffers [.., +NE*BN/2) | acc [.., +DP) | X [.., +DP/2) | X bias step [.., +8).
// Warp roles: 0 .. 4 EG - 1 epilogue (group e = warp / 4 owns columns [32e, 32e+32) of every score tile, q = warp % 4 the
// TMEM lane quarter), then TMA, UMMA issue of S = X Y^T, TMEM alloc, UMMA issue of acc += E Y (warp-uniform, elected lane).
#include <stdlib.h>

#include "ce_common.cuh"

namespace tt {

namespace {

#ifndef TT_CE3_CW
#define TT_CE3_CW 32
#endif

template <int DP>
struct Tile3 {
  static constexpr int BN = DP == 64 ? 128 : 96;  // score-tile width (columns of Y per tile); TMEM: 2 BN + BN + DP + DP/2 + 8 <= 512
  static constexpr int NB = 2;                    // score buffers (free again once loaded into registers)
  static constexpr int NE = 2;                    // E buffers (free again once acc += E Y has completed)
};

template <int DP, bool BIAS_X>
struct Cfg3 {
  static constexpr int BN = Tile3<DP>::BN;
  static constexpr int NB = Tile3<DP>::NB;
  static constexpr int NE = Tile3<DP>::NE;
  static constexpr int CW = TT_CE3_CW;                              // score columns per epilogue thread and tile (16 or 32)
  static constexpr int EG = BN / CW;                                // epilogue column groups (4 warps each)
  static constexpr int THREADS = 128 + EG * 128;
  static constexpr int XP = DP / 32;                                // 32-column parts of the X tile / accumulator
  static constexpr int KBOX = DP / 64;
  static constexpr int X_BYTES = 128 * DP * 2;
  static constexpr int EXT_BYTES = BN * 128;                        // one 128-byte-swizzled atom: BN rows x 64 bf16
  static constexpr int Y_MAIN = BN * DP * 2;
  static constexpr int Y_BYTES = Y_MAIN + (BIAS_X ? 0 : EXT_BYTES);  // dV: the users' bias step travels with the Y tile
  static constexpr int ONES_BYTES = BIAS_X ? 128 * 128 : 0;          // dU: constant ones tile on the item side
  static constexpr int STAGES = BIAS_X ? (DP == 64 ? 8 : 6) : (DP == 64 ? 5 : 5);
  static constexpr int SMEM_BYTES = X_BYTES + ONES_BYTES + STAGES * Y_BYTES + 1024 + 512;
  static constexpr int E_COL = NB * BN;
  static constexpr int ACC_COL = E_COL + NE * (BN / 2);
  static constexpr int X_COL = ACC_COL + DP;
  static constexpr int XE_COL = X_COL + DP / 2;
  static_assert(XE_COL + 8 <= 512, "TMEM budget");
  static_assert(SMEM_BYTES <= 232448, "shared memory budget");
  static_assert(BN % 32 == 0 && (BN * 128) % 1024 == 0, "tile width");
};

#ifdef TT_CE_BRINGUP
#define CE3_STAMP(role, tile, which)                                                                                   \
  do {                                                                                                                 \
    if (a.trace != nullptr && blockIdx.x == a.trace_cta && (tile) < 64) a.trace[((role) * 64 + (tile)) * 2 + (which)] = clock64(); \
  } while (0)
#define CE3_CTA_TIME(slot)                                                       \
  do {                                                                           \
    if (a.cta_times != nullptr && threadIdx.x == blockDim.x - 32) {                           \
      unsigned long long t_;                                                     \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                     \
      a.cta_times[(size_t)blockIdx.x * 4 + (slot)] = t_;                         \
    }                                                                            \
  } while (0)
#else
#define CE3_STAMP(role, tile, which) do { } while (0)
#define CE3_CTA_TIME(slot) do { } while (0)
#endif

// exp2 on the FMA/ALU pipes (degree-3 minimax of 2^f on [-0.5, 0.5], relative error 7.5e-5 - below the bf16 rounding
// of E): offloads a share of the exponentials from the MUFU pipe (16/clk/SM), see tools/micro/mufu_bench.cu.
__device__ __forceinline__ float exp2_poly(float x) {
  x = fmaxf(x, -126.f);
  const float t = x + 12582912.f;  // 1.5 * 2^23: the integer part lands in the low mantissa bits
  const float f = x - (t - 12582912.f);
  float p = fmaf(f, 0.05517166f, 0.24261112f);
  p = fmaf(p, f, 0.69326099f);
  p = fmaf(p, f, 0.99992807f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}

template <int DP, bool BIAS_X>
__global__ void __launch_bounds__(Cfg3<DP, BIAS_X>::THREADS, 1)
ce_bwd3_kernel(const __grid_constant__ TmapSet tmx, const __grid_constant__ TmapSet tmy,
               const __grid_constant__ CUtensorMap tme, const CeBwd3Args a) {
  using Cfg = Cfg3<DP, BIAS_X>;
  constexpr int BN = Cfg::BN, EG = Cfg::EG, NB = Cfg::NB, NE = Cfg::NE, CW = Cfg::CW;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sx = smem;
  uint8_t* sones = sx + Cfg::X_BYTES;
  uint8_t* sy = sones + Cfg::ONES_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sy + Cfg::STAGES * Cfg::Y_BYTES);
  uint64_t* x_full = bars;        // TMA landed the X tile
  uint64_t* x_empty = bars + 1;   // every S = X Y^T of the segment has completed (X in TMEM may be replaced)
  uint64_t* xt_full = bars + 2;   // X tile and its bias step staged in TMEM (16 epilogue warps)
  uint64_t* acc_full = bars + 3;
  uint64_t* acc_empty = bars + 4;
  uint64_t* sx_free = bars + 13;  // the epilogue warps are done using the X buffer as drain staging (one phase per segment)
  uint64_t* s_full = bars + 5;    // [NB]
  uint64_t* s_empty = bars + 7;   // [NB]: every epilogue warp holds its part of the score tile in registers
  uint64_t* e_full = bars + 9;    // [NE]: E of the whole tile is stored (all epilogue warps)
  uint64_t* y_full = bars + 14;
  uint64_t* y_empty = y_full + Cfg::STAGES;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(y_empty + Cfg::STAGES);

  // Roles by warp index: the epilogue warps come FIRST (0 .. 4 EG - 1), the TMA / UMMA-issue / TMEM-alloc warps LAST.
  // The scheduler of an SM sub-partition prefers its highest warp index; with the UMMA issuer at warp 1 its ~200
  // control instructions per tile queued behind three issue-hungry epilogue warps (200-400 idle cycles of the tensor
  // pipe per tile in the v3b timeline).
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int W_TMA = 4 * EG, W_MMA = 4 * EG + 1, W_ALLOC = 4 * EG + 2, W_MMA2 = 4 * EG + 3;
  CE3_CTA_TIME(0);
  if (warp == W_TMA && lane == 0) {
    for (int p = 0; p < tmx.n; ++p) tma_prefetch_desc(&tmx.m[p]);
    for (int p = 0; p < tmy.n; ++p) tma_prefetch_desc(&tmy.m[p]);
    if (!BIAS_X) tma_prefetch_desc(&tme);
  }
  if (warp == W_MMA && lane == 0) {
    mbar_init(x_full, 1);
    mbar_init(x_empty, 1);
    mbar_init(xt_full, 4 * EG);
    mbar_init(acc_full, 1);
    mbar_init(acc_empty, 4 * EG);
    mbar_init(sx_free, 4 * EG);
    for (int i = 0; i < NB; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&s_empty[i], 4 * EG);
    }
    for (int i = 0; i < NE; ++i) mbar_init(&e_full[i], 4 * EG);
    for (int i = 0; i < Cfg::STAGES; ++i) {
      mbar_init(&y_full[i], 1);
      mbar_init(&y_empty[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == W_ALLOC) tmem_alloc(tmem_holder, 512);
  if (BIAS_X) {
    // constant item-side operand of the bias step: K-major 128-byte-swizzled atom, columns k = 0..2 hold 1.0, the
    // rest 0 (only the 16-byte chunks 0 and 1 of a row are read by the K = 16 instruction)
    for (int i = threadIdx.x; i < 128 * 8; i += Cfg::THREADS) {
      const uint32_t row = i >> 3, ch = i & 7;
      const uint4 v = ch == 0 ? make_uint4(0x3F803F80u, 0x00003F80u, 0u, 0u) : make_uint4(0u, 0u, 0u, 0u);
      *reinterpret_cast<uint4*>(sones + sw128_offset(row, ch)) = v;
    }
    fence_proxy_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  CE3_CTA_TIME(1);

  if (warp == W_TMA) {
    if (lane == 0) {
      SegIter it(a.T, a.total, a.CT, a.CTr);
      int r, j0, j1, stage = 0;
      uint32_t phase = 0, xs = 0;
      while (it.next(r, j0, j1)) {
        mbar_wait(x_empty, (xs & 1) ^ 1);
        if (xs >= 2) mbar_wait(sx_free, xs & 1);  // the drain of segment xs - 2 staged through the X buffer
        mbar_arrive_expect_tx(x_full, Cfg::X_BYTES);
        int xrow;
        const CUtensorMap* mx = tmap_of(tmx, r * 128, xrow);
#pragma unroll
        for (int b = 0; b < Cfg::KBOX; ++b) tma_load_2d(sx + b * 16384, mx, x_full, b * 64, xrow);
        for (int j = j0; j < j1; ++j) {
          mbar_wait(&y_empty[stage], phase ^ 1);
          mbar_arrive_expect_tx(&y_full[stage], Cfg::Y_BYTES);
          uint8_t* dst = sy + stage * Cfg::Y_BYTES;
          int yrow;
          const CUtensorMap* my = tmap_of(tmy, j * BN, yrow);
#pragma unroll
          for (int b = 0; b < Cfg::KBOX; ++b) tma_load_2d(dst + b * (BN * 128), my, &y_full[stage], b * 64, yrow);
          if (!BIAS_X) tma_load_2d(dst + Cfg::Y_MAIN, &tme, &y_full[stage], 0, j * BN);
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
        }
        ++xs;
      }
    }
  } else if (warp == W_MMA) {
    // ---- issuer of the score tiles  S' = X Y^T (+ bias step) ----
    // One thread issuing BOTH contractions spends ~300 dependent instructions per tile on descriptors and barriers and
    // cannot keep the tensor pipe fed, so the two contractions have one issuing warp each, on different SM
    // sub-partitions.  Barriers are probed one batch ahead (non-blocking test_wait); in steady state every probe
    // succeeds and the thread never blocks.
    const uint32_t leader = elect_one();
    constexpr uint32_t idesc1 = make_idesc_bf16(128, BN, 0, 0);
    SegIter it(a.T, a.total, a.CT, a.CTr);
    int r, j0, j1;
    uint32_t t1 = 0, xs = 0;
    const uint64_t dy0 = make_smem_desc_sw128(smem_u32(sy), 0, 1024);
    const uint64_t dones = make_smem_desc_sw128(smem_u32(sones), 0, 1024);
    auto probe_next = [&](uint32_t t, uint32_t& y_ok, uint32_t& b_ok) {
      y_ok = mbar_probe(&y_full[t % Cfg::STAGES], (t / Cfg::STAGES) & 1);
      b_ok = mbar_probe(&s_empty[t % NB], ((t / NB) & 1) ^ 1);
    };
    uint32_t y_ok = 0, b_ok = 1;
    while (it.next(r, j0, j1)) {
      mbar_wait(xt_full, xs & 1);
      tc_fence_after();
      for (int j = j0; j < j1; ++j, ++t1) {
        const uint32_t buf = t1 % NB, stage = t1 % Cfg::STAGES;
        if (leader) CE3_STAMP(4, t1, 0);
        mbar_wait_probed(&y_full[stage], (t1 / Cfg::STAGES) & 1, y_ok);
        mbar_wait_probed(&s_empty[buf], ((t1 / NB) & 1) ^ 1, b_ok);
        tc_fence_after();
        if (leader) CE3_STAMP(0, t1, 0);
        probe_next(t1 + 1, y_ok, b_ok);
        const uint64_t dy = desc_advance(dy0, stage * Cfg::Y_BYTES);
#pragma unroll
        for (int k = 0; k < DP / 16; ++k)
          umma_bf16_ta_w(tmem_base + buf * BN, tmem_base + Cfg::X_COL + k * 8,
                         desc_advance(dy, (k >> 2) * (BN * 128) + (k & 3) * 32), idesc1, k > 0 ? 1u : 0u, leader);
        umma_bf16_ta_w(tmem_base + buf * BN, tmem_base + Cfg::XE_COL,
                       BIAS_X ? dones : desc_advance(dy, Cfg::Y_MAIN), idesc1, 1u, leader);
        if (leader) CE3_STAMP(4, t1, 1);
        umma_commit_w(&s_full[buf], leader);
        if (leader) CE3_STAMP(0, t1, 1);
      }
      umma_commit_w(x_empty, leader);
      ++xs;
    }
  } else if (warp == W_MMA2) {
    // ---- issuer of  acc += E Y  (E from its TMEM buffer, Y read MN-major); its completion frees the Y stage and E buffer ----
    const uint32_t leader = elect_one();
    constexpr uint32_t idesc2 = make_idesc_bf16(128, DP, 0, 1);
    SegIter it(a.T, a.total, a.CT, a.CTr);
    int r, j0, j1;
    uint32_t t2 = 0, xs = 0;
    const uint64_t dyt0 = make_smem_desc_sw128(smem_u32(sy), BN * 128, 1024);
    uint32_t e_ok = 0;
    while (it.next(r, j0, j1)) {
      for (int j = j0; j < j1; ++j, ++t2) {
        const uint32_t buf = t2 % NE, stage = t2 % Cfg::STAGES;
        if (leader) CE3_STAMP(1, t2, 0);
        mbar_wait_probed(&e_full[buf], (t2 / NE) & 1, e_ok);
        if (j == j0) mbar_wait(acc_empty, (xs & 1) ^ 1);
        tc_fence_after();
        if (leader) CE3_STAMP(5, t2, 0);
        e_ok = mbar_probe(&e_full[(t2 + 1) % NE], ((t2 + 1) / NE) & 1);
        const uint64_t dyt = desc_advance(dyt0, stage * Cfg::Y_BYTES);
#pragma unroll
        for (int k = 0; k < BN / 16; ++k)  // K = 16 rows of the Y tile per instruction
          umma_bf16_ta_w(tmem_base + Cfg::ACC_COL, tmem_base + Cfg::E_COL + buf * (BN / 2) + k * 8,
                         desc_advance(dyt, k * 2048), idesc2, (j > j0 || k > 0) ? 1u : 0u, leader);
        if (leader) CE3_STAMP(5, t2, 1);
        umma_commit_w(&y_empty[stage], leader);
        if (leader) CE3_STAMP(1, t2, 1);
      }
      umma_commit_w(acc_full, leader);
      ++xs;
    }
  } else if (warp < 4 * EG) {
    const int e = warp >> 2;
    const int q = warp & 3;
    const uint32_t prow = q * 32 + lane;
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    const float gsv = (a.g_scale != nullptr ? __ldg(a.g_scale) : 1.f) * (a.g_scale2 != nullptr ? __ldg(a.g_scale2) : 1.f);
    const float gabs = fabsf(gsv);
    const float c0 = log2f(gabs);  // -inf for a zero scale: every exponential becomes 0
    const uint32_t sgn_gs = gsv < 0.f ? 0x80008000u : 0u;
    SegIter it(a.T, a.total, a.CT, a.CTr);
    int r, j0, j1;
    uint32_t t = 0, xs = 0;
    // X tile (and the bias step of its rows) from shared memory into tensor memory
    auto stage_x = [&](int rr, uint32_t seg) {
      mbar_wait(x_full, seg & 1);
#pragma unroll
      for (int part = 0; part < Cfg::XP; ++part) {
        if (part % EG != e) continue;
        uint32_t xr[16];
        const uint8_t* atom = sx + ((part * 32) >> 6) * 16384;
        const uint32_t ch0 = ((part * 32) & 63) >> 3;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const uint4 u = *reinterpret_cast<const uint4*>(atom + sw128_offset(prow, ch0 + c));
          xr[4 * c] = u.x; xr[4 * c + 1] = u.y; xr[4 * c + 2] = u.z; xr[4 * c + 3] = u.w;
        }
        tmem_st16(lane_base + Cfg::X_COL + part * 16, xr);
      }
      if (e == EG - 1) {
        uint32_t xe[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) xe[i] = 0u;
        if (BIAS_X) {
          const long long row2 = (long long)rr * 128 + prow;
          float b = -30000.f;  // rows past the end and rows with g = 0 produce E = 0
          if (row2 < a.XR) {
            const float gi = fabsf(__ldg(a.g + row2));
            if (gi > 0.f) b = logf(gi) - __ldg(a.lse + row2);
          }
          const bf16 hi = __float2bfloat16(b);
          const float r1 = b - __bfloat162float(hi);
          const bf16 mid = __float2bfloat16(r1);
          const bf16 lo = __float2bfloat16(r1 - __bfloat162float(mid));
          xe[0] = (uint32_t)__bfloat16_as_ushort(hi) | ((uint32_t)__bfloat16_as_ushort(mid) << 16);
          xe[1] = (uint32_t)__bfloat16_as_ushort(lo);
        } else {
          xe[0] = 0x3F803F80u;
          xe[1] = 0x00003F80u;
        }
        tmem_st8(lane_base + Cfg::XE_COL, xe);
      }
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(xt_full);
    };
    bool have = it.next(r, j0, j1);
    if (have) stage_x(r, 0);
    while (have) {
      const long long row = (long long)r * 128 + prow;
      const long long tgt = row + a.diag_shift;
      const bool haspos = row < a.XR && tgt >= 0 && tgt < a.YR;
      const int jd = haspos ? (int)(tgt / BN) : -1;                  // tile that holds this row's positive
      const bool mine = haspos && (int)((tgt % BN) / CW) == e;       // ... and it is in this group's CW columns
      const int od = haspos ? (int)((tgt % BN) % CW) : 0;
      float sub = 0.f;
      uint32_t sgn_row = sgn_gs;
      if (BIAS_X) {
        const float gi = row < a.XR ? __ldg(a.g + row) : 0.f;
        sub = fabsf(gi) * gabs;
        if (gi < 0.f) sgn_row ^= 0x80008000u;
      } else if (mine) {
        sub = fabsf(__ldg(a.g + tgt)) * gabs;
      }
      // (Issuing the tcgen05.ld of tile t + 1 before the exponentials of tile t - two register sets - was measured
      // SLOWER: 1560 instead of 1050 cycles per tile, with spills at 128 registers.)
      for (int j = j0; j < j1; ++j, ++t) {
        const uint32_t buf = t % NB;
        uint32_t cmask = 0u;  // sign bits of g of this chunk's CW users (columns)
        if (!BIAS_X) cmask = CW == 32 ? __ldg(a.signmask + j * (BN / 32) + e)
                                      : (__ldg(a.signmask + j * (BN / 32) + (e >> 1)) >> ((e & 1) * 16)) & 0xffffu;
        mbar_wait(&s_full[buf], (t / NB) & 1);
        tc_fence_after();
        if (lane == 0 && e == 0) CE3_STAMP(q < 2 ? 2 + q : 4 + q, t, 0);  // rows 2, 3, 6, 7 of the trace: the four lane quarters
        float v[CW];
        if (CW == 32) tmem_ld32(lane_base + buf * BN + e * CW, v);
        else tmem_ld16(lane_base + buf * BN + e * CW, v);
        tmem_wait_ld();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_empty[buf]);  // the score buffer goes back to the issuer of S = X Y^T
#ifdef TT_CE_POLY
#pragma unroll
        for (int i = 0; i < CW; ++i) {
          const float x = fmaf(v[i], LOG2E, c0);
          v[i] = (i & 3) == 3 ? exp2_poly(x) : ex2f(x);
        }
#else
#pragma unroll
        for (int i = 0; i < CW; ++i) v[i] = ex2f(fmaf(v[i], LOG2E, c0));
#endif
        if (mine && j == jd) {
#pragma unroll
          for (int i = 0; i < CW; ++i) v[i] -= (i == od) ? sub : 0.f;
        }
        uint32_t p[CW / 2];
#pragma unroll
        for (int i = 0; i < CW / 2; ++i) p[i] = pack_bf16x2(v[2 * i], v[2 * i + 1]) ^ sgn_row;
        if (!BIAS_X && cmask != 0u) {
#pragma unroll
          for (int i = 0; i < CW / 2; ++i)
            p[i] ^= (((cmask >> (2 * i)) & 1u) << 15) | (((cmask >> (2 * i + 1)) & 1u) << 31);
        }
        // E buffer t % NE is free once acc += E(t - NE) Y(t - NE) has completed = the event y_empty tracks for that tile
        if (t >= (uint32_t)NE) {
          mbar_wait(&y_empty[(t - NE) % Cfg::STAGES], ((t - NE) / Cfg::STAGES) & 1);
          tc_fence_after();
        }
        if (CW == 32) tmem_st16(lane_base + Cfg::E_COL + (t % NE) * (BN / 2) + e * (CW / 2), p);
        else tmem_st8(lane_base + Cfg::E_COL + (t % NE) * (BN / 2) + e * (CW / 2), p);
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&e_full[t % NE]);
        if (lane == 0 && e == 0) CE3_STAMP(q < 2 ? 2 + q : 4 + q, t, 1);
      }
      // the next segment's X tile goes into tensor memory first (its first score MMAs then overlap the drain)
      int r2 = 0, j02 = 0, j12 = 0;
      const bool have2 = it.next(r2, j02, j12);
      if (q == 0 && lane == 0 && e == 0 && have2) CE3_STAMP(2, 60, 0);
      if (have2) stage_x(r2, xs + 1);
      if (q == 0 && lane == 0 && e == 0 && have2) CE3_STAMP(2, 60, 1);
      // segment accumulator -> partial slot
      mbar_wait(acc_full, xs & 1);
      tc_fence_after();
      if (q == 0 && lane == 0 && e == 0 && have2) CE3_STAMP(2, 61, 0);
      {
        // A thread holds 32 consecutive columns of ONE row; written straight to global memory every store instruction
        // would touch 32 different 128-byte lines (the first drafts did: 4000 cycles per drain, and the flood of partial
        // sectors delayed the TMA loads of the next segment).  Each warp transposes through a private slice of the X
        // staging buffer (idle after stage_x): CHF floats of its 32 rows per round, written with an XOR swizzle,
        // read back so that 4 (2) consecutive lanes cover 64 (32) contiguous bytes of a row.  (Four 32-byte stores per
        // thread straight from registers - full sectors, no transpose - measured 82.9 vs 80.5 us for both passes.)
        constexpr int CHF = DP == 64 ? 8 : 16;          // floats of a row per round
        constexpr int CPR = CHF / 4;                    // 16-byte chunks per row and round
        constexpr int RPI = 32 / CPR;                   // rows per read-back instruction
        static_assert(4 * (EG < Cfg::XP ? EG : Cfg::XP) * 32 * CHF * 4 <= Cfg::X_BYTES, "drain staging exceeds the X buffer");
        const int slot = (int)(blockIdx.x - ((long long)r * a.CT) / a.T);
        float* tile = a.partial + (long long)slot * a.slot_stride + ((long long)r * 128 + q * 32) * DP;  // row q*32 of the tile
        uint8_t* wbuf = sx + warp * (32 * CHF * 4);
        const uint32_t sw = CHF == 16 ? ((lane >> 1) & 3) : ((lane >> 2) & 1);
        int last_part = -1;
#pragma unroll
        for (int part = 0; part < Cfg::XP; ++part)
          if (part % EG == e) last_part = part;
        if (last_part < 0) {  // this warp has no accumulator columns to drain (d = 64: two of the four groups)
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(acc_empty);
        }
#pragma unroll
        for (int part = 0; part < Cfg::XP; ++part) {
          if (part % EG != e) continue;
          float w[32];
          tmem_ld32(lane_base + Cfg::ACC_COL + part * 32, w);
          tmem_wait_ld();
          if (part == last_part) {  // the accumulator columns of this warp are in registers: hand them back
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(acc_empty);
          }
          if (q == 0 && lane == 0 && e == 0 && have2) CE3_STAMP(3, part == last_part ? 61 : 60, 0);
#pragma unroll
          for (int round = 0; round < 32 / CHF; ++round) {
#pragma unroll
            for (int c = 0; c < CPR; ++c)
              *reinterpret_cast<float4*>(wbuf + lane * (CHF * 4) + ((c ^ sw) << 4)) =
                  make_float4(w[round * CHF + 4 * c], w[round * CHF + 4 * c + 1], w[round * CHF + 4 * c + 2], w[round * CHF + 4 * c + 3]);
            __syncwarp();
#pragma unroll
            for (int k = 0; k < 32 / RPI; ++k) {
              const int rr = k * RPI + lane / CPR, c = lane % CPR;
              const uint32_t swr = CHF == 16 ? ((rr >> 1) & 3) : ((rr >> 2) & 1);
              const float4 val = *reinterpret_cast<const float4*>(wbuf + rr * (CHF * 4) + ((c ^ swr) << 4));
              *reinterpret_cast<float4*>(tile + (long long)rr * DP + part * 32 + round * CHF + 4 * c) = val;
            }
            __syncwarp();
          }
          if (q == 0 && lane == 0 && e == 0 && have2) CE3_STAMP(3, part == last_part ? 61 : 60, 1);
        }
        if (lane == 0) mbar_arrive(sx_free);
      }
      if (q == 0 && lane == 0 && e == 0 && have2) CE3_STAMP(2, 61, 1);
      r = r2; j0 = j02; j1 = j12; have = have2;
      ++xs;
    }
  }

  tc_fence_before();
  __syncthreads();
  CE3_CTA_TIME(2);
  if (warp == W_ALLOC) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
  CE3_CTA_TIME(3);
}

// b = ln|g| - lse of every user as a bf16 triple in columns 0..2 of a [rows, 64] bf16 matrix (the Y-side operand of the
// dV pass's bias step, fetched by TMA beside the U tile), and the sign bits of g, 32 users per word.
__global__ void ce_bwd3_prep_kernel(int B, int rows_pad, const float* __restrict__ g, const float* __restrict__ lse,
                                    uint4* __restrict__ uext, uint32_t* __restrict__ signmask) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  float gi = 0.f, b = -30000.f;
  if (row < B) {
    gi = g[row];
    const float ga = fabsf(gi);
    if (ga > 0.f) b = logf(ga) - lse[row];
  }
  const unsigned neg = __ballot_sync(0xffffffffu, gi < 0.f);
  if (row >= rows_pad) return;
  if ((threadIdx.x & 31) == 0) signmask[row >> 5] = neg;
  const bf16 hi = __float2bfloat16(b);
  const float r1 = b - __bfloat162float(hi);
  const bf16 mid = __float2bfloat16(r1);
  const bf16 lo = __float2bfloat16(r1 - __bfloat162float(mid));
  uint4* dst = uext + (size_t)row * 8;
  dst[0] = make_uint4((uint32_t)__bfloat16_as_ushort(hi) | ((uint32_t)__bfloat16_as_ushort(mid) << 16),
                      (uint32_t)__bfloat16_as_ushort(lo), 0u, 0u);
#pragma unroll
  for (int c = 1; c < 8; ++c) dst[c] = make_uint4(0u, 0u, 0u, 0u);  // (only chunks 0 and 1 are read by the K = 16 instruction)
}

template <int DP, bool BIAS_X>
int launch3(const TmapSet& tx, const TmapSet& ty, const CUtensorMap& te, const CeBwd3Args& a, int grid, cudaStream_t st) {
  using Cfg = Cfg3<DP, BIAS_X>;
  static bool configured = false;
  if (!configured) {
    TT_CUDA(cudaFuncSetAttribute(ce_bwd3_kernel<DP, BIAS_X>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    configured = true;
  }
  KernelSpan span(BIAS_X ? "ce_bwd3_kernel_dU" : "ce_bwd3_kernel_dV", st);
  ce_bwd3_kernel<DP, BIAS_X><<<grid, Cfg::THREADS, Cfg::SMEM_BYTES, st>>>(tx, ty, te, a);
  TT_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

}  // namespace

int ce_bwd3_tile_cols(int DP) { return DP == 64 ? Tile3<64>::BN : Tile3<128>::BN; }

size_t ce_bwd3_ext_bytes(long long users) {
  const long long pad = (users + 127) / 128 * 128;
  return (size_t)pad * 128 + (size_t)pad / 32 * 4 + 256;
}

int ce_bwd3_prep(long long users, const float* g, const float* lse, void* ext, cudaStream_t st) {
  const long long pad = (users + 127) / 128 * 128;
  uint4* uext = reinterpret_cast<uint4*>(ext);
  uint32_t* signmask = reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(ext) + (size_t)pad * 128);
  KernelSpan span("ce_bwd3_prep_kernel", st);
  ce_bwd3_prep_kernel<<<(unsigned)(pad / 256 + (pad % 256 != 0)), 256, 0, st>>>((int)users, (int)pad, g, lse, uext, signmask);
  TT_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

int launch_ce_bwd3(int DP, bool bias_x, const TmapSet& tx, const TmapSet& ty, long long users, const void* ext,
                   CeBwd3Args a, int grid, cudaStream_t st) {
  const long long pad = (users + 127) / 128 * 128;
  a.signmask = reinterpret_cast<const uint32_t*>(reinterpret_cast<const uint8_t*>(ext) + (size_t)pad * 128);
  CUtensorMap te;
  int rc = make_tmap_bf16(&te, ext, 64, (uint64_t)pad, 64, 64, (uint32_t)ce_bwd3_tile_cols(DP));
  if (rc) return rc;
  if (DP == 64) return bias_x ? launch3<64, true>(tx, ty, te, a, grid, st) : launch3<64, false>(tx, ty, te, a, grid, st);
  return bias_x ? launch3<128, true>(tx, ty, te, a, grid, st) : launch3<128, false>(tx, ty, te, a, grid, st);
}

}  // namespace tt
