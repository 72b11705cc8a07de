// In-batch cross-entropy backward, v2: the E = g (softmax - onehot) tile and the X tile live in TENSOR MEMORY.
//
// Math and scheduling (see ce.cu):  acc[128, d] = sum_j E_j Y_j,  E_j = f(X Y_j^T), run
// once for dU (X=U, Y=V, row statistics) and once for dV (X=V, Y=U, column statistics).  What changed is
// where the UMMA A operands come from.  With both operands in shared memory a 128x128x16 UMMA reads 8 KB per
// 64 tensor-pipe cycles = the full 128 B/clk of shared-memory bandwidth, and the measured issue interval was
// ~100 cycles per instruction instead of 64 (profiles/r01_ce_bwd_timeline.txt).  Here
//   * E is written by the epilogue warps with tcgen05.st (bf16 pairs, lane = row, column = K/2) into its own
//     TMEM columns and consumed as the TMEM A operand of  acc += E Y  - no swizzled st.shared, no proxy fence;
//   * the X tile of the segment is copied once into TMEM and is the A operand of every S = X Y^T,
// so each UMMA reads only its 4 KB B slice from shared memory.
//
// TMEM columns (DP = 128): S0 [0,128) S1 [128,256) | acc [256,384) | E [384,448) | X [448,512).
// Warp roles as before: 0 TMA, 1 UMMA issue (warp-uniform, elected lane), 2 TMEM alloc, 4-11 epilogue:
// group e = column half of every score tile, q = TMEM lane quarter.
#include <stdlib.h>

#include "ce_common.cuh"

namespace tt {

namespace {

template <int DP>
struct Cfg2 {
  static constexpr int BN = DP == 256 ? 64 : 128;
  static constexpr int NS = 2;                   // score-tile buffers
  // epilogue column groups (4 warps each).  4 groups (32 columns per thread) were measured SLOWER here (113 vs 97 us
  // at B = N = 8192, d = 128): unlike the forward, every group also waits for its slice of E to be consumed, stores
  // it and signals the UMMA warp, and that per-warp handshake does not shrink with the column count.
  static constexpr int EG = 2;
  static constexpr int THREADS = 128 + EG * 128;
  static constexpr int XG = (DP / 32 < EG) ? DP / 32 : EG;  // groups that copy X / drain the accumulator (>= 32 columns each)
  static constexpr bool XT = DP <= 128;          // X tile resident in TMEM
  static constexpr int KBOX = DP / 64;
  static constexpr int X_BYTES = 128 * DP * 2;
  static constexpr int Y_BYTES = BN * DP * 2;
  static constexpr int STAGES = DP == 64 ? 8 : (DP == 128 ? 6 : 4);
  static constexpr int COLSTAT_BYTES = BN * 8;  // (g, lse) per column of the current tile
  static constexpr int SMEM_BYTES = X_BYTES + STAGES * Y_BYTES + COLSTAT_BYTES + 1024 + 256;
  static constexpr int ACC_COL = NS * BN;
  static constexpr int E_COL = ACC_COL + DP;
  static constexpr int X_COL = E_COL + BN / 2;
  static constexpr int TMEM_USED = X_COL + (XT ? DP / 2 : 0);
  static_assert(TMEM_USED <= 512, "TMEM budget");
  static_assert(SMEM_BYTES <= 232448, "shared memory budget");
};

template <int DP, bool COLSTATS>
__global__ void __launch_bounds__(Cfg2<DP>::THREADS, 1)
ce_bwd2_kernel(const __grid_constant__ TmapSet tmx, const __grid_constant__ TmapSet tmy, const CeBwdArgs a) {
  using Cfg = Cfg2<DP>;
  constexpr int BN = Cfg::BN, NS = Cfg::NS, EG = Cfg::EG, XG = Cfg::XG;
  constexpr bool XT = Cfg::XT;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sx = smem;
  uint8_t* sy = sx + Cfg::X_BYTES;
  float2* scol = reinterpret_cast<float2*>(sy + Cfg::STAGES * Cfg::Y_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(scol) + Cfg::COLSTAT_BYTES);
  uint64_t* x_full = bars;        // TMA landed the X tile
  uint64_t* x_empty = bars + 1;   // every S = X Y^T of the segment has completed
  uint64_t* xt_full = bars + 2;   // X tile copied into TMEM (8 epilogue warps)
  uint64_t* acc_full = bars + 3;
  uint64_t* acc_empty = bars + 4;
  uint64_t* e_full = bars + 5;    // [EG] per column group of E
  uint64_t* e_empty = bars + 9;   // [EG]
  uint64_t* s_full = bars + 13;   // [NS]
  uint64_t* s_empty = s_full + NS;
  uint64_t* y_full = s_empty + NS;
  uint64_t* y_empty = y_full + Cfg::STAGES;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(y_empty + Cfg::STAGES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
// TT_CE_BWD_LEAN (candidate for the next round, default off until measured on a B200): the epilogue's bookkeeping
// is cut the way the forward's was - bring-up stamps compiled out, 32-bit column arithmetic, the positive handled on
// the one tile that holds it - ncu counted ~150 bookkeeping instructions around 224 of math per warp and tile.
#if defined(TT_CE_BWD_LEAN) && !defined(TT_CE_BRINGUP)
#define CTA_TIME(slot) do { } while (0)
#define CE_STAMP(role, tile, which) do { } while (0)
#else
#define TT_CE_BWD_HOOKS 1
#define CTA_TIME(slot)                                                          \
  do {                                                                          \
    if (a.cta_times != nullptr && threadIdx.x == 96) {                          \
      unsigned long long t_;                                                    \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                    \
      a.cta_times[(size_t)blockIdx.x * 4 + (slot)] = t_;                        \
    }                                                                           \
  } while (0)
  CTA_TIME(0);
#define CE_STAMP(role, tile, which)                                                           \
  do {                                                                                        \
    if (a.trace != nullptr && blockIdx.x == 0 && (tile) < 64) a.trace[((role) * 64 + (tile)) * 2 + (which)] = clock64(); \
  } while (0)
#endif
  if (warp == 0 && lane == 0) {
    for (int p = 0; p < tmx.n; ++p) tma_prefetch_desc(&tmx.m[p]);
    for (int p = 0; p < tmy.n; ++p) tma_prefetch_desc(&tmy.m[p]);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(x_full, 1);
    mbar_init(x_empty, 1);
    mbar_init(xt_full, 4 * EG);
    mbar_init(acc_full, 1);
    mbar_init(acc_empty, 4 * EG);
    for (int i = 0; i < EG; ++i) {
      mbar_init(&e_full[i], 4);
      mbar_init(&e_empty[i], 1);
    }
    for (int i = 0; i < NS; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&s_empty[i], 4 * EG);
    }
    for (int i = 0; i < Cfg::STAGES; ++i) {
      mbar_init(&y_full[i], 1);
      mbar_init(&y_empty[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_holder, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  CTA_TIME(1);

  if (warp == 0) {
    if (lane == 0) {
      SegIter it(a.T, a.total, a.CT);
      int r, j0, j1, stage = 0;
      uint32_t phase = 0, xs = 0;
      while (it.next(r, j0, j1)) {
        mbar_wait(x_empty, (xs & 1) ^ 1);
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
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
        }
        ++xs;
      }
    }
  } else if (warp == 1) {
    const uint32_t leader = elect_one();
    constexpr uint32_t idesc1 = make_idesc_bf16(128, BN, 0, 0);  // S = X Y^T
    constexpr uint32_t idesc2 = make_idesc_bf16(128, DP, 0, 1);  // acc += E Y   (Y read MN-major)
    SegIter it(a.T, a.total, a.CT);
    int r, j0, j1;
    int stage1 = 0, stage2 = 0;
    uint32_t phase1 = 0;
    uint32_t t1 = 0, t2 = 0, xs = 0;
    const uint64_t dx0 = make_smem_desc_sw128(smem_u32(sx), 0, 1024);
    const uint64_t dy0 = make_smem_desc_sw128(smem_u32(sy), 0, 1024);
    const uint64_t dyt0 = make_smem_desc_sw128(smem_u32(sy), BN * 128, 1024);
    auto mma1 = [&]() {
      const uint32_t buf = t1 % NS, use = t1 / NS;
      mbar_wait(&y_full[stage1], phase1);
      if (leader) CE_STAMP(0, t1, 0);
      mbar_wait(&s_empty[buf], (use & 1) ^ 1);
      tc_fence_after();
      if (leader) CE_STAMP(0, t1, 1);
      const uint64_t dy = desc_advance(dy0, stage1 * Cfg::Y_BYTES);
#pragma unroll
      for (int k = 0; k < DP / 16; ++k) {
        const uint64_t db = desc_advance(dy, (k >> 2) * (BN * 128) + (k & 3) * 32);
        if (XT)
          umma_bf16_ta_w(tmem_base + buf * BN, tmem_base + Cfg::X_COL + k * 8, db, idesc1, k > 0 ? 1u : 0u, leader);
        else
          umma_bf16_w(tmem_base + buf * BN, desc_advance(dx0, (k >> 2) * 16384 + (k & 3) * 32), db, idesc1,
                      k > 0 ? 1u : 0u, leader);
      }
      umma_commit_w(&s_full[buf], leader);
      if (++stage1 == Cfg::STAGES) { stage1 = 0; phase1 ^= 1; }
      ++t1;
    };
    auto mma2 = [&](bool first) {
      constexpr int KH = BN / (16 * EG);  // K = 16 steps per column group of E
      const uint64_t dyt = desc_advance(dyt0, stage2 * Cfg::Y_BYTES);
#pragma unroll
      for (int h = 0; h < EG; ++h) {
        if (h == 0 && leader) CE_STAMP(1, t2, 0);
        mbar_wait(&e_full[h], t2 & 1);
        tc_fence_after();
        if (h == EG - 1 && leader) CE_STAMP(1, t2, 1);
#pragma unroll
        for (int kk = 0; kk < KH; ++kk) {
          const int k = h * KH + kk;
          umma_bf16_ta_w(tmem_base + Cfg::ACC_COL, tmem_base + Cfg::E_COL + k * 8, desc_advance(dyt, k * 2048), idesc2,
                         (!first || k > 0) ? 1u : 0u, leader);
        }
        umma_commit_w(&e_empty[h], leader);
      }
      umma_commit_w(&y_empty[stage2], leader);
      if (++stage2 == Cfg::STAGES) stage2 = 0;
      ++t2;
    };
    while (it.next(r, j0, j1)) {
      const int n = j1 - j0;
      int issued = 0;
      if (XT) mbar_wait(xt_full, xs & 1);
      else mbar_wait(x_full, xs & 1);
      tc_fence_after();
      for (int i = 0; i < n; ++i) {
        while (issued < n && issued <= i + (NS - 1)) {
          mma1();
          if (++issued == n) umma_commit_w(x_empty, leader);
        }
        if (i == 0) {
          mbar_wait(acc_empty, (xs & 1) ^ 1);
          tc_fence_after();
        }
        mma2(i == 0);
      }
      umma_commit_w(acc_full, leader);
      ++xs;
    }
  } else if (warp >= 4) {
    const int e = (warp - 4) >> 2;
    const int q = warp & 3;
    const int wg_tid = threadIdx.x - 128 - e * 128;
    constexpr int CH = BN / (32 * EG);  // 32-column chunks per group and tile
    SegIter it(a.T, a.total, a.CT);
    int r, j0, j1;
    uint32_t t = 0, xs = 0;
    const uint32_t prow = q * 32 + lane;
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    const float gs = (a.g_scale != nullptr ? __ldg(a.g_scale) : 1.f) * (a.g_scale2 != nullptr ? __ldg(a.g_scale2) : 1.f);
    while (it.next(r, j0, j1)) {
      const long long row = (long long)r * 128 + prow;
      const bool valid = row < a.XR;
      if (XT) {
        // copy this thread's part of its X row (DP / XG elements; groups >= XG have none) from (swizzled) shared
        // memory into tensor memory
        mbar_wait(x_full, xs & 1);
        if (e < XG) {
          constexpr int PART = DP / XG;   // elements per thread, a multiple of 32
          constexpr int NCH = PART / 8;   // 16-byte chunks
          uint32_t xr[NCH * 4];
          const uint8_t* atom = sx + ((e * PART) >> 6) * 16384;
          const uint32_t ch0 = ((e * PART) & 63) >> 3;
#pragma unroll
          for (int c = 0; c < NCH; ++c) {
            const uint4 u = *reinterpret_cast<const uint4*>(atom + sw128_offset(prow, ch0 + c));
            xr[4 * c] = u.x; xr[4 * c + 1] = u.y; xr[4 * c + 2] = u.z; xr[4 * c + 3] = u.w;
          }
#pragma unroll
          for (int c = 0; c < NCH / 4; ++c) tmem_st16(lane_base + Cfg::X_COL + e * (PART / 2) + c * 16, xr + 16 * c);
          tmem_wait_st();
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(xt_full);
      }
      float rs = 1.f, rl = 0.f;
      if (!COLSTATS) {
        rs = valid ? a.g[row] * gs : 0.f;
        rl = valid ? a.lse[row] * LOG2E : 0.f;
      }
      const long long tgt = row + a.diag_shift;
#ifdef TT_CE_BWD_LEAN
      // row pass: |g| folded into the exponent (g exp2(x) = sign(g) exp2(x + log2|g|); g = 0 gives exp2(-inf) = 0), the
      // sign goes onto the packed bf16 pairs - one multiply per element less
      const float rl2 = rl - log2f(fabsf(rs));
      const float rs_abs = fabsf(rs);
      const uint32_t sgn = (!COLSTATS && rs < 0.f) ? 0x80008000u : 0u;
      // tile / 32-column chunk / column of this row's positive (-1: none among the YR columns)
      const int tgt_i = (valid && tgt >= 0 && tgt < a.YR) ? (int)tgt : -1;
      const int jd = tgt_i >= 0 ? tgt_i / BN : -1, cd = tgt_i >= 0 ? (tgt_i % BN) >> 5 : -1, od = tgt_i & 31;
#endif
      // column statistics (g, lse) of the tile's columns live in shared memory: the loads for tile j + 1 are issued
      // before tile j is transformed and are written afterwards, so their L2 round trip hides behind the tile instead
      // of stalling every tile.
      auto colstat_load = [&](int jj) -> float2 {
        const long long col = (long long)jj * BN + e * (BN / EG) + wg_tid;
        const bool cv = wg_tid < BN / EG && col < a.YR;
        return make_float2(cv ? a.g[col] * gs : 0.f, cv ? a.lse[col] * LOG2E : 0.f);
      };
      if (COLSTATS) {
        const float2 cs0 = colstat_load(j0);
        if (wg_tid < BN / EG) scol[e * (BN / EG) + wg_tid] = cs0;
        asm volatile("bar.sync %0, 128;" ::"r"(1 + e) : "memory");
      }
      for (int j = j0; j < j1; ++j, ++t) {
        const uint32_t buf = t % NS;
        const float2* sc = scol;
        float2 cs_next = make_float2(0.f, 0.f);
        if (COLSTATS && j + 1 < j1) cs_next = colstat_load(j + 1);
        mbar_wait(&s_full[buf], (t / NS) & 1);
        tc_fence_after();
        if (q == 0 && lane == 0 && e < 2) CE_STAMP(2 + e, t, 0);
        // E = g (exp(S - lse) - [positive]) for 32 columns, packed to bf16 pairs
        auto transform = [&](float* v, int c, uint32_t* out) {
#ifndef TT_CE_BWD_LEAN
          const long long n0 = (long long)j * BN + c * 32;
#endif
          if (!COLSTATS) {
#ifdef TT_CE_BWD_LEAN
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = ex2f(fmaf(v[i], LOG2E, -rl2));
#else
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = rs * ex2f(fmaf(v[i], LOG2E, -rl));
#endif
#ifndef TT_CE_BWD_LEAN
            if (tgt >= n0 && tgt < n0 + 32) {
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (n0 + i == tgt) v[i] -= rs;
            }
#endif
          } else {
#pragma unroll
            for (int i = 0; i < 32; i += 2) {  // (g, lse) of two columns per 16-byte broadcast load
              const float4 cc4 = *reinterpret_cast<const float4*>(&sc[c * 32 + i]);
              v[i] = cc4.x * ex2f(fmaf(v[i], LOG2E, -cc4.y));
              v[i + 1] = cc4.z * ex2f(fmaf(v[i + 1], LOG2E, -cc4.w));
            }
#ifndef TT_CE_BWD_LEAN
            if (tgt >= n0 && tgt < n0 + 32) {
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (n0 + i == tgt) v[i] -= sc[c * 32 + i].x;
            }
#endif
          }
#ifdef TT_CE_BWD_LEAN
          if (j == jd && c == cd) {  // the one chunk of the one tile that holds this row's positive
            const float sub = COLSTATS ? sc[c * 32 + od].x : rs_abs;
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] -= (i == od) ? sub : 0.f;
          }
#endif
#ifdef TT_CE_BWD_LEAN
#pragma unroll
          for (int i = 0; i < 16; ++i) out[i] = pack_bf16x2(v[2 * i], v[2 * i + 1]) ^ sgn;
#else
#pragma unroll
          for (int i = 0; i < 16; ++i) out[i] = pack_bf16x2(v[2 * i], v[2 * i + 1]);
#endif
        };
        const uint32_t s_addr = lane_base + buf * BN + e * (CH * 32);
        const uint32_t e_addr = lane_base + Cfg::E_COL + e * (CH * 16);
        float v0[32];
        uint32_t p0[16];
        tmem_ld32(s_addr, v0);
        tmem_wait_ld();
        if (CH == 2) {
          float v1[32];
          uint32_t p1[16];
          tmem_ld32(s_addr + 32, v1);  // in flight while the first chunk is transformed
          transform(v0, e * CH, p0);
          mbar_wait(&e_empty[e], (t & 1) ^ 1);  // acc += E(t-1) Y(t-1) has consumed this half of E
          tc_fence_after();
          tmem_st16(e_addr, p0);
          tmem_wait_ld();
          transform(v1, e * CH + 1, p1);
          tmem_st16(e_addr + 16, p1);
        } else {
          transform(v0, e * CH, p0);
          mbar_wait(&e_empty[e], (t & 1) ^ 1);
          tc_fence_after();
          tmem_st16(e_addr, p0);
        }
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&s_empty[buf]);
          mbar_arrive(&e_full[e]);
        }
        if (COLSTATS && j + 1 < j1) {
          asm volatile("bar.sync %0, 128;" ::"r"(1 + e) : "memory");  // every thread of the group is done with tile j's
          if (wg_tid < BN / EG) scol[e * (BN / EG) + wg_tid] = cs_next;
          asm volatile("bar.sync %0, 128;" ::"r"(1 + e) : "memory");
        }
        if (q == 0 && lane == 0 && e < 2) CE_STAMP(2 + e, t, 1);
      }
      // segment accumulator -> partial slot (each group drains half of the columns)
      mbar_wait(acc_full, xs & 1);
      tc_fence_after();
      if (e < XG) {
        const int slot = (int)(blockIdx.x - ((long long)r * a.CT) / a.T);
        float* dst = a.partial + (long long)slot * a.slot_stride + row * DP;
#pragma unroll 1
        for (int c = 0; c < DP / (32 * XG); ++c) {
          const int col = e * (DP / XG) + c * 32;
          float v[32];
          tmem_ld32(lane_base + Cfg::ACC_COL + col, v);
          tmem_wait_ld();
#pragma unroll
          for (int i = 0; i < 8; ++i)
            *reinterpret_cast<float4*>(dst + col + 4 * i) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty);
      ++xs;
    }
  }
#undef CE_STAMP

  tc_fence_before();
  __syncthreads();
  CTA_TIME(2);
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
  CTA_TIME(3);
#undef CTA_TIME
}

template <int DP, bool COLSTATS>
int launch2(const TmapSet& tx, const TmapSet& ty, const CeBwdArgs& a, int grid, cudaStream_t st) {
  using Cfg = Cfg2<DP>;
  static bool configured = false;
  if (!configured) {
    TT_CUDA(cudaFuncSetAttribute(ce_bwd2_kernel<DP, COLSTATS>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    configured = true;
  }
  KernelSpan span(COLSTATS ? "ce_bwd2_kernel_dV" : "ce_bwd2_kernel_dU", st);
  ce_bwd2_kernel<DP, COLSTATS><<<grid, Cfg::THREADS, Cfg::SMEM_BYTES, st>>>(tx, ty, a);
  TT_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

}  // namespace

int launch_ce_bwd2(int DP, bool colstats, const TmapSet& tx, const TmapSet& ty, const CeBwdArgs& a, int grid,
                   cudaStream_t st) {
  if (DP == 64) return colstats ? launch2<64, true>(tx, ty, a, grid, st) : launch2<64, false>(tx, ty, a, grid, st);
  if (DP == 128) return colstats ? launch2<128, true>(tx, ty, a, grid, st) : launch2<128, false>(tx, ty, a, grid, st);
  return colstats ? launch2<256, true>(tx, ty, a, grid, st) : launch2<256, false>(tx, ty, a, grid, st);
}

}  // namespace tt
