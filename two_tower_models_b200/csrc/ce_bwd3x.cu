// In-batch cross-entropy backward, v3 for 128 < d <= 256: the X operand stays in SHARED memory, 128 x 64 score tiles.
//
// Same scheme as ce_bwd3.cu (statistics folded into the score MMA as one extra K = 16 step, two UMMA-issuing warps, score
// buffers released as soon as their values are in registers, separate E buffers, ghost tiles for balance) - what changes is
// where things live, because a 256-column fp32 accumulator takes half of tensor memory:
//   TMEM : S 2 x 64 | E 2 x 32 | acc 256                                   = 448 columns (no room for the X tile)
//   SMEM : X tile 128 x 256 bf16 (64 KB, the A operand of every S = X Y^T, read in place) | A-side bias step | Y stages
// With both operands in shared memory a 128 x 64 x 16 UMMA is bound by shared-memory bandwidth (32 + N/4 = 48 cycles,
// tools/micro/umma_rate.cu); per tile S = 17 x 48, E Y = 4 x 128 (N = 256) -> 1328 tensor cycles against 512 of MUFU: the
// kernel is tensor-bound and 8 epilogue warps suffice.  The constant "ones" operand of the bias step is ONE 1 KB swizzle
// atom addressed with a stride-byte-offset of 0 (every 8-row group reads the same 8 rows): shared memory is the scarce
// resource here (dV pass: 64 + 1 + 4 x 40 KB).
#include <stdlib.h>

#include "ce_common.cuh"

namespace tt {

namespace {

template <int DPV, bool BIAS_X>
struct CfgX {
  static constexpr int DP = DPV;
  static constexpr int BN = DPV == 256 ? 64 : 128;  // TMEM: 2 BN + BN + DP <= 512
  static constexpr int NB = 2, NE = 2;
  static constexpr int EG = BN / 32;                                  // 2 column groups x 4 lane quarters
  static constexpr int THREADS = 128 + EG * 128;
  static constexpr int X_BYTES = 128 * DP * 2;                        // 4 K-atoms of 128 rows x 128 B
  static constexpr int XE_BYTES = BIAS_X ? 128 * 128 : 1024;          // A-side bias step: bias rows (dU) / ones atom (dV)
  static constexpr int ONES_BYTES = BIAS_X ? 1024 : 0;                // B-side ones atom (dU)
  static constexpr int Y_MAIN = BN * DP * 2;                          // 4 K-atoms of 64 rows x 128 B
  static constexpr int EXT_BYTES = BN * 128;
  static constexpr int Y_BYTES = Y_MAIN + (BIAS_X ? 0 : EXT_BYTES);   // dV: the users' bias rows travel with the Y tile
  static constexpr int STAGES = (DPV == 128 && BIAS_X) ? 5 : 4;
  static constexpr int SMEM_BYTES = X_BYTES + XE_BYTES + ONES_BYTES + STAGES * Y_BYTES + 1024 + 512;
  static constexpr int E_COL = NB * BN;
  static constexpr int ACC_COL = E_COL + NE * (BN / 2);
  static_assert(ACC_COL + DP <= 512, "TMEM budget");
  static_assert(SMEM_BYTES <= 232448, "shared memory budget");
};

template <int DPV, bool BIAS_X>
__global__ void __launch_bounds__(CfgX<DPV, BIAS_X>::THREADS, 1)
ce_bwd3x_kernel(const __grid_constant__ TmapSet tmx, const __grid_constant__ TmapSet tmy,
                const __grid_constant__ CUtensorMap tme, const CeBwd3Args a) {
  using Cfg = CfgX<DPV, BIAS_X>;
  constexpr int DP = Cfg::DP, BN = Cfg::BN, EG = Cfg::EG, NB = Cfg::NB, NE = Cfg::NE;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sx = smem;
  uint8_t* sxe = sx + Cfg::X_BYTES;
  uint8_t* sones = sxe + Cfg::XE_BYTES;
  uint8_t* sy = sones + Cfg::ONES_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sy + Cfg::STAGES * Cfg::Y_BYTES);
  uint64_t* x_full = bars;        // TMA landed the X tile
  uint64_t* x_empty = bars + 1;   // every S = X Y^T of the segment has completed (X / bias rows may be replaced)
  uint64_t* xt_full = bars + 2;   // the epilogue warps saw the X tile and (dU) wrote the bias rows of the segment
  uint64_t* acc_full = bars + 3;
  uint64_t* acc_empty = bars + 4;
  uint64_t* s_full = bars + 5;    // [NB]
  uint64_t* s_empty = bars + 7;   // [NB]
  uint64_t* e_full = bars + 9;    // [NE]
  uint64_t* y_full = bars + 11;
  uint64_t* y_empty = y_full + Cfg::STAGES;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(y_empty + Cfg::STAGES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int W_TMA = 4 * EG, W_MMA = 4 * EG + 1, W_ALLOC = 4 * EG + 2, W_MMA2 = 4 * EG + 3;
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
  {
    // the constant ones atom (8 rows x 128 B, K-major, 128-byte swizzle): columns k = 0..2 hold 1.0
    uint8_t* ones = BIAS_X ? sones : sxe;
    if (threadIdx.x < 64) {
      const uint32_t row = threadIdx.x >> 3, ch = threadIdx.x & 7;
      const uint4 v = ch == 0 ? make_uint4(0x3F803F80u, 0x00003F80u, 0u, 0u) : make_uint4(0u, 0u, 0u, 0u);
      *reinterpret_cast<uint4*>(ones + sw128_offset(row, ch)) = v;
    }
    if (BIAS_X) {  // chunks 1..7 of the bias rows stay zero for the whole kernel; chunk 0 is rewritten per segment
      for (int i = threadIdx.x; i < 128 * 8; i += Cfg::THREADS)
        *reinterpret_cast<uint4*>(sxe + sw128_offset(i >> 3, i & 7)) = make_uint4(0u, 0u, 0u, 0u);
    }
    fence_proxy_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  if (warp == W_TMA) {
    if (lane == 0) {
      SegIter it(a.T, a.total, a.CT, a.CTr);
      int r, j0, j1, stage = 0;
      uint32_t phase = 0, xs = 0;
      while (it.next(r, j0, j1)) {
        mbar_wait(x_empty, (xs & 1) ^ 1);
        mbar_arrive_expect_tx(x_full, Cfg::X_BYTES);
        int xrow;
        const CUtensorMap* mx = tmap_of(tmx, r * 128, xrow);
#pragma unroll
        for (int b = 0; b < DP / 64; ++b) tma_load_2d(sx + b * 16384, mx, x_full, b * 64, xrow);
        for (int j = j0; j < j1; ++j) {
          mbar_wait(&y_empty[stage], phase ^ 1);
          mbar_arrive_expect_tx(&y_full[stage], Cfg::Y_BYTES);
          uint8_t* dst = sy + stage * Cfg::Y_BYTES;
          int yrow;
          const CUtensorMap* my = tmap_of(tmy, j * BN, yrow);
#pragma unroll
          for (int b = 0; b < DP / 64; ++b) tma_load_2d(dst + b * (BN * 128), my, &y_full[stage], b * 64, yrow);
          if (!BIAS_X) tma_load_2d(dst + Cfg::Y_MAIN, &tme, &y_full[stage], 0, j * BN);
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
        }
        ++xs;
      }
    }
  } else if (warp == W_MMA) {
    // ---- issuer of the score tiles  S' = X Y^T (+ bias step), both operands in shared memory ----
    const uint32_t leader = elect_one();
    constexpr uint32_t idesc1 = make_idesc_bf16(128, BN, 0, 0);
    SegIter it(a.T, a.total, a.CT, a.CTr);
    int r, j0, j1;
    uint32_t t1 = 0, xs = 0;
    const uint64_t dx0 = make_smem_desc_sw128(smem_u32(sx), 0, 1024);
    const uint64_t dy0 = make_smem_desc_sw128(smem_u32(sy), 0, 1024);
    // bias step: A = bias rows (dU, a full 128-row atom) or the ones atom (dV, stride-byte-offset 0: every 8-row group
    // aliases the same 1 KB); B = the ones atom (dU) or the bias rows that travel with the Y stage (dV)
    const uint64_t dxe = make_smem_desc_sw128(smem_u32(sxe), 0, BIAS_X ? 1024 : 0);
    const uint64_t dones = make_smem_desc_sw128(smem_u32(sones), 0, 0);
    uint32_t y_ok = 0, b_ok = 1;
    while (it.next(r, j0, j1)) {
      mbar_wait(xt_full, xs & 1);
      tc_fence_after();
      for (int j = j0; j < j1; ++j, ++t1) {
        const uint32_t buf = t1 % NB, stage = t1 % Cfg::STAGES;
        mbar_wait_probed(&y_full[stage], (t1 / Cfg::STAGES) & 1, y_ok);
        mbar_wait_probed(&s_empty[buf], ((t1 / NB) & 1) ^ 1, b_ok);
        tc_fence_after();
        y_ok = mbar_probe(&y_full[(t1 + 1) % Cfg::STAGES], ((t1 + 1) / Cfg::STAGES) & 1);
        b_ok = mbar_probe(&s_empty[(t1 + 1) % NB], (((t1 + 1) / NB) & 1) ^ 1);
        const uint64_t dy = desc_advance(dy0, stage * Cfg::Y_BYTES);
#pragma unroll
        for (int k = 0; k < DP / 16; ++k)
          umma_bf16_w(tmem_base + buf * BN, desc_advance(dx0, (k >> 2) * 16384 + (k & 3) * 32),
                      desc_advance(dy, (k >> 2) * (BN * 128) + (k & 3) * 32), idesc1, k > 0 ? 1u : 0u, leader);
        umma_bf16_w(tmem_base + buf * BN, dxe, BIAS_X ? dones : desc_advance(dy, Cfg::Y_MAIN), idesc1, 1u, leader);
        umma_commit_w(&s_full[buf], leader);
      }
      umma_commit_w(x_empty, leader);
      ++xs;
    }
  } else if (warp == W_MMA2) {
    // ---- issuer of  acc += E Y  (E from its TMEM buffer, Y read MN-major, N = 256) ----
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
        mbar_wait_probed(&e_full[buf], (t2 / NE) & 1, e_ok);
        if (j == j0) mbar_wait(acc_empty, (xs & 1) ^ 1);
        tc_fence_after();
        e_ok = mbar_probe(&e_full[(t2 + 1) % NE], ((t2 + 1) / NE) & 1);
        const uint64_t dyt = desc_advance(dyt0, stage * Cfg::Y_BYTES);
#pragma unroll
        for (int k = 0; k < BN / 16; ++k)
          umma_bf16_ta_w(tmem_base + Cfg::ACC_COL, tmem_base + Cfg::E_COL + buf * (BN / 2) + k * 8,
                         desc_advance(dyt, k * 2048), idesc2, (j > j0 || k > 0) ? 1u : 0u, leader);
        umma_commit_w(&y_empty[stage], leader);
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
    const float c0 = log2f(gabs);
    const uint32_t sgn_gs = gsv < 0.f ? 0x80008000u : 0u;
    SegIter it(a.T, a.total, a.CT, a.CTr);
    int r, j0, j1;
    uint32_t t = 0, xs = 0;
    // segment start: the X tile has landed (which also means every score MMA of the previous segment has completed), the
    // bias rows of the new segment go into the A-side atom of the bias step (dU pass)
    auto begin_segment = [&](int rr, uint32_t seg) {
      mbar_wait(x_full, seg & 1);
      if (BIAS_X && e == EG - 1) {
        const long long row2 = (long long)rr * 128 + prow;
        float b = -30000.f;
        if (row2 < a.XR) {
          const float gi = fabsf(__ldg(a.g + row2));
          if (gi > 0.f) b = logf(gi) - __ldg(a.lse + row2);
        }
        const bf16 hi = __float2bfloat16(b);
        const float r1 = b - __bfloat162float(hi);
        const bf16 mid = __float2bfloat16(r1);
        const bf16 lo = __float2bfloat16(r1 - __bfloat162float(mid));
        *reinterpret_cast<uint4*>(sxe + sw128_offset(prow, 0)) =
            make_uint4((uint32_t)__bfloat16_as_ushort(hi) | ((uint32_t)__bfloat16_as_ushort(mid) << 16),
                       (uint32_t)__bfloat16_as_ushort(lo), 0u, 0u);
        fence_proxy_async_smem();
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(xt_full);
    };
    bool have = it.next(r, j0, j1);
    if (have) begin_segment(r, 0);
    while (have) {
      const long long row = (long long)r * 128 + prow;
      const long long tgt = row + a.diag_shift;
      const bool haspos = row < a.XR && tgt >= 0 && tgt < a.YR;
      const int jd = haspos ? (int)(tgt / BN) : -1;
      const bool mine = haspos && (int)((tgt % BN) >> 5) == e;
      const int od = haspos ? (int)((tgt % BN) & 31) : 0;
      float sub = 0.f;
      uint32_t sgn_row = sgn_gs;
      if (BIAS_X) {
        const float gi = row < a.XR ? __ldg(a.g + row) : 0.f;
        sub = fabsf(gi) * gabs;
        if (gi < 0.f) sgn_row ^= 0x80008000u;
      } else if (mine) {
        sub = fabsf(__ldg(a.g + tgt)) * gabs;
      }
      for (int j = j0; j < j1; ++j, ++t) {
        const uint32_t buf = t % NB;
        uint32_t cmask = 0u;
        if (!BIAS_X) cmask = __ldg(a.signmask + j * (BN / 32) + e);
        mbar_wait(&s_full[buf], (t / NB) & 1);
        tc_fence_after();
        float v[32];
        tmem_ld32(lane_base + buf * BN + e * 32, v);
        tmem_wait_ld();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_empty[buf]);
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = ex2f(fmaf(v[i], LOG2E, c0));
        if (mine && j == jd) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] -= (i == od) ? sub : 0.f;
        }
        uint32_t p[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) p[i] = pack_bf16x2(v[2 * i], v[2 * i + 1]) ^ sgn_row;
        if (!BIAS_X && cmask != 0u) {
#pragma unroll
          for (int i = 0; i < 16; ++i)
            p[i] ^= (((cmask >> (2 * i)) & 1u) << 15) | (((cmask >> (2 * i + 1)) & 1u) << 31);
        }
        if (t >= (uint32_t)NE) {  // E buffer free once acc += E(t - NE) Y(t - NE) has completed
          mbar_wait(&y_empty[(t - NE) % Cfg::STAGES], ((t - NE) / Cfg::STAGES) & 1);
          tc_fence_after();
        }
        tmem_st16(lane_base + Cfg::E_COL + (t % NE) * (BN / 2) + e * 16, p);
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&e_full[t % NE]);
      }
      int r2 = 0, j02 = 0, j12 = 0;
      const bool have2 = it.next(r2, j02, j12);
      if (have2) begin_segment(r2, xs + 1);
      // segment accumulator -> partial slot (group e drains the 32-column parts with part % EG == e)
      mbar_wait(acc_full, xs & 1);
      tc_fence_after();
      {
        const int slot = (int)(blockIdx.x - ((long long)r * a.CT) / a.T);
        float* dst = a.partial + (long long)slot * a.slot_stride + row * DP;
#pragma unroll 1
        for (int part = e; part < DP / 32; part += EG) {
          float w[32];
          tmem_ld32(lane_base + Cfg::ACC_COL + part * 32, w);
          tmem_wait_ld();
          if (part + EG >= DP / 32) {  // last part of this warp: the accumulator columns are in registers
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(acc_empty);
          }
#pragma unroll
          for (int i = 0; i < 8; ++i)
            *reinterpret_cast<float4*>(dst + part * 32 + 4 * i) = make_float4(w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]);
        }
      }
      r = r2; j0 = j02; j1 = j12; have = have2;
      ++xs;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == W_ALLOC) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

template <int DPV, bool BIAS_X>
int launch3x(const TmapSet& tx, const TmapSet& ty, const CUtensorMap& te, const CeBwd3Args& a, int grid, cudaStream_t st) {
  using Cfg = CfgX<DPV, BIAS_X>;
  static bool configured = false;
  if (!configured) {
    TT_CUDA(cudaFuncSetAttribute(ce_bwd3x_kernel<DPV, BIAS_X>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    configured = true;
  }
  KernelSpan span(BIAS_X ? "ce_bwd3x_kernel_dU" : "ce_bwd3x_kernel_dV", st);
  ce_bwd3x_kernel<DPV, BIAS_X><<<grid, Cfg::THREADS, Cfg::SMEM_BYTES, st>>>(tx, ty, te, a);
  TT_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

}  // namespace

int launch_ce_bwd3x(int DP, bool bias_x, const TmapSet& tx, const TmapSet& ty, long long users, const void* ext, CeBwd3Args a,
                    int grid, cudaStream_t st) {
  const long long pad = (users + 127) / 128 * 128;
  a.signmask = reinterpret_cast<const uint32_t*>(reinterpret_cast<const uint8_t*>(ext) + (size_t)pad * 128);
  CUtensorMap te;
  int rc = make_tmap_bf16(&te, ext, 64, (uint64_t)pad, 64, 64, DP == 256 ? 64 : 128);
  if (rc) return rc;
  if (DP == 128) return bias_x ? launch3x<128, true>(tx, ty, te, a, grid, st) : launch3x<128, false>(tx, ty, te, a, grid, st);
  return bias_x ? launch3x<256, true>(tx, ty, te, a, grid, st) : launch3x<256, false>(tx, ty, te, a, grid, st);
}

}  // namespace tt
