// Fused in-batch sampled-softmax cross entropy for sm_100a: forward kernel, merge kernels, host side of the backward.
//
// Reference semantics: src/two_tower_base_retrieval.py:287 (S = U V^T), :301 (target = arange),
// :310-312 (F.cross_entropy(reduction="none")), :322-343 (label weights, batch max, weighted mean) and autograd
// of the same (train/train.py:124).  The [B,N] score matrix never exists in HBM: 128x128 score tiles are
// produced by tcgen05.mma into TMEM from TMA-staged bf16 operand tiles and consumed in place.
//
//   forward : per 128-row tile, running (max, sum-exp) over the column tiles (ce_fwd_kernel); a merge kernel folds
//             the per-(slot, column group) partials, recomputes the positive's logit from the operands and - with the
//             identity debias hook - also forms the label weights, their batch maximum and the weighted mean.
//   backward: ce_bwd2.cu runs  acc[128, d] = sum_j E_j * Y_j,  E_j = f(X Y_j^T)  twice:
//       pass A  X=U, Y=V, E_ij = g_i (exp(S_ij - lse_i) - [j == i+off])           -> dU
//       pass B  X=V, Y=U, E_ji = g_i (exp(S_ij - lse_i) - [j == i+off]) (col stats) -> dV
//     and ce_bwd_reduce_kernel (here) merges the slot partials of both passes in one launch.
//
// Work is the flattened (row tile, column tile) space cut into equal contiguous ranges, one per SM
// (persistent CTAs).  A CTA's range touches <= a few row tiles ("segments"); every segment writes a
// partial result into a slot and a small second kernel merges the slots.
//
// Warp roles of ce_fwd_kernel (640 threads): warp 0 TMA producer, warp 1 UMMA issuer, warp 2 TMEM allocator,
// warps 4-19 epilogue: 4 column groups x 4 TMEM lane quarters, every thread owns 32 columns of one row of each tile.
#include <stdlib.h>

#include "ce_common.cuh"
#include "kernels.h"

namespace tt {

// =============================================================================================
// Forward
// =============================================================================================
struct CeFwdArgs {
  int B, N;
  long long target_offset;
  long long T, total;
  int CT;
  long long Bpad;
  float* part_m;
  float* part_s;
  struct LossSync* sync;  // ticket of the fused combine+loss kernel, zeroed here
  unsigned skew_ns;       // start delay of the odd epilogue groups
  // optional zero fills riding in this launch (tt_inbatch_ce_attach_zero_fill): an otherwise idle warp streams zeros
  // from shared memory with TMA bulk stores while the scoring tiles run
  uint8_t* zero_ptr[2];
  long long zero_bytes[2];
  long long* trace;  // bring-up (TT_CE_TRACE)
  int dbg;           // bring-up (TT_CE_DBG): bit0 skip ex2, bit1 skip the maximum, bit2 skip the whole tile update
};

// epilogue column groups of the forward kernel: FWD_EG x 4 warps, each thread owns 128 / FWD_EG columns of a row
static constexpr int FWD_EG = 4;
static constexpr int FWD_THREADS = 128 + FWD_EG * 128;

template <int DP>
struct CeFwdCfg {
  static constexpr int BN = 128;
  static constexpr int NS = 4;  // score-tile buffers in TMEM (4 x 128 columns)
  static constexpr int KBOX = DP / 64;
  static constexpr int X_BYTES = 128 * DP * 2;
  static constexpr int Y_BYTES = BN * DP * 2;
  static constexpr int STAGES = DP == 64 ? 6 : (DP == 128 ? 4 : 2);
  static constexpr int ZERO_BYTES = 16384;  // source of the fused zero fills
  static constexpr int SMEM_BYTES = X_BYTES + STAGES * Y_BYTES + ZERO_BYTES + 1024 + 256;
};

// Both epilogue groups work on EVERY score tile, each on one half of its columns (group e: columns
// [64 e, 64 e + 64)), so all 8 epilogue warps are busy on the tile that is ready while the UMMA warp runs up
// to NS - 1 tiles ahead.  A row's running (max, sum-exp) is therefore split over two threads; the combine
// kernel merges the (slot, group) partials.
// bring-up hooks (clock64 timelines, partial epilogues) exist only when compiled with -DTT_CE_BRINGUP
#ifdef TT_CE_BRINGUP
#define CE_FWD_STAMP(tile, which)                                                                              \
  do {                                                                                                         \
    if (a.trace && !(a.dbg & 16) && blockIdx.x == 0 && (tile) < 64 && q == 0 && lane == 0 && e < 2)              \
      a.trace[((2 + e) * 64 + (tile)) * 2 + (which)] = clock64();                                               \
  } while (0)
#else
#define CE_FWD_STAMP(tile, which) do { } while (0)
#endif

template <int DP>
__global__ void __launch_bounds__(FWD_THREADS, 1)
ce_fwd_kernel(const __grid_constant__ CUtensorMap tmx, const __grid_constant__ TmapSet tmy, const CeFwdArgs a) {
  using Cfg = CeFwdCfg<DP>;
  constexpr int BN = Cfg::BN, NS = Cfg::NS;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sx = smem;
  uint8_t* sy = smem + Cfg::X_BYTES;
  uint8_t* sz = sy + Cfg::STAGES * Cfg::Y_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sz + Cfg::ZERO_BYTES);
  uint64_t* x_full = bars;
  uint64_t* x_empty = bars + 1;
  uint64_t* s_full = bars + 2;
  uint64_t* s_empty = s_full + NS;
  uint64_t* y_full = s_empty + NS;
  uint64_t* y_empty = y_full + Cfg::STAGES;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(y_empty + Cfg::STAGES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmx);
    for (int p = 0; p < tmy.n; ++p) tma_prefetch_desc(&tmy.m[p]);
  }
  if (blockIdx.x == 0 && threadIdx.x == 96) *reinterpret_cast<unsigned int*>(a.sync) = 0u;  // workspace is uninitialised
  if (warp == 1 && lane == 0) {
    mbar_init(x_full, 1);
    mbar_init(x_empty, 1);
    for (int i = 0; i < Cfg::STAGES; ++i) {
      mbar_init(&y_full[i], 1);
      mbar_init(&y_empty[i], 1);
    }
    for (int i = 0; i < NS; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&s_empty[i], 4 * FWD_EG);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_holder, NS * BN);
  const bool zfill = a.zero_bytes[0] > 0 || a.zero_bytes[1] > 0;
  if (zfill) {
    for (int i = threadIdx.x; i < Cfg::ZERO_BYTES / 16; i += FWD_THREADS) reinterpret_cast<uint4*>(sz)[i] = make_uint4(0u, 0u, 0u, 0u);
    fence_proxy_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;
#ifdef TT_CE_BRINGUP
  if ((a.dbg & 16) && a.trace && blockIdx.x == 0 && threadIdx.x == 128) {
    unsigned long long gt;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    a.trace[0] = clock64();
    a.trace[1] = (long long)gt;
  }
#endif

  if (warp == 0) {
    if (lane == 0) {
      SegIter it(a.T, a.total, a.CT);
      int r, j0, j1, stage = 0;
      uint32_t phase = 0, xs = 0;
      while (it.next(r, j0, j1)) {
        mbar_wait(x_empty, (xs & 1) ^ 1);
        mbar_arrive_expect_tx(x_full, Cfg::X_BYTES);
#pragma unroll
        for (int b = 0; b < Cfg::KBOX; ++b) tma_load_2d(sx + b * 16384, &tmx, x_full, b * 64, r * 128);
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
    {  // the whole warp walks the schedule; only the elected lane issues tcgen05 instructions
      const uint32_t leader = elect_one();
      constexpr uint32_t idesc = make_idesc_bf16(128, BN, 0, 0);
      SegIter it(a.T, a.total, a.CT);
      int r, j0, j1, stage = 0;
      uint32_t phase = 0, xs = 0, t = 0;
      const uint64_t dx0 = make_smem_desc_sw128(smem_u32(sx), 0, 1024);
      const uint64_t dy0 = make_smem_desc_sw128(smem_u32(sy), 0, 1024);
      while (it.next(r, j0, j1)) {
        mbar_wait(x_full, xs & 1);
        for (int j = j0; j < j1; ++j, ++t) {
          const uint32_t buf = t % NS, use = t / NS;
#ifdef TT_CE_BRINGUP
          if (a.trace && !(a.dbg & 16) && blockIdx.x == 0 && t < 64 && leader) a.trace[(0 * 64 + t) * 2 + 0] = clock64();
#endif
          mbar_wait(&y_full[stage], phase);
#ifdef TT_CE_BRINGUP
          if (a.trace && !(a.dbg & 16) && blockIdx.x == 0 && t < 64 && leader) a.trace[(1 * 64 + t) * 2 + 0] = clock64();
#endif
          mbar_wait(&s_empty[buf], (use & 1) ^ 1);
          tc_fence_after();
#ifdef TT_CE_BRINGUP
          if (a.trace && !(a.dbg & 16) && blockIdx.x == 0 && t < 64 && leader) a.trace[(0 * 64 + t) * 2 + 1] = clock64();
#endif
          const uint64_t dy = desc_advance(dy0, stage * Cfg::Y_BYTES);
#pragma unroll
          for (int k = 0; k < DP / 16; ++k)
            umma_bf16_w(tmem_base + buf * BN, desc_advance(dx0, (k >> 2) * 16384 + (k & 3) * 32),
                        desc_advance(dy, (k >> 2) * (BN * 128) + (k & 3) * 32), idesc, k > 0 ? 1u : 0u, leader);
          umma_commit_w(&y_empty[stage], leader);
          umma_commit_w(&s_full[buf], leader);
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit_w(x_empty, leader);
        ++xs;
      }
    }
  } else if (warp == 3) {
    // fused zero fills (the dense embedding-table gradients of the step: 2 x 51 MB at the benchmark shape): 16 KB chunks,
    // CTA-strided, all issued up front - the TMA engine drains them beside the scoring tiles, no SM instruction issue
    if (zfill && lane == 0) {
      constexpr long long CH = Cfg::ZERO_BYTES;
      for (int z = 0; z < 2; ++z) {
        const long long n = a.zero_bytes[z];
        for (long long c = blockIdx.x; c * CH < n; c += gridDim.x) {
          const long long rem = n - c * CH;
          const uint32_t bytes = (uint32_t)(rem < CH ? rem : CH);
          asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                       ::"l"(a.zero_ptr[z] + c * CH), "r"(smem_u32(sz)), "r"(bytes)
                       : "memory");
        }
      }
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
  } else if (warp >= 4) {
    const int e = (warp - 4) >> 2;  // column group of every tile
    const int q = warp & 3;         // TMEM lane quarter
    constexpr int CW = BN / FWD_EG;  // columns per thread and tile
    constexpr int NL = CW / 32;      // 32-column TMEM loads per thread and tile
    static_assert(BN == 128 && (CW == 32 || CW == 64), "epilogue column groups");
    SegIter it(a.T, a.total, a.CT);
    int r, j0, j1;
    uint32_t t = 0;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16) + e * CW;
    const uint32_t sfull0 = smem_u32(s_full), sempty0 = smem_u32(s_empty);
    // The per-tile bookkeeping is kept to a handful of instructions (32-bit indices, tile-level special-case tests,
    // bring-up hooks compiled out): with it at ~180 instructions per warp and tile the epilogue was issue-bound at
    // half of the MUFU rate (profiles/r01_ce_fwd_inst_mix.txt).
    // TMEM -> registers of this thread's CW scores of tile `tt` (asynchronous: complete after tmem_wait_ld)
    auto load_tile = [&](uint32_t tt, float* x) {
      const uint32_t buf = tt & (NS - 1);
      mbar_wait_addr(sfull0 + buf * 8, (tt / NS) & 1);
      tc_fence_after();
      CE_FWD_STAMP(tt, 0);
#ifdef TT_CE_BRINGUP
      if (a.dbg & 32) {  // synthetic scores instead of the TMEM load
#pragma unroll
        for (int i = 0; i < CW; ++i) x[i] = fmaf(__uint_as_float(tt), 0.001f * (float)(i + 1), -0.05f * (float)i);
        return;
      }
#endif
#pragma unroll
      for (int l = 0; l < NL; ++l) tmem_ld32(lane_addr + buf * BN + l * 32, x + l * 32);
    };
    // the score buffer goes back to the UMMA warp as soon as its values sit in registers
    auto release = [&](uint32_t tt) {
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_addr(sempty0 + (tt & (NS - 1)) * 8);
    };
    static_assert((NS & (NS - 1)) == 0, "NS must be a power of two");
    // The warps that share a scheduler (same q, different column group) would otherwise run in lock step - both in
    // their MUFU phase, then both in their FMA/max phase - and the two pipes never overlap.  Starting every second
    // group about half a tile late keeps them out of phase for the rest of the kernel (the score buffers give slack).
    if (a.skew_ns > 0 && (e & 1)) __nanosleep(a.skew_ns);
    while (it.next(r, j0, j1)) {
      const int row = r * 128 + q * 32 + lane;
      const bool valid = row < a.B;
      // the (only) tile with columns past N is the one special case; the positive's logit is not picked out of the
      // score tiles at all - the merge kernel recomputes that one dot product per row from the operands
      const int jpart = (a.N % BN) != 0 ? a.N / BN : -1;
      float m = -INFINITY, s = 0.f;
      // online (max, sum-exp) update with this thread's CW scores of a tile: one maximum and one rescale, then CW
      // independent exponentials (in place, back to back on the MUFU pipe) and their sum
      auto process = [&](float* x, int j) {
#ifdef TT_CE_BRINGUP
        if (a.dbg & 4) { s += x[0]; return; }
#endif
        if (j == jpart) {
          asm volatile("");  // keep this a branch: if-converted it costs 2 * CW instructions on every tile
          const int n0 = j * BN + e * CW;
#pragma unroll
          for (int i = 0; i < CW; ++i)
            if (n0 + i >= a.N) x[i] = -INFINITY;
        }
        float c[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) c[u] = fmaxf(x[u], x[u + 4]);
#pragma unroll
        for (int i = 8; i < CW; i += 4) {
#pragma unroll
          for (int u = 0; u < 4; ++u) c[u] = fmaxf(c[u], x[i + u]);
        }
        const float cm = fmaxf(fmaxf(c[0], c[1]), fmaxf(c[2], c[3]));
        const float m_new = fmaxf(m, cm * LOG2E);
        const float ms = (m_new == -INFINITY) ? 0.f : m_new;  // fully masked so far: avoid inf - inf
        s *= ex2f(m - ms);
#ifdef TT_CE_BRINGUP
        if (a.dbg & 1) {
#pragma unroll
          for (int i = 0; i < CW; ++i) x[i] = fmaf(x[i], LOG2E, -ms);
        } else
#endif
        {
#ifdef TT_CE_FWD_POLY
          // every fourth exponential on the FMA pipe (the forward epilogue is MUFU-bound: XU 71 %, tensor 34 %)
#pragma unroll
          for (int i = 0; i < CW; ++i) {
            const float y = fmaf(x[i], LOG2E, -ms);
            x[i] = (i & 3) == 3 ? exp2_poly3(y) : ex2f(y);
          }
#else
#pragma unroll
          for (int i = 0; i < CW; ++i) x[i] = ex2f(fmaf(x[i], LOG2E, -ms));
#endif
        }
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int i = 0; i < CW; i += 4) {
#pragma unroll
          for (int u = 0; u < 4; ++u) acc[u] += x[i + u];
        }
        s += (acc[0] + acc[1]) + (acc[2] + acc[3]);
        m = m_new;
      };
      const int n = j1 - j0;
      if (FWD_EG >= 4) {
        // 4 warps per scheduler: the other warps cover the TMEM load latency, one register set per thread suffices
        float xs[CW];
        for (int i = 0; i < n; ++i) {
          load_tile(t + i, xs);
          tmem_wait_ld();
          release(t + i);
          process(xs, j0 + i);
          CE_FWD_STAMP(t + i, 1);
        }
      } else {
      float xa[CW], xb[CW];
      load_tile(t, xa);
      for (int i = 0; i < n; i += 2) {
        tmem_wait_ld();
        release(t + i);
        if (i + 1 < n) load_tile(t + i + 1, xb);
        process(xa, j0 + i);
        CE_FWD_STAMP(t + i, 1);
        if (i + 1 < n) {
          tmem_wait_ld();
          release(t + i + 1);
          if (i + 2 < n) load_tile(t + i + 2, xa);
          process(xb, j0 + i + 1);
          CE_FWD_STAMP(t + i + 1, 1);
        }
      }
      }
      t += n;
      if (valid) {
        const int slot = (int)(blockIdx.x - ((long long)r * a.CT) / a.T);
        const long long o = (long long)(slot * FWD_EG + e) * a.Bpad + row;
        a.part_m[o] = m;
        a.part_s[o] = s;
      }
    }
#ifdef TT_CE_BRINGUP
    if ((a.dbg & 16) && a.trace && blockIdx.x == 0 && threadIdx.x == 128) {
      unsigned long long gt;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
      a.trace[2] = clock64();
      a.trace[3] = (long long)gt;
    }
#endif
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, NS * BN);
  }
}

// The positive's logit U_i . V_{i+off}: fp32 dot product of the same bf16 operands the score tiles were made from
// (the tensor cores accumulate the same exact products in fp32, in another order).
struct DiagOperands {
  const bf16* U;
  long long ldu;
  const bf16* Vp[8];
  long long rows_per_part, ldv, target_offset;
  int d;
};
static constexpr int COMBINE_RB = 64;  // rows per 256-thread block of the merge kernels
// Quad layout of the merge kernels: 256-thread blocks own COMBINE_RB = 64 rows, thread (row, e) = (tid >> 2, tid & 3).
// All loads of a thread are issued before the first use (the merge is a chain of L2 round trips otherwise).
// Positive logit of `row`: the quad splits the row's 16-byte chunks; valid in all four lanes afterwards.
__device__ __forceinline__ float positive_logit_quad(const DiagOperands& o, long long row, int B, int e) {
  const int chunks = (o.d + 7) / 8;  // <= 32
  uint4 a4[8], b4[8];
  const bool live = row < B;
  const long long t = row + o.target_offset;
  const long long p = live ? t / o.rows_per_part : 0;
  const bf16* urow = o.U + row * o.ldu;
  const bf16* vrow = o.Vp[p] + (t - p * o.rows_per_part) * o.ldv;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = e + 4 * i;
    const bool ok = live && c < chunks;
    a4[i] = ok ? *reinterpret_cast<const uint4*>(urow + c * 8) : make_uint4(0, 0, 0, 0);
    b4[i] = ok ? *reinterpret_cast<const uint4*>(vrow + c * 8) : make_uint4(0, 0, 0, 0);
  }
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = e + 4 * i;
    const __nv_bfloat162* a2 = reinterpret_cast<const __nv_bfloat162*>(&a4[i]);
    const __nv_bfloat162* b2 = reinterpret_cast<const __nv_bfloat162*>(&b4[i]);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 fa = __bfloat1622float2(a2[k]), fb = __bfloat1622float2(b2[k]);
      if (c * 8 + 2 * k < o.d) acc = fmaf(fa.x, fb.x, acc);
      if (c * 8 + 2 * k + 1 < o.d) acc = fmaf(fa.y, fb.y, acc);
    }
  }
  acc += __shfl_xor_sync(0xffffffffu, acc, 1);
  acc += __shfl_xor_sync(0xffffffffu, acc, 2);
  return acc;
}
// log-sum-exp (natural log) of `row` from the forward kernel's (max, sum-exp) partials in base 2: lane e of the quad
// merges epilogue group e of every slot (four slots per batch of loads), then the quad folds.  Valid in all four lanes.
__device__ __forceinline__ float merge_lse_quad(long long row, int B, long long Bpad, long long T, int CT,
                                                const float* __restrict__ part_m, const float* __restrict__ part_s, int e) {
  static_assert(FWD_EG == 4, "one quad lane per epilogue group");
  const bool live = row < B;
  const long long r = row / 128;
  const int first = (int)((r * CT) / T), nsl = live ? (int)(((r + 1) * CT - 1) / T) - first + 1 : 0;
  float M = -INFINITY, S = 0.f;
  for (int s0 = 0; s0 < nsl; s0 += 4) {
    float m[4], sm[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const bool ok = s0 + u < nsl;
      const long long o = (long long)((s0 + u) * FWD_EG + e) * Bpad + row;
      m[u] = ok ? part_m[o] : -INFINITY;
      sm[u] = ok ? part_s[o] : 0.f;
    }
    // an epilogue group without valid columns reports (-inf, 0): the clamp keeps -inf - -inf out of the exponent
    const float mx = fmaxf(fmaxf(fmaxf(fmaxf(m[0], m[1]), fmaxf(m[2], m[3])), M), -3.0e38f);
    S *= exp2f(M - mx);
#pragma unroll
    for (int u = 0; u < 4; ++u) S += sm[u] * exp2f(m[u] - mx);
    M = mx;
  }
#pragma unroll
  for (int off = 1; off <= 2; off <<= 1) {
    const float Mo = __shfl_xor_sync(0xffffffffu, M, off), So = __shfl_xor_sync(0xffffffffu, S, off);
    const float mx = fmaxf(fmaxf(M, Mo), -3.0e38f);
    S = (live ? S * exp2f(M - mx) + So * exp2f(Mo - mx) : 0.f);
    M = mx;
  }
  return live ? (M + log2f(S)) * LN2 : 0.f;
}

__global__ void __launch_bounds__(256)
ce_combine_kernel(int B, long long Bpad, long long T, int CT, const float* part_m, const float* part_s,
                  const DiagOperands dg, float* ce, float* lse) {
  const int e = threadIdx.x & 3;
  const long long row = (long long)blockIdx.x * COMBINE_RB + (threadIdx.x >> 2);
  const float pos = positive_logit_quad(dg, row, B, e);
  const float l = merge_lse_quad(row, B, Bpad, T, CT, part_m, part_s, e);
  if (e == 0 && row < B) {
    lse[row] = l;
    ce[row] = l - pos;
  }
}

// combine + value-weighted mean in ONE launch (identity debias hook, reference :322-343).  Every block merges the
// (max, sum-exp) partials of 64 rows into ce / lse, forms nuv = max(labels . w, 1e-6) and reduces (max nuv,
// sum ce nuv) over its rows; the last block to finish (atomic ticket) folds the per-block pairs in block order
// (deterministic) into  loss = sum / (max B)  and  g_norm = 1 / (max B).  The backward kernels take g = nuv and
// the device scalar g_norm, so the normalised weights never make a separate pass.
static constexpr int LOSS_MAX_BLOCKS = 1024;
struct LossSync {
  unsigned int ticket;
  unsigned int pad[3];
  float pmax[LOSS_MAX_BLOCKS];
  float psum[LOSS_MAX_BLOCKS];
};
__global__ void __launch_bounds__(256)
ce_combine_loss_kernel(int B, long long Bpad, long long T, int CT, const float* __restrict__ part_m,
                       const float* __restrict__ part_s, const DiagOperands dg, float* __restrict__ ce,
                       float* __restrict__ lse, const float* __restrict__ labels, long long ldl,
                       const float* __restrict__ uvw, int TL, float inv_rows, float* __restrict__ loss,
                       float* __restrict__ g, float* __restrict__ g_norm, LossSync* sync, float* __restrict__ stats) {
  __shared__ float red_m[8], red_s[8];
  __shared__ unsigned int last;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, e = tid & 3;
  const long long row = (long long)blockIdx.x * COMBINE_RB + (tid >> 2);
  const float pos = positive_logit_quad(dg, row, B, e);
  const float l = merge_lse_quad(row, B, Bpad, T, CT, part_m, part_s, e);
  float nuv = 0.f, cw = 0.f;
  if (e == 0 && row < B) {
    const float c = l - pos;
    lse[row] = l;
    ce[row] = c;
    for (int t = 0; t < TL; ++t) nuv = fmaf(labels[row * ldl + t], uvw[t], nuv);
    nuv = fmaxf(nuv, 0.000001f);
    g[row] = nuv;
    cw = c * nuv;
  }
  float mx = nuv;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    cw += __shfl_xor_sync(0xffffffffu, cw, o);
  }
  if (lane == 0) { red_m[warp] = mx; red_s[warp] = cw; }
  __syncthreads();
  if (tid == 0) {
    float m = red_m[0], sm = red_s[0];
    for (int w = 1; w < 8; ++w) { m = fmaxf(m, red_m[w]); sm += red_s[w]; }
    sync->pmax[blockIdx.x] = m;
    sync->psum[blockIdx.x] = sm;
    __threadfence();
    last = (atomicAdd(&sync->ticket, 1u) == gridDim.x - 1) ? 1u : 0u;
  }
  __syncthreads();
  if (last && warp == 0) {
    __threadfence();
    float m = 0.f, sm = 0.f;
    for (int i = lane; i < (int)gridDim.x; i += 32) {  // lane-strided, then a fixed-order butterfly: deterministic
      m = fmaxf(m, __ldcg(&sync->pmax[i]));
      sm += __ldcg(&sync->psum[i]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
      sm += __shfl_xor_sync(0xffffffffu, sm, o);
    }
    if (lane == 0) {
      const float gn = inv_rows / m;
      *loss = sm * gn;
      *g_norm = gn;
      if (stats != nullptr) {  // batch-sharded loss: this rank's (max nuv, sum ce nuv), merged across ranks afterwards
        stats[0] = m;
        stats[1] = sm;
      }
      sync->ticket = 0;  // ready for the next launch (CUDA-graph replay)
    }
  }
}

__global__ void set_scalar_kernel(float* p, float v) { *p = v; }

// zero fills attached to the next forward launch of the calling thread
static thread_local struct { void* p[2]; long long n[2]; } g_zero_jobs = {{nullptr, nullptr}, {0, 0}};
int inbatch_ce_attach_zero_fill(void* p0, long long bytes0, void* p1, long long bytes1) {
  TT_CHECK(((uintptr_t)p0 % 16) == 0 && ((uintptr_t)p1 % 16) == 0 && bytes0 >= 0 && bytes1 >= 0 && bytes0 % 16 == 0 && bytes1 % 16 == 0,
           "inbatch_ce_attach_zero_fill: buffers need 16-byte alignment and sizes that are multiples of 16 bytes");
  TT_CHECK(g_zero_jobs.n[0] == 0 && g_zero_jobs.n[1] == 0, "inbatch_ce_attach_zero_fill: zero fills are already pending");
  g_zero_jobs.p[0] = p0; g_zero_jobs.n[0] = p0 ? bytes0 : 0;
  g_zero_jobs.p[1] = p1; g_zero_jobs.n[1] = p1 ? bytes1 : 0;
  return 0;
}

// Batch-sharded loss: fold the all-gathered per-rank (max nuv, sum ce nuv) pairs into the global weighted mean
// loss = sum_r s_r / (max_r m_r * rows) and the scalar g_norm = 1 / (max * rows) the backward kernels multiply into g.
__global__ void sharded_loss_finalize_kernel(const float* __restrict__ stats_all, int world, float inv_rows,
                                             float* __restrict__ loss, float* __restrict__ g_norm) {
  float m = 0.f, sm = 0.f;
  for (int r = 0; r < world; ++r) {  // fixed order: every rank computes bit-identical scalars
    m = fmaxf(m, stats_all[2 * r]);
    sm += stats_all[2 * r + 1];
  }
  const float gn = inv_rows / m;
  *loss = sm * gn;
  *g_norm = gn;
}
int sharded_loss_finalize(const float* stats_all, int world, long long rows, float* loss, float* g_norm, cudaStream_t stream) {
  TT_CHECK(stats_all && loss && g_norm && world >= 1 && rows > 0, "sharded_loss_finalize: bad arguments");
  KernelSpan span("sharded_loss_finalize_kernel", stream);
  sharded_loss_finalize_kernel<<<1, 1, 0, stream>>>(stats_all, world, 1.f / (float)rows, loss, g_norm);
  TT_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

static int pick_dp(long long d) { return d <= 64 ? 64 : (d <= 128 ? 128 : 256); }

// Tensor maps of an [rows, d] operand given as `np` equally sized row blocks (np == 1: one matrix)
static int make_tmap_set(TmapSet* t, const void* const* parts, int np, long long rows_per_part, long long rows,
                         long long d, long long ld, int box_rows) {
  TT_CHECK(np >= 1 && np <= 8, "in-batch CE: 1..8 operand parts supported (got %d)", np);
  TT_CHECK(np == 1 || (rows_per_part % 128 == 0 && rows_per_part * np == rows),
           "in-batch CE: operand parts must be equally sized multiples of 128 rows (%lld x %d != %lld)", rows_per_part, np, rows);
  t->n = np;
  t->rows_per_map = (int)(np == 1 ? rows : rows_per_part);
  for (int p = 0; p < np; ++p) {
    TT_CHECK(((uintptr_t)parts[p] % 16) == 0, "in-batch CE: operand parts need 16-byte alignment");
    const int rc = make_tmap_bf16(&t->m[p], parts[p], d, np == 1 ? rows : rows_per_part, ld, 64, box_rows);
    if (rc) return rc;
  }
  for (int p = np; p < 8; ++p) t->m[p] = t->m[0];
  return 0;
}

static size_t fwd_ws_bytes(const Sched& s, long long Bpad) {
  return (size_t)(2 * (size_t)s.max_slots * FWD_EG * Bpad) * sizeof(float) + sizeof(LossSync);
}
static size_t bwd_ws_bytes(const Sched& s, int DP) { return (size_t)s.max_slots * s.XT * 128 * DP * sizeof(float); }

// v3 backward kernels (ce_bwd3.cu, d <= 128) unless TT_CE_BWD_V3=0
static bool use_bwd_v3() {
  static const bool on = !(getenv("TT_CE_BWD_V3") && atoi(getenv("TT_CE_BWD_V3")) == 0);
  return on;
}
// columns of Y per score tile of the backward kernels
// TT_CE_BWD_X128=1: the d = 128 problem on the shared-memory-operand variant (ce_bwd3x.cu, 128-wide tiles, 16 epilogue warps)
static bool use_x128() {
  static const bool on = getenv("TT_CE_BWD_X128") && atoi(getenv("TT_CE_BWD_X128")) != 0;
  return on;
}
static int bwd_tile_cols(int DP, bool v3) {
  if (v3 && DP == 128 && use_x128()) return 128;
  return (v3 && DP <= 128) ? ce_bwd3_tile_cols(DP) : (DP == 256 ? 64 : 128);
}
// ghost tiles per row tile (ce_common.cuh: SegIter): what a row-tile boundary inside a CTA's range costs the v3 kernels
// (accumulator drain at ~32 B/clk + pipeline refill = ~4 tiles, profiles/r02_ce_bwd3_timeline_*.txt)
static int bwd_ghost(int DP, bool v3) {
  static const int g = getenv("TT_CE_BWD_GHOST") ? atoi(getenv("TT_CE_BWD_GHOST")) : 4;
  return v3 ? g : 0;
}
// the d = 256 variant of the v3 kernels (ce_bwd3x.cu) unless TT_CE_BWD_V3X=0
static bool use_bwd_v3x() {
  static const bool on = !(getenv("TT_CE_BWD_V3X") && atoi(getenv("TT_CE_BWD_V3X")) == 0);
  return on;
}
static Sched bwd_sched(long long xr, long long yr, int DP, bool v3) { return make_sched(xr, yr, bwd_tile_cols(DP, v3), bwd_ghost(DP, v3)); }

// forward / backward partials (sized for either backward kernel generation); the v3 backward's ext block (bias rows +
// sign words of the users) follows at this offset
static size_t ce_ws_base(long long B, long long N, long long d) {
  const int DP = pick_dp(d);
  const long long Bpad = (B + 127) / 128 * 128;
  size_t best = fwd_ws_bytes(make_sched(B, N, 128), Bpad);
  size_t bmax = 0, cmax = 0;  // the two backward passes use disjoint regions: [dU partials | dV partials]
  for (int v3 = 0; v3 < 2; ++v3) {
    const size_t b = (bwd_ws_bytes(bwd_sched(B, N, DP, v3 != 0), DP) + 255) / 256 * 256;
    const size_t c = bwd_ws_bytes(bwd_sched(N, B, DP, v3 != 0), DP);
    if (b > bmax) bmax = b;
    if (c > cmax) cmax = c;
  }
  if (bmax + cmax > best) best = bmax + cmax;
  return (best + 256 + 255) / 256 * 256;
}

size_t inbatch_ce_workspace_bytes(long long B, long long N, long long d) {
  return ce_ws_base(B, N, d) + ce_bwd3_ext_bytes(B);
}

template <int DP>
static int launch_ce_fwd(const CUtensorMap& tx, const TmapSet& ty, const CeFwdArgs& a, int grid, cudaStream_t st) {
  using Cfg = CeFwdCfg<DP>;
  static bool configured = false;
  if (!configured) {
    TT_CUDA(cudaFuncSetAttribute(ce_fwd_kernel<DP>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    configured = true;
  }
  KernelSpan span("ce_fwd_kernel", st);
  ce_fwd_kernel<DP><<<grid, FWD_THREADS, Cfg::SMEM_BYTES, st>>>(tx, ty, a);
  TT_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

int inbatch_ce_fwd(const void* U, long long ldu, const void* V, long long ldv, long long B, long long N, long long d,
                   long long target_offset, float* ce, float* lse, void* ws, size_t ws_bytes, cudaStream_t stream) {
  return inbatch_ce_fwd_parts(U, ldu, &V, 1, N, ldv, B, N, d, target_offset, ce, lse, ws, ws_bytes, stream);
}

int inbatch_ce_fwd_parts(const void* U, long long ldu, const void* const* Vp, int np, long long rows_per_part,
                         long long ldv, long long B, long long N, long long d, long long target_offset, float* ce,
                         float* lse, void* ws, size_t ws_bytes, cudaStream_t stream) {
  return inbatch_ce_loss_fwd(U, ldu, Vp, np, rows_per_part, ldv, B, N, d, target_offset, ce, lse, nullptr, 0, nullptr, 0,
                             nullptr, nullptr, nullptr, ws, ws_bytes, stream, nullptr);
}

int inbatch_ce_loss_fwd(const void* U, long long ldu, const void* const* Vp, int np, long long rows_per_part,
                        long long ldv, long long B, long long N, long long d, long long target_offset, float* ce,
                        float* lse, const float* labels, long long ldl, const float* uvw, long long TL, float* loss,
                        float* g, float* g_norm, void* ws, size_t ws_bytes, cudaStream_t stream, float* stats) {
  const void* V = Vp[0];
  // the attached zero fills belong to THIS call: taken (and forgotten) before anything can fail, so that an error return
  // never leaves pointers behind for a later launch
  void* zero_ptr[2];
  long long zero_bytes[2];
  for (int z = 0; z < 2; ++z) {
    zero_ptr[z] = g_zero_jobs.p[z];
    zero_bytes[z] = g_zero_jobs.n[z];
    g_zero_jobs.p[z] = nullptr;
    g_zero_jobs.n[z] = 0;
  }
  TT_CHECK(B > 0 && N > 0 && d > 0, "inbatch_ce_fwd: empty problem");
  TT_CHECK(d <= 256, "inbatch_ce_fwd: embedding dim %lld > 256 is not supported", d);
  TT_CHECK(target_offset >= 0 && target_offset + B <= N, "inbatch_ce_fwd: targets [%lld, %lld) outside the %lld item columns",
           target_offset, target_offset + B, N);
  TT_CHECK((ldu % 8) == 0 && (ldv % 8) == 0 && ((uintptr_t)U % 16) == 0 && ((uintptr_t)V % 16) == 0,
           "inbatch_ce_fwd: operands need 16-byte aligned rows");
  const int DP = pick_dp(d);
  const Sched s = make_sched(B, N, 128);
  const long long Bpad = (B + 127) / 128 * 128;
  TT_CHECK(ws_bytes >= fwd_ws_bytes(s, Bpad), "inbatch_ce_fwd: workspace too small (%zu < %zu)", ws_bytes, fwd_ws_bytes(s, Bpad));
  CeFwdArgs a;
  a.B = (int)B; a.N = (int)N; a.target_offset = target_offset;
  a.T = s.T; a.total = s.total; a.CT = s.CT; a.Bpad = Bpad;
  a.part_m = (float*)ws;
  a.part_s = a.part_m + (size_t)s.max_slots * FWD_EG * Bpad;
  a.sync = reinterpret_cast<LossSync*>(a.part_s + (size_t)s.max_slots * FWD_EG * Bpad);
  DiagOperands dgo;
  dgo.U = (const bf16*)U; dgo.ldu = ldu; dgo.ldv = ldv; dgo.target_offset = target_offset; dgo.d = (int)d;
  dgo.rows_per_part = np == 1 ? N : rows_per_part;
  for (int p = 0; p < 8; ++p) dgo.Vp[p] = (const bf16*)Vp[p < np ? p : 0];
  for (int z = 0; z < 2; ++z) {
    a.zero_ptr[z] = (uint8_t*)zero_ptr[z];
    a.zero_bytes[z] = zero_bytes[z];
  }
  a.trace = nullptr;
  if (const char* tr = getenv("TT_CE_TRACE")) a.trace = (long long*)strtoull(tr, nullptr, 0);
  a.dbg = getenv("TT_CE_DBG") ? atoi(getenv("TT_CE_DBG")) : 0;
  a.skew_ns = getenv("TT_CE_SKEW_NS") ? (unsigned)atoi(getenv("TT_CE_SKEW_NS")) : 300u;
  CUtensorMap tx;
  TmapSet ty;
  int rc = make_tmap_bf16(&tx, U, d, B, ldu, 64, 128);
  if (rc) return rc;
  rc = make_tmap_set(&ty, Vp, np, rows_per_part, N, d, ldv, 128);
  if (rc) return rc;
  if (DP == 64) rc = launch_ce_fwd<64>(tx, ty, a, s.grid, stream);
  else if (DP == 128) rc = launch_ce_fwd<128>(tx, ty, a, s.grid, stream);
  else rc = launch_ce_fwd<256>(tx, ty, a, s.grid, stream);
  if (rc) return rc;
  const long long loss_blocks = (B + COMBINE_RB - 1) / COMBINE_RB;
  if (labels != nullptr && loss_blocks <= LOSS_MAX_BLOCKS) {
    TT_CHECK(TL > 0 && ldl >= TL && uvw && loss && g && g_norm, "inbatch_ce_loss_fwd: bad label arguments");
    KernelSpan span("ce_combine_loss_kernel", stream);
    ce_combine_loss_kernel<<<(unsigned)loss_blocks, 256, 0, stream>>>((int)B, Bpad, s.T, s.CT, a.part_m, a.part_s, dgo, ce,
                                                                      lse, labels, ldl, uvw, (int)TL, 1.f / (float)B, loss, g,
                                                                      g_norm, a.sync, stats);
    TT_CUDA(cudaGetLastError());
    count_launch();
    return 0;
  }
  {
    KernelSpan span("ce_combine_kernel", stream);
    ce_combine_kernel<<<(unsigned)((B + COMBINE_RB - 1) / COMBINE_RB), 256, 0, stream>>>((int)B, Bpad, s.T, s.CT, a.part_m, a.part_s, dgo, ce, lse);
    TT_CUDA(cudaGetLastError());
    count_launch();
  }
  TT_CHECK(stats == nullptr, "inbatch_ce_loss_fwd: the per-rank statistics need B <= %d rows per rank", LOSS_MAX_BLOCKS * COMBINE_RB);
  if (labels != nullptr) {  // very large batch: separate weighted-mean launch; g is already normalised
    set_scalar_kernel<<<1, 1, 0, stream>>>(g_norm, 1.f);
    TT_CUDA(cudaGetLastError());
    count_launch();
    return weighted_loss(ce, labels, ldl, uvw, B, TL, loss, g, stream);
  }
  return 0;
}

// =============================================================================================
// Backward: host side (the kernel is ce_bwd2.cu)
// =============================================================================================
// out[row, c] = sum over the slots that touched row's tile; one launch handles up to two results (dU and dV) and
// optionally accumulates the fp32 column sums of each result (= the bias gradient of the tower Linear above it).
struct ReduceJob {
  int rows, d, CT, CTr;  // CT = column tiles per row tile including ghost tiles (SegIter), CTr = real ones
  long long T, slot_stride;
  const float* partial;
  float* out32;
  long long ld32;
  bf16* out16;
  long long ld16;
  float* colsum;  // [d] or null; caller initialises
  int blocks;
};
// Persistent blocks stride over groups of (256 / (DP/4)) rows; a thread owns 4 columns, adds the slot partials with
// 16-byte loads, writes fp32 / bf16 results with vector stores and keeps the column sums of everything it has seen in
// registers (one shared-memory fold and one atomic per column per BLOCK at the end).  Launched with a few blocks per SM:
// the first version (one block per 256 outputs, scalar stores, an atomic per column per block) was instruction bound.
__global__ void __launch_bounds__(256)
ce_bwd_reduce_kernel(const ReduceJob j0, const ReduceJob j1, int DP) {
  __shared__ float red[1024];  // [rows of the block][DP columns]
  const int cpr = DP / 4, rpb = 256 / cpr;
  const int ry = threadIdx.x / cpr, c = (threadIdx.x % cpr) * 4;
#pragma unroll 1
  for (int which = 0; which < 2; ++which) {
    const ReduceJob& j = which ? j1 : j0;
    if (j.blocks == 0) continue;
    const int T = (int)j.T, CT = j.CT, CTr = j.CTr;
    const bool vec32 = j.out32 != nullptr && (j.ld32 % 4) == 0 && c + 3 < j.d;
    const bool vec16 = j.out16 != nullptr && (j.ld16 % 4) == 0 && c + 3 < j.d;
    float4 cs = make_float4(0.f, 0.f, 0.f, 0.f);
    const int groups = (j.rows + rpb - 1) / rpb;
    for (int gi = blockIdx.x; gi < groups; gi += gridDim.x) {
      const int row = gi * rpb + ry;
      if (row >= j.rows) continue;
      const int r = row >> 7;
      const int nsl = (int)(((long long)r * CT + CTr - 1) / T) - (int)(((long long)r * CT) / T);  // last - first slot
      const float* src = j.partial + (long long)row * DP + c;
      float4 acc = *reinterpret_cast<const float4*>(src);
      for (int sl = 1; sl <= nsl; ++sl) {
        const float4 p = *reinterpret_cast<const float4*>(src + (long long)sl * j.slot_stride);
        acc.x += p.x; acc.y += p.y; acc.z += p.z; acc.w += p.w;
      }
      if (vec32) {
        *reinterpret_cast<float4*>(j.out32 + (long long)row * j.ld32 + c) = acc;
      } else if (j.out32 != nullptr) {
        const float vals[4] = {acc.x, acc.y, acc.z, acc.w};
        for (int i = 0; i < 4; ++i)
          if (c + i < j.d) j.out32[(long long)row * j.ld32 + c + i] = vals[i];
      }
      if (vec16) {
        *reinterpret_cast<uint2*>(j.out16 + (long long)row * j.ld16 + c) =
            make_uint2(pack_bf16x2(acc.x, acc.y), pack_bf16x2(acc.z, acc.w));
      } else if (j.out16 != nullptr) {
        const float vals[4] = {acc.x, acc.y, acc.z, acc.w};
        for (int i = 0; i < 4; ++i)
          if (c + i < j.d) j.out16[(long long)row * j.ld16 + c + i] = __float2bfloat16(vals[i]);
      }
      cs.x += acc.x; cs.y += acc.y; cs.z += acc.z; cs.w += acc.w;
    }
    if (j.colsum != nullptr && cpr <= 32) {  // block = rpb rows x cpr column groups
      __syncthreads();
      float* rr = red + ry * DP + c;
      rr[0] = cs.x; rr[1] = cs.y; rr[2] = cs.z; rr[3] = cs.w;
      __syncthreads();
      if ((int)threadIdx.x < DP && (int)threadIdx.x < j.d) {
        float t = 0.f;
        for (int i = 0; i < rpb; ++i) t += red[i * DP + threadIdx.x];
        atomicAdd(j.colsum + threadIdx.x, t);
      }
    }
  }
}

template <int DP>
static int ce_bwd_pass(bool colstats, const void* const* Xp, int nxp, long long x_rows_per_part, long long ldx, long long xr,
                       const void* const* Yp, int nyp, long long y_rows_per_part, long long ldy,
                       long long yr, long long d, long long diag_shift, const float* g, const float* g_scale,
                       const float* g_scale2, const float* lse, float* out32,
                       long long ld32, void* out16, long long ld16, float* colsum, void* ws, size_t ws_bytes,
                       ReduceJob& job, cudaStream_t stream, const void* ext = nullptr) {
  const bool v3 = ext != nullptr && (DP == 256 || nyp == 1 || bwd_tile_cols(DP, true) == 128);  // (96-row tiles would straddle parts)
  const int BN = bwd_tile_cols(DP, v3);
  const Sched s = bwd_sched(xr, yr, DP, v3);
  TT_CHECK(ws_bytes >= bwd_ws_bytes(s, DP), "inbatch_ce_bwd: workspace too small (%zu < %zu)", ws_bytes, bwd_ws_bytes(s, DP));
  CeBwdArgs a;
  a.XR = (int)xr; a.YR = (int)yr; a.diag_shift = diag_shift;
  a.T = s.T; a.total = s.total; a.CT = s.CT;
  a.g = g; a.g_scale = g_scale; a.g_scale2 = g_scale2; a.lse = lse;
  a.partial = (float*)ws;
  a.slot_stride = (long long)s.XT * 128 * DP;
  a.trace = nullptr;
  a.dbg = 0;
  if (const char* tr = getenv("TT_CE_TRACE")) a.trace = (long long*)strtoull(tr, nullptr, 0);
  if (const char* db = getenv("TT_CE_DBG")) a.dbg = atoi(db);
  a.cta_times = nullptr;
  if (const char* ct = getenv("TT_CE_CTA_TIMES")) a.cta_times = (unsigned long long*)strtoull(ct, nullptr, 0);
  TmapSet tx, ty;
  int rc = make_tmap_set(&tx, Xp, nxp, x_rows_per_part, xr, d, ldx, 128);
  if (rc) return rc;
  rc = make_tmap_set(&ty, Yp, nyp, y_rows_per_part, yr, d, ldy, BN);
  if (rc) return rc;
  if (v3) {  // statistics inside the score MMA (ce_bwd3.cu)
    CeBwd3Args b;
    b.XR = a.XR; b.YR = a.YR; b.diag_shift = a.diag_shift; b.T = a.T; b.total = a.total; b.CT = a.CT; b.CTr = s.CTr;
    b.g = g; b.g_scale = g_scale; b.g_scale2 = g_scale2; b.lse = lse; b.signmask = nullptr;
    b.partial = a.partial; b.slot_stride = a.slot_stride; b.trace = a.trace; b.cta_times = a.cta_times;
    b.trace_cta = getenv("TT_CE_TRACE_CTA") ? atoi(getenv("TT_CE_TRACE_CTA")) : 0;
    rc = (DP == 256 || (DP == 128 && use_x128()))
             ? launch_ce_bwd3x(DP, !colstats, tx, ty, colstats ? yr : xr, ext, b, s.grid, stream)
             : launch_ce_bwd3(DP, !colstats, tx, ty, colstats ? yr : xr, ext, b, s.grid, stream);
  } else {
    rc = launch_ce_bwd2(DP, colstats, tx, ty, a, s.grid, stream);
  }
  if (rc) return rc;
  job.rows = (int)xr; job.d = (int)d; job.CT = s.CT; job.CTr = s.CTr; job.T = s.T; job.slot_stride = a.slot_stride;
  job.partial = a.partial; job.out32 = out32; job.ld32 = ld32; job.out16 = (bf16*)out16; job.ld16 = ld16;
  job.colsum = colsum;
  job.blocks = (int)((xr * (DP / 4) + 255) / 256);
  return 0;
}

int inbatch_ce_bwd(const void* U, long long ldu, const void* V, long long ldv, long long B, long long N, long long d,
                   long long target_offset, const float* lse, const float* g, float* dU, long long lddu, void* dU16,
                   long long lddu16, float* dV, long long lddv, void* dV16, long long lddv16, float* dU_colsum,
                   float* dV_colsum, void* ws, size_t ws_bytes, cudaStream_t stream, const float* g_scale,
                   const float* g_scale2) {
  return inbatch_ce_bwd_parts(U, ldu, &V, 1, N, ldv, B, N, d, target_offset, lse, g, dU, lddu, dU16, lddu16, dV, lddv, dV16,
                              lddv16, dU_colsum, dV_colsum, ws, ws_bytes, stream, g_scale, g_scale2);
}

int inbatch_ce_bwd_parts(const void* U, long long ldu, const void* const* Vp, int np, long long rows_per_part,
                         long long ldv, long long B, long long N, long long d, long long target_offset, const float* lse,
                         const float* g, float* dU, long long lddu, void* dU16, long long lddu16, float* dV, long long lddv,
                         void* dV16, long long lddv16, float* dU_colsum, float* dV_colsum, void* ws, size_t ws_bytes,
                         cudaStream_t stream, const float* g_scale, const float* g_scale2) {
  const void* V = Vp[0];
  TT_CHECK(B > 0 && N > 0 && d > 0, "inbatch_ce_bwd: empty problem");
  TT_CHECK(d <= 256, "inbatch_ce_bwd: embedding dim %lld > 256 is not supported", d);
  TT_CHECK((ldu % 8) == 0 && (ldv % 8) == 0 && ((uintptr_t)U % 16) == 0 && ((uintptr_t)V % 16) == 0,
           "inbatch_ce_bwd: operands need 16-byte aligned rows");
  TT_CHECK(ws_bytes >= inbatch_ce_workspace_bytes(B, N, d), "inbatch_ce_bwd: workspace too small");
  const int DP = pick_dp(d);
  const bool v3 = use_bwd_v3() && (DP <= 128 || use_bwd_v3x());
  size_t offB = 0;  // region of the dV pass: behind the larger of the two possible dU geometries
  for (int k = 0; k < 2; ++k) {
    const size_t o = (bwd_ws_bytes(bwd_sched(B, N, DP, k != 0), DP) + 255) / 256 * 256;
    if (o > offB) offB = o;
  }
  ReduceJob ja, jb;
  ja.blocks = jb.blocks = 0;
  ja.rows = jb.rows = 0;
  int rc = 0;
  const bool wantU = dU || dU16, wantV = dV || dV16;
  // v3 kernels: one small launch turns (g, lse) into the users' bias rows + sign words for the dV pass
  void* ext = nullptr;
  if (v3 && (wantU || wantV)) {
    ext = (char*)ws + ce_ws_base(B, N, d);
    TT_CHECK(ws_bytes >= ce_ws_base(B, N, d) + ce_bwd3_ext_bytes(B), "inbatch_ce_bwd: workspace too small for the ext block");
    if (wantV) {
      rc = ce_bwd3_prep(B, g, lse, ext, stream);
      if (rc) return rc;
    }
  }
#define TT_PASS(DPV)                                                                                                 \
  do {                                                                                                               \
    if (wantU)                                                                                                       \
      rc = ce_bwd_pass<DPV>(false, &U, 1, B, ldu, B, Vp, np, rows_per_part, ldv, N, d, target_offset, g, g_scale, g_scale2, lse, dU, lddu, dU16, \
                            lddu16, dU_colsum, ws, offB, ja, stream, ext);                                           \
    if (rc == 0 && wantV)                                                                                            \
      rc = ce_bwd_pass<DPV>(true, Vp, np, rows_per_part, ldv, N, &U, 1, B, ldu, B, d, -target_offset, g, g_scale, g_scale2, lse, dV, lddv, dV16, \
                            lddv16, dV_colsum, (char*)ws + offB, ws_bytes - offB, jb, stream, ext);                  \
  } while (0)
  if (DP == 64) TT_PASS(64);
  else if (DP == 128) TT_PASS(128);
  else TT_PASS(256);
#undef TT_PASS
  if (rc) return rc;
  if (!wantU) { ja = jb; jb.blocks = 0; }
  if (ja.blocks + jb.blocks == 0) return 0;
  if (DP > 128) { ja.colsum = nullptr; jb.colsum = nullptr; }  // handled by the caller (tt_colsum) for wide rows
  KernelSpan span("ce_bwd_reduce_kernel", stream);
  const int want = ja.blocks + jb.blocks, cap = 4 * num_sms();
  ce_bwd_reduce_kernel<<<(unsigned)(want < cap ? want : cap), 256, 0, stream>>>(ja, jb, DP);
  TT_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

}  // namespace tt
