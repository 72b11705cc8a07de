// Fused in-batch sampled-softmax cross entropy for sm_100a (forward and backward).
//
// Reference semantics: src/two_tower_base_retrieval.py:287 (S = U V^T), :301 (target = arange),
// :310-312 (F.cross_entropy(reduction="none")) and autograd of the same (train/train.py:124).
// The [B,N] score matrix never exists in HBM: 128xBN score tiles are produced by tcgen05.mma into
// TMEM from TMA-staged bf16 operand tiles and consumed in place.
//
//   forward : per 128-row tile, running (max, sum-exp) over column tiles + the target logit.
//   backward: one generic "flash" kernel  acc[128, d] = sum_j E_j * Y_j,  E_j = f(X Y_j^T), run twice:
//       pass A  X=U, Y=V, E_ij = g_i (exp(S_ij - lse_i) - [j == i+off])           -> dU
//       pass B  X=V, Y=U, E_ji = g_i (exp(S_ij - lse_i) - [j == i+off]) (col stats) -> dV
//     E is written as bf16 into a 128B-swizzled smem tile and fed back as the A operand of the second
//     UMMA, whose B operand is the same Y tile read MN-major.
//
// Work is the flattened (row tile, column tile) space cut into equal contiguous ranges, one per SM
// (persistent CTAs).  A CTA's range touches <= a few row tiles ("segments"); every segment writes a
// partial result into a slot and a small second kernel merges the slots.
//
// Warp roles (384 threads): warp 0 TMA producer, warp 1 UMMA issuer, warp 2 TMEM allocator,
// warps 4-7 / 8-11 two epilogue warpgroups that alternate score tiles (TMEM buffer e <-> group e).
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
  float* diag;
  long long* trace;  // bring-up (TT_CE_TRACE)
};

template <int DP>
struct CeFwdCfg {
  static constexpr int BN = 128;
  static constexpr int NS = 4;  // score-tile buffers in TMEM (4 x 128 columns)
  static constexpr int KBOX = DP / 64;
  static constexpr int X_BYTES = 128 * DP * 2;
  static constexpr int Y_BYTES = BN * DP * 2;
  static constexpr int STAGES = DP == 64 ? 6 : (DP == 128 ? 4 : 2);
  static constexpr int SMEM_BYTES = X_BYTES + STAGES * Y_BYTES + 1024 + 256;
};

// Both epilogue groups work on EVERY score tile, each on one half of its columns (group e: columns
// [64 e, 64 e + 64)), so all 8 epilogue warps are busy on the tile that is ready while the UMMA warp runs up
// to NS - 1 tiles ahead.  A row's running (max, sum-exp) is therefore split over two threads; the combine
// kernel merges the (slot, group) partials.
template <int DP>
__global__ void __launch_bounds__(384, 1)
ce_fwd_kernel(const __grid_constant__ CUtensorMap tmx, const __grid_constant__ TmapSet tmy, const CeFwdArgs a) {
  using Cfg = CeFwdCfg<DP>;
  constexpr int BN = Cfg::BN, NS = Cfg::NS;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sx = smem;
  uint8_t* sy = smem + Cfg::X_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sy + Cfg::STAGES * Cfg::Y_BYTES);
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
  if (warp == 1 && lane == 0) {
    mbar_init(x_full, 1);
    mbar_init(x_empty, 1);
    for (int i = 0; i < Cfg::STAGES; ++i) {
      mbar_init(&y_full[i], 1);
      mbar_init(&y_empty[i], 1);
    }
    for (int i = 0; i < NS; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&s_empty[i], 8);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_holder, NS * BN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;

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
          if (a.trace && blockIdx.x == 0 && t < 64 && leader) a.trace[(0 * 64 + t) * 2 + 0] = clock64();
          mbar_wait(&y_full[stage], phase);
          if (a.trace && blockIdx.x == 0 && t < 64 && leader) a.trace[(1 * 64 + t) * 2 + 0] = clock64();
          mbar_wait(&s_empty[buf], (use & 1) ^ 1);
          tc_fence_after();
          if (a.trace && blockIdx.x == 0 && t < 64 && leader) a.trace[(0 * 64 + t) * 2 + 1] = clock64();
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
  } else if (warp >= 4) {
    const int e = (warp - 4) >> 2;  // column half of every tile
    const int q = warp & 3;         // TMEM lane quarter
    SegIter it(a.T, a.total, a.CT);
    int r, j0, j1;
    uint32_t t = 0;
    while (it.next(r, j0, j1)) {
      const long long row = (long long)r * 128 + q * 32 + lane;
      const bool valid = row < a.B;
      const long long tgt = row + a.target_offset;
      float m = -INFINITY, s = 0.f;
      for (int j = j0; j < j1; ++j, ++t) {
        const uint32_t buf = t % NS, use = t / NS;
        mbar_wait(&s_full[buf], use & 1);
        tc_fence_after();
        if (a.trace && blockIdx.x == 0 && t < 64 && q == 0 && lane == 0) a.trace[((2 + e) * 64 + t) * 2 + 0] = clock64();
        // online (max, sum-exp) update with the 32 scores of tile columns [c*32, c*32+32)
        auto consume = [&](float* v, int c) {
          const long long n0 = (long long)j * BN + c * 32;
          if (n0 >= a.N) return;
          if (n0 + 32 > a.N) {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (n0 + i >= a.N) v[i] = -INFINITY;
          }
          if (valid && tgt >= n0 && tgt < n0 + 32) {
            float dg = 0.f;
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (n0 + i == tgt) dg = v[i];
            a.diag[row] = dg;
          }
          // chunk maximum as a tree (4 independent chains), then one rescale of the running sum
          float c0 = fmaxf(v[0], v[1]), c1 = fmaxf(v[2], v[3]), c2 = fmaxf(v[4], v[5]), c3 = fmaxf(v[6], v[7]);
#pragma unroll
          for (int i = 8; i < 32; i += 4) {
            c0 = fmaxf(c0, v[i]); c1 = fmaxf(c1, v[i + 1]); c2 = fmaxf(c2, v[i + 2]); c3 = fmaxf(c3, v[i + 3]);
          }
          const float cm = fmaxf(fmaxf(c0, c1), fmaxf(c2, c3));
          const float m_new = fmaxf(m, cm * LOG2E);
          const float ms = (m_new == -INFINITY) ? 0.f : m_new;  // fully masked so far: avoid inf - inf
          s *= ex2f(m - ms);
          float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            acc0 += ex2f(fmaf(v[i], LOG2E, -ms));
            acc1 += ex2f(fmaf(v[i + 1], LOG2E, -ms));
            acc2 += ex2f(fmaf(v[i + 2], LOG2E, -ms));
            acc3 += ex2f(fmaf(v[i + 3], LOG2E, -ms));
          }
          s += (acc0 + acc1) + (acc2 + acc3);
          m = m_new;
        };
        static_assert(BN == 128, "two 32-column chunks per epilogue group");
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * BN + e * 64;
        float v0[32], v1[32];
        tmem_ld32(taddr, v0);
        tmem_wait_ld();
        tmem_ld32(taddr + 32, v1);  // in flight while the first chunk is reduced
        consume(v0, e * 2);
        tmem_wait_ld();
        consume(v1, e * 2 + 1);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_empty[buf]);
        if (a.trace && blockIdx.x == 0 && t < 64 && q == 0 && lane == 0) a.trace[((2 + e) * 64 + t) * 2 + 1] = clock64();
      }
      if (valid) {
        const int slot = (int)(blockIdx.x - ((long long)r * a.CT) / a.T);
        const long long o = (long long)(slot * 2 + e) * a.Bpad + row;
        a.part_m[o] = m;
        a.part_s[o] = s;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, NS * BN);
  }
}

__global__ void ce_combine_kernel(int B, long long Bpad, long long T, int CT, const float* part_m, const float* part_s,
                                  const float* diag, float* ce, float* lse) {
  const long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= B) return;
  const long long r = row / 128;
  const int first = (int)((r * CT) / T), last = (int)(((r + 1) * CT - 1) / T);
  float M = -INFINITY;
  for (int sl = 0; sl <= last - first; ++sl)
    for (int e = 0; e < 2; ++e) M = fmaxf(M, part_m[(long long)(sl * 2 + e) * Bpad + row]);
  float S = 0.f;
  for (int sl = 0; sl <= last - first; ++sl)
    for (int e = 0; e < 2; ++e) {
      const long long o = (long long)(sl * 2 + e) * Bpad + row;
      S += part_s[o] * exp2f(part_m[o] - M);
    }
  const float l = (M + log2f(S)) * LN2;
  lse[row] = l;
  ce[row] = l - diag[row];
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
  return (size_t)(2 * (size_t)s.max_slots * 2 * Bpad + Bpad) * sizeof(float);
}
static size_t bwd_ws_bytes(const Sched& s, int DP) { return (size_t)s.max_slots * s.XT * 128 * DP * sizeof(float); }

size_t inbatch_ce_workspace_bytes(long long B, long long N, long long d) {
  const int DP = pick_dp(d);
  const int BNb = DP == 256 ? 64 : 128;
  const long long Bpad = (B + 127) / 128 * 128;
  size_t a = fwd_ws_bytes(make_sched(B, N, 128), Bpad);
  size_t b = bwd_ws_bytes(make_sched(B, N, BNb), DP);
  size_t c = bwd_ws_bytes(make_sched(N, B, BNb), DP);
  const size_t bw = (b + 255) / 256 * 256 + c;  // the two backward passes use disjoint regions
  return (a > bw ? a : bw) + 256;
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
  ce_fwd_kernel<DP><<<grid, 384, Cfg::SMEM_BYTES, st>>>(tx, ty, a);
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
  const void* V = Vp[0];
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
  a.part_s = a.part_m + (size_t)s.max_slots * 2 * Bpad;
  a.diag = a.part_s + (size_t)s.max_slots * 2 * Bpad;
  a.trace = nullptr;
  if (const char* tr = getenv("TT_CE_TRACE")) a.trace = (long long*)strtoull(tr, nullptr, 0);
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
  KernelSpan span("ce_combine_kernel", stream);
  ce_combine_kernel<<<(unsigned)((B + 255) / 256), 256, 0, stream>>>((int)B, Bpad, s.T, s.CT, a.part_m, a.part_s, a.diag, ce, lse);
  TT_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

// =============================================================================================
// Backward: host side (the kernel is ce_bwd2.cu)
// =============================================================================================
// out[row, c] = sum over the slots that touched row's tile; one launch handles up to two results (dU and dV) and
// optionally accumulates the fp32 column sums of each result (= the bias gradient of the tower Linear above it).
struct ReduceJob {
  int rows, d, CT;
  long long T, slot_stride;
  const float* partial;
  float* out32;
  long long ld32;
  bf16* out16;
  long long ld16;
  float* colsum;  // [d] or null; caller initialises
  int blocks;
};
__global__ void __launch_bounds__(256)
ce_bwd_reduce_kernel(const ReduceJob j0, const ReduceJob j1, int DP) {
  __shared__ float red[1024];  // [rows of the block][DP columns]
  const bool second = (int)blockIdx.x >= j0.blocks;
  const ReduceJob& j = second ? j1 : j0;
  const long long idx = (long long)(blockIdx.x - (second ? j0.blocks : 0)) * blockDim.x + threadIdx.x;
  const int cpr = DP / 4;
  const long long row = idx / cpr;
  const int c = (int)(idx % cpr) * 4;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (row < j.rows) {
    const long long r = row / 128;
    const int first = (int)((r * j.CT) / j.T), last = (int)(((r + 1) * j.CT - 1) / j.T);
    for (int sl = 0; sl <= last - first; ++sl) {
      const float4 p = *reinterpret_cast<const float4*>(j.partial + (long long)sl * j.slot_stride + row * DP + c);
      acc.x += p.x; acc.y += p.y; acc.z += p.z; acc.w += p.w;
    }
    const float vals[4] = {acc.x, acc.y, acc.z, acc.w};
    for (int i = 0; i < 4; ++i) {
      if (c + i < j.d) {
        if (j.out32) j.out32[row * j.ld32 + c + i] = vals[i];
        if (j.out16) j.out16[row * j.ld16 + c + i] = __float2bfloat16(vals[i]);
      }
    }
  }
  if (j.colsum != nullptr && cpr <= 32) {  // block = (256 / cpr) rows x cpr column groups
    const int rpb = 256 / cpr, ry = threadIdx.x / cpr, cx = threadIdx.x % cpr;
    float* rr = red + ry * DP + cx * 4;
    rr[0] = acc.x; rr[1] = acc.y; rr[2] = acc.z; rr[3] = acc.w;
    __syncthreads();
    if (threadIdx.x < DP && threadIdx.x < j.d) {
      float t = 0.f;
      for (int i = 0; i < rpb; ++i) t += red[i * DP + threadIdx.x];
      atomicAdd(j.colsum + threadIdx.x, t);
    }
  }
}

template <int DP>
static int ce_bwd_pass(bool colstats, const void* const* Xp, int nxp, long long x_rows_per_part, long long ldx, long long xr,
                       const void* const* Yp, int nyp, long long y_rows_per_part, long long ldy,
                       long long yr, long long d, long long diag_shift, const float* g, const float* lse, float* out32,
                       long long ld32, void* out16, long long ld16, float* colsum, void* ws, size_t ws_bytes,
                       ReduceJob& job, cudaStream_t stream) {
  constexpr int BN = DP == 256 ? 64 : 128;
  const Sched s = make_sched(xr, yr, BN);
  TT_CHECK(ws_bytes >= bwd_ws_bytes(s, DP), "inbatch_ce_bwd: workspace too small (%zu < %zu)", ws_bytes, bwd_ws_bytes(s, DP));
  CeBwdArgs a;
  a.XR = (int)xr; a.YR = (int)yr; a.diag_shift = diag_shift;
  a.T = s.T; a.total = s.total; a.CT = s.CT;
  a.g = g; a.lse = lse;
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
  rc = launch_ce_bwd2(DP, colstats, tx, ty, a, s.grid, stream);
  if (rc) return rc;
  job.rows = (int)xr; job.d = (int)d; job.CT = s.CT; job.T = s.T; job.slot_stride = a.slot_stride;
  job.partial = a.partial; job.out32 = out32; job.ld32 = ld32; job.out16 = (bf16*)out16; job.ld16 = ld16;
  job.colsum = colsum;
  job.blocks = (int)((xr * (DP / 4) + 255) / 256);
  return 0;
}

int inbatch_ce_bwd(const void* U, long long ldu, const void* V, long long ldv, long long B, long long N, long long d,
                   long long target_offset, const float* lse, const float* g, float* dU, long long lddu, void* dU16,
                   long long lddu16, float* dV, long long lddv, void* dV16, long long lddv16, float* dU_colsum,
                   float* dV_colsum, void* ws, size_t ws_bytes, cudaStream_t stream) {
  return inbatch_ce_bwd_parts(U, ldu, &V, 1, N, ldv, B, N, d, target_offset, lse, g, dU, lddu, dU16, lddu16, dV, lddv, dV16,
                              lddv16, dU_colsum, dV_colsum, ws, ws_bytes, stream);
}

int inbatch_ce_bwd_parts(const void* U, long long ldu, const void* const* Vp, int np, long long rows_per_part,
                         long long ldv, long long B, long long N, long long d, long long target_offset, const float* lse,
                         const float* g, float* dU, long long lddu, void* dU16, long long lddu16, float* dV, long long lddv,
                         void* dV16, long long lddv16, float* dU_colsum, float* dV_colsum, void* ws, size_t ws_bytes,
                         cudaStream_t stream) {
  const void* V = Vp[0];
  TT_CHECK(B > 0 && N > 0 && d > 0, "inbatch_ce_bwd: empty problem");
  TT_CHECK(d <= 256, "inbatch_ce_bwd: embedding dim %lld > 256 is not supported", d);
  TT_CHECK((ldu % 8) == 0 && (ldv % 8) == 0 && ((uintptr_t)U % 16) == 0 && ((uintptr_t)V % 16) == 0,
           "inbatch_ce_bwd: operands need 16-byte aligned rows");
  TT_CHECK(ws_bytes >= inbatch_ce_workspace_bytes(B, N, d) - 256, "inbatch_ce_bwd: workspace too small");
  const int DP = pick_dp(d);
  const int BNb = DP == 256 ? 64 : 128;
  const size_t offB = (bwd_ws_bytes(make_sched(B, N, BNb), DP) + 255) / 256 * 256;  // region of the dV pass
  ReduceJob ja, jb;
  ja.blocks = jb.blocks = 0;
  ja.rows = jb.rows = 0;
  int rc = 0;
  const bool wantU = dU || dU16, wantV = dV || dV16;
#define TT_PASS(DPV)                                                                                                 \
  do {                                                                                                               \
    if (wantU)                                                                                                       \
      rc = ce_bwd_pass<DPV>(false, &U, 1, B, ldu, B, Vp, np, rows_per_part, ldv, N, d, target_offset, g, lse, dU, lddu, dU16, \
                            lddu16, dU_colsum, ws, offB, ja, stream);                                                \
    if (rc == 0 && wantV)                                                                                            \
      rc = ce_bwd_pass<DPV>(true, Vp, np, rows_per_part, ldv, N, &U, 1, B, ldu, B, d, -target_offset, g, lse, dV, lddv, dV16, \
                            lddv16, dV_colsum, (char*)ws + offB, ws_bytes - offB, jb, stream);                       \
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
  ce_bwd_reduce_kernel<<<(unsigned)(ja.blocks + jb.blocks), 256, 0, stream>>>(ja, jb, DP);
  TT_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

}  // namespace tt
