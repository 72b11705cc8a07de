// Persistent warp-specialised tcgen05 GEMM for sm_100a:  C[M,N] (+)= alpha * A * B^T  (+bias, ReLU, mask)
//
// A is [M,K] (K-major) or, when a_mn, stored as [K,M] (M contiguous: "MN-major", used for X^T * dY weight
// gradients without materialising a transpose); likewise B is [N,K] or stored [K,N].  bf16 in, fp32
// accumulate in TMEM, fp32 and/or bf16 out.  One CTA per SM:
//   warp 0  TMA producer (one lane)     warp 1  UMMA issuer (one lane)     warp 2  TMEM allocator
//   warps 4-11 epilogue (TMEM -> registers -> global; two groups of 4 warps, each draining one column half of
//   the tile), overlapped with the next tile's main loop through two TMEM accumulator stages.
// Used by the tower MLPs (reference src/two_tower_base_retrieval.py:76-80,90-93,101-110), their
// backward (autograd of the same) and the history encoder's in/out projections
// (src/user_history_encoder.py:60-67).
#include <stdlib.h>

#include "common.cuh"
#include "kernels.h"

namespace tt {

static constexpr int BM = 128;
static constexpr int BK = 64;  // one 128-byte swizzle atom of bf16 along K

struct GemmArgs {
  int M, N, K;
  int a_mn, b_mn;
  int m_tiles, n_tiles, splits, kb_total, kb_per_split;
  const float* bias;
  int relu;
  const bf16* mask;
  long long ld_mask;
  float* c32;
  long long ldc32;
  int atomic32;
  bf16* c16;
  long long ldc16;
  float alpha;
  int vec32, vec16, vecmask;
  int mn_lbo, mn_sbo, mn_kadv;  // MN-major descriptor strides (bytes)
  float* colsum;                 // optional [N] fp32, += column sums of the final C values
  int dbg;                       // bring-up (TT_GEMM_DBG): bit0 skip global stores, bit1 skip bias shuffles
};

// Up to MAXP independent problems per launch (e.g. the same layer of the user tower and of the item tower):
// the work items of all problems form one list that the persistent CTAs stride over, so two half-empty
// launches become one launch that fills the SMs, and a dependent chain of small GEMMs is half as long.
static constexpr int MAXP = 4;
struct GemmBatch {
  CUtensorMap ta[MAXP], tb[MAXP];
  GemmArgs g[MAXP];
  int n;
  int work_start[MAXP + 1];
};

template <int BN>
struct GemmCfg {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (BN == 256) ? 4 : 6;
  static constexpr int TMEM_COLS = (2 * BN <= 32) ? 32 : (2 * BN <= 64 ? 64 : (2 * BN <= 128 ? 128 : (2 * BN <= 256 ? 256 : 512)));
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
};

template <int BN>
__global__ void __launch_bounds__(384, 1)
gemm_kernel(const __grid_constant__ GemmBatch bt) {
  using Cfg = GemmCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + Cfg::STAGES;
  uint64_t* tfull_bar = empty_bar + Cfg::STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    for (int p = 0; p < bt.n; ++p) {
      tma_prefetch_desc(&bt.ta[p]);
      tma_prefetch_desc(&bt.tb[p]);
    }
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < Cfg::STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 8);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_holder, Cfg::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  const int total_work = bt.work_start[bt.n];

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
        int p = 0;
        while (w >= bt.work_start[p + 1]) ++p;
        const GemmArgs& g = bt.g[p];
        const CUtensorMap* tma = &bt.ta[p];
        const CUtensorMap* tmb = &bt.tb[p];
        const int lw = w - bt.work_start[p];
        const int split = lw % g.splits;
        const int t = lw / g.splits;
        const int n_tile = t % g.n_tiles, m_tile = t / g.n_tiles;
        const int kb0 = split * g.kb_per_split;
        const int kb1 = min(g.kb_total, kb0 + g.kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
          uint8_t* sb = sa + Cfg::A_BYTES;
          mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
          if (!g.a_mn) {
            tma_load_2d(sa, tma, &full_bar[stage], kb * BK, m_tile * BM);
          } else {
#pragma unroll
            for (int b = 0; b < BM / 64; ++b)
              tma_load_2d(sa + b * (BK * 128), tma, &full_bar[stage], m_tile * BM + b * 64, kb * BK);
          }
          if (!g.b_mn) {
            tma_load_2d(sb, tmb, &full_bar[stage], kb * BK, n_tile * BN);
          } else {
#pragma unroll
            for (int b = 0; b < BN / 64; ++b)
              tma_load_2d(sb + b * (BK * 128), tmb, &full_bar[stage], n_tile * BN + b * 64, kb * BK);
          }
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    {  // the whole warp walks the schedule; only the elected lane issues tcgen05 instructions
      const uint32_t leader = elect_one();
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      const uint32_t s0 = smem_u32(smem);
      for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
        int p = 0;
        while (w >= bt.work_start[p + 1]) ++p;
        const GemmArgs& g = bt.g[p];
        const uint32_t idesc = make_idesc_bf16(BM, BN, g.a_mn, g.b_mn);
        const uint64_t da0 = g.a_mn ? make_smem_desc_sw128(s0, g.mn_lbo, g.mn_sbo) : make_smem_desc_sw128(s0, 0, 1024);
        const uint64_t db0 = g.b_mn ? make_smem_desc_sw128(s0 + Cfg::A_BYTES, g.mn_lbo, g.mn_sbo)
                                    : make_smem_desc_sw128(s0 + Cfg::A_BYTES, 0, 1024);
        const uint32_t ka = g.a_mn ? g.mn_kadv : 32, kb_ = g.b_mn ? g.mn_kadv : 32;
        const int lw = w - bt.work_start[p];
        const int split = lw % g.splits;
        const int kb0 = split * g.kb_per_split;
        const int kb1 = min(g.kb_total, kb0 + g.kb_per_split);
        mbar_wait(&tempty_bar[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint64_t da = desc_advance(da0, stage * Cfg::STAGE_BYTES);
          const uint64_t db = desc_advance(db0, stage * Cfg::STAGE_BYTES);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)
            umma_bf16_w(d_tmem, desc_advance(da, k * ka), desc_advance(db, k * kb_), idesc, (kb > kb0 || k > 0) ? 1u : 0u,
                        leader);
          umma_commit_w(&empty_bar[stage], leader);
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit_w(&tfull_bar[as], leader);
        if (++as == 2) { as = 0; aphase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    const int q = warp & 3;         // TMEM lane quarter owned by this warp
    const int e = (warp - 4) >> 2;  // epilogue group: column half of every tile (8 warps drain one accumulator)
    int as = 0;
    uint32_t aphase = 0;
    for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
      int p = 0;
      while (w >= bt.work_start[p + 1]) ++p;
      const GemmArgs g = bt.g[p];  // registers: the chunk loop below must not re-read kernel parameters
      const int t = (w - bt.work_start[p]) / g.splits;
      const int n_tile = t % g.n_tiles, m_tile = t / g.n_tiles;
      mbar_wait(&tfull_bar[as], aphase);
      tc_fence_after();
      const long long row = (long long)m_tile * BM + q * 32 + lane;
      const bool row_ok = row < g.M;
#pragma unroll 1
      for (int cc = 0; cc < BN / 64; ++cc) {
        const int c = e * (BN / 64) + cc;
        const int n0 = n_tile * BN + c * 32;
        if (n0 >= g.N) break;
        float v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + as * BN + c * 32, v);
        // operands of the epilogue are fetched while the TMEM load is in flight
        const bool full = (n0 + 32 <= g.N);
        // bias of the chunk: 8 independent broadcast 16-byte loads (a per-element __ldg + add gets serialised by
        // ptxas, a shuffle broadcast costs as much crossbar time as the stores)
        float bias_r[32];
        const bool vbias = g.bias != nullptr && full && ((reinterpret_cast<uintptr_t>(g.bias) & 15) == 0);
        if (vbias) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(g.bias + n0) + j);
            bias_r[4 * j] = b4.x; bias_r[4 * j + 1] = b4.y; bias_r[4 * j + 2] = b4.z; bias_r[4 * j + 3] = b4.w;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) bias_r[j] = (g.bias != nullptr && n0 + j < g.N) ? __ldg(g.bias + n0 + j) : 0.f;
        }
        uint4 mv[4];
        const bool vmask = g.mask != nullptr && row_ok && full && g.vecmask;
        if (vmask) {
          const bf16* mp = g.mask + row * g.ld_mask + n0;
#pragma unroll
          for (int j = 0; j < 4; ++j) mv[j] = *reinterpret_cast<const uint4*>(mp + j * 8);
        }
        tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          float x = fmaf(v[j], g.alpha, bias_r[j]);
          if (g.relu) x = fmaxf(x, 0.f);
          v[j] = x;
        }
        if (g.mask != nullptr && row_ok) {
          if (vmask) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const bf16* mb = reinterpret_cast<const bf16*>(&mv[j]);
#pragma unroll
              for (int e = 0; e < 8; ++e)
                if (!(__bfloat162float(mb[e]) > 0.f)) v[j * 8 + e] = 0.f;
            }
          } else {
            const bf16* mp = g.mask + row * g.ld_mask + n0;
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (n0 + j < g.N && !(__bfloat162float(mp[j]) > 0.f)) v[j] = 0.f;
          }
        }
        if (row_ok && !(g.dbg & 1)) {
          if (g.c32 != nullptr) {
            float* cp = g.c32 + row * g.ldc32 + n0;
            if (g.atomic32 && full && g.vec32) {
#pragma unroll
              for (int j = 0; j < 8; ++j)  // 16-byte vector reductions: 4x fewer L2 atomic transactions
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(cp + j * 4), "f"(v[4 * j]),
                             "f"(v[4 * j + 1]), "f"(v[4 * j + 2]), "f"(v[4 * j + 3])
                             : "memory");
            } else if (g.atomic32) {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (n0 + j < g.N) atomicAdd(cp + j, v[j]);
            } else if (full && g.vec32 == 2) {
#pragma unroll
              for (int j = 0; j < 4; ++j)
                asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(cp + j * 8), "f"(v[8 * j]),
                             "f"(v[8 * j + 1]), "f"(v[8 * j + 2]), "f"(v[8 * j + 3]), "f"(v[8 * j + 4]), "f"(v[8 * j + 5]),
                             "f"(v[8 * j + 6]), "f"(v[8 * j + 7])
                             : "memory");
            } else if (full && g.vec32) {
#pragma unroll
              for (int j = 0; j < 8; ++j)
                *reinterpret_cast<float4*>(cp + j * 4) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (n0 + j < g.N) cp[j] = v[j];
            }
          }
          if (g.c16 != nullptr) {
            bf16* cp = g.c16 + row * g.ldc16 + n0;
            if (full && g.vec16 == 2) {
#pragma unroll
              for (int j = 0; j < 2; ++j) {
                uint32_t o[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) o[i] = pack_bf16x2(v[16 * j + 2 * i], v[16 * j + 2 * i + 1]);
                asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(cp + j * 16), "r"(o[0]),
                             "r"(o[1]), "r"(o[2]), "r"(o[3]), "r"(o[4]), "r"(o[5]), "r"(o[6]), "r"(o[7])
                             : "memory");
              }
            } else if (full && g.vec16) {
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                uint4 o;
                o.x = pack_bf16x2(v[8 * j + 0], v[8 * j + 1]);
                o.y = pack_bf16x2(v[8 * j + 2], v[8 * j + 3]);
                o.z = pack_bf16x2(v[8 * j + 4], v[8 * j + 5]);
                o.w = pack_bf16x2(v[8 * j + 6], v[8 * j + 7]);
                *reinterpret_cast<uint4*>(cp + j * 8) = o;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (n0 + j < g.N) cp[j] = __float2bfloat16(v[j]);
            }
          }
        }
        if (g.colsum != nullptr) {
          // fp32 column sums of the tile (bias gradients): butterfly transpose-reduce over the 32 rows held
          // by this warp, after which lane l owns column n0 + l; one atomic per column per warp.
          if (!row_ok) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = 0.f;
          }
#pragma unroll
          for (int off = 16; off >= 1; off >>= 1) {
            const bool up = (lane & off) != 0;
#pragma unroll
            for (int i = 0; i < off; ++i) {
              const float send = up ? v[i] : v[i + off];
              const float keep = up ? v[i + off] : v[i];
              v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
            }
          }
          if (n0 + lane < g.N) atomicAdd(g.colsum + n0 + lane, v[0]);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[as]);
      if (++as == 2) { as = 0; aphase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

template <int BN>
static int launch_gemm(const GemmBatch& bt, cudaStream_t stream) {
  using Cfg = GemmCfg<BN>;
  static bool configured = false;
  if (!configured) {
    TT_CUDA(cudaFuncSetAttribute(gemm_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    configured = true;
  }
  const int total = bt.work_start[bt.n];
  const int grid = total < num_sms() ? total : num_sms();
  KernelSpan span(bt.g[0].atomic32 ? "gemm_splitk" : "gemm", stream);
  gemm_kernel<BN><<<grid, 384, Cfg::SMEM_BYTES, stream>>>(bt);
  TT_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

static int pick_bn(const GemmDesc& d, int share = 1) {
  // fewest padded columns first; then the widest tile that still gives ~one CTA per SM, else the narrowest
  int BN = 64;
  const long long m_tiles = (d.M + BM - 1) / BM;
  long long best_pad = -1;
  const int cand[3] = {256, 128, 64};
  for (int i = 0; i < 3; ++i) {
    const long long padded = (d.N + cand[i] - 1) / cand[i] * cand[i];
    if (best_pad < 0 || padded < best_pad) best_pad = padded;
  }
  bool found = false;
  for (int i = 0; i < 3 && !found; ++i) {
    const long long padded = (d.N + cand[i] - 1) / cand[i] * cand[i];
    if (padded != best_pad) continue;
    BN = cand[i];  // ends at the narrowest candidate with minimal padding
    if (m_tiles * (padded / cand[i]) * share * 4 >= 3ll * num_sms()) found = true;  // share = problems in the launch
  }
  return BN;
}

// validate one problem and fill its kernel arguments / tensor maps; `share` = problems in the same launch
static int prepare(const GemmDesc& d, int BN, int share, GemmArgs& g, CUtensorMap& ta, CUtensorMap& tb) {
  TT_CHECK(d.M > 0 && d.N > 0 && d.K > 0, "gemm: empty problem M=%lld N=%lld K=%lld", d.M, d.N, d.K);
  TT_CHECK(d.A && d.B, "gemm: null operand");
  TT_CHECK((d.lda % 8) == 0 && (d.ldb % 8) == 0, "gemm: operand pitches must be multiples of 8 elements (lda=%lld ldb=%lld)", d.lda, d.ldb);
  TT_CHECK(((uintptr_t)d.A % 16) == 0 && ((uintptr_t)d.B % 16) == 0, "gemm: operands must be 16-byte aligned");
  TT_CHECK(d.c32 || d.c16 || d.colsum, "gemm: no output");
  TT_CHECK(!(d.accumulate && (d.bias || d.relu || d.relu_mask || d.c16 || d.colsum)),
           "gemm: split-K accumulation only supports a plain fp32 output");
  g.M = (int)d.M; g.N = (int)d.N; g.K = (int)d.K;
  g.a_mn = d.a_mn_major; g.b_mn = d.b_mn_major;
  g.m_tiles = (int)((d.M + BM - 1) / BM);
  g.n_tiles = (int)((d.N + BN - 1) / BN);
  g.kb_total = (int)((d.K + BK - 1) / BK);
  int splits = 1;
  if (d.accumulate) {
    splits = d.split_k;
    if (splits <= 0) {
      const int tiles = g.m_tiles * g.n_tiles * share;
      splits = tiles >= num_sms() ? 1 : (num_sms() + tiles - 1) / tiles;
      const int max_useful = g.kb_total / 8 > 0 ? g.kb_total / 8 : 1;  // >= 8 k-blocks per CTA: fewer atomics
      if (splits > max_useful) splits = max_useful;
    }
    if (splits > g.kb_total) splits = g.kb_total;
  }
  g.kb_per_split = (g.kb_total + splits - 1) / splits;
  g.splits = (g.kb_total + g.kb_per_split - 1) / g.kb_per_split;
  g.bias = d.bias; g.relu = d.relu;
  g.mask = (const bf16*)d.relu_mask; g.ld_mask = d.ld_mask;
  g.c32 = d.c32; g.ldc32 = d.ldc32; g.atomic32 = d.accumulate ? 1 : 0;
  g.c16 = (bf16*)d.c16; g.ldc16 = d.ldc16;
  g.alpha = d.alpha;
  g.colsum = d.colsum;
  g.vec32 = d.c32 && (d.ldc32 % 4 == 0) && ((uintptr_t)d.c32 % 16 == 0);
  g.vec16 = d.c16 && (d.ldc16 % 8 == 0) && ((uintptr_t)d.c16 % 16 == 0);
  // 32-byte stores (one full L2 sector per instruction instead of two half-sector writes)
  if (g.vec16 && (d.ldc16 % 16 == 0) && ((uintptr_t)d.c16 % 32 == 0)) g.vec16 = 2;
  if (g.vec32 && (d.ldc32 % 8 == 0) && ((uintptr_t)d.c32 % 32 == 0)) g.vec32 = 2;
  g.vecmask = d.relu_mask && (d.ld_mask % 8 == 0) && ((uintptr_t)d.relu_mask % 16 == 0);
  g.mn_lbo = BK * 128; g.mn_sbo = 1024; g.mn_kadv = 2048;
  g.dbg = getenv("TT_GEMM_DBG") ? atoi(getenv("TT_GEMM_DBG")) : 0;
  int rc;
  if (!d.a_mn_major) rc = make_tmap_bf16(&ta, d.A, d.K, d.M, d.lda, 64, BM);
  else               rc = make_tmap_bf16(&ta, d.A, d.M, d.K, d.lda, 64, BK);
  if (rc) return rc;
  if (!d.b_mn_major) rc = make_tmap_bf16(&tb, d.B, d.K, d.N, d.ldb, 64, BN);
  else               rc = make_tmap_bf16(&tb, d.B, d.N, d.K, d.ldb, 64, BK);
  return rc;
}

static int launch_bn(int BN, const GemmBatch& bt, cudaStream_t stream) {
  switch (BN) {
    case 64:  return launch_gemm<64>(bt, stream);
    case 128: return launch_gemm<128>(bt, stream);
    default:  return launch_gemm<256>(bt, stream);
  }
}

int gemm_bf16_batched(const GemmDesc* d, int n, cudaStream_t stream) {
  TT_CHECK(n >= 1, "gemm: empty batch");
  int i = 0;
  while (i < n) {  // greedy groups of consecutive problems with the same tile width and accumulation mode
    // tile width chosen for the whole group: identical problems batched into one launch (the same layer of both towers)
    // together fill the SMs with wider tiles than each would alone
    int same = 1;
    while (i + same < n && same < MAXP && d[i + same].M == d[i].M && d[i + same].N == d[i].N && d[i + same].K == d[i].K &&
           (d[i + same].accumulate != 0) == (d[i].accumulate != 0))
      ++same;
    const int BN = pick_bn(d[i], same);
    int j = i + 1;
    while (j < n && j - i < MAXP && pick_bn(d[j], same) == BN && (d[j].accumulate != 0) == (d[i].accumulate != 0)) ++j;
    GemmBatch bt;
    bt.n = j - i;
    bt.work_start[0] = 0;
    for (int p = 0; p < bt.n; ++p) {
      const int rc = prepare(d[i + p], BN, bt.n, bt.g[p], bt.ta[p], bt.tb[p]);
      if (rc) return rc;
      bt.work_start[p + 1] = bt.work_start[p] + bt.g[p].m_tiles * bt.g[p].n_tiles * bt.g[p].splits;
    }
    for (int p = bt.n; p < MAXP; ++p) bt.work_start[p + 1] = bt.work_start[bt.n];
    const int rc = launch_bn(BN, bt, stream);
    if (rc) return rc;
    i = j;
  }
  return 0;
}

int gemm_bf16(const GemmDesc& d, cudaStream_t stream) { return gemm_bf16_batched(&d, 1, stream); }

}  // namespace tt
