// Fused tower backward chain for sm_100a (CANDIDATE, off by default: TT_B200_FUSED_TOWER_BWD=1; written at the end of
// round 1 without GPU time left, to be taken through the parity tests before it becomes the default).
//
// Autograd of compute_user_embedding / compute_item_embeddings (reference src/two_tower_base_retrieval.py:112-219,
// backward by autograd, train/train.py:124) up to the activations' gradients, one CTA per 128 batch rows of a tower:
//   GEMM a   acc[128, 2D] = demb Wt                (A: demb tile by TMA; B: Wt read MN-major, i.e. Wt^T without a copy)
//   epilogue dX -> bf16 (HBM);  id half: fp32 rows added straight into the dense table gradient (red.global.add.v4);
//            feature half dFe: bf16 into TENSOR MEMORY (A operand of GEMM b) + fp32 column sums (= db1)
//   GEMM b   acc[128, 256] = dFe W1                (A in TMEM; B: W1 read MN-major)
//   epilogue dH = acc where H > 0 else 0 -> bf16 (HBM) + fp32 column sums (= db0)
// The three weight gradients (dWt = demb^T X, dW1 = dFe^T H, dW0 = dH^T feats) stay split-K GEMMs (gemm.cu); this
// kernel replaces the dX GEMM, the dH GEMM and the two scatter-add launches (the critical chain of the tower backward).
// Column sums are reduced per CTA in shared memory: one global atomic per column per CTA.
#include <stdlib.h>

#include "common.cuh"
#include "kernels.h"

namespace tt {

namespace {

constexpr int HIDB = 256;
constexpr int MAXTB = 4;

struct TowerBwdArgs {
  const long long* ids;
  long long table_rows;
  const bf16* h16;   // [rows, 256] forward activations (ReLU mask)
  long long ldh;
  bf16* dx16;        // out [rows, 2D]
  long long lddx;
  bf16* dh16;        // out [rows, 256]
  long long lddh;
  float* dtable;     // [table_rows, D] fp32, += (may be null: no scatter, e.g. data parallel row exchange)
  float* dxsum;      // [2D] fp32, += column sums of dX (only the feature half is written: db1)
  float* db0;        // [256] fp32, += column sums of dH
  int rows;
  int tile0;
};
struct TowerBwdBatch {
  CUtensorMap demb[MAXTB], wt[MAXTB], w1[MAXTB];
  TowerBwdArgs t[MAXTB];
  int n;
};

template <int D, int DI>
struct TowerBwdCfg {
  static constexpr int N1 = 2 * D;                    // columns of dX
  static constexpr int A_BYTES = 128 * DI * 2;        // demb tile, K-major, DI/64 k-blocks of 16 KB
  static constexpr int B1_KB_BYTES = (N1 / 64) * 8192;   // one 64-row k-block of Wt (MN-major): N1/64 atoms of 8 KB
  static constexpr int B1_BYTES = (DI / 64) * B1_KB_BYTES;
  static constexpr int B2_KB_BYTES = (HIDB / 64) * 8192;
  static constexpr int B2_BYTES = (D / 64) * B2_KB_BYTES;
  static constexpr int CS_FLOATS = D + HIDB;          // shared-memory column sums: dFe | dH
  static constexpr int SMEM_BYTES = A_BYTES + B1_BYTES + B2_BYTES + CS_FLOATS * 4 + 1024 + 256;
  static constexpr int ACC = 0;       // acc a [128, 2D] then acc b [128, 256]
  static constexpr int DFE = 256;     // dFe as bf16 pairs: D/2 columns
  static_assert(D % 64 == 0 && DI % 64 == 0 && D <= 128 && DI <= 128, "tower shape");
  static_assert(SMEM_BYTES <= 232448, "shared memory budget");
};

// butterfly transpose-reduce over the warp's 32 rows: afterwards lane l holds the sum of column l of the chunk
__device__ __forceinline__ float warp_colsum32(float* v, int lane) {
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
  return v[0];
}

template <int D, int DI>
__global__ void __launch_bounds__(384, 1)
tower_bwd_kernel(const __grid_constant__ TowerBwdBatch tb) {
  using Cfg = TowerBwdCfg<D, DI>;
  constexpr int N1 = Cfg::N1;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB1 = sA + Cfg::A_BYTES;
  uint8_t* sB2 = sB1 + Cfg::B1_BYTES;
  float* sCS = reinterpret_cast<float*>(sB2 + Cfg::B2_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sCS + Cfg::CS_FLOATS);
  uint64_t* ab_full = bars + 0;    // demb tile + Wt landed
  uint64_t* w1_full = bars + 1;
  uint64_t* acca_full = bars + 2;
  uint64_t* dfe_full = bars + 3;   // every worker warp is done with acc a; dFe sits in TMEM (8 warps)
  uint64_t* accb_full = bars + 4;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 5);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int p = 0;
  while (p + 1 < tb.n && (int)blockIdx.x >= tb.t[p + 1].tile0) ++p;
  const TowerBwdArgs& ta = tb.t[p];
  const int row0 = ((int)blockIdx.x - ta.tile0) * 128;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tb.demb[p]);
    tma_prefetch_desc(&tb.wt[p]);
    tma_prefetch_desc(&tb.w1[p]);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(ab_full, 1);
    mbar_init(w1_full, 1);
    mbar_init(acca_full, 1);
    mbar_init(dfe_full, 8);
    mbar_init(accb_full, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_holder, 512);
  for (int i = threadIdx.x; i < Cfg::CS_FLOATS; i += blockDim.x) sCS[i] = 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(ab_full, Cfg::A_BYTES + Cfg::B1_BYTES);
#pragma unroll
      for (int kb = 0; kb < DI / 64; ++kb) tma_load_2d(sA + kb * 16384, &tb.demb[p], ab_full, kb * 64, row0);
#pragma unroll
      for (int kb = 0; kb < DI / 64; ++kb)
#pragma unroll
        for (int b = 0; b < N1 / 64; ++b)  // Wt stored [DI, 2D]: box = 64 N-elements x 64 K-rows
          tma_load_2d(sB1 + kb * Cfg::B1_KB_BYTES + b * 8192, &tb.wt[p], ab_full, b * 64, kb * 64);
      mbar_arrive_expect_tx(w1_full, Cfg::B2_BYTES);
#pragma unroll
      for (int kb = 0; kb < D / 64; ++kb)
#pragma unroll
        for (int b = 0; b < HIDB / 64; ++b)  // W1 stored [D, 256]
          tma_load_2d(sB2 + kb * Cfg::B2_KB_BYTES + b * 8192, &tb.w1[p], w1_full, b * 64, kb * 64);
    }
  } else if (warp == 1) {
    const uint32_t leader = elect_one();
    // GEMM a: acc = demb Wt  (B MN-major: 8-k groups every 1024 B, 64-element N atoms every 8192 B, 16 k-rows = 2048 B)
    mbar_wait(ab_full, 0);
    tc_fence_after();
    {
      constexpr uint32_t idesc = make_idesc_bf16(128, N1, 0, 1);
      const uint64_t da = make_smem_desc_sw128(smem_u32(sA), 0, 1024);
      const uint64_t db = make_smem_desc_sw128(smem_u32(sB1), 8192, 1024);
#pragma unroll
      for (int k = 0; k < DI / 16; ++k)
        umma_bf16_w(tmem_base + Cfg::ACC, desc_advance(da, (k >> 2) * 16384 + (k & 3) * 32),
                    desc_advance(db, (k >> 2) * Cfg::B1_KB_BYTES + (k & 3) * 2048), idesc, k > 0 ? 1u : 0u, leader);
      umma_commit_w(acca_full, leader);
    }
    // GEMM b: acc = dFe W1, dFe from tensor memory (acc a is dead once every worker has arrived on dfe_full)
    mbar_wait(w1_full, 0);
    mbar_wait(dfe_full, 0);
    tc_fence_after();
    {
      constexpr uint32_t idesc = make_idesc_bf16(128, HIDB, 0, 1);
      const uint64_t db = make_smem_desc_sw128(smem_u32(sB2), 8192, 1024);
#pragma unroll
      for (int k = 0; k < D / 16; ++k)
        umma_bf16_ta_w(tmem_base + Cfg::ACC, tmem_base + Cfg::DFE + k * 8,
                       desc_advance(db, (k >> 2) * Cfg::B2_KB_BYTES + (k & 3) * 2048), idesc, k > 0 ? 1u : 0u, leader);
      umma_commit_w(accb_full, leader);
    }
  } else if (warp >= 4) {
    const int q = warp & 3;         // TMEM lane quarter
    const int e = (warp - 4) >> 2;  // column half: 0 = id embedding half of dX, 1 = feature half
    const int r = q * 32 + lane;
    const int grow = row0 + r;
    const bool row_ok = grow < ta.rows;
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    long long id = 0;
    if (e == 0 && row_ok && ta.dtable != nullptr) {
      id = ta.ids[grow];
      id = id < 0 ? 0 : (id >= ta.table_rows ? ta.table_rows - 1 : id);  // forward clamps the same way (and flags it)
    }
    // ---- epilogue a: dX
    mbar_wait(acca_full, 0);
    tc_fence_after();
#pragma unroll 1
    for (int c = 0; c < D / 32; ++c) {
      const int col = e * D + c * 32;  // column of dX
      float v[32];
      uint32_t pk[16];
      tmem_ld32(lane_base + Cfg::ACC + col, v);
      tmem_wait_ld();
#pragma unroll
      for (int j = 0; j < 16; ++j) pk[j] = pack_bf16x2(v[2 * j], v[2 * j + 1]);
      if (row_ok) {
        bf16* dst = ta.dx16 + (long long)grow * ta.lddx + col;
#pragma unroll
        for (int j = 0; j < 4; ++j)
          *reinterpret_cast<uint4*>(dst + j * 8) = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
      }
      if (e == 0) {
        if (row_ok && ta.dtable != nullptr) {  // dense embedding gradient: duplicates accumulate
          float* g = ta.dtable + id * D + c * 32;
#pragma unroll
          for (int j = 0; j < 8; ++j)
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(g + j * 4), "f"(v[4 * j]), "f"(v[4 * j + 1]),
                         "f"(v[4 * j + 2]), "f"(v[4 * j + 3])
                         : "memory");
        }
      } else {
        tmem_st16(lane_base + Cfg::DFE + c * 16, pk);  // dFe columns [c*32, c*32+32) as bf16 pairs
        const float cs = warp_colsum32(v, lane);       // rows past the batch are zero (TMA zero fill)
        atomicAdd(&sCS[c * 32 + lane], cs);
      }
    }
    if (e == 1) tmem_wait_st();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(dfe_full);
    // ---- epilogue b: dH = acc masked by H > 0
    mbar_wait(accb_full, 0);
    tc_fence_after();
#pragma unroll 1
    for (int c = 0; c < HIDB / 64; ++c) {
      const int col = e * (HIDB / 2) + c * 32;
      float v[32];
      uint32_t pk[16];
      tmem_ld32(lane_base + Cfg::ACC + col, v);
      uint4 mv[4];
      if (row_ok) {
        const bf16* mp = ta.h16 + (long long)grow * ta.ldh + col;
#pragma unroll
        for (int j = 0; j < 4; ++j) mv[j] = *reinterpret_cast<const uint4*>(mp + j * 8);
      }
      tmem_wait_ld();
      if (row_ok) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const bf16* mb = reinterpret_cast<const bf16*>(&mv[j]);
#pragma unroll
          for (int k = 0; k < 8; ++k)
            if (!(__bfloat162float(mb[k]) > 0.f)) v[j * 8 + k] = 0.f;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0.f;
      }
#pragma unroll
      for (int j = 0; j < 16; ++j) pk[j] = pack_bf16x2(v[2 * j], v[2 * j + 1]);
      if (row_ok) {
        bf16* dst = ta.dh16 + (long long)grow * ta.lddh + col;
#pragma unroll
        for (int j = 0; j < 4; ++j)
          *reinterpret_cast<uint4*>(dst + j * 8) = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
      }
      const float cs = warp_colsum32(v, lane);
      atomicAdd(&sCS[D + col + lane], cs);
    }
    // ---- column sums of the CTA -> global (one atomic per column per CTA)
    asm volatile("bar.sync 1, 256;" ::: "memory");  // the 8 worker warps
    const int wt_id = threadIdx.x - 128;
    for (int i = wt_id; i < Cfg::CS_FLOATS; i += 256) {
      if (i < D) atomicAdd(ta.dxsum + D + i, sCS[i]);
      else atomicAdd(ta.db0 + (i - D), sCS[i]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

template <int D, int DI>
int launch_tower_bwd(const TowerBwdBatch& tb, int grid, cudaStream_t stream) {
  using Cfg = TowerBwdCfg<D, DI>;
  static bool configured = false;
  if (!configured) {
    TT_CUDA(cudaFuncSetAttribute(tower_bwd_kernel<D, DI>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    configured = true;
  }
  KernelSpan span("tower_bwd_kernel", stream);
  tower_bwd_kernel<D, DI><<<grid, 384, Cfg::SMEM_BYTES, stream>>>(tb);
  TT_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

}  // namespace

int tower_bwd_chain(const TowerBwdProblem* pr, int n, cudaStream_t stream) {
  TT_CHECK(n >= 1 && n <= MAXTB, "tower_bwd_chain: 1..%d towers per launch", MAXTB);
  TowerBwdBatch tb;
  tb.n = n;
  int grid = 0;
  const long long D = pr[0].D, DI = pr[0].DI;
  TT_CHECK(tower_fwd_supported(D, D, DI, pr[0].hidden), "tower_bwd_chain: unsupported shape D=%lld DI=%lld hidden=%lld", D, DI,
           pr[0].hidden);
  for (int i = 0; i < n; ++i) {
    const TowerBwdProblem& q = pr[i];
    TT_CHECK(q.D == D && q.DI == DI && q.hidden == HIDB, "tower_bwd_chain: towers of one launch must share a shape");
    TT_CHECK(q.rows > 0 && q.demb16 && q.wt && q.w1 && q.h16 && q.dx16 && q.dh16 && q.dxsum && q.db0, "tower_bwd_chain: null argument");
    TT_CHECK(q.dtable == nullptr || q.ids != nullptr, "tower_bwd_chain: ids needed for the table gradient");
    TT_CHECK((q.ld_demb % 8) == 0 && (q.ldh % 8) == 0 && (q.lddx % 8) == 0 && (q.lddh % 8) == 0 &&
                 ((uintptr_t)q.demb16 % 16) == 0 && ((uintptr_t)q.h16 % 16) == 0 && ((uintptr_t)q.dx16 % 16) == 0 &&
                 ((uintptr_t)q.dh16 % 16) == 0 && (q.dtable == nullptr || ((uintptr_t)q.dtable % 16) == 0),
             "tower_bwd_chain: operands need 16-byte aligned rows");
    TowerBwdArgs& t = tb.t[i];
    t.ids = q.ids; t.table_rows = q.table_rows;
    t.h16 = (const bf16*)q.h16; t.ldh = q.ldh;
    t.dx16 = (bf16*)q.dx16; t.lddx = q.lddx;
    t.dh16 = (bf16*)q.dh16; t.lddh = q.lddh;
    t.dtable = q.dtable; t.dxsum = q.dxsum; t.db0 = q.db0;
    t.rows = (int)q.rows;
    t.tile0 = grid;
    grid += (int)((q.rows + 127) / 128);
    int rc = make_tmap_bf16(&tb.demb[i], q.demb16, DI, q.rows, q.ld_demb, 64, 128);
    if (rc) return rc;
    rc = make_tmap_bf16(&tb.wt[i], q.wt, 2 * D, DI, q.ldwt, 64, 64);   // stored [DI, 2D]: inner = N, outer = K
    if (rc) return rc;
    rc = make_tmap_bf16(&tb.w1[i], q.w1, HIDB, D, q.ldw1, 64, 64);     // stored [D, 256]
    if (rc) return rc;
  }
  for (int i = n; i < MAXTB; ++i) {
    tb.t[i] = tb.t[0];
    tb.t[i].tile0 = 0x7fffffff;
    tb.demb[i] = tb.demb[0]; tb.wt[i] = tb.wt[0]; tb.w1[i] = tb.w1[0];
  }
  if (D == 128) return launch_tower_bwd<128, 128>(tb, grid, stream);
  return launch_tower_bwd<64, 64>(tb, grid, stream);
}

}  // namespace tt
