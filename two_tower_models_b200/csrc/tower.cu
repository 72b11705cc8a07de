// Fused tower forward for sm_100a:  emb = [table[ids] | MLP(feats)] Wt^T + bt  in ONE launch per set of towers.
//
// Reference semantics: src/two_tower_base_retrieval.py:112-219 - get_user_embedding (:126) / item id lookup (:209),
// user_features_arch / item_features_arch = Linear(F,256) -> ReLU -> Linear(256,D) (:76-80, :101-105, :153, :211),
// torch.cat (:159-161, :214-216) and user_tower_arch / item_tower_arch = Linear(2D, DI) (:90-93, :107-110, :190, :218).
//
// One CTA owns 128 batch rows of one tower and runs the whole chain on chip:
//   workers   gather table rows + cast the feature rows to bf16 straight into 128B-swizzled UMMA operand tiles
//   GEMM 1    acc1[128,256] = feats W0^T            (A, B in shared memory)
//   epilogue  H = relu(acc1 + b0) -> bf16 into TENSOR MEMORY (A operand of GEMM 2) and to HBM (saved for backward)
//   GEMM 2    acc2[128,D]   = H W1^T                (A in TMEM, B in shared memory)
//   epilogue  Fe = acc2 + b1 -> bf16 into TMEM (A operand of GEMM 3) and into X[:, D:2D] in HBM
//   GEMM 3    acc3[128,DI]  = [id_emb | Fe] Wt^T    (id half: A in shared memory; Fe half: A in TMEM)
//   epilogue  emb = acc3 + bt -> fp32 + bf16 in HBM
// The concatenation never exists as a tensor: it is the K-split of GEMM 3.  The hidden activations only leave the
// SM as the bf16 copies the backward pass needs.  Weights arrive by TMA (the tower weight re-uses W0's buffer once
// GEMM 1 has completed).  Warp roles: 0 TMA, 1 UMMA issue (warp-uniform, elected lane), 2 TMEM alloc, 4-11 workers.
#include <stdlib.h>

#include <type_traits>

#include "common.cuh"
#include "kernels.h"

namespace tt {

namespace {

constexpr int HID = 256;
constexpr int MAXT = 4;  // towers per launch

struct TowerArgs {
  const long long* ids;
  const float* table;
  long long table_rows;
  const float* feats;
  long long ld_feats;
  const float *b0, *b1, *bt;
  bf16* feats16;
  long long ld_feats16;
  bf16* h16;
  long long ldh;
  bf16* x16;
  long long ldx;
  float* emb32;
  long long ld_emb32;
  bf16* emb16;
  long long ld_emb16;
  int rows;
  int tile0;  // first CTA of this tower
};
struct TowerBatch {
  CUtensorMap w0[MAXT], w1[MAXT], wt[MAXT];
  TowerArgs t[MAXT];
  int n;
  int* oob_flag;
  unsigned long long* trace;  // bring-up (TT_TOWER_TRACE): %globaltimer stamps of CTA 0's first worker warp
};

template <int F, int D, int DI>
struct TowerCfg {
  static constexpr int W0_BYTES = HID * F * 2;
  static constexpr int W1_BYTES = D * HID * 2;
  static constexpr int WT_BYTES = DI * 2 * D * 2;
  static constexpr int WA_BYTES = W0_BYTES > WT_BYTES ? W0_BYTES : WT_BYTES;  // W0, later the tower weight
  static constexpr int AF_BYTES = 128 * F * 2;
  static constexpr int AI_BYTES = 128 * D * 2;
  static constexpr int BIAS_FLOATS = HID + D + DI;
  static constexpr int SMEM_BYTES = WA_BYTES + W1_BYTES + AF_BYTES + AI_BYTES + BIAS_FLOATS * 4 + 1024 + 256;
  // tensor memory columns
  static constexpr int ACC1 = 0;     // [128, 256] fp32
  static constexpr int HCOL = 256;   // H as bf16 pairs: 128 columns
  static constexpr int ACC2 = 384;   // [128, D] fp32
  static constexpr int XFCOL = 0;    // Fe as bf16 pairs: D/2 columns (acc1 is dead by then)
  static constexpr int ACC3 = 64;    // [128, DI] fp32
  static_assert(F % 64 == 0 && D % 64 == 0 && DI % 64 == 0 && F <= 128 && D <= 128 && DI <= 128, "tower shape");
  static_assert(SMEM_BYTES <= 232448, "shared memory budget");
};

// bias + (ReLU) of 32 accumulator columns
__device__ __forceinline__ void add_bias32(float* v, const float* bias, bool relu) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias) + j);
    v[4 * j] += b4.x; v[4 * j + 1] += b4.y; v[4 * j + 2] += b4.z; v[4 * j + 3] += b4.w;
  }
  if (relu) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
  }
}
__device__ __forceinline__ void pack32(const float* v, uint32_t* p) {
#pragma unroll
  for (int j = 0; j < 16; ++j) p[j] = pack_bf16x2(v[2 * j], v[2 * j + 1]);
}
__device__ __forceinline__ void store_bf16x32(bf16* dst, const uint32_t* p) {
  if ((reinterpret_cast<uintptr_t>(dst) & 31) == 0) {  // full-sector stores
    st_global_32B(dst, p[0], p[1], p[2], p[3], p[4], p[5], p[6], p[7]);
    st_global_32B(dst + 16, p[8], p[9], p[10], p[11], p[12], p[13], p[14], p[15]);
    return;
  }
#pragma unroll
  for (int j = 0; j < 4; ++j)
    *reinterpret_cast<uint4*>(dst + j * 8) = make_uint4(p[4 * j], p[4 * j + 1], p[4 * j + 2], p[4 * j + 3]);
}

#define TOWER_STAMP(i)                                                       \
  do {                                                                       \
    if (tb.trace != nullptr && blockIdx.x == 0 && threadIdx.x == 128) {      \
      unsigned long long t_;                                                 \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                 \
      tb.trace[i] = t_;                                                      \
    }                                                                        \
  } while (0)

template <int F, int D, int DI>
__global__ void __launch_bounds__(384, 1)
tower_fwd_kernel(const __grid_constant__ TowerBatch tb) {
  using Cfg = TowerCfg<F, D, DI>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sWA = smem;
  uint8_t* sW1 = sWA + Cfg::WA_BYTES;
  uint8_t* sAF = sW1 + Cfg::W1_BYTES;
  uint8_t* sAI = sAF + Cfg::AF_BYTES;
  float* sBias = reinterpret_cast<float*>(sAI + Cfg::AI_BYTES);  // b0 | b1 | bt
  uint64_t* bars = reinterpret_cast<uint64_t*>(sBias + Cfg::BIAS_FLOATS);
  uint64_t* w0_full = bars + 0;
  uint64_t* w1_full = bars + 1;
  uint64_t* wt_full = bars + 2;
  uint64_t* a_full = bars + 3;     // operand tiles written by the 8 worker warps
  uint64_t* acc1_full = bars + 4;
  uint64_t* h_full = bars + 5;     // H sits in TMEM (8 worker warps)
  uint64_t* acc2_full = bars + 6;
  uint64_t* xf_full = bars + 7;    // Fe sits in TMEM (8 worker warps)
  uint64_t* acc3_full = bars + 8;
  uint64_t* bias_full = bars + 9;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 10);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  TOWER_STAMP(0);
  int p = 0;
  while (p + 1 < tb.n && (int)blockIdx.x >= tb.t[p + 1].tile0) ++p;
  const TowerArgs& ta = tb.t[p];
  const int row0 = ((int)blockIdx.x - ta.tile0) * 128;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tb.w0[p]);
    tma_prefetch_desc(&tb.w1[p]);
    tma_prefetch_desc(&tb.wt[p]);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(w0_full, 1);
    mbar_init(w1_full, 1);
    mbar_init(wt_full, 1);
    mbar_init(a_full, 8);
    mbar_init(acc1_full, 1);
    mbar_init(h_full, 8);
    mbar_init(acc2_full, 1);
    mbar_init(xf_full, 8);
    mbar_init(acc3_full, 1);
    mbar_init(bias_full, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_holder, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  TOWER_STAMP(1);

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(w0_full, Cfg::W0_BYTES);
#pragma unroll
      for (int kb = 0; kb < F / 64; ++kb) tma_load_2d(sWA + kb * (HID * 128), &tb.w0[p], w0_full, kb * 64, 0);
      mbar_arrive_expect_tx(w1_full, Cfg::W1_BYTES);
#pragma unroll
      for (int kb = 0; kb < HID / 64; ++kb) tma_load_2d(sW1 + kb * (D * 128), &tb.w1[p], w1_full, kb * 64, 0);
      mbar_wait(acc1_full, 0);  // GEMM 1 has read W0: its buffer takes the tower weight
      mbar_arrive_expect_tx(wt_full, Cfg::WT_BYTES);
#pragma unroll
      for (int kb = 0; kb < 2 * D / 64; ++kb) tma_load_2d(sWA + kb * (DI * 128), &tb.wt[p], wt_full, kb * 64, 0);
    }
  } else if (warp == 1) {
    const uint32_t leader = elect_one();
    // GEMM 1: acc1 = feats W0^T
    mbar_wait(w0_full, 0);
    mbar_wait(a_full, 0);
    tc_fence_after();
    {
      constexpr uint32_t idesc = make_idesc_bf16(128, HID, 0, 0);
      const uint64_t da = make_smem_desc_sw128(smem_u32(sAF), 0, 1024);
      const uint64_t db = make_smem_desc_sw128(smem_u32(sWA), 0, 1024);
#pragma unroll
      for (int k = 0; k < F / 16; ++k)
        umma_bf16_w(tmem_base + Cfg::ACC1, desc_advance(da, (k >> 2) * 16384 + (k & 3) * 32),
                    desc_advance(db, (k >> 2) * (HID * 128) + (k & 3) * 32), idesc, k > 0 ? 1u : 0u, leader);
      umma_commit_w(acc1_full, leader);
    }
    // GEMM 2: acc2 = H W1^T, H read from tensor memory
    mbar_wait(w1_full, 0);
    mbar_wait(h_full, 0);
    tc_fence_after();
    {
      constexpr uint32_t idesc = make_idesc_bf16(128, D, 0, 0);
      const uint64_t db = make_smem_desc_sw128(smem_u32(sW1), 0, 1024);
#pragma unroll
      for (int k = 0; k < HID / 16; ++k)
        umma_bf16_ta_w(tmem_base + Cfg::ACC2, tmem_base + Cfg::HCOL + k * 8,
                       desc_advance(db, (k >> 2) * (D * 128) + (k & 3) * 32), idesc, k > 0 ? 1u : 0u, leader);
      umma_commit_w(acc2_full, leader);
    }
    // GEMM 3: acc3 = id_emb Wt[:, :D]^T + Fe Wt[:, D:]^T
    mbar_wait(wt_full, 0);
    mbar_wait(xf_full, 0);
    tc_fence_after();
    {
      constexpr uint32_t idesc = make_idesc_bf16(128, DI, 0, 0);
      const uint64_t da = make_smem_desc_sw128(smem_u32(sAI), 0, 1024);
      const uint64_t db = make_smem_desc_sw128(smem_u32(sWA), 0, 1024);
#pragma unroll
      for (int k = 0; k < D / 16; ++k)
        umma_bf16_w(tmem_base + Cfg::ACC3, desc_advance(da, (k >> 2) * 16384 + (k & 3) * 32),
                    desc_advance(db, (k >> 2) * (DI * 128) + (k & 3) * 32), idesc, k > 0 ? 1u : 0u, leader);
#pragma unroll
      for (int k = 0; k < D / 16; ++k) {
        const int kk = D / 16 + k;  // k-step of the tower weight
        umma_bf16_ta_w(tmem_base + Cfg::ACC3, tmem_base + Cfg::XFCOL + k * 8,
                       desc_advance(db, (kk >> 2) * (DI * 128) + (kk & 3) * 32), idesc, 1u, leader);
      }
      umma_commit_w(acc3_full, leader);
    }
  } else if (warp == 3) {
    // biases -> shared memory (the epilogues read them as broadcast LDS instead of one L2 round trip per chunk)
    constexpr int N4 = Cfg::BIAS_FLOATS / 4;
    for (int i = lane; i < N4; i += 32) {
      const int f = i * 4;
      const float* src = f < HID ? ta.b0 + f : (f < HID + D ? ta.b1 + (f - HID) : ta.bt + (f - HID - D));
      *reinterpret_cast<float4*>(sBias + f) = __ldg(reinterpret_cast<const float4*>(src));
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(bias_full);
  } else if (warp >= 4) {
    // ---- operand tiles: bf16(feats) and table[ids] into swizzled K-major tiles (and the copies backward needs).
    // Worker warp w owns rows [16 w, 16 w + 16) of the tile; all of a thread's loads are issued before the first
    // one is consumed (a loop of dependent load -> convert -> store steps costs one HBM latency per row).
    {
      const int w = warp - 4;
      // ids of this warp's 16 rows: lane l holds the id of row 16 w + (l & 15)
      long long my_id = -1;
      {
        const int grow = row0 + w * 16 + (lane & 15);
        if (grow < ta.rows) {
          my_id = ta.ids[grow];
          if (my_id < 0 || my_id >= ta.table_rows) {
            if (tb.oob_flag != nullptr) *tb.oob_flag = 1;
            my_id = my_id < 0 ? 0 : ta.table_rows - 1;
          }
        }
      }
      constexpr int F4 = F / 4, D4 = D / 4;              // float4 per row
      constexpr int FR = 32 / F4, DR = 32 / D4;          // rows per warp iteration (1 or 2)
      constexpr int FIT = 16 / FR, DIT = 16 / DR;
      const int fc4 = lane % F4, fsub = lane / F4, dc4 = lane % D4, dsub = lane / D4;
      float4 fv[FIT], dv[DIT];
#pragma unroll
      for (int i = 0; i < FIT; ++i) {  // feature rows: in flight while the ids arrive
        const int grow = row0 + w * 16 + i * FR + fsub;
        fv[i] = grow < ta.rows ? __ldg(reinterpret_cast<const float4*>(ta.feats + (long long)grow * ta.ld_feats) + fc4)
                               : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int i = 0; i < DIT; ++i) {  // embedding rows
        const long long id = __shfl_sync(0xffffffffu, my_id, i * DR + dsub);
        dv[i] = id >= 0 ? __ldg(reinterpret_cast<const float4*>(ta.table + id * D) + dc4) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      auto put = [&](uint8_t* stile, int r, int c4, const float4& v, bf16* copy, long long ld_copy) {
        const uint2 o = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
        if (row0 + r < ta.rows) *reinterpret_cast<uint2*>(copy + (long long)(row0 + r) * ld_copy + c4 * 4) = o;
        *reinterpret_cast<uint2*>(stile + (c4 >> 4) * 16384 + sw128_offset(r, (c4 & 15) >> 1) + (c4 & 1) * 8) = o;
      };
#pragma unroll
      for (int i = 0; i < FIT; ++i) put(sAF, w * 16 + i * FR + fsub, fc4, fv[i], ta.feats16, ta.ld_feats16);
#pragma unroll
      for (int i = 0; i < DIT; ++i) put(sAI, w * 16 + i * DR + dsub, dc4, dv[i], ta.x16, ta.ldx);
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(a_full);
      TOWER_STAMP(2);
    }
    const int q = warp & 3;         // TMEM lane quarter
    const int e = (warp - 4) >> 2;  // column half
    const int r = q * 32 + lane;
    const int grow = row0 + r;
    const bool row_ok = grow < ta.rows;
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    // Each epilogue handles its columns two 32-column chunks at a time: both TMEM loads are issued before the first is
    // consumed, biases come from shared memory.
    auto bias_act = [&](float* v, const float* sb, bool relu) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 b4 = *reinterpret_cast<const float4*>(sb + 4 * j);
        v[4 * j] += b4.x; v[4 * j + 1] += b4.y; v[4 * j + 2] += b4.z; v[4 * j + 3] += b4.w;
      }
      if (relu) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
      }
    };
    mbar_wait(bias_full, 0);
    // ---- epilogue 1: H = relu(acc1 + b0)
    mbar_wait(acc1_full, 0);
    tc_fence_after();
    TOWER_STAMP(3);
#pragma unroll
    for (int c = 0; c < 4; c += 2) {
      const int col = e * 128 + c * 32;
      float v0[32], v1[32];
      uint32_t pk[16];
      tmem_ld32(lane_base + Cfg::ACC1 + col, v0);
      tmem_ld32(lane_base + Cfg::ACC1 + col + 32, v1);
      tmem_wait_ld();
      bias_act(v0, sBias + col, true);
      pack32(v0, pk);
      tmem_st16(lane_base + Cfg::HCOL + col / 2, pk);
      if (row_ok) store_bf16x32(ta.h16 + (long long)grow * ta.ldh + col, pk);
      bias_act(v1, sBias + col + 32, true);
      pack32(v1, pk);
      tmem_st16(lane_base + Cfg::HCOL + col / 2 + 16, pk);
      if (row_ok) store_bf16x32(ta.h16 + (long long)grow * ta.ldh + col + 32, pk);
    }
    tmem_wait_st();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(h_full);
    TOWER_STAMP(4);
    // ---- epilogue 2: Fe = acc2 + b1  (acc1 is dead: Fe overlays its first columns)
    mbar_wait(acc2_full, 0);
    tc_fence_after();
    TOWER_STAMP(5);
    {
      constexpr int NCH = D / 64;  // 32-column chunks per thread (1 or 2)
      const int col = e * (D / 2);
      float v[NCH][32];
      uint32_t pk[16];
#pragma unroll
      for (int c = 0; c < NCH; ++c) tmem_ld32(lane_base + Cfg::ACC2 + col + c * 32, v[c]);
      tmem_wait_ld();
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        bias_act(v[c], sBias + HID + col + c * 32, false);
        pack32(v[c], pk);
        tmem_st16(lane_base + Cfg::XFCOL + (col + c * 32) / 2, pk);
        if (row_ok) store_bf16x32(ta.x16 + (long long)grow * ta.ldx + D + col + c * 32, pk);
      }
    }
    tmem_wait_st();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(xf_full);
    TOWER_STAMP(6);
    // ---- epilogue 3: emb = acc3 + bt
    mbar_wait(acc3_full, 0);
    tc_fence_after();
    TOWER_STAMP(7);
    {
      constexpr int NCH = DI / 64;
      const int col = e * (DI / 2);
      float v[NCH][32];
      uint32_t pk[16];
#pragma unroll
      for (int c = 0; c < NCH; ++c) tmem_ld32(lane_base + Cfg::ACC3 + col + c * 32, v[c]);
      tmem_wait_ld();
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        bias_act(v[c], sBias + HID + D + col + c * 32, false);
        if (row_ok) {
          float* dst = ta.emb32 + (long long)grow * ta.ld_emb32 + col + c * 32;
          if ((reinterpret_cast<uintptr_t>(dst) & 31) == 0) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
              st_global_32B(dst + 8 * j, __float_as_uint(v[c][8 * j]), __float_as_uint(v[c][8 * j + 1]), __float_as_uint(v[c][8 * j + 2]),
                            __float_as_uint(v[c][8 * j + 3]), __float_as_uint(v[c][8 * j + 4]), __float_as_uint(v[c][8 * j + 5]),
                            __float_as_uint(v[c][8 * j + 6]), __float_as_uint(v[c][8 * j + 7]));
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j)
              *reinterpret_cast<float4*>(dst + 4 * j) = make_float4(v[c][4 * j], v[c][4 * j + 1], v[c][4 * j + 2], v[c][4 * j + 3]);
          }
          pack32(v[c], pk);
          store_bf16x32(ta.emb16 + (long long)grow * ta.ld_emb16 + col + c * 32, pk);
        }
      }
    }
    TOWER_STAMP(8);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

template <int F, int D, int DI>
int launch_tower(const TowerBatch& tb, int grid, cudaStream_t stream) {
  using Cfg = TowerCfg<F, D, DI>;
  static bool configured = false;
  if (!configured) {
    TT_CUDA(cudaFuncSetAttribute(tower_fwd_kernel<F, D, DI>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    configured = true;
  }
  KernelSpan span("tower_fwd_kernel", stream);
  tower_fwd_kernel<F, D, DI><<<grid, 384, Cfg::SMEM_BYTES, stream>>>(tb);
  TT_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

}  // namespace

bool tower_fwd_supported(long long F, long long D, long long DI, long long hidden) {
  return hidden == HID && ((F == 128 && D == 128 && DI == 128) || (F == 64 && D == 64 && DI == 64));
}

int tower_fwd(const TowerProblem* pr, int n, int* oob_flag, cudaStream_t stream) {
  TT_CHECK(n >= 1 && n <= MAXT, "tower_fwd: 1..%d towers per launch", MAXT);
  TowerBatch tb;
  tb.n = n;
  tb.oob_flag = oob_flag;
  tb.trace = nullptr;
  if (const char* tr = getenv("TT_TOWER_TRACE")) tb.trace = (unsigned long long*)strtoull(tr, nullptr, 0);
  int grid = 0;
  const long long F = pr[0].F, D = pr[0].D, DI = pr[0].DI;
  TT_CHECK(tower_fwd_supported(F, D, DI, pr[0].hidden), "tower_fwd: unsupported shape F=%lld D=%lld DI=%lld hidden=%lld", F, D,
           DI, pr[0].hidden);
  for (int i = 0; i < n; ++i) {
    const TowerProblem& q = pr[i];
    TT_CHECK(q.F == F && q.D == D && q.DI == DI && q.hidden == HID, "tower_fwd: towers of one launch must share a shape");
    TT_CHECK(q.rows > 0 && q.ids && q.table && q.feats && q.w0 && q.w1 && q.wt && q.b0 && q.b1 && q.bt, "tower_fwd: null argument");
    TT_CHECK(q.feats16 && q.h16 && q.x16 && q.emb32 && q.emb16, "tower_fwd: null output");
    TT_CHECK((q.ld_feats % 4) == 0 && ((uintptr_t)q.feats % 16) == 0 && ((uintptr_t)q.table % 16) == 0,
             "tower_fwd: fp32 inputs need 16-byte aligned rows");
    TT_CHECK((q.ld_feats16 % 8) == 0 && (q.ldh % 8) == 0 && (q.ldx % 8) == 0 && (q.ld_emb16 % 8) == 0 && (q.ld_emb32 % 4) == 0 &&
                 ((uintptr_t)q.feats16 % 16) == 0 && ((uintptr_t)q.h16 % 16) == 0 && ((uintptr_t)q.x16 % 16) == 0 &&
                 ((uintptr_t)q.emb16 % 16) == 0 && ((uintptr_t)q.emb32 % 16) == 0,
             "tower_fwd: outputs need 16-byte aligned rows");
    TT_CHECK(((uintptr_t)q.b0 % 16) == 0 && ((uintptr_t)q.b1 % 16) == 0 && ((uintptr_t)q.bt % 16) == 0,
             "tower_fwd: biases need 16-byte alignment");
    TowerArgs& t = tb.t[i];
    t.ids = q.ids; t.table = q.table; t.table_rows = q.table_rows;
    t.feats = q.feats; t.ld_feats = q.ld_feats;
    t.b0 = q.b0; t.b1 = q.b1; t.bt = q.bt;
    t.feats16 = (bf16*)q.feats16; t.ld_feats16 = q.ld_feats16;
    t.h16 = (bf16*)q.h16; t.ldh = q.ldh;
    t.x16 = (bf16*)q.x16; t.ldx = q.ldx;
    t.emb32 = q.emb32; t.ld_emb32 = q.ld_emb32;
    t.emb16 = (bf16*)q.emb16; t.ld_emb16 = q.ld_emb16;
    t.rows = (int)q.rows;
    t.tile0 = grid;
    grid += (int)((q.rows + 127) / 128);
    int rc = make_tmap_bf16(&tb.w0[i], q.w0, F, HID, q.ldw0, 64, HID);
    if (rc) return rc;
    rc = make_tmap_bf16(&tb.w1[i], q.w1, HID, D, q.ldw1, 64, (uint32_t)D);
    if (rc) return rc;
    rc = make_tmap_bf16(&tb.wt[i], q.wt, 2 * D, DI, q.ldwt, 64, (uint32_t)DI);
    if (rc) return rc;
  }
  for (int i = n; i < MAXT; ++i) {
    tb.t[i] = tb.t[0];
    tb.t[i].tile0 = 0x7fffffff;
    tb.w0[i] = tb.w0[0]; tb.w1[i] = tb.w1[0]; tb.wt[i] = tb.wt[0];
  }
  if (F == 128) return launch_tower<128, 128, 128>(tb, grid, stream);
  return launch_tower<64, 64, 64>(tb, grid, stream);
}

}  // namespace tt
