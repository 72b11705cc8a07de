// Tensor-core self-attention forward for the history encoder (sm_100a, tcgen05 + TMEM + TMA).
//
// Reference: nn.MultiheadAttention inside src/user_history_encoder.py:60-67,103-108 (per head
// softmax(q k^T / sqrt(hd)) v, no mask, no dropout).  Sequences are short (H <= 128, head_dim 16..64), so a
// 128-row UMMA tile holds TWO sequences when H <= 64 (rows [0,64) and [64,128), zero padded) or one otherwise.
// Per tile and head:
//     S = Q_h K_h^T          UMMA 128 x 128 x hd, both operands from the TMA-staged q|k|v tile (128-B swizzle)
//     P = softmax rows        one thread per row reads ITS sequence's block of S from TMEM, writes bf16 P back
//                             into TMEM (tcgen05.st); the cross-sequence blocks of P stay zero
//     O_h = P V_h            UMMA 128 x hd x 128 with P as the TMEM A operand and V_h read MN-major from the
//                             same shared-memory tile
// S and P are double buffered in TMEM so that the UMMAs of head h+1 overlap the softmax of head h; the q|k|v
// tile is double buffered in shared memory so that the TMA loads of the next tile overlap this one.  The
// kernel is HBM bound (reads 3 D, writes D bf16 per token) - the tensor work per tile is ~1k cycles.
//
// Warps: 0 TMA producer, 1 UMMA issuer (warp-uniform, elected lane), 2 TMEM allocator, 4-7 softmax/epilogue.
#include "common.cuh"
#include "kernels.h"

namespace tt {

namespace {

constexpr float LOG2E_A = 1.4426950408889634f;

struct AttnTcArgs {
  int nseq, H, D, heads, hd, q_rows;
  int HP;      // rows reserved per sequence inside a 128-row tile (64 or 128)
  int ntiles;
  bf16* out;
  long long ldo;
  float scale_log2;
};

constexpr int S_COL = 0;    // 2 x 128 fp32 score columns
constexpr int P_COL = 256;  // 2 x 64 columns of packed bf16 probabilities (128 keys)
constexpr int O_COL = 384;  // up to 128 fp32 output columns

__global__ void __launch_bounds__(256, 1)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmq, const AttnTcArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int natoms = 3 * a.D / 64;
  const int tile_bytes = natoms * 16384;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * tile_bytes);
  uint64_t* t_full = bars;        // [2] q|k|v tile landed
  uint64_t* t_empty = bars + 2;   // [2] all UMMAs reading the tile completed
  uint64_t* s_full = bars + 4;    // [2]
  uint64_t* s_empty = bars + 6;   // [2] count 4
  uint64_t* p_full = bars + 8;    // [2] count 4
  uint64_t* p_empty = bars + 10;  // [2]
  uint64_t* o_full = bars + 12;
  uint64_t* o_empty = bars + 13;  // count 4
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 14);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int spt = 128 / a.HP;  // sequences per tile

  // zero both tile buffers once: the pad rows [H, HP) of every slot are never written by TMA
  for (int i = threadIdx.x; i < 2 * tile_bytes / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(smem)[i] = make_uint4(0u, 0u, 0u, 0u);
  fence_proxy_async_smem();
  if (warp == 0 && lane == 0) tma_prefetch_desc(&tmq);
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&t_full[i], 1);
      mbar_init(&t_empty[i], 1);
      mbar_init(&s_full[i], 1);
      mbar_init(&s_empty[i], 4);
      mbar_init(&p_full[i], 4);
      mbar_init(&p_empty[i], 1);
    }
    mbar_init(o_full, 1);
    mbar_init(o_empty, 4);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_holder, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x, ++it) {
        const uint32_t buf = it & 1;
        mbar_wait(&t_empty[buf], ((it >> 1) & 1) ^ 1);
        int nvalid = a.nseq - tile * spt;
        nvalid = nvalid < spt ? nvalid : spt;
        mbar_arrive_expect_tx(&t_full[buf], (uint32_t)(nvalid * natoms * a.H * 128));
        uint8_t* dst = smem + buf * tile_bytes;
        for (int s = 0; s < nvalid; ++s)
          for (int b = 0; b < natoms; ++b)
            tma_load_2d(dst + b * 16384 + s * a.HP * 128, &tmq, &t_full[buf], b * 64, (tile * spt + s) * a.H);
      }
    }
  } else if (warp == 1) {
    const uint32_t leader = elect_one();
    const uint32_t idesc_s = make_idesc_bf16(128, 128, 0, 0);
    const uint32_t idesc_o = make_idesc_bf16(128, a.hd, 0, 1);
    uint32_t it = 0, n = 0;  // n counts (tile, head) pairs: S / P buffer = n & 1
    for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x, ++it) {
      const uint32_t buf = it & 1;
      mbar_wait(&t_full[buf], (it >> 1) & 1);
      tc_fence_after();
      const uint32_t tb = smem_u32(smem + buf * tile_bytes);
      auto issue_s = [&](int h, uint32_t nn) {
        const uint32_t sb = nn & 1;
        mbar_wait(&s_empty[sb], ((nn >> 1) & 1) ^ 1);
        tc_fence_after();
        const int qc = h * a.hd, kc = a.D + h * a.hd;
        const uint64_t dq = make_smem_desc_sw128(tb + (qc >> 6) * 16384 + (qc & 63) * 2, 0, 1024);
        const uint64_t dk = make_smem_desc_sw128(tb + (kc >> 6) * 16384 + (kc & 63) * 2, 0, 1024);
        for (int kk = 0; kk < a.hd / 16; ++kk)
          umma_bf16_w(tmem_base + S_COL + sb * 128, desc_advance(dq, kk * 32), desc_advance(dk, kk * 32), idesc_s,
                      kk > 0 ? 1u : 0u, leader);
        umma_commit_w(&s_full[sb], leader);
      };
      issue_s(0, n);
      for (int h = 0; h < a.heads; ++h, ++n) {
        if (h + 1 < a.heads) issue_s(h + 1, n + 1);
        const uint32_t pb = n & 1;
        mbar_wait(&p_full[pb], (n >> 1) & 1);
        if (h == 0) mbar_wait(o_empty, (it & 1) ^ 1);
        tc_fence_after();
        const int vc = 2 * a.D + h * a.hd;
        const uint64_t dv = make_smem_desc_sw128(tb + (vc >> 6) * 16384 + (vc & 63) * 2, 16384, 1024);
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)
          umma_bf16_ta_w(tmem_base + O_COL + h * a.hd, tmem_base + P_COL + pb * 64 + kk * 8, desc_advance(dv, kk * 2048),
                         idesc_o, kk > 0 ? 1u : 0u, leader);
        umma_commit_w(&p_empty[pb], leader);
      }
      umma_commit_w(o_full, leader);
      umma_commit_w(&t_empty[buf], leader);
    }
  } else if (warp >= 4) {
    const int q = warp & 3;
    const int row = q * 32 + lane;       // row of the 128-row tile
    const int slot = row / a.HP;          // sequence slot inside the tile
    const int i = row - slot * a.HP;      // position inside the sequence
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    const int nch = a.HP / 32;            // 32-column chunks of this row's own score block
    {  // P buffers start as zeros: the cross-sequence blocks are never written again
      uint32_t z[16];
#pragma unroll
      for (int k = 0; k < 16; ++k) z[k] = 0u;
#pragma unroll
      for (int c = 0; c < 8; ++c) tmem_st16(lane_base + P_COL + c * 16, z);
      tmem_wait_st();
    }
    uint32_t it = 0, n = 0;
    for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x, ++it) {
      const long long seq = (long long)tile * spt + slot;
      for (int h = 0; h < a.heads; ++h, ++n) {
        const uint32_t sb = n & 1;
        mbar_wait(&s_full[sb], (n >> 1) & 1);
        tc_fence_after();
        // pass 1: row maximum over the H valid keys of this row's sequence
        float v[4][32];
        float mx = -INFINITY;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          if (c < nch) {
            tmem_ld32(lane_base + S_COL + sb * 128 + slot * a.HP + c * 32, v[c]);
            tmem_wait_ld();
#pragma unroll
            for (int k = 0; k < 32; ++k) {
              if (c * 32 + k >= a.H) v[c][k] = -INFINITY;
              mx = fmaxf(mx, v[c][k]);
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_empty[sb]);  // scores are in registers: the buffer can be refilled
        const float ms = mx * a.scale_log2;
        float sum = 0.f;
#pragma unroll
        for (int c = 0; c < 4; ++c)
          if (c < nch) {
#pragma unroll
            for (int k = 0; k < 32; ++k) {
              v[c][k] = ex2f(fmaf(v[c][k], a.scale_log2, -ms));
              sum += v[c][k];
            }
          }
        const float inv = 1.f / sum;
        mbar_wait(&p_empty[sb], ((n >> 1) & 1) ^ 1);
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < 4; ++c)
          if (c < nch) {
            uint32_t pk[16];
#pragma unroll
            for (int k = 0; k < 16; ++k) pk[k] = pack_bf16x2(v[c][2 * k] * inv, v[c][2 * k + 1] * inv);
            tmem_st16(lane_base + P_COL + sb * 64 + (slot * a.HP) / 2 + c * 16, pk);
          }
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[sb]);
      }
      // all heads done: O[128, D] -> bf16 rows of the output
      mbar_wait(o_full, it & 1);
      tc_fence_after();
      const bool wr = seq < a.nseq && i < a.q_rows;
      bf16* dst = a.out + (seq * a.q_rows + i) * a.ldo;
      const bool st32 = ((reinterpret_cast<uintptr_t>(a.out) | (uintptr_t)(a.ldo * 2)) & 31) == 0;
      for (int c = 0; c < a.D / 32; ++c) {
        float o[32];
        tmem_ld32(lane_base + O_COL + c * 32, o);
        tmem_wait_ld();
        if (wr && st32) {
#pragma unroll
          for (int k = 0; k < 2; ++k)
            st_global_32B(dst + c * 32 + k * 16, pack_bf16x2(o[16 * k + 0], o[16 * k + 1]), pack_bf16x2(o[16 * k + 2], o[16 * k + 3]),
                          pack_bf16x2(o[16 * k + 4], o[16 * k + 5]), pack_bf16x2(o[16 * k + 6], o[16 * k + 7]),
                          pack_bf16x2(o[16 * k + 8], o[16 * k + 9]), pack_bf16x2(o[16 * k + 10], o[16 * k + 11]),
                          pack_bf16x2(o[16 * k + 12], o[16 * k + 13]), pack_bf16x2(o[16 * k + 14], o[16 * k + 15]));
        } else if (wr) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            uint4 u;
            u.x = pack_bf16x2(o[8 * k + 0], o[8 * k + 1]);
            u.y = pack_bf16x2(o[8 * k + 2], o[8 * k + 3]);
            u.z = pack_bf16x2(o[8 * k + 4], o[8 * k + 5]);
            u.w = pack_bf16x2(o[8 * k + 6], o[8 * k + 7]);
            *reinterpret_cast<uint4*>(dst + c * 32 + k * 8) = u;
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(o_empty);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ---------------------------------------------------------------------------------------------
// backward: per tile and head
//     S  = Q_h K_h^T,  dP = dO_h V_h^T                      (UMMA, operands from the q|k|v and dO tiles)
//     P  = softmax rows, delta_i = sum_j P_ij dP_ij,  dS = P (dP - delta) / sqrt(hd)      (one thread per row)
//     dQ_h = dS K_h,   dK_h = dS^T Q_h,   dV_h = P^T dO_h   (UMMA; P and dS are staged as bf16 [query][key] tiles in
//                                                            shared memory and read K-major or MN-major)
// Query rows without an upstream gradient (pad rows, rows >= q_rows of the last layer) have dO = 0 => dS = 0.
// ---------------------------------------------------------------------------------------------
struct AttnBwdArgs {
  int nseq, H, D, heads, hd, q_rows, HP, ntiles, nacc;
  bf16* dqkv;
  long long lddqkv;
  float scale, scale_log2;
};

constexpr int BS_COL = 0;     // scores
constexpr int BDP_COL = 128;  // dP
constexpr int BACC_COL = 256; // nacc x [dQ_h | dK_h | dV_h] (3 hd columns each)

__global__ void __launch_bounds__(256, 1)
attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmq, const __grid_constant__ CUtensorMap tmdo, const AttnBwdArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int qatoms = 3 * a.D / 64, oatoms = a.D / 64;
  uint8_t* sq = smem;                         // q|k|v tile
  uint8_t* sdo = sq + qatoms * 16384;         // dO tile
  uint8_t* sp = sdo + oatoms * 16384;         // P  [128 queries][128 keys] bf16 (2 atoms)
  uint8_t* sds = sp + 32768;                  // dS
  uint64_t* bars = reinterpret_cast<uint64_t*>(sds + 32768);
  uint64_t* t_full = bars;          // tile (q|k|v + dO) landed
  uint64_t* t_empty = bars + 1;     // all UMMAs of the tile completed
  uint64_t* sdp_full = bars + 2;    // S and dP of a head are in TMEM
  uint64_t* sdp_empty = bars + 3;   // count 4: both are in registers
  uint64_t* pds_full = bars + 4;    // count 4: P and dS tiles written
  uint64_t* pds_empty = bars + 5;   // the three gradient UMMAs of a head completed
  uint64_t* acc_full = bars + 6;    // [2]
  uint64_t* acc_empty = bars + 8;   // [2] count 4
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 10);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int spt = 128 / a.HP;
  const int smem_data = (qatoms + oatoms) * 16384 + 65536;
  for (int i = threadIdx.x; i < smem_data / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0u, 0u, 0u, 0u);
  fence_proxy_async_smem();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmq);
    tma_prefetch_desc(&tmdo);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(t_full, 1);
    mbar_init(t_empty, 1);
    mbar_init(sdp_full, 1);
    mbar_init(sdp_empty, 4);
    mbar_init(pds_full, 4);
    mbar_init(pds_empty, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], 4);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_holder, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x, ++it) {
        mbar_wait(t_empty, (it & 1) ^ 1);
        int nvalid = a.nseq - tile * spt;
        nvalid = nvalid < spt ? nvalid : spt;
        mbar_arrive_expect_tx(t_full, (uint32_t)(nvalid * (qatoms * a.H + oatoms * a.q_rows) * 128));
        for (int s = 0; s < nvalid; ++s) {
          for (int b = 0; b < qatoms; ++b)
            tma_load_2d(sq + b * 16384 + s * a.HP * 128, &tmq, t_full, b * 64, (tile * spt + s) * a.H);
          for (int b = 0; b < oatoms; ++b)
            tma_load_2d(sdo + b * 16384 + s * a.HP * 128, &tmdo, t_full, b * 64, (tile * spt + s) * a.q_rows);
        }
      }
    }
  } else if (warp == 1) {
    const uint32_t leader = elect_one();
    const uint32_t idesc_s = make_idesc_bf16(128, 128, 0, 0);   // S, dP: K-major x K-major
    const uint32_t idesc_q = make_idesc_bf16(128, a.hd, 0, 1);  // dQ = dS K_h      : A K-major, B MN-major
    const uint32_t idesc_t = make_idesc_bf16(128, a.hd, 1, 1);  // dK, dV          : A MN-major (transposed), B MN-major
    const uint32_t sqa = smem_u32(sq), sdoa = smem_u32(sdo), spa = smem_u32(sp), sdsa = smem_u32(sds);
    uint32_t it = 0, n = 0;
    for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x, ++it) {
      mbar_wait(t_full, it & 1);
      tc_fence_after();
      // S = Q K^T and dP = dO V^T of head hh (global head counter nn); the softmax threads copy both into
      // registers right away, so the single TMEM buffer is free again for the next head almost immediately
      auto issue_sdp = [&](int hh, uint32_t nn) {
        const int qc = hh * a.hd, kc = a.D + hh * a.hd, vc = 2 * a.D + hh * a.hd;
        const uint32_t q_addr = sqa + (qc >> 6) * 16384 + (qc & 63) * 2;
        const uint32_t k_addr = sqa + (kc >> 6) * 16384 + (kc & 63) * 2;
        const uint32_t v_addr = sqa + (vc >> 6) * 16384 + (vc & 63) * 2;
        const uint32_t do_addr = sdoa + (qc >> 6) * 16384 + (qc & 63) * 2;
        mbar_wait(sdp_empty, (nn & 1) ^ 1);
        tc_fence_after();
        const uint64_t dq = make_smem_desc_sw128(q_addr, 0, 1024), dk = make_smem_desc_sw128(k_addr, 0, 1024);
        const uint64_t dd = make_smem_desc_sw128(do_addr, 0, 1024), dv = make_smem_desc_sw128(v_addr, 0, 1024);
        for (int kk = 0; kk < a.hd / 16; ++kk)
          umma_bf16_w(tmem_base + BS_COL, desc_advance(dq, kk * 32), desc_advance(dk, kk * 32), idesc_s, kk > 0 ? 1u : 0u, leader);
        for (int kk = 0; kk < a.hd / 16; ++kk)
          umma_bf16_w(tmem_base + BDP_COL, desc_advance(dd, kk * 32), desc_advance(dv, kk * 32), idesc_s, kk > 0 ? 1u : 0u, leader);
        umma_commit_w(sdp_full, leader);
      };
      issue_sdp(0, n);
      for (int h = 0; h < a.heads; ++h, ++n) {
        const int qc = h * a.hd, kc = a.D + h * a.hd;
        const uint32_t q_addr = sqa + (qc >> 6) * 16384 + (qc & 63) * 2;
        const uint32_t k_addr = sqa + (kc >> 6) * 16384 + (kc & 63) * 2;
        const uint32_t do_addr = sdoa + (qc >> 6) * 16384 + (qc & 63) * 2;
        if (h + 1 < a.heads) issue_sdp(h + 1, n + 1);
        // gradients of this head once P and dS are staged
        const uint32_t ab = a.nacc == 2 ? (n & 1) : 0;
        const uint32_t use = a.nacc == 2 ? (n >> 1) : n;
        mbar_wait(pds_full, n & 1);
        mbar_wait(&acc_empty[ab], (use & 1) ^ 1);
        tc_fence_after();
        {
          const uint32_t acc = tmem_base + BACC_COL + ab * 3 * a.hd;
          const uint64_t ds_k = make_smem_desc_sw128(sdsa, 0, 1024);        // dS as A[M = query, K = key]
          const uint64_t ds_t = make_smem_desc_sw128(sdsa, 16384, 1024);    // dS as A[M = key, K = query] (MN-major)
          const uint64_t p_t = make_smem_desc_sw128(spa, 16384, 1024);      // P  as A[M = key, K = query]
          const uint64_t bk = make_smem_desc_sw128(k_addr, 16384, 1024);    // K_h  as B[K = key,   N = hd] (MN-major)
          const uint64_t bq = make_smem_desc_sw128(q_addr, 16384, 1024);    // Q_h  as B[K = query, N = hd]
          const uint64_t bd = make_smem_desc_sw128(do_addr, 16384, 1024);   // dO_h as B[K = query, N = hd]
#pragma unroll
          for (int kk = 0; kk < 8; ++kk)  // dQ: K runs over the 128 keys = 2 atoms of 64 (K-major A: 32 B per step)
            umma_bf16_w(acc, desc_advance(ds_k, (kk >> 2) * 16384 + (kk & 3) * 32), desc_advance(bk, kk * 2048), idesc_q,
                        kk > 0 ? 1u : 0u, leader);
#pragma unroll
          for (int kk = 0; kk < 8; ++kk)  // dK: K runs over the 128 query rows (16 rows = 2048 B per step)
            umma_bf16_w(acc + a.hd, desc_advance(ds_t, kk * 2048), desc_advance(bq, kk * 2048), idesc_t, kk > 0 ? 1u : 0u, leader);
#pragma unroll
          for (int kk = 0; kk < 8; ++kk)  // dV
            umma_bf16_w(acc + 2 * a.hd, desc_advance(p_t, kk * 2048), desc_advance(bd, kk * 2048), idesc_t, kk > 0 ? 1u : 0u, leader);
        }
        umma_commit_w(&acc_full[ab], leader);
        umma_commit_w(pds_empty, leader);
      }
      umma_commit_w(t_empty, leader);
    }
  } else if (warp >= 4) {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int slot = row / a.HP;
    const int i = row - slot * a.HP;
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    constexpr int nch = 2;
    uint32_t it = 0, n = 0;
    auto drain = [&](uint32_t nn, long long seq, int h) {  // [dQ_h | dK_h | dV_h] of head h -> dqkv rows (bf16)
      const uint32_t ab = a.nacc == 2 ? (nn & 1) : 0;
      const uint32_t use = a.nacc == 2 ? (nn >> 1) : nn;
      mbar_wait(&acc_full[ab], use & 1);
      tc_fence_after();
      const bool wr = seq < a.nseq && i < a.H;
      bf16* dst = a.dqkv + (seq * a.H + i) * a.lddqkv + h * a.hd;
      for (int part = 0; part < 3; ++part)
        for (int c = 0; c < a.hd / 16; ++c) {
          float o[32];
          // 16 columns at a time (x32 would run past a 16-wide head); the upper half of o is unused
          uint32_t* r = reinterpret_cast<uint32_t*>(o);
          asm volatile(
              "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
              : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
              : "r"(lane_base + BACC_COL + ab * 3 * a.hd + part * a.hd + c * 16)
              : "memory");
          tmem_wait_ld();
          if (wr) {
            uint4 u0, u1;
            u0.x = pack_bf16x2(o[0], o[1]); u0.y = pack_bf16x2(o[2], o[3]); u0.z = pack_bf16x2(o[4], o[5]); u0.w = pack_bf16x2(o[6], o[7]);
            u1.x = pack_bf16x2(o[8], o[9]); u1.y = pack_bf16x2(o[10], o[11]); u1.z = pack_bf16x2(o[12], o[13]); u1.w = pack_bf16x2(o[14], o[15]);
            bf16* d2 = dst + part * a.D + c * 16;
            if ((reinterpret_cast<uintptr_t>(d2) & 31) == 0) {
              st_global_32B(d2, u0.x, u0.y, u0.z, u0.w, u1.x, u1.y, u1.z, u1.w);
            } else {
              *reinterpret_cast<uint4*>(d2) = u0;
              *reinterpret_cast<uint4*>(d2 + 8) = u1;
            }
          }
        }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[ab]);
    };
    long long prev_seq = 0;
    int prev_h = -1;
    for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x, ++it) {
      const long long seq = (long long)tile * spt + slot;
      for (int h = 0; h < a.heads; ++h, ++n) {
        mbar_wait(sdp_full, n & 1);
        tc_fence_after();
        float s[2][32], dp[2][32];  // H <= 64: this row's 64-key block
        float mx = -INFINITY;
#pragma unroll
        for (int c = 0; c < 2; ++c)
          if (c < nch) {
            tmem_ld32(lane_base + BS_COL + slot * a.HP + c * 32, s[c]);
            tmem_ld32(lane_base + BDP_COL + slot * a.HP + c * 32, dp[c]);
            tmem_wait_ld();
#pragma unroll
            for (int k = 0; k < 32; ++k) {
              if (c * 32 + k >= a.H) s[c][k] = -INFINITY;
              mx = fmaxf(mx, s[c][k]);
            }
          }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(sdp_empty);
        const float ms = mx * a.scale_log2;
        float sum = 0.f;
#pragma unroll
        for (int c = 0; c < 2; ++c)
          if (c < nch) {
#pragma unroll
            for (int k = 0; k < 32; ++k) {
              s[c][k] = ex2f(fmaf(s[c][k], a.scale_log2, -ms));
              sum += s[c][k];
            }
          }
        const float inv = 1.f / sum;
        float delta = 0.f;
#pragma unroll
        for (int c = 0; c < 2; ++c)
          if (c < nch) {
#pragma unroll
            for (int k = 0; k < 32; ++k) {
              s[c][k] *= inv;                         // P
              delta = fmaf(s[c][k], dp[c][k], delta);
            }
          }
#pragma unroll
        for (int c = 0; c < 2; ++c)
          if (c < nch) {
#pragma unroll
            for (int k = 0; k < 32; ++k) dp[c][k] = s[c][k] * (dp[c][k] - delta) * a.scale;  // dS
          }
        // the previous head's gradient UMMAs must have finished reading the P / dS tiles
        mbar_wait(pds_empty, (n & 1) ^ 1);
#pragma unroll
        for (int c = 0; c < 2; ++c)
          if (c < nch) {
            const int key0 = slot * a.HP + c * 32;  // first key column of this chunk inside the 128-key tile
            uint8_t* pbox = sp + (key0 >> 6) * 16384;
            uint8_t* dbox = sds + (key0 >> 6) * 16384;
            const uint32_t ch0 = (key0 & 63) >> 3;
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              uint4 up, ud;
              up.x = pack_bf16x2(s[c][8 * g + 0], s[c][8 * g + 1]); up.y = pack_bf16x2(s[c][8 * g + 2], s[c][8 * g + 3]);
              up.z = pack_bf16x2(s[c][8 * g + 4], s[c][8 * g + 5]); up.w = pack_bf16x2(s[c][8 * g + 6], s[c][8 * g + 7]);
              ud.x = pack_bf16x2(dp[c][8 * g + 0], dp[c][8 * g + 1]); ud.y = pack_bf16x2(dp[c][8 * g + 2], dp[c][8 * g + 3]);
              ud.z = pack_bf16x2(dp[c][8 * g + 4], dp[c][8 * g + 5]); ud.w = pack_bf16x2(dp[c][8 * g + 6], dp[c][8 * g + 7]);
              *reinterpret_cast<uint4*>(pbox + sw128_offset(row, ch0 + g)) = up;
              *reinterpret_cast<uint4*>(dbox + sw128_offset(row, ch0 + g)) = ud;
            }
          }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(pds_full);
        if (prev_h >= 0) drain(n - 1, prev_seq, prev_h);  // overlaps this head's gradient UMMAs
        prev_seq = seq;
        prev_h = h;
      }
    }
    if (prev_h >= 0) drain(n - 1, prev_seq, prev_h);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace

bool attn_fwd_tc_supported(long long H, long long D, long long heads, long long ld, long long ldo, const void* qkv,
                           const void* out) {
  if (heads <= 0 || D % heads) return false;
  const long long hd = D / heads;
  return H >= 1 && H <= 128 && (D == 64 || D == 128) && (hd == 16 || hd == 32 || hd == 64) && (ld % 8) == 0 &&
         (ldo % 8) == 0 && ((uintptr_t)qkv % 16) == 0 && ((uintptr_t)out % 16) == 0;
}

int attn_fwd_tc(const void* qkv, long long ld, long long nseq, long long H, long long D, long long heads, long long q_rows,
                void* out, long long ldo, cudaStream_t stream) {
  AttnTcArgs a;
  a.nseq = (int)nseq; a.H = (int)H; a.D = (int)D; a.heads = (int)heads; a.hd = (int)(D / heads); a.q_rows = (int)q_rows;
  a.HP = H <= 64 ? 64 : 128;
  const int spt = 128 / a.HP;
  a.ntiles = (int)((nseq + spt - 1) / spt);
  a.out = (bf16*)out; a.ldo = ldo;
  a.scale_log2 = LOG2E_A / sqrtf((float)a.hd);
  CUtensorMap tm;
  int rc = make_tmap_bf16(&tm, qkv, 3 * D, nseq * H, ld, 64, (uint32_t)H);
  if (rc) return rc;
  const int smem = 2 * (int)(3 * D / 64) * 16384 + 1024 + 256;
  static bool configured = false;
  if (!configured) {
    TT_CUDA(cudaFuncSetAttribute(attn_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 6 * 16384 + 1024 + 256));
    configured = true;
  }
  const int grid = a.ntiles < num_sms() ? a.ntiles : num_sms();
  KernelSpan span("attn_fwd_tc_kernel", stream);
  attn_fwd_tc_kernel<<<grid, 256, smem, stream>>>(tm, a);
  TT_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

}  // namespace tt

namespace tt {

bool attn_bwd_tc_supported(long long H, long long D, long long heads, long long ld, long long lddo, long long lddqkv,
                           const void* qkv, const void* dout, const void* dqkv) {
  if (!attn_fwd_tc_supported(H, D, heads, ld, lddo, qkv, dout)) return false;
  if (H > 64) return false;  // one thread keeps its row's S and dP blocks (2 x 64 values) in registers
  return (lddqkv % 8) == 0 && ((uintptr_t)dqkv % 16) == 0;
}

int attn_bwd_tc(const void* qkv, long long ld, const void* dout, long long lddo, long long nseq, long long H, long long D,
                long long heads, long long q_rows, void* dqkv, long long lddqkv, cudaStream_t stream) {
  AttnBwdArgs a;
  a.nseq = (int)nseq; a.H = (int)H; a.D = (int)D; a.heads = (int)heads; a.hd = (int)(D / heads); a.q_rows = (int)q_rows;
  a.HP = H <= 64 ? 64 : 128;
  const int spt = 128 / a.HP;
  a.ntiles = (int)((nseq + spt - 1) / spt);
  a.nacc = (256 + 2 * 3 * a.hd <= 512) ? 2 : 1;
  a.dqkv = (bf16*)dqkv; a.lddqkv = lddqkv;
  a.scale = 1.f / sqrtf((float)a.hd);
  a.scale_log2 = LOG2E_A * a.scale;
  CUtensorMap tq, td;
  int rc = make_tmap_bf16(&tq, qkv, 3 * D, nseq * H, ld, 64, (uint32_t)H);
  if (rc) return rc;
  rc = make_tmap_bf16(&td, dout, D, nseq * q_rows, lddo, 64, (uint32_t)q_rows);
  if (rc) return rc;
  const int smem = (int)(4 * D / 64) * 16384 + 65536 + 1024 + 256;
  static bool configured = false;
  if (!configured) {
    TT_CUDA(cudaFuncSetAttribute(attn_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 16384 + 65536 + 1024 + 256));
    configured = true;
  }
  const int grid = a.ntiles < num_sms() ? a.ntiles : num_sms();
  KernelSpan span("attn_bwd_tc_kernel", stream);
  attn_bwd_tc_kernel<<<grid, 256, smem, stream>>>(tq, td, a);
  TT_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

}  // namespace tt
