// Brute-force maximum-inner-product search for sm_100a:  top-k of Q C^T per query row.
//
// Reference semantics: src/baseline_mips_module.py:57-61 (torch.topk(torch.matmul(query, corpus.T), k)),
// indices int64, scores sorted descending.  The [Q, C] score matrix (262 GB at BASELINE config 4) never
// exists: 128 x 128 score tiles are produced by tcgen05.mma into TMEM from TMA-staged bf16 tiles and are
// filtered in place against a per-row running threshold.
//
// Two kernels:
//   1. mips_screen_kernel (persistent, one CTA per SM, 640 threads)
//        warp 0 TMA producer, warp 1 UMMA issuer, warp 2 TMEM allocator, warps 4-19 epilogue.
//        A CTA holds TWO 128-row query tiles; the corpus tile stream (128 rows per stage) is shared by both.
//        Each query tile has TWO accumulator buffers (2 x 2 x 128 = the 512 TMEM columns), tile t goes to buffer
//        t mod 2, and each (query tile, buffer, TMEM lane quarter) has its own epilogue warp: a query row is
//        scanned by two threads - one for the even, one for the odd corpus tiles - each with its own threshold
//        tau (a register) and its own candidate list (<= 512 packed (score, index) keys in an L2-resident
//        scratch area).  A warp therefore has two tile periods for one tile, and the UMMAs of the next tile run
//        while it reads.  Per 32-column chunk the thread reduces its 32 scores to four group maxima and compares
//        once; a row with a candidate parks the chunk in its own shared-memory slot (no vote, no cross-lane
//        traffic) and filters it after the accumulator has been handed back.  A full list is cut back to its
//        best `kp` entries by a warp-wide bisection on the score bits (no sort in the scan).
//        All CTAs walk the corpus from the same end at the same pace, so a corpus tile is fetched from
//        HBM once per wave and served to the other SMs from L2.
//        What was measured on the way (profiles/r02_summary.md): the 128 x 256 single-buffer layout of round 1
//        was bound by the accumulator read phase (hit-free ceiling 1034 TFLOP/s) and a warp-cooperative hit
//        queue (3.4 k cycles per tile); with double buffering the hit-free ceiling is 1390 TFLOP/s and the
//        remaining cost is the serial latency of the rare-hit path of a warp (about 7 cycles per instruction
//        with 5 warps per sub-partition), which is why it is lane-local, branch-light and off the read phase.
//   2. mips_finalize_kernel (one warp per query): merges the sorted per-part candidate lists (bitonic merge), re-scores the
//        best kp = k + margin candidates in fp32 against the fp32 corpus (the reference's arithmetic),
//        sorts by (score desc, index asc) and writes the top k.
// Ordering rule for exact ties: ascending corpus index (torch.topk leaves it unspecified).
#include <cstdio>
#include <vector>
#include "common.cuh"
#include "kernels.h"

namespace tt {

namespace {

constexpr int LCAP = 512;   // candidate-list capacity per query row (keys)
constexpr int KP_MAX = 256; // screening depth limit (k + margin)
static_assert(LCAP == 2 * KP_MAX, "the finalize merge lays two KP_MAX-key runs side by side");
constexpr int QT = 128;     // query rows per tile
constexpr int CN = 128;     // corpus rows per tile (UMMA N)
constexpr int NBUF = 2;     // accumulator buffers per query tile; 2 query tiles x NBUF x CN = the 512 TMEM columns
#ifndef TT_MIPS_CW
#define TT_MIPS_CW 32
#endif
constexpr int CW = TT_MIPS_CW;  // score columns per TMEM load / compare / parking slot (16: next load in flight during the compare)
static_assert(CW == 16 || CW == 32, "chunk width");
#ifndef TT_MIPS_TRIG
#define TT_MIPS_TRIG 128  // new keys between two threshold refreshes of a list (screen kernel 24.6 / 23.9 / 24.3 / 24.8 / 25.0 ms at 96 / 128 / 160 / 208 / 252)
#endif
#ifndef TT_MIPS_SLEEP
#define TT_MIPS_SLEEP 0
#endif
// query tiles per CTA: two for d <= 128; one for 128 < d <= 256 (the operand tiles are twice as large)
constexpr int ne_of(int DP) { return DP <= 128 ? 2 : 1; }

// ---- packed keys: (order-preserving score bits << 32) | ~index ; larger key = better candidate ----------
__device__ __forceinline__ uint32_t f2ord(float f) {
  f += 0.0f;  // -0 -> +0
  uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(uint32_t o) {
  return __uint_as_float((o & 0x80000000u) ? (o & 0x7FFFFFFFu) : ~o);
}
__device__ __forceinline__ unsigned long long make_key(float s, uint32_t idx) {
  return ((unsigned long long)f2ord(s) << 32) | (unsigned long long)(~idx);
}
__device__ __forceinline__ uint32_t key_idx(unsigned long long k) { return ~(uint32_t)k; }

// explicit shared-window accesses (the parking slots are addressed by 32-bit shared addresses, not generic pointers)
__device__ __forceinline__ void sts128(uint32_t addr, float x, float y, float z, float w) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 r;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(addr) : "memory");
  return r;
}
__device__ __forceinline__ float lds32(uint32_t addr) {
  float r;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(r) : "r"(addr) : "memory");
  return r;
}

// barrier wait of the producer roles: a short sleep between probes leaves the issue slots of the sub-partition to the
// epilogue warps that share it
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t* bar, uint32_t parity) {
  if (TT_MIPS_SLEEP == 0) { mbar_wait(bar, parity); return; }
  const uint32_t addr = smem_u32(bar);
  if (mbar_try_wait_addr(addr, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait_addr(addr, parity)) {
    __nanosleep(TT_MIPS_SLEEP);
    if (clock64() - t0 > TT_MBAR_TIMEOUT_CYCLES) __trap();
  }
}

#ifdef TT_MIPS_BRINGUP  // clock stamps of one epilogue warp of CTA 0 (tools/mips_trace.py) and partial epilogues
#define MIPS_STAMP(slot_, val_)                                                                         \
  do {                                                                                                  \
    if (a.trace != nullptr && blockIdx.x == 0 && warp == 4 && lane == 0 && u == 0)                      \
      a.trace[(size_t)(j - j0) * 16 + (slot_)] = (val_);                                                \
  } while (0)
#define MIPS_DBG(bit_) (a.dbg & (bit_))
#else
#define MIPS_STAMP(slot_, val_) do { } while (0)
#define MIPS_DBG(bit_) 0
#endif

// Bitonic sort (descending) of n keys (power of two, 64 <= n <= 512) held in shared memory, by one warp.
__device__ void warp_sort_desc(unsigned long long* s, int n, int lane) {
  for (int k = 2; k <= n; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = lane; t < (n >> 1); t += 32) {
        const int lo = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        const int hi = lo | j;
        const unsigned long long a = s[lo], b = s[hi];
        const bool desc = (lo & k) == 0;
        if (desc ? (a < b) : (a > b)) {
          s[lo] = b;
          s[hi] = a;
        }
      }
      __syncwarp();
    }
  }
}

// Merge step of the bitonic network: n keys, first half descending and second half ascending (or any bitonic order)
// -> all n descending.  log2(n) passes instead of the log2(n) (log2(n) + 1) / 2 of a full sort.
__device__ void warp_bitonic_merge_desc(unsigned long long* s, int n, int lane) {
  for (int j = n >> 1; j > 0; j >>= 1) {
    for (int t = lane; t < (n >> 1); t += 32) {
      const int lo = ((t & ~(j - 1)) << 1) | (t & (j - 1));
      const int hi = lo | j;
      const unsigned long long a = s[lo], b = s[hi];
      if (a < b) {
        s[lo] = b;
        s[hi] = a;
      }
    }
    __syncwarp();
  }
}

struct ScreenArgs {
  int nq, nc, kp;
  int NE;        // query tiles per CTA (= ne_of(DP))
  int G;         // grid size
  int R;         // full rounds: q-blocks [0, R*G) are scanned over the whole corpus
  int r, S;      // tail: r q-blocks, each split into S corpus parts
  int CT;        // corpus tiles
  unsigned long long* lists;  // [G][NE][NBUF][128][LCAP] scratch
  unsigned long long* part;   // [slots][NE][NBUF][128][kp] per-unit results (sorted, zero padded)
#ifdef TT_MIPS_BRINGUP
  int dbg;
  long long* trace;  // [tile][16] clock stamps of one epilogue warp of CTA 0
#endif
};

template <int DP>
struct ScreenCfg {
  static constexpr int NE = ne_of(DP);
  static constexpr int EPI_WARPS = 4 * NE * NBUF;  // one warp per (query tile, accumulator buffer, TMEM lane quarter)
  static constexpr int THREADS = 128 + 32 * EPI_WARPS;
  static constexpr int TMEM_COLS = NE * NBUF * CN;
  static constexpr int KBOX = DP / 64;
  static constexpr int Q_BYTES = QT * DP * 2;   // one query tile
  static constexpr int C_BYTES = CN * DP * 2;   // one corpus stage
  static constexpr int STAGES = DP == 64 ? 4 : (DP == 128 ? 3 : 2);
  static constexpr int SORT_BYTES = EPI_WARPS * LCAP * 8;  // one 4 KB scratch per epilogue warp
  static constexpr int TAG_BYTES = 0;
  static constexpr int SMEM_BYTES = NE * Q_BYTES + STAGES * C_BYTES + SORT_BYTES + TAG_BYTES + 1024 + 256;
  static_assert(SMEM_BYTES <= 232448, "shared memory budget");
};

// Cut the list of `row` (n keys in `list`) back to its best entries: finds by bisection on the score bits
// the largest threshold t with count(score >= t) >= kp, keeps those keys (>= kp of them; more only on exact
// score ties), returns the new count and threshold.  Warp-cooperative; falls back to an exact sort when ties
// would leave the list too full.
__device__ __noinline__ void warp_compact(unsigned long long* list, int n, int kp, float tau_in, unsigned long long* sscr, int lane,
                             int& n_out, float& tau_out) {
  unsigned long long k[LCAP / 32];
#pragma unroll
  for (int i = 0; i < LCAP / 32; ++i) {
    const int e = i * 32 + lane;
    k[i] = e < n ? list[e] : 0ull;
  }
  uint32_t mx = 0;
#pragma unroll
  for (int i = 0; i < LCAP / 32; ++i) mx = max(mx, (uint32_t)(k[i] >> 32));
  mx = __reduce_max_sync(0xffffffffu, mx);
  // invariant: count(score >= lo) >= kp, count(score >= hi) < kp
  // every key of the list scores >= tau_in (the threshold it was collected under), so the search can start there
  uint32_t lo = (n >= kp && tau_in > -INFINITY) ? f2ord(tau_in) : 1u, hi = mx + 1u;
  if (lo >= hi) lo = 1u;
  int clo = n;
  if (mx == 0xffffffffu) hi = mx;  // NaN-ish garbage; keep it bounded
  while (hi - lo > 1u && clo > kp + 32) {
    const uint32_t mid = lo + ((hi - lo) >> 1);
    int c = 0;
#pragma unroll
    for (int i = 0; i < LCAP / 32; ++i) c += ((uint32_t)(k[i] >> 32) >= mid) ? 1 : 0;
    c = __reduce_add_sync(0xffffffffu, c);
    if (c >= kp) { lo = mid; clo = c; } else { hi = mid; }
  }
  if (clo <= LCAP - 192) {
    int off = 0;
#pragma unroll
    for (int i = 0; i < LCAP / 32; ++i) {
      const bool keep = (uint32_t)(k[i] >> 32) >= lo;
      const uint32_t b = __ballot_sync(0xffffffffu, keep);
      if (keep) list[off + __popc(b & ((1u << lane) - 1u))] = k[i];
      off += __popc(b);
    }
    __syncwarp();
    n_out = off;
    tau_out = ord2f(lo);
    return;
  }
  // heavy ties (rare): exact (score, index) order, sorted in place in the list's own (global, L2-resident) storage -
  // the warp's shared-memory scratch may hold parked rows at this point and must not be touched
  (void)sscr;
#pragma unroll
  for (int i = 0; i < LCAP / 32; ++i) list[i * 32 + lane] = k[i];
  __syncwarp();
  warp_sort_desc(list, LCAP, lane);
  n_out = kp;
  tau_out = ord2f((uint32_t)(list[kp - 1] >> 32));
  __syncwarp();
}

template <int DP>
__global__ void __launch_bounds__(ScreenCfg<DP>::THREADS, 1)
mips_screen_kernel(const __grid_constant__ CUtensorMap tmq, const __grid_constant__ CUtensorMap tmc, const ScreenArgs a) {
  using Cfg = ScreenCfg<DP>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sq = smem;                                   // 2 query tiles
  constexpr int NE = Cfg::NE;
  uint8_t* sc = smem + NE * Cfg::Q_BYTES;               // corpus ring
  unsigned long long* ssort = reinterpret_cast<unsigned long long*>(sc + Cfg::STAGES * Cfg::C_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(ssort) + Cfg::SORT_BYTES + Cfg::TAG_BYTES);
  uint64_t* q_full = bars;          // [1]
  uint64_t* q_empty = bars + 1;     // [1]
  uint64_t* d_full = bars + 2;      // [2][NBUF]
  uint64_t* d_empty = bars + 6;     // [2][NBUF]
  uint64_t* c_full = bars + 10;     // [STAGES]
  uint64_t* c_empty = c_full + Cfg::STAGES;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(c_empty + Cfg::STAGES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmq);
    tma_prefetch_desc(&tmc);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(q_full, 1);
    mbar_init(q_empty, 1);
    for (int i = 0; i < NE * NBUF; ++i) {
      mbar_init(&d_full[i], 1);
      mbar_init(&d_empty[i], 4);
    }
    for (int i = 0; i < Cfg::STAGES; ++i) {
      mbar_init(&c_full[i], 1);
      mbar_init(&c_empty[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_holder, Cfg::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  // unit enumeration shared by all roles: u-th unit of this CTA -> (q-block, first tile, last tile, slot)
  const int n_full = a.R;                                                   // one unit per full round
  const int n_tail = (a.r * a.S > (int)blockIdx.x) ? ((a.r * a.S - 1 - (int)blockIdx.x) / a.G + 1) : 0;
  const int n_units = n_full + n_tail;
  auto unit = [&](int u, int& qb, int& j0, int& j1, int& slot) {
    if (u < n_full) {
      qb = u * a.G + blockIdx.x;
      j0 = 0; j1 = a.CT;
      slot = qb;
    } else {
      const int t = (u - n_full) * a.G + blockIdx.x;  // tail unit id, part-major
      const int p = t / a.r;
      qb = a.R * a.G + t % a.r;
      j0 = (int)((long long)a.CT * p / a.S);
      j1 = (int)((long long)a.CT * (p + 1) / a.S);
      slot = a.R * a.G + (t % a.r) * a.S + p;
    }
  };

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int u = 0; u < n_units; ++u) {
        int qb, j0, j1, slot;
        unit(u, qb, j0, j1, slot);
        mbar_wait(q_empty, (u & 1) ^ 1);
        mbar_arrive_expect_tx(q_full, NE * Cfg::Q_BYTES);
#pragma unroll
        for (int e = 0; e < NE; ++e)
#pragma unroll
          for (int b = 0; b < Cfg::KBOX; ++b)
            tma_load_2d(sq + e * Cfg::Q_BYTES + b * (QT * 128), &tmq, q_full, b * 64, (qb * NE + e) * QT);
        for (int j = j0; j < j1; ++j) {
          mbar_wait_relaxed(&c_empty[stage], phase ^ 1);
          mbar_arrive_expect_tx(&c_full[stage], Cfg::C_BYTES);
          uint8_t* dst = sc + stage * Cfg::C_BYTES;
#pragma unroll
          for (int b = 0; b < Cfg::KBOX; ++b) tma_load_2d(dst + b * (CN * 128), &tmc, &c_full[stage], b * 64, j * CN);
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    {  // the whole warp walks the schedule; only the elected lane issues tcgen05 instructions
      const uint32_t leader = elect_one();
      constexpr uint32_t idesc = make_idesc_bf16(QT, CN, 0, 0);
      int stage = 0;
      uint32_t phase = 0, t = 0;
      const uint64_t dq0 = make_smem_desc_sw128(smem_u32(sq), 0, 1024);
      const uint64_t dc0 = make_smem_desc_sw128(smem_u32(sc), 0, 1024);
      for (int u = 0; u < n_units; ++u) {
        int qb, j0, j1, slot;
        unit(u, qb, j0, j1, slot);
        mbar_wait(q_full, u & 1);
        for (int j = j0; j < j1; ++j, ++t) {
          mbar_wait(&c_full[stage], phase);
          const uint64_t dc = desc_advance(dc0, stage * Cfg::C_BYTES);
#pragma unroll
          const uint32_t buf = t % NBUF, use = t / NBUF;  // accumulator buffer of this tile and its use count
#pragma unroll
          for (int e = 0; e < NE; ++e) {
            mbar_wait_relaxed(&d_empty[e * NBUF + buf], (use & 1) ^ 1);
            tc_fence_after();
#pragma unroll
            for (int k = 0; k < DP / 16; ++k)
              umma_bf16_w(tmem_base + (e * NBUF + buf) * CN,
                          desc_advance(dq0, e * Cfg::Q_BYTES + (k >> 2) * (QT * 128) + (k & 3) * 32),
                          desc_advance(dc, (k >> 2) * (CN * 128) + (k & 3) * 32), idesc, k > 0 ? 1u : 0u, leader);
            umma_commit_w(&d_full[e * NBUF + buf], leader);
          }
          umma_commit_w(&c_empty[stage], leader);
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit_w(q_empty, leader);
      }
    }
  } else if (warp >= 4) {
    const int e = (warp - 4) / (4 * NBUF);  // query tile of the CTA
    const int par = ((warp - 4) >> 2) & 1;  // accumulator buffer = parity of the corpus tiles this warp scans
    const int q = warp & 3;         // TMEM lane quarter
    const int row = q * 32 + lane;  // row of the query tile owned by this thread
    unsigned long long* sscr = ssort + (warp - 4) * LCAP;  // unit-end sort scratch of this warp
    // this warp's 32 lists: [cta][e][row][LCAP]; a thread appends to the list of its own query row
    unsigned long long* wlists = a.lists + ((((size_t)blockIdx.x * NE + e) * NBUF + par) * QT + q * 32) * LCAP;
    unsigned long long* mylist = wlists + (size_t)lane * LCAP;
    const uint32_t park = smem_u32(sscr) + lane * (CW * 4);  // this row's parking slot in the warp's scratch
    const int sw = CW == 32 ? (lane & 7) : ((lane >> 1) & 3);  // 16-byte pieces of a slot are XOR-swizzled against bank conflicts
    // list length that triggers a threshold refresh (checked once per tile); a tile adds at most CN keys to a list of
    // fewer than max(trig, LCAP - 191) keys, which stays within LCAP
    const int trig = (a.kp + TT_MIPS_TRIG < LCAP - CN) ? a.kp + TT_MIPS_TRIG : LCAP - CN;
    uint32_t t = 0;
    for (int u = 0; u < n_units; ++u) {
      int qb, j0, j1, slot;
      unit(u, qb, j0, j1, slot);
      float tau = -INFINITY;
      int cnt = 0;
      for (int j = j0; j < j1; ++j, ++t) {
        const uint32_t buf = t % NBUF, use = t / NBUF;
        if ((int)buf != par) continue;  // the tiles of the other buffer belong to the sibling warp (own lists, own threshold)
        MIPS_STAMP(0, clock64());
        mbar_wait_relaxed(&d_full[e * NBUF + buf], use & 1);
        tc_fence_after();
        MIPS_STAMP(1, clock64());
        const uint32_t tcol = tmem_base + ((uint32_t)(q * 32) << 16) + (e * NBUF + buf) * CN;
        const bool ragged = (j + 1) * CN > a.nc;  // last corpus tile: columns past the corpus end score 0 (TMA zero fill)
        // Read phase (holds the accumulator buffer): CW-column chunks of the thread's query row are reduced to the maxima
        // of their 8-column groups and compared against the row's threshold (CW = 16: with the TMEM load of the next
        // chunk in flight).  A row with a candidate PARKS the chunk in its own slot of the warp's scratch (no vote,
        // no cross-lane traffic).  Filtering a parked chunk is lane-local too: it happens when the row needs its
        // slot again within the tile (rare outside the first tiles) or, normally, after the buffer has been handed
        // back.
        constexpr int NG = CW / 8;  // groups per chunk
        unsigned long long* dst = mylist + cnt;
        int pend = -1;      // parked chunk of this row (-1: none)
        uint32_t pgrp = 0;  // its groups that beat the threshold
        auto drain_own = [&]() {
          // usually one group with one score above the threshold: a bit mask of the group's eight compares, then the
          // survivors are re-read from the slot by index
          const uint32_t col0 = (uint32_t)(j * CN + pend * CW);
          while (pgrp) {
            const int k = __ffs(pgrp) - 1;
            pgrp &= pgrp - 1;
            const float4 x0 = lds128(park + (((2 * k) ^ sw) << 4)), x1 = lds128(park + (((2 * k + 1) ^ sw) << 4));
            uint32_t m = (x0.x > tau ? 1u : 0u) | (x0.y > tau ? 2u : 0u) | (x0.z > tau ? 4u : 0u) | (x0.w > tau ? 8u : 0u) |
                         (x1.x > tau ? 16u : 0u) | (x1.y > tau ? 32u : 0u) | (x1.z > tau ? 64u : 0u) | (x1.w > tau ? 128u : 0u);
            while (m) {
              const int i = __ffs(m) - 1;
              m &= m - 1;
              *dst++ = make_key(lds32(park + ((((2 * k + (i >> 2)) ^ sw) << 4) | ((i & 3) << 2))), col0 + 8 * k + i);
            }
          }
        };
        auto scan = [&](float* v, int c) {
          if (ragged) {
#pragma unroll
            for (int i = 0; i < CW; ++i)
              if (j * CN + c * CW + i >= a.nc) v[i] = -INFINITY;
          }
          float g[NG];
#pragma unroll
          for (int k = 0; k < NG; ++k) {  // three 3-input maxima and one 2-input per group
            const float m0 = fmaxf(fmaxf(v[8 * k], v[8 * k + 1]), v[8 * k + 2]);
            const float m1 = fmaxf(fmaxf(v[8 * k + 3], v[8 * k + 4]), v[8 * k + 5]);
            g[k] = fmaxf(fmaxf(fmaxf(m0, m1), v[8 * k + 6]), v[8 * k + 7]);
          }
          float gm = fmaxf(g[0], g[1]);
          if (NG == 4) gm = fmaxf(fmaxf(gm, g[2]), g[3]);
          if (gm > tau && !MIPS_DBG(1)) {
            if (pend >= 0) drain_own();
            pend = c;
            pgrp = 0;
#pragma unroll
            for (int k = 0; k < NG; ++k) pgrp |= g[k] > tau ? (1u << k) : 0u;
#pragma unroll
            for (int i = 0; i < CW / 4; ++i) sts128(park + ((i ^ sw) << 4), v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
          }
        };
        if constexpr (CW == 16) {
          float va[16], vb[16];
          tmem_ld16(tcol, va);
#pragma unroll 1
          for (int c = 0; c < CN / 16; c += 2) {
            tmem_wait_ld();
            tmem_ld16(tcol + (c + 1) * 16, vb);
            MIPS_STAMP(4 + c, clock64());
            if (!MIPS_DBG(2)) scan(va, c);
            tmem_wait_ld();
            if (c + 2 < CN / 16) {
              tmem_ld16(tcol + (c + 2) * 16, va);
            } else {  // accumulator fully read: hand it back
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(&d_empty[e * NBUF + buf]);
            }
            MIPS_STAMP(5 + c, clock64());
            if (!MIPS_DBG(2)) scan(vb, c + 1);
          }
        } else {
          float v[32];
#pragma unroll 1
          for (int c = 0; c < CN / 32; ++c) {
            tmem_ld32(tcol + c * 32, v);
            tmem_wait_ld();
            MIPS_STAMP(4 + 2 * c, clock64());
            if (c == CN / 32 - 1) {  // accumulator fully read: hand it back
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(&d_empty[e * NBUF + buf]);
            }
            if (!MIPS_DBG(2)) scan(v, c);
          }
        }
        MIPS_STAMP(2, clock64());
        if (__any_sync(0xffffffffu, pend >= 0)) {
          if (pend >= 0) drain_own();
          cnt = (int)(dst - mylist);
          MIPS_STAMP(12, clock64());
          // lists close to their capacity are cut back to their best entries (threshold refresh); rare
          uint32_t full = __ballot_sync(0xffffffffu, cnt >= trig);
          MIPS_STAMP(13, clock64());
          MIPS_STAMP(14, full);
          while (full) {
            const int rr = __ffs(full) - 1;
            full &= full - 1;
            __syncwarp();  // the owner's appends are visible to the whole warp
            int cnt_r = __shfl_sync(0xffffffffu, cnt, rr);
            float new_tau = __shfl_sync(0xffffffffu, tau, rr);
            warp_compact(wlists + (size_t)rr * LCAP, cnt_r, a.kp, new_tau, sscr, lane, cnt_r, new_tau);
            if (lane == rr) { cnt = cnt_r; tau = new_tau; }
          }
        }
        MIPS_STAMP(3, clock64());
      }
      __syncwarp();
      // unit done: exact order of each row's list, best kp keys -> part[slot][e][row][kp]
      for (int r = 0; r < 32; ++r) {
        const int n = __shfl_sync(0xffffffffu, cnt, r);
        const unsigned long long* list = wlists + (size_t)r * LCAP;
        const int ns = n <= LCAP / 2 ? LCAP / 2 : LCAP;  // kp <= LCAP / 2
        for (int i = lane; i < ns; i += 32) sscr[i] = i < n ? list[i] : 0ull;
        __syncwarp();
        warp_sort_desc(sscr, ns, lane);
        unsigned long long* dst = a.part + ((((size_t)slot * NE + e) * NBUF + par) * QT + q * 32 + r) * a.kp;
        for (int i = lane; i < a.kp; i += 32) dst[i] = sscr[i];
        __syncwarp();
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------
// finalize: merge parts, fp32 re-score, exact sort, write top-k
// ------------------------------------------------------------------------------------------------
struct FinalArgs {
  int nq, nc, d, k, kp;
  int NE, G, R, r, S;
  const unsigned long long* part;
  const float* q32;
  long long ldq;
  const float* c32;
  long long ldc;
  long long* idx_out;
  float* score_out;
};

__global__ void __launch_bounds__(128) mips_finalize_kernel(const FinalArgs a) {
  __shared__ unsigned long long ssort_all[4][LCAP];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * 4 + warp;
  if (row >= a.nq) return;
  unsigned long long* s = ssort_all[warp];
  const int qtile = (int)(row / QT), qb = qtile / a.NE, e = qtile % a.NE, rr = (int)(row % QT);
  int slot0, nparts;
  if (qb < a.R * a.G) { slot0 = qb; nparts = 1; }
  else { slot0 = a.R * a.G + (qb - a.R * a.G) * a.S; nparts = a.S; }
  // merged best kp by screening score: every part is sorted (descending, zero padded).  The running best lives in
  // s[0, KP_MAX); the next part is laid behind it in REVERSE order, which makes the 2 KP_MAX keys a bitonic sequence,
  // and one merge pass set sorts them.
  for (int p = 0; p < nparts * NBUF; ++p) {
    const unsigned long long* src = a.part + ((((size_t)(slot0 + p / NBUF) * a.NE + e) * NBUF + p % NBUF) * QT + rr) * a.kp;
    if (p == 0) {
      for (int i = lane; i < KP_MAX; i += 32) s[i] = i < a.kp ? src[i] : 0ull;
    } else {
      for (int i = lane; i < KP_MAX; i += 32) s[LCAP - 1 - i] = i < a.kp ? src[i] : 0ull;
    }
    __syncwarp();
    if (p > 0) warp_bitonic_merge_desc(s, LCAP, lane);
  }
  // fp32 re-score of the kp candidates (the reference scores in fp32: src/baseline_mips_module.py:58)
  const float* qrow = a.q32 + row * a.ldq;
  for (int c = 0; c < a.kp; ++c) {
    const unsigned long long key = s[c];
    if (key == 0ull) continue;  // fewer than kp candidates (tiny corpus)
    const uint32_t ci = key_idx(key);
    const float* crow = a.c32 + (long long)ci * a.ldc;
    float acc = 0.f;
    for (int x = lane; x < a.d; x += 32) acc = fmaf(qrow[x], crow[x], acc);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) s[c] = make_key(acc, ci);
  }
  __syncwarp();
  for (int i = a.kp + lane; i < LCAP; i += 32) s[i] = 0ull;
  __syncwarp();
  warp_sort_desc(s, a.kp <= 64 ? 64 : (a.kp <= 128 ? 128 : 256), lane);
  for (int i = lane; i < a.k; i += 32) {
    const unsigned long long key = s[i];
    a.idx_out[row * a.k + i] = (long long)key_idx(key);
    a.score_out[row * a.k + i] = ord2f((uint32_t)(key >> 32));
  }
}

struct Plan {
  int DP, NE, kp, NQB, CT, G, R, r, S, slots;
};

static int make_plan(long long nq, long long nc, long long d, long long k, Plan& p) {
  p.DP = d <= 64 ? 64 : (d <= 128 ? 128 : 256);
  p.NE = ne_of(p.DP);
  long long kp = k + (k / 4 > 32 ? k / 4 : 32);
  if (kp > nc) kp = nc;
  if (kp > KP_MAX) kp = KP_MAX;
  p.kp = (int)kp;
  p.NQB = (int)((nq + p.NE * QT - 1) / (p.NE * QT));
  p.CT = (int)((nc + CN - 1) / CN);
  const int sms = num_sms();
  p.G = sms;
  p.R = p.NQB / sms;  // full rounds: every SM scans the whole corpus for one q-block
  p.r = p.NQB - p.R * sms;
  p.S = 1;
  if (p.r > 0) {
    if (p.r * 2 <= sms) {
      p.S = sms / p.r;  // one tail unit per CTA, corpus cut into S parts
      if (p.S > 64) p.S = 64;
    } else {
      // several tail units per CTA: minimise ceil(r S / G) / S over small S
      double best = 1e30;
      for (int s = 1; s <= 8; ++s) {
        const double cost = (double)((p.r * s + sms - 1) / sms) / s;
        if (cost < best - 1e-9) { best = cost; p.S = s; }
      }
    }
    if (p.S > p.CT) p.S = p.CT;
  }
  if (p.R == 0 && p.r * p.S < sms) p.G = p.r * p.S;  // small batch: fewer CTAs than SMs
  p.slots = p.R * p.G + p.r * p.S;
  return 0;
}

}  // namespace

size_t mips_workspace_bytes(long long Q, long long C, long long d, long long k) {
  Plan p;
  make_plan(Q, C, d, k, p);
  return (size_t)p.G * p.NE * NBUF * QT * LCAP * 8 + (size_t)p.slots * p.NE * NBUF * QT * p.kp * 8 + 1024;
}

template <int DP>
static int launch_screen(const CUtensorMap& tq, const CUtensorMap& tc, const ScreenArgs& a, cudaStream_t st) {
  using Cfg = ScreenCfg<DP>;
  static bool configured = false;
  if (!configured) {
    TT_CUDA(cudaFuncSetAttribute(mips_screen_kernel<DP>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    configured = true;
  }
  KernelSpan span("mips_screen_kernel", st);
  mips_screen_kernel<DP><<<a.G, Cfg::THREADS, Cfg::SMEM_BYTES, st>>>(tq, tc, a);
  TT_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

int mips_topk(const void* Q16, long long ldq, const void* C16, long long ldc, const float* Q32, long long ldq32,
              const float* C32, long long ldc32, long long nq, long long nc, long long d, long long k, long long* idx,
              float* scores, void* ws, size_t ws_bytes, cudaStream_t stream) {
  TT_CHECK(nq > 0 && nc > 0 && d > 0, "mips_topk: empty problem");
  TT_CHECK(k > 0 && k <= nc, "mips_topk: selected index k out of range (k=%lld, corpus %lld)", k, nc);
  TT_CHECK(d <= 256, "mips_topk: embedding dim %lld > 256 is not supported by the fused kernel", d);
  TT_CHECK(k <= KP_MAX - 32, "mips_topk: k=%lld > %d is not supported by the fused kernel", k, KP_MAX - 32);
  TT_CHECK(nc < (1ll << 32) - 1, "mips_topk: corpus too large for 32-bit row indices");
  TT_CHECK((ldq % 8) == 0 && (ldc % 8) == 0 && ((uintptr_t)Q16 % 16) == 0 && ((uintptr_t)C16 % 16) == 0,
           "mips_topk: bf16 operands need 16-byte aligned rows");
  Plan p;
  make_plan(nq, nc, d, k, p);
  TT_CHECK(ws_bytes >= mips_workspace_bytes(nq, nc, d, k), "mips_topk: workspace too small");
  ScreenArgs a;
  a.nq = (int)nq; a.nc = (int)nc; a.kp = p.kp; a.NE = p.NE;
  a.G = p.G; a.R = p.R; a.r = p.r; a.S = p.S; a.CT = p.CT;
  a.lists = (unsigned long long*)ws;
#ifdef TT_MIPS_BRINGUP
  a.dbg = getenv("TT_MIPS_DBG") ? atoi(getenv("TT_MIPS_DBG")) : 0;
  a.trace = nullptr;
  static long long* trace_dev = nullptr;
  if (getenv("TT_MIPS_TRACE")) {
    if (!trace_dev) cudaMalloc(&trace_dev, (size_t)p.CT * 16 * 8);
    cudaMemset(trace_dev, 0, (size_t)p.CT * 16 * 8);
    a.trace = trace_dev;
  }
#endif
  a.part = a.lists + (size_t)p.G * p.NE * NBUF * QT * LCAP;
  CUtensorMap tq, tc;
  int rc = make_tmap_bf16(&tq, Q16, d, nq, ldq, 64, QT);
  if (rc) return rc;
  rc = make_tmap_bf16(&tc, C16, d, nc, ldc, 64, CN);
  if (rc) return rc;
  rc = p.DP == 64 ? launch_screen<64>(tq, tc, a, stream)
                  : (p.DP == 128 ? launch_screen<128>(tq, tc, a, stream) : launch_screen<256>(tq, tc, a, stream));
  if (rc) return rc;
  FinalArgs f;
  f.nq = (int)nq; f.nc = (int)nc; f.d = (int)d; f.k = (int)k; f.kp = p.kp;
  f.NE = p.NE; f.G = p.G; f.R = p.R; f.r = p.r; f.S = p.S;
  f.part = a.part;
  f.q32 = Q32; f.ldq = ldq32; f.c32 = C32; f.ldc = ldc32;
  f.idx_out = idx; f.score_out = scores;
  KernelSpan span("mips_finalize_kernel", stream);
  mips_finalize_kernel<<<(unsigned)((nq + 3) / 4), 128, 0, stream>>>(f);
#ifdef TT_MIPS_BRINGUP
  if (a.trace) {
    cudaStreamSynchronize(stream);
    std::vector<long long> h((size_t)p.CT * 16);
    cudaMemcpy(h.data(), a.trace, h.size() * 8, cudaMemcpyDeviceToHost);
    FILE* fp = fopen(getenv("TT_MIPS_TRACE"), "wb");
    if (fp) { fwrite(h.data(), 8, h.size(), fp); fclose(fp); }
  }
#endif
  TT_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

}  // namespace tt
