// Brute-force maximum-inner-product search for sm_100a:  top-k of Q C^T per query row.
//
// Reference semantics: src/baseline_mips_module.py:57-61 (torch.topk(torch.matmul(query, corpus.T), k)),
// indices int64, scores sorted descending.  The [Q, C] score matrix (262 GB at BASELINE config 4) never
// exists: 128 x 256 score tiles are produced by tcgen05.mma into TMEM from TMA-staged bf16 tiles and are
// filtered in place against a per-row running threshold.
//
// Three kernels:
//   1. mips_screen_kernel (persistent, one CTA per SM, 384 threads)
//        warp 0 TMA producer, warp 1 UMMA issuer, warp 2 TMEM allocator,
//        warps 4-7 / 8-11: epilogue group 0 / 1.  A CTA holds TWO 128-row query tiles; the corpus tile
//        stream (256 rows per stage) is shared by both, group e owns the accumulator of query tile e
//        (TMEM columns [256 e, 256 e + 256)), so every query row belongs to exactly one thread, which keeps
//        that row's threshold tau in a register and its candidate list (<= 512 packed (score, index) keys)
//        in an L2-resident scratch area.  Per 32-column chunk the thread reduces its 32 scores with a max
//        tree and compares once; the rare chunks with a hit are appended warp-cooperatively.  A full list is
//        cut back to its best `kp` entries by a warp-wide bisection on the score bits (no sort in the scan).
//        All CTAs walk the corpus from the same end at the same pace, so a corpus tile is fetched from
//        HBM once per wave and served to the other SMs from L2.
//   2. mips_finalize_kernel (one warp per query): merges the per-part candidate lists, re-scores the
//        best kp = k + margin candidates in fp32 against the fp32 corpus (the reference's arithmetic),
//        sorts by (score desc, index asc) and writes the top k.
// Ordering rule for exact ties: ascending corpus index (torch.topk leaves it unspecified).
#include "common.cuh"
#include "kernels.h"

namespace tt {

namespace {

constexpr int LCAP = 512;   // candidate-list capacity per query row (keys)
constexpr int KP_MAX = 256; // screening depth limit (k + margin)
constexpr int QT = 128;     // query rows per tile
constexpr int CN = 256;     // corpus rows per tile (UMMA N)
constexpr int QCAP = 32;    // parked rows per warp queue (32 x 128 B = the warp's 4 KB scratch; tags live beside it)

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

struct ScreenArgs {
  int nq, nc, kp;
  int G;         // grid size
  int R;         // full rounds: q-blocks [0, R*G) are scanned over the whole corpus
  int r, S;      // tail: r q-blocks, each split into S corpus parts
  int CT;        // corpus tiles
  unsigned long long* lists;  // [G][2][128][LCAP] scratch
  unsigned long long* part;   // [slots][2][128][kp] per-unit results (best kp, unsorted beyond "top kp")
};

template <int DP>
struct ScreenCfg {
  static constexpr int KBOX = DP / 64;
  static constexpr int Q_BYTES = QT * DP * 2;   // one query tile
  static constexpr int C_BYTES = CN * DP * 2;   // one corpus stage
  static constexpr int STAGES = DP == 64 ? 4 : 2;
  static constexpr int SORT_BYTES = 8 * LCAP * 8;  // one 4 KB scratch per epilogue warp
  static constexpr int TAG_BYTES = 8 * QCAP * 4;   // (chunk, row) tag per parked row and epilogue warp
  static constexpr int SMEM_BYTES = 2 * Q_BYTES + STAGES * C_BYTES + SORT_BYTES + TAG_BYTES + 1024 + 256;
  static_assert(SMEM_BYTES <= 232448, "shared memory budget");
};

// Cut the list of `row` (n keys in `list`) back to its best entries: finds by bisection on the score bits
// the largest threshold t with count(score >= t) >= kp, keeps those keys (>= kp of them; more only on exact
// score ties), returns the new count and threshold.  Warp-cooperative; falls back to an exact sort when ties
// would leave the list too full.
__device__ void warp_compact(unsigned long long* list, int n, int kp, float tau_in, unsigned long long* sscr, int lane,
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
__global__ void __launch_bounds__(384, 1)
mips_screen_kernel(const __grid_constant__ CUtensorMap tmq, const __grid_constant__ CUtensorMap tmc, const ScreenArgs a) {
  using Cfg = ScreenCfg<DP>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sq = smem;                                   // 2 query tiles
  uint8_t* sc = smem + 2 * Cfg::Q_BYTES;                // corpus ring
  unsigned long long* ssort = reinterpret_cast<unsigned long long*>(sc + Cfg::STAGES * Cfg::C_BYTES);
  int* stags = reinterpret_cast<int*>(reinterpret_cast<uint8_t*>(ssort) + Cfg::SORT_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(ssort) + Cfg::SORT_BYTES + Cfg::TAG_BYTES);
  uint64_t* q_full = bars;          // [1]
  uint64_t* q_empty = bars + 1;     // [1]
  uint64_t* d_full = bars + 2;      // [2]
  uint64_t* d_empty = bars + 4;     // [2]
  uint64_t* c_full = bars + 6;      // [STAGES]
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
    for (int i = 0; i < 2; ++i) {
      mbar_init(&d_full[i], 1);
      mbar_init(&d_empty[i], 4);
    }
    for (int i = 0; i < Cfg::STAGES; ++i) {
      mbar_init(&c_full[i], 1);
      mbar_init(&c_empty[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_holder, 512);
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
        mbar_arrive_expect_tx(q_full, 2 * Cfg::Q_BYTES);
#pragma unroll
        for (int e = 0; e < 2; ++e)
#pragma unroll
          for (int b = 0; b < Cfg::KBOX; ++b)
            tma_load_2d(sq + e * Cfg::Q_BYTES + b * (QT * 128), &tmq, q_full, b * 64, (qb * 2 + e) * QT);
        for (int j = j0; j < j1; ++j) {
          mbar_wait(&c_empty[stage], phase ^ 1);
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
          for (int e = 0; e < 2; ++e) {
            mbar_wait(&d_empty[e], (t & 1) ^ 1);
            tc_fence_after();
#pragma unroll
            for (int k = 0; k < DP / 16; ++k)
              umma_bf16_w(tmem_base + e * CN, desc_advance(dq0, e * Cfg::Q_BYTES + (k >> 2) * (QT * 128) + (k & 3) * 32),
                          desc_advance(dc, (k >> 2) * (CN * 128) + (k & 3) * 32), idesc, k > 0 ? 1u : 0u, leader);
            umma_commit_w(&d_full[e], leader);
          }
          umma_commit_w(&c_empty[stage], leader);
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit_w(q_empty, leader);
      }
    }
  } else if (warp >= 4) {
    const int e = (warp - 4) >> 2;  // epilogue group <-> query tile <-> accumulator
    const int q = warp & 3;         // TMEM lane quarter
    const int row = q * 32 + lane;  // row of the query tile owned by this thread
    unsigned long long* sscr = ssort + (warp - 4) * LCAP;
    float* fscr = reinterpret_cast<float*>(sscr);                 // parked rows: QCAP x 32 floats
    int* qmeta = stags + (warp - 4) * QCAP;                       // (chunk << 5) | row tag per parked row
    unsigned long long* sscr2 = sscr;  // exact-sort fallback of warp_compact: only entered with an empty queue (see drain)
    // this warp's 32 lists: [cta][e][row][LCAP]
    unsigned long long* wlists = a.lists + (((size_t)blockIdx.x * 2 + e) * QT + q * 32) * LCAP;
    const int trig = (a.kp + 160 < LCAP - 64) ? a.kp + 160 : LCAP - 64;  // list length that triggers a threshold refresh
    uint32_t t = 0;
    for (int u = 0; u < n_units; ++u) {
      int qb, j0, j1, slot;
      unit(u, qb, j0, j1, slot);
      float tau = -INFINITY;
      int cnt = 0;
      for (int j = j0; j < j1; ++j, ++t) {
        mbar_wait(&d_full[e], t & 1);
        tc_fence_after();
        // Pass 1 (holds the accumulator): per 32-column chunk a max tree and one compare per row; rows with a
        // candidate are PARKED (their 32 scores + (row, chunk) tag) in the warp's shared-memory queue.
        // Pass 2 (after the accumulator has been handed back, i.e. off the UMMA critical path): the parked rows
        // are filtered against their thresholds and appended to the candidate lists.
        int nq_ = 0;  // parked rows in the queue (warp-uniform)
        auto drain = [&]() {
          const uint32_t below = (1u << lane) - 1u;
          for (int s0 = 0; s0 < nq_; s0 += 2) {  // two rows per step: their load / ballot chains overlap
            const int s1 = s0 + 1 < nq_ ? s0 + 1 : s0;
            const bool two = s1 != s0;
            const int tag0 = qmeta[s0], tag1 = qmeta[s1];
            const int r0 = tag0 & 31, r1 = tag1 & 31;
            const float x0 = fscr[s0 * 32 + lane], x1 = fscr[s1 * 32 + lane];
            const float t0 = __shfl_sync(0xffffffffu, tau, r0), t1 = __shfl_sync(0xffffffffu, tau, r1);
            const int c0 = __shfl_sync(0xffffffffu, cnt, r0);
            const int col0 = j * CN + (tag0 >> 5) * 32 + lane, col1 = j * CN + (tag1 >> 5) * 32 + lane;
            const bool p0 = (x0 > t0) && (col0 < a.nc);
            const uint32_t b0 = __ballot_sync(0xffffffffu, p0);
            if (p0) wlists[(size_t)r0 * LCAP + c0 + __popc(b0 & below)] = make_key(x0, (uint32_t)col0);
            if (lane == r0) cnt = c0 + __popc(b0);
            // the second row may be the same query row (another chunk of it): read its count after the update
            const int c1 = __shfl_sync(0xffffffffu, cnt, r1);
            const bool p1 = two && (x1 > t1) && (col1 < a.nc);
            const uint32_t b1 = __ballot_sync(0xffffffffu, p1);
            if (p1) wlists[(size_t)r1 * LCAP + c1 + __popc(b1 & below)] = make_key(x1, (uint32_t)col1);
            if (two && lane == r1) cnt = c1 + __popc(b1);
            // lists close to their capacity are cut back to their best entries (threshold refresh); rare
            uint32_t full = __ballot_sync(0xffffffffu, cnt >= trig);
            while (full) {
              const int r = __ffs(full) - 1;
              full &= full - 1;
              int cnt_r = __shfl_sync(0xffffffffu, cnt, r);
              float new_tau = __shfl_sync(0xffffffffu, tau, r);
              warp_compact(wlists + (size_t)r * LCAP, cnt_r, a.kp, new_tau, sscr2, lane, cnt_r, new_tau);
              if (lane == r) { cnt = cnt_r; tau = new_tau; }
            }
          }
          __syncwarp();
          nq_ = 0;
        };
#pragma unroll 1
        for (int c = 0; c < CN / 32; ++c) {
          float v[32];
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + e * CN + c * 32, v);
          tmem_wait_ld();
          if (c == CN / 32 - 1) {  // accumulator fully read: hand it back
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&d_empty[e]);
          }
          float m0 = fmaxf(fmaxf(v[0], v[1]), v[2]), m1 = fmaxf(fmaxf(v[3], v[4]), v[5]);
          float m2 = fmaxf(fmaxf(v[6], v[7]), v[8]), m3 = fmaxf(fmaxf(v[9], v[10]), v[11]);
#pragma unroll
          for (int i = 12; i < 32; i += 4) {
            m0 = fmaxf(m0, v[i]); m1 = fmaxf(m1, v[i + 1]); m2 = fmaxf(m2, v[i + 2]); m3 = fmaxf(m3, v[i + 3]);
          }
          const float m = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
          const bool hit = m > tau;
          const uint32_t hits = __ballot_sync(0xffffffffu, hit);
          if (hits == 0) continue;
          const int nh = __popc(hits);
          if (nq_ + nh > QCAP) drain();  // queue full (early in a scan): filter what is parked first
          if (hit) {
            const int s_ = nq_ + __popc(hits & ((1u << lane) - 1u));
            qmeta[s_] = (c << 5) | lane;
#pragma unroll
            for (int i = 0; i < 8; ++i)
              *reinterpret_cast<float4*>(fscr + s_ * 32 + 4 * i) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
          }
          nq_ += nh;
          __syncwarp();
        }
        if (nq_ > 0) drain();
      }
      // unit done: exact order of each row's list, best kp keys -> part[slot][e][row][kp]
      for (int r = 0; r < 32; ++r) {
        const int n = __shfl_sync(0xffffffffu, cnt, r);
        const unsigned long long* list = wlists + (size_t)r * LCAP;
        for (int i = lane; i < LCAP; i += 32) sscr[i] = i < n ? list[i] : 0ull;
        __syncwarp();
        warp_sort_desc(sscr, LCAP, lane);
        unsigned long long* dst = a.part + (((size_t)slot * 2 + e) * QT + q * 32 + r) * a.kp;
        for (int i = lane; i < a.kp; i += 32) dst[i] = sscr[i];
        __syncwarp();
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------------
// finalize: merge parts, fp32 re-score, exact sort, write top-k
// ------------------------------------------------------------------------------------------------
struct FinalArgs {
  int nq, nc, d, k, kp;
  int G, R, r, S;
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
  const int qtile = (int)(row / QT), qb = qtile >> 1, e = qtile & 1, rr = (int)(row % QT);
  int slot0, nparts;
  if (qb < a.R * a.G) { slot0 = qb; nparts = 1; }
  else { slot0 = a.R * a.G + (qb - a.R * a.G) * a.S; nparts = a.S; }
  // merged best kp by screening score
  for (int i = lane; i < LCAP; i += 32) s[i] = 0ull;
  __syncwarp();
  for (int p = 0; p < nparts; ++p) {
    const unsigned long long* src = a.part + (((size_t)(slot0 + p) * 2 + e) * QT + rr) * a.kp;
    // s[0, kp) holds the running best (sorted); append the next part behind it and re-sort
    for (int i = lane; i < a.kp; i += 32) s[KP_MAX + i] = src[i];
    for (int i = a.kp + lane; i < KP_MAX; i += 32) { s[i] = 0ull; s[KP_MAX + i] = 0ull; }
    __syncwarp();
    if (nparts > 1 || p == 0) warp_sort_desc(s, LCAP, lane);
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
  int DP, kp, NQB, CT, G, R, r, S, slots;
};

static int make_plan(long long nq, long long nc, long long d, long long k, Plan& p) {
  p.DP = d <= 64 ? 64 : 128;
  long long kp = k + (k / 4 > 32 ? k / 4 : 32);
  if (kp > nc) kp = nc;
  if (kp > KP_MAX) kp = KP_MAX;
  p.kp = (int)kp;
  p.NQB = (int)((nq + 2 * QT - 1) / (2 * QT));
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
  return (size_t)p.G * 2 * QT * LCAP * 8 + (size_t)p.slots * 2 * QT * p.kp * 8 + 1024;
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
  mips_screen_kernel<DP><<<a.G, 384, Cfg::SMEM_BYTES, st>>>(tq, tc, a);
  TT_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

int mips_topk(const void* Q16, long long ldq, const void* C16, long long ldc, const float* Q32, long long ldq32,
              const float* C32, long long ldc32, long long nq, long long nc, long long d, long long k, long long* idx,
              float* scores, void* ws, size_t ws_bytes, cudaStream_t stream) {
  TT_CHECK(nq > 0 && nc > 0 && d > 0, "mips_topk: empty problem");
  TT_CHECK(k > 0 && k <= nc, "mips_topk: selected index k out of range (k=%lld, corpus %lld)", k, nc);
  TT_CHECK(d <= 128, "mips_topk: embedding dim %lld > 128 is not supported by the fused kernel", d);
  TT_CHECK(k <= KP_MAX - 32, "mips_topk: k=%lld > %d is not supported by the fused kernel", k, KP_MAX - 32);
  TT_CHECK(nc < (1ll << 32) - 1, "mips_topk: corpus too large for 32-bit row indices");
  TT_CHECK((ldq % 8) == 0 && (ldc % 8) == 0 && ((uintptr_t)Q16 % 16) == 0 && ((uintptr_t)C16 % 16) == 0,
           "mips_topk: bf16 operands need 16-byte aligned rows");
  Plan p;
  make_plan(nq, nc, d, k, p);
  TT_CHECK(ws_bytes >= mips_workspace_bytes(nq, nc, d, k), "mips_topk: workspace too small");
  ScreenArgs a;
  a.nq = (int)nq; a.nc = (int)nc; a.kp = p.kp;
  a.G = p.G; a.R = p.R; a.r = p.r; a.S = p.S; a.CT = p.CT;
  a.lists = (unsigned long long*)ws;
  a.part = a.lists + (size_t)p.G * 2 * QT * LCAP;
  CUtensorMap tq, tc;
  int rc = make_tmap_bf16(&tq, Q16, d, nq, ldq, 64, QT);
  if (rc) return rc;
  rc = make_tmap_bf16(&tc, C16, d, nc, ldc, 64, CN);
  if (rc) return rc;
  rc = p.DP == 64 ? launch_screen<64>(tq, tc, a, stream) : launch_screen<128>(tq, tc, a, stream);
  if (rc) return rc;
  FinalArgs f;
  f.nq = (int)nq; f.nc = (int)nc; f.d = (int)d; f.k = (int)k; f.kp = p.kp;
  f.G = p.G; f.R = p.R; f.r = p.r; f.S = p.S;
  f.part = a.part;
  f.q32 = Q32; f.ldq = ldq32; f.c32 = C32; f.ldc = ldc32;
  f.idx_out = idx; f.score_out = scores;
  KernelSpan span("mips_finalize_kernel", stream);
  mips_finalize_kernel<<<(unsigned)((nq + 3) / 4), 128, 0, stream>>>(f);
  TT_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

}  // namespace tt
