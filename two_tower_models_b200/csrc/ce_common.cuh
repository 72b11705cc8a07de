// Shared pieces of the in-batch cross-entropy kernels (ce.cu, ce_bwd2.cu).
#pragma once
#include "common.cuh"

namespace tt {

static constexpr float LOG2E = 1.4426950408889634f;
static constexpr float LN2 = 0.6931471805599453f;

// Walks the contiguous range of (row tile, column tile) work items of one CTA, segment by segment (a segment = the part
// of the range inside one row tile).  The flattened space may carry `CT - CTr` GHOST column tiles at the end of every row
// tile: they are never executed, so a CTA whose range crosses a row-tile boundary (and pays for an accumulator drain and
// a pipeline refill there) gets that many fewer real tiles than one that does not.
#ifdef __CUDACC__
// exp2 on the FMA/ALU pipes (degree-3 minimax of 2^f on [-0.5, 0.5], relative error 7.5e-5): offloads a share of the
// exponentials from the MUFU pipe (16/clk/SM), see tools/micro/mufu_bench.cu.  x <= 0 in the CE kernels; x < -126 -> ~0.
__device__ __forceinline__ float exp2_poly3(float x) {
  x = fmaxf(x, -126.f);
  const float t = x + 12582912.f;  // 1.5 * 2^23: the integer part lands in the low mantissa bits
  const float f = x - (t - 12582912.f);
  float p = fmaf(f, 0.05517166f, 0.24261112f);
  p = fmaf(p, f, 0.69326099f);
  p = fmaf(p, f, 0.99992807f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}
#endif

struct SegIter {
  long long f, f1;
  int CT, CTr;
  __device__ SegIter(long long T, long long total, int ct, int ctr = -1) {
    f = (long long)blockIdx.x * T;
    f1 = f + T < total ? f + T : total;
    CT = ct;
    CTr = ctr < 0 ? ct : ctr;
  }
  __device__ bool next(int& r, int& j0, int& j1) {
    while (f < f1) {
      r = (int)(f / CT);
      j0 = (int)(f % CT);
      long long rem = f1 - f;
      j1 = (long long)j0 + rem < CT ? (int)(j0 + rem) : CT;
      f += j1 - j0;
      if (j0 >= CTr) continue;  // only ghost tiles of this row tile
      if (j1 > CTr) j1 = CTr;
      return true;
    }
    return false;
  }
};

struct Sched {
  long long T, total;
  int XT, CT, CTr, max_slots, grid;  // CT = column tiles per row tile INCLUDING ghosts, CTr = real ones
};
static inline Sched make_sched(long long x_rows, long long y_rows, int BN, int ghost = 0) {
  Sched s;
  s.XT = (int)((x_rows + 127) / 128);
  s.CTr = (int)((y_rows + BN - 1) / BN);
  if ((long long)s.XT * s.CTr < 16LL * num_sms()) ghost = 0;  // short ranges: nothing to balance
  s.CT = s.CTr + ghost;
  s.total = (long long)s.XT * s.CT;
  s.grid = (int)(s.total < num_sms() ? s.total : num_sms());
  s.T = (s.total + s.grid - 1) / s.grid;
  s.grid = (int)((s.total + s.T - 1) / s.T);
  s.max_slots = (int)((s.CT + s.T - 1) / s.T) + 1;
  return s;
}

// An operand whose rows are spread over up to 8 equally sized parts (one tensor map each): the all-gathered item
// embeddings of the data-parallel loss, read IN PLACE from the peers' memory over NVLink.  n == 1: a plain matrix.
struct TmapSet {
  CUtensorMap m[8];
  int n;
  int rows_per_map;
};
// tensor map + row coordinate of global row `row` (row and rows_per_map are multiples of the box height)
__device__ __forceinline__ const CUtensorMap* tmap_of(const TmapSet& t, int row, int& local_row) {
  const int p = t.n == 1 ? 0 : row / t.rows_per_map;
  local_row = row - p * t.rows_per_map;
  return &t.m[p];
}

struct CeBwdArgs {
  int XR, YR;            // valid rows of X / Y
  long long diag_shift;  // element (row, col) is a positive when col == row + diag_shift
  long long T, total;
  int CT;
  const float* g;    // upstream dL/dce, indexed by user
  const float* g_scale;   // optional device scalars multiplied into g (incoming d loss; weight normalisation
  const float* g_scale2;  // 1 / (max nuv * B) of the fused weighted mean), or null
  const float* lse;  // indexed by user
  float* partial;    // [max_slots][XT*128][DP]
  long long slot_stride;
  long long* trace;  // bring-up (TT_CE_TRACE): clock64 stamps of CTA 0, normally null
  unsigned long long* cta_times;  // bring-up (TT_CE_CTA_TIMES): [grid][4] %globaltimer at start / setup / work done / exit
  int dbg;           // bring-up (TT_CE_DBG): bit0 skip ex2, bit1 skip E store, bit2 skip transform entirely
};


// v2 backward kernel (ce_bwd2.cu): E operand and X tile in tensor memory
int launch_ce_bwd2(int DP, bool colstats, const TmapSet& tx, const TmapSet& ty, const CeBwdArgs& a, int grid,
                   cudaStream_t st);

// v3 backward kernel (ce_bwd3.cu, d <= 128): statistics folded into the score MMA, E in place, 16 epilogue warps
struct CeBwd3Args {
  int XR, YR;
  long long diag_shift;
  long long T, total;
  int CT, CTr;               // column tiles per row tile with / without the ghost tiles (see SegIter)
  const float* g;            // upstream dL/dce, indexed by user (any sign)
  const float* g_scale;      // optional device scalars multiplied into g, or null
  const float* g_scale2;
  const float* lse;          // indexed by user
  const uint32_t* signmask;  // sign bits of g, 32 users per word (set by launch_ce_bwd3 from the ext block)
  float* partial;            // [max_slots][XT*128][DP]
  long long slot_stride;
  long long* trace;               // bring-up builds only (TT_CE_BRINGUP)
  unsigned long long* cta_times;
  int trace_cta;                  // CTA whose clock64 timeline is recorded (TT_CE_TRACE_CTA)
};
// bytes of the per-launch "ext" block: [users_pad, 64] bf16 bias rows + sign words; filled by ce_bwd3_prep
int ce_bwd3_tile_cols(int DP);  // columns of Y per score tile (128 at d <= 64, 96 at d <= 128)
size_t ce_bwd3_ext_bytes(long long users);
int ce_bwd3_prep(long long users, const float* g, const float* lse, void* ext, cudaStream_t st);
// 128 < d <= 256 (ce_bwd3x.cu): X operand in shared memory, 128 x 64 score tiles
int launch_ce_bwd3x(int DP, bool bias_x, const TmapSet& tx, const TmapSet& ty, long long users, const void* ext, CeBwd3Args a,
                    int grid, cudaStream_t st);
int launch_ce_bwd3(int DP, bool bias_x, const TmapSet& tx, const TmapSet& ty, long long users, const void* ext,
                   CeBwd3Args a, int grid, cudaStream_t st);

}  // namespace tt
