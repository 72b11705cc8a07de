// C ABI (include/tt_b200.h) over the kernel translation units + shared host utilities.
#include <stdarg.h>
#include <atomic>
#include <mutex>
#include <string>
#include <vector>
#include <map>
#include <string.h>

#include "../../include/tt_b200.h"
#include "common.cuh"
#include "kernels.h"

namespace tt {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static std::atomic<long long> g_launches{0};
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// ---- optional per-kernel timing --------------------------------------------------------------------------
struct SpanRec {
  const char* name;
  cudaEvent_t e0, e1;
};
static std::atomic<int> g_profile{0};  // 1: eager launches; 2: launches under stream capture too (external event nodes)
static std::mutex g_span_mu;
static std::vector<SpanRec> g_spans;        // eager launches since the last report
static std::vector<SpanRec> g_graph_spans;  // launches captured into CUDA graphs: re-stamped by every replay

KernelSpan::KernelSpan(const char* name, cudaStream_t stream) : rec_(nullptr), stream_(stream), captured_(false) {
  const int mode = g_profile.load(std::memory_order_relaxed);
  if (!mode) return;
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(stream, &cs) != cudaSuccess) return;
  captured_ = cs != cudaStreamCaptureStatusNone;
  if (captured_ && mode < 2) return;
  SpanRec* r = new SpanRec{name, nullptr, nullptr};
  cudaEventCreate(&r->e0);
  cudaEventCreate(&r->e1);
  if (captured_) cudaEventRecordWithFlags(r->e0, stream, cudaEventRecordExternal);
  else cudaEventRecord(r->e0, stream);
  rec_ = r;
}
KernelSpan::~KernelSpan() {
  if (!rec_) return;
  SpanRec* r = static_cast<SpanRec*>(rec_);
  if (captured_) cudaEventRecordWithFlags(r->e1, stream_, cudaEventRecordExternal);
  else cudaEventRecord(r->e1, stream_);
  std::lock_guard<std::mutex> lk(g_span_mu);
  (captured_ ? g_graph_spans : g_spans).push_back(*r);
  delete r;
}

static int g_sm_limit = 0;  // tt_set_sm_limit: SMs the persistent kernels may occupy (0 = all)
int num_sms() {
  static int cached = 0;
  if (cached > 0) return (g_sm_limit > 0 && g_sm_limit < cached) ? g_sm_limit : cached;
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 1;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) return 1;
  cached = n;
  return (g_sm_limit > 0 && g_sm_limit < n) ? g_sm_limit : n;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
      q != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<EncodeTiledFn>(p);
  return fn;
}

int make_tmap_bf16(CUtensorMap* map, const void* base, uint64_t inner, uint64_t outer, uint64_t pitch_elems,
                   uint32_t box_inner, uint32_t box_outer) {
  EncodeTiledFn fn = get_encode_fn();
  TT_CHECK(fn != nullptr, "cuTensorMapEncodeTiled is unavailable (no CUDA driver / device)");
  TT_CHECK(((uintptr_t)base % 16) == 0 && (pitch_elems % 8) == 0, "tensor map: base/pitch not 16-byte aligned");
  TT_CHECK(box_inner * 2 <= 128 && box_outer <= 256, "tensor map: box too large");
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {pitch_elems * 2};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  TT_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult %d (inner=%llu outer=%llu pitch=%llu box=%ux%u)",
           (int)r, (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)pitch_elems, box_inner, box_outer);
  return 0;
}

}  // namespace tt

using namespace tt;

static inline cudaStream_t S(void* s) { return reinterpret_cast<cudaStream_t>(s); }

extern "C" {

int tt_abi_version(void) { return TT_B200_ABI_VERSION; }
const char* tt_last_error(void) { return g_err; }

long long tt_launch_count(void) { return (long long)g_launches.load(); }

void tt_profile_enable(int on) { g_profile.store(on < 0 ? 0 : (on > 2 ? 2 : on)); }

int tt_profile_report(char* buf, int64_t buf_bytes) {
  TT_CUDA(cudaDeviceSynchronize());
  std::map<std::string, std::pair<double, long long>> agg;
  {
    std::lock_guard<std::mutex> lk(g_span_mu);
    for (auto& r : g_spans) {
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, r.e0, r.e1) == cudaSuccess) {
        auto& a = agg[r.name];
        a.first += ms;
        a.second += 1;
      }
      cudaEventDestroy(r.e0);
      cudaEventDestroy(r.e1);
    }
    g_spans.clear();
  }
  std::string out;
  for (auto& kv : agg) {
    char line[256];
    snprintf(line, sizeof(line), "%s %.6f %lld\n", kv.first.c_str(), kv.second.first, kv.second.second);
    out += line;
  }
  TT_CHECK((int64_t)out.size() + 1 <= buf_bytes, "tt_profile_report: buffer too small");
  memcpy(buf, out.c_str(), out.size() + 1);
  return 0;
}

int tt_profile_report_graph(char* buf, int64_t buf_bytes, int32_t clear) {
  TT_CUDA(cudaDeviceSynchronize());
  std::map<std::string, std::pair<double, long long>> agg;
  {
    std::lock_guard<std::mutex> lk(g_span_mu);
    for (auto& r : g_graph_spans) {
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, r.e0, r.e1) == cudaSuccess) {
        auto& a = agg[r.name];
        a.first += ms;
        a.second += 1;
      }
      if (clear) {
        cudaEventDestroy(r.e0);
        cudaEventDestroy(r.e1);
      }
    }
    if (clear) g_graph_spans.clear();
  }
  std::string out;
  for (auto& kv : agg) {
    char line[256];
    snprintf(line, sizeof(line), "%s %.6f %lld\n", kv.first.c_str(), kv.second.first, kv.second.second);
    out += line;
  }
  TT_CHECK((int64_t)out.size() + 1 <= buf_bytes, "tt_profile_report_graph: buffer too small");
  memcpy(buf, out.c_str(), out.size() + 1);
  return 0;
}

// An empty kernel inside a span: what the event pair itself adds to every kernel span measured the same way.
__global__ void null_kernel() {}
int tt_profile_null_span(void* stream) {
  KernelSpan span("null_kernel", (cudaStream_t)stream);
  null_kernel<<<1, 32, 0, (cudaStream_t)stream>>>();
  TT_CUDA(cudaGetLastError());
  return 0;
}

int tt_device_sm_count(void) {
  int dev = 0, n = 0;
  TT_CUDA(cudaGetDevice(&dev));
  TT_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
  return n;
}

int tt_cast_rows_bf16(const float* src, int64_t rows, int64_t cols, int64_t ld_src, void* dst, int64_t ld_dst,
                      int64_t dst_cols, void* stream) {
  return cast_rows_bf16(src, rows, cols, ld_src, dst, ld_dst, dst_cols, S(stream));
}
int tt_gather_rows_bf16(const float* table, int64_t table_rows, int64_t dim, const int64_t* ids, int64_t n, void* dst,
                        int64_t ld_dst, int32_t* oob_flag, void* stream) {
  return gather_rows_bf16(table, table_rows, dim, (const long long*)ids, n, dst, ld_dst, oob_flag, S(stream));
}
int tt_gather_rows_f32(const float* table, int64_t table_rows, int64_t dim, const int64_t* ids, int64_t n, float* dst,
                       int64_t ld_dst, int32_t* oob_flag, void* stream) {
  return gather_rows_f32(table, table_rows, dim, (const long long*)ids, n, dst, ld_dst, oob_flag, S(stream));
}
int tt_scatter_add_rows(const void* src16, const float* src32, int64_t ld_src, const int64_t* ids, int64_t n,
                        int64_t dim, float* table_grad, int64_t table_rows, void* stream) {
  return scatter_add_rows(src16, src32, ld_src, (const long long*)ids, n, dim, table_grad, table_rows, S(stream));
}
int tt_colsum(const void* src16, const float* src32, int64_t rows, int64_t cols, int64_t ld, float* out, void* stream) {
  return colsum(src16, src32, rows, cols, ld, out, S(stream));
}

int tt_gemm_bf16(const void* A, int64_t lda, int32_t a_mn_major, const void* B, int64_t ldb, int32_t b_mn_major,
                 int64_t M, int64_t N, int64_t K, const float* bias, int32_t relu, const void* relu_mask, int64_t ld_mask,
                 float alpha, float* c_f32, int64_t ldc_f32, void* c_bf16, int64_t ldc_bf16, int32_t accumulate,
                 int32_t split_k, float* colsum_f32, void* stream) {
  GemmDesc d;
  d.A = A; d.lda = lda; d.a_mn_major = a_mn_major;
  d.B = B; d.ldb = ldb; d.b_mn_major = b_mn_major;
  d.M = M; d.N = N; d.K = K;
  d.bias = bias; d.relu = relu; d.relu_mask = relu_mask; d.ld_mask = ld_mask;
  d.alpha = alpha;
  d.c32 = c_f32; d.ldc32 = ldc_f32; d.c16 = c_bf16; d.ldc16 = ldc_bf16;
  d.accumulate = accumulate; d.split_k = split_k; d.colsum = colsum_f32;
  return gemm_bf16(d, S(stream));
}

int tt_gemm_bf16_batched(const tt_gemm_problem* pr, int32_t count, void* stream) {
  TT_CHECK(pr != nullptr && count >= 1 && count <= 16, "tt_gemm_bf16_batched: 1..16 problems");
  GemmDesc d[16];
  for (int i = 0; i < count; ++i) {
    d[i].A = pr[i].A; d[i].lda = pr[i].lda; d[i].a_mn_major = pr[i].a_mn_major;
    d[i].B = pr[i].B; d[i].ldb = pr[i].ldb; d[i].b_mn_major = pr[i].b_mn_major;
    d[i].M = pr[i].M; d[i].N = pr[i].N; d[i].K = pr[i].K;
    d[i].bias = pr[i].bias; d[i].relu = pr[i].relu; d[i].relu_mask = pr[i].relu_mask_bf16; d[i].ld_mask = pr[i].ld_mask;
    d[i].alpha = pr[i].alpha;
    d[i].c32 = pr[i].c_f32; d[i].ldc32 = pr[i].ldc_f32; d[i].c16 = pr[i].c_bf16; d[i].ldc16 = pr[i].ldc_bf16;
    d[i].accumulate = pr[i].accumulate; d[i].split_k = pr[i].split_k; d[i].colsum = pr[i].colsum_f32;
  }
  return gemm_bf16_batched(d, count, S(stream));
}
int tt_cast_rows_bf16_batched(const tt_cast_problem* pr, int32_t count, void* stream) {
  TT_CHECK(pr != nullptr && count >= 1 && count <= 16, "tt_cast_rows_bf16_batched: 1..16 problems");
  CastProblem c[16];
  for (int i = 0; i < count; ++i) c[i] = CastProblem{pr[i].src, pr[i].rows, pr[i].cols, pr[i].ld_src, pr[i].dst_bf16, pr[i].ld_dst, pr[i].dst_cols};
  return cast_rows_bf16_batched(c, count, S(stream));
}
int tt_gather_rows_bf16_batched(const tt_gather_problem* pr, int32_t count, int32_t* oob_flag, void* stream) {
  TT_CHECK(pr != nullptr && count >= 1 && count <= 8, "tt_gather_rows_bf16_batched: 1..8 problems");
  GatherProblem g[8];
  for (int i = 0; i < count; ++i)
    g[i] = GatherProblem{pr[i].table, pr[i].table_rows, pr[i].dim, (const long long*)pr[i].ids, pr[i].n, pr[i].dst_bf16, pr[i].ld_dst};
  return gather_rows_bf16_batched(g, count, oob_flag, S(stream));
}

int32_t tt_tower_fwd_supported(int64_t F, int64_t D, int64_t DI, int64_t hidden) {
  return tower_fwd_supported(F, D, DI, hidden) ? 1 : 0;
}
int tt_tower_fwd(const tt_tower_problem* pr, int32_t count, int32_t* oob_flag, void* stream) {
  TT_CHECK(pr != nullptr && count >= 1 && count <= 4, "tt_tower_fwd: 1..4 towers");
  TowerProblem t[4];
  for (int i = 0; i < count; ++i) {
    const tt_tower_problem& q = pr[i];
    t[i] = TowerProblem{(const long long*)q.ids, q.table, q.table_rows, q.feats, q.ld_feats, q.w0_bf16, q.ldw0, q.b0,
                        q.w1_bf16, q.ldw1, q.b1, q.wt_bf16, q.ldwt, q.bt, q.feats_bf16, q.ld_feats16, q.h_bf16, q.ldh,
                        q.x_bf16, q.ldx, q.emb_f32, q.ld_emb, q.emb_bf16, q.ld_emb16, q.rows, q.F, q.D, q.DI, q.hidden};
  }
  return tower_fwd(t, count, oob_flag, S(stream));
}

int tt_tower_bwd_chain(const tt_tower_bwd_problem* pr, int32_t count, void* stream) {
  TT_CHECK(pr != nullptr && count >= 1 && count <= 4, "tt_tower_bwd_chain: 1..4 towers");
  TowerBwdProblem t[4];
  for (int i = 0; i < count; ++i) {
    const tt_tower_bwd_problem& q = pr[i];
    t[i] = TowerBwdProblem{q.demb_bf16, q.ld_demb, (const long long*)q.ids, q.table_rows, q.wt_bf16, q.ldwt, q.w1_bf16, q.ldw1,
                           q.h_bf16, q.ldh, q.dx_bf16, q.lddx, q.dh_bf16, q.lddh, q.table_grad, q.dxsum, q.db0, q.rows, q.D, q.DI,
                           q.hidden};
  }
  return tower_bwd_chain(t, count, S(stream));
}

int64_t tt_inbatch_ce_workspace_bytes(int64_t B, int64_t N, int64_t d) {
  return (int64_t)inbatch_ce_workspace_bytes(B, N, d);
}
int tt_inbatch_ce_fwd(const void* U, int64_t ldu, const void* V, int64_t ldv, int64_t B, int64_t N, int64_t d,
                      int64_t target_offset, float* ce, float* lse, void* ws, int64_t ws_bytes, void* stream) {
  return inbatch_ce_fwd(U, ldu, V, ldv, B, N, d, target_offset, ce, lse, ws, (size_t)ws_bytes, S(stream));
}
int tt_inbatch_ce_bwd(const void* U, int64_t ldu, const void* V, int64_t ldv, int64_t B, int64_t N, int64_t d,
                      int64_t target_offset, const float* lse, const float* g, float* dU, int64_t lddu, void* dU16,
                      int64_t lddu16, float* dV, int64_t lddv, void* dV16, int64_t lddv16, float* dU_colsum,
                      float* dV_colsum, void* ws, int64_t ws_bytes, void* stream) {
  return inbatch_ce_bwd(U, ldu, V, ldv, B, N, d, target_offset, lse, g, dU, lddu, dU16, lddu16, dV, lddv, dV16, lddv16,
                        dU_colsum, dV_colsum, ws, (size_t)ws_bytes, S(stream));
}

int tt_inbatch_ce_fwd_parts(const void* U, int64_t ldu, const void* const* Vp, int32_t np, int64_t rows_per_part, int64_t ldv,
                            int64_t B, int64_t N, int64_t d, int64_t target_offset, float* ce, float* lse, void* ws,
                            int64_t ws_bytes, void* stream) {
  TT_CHECK(Vp != nullptr, "tt_inbatch_ce_fwd_parts: null part list");
  return inbatch_ce_fwd_parts(U, ldu, Vp, np, rows_per_part, ldv, B, N, d, target_offset, ce, lse, ws, (size_t)ws_bytes,
                              S(stream));
}
int tt_inbatch_ce_bwd_parts(const void* U, int64_t ldu, const void* const* Vp, int32_t np, int64_t rows_per_part, int64_t ldv,
                            int64_t B, int64_t N, int64_t d, int64_t target_offset, const float* lse, const float* g,
                            float* dU, int64_t lddu, void* dU16, int64_t lddu16, float* dV, int64_t lddv, void* dV16,
                            int64_t lddv16, float* dU_colsum, float* dV_colsum, void* ws, int64_t ws_bytes, void* stream) {
  TT_CHECK(Vp != nullptr, "tt_inbatch_ce_bwd_parts: null part list");
  return inbatch_ce_bwd_parts(U, ldu, Vp, np, rows_per_part, ldv, B, N, d, target_offset, lse, g, dU, lddu, dU16, lddu16, dV,
                              lddv, dV16, lddv16, dU_colsum, dV_colsum, ws, (size_t)ws_bytes, S(stream));
}

int tt_inbatch_ce_loss_fwd(const void* U, int64_t ldu, const void* V, int64_t ldv, int64_t B, int64_t N, int64_t d,
                           int64_t target_offset, const float* labels, int64_t ldl, const float* weights, int64_t T,
                           float* ce, float* lse, float* loss, float* g, float* g_norm, void* ws, int64_t ws_bytes,
                           void* stream) {
  TT_CHECK(labels != nullptr && weights != nullptr && loss != nullptr && g != nullptr && g_norm != nullptr, "tt_inbatch_ce_loss_fwd: null argument");
  return inbatch_ce_loss_fwd(U, ldu, &V, 1, N, ldv, B, N, d, target_offset, ce, lse, labels, ldl, weights, T, loss, g, g_norm,
                             ws, (size_t)ws_bytes, S(stream));
}
int tt_inbatch_ce_attach_zero_fill(void* p0, int64_t bytes0, void* p1, int64_t bytes1) {
  return inbatch_ce_attach_zero_fill(p0, bytes0, p1, bytes1);
}
int tt_set_sm_limit(int32_t n) {
  const int prev = g_sm_limit;
  g_sm_limit = n > 0 ? n : 0;
  return prev;
}
int tt_history_last_supported(int64_t H, int64_t D, int64_t heads) { return history_last_supported(H, D, heads); }
int tt_history_last_fwd(const void* x16, int64_t ldx, const float* qt, int64_t B, int64_t H, int64_t D, int64_t heads, void* z16,
                        float* p32, void* stream) {
  TT_CHECK(x16 && qt && z16 && p32, "tt_history_last_fwd: null argument");
  return history_last_fwd(x16, ldx, qt, B, H, D, heads, z16, p32, S(stream));
}
int tt_history_last_bwd1(const void* x16, int64_t ldx, const float* dz, const float* p32, int64_t B, int64_t H, int64_t D,
                         int64_t heads, float* ds32, void* dqt16, void* stream) {
  TT_CHECK(x16 && dz && p32 && ds32 && dqt16, "tt_history_last_bwd1: null argument");
  return history_last_bwd1(x16, ldx, dz, p32, B, H, D, heads, ds32, dqt16, S(stream));
}
int tt_history_last_bwd2(const float* dz, const float* qt, const float* p32, const float* ds32, const float* extra, int64_t B,
                         int64_t H, int64_t D, int64_t heads, void* dx16, int64_t lddx, float* colsum, void* stream) {
  TT_CHECK(dz && qt && p32 && ds32 && dx16, "tt_history_last_bwd2: null argument");
  return history_last_bwd2(dz, qt, p32, ds32, extra, B, H, D, heads, dx16, lddx, colsum, S(stream));
}
int tt_inbatch_ce_loss_fwd_sharded(const void* U, int64_t ldu, const void* V, int64_t ldv, int64_t B, int64_t N, int64_t d,
                                   int64_t target_offset, const float* labels, int64_t ldl, const float* weights, int64_t T,
                                   float* ce, float* lse, float* g, float* stats, void* ws, int64_t ws_bytes, void* stream) {
  TT_CHECK(labels != nullptr && weights != nullptr && g != nullptr && stats != nullptr, "tt_inbatch_ce_loss_fwd_sharded: null argument");
  // the local loss / g_norm scalars are by-products nobody reads here: they land behind the two statistics
  return inbatch_ce_loss_fwd(U, ldu, &V, 1, N, ldv, B, N, d, target_offset, ce, lse, labels, ldl, weights, T, stats + 2, g,
                             stats + 3, ws, (size_t)ws_bytes, S(stream), stats);
}
int tt_sharded_loss_finalize(const float* stats_all, int32_t world, int64_t global_rows, float* loss, float* g_norm,
                             void* stream) {
  return sharded_loss_finalize(stats_all, world, global_rows, loss, g_norm, S(stream));
}
int tt_inbatch_ce_bwd_scaled(const void* U, int64_t ldu, const void* V, int64_t ldv, int64_t B, int64_t N, int64_t d,
                             int64_t target_offset, const float* lse, const float* g, const float* g_scale,
                             const float* g_scale2, float* dU,
                             int64_t lddu, void* dU16, int64_t lddu16, float* dV, int64_t lddv, void* dV16, int64_t lddv16,
                             float* dU_colsum, float* dV_colsum, void* ws, int64_t ws_bytes, void* stream) {
  return inbatch_ce_bwd(U, ldu, V, ldv, B, N, d, target_offset, lse, g, dU, lddu, dU16, lddu16, dV, lddv, dV16, lddv16,
                        dU_colsum, dV_colsum, ws, (size_t)ws_bytes, S(stream), g_scale, g_scale2);
}

int tt_adam_step(const tt_adam_tensor* ts, int32_t count, double lr, double beta1, double beta2, float eps, float weight_decay,
                 int64_t* step_dev, uint32_t* ticket_dev, void* stream) {
  TT_CHECK(ts != nullptr && count >= 1 && count <= 32, "tt_adam_step: 1..32 tensors");
  AdamTensor a[32];
  for (int i = 0; i < count; ++i) a[i] = AdamTensor{ts[i].param, ts[i].grad, ts[i].exp_avg, ts[i].exp_avg_sq, ts[i].numel, 0};
  return adam_step(a, count, lr, beta1, beta2, eps, weight_decay, (long long*)step_dev, ticket_dev, S(stream));
}

int tt_weighted_loss(const float* ce, const float* labels, int64_t ldl, const float* weights, int64_t B, int64_t T,
                     float* loss, float* g, void* stream) {
  return weighted_loss(ce, labels, ldl, weights, B, T, loss, g, S(stream));
}

int64_t tt_mips_workspace_bytes(int64_t nq, int64_t nc, int64_t d, int64_t k) {
  return (int64_t)mips_workspace_bytes(nq, nc, d, k);
}
int tt_mips_topk(const void* Q16, int64_t ldq16, const void* C16, int64_t ldc16, const float* Q32, int64_t ldq32,
                 const float* C32, int64_t ldc32, int64_t nq, int64_t nc, int64_t d, int64_t k, int64_t* idx,
                 float* scores, void* ws, int64_t ws_bytes, void* stream) {
  return mips_topk(Q16, ldq16, C16, ldc16, Q32, ldq32, C32, ldc32, nq, nc, d, k, (long long*)idx, scores, ws,
                   (size_t)ws_bytes, S(stream));
}

int tt_history_gather_pool(const float* table, int64_t table_rows, int64_t D, const int64_t* ids, int64_t B, int64_t H,
                           const float* pe, void* x16, int64_t ldx, float* mean, int64_t ldmean, int32_t* oob_flag,
                           void* stream) {
  return history_gather_pool(table, table_rows, D, (const long long*)ids, B, H, pe, x16, ldx, mean, ldmean, oob_flag, S(stream));
}
int tt_history_scatter_grad(const void* dx16, int64_t lddx, const float* dmean, int64_t lddmean, const int64_t* ids,
                            int64_t B, int64_t H, int64_t D, float* table_grad, int64_t table_rows, void* stream) {
  return history_scatter_grad(dx16, lddx, dmean, lddmean, (const long long*)ids, B, H, D, table_grad, table_rows, S(stream));
}

int tt_attn_fwd(const void* qkv, int64_t ldqkv, int64_t nseq, int64_t H, int64_t D, int64_t heads, int64_t q_rows,
                void* out, int64_t ldo, void* stream) {
  return attn_fwd(qkv, ldqkv, nseq, H, D, heads, q_rows, out, ldo, S(stream));
}
int tt_attn_bwd(const void* qkv, int64_t ldqkv, const void* dout, int64_t lddo, int64_t nseq, int64_t H, int64_t D,
                int64_t heads, int64_t q_rows, void* dqkv, int64_t lddqkv, void* stream) {
  return attn_bwd(qkv, ldqkv, dout, lddo, nseq, H, D, heads, q_rows, dqkv, lddqkv, S(stream));
}

}  // extern "C"
