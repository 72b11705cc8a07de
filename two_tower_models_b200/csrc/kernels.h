// Internal C++ interfaces between the C-ABI layer (api.cu) and the kernel translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace tt {

struct GemmDesc {
  const void* A = nullptr;  // bf16; [M,K] pitch lda, or [K,M] pitch lda when a_mn_major
  long long lda = 0;
  int a_mn_major = 0;
  const void* B = nullptr;  // bf16; [N,K] pitch ldb, or [K,N] pitch ldb when b_mn_major
  long long ldb = 0;
  int b_mn_major = 0;
  long long M = 0, N = 0, K = 0;
  const float* bias = nullptr;  // [N]
  int relu = 0;
  const void* relu_mask = nullptr;  // bf16 [M, ld_mask]; output zeroed where mask <= 0
  long long ld_mask = 0;
  float* c32 = nullptr;
  long long ldc32 = 0;
  void* c16 = nullptr;  // bf16
  long long ldc16 = 0;
  int accumulate = 0;  // fp32 atomicAdd into c32 (split-K); c32 must be initialised by the caller
  int split_k = 0;     // 0 = auto
  float alpha = 1.f;
  float* colsum = nullptr;  // optional [N] fp32: += column sums of the final C (caller initialises)
};
int gemm_bf16(const GemmDesc& d, cudaStream_t stream);
// n independent problems; consecutive problems with the same tile width share one launch (up to 4)
int gemm_bf16_batched(const GemmDesc* d, int n, cudaStream_t stream);

// In-batch sampled-softmax cross entropy.
size_t inbatch_ce_workspace_bytes(long long B, long long N, long long d);
int inbatch_ce_fwd(const void* U, long long ldu, const void* V, long long ldv, long long B, long long N, long long d,
                   long long target_offset, float* ce, float* lse, void* ws, size_t ws_bytes, cudaStream_t stream);
// The same with the item matrix V given as `np` equally sized row blocks (e.g. the peers' buffers of a data-parallel
// step, read in place over NVLink): V = [Vp[0]; Vp[1]; ...], rows_per_part rows each (a multiple of 128).
int inbatch_ce_fwd_parts(const void* U, long long ldu, const void* const* Vp, int np, long long rows_per_part,
                         long long ldv, long long B, long long N, long long d, long long target_offset, float* ce,
                         float* lse, void* ws, size_t ws_bytes, cudaStream_t stream);
int inbatch_ce_bwd_parts(const void* U, long long ldu, const void* const* Vp, int np, long long rows_per_part,
                         long long ldv, long long B, long long N, long long d, long long target_offset, const float* lse,
                         const float* g, float* dU, long long lddu, void* dU16, long long lddu16, float* dV, long long lddv,
                         void* dV16, long long lddv16, float* dU_colsum, float* dV_colsum, void* ws, size_t ws_bytes,
                         cudaStream_t stream, const float* g_scale = nullptr, const float* g_scale2 = nullptr);
// Forward fused with the value-weighted mean of the identity debias hook (labels != null): the kernel that merges the
// partials also writes g = nuv (unnormalised weights), *g_norm = 1 / (max nuv * B) and *loss; d loss / d ce = g * g_norm.
int inbatch_ce_loss_fwd(const void* U, long long ldu, const void* const* Vp, int np, long long rows_per_part,
                        long long ldv, long long B, long long N, long long d, long long target_offset, float* ce,
                        float* lse, const float* labels, long long ldl, const float* uvw, long long TL, float* loss,
                        float* g, float* g_norm, void* ws, size_t ws_bytes, cudaStream_t stream, float* stats = nullptr);
// the next forward launch of the calling thread also zero-fills these buffers (TMA bulk stores from an idle warp)
int inbatch_ce_attach_zero_fill(void* p0, long long bytes0, void* p1, long long bytes1);
// stats (optional, [2]): this rank's (max nuv, sum ce nuv) for the batch-sharded loss; merged by sharded_loss_finalize
int sharded_loss_finalize(const float* stats_all, int world, long long rows, float* loss, float* g_norm, cudaStream_t stream);
// dU (fp32 [B,d], optional bf16 copy) and dV (fp32 [N,d], optional bf16 copy) from upstream g[B].
int inbatch_ce_bwd(const void* U, long long ldu, const void* V, long long ldv, long long B, long long N, long long d,
                   long long target_offset, const float* lse, const float* g, float* dU, long long lddu, void* dU16,
                   long long lddu16, float* dV, long long lddv, void* dV16, long long lddv16, float* dU_colsum,
                   float* dV_colsum, void* ws, size_t ws_bytes, cudaStream_t stream, const float* g_scale = nullptr, const float* g_scale2 = nullptr);

// MIPS: top-k of Q C^T per query row.
size_t mips_workspace_bytes(long long Q, long long C, long long d, long long k);
int mips_topk(const void* Q16, long long ldq, const void* C16, long long ldc, const float* Q32, long long ldq32,
              const float* C32, long long ldc32, long long nq, long long nc, long long d, long long k, long long* idx,
              float* scores, void* ws, size_t ws_bytes, cudaStream_t stream);

// Self-attention core for the history encoder (per sequence, per head); qkv rows = [q | k | v] of width 3D.
// Only the first q_rows query rows of every sequence are computed (out / dout are [nseq*q_rows, D]).
int attn_fwd(const void* qkv, long long ld, long long nseq, long long H, long long D, long long heads, long long q_rows,
             void* out, long long ldo, cudaStream_t stream);
// tensor-core forward (attn_tc.cu); attn_fwd dispatches to it when the shape is supported
bool attn_fwd_tc_supported(long long H, long long D, long long heads, long long ld, long long ldo, const void* qkv,
                           const void* out);
int attn_fwd_tc(const void* qkv, long long ld, long long nseq, long long H, long long D, long long heads, long long q_rows,
                void* out, long long ldo, cudaStream_t stream);
bool attn_bwd_tc_supported(long long H, long long D, long long heads, long long ld, long long lddo, long long lddqkv,
                           const void* qkv, const void* dout, const void* dqkv);
int attn_bwd_tc(const void* qkv, long long ld, const void* dout, long long lddo, long long nseq, long long H, long long D,
                long long heads, long long q_rows, void* dqkv, long long lddqkv, cudaStream_t stream);
int attn_bwd(const void* qkv, long long ld, const void* dout, long long lddo, long long nseq, long long H, long long D,
             long long heads, long long q_rows, void* dqkv, long long lddqkv, cudaStream_t stream);

struct CastProblem {
  const float* src;
  long long rows, cols, ld_src;
  void* dst;  // bf16
  long long ld_dst, dst_cols;
};
struct GatherProblem {
  const float* table;
  long long table_rows, dim;
  const long long* ids;
  long long n;
  void* dst;  // bf16
  long long ld_dst;
};
int cast_rows_bf16_batched(const CastProblem* probs, int n, cudaStream_t stream);
int gather_rows_bf16_batched(const GatherProblem* probs, int n, int* oob_flag, cudaStream_t stream);

// Elementwise / gather / scatter helpers (elementwise.cu)
int cast_rows_bf16(const float* src, long long rows, long long cols, long long ld_src, void* dst, long long ld_dst,
                   long long dst_cols, cudaStream_t stream);
int gather_rows_bf16(const float* table, long long table_rows, long long dim, const long long* ids, long long n,
                     void* dst, long long ld_dst, int* oob_flag, cudaStream_t stream);
int gather_rows_f32(const float* table, long long table_rows, long long dim, const long long* ids, long long n,
                    float* dst, long long ld_dst, int* oob_flag, cudaStream_t stream);
int scatter_add_rows(const void* src16, const float* src32, long long ld_src, const long long* ids, long long n,
                     long long dim, float* table_grad, long long table_rows, cudaStream_t stream);
int colsum(const void* src16, const float* src32, long long rows, long long cols, long long ld, float* out,
           cudaStream_t stream);
int history_gather_pool(const float* table, long long table_rows, long long D, const long long* ids, long long B,
                        long long H, const float* pe, void* x16, long long ldx, float* mean, long long ldmean,
                        int* oob_flag, cudaStream_t stream);
int history_scatter_grad(const void* dx16, long long lddx, const float* dmean, long long lddmean,
                         const long long* ids, long long B, long long H, long long D, float* table_grad,
                         long long table_rows, cudaStream_t stream);

// Fused tower forward (tower.cu): [table[ids] | MLP(feats)] Wt^T + bt for up to 4 towers of one shape per launch.
struct TowerProblem {
  const long long* ids;
  const float* table;
  long long table_rows;
  const float* feats;
  long long ld_feats;
  const void* w0;  // bf16 [hidden, F]
  long long ldw0;
  const float* b0;
  const void* w1;  // bf16 [D, hidden]
  long long ldw1;
  const float* b1;
  const void* wt;  // bf16 [DI, 2D]
  long long ldwt;
  const float* bt;
  void* feats16;  // out: bf16 [rows, F]
  long long ld_feats16;
  void* h16;  // out: bf16 [rows, hidden]
  long long ldh;
  void* x16;  // out: bf16 [rows, 2D] = [id_emb | feat_emb]
  long long ldx;
  float* emb32;  // out: fp32 [rows, DI]
  long long ld_emb32;
  void* emb16;  // out: bf16 [rows, DI]
  long long ld_emb16;
  long long rows, F, D, DI, hidden;
};
// Fused tower backward chain (tower_bwd.cu, candidate): dX = demb Wt, dH = (dFe W1) masked, table scatter, db1 / db0.
struct TowerBwdProblem {
  const void* demb16;  // bf16 [rows, DI]
  long long ld_demb;
  const long long* ids;
  long long table_rows;
  const void* wt;  // bf16 [DI, 2D]
  long long ldwt;
  const void* w1;  // bf16 [D, hidden]
  long long ldw1;
  const void* h16;  // bf16 [rows, hidden]
  long long ldh;
  void* dx16;  // out bf16 [rows, 2D]
  long long lddx;
  void* dh16;  // out bf16 [rows, hidden]
  long long lddh;
  float* dtable;  // += [table_rows, D] or null
  float* dxsum;   // += [2D] (feature half only)
  float* db0;     // += [hidden]
  long long rows, D, DI, hidden;
};
int tower_bwd_chain(const TowerBwdProblem* problems, int n, cudaStream_t stream);
bool tower_fwd_supported(long long F, long long D, long long DI, long long hidden);
int tower_fwd(const TowerProblem* problems, int n, int* oob_flag, cudaStream_t stream);

// Last history-encoder layer evaluated for query row 0 only, against the raw layer input (csrc/history_last.cu)
int history_last_supported(long long H, long long D, long long heads);
int history_last_fwd(const void* x16, long long ldx, const float* qt, long long B, long long H, long long D, long long heads,
                     void* z16, float* p32, cudaStream_t stream);
int history_last_bwd1(const void* x16, long long ldx, const float* dz, const float* p32, long long B, long long H, long long D,
                      long long heads, float* ds32, void* dqt16, cudaStream_t stream);
int history_last_bwd2(const float* dz, const float* qt, const float* p32, const float* ds32, const float* extra, long long B,
                      long long H, long long D, long long heads, void* dx16, long long lddx, float* colsum, cudaStream_t stream);

// Fused multi-tensor Adam (torch.optim.Adam arithmetic, amsgrad off); *step is a device counter bumped by the launch.
struct AdamTensor {
  float* p;
  const float* g;
  float* m;
  float* v;
  long long n;
  int vec4;  // filled by adam_step
};
int adam_step(const AdamTensor* tensors, int n, double lr, double beta1, double beta2, float eps, float weight_decay,
              long long* step, unsigned int* ticket, cudaStream_t stream);

int weighted_loss(const float* ce, const float* labels, long long ldl, const float* uvw, long long B, long long T,
                  float* loss, float* g, cudaStream_t stream);

}  // namespace tt
