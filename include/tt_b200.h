/* tt_b200.h - C ABI of the B200-native two-tower hot path (libtt_b200.so).
 *
 * The reference (gauravchak/two_tower_models) has no FFI of its own: its hot path is a set of PyTorch
 * library calls inside nn.Module methods.  Each entry point below replaces one such call site; the
 * host-side mirror in two_tower_models_b200/ binds them with ctypes (see INTEGRATION.md for the stub a
 * reference maintainer would add).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the parameter name ends in _host;
 *   - matrices are row-major; `ld*` is the row pitch in ELEMENTS; bf16 operand matrices consumed by
 *     tensor-core kernels need 16-byte aligned bases and pitches that are multiples of 8 elements;
 *   - ids are int64 (the reference's dtype), floats are fp32, "bf16" is stored as uint16 bits;
 *   - `stream` is a cudaStream_t; all work is enqueued on it, nothing synchronises the host, so every
 *     entry point is CUDA-graph capturable;
 *   - inputs are borrowed and never written; outputs and workspaces are caller-owned;
 *   - return value 0 = success, negative = error, message via tt_last_error() (thread-local).
 *     There is no CPU fallback: without a CUDA device every compute entry point fails.
 */
#ifndef TT_B200_H
#define TT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TT_B200_ABI_VERSION 1

int tt_abi_version(void);
const char* tt_last_error(void);
/* Number of SMs of the current device (148 on B200); negative on error. */
int tt_device_sm_count(void);
/* Total number of kernels this library has launched in this process (for bench.py's gpu_launches). */
long long tt_launch_count(void);

/* Optional per-kernel device timing for benchmarks: while enabled, every kernel this library launches
 * outside of stream capture is bracketed by CUDA events on its launching stream.  tt_profile_report
 * synchronises the device and writes one line per kernel name, "name total_ms launches\n", into buf. */
void tt_profile_enable(int on);
int tt_profile_report(char* buf, int64_t buf_bytes);
/* on = 2 also brackets launches made under stream capture (external event-record nodes in the CUDA graph): every
 * replay of such a graph re-stamps its events, and tt_profile_report_graph reports the spans of the LAST replay
 * (same line format; clear != 0 forgets the captured spans - call it before the graph is destroyed). */
int tt_profile_report_graph(char* buf, int64_t buf_bytes, int32_t clear);
/* Launches an EMPTY kernel inside a span named "null_kernel": its reported time is what an event pair adds to every
 * span taken the same way (eager or inside a graph), i.e. the calibration of the per-kernel numbers. */
int tt_profile_null_span(void* stream);

/* ---- data movement (HBM-bound helpers) ------------------------------------------------------ */

/* dst[r, 0:dst_cols] = bf16(src[r, 0:cols]) zero-padded.  Packs fp32 features / weights for the
 * tensor-core kernels (inputs of nn.Linear, reference src/two_tower_base_retrieval.py:76-80). */
int tt_cast_rows_bf16(const float* src, int64_t rows, int64_t cols, int64_t ld_src, void* dst_bf16, int64_t ld_dst,
                      int64_t dst_cols, void* stream);

/* nn.Embedding lookup (reference :126, :209): dst[i, 0:dim] = table[ids[i], :].
 * Out-of-range ids are clamped and *oob_flag (may be NULL) is set to 1. */
int tt_gather_rows_bf16(const float* table, int64_t table_rows, int64_t dim, const int64_t* ids, int64_t n,
                        void* dst_bf16, int64_t ld_dst, int32_t* oob_flag, void* stream);
int tt_gather_rows_f32(const float* table, int64_t table_rows, int64_t dim, const int64_t* ids, int64_t n,
                       float* dst, int64_t ld_dst, int32_t* oob_flag, void* stream);

/* Dense embedding gradient (autograd of nn.Embedding): grad[ids[i], :] += src[i, :].
 * Exactly one of src_bf16 / src_f32 is non-NULL. */
int tt_scatter_add_rows(const void* src_bf16, const float* src_f32, int64_t ld_src, const int64_t* ids, int64_t n,
                        int64_t dim, float* table_grad, int64_t table_rows, void* stream);

/* Bias gradient: out[c] += sum_r src[r, c]. */
int tt_colsum(const void* src_bf16, const float* src_f32, int64_t rows, int64_t cols, int64_t ld, float* out,
              void* stream);

/* ---- tcgen05 GEMM --------------------------------------------------------------------------- */

/* C[M,N] (+)= alpha * A * B^T (+ bias[N]) (ReLU) (zeroed where relu_mask <= 0).
 * A: bf16 [M,K] pitch lda, or (a_mn_major) stored as [K,M] pitch lda.  B: bf16 [N,K] or (b_mn_major) [K,N].
 * Outputs: c_f32 (pitch ldc_f32) and/or c_bf16 (pitch ldc_bf16).  accumulate != 0: split-K with fp32
 * atomicAdd into c_f32 (no bias/relu/mask/bf16 output; caller initialises c_f32); split_k 0 = auto.
 * colsum_f32 (may be NULL): [N] fp32, += sum over rows of the final C values taken from the fp32
 * accumulators (bias gradients; caller initialises).
 * Replaces nn.Linear forward/backward (reference :76-80, :90-93, :101-110) and the MHA in/out
 * projections (src/user_history_encoder.py:60-67). */
int tt_gemm_bf16(const void* A, int64_t lda, int32_t a_mn_major, const void* B, int64_t ldb, int32_t b_mn_major,
                 int64_t M, int64_t N, int64_t K, const float* bias, int32_t relu, const void* relu_mask_bf16,
                 int64_t ld_mask, float alpha, float* c_f32, int64_t ldc_f32, void* c_bf16, int64_t ldc_bf16,
                 int32_t accumulate, int32_t split_k, float* colsum_f32, void* stream);

/* Several independent GEMMs in one launch (e.g. the same layer of the user tower and of the item tower):
 * consecutive problems that select the same tile width share a launch (up to 4), the rest follow in further
 * launches.  Field meanings as in tt_gemm_bf16. */
typedef struct tt_gemm_problem {
  const void* A;
  int64_t lda;
  const void* B;
  int64_t ldb;
  int64_t M, N, K;
  const float* bias;
  const void* relu_mask_bf16;
  int64_t ld_mask;
  float* c_f32;
  int64_t ldc_f32;
  void* c_bf16;
  int64_t ldc_bf16;
  float* colsum_f32;
  float alpha;
  int32_t a_mn_major, b_mn_major, relu, accumulate, split_k;
} tt_gemm_problem;
int tt_gemm_bf16_batched(const tt_gemm_problem* problems, int32_t count, void* stream);

/* Batched forms of tt_cast_rows_bf16 (count <= 16) and tt_gather_rows_bf16 (count <= 8): one launch. */
typedef struct tt_cast_problem {
  const float* src;
  int64_t rows, cols, ld_src;
  void* dst_bf16;
  int64_t ld_dst, dst_cols;
} tt_cast_problem;
int tt_cast_rows_bf16_batched(const tt_cast_problem* problems, int32_t count, void* stream);
typedef struct tt_gather_problem {
  const float* table;
  int64_t table_rows, dim;
  const int64_t* ids;
  int64_t n;
  void* dst_bf16;
  int64_t ld_dst;
} tt_gather_problem;
int tt_gather_rows_bf16_batched(const tt_gather_problem* problems, int32_t count, int32_t* oob_flag, void* stream);

/* ---- fused tower forward -------------------------------------------------------------------- */

/* emb = [table[ids] | Linear(hidden, D)(relu(Linear(F, hidden)(feats)))] Wt^T + bt for `count` (<= 4) towers of the same
 * shape in ONE launch: the id lookup, both MLP layers, the concatenation (a K-split of the last GEMM) and the tower
 * Linear run on chip per 128-row tile, hidden activations stay in tensor memory.  Replaces compute_user_embedding /
 * compute_item_embeddings of the base model (reference src/two_tower_base_retrieval.py:112-219).
 * Weights are the bf16 copies (w0 [hidden, F], w1 [D, hidden], wt [DI, 2D]); biases fp32.  Besides emb (fp32 and bf16)
 * the kernel writes the bf16 activations the backward pass needs: feats_bf16 [rows, F], h_bf16 [rows, hidden],
 * x_bf16 [rows, 2D] = [id_emb | feat_emb].  Supported shapes: tt_tower_fwd_supported() (F = D = DI in {64, 128},
 * hidden = 256); other shapes use the per-layer entry points above. */
typedef struct tt_tower_problem {
  const int64_t* ids;
  const float* table;
  int64_t table_rows;
  const float* feats;
  int64_t ld_feats;
  const void* w0_bf16;
  int64_t ldw0;
  const float* b0;
  const void* w1_bf16;
  int64_t ldw1;
  const float* b1;
  const void* wt_bf16;
  int64_t ldwt;
  const float* bt;
  void* feats_bf16;
  int64_t ld_feats16;
  void* h_bf16;
  int64_t ldh;
  void* x_bf16;
  int64_t ldx;
  float* emb_f32;
  int64_t ld_emb;
  void* emb_bf16;
  int64_t ld_emb16;
  int64_t rows, F, D, DI, hidden;
} tt_tower_problem;
int32_t tt_tower_fwd_supported(int64_t F, int64_t D, int64_t DI, int64_t hidden);
int tt_tower_fwd(const tt_tower_problem* problems, int32_t count, int32_t* oob_flag, void* stream);

/* CANDIDATE (not on the default path yet): the activation-gradient chain of the tower backward in one
 * launch for `count` (<= 4) towers of a shape tt_tower_fwd_supported() accepts (F = D): dx_bf16 = demb Wt (both halves),
 * table_grad[ids[r], :] += dX[r, 0:D] in fp32 (skipped when table_grad is NULL), dh_bf16 = (dX[:, D:2D] W1) where h > 0,
 * dxsum[D + c] += sum_r dX[r, D + c] (bias gradient of MLP layer 1), db0[c] += sum_r dH[r, c].  The weight gradients
 * stay tt_gemm_bf16 (split-K) launches. */
typedef struct tt_tower_bwd_problem {
  const void* demb_bf16;
  int64_t ld_demb;
  const int64_t* ids;
  int64_t table_rows;
  const void* wt_bf16;
  int64_t ldwt;
  const void* w1_bf16;
  int64_t ldw1;
  const void* h_bf16;
  int64_t ldh;
  void* dx_bf16;
  int64_t lddx;
  void* dh_bf16;
  int64_t lddh;
  float* table_grad;
  float* dxsum;
  float* db0;
  int64_t rows, D, DI, hidden;
} tt_tower_bwd_problem;
int tt_tower_bwd_chain(const tt_tower_bwd_problem* problems, int32_t count, void* stream);

/* ---- in-batch sampled-softmax loss ---------------------------------------------------------- */

/* Scratch bytes needed by tt_inbatch_ce_fwd / _bwd for this shape on the current device. */
int64_t tt_inbatch_ce_workspace_bytes(int64_t B, int64_t N, int64_t d);

/* ce[i] = logsumexp_j(U_i . V_j) - U_i . V_{i+target_offset},  lse[i] = logsumexp_j(U_i . V_j).
 * U: bf16 [B,d], V: bf16 [N,d].  Replaces torch.matmul + F.cross_entropy(reduction="none") with
 * target = arange (reference src/two_tower_base_retrieval.py:287, :301, :310-312); target_offset is the
 * multi-GPU generalisation (local users scored against all-gathered items).  d <= 256. */
int tt_inbatch_ce_fwd(const void* U_bf16, int64_t ldu, const void* V_bf16, int64_t ldv, int64_t B, int64_t N, int64_t d,
                      int64_t target_offset, float* ce, float* lse, void* workspace, int64_t workspace_bytes,
                      void* stream);

/* Backward for upstream g[i] = dL/dce[i]:  dS = g_i (softmax(S)_ij - [j == i+off]);  dU = dS V;  dV = dS^T U.
 * Any of dU_f32/dU_bf16/dV_f32/dV_bf16 may be NULL (a NULL pair skips that pass).
 * dU_colsum / dV_colsum (may be NULL, d <= 128): [d] fp32, += column sums of dU / dV (the bias gradients of the
 * tower Linear layers that produced U and V); the caller initialises them. */
int tt_inbatch_ce_bwd(const void* U_bf16, int64_t ldu, const void* V_bf16, int64_t ldv, int64_t B, int64_t N, int64_t d,
                      int64_t target_offset, const float* lse, const float* g, float* dU_f32, int64_t lddu,
                      void* dU_bf16, int64_t lddu16, float* dV_f32, int64_t lddv, void* dV_bf16, int64_t lddv16,
                      float* dU_colsum, float* dV_colsum, void* workspace, int64_t workspace_bytes, void* stream);

/* Value-weighted mean of the per-row loss with the identity debias hook (reference
 * src/two_tower_base_retrieval.py:322-343): nuv_i = sum_t labels[i,t] * weights[t]; w_i = max(nuv_i, 1e-6) /
 * max_i(max(nuv_i, 1e-6)); *loss = sum_i ce[i] w_i / B; g[i] = d loss / d ce[i] = w_i / B.  One launch. */
int tt_weighted_loss(const float* ce, const float* labels, int64_t ld_labels, const float* weights, int64_t B, int64_t T,
                     float* loss, float* g, void* stream);

/* tt_inbatch_ce_fwd and tt_weighted_loss in two launches instead of three: the kernel that merges the CE partials
 * also forms nuv_i = max(labels_i . weights, 1e-6) and reduces the weighted mean - the whole of compute_training_loss
 * with the identity hook (reference src/two_tower_base_retrieval.py:279-347).  Writes ce[B], lse[B], *loss,
 * g[i] = nuv_i (UNnormalised) and *g_norm = 1 / (max_i nuv_i * B), so that d loss / d ce[i] = g[i] * *g_norm. */
int tt_inbatch_ce_loss_fwd(const void* U_bf16, int64_t ldu, const void* V_bf16, int64_t ldv, int64_t B, int64_t N,
                           int64_t d, int64_t target_offset, const float* labels, int64_t ld_labels,
                           const float* weights, int64_t T, float* ce, float* lse, float* loss, float* g, float* g_norm,
                           void* workspace, int64_t workspace_bytes, void* stream);
/* Batch-sharded form of tt_inbatch_ce_loss_fwd (one process per GPU, local users against all-gathered items, positives at
 * column row + target_offset): writes ce[B], lse[B], g[i] = nuv_i and stats[0..1] = (max_i nuv_i, sum_i ce_i nuv_i) of
 * THIS rank (stats needs room for 4 floats; [2..3] are scratch).  The per-rank pairs are all-gathered by the caller and
 * folded by tt_sharded_loss_finalize into *loss = sum_r s_r / (max_r m_r * global_rows) and *g_norm = 1 / (max * rows),
 * the reference's batch-global max and mean (src/two_tower_base_retrieval.py:339-343) over the concatenated batch. */
int tt_inbatch_ce_loss_fwd_sharded(const void* U_bf16, int64_t ldu, const void* V_bf16, int64_t ldv, int64_t B, int64_t N,
                                   int64_t d, int64_t target_offset, const float* labels, int64_t ld_labels,
                                   const float* weights, int64_t T, float* ce, float* lse, float* g, float* stats,
                                   void* workspace, int64_t workspace_bytes, void* stream);
int tt_sharded_loss_finalize(const float* stats_all, int32_t world, int64_t global_rows, float* loss, float* g_norm,
                             void* stream);
/* tt_inbatch_ce_bwd with g multiplied by the DEVICE scalars *g_scale and *g_scale2 (each may be NULL = 1): the
 * incoming gradient of the scalar loss and *g_norm above are applied inside the kernels instead of by separate
 * elementwise launches. */
int tt_inbatch_ce_bwd_scaled(const void* U_bf16, int64_t ldu, const void* V_bf16, int64_t ldv, int64_t B, int64_t N,
                             int64_t d, int64_t target_offset, const float* lse, const float* g, const float* g_scale,
                             const float* g_scale2, float* dU_f32, int64_t lddu, void* dU_bf16, int64_t lddu16,
                             float* dV_f32, int64_t lddv, void* dV_bf16, int64_t lddv16, float* dU_colsum,
                             float* dV_colsum, void* workspace, int64_t workspace_bytes, void* stream);

/* ---- optimizer (the adjacent step of the training loop) --------------------------- */

/* One launch of torch.optim.Adam's update (amsgrad off; reference train/train.py:179, :123-125) over `count` <= 32
 * fp32 tensors: exp_avg += (grad - exp_avg)(1 - beta1); exp_avg_sq = beta2 exp_avg_sq + (1 - beta2) grad^2;
 * param -= lr / (1 - beta1^t) * exp_avg / (sqrt(exp_avg_sq) / sqrt(1 - beta2^t) + eps), t = *step_dev + 1;
 * weight_decay != 0 adds weight_decay * param to grad first.  *step_dev (int64, device) is incremented by the launch,
 * ticket_dev is a zero-initialised uint32 scratch word owned by the optimizer (both make CUDA-graph replay work). */
typedef struct tt_adam_tensor {
  float* param;
  const float* grad;
  float* exp_avg;
  float* exp_avg_sq;
  int64_t numel;
} tt_adam_tensor;
int tt_adam_step(const tt_adam_tensor* tensors, int32_t count, double lr, double beta1, double beta2, float eps,
                 float weight_decay, int64_t* step_dev, uint32_t* ticket_dev, void* stream);

/* ---- brute-force MIPS ------------------------------------------------------------------------ */

/* Scratch bytes needed by tt_mips_topk for this shape on the current device. */
int64_t tt_mips_workspace_bytes(int64_t nq, int64_t nc, int64_t d, int64_t k);

/* idx[q, 0:k], scores[q, 0:k] = top-k of Q C^T per query row, scores descending, exact ties by ascending
 * corpus index.  Replaces torch.topk(torch.matmul(query, corpus.T), k) (reference
 * src/baseline_mips_module.py:57-61).  Q_bf16 [nq,d] / C_bf16 [nc,d] are bf16 copies used to screen the
 * corpus on the tensor cores (no [nq,nc] matrix is materialised); the best k + margin candidates are then
 * re-scored in fp32 from Q_f32 / C_f32, which defines the returned scores and order.
 * d <= 128, k <= 224, nc < 2^32 - 1. */
int tt_mips_topk(const void* Q_bf16, int64_t ldq16, const void* C_bf16, int64_t ldc16, const float* Q_f32,
                 int64_t ldq32, const float* C_f32, int64_t ldc32, int64_t nq, int64_t nc, int64_t d, int64_t k,
                 int64_t* idx, float* scores, void* workspace, int64_t workspace_bytes, void* stream);

/* The same two entry points with the item matrix V given as n_parts equally sized row blocks
 * V = [V_parts[0]; V_parts[1]; ...] (rows_per_part rows each, a multiple of 128; n_parts <= 8).  The blocks may live in
 * the memory of peer GPUs (mapped through CUDA IPC / symmetric memory): the kernels read them in place over NVLink
 * with TMA, so the data-parallel loss needs no all-gather of the item embeddings.  V_parts is a HOST array of
 * device pointers. */
int tt_inbatch_ce_fwd_parts(const void* U_bf16, int64_t ldu, const void* const* V_parts_host, int32_t n_parts,
                            int64_t rows_per_part, int64_t ldv, int64_t B, int64_t N, int64_t d, int64_t target_offset,
                            float* ce, float* lse, void* workspace, int64_t workspace_bytes, void* stream);
int tt_inbatch_ce_bwd_parts(const void* U_bf16, int64_t ldu, const void* const* V_parts_host, int32_t n_parts,
                            int64_t rows_per_part, int64_t ldv, int64_t B, int64_t N, int64_t d, int64_t target_offset,
                            const float* lse, const float* g, float* dU_f32, int64_t lddu, void* dU_bf16, int64_t lddu16,
                            float* dV_f32, int64_t lddv, void* dV_bf16, int64_t lddv16, float* dU_colsum,
                            float* dV_colsum, void* workspace, int64_t workspace_bytes, void* stream);

/* The next tt_inbatch_ce_fwd / tt_inbatch_ce_loss_fwd* launch of the calling thread also fills up to two device buffers
 * with zeros (16-byte aligned, sizes multiples of 16 bytes; p1 may be NULL): an otherwise idle warp of the scoring kernel
 * streams them out with TMA bulk stores while the score tiles run.  Used for the dense embedding-table gradients of the
 * training step (the reference's nn.Embedding(sparse=False) semantics: 2 x hash x D floats zero-filled per step), which
 * would otherwise cost a ~8 us memset between the tower and the scoring kernels. */
int tt_inbatch_ce_attach_zero_fill(void* p0, int64_t bytes0, void* p1, int64_t bytes1);

/* Limit the number of SMs the persistent kernels launched AFTER this call size their grids for (0 = all SMs); returns the
 * previous limit.  Used to leave SMs to a collective that runs beside a kernel: the batch-sharded loss starts the NCCL
 * reduce-scatter of dV and runs the dU pass on the remaining SMs (a 148-CTA persistent kernel would otherwise wait for the
 * SMs NCCL holds and finish late by the collective's duration).  Host-side state, not thread safe. */
int tt_set_sm_limit(int32_t n);

/* ---- last history-encoder layer, query row 0 only ------------------------------------------ */

/* The reference consumes row 0 of the last nn.MultiheadAttention layer only (src/user_history_encoder.py:116).  For a
 * single query row the keys / values need not be projected: with qt_h = c Wk_h^T q0_h (c = head_dim^-1/2) the scores are
 * qt_h . x_j against the RAW layer input x and the head output is Wv_h (sum_j p_j x_j) + bv_h.  These three entry points
 * are the memory-bound part (one warp per sequence); the [B, .]-sized projections around them are tt_gemm_bf16 calls.
 *   fwd : p[b,h,:] = softmax_j(qt[b,h,:] . x[b*H+j,:]);  z[b,h,:] = sum_j p[b,h,j] x[b*H+j,:]   (z bf16 [B, heads*D])
 *   bwd1: ds = p (dz . x - sum p (dz . x));  dqt[b,h,:] = sum_j ds[b,h,j] x[b*H+j,:]               (dqt bf16)
 *   bwd2: dx[b*H+j,:] = sum_h (p[b,h,j] dz[b,h,:] + ds[b,h,j] qt[b,h,:]) (+ extra[b,:] on row j = 0); colsum += sum rows
 * D in {64, 128}, H <= 128, heads <= 8, head_dim % 8 == 0 (tt_history_last_supported). */
int tt_history_last_supported(int64_t H, int64_t D, int64_t heads);
int tt_history_last_fwd(const void* x_bf16, int64_t ldx, const float* qt, int64_t B, int64_t H, int64_t D, int64_t heads,
                        void* z_bf16, float* p, void* stream);
int tt_history_last_bwd1(const void* x_bf16, int64_t ldx, const float* dz, const float* p, int64_t B, int64_t H, int64_t D,
                         int64_t heads, float* ds, void* dqt_bf16, void* stream);
int tt_history_last_bwd2(const float* dz, const float* qt, const float* p, const float* ds, const float* extra, int64_t B,
                         int64_t H, int64_t D, int64_t heads, void* dx_bf16, int64_t lddx, float* colsum, void* stream);

/* ---- history encoder helpers ---------------------------------------------------------------- */

/* x_bf16[b*H+h, :] = bf16(table[ids[b,h]] + pe[h]) (pe may be NULL);  mean[b, :] = mean_h table[ids[b,h]].
 * Reference: src/two_tower_with_user_history_encoder.py:105, src/user_history_encoder.py:89-95. */
int tt_history_gather_pool(const float* table, int64_t table_rows, int64_t D, const int64_t* ids, int64_t B, int64_t H,
                           const float* pe, void* x_bf16, int64_t ldx, float* mean, int64_t ldmean, int32_t* oob_flag,
                           void* stream);
/* table_grad[ids[b,h], :] += dx_bf16[b*H+h, :] + dmean[b, :] / H   (either source may be NULL). */
int tt_history_scatter_grad(const void* dx_bf16, int64_t lddx, const float* dmean, int64_t lddmean, const int64_t* ids,
                            int64_t B, int64_t H, int64_t D, float* table_grad, int64_t table_rows, void* stream);

/* Self-attention core of nn.MultiheadAttention (reference src/user_history_encoder.py:60-67, 103-108;
 * torch.nn.functional.multi_head_attention_forward): per sequence and head,
 * out = softmax(q k^T / sqrt(D/heads)) v, no mask, no dropout.  qkv_bf16: [nseq*H, ldqkv] rows = [q | k | v]
 * (3D columns, head h in columns [h*D/heads, (h+1)*D/heads) of each block), i.e. the output of the packed
 * in-projection.  Only the first q_rows query rows of every sequence are produced: out_bf16 is
 * [nseq*q_rows, ldo] (the encoder consumes row 0 of its last layer only, reference :116). */
int tt_attn_fwd(const void* qkv_bf16, int64_t ldqkv, int64_t nseq, int64_t H, int64_t D, int64_t heads, int64_t q_rows,
                void* out_bf16, int64_t ldo, void* stream);
/* dqkv_bf16 [nseq*H, lddqkv] from dout_bf16 [nseq*q_rows, lddo] (softmax recomputed from qkv). */
int tt_attn_bwd(const void* qkv_bf16, int64_t ldqkv, const void* dout_bf16, int64_t lddo, int64_t nseq, int64_t H,
                int64_t D, int64_t heads, int64_t q_rows, void* dqkv_bf16, int64_t lddqkv, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TT_B200_H */
