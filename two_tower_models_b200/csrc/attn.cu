// Self-attention core of the history encoder: per (sequence, head)  O = softmax(Q K^T / sqrt(hd)) V  and
// its backward.  Reference: nn.MultiheadAttention as used by src/user_history_encoder.py:60-67,103-108
// (torch.nn.functional.multi_head_attention_forward: q scaled by head_dim^-0.5, softmax over keys, no mask,
// no dropout).  The packed in-projection and the out-projection run on the tcgen05 GEMM (gemm.cu); this
// file handles the H x H part, which at H <= 128 and head_dim <= 64 is HBM / latency bound: one CTA per
// sequence stages that sequence's q|k|v rows (bf16, 16-byte vector loads) into shared memory as fp32, one
// warp per head; every lane owns query rows (forward, dQ) or key rows (dK, dV) in registers and streams the
// other side from shared memory as warp-wide broadcasts, so there are no bank conflicts and no score matrix.
//
// qkv layout: [nseq*H, ld] bf16 with q at column 0, k at column D, v at column 2D, head h in columns
// [h*hd, (h+1)*hd) of each block.  `q_rows` limits the query rows computed per sequence (the last encoder
// layer only needs row 0: src/user_history_encoder.py:116).
#include <stdlib.h>

#include "common.cuh"
#include "kernels.h"

namespace tt {

namespace {

constexpr float LOG2E_F = 1.4426950408889634f;

// shared-memory row pitches (fp32 words) keep every head's slice 16-byte aligned when D and D/heads are
// multiples of 4; the +4 staggers consecutive rows over the banks
__host__ __device__ constexpr int round4(int x) { return (x + 3) / 4 * 4; }
__host__ __device__ inline int attn_pitch(int D) { return 2 * D + round4(D) + 4; }

__device__ __forceinline__ void load_rows_f32(float* dst, int dst_pitch, const bf16* src, long long ld, int rows,
                                              int cols, int tid, int nthreads) {
  // cols is a multiple of 8 when vec is true (16-byte aligned rows)
  const bool vec = (cols % 8 == 0) && (ld % 8 == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
  if (vec) {
    const int cpr = cols / 8;
    for (int i = tid; i < rows * cpr; i += nthreads) {
      const int r = i / cpr, c = (i % cpr) * 8;
      const uint4 v = *reinterpret_cast<const uint4*>(src + (long long)r * ld + c);
      const bf16* b = reinterpret_cast<const bf16*>(&v);
#pragma unroll
      for (int e = 0; e < 8; ++e) dst[r * dst_pitch + c + e] = __bfloat162float(b[e]);
    }
  } else {
    for (int i = tid; i < rows * cols; i += nthreads) {
      const int r = i / cols, c = i % cols;
      dst[r * dst_pitch + c] = __bfloat162float(src[(long long)r * ld + c]);
    }
  }
}

// dot product of a register vector with a shared-memory row, 4 independent accumulation chains (the lanes
// of a warp run in lock step on different rows, so instruction-level parallelism is the only latency hiding)
template <int HD>
__device__ __forceinline__ float dot_reg_smem(const float* a, const float* __restrict__ b, int hd) {
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  if ((hd & 3) == 0) {  // rows are 16-byte aligned (pitch % 4 == 0): one broadcast LDS.128 per 4 FMAs
#pragma unroll
    for (int c = 0; c < HD; c += 4) {
      if (c < hd) {
        const float4 v = *reinterpret_cast<const float4*>(b + c);
        s0 = fmaf(a[c], v.x, s0);
        s1 = fmaf(a[c + 1], v.y, s1);
        s2 = fmaf(a[c + 2], v.z, s2);
        s3 = fmaf(a[c + 3], v.w, s3);
      }
    }
  } else {
#pragma unroll
    for (int c = 0; c < HD; c += 4) {
      if (c < hd) s0 = fmaf(a[c], b[c], s0);
      if (c + 1 < hd) s1 = fmaf(a[c + 1], b[c + 1], s1);
      if (c + 2 < hd) s2 = fmaf(a[c + 2], b[c + 2], s2);
      if (c + 3 < hd) s3 = fmaf(a[c + 3], b[c + 3], s3);
    }
  }
  return (s0 + s1) + (s2 + s3);
}

// ---------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------
template <int HD>  // head dim padded to HD (zero columns)
__global__ void __launch_bounds__(128)
attn_fwd_kernel(const bf16* __restrict__ qkv, long long ld, int H, int D, int heads, int hd, int q_rows,
                bf16* __restrict__ out, long long ldo, float scale_log2) {
  extern __shared__ float sm[];
  const int pitch = attn_pitch(D);  // fp32 words, multiple of 4: broadcast rows can be read as float4
  const long long seq = blockIdx.x;
  const bf16* base = qkv + seq * H * ld;
  load_rows_f32(sm, pitch, base, ld, H, 3 * D, threadIdx.x, blockDim.x);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int h = warp; h < heads; h += nwarps) {
    const float* Q = sm + h * hd;
    const float* K = sm + D + h * hd;
    const float* V = sm + 2 * D + h * hd;
    for (int i = lane; i < q_rows; i += 32) {
      float q[HD], o[HD];
#pragma unroll
      for (int c = 0; c < HD; ++c) {
        q[c] = c < hd ? Q[i * pitch + c] : 0.f;
        o[c] = 0.f;
      }
      float m = -INFINITY, l = 0.f;
      int j = 0;
      for (; j + 1 < H; j += 2) {  // two keys per step: independent dot products, one rescale
        const float s0 = dot_reg_smem<HD>(q, K + j * pitch, hd) * scale_log2;
        const float s1 = dot_reg_smem<HD>(q, K + (j + 1) * pitch, hd) * scale_log2;
        const float mn = fmaxf(m, fmaxf(s0, s1));
        const float corr = ex2f(m - mn);
        const float p0 = ex2f(s0 - mn), p1 = ex2f(s1 - mn);
        l = l * corr + (p0 + p1);
#pragma unroll
        for (int c = 0; c < HD; ++c)
          if (c < hd) o[c] = fmaf(p1, V[(j + 1) * pitch + c], fmaf(p0, V[j * pitch + c], o[c] * corr));
        m = mn;
      }
      if (j < H) {
        const float s0 = dot_reg_smem<HD>(q, K + j * pitch, hd) * scale_log2;
        const float mn = fmaxf(m, s0);
        const float corr = ex2f(m - mn);
        const float p0 = ex2f(s0 - mn);
        l = l * corr + p0;
#pragma unroll
        for (int c = 0; c < HD; ++c)
          if (c < hd) o[c] = fmaf(p0, V[j * pitch + c], o[c] * corr);
        m = mn;
      }
      const float inv = 1.f / l;
      bf16* dst = out + (seq * q_rows + i) * ldo + h * hd;
#pragma unroll
      for (int c = 0; c < HD; ++c)
        if (c < hd) dst[c] = __float2bfloat16(o[c] * inv);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// backward: dq (lanes <-> query rows), then dk / dv (lanes <-> key rows); softmax recomputed
// ---------------------------------------------------------------------------------------------
template <int HD>
__global__ void __launch_bounds__(128)
attn_bwd_kernel(const bf16* __restrict__ qkv, long long ld, const bf16* __restrict__ dout, long long lddo, int H, int D,
                int heads, int hd, int q_rows, bf16* __restrict__ dqkv, long long lddqkv, float scale, float scale_log2) {
  extern __shared__ float sm[];
  const int pitch = attn_pitch(D);
  const int dpitch = attn_pitch(D) - 2 * D;  // == round4(D) + 4
  float* sdo = sm + H * pitch;                 // [q_rows][D+1] upstream gradient
  float* slse = sdo + q_rows * dpitch;         // [heads][q_rows] log2-sum-exp of the scaled scores
  float* sdelta = slse + heads * q_rows;       // [heads][q_rows] sum_c dO_ic O_ic
  const long long seq = blockIdx.x;
  load_rows_f32(sm, pitch, qkv + seq * H * ld, ld, H, 3 * D, threadIdx.x, blockDim.x);
  load_rows_f32(sdo, dpitch, dout + seq * q_rows * lddo, lddo, q_rows, D, threadIdx.x, blockDim.x);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  bf16* dbase = dqkv + seq * H * lddqkv;
  for (int h = warp; h < heads; h += nwarps) {
    const float* Q = sm + h * hd;
    const float* K = sm + D + h * hd;
    const float* V = sm + 2 * D + h * hd;
    const float* dO = sdo + h * hd;
    // phase A: per query row  lse, delta, dq
    for (int i0 = 0; i0 < H; i0 += 32) {
      const int i = i0 + lane;
      if (i < q_rows) {
        float q[HD], g[HD], acc[HD];
#pragma unroll
        for (int c = 0; c < HD; ++c) {
          q[c] = c < hd ? Q[i * pitch + c] : 0.f;
          g[c] = c < hd ? dO[i * dpitch + c] : 0.f;
          acc[c] = 0.f;
        }
        float m = -INFINITY, l = 0.f;
        for (int j = 0; j < H; ++j) {
          const float s = dot_reg_smem<HD>(q, K + j * pitch, hd) * scale_log2;
          const float mn = fmaxf(m, s);
          const float corr = ex2f(m - mn);
          const float p = ex2f(s - mn);
          l = l * corr + p;
#pragma unroll
          for (int c = 0; c < HD; ++c)
            if (c < hd) acc[c] = fmaf(p, V[j * pitch + c], acc[c] * corr);
          m = mn;
        }
        const float inv = 1.f / l;
        float delta = 0.f;
#pragma unroll
        for (int c = 0; c < HD; ++c)
          if (c < hd) delta = fmaf(g[c], acc[c] * inv, delta);
        const float lse2 = m + log2f(l);
        slse[h * q_rows + i] = lse2;
        sdelta[h * q_rows + i] = delta;
#pragma unroll
        for (int c = 0; c < HD; ++c) acc[c] = 0.f;  // now dq
        for (int j = 0; j < H; ++j) {
          const float s = dot_reg_smem<HD>(q, K + j * pitch, hd);
          const float dp = dot_reg_smem<HD>(g, V + j * pitch, hd);
          const float p = ex2f(s * scale_log2 - lse2);
          const float ds = p * (dp - delta) * scale;
#pragma unroll
          for (int c = 0; c < HD; ++c)
            if (c < hd) acc[c] = fmaf(ds, K[j * pitch + c], acc[c]);
        }
        bf16* dst = dbase + (long long)i * lddqkv + h * hd;
#pragma unroll
        for (int c = 0; c < HD; ++c)
          if (c < hd) dst[c] = __float2bfloat16(acc[c]);
      } else if (i < H) {  // query rows that were not computed in the forward (last layer): dq = 0
        bf16* dst = dbase + (long long)i * lddqkv + h * hd;
        for (int c = 0; c < hd; ++c) dst[c] = __float2bfloat16(0.f);
      }
    }
    __syncwarp();
    // phase B: per key row  dk, dv
    for (int j0 = 0; j0 < H; j0 += 32) {
      const int j = j0 + lane;
      if (j < H) {
        float k[HD], v[HD], dk[HD], dv[HD];
#pragma unroll
        for (int c = 0; c < HD; ++c) {
          k[c] = c < hd ? K[j * pitch + c] : 0.f;
          v[c] = c < hd ? V[j * pitch + c] : 0.f;
          dk[c] = 0.f;
          dv[c] = 0.f;
        }
        for (int i = 0; i < q_rows; ++i) {
          const float s = dot_reg_smem<HD>(k, Q + i * pitch, hd);
          const float dp = dot_reg_smem<HD>(v, dO + i * dpitch, hd);
          const float p = ex2f(s * scale_log2 - slse[h * q_rows + i]);
          const float ds = p * (dp - sdelta[h * q_rows + i]) * scale;
#pragma unroll
          for (int c = 0; c < HD; ++c)
            if (c < hd) {
              dk[c] = fmaf(ds, Q[i * pitch + c], dk[c]);
              dv[c] = fmaf(p, dO[i * dpitch + c], dv[c]);
            }
        }
        bf16* dkp = dbase + (long long)j * lddqkv + D + h * hd;
        bf16* dvp = dbase + (long long)j * lddqkv + 2 * D + h * hd;
#pragma unroll
        for (int c = 0; c < HD; ++c)
          if (c < hd) {
            dkp[c] = __float2bfloat16(dk[c]);
            dvp[c] = __float2bfloat16(dv[c]);
          }
      }
    }
  }
}

static int pick_hd(long long hd) { return hd <= 4 ? 4 : (hd <= 8 ? 8 : (hd <= 16 ? 16 : (hd <= 32 ? 32 : 64))); }

}  // namespace

int attn_fwd(const void* qkv, long long ld, long long nseq, long long H, long long D, long long heads, long long q_rows,
             void* out, long long ldo, cudaStream_t stream) {
  TT_CHECK(nseq > 0 && H > 0 && D > 0 && heads > 0 && D % heads == 0, "attn_fwd: bad shape");
  TT_CHECK(q_rows > 0 && q_rows <= H, "attn_fwd: q_rows out of range");
  const long long hd = D / heads;
  TT_CHECK(hd <= 64, "attn_fwd: head_dim %lld > 64 is not supported", hd);
  static const bool use_tc = !(getenv("TT_ATTN_TC") && atoi(getenv("TT_ATTN_TC")) == 0);
  if (use_tc && attn_fwd_tc_supported(H, D, heads, ld, ldo, qkv, out))
    return attn_fwd_tc(qkv, ld, nseq, H, D, heads, q_rows, out, ldo, stream);
  const size_t smem = (size_t)H * attn_pitch((int)D) * sizeof(float);
  TT_CHECK(smem <= 200 * 1024, "attn_fwd: sequence tile H=%lld D=%lld does not fit shared memory", H, D);
  const float scale_log2 = LOG2E_F / sqrtf((float)hd);
#define TT_LAUNCH(HDV)                                                                                           \
  do {                                                                                                           \
    static bool configured = false;                                                                              \
    if (!configured) {                                                                                           \
      TT_CUDA(cudaFuncSetAttribute(attn_fwd_kernel<HDV>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); \
      configured = true;                                                                                         \
    }                                                                                                            \
    KernelSpan span("attn_fwd_kernel", stream);                                                                  \
    attn_fwd_kernel<HDV><<<(unsigned)nseq, 128, smem, stream>>>((const bf16*)qkv, ld, (int)H, (int)D, (int)heads, \
                                                                (int)hd, (int)q_rows, (bf16*)out, ldo, scale_log2); \
  } while (0)
  switch (pick_hd(hd)) {
    case 4: TT_LAUNCH(4); break;
    case 8: TT_LAUNCH(8); break;
    case 16: TT_LAUNCH(16); break;
    case 32: TT_LAUNCH(32); break;
    default: TT_LAUNCH(64); break;
  }
#undef TT_LAUNCH
  TT_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

int attn_bwd(const void* qkv, long long ld, const void* dout, long long lddo, long long nseq, long long H, long long D,
             long long heads, long long q_rows, void* dqkv, long long lddqkv, cudaStream_t stream) {
  TT_CHECK(nseq > 0 && H > 0 && D > 0 && heads > 0 && D % heads == 0, "attn_bwd: bad shape");
  TT_CHECK(q_rows > 0 && q_rows <= H, "attn_bwd: q_rows out of range");
  const long long hd = D / heads;
  static const bool use_tc = !(getenv("TT_ATTN_TC") && atoi(getenv("TT_ATTN_TC")) == 0);
  if (use_tc && attn_bwd_tc_supported(H, D, heads, ld, lddo, lddqkv, qkv, dout, dqkv))
    return attn_bwd_tc(qkv, ld, dout, lddo, nseq, H, D, heads, q_rows, dqkv, lddqkv, stream);
  TT_CHECK(hd <= 32, "attn_bwd: head_dim %lld > 32 is not supported", hd);
  const size_t smem = ((size_t)H * attn_pitch((int)D) + (size_t)q_rows * (attn_pitch((int)D) - 2 * D) + 2 * (size_t)heads * q_rows) * sizeof(float);
  TT_CHECK(smem <= 200 * 1024, "attn_bwd: sequence tile H=%lld D=%lld does not fit shared memory", H, D);
  const float scale = 1.f / sqrtf((float)hd);
  const float scale_log2 = LOG2E_F * scale;
#define TT_LAUNCH(HDV)                                                                                           \
  do {                                                                                                           \
    static bool configured = false;                                                                              \
    if (!configured) {                                                                                           \
      TT_CUDA(cudaFuncSetAttribute(attn_bwd_kernel<HDV>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); \
      configured = true;                                                                                         \
    }                                                                                                            \
    KernelSpan span("attn_bwd_kernel", stream);                                                                  \
    attn_bwd_kernel<HDV><<<(unsigned)nseq, 128, smem, stream>>>((const bf16*)qkv, ld, (const bf16*)dout, lddo, (int)H, \
                                                                (int)D, (int)heads, (int)hd, (int)q_rows,        \
                                                                (bf16*)dqkv, lddqkv, scale, scale_log2);        \
  } while (0)
  switch (pick_hd(hd)) {
    case 4: TT_LAUNCH(4); break;
    case 8: TT_LAUNCH(8); break;
    case 16: TT_LAUNCH(16); break;
    default: TT_LAUNCH(32); break;
  }
#undef TT_LAUNCH
  TT_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

}  // namespace tt
