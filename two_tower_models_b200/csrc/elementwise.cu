// HBM-bound helper kernels: fp32->bf16 row packing, embedding gathers, embedding-gradient scatter-add,
// column sums (bias gradients) and the history gather + positional-encoding + mean-pool.
// All accesses are coalesced along the feature dimension; ids are int64 like the reference's.
#include "common.cuh"
#include "kernels.h"

namespace tt {

// ---------------------------------------------------------------------------------------------
// fp32 [rows, cols] -> bf16 [rows, dst_cols] (zero padded on the right)
// ---------------------------------------------------------------------------------------------
__global__ void cast_rows_kernel(const float* __restrict__ src, long long rows, int cols, long long ld_src,
                                 bf16* __restrict__ dst, long long ld_dst, int dst_cols) {
  const int cpr = (dst_cols + 1) / 2;  // column pairs per row
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long row = idx / cpr;
  const int c = (int)(idx % cpr) * 2;
  if (row >= rows) return;
  const float* s = src + row * ld_src;
  const float a = c < cols ? s[c] : 0.f;
  const float b = c + 1 < cols ? s[c + 1] : 0.f;
  bf16* d = dst + row * ld_dst + c;
  if (c + 1 < dst_cols && ((ld_dst & 1) == 0)) {
    *reinterpret_cast<__nv_bfloat162*>(d) = __floats2bfloat162_rn(a, b);
  } else {
    d[0] = __float2bfloat16(a);
    if (c + 1 < dst_cols) d[1] = __float2bfloat16(b);
  }
}

int cast_rows_bf16(const float* src, long long rows, long long cols, long long ld_src, void* dst, long long ld_dst,
                   long long dst_cols, cudaStream_t stream) {
  TT_CHECK(rows >= 0 && cols >= 0 && dst_cols >= cols && ld_dst >= dst_cols, "cast_rows_bf16: bad shape");
  if (rows == 0 || dst_cols == 0) return 0;
  TT_CHECK(((uintptr_t)dst % 4) == 0, "cast_rows_bf16: dst must be 4-byte aligned");
  const long long n = rows * ((dst_cols + 1) / 2);
  KernelSpan span("cast_rows_kernel", stream);
  cast_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(src, rows, (int)cols, ld_src, (bf16*)dst, ld_dst, (int)dst_cols);
  TT_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// embedding gather: one warp per output row
// ---------------------------------------------------------------------------------------------
template <typename OutT>
__global__ void gather_rows_kernel(const float* __restrict__ table, long long table_rows, int dim,
                                   const long long* __restrict__ ids, long long n, OutT* __restrict__ dst,
                                   long long ld_dst, int* oob_flag) {
  const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= n) return;
  long long id = ids[row];
  if (id < 0 || id >= table_rows) {
    if (lane == 0 && oob_flag) atomicExch(oob_flag, 1);
    id = id < 0 ? 0 : table_rows - 1;
  }
  const float* s = table + id * dim;
  OutT* d = dst + row * ld_dst;
  for (int c = lane; c < dim; c += 32) {
    if constexpr (sizeof(OutT) == 2) d[c] = __float2bfloat16(s[c]);
    else d[c] = s[c];
  }
}

int gather_rows_bf16(const float* table, long long table_rows, long long dim, const long long* ids, long long n,
                     void* dst, long long ld_dst, int* oob_flag, cudaStream_t stream) {
  if (n == 0) return 0;
  TT_CHECK(table_rows > 0 && dim > 0 && ld_dst >= dim, "gather_rows_bf16: bad shape");
  KernelSpan span("gather_rows_kernel", stream);
  gather_rows_kernel<bf16><<<(unsigned)((n * 32 + 255) / 256), 256, 0, stream>>>(table, table_rows, (int)dim, ids, n, (bf16*)dst, ld_dst, oob_flag);
  TT_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}
int gather_rows_f32(const float* table, long long table_rows, long long dim, const long long* ids, long long n,
                    float* dst, long long ld_dst, int* oob_flag, cudaStream_t stream) {
  if (n == 0) return 0;
  TT_CHECK(table_rows > 0 && dim > 0 && ld_dst >= dim, "gather_rows_f32: bad shape");
  KernelSpan span("gather_rows_kernel", stream);
  gather_rows_kernel<float><<<(unsigned)((n * 32 + 255) / 256), 256, 0, stream>>>(table, table_rows, (int)dim, ids, n, dst, ld_dst, oob_flag);
  TT_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// dense embedding gradient: table_grad[ids[i], :] += src[i, :]   (duplicates accumulate, like
// embedding_dense_backward in the reference's autograd)
// ---------------------------------------------------------------------------------------------
// 16-byte vector reduction: one L2 atomic transaction for four columns
__device__ __forceinline__ void red_add_v4(float* p, float x, float y, float z, float w) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
}
__device__ __forceinline__ float4 load4_as_f32(const bf16* s16, const float* s32, long long off) {
  if (s16) {
    const uint2 u = *reinterpret_cast<const uint2*>(s16 + off);
    const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x));
    const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
    return make_float4(a.x, a.y, b.x, b.y);
  }
  return *reinterpret_cast<const float4*>(s32 + off);
}
// one warp per source row; VEC: a lane owns 4 consecutive columns (dim % 4 == 0, 16-byte aligned rows on both sides)
template <bool VEC>
__global__ void scatter_add_rows_kernel(const bf16* __restrict__ s16, const float* __restrict__ s32, long long ld_src,
                                        const long long* __restrict__ ids, long long n, int dim,
                                        float* __restrict__ grad, long long table_rows) {
  const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= n) return;
  const long long id = ids[row];
  if (id < 0 || id >= table_rows) return;
  float* g = grad + id * dim;
  if (VEC) {
    for (int c = lane * 4; c < dim; c += 128) {
      const float4 v = load4_as_f32(s16, s32, row * ld_src + c);
      red_add_v4(g + c, v.x, v.y, v.z, v.w);
    }
  } else {
    for (int c = lane; c < dim; c += 32) {
      const float v = s16 ? __bfloat162float(s16[row * ld_src + c]) : s32[row * ld_src + c];
      atomicAdd(g + c, v);
    }
  }
}
static bool rows_vec4(const void* src16, const float* src32, long long ld_src, long long dim, const float* table_grad) {
  const bool src_ok = src16 ? (ld_src % 4 == 0 && (uintptr_t)src16 % 8 == 0) : (ld_src % 4 == 0 && (uintptr_t)src32 % 16 == 0);
  return dim % 4 == 0 && src_ok && (uintptr_t)table_grad % 16 == 0;
}
int scatter_add_rows(const void* src16, const float* src32, long long ld_src, const long long* ids, long long n,
                     long long dim, float* table_grad, long long table_rows, cudaStream_t stream) {
  if (n == 0) return 0;
  TT_CHECK((src16 != nullptr) != (src32 != nullptr), "scatter_add_rows: exactly one source");
  KernelSpan span("scatter_add_rows_kernel", stream);
  if (rows_vec4(src16, src32, ld_src, dim, table_grad))
    scatter_add_rows_kernel<true><<<(unsigned)((n * 32 + 255) / 256), 256, 0, stream>>>((const bf16*)src16, src32, ld_src, ids, n, (int)dim, table_grad, table_rows);
  else
    scatter_add_rows_kernel<false><<<(unsigned)((n * 32 + 255) / 256), 256, 0, stream>>>((const bf16*)src16, src32, ld_src, ids, n, (int)dim, table_grad, table_rows);
  TT_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// column sums (bias gradients): out[c] += sum_r src[r, c];  out must be initialised by the caller
// ---------------------------------------------------------------------------------------------
__global__ void colsum_kernel(const bf16* __restrict__ s16, const float* __restrict__ s32, long long rows, int cols,
                              long long ld, float* __restrict__ out, int rows_per_block) {
  __shared__ float red[8][33];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cx;
  const long long r0 = (long long)blockIdx.y * rows_per_block;
  const long long r1 = r0 + rows_per_block < rows ? r0 + rows_per_block : rows;
  float acc = 0.f;
  if (c < cols)
    for (long long r = r0 + ry; r < r1; r += 8) acc += s16 ? __bfloat162float(s16[r * ld + c]) : s32[r * ld + c];
  red[ry][cx] = acc;
  __syncthreads();
  if (ry == 0 && c < cols) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i][cx];
    atomicAdd(out + c, t);
  }
}
// vectorised variant: a thread owns 8 consecutive columns (one 16-byte bf16 load / two fp32 float4 loads per row),
// a block covers 256 columns x its share of the rows; rows are read as full contiguous segments.
template <typename T>
__global__ void __launch_bounds__(256)
colsum_vec_kernel(const T* __restrict__ src, long long rows, int cols, long long ld, float* __restrict__ out,
                  long long rows_per_block) {
  __shared__ float red[8][32][9];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c0 = (blockIdx.x * 32 + tx) * 8;
  const long long r0 = (long long)blockIdx.y * rows_per_block;
  const long long r1 = r0 + rows_per_block < rows ? r0 + rows_per_block : rows;
  float acc[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = 0.f;
  if (c0 < cols) {
#pragma unroll 4
    for (long long r = r0 + ty; r < r1; r += 8) {
      if (sizeof(T) == 2) {
        const uint4 v = *reinterpret_cast<const uint4*>(src + r * ld + c0);
        const bf16* b = reinterpret_cast<const bf16*>(&v);
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] += __bfloat162float(b[e]);
      } else {
        const float4 v0 = *reinterpret_cast<const float4*>(src + r * ld + c0);
        const float4 v1 = *reinterpret_cast<const float4*>(src + r * ld + c0 + 4);
        acc[0] += v0.x; acc[1] += v0.y; acc[2] += v0.z; acc[3] += v0.w;
        acc[4] += v1.x; acc[5] += v1.y; acc[6] += v1.z; acc[7] += v1.w;
      }
    }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) red[ty][tx][e] = acc[e];
  __syncthreads();
  if (ty == 0 && c0 < cols) {
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float t = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) t += red[i][tx][e];
      atomicAdd(out + c0 + e, t);
    }
  }
}

int colsum(const void* src16, const float* src32, long long rows, long long cols, long long ld, float* out,
           cudaStream_t stream) {
  if (rows == 0 || cols == 0) return 0;
  TT_CHECK((src16 != nullptr) != (src32 != nullptr), "colsum: exactly one source");
  const void* base = src16 ? src16 : (const void*)src32;
  const bool vec = (cols % 8 == 0) && (ld % 8 == 0) && (((uintptr_t)base) % 16 == 0);
  if (vec) {
    const unsigned gx = (unsigned)((cols + 255) / 256);
    long long gy = (4ll * num_sms() + gx - 1) / gx;            // ~4 blocks per SM in total
    long long rpb = (rows + gy - 1) / gy;
    rpb = rpb < 64 ? 64 : (rpb + 7) / 8 * 8;                   // at least 8 rows per thread-row
    gy = (rows + rpb - 1) / rpb;
    dim3 grid(gx, (unsigned)gy);
    KernelSpan span("colsum_kernel", stream);
    if (src16) colsum_vec_kernel<bf16><<<grid, 256, 0, stream>>>((const bf16*)src16, rows, (int)cols, ld, out, rpb);
    else colsum_vec_kernel<float><<<grid, 256, 0, stream>>>(src32, rows, (int)cols, ld, out, rpb);
    TT_CUDA(cudaGetLastError());
    count_launch();
    return 0;
  }
  const int rpb = 256;
  dim3 grid((unsigned)((cols + 31) / 32), (unsigned)((rows + rpb - 1) / rpb));
  KernelSpan span("colsum_kernel", stream);
  colsum_kernel<<<grid, 256, 0, stream>>>((const bf16*)src16, src32, rows, (int)cols, ld, out, rpb);
  TT_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// history: x16[b*H + h, :] = bf16(table[ids[b,h]] + pe[h]);  mean[b, :] = mean_h table[ids[b,h]]
// (mean pooling BEFORE the positional encoding: reference src/user_history_encoder.py:89 vs :95;
//  history ids index the ITEM table: src/two_tower_with_user_history_encoder.py:105)
// ---------------------------------------------------------------------------------------------
__global__ void history_gather_pool_kernel(const float* __restrict__ table, long long table_rows, int D,
                                           const long long* __restrict__ ids, int H, const float* __restrict__ pe,
                                           bf16* __restrict__ x16, long long ldx, float* __restrict__ mean,
                                           long long ldmean, int* oob_flag) {
  const long long b = blockIdx.x;
  const float inv = 1.f / (float)H;
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    float sum = 0.f;
    for (int h = 0; h < H; ++h) {
      long long id = ids[b * H + h];
      if (id < 0 || id >= table_rows) {
        if (oob_flag) atomicExch(oob_flag, 1);
        id = id < 0 ? 0 : table_rows - 1;
      }
      const float v = table[id * D + c];
      sum += v;
      x16[(b * H + h) * ldx + c] = __float2bfloat16(v + (pe ? pe[h * D + c] : 0.f));
    }
    mean[b * ldmean + c] = sum * inv;
  }
}
int history_gather_pool(const float* table, long long table_rows, long long D, const long long* ids, long long B,
                        long long H, const float* pe, void* x16, long long ldx, float* mean, long long ldmean,
                        int* oob_flag, cudaStream_t stream) {
  if (B == 0) return 0;
  TT_CHECK(H > 0 && D > 0 && ldx >= D && ldmean >= D, "history_gather_pool: bad shape");
  const int threads = D >= 256 ? 256 : (D >= 128 ? 128 : 64);
  KernelSpan span("history_gather_pool_kernel", stream);
  history_gather_pool_kernel<<<(unsigned)B, threads, 0, stream>>>(table, table_rows, (int)D, ids, (int)H, pe, (bf16*)x16, ldx, mean, ldmean, oob_flag);
  TT_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

// table_grad[ids[b,h], :] += dx[b*H+h, :] + dmean[b, :] / H
// One warp per batch row; VEC: a lane owns 4 consecutive columns, the H ids of the row are read once per warp.
template <bool VEC>
__global__ void __launch_bounds__(256)
history_scatter_grad_kernel(const bf16* __restrict__ dx16, long long lddx, const float* __restrict__ dmean,
                            long long lddmean, const long long* __restrict__ ids, long long B, int H, int D,
                            float* __restrict__ grad, long long table_rows) {
  const long long b = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (b >= B) return;
  const float inv = 1.f / (float)H;
  if (VEC) {
    for (int c0 = 0; c0 < D; c0 += 128) {  // every lane walks the loop (the ids travel by shuffle); idle lanes skip the work
      const int c = c0 + lane * 4;
      const bool act = c < D;
      float4 dm = make_float4(0.f, 0.f, 0.f, 0.f);
      if (dmean && act) {
        dm = *reinterpret_cast<const float4*>(dmean + b * lddmean + c);
        dm.x *= inv; dm.y *= inv; dm.z *= inv; dm.w *= inv;
      }
      for (int h0 = 0; h0 < H; h0 += 32) {
        const long long my_id = h0 + lane < H ? ids[b * H + h0 + lane] : -1;
        const int hn = H - h0 < 32 ? H - h0 : 32;
        for (int hh = 0; hh < hn; ++hh) {
          const long long id = __shfl_sync(0xffffffffu, my_id, hh);
          if (id < 0 || id >= table_rows || !act) continue;
          float4 v = dm;
          if (dx16) {
            const float4 x = load4_as_f32(dx16, nullptr, (b * H + h0 + hh) * lddx + c);
            v.x += x.x; v.y += x.y; v.z += x.z; v.w += x.w;
          }
          red_add_v4(grad + id * D + c, v.x, v.y, v.z, v.w);
        }
      }
    }
  } else {
    for (int c = lane; c < D; c += 32) {
      const float dm = dmean ? dmean[b * lddmean + c] * inv : 0.f;
      for (int h = 0; h < H; ++h) {
        const long long id = ids[b * H + h];
        if (id < 0 || id >= table_rows) continue;
        const float v = (dx16 ? __bfloat162float(dx16[(b * H + h) * lddx + c]) : 0.f) + dm;
        atomicAdd(grad + id * D + c, v);
      }
    }
  }
}
int history_scatter_grad(const void* dx16, long long lddx, const float* dmean, long long lddmean,
                         const long long* ids, long long B, long long H, long long D, float* table_grad,
                         long long table_rows, cudaStream_t stream) {
  if (B == 0) return 0;
  KernelSpan span("history_scatter_grad_kernel", stream);
  const bool vec = D % 4 == 0 && (uintptr_t)table_grad % 16 == 0 && (!dx16 || (lddx % 4 == 0 && (uintptr_t)dx16 % 8 == 0)) &&
                   (!dmean || (lddmean % 4 == 0 && (uintptr_t)dmean % 16 == 0));
  const unsigned blocks = (unsigned)((B * 32 + 255) / 256);
  if (vec)
    history_scatter_grad_kernel<true><<<blocks, 256, 0, stream>>>((const bf16*)dx16, lddx, dmean, lddmean, ids, B, (int)H, (int)D, table_grad, table_rows);
  else
    history_scatter_grad_kernel<false><<<blocks, 256, 0, stream>>>((const bf16*)dx16, lddx, dmean, lddmean, ids, B, (int)H, (int)D, table_grad, table_rows);
  TT_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// value-weighted mean of the per-row cross entropy (reference src/two_tower_base_retrieval.py:322-343
// with the identity debias hook):  nuv_i = sum_t labels[i,t] w_t;  w_i = max(nuv_i, 1e-6) / max_i(.);
// loss = sum_i ce_i w_i / B;  g_i = d loss / d ce_i = w_i / B.   One CTA, two passes over B rows.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
weighted_loss_kernel(const float* __restrict__ ce, const float* __restrict__ labels, long long ldl,
                     const float* __restrict__ uvw, int B, int T, float inv_rows, float* __restrict__ loss,
                     float* __restrict__ g) {
  __shared__ float red[32];
  __shared__ float bcast;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float mx = 0.f;
  for (int i = tid; i < B; i += blockDim.x) {
    float nuv = 0.f;
    for (int t = 0; t < T; ++t) nuv = fmaf(labels[(long long)i * ldl + t], uvw[t], nuv);
    nuv = fmaxf(nuv, 0.000001f);
    g[i] = nuv;  // parked until the maximum is known
    mx = fmaxf(mx, nuv);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if (lane == 0) red[warp] = mx;
  __syncthreads();
  if (warp == 0) {
    float m = lane < (blockDim.x >> 5) ? red[lane] : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0) bcast = m;
  }
  __syncthreads();
  const float inv_max = 1.f / bcast;
  float acc = 0.f;
  for (int i = tid; i < B; i += blockDim.x) {
    const float w = g[i] * inv_max;
    acc = fmaf(ce[i], w, acc);
    g[i] = w * inv_rows;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  __syncthreads();
  if (lane == 0) red[warp] = acc;
  __syncthreads();
  if (warp == 0) {
    float a = lane < (blockDim.x >> 5) ? red[lane] : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (lane == 0) *loss = a * inv_rows;
  }
}
int weighted_loss(const float* ce, const float* labels, long long ldl, const float* uvw, long long B, long long T,
                  float* loss, float* g, cudaStream_t stream) {
  TT_CHECK(B > 0 && T > 0 && ldl >= T, "weighted_loss: bad shape");
  KernelSpan span("weighted_loss_kernel", stream);
  weighted_loss_kernel<<<1, 1024, 0, stream>>>(ce, labels, ldl, uvw, (int)B, (int)T, 1.f / (float)B, loss, g);
  TT_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// batched variants: several independent cast / gather problems in one launch (both towers, all weights)
// ---------------------------------------------------------------------------------------------
struct CastBatch {
  CastProblem p[16];
  int n;
  int block_start[17];
};
__global__ void cast_rows_batched_kernel(const __grid_constant__ CastBatch b) {
  int p = 0;
  while ((int)blockIdx.x >= b.block_start[p + 1]) ++p;
  const CastProblem& c = b.p[p];
  const int half = (int)((c.dst_cols + 1) / 2);
  const long long idx = (long long)(blockIdx.x - b.block_start[p]) * blockDim.x + threadIdx.x;
  const long long r = idx / half;
  const int cp = (int)(idx % half) * 2;
  if (r >= c.rows) return;
  const float* s = c.src + r * c.ld_src;
  bf16* d = reinterpret_cast<bf16*>(c.dst) + r * c.ld_dst;
  const float a0 = cp < c.cols ? s[cp] : 0.f;
  const float a1 = cp + 1 < c.cols ? s[cp + 1] : 0.f;
  if (cp + 1 < c.dst_cols) *reinterpret_cast<__nv_bfloat162*>(d + cp) = __floats2bfloat162_rn(a0, a1);
  else d[cp] = __float2bfloat16(a0);
}
int cast_rows_bf16_batched(const CastProblem* probs, int n, cudaStream_t stream) {
  TT_CHECK(n >= 1 && n <= 16, "cast_rows_bf16_batched: 1..16 problems");
  CastBatch b;
  b.n = n;
  b.block_start[0] = 0;
  for (int i = 0; i < n; ++i) {
    const CastProblem& c = probs[i];
    TT_CHECK(c.cols <= c.dst_cols && c.dst_cols <= c.ld_dst && (c.ld_dst % 2) == 0 && ((uintptr_t)c.dst % 4) == 0,
             "cast_rows_bf16_batched: bad shape in problem %d", i);
    b.p[i] = c;
    const long long work = c.rows * ((c.dst_cols + 1) / 2);
    b.block_start[i + 1] = b.block_start[i] + (int)((work + 255) / 256);
  }
  for (int i = n; i < 16; ++i) b.block_start[i + 1] = b.block_start[n];
  if (b.block_start[n] == 0) return 0;
  KernelSpan span("cast_rows_batched_kernel", stream);
  cast_rows_batched_kernel<<<b.block_start[n], 256, 0, stream>>>(b);
  TT_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

struct GatherBatch {
  GatherProblem p[8];
  int n;
  int block_start[9];
};
__global__ void gather_rows_batched_kernel(const __grid_constant__ GatherBatch b, int* oob_flag) {
  int p = 0;
  while ((int)blockIdx.x >= b.block_start[p + 1]) ++p;
  const GatherProblem& g = b.p[p];
  const long long w = ((long long)(blockIdx.x - b.block_start[p]) * blockDim.x + threadIdx.x) >> 5;  // one warp per row
  const int lane = threadIdx.x & 31;
  if (w >= g.n) return;
  long long id = g.ids[w];
  if (id < 0 || id >= g.table_rows) {
    if (oob_flag && lane == 0) atomicExch(oob_flag, 1);
    id = id < 0 ? 0 : g.table_rows - 1;
  }
  const float* src = g.table + id * g.dim;
  bf16* dst = reinterpret_cast<bf16*>(g.dst) + w * g.ld_dst;
  for (int c = lane; c < g.dim; c += 32) dst[c] = __float2bfloat16(src[c]);
}
int gather_rows_bf16_batched(const GatherProblem* probs, int n, int* oob_flag, cudaStream_t stream) {
  TT_CHECK(n >= 1 && n <= 8, "gather_rows_bf16_batched: 1..8 problems");
  GatherBatch b;
  b.n = n;
  b.block_start[0] = 0;
  for (int i = 0; i < n; ++i) {
    TT_CHECK(probs[i].table_rows > 0 && probs[i].dim > 0 && probs[i].ld_dst >= probs[i].dim, "gather_rows_bf16_batched: bad shape");
    b.p[i] = probs[i];
    b.block_start[i + 1] = b.block_start[i] + (int)((probs[i].n * 32 + 255) / 256);
  }
  for (int i = n; i < 8; ++i) b.block_start[i + 1] = b.block_start[n];
  if (b.block_start[n] == 0) return 0;
  KernelSpan span("gather_rows_batched_kernel", stream);
  gather_rows_batched_kernel<<<b.block_start[n], 256, 0, stream>>>(b, oob_flag);
  TT_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Fused multi-tensor Adam (the reference trains with torch.optim.Adam on dense gradients, train/train.py:179,123-125):
// one launch updates every parameter tensor of the model, HBM-bound at 28 bytes per element (read p, g, m, v; write
// p, m, v).  Arithmetic follows torch.optim.Adam (amsgrad off): m += (g - m)(1 - b1); v = b2 v + (1 - b2) g^2;
// p -= lr / (1 - b1^t) * m / (sqrt(v) / sqrt(1 - b2^t) + eps), optional L2 weight decay folded into g.
// The step counter lives on the device (CUDA-graph replay): every block reads it, the last block to finish bumps it.
// ---------------------------------------------------------------------------------------------
static constexpr int ADAM_MAXT = 32;
static constexpr int ADAM_CHUNK = 4096;  // elements per block
struct AdamBatch {
  AdamTensor t[ADAM_MAXT];
  int block_start[ADAM_MAXT + 1];
  int n;
};
__global__ void __launch_bounds__(256)
adam_kernel(const __grid_constant__ AdamBatch b, double lr, double beta1d, double beta2d, float eps, float weight_decay,
            long long* step, unsigned int* ticket) {
  __shared__ float consts[2];
  int p = 0;
  while ((int)blockIdx.x >= b.block_start[p + 1]) ++p;
  const AdamTensor& a = b.t[p];
  if (threadIdx.x == 0) {
    // bias corrections in double, as torch computes them on the host: 1 - beta^t cancels badly in fp32 for small t
    const double t = (double)(*step + 1);
    consts[0] = (float)(lr / (1.0 - pow(beta1d, t)));
    consts[1] = (float)sqrt(1.0 - pow(beta2d, t));
  }
  __syncthreads();
  const float step_size = consts[0], bc2_sqrt = consts[1];
  const float beta2 = (float)beta2d, omb1 = (float)(1.0 - beta1d), omb2 = (float)(1.0 - beta2d);
  const long long base = (long long)(blockIdx.x - b.block_start[p]) * ADAM_CHUNK;
  auto update = [&](float& pv, float gv, float& mv, float& vv) {
    if (weight_decay != 0.f) gv = fmaf(weight_decay, pv, gv);
    mv = fmaf(gv - mv, omb1, mv);
    vv = fmaf(vv, beta2, omb2 * gv * gv);
    const float denom = sqrtf(vv) / bc2_sqrt + eps;
    pv = pv - step_size * (mv / denom);
  };
  if (a.vec4) {
#pragma unroll
    for (int i = 0; i < ADAM_CHUNK / (256 * 4); ++i) {
      const long long e = base + ((long long)i * 256 + threadIdx.x) * 4;
      if (e + 3 < a.n) {
        float4 pv = *reinterpret_cast<float4*>(a.p + e);
        const float4 gv = *reinterpret_cast<const float4*>(a.g + e);
        float4 mv = *reinterpret_cast<float4*>(a.m + e), vv = *reinterpret_cast<float4*>(a.v + e);
        update(pv.x, gv.x, mv.x, vv.x); update(pv.y, gv.y, mv.y, vv.y);
        update(pv.z, gv.z, mv.z, vv.z); update(pv.w, gv.w, mv.w, vv.w);
        *reinterpret_cast<float4*>(a.p + e) = pv;
        *reinterpret_cast<float4*>(a.m + e) = mv;
        *reinterpret_cast<float4*>(a.v + e) = vv;
      } else {
        for (long long k = e; k < a.n && k < e + 4; ++k) update(a.p[k], a.g[k], a.m[k], a.v[k]);
      }
    }
  } else {
    for (int i = threadIdx.x; i < ADAM_CHUNK; i += 256) {
      const long long k = base + i;
      if (k < a.n) update(a.p[k], a.g[k], a.m[k], a.v[k]);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(ticket, 1u) == gridDim.x - 1) {  // every block has read *step by now
      *step += 1;
      *ticket = 0;
    }
  }
}
int adam_step(const AdamTensor* tensors, int n, double lr, double beta1, double beta2, float eps, float weight_decay,
              long long* step, unsigned int* ticket, cudaStream_t stream) {
  TT_CHECK(n >= 1 && step != nullptr && ticket != nullptr, "adam_step: bad arguments");
  for (int i0 = 0; i0 < n; i0 += ADAM_MAXT) {
    TT_CHECK(i0 == 0, "adam_step: at most %d tensors per call (the step counter advances once per launch)", ADAM_MAXT);
    AdamBatch b;
    b.n = n - i0 < ADAM_MAXT ? n - i0 : ADAM_MAXT;
    b.block_start[0] = 0;
    for (int i = 0; i < b.n; ++i) {
      AdamTensor a = tensors[i0 + i];
      TT_CHECK(a.n >= 0 && (a.n == 0 || (a.p && a.g && a.m && a.v)), "adam_step: null tensor %d", i0 + i);
      a.vec4 = (((uintptr_t)a.p | (uintptr_t)a.g | (uintptr_t)a.m | (uintptr_t)a.v) % 16) == 0;
      b.t[i] = a;
      b.block_start[i + 1] = b.block_start[i] + (int)((a.n + ADAM_CHUNK - 1) / ADAM_CHUNK);
    }
    for (int i = b.n; i < ADAM_MAXT; ++i) b.block_start[i + 1] = b.block_start[b.n];
    if (b.block_start[b.n] == 0) return 0;
    KernelSpan span("adam_kernel", stream);
    adam_kernel<<<b.block_start[b.n], 256, 0, stream>>>(b, (double)lr, (double)beta1, (double)beta2, eps, weight_decay, step, ticket);
    TT_CUDA(cudaGetLastError());
    count_launch();
  }
  return 0;
}

}  // namespace tt
