// Last layer of the history encoder, evaluated for query row 0 only - WITHOUT projecting keys and values.
//
// The reference consumes a single row of the last nn.MultiheadAttention layer (src/user_history_encoder.py:116: row 0
// of the sequence).  For one query row the packed in-projection of all H rows is wasted work and traffic:
//     s_j = c q0_h . (Wk_h x_j + bk_h) = (c Wk_h^T q0_h) . x_j + const        (the constant cancels in the softmax)
//     o_h = sum_j p_j (Wv_h x_j + bv_h) = Wv_h (sum_j p_j x_j) + bv_h
// so the attention runs against the RAW layer input x with a per-head transformed query  qt_h = c Wk_h^T q0_h  (D wide),
// and the value projection is applied to the p-weighted mean  z_h = sum_j p_j x_j  afterwards.  The dense pieces
// (q0, qt, o, the out-projection and their gradients) are [B, .]-sized GEMMs on the tcgen05 GEMM kernel; the kernels here
// are the memory-bound rest: one pass over x forward (105 MB at BASELINE configs[2] instead of 0.7 GB through qkv), one
// pass over x plus one write of dx backward.  One warp owns a sequence: lane l holds columns [l*DL, l*DL+DL) of every row.
//
//   forward : p = softmax_j(qt_h . x_j),  z_h = sum_j p_j x_j
//   backward: dp_j = dz_h . x_j,  ds = p (dp - sum p dp),  dqt_h = sum_j ds_j x_j            (pass 1, reads x)
//             dx_j = sum_h (p_j dz_h + ds_j qt_h)  (+ dq0 Wq on row 0),  column sums of dx   (pass 2, writes dx)
#include "common.cuh"
#include "kernels.h"

namespace tt {

namespace {

constexpr int MAXH = 8;  // heads

template <int DL>
__device__ __forceinline__ void load_slice(const bf16* p, float* v) {
  if (DL == 4) {
    const uint2 u = *reinterpret_cast<const uint2*>(p);
    const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x));
    const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
  } else {
    const uint32_t u = *reinterpret_cast<const uint32_t*>(p);
    const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u));
    v[0] = a.x; v[1] = a.y;
  }
}
template <int DL>
__device__ __forceinline__ void store_slice(bf16* p, const float* v) {
  if (DL == 4) *reinterpret_cast<uint2*>(p) = make_uint2(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]));
  else *reinterpret_cast<uint32_t*>(p) = pack_bf16x2(v[0], v[1]);
}
template <int DL>
__device__ __forceinline__ void load_f32(const float* p, float* v) {
  if (DL == 4) {
    const float4 u = *reinterpret_cast<const float4*>(p);
    v[0] = u.x; v[1] = u.y; v[2] = u.z; v[3] = u.w;
  } else {
    const float2 u = *reinterpret_cast<const float2*>(p);
    v[0] = u.x; v[1] = u.y;
  }
}

// Stage the H x D tile of sequence b into the warp's shared-memory slice (rows of D bf16, same layout as in global
// memory).  Contiguous rows (ldx == D): 16-byte fully coalesced loads, eight in flight per lane - a row-by-row loop
// serialised one global round trip per row (50 x ~600 cycles per sequence, 4x the kernel's instruction time).
template <int DL>
__device__ __forceinline__ void stage_tile(const bf16* x, long long ldx, long long row0, int H, bf16* tile, int lane) {
  constexpr int D = DL * 32;
  if (ldx == D) {
    const uint4* src = reinterpret_cast<const uint4*>(x + row0 * ldx);
    uint4* dst = reinterpret_cast<uint4*>(tile);
    const int n16 = H * D / 8;  // 16-byte chunks of the tile
    int i = lane;
    for (; i + 7 * 32 < n16; i += 8 * 32) {
      uint4 r[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) r[u] = src[i + u * 32];
#pragma unroll
      for (int u = 0; u < 8; ++u) dst[i + u * 32] = r[u];
    }
    for (; i < n16; i += 32) dst[i] = src[i];
  } else {
    for (int j = 0; j < H; ++j) {
      if (DL == 4) *reinterpret_cast<uint2*>(tile + j * D + lane * DL) = *reinterpret_cast<const uint2*>(x + (row0 + j) * ldx + lane * DL);
      else *reinterpret_cast<uint32_t*>(tile + j * D + lane * DL) = *reinterpret_cast<const uint32_t*>(x + (row0 + j) * ldx + lane * DL);
    }
  }
  __syncwarp();
}

// dots[c] (lane = row c*32 + lane) = q . x_row for up to 4 chunks of 32 rows: 32 lane-partial dot products are folded
// with a butterfly transpose-reduce (31 shuffles per 32 rows instead of 5 per row)
template <int DL>
__device__ __forceinline__ void row_dots(const bf16* tile, int H, const float* q, int lane, float* dots) {
  constexpr int D = DL * 32;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    dots[c] = 0.f;
    if (c * 32 >= H) continue;
    float v[32];
#pragma unroll
    for (int jj = 0; jj < 32; ++jj) {
      const int j = c * 32 + jj;
      float acc = 0.f;
      if (j < H) {
        float xr[DL];
        load_slice<DL>(tile + j * D + lane * DL, xr);
#pragma unroll
        for (int i = 0; i < DL; ++i) acc = fmaf(q[i], xr[i], acc);
      }
      v[jj] = acc;
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
      const bool up = (lane & off) != 0;
#pragma unroll
      for (int i = 0; i < off; ++i) {
        const float send = up ? v[i] : v[i + off];
        const float keep = up ? v[i + off] : v[i];
        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
      }
    }
    dots[c] = v[0];
  }
}

// out[DL] = sum_j w_j x_j (w_j lives in lane j % 32 of chunk j / 32)
template <int DL>
__device__ __forceinline__ void weighted_rows(const bf16* tile, int H, const float* w, int lane, float* out) {
  constexpr int D = DL * 32;
#pragma unroll
  for (int i = 0; i < DL; ++i) out[i] = 0.f;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    if (c * 32 >= H) continue;
    const int n = H - c * 32 < 32 ? H - c * 32 : 32;
    for (int jj = 0; jj < n; ++jj) {
      const float wj = __shfl_sync(0xffffffffu, w[c], jj);
      float xr[DL];
      load_slice<DL>(tile + (c * 32 + jj) * D + lane * DL, xr);
#pragma unroll
      for (int i = 0; i < DL; ++i) out[i] = fmaf(wj, xr[i], out[i]);
    }
  }
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <int DL>
__global__ void __launch_bounds__(256)
hist_last_fwd_kernel(const bf16* __restrict__ x, long long ldx, const float* __restrict__ qt, int B, int H, int heads,
                     bf16* __restrict__ z16, float* __restrict__ p32) {
  constexpr int D = DL * 32;
  extern __shared__ uint8_t hl_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  bf16* tile = reinterpret_cast<bf16*>(hl_smem) + (size_t)warp * H * D;
  for (int b = blockIdx.x * wpb + warp; b < B; b += gridDim.x * wpb) {
    stage_tile<DL>(x, ldx, (long long)b * H, H, tile, lane);
    for (int h = 0; h < heads; ++h) {
      float q[DL];
      load_f32<DL>(qt + ((long long)b * heads + h) * D + lane * DL, q);
      float s[4];
      row_dots<DL>(tile, H, q, lane, s);
      float m = -INFINITY;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        if (c * 32 + lane >= H) s[c] = -INFINITY;
        m = fmaxf(m, s[c]);
      }
      m = warp_max(m);
      float sum = 0.f;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        s[c] = (c * 32 + lane < H) ? __expf(s[c] - m) : 0.f;
        sum += s[c];
      }
      const float inv = 1.f / warp_sum(sum);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        s[c] *= inv;
        if (c * 32 + lane < H) p32[((long long)b * heads + h) * H + c * 32 + lane] = s[c];
      }
      float z[DL];
      weighted_rows<DL>(tile, H, s, lane, z);
      store_slice<DL>(z16 + ((long long)b * heads + h) * D + lane * DL, z);
    }
    __syncwarp();
  }
}

template <int DL>
__global__ void __launch_bounds__(256)
hist_last_bwd1_kernel(const bf16* __restrict__ x, long long ldx, const float* __restrict__ dz, const float* __restrict__ p32,
                      int B, int H, int heads, float* __restrict__ ds32, bf16* __restrict__ dqt16) {
  constexpr int D = DL * 32;
  extern __shared__ uint8_t hl_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  bf16* tile = reinterpret_cast<bf16*>(hl_smem) + (size_t)warp * H * D;
  for (int b = blockIdx.x * wpb + warp; b < B; b += gridDim.x * wpb) {
    stage_tile<DL>(x, ldx, (long long)b * H, H, tile, lane);
    for (int h = 0; h < heads; ++h) {
      float g[DL];
      load_f32<DL>(dz + ((long long)b * heads + h) * D + lane * DL, g);
      float dp[4], p[4];
      row_dots<DL>(tile, H, g, lane, dp);
      float t = 0.f;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        p[c] = (c * 32 + lane < H) ? p32[((long long)b * heads + h) * H + c * 32 + lane] : 0.f;
        t = fmaf(p[c], dp[c], t);
      }
      t = warp_sum(t);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        dp[c] = p[c] * (dp[c] - t);  // ds
        if (c * 32 + lane < H) ds32[((long long)b * heads + h) * H + c * 32 + lane] = dp[c];
      }
      float dq[DL];
      weighted_rows<DL>(tile, H, dp, lane, dq);
      store_slice<DL>(dqt16 + ((long long)b * heads + h) * D + lane * DL, dq);
    }
    __syncwarp();
  }
}

template <int DL>
__global__ void __launch_bounds__(256)
hist_last_bwd2_kernel(const float* __restrict__ dz, const float* __restrict__ qt, const float* __restrict__ p32,
                      const float* __restrict__ ds32, const float* __restrict__ extra, int B, int H, int heads,
                      bf16* __restrict__ dx16, long long lddx, float* __restrict__ colsum) {
  constexpr int D = DL * 32;
  extern __shared__ uint8_t hl_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  float* wsm = reinterpret_cast<float*>(hl_smem) + (size_t)warp * 2 * heads * H;  // p | ds of the sequence
  float cs[DL];
#pragma unroll
  for (int i = 0; i < DL; ++i) cs[i] = 0.f;
  for (int b = blockIdx.x * wpb + warp; b < B; b += gridDim.x * wpb) {
    const int n = heads * H;
    {  // p | ds of the sequence -> shared memory, all loads of a lane in flight together
      float tp[8], td[8];
      for (int i0 = 0; i0 < n; i0 += 8 * 32) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int i = i0 + u * 32 + lane;
          tp[u] = i < n ? p32[(long long)b * n + i] : 0.f;
          td[u] = i < n ? ds32[(long long)b * n + i] : 0.f;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int i = i0 + u * 32 + lane;
          if (i < n) { wsm[i] = tp[u]; wsm[n + i] = td[u]; }
        }
      }
    }
    float g[MAXH][DL], q[MAXH][DL];
#pragma unroll
    for (int h = 0; h < MAXH; ++h) {
      if (h < heads) {
        load_f32<DL>(dz + ((long long)b * heads + h) * D + lane * DL, g[h]);
        load_f32<DL>(qt + ((long long)b * heads + h) * D + lane * DL, q[h]);
      }
    }
    __syncwarp();
    for (int j = 0; j < H; ++j) {
      float acc[DL];
#pragma unroll
      for (int i = 0; i < DL; ++i) acc[i] = 0.f;
      if (j == 0 && extra != nullptr) load_f32<DL>(extra + (long long)b * D + lane * DL, acc);
#pragma unroll
      for (int h = 0; h < MAXH; ++h) {
        if (h < heads) {
          const float pj = wsm[h * H + j], dsj = wsm[n + h * H + j];
#pragma unroll
          for (int i = 0; i < DL; ++i) acc[i] = fmaf(pj, g[h][i], fmaf(dsj, q[h][i], acc[i]));
        }
      }
      store_slice<DL>(dx16 + ((long long)b * H + j) * lddx + lane * DL, acc);
#pragma unroll
      for (int i = 0; i < DL; ++i) cs[i] += acc[i];
    }
    __syncwarp();
  }
  if (colsum != nullptr) {  // fold the warps of the block, then one atomic per column per block
    __shared__ float red[8][128];
    __syncthreads();
#pragma unroll
    for (int i = 0; i < DL; ++i) red[warp][lane * DL + i] = cs[i];
    __syncthreads();
    if ((int)threadIdx.x < D) {
      float t = 0.f;
      for (int w = 0; w < wpb; ++w) t += red[w][threadIdx.x];
      atomicAdd(colsum + threadIdx.x, t);
    }
  }
}

int check_shape(long long B, long long H, long long D, long long heads, const char* what) {
  TT_CHECK(B > 0 && H > 0 && H <= 128 && (D == 64 || D == 128) && heads >= 1 && heads <= MAXH && D % heads == 0,
           "%s: unsupported shape B=%lld H=%lld D=%lld heads=%lld (D in {64, 128}, H <= 128, heads <= 8)", what, B, H, D, heads);
  return 0;
}
int pick_wpb(size_t per_warp) {  // warps per block: as many as fit into ~96 KB of shared memory, at most 8
  int w = (int)((96 * 1024) / (per_warp ? per_warp : 1));
  return w > 8 ? 8 : (w < 1 ? 1 : w);
}

}  // namespace

int history_last_supported(long long H, long long D, long long heads) {
  return (H > 0 && H <= 128 && (D == 64 || D == 128) && heads >= 1 && heads <= MAXH && D % heads == 0 && (D / heads) % 8 == 0) ? 1 : 0;
}

int history_last_fwd(const void* x16, long long ldx, const float* qt, long long B, long long H, long long D, long long heads,
                     void* z16, float* p32, cudaStream_t stream) {
  if (check_shape(B, H, D, heads, "history_last_fwd")) return -1;
  TT_CHECK((ldx % 4) == 0, "history_last_fwd: row pitch must be a multiple of 4 elements");
  const size_t per_warp = (size_t)H * D * 2;
  const int wpb = pick_wpb(per_warp);
  const size_t smem = per_warp * wpb;
  const int blocks = (int)((B + wpb - 1) / wpb < 4LL * num_sms() ? (B + wpb - 1) / wpb : 4LL * num_sms());
  KernelSpan span("hist_last_fwd_kernel", stream);
  if (D == 128) {
    TT_CUDA(cudaFuncSetAttribute(hist_last_fwd_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    hist_last_fwd_kernel<4><<<blocks, wpb * 32, smem, stream>>>((const bf16*)x16, ldx, qt, (int)B, (int)H, (int)heads, (bf16*)z16, p32);
  } else {
    TT_CUDA(cudaFuncSetAttribute(hist_last_fwd_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    hist_last_fwd_kernel<2><<<blocks, wpb * 32, smem, stream>>>((const bf16*)x16, ldx, qt, (int)B, (int)H, (int)heads, (bf16*)z16, p32);
  }
  TT_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

int history_last_bwd1(const void* x16, long long ldx, const float* dz, const float* p32, long long B, long long H, long long D,
                      long long heads, float* ds32, void* dqt16, cudaStream_t stream) {
  if (check_shape(B, H, D, heads, "history_last_bwd1")) return -1;
  const size_t per_warp = (size_t)H * D * 2;
  const int wpb = pick_wpb(per_warp);
  const size_t smem = per_warp * wpb;
  const int blocks = (int)((B + wpb - 1) / wpb < 4LL * num_sms() ? (B + wpb - 1) / wpb : 4LL * num_sms());
  KernelSpan span("hist_last_bwd1_kernel", stream);
  if (D == 128) {
    TT_CUDA(cudaFuncSetAttribute(hist_last_bwd1_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    hist_last_bwd1_kernel<4><<<blocks, wpb * 32, smem, stream>>>((const bf16*)x16, ldx, dz, p32, (int)B, (int)H, (int)heads, ds32, (bf16*)dqt16);
  } else {
    TT_CUDA(cudaFuncSetAttribute(hist_last_bwd1_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    hist_last_bwd1_kernel<2><<<blocks, wpb * 32, smem, stream>>>((const bf16*)x16, ldx, dz, p32, (int)B, (int)H, (int)heads, ds32, (bf16*)dqt16);
  }
  TT_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

int history_last_bwd2(const float* dz, const float* qt, const float* p32, const float* ds32, const float* extra, long long B,
                      long long H, long long D, long long heads, void* dx16, long long lddx, float* colsum, cudaStream_t stream) {
  if (check_shape(B, H, D, heads, "history_last_bwd2")) return -1;
  TT_CHECK((lddx % 4) == 0, "history_last_bwd2: row pitch must be a multiple of 4 elements");
  const size_t per_warp = (size_t)2 * heads * H * 4;
  const int wpb = 8;
  const size_t smem = per_warp * wpb;
  const int blocks = (int)((B + wpb - 1) / wpb < 4LL * num_sms() ? (B + wpb - 1) / wpb : 4LL * num_sms());
  KernelSpan span("hist_last_bwd2_kernel", stream);
  if (smem > 48 * 1024) {
    TT_CUDA(cudaFuncSetAttribute(hist_last_bwd2_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    TT_CUDA(cudaFuncSetAttribute(hist_last_bwd2_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  if (D == 128)
    hist_last_bwd2_kernel<4><<<blocks, wpb * 32, smem, stream>>>(dz, qt, p32, ds32, extra, (int)B, (int)H, (int)heads, (bf16*)dx16, lddx, colsum);
  else
    hist_last_bwd2_kernel<2><<<blocks, wpb * 32, smem, stream>>>(dz, qt, p32, ds32, extra, (int)B, (int)H, (int)heads, (bf16*)dx16, lddx, colsum);
  TT_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

}  // namespace tt
