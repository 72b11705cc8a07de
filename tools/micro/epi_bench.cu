// Micro-benchmark of the CE forward epilogue math in isolation: 8 warps per SM (2 per scheduler), each thread owns 64
// "scores" and repeats the online (max, sum-exp) update; reports cycles per tile-equivalent for several formulations.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o epi_bench epi_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
#define LOG2E 1.4426950408889634f
__device__ __forceinline__ float ex2f(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

template <int MODE>
__global__ void __launch_bounds__(256, 1) k(const float* in, float* out, long long* cyc, int iters) {
  float x[64];
  const float base = in[threadIdx.x];
  float m = -1e30f, s = 0.f;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 64; ++i) x[i] = fmaf(base, 0.001f * (float)(i + 1), (float)it * 1e-4f - 0.05f * (float)i);
    if (MODE == 0) {  // as in ce_fwd_kernel: max tree, rescale, exp in place, 4 sum chains
      float c[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) c[u] = fmaxf(x[u], x[u + 4]);
#pragma unroll
      for (int i = 8; i < 64; i += 4)
#pragma unroll
        for (int u = 0; u < 4; ++u) c[u] = fmaxf(c[u], x[i + u]);
      const float cm = fmaxf(fmaxf(c[0], c[1]), fmaxf(c[2], c[3]));
      const float m_new = fmaxf(m, cm * LOG2E);
      s *= ex2f(m - m_new);
#pragma unroll
      for (int i = 0; i < 64; ++i) x[i] = ex2f(fmaf(x[i], LOG2E, -m_new));
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int i = 0; i < 64; i += 4)
#pragma unroll
        for (int u = 0; u < 4; ++u) acc[u] += x[i + u];
      s += (acc[0] + acc[1]) + (acc[2] + acc[3]);
      m = m_new;
    } else if (MODE == 1) {  // no exp (FFMA only)
      float c[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) c[u] = fmaxf(x[u], x[u + 4]);
#pragma unroll
      for (int i = 8; i < 64; i += 4)
#pragma unroll
        for (int u = 0; u < 4; ++u) c[u] = fmaxf(c[u], x[i + u]);
      const float cm = fmaxf(fmaxf(c[0], c[1]), fmaxf(c[2], c[3]));
      const float m_new = fmaxf(m, cm * LOG2E);
#pragma unroll
      for (int i = 0; i < 64; ++i) x[i] = fmaf(x[i], LOG2E, -m_new);
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int i = 0; i < 64; i += 4)
#pragma unroll
        for (int u = 0; u < 4; ++u) acc[u] += x[i + u];
      s += (acc[0] + acc[1]) + (acc[2] + acc[3]);
      m = m_new;
    } else if (MODE == 2) {  // exp only, tree sum afterwards (pairwise)
#pragma unroll
      for (int i = 0; i < 64; ++i) x[i] = ex2f(x[i]);
#pragma unroll
      for (int w = 32; w >= 1; w >>= 1)
#pragma unroll
        for (int i = 0; i < w; ++i) x[i] += x[i + w];
      s += x[0];
    } else if (MODE == 3) {  // 48 MUFU + 16 polynomial exps, rest as MODE 0
      float c[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) c[u] = fmaxf(x[u], x[u + 4]);
#pragma unroll
      for (int i = 8; i < 64; i += 4)
#pragma unroll
        for (int u = 0; u < 4; ++u) c[u] = fmaxf(c[u], x[i + u]);
      const float cm = fmaxf(fmaxf(c[0], c[1]), fmaxf(c[2], c[3]));
      const float m_new = fmaxf(m, cm * LOG2E);
      s *= ex2f(m - m_new);
#pragma unroll
      for (int i = 0; i < 64; ++i) {
        const float v = fmaf(x[i], LOG2E, -m_new);
        if ((i & 3) != 3) {
          x[i] = ex2f(v);
        } else {
          const float vc = fmaxf(v, -120.f);
          const float r = vc + 12582912.f;
          const float f = vc - (r - 12582912.f);
          float p = fmaf(f, 0.0555041f, 0.2402265f);
          p = fmaf(p, f, 0.6931472f);
          p = fmaf(p, f, 1.0f);
          x[i] = __int_as_float(__float_as_int(p) + (__float_as_int(r) << 23));
        }
      }
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int i = 0; i < 64; i += 4)
#pragma unroll
        for (int u = 0; u < 4; ++u) acc[u] += x[i + u];
      s += (acc[0] + acc[1]) + (acc[2] + acc[3]);
      m = m_new;
    } else if (MODE == 4) {  // packed fp32x2 scale and sum (fma.rn.f32x2 / add.rn.f32x2), max tree as before
      float c[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) c[u] = fmaxf(x[u], x[u + 4]);
#pragma unroll
      for (int i = 8; i < 64; i += 4)
#pragma unroll
        for (int u = 0; u < 4; ++u) c[u] = fmaxf(c[u], x[i + u]);
      const float cm = fmaxf(fmaxf(c[0], c[1]), fmaxf(c[2], c[3]));
      const float m_new = fmaxf(m, cm * LOG2E);
      s *= ex2f(m - m_new);
      unsigned long long sc, mn, acc0 = 0ull, acc1 = 0ull;
      const float nm = -m_new;
      asm("mov.b64 %0, {%1, %1};" : "=l"(sc) : "f"(LOG2E));
      asm("mov.b64 %0, {%1, %1};" : "=l"(mn) : "f"(nm));
#pragma unroll
      for (int i = 0; i < 64; i += 2) {
        unsigned long long v;
        asm("mov.b64 %0, {%1, %2};" : "=l"(v) : "f"(x[i]), "f"(x[i + 1]));
        asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(v) : "l"(v), "l"(sc), "l"(mn));
        asm("mov.b64 {%0, %1}, %2;" : "=f"(x[i]), "=f"(x[i + 1]) : "l"(v));
      }
#pragma unroll
      for (int i = 0; i < 64; ++i) x[i] = ex2f(x[i]);
#pragma unroll
      for (int i = 0; i < 64; i += 4) {
        unsigned long long v0, v1;
        asm("mov.b64 %0, {%1, %2};" : "=l"(v0) : "f"(x[i]), "f"(x[i + 1]));
        asm("mov.b64 %0, {%1, %2};" : "=l"(v1) : "f"(x[i + 2]), "f"(x[i + 3]));
        asm("add.rn.f32x2 %0, %0, %1;" : "+l"(acc0) : "l"(v0));
        asm("add.rn.f32x2 %0, %0, %1;" : "+l"(acc1) : "l"(v1));
      }
      float a0, a1, a2, a3;
      asm("mov.b64 {%0, %1}, %2;" : "=f"(a0), "=f"(a1) : "l"(acc0));
      asm("mov.b64 {%0, %1}, %2;" : "=f"(a2), "=f"(a3) : "l"(acc1));
      s += (a0 + a1) + (a2 + a3);
      m = m_new;
    }
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + m;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

template <int MODE>
void run(const char* name) {
  float *in, *out; long long* cyc;
  cudaMalloc(&in, 1024 * 4); cudaMalloc(&out, 148 * 256 * 4); cudaMalloc(&cyc, 8);
  cudaMemset(in, 0x3c, 1024 * 4);
  const int iters = 2000;
  k<MODE><<<148, 256>>>(in, out, cyc, iters);
  cudaDeviceSynchronize();
  k<MODE><<<148, 256>>>(in, out, cyc, iters);
  cudaDeviceSynchronize();
  long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%-44s %8.1f cycles per 128x128 tile-equivalent (8 warps, 64 elements per thread)\n", name, (double)h / iters);
}
int main() {
  run<0>("max + rescale + 64 exp + sum (kernel form)");
  run<1>("same without the exponentials");
  run<2>("64 exp + pairwise tree sum only");
  run<3>("48 MUFU + 16 polynomial exp2");
  run<4>("packed f32x2 scale + sum");
  return 0;
}
