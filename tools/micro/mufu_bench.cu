// Micro-benchmark: per-SM throughput of ex2 in fp32 / f16x2 / bf16x2 form and of an FMA-pipe polynomial exp2.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu_bench mufu_bench.cu
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k(float* out, int iters, float seed) {
  float x[8];
  for (int i = 0; i < 8; ++i) x[i] = seed * (threadIdx.x + i) * 1e-3f;
  unsigned h[8];
  for (int i = 0; i < 8; ++i) h[i] = 0x3c003c00u ^ (threadIdx.x + i);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) {
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
      } else if (MODE == 1) {
        asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h[i]));
      } else if (MODE == 2) {
        asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(h[i]));
      } else if (MODE == 3) {  // Cody-Waite + degree-3 polynomial on the FMA pipe
        float v = x[i];
        float r = v + 12582912.f;           // round to integer in the mantissa
        float n = r - 12582912.f;
        float f = v - n;                     // [-0.5, 0.5]
        float p = fmaf(f, 0.0555041f, 0.2402265f);
        p = fmaf(p, f, 0.6931472f);
        p = fmaf(p, f, 1.0f);
        int bits = __float_as_int(p) + (__float_as_int(r) << 23);
        x[i] = __int_as_float(bits) * 1e-9f;
      } else if (MODE == 4) {  // mixed: 6 MUFU + 2 polynomial per 8
        if (i < 6) {
          asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
        } else {
          float v = x[i];
          float r = v + 12582912.f;
          float n = r - 12582912.f;
          float f = v - n;
          float p = fmaf(f, 0.0555041f, 0.2402265f);
          p = fmaf(p, f, 0.6931472f);
          p = fmaf(p, f, 1.0f);
          int bits = __float_as_int(p) + (__float_as_int(r) << 23);
          x[i] = __int_as_float(bits) * 1e-9f;
        }
      }
    }
  }
  float s = 0.f;
  for (int i = 0; i < 8; ++i) s += x[i] + __uint_as_float(h[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, int per_lane, int warps) {
  float* out;
  cudaMalloc(&out, 148 * 1024 * 4);
  const int iters = 20000;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<148, warps * 32>>>(out, 100, 0.5f);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  k<MODE><<<148, warps * 32>>>(out, iters, 0.5f);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  double ops = 148.0 * warps * 32 * 8.0 * iters * per_lane;
  printf("%-28s warps/SM=%2d: %.3f ms  %.2f Texp/s  (%.1f exp/ns/SM)\n", name, warps, ms, ops / ms / 1e9, ops / ms / 1e6 / 148);
  cudaFree(out);
}

int main() {
  for (int w : {8, 16, 32}) {
    run<0>("ex2.approx.ftz.f32", 1, w);
    run<1>("ex2.approx.f16x2", 2, w);
    run<2>("ex2.approx.ftz.bf16x2", 2, w);
    run<3>("poly3 exp2 (FMA pipe)", 1, w);
    run<4>("6 MUFU + 2 poly per 8", 1, w);
  }
  return 0;
}
