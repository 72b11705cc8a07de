// Issue-to-completion rate of tcgen05.mma (kind::f16, bf16 -> fp32, M = 128, K = 16, cta_group::1) by operand source and N.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I two_tower_models_b200/csrc tools/micro/umma_rate.cu -o gpurun_out/umma_rate
// One CTA per SM (148 CTAs, like the real kernels), one thread issues REPS back-to-back instructions into the same accumulator,
// commits, waits; cycles = clock64 around the whole batch / REPS.  Operand contents are zeros (rates do not depend on data).
#include <cstdio>
#include <cstdlib>
#include "common.cuh"
namespace tt { void set_error(const char*, ...) {} }
using namespace tt;

struct Case { int N; int a_tmem; int b_mn; int reps; int mode; int bgwarps; };

template <int A_TMEM, int B_MN>
__global__ void __launch_bounds__(640, 1) rate_kernel(Case c, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar, bar2, bar3;
  __shared__ uint32_t holder;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __shared__ volatile int done_flag;
  if (threadIdx.x == 0) done_flag = 0;
  for (int i = threadIdx.x; i < (128 * 1024) / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async_smem();
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_init(&bar2, 1); mbar_init(&bar3, 1); mbar_arrive(&bar3); fence_barrier_init(); }
  if (warp == 1) tmem_alloc(&holder, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = holder;
  if (warp == 0) {
    const uint32_t leader = elect_one();
    const uint32_t idesc = make_idesc_bf16(128, c.N, 0, B_MN);
    const uint64_t da = make_smem_desc_sw128(smem_u32(smem), 0, 1024);                       // A: 128 rows, K-major
    const uint64_t db = B_MN ? make_smem_desc_sw128(smem_u32(smem + 32768), 256 * 128, 1024)  // B MN-major: N atoms of 64
                             : make_smem_desc_sw128(smem_u32(smem + 32768), 0, 1024);          // B K-major: N rows
    long long t0 = clock64();
    if (c.reps == 0) { while (clock64() - t0 < 16000) { } }
    for (int it = 0; it < c.reps / 8; ++it) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {  // fully unrolled batch: descriptors are compile-time offsets (the loop is not issue-bound)
        if (A_TMEM) umma_bf16_ta_w(tm, tm + 256 + i * 8, desc_advance(db, B_MN ? (i & 3) * 2048 : (i & 3) * 32), idesc, 1u, leader);
        else umma_bf16_w(tm, desc_advance(da, (i & 3) * 32), desc_advance(db, B_MN ? (i & 3) * 2048 : (i & 3) * 32), idesc, 1u, leader);
      }
      // what the issuing warp of the real kernels does between two batches
      if (c.mode == 1) tc_fence_after();
      else if (c.mode == 2) umma_commit_w(&bar2, leader);
      else if (c.mode == 3) { if (!mbar_probe(&bar3, 0)) __trap(); }
      else if (c.mode == 4) mbar_wait(&bar3, 0);
      else if (c.mode == 5) { mbar_wait(&bar3, 0); tc_fence_after(); umma_commit_w(&bar2, leader); }
    }
    long long t1 = clock64();
    umma_commit_w(&bar, leader);
    mbar_wait(&bar, 0);
    long long t2 = clock64();
    if (lane == 0 && blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    done_flag = 1;
  } else if (warp >= 4 && c.mode >= 6) {
    // background TMEM traffic like an epilogue: mode 6 = tcgen05.ld x32 loops, 7 = ld + st, 8 = ld + 32 MUFU + st, from `c.mode2` warps
    const int q = warp & 3;
    if (warp - 4 < c.bgwarps) {
      const uint32_t addr = tm + ((uint32_t)(q * 32) << 16) + 320 + ((warp - 4) >> 2) * 32;
      float v[32];
      float acc = 0.f;
      long long n = 0;
      while (!done_flag) {
        if (c.mode == 9) {  // one 32x32b.x64 load (8 KB per warp) instead of x32
          float w[64];
          uint32_t* r = reinterpret_cast<uint32_t*>(w);
          asm volatile("tcgen05.ld.sync.aligned.32x32b.x64.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32,%33,%34,%35,%36,%37,%38,%39,%40,%41,%42,%43,%44,%45,%46,%47,%48,%49,%50,%51,%52,%53,%54,%55,%56,%57,%58,%59,%60,%61,%62,%63}, [%64];"
            : "=r"(r[0]),"=r"(r[1]),"=r"(r[2]),"=r"(r[3]),"=r"(r[4]),"=r"(r[5]),"=r"(r[6]),"=r"(r[7]),"=r"(r[8]),"=r"(r[9]),"=r"(r[10]),"=r"(r[11]),"=r"(r[12]),"=r"(r[13]),"=r"(r[14]),"=r"(r[15]),
              "=r"(r[16]),"=r"(r[17]),"=r"(r[18]),"=r"(r[19]),"=r"(r[20]),"=r"(r[21]),"=r"(r[22]),"=r"(r[23]),"=r"(r[24]),"=r"(r[25]),"=r"(r[26]),"=r"(r[27]),"=r"(r[28]),"=r"(r[29]),"=r"(r[30]),"=r"(r[31]),
              "=r"(r[32]),"=r"(r[33]),"=r"(r[34]),"=r"(r[35]),"=r"(r[36]),"=r"(r[37]),"=r"(r[38]),"=r"(r[39]),"=r"(r[40]),"=r"(r[41]),"=r"(r[42]),"=r"(r[43]),"=r"(r[44]),"=r"(r[45]),"=r"(r[46]),"=r"(r[47]),
              "=r"(r[48]),"=r"(r[49]),"=r"(r[50]),"=r"(r[51]),"=r"(r[52]),"=r"(r[53]),"=r"(r[54]),"=r"(r[55]),"=r"(r[56]),"=r"(r[57]),"=r"(r[58]),"=r"(r[59]),"=r"(r[60]),"=r"(r[61]),"=r"(r[62]),"=r"(r[63])
            : "r"(tm + ((uint32_t)(q * 32) << 16) + 320) : "memory");
          tmem_wait_ld();
          float sacc = 0.f;
#pragma unroll
          for (int i = 0; i < 64; ++i) sacc += w[i];
          acc += sacc;
          ++n;
          continue;
        }
        tmem_ld32(addr, v);
        tmem_wait_ld();
        if (c.mode >= 8) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = ex2f(v[i] * 1.4426950f);
        }
        if (c.mode >= 7) {
          uint32_t p[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) p[i] = pack_bf16x2(v[2 * i], v[2 * i + 1]);
          tmem_st16(addr, p);
          tmem_wait_st();
        }
        acc += v[lane & 31];
        ++n;
      }
      if (acc == 123.456f) out[1] = n;
      if (lane == 0 && blockIdx.x == 0 && warp == 4) out[2] = n;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tm, 512); }
}

template <int A_TMEM, int B_MN>
int run(long long* out, int smem_bytes, int mode, int bgwarps = 0, int reps = 256) {
  cudaFuncSetAttribute(rate_kernel<A_TMEM, B_MN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  const int Ns[] = {32, 64, 96, 128, 192, 256};
  for (int N : Ns) {
    if (mode != 0 && N != 128 && N != 96) continue;
    if (reps == 0 && N != 128) continue;
    Case c{N, A_TMEM, B_MN, reps, mode, bgwarps};
    long long h[3];
    for (int rep = 0; rep < 2; ++rep) {
      rate_kernel<A_TMEM, B_MN><<<148, 640, smem_bytes>>>(c, out);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("N=%d a_tmem=%d b_mn=%d: %s\n", N, A_TMEM, B_MN, cudaGetErrorString(e)); return 1; }
    }
    cudaMemcpy(h, out, 24, cudaMemcpyDeviceToHost);
    printf("mode %d bg %2d A=%s B=%s N=%3d : issue %6.1f  complete %6.1f   (bg iterations per warp %lld)\n", mode, bgwarps, A_TMEM ? "tmem" : "smem", B_MN ? "MN-major" : "K-major ", N,
           h[0] / 256.0, h[1] / 256.0, mode >= 6 ? h[2] : 0LL);
  }
  return 0;
}

int main() {
  long long* out;
  cudaMalloc(&out, 32);
  const int smem_bytes = 128 * 1024 + 32768 + 1024;
  printf("M=128 K=16 bf16, 148 CTAs, 256 instructions in batches of 8: cycles per instruction (issue loop | until complete)\n");
  printf("between batches: mode 0 nothing, 1 tcgen05.fence::after_thread_sync, 2 tcgen05.commit, 3 mbarrier.test_wait (complete), "
         "4 mbarrier.try_wait (complete), 5 try_wait + fence + commit\n");
  if (run<0, 0>(out, smem_bytes, 0) || run<0, 1>(out, smem_bytes, 0) || run<1, 0>(out, smem_bytes, 0) || run<1, 1>(out, smem_bytes, 0)) return 1;
  for (int mode = 1; mode <= 5; ++mode)
    if (run<1, 0>(out, smem_bytes, mode)) return 1;
  printf("with background warps on the same SM: mode 6 = tcgen05.ld x32 loop, 7 = ld + st x16, 8 = ld + 32 ex2 + st (an epilogue)\n");
  for (int mode = 6; mode <= 9; ++mode)
    for (int bg : {4, 8, 16})
      if (run<1, 0>(out, smem_bytes, mode, bg)) return 1;
  printf("the same background loops with the tensor pipe IDLE (issuer spins 16000 cycles; divide by the iteration count)\n");
  for (int mode = 6; mode <= 9; ++mode)
    for (int bg : {4, 8, 16})
      if (run<1, 0>(out, smem_bytes, mode, bg, 0)) return 1;
  return 0;
}
