"""Micro-benchmarks of the native kernels on one B200 (CUDA events, L2-warm repeated launches).
Usage: python tools/bench_kernels.py [gemm] [ce] [mips] [attn]   -> prints one line per case."""
import sys
import time

import torch

sys.path.insert(0, ".")
from two_tower_models_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")


def timeit(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3  # us


def timeit_graph(fn, reps=20, iters=10):
    """Per-launch device time without host launch overhead: `reps` launches captured in one CUDA graph."""
    fn()
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (iters * reps) * 1e3


def bench_small():
    """Launch-latency-bound kernels of the base step, timed back-to-back inside a CUDA graph."""
    M = 8192
    for N, K, a_mn, b_mn, acc, note in [(256, 128, 0, 0, 0, "mlp0 fwd"), (128, 256, 0, 0, 0, "mlp1 fwd"),
                                        (256, 128, 0, 1, 0, "dX"), (128, 8192, 1, 1, 1, "dW0 (M=256)"),
                                        (256, 8192, 1, 1, 1, "dW1 (M=128)")]:
        m = 256 if "dW0" in note else (128 if "dW1" in note else M)
        A = torch.randn((K, m) if a_mn else (m, K), device=dev).to(torch.bfloat16)
        B = torch.randn((K, N) if b_mn else (N, K), device=dev).to(torch.bfloat16)
        out32 = torch.zeros((m, N), device=dev) if acc else None
        out16 = None if acc else torch.empty((m, N), dtype=torch.bfloat16, device=dev)
        bias = None if acc else torch.randn(N, device=dev)
        us = timeit_graph(lambda: ops.gemm(A, B, m, N, K, a_mn=bool(a_mn), b_mn=bool(b_mn), out32=out32, out16=out16,
                                           bias=bias, accumulate=bool(acc)))
        print(f"graph gemm {note:12s} M={m} N={N} K={K}: {us:7.2f} us", flush=True)
    x = torch.randn(M, 128, device=dev)
    print(f"graph cast_rows 8192x128: {timeit_graph(lambda: ops.cast_rows_bf16(x)):7.2f} us")
    x16 = x.to(torch.bfloat16)
    print(f"graph colsum 8192x128 bf16: {timeit_graph(lambda: ops.colsum(x16, 128)):7.2f} us")
    table = torch.randn(100000, 128, device=dev)
    ids = torch.randint(0, 100000, (M,), device=dev)
    out = torch.empty(M, 256, dtype=torch.bfloat16, device=dev)
    print(f"graph gather 8192 rows: {timeit_graph(lambda: ops.gather_rows(table, ids, out)):7.2f} us")
    print(f"graph scatter_add (incl. 51 MB memset): {timeit_graph(lambda: ops.scatter_add_rows(x16, ids, 128, 100000)):7.2f} us")
    y = torch.empty(1 << 20, device=dev)
    print(f"graph torch fill 4MB: {timeit_graph(lambda: y.zero_()):7.2f} us")
    U = (torch.randn(M, 128, device=dev) * 0.4).to(torch.bfloat16)
    V = (torch.randn(M, 128, device=dev) * 0.4).to(torch.bfloat16)
    ce, lse = ops.inbatch_ce_forward_raw(U, V, M, M, 128, 0)
    g = torch.full((M,), 1.0 / M, device=dev)
    print(f"graph ce fwd 8192^2 d128: {timeit_graph(lambda: ops.inbatch_ce_forward_raw(U, V, M, M, 128, 0), reps=5):7.2f} us")
    print(f"graph ce bwd 8192^2 d128: {timeit_graph(lambda: ops.inbatch_ce_backward_raw(U, V, M, M, 128, 0, lse, g), reps=5):7.2f} us")


def bench_gemm():
    cases = [  # M, N, K, a_mn, b_mn, accumulate, note
        (8192, 256, 128, 0, 0, 0, "mlp layer0 fwd"),
        (8192, 128, 256, 0, 0, 0, "mlp layer1 / tower fwd"),
        (8192, 256, 128, 0, 1, 0, "dX = demb Wt"),
        (256, 128, 8192, 1, 1, 1, "dW0 split-K"),
        (128, 256, 8192, 1, 1, 1, "dW1 / dWt split-K"),
        (409600, 384, 128, 0, 0, 0, "history in-proj (C3)"),
        (409600, 128, 128, 0, 0, 0, "history out-proj (C3)"),
        (384, 128, 409600, 1, 1, 1, "history d_in_w split-K"),
        (8192, 8192, 8192, 0, 0, 0, "square 8k"),
    ]
    for M, N, K, a_mn, b_mn, acc, note in cases:
        A = torch.randn((K, M) if a_mn else (M, K), device=dev).to(torch.bfloat16)
        B = torch.randn((K, N) if b_mn else (N, K), device=dev).to(torch.bfloat16)
        out32 = torch.zeros((M, N), device=dev) if acc else None
        out16 = None if acc else torch.empty((M, N), dtype=torch.bfloat16, device=dev)
        us = timeit(lambda: ops.gemm(A, B, M, N, K, a_mn=bool(a_mn), b_mn=bool(b_mn), out32=out32, out16=out16,
                                     accumulate=bool(acc)))
        print(f"gemm {note:28s} M={M} N={N} K={K}: {us:9.1f} us  {2.0*M*N*K/us/1e6:8.1f} TFLOP/s", flush=True)


def bench_ce():
    for B, N, d in [(8192, 8192, 128), (8192, 8192, 256), (8192, 65536, 256), (8192, 16384, 128)]:
        U = (torch.randn(B, d, device=dev) * 0.4).to(torch.bfloat16)
        V = (torch.randn(N, d, device=dev) * 0.4).to(torch.bfloat16)
        ce, lse = ops.inbatch_ce_forward_raw(U, V, B, N, d, 0)
        g = torch.full((B,), 1.0 / B, device=dev)
        us_f = timeit(lambda: ops.inbatch_ce_forward_raw(U, V, B, N, d, 0))
        us_b = timeit(lambda: ops.inbatch_ce_backward_raw(U, V, B, N, d, 0, lse, g))
        print(f"ce B={B} N={N} d={d}: fwd {us_f:8.1f} us ({2.0*B*N*d/us_f/1e6:7.1f} TF/s alg)  "
              f"bwd {us_b:8.1f} us ({4.0*B*N*d/us_b/1e6:7.1f} TF/s alg)", flush=True)


def bench_mips():
    import two_tower_models_b200 as tt

    for nq, nc, d, k in [(4096, 1_000_000, 128, 100), (16384, 1_000_000, 128, 100), (256, 1_000_000, 128, 100)]:
        m = tt.BaselineMIPSModule(nc, d).to(dev)
        q = torch.randn(nq, d, device=dev)
        c16 = m._packed.get("corpus", m.corpus)
        t0 = time.perf_counter()
        ops.mips_topk(q, m.corpus, c16, k)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        us = timeit(lambda: ops.mips_topk(q, m.corpus, c16, k), iters=3, warm=1)
        print(f"mips nq={nq} nc={nc} d={d} k={k}: {us/1e3:9.2f} ms  {nq/us*1e6:9.0f} q/s  "
              f"{2.0*nq*nc*d/us/1e6:7.1f} TF/s (first call {1e3*(t1-t0):.1f} ms)", flush=True)
        del m


def bench_attn():
    B, H, D, heads = 8192, 50, 128, 4
    qkv = torch.randn(B * H, 3 * D, device=dev).to(torch.bfloat16)
    for q_rows in (H, 1):
        us_f = timeit(lambda: ops.attn_forward(qkv, B, H, D, heads, q_rows), iters=5)
        do = torch.randn(B * q_rows, D, device=dev).to(torch.bfloat16)
        us_b = timeit(lambda: ops.attn_backward(qkv, do, B, H, D, heads, q_rows), iters=5)
        print(f"attn B={B} H={H} D={D} heads={heads} q_rows={q_rows}: fwd {us_f:8.1f} us  bwd {us_b:8.1f} us", flush=True)


if __name__ == "__main__":
    which = sys.argv[1:] or ["gemm", "ce", "mips", "attn"]
    for w in which:
        globals()["bench_" + w]()
