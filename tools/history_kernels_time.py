"""Device time and achieved HBM GB/s of the history-encoder kernels at BASELINE configs[2] (B=8192, H=50, D=128, 4 heads),
each timed back to back inside a CUDA graph.  usage: python tools/history_kernels_time.py"""
import sys, torch
sys.path.insert(0, "."); sys.path.insert(0, "tools")
from two_tower_models_b200 import ops
from bench_kernels import timeit_graph
dev = torch.device("cuda:0")
B, H, D, heads = 8192, 50, 128, 4
M = B * H
bf = torch.bfloat16
x = torch.randn(M, D, device=dev).to(bf); w_in = torch.randn(3 * D, D, device=dev).to(bf) * 0.05; w_out = torch.randn(D, D, device=dev).to(bf) * 0.05
b_in = torch.zeros(3 * D, device=dev); b_out = torch.zeros(D, device=dev)
qkv = torch.empty(M, 3 * D, dtype=bf, device=dev); y = torch.empty(M, D, dtype=bf, device=dev); o = torch.empty(M, D, dtype=bf, device=dev)
def rep(name, us, bytes_, flops):
    print(f"{name:46s} {us:8.1f} us  {bytes_/us/1e3:7.0f} GB/s  {flops/us/1e6:7.1f} TFLOP/s", flush=True)
t = timeit_graph(lambda: ops.gemm(x, w_in, M, 3 * D, D, bias=b_in, out16=qkv), reps=5)
rep("in-proj  [M,128]x[384,128]^T -> bf16 [M,384]", t, M * D * 2 + M * 3 * D * 2, 2.0 * M * 3 * D * D)
t = timeit_graph(lambda: ops.gemm(o, w_out, M, D, D, bias=b_out, out16=y), reps=5)
rep("out-proj [M,128]x[128,128]^T -> bf16 [M,128]", t, M * D * 4, 2.0 * M * D * D)
dx = torch.empty(M, D, dtype=bf, device=dev); cs = torch.zeros(D, device=dev)
t = timeit_graph(lambda: ops.gemm(qkv, w_in, M, D, 3 * D, b_mn=True, out16=dx, colsum=cs), reps=5)
rep("dx = dqkv W_in [M,384]x[384,128] -> bf16 + colsum", t, M * 3 * D * 2 + M * D * 2, 2.0 * M * 3 * D * D)
dw = torch.zeros(3 * D, D, device=dev)
t = timeit_graph(lambda: ops.gemm(qkv, x, 3 * D, D, M, a_mn=True, b_mn=True, out32=dw, accumulate=True), reps=5)
rep("d_in_w = dqkv^T x (split-K over M)", t, M * 3 * D * 2 + M * D * 2, 2.0 * M * 3 * D * D)
dwo = torch.zeros(D, D, device=dev)
t = timeit_graph(lambda: ops.gemm(y, o, D, D, M, a_mn=True, b_mn=True, out32=dwo, accumulate=True), reps=5)
rep("d_out_w = dy^T o (split-K over M)", t, M * D * 4, 2.0 * M * D * D)
t = timeit_graph(lambda: ops.colsum(qkv, 3 * D), reps=5)
rep("colsum of dqkv [M,384]", t, M * 3 * D * 2, 0)
t = timeit_graph(lambda: ops.attn_forward(qkv, B, H, D, heads, H), reps=5)
rep("attention forward, all rows", t, M * 3 * D * 2 + M * D * 2, 4.0 * B * heads * H * H * (D // heads))
t = timeit_graph(lambda: ops.attn_forward(qkv, B, H, D, heads, 1), reps=5)
rep("attention forward, row 0 only (last layer)", t, M * 2 * D * 2 + B * D * 4, 4.0 * B * heads * H * (D // heads))
do = torch.randn(M, D, device=dev).to(bf)
t = timeit_graph(lambda: ops.attn_backward(qkv, do, B, H, D, heads, H), reps=5)
rep("attention backward, all rows", t, M * 3 * D * 2 * 2 + M * D * 2, 10.0 * B * heads * H * H * (D // heads))
do1 = torch.randn(B, D, device=dev).to(bf)
t = timeit_graph(lambda: ops.attn_backward(qkv, do1, B, H, D, heads, 1), reps=5)
rep("attention backward, row 0 only (last layer)", t, M * 3 * D * 2 * 2, 10.0 * B * heads * H * (D // heads))
