"""One eager launch of each big history-encoder kernel (for ncu).  usage: python tools/history_kernels_once.py"""
import sys, torch
sys.path.insert(0, ".")
from two_tower_models_b200 import ops
dev = torch.device("cuda:0")
B, H, D, heads = 8192, 50, 128, 4
M = B * H
bf = torch.bfloat16
x = torch.randn(M, D, device=dev).to(bf); w_in = torch.randn(3 * D, D, device=dev).to(bf) * 0.05
b_in = torch.zeros(3 * D, device=dev)
qkv = torch.empty(M, 3 * D, dtype=bf, device=dev)
do = torch.randn(M, D, device=dev).to(bf)
dw = torch.zeros(3 * D, D, device=dev)
for _ in range(2):
    ops.gemm(x, w_in, M, 3 * D, D, bias=b_in, out16=qkv)
    ops.gemm(qkv, x, 3 * D, D, M, a_mn=True, b_mn=True, out32=dw, accumulate=True)
    ops.attn_forward(qkv, B, H, D, heads, H)
    ops.attn_backward(qkv, do, B, H, D, heads, H)
torch.cuda.synchronize()
