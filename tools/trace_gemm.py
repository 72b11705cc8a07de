"""Bring-up: per-CTA %globaltimer stamps of one small GEMM (TT_GEMM_TRACE)."""
import os, sys
import torch
sys.path.insert(0, ".")
from two_tower_models_b200 import ops
dev = torch.device("cuda:0")
mode = sys.argv[1] if len(sys.argv) > 1 else "fwd"
if mode == "fwd":
    M, N, K = 8192, 256, 128
    A = torch.randn(M, K, device=dev).to(torch.bfloat16); B = torch.randn(N, K, device=dev).to(torch.bfloat16)
    out16 = torch.empty(M, N, dtype=torch.bfloat16, device=dev); bias = torch.randn(N, device=dev)
    run = lambda: ops.gemm(A, B, M, N, K, out16=out16, bias=bias)
else:  # split-K weight gradient
    M, N, K = 256, 128, 8192
    A = torch.randn(K, M, device=dev).to(torch.bfloat16); B = torch.randn(K, N, device=dev).to(torch.bfloat16)
    out32 = torch.zeros(M, N, device=dev)
    run = lambda: ops.gemm(A, B, M, N, K, a_mn=True, b_mn=True, out32=out32, accumulate=True)
for _ in range(3): run()
torch.cuda.synchronize()
tr = torch.zeros(148 * 8, dtype=torch.int64, device=dev)
os.environ["TT_GEMM_TRACE"] = str(tr.data_ptr())
for rep in range(2):
    tr.zero_(); torch.cuda.synchronize()
    run()
    torch.cuda.synchronize()
    t = tr.cpu().view(148, 8)
    t = t[t[:, 0] > 0]
    t0 = int(t[:, 0].min())
    names = ["start", "setup done", "first TMA landed", "acc ready", "epilogue done", "exit"]
    print(f"rep {rep}: {t.shape[0]} CTAs; ns since first CTA start (min / median / max over CTAs)")
    for i, n in enumerate(names):
        col = (t[:, i] - t0).float()
        print(f"  {n:18s} {col.min():8.0f} {col.median():8.0f} {col.max():8.0f}")
