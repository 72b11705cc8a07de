"""In-graph time of the batched dX / dH gradient GEMMs of the tower backward with and without the fused column sums."""
import sys, torch
sys.path.insert(0, "."); sys.path.insert(0, "tools")
from two_tower_models_b200 import ops
from bench_kernels import timeit_graph
dev = torch.device("cuda:0")
B = 8192
def mk(n):
    return [dict(A=torch.randn(B, 128, device=dev).to(torch.bfloat16), Bm=torch.randn(128, 256, device=dev).to(torch.bfloat16),
                 out=torch.empty(B, 256, dtype=torch.bfloat16, device=dev), cs=torch.zeros(256, device=dev),
                 mask=torch.randn(B, 256, device=dev).to(torch.bfloat16)) for _ in range(n)]
for n in (1, 2):
    ps = mk(n)
    for name, kw in [("plain", {}), ("colsum", {"colsum": True}), ("mask+colsum", {"colsum": True, "mask": True})]:
        def run():
            ops.gemm_batched([dict(A=p["A"], B=p["Bm"], M=B, N=256, K=128, b_mn=True, out16=p["out"],
                                   **({"colsum": p["cs"]} if kw.get("colsum") else {}),
                                   **({"relu_mask": p["mask"]} if kw.get("mask") else {})) for p in ps])
        print(f"{n} problem(s) M=8192 N=256 K=128 b_mn {name:12s}: {timeit_graph(run):6.2f} us", flush=True)
