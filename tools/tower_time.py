"""In-graph time of the tower forward pieces at the benchmark shape (B=8192, F=D=DI=128, two towers)."""
import sys, torch
sys.path.insert(0, "."); sys.path.insert(0, "tools")
from two_tower_models_b200 import ops
from bench_kernels import timeit_graph
dev = torch.device("cuda:0")
B, F, D = 8192, 128, 128
g = torch.Generator().manual_seed(0)
def mk():
    d = dict(ids=torch.randint(0, 100000, (B,), generator=g).to(dev), feats=torch.randn(B, F, generator=g).to(dev),
             table=torch.randn(100000, D, generator=g).to(dev),
             w0=torch.randn(256, F, generator=g).to(dev) * 0.1, w1=torch.randn(D, 256, generator=g).to(dev) * 0.1,
             wt=torch.randn(D, 2 * D, generator=g).to(dev) * 0.1,
             b0=torch.randn(256, generator=g).to(dev), b1=torch.randn(D, generator=g).to(dev), bt=torch.randn(D, generator=g).to(dev),
             feats16=torch.empty(B, F, dtype=torch.bfloat16, device=dev), H16=torch.empty(B, 256, dtype=torch.bfloat16, device=dev),
             X16=torch.empty(B, 2 * D, dtype=torch.bfloat16, device=dev), emb=torch.empty(B, D, device=dev),
             emb16=torch.empty(B, D, dtype=torch.bfloat16, device=dev), B=B, F=F, D=D, DI=D, hid=256)
    for k in ("w0", "w1", "wt"):
        d[k + "_16"] = torch.empty(d[k].shape, dtype=torch.bfloat16, device=dev)
    return d
tw = [mk(), mk()]
casts = [(d[k], 0, d[k].shape[1], d[k + "_16"], 0, d[k].shape[1]) for d in tw for k in ("w0", "w1", "wt")]
ops.cast_batched(casts)
print(f"cast of 6 weights (one launch)   : {timeit_graph(lambda: ops.cast_batched(casts)):7.2f} us")
print(f"fused tower forward, 2 towers    : {timeit_graph(lambda: ops.tower_forward_fused(tw)):7.2f} us")
print(f"fused tower forward, 1 tower     : {timeit_graph(lambda: ops.tower_forward_fused(tw[:1])):7.2f} us")

import os
tr = torch.zeros(16, dtype=torch.int64, device=dev)
os.environ["TT_TOWER_TRACE"] = str(tr.data_ptr())
for _ in range(3):
    ops.tower_forward_fused(tw)
torch.cuda.synchronize()
t = tr.cpu().tolist()
names = ["start", "setup done", "operands in smem", "acc1 ready", "H in TMEM", "acc2 ready", "Fe in TMEM", "acc3 ready", "emb stored"]
for i, n in enumerate(names):
    print(f"  {n:18s} +{t[i] - t[0]:6d} ns")
