"""Bring-up: CE forward time and effective SM clock with parts of the epilogue disabled (TT_CE_DBG bits:
1 no ex2, 2 no max, 4 no tile update at all, 8 no TMEM loads, 16 = stamp clock64/globaltimer of CTA 0)."""
import os, sys, torch
sys.path.insert(0, ".")
from two_tower_models_b200 import ops
sys.path.insert(0, "tools")
from bench_kernels import timeit_graph
dev = torch.device("cuda:0")
M = 8192
d = int(sys.argv[1]) if len(sys.argv) > 1 else 128
U = (torch.randn(M, d, device=dev) * 0.4).to(torch.bfloat16)
V = (torch.randn(M, d, device=dev) * 0.4).to(torch.bfloat16)
tr = torch.zeros(4 * 64 * 2, dtype=torch.int64, device=dev)
os.environ["TT_CE_TRACE"] = str(tr.data_ptr())
for dbg in [0, 1, 32, 33, 36]:
    os.environ["TT_CE_DBG"] = str(dbg | 16)
    for _ in range(20):
        ops.inbatch_ce_forward_raw(U, V, M, M, d, 0)
    torch.cuda.synchronize()
    t = tr.cpu()
    cyc, ns = int(t[2] - t[0]), int(t[3] - t[1])
    f = timeit_graph(lambda: ops.inbatch_ce_forward_raw(U, V, M, M, d, 0), reps=5)
    print(f"TT_CE_DBG={dbg:2d}: fwd {f:7.2f} us | CTA0 {cyc} cycles in {ns} ns = {cyc/max(ns,1):.3f} GHz, {cyc/28:.0f} cycles/tile", flush=True)
