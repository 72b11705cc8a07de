"""CE forward / backward time (graph-replayed, us) at B = N = 8192.  usage: python tools/ce_time.py [d ...]"""
import sys, torch
sys.path.insert(0, ".")
from two_tower_models_b200 import ops
sys.path.insert(0, "tools")
from bench_kernels import timeit_graph
dev = torch.device("cuda:0")
M = 8192
for d in [int(x) for x in sys.argv[1:]] or [128]:
    U = (torch.randn(M, d, device=dev) * 0.4).to(torch.bfloat16)
    V = (torch.randn(M, d, device=dev) * 0.4).to(torch.bfloat16)
    ce, lse = ops.inbatch_ce_forward_raw(U, V, M, M, d, 0)
    g = torch.full((M,), 1.0 / M, device=dev)
    f = timeit_graph(lambda: ops.inbatch_ce_forward_raw(U, V, M, M, d, 0), reps=5)
    b = timeit_graph(lambda: ops.inbatch_ce_backward_raw(U, V, M, M, d, 0, lse, g), reps=5)
    print(f"d={d}: ce fwd {f:7.2f} us ({2.0*M*M*d/f/1e6:6.1f} TF/s)   bwd {b:7.2f} us ({4.0*M*M*d/b/1e6:6.1f} TF/s)", flush=True)
