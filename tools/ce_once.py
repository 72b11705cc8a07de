"""One CE forward + backward launch at B = N = 8192 (for ncu captures).  usage: python tools/ce_once.py [d]"""
import sys, torch
sys.path.insert(0, ".")
from two_tower_models_b200 import ops
dev = torch.device("cuda:0")
M = 8192
d = int(sys.argv[1]) if len(sys.argv) > 1 else 128
U = (torch.randn(M, d, device=dev) * 0.4).to(torch.bfloat16)
V = (torch.randn(M, d, device=dev) * 0.4).to(torch.bfloat16)
g = torch.full((M,), 1.0 / M, device=dev)
for _ in range(3):
    ce, lse = ops.inbatch_ce_forward_raw(U, V, M, M, d, 0)
    ops.inbatch_ce_backward_raw(U, V, M, M, d, 0, lse, g)
torch.cuda.synchronize()
