"""Where does the base training step spend its time?  Sub-sequences of the step captured as CUDA graphs."""
import sys, torch
sys.path.insert(0, ".")
import two_tower_models_b200 as tt
from two_tower_models_b200 import ops
import bench

dev = torch.device("cuda:0")
B, d, F = 8192, 128, 128
torch.manual_seed(0)
m = tt.TwoTowerBaseRetrieval(100, bench.HASH, d, F, bench.HASH, d, F, [1.0], tt.BaselineMIPSModule(16, d)).to(dev)
b = {k: v.to(dev) for k, v in bench.make_batch(B, F, torch.Generator().manual_seed(1)).items()}
O = bench.ORDER


def graph_time(fn, iters=50):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            fn()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


def towers_fwd():
    m._packed.invalidate()
    with torch.no_grad():
        fu, fi = m.user_features_arch, m.item_features_arch
        out = ops.TowerSetFunction.apply([("user", None), ("item", None)], m._packed,
            b["user_id"], b["user_features"], None, m.user_id_embedding_arch.weight, fu[0].weight, fu[0].bias, fu[2].weight, fu[2].bias, m.user_tower_arch.weight, m.user_tower_arch.bias,
            b["item_id"], b["item_features"], None, m.item_id_embedding_arch.weight, fi[0].weight, fi[0].bias, fi[2].weight, fi[2].bias, m.item_tower_arch.weight, m.item_tower_arch.bias)
        ops.join_pending_fills()  # train_forward does this after the loss forward
        return out

def fwd_only():
    m._packed.invalidate()
    with torch.no_grad():
        return m.train_forward(*[b[k] for k in O])

def full():
    m._packed.invalidate()
    for p in m.parameters(): p.grad = None
    l = m.train_forward(*[b[k] for k in O]); l.backward(); return l

U, V = towers_fwd()
U16, V16 = U._tt_bf16, V._tt_bf16
ce, lse = ops.inbatch_ce_forward_raw(U16, V16, B, B, d, 0)
g = torch.full((B,), 1.0 / B, device=dev)
print(f"towers forward (cast+gather+3 GEMM launches): {graph_time(towers_fwd):7.1f} us")
print(f"CE forward (+combine)                        : {graph_time(lambda: ops.inbatch_ce_forward_raw(U16, V16, B, B, d, 0)):7.1f} us")
print(f"train_forward (no grad)                      : {graph_time(fwd_only):7.1f} us")
print(f"CE backward (2 passes + reduce)              : {graph_time(lambda: ops.inbatch_ce_backward_raw(U16, V16, B, B, d, 0, lse, g)):7.1f} us")
print(f"full step fwd+bwd                            : {graph_time(full):7.1f} us")
z = torch.empty(bench.HASH, d, device=dev)
print(f"zero-fill of one dense table gradient        : {graph_time(lambda: z.zero_()):7.1f} us")
