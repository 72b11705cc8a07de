"""Kernel timeline of one CUDA-graph replay of the base training step (torch.profiler / CUPTI): start offset, duration
and stream of every kernel and memset, to see the gaps and the overlap between the parallel branches."""
import sys, torch
sys.path.insert(0, ".")
import two_tower_models_b200 as tt
import bench
from torch.profiler import profile, ProfilerActivity

dev = torch.device("cuda:0")
B, d, F = 8192, 128, 128
torch.manual_seed(0)
m = tt.TwoTowerBaseRetrieval(100, bench.HASH, d, F, bench.HASH, d, F, [1.0], tt.BaselineMIPSModule(16, d)).to(dev)
b = {k: v.to(dev) for k, v in bench.make_batch(B, F, torch.Generator().manual_seed(1)).items()}

def full():
    m._packed.invalidate()
    for p in m.parameters():
        p.grad = None
    l = m.train_forward(*[b[k] for k in bench.ORDER]); l.backward(); return l

s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for _ in range(3): full()
torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    full()
for _ in range(5): g.replay()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    g.replay(); torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ev.sort(key=lambda e: e.time_range.start)
t0 = ev[0].time_range.start
for e in ev:
    print(f"{e.time_range.start - t0:8.1f} us  +{e.time_range.end - e.time_range.start:7.1f} us  {e.name[:90]}")
print(f"total span {ev[-1].time_range.end - t0:.1f} us")
