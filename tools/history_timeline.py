"""Kernel timeline of one CUDA-graph replay of the history-encoder training step (BASELINE configs[2]): torch.profiler /
CUPTI durations summed per kernel name, plus the ordered list.  usage: python tools/history_timeline.py"""
import sys, torch
sys.path.insert(0, ".")
import two_tower_models_b200 as tt
import bench
from torch.profiler import profile, ProfilerActivity

dev = torch.device("cuda:0")
B, d, F, H, L = 8192, 128, 128, 50, 2
torch.manual_seed(0)
m = tt.TwoTowerWithUserHistoryEncoder(100, bench.HASH, d, F, H, bench.HASH, d, F, [1.0], tt.BaselineMIPSModule(16, d),
                                      num_attention_heads=4, num_attention_layers=L).to(dev)
gen = torch.Generator().manual_seed(1)
b = bench.make_batch(B, F, gen)
b["user_history"] = torch.randint(0, bench.HASH, (B, H), generator=gen)
b = {k: v.to(dev) for k, v in b.items()}

def full():
    m._packed.invalidate(); m.user_history_encoder._packed.invalidate()
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
for _ in range(3): g.replay()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    g.replay(); torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ev.sort(key=lambda e: e.time_range.start)
t0 = ev[0].time_range.start
for e in ev:
    print(f"{e.time_range.start - t0:8.1f} us  +{e.time_range.end - e.time_range.start:7.1f} us  {e.name[:100]}")
print(f"total span {ev[-1].time_range.end - t0:.1f} us")
agg = {}
for e in ev:
    k = e.name[:70]
    a = agg.setdefault(k, [0.0, 0]); a[0] += e.time_range.end - e.time_range.start; a[1] += 1
print("--- per kernel name")
for k, (t, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"{t:8.1f} us  x{n:3d}  {k}")
