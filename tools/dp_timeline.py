"""Kernel timeline (torch.profiler / CUPTI) of one CUDA-graph replay of the batch-sharded training step on rank 0.
Run it under a SHORT timeout (`timeout 120 ...`): with 8 ranks under the profiler the process-group shutdown has hung
once and, at 8x GPU-minutes per second, spent the rest of a round's GPU budget.  The script therefore leaves with
os._exit after printing instead of tearing the group down.
usage: timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 tools/dp_timeline.py [d]"""
import os, sys, torch
import torch.distributed as dist
sys.path.insert(0, ".")
import two_tower_models_b200 as tt
from two_tower_models_b200 import distributed as ttd
from two_tower_models_b200.graph import GraphedTrainStep
import bench
from torch.profiler import profile, ProfilerActivity

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
d = int(sys.argv[1]) if len(sys.argv) > 1 else 128
B, F = 8192, 128
torch.manual_seed(0)
m = tt.TwoTowerBaseRetrieval(100, bench.HASH, d, F, bench.HASH, d, F, [1.0], tt.BaselineMIPSModule(16, d)).to(dev)
ctx = ttd.enable_data_parallel(m)
b = {k: v.to(dev) for k, v in bench.make_batch(B, F, torch.Generator().manual_seed(1 + rank)).items()}
g = GraphedTrainStep(m, b, post_backward=ctx.sync_gradients)
for _ in range(5):
    g(b)
torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    g(b); torch.cuda.synchronize()
dist.barrier()
if rank == 0:
    ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    ev.sort(key=lambda e: e.time_range.start)
    t0 = ev[0].time_range.start
    busy_end = 0.0
    for e in ev:
        s, t = e.time_range.start - t0, e.time_range.end - t0
        gap = s - busy_end if s > busy_end else 0.0
        busy_end = max(busy_end, t)
        print(f"{s:8.1f} us  +{t - s:7.1f} us  gap {gap:6.1f}  {e.name[:95]}")
    print(f"total span {ev[-1].time_range.end - t0:.1f} us", flush=True)
sys.stdout.flush()
os._exit(0)
