"""Bring-up: per-tile clock stamps of one epilogue warp of the MIPS screen kernel (TT_MIPS_TRACE build hook)."""
import os, sys
import numpy as np
import torch
os.environ["TT_MIPS_TRACE"] = "/tmp/mips_trace.bin"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import two_tower_models_b200 as tt
torch.manual_seed(0)
Q, C, d, k = 37888, 1_000_000, 128, 100
m = tt.BaselineMIPSModule(C, d).cuda()
q = torch.randn(Q, d, device="cuda")
idx, sc = m(q, k)[:2] if isinstance(m(q, k), tuple) else (m(q, k), None)
torch.cuda.synchronize()
t = np.fromfile("/tmp/mips_trace.bin", dtype=np.int64).reshape(-1, 16)
t = t[t[:, 0] != 0]
print("tiles traced", len(t), "total cycles", t[-1, 3] - t[0, 0])
wait = t[:, 1] - t[:, 0]; read = t[:, 2] - t[:, 1]; drain = t[:, 3] - t[:, 2]
period = np.diff(t[:, 0])
edges = [0, 4, 16, 64, 256, 1024, 2048, len(t)]
print("tile range      wait   read  drain  period (mean cycles of this warp's own tiles)")
for a, b in zip(edges[:-1], edges[1:]):
    if a >= len(t): break
    b = min(b, len(t))
    print(f"[{a:5d},{b:5d})  {wait[a:b].mean():7.0f} {read[a:b].mean():6.0f} {drain[a:b].mean():6.0f}  {period[a:b-1].mean() if b-1>a else 0:7.0f}   sum={(t[b-1,3]-t[a,0])/1e6:6.2f} Mcyc")
big = np.argsort(-(drain))[:10]
print("largest drains (own-tile index, cycles):", [(int(i), int(drain[i])) for i in big])

late = t[2048:]
print("late tiles: mean cycles from d_full to: " + " ".join(f"{(late[:, 4 + i] - late[:, 1]).mean():6.0f}" for i in range(8)) + "  (stamps after the wait of chunks 0..7)")
print("read end", (late[:, 2] - late[:, 1]).mean())

has = late[:, 12] != 0
print("late tiles with a drain:", has.mean(), " own-slot part", (late[has, 12] - late[has, 2]).mean(), " ballot", (late[has, 13] - late[has, 12]).mean(),
      " rest (compaction)", (late[has, 3] - late[has, 13]).mean(), " tiles with compaction", (late[has, 14] != 0).mean())
comp = late[has][late[has, 14] != 0]
if len(comp): print("compaction tiles: rest mean", (comp[:, 3] - comp[:, 13]).mean(), "rows per compaction tile", np.mean([bin(int(x)).count("1") for x in comp[:, 14]]))
nocomp = late[has][late[has, 14] == 0]
print("no-compaction drain total", (nocomp[:, 3] - nocomp[:, 2]).mean())
