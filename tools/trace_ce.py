"""Bring-up: clock64 timeline of CTA 0 of the CE backward kernel (TT_CE_TRACE)."""
import os, sys
import torch
sys.path.insert(0, ".")
from two_tower_models_b200 import ops
dev = torch.device("cuda:0")
B = N = 8192; d = int(sys.argv[1]) if len(sys.argv) > 1 else 128
U = (torch.randn(B, d, device=dev) * 0.4).to(torch.bfloat16); V = (torch.randn(N, d, device=dev) * 0.4).to(torch.bfloat16)
ce, lse = ops.inbatch_ce_forward_raw(U, V, B, N, d, 0)
g = torch.full((B,), 1.0 / B, device=dev)
for _ in range(2): ops.inbatch_ce_backward_raw(U, V, B, N, d, 0, lse, g)
torch.cuda.synchronize()
tr = torch.zeros(8 * 64 * 2, dtype=torch.int64, device=dev)
os.environ["TT_CE_TRACE"] = str(tr.data_ptr())
ct = torch.zeros(148 * 4, dtype=torch.int64, device=dev)
os.environ["TT_CE_CTA_TIMES"] = str(ct.data_ptr())
lib = __import__("two_tower_models_b200._native", fromlist=["lib"]).lib()
if len(sys.argv) > 2 and sys.argv[2] == "fwd":
    ops.inbatch_ce_forward_raw(U, V, B, N, d, 0)
    torch.cuda.synchronize()
    t = tr.cpu().view(8, 64, 2)
    t0 = int(t[t > 0].min())
    print("tile | MMA loop top, Y landed, S issued | epilogue g0 start->end | epilogue g1 start->end")
    for i in range(40):
        g = lambda r, w: int(t[r, i, w]) - t0
        print(f"{i:3d} | {g(0,0):7d} {g(1,0):7d} {g(0,1):7d} | {g(2,0):7d} {g(2,1):7d} | {g(3,0):7d} {g(3,1):7d}")
    sys.exit(0)
# only ONE pass (the other outputs NULL) so that the trace is not overwritten by the second pass: "dv" selects the dV pass
dU = torch.empty(B, d, device=dev); ws = ops._ce_workspace(B, N, d, dev)
if len(sys.argv) > 2 and sys.argv[2] == "dv":
    rc = lib.tt_inbatch_ce_bwd(U.data_ptr(), U.stride(0), V.data_ptr(), V.stride(0), B, N, d, 0, lse.data_ptr(), g.data_ptr(),
                               None, 0, None, 0, dU.data_ptr(), dU.stride(0), None, 0, None, None, ws.data_ptr(), ws.numel(),
                               torch.cuda.current_stream().cuda_stream)
else:
    rc = lib.tt_inbatch_ce_bwd(U.data_ptr(), U.stride(0), V.data_ptr(), V.stride(0), B, N, d, 0, lse.data_ptr(), g.data_ptr(),
                               dU.data_ptr(), dU.stride(0), None, 0, None, 0, None, 0, None, None, ws.data_ptr(), ws.numel(),
                               torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize()
c = ct.cpu().view(148, 4); c = c[c[:, 0] > 0]
c0 = int(c[:, 0].min())
for i, nm in enumerate(["start", "setup", "work done", "exit"]):
    col = (c[:, i] - c0).float()
    print(f"CTA {nm:10s} ns: min {col.min():8.0f} median {col.median():8.0f} max {col.max():8.0f}")
dur = (c[:, 2] - c[:, 1]).float()
T = (64 * 64 + 146) // 147
print("work us by CTA (b = has a row-tile boundary inside its range):")
print(" ".join(f"{i}{'b' if (i * T) // 64 != ((i + 1) * T - 1) // 64 else ''}:{float(d)/1e3:.0f}" for i, d in enumerate(dur)))
print("per-CTA work ns: min %.0f median %.0f max %.0f; slowest CTAs: %s" % (dur.min(), dur.median(), dur.max(), torch.topk(dur, 5).indices.tolist()))
t = tr.cpu().view(8, 64, 2)
t0 = int(t[t > 0].min())
print("tile | S-MMA wait->issue | PV-MMA wait->issue | epilogue g0 start->end | epilogue g1 start->end   (cycles since start)")
print("     (v3 bring-up build, after the 4 columns: S batch entry / UMMAs issued | E Y batch barrier passed / UMMAs issued | epilogue group 0 of lane quarters 2 and 3 (columns 3, 4: quarters 0 and 1))")
for i in list(range(40)) + [60, 61]:  # 60 / 61: (v3) segment boundary - MMA warp waits for X-in-TMEM / drained accumulator; epilogue stages X / drains
    def f(r):
        a, b = int(t[r, i, 0]), int(t[r, i, 1])
        return f"{a - t0 if a else -1:7d} {b - t0 if b else -1:7d}"
    print(f"{i:3d} | {f(0)} | {f(1)} | {f(2)} | {f(3)} || {f(4)} | {f(5)} | {f(6)} | {f(7)}")
