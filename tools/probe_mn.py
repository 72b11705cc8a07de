"""Bring-up probe: which MN-major shared-memory descriptor strides does the hardware accept?"""
import os, subprocess, sys
CASE = r'''
import torch, sys
sys.path.insert(0, ".")
from two_tower_models_b200 import ops
g = torch.Generator().manual_seed(1)
M,N,K = 128,128,128
A = torch.randint(-3,4,(M,K),generator=g).float(); B = torch.randint(-3,4,(N,K),generator=g).float()
ref = A@B.t()
dev = torch.device("cuda:0")
for a_mn,b_mn in ((False,False),(True,False),(False,True),(True,True)):
    A16 = ops.cast_rows_bf16((A.t().contiguous() if a_mn else A).to(dev))
    B16 = ops.cast_rows_bf16((B.t().contiguous() if b_mn else B).to(dev))
    out = torch.zeros((M,N),device=dev)
    try:
        ops.gemm(A16,B16,M,N,K,a_mn=a_mn,b_mn=b_mn,out32=out); torch.cuda.synchronize()
        bad = int((out.cpu()!=ref).sum())
    except Exception as e:
        bad = repr(e)[:100]
    print(f"   a_mn={int(a_mn)} b_mn={int(b_mn)} mismatches={bad}", flush=True)
'''
for var in ("", "8192,1024,2048", "1024,8192,2048", "8192,1024,1024", "1024,8192,1024", "16384,1024,2048"):
    env = dict(os.environ)
    if var: env["TT_DBG_MN"] = var
    print("TT_DBG_MN=", var or "(default)", flush=True)
    r = subprocess.run([sys.executable, "-c", CASE], env=env, capture_output=True, text=True, timeout=120)
    print(r.stdout, r.stderr[-400:] if r.returncode else "", flush=True)
