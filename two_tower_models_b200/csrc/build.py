"""Build libtt_b200.so in-tree with nvcc for sm_100a (no torch headers involved).

    python two_tower_models_b200/csrc/build.py [--force] [--verbose]

The library links cudart statically and resolves cuTensorMapEncodeTiled at run time through
cudaGetDriverEntryPoint, so it loads on a machine without a GPU driver (symbol-export test).
"""
import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SOURCES = ["api.cu", "gemm.cu", "ce.cu", "ce_bwd2.cu", "ce_bwd3.cu", "ce_bwd3x.cu", "elementwise.cu", "mips.cu", "attn.cu", "attn_tc.cu", "history_last.cu", "tower.cu", "tower_bwd.cu"]  # missing files are skipped
HEADERS = ["common.cuh", "ce_common.cuh", "kernels.h", os.path.join(ROOT, "include", "tt_b200.h")]
LIB = os.environ.get("TT_B200_LIB_OUT") or os.path.join(HERE, "libtt_b200.so")  # bring-up builds go to another file
VARIANT = os.path.splitext(os.path.basename(LIB))[0]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xptxas", "-v",
]
if os.environ.get("TT_CE_BWD_LEAN") == "1":  # candidate epilogue of the CE backward (ce_bwd2.cu), not yet measured
    FLAGS.append("-DTT_CE_BWD_LEAN")
if os.environ.get("TT_CE_POLY") == "1":  # every fourth exponential of the v3 CE backward on the FMA pipe (ce_bwd3.cu)
    FLAGS.append("-DTT_CE_POLY")
if os.environ.get("TT_CE_FWD_POLY") == "1":  # every fourth exponential of the CE forward on the FMA pipe
    FLAGS.append("-DTT_CE_FWD_POLY")
if os.environ.get("TT_CE3_CW"):  # score columns per epilogue thread of the v3 CE backward (16 -> 24 epilogue warps)
    FLAGS.append("-DTT_CE3_CW=" + os.environ["TT_CE3_CW"])
for _k in ("TT_MIPS_CW", "TT_MIPS_SLEEP", "TT_MIPS_TRIG"):  # MIPS epilogue experiments: chunk width (16 / 32), sleep (ns) between barrier probes
    if os.environ.get(_k):
        FLAGS.append("-D%s=%s" % (_k, os.environ[_k]))
if os.environ.get("TT_MIPS_BRINGUP") == "1":  # per-tile clock stamps / partial epilogues of the MIPS screen kernel (tools/mips_trace.py)
    FLAGS.append("-DTT_MIPS_BRINGUP")
if os.environ.get("TT_CE_BRINGUP") == "1":  # clock64 timelines / partial epilogues of the CE kernels (tools/trace_ce.py)
    FLAGS.append("-DTT_CE_BRINGUP")


def _digest(paths):
    h = hashlib.sha256()
    for p in paths:
        with open(p, "rb") as f:
            h.update(f.read())
    h.update(" ".join(FLAGS).encode())
    return h.hexdigest()


def build(force=False, verbose=False):
    srcs = [os.path.join(HERE, s) for s in SOURCES if os.path.exists(os.path.join(HERE, s))]
    hdrs = [h if os.path.isabs(h) else os.path.join(HERE, h) for h in HEADERS]
    stamp = os.path.join(HERE, ".build_stamp" + ("" if VARIANT == "libtt_b200" else "." + VARIANT))
    dig = _digest(srcs + hdrs)
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == dig:
        return LIB
    objs = []

    def compile_one(src):
        obj = src[:-3] + ("" if VARIANT == "libtt_b200" else "." + VARIANT) + ".o"
        cmd = [NVCC] + FLAGS + ["-c", src, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        return src, obj, r

    with ThreadPoolExecutor(max_workers=8) as ex:
        results = list(ex.map(compile_one, srcs))
    log = []
    for src, obj, r in results:
        err = "\n".join(l for l in r.stderr.splitlines() if "Compile time" not in l)
        log.append(f"== {os.path.basename(src)}\n{err}")
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError(f"nvcc failed on {src}")
        objs.append(obj)
    with open(os.path.join(HERE, "ptxas_report.txt" if VARIANT == "libtt_b200" else "ptxas_report." + VARIANT + ".txt"), "w") as f:
        f.write("\n".join(log))
    if verbose:
        print("\n".join(log))
    cmd = [NVCC, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("link failed")
    with open(stamp, "w") as f:
        f.write(dig)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
