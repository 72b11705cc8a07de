"""Top stall locations of a kernel from an ncu report (source page, SASS view).
usage: python tools/ncu_source_top.py report.ncu-rep kernel_regex [N]"""
import csv, subprocess, sys, io
rep, rx = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{rx}"],
                     capture_output=True, text=True).stdout
blocks = out.split('"Kernel Name",')
for blk in blocks[1:2]:
    lines = blk.split("\n")
    print("kernel:", lines[0][:120])
    rows = list(csv.reader(io.StringIO("\n".join(lines[1:]))))
    hdr = rows[0]
    iS, iA, iN, iE = hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Warp Stall Sampling (Not-issued Samples)"), hdr.index("Instructions Executed")
    data = [(int(r[iA] or 0), int(r[iN] or 0), int(r[iE] or 0), k, r[iS].strip()) for k, r in enumerate(rows[1:]) if len(r) > iE]
    tot = sum(d[0] for d in data)
    print(f"total samples {tot}, instructions {len(data)}")
    for a, ni, ex, k, src in sorted(data, reverse=True)[:n]:
        print(f"{100.0*a/tot:5.1f}%  samples={a:6d} notissued={ni:6d} exec={ex:8d}  [{k:5d}] {src[:90]}")
