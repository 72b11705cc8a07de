"""Executed warp-instruction mix of a kernel from an ncu report (source page).
usage: python tools/ncu_inst_mix.py report.ncu-rep kernel_regex [per_unit_divisor]"""
import csv, subprocess, sys, io, collections, re
rep, rx = sys.argv[1], sys.argv[2]
div = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{rx}"],
                     capture_output=True, text=True).stdout
blk = out.split('"Kernel Name",')[1]
lines = blk.split("\n")
print("kernel:", lines[0][:120])
rows = list(csv.reader(io.StringIO("\n".join(lines[1:]))))
hdr = rows[0]
iS, iE = hdr.index("Source"), hdr.index("Instructions Executed")
mix = collections.Counter()
tot = 0
for r in rows[1:]:
    if len(r) <= iE:
        continue
    ex = int(r[iE] or 0)
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[iS])
    op = m.group(2) if m else "?"
    mix[op] += ex
    tot += ex
print(f"total warp instructions {tot}  ({tot/div:.1f} per unit)")
for op, n in mix.most_common(28):
    print(f"  {op:12s} {n:10d}  {100.0*n/tot:5.1f}%  {n/div:8.1f} per unit")
