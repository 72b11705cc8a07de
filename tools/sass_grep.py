"""Per-kernel counts of the Blackwell-specific SASS mnemonics in libtt_b200.so (evidence that the contractions run on
tcgen05 with TMEM accumulators and TMA-staged operands).  usage: python tools/sass_grep.py > profiles/r02_sass_grep.txt"""
import collections, re, subprocess, sys
lib = sys.argv[1] if len(sys.argv) > 1 else "two_tower_models_b200/csrc/libtt_b200.so"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
keys = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "MUFU.EX2", "HMMA", "ELECT"]
cur, counts = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    for k in keys:
        if re.search(r"\b" + re.escape(k), line):
            counts[cur][k] += 1
    if re.search(r"/\*[0-9a-f]{4}\*/", line):
        counts[cur]["instructions"] += 1
print(f"# {lib}: SASS mnemonic counts per kernel (cuobjdump -sass); UTCHMMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st,")
print("# UTMALDG = cp.async.bulk.tensor (TMA load), UTCBAR = tcgen05.commit, SYNCS = mbarrier ops, HMMA = legacy mma.sync (none expected)")
print(" ".join(f"{k:>8s}" for k in keys) + "   instr  kernel")
for fn, c in counts.items():
    name = re.sub(r"\(.*", "", demangle(fn).replace("(anonymous namespace)::", ""))
    print(" ".join(f"{c[k]:8d}" for k in keys) + f" {c['instructions']:7d}  {name}")
