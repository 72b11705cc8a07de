"""DRAM traffic per launch (dram__bytes_read.sum + dram__bytes_write.sum) of the kernels in an `ncu --set full` report
-> profiles/ncu_traffic.json, which bench.py quotes as roofline.traffic.
usage: python tools/ncu_traffic.py report.ncu-rep "config string" """
import csv, io, json, subprocess, sys
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
out = {}
for r in rows[2:]:
    name = r[idx["Kernel Name"]]
    key = None
    for k, pat in [("ce_fwd_kernel", "ce_fwd_kernel"), ("ce_bwd2_kernel_dV", "ce_bwd2_kernel<128, 1>"), ("ce_bwd2_kernel_dU", "ce_bwd2_kernel<128, 0>"),
                   ("ce_bwd3_kernel_dU", "ce_bwd3_kernel<128, 1>"), ("ce_bwd3_kernel_dV", "ce_bwd3_kernel<128, 0>"),
                   ("ce_bwd3x_kernel_dU", "ce_bwd3x_kernel<256, 1>"), ("ce_bwd3x_kernel_dV", "ce_bwd3x_kernel<256, 0>"),
                   ("tower_fwd_kernel", "tower_fwd_kernel"), ("ce_bwd_reduce_kernel", "ce_bwd_reduce_kernel"),
                   ("ce_combine_loss_kernel", "ce_combine_loss_kernel"), ("adam_kernel", "adam_kernel")]:
        if pat in name:
            key = k
    if key is None or key in out:
        continue
    tot = 0.0
    for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        tot += float(r[idx[m]].replace(",", "")) * scale[units[idx[m]]]
    out[key] = tot
json.dump({"config": sys.argv[2], "source": sys.argv[1].split("/")[-1], "bytes_per_launch": out},
          open("profiles/ncu_traffic.json", "w"), indent=1)
print(json.dumps(out, indent=1))
