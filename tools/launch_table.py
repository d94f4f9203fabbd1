"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel:
python tools/launch_table.py gpurun_out/x_launches.csv [steps_in_window]"""
import csv, re, sys
from collections import defaultdict
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr = rows[0]; ix = {h: i for i, h in enumerate(hdr)}
agg = defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    if r[ix["Metric Name"]] != "gpu__time_duration.sum": continue
    name = r[ix["Kernel Name"]]
    name = re.sub(r"\(.*", "", name)                     # drop the argument list
    name = re.sub(r"(stgcn::)?(simt|umma)::", "", name)
    grid = r[ix["Grid Size"]]
    v = float(r[ix["Metric Value"]].replace(",", ""))
    u = r[ix["Metric Unit"]]
    v *= {"ns": 1e-3, "us": 1, "ms": 1e3}.get(u, 1e-3)
    a = agg[name]; a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
print(f"total {tot:.0f} us over {sum(a[0] for a in agg.values())} launches\n")
print("| kernel | launches | time (us) | share |\n|---|---|---|---|")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{k[:90]}` | {n} | {t:.0f} | {100 * t / tot:.1f} % |")
