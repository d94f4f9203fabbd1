"""Hottest SASS instructions (by warp-stall samples) of one launch in an .ncu-rep:
python tools/ncu_hot.py rep launch_index [top]"""
import csv, io, subprocess, sys
rep, idx = sys.argv[1], int(sys.argv[2])
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", str(idx), "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
print(rows[0][1][:100])
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
body = [r for r in rows[2:] if len(r) == len(hdr) and r[ix["# Samples"]].isdigit()]
tot = sum(int(r[ix["# Samples"]]) for r in body)
agg = {s: sum(int(r[ix[s]]) for r in body) for s in stalls}
print("total samples", tot, "| by reason:", ", ".join(f"{k[6:]}={v * 100 // max(tot, 1)}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
body_i = list(enumerate(body))
for n, r in sorted(body_i, key=lambda t: -int(t[1][ix["# Samples"]]))[:top]:
    s = int(r[ix["# Samples"]])
    why = sorted(((int(r[ix[k]]), k[6:]) for k in stalls), reverse=True)[:2]
    print(f"{n:5d} {s * 100 / max(tot, 1):5.1f}%  {r[ix['Source']].strip()[:70]:70s} exec={r[ix['Instructions Executed']]:>8} {why[0][1]}:{why[0][0]} {why[1][1]}:{why[1][0]}")
