"""Instruction histogram of the built library per kernel family: the SASS evidence that the bf16 path runs on tcgen05 /
TMA / TMEM (UTCHMMA = tcgen05.mma, UTCBAR = tcgen05.commit, LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG = TMA tensor
load / store, UBLKCP = bulk copy, SYNCS = mbarrier, NANOSLEEP = hinted try_wait, LDGSTS = cp.async).  No GPU needed.

    python tools/sass_histogram.py > profiles/r02_sass_histogram.md
"""
import collections, os, re, subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "stgcn_b200", "lib", "libstgcn_b200.so")
KEYS = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "LDGSTS", "SYNCS", "NANOSLEEP", "MUFU", "ATOMG", "RED"]

out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
fam = collections.OrderedDict()
cur = None
for line in out.splitlines():
    if "Function :" in line:
        name = subprocess.run(["c++filt", line.split("Function :")[1].strip()], capture_output=True, text=True).stdout.strip()
        m = re.search(r"(umma_\w+_kernel|\w+_kernel)", name)
        cur = m.group(1) if m else name[:40]
        fam.setdefault(cur, {"variants": 0, "instr": 0, **{k: 0 for k in KEYS}})
        fam[cur]["variants"] += 1
    elif cur and re.match(r"^\s+/\*[0-9a-f]{4,5}\*/", line):
        fam[cur]["instr"] += 1
        op = re.sub(r"^\s+/\*[0-9a-f]+\*/\s+(@!?U?P\d+\s+)?", "", line).split()[0].rstrip(";")
        for k in KEYS:
            if op.startswith(k):
                fam[cur][k] += 1
print("| kernel family | variants | SASS instructions | " + " | ".join(KEYS) + " |")
print("|---|---|---|" + "---|" * len(KEYS))
for k, v in sorted(fam.items(), key=lambda kv: -kv[1]["UTCHMMA"] * 100000 - kv[1]["instr"]):
    if v["instr"] < 50:
        continue
    print(f"| `{k}` | {v['variants']} | {v['instr']} | " + " | ".join(str(v[x]) for x in KEYS) + " |")
