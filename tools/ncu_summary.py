"""Table of the launches in an .ncu-rep (run where ncu is installed; no GPU needed):
python tools/ncu_summary.py gpurun_out/x.ncu-rep"""
import csv, io, subprocess, sys
M = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
     "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
     "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__grid_size", "launch__registers_per_thread",
     "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv", "--metrics", ",".join(M)],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
def f(r, k):
    if k not in ix: return float("nan")
    try: return float(r[ix[k]].replace(",", ""))
    except Exception: return float("nan")
def scale(k, v):   # normalise to us / MB
    if k not in ix: return float("nan")
    u = units[ix[k]]
    return v * {"ns": 1e-3, "us": 1, "ms": 1e3, "byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1, "Gbyte": 1e3}.get(u, 1)
print("| id | kernel | grid | regs | time us | DRAM rd MB | DRAM wr MB | DRAM GB/s | L2 MB | tensor % | issue % | warps % |")
print("|---|---|---|---|---|---|---|---|---|---|---|---|")
for r in rows[2:]:
    t = scale(M[0], f(r, M[0])); rd = scale(M[1], f(r, M[1])); wr = scale(M[2], f(r, M[2])); l2 = scale(M[9], f(r, M[9]))
    print(f"| {r[ix['ID']]} | {r[ix['Kernel Name']][:70]} | {r[ix[M[6]]]} | {r[ix[M[7]]]} | {t:.1f} | {rd:.1f} | {wr:.1f} | "
          f"{(rd + wr) / t * 1e3:.0f} | {l2:.0f} | {f(r, M[3]):.1f} | {f(r, M[5]):.1f} | {f(r, M[4]):.1f} |")
