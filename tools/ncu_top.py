"""Read an .ncu-rep here (no GPU): headline metrics + the instructions with the most stall samples."""
import csv, subprocess, sys
rep = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[0]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "launch__grid_size",
        "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active"]
for r in rows[2:]:
    for w in want:
        if w in hdr:
            print(f"  {w} = {r[hdr.index(w)]} {rows[1][hdr.index(w)]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
s = starts[0]
e = starts[1] if len(starts) > 1 else len(rows)
hdr = rows[s + 1]
body = rows[s + 2:e]
iS, iSrc, iI = hdr.index("# Samples"), hdr.index("Source"), hdr.index("Instructions Executed")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[iS]) for r in body)
print("total samples", tot, "instructions", len(body))
agg = {}
for r in body:
    for i in stall_cols:
        agg[hdr[i]] = agg.get(hdr[i], 0) + int(r[i])
print("stall totals:", sorted(agg.items(), key=lambda kv: -kv[1])[:8])
idx = {id(r): k for k, r in enumerate(body)}
for r in sorted(body, key=lambda r: -int(r[iS]))[:n]:
    st = sorted([(int(r[i]), hdr[i]) for i in stall_cols], reverse=True)[:2]
    print(f"{r[iS]:>7} exec={r[iI]:>10} #{idx[id(r)]:<5} {r[iSrc].strip()[:80]:<80} {st}")
