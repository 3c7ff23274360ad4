"""Summarise an .ncu-rep (raw page) into the handful of numbers quoted in DESIGN.md / bench.py: duration, DRAM bytes,
tensor-pipe and XU utilisation, issue activity, registers, shared memory, top stall reasons."""
import csv, subprocess, sys, json
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
want = {
 "gpu__time_duration.sum": "duration", "dram__bytes_read.sum": "dram_read", "dram__bytes_write.sum": "dram_write",
 "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_pct_active",
 "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active": "xu_pipe_pct",
 "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
 "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
 "launch__registers_per_thread": "regs_per_thread", "launch__shared_mem_per_block_dynamic": "dyn_smem",
 "launch__grid_size": "grid", "launch__block_size": "block", "sm__cycles_elapsed.avg": "sm_cycles",
 "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct_of_peak",
 "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
}
allk = []
for r in rows[2:]:
    d = {"kernel": r[hdr.index("Kernel Name")][:90]}
    stalls = {}
    for i, h in enumerate(hdr):
        if h in want:
            d[want[h]] = f"{r[i]} {units[i]}"
        if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio"):
            try:
                stalls[h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]] = float(r[i].replace(",", ""))
            except ValueError:
                pass
    d["top_stalls_per_issue"] = dict(sorted(stalls.items(), key=lambda kv: -kv[1])[:5])
    allk.append(d)
print(json.dumps(allk[0] if len(allk) == 1 else allk, indent=1))
