#!/usr/bin/env python3
"""Summarise one kernel of an .ncu-rep (raw page) into the handful of metrics the design cites.
Usage: python tools/ncu_summary.py REPORT.ncu-rep [kernel-substring [scenarios-per-gpu]] > profiles/xxx.json"""
import csv, json, subprocess, sys
rep = sys.argv[1]
want = sys.argv[2] if len(sys.argv) > 2 else ""
batch = int(sys.argv[3]) if len(sys.argv) > 3 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_uniform.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__registers_per_thread", "launch__block_size",
        "launch__grid_size", "launch__shared_mem_per_block_dynamic", "smsp__cycles_active.avg", "sm__cycles_elapsed.max"]
out = []
for r in rows[2:]:
    d = dict(zip(hdr, r))
    if want not in d.get("Kernel Name", ""):
        continue
    rec = {"kernel": d["Kernel Name"][:120]}
    if batch is not None:
        rec["scenarios_per_gpu"] = batch
    for k, u in zip(hdr, units):
        if k in KEYS or k.startswith("smsp__average_warps_issue_stalled") and k.endswith("per_issue_active.ratio") or "tensor" in k and "pct" in k and ".avg." in k:
            try:
                value = float(d[k].replace(",", ""))
                rec[k + (f" [{u}]" if u else "")] = value if value == value else None
            except ValueError:
                rec[k] = d[k]
    if rec.get("gpu__time_duration.sum [ms]") is not None and not any(v is None for v in rec.values()):
        out.append(rec)
json.dump(out, sys.stdout, indent=1)
