"""Compact per-kernel digest of an `ncu --set full` report: duration, issue rate, pipe loads, stall reasons.
usage: python scripts/ncu_digest.py gpurun_out/x.ncu-rep [name-filter]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; filt = sys.argv[2] if len(sys.argv) > 2 else ""
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr = rows[0]; idx = {h: i for i, h in enumerate(hdr)}
def f(r, h):
    try: return float(r[idx[h]].replace(",", ""))
    except Exception: return float("nan")
keys = [("dur_us", "gpu__time_duration.sum"), ("regs", "launch__registers_per_thread"),
        ("issue%", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
        ("ipc", "smsp__inst_executed.avg.per_cycle_active"),
        ("inst", "smsp__inst_executed.sum"),
        ("alu%", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
        ("fma%", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
        ("xu%", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
        ("lsu%", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
        ("tensor%", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
        ("hmma%", "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active"),
        ("dram%", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        ("l2%", "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
        ("l1%", "l1tex__throughput.avg.pct_of_peak_sustained_active"),
        ("warps/sched", "smsp__warps_active.avg.per_cycle_active"),
        ("eligible/sched", "smsp__warps_eligible.avg.per_cycle_active")]
stall = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
if not stall:
    stall = [h for h in hdr if "issue_stalled" in h and h.endswith(".pct")]
for r in rows[2:]:
    name = r[idx["Kernel Name"]]
    if filt and filt not in name: continue
    print("==", name[:90])
    print("   " + "  ".join(f"{k}={f(r, h):.4g}" for k, h in keys if h in idx))
    st = sorted(((f(r, h), h) for h in stall), reverse=True)
    print("   stalls: " + ", ".join(f"{h.split('issue_stalled_')[1].split('_per_')[0].split('.')[0]}={v:.2f}" for v, h in st[:7] if v == v))
