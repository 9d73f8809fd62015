"""Turn the raw gpurun artefacts (ncu launch list CSV, ncu --set full report, bench JSON) into the small text
summaries committed under profiles/.   python scripts/summarize_profiles.py <round-tag>"""
import collections
import csv
import json
import os
import re
import subprocess
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
out_dir = "profiles"
os.makedirs(out_dir, exist_ok=True)

# 1. launch list -------------------------------------------------------------------------------------------------
src = f"gpurun_out/launches_{tag}.csv"
if os.path.exists(src):
    rows = [r for r in csv.reader(open(src)) if len(r) > 5]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg, tot = collections.defaultdict(lambda: [0, 0.0]), 0.0
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        name = re.sub(r"\(.*", "", r[ki]).replace("void ", "").replace("<unnamed>::", "")
        agg[name][0] += 1
        agg[name][1] += v
        tot += v
    with open(f"{out_dir}/{tag}_launch_list_one_step.txt", "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none, ONE TBSRN train step, B=256 (scripts/one_step.py)\n")
        f.write(f"# serialised cold-cache per-launch times: compare SHARES, not absolutes\n")
        f.write(f"# total {tot / 1e6:.3f} ms over {sum(a[0] for a in agg.values())} launches\n")
        f.write(f"{'ms':>10} {'share':>7} {'count':>6}  kernel\n")
        for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{t / 1e6:10.4f} {100 * t / tot:6.2f}% {c:6d}  {k}\n")
    print("wrote launch list summary")

# 2. full report ---------------------------------------------------------------------------------------------------
rep = f"gpurun_out/hot_{tag}.ncu-rep"
if os.path.exists(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum"]
    traffic = {}
    with open(f"{out_dir}/{tag}_ncu_hot_kernels.txt", "w") as f:
        f.write("# ncu --set full --clock-control none --import-source on, hot kernels at the B=256 shapes (scripts/prof_ops.py)\n")
        for r in rows[2:]:
            name = re.sub(r"\(.*", "", r[idx["Kernel Name"]]).replace("void ", "").replace("<unnamed>::", "")
            f.write(f"\n== {name}\n")
            for w in want:
                if w in idx:
                    f.write(f"   {w:72s} {r[idx[w]]:>16s} {units[idx[w]]}\n")
            try:
                rd = float(r[idx["dram__bytes_read.sum"]].replace(",", ""))
                wr = float(r[idx["dram__bytes_write.sum"]].replace(",", ""))
                scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}
                traffic[name] = rd * scale[units[idx["dram__bytes_read.sum"]]] + wr * scale[units[idx["dram__bytes_write.sum"]]]
            except Exception:
                pass
    json.dump(traffic, open(f"{out_dir}/{tag}_dram_traffic_bytes.json", "w"), indent=1)
    print("wrote ncu hot-kernel summary")

# 3. bench line ------------------------------------------------------------------------------------------------------
for name in ("bench2.json", "bench1.json"):
    p = f"gpurun_out/{name}"
    if os.path.exists(p) and os.path.getsize(p) > 10:
        line = open(p).read().strip().splitlines()[-1]
        json.loads(line)
        open(f"{out_dir}/{tag}_bench.json", "w").write(line + "\n")
        print("wrote bench line")
        break
