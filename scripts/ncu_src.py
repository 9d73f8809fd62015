"""SASS-level digest of one kernel of an ncu report: opcode histogram (executed warp instructions) and the top stall sites.
usage: python scripts/ncu_src.py rep.ncu-rep kernel-regex [topN]"""
import csv, io, subprocess, sys, collections
rep, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + pat], capture_output=True, text=True).stdout
lines = txt.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
rows = list(csv.reader(io.StringIO("\n".join(lines[start:]))))
hdr = rows[0]; ix = {h: i for i, h in enumerate(hdr)}
ops = collections.Counter(); samp = collections.Counter()
tot_i = tot_s = 0
recs = []
for r in rows[1:]:
    if len(r) < len(hdr) or r[0] == "Address": break
    src = r[ix["Source"]].strip()
    parts = src.split()
    op = parts[1] if parts and parts[0].startswith("@") else (parts[0] if parts else "?")
    n = int(r[ix["Instructions Executed"]] or 0); s = int(r[ix["# Samples"]] or 0)
    ops[op.split(".")[0]] += n; samp[op.split(".")[0]] += s; tot_i += n; tot_s += s
    recs.append((s, n, src, r))
print(f"total warp instructions {tot_i:,}  samples {tot_s:,}")
print("opcode histogram (share of executed instructions | share of samples):")
for op, n in ops.most_common(28):
    print(f"  {op:12s} {100*n/tot_i:5.1f}%  | {100*samp[op]/max(tot_s,1):5.1f}%")
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
print("top stall sites:")
for s, n, src, r in sorted(recs, key=lambda t: -t[0])[:top]:
    st = sorted(((int(r[ix[c]] or 0), c[6:]) for c in stall_cols), reverse=True)[:3]
    print(f"  {100*s/max(tot_s,1):5.2f}%  exec {n:>10,}  {src[:70]:70s} " + " ".join(f"{c}={v}" for v, c in st if v))
