"""Opcode histogram of the innermost HMMA loop of every kernel whose name matches argv[1] (default: attn)."""
import re, collections, subprocess, sys
pat = sys.argv[1] if len(sys.argv) > 1 else "attn"
txt = subprocess.run(["cuobjdump", "-sass", "fudanocr_b200/libfocr_sm100.so"], capture_output=True, text=True).stdout
lines = txt.split("\n")
funcs = [(i, l.split("Function : ")[1]) for i, l in enumerate(lines) if "Function :" in l] + [(len(lines), "")]
for (a, name), (b, _) in zip(funcs, funcs[1:]):
    if pat not in name: continue
    ins = []
    for l in lines[a:b]:
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", l)
        if m: ins.append((int(m.group(1), 16), m.group(2).strip()))
    best = None
    for addr, t in ins:
        m = re.search(r"BRA\s+0x([0-9a-f]+)", t)
        if m and int(m.group(1), 16) < addr:
            body = [x for x in ins if int(m.group(1), 16) <= x[0] <= addr]
            if sum("HMMA" in x[1] for x in body) >= 16 and (best is None or len(body) < len(best)): best = body
    short = re.search(r"\d+([a-z_0-9]+_kernelI\w+?)EEv", name)
    short = short.group(1) if short else name[:60]
    if best is None: print(short, "no loop"); continue
    c = collections.Counter(re.sub(r"^@!?U?P\d\s+", "", t).split()[0].split(".")[0] for _, t in best)
    print(short, len(best), dict(c.most_common(20)))
    if len(sys.argv) > 2:
        for _, t in best: print("     ", t)
