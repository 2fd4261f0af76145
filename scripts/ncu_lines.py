#!/usr/bin/env python
"""Per-source-line profile from an ncu report: joins `ncu --page source --csv` (SASS rows with executed
instruction counts and stall samples) with `nvdisasm -g` line info of the same kernel, by instruction order.

    python scripts/ncu_lines.py REPORT.ncu-rep KERNEL_REGEX CUBIN MANGLED_SUBSTR [top_n] [NAME_SUBSTR]

KERNEL_REGEX matches the base function name only (ncu); NAME_SUBSTR picks the instance whose full name (template
arguments included, e.g. "(int)1, (int)1, (int)4") contains it.
"""
import csv, re, subprocess, sys, collections

rep, kre, cubin, mangled = sys.argv[1:5]
topn = int(sys.argv[5]) if len(sys.argv) > 5 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kre}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
# several kernel instances may be concatenated: keep the first table
tables, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "rows": []}; tables.append(cur); continue
    if cur is None: continue
    if cur["hdr"] is None: cur["hdr"] = r; continue
    cur["rows"].append(r)
want = sys.argv[6] if len(sys.argv) > 6 else ""
t = next(x for x in tables if want in x["name"])
h = {n: i for i, n in enumerate(t["hdr"])}
sass = subprocess.run(["nvdisasm", "-g", cubin], capture_output=True, text=True).stdout.splitlines()
# locate function
start = next(i for i, l in enumerate(sass) if l.startswith(".text.") and mangled in l)
lines = []  # (file,line) per instruction
curline = ("?", 0)
ins_re = re.compile(r"^\s+/\*([0-9a-f]{4,})\*/\s+(.*?);")
for l in sass[start + 1:]:
    if l.startswith(".text.") or l.startswith("\t.section") or l.startswith(".section"): 
        if lines: break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        curline = (m.group(1).split("/")[-1], int(m.group(2))); continue
    m = ins_re.match(l)
    if m: lines.append((curline, m.group(2)))
n = min(len(lines), len(t["rows"]))
print(f"kernel: {t['name'][:80]}  sass rows ncu={len(t['rows'])} nvdisasm={len(lines)}")
mism = sum(1 for i in range(n) if t["rows"][i][h["Source"]].split()[0].strip("@!P0123456789 ") [:3] != lines[i][1].split()[0].strip("@!P0123456789 ")[:3])
print("opcode mismatches:", mism)
agg = collections.defaultdict(lambda: [0, 0, collections.Counter()])
tot_i = tot_s = 0
stall_cols = [c for c in t["hdr"] if c.startswith("stall_") and "Not Issued" not in c]
for i in range(n):
    r = t["rows"][i]
    ie = int(float(r[h["Instructions Executed"]] or 0)); ss = int(float(r[h["# Samples"]] or 0))
    a = agg[lines[i][0]]; a[0] += ie; a[1] += ss; tot_i += ie; tot_s += ss
    for c in stall_cols:
        v = int(float(r[h[c]] or 0))
        if v: a[2][c[6:]] += v
print(f"total warp-instructions {tot_i}, samples {tot_s}")
print(f"{'file:line':34s} {'inst%':>6s} {'samp%':>6s}  top stalls")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:topn]:
    st = ", ".join(f"{n}:{v}" for n, v in a[2].most_common(3))
    print(f"{k[0]+':'+str(k[1]):34s} {100*a[0]/max(tot_i,1):6.2f} {100*a[1]/max(tot_s,1):6.2f}  {st}")
