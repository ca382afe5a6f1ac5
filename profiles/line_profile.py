#!/usr/bin/env python
"""Join an ncu per-SASS-instruction CSV (ncu -i X.ncu-rep --page source --csv) with nvdisasm -g
line info of the same cubin, and print warp-instructions executed per CUDA source line.

usage: line_profile.py <src.csv> <dis.txt> <mangled-substring> <demangled-substring> [top] [source.cu]
"""
import csv
import re
import sys
from collections import defaultdict

src_csv, dis_txt, kname, dname = sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4]
top = int(sys.argv[5]) if len(sys.argv) > 5 else 60

# --- nvdisasm: ordered list of (line, sass-text) for the kernel -------------------------------
lines, cur, infn = [], None, False
stack_line = None
for ln in open(dis_txt, errors="replace"):
    if ln.startswith(".text."):
        infn = kname in ln
        continue
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', ln)
    if m:
        # attribute inlined code to the line in the outermost file position given
        cur = (int(m.group(2)), int(m.group(4)) if m.group(4) else None)
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m:
        lines.append((cur, m.group(2)))

# --- ncu: ordered per-instruction counters --------------------------------------------------------
rows = list(csv.reader(open(src_csv)))
start = None
for i, r in enumerate(rows):
    if r and r[0] == "Kernel Name" and dname in r[1]:
        start = i
        break
hdr = rows[start + 1]
ci, cs, ct = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Thread Instructions Executed")
inst = []
for r in rows[start + 2:]:
    if not r or r[0] == "Kernel Name":
        break
    inst.append((r[1].strip(), int(r[ci] or 0), int(r[cs] or 0), int(r[ct] or 0)))
n = min(len(inst), len(lines))
per = defaultdict(lambda: [0, 0, 0])
tot = [0, 0, 0]
for k in range(n):
    (ln, outer), _ = lines[k] if lines[k][0] else ((0, None), None)
    key = ln
    for j in range(3):
        per[key][j] += inst[k][1 + j]
        tot[j] += inst[k][1 + j]
# --- phase table: bucket by the outermost line (call site of inlined helpers) ----------------
if len(sys.argv) > 6:
    marks = []
    for i, t in enumerate(open(sys.argv[6]).read().splitlines(), 1):
        m = re.search(r"// ---- (phase [0-9a-z]+:[^-]*)", t)
        if m:
            marks.append((i, m.group(1).strip()))
    if marks:
        ph = defaultdict(lambda: [0, 0, 0])
        for k in range(n):
            pos = lines[k][0] or (0, None)
            ln = pos[1] or pos[0]
            name = "(before phase 0 / helpers)"
            for (a, nm) in marks:
                if ln >= a:
                    name = nm
            if ln >= 846 or ln < marks[0][0]:
                name = "(outside step body: ln %d..)" % (ln // 1000 * 1000)
            for j in range(3):
                ph[name][j] += inst[k][1 + j]
        T = sum(v[0] for v in ph.values())
        S = sum(v[1] for v in ph.values())
        print("%-70s %10s %6s %8s %6s %s" % ("phase", "warp-inst", "%", "samples", "%", "thr/inst"))
        for nm, v in sorted(ph.items(), key=lambda kv: -kv[1][0]):
            print("%-70s %10d %5.1f%% %8d %5.1f%% %5.1f" % (nm[:70], v[0], 100.0 * v[0] / T, v[1], 100.0 * v[1] / max(S, 1), v[2] / max(v[0], 1)))
        print()
print("sass instructions: ncu %d, nvdisasm %d; total warp-inst %d, samples %d, thread-inst %d" %
      (len(inst), len(lines), tot[0], tot[1], tot[2]))
src = open(sys.argv[6]).read().splitlines() if len(sys.argv) > 6 else None
for key, v in sorted(per.items(), key=lambda kv: -kv[1][0])[:top]:
    text = src[key - 1].strip()[:90] if src and 0 < key <= len(src) else ""
    print("%5d  inst %9d (%4.1f%%)  samples %6d (%4.1f%%)  thr/inst %4.1f  %s" %
          (key, v[0], 100.0 * v[0] / tot[0], v[1], 100.0 * v[1] / max(tot[1], 1), v[2] / max(v[0], 1), text))
