#!/usr/bin/env python
"""Executed warp-instructions per source region of the N=64 Flock step kernel.

usage: line_buckets.py <ncu source-page csv> <nvdisasm -g -c output> [macm_kernels.cu]
Regions are cut at the device functions and at the `// ---- phase` markers of the source file, so the
table follows the file as it changes.  ALU = compare/select/integer pipe, FMA = fp32/IMAD pipes."""
import csv
import os
import re
import sys
from collections import defaultdict

src_csv, dis_txt = sys.argv[1:3]
cu = sys.argv[3] if len(sys.argv) > 3 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                        "gym-macm_b200", "csrc", "macm_kernels.cu")
kname = sys.argv[4] if len(sys.argv) > 4 else "macm_step_kernelILi32ELi2ELi0ELi0"
dname = sys.argv[5] if len(sys.argv) > 5 else "macm_step_kernel<(int)32, (int)2, (int)0, (int)0>"
E = int(sys.argv[6]) if len(sys.argv) > 6 else 4096

marks = []   # (first line, name)
for i, t in enumerate(open(cu).read().splitlines(), 1):
    m = re.match(r"\s*// ---- (phase [0-9a-z]+)[: ]", t)
    if m:
        marks.append((i, m.group(1)))
        continue
    m = re.match(r"__device__ (?:__forceinline__ |__noinline__ )?[\w:<>\*& ]+?\b(\w+)\(", t)
    if m:
        marks.append((i, m.group(1)))
    m = re.match(r"__global__ void .*?(\w+)\(", t)
    if m:
        marks.append((i, m.group(1) + " (prologue)"))
marks.sort()


def region(ln):
    name = "(top of file)"
    for first, nm in marks:
        if first <= ln:
            name = nm
        else:
            break
    return name


lines, cur, infn = [], None, False
for ln in open(dis_txt, errors="replace"):
    if ln.startswith(".text."):
        infn = kname in ln
        continue
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = int(m.group(2))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m:
        lines.append((cur, m.group(2)))
rows = list(csv.reader(open(src_csv)))
start = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name" and dname in r[1]][0]
hdr = rows[start + 1]
ci, cs, ct = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Thread Instructions Executed")
inst = []
for r in rows[start + 2:]:
    if not r or r[0] == "Kernel Name":
        break
    inst.append((r[1].strip(), int(r[ci] or 0), int(r[cs] or 0), int(r[ct] or 0)))
ALU = ("FSETP", "ISETP", "FSEL", "SEL", "VIADD", "LOP3", "IADD3", "SHF", "LEA", "MOV", "FMNMX", "FMNMX3", "PRMT", "VIMNMX",
       "VIMNMX3", "PLOP3", "IABS", "BMSK", "SGXT", "FCHK", "VOTE", "P2R", "R2P", "CS2R", "POPC", "FLO", "BREV")
FMA = ("FADD", "FMUL", "FFMA", "IMAD", "FADD2", "FMUL2", "FFMA2", "HFMA2", "HADD2", "HMUL2")
per, tot = defaultdict(lambda: [0, 0, 0, 0, 0]), [0, 0, 0, 0, 0]
for k in range(min(len(inst), len(lines))):
    name = region(lines[k][0] or 0)
    s = re.sub(r"^@!?U?P\d+\s+", "", inst[k][0])
    op = s.split()[0].split(".")[0]
    v = [inst[k][1], inst[k][2], inst[k][3], inst[k][1] if op in ALU else 0, inst[k][1] if op in FMA else 0]
    for j in range(5):
        per[name][j] += v[j]
        tot[j] += v[j]
print("total warp-inst %d (%.0f/env) samples %d thr/inst %.1f  ALU %.0f/env FMA %.0f/env" % (
    tot[0], tot[0] / E, tot[1], tot[2] / max(tot[0], 1), tot[3] / E, tot[4] / E))
for nm, v in sorted(per.items(), key=lambda x: -x[1][0]):
    if v[0] == 0:
        continue
    print("%-28s inst/env %6.0f %5.1f%%  samples %5.1f%%  thr/inst %4.1f  alu/env %5.0f fma/env %5.0f" % (
        nm, v[0] / E, 100 * v[0] / tot[0], 100 * v[1] / max(tot[1], 1), v[2] / max(v[0], 1), v[3] / E, v[4] / E))
