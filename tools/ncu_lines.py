#!/usr/bin/env python
"""Attribute an ncu SASS-level source page to CUDA source lines.

  ncu -i rep.ncu-rep --page source --csv --kernel-name regex:K > sass.csv
  cuobjdump -xelf all libppp_gpu.so ; nvdisasm -g -c knn.sm_100a.cubin > knn.dis
  python tools/ncu_lines.py sass.csv knn.dis K [file.cu]

nvdisasm -g interleaves `//## File "...", line N` markers with the instructions; the ncu page lists
the same instructions in the same order (checked by opcode), so instruction i of one is instruction
i of the other.  Prints executed warp instructions and stall samples per source line.
"""
import csv, re, sys
from collections import defaultdict

sass_csv, dis, kname = sys.argv[1:4]
only = sys.argv[4] if len(sys.argv) > 4 else None
rows = list(csv.reader(open(sass_csv)))
hi = next(i for i, r in enumerate(rows) if "Instructions Executed" in r)
hdr = rows[hi]
body = []
for r in rows[hi + 1:]:          # first profiled launch only
    if r and r[0] in ("Kernel Name", "Address"):
        break
    body.append(r)
ci, cs, csamp = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("# Samples")
ins = [(r[cs].split()[0] if not r[cs].strip().startswith("@") else r[cs].split()[1], int(r[ci]), int(r[csamp]))
       for r in body if len(r) > ci]
lines = open(dis).read().split("\n")
start = next(i for i, l in enumerate(lines) if l.startswith("_Z") and kname in l and l.rstrip().endswith(":"))
cur = ("?", 0)
seq = []
for l in lines[start + 1:]:
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(.*?);", l)
    if m:
        t = m.group(1).split()
        op = t[1] if t[0].startswith("@") else t[0]
        seq.append((op, cur))
    if l.startswith("\t.section") or l.startswith("//-----"):
        if seq:
            break
n = min(len(seq), len(ins))
bad = sum(1 for i in range(n) if seq[i][0].split(".")[0] != ins[i][0].split(".")[0])
print(f"# {len(ins)} profiled instructions, {len(seq)} disassembled, opcode mismatches in first {n}: {bad}")
agg = defaultdict(lambda: [0, 0, 0])
for i in range(n):
    a = agg[seq[i][1]]
    a[0] += ins[i][1]; a[1] += ins[i][2]; a[2] += 1
tot_i = sum(a[0] for a in agg.values()); tot_s = sum(a[1] for a in agg.values())
print(f"# total warp instructions {tot_i}, samples {tot_s}")
src = {}
for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    if only and f != only:
        continue
    if a[1] < tot_s * 0.004 and a[0] < tot_i * 0.004:
        continue
    if f not in src:
        try:
            src[f] = open(f"polishpathplanning_b200/csrc/{f}").read().split("\n")
        except OSError:
            src[f] = []
    text = src[f][ln - 1].strip()[:90] if 0 < ln <= len(src[f]) else ""
    print(f"{f}:{ln:5d}  inst {100*a[0]/tot_i:5.1f}%  samples {100*a[1]/tot_s:5.1f}%  sass {a[2]:4d}  | {text}")
