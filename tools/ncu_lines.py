#!/usr/bin/env python
"""Per-source-line instruction counts and stall samples of one kernel in an .ncu-rep.

  tools/ncu_lines.py file.ncu-rep <kernel-regex> <cubin> <mangled-substring> [top]

ncu's CSV source page is per SASS instruction; `nvdisasm -g` of the same cubin (extract with
`cuobjdump -xelf all lib.so`) carries the file/line of each instruction. The two listings are joined
by instruction offset. Inlined device functions are attributed to the line of the innermost frame."""
import collections
import csv
import io
import re
import subprocess
import sys

rep, kre, cubin, mangled = sys.argv[1:5]
top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]
ia, isrc, isamp, iex = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
ins = []
base = None
for r in rows[2:]:
    if len(r) <= iex or not r[ia].startswith("0x"):
        if ins:
            break      # next kernel instance
        continue
    a = int(r[ia], 16)
    if base is None:
        base = a
    ins.append((a - base, r[isrc].strip(), int(r[isamp] or 0), int(r[iex] or 0), {h: int(r[i] or 0) for i, h in stall_cols}))

dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
line_of = {}
cur = None
inside = False
for ln in dis:
    if ln.startswith(".text.") and ln.rstrip().endswith(":"):
        inside = mangled in ln
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        if "inlined at" in ln and cur is not None and False:
            pass
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]+)\*/", ln)
    if m:
        line_of[int(m.group(1), 16)] = cur

by_line = collections.defaultdict(lambda: [0, 0, collections.Counter(), collections.Counter()])
tot_s = tot_e = 0
for off, sass, s, e, st in ins:
    key = line_of.get(off, ("?", 0))
    b = by_line[key]
    b[0] += s; b[1] += e
    b[2].update(st)
    toks = sass.split()
    op = (toks[1] if toks[0].startswith("@") else toks[0]).split(".")[0]
    b[3][op] += e
    tot_s += s; tot_e += e
print(f"instructions {tot_e}  samples {tot_s}  ({len(ins)} SASS instructions, {len(line_of)} mapped)")
print("by executed instructions:")
for key, b in sorted(by_line.items(), key=lambda kv: -kv[1][1])[:top]:
    ops = " ".join(f"{o}:{100*c/max(b[1],1):.0f}" for o, c in b[3].most_common(4))
    st = " ".join(f"{h[6:]}:{c}" for h, c in b[2].most_common(3) if c)
    print(f"  {key[0]}:{key[1]:<5d} exec {100*b[1]/tot_e:5.1f}%  samples {100*b[0]/max(tot_s,1):5.1f}%  [{ops}]  {st}")
print("by stall samples:")
for key, b in sorted(by_line.items(), key=lambda kv: -kv[1][0])[:top//2]:
    st = " ".join(f"{h[6:]}:{c}" for h, c in b[2].most_common(4) if c)
    print(f"  {key[0]}:{key[1]:<5d} samples {100*b[0]/max(tot_s,1):5.1f}%  exec {100*b[1]/tot_e:5.1f}%  {st}")

# optional phase summary: FVG_PHASES="name:file:lo-hi,..." groups lines into ranges
import os
ph = os.environ.get("FVG_PHASES")
if ph:
    print("by phase:")
    for spec in ph.split(","):
        name, fn, rng = spec.split(":")
        lo, hi = (int(x) for x in rng.split("-"))
        s = sum(b[0] for k, b in by_line.items() if k[0] == fn and lo <= k[1] <= hi)
        e = sum(b[1] for k, b in by_line.items() if k[0] == fn and lo <= k[1] <= hi)
        st = collections.Counter()
        for k, b in by_line.items():
            if k[0] == fn and lo <= k[1] <= hi:
                st.update(b[2])
        print(f"  {name:12s} samples {100*s/max(tot_s,1):5.1f}%  exec {100*e/tot_e:5.1f}%  " + " ".join(f"{h[6:]}:{100*c/max(tot_s,1):.1f}%" for h, c in st.most_common(5)))
