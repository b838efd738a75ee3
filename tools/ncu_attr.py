#!/usr/bin/env python
"""Attribute the PC samples / executed instructions of an `ncu --set full --import-source on` capture to source lines
and device functions: joins the SASS source page (ncu -i X.ncu-rep --page source --csv) with the line table of
`nvdisasm -g -c` on the cubin of the SAME build.

    python tools/ncu_attr.py src.csv dis.txt <kernel-mangled-substring> [top_n]
"""
import csv
import re
import sys
from collections import defaultdict

src_csv, dis, kern = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
# offset -> (file, line) from nvdisasm
loc = {}
cur = None
inside = False
for line in open(dis):
    if line.startswith(".text."):
        inside = kern in line
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(\S.*?);", line)
    if m:
        loc[int(m.group(1), 16)] = (cur, m.group(2))
rows = list(csv.reader(open(src_csv)))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
H = rows[hdr]
ia, isrc, isamp, iexec = H.index("Address"), H.index("Source"), H.index("# Samples"), H.index("Instructions Executed")
stall_cols = [i for i, h in enumerate(H) if h.startswith("stall_") and "Not Issued" not in h]
base = None
by_line = defaultdict(lambda: [0, 0])
stall_line = defaultdict(lambda: defaultdict(int))
opc = defaultdict(int)
tot_s = tot_e = 0
for r in rows[hdr + 1:]:
    if len(r) <= iexec or not r[ia]:
        continue
    a = int(r[ia], 16) if r[ia].startswith("0x") else int(r[ia])
    if base is None:
        base = a
    off = a - base
    s, e = int(r[isamp] or 0), int(r[iexec] or 0)
    l = loc.get(off, ((None, 0), ""))[0] or ("?", 0)
    by_line[l][0] += s
    by_line[l][1] += e
    for i in stall_cols:
        v = int(r[i] or 0)
        if v:
            stall_line[l][H[i]] += v
    opc[r[isrc].split()[0].split(".")[0] if r[isrc] else "?"] += e
    tot_s += s
    tot_e += e
print(f"samples {tot_s}  executed warp-instructions {tot_e}")
print("--- top source lines by samples")
for l, (s, e) in sorted(by_line.items(), key=lambda kv: -kv[1][0])[:top]:
    st = sorted(stall_line[l].items(), key=lambda kv: -kv[1])[:3]
    print(f"{100*s/tot_s:5.1f}% samples {100*e/tot_e:5.1f}% exec  {l[0]}:{l[1]}   " + " ".join(f"{k[6:]}={100*v/max(s,1):.0f}%" for k, v in st))
print("--- opcode mix (executed)")
for k, v in sorted(opc.items(), key=lambda kv: -kv[1])[:25]:
    print(f"{100*v/tot_e:5.1f}%  {k}")
