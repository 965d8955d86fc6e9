"""Attribute an `ncu --page source --csv` export (SASS view: instructions executed, stall samples per address) to SOURCE
LINES and to named code regions, using the line table of the shipped binary:

    cuobjdump -xelf all vdlm2dec_b200/libvdl2gpu.so            (-> vdl2_kernel.sm_100a.cubin)
    nvdisasm -g -c vdl2_kernel.sm_100a.cubin > sass_lines.txt
    ncu -i prof.ncu-rep --page source --csv > src.csv
    python tools/ncu_line_summary.py src.csv sass_lines.txt <mangled kernel name substring> [regions.json]

The binary must be the one that was profiled (addresses are matched).  Regions: a list of [name, file substring, first line,
last line]; the innermost source line of an instruction decides (inlined callees count for the callee's lines).
"""
import collections
import csv
import json
import re
import sys

src_csv, sass_txt, kname = sys.argv[1:4]
regions = json.load(open(sys.argv[4])) if len(sys.argv) > 4 else []

# address -> (file, line) from nvdisasm -g
addr2line = {}
cur = None
infn = False
for l in open(sass_txt):
    if l.startswith(".text."):
        infn = kname in l
        continue
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*)", l)
    if m:
        addr2line[int(m.group(1), 16)] = cur

rows = list(csv.reader(open(src_csv)))
hdr = rows[1] if rows[0][0] != "Address" and "Address" in rows[1] else rows[0]
start = rows.index(hdr) + 1
ix = {h: i for i, h in enumerate(hdr)}
byline = collections.defaultdict(lambda: [0, 0])
byreg = collections.defaultdict(lambda: [0, 0])
tot_i = tot_s = 0
base = None
for r in rows[start:]:
    try:
        n = int(r[ix["Instructions Executed"]])
        s = int(r[ix["# Samples"]])
        a = int(r[ix["Address"]], 16)
    except Exception:
        continue
    if base is None:
        base = a
    fl = addr2line.get(a - base)
    tot_i += n
    tot_s += s
    byline[fl][0] += n
    byline[fl][1] += s
    name = "other"
    if fl:
        for rn, fsub, l0, l1 in regions:
            if fsub in fl[0] and l0 <= fl[1] <= l1:
                name = rn
                break
    byreg[name][0] += n
    byreg[name][1] += s
print(f"total warp instructions {tot_i}, stall samples {tot_s}, mapped addresses {len(addr2line)}")
if regions:
    print("regions (share of instructions, share of samples):")
    for name, (n, s) in sorted(byreg.items(), key=lambda kv: -kv[1][1]):
        print(f"  {name:34s} {100 * n / tot_i:5.1f}% instr  {100 * s / max(1, tot_s):5.1f}% samples")
print("top source lines by samples:")
for fl, (n, s) in sorted(byline.items(), key=lambda kv: -kv[1][1])[:45]:
    print(f"  {str(fl):38s} {100 * n / tot_i:5.1f}% instr  {100 * s / max(1, tot_s):5.1f}% samples")
