"""Summarise an `ncu --page source --csv` export: instructions executed / stall samples by opcode and by region."""
import csv, sys, collections, re
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
tot = 0; byop = collections.Counter(); samp = collections.Counter(); lines = []
for r in data:
    try:
        n = int(r[ix["Instructions Executed"]]); s = int(r[ix["# Samples"]])
    except Exception:
        continue
    src = r[ix["Source"]].strip()
    op = src.split()[0] if src else "?"
    if op.startswith("@"):
        op = src.split()[1]
    op = op.split(".")[0]
    byop[op] += n; samp[op] += s; tot += n
    lines.append((int(r[ix["Address"]], 16) if r[ix["Address"]].startswith("0x") else len(lines), n, s, src, r))
print("total warp instrs", tot, "SASS lines", len(lines))
for op, n in byop.most_common(28):
    print(f"  {op:12s} {n:12d} {100*n/tot:5.1f}%   samples {samp[op]:7d}")
# shared memory wavefronts by line
if "L1 Wavefronts Shared" in ix:
    ws = []
    for a, n, s, src, r in lines:
        try:
            w = int(r[ix["L1 Wavefronts Shared"]]); wi = int(r[ix["L1 Wavefronts Shared Ideal"]])
        except Exception:
            continue
        if w: ws.append((w, wi, n, src))
    ws.sort(reverse=True)
    print("top shared-memory lines (wavefronts, ideal, execs):")
    for w, wi, n, src in ws[:14]:
        print(f"  {w:11d} {wi:11d} {n:10d}  {src[:70]}")
# execution-count histogram -> regions
print("regions by execution count:")
reg = collections.Counter(); regs = collections.Counter()
for a, n, s, src, r in lines:
    reg[n] += 1; regs[n] += s
for n, c in sorted(reg.items(), key=lambda kv: -kv[0] * kv[1])[:16]:
    print(f"  exec {n:10d} x {c:5d} lines = {n*c:12d} ({100*n*c/tot:5.1f}%)  samples {regs[n]}")
if len(sys.argv) > 2:
    for a, n, s, src, r in lines:
        print(f"{a:6x} {n:10d} {s:6d}  {src}")
