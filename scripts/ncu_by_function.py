"""Aggregate an ncu source-page CSV (SASS view) by device function using the cubin symbol table.
usage: ncu_by_function.py <source.csv> <lib.so> <kernel mangled-name pattern>"""
import csv, re, subprocess, sys
src, so, pat = sys.argv[1], sys.argv[2], sys.argv[3]
out = subprocess.run(["cuobjdump", "-elf", so], capture_output=True, text=True).stdout
syms = []
for ln in out.splitlines():
    p = ln.split()
    if len(p) >= 7 and pat in p[-1] and p[-1].startswith("$"):
        try:
            syms.append((int(p[1], 16), int(p[2], 16), p[-1].split("$")[-1]))
        except ValueError:
            pass
syms.sort()
rows = list(csv.reader(open(src)))
hdr = rows[1]
ia, ii, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
base = None
agg = {}
for r in rows[2:]:
    if len(r) <= ii:
        continue
    a = int(r[ia], 16)
    if base is None:
        base = a
    off = a - base
    name = "<kernel body>"
    for s0, sz, n in syms:
        if s0 <= off < s0 + sz:
            name = n
            break
    d = agg.setdefault(name, [0, 0])
    d[0] += int(r[ii] or 0)
    d[1] += int(r[isamp] or 0)
ti = sum(v[0] for v in agg.values()); ts = sum(v[1] for v in agg.values())
print("%-60s %14s %7s %9s %7s" % ("function", "warp-instr", "%instr", "samples", "%smpl"))
for n, v in sorted(agg.items(), key=lambda x: -x[1][1]):
    print("%-60s %14d %6.1f%% %9d %6.1f%%" % (n[:60], v[0], 100.0 * v[0] / ti, v[1], 100.0 * v[1] / ts))
print("total warp-instr", ti, "samples", ts)
