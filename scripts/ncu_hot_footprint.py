"""Hot instruction footprint per device function from an ncu source-page CSV: bytes of SASS whose execution count is at
least `frac` of the per-warp iteration count (instructions are 16 B).  usage: ncu_hot_footprint.py <source.csv> <lib.so> <kernel pattern> <warp-iterations>"""
import csv, subprocess, sys
src, so, pat, iters = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
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
ia, ii = hdr.index("Address"), hdr.index("Instructions Executed")
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
    n = int(r[ii] or 0)
    d = agg.setdefault(name, [0, 0, 0, 0])
    d[0] += 16
    if n >= 0.05 * iters: d[1] += 16
    if n >= 0.3 * iters: d[2] += 16
    if n >= 0.9 * iters: d[3] += 16
print("%-50s %8s %10s %10s %10s" % ("function", "KB", ">=5%/it", ">=30%/it", ">=90%/it"))
tot = [0, 0, 0, 0]
for n, v in sorted(agg.items(), key=lambda x: -x[1][1]):
    print("%-50s %8.1f %10.1f %10.1f %10.1f" % (n[:50], v[0] / 1024, v[1] / 1024, v[2] / 1024, v[3] / 1024))
    tot = [a + b for a, b in zip(tot, v)]
print("%-50s %8.1f %10.1f %10.1f %10.1f" % ("TOTAL", tot[0] / 1024, tot[1] / 1024, tot[2] / 1024, tot[3] / 1024))
