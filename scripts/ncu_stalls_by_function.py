"""Per-function stall-reason breakdown from an ncu source-page CSV."""
import csv, subprocess, sys
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
cols = ["stall_no_inst", "stall_long_sb", "stall_short_sb", "stall_wait", "stall_math", "stall_branch_resolving", "stall_dispatch", "stall_not_selected", "stall_selected", "stall_mio"]
idx = [hdr.index(c) for c in cols]
ia = hdr.index("Address")
base = None
agg = {}
for r in rows[2:]:
    if len(r) <= max(idx):
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
    d = agg.setdefault(name, [0] * len(cols))
    for j, i in enumerate(idx):
        d[j] += int(r[i] or 0)
tot = [sum(v[j] for v in agg.values()) for j in range(len(cols))]
print("%-44s" % "function" + "".join("%10s" % c.replace("stall_", "")[:9] for c in cols))
for n, v in sorted(agg.items(), key=lambda x: -sum(x[1]))[:14]:
    print("%-44s" % n[:44] + "".join("%10d" % x for x in v))
print("%-44s" % "TOTAL" + "".join("%10d" % x for x in tot))
