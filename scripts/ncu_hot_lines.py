"""Hot SASS bytes per source line of one kernel: joins `nvdisasm -g -c` line info with the execution counts of an ncu
source-page CSV.  usage: ncu_hot_lines.py <source.csv> <lineinfo.sass> <kernel mangled pattern> <warp-iterations> [frac]"""
import csv, re, sys
from collections import defaultdict
src, sass, pat, iters = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
frac = float(sys.argv[5]) if len(sys.argv) > 5 else 0.05
line_of = {}
cur, inside = None, False
for ln in open(sass):
    if ln.startswith("//---") and ".text." in ln:
        inside = pat in ln
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", ln)
    if m:
        line_of[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(src)))
hdr = rows[1]
ia, ii = hdr.index("Address"), hdr.index("Instructions Executed")
base = None
hot = defaultdict(lambda: [0, 0])
for r in rows[2:]:
    if len(r) <= ii:
        continue
    a = int(r[ia], 16)
    if base is None:
        base = a
    n = int(r[ii] or 0)
    key = line_of.get(a - base)
    if key is None:
        continue
    if n >= frac * iters:
        hot[key][0] += 16
        hot[key][1] += n
byfile = defaultdict(int)
for (f, l), (b, n) in hot.items():
    byfile[f] += b
print({k: round(v / 1024, 1) for k, v in byfile.items()})
# group by file and 10-line buckets
buck = defaultdict(lambda: [0, 0])
for (f, l), (b, n) in hot.items():
    k = (f, l // 10 * 10)
    buck[k][0] += b
    buck[k][1] += n
for (f, l), (b, n) in sorted(buck.items(), key=lambda x: -x[1][0])[:70]:
    print("%-22s %5d-%-5d %6.2f KB  %6.2f instr/iter" % (f, l, l + 9, b / 1024, n / iters))
