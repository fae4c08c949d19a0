"""Print SASS size of each device function of one kernel (instruction-cache footprint)."""
import re, subprocess, sys
so = sys.argv[1] if len(sys.argv) > 1 else "geobipy_b200/libgeobipy_b200.so"
pat = sys.argv[2] if len(sys.argv) > 2 else "rjmcmc_kernelIfLi12"
out = subprocess.run(["cuobjdump", "-elf", so], capture_output=True, text=True).stdout
rows, total = [], 0
for ln in out.splitlines():
    p = ln.split()
    if len(p) >= 7 and pat in p[-1] and p[-1].startswith("$"):
        try:
            size = int(p[2], 16)
        except ValueError:
            continue
        name = p[-1].split("$")[-1]
        rows.append((size, name))
for ln in out.splitlines():
    m = re.match(r"\s*\w+\s+\w+\s+(\w+)\s+\w+\s+\w+\s+PROGBITS.*\.text\.(\S*%s\S*)" % pat, ln)
    if m:
        total = int(m.group(1), 16)
rows.sort()
for s, n in rows:
    print("%7.1f KB  %s" % (s / 1024, n[:100]))
print("sum of callees %.1f KB; .text section %.1f KB" % (sum(s for s, _ in rows) / 1024, total / 1024))
