# usage: python benchmarks/ncu_source_hist.py <source_page.csv> [regions|ops|dump]   -- executed-instruction histogram from
#        `ncu -i rep --page source --csv --print-source sass`
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; rows = rows[2:]
ia = hdr.index("Source"); ie = hdr.index("Instructions Executed"); isamp = hdr.index("# Samples")
tot = sum(int(r[ie]) for r in rows); tots = sum(int(r[isamp]) for r in rows)
print("total inst", tot, "samples", tots)
# print contiguous regions with exec count ranges
mode = sys.argv[2] if len(sys.argv) > 2 else "regions"
if mode == "dump":
    for i, r in enumerate(rows):
        print(f"{i:5d} {int(r[ie]):>10d} {int(r[isamp]):>6d}  {r[ia].strip()}")
elif mode == "ops":
    c = collections.Counter(); s = collections.Counter()
    for r in rows:
        toks = r[ia].split()
        op = toks[1] if toks[0].startswith("@") else toks[0]
        op = op.split(".")[0]
        c[op] += int(r[ie]); s[op] += int(r[isamp])
    for op, n in c.most_common(40):
        print(f"{op:12s} {n:>12d} {100*n/tot:6.2f}%  samples {100*s[op]/tots:6.2f}%")
else:
    # regions: group consecutive instructions with same exec count (within 1%)
    start = 0
    for i in range(1, len(rows) + 1):
        if i == len(rows) or abs(int(rows[i][ie]) - int(rows[start][ie])) > 0.02 * max(1, int(rows[start][ie])):
            n = i - start; e = int(rows[start][ie]); smp = sum(int(rows[j][isamp]) for j in range(start, i))
            if e * n > 0.003 * tot:
                print(f"rows {start:5d}-{i-1:5d} n={n:4d} exec={e:>10d} total={e*n:>11d} ({100*e*n/tot:5.1f}%) samples={100*smp/tots:5.1f}%  first: {rows[start][ia].strip()[:60]}")
            start = i
