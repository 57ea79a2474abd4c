"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel (+grid)."""
import collections
import csv
import io
import sys

txt = open(sys.argv[1]).read().splitlines()
i = [k for k, l in enumerate(txt) if l.startswith('"ID"')][0]
rows = list(csv.DictReader(io.StringIO("\n".join(txt[i:]))))
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[skip:]:
    n = r["Kernel Name"]
    n = n.split("<")[0] if n.startswith("void at") else n.split("(")[0]
    key = (n[-60:], r["Grid Size"], r["Block Size"])
    agg[key][0] += 1
    agg[key][1] += float(r["Metric Value"])
tot = sum(v[1] for v in agg.values())
print(f"{len(rows) - skip} launches, {tot / 1e3:.1f} us total")
for (n, g, b), (c, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:30]:
    print(f"{t / 1e3:10.1f} us {100 * t / tot:5.1f}% {c:5d} x {t / c / 1e3:8.1f} us  grid {g:>14} block {b:>12}  {n}")
