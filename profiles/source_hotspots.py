"""Hot source lines of one kernel from `ncu -i X.ncu-rep --page source --csv --print-source sass,cuda`."""
import csv
import sys


def num(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return 0.0


rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur, hdr, agg = None, None, {}
for r in rows:
    if not r:
        continue
    if r[0] in ("File Path", "File Name"):
        cur = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) < len(hdr) or r[0] == "":
        continue
    try:
        line = int(r[0])
    except ValueError:
        continue
    agg[(cur, line, r[1].strip()[:100])] = (num(r[hdr.index("Instructions Executed")]), num(r[hdr.index("# Samples")]))
ti = sum(v[0] for v in agg.values()) or 1
ts = sum(v[1] for v in agg.values()) or 1
files = {}
for (f, _, _), (i, s) in agg.items():
    files.setdefault(f, [0, 0])
    files[f][0] += i
    files[f][1] += s
for f, v in files.items():
    print(f"{f}: instructions {v[0] / ti:.3f}  stall samples {v[1] / ts:.3f}")
for k, v in sorted(agg.items(), key=lambda x: -x[1][1])[:top]:
    print(f"{k[0]:18s}:{k[1]:5d} inst={v[0] / ti:.3f} samp={v[1] / ts:.3f}  {k[2]}")
