"""Turn an `ncu --metrics gpu__time_duration.sum --csv` launch list into a per-kernel share table (markdown).

usage: summarize_launches.py launches.csv out.md [title [delimiter-kernel k]]
With a delimiter (e.g. `adam_kernel 4`) only the launches after its (k-1)-th and up to its k-th occurrence are kept: one
whole step of a multi-step capture.
"""
import collections
import csv
import re
import sys

src, dst = sys.argv[1], sys.argv[2]
title = sys.argv[3] if len(sys.argv) > 3 else src
lines = [l for l in open(src) if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0.0, 0])
tot = 0.0
n = 0
rows = list(csv.DictReader(lines))
if len(sys.argv) > 5:
    delim, k = sys.argv[4], int(sys.argv[5])
    hits = [i for i, r in enumerate(rows) if delim in r.get("Kernel Name", "")]
    rows = rows[hits[k - 2] + 1:hits[k - 1] + 1]
for row in rows:
    try:
        t = float(row["Metric Value"].replace(",", ""))
    except (ValueError, KeyError):
        continue
    t *= {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(row["Metric Unit"], 1.0)
    name = re.sub(r"\(.*", "", row["Kernel Name"])
    name = re.sub(r"^void ", "", name)
    agg[name][0] += t
    agg[name][1] += 1
    tot += t
    n += 1
with open(dst, "w") as f:
    f.write("# %s\n\n" % title)
    f.write("ncu `--metrics gpu__time_duration.sum --clock-control none` (cold-cache, serialised launches: compare SHARES, "
            "not absolutes).  %d launches, %.2f ms of kernel time.\n\n" % (n, tot / 1e6))
    f.write("| kernel | launches | total ms | share |\n|---|---:|---:|---:|\n")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        f.write("| `%s` | %d | %.3f | %.1f %% |\n" % (k[:90], v[1], v[0] / 1e6, 100 * v[0] / tot))
print("wrote", dst, n, "launches", tot / 1e6, "ms")
