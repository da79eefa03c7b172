"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv --log-file X): per kernel count, mean, share.
usage: python scripts/launch_summary.py gpurun_out/launches.csv"""
import csv, sys, collections, re
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if not l.startswith("=="))]
hdr = rows[0]
ki, mi, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[1:]:
    if len(r) <= vi or r[mi] != "gpu__time_duration.sum":
        continue
    v = float(r[vi].replace(",", ""))
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[ui], 1.0)
    name = re.sub(r"\(.*", "", r[ki])
    agg.setdefault(name, []).append(v)
tot = sum(sum(v) for v in agg.values())
print("# kernel, launches, total_us, mean_us, share")
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print("%s, %d, %.1f, %.1f, %.1f%%" % (k, len(v), sum(v), sum(v) / len(v), 100 * sum(v) / tot))
