"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel:
kernel, launches, total ms, share of the summed (serialised, cold-cache) durations.
usage: python tools/launch_shares.py launches.csv > shares.csv"""
import collections, csv, re, sys
tot = collections.OrderedDict()
rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if len(r) > 5]
hdr = next(r for r in rows if "Kernel Name" in r)
ki, mi, vi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
for r in rows:
    if len(r) <= vi or r[mi] != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r[ki]).replace("void ", "").strip()
    e = tot.setdefault(name, [0, 0.0])
    e[0] += 1
    e[1] += float(r[vi].replace(",", "")) / 1e6
s = sum(v[1] for v in tot.values())
print("kernel,launches,total_ms,share_pct")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print("%s,%d,%.3f,%.2f" % (k, v[0], v[1], 100.0 * v[1] / s))
