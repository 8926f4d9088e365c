"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel launches, total time, share.
usage: python tools/summarise_launches.py gpurun_out/r01d_launches_raw.csv > profiles/r01d_launches_summary.csv"""
import collections, csv, re, sys

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot = collections.defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    v = float(r[vi].replace(",", ""))
    v = v / 1e3 if r[ui] == "ns" else v * (1e3 if r[ui] == "ms" else 1.0)      # -> us
    name = re.sub(r"\(.*$", "", r[ki]).strip()
    tot[name][0] += 1
    tot[name][1] += v
total = sum(v[1] for v in tot.values())
w = csv.writer(sys.stdout)
w.writerow(["kernel", "launches", "total_us", "share_pct"])
for k, (n, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    w.writerow([k, n, f"{t:.1f}", f"{100 * t / total:.2f}"])
