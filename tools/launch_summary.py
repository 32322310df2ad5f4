"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel name, launches and mean / total time.

    python tools/launch_summary.py gpurun_out/launches.csv [name-filter]
"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1], errors="replace")))
flt = sys.argv[2] if len(sys.argv) > 2 else ""
hdr = None
agg = collections.OrderedDict()
for r in rows:
    if "Kernel Name" in r:
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        try:
            v = float(d["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        name = d["Kernel Name"]
        name = name[:name.index("(")] if "(" in name else name
        if flt and flt not in name:
            continue
        agg.setdefault(name[-70:], []).append(v / 1e3)
for k, v in agg.items():
    print(f"{k:72s} n={len(v):4d} mean={sum(v) / len(v):9.2f} us  last={v[-1]:9.2f} us")
