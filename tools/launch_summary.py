"""Summarise an ncu --csv launch list (gpu__time_duration.sum): per kernel count / mean / total and share of the sum."""
import collections
import csv
import sys

for f in sys.argv[1:]:
    rows = [r for r in csv.reader(open(f)) if len(r) > 5]
    hdr = [i for i, r in enumerate(rows) if r[0] == "ID"]
    if not hdr:
        print(f, "no header")
        continue
    h = rows[hdr[0]]
    ki, vi = h.index("Kernel Name"), h.index("Metric Value")
    d = collections.OrderedDict()
    for r in rows[hdr[0] + 1:]:
        try:
            d.setdefault(r[ki][:70], []).append(float(r[vi].replace(",", "")))
        except ValueError:
            pass
    tot = sum(sum(v) for v in d.values())
    print(f)
    for k, v in d.items():
        print(f"  {k:70s} n={len(v):3d} mean={sum(v)/len(v)/1e3:9.1f} us  share={100*sum(v)/tot:5.1f}%")
