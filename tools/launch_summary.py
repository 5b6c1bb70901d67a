"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel name."""
import csv, collections, sys
for f in sys.argv[1:]:
    rows = [r for r in csv.reader(open(f)) if len(r) > 10 and r[0].isdigit()]
    agg = collections.OrderedDict()
    for r in rows:
        name = r[4].split("(")[0][-44:]
        v = float(r[-1].replace(",", ""))
        unit = r[-2]
        us = v / 1000 if unit in ("ns", "nsecond") else (v if unit.startswith("us") else v * 1000)
        a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += us
    print(f, "total us %.1f" % sum(a[1] for a in agg.values()))
    for k, (n, t) in agg.items():
        print(f"  {k:46s} n={n:4d} total={t:9.1f} us avg={t/n:8.1f}")
