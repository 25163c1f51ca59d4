"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel count, total and share."""
import collections, csv, io, sys
rows = [l for l in open(sys.argv[1]) if l.startswith('"')]
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
agg = collections.OrderedDict()
tot = 0.0
for i, x in enumerate(csv.DictReader(io.StringIO(''.join(rows)))):
    if i < skip:
        continue
    k = x['Kernel Name'].split('(')[0].replace('void ', '').replace('g16::', '')[:60]
    v = float(x['Metric Value'].replace(',', ''))
    v = v / 1e3 if x['Metric Unit'] == 'ns' else (v * 1e3 if x['Metric Unit'] == 'ms' else v)
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1; a[1] += v; tot += v
for k, (n, v) in agg.items():
    print(f"{k:62s} n={n:4d} total={v:10.1f} us  share={100*v/tot:5.1f}%")
print("total", round(tot, 1), "us")
