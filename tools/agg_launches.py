"""Dev tool: aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (last step only)."""
import csv, collections, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
data = rows[hdr + 1:]
names = [r[4] for r in data]
adam = [i for i, n in enumerate(names) if 'adam' in n]
seg = data[adam[-2] + 1:adam[-1] + 1] if len(adam) >= 2 else data
agg = collections.OrderedDict()
tot = 0
for r in seg:
    n = re.sub(r'\(.*', '', r[4]).replace('void wspc::<unnamed>::', '').replace('wspc::<unnamed>::', '')
    t = float(r[-1]) / 1e6
    tot += t
    agg.setdefault(n, [0, 0.0])
    agg[n][0] += 1
    agg[n][1] += t
for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{t:9.3f} ms {100*t/tot:5.1f}% {c:3d}x  {n}")
print("total", round(tot, 3), "ms; launches", len(seg))
