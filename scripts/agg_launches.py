"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import collections, csv, re, sys
path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 50
hdr = None
agg = collections.defaultdict(lambda: [0, 0.0])
for r in csv.reader(open(path)):
    if hdr is None:
        if 'Kernel Name' in r:
            hdr = r; ki = r.index('Kernel Name'); vi = r.index('Metric Value'); ui = r.index('Metric Unit')
        continue
    if len(r) <= vi:
        continue
    v = float(r[vi].replace(',', '')); u = r[ui]
    v = v / 1e6 if u == 'ns' else (v / 1e3 if u in ('us', 'usecond') else v)
    name = re.sub(r'\(.*', '', r[ki])[:100]
    agg[name][0] += 1; agg[name][1] += v
tot = sum(a[1] for a in agg.values())
print("total GPU time %.2f ms over %d launches" % (tot, sum(a[0] for a in agg.values())))
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print("%9.3f ms %5.1f%% n=%4d  %s" % (a[1], 100 * a[1] / tot, a[0], k))
