"""Aggregate an `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --csv` launch list by
kernel name: launches, total time, DRAM bytes read + written (total and per launch) -> JSON on stdout."""
import collections, csv, json, re, sys
hdr = None
agg = collections.defaultdict(lambda: {"launches": set(), "ms": 0.0, "bytes": 0.0})
for r in csv.reader(open(sys.argv[1])):
    if hdr is None:
        if 'Kernel Name' in r:
            hdr = r; ki = r.index('Kernel Name'); vi = r.index('Metric Value'); ui = r.index('Metric Unit'); mi = r.index('Metric Name'); ii = r.index('ID')
        continue
    if len(r) <= vi:
        continue
    name = re.sub(r'\(.*', '', r[ki]).replace('void ', '')[:80]
    v = float(r[vi].replace(',', '')); u = r[ui]; m = r[mi]
    a = agg[name]
    a["launches"].add(r[ii])
    if m.startswith('gpu__time'):
        a["ms"] += v / 1e6 if u == 'ns' else (v / 1e3 if u in ('us', 'usecond') else v)
    else:
        mult = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(u, 1)
        a["bytes"] += v * mult
out = {}
for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["ms"]):
    n = len(a["launches"])
    out[k] = {"launches": n, "ms": round(a["ms"], 3), "dram_bytes": a["bytes"], "dram_bytes_per_launch": a["bytes"] / max(n, 1),
              "dram_gbs": a["bytes"] / max(a["ms"], 1e-9) / 1e6}
print(json.dumps(out, indent=1))
