"""Per-launch conv-engine timing of one eager iteration (T2V_PROFILE_DUMP csv) aggregated by shape."""
import collections
import csv
import os
import subprocess
import sys

batch = sys.argv[1] if len(sys.argv) > 1 else "256"
out = "gpurun_out/conv_launches_b%s.csv" % batch
os.makedirs("gpurun_out", exist_ok=True)
if os.path.exists(out):
    os.remove(out)
env = dict(os.environ, T2V_PROFILE_DUMP=out)
subprocess.check_call([sys.executable, "scripts/iter_once.py", "--batch", batch, "--convprof"], env=env)
agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0])
for r in csv.reader(open(out)):
    key = tuple(r[:10])
    a = agg[key]
    a[0] += 1
    a[1] += float(r[11])
    a[2] += float(r[12])
    a[3] = int(r[10])
tot = sum(a[1] for a in agg.values())
print("total conv-engine time %.2f ms" % tot)
with open("gpurun_out/conv_shapes_b%s.txt" % batch, "w") as f:
    for key, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(os.environ.get('TOPN', '45'))]:
        line = "%s N=%s D=%s H=%s W=%s Cin=%s Cout=%s k=%s%s%s  n=%d ctas=%d  %.3f ms (%.1f%%)  %.1f TF/s" % (
            ("fprop", "wgrad")[int(key[0])], key[1], key[2], key[3], key[4], key[5], key[6], key[7], key[8], key[9],
            a[0], a[3], a[1], 100 * a[1] / tot, a[2] / a[1] / 1e9)
        print(line)
        f.write(line + "\n")
