"""one training step out of an ncu launch list (gpu__time_duration.sum csv): python scripts/launch_step.py CSV [marker]"""
import csv, collections, sys
path = sys.argv[1]; marker = sys.argv[2] if len(sys.argv) > 2 else "fused2_fwd"
rows = [r for r in csv.reader(open(path)) if len(r) > 5]
hdr = rows[0]; ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value')
seq = [(r[ki], float(r[vi].replace(',', ''))) for r in rows[1:] if r[vi].replace(',', '').replace('.', '').isdigit()]
idx = [i for i, (k, v) in enumerate(seq) if marker in k]
a, b = idx[-3], idx[-2]
step = seq[a:b]
tot = sum(v for k, v in step)
print('launches in one step: %d, sum of durations %.1f us (cold-cache, serialised: compare SHARES)' % (len(step), tot / 1e3))
agg = collections.OrderedDict()
for k, v in step:
    kk = k.replace('void ', '')[:90]; agg.setdefault(kk, [0, 0.0]); agg[kk][0] += 1; agg[kk][1] += v
print('| share | us | launches | kernel |\n|---|---|---|---|')
for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[3]) if len(sys.argv) > 3 else 24]:
    print('| %4.1f%% | %7.1f | %d | `%s` |' % (100 * v / tot, v / 1e3, n, k))
