"""Trim an `ncu --metrics gpu__time_duration.sum --csv` launch list (kernel names cut to 80
chars) and print the per-kernel share of the total.
usage: python tools/launch_list.py gpurun_out/launches.csv profiles/rXX_launches.csv"""
import collections, csv, sys
src, dst = sys.argv[1], sys.argv[2]
rows = [r for r in csv.reader(l for l in open(src) if not l.startswith('=='))]
hdr = rows[0]; ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value')
with open(dst, 'w', newline='') as f:
    w = csv.writer(f)
    for r in rows:
        if len(r) > ki: r[ki] = r[ki][:80]
        w.writerow(r)
t = collections.Counter(); n = collections.Counter()
for r in rows[1:]:
    t[r[ki][:60]] += float(r[vi]); n[r[ki][:60]] += 1
tot = sum(t.values())
for k, v in t.most_common(10):
    print(f'{v/1e6:9.3f} ms {n[k]:3d}x {100*v/tot:5.1f}%  {k}')
