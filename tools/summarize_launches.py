"""Per-kernel time shares of ONE step from an ncu launch list (gpu__time_duration.sum, --csv):
python tools/summarize_launches.py gpurun_out/launches_all.csv  -> the last full step (weight pack .. adamw_flat_kernel)."""
import collections, csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
h = rows[hi]; body = [r for r in rows[hi + 2:] if len(r) > h.index('Metric Value')]
ki, vi = h.index('Kernel Name'), h.index('Metric Value')
ends = [i for i, r in enumerate(body) if 'adamw_flat_kernel' in r[ki]]  # last kernel of a step
step = body[ends[-2] + 1:ends[-1] + 1]
agg = collections.OrderedDict()
for r in step:
    name = re.sub(r'\(.*', '', r[ki]); name = re.sub(r'^void ', '', name).replace('theia::', '')
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += float(r[vi].replace(',', '')) / 1e3
tot = sum(v[1] for v in agg.values())
print(f"one step: {len(step)} launches, {tot/1e3:.2f} ms of kernel time (ncu: serialized, cold clocks -- shares, not absolutes)")
print(f"{'kernel':72s} {'n':>5s} {'us':>10s} {'share':>6s} {'us/launch':>9s}")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    if v[1] / tot < 0.0005: continue
    print(f"{k[:72]:72s} {v[0]:5d} {v[1]:10.1f} {100*v[1]/tot:5.1f}% {v[1]/v[0]:9.1f}")
