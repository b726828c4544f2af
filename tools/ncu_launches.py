"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per step (steps start at KApplyGravity).
usage: python tools/ncu_launches.py launches.csv"""
import csv, re, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == 'ID')
h = rows[hi]; kn = h.index('Kernel Name'); mv = h.index('Metric Value'); mu = h.index('Metric Unit')
def name(s):
    m = re.search(r'run_kernel\w*<(?:b2j::)?(\w+(?:<[^>]*>)?)', s)
    if m: return m.group(1)
    m = re.search(r'(DeviceRadixSort\w+|DeviceScan\w+|\w+Kernel)', s)
    return m.group(1) if m else s[:40]
def us(r):
    v = float(r[mv].replace(',', '')); u = r[mu]
    return v / 1000.0 if u.startswith('ns') else v if u.startswith('us') else v * 1000.0 if u.startswith('ms') else v * 1e6
steps = []
for r in rows[hi + 1:]:
    n = name(r[kn])
    if n == 'KApplyGravity' or not steps: steps.append([])
    steps[-1].append((n, us(r)))
if len(sys.argv) > 2 and sys.argv[2] == "--total":
    steps = [[x for st in steps for x in st]]  # concurrent batch groups interleave their launches: one table for the whole capture
for si, st in enumerate(steps):
    tot = sum(t for _, t in st)
    agg = collections.defaultdict(lambda: [0, 0.0])
    for n, t in st: agg[n][0] += 1; agg[n][1] += t
    print(f"step {si}: {len(st)} launches, {tot/1000:.2f} ms of kernel time (cold cache, serialised)")
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:22]:
        print(f"  {n:44s} x{c:4d} {t/1000:8.3f} ms {100*t/tot:5.1f}%")
