"""Join an `ncu --page source --csv` SASS profile with `nvdisasm --print-line-info` of the same kernel: executed warp
instructions, average active lanes and stall samples per SOURCE line.
usage: python tools/ncu_lines.py sass_profile.csv disasm_with_lines.txt [top_n]"""
import csv, re, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
h = rows[hi]; data = []
for r in rows[hi + 1:]:
    if len(r) != len(h) or r[0] == 'Address': break
    data.append(r)
ie = h.index('Instructions Executed'); te = h.index('Thread Instructions Executed'); sa = h.index('# Samples')
lines = []; cur = None
for l in open(sys.argv[2]):
    if l.startswith('//---') and lines: break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    if re.match(r'\s+/\*[0-9a-f]+\*/', l): lines.append(cur)
assert len(lines) == len(data), (len(lines), len(data))
agg = collections.defaultdict(lambda: [0, 0, 0])
for r, ln in zip(data, lines):
    a = agg[ln]; a[0] += int(r[ie] or 0); a[1] += int(r[te] or 0); a[2] += int(r[sa] or 0)
tot = sum(a[2] for a in agg.values()) or 1
for ln, a in sorted(agg.items(), key=lambda kv: -kv[1][2])[:int(sys.argv[3]) if len(sys.argv) > 3 else 40]:
    print(f"{str(ln):40s} warp-inst {a[0]:12d} lanes {a[1]/max(a[0],1):5.2f} samples {100*a[2]/tot:5.1f}%")
