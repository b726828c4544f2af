"""Aggregate the `ncu --page source --csv` output of one kernel: stall reasons (totals) and the hottest SASS lines.
usage: ncu -i rep --page source --csv -k regex:NAME -c 1 | python tools/ncu_stalls.py [top_n]"""
import csv, sys
rows = list(csv.reader(sys.stdin))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hi]
data = []
for r in rows[hi + 1:]:
    if len(r) != len(h) or r[0] == "Address": break  # first (SASS) section only
    data.append(r)
stall_cols = [i for i, n in enumerate(h) if n.startswith("stall_") and "Not Issued" not in n]
samp = h.index("# Samples"); src = h.index("Source"); inst = h.index("Instructions Executed")
tot = {h[i]: 0 for i in stall_cols}
for r in data:
    for i in stall_cols:
        try: tot[h[i]] += int(r[i])
        except ValueError: pass
total = sum(tot.values()) or 1
print("instructions", len(data), "warp-inst executed", sum(int(r[inst] or 0) for r in data), "samples", total)
print({k: round(100.0 * v / total, 1) for k, v in sorted(tot.items(), key=lambda kv: -kv[1]) if v})
top = int(sys.argv[1]) if len(sys.argv) > 1 else 25
for r in sorted(data, key=lambda r: -int(r[samp] or 0))[:top]:
    reasons = sorted(((int(r[i] or 0), h[i][6:]) for i in stall_cols), reverse=True)[:2]
    print(r[samp].rjust(6), r[inst].rjust(8), r[src][:90].ljust(90), reasons)
