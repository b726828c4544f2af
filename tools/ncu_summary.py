"""Text summary of an .ncu-rep for profiles/: per captured launch the headline metrics (raw page), then the stall reason mix of the
first launch of every kernel (source page). Runs where ncu is installed, no GPU needed.
usage: python tools/ncu_summary.py report.ncu-rep > profiles/NAME.txt"""
import csv, io, subprocess, sys, re

rep = sys.argv[1]
def ncu(*args):
    return subprocess.run(["ncu", "-i", rep, *args], capture_output=True, text=True).stdout

METRICS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "regs"),
    ("dram__bytes_read.sum", "dram_read"), ("dram__bytes_write.sum", "dram_write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("l1tex__t_sector_hit_rate.pct", "l1_hit_pct"), ("lts__t_sector_hit_rate.pct", "l2_hit_pct"),
    ("sm__inst_executed.avg.per_cycle_elapsed", "ipc"), ("smsp__inst_executed.sum", "warp_inst"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active_lanes"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy_pct"),
]
rows = list(csv.reader(io.StringIO(ncu("--page", "raw", "--csv", "--kernel-name-base", "demangled"))))
h, units = rows[0], rows[1]
kn = h.index("Kernel Name")
def short(n):
    m = re.search(r"run_kernel\w*<(?:b2j::)?(.*?)>?\(", n)
    return (m.group(1) if m else n)[:70]
print(f"# {rep.split('/')[-1]}: {len(rows) - 2} captured launches (ncu --set full --clock-control none; cold cache, serialised)")
seen = {}
for r in rows[2:]:
    name = short(r[kn])
    seen.setdefault(name, len(seen))
    out = [f"{label}={r[h.index(m)]}{(' ' + units[h.index(m)]) if units[h.index(m)] and label in ('duration', 'dram_read', 'dram_write') else ''}" for m, label in METRICS if m in h]
    print(f"- {name}\n    " + "  ".join(out))
print()
for name in seen:
    pat = re.escape(name.split("<")[0])
    src = list(csv.reader(io.StringIO(ncu("--page", "source", "--csv", "--kernel-name-base", "demangled", "-k", "regex:" + pat, "-c", "1"))))
    try:
        hi = next(i for i, r in enumerate(src) if r and r[0] == "Address")
    except StopIteration:
        continue
    hh = src[hi]; data = []
    for r in src[hi + 1:]:
        if len(r) != len(hh) or r[0] == "Address": break
        data.append(r)
    cols = [i for i, n in enumerate(hh) if n.startswith("stall_") and "Not Issued" not in n]
    tot = {hh[i]: sum(int(r[i] or 0) for r in data) for i in cols}
    s = sum(tot.values()) or 1
    mix = ", ".join(f"{k[6:]} {100 * v / s:.1f}%" for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:6] if v)
    print(f"stalls {name}: {len(data)} SASS instructions; {mix}")
