"""DRAM traffic of the velocity solve kernel against its algorithmic bytes, from an `ncu --set full` capture of a bench.py step.
usage: python tools/ncu_traffic.py capture.ncu-rep bench_stdout.log > profiles/rN_solve_traffic.json
Per captured KSolveVelocity launch: dram__bytes_read.sum + dram__bytes_write.sum and the constraints it processed (grid size x 128
threads: one thread per constraint, the last block may be partial, < 0.1 % for these grids); algorithmic bytes per constraint and
iteration = C(c) + 4 S_v + 4 (3 + c) with C(c) = 220 + 64 c and c = mean points per constraint of the SAME step
(the bench line in the log). bench.py scales the ratio to its own bytes per launch for `roofline.traffic`."""
import csv, io, json, re, subprocess, sys

rep, log = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--kernel-name-base", "demangled"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h, units = rows[0], rows[1]
col = {n: i for i, n in enumerate(h)}
def val(r, name):
    v = float(r[col[name]].replace(",", "")); u = units[col[name]]
    return v * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)
cbar = None
for line in open(log):
    if line.startswith("{") and "step_counters_mean" in line:
        p = json.loads(line)
        c = p["step_counters_mean"]
        cbar = c["num_contact_points"] / max(c["num_constraints"], 1)
launches = []
for r in rows[2:]:
    name = r[col["Kernel Name"]]
    if "KSolveVelocity" not in name or "All" in name:
        continue
    n = int(float(r[col["launch__grid_size"]])) * int(float(r[col["launch__block_size"]]))
    launches.append({"constraints_upper_bound": n, "dram_read_bytes": val(r, "dram__bytes_read.sum"), "dram_write_bytes": val(r, "dram__bytes_write.sum"),
                     "duration_us": float(r[col["gpu__time_duration.sum"]].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(units[col["gpu__time_duration.sum"]], 1.0)})
big = [l for l in launches if l["constraints_upper_bound"] >= 32768]  # small launches are dominated by sector granularity
use = big or launches
per_constraint = (220 + 64 * cbar) + 4 * 24 + 4 * (3 + cbar) if cbar is not None else None
dram = sum(l["dram_read_bytes"] + l["dram_write_bytes"] for l in use)
alg = sum(l["constraints_upper_bound"] for l in use) * per_constraint if per_constraint else None
print(json.dumps({"kernel": "KSolveVelocity", "capture": rep.split("/")[-1], "points_per_constraint": cbar, "algorithmic_bytes_per_constraint": per_constraint,
                  "launches": launches, "launches_used": len(use), "dram_bytes_per_algorithmic_byte": dram / alg if alg else None}, indent=1))
