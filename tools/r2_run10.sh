#!/bin/bash
# round 2 run 10: the warp cooperative mesh kernel -- parity, then its time next to the serial form (B2J_MESH_SERIAL=1)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/run10_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/run10_tests.log
tail -3 gpurun_out/run10_tests.log
for serial in 0 1; do
  B2J_MESH_SERIAL=$serial timeout 300 python tools/diag_small.py > gpurun_out/run10_small_serial$serial.log 2>&1
  grep -E "^(convex_vs_mesh|pyramid):" gpurun_out/run10_small_serial$serial.log
  B2J_MESH_SERIAL=$serial timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:KCollideMesh -c 60 --csv --log-file gpurun_out/run10_mesh_serial$serial.csv python tools/diag_mesh.py > gpurun_out/run10_ncu_serial$serial.log 2>&1
done
python - <<'PY'
import csv, statistics
for serial in (0, 1):
    rows = [r for r in csv.reader(open(f"gpurun_out/run10_mesh_serial{serial}.csv")) if len(r) > 5 and r[0].isdigit()]
    d = [float(r[-1].replace(",", "")) for r in rows]
    if d: print(f"serial={serial}: {len(d)} launches, median {statistics.median(d)/1000:.1f} us, max {max(d)/1000:.1f} us")
PY
