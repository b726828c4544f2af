#!/bin/bash
mkdir -p gpurun_out
B2J_BATCH_GROUPS=1 timeout 600 python tools/diag_landing.py 512 18 26 > gpurun_out/run21_landing.log 2>&1; grep -A1 "^\[profiled\]" gpurun_out/run21_landing.log | cut -c1-330 | tail -24
for worlds in 4096 512; do
  timeout 600 python bench.py --steps 20 --warmup 5 --worlds $worlds --no-pile --no-extras --no-cpu-baseline > gpurun_out/run21_w$worlds.json 2> gpurun_out/run21_w$worlds.err
  python - "$worlds" <<'PY'
import json, sys
w = sys.argv[1]
for l in open(f"gpurun_out/run21_w{w}.json"):
    if l.startswith("{"):
        d = json.loads(l); print(f"worlds {w}: {d['ms_per_step']:.2f} ms/step e2e {d['e2e']['value']/1e6:.1f}M", [round(x, 1) for x in d['ms_per_step_series']])
PY
done
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/run21_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/run21_tests.log
