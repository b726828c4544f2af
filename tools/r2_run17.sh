#!/bin/bash
mkdir -p gpurun_out
B2J_TRACE_EPA=1 B2J_BATCH_GROUPS=1 timeout 600 python tools/diag_landing.py 512 21 24 > gpurun_out/run17_landing.log 2>&1; grep -v "^\[plain\]" gpurun_out/run17_landing.log | cut -c1-420 | tail -34
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/run17_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/run17_tests.log
for worlds in 4096 512; do
  timeout 600 python bench.py --steps 20 --warmup 5 --worlds $worlds --no-pile --no-extras --no-cpu-baseline > gpurun_out/run17_w$worlds.json 2> gpurun_out/run17_w$worlds.err
  python - "$worlds" <<'PY'
import json, sys
w = sys.argv[1]
for l in open(f"gpurun_out/run17_w{w}.json"):
    if l.startswith("{"):
        d = json.loads(l); print(f"worlds {w}: {d['ms_per_step']:.2f} ms/step e2e {d['e2e']['value']/1e6:.1f}M", [round(x, 1) for x in d['ms_per_step_series']])
PY
done
timeout 600 python bench.py --workload pile --steps 30 --warmup 120 --no-cpu-baseline > gpurun_out/run17_pile.json 2> gpurun_out/run17_pile.err
python - <<'PY'
import json
for l in open("gpurun_out/run17_pile.json"):
    if l.startswith("{"):
        d = json.loads(l); print(f"pile formed: {d['ms_per_step']:.2f} ms/step", d.get("kernel_ms_per_step"))
PY
timeout 300 python tools/diag_small.py 2>&1 | grep -E "^(convex_vs_mesh|pyramid):"
