#!/bin/bash
# decorated shapes (f4): parity on the GPU, and that the hot path did not slow down
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -x > gpurun_out/run22_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/run22_tests.log
for worlds in 4096 512; do
  timeout 600 python bench.py --steps 20 --warmup 5 --worlds $worlds --no-pile --no-extras --no-cpu-baseline > gpurun_out/run22_w$worlds.json 2> gpurun_out/run22_w$worlds.err
  python - "$worlds" <<'PY'
import json, sys
w = sys.argv[1]
for l in open(f"gpurun_out/run22_w{w}.json"):
    if l.startswith("{"):
        d = json.loads(l); print(f"worlds {w}: {d['ms_per_step']:.2f} ms/step e2e {d['e2e']['value']/1e6:.1f}M", [round(x, 1) for x in d['ms_per_step_series']])
PY
done
timeout 600 python bench.py --workload pile --steps 30 --warmup 120 --no-cpu-baseline > gpurun_out/run22_pile.json 2> gpurun_out/run22_pile.err
python - <<'PY'
import json
for l in open("gpurun_out/run22_pile.json"):
    if l.startswith("{"):
        d = json.loads(l); print(f"pile formed: {d['ms_per_step']:.2f} ms/step", d.get("kernel_ms_per_step"))
PY
timeout 300 python tools/diag_small.py 2>&1 | grep -E "^(convex_vs_mesh|pyramid):"
