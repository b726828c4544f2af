#!/bin/bash
mkdir -p gpurun_out
B2J_BATCH_GROUPS=1 timeout 600 python tools/diag_landing.py 512 15 26 > gpurun_out/run13_landing.log 2>&1; cat gpurun_out/run13_landing.log | tail -40
for g in 1 4 8; do
  B2J_BATCH_GROUPS=$g timeout 600 python bench.py --steps 20 --warmup 5 --worlds 512 --no-pile --no-extras --no-cpu-baseline > gpurun_out/run13_w512_g$g.json 2> gpurun_out/run13_w512_g$g.err
  python - "$g" <<'PY'
import json, sys
g = sys.argv[1]
for l in open(f"gpurun_out/run13_w512_g{g}.json"):
    if l.startswith("{"):
        d = json.loads(l); print(f"worlds 512 groups {g}: {d['ms_per_step']:.2f} ms/step", [round(x, 1) for x in d['ms_per_step_series']])
PY
done
