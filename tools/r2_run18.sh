#!/bin/bash
# group count rule (groups of >= 128 worlds) at the per GPU shares of N = 8 / 4 / 2, two EPA tiers again
mkdir -p gpurun_out
for cfg in "512 0" "1024 0" "1024 4" "2048 0" "256 0" "256 1"; do
  set -- $cfg
  if [ "$2" != "0" ]; then export B2J_BATCH_GROUPS=$2; else unset B2J_BATCH_GROUPS; fi
  timeout 600 python bench.py --steps 20 --warmup 5 --worlds $1 --no-pile --no-extras --no-cpu-baseline > gpurun_out/run18_w$1_g$2.json 2> gpurun_out/run18_w$1_g$2.err
  python - "$1" "$2" <<'PY'
import json, sys
w, g = sys.argv[1:3]
for l in open(f"gpurun_out/run18_w{w}_g{g}.json"):
    if l.startswith("{"):
        d = json.loads(l); print(f"worlds {w} groups {g or 'default'}: {d['ms_per_step']:.2f} ms/step e2e {d['e2e']['value']/1e6:.1f}M", [round(x, 1) for x in d['ms_per_step_series']][:3])
PY
done
unset B2J_BATCH_GROUPS
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/run18_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/run18_tests.log
