#!/bin/bash
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
tools/r2_solve_ab.sh "0:8 0:16"
echo "---- 512 worlds"; B2J_BENCH_WORLDS=512 tools/r2_solve_ab.sh "0:2 0:4 0:8"
echo "---- single big worlds: per phase launches vs persistent TMA"
for mode in 0 2; do for wl in pile max_bodies; do
  B2J_SOLVE_MODE=$mode B2J_SOLVE_TMA_SHAPE=2 timeout 300 python bench.py --workload $wl --bodies 1000000 --steps 20 --warmup 100 --no-cpu-baseline > gpurun_out/big_${wl}_$mode.json 2> gpurun_out/big_${wl}_$mode.err
  python - gpurun_out/big_${wl}_$mode.json <<'PY'
import json, sys
try:
    p = json.load(open(sys.argv[1])); r = p["roofline"] or {}
    print(sys.argv[1], "ms/step %.2f e2e/value %.2f roofline %s %.3f" % (p["ms_per_step"], p["e2e"]["value"] / p["value"], r.get("kernel"), r.get("frac", 0)), {k: round(v, 2) for k, v in list(p["kernel_ms_per_step"].items())[:6]})
except Exception as e:
    print(sys.argv[1], "failed", e)
PY
done; done
python tools/diag_small.py 2>&1 | tail -14
