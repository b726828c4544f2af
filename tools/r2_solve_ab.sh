#!/bin/bash
# A/B of the velocity solve forms on the batch workload (run under gpurun): B2J_SOLVE_MODE x B2J_BATCH_GROUPS
# usage: tools/r2_solve_ab.sh "<mode>:<groups>[:<griddiv>] ..."   results -> gpurun_out/ab_<mode>_<groups>_<div>.json
mkdir -p gpurun_out
for cfg in $1; do
  IFS=: read mode groups div shape <<< "$cfg"
  div=${div:-1}
  out=gpurun_out/ab_${mode}_${groups}_${div}_${shape:-1}.json
  B2J_SOLVE_MODE=$mode B2J_BATCH_GROUPS=$groups B2J_SOLVE_GRID_DIV=$div B2J_SOLVE_TMA_SHAPE=${shape:-1} timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-pile --no-extras --no-cpu-baseline > $out 2> ${out%.json}.err
  echo "== mode $mode groups $groups div $div rc $?"
  python - "$out" <<'PY'
import json, sys
try:
    p = json.load(open(sys.argv[1]))
    r = p["roofline"] or {}
    print("  ms/step %.2f  value %.1fM  e2e %.1fM (%.2f of value, split %s)  roofline %s frac %.3f launches %s avg_us %.1f share %.2f  gpu_launches %d" % (
        p["ms_per_step"], p["value"] / 1e6, p["e2e"]["value"] / 1e6, p["e2e"]["value"] / p["value"], p["e2e"].get("forces_in__step__positions_out_ms"),
        r.get("kernel"), r.get("frac", 0), r.get("launches"), r.get("avg_launch_us", 0), r.get("share_of_step", 0), p["gpu_launches"]))
    print("  kernels", {k: round(v, 2) for k, v in list(p["kernel_ms_per_step"].items())[:8]}, "profiled", round(p["profiled_ms_per_step"], 1))
except Exception as e:
    print("  no result:", e)
    print(open(sys.argv[1].replace(".json", ".err")).read()[-800:])
PY
done
