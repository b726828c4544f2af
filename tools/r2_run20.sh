#!/bin/bash
# A/B: register cap on the per phase velocity solve kernel (B2J_SOLVE_MINB = resident blocks of 128 threads per SM the kernel is compiled for)
mkdir -p gpurun_out
for worlds in 4096 512; do
for minb in 0 4 5; do
  B2J_SOLVE_MINB=$minb timeout 600 python bench.py --steps 20 --warmup 5 --worlds $worlds --no-pile --no-extras --no-cpu-baseline > gpurun_out/run20_w${worlds}_m$minb.json 2> gpurun_out/run20_w${worlds}_m$minb.err
  python - "$worlds" "$minb" <<'PY'
import json, sys
w, m = sys.argv[1:3]
for l in open(f"gpurun_out/run20_w{w}_m{m}.json"):
    if l.startswith("{"):
        d = json.loads(l); r = d.get("roofline") or {}
        print(f"worlds {w} minb {m}: {d['ms_per_step']:.2f} ms/step frac {r.get('frac'):.3f} solve {d['kernel_ms_per_step'].get('KSolveVelocity')}", [round(x, 1) for x in d['ms_per_step_series']][:3])
PY
done
done
