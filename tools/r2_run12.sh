#!/bin/bash
# programmatic dependent launch of the solver phase kernels: parity, then A/B (B2J_SOLVE_PDL=0/1) at 4096 / 512 worlds and on the single worlds
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/run12_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/run12_tests.log
for worlds in 4096 512; do
  for pdl in 1 0; do
    B2J_SOLVE_PDL=$pdl timeout 600 python bench.py --steps 20 --warmup 5 --worlds $worlds --no-pile --no-extras --no-cpu-baseline > gpurun_out/run12_w${worlds}_pdl$pdl.json 2> gpurun_out/run12_w${worlds}_pdl$pdl.err
    python - "$worlds" "$pdl" <<'PY'
import json, sys
w, p = sys.argv[1:3]
for l in open(f"gpurun_out/run12_w{w}_pdl{p}.json"):
    if l.startswith("{"):
        d = json.loads(l); r = d.get("roofline") or {}
        print(f"worlds {w} pdl {p}: {d['ms_per_step']:.2f} ms/step value {d['value']/1e6:.1f}M e2e {d['e2e']['value']/1e6:.1f}M frac {r.get('frac')}")
PY
  done
done
for pdl in 1 0; do
  B2J_SOLVE_PDL=$pdl timeout 600 python bench.py --workload pile --steps 30 --warmup 120 --no-cpu-baseline > gpurun_out/run12_pile_pdl$pdl.json 2> gpurun_out/run12_pile_pdl$pdl.err
  python - "$pdl" <<'PY'
import json, sys
p = sys.argv[1]
for l in open(f"gpurun_out/run12_pile_pdl{p}.json"):
    if l.startswith("{"):
        d = json.loads(l); print(f"pile formed pdl {p}: {d['ms_per_step']:.2f} ms/step", d.get("kernel_ms_per_step"))
PY
done
