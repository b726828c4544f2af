#!/bin/bash
# A/B of the velocity solve forms with the driver's bench command (4096 Pyramid worlds, steps [5, 25)): the default kernel against the
# late point-part loads under a register budget of 128 / 96 (B2J_SOLVE_LATE=1 / 2); also at 512 worlds (the per GPU share of N = 8)
for worlds in 4096 512; do
  for late in 0 1 2; do
    echo "---- worlds $worlds B2J_SOLVE_LATE=$late"
    B2J_SOLVE_LATE=$late python bench.py --gpus 1 --steps 20 --warmup 5 --worlds $worlds --no-pile --no-extras --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('ms_per_step', round(d['ms_per_step'], 2), 'value', round(d['value'] / 1e6, 2), 'M body-steps/s, e2e', round(((d.get('e2e') or {}).get('value') or 0) / 1e6, 2), 'roofline.frac', (d.get('roofline') or {}).get('frac'), {k: v for k, v in list((d.get('kernel_ms_per_step') or {}).items())[:3]})
"
  done
done
