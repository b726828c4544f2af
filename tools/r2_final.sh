#!/bin/bash
# the driver's commands on one GPU: reference arm, then the B200 arm; smoke
mkdir -p gpurun_out
timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/final_ref.json 2> gpurun_out/final_ref.err; echo "ref rc=$?"
timeout 1500 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/final_b200.json 2> gpurun_out/final_b200.err; echo "b200 rc=$?"
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/final_smoke.log 2>&1; tail -2 gpurun_out/final_smoke.log
python - <<'PY'
import json
for f in ("final_ref", "final_b200"):
    for l in open(f"gpurun_out/{f}.json"):
        if l.startswith("{"):
            d = json.loads(l)
            print(f, "value", round(d["value"] / 1e6, 2), "M", "ms/step", round(d.get("ms_per_step", 0), 2), "e2e", round((d.get("e2e") or {}).get("value", 0) / 1e6, 2), "roofline", (d.get("roofline") or {}).get("frac"), "launches", d.get("gpu_launches"))
PY
