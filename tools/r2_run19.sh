#!/bin/bash
mkdir -p gpurun_out
B2J_TRACE_EPA=1 B2J_BATCH_GROUPS=1 timeout 600 python tools/diag_landing.py 512 19 26 > gpurun_out/run19_landing.log 2>&1; grep "b2j epa" gpurun_out/run19_landing.log | grep -v "collide 0 " | head -12
B2J_TRACE_EPA=1 timeout 600 python bench.py --workload pile --steps 2 --warmup 120 --no-cpu-baseline 2>&1 >/dev/null | grep "b2j epa" | tail -2
