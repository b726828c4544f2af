#!/bin/bash
# launch lists of the final round 2 code (the default velocity solve = KSolveVelocityT<true>, late point part loads): every launch of
# simulation step 22 (impact) and 110 (rest) of the driver's bench window at 256 worlds, one group; per kernel totals
export B2J_BENCH_CUPROFILE=1 B2J_BATCH_GROUPS=1
for PHASE in impact:22 rest:110; do
  NAME=${PHASE%%:*}; WARM=${PHASE##*:}
  BENCH="python bench.py --worlds 256 --steps 1 --warmup $WARM --no-cpu-baseline --no-pile --no-extras"
  ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2c_launches_${NAME}.csv $BENCH > gpurun_out/r2c_launches_${NAME}.log 2>&1
  python tools/ncu_launches.py gpurun_out/r2c_launches_${NAME}.csv > gpurun_out/r2c_launches_${NAME}.txt 2>&1
  head -12 gpurun_out/r2c_launches_${NAME}.txt
done
