#!/bin/bash
# ncu --set full capture of the one launch velocity solve (run under gpurun): tools/r2_ncu_solve.sh <shape> <worlds> <tag>
shape=$1; worlds=$2; tag=$3
B2J_SOLVE_TMA_SHAPE=$shape B2J_BATCH_GROUPS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:solve_velocity_tma -s 14 -c 1 -o gpurun_out/$tag -f \
  python bench.py --gpus 1 --steps 4 --warmup 12 --worlds $worlds --no-pile --no-extras --no-cpu-baseline > gpurun_out/$tag.log 2>&1
echo "ncu $tag rc $?"; tail -3 gpurun_out/$tag.log | cut -c1-300
