#!/bin/bash
# full capture of the default velocity solve kernel (late point part loads, 110 registers) on a resting step of 256 Pyramid worlds,
# programmatic dependent launch on (ncu serialises the kernels itself); summary + DRAM traffic against the algorithmic bytes
export B2J_BENCH_CUPROFILE=1 B2J_BATCH_GROUPS=1
BENCH="python bench.py --worlds 256 --steps 1 --warmup 110 --no-cpu-baseline --no-pile --no-extras"
ncu --set full --clock-control none --profile-from-start off --import-source on --kernel-name-base demangled -k regex:"KSolveVelocity" -c 8 -f -o gpurun_out/r2c_hot_solve_late $BENCH > gpurun_out/r2c_hot_solve_late.log 2>&1
python tools/ncu_summary.py gpurun_out/r2c_hot_solve_late.ncu-rep > gpurun_out/r2c_hot_solve_late.txt 2>&1
python tools/ncu_traffic.py gpurun_out/r2c_hot_solve_late.ncu-rep gpurun_out/r2c_hot_solve_late.log > gpurun_out/r2c_solve_traffic.json 2> gpurun_out/r2c_solve_traffic.err
head -40 gpurun_out/r2c_hot_solve_late.txt; tail -5 gpurun_out/r2c_solve_traffic.json; tail -3 gpurun_out/r2c_solve_traffic.err
