#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -x > gpurun_out/run23_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/run23_tests.log
timeout 300 python tools/diag_small.py > gpurun_out/run23_small.log 2>&1; grep -E "^(convex_vs_mesh|pyramid):|b2j step us" gpurun_out/run23_small.log | cut -c1-260
B2J_SOLVE_PDL=0 timeout 300 python tools/diag_small.py 2>&1 | grep -E "^(convex_vs_mesh|pyramid):"
