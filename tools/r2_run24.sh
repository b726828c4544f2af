#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -x > gpurun_out/run24_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/run24_tests.log
timeout 300 python tools/diag_small.py 2>&1 | grep -E "^(convex_vs_mesh|pyramid):"
