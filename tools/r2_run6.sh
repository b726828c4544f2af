#!/bin/bash
B2J_SOLVE_MODE=2 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_batch.py tests/test_snapshot.py -m gpu -x -q 2>&1 | tail -4
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for w in 512 1024 2048 4096; do echo "---- $w worlds"; B2J_BENCH_WORLDS=$w tools/r2_solve_ab.sh "2:1"; done
echo "---- 2048 / 4096 worlds, 2 groups"; B2J_BENCH_WORLDS=2048 tools/r2_solve_ab.sh "2:2 0:8"; B2J_BENCH_WORLDS=4096 tools/r2_solve_ab.sh "2:2"
