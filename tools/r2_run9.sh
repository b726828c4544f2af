#!/bin/bash
# two GPUs, diagnostics: where does the N=2 bench stop, what fails in the two device batch
export PYTHONFAULTHANDLER=1
timeout 300 python -m pytest tests/test_batch.py -m gpu -x -q -k "two_devices" > gpurun_out/r2_two_devices.log 2>&1; echo "two devices rc $?"; grep -E "Error|error|assert|passed|failed" gpurun_out/r2_two_devices.log | head -20
timeout -s ABRT 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 4 --warmup 3 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err; echo "bench N=2 rc $?"
head -c 600 gpurun_out/r2_bench_n2.json; echo; grep -v "^$" gpurun_out/r2_bench_n2.err | tail -60
