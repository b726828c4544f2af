#!/bin/bash
# two GPUs: the library level two device batch test, then the driver's N=2 bench command
export PYTHONFAULTHANDLER=1
timeout 300 python -m pytest tests/test_batch.py -m gpu -x -q -k "two_devices" > gpurun_out/r2_two_devices.log 2>&1; echo "two devices rc $?"; tail -5 gpurun_out/r2_two_devices.log
timeout -s ABRT 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err; echo "bench N=2 rc $?"
grep '^{' gpurun_out/r2_bench_n2.json | head -c 1500; echo
