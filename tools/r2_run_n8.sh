#!/bin/bash
# the driver's N=8 command (one rank per GPU)
mkdir -p gpurun_out
timeout -s ABRT 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err; echo "bench N=8 rc $?"
grep '^{' gpurun_out/r2_bench_n8.json | head -c 1800; echo; tail -5 gpurun_out/r2_bench_n8.err
