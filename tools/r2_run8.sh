#!/bin/bash
# two GPUs: the library level multi device batch test, the bench at N=2 (torchrun), compute-sanitizer on small parity cases
timeout 600 python -m pytest tests/test_batch.py -m gpu -x -q -k "two_devices" 2>&1 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err
python - <<'PY'
import json
try:
    p = json.load(open("gpurun_out/r2_bench_n2.json"))
    print("N=2: ms/step %.2f value %.1fM e2e %.1fM launches %d clocks %s" % (p["ms_per_step"], p["value"] / 1e6, p["e2e"]["value"] / 1e6, p["gpu_launches"], p["clocks"]))
except Exception as e:
    print("N=2 bench failed", e); print(open("gpurun_out/r2_bench_n2.err").read()[-1500:])
PY
timeout 600 python bench.py --impl reference --gpus 2 --steps 20 --warmup 5 | cut -c1-300
echo "---- compute-sanitizer memcheck"
CUDA_VISIBLE_DEVICES=0 timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "pyramid-4-0-40 or small_stack-4-0-45 or zoo-step60 or convex_vs_mesh" > gpurun_out/r2_sanitizer_memcheck.log 2>&1; echo "memcheck rc $?"; tail -8 gpurun_out/r2_sanitizer_memcheck.log
echo "---- compute-sanitizer racecheck (shared memory hazards: the TMA solve kernel, the block scheduler)"
CUDA_VISIBLE_DEVICES=0 B2J_SOLVE_MODE=2 timeout 900 compute-sanitizer --tool racecheck --error-exitcode 3 python -m pytest tests/test_batch.py -m gpu -x -q -k "batch_vs_oracle_gpu and pyramid" > gpurun_out/r2_sanitizer_racecheck.log 2>&1; echo "racecheck rc $?"; tail -8 gpurun_out/r2_sanitizer_racecheck.log
