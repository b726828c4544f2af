#!/bin/bash
# round 2 (third session): the whole GPU suite with the constraint / CollideShape kernels, compute-sanitizer memcheck over the new kernels
python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r2e_gputest.log; cat gpurun_out/r2e_gputest.log
echo "---- compute-sanitizer memcheck (constraint kernels, CollideShape / sphere / point queries, compound query path)"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_parity.py tests/test_queries.py tests/test_constraints.py -m gpu -x -q \
  -k "joints-step30 or joints-step1 or (queries_gpu and small_stack) or (queries_gpu and compound-0) or constraints_errors_gpu" > gpurun_out/r2e_sanitizer_memcheck.log 2>&1
echo "memcheck rc $?"; tail -6 gpurun_out/r2e_sanitizer_memcheck.log
