#!/bin/bash
# round 2, GPU call 2: full GPU suite, kernel variants after the first-two fix, bench with the 2-D records
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --durations=15 > gpurun_out/r2_2_pytest_gpu.log 2>&1
echo "pytest gpu rc=$?" >> gpurun_out/r2_2_pytest_gpu.log
timeout 600 python tests/probe/tune_kernels.py 512 > gpurun_out/r2_2_tune.log 2>&1
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_2_bench.json 2> gpurun_out/r2_2_bench.err
tail -25 gpurun_out/r2_2_pytest_gpu.log
cat gpurun_out/r2_2_tune.log
cut -c1-300 gpurun_out/r2_2_bench.json; tail -3 gpurun_out/r2_2_bench.err
