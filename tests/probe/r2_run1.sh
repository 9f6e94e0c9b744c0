#!/bin/bash
# round 2, GPU call 1: correctness of the rewritten staged kernels, kernel variants, first bench lines
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv > gpurun_out/r2_1_smi.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/r2_1_pytest_parity.log 2>&1
echo "pytest parity rc=$?" >> gpurun_out/r2_1_pytest_parity.log
timeout 900 python tests/probe/tune_kernels.py 512 > gpurun_out/r2_1_tune.log 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 --no-configs --no-cpu > gpurun_out/r2_1_bench_s4.json 2> gpurun_out/r2_1_bench_s4.err
timeout 600 python bench.py --steps 5 --warmup 3 --no-configs --no-cpu --no-fp64-pass --sweeps-shifted 5 > gpurun_out/r2_1_bench_s5.json 2> gpurun_out/r2_1_bench_s5.err
timeout 600 python bench.py --steps 5 --warmup 3 --no-configs --no-cpu --no-fp64-pass --sweeps-shifted 5 --option cheb_bounds=0 > gpurun_out/r2_1_bench_s5_generic.json 2> gpurun_out/r2_1_bench_s5_generic.err
tail -3 gpurun_out/r2_1_pytest_parity.log
tail -5 gpurun_out/r2_1_tune.log
cut -c1-400 gpurun_out/r2_1_bench_s4.json
