#!/bin/bash
# round 2, GPU call 5: producer/consumer scalar kernel: correctness, variants, bench, 100-step trajectories (if fixtures are there)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "tma or precond or b13d or large or backtracking" > gpurun_out/r2_5_pytest.log 2>&1
echo "rc=$?" >> gpurun_out/r2_5_pytest.log
timeout 600 python tests/probe/tune_kernels.py 512 > gpurun_out/r2_5_tune.log 2>&1
timeout 900 python bench.py --steps 5 --warmup 3 --no-configs --no-cpu > gpurun_out/r2_5_bench.json 2> gpurun_out/r2_5_bench.err
ls tests/golden/b13d_5*.npz >/dev/null 2>&1 && timeout 900 python -m pytest tests/test_gpu_b13d_trajectory.py -x -q -s > gpurun_out/r2_5_traj.log 2>&1
tail -5 gpurun_out/r2_5_pytest.log; cat gpurun_out/r2_5_tune.log; cut -c1-300 gpurun_out/r2_5_bench.json; tail -3 gpurun_out/r2_5_bench.err; tail -8 gpurun_out/r2_5_traj.log 2>/dev/null
