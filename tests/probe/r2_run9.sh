#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "tma or precond or b13d" > gpurun_out/r2_9_pytest.log 2>&1
echo "rc=$?" >> gpurun_out/r2_9_pytest.log
B="--steps 5 --warmup 3 --no-configs --no-cpu --no-fp64-pass"
timeout 900 python bench.py $B > gpurun_out/r2_9_bench.json 2> gpurun_out/r2_9_bench.err
timeout 900 python bench.py $B --option shell_async=0 > gpurun_out/r2_9_bench_sync.json 2> gpurun_out/r2_9_bench_sync.err
timeout 900 python bench.py $B --option scalar_variant=3 > gpurun_out/r2_9_bench_v3.json 2> gpurun_out/r2_9_bench_v3.err
tail -3 gpurun_out/r2_9_pytest.log
for f in bench bench_sync bench_v3; do cut -c1-260 gpurun_out/r2_9_$f.json; tail -2 gpurun_out/r2_9_$f.err; done
