#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "tma or b13d" > gpurun_out/r2_13_pytest.log 2>&1
echo "rc=$?" >> gpurun_out/r2_13_pytest.log
timeout 600 python tests/probe/tune_kernels.py 512 > gpurun_out/r2_13_tune.log 2>&1
B="--steps 5 --warmup 3 --no-configs --no-cpu --no-fp64-pass"
timeout 600 python bench.py $B --option jac_variant=2 > gpurun_out/r2_13_bench_ws.json 2> gpurun_out/r2_13_bench_ws.err
timeout 600 python bench.py $B > gpurun_out/r2_13_bench.json 2> gpurun_out/r2_13_bench.err
tail -3 gpurun_out/r2_13_pytest.log; cat gpurun_out/r2_13_tune.log; for f in bench_ws bench; do cut -c1-250 gpurun_out/r2_13_$f.json; tail -2 gpurun_out/r2_13_$f.err; done
