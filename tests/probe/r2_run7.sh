#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out /tmp/ncu
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "tma or precond or b13d" > gpurun_out/r2_8_pytest.log 2>&1
echo "rc=$?" >> gpurun_out/r2_8_pytest.log
timeout 600 python tests/probe/tune_kernels.py 512 > gpurun_out/r2_8_tune.log 2>&1
timeout 900 python bench.py --steps 5 --warmup 3 --no-configs --no-cpu --no-fp64-pass > gpurun_out/r2_8_bench.json 2> gpurun_out/r2_8_bench.err
timeout 900 ncu --set full --clock-control none --import-source on -f -o /tmp/ncu/sc -k regex:"k3_scalar_tma" -c 3 python tests/probe/ncu_kernels.py 3d 512 > gpurun_out/r2_8_ncu.log 2>&1
ncu -i /tmp/ncu/sc.ncu-rep --page raw --csv > gpurun_out/r2_8_raw.csv 2>/dev/null
for id in 0 1 2; do ncu -i /tmp/ncu/sc.ncu-rep --page source --csv --launch-skip $id --launch-count 1 > gpurun_out/r2_8_source_$id.csv 2>/dev/null; done
tail -3 gpurun_out/r2_8_pytest.log; cat gpurun_out/r2_8_tune.log; cut -c1-300 gpurun_out/r2_8_bench.json; tail -3 gpurun_out/r2_8_bench.err
