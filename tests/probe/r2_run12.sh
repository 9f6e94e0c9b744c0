#!/bin/bash
# evidence: ncu --set full for every kernel family (3-D step, 2-D step), launch list of a short bench run
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out /tmp/ncu
K='regex:k3_|k2_|k_multi|k_norm|k_axp|k_reduce|k_jacobi'
timeout 1200 ncu --set full --clock-control none -f -o /tmp/ncu/f3 -k "$K" -c 34 python tests/probe/ncu_kernels.py 3d 512 > gpurun_out/r2_12_ncu3.log 2>&1
ncu -i /tmp/ncu/f3.ncu-rep --page raw --csv > gpurun_out/r2_12_raw3d.csv 2>/dev/null
timeout 1200 ncu --set full --clock-control none -f -o /tmp/ncu/f2 -k "$K" -c 30 python tests/probe/ncu_kernels.py 2d 4096 > gpurun_out/r2_12_ncu2.log 2>&1
ncu -i /tmp/ncu/f2.ncu-rep --page raw --csv > gpurun_out/r2_12_raw2d.csv 2>/dev/null
python bench.py --steps 2 --warmup 1 --no-configs --no-cpu --no-fp64-pass > gpurun_out/r2_12_short.json 2> gpurun_out/r2_12_short.err
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2_12_launches.csv python bench.py --steps 2 --warmup 1 --no-configs --no-cpu --no-fp64-pass > gpurun_out/r2_12_launch_run.log 2>&1
du -sh gpurun_out; wc -l gpurun_out/r2_12_*.csv
