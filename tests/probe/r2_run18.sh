#!/bin/bash
# final code: ncu --set full of every kernel family of the 3-D step (CSV export only)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out /tmp/ncu
K='regex:k3_|k_multi|k_norm|k_axp|k_reduce|k_publish'
timeout 900 ncu --set full --clock-control none -f -o /tmp/ncu/f3 -k "$K" -c 36 python tests/probe/ncu_kernels.py 3d 512 > gpurun_out/r2_18_ncu3.log 2>&1
ncu -i /tmp/ncu/f3.ncu-rep --page raw --csv > gpurun_out/r2_18_raw3d.csv 2>/dev/null
wc -l gpurun_out/r2_18_raw3d.csv
