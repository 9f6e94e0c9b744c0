#!/bin/bash
# ncu of the pipelined scalar kernel (3 modes); CSV exports only
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out /tmp/ncu
timeout 900 ncu --set full --clock-control none --import-source on -f -o /tmp/ncu/sc -k regex:"k3_scalar_tma" -c 7 python tests/probe/ncu_kernels.py 3d 512 > gpurun_out/r2_6_ncu.log 2>&1
ncu -i /tmp/ncu/sc.ncu-rep --page raw --csv > gpurun_out/r2_6_raw.csv 2>/dev/null
for id in 0 1 2; do
  ncu -i /tmp/ncu/sc.ncu-rep --page source --csv --launch-skip $id --launch-count 1 > gpurun_out/r2_6_source_$id.csv 2>/dev/null
done
du -sh gpurun_out
