#!/bin/bash
# round 2, GPU call on 2 GPUs: N-GPU == 1-GPU tests, strong scaling at N = 2 with the wide-ghost preconditioner
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -s > gpurun_out/r2_4_pytest_multi.log 2>&1
echo "rc=$?" >> gpurun_out/r2_4_pytest_multi.log
B="--steps 5 --warmup 3 --no-configs --no-cpu --no-fp64-pass"
timeout 600 python bench.py $B > gpurun_out/r2_4_n1.json 2> gpurun_out/r2_4_n1.err
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555"
timeout 600 $TR bench.py --gpus 2 $B > gpurun_out/r2_4_n2.json 2> gpurun_out/r2_4_n2.err
timeout 600 $TR bench.py --gpus 2 $B --option overlap=0 > gpurun_out/r2_4_n2_noov.json 2> gpurun_out/r2_4_n2_noov.err
PFH_WIDE=0 timeout 600 $TR bench.py --gpus 2 $B > gpurun_out/r2_4_n2_narrow.json 2> gpurun_out/r2_4_n2_narrow.err
tail -12 gpurun_out/r2_4_pytest_multi.log
for f in n1 n2 n2_noov n2_narrow; do tail -n 1 gpurun_out/r2_4_$f.json | cut -c1-330; tail -n 2 gpurun_out/r2_4_$f.err; done
