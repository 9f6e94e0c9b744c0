#!/bin/bash
# 8-GPU box: N-GPU == 1-GPU tests at 8 / 4 / 2 ranks, strong scaling at N = 8 and 4
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q -s -k "8-44-1 or 8-66-0 or 4-40-1" > gpurun_out/r2_16_pytest_multi.log 2>&1
echo "rc=$?" >> gpurun_out/r2_16_pytest_multi.log
B="--steps 5 --warmup 3 --no-configs --no-cpu --no-fp64-pass"
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --master-port 29555"
export NCCL_DEBUG=VERSION
timeout 300 $TR --nproc-per-node 8 bench.py --gpus 8 $B > gpurun_out/r2_16_n8.json 2> gpurun_out/r2_16_n8.err
timeout 300 $TR --nproc-per-node 8 bench.py --gpus 8 $B --option overlap=0 > gpurun_out/r2_16_n8_noov.json 2> gpurun_out/r2_16_n8_noov.err
timeout 300 $TR --nproc-per-node 4 bench.py --gpus 4 $B > gpurun_out/r2_16_n4.json 2> gpurun_out/r2_16_n4.err
timeout 300 python bench.py $B > gpurun_out/r2_16_n1.json 2> gpurun_out/r2_16_n1.err
tail -8 gpurun_out/r2_16_pytest_multi.log | cut -c1-250
for f in n8 n8_noov n4 n1; do tail -n 1 gpurun_out/r2_16_$f.json | cut -c1-260; tail -n 2 gpurun_out/r2_16_$f.err | cut -c1-200; done
