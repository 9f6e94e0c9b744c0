#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
export NCCL_DEBUG=VERSION
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q -s -k "2-4" > gpurun_out/r2_11_pytest_multi.log 2>&1
echo "rc=$?" >> gpurun_out/r2_11_pytest_multi.log
B="--steps 5 --warmup 3 --no-configs --no-cpu --no-fp64-pass"
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --master-port 29555"
timeout 300 $TR --nproc-per-node 2 bench.py --gpus 2 $B > gpurun_out/r2_11_n2.json 2> gpurun_out/r2_11_n2.err
timeout 300 python bench.py $B > gpurun_out/r2_11_n1.json 2> gpurun_out/r2_11_n1.err
grep mgpu gpurun_out/r2_11_pytest_multi.log | cut -c1-250; tail -2 gpurun_out/r2_11_pytest_multi.log
for f in n2 n1; do tail -n 1 gpurun_out/r2_11_$f.json | cut -c1-260; tail -n 2 gpurun_out/r2_11_$f.err | cut -c1-200; done
