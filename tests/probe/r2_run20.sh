#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
export NCCL_DEBUG=VERSION
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q -s -k "2-40-1 or 2-40-0 or 2-41-1" > gpurun_out/r2_20_pytest.log 2>&1
echo "rc=$?" >> gpurun_out/r2_20_pytest.log
B="--steps 5 --warmup 3 --no-configs --no-cpu --no-fp64-pass"
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --master-port 29555"
timeout 300 $TR --nproc-per-node 2 bench.py --gpus 2 $B > gpurun_out/r2_20_n2.json 2> gpurun_out/r2_20_n2.err
grep "mgpu world\|passed\|failed\|rc=" gpurun_out/r2_20_pytest.log | cut -c1-220
tail -n 1 gpurun_out/r2_20_n2.json | cut -c1-260; tail -n 2 gpurun_out/r2_20_n2.err | cut -c1-200
