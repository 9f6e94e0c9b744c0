#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
export NCCL_DEBUG=VERSION
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_multi.py -x -q -s -k "precond or b13d or host or 2-40-1 or 2-41-1" > gpurun_out/r2_17_pytest.log 2>&1
echo "rc=$?" >> gpurun_out/r2_17_pytest.log
B="--steps 5 --warmup 3 --no-configs --no-cpu --no-fp64-pass"
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --master-port 29555"
timeout 300 python bench.py $B > gpurun_out/r2_17_n1.json 2> gpurun_out/r2_17_n1.err
timeout 300 $TR --nproc-per-node 2 bench.py --gpus 2 $B > gpurun_out/r2_17_n2.json 2> gpurun_out/r2_17_n2.err
timeout 300 python bench.py --workload bench1 --grid 100 --steps 20 --warmup 5 --no-configs --no-cpu --no-fp64-pass > gpurun_out/r2_17_b1_100.json 2> gpurun_out/r2_17_b1_100.err
grep "mgpu\|passed\|failed" gpurun_out/r2_17_pytest.log | cut -c1-220
for f in n1 n2 b1_100; do tail -n 1 gpurun_out/r2_17_$f.json | cut -c1-260; tail -n 2 gpurun_out/r2_17_$f.err | cut -c1-200; done
