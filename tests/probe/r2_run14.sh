#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "tma or precond or b13d or large" > gpurun_out/r2_15_pytest.log 2>&1
echo "rc=$?" >> gpurun_out/r2_15_pytest.log
timeout 600 python tests/probe/tune_kernels.py 512 > gpurun_out/r2_15_tune.log 2>&1
B="--steps 5 --warmup 3 --no-configs --no-cpu --no-fp64-pass"
timeout 600 python bench.py $B > gpurun_out/r2_15_bench.json 2> gpurun_out/r2_15_bench.err
tail -4 gpurun_out/r2_15_pytest.log; grep jac gpurun_out/r2_15_tune.log; cut -c1-250 gpurun_out/r2_15_bench.json; tail -2 gpurun_out/r2_15_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_15_bench.json').read().strip().splitlines()[-1])
print({k:(round(v['ms']),round(v['ms']/v['launches'],3)) for k,v in d['kernels'].items()})
print(d['roofline'])
PY
