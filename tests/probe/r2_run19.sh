#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "b13d or host or large or precond" > gpurun_out/r2_19_pytest.log 2>&1
echo "rc=$?" >> gpurun_out/r2_19_pytest.log
B="--steps 5 --warmup 3 --no-configs --no-cpu --no-fp64-pass"
timeout 300 python bench.py $B > gpurun_out/r2_19_n1.json 2> gpurun_out/r2_19_n1.err
tail -2 gpurun_out/r2_19_pytest.log
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_19_n1.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], {k:(round(v['ms']),round(v['ms']/v['launches'],3),v['gbs']) for k,v in d['kernels'].items()})
PY
