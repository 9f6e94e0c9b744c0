#!/bin/bash
# final 1-GPU validation: whole GPU suite, smoke(), default bench line, reference arm
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/r2_final_pytest.log 2>&1
echo "rc=$?" >> gpurun_out/r2_final_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_final_smoke.log 2>&1
echo "rc=$?" >> gpurun_out/r2_final_smoke.log
( time timeout 900 python bench.py ) > gpurun_out/r2_final_bench.json 2> gpurun_out/r2_final_bench.err
( time timeout 600 python bench.py --impl reference --steps 3 --warmup 3 ) > gpurun_out/r2_final_ref.json 2> gpurun_out/r2_final_ref.err
tail -14 gpurun_out/r2_final_pytest.log; cat gpurun_out/r2_final_smoke.log; cut -c1-300 gpurun_out/r2_final_bench.json; tail -4 gpurun_out/r2_final_bench.err; cut -c1-400 gpurun_out/r2_final_ref.json; tail -4 gpurun_out/r2_final_ref.err
