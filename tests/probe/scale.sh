N=$1
mkdir -p gpurun_out
if [ "$N" = "2" ]; then timeout 500 python -m pytest tests/test_gpu_multi.py -m gpu -q -x 2>&1 | tail -2; fi
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2957$N bench.py --gpus $N 2>gpurun_out/scale_n$N.err > gpurun_out/scale_n$N.json
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/scale_n$N.json") if l.startswith("{")][0])
print("N=$N: value %.3e ms/step %.1f krylov %d e2e %.3e lines_on_stdout %d" % (d["value"], d["ms_per_step"], d["krylov_its"], d["e2e"]["value"], sum(1 for _ in open("gpurun_out/scale_n$N.json"))))
print({k:(v["ms"],v["launches"]) for k,v in list(d["kernels"].items())[:8]})
PY
