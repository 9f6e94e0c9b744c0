mkdir -p gpurun_out
for ov in 1 0; do
PFH_OVERLAP=$ov timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2956$ov bench.py --gpus 8 --steps 5 --warmup 3 2>gpurun_out/n8_$ov.err > gpurun_out/n8_$ov.json
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/n8_$ov.json") if l.startswith("{")][0])
print("overlap $ov: value %.3e ms/step %.1f krylov %d e2e %.3e" % (d["value"], d["ms_per_step"], d["krylov_its"], d["e2e"]["value"]))
print({k:(v["ms"],v["launches"]) for k,v in d["kernels"].items()})
PY
done | tee gpurun_out/n8_overlap.log
