mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_gpu_multi.py -m gpu -q -x -s 2>&1 | tail -5 | tee gpurun_out/mgpu2.log
for ov in 1 0; do
PFH_OVERLAP=$ov timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2955$ov bench.py --gpus 2 --grid 256 --steps 4 --warmup 3 2>gpurun_out/n2_$ov.err | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][0]); print('overlap $ov: value %.3e ms/step %.1f krylov %d' % (d['value'], d['ms_per_step'], d['krylov_its'])); print({k:(v['ms'],v['launches']) for k,v in d['kernels'].items() if k in ('halo','cheb_sweep','cheb_first')})"
done | tee gpurun_out/n2_overlap.log
