mkdir -p gpurun_out
for rs in 30 20 15 10 8; do
python bench.py --grid 256 --steps 4 --warmup 3 --no-cpu --restart $rs 2>/dev/null | python -c "
import json,sys; d=json.loads([l for l in sys.stdin if l.startswith('{')][0]); print('restart $rs: value %.3e ms/step %.1f newton %d krylov %d' % (d['value'], d['ms_per_step'], d['newton_its'], d['krylov_its']), {k:round(v['ms']) for k,v in list(d['kernels'].items())[:5]})"
done | tee gpurun_out/restart256.log
