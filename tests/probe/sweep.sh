mkdir -p gpurun_out
for v in 0 1; do for sw in "2 4" "3 5" "3 6" "2 6" "4 8"; do set -- $sw
python bench.py --grid 256 --steps 4 --warmup 3 --no-cpu --pc-variant $v --sweeps-mass $1 --sweeps-shifted $2 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('variant $v sweeps $1 $2: value %.3e ms/step %.1f newton %d krylov %d' % (d['value'], d['ms_per_step'], d['newton_its'], d['krylov_its']))"
done; done | tee gpurun_out/sweep256.log
