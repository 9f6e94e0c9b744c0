# Chebyshev sweep counts revisited with the compressed Krylov basis (cheaper orthogonalisation shifts the optimum?)
mkdir -p gpurun_out
for sw in "3 5" "2 4" "3 4" "2 5" "2 3"; do set -- $sw
python bench.py --grid 512 --steps 5 --warmup 3 --no-cpu --sweeps-mass $1 --sweeps-shifted $2 2>/dev/null | python -c "
import json,sys; d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{')][0]); print('sweeps $1 $2: value %.3e ms/step %.1f newton %d krylov %d' % (d['value'], d['ms_per_step'], d['newton_its'], d['krylov_its']))"
done | tee gpurun_out/sweep512.log
