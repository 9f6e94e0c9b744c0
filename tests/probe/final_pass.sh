# round-1 final single-GPU pass: parity subset, headline bench, ncu launch list, ncu --set full of the top kernels
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-cpu"
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tma or precond or compressed" > gpurun_out/fp_tests.log 2>&1; tail -3 gpurun_out/fp_tests.log
python bench.py > gpurun_out/fp_bench.json 2> gpurun_out/fp_bench.err || exit 1
$CMD > gpurun_out/fp_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -s 3500 -c 2500 --csv --log-file gpurun_out/fp_launches.csv $CMD > gpurun_out/fp_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k3_scalar_tma -s 300 -c 4 -f -o gpurun_out/fp_scalar $CMD > gpurun_out/fp_ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_multi -s 200 -c 4 -f -o gpurun_out/fp_vec $CMD > gpurun_out/fp_ncu3.log 2>&1
ls -la gpurun_out | grep fp_
