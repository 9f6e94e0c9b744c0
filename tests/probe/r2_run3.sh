#!/bin/bash
# round 2, GPU call: ncu --set full of one launch of every kernel family of the 3-D step; only CSV exports travel back
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out /tmp/ncu
python tests/probe/ncu_kernels.py 3d 512 > gpurun_out/r2_3_plain.log 2>&1 || exit 1
# launch order of the script: residual, [first2, sweep, lin, first2, sweep, sweep, lin, first2, sweep, sweep], J.v, multidot, multiaxpy ...
timeout 1200 ncu --set full --clock-control none --import-source on -f -o /tmp/ncu/full3d \
   -k regex:"k3_scalar_tma|k3_jac_tma|k_multi|k3_energy|k_norm|k_axp" -c 22 python tests/probe/ncu_kernels.py 3d 512 > gpurun_out/r2_3_ncu.log 2>&1
ls -la /tmp/ncu
ncu -i /tmp/ncu/full3d.ncu-rep --page raw --csv > gpurun_out/r2_3_raw.csv 2>/dev/null
ncu -i /tmp/ncu/full3d.ncu-rep --page details --csv > gpurun_out/r2_3_details.csv 2>/dev/null
# source pages (SASS + source correlation) of the kernels under work: ids are positions in the capture
for id in 0 1 2 3 11; do
  ncu -i /tmp/ncu/full3d.ncu-rep --page source --csv --print-source sass --kernel-id ::: 2>/dev/null | head -0
done
python - <<'PY'
import subprocess, csv, io
raw = subprocess.run(["ncu", "-i", "/tmp/ncu/full3d.ncu-rep", "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
idi, ni = hdr.index("ID"), hdr.index("Kernel Name")
seen = {}
for r in rows[2:]:
    name = r[ni]
    key = name[:60]
    if key not in seen:
        seen[key] = r[idi]
print(seen)
for key, kid in seen.items():
    out = subprocess.run(["ncu", "-i", "/tmp/ncu/full3d.ncu-rep", "--page", "source", "--csv", "--launch-skip", kid, "--launch-count", "1"],
                         capture_output=True, text=True).stdout
    open("gpurun_out/r2_3_source_%s.csv" % kid, "w").write(out)
PY
du -sh gpurun_out; ls -la gpurun_out | head -30
