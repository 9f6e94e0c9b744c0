"""Generate tests/golden/*.npz from the reference's committed artefacts.

Run HERE (the build container, where /root/reference is mounted):
    python -m oracle.make_golden [/root/reference]
The GPU box has no /root/reference, so the vectors travel as committed fixtures.
Nothing is computed: the arrays are the reference's own FEniCS outputs
(``results/bench*_out.csv`` rows; PointData of the VTU snapshots listed in the PVD
files), stored as float64.
"""
import os
import sys

import numpy as np

from .golden_io import read_csv, read_pvd, read_vtu_scalar


def main(ref="/root/reference"):
    res = os.path.join(ref, "results")
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
    os.makedirs(out, exist_ok=True)

    csv = {}
    for b in ("bench1", "bench2", "bench3", "bench6"):
        hdr, rows = read_csv(os.path.join(res, f"{b}_out.csv"))
        csv[b] = rows
        csv[b + "_header"] = np.array(",".join(hdr))
    np.savez_compressed(os.path.join(out, "csv.npz"), **csv)

    # bench1: all six c snapshots (results/bench1/conc.pvd:4-9)
    pvd = read_pvd(os.path.join(res, "bench1", "conc.pvd"))
    np.savez_compressed(os.path.join(out, "bench1_fields.npz"),
                        times=np.array([t for t, _ in pvd]),
                        c=np.stack([read_vtu_scalar(os.path.join(res, "bench1", f)) for _, f in pvd]))

    # bench2: snapshots 0,1 (pins the IC quirk + one step) and 27,28 (late step); fields c, eta1..4
    pvd = read_pvd(os.path.join(res, "bench2", "conc.pvd"))
    idx = [0, 1, 27, 28]
    names = ["conc", "eta1", "eta2", "eta3", "eta4"]
    fields = np.stack([np.stack([read_vtu_scalar(os.path.join(res, "bench2", f"{n}{k:06d}.vtu"))
                                 for n in names]) for k in idx])
    np.savez_compressed(os.path.join(out, "bench2_fields.npz"), index=np.array(idx),
                        times=np.array([pvd[k][0] for k in idx]), fields=fields)

    # bench6: snapshots 0,1,2,6,7 of c and phi (bench6/conc.pvd is a different run from the CSV)
    pvd = read_pvd(os.path.join(res, "bench6", "conc.pvd"))
    idx = [0, 1, 2, 6, 7]
    np.savez_compressed(
        os.path.join(out, "bench6_fields.npz"), index=np.array(idx),
        times=np.array([pvd[k][0] for k in idx]),
        c=np.stack([read_vtu_scalar(os.path.join(res, "bench6", f"conc{k:06d}.vtu")) for k in idx]),
        phi=np.stack([read_vtu_scalar(os.path.join(res, "bench6", f"phi{k:06d}.vtu")) for k in idx]))

    # mesh pin: first cells / a few coordinates of the golden bench1 mesh
    from .golden_io import read_vtu
    d = read_vtu(os.path.join(res, "bench1", "conc000000.vtu"))
    np.savez_compressed(os.path.join(out, "bench1_mesh.npz"), points=d["points"][:, :2].astype(np.float32),
                        cells=d["cells"].astype(np.int32))
    for f in sorted(os.listdir(out)):
        print(f, os.path.getsize(os.path.join(out, f)))


if __name__ == "__main__":
    main(*sys.argv[1:])
