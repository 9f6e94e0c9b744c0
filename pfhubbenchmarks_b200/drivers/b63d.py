"""dolfin/b63d.py of the reference.

Despite its name that script is NOT a 3-D problem: it is dolfin/bench6.py on the same 100 x 100 crossed RectangleMesh of
[0,100]^2, byte for byte except for the `import matplotlib.pyplot as plt` line (bench6.py:4) and a trailing `pass`
(`diff dolfin/b63d.py dolfin/bench6.py`): a working copy from which a 3-D version of benchmark 6 was to be made, as b13d.py
was made from bench1.py.  It writes the same results/bench6_out.csv (b63d.py:246).  This driver therefore runs the
benchmark-6 model (fields c, mu, phi; Dirichlet data on phi) under the b63d name, with the same options:

    python -m pfhubbenchmarks_b200.drivers.b63d [--end-time 3.0] [--max-steps N] [--outdir DIR]
"""
from . import bench6


def main():
    return bench6.main()


if __name__ == "__main__":
    main()
