"""CPU oracle for the PFHub implicit time step -- TEST INFRASTRUCTURE, NOT PRODUCT.

This package is a numpy/scipy restatement of the discrete problem that the
reference (vpuri3/PFHubBenchmarks, ``dolfin/pfbase.py``, ``dolfin/bench{1,2,3,6}.py``,
``dolfin/b13d.py``) hands to FEniCS + PETSc: mixed-P1 weak forms on DOLFIN's
structured simplicial meshes, integrated with the degree-3 quadrature rules the
drivers select (``dolfin/bench1.py:14-16``), advanced with backward Euler and
solved with Newton.

The arithmetic of the reference lives in third-party, un-vendored and un-pinned
dependencies (conda-forge ``fenics`` = DOLFIN/FFC/UFL/FIAT 2019.1.x, PETSc;
``README.md:6-11`` of the reference). None of them is installable here, so the
oracle restates their published algorithms and is PINNED against the reference's
own committed result artefacts (``results/bench*_out.csv`` and the VTU field
snapshots) -- see ``tests/test_oracle_golden.py`` and ``tests/golden/``.
3-D (b13d) has no committed reference output: 3-D parity is "unpinned"
(self-consistency only) and DESIGN.md says so.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import this package, and only as the
checker / the CPU arm.  The product (``pfhubbenchmarks_b200``) never imports it.
"""
