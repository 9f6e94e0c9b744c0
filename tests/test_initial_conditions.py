"""Host evaluation of the reference's initial-condition expressions (dolfin/pfbase.py:177-339) -- the path a generated
model or a mismatched benchmark takes -- against the oracle's statement of the same formulas; `h` / `dh` of pfbase."""
import numpy as np
import sympy as sp

from oracle import mesh, problems
from pfhubbenchmarks_b200 import api


def test_ic_classes_evaluate_the_reference_formulas():
    m = mesh.crossed2d(10, 8, 20.0, 16.0)
    cases = [(api.InitialConditionsBench1, problems.CahnHilliard(), (0.5, 0.05)),
             (api.InitialConditionsBench2, problems.Bench2(), (0.5, 0.05, 0.1, 1.5)),
             (api.InitialConditionsBench3, problems.Bench3(), (-0.3, 8.0, 1.0, 1.0, -1.0)),
             (api.InitialConditionsBench6, problems.Bench6(), (0.5, 0.04))]
    for cls, P, args in cases:
        ic = cls(*args, degree=2)
        v = ic.values(m.coords)
        assert v.shape == (ic.nvalues, m.nv) and ic.value_shape() == (ic.nvalues,)
        assert np.abs(v - P.ic(m)).max() < 1e-15, cls.__name__
        one = np.zeros(ic.nvalues)
        ic.eval(one, m.coords[7])                     # the UserExpression signature, one point
        assert np.array_equal(one, v[:, 7])
    # the bench2 quirk (i instead of ii in the squared term, pfbase.py:285-286) is in the host formula too
    v = api.InitialConditionsBench2(0.5, 0.05, 0.1, 1.5).values(np.array([[30.0, 40.0]]))
    x, y = 30.0, 40.0
    sq_ii = (np.cos(0.047 * x - 0.0415 * y) * np.cos(0.032 * x - 0.005 * y)) ** 2
    wrong = 0.1 * (np.cos(0.01 * x - 4.0) * np.cos(0.017 * y) + np.cos(0.12 * x) * np.cos(0.12 * y) + 1.5 * sq_ii) ** 2
    assert abs(v[2, 0] - wrong) > 0.5 * abs(v[2, 0])


def test_h_and_dh():
    from pfhubbenchmarks_b200 import pfbase
    u = sp.Symbol("u")
    assert sp.simplify(sp.diff(pfbase.h(u), u) - pfbase.dh(u)) == 0
    assert sp.simplify(pfbase.dh(u) - 30 * u ** 2 * (u - 1) ** 2) == 0
