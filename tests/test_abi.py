"""CPU-only: the C-ABI library builds/loads and exports every symbol include/pfh_b200.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib_path():
    from pfhubbenchmarks_b200 import build
    return build.build()


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "pfh_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(pfh_[a-z0-9_]+)\s*\(", hdr)))


def test_exports_every_declared_symbol(lib_path):
    lib = ctypes.CDLL(lib_path)
    syms = declared_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(lib, s), s


def test_binding_covers_header(lib_path):
    from pfhubbenchmarks_b200 import _lib
    assert sorted(_lib.SIGNATURES) == declared_symbols()
    lib = _lib.load()
    assert lib.pfh_abi_version() == _lib.ABI_VERSION == 2
    o = _lib.default_opts()
    assert o.absolute_tolerance == 1e-6 and o.maximum_iterations == 10 and o.krylov_restart == 30


def test_struct_layout_matches_c(lib_path, tmp_path):
    """sizeof() of the three ABI structs as seen by a C compiler == ctypes."""
    from pfhubbenchmarks_b200 import _lib
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "pfh_b200.h"\nint main(){printf("%zu %zu %zu\\n",'
                   'sizeof(pfh_config),sizeof(pfh_solver_opts),sizeof(pfh_solve_report));return 0;}\n')
    exe = tmp_path / "sz"
    import subprocess
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    a, b, c = map(int, subprocess.check_output([str(exe)]).split())
    assert (a, b, c) == (ctypes.sizeof(_lib.Config), ctypes.sizeof(_lib.SolverOpts), ctypes.sizeof(_lib.SolveReport))


def test_no_product_import_of_oracle():
    """The product package must never import the oracle (no CPU fallback)."""
    pkg = os.path.join(ROOT, "pfhubbenchmarks_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                txt = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f


def test_slab_range_partitions_planes():
    from pfhubbenchmarks_b200.engine import slab_range
    for n, r in [(512, 8), (51, 4), (129, 2), (10, 3)]:
        cuts = [slab_range(n, k, r) for k in range(r)]
        assert cuts[0][0] == 0 and cuts[-1][1] == n
        assert all(cuts[k][1] == cuts[k + 1][0] for k in range(r - 1))
        sizes = [b - a for a, b in cuts]
        assert max(sizes) - min(sizes) <= 1
