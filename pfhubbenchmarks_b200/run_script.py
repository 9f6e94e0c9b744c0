"""Run a reference-style driver script unchanged on the B200 backend:

    python -m pfhubbenchmarks_b200.run_script path/to/bench1.py [script arguments]

The reference drivers start with `import dolfin as df` and `from pfbase import *` (dolfin/bench1.py:2-7).  This
launcher registers `pfhubbenchmarks_b200.dolfin` and `pfhubbenchmarks_b200.pfbase` under those two module names (and an
inert `matplotlib.pyplot` if matplotlib is not installed -- the drivers import it but the stepping path never plots),
then executes the script as `__main__` from its own directory, so that its relative `results/...` paths resolve as
they do under FEniCS.  Nothing in the script is rewritten.
"""
import os
import runpy
import sys
import types


def install_aliases():
    """Make `import dolfin` / `from pfbase import *` resolve to the B200 front end in this process."""
    from . import dolfin, pfbase
    sys.modules["dolfin"] = dolfin
    sys.modules["pfbase"] = pfbase
    try:
        import matplotlib.pyplot  # noqa: F401
    except ImportError:
        mpl, plt = types.ModuleType("matplotlib"), types.ModuleType("matplotlib.pyplot")
        mpl.pyplot = plt
        sys.modules["matplotlib"], sys.modules["matplotlib.pyplot"] = mpl, plt


def run(path, argv=()):
    path = os.path.abspath(path)
    install_aliases()
    old_argv, old_cwd = sys.argv, os.getcwd()
    sys.argv = [path] + list(argv)
    os.chdir(os.path.dirname(path))
    sys.path.insert(0, os.path.dirname(path))
    try:
        return runpy.run_path(path, run_name="__main__")
    finally:
        sys.argv = old_argv
        os.chdir(old_cwd)
        sys.path.remove(os.path.dirname(path))


if __name__ == "__main__":
    if len(sys.argv) < 2:
        sys.exit(__doc__)
    run(sys.argv[1], sys.argv[2:])
