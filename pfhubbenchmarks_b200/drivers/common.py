"""The time-stepping loop every reference driver repeats (dolfin/bench1.py:128-217), written once.

Kept verbatim in behaviour: w0 <- w; solve; on failure halve dt (floor dt_min), restore w <- w0 and
retry; `niters < 5` doubles dt else halves it; the row appended after each accepted step is
(t, total_free_energy, total_solute|solid_fraction); the loop overshoots end_time; the CSV is
written with '%1.10f'.  Differences: the retry loop gives up after `max_retries` at dt_min (the
reference spins forever, bench1.py:164-177), and `--backend fenics` hands over to the reference script.
"""
import argparse
import os
import runpy
import sys
import time

from .. import api as _api
from .. import dolfin as df


def parse_args(name, end_time):
    ap = argparse.ArgumentParser(description="PFHub %s (B200 backend of vpuri3/PFHubBenchmarks dolfin/%s.py)" % (name, name))
    ap.add_argument("--backend", default="b200", choices=["b200", "fenics"])
    ap.add_argument("--end-time", type=float, default=end_time)
    ap.add_argument("--max-steps", type=int, default=0, help="stop after this many accepted steps (0 = until end_time)")
    ap.add_argument("--outdir", default="results")
    ap.add_argument("--reference-dir", default=os.environ.get("PFHUB_REFERENCE", "/root/reference"))
    ap.add_argument("--atol", type=float, default=1e-6, help="snes absolute_tolerance (reference: 1e-6)")
    ap.add_argument("--quiet", action="store_true")
    ap.add_argument("--save-solution", action="store_true", help="write <outdir>/<name>/<field>.pvd + VTU per step (bench1.py:9,116-119,191)")
    return ap.parse_args()


def run_fenics(args, script):
    """--backend fenics: run the unmodified reference driver (needs a FEniCS install, absent in this image)."""
    try:
        import dolfin  # noqa: F401
    except ImportError:
        sys.exit("--backend fenics needs legacy FEniCS (import dolfin failed); use --backend b200")
    sys.path.insert(0, os.path.join(args.reference_dir, "dolfin"))
    runpy.run_path(os.path.join(args.reference_dir, "dolfin", script), run_name="__main__")


def mixed_p1(mesh, nfields):
    """W, w, w0, w_, dw: the mixed P1 space of `nfields` components and the Functions every driver builds on it
    (bench1.py:39-56)."""
    pe = df.FunctionSpace(mesh, 'P', 1).ufl_element()
    W = df.FunctionSpace(mesh, df.MixedElement([pe] * nfields))
    return W, df.Function(W), df.Function(W), df.TestFunction(W), df.TrialFunction(W)


def snes_setup(solver, atol, line_search, krylov_max_it, linear_solver='gmres'):
    """The parameter block every reference driver repeats (bench1.py:82-110)."""
    solver.parameters['nonlinear_solver'] = 'snes'
    nl = solver.parameters['snes_solver']
    nl['report'] = True
    nl['error_on_nonconvergence'] = False
    nl['absolute_tolerance'] = atol
    nl['maximum_iterations'] = 10
    nl['line_search'] = line_search
    nl['linear_solver'] = linear_solver   # 'gmres' in bench1/2/3/6 and b13d (bench1.py:98), 'bicgstab' in bench2_eta1/2
    nl['preconditioner'] = 'sor'          # mapped to the Schur/Chebyshev preconditioner
    nl['krylov_solver']['maximum_iterations'] = krylov_max_it
    nl['krylov_solver']['error_on_nonconvergence'] = False
    return nl


def time_loop(problem, solver, w_ic, dt0, dt_min, end_time, csv_path, header, max_steps=0, quiet=False, max_retries=20,
              abort_on_negative_energy=False, save_dir=None, save_fields=(0,), functionals=None):
    """`functionals`: the driver's own (total_free_energy, total_solute) callables built on df.assemble, as the
    reference scripts define them (bench1.py:121-125); None uses both values of the fused energy kernel."""
    w, w0, dt = problem.w, problem.w0, problem.dt
    rank0 = df.MPI.rank() == 0
    t = df.Constant(0.0)
    w.interpolate(w_ic)
    w0.interpolate(w_ic)
    benchmark_output = []
    iteration_count = 0
    files = {}
    if save_dir and rank0 and problem.ctx.nranks == 1:      # field dumps: `cfile << (c, t)` (b13d.py:179, bench2.py:269-274)
        from .. import io
        os.makedirs(save_dir, exist_ok=True)
        names = problem.ctx.field_names
        files = {k: io.File(os.path.join(save_dir, ('conc' if names[k] == 'c' else names[k]) + '.pvd'), 'ascii', mesh=problem.mesh,
                            name='f_%d' % k) for k in save_fields}
    dt.assign(dt0)
    t1 = time.time()
    while float(t) < end_time + df.DOLFIN_EPS:
        tprev = float(t)
        iteration_count += 1
        if rank0 and not quiet:
            print("Iteration #%d. Time: %s, dt: %s" % (iteration_count, float(t), float(dt)))
        w0.assign(w)                                          # set IC
        t.assign(tprev + float(dt))
        niters, converged = solver.solve()                    # solve
        retries = 0
        while not converged:
            if float(dt) < dt_min + 1e-8 and retries >= max_retries:
                raise RuntimeError("Newton failed at dt_min; giving up")
            retries += 1
            dt.assign(max(0.5 * float(dt), dt_min))
            t.assign(tprev + float(dt))
            w.assign(w0)
            if rank0 and not quiet:
                print("REPEATING Iteration #%d. Time: %s, dt: %s" % (iteration_count, float(t), float(dt)))
            niters, converged = solver.solve()
        if niters < 5:                                        # simple rule for adaptive timestepping
            dt.assign(2 * float(dt))
        else:
            dt.assign(max(0.5 * float(dt), dt_min))
        if functionals is None:                               # analysis
            F_total, C_total = _api.diagnostics(problem)
        else:
            F_total, C_total = functionals[0](), functionals[1]()
        benchmark_output.append([float(t), F_total, C_total])
        for k, fh in files.items():
            fh << (w.split()[k], float(t))
        if rank0 and not quiet:
            print("Converged in %d Newton iterations; C_total: %r TFE: %r" % (niters, C_total, F_total))
        if abort_on_negative_energy and F_total < 0.0:        # bench3.py:254-258
            break
        if max_steps and iteration_count >= max_steps:
            break
    spent = time.time() - t1
    if rank0:
        if not quiet:
            print("Time spent is %s" % spent)
        os.makedirs(os.path.dirname(os.path.abspath(csv_path)), exist_ok=True)
        _api.write_csv(csv_path, benchmark_output, header)      # post process
    return benchmark_output, spent
