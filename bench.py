#!/usr/bin/env python
"""Headline benchmark: implicit Cahn-Hilliard time steps (b13d family) on a synthetic 3-D P1 grid.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--grid 512]

* step      = one backward-Euler time step = `niters, converged = solver.solve()` plus the
              free-energy / solute reductions (dolfin/b13d.py:149-186), reference solver settings
              (atol 1e-6, 10 Newton its, Krylov rtol 1e-5), dt list 0.1,0.2,0.4,0.8,1.6,1.6,...
* workload  = b13d on a Kuhn grid with `grid` vertices per edge (default 512^3 = 2.68e8 DOF), h = 2,
              bench1 initial condition; for N > 1 the SAME grid is z-slab sharded (strong scaling).
* value     = DOF x Newton iterations / s over the K timed steps, inputs resident in HBM,
              CUDA events on the ctx stream, barrier + synchronize on both sides, max over ranks.
* e2e       = the same steps through the host-buffer C-ABI call pfh_step_host (pinned host arrays,
              H2D of w and w0 and D2H of w inside the timed region).
* roofline  = the kernel family with the largest share of the timed region: algorithmic bytes
               / mean launch duration (CUDA events inside the library).
* cpu_baseline / --impl reference = the oracle port (numpy/scipy restatement of the FEniCS path;
              FEniCS itself is not installable here) timed on the host on a bounded sample grid.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# NCCL_DEBUG is left as the launcher set it (the harness reads the rank count from NCCL's own log); whatever NCCL prints
# on stdout comes first, the JSON line is the LAST line rank 0 prints.

METRIC = "CH DOF*Newton-steps/s (b13d 3-D Cahn-Hilliard, implicit step)"
UNIT = "DOF*Newton-steps/s"


def dt_schedule(n):
    """0.1, 0.2, 0.4, 0.8, 1.6, 1.6, ... (the doubling schedule capped at 1.6)."""
    return [min(0.1 * 2 ** k, 1.6) for k in range(n)]


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag, self.proc = index, [], False, None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                if self.stop_flag:
                    break
                self.samples.append([x.strip() for x in line.split(",")])
        except Exception:
            pass

    def finish(self):
        self.stop_flag = True
        if self.proc is not None:
            self.proc.terminate()
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            try:
                sm.append(float(s[0])); mx = max(mx, float(s[1]))
                for n, v in zip(names, s[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle port on a bounded sample of the same workload
# ------------------------------------------------------------------------------------------------
def oracle_steps(sample_grid, n_steps, skip=0, n_m=3, n_a=4):
    """Run backward-Euler steps of the oracle (assemble + Newton + preconditioned GMRES, reference solver
    tolerances) on a Kuhn grid with `sample_grid` vertices per edge. Returns (dof, newton_its list, seconds list)."""
    import numpy as np
    from oracle import krylov, mesh, problems, stepper
    n = sample_grid - 1
    m = mesh.kuhn3d(n, n, n, 2.0 * n, 2.0 * n, 2.0 * n)
    P = problems.CahnHilliard()
    D = stepper.Discretization(m, P)
    u = P.ic(m)
    its, secs = [], []
    for k, dt in enumerate(dt_schedule(skip + n_steps)):
        t0 = time.perf_counter()
        u, nit, ok = D.solve_step(u, u, dt, atol=1e-6, rtol=1e-9, max_it=10,
                                  linear=krylov.make_linear(rtol=1e-5, n_m=n_m, n_a=n_a, variant="upper", symbol_lmax=True))
        E, C = D.energy(u), D.mass(u)
        t1 = time.perf_counter()
        assert ok and np.isfinite(E) and np.isfinite(C)
        if k >= skip:
            its.append(nit); secs.append(t1 - t0)
    return 2 * m.nv, its, secs


def _oracle_worker(job):
    """One replica of the CPU arm (runs in its own process, pinned to nothing: the OS spreads the replicas over the cores)."""
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    grid, n_steps, skip, n_m, n_a = job
    dof, its, secs = oracle_steps(grid, n_steps, skip, n_m, n_a)
    return dof, its, secs


def host_cores():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def oracle_all_cores(sample_grid, n_steps, skip=0, n_m=3, n_a=4, cores=None):
    """The CPU arm on all host cores: the oracle's solver is a single-threaded scipy code (as one FEniCS/PETSc rank is),
    so one replica per core runs the same sample concurrently -- the analogue of `mpirun -np <cores>` for a code with no
    MPI here.  Returns (aggregate DOF*Newton-steps/s, cores, mean seconds per step of a replica, newton its of a replica)."""
    import multiprocessing as mp
    cores = cores or host_cores()
    if cores == 1:
        res = [_oracle_worker((sample_grid, n_steps, skip, n_m, n_a))]
    else:
        with mp.get_context("spawn").Pool(cores) as pool:
            res = pool.map(_oracle_worker, [(sample_grid, n_steps, skip, n_m, n_a)] * cores)
    value = sum(dof * sum(its) / sum(secs) for dof, its, secs in res)
    step_s = sum(sum(secs) for _, _, secs in res) / (len(res) * n_steps)
    return value, cores, step_s, sum(res[0][1])


def cpu_baseline(sample_grid=48, n_steps=2, n_m=3, n_a=4):
    value, cores, step_s, _ = oracle_all_cores(sample_grid, n_steps, 0, n_m, n_a)
    return {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "oracle port (numpy/scipy: assembled P1 operators, Newton, Schur-Chebyshev(%d,%d) GMRES, the product's algorithm), "
                      "%d^3-vertex Kuhn grid, first %d steps of the dt list, one single-threaded replica per host core (%d), "
                      "aggregate throughput" % (n_m, n_a, sample_grid, n_steps, cores)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    value, cores, step_s, _ = oracle_all_cores(args.ref_grid, args.steps, args.warmup, args.sweeps_mass, args.sweeps_shifted)
    sample = ("oracle port (numpy/scipy restatement; FEniCS/PETSc not installable here), %d^3-vertex Kuhn grid sample of the "
              "b13d workload, steps %d..%d of the dt list, one single-threaded replica per host core (%d), aggregate throughput"
              % (args.ref_grid, args.warmup, args.warmup + args.steps - 1, cores))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * step_s, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "b13d_kuhn3d_%d" % args.grid, "sample_grid": args.ref_grid, "fields": "c,mu"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def bench_case(args, workload, grid, steps, warmup, dist_env, e2e=True, fp64_pass=False):
    """W warm-up + K timed steps of one workload on its own ctx; returns the measured quantities of this rank's view
    (times already reduced to the maximum over ranks)."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from pfhubbenchmarks_b200 import _lib, engine
    rank, world, local, nccl_id = dist_env

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if workload != "b13d":       # 2-D crossed mesh with `grid` cells per edge (reference: 100 / 350), replicas only
        assert world == 1, "2-D cases do not shard (DESIGN.md section 6)"
        n = grid
        h = {"bench1": 2.0, "bench2": 2.0, "bench3": 960.0 / 350.0, "bench6": 1.0}[workload]   # reference cell sizes
        ctx = engine.Context(workload, (n, n), (h * n, h * n), device=local, krylov_basis=args.restart)
    else:
        n = grid - 1
        L = 2.0 * n
        ctx = engine.Context("b13d", (n, n, n), (L, L, L), device=local, rank=rank, nranks=world, nccl_id=nccl_id,
                             krylov_basis=args.restart)
    dof = ctx.nf * ctx.n_global
    if world > 1 and os.environ.get("PFH_WIDE", "1") == "0":
        ctx.set_option("use_wide", 0)
    for kv in args.option:
        k, v = kv.split("=")
        ctx.set_option(k, float(v))
    ctx.set_option("basis_precision", {"auto": 0, "fp64": 1, "fp32": 2}[args.basis])
    if args.jac_fp32:
        ctx.set_option("jac_quadrature_fp32", 1)
    opts = _lib.default_opts(krylov_restart=args.restart, pc_variant=args.pc_variant, pc_sweeps_mass=args.sweeps_mass,
                             pc_sweeps_shifted=args.sweeps_shifted, krylov_maximum_iterations=max(1000, int(n ** 3 / 10)) if workload == "b13d" else 1000,
                             line_search=1 if workload in ("bench2", "bench6") else 0)
    stream = torch.cuda.ExternalStream(ctx.lib.pfh_stream(ctx._h), device=torch.device("cuda", local))
    dts = dt_schedule(warmup + steps)
    if workload in ("bench2", "bench3", "bench6"):      # these drivers start at dt = 1e-2 (bench2.py:222)
        # capped: the reference's own run rejects dt = 0.16 at t = 0.15 for bench2 (results/bench2_out.csv rows 4-5), and on a
        # 2048^2 mesh Newton with the 'cp' search already fails at 0.08 (first step at that dt: 6 iterations at 512^2 and
        # 1024^2, a stagnating linear solve at 2048^2, profiles/r2_logs/r2_21_bench2.log) -- the drivers' loop would halve dt there
        cap = 0.04 if workload == "bench2" else 0.08
        dts = [min(0.01 * 2 ** k, cap) for k in range(warmup + steps)]

    w = ctx.initial_condition()
    w0 = torch.empty_like(w)

    def one_step(dt):
        w0.copy_(w)                      # w0.assign(w)       (b13d.py:146)
        ctx.dt = dt                      # t = tprev + dt
        rep = ctx.newton_solve(w, w0, opts=opts)    # niters, converged = solver.solve()
        E, C = ctx.energy(w)             # total_free_energy, total_solute
        rep["energy"], rep["solute"] = E, C
        return rep

    for k in range(warmup):
        rep = one_step(dts[k])
        assert rep["converged"], rep
    # state at the start of the timed region, kept on the host for the e2e pass
    start_host = torch.empty(w.shape, dtype=w.dtype, pin_memory=True)
    start_host.copy_(w)
    torch.cuda.synchronize()

    # ---- timed region: K steps, inputs resident in HBM ----
    # The timed region runs without the library's per-launch CUDA events (two cudaEventRecord per kernel cost the GPU a
    # few microseconds each, which is visible once a slab's kernels take 100 us); the per-family kernel table comes from
    # an identical second pass with the events on.
    bytes_vec = 8 * ctx.nf * ctx.n_owned
    profile_in_timed = False

    def timed_pass(profile):
        w.copy_(start_host, non_blocking=False)
        torch.cuda.synchronize()
        ctx.timers(reset=True)
        ctx.set_option("profiling", 1 if profile else 0)
        ctx.set_option("reset_launch_count", 0)
        reps = []
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for k in range(steps):
            reps.append(one_step(dts[warmup + k]))
        e1.record(stream)
        barrier()
        t_ms = e0.elapsed_time(e1)
        n_launch = ctx.launch_count()
        tm = ctx.timers(reset=True)
        ctx.set_option("profiling", 0)
        return t_ms, n_launch, tm, reps

    def allmax(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(x):
        if world == 1:
            return x
        t = torch.tensor([float(x)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    sampler = ClockSampler(local)
    sampler.start()
    ms, launches, timers, reports = timed_pass(profile_in_timed)
    clocks = sampler.finish()
    if not profile_in_timed:
        ms_prof, _, timers, _ = timed_pass(True)
    else:
        ms_prof = ms
    assert all(r["converged"] for r in reports), reports
    out = {"dof": dof, "nf": ctx.nf, "fields": ",".join(ctx.field_names), "bytes_vec": bytes_vec, "dts": dts[warmup:],
           "newton": sum(r["niters"] for r in reports), "krylov": sum(r["krylov_its"] for r in reports),
           "energy": reports[-1]["energy"], "solute": reports[-1]["solute"], "ms": allmax(ms), "ms_prof": ms_prof,
           "launches": int(allsum(launches)), "timers": timers, "clocks": clocks, "profile_in_timed": bool(profile_in_timed)}

    # ---- the same K steps with every vector stored in fp64 (the Krylov basis included) ----
    if fp64_pass and args.basis != "fp64":
        ctx.set_option("basis_precision", 1)
        ms64, _, _, reps64 = timed_pass(False)
        ctx.set_option("basis_precision", {"auto": 0, "fp64": 1, "fp32": 2}[args.basis])
        assert all(r["converged"] for r in reps64), reps64
        out["fp64"] = {"ms": allmax(ms64), "newton": sum(r["niters"] for r in reps64), "krylov": sum(r["krylov_its"] for r in reps64),
                       "energy": reps64[-1]["energy"]}

    # ---- e2e: the same K steps through the host-buffer plugin call ----
    if e2e:
        w.copy_(start_host)
        h_w = start_host.numpy()
        h_w0 = torch.empty(w.shape, dtype=w.dtype, pin_memory=True).numpy()
        e2e_ms, e2e_newton = 0.0, 0
        for k in range(steps):
            np.copyto(h_w0, h_w)             # w0.assign(w) on the host, outside the timed call
            barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            rep = ctx.step_host(h_w, h_w0, dts[warmup + k], opts=opts)
            b.record(stream)
            barrier()
            e2e_ms += a.elapsed_time(b)
            e2e_newton += rep["niters"]
            assert rep["converged"], rep
        out["e2e_ms"], out["e2e_newton"] = allmax(e2e_ms), e2e_newton
    ctx.close()
    del w, w0
    torch.cuda.empty_cache()
    return out


SCALAR_MODES = ("cheb_sweep", "cheb_first", "lin_apply")   # the three modes of ONE kernel, k3_scalar_tma / k2_scalar_tiled


def kernel_table(res):
    fam = {k: v for k, v in res["timers"].items() if v["launches"] > 0}
    total = sum(v["ms"] for v in fam.values())
    ours = {k: dict(v) for k, v in fam.items() if k not in ("halo", "allreduce", "copy")}
    # the scalar stencil kernel appears as three timer families (its modes): as a kernel it is one entry
    modes = [ours.pop(k) for k in SCALAR_MODES if k in ours]
    if modes:
        ours["scalar_stencil"] = {key: sum(m[key] for m in modes) for key in ("ms", "launches", "bytes")}
    dom = max(ours, key=lambda k: ours[k]["ms"])
    kernels = {k: {"ms": round(v["ms"], 3), "launches": v["launches"],
                   "gbs": round(v["bytes"] / (v["ms"] * 1e-3) / 1e9, 1) if v["ms"] > 0 else None,
                   "share": round(v["ms"] / res["ms_prof"], 4)} for k, v in sorted(fam.items(), key=lambda kv: -kv[1]["ms"])}
    return dom, ours[dom], kernels, total / res["ms_prof"], ours


def workload_name(workload, grid):
    return ("b13d_kuhn3d_%d" % grid) if workload == "b13d" else "%s_crossed2d_%d" % (workload, grid)


FP64_INSTR_PER_VERTEX_JAC = 350.0     # fp64 instructions per vertex of k3_jac_tma<CH> (ncu instruction counts by opcode: 355 before the
                                      # interior-only vertex phase saved ~6)


def roofline_of(res, name, world):
    """`roofline`: the dominant HBM-bound kernel (the scalar stencil kernel in its three modes, or whichever family is
    larger) against the measured copy bandwidth; `whole_step`: algorithmic bytes of every family over the step's wall
    time; `fp64_bound`: the J.v family, which is bound by the fp64 pipe and not by HBM, against 64 fp64 lanes/clk/SM at
    the SM clock sampled during the run."""
    peak, peak_src = measured_peak()
    dom, d, kernels, cover, ours = kernel_table(res)
    hbm = {k: v for k, v in ours.items() if k not in ("jac_apply", "residual", "energy", "reduce_final", "initial_condition", "other")}
    if hbm:
        dom = max(hbm, key=lambda k: hbm[k]["ms"])
        d = hbm[dom]
    achieved = d["bytes"] / (d["ms"] * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath) and world == 1:
        try:
            traffic = json.load(open(tpath)).get(name, {}).get(dom)
        except Exception:
            traffic = None
    # whole-step figure: algorithmic bytes of every family of the timed region over its wall time
    tot_bytes = sum(v["bytes"] for k, v in res["timers"].items() if k not in ("halo", "allreduce"))
    stencil = "k3_scalar_tma" if name.startswith("b13d") else "k2_scalar_tiled / k2_scalar"
    roof = {"bound": "hbm", "kernel": dom if dom != "scalar_stencil" else "scalar_stencil (%s: cheb_sweep + cheb_first + lin_apply)" % stencil,
            "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": traffic, "peak_source": peak_src, "bytes_per_launch": d["bytes"] / d["launches"],
            "ms_per_launch": d["ms"] / d["launches"], "share_of_step": d["ms"] / res["ms_prof"],
            "timed_region_profiled": res["profile_in_timed"],
            "whole_step": {"algorithmic_bytes": tot_bytes, "gbs": tot_bytes / (res["ms_prof"] * 1e-3) / 1e9,
                           "frac": tot_bytes / (res["ms_prof"] * 1e-3) / 1e9 / peak}}
    j = ours.get("jac_apply")
    if j and name.startswith("b13d") and res["clocks"].get("sm_mhz"):
        nvert = res["dof"] / res["nf"] / max(world, 1)
        cycles = FP64_INSTR_PER_VERTEX_JAC * nvert / (64.0 * 148.0)
        avail = res["clocks"]["sm_mhz"] * 1e6 * (j["ms"] / j["launches"]) * 1e-3
        roof["fp64_bound"] = {"kernel": "jac_apply (k3_jac_tma)", "fp64_instr_per_vertex": FP64_INSTR_PER_VERTEX_JAC,
                              "ms_per_launch": j["ms"] / j["launches"], "share_of_step": j["ms"] / res["ms_prof"],
                              "frac_of_fp64_pipe": cycles / avail, "peak": "64 fp64 lanes/clk/SM x 148 SMs at the sampled SM clock"}
    return roof, kernels, cover


def run_b200(args):
    import torch
    import torch.distributed as dist
    from pfhubbenchmarks_b200 import _lib, engine

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus, "launch with torchrun --nproc-per-node N for --gpus N"
    # the CPU leg first: its worker processes are spawned before this process holds a CUDA context
    cpu = cpu_baseline(args.cpu_grid, 2, args.sweeps_mass, args.sweeps_shifted) if (world == 1 and rank == 0 and not args.no_cpu) else None
    torch.cuda.set_device(local)
    nccl_id = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        buf = torch.zeros(_lib.NCCL_ID_BYTES, dtype=torch.uint8, device="cuda")
        if rank == 0:
            buf.copy_(torch.frombuffer(bytearray(engine.nccl_unique_id()), dtype=torch.uint8))
        dist.broadcast(buf, 0)
        nccl_id = bytes(buf.cpu().numpy().tobytes())
    env = (rank, world, local, nccl_id)

    res = bench_case(args, args.workload, args.grid, args.steps, args.warmup, env, e2e=True, fp64_pass=not args.no_fp64_pass)
    name = workload_name(args.workload, args.grid)
    dof, ms = res["dof"], res["ms"]
    value = dof * res["newton"] / (ms * 1e-3)
    e2e_value = dof * res["e2e_newton"] / (res["e2e_ms"] * 1e-3)
    roofline, kernels, cover = roofline_of(res, name, world)
    stored = "fp64" if args.basis == "fp64" else "fp32"

    # ---- the other sizes BASELINE.json's metric names (bench1 2-D), one record each; single GPU only ----
    extra = []
    if world == 1 and args.workload == "b13d" and not args.no_configs:
        for wl, grid in [("bench1", 1024), ("bench1", 4096), ("bench1", 8192)]:
            r2 = bench_case(args, wl, grid, min(args.steps, 3), 3, env, e2e=False)
            roof2, kern2, _ = roofline_of(r2, workload_name(wl, grid), 1)
            extra.append({"workload": workload_name(wl, grid), "dof": r2["dof"], "steps": min(args.steps, 3), "warmup": 3,
                          "value": r2["dof"] * r2["newton"] / (r2["ms"] * 1e-3), "unit": UNIT, "ms_per_step": r2["ms"] / min(args.steps, 3),
                          "newton_its": r2["newton"], "krylov_its": r2["krylov"], "energy": r2["energy"], "solute": r2["solute"],
                          "roofline": {k: roof2[k] for k in ("kernel", "achieved", "peak", "frac", "share_of_step", "whole_step")}})

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms / args.steps,
                # results of the run first: the N-GPU lines must reproduce the N = 1 values (same grid, same steps)
                "newton_its": res["newton"], "krylov_its": res["krylov"], "energy": res["energy"], "solute": res["solute"],
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f64 arithmetic; Krylov basis stored %s, everything else fp64" % stored, "data": "synthetic",
                "config": {"workload": name,
                           "grid_vertices": [args.grid] * 3 if args.workload == "b13d" else [args.grid + 1] * 2, "dof": dof,
                           "fields": res["fields"], "dt_list": res["dts"], "parallelism": "z-slabs x%d" % world,
                           "solver": "newton(atol 1e-6, max 10) + GMRES(%d, rtol 1e-5, basis stored %s) + Schur/Chebyshev(%d,%d,%s)" % (
                               args.restart, stored, args.sweeps_mass, args.sweeps_shifted, "upper" if args.pc_variant else "full"),
                           "l2_policy": ("inputs_exceed_l2 (every vector is %.0f MB >> 126 MB L2)" % (res["bytes_vec"] / 1e6)) if res["bytes_vec"] > 2.6e8
                           else "L2-resident working set (%.1f MB per vector): launch-bound regime, no flush" % (res["bytes_vec"] / 1e6)},
                "gpu_launches": res["launches"], "clocks": res["clocks"],
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 2 * res["bytes_vec"] * world,
                        "d2h_bytes_per_step": res["bytes_vec"] * world + 32, "ms_per_step": res["e2e_ms"] / args.steps},
                "roofline": roofline, "kernels": kernels, "kernel_time_coverage": cover}
        if "fp64" in res:
            f = res["fp64"]
            line["value_fp64_storage"] = dof * f["newton"] / (f["ms"] * 1e-3)
            line["fp64_storage"] = {"ms_per_step": f["ms"] / args.steps, "newton_its": f["newton"], "krylov_its": f["krylov"],
                                    "energy": f["energy"]}
        if cpu is not None:
            line["cpu_baseline"] = cpu
        if extra:
            line["configs"] = extra
        sys.stdout.flush()
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--grid", type=int, default=512, help="vertices per edge of the synthetic 3-D grid (cells per edge for bench1)")
    ap.add_argument("--workload", default="b13d", choices=["b13d", "bench1", "bench2", "bench3", "bench6"],
                    help="b13d = 3-D headline; bench1/2/3/6 = the 2-D problems on a crossed mesh with `grid` cells per edge")
    ap.add_argument("--restart", type=int, default=30)
    ap.add_argument("--pc-variant", type=int, default=1)
    ap.add_argument("--sweeps-mass", type=int, default=3)
    ap.add_argument("--sweeps-shifted", type=int, default=4)
    ap.add_argument("--cpu-grid", type=int, default=48, help="sample grid of the cpu_baseline leg")
    ap.add_argument("--ref-grid", type=int, default=48, help="sample grid of --impl reference")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the bench1 2-D records (1024^2, 4096^2, 8192^2)")
    ap.add_argument("--no-fp64-pass", action="store_true", help="skip the second timed pass with the Krylov basis stored in fp64")
    ap.add_argument("--option", action="append", default=[], help="library tuning option key=value (pfh_set_option), repeatable")
    ap.add_argument("--basis", default="auto", choices=["auto", "fp64", "fp32"],
                    help="Krylov basis storage (auto: fp32 at the default linear tolerance, DESIGN.md section 5)")
    ap.add_argument("--jac-fp32", action="store_true",
                    help="experimental: quadrature block of the Krylov J.v in single precision (DESIGN.md section 9, not the default)")
    args = ap.parse_args()
    if args.warmup < 3:
        print("warning: timing rules ask for >= 3 warm-up steps", file=sys.stderr)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
