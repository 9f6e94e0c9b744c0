/*
 * pfh_b200.h -- C ABI of the B200-native PFHub implicit time step.
 *
 * This is the drop-in boundary for the hot path of vpuri3/PFHubBenchmarks: everything
 * the reference drivers do behind
 *
 *     niters, converged = solver.solve()                 dolfin/bench1.py:162,177
 *     F_total = df.assemble(f_chem*dx + kappa/2*|grad c|^2*dx)   dolfin/bench1.py:124-125
 *     C_total = df.assemble(c*dx)                                dolfin/bench1.py:121-122
 *
 * i.e. FFC-generated element kernels + DOLFIN assembly of the mixed-P1 residual/Jacobian
 * (dolfin/pfbase.py:361-460) + PETSc SNES/KSP.  The library owns no field data: the caller
 * passes fp64 buffers laid out SoA, one contiguous array per field in DOLFIN vertex order
 * (2-D crossed mesh: grid corners x-fastest, then cell centres; 3-D BoxMesh: x, then y,
 * then z), fields concatenated: w = [field0 | field1 | ...], each of length n_owned.
 * "dev" entry points take device pointers on the ctx device; "host" entry points take host
 * pointers and do the host<->device copies themselves (the reference-facing plugin call).
 *
 * All functions return 0 on success, <0 on error (pfh_last_error() gives the message).
 * Numerical non-convergence is NOT an error: it is report.converged == 0, because the
 * reference drivers rely on it for their dt-halving retry (dolfin/bench1.py:164-177,
 * error_on_nonconvergence=False at :86).
 *
 * Threading: one host thread per ctx; a ctx is not thread-safe, distinct ctxs are independent.
 */
#ifndef PFH_B200_H
#define PFH_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define PFH_ABI_VERSION 2

/* mesh kinds: df.RectangleMesh(..., 'crossed') (dolfin/bench1.py:23), df.BoxMesh (dolfin/b13d.py:26),
 * df.UnitSquareMesh(Nx, Ny) with DOLFIN's default right-diagonal split (dolfin/nl.py:20; grid vertices only, x fastest) */
enum { PFH_MESH_RIGHT2D = 1, PFH_MESH_CROSSED2D = 2, PFH_MESH_KUHN3D = 3 };

/* problems = the weak forms the drivers build from dolfin/pfbase.py:361-460 */
enum {
  PFH_PROBLEM_CH = 1,     /* bench1.py / b13d.py : fields (c, mu);  params = {c_alpha, c_beta, rho_s, kappa, M} */
  PFH_PROBLEM_BENCH2 = 2, /* bench2.py : (c, mu, eta1..4); params = {c_alpha, c_beta, rho^2, kappa_c, kappa_eta, M, w, alpha, L} */
  PFH_PROBLEM_BENCH3 = 3, /* bench3.py : (U, phi); params = {W0, tau0, D, lambda, Lx*Ly} */
  PFH_PROBLEM_BENCH6 = 6, /* bench6.py : (c, mu, phi); params = {c_alpha, c_beta, rho, kappa, M, k, epsilon} */
  /* nl.py : one field u, STEADY (w0 and dt are ignored): F = int src v - int kc sqrt(u u) grad u . grad v - int reac u v
   * (nl.py:37), u = g on x = 0 and x = Lx (nl.py:31-35); params = {src, kc, reac, g} = {5, 1, 1, 0.9}.  Only on
   * PFH_MESH_RIGHT2D, and that mesh kind only carries this problem.  No energy functional (pfh_energy_* is an error). */
  PFH_PROBLEM_NLDIFF = 7,
  /* A model generated from the driver's own weak form (pfhubbenchmarks_b200/codegen.py: sympy -> functor) and compiled
   * into its own copy of this library; only that copy accepts it.  No params: the coefficients are in the code. */
  PFH_PROBLEM_USER = 100
};

#define PFH_MAX_PARAMS 16
#define PFH_NCCL_ID_BYTES 128

typedef struct pfh_ctx pfh_ctx;

typedef struct {
  int mesh_kind;                 /* PFH_MESH_* */
  int problem;                   /* PFH_PROBLEM_* */
  int ncell[3];                  /* cells per direction (Nx, Ny, Nz); ncell[2] ignored in 2-D */
  double length[3];              /* Lx, Ly, Lz */
  int n_params;
  double params[PFH_MAX_PARAMS];
  int device;                    /* CUDA device ordinal */
  /* z-slab decomposition (3-D only; 2-D and single-GPU: rank=0, nranks=1, z range = all planes) */
  int rank, nranks;
  int z_begin, z_end;            /* owned vertex planes [z_begin, z_end) out of ncell[2]+1 */
  const void* nccl_unique_id;    /* PFH_NCCL_ID_BYTES bytes from pfh_nccl_unique_id(), or NULL if nranks==1 */
  void* stream;                  /* cudaStream_t to enqueue on, or NULL: the ctx creates its own */
  int krylov_basis;              /* GMRES restart length the workspace is sized for (PETSc default 30) */
} pfh_config;

/* mirrors solver.parameters['snes_solver'][...] of the drivers (dolfin/bench1.py:82-110) */
typedef struct {
  double absolute_tolerance;     /* ||F||_2 < atol            (drivers: 1e-6) */
  double relative_tolerance;     /* ||F|| < rtol ||F_0||      (DOLFIN default 1e-9) */
  double solution_tolerance;     /* ||dw|| < stol ||w||       (DOLFIN default 1e-16) */
  int maximum_iterations;        /* drivers: 10 */
  int line_search;               /* 0 = 'basic' (full step), 1 = 'cp' (critical point, one secant iteration), 2 = 'bt' (backtracking) */
  double krylov_rtol;            /* relative residual of each linear solve (PETSc default 1e-5) */
  double krylov_atol;
  int krylov_maximum_iterations; /* drivers: 1000 / 5000 / N^3/10 */
  int krylov_restart;            /* <= cfg.krylov_basis */
  int preconditioner;            /* 0 = Schur-complement + Jacobi-Chebyshev (default), 1 = point Jacobi, 2 = none */
  int pc_variant;                /* Schur: 0 = full block-LDU, 1 = block upper-triangular */
  int pc_sweeps_mass;            /* Chebyshev sweeps per mass solve */
  int pc_sweeps_shifted;         /* Chebyshev sweeps per (M + alpha K) solve; 0 = automatic (5; 4 in 3-D) */
  int verbose;                   /* 1: print Newton/Krylov history to stderr */
  int linear_solver;             /* 0 = 'gmres' (drivers, bench1.py:98), 1 = 'bicgstab' (bench2_eta1.py:113); ABI 2 */
} pfh_solver_opts;

typedef struct {
  int niters;                    /* SNESGetIterationNumber analogue */
  int converged;                 /* reason > 0 */
  int reason;                    /* 2 atol, 3 rtol, 4 stol; -5 max_it, -3 linear solve failed, -4 NaN */
  double fnorm0, fnorm;          /* ||F|| at entry / exit */
  int krylov_its;                /* total Krylov iterations of this solve */
  int residual_evals, jac_applies, pc_applies;
  double last_krylov_relres;
} pfh_solve_report;

void pfh_default_solver_opts(pfh_solver_opts* opts);
int pfh_abi_version(void);
const char* pfh_last_error(void);

/* NCCL bootstrap for multi-GPU slabs: rank 0 calls this, ships the bytes to the other ranks
 * (torch.distributed broadcast), every rank passes them in pfh_config.nccl_unique_id. */
int pfh_nccl_unique_id(void* out /* PFH_NCCL_ID_BYTES */);

int pfh_create(const pfh_config* cfg, pfh_ctx** out);
int pfh_destroy(pfh_ctx* ctx);
int pfh_n_fields(const pfh_ctx* ctx);
long long pfh_n_owned(const pfh_ctx* ctx);   /* vertices owned by this rank (per field) */
long long pfh_n_global(const pfh_ctx* ctx);  /* vertices of the whole mesh (per field) */

/* dt is a mutable constant in the drivers: dt.assign(x) (dolfin/bench1.py:171,180-183) */
int pfh_set_dt(pfh_ctx* ctx, double dt);
double pfh_get_dt(const pfh_ctx* ctx);

/* w0.interpolate(InitialConditionsBench*(...)) (dolfin/pfbase.py:177-339): evaluate the
 * problem's initial condition at the owned vertices.  ic_params: bench1/6 {c0, epsilon},
 * bench2 {c0, epsilon, epsilon_eta, psi}, bench3 {Delta, r, w, vin, vout}. */
int pfh_initial_condition_dev(pfh_ctx* ctx, const double* ic_params, int n, double* w_dev);

/* Operator-level contract (the dolfinx statement of the same path,
 * dolfinx/pfbase/pde_problems.py:138-177: F(snes,x,b), J(snes,x,J,P)), matrix-free: */
int pfh_residual_dev(pfh_ctx* ctx, const double* w, const double* w0, double* F);  /* F <- R(w; w0, dt) */
int pfh_jac_setup_dev(pfh_ctx* ctx, const double* w);                              /* linearisation point */
int pfh_jac_apply_dev(pfh_ctx* ctx, const double* x, double* y);                   /* y <- J(w) x */
int pfh_precond_setup(pfh_ctx* ctx, const pfh_solver_opts* opts);                  /* after jac_setup / set_dt */
int pfh_precond_apply_dev(pfh_ctx* ctx, const double* r, double* z);               /* z <- P^-1 r */

/* niters, converged = solver.solve() (dolfin/bench1.py:162): w in/out (initial guess = entry value) */
int pfh_newton_solve_dev(pfh_ctx* ctx, double* w, const double* w0,
                         const pfh_solver_opts* opts, pfh_solve_report* report);

/* df.assemble reductions after the step (dolfin/bench1.py:121-125, bench2.py:184-194,
 * bench3.py:162-170, bench6.py:155-165): out[0] = total free energy, out[1] = total solute
 * (bench3: solid fraction), out[2], out[3] reserved (0).  Global (all ranks) values. */
int pfh_energy_dev(pfh_ctx* ctx, const double* w, double out[4]);

/* One piece of a df.assemble functional: out = u_a^T (am M + ak K) u_b over the global mesh, u_a and u_b fields
 * of w (M = consistent mass, K = stiffness).  bench2.py:190-192 assembles kappa_eta/2 |grad eta_i|^2 per order
 * parameter: that is ak = kappa_eta/2, a = b = 2+i.  am = 1, ak = 0 gives int u_a u_b dx. */
int pfh_bilinear_dev(pfh_ctx* ctx, const double* w, int field_a, int field_b, double am, double ak, double* out);

/* f(x) of the reference's peval / sample (dolfin/pfbase.py:50-94): the P1 interpolant of field `field` of w at
 * npts points (pts: npts x dim, row-major; pts and out are DEVICE pointers).  Points outside the mesh give +inf
 * (peval's fill value for points a rank cannot evaluate).  Slab-decomposed meshes: every rank passes the same points
 * and receives all values. */
int pfh_eval_points_dev(pfh_ctx* ctx, const double* w, int field, const double* pts, long long npts, double* out);

/* Host-buffer entry points (reference-facing plugin call; copies inside): one full
 * backward-Euler step w0 -> w at the given dt, plus the step's diagnostics. */
int pfh_step_host(pfh_ctx* ctx, double* w_host /* in: guess, out: solution */, const double* w0_host,
                  double dt, const pfh_solver_opts* opts, pfh_solve_report* report, double diag[4]);
int pfh_residual_host(pfh_ctx* ctx, const double* w_host, const double* w0_host, double* F_host);
int pfh_jac_apply_host(pfh_ctx* ctx, const double* w_host, const double* x_host, double* y_host);
int pfh_energy_host(pfh_ctx* ctx, const double* w_host, double out[4]);

/* Instrumentation: kernels launched by this ctx since creation / last reset; and device
 * time (ms, CUDA events on the ctx stream) accumulated per kernel family when profiling is on. */
long long pfh_launch_count(const pfh_ctx* ctx);
/* Options (key, value): "profiling" 0/1 per-family event timers; "reset_launch_count"; "use_tma" 0/1 staged 3-D
 * kernels (default 1; 0 = gather kernels); "use_tma_loader" 0/1 fill the stages by TMA when rows are 16-byte aligned
 * (default 1; 0 = always the cp.async loader that also serves odd vertex counts); "use_wide" 0/1 wide-ghost
 * preconditioner on slab-decomposed meshes (default 1; 0 = one ghost-plane exchange per sweep); "use_first2" 0/1 first
 * two Chebyshev iterates in one pass (default 1); "scalar_variant", "scalar_zchunk", "jac_variant", "jac_zchunk" tile /
 * ring / chunk variants of the staged kernels (tuning); "cheb_bounds" 1 = Chebyshev interval of the 3-D shifted solves
 * from the operator's symbol (default), 0 = generic bounds, with factors "cheb_flo", "cheb_fhi"; "basis_precision" storage
 * of the GMRES basis: 0 = fp32 when krylov_rtol >= 1e-6 else fp64 (default), 1 = fp64, 2 = fp32 (arithmetic is fp64
 * in all cases); "jac_quadrature_fp32" 0/1 (default 0): the quadrature block of the Krylov J.v of the 3-D fast path in
 * single precision (Newton's residual stays fp64); "op2_tiled" 0/1 (default 1): square-centric residual / J.v kernels on the
 * 2-D crossed mesh (0 = vertex-centric gather kernels); "mg_cycles" (default 1): geometric-multigrid V-cycles per application
 * of the preconditioner's Poisson block (bench6), 0 = a 240-sweep Chebyshev solve instead; "nl_tiled" 0/1 (default 1): the same choice for PFH_PROBLEM_NLDIFF.
 * Unknown keys are an error. */
int pfh_set_option(pfh_ctx* ctx, const char* key, double value);
#define PFH_N_TIMERS 16
int pfh_timers(pfh_ctx* ctx, double ms[PFH_N_TIMERS], long long launches[PFH_N_TIMERS],
               double algorithmic_bytes[PFH_N_TIMERS], int reset);
const char* pfh_timer_name(int i);
void* pfh_stream(pfh_ctx* ctx);
int pfh_synchronize(pfh_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* PFH_B200_H */
