// Internal declarations shared by the translation units of libpfh_b200.so.
// Host-side C++; device-side descriptors are PODs passed to kernels by value.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/pfh_b200.h"

#define PFH_MAX_FIELDS 6

struct PfhError : std::runtime_error {
  using std::runtime_error::runtime_error;
};

void pfh_set_error(const std::string& msg);

#define PFH_CUDA(call)                                                                         \
  do {                                                                                         \
    cudaError_t e__ = (call);                                                                  \
    if (e__ != cudaSuccess) {                                                                  \
      char b__[512];                                                                           \
      snprintf(b__, sizeof b__, "CUDA error %s at %s:%d: %s", cudaGetErrorName(e__), __FILE__, \
               __LINE__, cudaGetErrorString(e__));                                             \
      throw PfhError(b__);                                                                     \
    }                                                                                          \
  } while (0)

#define PFH_CHECK(cond, msg)                  \
  do {                                        \
    if (!(cond)) throw PfhError(std::string(msg)); \
  } while (0)

// ------------------------------------------------------------------------------------------
// Mesh descriptors (device PODs).  Stencil coefficient tables are built on the host by
// assembling the P1 element matrices of the reference's meshes (pfh_tables.cpp).
// ------------------------------------------------------------------------------------------

// 3-D Kuhn mesh = df.BoxMesh (dolfin/b13d.py:24-26).  Offsets of the 15-point P1 stencil:
//  0 self | 1 +x 2 -x 3 +y 4 -y 5 +z 6 -z | 7 +x+y 8 -x-y 9 +y+z 10 -y-z 11 +x+z 12 -x-z 13 +xyz 14 -xyz
// The stiffness matrix only populates offsets 0..6 (checked when the tables are built).
#define PFH_NOFF3 15
#define PFH_NKOFF3 7
struct Mesh3 {
  int nxv, nyv, nzv;   // global vertex counts per direction
  int z0, nzl;         // first owned global plane, number of owned planes
  long long nxy;       // nxv*nyv
  long long n_owned;   // nxy*nzl
  double hx, hy, hz, detJ;  // detJ = hx*hy*hz = 6|T|
  double mI[PFH_NOFF3];     // interior mass row
  double kI[PFH_NKOFF3];    // interior stiffness row
  const double* mT;         // [27][15] rows by boundary type (tz*9+ty*3+tx; 0 lo, 1 interior, 2 hi)
  const double* kT;         // [27][7]
};

// 2-D crossed mesh = df.RectangleMesh(..., 'crossed') (dolfin/bench1.py:21-23).
// Corner row offsets: 0 self | 1 +x 2 -x 3 +y 4 -y | centres: 5 NE(ix,iy) 6 NW(ix-1,iy) 7 SE(ix,iy-1) 8 SW(ix-1,iy-1)
// Centre row offsets: 0 self | corners 1 v0(ix,iy) 2 v1(ix+1,iy) 3 v2(ix,iy+1) 4 v3(ix+1,iy+1)
#define PFH_NOFF2 9
struct Mesh2 {
  int nx, ny;          // cells
  int nxv, nyv;        // corner counts
  long long ncorner, ncentre, n_owned;
  double hx, hy, detJ; // detJ = hx*hy/2 = 2|T|
  double mC[5], kC[5]; // centre rows
  double mI[PFH_NOFF2], kI[PFH_NOFF2];  // interior corner rows
  const double* mT;    // [9][9] corner rows by type (ty*3+tx)
  const double* kT;    // [9][9]
};

// 2-D right-diagonal mesh = df.UnitSquareMesh(Nx, Ny) (dolfin/nl.py:20): grid vertices only, x fastest; every square is
// cut by the diagonal v0-v3.  Used by the steady quasi-linear diffusion problem PFH_PROBLEM_NLDIFF (pfh_nl.cuh), whose
// coefficients depend on the solution: no stencil tables.
struct MeshR2 {
  int nx, ny, nxv, nyv;
  long long n;       // vertices = nxv * nyv
  double hx, hy;
};
struct NlParams {
  double src, kc, reac, bc;  // F = int src v - int kc |u| grad u . grad v - int reac u v, u = bc on x = 0 and x = Lx
};

void pfh_build_tables3(Mesh3& m, std::vector<double>& mT, std::vector<double>& kT);
void pfh_build_tables2(Mesh2& m, std::vector<double>& mT, std::vector<double>& kT);
// Coarse levels 1, 2, .. of the crossed mesh `fine` for the geometric multigrid (pfh_mg2.cuh): cells per direction halved
// while both counts are even and the coarse mesh keeps at least 2 x 2 cells.  meshes[l].mT / kT point into tables[l] (host;
// layout mT then kT, 81 doubles each).
void pfh_mg_build_levels(const Mesh2& fine, std::vector<Mesh2>& meshes, std::vector<std::vector<double>>& tables);

// one coarse level: its mesh (device table pointers in the product) and work vectors of n_owned doubles
struct MgLevel {
  Mesh2 m;
  double *x = nullptr, *b = nullptr, *t = nullptr, *tmp = nullptr;
};

// Block structure the Schur/Chebyshev preconditioner works on (pfh_solver.cpp): fields (0, 1) as a Cahn-Hilliard
// (c, mu) pair, up to four further fields with an SPD parabolic diagonal block am M + ak K, or plain Jacobi.
struct PfhPcLayout {
  bool has_ch = false;
  double Mob = 0, kappa = 0;
  int n_spd = 0;
  int spd_field[4] = {0, 0, 0, 0};
  double spd_am[4] = {0, 0, 0, 0}, spd_ak[4] = {0, 0, 0, 0}, spd_sign[4] = {1, 1, 1, 1};
  bool jacobi_only = false;  // some field fits neither block: fall back to the point-Jacobi preconditioner
};

// Scalar field operand of a stencil kernel: owned planes + (3-D slabs) one ghost plane per side.
struct FieldRef {
  const double* p;
  const double* glo;
  const double* ghi;
};

// Linear part of the residual / Jacobian:
//   R_a = sum_b (Mc[a][b] M + Kc[a][b] K) u_b - Mp[a] M u0_a + N[g_a](u)
struct LinCoefs {
  double Mc[PFH_MAX_FIELDS][PFH_MAX_FIELDS];
  double Kc[PFH_MAX_FIELDS][PFH_MAX_FIELDS];
  double Mp[PFH_MAX_FIELDS];
};

struct ProblemParams {
  int id;
  int nf;
  double p[PFH_MAX_PARAMS];
};

// ------------------------------------------------------------------------------------------
// Context
// ------------------------------------------------------------------------------------------
enum TimerId {
  T_RESIDUAL = 0, T_JAC_APPLY, T_CHEB_SWEEP, T_CHEB_FIRST, T_LIN_APPLY, T_MULTIDOT, T_MULTIAXPY,
  T_VEC, T_REDUCE, T_ENERGY, T_HALO, T_ALLREDUCE, T_IC, T_JACOBI, T_COPY, T_OTHER
};

struct EventPair {
  cudaEvent_t a, b;
  int id;
};

struct Comm;  // pfh_comm.cpp

struct pfh_ctx {
  pfh_config cfg;
  int dim = 0;
  int nf = 0;
  ProblemParams prob;
  Mesh2 m2;
  Mesh3 m3;
  MeshR2 mr;     // cfg.mesh_kind == PFH_MESH_RIGHT2D
  NlParams nlp;  // cfg.problem == PFH_PROBLEM_NLDIFF
  long long n_owned = 0, n_global = 0;
  long long nplane = 0;  // ghost plane size (3-D) or 0
  double dt = 0.1;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  // halo/compute overlap (nranks > 1): ghost planes travel on comm_stream while interior planes are computed
  cudaStream_t comm_stream = nullptr;
  cudaEvent_t ev_bnd = nullptr, ev_halo = nullptr;
  bool overlap = true;
  const double* gS_prefetched = nullptr;  // scalar field whose ghost planes are already travelling into gS (ev_halo)
  const double* gR_prefetched = nullptr;  // right-hand side whose wide ghost planes are already travelling into gR (ev_rhs)
  int gR_prefetched_w = 0;
  cudaEvent_t ev_rhs = nullptr;
  const double* gX_prefetched = nullptr;  // state vector whose ghost planes are already travelling into gX (ev_halo): the
                                          // wide-ghost preconditioner ships its result for the J.v that follows
  double* d_tables = nullptr;

  // workspaces (device)
  std::vector<double*> allocs;
  double *U = nullptr, *U0 = nullptr;  // device staging of caller host buffers (host entry points)
  double *R = nullptr, *B = nullptr, *DX = nullptr, *DX2 = nullptr;  // nf*n_owned each
  double* V = nullptr;                                               // (basis+1) * nf*n_owned
  double *W1 = nullptr, *W2 = nullptr;                               // nf*n_owned Krylov temporaries
  double* S[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};  // scalar temporaries (n_owned)
  // ghost planes: [nf][nplane] each
  double *gU_lo = nullptr, *gU_hi = nullptr;    // linearisation point / current iterate
  double *gU0_lo = nullptr, *gU0_hi = nullptr;  // previous step
  double *gX_lo = nullptr, *gX_hi = nullptr;    // operand of an operator apply
  double *gS_lo = nullptr, *gS_hi = nullptr;    // scalar operand (preconditioner sweeps)
  // Wide-ghost preconditioner (nranks > 1): the scalar work vectors S[] are allocated with `gw`
  // extra planes below and above the owned slab (S[i] points at owned plane 0), so a Chebyshev solve exchanges its
  // right-hand side ONCE at the width it needs and then sweeps redundantly on a shrinking halo; gR_* hold the wide
  // ghost planes of a right-hand side that lives in a Krylov vector.
  int gw = 0;
  int min_slab = 0;  // smallest slab height over the ranks (an exchange of width w needs w <= min_slab)
  double *gR_lo = nullptr, *gR_hi = nullptr;
  const double* jac_point = nullptr;            // device pointer to linearisation point (nf*n_owned)

  // reductions
  double* d_partial = nullptr;  // [PFH_RED_MAXBLOCKS][PFH_RED_MAXVALS]
  double* d_red = nullptr;      // final values (device)
  double* h_red = nullptr;      // mapped pinned host memory the final reduction kernel writes into (zero copy)
  unsigned long long* h_seq = nullptr;  // sequence number next to it, written after the values; the host polls it
  unsigned long long red_seq = 0;

  // preconditioner state
  pfh_solver_opts pc_opts;
  bool pc_ready = false;
  double pc_alpha = 0.0;
  double lmax = 2.5, lmin_m = 0.5, lmin_a = 0.1;
  double lmax_m = 2.5, lmax_a = 2.5;       // Chebyshev intervals of the mass and the shifted (M + alpha K) solves
  double lmin_spd[4] = {0.1, 0.1, 0.1, 0.1};
  int cheb_bounds = 1;                     // 3-D shifted solve: 1 = interval from the operator's symbol, 0 = generic bounds
  double cheb_flo = 0.0, cheb_fhi = 1.015;  // factors on the symbol's extreme values (flo 0: keep the Rayleigh estimate)
  std::vector<double> h_mT, h_kT;          // host copies of the stencil type tables

  Comm* comm = nullptr;

  // instrumentation
  long long launches = 0;
  bool profiling = false;
  std::vector<EventPair> ev_pending;
  std::vector<EventPair> ev_free;
  double t_ms[PFH_N_TIMERS] = {0};
  double t_bytes[PFH_N_TIMERS] = {0};  // algorithmic bytes of the launches timed above
  long long t_n[PFH_N_TIMERS] = {0};
  const double* scalar_sig = nullptr;  // 2-D scalar operators: optional per-vertex diagonal shift (lumped reaction term)
  double* SIG[4] = {nullptr, nullptr, nullptr, nullptr};
  int scalar_bc = 0;    // 2-D scalar operators: 1 = Dirichlet (identity) rows on the corner vertices of x = 0 and x = Lx
  int scalar_variant = 0;  // tile/ring variant of the staged scalar kernel (tuning; pfh_set_option "scalar_variant")
  int scalar_zchunk = 0;   // planes per CTA of the staged scalar kernel (0 = default 32; tuning)
  int jac_variant = 0;     // ring depth / occupancy variant of the cube-centric kernel (tuning)
  int jac_zchunk = 0;      // planes per CTA of the cube-centric J.v / residual kernel (0 = default 32; tuning)
  bool use_wide = true;    // nranks > 1: wide-ghost Schur preconditioner (false: one ghost-plane exchange per sweep)
  void* tmap_cache = nullptr;  // tensor-map cache (pfh_scalar3_tma.cu)
  int n_sm = 0;                // multiprocessors of the device (grid sizing)
  // the boundary-shell kernel of a scalar-operator launch runs on aux_stream, concurrently with the staged kernel
  cudaStream_t aux_stream = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  bool shell_async = true;
  // Krylov J.v of the 3-D TMA path: quadrature block in single precision (elem::cube_jac_quartic_f32).  Opt-in: validated
  // numerically on the CPU solver leg, not yet measured on hardware.
  bool jac_quadrature_fp32 = false;
  bool use_first2 = true;  // 3-D: first two Chebyshev iterates in one pass over r (k_cheb_first2)
  int basis_precision = 0;  // Krylov basis storage: 0 = fp32 when krylov_rtol >= 1e-6 else fp64, 1 = fp64, 2 = fp32
  // bench6 Poisson block: geometric multigrid V-cycles instead of a long Chebyshev solve (pfh_solver.cpp)
  std::vector<MgLevel> mg;  // coarse levels 1.. of the crossed mesh (empty: no hierarchy)
  int mg_cycles = 1;        // V-cycles per preconditioner application; 0 = the 240-sweep Chebyshev solve of round 1
  bool op2_tiled = true;  // 2-D crossed mesh: square-centric residual / J.v (pfh_op2_tiled.cuh); false: the vertex-centric gather k2_op
  bool nl_tiled = true;  // PFH_PROBLEM_NLDIFF: square-centric tiled kernels (false: the vertex-centric gather kernels)
  bool use_tma = true;  // 3-D fast path: staged kernels (pfh_scalar3_tma.cu, pfh_stencil3_tma.cu); false = gather kernels only
  bool use_tma_loader = true;  // stage with TMA when the operands allow it (false: always the cp.async loader; testing)
};

#define PFH_RED_MAXBLOCKS 1184  // 8 CTAs per SM * 148 SMs
#define PFH_RED_MAXVALS 36

double* pfh_alloc(pfh_ctx* c, size_t n_doubles);
void pfh_prof_begin(pfh_ctx* c, int id, cudaStream_t s = nullptr);
void pfh_prof_end(pfh_ctx* c, int id, cudaStream_t s = nullptr);
void pfh_prof_resolve(pfh_ctx* c);

struct ProfScope {
  pfh_ctx* c;
  int id;
  cudaStream_t s;
  ProfScope(pfh_ctx* c_, int id_, double algorithmic_bytes = 0.0, cudaStream_t s_ = nullptr) : c(c_), id(id_), s(s_) {
    if (id != T_COPY && id != T_HALO && id != T_ALLREDUCE) c->launches++;  // our own kernels only
    if (c->profiling) {
      c->t_bytes[id] += algorithmic_bytes;
      pfh_prof_begin(c, id, s);
    }
  }
  ~ProfScope() {
    if (c->profiling) pfh_prof_end(c, id, s);
  }
};

// ---- stencil launchers (pfh_stencil.cu); operands are device pointers, ghosts must be current ----
LinCoefs pfh_lin_coefs(const pfh_ctx* c, double dt);
void k_residual(pfh_ctx* c, const double* u, const double* u0, double* R);
void k_jac_apply(pfh_ctx* c, const double* ulin, const double* x, double* y);
// out = s * add + (am M + ak K) in   (add may be null when s == 0)
void k_lin_apply(pfh_ctx* c, double am, double ak, FieldRef in, double s, const double* add, double* out);
// z = b * Dinv r,  D = diag(am M + ak K)
void k_cheb_first(pfh_ctx* c, double am, double ak, double b, const double* r, double* z);
// first two Chebyshev iterates from r alone (3-D TMA path): z1 = b0 D^-1 r, z2 = z1 + a1 z1 + b1 D^-1 (r - A z1); false = not available
// (z1 may be null: the third iterate can recompute it from r, see k_cheb_sweep)
bool k_cheb_first2(pfh_ctx* c, double am, double ak, double b0, double a1, double b1, FieldRef r, double* z1, double* z2);
// zprev <- z + a (z - zprev) + b Dinv (r - (am M + ak K) z)
// zprev_is_zero: zprev is write-only, z_{k-1} = 0; or, with zprev_b0 != 0 (staged 3-D kernel only), z_{k-1} = zprev_b0 Dinv r
void k_cheb_sweep(pfh_ctx* c, double am, double ak, double a, double b, const double* r, FieldRef z,
                  double* zprev, bool zprev_is_zero, double zprev_b0 = 0.0);
// point-Jacobi on the Jacobian diagonal blocks: z_a = r_a / diag_a
void k_jacobi(pfh_ctx* c, const double* ulin, const double* r, double* z);
void k_eval_points(pfh_ctx* c, FieldRef f, const double* pts, long long npts, double* out);  // pfh_sample.cu
void k_energy(pfh_ctx* c, const double* u, double out[4]);  // includes final reduction + allreduce + D2H
void k_initial_condition(pfh_ctx* c, const double* icp, int n, double* w);
void k_dirichlet_fix(pfh_ctx* c, double* w);  // impose Dirichlet values on the iterate (bench6)
void k_dirichlet_copy(pfh_ctx* c, const double* src, double* dst);
// sig_i = max(d g_field / d u_field (u_i), floor_coef) * rowsum(M)_i  (2-D)
void k_reaction_diag(pfh_ctx* c, const double* u, int field, double floor_coef, double* sig);  // dst = src on the Dirichlet vertices of a scalar field

// ---- steady quasi-linear diffusion on the right-diagonal mesh (pfh_nl.cu; dolfin/nl.py) ----
inline bool pfh_is_nl(const pfh_ctx* c) { return c->prob.id == PFH_PROBLEM_NLDIFF; }
void k_nl_residual(pfh_ctx* c, const double* u, double* R);
void k_nl_jac(pfh_ctx* c, const double* ulin, const double* x, double* y);
// one Jacobi-Chebyshev sweep on A = K_{kc|ulin|} + reac M (the SPD part of -J) with right-hand side sgn*r:
// zprev <- z + a (z - zprev) + b D^-1 (sgn r - A z) on the free vertices, zprev <- r on the Dirichlet vertices
void k_nl_sweep(pfh_ctx* c, const double* ulin, double a, double b, double sgn, const double* r, const double* z, double* zprev);

// ---- multigrid transfers on the crossed mesh (pfh_stencil.cu; arithmetic in pfh_mg2.cuh) ----
// bc_ = scale * P^T tf on the coarse mesh mc, zero on its Dirichlet vertices when bc != 0
void k2_mg_restrict(pfh_ctx* c, const Mesh2& mf, const Mesh2& mc, const double* tf, double scale, double* bc_, int bc);
// xf += P xc (Dirichlet vertices of the fine mesh untouched when bc != 0)
void k2_mg_prolong_add(pfh_ctx* c, const Mesh2& mf, const Mesh2& mc, const double* xc, double* xf, int bc);

// ---- staged 3-D kernels (pfh_scalar3_tma.cu, pfh_stencil3_tma.cu) ----
bool pfh_fast3_ok(const pfh_ctx* c);   // 3-D and the staged kernels are enabled
void pfh_scalar3_init_device();        // per-device kernel attributes (dynamic shared memory opt-in)
void pfh_jac3_init_device();
void pfh_tmap_cache_free(pfh_ctx* c);
// One launch of the scalar operator A = am M + ak K on a window of planes (pfh_scalar3_tma.cu).  Output planes are
// launch-local zl = 0 .. nzo-1 = global planes gz0 .. gz0+nzo-1; out / out1 / r / o point at launch-local plane 0.
// The stencil operand is read on launch-local planes -1 .. nzo: plane zi has index zi + in_off in the main array `in`
// (in_n planes); indices below 0 come from the last planes of `in_lo` (n_lo planes), indices >= in_n from `in_hi`;
// whatever lies outside every array reads as zero (physical boundary).
struct ScalarOp {
  int mode = 0;                    // 0: out = a r + A in [out1 = in]; 2: out = in + a (in - o) + b D^-1 (r - A in);
                                   // 3: out1 = b0 D^-1 in, out = out1 + a out1 + b D^-1 (in - A out1)
  double am = 0, ak = 0, a = 0, b = 0, b0 = 0;
  double ob0 = 0;                  // mode 2 with o == nullptr: o = ob0 D^-1 r, recomputed from r (0: o = 0)
  const double* in = nullptr;
  int in_off = 0, in_n = 0;
  const double* in_lo = nullptr;
  int n_lo = 0;
  const double* in_hi = nullptr;
  int n_hi = 0;
  const double* r = nullptr;       // pointwise operand, addressed like `in`: plane zl has index zl + r_off in `r` (r_n planes,
  int r_off = 0, r_n = 0;          // 0 = nzo), below / above that in r_lo / r_hi
  const double* r_lo = nullptr;
  int n_rlo = 0;
  const double* r_hi = nullptr;
  int n_rhi = 0;
  const double* o = nullptr;
  double* out = nullptr;
  double* out1 = nullptr;
  int gz0 = 0, nzo = 0;
};
void k3_scalar_launch(pfh_ctx* c, const ScalarOp& op);
bool k3_jac_tma_launch(pfh_ctx* c, const LinCoefs& L, const double* ulin, const double* x, double* y);
bool k3_res_tma_launch(pfh_ctx* c, const LinCoefs& L, const double* u, const double* u0, double* R);
bool k3_ch_crow_tma_launch(pfh_ctx* c, const LinCoefs& L, const double* ulin, FieldRef xc, double* out);
// out = [J(ulin) (xc, 0)]_mu for Cahn-Hilliard (scalar in, scalar out); false if no fast path applies
bool k_ch_crow(pfh_ctx* c, const double* ulin, FieldRef xc, double* out);

// ---- vector kernels (pfh_vec.cu) ----
void v_copy(pfh_ctx* c, double* dst, const double* src, long long n);
void v_zero(pfh_ctx* c, double* dst, long long n);
void v_axpy(pfh_ctx* c, double a, const double* x, double* y, long long n);               // y += a x
void v_axpby(pfh_ctx* c, double a, const double* x, double b, const double* y, double* out, long long n);
void v_scale(pfh_ctx* c, double a, const double* x, double* out, long long n);            // out = a x
// dots[k] = V_k . w (k < nv), dots[nv] = w . w ; global (allreduced) values returned on the host
void v_multidot(pfh_ctx* c, const double* V, long long stride, int nv, const double* w, long long n,
                double* host_out);
// out = scale * (w - sum_k h[k] V_k)
void v_multiaxpy(pfh_ctx* c, const double* V, long long stride, int nv, const double* h,
                 const double* w, double scale, double* out, long long n);
// out = sum_k y[k] V_k
void v_combine(pfh_ctx* c, const double* V, long long stride, int nv, const double* y, double* out,
               long long n);
// the same three with the basis stored in fp32 (compressed-basis GMRES); multiaxpy32 also stores the fp32-rounded
// result in out32 (may be null), writes that rounded value, widened, to out, and returns its global squared norm
// in *norm2 (may be null) from the same pass
void v_multidot32(pfh_ctx* c, const float* V, long long stride, int nv, const double* w, long long n, double* host_out);
void v_multiaxpy32(pfh_ctx* c, const float* V, long long stride, int nv, const double* h, const double* w, double scale,
                   double* out, float* out32, long long n, double* norm2);
// called by v_multiaxpy32 between the launch of its last kernel and the (blocking) read of the norm: lets the solver
// start shipping the ghost planes of the new vector while the host waits for the reduction
void s_after_vector_written(pfh_ctx* c, const double* vec);
void v_combine32(pfh_ctx* c, const float* V, long long stride, int nv, const double* y, double* out, long long n);
double v_norm2(pfh_ctx* c, const double* x, long long n);  // global ||x||_2^2 on the host
// finish a two-stage reduction: sums d_partial[nblocks][nvals] -> allreduce -> host; returns h_red
const double* pfh_reduce_finish(pfh_ctx* c, int nblocks, int nvals);

// ---- communication (pfh_comm.cpp) ----
Comm* comm_create(const pfh_config& cfg, cudaStream_t s);
void comm_destroy(Comm*);
void comm_allreduce_sum(pfh_ctx* c, double* d_buf, int n);
// exchange boundary planes of nfields fields stored field-major with stride fstride
void comm_halo(pfh_ctx* c, const double* base, long long fstride, int nfields, double* glo, double* ghi,
               cudaStream_t stream = nullptr);  // stream == nullptr: the ctx stream
// exchange the first / last `nplanes` planes of one scalar field (ghost width nplanes): glo receives the lower
// neighbour's last planes, ghi the upper neighbour's first planes
void comm_halo_n(pfh_ctx* c, const double* base, int nplanes, double* glo, double* ghi, cudaStream_t stream = nullptr);
int comm_unique_id(void* out);
bool comm_halo_is_separate(const pfh_ctx* c);  // the halo exchanges have their own communicator (may overlap a reduction)

// ---- solver (pfh_solver.cpp) ----
void s_exchange_state(pfh_ctx* c, const double* u, double* glo, double* ghi);
void s_residual(pfh_ctx* c, const double* u, const double* u0, double* R, bool u0_ghost_current);
FieldRef s_field_ref(pfh_ctx* c, const double* u);  // one field + its freshly exchanged ghost planes
void s_bilinear_apply(pfh_ctx* c, const double* u, double am, double ak, double* out);  // out = (am M + ak K) u, one field
void s_jac_setup(pfh_ctx* c, const double* u);
void s_jac_apply(pfh_ctx* c, const double* x, double* y);
void s_precond_setup(pfh_ctx* c, const pfh_solver_opts* o);
void s_precond_apply(pfh_ctx* c, const double* r, double* z);
void s_newton(pfh_ctx* c, double* w, const double* w0, const pfh_solver_opts* o, pfh_solve_report* rep);
