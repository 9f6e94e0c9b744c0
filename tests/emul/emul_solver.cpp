// TEST INFRASTRUCTURE: the product's host-side solver (csrc/pfh_solver.cpp: Newton with the 'basic' and 'cp' line
// searches, GMRES with the fp64 and the fp32-stored basis, the Schur/Chebyshev preconditioner and its block layouts)
// compiled for the CPU, with the kernel launchers it calls replaced by plain loops over the same __host__ __device__
// per-vertex functions the CUDA kernels call (csrc/pfh_vertex.cuh).  This lets the whole Newton-Krylov step of the
// product be compared with the oracle WITHOUT a GPU; the launch layer itself (grids, TMA tiles, streams, NCCL) is what
// the GPU tests cover.  Never loaded by the product.  Single rank only: the communication hooks are no-ops.
#include "emul.cpp"  // Setup, setup(), run<P>(): residual / J.v / energy loops over the per-vertex code

#include <cstring>
#include <map>

// ---- CUDA runtime entry points the solver touches for stream ordering: nothing to order on the host ----
extern "C" {
cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned int) { return cudaSuccess; }
cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
const char* cudaGetErrorString(cudaError_t) { return "cuda stub"; }
const char* cudaGetErrorName(cudaError_t) { return "cudaStub"; }
}
void pfh_prof_begin(pfh_ctx*, int, cudaStream_t) {}
void pfh_prof_end(pfh_ctx*, int, cudaStream_t) {}

namespace {
std::map<const pfh_ctx*, Setup*> g_setup;
const Setup& setup_of(const pfh_ctx* c) { return *g_setup.at(c); }

#ifdef PFH_USER_PROBLEM_HEADER
#define EMUL_USER_CASE(BODY) case PFH_PROBLEM_USER: { using P = ProbUser; BODY; } break;   // a generated model (codegen.py)
#else
#define EMUL_USER_CASE(BODY)
#endif
#define EMUL_DISPATCH(c, BODY)                                           \
  switch ((c)->prob.id) {                                                \
    case PFH_PROBLEM_CH: { using P = ProbCH; BODY; } break;              \
    case PFH_PROBLEM_BENCH2: { using P = ProbBench2; BODY; } break;      \
    case PFH_PROBLEM_BENCH3: { using P = ProbBench3; BODY; } break;      \
    case PFH_PROBLEM_BENCH6: { using P = ProbBench6; BODY; } break;      \
    EMUL_USER_CASE(BODY)                                                 \
    default: throw PfhError("emul: unknown problem id");                 \
  }

template <class P>
LinCoefs lin_of(const pfh_ctx* c, double dt) {
  LinCoefs L;
  memset(&L, 0, sizeof L);
  P(c->prob).lin(dt, L);
  return L;
}

// the three scalar-operator modes of k2_scalar / k3_scalar (csrc/pfh_stencil.cu), one vertex
// MODE 0: out = s*add + A in ; MODE 1: z = b Dinv r ; MODE 2: out <- in + a (in - out) + b Dinv (r - A in)
template <int MODE>
void scalar_all(pfh_ctx* c, double am, double ak, double a, double b, FieldRef in, const double* r, double* out, bool out_is_zero) {
  if (c->dim == 3) {
    const Mesh3& m = c->m3;
    for (int zl = 0; zl < m.nzl; ++zl)
      for (int iy = 0; iy < m.nyv; ++iy)
        for (int ix = 0; ix < m.nxv; ++ix) {
          const long long i = (long long)zl * m.nxy + (long long)iy * m.nxv + ix;
          if (MODE == 1) {
            out[i] = b / vtx::scalar_diag3(m, am, ak, ix, iy, zl) * r[i];
          } else {
            double Av, d, self;
            vtx::scalar_op3(m, am, ak, in, ix, iy, zl, Av, d, self);
            if (MODE == 0) out[i] = (a != 0.0 ? a * r[i] : 0.0) + Av;
            else out[i] = self + a * (self - (out_is_zero ? 0.0 : out[i])) + (b / d) * (r[i] - Av);
          }
        }
    return;
  }
  const Mesh2& m = c->m2;
  const double* sig = c->scalar_sig;
  for (long long i = 0; i < m.n_owned; ++i) {
    if (c->scalar_bc && i < m.ncorner) {
      const int ix = (int)(i % m.nxv);
      if (ix == 0 || ix == m.nxv - 1) {
        out[i] = r[i];
        continue;
      }
    }
    const double sg = sig ? sig[i] : 0.0;
    if (MODE == 1) {
      out[i] = b / (vtx::scalar_diag2(m, am, ak, i) + sg) * r[i];
    } else {
      double Av, d, self;
      vtx::scalar_op2(m, am, ak, in.p, i, Av, d, self);
      Av += sg * self;
      d += sg;
      if (MODE == 0) out[i] = (a != 0.0 ? a * r[i] : 0.0) + Av;
      else out[i] = self + a * (self - (out_is_zero ? 0.0 : out[i])) + (b / d) * (r[i] - Av);
    }
  }
}

template <class P>
void reaction_diag(pfh_ctx* c, const double* U, int field, double floor_coef, double* sig) {
  const P prob(c->prob);
  const Mesh2& m = c->m2;
  for (long long i = 0; i < m.n_owned; ++i) {
    double u[P::NIN], x[P::NIN], o[P::NOUT];
    int oi = -1;
    for (int f = 0; f < P::NIN; ++f) {
      u[f] = U[P::in(f) * m.n_owned + i];
      x[f] = P::in(f) == field ? 1.0 : 0.0;
    }
    for (int k = 0; k < P::NOUT; ++k)
      if (P::out(k) == field) oi = k;
    double s = 0.0;
    if (oi >= 0) {
      prob.dg(u, x, o);
      s = o[oi];
    }
    double ml;
    if (i < m.ncorner) {
      vtx::Row2 row;
      vtx::get_row2(m, vtx::type2(m, (int)(i % m.nxv), (int)(i / m.nxv)), row);
      ml = 0.0;
      for (int k = 0; k < PFH_NOFF2; ++k) ml += row.m[k];
    } else {
      ml = m.mC[0] + m.mC[1] + m.mC[2] + m.mC[3] + m.mC[4];
    }
    sig[i] = fmax(s, floor_coef) * ml;
  }
}
}  // namespace

// ---------------------------------------------------------------------------------------------------------------------
// launch layer, host version (declarations: csrc/pfh_internal.h)
// ---------------------------------------------------------------------------------------------------------------------
LinCoefs pfh_lin_coefs(const pfh_ctx* c, double dt) {
  LinCoefs L;
  EMUL_DISPATCH(c, L = lin_of<P>(c, dt));
  return L;
}
bool pfh_fast3_ok(const pfh_ctx*) { return false; }   // the staged kernels (and the wide-ghost path built on them) are GPU-only
void k3_scalar_launch(pfh_ctx*, const ScalarOp&) { throw PfhError("emul: staged scalar kernel is not emulated"); }
bool k_cheb_first2(pfh_ctx*, double, double, double, double, double, FieldRef, double*, double*) { return false; }
bool k_ch_crow(pfh_ctx*, const double*, FieldRef, double*) { return false; }

// dolfin/nl.py: loops over nl::vertex_op (csrc/pfh_nl.cuh), the function k_nl_gather wraps on the device
void k_nl_residual(pfh_ctx* c, const double* u, double* R) {
  const MeshR2& m = c->mr;
  for (int iy = 0; iy < m.nyv; ++iy)
    for (int ix = 0; ix < m.nxv; ++ix) {
      double d;
      nl::vertex_op<0>(m, c->nlp, u, nullptr, ix, iy, R[(long long)iy * m.nxv + ix], d);
    }
}
void k_nl_jac(pfh_ctx* c, const double* ulin, const double* x, double* y) {
  const MeshR2& m = c->mr;
  for (int iy = 0; iy < m.nyv; ++iy)
    for (int ix = 0; ix < m.nxv; ++ix) {
      double d;
      nl::vertex_op<1>(m, c->nlp, ulin, x, ix, iy, y[(long long)iy * m.nxv + ix], d);
    }
}
void k_nl_sweep(pfh_ctx* c, const double* ulin, double a, double b, double sgn, const double* r, const double* z, double* zprev) {
  const MeshR2& m = c->mr;
  for (int iy = 0; iy < m.nyv; ++iy)
    for (int ix = 0; ix < m.nxv; ++ix) {
      const long long i = (long long)iy * m.nxv + ix;
      if (nl::is_dirichlet(m, ix)) {
        zprev[i] = r[i];
        continue;
      }
      double Az, d;
      nl::vertex_op<2>(m, c->nlp, ulin, z, ix, iy, Az, d);
      const double zp = a != 0.0 ? zprev[i] : z[i];
      zprev[i] = z[i] + a * (z[i] - zp) + (b / d) * (sgn * r[i] - Az);
    }
}

void k_residual(pfh_ctx* c, const double* u, const double* u0, double* R) {
  if (pfh_is_nl(c)) return k_nl_residual(c, u, R);
  EMUL_DISPATCH(c, (run<P>(setup_of(c), c->prob, c->dt, 0, u, nullptr, u0, nullptr, nullptr, R, nullptr)));
}
void k_jac_apply(pfh_ctx* c, const double* ulin, const double* x, double* y) {
  if (pfh_is_nl(c)) return k_nl_jac(c, ulin, x, y);
  EMUL_DISPATCH(c, (run<P>(setup_of(c), c->prob, c->dt, 1, ulin, x, nullptr, nullptr, nullptr, y, nullptr)));
}
void k_lin_apply(pfh_ctx* c, double am, double ak, FieldRef in, double s, const double* add, double* out) {
  scalar_all<0>(c, am, ak, s, 0.0, in, add, out, false);
}
void k_cheb_first(pfh_ctx* c, double am, double ak, double b, const double* r, double* z) {
  scalar_all<1>(c, am, ak, 0.0, b, FieldRef{nullptr, nullptr, nullptr}, r, z, false);
}
void k_cheb_sweep(pfh_ctx* c, double am, double ak, double a, double b, const double* r, FieldRef z, double* zprev, bool zprev_is_zero, double) {
  scalar_all<2>(c, am, ak, a, b, z, r, zprev, zprev_is_zero);
}
void k_jacobi(pfh_ctx* c, const double* ulin, const double* r, double* z) {
  if (pfh_is_nl(c)) {
    memset(c->S[0], 0, c->n_owned * sizeof(double));
    return k_nl_sweep(c, ulin, 0.0, 1.0, -1.0, r, c->S[0], z);
  }
  const LinCoefs L = pfh_lin_coefs(c, c->dt);
  const long long n = c->n_owned;
  for (long long i = 0; i < n; ++i)
    for (int a = 0; a < c->nf; ++a) {
      double d;
      if (c->dim == 2) {
        d = vtx::scalar_diag2(c->m2, L.Mc[a][a], L.Kc[a][a], i);
      } else {
        const int zl = (int)(i / c->m3.nxy), rem = (int)(i - (long long)zl * c->m3.nxy);
        d = vtx::scalar_diag3(c->m3, L.Mc[a][a], L.Kc[a][a], rem % c->m3.nxv, rem / c->m3.nxv, zl);
      }
      z[a * n + i] = r[a * n + i] / (d != 0.0 ? d : 1.0);
    }
}
void k_dirichlet_fix(pfh_ctx* c, double* w) {
  if (c->dim != 2 || c->prob.id != PFH_PROBLEM_BENCH6) return;
  const ProbBench6 prob(c->prob);
  const Mesh2& m = c->m2;
  for (long long i = 0; i < m.ncorner; ++i) {
    const int ix = (int)(i % m.nxv), iy = (int)(i / m.nxv);
    for (int a = 0; a < ProbBench6::NF; ++a) {
      double val;
      if (prob.dirichlet(a, ix, iy, m.nx, iy * m.hy, val)) w[a * m.n_owned + i] = val;
    }
  }
}
void k_dirichlet_copy(pfh_ctx* c, const double* src, double* dst) {
  if (c->dim != 2) return;
  const Mesh2& m = c->m2;
  for (long long i = 0; i < m.ncorner; ++i) {
    const int ix = (int)(i % m.nxv);
    if (ix == 0 || ix == m.nxv - 1) dst[i] = src[i];
  }
}
void k2_mg_restrict(pfh_ctx*, const Mesh2& mf, const Mesh2& mc, const double* tf, double scale, double* b, int bc) {
  for (long long i = 0; i < mc.n_owned; ++i) b[i] = (bc && mg2::is_dirichlet(mc, i)) ? 0.0 : scale * mg2::restrict_node(mf, mc, tf, i);
}
void k2_mg_prolong_add(pfh_ctx*, const Mesh2& mf, const Mesh2& mc, const double* xc, double* xf, int bc) {
  for (long long i = 0; i < mf.n_owned; ++i)
    if (!(bc && mg2::is_dirichlet(mf, i))) xf[i] += mg2::prolong_node(mf, mc, xc, i);
}
void k_reaction_diag(pfh_ctx* c, const double* u, int field, double floor_coef, double* sig) {
  PFH_CHECK(c->dim == 2, "reaction diagonal is only used by the 2-D multi-field problems");
  EMUL_DISPATCH(c, (reaction_diag<P>(c, u, field, floor_coef, sig)));
}

// ---- vector kernels ----
void v_copy(pfh_ctx*, double* dst, const double* src, long long n) { memmove(dst, src, n * sizeof(double)); }
void v_zero(pfh_ctx*, double* dst, long long n) { memset(dst, 0, n * sizeof(double)); }
void v_axpby(pfh_ctx*, double a, const double* x, double b, const double* y, double* out, long long n) {
  for (long long i = 0; i < n; ++i) out[i] = a * x[i] + (b != 0.0 ? b * y[i] : 0.0);
}
void v_axpy(pfh_ctx* c, double a, const double* x, double* y, long long n) { v_axpby(c, a, x, 1.0, y, y, n); }
void v_scale(pfh_ctx* c, double a, const double* x, double* out, long long n) { v_axpby(c, a, x, 0.0, nullptr, out, n); }
double v_norm2(pfh_ctx*, const double* x, long long n) {
  double s = 0.0;
  for (long long i = 0; i < n; ++i) s += x[i] * x[i];
  return s;
}
namespace {
template <class TV>
void multidot_t(const TV* V, long long stride, int nv, const double* w, long long n, double* out) {
  for (int k = 0; k < nv; ++k) {
    double s = 0.0;
    for (long long i = 0; i < n; ++i) s += (double)V[k * stride + i] * w[i];
    out[k] = s;
  }
  double ww = 0.0;
  for (long long i = 0; i < n; ++i) ww += w[i] * w[i];
  out[nv] = ww;
}
// out = wscale * w - sum_k coef_k V_k ; optional fp32 copy (rounded value also returned in out) and its squared norm
template <class TV>
void multiaxpy_t(const TV* V, long long stride, int nv, const double* coef, const double* w, double wscale, double* out, float* out32,
                 long long n, double* norm2) {
  double acc = 0.0;
  for (long long i = 0; i < n; ++i) {
    double r = w ? wscale * w[i] : 0.0;
    for (int k = 0; k < nv; ++k) r -= coef[k] * (double)V[k * stride + i];
    if (out32) {
      out32[i] = (float)r;
      r = (double)out32[i];
    }
    acc += r * r;
    out[i] = r;
  }
  if (norm2) *norm2 = acc;
}
}  // namespace
void v_multidot(pfh_ctx*, const double* V, long long stride, int nv, const double* w, long long n, double* o) { multidot_t(V, stride, nv, w, n, o); }
void v_multidot32(pfh_ctx*, const float* V, long long stride, int nv, const double* w, long long n, double* o) { multidot_t(V, stride, nv, w, n, o); }
void v_multiaxpy(pfh_ctx*, const double* V, long long stride, int nv, const double* h, const double* w, double scale, double* out, long long n) {
  std::vector<double> cf(nv);
  for (int k = 0; k < nv; ++k) cf[k] = scale * h[k];
  multiaxpy_t(V, stride, nv, cf.data(), w, scale, out, nullptr, n, nullptr);
}
void v_multiaxpy32(pfh_ctx*, const float* V, long long stride, int nv, const double* h, const double* w, double scale, double* out,
                   float* out32, long long n, double* norm2) {
  std::vector<double> cf(nv);
  for (int k = 0; k < nv; ++k) cf[k] = scale * h[k];
  multiaxpy_t(V, stride, nv, cf.data(), w, scale, out, out32, n, norm2);
}
void v_combine(pfh_ctx*, const double* V, long long stride, int nv, const double* y, double* out, long long n) {
  std::vector<double> cf(nv);
  for (int k = 0; k < nv; ++k) cf[k] = -y[k];
  multiaxpy_t(V, stride, nv, cf.data(), nullptr, 0.0, out, nullptr, n, nullptr);
}
void v_combine32(pfh_ctx*, const float* V, long long stride, int nv, const double* y, double* out, long long n) {
  std::vector<double> cf(nv);
  for (int k = 0; k < nv; ++k) cf[k] = -y[k];
  multiaxpy_t(V, stride, nv, cf.data(), nullptr, 0.0, out, nullptr, n, nullptr);
}

// ---- communication: single rank ----
void comm_allreduce_sum(pfh_ctx*, double*, int) {}
bool comm_halo_is_separate(const pfh_ctx*) { return false; }
void comm_halo(pfh_ctx*, const double*, long long, int, double*, double*, cudaStream_t) {}
void comm_halo_n(pfh_ctx*, const double*, int, double*, double*, cudaStream_t) {}

// the solver itself, unchanged
#include "../../pfhubbenchmarks_b200/csrc/pfh_solver.cpp"

// ---------------------------------------------------------------------------------------------------------------------
// entry point: one backward-Euler step  w0 -> w  (w in: initial guess) with the product's solver on host arrays
// report[8] = niters, converged (reason > 0), reason, fnorm0, fnorm, krylov_its, residual_evals, pc_applies
// ---------------------------------------------------------------------------------------------------------------------
// the multigrid solve of the bench6 Poisson block alone: x ~ K_D^-1 rhs by `cycles` V-cycles (mg_solve, csrc/pfh_solver.cpp)
extern "C" int emul_mg_solve(const int* ncell, const double* length, int cycles, const double* rhs, double* x) {
  try {
    Setup S;
    setup(S, PFH_MESH_CROSSED2D, ncell, length, 0, -1);
    pfh_ctx c;
    memset(&c.cfg, 0, sizeof c.cfg);
    c.cfg.nranks = 1;
    c.dim = 2;
    c.m2 = S.m2;
    memset(&c.prob, 0, sizeof c.prob);
    c.prob.id = PFH_PROBLEM_BENCH6;
    c.nf = c.prob.nf = 3;
    c.n_owned = c.n_global = S.m2.n_owned;
    g_setup[&c] = &S;
    std::vector<std::vector<double>> store;
    auto alloc = [&](size_t n) {
      store.emplace_back(n, 0.0);
      return store.back().data();
    };
    for (int i = 0; i < 5; ++i) c.S[i] = alloc(c.n_owned + 32);
    std::vector<Mesh2> meshes;
    std::vector<std::vector<double>> tabs;
    pfh_mg_build_levels(c.m2, meshes, tabs);
    for (const Mesh2& m : meshes) {
      MgLevel L;
      L.m = m;
      L.x = alloc(m.n_owned + 32); L.b = alloc(m.n_owned + 32); L.t = alloc(m.n_owned + 32); L.tmp = alloc(m.n_owned + 32);
      c.mg.push_back(L);
    }
    if (c.mg.empty()) return -2;
    c.mg_cycles = cycles;
    c.scalar_bc = 1;
    mg_solve(&c, rhs, x);
    g_setup.erase(&c);
    return (int)c.mg.size();
  } catch (const std::exception& e) {
    fprintf(stderr, "emul_solver: %s\n", e.what());
    return -1;
  }
}

extern "C" int emul_newton_step(int mesh_kind, int problem, const int* ncell, const double* length, const double* params, int nparams,
                                double dt, const pfh_solver_opts* opts, int basis_precision, double* w, const double* w0,
                                double* report, int mg_cycles) {
  try {
    Setup S;
    setup(S, mesh_kind, ncell, length, 0, -1);
    pfh_ctx c;
    memset(&c.cfg, 0, sizeof c.cfg);
    c.cfg.nranks = 1;
    c.cfg.krylov_basis = 30;
    c.dim = S.dim;
    c.m2 = S.m2;
    c.m3 = S.m3;
    c.mr = S.mr;
    if (problem == PFH_PROBLEM_NLDIFF) c.nlp = NlParams{params[0], params[1], params[2], params[3]};
    memset(&c.prob, 0, sizeof c.prob);
    c.prob.id = problem;
    for (int i = 0; i < nparams; ++i) c.prob.p[i] = params[i];
    c.nf = problem == PFH_PROBLEM_BENCH2 ? 6 : (problem == PFH_PROBLEM_BENCH6 ? 3 : (problem == PFH_PROBLEM_NLDIFF ? 1 : 2));
#ifdef PFH_USER_PROBLEM_HEADER
    if (problem == PFH_PROBLEM_USER) c.nf = ProbUser::NF;
#endif
    c.prob.nf = c.nf;
    c.n_owned = c.n_global = S.dim == 3 ? S.m3.n_owned : (mesh_kind == PFH_MESH_RIGHT2D ? S.mr.n : S.m2.n_owned);
    c.dt = dt;
    c.basis_precision = basis_precision;
    g_setup[&c] = &S;
    const long long N = c.nf * c.n_owned, vs = (N + 31) / 32 * 32;
    std::vector<std::vector<double>> store;
    auto alloc = [&](size_t n) {
      store.emplace_back(n, 0.0);
      return store.back().data();
    };
    c.R = alloc(vs); c.B = alloc(vs); c.DX = alloc(vs); c.DX2 = alloc(vs); c.W1 = alloc(vs); c.W2 = alloc(vs);
    c.V = alloc((size_t)vs * (c.cfg.krylov_basis + 1));
    for (int i = 0; i < 5; ++i) c.S[i] = alloc(c.n_owned + 32);
    if (c.dim == 2 && (problem == PFH_PROBLEM_BENCH2 || problem == PFH_PROBLEM_BENCH3 || problem == PFH_PROBLEM_USER))
      for (int i = 0; i < 4; ++i) c.SIG[i] = alloc(c.n_owned + 32);
    std::vector<Mesh2> mg_meshes;
    std::vector<std::vector<double>> mg_tabs;
    c.mg_cycles = mg_cycles;
    if (mesh_kind == PFH_MESH_CROSSED2D && problem == PFH_PROBLEM_BENCH6) {   // as pfh_create: the Poisson block's hierarchy
      pfh_mg_build_levels(c.m2, mg_meshes, mg_tabs);
      for (const Mesh2& m : mg_meshes) {
        MgLevel L;
        L.m = m;
        L.x = alloc(m.n_owned + 32); L.b = alloc(m.n_owned + 32); L.t = alloc(m.n_owned + 32); L.tmp = alloc(m.n_owned + 32);
        c.mg.push_back(L);
      }
    }
    pfh_solve_report rep;
    s_newton(&c, w, w0, opts, &rep);
    g_setup.erase(&c);
    report[0] = rep.niters; report[1] = rep.reason > 0 ? 1.0 : 0.0; report[2] = rep.reason; report[3] = rep.fnorm0;
    report[4] = rep.fnorm; report[5] = rep.krylov_its; report[6] = rep.residual_evals; report[7] = rep.pc_applies;
    return 0;
  } catch (const std::exception& e) {
    fprintf(stderr, "emul_solver: %s\n", e.what());
    return -1;
  }
}
