// extern "C" boundary (include/pfh_b200.h): argument checking, workspace ownership, error
// translation.  No field data is owned here except internal workspaces and, for the *_host
// entry points, device staging buffers for the caller's host arrays.
#include <algorithm>
#include <cmath>
#include <mutex>

#include "pfh_internal.h"
#ifdef PFH_USER_PROBLEM_HEADER
#include "pfh_problems.cuh"  // ProbUser: the generated model this copy of the library is built for
#endif

static thread_local std::string g_err;
void pfh_set_error(const std::string& m) { g_err = m; }

#define PFH_API_BEGIN try {
#define PFH_API_END                 \
  }                                 \
  catch (const std::exception& e) { \
    pfh_set_error(e.what());        \
    return -1;                      \
  }                                 \
  return 0;

double* pfh_alloc(pfh_ctx* c, size_t n) {
  double* p = nullptr;
  PFH_CUDA(cudaMalloc(&p, (n > 0 ? n : 1) * sizeof(double)));
  c->allocs.push_back(p);
  return p;
}

void pfh_prof_begin(pfh_ctx* c, int id, cudaStream_t s) {
  EventPair ep;
  if (!c->ev_free.empty()) {
    ep = c->ev_free.back();
    c->ev_free.pop_back();
  } else {
    PFH_CUDA(cudaEventCreate(&ep.a));
    PFH_CUDA(cudaEventCreate(&ep.b));
  }
  ep.id = id;
  if (c->ev_pending.size() >= 8192) pfh_prof_resolve(c);
  PFH_CUDA(cudaEventRecord(ep.a, s ? s : c->stream));
  c->ev_pending.push_back(ep);
}
void pfh_prof_end(pfh_ctx* c, int id, cudaStream_t s) {
  // scopes nest at most trivially (reduce inside multidot is sequential, not nested): close the last open one of this id
  for (size_t i = c->ev_pending.size(); i-- > 0;)
    if (c->ev_pending[i].id == id) {
      cudaEventRecord(c->ev_pending[i].b, s ? s : c->stream);  // called from a destructor: must not throw
      c->ev_pending[i].id = id | 0x100;  // closed
      break;
    }
}
void pfh_prof_resolve(pfh_ctx* c) {
  if (c->ev_pending.empty()) return;
  PFH_CUDA(cudaStreamSynchronize(c->stream));
  if (c->comm_stream) PFH_CUDA(cudaStreamSynchronize(c->comm_stream));
  for (EventPair& ep : c->ev_pending) {
    if (ep.id & 0x100) {
      float ms = 0.f;
      PFH_CUDA(cudaEventElapsedTime(&ms, ep.a, ep.b));
      c->t_ms[ep.id & 0xff] += ms;
      c->t_n[ep.id & 0xff] += 1;
    }
    c->ev_free.push_back(ep);
  }
  c->ev_pending.clear();
}

extern "C" {

int pfh_abi_version(void) { return PFH_ABI_VERSION; }
const char* pfh_last_error(void) { return g_err.c_str(); }

void pfh_default_solver_opts(pfh_solver_opts* o) {
  memset(o, 0, sizeof *o);
  o->absolute_tolerance = 1e-6;   // drivers: nlparams['absolute_tolerance'] = 1e-6 (bench1.py:87)
  o->relative_tolerance = 1e-9;   // DOLFIN PETScSNESSolver default
  o->solution_tolerance = 1e-16;  // DOLFIN default
  o->maximum_iterations = 10;     // bench1.py:88
  o->line_search = 0;
  o->krylov_rtol = 1e-5;          // PETSc KSP default
  o->krylov_atol = 1e-50;
  o->krylov_maximum_iterations = 1000;  // bench1.py:110
  o->krylov_restart = 30;               // PETSc GMRES default
  o->preconditioner = 0;
  o->pc_variant = 1;        // block upper-triangular: best time-to-solution in the 256^3 sweep (profiles/r1_pc_sweep_256.txt)
  o->pc_sweeps_mass = 3;
  o->pc_sweeps_shifted = 0;   // automatic: 5, or 4 in 3-D where the Chebyshev interval comes from the operator's symbol (pfh_solver.cpp)
  o->verbose = 0;
  o->linear_solver = 0;
}

int pfh_nccl_unique_id(void* out) {
  PFH_API_BEGIN
  PFH_CHECK(out != nullptr, "null output");
  comm_unique_id(out);
  PFH_API_END
}

int pfh_create(const pfh_config* cfg, pfh_ctx** out) {
  pfh_ctx* c = nullptr;
  try {
    PFH_CHECK(cfg && out, "null argument");
    c = new pfh_ctx;
    c->cfg = *cfg;
    c->cfg.nccl_unique_id = nullptr;
    PFH_CHECK(cfg->mesh_kind == PFH_MESH_CROSSED2D || cfg->mesh_kind == PFH_MESH_KUHN3D || cfg->mesh_kind == PFH_MESH_RIGHT2D,
              "unknown mesh kind");
    PFH_CHECK((cfg->mesh_kind == PFH_MESH_RIGHT2D) == (cfg->problem == PFH_PROBLEM_NLDIFF),
              "the right-diagonal mesh carries PFH_PROBLEM_NLDIFF (dolfin/nl.py) and nothing else does");
    c->dim = cfg->mesh_kind == PFH_MESH_KUHN3D ? 3 : 2;
    for (int d = 0; d < c->dim; ++d) PFH_CHECK(cfg->ncell[d] >= 1 && cfg->length[d] > 0.0, "bad mesh size");
    memset(&c->prob, 0, sizeof c->prob);
    c->prob.id = cfg->problem;
    int need = 0;
    switch (cfg->problem) {
      case PFH_PROBLEM_CH: c->nf = 2; need = 5; break;
      case PFH_PROBLEM_BENCH2: c->nf = 6; need = 9; break;
      case PFH_PROBLEM_BENCH3: c->nf = 2; need = 5; break;
      case PFH_PROBLEM_BENCH6: c->nf = 3; need = 7; break;
      case PFH_PROBLEM_NLDIFF: c->nf = 1; need = 4; break;
#ifdef PFH_USER_PROBLEM_HEADER
      case PFH_PROBLEM_USER: c->nf = ProbUser::NF; need = 0; break;
#endif
      default: throw PfhError("unknown problem id (a generated model needs the library built for it)");
    }
    PFH_CHECK(cfg->n_params >= need && cfg->n_params <= PFH_MAX_PARAMS, "wrong number of problem parameters");
    c->prob.nf = c->nf;
    for (int i = 0; i < cfg->n_params; ++i) c->prob.p[i] = cfg->params[i];
    PFH_CHECK(c->dim == 2 || cfg->problem == PFH_PROBLEM_CH || cfg->problem == PFH_PROBLEM_USER,
              "3-D meshes support PFH_PROBLEM_CH (b13d) and generated models only");
    PFH_CHECK(cfg->nranks >= 1 && cfg->rank >= 0 && cfg->rank < cfg->nranks, "bad rank/nranks");
    PFH_CHECK(c->dim == 3 || cfg->nranks == 1, "2-D problems do not shard (replicas only)");

    PFH_CUDA(cudaSetDevice(cfg->device));
    PFH_CUDA(cudaDeviceGetAttribute(&c->n_sm, cudaDevAttrMultiProcessorCount, cfg->device));
    if (c->dim == 3) {  // kernel attributes are per device: set them for the device this ctx lives on
      pfh_scalar3_init_device();
      pfh_jac3_init_device();
      PFH_CUDA(cudaStreamCreateWithFlags(&c->aux_stream, cudaStreamNonBlocking));
      PFH_CUDA(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
      PFH_CUDA(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
    }
    if (cfg->stream) {
      c->stream = (cudaStream_t)cfg->stream;
    } else {
      PFH_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
      c->own_stream = true;
    }

    std::vector<double> mT, kT;
    if (c->dim == 3) {
      Mesh3& m = c->m3;
      m.nxv = cfg->ncell[0] + 1; m.nyv = cfg->ncell[1] + 1; m.nzv = cfg->ncell[2] + 1;
      int zb = cfg->z_begin, ze = cfg->z_end;
      if (cfg->nranks == 1 && zb == 0 && ze == 0) ze = m.nzv;
      PFH_CHECK(zb >= 0 && ze <= m.nzv && ze - zb >= 2, "each rank must own at least two z-planes");
      PFH_CHECK(cfg->nranks > 1 || (zb == 0 && ze == m.nzv), "single rank must own all planes");
      m.z0 = zb; m.nzl = ze - zb;
      m.nxy = (long long)m.nxv * m.nyv; m.n_owned = m.nxy * m.nzl;
      m.hx = cfg->length[0] / cfg->ncell[0]; m.hy = cfg->length[1] / cfg->ncell[1]; m.hz = cfg->length[2] / cfg->ncell[2];
      m.detJ = m.hx * m.hy * m.hz;
      pfh_build_tables3(m, mT, kT);
      c->n_owned = m.n_owned; c->n_global = m.nxy * m.nzv;
      c->nplane = cfg->nranks > 1 ? m.nxy : 0;
    } else if (cfg->mesh_kind == PFH_MESH_RIGHT2D) {
      MeshR2& m = c->mr;
      m.nx = cfg->ncell[0]; m.ny = cfg->ncell[1]; m.nxv = m.nx + 1; m.nyv = m.ny + 1;
      m.n = (long long)m.nxv * m.nyv;
      m.hx = cfg->length[0] / m.nx; m.hy = cfg->length[1] / m.ny;
      c->nlp = NlParams{cfg->params[0], cfg->params[1], cfg->params[2], cfg->params[3]};
      PFH_CHECK(c->nlp.kc > 0.0 && c->nlp.reac >= 0.0, "nl.py-type problem needs kc > 0 and reac >= 0");
      memset(&c->m2, 0, sizeof c->m2);
      c->n_owned = c->n_global = m.n;
    } else {
      Mesh2& m = c->m2;
      m.nx = cfg->ncell[0]; m.ny = cfg->ncell[1]; m.nxv = m.nx + 1; m.nyv = m.ny + 1;
      m.ncorner = (long long)m.nxv * m.nyv; m.ncentre = (long long)m.nx * m.ny; m.n_owned = m.ncorner + m.ncentre;
      m.hx = cfg->length[0] / m.nx; m.hy = cfg->length[1] / m.ny; m.detJ = 0.5 * m.hx * m.hy;
      pfh_build_tables2(m, mT, kT);
      c->n_owned = c->n_global = m.n_owned;
    }
    c->d_tables = pfh_alloc(c, mT.size() + kT.size());
    PFH_CUDA(cudaMemcpy(c->d_tables, mT.data(), mT.size() * sizeof(double), cudaMemcpyHostToDevice));
    PFH_CUDA(cudaMemcpy(c->d_tables + mT.size(), kT.data(), kT.size() * sizeof(double), cudaMemcpyHostToDevice));
    if (c->dim == 3) { c->m3.mT = c->d_tables; c->m3.kT = c->d_tables + mT.size(); }
    else { c->m2.mT = c->d_tables; c->m2.kT = c->d_tables + mT.size(); }
    if (cfg->mesh_kind == PFH_MESH_CROSSED2D && cfg->problem == PFH_PROBLEM_BENCH6) {
      // coarse levels of the multigrid cycle for the Poisson block (pfh_solver.cpp): meshes, their stencil tables, work vectors
      std::vector<Mesh2> meshes;
      std::vector<std::vector<double>> tabs;
      pfh_mg_build_levels(c->m2, meshes, tabs);
      for (size_t l = 0; l < meshes.size(); ++l) {
        MgLevel L;
        L.m = meshes[l];
        double* d = pfh_alloc(c, tabs[l].size());
        PFH_CUDA(cudaMemcpy(d, tabs[l].data(), tabs[l].size() * sizeof(double), cudaMemcpyHostToDevice));
        L.m.mT = d;
        L.m.kT = d + tabs[l].size() / 2;
        L.x = pfh_alloc(c, L.m.n_owned + 32); L.b = pfh_alloc(c, L.m.n_owned + 32);
        L.t = pfh_alloc(c, L.m.n_owned + 32); L.tmp = pfh_alloc(c, L.m.n_owned + 32);
        c->mg.push_back(L);
      }
    }

    if (c->cfg.krylov_basis <= 0) c->cfg.krylov_basis = 30;
    PFH_CHECK(c->cfg.krylov_basis + 2 <= PFH_RED_MAXVALS, "krylov_basis too large");
    const long long N = c->nf * c->n_owned;
    const long long vs = (N + 31) / 32 * 32;
    c->R = pfh_alloc(c, vs); c->B = pfh_alloc(c, vs); c->DX = pfh_alloc(c, vs); c->DX2 = pfh_alloc(c, vs);
    c->W1 = pfh_alloc(c, vs); c->W2 = pfh_alloc(c, vs);
    c->V = pfh_alloc(c, (size_t)vs * (c->cfg.krylov_basis + 1));
    // scalar work vectors; on a slab (nranks > 1) with `gw` spare planes on each side for the wide-ghost preconditioner
    if (c->dim == 3 && cfg->nranks > 1) {
      c->gw = 8;
      c->min_slab = c->m3.nzv / cfg->nranks;  // slab_range (engine.py) balances the planes: every rank owns at least this many
    }
    const size_t ext = (size_t)c->gw * (size_t)c->nplane;
    for (int i = 0; i < 5; ++i) c->S[i] = pfh_alloc(c, c->n_owned + 2 * ext + 32) + ext;
    if (c->dim == 2 && (cfg->problem == PFH_PROBLEM_BENCH2 || cfg->problem == PFH_PROBLEM_BENCH3 || cfg->problem == PFH_PROBLEM_USER))
      for (int i = 0; i < 4; ++i) c->SIG[i] = pfh_alloc(c, c->n_owned + 32);
    if (c->nplane) {
      const size_t g = (size_t)c->nf * c->nplane;
      c->gU_lo = pfh_alloc(c, g); c->gU_hi = pfh_alloc(c, g);
      c->gU0_lo = pfh_alloc(c, g); c->gU0_hi = pfh_alloc(c, g);
      c->gX_lo = pfh_alloc(c, g); c->gX_hi = pfh_alloc(c, g);
      c->gS_lo = pfh_alloc(c, c->nplane); c->gS_hi = pfh_alloc(c, c->nplane);
      c->gR_lo = pfh_alloc(c, (size_t)c->gw * c->nplane); c->gR_hi = pfh_alloc(c, (size_t)c->gw * c->nplane);
    }
    c->d_partial = pfh_alloc(c, (size_t)PFH_RED_MAXBLOCKS * PFH_RED_MAXVALS);
    c->d_red = pfh_alloc(c, PFH_RED_MAXVALS);
    PFH_CUDA(cudaHostAlloc(&c->h_red, (PFH_RED_MAXVALS + 8) * sizeof(double), cudaHostAllocMapped));
    memset(c->h_red, 0, (PFH_RED_MAXVALS + 8) * sizeof(double));
    c->h_seq = reinterpret_cast<unsigned long long*>(c->h_red + PFH_RED_MAXVALS);
    pfh_default_solver_opts(&c->pc_opts);

    pfh_config cc = *cfg;
    c->comm = comm_create(cc, c->stream);
    if (cfg->nranks > 1) {
      int lo_prio = 0, hi_prio = 0;
      PFH_CUDA(cudaDeviceGetStreamPriorityRange(&lo_prio, &hi_prio));
      PFH_CUDA(cudaStreamCreateWithPriority(&c->comm_stream, cudaStreamNonBlocking, hi_prio));
      PFH_CUDA(cudaEventCreateWithFlags(&c->ev_bnd, cudaEventDisableTiming));
      PFH_CUDA(cudaEventCreateWithFlags(&c->ev_halo, cudaEventDisableTiming));
      PFH_CUDA(cudaEventCreateWithFlags(&c->ev_rhs, cudaEventDisableTiming));
    }
    if (cfg->nranks > 1 && c->dim == 3 && cfg->nranks <= PFH_RED_MAXVALS) {
      // smallest slab over the ranks, from the ranks themselves (the caller may have chosen any split): every rank adds
      // its height at its own index
      std::vector<double> h(cfg->nranks, 0.0);
      h[cfg->rank] = (double)c->m3.nzl;
      PFH_CUDA(cudaMemcpyAsync(c->d_red, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
      comm_allreduce_sum(c, c->d_red, cfg->nranks);
      PFH_CUDA(cudaMemcpyAsync(h.data(), c->d_red, h.size() * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
      PFH_CUDA(cudaStreamSynchronize(c->stream));
      int mn = c->m3.nzl;
      for (double v : h) mn = std::min(mn, (int)(v + 0.5));
      c->min_slab = mn;
    }
    PFH_CUDA(cudaStreamSynchronize(c->stream));
    *out = c;
  } catch (const std::exception& e) {
    pfh_set_error(e.what());
    if (c) pfh_destroy(c);
    return -1;
  }
  return 0;
}

int pfh_destroy(pfh_ctx* c) {
  if (!c) return 0;
  cudaSetDevice(c->cfg.device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  if (c->comm_stream) cudaStreamSynchronize(c->comm_stream);
  comm_destroy(c->comm);
  if (c->aux_stream) cudaStreamSynchronize(c->aux_stream);
  if (c->ev_fork) cudaEventDestroy(c->ev_fork);
  if (c->ev_join) cudaEventDestroy(c->ev_join);
  if (c->aux_stream) cudaStreamDestroy(c->aux_stream);
  if (c->ev_bnd) cudaEventDestroy(c->ev_bnd);
  if (c->ev_halo) cudaEventDestroy(c->ev_halo);
  if (c->ev_rhs) cudaEventDestroy(c->ev_rhs);
  if (c->comm_stream) cudaStreamDestroy(c->comm_stream);
  pfh_tmap_cache_free(c);
  for (double* p : c->allocs) cudaFree(p);
  if (c->h_red) cudaFreeHost(c->h_red);
  for (EventPair& e : c->ev_pending) { cudaEventDestroy(e.a); cudaEventDestroy(e.b); }
  for (EventPair& e : c->ev_free) { cudaEventDestroy(e.a); cudaEventDestroy(e.b); }
  if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
  delete c;
  return 0;
}

int pfh_n_fields(const pfh_ctx* c) { return c ? c->nf : -1; }
long long pfh_n_owned(const pfh_ctx* c) { return c ? c->n_owned : -1; }
long long pfh_n_global(const pfh_ctx* c) { return c ? c->n_global : -1; }

int pfh_set_dt(pfh_ctx* c, double dt) {
  PFH_API_BEGIN
  PFH_CHECK(c && dt > 0.0 && std::isfinite(dt), "dt must be positive");
  c->dt = dt;
  c->pc_ready = false;
  PFH_API_END
}
double pfh_get_dt(const pfh_ctx* c) { return c ? c->dt : NAN; }

int pfh_initial_condition_dev(pfh_ctx* c, const double* icp, int n, double* w) {
  PFH_API_BEGIN
  PFH_CHECK(c && icp && w, "null argument");
  PFH_CHECK(!pfh_is_nl(c), "PFH_PROBLEM_NLDIFF has no initial-condition expression (nl.py:48 fills the vector with a constant)");
  PFH_CUDA(cudaSetDevice(c->cfg.device));
  k_initial_condition(c, icp, n, w);
  PFH_API_END
}

int pfh_residual_dev(pfh_ctx* c, const double* w, const double* w0, double* F) {
  PFH_API_BEGIN
  PFH_CHECK(c && w && w0 && F, "null argument");
  PFH_CUDA(cudaSetDevice(c->cfg.device));
  s_residual(c, w, w0, F, false);
  PFH_CUDA(cudaStreamSynchronize(c->stream));
  PFH_API_END
}

int pfh_jac_setup_dev(pfh_ctx* c, const double* w) {
  PFH_API_BEGIN
  PFH_CHECK(c && w, "null argument");
  PFH_CUDA(cudaSetDevice(c->cfg.device));
  s_jac_setup(c, w);
  PFH_API_END
}

int pfh_jac_apply_dev(pfh_ctx* c, const double* x, double* y) {
  PFH_API_BEGIN
  PFH_CHECK(c && x && y, "null argument");
  PFH_CUDA(cudaSetDevice(c->cfg.device));
  s_jac_apply(c, x, y);
  PFH_CUDA(cudaStreamSynchronize(c->stream));
  PFH_API_END
}

int pfh_precond_setup(pfh_ctx* c, const pfh_solver_opts* o) {
  PFH_API_BEGIN
  PFH_CHECK(c && o, "null argument");
  s_precond_setup(c, o);
  PFH_API_END
}

int pfh_precond_apply_dev(pfh_ctx* c, const double* r, double* z) {
  PFH_API_BEGIN
  PFH_CHECK(c && r && z, "null argument");
  PFH_CUDA(cudaSetDevice(c->cfg.device));
  s_precond_apply(c, r, z);
  PFH_CUDA(cudaStreamSynchronize(c->stream));
  c->gX_prefetched = nullptr;   // the caller owns z from here on: its ghost planes are exchanged again when it is used
  PFH_API_END
}

int pfh_newton_solve_dev(pfh_ctx* c, double* w, const double* w0, const pfh_solver_opts* o, pfh_solve_report* rep) {
  PFH_API_BEGIN
  PFH_CHECK(c && w && w0 && o && rep, "null argument");
  PFH_CHECK(o->krylov_restart <= c->cfg.krylov_basis, "krylov_restart exceeds the basis the ctx was created for");
  PFH_CUDA(cudaSetDevice(c->cfg.device));
  s_newton(c, w, w0, o, rep);
  PFH_API_END
}

int pfh_energy_dev(pfh_ctx* c, const double* w, double out[4]) {
  PFH_API_BEGIN
  PFH_CHECK(c && w && out, "null argument");
  PFH_CHECK(!pfh_is_nl(c), "PFH_PROBLEM_NLDIFF has no energy functional");
  PFH_CUDA(cudaSetDevice(c->cfg.device));
  s_exchange_state(c, w, c->gU_lo, c->gU_hi);
  k_energy(c, w, out);
  PFH_API_END
}

int pfh_bilinear_dev(pfh_ctx* c, const double* w, int fa, int fb, double am, double ak, double* out) {
  PFH_API_BEGIN
  PFH_CHECK(c && w && out, "null argument");
  PFH_CHECK(fa >= 0 && fa < c->nf && fb >= 0 && fb < c->nf, "field index out of range");
  PFH_CHECK(!pfh_is_nl(c), "no tabulated mass / stiffness operators on the right-diagonal mesh");
  PFH_CUDA(cudaSetDevice(c->cfg.device));
  const long long n = c->n_owned;
  const int bc = c->scalar_bc;
  const double* sig = c->scalar_sig;
  c->scalar_bc = 0;       // the plain operators: no Dirichlet rows, no reaction diagonal
  c->scalar_sig = nullptr;
  s_bilinear_apply(c, w + fb * n, am, ak, c->S[0]);
  c->scalar_bc = bc;
  c->scalar_sig = sig;
  v_copy(c, c->S[1], w + fa * n, n);   // field offsets are only 8-byte aligned when n_owned is odd
  double dots[2];
  v_multidot(c, c->S[0], 0, 1, c->S[1], n, dots);
  *out = dots[0];
  PFH_API_END
}

int pfh_eval_points_dev(pfh_ctx* c, const double* w, int field, const double* pts, long long npts, double* out) {
  PFH_API_BEGIN
  PFH_CHECK(c && w && (npts == 0 || (pts && out)), "null argument");
  PFH_CHECK(field >= 0 && field < c->nf && npts >= 0, "field index out of range");
  PFH_CHECK(npts < (1LL << 31), "too many points for one call");
  PFH_CHECK(!pfh_is_nl(c), "point evaluation is not implemented on the right-diagonal mesh");
  PFH_CUDA(cudaSetDevice(c->cfg.device));
  k_eval_points(c, s_field_ref(c, w + (long long)field * c->n_owned), pts, npts, out);
  if (c->cfg.nranks > 1 && npts > 0) comm_allreduce_sum(c, out, (int)npts);
  PFH_API_END
}

static void stage(pfh_ctx* c) {
  const long long N = c->nf * c->n_owned;
  if (!c->U) {
    c->U = pfh_alloc(c, N);
    c->U0 = pfh_alloc(c, N);
  }
}

int pfh_step_host(pfh_ctx* c, double* w_host, const double* w0_host, double dt, const pfh_solver_opts* o,
                  pfh_solve_report* rep, double diag[4]) {
  PFH_API_BEGIN
  PFH_CHECK(c && w_host && w0_host && o && rep, "null argument");
  PFH_CUDA(cudaSetDevice(c->cfg.device));
  stage(c);
  const size_t bytes = (size_t)c->nf * c->n_owned * sizeof(double);
  PFH_CUDA(cudaMemcpyAsync(c->U, w_host, bytes, cudaMemcpyHostToDevice, c->stream));
  PFH_CUDA(cudaMemcpyAsync(c->U0, w0_host, bytes, cudaMemcpyHostToDevice, c->stream));
  PFH_CHECK(dt > 0.0, "dt must be positive");
  c->dt = dt;
  c->pc_ready = false;
  s_newton(c, c->U, c->U0, o, rep);
  if (diag && pfh_is_nl(c)) {
    diag[0] = diag[1] = diag[2] = diag[3] = 0.0;   // no functional is assembled after nl.py's solves
  } else if (diag) {
    s_exchange_state(c, c->U, c->gU_lo, c->gU_hi);
    k_energy(c, c->U, diag);
  }
  PFH_CUDA(cudaMemcpyAsync(w_host, c->U, bytes, cudaMemcpyDeviceToHost, c->stream));
  PFH_CUDA(cudaStreamSynchronize(c->stream));
  PFH_API_END
}

int pfh_residual_host(pfh_ctx* c, const double* w_host, const double* w0_host, double* F_host) {
  PFH_API_BEGIN
  PFH_CHECK(c && w_host && w0_host && F_host, "null argument");
  PFH_CUDA(cudaSetDevice(c->cfg.device));
  stage(c);
  const size_t bytes = (size_t)c->nf * c->n_owned * sizeof(double);
  PFH_CUDA(cudaMemcpyAsync(c->U, w_host, bytes, cudaMemcpyHostToDevice, c->stream));
  PFH_CUDA(cudaMemcpyAsync(c->U0, w0_host, bytes, cudaMemcpyHostToDevice, c->stream));
  s_residual(c, c->U, c->U0, c->R, false);
  PFH_CUDA(cudaMemcpyAsync(F_host, c->R, bytes, cudaMemcpyDeviceToHost, c->stream));
  PFH_CUDA(cudaStreamSynchronize(c->stream));
  PFH_API_END
}

int pfh_jac_apply_host(pfh_ctx* c, const double* w_host, const double* x_host, double* y_host) {
  PFH_API_BEGIN
  PFH_CHECK(c && w_host && x_host && y_host, "null argument");
  PFH_CUDA(cudaSetDevice(c->cfg.device));
  stage(c);
  const size_t bytes = (size_t)c->nf * c->n_owned * sizeof(double);
  PFH_CUDA(cudaMemcpyAsync(c->U, w_host, bytes, cudaMemcpyHostToDevice, c->stream));
  PFH_CUDA(cudaMemcpyAsync(c->U0, x_host, bytes, cudaMemcpyHostToDevice, c->stream));
  s_jac_setup(c, c->U);
  s_jac_apply(c, c->U0, c->R);
  PFH_CUDA(cudaMemcpyAsync(y_host, c->R, bytes, cudaMemcpyDeviceToHost, c->stream));
  PFH_CUDA(cudaStreamSynchronize(c->stream));
  c->jac_point = nullptr;
  PFH_API_END
}

int pfh_energy_host(pfh_ctx* c, const double* w_host, double out[4]) {
  PFH_API_BEGIN
  PFH_CHECK(c && w_host && out, "null argument");
  PFH_CHECK(!pfh_is_nl(c), "PFH_PROBLEM_NLDIFF has no energy functional");
  PFH_CUDA(cudaSetDevice(c->cfg.device));
  stage(c);
  const size_t bytes = (size_t)c->nf * c->n_owned * sizeof(double);
  PFH_CUDA(cudaMemcpyAsync(c->U, w_host, bytes, cudaMemcpyHostToDevice, c->stream));
  s_exchange_state(c, c->U, c->gU_lo, c->gU_hi);
  k_energy(c, c->U, out);
  PFH_API_END
}

long long pfh_launch_count(const pfh_ctx* c) { return c ? c->launches : -1; }

int pfh_set_option(pfh_ctx* c, const char* key, double value) {
  PFH_API_BEGIN
  PFH_CHECK(c && key, "null argument");
  const std::string k(key);
  if (k == "profiling") {
    if (c->profiling && value == 0.0) pfh_prof_resolve(c);
    c->profiling = value != 0.0;
  } else if (k == "reset_launch_count") {
    c->launches = 0;
  } else if (k == "overlap") {
    c->overlap = value != 0.0;
  } else if (k == "scalar_variant") {
    c->scalar_variant = (int)value;
  } else if (k == "scalar_zchunk") {
    c->scalar_zchunk = (int)value;
  } else if (k == "jac_zchunk") {
    c->jac_zchunk = (int)value;
  } else if (k == "jac_variant") {
    c->jac_variant = (int)value;
  } else if (k == "shell_async") {
    c->shell_async = value != 0.0;
  } else if (k == "use_wide") {
    c->use_wide = value != 0.0;
  } else if (k == "use_tma_loader") {
    c->use_tma_loader = value != 0.0;
  } else if (k == "cheb_bounds") {
    c->cheb_bounds = (int)value;
    c->pc_ready = false;
  } else if (k == "cheb_flo") {
    c->cheb_flo = value;
    c->pc_ready = false;
  } else if (k == "cheb_fhi") {
    c->cheb_fhi = value;
    c->pc_ready = false;
  } else if (k == "basis_precision") {
    PFH_CHECK(value == 0.0 || value == 1.0 || value == 2.0, "basis_precision: 0 auto, 1 fp64, 2 fp32");
    c->basis_precision = (int)value;
  } else if (k == "jac_quadrature_fp32") {
    c->jac_quadrature_fp32 = value != 0.0;
  } else if (k == "use_first2") {
    c->use_first2 = value != 0.0;
  } else if (k == "use_tma") {
    c->use_tma = value != 0.0;
  } else if (k == "mg_cycles") {
    PFH_CHECK(value >= 0.0 && value <= 8.0, "mg_cycles: 0 (Chebyshev solve) .. 8 V-cycles per application");
    c->mg_cycles = (int)value;
    c->pc_ready = false;
  } else if (k == "op2_tiled") {
    c->op2_tiled = value != 0.0;
  } else if (k == "nl_tiled") {
    c->nl_tiled = value != 0.0;
  } else {
    throw PfhError("unknown option: " + k);
  }
  PFH_API_END
}

static const char* kTimerNames[PFH_N_TIMERS] = {"residual", "jac_apply", "cheb_sweep", "cheb_first", "lin_apply", "multidot",
                                               "multiaxpy", "vec", "reduce_final", "energy", "halo", "allreduce",
                                               "initial_condition", "jacobi", "copy", "other"};
const char* pfh_timer_name(int i) { return (i >= 0 && i < PFH_N_TIMERS) ? kTimerNames[i] : ""; }

int pfh_timers(pfh_ctx* c, double ms[PFH_N_TIMERS], long long launches[PFH_N_TIMERS], double bytes[PFH_N_TIMERS], int reset) {
  PFH_API_BEGIN
  PFH_CHECK(c, "null argument");
  pfh_prof_resolve(c);
  for (int i = 0; i < PFH_N_TIMERS; ++i) {
    if (ms) ms[i] = c->t_ms[i];
    if (launches) launches[i] = c->t_n[i];
    if (bytes) bytes[i] = c->t_bytes[i];
    if (reset) { c->t_ms[i] = 0.0; c->t_n[i] = 0; c->t_bytes[i] = 0.0; }
  }
  PFH_API_END
}

void* pfh_stream(pfh_ctx* c) { return c ? (void*)c->stream : nullptr; }
int pfh_synchronize(pfh_ctx* c) {
  PFH_API_BEGIN
  PFH_CHECK(c, "null argument");
  PFH_CUDA(cudaStreamSynchronize(c->stream));
  PFH_API_END
}

}  // extern "C"
