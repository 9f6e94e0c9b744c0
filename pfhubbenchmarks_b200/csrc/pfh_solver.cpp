// Newton-Krylov driver of one backward-Euler step: the B200-native stand-in for
//   niters, converged = solver.solve()        (dolfin/bench1.py:162; PETSc SNES newtonls + KSP)
// Host logic only (control flow, Hessenberg/Givens); all O(n) work is in the CUDA kernels of
// pfh_stencil.cu / pfh_vec.cu.  Algorithm statement on the CPU: oracle/krylov.py.
//
//  * Newton: residual norm tests as SNES (atol, rtol*||F0||), 'basic' line search (full step),
//    returns (niters, converged) and never throws on non-convergence (bench1.py:86,164-177).
//  * Linear solve: right-preconditioned restarted GMRES, classical Gram-Schmidt with a fused
//    multi-dot (one reduction per iteration; ||w||^2 rides along and h_{j+1,j} comes from the
//    Pythagorean identity, with an explicit-norm fallback when cancellation is severe) and a
//    fused multi-axpy + normalisation.
//  * Preconditioner: for the Cahn-Hilliard pair (c, mu)
//        J = [[M/dt, Mob K], [-(M_f'' + kappa K), M]],
//    block factorisation with the Schur complement S = M + dt Mob K M^-1 (kappa K + M_f'')
//    approximated by S^ = (M + alpha K) M^-1 (M + alpha K), alpha = sqrt(dt Mob kappa); every
//    inner solve is a fixed number of Jacobi-Chebyshev sweeps on an SPD operator am M + ak K.
//    Other fields (Allen-Cahn eta_i, heat U, Poisson phi) get Jacobi-Chebyshev on their own
//    diagonal block.
#include <algorithm>
#include <cmath>

#include "pfh_internal.h"
#ifdef PFH_USER_PROBLEM_HEADER
#include "pfh_problems.cuh"  // ProbUser: the generated model this copy of the library is built for
#endif

// ------------------------------------------------------------------------------------------
// operator wrappers (ghost exchange + kernel)
// ------------------------------------------------------------------------------------------
// An early shipment (s_after_vector_written) may still be in flight on comm_stream: exchanges issued on the ctx stream
// use the same communicator and must be ordered after it.
static void settle_prefetch(pfh_ctx* c) {
  if (c->gR_prefetched != nullptr) {
    PFH_CUDA(cudaStreamWaitEvent(c->stream, c->ev_rhs, 0));
    c->gR_prefetched = nullptr;
  }
}

void s_exchange_state(pfh_ctx* c, const double* u, double* glo, double* ghi) {
  if (c->cfg.nranks > 1) {
    settle_prefetch(c);
    comm_halo(c, u, c->n_owned, c->nf, glo, ghi);
  }
}

static FieldRef scalar_ref(pfh_ctx* c, const double* p) {
  if (c->cfg.nranks > 1) {
    if (c->gS_prefetched == p) {
      // the producer (cheb_solve) already sent these ghost planes on comm_stream: just order the consumer after them
      PFH_CUDA(cudaStreamWaitEvent(c->stream, c->ev_halo, 0));
    } else {
      settle_prefetch(c);
      comm_halo(c, p, 0, 1, c->gS_lo, c->gS_hi);
    }
  }
  c->gS_prefetched = nullptr;
  return FieldRef{p, c->gS_lo, c->gS_hi};
}

FieldRef s_field_ref(pfh_ctx* c, const double* u) { return scalar_ref(c, u); }

void s_bilinear_apply(pfh_ctx* c, const double* u, double am, double ak, double* out) {
  k_lin_apply(c, am, ak, scalar_ref(c, u), 0.0, nullptr, out);
}

void s_residual(pfh_ctx* c, const double* u, const double* u0, double* R, bool u0_ghost_current) {
  s_exchange_state(c, u, c->gU_lo, c->gU_hi);
  if (!u0_ghost_current) s_exchange_state(c, u0, c->gU0_lo, c->gU0_hi);
  k_residual(c, u, u0, R);
}

void s_jac_setup(pfh_ctx* c, const double* u) {
  c->jac_point = u;
  s_exchange_state(c, u, c->gU_lo, c->gU_hi);
  c->pc_ready = false;
}

void s_jac_apply(pfh_ctx* c, const double* x, double* y) {
  PFH_CHECK(c->jac_point != nullptr, "jac_apply before jac_setup");
  if (c->cfg.nranks > 1 && c->gX_prefetched == x) {
    // the producer (wide-ghost preconditioner) already sent these ghost planes on comm_stream: order the consumer after them
    PFH_CUDA(cudaStreamWaitEvent(c->stream, c->ev_halo, 0));
  } else {
    s_exchange_state(c, x, c->gX_lo, c->gX_hi);
  }
  c->gX_prefetched = nullptr;
  k_jac_apply(c, c->jac_point, x, y);
}

// ------------------------------------------------------------------------------------------
// Jacobi-Chebyshev on A = am M + ak K
// ------------------------------------------------------------------------------------------
namespace {

struct ChebCoef {
  double a, b;
};

// z_{k+1} = z_k + a_k (z_k - z_{k-1}) + b_k D^-1 (r - A z_k), z_0 = z_{-1} = 0
std::vector<ChebCoef> cheb_coeffs(double lmin, double lmax, int n) {
  const double theta = 0.5 * (lmax + lmin), delta = 0.5 * (lmax - lmin), sigma = theta / delta;
  double rho = 1.0 / sigma;
  std::vector<ChebCoef> out;
  out.push_back({0.0, 1.0 / theta});
  for (int k = 1; k < n; ++k) {
    const double rn = 1.0 / (2.0 * sigma - rho);
    out.push_back({rn * rho, 2.0 * rn / delta});
    rho = rn;
  }
  return out;
}

// min over vertex types of rowsum(M)/diag(am M + ak K): Rayleigh quotient of the constant mode
double lmin_estimate(const pfh_ctx* c, double am, double ak, const std::vector<double>& mT, const std::vector<double>& kT) {
  double lo = 1e300;
  if (c->dim == 3) {
    for (int t = 0; t < 27; ++t) {
      double rs = 0.0;
      for (int o = 0; o < PFH_NOFF3; ++o) rs += mT[t * PFH_NOFF3 + o];
      const double d = am * mT[t * PFH_NOFF3] + ak * kT[t * PFH_NKOFF3];
      if (d > 0.0) lo = std::fmin(lo, am * rs / d);
    }
  } else {
    for (int t = 0; t < 9; ++t) {
      double rs = 0.0;
      for (int o = 0; o < PFH_NOFF2; ++o) rs += mT[t * PFH_NOFF2 + o];
      const double d = am * mT[t * PFH_NOFF2] + ak * kT[t * PFH_NOFF2];
      if (d > 0.0) lo = std::fmin(lo, am * rs / d);
    }
    double rs = 0.0;
    for (int o = 0; o < 5; ++o) rs += c->m2.mC[o];
    lo = std::fmin(lo, am * rs / (am * c->m2.mC[0] + ak * c->m2.kC[0]));
  }
  return lo;
}

// result is guaranteed to land in `target`; `tmp` is scratch of the same size.  Ghost planes (nranks > 1) are
// exchanged synchronously before every launch that reads a neighbour's plane; the wide-ghost path below avoids that.
void cheb_solve(pfh_ctx* c, double am, double ak, double lmin, double lmax, int nsweeps, const double* r, double* target,
                double* tmp) {
  const std::vector<ChebCoef> cf = cheb_coeffs(lmin, lmax, nsweeps);
  double* cur = ((nsweeps - 1) % 2 == 0) ? target : tmp;
  double* prev = (cur == target) ? tmp : target;
  // 3-D fast path: the first two iterates come out of one pass over r (k_cheb_first2), the loop then starts at k = 2
  const bool first2 = nsweeps >= 2 && c->use_first2 && pfh_fast3_ok(c);
  int k0 = 1;
  // with more than two sweeps the first iterate is never stored: the third one recomputes it from r (which it reads anyway)
  const bool skip_z1 = nsweeps >= 3;
  if (first2 && k_cheb_first2(c, am, ak, cf[0].b, cf[1].a, cf[1].b, scalar_ref(c, r), skip_z1 ? nullptr : cur, prev)) {
    std::swap(cur, prev);
    k0 = 2;
  } else {
    k_cheb_first(c, am, ak, cf[0].b, r, cur);
  }
  for (int k = k0; k < nsweeps; ++k) {
    const bool from_r = k0 == 2 && k == 2;
    k_cheb_sweep(c, am, ak, cf[k].a, cf[k].b, r, scalar_ref(c, cur), prev, k == 1 || from_r, from_r ? cf[0].b : 0.0);
    std::swap(cur, prev);
  }
  PFH_CHECK(cur == target, "internal: chebyshev ping-pong parity");
}

// ------------------------------------------------------------------------------------------
// Geometric multigrid for K_D x = b on the 2-D crossed mesh (K with identity rows on the Dirichlet vertices): the Poisson
// block of bench6 (dolfin/bench6.py:72-90).  V(4,4) cycle, Jacobi-Chebyshev smoothing on the upper part
// [lmax/10, lmax] of spec(D^-1 K), re-discretised coarse operators (= Galerkin, the P1 spaces are nested; transfers in
// pfh_mg2.cuh), a Chebyshev solve on the coarsest level.  Measured contraction 0.02-0.05 per cycle independent of the mesh
// size (oracle prototype, 64^2 .. 256^2; tests/test_emul_solver.py), against 0.14 for the 240 sweeps it replaces at 100^2
// -- which do not even resolve the smooth modes at 2048^2.  Every launcher works on the level whose mesh is installed in
// the ctx for the duration (LevelScope).
// ------------------------------------------------------------------------------------------
struct LevelScope {
  pfh_ctx* c;
  Mesh2 m;
  long long n;
  LevelScope(pfh_ctx* c_, const Mesh2& level) : c(c_), m(c_->m2), n(c_->n_owned) {
    c->m2 = level;
    c->n_owned = level.n_owned;
  }
  ~LevelScope() {
    c->m2 = m;
    c->n_owned = n;
  }
};
constexpr int MG_NU = 4;
constexpr int COARSE = 2;   // coarsest-level Chebyshev sweeps = COARSE/2 x its cells per direction (100^2 ends at 25^2: per-cycle
                            // residual reduction 0.012, 0.019, 0.06 with 25 sweeps; 0.012, 0.019, 0.023 with 75)
constexpr double MG_RATIO = 10.0, MG_LMAX = 2.0;

// x <- V-cycle(b) from a zero initial guess on level l (0 = the ctx mesh, l >= 1 = c->mg[l-1]); t, tmp: scratch of the level
void mg_vcycle(pfh_ctx* c, const Mesh2& fine0, size_t l, const double* b, double* x, double* t, double* tmp) {
  const Mesh2 ml = l == 0 ? fine0 : c->mg[l - 1].m;
  LevelScope scope(c, ml);
  if (l == c->mg.size()) {   // coarsest level: Chebyshev on the whole spectrum (bounds as in the single-level solve)
    const double pi = 3.14159265358979323846;
    const double lmin = 0.5 * (pi / ml.nx) * (pi / ml.nx) / 8.0;
    cheb_solve(c, 0.0, 1.0, lmin, MG_LMAX, std::max(8, COARSE * std::max(ml.nx, ml.ny) / 2), b, x, tmp);
    return;
  }
  const std::vector<ChebCoef> cf = cheb_coeffs(MG_LMAX / MG_RATIO, MG_LMAX, MG_NU);
  cheb_solve(c, 0.0, 1.0, MG_LMAX / MG_RATIO, MG_LMAX, MG_NU, b, x, tmp);                    // pre-smoothing from x = 0
  k_lin_apply(c, 0.0, 1.0, FieldRef{x, nullptr, nullptr}, -1.0, b, t);                       // t = K x - b (free rows)
  MgLevel& nx = c->mg[l];
  k2_mg_restrict(c, ml, nx.m, t, -1.0, nx.b, 1);                                             // coarse rhs = P^T (b - K x), 0 on Dirichlet rows
  mg_vcycle(c, fine0, l + 1, nx.b, nx.x, nx.t, nx.tmp);
  // (the recursion's own scope has put this level's mesh back on its way out)
  k2_mg_prolong_add(c, ml, nx.m, nx.x, x, 1);
  double *cur = x, *prev = tmp;                                                              // post-smoothing from the corrected x
  for (int k = 0; k < MG_NU; ++k) {
    k_cheb_sweep(c, 0.0, 1.0, cf[k].a, cf[k].b, b, FieldRef{cur, nullptr, nullptr}, prev, k == 0);
    std::swap(cur, prev);
  }
  static_assert(MG_NU % 2 == 0, "post-smoothing ping-pong must end in x");
}

// x ~ K_D^-1 rhs by c->mg_cycles V-cycles: the first from x = 0, the others on the residual equation x <- x + V(rhs - K_D x).
// c->scalar_bc must be 1; scratch S[0], S[1], S[3], S[4] (rhs and x must be none of them).
void mg_solve(pfh_ctx* c, const double* rhs, double* x) {
  const Mesh2 fine = c->m2;
  const long long n = c->n_owned;
  double *t = c->S[0], *tmp = c->S[1], *corr = c->S[3], *res = c->S[4];
  mg_vcycle(c, fine, 0, rhs, x, t, tmp);
  for (int cyc = 1; cyc < c->mg_cycles; ++cyc) {
    k_lin_apply(c, 0.0, 1.0, FieldRef{x, nullptr, nullptr}, -1.0, rhs, res);   // K x - rhs on the free rows (rhs on the Dirichlet rows)
    v_scale(c, -1.0, res, res, n);
    v_zero(c, corr, n);
    k_dirichlet_copy(c, corr, res);   // the Dirichlet rows already hold their values: zero correction there
    mg_vcycle(c, fine, 0, res, corr, t, tmp);
    v_axpy(c, 1.0, corr, x, n);
  }
}

// ------------------------------------------------------------------------------------------
// Wide-ghost scalar operators (3-D slabs, nranks > 1)
//
// A Chebyshev solve with n sweeps is a polynomial of degree n-1 in the 15-point operator: its result on the owned
// planes depends on the right-hand side up to n-1 planes beyond them.  Instead of exchanging one ghost plane before
// every sweep (n exchanges of 2-4 MB, each costing a launch + NVLink round trip that a 100 us sweep cannot hide), the
// right-hand side is exchanged ONCE at width n-1+d and the sweeps run redundantly on a window that shrinks by one
// plane per sweep until it reaches the depth d the consumer needs.  The scalar work vectors S[] carry `gw` spare
// planes on each side for this (pfh_ctx::gw); an operand that lives in a Krylov vector uses the wide ghost arrays gR_*.
// Redundant work: at most (n-1)/nzl of a sweep per side, ~6 % of the preconditioner on a 64-plane slab.
// ------------------------------------------------------------------------------------------
struct Win {
  int lo, hi;  // output planes [lo, hi) relative to the owned slab
};
// owned slab widened by d planes on the sides that have a neighbour
Win win_of(const pfh_ctx* c, int d) {
  const Mesh3& m = c->m3;
  return Win{-std::min(d, m.z0), m.nzl + std::min(d, m.nzv - (m.z0 + m.nzl))};
}
struct WideOperand {
  const double* p;       // owned plane 0
  bool ext;              // true: one of the S[] vectors (gw spare planes on each side, contiguous)
  const double* glo;     // !ext: separate ghost arrays holding planes -n..-1 and nzl..nzl+n-1
  const double* ghi;
  int n;
};
WideOperand ext_op(const double* p) { return WideOperand{p, true, nullptr, nullptr, 0}; }

const WideOperand kNoOperand{nullptr, true, nullptr, nullptr, 0};

void wide_launch(pfh_ctx* c, int timer, double bytes_per_vertex, int mode, double am, double ak, double a, double b, double b0,
                 const WideOperand& in, const WideOperand& r, const double* o, double* out, double* out1, Win w, double ob0 = 0.0) {
  const Mesh3& m = c->m3;
  const bool has_lo = c->cfg.rank > 0, has_hi = c->cfg.rank < c->cfg.nranks - 1;
  ScalarOp op;
  op.mode = mode; op.am = am; op.ak = ak; op.a = a; op.b = b; op.b0 = b0; op.ob0 = ob0;
  if (in.ext) {
    const int elo = has_lo ? c->gw : 0, ehi = has_hi ? c->gw : 0;
    op.in = in.p - (long long)elo * m.nxy;
    op.in_n = elo + m.nzl + ehi;
    op.in_off = elo + w.lo;
  } else {
    op.in = in.p;
    op.in_n = m.nzl;
    op.in_off = w.lo;
    if (has_lo && in.glo) { op.in_lo = in.glo; op.n_lo = in.n; }
    if (has_hi && in.ghi) { op.in_hi = in.ghi; op.n_hi = in.n; }
  }
  const long long off = (long long)w.lo * m.nxy;
  if (r.p != nullptr) {
    if (r.ext) {
      op.r = r.p + off;  // covers the window contiguously
    } else {
      op.r = r.p;
      op.r_off = w.lo;
      op.r_n = m.nzl;
      if (has_lo && r.glo) { op.r_lo = r.glo; op.n_rlo = r.n; }
      if (has_hi && r.ghi) { op.r_hi = r.ghi; op.n_rhi = r.n; }
    }
  }
  op.o = o ? o + off : nullptr;
  op.out = out + off;
  op.out1 = out1 ? out1 + off : nullptr;
  op.gz0 = m.z0 + w.lo;
  op.nzo = w.hi - w.lo;
  ProfScope ps(c, timer, bytes_per_vertex * (double)op.nzo * (double)m.nxy);
  k3_scalar_launch(c, op);
}

// exchange `w` planes of a scalar field: ext vectors receive in place, others into gR_*
void wide_exchange(pfh_ctx* c, const double* p, bool ext, int w) {
  if (w <= 0) return;
  const Mesh3& m = c->m3;
  if (ext) comm_halo_n(c, p, w, const_cast<double*>(p) - (long long)w * m.nxy, const_cast<double*>(p) + (long long)m.nzl * m.nxy);
  else comm_halo_n(c, p, w, c->gR_lo, c->gR_hi);
}

// Halo/compute overlap of the wide path: a producer whose result must travel computes the `bw` planes next to each
// slab face first; their exchange runs on the high-priority comm_stream while the interior planes are computed on the
// ctx stream; the consumer is ordered after the exchange (ev_halo).  `launch(Win)` computes a window of the result,
// `exchange(stream)` ships it.
template <class Launch, class Exchange>
void produce_and_ship(pfh_ctx* c, int bw, Launch launch, Exchange exchange) {
  const Mesh3& m = c->m3;
  if (!c->overlap || c->comm_stream == nullptr || m.nzl < 2 * bw + 2) {
    launch(Win{0, m.nzl});
    exchange(c->stream);
    return;
  }
  launch(Win{0, bw});
  launch(Win{m.nzl - bw, m.nzl});
  PFH_CUDA(cudaEventRecord(c->ev_bnd, c->stream));
  PFH_CUDA(cudaStreamWaitEvent(c->comm_stream, c->ev_bnd, 0));
  exchange(c->comm_stream);
  PFH_CUDA(cudaEventRecord(c->ev_halo, c->comm_stream));
  launch(Win{bw, m.nzl - bw});
  PFH_CUDA(cudaStreamWaitEvent(c->stream, c->ev_halo, 0));
}

// z_n = p_{n-1}(D^-1 A) D^-1 rhs on the window of depth d_out.  `rhs` must be valid to depth d_out + n - 1; iterates
// alternate between the ext vectors bufA / bufB; the last sweep writes `out` (any vector that covers the window).
// `last(Win)`, when given, replaces the plain launch of the last sweep: the caller splits it to ship the result early.
struct LastSweep {
  double am, ak, a, b, ob0;
  WideOperand z, rhs;
  const double* zprev;
  double* out;
};
void cheb_solve_wide(pfh_ctx* c, double am, double ak, double lmin, double lmax, int n, const WideOperand& rhs, int d_out,
                     double* bufA, double* bufB, double* out, LastSweep* defer_last = nullptr) {
  PFH_CHECK(n >= 2, "internal: wide Chebyshev solve needs at least two sweeps");
  const std::vector<ChebCoef> cf = cheb_coeffs(lmin, lmax, n);
  // first two iterates on the window of depth d_out + n - 2; if n == 2 the second one is the result.  With n >= 3 the
  // first iterate is not stored: the third one recomputes it from the right-hand side it reads anyway.
  double* z1 = bufA;
  double* z2 = n == 2 ? out : bufB;
  wide_launch(c, T_CHEB_FIRST, n == 2 ? 24.0 : 16.0, 3, am, ak, cf[1].a, cf[1].b, cf[0].b, rhs, kNoOperand, nullptr, z2, n == 2 ? z1 : nullptr,
              win_of(c, d_out + n - 2));
  double *cur = z2, *prev = z1;
  for (int k = 2; k < n; ++k) {
    double* dst = k == n - 1 ? out : prev;  // in place over z_{k-1} except for the last sweep
    if (k == n - 1 && defer_last != nullptr) {
      *defer_last = LastSweep{am, ak, cf[k].a, cf[k].b, k == 2 ? cf[0].b : 0.0, ext_op(cur), rhs, k == 2 ? nullptr : prev, dst};
      return;
    }
    wide_launch(c, T_CHEB_SWEEP, k == 2 ? 24.0 : 32.0, 2, am, ak, cf[k].a, cf[k].b, 0.0, ext_op(cur), rhs, k == 2 ? nullptr : prev, dst, nullptr,
                win_of(c, d_out + n - 1 - k), k == 2 ? cf[0].b : 0.0);
    prev = cur;
    cur = dst;
  }
}

// host copies of the stencil type tables, built once per ctx
void ensure_host_tables(pfh_ctx* c) {
  if (!c->h_mT.empty()) return;
  if (c->dim == 3) {
    Mesh3 m = c->m3;
    pfh_build_tables3(m, c->h_mT, c->h_kT);
  } else {
    Mesh2 m = c->m2;
    pfh_build_tables2(m, c->h_mT, c->h_kT);
  }
}

// which fields form the Cahn-Hilliard pair, and what the remaining fields' diagonal blocks are
using PcLayout = PfhPcLayout;
PcLayout pc_layout(const pfh_ctx* c) {
  PcLayout L;
  const LinCoefs lc = pfh_lin_coefs(c, c->dt);
  const double* p = c->prob.p;
  switch (c->prob.id) {
    case PFH_PROBLEM_CH:
      L.has_ch = true; L.Mob = p[4]; L.kappa = p[3];
      break;
    case PFH_PROBLEM_BENCH6:
      L.has_ch = true; L.Mob = p[4]; L.kappa = p[3];
      // the Poisson row of phi is handled separately in s_precond_apply (Dirichlet rows, many sweeps)
      break;
    case PFH_PROBLEM_BENCH2:
      L.has_ch = true; L.Mob = p[5]; L.kappa = p[3];
      for (int i = 0; i < 4; ++i) {
        L.spd_field[i] = 2 + i; L.spd_am[i] = lc.Mc[2 + i][2 + i]; L.spd_ak[i] = lc.Kc[2 + i][2 + i]; L.spd_sign[i] = 1.0;
      }
      L.n_spd = 4;
      break;
    case PFH_PROBLEM_BENCH3:
      for (int i = 0; i < 2; ++i) {
        L.spd_field[i] = i; L.spd_am[i] = lc.Mc[i][i]; L.spd_ak[i] = lc.Kc[i][i]; L.spd_sign[i] = 1.0;
      }
      L.n_spd = 2;
      break;
#ifdef PFH_USER_PROBLEM_HEADER
    case PFH_PROBLEM_USER:
      ProbUser::pc_layout(lc, L);
      break;
#endif
  }
  return L;
}

// restores the scalar-operator modifiers of the ctx however the scope is left
struct ScalarModGuard {
  pfh_ctx* c;
  int bc;
  const double* sig;
  explicit ScalarModGuard(pfh_ctx* c_) : c(c_), bc(c_->scalar_bc), sig(c_->scalar_sig) {}
  ~ScalarModGuard() {
    c->scalar_bc = bc;
    c->scalar_sig = sig;
  }
};

// the wide-ghost Schur preconditioner applies when the slabs are thick enough for the widest exchange
bool wide_ok(const pfh_ctx* c, const PcLayout& L, int nM, int nA) {
  return c->cfg.nranks > 1 && c->use_wide && pfh_fast3_ok(c) && c->prob.id == PFH_PROBLEM_CH && L.has_ch && c->pc_opts.pc_variant == 1 &&
         nM >= 2 && nA >= 3 && std::max(nM, nA) <= std::min(c->gw, c->min_slab);
}

// Block upper-triangular Schur preconditioner of the (c, mu) pair with three wide exchanges (r_mu, q, M y) instead of
// one per sweep:  t = M^-1 r_mu ;  q = dt r_c - dt Mob K t ;  dc = A^-1 M A^-1 q ;  dmu = t
void precond_apply_ch_wide(pfh_ctx* c, const PcLayout& L, int nM, int nA, const double* r, double* z) {
  const long long n = c->n_owned;
  const double dt = c->dt, alpha = c->pc_alpha;
  const double *r1 = r, *r2 = r + n;
  double *S0 = c->S[0], *S1 = c->S[1], *S2 = c->S[2], *S3 = c->S[3], *S4 = c->S[4];
  // t = M^-1 r_mu to depth 1 (K t is needed on the owned planes): r_mu travels at width nM -- already on its way if the
  // Krylov loop announced the vector when it was written (s_after_vector_written)
  if (c->gR_prefetched == r2 && c->gR_prefetched_w == nM) PFH_CUDA(cudaStreamWaitEvent(c->stream, c->ev_rhs, 0));
  else wide_exchange(c, r2, false, nM);
  c->gR_prefetched = nullptr;
  cheb_solve_wide(c, 1.0, 0.0, c->lmin_m, c->lmax_m, nM, WideOperand{r2, false, c->gR_lo, c->gR_hi, nM}, 1, S1, S3, S0);
  // q = dt r_c - dt Mob K t on the owned planes; the same pass copies t into z_mu (dmu = t).
  // y = A^-1 q to depth 1 (M y is needed on the owned planes): q travels at width nA, under the interior planes of its producer
  const Mesh3& m = c->m3;
  auto ship_S2 = [&](int w) {
    return [c, S2, w, &m](cudaStream_t st) {
      comm_halo_n(c, S2, w, S2 - (long long)w * m.nxy, S2 + (long long)m.nzl * m.nxy, st);
    };
  };
  produce_and_ship(c, nA, [&](Win w) {
    wide_launch(c, T_LIN_APPLY, 32.0, 0, 0.0, -dt * L.Mob, dt, 0.0, 0.0, ext_op(S0), ext_op(r1), nullptr, S2, z + n, w);
  }, ship_S2(nA));
  cheb_solve_wide(c, 1.0, alpha, c->lmin_a, c->lmax_a, nA, ext_op(S2), 1, S3, S4, S1);
  // dc = A^-1 (M y) on the owned planes: M y travels at width nA - 1
  produce_and_ship(c, nA - 1, [&](Win w) {
    wide_launch(c, T_LIN_APPLY, 16.0, 0, 1.0, 0.0, 0.0, 0.0, 0.0, ext_op(S1), kNoOperand, nullptr, S2, nullptr, w);
  }, ship_S2(nA - 1));
  // the last sweep writes dc = z_c; its slab-face planes go first and the ghost planes of z = (dc, dmu) travel into gX under
  // the interior planes, for the J.v that follows a preconditioner application in GMRES
  LastSweep ls;
  cheb_solve_wide(c, 1.0, alpha, c->lmin_a, c->lmax_a, nA, ext_op(S2), 0, S3, S4, z, &ls);
  produce_and_ship(c, 1, [&](Win w) {
    wide_launch(c, T_CHEB_SWEEP, ls.zprev ? 32.0 : 24.0, 2, ls.am, ls.ak, ls.a, ls.b, 0.0, ls.z, ls.rhs, ls.zprev, ls.out, nullptr, w, ls.ob0);
  }, [c, z](cudaStream_t st) { comm_halo(c, z, c->n_owned, c->nf, c->gX_lo, c->gX_hi, st); });
  c->gX_prefetched = z;
}

}  // namespace

// Extreme values of the symbol of D^-1 (am M + ak K) on the interior rows of the Kuhn mesh: the spectrum of the
// Jacobi-scaled operator up to boundary effects of a fraction of a per cent (checked against the assembled operator in
// tests/test_oracle_properties.py).  The generic bounds, (d + 2)/2 from above and the Rayleigh quotient of the constant
// from below, are up to 30 % wide of this interval for am M + ak K with ak > 0.
static void symbol_bounds3(const Mesh3& m, double am, double ak, double& lo, double& hi) {
  static const int DX[PFH_NOFF3] = {0, 1, -1, 0, 0, 0, 0, 1, -1, 0, 0, 1, -1, 1, -1};
  static const int DY[PFH_NOFF3] = {0, 0, 0, 1, -1, 0, 0, 1, -1, 1, -1, 0, 0, 1, -1};
  static const int DZ[PFH_NOFF3] = {0, 0, 0, 0, 0, 1, -1, 0, 0, 1, -1, 1, -1, 1, -1};
  double cf[PFH_NOFF3];
  for (int o = 0; o < PFH_NOFF3; ++o) cf[o] = am * m.mI[o] + (o < PFH_NKOFF3 ? ak * m.kI[o] : 0.0);
  const int ng = 24;
  const double pi = 3.14159265358979323846;
  lo = 1e300;
  hi = -1e300;
  for (int i = 0; i <= ng; ++i)
    for (int j = -ng; j <= ng; ++j)
      for (int k = -ng; k <= ng; ++k) {
        const double tx = pi * i / ng, ty = pi * j / ng, tz = pi * k / ng;
        double s = 0.0;
        for (int o = 0; o < PFH_NOFF3; ++o) s += cf[o] * std::cos(DX[o] * tx + DY[o] * ty + DZ[o] * tz);
        lo = std::fmin(lo, s);
        hi = std::fmax(hi, s);
      }
  lo /= cf[0];
  hi /= cf[0];
}

// GMRES has just launched the kernel that writes its next basis vector `vec` and is about to block on that vector's
// norm: start the wide exchange of vec's mu half -- the first thing the preconditioner will need -- on comm_stream now.
// Only with a separate halo communicator (the reduction that follows runs on the other one, concurrently).
void s_after_vector_written(pfh_ctx* c, const double* vec) {
  if (c->cfg.nranks <= 1 || !c->overlap || c->comm_stream == nullptr || !c->pc_ready || !comm_halo_is_separate(c)) return;
  const pfh_solver_opts& o = c->pc_opts;
  if (o.preconditioner != 0) return;
  const PcLayout L = pc_layout(c);
  const int nM = o.pc_sweeps_mass > 0 ? o.pc_sweeps_mass : 3;
  const int nA = o.pc_sweeps_shifted > 0 ? o.pc_sweeps_shifted : ((c->dim == 3 && c->cheb_bounds == 1) ? 4 : 5);
  if (!wide_ok(c, L, nM, nA)) return;
  const double* r2 = vec + c->n_owned;
  PFH_CUDA(cudaEventRecord(c->ev_bnd, c->stream));
  PFH_CUDA(cudaStreamWaitEvent(c->comm_stream, c->ev_bnd, 0));
  comm_halo_n(c, r2, nM, c->gR_lo, c->gR_hi, c->comm_stream);
  PFH_CUDA(cudaEventRecord(c->ev_rhs, c->comm_stream));
  c->gR_prefetched = r2;
  c->gR_prefetched_w = nM;
}

// PFH_PROBLEM_NLDIFF (dolfin/nl.py): J = -(K_{kc|u|} + reac M + C(u)) with identity Dirichlet rows.  z = P^-1 r with
// P = -A, A = K_{kc|u|} + reac M (the SPD part, frozen at the linearisation point): n Jacobi-Chebyshev sweeps on the
// upper part [lmax / (5 n), lmax] of spec(D^-1 A), lmax = 2 (both the stiffness and the mass part of a P1 triangle mesh
// satisfy x.Ax <= 2 x.Dx); the Krylov method deals with the smooth modes the polynomial leaves.  The sweeps run on -r and
// keep z = r on the Dirichlet vertices, so their columns enter the neighbouring rows exactly as in J.  On the reference
// mesh (60 x 75) 12 sweeps give 9-12 GMRES iterations per Newton step (the reference's gmres + sor is order dependent).
void precond_apply_nl(pfh_ctx* c, const double* r, double* z) {
  const pfh_solver_opts& o = c->pc_opts;
  if (o.preconditioner == 1) {
    k_jacobi(c, c->jac_point, r, z);
    return;
  }
  const int n = o.pc_sweeps_shifted > 0 ? o.pc_sweeps_shifted : 12;
  const double lmax = 2.0, lmin = lmax / (5.0 * n);
  const std::vector<ChebCoef> cf = cheb_coeffs(lmin, lmax, n);
  // ping-pong between z and a scratch vector so that sweep n writes z
  double* P = (n & 1) ? z : c->S[0];
  double* Q = (n & 1) ? c->S[0] : z;
  v_zero(c, Q, c->n_owned);   // z_0 = 0; z_{-1} is not read by the first sweep (a_0 = 0)
  for (int k = 0; k < n; ++k) {
    k_nl_sweep(c, c->jac_point, cf[k].a, cf[k].b, -1.0, r, Q, P);
    std::swap(P, Q);
  }
}

void s_precond_setup(pfh_ctx* c, const pfh_solver_opts* o) {
  c->pc_opts = *o;
  if (pfh_is_nl(c)) {   // nothing to prepare: the sweeps read the coefficient from the linearisation point
    c->pc_ready = true;
    return;
  }
  ensure_host_tables(c);
  const PcLayout L = pc_layout(c);
  c->lmax = 0.5 * (c->dim + 2);
  c->lmin_m = 0.5;
  c->lmax_m = c->lmax_a = c->lmax;
  if (L.has_ch) {
    c->pc_alpha = std::sqrt(c->dt * L.Mob * L.kappa);
    c->lmin_a = std::fmin(0.5, lmin_estimate(c, 1.0, c->pc_alpha, c->h_mT, c->h_kT));
    if (c->dim == 3 && c->cheb_bounds == 1) {
      // upper end from the symbol; the lower end stays at the Rayleigh estimate (about 0.76 of the symbol's minimum):
      // scanned on the CPU solver leg, this pair needs one sweep less per solve than the generic interval
      double lo, hi;
      symbol_bounds3(c->m3, 1.0, c->pc_alpha, lo, hi);
      c->lmax_a = c->cheb_fhi * hi;
      if (c->cheb_flo > 0.0) c->lmin_a = c->cheb_flo * lo;
    }
  }
  for (int k = 0; k < L.n_spd; ++k)
    // with the clipped reaction shift the constant mode's Rayleigh quotient is at least half the unshifted one
    c->lmin_spd[k] = 0.5 * std::fmin(0.5, lmin_estimate(c, L.spd_am[k], L.spd_ak[k], c->h_mT, c->h_kT));
  // lumped reaction term d g_a / d u_a of the parabolic fields' diagonal blocks at the linearisation point
  // (Allen-Cahn eta_i of bench2, U and phi of bench3); clipped at -1/2 of the block's mass coefficient
  if (c->jac_point != nullptr)
    for (int k = 0; k < L.n_spd; ++k)
      if (c->SIG[k] != nullptr) k_reaction_diag(c, c->jac_point, L.spd_field[k], -0.5 * L.spd_am[k], c->SIG[k]);
  c->pc_ready = true;
}

void s_precond_apply(pfh_ctx* c, const double* r, double* z) {
  PFH_CHECK(c->pc_ready, "precond_apply before precond_setup");
  c->gX_prefetched = nullptr;
  if (c->gR_prefetched != nullptr && c->gR_prefetched != r + c->n_owned) {
    // a prefetch for another vector is in flight into gR: let it land before gR is reused
    PFH_CUDA(cudaStreamWaitEvent(c->stream, c->ev_rhs, 0));
    c->gR_prefetched = nullptr;
  }
  const long long n = c->n_owned;
  const pfh_solver_opts& o = c->pc_opts;
  if (o.preconditioner == 2) {
    v_copy(c, z, r, c->nf * n);
    return;
  }
  if (pfh_is_nl(c)) {
    PFH_CHECK(c->jac_point != nullptr, "precond_apply before jac_setup");
    precond_apply_nl(c, r, z);
    return;
  }
  const PcLayout L = pc_layout(c);
  if (o.preconditioner == 1 || L.jacobi_only) {
    k_jacobi(c, c->jac_point, r, z);
    return;
  }
  const int nM = o.pc_sweeps_mass > 0 ? o.pc_sweeps_mass : 3;
  const int nA = o.pc_sweeps_shifted > 0 ? o.pc_sweeps_shifted : ((c->dim == 3 && c->cheb_bounds == 1) ? 4 : 5);
  ScalarModGuard guard(c);
  if (L.has_ch && wide_ok(c, L, nM, nA)) {
    precond_apply_ch_wide(c, L, nM, nA, r, z);
  } else if (L.has_ch) {
    const double dt = c->dt, alpha = c->pc_alpha;
    const double *r1 = r, *r2 = r + n;
    double *S0 = c->S[0], *S1 = c->S[1], *S2 = c->S[2], *S3 = c->S[3], *S4 = c->S[4];
    // t = M^-1 r2   (block upper-triangular variant: t IS dmu, so it is solved straight into z_mu)
    double* T = o.pc_variant == 1 ? z + n : S0;
    cheb_solve(c, 1.0, 0.0, c->lmin_m, c->lmax_m, nM, r2, T, S1);
    // q = dt r1 - dt Mob K t
    k_lin_apply(c, 0.0, -dt * L.Mob, scalar_ref(c, T), dt, r1, S2);
    // dc = A^-1 M A^-1 q,  A = M + alpha K
    cheb_solve(c, 1.0, alpha, c->lmin_a, c->lmax_a, nA, S2, S3, S4);
    k_lin_apply(c, 1.0, 0.0, scalar_ref(c, S3), 0.0, nullptr, S2);
    cheb_solve(c, 1.0, alpha, c->lmin_a, c->lmax_a, nA, S2, z, S4);
    if (o.pc_variant != 1) {
      // dmu = t + M^-1 (kappa K + M_f'') dc = t - M^-1 [J (dc, 0, ..)]_mu
      if (k_ch_crow(c, c->jac_point, scalar_ref(c, z), S2)) {
        // fused (mu, c) block apply: reads c and dc only, writes the mu row only
        cheb_solve(c, 1.0, 0.0, c->lmin_m, c->lmax_m, nM, S2, S3, S1);
        v_axpby(c, 1.0, S0, -1.0, S3, z + n, n);
      } else {
        double *PX = c->W2, *PY = c->R;  // free during a linear solve (see gmres below)
        v_zero(c, PX, c->nf * n);
        v_copy(c, PX, z, n);
        s_jac_apply(c, PX, PY);
        cheb_solve(c, 1.0, 0.0, c->lmin_m, c->lmax_m, nM, PY + n, S2, S1);
        v_axpby(c, 1.0, S0, -1.0, S2, z + n, n);
      }
    }
  }
  for (int k = 0; k < L.n_spd; ++k) {
    const int f = L.spd_field[k];
    c->scalar_sig = (c->SIG[k] != nullptr && c->jac_point != nullptr) ? c->SIG[k] : nullptr;
    cheb_solve(c, L.spd_am[k], L.spd_ak[k], c->lmin_spd[k], c->lmax, nA, r + f * n, z + f * n, c->S[1]);
    c->scalar_sig = nullptr;
    if (L.spd_sign[k] < 0.0) v_scale(c, -1.0, z + f * n, z + f * n, n);
  }
  if (c->prob.id == PFH_PROBLEM_BENCH6) {
    // phi row of J: (k/eps) M dc - K dphi = r_phi, identity on the Dirichlet vertices (bench6.py:72-90).
    // Block lower-triangular step: dphi = K_D^-1 ((k/eps) M dc - r_phi), K_D = K with Dirichlet rows, solved by
    // Jacobi-Chebyshev with the analytic bounds of the Dirichlet/Neumann Laplacian on the crossed mesh
    // (a multigrid cycle would do this in O(1) sweeps: "next" row N4).
    const double k_eps = c->prob.p[5] / c->prob.p[6];
    double *rhs = c->S[2], *tmp = c->S[1];
    c->scalar_bc = 1;
    v_scale(c, -1.0, r + 2 * n, c->S[0], n);                                        // -r_phi
    k_lin_apply(c, k_eps, 0.0, FieldRef{z, nullptr, nullptr}, 1.0, c->S[0], rhs);  // (k/eps) M dc - r_phi (interior rows)
    k_dirichlet_copy(c, r + 2 * n, rhs);                                            // identity rows: dphi = r_phi
    if (c->mg_cycles > 0 && !c->mg.empty()) {
      mg_solve(c, rhs, z + 2 * n);   // scratch: S[0], S[1], S[3], S[4]
    } else {
      const double pi = 3.14159265358979323846;
      const double lmin = 0.5 * (pi / c->m2.nx) * (pi / c->m2.nx) / 8.0;
      cheb_solve(c, 0.0, 1.0, lmin, 2.0, 240, rhs, z + 2 * n, tmp);
    }
  }
}

// ------------------------------------------------------------------------------------------
// GMRES
// ------------------------------------------------------------------------------------------
namespace {

struct KspResult {
  int its;
  double relres;
  bool converged;
};

// Solves J x = b. Workspace use: V (basis), W1 (w), W2 (z / precond scratch), R is free for the
// preconditioner because the Newton loop has already moved -R into `b` (= DX's companion buffer).
//
KspResult gmres_compressed(pfh_ctx* c, const double* b, double* x, const pfh_solver_opts& o, pfh_solve_report* rep);

// Hessenberg matrix of one GMRES cycle, kept upper triangular by Givens rotations; g carries the rotated right-hand
// side, |g[j+1]| is the residual norm after column j.
struct Arnoldi {
  int m;
  std::vector<double> H, cs, sn, g, y;
  explicit Arnoldi(int m_) : m(m_), H((size_t)(m_ + 1) * m_, 0.0), cs(m_), sn(m_), g(m_ + 1), y(m_) {}
  double& at(int i, int j) { return H[(size_t)i * m + j]; }
  void start(double beta) {
    std::fill(g.begin(), g.end(), 0.0);
    g[0] = beta;
  }
  // append column j (h[0..j] projections, hn the sub-diagonal entry); returns the new residual norm estimate
  double push(int j, const double* h, double hn) {
    for (int i = 0; i <= j; ++i) at(i, j) = h[i];
    at(j + 1, j) = hn;
    for (int i = 0; i < j; ++i) {
      const double t = cs[i] * at(i, j) + sn[i] * at(i + 1, j);
      at(i + 1, j) = -sn[i] * at(i, j) + cs[i] * at(i + 1, j);
      at(i, j) = t;
    }
    const double den = std::hypot(at(j, j), at(j + 1, j));
    cs[j] = at(j, j) / den;
    sn[j] = at(j + 1, j) / den;
    at(j, j) = den;
    at(j + 1, j) = 0.0;
    g[j + 1] = -sn[j] * g[j];
    g[j] = cs[j] * g[j];
    return std::fabs(g[j + 1]);
  }
  // y = H^-1 g for the first jd columns
  void solve(int jd) {
    for (int i = jd - 1; i >= 0; --i) {
      double t = g[i];
      for (int k = i + 1; k < jd; ++k) t -= at(i, k) * y[k];
      y[i] = t / at(i, i);
    }
  }
};

KspResult gmres(pfh_ctx* c, const double* b, double* x, const pfh_solver_opts& o, pfh_solve_report* rep) {
  const long long N = c->nf * c->n_owned;
  const long long vs = (N + 31) / 32 * 32;
  const int m = std::max(1, std::min(o.krylov_restart > 0 ? o.krylov_restart : 30, c->cfg.krylov_basis));
  double* V = c->V;
  double* w = c->W1;
  // the fp32 basis and its fp64 working copy share the basis allocation: that needs room for two vectors behind half of it
  const bool fits = m >= 2 && vs >= 64;
  if (fits && (c->basis_precision == 2 || (c->basis_precision == 0 && o.krylov_rtol >= 1e-6))) return gmres_compressed(c, b, x, o, rep);
  Arnoldi A(m);
  std::vector<double> h(m + 2);

  v_zero(c, x, N);
  const double bnorm = std::sqrt(v_norm2(c, b, N));
  if (bnorm == 0.0) return {0, 0.0, true};
  const double tol = std::fmax(o.krylov_rtol * bnorm, o.krylov_atol);
  double beta = bnorm;
  int its = 0;
  bool first_cycle = true;
  while (true) {
    // V0 = r / beta  (first cycle: r = b)
    if (first_cycle) v_scale(c, 1.0 / beta, b, V, N);
    first_cycle = false;
    A.start(beta);
    int jd = 0;
    bool done = false;
    for (int j = 0; j < m; ++j) {
      double* Vj = V + (long long)j * vs;
      double* Vn = V + (long long)(j + 1) * vs;
      s_precond_apply(c, Vj, c->DX2);
      rep->pc_applies++;
      s_jac_apply(c, c->DX2, w);
      rep->jac_applies++;
      v_multidot(c, V, vs, j + 1, w, N, h.data());
      const double ww = h[j + 1];
      double s2 = 0.0;
      for (int i = 0; i <= j; ++i) s2 += h[i] * h[i];
      double hn2 = ww - s2;
      double hn;
      if (!(hn2 > 1e-8 * ww)) {
        // severe cancellation: orthogonalise, then take the norm explicitly
        v_multiaxpy(c, V, vs, j + 1, h.data(), w, 1.0, Vn, N);
        hn = std::sqrt(v_norm2(c, Vn, N));
        if (hn > 0.0) v_scale(c, 1.0 / hn, Vn, Vn, N);
      } else {
        hn = std::sqrt(hn2);
        v_multiaxpy(c, V, vs, j + 1, h.data(), w, 1.0 / hn, Vn, N);
      }
      const double res = A.push(j, h.data(), hn);
      ++its;
      jd = j + 1;
      if (o.verbose > 1) fprintf(stderr, "      gmres it %3d  |r|/|b| = %.3e\n", its, res / bnorm);
      if (!(res == res)) return {its, NAN, false};
      if (res <= tol || its >= o.krylov_maximum_iterations) {
        done = true;
        break;
      }
    }
    // y = H^-1 g ; x += P^-1 (V y)
    A.solve(jd);
    v_combine(c, V, vs, jd, A.y.data(), w, N);
    s_precond_apply(c, w, c->DX2);
    rep->pc_applies++;
    v_axpy(c, 1.0, c->DX2, x, N);
    const double est = std::fabs(A.g[jd]);
    if (done) return {its, est / bnorm, est <= tol};
    // restart: true residual r = b - J x -> V0
    s_jac_apply(c, x, w);
    rep->jac_applies++;
    v_axpby(c, 1.0, b, -1.0, w, w, N);
    beta = std::sqrt(v_norm2(c, w, N));
    if (beta <= tol) return {its, beta / bnorm, true};
    v_scale(c, 1.0 / beta, w, V, N);
  }
}

// Compressed-basis GMRES.  The orthogonalisation (multi-dot + multi-axpy over the whole basis) is pure HBM
// streaming and a third of the step, so at the default linear tolerance the basis is STORED in fp32 while all
// arithmetic stays fp64.  Stored vector u_i is the fp32 rounding of a multiple of the Arnoldi vector, u_i = s_i v_i
// with s_i = ||u_i|| measured (in the pass that writes it) on the rounded values; the fp64 working copy handed to
// the preconditioner is that same rounded vector.  With w~ = J P^-1 u_j = s_j A v_j and d_i = u_i . w~:
//     h_ij = d_i / (s_i s_j),   r~ = w~ - sum_i (d_i / s_i^2) u_i = s_j h_{j+1,j} v_{j+1},
//     u_{j+1} = round(a r~)  (a = 1 / Pythagoras estimate of ||r~||, only to keep s near 1),   h_{j+1,j} = s_{j+1} / (a s_j).
// The Arnoldi relation then holds for the stored basis up to the rounding of u_{j+1} (6e-8 relative): no norm is
// ever taken from the Pythagorean difference, which loses all accuracy against a basis that is orthonormal only to
// fp32 (measured: +26 % iterations, tests/probe/README.md).  The true residual is recomputed in fp64 at every restart.
KspResult gmres_compressed(pfh_ctx* c, const double* b, double* x, const pfh_solver_opts& o, pfh_solve_report* rep) {
  const long long N = c->nf * c->n_owned;
  const long long vs = (N + 31) / 32 * 32;
  const int m = std::max(1, std::min(o.krylov_restart > 0 ? o.krylov_restart : 30, c->cfg.krylov_basis));
  float* U = reinterpret_cast<float*>(c->V);                            // fp32 basis shares the allocation ...
  double* ucur = c->V + ((long long)(m + 1) * vs / 2 + 31) / 32 * 32;   // ... and the fp64 working copy sits behind it
  double* w = c->W1;
  Arnoldi A(m);
  std::vector<double> d(m + 2), sv(m + 1), cf(m + 1), h(m + 1);

  v_zero(c, x, N);
  const double bnorm = std::sqrt(v_norm2(c, b, N));
  if (bnorm == 0.0) return {0, 0.0, true};
  const double tol = std::fmax(o.krylov_rtol * bnorm, o.krylov_atol);
  double beta = bnorm;
  const double* r = b;
  int its = 0;
  while (true) {
    double s2 = 0.0;
    v_multiaxpy32(c, U, vs, 0, nullptr, r, 1.0 / beta, ucur, U, N, &s2);   // u_0 = round(r / beta)
    sv[0] = std::sqrt(s2);
    A.start(beta * sv[0]);                                                 // r = beta u_0 = beta s_0 v_0
    int jd = 0;
    bool done = false;
    for (int j = 0; j < m; ++j) {
      s_precond_apply(c, ucur, c->DX2);
      rep->pc_applies++;
      s_jac_apply(c, c->DX2, w);
      rep->jac_applies++;
      v_multidot32(c, U, vs, j + 1, w, N, d.data());
      const double ww = d[j + 1];
      double est = ww;
      for (int i = 0; i <= j; ++i) {
        cf[i] = d[i] / (sv[i] * sv[i]);
        est -= d[i] * d[i] / (sv[i] * sv[i]);
      }
      const double a = ww > 0.0 ? 1.0 / std::sqrt(std::fmax(est, 1e-8 * ww)) : 1.0;
      v_multiaxpy32(c, U, vs, j + 1, cf.data(), w, a, ucur, U + (long long)(j + 1) * vs, N, &s2);
      sv[j + 1] = std::sqrt(s2);
      const double hn = sv[j + 1] / (a * sv[j]);
      for (int i = 0; i <= j; ++i) h[i] = d[i] / (sv[i] * sv[j]);
      const double res = A.push(j, h.data(), hn);
      ++its;
      jd = j + 1;
      if (o.verbose > 1) fprintf(stderr, "      gmres it %3d  |r|/|b| = %.3e\n", its, res / bnorm);
      if (!(res == res)) return {its, NAN, false};
      if (res <= tol || its >= o.krylov_maximum_iterations || !(sv[j + 1] > 0.0)) {
        done = true;
        break;
      }
    }
    // y = H^-1 g ; x += P^-1 (sum_j (y_j / s_j) u_j)
    A.solve(jd);
    for (int i = 0; i < jd; ++i) cf[i] = A.y[i] / sv[i];
    v_combine32(c, U, vs, jd, cf.data(), w, N);
    s_precond_apply(c, w, c->DX2);
    rep->pc_applies++;
    v_axpy(c, 1.0, c->DX2, x, N);
    const double estr = std::fabs(A.g[jd]);
    if (done) return {its, estr / bnorm, estr <= tol};
    // restart: true residual r = b - J x
    s_jac_apply(c, x, w);
    rep->jac_applies++;
    v_axpby(c, 1.0, b, -1.0, w, w, N);
    beta = std::sqrt(v_norm2(c, w, N));
    if (beta <= tol) return {its, beta / bnorm, true};
    r = w;
  }
}

// Right-preconditioned BiCGStab (PETSc KSPBCGS, the reduced Ostwald drivers' choice: dolfin/bench2_eta1.py:113).  Two J.v and
// two preconditioner applications per iteration, no growing basis: the orthogonalisation passes of GMRES are traded for
// a second operator apply.  Work vectors live in the (otherwise idle) Krylov-basis allocation.  Dots are the fused
// two-stage reductions of pfh_vec.cu (r^.v and v.v, t.s and t.t come out of one pass each).
KspResult bicgstab(pfh_ctx* c, const double* b, double* x, const pfh_solver_opts& o, pfh_solve_report* rep) {
  const long long N = c->nf * c->n_owned;
  const long long vs = (N + 31) / 32 * 32;
  PFH_CHECK(c->cfg.krylov_basis + 1 >= 6, "bicgstab needs a ctx created with krylov_basis >= 5 (six work vectors)");
  double *r = c->V, *rhat = c->V + vs, *p = c->V + 2 * vs, *v = c->V + 3 * vs, *phat = c->V + 4 * vs, *shat = c->V + 5 * vs;
  double* t = c->W1;
  double d[2];
  v_zero(c, x, N);
  v_copy(c, r, b, N);
  v_copy(c, rhat, b, N);
  const double bnorm = std::sqrt(v_norm2(c, b, N));
  if (bnorm == 0.0) return {0, 0.0, true};
  const double tol = std::fmax(o.krylov_rtol * bnorm, o.krylov_atol);
  double rho = 1.0, alpha = 1.0, omega = 1.0, rnorm = bnorm;
  v_zero(c, p, N);
  v_zero(c, v, N);
  int its = 0;
  while (its < o.krylov_maximum_iterations) {
    v_multidot(c, rhat, 0, 1, r, N, d);                 // rhat . r
    const double rho_new = d[0];
    if (rho_new == 0.0 || !(rho_new == rho_new)) return {its, rnorm / bnorm, false};   // breakdown
    const double beta = (rho_new / rho) * (alpha / omega);
    v_axpby(c, 1.0, p, -omega, v, p, N);                // p = r + beta (p - omega v)
    v_axpby(c, 1.0, r, beta, p, p, N);
    s_precond_apply(c, p, phat);
    rep->pc_applies++;
    s_jac_apply(c, phat, v);
    rep->jac_applies++;
    v_multidot(c, rhat, 0, 1, v, N, d);                 // rhat . v
    if (d[0] == 0.0 || !(d[0] == d[0])) return {its, rnorm / bnorm, false};
    alpha = rho_new / d[0];
    v_axpby(c, 1.0, r, -alpha, v, r, N);                // s = r - alpha v   (kept in r)
    const double snorm = std::sqrt(v_norm2(c, r, N));
    ++its;
    if (snorm <= tol) {                                 // half step already converged
      v_axpy(c, alpha, phat, x, N);
      if (o.verbose > 1) fprintf(stderr, "      bicgstab it %3d  |s|/|b| = %.3e\n", its, snorm / bnorm);
      return {its, snorm / bnorm, true};
    }
    s_precond_apply(c, r, shat);
    rep->pc_applies++;
    s_jac_apply(c, shat, t);
    rep->jac_applies++;
    v_multidot(c, r, 0, 1, t, N, d);                    // s . t and t . t in one pass
    if (d[1] == 0.0) return {its, snorm / bnorm, false};
    omega = d[0] / d[1];
    v_axpy(c, alpha, phat, x, N);                       // x += alpha phat + omega shat
    v_axpy(c, omega, shat, x, N);
    v_axpy(c, -omega, t, r, N);                         // r = s - omega t
    rnorm = std::sqrt(v_norm2(c, r, N));
    if (o.verbose > 1) fprintf(stderr, "      bicgstab it %3d  |r|/|b| = %.3e\n", its, rnorm / bnorm);
    if (!(rnorm == rnorm)) return {its, NAN, false};
    if (rnorm <= tol) return {its, rnorm / bnorm, true};
    if (omega == 0.0) return {its, rnorm / bnorm, false};
    rho = rho_new;
  }
  return {its, rnorm / bnorm, false};
}

}  // namespace

// ------------------------------------------------------------------------------------------
// Newton
// ------------------------------------------------------------------------------------------
void s_newton(pfh_ctx* c, double* w, const double* w0, const pfh_solver_opts* opts, pfh_solve_report* rep) {
  const pfh_solver_opts& o = *opts;
  const long long N = c->nf * c->n_owned;
  memset(rep, 0, sizeof *rep);
  double* U = w;  // iterate in place in the caller's buffer (dolfin: w is the solution Function)
  k_dirichlet_fix(c, U);
  s_exchange_state(c, w0, c->gU0_lo, c->gU0_hi);
  double fnorm = 0.0;
  for (int it = 0;; ++it) {
    s_residual(c, U, w0, c->R, true);
    rep->residual_evals++;
    fnorm = std::sqrt(v_norm2(c, c->R, N));
    if (it == 0) rep->fnorm0 = fnorm;
    rep->fnorm = fnorm;
    rep->niters = it;
    if (o.verbose) fprintf(stderr, "  newton it %2d  |F| = %.6e\n", it, fnorm);
    if (!(fnorm == fnorm) || std::isinf(fnorm)) {
      rep->reason = -4;
      break;
    }
    if (fnorm < o.absolute_tolerance) {
      rep->reason = 2;
      break;
    }
    if (it > 0 && fnorm < o.relative_tolerance * rep->fnorm0) {
      rep->reason = 3;
      break;
    }
    if (it >= o.maximum_iterations) {
      rep->reason = -5;
      break;
    }
    // J dx = -R
    c->jac_point = U;  // ghosts of U are current from s_residual
    s_precond_setup(c, opts);
    v_scale(c, -1.0, c->R, c->B, N);
    const KspResult ks = o.linear_solver == 1 ? bicgstab(c, c->B, c->DX, o, rep) : gmres(c, c->B, c->DX, o, rep);
    rep->krylov_its += ks.its;
    rep->last_krylov_relres = ks.relres;
    if (o.verbose) fprintf(stderr, "    gmres: %d its, relres %.3e%s\n", ks.its, ks.relres, ks.converged ? "" : " (NOT converged)");
    if (!(ks.relres == ks.relres)) {
      rep->reason = -3;
      break;
    }
    // PETSc's SNES flags a failed linear solve; with error_on_nonconvergence=False the reference
    // carries on only if KSP converged, so a Krylov failure ends the solve as diverged.
    if (!ks.converged) {
      rep->reason = -3;
      break;
    }
    v_axpy(c, 1.0, c->DX, U, N);  // 'basic' line search: full step
    if (o.line_search == 1) {
      // 'cp' (critical point, PETSc SNESLINESEARCHCP with its default single secant iteration; bench2.py:140,
      // bench6.py:117): find lambda with F(x + lambda dx) . dx = 0 from the values at lambda = 0 and 1
      double d0[2], d1[2];
      v_multidot(c, c->DX, 0, 1, c->B, N, d0);   // B = -F(x) survives the linear solve (R may serve as its scratch)
      d0[0] = -d0[0];                            // F(x) . dx
      s_residual(c, U, w0, c->W1, true);
      rep->residual_evals++;
      v_multidot(c, c->DX, 0, 1, c->W1, N, d1);  // F(x + dx) . dx
      const double den = d1[0] - d0[0];
      double lambda = 1.0;
      if (std::fabs(den) > 1e-300) lambda = -d0[0] / den;
      if (!(lambda == lambda) || lambda <= 1e-3 || lambda > 4.0) lambda = 1.0;  // guard: keep the full step
      if (o.verbose) fprintf(stderr, "    cp line search: lambda = %.4f\n", lambda);
      if (lambda != 1.0) v_axpy(c, lambda - 1.0, c->DX, U, N);
    }
    if (o.line_search == 2) {
      // 'bt' (PETSc SNESLINESEARCHBT, its default; the drivers mention it as working, bench3.py:127): backtracking on
      // phi(lambda) = 1/2 |F(x + lambda dx)|^2 with the Armijo condition phi(lambda) <= phi(0) + alpha lambda phi'(0),
      // alpha = 1e-4; first a quadratic, then cubic fits through the last two trial points, each new lambda clamped to
      // [0.1, 0.5] of the previous one, at most 40 trials.  phi'(0) = F . (J dx) is taken from one extra J.v.
      const double alpha = 1e-4, f2 = fnorm * fnorm;
      double d[2];
      s_jac_apply(c, c->DX, c->W2);
      rep->jac_applies++;
      v_multidot(c, c->W2, 0, 1, c->B, N, d);     // B = -F(x) survives the linear solve (R may serve as its scratch)
      double slope = -d[0];
      if (slope > 0.0) slope = -slope;
      if (slope == 0.0) slope = -1.0;
      auto gnorm2_at = [&](double lam_new, double lam_cur) {   // moves U from x + lam_cur dx to x + lam_new dx
        if (lam_new != lam_cur) v_axpy(c, lam_new - lam_cur, c->DX, U, N);
        s_residual(c, U, w0, c->W1, true);
        rep->residual_evals++;
        return v_norm2(c, c->W1, N);
      };
      double lam = 1.0, g2 = gnorm2_at(1.0, 1.0);
      if (!(0.5 * g2 <= 0.5 * f2 + alpha * lam * slope)) {
        double lam_prev = lam, g2_prev = g2;
        double lt = -slope / (g2 - f2 - 2.0 * slope);             // minimiser of the quadratic through phi(0), phi'(0), phi(1)
        lt = std::fmin(std::fmax(lt, 0.1 * lam), 0.5 * lam);
        if (!(lt == lt)) lt = 0.5 * lam;
        g2 = gnorm2_at(lt, lam);
        lam = lt;
        for (int k = 0; k < 40 && !(0.5 * g2 <= 0.5 * f2 + alpha * lam * slope); ++k) {
          const double t1 = 0.5 * (g2 - f2) - lam * slope, t2 = 0.5 * (g2_prev - f2) - lam_prev * slope;
          const double a = (t1 / (lam * lam) - t2 / (lam_prev * lam_prev)) / (lam - lam_prev);
          const double b = (-lam_prev * t1 / (lam * lam) + lam * t2 / (lam_prev * lam_prev)) / (lam - lam_prev);
          double disc = b * b - 3.0 * a * slope;
          if (disc < 0.0) disc = 0.0;
          lt = a == 0.0 ? -slope / (2.0 * b) : (-b + std::sqrt(disc)) / (3.0 * a);
          lt = std::fmin(std::fmax(lt, 0.1 * lam), 0.5 * lam);
          if (!(lt == lt)) lt = 0.5 * lam;
          lam_prev = lam;
          g2_prev = g2;
          g2 = gnorm2_at(lt, lam);
          lam = lt;
        }
      }
      if (o.verbose) fprintf(stderr, "    bt line search: lambda = %.4g\n", lam);
    }
    if (o.solution_tolerance > 1e-14) {
      const double dn = std::sqrt(v_norm2(c, c->DX, N)), un = std::sqrt(v_norm2(c, U, N));
      if (dn < o.solution_tolerance * un) {
        rep->niters = it + 1;
        rep->reason = 4;
        break;
      }
    }
  }
  rep->converged = rep->reason > 0;
  c->jac_point = nullptr;
  PFH_CUDA(cudaStreamSynchronize(c->stream));
}
