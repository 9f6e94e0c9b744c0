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
#include <cmath>

#include "pfh_internal.h"
#ifdef PFH_USER_PROBLEM_HEADER
#include "pfh_problems.cuh"  // ProbUser: the generated model this copy of the library is built for
#endif

// ------------------------------------------------------------------------------------------
// operator wrappers (ghost exchange + kernel)
// ------------------------------------------------------------------------------------------
void s_exchange_state(pfh_ctx* c, const double* u, double* glo, double* ghi) {
  if (c->cfg.nranks > 1) comm_halo(c, u, c->n_owned, c->nf, glo, ghi);
}

static FieldRef scalar_ref(pfh_ctx* c, const double* p) {
  if (c->cfg.nranks > 1) {
    if (c->gS_prefetched == p) {
      // the producer (cheb_solve) already sent these ghost planes on comm_stream: just order the consumer after them
      PFH_CUDA(cudaStreamWaitEvent(c->stream, c->ev_halo, 0));
    } else {
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
  s_exchange_state(c, x, c->gX_lo, c->gX_hi);
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

// result is guaranteed to land in `target`; `tmp` is scratch of the same size
// Whole solve in one (3 sweeps) or two (5 sweeps) temporally blocked passes (pfh_cheb2_tma.cu); false: not applicable
bool cheb_solve_fused(pfh_ctx* c, double am, double ak, const std::vector<ChebCoef>& cf, const double* r, double* target,
                      double* tmp, double* tmp2) {
  const int ns = (int)cf.size();
  if (c->dim != 3 || !c->use_fused || !pfh_tma_ok(c) || (ns != 3 && ns != 5) || c->m3.nzl < 2) return false;
  if (ns == 5 && (tmp == nullptr || tmp2 == nullptr)) return false;
  const bool multi = c->cfg.nranks > 1;
  const double n = (double)c->n_owned;
  if (multi) comm_halo_n(c, r, 2, c->g2R_lo, c->g2R_hi);
  if (ns == 3) {
    ProfScope ps(c, T_CHEB_SWEEP, 16.0 * n);
    return k3_cheb2_launch(c, true, am, ak, cf[0].b, cf[1].a, cf[1].b, cf[2].a, cf[2].b, nullptr, nullptr, r, nullptr, target);
  }
  {  // r -> z2 (tmp), z3 (tmp2)
    ProfScope ps(c, T_CHEB_SWEEP, 24.0 * n);
    if (!k3_cheb2_launch(c, true, am, ak, cf[0].b, cf[1].a, cf[1].b, cf[2].a, cf[2].b, nullptr, nullptr, r, tmp, tmp2)) return false;
  }
  if (multi) {
    comm_halo_n(c, tmp2, 2, c->g2S_lo, c->g2S_hi);
    comm_halo_n(c, tmp, 2, c->g2P_lo, c->g2P_hi);
  }
  ProfScope ps(c, T_CHEB_SWEEP, 32.0 * n);  // z3, z2, r -> z5
  return k3_cheb2_launch(c, false, am, ak, 0.0, cf[3].a, cf[3].b, cf[4].a, cf[4].b, tmp2, tmp, r, nullptr, target);
}

void cheb_solve(pfh_ctx* c, double am, double ak, double lmin, double lmax, int nsweeps, const double* r, double* target,
                double* tmp, bool consumer_needs_ghosts = false, double* tmp2 = nullptr) {
  const std::vector<ChebCoef> cf = cheb_coeffs(lmin, lmax, nsweeps);
  if (cheb_solve_fused(c, am, ak, cf, r, target, tmp, tmp2)) return;
  double* cur = ((nsweeps - 1) % 2 == 0) ? target : tmp;
  double* prev = (cur == target) ? tmp : target;
  const bool ov = c->cfg.nranks > 1 && c->overlap && c->comm_stream && c->dim == 3 && c->m3.nzl >= 4;
  // 3-D fast path: the first two iterates come out of one pass over r (k_cheb_first2), the loop then starts at k = 2
  const bool first2 = nsweeps >= 2 && c->use_first2 && c->dim == 3 && pfh_tma_ok(c) && c->scalar_variant == 0;
  if (!ov) {
    int k0 = 1;
    if (first2 && k_cheb_first2(c, am, ak, cf[0].b, cf[1].a, cf[1].b, scalar_ref(c, r), cur, prev)) {
      std::swap(cur, prev);
      k0 = 2;
    } else {
      k_cheb_first(c, am, ak, cf[0].b, r, cur);
    }
    for (int k = k0; k < nsweeps; ++k) {
      k_cheb_sweep(c, am, ak, cf[k].a, cf[k].b, r, scalar_ref(c, cur), prev, k == 1);
      std::swap(cur, prev);
    }
  } else if (first2) {
    // as below, but the first launch pair reads r's ghost planes: they travel under the interior launch
    if (c->gS_prefetched != r) {
      PFH_CUDA(cudaEventRecord(c->ev_bnd, c->stream));
      PFH_CUDA(cudaStreamWaitEvent(c->comm_stream, c->ev_bnd, 0));
      comm_halo(c, r, 0, 1, c->gS_lo, c->gS_hi, c->comm_stream);
      PFH_CUDA(cudaEventRecord(c->ev_halo, c->comm_stream));
    }
    c->gS_prefetched = nullptr;
    const FieldRef rref{r, c->gS_lo, c->gS_hi};
    c->zpart = 2;
    PFH_CHECK(k_cheb_first2(c, am, ak, cf[0].b, cf[1].a, cf[1].b, rref, cur, prev), "internal: fused first sweeps unavailable");
    PFH_CUDA(cudaStreamWaitEvent(c->stream, c->ev_halo, 0));
    c->zpart = 1;
    k_cheb_first2(c, am, ak, cf[0].b, cf[1].a, cf[1].b, rref, cur, prev);
    PFH_CUDA(cudaEventRecord(c->ev_bnd, c->stream));
    std::swap(cur, prev);
    if (nsweeps == 2 && consumer_needs_ghosts) {
      PFH_CUDA(cudaStreamWaitEvent(c->comm_stream, c->ev_bnd, 0));
      comm_halo(c, cur, 0, 1, c->gS_lo, c->gS_hi, c->comm_stream);
      PFH_CUDA(cudaEventRecord(c->ev_halo, c->comm_stream));
      c->gS_prefetched = cur;
    }
    for (int k = 2; k < nsweeps; ++k) {
      PFH_CUDA(cudaStreamWaitEvent(c->comm_stream, c->ev_bnd, 0));
      comm_halo(c, cur, 0, 1, c->gS_lo, c->gS_hi, c->comm_stream);
      PFH_CUDA(cudaEventRecord(c->ev_halo, c->comm_stream));
      PFH_CUDA(cudaStreamWaitEvent(c->stream, c->ev_halo, 0));
      const FieldRef ref{cur, c->gS_lo, c->gS_hi};
      c->zpart = 1;
      k_cheb_sweep(c, am, ak, cf[k].a, cf[k].b, r, ref, prev, false);
      PFH_CUDA(cudaEventRecord(c->ev_bnd, c->stream));
      if (k == nsweeps - 1 && consumer_needs_ghosts) {
        PFH_CUDA(cudaStreamWaitEvent(c->comm_stream, c->ev_bnd, 0));
        comm_halo(c, prev, 0, 1, c->gS_lo, c->gS_hi, c->comm_stream);
        PFH_CUDA(cudaEventRecord(c->ev_halo, c->comm_stream));
        c->gS_prefetched = prev;
      }
      c->zpart = 2;
      k_cheb_sweep(c, am, ak, cf[k].a, cf[k].b, r, ref, prev, false);
      std::swap(cur, prev);
    }
    c->zpart = 0;
  } else {
    // Halo/compute overlap: every producer writes its two slab-boundary planes first and signals ev_bnd; the ghost
    // exchange of those planes runs on comm_stream while the interior planes are computed on the ctx stream; the
    // next sweep's boundary launch waits for ev_halo.  Only the interior launch never touches ghost planes.
    c->zpart = 1;
    k_cheb_first(c, am, ak, cf[0].b, r, cur);
    PFH_CUDA(cudaEventRecord(c->ev_bnd, c->stream));
    c->zpart = 2;
    k_cheb_first(c, am, ak, cf[0].b, r, cur);
    for (int k = 1; k < nsweeps; ++k) {
      PFH_CUDA(cudaStreamWaitEvent(c->comm_stream, c->ev_bnd, 0));
      comm_halo(c, cur, 0, 1, c->gS_lo, c->gS_hi, c->comm_stream);
      PFH_CUDA(cudaEventRecord(c->ev_halo, c->comm_stream));
      PFH_CUDA(cudaStreamWaitEvent(c->stream, c->ev_halo, 0));
      const FieldRef ref{cur, c->gS_lo, c->gS_hi};
      c->zpart = 1;
      k_cheb_sweep(c, am, ak, cf[k].a, cf[k].b, r, ref, prev, k == 1);
      PFH_CUDA(cudaEventRecord(c->ev_bnd, c->stream));
      if (k == nsweeps - 1 && consumer_needs_ghosts) {
        // the result's slab-boundary planes are final: ship its ghost planes now, under the last interior launch
        PFH_CUDA(cudaStreamWaitEvent(c->comm_stream, c->ev_bnd, 0));
        comm_halo(c, prev, 0, 1, c->gS_lo, c->gS_hi, c->comm_stream);
        PFH_CUDA(cudaEventRecord(c->ev_halo, c->comm_stream));
        c->gS_prefetched = prev;
      }
      c->zpart = 2;
      k_cheb_sweep(c, am, ak, cf[k].a, cf[k].b, r, ref, prev, k == 1);
      std::swap(cur, prev);
    }
    c->zpart = 0;
  }
  PFH_CHECK(cur == target, "internal: chebyshev ping-pong parity");
}

struct HostTables {
  std::vector<double> mT, kT;
};
HostTables host_tables(const pfh_ctx* c) {
  HostTables t;
  if (c->dim == 3) {
    Mesh3 m = c->m3;
    pfh_build_tables3(m, t.mT, t.kT);
  } else {
    Mesh2 m = c->m2;
    pfh_build_tables2(m, t.mT, t.kT);
  }
  return t;
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

}  // namespace

void s_precond_setup(pfh_ctx* c, const pfh_solver_opts* o) {
  c->pc_opts = *o;
  const PcLayout L = pc_layout(c);
  c->lmax = 0.5 * (c->dim + 2);
  c->lmin_m = 0.5;
  if (L.has_ch) {
    c->pc_alpha = std::sqrt(c->dt * L.Mob * L.kappa);
    const HostTables t = host_tables(c);
    c->lmin_a = std::fmin(0.5, lmin_estimate(c, 1.0, c->pc_alpha, t.mT, t.kT));
  }
  // lumped reaction term d g_a / d u_a of the parabolic fields' diagonal blocks at the linearisation point
  // (Allen-Cahn eta_i of bench2, U and phi of bench3); clipped at -1/2 of the block's mass coefficient
  if (c->jac_point != nullptr)
    for (int k = 0; k < L.n_spd; ++k)
      if (c->SIG[k] != nullptr) k_reaction_diag(c, c->jac_point, L.spd_field[k], -0.5 * L.spd_am[k], c->SIG[k]);
  c->pc_ready = true;
}

void s_precond_apply(pfh_ctx* c, const double* r, double* z) {
  PFH_CHECK(c->pc_ready, "precond_apply before precond_setup");
  const long long n = c->n_owned;
  const pfh_solver_opts& o = c->pc_opts;
  if (o.preconditioner == 2) {
    v_copy(c, z, r, c->nf * n);
    return;
  }
  const PcLayout L = pc_layout(c);
  if (o.preconditioner == 1 || L.jacobi_only) {
    k_jacobi(c, c->jac_point, r, z);
    return;
  }
  const int nM = o.pc_sweeps_mass > 0 ? o.pc_sweeps_mass : 3;
  const int nA = o.pc_sweeps_shifted > 0 ? o.pc_sweeps_shifted : 5;
  if (L.has_ch) {
    const double dt = c->dt, alpha = c->pc_alpha;
    const double *r1 = r, *r2 = r + n;
    double *S0 = c->S[0], *S1 = c->S[1], *S2 = c->S[2], *S3 = c->S[3], *S4 = c->S[4];
    // t = M^-1 r2   (block upper-triangular variant: t IS dmu, so it is solved straight into z_mu)
    double* T = o.pc_variant == 1 ? z + n : S0;
    cheb_solve(c, 1.0, 0.0, c->lmin_m, c->lmax, nM, r2, T, S1, true);
    // q = dt r1 - dt Mob K t
    k_lin_apply(c, 0.0, -dt * L.Mob, scalar_ref(c, T), dt, r1, S2);
    // dc = A^-1 M A^-1 q,  A = M + alpha K
    cheb_solve(c, 1.0, alpha, c->lmin_a, c->lmax, nA, S2, S3, S4, true, S1);
    k_lin_apply(c, 1.0, 0.0, scalar_ref(c, S3), 0.0, nullptr, S2);
    cheb_solve(c, 1.0, alpha, c->lmin_a, c->lmax, nA, S2, z, S4, o.pc_variant != 1, S1);
    if (o.pc_variant != 1) {
      // dmu = t + M^-1 (kappa K + M_f'') dc = t - M^-1 [J (dc, 0, ..)]_mu
      if (k_ch_crow(c, c->jac_point, scalar_ref(c, z), S2)) {
        // fused (mu, c) block apply: reads c and dc only, writes the mu row only
        cheb_solve(c, 1.0, 0.0, c->lmin_m, c->lmax, nM, S2, S3, S1);
        v_axpby(c, 1.0, S0, -1.0, S3, z + n, n);
      } else {
        double *PX = c->W2, *PY = c->R;  // free during a linear solve (see gmres below)
        v_zero(c, PX, c->nf * n);
        v_copy(c, PX, z, n);
        s_jac_apply(c, PX, PY);
        cheb_solve(c, 1.0, 0.0, c->lmin_m, c->lmax, nM, PY + n, S2, S1);
        v_axpby(c, 1.0, S0, -1.0, S2, z + n, n);
      }
    }
  }
  for (int k = 0; k < L.n_spd; ++k) {
    const int f = L.spd_field[k];
    const HostTables t = host_tables(c);
    // with the clipped reaction shift the constant mode's Rayleigh quotient is at least half the unshifted one
    const double lmin = 0.5 * std::fmin(0.5, lmin_estimate(c, L.spd_am[k], L.spd_ak[k], t.mT, t.kT));
    c->scalar_sig = (c->SIG[k] != nullptr && c->jac_point != nullptr) ? c->SIG[k] : nullptr;
    cheb_solve(c, L.spd_am[k], L.spd_ak[k], lmin, c->lmax, nA, r + f * n, z + f * n, c->S[1]);
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
    const double pi = 3.14159265358979323846;
    const double lmin = 0.5 * (pi / c->m2.nx) * (pi / c->m2.nx) / 8.0;
    cheb_solve(c, 0.0, 1.0, lmin, 2.0, 240, rhs, z + 2 * n, tmp);
    c->scalar_bc = 0;
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
    const KspResult ks = gmres(c, c->B, c->DX, o, rep);
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
