// TEST INFRASTRUCTURE: compiles the product's __host__ __device__ per-vertex operator code
// (pfh_vertex.cuh / pfh_elem.cuh / pfh_problems.cuh / pfh_tables.cpp) for the CPU so the
// index arithmetic, stencil tables, quadrature and hand-written f', f'' can be checked against
// the oracle without a GPU.  Never loaded by the product.
#include <cmath>
#include <vector>

#include "../../pfhubbenchmarks_b200/csrc/pfh_vertex.cuh"
#include "../../pfhubbenchmarks_b200/csrc/pfh_sample.cuh"
#include "../../pfhubbenchmarks_b200/csrc/pfh_nl.cuh"
#include "../../pfhubbenchmarks_b200/csrc/pfh_op2_tiled.cuh"
#include "../../pfhubbenchmarks_b200/csrc/pfh_mg2.cuh"

void pfh_set_error(const std::string&) {}

namespace {
struct Setup {
  Mesh2 m2;
  Mesh3 m3;
  MeshR2 mr;
  std::vector<double> mT, kT;
  int dim;
};

void setup(Setup& S, int mesh_kind, const int* nc, const double* len, int z0, int nzl) {
  if (mesh_kind == PFH_MESH_KUHN3D) {
    S.dim = 3;
    Mesh3& m = S.m3;
    m.nxv = nc[0] + 1; m.nyv = nc[1] + 1; m.nzv = nc[2] + 1;
    m.z0 = z0; m.nzl = nzl < 0 ? m.nzv : nzl;
    m.nxy = (long long)m.nxv * m.nyv; m.n_owned = m.nxy * m.nzl;
    m.hx = len[0] / nc[0]; m.hy = len[1] / nc[1]; m.hz = len[2] / nc[2];
    m.detJ = m.hx * m.hy * m.hz;
    pfh_build_tables3(m, S.mT, S.kT);
    m.mT = S.mT.data(); m.kT = S.kT.data();
  } else if (mesh_kind == PFH_MESH_RIGHT2D) {
    S.dim = 2;
    memset(&S.m2, 0, sizeof S.m2);
    MeshR2& m = S.mr;
    m.nx = nc[0]; m.ny = nc[1]; m.nxv = nc[0] + 1; m.nyv = nc[1] + 1;
    m.n = (long long)m.nxv * m.nyv;
    m.hx = len[0] / nc[0]; m.hy = len[1] / nc[1];
  } else {
    S.dim = 2;
    Mesh2& m = S.m2;
    m.nx = nc[0]; m.ny = nc[1]; m.nxv = nc[0] + 1; m.nyv = nc[1] + 1;
    m.ncorner = (long long)m.nxv * m.nyv; m.ncentre = (long long)m.nx * m.ny; m.n_owned = m.ncorner + m.ncentre;
    m.hx = len[0] / nc[0]; m.hy = len[1] / nc[1]; m.detJ = 0.5 * m.hx * m.hy;
    pfh_build_tables2(m, S.mT, S.kT);
    m.mT = S.mT.data(); m.kT = S.kT.data();
  }
}

// the square-centric 2-D kernel k2_op_tiled (csrc/pfh_stencil.cu) block by block: all threads of a CTA run a phase before
// the next one starts, as between the kernel's barriers
template <class P, bool JAC>
int run_tiled(const Mesh2& m, const P& prob, const LinCoefs& L, const double* OP, const double* EX, double* out) {
  if constexpr (op2t::fits<P, JAC>()) {
    using namespace op2t;
    std::vector<Tile<P, JAC>> tile(1);
    for (int by = 0; by < (m.nyv + WY - 1) / WY; ++by)
      for (int bx = 0; bx < (m.nxv + WX - 1) / WX; ++bx) {
        const int x0 = bx * WX - 1, y0 = by * WY - 1;
        memset(&tile[0], 0xff, sizeof tile[0]);   // NaN fill: a read of an entry no phase wrote would show
        for (int t = 0; t < SX * SY; ++t) stage<P, JAC>(tile[0], m, L, OP, EX, x0, y0, t, SX * SY);
        for (int ty = 0; ty < SY; ++ty)
          for (int tx = 0; tx < SX; ++tx) square_phase<P, JAC>(tile[0], prob, m, L, x0, y0, tx, ty, out);
        for (int ty = 0; ty < SY; ++ty)
          for (int tx = 0; tx < SX; ++tx) vertex_phase<P, JAC>(tile[0], prob, m, L, x0, y0, tx, ty, out);
      }
    return 0;
  } else {
    return -3;   // tile does not fit the static shared memory: the product uses the gather kernel for this problem
  }
}

template <class P>
int run(const Setup& S, const ProblemParams& pp, double dt, int mode, const double* U, const double* X, const double* U0,
        const double* glo, const double* ghi, double* out, const double* icp) {
  P prob(pp);
  LinCoefs L;
  memset(&L, 0, sizeof L);
  prob.lin(dt, L);
  if (S.dim == 3) {
    const Mesh3& m = S.m3;
    const long long n = m.n_owned;
    FieldRef u[P::NF], x[P::NF], u0[P::NF];
    // ghost layout (slab tests): glo/ghi hold [3 vectors (U, X, U0)][NF][nxy]
    for (int f = 0; f < P::NF; ++f) {
      u[f] = {U + f * n, glo ? glo + (0 * P::NF + f) * m.nxy : nullptr, ghi ? ghi + (0 * P::NF + f) * m.nxy : nullptr};
      x[f] = {X ? X + f * n : nullptr, glo ? glo + (1 * P::NF + f) * m.nxy : nullptr, ghi ? ghi + (1 * P::NF + f) * m.nxy : nullptr};
      u0[f] = {U0 ? U0 + f * n : nullptr, glo ? glo + (2 * P::NF + f) * m.nxy : nullptr, ghi ? ghi + (2 * P::NF + f) * m.nxy : nullptr};
    }
    double E = 0, Mm = 0;
#pragma omp parallel for reduction(+ : E, Mm) collapse(2)
    for (int zl = 0; zl < m.nzl; ++zl)
      for (int iy = 0; iy < m.nyv; ++iy)
        for (int ix = 0; ix < m.nxv; ++ix) {
          const long long i = (long long)zl * m.nxy + (long long)iy * m.nxv + ix;
          if (mode == 0 || mode == 1) {
            double o[P::NF];
            if (mode == 0) vtx::vertex_op3<P, false>(prob, m, L, u, x, u0, ix, iy, zl, o);
            else vtx::vertex_op3<P, true>(prob, m, L, u, x, u0, ix, iy, zl, o);
            for (int a = 0; a < P::NF; ++a) out[a * n + i] = o[a];
          } else if (mode == 2) {
            double e, mm;
            vtx::vertex_energy3<P>(prob, m, u, ix, iy, zl, e, mm);
            E += e; Mm += mm;
          } else if (mode == 3) {  // scalar op: out = (am M + ak K) U[0], am = icp[0], ak = icp[1]
            double Av, d, s;
            vtx::scalar_op3(m, icp[0], icp[1], u[0], ix, iy, zl, Av, d, s);
            out[i] = Av; out[n + i] = d;
          } else if (mode == 4) {
            double o[P::NF];
            prob.ic(icp, ix * m.hx, iy * m.hy, o);
            for (int a = 0; a < P::NF; ++a) out[a * n + i] = o[a];
          }
        }
    if (mode == 2) { out[0] = E; out[1] = Mm; }
    if constexpr (P::QUARTIC) {
      if (mode == 6) {  // nonlinear residual load only, cube-centric scatter with the low-flop routine
        double a3, a2, a1, a0;
        prob.res_poly(a3, a2, a1, a0);
        const elem::QuarticRes qr(a3, a2, a1, a0, m.detJ);
        for (long long i = 0; i < n; ++i) out[i] = 0.0;
        const double* cu = U + P::in(0) * n;
        for (int cz = 0; cz < m.nzv - 1; ++cz)
          for (int cy = 0; cy < m.nyv - 1; ++cy)
            for (int cxi = 0; cxi < m.nxv - 1; ++cxi) {
              double c8[8], o8[8] = {0};
              long long idx[8];
              for (int l = 0; l < 8; ++l) {
                idx[l] = (long long)(cz + (l >> 2)) * m.nxy + (long long)(cy + ((l >> 1) & 1)) * m.nxv + cxi + (l & 1);
                c8[l] = cu[idx[l]];
              }
              elem::cube_res_quartic(qr, c8, o8);
              for (int l = 0; l < 8; ++l) out[idx[l]] += o8[l];
            }
      }
      if (mode == 5 || mode == 7) {  // nonlinear Jacobian block only, cube-centric scatter with the low-flop routine (7: its fp32 variant)
        double a2, a1, a0;
        prob.jac_poly(a2, a1, a0);
        const elem::QuarticJac qj(a2, a1, a0, m.detJ);
        for (long long i = 0; i < n; ++i) out[i] = 0.0;
        const double* cu = U + P::in(0) * n;
        const double* cx = X + P::in(0) * n;
        for (int cz = 0; cz < m.nzv - 1; ++cz)
          for (int cy = 0; cy < m.nyv - 1; ++cy)
            for (int cxi = 0; cxi < m.nxv - 1; ++cxi) {
              double c8[8], x8[8], o8[8] = {0};
              long long idx[8];
              for (int l = 0; l < 8; ++l) {
                idx[l] = (long long)(cz + (l >> 2)) * m.nxy + (long long)(cy + ((l >> 1) & 1)) * m.nxv + cxi + (l & 1);
                c8[l] = cu[idx[l]]; x8[l] = cx[idx[l]];
              }
              if (mode == 7) elem::cube_jac_quartic_f32(qj, c8, x8, o8);
              else elem::cube_jac_quartic(qj, c8, x8, o8);
              for (int l = 0; l < 8; ++l) out[idx[l]] += o8[l];
            }
      }
    }
  } else {
    const Mesh2& m = S.m2;
    const long long n = m.n_owned;
    if (mode == 8 || mode == 9) {
      for (long long i = 0; i < P::NF * n; ++i) out[i] = NAN;   // every entry must be written exactly by its owner
      return mode == 8 ? run_tiled<P, false>(m, prob, L, U, U0, out) : run_tiled<P, true>(m, prob, L, X, U, out);
    }
    double E = 0, Mm = 0;
#pragma omp parallel for reduction(+ : E, Mm)
    for (long long i = 0; i < n; ++i) {
      if (mode == 0 || mode == 1) {
        double o[P::NF];
        if (mode == 0) vtx::vertex_op2<P, false>(prob, m, L, U, X, U0, i, o);
        else vtx::vertex_op2<P, true>(prob, m, L, U, X, U0, i, o);
        for (int a = 0; a < P::NF; ++a) out[a * n + i] = o[a];
      } else if (mode == 2) {
        double e, mm;
        vtx::vertex_energy2<P>(prob, m, U, i, e, mm);
        E += e; Mm += mm;
      } else if (mode == 3) {
        double Av, d, s;
        vtx::scalar_op2(m, icp[0], icp[1], U, i, Av, d, s);
        out[i] = Av; out[n + i] = d;
      } else if (mode == 4) {
        double x, y;
        if (i < m.ncorner) { x = (i % m.nxv) * m.hx; y = (i / m.nxv) * m.hy; }
        else { const long long ci = i - m.ncorner; x = (ci % m.nx + 0.5) * m.hx; y = (ci / m.nx + 0.5) * m.hy; }
        double o[P::NF];
        prob.ic(icp, x, y, o);
        for (int a = 0; a < P::NF; ++a) out[a * n + i] = o[a];
      }
    }
    if (mode == 2) { out[0] = E; out[1] = Mm; }
  }
  return 0;
}
}  // namespace

// point evaluation (csrc/pfh_sample.cuh) on the CPU: out[i] = value, +inf outside; 3-D slabs: NaN where not owned
extern "C" int emul_sample(int mesh_kind, const int* ncell, const double* length, int z0, int nzl, const double* u,
                           const double* glo, const double* ghi, const double* pts, long long npts, double* out) {
  try {
    Setup S;
    setup(S, mesh_kind, ncell, length, z0, nzl);
    for (long long i = 0; i < npts; ++i) {
      double v = 0.0;
      if (mesh_kind == 3) {
        bool inside, owned;
        smp::eval3(S.m3, FieldRef{u, glo, ghi}, pts[3 * i], pts[3 * i + 1], pts[3 * i + 2], inside, owned, v);
        out[i] = !inside ? INFINITY : (owned ? v : NAN);
      } else {
        out[i] = smp::eval2(S.m2, u, pts[2 * i], pts[2 * i + 1], v) ? v : INFINITY;
      }
    }
    return 0;
  } catch (const std::exception& e) {
    fprintf(stderr, "emul: %s\n", e.what());
    return -1;
  }
}

// dolfin/nl.py on the right-diagonal mesh (csrc/pfh_nl.cuh): mode 0 residual, 1 J(U) X, 2 A X and the diagonal of A
// (out[0..n) values, out[n..2n) diagonal).  square != 0: square-centric scatter (the order of work of the tiled CUDA kernel)
// instead of the vertex-centric gather.
extern "C" int emul_nl(const int* ncell, const double* length, const double* params, int mode, int square, const double* U,
                       const double* X, double* out) {
  Setup S;
  setup(S, PFH_MESH_RIGHT2D, ncell, length, 0, -1);
  const MeshR2& m = S.mr;
  const NlParams p{params[0], params[1], params[2], params[3]};
  for (long long i = 0; i < 2 * m.n; ++i) out[i] = 0.0;
  if (square) {
    for (int oy = 0; oy < m.ny; ++oy)
      for (int ox = 0; ox < m.nx; ++ox) {
        const long long v0 = (long long)oy * m.nxv + ox, idx[4] = {v0, v0 + 1, v0 + m.nxv, v0 + m.nxv + 1};
        double u[4], x[4] = {0, 0, 0, 0}, o[4] = {0, 0, 0, 0}, d[4] = {0, 0, 0, 0};
        for (int l = 0; l < 4; ++l) {
          u[l] = U[idx[l]];
          if (mode != 0) x[l] = X[idx[l]];
        }
        if (mode == 0) nl::square_op<0>(m, p, u, x, o, d);
        else if (mode == 1) nl::square_op<1>(m, p, u, x, o, d);
        else nl::square_op<2>(m, p, u, x, o, d);
        for (int l = 0; l < 4; ++l) {
          out[idx[l]] += o[l];
          out[m.n + idx[l]] += d[l];
        }
      }
    for (int iy = 0; iy < m.nyv; ++iy)
      for (int ix = 0; ix < m.nxv; ++ix)
        if (nl::is_dirichlet(m, ix)) {
          const long long i = (long long)iy * m.nxv + ix;
          out[i] = mode == 0 ? U[i] - p.bc : X[i];
          out[m.n + i] = 1.0;
        }
    return 0;
  }
  for (int iy = 0; iy < m.nyv; ++iy)
    for (int ix = 0; ix < m.nxv; ++ix) {
      const long long i = (long long)iy * m.nxv + ix;
      double v, d;
      if (mode == 0) nl::vertex_op<0>(m, p, U, X, ix, iy, v, d);
      else if (mode == 1) nl::vertex_op<1>(m, p, U, X, ix, iy, v, d);
      else nl::vertex_op<2>(m, p, U, X, ix, iy, v, d);
      out[i] = v;
      out[m.n + i] = d;
    }
  return 0;
}

// multigrid transfers on the crossed mesh (csrc/pfh_mg2.cuh): fine mesh = ncell, coarse = ncell / 2.  mode 0: out (coarse) =
// P^T in (fine); mode 1: out (fine) = P in (coarse)
extern "C" int emul_mg_transfer(const int* ncell, const double* length, int mode, const double* in, double* out) {
  Setup S;
  setup(S, PFH_MESH_CROSSED2D, ncell, length, 0, -1);
  std::vector<Mesh2> meshes;
  std::vector<std::vector<double>> tabs;
  pfh_mg_build_levels(S.m2, meshes, tabs);
  if (meshes.empty()) return -2;
  const Mesh2 &mf = S.m2, &mc = meshes[0];
  if (mode == 0)
    for (long long i = 0; i < mc.n_owned; ++i) out[i] = mg2::restrict_node(mf, mc, in, i);
  else
    for (long long i = 0; i < mf.n_owned; ++i) out[i] = mg2::prolong_node(mf, mc, in, i);
  return (int)meshes.size();
}

extern "C" int emul_run(int mesh_kind, int problem, const int* ncell, const double* length, const double* params, int nparams,
                        double dt, int mode, int z0, int nzl, const double* U, const double* X, const double* U0,
                        const double* glo, const double* ghi, double* out, const double* icp) {
  try {
    Setup S;
    setup(S, mesh_kind, ncell, length, z0, nzl);
    ProblemParams pp;
    memset(&pp, 0, sizeof pp);
    pp.id = problem;
    for (int i = 0; i < nparams; ++i) pp.p[i] = params[i];
    switch (problem) {
      case PFH_PROBLEM_CH: return run<ProbCH>(S, pp, dt, mode, U, X, U0, glo, ghi, out, icp);
      case PFH_PROBLEM_BENCH2: return run<ProbBench2>(S, pp, dt, mode, U, X, U0, glo, ghi, out, icp);
      case PFH_PROBLEM_BENCH3: return run<ProbBench3>(S, pp, dt, mode, U, X, U0, glo, ghi, out, icp);
      case PFH_PROBLEM_BENCH6: return run<ProbBench6>(S, pp, dt, mode, U, X, U0, glo, ghi, out, icp);
#ifdef PFH_USER_PROBLEM_HEADER
      case PFH_PROBLEM_USER: return run<ProbUser>(S, pp, dt, mode, U, X, U0, glo, ghi, out, icp);   // a generated model (codegen.py)
#endif
    }
    return -2;
  } catch (const std::exception& e) {
    fprintf(stderr, "emul: %s\n", e.what());
    return -1;
  }
}
