// Steady quasi-linear diffusion on the right-diagonal 2-D mesh: the reference's nonlinear-solver script dolfin/nl.py.
//
//   mesh   UnitSquareMesh(Nx, Ny)                   (nl.py:20; DOLFIN's default 'right' split: every grid square is cut by
//                                                    the diagonal v0-v3 into (v0, v1, v3) and (v0, v2, v3); vertices x-fastest)
//   form   F(u; v) = int src v - int kc |u| grad u . grad v - int reac u v        (nl.py:37, sqrt(u*u) = |u|)
//   J x    = -int kc sign(u) x grad u . grad v - int kc |u| grad x . grad v - int reac x v   (derivative(F, u, du), nl.py:38)
//   bc     u = g on the vertices of x = 0 and x = Lx (nl.py:31-35): rows F_i = u_i - g, (J x)_i = x_i
//
// Quadrature degree 3 (nl.py:12) = the 6-point Strang-Fix rule on every triangle; grad u is constant per triangle, so the
// diffusion term of a triangle is (grad u . grad phi_i) * sum_q w_q kc |u_q|; only a triangle on which u_h changes sign needs
// the six points for it, and the polynomial terms have closed forms (tri_op).  Nothing is assembled: one call computes the
// row of one vertex from the (up to) six triangles around it.  The functions are __host__ __device__: the CUDA kernels of
// pfh_nl.cu are thread -> vertex mappings around them, and tests/emul compiles them for the CPU against the oracle
// (oracle/nl.py).
#pragma once
#include "pfh_elem.cuh"

#ifdef __CUDA_ARCH__
#define PFH_NL_LD(p) __ldg(p)
#else
#define PFH_NL_LD(p) (*(p))
#endif

// MeshR2, NlParams: pfh_internal.h

namespace nl {

// MODE 0: residual F(u); MODE 1: J(u) x; MODE 2: A x with A = K_{kc|u|} + reac M (= -J without the convective term, SPD)
// One triangle with local vertex values u[3] (and x[3]), constant basis gradients g[3][2]; adds to out[3] and, in MODE 2,
// the diagonal entries to dg[3].  w = |det J| / 12 = hx hy / 12 is the weight of one quadrature point.
template <int MODE>
PFH_HD void tri_op(const NlParams& p, double w, const double (&u)[3], const double (&x)[3], const double (&g)[3][2],
                   double (&out)[3], double (&dg)[3]) {
  // sum_q |u_q| and sum_q sign(u_q) x_q over the six points.  The barycentric coordinates of the points are the six
  // permutations of (A, B, C), A + B + C = 1, so sum_q u_q = 2 (u0 + u1 + u2): where u_h keeps one sign on the triangle
  // (everywhere in nl.py's solution, which stays above its boundary value 0.9) the two sums need no quadrature loop.
  const double Su = u[0] + u[1] + u[2];
  const double Sx = MODE != 0 ? x[0] + x[1] + x[2] : 0.0;
  double kbar = 0.0, sbar = 0.0;
  const bool pos = u[0] > 0.0 && u[1] > 0.0 && u[2] > 0.0, neg = u[0] < 0.0 && u[1] < 0.0 && u[2] < 0.0;
  if (pos || neg) {
    const double sg = pos ? 2.0 : -2.0;
    kbar = sg * Su;
    if (MODE == 1) sbar = sg * Sx;
  } else {
#pragma unroll
    for (int q = 0; q < 6; ++q) {
      const double l0 = elem::sf(q, 0), l1 = elem::sf(q, 1), l2 = elem::sf(q, 2);
      const double uq = l0 * u[0] + l1 * u[1] + l2 * u[2];
      kbar += fabs(uq);
      if (MODE == 1) sbar += (uq > 0.0 ? 1.0 : (uq < 0.0 ? -1.0 : 0.0)) * (l0 * x[0] + l1 * x[1] + l2 * x[2]);
    }
  }
  kbar *= w * p.kc;
  sbar *= w * p.kc;
  const double gu[2] = {u[0] * g[0][0] + u[1] * g[1][0] + u[2] * g[2][0], u[0] * g[0][1] + u[1] * g[1][1] + u[2] * g[2][1]};
  double gx[2] = {0.0, 0.0};
  if (MODE != 0) {
    gx[0] = x[0] * g[0][0] + x[1] * g[1][0] + x[2] * g[2][0];
    gx[1] = x[0] * g[0][1] + x[1] * g[1][1] + x[2] * g[2][1];
  }
  // the source / reaction terms are polynomials of degree <= 2, which the rule integrates exactly: closed forms
  // int phi_l = |T|/3 = 2w, int v phi_l = |T|/12 (v0 + v1 + v2 + v_l) = (w/2)(S + v_l), int phi_l^2 = |T|/6 = w
  const double hm = 0.5 * w * p.reac;
#pragma unroll
  for (int l = 0; l < 3; ++l) {
    const double gul = gu[0] * g[l][0] + gu[1] * g[l][1];
    const double gxl = gx[0] * g[l][0] + gx[1] * g[l][1];
    if (MODE == 0) out[l] += (2.0 * w * p.src - hm * (Su + u[l])) - kbar * gul;
    if (MODE == 1) out[l] += -hm * (Sx + x[l]) - sbar * gul - kbar * gxl;
    if (MODE == 2) {
      out[l] += hm * (Sx + x[l]) + kbar * gxl;
      dg[l] += p.reac * w + kbar * (g[l][0] * g[l][0] + g[l][1] * g[l][1]);
    }
  }
}

// One grid square: corner values in local order v0 (0,0), v1 (1,0), v2 (0,1), v3 (1,1); adds the contributions of its two
// triangles to out[4] (and the diagonal entries to dg[4]).
template <int MODE>
PFH_HD void square_op(const MeshR2& m, const NlParams& p, const double (&u)[4], const double (&x)[4], double (&out)[4],
                      double (&dg)[4]) {
  const double ix = 1.0 / m.hx, iy = 1.0 / m.hy, w = m.hx * m.hy / 12.0;
  {  // (v0, v1, v3): lambda0 = 1 - x/hx, lambda1 = x/hx - y/hy, lambda3 = y/hy
    const double g[3][2] = {{-ix, 0.0}, {ix, -iy}, {0.0, iy}};
    const double ul[3] = {u[0], u[1], u[3]}, xl[3] = {x[0], x[1], x[3]};
    double o[3] = {0.0, 0.0, 0.0}, d[3] = {0.0, 0.0, 0.0};
    tri_op<MODE>(p, w, ul, xl, g, o, d);
    out[0] += o[0]; out[1] += o[1]; out[3] += o[2];
    dg[0] += d[0]; dg[1] += d[1]; dg[3] += d[2];
  }
  {  // (v0, v2, v3): lambda0 = 1 - y/hy, lambda2 = y/hy - x/hx, lambda3 = x/hx
    const double g[3][2] = {{0.0, -iy}, {-ix, iy}, {ix, 0.0}};
    const double ul[3] = {u[0], u[2], u[3]}, xl[3] = {x[0], x[2], x[3]};
    double o[3] = {0.0, 0.0, 0.0}, d[3] = {0.0, 0.0, 0.0};
    tri_op<MODE>(p, w, ul, xl, g, o, d);
    out[0] += o[0]; out[2] += o[1]; out[3] += o[2];
    dg[0] += d[0]; dg[2] += d[1]; dg[3] += d[2];
  }
}

PFH_HD bool is_dirichlet(const MeshR2& m, int ix) { return ix == 0 || ix == m.nx; }

// Row of vertex (ix, iy): gathers its own corner entry from the squares NE, NW, SE, SW (local index 0, 1, 2, 3 in them).
// val = F_i / (J x)_i / (A x)_i, diag = A_ii (MODE 2).  U = linearisation point / iterate, X = operand (MODE 1, 2).
template <int MODE>
PFH_HD void vertex_op(const MeshR2& m, const NlParams& p, const double* U, const double* X, int ix, int iy, double& val,
                      double& diag) {
  const long long i = (long long)iy * m.nxv + ix;
  if (is_dirichlet(m, ix)) {
    val = MODE == 0 ? PFH_NL_LD(U + i) - p.bc : PFH_NL_LD(X + i);
    diag = 1.0;
    return;
  }
  double v = 0.0, d = 0.0;
#pragma unroll
  for (int s = 0; s < 4; ++s) {
    const int ox = ix - (s & 1), oy = iy - (s >> 1);
    if (ox < 0 || ox > m.nx - 1 || oy < 0 || oy > m.ny - 1) continue;
    const long long v0 = (long long)oy * m.nxv + ox;
    const double u[4] = {PFH_NL_LD(U + v0), PFH_NL_LD(U + v0 + 1), PFH_NL_LD(U + v0 + m.nxv), PFH_NL_LD(U + v0 + m.nxv + 1)};
    double x[4] = {0.0, 0.0, 0.0, 0.0};
    if (MODE != 0) {
      x[0] = PFH_NL_LD(X + v0); x[1] = PFH_NL_LD(X + v0 + 1); x[2] = PFH_NL_LD(X + v0 + m.nxv); x[3] = PFH_NL_LD(X + v0 + m.nxv + 1);
    }
    double o[4] = {0.0, 0.0, 0.0, 0.0}, dd[4] = {0.0, 0.0, 0.0, 0.0};
    square_op<MODE>(m, p, u, x, o, dd);
    v += o[s];
    d += dd[s];
  }
  val = v;
  diag = d;
}

}  // namespace nl
