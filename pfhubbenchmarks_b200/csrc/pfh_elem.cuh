// Element-level quadrature of the nonlinear terms, shared by all stencil kernels.
//
// The reference integrates every form with FIAT's degree-3 "default" rules because every
// driver sets parameters["form_compiler"]["quadrature_degree"] = 3 (dolfin/bench1.py:14-16):
//   triangle    : Strang-Fix 6 points, weights 1/12, barycentric = permutations of (A, B, C)
//   tetrahedron : Zienkiewicz-Taylor 5 points: centroid (weight -0.8/6) and four points with
//                 one barycentric 1/2 and three 1/6 (weight 0.45/6 each)
// Both are inexact for the quartic/.. free energies, and parity with FEniCS depends on using
// exactly these rules.  A "cell patch" is what one structured cell holds:
// the 6 Kuhn tetrahedra of a cube (DOLFIN BoxMesh) or the 4 triangles of a crossed square.
#pragma once
#include "pfh_problems.cuh"

#define PFH_DEV PFH_HD

namespace elem {

constexpr double SF_A = 0.659027622374092, SF_B = 0.231933368553031, SF_C = 0.109039009072877;
constexpr double TET_W0 = -0.8 / 6.0, TET_W1 = 0.45 / 6.0;

// local vertices of the six tetrahedra of a cube; local index = lx + 2 ly + 4 lz
PFH_HD constexpr int kuhn(int t, int a) {
  constexpr int K[6][4] = {{0, 1, 3, 7}, {0, 1, 7, 5}, {0, 5, 7, 4}, {0, 3, 2, 7}, {0, 6, 4, 7}, {0, 2, 6, 7}};
  return K[t][a];
}
// local vertices of the four triangles of a crossed square: 0..3 corners v0..v3, 4 centre
PFH_HD constexpr int tri(int t, int a) {
  constexpr int T[4][3] = {{0, 1, 4}, {0, 2, 4}, {1, 3, 4}, {2, 3, 4}};
  return T[t][a];
}
// barycentric coordinates of the 6 Strang-Fix points
PFH_HD constexpr double sf(int q, int a) {
  constexpr double L[6][3] = {{SF_C, SF_A, SF_B}, {SF_B, SF_A, SF_C}, {SF_C, SF_B, SF_A},
                              {SF_A, SF_B, SF_C}, {SF_B, SF_C, SF_A}, {SF_A, SF_C, SF_B}};
  return L[q][a];
}

// Nonlinear load of one cube: out[o][k] += int_{cube} phi_k g_o(u_h)   (JAC: g -> dg(u_h; x_h))
template <class P, bool JAC>
PFH_DEV void cube_load(const P& prob, const double (&u)[P::NIN][8], const double (&x)[P::NIN][8], double detJ,
                       double (&out)[P::NOUT][8]) {
#pragma unroll
  for (int t = 0; t < 6; ++t) {
    double s[P::NIN], sx[P::NIN];
#pragma unroll
    for (int f = 0; f < P::NIN; ++f) {
      s[f] = u[f][kuhn(t, 0)] + u[f][kuhn(t, 1)] + u[f][kuhn(t, 2)] + u[f][kuhn(t, 3)];
      sx[f] = JAC ? x[f][kuhn(t, 0)] + x[f][kuhn(t, 1)] + x[f][kuhn(t, 2)] + x[f][kuhn(t, 3)] : 0.0;
    }
    double uq[P::NIN], xq[P::NIN], g0[P::NOUT], gj[4][P::NOUT];
#pragma unroll
    for (int f = 0; f < P::NIN; ++f) {
      uq[f] = 0.25 * s[f];
      xq[f] = 0.25 * sx[f];
    }
    if (JAC) prob.dg(uq, xq, g0); else prob.g(uq, g0);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
#pragma unroll
      for (int f = 0; f < P::NIN; ++f) {
        uq[f] = s[f] * (1.0 / 6.0) + u[f][kuhn(t, j)] * (1.0 / 3.0);
        xq[f] = JAC ? sx[f] * (1.0 / 6.0) + x[f][kuhn(t, j)] * (1.0 / 3.0) : 0.0;
      }
      if (JAC) prob.dg(uq, xq, gj[j]); else prob.g(uq, gj[j]);
    }
#pragma unroll
    for (int o = 0; o < P::NOUT; ++o) {
      const double S = gj[0][o] + gj[1][o] + gj[2][o] + gj[3][o];
      const double base = detJ * (TET_W0 * 0.25 * g0[o] + (TET_W1 / 6.0) * S);
#pragma unroll
      for (int j = 0; j < 4; ++j) out[o][kuhn(t, j)] += base + detJ * (TET_W1 / 3.0) * gj[j][o];
    }
  }
}

// Low-flop form of cube_load<P, true> for a single nonlinear function whose derivative is a
// quadratic polynomial, dg(c; x) = (a2 c^2 + a1 c + a0) x (Cahn-Hilliard double well).  Same
// quadrature, with every constant folded into two pre-scaled polynomials:
//   point j of a tet : c_q = (s + 2 c_j)/6, x_q = (sx + 2 x_j)/6   ->  h_j = Q(s + 2 c_j) (sx + 2 x_j)
//   centroid         : c_q = s/4, x_q = sx/4                        ->  t0  = R(s) sx
//   corner k gets      h_k + S/2 + t0,  S = sum_j h_j      (39 fp64 ops per tet instead of ~90)
struct QuarticJac {
  double Q2, Q1, Q0, R2, R1, R0;
  QuarticJac() = default;
  PFH_HD QuarticJac(double a2, double a1, double a0, double detJ) {
    Q2 = detJ * TET_W1 * a2 / 648.0; Q1 = detJ * TET_W1 * a1 / 108.0; Q0 = detJ * TET_W1 * a0 / 18.0;
    R2 = detJ * TET_W0 * a2 / 256.0; R1 = detJ * TET_W0 * a1 / 64.0; R0 = detJ * TET_W0 * a0 / 16.0;
  }
};
// the two vertices of Kuhn tet t other than v0 and v7
PFH_HD constexpr int kuhn_mid(int t, int i) {
  constexpr int M[6][2] = {{1, 3}, {1, 5}, {5, 4}, {3, 2}, {6, 4}, {2, 6}};
  return M[t][i];
}
PFH_HD void cube_jac_quartic(const QuarticJac& k, const double (&c)[8], const double (&x)[8], double (&out)[8]) {
  const double c07 = c[0] + c[7], x07 = x[0] + x[7];
#pragma unroll
  for (int t = 0; t < 6; ++t) {
    const int p = kuhn_mid(t, 0), q = kuhn_mid(t, 1);
    const double s = c07 + c[p] + c[q], sx = x07 + x[p] + x[q];
    const int v[4] = {0, 7, p, q};
    double h[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const double tc = 2.0 * c[v[j]] + s, tx = 2.0 * x[v[j]] + sx;
      h[j] = ((k.Q2 * tc + k.Q1) * tc + k.Q0) * tx;
    }
    const double S = (h[0] + h[1]) + (h[2] + h[3]);
    const double base = 0.5 * S + ((k.R2 * s + k.R1) * s + k.R0) * sx;
#pragma unroll
    for (int j = 0; j < 4; ++j) out[v[j]] += base + h[j];
  }
}

// The same block with single-precision arithmetic (operands and accumulation in float, result widened): the Krylov
// J.v tolerates it -- iteration counts of the Newton-Krylov solve are unchanged (tests/probe/README.md) -- and it takes
// the ~230 operations per cube off the fp64 pipe.  Opt-in (pfh_set_option "jac_quadrature_fp32"); Newton's residual
// always uses the fp64 routines.
PFH_HD void cube_jac_quartic_f32(const QuarticJac& k, const double (&c)[8], const double (&x)[8], double (&out)[8]) {
  const float Q2 = (float)k.Q2, Q1 = (float)k.Q1, Q0 = (float)k.Q0, R2 = (float)k.R2, R1 = (float)k.R1, R0 = (float)k.R0;
  float cf[8], xf[8], of[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    cf[i] = (float)c[i];
    xf[i] = (float)x[i];
    of[i] = 0.0f;
  }
  const float c07 = cf[0] + cf[7], x07 = xf[0] + xf[7];
#pragma unroll
  for (int t = 0; t < 6; ++t) {
    const int p = kuhn_mid(t, 0), q = kuhn_mid(t, 1);
    const float s = c07 + cf[p] + cf[q], sx = x07 + xf[p] + xf[q];
    const int v[4] = {0, 7, p, q};
    float h[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float tc = 2.0f * cf[v[j]] + s, tx = 2.0f * xf[v[j]] + sx;
      h[j] = ((Q2 * tc + Q1) * tc + Q0) * tx;
    }
    const float S = (h[0] + h[1]) + (h[2] + h[3]);
    const float base = 0.5f * S + ((R2 * s + R1) * s + R0) * sx;
#pragma unroll
    for (int j = 0; j < 4; ++j) of[v[j]] += base + h[j];
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) out[i] += (double)of[i];
}

// Residual counterpart: load of a cubic g(c) = a3 c^3 + a2 c^2 + a1 c + a0 (= -f'(c) of the double well) over one cube,
// with the same folding: corner k gets Q(s + 2 c_k) + 1/2 sum_j Q(s + 2 c_j) + R(s).
struct QuarticRes {
  double Q3, Q2, Q1, Q0, R3, R2, R1, R0;
  QuarticRes() = default;
  PFH_HD QuarticRes(double a3, double a2, double a1, double a0, double detJ) {
    const double q = detJ * TET_W1 / 3.0, r = detJ * TET_W0 / 4.0;
    Q3 = q * a3 / 216.0; Q2 = q * a2 / 36.0; Q1 = q * a1 / 6.0; Q0 = q * a0;
    R3 = r * a3 / 64.0; R2 = r * a2 / 16.0; R1 = r * a1 / 4.0; R0 = r * a0;
  }
};
PFH_HD void cube_res_quartic(const QuarticRes& k, const double (&c)[8], double (&out)[8]) {
  const double c07 = c[0] + c[7];
#pragma unroll
  for (int t = 0; t < 6; ++t) {
    const int p = kuhn_mid(t, 0), q = kuhn_mid(t, 1);
    const double s = c07 + c[p] + c[q];
    const int v[4] = {0, 7, p, q};
    double h[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const double tc = 2.0 * c[v[j]] + s;
      h[j] = ((k.Q3 * tc + k.Q2) * tc + k.Q1) * tc + k.Q0;
    }
    const double S = (h[0] + h[1]) + (h[2] + h[3]);
    const double base = 0.5 * S + (((k.R3 * s + k.R2) * s + k.R1) * s + k.R0);
#pragma unroll
    for (int j = 0; j < 4; ++j) out[v[j]] += base + h[j];
  }
}

// Nonlinear load of one crossed square: out[o][k] += int_{square} phi_k g_o(u_h), k = 0..4
template <class P, bool JAC>
PFH_DEV void square_load(const P& prob, const double (&u)[P::NIN][5], const double (&x)[P::NIN][5], double detJ,
                         double (&out)[P::NOUT][5]) {
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    double acc[P::NOUT][3];
#pragma unroll
    for (int o = 0; o < P::NOUT; ++o) acc[o][0] = acc[o][1] = acc[o][2] = 0.0;
#pragma unroll
    for (int q = 0; q < 6; ++q) {
      double uq[P::NIN], xq[P::NIN], g[P::NOUT];
#pragma unroll
      for (int f = 0; f < P::NIN; ++f) {
        uq[f] = sf(q, 0) * u[f][tri(t, 0)] + sf(q, 1) * u[f][tri(t, 1)] + sf(q, 2) * u[f][tri(t, 2)];
        xq[f] = JAC ? sf(q, 0) * x[f][tri(t, 0)] + sf(q, 1) * x[f][tri(t, 1)] + sf(q, 2) * x[f][tri(t, 2)] : 0.0;
      }
      if (JAC) prob.dg(uq, xq, g); else prob.g(uq, g);
#pragma unroll
      for (int o = 0; o < P::NOUT; ++o) {
        acc[o][0] += sf(q, 0) * g[o];
        acc[o][1] += sf(q, 1) * g[o];
        acc[o][2] += sf(q, 2) * g[o];
      }
    }
#pragma unroll
    for (int o = 0; o < P::NOUT; ++o)
#pragma unroll
      for (int a = 0; a < 3; ++a) out[o][tri(t, a)] += (detJ / 12.0) * acc[o][a];
  }
}

// int_{cube} e(u_h), int_{cube} m(u_h)  (all NF fields)
template <class P>
PFH_DEV void cube_energy(const P& prob, const double (&u)[P::NF][8], double detJ, double& E, double& Mm) {
#pragma unroll
  for (int t = 0; t < 6; ++t) {
    double s[P::NF], uq[P::NF];
#pragma unroll
    for (int f = 0; f < P::NF; ++f) {
      s[f] = u[f][kuhn(t, 0)] + u[f][kuhn(t, 1)] + u[f][kuhn(t, 2)] + u[f][kuhn(t, 3)];
      uq[f] = 0.25 * s[f];
    }
    double e = TET_W0 * prob.e(uq), m = TET_W0 * prob.m(uq);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
#pragma unroll
      for (int f = 0; f < P::NF; ++f) uq[f] = s[f] * (1.0 / 6.0) + u[f][kuhn(t, j)] * (1.0 / 3.0);
      e += TET_W1 * prob.e(uq);
      m += TET_W1 * prob.m(uq);
    }
    E += detJ * e;
    Mm += detJ * m;
  }
}

template <class P>
PFH_DEV void square_energy(const P& prob, const double (&u)[P::NF][5], double detJ, double& E, double& Mm) {
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    double e = 0.0, m = 0.0;
#pragma unroll
    for (int q = 0; q < 6; ++q) {
      double uq[P::NF];
#pragma unroll
      for (int f = 0; f < P::NF; ++f)
        uq[f] = sf(q, 0) * u[f][tri(t, 0)] + sf(q, 1) * u[f][tri(t, 1)] + sf(q, 2) * u[f][tri(t, 2)];
      e += prob.e(uq);
      m += prob.m(uq);
    }
    E += (detJ / 12.0) * e;
    Mm += (detJ / 12.0) * m;
  }
}

}  // namespace elem
