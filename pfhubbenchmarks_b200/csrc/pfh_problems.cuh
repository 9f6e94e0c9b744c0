// Pointwise physics of the PFHub problems as device functors.
//
// Each problem is in the common discrete form
//     R_a = sum_b (Mc[a][b] M + Kc[a][b] K) u_b - Mp[a] M u0_a + N[g_a](u),
// with N[g]_i = int phi_i g(u_h) by the degree-3 quadrature rule.  A functor supplies
//   NF            number of fields
//   NIN / in(i)   the fields the nonlinear functions read
//   NOUT / out(o) the residual rows that carry a nonlinear load
//   g(u, o[])     the loads g_out(o) at a quadrature point (u indexed by in())
//   dg(u, x, o[]) their directional derivative sum_b dg/du_b x_b (the Jacobian's M_{g'} blocks)
//   e(u), m(u)    energy density and "mass" density of the post-step reductions (all NF fields)
// Reference definitions: bench1.py:59-67, bench2.py:69-111, bench3.py:64-98, bench6.py:61-72
// with the weak forms of pfbase.py:361-460.  The reference obtains f' and f'' symbolically
// (df.diff / df.derivative); here they are written out by hand and checked against the
// sympy-differentiated oracle in tests/.
#pragma once
#include "pfh_internal.h"

#ifdef __CUDACC__
#define PFH_HD __host__ __device__ __forceinline__
#else
#define PFH_HD inline
#endif

// ---------------- bench1 / b13d : Cahn-Hilliard, fields (c, mu) ----------------
struct ProbCH {
  static constexpr int NF = 2, NIN = 1, NOUT = 1;
  static constexpr bool QUARTIC = true;  // dg(c; x) = (a2 c^2 + a1 c + a0) x : enables the low-flop cube routine
  PFH_HD void jac_poly(double& a2, double& a1, double& a0) const {  // -f''(c), f = rho (c-ca)^2 (cb-c)^2
    const double s = ca + cb;
    a2 = -12.0 * rho; a1 = 12.0 * rho * s; a0 = -rho * (2.0 * s * s + 4.0 * ca * cb);
  }
  PFH_HD void res_poly(double& a3, double& a2, double& a1, double& a0) const {  // g(c) = -f'(c)
    const double s = ca + cb;
    a3 = -4.0 * rho; a2 = 6.0 * rho * s; a1 = -2.0 * rho * (s * s + 2.0 * ca * cb); a0 = 2.0 * rho * s * ca * cb;
  }
  PFH_HD static constexpr int in(int) { return 0; }
  PFH_HD static constexpr int out(int) { return 1; }
  double ca, cb, rho, kappa, Mob;
  explicit ProbCH(const ProblemParams& pp) : ca(pp.p[0]), cb(pp.p[1]), rho(pp.p[2]), kappa(pp.p[3]), Mob(pp.p[4]) {}
  PFH_HD void g(const double* u, double* o) const {  // g_mu = -f'(c), f = rho (c-ca)^2 (cb-c)^2
    const double a = u[0] - ca, b = cb - u[0];
    o[0] = -2.0 * rho * a * b * (b - a);
  }
  PFH_HD void dg(const double* u, const double* x, double* o) const {
    const double a = u[0] - ca, b = cb - u[0], d = b - a;
    o[0] = -2.0 * rho * (d * d - 2.0 * a * b) * x[0];
  }
  PFH_HD double e(const double* u) const {
    const double a = u[0] - ca, b = cb - u[0];
    return rho * a * a * b * b;
  }
  PFH_HD double m(const double* u) const { return u[0]; }
  PFH_HD static constexpr int n_grad() { return 1; }
  PFH_HD int grad_field(int) const { return 0; }
  PFH_HD double grad_coef(int) const { return 0.5 * kappa; }
  PFH_HD bool dirichlet(int, int, int, int, double, double&) const { return false; }
  static constexpr bool HAS_BC = false;
  void lin(double dt, LinCoefs& L) const {
    L.Mc[0][0] = 1.0 / dt; L.Mp[0] = 1.0 / dt; L.Kc[0][1] = Mob;
    L.Mc[1][1] = 1.0; L.Kc[1][0] = -kappa;
  }
  // c = c0 + eps (cos(.105x)cos(.11y) + (cos(.13x)cos(.087y))^2 + cos(.025x-.15y)cos(.07x-.02y)), mu = 0
  // (pfbase.py:187-190; z is ignored in 3-D, b13d.py:55)
  PFH_HD void ic(const double* q, double x, double y, double* u) const {
    u[0] = q[0] + q[1] * (cos(0.105 * x) * cos(0.11 * y) + (cos(0.13 * x) * cos(0.087 * y)) * (cos(0.13 * x) * cos(0.087 * y)) +
                          cos(0.025 * x - 0.15 * y) * cos(0.07 * x - 0.02 * y));
    u[1] = 0.0;
  }
};

// ---------------- bench6 : CH + electrostatics, fields (c, mu, phi) ----------------
struct ProbBench6 {
  static constexpr int NF = 3, NIN = 1, NOUT = 1;
  static constexpr bool QUARTIC = true;
  PFH_HD void jac_poly(double& a2, double& a1, double& a0) const {
    const double s = ca + cb;
    a2 = -12.0 * rho; a1 = 12.0 * rho * s; a0 = -rho * (2.0 * s * s + 4.0 * ca * cb);
  }
  PFH_HD void res_poly(double& a3, double& a2, double& a1, double& a0) const {  // g(c) = -f'(c)
    const double s = ca + cb;
    a3 = -4.0 * rho; a2 = 6.0 * rho * s; a1 = -2.0 * rho * (s * s + 2.0 * ca * cb); a0 = 2.0 * rho * s * ca * cb;
  }
  PFH_HD static constexpr int in(int) { return 0; }
  PFH_HD static constexpr int out(int) { return 1; }
  double ca, cb, rho, kappa, Mob, k, eps;
  explicit ProbBench6(const ProblemParams& pp)
      : ca(pp.p[0]), cb(pp.p[1]), rho(pp.p[2]), kappa(pp.p[3]), Mob(pp.p[4]), k(pp.p[5]), eps(pp.p[6]) {}
  PFH_HD void g(const double* u, double* o) const {
    const double a = u[0] - ca, b = cb - u[0];
    o[0] = -2.0 * rho * a * b * (b - a);
  }
  PFH_HD void dg(const double* u, const double* x, double* o) const {
    const double a = u[0] - ca, b = cb - u[0], d = b - a;
    o[0] = -2.0 * rho * (d * d - 2.0 * a * b) * x[0];
  }
  PFH_HD double e(const double* u) const {  // f_chem + k c phi / 2   (bench6.py:61-65,158-165)
    const double a = u[0] - ca, b = cb - u[0];
    return rho * a * a * b * b + 0.5 * k * u[0] * u[2];
  }
  PFH_HD double m(const double* u) const { return u[0]; }
  PFH_HD static constexpr int n_grad() { return 1; }
  PFH_HD int grad_field(int) const { return 0; }
  PFH_HD double grad_coef(int) const { return 0.5 * kappa; }
  static constexpr bool HAS_BC = true;
  // DirichletBC on phi: 0 on x = 0, sin(y/7) on x = Lx (bench6.py:77-90); corner vertices only
  PFH_HD bool dirichlet(int field, int ix, int /*iy*/, int nx, double y, double& val) const {
    if (field != 2) return false;
    if (ix == 0) { val = 0.0; return true; }
    if (ix == nx) { val = sin(y / 7.0); return true; }
    return false;
  }
  void lin(double dt, LinCoefs& L) const {
    L.Mc[0][0] = 1.0 / dt; L.Mp[0] = 1.0 / dt; L.Kc[0][1] = Mob;
    L.Mc[1][1] = 1.0; L.Kc[1][0] = -kappa; L.Mc[1][2] = -k;
    L.Kc[2][2] = -1.0; L.Mc[2][0] = k / eps;
  }
  PFH_HD void ic(const double* q, double x, double y, double* u) const {  // pfbase.py:332-336
    u[0] = q[0] + q[1] * (cos(0.2 * x) * cos(0.11 * y) + (cos(0.13 * x) * cos(0.087 * y)) * (cos(0.13 * x) * cos(0.087 * y)) +
                          cos(0.025 * x - 0.15 * y) * cos(0.07 * x - 0.02 * y));
    u[1] = 0.0;
    u[2] = 0.0;
  }
};

// ---------------- bench3 : dendrite (isotropic as coded), fields (U, phi) ----------------
struct ProbBench3 {
  static constexpr int NF = 2, NIN = 2, NOUT = 2;
  static constexpr bool QUARTIC = false;
  PFH_HD static constexpr int in(int i) { return i; }
  PFH_HD static constexpr int out(int o) { return o; }
  double W, tau, D, lam, area;
  explicit ProbBench3(const ProblemParams& pp) : W(pp.p[0]), tau(pp.p[1]), D(pp.p[2]), lam(pp.p[3]), area(pp.p[4]) {}
  PFH_HD void g(const double* u, double* o) const {  // dfdp = (phi - lam U (1-phi^2)) (1-phi^2)  (bench3.py:81)
    const double U = u[0], p = u[1], q = 1.0 - p * p;
    const double dfdp = (p - lam * U * q) * q;
    o[0] = -(0.5 / tau) * dfdp;
    o[1] = -(1.0 / tau) * dfdp;
  }
  PFH_HD void dg(const double* u, const double* x, double* o) const {
    const double U = u[0], p = u[1], q = 1.0 - p * p;
    const double a = p - lam * U * q;
    const double da = x[1] - lam * x[0] * q + 2.0 * lam * U * p * x[1];
    const double d = da * q + a * (-2.0 * p * x[1]);
    o[0] = -(0.5 / tau) * d;
    o[1] = -(1.0 / tau) * d;
  }
  PFH_HD double e(const double* u) const {  // bench3.py:69-70
    const double U = u[0], p = u[1], p2 = p * p;
    return -0.5 * p2 + 0.25 * p2 * p2 + lam * U * p * (1.0 - (2.0 / 3.0) * p2 + 0.2 * p2 * p2);
  }
  PFH_HD double m(const double* u) const { return 0.5 * (u[1] + 1.0) / area; }  // solid fraction, bench3.py:169-170
  PFH_HD static constexpr int n_grad() { return 1; }
  PFH_HD int grad_field(int) const { return 1; }
  PFH_HD double grad_coef(int) const { return 0.5 * W * W; }
  static constexpr bool HAS_BC = false;
  PFH_HD bool dirichlet(int, int, int, int, double, double&) const { return false; }
  void lin(double dt, LinCoefs& L) const {
    L.Mc[0][0] = 1.0 / dt; L.Mp[0] = 1.0 / dt; L.Kc[0][0] = D; L.Kc[0][1] = 0.5 * W * W / tau;
    L.Mc[1][1] = 1.0 / dt; L.Mp[1] = 1.0 / dt; L.Kc[1][1] = W * W / tau;
  }
  PFH_HD void ic(const double* q, double x, double y, double* u) const {  // pfbase.py:306-317
    const double Delta = q[0], r0 = q[1], w = q[2], vin = q[3], vout = q[4];
    const double r = sqrt(x * x + y * y);
    u[0] = Delta;
    if (r < r0 - 0.5 * w) u[1] = vin;
    else if (r > r0 + 0.5 * w) u[1] = vout;
    else u[1] = vout + 0.5 * (vin - vout) * (1.0 + cos(3.14159265358979323846 * (r - r0 + 0.5 * w) / w));
  }
};

// ---------------- bench2 : Ostwald ripening, fields (c, mu, eta1..4) ----------------
struct ProbBench2 {
  static constexpr int NF = 6, NIN = 5, NOUT = 5;
  static constexpr bool QUARTIC = false;
  PFH_HD static constexpr int in(int i) { return i == 0 ? 0 : i + 1; }   // c, eta1..4
  PFH_HD static constexpr int out(int o) { return o + 1; }               // mu, eta1..4
  double ca, cb, rho2, kc, ke, Mob, w, alpha, L;
  explicit ProbBench2(const ProblemParams& pp)
      : ca(pp.p[0]), cb(pp.p[1]), rho2(pp.p[2]), kc(pp.p[3]), ke(pp.p[4]), Mob(pp.p[5]), w(pp.p[6]), alpha(pp.p[7]), L(pp.p[8]) {}
  // f = rho^2 (c-ca)^2 (1-H) + rho^2 (c-cb)^2 H + w [sum e^2(1-e)^2 + alpha sum_{i<j} ei^2 ej^2],
  // H = sum e^3 (6e^2 - 15e + 10)   (bench2.py:76-99)
  PFH_HD void g(const double* u, double* o) const {
    const double c = u[0];
    double H = 0.0, s2 = 0.0;
    for (int i = 0; i < 4; ++i) {
      const double e = u[1 + i];
      H += e * e * e * (6.0 * e * e - 15.0 * e + 10.0);
      s2 += e * e;
    }
    const double da = c - ca, db = c - cb;
    o[0] = -(2.0 * rho2 * da * (1.0 - H) + 2.0 * rho2 * db * H);
    const double dfdH = rho2 * (db * db - da * da);
    for (int i = 0; i < 4; ++i) {
      const double e = u[1 + i], om = 1.0 - e;
      const double hp = 30.0 * e * e * om * om;
      o[1 + i] = L * (dfdH * hp + w * (2.0 * e * om * om - 2.0 * e * e * om + 2.0 * alpha * e * (s2 - e * e)));
    }
  }
  PFH_HD void dg(const double* u, const double* x, double* o) const {
    const double c = u[0], xc = x[0];
    double s2 = 0.0, sx = 0.0, shx = 0.0;  // sum e^2, sum e x, sum h'(e) x
    for (int i = 0; i < 4; ++i) {
      const double e = u[1 + i], om = 1.0 - e;
      s2 += e * e;
      sx += e * x[1 + i];
      shx += 30.0 * e * e * om * om * x[1 + i];
    }
    const double da = c - ca, db = c - cb;
    o[0] = -(2.0 * rho2 * xc + 2.0 * rho2 * (ca - cb) * shx);
    const double dfdH = rho2 * (db * db - da * da);
    const double ddfdH = 2.0 * rho2 * (ca - cb) * xc;  // d/dc of dfdH times x_c
    for (int i = 0; i < 4; ++i) {
      const double e = u[1 + i], om = 1.0 - e, xi = x[1 + i];
      const double hp = 30.0 * e * e * om * om;
      const double hpp = 60.0 * e * om * (1.0 - 2.0 * e);
      const double dw = (2.0 * om * om - 8.0 * e * om + 2.0 * e * e) * xi + 2.0 * alpha * (s2 - e * e) * xi +
                        4.0 * alpha * e * (sx - e * xi);
      o[1 + i] = L * (ddfdH * hp + dfdH * hpp * xi + w * dw);
    }
  }
  PFH_HD double e(const double* u) const {
    const double c = u[0];
    double H = 0.0, W1 = 0.0, s2 = 0.0, s4 = 0.0;
    for (int i = 0; i < 4; ++i) {
      const double e = u[2 + i];
      H += e * e * e * (6.0 * e * e - 15.0 * e + 10.0);
      W1 += e * e * (1.0 - e) * (1.0 - e);
      s2 += e * e;
      s4 += e * e * e * e;
    }
    const double cross = 0.5 * (s2 * s2 - s4);  // sum_{i<j} ei^2 ej^2
    return rho2 * (c - ca) * (c - ca) * (1.0 - H) + rho2 * (c - cb) * (c - cb) * H + w * (W1 + alpha * cross);
  }
  PFH_HD double m(const double* u) const { return u[0]; }
  PFH_HD static constexpr int n_grad() { return 5; }
  PFH_HD int grad_field(int i) const { return i == 0 ? 0 : i + 1; }
  PFH_HD double grad_coef(int i) const { return i == 0 ? 0.5 * kc : 0.5 * ke; }
  static constexpr bool HAS_BC = false;
  PFH_HD bool dirichlet(int, int, int, int, double, double&) const { return false; }
  void lin(double dt, LinCoefs& Lc) const {
    Lc.Mc[0][0] = 1.0 / dt; Lc.Mp[0] = 1.0 / dt; Lc.Kc[0][1] = Mob;
    Lc.Mc[1][1] = 1.0; Lc.Kc[1][0] = -kc;
    for (int i = 2; i < 6; ++i) {
      Lc.Mc[i][i] = 1.0 / dt; Lc.Mp[i] = 1.0 / dt; Lc.Kc[i][i] = L * ke;
    }
  }
  // pfbase.py:274-288, including the i/ii inconsistency in the psi term that the golden data confirms
  PFH_HD void ic(const double* q, double x, double y, double* u) const {
    const double c0 = q[0], eps = q[1], eps_eta = q[2], psi = q[3];
    const double t2 = cos(0.13 * x) * cos(0.087 * y);
    u[0] = c0 + eps * (cos(0.105 * x) * cos(0.11 * y) + t2 * t2 + cos(0.025 * x - 0.15 * y) * cos(0.07 * x - 0.02 * y));
    u[1] = 0.0;
    for (int i = 0; i < 4; ++i) {
      const double ii = i + 1.0, di = (double)i;
      const double t3 = cos((0.046 + 0.001 * di) * x - (0.0405 + 0.001 * di) * y) *
                        cos((0.031 + 0.001 * di) * x - (0.004 + 0.001 * di) * y);
      const double v = cos((0.01 * ii) * x - 4.0) * cos((0.007 + 0.01 * ii) * y) +
                       cos((0.11 + 0.01 * ii) * x) * cos((0.11 + 0.01 * ii) * y) + psi * t3 * t3;
      u[2 + i] = eps_eta * v * v;
    }
  }
};

// A generated model (codegen.py): `struct ProbUser` with the interface of the functors above.
#ifdef PFH_USER_PROBLEM_HEADER
#include PFH_USER_PROBLEM_HEADER
#endif
