// Per-vertex (row-wise) evaluation of the matrix-free operators: one call computes all NF
// entries of the residual / Jacobian-vector product / smoother update that belong to one mesh
// vertex, gathering from the neighbouring vertices.  The CUDA kernels in pfh_stencil.cu are
// thread->vertex mappings around these functions; being __host__ __device__ they are also
// compiled for the CPU by tests/emul to check the index arithmetic against the oracle.
//
// What is computed (reference forms dolfin/pfbase.py:361-460):
//   residual : R_a = sum_b (Mc_ab M + Kc_ab K) u_b - Mp_a M u0_a + N[g_a](u)
//   jac_apply: y_a = sum_b (Mc_ab M + Kc_ab K) x_b + N[dg_a(u; x)]          (= df.derivative(F, w, dw) . x)
// M, K rows come from the stencil tables (pfh_tables.cpp), N[.] from the degree-3 quadrature
// of the cells incident to the vertex (pfh_elem.cuh).
#pragma once
#include "pfh_elem.cuh"

#ifdef __CUDA_ARCH__
#define PFH_LDG(p) __ldg(p)
#else
#define PFH_LDG(p) (*(p))
#endif

namespace vtx {

// =============================== 3-D Kuhn mesh ===============================
struct Planes3 {
  const double *pm, *p0, *pp;  // planes zl-1, zl, zl+1 (nullptr outside the global domain)
};

PFH_HD const double* plane_ptr(const FieldRef& f, const Mesh3& m, int zl) {
  const int gz = m.z0 + zl;
  if (gz < 0 || gz >= m.nzv) return nullptr;
  if (zl < 0) return f.glo;
  if (zl >= m.nzl) return f.ghi;
  return f.p + (long long)zl * m.nxy;
}
PFH_HD Planes3 planes(const FieldRef& f, const Mesh3& m, int zl) {
  Planes3 P;
  P.pm = plane_ptr(f, m, zl - 1);
  P.p0 = plane_ptr(f, m, zl);
  P.pp = plane_ptr(f, m, zl + 1);
  return P;
}
PFH_HD int type3(const Mesh3& m, int ix, int iy, int gz) {
  const int tx = ix == 0 ? 0 : (ix == m.nxv - 1 ? 2 : 1);
  const int ty = iy == 0 ? 0 : (iy == m.nyv - 1 ? 2 : 1);
  const int tz = gz == 0 ? 0 : (gz == m.nzv - 1 ? 2 : 1);
  return tz * 9 + ty * 3 + tx;
}
struct Row3 {
  double m[PFH_NOFF3];
  double k[PFH_NKOFF3];
};
PFH_HD void get_row3(const Mesh3& m, int type, Row3& r) {
  if (type == 13) {
#pragma unroll
    for (int o = 0; o < PFH_NOFF3; ++o) r.m[o] = m.mI[o];
#pragma unroll
    for (int o = 0; o < PFH_NKOFF3; ++o) r.k[o] = m.kI[o];
  } else {
#pragma unroll
    for (int o = 0; o < PFH_NOFF3; ++o) r.m[o] = PFH_LDG(m.mT + type * PFH_NOFF3 + o);
#pragma unroll
    for (int o = 0; o < PFH_NKOFF3; ++o) r.k[o] = PFH_LDG(m.kT + type * PFH_NKOFF3 + o);
  }
}

template <bool GUARD>
PFH_HD double ld3(const double* p, const Mesh3& m, int ix, int iy, int dx, int dy) {
  if (GUARD) {
    if (p == nullptr) return 0.0;
    const int x = ix + dx, y = iy + dy;
    if (x < 0 || x >= m.nxv || y < 0 || y >= m.nyv) return 0.0;
  }
  return PFH_LDG(p + ((long long)(iy + dy) * m.nxv + (ix + dx)));
}

// the 15 stencil neighbours in the offset order documented in pfh_internal.h
template <bool GUARD>
PFH_HD void gather15(const Planes3& P, const Mesh3& m, int ix, int iy, double (&v)[PFH_NOFF3]) {
  v[0] = ld3<GUARD>(P.p0, m, ix, iy, 0, 0);
  v[1] = ld3<GUARD>(P.p0, m, ix, iy, 1, 0);
  v[2] = ld3<GUARD>(P.p0, m, ix, iy, -1, 0);
  v[3] = ld3<GUARD>(P.p0, m, ix, iy, 0, 1);
  v[4] = ld3<GUARD>(P.p0, m, ix, iy, 0, -1);
  v[5] = ld3<GUARD>(P.pp, m, ix, iy, 0, 0);
  v[6] = ld3<GUARD>(P.pm, m, ix, iy, 0, 0);
  v[7] = ld3<GUARD>(P.p0, m, ix, iy, 1, 1);
  v[8] = ld3<GUARD>(P.p0, m, ix, iy, -1, -1);
  v[9] = ld3<GUARD>(P.pp, m, ix, iy, 0, 1);
  v[10] = ld3<GUARD>(P.pm, m, ix, iy, 0, -1);
  v[11] = ld3<GUARD>(P.pp, m, ix, iy, 1, 0);
  v[12] = ld3<GUARD>(P.pm, m, ix, iy, -1, 0);
  v[13] = ld3<GUARD>(P.pp, m, ix, iy, 1, 1);
  v[14] = ld3<GUARD>(P.pm, m, ix, iy, -1, -1);
}

PFH_HD void mk_sums3(const Row3& r, const double (&v)[PFH_NOFF3], double& Ms, double& Ks) {
  double ms = 0.0, ks = 0.0;
#pragma unroll
  for (int o = 0; o < PFH_NOFF3; ++o) ms += r.m[o] * v[o];
#pragma unroll
  for (int o = 0; o < PFH_NKOFF3; ++o) ks += r.k[o] * v[o];
  Ms = ms;
  Ks = ks;
}

// (M v)_i and (K v)_i of a scalar field at vertex (ix, iy, zl)
PFH_HD void lin_sums3(const FieldRef& f, const Mesh3& m, const Row3& r, bool guard, int ix, int iy, int zl, double& self,
                      double& Ms, double& Ks) {
  const Planes3 P = planes(f, m, zl);
  double v[PFH_NOFF3];
  if (guard) gather15<true>(P, m, ix, iy, v); else gather15<false>(P, m, ix, iy, v);
  self = v[0];
  mk_sums3(r, v, Ms, Ks);
}

// 27-neighbourhood of a field, index (dz+1)*9 + (dy+1)*3 + (dx+1); zeros outside the domain
PFH_HD void gather27(const FieldRef& f, const Mesh3& m, int ix, int iy, int zl, double (&nb)[27]) {
  const Planes3 P = planes(f, m, zl);
#pragma unroll
  for (int dz = -1; dz <= 1; ++dz) {
    const double* p = dz < 0 ? P.pm : (dz == 0 ? P.p0 : P.pp);
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
      for (int dx = -1; dx <= 1; ++dx) nb[(dz + 1) * 9 + (dy + 1) * 3 + (dx + 1)] = ld3<true>(p, m, ix, iy, dx, dy);
  }
}

// nonlinear load at one vertex by gathering over its (up to) 8 incident cubes
template <class P, bool JAC>
PFH_HD void nl_gather3(const P& prob, const Mesh3& m, const FieldRef* U, const FieldRef* X, int ix, int iy, int zl,
                       double (&acc)[P::NOUT]) {
  double nu[P::NIN][27], nx[P::NIN][27];
#pragma unroll
  for (int f = 0; f < P::NIN; ++f) {
    gather27(U[P::in(f)], m, ix, iy, zl, nu[f]);
    if (JAC) gather27(X[P::in(f)], m, ix, iy, zl, nx[f]);
  }
  const int gz = m.z0 + zl;
#pragma unroll
  for (int o = 0; o < P::NOUT; ++o) acc[o] = 0.0;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const int bx = c & 1, by = (c >> 1) & 1, bz = (c >> 2) & 1;
    const int ox = ix - bx, oy = iy - by, oz = gz - bz;
    if (ox < 0 || ox > m.nxv - 2 || oy < 0 || oy > m.nyv - 2 || oz < 0 || oz > m.nzv - 2) continue;
    double u[P::NIN][8], x[P::NIN][8], out[P::NOUT][8];
#pragma unroll
    for (int l = 0; l < 8; ++l) {
      const int lx = l & 1, ly = (l >> 1) & 1, lz = (l >> 2) & 1;
      const int idx = (lz - bz + 1) * 9 + (ly - by + 1) * 3 + (lx - bx + 1);
#pragma unroll
      for (int f = 0; f < P::NIN; ++f) {
        u[f][l] = nu[f][idx];
        x[f][l] = JAC ? nx[f][idx] : 0.0;
      }
    }
#pragma unroll
    for (int o = 0; o < P::NOUT; ++o)
#pragma unroll
      for (int l = 0; l < 8; ++l) out[o][l] = 0.0;
    elem::cube_load<P, JAC>(prob, u, x, m.detJ, out);
#pragma unroll
    for (int o = 0; o < P::NOUT; ++o) acc[o] += out[o][c];  // our local index in that cube is c
  }
}

template <class P, bool JAC>
PFH_HD void vertex_op3(const P& prob, const Mesh3& m, const LinCoefs& L, const FieldRef* U, const FieldRef* X,
                       const FieldRef* U0, int ix, int iy, int zl, double (&out)[P::NF]) {
  const int type = type3(m, ix, iy, m.z0 + zl);
  const bool guard = type != 13;
  Row3 r;
  get_row3(m, type, r);
  double Ms[P::NF], Ks[P::NF], self;
#pragma unroll
  for (int b = 0; b < P::NF; ++b) lin_sums3(JAC ? X[b] : U[b], m, r, guard, ix, iy, zl, self, Ms[b], Ks[b]);
#pragma unroll
  for (int a = 0; a < P::NF; ++a) {
    double s = 0.0;
#pragma unroll
    for (int b = 0; b < P::NF; ++b) s += L.Mc[a][b] * Ms[b] + L.Kc[a][b] * Ks[b];
    if (!JAC && L.Mp[a] != 0.0) {
      double m0, k0;
      lin_sums3(U0[a], m, r, guard, ix, iy, zl, self, m0, k0);
      s -= L.Mp[a] * m0;
    }
    out[a] = s;
  }
  double acc[P::NOUT];
  nl_gather3<P, JAC>(prob, m, U, X, ix, iy, zl, acc);
#pragma unroll
  for (int o = 0; o < P::NOUT; ++o) out[P::out(o)] += acc[o];
}

// energy / mass integrands attributed to one vertex: gradient energy u_i (K u)_i and the
// quadrature of the cube whose origin is this vertex
template <class P>
PFH_HD void vertex_energy3(const P& prob, const Mesh3& m, const FieldRef* U, int ix, int iy, int zl, double& E, double& Mm) {
  const int gz = m.z0 + zl;
  const int type = type3(m, ix, iy, gz);
  Row3 r;
  get_row3(m, type, r);
  double e = 0.0, mm = 0.0;
#pragma unroll
  for (int g = 0; g < P::n_grad(); ++g) {
    double self, Ms, Ks;
    lin_sums3(U[prob.grad_field(g)], m, r, type != 13, ix, iy, zl, self, Ms, Ks);
    e += prob.grad_coef(g) * self * Ks;
  }
  if (ix < m.nxv - 1 && iy < m.nyv - 1 && gz < m.nzv - 1) {
    double u[P::NF][8];
#pragma unroll
    for (int f = 0; f < P::NF; ++f) {
      const Planes3 Pl = planes(U[f], m, zl);
#pragma unroll
      for (int l = 0; l < 8; ++l)
        u[f][l] = ld3<false>((l >> 2) ? Pl.pp : Pl.p0, m, ix, iy, l & 1, (l >> 1) & 1);
    }
    elem::cube_energy<P>(prob, u, m.detJ, e, mm);
  }
  E = e;
  Mm = mm;
}

// scalar operator A = am M + ak K at one vertex: returns (A v)_i, the diagonal and v_i
PFH_HD void scalar_op3(const Mesh3& m, double am, double ak, const FieldRef& f, int ix, int iy, int zl, double& Av,
                       double& diag, double& self) {
  const int type = type3(m, ix, iy, m.z0 + zl);
  Row3 r;
  get_row3(m, type, r);
  double Ms, Ks;
  lin_sums3(f, m, r, type != 13, ix, iy, zl, self, Ms, Ks);
  Av = am * Ms + ak * Ks;
  diag = am * r.m[0] + ak * r.k[0];
}
PFH_HD double scalar_diag3(const Mesh3& m, double am, double ak, int ix, int iy, int zl) {
  const int type = type3(m, ix, iy, m.z0 + zl);
  if (type == 13) return am * m.mI[0] + ak * m.kI[0];
  return am * PFH_LDG(m.mT + type * PFH_NOFF3) + ak * PFH_LDG(m.kT + type * PFH_NKOFF3);
}

// =============================== 2-D crossed mesh ===============================
PFH_HD int type2(const Mesh2& m, int ix, int iy) {
  const int tx = ix == 0 ? 0 : (ix == m.nxv - 1 ? 2 : 1);
  const int ty = iy == 0 ? 0 : (iy == m.nyv - 1 ? 2 : 1);
  return ty * 3 + tx;
}
struct Row2 {
  double m[PFH_NOFF2], k[PFH_NOFF2];
};
PFH_HD void get_row2(const Mesh2& m, int type, Row2& r) {
  if (type == 4) {
#pragma unroll
    for (int o = 0; o < PFH_NOFF2; ++o) {
      r.m[o] = m.mI[o];
      r.k[o] = m.kI[o];
    }
  } else {
#pragma unroll
    for (int o = 0; o < PFH_NOFF2; ++o) {
      r.m[o] = PFH_LDG(m.mT + type * PFH_NOFF2 + o);
      r.k[o] = PFH_LDG(m.kT + type * PFH_NOFF2 + o);
    }
  }
}
PFH_HD double ldc(const double* p, const Mesh2& m, int ix, int iy) {  // corner value or 0 outside
  if (ix < 0 || ix >= m.nxv || iy < 0 || iy >= m.nyv) return 0.0;
  return PFH_LDG(p + ((long long)iy * m.nxv + ix));
}
PFH_HD double ldz(const double* p, const Mesh2& m, int cx, int cy) {  // centre value or 0 outside
  if (cx < 0 || cx >= m.nx || cy < 0 || cy >= m.ny) return 0.0;
  return PFH_LDG(p + (m.ncorner + (long long)cy * m.nx + cx));
}
// corner neighbourhood: 0..8 = 3x3 corners ((dy+1)*3 + dx+1), 9..12 = centres NE, NW, SE, SW
PFH_HD void gather13(const double* p, const Mesh2& m, int ix, int iy, double (&nb)[13]) {
#pragma unroll
  for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
    for (int dx = -1; dx <= 1; ++dx) nb[(dy + 1) * 3 + dx + 1] = ldc(p, m, ix + dx, iy + dy);
  nb[9] = ldz(p, m, ix, iy);
  nb[10] = ldz(p, m, ix - 1, iy);
  nb[11] = ldz(p, m, ix, iy - 1);
  nb[12] = ldz(p, m, ix - 1, iy - 1);
}
PFH_HD void mk_sums2c(const Row2& r, const double (&nb)[13], double& Ms, double& Ks) {
  // offsets: 0 self | 1 +x 2 -x 3 +y 4 -y | 5 NE 6 NW 7 SE 8 SW
  const double v[PFH_NOFF2] = {nb[4], nb[5], nb[3], nb[7], nb[1], nb[9], nb[10], nb[11], nb[12]};
  double ms = 0.0, ks = 0.0;
#pragma unroll
  for (int o = 0; o < PFH_NOFF2; ++o) {
    ms += r.m[o] * v[o];
    ks += r.k[o] * v[o];
  }
  Ms = ms;
  Ks = ks;
}
// centre neighbourhood: 0..3 corners v0..v3, 4 self
PFH_HD void gather5(const double* p, const Mesh2& m, int cx, int cy, double (&q)[5]) {
  const long long v0 = (long long)cy * m.nxv + cx;
  q[0] = PFH_LDG(p + v0);
  q[1] = PFH_LDG(p + v0 + 1);
  q[2] = PFH_LDG(p + v0 + m.nxv);
  q[3] = PFH_LDG(p + v0 + m.nxv + 1);
  q[4] = PFH_LDG(p + (m.ncorner + (long long)cy * m.nx + cx));
}
PFH_HD void mk_sums2z(const Mesh2& m, const double (&q)[5], double& Ms, double& Ks) {
  Ms = m.mC[0] * q[4] + m.mC[1] * q[0] + m.mC[2] * q[1] + m.mC[3] * q[2] + m.mC[4] * q[3];
  Ks = m.kC[0] * q[4] + m.kC[1] * q[0] + m.kC[2] * q[1] + m.kC[3] * q[2] + m.kC[4] * q[3];
}

// one vertex i (DOLFIN order: corners first, then centres) of the 2-D operators
// corner vertex (ix, iy)
template <class P, bool JAC>
PFH_HD void vertex_op2_corner(const P& prob, const Mesh2& m, const LinCoefs& L, const double* U, const double* X,
                              const double* U0, int ix, int iy, double (&out)[P::NF]) {
  const long long n = m.n_owned;
  {
    Row2 r;
    get_row2(m, type2(m, ix, iy), r);
    double nu[P::NF][13];
    double Ms[P::NF], Ks[P::NF];
#pragma unroll
    for (int b = 0; b < P::NF; ++b) {
      gather13((JAC ? X : U) + b * n, m, ix, iy, nu[b]);
      mk_sums2c(r, nu[b], Ms[b], Ks[b]);
    }
#pragma unroll
    for (int a = 0; a < P::NF; ++a) {
      double s = 0.0;
#pragma unroll
      for (int b = 0; b < P::NF; ++b) s += L.Mc[a][b] * Ms[b] + L.Kc[a][b] * Ks[b];
      if (!JAC && L.Mp[a] != 0.0) {
        double n0[13], m0, k0;
        gather13(U0 + a * n, m, ix, iy, n0);
        mk_sums2c(r, n0, m0, k0);
        s -= L.Mp[a] * m0;
      }
      out[a] = s;
    }
    // nonlinear load: squares NE, NW, SE, SW; our local corner index in square s is s
    double un[P::NIN][13], xn[P::NIN][13];
#pragma unroll
    for (int f = 0; f < P::NIN; ++f) {
      if (JAC) {
        gather13(U + P::in(f) * n, m, ix, iy, un[f]);
#pragma unroll
        for (int k = 0; k < 13; ++k) xn[f][k] = nu[P::in(f)][k];
      } else {
#pragma unroll
        for (int k = 0; k < 13; ++k) {
          un[f][k] = nu[P::in(f)][k];
          xn[f][k] = 0.0;
        }
      }
    }
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      const int bx = s & 1, by = (s >> 1) & 1;
      const int ox = ix - bx, oy = iy - by;
      if (ox < 0 || ox > m.nx - 1 || oy < 0 || oy > m.ny - 1) continue;
      double u[P::NIN][5], x[P::NIN][5], o5[P::NOUT][5];
#pragma unroll
      for (int f = 0; f < P::NIN; ++f) {
#pragma unroll
        for (int l = 0; l < 4; ++l) {
          const int idx = ((l >> 1) - by + 1) * 3 + ((l & 1) - bx + 1);
          u[f][l] = un[f][idx];
          x[f][l] = xn[f][idx];
        }
        u[f][4] = un[f][9 + s];
        x[f][4] = xn[f][9 + s];
      }
#pragma unroll
      for (int o = 0; o < P::NOUT; ++o)
#pragma unroll
        for (int l = 0; l < 5; ++l) o5[o][l] = 0.0;
      elem::square_load<P, JAC>(prob, u, x, m.detJ, o5);
#pragma unroll
      for (int o = 0; o < P::NOUT; ++o) out[P::out(o)] += o5[o][s];
    }
    if (P::HAS_BC) {
#pragma unroll
      for (int a = 0; a < P::NF; ++a) {
        double val;
        if (prob.dirichlet(a, ix, iy, m.nx, iy * m.hy, val)) out[a] = JAC ? nu[a][4] : nu[a][4] - val;
      }
    }
  }
}

// cell centre (cx, cy)
template <class P, bool JAC>
PFH_HD void vertex_op2_centre(const P& prob, const Mesh2& m, const LinCoefs& L, const double* U, const double* X,
                              const double* U0, int cx, int cy, double (&out)[P::NF]) {
  const long long n = m.n_owned;
  {
    double q[P::NF][5], Ms[P::NF], Ks[P::NF];
#pragma unroll
    for (int b = 0; b < P::NF; ++b) {
      gather5((JAC ? X : U) + b * n, m, cx, cy, q[b]);
      mk_sums2z(m, q[b], Ms[b], Ks[b]);
    }
#pragma unroll
    for (int a = 0; a < P::NF; ++a) {
      double s = 0.0;
#pragma unroll
      for (int b = 0; b < P::NF; ++b) s += L.Mc[a][b] * Ms[b] + L.Kc[a][b] * Ks[b];
      if (!JAC && L.Mp[a] != 0.0) {
        double q0[5], m0, k0;
        gather5(U0 + a * n, m, cx, cy, q0);
        mk_sums2z(m, q0, m0, k0);
        s -= L.Mp[a] * m0;
      }
      out[a] = s;
    }
    double u[P::NIN][5], x[P::NIN][5], o5[P::NOUT][5];
#pragma unroll
    for (int f = 0; f < P::NIN; ++f) {
      if (JAC) {
        gather5(U + P::in(f) * n, m, cx, cy, u[f]);
#pragma unroll
        for (int l = 0; l < 5; ++l) x[f][l] = q[P::in(f)][l];
      } else {
#pragma unroll
        for (int l = 0; l < 5; ++l) {
          u[f][l] = q[P::in(f)][l];
          x[f][l] = 0.0;
        }
      }
    }
#pragma unroll
    for (int o = 0; o < P::NOUT; ++o)
#pragma unroll
      for (int l = 0; l < 5; ++l) o5[o][l] = 0.0;
    elem::square_load<P, JAC>(prob, u, x, m.detJ, o5);
#pragma unroll
    for (int o = 0; o < P::NOUT; ++o) out[P::out(o)] += o5[o][4];
  }
}

// one vertex i in DOLFIN order (corners first, then centres)
template <class P, bool JAC>
PFH_HD void vertex_op2(const P& prob, const Mesh2& m, const LinCoefs& L, const double* U, const double* X, const double* U0,
                       long long i, double (&out)[P::NF]) {
  if (i < m.ncorner) {
    vertex_op2_corner<P, JAC>(prob, m, L, U, X, U0, (int)(i % m.nxv), (int)(i / m.nxv), out);
  } else {
    const long long ci = i - m.ncorner;
    vertex_op2_centre<P, JAC>(prob, m, L, U, X, U0, (int)(ci % m.nx), (int)(ci / m.nx), out);
  }
}

template <class P>
PFH_HD void vertex_energy2(const P& prob, const Mesh2& m, const double* U, long long i, double& E, double& Mm) {
  const long long n = m.n_owned;
  double e = 0.0, mm = 0.0;
  if (i < m.ncorner) {
    const int ix = (int)(i % m.nxv), iy = (int)(i / m.nxv);
    Row2 r;
    get_row2(m, type2(m, ix, iy), r);
#pragma unroll
    for (int g = 0; g < P::n_grad(); ++g) {
      double nb[13], Ms, Ks;
      gather13(U + prob.grad_field(g) * n, m, ix, iy, nb);
      mk_sums2c(r, nb, Ms, Ks);
      e += prob.grad_coef(g) * nb[4] * Ks;
    }
  } else {
    const long long ci = i - m.ncorner;
    const int cx = (int)(ci % m.nx), cy = (int)(ci / m.nx);
    double u[P::NF][5];
#pragma unroll
    for (int f = 0; f < P::NF; ++f) gather5(U + f * n, m, cx, cy, u[f]);
#pragma unroll
    for (int g = 0; g < P::n_grad(); ++g) {
      double Ms, Ks;
      mk_sums2z(m, u[prob.grad_field(g)], Ms, Ks);
      e += prob.grad_coef(g) * u[prob.grad_field(g)][4] * Ks;
    }
    elem::square_energy<P>(prob, u, m.detJ, e, mm);
  }
  E = e;
  Mm = mm;
}

// scalar operator A = am M + ak K at vertex i of the crossed mesh
PFH_HD void scalar_op2(const Mesh2& m, double am, double ak, const double* f, long long i, double& Av, double& diag,
                       double& self) {
  double Ms, Ks;
  if (i < m.ncorner) {
    const int ix = (int)(i % m.nxv), iy = (int)(i / m.nxv);
    Row2 r;
    get_row2(m, type2(m, ix, iy), r);
    double nb[13];
    gather13(f, m, ix, iy, nb);
    mk_sums2c(r, nb, Ms, Ks);
    self = nb[4];
    diag = am * r.m[0] + ak * r.k[0];
  } else {
    const long long ci = i - m.ncorner;
    double q[5];
    gather5(f, m, (int)(ci % m.nx), (int)(ci / m.nx), q);
    mk_sums2z(m, q, Ms, Ks);
    self = q[4];
    diag = am * m.mC[0] + ak * m.kC[0];
  }
  Av = am * Ms + ak * Ks;
}
PFH_HD double scalar_diag2(const Mesh2& m, double am, double ak, long long i) {
  if (i >= m.ncorner) return am * m.mC[0] + ak * m.kC[0];
  const int type = type2(m, (int)(i % m.nxv), (int)(i / m.nxv));
  if (type == 4) return am * m.mI[0] + ak * m.kI[0];
  return am * PFH_LDG(m.mT + type * PFH_NOFF2) + ak * PFH_LDG(m.kT + type * PFH_NOFF2);
}

}  // namespace vtx
