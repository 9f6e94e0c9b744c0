// Transfer operators of the geometric multigrid on the 2-D crossed mesh (the Poisson block of bench6, dolfin/bench6.py:72-90).
//
// Halving the cells per direction of a crossed mesh gives a crossed mesh whose triangles are unions of fine triangles (the
// coarse diagonals are fine diagonals), so the P1 spaces are nested: the coarse nodes are the fine corners with even indices
// (coarse corners) and the fine corners with odd indices (coarse cell centres), and the prolongation is plain P1
// interpolation of the coarse function at the fine nodes:
//   fine corner (even, even) = coarse corner;  (odd, odd) = coarse centre;  (odd, even) / (even, odd) = mean of the two
//   coarse corners of the coarse edge it halves;  fine centre = mean of the nearest coarse corner and the coarse centre of
//   the coarse square it lies in (it is the midpoint of that half-diagonal).
// The restriction is its transpose; with nested spaces the re-discretised coarse stiffness matrix IS the Galerkin
// operator P^T K P.  Both are written as gathers (one output node per call, no atomics) and are __host__ __device__:
// tests/emul runs them on the CPU against explicitly assembled P, P^T.
#pragma once
#include "pfh_vertex.cuh"

namespace mg2 {

// (P xc)_i for fine node i (DOLFIN order of the fine mesh: corners, then centres)
PFH_HD double prolong_node(const Mesh2& mf, const Mesh2& mc, const double* xc, long long i) {
  if (i < mf.ncorner) {
    const int ix = (int)(i % mf.nxv), iy = (int)(i / mf.nxv);
    const int I = ix >> 1, J = iy >> 1;
    if ((ix & 1) == 0 && (iy & 1) == 0) return vtx::ldc(xc, mc, I, J);
    if ((ix & 1) && (iy & 1)) return vtx::ldz(xc, mc, I, J);
    if (ix & 1) return 0.5 * (vtx::ldc(xc, mc, I, J) + vtx::ldc(xc, mc, I + 1, J));
    return 0.5 * (vtx::ldc(xc, mc, I, J) + vtx::ldc(xc, mc, I, J + 1));
  }
  const long long ci = i - mf.ncorner;
  const int cx = (int)(ci % mf.nx), cy = (int)(ci / mf.nx);
  const int I = cx >> 1, J = cy >> 1;
  return 0.5 * (vtx::ldc(xc, mc, I + (cx & 1), J + (cy & 1)) + vtx::ldz(xc, mc, I, J));
}

// (P^T t)_i for coarse node i (DOLFIN order of the coarse mesh); ldc / ldz read 0 outside the fine mesh
PFH_HD double restrict_node(const Mesh2& mf, const Mesh2& mc, const double* t, long long i) {
  if (i < mc.ncorner) {
    const int I = (int)(i % mc.nxv), J = (int)(i / mc.nxv);
    const int x = 2 * I, y = 2 * J;
    const double edges = (vtx::ldc(t, mf, x + 1, y) + vtx::ldc(t, mf, x - 1, y)) + (vtx::ldc(t, mf, x, y + 1) + vtx::ldc(t, mf, x, y - 1));
    const double cents = (vtx::ldz(t, mf, x, y) + vtx::ldz(t, mf, x - 1, y)) + (vtx::ldz(t, mf, x, y - 1) + vtx::ldz(t, mf, x - 1, y - 1));
    return vtx::ldc(t, mf, x, y) + 0.5 * (edges + cents);
  }
  const long long ci = i - mc.ncorner;
  const int I = (int)(ci % mc.nx), J = (int)(ci / mc.nx);
  const int x = 2 * I, y = 2 * J;
  const double cents = (vtx::ldz(t, mf, x, y) + vtx::ldz(t, mf, x + 1, y)) + (vtx::ldz(t, mf, x, y + 1) + vtx::ldz(t, mf, x + 1, y + 1));
  return vtx::ldc(t, mf, x + 1, y + 1) + 0.5 * cents;
}

// Dirichlet vertices of the scalar operators with scalar_bc = 1: the corner vertices of x = 0 and x = Lx
PFH_HD bool is_dirichlet(const Mesh2& m, long long i) {
  if (i >= m.ncorner) return false;
  const int ix = (int)(i % m.nxv);
  return ix == 0 || ix == m.nxv - 1;
}

}  // namespace mg2
