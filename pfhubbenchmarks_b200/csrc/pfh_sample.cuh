// Point evaluation of a P1 vertex field on the two structured simplicial meshes: what `f(x)` does inside the
// reference's peval / sample (dolfin/pfbase.py:50-94, a bounding-box-tree search in DOLFIN).  Here the containing
// simplex follows from the point's cell and the ordering of its local coordinates.
#pragma once
#include <cmath>

#include "pfh_vertex.cuh"

namespace smp {

PFH_HD int cell_of(double f, int ncell) {
  int i = (int)floor(f);
  return i < 0 ? 0 : (i > ncell - 1 ? ncell - 1 : i);
}

// Crossed mesh: square (ix, iy) is split by its centre vertex into bottom / left / right / top triangles.
// Returns false for a point outside the rectangle (beyond a 1e-12 h slack).
PFH_HD bool eval2(const Mesh2& m, const double* u, double x, double y, double& val) {
  const double fx = x / m.hx, fy = y / m.hy, tol = 1e-12;
  if (fx < -tol || fy < -tol || fx > m.nx + tol || fy > m.ny + tol) return false;
  const int ix = cell_of(fx, m.nx), iy = cell_of(fy, m.ny);
  const double xi = fx - ix, et = fy - iy;
  const double* row = u + (long long)iy * m.nxv + ix;
  const double v0 = row[0], v1 = row[1], v2 = row[m.nxv], v3 = row[m.nxv + 1];
  const double vm = u[m.ncorner + (long long)iy * m.nx + ix];
  if (et <= xi && et <= 1.0 - xi) val = v0 * (1.0 - xi - et) + v1 * (xi - et) + vm * (2.0 * et);               // bottom
  else if (xi <= et && xi <= 1.0 - et) val = v0 * (1.0 - xi - et) + v2 * (et - xi) + vm * (2.0 * xi);          // left
  else if (xi >= et) val = v1 * (xi - et) + v3 * (xi + et - 1.0) + vm * (2.0 * (1.0 - xi));                    // right
  else val = v2 * (et - xi) + v3 * (xi + et - 1.0) + vm * (2.0 * (1.0 - et));                                  // top
  return true;
}

// Kuhn mesh: the six tets of a cube are the six orderings of the local coordinates; with t0 >= t1 >= t2 along
// axes a0, a1, a2 the interpolant walks v0 -> +e_a0 -> +e_a1 -> +e_a2 (= v7).
// inside = false: outside the box.  owned = false: the cube's lower plane belongs to another slab (value untouched).
PFH_HD void eval3(const Mesh3& m, const FieldRef& f, double x, double y, double z, bool& inside, bool& owned, double& val) {
  const double fx = x / m.hx, fy = y / m.hy, fz = z / m.hz, tol = 1e-12;
  inside = !(fx < -tol || fy < -tol || fz < -tol || fx > m.nxv - 1 + tol || fy > m.nyv - 1 + tol || fz > m.nzv - 1 + tol);
  owned = false;
  if (!inside) return;
  const int ix = cell_of(fx, m.nxv - 1), iy = cell_of(fy, m.nyv - 1), iz = cell_of(fz, m.nzv - 1);
  const int zl = iz - m.z0;
  if (zl < 0 || zl >= m.nzl) return;
  owned = true;
  const double* pl[2] = {vtx::plane_ptr(f, m, zl), vtx::plane_ptr(f, m, zl + 1)};
  double t[3] = {fx - ix, fy - iy, fz - iz};
  int ax[3] = {0, 1, 2};
#define PFH_SWAP(i, j) if (t[i] < t[j]) { const double tt = t[i]; t[i] = t[j]; t[j] = tt; const int aa = ax[i]; ax[i] = ax[j]; ax[j] = aa; }
  PFH_SWAP(0, 1) PFH_SWAP(1, 2) PFH_SWAP(0, 1)
#undef PFH_SWAP
  int d[3] = {0, 0, 0};
  auto at = [&]() { return pl[d[2]][(long long)(iy + d[1]) * m.nxv + ix + d[0]]; };
  val = at() * (1.0 - t[0]);
  d[ax[0]] = 1;
  val += at() * (t[0] - t[1]);
  d[ax[1]] = 1;
  val += at() * (t[1] - t[2]);
  d[ax[2]] = 1;
  val += at() * t[2];
}

}  // namespace smp
