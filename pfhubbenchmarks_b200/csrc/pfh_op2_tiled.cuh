// Square-centric residual / J.v on the 2-D crossed mesh: the cube-centric idea of pfh_stencil3_tma.cu in two dimensions.
//
// The gather kernel k2_op (pfh_stencil.cu) evaluates, for every corner vertex, the four crossed squares around it (16
// triangles x 6 quadrature points) and for every cell centre its own square again: each square's quadrature is done five
// times.  Here a CTA of 32 x 8 threads owns 31 x 7 corner vertices and the cell centres of the squares whose lower-left
// corner they are; it stages the operand values of the 33 x 9 corners and 32 x 8 centres around the tile in shared memory
// (one read per array element + halo), thread (sx, sy) evaluates square (x0 + sx, y0 + sy) ONCE, keeps the centre's load for
// the centre row it writes itself and leaves the four corner loads in shared memory; after one barrier every corner vertex
// adds its four squares' entries to its stencil part, which also comes from the staged tile.
//
// The three phases are __host__ __device__ functions of (block origin, thread index, staged tile): the kernel calls them
// with a barrier in between, tests/emul calls them block by block on the CPU, so the index arithmetic is checked against
// the oracle without a GPU.  Same arithmetic as vtx::vertex_op2_corner / vertex_op2_centre (same element routine, same
// row tables, same order of the four squares).
#pragma once
#include "pfh_vertex.cuh"

namespace op2t {

constexpr int SX = 32, SY = 8;                                      // squares per CTA = threads per CTA
constexpr int VX = SX + 1, VY = SY + 1, WX = SX - 1, WY = SY - 1;   // corners staged / corners owned per CTA

// OP = operand (residual: the iterate u; J.v: x), EX = residual: previous step u0 (fields with Mp != 0), J.v: the fields of
// the linearisation point the nonlinear functions read
template <class P, bool JAC>
struct Tile {
  static constexpr int NEX = JAC ? P::NIN : P::NF;
  double opc[P::NF][VY][VX], opz[P::NF][SY][SX];
  double exc[NEX][VY][VX], exz[NEX][SY][SX];
  double cv[P::NOUT][4][SY][SX];                                    // corner loads of each square
};
template <class P, bool JAC>
constexpr bool fits() { return sizeof(Tile<P, JAC>) <= 48 * 1024; }   // static shared memory

template <class P, bool JAC>
PFH_HD void stage(Tile<P, JAC>& s, const Mesh2& m, const LinCoefs& L, const double* OP, const double* EX, int x0, int y0,
                  int tid, int nthreads) {
  const long long n = m.n_owned;
  for (int k = tid; k < VX * VY; k += nthreads) {
    const int lx = k % VX, ly = k / VX;
#pragma unroll
    for (int f = 0; f < P::NF; ++f) s.opc[f][ly][lx] = vtx::ldc(OP + f * n, m, x0 + lx, y0 + ly);
#pragma unroll
    for (int f = 0; f < Tile<P, JAC>::NEX; ++f)
      if (JAC || L.Mp[f] != 0.0) s.exc[f][ly][lx] = vtx::ldc(EX + (JAC ? P::in(f) : f) * n, m, x0 + lx, y0 + ly);
  }
  for (int k = tid; k < SX * SY; k += nthreads) {
    const int lx = k % SX, ly = k / SX;
#pragma unroll
    for (int f = 0; f < P::NF; ++f) s.opz[f][ly][lx] = vtx::ldz(OP + f * n, m, x0 + lx, y0 + ly);
#pragma unroll
    for (int f = 0; f < Tile<P, JAC>::NEX; ++f)
      if (JAC || L.Mp[f] != 0.0) s.exz[f][ly][lx] = vtx::ldz(EX + (JAC ? P::in(f) : f) * n, m, x0 + lx, y0 + ly);
  }
}

// the five nodes of local square (sx, sy) of one staged field: corners v0..v3, centre
template <class TC, class TZ>
PFH_HD void nodes5(const TC& c, const TZ& z, int sx, int sy, double (&q)[5]) {
  q[0] = c[sy][sx];
  q[1] = c[sy][sx + 1];
  q[2] = c[sy + 1][sx];
  q[3] = c[sy + 1][sx + 1];
  q[4] = z[sy][sx];
}

template <class P, bool JAC>
PFH_HD void square_phase(Tile<P, JAC>& s, const P& prob, const Mesh2& m, const LinCoefs& L, int x0, int y0, int sx, int sy,
                         double* out) {
  const int ox = x0 + sx, oy = y0 + sy;
  const bool valid = ox >= 0 && ox < m.nx && oy >= 0 && oy < m.ny;
  double o5[P::NOUT][5];
#pragma unroll
  for (int o = 0; o < P::NOUT; ++o)
#pragma unroll
    for (int l = 0; l < 5; ++l) o5[o][l] = 0.0;
  if (valid) {
    double u[P::NIN][5], x[P::NIN][5];
#pragma unroll
    for (int f = 0; f < P::NIN; ++f) {
      if (JAC) {
        nodes5(s.exc[f], s.exz[f], sx, sy, u[f]);
        nodes5(s.opc[P::in(f)], s.opz[P::in(f)], sx, sy, x[f]);
      } else {
        nodes5(s.opc[P::in(f)], s.opz[P::in(f)], sx, sy, u[f]);
#pragma unroll
        for (int l = 0; l < 5; ++l) x[f][l] = 0.0;
      }
    }
    elem::square_load<P, JAC>(prob, u, x, m.detJ, o5);
  }
#pragma unroll
  for (int o = 0; o < P::NOUT; ++o)
#pragma unroll
    for (int c = 0; c < 4; ++c) s.cv[o][c][sy][sx] = o5[o][c];
  if (!valid || sx < 1 || sy < 1) return;   // the centre belongs to the CTA that owns the square's lower-left corner
  double Ms[P::NF], Ks[P::NF], row[P::NF];
#pragma unroll
  for (int b = 0; b < P::NF; ++b) {
    double q[5];
    nodes5(s.opc[b], s.opz[b], sx, sy, q);
    vtx::mk_sums2z(m, q, Ms[b], Ks[b]);
  }
#pragma unroll
  for (int a = 0; a < P::NF; ++a) {
    double v = 0.0;
#pragma unroll
    for (int b = 0; b < P::NF; ++b) v += L.Mc[a][b] * Ms[b] + L.Kc[a][b] * Ks[b];
    if (!JAC && L.Mp[a] != 0.0) {
      double q0[5], m0, k0;
      nodes5(s.exc[a], s.exz[a], sx, sy, q0);
      vtx::mk_sums2z(m, q0, m0, k0);
      v -= L.Mp[a] * m0;
    }
    row[a] = v;
  }
#pragma unroll
  for (int o = 0; o < P::NOUT; ++o) row[P::out(o)] += o5[o][4];
  const long long i = m.ncorner + (long long)oy * m.nx + ox;
#pragma unroll
  for (int a = 0; a < P::NF; ++a) out[a * m.n_owned + i] = row[a];
}

// 13-neighbourhood of owned vertex (tx, ty) of one staged field, in vtx::gather13's order
template <class TC, class TZ>
PFH_HD void nodes13(const TC& c, const TZ& z, int tx, int ty, double (&nb)[13]) {
#pragma unroll
  for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
    for (int dx = -1; dx <= 1; ++dx) nb[(dy + 1) * 3 + dx + 1] = c[ty + 1 + dy][tx + 1 + dx];
  nb[9] = z[ty + 1][tx + 1];   // NE: square (ix, iy)
  nb[10] = z[ty + 1][tx];      // NW: square (ix-1, iy)
  nb[11] = z[ty][tx + 1];      // SE: square (ix, iy-1)
  nb[12] = z[ty][tx];          // SW: square (ix-1, iy-1)
}

template <class P, bool JAC>
PFH_HD void vertex_phase(const Tile<P, JAC>& s, const P& prob, const Mesh2& m, const LinCoefs& L, int x0, int y0, int tx, int ty,
                         double* out) {
  if (tx >= WX || ty >= WY) return;
  const int ix = x0 + 1 + tx, iy = y0 + 1 + ty;
  if (ix >= m.nxv || iy >= m.nyv) return;
  vtx::Row2 r;
  vtx::get_row2(m, vtx::type2(m, ix, iy), r);
  double Ms[P::NF], Ks[P::NF], self[P::NF], row[P::NF];
#pragma unroll
  for (int b = 0; b < P::NF; ++b) {
    double nb[13];
    nodes13(s.opc[b], s.opz[b], tx, ty, nb);
    vtx::mk_sums2c(r, nb, Ms[b], Ks[b]);
    self[b] = nb[4];
  }
#pragma unroll
  for (int a = 0; a < P::NF; ++a) {
    double v = 0.0;
#pragma unroll
    for (int b = 0; b < P::NF; ++b) v += L.Mc[a][b] * Ms[b] + L.Kc[a][b] * Ks[b];
    if (!JAC && L.Mp[a] != 0.0) {
      double n0[13], m0, k0;
      nodes13(s.exc[a], s.exz[a], tx, ty, n0);
      vtx::mk_sums2c(r, n0, m0, k0);
      v -= L.Mp[a] * m0;
    }
    row[a] = v;
  }
  // squares NE, NW, SE, SW = local squares (tx+1, ty+1), (tx, ty+1), (tx+1, ty), (tx, ty); the vertex is their corner 0, 1, 2, 3
#pragma unroll
  for (int o = 0; o < P::NOUT; ++o) {
    double v = row[P::out(o)];
    v += s.cv[o][0][ty + 1][tx + 1];
    v += s.cv[o][1][ty + 1][tx];
    v += s.cv[o][2][ty][tx + 1];
    v += s.cv[o][3][ty][tx];
    row[P::out(o)] = v;
  }
  if (P::HAS_BC) {
#pragma unroll
    for (int a = 0; a < P::NF; ++a) {
      double val;
      if (prob.dirichlet(a, ix, iy, m.nx, iy * m.hy, val)) row[a] = JAC ? self[a] : self[a] - val;
    }
  }
  const long long i = (long long)iy * m.nxv + ix;
#pragma unroll
  for (int a = 0; a < P::NF; ++a) out[a * m.n_owned + i] = row[a];
}

}  // namespace op2t
