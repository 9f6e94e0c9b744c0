// Shared-memory staged form of the 2-D scalar operator A = am M + ak K on the crossed mesh (Chebyshev sweeps, lin_apply).
//
// k2_scalar_tiled (pfh_stencil.cu) gathers the 9 + 5 operand values of a corner / centre pair straight from global memory and
// relies on L1 for the reuse between neighbours; under ncu it is latency-bound (long-scoreboard stalls, 45-49 % of the DRAM
// peak).  Here a CTA of 32 x 8 threads first copies the operand values its tile touches -- 34 x 10 corners, 33 x 9 centres,
// 1.24 loads per output instead of 14 per thread -- into shared memory with coalesced loads, and after one barrier every
// thread computes its corner row and its centre row from the staged values; the pointwise operands (r, z_{k-1}) are read
// once, coalesced.  Same arithmetic as k2_scalar_tiled: constant-bank coefficients for interior rows, table rows by boundary
// type otherwise, optional Dirichlet identity rows (bc) and lumped reaction diagonal (sig).
// The phases are __host__ __device__ (tests/emul runs them block by block against the oracle).
#pragma once
#include "pfh_vertex.cuh"

namespace sc2 {

constexpr int TX = 32, TY = 8;                       // outputs per CTA: TX x TY corners and as many centres
constexpr int CX = TX + 2, CY = TY + 2;              // staged corners: (x0-1 .. x0+TX, y0-1 .. y0+TY)
constexpr int ZX = TX + 1, ZY = TY + 1;              // staged centres: (x0-1 .. x0+TX-1, y0-1 .. y0+TY-1)

struct Tile {
  double c[CY][CX];
  double z[ZY][ZX];
};
struct Coefs {
  double cc[PFH_NOFF2];  // am M + ak K, interior corner row (offsets of pfh_internal.h)
  double zc[5];          // centre row
};

PFH_HD void stage(Tile& s, const Mesh2& m, const double* in, int x0, int y0, int tid, int nthreads) {
  for (int k = tid; k < CX * CY; k += nthreads) {
    const int lx = k % CX, ly = k / CX;
    s.c[ly][lx] = vtx::ldc(in, m, x0 - 1 + lx, y0 - 1 + ly);
  }
  for (int k = tid; k < ZX * ZY; k += nthreads) {
    const int lx = k % ZX, ly = k / ZX;
    s.z[ly][lx] = vtx::ldz(in, m, x0 - 1 + lx, y0 - 1 + ly);
  }
}

// MODE 0: out = a r + A in;  MODE 2: out <- in + a (in - out) + b D^-1 (r - A in)   (out_is_zero: z_{k-1} = 0, out write-only)
template <int MODE>
PFH_HD void compute(const Tile& s, const Mesh2& m, double am, double ak, const Coefs& cf, double a, double b, const double* r,
                    double* out, int out_is_zero, int bc, const double* sig, int x0, int y0, int tx, int ty) {
  const int ix = x0 + tx, iy = y0 + ty;
  const int lx = tx + 1, ly = ty + 1;               // position of corner (ix, iy) in the staged corner window
  if (ix < m.nxv && iy < m.nyv) {
    const long long i = (long long)iy * m.nxv + ix;
    const double ri = (MODE == 2 || a != 0.0) ? PFH_LDG(r + i) : 0.0;
    if (bc && (ix == 0 || ix == m.nxv - 1)) {
      out[i] = ri;                                   // Dirichlet identity row, value carried by r
    } else {
      const int type = vtx::type2(m, ix, iy);
      const double self = s.c[ly][lx];
      // centres NE (ix, iy), NW (ix-1, iy), SE (ix, iy-1), SW (ix-1, iy-1): window index = centre index - (x0-1, y0-1)
      const double v[PFH_NOFF2] = {self, s.c[ly][lx + 1], s.c[ly][lx - 1], s.c[ly + 1][lx], s.c[ly - 1][lx],
                                   s.z[ty + 1][tx + 1], s.z[ty + 1][tx], s.z[ty][tx + 1], s.z[ty][tx]};
      double Av = 0.0, d;
      if (type == 4) {
#pragma unroll
        for (int o = 0; o < PFH_NOFF2; ++o) Av += cf.cc[o] * v[o];
        d = cf.cc[0];
      } else {
        vtx::Row2 row;
        vtx::get_row2(m, type, row);
        double Ms = 0.0, Ks = 0.0;
#pragma unroll
        for (int o = 0; o < PFH_NOFF2; ++o) {
          Ms += row.m[o] * v[o];
          Ks += row.k[o] * v[o];
        }
        Av = am * Ms + ak * Ks;
        d = am * row.m[0] + ak * row.k[0];
      }
      if (sig) {
        const double sg = PFH_LDG(sig + i);
        Av += sg * self;
        d += sg;
      }
      if (MODE == 0) out[i] = a * ri + Av;
      else out[i] = self + a * (self - (out_is_zero ? 0.0 : out[i])) + (b / d) * (ri - Av);
    }
  }
  if (ix < m.nx && iy < m.ny) {                      // cell centre (ix, iy): corners v0..v3 = (ix, iy), (ix+1, iy), (ix, iy+1), (ix+1, iy+1)
    const long long ci = m.ncorner + (long long)iy * m.nx + ix;
    const double self = s.z[ty + 1][tx + 1];
    const double sg = sig ? PFH_LDG(sig + ci) : 0.0;
    const double Av = cf.zc[0] * self + cf.zc[1] * s.c[ly][lx] + cf.zc[2] * s.c[ly][lx + 1] + cf.zc[3] * s.c[ly + 1][lx] +
                      cf.zc[4] * s.c[ly + 1][lx + 1] + sg * self;
    const double ri = (MODE == 2 || a != 0.0) ? PFH_LDG(r + ci) : 0.0;
    if (MODE == 0) out[ci] = a * ri + Av;
    else out[ci] = self + a * (self - (out_is_zero ? 0.0 : out[ci])) + (b / (cf.zc[0] + sg)) * (ri - Av);
  }
}

}  // namespace sc2
