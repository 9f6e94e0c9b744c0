// Kernels of the steady quasi-linear diffusion problem on the right-diagonal mesh (dolfin/nl.py; per-vertex / per-square
// arithmetic in pfh_nl.cuh): residual, J.v and the Jacobi-Chebyshev sweep of the preconditioner, all matrix-free (the
// coefficient kc |u| changes with every Newton iterate, nothing could be tabulated).
//
// Two mappings of the same arithmetic:
//  * k_nl_tiled (default): SQUARE-centric, the 2-D form of the cube-centric J.v of pfh_stencil3_tma.cu.  A CTA of 32 x 8
//    threads owns a 31 x 7 tile of vertices: it stages the 33 x 9 operand values around the tile in shared memory, every
//    thread evaluates ONE of the 32 x 8 squares that touch the tile (two triangles, six quadrature points each; one
//    pass, no idle second pass) and leaves its four corner contributions in shared memory; after one barrier
//    every vertex adds up the entries of its four squares.  2.4 triangle evaluations per vertex instead of the 6 of a
//    vertex-centric gather; operand traffic = one read per array + halo.
//  * k_nl_gather: thread -> vertex around nl::vertex_op, the mapping tests/emul runs on the CPU (pfh_set_option
//    "nl_tiled" 0); the tiled kernel is tested against it.
// Algorithmic bytes per vertex: residual 16 (u, F), J.v 24 (u, x, y), sweep 40 (u, z, zprev, r, out).
#include "pfh_nl.cuh"

namespace {

constexpr int NTX = 32, NTY = 8;                                  // gather kernel: vertices per CTA
constexpr int NSX = 32, NSY = 8;                                  // tiled kernel: squares per CTA = threads per CTA
constexpr int NWX = NSX - 1, NWY = NSY - 1, NVX = NSX + 1, NVY = NSY + 1;   // vertices owned / staged per CTA

struct SweepArgs {
  double a, b, sgn;
  const double* r;
  const double* zprev;
};

template <int MODE>
__device__ __forceinline__ double finish(const MeshR2& m, const NlParams& p, int ix, double val, double diag, double self,
                                         long long i, const SweepArgs& s) {
  if (MODE != 2) return val;
  if (nl::is_dirichlet(m, ix)) return s.r[i];
  const double zp = s.a != 0.0 ? s.zprev[i] : self;   // first sweep: z_{k-1} is never read (the buffer may hold anything)
  return self + s.a * (self - zp) + (s.b / diag) * (s.sgn * s.r[i] - val);
}

// MODE 0: out = F(U); 1: out = J(U) X; 2: out = X + a (X - zprev) + b D^-1 (sgn r - A X)   (out may alias zprev)
template <int MODE>
__global__ void __launch_bounds__(NTX* NTY) k_nl_gather(MeshR2 m, NlParams p, const double* __restrict__ U,
                                                        const double* __restrict__ X, SweepArgs s, double* out) {
  const int ix = blockIdx.x * NTX + threadIdx.x, iy = blockIdx.y * NTY + threadIdx.y;
  if (ix >= m.nxv || iy >= m.nyv) return;
  const long long i = (long long)iy * m.nxv + ix;
  double val, diag;
  nl::vertex_op<MODE>(m, p, U, X, ix, iy, val, diag);
  out[i] = finish<MODE>(m, p, ix, val, diag, MODE == 2 ? X[i] : 0.0, i, s);
}

template <int MODE>
__global__ void __launch_bounds__(NSX* NSY) k_nl_tiled(MeshR2 m, NlParams p, const double* __restrict__ U,
                                                       const double* __restrict__ X, SweepArgs s, double* out) {
  __shared__ double us[NVY][NVX];
  __shared__ double xs[MODE == 0 ? 1 : NVY][MODE == 0 ? 1 : NVX];
  __shared__ double cv[4][NSY][NSX];
  __shared__ double cd[MODE == 2 ? 4 : 1][MODE == 2 ? NSY : 1][MODE == 2 ? NSX : 1];
  const int tid = threadIdx.y * NSX + threadIdx.x;
  const int x0 = blockIdx.x * NWX - 1, y0 = blockIdx.y * NWY - 1;  // vertex (and square) origin of the staged window
  for (int k = tid; k < NVX * NVY; k += NSX * NSY) {
    const int lx = k % NVX, ly = k / NVX, gx = x0 + lx, gy = y0 + ly;
    const bool in = gx >= 0 && gx < m.nxv && gy >= 0 && gy < m.nyv;
    const long long g = (long long)gy * m.nxv + gx;
    us[ly][lx] = in ? U[g] : 0.0;
    if (MODE != 0) xs[ly][lx] = in ? X[g] : 0.0;
  }
  __syncthreads();
  {
    const int sx = threadIdx.x, sy = threadIdx.y, ox = x0 + sx, oy = y0 + sy;
    double o[4] = {0.0, 0.0, 0.0, 0.0}, d[4] = {0.0, 0.0, 0.0, 0.0};
    if (ox >= 0 && ox <= m.nx - 1 && oy >= 0 && oy <= m.ny - 1) {
      const double u[4] = {us[sy][sx], us[sy][sx + 1], us[sy + 1][sx], us[sy + 1][sx + 1]};
      double x[4] = {0.0, 0.0, 0.0, 0.0};
      if (MODE != 0) {
        x[0] = xs[sy][sx]; x[1] = xs[sy][sx + 1]; x[2] = xs[sy + 1][sx]; x[3] = xs[sy + 1][sx + 1];
      }
      nl::square_op<MODE>(m, p, u, x, o, d);
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      cv[c][sy][sx] = o[c];
      if (MODE == 2) cd[c][sy][sx] = d[c];
    }
  }
  __syncthreads();
  const int lx = threadIdx.x, ly = threadIdx.y, ix = x0 + 1 + lx, iy = y0 + 1 + ly;
  if (lx >= NWX || ly >= NWY || ix >= m.nxv || iy >= m.nyv) return;
  const long long i = (long long)iy * m.nxv + ix;
  double val, diag = 1.0;
  if (nl::is_dirichlet(m, ix)) {
    val = MODE == 0 ? us[ly + 1][lx + 1] - p.bc : xs[MODE == 0 ? 0 : ly + 1][MODE == 0 ? 0 : lx + 1];
  } else {
    // squares NE, NW, SE, SW of the vertex = local squares (lx+1, ly+1), (lx, ly+1), (lx+1, ly), (lx, ly); our corner index
    // in them is 0, 1, 2, 3 (squares outside the mesh hold zeros)
    val = ((cv[0][ly + 1][lx + 1] + cv[1][ly + 1][lx]) + cv[2][ly][lx + 1]) + cv[3][ly][lx];
    if (MODE == 2) diag = ((cd[0][ly + 1][lx + 1] + cd[1][ly + 1][lx]) + cd[2][ly][lx + 1]) + cd[3][ly][lx];
  }
  out[i] = finish<MODE>(m, p, ix, val, diag, MODE == 2 ? xs[MODE == 0 ? 0 : ly + 1][MODE == 0 ? 0 : lx + 1] : 0.0, i, s);
}

template <int MODE>
void launch(pfh_ctx* c, const double* U, const double* X, const SweepArgs& s, double* out) {
  const MeshR2& m = c->mr;
  if (c->nl_tiled)
    k_nl_tiled<MODE><<<dim3((m.nxv + NWX - 1) / NWX, (m.nyv + NWY - 1) / NWY), dim3(NSX, NSY), 0, c->stream>>>(m, c->nlp, U, X, s, out);
  else
    k_nl_gather<MODE><<<dim3((m.nxv + NTX - 1) / NTX, (m.nyv + NTY - 1) / NTY), dim3(NTX, NTY), 0, c->stream>>>(m, c->nlp, U, X, s, out);
  PFH_CUDA(cudaGetLastError());
}

}  // namespace

void k_nl_residual(pfh_ctx* c, const double* u, double* R) {
  ProfScope ps(c, T_RESIDUAL, 16.0 * c->n_owned);
  launch<0>(c, u, nullptr, SweepArgs{0, 0, 0, nullptr, nullptr}, R);
}

void k_nl_jac(pfh_ctx* c, const double* ulin, const double* x, double* y) {
  ProfScope ps(c, T_JAC_APPLY, 24.0 * c->n_owned);
  launch<1>(c, ulin, x, SweepArgs{0, 0, 0, nullptr, nullptr}, y);
}

void k_nl_sweep(pfh_ctx* c, const double* ulin, double a, double b, double sgn, const double* r, const double* z, double* zprev) {
  ProfScope ps(c, T_CHEB_SWEEP, 40.0 * c->n_owned);
  launch<2>(c, ulin, z, SweepArgs{a, b, sgn, r, zprev}, zprev);
}
