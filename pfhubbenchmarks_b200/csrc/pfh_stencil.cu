// Matrix-free stencil kernels (sm_100a, fp64): residual, Jacobian-vector product, the scalar
// SPD operators of the preconditioner (Jacobi-Chebyshev sweeps on am M + ak K), the energy /
// solute reductions and the initial conditions.  These replace FFC's generated tabulate_tensor
// kernels + DOLFIN's assembler + PETSc MatMult on the reference's hot path
// (dolfin/bench1.py:162 -> SNESSolve; SURVEY.md section 3.3, hot loops A-D).
//
// This file holds the general (any problem, any mesh size) gather kernels and the launch dispatch.  The 3-D fast
// path -- staged scalar operators (pfh_scalar3_tma.cu, any problem) and the cube-centric J.v / residual of
// Cahn-Hilliard (pfh_stencil3_tma.cu), both with a TMA and a cp.async loader -- is tried first; the kernels here serve
// the 2-D problems, generated models in 3-D, the boundary shells of the fast path (vtx::vertex_op3) and are the reference
// the fast path is tested against (pfh_set_option "use_tma" 0).
// Thread mapping: 3-D kernels use (BX x BY) tiles of an x-y plane and march along z so the three planes a vertex
// reads stay in L1; 2-D kernels use 64 x 4 tiles with one corner vertex and one cell centre per thread.
#include "pfh_vertex.cuh"
#include "pfh_op2_tiled.cuh"
#include "pfh_mg2.cuh"

namespace {

constexpr int BX = 64, BY = 4, ZCHUNK = 16;
constexpr int TPB = 256;

template <int NF>
struct Fields {
  FieldRef f[NF];
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}
// block-wide sum of NV values; result valid in thread 0
template <int NV>
__device__ __forceinline__ void block_sum(double (&v)[NV]) {
  __shared__ double sh[NV][32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    v[k] = warp_sum(v[k]);
    if (lane == 0) sh[k][wid] = v[k];
  }
  __syncthreads();
  if (wid == 0) {
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      double t = lane < nw ? sh[k][lane] : 0.0;
      v[k] = warp_sum(t);
    }
  }
}

// ------------------------------- 3-D kernels -------------------------------
template <class P, bool JAC>
__global__ void __launch_bounds__(BX* BY) k3_op(P prob, Mesh3 m, LinCoefs L, Fields<P::NF> U, Fields<P::NF> X,
                                                 Fields<P::NF> U0, double* __restrict__ out) {
  const int ix = blockIdx.x * BX + threadIdx.x, iy = blockIdx.y * BY + threadIdx.y;
  if (ix >= m.nxv || iy >= m.nyv) return;
  const int z_end = min(m.nzl, (int)(blockIdx.z + 1) * ZCHUNK);
  for (int zl = blockIdx.z * ZCHUNK; zl < z_end; ++zl) {
    double o[P::NF];
    vtx::vertex_op3<P, JAC>(prob, m, L, U.f, X.f, U0.f, ix, iy, zl, o);
    const long long i = (long long)zl * m.nxy + (long long)iy * m.nxv + ix;
#pragma unroll
    for (int a = 0; a < P::NF; ++a) out[a * m.n_owned + i] = o[a];
  }
}

// MODE 0: out = s*add + A in ; MODE 1: z = b Dinv r ; MODE 2: zprev <- z + a (z - zprev) + b Dinv (r - A z)
template <int MODE>
__global__ void __launch_bounds__(BX* BY) k3_scalar(Mesh3 m, double am, double ak, double a, double b, FieldRef in,
                                                    const double* __restrict__ r, double* __restrict__ out, bool out_is_zero) {
  const int ix = blockIdx.x * BX + threadIdx.x, iy = blockIdx.y * BY + threadIdx.y;
  if (ix >= m.nxv || iy >= m.nyv) return;
  const int z_end = min(m.nzl, (int)(blockIdx.z + 1) * ZCHUNK);
  for (int zl = blockIdx.z * ZCHUNK; zl < z_end; ++zl) {
    const long long i = (long long)zl * m.nxy + (long long)iy * m.nxv + ix;
    if (MODE == 1) {
      out[i] = b / vtx::scalar_diag3(m, am, ak, ix, iy, zl) * r[i];
    } else {
      double Av, d, self;
      vtx::scalar_op3(m, am, ak, in, ix, iy, zl, Av, d, self);
      if (MODE == 0) out[i] = (a != 0.0 ? a * r[i] : 0.0) + Av;
      else out[i] = self + a * (self - (out_is_zero ? 0.0 : out[i])) + (b / d) * (r[i] - Av);
    }
  }
}

template <class P>
__global__ void __launch_bounds__(TPB) k3_energy(P prob, Mesh3 m, Fields<P::NF> U, double* __restrict__ partial) {
  double acc[2] = {0.0, 0.0};
  for (long long i = blockIdx.x * (long long)TPB + threadIdx.x; i < m.n_owned; i += (long long)gridDim.x * TPB) {
    const int zl = (int)(i / m.nxy);
    const int rem = (int)(i - (long long)zl * m.nxy);
    double e, mm;
    vtx::vertex_energy3<P>(prob, m, U.f, rem % m.nxv, rem / m.nxv, zl, e, mm);
    acc[0] += e;
    acc[1] += mm;
  }
  block_sum<2>(acc);
  if (threadIdx.x == 0) {
    partial[blockIdx.x * PFH_RED_MAXVALS + 0] = acc[0];
    partial[blockIdx.x * PFH_RED_MAXVALS + 1] = acc[1];
  }
}

template <class P>
__global__ void __launch_bounds__(TPB) k3_ic(P prob, Mesh3 m, const double* __restrict__ q, double* __restrict__ w) {
  double qq[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) qq[k] = q[k];
  for (long long i = blockIdx.x * (long long)TPB + threadIdx.x; i < m.n_owned; i += (long long)gridDim.x * TPB) {
    const int rem = (int)(i % m.nxy);
    double u[P::NF];
    prob.ic(qq, (rem % m.nxv) * m.hx, (rem / m.nxv) * m.hy, u);
#pragma unroll
    for (int a = 0; a < P::NF; ++a) w[a * m.n_owned + i] = u[a];
  }
}

// ------------------------------- 2-D kernels -------------------------------
template <class P, bool JAC>
__global__ void __launch_bounds__(TPB) k2_op(P prob, Mesh2 m, LinCoefs L, const double* __restrict__ U,
                                             const double* __restrict__ X, const double* __restrict__ U0,
                                             double* __restrict__ out) {
  // 2-D tiles (64 x 4): thread (ix, iy) evaluates corner vertex (ix, iy) and cell centre (ix, iy)
  const int ix = blockIdx.x * 64 + (threadIdx.x & 63), iy = blockIdx.y * 4 + (threadIdx.x >> 6);
  double o[P::NF];
  if (ix < m.nxv && iy < m.nyv) {
    vtx::vertex_op2_corner<P, JAC>(prob, m, L, U, X, U0, ix, iy, o);
    const long long i = (long long)iy * m.nxv + ix;
#pragma unroll
    for (int a = 0; a < P::NF; ++a) out[a * m.n_owned + i] = o[a];
  }
  if (ix < m.nx && iy < m.ny) {
    vtx::vertex_op2_centre<P, JAC>(prob, m, L, U, X, U0, ix, iy, o);
    const long long i = m.ncorner + (long long)iy * m.nx + ix;
#pragma unroll
    for (int a = 0; a < P::NF; ++a) out[a * m.n_owned + i] = o[a];
  }
}

// square-centric form of k2_op (pfh_op2_tiled.cuh): every crossed square is evaluated once per launch
template <class P, bool JAC>
__global__ void __launch_bounds__(op2t::SX* op2t::SY) k2_op_tiled(P prob, Mesh2 m, LinCoefs L, const double* __restrict__ OP,
                                                                   const double* __restrict__ EX, double* __restrict__ out) {
  __shared__ op2t::Tile<P, JAC> s;
  const int x0 = blockIdx.x * op2t::WX - 1, y0 = blockIdx.y * op2t::WY - 1;
  op2t::stage<P, JAC>(s, m, L, OP, EX, x0, y0, threadIdx.y * op2t::SX + threadIdx.x, op2t::SX * op2t::SY);
  __syncthreads();
  op2t::square_phase<P, JAC>(s, prob, m, L, x0, y0, threadIdx.x, threadIdx.y, out);
  __syncthreads();
  op2t::vertex_phase<P, JAC>(s, prob, m, L, x0, y0, threadIdx.x, threadIdx.y, out);
}
template <class P, bool JAC>
bool launch_op2_tiled(pfh_ctx* c, const LinCoefs& L, const double* op, const double* ex, double* out) {
  if constexpr (op2t::fits<P, JAC>()) {
    const Mesh2& m = c->m2;
    k2_op_tiled<P, JAC><<<dim3((m.nxv + op2t::WX - 1) / op2t::WX, (m.nyv + op2t::WY - 1) / op2t::WY), dim3(op2t::SX, op2t::SY), 0,
                          c->stream>>>(P(c->prob), m, L, op, ex, out);
    return true;
  } else {
    return false;   // the staged tile of this problem (bench2: six fields) exceeds the static shared memory: gather kernel
  }
}

template <int MODE>
__global__ void __launch_bounds__(TPB) k2_scalar(Mesh2 m, double am, double ak, double a, double b,
                                                 const double* __restrict__ in, const double* __restrict__ r,
                                                 double* __restrict__ out, bool out_is_zero, int bc,
                                                 const double* __restrict__ sig) {
  const long long i = blockIdx.x * (long long)TPB + threadIdx.x;
  if (i >= m.n_owned) return;
  if (bc && i < m.ncorner) {  // Dirichlet rows (bench6 phi, bench6.py:77-90): identity row, value carried by r
    const int ix = (int)(i % m.nxv);
    if (ix == 0 || ix == m.nxv - 1) {
      out[i] = r[i];
      return;
    }
  }
  const double sg = sig ? sig[i] : 0.0;  // lumped reaction term added to the diagonal
  if (MODE == 1) {
    out[i] = b / (vtx::scalar_diag2(m, am, ak, i) + sg) * r[i];
  } else {
    double Av, d, self;
    vtx::scalar_op2(m, am, ak, in, i, Av, d, self);
    Av += sg * self;
    d += sg;
    if (MODE == 0) out[i] = (a != 0.0 ? a * r[i] : 0.0) + Av;
    else out[i] = self + a * (self - (out_is_zero ? 0.0 : out[i])) + (b / d) * (r[i] - Av);
  }
}

// Tiled 2-D scalar operator: one corner vertex and one cell centre per thread, (T2X x T2Y) tiles so that the rows a
// tile shares stay in L1; interior rows use constant-bank coefficients, boundary rows the stencil tables.
// MODE 0: out = a*r + A in ; MODE 2: out <- in + a (in - out) + b D^-1 (r - A in).  32-bit indices.
constexpr int T2X = 64, T2Y = 4;
struct Coef2 {
  double cc[PFH_NOFF2];  // am*M + ak*K, interior corner row
  double zc[5];          // am*M + ak*K, centre row
};
template <int MODE>
__global__ void __launch_bounds__(T2X* T2Y, 8) k2_scalar_tiled(Mesh2 m, double am, double ak, Coef2 cf, double a, double b,
                                                             const double* __restrict__ in, const double* __restrict__ r,
                                                             double* __restrict__ out, int out_is_zero, int bc,
                                                             const double* __restrict__ sig) {
  const int ix = blockIdx.x * T2X + threadIdx.x, iy = blockIdx.y * T2Y + threadIdx.y;
  const int nxv = m.nxv, nx = m.nx, nc = (int)m.ncorner;
  if (ix < nxv && iy < m.nyv) {  // corner vertex (ix, iy)
    const int i = iy * nxv + ix;
    const double ri = (MODE == 2 || a != 0.0) ? __ldg(r + i) : 0.0;
    if (bc && (ix == 0 || ix == nxv - 1)) {
      out[i] = ri;  // Dirichlet identity row
    } else {
      const int type = vtx::type2(m, ix, iy);
      double Av, d, self;
      if (type == 4) {
        const int cb = nc + iy * nx + ix;
        self = __ldg(in + i);
        Av = cf.cc[0] * self + cf.cc[1] * __ldg(in + i + 1) + cf.cc[2] * __ldg(in + i - 1) + cf.cc[3] * __ldg(in + i + nxv) +
             cf.cc[4] * __ldg(in + i - nxv) + cf.cc[5] * __ldg(in + cb) + cf.cc[6] * __ldg(in + cb - 1) +
             cf.cc[7] * __ldg(in + cb - nx) + cf.cc[8] * __ldg(in + cb - nx - 1);
        d = cf.cc[0];
      } else {
        vtx::Row2 row;
        vtx::get_row2(m, type, row);
        double nb[13], Ms, Ks;
        vtx::gather13(in, m, ix, iy, nb);
        vtx::mk_sums2c(row, nb, Ms, Ks);
        self = nb[4];
        Av = am * Ms + ak * Ks;
        d = am * row.m[0] + ak * row.k[0];
      }
      if (sig) {
        const double sg = sig[i];
        Av += sg * self;
        d += sg;
      }
      if (MODE == 0) out[i] = a * ri + Av;
      else out[i] = self + a * (self - (out_is_zero ? 0.0 : out[i])) + (b / d) * (ri - Av);
    }
  }
  if (ix < nx && iy < m.ny) {  // cell centre (ix, iy)
    const int ci = nc + iy * nx + ix, v0 = iy * nxv + ix;
    const double self = __ldg(in + ci);
    const double sg = sig ? sig[ci] : 0.0;
    const double Av = cf.zc[0] * self + cf.zc[1] * __ldg(in + v0) + cf.zc[2] * __ldg(in + v0 + 1) + cf.zc[3] * __ldg(in + v0 + nxv) +
                      cf.zc[4] * __ldg(in + v0 + nxv + 1) + sg * self;
    const double ri = (MODE == 2 || a != 0.0) ? __ldg(r + ci) : 0.0;
    if (MODE == 0) out[ci] = a * ri + Av;
    else out[ci] = self + a * (self - (out_is_zero ? 0.0 : out[ci])) + (b / (cf.zc[0] + sg)) * (ri - Av);
  }
}

// multigrid transfers on the crossed mesh (pfh_mg2.cuh): one coarse / fine node per thread, gathers only
__global__ void __launch_bounds__(TPB) k2_mg_restrict_kernel(Mesh2 mf, Mesh2 mc, const double* __restrict__ t, double scale,
                                                             double* __restrict__ b, int bc) {
  const long long i = blockIdx.x * (long long)TPB + threadIdx.x;
  if (i >= mc.n_owned) return;
  b[i] = (bc && mg2::is_dirichlet(mc, i)) ? 0.0 : scale * mg2::restrict_node(mf, mc, t, i);
}
__global__ void __launch_bounds__(TPB) k2_mg_prolong_kernel(Mesh2 mf, Mesh2 mc, const double* __restrict__ xc, double* __restrict__ xf,
                                                            int bc) {
  const long long i = blockIdx.x * (long long)TPB + threadIdx.x;
  if (i >= mf.n_owned || (bc && mg2::is_dirichlet(mf, i))) return;
  xf[i] += mg2::prolong_node(mf, mc, xc, i);
}

template <class P>
__global__ void __launch_bounds__(TPB) k2_energy(P prob, Mesh2 m, const double* __restrict__ U, double* __restrict__ partial) {
  double acc[2] = {0.0, 0.0};
  for (long long i = blockIdx.x * (long long)TPB + threadIdx.x; i < m.n_owned; i += (long long)gridDim.x * TPB) {
    double e, mm;
    vtx::vertex_energy2<P>(prob, m, U, i, e, mm);
    acc[0] += e;
    acc[1] += mm;
  }
  block_sum<2>(acc);
  if (threadIdx.x == 0) {
    partial[blockIdx.x * PFH_RED_MAXVALS + 0] = acc[0];
    partial[blockIdx.x * PFH_RED_MAXVALS + 1] = acc[1];
  }
}

template <class P>
__global__ void __launch_bounds__(TPB) k2_ic(P prob, Mesh2 m, const double* __restrict__ q, double* __restrict__ w) {
  const long long i = blockIdx.x * (long long)TPB + threadIdx.x;
  if (i >= m.n_owned) return;
  double qq[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) qq[k] = q[k];
  double x, y;
  if (i < m.ncorner) {
    x = (i % m.nxv) * m.hx;
    y = (i / m.nxv) * m.hy;
  } else {
    const long long ci = i - m.ncorner;
    x = (ci % m.nx + 0.5) * m.hx;
    y = (ci / m.nx + 0.5) * m.hy;
  }
  double u[P::NF];
  prob.ic(qq, x, y, u);
#pragma unroll
  for (int a = 0; a < P::NF; ++a) w[a * m.n_owned + i] = u[a];
}

// z_a = r_a / (Mc_aa M_ii + Kc_aa K_ii): point Jacobi on the linear part of the diagonal blocks
template <int DIM>
__global__ void __launch_bounds__(TPB) k_jacobi_kernel(Mesh2 m2, Mesh3 m3, LinCoefs L, int nf, const double* __restrict__ r,
                                                       double* __restrict__ z) {
  const long long n = DIM == 2 ? m2.n_owned : m3.n_owned;
  const long long i = blockIdx.x * (long long)TPB + threadIdx.x;
  if (i >= n) return;
  for (int a = 0; a < nf; ++a) {
    double d;
    if (DIM == 2) {
      d = vtx::scalar_diag2(m2, L.Mc[a][a], L.Kc[a][a], i);
    } else {
      const int zl = (int)(i / m3.nxy);
      const int rem = (int)(i - (long long)zl * m3.nxy);
      d = vtx::scalar_diag3(m3, L.Mc[a][a], L.Kc[a][a], rem % m3.nxv, rem / m3.nxv, zl);
    }
    z[a * n + i] = r[a * n + i] / (d != 0.0 ? d : 1.0);
  }
}

template <class P>
__global__ void __launch_bounds__(TPB) k2_dirichlet(P prob, Mesh2 m, double* __restrict__ w) {
  const long long i = blockIdx.x * (long long)TPB + threadIdx.x;
  if (i >= m.ncorner) return;
  const int ix = (int)(i % m.nxv), iy = (int)(i / m.nxv);
#pragma unroll
  for (int a = 0; a < P::NF; ++a) {
    double val;
    if (prob.dirichlet(a, ix, iy, m.nx, iy * m.hy, val)) w[a * m.n_owned + i] = val;
  }
}

__global__ void __launch_bounds__(TPB) k2_dirichlet_copy(Mesh2 m, const double* __restrict__ src, double* __restrict__ dst) {
  const long long i = blockIdx.x * (long long)TPB + threadIdx.x;
  if (i >= m.ncorner) return;
  const int ix = (int)(i % m.nxv);
  if (ix == 0 || ix == m.nxv - 1) dst[i] = src[i];
}

// the scalar operator on the owned planes of a slab through the staged kernel (pfh_scalar3_tma.cu): `in` with its
// one-plane ghost arrays
void scalar_owned(pfh_ctx* c, int mode, double am, double ak, double a, double b, double b0, FieldRef in, const double* r,
                  const double* o, double* out, double* out1, double ob0 = 0.0) {
  const Mesh3& m = c->m3;
  ScalarOp op;
  op.mode = mode; op.am = am; op.ak = ak; op.a = a; op.b = b; op.b0 = b0; op.ob0 = ob0;
  op.in = in.p; op.in_off = 0; op.in_n = m.nzl;
  if (in.glo != nullptr && c->cfg.rank > 0) { op.in_lo = in.glo; op.n_lo = 1; }
  if (in.ghi != nullptr && c->cfg.rank < c->cfg.nranks - 1) { op.in_hi = in.ghi; op.n_hi = 1; }
  op.r = r; op.o = o; op.out = out; op.out1 = out1;
  op.gz0 = m.z0; op.nzo = m.nzl;
  k3_scalar_launch(c, op);
}

// sig_i = clip(d g_a / d u_a (u_i)) * rowsum(M)_i : lumped reaction term of the diagonal Jacobian block of field a
// (the M_{g'} block is replaced by its row-sum lumping at the vertex value; clipped so that the block stays positive)
template <class P>
__global__ void __launch_bounds__(TPB) k2_reaction_diag(P prob, Mesh2 m, const double* __restrict__ U, int field, double floor_coef,
                                                         double* __restrict__ sig) {
  const long long i = blockIdx.x * (long long)TPB + threadIdx.x;
  if (i >= m.n_owned) return;
  double u[P::NIN], x[P::NIN], o[P::NOUT];
  int oi = -1;
#pragma unroll
  for (int f = 0; f < P::NIN; ++f) {
    u[f] = U[P::in(f) * m.n_owned + i];
    x[f] = P::in(f) == field ? 1.0 : 0.0;
  }
#pragma unroll
  for (int k = 0; k < P::NOUT; ++k)
    if (P::out(k) == field) oi = k;
  double s = 0.0;
  if (oi >= 0) {
    prob.dg(u, x, o);
    s = o[oi];
  }
  // row sum of M at this vertex
  double ml;
  if (i < m.ncorner) {
    vtx::Row2 row;
    vtx::get_row2(m, vtx::type2(m, (int)(i % m.nxv), (int)(i / m.nxv)), row);
    ml = 0.0;
#pragma unroll
    for (int k = 0; k < PFH_NOFF2; ++k) ml += row.m[k];
  } else {
    ml = m.mC[0] + m.mC[1] + m.mC[2] + m.mC[3] + m.mC[4];
  }
  sig[i] = fmax(s, floor_coef) * ml;
}

// -------- launch helpers --------
Coef2 coef2(const Mesh2& m, double am, double ak) {
  Coef2 cf;
  for (int o = 0; o < PFH_NOFF2; ++o) cf.cc[o] = am * m.mI[o] + ak * m.kI[o];
  for (int o = 0; o < 5; ++o) cf.zc[o] = am * m.mC[o] + ak * m.kC[o];
  return cf;
}
dim3 grid2_tiled(const Mesh2& m) { return dim3((m.nxv + T2X - 1) / T2X, (m.nyv + T2Y - 1) / T2Y); }

dim3 grid3(const Mesh3& m) { return dim3((m.nxv + BX - 1) / BX, (m.nyv + BY - 1) / BY, (m.nzl + ZCHUNK - 1) / ZCHUNK); }
int grid1(long long n) { return (int)((n + TPB - 1) / TPB); }
int grid_red(long long n) {
  long long g = (n + TPB - 1) / TPB;
  return (int)(g < PFH_RED_MAXBLOCKS ? g : PFH_RED_MAXBLOCKS);
}

template <class P>
Fields<P::NF> state_fields(const pfh_ctx* c, const double* u, const double* glo, const double* ghi) {
  Fields<P::NF> F;
  for (int f = 0; f < P::NF; ++f) {
    F.f[f].p = u ? u + f * c->n_owned : nullptr;
    F.f[f].glo = (glo && c->nplane) ? glo + f * c->nplane : nullptr;
    F.f[f].ghi = (ghi && c->nplane) ? ghi + f * c->nplane : nullptr;
  }
  return F;
}

// A library generated for one user model (PFH_USER_PROBLEM_HEADER, codegen.py) instantiates the kernels for that
// model in addition to Cahn-Hilliard (whose TMA kernels are compiled in any case); PFH_ONLY_USER drops the other
// shipped systems from it to keep the just-in-time build short.
#ifdef PFH_USER_PROBLEM_HEADER
#define PFH_USER_CASE(BODY) case PFH_PROBLEM_USER: { using P = ProbUser; BODY; } break;
#else
#define PFH_USER_CASE(BODY)
#endif
#ifdef PFH_ONLY_USER
#define PFH_SHIPPED_CASES(BODY)
#else
#define PFH_SHIPPED_CASES(BODY)                                           \
    case PFH_PROBLEM_BENCH2: { using P = ProbBench2; BODY; } break;       \
    case PFH_PROBLEM_BENCH3: { using P = ProbBench3; BODY; } break;       \
    case PFH_PROBLEM_BENCH6: { using P = ProbBench6; BODY; } break;
#endif
#define PFH_DISPATCH2(c, BODY)                                            \
  switch ((c)->prob.id) {                                                 \
    case PFH_PROBLEM_CH: { using P = ProbCH; BODY; } break;               \
    PFH_SHIPPED_CASES(BODY)                                               \
    PFH_USER_CASE(BODY)                                                   \
    default: throw PfhError("unknown problem id");                        \
  }
#define PFH_DISPATCH3(c, BODY)                                                     \
  switch ((c)->prob.id) {                                                          \
    case PFH_PROBLEM_CH: { using P = ProbCH; BODY; } break;                        \
    PFH_USER_CASE(BODY)                                                            \
    default: throw PfhError("3-D meshes support PFH_PROBLEM_CH (b13d) and generated models only"); \
  }

template <class P>
LinCoefs lin_of(const pfh_ctx* c, double dt) {
  LinCoefs L;
  memset(&L, 0, sizeof L);
  P(c->prob).lin(dt, L);
  return L;
}

}  // namespace

LinCoefs pfh_lin_coefs(const pfh_ctx* c, double dt) {
  LinCoefs L;
  PFH_DISPATCH2(c, L = lin_of<P>(c, dt));
  return L;
}

void k_residual(pfh_ctx* c, const double* u, const double* u0, double* R) {
  if (pfh_is_nl(c)) return k_nl_residual(c, u, R);   // steady: no previous step
  const LinCoefs L = pfh_lin_coefs(c, c->dt);
  int nprev = 0;
  for (int a = 0; a < c->nf; ++a) nprev += L.Mp[a] != 0.0;
  ProfScope ps(c, T_RESIDUAL, 8.0 * (2 * c->nf + nprev) * c->n_owned);
  if (c->dim == 3 && k3_res_tma_launch(c, L, u, u0, R)) return;
  if (c->dim == 3) {
    PFH_DISPATCH3(c, (k3_op<P, false><<<grid3(c->m3), dim3(BX, BY), 0, c->stream>>>(
                         P(c->prob), c->m3, L, state_fields<P>(c, u, c->gU_lo, c->gU_hi), state_fields<P>(c, nullptr, nullptr, nullptr),
                         state_fields<P>(c, u0, c->gU0_lo, c->gU0_hi), R)));
  } else {
    bool done = false;
    if (c->op2_tiled) PFH_DISPATCH2(c, (done = launch_op2_tiled<P, false>(c, L, u, u0, R)));
    if (!done)
      PFH_DISPATCH2(c, (k2_op<P, false><<<dim3((c->m2.nxv + 63) / 64, (c->m2.nyv + 3) / 4), TPB, 0, c->stream>>>(P(c->prob), c->m2, L, u, nullptr, u0, R)));
  }
  PFH_CUDA(cudaGetLastError());
}

void k_jac_apply(pfh_ctx* c, const double* ulin, const double* x, double* y) {
  if (pfh_is_nl(c)) return k_nl_jac(c, ulin, x, y);
  const LinCoefs L = pfh_lin_coefs(c, c->dt);
  int nin = c->prob.id == PFH_PROBLEM_BENCH2 ? 5 : (c->prob.id == PFH_PROBLEM_BENCH3 ? 2 : 1);
#ifdef PFH_USER_PROBLEM_HEADER
  if (c->prob.id == PFH_PROBLEM_USER) nin = ProbUser::NIN;
#endif
  ProfScope ps(c, T_JAC_APPLY, 8.0 * (2 * c->nf + nin) * c->n_owned);
  if (c->dim == 3 && k3_jac_tma_launch(c, L, ulin, x, y)) return;
  if (c->dim == 3) {
    PFH_DISPATCH3(c, (k3_op<P, true><<<grid3(c->m3), dim3(BX, BY), 0, c->stream>>>(
                         P(c->prob), c->m3, L, state_fields<P>(c, ulin, c->gU_lo, c->gU_hi), state_fields<P>(c, x, c->gX_lo, c->gX_hi),
                         state_fields<P>(c, nullptr, nullptr, nullptr), y)));
  } else {
    bool done = false;
    if (c->op2_tiled) PFH_DISPATCH2(c, (done = launch_op2_tiled<P, true>(c, L, x, ulin, y)));
    if (!done)
      PFH_DISPATCH2(c, (k2_op<P, true><<<dim3((c->m2.nxv + 63) / 64, (c->m2.nyv + 3) / 4), TPB, 0, c->stream>>>(P(c->prob), c->m2, L, ulin, x, nullptr, y)));
  }
  PFH_CUDA(cudaGetLastError());
}

bool k_ch_crow(pfh_ctx* c, const double* ulin, FieldRef xc, double* out) {
  if (!pfh_fast3_ok(c) || c->prob.id != PFH_PROBLEM_CH) return false;
  const LinCoefs L = pfh_lin_coefs(c, c->dt);
  ProfScope ps(c, T_JAC_APPLY, 24.0 * c->n_owned);
  return k3_ch_crow_tma_launch(c, L, ulin, xc, out);
}

void k_lin_apply(pfh_ctx* c, double am, double ak, FieldRef in, double s, const double* add, double* out) {
  ProfScope ps(c, T_LIN_APPLY, 8.0 * (s != 0.0 ? 3 : 2) * c->n_owned);
  PFH_CHECK(s == 0.0 || add != nullptr, "k_lin_apply: add operand missing");
  if (pfh_fast3_ok(c)) {
    scalar_owned(c, 0, am, ak, s, 0.0, 0.0, in, add, nullptr, out, nullptr);
    return;
  }
  if (c->dim == 3)
    k3_scalar<0><<<grid3(c->m3), dim3(BX, BY), 0, c->stream>>>(c->m3, am, ak, s, 0.0, in, add, out, false);
  else
    k2_scalar_tiled<0><<<grid2_tiled(c->m2), dim3(T2X, T2Y), 0, c->stream>>>(c->m2, am, ak, coef2(c->m2, am, ak), s, 0.0, in.p, add, out, 0,
                                                                                  c->scalar_bc, c->scalar_sig);
  PFH_CUDA(cudaGetLastError());
}

void k_cheb_first(pfh_ctx* c, double am, double ak, double b, const double* r, double* z) {
  ProfScope ps(c, T_CHEB_FIRST, 16.0 * c->n_owned);
  if (c->dim == 3)
    k3_scalar<1><<<grid3(c->m3), dim3(BX, BY), 0, c->stream>>>(c->m3, am, ak, 0.0, b, FieldRef{nullptr, nullptr, nullptr}, r, z, false);
  else
    k2_scalar<1><<<grid1(c->n_owned), TPB, 0, c->stream>>>(c->m2, am, ak, 0.0, b, nullptr, r, z, false, c->scalar_bc, c->scalar_sig);
  PFH_CUDA(cudaGetLastError());
}

bool k_cheb_first2(pfh_ctx* c, double am, double ak, double b0, double a1, double b1, FieldRef r, double* z1, double* z2) {
  if (!pfh_fast3_ok(c)) return false;
  ProfScope ps(c, T_CHEB_FIRST, (z1 ? 24.0 : 16.0) * c->n_owned);
  scalar_owned(c, 3, am, ak, a1, b1, b0, r, nullptr, nullptr, z2, z1);
  return true;
}

void k_cheb_sweep(pfh_ctx* c, double am, double ak, double a, double b, const double* r, FieldRef z, double* zprev,
                  bool zprev_is_zero, double zprev_b0) {
  ProfScope ps(c, T_CHEB_SWEEP, (zprev_is_zero ? 24.0 : 32.0) * c->n_owned);
  if (pfh_fast3_ok(c)) {
    scalar_owned(c, 2, am, ak, a, b, 0.0, z, r, zprev_is_zero ? nullptr : zprev, zprev, nullptr, zprev_is_zero ? zprev_b0 : 0.0);
    return;
  }
  PFH_CHECK(zprev_b0 == 0.0, "internal: recomputed first iterate needs the staged 3-D kernel");
  if (c->dim == 3)
    k3_scalar<2><<<grid3(c->m3), dim3(BX, BY), 0, c->stream>>>(c->m3, am, ak, a, b, z, r, zprev, zprev_is_zero);
  else
    k2_scalar_tiled<2><<<grid2_tiled(c->m2), dim3(T2X, T2Y), 0, c->stream>>>(c->m2, am, ak, coef2(c->m2, am, ak), a, b, z.p, r, zprev,
                                                                                  zprev_is_zero ? 1 : 0, c->scalar_bc, c->scalar_sig);
  PFH_CUDA(cudaGetLastError());
}

void k_jacobi(pfh_ctx* c, const double* ulin, const double* r, double* z) {
  if (pfh_is_nl(c)) {  // z = -D^-1 r with D = diag(K_{kc|u|} + reac M): one sweep from z = 0
    v_zero(c, c->S[0], c->n_owned);
    return k_nl_sweep(c, ulin, 0.0, 1.0, -1.0, r, c->S[0], z);
  }
  ProfScope ps(c, T_JACOBI, 16.0 * c->nf * c->n_owned);
  const LinCoefs L = pfh_lin_coefs(c, c->dt);
  if (c->dim == 3)
    k_jacobi_kernel<3><<<grid1(c->n_owned), TPB, 0, c->stream>>>(c->m2, c->m3, L, c->nf, r, z);
  else
    k_jacobi_kernel<2><<<grid1(c->n_owned), TPB, 0, c->stream>>>(c->m2, c->m3, L, c->nf, r, z);
  PFH_CUDA(cudaGetLastError());
}

void k_energy(pfh_ctx* c, const double* u, double out[4]) {
  const int nb = grid_red(c->n_owned);
  {
    ProfScope ps(c, T_ENERGY, 8.0 * c->nf * c->n_owned);
    if (c->dim == 3) {
      PFH_DISPATCH3(c, (k3_energy<P><<<nb, TPB, 0, c->stream>>>(P(c->prob), c->m3, state_fields<P>(c, u, c->gU_lo, c->gU_hi), c->d_partial)));
    } else {
      PFH_DISPATCH2(c, (k2_energy<P><<<nb, TPB, 0, c->stream>>>(P(c->prob), c->m2, u, c->d_partial)));
    }
    PFH_CUDA(cudaGetLastError());
  }
  const double* h = pfh_reduce_finish(c, nb, 2);
  out[0] = h[0];
  out[1] = h[1];
  out[2] = out[3] = 0.0;
}

void k_initial_condition(pfh_ctx* c, const double* icp, int n, double* w) {
  PFH_CHECK(n <= 8, "too many initial-condition parameters");
  double q[8] = {0};
  for (int i = 0; i < n; ++i) q[i] = icp[i];
  // small parameter block travels through the reduction scratch buffer
  PFH_CUDA(cudaMemcpyAsync(c->d_red, q, sizeof q, cudaMemcpyHostToDevice, c->stream));
  ProfScope ps(c, T_IC);
  if (c->dim == 3) {
    PFH_DISPATCH3(c, (k3_ic<P><<<grid_red(c->n_owned), TPB, 0, c->stream>>>(P(c->prob), c->m3, c->d_red, w)));
  } else {
    PFH_DISPATCH2(c, (k2_ic<P><<<grid1(c->n_owned), TPB, 0, c->stream>>>(P(c->prob), c->m2, c->d_red, w)));
  }
  PFH_CUDA(cudaGetLastError());
  PFH_CUDA(cudaStreamSynchronize(c->stream));  // q is a stack buffer
}

void k_dirichlet_fix(pfh_ctx* c, double* w) {
  if (c->dim != 2 || c->prob.id != PFH_PROBLEM_BENCH6) return;
  ProfScope ps(c, T_OTHER);
  k2_dirichlet<ProbBench6><<<grid1(c->m2.ncorner), TPB, 0, c->stream>>>(ProbBench6(c->prob), c->m2, w);
  PFH_CUDA(cudaGetLastError());
}

void k_dirichlet_copy(pfh_ctx* c, const double* src, double* dst) {
  if (c->dim != 2) return;
  ProfScope ps(c, T_OTHER);
  k2_dirichlet_copy<<<grid1(c->m2.ncorner), TPB, 0, c->stream>>>(c->m2, src, dst);
  PFH_CUDA(cudaGetLastError());
}

void k2_mg_restrict(pfh_ctx* c, const Mesh2& mf, const Mesh2& mc, const double* tf, double scale, double* b, int bc) {
  ProfScope ps(c, T_OTHER, 8.0 * (mf.n_owned + mc.n_owned));
  k2_mg_restrict_kernel<<<grid1(mc.n_owned), TPB, 0, c->stream>>>(mf, mc, tf, scale, b, bc);
  PFH_CUDA(cudaGetLastError());
}
void k2_mg_prolong_add(pfh_ctx* c, const Mesh2& mf, const Mesh2& mc, const double* xc, double* xf, int bc) {
  ProfScope ps(c, T_OTHER, 8.0 * (2 * mf.n_owned + mc.n_owned));
  k2_mg_prolong_kernel<<<grid1(mf.n_owned), TPB, 0, c->stream>>>(mf, mc, xc, xf, bc);
  PFH_CUDA(cudaGetLastError());
}

void k_reaction_diag(pfh_ctx* c, const double* u, int field, double floor_coef, double* sig) {
  PFH_CHECK(c->dim == 2, "reaction diagonal is only used by the 2-D multi-field problems");
  ProfScope ps(c, T_OTHER);
  PFH_DISPATCH2(c, (k2_reaction_diag<P><<<grid1(c->n_owned), TPB, 0, c->stream>>>(P(c->prob), c->m2, u, field, floor_coef, sig)));
  PFH_CUDA(cudaGetLastError());
}
