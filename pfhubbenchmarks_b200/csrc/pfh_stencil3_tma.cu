// TMA-staged 3-D stencil kernels for sm_100a (the fast path of the b13d hot loop).
//
// Both kernels tile the x-y plane, march along z and keep the vertex planes they need in a ring
// of shared-memory stages that one elected thread fills with cp.async.bulk.tensor (TMA) loads
// signalled through mbarriers, so the loads of plane z+2.. overlap the arithmetic of plane z.
// TMA's out-of-bounds zero fill supplies the halo outside the domain (the boundary rows of M
// and K come from the stencil tables), and the ghost planes of a z-slab are separate one-plane
// tensors selected per plane.
//
//  k3_scalar_tma : A = am M + ak K on a scalar field (15-point P1 stencil):
//                  MODE 0  out = s*add + A in                      (lin_apply)
//                  MODE 2  zprev <- z + a (z - zprev) + b D^-1 (r - A z)   (Jacobi-Chebyshev sweep)
//                  MODE 3  z1 = b0 D^-1 r ; z2 = z1 + a z1 + b D^-1 (r - A z1)   (first two Chebyshev iterates from r alone:
//                          the stencil operand is r, scaled on the fly by b0 / diag of each neighbour's row) 24 B/vertex
//                  algorithmic traffic 32 B/vertex (MODE 2), 16-24 B/vertex (MODE 0).
//  k3_jac_tma    : y = J(u) x for the mixed system.  The nonlinear block int phi_i g'(u_h) x_h is
//                  evaluated ONCE per tetrahedron (cube-centric, 5-point Zienkiewicz-Taylor rule,
//                  pfh_elem.cuh) and scattered to the 8 cube corners through shared memory; each
//                  vertex then gathers its 8 cube contributions, so no quadrature is recomputed
//                  across the 24 tets around a vertex (the v1 gather kernel does 4x the flops).
//                  algorithmic traffic 8 (NIN + 2 NF) B/vertex = 40 B/vertex for Cahn-Hilliard.
//
// Reference hot loops replaced: PETSc MatMult / PCApply and FFC tabulate_tensor + assembly of
// J = df.derivative(F, w, dw) (dolfin/b13d.py:68; SURVEY.md section 3.3 loops B-D).
#include "pfh_tma.cuh"
#include "pfh_vertex.cuh"

namespace {
using namespace tma;

// ---------------------------------------------------------------------------------------------
// scalar operator kernel
// ---------------------------------------------------------------------------------------------
struct Coef15 {
  double c[PFH_NOFF3];
};
struct InvDiag27 {
  double v[27];  // 1 / diagonal of (am M + ak K) by boundary type
};

template <int MODE, int TX, int TY, int RY, int NS>
__global__ void __launch_bounds__(TX*(TY / RY), (TX * TY / RY <= 256 ? 4 : 2))
    k3_scalar_tma(const __grid_constant__ CUtensorMap tm_main, const __grid_constant__ CUtensorMap tm_lo,
                  const __grid_constant__ CUtensorMap tm_hi, const __grid_constant__ CUtensorMap tm_r,
                  const __grid_constant__ CUtensorMap tm_o, Mesh3 m, int has_lo, int has_hi, int has_r, int has_o, int zchunk,
                  int zpart, double am, double ak, Coef15 cI, double a, double b, double bdI, double* __restrict__ out,
                  double b0, double* __restrict__ out1, const __grid_constant__ InvDiag27 invd) {
  constexpr int W = TX + 4, H = TY + 2;  // x halo of 2: the TMA box start (x0-2) must be 16-byte aligned
  constexpr int ZT_BYTES = W * H * 8, PT_BYTES = TX * TY * 8;
  constexpr int ZT_STRIDE = round_up(ZT_BYTES, 128);
  constexpr int NR = 3;  // ring depth of the pointwise operand tiles (needed for one plane only) vs NS for the stencil operand
  constexpr int ROWS = TY / RY;
  extern __shared__ __align__(128) unsigned char smem[];
  // layout: NS stencil-operand tiles (tile + halo) | NR x [r tile | zprev tile] | NS + NR mbarriers
  unsigned char* ro_base = smem + NS * ZT_STRIDE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(ro_base + NR * 2 * PT_BYTES);
  uint64_t* bars_ro = bars + NS;
  auto ztile = [&](int p) { return reinterpret_cast<const double*>(smem + (p % NS) * ZT_STRIDE); };
  auto rtile = [&](int k) { return reinterpret_cast<const double*>(ro_base + (k % NR) * 2 * PT_BYTES); };
  auto otile = [&](int k) { return reinterpret_cast<const double*>(ro_base + (k % NR) * 2 * PT_BYTES + PT_BYTES); };

  const int tid = threadIdx.y * TX + threadIdx.x;
  const int x0 = blockIdx.x * TX, y0 = blockIdx.y * TY;
  // zpart 0: all planes in chunks; 1: the two slab-boundary planes (one per blockIdx.z); 2: interior planes [1, nzl-1)
  int zb, ze;
  if (zpart == 1) {
    zb = blockIdx.z == 0 ? 0 : m.nzl - 1;
    ze = zb + 1;
  } else {
    const int zoff = zpart == 2 ? 1 : 0, zlim = zpart == 2 ? m.nzl - 1 : m.nzl;
    zb = zoff + blockIdx.z * zchunk;
    ze = min(zlim, zb + zchunk);
  }
  const int nplanes = ze - zb + 2;  // planes zb-1 .. ze  <->  p = 0 .. nplanes-1

  if (tid == 0) {
    for (int s = 0; s < NS + NR; ++s) mbar_init(&bars[s], 1);
    fence_barrier_init();
  }
  __syncthreads();

  // plane p <-> zl = zb-1+p: the stencil operand with halo
  auto issue = [&](int p) {
    const int zl = zb - 1 + p;
    const CUtensorMap* map = &tm_main;
    int zc = zl;  // zl = -1 or nzl on a physical boundary: out of bounds in z -> zero fill
    if (zl < 0 && has_lo) {
      map = &tm_lo;
      zc = 0;
    } else if (zl >= m.nzl && has_hi) {
      map = &tm_hi;
      zc = 0;
    }
    uint64_t* bar = &bars[p % NS];
    mbar_expect_tx(bar, ZT_BYTES);
    tma_load_3d(smem + (p % NS) * ZT_STRIDE, map, x0 - 2, y0 - 1, zc, bar);
  };
  // output plane k <-> zl = zb+k: the pointwise operands r and zprev (tile only)
  auto issue_ro = [&](int k) {
    if (!has_r && !has_o) return;
    uint64_t* bar = &bars_ro[k % NR];
    unsigned char* st = ro_base + (k % NR) * 2 * PT_BYTES;
    mbar_expect_tx(bar, (has_r ? PT_BYTES : 0) + (has_o ? PT_BYTES : 0));
    if (has_r) tma_load_3d(st, &tm_r, x0, y0, zb + k, bar);
    if (has_o) tma_load_3d(st + PT_BYTES, &tm_o, x0, y0, zb + k, bar);
  };
  if (tid == 0) {
    for (int p = 0; p < NS - 1 && p < nplanes; ++p) issue(p);
    for (int k = 0; k < NR - 1 && k < ze - zb; ++k) issue_ro(k);
  }

  const int ix = x0 + threadIdx.x;
  const int txt = ix == 0 ? 0 : (ix == m.nxv - 1 ? 2 : 1);
  // per-thread row state: validity, boundary type in x-y, running global index (advanced by nxy per plane)
  bool valid[RY];
  int txy[RY];
  long long idx[RY];
#pragma unroll
  for (int rr = 0; rr < RY; ++rr) {
    const int iy = y0 + threadIdx.y + rr * ROWS;
    valid[rr] = ix < m.nxv && iy < m.nyv;
    txy[rr] = (iy == 0 ? 0 : (iy == m.nyv - 1 ? 2 : 1)) * 3 + txt;
    idx[rr] = (long long)zb * m.nxy + (long long)iy * m.nxv + ix;
  }

  for (int k = 0; k < ze - zb; ++k) {
    const int zl = zb + k;
    __syncthreads();  // every thread is done with plane k-1: its stage may be refilled
    if (tid == 0) {
      if (k + NS - 1 < nplanes) issue(k + NS - 1);
      if (k + NR - 1 < ze - zb) issue_ro(k + NR - 1);
    }
    if (k == 0) {
      mbar_wait(&bars[0], 0);
      mbar_wait(&bars[1 % NS], (1 / NS) & 1);
    }
    mbar_wait(&bars[(k + 2) % NS], ((k + 2) / NS) & 1);
    if (has_r || has_o) mbar_wait(&bars_ro[k % NR], (k / NR) & 1);

    const double* S0 = ztile(k);
    const double* S1 = ztile(k + 1);
    const double* S2 = ztile(k + 2);
    const double* Rt = rtile(k);
    const double* Ot = otile(k);
    const int gz = m.z0 + zl;
    const int tzt = gz == 0 ? 0 : (gz == m.nzv - 1 ? 2 : 1);
#pragma unroll
    for (int rr = 0; rr < RY; ++rr) {
      if (!valid[rr]) continue;
      const int ly = threadIdx.y + rr * ROWS;
      const int c = (ly + 1) * W + threadIdx.x + 2;
      const int pc = ly * TX + threadIdx.x;
      const double rv = has_r ? Rt[pc] : 0.0;
      const double ov = has_o ? Ot[pc] : 0.0;
      double v[PFH_NOFF3];
      v[0] = S1[c]; v[1] = S1[c + 1]; v[2] = S1[c - 1]; v[3] = S1[c + W]; v[4] = S1[c - W];
      v[5] = S2[c]; v[6] = S0[c]; v[7] = S1[c + W + 1]; v[8] = S1[c - W - 1];
      v[9] = S2[c + W]; v[10] = S0[c - W]; v[11] = S2[c + 1]; v[12] = S0[c - 1];
      v[13] = S2[c + W + 1]; v[14] = S0[c - W - 1];
      const int type = tzt * 9 + txy[rr];
      double Av, bd;
      double rc = 0.0, s1 = 0.0;  // MODE 3: r at this vertex; uniform operand scale b0 / diag when every neighbour is an interior row
      if (MODE == 3) {
        rc = v[0];
        const int iy = y0 + ly;
        const bool deep = ix >= 2 && ix <= m.nxv - 3 && iy >= 2 && iy <= m.nyv - 3 && gz >= 2 && gz <= m.nzv - 3;
        if (deep) {
          s1 = b0 * invd.v[13];
        } else {
          // next to the boundary the neighbours' rows differ: scale operand by operand (out-of-domain operands are the
          // zeros TMA filled in, whatever scale they get)
          constexpr int DX[PFH_NOFF3] = {0, 1, -1, 0, 0, 0, 0, 1, -1, 0, 0, 1, -1, 1, -1};
          constexpr int DY[PFH_NOFF3] = {0, 0, 0, 1, -1, 0, 0, 1, -1, 1, -1, 0, 0, 1, -1};
          constexpr int DZ[PFH_NOFF3] = {0, 0, 0, 0, 0, 1, -1, 0, 0, 1, -1, 1, -1, 1, -1};
          auto cls = [](int i, int n) { return i == 0 ? 0 : (i == n - 1 ? 2 : 1); };
          int tX[3], tY[3], tZ[3];
#pragma unroll
          for (int d = 0; d < 3; ++d) {
            tX[d] = cls(ix + d - 1, m.nxv);
            tY[d] = 3 * cls(iy + d - 1, m.nyv);
            tZ[d] = 9 * cls(gz + d - 1, m.nzv);
          }
#pragma unroll
          for (int o = 0; o < PFH_NOFF3; ++o) v[o] *= b0 * invd.v[tZ[DZ[o] + 1] + tY[DY[o] + 1] + tX[DX[o] + 1]];
          s1 = 1.0;
        }
      }
      if (type == 13) {
        // four independent partial sums: the 15 fused multiply-adds are not one dependent chain
        double s0 = cI.c[0] * v[0], s1 = cI.c[1] * v[1], s2 = cI.c[2] * v[2], s3 = cI.c[3] * v[3];
        s0 += cI.c[4] * v[4]; s1 += cI.c[5] * v[5]; s2 += cI.c[6] * v[6]; s3 += cI.c[7] * v[7];
        s0 += cI.c[8] * v[8]; s1 += cI.c[9] * v[9]; s2 += cI.c[10] * v[10]; s3 += cI.c[11] * v[11];
        s0 += cI.c[12] * v[12]; s1 += cI.c[13] * v[13]; s2 += cI.c[14] * v[14];
        Av = (s0 + s1) + (s2 + s3);
        bd = bdI;
      } else {
        vtx::Row3 row;
        vtx::get_row3(m, type, row);
        double Ms, Ks;
        vtx::mk_sums3(row, v, Ms, Ks);
        Av = am * Ms + ak * Ks;
        bd = MODE >= 2 ? b / (am * row.m[0] + ak * row.k[0]) : 0.0;
      }
      if (MODE == 0) out[idx[rr]] = a * rv + Av;
      else if (MODE == 2) out[idx[rr]] = v[0] + a * (v[0] - ov) + bd * (rv - Av);
      else {
        const double z1 = s1 * v[0];
        out1[idx[rr]] = z1;
        out[idx[rr]] = z1 + a * z1 + bd * (rc - s1 * Av);
      }
    }
#pragma unroll
    for (int rr = 0; rr < RY; ++rr) idx[rr] += m.nxy;
  }
}

// ---------------------------------------------------------------------------------------------
// Jacobian-vector product, cube-centric
// ---------------------------------------------------------------------------------------------
template <int N>
struct MapSet {
  CUtensorMap main[N], lo[N], hi[N];
};

// KIND 0: y = J(u) x (all NF rows).
// KIND 1: only the nonlinear row, from its own nonlinear column: out = (Mc[o][i] M + Kc[o][i] K) x_i + N'[x_i],
//         o = P::out(0), i = P::in(0) (the (kappa K + M_f'') block of the Cahn-Hilliard Jacobian for the Schur preconditioner).
// KIND 2: residual R(u; u0): the operand fields are u, the extra ring fields hold u0 of the fields in(.) (for the
//         problems served here the fields with a time derivative are the nonlinear inputs: c for Cahn-Hilliard).
// F32Q (KIND 0 of a QUARTIC problem only): the quadrature block in single precision (elem::cube_jac_quartic_f32)
template <class P, int KIND, int TX, int TY, int NT, int NS, bool F32Q = false>
__global__ void __launch_bounds__(NT, 2)
    k3_jac_tma(const __grid_constant__ MapSet<(KIND == 1 ? 1 : P::NF) + P::NIN> maps, P prob, Mesh3 m, LinCoefs L, int has_lo,
               int has_hi, int zchunk, double* __restrict__ out) {
  constexpr bool CROW = KIND == 1, RES = KIND == 2;
  constexpr int NX = CROW ? 1 : P::NF;  // ring-buffered operand fields x, followed by NIN fields (u for J.v, u0 for RES)
  constexpr int NFLD = NX + P::NIN;
  constexpr int W = TX + 4, H = TY + 2;  // x halo of 2 (16-byte aligned TMA box start)
  constexpr int STAGE_BYTES = W * H * 8;
  constexpr int STAGE_STRIDE = round_up(STAGE_BYTES, 128);
  constexpr int CW = TX + 1, CH = TY + 1, NCUBE = CW * CH;
  constexpr int CPAD = NCUBE | 1;  // odd pitch: corner arrays start in different banks
  extern __shared__ __align__(128) unsigned char smem[];
  // layout: [NFLD][NS] stages | contrib[NOUT][8][CPAD] | mbarriers[NS]
  double* contrib = reinterpret_cast<double*>(smem + NFLD * NS * STAGE_STRIDE);
  uint64_t* bars = reinterpret_cast<uint64_t*>(contrib + P::NOUT * 8 * CPAD);
  auto stage = [&](int f, int p) { return reinterpret_cast<const double*>(smem + (f * NS + (p % NS)) * STAGE_STRIDE); };
  auto xstage = [&](int field, int p) { return stage(CROW ? 0 : field, p); };  // operand field `field`

  const int tid = threadIdx.x;
  const int x0 = blockIdx.x * TX, y0 = blockIdx.y * TY;
  const int zb = blockIdx.z * zchunk, ze = min(m.nzl, zb + zchunk);
  const int nplanes = ze - zb + 2;  // planes zb-1 .. ze

  if (tid == 0) {
    for (int s = 0; s < NS; ++s) mbar_init(&bars[s], 1);
    fence_barrier_init();
  }
  __syncthreads();

  auto issue = [&](int p) {
    const int zl = zb - 1 + p;
    uint64_t* bar = &bars[p % NS];
    mbar_expect_tx(bar, NFLD * STAGE_BYTES);
#pragma unroll
    for (int f = 0; f < NFLD; ++f) {
      const CUtensorMap* map = &maps.main[f];
      int zc = zl;
      if (zl < 0 && has_lo) {
        map = &maps.lo[f];
        zc = 0;
      } else if (zl >= m.nzl && has_hi) {
        map = &maps.hi[f];
        zc = 0;
      }
      tma_load_3d(smem + (f * NS + (p % NS)) * STAGE_STRIDE, map, x0 - 2, y0 - 1, zc, bar);
    }
  };
  if (tid == 0)
    for (int p = 0; p < NS - 2 && p < nplanes; ++p) issue(p);

  // vertex role (tid < TX*TY)
  const int vtx_x = tid % TX, vtx_y = tid / TX;
  const int ix = x0 + vtx_x, iy = y0 + vtx_y;
  const bool is_vertex = tid < TX * TY && ix < m.nxv && iy < m.nyv;
  const int txy = (iy == 0 ? 0 : (iy == m.nyv - 1 ? 2 : 1)) * 3 + (ix == 0 ? 0 : (ix == m.nxv - 1 ? 2 : 1));
  double acc_hi[P::NOUT];
#pragma unroll
  for (int o = 0; o < P::NOUT; ++o) acc_hi[o] = 0.0;
  double a3 = 0.0, a2 = 0.0, a1 = 0.0, a0 = 0.0;
  if constexpr (P::QUARTIC) {
    if (RES) prob.res_poly(a3, a2, a1, a0);
    else prob.jac_poly(a2, a1, a0);
  }
  const elem::QuarticJac qj(a2, a1, a0, m.detJ);
  const elem::QuarticRes qr(a3, a2, a1, a0, m.detJ);

  // iteration `it` handles cube layer l = zb-1+it (between planes l and l+1) and writes plane l (it >= 1)
  for (int it = 0; it <= ze - zb; ++it) {
    const int l = zb - 1 + it;
    __syncthreads();  // plane it-2 and the previous contributions are no longer needed
    if (tid == 0 && it + NS - 2 < nplanes) issue(it + NS - 2);
    if (it == 0) mbar_wait(&bars[0], 0);
    mbar_wait(&bars[(it + 1) % NS], ((it + 1) / NS) & 1);

    // ---- cube phase: one cube (6 tets) per thread, quadrature evaluated once per tet ----
    const int gl = m.z0 + l;
    for (int q = tid; q < NCUBE; q += NT) {
      const int qx = q % CW, qy = q / CW;
      const int cx = x0 - 1 + qx, cy = y0 - 1 + qy;
      double o8[P::NOUT][8];
#pragma unroll
      for (int o = 0; o < P::NOUT; ++o)
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) o8[o][kk] = 0.0;
      if (cx >= 0 && cx <= m.nxv - 2 && cy >= 0 && cy <= m.nyv - 2 && gl >= 0 && gl <= m.nzv - 2) {
        double u[P::NIN][8], x[P::NIN][8];
        const int base = qy * W + qx + 1;
#pragma unroll
        for (int f = 0; f < P::NIN; ++f) {
          // J.v: u = linearisation point (extra ring field), x = operand; residual: u = the operand itself
          const double* u0 = RES ? xstage(P::in(f), it) : stage(NX + f, it);
          const double* u1 = RES ? xstage(P::in(f), it + 1) : stage(NX + f, it + 1);
          const double* xa = xstage(P::in(f), it);
          const double* xb = xstage(P::in(f), it + 1);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            const int off = base + (kk >> 1) * W + (kk & 1);
            u[f][kk] = u0[off];
            u[f][kk + 4] = u1[off];
            x[f][kk] = xa[off];
            x[f][kk + 4] = xb[off];
          }
        }
        if constexpr (P::QUARTIC) {
          if (RES) elem::cube_res_quartic(qr, u[0], o8[0]);
          else if (F32Q) elem::cube_jac_quartic_f32(qj, u[0], x[0], o8[0]);
          else elem::cube_jac_quartic(qj, u[0], x[0], o8[0]);
        } else {
          if (RES) elem::cube_load<P, false>(prob, u, x, m.detJ, o8);
          else elem::cube_load<P, true>(prob, u, x, m.detJ, o8);
        }
      }
#pragma unroll
      for (int o = 0; o < P::NOUT; ++o)
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) contrib[(o * 8 + kk) * CPAD + q] = o8[o][kk];
    }
    __syncthreads();

    // ---- vertex phase: gather the 8 cube contributions, add the linear 15-point part, store plane l ----
    if (is_vertex) {
      double lo[P::NOUT], hi[P::NOUT];
#pragma unroll
      for (int o = 0; o < P::NOUT; ++o) {
        double slo = 0.0, shi = 0.0;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const int q = (vtx_y + 1 - (kk >> 1)) * CW + (vtx_x + 1 - (kk & 1));
          slo += contrib[(o * 8 + kk) * CPAD + q];
          shi += contrib[(o * 8 + 4 + kk) * CPAD + q];
        }
        lo[o] = slo;
        hi[o] = shi;
      }
      if (it >= 1) {
        const int gz = m.z0 + l;
        const int type = (gz == 0 ? 0 : (gz == m.nzv - 1 ? 2 : 1)) * 9 + txy;
        vtx::Row3 row;
        vtx::get_row3(m, type, row);
        const int c = (vtx_y + 1) * W + vtx_x + 2;
        double Ms[NX], Ks[NX];
#pragma unroll
        for (int f = 0; f < NX; ++f) {
          const double* S0 = stage(f, it - 1);
          const double* S1 = stage(f, it);
          const double* S2 = stage(f, it + 1);
          double v[PFH_NOFF3];
          v[0] = S1[c]; v[1] = S1[c + 1]; v[2] = S1[c - 1]; v[3] = S1[c + W]; v[4] = S1[c - W];
          v[5] = S2[c]; v[6] = S0[c]; v[7] = S1[c + W + 1]; v[8] = S1[c - W - 1];
          v[9] = S2[c + W]; v[10] = S0[c - W]; v[11] = S2[c + 1]; v[12] = S0[c - 1];
          v[13] = S2[c + W + 1]; v[14] = S0[c - W - 1];
          vtx::mk_sums3(row, v, Ms[f], Ks[f]);
        }
        const long long i = (long long)l * m.nxy + (long long)iy * m.nxv + ix;
        if (CROW) {
          out[i] = L.Mc[P::out(0)][P::in(0)] * Ms[0] + L.Kc[P::out(0)][P::in(0)] * Ks[0] + acc_hi[0] + lo[0];
        } else {
          double y[P::NF];
#pragma unroll
          for (int aa = 0; aa < P::NF; ++aa) {
            double s = 0.0;
#pragma unroll
            for (int bb = 0; bb < P::NF; ++bb) s += L.Mc[aa][bb] * Ms[bb % NX] + L.Kc[aa][bb] * Ks[bb % NX];
            y[aa] = s;
          }
#pragma unroll
          for (int o = 0; o < P::NOUT; ++o) y[P::out(o)] += acc_hi[o] + lo[o];
          if (RES) {  // - Mp_a M u0_a for the fields with a time derivative
#pragma unroll
            for (int f = 0; f < P::NIN; ++f) {
              const double* S0 = stage(NX + f, it - 1);
              const double* S1 = stage(NX + f, it);
              const double* S2 = stage(NX + f, it + 1);
              double v[PFH_NOFF3], ms0, ks0;
              v[0] = S1[c]; v[1] = S1[c + 1]; v[2] = S1[c - 1]; v[3] = S1[c + W]; v[4] = S1[c - W];
              v[5] = S2[c]; v[6] = S0[c]; v[7] = S1[c + W + 1]; v[8] = S1[c - W - 1];
              v[9] = S2[c + W]; v[10] = S0[c - W]; v[11] = S2[c + 1]; v[12] = S0[c - 1];
              v[13] = S2[c + W + 1]; v[14] = S0[c - W - 1];
              vtx::mk_sums3(row, v, ms0, ks0);
              y[P::in(f)] -= L.Mp[P::in(f)] * ms0;
            }
          }
#pragma unroll
          for (int aa = 0; aa < P::NF; ++aa) out[aa * m.n_owned + i] = y[aa];
        }
      }
#pragma unroll
      for (int o = 0; o < P::NOUT; ++o) acc_hi[o] = hi[o];
    }
  }
}

constexpr int S_TX = 64, S_TY = 8, S_RY = 2, S_NS = 4, S_ZCH = 32;
constexpr int J_TX = 32, J_TY = 8, J_NT = 320, J_NS = 4, J_ZCH = 32;


}  // namespace

bool pfh_tma_ok(const pfh_ctx* c) {
  return c->dim == 3 && c->use_tma && (c->m3.nxv % 2) == 0 && c->m3.nxv >= 8 && c->m3.nyv >= 4 && c->m3.nzl >= 2;
}

struct ScalarArgs {
  CUtensorMap tm, tl, th, tr, to;
  Mesh3 m;
  bool lo, hi, has_r, has_o;
  int zpart;
  double am, ak;
  Coef15 cI;
  double a, b, bdI;
  FieldRef in;
  const double* r;
  double* out;
  int mode;
  double b0 = 0.0;       // MODE 3
  double* out1 = nullptr;
  InvDiag27 invd = {};
};

template <int TX, int TY, int RY, int NS, int ZCH>
bool scalar_launch(pfh_ctx* c, ScalarArgs& A) {
  const Mesh3& m = A.m;
  A.tm = make_map(A.in.p, m.nxv, m.nyv, m.nzl, TX + 4, TY + 2);
  A.tl = A.lo ? make_map(A.in.glo, m.nxv, m.nyv, 1, TX + 4, TY + 2) : A.tm;
  A.th = A.hi ? make_map(A.in.ghi, m.nxv, m.nyv, 1, TX + 4, TY + 2) : A.tm;
  A.tr = A.has_r ? make_map(A.r, m.nxv, m.nyv, m.nzl, TX, TY) : A.tm;
  A.to = A.has_o ? make_map(A.out, m.nxv, m.nyv, m.nzl, TX, TY) : A.tm;
  const int zp = A.zpart;
  const int nz_part = zp == 1 ? 2 * ZCH : (zp == 2 ? m.nzl - 2 : m.nzl);  // zpart 1: gridDim.z = 2
  if (nz_part <= 0) return true;
  const dim3 grid((m.nxv + TX - 1) / TX, (m.nyv + TY - 1) / TY, (nz_part + ZCH - 1) / ZCH);
  const dim3 block(TX, TY / RY);
  constexpr size_t smem = NS * round_up((TX + 4) * (TY + 2) * 8, 128) + 3 * 2 * TX * TY * 8 + (NS + 3) * sizeof(uint64_t);
  static bool attr_set = false;
  if (!attr_set) {
    PFH_CUDA(cudaFuncSetAttribute(k3_scalar_tma<0, TX, TY, RY, NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    PFH_CUDA(cudaFuncSetAttribute(k3_scalar_tma<2, TX, TY, RY, NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    PFH_CUDA(cudaFuncSetAttribute(k3_scalar_tma<3, TX, TY, RY, NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  if (A.mode == 0)
    k3_scalar_tma<0, TX, TY, RY, NS><<<grid, block, smem, c->stream>>>(A.tm, A.tl, A.th, A.tr, A.to, m, A.lo, A.hi, A.has_r, 0, ZCH, zp, A.am,
                                                                      A.ak, A.cI, A.a, A.b, 0.0, A.out, 0.0, nullptr, A.invd);
  else if (A.mode == 2)
    k3_scalar_tma<2, TX, TY, RY, NS><<<grid, block, smem, c->stream>>>(A.tm, A.tl, A.th, A.tr, A.to, m, A.lo, A.hi, A.has_r, A.has_o, ZCH, zp,
                                                                      A.am, A.ak, A.cI, A.a, A.b, A.bdI, A.out, 0.0, nullptr, A.invd);
  else
    k3_scalar_tma<3, TX, TY, RY, NS><<<grid, block, smem, c->stream>>>(A.tm, A.tl, A.th, A.tr, A.to, m, A.lo, A.hi, 0, 0, ZCH, zp, A.am, A.ak,
                                                                      A.cI, A.a, A.b, A.bdI, A.out, A.b0, A.out1, A.invd);
  PFH_CUDA(cudaGetLastError());
  return true;
}

// MODE 0 (lin_apply) and MODE 2 (Chebyshev sweep) through the TMA kernel; returns false if the
// operands do not qualify (caller falls back to the gather kernel)
bool k3_scalar_tma_launch(pfh_ctx* c, int mode, double am, double ak, double a, double b, FieldRef in, const double* r,
                          double* out, bool out_is_zero) {
  if (!pfh_tma_ok(c) || !aligned16(in.p) || !aligned16(out) || (r && !aligned16(r))) return false;
  const Mesh3& m = c->m3;
  const bool lo = in.glo != nullptr && c->cfg.rank > 0, hi = in.ghi != nullptr && c->cfg.rank < c->cfg.nranks - 1;
  if ((lo && !aligned16(in.glo)) || (hi && !aligned16(in.ghi))) return false;
  const bool has_r = (mode == 2 || a != 0.0) && r != nullptr;
  const bool has_o = mode == 2 && !out_is_zero;
  CUtensorMap tm;
  memset(&tm, 0, sizeof tm);
  Coef15 cI;
  for (int o = 0; o < PFH_NOFF3; ++o) cI.c[o] = am * m.mI[o] + (o < PFH_NKOFF3 ? ak * m.kI[o] : 0.0);
  ScalarArgs A{tm, tm, tm, tm, tm, m, lo, hi, has_r, has_o, c->zpart, am, ak, cI, a, b, b / cI.c[0], in, r, out, mode};
  switch (c->scalar_variant) {
    case 1: return scalar_launch<64, 8, 2, 5, 32>(c, A);
    case 2: return scalar_launch<64, 8, 2, 4, 64>(c, A);
    case 3: return scalar_launch<64, 16, 4, 3, 32>(c, A);
    case 4: return scalar_launch<64, 8, 2, 3, 32>(c, A);
    case 5: return scalar_launch<64, 7, 1, 4, 64>(c, A);   // 8 x 74 = 592 tiles on a 512^2 plane = 148 SMs x 4 exactly
    case 6: return scalar_launch<64, 8, 2, 4, 64>(c, A);
    case 7: return scalar_launch<64, 7, 1, 4, 32>(c, A);
    default: return scalar_launch<S_TX, S_TY, S_RY, S_NS, S_ZCH>(c, A);
  }
}

// z1 = b0 D^-1 r and z2 = z1 + a1 z1 + b1 D^-1 (r - A z1) in one pass over r (MODE 3): replaces the pointwise first
// Chebyshev step and the first sweep (40 B/vertex together) by 24 B/vertex
bool k3_cheb_first2_tma_launch(pfh_ctx* c, double am, double ak, double b0, double a1, double b1, FieldRef r, double* z1, double* z2) {
  if (!pfh_tma_ok(c) || !aligned16(r.p) || !aligned16(z1) || !aligned16(z2)) return false;
  const Mesh3& m = c->m3;
  const bool lo = r.glo != nullptr && c->cfg.rank > 0, hi = r.ghi != nullptr && c->cfg.rank < c->cfg.nranks - 1;
  if ((lo && !aligned16(r.glo)) || (hi && !aligned16(r.ghi))) return false;
  CUtensorMap tm;
  memset(&tm, 0, sizeof tm);
  Coef15 cI;
  for (int o = 0; o < PFH_NOFF3; ++o) cI.c[o] = am * m.mI[o] + (o < PFH_NKOFF3 ? ak * m.kI[o] : 0.0);
  ScalarArgs A{tm, tm, tm, tm, tm, m, lo, hi, false, false, c->zpart, am, ak, cI, a1, b1, b1 / cI.c[0], r, nullptr, z2, 3, b0, z1};
  if (c->h_mT3.empty()) {  // host copy of the type tables, for the inverse diagonals
    Mesh3 mm = m;
    pfh_build_tables3(mm, c->h_mT3, c->h_kT3);
  }
  for (int t = 0; t < 27; ++t) A.invd.v[t] = 1.0 / (am * c->h_mT3[t * PFH_NOFF3] + ak * c->h_kT3[t * PFH_NKOFF3]);
  return scalar_launch<S_TX, S_TY, S_RY, S_NS, S_ZCH>(c, A);
}

// y = J(ulin) x for Cahn-Hilliard on the Kuhn mesh; ghosts: gX_* for x, gU_* for ulin
bool k3_jac_tma_launch(pfh_ctx* c, const LinCoefs& L, const double* ulin, const double* x, double* y) {
  if (!pfh_tma_ok(c) || c->prob.id != PFH_PROBLEM_CH || !aligned16(ulin) || !aligned16(x)) return false;
  using P = ProbCH;
  const Mesh3& m = c->m3;
  if ((m.n_owned & 1) || (c->nplane & 1)) return false;  // field offsets must stay 16-byte aligned
  const bool lo = c->nplane > 0 && c->cfg.rank > 0, hi = c->nplane > 0 && c->cfg.rank < c->cfg.nranks - 1;
  MapSet<P::NF + P::NIN> maps;
  for (int f = 0; f < P::NF + P::NIN; ++f) {
    const bool isx = f < P::NF;
    const int fld = isx ? f : P::in(f - P::NF);
    const double* base = (isx ? x : ulin) + (long long)fld * m.n_owned;
    const double* glo = (isx ? c->gX_lo : c->gU_lo);
    const double* ghi = (isx ? c->gX_hi : c->gU_hi);
    maps.main[f] = make_map(base, m.nxv, m.nyv, m.nzl, J_TX + 4, J_TY + 2);
    maps.lo[f] = lo ? make_map(glo + (long long)fld * c->nplane, m.nxv, m.nyv, 1, J_TX + 4, J_TY + 2) : maps.main[f];
    maps.hi[f] = hi ? make_map(ghi + (long long)fld * c->nplane, m.nxv, m.nyv, 1, J_TX + 4, J_TY + 2) : maps.main[f];
  }
  constexpr int stage = round_up((J_TX + 4) * (J_TY + 2) * 8, 128);
  constexpr int ncube = (J_TX + 1) * (J_TY + 1);
  constexpr size_t smem = (size_t)(P::NF + P::NIN) * J_NS * stage + (size_t)P::NOUT * 8 * (ncube | 1) * 8 + J_NS * sizeof(uint64_t);
  auto kern = k3_jac_tma<P, 0, J_TX, J_TY, J_NT, J_NS>;
  auto kern32 = k3_jac_tma<P, 0, J_TX, J_TY, J_NT, J_NS, true>;
  static bool attr_set = false;
  if (!attr_set) {
    PFH_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    PFH_CUDA(cudaFuncSetAttribute(kern32, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  const dim3 grid((m.nxv + J_TX - 1) / J_TX, (m.nyv + J_TY - 1) / J_TY, (m.nzl + J_ZCH - 1) / J_ZCH);
  if (c->jac_quadrature_fp32) kern32<<<grid, J_NT, smem, c->stream>>>(maps, P(c->prob), m, L, lo, hi, J_ZCH, y);
  else kern<<<grid, J_NT, smem, c->stream>>>(maps, P(c->prob), m, L, lo, hi, J_ZCH, y);
  PFH_CUDA(cudaGetLastError());
  return true;
}

// out = (Mc[mu][c] M + Kc[mu][c] K) xc + M_{g'}(ulin) xc : the (mu, c) block of the Cahn-Hilliard Jacobian applied to a
// scalar field (ghosts of xc in `xc.glo/ghi`, of ulin in gU_*)
bool k3_ch_crow_tma_launch(pfh_ctx* c, const LinCoefs& L, const double* ulin, FieldRef xc, double* out) {
  if (!pfh_tma_ok(c) || c->prob.id != PFH_PROBLEM_CH || !aligned16(ulin) || !aligned16(xc.p)) return false;
  using P = ProbCH;
  const Mesh3& m = c->m3;
  if ((m.n_owned & 1) || (c->nplane & 1)) return false;
  const bool lo = c->nplane > 0 && c->cfg.rank > 0, hi = c->nplane > 0 && c->cfg.rank < c->cfg.nranks - 1;
  MapSet<1 + P::NIN> maps;
  maps.main[0] = make_map(xc.p, m.nxv, m.nyv, m.nzl, J_TX + 4, J_TY + 2);
  maps.lo[0] = lo ? make_map(xc.glo, m.nxv, m.nyv, 1, J_TX + 4, J_TY + 2) : maps.main[0];
  maps.hi[0] = hi ? make_map(xc.ghi, m.nxv, m.nyv, 1, J_TX + 4, J_TY + 2) : maps.main[0];
  const int fld = P::in(0);
  maps.main[1] = make_map(ulin + (long long)fld * m.n_owned, m.nxv, m.nyv, m.nzl, J_TX + 4, J_TY + 2);
  maps.lo[1] = lo ? make_map(c->gU_lo + (long long)fld * c->nplane, m.nxv, m.nyv, 1, J_TX + 4, J_TY + 2) : maps.main[1];
  maps.hi[1] = hi ? make_map(c->gU_hi + (long long)fld * c->nplane, m.nxv, m.nyv, 1, J_TX + 4, J_TY + 2) : maps.main[1];
  constexpr int stage = round_up((J_TX + 4) * (J_TY + 2) * 8, 128);
  constexpr int ncube = (J_TX + 1) * (J_TY + 1);
  constexpr size_t smem = (size_t)(1 + P::NIN) * J_NS * stage + (size_t)P::NOUT * 8 * (ncube | 1) * 8 + J_NS * sizeof(uint64_t);
  auto kern = k3_jac_tma<P, 1, J_TX, J_TY, J_NT, J_NS>;
  static bool attr_set = false;
  if (!attr_set) {
    PFH_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  const dim3 grid((m.nxv + J_TX - 1) / J_TX, (m.nyv + J_TY - 1) / J_TY, (m.nzl + J_ZCH - 1) / J_ZCH);
  kern<<<grid, J_NT, smem, c->stream>>>(maps, P(c->prob), m, L, lo, hi, J_ZCH, out);
  PFH_CUDA(cudaGetLastError());
  return true;
}

// R = R(u; u0, dt) for Cahn-Hilliard on the Kuhn mesh (ghosts: gU_* for u, gU0_* for u0)
bool k3_res_tma_launch(pfh_ctx* c, const LinCoefs& L, const double* u, const double* u0, double* R) {
  if (!pfh_tma_ok(c) || c->prob.id != PFH_PROBLEM_CH || !aligned16(u) || !aligned16(u0)) return false;
  using P = ProbCH;
  const Mesh3& m = c->m3;
  if ((m.n_owned & 1) || (c->nplane & 1)) return false;
  for (int a = 0; a < P::NF; ++a) {  // the kernel subtracts Mp_a M u0_a only for the nonlinear input fields
    bool is_in = false;
    for (int f = 0; f < P::NIN; ++f) is_in |= P::in(f) == a;
    if (L.Mp[a] != 0.0 && !is_in) return false;
  }
  const bool lo = c->nplane > 0 && c->cfg.rank > 0, hi = c->nplane > 0 && c->cfg.rank < c->cfg.nranks - 1;
  MapSet<P::NF + P::NIN> maps;
  for (int f = 0; f < P::NF + P::NIN; ++f) {
    const bool isx = f < P::NF;
    const int fld = isx ? f : P::in(f - P::NF);
    const double* base = (isx ? u : u0) + (long long)fld * m.n_owned;
    const double* glo = (isx ? c->gU_lo : c->gU0_lo);
    const double* ghi = (isx ? c->gU_hi : c->gU0_hi);
    maps.main[f] = make_map(base, m.nxv, m.nyv, m.nzl, J_TX + 4, J_TY + 2);
    maps.lo[f] = lo ? make_map(glo + (long long)fld * c->nplane, m.nxv, m.nyv, 1, J_TX + 4, J_TY + 2) : maps.main[f];
    maps.hi[f] = hi ? make_map(ghi + (long long)fld * c->nplane, m.nxv, m.nyv, 1, J_TX + 4, J_TY + 2) : maps.main[f];
  }
  constexpr int stage = round_up((J_TX + 4) * (J_TY + 2) * 8, 128);
  constexpr int ncube = (J_TX + 1) * (J_TY + 1);
  constexpr size_t smem = (size_t)(P::NF + P::NIN) * J_NS * stage + (size_t)P::NOUT * 8 * (ncube | 1) * 8 + J_NS * sizeof(uint64_t);
  auto kern = k3_jac_tma<P, 2, J_TX, J_TY, J_NT, J_NS>;
  static bool attr_set = false;
  if (!attr_set) {
    PFH_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  const dim3 grid((m.nxv + J_TX - 1) / J_TX, (m.nyv + J_TY - 1) / J_TY, (m.nzl + J_ZCH - 1) / J_ZCH);
  kern<<<grid, J_NT, smem, c->stream>>>(maps, P(c->prob), m, L, lo, hi, J_ZCH, R);
  PFH_CUDA(cudaGetLastError());
  return true;
}
