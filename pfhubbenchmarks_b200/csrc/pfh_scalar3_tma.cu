// Scalar SPD operator A = am M + ak K on the 3-D Kuhn mesh (15-point P1 stencil), the workhorse of the
// Schur/Chebyshev preconditioner (pfh_solver.cpp) -- the B200 stand-in for PETSc's PCApply inside KSPSolve
// (dolfin/b13d.py:86-98 'gmres' + preconditioner; SURVEY.md section 3.3 loop D).
//
//   MODE 0  out = a*add + A in   [out1 = in]                          (lin_apply)                16-24 B/vertex
//   MODE 2  out = z + a (z - zprev) + b D^-1 (r - A z)                  (Jacobi-Chebyshev sweep)   32 B/vertex
//           (zprev either loaded, or -- third iterate -- recomputed as b0 D^-1 r from the r it reads anyway: 24 B/vertex)
//   MODE 3  z1 = b0 D^-1 r ; z2 = z1 + a z1 + b D^-1 (r - A z1)         (first two iterates from r alone) 24 B/vertex
//
// One CTA owns a TX x TY tile of the x-y plane and marches along z over `zchunk` planes.  Every plane of the stencil
// operand is staged once in shared memory (tile + halo) by a ring of NS stages; stage completion is signalled on
// mbarriers.  Two loaders fill the ring:
//   TMA      one elected thread issues cp.async.bulk.tensor.3d boxes; out-of-bounds zero fill supplies the halo
//            outside the domain.  Needs 16-byte aligned rows, i.e. an even vertex count in x.
//   CPASYNC  every thread issues 8-byte cp.async copies (zero fill by src-size 0) and arrives on the same mbarrier
//            with cp.async.mbarrier.arrive.noinc.  Serves odd vertex counts (the reference's 51^3 b13d mesh).
// Along the z-march every thread keeps the four operands of plane k-1 and four of plane k that it already loaded
// in registers, so a vertex costs 7 shared-memory loads instead of 15.
//
// A launch works on an arbitrary window of planes: output planes are launch-local zl = 0 .. nzo-1 (global plane
// gz0 + zl); the stencil operand is addressed through a main array plus optional lower / upper ghost arrays of any
// width.  This serves both the classic one-ghost-plane layout and the wide-ghost preconditioner of the multi-GPU
// path, where Chebyshev sweeps run redundantly on a shrinking halo instead of exchanging a plane per sweep.
#include "pfh_tma.cuh"
#include "pfh_vertex.cuh"

namespace {
using namespace tma;

struct Coef15 {
  double c[PFH_NOFF3];
};
struct InvDiag27 {
  double v[27];  // 1 / diagonal of (am M + ak K) by boundary type
};
struct ScalarMaps {
  CUtensorMap main, lo, hi, r, r_lo, r_hi, o;
};
struct ScalarParams {
  Mesh3 m;                     // nxv, nyv, nzv, nxy and the stencil tables; z0 / nzl are not used here
  int gz0, nzo, zchunk;        // global z of launch-local plane 0, number of output planes, planes per CTA
  int in_off, in_n;            // main operand array: index of launch-local plane 0, number of planes it holds
  int n_lo, n_hi;              // planes in the lower / upper ghost arrays (main index -n_lo..-1 and in_n..in_n+n_hi-1)
  int has_r, has_o;
  int r_off, r_n, r_nlo, r_nhi;  // pointwise operand r: same addressing as the stencil operand
  double am, ak, a, b, bdI, b0;
  double ob0;                  // MODE 2 without a zprev operand: zprev = ob0 D^-1 r (0: zprev = 0)
  Coef15 cI;
  InvDiag27 invd;
  double* out;                 // launch-local plane 0
  double* out1;                // MODE 3: z1; MODE 0: optional copy of the operand
  const double *in, *in_lo, *in_hi, *r, *r_lo, *r_hi, *o;  // raw operand pointers (CPASYNC loader); o at launch-local plane 0
};

// next to the boundary the neighbours' rows have different diagonals: scale operand by operand
// (only on the two planes next to each z boundary: kept out of line, working on a copy in local memory, so that it
// does not cost the main loop registers)
__device__ __noinline__ void first2_scale_boundary(double* v, const InvDiag27& invd, double b0, int ix, int iy, int gz,
                                                   const Mesh3& m) {
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
  // out-of-domain operands are zeros, whatever scale they get
#pragma unroll
  for (int o = 0; o < PFH_NOFF3; ++o) v[o] *= b0 * invd.v[tZ[DZ[o] + 1] + tY[DY[o] + 1] + tX[DX[o] + 1]];
}

template <int MODE, int TX, int TY, int RY, int NS, bool TMA>
__global__ void __launch_bounds__(TX*(TY / RY), (TX * TY / RY <= 128 ? 8 : (TX * TY / RY <= 256 ? 4 : 2)))
    k3_scalar_tma(const __grid_constant__ ScalarMaps M, const __grid_constant__ ScalarParams P) {
  constexpr int W = TX + 4, H = TY + 2;  // x halo of 2: the TMA box start (x0-2) must be 16-byte aligned
  constexpr int NT = TX * (TY / RY), ROWS = TY / RY;
  constexpr int ZT_BYTES = W * H * 8, PT_BYTES = TX * TY * 8;
  constexpr int ZT_STRIDE = round_up(ZT_BYTES, 128);
  constexpr int NR = 3;  // ring depth of the pointwise operand tiles (r, zprev)
  extern __shared__ __align__(128) unsigned char smem[];
  // layout: NS stencil-operand tiles (tile + halo) | NR x [r tile | zprev tile] | NS + NR mbarriers
  unsigned char* ro_base = smem + NS * ZT_STRIDE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(ro_base + NR * 2 * PT_BYTES);
  uint64_t* bars_ro = bars + NS;
  auto ztile = [&](int p) { return reinterpret_cast<const double*>(smem + (p % NS) * ZT_STRIDE); };

  const Mesh3& m = P.m;
  const int tid = threadIdx.y * TX + threadIdx.x;
  const int x0 = blockIdx.x * TX, y0 = blockIdx.y * TY;
  const int zb = blockIdx.z * P.zchunk, ze = min(P.nzo, zb + P.zchunk);
  const int nout = ze - zb, nplanes = nout + 2;  // input planes zb-1 .. ze  <->  p = 0 .. nplanes-1
  const bool has_ro = P.has_r || P.has_o;

  if (tid == 0) {
    for (int s = 0; s < NS + NR; ++s) mbar_init(&bars[s], TMA ? 1 : NT);
    fence_barrier_init();
  }
  __syncthreads();

  // plane p <-> launch-local input plane zb-1+p: the stencil operand with halo
  auto issue = [&](int p) {
    const int zm = zb - 1 + p + P.in_off;  // index in the main array
    uint64_t* bar = &bars[p % NS];
    unsigned char* dst = smem + (p % NS) * ZT_STRIDE;
    if constexpr (TMA) {
      if (tid == 0) {
        const CUtensorMap* map = &M.main;
        int zc = zm;  // outside every array: beyond the physical boundary (or a plane nobody needs) -> zero fill
        if (zm < 0 && P.n_lo > 0) {
          map = &M.lo;
          zc = zm + P.n_lo;
        } else if (zm >= P.in_n && P.n_hi > 0) {
          map = &M.hi;
          zc = zm - P.in_n;
        }
        mbar_expect_tx(bar, ZT_BYTES);
        tma_load_3d(dst, map, x0 - 2, y0 - 1, zc, bar);
      }
    } else {
      const double* src = nullptr;
      if (zm >= 0 && zm < P.in_n) src = P.in + (long long)zm * m.nxy;
      else if (zm < 0 && zm + P.n_lo >= 0) src = P.in_lo + (long long)(zm + P.n_lo) * m.nxy;
      else if (zm >= P.in_n && zm - P.in_n < P.n_hi) src = P.in_hi + (long long)(zm - P.in_n) * m.nxy;
      for (int i = tid; i < W * H; i += NT) {
        const int gx = x0 - 2 + i % W, gy = y0 - 1 + i / W;
        const bool ok = src != nullptr && gx >= 0 && gx < m.nxv && gy >= 0 && gy < m.nyv;
        cp_async8(dst + i * 8, ok ? src + (long long)gy * m.nxv + gx : P.out, ok ? 8 : 0);
      }
      cp_async_mbar_arrive(bar);
    }
  };
  // output plane k <-> launch-local plane zb+k: the pointwise operands r and zprev (tile only)
  auto issue_ro = [&](int k) {
    if (!has_ro) return;
    uint64_t* bar = &bars_ro[k % NR];
    unsigned char* st = ro_base + (k % NR) * 2 * PT_BYTES;
    const int zr = zb + k + P.r_off;  // index of this plane in the main array of r
    if constexpr (TMA) {
      if (tid == 0) {
        mbar_expect_tx(bar, (P.has_r ? PT_BYTES : 0) + (P.has_o ? PT_BYTES : 0));
        if (P.has_r) {
          const CUtensorMap* map = &M.r;
          int zc = zr;
          if (zr < 0 && P.r_nlo > 0) {
            map = &M.r_lo;
            zc = zr + P.r_nlo;
          } else if (zr >= P.r_n && P.r_nhi > 0) {
            map = &M.r_hi;
            zc = zr - P.r_n;
          }
          tma_load_3d(st, map, x0, y0, zc, bar);
        }
        if (P.has_o) tma_load_3d(st + PT_BYTES, &M.o, x0, y0, zb + k, bar);
      }
    } else {
      const double* rs = nullptr;
      if (P.has_r) {
        if (zr >= 0 && zr < P.r_n) rs = P.r + (long long)zr * m.nxy;
        else if (zr < 0 && zr + P.r_nlo >= 0) rs = P.r_lo + (long long)(zr + P.r_nlo) * m.nxy;
        else if (zr >= P.r_n && zr - P.r_n < P.r_nhi) rs = P.r_hi + (long long)(zr - P.r_n) * m.nxy;
      }
      const long long pl = (long long)(zb + k) * m.nxy;
      for (int i = tid; i < TX * TY; i += NT) {
        const int gx = x0 + i % TX, gy = y0 + i / TX;
        const bool ok = gx < m.nxv && gy < m.nyv;
        const long long g = (long long)gy * m.nxv + gx;
        if (P.has_r) cp_async8(st + i * 8, (ok && rs) ? rs + g : P.out, (ok && rs) ? 8 : 0);
        if (P.has_o) cp_async8(st + PT_BYTES + i * 8, ok ? P.o + pl + g : P.out, ok ? 8 : 0);
      }
      cp_async_mbar_arrive(bar);
    }
  };
  for (int p = 0; p < NS && p < nplanes; ++p) issue(p);
  for (int k = 0; k < NR - 1 && k < nout; ++k) issue_ro(k);

  const int ix = x0 + threadIdx.x;
  const int txt = ix == 0 ? 0 : (ix == m.nxv - 1 ? 2 : 1);
  // per-thread row state: validity, boundary type in x-y, running output index (advanced by nxy per plane)
  bool valid[RY], deep_xy[RY];
  int txy[RY], cidx[RY];
  long long idx[RY];
  // z-march register reuse (MODE 0, 2): the operands of planes k-1 and k a thread has already loaded.  MODE 3 spends
  // the registers on its operand scales instead and reads all 15 operands from shared memory.
  constexpr bool REUSE = MODE != 3;
  double pa[REUSE ? RY : 1][4];  // plane k-1: c, c-W, c-1, c-W-1
  double pb[REUSE ? RY : 1][4];  // plane k  : c, c+W, c+1, c+W+1
  // MODE 3: scale b0 / diag of the row of each of the 7 distinct x-y neighbours (0,0) (1,0) (-1,0) (0,1) (0,-1) (1,1) (-1,-1)
  // on a plane away from the z boundaries; fixed per thread, so the vertices next to an x or y boundary cost 15
  // multiplications instead of a table walk per plane
  double sc[MODE == 3 ? RY : 1][7];
  mbar_wait(&bars[0], 0);
  mbar_wait(&bars[1 % NS], (1 / NS) & 1);
  {
    const double* S0 = ztile(0);
    const double* S1 = ztile(1);
#pragma unroll
    for (int rr = 0; rr < RY; ++rr) {
      const int ly = threadIdx.y + rr * ROWS;
      const int iy = y0 + ly;
      valid[rr] = ix < m.nxv && iy < m.nyv;
      deep_xy[rr] = ix >= 2 && ix <= m.nxv - 3 && iy >= 2 && iy <= m.nyv - 3;
      txy[rr] = (iy == 0 ? 0 : (iy == m.nyv - 1 ? 2 : 1)) * 3 + txt;
      idx[rr] = (long long)zb * m.nxy + (long long)iy * m.nxv + ix;
      const int c = (ly + 1) * W + threadIdx.x + 2;
      cidx[rr] = c;
      if constexpr (REUSE) {
        pa[rr][0] = S0[c]; pa[rr][1] = S0[c - W]; pa[rr][2] = S0[c - 1]; pa[rr][3] = S0[c - W - 1];
        pb[rr][0] = S1[c]; pb[rr][1] = S1[c + W]; pb[rr][2] = S1[c + 1]; pb[rr][3] = S1[c + W + 1];
      }
      if constexpr (MODE == 3) {
        constexpr int DX7[7] = {0, 1, -1, 0, 0, 1, -1}, DY7[7] = {0, 0, 0, 1, -1, 1, -1};
        auto cls = [](int i, int n) { return i == 0 ? 0 : (i == n - 1 ? 2 : 1); };
#pragma unroll
        for (int j = 0; j < 7; ++j) sc[rr][j] = P.b0 * P.invd.v[9 + 3 * cls(iy + DY7[j], m.nyv) + cls(ix + DX7[j], m.nxv)];
      }
    }
  }

  for (int k = 0; k < nout; ++k) {
    // REUSE: every thread has read plane k (kept in registers from here on), its stage may be refilled; otherwise plane k
    // is still read in this iteration and the stage of plane k-1 is the free one
    __syncthreads();
    if constexpr (REUSE) {
      if (k + NS < nplanes) issue(k + NS);
    } else {
      if (k >= 1 && k - 1 + NS < nplanes) issue(k - 1 + NS);
    }
    if (k + NR - 1 < nout) issue_ro(k + NR - 1);
    mbar_wait(&bars[(k + 2) % NS], ((k + 2) / NS) & 1);
    if (has_ro) mbar_wait(&bars_ro[k % NR], (k / NR) & 1);

    const double* S1 = ztile(k + 1);
    const double* S2 = ztile(k + 2);
    const double* Rt = reinterpret_cast<const double*>(ro_base + (k % NR) * 2 * PT_BYTES);
    const double* Ot = Rt + TX * TY;
    const int gz = P.gz0 + zb + k;
    const int tzt = gz == 0 ? 0 : (gz == m.nzv - 1 ? 2 : 1);
    const bool deep_z = gz >= 2 && gz <= m.nzv - 3;
#pragma unroll
    for (int rr = 0; rr < RY; ++rr) {
      const int ly = threadIdx.y + rr * ROWS;
      const int c = cidx[rr];
      const int pc = ly * TX + threadIdx.x;
      const double rv = P.has_r ? Rt[pc] : 0.0;
      double ov = P.has_o ? Ot[pc] : 0.0;
      double v[PFH_NOFF3];
      double self;
      if constexpr (REUSE) {
        v[0] = pb[rr][0]; v[3] = pb[rr][1]; v[1] = pb[rr][2]; v[7] = pb[rr][3];
        v[2] = S1[c - 1]; v[4] = S1[c - W]; v[8] = S1[c - W - 1];
        v[6] = pa[rr][0]; v[10] = pa[rr][1]; v[12] = pa[rr][2]; v[14] = pa[rr][3];
        v[5] = S2[c]; v[9] = S2[c + W]; v[11] = S2[c + 1]; v[13] = S2[c + W + 1];
        // the march: plane k becomes plane k-1, plane k+1 becomes plane k
        pa[rr][0] = v[0]; pa[rr][1] = v[4]; pa[rr][2] = v[2]; pa[rr][3] = v[8];
        pb[rr][0] = v[5]; pb[rr][1] = v[9]; pb[rr][2] = v[11]; pb[rr][3] = v[13];
      } else {
        const double* S0 = ztile(k);
        v[0] = S1[c]; v[1] = S1[c + 1]; v[2] = S1[c - 1]; v[3] = S1[c + W]; v[4] = S1[c - W];
        v[5] = S2[c]; v[6] = S0[c]; v[7] = S1[c + W + 1]; v[8] = S1[c - W - 1];
        v[9] = S2[c + W]; v[10] = S0[c - W]; v[11] = S2[c + 1]; v[12] = S0[c - 1];
        v[13] = S2[c + W + 1]; v[14] = S0[c - W - 1];
      }
      self = v[0];
      const int type = tzt * 9 + txy[rr];
      if (MODE == 2 && P.ob0 != 0.0) ov = (P.ob0 * P.invd.v[type]) * rv;  // the first iterate, exactly as MODE 3 stores it
      double Av, bd;
      double rc = 0.0, s1 = 0.0;  // MODE 3: r at this vertex; uniform operand scale b0 / diag when every neighbour is an interior row
      if constexpr (MODE == 3) {
        rc = v[0];
        if (deep_z && deep_xy[rr]) {
          s1 = sc[rr][0];
        } else if (deep_z) {
          // next to an x or y boundary, away from the z boundaries: the per-thread scales of the 7 x-y neighbours
          constexpr int J7[PFH_NOFF3] = {0, 1, 2, 3, 4, 0, 0, 5, 6, 3, 4, 1, 2, 5, 6};
#pragma unroll
          for (int o = 0; o < PFH_NOFF3; ++o) v[o] *= sc[rr][J7[o]];
          s1 = 1.0;
        } else {
          double t[PFH_NOFF3];
#pragma unroll
          for (int o = 0; o < PFH_NOFF3; ++o) t[o] = v[o];
          first2_scale_boundary(t, P.invd, P.b0, ix, y0 + ly, gz, m);
#pragma unroll
          for (int o = 0; o < PFH_NOFF3; ++o) v[o] = t[o];
          s1 = 1.0;
        }
      }
      if (type == 13) {
        // four independent partial sums: the 15 fused multiply-adds are not one dependent chain
        double s0 = P.cI.c[0] * v[0], t1 = P.cI.c[1] * v[1], t2 = P.cI.c[2] * v[2], t3 = P.cI.c[3] * v[3];
        s0 += P.cI.c[4] * v[4]; t1 += P.cI.c[5] * v[5]; t2 += P.cI.c[6] * v[6]; t3 += P.cI.c[7] * v[7];
        s0 += P.cI.c[8] * v[8]; t1 += P.cI.c[9] * v[9]; t2 += P.cI.c[10] * v[10]; t3 += P.cI.c[11] * v[11];
        s0 += P.cI.c[12] * v[12]; t1 += P.cI.c[13] * v[13]; t2 += P.cI.c[14] * v[14];
        Av = (s0 + t1) + (t2 + t3);
        bd = P.bdI;
      } else {
        vtx::Row3 row;
        vtx::get_row3(m, type, row);
        double Ms, Ks;
        vtx::mk_sums3(row, v, Ms, Ks);
        Av = P.am * Ms + P.ak * Ks;
        bd = MODE >= 2 ? P.b / (P.am * row.m[0] + P.ak * row.k[0]) : 0.0;
      }
      if (valid[rr]) {
        if (MODE == 0) {
          P.out[idx[rr]] = P.a * rv + Av;
          if (P.out1 != nullptr) P.out1[idx[rr]] = self;
        } else if (MODE == 2) {
          P.out[idx[rr]] = self + P.a * (self - ov) + bd * (rv - Av);
        } else {
          const double z1 = s1 * v[0];
          if (P.out1 != nullptr) P.out1[idx[rr]] = z1;
          P.out[idx[rr]] = z1 + P.a * z1 + bd * (rc - s1 * Av);
        }
      }
      idx[rr] += m.nxy;
    }
  }
}

constexpr int S_TX = 64, S_TY = 4, S_RY = 1, S_NS = 4;   // 256 threads, 4 CTAs per SM: the fastest of the variants at 512^3

template <int TX, int TY, int NS>
constexpr size_t scalar_smem() {
  return NS * round_up((TX + 4) * (TY + 2) * 8, 128) + 3 * 2 * TX * TY * 8 + (NS + 3) * sizeof(uint64_t);
}

template <int MODE, int TX, int TY, int RY, int NS>
void scalar_set_attr() {
  PFH_CUDA(cudaFuncSetAttribute(k3_scalar_tma<MODE, TX, TY, RY, NS, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)scalar_smem<TX, TY, NS>()));
  PFH_CUDA(cudaFuncSetAttribute(k3_scalar_tma<MODE, TX, TY, RY, NS, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)scalar_smem<TX, TY, NS>()));
}

template <int TX, int TY, int RY, int NS>
void scalar_launch(pfh_ctx* c, const ScalarOp& op, ScalarParams& P, bool use_tma) {
  const Mesh3& m = P.m;
  const dim3 grid((m.nxv + TX - 1) / TX, (m.nyv + TY - 1) / TY, (P.nzo + P.zchunk - 1) / P.zchunk);
  const dim3 block(TX, TY / RY);
  constexpr size_t smem = scalar_smem<TX, TY, NS>();
  ScalarMaps M;
  if (use_tma) {
    M.main = pfh_tensor_map(c, op.in, m.nxv, m.nyv, op.in_n, TX + 4, TY + 2);
    M.lo = op.n_lo > 0 ? pfh_tensor_map(c, op.in_lo, m.nxv, m.nyv, op.n_lo, TX + 4, TY + 2) : M.main;
    M.hi = op.n_hi > 0 ? pfh_tensor_map(c, op.in_hi, m.nxv, m.nyv, op.n_hi, TX + 4, TY + 2) : M.main;
    M.r = P.has_r ? pfh_tensor_map(c, op.r, m.nxv, m.nyv, P.r_n, TX, TY) : M.main;
    M.r_lo = (P.has_r && P.r_nlo > 0) ? pfh_tensor_map(c, op.r_lo, m.nxv, m.nyv, P.r_nlo, TX, TY) : M.r;
    M.r_hi = (P.has_r && P.r_nhi > 0) ? pfh_tensor_map(c, op.r_hi, m.nxv, m.nyv, P.r_nhi, TX, TY) : M.r;
    M.o = P.has_o ? pfh_tensor_map(c, op.o, m.nxv, m.nyv, op.nzo, TX, TY) : M.main;
  } else {
    memset(&M, 0, sizeof M);
  }
#define PFH_SCALAR_GO(MODE)                                                                        \
  if (use_tma) k3_scalar_tma<MODE, TX, TY, RY, NS, true><<<grid, block, smem, c->stream>>>(M, P);  \
  else k3_scalar_tma<MODE, TX, TY, RY, NS, false><<<grid, block, smem, c->stream>>>(M, P)
  if (op.mode == 0) { PFH_SCALAR_GO(0); }
  else if (op.mode == 2) { PFH_SCALAR_GO(2); }
  else { PFH_SCALAR_GO(3); }
#undef PFH_SCALAR_GO
  PFH_CUDA(cudaGetLastError());
}

}  // namespace

// ---- tensor-map cache ----
namespace {
struct TmapEntry {
  const void* p;
  int nxv, nyv, nz, bx, by;
  CUtensorMap map;
};
struct TmapCache {
  std::vector<TmapEntry> e;
};
}  // namespace

CUtensorMap pfh_tensor_map(pfh_ctx* c, const double* ptr, int nxv, int nyv, int nz, int bx, int by) {
  if (c->tmap_cache == nullptr) c->tmap_cache = new TmapCache;
  TmapCache* tc = static_cast<TmapCache*>(c->tmap_cache);
  for (const TmapEntry& t : tc->e)
    if (t.p == ptr && t.nz == nz && t.bx == bx && t.by == by && t.nxv == nxv && t.nyv == nyv) return t.map;
  if (tc->e.size() >= 512) tc->e.clear();  // caller-owned buffers may come and go
  TmapEntry t{ptr, nxv, nyv, nz, bx, by, tma::make_map(ptr, nxv, nyv, nz, bx, by)};
  tc->e.push_back(t);
  return t.map;
}

void pfh_tmap_cache_free(pfh_ctx* c) {
  delete static_cast<TmapCache*>(c->tmap_cache);
  c->tmap_cache = nullptr;
}

// dynamic shared memory opt-in of every instantiation, for the device the ctx lives on (the attribute is per device)
void pfh_scalar3_init_device() {
  scalar_set_attr<0, S_TX, S_TY, S_RY, S_NS>();
  scalar_set_attr<2, S_TX, S_TY, S_RY, S_NS>();
  scalar_set_attr<3, S_TX, S_TY, S_RY, S_NS>();
  scalar_set_attr<0, 64, 8, 2, 4>();
  scalar_set_attr<2, 64, 8, 2, 4>();
  scalar_set_attr<3, 64, 8, 2, 4>();
  scalar_set_attr<0, 64, 8, 1, 3>();
  scalar_set_attr<2, 64, 8, 1, 3>();
  scalar_set_attr<3, 64, 8, 1, 3>();
  scalar_set_attr<0, 64, 8, 1, 4>();
  scalar_set_attr<2, 64, 8, 1, 4>();
  scalar_set_attr<3, 64, 8, 1, 4>();
  scalar_set_attr<0, 64, 4, 1, 5>();
  scalar_set_attr<2, 64, 4, 1, 5>();
  scalar_set_attr<3, 64, 4, 1, 5>();
  scalar_set_attr<0, 64, 2, 1, 4>();
  scalar_set_attr<2, 64, 2, 1, 4>();
  scalar_set_attr<3, 64, 2, 1, 4>();
  scalar_set_attr<0, 128, 4, 1, 4>();
  scalar_set_attr<2, 128, 4, 1, 4>();
  scalar_set_attr<3, 128, 4, 1, 4>();
}

bool pfh_fast3_ok(const pfh_ctx* c) { return c->dim == 3 && c->use_tma; }

void k3_scalar_launch(pfh_ctx* c, const ScalarOp& op) {
  PFH_CHECK(op.mode == 0 || op.mode == 2 || op.mode == 3, "scalar operator: unknown mode");
  PFH_CHECK(op.nzo >= 1 && op.in != nullptr && op.out != nullptr, "scalar operator: empty launch");
  const Mesh3& m = c->m3;
  ScalarParams P;
  memset(&P, 0, sizeof P);
  P.m = m;
  P.gz0 = op.gz0; P.nzo = op.nzo;
  P.in_off = op.in_off; P.in_n = op.in_n;
  P.n_lo = op.in_lo ? op.n_lo : 0; P.n_hi = op.in_hi ? op.n_hi : 0;
  P.has_r = (op.mode == 2 || (op.mode == 0 && op.a != 0.0)) && op.r != nullptr;
  P.has_o = op.mode == 2 && op.o != nullptr;
  if (P.has_r) {
    P.r_off = op.r_off;
    P.r_n = op.r_n > 0 ? op.r_n : op.nzo;  // default: r points at launch-local plane 0 and covers the window
    P.r_nlo = op.r_lo ? op.n_rlo : 0;
    P.r_nhi = op.r_hi ? op.n_rhi : 0;
  }
  P.am = op.am; P.ak = op.ak; P.a = op.a; P.b = op.b; P.b0 = op.b0;
  for (int o = 0; o < PFH_NOFF3; ++o) P.cI.c[o] = op.am * m.mI[o] + (o < PFH_NKOFF3 ? op.ak * m.kI[o] : 0.0);
  P.bdI = op.b / P.cI.c[0];
  P.ob0 = (op.mode == 2 && op.o == nullptr) ? op.ob0 : 0.0;
  if (op.mode == 3 || P.ob0 != 0.0) {
    if (c->h_mT.empty()) {  // host copy of the type tables, for the inverse diagonals
      Mesh3 mm = m;
      pfh_build_tables3(mm, c->h_mT, c->h_kT);
    }
    for (int t = 0; t < 27; ++t) P.invd.v[t] = 1.0 / (op.am * c->h_mT[t * PFH_NOFF3] + op.ak * c->h_kT[t * PFH_NKOFF3]);
  }
  P.out = op.out; P.out1 = op.out1;
  P.in = op.in; P.in_lo = op.in_lo; P.in_hi = op.in_hi; P.r = op.r; P.r_lo = op.r_lo; P.r_hi = op.r_hi; P.o = op.o;
  // planes per CTA: balanced chunks of at most `scalar_zchunk` planes
  const int target = c->scalar_zchunk > 0 ? c->scalar_zchunk : 32;
  const int nch = (op.nzo + target - 1) / target;
  P.zchunk = (op.nzo + nch - 1) / nch;
  // TMA needs 16-byte aligned rows and plane starts: an even vertex count in x and aligned base pointers
  bool use_tma = c->use_tma_loader && (m.nxv % 2) == 0 && aligned16(op.in) && aligned16(op.out);
  if (P.n_lo > 0) use_tma = use_tma && aligned16(op.in_lo);
  if (P.n_hi > 0) use_tma = use_tma && aligned16(op.in_hi);
  if (P.has_r) use_tma = use_tma && aligned16(op.r) && (P.r_nlo == 0 || aligned16(op.r_lo)) && (P.r_nhi == 0 || aligned16(op.r_hi));
  if (P.has_o) use_tma = use_tma && aligned16(op.o);
  switch (c->scalar_variant) {
    case 1: scalar_launch<64, 8, 2, 4>(c, op, P, use_tma); break;   // two rows per thread: spills at 64 registers
    case 2: scalar_launch<64, 8, 1, 3>(c, op, P, use_tma); break;
    case 3: scalar_launch<64, 8, 1, 4>(c, op, P, use_tma); break;
    case 4: scalar_launch<128, 4, 1, 4>(c, op, P, use_tma); break;
    case 5: scalar_launch<64, 4, 1, 5>(c, op, P, use_tma); break;
    case 6: scalar_launch<64, 2, 1, 4>(c, op, P, use_tma); break;
    default: scalar_launch<S_TX, S_TY, S_RY, S_NS>(c, op, P, use_tma); break;
  }
}
