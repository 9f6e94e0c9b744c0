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

// One CTA = TX*TY consumer threads (one vertex column each) + one producer warp.  The producer fills a ring of NS stages
// -- each stage holds one plane of the stencil operand (tile + halo) and the r / zprev tiles of the output plane two
// below it, i.e. everything that becomes necessary in the same iteration -- and signals full[stage]; a consumer warp
// that has finished with a stage arrives on empty[stage].  There is no block-wide barrier in the z-march: warps drift
// apart by up to NS-2 planes, which hides the latency of the loads.  The march is unrolled NS-fold so that stage
// addresses and barrier parities are compile-time constants and the register rotation of the z-reuse needs no moves.
template <int MODE, int TX, int TY, int NS, int MINB, bool TMA>
__global__ void __launch_bounds__(TX* TY + 32, (MODE == 3 && MINB > 2) ? MINB - 1 : MINB)  // MODE 3 needs ~100 registers
    k3_scalar_tma(const __grid_constant__ ScalarMaps M, const __grid_constant__ ScalarParams P) {
  constexpr int W = TX + 4, H = TY + 2;  // x halo of 2: the TMA box start (x0-2) must be 16-byte aligned
  constexpr int NC = TX * TY, NCW = NC / 32;  // consumer threads / warps
  constexpr int ZT_BYTES = W * H * 8, PT_BYTES = TX * TY * 8;
  constexpr int ZT_STRIDE = round_up(ZT_BYTES, 128);
  constexpr int STAGE = ZT_STRIDE + 2 * PT_BYTES;  // stencil tile | r tile | zprev tile
  constexpr bool REUSE = MODE != 3;                // z-march register reuse (MODE 3 spends the registers on its operand scales)
  static_assert(NC % 32 == 0 && NS >= 3, "tile must be whole warps; the march reads three planes");
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + NS * STAGE);
  uint64_t* empty = full + NS;

  const Mesh3& m = P.m;
  const int tid = threadIdx.x;
  const int x0 = blockIdx.x * TX, y0 = blockIdx.y * TY;
  const int zb = blockIdx.z * P.zchunk, ze = min(P.nzo, zb + P.zchunk);
  const int nout = ze - zb, nplanes = nout + 2;  // input planes zb-1 .. ze  <->  p = 0 .. nplanes-1

  if (tid == 0) {
    for (int s = 0; s < NS; ++s) {
      mbar_init(&full[s], TMA ? 1 : 32);
      mbar_init(&empty[s], NCW);
    }
    fence_barrier_init();
  }
  __syncthreads();

  if (tid >= NC) {
    // ------------------------------------------------ producer warp ------------------------------------------------
    const int lane = tid - NC;
    if (TMA && lane != 0) return;
    for (int p = 0; p < nplanes; ++p) {
      const int s = p % NS;
      if (p >= NS) mbar_wait(&empty[s], ((p / NS) - 1) & 1);
      unsigned char* dst = smem + s * STAGE;
      const int zm = zb - 1 + p + P.in_off;  // index of the plane in the main array of the stencil operand
      const int k = p - 2;                   // output plane whose pointwise operands travel with this stage
      const bool ro = k >= 0 && (P.has_r || P.has_o);
      const int zr = zb + k + P.r_off;       // index of output plane k in the main array of r
      if constexpr (TMA) {
        const CUtensorMap* map = &M.main;
        int zc = zm;  // outside every array: beyond the physical boundary (or a plane nobody needs) -> zero fill
        if (zm < 0 && P.n_lo > 0) {
          map = &M.lo;
          zc = zm + P.n_lo;
        } else if (zm >= P.in_n && P.n_hi > 0) {
          map = &M.hi;
          zc = zm - P.in_n;
        }
        mbar_expect_tx(&full[s], ZT_BYTES + (ro ? (P.has_r ? PT_BYTES : 0) + (P.has_o ? PT_BYTES : 0) : 0));
        tma_load_3d(dst, map, x0 - 2, y0 - 1, zc, &full[s]);
        if (ro && P.has_r) {
          const CUtensorMap* rm = &M.r;
          int rc = zr;
          if (zr < 0 && P.r_nlo > 0) {
            rm = &M.r_lo;
            rc = zr + P.r_nlo;
          } else if (zr >= P.r_n && P.r_nhi > 0) {
            rm = &M.r_hi;
            rc = zr - P.r_n;
          }
          tma_load_3d(dst + ZT_STRIDE, rm, x0, y0, rc, &full[s]);
        }
        if (ro && P.has_o) tma_load_3d(dst + ZT_STRIDE + PT_BYTES, &M.o, x0, y0, zb + k, &full[s]);
      } else {
        const double* src = nullptr;
        if (zm >= 0 && zm < P.in_n) src = P.in + (long long)zm * m.nxy;
        else if (zm < 0 && zm + P.n_lo >= 0) src = P.in_lo + (long long)(zm + P.n_lo) * m.nxy;
        else if (zm >= P.in_n && zm - P.in_n < P.n_hi) src = P.in_hi + (long long)(zm - P.in_n) * m.nxy;
        for (int i = lane; i < W * H; i += 32) {
          const int gx = x0 - 2 + i % W, gy = y0 - 1 + i / W;
          const bool ok = src != nullptr && gx >= 0 && gx < m.nxv && gy >= 0 && gy < m.nyv;
          cp_async8(dst + i * 8, ok ? src + (long long)gy * m.nxv + gx : P.out, ok ? 8 : 0);
        }
        if (ro) {
          const double* rs = nullptr;
          if (P.has_r) {
            if (zr >= 0 && zr < P.r_n) rs = P.r + (long long)zr * m.nxy;
            else if (zr < 0 && zr + P.r_nlo >= 0) rs = P.r_lo + (long long)(zr + P.r_nlo) * m.nxy;
            else if (zr >= P.r_n && zr - P.r_n < P.r_nhi) rs = P.r_hi + (long long)(zr - P.r_n) * m.nxy;
          }
          const long long pl = (long long)(zb + k) * m.nxy;
          for (int i = lane; i < NC; i += 32) {
            const int gx = x0 + i % TX, gy = y0 + i / TX;
            const bool ok = gx < m.nxv && gy < m.nyv;
            const long long g = (long long)gy * m.nxv + gx;
            if (P.has_r) cp_async8(dst + ZT_STRIDE + i * 8, (ok && rs) ? rs + g : P.out, (ok && rs) ? 8 : 0);
            if (P.has_o) cp_async8(dst + ZT_STRIDE + PT_BYTES + i * 8, ok ? P.o + pl + g : P.out, ok ? 8 : 0);
          }
        }
        cp_async_mbar_arrive(&full[s]);
      }
    }
    return;
  }

  // -------------------------------------------------- consumer warps --------------------------------------------------
  const int lx = tid % TX, ly = tid / TX;
  const int ix = x0 + lx, iy = y0 + ly;
  const bool valid = ix < m.nxv && iy < m.nyv;
  const bool deep_xy = ix >= 2 && ix <= m.nxv - 3 && iy >= 2 && iy <= m.nyv - 3;
  const int txy = (iy == 0 ? 0 : (iy == m.nyv - 1 ? 2 : 1)) * 3 + (ix == 0 ? 0 : (ix == m.nxv - 1 ? 2 : 1));
  const int c = (ly + 1) * W + lx + 2;  // this vertex in a stencil tile
  const int pc = ly * TX + lx;          // and in a pointwise tile
  double* outp = P.out + ((long long)zb * m.nxy + (long long)iy * m.nxv + ix);
  double* out1p = P.out1 ? P.out1 + ((long long)zb * m.nxy + (long long)iy * m.nxv + ix) : nullptr;
  const bool elect = (tid & 31) == 0;
  auto ztile = [&](int s) { return reinterpret_cast<const double*>(smem + s * STAGE); };
  auto release = [&](int s) {  // this warp is done with stage s
    __syncwarp();
    if (elect) mbar_arrive(&empty[s]);
  };

  double pa[4], pb[4];  // REUSE: plane k-1: c, c-W, c-1, c-W-1; plane k: c, c+W, c+1, c+W+1
  // MODE 3: scale b0 / diag of the row of each of the 7 distinct x-y neighbours (0,0) (1,0) (-1,0) (0,1) (0,-1) (1,1) (-1,-1)
  // on a plane away from the z boundaries; fixed per thread, so the vertices next to an x or y boundary cost 15
  // multiplications instead of a table walk per plane
  double sc[MODE == 3 ? 7 : 1];
  if constexpr (MODE == 3) {
    constexpr int DX7[7] = {0, 1, -1, 0, 0, 1, -1}, DY7[7] = {0, 0, 0, 1, -1, 1, -1};
    auto cls = [](int i, int n) { return i == 0 ? 0 : (i == n - 1 ? 2 : 1); };
#pragma unroll
    for (int j = 0; j < 7; ++j) sc[j] = P.b0 * P.invd.v[9 + 3 * cls(iy + DY7[j], m.nyv) + cls(ix + DX7[j], m.nxv)];
  }
  mbar_wait(&full[0], 0);
  mbar_wait(&full[1], 0);
  if constexpr (REUSE) {
    const double* S0 = ztile(0);
    const double* S1 = ztile(1);
    pa[0] = S0[c]; pa[1] = S0[c - W]; pa[2] = S0[c - 1]; pa[3] = S0[c - W - 1];
    pb[0] = S1[c]; pb[1] = S1[c + W]; pb[2] = S1[c + 1]; pb[3] = S1[c + W + 1];
    release(0);
  }

  uint32_t ph = 0;  // parity of the ring pass the iteration's base plane k0 belongs to
  for (int k0 = 0; k0 < nout; k0 += NS, ph ^= 1) {
#pragma unroll
    for (int u = 0; u < NS; ++u) {
      const int k = k0 + u;
      if (k >= nout) break;
      const int s1i = (u + 1) % NS, s2i = (u + 2) % NS;
      mbar_wait(&full[s2i], ph ^ ((u + 2) / NS));  // plane k+2 and the pointwise tiles of output plane k
      const double* S1 = ztile(s1i);
      const double* S2 = ztile(s2i);
      const double* Rt = reinterpret_cast<const double*>(smem + s2i * STAGE + ZT_STRIDE);
      const double* Ot = Rt + NC;
      const int gz = P.gz0 + zb + k;
      const int tzt = gz == 0 ? 0 : (gz == m.nzv - 1 ? 2 : 1);
      const bool deep_z = gz >= 2 && gz <= m.nzv - 3;
      const double rv = P.has_r ? Rt[pc] : 0.0;
      double ov = P.has_o ? Ot[pc] : 0.0;
      double v[PFH_NOFF3];
      if constexpr (REUSE) {
        v[0] = pb[0]; v[3] = pb[1]; v[1] = pb[2]; v[7] = pb[3];
        v[2] = S1[c - 1]; v[4] = S1[c - W]; v[8] = S1[c - W - 1];
        v[6] = pa[0]; v[10] = pa[1]; v[12] = pa[2]; v[14] = pa[3];
        v[5] = S2[c]; v[9] = S2[c + W]; v[11] = S2[c + 1]; v[13] = S2[c + W + 1];
        // the march: plane k becomes plane k-1, plane k+1 becomes plane k
        pa[0] = v[0]; pa[1] = v[4]; pa[2] = v[2]; pa[3] = v[8];
        pb[0] = v[5]; pb[1] = v[9]; pb[2] = v[11]; pb[3] = v[13];
        release(s1i);  // the rest of plane k+1 lives in registers
      } else {
        const double* S0 = ztile(u % NS);
        v[0] = S1[c]; v[1] = S1[c + 1]; v[2] = S1[c - 1]; v[3] = S1[c + W]; v[4] = S1[c - W];
        v[5] = S2[c]; v[6] = S0[c]; v[7] = S1[c + W + 1]; v[8] = S1[c - W - 1];
        v[9] = S2[c + W]; v[10] = S0[c - W]; v[11] = S2[c + 1]; v[12] = S0[c - 1];
        v[13] = S2[c + W + 1]; v[14] = S0[c - W - 1];
        release(u % NS);
      }
      const double self = v[0];
      const int type = tzt * 9 + txy;
      if (MODE == 2 && P.ob0 != 0.0) ov = (P.ob0 * P.invd.v[type]) * rv;  // the first iterate, exactly as MODE 3 stores it
      double Av, bd;
      double rc = 0.0, s1 = 0.0;  // MODE 3: r at this vertex; uniform operand scale b0 / diag when every neighbour is an interior row
      if constexpr (MODE == 3) {
        rc = v[0];
        if (deep_z && deep_xy) {
          s1 = sc[0];
        } else if (deep_z) {
          // next to an x or y boundary, away from the z boundaries: the per-thread scales of the 7 x-y neighbours
          constexpr int J7[PFH_NOFF3] = {0, 1, 2, 3, 4, 0, 0, 5, 6, 3, 4, 1, 2, 5, 6};
#pragma unroll
          for (int o = 0; o < PFH_NOFF3; ++o) v[o] *= sc[J7[o]];
          s1 = 1.0;
        } else {
          double t[PFH_NOFF3];
#pragma unroll
          for (int o = 0; o < PFH_NOFF3; ++o) t[o] = v[o];
          first2_scale_boundary(t, P.invd, P.b0, ix, iy, gz, m);
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
      if (valid) {
        if (MODE == 0) {
          *outp = P.a * rv + Av;
          if (out1p != nullptr) *out1p = self;
        } else if (MODE == 2) {
          *outp = self + P.a * (self - ov) + bd * (rv - Av);
        } else {
          const double z1 = s1 * v[0];
          if (out1p != nullptr) *out1p = z1;
          *outp = z1 + P.a * z1 + bd * (rc - s1 * Av);
        }
      }
      outp += m.nxy;
      if (out1p != nullptr) out1p += m.nxy;
    }
  }
}

constexpr int S_TX = 64, S_TY = 4, S_NS = 4, S_MINB = 3;

template <int TX, int TY, int NS>
constexpr size_t scalar_smem() {
  return NS * (round_up((TX + 4) * (TY + 2) * 8, 128) + 2 * TX * TY * 8) + 2 * NS * sizeof(uint64_t);
}

template <int MODE, int TX, int TY, int NS, int MINB>
void scalar_set_attr() {
  PFH_CUDA(cudaFuncSetAttribute(k3_scalar_tma<MODE, TX, TY, NS, MINB, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)scalar_smem<TX, TY, NS>()));
  PFH_CUDA(cudaFuncSetAttribute(k3_scalar_tma<MODE, TX, TY, NS, MINB, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)scalar_smem<TX, TY, NS>()));
}
template <int TX, int TY, int NS, int MINB>
void scalar_set_attr_all() {
  scalar_set_attr<0, TX, TY, NS, MINB>();
  scalar_set_attr<2, TX, TY, NS, MINB>();
  scalar_set_attr<3, TX, TY, NS, MINB>();
}

template <int TX, int TY, int NS, int MINB>
void scalar_launch(pfh_ctx* c, const ScalarOp& op, ScalarParams& P, bool use_tma) {
  const Mesh3& m = P.m;
  const dim3 grid((m.nxv + TX - 1) / TX, (m.nyv + TY - 1) / TY, (P.nzo + P.zchunk - 1) / P.zchunk);
  const dim3 block(TX * TY + 32);
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
#define PFH_SCALAR_GO(MODE)                                                                              \
  if (use_tma) k3_scalar_tma<MODE, TX, TY, NS, MINB, true><<<grid, block, smem, c->stream>>>(M, P);      \
  else k3_scalar_tma<MODE, TX, TY, NS, MINB, false><<<grid, block, smem, c->stream>>>(M, P)
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
  scalar_set_attr_all<S_TX, S_TY, S_NS, S_MINB>();
  scalar_set_attr_all<64, 4, 6, 3>();
  scalar_set_attr_all<64, 4, 4, 4>();
  scalar_set_attr_all<64, 8, 4, 2>();
  scalar_set_attr_all<64, 2, 6, 6>();
  scalar_set_attr_all<128, 2, 5, 3>();
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
  // planes per CTA: balanced chunks of at most `scalar_zchunk` planes; on a short window (a slab of a multi-GPU run) the
  // chunk count is chosen so that the CTAs fill whole waves of the device (each CTA also pays ~3 planes of ramp-up)
  const int target = c->scalar_zchunk > 0 ? c->scalar_zchunk : 32;
  int nch = (op.nzo + target - 1) / target;
  if (c->scalar_zchunk <= 0 && c->n_sm > 0 && op.nzo <= 160) {
    const long long tiles = (long long)((m.nxv + 63) / 64) * ((m.nyv + 3) / 4), slots = 3LL * c->n_sm;
    double best = 1e300;
    for (int n = nch; n <= (op.nzo + 7) / 8; ++n) {
      const int ch = (op.nzo + n - 1) / n;
      const double waves = (double)((tiles * n + slots - 1) / slots);
      const double cost = waves * (ch + 3.0);
      if (cost < best - 1e-9) {
        best = cost;
        nch = n;
      }
    }
  }
  P.zchunk = (op.nzo + nch - 1) / nch;
  // TMA needs 16-byte aligned rows and plane starts: an even vertex count in x and aligned base pointers
  bool use_tma = c->use_tma_loader && (m.nxv % 2) == 0 && aligned16(op.in) && aligned16(op.out);
  if (P.n_lo > 0) use_tma = use_tma && aligned16(op.in_lo);
  if (P.n_hi > 0) use_tma = use_tma && aligned16(op.in_hi);
  if (P.has_r) use_tma = use_tma && aligned16(op.r) && (P.r_nlo == 0 || aligned16(op.r_lo)) && (P.r_nhi == 0 || aligned16(op.r_hi));
  if (P.has_o) use_tma = use_tma && aligned16(op.o);
  switch (c->scalar_variant) {
    case 1: scalar_launch<64, 4, 6, 3>(c, op, P, use_tma); break;    // deeper ring
    case 2: scalar_launch<64, 4, 4, 4>(c, op, P, use_tma); break;    // 4 CTAs per SM at 56 registers
    case 3: scalar_launch<64, 8, 4, 2>(c, op, P, use_tma); break;    // 544 threads, 2 CTAs per SM
    case 4: scalar_launch<64, 2, 6, 6>(c, op, P, use_tma); break;    // 160 threads, 6 CTAs per SM
    case 5: scalar_launch<128, 2, 5, 3>(c, op, P, use_tma); break;
    default: scalar_launch<S_TX, S_TY, S_NS, S_MINB>(c, op, P, use_tma); break;
  }
}
