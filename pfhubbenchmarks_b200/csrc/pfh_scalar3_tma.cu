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
// The staged kernel treats EVERY vertex as an interior row (constant 15-point stencil, uniform Jacobi scale): no row
// tables, no boundary-type logic in the z-march.  It skips the stores of the vertices whose result that would get
// wrong -- the boundary vertices, and for MODE 3 also their neighbours (their operand b0 D^-1 r needs the neighbour's own
// diagonal) -- and a small gather kernel (k3_scalar_shell) computes those from global memory with the general
// per-type rows: 1-2 % of the vertices, ~5 % of the time.
//
// A launch works on an arbitrary window of planes: output planes are launch-local zl = 0 .. nzo-1 (global plane
// gz0 + zl); the stencil operand is addressed through a main array plus optional lower / upper ghost arrays of any
// width.  This serves both the classic one-ghost-plane layout and the wide-ghost preconditioner of the multi-GPU
// path, where Chebyshev sweeps run redundantly on a shrinking halo instead of exchanging a plane per sweep.
#include <algorithm>
#include <unordered_map>

#include "pfh_tma.cuh"
#include "pfh_vertex.cuh"

namespace {
using namespace tma;

struct Coef15 {
  double c[PFH_NOFF3];
};
// interior row of am M + ak K by symmetric pair: self, +-x, +-y, +-z, +-(x+y), +-(y+z), +-(x+z), +-(x+y+z)
struct Coef8 {
  double c[8];
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
  double s1I, ob0I;            // b0 / diag and ob0 / diag of the interior row
  double rowsB[9][PFH_NOFF3];  // MODE 0 / 2: rows of am M + ak K by x-y boundary type (ty*3+tx) on a plane away from the z boundaries
  int shell;                   // depth of the boundary shell the staged kernel leaves to k3_scalar_shell (1; MODE 3: 2)
  Coef8 cS;
  InvDiag27 invd;
  double* out;                 // launch-local plane 0
  double* out1;                // MODE 3: z1; MODE 0: optional copy of the operand
  const double *in, *in_lo, *in_hi, *r, *r_lo, *r_hi, *o;  // raw operand pointers (CPASYNC loader); o at launch-local plane 0
};

// One CTA = TX*TY consumer threads (one vertex column each) + one producer warp.  The producer fills a ring of NS stages
// -- each stage holds one plane of the stencil operand (tile + halo) and the r / zprev tiles of the output plane two
// below it, i.e. everything that becomes necessary in the same iteration -- and signals full[stage]; a consumer warp
// that has finished with a stage arrives on empty[stage].  There is no block-wide barrier in the z-march: warps drift
// apart by up to NS-2 planes, which hides the latency of the loads.  The march is unrolled NS-fold so that stage
// addresses and barrier parities are compile-time constants and the register rotation of the z-reuse needs no moves.
template <int MODE, int TX, int TY, int NS, int MINB, bool TMA>
__global__ void __launch_bounds__(TX* TY + 32, MINB)
    k3_scalar_tma(const __grid_constant__ ScalarMaps M, const __grid_constant__ ScalarParams P) {
  constexpr int W = TX + 4, H = TY + 2;  // x halo of 2: the TMA box start (x0-2) must be 16-byte aligned
  constexpr int NC = TX * TY, NCW = NC / 32;  // consumer threads / warps
  constexpr int ZT_BYTES = W * H * 8, PT_BYTES = TX * TY * 8;
  constexpr int ZT_STRIDE = round_up(ZT_BYTES, 128);
  constexpr int STAGE = ZT_STRIDE + 2 * PT_BYTES;  // stencil tile | r tile | zprev tile
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
  const int d = P.shell;
  // Vertices this kernel stores.  MODE 3: inside the mesh and not within `shell` (2) layers of any boundary.  MODE 0 / 2: every
  // vertex of the planes away from the z boundaries -- the vertices on an x or y face take their row from rowsB (fixed per
  // thread along the march; only the warps that contain such a vertex diverge) -- k3_scalar_shell computes the rest.
  const bool store_xy = MODE == 3 ? (ix >= d && ix < m.nxv - d && iy >= d && iy < m.nyv - d) : (ix < m.nxv && iy < m.nyv);
  const int txy = (iy == 0 ? 0 : (iy == m.nyv - 1 ? 2 : 1)) * 3 + (ix == 0 ? 0 : (ix == m.nxv - 1 ? 2 : 1));
  const bool face = MODE != 3 && txy != 4 && store_xy;
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

  // z-march register reuse: plane k-1: c, c-W, c-1, c-W-1; plane k: c, c+W, c+1, c+W+1
  double pa[4], pb[4];
  mbar_wait(&full[0], 0);
  mbar_wait(&full[1], 0);
  {
    const double* S0 = ztile(0);
    const double* S1 = ztile(1);
    pa[0] = S0[c]; pa[1] = S0[c - W]; pa[2] = S0[c - 1]; pa[3] = S0[c - W - 1];
    pb[0] = S1[c]; pb[1] = S1[c + W]; pb[2] = S1[c + 1]; pb[3] = S1[c + W + 1];
    release(0);
  }

  // one output plane: `u` = k mod NS (compile-time in the unrolled groups), `ph` = parity of ring pass k / NS
  auto plane = [&](const int k, const int u, const uint32_t ph) {
    const int s1i = (u + 1) % NS, s2i = (u + 2) % NS;
    mbar_wait(&full[s2i], ph ^ ((u + 2) / NS));  // plane k+2 and the pointwise tiles of output plane k
    const double* S1 = ztile(s1i);
    const double* S2 = ztile(s2i);
    const double* Rt = reinterpret_cast<const double*>(smem + s2i * STAGE + ZT_STRIDE);
    const double* Ot = Rt + NC;
    const int gz = P.gz0 + zb + k;
    const bool store = store_xy && gz >= d && gz < m.nzv - d;
    double bd = P.bdI;
    const double rv = P.has_r ? Rt[pc] : 0.0;
    double ov = P.has_o ? Ot[pc] : 0.0;
    const double v0 = pb[0], v3 = pb[1], v1 = pb[2], v7 = pb[3];
    const double v2 = S1[c - 1], v4 = S1[c - W], v8 = S1[c - W - 1];
    const double v6 = pa[0], v10 = pa[1], v12 = pa[2], v14 = pa[3];
    const double v5 = S2[c], v9 = S2[c + W], v11 = S2[c + 1], v13 = S2[c + W + 1];
    // the march: plane k becomes plane k-1, plane k+1 becomes plane k
    pa[0] = v0; pa[1] = v4; pa[2] = v2; pa[3] = v8;
    pb[0] = v5; pb[1] = v9; pb[2] = v11; pb[3] = v13;
    release(s1i);  // the rest of plane k+1 lives in registers
    // the interior row is symmetric: one coefficient per +- pair (8 constants), two partial sums
    double s0 = P.cS.c[0] * v0;
    double t1 = P.cS.c[1] * (v1 + v2);
    s0 += P.cS.c[2] * (v3 + v4);
    t1 += P.cS.c[3] * (v5 + v6);
    s0 += P.cS.c[4] * (v7 + v8);
    t1 += P.cS.c[5] * (v9 + v10);
    s0 += P.cS.c[6] * (v11 + v12);
    t1 += P.cS.c[7] * (v13 + v14);
    double Av = s0 + t1;
    if (MODE != 3 && face) {
      // a vertex on an x or y face (or a z-parallel edge): its own row, out-of-domain operands are the zeros TMA filled in
      const double* row = P.rowsB[txy];
      double a0 = row[0] * v0 + row[1] * v1, a1 = row[2] * v2 + row[3] * v3;
      a0 += row[4] * v4; a1 += row[5] * v5; a0 += row[6] * v6; a1 += row[7] * v7;
      a0 += row[8] * v8; a1 += row[9] * v9; a0 += row[10] * v10; a1 += row[11] * v11;
      a0 += row[12] * v12; a1 += row[13] * v13; a0 += row[14] * v14;
      Av = a0 + a1;
      bd = P.b * P.invd.v[9 + txy];
      if (MODE == 2 && P.ob0 != 0.0) ov = (P.ob0 * P.invd.v[9 + txy]) * rv;
    } else if (MODE == 2 && P.ob0I != 0.0) {
      ov = P.ob0I * rv;  // the first iterate recomputed from r, exactly as MODE 3 would store it
    }
    if (store) {
      if (MODE == 0) {
        *outp = P.a * rv + Av;
        if (out1p != nullptr) *out1p = v0;
      } else if (MODE == 2) {
        *outp = v0 + P.a * (v0 - ov) + bd * (rv - Av);
      } else {
        const double z1 = P.s1I * v0;
        if (out1p != nullptr) *out1p = z1;
        *outp = z1 + P.a * z1 + P.bdI * (v0 - P.s1I * Av);
      }
    }
    outp += m.nxy;
    if (out1p != nullptr) out1p += m.nxy;
  };
  // full groups of NS planes, unrolled: stage addresses and parities are constants and the z-reuse registers rotate by
  // renaming; then the remaining planes one by one
  int k0 = 0;
  uint32_t ph = 0;
  for (; k0 + NS <= nout; k0 += NS, ph ^= 1) {
#pragma unroll
    for (int u = 0; u < NS; ++u) plane(k0 + u, u, ph);
  }
  for (int k = k0; k < nout; ++k) plane(k, k - k0, ph);
}

// ---------------------------------------------------------------------------------------------
// Boundary shell: the vertices within `shell` layers of the mesh boundary, computed from global memory with the general
// per-type rows (vtx::get_row3) and, for MODE 3, the operand scaled by each neighbour's own b0 / diag.  One thread per
// shell vertex; the index space is [z-shell planes: whole planes | other planes: y-shell rows | x-shell columns].
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int shell_index(int j, int n, int d) { return (2 * d >= n || j < d) ? j : n - 2 * d + j; }

// plane `z` (index in the main array, may lie in a ghost array) of an operand addressed like the stencil operand
__device__ __forceinline__ const double* operand_plane(const double* main, int n_main, const double* lo, int n_lo, const double* hi,
                                                       int n_hi, int z, long long nxy) {
  if (z >= 0 && z < n_main) return main + (long long)z * nxy;
  if (z < 0 && z + n_lo >= 0) return lo + (long long)(z + n_lo) * nxy;
  if (z >= n_main && z - n_main < n_hi) return hi + (long long)(z - n_main) * nxy;
  return nullptr;
}

template <int MODE>
__global__ void __launch_bounds__(128) k3_scalar_shell(const __grid_constant__ ScalarParams P, int nzs, int zs0, int zs1, int zs2, int zs3,
                                                       int zi0, int nzi, long long n_threads) {
  const Mesh3& m = P.m;
  const int d = P.shell;
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= n_threads) return;
  const int csx = min(2 * d, m.nxv), csy = min(2 * d, m.nyv);
  const long long n1 = (long long)nzs * m.nxy, n2 = (long long)nzi * csy * m.nxv;
  int ix, iy, zl;
  if (t < n1) {  // a plane of the z shell
    const int j = (int)(t / m.nxy);
    const long long r = t - (long long)j * m.nxy;
    zl = j == 0 ? zs0 : (j == 1 ? zs1 : (j == 2 ? zs2 : zs3));
    iy = (int)(r / m.nxv);
    ix = (int)(r - (long long)iy * m.nxv);
  } else if (t < n1 + n2) {  // a row of the y shell
    const long long r = t - n1;
    const long long per = (long long)csy * m.nxv;
    zl = zi0 + (int)(r / per);
    const long long q = r % per;
    iy = shell_index((int)(q / m.nxv), m.nyv, d);
    ix = (int)(q % m.nxv);
  } else {  // a column of the x shell, rows outside the y shell
    const long long r = t - n1 - n2;
    const int ny_in = m.nyv - csy;
    const long long per = (long long)ny_in * csx;
    zl = zi0 + (int)(r / per);
    const long long q = r % per;
    iy = d + (int)(q / csx);
    ix = shell_index((int)(q % csx), m.nxv, d);
  }
  const int gz = P.gz0 + zl;
  const double* pl[3];
#pragma unroll
  for (int dz = -1; dz <= 1; ++dz) {
    const int g = gz + dz;
    pl[dz + 1] = (g < 0 || g >= m.nzv) ? nullptr : operand_plane(P.in, P.in_n, P.in_lo, P.n_lo, P.in_hi, P.n_hi, zl + dz + P.in_off, m.nxy);
  }
  const vtx::Planes3 PL{pl[0], pl[1], pl[2]};
  double v[PFH_NOFF3];
  vtx::gather15<true>(PL, m, ix, iy, v);
  const int type = vtx::type3(m, ix, iy, gz);
  const long long i = (long long)zl * m.nxy + (long long)iy * m.nxv + ix;
  double rv = 0.0, ov = 0.0;
  if (P.has_r) {
    const double* rp = operand_plane(P.r, P.r_n, P.r_lo, P.r_nlo, P.r_hi, P.r_nhi, zl + P.r_off, m.nxy);
    rv = rp ? rp[(long long)iy * m.nxv + ix] : 0.0;
  }
  if (P.has_o) ov = P.o[i];
  if (MODE == 2 && P.ob0 != 0.0) ov = (P.ob0 * P.invd.v[type]) * rv;
  const double rc = v[0], self = v[0];
  if (MODE == 3) {
    constexpr int DX[PFH_NOFF3] = {0, 1, -1, 0, 0, 0, 0, 1, -1, 0, 0, 1, -1, 1, -1};
    constexpr int DY[PFH_NOFF3] = {0, 0, 0, 1, -1, 0, 0, 1, -1, 1, -1, 0, 0, 1, -1};
    constexpr int DZ[PFH_NOFF3] = {0, 0, 0, 0, 0, 1, -1, 0, 0, 1, -1, 1, -1, 1, -1};
    auto cls = [](int q, int n) { return q == 0 ? 0 : (q == n - 1 ? 2 : 1); };
    // operand o is r at the neighbour times b0 / diag of the NEIGHBOUR's row; out-of-domain operands are zero anyway
#pragma unroll
    for (int o = 0; o < PFH_NOFF3; ++o)
      v[o] *= P.b0 * P.invd.v[9 * cls(gz + DZ[o], m.nzv) + 3 * cls(iy + DY[o], m.nyv) + cls(ix + DX[o], m.nxv)];
  }
  vtx::Row3 row;
  vtx::get_row3(m, type, row);
  double Ms, Ks;
  vtx::mk_sums3(row, v, Ms, Ks);
  const double Av = P.am * Ms + P.ak * Ks;
  const double bd = MODE >= 2 ? P.b / (P.am * row.m[0] + P.ak * row.k[0]) : 0.0;
  if (MODE == 0) {
    P.out[i] = P.a * rv + Av;
    if (P.out1 != nullptr) P.out1[i] = self;
  } else if (MODE == 2) {
    P.out[i] = self + P.a * (self - ov) + bd * (rv - Av);
  } else {
    const double z1 = v[0];  // already scaled by b0 / diag of this row
    if (P.out1 != nullptr) P.out1[i] = z1;
    P.out[i] = z1 + P.a * z1 + bd * (rc - Av);
  }
}

constexpr int S_TX = 64, S_TY = 4, S_NS = 8, S_MINB = 3;   // 59 KB per CTA, three CTAs per SM

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
struct TmapKey {
  const void* p;
  int nxv, nyv, nz, bx, by;
  bool operator==(const TmapKey& o) const { return p == o.p && nz == o.nz && bx == o.bx && by == o.by && nxv == o.nxv && nyv == o.nyv; }
};
struct TmapHash {
  size_t operator()(const TmapKey& k) const {
    size_t h = reinterpret_cast<size_t>(k.p) >> 4;
    h = h * 1000003u ^ (size_t)k.nz;
    h = h * 1000003u ^ (size_t)(k.bx * 64 + k.by);
    return h;
  }
};
struct TmapCache {
  std::unordered_map<TmapKey, CUtensorMap, TmapHash> e;
};
}  // namespace

CUtensorMap pfh_tensor_map(pfh_ctx* c, const double* ptr, int nxv, int nyv, int nz, int bx, int by) {
  if (c->tmap_cache == nullptr) c->tmap_cache = new TmapCache;
  TmapCache* tc = static_cast<TmapCache*>(c->tmap_cache);
  const TmapKey key{ptr, nxv, nyv, nz, bx, by};
  auto it = tc->e.find(key);
  if (it != tc->e.end()) return it->second;
  if (tc->e.size() >= 4096) tc->e.clear();  // caller-owned buffers may come and go
  const CUtensorMap map = tma::make_map(ptr, nxv, nyv, nz, bx, by);
  tc->e.emplace(key, map);
  return map;
}

void pfh_tmap_cache_free(pfh_ctx* c) {
  delete static_cast<TmapCache*>(c->tmap_cache);
  c->tmap_cache = nullptr;
}

// dynamic shared memory opt-in of every instantiation, for the device the ctx lives on (the attribute is per device)
void pfh_scalar3_init_device() {
  scalar_set_attr_all<S_TX, S_TY, S_NS, S_MINB>();
  scalar_set_attr_all<64, 4, 6, 3>();
  scalar_set_attr_all<64, 4, 4, 3>();
  scalar_set_attr_all<64, 4, 10, 3>();
  scalar_set_attr_all<64, 8, 6, 2>();
  scalar_set_attr_all<128, 2, 8, 3>();
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
  double cI[PFH_NOFF3];
  for (int o = 0; o < PFH_NOFF3; ++o) cI[o] = op.am * m.mI[o] + (o < PFH_NKOFF3 ? op.ak * m.kI[o] : 0.0);
  for (int o = 1; o < PFH_NOFF3; o += 2) PFH_CHECK(cI[o] == cI[o + 1], "internal: interior stencil row is not symmetric");
  P.cS.c[0] = cI[0];
  for (int j = 1; j < 8; ++j) P.cS.c[j] = cI[2 * j - 1];
  P.bdI = op.b / cI[0];
  P.ob0 = (op.mode == 2 && op.o == nullptr) ? op.ob0 : 0.0;
  if (c->h_mT.empty()) {  // host copy of the type tables, for the inverse diagonals
    Mesh3 mm = m;
    pfh_build_tables3(mm, c->h_mT, c->h_kT);
  }
  for (int t = 0; t < 27; ++t) P.invd.v[t] = 1.0 / (op.am * c->h_mT[t * PFH_NOFF3] + op.ak * c->h_kT[t * PFH_NKOFF3]);
  for (int t = 0; t < 9; ++t)
    for (int o = 0; o < PFH_NOFF3; ++o)
      P.rowsB[t][o] = op.am * c->h_mT[(9 + t) * PFH_NOFF3 + o] + (o < PFH_NKOFF3 ? op.ak * c->h_kT[(9 + t) * PFH_NKOFF3 + o] : 0.0);
  P.s1I = op.b0 * P.invd.v[13];
  P.ob0I = P.ob0 * P.invd.v[13];
  P.shell = op.mode == 3 ? 2 : 1;
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
  // the boundary shell of the window
  const int d = P.shell;
  int zs[4] = {0, 0, 0, 0}, nzs = 0, zi0 = 0, nzi = 0;
  for (int zl = 0; zl < op.nzo; ++zl) {
    const int gz = op.gz0 + zl;
    if (gz < d || gz >= m.nzv - d) {
      PFH_CHECK(nzs < 4, "internal: more than four shell planes in a window");
      zs[nzs++] = zl;
    } else {
      if (nzi == 0) zi0 = zl;
      ++nzi;
    }
  }
  const long long csx = std::min(2 * d, m.nxv), csy = std::min(2 * d, m.nyv);
  // MODE 0 / 2: the staged kernel does the x and y faces itself; only the planes on a z boundary are left
  const long long n_threads = op.mode == 3 ? (long long)nzs * m.nxy + (long long)nzi * (csy * m.nxv + (m.nyv - csy) * csx) : (long long)nzs * m.nxy;
  // The shell kernel reads what the staged kernel reads and stores the vertices it skips: the two are independent, so
  // the (latency-bound, partly uncoalesced) gather runs on aux_stream underneath the staged kernel.
  const bool fork = c->shell_async && c->aux_stream != nullptr && n_threads > 0;
  cudaStream_t sst = fork ? c->aux_stream : c->stream;
  if (fork) {
    PFH_CUDA(cudaEventRecord(c->ev_fork, c->stream));
    PFH_CUDA(cudaStreamWaitEvent(c->aux_stream, c->ev_fork, 0));
  }
  if (n_threads > 0) {
    const unsigned nb = (unsigned)((n_threads + 127) / 128);
    if (op.mode == 0) k3_scalar_shell<0><<<nb, 128, 0, sst>>>(P, nzs, zs[0], zs[1], zs[2], zs[3], zi0, nzi, n_threads);
    else if (op.mode == 2) k3_scalar_shell<2><<<nb, 128, 0, sst>>>(P, nzs, zs[0], zs[1], zs[2], zs[3], zi0, nzi, n_threads);
    else k3_scalar_shell<3><<<nb, 128, 0, sst>>>(P, nzs, zs[0], zs[1], zs[2], zs[3], zi0, nzi, n_threads);
    PFH_CUDA(cudaGetLastError());
    c->launches++;
  }
  if (fork) PFH_CUDA(cudaEventRecord(c->ev_join, c->aux_stream));
  switch (c->scalar_variant) {
    case 1: scalar_launch<64, 4, 6, 3>(c, op, P, use_tma); break;
    case 2: scalar_launch<64, 4, 4, 3>(c, op, P, use_tma); break;
    case 3: scalar_launch<64, 4, 10, 3>(c, op, P, use_tma); break;
    case 4: scalar_launch<64, 8, 6, 2>(c, op, P, use_tma); break;    // 544 threads, 2 CTAs per SM
    case 5: scalar_launch<128, 2, 8, 3>(c, op, P, use_tma); break;
    default: scalar_launch<S_TX, S_TY, S_NS, S_MINB>(c, op, P, use_tma); break;
  }
  if (fork) PFH_CUDA(cudaStreamWaitEvent(c->stream, c->ev_join, 0));

}
