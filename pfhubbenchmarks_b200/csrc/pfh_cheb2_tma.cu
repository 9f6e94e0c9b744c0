// Temporally blocked Jacobi-Chebyshev kernels for the 3-D scalar SPD operators A = am M + ak K.
//
// A Chebyshev sweep z_{k+1} = z_k + a_k (z_k - z_{k-1}) + b_k D^-1 (r - A z_k) streams 32 B/vertex for 15 fused
// multiply-adds.  Two consecutive sweeps only need data within radius 2 of a vertex, so they are fused into one
// pass: stage X evaluates the first sweep on the tile plus a halo of 1 (from the operand on a halo of 2), keeps it in
// shared memory, and stage Y evaluates the second sweep on the tile.  Two kernels cover a whole solve:
//   FIRST = true  : r            ->  z1 = b0 D^-1 r (pointwise, in shared memory), z2 (X), z3 (Y)     16-24 B/vertex
//   FIRST = false : z_k, z_{k-1}, r ->  z_{k+1} (X, optional store), z_{k+2} (Y)                      32-40 B/vertex
// so the 3-sweep mass solve is ONE pass over r (16 B/vertex instead of 72) and the 5-sweep shifted solve is two passes
// (56 B/vertex instead of 152).  Planes travel by TMA into a ring of shared-memory slots (one mbarrier per slot);
// out-of-bounds zero fill supplies everything outside the domain; slab ghosts are two-plane tensors.
// Replaces PETSc PCApply (SOR) on the reference's hot path; algorithm statement: oracle/krylov.py:cheb_solve.
//
// STATUS (round 1): bit-compatible with the per-sweep path (tests/test_gpu_parity.py::test_fused_chebyshev_*), but
// MEASURED SLOWER at 512^3: 2.58 ms per fused pass vs 0.77 ms per sweep -> 12.9 ms vs 8.9 ms per preconditioner
// application.  The 15-point fp64 stencil re-reads ~18 shared-memory words per point in each of the two stages
// (~180 KB of LDS traffic per CTA-plane = 1.3 ms at 128 B/clk/SM), so the fused kernel trades an HBM bound for a
// shared-memory bound.  Kept as an opt-in (pfh_set_option "use_fused") for the register-blocked follow-up.
#include "pfh_tma.cuh"
#include "pfh_vertex.cuh"

namespace {
using namespace tma;

struct Coef15 {
  double c[PFH_NOFF3];
};
struct Maps3 {
  CUtensorMap main, lo, hi;
};

constexpr int C_TX = 64, C_TY = 8, C_NT = 256, C_NG = 4, C_ZCH = 32;

// A v at tile-local index c of three consecutive planes (pitch W), interior row
template <int W>
__device__ __forceinline__ double apply_interior(const Coef15& k, const double* S0, const double* S1, const double* S2, int c) {
  double s0 = k.c[0] * S1[c], s1 = k.c[1] * S1[c + 1], s2 = k.c[2] * S1[c - 1], s3 = k.c[3] * S1[c + W];
  s0 += k.c[4] * S1[c - W]; s1 += k.c[5] * S2[c]; s2 += k.c[6] * S0[c]; s3 += k.c[7] * S1[c + W + 1];
  s0 += k.c[8] * S1[c - W - 1]; s1 += k.c[9] * S2[c + W]; s2 += k.c[10] * S0[c - W]; s3 += k.c[11] * S2[c + 1];
  s0 += k.c[12] * S0[c - 1]; s1 += k.c[13] * S2[c + W + 1]; s2 += k.c[14] * S0[c - W - 1];
  return (s0 + s1) + (s2 + s3);
}
template <int W>
__device__ __forceinline__ double apply_table(const Mesh3& m, int type, double am, double ak, const double* S0, const double* S1,
                                              const double* S2, int c, double& d) {
  vtx::Row3 row;
  vtx::get_row3(m, type, row);
  double v[PFH_NOFF3];
  v[0] = S1[c]; v[1] = S1[c + 1]; v[2] = S1[c - 1]; v[3] = S1[c + W]; v[4] = S1[c - W];
  v[5] = S2[c]; v[6] = S0[c]; v[7] = S1[c + W + 1]; v[8] = S1[c - W - 1];
  v[9] = S2[c + W]; v[10] = S0[c - W]; v[11] = S2[c + 1]; v[12] = S0[c - 1];
  v[13] = S2[c + W + 1]; v[14] = S0[c - W - 1];
  double Ms, Ks;
  vtx::mk_sums3(row, v, Ms, Ks);
  d = am * row.m[0] + ak * row.k[0];
  return am * Ms + ak * Ks;
}
// boundary type of global vertex (gx, gy, gz), -1 outside the domain
__device__ __forceinline__ int vtype(const Mesh3& m, int gx, int gy, int gz) {
  if (gx < 0 || gx >= m.nxv || gy < 0 || gy >= m.nyv || gz < 0 || gz >= m.nzv) return -1;
  return vtx::type3(m, gx, gy, gz);
}
__device__ __forceinline__ double diag_of(const Mesh3& m, int type, double am, double ak, double dI) {
  return type == 13 ? dI : am * __ldg(m.mT + type * PFH_NOFF3) + ak * __ldg(m.kT + type * PFH_NKOFF3);
}

template <bool FIRST, int TX, int TY, int NT, int NG>
__global__ void __launch_bounds__(NT, 2)
    k3_cheb2_tma(const __grid_constant__ Maps3 mS, const __grid_constant__ Maps3 mR, const __grid_constant__ Maps3 mP, Mesh3 m,
                 int has_lo, int has_hi, int zchunk, double am, double ak, Coef15 cI, double b0, double aX, double bX, double aY,
                 double bY, double* __restrict__ outX, double* __restrict__ outY) {
  constexpr int W = TX + 4, H = TY + 4, TILE = W * H;  // tile + halo 2; 68 x 12 doubles = 6528 B (multiple of 128)
  constexpr int TILE_BYTES = TILE * 8;
  constexpr int NT_TILES = FIRST ? 2 : 3;  // FIRST: [r | z1]   else: [z_k | r | z_{k-1}]
  constexpr int SLOT_BYTES = NT_TILES * TILE_BYTES;
  constexpr int XW = TX + 2, XH = TY + 2;  // stage-X region: tile + halo 1
  extern __shared__ __align__(128) unsigned char smem[];
  double* xring = reinterpret_cast<double*>(smem + NG * SLOT_BYTES);  // 3 planes, same geometry as a tile
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + NG * SLOT_BYTES + 3 * TILE_BYTES);
  auto tile = [&](int g, int which) { return reinterpret_cast<double*>(smem + (g % NG) * SLOT_BYTES + which * TILE_BYTES); };
  // which-index of the operand (stencil) field S, of r, and of z_{k-1}
  constexpr int IS = FIRST ? 1 : 0, IR = FIRST ? 0 : 1, IP = 2;

  const int tid = threadIdx.x;
  const int x0 = blockIdx.x * TX, y0 = blockIdx.y * TY;
  const int zb = blockIdx.z * zchunk, ze = min(m.nzl, zb + zchunk);
  const int n = ze - zb;
  const int ngroups = n + 4;  // group g <-> local plane zl = zb - 2 + g
  const double dI = cI.c[0];
  const double b0I = b0 / dI, bXI = bX / dI, bYI = bY / dI;

  if (tid == 0) {
    for (int s = 0; s < NG; ++s) mbar_init(&bars[s], 1);
    fence_barrier_init();
  }
  __syncthreads();

  // coordinates of plane zl in the (main | 2-plane lo ghost | 2-plane hi ghost) tensors; physical boundaries: main map, OOB
  auto pick = [&](const Maps3& M, int zl, int& zc) -> const CUtensorMap* {
    if (zl < 0 && has_lo) {
      zc = zl + 2;
      return &M.lo;
    }
    if (zl >= m.nzl && has_hi) {
      zc = zl - m.nzl;
      return &M.hi;
    }
    zc = zl;
    return &M.main;
  };
  auto issue = [&](int g) {
    const int zl = zb - 2 + g;
    uint64_t* bar = &bars[g % NG];
    const bool mid = g >= 1 && g <= n + 2;  // planes zb-1 .. ze also carry the pointwise operands
    int zc;
    if (FIRST) {
      mbar_expect_tx(bar, TILE_BYTES);
      const CUtensorMap* mp = pick(mR, zl, zc);
      tma_load_3d(tile(g, IR), mp, x0 - 2, y0 - 2, zc, bar);
    } else {
      mbar_expect_tx(bar, TILE_BYTES + (mid ? 2 * TILE_BYTES : 0));
      const CUtensorMap* mp = pick(mS, zl, zc);
      tma_load_3d(tile(g, IS), mp, x0 - 2, y0 - 2, zc, bar);
      if (mid) {
        mp = pick(mR, zl, zc);
        tma_load_3d(tile(g, IR), mp, x0 - 2, y0 - 2, zc, bar);
        mp = pick(mP, zl, zc);
        tma_load_3d(tile(g, IP), mp, x0 - 2, y0 - 2, zc, bar);
      }
    }
  };
  if (tid == 0)
    for (int g = 0; g < NG - 1 && g < ngroups; ++g) issue(g);

  // Per-thread point lists, fixed for the whole z march: shared-memory index, x-y boundary type (-1 outside the
  // domain) and global in-plane offset.  Only the z type changes from plane to plane (uniform over the CTA).
  constexpr int NCV = (TILE + NT - 1) / NT, NXP = (XW * XH + NT - 1) / NT, NYP = (TX * TY + NT - 1) / NT;
  auto xy_type = [&](int gx, int gy) -> int {
    if (gx < 0 || gx >= m.nxv || gy < 0 || gy >= m.nyv) return -1;
    return (gy == 0 ? 0 : (gy == m.nyv - 1 ? 2 : 1)) * 3 + (gx == 0 ? 0 : (gx == m.nxv - 1 ? 2 : 1));
  };
  int cv_t[NCV];                       // convert: point q = tid + k NT of the whole tile (index = q)
  int xp_c[NXP], xp_t[NXP], xp_g[NXP]; // stage X
  int yp_c[NYP], yp_t[NYP], yp_g[NYP]; // stage Y
#pragma unroll
  for (int k = 0; k < NCV; ++k) {
    const int q = tid + k * NT;
    cv_t[k] = q < TILE ? xy_type(x0 - 2 + q % W, y0 - 2 + q / W) : -2;
  }
#pragma unroll
  for (int k = 0; k < NXP; ++k) {
    const int q = tid + k * NT;
    const int lx = 1 + q % XW, ly = 1 + q / XW;
    xp_c[k] = ly * W + lx;
    xp_t[k] = q < XW * XH ? xy_type(x0 - 2 + lx, y0 - 2 + ly) : -2;
    const bool own = lx >= 2 && lx < W - 2 && ly >= 2 && ly < H - 2;
    xp_g[k] = own ? (y0 - 2 + ly) * m.nxv + (x0 - 2 + lx) : -1;
  }
#pragma unroll
  for (int k = 0; k < NYP; ++k) {
    const int q = tid + k * NT;
    const int lx = 2 + q % TX, ly = 2 + q / TX;
    yp_c[k] = ly * W + lx;
    yp_t[k] = q < TX * TY ? xy_type(x0 - 2 + lx, y0 - 2 + ly) : -2;
    yp_g[k] = (y0 - 2 + ly) * m.nxv + (x0 - 2 + lx);
  }
  auto z_type = [&](int gz) -> int { return gz < 0 || gz >= m.nzv ? -1 : (gz == 0 ? 0 : (gz == m.nzv - 1 ? 2 : 1)); };

  // FIRST: z1 = b0 D^-1 r on the whole tile of group g (the operand of stage X)
  auto convert = [&](int g) {
    const int tz = z_type(m.z0 + zb - 2 + g);
    const double* R = tile(g, IR);
    double* S = tile(g, IS);
#pragma unroll
    for (int k = 0; k < NCV; ++k) {
      if (cv_t[k] == -2) continue;
      const int q = tid + k * NT;
      const int t = (tz < 0 || cv_t[k] < 0) ? -1 : tz * 9 + cv_t[k];
      S[q] = t < 0 ? 0.0 : (t == 13 ? b0I : b0 / diag_of(m, t, am, ak, dI)) * R[q];
    }
  };

  for (int j = 0; j <= n + 1; ++j) {
    __syncthreads();  // slot of group j-1 and x-ring plane j-3 are free
    if (tid == 0 && j + NG - 1 < ngroups) issue(j + NG - 1);
    if (j == 0) {
      mbar_wait(&bars[0], 0);
      mbar_wait(&bars[1 % NG], (1 / NG) & 1);
    }
    mbar_wait(&bars[(j + 2) % NG], ((j + 2) / NG) & 1);
    if (FIRST) {
      if (j == 0) {
        convert(0);
        convert(1);
      }
      convert(j + 2);
      __syncthreads();
    }

    // ---- stage X: plane zx = zb - 1 + j (group j+1) on tile + halo 1 ----
    {
      const int zx = zb - 1 + j;
      const int tz = z_type(m.z0 + zx);
      const double* S0 = tile(j, IS);
      const double* S1 = tile(j + 1, IS);
      const double* S2 = tile(j + 2, IS);
      const double* R = tile(j + 1, IR);
      const double* P = tile(j + 1, IP);
      double* X = xring + (j % 3) * TILE;
      const bool zowned = outX != nullptr && zx >= zb && zx < ze;
      const long long zoff = (long long)zx * m.nxy;
#pragma unroll
      for (int k = 0; k < NXP; ++k) {
        if (xp_t[k] == -2) continue;
        const int c = xp_c[k];
        const int t = (tz < 0 || xp_t[k] < 0) ? -1 : tz * 9 + xp_t[k];
        double xv = 0.0;
        if (t >= 0) {
          double Av, bd;
          if (t == 13) {
            Av = apply_interior<W>(cI, S0, S1, S2, c);
            bd = bXI;
          } else {
            double d;
            Av = apply_table<W>(m, t, am, ak, S0, S1, S2, c, d);
            bd = bX / d;
          }
          const double s = S1[c];
          const double prev = FIRST ? 0.0 : P[c];
          xv = s + aX * (s - prev) + bd * (R[c] - Av);
          if (zowned && xp_g[k] >= 0) outX[zoff + xp_g[k]] = xv;
        }
        X[c] = xv;
      }
    }
    __syncthreads();

    // ---- stage Y: plane zy = zb - 2 + j (group j) on the tile ----
    if (j >= 2) {
      const int zy = zb - 2 + j;
      const int tz = z_type(m.z0 + zy);
      const double* X0 = xring + ((j - 2) % 3) * TILE;
      const double* X1 = xring + ((j - 1) % 3) * TILE;
      const double* X2 = xring + (j % 3) * TILE;
      const double* S = tile(j, IS);
      const double* R = tile(j, IR);
      const long long zoff = (long long)zy * m.nxy;
#pragma unroll
      for (int k = 0; k < NYP; ++k) {
        if (yp_t[k] < 0 || tz < 0) continue;
        const int c = yp_c[k];
        const int t = tz * 9 + yp_t[k];
        double Av, bd;
        if (t == 13) {
          Av = apply_interior<W>(cI, X0, X1, X2, c);
          bd = bYI;
        } else {
          double d;
          Av = apply_table<W>(m, t, am, ak, X0, X1, X2, c, d);
          bd = bY / d;
        }
        const double u = X1[c];
        outY[zoff + yp_g[k]] = u + aY * (u - S[c]) + bd * (R[c] - Av);
      }
    }
  }
}

Maps3 maps3(const pfh_ctx* c, const double* p, const double* glo2, const double* ghi2, bool lo, bool hi) {
  const Mesh3& m = c->m3;
  Maps3 M;
  M.main = make_map(p, m.nxv, m.nyv, m.nzl, C_TX + 4, C_TY + 4);
  M.lo = lo ? make_map(glo2, m.nxv, m.nyv, 2, C_TX + 4, C_TY + 4) : M.main;
  M.hi = hi ? make_map(ghi2, m.nxv, m.nyv, 2, C_TX + 4, C_TY + 4) : M.main;
  return M;
}

}  // namespace

// One fused pass of two Chebyshev sweeps.  first: (r) -> z2 (outX, optional), z3 (outY) with z1 = b0 D^-1 r implicit;
// otherwise (zk, zkm1, r) -> z_{k+1} (outX, optional), z_{k+2} (outY).  Ghost operands: two planes per side in
// g*_lo / g*_hi (zk: c->g2S, r: c->g2R, zkm1: c->g2P), already exchanged by the caller.
bool k3_cheb2_launch(pfh_ctx* c, bool first, double am, double ak, double b0, double aX, double bX, double aY, double bY,
                     const double* zk, const double* zkm1, const double* r, double* outX, double* outY) {
  if (!pfh_tma_ok(c) || !aligned16(r) || !aligned16(outY) || (outX && !aligned16(outX))) return false;
  if (!first && (!aligned16(zk) || !aligned16(zkm1))) return false;
  const Mesh3& m = c->m3;
  const bool lo = c->nplane > 0 && c->cfg.rank > 0, hi = c->nplane > 0 && c->cfg.rank < c->cfg.nranks - 1;
  const Maps3 mR = maps3(c, r, c->g2R_lo, c->g2R_hi, lo, hi);
  const Maps3 mS = first ? mR : maps3(c, zk, c->g2S_lo, c->g2S_hi, lo, hi);
  const Maps3 mP = first ? mR : maps3(c, zkm1, c->g2P_lo, c->g2P_hi, lo, hi);
  Coef15 cI;
  for (int o = 0; o < PFH_NOFF3; ++o) cI.c[o] = am * m.mI[o] + (o < PFH_NKOFF3 ? ak * m.kI[o] : 0.0);
  const dim3 grid((m.nxv + C_TX - 1) / C_TX, (m.nyv + C_TY - 1) / C_TY, (m.nzl + C_ZCH - 1) / C_ZCH);
  constexpr int tile_bytes = (C_TX + 4) * (C_TY + 4) * 8;
  static bool attr_set = false;
  if (!attr_set) {
    PFH_CUDA(cudaFuncSetAttribute(k3_cheb2_tma<true, C_TX, C_TY, C_NT, C_NG>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  C_NG * 2 * tile_bytes + 3 * tile_bytes + C_NG * 8));
    PFH_CUDA(cudaFuncSetAttribute(k3_cheb2_tma<false, C_TX, C_TY, C_NT, C_NG>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  C_NG * 3 * tile_bytes + 3 * tile_bytes + C_NG * 8));
    attr_set = true;
  }
  if (first) {
    const size_t smem = C_NG * 2 * tile_bytes + 3 * tile_bytes + C_NG * 8;
    k3_cheb2_tma<true, C_TX, C_TY, C_NT, C_NG><<<grid, C_NT, smem, c->stream>>>(mS, mR, mP, m, lo, hi, C_ZCH, am, ak, cI, b0, aX, bX, aY, bY,
                                                                              outX, outY);
  } else {
    const size_t smem = C_NG * 3 * tile_bytes + 3 * tile_bytes + C_NG * 8;
    k3_cheb2_tma<false, C_TX, C_TY, C_NT, C_NG><<<grid, C_NT, smem, c->stream>>>(mS, mR, mP, m, lo, hi, C_ZCH, am, ak, cI, b0, aX, bX, aY,
                                                                               bY, outX, outY);
  }
  PFH_CUDA(cudaGetLastError());
  return true;
}
