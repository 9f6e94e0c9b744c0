// Cube-centric Jacobian-vector product and residual on the 3-D Kuhn mesh for sm_100a -- the B200 stand-in for FFC's
// tabulate_tensor + DOLFIN assembly of F and J = df.derivative(F, w, dw) followed by PETSc MatMult
// (dolfin/b13d.py:64-68; SURVEY.md section 3.3 loops A-C), matrix-free.
//
// One CTA owns a TX x TY tile of the x-y plane and marches along z.  The planes of every operand field are staged in
// a ring of shared-memory stages (tile + halo), filled either by TMA (one elected thread, cp.async.bulk.tensor.3d,
// out-of-bounds zero fill = the halo outside the domain; needs an even vertex count in x) or by 8-byte cp.async copies
// of all threads (any vertex count, e.g. the reference's 51^3 mesh), both signalling mbarriers.
//
// The nonlinear block int phi_i g'(u_h) x_h is evaluated ONCE per tetrahedron (5-point Zienkiewicz-Taylor rule,
// pfh_elem.cuh): in the "cube phase" of a layer every thread evaluates one cube (6 Kuhn tets) and writes its 8 corner
// contributions to shared memory; in the "vertex phase" every vertex gathers the contributions of its 8 cubes (4 of
// this layer, 4 carried in a register from the layer below) and adds the linear 15-point part.  The contribution
// buffer is double buffered, so a layer needs ONE block-wide barrier (between its cube and vertex phase): warps that
// finish a vertex phase early start the next layer's cubes while the others still gather.
//
//  KIND 0: y = J(u) x (all NF rows).                                              8 (NIN + 2 NF) B/vertex = 40 (CH)
//  KIND 1: only the nonlinear row from its own nonlinear column: out = (Mc[o][i] M + Kc[o][i] K) x_i + N'[x_i]
//          (the (kappa K + M_f'') block of the Cahn-Hilliard Jacobian for the full Schur preconditioner)    24 B/vertex
//  KIND 2: residual R(u; u0): operand fields u, extra ring fields u0 of the fields in(.)                       40 (CH)
//  F32Q  : KIND 0 of a QUARTIC problem with the quadrature block in single precision (opt-in experiment).
//
// Measured dead end (round 2, profiles/README.md): cube and vertex phases as different WARPS of one CTA (10 cube warps + 8
// vertex warps + a producer warp, one CTA per SM because the cube threads want ~100 registers) is correct but slower,
// 4.6 ms against 4.2 ms at 512^3: 19 warps per SM do not keep the fp64 pipe busier than three CTAs of this kernel.
#include <algorithm>

#include "pfh_tma.cuh"
#include "pfh_vertex.cuh"

namespace {
using namespace tma;

template <int N>
struct MapSet {
  CUtensorMap main[N], lo[N], hi[N];
};
template <int N>
struct JacParams {
  Mesh3 m;
  LinCoefs L;
  int has_lo, has_hi, zchunk;
  elem::QuarticJac qj;
  elem::QuarticRes qr;
  double* out;
  const double *base[N], *glo[N], *ghi[N];  // raw operand pointers (cp.async loader)
};

template <class P, int KIND, int TX, int TY, int NT, int NS, int OCC, bool TMA, bool F32Q = false>
__global__ void __launch_bounds__(NT, OCC)
    k3_jac_tma(const __grid_constant__ MapSet<(KIND == 1 ? 1 : P::NF) + P::NIN> maps,
               const __grid_constant__ JacParams<(KIND == 1 ? 1 : P::NF) + P::NIN> A, P prob) {
  constexpr bool CROW = KIND == 1, RES = KIND == 2;
  // KIND 0 / 2: every vertex is treated as an interior row here (symmetric-pair sums, constants from the parameter bank,
  // no row registers); the stores of the boundary vertices are skipped and k3_jac_shell computes them.  KIND 1 keeps the
  // general rows in the kernel.
  constexpr bool SHELL = !CROW;
  constexpr int NX = CROW ? 1 : P::NF;  // ring-buffered operand fields x, followed by NIN fields (u for J.v, u0 for RES)
  constexpr int NFLD = NX + P::NIN;
  constexpr int W = TX + 4, H = TY + 2;  // x halo of 2 (16-byte aligned TMA box start)
  constexpr int STAGE_BYTES = W * H * 8;
  constexpr int STAGE_STRIDE = round_up(STAGE_BYTES, 128);
  constexpr int CW = TX + 1, CH = TY + 1, NCUBE = CW * CH;
  constexpr int CPAD = NCUBE | 1;  // odd pitch: corner arrays start in different banks
  constexpr int CSIZE = P::NOUT * 8 * CPAD;
  extern __shared__ __align__(128) unsigned char smem[];
  // layout: [NFLD][NS] stages | contrib[2][NOUT][8][CPAD] | mbarriers[NS]
  double* contrib_base = reinterpret_cast<double*>(smem + NFLD * NS * STAGE_STRIDE);
  uint64_t* bars = reinterpret_cast<uint64_t*>(contrib_base + 2 * CSIZE);
  auto stage = [&](int f, int p) { return reinterpret_cast<const double*>(smem + (f * NS + (p % NS)) * STAGE_STRIDE); };
  auto xstage = [&](int field, int p) { return stage(CROW ? 0 : field, p); };  // operand field `field`

  const Mesh3& m = A.m;
  const LinCoefs& L = A.L;
  const int tid = threadIdx.x;
  const int x0 = blockIdx.x * TX, y0 = blockIdx.y * TY;
  const int zb = blockIdx.z * A.zchunk, ze = min(m.nzl, zb + A.zchunk);
  const int nplanes = ze - zb + 2;  // planes zb-1 .. ze

  if (tid == 0) {
    for (int s = 0; s < NS; ++s) mbar_init(&bars[s], TMA ? 1 : NT);
    fence_barrier_init();
  }
  __syncthreads();

  auto issue = [&](int p) {
    const int zl = zb - 1 + p;
    uint64_t* bar = &bars[p % NS];
    if constexpr (TMA) {
      if (tid == 0) {
        mbar_expect_tx(bar, NFLD * STAGE_BYTES);
#pragma unroll
        for (int f = 0; f < NFLD; ++f) {
          const CUtensorMap* map = &maps.main[f];
          int zc = zl;  // -1 or nzl without a ghost plane: beyond the physical boundary -> zero fill
          if (zl < 0 && A.has_lo) {
            map = &maps.lo[f];
            zc = 0;
          } else if (zl >= m.nzl && A.has_hi) {
            map = &maps.hi[f];
            zc = 0;
          }
          tma_load_3d(smem + (f * NS + (p % NS)) * STAGE_STRIDE, map, x0 - 2, y0 - 1, zc, bar);
        }
      }
    } else {
#pragma unroll
      for (int f = 0; f < NFLD; ++f) {
        const double* src = nullptr;
        if (zl >= 0 && zl < m.nzl) src = A.base[f] + (long long)zl * m.nxy;
        else if (zl < 0 && A.has_lo) src = A.glo[f];
        else if (zl >= m.nzl && A.has_hi) src = A.ghi[f];
        unsigned char* dst = smem + (f * NS + (p % NS)) * STAGE_STRIDE;
        for (int i = tid; i < W * H; i += NT) {
          const int gx = x0 - 2 + i % W, gy = y0 - 1 + i / W;
          const bool ok = src != nullptr && gx >= 0 && gx < m.nxv && gy >= 0 && gy < m.nyv;
          cp_async8(dst + i * 8, ok ? src + (long long)gy * m.nxv + gx : A.base[0], ok ? 8 : 0);
        }
      }
      cp_async_mbar_arrive(bar);
    }
  };
  for (int p = 0; p < NS && p < nplanes; ++p) issue(p);

  // vertex role (tid < TX*TY)
  const int vtx_x = tid % TX, vtx_y = tid / TX;
  const int ix = x0 + vtx_x, iy = y0 + vtx_y;
  const bool is_vertex = tid < TX * TY && ix < m.nxv && iy < m.nyv;
  const int txy = (iy == 0 ? 0 : (iy == m.nyv - 1 ? 2 : 1)) * 3 + (ix == 0 ? 0 : (ix == m.nxv - 1 ? 2 : 1));
  double acc_hi[P::NOUT];
#pragma unroll
  for (int o = 0; o < P::NOUT; ++o) acc_hi[o] = 0.0;

  // iteration `it` handles cube layer l = zb-1+it (between planes l and l+1) and writes plane l (it >= 1)
  for (int it = 0; it <= ze - zb; ++it) {
    const int l = zb - 1 + it;
    double* contrib = contrib_base + (it & 1) * CSIZE;
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
          if (RES) elem::cube_res_quartic(A.qr, u[0], o8[0]);
          else if (F32Q) elem::cube_jac_quartic_f32(A.qj, u[0], x[0], o8[0]);
          else elem::cube_jac_quartic(A.qj, u[0], x[0], o8[0]);
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
    // the only block-wide barrier of the layer: every cube of layer `it` is written, and every thread has finished
    // the vertex phase of layer it-1, so plane it-2 of the ring and the other contribution buffer are free
    __syncthreads();
    if (it >= 2 && it - 2 + NS < nplanes) issue(it - 2 + NS);

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
        const int c = (vtx_y + 1) * W + vtx_x + 2;
        // (M v)_i and (K v)_i of the field staged as ring field f
        auto lin_sums = [&](int f, const vtx::Row3& row, double& Ms, double& Ks) {
          const double* S0 = stage(f, it - 1);
          const double* S1 = stage(f, it);
          const double* S2 = stage(f, it + 1);
          double v[PFH_NOFF3];
          v[0] = S1[c]; v[1] = S1[c + 1]; v[2] = S1[c - 1]; v[3] = S1[c + W]; v[4] = S1[c - W];
          v[5] = S2[c]; v[6] = S0[c]; v[7] = S1[c + W + 1]; v[8] = S1[c - W - 1];
          v[9] = S2[c + W]; v[10] = S0[c - W]; v[11] = S2[c + 1]; v[12] = S0[c - 1];
          v[13] = S2[c + W + 1]; v[14] = S0[c - W - 1];
          if constexpr (SHELL) {
            // interior row: symmetric, one coefficient per +- pair; K has no diagonal-neighbour entries
            const double px = v[1] + v[2], py = v[3] + v[4], pz = v[5] + v[6];
            Ms = (m.mI[0] * v[0] + m.mI[1] * px + m.mI[3] * py + m.mI[5] * pz) +
                 (m.mI[7] * (v[7] + v[8]) + m.mI[9] * (v[9] + v[10]) + m.mI[11] * (v[11] + v[12]) + m.mI[13] * (v[13] + v[14]));
            Ks = m.kI[0] * v[0] + m.kI[1] * px + m.kI[3] * py + m.kI[5] * pz;
          } else {
            vtx::mk_sums3(row, v, Ms, Ks);
          }
        };
        vtx::Row3 row;
        bool store = true;
        if constexpr (SHELL) {
          store = ix >= 1 && ix <= m.nxv - 2 && iy >= 1 && iy <= m.nyv - 2 && gz >= 1 && gz <= m.nzv - 2;
        } else {
          vtx::get_row3(m, (gz == 0 ? 0 : (gz == m.nzv - 1 ? 2 : 1)) * 9 + txy, row);
        }
        double Ms[NX], Ks[NX];
#pragma unroll
        for (int f = 0; f < NX; ++f) lin_sums(f, row, Ms[f], Ks[f]);
        const long long i = (long long)l * m.nxy + (long long)iy * m.nxv + ix;
        if (CROW) {
          A.out[i] = L.Mc[P::out(0)][P::in(0)] * Ms[0] + L.Kc[P::out(0)][P::in(0)] * Ks[0] + acc_hi[0] + lo[0];
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
              double ms0, ks0;
              lin_sums(NX + f, row, ms0, ks0);
              y[P::in(f)] -= L.Mp[P::in(f)] * ms0;
            }
          }
          if (store) {
#pragma unroll
            for (int aa = 0; aa < P::NF; ++aa) A.out[aa * m.n_owned + i] = y[aa];
          }
        }
      }
#pragma unroll
      for (int o = 0; o < P::NOUT; ++o) acc_hi[o] = hi[o];
    }
  }
}

// Boundary vertices of the owned slab (one layer) through the general per-vertex operator of pfh_vertex.cuh, gathered
// from global memory: what the cube-centric kernel skips.  Index space as in k3_scalar_shell.
template <int NF>
struct ShellFields {
  FieldRef U[NF], X[NF], U0[NF];
};
__device__ __forceinline__ int jshell_index(int j, int n) { return (2 >= n || j < 1) ? j : n - 2 + j; }

template <class P, bool JAC>
__global__ void __launch_bounds__(128) k3_jac_shell(P prob, Mesh3 m, LinCoefs L, ShellFields<P::NF> F, double* __restrict__ out, int nzs,
                                                    int zs0, int zs1, int zi0, int nzi, long long n_threads) {
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= n_threads) return;
  const int csx = min(2, m.nxv), csy = min(2, m.nyv);
  const long long n1 = (long long)nzs * m.nxy, n2 = (long long)nzi * csy * m.nxv;
  int ix, iy, zl;
  if (t < n1) {
    const int j = (int)(t / m.nxy);
    const long long r = t - (long long)j * m.nxy;
    zl = j == 0 ? zs0 : zs1;
    iy = (int)(r / m.nxv);
    ix = (int)(r - (long long)iy * m.nxv);
  } else if (t < n1 + n2) {
    const long long r = t - n1, per = (long long)csy * m.nxv;
    zl = zi0 + (int)(r / per);
    const long long q = r % per;
    iy = jshell_index((int)(q / m.nxv), m.nyv);
    ix = (int)(q % m.nxv);
  } else {
    const long long r = t - n1 - n2, per = (long long)(m.nyv - csy) * csx;
    zl = zi0 + (int)(r / per);
    const long long q = r % per;
    iy = 1 + (int)(q / csx);
    ix = jshell_index((int)(q % csx), m.nxv);
  }
  double o[P::NF];
  vtx::vertex_op3<P, JAC>(prob, m, L, F.U, F.X, F.U0, ix, iy, zl, o);
  const long long i = (long long)zl * m.nxy + (long long)iy * m.nxv + ix;
#pragma unroll
  for (int a = 0; a < P::NF; ++a) out[a * m.n_owned + i] = o[a];
}

constexpr int J_TX = 32, J_TY = 8, J_NT = 320;

template <class P, int KIND, int NS>
constexpr size_t jac_smem() {
  constexpr int nfld = (KIND == 1 ? 1 : P::NF) + P::NIN;
  constexpr int stage = round_up((J_TX + 4) * (J_TY + 2) * 8, 128);
  constexpr int ncube = (J_TX + 1) * (J_TY + 1);
  return (size_t)nfld * NS * stage + (size_t)2 * P::NOUT * 8 * (ncube | 1) * 8 + NS * sizeof(uint64_t);
}

template <class P, int KIND, int NS, int OCC, bool F32Q>
void jac_set_attr() {
  PFH_CUDA(cudaFuncSetAttribute(k3_jac_tma<P, KIND, J_TX, J_TY, J_NT, NS, OCC, true, F32Q>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)jac_smem<P, KIND, NS>()));
  PFH_CUDA(cudaFuncSetAttribute(k3_jac_tma<P, KIND, J_TX, J_TY, J_NT, NS, OCC, false, F32Q>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)jac_smem<P, KIND, NS>()));
}

// operand f of the ring: `base` (owned planes) with one-plane ghost arrays glo / ghi
struct JacOperand {
  const double *base, *glo, *ghi;
};

template <class P, int KIND, int NS, int OCC, bool F32Q = false>
bool jac_launch(pfh_ctx* c, const LinCoefs& L, const JacOperand* ops, double* out) {
  constexpr int NFLD = (KIND == 1 ? 1 : P::NF) + P::NIN;
  const Mesh3& m = c->m3;
  const bool lo = c->nplane > 0 && c->cfg.rank > 0, hi = c->nplane > 0 && c->cfg.rank < c->cfg.nranks - 1;
  bool use_tma = c->use_tma_loader && (m.nxv % 2) == 0 && m.nzl >= 1;
  for (int f = 0; f < NFLD; ++f) {
    use_tma = use_tma && aligned16(ops[f].base);
    if (lo) use_tma = use_tma && aligned16(ops[f].glo);
    if (hi) use_tma = use_tma && aligned16(ops[f].ghi);
  }
  MapSet<NFLD> maps;
  JacParams<NFLD> A;
  memset(&maps, 0, sizeof maps);
  memset(&A, 0, sizeof A);
  A.m = m;
  A.L = L;
  A.has_lo = lo;
  A.has_hi = hi;
  const int target = c->jac_zchunk > 0 ? c->jac_zchunk : 32;
  const int nch = (m.nzl + target - 1) / target;
  A.zchunk = (m.nzl + nch - 1) / nch;
  const P prob(c->prob);
  if constexpr (P::QUARTIC) {
    double a3 = 0.0, a2 = 0.0, a1 = 0.0, a0 = 0.0;
    if (KIND == 2) prob.res_poly(a3, a2, a1, a0);
    else prob.jac_poly(a2, a1, a0);
    A.qj = elem::QuarticJac(a2, a1, a0, m.detJ);
    A.qr = elem::QuarticRes(a3, a2, a1, a0, m.detJ);
  }
  A.out = out;
  for (int f = 0; f < NFLD; ++f) {
    A.base[f] = ops[f].base;
    A.glo[f] = ops[f].glo;
    A.ghi[f] = ops[f].ghi;
    if (use_tma) {
      maps.main[f] = pfh_tensor_map(c, ops[f].base, m.nxv, m.nyv, m.nzl, J_TX + 4, J_TY + 2);
      maps.lo[f] = lo ? pfh_tensor_map(c, ops[f].glo, m.nxv, m.nyv, 1, J_TX + 4, J_TY + 2) : maps.main[f];
      maps.hi[f] = hi ? pfh_tensor_map(c, ops[f].ghi, m.nxv, m.nyv, 1, J_TX + 4, J_TY + 2) : maps.main[f];
    }
  }
  // boundary shell of the slab (KIND 0 / 2), on aux_stream underneath the cube-centric kernel
  bool fork = false;
  if constexpr (KIND != 1) {
    ShellFields<P::NF> F;
    memset(&F, 0, sizeof F);
    auto ref = [&](const JacOperand& o) { return FieldRef{o.base, lo ? o.glo : nullptr, hi ? o.ghi : nullptr}; };
    for (int b = 0; b < P::NF; ++b) {
      F.U[b] = F.X[b] = F.U0[b] = ref(ops[b]);          // operand fields: x for J.v, u for the residual
    }
    for (int f = 0; f < P::NIN; ++f) {
      if (KIND == 0) F.U[P::in(f)] = ref(ops[P::NF + f]);  // linearisation point
      else F.U0[P::in(f)] = ref(ops[P::NF + f]);           // previous step
    }
    int zs[2] = {0, 0}, nzs = 0, zi0 = 0, nzi = 0;
    for (int zl = 0; zl < m.nzl; ++zl) {
      const int gz = m.z0 + zl;
      if (gz == 0 || gz == m.nzv - 1) {
        if (nzs < 2) zs[nzs] = zl;
        ++nzs;
      } else {
        if (nzi == 0) zi0 = zl;
        ++nzi;
      }
    }
    PFH_CHECK(nzs <= 2, "internal: more than two boundary planes in a slab");
    const long long csx = std::min(2, m.nxv), csy = std::min(2, m.nyv);
    const long long n_threads = (long long)nzs * m.nxy + (long long)nzi * (csy * m.nxv + (m.nyv - csy) * csx);
    if (n_threads > 0) {
      fork = c->shell_async && c->aux_stream != nullptr;
      cudaStream_t sst = fork ? c->aux_stream : c->stream;
      if (fork) {
        PFH_CUDA(cudaEventRecord(c->ev_fork, c->stream));
        PFH_CUDA(cudaStreamWaitEvent(c->aux_stream, c->ev_fork, 0));
      }
      const unsigned nb = (unsigned)((n_threads + 127) / 128);
      k3_jac_shell<P, KIND == 0><<<nb, 128, 0, sst>>>(prob, m, L, F, out, nzs, zs[0], zs[1], zi0, nzi, n_threads);
      PFH_CUDA(cudaGetLastError());
      c->launches++;
      if (fork) PFH_CUDA(cudaEventRecord(c->ev_join, c->aux_stream));
    }
  }
  const dim3 grid((m.nxv + J_TX - 1) / J_TX, (m.nyv + J_TY - 1) / J_TY, (m.nzl + A.zchunk - 1) / A.zchunk);
  constexpr size_t smem = jac_smem<P, KIND, NS>();
  if (use_tma) k3_jac_tma<P, KIND, J_TX, J_TY, J_NT, NS, OCC, true, F32Q><<<grid, J_NT, smem, c->stream>>>(maps, A, prob);
  else k3_jac_tma<P, KIND, J_TX, J_TY, J_NT, NS, OCC, false, F32Q><<<grid, J_NT, smem, c->stream>>>(maps, A, prob);
  PFH_CUDA(cudaGetLastError());
  if (fork) PFH_CUDA(cudaStreamWaitEvent(c->stream, c->ev_join, 0));
  return true;
}

// ring depth / CTAs per SM of the shipped instantiations: 4 stages and 3 CTAs per SM (73 KB each)
constexpr int J_NS = 4, J_OCC = 3;

}  // namespace

void pfh_jac3_init_device() {
  jac_set_attr<ProbCH, 0, J_NS, J_OCC, false>();
  jac_set_attr<ProbCH, 0, J_NS, J_OCC, true>();
  jac_set_attr<ProbCH, 0, 5, 2, false>();
  jac_set_attr<ProbCH, 1, J_NS, J_OCC, false>();
  jac_set_attr<ProbCH, 2, J_NS, J_OCC, false>();
}

// the field offsets of a state vector must keep the rows 8-byte aligned only; 16-byte alignment is checked per launch
static JacOperand state_operand(const pfh_ctx* c, const double* v, const double* glo, const double* ghi, int fld) {
  return JacOperand{v + (long long)fld * c->m3.n_owned, glo ? glo + (long long)fld * c->nplane : nullptr,
                    ghi ? ghi + (long long)fld * c->nplane : nullptr};
}

// y = J(ulin) x for Cahn-Hilliard on the Kuhn mesh; ghosts: gX_* for x, gU_* for ulin
bool k3_jac_tma_launch(pfh_ctx* c, const LinCoefs& L, const double* ulin, const double* x, double* y) {
  if (!pfh_fast3_ok(c) || c->prob.id != PFH_PROBLEM_CH) return false;
  using P = ProbCH;
  JacOperand ops[P::NF + P::NIN];
  for (int f = 0; f < P::NF; ++f) ops[f] = state_operand(c, x, c->gX_lo, c->gX_hi, f);
  for (int f = 0; f < P::NIN; ++f) ops[P::NF + f] = state_operand(c, ulin, c->gU_lo, c->gU_hi, P::in(f));
  if (c->jac_quadrature_fp32) return jac_launch<P, 0, J_NS, J_OCC, true>(c, L, ops, y);
  if (c->jac_variant == 1) return jac_launch<P, 0, 5, 2>(c, L, ops, y);
  return jac_launch<P, 0, J_NS, J_OCC>(c, L, ops, y);
}

// out = (Mc[mu][c] M + Kc[mu][c] K) xc + M_{g'}(ulin) xc : the (mu, c) block of the Cahn-Hilliard Jacobian applied to a
// scalar field (ghosts of xc in `xc.glo/ghi`, of ulin in gU_*)
bool k3_ch_crow_tma_launch(pfh_ctx* c, const LinCoefs& L, const double* ulin, FieldRef xc, double* out) {
  if (!pfh_fast3_ok(c) || c->prob.id != PFH_PROBLEM_CH) return false;
  using P = ProbCH;
  JacOperand ops[1 + P::NIN];
  ops[0] = JacOperand{xc.p, xc.glo, xc.ghi};
  ops[1] = state_operand(c, ulin, c->gU_lo, c->gU_hi, P::in(0));
  return jac_launch<P, 1, J_NS, J_OCC>(c, L, ops, out);
}

// R = R(u; u0, dt) for Cahn-Hilliard on the Kuhn mesh (ghosts: gU_* for u, gU0_* for u0)
bool k3_res_tma_launch(pfh_ctx* c, const LinCoefs& L, const double* u, const double* u0, double* R) {
  if (!pfh_fast3_ok(c) || c->prob.id != PFH_PROBLEM_CH) return false;
  using P = ProbCH;
  for (int a = 0; a < P::NF; ++a) {  // the kernel subtracts Mp_a M u0_a only for the nonlinear input fields
    bool is_in = false;
    for (int f = 0; f < P::NIN; ++f) is_in |= P::in(f) == a;
    if (L.Mp[a] != 0.0 && !is_in) return false;
  }
  JacOperand ops[P::NF + P::NIN];
  for (int f = 0; f < P::NF; ++f) ops[f] = state_operand(c, u, c->gU_lo, c->gU_hi, f);
  for (int f = 0; f < P::NIN; ++f) ops[P::NF + f] = state_operand(c, u0, c->gU0_lo, c->gU0_hi, P::in(f));
  return jac_launch<P, 2, J_NS, J_OCC>(c, L, ops, R);
}
