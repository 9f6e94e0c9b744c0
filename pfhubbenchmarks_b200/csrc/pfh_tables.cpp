// Stencil rows of the consistent P1 mass matrix M and stiffness matrix K on the reference's
// structured meshes, built by assembling element matrices over the cells incident to one
// vertex of each boundary type.  This is what DOLFIN's assembler produces for the
// `u*v*dx` / `inner(grad(u),grad(v))*dx` terms of dolfin/pfbase.py:361-460 (both are integrated
// exactly by the degree-3 rule the drivers select, so no quadrature appears here).
#include <cmath>

#include "pfh_internal.h"

namespace {

// P1 element mass / stiffness of a simplex with vertices X[0..d] (d = 2 or 3)
void p1_element(int d, const double X[4][3], double Me[4][4], double Ke[4][4]) {
  double J[3][3];
  for (int i = 0; i < d; ++i)
    for (int j = 0; j < d; ++j) J[i][j] = X[j + 1][i] - X[0][i];  // columns = edge vectors
  double det, Jinv[3][3];
  if (d == 2) {
    det = J[0][0] * J[1][1] - J[0][1] * J[1][0];
    Jinv[0][0] = J[1][1] / det; Jinv[0][1] = -J[0][1] / det;
    Jinv[1][0] = -J[1][0] / det; Jinv[1][1] = J[0][0] / det;
  } else {
    det = J[0][0] * (J[1][1] * J[2][2] - J[1][2] * J[2][1]) - J[0][1] * (J[1][0] * J[2][2] - J[1][2] * J[2][0]) +
          J[0][2] * (J[1][0] * J[2][1] - J[1][1] * J[2][0]);
    Jinv[0][0] = (J[1][1] * J[2][2] - J[1][2] * J[2][1]) / det;
    Jinv[0][1] = (J[0][2] * J[2][1] - J[0][1] * J[2][2]) / det;
    Jinv[0][2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) / det;
    Jinv[1][0] = (J[1][2] * J[2][0] - J[1][0] * J[2][2]) / det;
    Jinv[1][1] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) / det;
    Jinv[1][2] = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) / det;
    Jinv[2][0] = (J[1][0] * J[2][1] - J[1][1] * J[2][0]) / det;
    Jinv[2][1] = (J[0][1] * J[2][0] - J[0][0] * J[2][1]) / det;
    Jinv[2][2] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) / det;
  }
  const double vol = std::fabs(det) / (d == 2 ? 2.0 : 6.0);
  // physical gradients: grad phi_a = Jinv^T * gref_a ; gref_0 = -1, gref_{a>0} = e_{a-1}
  double g[4][3];
  for (int a = 0; a <= d; ++a)
    for (int i = 0; i < d; ++i) {
      double s = 0.0;
      for (int k = 0; k < d; ++k) {
        const double gref = (a == 0) ? -1.0 : (a - 1 == k ? 1.0 : 0.0);
        s += gref * Jinv[k][i];
      }
      g[a][i] = s;
    }
  for (int a = 0; a <= d; ++a)
    for (int b = 0; b <= d; ++b) {
      double s = 0.0;
      for (int i = 0; i < d; ++i) s += g[a][i] * g[b][i];
      Ke[a][b] = s * vol;
      Me[a][b] = vol * (a == b ? 2.0 : 1.0) / ((d + 1) * (d + 2));
    }
}

const int KUHN[6][4] = {{0, 1, 3, 7}, {0, 1, 7, 5}, {0, 5, 7, 4}, {0, 3, 2, 7}, {0, 6, 4, 7}, {0, 2, 6, 7}};

int off3(int dx, int dy, int dz) {
  static const int tab[15][3] = {{0, 0, 0}, {1, 0, 0}, {-1, 0, 0}, {0, 1, 0}, {0, -1, 0}, {0, 0, 1}, {0, 0, -1},
                                 {1, 1, 0}, {-1, -1, 0}, {0, 1, 1}, {0, -1, -1}, {1, 0, 1}, {-1, 0, -1},
                                 {1, 1, 1}, {-1, -1, -1}};
  for (int o = 0; o < 15; ++o)
    if (tab[o][0] == dx && tab[o][1] == dy && tab[o][2] == dz) return o;
  return -1;
}

}  // namespace

void pfh_build_tables3(Mesh3& m, std::vector<double>& mT, std::vector<double>& kT) {
  mT.assign(27 * PFH_NOFF3, 0.0);
  kT.assign(27 * PFH_NKOFF3, 0.0);
  const double h[3] = {m.hx, m.hy, m.hz};
  for (int tz = 0; tz < 3; ++tz)
    for (int ty = 0; ty < 3; ++ty)
      for (int tx = 0; tx < 3; ++tx) {
        const int type = tz * 9 + ty * 3 + tx;
        double mrow[15] = {0}, krow[15] = {0};
        const int t[3] = {tx, ty, tz};
        for (int c = 0; c < 8; ++c) {
          const int d[3] = {-(c & 1), -((c >> 1) & 1), -((c >> 2) & 1)};  // cube offset in {-1,0}^3
          bool exists = true;
          for (int a = 0; a < 3; ++a) {
            if (d[a] == -1 && t[a] == 0) exists = false;
            if (d[a] == 0 && t[a] == 2) exists = false;
          }
          if (!exists) continue;
          const int k = (-d[0]) + 2 * (-d[1]) + 4 * (-d[2]);  // local index of our vertex in that cube
          for (int tet = 0; tet < 6; ++tet) {
            int row = -1;
            for (int a = 0; a < 4; ++a)
              if (KUHN[tet][a] == k) row = a;
            if (row < 0) continue;
            double X[4][3], Me[4][4], Ke[4][4];
            for (int a = 0; a < 4; ++a)
              for (int i = 0; i < 3; ++i) X[a][i] = ((KUHN[tet][a] >> i) & 1) * h[i];
            p1_element(3, X, Me, Ke);
            for (int b = 0; b < 4; ++b) {
              const int l = KUHN[tet][b];
              const int o = off3((l & 1) - (k & 1), ((l >> 1) & 1) - ((k >> 1) & 1), ((l >> 2) & 1) - ((k >> 2) & 1));
              PFH_CHECK(o >= 0, "internal: unexpected Kuhn stencil offset");
              mrow[o] += Me[row][b];
              krow[o] += Ke[row][b];
            }
          }
        }
        double kscale = 0.0;
        for (int o = 0; o < 15; ++o) kscale = std::fmax(kscale, std::fabs(krow[o]));
        for (int o = 0; o < 15; ++o) {
          mT[type * PFH_NOFF3 + o] = mrow[o];
          if (o < PFH_NKOFF3)
            kT[type * PFH_NKOFF3 + o] = krow[o];
          else
            PFH_CHECK(std::fabs(krow[o]) <= 1e-12 * kscale, "Kuhn stiffness has diagonal-edge entries");
        }
      }
  for (int o = 0; o < PFH_NOFF3; ++o) m.mI[o] = mT[13 * PFH_NOFF3 + o];
  for (int o = 0; o < PFH_NKOFF3; ++o) m.kI[o] = kT[13 * PFH_NKOFF3 + o];
}

void pfh_mg_build_levels(const Mesh2& fine, std::vector<Mesh2>& meshes, std::vector<std::vector<double>>& tables) {
  meshes.clear();
  tables.clear();
  Mesh2 cur = fine;
  while (cur.nx % 2 == 0 && cur.ny % 2 == 0 && cur.nx >= 4 && cur.ny >= 4) {
    Mesh2 m = cur;
    m.nx = cur.nx / 2; m.ny = cur.ny / 2; m.nxv = m.nx + 1; m.nyv = m.ny + 1;
    m.ncorner = (long long)m.nxv * m.nyv; m.ncentre = (long long)m.nx * m.ny; m.n_owned = m.ncorner + m.ncentre;
    m.hx = 2.0 * cur.hx; m.hy = 2.0 * cur.hy; m.detJ = 0.5 * m.hx * m.hy;
    std::vector<double> mT, kT;
    pfh_build_tables2(m, mT, kT);
    tables.emplace_back(mT);
    tables.back().insert(tables.back().end(), kT.begin(), kT.end());
    meshes.push_back(m);
    cur = m;
  }
  // the vectors no longer move: fix the table pointers
  for (size_t l = 0; l < meshes.size(); ++l) {
    meshes[l].mT = tables[l].data();
    meshes[l].kT = tables[l].data() + tables[l].size() / 2;
  }
}

void pfh_build_tables2(Mesh2& m, std::vector<double>& mT, std::vector<double>& kT) {
  mT.assign(9 * PFH_NOFF2, 0.0);
  kT.assign(9 * PFH_NOFF2, 0.0);
  const double hx = m.hx, hy = m.hy;
  // local square vertices: 0 v0, 1 v1, 2 v2, 3 v3, 4 centre
  const double P[5][2] = {{0, 0}, {hx, 0}, {0, hy}, {hx, hy}, {0.5 * hx, 0.5 * hy}};
  const int TRI[4][3] = {{0, 1, 4}, {0, 2, 4}, {1, 3, 4}, {2, 3, 4}};
  double Me[4][4][4], Ke[4][4][4];
  for (int t = 0; t < 4; ++t) {
    double X[4][3] = {{0}};
    for (int a = 0; a < 3; ++a) {
      X[a][0] = P[TRI[t][a]][0];
      X[a][1] = P[TRI[t][a]][1];
    }
    p1_element(2, X, Me[t], Ke[t]);
  }
  // centre rows
  for (int o = 0; o < 5; ++o) m.mC[o] = m.kC[o] = 0.0;
  for (int t = 0; t < 4; ++t)
    for (int b = 0; b < 3; ++b) {
      const int l = TRI[t][b];
      const int o = (l == 4) ? 0 : 1 + l;
      m.mC[o] += Me[t][2][b];
      m.kC[o] += Ke[t][2][b];
    }
  // corner rows by type
  for (int ty = 0; ty < 3; ++ty)
    for (int tx = 0; tx < 3; ++tx) {
      const int type = ty * 3 + tx;
      double* mrow = &mT[type * PFH_NOFF2];
      double* krow = &kT[type * PFH_NOFF2];
      for (int s = 0; s < 4; ++s) {  // squares NE, NW, SE, SW
        const int dx = -(s & 1), dy = -((s >> 1) & 1);
        if ((dx == -1 && tx == 0) || (dx == 0 && tx == 2) || (dy == -1 && ty == 0) || (dy == 0 && ty == 2)) continue;
        const int k = (-dx) + 2 * (-dy);  // our corner's local index in that square
        for (int t = 0; t < 4; ++t) {
          int row = -1;
          for (int a = 0; a < 3; ++a)
            if (TRI[t][a] == k) row = a;
          if (row < 0) continue;
          for (int b = 0; b < 3; ++b) {
            const int l = TRI[t][b];
            int o;
            if (l == 4) {
              o = 5 + s;
            } else if (l == k) {
              o = 0;
            } else {
              const int ex = (l & 1) - (k & 1), ey = ((l >> 1) & 1) - ((k >> 1) & 1);
              PFH_CHECK((ex == 0) != (ey == 0), "internal: unexpected crossed-mesh edge");
              o = ex == 1 ? 1 : ex == -1 ? 2 : ey == 1 ? 3 : 4;
            }
            mrow[o] += Me[t][row][b];
            krow[o] += Ke[t][row][b];
          }
        }
      }
    }
  for (int o = 0; o < PFH_NOFF2; ++o) {
    m.mI[o] = mT[4 * PFH_NOFF2 + o];
    m.kI[o] = kT[4 * PFH_NOFF2 + o];
  }
}
