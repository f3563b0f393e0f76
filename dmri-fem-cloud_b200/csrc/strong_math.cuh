// Element arithmetic of the strongly periodic (transformed) equation, shared by the kernels of strong.cu and by a
// host-compiled test harness (tests/strong_host_check.cu) that checks it against the oracle on the CPU.
#pragma once

#include <math.h>
#include <stdint.h>

#ifdef __CUDACC__
#define BT_HD __host__ __device__
#else
#define BT_HD
#endif

// measure |T| and the gradients of the barycentric functions of cell t (tetrahedron or triangle in R^3)
BT_HD inline double cell_geometry(const double* xyz, const int32_t* cells, int64_t t,
                                       int cell_nv, double (&g)[4][3]) {
  double x[4][3];
  for (int k = 0; k < cell_nv; ++k)
    for (int d = 0; d < 3; ++d) x[k][d] = xyz[3 * (int64_t)cells[4 * t + k] + d];
  if (cell_nv == 4) {
    double e[3][3];
    for (int k = 0; k < 3; ++k)
      for (int d = 0; d < 3; ++d) e[k][d] = x[k + 1][d] - x[0][d];
    double c[4][3];
    c[1][0] = e[1][1] * e[2][2] - e[1][2] * e[2][1];
    c[1][1] = e[1][2] * e[2][0] - e[1][0] * e[2][2];
    c[1][2] = e[1][0] * e[2][1] - e[1][1] * e[2][0];
    c[2][0] = e[2][1] * e[0][2] - e[2][2] * e[0][1];
    c[2][1] = e[2][2] * e[0][0] - e[2][0] * e[0][2];
    c[2][2] = e[2][0] * e[0][1] - e[2][1] * e[0][0];
    c[3][0] = e[0][1] * e[1][2] - e[0][2] * e[1][1];
    c[3][1] = e[0][2] * e[1][0] - e[0][0] * e[1][2];
    c[3][2] = e[0][0] * e[1][1] - e[0][1] * e[1][0];
    const double det = e[0][0] * c[1][0] + e[0][1] * c[1][1] + e[0][2] * c[1][2];
    const double inv = 1.0 / det;
    for (int d = 0; d < 3; ++d) {
      g[1][d] = c[1][d] * inv; g[2][d] = c[2][d] * inv; g[3][d] = c[3][d] * inv;
      g[0][d] = -(g[1][d] + g[2][d] + g[3][d]);
    }
    return fabs(det) / 6.0;
  }
  double e1[3], e2[3], nn[3];
  for (int d = 0; d < 3; ++d) { e1[d] = x[1][d] - x[0][d]; e2[d] = x[2][d] - x[0][d]; }
  nn[0] = e1[1] * e2[2] - e1[2] * e2[1];
  nn[1] = e1[2] * e2[0] - e1[0] * e2[2];
  nn[2] = e1[0] * e2[1] - e1[1] * e2[0];
  const double n2 = nn[0] * nn[0] + nn[1] * nn[1] + nn[2] * nn[2];
  const double inv = 1.0 / n2;
  g[1][0] = (e2[1] * nn[2] - e2[2] * nn[1]) * inv;
  g[1][1] = (e2[2] * nn[0] - e2[0] * nn[2]) * inv;
  g[1][2] = (e2[0] * nn[1] - e2[1] * nn[0]) * inv;
  g[2][0] = (nn[1] * e1[2] - nn[2] * e1[1]) * inv;
  g[2][1] = (nn[2] * e1[0] - nn[0] * e1[2]) * inv;
  g[2][2] = (nn[0] * e1[1] - nn[1] * e1[0]) * inv;
  for (int d = 0; d < 3; ++d) { g[0][d] = -(g[1][d] + g[2][d]); g[3][d] = 0.0; }
  return 0.5 * sqrt(n2);
}

// D g and D^T g of cell t
BT_HD inline void cell_Dg(int dkind, const double* D, int64_t t, const double gd[3], double (&Dg)[3],
                               double (&DTg)[3]) {
  if (dkind == 2) {
    const double* Dt = D + 9 * t;
    for (int a = 0; a < 3; ++a) {
      Dg[a] = Dt[3 * a] * gd[0] + Dt[3 * a + 1] * gd[1] + Dt[3 * a + 2] * gd[2];
      DTg[a] = Dt[a] * gd[0] + Dt[3 + a] * gd[1] + Dt[6 + a] * gd[2];
    }
  } else {
    const double d0 = dkind == 0 ? D[0] : D[t];
    for (int a = 0; a < 3; ++a) Dg[a] = DTg[a] = d0 * gd[a];
  }
}


// one cell contribution (i, j) to W and C
BT_HD inline void strong_cell_wc(const double* xyz, const int32_t* cells, int64_t t, int cell_nv, int i, int j, int dkind,
                                 const double* D, const double gd[3], double* w, double* c) {
  const double dm = (double)(cell_nv - 1);
  double g[4][3];
  const double vol = cell_geometry(xyz, cells, t, cell_nv, g);
  double Dg[3], DTg[3];
  cell_Dg(dkind, D, t, gd, Dg, DTg);
  const double gDg = gd[0] * Dg[0] + gd[1] * Dg[1] + gd[2] * Dg[2];
  *w = gDg * vol * (i == j ? 2.0 : 1.0) / ((dm + 1.0) * (dm + 2.0));
  *c = vol / (dm + 1.0) * ((Dg[0] + DTg[0]) * g[j][0] + (Dg[1] + DTg[1]) * g[j][1] + (Dg[2] + DTg[2]) * g[j][2]);
}

// facet of cell t opposite its local vertex lf:  (Dg.n_out) int_F phi_a phi_b = coef * (1 + d_ab),
// coef = -|T| (Dg . grad lambda_lf) / (d + 1)
BT_HD inline double strong_facet_coef(const double* xyz, const int32_t* cells, int64_t t, int cell_nv, int lf, int dkind,
                                      const double* D, const double gd[3]) {
  double g[4][3];
  const double vol = cell_geometry(xyz, cells, t, cell_nv, g);
  double Dg[3], DTg[3];
  cell_Dg(dkind, D, t, gd, Dg, DTg);
  return -vol * (Dg[0] * g[lf][0] + Dg[1] * g[lf][1] + Dg[2] * g[lf][2]) / (double)cell_nv;
}

// SELL operator values of one nonzero for the step with scalars aA = theta (q F_n)^2, aP = theta (q F_p)^2
BT_HD inline void strong_combine_entry(double mk, double k0t /* theta (S+R+I) */, double w, double gv, double aA, double aP,
                                       double di, double* p, double* q, double* j) {
  *p = (mk + k0t + aA * w) * di;
  *q = (mk - k0t - aP * w) * di;
  *j = gv * di;
}
