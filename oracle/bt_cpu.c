/* TEST INFRASTRUCTURE ONLY (oracle/): plain-C + OpenMP restatement of the reference's time loop,
 * used as the checker at sizes the numpy oracle is too slow for and as the timed CPU baseline
 * (bench.py `cpu_baseline` / `--impl reference`).  Never linked into the product.
 *
 * Restates  MRI_simulation.solve, DmriFemLib.py:897-910:
 *     A = 1/k*M + assemble(F); b = assemble(L); linsolver.solve(A, u, b)
 * in the pre-assembled form of the comri solvers (comri/one-comp/hpc-fenics-cpp/main.cpp:296-328):
 *     A_n = P + i*theta*cA[n]*Jg ,  b_n = (Q - i*(1-theta)*cb[n]*Jg) u^n
 * and PETSc 3.7 KSPSolve_BCGS + PCJACOBI + KSPConvergedDefault (third party; algorithm restated,
 * see oracle/bt_oracle.py:bicgstab_petsc -- real inner products on the (re,im)-split system,
 * left preconditioning, test ||K^-1 r|| <= max(rtol*||K^-1 b||, atol)).
 * All matrices real CSR on one pattern (int32), vectors interleaved complex.
 *
 * btcpu_theta_loop_reassemble: the work pattern of DmriFemLib.solve itself:
 *         EVERY time step A = 1/k*M + assemble(F) and b = assemble(L) are re-assembled from element
 *         integrals (cells + interface facets), the Jacobi diagonal is re-extracted, then the Krylov solve.
 *         Element integrals use the closed forms (cheaper than the reference's FFC quadrature kernels), so
 *         timings of this mode are a LOWER bound on what FEniCS spends.
 * mode 0: operators pre-combined once (what the comri C++ solvers do).
 * mode 1: "reference-faithful" work pattern: P, Q and the Jacobi diagonal are re-formed from
 *         M, K0 (= S+R+I), B every time step, like `assemble` + a fresh PC each step do
 *         (DmriFemLib.py:904-907); arithmetic result identical to mode 0.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct { double re, im; } cplx;

int btcpu_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

void btcpu_set_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

/* y = Kinv .* (V + i*c*J) x */
static void spmv(int n, const int* rp, const int* ci, const double* V, const double* J, double c,
                 const double* kinv, const cplx* x, cplx* y) {
#pragma omp parallel for schedule(static)
  for (int r = 0; r < n; ++r) {
    double ar = 0.0, ai = 0.0;
    for (int k = rp[r]; k < rp[r + 1]; ++k) {
      const double a = V[k], b = c * J[k];
      const cplx xv = x[ci[k]];
      ar += a * xv.re - b * xv.im;
      ai += a * xv.im + b * xv.re;
    }
    y[r].re = ar * kinv[r];
    y[r].im = ai * kinv[r];
  }
}

static double rdot(int n, const cplx* a, const cplx* b) {
  double s = 0.0;
#pragma omp parallel for reduction(+ : s) schedule(static)
  for (int i = 0; i < n; ++i) s += a[i].re * b[i].re + a[i].im * b[i].im;
  return s;
}

/* One Jacobi-BiCGStab solve (zero initial guess) of K^-1 (V + i c J) x = r0 where r0 = K^-1 b is passed in r.
 * Work vectors w[5] = rhat, p, v, s, t.  Returns iterations (>= 0) or a negative reason. */
static int bicgstab(int n, const int* rp, const int* ci, const double* V, const double* J, double c,
                    const double* kinv, cplx* r, cplx* x, cplx** w, double rtol, double atol, int maxit) {
  cplx *rhat = w[0], *p = w[1], *v = w[2], *s = w[3], *t = w[4];
  const double bnorm = sqrt(rdot(n, r, r));
  const double ttol = fmax(rtol * bnorm, atol);
  double dp = bnorm;
  int it = 0;
  memset(x, 0, sizeof(cplx) * n);
  if (dp <= ttol) return 0;
  memcpy(rhat, r, sizeof(cplx) * n);
  memset(p, 0, sizeof(cplx) * n);
  memset(v, 0, sizeof(cplx) * n);
  double rhoold = 1.0, alpha = 1.0, omegaold = 1.0;
  for (;;) {
    const double rho = rdot(n, r, rhat);
    const double beta = (rho / rhoold) * (alpha / omegaold);
    const double ob = omegaold * beta;
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; ++i) {
      p[i].re = r[i].re - ob * v[i].re + beta * p[i].re;
      p[i].im = r[i].im - ob * v[i].im + beta * p[i].im;
    }
    spmv(n, rp, ci, V, J, c, kinv, p, v);
    const double d1 = rdot(n, v, rhat);
    if (d1 == 0.0) return -4;
    alpha = rho / d1;
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; ++i) {
      s[i].re = r[i].re - alpha * v[i].re;
      s[i].im = r[i].im - alpha * v[i].im;
    }
    spmv(n, rp, ci, V, J, c, kinv, s, t);
    const double st = rdot(n, s, t), tt = rdot(n, t, t);
    const double omega = tt == 0.0 ? 0.0 : st / tt;
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; ++i) {
      x[i].re += alpha * p[i].re + omega * s[i].re;
      x[i].im += alpha * p[i].im + omega * s[i].im;
      r[i].re = s[i].re - omega * t[i].re;
      r[i].im = s[i].im - omega * t[i].im;
    }
    dp = sqrt(rdot(n, r, r));
    rhoold = rho;
    omegaold = omega;
    ++it;
    if (!(dp == dp) || isinf(dp)) return -5;
    if (dp <= ttol) return it;
    if (dp >= 1e4 * bnorm) return -6;
    if (rho == 0.0 || omega == 0.0) return -4;
    if (it >= maxit) return -3;
  }
}

static inline int find_col(const int* ci, int lo, int hi, int c) {
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (ci[mid] < c) lo = mid + 1; else hi = mid;
  }
  return lo;
}

/* DmriFemLib.solve work pattern: per step assemble A (Are + i*Aim) and b from the elements, then solve.
 * Scalar D and 1/T2; interface facets given as dof sextuples (d0[3], d1[3]) with coef = kappa*area/12. */
int btcpu_theta_loop_reassemble(int n, const int* rp, const int* ci, int nc, const double* xyz, const int* tets,
                                const int* cell_dofs, double D, double invT2, int nif, const int* if_dofs,
                                const double* if_coef, const double* g, double dt, double theta, int nsteps,
                                const double* cA, const double* cb, double rtol, double atol, int maxit,
                                double* u_io, int* iters_out) {
  const int nnz = rp[n];
  cplx* u = (cplx*)u_io;
  double* Are = (double*)malloc(sizeof(double) * nnz);
  double* Aim = (double*)malloc(sizeof(double) * nnz);
  double* kinv = (double*)malloc(sizeof(double) * n);
  cplx* b = malloc(sizeof(cplx) * n);
  cplx* x = malloc(sizeof(cplx) * n);
  cplx* w[5];
  for (int i = 0; i < 5; ++i) w[i] = malloc(sizeof(cplx) * n);
  int rc = 0;
  for (int step = 0; step < nsteps && rc == 0; ++step) {
    const double ca = theta * cA[step], cbb = -(1.0 - theta) * cb[step];
    memset(Are, 0, sizeof(double) * nnz);
    memset(Aim, 0, sizeof(double) * nnz);
    memset(b, 0, sizeof(cplx) * n);
#pragma omp parallel for schedule(static)
    for (int c = 0; c < nc; ++c) {
      const int* tv = tets + 4 * c;
      const int* dof = cell_dofs + 4 * c;
      double X[4][3], e[3][3], cf[4][3];
      for (int k = 0; k < 4; ++k)
        for (int d = 0; d < 3; ++d) X[k][d] = xyz[3 * tv[k] + d];
      for (int k = 0; k < 3; ++k)
        for (int d = 0; d < 3; ++d) e[k][d] = X[k + 1][d] - X[0][d];
      cf[1][0] = e[1][1] * e[2][2] - e[1][2] * e[2][1]; cf[1][1] = e[1][2] * e[2][0] - e[1][0] * e[2][2]; cf[1][2] = e[1][0] * e[2][1] - e[1][1] * e[2][0];
      cf[2][0] = e[2][1] * e[0][2] - e[2][2] * e[0][1]; cf[2][1] = e[2][2] * e[0][0] - e[2][0] * e[0][2]; cf[2][2] = e[2][0] * e[0][1] - e[2][1] * e[0][0];
      cf[3][0] = e[0][1] * e[1][2] - e[0][2] * e[1][1]; cf[3][1] = e[0][2] * e[1][0] - e[0][0] * e[1][2]; cf[3][2] = e[0][0] * e[1][1] - e[0][1] * e[1][0];
      const double det = e[0][0] * cf[1][0] + e[0][1] * cf[1][1] + e[0][2] * cf[1][2];
      for (int d = 0; d < 3; ++d) cf[0][d] = -(cf[1][d] + cf[2][d] + cf[3][d]);
      const double vol = fabs(det) / 6.0, inv = 1.0 / det;
      double gx[4], sg = 0.0;
      for (int k = 0; k < 4; ++k) { gx[k] = g[0] * X[k][0] + g[1] * X[k][1] + g[2] * X[k][2]; sg += gx[k]; }
      cplx ul[4];
      for (int k = 0; k < 4; ++k) ul[k] = u[dof[k]];
      for (int i = 0; i < 4; ++i) {
        double bre = 0.0, bim = 0.0;
        const int r0 = rp[dof[i]], r1 = rp[dof[i] + 1];
        for (int j = 0; j < 4; ++j) {
          const double m = vol * (i == j ? 2.0 : 1.0) / 20.0;
          const double sij = vol * D * inv * inv * (cf[i][0] * cf[j][0] + cf[i][1] * cf[j][1] + cf[i][2] * cf[j][2]);
          const double jij = (i == j) ? vol * (2.0 * sg + 4.0 * gx[i]) / 120.0 : vol * (sg + gx[i] + gx[j]) / 120.0;
          const double k0 = sij + invT2 * m;
          const int pos = find_col(ci, r0, r1, dof[j]);
#pragma omp atomic
          Are[pos] += m / dt + theta * k0;
#pragma omp atomic
          Aim[pos] += ca * jij;
          const double qre = m / dt - (1.0 - theta) * k0, qim = cbb * jij;      /* (Q + i cb' J) u */
          bre += qre * ul[j].re - qim * ul[j].im;
          bim += qre * ul[j].im + qim * ul[j].re;
        }
#pragma omp atomic
        b[dof[i]].re += bre;
#pragma omp atomic
        b[dof[i]].im += bim;
      }
    }
#pragma omp parallel for schedule(static)
    for (int f = 0; f < nif; ++f) {            /* kappa (u0-u1)(v0-v1) on interface facets */
      const int* d = if_dofs + 6 * f;
      for (int blk = 0; blk < 4; ++blk) {
        const int rs = (blk == 1 || blk == 3) ? 3 : 0, cs = (blk == 1 || blk == 2) ? 3 : 0;
        const double sgn = blk < 2 ? 1.0 : -1.0;
        for (int i = 0; i < 3; ++i) {
          const int row = d[rs + i];
          double bre = 0.0, bim = 0.0;
          for (int j = 0; j < 3; ++j) {
            const int col = d[cs + j];
            const double v = sgn * if_coef[f] * (i == j ? 2.0 : 1.0);
            const int pos = find_col(ci, rp[row], rp[row + 1], col);
#pragma omp atomic
            Are[pos] += theta * v;
            bre -= (1.0 - theta) * v * u[col].re;
            bim -= (1.0 - theta) * v * u[col].im;
          }
#pragma omp atomic
          b[row].re += bre;
#pragma omp atomic
          b[row].im += bim;
        }
      }
    }
#pragma omp parallel for schedule(static)
    for (int row = 0; row < n; ++row) {        /* PCJACOBI set-up on the new matrix + left preconditioning of b */
      const int pos = find_col(ci, rp[row], rp[row + 1], row);
      kinv[row] = 1.0 / Are[pos];
      b[row].re *= kinv[row];
      b[row].im *= kinv[row];
    }
    const int it = bicgstab(n, rp, ci, Are, Aim, 1.0, kinv, b, x, w, rtol, atol, maxit);
    if (it < 0) { rc = it; break; }
    memcpy(u, x, sizeof(cplx) * n);
    if (iters_out) iters_out[step] = it;
  }
  free(Are); free(Aim); free(kinv); free(b); free(x);
  for (int i = 0; i < 5; ++i) free(w[i]);
  return rc;
}

/* returns 0 or a negative reason; u is in/out (interleaved complex, n entries) */
int btcpu_theta_loop(int n, const int* rp, const int* ci, const double* M, const double* K0, const double* B,
                     const double* Jg, double dt, double theta, int nsteps, const double* cA, const double* cb,
                     double rtol, double atol, int maxit, int mode, double* u_io, int* iters_out) {
  const int nnz = rp[n];
  cplx* u = (cplx*)u_io;
  double* P = (double*)malloc(sizeof(double) * nnz);
  double* Q = (double*)malloc(sizeof(double) * nnz);
  double* kinv = (double*)malloc(sizeof(double) * n);
  cplx* r = malloc(sizeof(cplx) * n);
  cplx* x = malloc(sizeof(cplx) * n);
  cplx* w[5];
  for (int i = 0; i < 5; ++i) w[i] = malloc(sizeof(cplx) * n);
  int rc = 0;
  for (int step = 0; step < nsteps && rc == 0; ++step) {
    if (step == 0 || mode == 1) {
#pragma omp parallel for schedule(static)
      for (int row = 0; row < n; ++row) {
        for (int k = rp[row]; k < rp[row + 1]; ++k) {
          P[k] = M[k] / dt + theta * (K0[k] + B[k]);
          Q[k] = M[k] / dt - (1.0 - theta) * K0[k];
          if (ci[k] == row) kinv[row] = 1.0 / P[k];
        }
      }
    }
    /* b^ = K^-1 (Q - i(1-theta) cb J) u ; zero initial guess -> r = b^ */
    spmv(n, rp, ci, Q, Jg, -(1.0 - theta) * cb[step], kinv, u, r);
    const int it = bicgstab(n, rp, ci, P, Jg, theta * cA[step], kinv, r, x, w, rtol, atol, maxit);
    if (it < 0) { rc = it; break; }
    memcpy(u, x, sizeof(cplx) * n);
    if (iters_out) iters_out[step] = it;
  }
  free(P); free(Q); free(kinv); free(r); free(x);
  for (int i = 0; i < 5; ++i) free(w[i]);
  return rc;
}
