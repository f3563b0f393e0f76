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

/* returns 0 or a negative reason; u is in/out (interleaved complex, n entries) */
int btcpu_theta_loop(int n, const int* rp, const int* ci, const double* M, const double* K0, const double* B,
                     const double* Jg, double dt, double theta, int nsteps, const double* cA, const double* cb,
                     double rtol, double atol, int maxit, int mode, double* u_io, int* iters_out) {
  const int nnz = rp[n];
  cplx* u = (cplx*)u_io;
  double* P = (double*)malloc(sizeof(double) * nnz);
  double* Q = (double*)malloc(sizeof(double) * nnz);
  double* kinv = (double*)malloc(sizeof(double) * n);
  double* ones = (double*)malloc(sizeof(double) * n);
  cplx *r = malloc(sizeof(cplx) * n), *rhat = malloc(sizeof(cplx) * n), *p = malloc(sizeof(cplx) * n),
       *v = malloc(sizeof(cplx) * n), *s = malloc(sizeof(cplx) * n), *t = malloc(sizeof(cplx) * n),
       *x = malloc(sizeof(cplx) * n);
  int rc = 0;
  for (int i = 0; i < n; ++i) ones[i] = 1.0;
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
    const double bnorm = sqrt(rdot(n, r, r));
    const double ttol = fmax(rtol * bnorm, atol);
    double dp = bnorm;
    int it = 0;
    memset(x, 0, sizeof(cplx) * n);
    if (dp > ttol) {
      memcpy(rhat, r, sizeof(cplx) * n);
      memset(p, 0, sizeof(cplx) * n);
      memset(v, 0, sizeof(cplx) * n);
      double rhoold = 1.0, alpha = 1.0, omegaold = 1.0;
      const double c = theta * cA[step];
      for (;;) {
        const double rho = rdot(n, r, rhat);
        const double beta = (rho / rhoold) * (alpha / omegaold);
        const double ob = omegaold * beta;
#pragma omp parallel for schedule(static)
        for (int i = 0; i < n; ++i) {
          p[i].re = r[i].re - ob * v[i].re + beta * p[i].re;
          p[i].im = r[i].im - ob * v[i].im + beta * p[i].im;
        }
        spmv(n, rp, ci, P, Jg, c, kinv, p, v);
        const double d1 = rdot(n, v, rhat);
        if (d1 == 0.0) { rc = -4; break; }
        alpha = rho / d1;
#pragma omp parallel for schedule(static)
        for (int i = 0; i < n; ++i) {
          s[i].re = r[i].re - alpha * v[i].re;
          s[i].im = r[i].im - alpha * v[i].im;
        }
        spmv(n, rp, ci, P, Jg, c, kinv, s, t);
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
        if (!(dp == dp) || isinf(dp)) { rc = -5; break; }
        if (dp <= ttol) break;
        if (dp >= 1e4 * bnorm) { rc = -6; break; }
        if (rho == 0.0 || omega == 0.0) { rc = -4; break; }
        if (it >= maxit) { rc = -3; break; }
      }
    }
    memcpy(u, x, sizeof(cplx) * n);
    if (iters_out) iters_out[step] = it;
  }
  free(P); free(Q); free(kinv); free(ones);
  free(r); free(rhat); free(p); free(v); free(s); free(t); free(x);
  return rc;
}
