/*
 * vlc_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * Plain-C restatement of VOLCANOR's Biot-Savart hot path.  Loop nests, summation
 * order, guards and quirks follow the Fortran reference line by line; every
 * function names the reference file:line it restates.  See vlc_oracle.h for the
 * parity pin.  Build: oracle/Makefile (strict: -O2 -ffp-contract=off; the timed
 * CPU-baseline build adds -march=native -fopenmp like CMakeLists.txt:37-40).
 *
 * Indexing: Fortran arrays are 1-based column-major; the macros below keep the
 * reference's 1-based subscripts so loops read like the original.
 */
#include "vlc_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define WIP(b, i, j) ((b)->wiP[((i)-1) + (size_t)(b)->nc * ((j)-1)])
#define WAN(b, i, j) ((b)->waN[((i)-1) + (size_t)(b)->nNwake * ((j)-1)])
#define WANP(b, i, j) ((b)->waNPredicted[((i)-1) + (size_t)(b)->nNwake * ((j)-1)])
#define WAF(b, i) ((b)->waF[(i)-1])
#define WAFP(b, i) ((b)->waFPredicted[(i)-1])
#define VELN(arr, b, i, j) (&(arr)[3 * (((i)-1) + (size_t)(b)->nNwake * ((j)-1))])
#define VELF(arr, i) (&(arr)[3 * ((i)-1)])

/* classdef.f90:13-14, libMath.f90:9 : pi = atan(1)*4, inv4pi = 0.25/pi, twoPi = 2*pi */
static double orc_pi(void) { return atan(1.0) * 4.0; }

int orc_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* ------------------------------------------------------------------ libMath */

double orc_norm2(const double a[3]) { return sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]); }

/* libMath.f90:249-262 */
void orc_unitVec(const double a[3], double u[3]) {
  double n = orc_norm2(a);
  if (n > ORC_EPS) {
    u[0] = a[0] / n;
    u[1] = a[1] / n;
    u[2] = a[2] / n;
  } else {
    u[0] = u[1] = u[2] = 0.0;
  }
}

/* libMath.f90:202-212 */
void orc_cross(const double a[3], const double b[3], double c[3]) {
  c[0] = a[1] * b[2] - a[2] * b[1];
  c[1] = a[2] * b[0] - a[0] * b[2];
  c[2] = a[0] * b[1] - a[1] * b[0];
}

/* libMath.f90:695-726.  T is column-major: T[r + 3*c]. */
void orc_getTransformAxis(double theta, const double axisVec[3], double T[9]) {
  double n = orc_norm2(axisVec);
  double ax[3] = {axisVec[0] / n, axisVec[1] / n, axisVec[2] / n};
  double ct = cos(theta), st = sin(theta), omct = 1.0 - ct;
  T[0] = ct + ax[0] * ax[0] * omct;
  T[1] = ax[2] * st + ax[1] * ax[0] * omct;
  T[2] = -ax[1] * st + ax[2] * ax[0] * omct;
  T[3] = -ax[2] * st + ax[0] * ax[1] * omct;
  T[4] = ct + ax[1] * ax[1] * omct;
  T[5] = ax[0] * st + ax[2] * ax[1] * omct;
  T[6] = ax[1] * st + ax[0] * ax[2] * omct;
  T[7] = -ax[0] * st + ax[1] * ax[2] * omct;
  T[8] = ct + ax[2] * ax[2] * omct;
}

/* matmul(Tmat, x - origin) + origin  (classdef.f90:638-639, 971-972) */
static void rot_point(const double T[9], const double origin[3], double x[3]) {
  double d[3] = {x[0] - origin[0], x[1] - origin[1], x[2] - origin[2]};
  for (int r = 0; r < 3; ++r) x[r] = (T[r] * d[0] + T[r + 3] * d[1] + T[r + 6] * d[2]) + origin[r];
}

/*
 * libMath.f90:48-83 inv2 = LAPACK DGETRF + DGETRI.  LAPACK is a third-party
 * dependency that is not vendored in the reference (find_package(LAPACK),
 * unpinned system library).  Restated here from the published algorithms:
 * DGETF2-style right-looking LU with partial (row) pivoting, then DGETRI:
 * invert U in place (DTRTRI, upper, non-unit), then solve inv(A)*L = inv(U)
 * column by column from the right and undo the row interchanges as column
 * swaps.  Returns LAPACK's info (0 ok, >0 singular pivot index).
 */
int orc_inv2(int n, const double *A, double *Ainv) {
  int *ipiv = (int *)malloc(sizeof(int) * (size_t)n);
  double *work = (double *)malloc(sizeof(double) * (size_t)n);
  double *a = Ainv;
  memcpy(a, A, sizeof(double) * (size_t)n * n);
#define AA(i, j) a[(i) + (size_t)n * (j)]
  /* DGETF2 */
  for (int j = 0; j < n; ++j) {
    int p = j;
    double big = fabs(AA(j, j));
    for (int i = j + 1; i < n; ++i)
      if (fabs(AA(i, j)) > big) {
        big = fabs(AA(i, j));
        p = i;
      }
    ipiv[j] = p;
    if (AA(p, j) == 0.0) {
      free(ipiv);
      free(work);
      return j + 1;
    }
    if (p != j)
      for (int k = 0; k < n; ++k) {
        double t = AA(j, k);
        AA(j, k) = AA(p, k);
        AA(p, k) = t;
      }
    double rp = 1.0 / AA(j, j);
    for (int i = j + 1; i < n; ++i) AA(i, j) *= rp;
    for (int k = j + 1; k < n; ++k) {
      double ajk = AA(j, k);
      if (ajk != 0.0)
        for (int i = j + 1; i < n; ++i) AA(i, k) -= AA(i, j) * ajk;
    }
  }
  /* DTRTI2 (upper, non-unit): inv(U) in place */
  for (int j = 0; j < n; ++j) {
    AA(j, j) = 1.0 / AA(j, j);
    double ajj = -AA(j, j);
    /* x := U(0:j,0:j)^-1-so-far * U(0:j, j)  (DTRMV upper, no-trans, non-unit) */
    for (int i = 0; i < j; ++i) {
      double t = 0.0;
      for (int k = i; k < j; ++k) t += AA(i, k) * AA(k, j);
      work[i] = t;
    }
    for (int i = 0; i < j; ++i) AA(i, j) = work[i] * ajj;
  }
  /* DGETRI unblocked: for j = n-1..0: copy L(:,j) to work, zero it, A(:,j) -= A(:,j+1:) * work(j+1:) */
  for (int j = n - 2; j >= 0; --j) {
    for (int i = j + 1; i < n; ++i) {
      work[i] = AA(i, j);
      AA(i, j) = 0.0;
    }
    for (int k = j + 1; k < n; ++k) {
      double w = work[k];
      if (w != 0.0)
        for (int i = 0; i < n; ++i) AA(i, j) -= AA(i, k) * w;
    }
  }
  /* undo interchanges: columns */
  for (int j = n - 2; j >= 0; --j) {
    int p = ipiv[j];
    if (p != j)
      for (int i = 0; i < n; ++i) {
        double t = AA(i, j);
        AA(i, j) = AA(i, p);
        AA(i, p) = t;
      }
  }
#undef AA
  free(ipiv);
  free(work);
  return 0;
}

/* libMath.f90:105-122  DGEMV('N'): AX = A*X, reference BLAS loop order (axpy by column). */
void orc_matmulAX(int m, int n, const double *A, const double *X, double *AX) {
  for (int i = 0; i < m; ++i) AX[i] = 0.0;
  for (int j = 0; j < n; ++j) {
    double t = X[j];
    for (int i = 0; i < m; ++i) AX[i] += t * A[i + (size_t)m * j];
  }
}

/* -------------------------------------------------------------- pair kernel */

/* classdef.f90:476-503  vf_vind: unit-strength filament, Vatistas n=2 core. */
void orc_vf_vind(const orc_vf_t *f, const double P[3], double v[3]) {
  const double inv4pi = 0.25 / orc_pi();
  double r1[3], r2[3], r0[3], c[3], u1[3], u2[3], du[3];
  for (int k = 0; k < 3; ++k) {
    r1[k] = P[k] - f->fc[0][k]; /* :486 */
    r2[k] = P[k] - f->fc[1][k]; /* :487 */
    r0[k] = r1[k] - r2[k];      /* :488 */
  }
  c[0] = r1[1] * r2[2] - r1[2] * r2[1]; /* :491-493 */
  c[1] = r1[2] * r2[0] - r1[0] * r2[2];
  c[2] = r1[0] * r2[1] - r1[1] * r2[0];
  double c2 = c[0] * c[0] + c[1] * c[1] + c[2] * c[2]; /* :494 */
  v[0] = v[1] = v[2] = 0.0;                             /* :496 */
  if (c2 > ORC_EPS * ORC_EPS) {                         /* :498 */
    orc_unitVec(r1, u1);
    orc_unitVec(r2, u2);
    for (int k = 0; k < 3; ++k) du[k] = u1[k] - u2[k];
    double d = r0[0] * du[0] + r0[1] * du[1] + r0[2] * du[2];
    /* :500-501  (r1Xr2*inv4pi*dot)/sqrt((rVc*|r0|)**4._dp + c2**2._dp); **4._dp is a pow() call */
    double den = sqrt(pow(f->rVc * orc_norm2(r0), 4.0) + c2 * c2);
    for (int k = 0; k < 3; ++k) v[k] = ((c[k] * inv4pi) * d) / den;
  }
}

/* classdef.f90:527-542  vr_vind: vindMat(i,:) then sum(vindMat,1) in order i=1..4 */
void orc_vr_vind(const orc_vr_t *r, const double P[3], double v[3]) {
  double m[4][3];
  for (int i = 0; i < 4; ++i) orc_vf_vind(&r->vf[i], P, m[i]);
  for (int k = 0; k < 3; ++k) v[k] = ((m[0][k] + m[1][k]) + m[2][k]) + m[3][k];
}

/* ------------------------------------------------------------- source loops */

/* classdef.f90:1342-1357 */
void orc_blade_vind_bywing(const orc_blade_t *b, const double P[3], double v[3]) {
  double t[3];
  v[0] = v[1] = v[2] = 0.0;
  for (int j = 1; j <= b->ns; ++j)
    for (int i = 1; i <= b->nc; ++i) {
      const orc_vr_t *vr = &WIP(b, i, j).vr;
      orc_vr_vind(vr, P, t);
      for (int k = 0; k < 3; ++k) v[k] = v[k] + t[k] * vr->gam;
    }
}

/* classdef.f90:1376-1396 */
void orc_blade_vind_bywing_boundVortices(const orc_blade_t *b, const double P[3], double v[3]) {
  double t2[3], t4[3];
  v[0] = v[1] = v[2] = 0.0;
  for (int j = 1; j <= b->ns; ++j)
    for (int i = 1; i <= b->nc; ++i) {
      const orc_vr_t *vr = &WIP(b, i, j).vr;
      orc_vf_vind(&vr->vf[1], P, t2);
      orc_vf_vind(&vr->vf[3], P, t4);
      for (int k = 0; k < 3; ++k) v[k] = v[k] + (t2[k] + t4[k]) * vr->gam;
    }
  for (int j = 1; j <= b->ns; ++j) {
    const orc_vr_t *vr = &WIP(b, b->nc, j).vr;
    orc_vf_vind(&vr->vf[1], P, t2);
    for (int k = 0; k < 3; ++k) v[k] = v[k] - t2[k] * vr->gam;
  }
}

/* classdef.f90:1398-1418 */
void orc_blade_vind_bywing_chordwiseVortices(const orc_blade_t *b, const double P[3], double v[3]) {
  double t1[3], t3[3];
  v[0] = v[1] = v[2] = 0.0;
  for (int j = 1; j <= b->ns; ++j)
    for (int i = 1; i <= b->nc; ++i) {
      const orc_vr_t *vr = &WIP(b, i, j).vr;
      orc_vf_vind(&vr->vf[0], P, t1);
      orc_vf_vind(&vr->vf[2], P, t3);
      for (int k = 0; k < 3; ++k) v[k] = v[k] + (t1[k] + t3[k]) * vr->gam;
    }
  for (int j = 1; j <= b->ns; ++j) {
    const orc_vr_t *vr = &WIP(b, b->nc, j).vr;
    orc_vf_vind(&vr->vf[1], P, t1);
    for (int k = 0; k < 3; ++k) v[k] = v[k] + t1[k] * vr->gam;
  }
}

/* classdef.f90:1420-1435 */
void orc_blade_vind_boundVortex(const orc_blade_t *b, int ic, int is, const double P[3], double v[3]) {
  double t4[3], t2[3];
  const orc_vr_t *vr = &WIP(b, ic, is).vr;
  orc_vf_vind(&vr->vf[3], P, t4);
  if (ic > 1) {
    const orc_vr_t *vm = &WIP(b, ic - 1, is).vr;
    orc_vf_vind(&vm->vf[1], P, t2);
    for (int k = 0; k < 3; ++k) v[k] = t4[k] * vr->gam + t2[k] * vm->gam;
  } else {
    for (int k = 0; k < 3; ++k) v[k] = t4[k] * vr->gam;
  }
}

/* classdef.f90:1437-1513.  predicted != 0 <=> optionalChar == 'P'. */
void orc_blade_vind_bywake(const orc_blade_t *b, int rowNear, int rowFar, const double P[3], int predicted,
                           double v[3]) {
  const int nNwake = b->nNwake, nFwake = b->nFwake; /* :1446-1447 */
  const orc_vr_t *waN = predicted ? b->waNPredicted : b->waN;
  const orc_fwake_t *waF = predicted ? b->waFPredicted : b->waF;
  const orc_fwake_t *wapF = predicted ? b->wapFPredicted : b->wapF;
  double t[3];
  v[0] = v[1] = v[2] = 0.0;
  for (int j = 1; j <= b->ns; ++j) /* :1450-1456 / :1480-1486 */
    for (int i = rowNear; i <= nNwake; ++i) {
      const orc_vr_t *vr = &waN[(i - 1) + (size_t)nNwake * (j - 1)];
      if (fabs(vr->gam) > ORC_EPS) {
        orc_vr_vind(vr, P, t);
        for (int k = 0; k < 3; ++k) v[k] = v[k] + t[k] * vr->gam;
      }
    }
  if (rowFar <= nFwake) { /* :1458 */
    for (int j = 1; j <= b->ns; ++j) { /* horseshoe correction :1460-1463 */
      const orc_vr_t *vr = &waN[(nNwake - 1) + (size_t)nNwake * (j - 1)];
      orc_vf_vind(&vr->vf[1], P, t);
      for (int k = 0; k < 3; ++k) v[k] = v[k] - t[k] * vr->gam;
    }
    for (int i = rowFar; i <= nFwake; ++i) { /* :1465-1469 */
      const orc_fwake_t *f = &waF[i - 1];
      if (fabs(f->gam) > ORC_EPS) {
        orc_vf_vind(&f->vf, P, t);
        for (int k = 0; k < 3; ++k) v[k] = v[k] + t[k] * f->gam;
      }
    }
    for (int i = 1; i <= ORC_NPFWAKE; ++i) { /* :1471-1476 */
      const orc_fwake_t *f = &wapF[i - 1];
      if (fabs(f->gam) > ORC_EPS) {
        orc_vf_vind(&f->vf, P, t);
        for (int k = 0; k < 3; ++k) v[k] = v[k] + t[k] * f->gam;
      }
    }
  }
}

/* classdef.f90:4424-4443 (surfaceType 2 = source panels returns 0: :544-554) */
void orc_rotor_vind_bywing(const orc_rotor_t *r, const double P[3], double v[3]) {
  double t[3];
  v[0] = v[1] = v[2] = 0.0;
  if (abs(r->surfaceType) == 1)
    for (int ib = 0; ib < r->nb; ++ib) {
      orc_blade_vind_bywing(&r->blade[ib], P, t);
      for (int k = 0; k < 3; ++k) v[k] = v[k] + t[k];
    }
}

/* classdef.f90:4445-4457 */
void orc_rotor_vind_bywing_boundVortices(const orc_rotor_t *r, const double P[3], double v[3]) {
  double t[3];
  v[0] = v[1] = v[2] = 0.0;
  for (int ib = 0; ib < r->nb; ++ib) {
    orc_blade_vind_bywing_boundVortices(&r->blade[ib], P, t);
    for (int k = 0; k < 3; ++k) v[k] = v[k] + t[k];
  }
}

/* classdef.f90:4459-4479 */
void orc_rotor_vind_bywake(const orc_rotor_t *r, const double P[3], int predicted, double v[3]) {
  double t[3];
  v[0] = v[1] = v[2] = 0.0;
  for (int ib = 0; ib < r->nb; ++ib) {
    orc_blade_vind_bywake(&r->blade[ib], r->rowNear, r->rowFar, P, predicted, t);
    for (int k = 0; k < 3; ++k) v[k] = v[k] + t[k];
  }
}

/* ------------------------------------------------------------ target sweeps */

/* libCommon.f90:114-171 */
void orc_vind_onNwake_byRotor(const orc_rotor_t *src, const orc_vr_t *Nwake, int rows, int cols, int ld,
                              int predicted, double *out) {
#ifdef _OPENMP
#pragma omp parallel for collapse(2) schedule(runtime)
#endif
  for (int j = 1; j <= cols; ++j)
    for (int i = 1; i <= rows; ++i) { /* :133-138 */
      double a[3], w[3];
      const double *P = Nwake[(i - 1) + (size_t)ld * (j - 1)].vf[1].fc[0];
      orc_rotor_vind_bywing(src, P, a);
      orc_rotor_vind_bywake(src, P, predicted, w);
      double *o = &out[3 * ((i - 1) + (size_t)rows * (j - 1))];
      for (int k = 0; k < 3; ++k) o[k] = a[k] + w[k];
    }
#ifdef _OPENMP
#pragma omp parallel for schedule(runtime)
#endif
  for (int i = 1; i <= rows; ++i) { /* :142-145 */
    double a[3], w[3];
    const double *P = Nwake[(i - 1) + (size_t)ld * (cols - 1)].vf[2].fc[0];
    orc_rotor_vind_bywing(src, P, a);
    orc_rotor_vind_bywake(src, P, predicted, w);
    double *o = &out[3 * ((i - 1) + (size_t)rows * cols)];
    for (int k = 0; k < 3; ++k) o[k] = a[k] + w[k];
  }
}

/* libCommon.f90:173-211 */
void orc_vind_onFwake_byRotor(const orc_rotor_t *src, const orc_fwake_t *Fwake, int rows, int predicted,
                              double *out) {
#ifdef _OPENMP
#pragma omp parallel for schedule(runtime)
#endif
  for (int i = 1; i <= rows; ++i) {
    double a[3], w[3];
    const double *P = Fwake[i - 1].vf.fc[0];
    orc_rotor_vind_bywing(src, P, a);
    orc_rotor_vind_bywake(src, P, predicted, w);
    for (int k = 0; k < 3; ++k) out[3 * (i - 1) + k] = a[k] + w[k];
  }
}

/* libCommon.f90:213-235.  All arrays (3, rows, cols). */
void orc_vel_order2_Nwake(const double *vn, const double *vnp1, int rows, int cols, double *out) {
#define V3(a, i, j) (&(a)[3 * (((i)-1) + (size_t)rows * ((j)-1))])
  for (int j = 1; j <= cols; ++j) {
    for (int k = 0; k < 3; ++k) V3(out, 1, j)[k] = (V3(vnp1, 1, j)[k] + V3(vn, 1, j)[k]) * 0.5;
    for (int i = 2; i <= rows - 1; ++i)
      for (int k = 0; k < 3; ++k)
        V3(out, i, j)[k] =
            (((V3(vnp1, i, j)[k] + V3(vnp1, i - 1, j)[k]) + V3(vn, i + 1, j)[k]) + V3(vn, i, j)[k]) * 0.25;
    for (int k = 0; k < 3; ++k) V3(out, rows, j)[k] = (V3(vnp1, rows, j)[k] + V3(vn, rows, j)[k]) * 0.5;
  }
#undef V3
}

/* libCommon.f90:237-258 */
void orc_vel_order2_Fwake(const double *vn, const double *vnp1, int rows, double *out) {
  if (rows < 1) return;
  for (int k = 0; k < 3; ++k) out[k] = (vnp1[k] + vn[k]) * 0.5;
  for (int i = 2; i <= rows - 1; ++i)
    for (int k = 0; k < 3; ++k)
      VELF(out, i)[k] = (((VELF(vnp1, i)[k] + VELF(vnp1, i - 1)[k]) + VELF(vn, i + 1)[k]) + VELF(vn, i)[k]) * 0.25;
  for (int k = 0; k < 3; ++k) VELF(out, rows)[k] = (VELF(vnp1, rows)[k] + VELF(vn, rows)[k]) * 0.5;
}

/* ------------------------------------------------------ vr / fwake helpers */

/* classdef.f90:569-592 */
void orc_vr_assignP(orc_vr_t *r, int n, const double P[3]) {
  static const int a[5] = {0, 3, 0, 1, 2}; /* filament whose fc(:,2) is corner n */
  static const int b[5] = {0, 0, 1, 2, 3}; /* filament whose fc(:,1) is corner n */
  for (int k = 0; k < 3; ++k) {
    r->vf[a[n]].fc[1][k] = P[k];
    r->vf[b[n]].fc[0][k] = P[k];
  }
}

/* classdef.f90:594-624 */
void orc_vr_shiftdP(orc_vr_t *r, int n, const double d[3]) {
  static const int a[5] = {0, 3, 0, 1, 2};
  static const int b[5] = {0, 0, 1, 2, 3};
  if (n == 0) {
    for (int f = 0; f < 4; ++f)
      for (int k = 0; k < 3; ++k) {
        r->vf[f].fc[0][k] += d[k];
        r->vf[f].fc[1][k] += d[k];
      }
    return;
  }
  for (int k = 0; k < 3; ++k) {
    r->vf[a[n]].fc[1][k] = r->vf[a[n]].fc[1][k] + d[k];
    r->vf[b[n]].fc[0][k] = r->vf[b[n]].fc[0][k] + d[k];
  }
}

/* classdef.f90:505-515 */
void orc_vf_calclength(orc_vf_t *f, int isOriginal) {
  double d[3] = {f->fc[0][0] - f->fc[1][0], f->fc[0][1] - f->fc[1][1], f->fc[0][2] - f->fc[1][2]};
  f->lc = orc_norm2(d);
  if (isOriginal) f->l0 = orc_norm2(d);
}

static void vr_rot(orc_vr_t *r, const double T[9], const double origin[3]) { /* classdef.f90:626-642 */
  for (int i = 0; i < 4; ++i) {
    rot_point(T, origin, r->vf[i].fc[0]);
    rot_point(T, origin, r->vf[i].fc[1]);
  }
}

/* ------------------------------------------------------- wake state update */

/* classdef.f90:1515-1575.  NOTE the reference quirk: the predicted
 * near-wake loops are `do i = 1, rowNear, nNwake` (start 1, END rowNear, STRIDE
 * nNwake), so only row 1 is shifted. Replicated, not fixed. */
void orc_blade_convectwake(orc_blade_t *b, int rowNear, int rowFar, double dt, char wakeType, int ductSwitch) {
  const int nNwake = b->nNwake, nFwake = b->nFwake;
  double d[3];
  if (wakeType == 'C') {
    for (int j = 1; j <= b->ns; ++j)
      for (int i = rowNear; i <= nNwake; ++i) { /* :1529-1533 */
        const double *vel = VELN(b->velNwake, b, i, j);
        for (int k = 0; k < 3; ++k) d[k] = vel[k] * dt;
        orc_vr_shiftdP(&WAN(b, i, j), 2, d);
      }
    for (int i = rowNear; i <= nNwake; ++i) { /* :1537-1539 */
      const double *vel = VELN(b->velNwake, b, i, b->ns + 1);
      for (int k = 0; k < 3; ++k) d[k] = vel[k] * dt;
      orc_vr_shiftdP(&WAN(b, i, b->ns), 3, d);
    }
    for (int i = rowFar; i <= nFwake; ++i) { /* :1544-1546  shift only TE: fc(:,1) */
      const double *vel = VELF(b->velFwake, i);
      for (int k = 0; k < 3; ++k) WAF(b, i).vf.fc[0][k] = WAF(b, i).vf.fc[0][k] + vel[k] * dt;
    }
  } else { /* 'P' */
    for (int j = 1; j <= b->ns; ++j)
      for (int i = 1; i <= rowNear; i += nNwake) { /* :1552 (quirk) */
        const double *vel = VELN(b->velNwake, b, i, j);
        for (int k = 0; k < 3; ++k) d[k] = vel[k] * dt;
        orc_vr_shiftdP(&WANP(b, i, j), 2, d);
      }
    for (int i = 1; i <= rowNear; i += nNwake) { /* :1559 (quirk) */
      const double *vel = VELN(b->velNwake, b, i, b->ns + 1);
      for (int k = 0; k < 3; ++k) d[k] = vel[k] * dt;
      orc_vr_shiftdP(&WANP(b, i, b->ns), 3, d);
    }
    for (int i = rowFar; i <= nFwake; ++i) { /* :1566-1568 */
      const double *vel = VELF(b->velFwake, i);
      for (int k = 0; k < 3; ++k) WAFP(b, i).vf.fc[0][k] = WAFP(b, i).vf.fc[0][k] + vel[k] * dt;
    }
  }
  orc_blade_wake_continuity(b, rowNear, rowFar, wakeType, ductSwitch);
}

/* classdef.f90:1609-1702 */
void orc_blade_wake_continuity(orc_blade_t *b, int rowNear, int rowFar, char wakeType, int ductSwitch) {
  const int nNwake = b->nNwake, nFwake = b->nFwake, ns = b->ns;
  orc_vr_t *waN = (wakeType == 'C') ? b->waN : b->waNPredicted;
  orc_fwake_t *waF = (wakeType == 'C') ? b->waF : b->waFPredicted;
#define W(i, j) waN[((i)-1) + (size_t)nNwake * ((j)-1)]
  for (int j = 1; j <= ns - 1; ++j)
    for (int i = rowNear + 1; i <= nNwake; ++i) { /* :1623-1629 / :1669-1675 */
      orc_vr_assignP(&W(i, j), 1, W(i - 1, j).vf[1].fc[0]);
      orc_vr_assignP(&W(i, j), 3, W(i, j + 1).vf[1].fc[0]);
      orc_vr_assignP(&W(i, j), 4, W(i - 1, j + 1).vf[1].fc[0]);
    }
  for (int j = 1; j <= ns - 1; ++j) /* :1633-1635 */
    orc_vr_assignP(&W(rowNear, j), 3, W(rowNear, j + 1).vf[1].fc[0]);
  for (int i = rowNear + 1; i <= nNwake; ++i) { /* :1639-1642 */
    orc_vr_assignP(&W(i, ns), 1, W(i - 1, ns).vf[1].fc[0]);
    orc_vr_assignP(&W(i, ns), 4, W(i - 1, ns).vf[2].fc[0]);
  }
  if (wakeType == 'C' && ductSwitch == 1) /* :1645-1656 (only in the 'C' branch) */
    for (int i = rowNear + 1; i <= nNwake; ++i) {
      orc_vr_assignP(&W(i, ns), 4, W(i, 1).vf[0].fc[0]);
      orc_vr_assignP(&W(i, ns), 3, W(i, 1).vf[0].fc[1]);
    }
  for (int i = rowFar + 1; i <= nFwake; ++i) /* :1660-1662 / :1693-1695 */
    for (int k = 0; k < 3; ++k) waF[i - 1].vf.fc[1][k] = waF[i - 2].vf.fc[0][k];
#undef W
}

/* classdef.f90:1283-1340 */
static void blade_rot_wake_axis(orc_blade_t *b, double theta, const double axisVec[3], const double origin[3],
                                int rowNear, int rowFar, char wakeType) {
  if (fabs(theta) > ORC_EPS) {
    double T[9];
    orc_getTransformAxis(theta, axisVec, T);
    orc_vr_t *waN = (wakeType == 'C') ? b->waN : b->waNPredicted;
    orc_fwake_t *waF = (wakeType == 'C') ? b->waF : b->waFPredicted;
    for (int j = 1; j <= b->ns; ++j)
      for (int i = rowNear; i <= b->nNwake; ++i) vr_rot(&waN[(i - 1) + (size_t)b->nNwake * (j - 1)], T, origin);
    for (int i = rowFar; i <= b->nFwake; ++i) {
      rot_point(T, origin, waF[i - 1].vf.fc[0]);
      rot_point(T, origin, waF[i - 1].vf.fc[1]);
    }
  }
}

/* classdef.f90:998-1066 pFwake_update: a helix of 10 revolutions in 240 filaments of 15 degrees, fitted to the mean radius
 * and pitch of the far-wake filaments waF(1:nFwake) handed in, relaxed against the previous fit, attached to the last of
 * them.  PARITY UNPINNED (no shipped case, no reference test).  isClockwiseRotor keeps its default .true. everywhere. */
int orc_pfwake_update(orc_fwake_t *pf, double *helixPitch, double *helixRadius, const orc_fwake_t *waF, int nFwake,
                      const double hubCoords[3], const double shaftAxis[3], double deltaPsi) {
  const double twoPi = 2.0 * orc_pi(), relaxFactor = 0.5, nRevs = 10.0;
  if (fabs(shaftAxis[0]) > ORC_EPS || fabs(shaftAxis[1]) > ORC_EPS) return 1; /* "only implemented for shaft along Z-axis" */
  if (nFwake < 1) return 2;
  double anchor[3] = {waF[nFwake - 1].vf.fc[0][0], waF[nFwake - 1].vf.fc[0][1], waF[nFwake - 1].vf.fc[0][2]};
  double pitchCur = 0.0, radiusCur = 0.0;
  for (int i = 1; i <= nFwake; ++i) {
    const double y = waF[i - 1].vf.fc[0][1], x = waF[i - 1].vf.fc[0][0];
    radiusCur = radiusCur + sqrt(y * y + x * x); /* norm2([fc(2,1), fc(1,1)]) */
    if (i < nFwake) pitchCur = pitchCur + waF[i - 1].vf.fc[0][2] - waF[i].vf.fc[0][2];
  }
  pitchCur = fabs(pitchCur) * (-twoPi / deltaPsi) / (nFwake - 1);
  radiusCur = radiusCur / nFwake;
  *helixPitch = relaxFactor * pitchCur + (1 - relaxFactor) * *helixPitch;
  *helixRadius = relaxFactor * radiusCur + (1 - relaxFactor) * *helixRadius;
  const double dTheta = atan2(anchor[1], anchor[0]);
  const double deltaZ = anchor[2] - hubCoords[2];
  double coords[ORC_NPFWAKE + 1][3];
  const double dx = (twoPi * nRevs - 0.0) / ((ORC_NPFWAKE + 1) - 1); /* linspace, libMath.f90:138-157 */
  for (int i = 0; i <= ORC_NPFWAKE; ++i) {
    double theta = i * dx;
    theta = theta + 0.0;
    theta = -1.0 * theta; /* isClockwiseRotor */
    coords[i][0] = *helixRadius * cos(theta + dTheta);
    coords[i][1] = *helixRadius * sin(theta + dTheta);
    coords[i][2] = *helixPitch * fabs(theta) / twoPi + deltaZ;
  }
  for (int i = 1; i <= ORC_NPFWAKE; ++i)
    for (int k = 0; k < 3; ++k) {
      pf[i - 1].vf.fc[1][k] = hubCoords[k] + coords[i - 1][k]; /* assignP(2, ...) */
      pf[i - 1].vf.fc[0][k] = hubCoords[k] + coords[i][k];     /* assignP(1, ...) */
    }
  for (int k = 0; k < 3; ++k) pf[0].vf.fc[1][k] = anchor[k]; /* continuity with the far wake */
  for (int i = 0; i < ORC_NPFWAKE; ++i) {
    pf[i].gam = waF[nFwake - 1].gam;
    pf[i].vf.rVc = waF[nFwake - 1].vf.rVc;
  }
  return 0;
}

/* classdef.f90:5170-5218 */
int orc_rotor_updatePrescribedWake(orc_rotor_t *r, double dt, char wakeType) {
  const double twoPi = 2.0 * orc_pi();
  const int s = (wakeType == 'C') ? 0 : 1;
  const int rowStart = (r->prescWakeGenNt == 0) ? r->rowFar : r->nFwakeEnd - r->prescWakeGenNt;
  if (rowStart < 1 || rowStart > r->nFwakeEnd) return 2;
  for (int ib = 0; ib < r->nbConvect; ++ib) {
    orc_blade_t *b = &r->blade[ib];
    int rc = orc_pfwake_update(s ? b->wapFPredicted : b->wapF, &b->pfHelixPitch[s], &b->pfHelixRadius[s],
                               (s ? b->waFPredicted : b->waF) + (rowStart - 1), r->nFwakeEnd - rowStart + 1, r->hubCoords,
                               r->shaftAxis, r->omegaSlow * dt);
    if (rc) return rc;
  }
  if (r->axisymmetrySwitch == 1)
    for (int ib = 2; ib <= r->nb; ++ib) {
      const double bladeOffset = twoPi / r->nb * (ib - 1);
      orc_blade_t *b = &r->blade[ib - 1], *b1 = &r->blade[0];
      orc_fwake_t *pf = s ? b->wapFPredicted : b->wapF;
      memcpy(pf, s ? b1->wapFPredicted : b1->wapF, sizeof(orc_fwake_t) * ORC_NPFWAKE);
      b->pfHelixPitch[s] = b1->pfHelixPitch[s];
      b->pfHelixRadius[s] = b1->pfHelixRadius[s];
      if (fabs(bladeOffset) > ORC_EPS) { /* pFwake_rot_wake_axis :1068-1086 */
        double T[9];
        orc_getTransformAxis(bladeOffset, r->shaftAxis, T);
        for (int i = 0; i < ORC_NPFWAKE; ++i) {
          rot_point(T, r->hubCoords, pf[i].vf.fc[0]);
          rot_point(T, r->hubCoords, pf[i].vf.fc[1]);
        }
      }
    }
  return 0;
}

/* classdef.f90:4786-4830 */
void orc_rotor_convectwake(orc_rotor_t *r, int iter, double dt, char wakeType) {
  const double twoPi = 2.0 * orc_pi();
  for (int ib = 0; ib < r->nbConvect; ++ib)
    orc_blade_convectwake(&r->blade[ib], r->rowNear, r->rowFar, dt, wakeType, r->ductSwitch);
  if (r->axisymmetrySwitch == 1) {
    for (int ib = 2; ib <= r->nb; ++ib) {
      double bladeOffset = twoPi / r->nb * (ib - 1);
      orc_blade_t *b = &r->blade[ib - 1], *b1 = &r->blade[0];
      orc_vr_t *dst = (wakeType == 'C') ? b->waN : b->waNPredicted;
      const orc_vr_t *src = (wakeType == 'C') ? b1->waN : b1->waNPredicted;
      for (int j = 1; j <= b->ns; ++j) /* waN(rowNear:, :) = blade(1)%waN(rowNear:, :) */
        for (int i = r->rowNear; i <= b->nNwake; ++i)
          dst[(i - 1) + (size_t)b->nNwake * (j - 1)] = src[(i - 1) + (size_t)b->nNwake * (j - 1)];
      orc_fwake_t *dF = (wakeType == 'C') ? b->waF : b->waFPredicted;
      const orc_fwake_t *sF = (wakeType == 'C') ? b1->waF : b1->waFPredicted;
      for (int i = r->rowFar; i <= b->nFwake; ++i) dF[i - 1] = sF[i - 1];
      blade_rot_wake_axis(b, bladeOffset, r->shaftAxis, r->hubCoords, r->rowNear, r->rowFar, wakeType);
    }
  }
  if (r->prescWakeNt > 0 && iter > r->prescWakeNt) orc_rotor_updatePrescribedWake(r, dt, wakeType); /* :4826-4828 */
}

/* classdef.f90:4297-4325 */
void orc_rotor_assignshed(orc_rotor_t *r, const char *edge) {
  if (edge[0] == 'L') { /* 'LE' */
    for (int ib = 0; ib < r->nb; ++ib) {
      orc_blade_t *b = &r->blade[ib];
      for (int i = 1; i <= r->ns; ++i) {
        orc_vr_t *w = &WAN(b, r->rowNear, i);
        orc_vr_assignP(w, 1, WIP(b, r->nc, i).vr.vf[1].fc[0]);
        orc_vr_assignP(w, 4, WIP(b, r->nc, i).vr.vf[2].fc[0]);
        for (int f = 0; f < 4; ++f) orc_vf_calclength(&w->vf[f], 1);
      }
      for (int i = 1; i <= r->ns; ++i) WAN(b, r->rowNear, i).gam = WIP(b, r->nc, i).vr.gam; /* :4311 */
    }
  } else { /* 'TE' */
    int row = r->rowNear - 1 > 1 ? r->rowNear - 1 : 1;
    for (int ib = 0; ib < r->nb; ++ib) {
      orc_blade_t *b = &r->blade[ib];
      for (int i = 1; i <= r->ns; ++i) {
        orc_vr_assignP(&WAN(b, row, i), 2, WIP(b, r->nc, i).vr.vf[1].fc[0]);
        orc_vr_assignP(&WAN(b, row, i), 3, WIP(b, r->nc, i).vr.vf[2].fc[0]);
      }
    }
  }
}

/* classdef.f90:4331-4354 */
void orc_rotor_age_wake(orc_rotor_t *r, double dt) {
  for (int ib = 0; ib < r->nb; ++ib) {
    orc_blade_t *b = &r->blade[ib];
    for (int f = 0; f < 4; ++f)
      for (int j = 1; j <= r->ns; ++j)
        for (int i = r->rowNear; i <= r->nNwake; ++i) {
          WAN(b, i, j).vf[f].age = WAN(b, i, j).vf[f].age + dt;
          WAN(b, i, j).vf[f].ageAzimuthal = WAN(b, i, j).vf[f].ageAzimuthal + dt * r->omegaSlow;
        }
    for (int i = r->rowFar; i <= r->nFwake; ++i) {
      WAF(b, i).vf.age = WAF(b, i).vf.age + dt;
      WAF(b, i).vf.ageAzimuthal = WAF(b, i).vf.ageAzimuthal + dt * r->omegaSlow;
    }
  }
}

/* classdef.f90:4356-4408 (quirk C2: vf(3).rVc <- vf(1).rVc; vr_decay :662-668; Fwake_decay :975-980) */
void orc_rotor_dissipate_wake(orc_rotor_t *r, double dt, double kinematicViscosity) {
  const double oseenParameter = 1.2564;
  for (int ib = 0; ib < r->nb; ++ib) {
    orc_blade_t *b = &r->blade[ib];
    for (int is = 1; is <= r->ns; ++is) {
      for (int ic = r->rowNear; ic <= r->nNwake; ++ic) { /* :4367-4375 */
        orc_vr_t *w = &WAN(b, ic, is);
        w->vf[0].rVc = sqrt(w->vf[0].rVc * w->vf[0].rVc +
                            4.0 * oseenParameter * r->apparentViscCoeff * kinematicViscosity * dt);
        w->vf[2].rVc = w->vf[0].rVc;
        w->gam = w->gam * exp(-r->decayCoeff * dt);
      }
      for (int ic = r->rowNear; ic <= r->nNwake; ++ic) { /* :4380-4383 */
        orc_vr_t *w = &WAN(b, ic, is);
        w->vf[1].rVc = sqrt(w->vf[1].rVc * w->vf[1].rVc +
                            4.0 * oseenParameter * r->apparentViscCoeff * kinematicViscosity * dt);
      }
      if (r->rowNear != r->nNwake) /* :4386-4392 */
        for (int ic = r->rowNear + 1; ic <= r->nNwake; ++ic) WAN(b, ic, is).vf[3].rVc = WAN(b, ic - 1, is).vf[1].rVc;
    }
    for (int ic = r->rowFar; ic <= r->nFwake; ++ic) { /* :4397-4404 */
      orc_fwake_t *f = &WAF(b, ic);
      f->vf.rVc = sqrt(f->vf.rVc * f->vf.rVc + 4.0 * oseenParameter * r->apparentViscCoeff * kinematicViscosity * dt);
      f->gam = f->gam * exp(-r->decayCoeff * dt);
    }
  }
}

/* classdef.f90:4410-4422, :505-521 (quirk C3: rVc recomputed from rVc0) */
void orc_rotor_strain_wake(orc_rotor_t *r) {
  for (int ib = 0; ib < r->nb; ++ib) {
    orc_blade_t *b = &r->blade[ib];
    for (int i = r->rowFar; i <= r->nFwake; ++i) {
      orc_vf_t *f = &WAF(b, i).vf;
      orc_vf_calclength(f, 0);
      f->rVc = f->rVc0 * sqrt(f->l0 / f->lc);
    }
  }
}

/* classdef.f90:4919-4936 rotor_calc_skew -> :2341-2353 blade_calc_skew -> :737-747 calc_skew -> :704-721 vr_getBimedianCos:
 * skew of a wake ring = |cos| of the angle between its bimedians (0 good, 1 bad), 0 for rings without circulation; an
 * output for skew2file only.  PARITY UNPINNED (skewPlotSwitch = 0 in every shipped case, no reference test). */
double orc_vr_skew(const orc_vr_t *v) {
  if (!(fabs(v->gam) > ORC_EPS)) return 0.0;
  const double *p1 = v->vf[0].fc[0], *p2 = v->vf[1].fc[0], *p3 = v->vf[2].fc[0], *p4 = v->vf[3].fc[0];
  double x1[3], x2[3];
  for (int k = 0; k < 3; ++k) {
    x1[k] = p3[k] + p4[k] - p1[k] - p2[k];
    x2[k] = p4[k] + p1[k] - p2[k] - p3[k];
  }
  const double d12 = x1[0] * x2[0] + x1[1] * x2[1] + x1[2] * x2[2];
  const double d11 = x1[0] * x1[0] + x1[1] * x1[1] + x1[2] * x1[2], d22 = x2[0] * x2[0] + x2[1] * x2[1] + x2[2] * x2[2];
  return fabs(d12 / sqrt(d11 * d22));
}
void orc_rotor_calc_skew(orc_rotor_t *r) {
  for (int ib = 0; ib < r->nbConvect; ++ib)
    for (int j = 1; j <= r->ns; ++j)
      for (int i = r->rowNear; i <= r->nNwake; ++i) WAN(&r->blade[ib], i, j).skew = orc_vr_skew(&WAN(&r->blade[ib], i, j));
  if (r->axisymmetrySwitch == 1)
    for (int ib = 1; ib < r->nb; ++ib)
      for (int j = 1; j <= r->ns; ++j)
        for (int i = r->rowNear; i <= r->nNwake; ++i) WAN(&r->blade[ib], i, j).skew = WAN(&r->blade[0], i, j).skew;
}

/* classdef.f90:4911-4917 rotor_burst_wake -> :2306-2339 blade_burst_wake (far wake only; the near-wake branch is commented
 * out in the source): a kink between successive far filaments beyond skewLimit gives both the core radius `chord`.
 * getAngleCos libMath.f90:238-247.  PARITY UNPINNED: wakeBurst = 0 in every shipped case, no reference test. */
int orc_burst_pair(const orc_fwake_t *f0, const orc_fwake_t *f1, double skewLimit) {
  const double pi = orc_pi();
  double a[3], b[3];
  for (int k = 0; k < 3; ++k) {
    a[k] = f0->vf.fc[1][k] - f0->vf.fc[0][k];
    b[k] = f1->vf.fc[0][k] - f1->vf.fc[1][k];
  }
  const double dot = a[0] * b[0] + a[1] * b[1] + a[2] * b[2];
  const double sa = a[0] * a[0] + a[1] * a[1] + a[2] * a[2], sb = b[0] * b[0] + b[1] * b[1] + b[2] * b[2];
  const double skewVal = fabs(acos(dot / sqrt(sa * sb)) - pi) / pi;
  return skewVal >= skewLimit;
}
void orc_rotor_burst_wake(orc_rotor_t *r) {
  for (int ib = 0; ib < r->nb; ++ib) {
    orc_blade_t *b = &r->blade[ib];
    if (r->rowFar > r->nFwake) continue;
    for (int i = r->rowFar; i <= r->nFwake - 1; ++i)
      if (orc_burst_pair(&WAF(b, i), &WAF(b, i + 1), r->skewLimit)) {
        WAF(b, i + 1).vf.rVc = r->chord;
        WAF(b, i).vf.rVc = r->chord;
      }
  }
}

/* classdef.f90:4481-4498 */
void orc_rotor_shiftwake(orc_rotor_t *r) {
  for (int ib = 0; ib < r->nb; ++ib) {
    orc_blade_t *b = &r->blade[ib];
    for (int i = r->nNwake; i >= 2; --i)
      for (int j = 1; j <= r->ns; ++j) WAN(b, i, j) = WAN(b, i - 1, j);
    for (int f = 0; f < 4; ++f)
      for (int j = 1; j <= r->ns; ++j) WAN(b, 1, j).vf[f].age = 0.0;
  }
}

/* classdef.f90:4500-4513 */
void orc_rotor_shiftFwake(orc_rotor_t *r) {
  for (int ib = 0; ib < r->nb; ++ib) {
    orc_blade_t *b = &r->blade[ib];
    for (int i = r->nFwake; i >= 2; --i) WAF(b, i) = WAF(b, i - 1);
    WAF(b, 1).vf.age = 0.0;
  }
}

/* classdef.f90:4515-4605.  NOTE: rowFarNext is a local that is updated to 1 inside
 * the blade loop after shiftFwake() (which itself shifts ALL blades), so later
 * blades do not shift again -- restated as written. */
void orc_rotor_rollup(orc_rotor_t *r) {
  int rowFarNext = r->rowFar - 1;
  for (int ib = 0; ib < r->nb; ++ib) {
    orc_blade_t *b = &r->blade[ib];
    double gamRollup = WAN(b, r->nNwake, r->ns).gam;
    double centroidLE[3] = {0, 0, 0}, centroidTE[3] = {0, 0, 0};
    double radiusRollup = 0.0, gamSum = 0.0;
    const double sgn = copysign(1.0, r->Omega * r->controlPitch[0]); /* sign(1._dp, Omega*controlPitch(1)) */
    for (int ispan = r->rollupStart; ispan <= r->rollupEnd; ++ispan) {
      const orc_vr_t *w = &WAN(b, r->nNwake, ispan);
      for (int k = 0; k < 3; ++k) {
        centroidLE[k] = centroidLE[k] + w->vf[3].fc[0][k] * w->gam;
        centroidTE[k] = centroidTE[k] + w->vf[2].fc[0][k] * w->gam;
      }
      gamSum = gamSum + w->gam;
      if (sgn > ORC_EPS) {
        if (w->gam < gamRollup) gamRollup = w->gam;
      } else {
        if (w->gam > gamRollup) gamRollup = w->gam;
      }
      radiusRollup = radiusRollup + w->vf[2].rVc * w->gam;
    }
    double ageRollup = WAN(b, r->nNwake, r->ns).vf[2].age;
    if (fabs(gamSum) > ORC_EPS) {
      for (int k = 0; k < 3; ++k) {
        centroidLE[k] = centroidLE[k] / gamSum;
        centroidTE[k] = centroidTE[k] / gamSum;
      }
      radiusRollup = radiusRollup / gamSum;
    } else {
      const orc_vr_t *w = &WAN(b, r->nNwake, r->rollupEnd);
      for (int k = 0; k < 3; ++k) {
        centroidLE[k] = w->vf[1].fc[0][k];
        centroidTE[k] = w->vf[2].fc[0][k];
      }
      radiusRollup = w->vf[2].rVc;
    }
    if (r->suppressFwakeSwitch == 1) gamRollup = 0.0;
    if (r->nFwake > 0) {
      if (rowFarNext == 0) {
        orc_rotor_shiftFwake(r);
        rowFarNext = 1;
      }
      orc_fwake_t *f = &WAF(b, rowFarNext);
      for (int k = 0; k < 3; ++k) {
        f->vf.fc[1][k] = centroidLE[k];
        f->vf.fc[0][k] = centroidTE[k];
      }
      f->gam = gamRollup;
      f->vf.age = ageRollup;
      f->vf.rVc0 = radiusRollup;
      f->vf.rVc = radiusRollup;
      orc_vf_calclength(&f->vf, 1);
      if (rowFarNext < r->nFwake)
        for (int k = 0; k < 3; ++k) WAF(b, rowFarNext + 1).vf.fc[1][k] = centroidTE[k];
    }
  }
  orc_rotor_shiftwake(r);
}

/* --------------------------------------------------------------------- AIC */

/* classdef.f90:4151-4179.  Returns orc_inv2's info. */
int orc_rotor_calcAIC(orc_rotor_t *r) {
  const int N = r->nc * r->ns * r->nb;
  double vec[3];
  for (int ib = 1; ib <= r->nb; ++ib)
    for (int is = 1; is <= r->ns; ++is)
      for (int ic = 1; ic <= r->nc; ++ic) {
        const int row = ic + r->nc * (is - 1) + r->ns * r->nc * (ib - 1);
        const orc_wingpanel_t *pr = &WIP(&r->blade[ib - 1], ic, is);
        for (int jb = 1; jb <= r->nb; ++jb)
          for (int j = 1; j <= r->ns; ++j)
            for (int i = 1; i <= r->nc; ++i) {
              const int col = i + r->nc * (j - 1) + r->ns * r->nc * (jb - 1);
              orc_vr_vind(&WIP(&r->blade[jb - 1], i, j).vr, pr->CP, vec);
              r->AIC[(row - 1) + (size_t)N * (col - 1)] =
                  vec[0] * pr->nCap[0] + vec[1] * pr->nCap[1] + vec[2] * pr->nCap[2];
            }
      }
  return orc_inv2(N, r->AIC, r->AIC_inv);
}

/* classdef.f90:4181-4196 */
void orc_rotor_map_gam(orc_rotor_t *r) {
  const int n = r->nc * r->ns;
  for (int ib = 1; ib <= r->nbConvect; ++ib)
    for (int k = 0; k < n; ++k) r->blade[ib - 1].wiP[k].vr.gam = r->gamVec[k + (size_t)n * (ib - 1)];
  if (r->axisymmetrySwitch == 1)
    for (int ib = 2; ib <= r->nb; ++ib)
      for (int k = 0; k < n; ++k) r->blade[ib - 1].wiP[k].vr.gam = r->blade[0].wiP[k].vr.gam;
}

/* ------------------------------------------------------------ flat helpers */

void orc_vind_flat(long n, const orc_vf_t *fil, const double *gam, const unsigned char *skip, long m,
                   const double *P, double *V) {
#ifdef _OPENMP
#pragma omp parallel for schedule(runtime)
#endif
  for (long t = 0; t < m; ++t) {
    double acc[3] = {0, 0, 0}, v[3];
    for (long s = 0; s < n; ++s) {
      if (skip && skip[s] && !(fabs(gam[s]) > ORC_EPS)) continue;
      orc_vf_vind(&fil[s], &P[3 * t], v);
      for (int k = 0; k < 3; ++k) acc[k] = acc[k] + v[k] * gam[s];
    }
    for (int k = 0; k < 3; ++k) V[3 * t + k] = acc[k];
  }
}

/* Same pair formula evaluated and accumulated in long double (x87 80-bit here). */
void orc_vind_flat_ld(long n, const orc_vf_t *fil, const double *gam, const unsigned char *skip, long m,
                      const double *P, double *V, double *Vabs) {
  const long double inv4pi = 0.25L / (atanl(1.0L) * 4.0L);
#ifdef _OPENMP
#pragma omp parallel for schedule(runtime)
#endif
  for (long t = 0; t < m; ++t) {
    long double acc[3] = {0, 0, 0}, aabs[3] = {0, 0, 0};
    for (long s = 0; s < n; ++s) {
      if (skip && skip[s] && !(fabs(gam[s]) > ORC_EPS)) continue;
      const orc_vf_t *f = &fil[s];
      long double r1[3], r2[3], r0[3], c[3];
      for (int k = 0; k < 3; ++k) {
        r1[k] = (long double)P[3 * t + k] - f->fc[0][k];
        r2[k] = (long double)P[3 * t + k] - f->fc[1][k];
        r0[k] = r1[k] - r2[k];
      }
      c[0] = r1[1] * r2[2] - r1[2] * r2[1];
      c[1] = r1[2] * r2[0] - r1[0] * r2[2];
      c[2] = r1[0] * r2[1] - r1[1] * r2[0];
      long double c2 = c[0] * c[0] + c[1] * c[1] + c[2] * c[2];
      if (c2 > (long double)ORC_EPS * ORC_EPS) {
        long double n1 = sqrtl(r1[0] * r1[0] + r1[1] * r1[1] + r1[2] * r1[2]);
        long double n2 = sqrtl(r2[0] * r2[0] + r2[1] * r2[1] + r2[2] * r2[2]);
        long double n0 = sqrtl(r0[0] * r0[0] + r0[1] * r0[1] + r0[2] * r0[2]);
        long double d = 0;
        for (int k = 0; k < 3; ++k)
          d += r0[k] * ((n1 > ORC_EPS ? r1[k] / n1 : 0.0L) - (n2 > ORC_EPS ? r2[k] / n2 : 0.0L));
        long double q = f->rVc * n0;
        long double den = sqrtl(q * q * q * q + c2 * c2);
        for (int k = 0; k < 3; ++k) {
          long double term = ((c[k] * inv4pi) * d) / den * gam[s];
          acc[k] += term;
          aabs[k] += fabsl(term);
        }
      }
    }
    for (int k = 0; k < 3; ++k) {
      V[3 * t + k] = (double)acc[k];
      if (Vabs) Vabs[3 * t + k] = (double)aabs[k];
    }
  }
}

/* ------------------------------------------------------------------- gridgen */

/* src/gridgen.f90:62-145: Cartesian grid, cell centres, velocity induced by the filaments of one
 * filamentsNNNNN.dat (libPostprocess.f90:363-473) at every cell centre, plus the free stream.
 * Quirk C12 restated: the wing loop ASSIGNS (gridgen.f90:122), so only the last wing ring survives.
 * No |gam| > eps rule here (:125-135).  velCentre, gridCentre: (3, nx-1, ny-1, nz-1) column-major. */
void orc_gridgen(int nx, int ny, int nz, const double xyzMin[3], const double xyzMax[3], const double vel[3],
                 long nVrWing, const orc_vr_t *vrWing, long nVrNwake, const orc_vr_t *vrNwake, long nVfNwakeTE,
                 const orc_vf_t *vfNwakeTE, const double *gamNwakeTE, long nVfFwake, const orc_vf_t *vfFwake,
                 const double *gamFwake, double *gridCentre, double *velCentre) {
  double *ax[3];
  const int n[3] = {nx, ny, nz};
  for (int d = 0; d < 3; ++d) { /* linspace, libMath.f90:138-157 */
    ax[d] = (double *)malloc(sizeof(double) * (size_t)n[d]);
    const double dx = (xyzMax[d] - xyzMin[d]) / (n[d] - 1);
    for (int i = 0; i < n[d]; ++i) ax[d][i] = i * dx;
    for (int i = 0; i < n[d]; ++i) ax[d][i] = ax[d][i] + xyzMin[d];
  }
  const long cx = nx - 1, cy = ny - 1, cz = nz - 1;
  /* corner order of gridgen.f90:77-81 as offsets (dx, dy, dz) */
  static const int off[8][3] = {{0, 0, 0}, {1, 0, 0}, {1, 1, 0}, {1, 1, 1}, {0, 1, 0}, {0, 1, 1}, {0, 0, 1}, {1, 0, 1}};
#ifdef _OPENMP
#pragma omp parallel for collapse(3) schedule(runtime)
#endif
  for (long iz = 0; iz < cz; ++iz)
    for (long iy = 0; iy < cy; ++iy)
      for (long ix = 0; ix < cx; ++ix) {
        const long q = ix + cx * (iy + cy * iz);
        const long idx[3] = {ix, iy, iz};
        double P[3], v[3] = {0, 0, 0}, t[3];
        for (int d = 0; d < 3; ++d) {
          double sum = ax[d][idx[d] + off[0][d]];
          for (int k = 1; k < 8; ++k) sum = sum + ax[d][idx[d] + off[k][d]];
          P[d] = sum * 0.125;
          gridCentre[3 * q + d] = P[d];
        }
        for (long f = 0; f < nVrWing; ++f) { /* :121-123 (assignment, not accumulation) */
          orc_vr_vind(&vrWing[f], P, t);
          for (int d = 0; d < 3; ++d) v[d] = t[d] * vrWing[f].gam;
        }
        for (long f = 0; f < nVrNwake; ++f) { /* :125-127 */
          orc_vr_vind(&vrNwake[f], P, t);
          for (int d = 0; d < 3; ++d) v[d] = v[d] + t[d] * vrNwake[f].gam;
        }
        for (long f = 0; f < nVfNwakeTE; ++f) { /* :129-131 */
          orc_vf_vind(&vfNwakeTE[f], P, t);
          for (int d = 0; d < 3; ++d) v[d] = v[d] + t[d] * gamNwakeTE[f];
        }
        for (long f = 0; f < nVfFwake; ++f) { /* :133-135 */
          orc_vf_vind(&vfFwake[f], P, t);
          for (int d = 0; d < 3; ++d) v[d] = v[d] + t[d] * gamFwake[f];
        }
        for (int d = 0; d < 3; ++d) velCentre[3 * q + d] = v[d] + vel[d]; /* :142-144 */
      }
  for (int d = 0; d < 3; ++d) free(ax[d]);
}

/* ------------------------------------------------- allocation (test harness) */

/* Mirrors the allocations of rotor_init (classdef.f90:3057-3123, :3733-3824 for fdScheme 3)
 * with everything zero-initialised (gam = 0 like :3835-3836). */
orc_rotor_t *orc_rotor_new(int nb, int nc, int ns, int nNwake, int nFwake) {
  orc_rotor_t *r = (orc_rotor_t *)calloc(1, sizeof(orc_rotor_t));
  r->nb = nb;
  r->nc = nc;
  r->ns = ns;
  r->nNwake = nNwake;
  r->nFwake = nFwake;
  r->nbConvect = nb;
  r->nNwakeEnd = nNwake;
  r->nFwakeEnd = nFwake;
  r->rowNear = nNwake + 1; /* main.f90:228-230 */
  r->rowFar = nFwake + 1;
  r->surfaceType = 1;
  r->shaftAxis[2] = 1.0;
  const size_t N = (size_t)nc * ns * nb;
  r->AIC = (double *)calloc(N * N, sizeof(double));
  r->AIC_inv = (double *)calloc(N * N, sizeof(double));
  r->gamVec = (double *)calloc(N, sizeof(double));
  r->gamVecPrev = (double *)calloc(N, sizeof(double));
  r->RHS = (double *)calloc(N, sizeof(double));
  r->blade = (orc_blade_t *)calloc((size_t)nb, sizeof(orc_blade_t));
  for (int ib = 0; ib < nb; ++ib) {
    orc_blade_t *b = &r->blade[ib];
    b->nc = nc;
    b->ns = ns;
    b->nNwake = nNwake;
    b->nFwake = nFwake;
    b->wiP = (orc_wingpanel_t *)calloc((size_t)nc * ns, sizeof(orc_wingpanel_t));
    const size_t nw = (size_t)nNwake * ns + 1, nf = (size_t)nFwake + 1;
    b->waN = (orc_vr_t *)calloc(nw, sizeof(orc_vr_t));
    b->waNPredicted = (orc_vr_t *)calloc(nw, sizeof(orc_vr_t));
    b->waF = (orc_fwake_t *)calloc(nf, sizeof(orc_fwake_t));
    b->waFPredicted = (orc_fwake_t *)calloc(nf, sizeof(orc_fwake_t));
    const size_t vn = 3 * (size_t)nNwake * (ns + 1) + 3, vf = 3 * (size_t)nFwake + 3;
    b->velNwake = (double *)calloc(vn, sizeof(double));
    b->velNwake1 = (double *)calloc(vn, sizeof(double));
    b->velNwakePredicted = (double *)calloc(vn, sizeof(double));
    b->velNwakeStep = (double *)calloc(vn, sizeof(double));
    b->velFwake = (double *)calloc(vf, sizeof(double));
    b->velFwake1 = (double *)calloc(vf, sizeof(double));
    b->velFwakePredicted = (double *)calloc(vf, sizeof(double));
    b->velFwakeStep = (double *)calloc(vf, sizeof(double));
    b->velNwake2 = (double *)calloc(vn, sizeof(double));
    b->velNwake3 = (double *)calloc(vn, sizeof(double));
    b->velFwake2 = (double *)calloc(vf, sizeof(double));
    b->velFwake3 = (double *)calloc(vf, sizeof(double));
    /* sectional arrays of the case driver (vlc_case.c; classdef.f90:3091-3122) */
    double **s1[] = {&b->secChord, &b->secArea, &b->secAlpha, &b->secCL, &b->secCLu, &b->secCD, &b->secMflapArm};
    double **s3[] = {&b->secForceInertial, &b->secLift, &b->secDrag, &b->secLiftDir, &b->secDragDir, &b->secLiftUnsteady,
                     &b->secTauCapChord, &b->secTauCapSpan, &b->secNormalVec, &b->secCP, &b->secChordwiseResVel};
    for (size_t k = 0; k < sizeof(s1) / sizeof(s1[0]); ++k) *s1[k] = (double *)calloc((size_t)ns + 1, sizeof(double));
    for (size_t k = 0; k < sizeof(s3) / sizeof(s3[0]); ++k) *s3[k] = (double *)calloc(3 * (size_t)ns + 3, sizeof(double));
  }
  r->streamwiseCoreVec = (double *)calloc((size_t)ns + 2, sizeof(double));
  return r;
}

void orc_rotor_free(orc_rotor_t *r) {
  if (!r) return;
  for (int ib = 0; ib < r->nb; ++ib) {
    orc_blade_t *b = &r->blade[ib];
    free(b->wiP);
    free(b->waN);
    free(b->waNPredicted);
    free(b->waF);
    free(b->waFPredicted);
    free(b->velNwake);
    free(b->velNwake1);
    free(b->velNwakePredicted);
    free(b->velNwakeStep);
    free(b->velFwake);
    free(b->velFwake1);
    free(b->velFwakePredicted);
    free(b->velFwakeStep);
    free(b->velNwake2);
    free(b->velNwake3);
    free(b->velFwake2);
    free(b->velFwake3);
    double *s[] = {b->secChord, b->secArea, b->secAlpha, b->secCL, b->secCLu, b->secCD, b->secMflapArm,
                   b->secForceInertial, b->secLift, b->secDrag, b->secLiftDir, b->secDragDir, b->secLiftUnsteady,
                   b->secTauCapChord, b->secTauCapSpan, b->secNormalVec, b->secCP, b->secChordwiseResVel};
    for (size_t k = 0; k < sizeof(s) / sizeof(s[0]); ++k) free(s[k]);
  }
  free(r->streamwiseCoreVec);
  free(r->blade);
  free(r->AIC);
  free(r->AIC_inv);
  free(r->gamVec);
  free(r->gamVecPrev);
  free(r->RHS);
  free(r);
}

/* raw views for the Python harness (numpy.ctypeslib.as_array) */
double *orc_rotor_wiP(orc_rotor_t *r, int ib) { return (double *)r->blade[ib].wiP; }
double *orc_rotor_waN(orc_rotor_t *r, int ib, int predicted) {
  return (double *)(predicted ? r->blade[ib].waNPredicted : r->blade[ib].waN);
}
double *orc_rotor_waF(orc_rotor_t *r, int ib, int predicted) {
  return (double *)(predicted ? r->blade[ib].waFPredicted : r->blade[ib].waF);
}
void orc_rotor_get_presc(const orc_rotor_t *r, int out[3]) {
  out[0] = r->prescWakeNt;
  out[1] = r->prescWakeAfterTruncNt;
  out[2] = r->prescWakeGenNt;
}
void orc_rotor_set_pfHelix(orc_rotor_t *r, int ib, int predicted, const double in[2]) {
  r->blade[ib].pfHelixPitch[predicted ? 1 : 0] = in[0];
  r->blade[ib].pfHelixRadius[predicted ? 1 : 0] = in[1];
}
void orc_rotor_get_pfHelix(const orc_rotor_t *r, int ib, int predicted, double out[2]) {
  out[0] = r->blade[ib].pfHelixPitch[predicted ? 1 : 0];
  out[1] = r->blade[ib].pfHelixRadius[predicted ? 1 : 0];
}
double *orc_rotor_wapF(orc_rotor_t *r, int ib, int predicted) {
  return (double *)(predicted ? r->blade[ib].wapFPredicted : r->blade[ib].wapF);
}
/* which: 0 velNwake 1 velNwake1 2 velNwakePredicted 3 velNwakeStep 4 velFwake 5 velFwake1 6 velFwakePredicted 7 velFwakeStep
 *        8 velNwake2 9 velNwake3 10 velFwake2 11 velFwake3 */
double *orc_rotor_vel(orc_rotor_t *r, int ib, int which) {
  orc_blade_t *b = &r->blade[ib];
  double *t[12] = {b->velNwake, b->velNwake1, b->velNwakePredicted, b->velNwakeStep,
                   b->velFwake, b->velFwake1, b->velFwakePredicted, b->velFwakeStep,
                   b->velNwake2, b->velNwake3, b->velFwake2, b->velFwake3};
  return t[which];
}
double *orc_rotor_AIC(orc_rotor_t *r, int inverse) { return inverse ? r->AIC_inv : r->AIC; }
double *orc_rotor_vec(orc_rotor_t *r, int which) { return which == 0 ? r->gamVec : (which == 1 ? r->RHS : r->gamVecPrev); }
/* out[0..9] = nb nc ns nNwake nFwake rowNear rowFar nbConvect nNwakeEnd nFwakeEnd */
void orc_rotor_dims(const orc_rotor_t *r, int *out) {
  out[0] = r->nb; out[1] = r->nc; out[2] = r->ns; out[3] = r->nNwake; out[4] = r->nFwake;
  out[5] = r->rowNear; out[6] = r->rowFar; out[7] = r->nbConvect; out[8] = r->nNwakeEnd; out[9] = r->nFwakeEnd;
}
void orc_rotor_gens(const orc_rotor_t *r, unsigned long out[3]) {
  out[0] = r->gen_wing;
  out[1] = r->gen_wake[0];
  out[2] = r->gen_wake[1];
}
/* out[0..17] = nbConvect axisymmetrySwitch ductSwitch suppressFwakeSwitch rollupStart rollupEnd Omega controlPitch(1)
 * omegaSlow apparentViscCoeff decayCoeff initWakeVel shaftAxis(3) hubCoords(3): what the wake mutators read */
void orc_rotor_get_params(const orc_rotor_t *r, double *out) {
  out[0] = r->nbConvect; out[1] = r->axisymmetrySwitch; out[2] = r->ductSwitch; out[3] = r->suppressFwakeSwitch;
  out[4] = r->rollupStart; out[5] = r->rollupEnd; out[6] = r->Omega; out[7] = r->controlPitch[0];
  out[8] = r->omegaSlow; out[9] = r->apparentViscCoeff; out[10] = r->decayCoeff; out[11] = r->initWakeVel;
  for (int k = 0; k < 3; ++k) {
    out[12 + k] = r->shaftAxis[k];
    out[15 + k] = r->hubCoords[k];
  }
}
void orc_rotor_set_rows(orc_rotor_t *r, int rowNear, int rowFar) {
  r->rowNear = rowNear;
  r->rowFar = rowFar;
}
void orc_rotor_set_params(orc_rotor_t *r, int surfaceType, int axisymmetrySwitch, int nbConvect, double Omega,
                          double omegaSlow, const double *shaftAxis, const double *hubCoords, double theta0,
                          double apparentViscCoeff, double decayCoeff, int rollupStart, int rollupEnd) {
  r->surfaceType = surfaceType;
  r->axisymmetrySwitch = axisymmetrySwitch;
  r->nbConvect = nbConvect;
  r->Omega = Omega;
  r->omegaSlow = omegaSlow;
  for (int k = 0; k < 3; ++k) {
    r->shaftAxis[k] = shaftAxis[k];
    r->hubCoords[k] = hubCoords[k];
  }
  r->controlPitch[0] = theta0;
  r->apparentViscCoeff = apparentViscCoeff;
  r->decayCoeff = decayCoeff;
  r->rollupStart = rollupStart;
  r->rollupEnd = rollupEnd;
}

/* batched point evaluations for the harness (targets (3,m)) */
void orc_rotor_vind_points(const orc_rotor_t *r, int what, int predicted, long m, const double *P, double *V) {
#ifdef _OPENMP
#pragma omp parallel for schedule(runtime)
#endif
  for (long t = 0; t < m; ++t) {
    double a[3] = {0, 0, 0}, w[3] = {0, 0, 0};
    if (what == 0 || what == 2) orc_rotor_vind_bywing(r, &P[3 * t], a);
    if (what == 1 || what == 2) orc_rotor_vind_bywake(r, &P[3 * t], predicted, w);
    if (what == 3) orc_rotor_vind_bywing_boundVortices(r, &P[3 * t], a);
    if (what == 4) /* sum over blades of blade%vind_bywing_chordwiseVortices (classdef.f90:1398-1418) */
      for (int ib = 0; ib < r->nb; ++ib) {
        double t3[3];
        orc_blade_vind_bywing_chordwiseVortices(&r->blade[ib], &P[3 * t], t3);
        for (int k = 0; k < 3; ++k) a[k] = a[k] + t3[k];
      }
    for (int k = 0; k < 3; ++k) V[3 * t + k] = a[k] + w[k];
  }
}
