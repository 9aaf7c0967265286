/*
 * vlc_case.c -- CPU ORACLE, case driver (test infrastructure, NOT product code).
 * See vlc_case.h.  Every function cites the reference file:line it restates; statements keep the
 * reference's evaluation order (left-to-right sums, the same intermediate roundings) so that the
 * CT/CL histories can be compared with the reference's golden files digit for digit.
 */
#include "vlc_case.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define WIP(b, i, j) ((b)->wiP[((i)-1) + (size_t)(b)->nc * ((j)-1)])
#define WAN(b, i, j) ((b)->waN[((i)-1) + (size_t)(b)->nNwake * ((j)-1)])
#define SEC3(a, j) (&(a)[3 * ((j)-1)])

static double pi_(void) { return atan(1.0) * 4.0; } /* libMath.f90:9 */

/* ------------------------------------------------------------------ small vector helpers */
static void v_set(double a[3], double x, double y, double z) { a[0] = x; a[1] = y; a[2] = z; }
static void v_copy(double a[3], const double b[3]) { a[0] = b[0]; a[1] = b[1]; a[2] = b[2]; }
static double v_dot(const double a[3], const double b[3]) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static double sgn1(double x) { return copysign(1.0, x); } /* sign(1._dp, x) */
static void matvec(const double T[9], const double x[3], double y[3]) { /* matmul(T, x), T column-major */
  double t[3];
  for (int r = 0; r < 3; ++r) t[r] = T[r] * x[0] + T[r + 3] * x[1] + T[r + 6] * x[2];
  v_copy(y, t);
}
static void rot_about(const double T[9], const double o[3], double x[3]) { /* matmul(T, x-o)+o */
  double d[3] = {x[0] - o[0], x[1] - o[1], x[2] - o[2]}, y[3];
  matvec(T, d, y);
  for (int k = 0; k < 3; ++k) x[k] = y[k] + o[k];
}
/* libMath.f90:264-276 */
static void projVec(const double a[3], const double d[3], double out[3]) {
  double nsq = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
  if (nsq > ORC_EPS) {
    double s = v_dot(a, d);
    for (int k = 0; k < 3; ++k) out[k] = s * d[k] / nsq;
  } else {
    out[0] = out[1] = out[2] = 0.0;
  }
}
/* libMath.f90:278-291 */
static void noProjVec(const double a[3], const double d[3], double out[3]) {
  double nsq = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
  if (nsq > ORC_EPS) {
    double s = v_dot(a, d);
    for (int k = 0; k < 3; ++k) out[k] = a[k] - s * d[k] / nsq;
  } else {
    v_copy(out, a);
  }
}
/* libMath.f90:138-200 */
static void linspace(double a, double b, int n, double *x) {
  double dx = (b - a) / (n - 1);
  for (int i = 0; i < n; ++i) x[i] = i * dx;
  for (int i = 0; i < n; ++i) x[i] = x[i] + a;
}
static void spacing(int kind, double a, double b, int n, double *x) {
  double *th = (double *)malloc(sizeof(double) * (size_t)n);
  switch (kind) {
    case 2: /* cosspace */
      linspace(0.0, pi_(), n, th);
      for (int i = 0; i < n; ++i) x[i] = a + (b - a) * 0.5 * (1.0 - cos(th[i]));
      break;
    case 3: /* halfsinspace */
      linspace(0.0, pi_() * 0.5, n, th);
      for (int i = 0; i < n; ++i) x[i] = a + (b - a) * sin(th[i]);
      break;
    case 4: /* tanspace */
      linspace(-1.2, 1.2, n, th);
      for (int i = 0; i < n; ++i) x[i] = a + (b - a) * tan(th[i]) / tan(1.2);
      break;
    default:
      linspace(a, b, n, x);
  }
  free(th);
}
/* libMath.f90:476-517 */
static double pwl_interp1d(int n, const double *x, const double *y, double q) {
  if (fabs(x[0] - q) < ORC_EPS) return y[0];
  if (fabs(x[n - 1] - q) < ORC_EPS) return y[n - 1];
  int asc = x[0] < x[n - 1];
  int idx = -1;
  for (int i = 0; i < n; ++i) {
    int t = asc ? (x[i] <= q) : (x[i] >= q);
    if (!t) {
      idx = i - 1;
      break;
    }
  }
  if (idx < 0) idx = 0; /* the reference error-stops when out of range; never hit for chordwiseFraction in (0,1) */
  return y[idx] + (y[idx + 1] - y[idx]) / (x[idx + 1] - x[idx]) * (q - x[idx]);
}
double orc_pwl_interp1d(int n, const double *x, const double *y, double q) { return pwl_interp1d(n, x, y, q); } /* KAT access */
/* libMath.f90:577-605 lsq2_scalar.  The reference inverts the 3x3 normal matrix with its native Doolittle
 * `inv` (libMath.f90:293-426); here the same system is solved by Gaussian elimination with partial pivoting
 * (the quantity only feeds the direction of secChordwiseResVel; agreement is to rounding). */
static double lsq2(double xq, int n, const double *xd, const double *yd) {
  double A[3][4];
  double s1 = 0, s2 = 0, s3 = 0, s4 = 0, r1 = 0, r2 = 0, r3 = 0;
  for (int i = 0; i < n; ++i) s1 += xd[i];
  for (int i = 0; i < n; ++i) s2 += xd[i] * xd[i];
  for (int i = 0; i < n; ++i) s3 += xd[i] * xd[i] * xd[i];
  for (int i = 0; i < n; ++i) s4 += xd[i] * xd[i] * xd[i] * xd[i];
  for (int i = 0; i < n; ++i) r1 += yd[i];
  for (int i = 0; i < n; ++i) r2 += yd[i] * xd[i];
  for (int i = 0; i < n; ++i) r3 += yd[i] * (xd[i] * xd[i]);
  A[0][0] = n;  A[0][1] = s1; A[0][2] = s2; A[0][3] = r1;
  A[1][0] = s1; A[1][1] = s2; A[1][2] = s3; A[1][3] = r2;
  A[2][0] = s2; A[2][1] = s3; A[2][2] = s4; A[2][3] = r3;
  for (int c = 0; c < 3; ++c) {
    int p = c;
    for (int r = c + 1; r < 3; ++r)
      if (fabs(A[r][c]) > fabs(A[p][c])) p = r;
    if (p != c)
      for (int k = 0; k < 4; ++k) {
        double t = A[c][k];
        A[c][k] = A[p][k];
        A[p][k] = t;
      }
    for (int r = c + 1; r < 3; ++r) {
      double f = A[r][c] / A[c][c];
      for (int k = c; k < 4; ++k) A[r][k] -= f * A[c][k];
    }
  }
  double co[3];
  for (int r = 2; r >= 0; --r) {
    double s = A[r][3];
    for (int k = r + 1; k < 3; ++k) s -= A[r][k] * co[k];
    co[r] = s / A[r][r];
  }
  return co[0] + co[1] * xq + co[2] * xq * xq;
}
/* libMath.f90:672-693 Tgb, column-major */
static void Tgb(const double pts[3], double T[9]) {
  double cp = cos(pts[0]), sp = sin(pts[0]), ct = cos(pts[1]), st = sin(pts[1]), cs = cos(pts[2]), ss = sin(pts[2]);
#define TT(i, j) T[((i)-1) + 3 * ((j)-1)]
  TT(1, 1) = cs * ct;
  TT(1, 2) = sp * st * cs - ss * cp;
  TT(1, 3) = sp * ss + st * cp * cs;
  TT(2, 1) = ss * ct;
  TT(2, 2) = sp * ss * st + cp * cs;
  TT(2, 3) = ss * st * cp - sp * cs;
  TT(3, 1) = -st;
  TT(3, 2) = sp * ct;
  TT(3, 3) = cp * ct;
#undef TT
}

/* ------------------------------------------------------------------ wingpanel_class methods */
/* classdef.f90:782-797 */
static void wp_calcCP(orc_wingpanel_t *p) {
  for (int k = 0; k < 3; ++k)
    p->CP[k] = ((p->PC[0][k] + p->PC[3][k]) * 0.25 + (p->PC[1][k] + p->PC[2][k]) * 0.75) * 0.5;
}
/* classdef.f90:799-815 */
static void wp_calcN(orc_wingpanel_t *p) {
  double a[3], b[3], c[3];
  for (int k = 0; k < 3; ++k) {
    a[k] = p->PC[2][k] - p->PC[0][k];
    b[k] = p->PC[3][k] - p->PC[1][k];
  }
  orc_cross(a, b, c);
  orc_unitVec(c, p->nCap);
}
/* classdef.f90:823-842 */
static void wp_calcTau(orc_wingpanel_t *p) {
  double a[3], b[3];
  for (int k = 0; k < 3; ++k) {
    a[k] = 0.5 * ((p->PC[1][k] + p->PC[2][k]) - (p->PC[0][k] + p->PC[3][k]));
    b[k] = 0.5 * ((p->PC[2][k] + p->PC[3][k]) - (p->PC[1][k] + p->PC[0][k]));
  }
  orc_unitVec(a, p->tauCapChord);
  orc_unitVec(b, p->tauCapSpan);
}
/* classdef.f90:844-863 (origin optional there; 0 when absent) */
static void wp_rot(orc_wingpanel_t *p, const double T[9], const double origin[3]) {
  for (int i = 0; i < 4; ++i) rot_about(T, origin, p->PC[i]);
  for (int i = 0; i < 4; ++i) { /* vr_rot classdef.f90:626-642 */
    rot_about(T, origin, p->vr.vf[i].fc[0]);
    rot_about(T, origin, p->vr.vf[i].fc[1]);
  }
  rot_about(T, origin, p->CP);
  matvec(T, p->nCap, p->nCap);
  matvec(T, p->tauCapChord, p->tauCapChord);
  matvec(T, p->tauCapSpan, p->tauCapSpan);
}
/* classdef.f90:865-877 */
static void wp_shiftdP(orc_wingpanel_t *p, const double d[3]) {
  for (int k = 0; k < 3; ++k) p->CP[k] = p->CP[k] + d[k];
  for (int i = 1; i <= 4; ++i) {
    for (int k = 0; k < 3; ++k) p->PC[i - 1][k] = p->PC[i - 1][k] + d[k];
    orc_vr_shiftdP(&p->vr, i, d);
  }
}
/* classdef.f90:879-884 */
static void wp_calc_area(orc_wingpanel_t *p) {
  double a[3], b[3], c[3];
  for (int k = 0; k < 3; ++k) {
    a[k] = p->PC[2][k] - p->PC[0][k];
    b[k] = p->PC[3][k] - p->PC[1][k];
  }
  orc_cross(a, b, c);
  p->panelArea = 0.5 * orc_norm2(c);
}
/* classdef.f90:886-893 */
static void wp_calc_mean_dimensions(orc_wingpanel_t *p) {
  double a[3], b[3];
  for (int k = 0; k < 3; ++k) {
    a[k] = p->PC[3][k] - p->PC[0][k];
    b[k] = p->PC[2][k] - p->PC[1][k];
  }
  p->meanSpan = 0.5 * (orc_norm2(a) + orc_norm2(b));
  for (int k = 0; k < 3; ++k) {
    a[k] = p->PC[1][k] - p->PC[0][k];
    b[k] = p->PC[2][k] - p->PC[3][k];
  }
  p->meanChord = 0.5 * (orc_norm2(a) + orc_norm2(b));
}

/* ------------------------------------------------------------------ blade_class methods */
/* classdef.f90:1092-1112 */
static void blade_move(orc_blade_t *b, const double d[3]) {
  for (int j = 1; j <= b->ns; ++j)
    for (int i = 1; i <= b->nc; ++i) wp_shiftdP(&WIP(b, i, j), d);
  for (int j = 1; j <= b->ns; ++j)
    for (int k = 0; k < 3; ++k) SEC3(b->secCP, j)[k] = SEC3(b->secCP, j)[k] + d[k];
  for (int k = 0; k < 3; ++k) b->flapOrigin[k] = b->flapOrigin[k] + d[k];
}

enum { ROT_AZIMUTH, ROT_FLAP, ROT_PITCH };

/* classdef.f90:1194-1281 */
static void blade_rotate(orc_blade_t *b, double angle, const double axis[3], const double origin[3], int type) {
  if (!(fabs(angle) > ORC_EPS)) return;
  const double zero[3] = {0, 0, 0};
  double mo[3] = {-1.0 * origin[0], -1.0 * origin[1], -1.0 * origin[2]};
  double T[9];
  blade_move(b, mo);
  orc_getTransformAxis(angle, axis, T);
  for (int j = 1; j <= b->ns; ++j)
    for (int i = 1; i <= b->nc; ++i) wp_rot(&WIP(b, i, j), T, zero);
  blade_move(b, origin);
  for (int j = 1; j <= b->ns; ++j) {
    rot_about(T, origin, SEC3(b->secCP, j));
    matvec(T, SEC3(b->secTauCapChord, j), SEC3(b->secTauCapChord, j));
    matvec(T, SEC3(b->secTauCapSpan, j), SEC3(b->secTauCapSpan, j));
    matvec(T, SEC3(b->secNormalVec, j), SEC3(b->secNormalVec, j));
  }
  if (type == ROT_AZIMUTH) {
    matvec(T, b->xAxisAzi, b->xAxisAzi);
    matvec(T, b->yAxisAzi, b->yAxisAzi);
    matvec(T, b->zAxisAzi, b->zAxisAzi);
  }
  if (type == ROT_AZIMUTH || type == ROT_FLAP) {
    matvec(T, b->xAxisAziFlap, b->xAxisAziFlap);
    matvec(T, b->yAxisAziFlap, b->yAxisAziFlap);
    matvec(T, b->zAxisAziFlap, b->zAxisAziFlap);
  }
  matvec(T, b->xAxis, b->xAxis);
  matvec(T, b->yAxis, b->yAxis);
  matvec(T, b->zAxis, b->zAxis);
}

/* classdef.f90:1164-1181 */
void orc_blade_rot_pitch(orc_blade_t *b, double theta) {
  if (fabs(theta) > ORC_EPS) {
    double o[3];
    for (int k = 0; k < 3; ++k)
      o[k] = WIP(b, 1, 1).PC[0][k] * (1.0 - b->pivotLE) + WIP(b, b->nc, 1).PC[1][k] * b->pivotLE;
    blade_rotate(b, theta, b->yAxis, o, ROT_PITCH);
  }
}
/* classdef.f90:1183-1192 */
static void blade_rot_flap(orc_blade_t *b, double beta) { blade_rotate(b, beta, b->xAxisAzi, b->flapOrigin, ROT_FLAP); }

/* classdef.f90:1114-1162 (order = 1: Tgb) */
static void blade_rot_pts(orc_blade_t *b, const double T[9], const double origin[3]) {
  for (int j = 1; j <= b->ns; ++j) {
    for (int i = 1; i <= b->nc; ++i) wp_rot(&WIP(b, i, j), T, origin);
    rot_about(T, origin, SEC3(b->secCP, j));
    matvec(T, SEC3(b->secTauCapChord, j), SEC3(b->secTauCapChord, j));
    matvec(T, SEC3(b->secNormalVec, j), SEC3(b->secNormalVec, j));
  }
  double *ax[9] = {b->xAxis, b->yAxis, b->zAxis, b->xAxisAzi, b->yAxisAzi, b->zAxisAzi,
                   b->xAxisAziFlap, b->yAxisAziFlap, b->zAxisAziFlap};
  for (int k = 0; k < 9; ++k) matvec(T, ax[k], ax[k]);
}

/* classdef.f90:2071-2089 */
static void blade_calc_secArea_secChord(orc_blade_t *b) {
  for (int is = 1; is <= b->ns; ++is) {
    double s = 0.0;
    for (int ic = 1; ic <= b->nc; ++ic) s += WIP(b, ic, is).panelArea;
    b->secArea[is - 1] = s;
    double d[3];
    for (int k = 0; k < 3; ++k)
      d[k] = 0.5 * ((WIP(b, 1, is).PC[0][k] + WIP(b, 1, is).PC[3][k]) -
                    (WIP(b, b->nc, is).PC[1][k] + WIP(b, b->nc, is).PC[2][k]));
    b->secChord[is - 1] = orc_norm2(d);
  }
}

/* classdef.f90:2267-2304 */
static void blade_calc_secLocations(orc_blade_t *b, double chordwiseFraction, double flapHingeRadius) {
  const int nc = b->nc;
  double *xz0 = (double *)malloc(sizeof(double) * (size_t)(nc + 1));
  double *xz1 = (double *)malloc(sizeof(double) * (size_t)(nc + 1));
  for (int is = 1; is <= b->ns; ++is) {
    double vecLE[3], vecPC[3] = {0, 0, 0}, pv[3];
    for (int k = 0; k < 3; ++k) vecLE[k] = 0.5 * (WIP(b, 1, is).PC[0][k] + WIP(b, 1, is).PC[3][k]);
    xz0[0] = xz1[0] = 0.0;
    for (int ic = 1; ic <= nc; ++ic) {
      for (int k = 0; k < 3; ++k) vecPC[k] = 0.5 * (WIP(b, ic, is).PC[1][k] + WIP(b, ic, is).PC[2][k]) - vecLE[k];
      projVec(vecPC, SEC3(b->secTauCapChord, is), pv);
      xz0[ic] = orc_norm2(pv);
      xz1[ic] = v_dot(vecPC, SEC3(b->secNormalVec, is));
    }
    double xcp = orc_norm2(vecPC) * chordwiseFraction;
    double zcp = pwl_interp1d(nc + 1, xz0, xz1, xcp);
    for (int k = 0; k < 3; ++k)
      SEC3(b->secCP, is)[k] = vecLE[k] + xcp * SEC3(b->secTauCapChord, is)[k] + zcp * SEC3(b->secNormalVec, is)[k];
    projVec(SEC3(b->secCP, is), b->yAxis, pv);
    b->secMflapArm[is - 1] = orc_norm2(pv) - flapHingeRadius;
  }
  free(xz0);
  free(xz1);
}

/* classdef.f90:2197-2232 (+ wingpanel_calc_chordwiseResVel :917-923) */
static void blade_calc_secChordwiseResVel(orc_blade_t *b) {
  const int nc = b->nc;
  double *xDist = (double *)malloc(sizeof(double) * (size_t)nc);
  double *y = (double *)malloc(sizeof(double) * (size_t)nc);
  for (int is = 1; is <= b->ns; ++is) {
    for (int ic = 1; ic <= nc; ++ic) {
      orc_wingpanel_t *p = &WIP(b, ic, is);
      noProjVec(p->velCPTotal, p->tauCapSpan, p->chordwiseResVel);
      double d[3];
      for (int k = 0; k < 3; ++k) d[k] = p->CP[k] - WIP(b, 1, is).PC[0][k];
      xDist[ic - 1] = v_dot(d, SEC3(b->secTauCapChord, is));
    }
    if (nc >= 3) {
      double d[3];
      for (int k = 0; k < 3; ++k) d[k] = SEC3(b->secCP, is)[k] - WIP(b, 1, is).PC[0][k];
      const double xq = v_dot(d, SEC3(b->secTauCapChord, is));
      for (int i = 0; i < 3; ++i) {
        for (int ic = 1; ic <= nc; ++ic) y[ic - 1] = WIP(b, ic, is).chordwiseResVel[i];
        SEC3(b->secChordwiseResVel, is)[i] = lsq2(xq, nc, xDist, y);
      }
    } else {
      for (int i = 0; i < 3; ++i) {
        double s = 0.0;
        for (int ic = 1; ic <= nc; ++ic) s += WIP(b, ic, is).chordwiseResVel[i];
        SEC3(b->secChordwiseResVel, is)[i] = s / nc;
      }
    }
  }
  free(xDist);
  free(y);
}

/* classdef.f90:2234-2265 (secAlpha only; secPhi/secViz/secVix/secTheta are output-only diagnostics) */
static void blade_calc_secAlpha(orc_blade_t *b) {
  blade_calc_secChordwiseResVel(b);
  for (int is = 1; is <= b->ns; ++is)
    b->secAlpha[is - 1] = atan2(v_dot(SEC3(b->secChordwiseResVel, is), SEC3(b->secNormalVec, is)),
                                v_dot(SEC3(b->secChordwiseResVel, is), SEC3(b->secTauCapChord, is)));
}

/* classdef.f90:2355-2366 */
static void blade_dirLiftDrag(orc_blade_t *b, double Omega) {
  for (int is = 1; is <= b->ns; ++is) {
    double c[3], u[3];
    orc_unitVec(SEC3(b->secChordwiseResVel, is), SEC3(b->secDragDir, is));
    orc_cross(SEC3(b->secDragDir, is), b->yAxisAziFlap, c);
    orc_unitVec(c, u);
    for (int k = 0; k < 3; ++k) SEC3(b->secLiftDir, is)[k] = sgn1(Omega) * u[k];
  }
}

/* classdef.f90:1704-1896.  The dead velInduced sweep (:1733-1735) is not evaluated. */
static void blade_calc_force(orc_blade_t *b, double density, double Omega, double dt) {
  const int nc = b->nc, ns = b->ns;
  const double inv = -1.0 * sgn1(Omega); /* invertGammaSign :1726 */
  v_set(b->forceInertial, 0, 0, 0);
  memset(b->secForceInertial, 0, sizeof(double) * 3 * (size_t)ns);
  memset(b->secLift, 0, sizeof(double) * 3 * (size_t)ns);
  memset(b->secDrag, 0, sizeof(double) * 3 * (size_t)ns);
  memset(b->secLiftUnsteady, 0, sizeof(double) * 3 * (size_t)ns);
  for (int is = 1; is <= ns; ++is) {
    for (int ic = 1; ic <= nc; ++ic) {
      orc_wingpanel_t *p = &WIP(b, ic, is);
      const double velTangentialChord = v_dot(p->velCP, p->tauCapChord); /* :1731 */
      const double velTangentialSpan = v_dot(p->velCP, p->tauCapSpan);   /* :1732 */
      /* :1739-1767 elemental circulations */
      double gamElementChord = (ic == 1) ? p->vr.gam : p->vr.gam - WIP(b, ic - 1, is).vr.gam;
      double gamElementSpan = (is == 1) ? p->vr.gam : p->vr.gam - WIP(b, ic, is - 1).vr.gam;
      gamElementChord = inv * gamElementChord;
      gamElementSpan = inv * gamElementSpan;
      if (ic > 1) /* :1774-1780 */
        p->gamTrapz = inv * 0.5 * (p->vr.gam + WIP(b, ic - 1, is).vr.gam);
      else
        p->gamTrapz = inv * 0.5 * p->vr.gam;
      p->delPUnsteady = density * (p->gamTrapz - p->gamPrev) / dt;                                   /* :1786 */
      p->delP = p->delPUnsteady + density * velTangentialChord * gamElementChord / p->meanChord;     /* :1789 */
      if (b->spanwiseLiftSwitch != 0)
        p->delP = p->delP + density * velTangentialSpan * gamElementSpan / p->meanSpan;              /* :1793 */
      p->gamPrev = p->gamTrapz;
      double pl[3], plu[3];
      for (int k = 0; k < 3; ++k) {
        p->normalForce[k] = p->delP * p->panelArea * p->nCap[k];                 /* :1813 */
        p->normalForceUnsteady[k] = p->delPUnsteady * p->panelArea * p->nCap[k]; /* :1816 */
        SEC3(b->secForceInertial, is)[k] = SEC3(b->secForceInertial, is)[k] + p->normalForce[k];
      }
      projVec(p->normalForce, SEC3(b->secLiftDir, is), pl);
      projVec(p->normalForceUnsteady, SEC3(b->secLiftDir, is), plu);
      for (int k = 0; k < 3; ++k) {
        SEC3(b->secLift, is)[k] = SEC3(b->secLift, is)[k] + pl[k];
        SEC3(b->secLiftUnsteady, is)[k] = SEC3(b->secLiftUnsteady, is)[k] + plu[k];
      }
    }
  }
  /* :1861-1892 sectional coefficients (drag terms are zero in the reference) */
  for (int is = 1; is <= ns; ++is) {
    const double mag = orc_norm2(SEC3(b->secChordwiseResVel, is));
    const double q = 0.5 * density * (mag * mag); /* getSecDynamicPressure :2058-2069 */
    if (fabs(q) > ORC_EPS) {
      const double s = sgn1(v_dot(SEC3(b->secLift, is), b->zAxisAziFlap));
      b->secCL[is - 1] = orc_norm2(SEC3(b->secLift, is)) * s / (q * b->secArea[is - 1]);
      b->secCD[is - 1] = orc_norm2(SEC3(b->secDrag, is)) / (q * b->secArea[is - 1]);
      b->secCLu[is - 1] = orc_norm2(SEC3(b->secLiftUnsteady, is)) * s / (q * b->secArea[is - 1]);
    } else {
      b->secCL[is - 1] = b->secCD[is - 1] = b->secCLu[is - 1] = 0.0;
    }
  }
  /* sumSecToNetForces :2368-2380 */
  v_set(b->lift, 0, 0, 0);
  v_set(b->drag, 0, 0, 0);
  v_set(b->liftUnsteady, 0, 0, 0);
  for (int is = 1; is <= ns; ++is)
    for (int k = 0; k < 3; ++k) {
      b->forceInertial[k] += SEC3(b->secForceInertial, is)[k];
      b->lift[k] += SEC3(b->secLift, is)[k];
      b->drag[k] += SEC3(b->secDrag, is)[k];
      b->liftUnsteady[k] += SEC3(b->secLiftUnsteady, is)[k];
    }
}

/* ------------------------------------------------------------------ rotor_class methods */
/* classdef.f90:4114-4136 (pitchDynamicsSwitch = 0) */
double orc_rotor_gettheta(const orc_rotor_t *r, double psi, int ib) {
  const double bladeOffset = 2.0 * pi_() / r->nb * (ib - 1);
  return r->controlPitch[0] + r->controlPitch[1] * cos(psi + bladeOffset) + r->controlPitch[2] * sin(psi + bladeOffset);
}
/* classdef.f90:4202-4214 */
static void rotor_move(orc_rotor_t *r, const double d[3]) {
  for (int ib = 0; ib < r->nb; ++ib) blade_move(&r->blade[ib], d);
  for (int k = 0; k < 3; ++k) {
    r->hubCoords[k] = r->hubCoords[k] + d[k];
    r->cgCoords[k] = r->cgCoords[k] + d[k];
  }
}
/* classdef.f90:4216-4252 (order = 1) */
static void rotor_rot_pts(orc_rotor_t *r, const double pts[3], const double origin_in[3]) {
  double T[9], origin[3];
  v_copy(origin, origin_in); /* the reference passes this%cgCoords, which is itself rotated at the end */
  Tgb(pts, T);
  for (int ib = 0; ib < r->nb; ++ib) blade_rot_pts(&r->blade[ib], T, origin);
  matvec(T, r->shaftAxis, r->shaftAxis);
  matvec(T, r->xAxisBody, r->xAxisBody);
  matvec(T, r->yAxisBody, r->yAxisBody);
  matvec(T, r->zAxisBody, r->zAxisBody);
  rot_about(T, origin, r->hubCoords);
  rot_about(T, origin, r->cgCoords);
}
/* classdef.f90:4265-4291 */
static void rotor_rot_advance(orc_rotor_t *r, double dpsi, int nopitch) {
  r->psi = r->psi + dpsi;
  for (int ib = 1; ib <= r->nb; ++ib) {
    orc_blade_t *b = &r->blade[ib - 1];
    blade_rotate(b, dpsi, r->shaftAxis, r->hubCoords, ROT_AZIMUTH);
    b->psi = b->psi + dpsi;
    if (!nopitch) {
      const double thetaNext = orc_rotor_gettheta(r, r->psi, ib);
      orc_blade_rot_pitch(b, thetaNext - b->theta);
      b->theta = thetaNext;
    }
  }
}
/* classdef.f90:4938-4952 */
void orc_rotor_dirLiftDrag(orc_rotor_t *r) {
  for (int ib = 0; ib < r->nbConvect; ++ib) {
    blade_calc_secChordwiseResVel(&r->blade[ib]);
    blade_dirLiftDrag(&r->blade[ib], r->Omega);
  }
  if (r->axisymmetrySwitch == 1)
    for (int ib = 1; ib < r->nb; ++ib) {
      memcpy(r->blade[ib].secDragDir, r->blade[0].secDragDir, sizeof(double) * 3 * (size_t)r->ns);
      memcpy(r->blade[ib].secLiftDir, r->blade[0].secLiftDir, sizeof(double) * 3 * (size_t)r->ns);
    }
}
/* classdef.f90:4766-4784 */
void orc_rotor_calc_secAlpha(orc_rotor_t *r) {
  for (int ib = 0; ib < r->nbConvect; ++ib) blade_calc_secAlpha(&r->blade[ib]);
  if (r->axisymmetrySwitch == 1)
    for (int ib = 1; ib < r->nb; ++ib)
      memcpy(r->blade[ib].secAlpha, r->blade[0].secAlpha, sizeof(double) * (size_t)r->ns);
}
/* classdef.f90:4607-4671 + sumBladeToNetForces :4954-4988 */
void orc_rotor_calc_force(orc_rotor_t *r, double density, double dt) {
  orc_rotor_dirLiftDrag(r);
  for (int ib = 0; ib < r->nbConvect; ++ib) blade_calc_force(&r->blade[ib], density, r->Omega, dt);
  orc_rotor_sum_forces(r);
}
void orc_rotor_get_force_params(const orc_rotor_t *r, double out[4]) {
  out[0] = r->Omega;
  out[1] = r->spanwiseLiftSwitch;
  out[2] = r->axisymmetrySwitch;
  out[3] = r->nbConvect;
}
/* what params2file writes about a rotor (libPostprocess.f90:76-128) as far as rotor_init derives it:
 * out[0..7] = radius root_cut chord Omega nonDimforceDenominator nNwake wakeTruncateNt prescWakeNt */
void orc_rotor_get_file_params(const orc_rotor_t *r, double out[8]) {
  out[0] = r->radius;
  out[1] = r->root_cut;
  out[2] = r->chord;
  out[3] = r->Omega;
  out[4] = r->nonDimforceDenominator;
  out[5] = r->nNwake;
  out[6] = r->wakeTruncateNt;
  out[7] = r->prescWakeNt;
}
/* classdef.f90:4623-4671: the copies for an axisymmetric rotor + sumBladeToNetForces :4954-4988 */
void orc_rotor_sum_forces(orc_rotor_t *r) {
  if (r->axisymmetrySwitch == 1) {
    const size_t n3 = sizeof(double) * 3 * (size_t)r->ns, n1 = sizeof(double) * (size_t)r->ns;
    for (int ib = 1; ib < r->nb; ++ib) {
      orc_blade_t *b = &r->blade[ib], *b1 = &r->blade[0];
      for (int q = 0; q < r->nc * r->ns; ++q) {
        b->wiP[q].delP = b1->wiP[q].delP;
        b->wiP[q].delPUnsteady = b1->wiP[q].delPUnsteady;
        b->wiP[q].gamPrev = b1->wiP[q].gamPrev;
        v_copy(b->wiP[q].normalForce, b1->wiP[q].normalForce);
        v_copy(b->wiP[q].normalForceUnsteady, b1->wiP[q].normalForceUnsteady);
      }
      memcpy(b->secForceInertial, b1->secForceInertial, n3);
      memcpy(b->secLift, b1->secLift, n3);
      memcpy(b->secLiftDir, b1->secLiftDir, n3);
      memcpy(b->secLiftUnsteady, b1->secLiftUnsteady, n3);
      memcpy(b->secDrag, b1->secDrag, n3);
      memcpy(b->secCL, b1->secCL, n1);
      memcpy(b->secCD, b1->secCD, n1);
      memcpy(b->secCLu, b1->secCLu, n1);
      v_copy(b->forceInertial, b1->forceInertial);
      v_copy(b->lift, b1->lift);
      v_copy(b->drag, b1->drag);
      v_copy(b->liftUnsteady, b1->liftUnsteady);
    }
    v_copy(r->liftPrev, r->lift);
    for (int k = 0; k < 3; ++k) {
      r->forceInertial[k] = r->nb * r->blade[0].forceInertial[k];
      r->lift[k] = r->nb * r->blade[0].lift[k];
      r->drag[k] = r->nb * r->blade[0].drag[k];
      r->liftUnsteady[k] = r->nb * r->blade[0].liftUnsteady[k];
    }
  } else {
    v_copy(r->liftPrev, r->lift);
    v_set(r->forceInertial, 0, 0, 0);
    v_set(r->lift, 0, 0, 0);
    v_set(r->drag, 0, 0, 0);
    v_set(r->liftUnsteady, 0, 0, 0);
    for (int ib = 0; ib < r->nbConvect; ++ib)
      for (int k = 0; k < 3; ++k) {
        r->forceInertial[k] = r->forceInertial[k] + r->blade[ib].forceInertial[k];
        r->lift[k] = r->lift[k] + r->blade[ib].lift[k];
        r->drag[k] = r->drag[k] + r->blade[ib].drag[k];
        r->liftUnsteady[k] = r->liftUnsteady[k] + r->blade[ib].liftUnsteady[k];
      }
  }
}

/* ------------------------------------------------------------------ geometry input + rotor_init */
typedef struct {
  int surfaceType, nb, propConvention, spanSpacing, chordSpacing, nc, ns, nNwake;
  double *grid; /* PLOT3D (3, nc+1, ns+1) or NULL */
  int grid_n;
  double hubCoords[3], cgCoords[3], fromCoords[3], phiThetaPsi[3];
  double span, rootcut, chord, preconeAngle, Omega, shaftAxis[3];
  double theta0, thetaC, thetaS, thetaTwist;
  int ductSwitch, axisymmetrySwitch, spanwiseLiftSwitch, symmetricTau, forceCalcSwitch;
  double pivotLE, flapHinge, velBody[3], omegaBody[3];
  double apparentViscCoeff, decayCoeff;
  int wakeTruncateNt, prescWakeAfterTruncNt, prescWakeGenNt;
  double spanwiseCore, *streamwiseCoreVec;
  int nStream;
  double rollupStartRadius, rollupEndRadius, initWakeVel, psiStart, skewLimit;
  double dragUnitVec[3], sideUnitVec[3], liftUnitVec[3];
} geom_t;

struct orc_case {
  int nr, iter, inited, rotors_inited;
  double t, pairs;
  orc_config_t cfg;
  geom_t *geom;
  orc_rotor_t **rotor;
  orc_hooks_t hooks;
  void *stage_user; /* `user` of the two wake-stage hooks when they were installed separately (orc_case_set_stage_hooks) */
  void *cp_user;    /* the same for the two collocation-point hooks (orc_case_set_cp_hooks) */
  char err[256];
};

/* classdef.f90:5133-5148 */
static void toChordsRevs(const geom_t *g, int *nsteps, double dt) {
  if (*nsteps < 0) {
    if (fabs(g->Omega) < ORC_EPS)
      *nsteps = (int)ceil(abs(*nsteps) * g->chord / (dt * orc_norm2(g->velBody)));
    else
      *nsteps = (int)ceil(2.0 * pi_() * abs(*nsteps) / (fabs(g->Omega) * dt));
  }
}

/* classdef.f90:2769-3891, lifting surfaces (surfaceType 0/1) */
static orc_rotor_t *rotor_init(orc_case_t *c, geom_t *g) {
  orc_config_t *cfg = &c->cfg;
  const double degToRad = pi_() / 180.0, twoPi = 2.0 * pi_();
  double dt = cfg->dt;
  int nt = cfg->nt;
  /* :2966-2987 dt */
  if (sgn1(dt) < 0.0) {
    if (fabs(g->Omega) < ORC_EPS)
      dt = fabs(dt) * g->chord / orc_norm2(g->velBody);
    else
      dt = twoPi * fabs(dt) / fabs(g->Omega);
  }
  if (fabs(dt) <= ORC_EPS) {
    if (fabs(g->Omega) < ORC_EPS)
      dt = (g->chord / g->nc) / orc_norm2(g->velBody);
    else
      dt = 5.0 * degToRad / fabs(g->Omega);
  }
  /* :2990-3011 */
  if (nt <= 0) {
    if (nt == 0) nt = -10;
    toChordsRevs(g, &nt, dt);
  }
  if (cfg->slowStart != 0) toChordsRevs(g, &cfg->slowStartNt, dt);
  toChordsRevs(g, &g->wakeTruncateNt, dt);
  toChordsRevs(g, &g->prescWakeAfterTruncNt, dt);
  toChordsRevs(g, &g->prescWakeGenNt, dt);
  toChordsRevs(g, &g->nNwake, dt);
  const int prescWakeNt = (g->wakeTruncateNt > 0 && g->prescWakeAfterTruncNt > 0) ? g->wakeTruncateNt + g->prescWakeAfterTruncNt : 0; /* :3013-3017 */
  if (g->wakeTruncateNt > 0 && g->wakeTruncateNt < g->nNwake + 1) g->wakeTruncateNt = g->nNwake + 1; /* :3019-3021 */
  if (g->surfaceType == 0) g->surfaceType = 1;
  if (g->nNwake > 0 && g->nNwake < 2) {
    snprintf(c->err, sizeof c->err, "ERROR: Atleast 2 near wake rows mandatory");
    return NULL;
  }
  cfg->dt = dt;
  cfg->nt = nt;
  /* :3039-3055 */
  int nNwake = g->nNwake < nt ? g->nNwake : nt;
  int nFwake = (g->wakeTruncateNt == 0) ? nt - nNwake : g->wakeTruncateNt - nNwake;
  g->nNwake = nNwake;

  const int nc = g->nc, ns = g->ns, nb = g->nb;
  orc_rotor_t *r = orc_rotor_new(nb, nc, ns, nNwake, nFwake);
  r->surfaceType = g->surfaceType;
  r->axisymmetrySwitch = g->axisymmetrySwitch;
  r->ductSwitch = g->ductSwitch;
  r->nbConvect = (g->axisymmetrySwitch == 1) ? 1 : nb;
  r->propConvention = g->propConvention;
  r->spanSpacing = g->spanSpacing;
  r->chordSpacing = g->chordSpacing;
  r->spanwiseLiftSwitch = g->spanwiseLiftSwitch;
  r->symmetricTau = g->symmetricTau;
  r->forceCalcSwitch = g->forceCalcSwitch;
  r->wakeTruncateNt = g->wakeTruncateNt;
  r->prescWakeNt = prescWakeNt;
  r->prescWakeAfterTruncNt = g->prescWakeAfterTruncNt;
  r->prescWakeGenNt = g->prescWakeGenNt;
  r->radius = g->span;
  r->root_cut = g->rootcut;
  r->chord = g->chord;
  r->Omega = g->Omega;
  r->pivotLE = g->pivotLE;
  r->flapHinge = g->flapHinge;
  r->apparentViscCoeff = g->apparentViscCoeff;
  r->decayCoeff = g->decayCoeff;
  r->rollupStartRadius = g->rollupStartRadius;
  r->rollupEndRadius = g->rollupEndRadius;
  r->initWakeVel = g->initWakeVel;
  r->skewLimit = g->skewLimit;
  v_copy(r->hubCoords, g->hubCoords);
  v_copy(r->cgCoords, g->cgCoords);
  v_copy(r->fromCoords, g->fromCoords);
  v_copy(r->shaftAxis, g->shaftAxis);
  v_copy(r->velBody, g->velBody);
  v_copy(r->omegaBody, g->omegaBody);
  v_copy(r->dragUnitVec, g->dragUnitVec);
  v_copy(r->sideUnitVec, g->sideUnitVec);
  v_copy(r->liftUnitVec, g->liftUnitVec);
  v_set(r->xAxisBody, 1, 0, 0);
  v_set(r->yAxisBody, 0, 1, 0);
  v_set(r->zAxisBody, 0, 0, 1);
  /* :3126-3137 conversions */
  r->controlPitch[0] = g->theta0 * degToRad;
  r->controlPitch[1] = g->thetaC * degToRad;
  r->controlPitch[2] = g->thetaS * degToRad;
  for (int k = 0; k < 3; ++k) r->pts[k] = g->phiThetaPsi[k] * degToRad;
  r->thetaTwist = g->thetaTwist * degToRad;
  r->preconeAngle = g->preconeAngle * degToRad;
  r->psiStart = g->psiStart * degToRad;
  r->spanwiseCore = g->spanwiseCore * g->chord;
  { /* classdef.f90:2722-2726 broadcast of a single-valued streamwiseCoreVec */
    double rest = 0.0;
    for (int j = 1; j < g->nStream; ++j) rest += g->streamwiseCoreVec[j] * g->streamwiseCoreVec[j];
    for (int j = 0; j <= ns; ++j) {
      double v = (g->nStream <= 1 || sqrt(rest) < ORC_EPS) ? g->streamwiseCoreVec[0]
                                                           : (j < g->nStream ? g->streamwiseCoreVec[j] : 0.0);
      r->streamwiseCoreVec[j] = v * g->chord;
    }
  }
  r->rollupStart = (int)ceil(g->rollupStartRadius * ns);
  r->rollupEnd = (int)floor(g->rollupEndRadius * ns);

  /* :3152-3260 panel corner coordinates */
  double *xVec = (double *)malloc(sizeof(double) * (size_t)(nc + 1));
  double *yVec = (double *)malloc(sizeof(double) * (size_t)(ns + 1));
  if (!g->grid) {
    const double c0 = (g->Omega >= 0) ? -g->chord : g->chord;
    spacing(g->chordSpacing, c0, 0.0, nc + 1, xVec);
    spacing(g->spanSpacing, g->rootcut * g->span, g->span, ns + 1, yVec);
  }
  for (int ib = 0; ib < nb; ++ib) {
    orc_blade_t *b = &r->blade[ib];
    for (int j = 1; j <= ns; ++j)
      for (int i = 1; i <= nc; ++i) {
        orc_wingpanel_t *p = &WIP(b, i, j);
        if (!g->grid) {
          v_set(p->PC[0], xVec[i - 1], yVec[j - 1], 0.0);
          v_set(p->PC[1], xVec[i], yVec[j - 1], 0.0);
          v_set(p->PC[2], xVec[i], yVec[j], 0.0);
          v_set(p->PC[3], xVec[i - 1], yVec[j], 0.0);
        } else { /* rotor_plot3dtoblade :3957-4020: grid(:, ic, is) */
#define GRID(ic, is) (&g->grid[3 * (((ic)-1) + (size_t)(nc + 1) * ((is)-1))])
          v_copy(p->PC[0], GRID(i, j));
          v_copy(p->PC[1], GRID(i + 1, j));
          v_copy(p->PC[2], GRID(i + 1, j + 1));
          v_copy(p->PC[3], GRID(i, j + 1));
#undef GRID
        }
      }
  }
  free(xVec);
  free(yVec);

  for (int ib = 0; ib < r->nbConvect; ++ib) { /* :3268-3536 */
    orc_blade_t *b = &r->blade[ib];
    v_set(b->xAxis, 1, 0, 0);
    v_set(b->yAxis, 0, 1, 0);
    v_set(b->zAxis, 0, 0, 1);
    v_set(b->xAxisAzi, 1, 0, 0);
    v_set(b->yAxisAzi, 0, 1, 0);
    v_set(b->zAxisAzi, 0, 0, 1);
    v_set(b->xAxisAziFlap, 1, 0, 0);
    v_set(b->yAxisAziFlap, 0, 1, 0);
    v_set(b->zAxisAziFlap, 0, 0, 1);
    for (int k = 0; k < 3; ++k) b->flapOrigin[k] = b->yAxis[k] * r->radius * r->flapHinge;
    for (int j = 1; j <= ns; ++j) { /* :3294-3309 sec vectors */
      double t[3], a[3], d[3], n[3];
      for (int k = 0; k < 3; ++k) {
        t[k] = (WIP(b, nc, j).PC[2][k] + WIP(b, nc, j).PC[1][k] - (WIP(b, 1, j).PC[3][k] + WIP(b, 1, j).PC[0][k])) * 0.5;
        a[k] = WIP(b, nc, j).PC[1][k] - WIP(b, 1, j).PC[3][k];
        d[k] = WIP(b, nc, j).PC[2][k] - WIP(b, 1, j).PC[0][k];
      }
      orc_unitVec(t, SEC3(b->secTauCapChord, j));
      v_set(SEC3(b->secTauCapSpan, j), 0, 1, 0);
      orc_cross(a, d, n);
      orc_unitVec(n, n);
      for (int k = 0; k < 3; ++k) SEC3(b->secNormalVec, j)[k] = sgn1(r->Omega) * n[k];
    }
    /* :3311-3367 vortex-ring corners at the quarter-panel shift (the last row reads PC of panel `i` = nc
     * after the loop, SURVEY C11, which is the same panel) */
    for (int j = 1; j <= ns; ++j)
      for (int i = 1; i <= nc; ++i) {
        orc_wingpanel_t *p = &WIP(b, i, j);
        double xs[4];
        xs[0] = (p->PC[1][0] - p->PC[0][0]) * 0.25;
        xs[3] = (p->PC[2][0] - p->PC[3][0]) * 0.25;
        if (i < nc) {
          xs[1] = (WIP(b, i + 1, j).PC[1][0] - p->PC[1][0]) * 0.25;
          xs[2] = (WIP(b, i + 1, j).PC[2][0] - p->PC[2][0]) * 0.25;
        } else {
          xs[1] = xs[2] = 0.0;
        }
        for (int n = 1; n <= 4; ++n) {
          double P[3] = {p->PC[n - 1][0] + xs[n - 1], p->PC[n - 1][1], p->PC[n - 1][2]};
          orc_vr_assignP(&p->vr, n, P);
        }
      }
    /* :3377-3386 */
    double dxdymin = 1e300;
    for (int is = 1; is <= ns; ++is)
      for (int ic = 1; ic <= nc; ++ic) {
        orc_wingpanel_t *p = &WIP(b, ic, is);
        double a[3], d[3];
        for (int k = 0; k < 3; ++k) {
          a[k] = p->PC[1][k] - p->PC[0][k];
          d[k] = p->PC[2][k] - p->PC[1][k];
        }
        double dx = fabs(orc_norm2(a)), dy = fabs(orc_norm2(d));
        if (dx < dxdymin) dxdymin = dx;
        if (dy < dxdymin) dxdymin = dy;
      }
    /* :3388-3400 shed the last row (0.05 is a default-real literal in the reference) */
    {
      double velShed;
      if (fabs(r->Omega) > ORC_EPS) {
        double d[3];
        for (int k = 0; k < 3; ++k) d[k] = WIP(b, nc, ns).vr.vf[1].fc[0][k] - r->hubCoords[k];
        double a = (double)0.05f * fabs(r->Omega) * orc_norm2(d), q = 0.125 * r->chord / dt;
        velShed = a < q ? a : q;
      } else {
        double mv[3] = {-1.0 * r->velBody[0], -1.0 * r->velBody[1], -1.0 * r->velBody[2]};
        velShed = 0.3 * orc_norm2(mv);
      }
      double d[3] = {sgn1(r->Omega) * velShed * dt, 0.0, 0.0};
      for (int j = 1; j <= ns; ++j) {
        orc_vr_shiftdP(&WIP(b, nc, j).vr, 2, d);
        orc_vr_shiftdP(&WIP(b, nc, j).vr, 3, d);
      }
    }
    /* :3402-3417 */
    for (int j = 1; j <= ns; ++j)
      for (int i = 1; i <= nc; ++i) {
        orc_wingpanel_t *p = &WIP(b, i, j);
        wp_calcCP(p);
        wp_calcN(p);
        if (sgn1(r->Omega) < 0.0)
          for (int k = 0; k < 3; ++k) p->nCap[k] = -1.0 * p->nCap[k];
        wp_calcTau(p);
        double m[3];
        for (int k = 0; k < 3; ++k) m[k] = (WIP(b, 1, j).PC[0][k] + WIP(b, 1, j).PC[3][k]) * 0.5 - p->CP[k];
        p->rHinge = orc_norm2(m);
        wp_calc_area(p);
        wp_calc_mean_dimensions(p);
      }
    blade_calc_secArea_secChord(b); /* :3446-3447 */
    b->spanwiseLiftSwitch = r->spanwiseLiftSwitch;
    for (int j = 1; j <= ns; ++j) { /* :3455-3464 overrideTauSpan */
      v_copy(SEC3(b->secTauCapSpan, j), b->yAxis);
      for (int i = 1; i <= nc; ++i) v_copy(WIP(b, i, j).tauCapSpan, b->yAxis);
    }
    if (r->symmetricTau == 1) /* :3467-3476 */
      for (int j = 1; j <= ns / 2; ++j) {
        for (int k = 0; k < 3; ++k) SEC3(b->secTauCapSpan, j)[k] = -1.0 * SEC3(b->secTauCapSpan, j)[k];
        for (int i = 1; i <= nc; ++i)
          for (int k = 0; k < 3; ++k) WIP(b, i, j).tauCapSpan[k] = -1.0 * WIP(b, i, j).tauCapSpan[k];
      }
    blade_calc_secLocations(b, 0.5, r->flapHinge * r->radius); /* :3479-3480 */
    b->pivotLE = r->pivotLE;
    /* :3498-3519 wing core radii */
    const double core = r->spanwiseCore < dxdymin * 0.1 ? r->spanwiseCore : dxdymin * 0.1;
    for (int q = 0; q < nc * ns; ++q)
      for (int f = 0; f < 4; ++f) b->wiP[q].vr.vf[f].rVc0 = core;
    for (int j = 1; j <= ns; ++j) WIP(b, nc, j).vr.vf[1].rVc0 = r->spanwiseCore;
    for (int q = 0; q < nc * ns; ++q)
      for (int f = 0; f < 4; ++f) b->wiP[q].vr.vf[f].rVc = b->wiP[q].vr.vf[f].rVc0;
  }
  if (r->axisymmetrySwitch == 1) /* :3539-3630 copy blade 1 to the others */
    for (int ib = 1; ib < nb; ++ib) {
      orc_blade_t *b = &r->blade[ib], *b1 = &r->blade[0];
      const size_t n3 = sizeof(double) * 3 * (size_t)ns, n1 = sizeof(double) * (size_t)ns;
      memcpy(b->xAxis, b1->xAxis, sizeof(double) * 27); /* the nine axis triplets are contiguous */
      v_copy(b->flapOrigin, b1->flapOrigin);
      memcpy(b->secTauCapChord, b1->secTauCapChord, n3);
      memcpy(b->secTauCapSpan, b1->secTauCapSpan, n3);
      memcpy(b->secNormalVec, b1->secNormalVec, n3);
      for (int q = 0; q < nc * ns; ++q) {
        orc_wingpanel_t *p = &b->wiP[q], *p1 = &b1->wiP[q];
        p->vr = p1->vr;
        v_copy(p->CP, p1->CP);
        v_copy(p->nCap, p1->nCap);
        v_copy(p->tauCapChord, p1->tauCapChord);
        v_copy(p->tauCapSpan, p1->tauCapSpan);
        p->rHinge = p1->rHinge;
        p->panelArea = p1->panelArea;
        p->meanChord = p1->meanChord;
        p->meanSpan = p1->meanSpan;
      }
      memcpy(b->secArea, b1->secArea, n1);
      memcpy(b->secChord, b1->secChord, n1);
      b->spanwiseLiftSwitch = b1->spanwiseLiftSwitch;
      memcpy(b->secCP, b1->secCP, n3);
      memcpy(b->secMflapArm, b1->secMflapArm, n1);
      b->pivotLE = b1->pivotLE;
    }
  { /* :3632-3635 */
    double d[3];
    for (int k = 0; k < 3; ++k) d[k] = r->hubCoords[k] - r->fromCoords[k];
    for (int ib = 0; ib < nb; ++ib) blade_move(&r->blade[ib], d);
  }
  for (int ib = 0; ib < nb; ++ib) { /* :3638-3657 (flapInitial = 0: bladeDynamicsSwitch = 0 everywhere) */
    r->blade[ib].preconeAngle = r->preconeAngle;
    blade_rot_flap(&r->blade[ib], r->preconeAngle);
    blade_rot_flap(&r->blade[ib], 0.0);
  }
  for (int ib = 2; ib <= nb; ++ib) { /* :3661-3667 */
    const double bladeOffset = sgn1(r->Omega) * twoPi / nb * (ib - 1);
    blade_rotate(&r->blade[ib - 1], bladeOffset, r->shaftAxis, r->hubCoords, ROT_AZIMUTH);
  }
  rotor_rot_pts(r, r->pts, r->cgCoords);                  /* :3670 */
  rotor_rot_advance(r, sgn1(r->Omega) * r->psiStart, 1);  /* :3673 */
  /* :3676-3691 */
  if (fabs(r->Omega) > ORC_EPS) {
    if (r->propConvention == 0)
      r->nonDimforceDenominator = cfg->density * (pi_() * (r->radius * r->radius)) * ((r->radius * r->Omega) * (r->radius * r->Omega));
    else
      r->nonDimforceDenominator = cfg->density * ((r->Omega / twoPi) * (r->Omega / twoPi)) * pow(2.0 * r->radius, 4.0);
  } else {
    r->nonDimforceDenominator = 0.5 * cfg->density * (r->radius * (1.0 - r->root_cut) * r->chord) * v_dot(r->velBody, r->velBody);
  }
  /* :3826-3859 wake core radii; gam = 0 (calloc) */
  for (int ib = 0; ib < nb; ++ib) {
    orc_blade_t *b = &r->blade[ib];
    for (int j = 1; j <= ns; ++j)
      for (int i = 1; i <= nNwake; ++i) {
        orc_vr_t *w = &WAN(b, i, j);
        w->vf[1].rVc0 = w->vf[1].rVc = r->spanwiseCore;
        w->vf[3].rVc0 = w->vf[3].rVc = r->spanwiseCore;
        w->vf[0].rVc0 = w->vf[0].rVc = r->streamwiseCoreVec[j - 1];
        w->vf[2].rVc0 = w->vf[2].rVc = r->streamwiseCoreVec[j];
      }
    for (int i = 1; i <= nFwake; ++i) b->waF[i - 1].vf.rVc0 = b->waF[i - 1].vf.rVc = r->streamwiseCoreVec[ns];
  }
  /* :3861-3889 wind frame (output only; kept for completeness of force2file's signLift) */
  if (orc_norm2(r->dragUnitVec) <= ORC_EPS && orc_norm2(r->sideUnitVec) <= ORC_EPS && orc_norm2(r->liftUnitVec) <= ORC_EPS) {
    if (fabs(r->Omega) <= ORC_EPS) {
      double u[3];
      if (fabs(r->velBody[0]) > ORC_EPS) {
        v_set(u, r->velBody[0], 0.0, r->velBody[2]);
        v_set(r->sideUnitVec, 0, 1, 0);
      } else {
        v_set(u, 0.0, r->velBody[1], r->velBody[2]);
        v_set(r->sideUnitVec, 1, 0, 0);
      }
      orc_unitVec(u, u);
      for (int k = 0; k < 3; ++k) r->dragUnitVec[k] = -1.0 * u[k];
      orc_cross(r->dragUnitVec, r->sideUnitVec, r->liftUnitVec);
    } else {
      v_copy(r->liftUnitVec, r->shaftAxis);
    }
  }
  return r;
}

/* ------------------------------------------------------------------ default hooks = the CPU oracle */
static int cpu_vind_points(void *user, int jr, int what, int predicted, long m, const double *P, double *V) {
  orc_case_t *c = (orc_case_t *)user;
  orc_rotor_vind_points(c->rotor[jr], what, predicted, m, P, V);
  return 0;
}
static int cpu_vind_onNwake(void *user, int jr, const double *Nwake, int rows, int cols, int ld, int predicted, double *out) {
  orc_case_t *c = (orc_case_t *)user;
  orc_vind_onNwake_byRotor(c->rotor[jr], (const orc_vr_t *)Nwake, rows, cols, ld, predicted, out);
  return 0;
}
static int cpu_vind_onFwake(void *user, int jr, const double *Fwake, int rows, int predicted, double *out) {
  orc_case_t *c = (orc_case_t *)user;
  orc_vind_onFwake_byRotor(c->rotor[jr], (const orc_fwake_t *)Fwake, rows, predicted, out);
  return 0;
}
static int cpu_calcAIC(void *user, int ir, double *AIC, double *AIC_inv) {
  orc_case_t *c = (orc_case_t *)user;
  (void)AIC;
  (void)AIC_inv;
  return orc_rotor_calcAIC(c->rotor[ir]);
}
static int cpu_solve(void *user, int ir, const double *RHS, double *gamVec) {
  orc_case_t *c = (orc_case_t *)user;
  orc_rotor_t *r = c->rotor[ir];
  const int N = r->nc * r->ns * r->nb;
  orc_matmulAX(N, N, r->AIC_inv, RHS, gamVec);
  return 0;
}

/* ------------------------------------------------------------------ case API */
orc_case_t *orc_case_new(int nr) {
  orc_case_t *c = (orc_case_t *)calloc(1, sizeof(orc_case_t));
  c->nr = nr;
  c->cfg.nr = nr;
  c->geom = (geom_t *)calloc((size_t)nr, sizeof(geom_t));
  c->rotor = (orc_rotor_t **)calloc((size_t)nr, sizeof(orc_rotor_t *));
  for (int ir = 0; ir < nr; ++ir) {
    geom_t *g = &c->geom[ir];
    g->nb = 1;
    g->spanSpacing = 1;
    g->chordSpacing = 1;
    g->shaftAxis[2] = 1.0;
    g->streamwiseCoreVec = (double *)calloc(1, sizeof(double));
    g->nStream = 1;
  }
  orc_case_set_hooks(c, NULL);
  return c;
}

void orc_case_free(orc_case_t *c) {
  if (!c) return;
  for (int ir = 0; ir < c->nr; ++ir) {
    free(c->geom[ir].grid);
    free(c->geom[ir].streamwiseCoreVec);
    orc_rotor_free(c->rotor[ir]);
  }
  free(c->geom);
  free(c->rotor);
  free(c);
}

void orc_case_set_stage_hooks(orc_case_t *c, void *user, int (*prestep)(void *, int), int (*convect)(void *, int)) {
  c->stage_user = user;
  c->hooks.wake_prestep = prestep;
  c->hooks.wake_convect = convect;
}

void orc_case_set_cp_hooks(orc_case_t *c, void *user, int (*rhs_solve)(void *), int (*forces)(void *, int)) {
  c->cp_user = user;
  c->hooks.cp_rhs_solve = rhs_solve;
  c->hooks.cp_forces = forces;
}

void orc_case_set_hooks(orc_case_t *c, const orc_hooks_t *h) {
  c->stage_user = NULL;
  c->cp_user = NULL;
  if (h) {
    c->hooks = *h;
  } else {
    c->hooks.user = c;
    c->hooks.vind_points = cpu_vind_points;
    c->hooks.vind_onNwake = cpu_vind_onNwake;
    c->hooks.vind_onFwake = cpu_vind_onFwake;
    c->hooks.calcAIC = cpu_calcAIC;
    c->hooks.solve = cpu_solve;
    c->hooks.wake_prestep = NULL;
    c->hooks.wake_convect = NULL;
    c->hooks.cp_rhs_solve = NULL;
    c->hooks.cp_forces = NULL;
  }
}

#define KEY(name) (strcmp(key, #name) == 0)
int orc_case_set_config(orc_case_t *c, const char *key, double v) {
  orc_config_t *g = &c->cfg;
  if (KEY(nt)) g->nt = (int)v;
  else if (KEY(dt)) g->dt = v;
  else if (KEY(nr)) { if ((int)v != c->nr) return 2; }
  else if (KEY(density)) g->density = v;
  else if (KEY(velSound)) g->velSound = v;
  else if (KEY(kinematicVisc)) g->kinematicVisc = v;
  else if (KEY(ntSub)) g->ntSub = (int)v;
  else if (KEY(ntSubInit)) g->ntSubInit = (int)v;
  else if (KEY(rotorForcePlot)) g->rotorForcePlot = (int)v;
  else if (KEY(wakeDissipation)) g->wakeDissipation = (int)v;
  else if (KEY(wakeStrain)) g->wakeStrain = (int)v;
  else if (KEY(wakeBurst)) g->wakeBurst = (int)v;
  else if (KEY(wakeSuppress)) g->wakeSuppress = (int)v;
  else if (KEY(slowStart)) g->slowStart = (int)v;
  else if (KEY(slowStartNt)) g->slowStartNt = (int)v;
  else if (KEY(fdScheme)) g->fdScheme = (int)v;
  else if (KEY(initWakeVelNt)) g->initWakeVelNt = (int)v;
  else return 1; /* plot / restart / probe switches: accepted by the caller as "not hot-path" */
  return 0;
}

int orc_case_set_geom(orc_case_t *c, int ir, const char *key, int n, const double *x) {
  if (ir < 0 || ir >= c->nr || n < 1) return 2;
  geom_t *g = &c->geom[ir];
#define V3(dst) do { if (n != 3) return 2; v_copy(g->dst, x); } while (0)
  if (KEY(surfaceType)) g->surfaceType = (int)x[0];
  else if (KEY(nb)) g->nb = (int)x[0];
  else if (KEY(propConvention)) g->propConvention = (int)x[0];
  else if (KEY(spanSpacing)) g->spanSpacing = (int)x[0];
  else if (KEY(chordSpacing)) g->chordSpacing = (int)x[0];
  else if (KEY(nc)) g->nc = (int)x[0];
  else if (KEY(ns)) g->ns = (int)x[0];
  else if (KEY(nNwake)) g->nNwake = (int)x[0];
  else if (KEY(hubCoords)) V3(hubCoords);
  else if (KEY(cgCoords)) V3(cgCoords);
  else if (KEY(fromCoords)) V3(fromCoords);
  else if (KEY(phiThetaPsi)) V3(phiThetaPsi);
  else if (KEY(span)) g->span = x[0];
  else if (KEY(rootcut)) g->rootcut = x[0];
  else if (KEY(chord)) g->chord = x[0];
  else if (KEY(preconeAngle)) g->preconeAngle = x[0];
  else if (KEY(Omega)) g->Omega = x[0];
  else if (KEY(shaftAxis)) V3(shaftAxis);
  else if (KEY(theta0)) g->theta0 = x[0];
  else if (KEY(thetaC)) g->thetaC = x[0];
  else if (KEY(thetaS)) g->thetaS = x[0];
  else if (KEY(thetaTwist)) g->thetaTwist = x[0];
  else if (KEY(ductSwitch)) g->ductSwitch = (int)x[0];
  else if (KEY(axisymmetrySwitch)) g->axisymmetrySwitch = (int)x[0];
  else if (KEY(pivotLE)) g->pivotLE = x[0];
  else if (KEY(flapHinge)) g->flapHinge = x[0];
  else if (KEY(spanwiseLiftSwitch)) g->spanwiseLiftSwitch = (int)x[0];
  else if (KEY(symmetricTau)) g->symmetricTau = (int)x[0];
  else if (KEY(velBody)) V3(velBody);
  else if (KEY(omegaBody)) V3(omegaBody);
  else if (KEY(forceCalcSwitch)) g->forceCalcSwitch = (int)x[0];
  else if (KEY(apparentViscCoeff)) g->apparentViscCoeff = x[0];
  else if (KEY(decayCoeff)) g->decayCoeff = x[0];
  else if (KEY(wakeTruncateNt)) g->wakeTruncateNt = (int)x[0];
  else if (KEY(prescWakeAfterTruncNt)) g->prescWakeAfterTruncNt = (int)x[0];
  else if (KEY(prescWakeGenNt)) g->prescWakeGenNt = (int)x[0];
  else if (KEY(spanwiseCore)) g->spanwiseCore = x[0];
  else if (KEY(streamwiseCoreVec)) {
    free(g->streamwiseCoreVec);
    g->streamwiseCoreVec = (double *)malloc(sizeof(double) * (size_t)n);
    memcpy(g->streamwiseCoreVec, x, sizeof(double) * (size_t)n);
    g->nStream = n;
  } else if (KEY(rollupStartRadius)) g->rollupStartRadius = x[0];
  else if (KEY(rollupEndRadius)) g->rollupEndRadius = x[0];
  else if (KEY(initWakeVel)) g->initWakeVel = x[0];
  else if (KEY(psiStart)) g->psiStart = x[0];
  else if (KEY(skewLimit)) g->skewLimit = x[0];
  else if (KEY(dragUnitVec)) V3(dragUnitVec);
  else if (KEY(sideUnitVec)) V3(sideUnitVec);
  else if (KEY(liftUnitVec)) V3(liftUnitVec);
  else if (KEY(grid)) {
    free(g->grid);
    g->grid = (double *)malloc(sizeof(double) * (size_t)n);
    memcpy(g->grid, x, sizeof(double) * (size_t)n);
    g->grid_n = n;
  } else return 1;
#undef V3
  return 0;
}
#undef KEY

int orc_case_iter(const orc_case_t *c) { return c->iter; }
const orc_config_t *orc_case_config(const orc_case_t *c) { return &c->cfg; }
orc_rotor_t *orc_case_rotor(orc_case_t *c, int ir) { return (ir >= 0 && ir < c->nr) ? c->rotor[ir] : NULL; }
const char *orc_case_error(const orc_case_t *c) { return c->err; }
double orc_case_pairs_last_step(const orc_case_t *c) { return c->pairs; }

/* number of source filaments rotor jr presents to vind_bywing / vind_bywake (reference enumeration) */
static double n_wing_fil(const orc_rotor_t *r) { return (abs(r->surfaceType) == 1) ? 4.0 * r->nc * r->ns * r->nb : 0.0; }
static double n_wake_fil(const orc_rotor_t *r) {
  if (r->nNwake <= 0) return 0.0;
  double n = 4.0 * (r->nNwake - r->rowNear + 1) * r->ns;
  if (r->rowFar <= r->nFwake) n += r->ns + (r->nFwake - r->rowFar + 1) + ORC_NPFWAKE;
  return n * r->nb;
}

/* main.f90:31-58 */
int orc_case_init_rotors(orc_case_t *c) {
  if (c->rotors_inited) return 0;
  for (int ir = 0; ir < c->nr; ++ir) {
    geom_t *g = &c->geom[ir];
    if (g->surfaceType < 0 || g->surfaceType == 2) {
      snprintf(c->err, sizeof c->err, "image / non-lifting surfaces are outside the oracle's scope");
      return 3;
    }
    if (g->grid && g->grid_n != 3 * (g->nc + 1) * (g->ns + 1)) {
      snprintf(c->err, sizeof c->err, "ERROR: Wrong or conflicting data in PLOT3D file");
      return 3;
    }
    c->rotor[ir] = rotor_init(c, g);
    if (!c->rotor[ir]) return 3;
  }
  for (int ir = 0; ir < c->nr; ++ir) { /* :43-58 */
    orc_rotor_t *r = c->rotor[ir];
    for (int ib = 1; ib <= r->nb; ++ib) {
      r->blade[ib - 1].theta = orc_rotor_gettheta(r, r->psiStart, ib);
      orc_blade_rot_pitch(&r->blade[ib - 1], sgn1(r->Omega) * r->blade[ib - 1].theta);
    }
    r->gen_wing++;
    r->gen_wake[0]++;
    r->gen_wake[1]++;
  }
  c->rotors_inited = 1;
  return 0;
}

/* velCP = kinematic velocity (main.f90:132-139 / :536-543) for the CPs of blades 1..nbConvect; P gets the CPs */
static void kinematic_velCP(orc_rotor_t *r, double *P) {
  long q = 0;
  for (int ib = 0; ib < r->nbConvect; ++ib) {
    orc_blade_t *b = &r->blade[ib];
    for (int is = 1; is <= r->ns; ++is)
      for (int ic = 1; ic <= r->nc; ++ic, ++q) {
        orc_wingpanel_t *p = &WIP(b, ic, is);
        double d1[3], d2[3], w[3], c1[3], c2[3];
        for (int k = 0; k < 3; ++k) {
          d1[k] = p->CP[k] - r->cgCoords[k];
          d2[k] = p->CP[k] - r->hubCoords[k];
          w[k] = r->omegaSlow * r->shaftAxis[k];
        }
        orc_cross(r->omegaBody, d1, c1);
        orc_cross(w, d2, c2);
        const double flapTerm = b->secMflapArm[is - 1] * b->dflap; /* scalar, broadcast over xyz as in the reference */
        for (int k = 0; k < 3; ++k) {
          p->velCP[k] = ((-1.0 * r->velBody[k] - c1[k]) - c2[k]) - flapTerm;
          p->velCPm[k] = p->velCP[k];
          P[3 * q + k] = p->CP[k];
        }
      }
  }
}

/* RHS = -(velCP . nCap), solve, map_gam for one rotor pass (main.f90:155-197 / :563-603) */
static void finish_rhs(orc_rotor_t *r) {
  const int npb = r->nc * r->ns;
  const int N = npb * r->nb;
  for (int i = 0; i < N; ++i) r->RHS[i] = 0.0;
  for (int ib = 0; ib < r->nbConvect; ++ib)
    for (int q = 0; q < npb; ++q) {
      orc_wingpanel_t *p = &r->blade[ib].wiP[q];
      r->RHS[q + npb * ib] = v_dot(p->velCP, p->nCap);
    }
  if (r->axisymmetrySwitch == 1)
    for (int ib = 1; ib < r->nb; ++ib)
      for (int q = 0; q < npb; ++q) r->RHS[q + npb * ib] = r->RHS[q];
  for (int i = 0; i < N; ++i) r->RHS[i] = -1.0 * r->RHS[i];
}

static int solve_all(orc_case_t *c) {
  for (int ir = 0; ir < c->nr; ++ir) {
    orc_rotor_t *r = c->rotor[ir];
    const int N = r->nc * r->ns * r->nb;
    memcpy(r->gamVecPrev, r->gamVec, sizeof(double) * (size_t)N);
    int rc = c->hooks.solve(c->hooks.user, ir, r->RHS, r->gamVec);
    if (rc) return rc;
    orc_rotor_map_gam(r);
    r->gen_wing++;
  }
  return 0;
}

/* forces, forceCalcSwitch = 0 (main.f90:244-291 / :630-670) */
static int compute_forces(orc_case_t *c) {
  for (int ir = 0; ir < c->nr; ++ir) {
    orc_rotor_t *r = c->rotor[ir];
    if (r->forceCalcSwitch != 0) {
      snprintf(c->err, sizeof c->err, "forceCalcSwitch %d needs C81 tables: outside the oracle's scope", r->forceCalcSwitch);
      return 3;
    }
    const long m = (long)r->nbConvect * r->ns * r->nc;
    if (c->hooks.cp_forces) { /* :630-663, calc_secAlpha and calc_force down to the blade sums by the hook owner */
      for (int jr = 0; jr < c->nr; ++jr)
        c->pairs += (double)m * (2.0 * c->rotor[jr]->nc * c->rotor[jr]->ns + c->rotor[jr]->ns) * c->rotor[jr]->nb;
      c->pairs += (double)m * n_wing_fil(r);
      int rc = c->hooks.cp_forces(c->cp_user ? c->cp_user : c->hooks.user, ir);
      if (rc) return rc;
      orc_rotor_sum_forces(r);
      continue;
    }
    double *P = (double *)malloc(sizeof(double) * 3 * (size_t)m);
    double *V = (double *)malloc(sizeof(double) * 3 * (size_t)m);
    long q = 0;
    for (int ib = 0; ib < r->nbConvect; ++ib)
      for (int k = 0; k < r->nc * r->ns; ++k, ++q) {
        v_copy(&P[3 * q], r->blade[ib].wiP[k].CP);
        v_copy(r->blade[ib].wiP[k].velCPTotal, r->blade[ib].wiP[k].velCP);
      }
    for (int jr = 0; jr <= c->nr; ++jr) {
      /* jr < nr: minus bound vortices of every rotor; jr == nr: plus the full wing of rotor ir */
      const int src = (jr < c->nr) ? jr : ir, what = (jr < c->nr) ? 3 : 0;
      int rc = c->hooks.vind_points(c->hooks.user, src, what, 0, m, P, V);
      if (rc) {
        free(P);
        free(V);
        return rc;
      }
      c->pairs += (double)m * (what == 3 ? (2.0 * c->rotor[src]->nc * c->rotor[src]->ns + c->rotor[src]->ns) * c->rotor[src]->nb
                                         : n_wing_fil(c->rotor[src]));
      q = 0;
      for (int ib = 0; ib < r->nbConvect; ++ib)
        for (int k = 0; k < r->nc * r->ns; ++k, ++q)
          for (int d = 0; d < 3; ++d) {
            double *t = &r->blade[ib].wiP[k].velCPTotal[d];
            *t = (what == 3) ? *t - V[3 * q + d] : *t + V[3 * q + d];
          }
    }
    free(P);
    free(V);
    if (r->axisymmetrySwitch == 1)
      for (int ib = 1; ib < r->nb; ++ib)
        for (int k = 0; k < r->nc * r->ns; ++k) v_copy(r->blade[ib].wiP[k].velCPTotal, r->blade[0].wiP[k].velCPTotal);
    orc_rotor_calc_secAlpha(r);
    orc_rotor_calc_force(r, c->cfg.density, c->cfg.dt);
  }
  return 0;
}

/* main.f90:65-382 */
int orc_case_init(orc_case_t *c) {
  int rc = orc_case_init_rotors(c);
  if (rc) return rc;
  if (c->inited) return 0;
  for (int ir = 0; ir < c->nr; ++ir) { /* :65-81 */
    orc_rotor_t *r = c->rotor[ir];
    rc = c->hooks.calcAIC(c->hooks.user, ir, r->AIC, r->AIC_inv);
    if (rc) {
      snprintf(c->err, sizeof c->err, "Matrix is numerically singular!");
      return rc;
    }
  }
  for (int ir = 0; ir < c->nr; ++ir) c->rotor[ir]->omegaSlow = (c->cfg.slowStart > 0) ? 0.0 : c->rotor[ir]->Omega; /* :97-105 */
  c->t = 0.0;
  c->iter = 0;
  c->pairs = 0.0;
  for (int i = 0; i <= c->cfg.ntSubInit; ++i) { /* :120-211 */
    for (int ir = 0; ir < c->nr; ++ir) {
      orc_rotor_t *r = c->rotor[ir];
      const long m = (long)r->nbConvect * r->ns * r->nc;
      double *P = (double *)malloc(sizeof(double) * 3 * (size_t)m);
      double *V = (double *)malloc(sizeof(double) * 3 * (size_t)m);
      kinematic_velCP(r, P);
      for (int jr = 0; jr < c->nr; ++jr)
        if (jr != ir) {
          rc = c->hooks.vind_points(c->hooks.user, jr, 0, 0, m, P, V);
          if (rc) return rc;
          long q = 0;
          for (int ib = 0; ib < r->nbConvect; ++ib)
            for (int k = 0; k < r->nc * r->ns; ++k, ++q)
              for (int d = 0; d < 3; ++d) r->blade[ib].wiP[k].velCP[d] = r->blade[ib].wiP[k].velCP[d] + V[3 * q + d];
        }
      free(P);
      free(V);
      finish_rhs(r);
    }
    rc = solve_all(c);
    if (rc) return rc;
    if (c->cfg.ntSubInit != 0) {
      double res = 0.0;
      for (int ir = 0; ir < c->nr; ++ir) {
        orc_rotor_t *r = c->rotor[ir];
        double s = 0.0;
        for (int k = 0; k < r->nc * r->ns * r->nb; ++k) s += (r->gamVec[k] - r->gamVecPrev[k]) * (r->gamVec[k] - r->gamVecPrev[k]);
        if (sqrt(s) > res) res = sqrt(s);
      }
      if (res <= ORC_EPS) break;
    }
  }
  for (int ir = 0; ir < c->nr; ++ir) { /* :227-234 */
    orc_rotor_t *r = c->rotor[ir];
    r->rowFar = r->nFwake + 1;
    r->rowNear = r->nNwake + 1;
    if (r->nNwake > 0) orc_rotor_assignshed(r, "TE");
    r->gen_wake[0]++;
  }
  if (c->cfg.rotorForcePlot != 0) {
    rc = compute_forces(c);
    if (rc) return rc;
  }
  c->inited = 1;
  return 0;
}

/* velNwake(:, rowNear:nNwakeEnd, :) += vind_onNwake_byRotor(...) etc. for all (ir, ib, jr): the two wake sweeps
 * main.f90:814-827 (predicted = 0) and :1057-1069 / :889-901 (predicted = 1, into vel*Predicted). */
static int wake_sweep(orc_case_t *c, int predicted) {
  for (int ir = 0; ir < c->nr; ++ir) {
    orc_rotor_t *r = c->rotor[ir];
    if (r->nNwake <= 0) continue;
    const int rowsN = r->nNwakeEnd - r->rowNear + 1, rowsF = r->nFwakeEnd - r->rowFar + 1;
    double *on = (double *)malloc(sizeof(double) * 3 * (size_t)(rowsN > 0 ? rowsN : 1) * (r->ns + 1));
    double *of = (double *)malloc(sizeof(double) * 3 * (size_t)(rowsF > 0 ? rowsF : 1));
    for (int ib = 0; ib < r->nbConvect; ++ib) {
      orc_blade_t *b = &r->blade[ib];
      double *vN = predicted ? b->velNwakePredicted : b->velNwake;
      double *vF = predicted ? b->velFwakePredicted : b->velFwake;
      const orc_vr_t *waN = predicted ? b->waNPredicted : b->waN;
      const orc_fwake_t *waF = predicted ? b->waFPredicted : b->waF;
      /* zero the active rows (:802-811 zeroes rowNear:nNwake; :1058-1059 rowNear:nNwakeEnd) */
      for (int j = 1; j <= r->ns + 1; ++j)
        for (int i = r->rowNear; i <= r->nNwake; ++i)
          for (int k = 0; k < 3; ++k) vN[3 * ((i - 1) + (size_t)r->nNwake * (j - 1)) + k] = 0.0;
      for (int i = r->rowFar; i <= r->nFwake; ++i)
        for (int k = 0; k < 3; ++k) vF[3 * (i - 1) + k] = 0.0;
      for (int jr = 0; jr < c->nr; ++jr) {
        const orc_rotor_t *s = c->rotor[jr];
        const double nsrc = n_wing_fil(s) + n_wake_fil(s);
        if (rowsN > 0) {
          int rc = c->hooks.vind_onNwake(c->hooks.user, jr, (const double *)&waN[r->rowNear - 1], rowsN, r->ns, r->nNwake,
                                         predicted, on);
          if (rc) return rc;
          c->pairs += (double)rowsN * (r->ns + 1) * nsrc;
          for (int j = 1; j <= r->ns + 1; ++j)
            for (int i = 1; i <= rowsN; ++i)
              for (int k = 0; k < 3; ++k) {
                double *t = &vN[3 * ((r->rowNear + i - 2) + (size_t)r->nNwake * (j - 1)) + k];
                *t = *t + on[3 * ((i - 1) + (size_t)rowsN * (j - 1)) + k];
              }
        }
        if (rowsF > 0) {
          int rc = c->hooks.vind_onFwake(c->hooks.user, jr, (const double *)&waF[r->rowFar - 1], rowsF, predicted, of);
          if (rc) return rc;
          c->pairs += (double)rowsF * nsrc;
          for (int i = 1; i <= rowsF; ++i)
            for (int k = 0; k < 3; ++k) vF[3 * (r->rowFar + i - 2) + k] = vF[3 * (r->rowFar + i - 2) + k] + of[3 * (i - 1) + k];
        }
      }
      if (c->iter < c->cfg.initWakeVelNt) { /* :829-838 / :1070-1081 (sign asymmetry: SURVEY C4) */
        for (int k = 0; k < 3; ++k) {
          const double w = r->initWakeVel * r->shaftAxis[k];
          for (int j = 1; j <= r->ns + 1; ++j)
            for (int i = r->rowNear; i <= r->nNwakeEnd; ++i) {
              double *t = &vN[3 * ((i - 1) + (size_t)r->nNwake * (j - 1)) + k];
              *t = predicted ? *t - w : *t + w;
            }
          for (int i = r->rowFar; i <= r->nFwakeEnd; ++i) vF[3 * (i - 1) + k] = vF[3 * (i - 1) + k] - w;
        }
      }
    }
    free(on);
    free(of);
  }
  return 0;
}

static void copy_wake_to_predicted(orc_rotor_t *r) { /* main.f90:869-872 / :1028-1036 */
  r->gen_wake[1]++; /* the 'P' set is rewritten here and convected right after */
  for (int ib = 0; ib < r->nbConvect; ++ib) {
    orc_blade_t *b = &r->blade[ib];
    for (int j = 1; j <= r->ns; ++j)
      for (int i = r->rowNear; i <= r->nNwakeEnd; ++i) {
        const size_t q = (i - 1) + (size_t)r->nNwake * (j - 1);
        b->waNPredicted[q] = b->waN[q];
      }
    for (int i = r->rowFar; i <= r->nFwakeEnd; ++i) b->waFPredicted[i - 1] = b->waF[i - 1];
  }
}

/* main.f90:400-1452 */
int orc_case_step(orc_case_t *c) {
  if (!c->inited) {
    int rc = orc_case_init(c);
    if (rc) return rc;
  }
  const orc_config_t *cfg = &c->cfg;
  const double dt = cfg->dt;
  c->iter += 1;
  c->t = c->t + dt;
  c->pairs = 0.0;
  const int iter = c->iter;
  for (int ir = 0; ir < c->nr; ++ir) { /* :412-417 */
    orc_rotor_t *r = c->rotor[ir];
    r->rowNear = r->rowNear - 1 > 1 ? r->rowNear - 1 : 1;
    if (iter > r->nNwake) r->rowFar = r->rowFar - 1 > 1 ? r->rowFar - 1 : 1;
  }
  for (int ir = 0; ir < c->nr; ++ir) { /* :428-452 */
    orc_rotor_t *r = c->rotor[ir];
    switch (cfg->slowStart) {
      case 0: r->omegaSlow = r->Omega; break;
      case 1: {
        float a = (float)cfg->slowStartNt, bq = (float)(iter + 1); /* min(real(..), real(..)): default real */
        r->omegaSlow = (double)(a < bq ? a : bq) * r->Omega / cfg->slowStartNt;
      } break;
      case 2: r->omegaSlow = tanh(5.0 * iter / cfg->slowStartNt) * r->Omega; break;
      case 3: r->omegaSlow = (tanh((double)(6.0f * 1.0f) * (double)((float)iter / (float)cfg->slowStartNt) - 3.0) + 1.0) * 0.5 * r->Omega; break;
      default: break;
    }
  }
  for (int ir = 0; ir < c->nr; ++ir) { /* :455-463 */
    orc_rotor_t *r = c->rotor[ir];
    double d[3] = {r->velBody[0] * dt, r->velBody[1] * dt, r->velBody[2] * dt};
    double w[3] = {r->omegaBody[0] * dt, r->omegaBody[1] * dt, r->omegaBody[2] * dt};
    rotor_move(r, d);
    rotor_rot_pts(r, w, r->cgCoords);
    rotor_rot_advance(r, r->omegaSlow * dt, 0);
    r->gen_wing++;
  }
  for (int ir = 0; ir < c->nr; ++ir) c->rotor[ir]->gen_wake[0]++; /* rows moved, shed row attached, aged, dissipated below */
  if (cfg->wakeSuppress == 0 && c->hooks.wake_prestep) { /* device-resident wake: the hook owner does :466-506 */
    int rc = c->hooks.wake_prestep(c->stage_user ? c->stage_user : c->hooks.user, iter);
    if (rc) return rc;
  } else if (cfg->wakeSuppress == 0) { /* :466-506 */
    for (int ir = 0; ir < c->nr; ++ir)
      if (c->rotor[ir]->nNwake > 0) orc_rotor_assignshed(c->rotor[ir], "LE");
    for (int ir = 0; ir < c->nr; ++ir)
      if (c->rotor[ir]->nNwake > 0) orc_rotor_age_wake(c->rotor[ir], dt);
    if (cfg->wakeDissipation == 1)
      for (int ir = 0; ir < c->nr; ++ir)
        if (c->rotor[ir]->nNwake > 0) orc_rotor_dissipate_wake(c->rotor[ir], dt, cfg->kinematicVisc);
    if (cfg->wakeBurst != 0 && iter % cfg->wakeBurst == 0) /* :490-497 */
      for (int ir = 0; ir < c->nr; ++ir)
        if (c->rotor[ir]->nNwake > 0) orc_rotor_burst_wake(c->rotor[ir]);
  }
  /* RHS, :522-615 */
  if (c->hooks.cp_rhs_solve && cfg->ntSub == 0) { /* collocation-point stage by the hook owner (tier 2c of the C ABI) */
    for (int ir = 0; ir < c->nr; ++ir) {
      orc_rotor_t *r = c->rotor[ir];
      const long m = (long)r->nbConvect * r->ns * r->nc;
      double *P = (double *)malloc(sizeof(double) * 3 * (size_t)m);
      kinematic_velCP(r, P); /* :528-547 stays with the driver */
      free(P);
      r->gen_wing++;
      memcpy(r->gamVecPrev, r->gamVec, sizeof(double) * (size_t)(r->nc * r->ns * r->nb));
      for (int jr = 0; jr < c->nr; ++jr)
        c->pairs += (double)m * (n_wake_fil(c->rotor[jr]) + (jr != ir ? n_wing_fil(c->rotor[jr]) : 0.0));
    }
    int rc = c->hooks.cp_rhs_solve(c->cp_user ? c->cp_user : c->hooks.user);
    if (rc) return rc;
    for (int ir = 0; ir < c->nr; ++ir) orc_rotor_map_gam(c->rotor[ir]); /* the hook owner holds this circulation already */
  } else
  for (int i = 0; i <= cfg->ntSub; ++i) {
    for (int ir = 0; ir < c->nr; ++ir) {
      orc_rotor_t *r = c->rotor[ir];
      const long m = (long)r->nbConvect * r->ns * r->nc;
      double *P = (double *)malloc(sizeof(double) * 3 * (size_t)m);
      double *V = (double *)malloc(sizeof(double) * 3 * (size_t)m);
      kinematic_velCP(r, P);
      for (int jr = 0; jr < c->nr; ++jr)
        for (int pass = 0; pass < 2; ++pass) { /* wake of every rotor, then the wing of every other rotor (:551-560) */
          if (pass == 1 && jr == ir) continue;
          int rc = c->hooks.vind_points(c->hooks.user, jr, pass == 0 ? 1 : 0, 0, m, P, V);
          if (rc) {
            free(P);
            free(V);
            return rc;
          }
          c->pairs += (double)m * (pass == 0 ? n_wake_fil(c->rotor[jr]) : n_wing_fil(c->rotor[jr]));
          long q = 0;
          for (int ib = 0; ib < r->nbConvect; ++ib)
            for (int k = 0; k < r->nc * r->ns; ++k, ++q)
              for (int d = 0; d < 3; ++d) r->blade[ib].wiP[k].velCP[d] = r->blade[ib].wiP[k].velCP[d] + V[3 * q + d];
        }
      free(P);
      free(V);
      finish_rhs(r);
    }
    int rc = solve_all(c);
    if (rc) return rc;
    if (cfg->ntSub != 0) {
      double res = 0.0;
      int done = 0;
      for (int ir = 0; ir < c->nr && !done; ++ir) {
        orc_rotor_t *r = c->rotor[ir];
        double s = 0.0;
        for (int k = 0; k < r->nc * r->ns * r->nb; ++k) s += (r->gamVec[k] - r->gamVecPrev[k]) * (r->gamVec[k] - r->gamVecPrev[k]);
        if (sqrt(s) > res) res = sqrt(s);
        if (res <= ORC_EPS) done = 1;
      }
      if (done) break;
    }
  }
  /* forces, :624-722 */
  if (cfg->rotorForcePlot != 0 && iter % cfg->rotorForcePlot == 0) {
    int rc = compute_forces(c);
    if (rc) return rc;
  }
  /* wake convection, :800-1440 */
  if (cfg->wakeSuppress == 0 && c->hooks.wake_convect) { /* device-resident wake: the hook owner does :800-1440 */
    if (cfg->fdScheme < 0 || cfg->fdScheme > 5) {
      snprintf(c->err, sizeof c->err, "fdScheme %d is outside the oracle's scope (0 ... 5)", cfg->fdScheme);
      return 3;
    }
    int rc = c->hooks.wake_convect(c->stage_user ? c->stage_user : c->hooks.user, iter);
    if (rc) return rc;
    const int nsweeps = (cfg->fdScheme == 0 || cfg->fdScheme == 2 || (cfg->fdScheme == 3 && iter == 1) ||
                         (cfg->fdScheme == 4 && iter == 2) || (cfg->fdScheme == 5 && iter <= 3)) ? 1 : 2;
    for (int ir = 0; ir < c->nr; ++ir) { /* the same count wake_sweep() keeps */
      const orc_rotor_t *r = c->rotor[ir];
      if (r->nNwake <= 0) continue;
      const int rowsN = r->nNwakeEnd - r->rowNear + 1, rowsF = r->nFwakeEnd - r->rowFar + 1;
      for (int jr = 0; jr < c->nr; ++jr) {
        const double nsrc = n_wing_fil(c->rotor[jr]) + n_wake_fil(c->rotor[jr]);
        c->pairs += nsweeps * (double)r->nbConvect * ((rowsN > 0 ? (double)rowsN * (r->ns + 1) : 0.0) + (rowsF > 0 ? rowsF : 0)) * nsrc;
      }
    }
  } else if (cfg->wakeSuppress == 0) {
    int rc = wake_sweep(c, 0);
    if (rc) return rc;
    switch (cfg->fdScheme) {
      case 0: /* :846-859 */
        for (int ir = 0; ir < c->nr; ++ir)
          if (c->rotor[ir]->nNwake > 0) orc_rotor_convectwake(c->rotor[ir], iter, dt, 'C');
        break;
      case 1: /* :861-949 */
        for (int ir = 0; ir < c->nr; ++ir)
          if (c->rotor[ir]->nNwake > 0) {
            copy_wake_to_predicted(c->rotor[ir]);
            orc_rotor_convectwake(c->rotor[ir], iter, dt, 'P');
          }
        rc = wake_sweep(c, 1);
        if (rc) return rc;
        for (int ir = 0; ir < c->nr; ++ir) {
          orc_rotor_t *r = c->rotor[ir];
          if (r->nNwake <= 0) continue;
          const int rowsN = r->nNwakeEnd - r->rowNear + 1, rowsF = r->nFwakeEnd - r->rowFar + 1;
          for (int ib = 0; ib < r->nbConvect; ++ib) {
            orc_blade_t *b = &r->blade[ib];
            /* vel_order2_Nwake on the slice (:, rowNear:nNwakeEnd, :) (libCommon.f90:213-235) */
            const size_t cnt = 3 * (size_t)(rowsN > 0 ? rowsN : 1) * (r->ns + 1);
            double *a = (double *)malloc(sizeof(double) * cnt), *p = (double *)malloc(sizeof(double) * cnt),
                   *o = (double *)malloc(sizeof(double) * cnt);
            for (int j = 1; j <= r->ns + 1; ++j)
              for (int i = 1; i <= rowsN; ++i)
                for (int k = 0; k < 3; ++k) {
                  const size_t src = 3 * ((r->rowNear + i - 2) + (size_t)r->nNwake * (j - 1)) + k;
                  a[3 * ((i - 1) + (size_t)rowsN * (j - 1)) + k] = b->velNwake[src];
                  p[3 * ((i - 1) + (size_t)rowsN * (j - 1)) + k] = b->velNwakePredicted[src];
                }
            if (rowsN > 0) orc_vel_order2_Nwake(a, p, rowsN, r->ns + 1, o);
            for (int j = 1; j <= r->ns + 1; ++j)
              for (int i = 1; i <= rowsN; ++i)
                for (int k = 0; k < 3; ++k)
                  b->velNwake[3 * ((r->rowNear + i - 2) + (size_t)r->nNwake * (j - 1)) + k] = o[3 * ((i - 1) + (size_t)rowsN * (j - 1)) + k];
            free(a);
            free(p);
            free(o);
            if (rowsF > 0) {
              double *of = (double *)malloc(sizeof(double) * 3 * (size_t)rowsF);
              orc_vel_order2_Fwake(&b->velFwake[3 * (r->rowFar - 1)], &b->velFwakePredicted[3 * (r->rowFar - 1)], rowsF, of);
              memcpy(&b->velFwake[3 * (r->rowFar - 1)], of, sizeof(double) * 3 * (size_t)rowsF);
              free(of);
            }
          }
          orc_rotor_convectwake(r, iter, dt, 'C');
        }
        break;
      case 2: /* :951-1000 explicit Adams-Bashforth: one sweep per step */
        for (int ir = 0; ir < c->nr; ++ir) {
          orc_rotor_t *r = c->rotor[ir];
          if (r->nNwake <= 0) continue;
          const size_t nn = 3 * (size_t)r->nNwake * (r->ns + 1), nf = 3 * (size_t)r->nFwake;
          if (iter == 1) {
            orc_rotor_convectwake(r, iter, dt, 'C');
            for (int ib = 0; ib < r->nbConvect; ++ib) {
              orc_blade_t *b = &r->blade[ib];
              memcpy(b->velNwake1, b->velNwake, sizeof(double) * nn);
              memcpy(b->velFwake1, b->velFwake, sizeof(double) * nf);
            }
          } else {
            for (int ib = 0; ib < r->nbConvect; ++ib) { /* :975-988 */
              orc_blade_t *b = &r->blade[ib];
              for (size_t q = 0; q < nn; ++q) b->velNwakeStep[q] = 0.5 * (3.0 * b->velNwake[q] - b->velNwake1[q]);
              for (size_t q = 0; q < nf; ++q) b->velFwakeStep[q] = 0.5 * (3.0 * b->velFwake[q] - b->velFwake1[q]);
              memcpy(b->velNwake1, b->velNwakeStep, sizeof(double) * nn);
              memcpy(b->velFwake1, b->velFwakeStep, sizeof(double) * nf);
              memcpy(b->velNwake, b->velNwakeStep, sizeof(double) * nn);
              memcpy(b->velFwake, b->velFwakeStep, sizeof(double) * nf);
            }
            orc_rotor_convectwake(r, iter, dt, 'C');
          }
        }
        break;
      case 3: /* :1002-1115 */
        if (iter == 1) {
          for (int ir = 0; ir < c->nr; ++ir) {
            orc_rotor_t *r = c->rotor[ir];
            if (r->nNwake <= 0) continue;
            orc_rotor_convectwake(r, iter, dt, 'C');
            for (int ib = 0; ib < r->nbConvect; ++ib) {
              orc_blade_t *b = &r->blade[ib];
              memcpy(b->velNwake1, b->velNwake, sizeof(double) * 3 * (size_t)r->nNwake * (r->ns + 1));
              memcpy(b->velFwake1, b->velFwake, sizeof(double) * 3 * (size_t)r->nFwake);
            }
          }
        } else {
          for (int ir = 0; ir < c->nr; ++ir) {
            orc_rotor_t *r = c->rotor[ir];
            if (r->nNwake <= 0) continue;
            copy_wake_to_predicted(r);
            for (int ib = 0; ib < r->nbConvect; ++ib) { /* :1031-1041, whole arrays */
              orc_blade_t *b = &r->blade[ib];
              const size_t nn = 3 * (size_t)r->nNwake * (r->ns + 1), nf = 3 * (size_t)r->nFwake;
              memcpy(b->velNwakeStep, b->velNwake, sizeof(double) * nn);
              for (size_t q = 0; q < nn; ++q) b->velNwake[q] = 0.5 * (3.0 * b->velNwake[q] - b->velNwake1[q]);
              memcpy(b->velFwakeStep, b->velFwake, sizeof(double) * nf);
              for (size_t q = 0; q < nf; ++q) b->velFwake[q] = 0.5 * (3.0 * b->velFwake[q] - b->velFwake1[q]);
            }
            orc_rotor_convectwake(r, iter, dt, 'P');
          }
          rc = wake_sweep(c, 1);
          if (rc) return rc;
          for (int ir = 0; ir < c->nr; ++ir) {
            orc_rotor_t *r = c->rotor[ir];
            if (r->nNwake <= 0) continue;
            for (int ib = 0; ib < r->nbConvect; ++ib) { /* :1094-1099 */
              orc_blade_t *b = &r->blade[ib];
              const size_t nn = 3 * (size_t)r->nNwake * (r->ns + 1), nf = 3 * (size_t)r->nFwake;
              for (size_t q = 0; q < nn; ++q) b->velNwake[q] = (b->velNwakePredicted[q] + b->velNwakeStep[q]) * 0.5;
              for (size_t q = 0; q < nf; ++q) b->velFwake[q] = (b->velFwakePredicted[q] + b->velFwakeStep[q]) * 0.5;
            }
            orc_rotor_convectwake(r, iter, dt, 'C');
            for (int ib = 0; ib < r->nbConvect; ++ib) { /* :1103-1107 */
              orc_blade_t *b = &r->blade[ib];
              memcpy(b->velNwake1, b->velNwakeStep, sizeof(double) * 3 * (size_t)r->nNwake * (r->ns + 1));
              memcpy(b->velFwake1, b->velFwakeStep, sizeof(double) * 3 * (size_t)r->nFwake);
            }
          }
        }
        break;
      case 4: /* :1117-1248 predictor-corrector Adams-Moulton, third order.  The first branch tests `iter == 0`, which
               * never holds inside the time loop (iter = 1..nt): step 1 already takes the multistep branch with zero
               * histories, step 2 only stores velNwake2 -- restated as written */
      case 5: /* :1250-1404 fourth order: steps 1, 2, 3 fill velNwake1, 2, 3 */
      {
        const int order = cfg->fdScheme == 4 ? 3 : 4;
        const int start = (order == 3) ? (iter == 2 ? 2 : 0) : (iter <= 3 ? iter : 0); /* which history this step fills */
        if (start) {
          for (int ir = 0; ir < c->nr; ++ir) {
            orc_rotor_t *r = c->rotor[ir];
            if (r->nNwake <= 0) continue;
            orc_rotor_convectwake(r, iter, dt, 'C');
            orc_rotor_wakevel_copy(r, start == 1 ? ORC_VEL_1 : (start == 2 ? ORC_VEL_2 : ORC_VEL_3), ORC_VEL);
          }
        } else {
          for (int ir = 0; ir < c->nr; ++ir) {
            orc_rotor_t *r = c->rotor[ir];
            if (r->nNwake <= 0) continue;
            copy_wake_to_predicted(r);
            for (int ib = 0; ib < r->nbConvect; ++ib) { /* :1158-1172 / :1307-1325 */
              orc_blade_t *b = &r->blade[ib];
              const size_t nn = 3 * (size_t)r->nNwake * (r->ns + 1), nf = 3 * (size_t)r->nFwake;
              memcpy(b->velNwakeStep, b->velNwake, sizeof(double) * nn);
              memcpy(b->velFwakeStep, b->velFwake, sizeof(double) * nf);
              if (order == 3) {
                for (size_t q = 0; q < nn; ++q)
                  b->velNwake[q] = (23.0 * b->velNwake[q] - 16.0 * b->velNwake2[q] + 5.0 * b->velNwake1[q]) / 12.0;
                for (size_t q = 0; q < nf; ++q)
                  b->velFwake[q] = (23.0 * b->velFwake[q] - 16.0 * b->velFwake2[q] + 5.0 * b->velFwake1[q]) / 12.0;
              } else {
                for (size_t q = 0; q < nn; ++q)
                  b->velNwake[q] = (55.0 * b->velNwake[q] - 59.0 * b->velNwake3[q] + 37.0 * b->velNwake2[q] - 9.0 * b->velNwake1[q]) / 24.0;
                for (size_t q = 0; q < nf; ++q)
                  b->velFwake[q] = (55.0 * b->velFwake[q] - 59.0 * b->velFwake3[q] + 37.0 * b->velFwake2[q] - 9.0 * b->velFwake1[q]) / 24.0;
              }
            }
            orc_rotor_convectwake(r, iter, dt, 'P');
          }
          rc = wake_sweep(c, 1);
          if (rc) return rc;
          for (int ir = 0; ir < c->nr; ++ir) {
            orc_rotor_t *r = c->rotor[ir];
            if (r->nNwake <= 0) continue;
            for (int ib = 0; ib < r->nbConvect; ++ib) { /* :1222-1231 / :1370-1381 */
              orc_blade_t *b = &r->blade[ib];
              const size_t nn = 3 * (size_t)r->nNwake * (r->ns + 1), nf = 3 * (size_t)r->nFwake;
              if (order == 3) {
                for (size_t q = 0; q < nn; ++q)
                  b->velNwake[q] = (5.0 * b->velNwakePredicted[q] + 8.0 * b->velNwakeStep[q] - 1.0 * b->velNwake2[q]) / 12.0;
                for (size_t q = 0; q < nf; ++q)
                  b->velFwake[q] = (5.0 * b->velFwakePredicted[q] + 8.0 * b->velFwakeStep[q] - 1.0 * b->velFwake2[q]) / 12.0;
              } else {
                for (size_t q = 0; q < nn; ++q)
                  b->velNwake[q] = (9.0 * b->velNwakePredicted[q] + 19.0 * b->velNwakeStep[q] - 5.0 * b->velNwake3[q] + 1.0 * b->velNwake2[q]) / 24.0;
                for (size_t q = 0; q < nf; ++q)
                  b->velFwake[q] = (9.0 * b->velFwakePredicted[q] + 19.0 * b->velFwakeStep[q] - 5.0 * b->velFwake3[q] + 1.0 * b->velFwake2[q]) / 24.0;
              }
            }
            orc_rotor_convectwake(r, iter, dt, 'C');
            orc_rotor_wakevel_copy(r, ORC_VEL_1, ORC_VEL_2); /* :1234-1239 / :1384-1391: shift the histories */
            if (order == 3) {
              orc_rotor_wakevel_copy(r, ORC_VEL_2, ORC_VEL_STEP);
            } else {
              orc_rotor_wakevel_copy(r, ORC_VEL_2, ORC_VEL_3);
              orc_rotor_wakevel_copy(r, ORC_VEL_3, ORC_VEL_STEP);
            }
          }
        }
      } break;
      default:
        snprintf(c->err, sizeof c->err, "fdScheme %d is outside the oracle's scope (0 ... 5)", cfg->fdScheme);
        return 3;
    }
    for (int ir = 0; ir < c->nr; ++ir) c->rotor[ir]->gen_wake[0]++; /* convectwake('C'), strain, roll-up, shed below */
    if (cfg->wakeStrain == 1) /* :1409-1416 */
      for (int ir = 0; ir < c->nr; ++ir)
        if (c->rotor[ir]->nNwake > 0) orc_rotor_strain_wake(c->rotor[ir]);
    for (int ir = 0; ir < c->nr; ++ir) { /* :1419-1439 */
      orc_rotor_t *r = c->rotor[ir];
      if (r->nNwake <= 0) continue;
      if (r->rowNear == 1) orc_rotor_rollup(r);
      orc_rotor_assignshed(r, "TE");
    }
  }
  return 0;
}

/* ---- the wake stages of the time loop as separate calls (what a hook owner in resident mode drives; the CPU versions
 * let tests/native/case_gpu_hooks.c run its orchestration against this file's inline statement of main.f90) ---- */
int orc_case_wake_sweep(orc_case_t *c, int predicted) { /* main.f90:800-838 / :889-911, :1057-1081 */
  const double pairs = c->pairs; /* the driver counts the sweeps of a staged step itself */
  const int rc = wake_sweep(c, predicted);
  c->pairs = pairs;
  return rc;
}
void orc_rotor_wake_to_predicted(orc_rotor_t *r) { copy_wake_to_predicted(r); }
/* op: 0 first-step copy (main.f90:1013-1020), 1 AB2 (:1031-1041), 2 AM2 (:1094-1099), 3 history (:1103-1107),
 * 4 vel_order2 on the active slices (:927-940) -- convected blades */
/* the six velocity arrays of a blade by id (near / far): 0 vel, 1 vel1, 2 velPredicted, 3 velStep, 4 vel2, 5 vel3 */
static void vel_arrays(orc_blade_t *b, double *n[6], double *f[6]) {
  double *nn[6] = {b->velNwake, b->velNwake1, b->velNwakePredicted, b->velNwakeStep, b->velNwake2, b->velNwake3};
  double *ff[6] = {b->velFwake, b->velFwake1, b->velFwakePredicted, b->velFwakeStep, b->velFwake2, b->velFwake3};
  memcpy(n, nn, sizeof nn);
  memcpy(f, ff, sizeof ff);
}
/* dst = src, whole arrays of the convected blades */
int orc_rotor_wakevel_copy(orc_rotor_t *r, int dst, int src) {
  if (dst < 0 || dst > 5 || src < 0 || src > 5) return 2;
  const size_t nn = 3 * (size_t)r->nNwake * (r->ns + 1), nf = 3 * (size_t)r->nFwake;
  for (int ib = 0; ib < r->nbConvect; ++ib) {
    double *n[6], *f[6];
    vel_arrays(&r->blade[ib], n, f);
    if (dst != src) {
      memcpy(n[dst], n[src], sizeof(double) * nn);
      memcpy(f[dst], f[src], sizeof(double) * nf);
    }
  }
  return 0;
}
/* dst = (coef[0]*src[0] + coef[1]*src[1] + ...)/divisor, terms added left to right (the multistep formulas of
 * main.f90:1160-1172, :1222-1231, :1309-1325, :1370-1381 with the signs in the coefficients); dst may be one of src */
int orc_rotor_wakevel_lincomb(orc_rotor_t *r, int dst, int nterms, const int *src, const double *coef, double divisor) {
  if (dst < 0 || dst > 5 || nterms < 1 || nterms > 4) return 2;
  for (int k = 0; k < nterms; ++k)
    if (src[k] < 0 || src[k] > 5) return 2;
  const size_t nn = 3 * (size_t)r->nNwake * (r->ns + 1), nf = 3 * (size_t)r->nFwake;
  for (int ib = 0; ib < r->nbConvect; ++ib) {
    double *n[6], *f[6];
    vel_arrays(&r->blade[ib], n, f);
    for (int far = 0; far < 2; ++far) {
      double **a = far ? f : n;
      const size_t cnt = far ? nf : nn;
      for (size_t q = 0; q < cnt; ++q) {
        double acc = coef[0] * a[src[0]][q];
        for (int k = 1; k < nterms; ++k) acc = acc + coef[k] * a[src[k]][q];
        a[dst][q] = acc / divisor;
      }
    }
  }
  return 0;
}

int orc_rotor_wakevel_op(orc_rotor_t *r, int op) {
  const size_t nn = 3 * (size_t)r->nNwake * (r->ns + 1), nf = 3 * (size_t)r->nFwake;
  for (int ib = 0; ib < r->nbConvect; ++ib) {
    orc_blade_t *b = &r->blade[ib];
    switch (op) {
      case 0:
        memcpy(b->velNwake1, b->velNwake, sizeof(double) * nn);
        memcpy(b->velFwake1, b->velFwake, sizeof(double) * nf);
        break;
      case 1:
        memcpy(b->velNwakeStep, b->velNwake, sizeof(double) * nn);
        memcpy(b->velFwakeStep, b->velFwake, sizeof(double) * nf);
        for (size_t q = 0; q < nn; ++q) b->velNwake[q] = 0.5 * (3.0 * b->velNwakeStep[q] - b->velNwake1[q]);
        for (size_t q = 0; q < nf; ++q) b->velFwake[q] = 0.5 * (3.0 * b->velFwakeStep[q] - b->velFwake1[q]);
        break;
      case 2:
        for (size_t q = 0; q < nn; ++q) b->velNwake[q] = (b->velNwakePredicted[q] + b->velNwakeStep[q]) * 0.5;
        for (size_t q = 0; q < nf; ++q) b->velFwake[q] = (b->velFwakePredicted[q] + b->velFwakeStep[q]) * 0.5;
        break;
      case 3:
        memcpy(b->velNwake1, b->velNwakeStep, sizeof(double) * nn);
        memcpy(b->velFwake1, b->velFwakeStep, sizeof(double) * nf);
        break;
      case 5: /* velStep = vel (whole arrays) */
        memcpy(b->velNwakeStep, b->velNwake, sizeof(double) * nn);
        memcpy(b->velFwakeStep, b->velFwake, sizeof(double) * nf);
        break;
      case 4: {
        const int rowsN = r->nNwakeEnd - r->rowNear + 1, rowsF = r->nFwakeEnd - r->rowFar + 1;
        if (rowsN > 0) { /* column by column on the slice (:, rowNear:nNwakeEnd, j): rows are contiguous inside a column */
          double *o = (double *)malloc(sizeof(double) * 3 * (size_t)rowsN);
          for (int j = 1; j <= r->ns + 1; ++j) {
            double *vn = &b->velNwake[3 * ((r->rowNear - 1) + (size_t)r->nNwake * (j - 1))];
            const double *vp = &b->velNwakePredicted[3 * ((r->rowNear - 1) + (size_t)r->nNwake * (j - 1))];
            orc_vel_order2_Nwake(vn, vp, rowsN, 1, o);
            memcpy(vn, o, sizeof(double) * 3 * (size_t)rowsN);
          }
          free(o);
        }
        if (rowsF > 0) {
          double *o = (double *)malloc(sizeof(double) * 3 * (size_t)rowsF);
          orc_vel_order2_Fwake(&b->velFwake[3 * (r->rowFar - 1)], &b->velFwakePredicted[3 * (r->rowFar - 1)], rowsF, o);
          memcpy(&b->velFwake[3 * (r->rowFar - 1)], o, sizeof(double) * 3 * (size_t)rowsF);
          free(o);
        }
      } break;
      default: return 2;
    }
  }
  return 0;
}

/* libPostprocess.f90:814-837 */
void orc_case_force_nondim(orc_case_t *c, int ir, double out[9]) {
  orc_rotor_t *r = c->rotor[ir];
  const double den = r->nonDimforceDenominator;
  const double zero[3] = {0, 0, 0};
  const double signLift = sgn1(v_dot(r->lift, r->zAxisBody));
  out[0] = signLift * orc_norm2(r->lift) / den;
  out[1] = orc_norm2(r->drag) / den;
  out[2] = orc_norm2(r->liftUnsteady) / den;
  out[3] = orc_norm2(zero) / den;
  out[4] = orc_norm2(zero) / den;
  out[5] = orc_norm2(zero) / den;
  out[6] = r->forceInertial[0] / den;
  out[7] = r->forceInertial[1] / den;
  out[8] = r->forceInertial[2] / den;
}

double *orc_blade_sec(orc_rotor_t *r, int ib, const char *name) {
  orc_blade_t *b = &r->blade[ib];
#define S(n) if (strcmp(name, #n) == 0) return b->n
  S(secChord); S(secArea); S(secAlpha); S(secCL); S(secCLu); S(secCD); S(secMflapArm);
  S(secForceInertial); S(secLift); S(secDrag); S(secLiftDir); S(secDragDir); S(secLiftUnsteady);
  S(secTauCapChord); S(secTauCapSpan); S(secNormalVec); S(secCP); S(secChordwiseResVel);
  S(forceInertial); S(lift); S(drag); S(liftUnsteady); S(yAxisAziFlap); S(zAxisAziFlap); S(yAxis);
#undef S
  return NULL;
}
