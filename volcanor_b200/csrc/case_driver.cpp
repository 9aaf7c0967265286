// case_driver.cpp -- PRODUCT-SIDE driver of an unmodified reference case (round-1 review, task 8).
//
// The reference is one program (src/main.f90) around the hot path.  A user without a Fortran toolchain -- this image has
// none -- still needs something that takes a `.case` directory (config.nml, geomNN.nml, PLOT3D grids) and produces the
// reference's force history with the hot path on the GPU.  This file is that program flow, and nothing more:
//   * what rotor%init derives from the namelists (classdef.f90:2769-3891: dt / nt / wake sizes from chords and
//     revolutions, the panel grid, vortex-ring corners, core radii, the pre-shed trailing edge),
//   * the rigid motion of the wing every step (rotor%move / rot_pts / rot_advance, classdef.f90:4202-4291) and the
//     kinematic part of the collocation-point velocity (main.f90:528-547),
//   * the order of the stages of main.f90:400-1452 for fdScheme 0 ... 5, ntSub sub-iterations included,
//   * the sum over blades and the non-dimensional row of force2file (libPostprocess.f90:814-838).
// EVERYTHING on the hot path is a call into the C ABI (include/volcanor_b200.h, tiers 2 / 2b / 2c): AIC + LU, the
// right-hand side at the collocation points, the solve, map_gam, velCPTotal, the sectional loads, every wake mutator,
// both wake sweeps.  There is no CPU implementation of any of them here -- without the CUDA library the driver cannot
// take a step (vcase_attach fails) -- and nothing under oracle/ is used: the oracle (oracle/vlc_case.c) stays the
// checker that tests/test_case_driver.py compares this driver with, next to the reference's own golden files.
//
// Scope = the oracle's: lifting surfaces (surfaceType 0 / 1), geometryFile '0' or PLOT3D, forceCalcSwitch 0,
// slowStart 0-3, dissipation, strain, burst, axisymmetry, far-wake roll-up / truncation, prescribed far wake.
// Records are the reference's derived types as arrays of doubles (classdef.f90:57-220), as everywhere in the ABI.
//
// Built by volcanor_b200/api.py:build_case_driver with g++ -O2 -ffp-contract=off (the reference's statement order,
// no FMA contraction) into volcanor_b200/libvolcanor_case.so, which links libvolcanor_b200.so.
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/volcanor_b200.h"

namespace {

constexpr double kEps = 2.220446049250313e-16;  // libMath.f90:11
inline double pi_() { return std::atan(1.0) * 4.0; }  // libMath.f90:9

// ---- the reference's records (classdef.f90:57-220), bit-compatible with transfer(blade%wiP, buf) ----
struct Vf {
  double fc[2][3];
  double l0, lc, rVc0, rVc, age, ageAzimuthal;
};
struct Vr {
  Vf vf[4];
  double gam, skew;
};
struct WingPanel {
  Vr vr;
  double gamPrev, gamTrapz;
  double PC[4][3];
  double CP[3], nCap[3], tauCapChord[3], tauCapSpan[3];
  double velCP[3], velCPTotal[3], velCPm[3];
  double normalForce[3], normalForceUnsteady[3], chordwiseResVel[3];
  double velPitch, delP, delPUnsteady, delDiConstant, delDiUnsteady;
  double meanChord, meanSpan, panelArea, rHinge, alpha;
};
struct Fwake {
  Vf vf;
  double gam;
};
static_assert(sizeof(Vf) == 8 * VLC_VF_DOUBLES && sizeof(Vr) == 8 * VLC_VR_DOUBLES, "vf / vr record layout");
static_assert(sizeof(WingPanel) == 8 * VLC_WINGPANEL_DOUBLES && sizeof(Fwake) == 8 * VLC_FWAKE_DOUBLES, "record layout");

// ---- small vector helpers (libMath.f90) ----
inline void v_set(double a[3], double x, double y, double z) { a[0] = x, a[1] = y, a[2] = z; }
inline void v_copy(double a[3], const double b[3]) { a[0] = b[0], a[1] = b[1], a[2] = b[2]; }
inline double v_dot(const double a[3], const double b[3]) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
inline double v_norm(const double a[3]) { return std::sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]); }
inline double sgn1(double x) { return std::copysign(1.0, x); }
inline void v_unit(const double a[3], double u[3]) {  // libMath.f90:249-262
  const double n = v_norm(a);
  if (n > kEps) {
    u[0] = a[0] / n, u[1] = a[1] / n, u[2] = a[2] / n;
  } else {
    u[0] = u[1] = u[2] = 0.0;
  }
}
inline void v_cross(const double a[3], const double b[3], double c[3]) {  // libMath.f90:202-212
  const double x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
  c[0] = x, c[1] = y, c[2] = z;
}
inline void matvec(const double T[9], const double x[3], double y[3]) {  // matmul(T, x), T column-major
  double t[3];
  for (int r = 0; r < 3; ++r) t[r] = T[r] * x[0] + T[r + 3] * x[1] + T[r + 6] * x[2];
  v_copy(y, t);
}
inline void rot_about(const double T[9], const double o[3], double x[3]) {  // matmul(T, x - o) + o
  double d[3] = {x[0] - o[0], x[1] - o[1], x[2] - o[2]}, y[3];
  matvec(T, d, y);
  for (int k = 0; k < 3; ++k) x[k] = y[k] + o[k];
}
void projVec(const double a[3], const double d[3], double out[3]) {  // libMath.f90:264-276
  const double nsq = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
  if (nsq > kEps) {
    const double s = v_dot(a, d);
    for (int k = 0; k < 3; ++k) out[k] = s * d[k] / nsq;
  } else {
    out[0] = out[1] = out[2] = 0.0;
  }
}
void transform_axis(double theta, const double axisVec[3], double T[9]) {  // getTransformAxis, libMath.f90:695-726
  const double n = v_norm(axisVec);
  const double ax[3] = {axisVec[0] / n, axisVec[1] / n, axisVec[2] / n};
  const double ct = std::cos(theta), st = std::sin(theta), omct = 1.0 - ct;
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
void Tgb(const double pts[3], double T[9]) {  // libMath.f90:214-236: body -> global for (phi, theta, psi)
  const double cp = std::cos(pts[0]), sp = std::sin(pts[0]), ct = std::cos(pts[1]), st = std::sin(pts[1]);
  const double cs = std::cos(pts[2]), ss = std::sin(pts[2]);
  T[0] = cs * ct, T[3] = sp * st * cs - ss * cp, T[6] = sp * ss + st * cp * cs;
  T[1] = ss * ct, T[4] = sp * ss * st + cp * cs, T[7] = ss * st * cp - sp * cs;
  T[2] = -st, T[5] = sp * ct, T[8] = cp * ct;
}
void linspace(double a, double b, int n, double* x) {  // libMath.f90:138-157
  const double dx = (b - a) / (n - 1);
  for (int i = 0; i < n; ++i) x[i] = i * dx;
  for (int i = 0; i < n; ++i) x[i] = x[i] + a;
}
void spacing(int kind, double a, double b, int n, double* x) {  // linspace / cosspace / halfsinspace / tanspace :138-200
  std::vector<double> th((size_t)n);
  switch (kind) {
    case 2:
      linspace(0.0, pi_(), n, th.data());
      for (int i = 0; i < n; ++i) x[i] = a + (b - a) * 0.5 * (1.0 - std::cos(th[i]));
      break;
    case 3:
      linspace(0.0, pi_() * 0.5, n, th.data());
      for (int i = 0; i < n; ++i) x[i] = a + (b - a) * std::sin(th[i]);
      break;
    case 4:
      linspace(-1.2, 1.2, n, th.data());
      for (int i = 0; i < n; ++i) x[i] = a + (b - a) * std::tan(th[i]) / std::tan(1.2);
      break;
    default: linspace(a, b, n, x);
  }
}
double pwl_interp1d(int n, const double* x, const double* y, double q) {  // libMath.f90:476-517
  if (std::fabs(x[0] - q) < kEps) return y[0];
  if (std::fabs(x[n - 1] - q) < kEps) return y[n - 1];
  const bool asc = x[0] < x[n - 1];
  int idx = -1;
  for (int i = 0; i < n; ++i) {
    const bool t = asc ? (x[i] <= q) : (x[i] >= q);
    if (!t) {
      idx = i - 1;
      break;
    }
  }
  if (idx < 0) idx = 0;
  return y[idx] + (y[idx + 1] - y[idx]) / (x[idx + 1] - x[idx]) * (q - x[idx]);
}

// ---- vr_class / wingpanel_class methods the driver needs (geometry only) ----
void vr_assignP(Vr& r, int n, const double P[3]) {  // classdef.f90:569-592: corner n = fc(:,1) of filament n = fc(:,2) of n-1
  static const int a[5] = {0, 3, 0, 1, 2}, b[5] = {0, 0, 1, 2, 3};
  for (int k = 0; k < 3; ++k) r.vf[a[n]].fc[1][k] = P[k], r.vf[b[n]].fc[0][k] = P[k];
}
void vr_shiftdP(Vr& r, int n, const double d[3]) {  // classdef.f90:594-624
  static const int a[5] = {0, 3, 0, 1, 2}, b[5] = {0, 0, 1, 2, 3};
  for (int k = 0; k < 3; ++k) {
    r.vf[a[n]].fc[1][k] = r.vf[a[n]].fc[1][k] + d[k];
    r.vf[b[n]].fc[0][k] = r.vf[b[n]].fc[0][k] + d[k];
  }
}
void wp_calcCP(WingPanel& p) {  // classdef.f90:782-797
  for (int k = 0; k < 3; ++k) p.CP[k] = ((p.PC[0][k] + p.PC[3][k]) * 0.25 + (p.PC[1][k] + p.PC[2][k]) * 0.75) * 0.5;
}
void wp_calcN(WingPanel& p) {  // classdef.f90:799-815
  double a[3], b[3], c[3];
  for (int k = 0; k < 3; ++k) a[k] = p.PC[2][k] - p.PC[0][k], b[k] = p.PC[3][k] - p.PC[1][k];
  v_cross(a, b, c);
  v_unit(c, p.nCap);
}
void wp_calcTau(WingPanel& p) {  // classdef.f90:823-842
  double a[3], b[3];
  for (int k = 0; k < 3; ++k) {
    a[k] = 0.5 * ((p.PC[1][k] + p.PC[2][k]) - (p.PC[0][k] + p.PC[3][k]));
    b[k] = 0.5 * ((p.PC[2][k] + p.PC[3][k]) - (p.PC[1][k] + p.PC[0][k]));
  }
  v_unit(a, p.tauCapChord);
  v_unit(b, p.tauCapSpan);
}
void wp_rot(WingPanel& p, const double T[9], const double origin[3]) {  // classdef.f90:844-863, vr_rot :626-642
  for (int i = 0; i < 4; ++i) rot_about(T, origin, p.PC[i]);
  for (int i = 0; i < 4; ++i) rot_about(T, origin, p.vr.vf[i].fc[0]), rot_about(T, origin, p.vr.vf[i].fc[1]);
  rot_about(T, origin, p.CP);
  matvec(T, p.nCap, p.nCap);
  matvec(T, p.tauCapChord, p.tauCapChord);
  matvec(T, p.tauCapSpan, p.tauCapSpan);
}
void wp_shiftdP(WingPanel& p, const double d[3]) {  // classdef.f90:865-877
  for (int k = 0; k < 3; ++k) p.CP[k] = p.CP[k] + d[k];
  for (int i = 1; i <= 4; ++i) {
    for (int k = 0; k < 3; ++k) p.PC[i - 1][k] = p.PC[i - 1][k] + d[k];
    vr_shiftdP(p.vr, i, d);
  }
}
void wp_calc_area(WingPanel& p) {  // classdef.f90:879-884
  double a[3], b[3], c[3];
  for (int k = 0; k < 3; ++k) a[k] = p.PC[2][k] - p.PC[0][k], b[k] = p.PC[3][k] - p.PC[1][k];
  v_cross(a, b, c);
  p.panelArea = 0.5 * v_norm(c);
}
void wp_calc_mean_dimensions(WingPanel& p) {  // classdef.f90:886-893
  double a[3], b[3];
  for (int k = 0; k < 3; ++k) a[k] = p.PC[3][k] - p.PC[0][k], b[k] = p.PC[2][k] - p.PC[1][k];
  p.meanSpan = 0.5 * (v_norm(a) + v_norm(b));
  for (int k = 0; k < 3; ++k) a[k] = p.PC[1][k] - p.PC[0][k], b[k] = p.PC[2][k] - p.PC[3][k];
  p.meanChord = 0.5 * (v_norm(a) + v_norm(b));
}

// ---- blade_class / rotor_class: what the driver keeps on the host (classdef.f90:238-467) ----
struct Blade {
  int nc = 0, ns = 0;
  std::vector<WingPanel> wiP;  // wiP(ic, is) at (ic-1) + nc*(is-1)
  double theta = 0, psi = 0, pivotLE = 0, preconeAngle = 0, dflap = 0;
  double flapOrigin[3] = {0, 0, 0};
  double forceInertial[3] = {0, 0, 0}, lift[3] = {0, 0, 0}, drag[3] = {0, 0, 0}, liftUnsteady[3] = {0, 0, 0};
  double axes[9][3];  // xAxis yAxis zAxis | xAxisAzi yAxisAzi zAxisAzi | xAxisAziFlap yAxisAziFlap zAxisAziFlap
  std::vector<double> secChord, secArea, secMflapArm;                           // (ns)
  std::vector<double> secTauCapChord, secTauCapSpan, secNormalVec, secCP;       // (3, ns)
  std::vector<double> loads;                                                    // the block of vlc_rotor_get_loads
  WingPanel& P(int ic, int is) { return wiP[(size_t)(ic - 1) + (size_t)nc * (is - 1)]; }
  double* s3(std::vector<double>& v, int is) { return &v[3 * (size_t)(is - 1)]; }
  double* xAxis() { return axes[0]; }
  double* yAxis() { return axes[1]; }
  double* xAxisAzi() { return axes[3]; }
  double* yAxisAziFlap() { return axes[7]; }
  double* zAxisAziFlap() { return axes[8]; }
};

enum { ROT_AZIMUTH, ROT_FLAP, ROT_PITCH };

void blade_move(Blade& b, const double d[3]) {  // classdef.f90:1092-1112
  for (auto& p : b.wiP) wp_shiftdP(p, d);
  for (int j = 1; j <= b.ns; ++j)
    for (int k = 0; k < 3; ++k) b.s3(b.secCP, j)[k] = b.s3(b.secCP, j)[k] + d[k];
  for (int k = 0; k < 3; ++k) b.flapOrigin[k] = b.flapOrigin[k] + d[k];
}
void blade_rotate(Blade& b, double angle, const double axis[3], const double origin[3], int type) {  // classdef.f90:1194-1281
  if (!(std::fabs(angle) > kEps)) return;
  const double zero[3] = {0, 0, 0};
  const double mo[3] = {-1.0 * origin[0], -1.0 * origin[1], -1.0 * origin[2]};
  double T[9];
  blade_move(b, mo);
  transform_axis(angle, axis, T);
  for (auto& p : b.wiP) wp_rot(p, T, zero);
  blade_move(b, origin);
  for (int j = 1; j <= b.ns; ++j) {
    rot_about(T, origin, b.s3(b.secCP, j));
    matvec(T, b.s3(b.secTauCapChord, j), b.s3(b.secTauCapChord, j));
    matvec(T, b.s3(b.secTauCapSpan, j), b.s3(b.secTauCapSpan, j));
    matvec(T, b.s3(b.secNormalVec, j), b.s3(b.secNormalVec, j));
  }
  if (type == ROT_AZIMUTH)
    for (int a = 3; a < 6; ++a) matvec(T, b.axes[a], b.axes[a]);
  if (type == ROT_AZIMUTH || type == ROT_FLAP)
    for (int a = 6; a < 9; ++a) matvec(T, b.axes[a], b.axes[a]);
  for (int a = 0; a < 3; ++a) matvec(T, b.axes[a], b.axes[a]);
}
void blade_rot_pitch(Blade& b, double theta) {  // classdef.f90:1164-1181
  if (std::fabs(theta) > kEps) {
    double o[3];
    for (int k = 0; k < 3; ++k) o[k] = b.P(1, 1).PC[0][k] * (1.0 - b.pivotLE) + b.P(b.nc, 1).PC[1][k] * b.pivotLE;
    blade_rotate(b, theta, b.yAxis(), o, ROT_PITCH);
  }
}
void blade_rot_flap(Blade& b, double beta) { blade_rotate(b, beta, b.xAxisAzi(), b.flapOrigin, ROT_FLAP); }  // :1183-1192
void blade_rot_pts(Blade& b, const double T[9], const double origin[3]) {  // classdef.f90:1114-1162 (order 1)
  for (int j = 1; j <= b.ns; ++j) {
    for (int i = 1; i <= b.nc; ++i) wp_rot(b.P(i, j), T, origin);
    rot_about(T, origin, b.s3(b.secCP, j));
    matvec(T, b.s3(b.secTauCapChord, j), b.s3(b.secTauCapChord, j));
    matvec(T, b.s3(b.secNormalVec, j), b.s3(b.secNormalVec, j));
  }
  for (int a = 0; a < 9; ++a) matvec(T, b.axes[a], b.axes[a]);
}
void blade_calc_secArea_secChord(Blade& b) {  // classdef.f90:2071-2089
  for (int is = 1; is <= b.ns; ++is) {
    double s = 0.0;
    for (int ic = 1; ic <= b.nc; ++ic) s += b.P(ic, is).panelArea;
    b.secArea[is - 1] = s;
    double d[3];
    for (int k = 0; k < 3; ++k)
      d[k] = 0.5 * ((b.P(1, is).PC[0][k] + b.P(1, is).PC[3][k]) - (b.P(b.nc, is).PC[1][k] + b.P(b.nc, is).PC[2][k]));
    b.secChord[is - 1] = v_norm(d);
  }
}
void blade_calc_secLocations(Blade& b, double chordwiseFraction, double flapHingeRadius) {  // classdef.f90:2267-2304
  const int nc = b.nc;
  std::vector<double> xz0((size_t)nc + 1), xz1((size_t)nc + 1);
  for (int is = 1; is <= b.ns; ++is) {
    double vecLE[3], vecPC[3] = {0, 0, 0}, pv[3];
    for (int k = 0; k < 3; ++k) vecLE[k] = 0.5 * (b.P(1, is).PC[0][k] + b.P(1, is).PC[3][k]);
    xz0[0] = xz1[0] = 0.0;
    for (int ic = 1; ic <= nc; ++ic) {
      for (int k = 0; k < 3; ++k) vecPC[k] = 0.5 * (b.P(ic, is).PC[1][k] + b.P(ic, is).PC[2][k]) - vecLE[k];
      projVec(vecPC, b.s3(b.secTauCapChord, is), pv);
      xz0[ic] = v_norm(pv);
      xz1[ic] = v_dot(vecPC, b.s3(b.secNormalVec, is));
    }
    const double xcp = v_norm(vecPC) * chordwiseFraction;
    const double zcp = pwl_interp1d(nc + 1, xz0.data(), xz1.data(), xcp);
    for (int k = 0; k < 3; ++k)
      b.s3(b.secCP, is)[k] = vecLE[k] + xcp * b.s3(b.secTauCapChord, is)[k] + zcp * b.s3(b.secNormalVec, is)[k];
    projVec(b.s3(b.secCP, is), b.yAxis(), pv);
    b.secMflapArm[is - 1] = v_norm(pv) - flapHingeRadius;
  }
}

struct Geom {  // geomNN.nml (classdef.f90:2541-2767); missing keys are 0
  int surfaceType = 0, nb = 1, propConvention = 0, spanSpacing = 0, chordSpacing = 0, nc = 0, ns = 0, nNwake = 0;
  std::vector<double> grid;  // PLOT3D (3, nc+1, ns+1), empty for geometryFile '0'
  double hubCoords[3] = {0, 0, 0}, cgCoords[3] = {0, 0, 0}, fromCoords[3] = {0, 0, 0}, phiThetaPsi[3] = {0, 0, 0};
  double span = 0, rootcut = 0, chord = 0, preconeAngle = 0, Omega = 0, shaftAxis[3] = {0, 0, 0};
  double theta0 = 0, thetaC = 0, thetaS = 0, thetaTwist = 0;
  int ductSwitch = 0, axisymmetrySwitch = 0, spanwiseLiftSwitch = 0, symmetricTau = 0, forceCalcSwitch = 0;
  double pivotLE = 0, flapHinge = 0, velBody[3] = {0, 0, 0}, omegaBody[3] = {0, 0, 0};
  double apparentViscCoeff = 0, decayCoeff = 0;
  int wakeTruncateNt = 0, prescWakeAfterTruncNt = 0, prescWakeGenNt = 0;
  double spanwiseCore = 0;
  std::vector<double> streamwiseCoreVec = {0.0};
  double rollupStartRadius = 0, rollupEndRadius = 0, initWakeVel = 0, psiStart = 0, skewLimit = 0;
  double dragUnitVec[3] = {0, 0, 0}, sideUnitVec[3] = {0, 0, 0}, liftUnitVec[3] = {0, 0, 0};
};

struct Config {  // config.nml (libCommon.f90:51-108); missing keys are 0
  int nt = 0, nr = 1;
  double dt = 0, density = 0, velSound = 0, kinematicVisc = 0;
  int ntSub = 0, ntSubInit = 0, rotorForcePlot = 0, wakeDissipation = 0, wakeStrain = 0, wakeBurst = 0, wakeSuppress = 0;
  int slowStart = 0, slowStartNt = 0, fdScheme = 0, initWakeVelNt = 0;
};

struct Rotor {
  int nb = 0, nc = 0, ns = 0, nNwake = 0, nFwake = 0, nbConvect = 0, rowNear = 1, rowFar = 1;
  int surfaceType = 1, axisymmetrySwitch = 0, ductSwitch = 0, suppressFwakeSwitch = 0, rollupStart = 1, rollupEnd = 1;
  int prescWakeNt = 0, prescWakeAfterTruncNt = 0, prescWakeGenNt = 0, wakeTruncateNt = 0;
  int propConvention = 0, spanwiseLiftSwitch = 0, symmetricTau = 0, forceCalcSwitch = 0;
  double Omega = 0, omegaSlow = 0, shaftAxis[3] = {0, 0, 1}, hubCoords[3] = {0, 0, 0}, controlPitch[3] = {0, 0, 0};
  double apparentViscCoeff = 0, decayCoeff = 0, radius = 0, root_cut = 0, chord = 0, preconeAngle = 0, thetaTwist = 0;
  double pivotLE = 0, flapHinge = 0, psi = 0, psiStart = 0, pts[3] = {0, 0, 0}, cgCoords[3] = {0, 0, 0}, fromCoords[3] = {0, 0, 0};
  double velBody[3] = {0, 0, 0}, omegaBody[3] = {0, 0, 0};
  double xAxisBody[3] = {1, 0, 0}, yAxisBody[3] = {0, 1, 0}, zAxisBody[3] = {0, 0, 1};
  double dragUnitVec[3] = {0, 0, 0}, sideUnitVec[3] = {0, 0, 0}, liftUnitVec[3] = {0, 0, 0};
  double spanwiseCore = 0, rollupStartRadius = 0, rollupEndRadius = 0, initWakeVel = 0, skewLimit = 0, nonDimforceDenominator = 0;
  std::vector<double> streamwiseCoreVec;  // (ns+1)
  double forceInertial[3] = {0, 0, 0}, lift[3] = {0, 0, 0}, liftPrev[3] = {0, 0, 0}, drag[3] = {0, 0, 0}, liftUnsteady[3] = {0, 0, 0};
  std::vector<Blade> blade;
  std::vector<double> gamVec, gamVecPrev;
  int N() const { return nc * ns * nb; }
};

}  // namespace

struct vcase {
  int nr = 0, iter = 0;
  bool rotors_inited = false, inited = false, resident_started = false;
  double t = 0.0;
  Config cfg;
  std::vector<Geom> geom;
  std::vector<Rotor> rotor;
  vlc_ctx* ctx = nullptr;
  std::string err;
  long wing_uploads = 0;
  double stage_s[6] = {0, 0, 0, 0, 0, 0};  // host wall time per stage: motion, upload + prestep, RHS + solve, forces, wake stage, (unused)
};

namespace {

int fail(vcase* c, int code, const std::string& msg) {
  c->err = msg;
  return code;
}
// a status of the library becomes the case's error text (the shim turns it into `error stop`)
#define VK(c, call)                                                                                     \
  do {                                                                                                  \
    const int rc_ = (call);                                                                             \
    if (rc_) return fail((c), rc_, std::string(#call) + ": " + vlc_last_error((c)->ctx));              \
  } while (0)

double rotor_gettheta(const Rotor& r, double psi, int ib) {  // classdef.f90:4114-4136 (pitchDynamicsSwitch = 0)
  const double bladeOffset = 2.0 * pi_() / r.nb * (ib - 1);
  return r.controlPitch[0] + r.controlPitch[1] * std::cos(psi + bladeOffset) + r.controlPitch[2] * std::sin(psi + bladeOffset);
}
void rotor_move(Rotor& r, const double d[3]) {  // classdef.f90:4202-4214
  for (auto& b : r.blade) blade_move(b, d);
  for (int k = 0; k < 3; ++k) r.hubCoords[k] = r.hubCoords[k] + d[k], r.cgCoords[k] = r.cgCoords[k] + d[k];
}
void rotor_rot_pts(Rotor& r, const double pts[3], const double origin_in[3]) {  // classdef.f90:4216-4252 (order 1)
  double T[9], origin[3];
  v_copy(origin, origin_in);  // the reference passes this%cgCoords, which is itself rotated at the end
  Tgb(pts, T);
  for (auto& b : r.blade) blade_rot_pts(b, T, origin);
  matvec(T, r.shaftAxis, r.shaftAxis);
  matvec(T, r.xAxisBody, r.xAxisBody);
  matvec(T, r.yAxisBody, r.yAxisBody);
  matvec(T, r.zAxisBody, r.zAxisBody);
  rot_about(T, origin, r.hubCoords);
  rot_about(T, origin, r.cgCoords);
}
void rotor_rot_advance(Rotor& r, double dpsi, bool nopitch) {  // classdef.f90:4265-4291
  r.psi = r.psi + dpsi;
  for (int ib = 1; ib <= r.nb; ++ib) {
    Blade& b = r.blade[ib - 1];
    blade_rotate(b, dpsi, r.shaftAxis, r.hubCoords, ROT_AZIMUTH);
    b.psi = b.psi + dpsi;
    if (!nopitch) {
      const double thetaNext = rotor_gettheta(r, r.psi, ib);
      blade_rot_pitch(b, thetaNext - b.theta);
      b.theta = thetaNext;
    }
  }
}

void toChordsRevs(const Geom& g, int* nsteps, double dt) {  // classdef.f90:5133-5148
  if (*nsteps < 0) {
    if (std::fabs(g.Omega) < kEps)
      *nsteps = (int)std::ceil(std::abs(*nsteps) * g.chord / (dt * v_norm(g.velBody)));
    else
      *nsteps = (int)std::ceil(2.0 * pi_() * std::abs(*nsteps) / (std::fabs(g.Omega) * dt));
  }
}

// rotor%init for lifting surfaces (classdef.f90:2769-3891).  Also returns the wake records as rotor_init leaves them
// (zero circulation, core radii from spanwiseCore / streamwiseCoreVec, :3826-3859): they go up once (resident_begin).
int rotor_init(vcase* c, Geom& g, Rotor& r, std::vector<std::vector<Vr>>& waN, std::vector<std::vector<Fwake>>& waF) {
  Config& cfg = c->cfg;
  const double degToRad = pi_() / 180.0, twoPi = 2.0 * pi_();
  double dt = cfg.dt;
  int nt = cfg.nt;
  if (sgn1(dt) < 0.0) dt = (std::fabs(g.Omega) < kEps) ? std::fabs(dt) * g.chord / v_norm(g.velBody) : twoPi * std::fabs(dt) / std::fabs(g.Omega);  // :2966-2977
  if (std::fabs(dt) <= kEps) dt = (std::fabs(g.Omega) < kEps) ? (g.chord / g.nc) / v_norm(g.velBody) : 5.0 * degToRad / std::fabs(g.Omega);     // :2979-2987
  if (nt <= 0) {  // :2990-2996
    if (nt == 0) nt = -10;
    toChordsRevs(g, &nt, dt);
  }
  if (cfg.slowStart != 0) toChordsRevs(g, &cfg.slowStartNt, dt);
  toChordsRevs(g, &g.wakeTruncateNt, dt);
  toChordsRevs(g, &g.prescWakeAfterTruncNt, dt);
  toChordsRevs(g, &g.prescWakeGenNt, dt);
  toChordsRevs(g, &g.nNwake, dt);
  const int prescWakeNt = (g.wakeTruncateNt > 0 && g.prescWakeAfterTruncNt > 0) ? g.wakeTruncateNt + g.prescWakeAfterTruncNt : 0;  // :3013-3017
  if (g.wakeTruncateNt > 0 && g.wakeTruncateNt < g.nNwake + 1) g.wakeTruncateNt = g.nNwake + 1;                                   // :3019-3021
  if (g.surfaceType == 0) g.surfaceType = 1;
  if (g.nNwake > 0 && g.nNwake < 2) return fail(c, 3, "ERROR: Atleast 2 near wake rows mandatory");
  cfg.dt = dt;
  cfg.nt = nt;
  const int nNwake = g.nNwake < nt ? g.nNwake : nt;  // :3039-3055
  const int nFwake = (g.wakeTruncateNt == 0) ? nt - nNwake : g.wakeTruncateNt - nNwake;
  g.nNwake = nNwake;
  const int nc = g.nc, ns = g.ns, nb = g.nb;
  r.nb = nb, r.nc = nc, r.ns = ns, r.nNwake = nNwake, r.nFwake = nFwake;
  r.surfaceType = g.surfaceType;
  r.axisymmetrySwitch = g.axisymmetrySwitch;
  r.ductSwitch = g.ductSwitch;
  r.nbConvect = (g.axisymmetrySwitch == 1) ? 1 : nb;
  r.propConvention = g.propConvention;
  r.spanwiseLiftSwitch = g.spanwiseLiftSwitch;
  r.symmetricTau = g.symmetricTau;
  r.forceCalcSwitch = g.forceCalcSwitch;
  r.wakeTruncateNt = g.wakeTruncateNt;
  r.prescWakeNt = prescWakeNt;
  r.prescWakeAfterTruncNt = g.prescWakeAfterTruncNt;
  r.prescWakeGenNt = g.prescWakeGenNt;
  r.radius = g.span, r.root_cut = g.rootcut, r.chord = g.chord, r.Omega = g.Omega;
  r.pivotLE = g.pivotLE, r.flapHinge = g.flapHinge;
  r.apparentViscCoeff = g.apparentViscCoeff, r.decayCoeff = g.decayCoeff;
  r.rollupStartRadius = g.rollupStartRadius, r.rollupEndRadius = g.rollupEndRadius;
  r.initWakeVel = g.initWakeVel, r.skewLimit = g.skewLimit;
  v_copy(r.hubCoords, g.hubCoords), v_copy(r.cgCoords, g.cgCoords), v_copy(r.fromCoords, g.fromCoords);
  v_copy(r.shaftAxis, g.shaftAxis), v_copy(r.velBody, g.velBody), v_copy(r.omegaBody, g.omegaBody);
  v_copy(r.dragUnitVec, g.dragUnitVec), v_copy(r.sideUnitVec, g.sideUnitVec), v_copy(r.liftUnitVec, g.liftUnitVec);
  r.controlPitch[0] = g.theta0 * degToRad, r.controlPitch[1] = g.thetaC * degToRad, r.controlPitch[2] = g.thetaS * degToRad;  // :3126-3137
  for (int k = 0; k < 3; ++k) r.pts[k] = g.phiThetaPsi[k] * degToRad;
  r.thetaTwist = g.thetaTwist * degToRad;
  r.preconeAngle = g.preconeAngle * degToRad;
  r.psiStart = g.psiStart * degToRad;
  r.spanwiseCore = g.spanwiseCore * g.chord;
  {  // classdef.f90:2722-2726: a single-valued streamwiseCoreVec is broadcast
    const int nS = (int)g.streamwiseCoreVec.size();
    double rest = 0.0;
    for (int j = 1; j < nS; ++j) rest += g.streamwiseCoreVec[j] * g.streamwiseCoreVec[j];
    r.streamwiseCoreVec.assign((size_t)ns + 1, 0.0);
    for (int j = 0; j <= ns; ++j) {
      const double v = (nS <= 1 || std::sqrt(rest) < kEps) ? g.streamwiseCoreVec[0] : (j < nS ? g.streamwiseCoreVec[j] : 0.0);
      r.streamwiseCoreVec[j] = v * g.chord;
    }
  }
  r.rollupStart = (int)std::ceil(g.rollupStartRadius * ns);
  r.rollupEnd = (int)std::floor(g.rollupEndRadius * ns);
  r.gamVec.assign((size_t)r.N(), 0.0);
  r.gamVecPrev.assign((size_t)r.N(), 0.0);

  // :3152-3260 panel corner coordinates
  std::vector<double> xVec((size_t)nc + 1), yVec((size_t)ns + 1);
  if (g.grid.empty()) {
    const double c0 = (g.Omega >= 0) ? -g.chord : g.chord;
    spacing(g.chordSpacing, c0, 0.0, nc + 1, xVec.data());
    spacing(g.spanSpacing, g.rootcut * g.span, g.span, ns + 1, yVec.data());
  }
  r.blade.assign((size_t)nb, Blade());
  for (Blade& b : r.blade) {
    b.nc = nc, b.ns = ns;
    b.wiP.assign((size_t)nc * ns, WingPanel());
    std::memset(b.wiP.data(), 0, sizeof(WingPanel) * b.wiP.size());
    for (auto* v : {&b.secChord, &b.secArea, &b.secMflapArm}) v->assign((size_t)ns, 0.0);
    for (auto* v : {&b.secTauCapChord, &b.secTauCapSpan, &b.secNormalVec, &b.secCP}) v->assign(3 * (size_t)ns, 0.0);
    b.loads.assign((size_t)12 + 25 * ns, 0.0);
    std::memset(b.axes, 0, sizeof b.axes);
    for (int j = 1; j <= ns; ++j)
      for (int i = 1; i <= nc; ++i) {
        WingPanel& p = b.P(i, j);
        if (g.grid.empty()) {
          v_set(p.PC[0], xVec[i - 1], yVec[j - 1], 0.0);
          v_set(p.PC[1], xVec[i], yVec[j - 1], 0.0);
          v_set(p.PC[2], xVec[i], yVec[j], 0.0);
          v_set(p.PC[3], xVec[i - 1], yVec[j], 0.0);
        } else {  // rotor_plot3dtoblade :3957-4020: grid(:, ic, is)
          auto G = [&](int ic, int is) { return &g.grid[3 * ((size_t)(ic - 1) + (size_t)(nc + 1) * (is - 1))]; };
          v_copy(p.PC[0], G(i, j)), v_copy(p.PC[1], G(i + 1, j)), v_copy(p.PC[2], G(i + 1, j + 1)), v_copy(p.PC[3], G(i, j + 1));
        }
      }
  }
  for (int ib = 0; ib < r.nbConvect; ++ib) {  // :3268-3536
    Blade& b = r.blade[ib];
    for (int a = 0; a < 9; ++a) v_set(b.axes[a], a % 3 == 0, a % 3 == 1, a % 3 == 2);
    for (int k = 0; k < 3; ++k) b.flapOrigin[k] = b.yAxis()[k] * r.radius * r.flapHinge;
    for (int j = 1; j <= ns; ++j) {  // :3294-3309 section vectors
      double t[3], a[3], d[3], n[3];
      for (int k = 0; k < 3; ++k) {
        t[k] = (b.P(nc, j).PC[2][k] + b.P(nc, j).PC[1][k] - (b.P(1, j).PC[3][k] + b.P(1, j).PC[0][k])) * 0.5;
        a[k] = b.P(nc, j).PC[1][k] - b.P(1, j).PC[3][k];
        d[k] = b.P(nc, j).PC[2][k] - b.P(1, j).PC[0][k];
      }
      v_unit(t, b.s3(b.secTauCapChord, j));
      v_set(b.s3(b.secTauCapSpan, j), 0, 1, 0);
      v_cross(a, d, n);
      v_unit(n, n);
      for (int k = 0; k < 3; ++k) b.s3(b.secNormalVec, j)[k] = sgn1(r.Omega) * n[k];
    }
    for (int j = 1; j <= ns; ++j)  // :3311-3367 vortex-ring corners at the quarter-panel shift
      for (int i = 1; i <= nc; ++i) {
        WingPanel& p = b.P(i, j);
        double xs[4];
        xs[0] = (p.PC[1][0] - p.PC[0][0]) * 0.25;
        xs[3] = (p.PC[2][0] - p.PC[3][0]) * 0.25;
        if (i < nc) {
          xs[1] = (b.P(i + 1, j).PC[1][0] - p.PC[1][0]) * 0.25;
          xs[2] = (b.P(i + 1, j).PC[2][0] - p.PC[2][0]) * 0.25;
        } else {
          xs[1] = xs[2] = 0.0;
        }
        for (int n = 1; n <= 4; ++n) {
          const double P[3] = {p.PC[n - 1][0] + xs[n - 1], p.PC[n - 1][1], p.PC[n - 1][2]};
          vr_assignP(p.vr, n, P);
        }
      }
    double dxdymin = 1e300;  // :3377-3386
    for (auto& p : b.wiP) {
      double a[3], d[3];
      for (int k = 0; k < 3; ++k) a[k] = p.PC[1][k] - p.PC[0][k], d[k] = p.PC[2][k] - p.PC[1][k];
      dxdymin = std::fmin(dxdymin, std::fmin(std::fabs(v_norm(a)), std::fabs(v_norm(d))));
    }
    {  // :3388-3400 shed the last row (0.05 is a default-real literal in the reference)
      double velShed;
      if (std::fabs(r.Omega) > kEps) {
        double d[3];
        for (int k = 0; k < 3; ++k) d[k] = b.P(nc, ns).vr.vf[1].fc[0][k] - r.hubCoords[k];
        const double a = (double)0.05f * std::fabs(r.Omega) * v_norm(d), q = 0.125 * r.chord / dt;
        velShed = a < q ? a : q;
      } else {
        const double mv[3] = {-1.0 * r.velBody[0], -1.0 * r.velBody[1], -1.0 * r.velBody[2]};
        velShed = 0.3 * v_norm(mv);
      }
      const double d[3] = {sgn1(r.Omega) * velShed * dt, 0.0, 0.0};
      for (int j = 1; j <= ns; ++j) vr_shiftdP(b.P(nc, j).vr, 2, d), vr_shiftdP(b.P(nc, j).vr, 3, d);
    }
    for (int j = 1; j <= ns; ++j)  // :3402-3417
      for (int i = 1; i <= nc; ++i) {
        WingPanel& p = b.P(i, j);
        wp_calcCP(p);
        wp_calcN(p);
        if (sgn1(r.Omega) < 0.0)
          for (int k = 0; k < 3; ++k) p.nCap[k] = -1.0 * p.nCap[k];
        wp_calcTau(p);
        double m[3];
        for (int k = 0; k < 3; ++k) m[k] = (b.P(1, j).PC[0][k] + b.P(1, j).PC[3][k]) * 0.5 - p.CP[k];
        p.rHinge = v_norm(m);
        wp_calc_area(p);
        wp_calc_mean_dimensions(p);
      }
    blade_calc_secArea_secChord(b);  // :3446-3447
    for (int j = 1; j <= ns; ++j) {   // :3455-3464 overrideTauSpan
      v_copy(b.s3(b.secTauCapSpan, j), b.yAxis());
      for (int i = 1; i <= nc; ++i) v_copy(b.P(i, j).tauCapSpan, b.yAxis());
    }
    if (r.symmetricTau == 1)  // :3467-3476
      for (int j = 1; j <= ns / 2; ++j) {
        for (int k = 0; k < 3; ++k) b.s3(b.secTauCapSpan, j)[k] = -1.0 * b.s3(b.secTauCapSpan, j)[k];
        for (int i = 1; i <= nc; ++i)
          for (int k = 0; k < 3; ++k) b.P(i, j).tauCapSpan[k] = -1.0 * b.P(i, j).tauCapSpan[k];
      }
    blade_calc_secLocations(b, 0.5, r.flapHinge * r.radius);  // :3479-3480
    b.pivotLE = r.pivotLE;
    const double core = r.spanwiseCore < dxdymin * 0.1 ? r.spanwiseCore : dxdymin * 0.1;  // :3498-3519
    for (auto& p : b.wiP)
      for (int f = 0; f < 4; ++f) p.vr.vf[f].rVc0 = core;
    for (int j = 1; j <= ns; ++j) b.P(nc, j).vr.vf[1].rVc0 = r.spanwiseCore;
    for (auto& p : b.wiP)
      for (int f = 0; f < 4; ++f) p.vr.vf[f].rVc = p.vr.vf[f].rVc0;
  }
  if (r.axisymmetrySwitch == 1)  // :3539-3630 copy blade 1 to the others
    for (int ib = 1; ib < nb; ++ib) {
      Blade &b = r.blade[ib], &b1 = r.blade[0];
      std::memcpy(b.axes, b1.axes, sizeof b.axes);
      v_copy(b.flapOrigin, b1.flapOrigin);
      b.secTauCapChord = b1.secTauCapChord, b.secTauCapSpan = b1.secTauCapSpan, b.secNormalVec = b1.secNormalVec;
      for (size_t q = 0; q < b.wiP.size(); ++q) {
        WingPanel &p = b.wiP[q], &p1 = b1.wiP[q];
        p.vr = p1.vr;
        v_copy(p.CP, p1.CP), v_copy(p.nCap, p1.nCap), v_copy(p.tauCapChord, p1.tauCapChord), v_copy(p.tauCapSpan, p1.tauCapSpan);
        p.rHinge = p1.rHinge, p.panelArea = p1.panelArea, p.meanChord = p1.meanChord, p.meanSpan = p1.meanSpan;
      }
      b.secArea = b1.secArea, b.secChord = b1.secChord, b.secCP = b1.secCP, b.secMflapArm = b1.secMflapArm;
      b.pivotLE = b1.pivotLE;
    }
  {  // :3632-3635
    double d[3];
    for (int k = 0; k < 3; ++k) d[k] = r.hubCoords[k] - r.fromCoords[k];
    for (auto& b : r.blade) blade_move(b, d);
  }
  for (auto& b : r.blade) {  // :3638-3657 (flapInitial = 0: bladeDynamicsSwitch = 0 everywhere)
    b.preconeAngle = r.preconeAngle;
    blade_rot_flap(b, r.preconeAngle);
    blade_rot_flap(b, 0.0);
  }
  for (int ib = 2; ib <= nb; ++ib) {  // :3661-3667
    const double bladeOffset = sgn1(r.Omega) * twoPi / nb * (ib - 1);
    blade_rotate(r.blade[ib - 1], bladeOffset, r.shaftAxis, r.hubCoords, ROT_AZIMUTH);
  }
  rotor_rot_pts(r, r.pts, r.cgCoords);                       // :3670
  rotor_rot_advance(r, sgn1(r.Omega) * r.psiStart, true);    // :3673
  if (std::fabs(r.Omega) > kEps) {                           // :3676-3691
    if (r.propConvention == 0)
      r.nonDimforceDenominator = cfg.density * (pi_() * (r.radius * r.radius)) * ((r.radius * r.Omega) * (r.radius * r.Omega));
    else
      r.nonDimforceDenominator = cfg.density * ((r.Omega / twoPi) * (r.Omega / twoPi)) * std::pow(2.0 * r.radius, 4.0);
  } else {
    r.nonDimforceDenominator = 0.5 * cfg.density * (r.radius * (1.0 - r.root_cut) * r.chord) * v_dot(r.velBody, r.velBody);
  }
  // :3826-3859 wake core radii; gam = 0
  waN.assign((size_t)nb, std::vector<Vr>((size_t)nNwake * ns));
  waF.assign((size_t)nb, std::vector<Fwake>((size_t)nFwake));
  for (int ib = 0; ib < nb; ++ib) {
    if (!waN[ib].empty()) std::memset(waN[ib].data(), 0, sizeof(Vr) * waN[ib].size());
    if (!waF[ib].empty()) std::memset(waF[ib].data(), 0, sizeof(Fwake) * waF[ib].size());
    for (int j = 1; j <= ns; ++j)
      for (int i = 1; i <= nNwake; ++i) {
        Vr& w = waN[ib][(size_t)(i - 1) + (size_t)nNwake * (j - 1)];
        w.vf[1].rVc0 = w.vf[1].rVc = r.spanwiseCore;
        w.vf[3].rVc0 = w.vf[3].rVc = r.spanwiseCore;
        w.vf[0].rVc0 = w.vf[0].rVc = r.streamwiseCoreVec[j - 1];
        w.vf[2].rVc0 = w.vf[2].rVc = r.streamwiseCoreVec[j];
      }
    for (auto& f : waF[ib]) f.vf.rVc0 = f.vf.rVc = r.streamwiseCoreVec[ns];
  }
  // :3861-3889 wind frame (force2file's signLift uses zAxisBody only; kept for completeness)
  if (v_norm(r.dragUnitVec) <= kEps && v_norm(r.sideUnitVec) <= kEps && v_norm(r.liftUnitVec) <= kEps) {
    if (std::fabs(r.Omega) <= kEps) {
      double u[3];
      if (std::fabs(r.velBody[0]) > kEps) {
        v_set(u, r.velBody[0], 0.0, r.velBody[2]);
        v_set(r.sideUnitVec, 0, 1, 0);
      } else {
        v_set(u, 0.0, r.velBody[1], r.velBody[2]);
        v_set(r.sideUnitVec, 1, 0, 0);
      }
      v_unit(u, u);
      for (int k = 0; k < 3; ++k) r.dragUnitVec[k] = -1.0 * u[k];
      v_cross(r.dragUnitVec, r.sideUnitVec, r.liftUnitVec);
    } else {
      v_copy(r.liftUnitVec, r.shaftAxis);
    }
  }
  return 0;
}

// velCP = velCPm = the kinematic velocity at the collocation points (main.f90:528-547): stays with the driver
void kinematic_velCP(Rotor& r) {
  for (int ib = 0; ib < r.nbConvect; ++ib) {
    Blade& b = r.blade[ib];
    for (int is = 1; is <= r.ns; ++is)
      for (int ic = 1; ic <= r.nc; ++ic) {
        WingPanel& p = b.P(ic, is);
        double d1[3], d2[3], w[3], c1[3], c2[3];
        for (int k = 0; k < 3; ++k) {
          d1[k] = p.CP[k] - r.cgCoords[k];
          d2[k] = p.CP[k] - r.hubCoords[k];
          w[k] = r.omegaSlow * r.shaftAxis[k];
        }
        v_cross(r.omegaBody, d1, c1);
        v_cross(w, d2, c2);
        const double flapTerm = b.secMflapArm[is - 1] * b.dflap;  // scalar, broadcast over xyz as in the reference
        for (int k = 0; k < 3; ++k) {
          p.velCP[k] = ((-1.0 * r.velBody[k] - c1[k]) - c2[k]) - flapTerm;
          p.velCPm[k] = p.velCP[k];
        }
      }
  }
}

// the library's copy of rotor ir becomes current: row counters, frame, and the wing records of every blade
int sync_wing(vcase* c, int ir) {
  Rotor& r = c->rotor[ir];
  VK(c, vlc_rotor_set_rows(c->ctx, ir, r.rowNear, r.rowFar));
  VK(c, vlc_rotor_set_frame(c->ctx, ir, r.shaftAxis, r.hubCoords));
  for (int ib = 0; ib < r.nb; ++ib) VK(c, vlc_rotor_put_wing(c->ctx, ir, ib, reinterpret_cast<const double*>(r.blade[ib].wiP.data())));
  c->wing_uploads++;
  return 0;
}

// map_gam on the driver's own records (classdef.f90:4181-4196): they travel again with the next move of the wing
void map_gam(Rotor& r) {
  const int npb = r.nc * r.ns;
  for (int ib = 0; ib < r.nb; ++ib)
    for (int q = 0; q < npb; ++q) r.blade[ib].wiP[q].vr.gam = r.gamVec[(size_t)q + (size_t)npb * ib];
}

// One pass of main.f90:548-603 for every rotor on the device (tier 2c): induced velCP + RHS of every rotor first, then
// every rotor's solve + map_gam -- the order of the reference's two loops over ir.
int rhs_solve_pass(vcase* c, bool reset) {
  for (int ir = 0; ir < c->nr; ++ir) {
    if (reset) VK(c, vlc_rotor_reset_velCP(c->ctx, ir));
    VK(c, vlc_rotor_calc_RHS(c->ctx, ir, nullptr, nullptr));
  }
  for (int ir = 0; ir < c->nr; ++ir) {
    Rotor& r = c->rotor[ir];
    r.gamVecPrev = r.gamVec;  // :591
    VK(c, vlc_rotor_solve_map_gam(c->ctx, ir, r.gamVec.data()));
    map_gam(r);
  }
  return 0;
}
// ntSubLoop (main.f90:522-615, initial :120-211): passes 0 .. ntSub, left early when the circulation stops changing
int rhs_solve(vcase* c, int ntSub) {
  for (int i = 0; i <= ntSub; ++i) {
    int rc = rhs_solve_pass(c, i > 0);
    if (rc) return rc;
    if (ntSub != 0) {  // :606-613
      double res = 0.0;
      bool done = false;
      for (int ir = 0; ir < c->nr && !done; ++ir) {
        const Rotor& r = c->rotor[ir];
        double s = 0.0;
        for (size_t k = 0; k < r.gamVec.size(); ++k) s += (r.gamVec[k] - r.gamVecPrev[k]) * (r.gamVec[k] - r.gamVecPrev[k]);
        res = std::fmax(res, std::sqrt(s));
        if (res <= kEps) done = true;
      }
      if (done) break;
    }
  }
  return 0;
}

// sumBladeToNetForces (classdef.f90:4954-4988) with the copies of an axisymmetric rotor (:4623-4671)
void sum_forces(Rotor& r) {
  v_copy(r.liftPrev, r.lift);
  if (r.axisymmetrySwitch == 1) {
    for (int ib = 1; ib < r.nb; ++ib) {
      Blade &b = r.blade[ib], &b1 = r.blade[0];
      v_copy(b.forceInertial, b1.forceInertial), v_copy(b.lift, b1.lift), v_copy(b.drag, b1.drag), v_copy(b.liftUnsteady, b1.liftUnsteady);
    }
    for (int k = 0; k < 3; ++k) {
      r.forceInertial[k] = r.nb * r.blade[0].forceInertial[k];
      r.lift[k] = r.nb * r.blade[0].lift[k];
      r.drag[k] = r.nb * r.blade[0].drag[k];
      r.liftUnsteady[k] = r.nb * r.blade[0].liftUnsteady[k];
    }
  } else {
    v_set(r.forceInertial, 0, 0, 0), v_set(r.lift, 0, 0, 0), v_set(r.drag, 0, 0, 0), v_set(r.liftUnsteady, 0, 0, 0);
    for (int ib = 0; ib < r.nbConvect; ++ib)
      for (int k = 0; k < 3; ++k) {
        r.forceInertial[k] = r.forceInertial[k] + r.blade[ib].forceInertial[k];
        r.lift[k] = r.lift[k] + r.blade[ib].lift[k];
        r.drag[k] = r.drag[k] + r.blade[ib].drag[k];
        r.liftUnsteady[k] = r.liftUnsteady[k] + r.blade[ib].liftUnsteady[k];
      }
  }
}

// main.f90:630-670 for every rotor: velCPTotal, calc_secAlpha, calc_force on the device down to the blade sums; the
// wing records come back (gamPrev, delP, normalForce ... travel again with the next move) and the blades are added here
int compute_forces(vcase* c) {
  for (int ir = 0; ir < c->nr; ++ir) {
    Rotor& r = c->rotor[ir];
    if (r.forceCalcSwitch != 0) return fail(c, 3, "forceCalcSwitch " + std::to_string(r.forceCalcSwitch) + " needs C81 tables: outside this driver's scope");
    const int ns = r.ns;
    std::vector<double> sec((size_t)10 * ns + 6);
    for (int ib = 0; ib < r.nb; ++ib) {
      Blade& b = r.blade[ib];
      std::memcpy(&sec[0], b.secTauCapChord.data(), sizeof(double) * 3 * (size_t)ns);
      std::memcpy(&sec[3 * (size_t)ns], b.secNormalVec.data(), sizeof(double) * 3 * (size_t)ns);
      std::memcpy(&sec[6 * (size_t)ns], b.secCP.data(), sizeof(double) * 3 * (size_t)ns);
      std::memcpy(&sec[9 * (size_t)ns], b.secArea.data(), sizeof(double) * (size_t)ns);
      std::memcpy(&sec[10 * (size_t)ns], b.yAxisAziFlap(), 3 * sizeof(double));
      std::memcpy(&sec[10 * (size_t)ns + 3], b.zAxisAziFlap(), 3 * sizeof(double));
      VK(c, vlc_rotor_put_sections(c->ctx, ir, ib, sec.data()));
    }
    VK(c, vlc_rotor_calc_velCPTotal(c->ctx, ir));
    VK(c, vlc_rotor_calc_force(c->ctx, ir, c->cfg.density, c->cfg.dt, r.Omega, r.spanwiseLiftSwitch));
    for (int ib = 0; ib < r.nb; ++ib) {
      Blade& b = r.blade[ib];
      if (ib >= r.nbConvect && r.axisymmetrySwitch != 1) continue;
      VK(c, vlc_rotor_get_wing(c->ctx, ir, ib, reinterpret_cast<double*>(b.wiP.data())));
      VK(c, vlc_rotor_get_loads(c->ctx, ir, ib, b.loads.data()));
      std::memcpy(b.forceInertial, &b.loads[0], 3 * sizeof(double));
      std::memcpy(b.lift, &b.loads[3], 3 * sizeof(double));
      std::memcpy(b.drag, &b.loads[6], 3 * sizeof(double));
      std::memcpy(b.liftUnsteady, &b.loads[9], 3 * sizeof(double));
    }
    sum_forces(r);
  }
  return 0;
}

// first wake stage of a step (main.f90:466-506): the wake goes up once, whole arrays, current and predicted
int resident_begin(vcase* c, std::vector<std::vector<std::vector<Vr>>>& waN, std::vector<std::vector<std::vector<Fwake>>>& waF) {
  for (int ir = 0; ir < c->nr; ++ir) {
    Rotor& r = c->rotor[ir];
    VK(c, vlc_rotor_set_wake_params(c->ctx, ir, r.nbConvect, r.axisymmetrySwitch, r.ductSwitch, r.suppressFwakeSwitch, r.rollupStart,
                                    r.rollupEnd, r.Omega * r.controlPitch[0], r.apparentViscCoeff, r.decayCoeff, r.initWakeVel));
    VK(c, vlc_rotor_set_rows(c->ctx, ir, 1, 1));
    for (int ib = 0; ib < r.nb; ++ib)
      for (int s = 0; s < 2; ++s) {
        if (r.nNwake > 0) VK(c, vlc_rotor_put_nwake(c->ctx, ir, ib, s, reinterpret_cast<const double*>(waN[ir][ib].data())));
        if (r.nFwake > 0) VK(c, vlc_rotor_put_fwake(c->ctx, ir, ib, s, reinterpret_cast<const double*>(waF[ir][ib].data())));
      }
    VK(c, vlc_rotor_set_rows(c->ctx, ir, r.rowNear, r.rowFar));
  }
  c->resident_started = true;
  return 0;
}

int convect(vcase* c, int ir, int iter, double dt, int p) {  // rotor%convectwake incl. its last statement (:4826-4828)
  Rotor& r = c->rotor[ir];
  VK(c, vlc_rotor_convectwake(c->ctx, ir, dt, p));
  if (r.prescWakeNt > 0 && iter > r.prescWakeNt) VK(c, vlc_rotor_updatePrescribedWake(c->ctx, ir, r.omegaSlow * dt, r.prescWakeGenNt, p));
  return 0;
}

// main.f90:800-1440: the wake sweeps, the fdScheme switch with its velocity bookkeeping, strain_wake, rollup, assignshed('TE')
int wake_convect(vcase* c, int iter) {
  const Config& cfg = c->cfg;
  const double dt = cfg.dt;
  const int addInit = iter < cfg.initWakeVelNt, nr = c->nr;
  vlc_ctx* x = c->ctx;
  VK(c, vlc_wake_sweep(x, 0, addInit));
  switch (cfg.fdScheme) {
    case 0:  // :846-859 explicit Euler
      for (int ir = 0; ir < nr; ++ir)
        if (int rc = convect(c, ir, iter, dt, 0)) return rc;
      break;
    case 1:  // :861-949 predictor-corrector
      for (int ir = 0; ir < nr; ++ir) {
        VK(c, vlc_rotor_wake_to_predicted(x, ir));
        if (int rc = convect(c, ir, iter, dt, 1)) return rc;
      }
      VK(c, vlc_wake_sweep(x, 1, addInit));
      for (int ir = 0; ir < nr; ++ir) {
        VK(c, vlc_rotor_wakevel_op(x, ir, VLC_VEL_ORDER2));
        if (int rc = convect(c, ir, iter, dt, 0)) return rc;
      }
      break;
    case 2:  // :951-1000 explicit Adams-Bashforth
      for (int ir = 0; ir < nr; ++ir) {
        if (iter == 1) {
          if (int rc = convect(c, ir, iter, dt, 0)) return rc;
          VK(c, vlc_rotor_wakevel_op(x, ir, VLC_VEL_FIRST_STEP));
        } else {
          VK(c, vlc_rotor_wakevel_op(x, ir, VLC_VEL_AB2));
          VK(c, vlc_rotor_wakevel_op(x, ir, VLC_VEL_FIRST_STEP));
          VK(c, vlc_rotor_wakevel_op(x, ir, VLC_VEL_COPY_TO_STEP));
          if (int rc = convect(c, ir, iter, dt, 0)) return rc;
        }
      }
      break;
    case 3:  // :1002-1115 Adams-Bashforth predictor / Adams-Moulton corrector
      if (iter == 1) {
        for (int ir = 0; ir < nr; ++ir) {
          if (int rc = convect(c, ir, iter, dt, 0)) return rc;
          VK(c, vlc_rotor_wakevel_op(x, ir, VLC_VEL_FIRST_STEP));
        }
      } else {
        for (int ir = 0; ir < nr; ++ir) {
          VK(c, vlc_rotor_wake_to_predicted(x, ir));
          VK(c, vlc_rotor_wakevel_op(x, ir, VLC_VEL_AB2));
          if (int rc = convect(c, ir, iter, dt, 1)) return rc;
        }
        VK(c, vlc_wake_sweep(x, 1, addInit));
        for (int ir = 0; ir < nr; ++ir) {
          VK(c, vlc_rotor_wakevel_op(x, ir, VLC_VEL_AM2));
          if (int rc = convect(c, ir, iter, dt, 0)) return rc;
          VK(c, vlc_rotor_wakevel_op(x, ir, VLC_VEL_SHIFT_HISTORY));
        }
      }
      break;
    case 4:    // :1117-1248 third-order Adams-Bashforth / Adams-Moulton
    case 5: {  // :1250-1404 fourth order: steps 1, 2, 3 fill vel1, vel2, vel3
      const int order = cfg.fdScheme == 4 ? 3 : 4;
      const int start = (order == 3) ? (iter == 2 ? 2 : 0) : (iter <= 3 ? iter : 0);
      static const int hist[4] = {0, VLC_VEL_ARRAY_1, VLC_VEL_ARRAY_2, VLC_VEL_ARRAY_3};
      static const int p3[3] = {VLC_VEL_ARRAY, VLC_VEL_ARRAY_2, VLC_VEL_ARRAY_1};
      static const double pc3[3] = {23.0, -16.0, 5.0};
      static const int c3[3] = {VLC_VEL_ARRAY_PREDICTED, VLC_VEL_ARRAY_STEP, VLC_VEL_ARRAY_2};
      static const double cc3[3] = {5.0, 8.0, -1.0};
      static const int p4[4] = {VLC_VEL_ARRAY, VLC_VEL_ARRAY_3, VLC_VEL_ARRAY_2, VLC_VEL_ARRAY_1};
      static const double pc4[4] = {55.0, -59.0, 37.0, -9.0};
      static const int c4[4] = {VLC_VEL_ARRAY_PREDICTED, VLC_VEL_ARRAY_STEP, VLC_VEL_ARRAY_3, VLC_VEL_ARRAY_2};
      static const double cc4[4] = {9.0, 19.0, -5.0, 1.0};
      if (start) {
        for (int ir = 0; ir < nr; ++ir) {
          if (int rc = convect(c, ir, iter, dt, 0)) return rc;
          VK(c, vlc_rotor_wakevel_copy(x, ir, hist[start], VLC_VEL_ARRAY));
        }
      } else {
        for (int ir = 0; ir < nr; ++ir) {
          VK(c, vlc_rotor_wake_to_predicted(x, ir));
          VK(c, vlc_rotor_wakevel_copy(x, ir, VLC_VEL_ARRAY_STEP, VLC_VEL_ARRAY));
          if (order == 3) VK(c, vlc_rotor_wakevel_lincomb(x, ir, VLC_VEL_ARRAY, 3, p3, pc3, 12.0));
          else VK(c, vlc_rotor_wakevel_lincomb(x, ir, VLC_VEL_ARRAY, 4, p4, pc4, 24.0));
          if (int rc = convect(c, ir, iter, dt, 1)) return rc;
        }
        VK(c, vlc_wake_sweep(x, 1, addInit));
        for (int ir = 0; ir < nr; ++ir) {
          if (order == 3) VK(c, vlc_rotor_wakevel_lincomb(x, ir, VLC_VEL_ARRAY, 3, c3, cc3, 12.0));
          else VK(c, vlc_rotor_wakevel_lincomb(x, ir, VLC_VEL_ARRAY, 4, c4, cc4, 24.0));
          if (int rc = convect(c, ir, iter, dt, 0)) return rc;
          VK(c, vlc_rotor_wakevel_copy(x, ir, VLC_VEL_ARRAY_1, VLC_VEL_ARRAY_2));
          if (order == 3) {
            VK(c, vlc_rotor_wakevel_copy(x, ir, VLC_VEL_ARRAY_2, VLC_VEL_ARRAY_STEP));
          } else {
            VK(c, vlc_rotor_wakevel_copy(x, ir, VLC_VEL_ARRAY_2, VLC_VEL_ARRAY_3));
            VK(c, vlc_rotor_wakevel_copy(x, ir, VLC_VEL_ARRAY_3, VLC_VEL_ARRAY_STEP));
          }
        }
      }
    } break;
    default: return fail(c, 3, "fdScheme " + std::to_string(cfg.fdScheme) + " is outside this driver's scope (0 ... 5)");
  }
  if (cfg.wakeStrain == 1)  // :1409-1416
    for (int ir = 0; ir < nr; ++ir) VK(c, vlc_rotor_strain_wake(x, ir));
  for (int ir = 0; ir < nr; ++ir) {  // :1419-1439
    const Rotor& r = c->rotor[ir];
    if (r.nNwake <= 0) continue;
    if (r.rowNear == 1) VK(c, vlc_rotor_rollup(x, ir));
    VK(c, vlc_rotor_assignshed(x, ir, 1));
  }
  return 0;
}

}  // namespace

// ============================================================================ C API (volcanor_b200/run_case.py binds it)
extern "C" {

vcase* vcase_new(int nr) {
  if (nr < 1 || nr > 64) return nullptr;
  vcase* c = new vcase();
  c->nr = nr;
  c->cfg.nr = nr;
  c->geom.assign((size_t)nr, Geom());
  c->rotor.assign((size_t)nr, Rotor());
  return c;
}
void vcase_free(vcase* c) { delete c; }
const char* vcase_error(const vcase* c) { return c ? c->err.c_str() : "null case"; }

// key = namelist variable of config.nml (libCommon.f90:51-108); 0 = known key
int vcase_set_config(vcase* c, const char* key, double v) {
  Config& g = c->cfg;
  const std::string k(key);
#define CI(name) if (k == #name) { g.name = (int)v; return 0; }
#define CD(name) if (k == #name) { g.name = v; return 0; }
  CI(nt) CD(dt) CD(density) CD(velSound) CD(kinematicVisc) CI(ntSub) CI(ntSubInit) CI(rotorForcePlot) CI(wakeDissipation)
  CI(wakeStrain) CI(wakeBurst) CI(wakeSuppress) CI(slowStart) CI(slowStartNt) CI(fdScheme) CI(initWakeVelNt)
#undef CI
#undef CD
  if (k == "nr") return ((int)v == c->nr) ? 0 : 2;
  return 1;  // a key the hot path does not need (plot switches, restart, ...): ignored by the caller
}

// key = namelist variable of geomNN.nml (classdef.f90:2541-2767), ir 0-based; "grid" = PLOT3D (3, nc+1, ns+1)
int vcase_set_geom(vcase* c, int ir, const char* key, int n, const double* x) {
  if (ir < 0 || ir >= c->nr || n < 1 || !x) return 2;
  Geom& g = c->geom[ir];
  const std::string k(key);
#define GI(name) if (k == #name) { g.name = (int)x[0]; return 0; }
#define GD(name) if (k == #name) { g.name = x[0]; return 0; }
#define G3(name) if (k == #name) { if (n != 3) return 2; for (int q = 0; q < 3; ++q) g.name[q] = x[q]; return 0; }
  GI(surfaceType) GI(nb) GI(propConvention) GI(spanSpacing) GI(chordSpacing) GI(nc) GI(ns) GI(nNwake)
  G3(hubCoords) G3(cgCoords) G3(fromCoords) G3(phiThetaPsi) GD(span) GD(rootcut) GD(chord) GD(preconeAngle) GD(Omega) G3(shaftAxis)
  GD(theta0) GD(thetaC) GD(thetaS) GD(thetaTwist) GI(ductSwitch) GI(axisymmetrySwitch) GI(spanwiseLiftSwitch) GI(symmetricTau)
  GI(forceCalcSwitch) GD(pivotLE) GD(flapHinge) G3(velBody) G3(omegaBody) GD(apparentViscCoeff) GD(decayCoeff) GI(wakeTruncateNt)
  GI(prescWakeAfterTruncNt) GI(prescWakeGenNt) GD(spanwiseCore) GD(rollupStartRadius) GD(rollupEndRadius) GD(initWakeVel)
  GD(psiStart) GD(skewLimit) G3(dragUnitVec) G3(sideUnitVec) G3(liftUnitVec)
#undef GI
#undef GD
#undef G3
  if (k == "streamwiseCoreVec") {
    g.streamwiseCoreVec.assign(x, x + n);
    return 0;
  }
  if (k == "grid") {
    g.grid.assign(x, x + n);
    return 0;
  }
  return 1;
}

// The library context that does the work: vlc_create (one GPU) or vlc_create_multi (several behind one handle).
int vcase_attach(vcase* c, vlc_ctx* ctx) {
  if (!ctx) return fail(c, VLC_ERR_NODEVICE, "no library context: volcanor_b200 has no CPU path");
  c->ctx = ctx;
  return 0;
}

// main.f90:1-382: rotor%init of every rotor, initial pitch, AIC, the initial solution and forces
int vcase_init(vcase* c) {
  if (c->inited) return 0;
  if (!c->ctx) return fail(c, VLC_ERR_NODEVICE, "vcase_attach first: every stage of the hot path runs in the CUDA library");
  std::vector<std::vector<std::vector<Vr>>> waN((size_t)c->nr);
  std::vector<std::vector<std::vector<Fwake>>> waF((size_t)c->nr);
  for (int ir = 0; ir < c->nr; ++ir) {  // :31-40
    Geom& g = c->geom[ir];
    if (g.surfaceType < 0 || g.surfaceType == 2) return fail(c, 3, "image / non-lifting surfaces are outside this driver's scope");
    if (!g.grid.empty() && (int)g.grid.size() != 3 * (g.nc + 1) * (g.ns + 1)) return fail(c, 3, "ERROR: Wrong or conflicting data in PLOT3D file");
    if (int rc = rotor_init(c, g, c->rotor[ir], waN[ir], waF[ir])) return rc;
  }
  for (Rotor& r : c->rotor)  // :43-58
    for (int ib = 1; ib <= r.nb; ++ib) {
      r.blade[ib - 1].theta = rotor_gettheta(r, r.psiStart, ib);
      blade_rot_pitch(r.blade[ib - 1], sgn1(r.Omega) * r.blade[ib - 1].theta);
    }
  c->rotors_inited = true;
  VK(c, vlc_rotors_clear(c->ctx));
  for (int ir = 0; ir < c->nr; ++ir) {
    const Rotor& r = c->rotor[ir];
    VK(c, vlc_rotor_define(c->ctx, ir, r.nb, r.nc, r.ns, r.nNwake, r.nFwake, r.surfaceType));
  }
  for (int ir = 0; ir < c->nr; ++ir) {  // :97-105, :227-230
    Rotor& r = c->rotor[ir];
    r.omegaSlow = (c->cfg.slowStart > 0) ? 0.0 : r.Omega;
    r.rowFar = r.nFwake + 1;
    r.rowNear = r.nNwake + 1;
  }
  if (int rc = resident_begin(c, waN, waF)) return rc;
  for (int ir = 0; ir < c->nr; ++ir) {  // wing up with the kinematic velCP in its records, then :65-81
    kinematic_velCP(c->rotor[ir]);
    if (int rc = sync_wing(c, ir)) return rc;
  }
  for (int ir = 0; ir < c->nr; ++ir) VK(c, vlc_rotor_calcAIC(c->ctx, ir, nullptr));  // 'Matrix is numerically singular!' -> VLC_ERR_SINGULAR
  c->t = 0.0;
  c->iter = 0;
  // :120-211 the initial solution: no wake row is active (rowNear = nNwake + 1), so vlc_rotor_calc_RHS adds exactly what
  // the reference's loop adds -- vind_bywing of the other rotors
  if (int rc = rhs_solve(c, c->cfg.ntSubInit)) return rc;
  for (int ir = 0; ir < c->nr; ++ir)  // :231-234
    if (c->rotor[ir].nNwake > 0) VK(c, vlc_rotor_assignshed(c->ctx, ir, 1));
  if (c->cfg.rotorForcePlot != 0)
    if (int rc = compute_forces(c)) return rc;
  c->inited = true;
  return 0;
}

// one pass of the time loop main.f90:400-1452
int vcase_step(vcase* c) {
  if (!c->inited)
    if (int rc = vcase_init(c)) return rc;
  const Config& cfg = c->cfg;
  const double dt = cfg.dt;
  c->iter += 1;
  c->t = c->t + dt;
  const int iter = c->iter;
  using clk = std::chrono::steady_clock;
  auto t0 = clk::now();
  auto lap = [&](int k) {
    const auto t1 = clk::now();
    c->stage_s[k] += std::chrono::duration<double>(t1 - t0).count();
    t0 = t1;
  };
  for (Rotor& r : c->rotor) {  // :412-417
    r.rowNear = r.rowNear - 1 > 1 ? r.rowNear - 1 : 1;
    if (iter > r.nNwake) r.rowFar = r.rowFar - 1 > 1 ? r.rowFar - 1 : 1;
  }
  for (Rotor& r : c->rotor) {  // :428-452
    switch (cfg.slowStart) {
      case 0: r.omegaSlow = r.Omega; break;
      case 1: {
        const float a = (float)cfg.slowStartNt, bq = (float)(iter + 1);  // min(real(..), real(..)): default real
        r.omegaSlow = (double)(a < bq ? a : bq) * r.Omega / cfg.slowStartNt;
      } break;
      case 2: r.omegaSlow = std::tanh(5.0 * iter / cfg.slowStartNt) * r.Omega; break;
      case 3: r.omegaSlow = (std::tanh((double)(6.0f * 1.0f) * (double)((float)iter / (float)cfg.slowStartNt) - 3.0) + 1.0) * 0.5 * r.Omega; break;
      default: break;
    }
  }
  for (Rotor& r : c->rotor) {  // :455-463
    const double d[3] = {r.velBody[0] * dt, r.velBody[1] * dt, r.velBody[2] * dt};
    const double w[3] = {r.omegaBody[0] * dt, r.omegaBody[1] * dt, r.omegaBody[2] * dt};
    rotor_move(r, d);
    rotor_rot_pts(r, w, r.cgCoords);
    rotor_rot_advance(r, r.omegaSlow * dt, false);
  }
  // the moved wing goes up ONCE per step, with the kinematic velCP (:528-547) already in its records
  for (int ir = 0; ir < c->nr; ++ir) kinematic_velCP(c->rotor[ir]);
  lap(0);
  for (int ir = 0; ir < c->nr; ++ir)
    if (int rc = sync_wing(c, ir)) return rc;
  if (cfg.wakeSuppress == 0) {  // :466-506
    for (int ir = 0; ir < c->nr; ++ir)
      if (c->rotor[ir].nNwake > 0) VK(c, vlc_rotor_assignshed(c->ctx, ir, 0));
    for (int ir = 0; ir < c->nr; ++ir) VK(c, vlc_rotor_age_wake(c->ctx, ir, dt, c->rotor[ir].omegaSlow));
    if (cfg.wakeDissipation == 1)
      for (int ir = 0; ir < c->nr; ++ir) VK(c, vlc_rotor_dissipate_wake(c->ctx, ir, dt, cfg.kinematicVisc));
    if (cfg.wakeBurst != 0 && iter % cfg.wakeBurst == 0)  // :490-497
      for (int ir = 0; ir < c->nr; ++ir)
        if (c->rotor[ir].nNwake > 0) VK(c, vlc_rotor_burst_wake(c->ctx, ir, c->rotor[ir].skewLimit, c->rotor[ir].chord));
  }
  lap(1);
  if (int rc = rhs_solve(c, cfg.ntSub)) return rc;  // :522-615
  lap(2);
  if (cfg.rotorForcePlot != 0 && iter % cfg.rotorForcePlot == 0)  // :624-722
    if (int rc = compute_forces(c)) return rc;
  lap(3);
  if (cfg.wakeSuppress == 0)  // :800-1440
    if (int rc = wake_convect(c, iter)) return rc;
  lap(4);
  return 0;
}

// host wall time spent in the stages of vcase_step so far: motion | wing upload + wake pre-step | RHS + solve | forces |
// wake stage.  The stages end in a device read-back (gamVec; wing + loads) or run asynchronously (wake stage), so the
// numbers say where the HOST waits, which for a small case is where the time goes.
void vcase_stage_seconds(const vcase* c, double out[5]) { std::memcpy(out, c->stage_s, 5 * sizeof(double)); }

int vcase_iter(const vcase* c) { return c->iter; }
// nt, dt as rotor%init left them (chords / revolutions resolved), per-rotor sizes and row counters
int vcase_info(const vcase* c, int ir, double* out /* [12] */) {
  if (ir < 0 || ir >= c->nr) return 2;
  const Rotor& r = c->rotor[ir];
  const double v[12] = {(double)c->cfg.nt, c->cfg.dt, (double)r.nb, (double)r.nc, (double)r.ns, (double)r.nNwake, (double)r.nFwake,
                        (double)r.rowNear, (double)r.rowFar, r.nonDimforceDenominator, (double)r.nbConvect, (double)c->wing_uploads};
  std::memcpy(out, v, sizeof v);
  return 0;
}
// the columns force2file writes to rNNForceNonDim.csv (libPostprocess.f90:814-837): CL/CT CD/CQ CLu CDi CD0 CDu CFx CFy CFz
void vcase_force_nondim(const vcase* c, int ir, double out[9]) {
  const Rotor& r = c->rotor[ir];
  const double den = r.nonDimforceDenominator, zero[3] = {0, 0, 0};
  const double signLift = sgn1(v_dot(r.lift, r.zAxisBody));
  out[0] = signLift * v_norm(r.lift) / den;
  out[1] = v_norm(r.drag) / den;
  out[2] = v_norm(r.liftUnsteady) / den;
  out[3] = out[4] = out[5] = v_norm(zero) / den;
  out[6] = r.forceInertial[0] / den;
  out[7] = r.forceInertial[1] / den;
  out[8] = r.forceInertial[2] / den;
}
// gamVec of rotor ir (N = nc*ns*nb) and the sectional block of blade ib as vlc_rotor_get_loads returned it last
int vcase_get_gamvec(const vcase* c, int ir, double* out) {
  if (ir < 0 || ir >= c->nr) return 2;
  std::memcpy(out, c->rotor[ir].gamVec.data(), sizeof(double) * c->rotor[ir].gamVec.size());
  return 0;
}
int vcase_get_loads(const vcase* c, int ir, int ib, double* out) {
  if (ir < 0 || ir >= c->nr || ib < 0 || ib >= c->rotor[ir].nb) return 2;
  const auto& l = c->rotor[ir].blade[ib].loads;
  std::memcpy(out, l.data(), sizeof(double) * l.size());
  return 0;
}
}  // extern "C"
