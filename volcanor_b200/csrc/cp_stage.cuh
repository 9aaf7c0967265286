// cp_stage.cuh -- the collocation-point stage of a time step on the DEVICE copies of the reference's wing records
// (SURVEY 8a10 and 8f rank 3): velCP at the collocation points of a rotor from every rotor's wake and every other
// rotor's wing, RHS = -velCP.nCap (main.f90:522-603), map_gam (classdef.f90:4181-4196), velCPTotal (main.f90:630-656)
// and the circulation-based sectional loads of rotor%calc_secAlpha / rotor%calc_force with forceCalcSwitch = 0
// (classdef.f90:1704-1896, :2197-2265, :2355-2380, :4607-4671).  The N-body part of the stage is the sweep kernels
// (bs_sweep.cuh / bs_lattice.cuh); what is here is the O(nc*ns) bookkeeping around them, so that after the wing has been
// uploaded neither the collocation-point velocities nor the right-hand side cross the bus.
//
// Arithmetic: every operation is written out with explicit, unfused IEEE operations in the reference's statement order,
// so the loads are bit-identical to the CPU restatement given the same velCPTotal (atan2 of secAlpha excepted: libm).
// The per-section routines are `VLC_HD` (host + device): tests/native/cp_stage_host.cpp compiles THIS file with g++ and
// checks it against the oracle without a GPU; the product only ever launches the kernels at the bottom.
#pragma once

#if defined(__CUDACC__)
#include <cuda_runtime.h>
#define VLC_HD __host__ __device__ __forceinline__
#else
#include <cmath>
#define VLC_HD inline
#endif

namespace vlc {
namespace cp {

// wingpanel_class (classdef.f90:106-179) as 104 doubles: vr(50) | gamPrev gamTrapz | PC(3,4) | CP nCap tauCapChord
// tauCapSpan | velCP velCPTotal velCPm | normalForce normalForceUnsteady chordwiseResVel | velPitch delP delPUnsteady
// delDiConstant delDiUnsteady | meanChord meanSpan panelArea rHinge alpha
constexpr int kRec = 104;
constexpr int kGam = 48, kGamPrev = 50, kGamTrapz = 51, kPC1 = 52, kCP = 64, kNcap = 67, kTauChord = 70, kTauSpan = 73;
constexpr int kVelCP = 76, kVelCPTotal = 79, kVelCPm = 82, kNormalForce = 85, kNormalForceUnsteady = 88, kChordwiseResVel = 91;
constexpr int kDelP = 95, kDelPUnsteady = 96, kMeanChord = 99, kMeanSpan = 100, kPanelArea = 101;

// Section frames of one blade as the driver holds them after moving the wing (blade_class, classdef.f90:238-358):
// secTauCapChord(3,ns) | secNormalVec(3,ns) | secCP(3,ns) | secArea(ns) | yAxisAziFlap(3) | zAxisAziFlap(3)
VLC_HD int sec_doubles(int ns) { return 10 * ns + 6; }
// Loads of one blade: forceInertial lift drag liftUnsteady (3 each) | secChordwiseResVel secDragDir secLiftDir
// secForceInertial secLift secDrag secLiftUnsteady (3,ns each) | secAlpha secCL secCD secCLu (ns each)
VLC_HD int loads_doubles(int ns) { return 12 + 25 * ns; }
constexpr int kLdResVel = 0, kLdDragDir = 1, kLdLiftDir = 2, kLdForceInertial = 3, kLdLift = 4, kLdDrag = 5, kLdLiftUnsteady = 6;
constexpr int kLdAlpha = 0, kLdCL = 1, kLdCD = 2, kLdCLu = 3;

constexpr double kEps = 2.220446049250313e-16;  // libMath.f90:11

#if defined(__CUDA_ARCH__)
VLC_HD double mul(double a, double b) { return __dmul_rn(a, b); }
VLC_HD double add(double a, double b) { return __dadd_rn(a, b); }
VLC_HD double sub(double a, double b) { return __dsub_rn(a, b); }
VLC_HD double quo(double a, double b) { return __ddiv_rn(a, b); }
VLC_HD double root(double a) { return __dsqrt_rn(a); }
#else  // host build of the tests: compiled with -ffp-contract=off
VLC_HD double mul(double a, double b) { return a * b; }
VLC_HD double add(double a, double b) { return a + b; }
VLC_HD double sub(double a, double b) { return a - b; }
VLC_HD double quo(double a, double b) { return a / b; }
VLC_HD double root(double a) { return std::sqrt(a); }
#endif

VLC_HD double dot3(const double* a, const double* b) { return add(add(mul(a[0], b[0]), mul(a[1], b[1])), mul(a[2], b[2])); }
VLC_HD double norm3(const double* a) { return root(dot3(a, a)); }
VLC_HD double sign1(double x) { return copysign(1.0, x); }  // sign(1._dp, x)
// libMath.f90:249-262
VLC_HD void unit3(const double* a, double* u) {
  const double n = norm3(a);
  if (n > kEps) {
    u[0] = quo(a[0], n);
    u[1] = quo(a[1], n);
    u[2] = quo(a[2], n);
  } else {
    u[0] = u[1] = u[2] = 0.0;
  }
}
// libMath.f90:202-212
VLC_HD void cross3(const double* a, const double* b, double* c) {
  c[0] = sub(mul(a[1], b[2]), mul(a[2], b[1]));
  c[1] = sub(mul(a[2], b[0]), mul(a[0], b[2]));
  c[2] = sub(mul(a[0], b[1]), mul(a[1], b[0]));
}
// libMath.f90:264-276: component of a along d
VLC_HD void proj3(const double* a, const double* d, double* out) {
  const double nsq = dot3(d, d);
  if (nsq > kEps) {
    const double s = dot3(a, d);
    for (int k = 0; k < 3; ++k) out[k] = quo(mul(s, d[k]), nsq);
  } else {
    out[0] = out[1] = out[2] = 0.0;
  }
}
// libMath.f90:278-291: a minus its component along d
VLC_HD void noproj3(const double* a, const double* d, double* out) {
  const double nsq = dot3(d, d);
  if (nsq > kEps) {
    const double s = dot3(a, d);
    for (int k = 0; k < 3; ++k) out[k] = sub(a[k], quo(mul(s, d[k]), nsq));
  } else {
    out[0] = a[0];
    out[1] = a[1];
    out[2] = a[2];
  }
}

// lsq2 (libMath.f90:577-605): value at xq of the least-squares parabola through (xd, yd); the 3x3 normal equations
// are solved by elimination with partial pivoting from their seven moments (same operation order as the CPU restatement).
VLC_HD double lsq2_from_moments(double xq, int n, double s1, double s2, double s3, double s4, double r1, double r2, double r3) {
  double A[3][4] = {{(double)n, s1, s2, r1}, {s1, s2, s3, r2}, {s2, s3, s4, r3}};
  for (int c = 0; c < 3; ++c) {
    int p = c;
    for (int r = c + 1; r < 3; ++r)
      if (fabs(A[r][c]) > fabs(A[p][c])) p = r;
    if (p != c)
      for (int k = 0; k < 4; ++k) {
        const double t = A[c][k];
        A[c][k] = A[p][k];
        A[p][k] = t;
      }
    for (int r = c + 1; r < 3; ++r) {
      const double f = quo(A[r][c], A[c][c]);
      for (int k = c; k < 4; ++k) A[r][k] = sub(A[r][k], mul(f, A[c][k]));
    }
  }
  double co[3] = {0.0, 0.0, 0.0};
  for (int r = 2; r >= 0; --r) {
    double s = A[r][3];
    for (int k = r + 1; k < 3; ++k) s = sub(s, mul(A[r][k], co[k]));
    co[r] = quo(s, A[r][r]);
  }
  return add(add(co[0], mul(co[1], xq)), mul(mul(co[2], xq), xq));
}

VLC_HD double* panel(double* wiP, int nc, int ic, int is) {  // wiP(ic, is), 1-based, one blade
  return wiP + (size_t)kRec * ((size_t)(ic - 1) + (size_t)nc * (is - 1));
}

// ---- loads of one blade (blade_calc_secChordwiseResVel + secAlpha classdef.f90:2197-2265, blade_dirLiftDrag :2355-2366,
// blade_calc_force :1726-1892) in four phases, so that the device gives every PANEL a thread where the work is per panel
// and every SECTION a thread where the reference adds panels in order:
//   1 loads_panel_resvel   per panel   chordwise resultant velocity of the panel, its abscissa along the chord
//   2 loads_section_dirs   per section moments added in panel order, least-squares value at the section point, alpha,
//                                      drag and lift directions
//   3 loads_panel_forces   per panel   pressures, normal forces, their components along the section's lift direction
//   4 loads_section_sums   per section forces added in panel order, sectional coefficients
// Panel results that a section phase adds travel through a scratch block of kScr doubles per panel.  A phase reads only
// what an EARLIER phase (or nobody) writes: panels and sections are independent within a phase.  Before r03t one thread per
// section walked everything (29 us for a 4 x 26 wing: ~3 000 dependent FP64 instructions in one warp); the arithmetic and
// its order are unchanged (tests/test_cp_stage_host.py: bit for bit against the oracle).
constexpr int kScr = 16;  // x | chordwiseResVel(3) | lift part of normalForce(3) | of normalForceUnsteady(3) | normalForce(3)

struct SecPtr {
  const double *tauChord, *normalVec, *secCP, *yAxisAziFlap, *zAxisAziFlap;
  double secArea;
  double *resVel, *dragDir, *liftDir, *secForceInertial, *secLift, *secDrag, *secLiftUnsteady, *sec1;
};
VLC_HD SecPtr sec_ptr(int ns, int is, const double* sec, double* loads) {
  SecPtr q;
  q.tauChord = sec + 3 * (is - 1);
  q.normalVec = sec + 3 * ns + 3 * (is - 1);
  q.secCP = sec + 6 * ns + 3 * (is - 1);
  q.secArea = sec[9 * ns + (is - 1)];
  q.yAxisAziFlap = sec + 10 * ns;
  q.zAxisAziFlap = sec + 10 * ns + 3;
  double* sec3 = loads + 12;    // (3, ns) blocks
  q.sec1 = loads + 12 + 21 * ns;  // (ns) blocks
  q.resVel = sec3 + 3 * ns * kLdResVel + 3 * (is - 1);
  q.dragDir = sec3 + 3 * ns * kLdDragDir + 3 * (is - 1);
  q.liftDir = sec3 + 3 * ns * kLdLiftDir + 3 * (is - 1);
  q.secForceInertial = sec3 + 3 * ns * kLdForceInertial + 3 * (is - 1);
  q.secLift = sec3 + 3 * ns * kLdLift + 3 * (is - 1);
  q.secDrag = sec3 + 3 * ns * kLdDrag + 3 * (is - 1);
  q.secLiftUnsteady = sec3 + 3 * ns * kLdLiftUnsteady + 3 * (is - 1);
  return q;
}
VLC_HD double* scr_of(double* scr, int nc, int ic, int is) { return scr + (size_t)kScr * ((size_t)(ic - 1) + (size_t)nc * (is - 1)); }

// phase 1 (:2197-2232, wingpanel_calc_chordwiseResVel :917-923)
VLC_HD void loads_panel_resvel(int nc, int ns, int ic, int is, double* wiP, const double* sec, double* scr) {
  double* p = panel(wiP, nc, ic, is);
  const double* PC1 = panel(wiP, nc, 1, is) + kPC1;
  const double* tauChord = sec + 3 * (is - 1);
  double* o = scr_of(scr, nc, ic, is);
  double crv[3];
  noproj3(p + kVelCPTotal, p + kTauSpan, crv);
  for (int i = 0; i < 3; ++i) p[kChordwiseResVel + i] = crv[i];
  const double d[3] = {sub(p[kCP], PC1[0]), sub(p[kCP + 1], PC1[1]), sub(p[kCP + 2], PC1[2])};
  o[0] = dot3(d, tauChord);
  for (int i = 0; i < 3; ++i) o[1 + i] = crv[i];
  (void)ns;
}

// phase 2 (:2197-2265, :2355-2366)
VLC_HD void loads_section_dirs(int nc, int ns, int is, double* wiP, const double* sec, double Omega, double* loads, double* scr) {
  const SecPtr q = sec_ptr(ns, is, sec, loads);
  const double* PC1 = panel(wiP, nc, 1, is) + kPC1;
  double s1 = 0.0, s2 = 0.0, s3 = 0.0, s4 = 0.0, r1[3] = {0, 0, 0}, r2[3] = {0, 0, 0}, r3[3] = {0, 0, 0};
  for (int ic = 1; ic <= nc; ++ic) {
    const double* o = scr_of(scr, nc, ic, is);
    const double x = o[0];
    const double xx = mul(x, x);
    s1 = add(s1, x);
    s2 = add(s2, xx);
    s3 = add(s3, mul(xx, x));
    s4 = add(s4, mul(mul(xx, x), x));
    for (int i = 0; i < 3; ++i) {
      const double y = o[1 + i];
      r1[i] = add(r1[i], y);
      r2[i] = add(r2[i], mul(y, x));
      r3[i] = add(r3[i], mul(y, xx));
    }
  }
  double rv[3];
  if (nc >= 3) {
    const double d[3] = {sub(q.secCP[0], PC1[0]), sub(q.secCP[1], PC1[1]), sub(q.secCP[2], PC1[2])};
    const double xq = dot3(d, q.tauChord);
    for (int i = 0; i < 3; ++i) rv[i] = lsq2_from_moments(xq, nc, s1, s2, s3, s4, r1[i], r2[i], r3[i]);
  } else {
    for (int i = 0; i < 3; ++i) rv[i] = quo(r1[i], (double)nc);
  }
  for (int i = 0; i < 3; ++i) q.resVel[i] = rv[i];
  q.sec1[ns * kLdAlpha + (is - 1)] = atan2(dot3(rv, q.normalVec), dot3(rv, q.tauChord));  // :2250-2252
  double dd[3], c[3], u[3];
  unit3(rv, dd);
  cross3(dd, q.yAxisAziFlap, c);
  unit3(c, u);
  const double sg = sign1(Omega);
  for (int k = 0; k < 3; ++k) {
    q.dragDir[k] = dd[k];
    q.liftDir[k] = mul(sg, u[k]);
  }
}

// phase 3 (:1726-1850)
VLC_HD void loads_panel_forces(int nc, int ns, int ic, int is, double* wiP, double density, double dt, double Omega,
                               int spanwiseLiftSwitch, const double* loads, double* scr) {
  const double* liftDir = loads + 12 + 3 * ns * kLdLiftDir + 3 * (is - 1);
  double* o = scr_of(scr, nc, ic, is);
  const double inv = mul(-1.0, sign1(Omega));  // invertGammaSign :1726
  double* p = panel(wiP, nc, ic, is);
  const double gam = p[kGam];
  const double velTangentialChord = dot3(p + kVelCP, p + kTauChord);
  const double velTangentialSpan = dot3(p + kVelCP, p + kTauSpan);
  const double gamChordPrev = ic > 1 ? panel(wiP, nc, ic - 1, is)[kGam] : 0.0;
  double gamElementChord = ic == 1 ? gam : sub(gam, gamChordPrev);
  double gamElementSpan = is == 1 ? gam : sub(gam, panel(wiP, nc, ic, is - 1)[kGam]);
  gamElementChord = mul(inv, gamElementChord);
  gamElementSpan = mul(inv, gamElementSpan);
  const double gamTrapz = ic > 1 ? mul(mul(inv, 0.5), add(gam, gamChordPrev)) : mul(mul(inv, 0.5), gam);  // :1774-1780
  p[kGamTrapz] = gamTrapz;
  const double delPUnsteady = quo(mul(density, sub(gamTrapz, p[kGamPrev])), dt);  // :1786
  double delP = add(delPUnsteady, quo(mul(mul(density, velTangentialChord), gamElementChord), p[kMeanChord]));  // :1789
  if (spanwiseLiftSwitch != 0) delP = add(delP, quo(mul(mul(density, velTangentialSpan), gamElementSpan), p[kMeanSpan]));
  p[kDelPUnsteady] = delPUnsteady;
  p[kDelP] = delP;
  p[kGamPrev] = gamTrapz;
  double nf[3], nfu[3];
  for (int k = 0; k < 3; ++k) {
    nf[k] = mul(mul(delP, p[kPanelArea]), p[kNcap + k]);            // :1813
    nfu[k] = mul(mul(delPUnsteady, p[kPanelArea]), p[kNcap + k]);   // :1816
    p[kNormalForce + k] = nf[k];
    p[kNormalForceUnsteady + k] = nfu[k];
    o[10 + k] = nf[k];
  }
  proj3(nf, liftDir, o + 4);
  proj3(nfu, liftDir, o + 7);
}

// phase 4 (:1813-1892; the drag terms are zero in the reference)
VLC_HD void loads_section_sums(int nc, int ns, int is, const double* sec, double density, double* loads, double* scr) {
  const SecPtr q = sec_ptr(ns, is, sec, loads);
  double fi[3] = {0.0, 0.0, 0.0}, sl[3] = {0.0, 0.0, 0.0}, slu[3] = {0.0, 0.0, 0.0};
  for (int ic = 1; ic <= nc; ++ic) {
    const double* o = scr_of(scr, nc, ic, is);
    for (int k = 0; k < 3; ++k) fi[k] = add(fi[k], o[10 + k]);
    for (int k = 0; k < 3; ++k) {
      sl[k] = add(sl[k], o[4 + k]);
      slu[k] = add(slu[k], o[7 + k]);
    }
  }
  for (int k = 0; k < 3; ++k) {
    q.secForceInertial[k] = fi[k];
    q.secLift[k] = sl[k];
    q.secDrag[k] = 0.0;
    q.secLiftUnsteady[k] = slu[k];
  }
  const double rv[3] = {q.resVel[0], q.resVel[1], q.resVel[2]};
  const double mag = norm3(rv);
  const double qd = mul(mul(0.5, density), mul(mag, mag));  // getSecDynamicPressure :2058-2069
  double cl = 0.0, cd = 0.0, clu = 0.0;
  if (fabs(qd) > kEps) {
    const double zero3[3] = {0.0, 0.0, 0.0};  // secDrag
    const double s = sign1(dot3(sl, q.zAxisAziFlap));
    const double den = mul(qd, q.secArea);
    cl = quo(mul(norm3(sl), s), den);
    cd = quo(norm3(zero3), den);
    clu = quo(mul(norm3(slu), s), den);
  }
  q.sec1[ns * kLdCL + (is - 1)] = cl;
  q.sec1[ns * kLdCD + (is - 1)] = cd;
  q.sec1[ns * kLdCLu + (is - 1)] = clu;
}

// sumSecToNetForces (classdef.f90:2368-2380): sections added in order is = 1..ns
VLC_HD void blade_sum_loads(int ns, double* loads) {
  const int which[4] = {kLdForceInertial, kLdLift, kLdDrag, kLdLiftUnsteady};
  for (int f = 0; f < 4; ++f) {
    const double* s = loads + 12 + 3 * ns * which[f];
    double acc[3] = {0.0, 0.0, 0.0};
    for (int is = 0; is < ns; ++is)
      for (int k = 0; k < 3; ++k) acc[k] = add(acc[k], s[3 * is + k]);
    for (int k = 0; k < 3; ++k) loads[3 * f + k] = acc[k];
  }
}

// RHS(q + npb*ib) = -1*dot(velCP, nCap) of the convected blades, blade 1's values for the other blades of an
// axisymmetric rotor, -0 for blades that are neither (main.f90:563-603 as restated in the oracle)
VLC_HD double rhs_entry(int i, int npb, int nbConvect, int axisym, const double* wiP) {
  int ib = i / npb;
  const int q = i % npb;
  double v = 0.0;
  if (axisym == 1 && ib >= 1) ib = 0;
  if (ib < nbConvect) {
    const double* p = wiP + (size_t)kRec * ((size_t)q + (size_t)npb * ib);
    v = dot3(p + kVelCP, p + kNcap);
  }
  return mul(-1.0, v);
}

// rotor_map_gam (classdef.f90:4181-4196): which entry of gamVec panel q of blade ib takes (-1: keeps its gam)
VLC_HD int map_gam_source(int ib, int q, int npb, int nbConvect, int axisym) {
  if (axisym == 1 && ib >= 1) return nbConvect >= 1 ? q : -1;
  return ib < nbConvect ? q + npb * ib : -1;
}

}  // namespace cp

#if defined(__CUDACC__)

// CP of every panel of the convected blades (blades are contiguous, npb panels each) -> targets (3, m)
__global__ void cp_targets_kernel(long long m, const double* __restrict__ wiP, double* __restrict__ P) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 3 * m) return;
  P[i] = wiP[(size_t)cp::kRec * (i / 3) + cp::kCP + (i % 3)];
}

// field(:, panel) = field + V (sign > 0) or field - V (sign < 0), one source rotor at a time in the driver's order
// (main.f90:551-560: velCP = velCP + vind; :639-652: velCPTotal = velCPTotal -/+ vind)
__global__ void cp_accumulate_kernel(long long m, int field, int sign, const double* __restrict__ V, double* __restrict__ wiP) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 3 * m) return;
  double* x = wiP + (size_t)cp::kRec * (i / 3) + field + (i % 3);
  *x = sign > 0 ? __dadd_rn(*x, V[i]) : __dsub_rn(*x, V[i]);
}

// dst field <- src field of the same panel, panels [0, m) (velCPTotal = velCP, main.f90:634)
__global__ void cp_copy_field_kernel(long long m, int src, int dst, double* __restrict__ wiP) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 3 * m) return;
  double* p = wiP + (size_t)cp::kRec * (i / 3);
  p[dst + (i % 3)] = p[src + (i % 3)];
}

// blades 2..nb of an axisymmetric rotor take blade 1's field (main.f90:658-663; n doubles starting at `field`)
__global__ void cp_axisym_field_kernel(int nb, int npb, int field, int n, double* __restrict__ wiP) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)(nb - 1) * npb * n) return;
  const int k = (int)(i % n), q = (int)((i / n) % npb), ib = 1 + (int)(i / ((long long)n * npb));
  wiP[(size_t)cp::kRec * ((size_t)q + (size_t)npb * ib) + field + k] = wiP[(size_t)cp::kRec * q + field + k];
}

__global__ void cp_rhs_kernel(int N, int npb, int nbConvect, int axisym, const double* __restrict__ wiP, double* __restrict__ RHS) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) RHS[i] = cp::rhs_entry(i, npb, nbConvect, axisym, wiP);
}

__global__ void cp_map_gam_kernel(int nb, int npb, int nbConvect, int axisym, const double* __restrict__ gamVec, double* __restrict__ wiP) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nb * npb) return;
  const int src = cp::map_gam_source(i / npb, i % npb, npb, nbConvect, axisym);
  if (src >= 0) wiP[(size_t)cp::kRec * i + cp::kGam] = gamVec[src];
}

// One CTA per convected blade; the four phases of the loads (cp::loads_*) with the CTA's threads over panels, sections,
// panels, sections (strided), a barrier between phases (the scratch block and the loads block are global memory: writes of
// a phase are visible to the CTA after __syncthreads); thread 0 then adds the sections in order.
__global__ void cp_loads_kernel(int nc, int ns, double density, double dt, double Omega, int spanwiseLiftSwitch,
                                double* __restrict__ wiP, const double* __restrict__ sec, double* __restrict__ loads,
                                double* __restrict__ scratch) {
  const int ib = blockIdx.x, np = nc * ns;
  double* w = wiP + (size_t)cp::kRec * np * ib;
  const double* s = sec + (size_t)cp::sec_doubles(ns) * ib;
  double* l = loads + (size_t)cp::loads_doubles(ns) * ib;
  double* scr = scratch + (size_t)cp::kScr * np * ib;
  for (int q = threadIdx.x; q < np; q += blockDim.x) cp::loads_panel_resvel(nc, ns, q % nc + 1, q / nc + 1, w, s, scr);
  __syncthreads();
  for (int is = 1 + threadIdx.x; is <= ns; is += blockDim.x) cp::loads_section_dirs(nc, ns, is, w, s, Omega, l, scr);
  __syncthreads();
  for (int q = threadIdx.x; q < np; q += blockDim.x)
    cp::loads_panel_forces(nc, ns, q % nc + 1, q / nc + 1, w, density, dt, Omega, spanwiseLiftSwitch, l, scr);
  __syncthreads();
  for (int is = 1 + threadIdx.x; is <= ns; is += blockDim.x) cp::loads_section_sums(nc, ns, is, s, density, l, scr);
  __syncthreads();
  if (threadIdx.x == 0) cp::blade_sum_loads(ns, l);
}

// The copies of rotor_calc_secAlpha / rotor_dirLiftDrag / rotor_calc_force for an axisymmetric rotor (classdef.f90:4766-4784,
// :4623-4650): blades 2..nb take blade 1's loads block -- all of it but secChordwiseResVel, which the reference leaves alone
__global__ void cp_axisym_loads_kernel(int nb, int ns, int nld, double* __restrict__ loads) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)(nb - 1) * nld) return;
  const int k = (int)(i % nld);
  if (k >= 12 + 3 * ns * cp::kLdResVel && k < 12 + 3 * ns * (cp::kLdResVel + 1)) return;
  loads[(size_t)nld + i] = loads[k];
}

#endif  // __CUDACC__

}  // namespace vlc
