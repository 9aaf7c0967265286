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

// One spanwise section `is` (1-based) of one blade: blade_calc_secChordwiseResVel + secAlpha (classdef.f90:2197-2265),
// blade_dirLiftDrag (:2355-2366), the section's share of blade_calc_force (:1726-1892).  Reads gam of section is-1
// (never written here), writes only records of section `is` and slot `is` of the loads block: sections are independent.
VLC_HD void section_loads(int nc, int ns, int is, double* wiP, const double* sec, double density, double dt, double Omega,
                          int spanwiseLiftSwitch, double* loads) {
  const double* tauChord = sec + 3 * (is - 1);
  const double* normalVec = sec + 3 * ns + 3 * (is - 1);
  const double* secCP = sec + 6 * ns + 3 * (is - 1);
  const double secArea = sec[9 * ns + (is - 1)];
  const double* yAxisAziFlap = sec + 10 * ns;
  const double* zAxisAziFlap = sec + 10 * ns + 3;
  double* sec3 = loads + 12;            // (3, ns) blocks
  double* sec1 = loads + 12 + 21 * ns;  // (ns) blocks
  double* resVel = sec3 + 3 * ns * kLdResVel + 3 * (is - 1);
  double* dragDir = sec3 + 3 * ns * kLdDragDir + 3 * (is - 1);
  double* liftDir = sec3 + 3 * ns * kLdLiftDir + 3 * (is - 1);
  double* secForceInertial = sec3 + 3 * ns * kLdForceInertial + 3 * (is - 1);
  double* secLift = sec3 + 3 * ns * kLdLift + 3 * (is - 1);
  double* secDrag = sec3 + 3 * ns * kLdDrag + 3 * (is - 1);
  double* secLiftUnsteady = sec3 + 3 * ns * kLdLiftUnsteady + 3 * (is - 1);

  // ---- chordwise resultant velocity of the section (:2197-2232)
  const double* PC1 = panel(wiP, nc, 1, is) + kPC1;
  double s1 = 0.0, s2 = 0.0, s3 = 0.0, s4 = 0.0, r1[3] = {0, 0, 0}, r2[3] = {0, 0, 0}, r3[3] = {0, 0, 0};
  // Values that are written to the records or the loads block and used again are kept in locals (rv, dd, ld, fi, sl, slu,
  // nf, nfu below): the same numbers without reading them back from global memory.  (The kernel stays at ~29 us for a
  // 4 x 26 wing, r03n: ONE warp walking ~3 000 dependent FP64 instructions -- IEEE divisions, atan2, square roots -- per
  // section; only spreading a section over several threads would shorten it.)
  for (int ic = 1; ic <= nc; ++ic) {
    double* p = panel(wiP, nc, ic, is);
    double crv[3];
    noproj3(p + kVelCPTotal, p + kTauSpan, crv);  // wingpanel_calc_chordwiseResVel :917-923
    for (int i = 0; i < 3; ++i) p[kChordwiseResVel + i] = crv[i];
    const double d[3] = {sub(p[kCP], PC1[0]), sub(p[kCP + 1], PC1[1]), sub(p[kCP + 2], PC1[2])};
    const double x = dot3(d, tauChord);
    const double xx = mul(x, x);
    s1 = add(s1, x);
    s2 = add(s2, xx);
    s3 = add(s3, mul(xx, x));
    s4 = add(s4, mul(mul(xx, x), x));
    for (int i = 0; i < 3; ++i) {
      const double y = crv[i];
      r1[i] = add(r1[i], y);
      r2[i] = add(r2[i], mul(y, x));
      r3[i] = add(r3[i], mul(y, xx));
    }
  }
  double rv[3];
  if (nc >= 3) {
    const double d[3] = {sub(secCP[0], PC1[0]), sub(secCP[1], PC1[1]), sub(secCP[2], PC1[2])};
    const double xq = dot3(d, tauChord);
    for (int i = 0; i < 3; ++i) rv[i] = lsq2_from_moments(xq, nc, s1, s2, s3, s4, r1[i], r2[i], r3[i]);
  } else {
    for (int i = 0; i < 3; ++i) rv[i] = quo(r1[i], (double)nc);
  }
  for (int i = 0; i < 3; ++i) resVel[i] = rv[i];
  sec1[ns * kLdAlpha + (is - 1)] = atan2(dot3(rv, normalVec), dot3(rv, tauChord));  // :2250-2252

  // ---- lift and drag directions (:2355-2366)
  double dd[3], ld[3];
  {
    double c[3], u[3];
    unit3(rv, dd);
    cross3(dd, yAxisAziFlap, c);
    unit3(c, u);
    const double sg = sign1(Omega);
    for (int k = 0; k < 3; ++k) ld[k] = mul(sg, u[k]);
    for (int k = 0; k < 3; ++k) {
      dragDir[k] = dd[k];
      liftDir[k] = ld[k];
    }
  }

  // ---- panel pressures and forces of the section (:1726-1850)
  const double inv = mul(-1.0, sign1(Omega));  // invertGammaSign :1726
  double fi[3] = {0.0, 0.0, 0.0}, sl[3] = {0.0, 0.0, 0.0}, slu[3] = {0.0, 0.0, 0.0};
  for (int k = 0; k < 3; ++k) secDrag[k] = 0.0;
  for (int ic = 1; ic <= nc; ++ic) {
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
    double pl[3], plu[3], nf[3], nfu[3];
    for (int k = 0; k < 3; ++k) {
      nf[k] = mul(mul(delP, p[kPanelArea]), p[kNcap + k]);            // :1813
      nfu[k] = mul(mul(delPUnsteady, p[kPanelArea]), p[kNcap + k]);   // :1816
      p[kNormalForce + k] = nf[k];
      p[kNormalForceUnsteady + k] = nfu[k];
      fi[k] = add(fi[k], nf[k]);
    }
    proj3(nf, ld, pl);
    proj3(nfu, ld, plu);
    for (int k = 0; k < 3; ++k) {
      sl[k] = add(sl[k], pl[k]);
      slu[k] = add(slu[k], plu[k]);
    }
  }
  for (int k = 0; k < 3; ++k) {
    secForceInertial[k] = fi[k];
    secLift[k] = sl[k];
    secLiftUnsteady[k] = slu[k];
  }

  // ---- sectional coefficients (:1861-1892; the drag terms are zero in the reference)
  {
    const double mag = norm3(rv);
    const double q = mul(mul(0.5, density), mul(mag, mag));  // getSecDynamicPressure :2058-2069
    double cl = 0.0, cd = 0.0, clu = 0.0;
    if (fabs(q) > kEps) {
      const double zero3[3] = {0.0, 0.0, 0.0};  // secDrag: zero in the reference
      const double s = sign1(dot3(sl, zAxisAziFlap));
      const double den = mul(q, secArea);
      cl = quo(mul(norm3(sl), s), den);
      cd = quo(norm3(zero3), den);
      clu = quo(mul(norm3(slu), s), den);
    }
    sec1[ns * kLdCL + (is - 1)] = cl;
    sec1[ns * kLdCD + (is - 1)] = cd;
    sec1[ns * kLdCLu + (is - 1)] = clu;
  }
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

// One CTA per convected blade, one thread per spanwise section (strided); thread 0 then adds the sections in order.
__global__ void cp_loads_kernel(int nc, int ns, double density, double dt, double Omega, int spanwiseLiftSwitch,
                                double* __restrict__ wiP, const double* __restrict__ sec, double* __restrict__ loads) {
  const int ib = blockIdx.x;
  double* w = wiP + (size_t)cp::kRec * nc * ns * ib;
  const double* s = sec + (size_t)cp::sec_doubles(ns) * ib;
  double* l = loads + (size_t)cp::loads_doubles(ns) * ib;
  for (int is = 1 + threadIdx.x; is <= ns; is += blockDim.x)
    cp::section_loads(nc, ns, is, w, s, density, dt, Omega, spanwiseLiftSwitch, l);
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
