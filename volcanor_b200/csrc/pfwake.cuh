// pfwake.cuh -- the prescribed far wake's generator on the DEVICE copies of the reference's records (its
// last item): pFwake_update (classdef.f90:998-1066) and rotor_updatePrescribedWake (:5170-5218), the statement that ends
// rotor_convectwake when prescWakeNt > 0 and iter > prescWakeNt (:4826-4828).  A helix of 10 revolutions in 240 filaments
// of 15 degrees per blade, fitted to the mean radius and the mean pitch of the far-wake filaments handed in, relaxed
// against the previous fit, attached to the last of them; with axisymmetry blades 2..nb are copies of blade 1 rotated
// about the shaft.  The 240 records per blade are sources of every sweep (classdef.f90:1471-1476; pack_fwake_kernel).
//
// Two tiny kernels in wake_records.cuh (pf_fit_kernel: a fit per convected blade, serial in the reference's summation
// order; pf_helix_kernel: one thread per helix filament):
// latency-bound, a few microseconds, but they keep the far rows and the helix from crossing the bus in each stage.
// Arithmetic: explicit unfused IEEE operations in the reference's statement order; the only calls whose low bits may
// differ from a CPU run are cos / sin / atan2 (CUDA's vs libm's, <= 2 ulp each).  The routines are `VLC_HD` (host +
// device): tests/native/kernels_emul.cpp compiles THIS file and the kernels that call it with g++, where they are
// bit-identical to the oracle (tests/test_kernels_emul.py).
#pragma once

#include "cp_stage.cuh"  // VLC_HD, the unfused mul / add / sub / quo / root

namespace vlc {
namespace pf {

using cp::add;
using cp::mul;
using cp::quo;
using cp::root;
using cp::sub;

constexpr int kFwRec = 13;   // Fwake_class (classdef.f90:198-220): fc(3,2) | l0 lc rVc0 rVc age ageAzimuthal | gam
constexpr int kFc1 = 0, kFc2 = 3, kRvc = 9, kGam = 12;
constexpr int kNpf = 240;    // pFwake_class%waF (classdef.f90:225)
constexpr double kNRevs = 10.0, kRelax = 0.5;  // :227, :230
constexpr double kEps = 2.220446049250313e-16;

// What one blade's helix needs once the fit is done.
struct Fit {
  double pitch, radius, dTheta, deltaZ, anchor[3], gam, rVc;
};

// libMath.f90:9-10: pi = atan(1._dp)*4._dp, folded by the compiler to the double nearest pi; twoPi = 2._dp*pi is exact
VLC_HD double two_pi() { return 2.0 * 3.141592653589793; }

// pFwake_update :1014-1045.  waF = the first of n far-wake records (rows rowStart..nFwakeEnd of one blade), helix[2] =
// this%helixPitch, this%helixRadius (read and updated).  n = 1 divides 0 by 0 like the source.
VLC_HD void fit(const double* waF, int n, double deltaPsi, double hubZ, double* helix, Fit* f) {
  const double* last = waF + (size_t)kFwRec * (n - 1);
  double pitchCur = 0.0, radiusCur = 0.0;
  for (int i = 0; i < n; ++i) {
    const double* w = waF + (size_t)kFwRec * i;
    radiusCur = add(radiusCur, root(add(mul(w[kFc1 + 1], w[kFc1 + 1]), mul(w[kFc1], w[kFc1]))));  // norm2([fc(2,1), fc(1,1)])
    if (i < n - 1) pitchCur = sub(add(pitchCur, w[kFc1 + 2]), w[kFwRec + kFc1 + 2]);
  }
  const double twoPi = two_pi();
  pitchCur = quo(mul(fabs(pitchCur), quo(-twoPi, deltaPsi)), (double)(n - 1));
  radiusCur = quo(radiusCur, (double)n);
  helix[0] = add(mul(kRelax, pitchCur), mul(sub(1.0, kRelax), helix[0]));
  helix[1] = add(mul(kRelax, radiusCur), mul(sub(1.0, kRelax), helix[1]));
  f->pitch = helix[0];
  f->radius = helix[1];
  for (int k = 0; k < 3; ++k) f->anchor[k] = last[kFc1 + k];
  f->dTheta = atan2(f->anchor[1], f->anchor[0]);
  f->deltaZ = sub(f->anchor[2], hubZ);
  f->gam = last[kGam];
  f->rVc = last[kRvc];
}

// this%coords(:, i) for i = 0..240 (:1047-1052): theta = linspace(0, twoPi*nRevs, 241)(i), negated (isClockwiseRotor)
VLC_HD void helix_point(const Fit& f, int i, double* x) {
  const double twoPi = two_pi();
  const double dx = quo(sub(mul(twoPi, kNRevs), 0.0), (double)((kNpf + 1) - 1));  // libMath.f90:152
  double theta = add(mul((double)i, dx), 0.0);
  theta = mul(-1.0, theta);
  const double a = add(theta, f.dTheta);
  x[0] = mul(f.radius, cos(a));
  x[1] = mul(f.radius, sin(a));
  x[2] = add(quo(mul(f.pitch, fabs(theta)), twoPi), f.deltaZ);
}

// Record i (0-based) of the helix (:1055-1065): fc(:,2) = hub + coords(:, i), fc(:,1) = hub + coords(:, i+1), the first
// filament starts at the anchor, gam and rVc of the last far filament; the other members are not touched.
VLC_HD void filament(const Fit& f, int i, const double* hub, double* rec) {
  double a[3], b[3];
  helix_point(f, i, a);
  helix_point(f, i + 1, b);
  for (int k = 0; k < 3; ++k) {
    rec[kFc2 + k] = (i == 0) ? f.anchor[k] : add(hub[k], a[k]);
    rec[kFc1 + k] = add(hub[k], b[k]);
  }
  rec[kGam] = f.gam;
  rec[kRvc] = f.rVc;
}

// Fwake rot (classdef.f90:969-973) = x <- matmul(T, x - o) + o with T column-major
VLC_HD void rot(const double* T, const double* o, double* x) {
  const double d0 = sub(x[0], o[0]), d1 = sub(x[1], o[1]), d2 = sub(x[2], o[2]);
  for (int r = 0; r < 3; ++r) x[r] = add(add(add(mul(T[r], d0), mul(T[r + 3], d1)), mul(T[r + 6], d2)), o[r]);
}

// Which blade's fit record i of blade ib is built from (:5186-5216): its own when convected, blade 1's for the
// axisymmetric copies, -1 = not touched.
VLC_HD int source_blade(int ib, int nbConvect, int axisym) {
  if (axisym == 1 && ib > 0) return 0;
  return ib < nbConvect ? ib : -1;
}

// One (blade, filament) of rotor_updatePrescribedWake after the fits: the blade's own helix, or (axisymmetry) blade 1's
// record -- all 13 members, `wapF(ib) = wapF(1)` is a whole-object copy -- with both end points rotated by T(ib) when
// |bladeOffset| > eps (pFwake_rot_wake_axis :1068-1086).  T = 9 doubles per blade, rotate = its flag.
VLC_HD void blade_filament(int ib, int i, int nbConvect, int axisym, const Fit* fits, const double* T, int rotate,
                           const double* hub, double* wapF, double* helix) {
  const int src = source_blade(ib, nbConvect, axisym);
  if (src < 0) return;
  double* rec = wapF + (size_t)kFwRec * ((size_t)i + (size_t)kNpf * ib);
  filament(fits[src], i, hub, rec);
  if (src == ib) return;
  const double* first = wapF + (size_t)kFwRec * ((size_t)i + (size_t)kNpf * src);
  rec[6] = first[6];    // l0, lc, rVc0, age, ageAzimuthal: never written by the update, copied with the object
  rec[7] = first[7];
  rec[8] = first[8];
  rec[10] = first[10];
  rec[11] = first[11];
  if (rotate) {
    rot(T, hub, rec + kFc1);
    rot(T, hub, rec + kFc2);
  }
  if (i == 0) {
    helix[2 * ib] = fits[src].pitch;
    helix[2 * ib + 1] = fits[src].radius;
  }
}

// ---- wake burst (classdef.f90:4911-4917 rotor_burst_wake -> :2306-2339 blade_burst_wake; far wake only, the near-wake
// branch is commented out in the source).  skewVal = |getAngleCos(fc2(i) - fc1(i), fc1(i+1) - fc2(i+1)) - pi|/pi
// (libMath.f90:238-247): 0 for a straight chain; at or beyond skewLimit both filaments get the core radius
// largeCoreRadius (= rotor%chord).  f0, f1 = records irow, irow + 1.  acos is the device's (<= 2 ulp from libm's): a
// decision can differ from a CPU run only for a kink within that distance of the limit.
VLC_HD bool burst_pair(const double* f0, const double* f1, double skewLimit) {
  const double pi = 3.141592653589793;
  double a[3], b[3];
  for (int k = 0; k < 3; ++k) {
    a[k] = sub(f0[kFc2 + k], f0[kFc1 + k]);
    b[k] = sub(f1[kFc1 + k], f1[kFc2 + k]);
  }
  const double dot = add(add(mul(a[0], b[0]), mul(a[1], b[1])), mul(a[2], b[2]));
  const double sa = add(add(mul(a[0], a[0]), mul(a[1], a[1])), mul(a[2], a[2]));
  const double sb = add(add(mul(b[0], b[0]), mul(b[1], b[1])), mul(b[2], b[2]));
  const double skewVal = quo(fabs(sub(acos(quo(dot, root(mul(sa, sb)))), pi)), pi);
  return skewVal >= skewLimit;
}

// ---- wake skew (classdef.f90:737-747 calc_skew, :704-721 vr_getBimedianCos): |cos| of the angle between the bimedians of a
// wake ring (vr_class record: 4 filaments of 12 doubles, corner n = fc(:,1) of filament n; gam at 48), 0 without circulation.
VLC_HD double ring_skew(const double* ring) {
  if (!(fabs(ring[48]) > kEps)) return 0.0;
  const double *p1 = ring, *p2 = ring + 12, *p3 = ring + 24, *p4 = ring + 36;
  double x1[3], x2[3];
  for (int k = 0; k < 3; ++k) {
    x1[k] = sub(sub(add(p3[k], p4[k]), p1[k]), p2[k]);
    x2[k] = sub(sub(add(p4[k], p1[k]), p2[k]), p3[k]);
  }
  const double d12 = add(add(mul(x1[0], x2[0]), mul(x1[1], x2[1])), mul(x1[2], x2[2]));
  const double d11 = add(add(mul(x1[0], x1[0]), mul(x1[1], x1[1])), mul(x1[2], x1[2]));
  const double d22 = add(add(mul(x2[0], x2[0]), mul(x2[1], x2[1])), mul(x2[2], x2[2]));
  return fabs(quo(d12, root(mul(d11, d22))));
}

}  // namespace pf
}  // namespace vlc
