// wake_records.cuh -- the reference's O(N) wake mutators on the DEVICE copies of its own records:
// assignshed, age_wake, dissipate_wake, strain_wake, convectwake (+ wake_continuity, axisymmetric copy/rotate), rollup
// (+ shiftwake / shiftFwake), and the velocity-array bookkeeping of the convection driver, so that a whole time step of
// a case runs without the wake ever leaving the GPU.  Records keep the reference's layout (vr_class = 50 doubles,
// Fwake_class = 13, wingpanel_class = 104; arrays column-major, waN(i, j) at 50*((i-1) + nNwake*(j-1)), blades
// contiguous), so the sweeps' pack kernels read them unchanged.  All arithmetic is written with __dadd_rn/__dmul_rn
// (no FMA contraction) in the reference's statement order: the results are bit-identical to the CPU restatement.
// HBM-bound elementwise maps, a few microseconds each at the shipped case sizes.
#pragma once
#include <cuda_runtime.h>

#include "pack.cuh"
#include "pfwake.cuh"

namespace vlc {

// vf_class (classdef.f90:57-79) as 12 doubles: fc(3,2) | l0 lc rVc0 rVc age ageAzimuthal
constexpr int kVfFc1 = 0, kVfFc2 = 3, kVfL0 = 6, kVfLc = 7, kVfRvc0 = 8, kVfAge = 10, kVfAgeAz = 11;

__device__ __forceinline__ double* ring_at(double* waN, int nNwake, int i, int j) {  // 1-based (i, j)
  return waN + (size_t)kVr * ((size_t)(i - 1) + (size_t)nNwake * (j - 1));
}
__device__ __forceinline__ const double* ring_at(const double* waN, int nNwake, int i, int j) {
  return waN + (size_t)kVr * ((size_t)(i - 1) + (size_t)nNwake * (j - 1));
}
__device__ __forceinline__ void copy3(double* d, const double* s) {
  d[0] = s[0];
  d[1] = s[1];
  d[2] = s[2];
}
// vr_assignP (classdef.f90:569-592): corner n is fc(:,2) of filament n-1 (cyclic) and fc(:,1) of filament n
__device__ __forceinline__ void ring_assignP(double* ring, int n, const double* P) {
  const int a = (n + 2) % 4, b = n - 1;  // 0-based filaments: n=1 -> (3, 0), 2 -> (0, 1), 3 -> (1, 2), 4 -> (2, 3)
  copy3(ring + kVf * a + kVfFc2, P);
  copy3(ring + kVf * b + kVfFc1, P);
}
// intrinsic norm2 of a 3-vector as the CPU restatement evaluates it: sqrt((x*x + y*y) + z*z), unfused
__device__ __forceinline__ double norm2_3(double x, double y, double z) {
  return sqrt(__dadd_rn(__dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y)), __dmul_rn(z, z)));
}
// vf_calclength (classdef.f90:505-515)
__device__ __forceinline__ void vf_calclength(double* vf, bool isOriginal) {
  const double l = norm2_3(vf[kVfFc1 + 0] - vf[kVfFc2 + 0], vf[kVfFc1 + 1] - vf[kVfFc2 + 1], vf[kVfFc1 + 2] - vf[kVfFc2 + 2]);
  vf[kVfLc] = l;
  if (isOriginal) vf[kVfL0] = l;
}

// rotor_assignshed (classdef.f90:4297-4325).  edge 0 = 'LE': corners 1, 4 of wake row rowNear <- corners 2, 3 of the
// wing's trailing-edge panels, lengths recorded as original, gam <- gam of the TE panel.  edge 1 = 'TE': corners 2, 3
// of row max(rowNear-1, 1).  One thread per (blade, spanwise station).
__global__ void rec_assignshed_kernel(int edge, int nb, int nc, int ns, int nNwake, int rowNear, const double* __restrict__ wiP,
                                      double* __restrict__ waN) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= nb * ns) return;
  const int ib = q / ns, is = q % ns + 1;
  const double* wp = wiP + (size_t)kWp * ((size_t)(nc - 1) + (size_t)nc * (is - 1) + (size_t)nc * ns * ib);
  double* blade = waN + (size_t)ib * nNwake * ns * kVr;
  if (edge == 0) {
    double* w = ring_at(blade, nNwake, rowNear, is);
    ring_assignP(w, 1, wp + kVf * 1 + kVfFc1);
    ring_assignP(w, 4, wp + kVf * 2 + kVfFc1);
    for (int f = 0; f < 4; ++f) vf_calclength(w + kVf * f, true);
    w[kVrGam] = wp[kVrGam];
  } else {
    const int row = rowNear - 1 > 1 ? rowNear - 1 : 1;
    double* w = ring_at(blade, nNwake, row, is);
    ring_assignP(w, 2, wp + kVf * 1 + kVfFc1);
    ring_assignP(w, 3, wp + kVf * 2 + kVfFc1);
  }
}

// rotor_age_wake (classdef.f90:4331-4354): age += dt, ageAzimuthal += dt*omegaSlow on rows rowNear..nNwake (4 filaments)
// and far rows rowFar..nFwake.  One thread per (blade, column, active row) and per (blade, far row).
__global__ void rec_age_kernel(int nb, int ns, int nNwake, int nFwake, int rowNear, int rowFar, double dt, double dtOmega,
                               double* __restrict__ waN, double* __restrict__ waF) {
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int nact = nNwake - rowNear + 1, nfar = nFwake - rowFar + 1;
  const long long nn = (long long)nb * ns * (nact > 0 ? nact : 0);
  if (q < nn) {
    const int i = rowNear + (int)(q % nact), j = (int)((q / nact) % ns) + 1, ib = (int)(q / ((long long)nact * ns));
    double* w = ring_at(waN + (size_t)ib * nNwake * ns * kVr, nNwake, i, j);
    for (int f = 0; f < 4; ++f) {
      w[kVf * f + kVfAge] = __dadd_rn(w[kVf * f + kVfAge], dt);
      w[kVf * f + kVfAgeAz] = __dadd_rn(w[kVf * f + kVfAgeAz], dtOmega);
    }
    return;
  }
  const long long p = q - nn;
  if (nfar > 0 && p < (long long)nb * nfar) {
    const int i = rowFar + (int)(p % nfar), ib = (int)(p / nfar);
    double* f = waF + (size_t)kFw * ((size_t)(i - 1) + (size_t)nFwake * ib);
    f[kVfAge] = __dadd_rn(f[kVfAge], dt);
    f[kVfAgeAz] = __dadd_rn(f[kVfAgeAz], dtOmega);
  }
}

// rotor_dissipate_wake (classdef.f90:4356-4408), pass 1: vf(1)%rVc grows, vf(3)%rVc <- vf(1)%rVc (quirk C2), gam decays,
// vf(2)%rVc grows; far rows grow and decay.  growTerm = 4*oseenParameter*apparentViscCoeff*nu*dt and
// decayFactor = exp(-decayCoeff*dt) are formed on the host exactly as the reference writes them.
__global__ void rec_dissipate_kernel(int nb, int ns, int nNwake, int nFwake, int rowNear, int rowFar, double growTerm,
                                     double decayFactor, double* __restrict__ waN, double* __restrict__ waF) {
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int nact = nNwake - rowNear + 1, nfar = nFwake - rowFar + 1;
  const long long nn = (long long)nb * ns * (nact > 0 ? nact : 0);
  if (q < nn) {
    const int i = rowNear + (int)(q % nact), j = (int)((q / nact) % ns) + 1, ib = (int)(q / ((long long)nact * ns));
    double* w = ring_at(waN + (size_t)ib * nNwake * ns * kVr, nNwake, i, j);
    const double r1 = sqrt(__dadd_rn(__dmul_rn(w[kVfRvc], w[kVfRvc]), growTerm));
    w[kVfRvc] = r1;
    w[kVf * 2 + kVfRvc] = r1;
    w[kVrGam] = __dmul_rn(w[kVrGam], decayFactor);
    const double r2 = w[kVf * 1 + kVfRvc];
    w[kVf * 1 + kVfRvc] = sqrt(__dadd_rn(__dmul_rn(r2, r2), growTerm));
    return;
  }
  const long long p = q - nn;
  if (nfar > 0 && p < (long long)nb * nfar) {
    const int i = rowFar + (int)(p % nfar), ib = (int)(p / nfar);
    double* f = waF + (size_t)kFw * ((size_t)(i - 1) + (size_t)nFwake * ib);
    f[kVfRvc] = sqrt(__dadd_rn(__dmul_rn(f[kVfRvc], f[kVfRvc]), growTerm));
    f[kFwGam] = __dmul_rn(f[kFwGam], decayFactor);
  }
}
// pass 2 (classdef.f90:4386-4392): vf(4)%rVc of row i <- vf(2)%rVc of row i-1, rows rowNear+1..nNwake
__global__ void rec_dissipate_vf4_kernel(int nb, int ns, int nNwake, int rowNear, double* __restrict__ waN) {
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int nact = nNwake - rowNear;  // rows rowNear+1..nNwake
  if (nact <= 0 || q >= (long long)nb * ns * nact) return;
  const int i = rowNear + 1 + (int)(q % nact), j = (int)((q / nact) % ns) + 1, ib = (int)(q / ((long long)nact * ns));
  double* blade = waN + (size_t)ib * nNwake * ns * kVr;
  ring_at(blade, nNwake, i, j)[kVf * 3 + kVfRvc] = ring_at(blade, nNwake, i - 1, j)[kVf * 1 + kVfRvc];
}

// rotor_strain_wake (classdef.f90:4410-4422; vf_calclength :505-515, vf_strain :517-521): far rows only, rVc recomputed
// from rVc0 (quirk C3).
__global__ void rec_strain_kernel(int nb, int nFwake, int rowFar, double* __restrict__ waF) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  const int nfar = nFwake - rowFar + 1;
  if (nfar <= 0 || q >= nb * nfar) return;
  const int i = rowFar + q % nfar, ib = q / nfar;
  double* f = waF + (size_t)kFw * ((size_t)(i - 1) + (size_t)nFwake * ib);
  vf_calclength(f, false);
  f[kVfRvc] = __dmul_rn(f[kVfRvc0], sqrt(f[kVfL0] / f[kVfLc]));
}

// blade_convectwake (classdef.f90:1515-1575), the shifts: corner 2 of ring (i, j) by velNwake(:, i, j)*dt for j <= ns,
// corner 3 of ring (i, ns) by velNwake(:, i, ns+1)*dt, far end point fc(:,1) by velFwake(:, i)*dt.  'C' moves rows
// rowNear..nNwake; 'P' has the reference's loop `do i = 1, rowNear, nNwake` (quirk C1: start 1, end rowNear, stride
// nNwake), restated as written.  vel arrays are (3, nNwake, ns+1) / (3, nFwake) per blade.  Blades 0..nbConvect-1.
__global__ void rec_convect_kernel(int predicted, int nbConvect, int ns, int nNwake, int nFwake, int rowNear, int rowFar,
                                   double dt, const double* __restrict__ velN, const double* __restrict__ velF,
                                   double* __restrict__ waN, double* __restrict__ waF) {
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int nfar = nFwake - rowFar + 1;
  const long long nn = (long long)nbConvect * (ns + 1) * nNwake;
  if (q < nn) {
    const int i = (int)(q % nNwake) + 1, j = (int)((q / nNwake) % (ns + 1)) + 1, ib = (int)(q / ((long long)nNwake * (ns + 1)));
    const bool active = predicted ? (i <= rowNear && (i - 1) % nNwake == 0) : (i >= rowNear);
    if (!active) return;
    const double* v = velN + 3 * ((size_t)(i - 1) + (size_t)nNwake * (j - 1) + (size_t)nNwake * (ns + 1) * ib);
    const double d0 = __dmul_rn(v[0], dt), d1 = __dmul_rn(v[1], dt), d2 = __dmul_rn(v[2], dt);
    double* blade = waN + (size_t)ib * nNwake * ns * kVr;
    // vr_shiftdP (classdef.f90:594-624): corner 2 = vf(1)%fc(:,2), vf(2)%fc(:,1); corner 3 = vf(2)%fc(:,2), vf(3)%fc(:,1)
    double* w = ring_at(blade, nNwake, i, j <= ns ? j : ns);
    double* a = (j <= ns) ? w + kVf * 0 + kVfFc2 : w + kVf * 1 + kVfFc2;
    double* b = (j <= ns) ? w + kVf * 1 + kVfFc1 : w + kVf * 2 + kVfFc1;
    a[0] = __dadd_rn(a[0], d0);
    a[1] = __dadd_rn(a[1], d1);
    a[2] = __dadd_rn(a[2], d2);
    b[0] = __dadd_rn(b[0], d0);
    b[1] = __dadd_rn(b[1], d1);
    b[2] = __dadd_rn(b[2], d2);
    return;
  }
  const long long p = q - nn;
  if (nfar > 0 && p < (long long)nbConvect * nfar) {
    const int i = rowFar + (int)(p % nfar), ib = (int)(p / nfar);
    const double* v = velF + 3 * ((size_t)(i - 1) + (size_t)nFwake * ib);
    double* f = waF + (size_t)kFw * ((size_t)(i - 1) + (size_t)nFwake * ib);
    f[0] = __dadd_rn(f[0], __dmul_rn(v[0], dt));
    f[1] = __dadd_rn(f[1], __dmul_rn(v[1], dt));
    f[2] = __dadd_rn(f[2], __dmul_rn(v[2], dt));
  }
}

// blade_wake_continuity (classdef.f90:1609-1702): re-stitch the shared corners from the convected ones.  Every value
// read here (corner 2 = vf(2)%fc(:,1) of a neighbour; corner 3 = vf(3)%fc(:,1) of the last column) is never written
// by this pass, so one thread per ring reproduces the reference's sequential loops exactly.  The duct closure
// (:1645-1656) reads corner 1 of column 1, which this pass writes: it runs as a second launch (duct != 0).
__global__ void rec_continuity_kernel(int duct, int nbConvect, int ns, int nNwake, int nFwake, int rowNear, int rowFar,
                                      double* __restrict__ waN, double* __restrict__ waF) {
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int nact = nNwake - rowNear + 1, nfar = nFwake - rowFar;  // far rows rowFar+1..nFwake
  const long long nn = (long long)nbConvect * ns * (nact > 0 ? nact : 0);
  if (q < nn) {
    const int i = rowNear + (int)(q % nact), j = (int)((q / nact) % ns) + 1, ib = (int)(q / ((long long)nact * ns));
    double* blade = waN + (size_t)ib * nNwake * ns * kVr;
    double* w = ring_at(blade, nNwake, i, j);
    if (duct) {
      if (j == ns && i > rowNear) {
        const double* first = ring_at(blade, nNwake, i, 1);
        ring_assignP(w, 4, first + kVf * 0 + kVfFc1);
        ring_assignP(w, 3, first + kVf * 0 + kVfFc2);
      }
      return;
    }
    if (j < ns) {
      if (i > rowNear) {
        ring_assignP(w, 1, ring_at(blade, nNwake, i - 1, j) + kVf * 1 + kVfFc1);
        ring_assignP(w, 3, ring_at(blade, nNwake, i, j + 1) + kVf * 1 + kVfFc1);
        ring_assignP(w, 4, ring_at(blade, nNwake, i - 1, j + 1) + kVf * 1 + kVfFc1);
      } else {
        ring_assignP(w, 3, ring_at(blade, nNwake, i, j + 1) + kVf * 1 + kVfFc1);
      }
    } else if (i > rowNear) {
      const double* up = ring_at(blade, nNwake, i - 1, ns);
      ring_assignP(w, 1, up + kVf * 1 + kVfFc1);
      ring_assignP(w, 4, up + kVf * 2 + kVfFc1);
    }
    return;
  }
  const long long p = q - nn;
  if (!duct && nfar > 0 && p < (long long)nbConvect * nfar) {
    const int i = rowFar + 1 + (int)(p % nfar), ib = (int)(p / nfar);
    double* f = waF + (size_t)kFw * ((size_t)(i - 1) + (size_t)nFwake * ib);
    copy3(f + kVfFc2, f - kFw + kVfFc1);
  }
}

// rotor_convectwake, axisymmetric branch (classdef.f90:4801-4823): blade ib >= 2 <- copy of blade 1's active rows, every
// filament end point rotated by Tmat(ib) about the hub (vr_rot :626-642, Fwake rot :969-973): x <- matmul(T, x - o) + o.
// T (column-major 3x3 per blade, identity flag = |theta| <= eps skips the rotation) is formed on the host with
// getTransformAxis (libMath.f90:695-726).
struct AxiT {
  double T[9];
  int rotate;
};
__device__ __forceinline__ void rot_point(const AxiT& t, const double* o, double* x) {
  const double d0 = x[0] - o[0], d1 = x[1] - o[1], d2 = x[2] - o[2];
  for (int r = 0; r < 3; ++r)
    x[r] = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(t.T[r], d0), __dmul_rn(t.T[r + 3], d1)), __dmul_rn(t.T[r + 6], d2)), o[r]);
}
__global__ void rec_axisym_kernel(int nb, int ns, int nNwake, int nFwake, int rowNear, int rowFar, const AxiT* __restrict__ Ts,
                                  double ox, double oy, double oz, double* __restrict__ waN, double* __restrict__ waF) {
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int nact = nNwake - rowNear + 1, nfar = nFwake - rowFar + 1;
  const long long nn = (long long)(nb - 1) * ns * (nact > 0 ? nact : 0);
  const double o[3] = {ox, oy, oz};
  if (q < nn) {
    const int i = rowNear + (int)(q % nact), j = (int)((q / nact) % ns) + 1, ib = 1 + (int)(q / ((long long)nact * ns));
    const double* src = ring_at(waN, nNwake, i, j);
    double* dst = ring_at(waN + (size_t)ib * nNwake * ns * kVr, nNwake, i, j);
    for (int k = 0; k < kVr; ++k) dst[k] = src[k];
    const AxiT t = Ts[ib];
    if (t.rotate)
      for (int f = 0; f < 4; ++f) {
        rot_point(t, o, dst + kVf * f + kVfFc1);
        rot_point(t, o, dst + kVf * f + kVfFc2);
      }
    return;
  }
  const long long p = q - nn;
  if (nfar > 0 && p < (long long)(nb - 1) * nfar) {
    const int i = rowFar + (int)(p % nfar), ib = 1 + (int)(p / nfar);
    const double* src = waF + (size_t)kFw * (i - 1);
    double* dst = waF + (size_t)kFw * ((size_t)(i - 1) + (size_t)nFwake * ib);
    for (int k = 0; k < kFw; ++k) dst[k] = src[k];
    const AxiT t = Ts[ib];
    if (t.rotate) {
      rot_point(t, o, dst + kVfFc1);
      rot_point(t, o, dst + kVfFc2);
    }
  }
}

// rotor_updatePrescribedWake (classdef.f90:5170-5218) in two launches, arithmetic in pfwake.cuh.  helix = (helixPitch,
// helixRadius) per blade of this record set; fits = scratch, one per blade; the far rows rowStart..rowStart+n-1 are fitted.
__global__ void pf_fit_kernel(int nbConvect, int nFwake, int rowStart, int n, double deltaPsi, double hubZ,
                              const double* __restrict__ waF, double* __restrict__ helix, pf::Fit* __restrict__ fits) {
  const int ib = blockIdx.x * blockDim.x + threadIdx.x;
  if (ib >= nbConvect) return;
  pf::fit(waF + (size_t)kFw * ((size_t)(rowStart - 1) + (size_t)nFwake * ib), n, deltaPsi, hubZ, helix + 2 * ib, fits + ib);
}
// Ts: the blade rotations of the axisymmetric branch (as rec_axisym_kernel), may be null when axisym != 1
__global__ void pf_helix_kernel(int nb, int nbConvect, int axisym, const pf::Fit* __restrict__ fits, const AxiT* __restrict__ Ts,
                                double ox, double oy, double oz, double* __restrict__ wapF, double* __restrict__ helix) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= nb * pf::kNpf) return;
  const int ib = q / pf::kNpf, i = q % pf::kNpf;
  const double hub[3] = {ox, oy, oz};
  const bool copy = axisym == 1 && ib > 0;
  pf::blade_filament(ib, i, nbConvect, axisym, fits, copy ? Ts[ib].T : nullptr, copy ? Ts[ib].rotate : 0, hub, wapF, helix);
}

// rotor_burst_wake (classdef.f90:4911-4917, :2306-2339): one thread per (blade, pair of successive far filaments irow,
// irow + 1 with irow = rowFar..nFwake-1).  The source's loop is sequential but order-free: the test reads only end points,
// which nothing here writes, and every write stores the same value, so overlapping pairs may race harmlessly.
__global__ void rec_burst_kernel(int nb, int nFwake, int rowFar, double skewLimit, double largeCoreRadius, double* __restrict__ waF) {
  const int npair = nFwake - rowFar;
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (npair <= 0 || q >= nb * npair) return;
  const int ib = q / npair, irow = rowFar + q % npair;
  double* f0 = waF + (size_t)kFw * ((size_t)(irow - 1) + (size_t)nFwake * ib);
  if (pf::burst_pair(f0, f0 + kFw, skewLimit)) {
    f0[kFw + kVfRvc] = largeCoreRadius;
    f0[kVfRvc] = largeCoreRadius;
  }
}

// rotor_calc_skew (classdef.f90:4919-4936): vr%skew of every active near-wake ring of the convected blades; with
// axisymmetry blades 2..nb receive blade 1's value (computed here from blade 1's ring: the same operands, the same result).
// One thread per (blade, column, active row); reads corners and gam, writes only the skew member (49).
__global__ void rec_skew_kernel(int nb, int nbConvect, int axisym, int ns, int nNwake, int rowNear, double* __restrict__ waN) {
  const int nact = nNwake - rowNear + 1;
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (nact <= 0 || q >= (long long)nb * ns * nact) return;
  const int i = rowNear + (int)(q % nact), j = (int)((q / nact) % ns) + 1, ib = (int)(q / ((long long)nact * ns));
  const int src = pf::source_blade(ib, nbConvect, axisym);
  if (src < 0) return;
  const size_t blade = (size_t)nNwake * ns * kVr;
  ring_at(waN + blade * ib, nNwake, i, j)[kVrGam + 1] = pf::ring_skew(ring_at(waN + blade * src, nNwake, i, j));
}

// rotor_shiftFwake (classdef.f90:4500-4513): waF(i) = waF(i-1), i = nFwake..2, then waF(1)%vf%age = 0.  One thread per
// (blade, double of the record), walking the rows downwards like the reference (nFwake is a few hundred at most).
__global__ void rec_shiftFwake_kernel(int nb, int nFwake, double* __restrict__ waF) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= nb * kFw) return;
  const int k = q % kFw, ib = q / kFw;
  double* f = waF + (size_t)kFw * nFwake * ib + k;
  for (int i = nFwake - 1; i >= 1; --i) f[(size_t)kFw * i] = f[(size_t)kFw * (i - 1)];
  if (k == kVfAge) f[0] = 0.0;
}

// rotor_rollup (classdef.f90:4515-4605) up to its final shiftwake: gam-weighted centroid of corners 4 / 3 and of vf(3)%rVc
// of the last near row over columns rollupStart..rollupEnd, extreme gam (sign rule), written into far row rowFarNext.
// One thread per blade, sums in the reference's order.  sgnPositive = sign(1, Omega*controlPitch(1)) > eps.
__global__ void rec_rollup_kernel(int nb, int ns, int nNwake, int nFwake, int rollupStart, int rollupEnd, int sgnPositive,
                                  int suppressFwake, int rowFarNext, const double* __restrict__ waN, double* __restrict__ waF) {
  const int ib = blockIdx.x * blockDim.x + threadIdx.x;
  if (ib >= nb) return;
  const double* blade = waN + (size_t)ib * nNwake * ns * kVr;
  double gamRollup = ring_at(blade, nNwake, nNwake, ns)[kVrGam];
  double cLE[3] = {0, 0, 0}, cTE[3] = {0, 0, 0}, radiusRollup = 0.0, gamSum = 0.0;
  for (int is = rollupStart; is <= rollupEnd; ++is) {
    const double* w = ring_at(blade, nNwake, nNwake, is);
    const double g = w[kVrGam];
    for (int k = 0; k < 3; ++k) {
      cLE[k] = __dadd_rn(cLE[k], __dmul_rn(w[kVf * 3 + kVfFc1 + k], g));
      cTE[k] = __dadd_rn(cTE[k], __dmul_rn(w[kVf * 2 + kVfFc1 + k], g));
    }
    gamSum = __dadd_rn(gamSum, g);
    if (sgnPositive) {
      if (g < gamRollup) gamRollup = g;
    } else {
      if (g > gamRollup) gamRollup = g;
    }
    radiusRollup = __dadd_rn(radiusRollup, __dmul_rn(w[kVf * 2 + kVfRvc], g));
  }
  const double ageRollup = ring_at(blade, nNwake, nNwake, ns)[kVf * 2 + kVfAge];
  if (fabs(gamSum) > 2.220446049250313e-16) {
    for (int k = 0; k < 3; ++k) {
      cLE[k] = cLE[k] / gamSum;
      cTE[k] = cTE[k] / gamSum;
    }
    radiusRollup = radiusRollup / gamSum;
  } else {
    const double* w = ring_at(blade, nNwake, nNwake, rollupEnd);
    for (int k = 0; k < 3; ++k) {
      cLE[k] = w[kVf * 1 + kVfFc1 + k];
      cTE[k] = w[kVf * 2 + kVfFc1 + k];
    }
    radiusRollup = w[kVf * 2 + kVfRvc];
  }
  if (suppressFwake) gamRollup = 0.0;
  if (nFwake > 0) {
    double* f = waF + (size_t)kFw * ((size_t)(rowFarNext - 1) + (size_t)nFwake * ib);
    for (int k = 0; k < 3; ++k) {
      f[kVfFc2 + k] = cLE[k];
      f[kVfFc1 + k] = cTE[k];
    }
    f[kFwGam] = gamRollup;
    f[kVfAge] = ageRollup;
    f[kVfRvc0] = radiusRollup;
    f[kVfRvc] = radiusRollup;
    vf_calclength(f, true);
    if (rowFarNext < nFwake)
      for (int k = 0; k < 3; ++k) f[kFw + kVfFc2 + k] = cTE[k];
  }
}

// rotor_shiftwake (classdef.f90:4481-4498): waN(i, :) = waN(i-1, :) for i = nNwake..2, ages of row 1 zeroed -- as a copy
// into a second buffer (the caller swaps the two), one thread per double: dst(i) = src(i-1), dst(1) = src(1).
__global__ void rec_shiftwake_kernel(long long nrec_cols, int nNwake, const double* __restrict__ src, double* __restrict__ dst) {
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // over nb*ns columns x nNwake rows x 50 doubles
  if (q >= nrec_cols * nNwake * kVr) return;
  const int k = (int)(q % kVr);
  const int i = (int)((q / kVr) % nNwake);  // 0-based destination row
  double v = src[i > 0 ? q - kVr : q];
  if (i == 0 && (k % kVf) == kVfAge && k < 4 * kVf) v = 0.0;
  dst[q] = v;
}

// ---- velocity arrays of the convection driver: velNwake (3, nNwake, ns+1), velFwake (3, nFwake) per blade ----------

// Targets of one rotor's wake sweep in the order of vind_onNwake_byRotor / vind_onFwake_byRotor (libCommon.f90:133-145,
// :190-195), blades 0..nbConvect-1: per blade [ corner 2 of ring (i, j), i = rowNear..nNwake, j = 1..ns | corner 3 of
// ring (i, ns) | fc(:,1) of far rows rowFar..nFwake ].
__global__ void rec_targets_kernel(int nbConvect, int ns, int nNwake, int nFwake, int rowNear, int rowFar,
                                   const double* __restrict__ waN, const double* __restrict__ waF, double* __restrict__ P) {
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int nact = nNwake - rowNear + 1 > 0 ? nNwake - rowNear + 1 : 0, nfar = nFwake - rowFar + 1 > 0 ? nFwake - rowFar + 1 : 0;
  const long long per = (long long)nact * (ns + 1) + nfar;
  if (q >= per * nbConvect) return;
  const int ib = (int)(q / per);
  const long long t = q % per;
  const double* s;
  if (t < (long long)nact * (ns + 1)) {
    const int i = rowNear + (int)(t % nact), j = (int)(t / nact) + 1;
    const double* w = ring_at(waN + (size_t)ib * nNwake * ns * kVr, nNwake, i, j <= ns ? j : ns);
    s = (j <= ns) ? w + kVf * 1 + kVfFc1 : w + kVf * 2 + kVfFc1;
  } else {
    const int i = rowFar + (int)(t - (long long)nact * (ns + 1));
    s = waF + (size_t)kFw * ((size_t)(i - 1) + (size_t)nFwake * ib) + kVfFc1;
  }
  copy3(P + 3 * q, s);
}

// acc = first ? v : acc + v   (main.f90:818-825: vel = vel + vind_on?wake_byRotor(rotor(jr), ...), jr = 1..nr)
__global__ void rec_accumulate_kernel(long long n, int first, const double* __restrict__ v, double* __restrict__ acc) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) acc[i] = first ? __dadd_rn(0.0, v[i]) : __dadd_rn(acc[i], v[i]);
}

// dst = (c0*s0 + c1*s1 + c2*s2 + c3*s3)/div with the terms added left to right: the multistep velocity formulas of
// fdScheme 4 / 5 (main.f90:1160-1172, :1222-1231, :1309-1325, :1370-1381; signs in the coefficients: a - 16*b is
// a + (-16)*b bit for bit).  dst may be one of the sources: a thread reads its element of every source before it writes.
__global__ void rec_lincomb_kernel(long long n, int nterms, const double* s0, const double* s1, const double* s2,
                                   const double* s3, double c0, double c1, double c2, double c3, double div, double* dst) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double acc = __dmul_rn(c0, s0[i]);
  if (nterms > 1) acc = __dadd_rn(acc, __dmul_rn(c1, s1[i]));
  if (nterms > 2) acc = __dadd_rn(acc, __dmul_rn(c2, s2[i]));
  if (nterms > 3) acc = __dadd_rn(acc, __dmul_rn(c3, s3[i]));
  dst[i] = __ddiv_rn(acc, div);
}

// Scatter the swept velocities into velNwake(:, rowNear:nNwake, :) / velFwake(:, rowFar:nFwake) of every convected
// blade and add the initial wake velocity w = initWakeVel*shaftAxis with the reference's signs (main.f90:829-838,
// :904-911): 'C' near +w, everything else -w; addInit = 0 leaves it out (iter >= initWakeVelNt).
__global__ void rec_scatter_vel_kernel(int nbConvect, int ns, int nNwake, int nFwake, int rowNear, int rowFar, int predicted,
                                       int addInit, double wx, double wy, double wz, const double* __restrict__ V,
                                       double* __restrict__ velN, double* __restrict__ velF) {
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int nact = nNwake - rowNear + 1 > 0 ? nNwake - rowNear + 1 : 0, nfar = nFwake - rowFar + 1 > 0 ? nFwake - rowFar + 1 : 0;
  const long long per = (long long)nact * (ns + 1) + nfar;
  if (q >= per * nbConvect) return;
  const int ib = (int)(q / per);
  const long long t = q % per;
  const double w[3] = {wx, wy, wz};
  double* d;
  bool plus;
  if (t < (long long)nact * (ns + 1)) {
    const int i = rowNear + (int)(t % nact), j = (int)(t / nact) + 1;
    d = velN + 3 * ((size_t)(i - 1) + (size_t)nNwake * (j - 1) + (size_t)nNwake * (ns + 1) * ib);
    plus = !predicted;
  } else {
    const int i = rowFar + (int)(t - (long long)nact * (ns + 1));
    d = velF + 3 * ((size_t)(i - 1) + (size_t)nFwake * ib);
    plus = false;
  }
  for (int k = 0; k < 3; ++k) {
    double v = V[3 * q + k];
    if (addInit) v = plus ? __dadd_rn(v, w[k]) : __dadd_rn(v, -w[k]);
    d[k] = v;
  }
}

// vel_order2_Nwake / vel_order2_Fwake (libCommon.f90:213-258) on the active slice (:, r0:r1, :) of (3, ld, cols) arrays,
// in place into vn (main.f90:927-940): first and last row of the slice (vnp1 + vn)*0.5, inner rows
// (((vnp1(i) + vnp1(i-1)) + vn(i+1)) + vn(i))*0.25 with vn the values BEFORE the update -> out of place via `out`.
__global__ void rec_vel_order2_kernel(int nblk, int ld, int cols, int r0, int rows, const double* __restrict__ vn,
                                      const double* __restrict__ vnp1, double* __restrict__ out) {
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (rows <= 0 || q >= 3LL * rows * cols * nblk) return;
  const int k = (int)(q % 3), i = (int)((q / 3) % rows), j = (int)(q / (3LL * rows));  // j runs over cols*nblk columns
  const size_t a = 3 * ((size_t)(r0 - 1 + i) + (size_t)ld * j) + k;
  if (i == 0 || i == rows - 1)
    out[a] = __dmul_rn(__dadd_rn(vnp1[a], vn[a]), 0.5);
  else
    out[a] = __dmul_rn(__dadd_rn(__dadd_rn(__dadd_rn(vnp1[a], vnp1[a - 3]), vn[a + 3]), vn[a]), 0.25);
}
__global__ void rec_copy_slice_kernel(int nblk, int ld, int cols, int r0, int rows, const double* __restrict__ src,
                                      double* __restrict__ dst) {
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (rows <= 0 || q >= 3LL * rows * cols * nblk) return;
  const int k = (int)(q % 3), i = (int)((q / 3) % rows), j = (int)(q / (3LL * rows));
  const size_t a = 3 * ((size_t)(r0 - 1 + i) + (size_t)ld * j) + k;
  dst[a] = src[a];
}

}  // namespace vlc
