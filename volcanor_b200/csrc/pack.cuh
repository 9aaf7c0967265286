// pack.cuh -- device kernels that turn the caller's filament descriptions into packed 96-byte
// source records (vlc_device.cuh: p1, p2, r0, L2, K, G), in the REFERENCE's enumeration order
// (src/classdef.f90:1342-1357 wing rings; :1437-1513 wake rings, horseshoe correction, far and
// prescribed filaments), and the AIC assembly kernel (classdef.f90:4151-4176).
#pragma once
#include "vlc_device.cuh"

namespace vlc {

#define VLC_INV4PI 0.07957747154594767 /* 0.25/pi, classdef.f90:13 */

// Record layout offsets (doubles) of the reference derived types, see include/volcanor_b200.h.
constexpr int kVf = 12, kVr = 50, kFw = 13, kWp = 104;
constexpr int kVfRvc = 9;      // vf%rVc
constexpr int kVrGam = 48;     // vr%gam
constexpr int kFwGam = 12;     // Fwake%gam
constexpr int kWpCP = 64;      // wingpanel%CP   (vr 50 + gamPrev,gamTrapz 2 + PC 12)
constexpr int kWpNcap = 67;    // wingpanel%nCap

__device__ __forceinline__ void write_rec(double* __restrict__ rec, double p1x, double p1y, double p1z, double p2x,
                                          double p2y, double p2z, double rvc, double G) {
  const double r0x = p2x - p1x, r0y = p2y - p1y, r0z = p2z - p1z;
  const double L2 = fma(r0z, r0z, fma(r0y, r0y, r0x * r0x));
  const double q = rvc * rvc * L2;  // (rVc*|r0|)^2 ; reference: (rVc*norm2(r0))**4 (classdef.f90:501)
  double2* o = reinterpret_cast<double2*>(rec);
  o[0] = make_double2(p1x, p1y);
  o[1] = make_double2(p1z, p2x);
  o[2] = make_double2(p2y, p2z);
  o[3] = make_double2(G * r0x, G * r0y);
  o[4] = make_double2(G * r0z, G * L2);
  o[5] = make_double2(q * q, G);
}

__device__ __forceinline__ void write_null_rec(double* __restrict__ rec) {
  double2* o = reinterpret_cast<double2*>(rec);
  const double2 z = make_double2(0.0, 0.0);
#pragma unroll
  for (int k = 0; k < 6; ++k) o[k] = z;
}

// gam -> G with the reference's wake rule `abs(gam) > eps` (classdef.f90:1452) when wake != 0.
__device__ __forceinline__ double strength(double gam, bool wake) {
  if (wake && !(fabs(gam) > 2.220446049250313e-16)) return 0.0;
  return gam * VLC_INV4PI;
}

// Flat filament arrays (tier 1). Records [n, n_pad) become null filaments.
__global__ void pack_flat_kernel(long long n, long long n_pad, const double* __restrict__ p1,
                                 const double* __restrict__ p2, const double* __restrict__ rvc,
                                 const double* __restrict__ gam, const unsigned char* __restrict__ flag,
                                 double* __restrict__ rec) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_pad) return;
  double* r = rec + i * kSrcDoubles;
  if (i >= n) {
    write_null_rec(r);
    return;
  }
  const bool wake = flag ? (flag[i] != 0) : false;
  write_rec(r, p1[3 * i], p1[3 * i + 1], p1[3 * i + 2], p2[3 * i], p2[3 * i + 1], p2[3 * i + 2], rvc[i],
            strength(gam[i], wake));
}

__global__ void pack_null_kernel(long long count, double* __restrict__ rec) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) write_null_rec(rec + i * kSrcDoubles);
}

// Filaments of vortex rings stored as reference records.
//   ring(i, j) at base + stride*(i + ld*j) doubles, i in [i0, i0+ni), j in [0, nj); 4 filaments each.
//   Output order: j outer, i inner, filament innermost  (classdef.f90:1350-1355, :1450-1456).
//   fil_mask selects filaments (bit f), sign scales gam.
//   blockIdx.y = blade: source arrays and record blocks of the blades of a rotor are equally shaped and equally spaced
//   (src_blade doubles, dst_blade records apart), so one launch packs them all (gridDim.y = 1 and 0, 0 otherwise).
__device__ __forceinline__ void pack_rings_body(const double* __restrict__ base, int stride, int ld, int i0, int ni, int nfil,
                                                int fil_mask, double sign, int wake, double* __restrict__ rec, long long q) {
  const int fsel = (int)(q % nfil);
  const long long ring = q / nfil;
  const int i = (int)(ring % ni) + i0;
  const int j = (int)(ring / ni);
  // fsel-th set bit of fil_mask
  int f = 0, seen = -1;
  for (int b = 0; b < 4; ++b)
    if (fil_mask & (1 << b)) {
      ++seen;
      if (seen == fsel) f = b;
    }
  const double* vr = base + (size_t)stride * ((size_t)i + (size_t)ld * j);
  const double* vf = vr + kVf * f;
  write_rec(rec + q * kSrcDoubles, vf[0], vf[1], vf[2], vf[3], vf[4], vf[5], vf[kVfRvc],
            strength(sign * vr[kVrGam], wake != 0));
}

__global__ void pack_rings_kernel(const double* __restrict__ base, int stride, int ld, int i0, int ni, int nj,
                                  int fil_mask, int nfil, double sign, int wake, double* __restrict__ rec,
                                  long long src_blade = 0, long long dst_blade = 0) {
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)ni * nj * nfil;
  if (q >= total) return;
  base += (size_t)blockIdx.y * src_blade;
  rec += (size_t)blockIdx.y * dst_blade * kSrcDoubles;
  pack_rings_body(base, stride, ld, i0, ni, nfil, fil_mask, sign, wake, rec, q);
}

// Far-wake / prescribed-wake filaments stored as Fwake_class records (13 doubles), i in [i0, i0+ni).
__device__ __forceinline__ void pack_fwake_body(const double* __restrict__ base, int i0, double* __restrict__ rec, long long q) {
  const double* fw = base + (size_t)kFw * (i0 + q);
  write_rec(rec + (size_t)q * kSrcDoubles, fw[0], fw[1], fw[2], fw[3], fw[4], fw[5], fw[kVfRvc],
            strength(fw[kFwGam], true));
}

__global__ void pack_fwake_kernel(const double* __restrict__ base, int i0, int ni, double* __restrict__ rec,
                                  long long src_blade = 0, long long dst_blade = 0) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= ni) return;
  base += (size_t)blockIdx.y * src_blade;
  rec += (size_t)blockIdx.y * dst_blade * kSrcDoubles;
  pack_fwake_body(base, i0, rec, q);
}

// AIC(row, col) = vr(col)%vind(CP(row)) . nCap(row)   classdef.f90:4159-4176 (serial 6-deep loop there).
// One thread per entry; the 4 filaments of a ring are summed in order like vr_vind (:537-540).
// wiP: all blades' wingpanel records, blade-major: panel (ic, is) of blade ib at (ib*nc*ns + ic + nc*is)*104.
__global__ void aic_assemble_kernel(const double* __restrict__ wiP, int N, double* __restrict__ A) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  const int col = blockIdx.y;
  if (row >= N) return;
  const double* pr = wiP + (size_t)kWp * row;
  const double* pc = wiP + (size_t)kWp * col;
  const double px = pr[kWpCP], py = pr[kWpCP + 1], pz = pr[kWpCP + 2];
  double vx = 0.0, vy = 0.0, vz = 0.0;
#pragma unroll
  for (int f = 0; f < 4; ++f) {
    const double* vf = pc + kVf * f;
    Src s;
    s.p1x = vf[0]; s.p1y = vf[1]; s.p1z = vf[2]; s.p2x = vf[3]; s.p2y = vf[4]; s.p2z = vf[5];
    const double r0x = s.p2x - s.p1x, r0y = s.p2y - s.p1y, r0z = s.p2z - s.p1z;
    const double L2 = fma(r0z, r0z, fma(r0y, r0y, r0x * r0x));
    const double q = vf[kVfRvc] * vf[kVfRvc] * L2;
    s.K = q * q;
    s.r0gx = VLC_INV4PI * r0x; s.r0gy = VLC_INV4PI * r0y; s.r0gz = VLC_INV4PI * r0z;
    s.L2g = VLC_INV4PI * L2;
    s.spare = VLC_INV4PI;
    pair_accumulate<false>(s, px, py, pz, vx, vy, vz);
  }
  A[(size_t)row + (size_t)N * col] = vx * pr[kWpNcap] + vy * pr[kWpNcap + 1] + vz * pr[kWpNcap + 2];
}

// ---- tier 3: node-indexed lattice -> packed records ------------------------------------------
// nodes(3, nrows+1, ns+1): node (r, c) at 3*(r + (nrows+1)*c).  Ring (r, j): corners
// 1=(r,j) 2=(r+1,j) 3=(r+1,j+1) 4=(r,j+1); filament k runs corner k -> k+1 (classdef.f90:569-592).
// Output order: j outer, r inner, filament innermost (classdef.f90:1450-1456).
__global__ void pack_lattice_kernel(int nrows, int ns, const double* __restrict__ nodes,
                                    const double* __restrict__ gam, const double* __restrict__ rvc4,
                                    double* __restrict__ rec) {
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)nrows * ns * 4;
  if (q >= total) return;
  const int f = (int)(q & 3);
  const long long ring = q >> 2;
  const int r = (int)(ring % nrows), j = (int)(ring / nrows);
  const int nr1 = nrows + 1;
  // corner (row, col) of filament start / end
  const int cr[5] = {r, r + 1, r + 1, r, r};
  const int cc[5] = {j, j, j + 1, j + 1, j};
  const double* a = nodes + 3 * ((size_t)cr[f] + (size_t)nr1 * cc[f]);
  const double* b = nodes + 3 * ((size_t)cr[f + 1] + (size_t)nr1 * cc[f + 1]);
  write_rec(rec + q * kSrcDoubles, a[0], a[1], a[2], b[0], b[1], b[2], rvc4[q],
            strength(gam[(size_t)r + (size_t)nrows * j], true));
}

// Horseshoe correction: -vf2 of the last near row (classdef.f90:1460-1463), one per column.
__global__ void pack_horseshoe_kernel(int nrows, int ns, const double* __restrict__ nodes,
                                      const double* __restrict__ gam, const double* __restrict__ rvc4,
                                      double* __restrict__ rec) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= ns) return;
  const int nr1 = nrows + 1;
  const size_t ring = (size_t)(nrows - 1) + (size_t)nrows * j;
  const double* a = nodes + 3 * ((size_t)nrows + (size_t)nr1 * j);
  const double* b = nodes + 3 * ((size_t)nrows + (size_t)nr1 * (j + 1));
  write_rec(rec + (size_t)j * kSrcDoubles, a[0], a[1], a[2], b[0], b[1], b[2], rvc4[4 * ring + 1],
            strength(-gam[ring], false));
}

// Far-wake chain: filament i has fc(:,1) = p(:, i+1) (downstream / TE end) and fc(:,2) = p(:, i)
// (classdef.f90:198-211: endpoint 1 is the trailing end, 2 the leading end).
__global__ void pack_chain_kernel(int nfar, const double* __restrict__ p, const double* __restrict__ gamF,
                                  const double* __restrict__ rvcF, double* __restrict__ rec) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nfar) return;
  const double* a = p + 3 * (size_t)(i + 1);
  const double* b = p + 3 * (size_t)i;
  write_rec(rec + (size_t)i * kSrcDoubles, a[0], a[1], a[2], b[0], b[1], b[2], rvcF[i], strength(gamF[i], true));
}


// ---- shared-node form (bs_lattice.cuh): strip records of width W --------------------------------------------
// One lattice of nrows x ns rings = (nrows+1) x (ns+1) nodes.  Strip s covers node columns s*W .. s*W+W; record
// index = strip*(nrows+1) + rr for node row rr = 0..nrows.  Merged edge strengths (Gamma' = Gamma, or 0 under the
// wake rule |Gamma| > eps, classdef.f90:1452; 0 outside the lattice):
//   spanwise   (rr,c)->(rr,c+1): Gamma'(rr-1,c) - Gamma'(rr,c)     [f2 of ring (rr-1,c), reversed f4 of ring (rr,c)]
//   streamwise (rr-1,c)->(rr,c): Gamma'(rr-1,c) - Gamma'(rr-1,c-1) [f1 of ring (rr-1,c), reversed f3 of ring (rr-1,c-1)]
// The streamwise edges of the LAST node column (c = ns) are left to the flat remainder (pack_lastcol / fil_mask 0x4).
// Core radii: the two copies of a shared edge may differ (non-uniform streamwiseCoreVec).  `fmt` says what to
// do about it:
//   kFmtDetect (tier 3, node-indexed lattices): write the merged form and raise *flag (bit 0) on ANY difference -- the
//              sweep then uses the flat records instead;
//   kFmtMerged / kFmtDual (tier 2, after check_rings_kernel has classified the set): write that form, never touch the flag.
//              Dual: the streamwise slot holds r0[3], |r0|^2, gA and the record's extra doubles KA, gB, KB (bs_lattice.cuh).
// Acc supplies node(rr, c, out[3]), gam(r, j) (raw circulation) and rvc(r, j, f).
constexpr int kFmtDetect = -1, kFmtMerged = 0, kFmtDual = 2;

__device__ __forceinline__ void write_edge(double* __restrict__ e, const double* U, const double* V, double g, double rvc) {
  const double x = V[0] - U[0], y = V[1] - U[1], z = V[2] - U[2];
  const double L = fma(z, z, fma(y, y, x * x));
  const double q = rvc * rvc * L;  // (rVc*|r0|)^2 ; reference: (rVc*norm2(r0))**4 (classdef.f90:501)
  e[0] = g * x;
  e[1] = g * y;
  e[2] = g * z;
  e[3] = g * L;
  e[4] = q * q;
}
// the dual form of a streamwise edge: e[0..4] = r0, |r0|^2, gA; x[0..3] = KA, gB, KB, 0
__device__ __forceinline__ void write_edge_dual(double* __restrict__ e, double* __restrict__ x, const double* U, const double* V,
                                                double gA, double rvcA, double gB, double rvcB) {
  const double rx = V[0] - U[0], ry = V[1] - U[1], rz = V[2] - U[2];
  const double L = fma(rz, rz, fma(ry, ry, rx * rx));
  const double qA = rvcA * rvcA * L, qB = rvcB * rvcB * L;
  e[0] = rx;
  e[1] = ry;
  e[2] = rz;
  e[3] = L;
  e[4] = gA;
  x[0] = qA * qA;
  x[1] = gB;
  x[2] = qB * qB;
  x[3] = 0.0;
}

template <int W, class Acc>
__device__ __forceinline__ void fill_strip_record(const Acc& acc, int nrows, int ns, int c0, int rr,
                                                  double* __restrict__ rec, int* __restrict__ unmergeable, int fmt) {
  constexpr int NP = (3 * (W + 1) + 1) / 2 * 2;  // c0 = first ring column of the strip (global column index)
  auto G = [&](int r, int j) -> double {
    return (r >= 0 && r < nrows && j >= 0 && j < ns) ? strength(acc.gam(r, j), true) : 0.0;
  };
  double N[W + 1][3], Np[W][3];
#pragma unroll
  for (int k = 0; k <= W; ++k) {
    const int c = c0 + k;
    if (c <= ns) {
      acc.node(rr, c, N[k]);
    } else {  // past the lattice: a well-separated dummy node; every edge touching it has zero strength
      acc.node(rr, ns, N[k]);
      const double off = (double)(c - ns);
      N[k][0] += off;
      N[k][1] += off;
      N[k][2] += off;
    }
    if (k < W) {
      if (rr >= 1 && c <= ns) {
        acc.node(rr - 1, c, Np[k]);
      } else {
        Np[k][0] = N[k][0];
        Np[k][1] = N[k][1];
        Np[k][2] = N[k][2];
      }
    }
    rec[3 * k] = N[k][0];
    rec[3 * k + 1] = N[k][1];
    rec[3 * k + 2] = N[k][2];
  }
  if (NP > 3 * (W + 1)) rec[NP - 1] = 0.0;
#pragma unroll
  for (int k = 0; k < W; ++k) {
    const int c = c0 + k;
    double* e = rec + NP + 10 * k;
    double* x = rec + NP + 10 * W + 4 * k;
    // spanwise edge (rr, c) -> (rr, c+1), real when c + 1 <= ns
    double gp = 0.0, rvcp = 0.0;
    if (c + 1 <= ns) {
      gp = G(rr - 1, c) - G(rr, c);
      if (rr >= 1) {
        rvcp = acc.rvc(rr - 1, c, 1);
        if (fmt == kFmtDetect && rr < nrows && acc.rvc(rr, c, 3) != rvcp) *unmergeable = 1;
      } else {
        rvcp = acc.rvc(0, c, 3);
      }
    }
    write_edge(e, N[k], N[k + 1], gp, rvcp);
    // streamwise edge (rr-1, c) -> (rr, c), real when rr >= 1 and c <= ns - 1: vf(1) of ring (rr-1, c) and the reversed
    // vf(3) of ring (rr-1, c-1)
    double gA = 0.0, gB = 0.0, rvcA = 0.0, rvcB = 0.0;
    if (rr >= 1 && c <= ns - 1) {
      gA = G(rr - 1, c);
      rvcA = acc.rvc(rr - 1, c, 0);
      rvcB = rvcA;
      if (c >= 1) {
        gB = -G(rr - 1, c - 1);
        rvcB = acc.rvc(rr - 1, c - 1, 2);
        if (fmt == kFmtDetect && rvcB != rvcA) *unmergeable = 1;
      }
    }
    if (fmt == kFmtDual) {
      write_edge_dual(e + 5, x, Np[k], N[k], gA, rvcA, gB, rvcB);
    } else {
      write_edge(e + 5, Np[k], N[k], gA + gB, rvcA);  // = Gamma'(rr-1, c) - Gamma'(rr-1, c-1)
      x[0] = x[1] = x[2] = x[3] = 0.0;
    }
  }
}

// null strip records (padding): distinct finite nodes, zero strengths
template <int W>
__global__ void pack_null_lat_kernel(long long count, double* __restrict__ rec) {
  constexpr int NP = (3 * (W + 1) + 1) / 2 * 2, RD = NP + 14 * W;  // = lat_rec_doubles(W); zero strengths in both forms
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  double* r = rec + i * RD;
  for (int k = 0; k < RD; ++k) r[k] = 0.0;
  for (int k = 0; k <= W; ++k) r[3 * k] = (double)k;  // N_k = (k, 0, 0)
}

// tier 3: node-indexed lattice arrays (include/volcanor_b200.h)
struct LatticeAcc {
  const double* nodes;
  const double* gamv;
  const double* rvc4;
  int nrows, ns;
  __device__ __forceinline__ void node(int rr, int c, double* o) const {
    const double* p = nodes + 3 * ((size_t)rr + (size_t)(nrows + 1) * c);
    o[0] = p[0];
    o[1] = p[1];
    o[2] = p[2];
  }
  __device__ __forceinline__ double gam(int r, int j) const { return gamv[(size_t)r + (size_t)nrows * j]; }
  __device__ __forceinline__ double rvc(int r, int j, int f) const { return rvc4[4 * ((size_t)r + (size_t)nrows * j) + f]; }
};

template <int W>
__global__ void pack_lattice_shared_kernel(int nrows, int ns, const double* __restrict__ nodes,
                                           const double* __restrict__ gam, const double* __restrict__ rvc4,
                                           double* __restrict__ rec, int* __restrict__ unmergeable) {
  constexpr int RD = (3 * (W + 1) + 1) / 2 * 2 + 14 * W;
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int nr1 = nrows + 1, nstrips = (ns + W - 1) / W;
  if (q >= (long long)nstrips * nr1) return;
  const LatticeAcc acc{nodes, gam, rvc4, nrows, ns};
  fill_strip_record<W>(acc, nrows, ns, (int)(q / nr1) * W, (int)(q % nr1), rec + q * RD, unmergeable, kFmtDetect);
}

// Streamwise edges of the last column (f3 of ring (r, ns-1): corner 3 -> corner 4), which no strip covers.
__global__ void pack_lastcol_kernel(int nrows, int ns, const double* __restrict__ nodes,
                                    const double* __restrict__ gam, const double* __restrict__ rvc4,
                                    double* __restrict__ rec) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= nrows) return;
  const int nr1 = nrows + 1;
  const double* a = nodes + 3 * ((size_t)(r + 1) + (size_t)nr1 * ns);
  const double* b = nodes + 3 * ((size_t)r + (size_t)nr1 * ns);
  const size_t ring = (size_t)r + (size_t)nrows * (ns - 1);
  write_rec(rec + (size_t)r * kSrcDoubles, a[0], a[1], a[2], b[0], b[1], b[2], rvc4[4 * ring + 2],
            strength(gam[ring], true));
}


// ---- tier 2, shared-node form from reference records (waN slices of vr_class records) ----------------------
// ring(r, j) at base + stride*((i0 + r) + ld*j), r in [0, nrows), j in [0, ns).
__device__ __forceinline__ const double* ring_ptr(const double* base, int stride, int ld, int i0, int r, int j) {
  return base + (size_t)stride * ((size_t)(i0 + r) + (size_t)ld * j);
}
// Two copies of a lattice corner: equal up to 16 ulp of the largest coordinate.  The reference's own records are
// bitwise equal where blade_wake_continuity copied them, but the axisymmetric blades (rotor_convectwake,
// classdef.f90:4801-4823) hold ROTATED copies of blade 1's wake whose newest row meets the blade's own wing-TE
// corners only to rounding (a few ulp).  The strip records use the upstream ring's TE corner as the node.
__device__ __forceinline__ bool same3(const double* a, const double* b) {
  const double m = fmax(fmax(fmax(fabs(a[0]), fabs(a[1])), fabs(a[2])), fmax(fmax(fabs(b[0]), fabs(b[1])), fabs(b[2])));
  const double tol = 16.0 * 2.220446049250313e-16 * m;
  return fabs(a[0] - b[0]) <= tol && fabs(a[1] - b[1]) <= tol && fabs(a[2] - b[2]) <= tol;
}

// The shared-node form needs the records to describe a LATTICE: filament k ends where filament k+1 starts and
// neighbouring rings share their corners (same3: bitwise after blade_wake_continuity, classdef.f90:1609-1702, and
// assignshed, :4297-4325; to rounding for rotated axisymmetric copies).  Anything else raises the flag and the sweep
// uses the flat enumeration.
__global__ void check_rings_kernel(const double* __restrict__ base, int stride, int ld, int i0, int nrows, int ns,
                                   int* __restrict__ unmergeable, long long src_blade = 0) {
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= (long long)nrows * ns) return;
  base += (size_t)blockIdx.y * src_blade;
  const int r = (int)(q % nrows), j = (int)(q / nrows);
  const double* g = ring_ptr(base, stride, ld, i0, r, j);
  bool ok = true;
#pragma unroll
  for (int f = 0; f < 4; ++f) ok = ok && same3(g + kVf * f + 3, g + kVf * ((f + 1) & 3));  // fc(:,2) of f == fc(:,1) of f+1
  if (r >= 1) {
    const double* up = ring_ptr(base, stride, ld, i0, r - 1, j);
    ok = ok && same3(g, up + kVf * 1) && same3(g + kVf * 3, up + kVf * 2);  // corner 1 == up.corner2, corner 4 == up.corner3
  }
  if (j >= 1) {
    const double* lf = ring_ptr(base, stride, ld, i0, r, j - 1);
    ok = ok && same3(g, lf + kVf * 3) && same3(g + kVf * 1, lf + kVf * 2);  // corner 1 == left.corner4, corner 2 == left.corner3
  }
  if (!ok) atomicOr(unmergeable, 1);
  // core radii of the two copies of the shared edges (bitwise): spanwise copies that differ -- vf(4) of this ring against
  // vf(2) of the ring upstream; only transiently, between shiftwake and the next dissipate_wake (classdef.f90:4386-4392) --
  // send the set to the flat enumeration (bit 0); streamwise copies that differ -- vf(1) of this ring against vf(3) of its
  // left neighbour: every interior edge of a wake shed with a non-uniform streamwiseCoreVec -- select the dual form (bit 1)
  if (r >= 1 && g[kVf * 3 + kVfRvc] != ring_ptr(base, stride, ld, i0, r - 1, j)[kVf * 1 + kVfRvc]) atomicOr(unmergeable, 1);
  if (j >= 1 && g[kVfRvc] != ring_ptr(base, stride, ld, i0, r, j - 1)[kVf * 2 + kVfRvc]) atomicOr(unmergeable, 2);
}

// tier 2: reference records (vr_class) of one blade's near wake
struct RingsAcc {
  const double* base;
  int stride, ld, i0, nrows, ns;
  __device__ __forceinline__ const double* ring(int r, int j) const { return ring_ptr(base, stride, ld, i0, r, j); }
  // node (rr, c) = corner 2 of ring (rr-1, c) [corner 3 of ring (rr-1, ns-1) for c = ns], or corner 1 [4] of row 0
  __device__ __forceinline__ void node(int rr, int c, double* o) const {
    const double* p;
    if (c < ns)
      p = (rr >= 1) ? ring(rr - 1, c) + kVf * 1 : ring(0, c);
    else
      p = (rr >= 1) ? ring(rr - 1, ns - 1) + kVf * 2 : ring(0, ns - 1) + kVf * 3;
    o[0] = p[0];
    o[1] = p[1];
    o[2] = p[2];
  }
  __device__ __forceinline__ double gam(int r, int j) const { return ring(r, j)[kVrGam]; }
  __device__ __forceinline__ double rvc(int r, int j, int f) const { return ring(r, j)[kVf * f + kVfRvc]; }
};

// nstrips strips of width W starting at ring column col_base (a lattice may be covered by strips of two widths:
// ns = 4*floor(ns/4) columns of width-4 strips + one tail strip of width ns mod 4, capi.cu: plan_strips).
template <int W>
__global__ void pack_rings_shared_kernel(const double* __restrict__ base, int stride, int ld, int i0, int nrows,
                                         int ns, int col_base, int nstrips, double* __restrict__ rec,
                                         int* __restrict__ unmergeable, long long src_blade = 0) {
  constexpr int RD = (3 * (W + 1) + 1) / 2 * 2 + 14 * W;
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int nr1 = nrows + 1;
  if (q >= (long long)nstrips * nr1) return;
  base += (size_t)blockIdx.y * src_blade;          // blade blockIdx.y: its records follow the previous blade's
  rec += (size_t)blockIdx.y * nstrips * nr1 * RD;
  const RingsAcc acc{base, stride, ld, i0, nrows, ns};
  // the form check_rings_kernel chose for the whole set (launched before this kernel on the same stream); a set that goes
  // to the flat enumeration (odd flag) gets merged-form records nobody reads
  const int fmt = (*unmergeable == kFmtDual) ? kFmtDual : kFmtMerged;
  fill_strip_record<W>(acc, nrows, ns, col_base + (int)(q / nr1) * W, (int)(q % nr1), rec + q * RD, unmergeable, fmt);
}


// ---- one launch for everything a rotor's packed set consists of ------------------------------------------------------
// A packed set is a handful of SEGMENTS (wing rings | padding | wake rings of every blade | horseshoe corrections | far wake
// | prescribed wake | padding; the same again for the flat remainder of the shared-node form; strip records of one or two
// widths | padding).  One kernel per segment made ~14 launches per pack and ~40 per time step of a small case, where a
// launch costs more than the work (K&P: 0.95 ms per step for 0.15 ms of sweeps).  The table below names the segments; one
// thread packs one record, found by its global index.  Segment kinds: 0 ring filaments (pack_rings_body), 1 far-wake
// records (pack_fwake_body), 2 null flat records, 3 strip records of width W (fill_strip_record), 4 null strip records.
struct PackSeg {
  int kind, W;
  const double* src;   // blade 0
  double* dst;         // blade 0
  long long src_blade; // doubles between the blades' source arrays
  long long dst_blade; // records between the blades' destination blocks
  long long count;     // records per blade
  int nb;
  int stride, ld, i0, ni, mask, nfil, wake;  // ring segments
  double sign;
  int nrows, ns, col_base, nstrips;          // strip segments
  int* flag;
};
constexpr int kPackSegs = 20;
__host__ __device__ constexpr int strip_rd(int W) { return (3 * (W + 1) + 1) / 2 * 2 + 14 * W; }  // = lat_rec_doubles(W), bs_lattice.cuh
struct PackTable {
  int n;
  long long start[kPackSegs + 1];  // first global index of every segment; start[n] = total
  PackSeg seg[kPackSegs];
};

template <int W>
__device__ __forceinline__ void pack_strip_seg(const PackSeg& g, const double* src, double* dst, long long idx) {
  constexpr int RD = (3 * (W + 1) + 1) / 2 * 2 + 14 * W;
  const int nr1 = g.nrows + 1;
  const RingsAcc acc{src, g.stride, g.ld, g.i0, g.nrows, g.ns};
  const int fmt = (*g.flag == kFmtDual) ? kFmtDual : kFmtMerged;
  fill_strip_record<W>(acc, g.nrows, g.ns, g.col_base + (int)(idx / nr1) * W, (int)(idx % nr1), dst + idx * RD, g.flag, fmt);
}
template <int W>
__device__ __forceinline__ void pack_null_strip(double* dst, long long idx) {
  constexpr int NP = (3 * (W + 1) + 1) / 2 * 2, RD = NP + 14 * W;
  double* r = dst + idx * RD;
  for (int k = 0; k < RD; ++k) r[k] = 0.0;
  for (int k = 0; k <= W; ++k) r[3 * k] = (double)k;
}

__global__ void pack_table_kernel(const PackTable t) {
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= t.start[t.n]) return;
  int k = 0;
  while (q >= t.start[k + 1]) ++k;
  const PackSeg& g = t.seg[k];
  const long long local = q - t.start[k];
  const int blade = (int)(local / g.count);
  const long long idx = local % g.count;
  const double* src = g.src + (size_t)blade * g.src_blade;
  switch (g.kind) {
    case 0: pack_rings_body(src, g.stride, g.ld, g.i0, g.ni, g.nfil, g.mask, g.sign, g.wake, g.dst + (size_t)blade * g.dst_blade * kSrcDoubles, idx); break;
    case 1: pack_fwake_body(src, g.i0, g.dst + (size_t)blade * g.dst_blade * kSrcDoubles, idx); break;
    case 2: write_null_rec(g.dst + (size_t)idx * kSrcDoubles); break;
    case 3: {
      const long long per = (long long)g.nstrips * (g.nrows + 1);
      switch (g.W) {
        case 1: pack_strip_seg<1>(g, src, g.dst + (size_t)blade * per * strip_rd(1), idx); break;
        case 2: pack_strip_seg<2>(g, src, g.dst + (size_t)blade * per * strip_rd(2), idx); break;
        case 3: pack_strip_seg<3>(g, src, g.dst + (size_t)blade * per * strip_rd(3), idx); break;
        default: pack_strip_seg<4>(g, src, g.dst + (size_t)blade * per * strip_rd(4), idx); break;
      }
    } break;
    default:
      switch (g.W) {
        case 1: pack_null_strip<1>(g.dst, idx); break;
        case 2: pack_null_strip<2>(g.dst, idx); break;
        case 3: pack_null_strip<3>(g.dst, idx); break;
        default: pack_null_strip<4>(g.dst, idx); break;
      }
  }
}

// ---- gridgen (src/gridgen.f90) ---------------------------------------------------------------------------
// vf_class records (12 doubles) with a separate circulation array; no |gam| > eps rule (gridgen.f90:129-135).
__global__ void pack_vf_gam_kernel(long long n, const double* __restrict__ vf, const double* __restrict__ gam,
                                   double* __restrict__ rec) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double* f = vf + (size_t)kVf * i;
  write_rec(rec + i * kSrcDoubles, f[0], f[1], f[2], f[3], f[4], f[5], f[kVfRvc], strength(gam[i], false));
}

// Cell centres of the Cartesian grid exactly as gridgen.f90:62-84 computes them: linspace (libMath.f90:138-157:
// i*dx then + xstart), the 8 corners summed in the file's order, times 0.125.  P: (3, nx-1, ny-1, nz-1).
__global__ void grid_centres_kernel(int nx, int ny, int nz, double x0, double y0, double z0, double dx, double dy,
                                    double dz, double* __restrict__ P) {
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long cx = nx - 1, cy = ny - 1, cz = nz - 1;
  if (q >= cx * cy * cz) return;
  const int ix = (int)(q % cx), iy = (int)((q / cx) % cy), iz = (int)(q / (cx * cy));
  const double xa = __dadd_rn(__dmul_rn((double)ix, dx), x0), xb = __dadd_rn(__dmul_rn((double)(ix + 1), dx), x0);
  const double ya = __dadd_rn(__dmul_rn((double)iy, dy), y0), yb = __dadd_rn(__dmul_rn((double)(iy + 1), dy), y0);
  const double za = __dadd_rn(__dmul_rn((double)iz, dz), z0), zb = __dadd_rn(__dmul_rn((double)(iz + 1), dz), z0);
  // corner order :77-81: (0,0,0) (1,0,0) (1,1,0) (1,1,1) (0,1,0) (0,1,1) (0,0,1) (1,0,1)
  double sx = xa, sy = ya, sz = za;
  sx = __dadd_rn(sx, xb); sy = __dadd_rn(sy, ya); sz = __dadd_rn(sz, za);
  sx = __dadd_rn(sx, xb); sy = __dadd_rn(sy, yb); sz = __dadd_rn(sz, za);
  sx = __dadd_rn(sx, xb); sy = __dadd_rn(sy, yb); sz = __dadd_rn(sz, zb);
  sx = __dadd_rn(sx, xa); sy = __dadd_rn(sy, yb); sz = __dadd_rn(sz, za);
  sx = __dadd_rn(sx, xa); sy = __dadd_rn(sy, yb); sz = __dadd_rn(sz, zb);
  sx = __dadd_rn(sx, xa); sy = __dadd_rn(sy, ya); sz = __dadd_rn(sz, zb);
  sx = __dadd_rn(sx, xb); sy = __dadd_rn(sy, ya); sz = __dadd_rn(sz, zb);
  P[3 * q + 0] = sx * 0.125;
  P[3 * q + 1] = sy * 0.125;
  P[3 * q + 2] = sz * 0.125;
}

__global__ void add_freestream_kernel(long long m, double vx, double vy, double vz, double* __restrict__ V) {
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= m) return;
  V[3 * q + 0] = V[3 * q + 0] + vx;
  V[3 * q + 1] = V[3 * q + 1] + vy;
  V[3 * q + 2] = V[3 * q + 2] + vz;
}

}  // namespace vlc
