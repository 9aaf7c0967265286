// wake_state.cuh -- O(N) state-update kernels of the wake (K6-K8 of SURVEY 2.1) and small helpers.
// HBM-bound elementwise maps; each cites the reference statement it restates.
#pragma once
#include <cuda_runtime.h>

namespace vlc {

// vr_shiftdP / Fwake_shiftdP with dshift = vel*dt (classdef.f90:1531, :1538, :1545; :611-616, :946)
__global__ void convect_kernel(long long n, double* __restrict__ x, const double* __restrict__ v, double dt) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) x[i] = __dadd_rn(x[i], __dmul_rn(v[i], dt));  // no FMA contraction: bit-parity with the reference arithmetic
}

// main.f90:1032-1034  velNwake = 0.5*(3*velNwake - velNwake1)   (Adams-Bashforth predictor velocity)
__global__ void ab2_kernel(long long n, const double* __restrict__ v, const double* __restrict__ v1,
                           double* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = __dmul_rn(0.5, __dadd_rn(__dmul_rn(3.0, v[i]), -v1[i]));
}

// The velocity bookkeeping of a predictor / corrector stage on the near- and far-wake arrays of a rotor in ONE launch
// (near array = elements [0, nn), far array = [nn, nn + nf)); same arithmetic as ab2_kernel / am2_kernel.
//   op 0: main.f90:1031-1041  velStep = vel; vel = 0.5*(3*vel - vel1)
//   op 1: main.f90:1094-1099  vel = (velPredicted + velStep)*0.5
//   op 2: dst = src (vel1 = vel, vel1 = velStep, velStep = vel)
struct VelArrays {
  double *n0, *n1, *n2, *n3, *f0, *f1, *f2, *f3;  // vel, vel1, velPredicted, velStep: near / far
};
__global__ void wakevel_fused_kernel(int op, long long nn, long long nf, VelArrays a, int dst, int src) {
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= nn + nf) return;
  const bool far = q >= nn;
  const long long i = far ? q - nn : q;
  double* v[4] = {far ? a.f0 : a.n0, far ? a.f1 : a.n1, far ? a.f2 : a.n2, far ? a.f3 : a.n3};
  if (op == 0) {
    const double s = v[0][i];
    v[3][i] = s;
    v[0][i] = __dmul_rn(0.5, __dadd_rn(__dmul_rn(3.0, s), -v[1][i]));
  } else if (op == 1) {
    v[0][i] = __dmul_rn(__dadd_rn(v[2][i], v[3][i]), 0.5);
  } else {
    v[dst][i] = v[src][i];
  }
}

// main.f90:1094-1096  velNwake = (velNwakePredicted + velNwakeStep)*0.5   (Adams-Moulton corrector)
__global__ void am2_kernel(long long n, const double* __restrict__ vp, const double* __restrict__ vs,
                           double* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (vp[i] + vs[i]) * 0.5;
}

// libCommon.f90:213-258  vel_order2_Nwake / vel_order2_Fwake (fdScheme 1): arrays (3, rows, cols), cols = 1 for the
// far wake.  Row 1 and row `rows`: (vnp1 + vn)*0.5; inner rows: (((vnp1(i) + vnp1(i-1)) + vn(i+1)) + vn(i))*0.25.
__global__ void vel_order2_kernel(int rows, int cols, const double* __restrict__ vn, const double* __restrict__ vnp1,
                                  double* __restrict__ out) {
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= 3LL * rows * cols) return;
  const int i = (int)((q / 3) % rows);
  if (i == 0 || i == rows - 1)
    out[q] = __dmul_rn(__dadd_rn(vnp1[q], vn[q]), 0.5);
  else
    out[q] = __dmul_rn(__dadd_rn(__dadd_rn(__dadd_rn(vnp1[q], vnp1[q - 3]), vn[q + 3]), vn[q]), 0.25);
}

// classdef.f90:4368-4370 / :4398-4401  rVc = sqrt(rVc**2 + 4*oseenParameter*apparentViscCoeff*nu*dt)
__device__ __forceinline__ double grow(double rvc, double a, double nu, double dt) {
  return sqrt(__dadd_rn(__dmul_rn(rvc, rvc), 4.0 * 1.2564 * a * nu * dt));
}

__global__ void core_growth_kernel(long long n, double* __restrict__ rvc, double a, double nu, double dt) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) rvc[i] = grow(rvc[i], a, nu, dt);
}

// classdef.f90:662-668, :975-980  gam = gam*exp(-decayCoeff*dt)
__global__ void decay_kernel(long long n, double* __restrict__ gam, double decayCoeff, double dt) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) gam[i] = gam[i] * exp(-decayCoeff * dt);
}

// classdef.f90:4364-4384 on rvc4 (4, nrows, ns): vf1 grows, vf3 <- vf1 (quirk C2), gam decays, vf2 grows.
__global__ void dissipate_lattice_kernel(int nrows, int ns, double* __restrict__ rvc4, double* __restrict__ gam,
                                         double a, double nu, double decayCoeff, double dt) {
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= (long long)nrows * ns) return;
  double* r = rvc4 + 4 * q;
  const double r1 = grow(r[0], a, nu, dt);
  r[0] = r1;
  r[2] = r1;
  gam[q] = gam[q] * exp(-decayCoeff * dt);
  r[1] = grow(r[1], a, nu, dt);
}

// classdef.f90:4386-4392  vf4(i) <- vf2(i-1) for i > first row (second pass: needs the updated vf2).
__global__ void dissipate_lattice_vf4_kernel(int nrows, int ns, double* __restrict__ rvc4) {
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= (long long)nrows * ns) return;
  const int r = (int)(q % nrows);
  if (r > 0) rvc4[4 * q + 3] = rvc4[4 * (q - 1) + 1];
}

// classdef.f90:4414-4421 + :505-521  lc = |fc1 - fc2|; rVc = rVc0*sqrt(l0/lc)
__global__ void strain_kernel(long long n, const double* __restrict__ p1, const double* __restrict__ p2,
                              const double* __restrict__ l0, const double* __restrict__ rvc0,
                              double* __restrict__ rvc) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double dx = p1[3 * i] - p2[3 * i], dy = p1[3 * i + 1] - p2[3 * i + 1], dz = p1[3 * i + 2] - p2[3 * i + 2];
  const double lc = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)));
  rvc[i] = rvc0[i] * sqrt(l0[i] / lc);
}

// nodes(3, nrows+1, ns+1) rows 1..nrows <-> P(3, nrows, ns+1)  (targets of libCommon.f90:133-145)
__global__ void lattice_gather_kernel(int nrows, int ns, const double* __restrict__ nodes, double* __restrict__ P) {
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= 3LL * nrows * (ns + 1)) return;
  const int d = (int)(q % 3);
  const long long t = q / 3;
  const int r = (int)(t % nrows), c = (int)(t / nrows);
  P[q] = nodes[3 * ((size_t)(r + 1) + (size_t)(nrows + 1) * c) + d];
}
__global__ void lattice_scatter_kernel(int nrows, int ns, double* __restrict__ nodes, const double* __restrict__ P) {
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= 3LL * nrows * (ns + 1)) return;
  const int d = (int)(q % 3);
  const long long t = q / 3;
  const int r = (int)(t % nrows), c = (int)(t / nrows);
  nodes[3 * ((size_t)(r + 1) + (size_t)(nrows + 1) * c) + d] = P[q];
}

__global__ void identity_kernel(int N, double* __restrict__ A) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)N * N) return;
  A[i] = ((i % N) == (i / N)) ? 1.0 : 0.0;
}

// gamVec = AIC_inv . RHS (main.f90:190, :596: matmulAX = DGEMV 'N', libMath.f90:105-122), A column-major N x N.
// One thread per (row, column slice): kGemvSlices partial sums over contiguous column ranges, each accumulated in column
// order like DGEMV's axpy sweep, combined in slice order -- a fixed order, independent of the launch.
constexpr int kGemvRows = 32;
constexpr int kGemvSlices = 8;
__global__ void __launch_bounds__(kGemvRows* kGemvSlices)
    ainv_gemv_kernel(int N, const double* __restrict__ A, const double* __restrict__ x, double* __restrict__ y) {
  __shared__ double part[kGemvSlices][kGemvRows];
  const int i = blockIdx.x * kGemvRows + threadIdx.x, s = threadIdx.y;
  const int per = (N + kGemvSlices - 1) / kGemvSlices;
  const int j0 = s * per, j1 = min(N, j0 + per);
  double acc = 0.0;
  if (i < N)
    for (int j = j0; j < j1; ++j) acc = fma(A[(size_t)j * N + i], x[j], acc);
  part[s][threadIdx.x] = acc;
  __syncthreads();
  if (s == 0 && i < N) {
    double t = part[0][threadIdx.x];
#pragma unroll
    for (int k = 1; k < kGemvSlices; ++k) t += part[k][threadIdx.x];
    y[i] = t;
  }
}

// FP64 roofline denominator: kPeakChains independent register-resident DFMA chains per thread.
constexpr int kPeakChains = 8;
constexpr int kPeakUnroll = 16;
__global__ void dfma_peak_kernel(int iters, double* __restrict__ out) {
  double a[kPeakChains];
  const double b = 1.0000001, c = 1e-9 * (threadIdx.x + 1);
#pragma unroll
  for (int k = 0; k < kPeakChains; ++k) a[k] = 1.0 + k * 1e-3 + threadIdx.x * 1e-6;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < kPeakUnroll; ++u)
#pragma unroll
      for (int k = 0; k < kPeakChains; ++k) a[k] = fma(a[k], b, c);
  }
  double s = 0.0;
#pragma unroll
  for (int k = 0; k < kPeakChains; ++k) s += a[k];
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// The same with THREE distinct, changing register operands per DFMA (a = b*c + a; b = c*a + b; c = a*b + c): what a real
// kernel's instructions look like to the register file (the chain above feeds two of its three operands from a
// constant and a loop-invariant register).  8 independent chains per phase.
__global__ void dfma_peak3_kernel(int iters, double* __restrict__ out) {
  double a[kPeakChains], b[kPeakChains], c[kPeakChains];
#pragma unroll
  for (int k = 0; k < kPeakChains; ++k) {
    a[k] = 1e-3 * (k + 1) + threadIdx.x * 1e-9;
    b[k] = 0.5 + 1e-2 * k;
    c[k] = 1e-4 * (threadIdx.x + 1) + 1e-5 * k;
  }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < kPeakUnroll / 4; ++u) {
#pragma unroll
      for (int k = 0; k < kPeakChains; ++k) a[k] = fma(b[k], c[k], a[k]);
#pragma unroll
      for (int k = 0; k < kPeakChains; ++k) b[k] = fma(c[k], a[k], b[k]);
#pragma unroll
      for (int k = 0; k < kPeakChains; ++k) c[k] = fma(a[k], b[k], c[k]);
#pragma unroll
      for (int k = 0; k < kPeakChains; ++k) a[k] = fma(c[k], b[k], -a[k]);
    }
  }
  double s = 0.0;
#pragma unroll
  for (int k = 0; k < kPeakChains; ++k) s += a[k] + b[k] + c[k];
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

}  // namespace vlc

// ---- probe: raw MUFU.RSQ64H seed and both Newton refinements (tests measure the seed error bound) ----
#include "vlc_device.cuh"
namespace vlc {
__global__ void rsqrt_probe_kernel(long long n, const double* __restrict__ x, double* __restrict__ seed,
                                   double* __restrict__ full, double* __restrict__ fast) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x[i]));
  seed[i] = y;
  full[i] = rsqrt_fp64<false>(x[i]);
  fast[i] = rsqrt_fp64<true>(x[i]);
}
}  // namespace vlc
