// bs_lattice.cuh -- shared-node sweep over a vortex-ring LATTICE (near wake), the structured fast path of K1/K2.
//
// The reference enumerates a near wake ring by ring: 4 filaments per ring, every interior lattice edge twice (once
// per adjacent ring, opposite directions), every node 8 times as a filament end point
// (src/classdef.f90:1450-1456 -> vr_vind :527-542 -> vf_vind :476-503).  On a lattice the same sum can be
// regrouped exactly:
//   * per (target, NODE):  r = P - X,  u = 1/|r|                      -- once instead of 8 times
//   * per (target, EDGE U->V) with merged strength g = (Gamma_a - Gamma_b)/4pi of the two rings that share it
//     (the reversed copy of a filament induces exactly the negated velocity):
//        c = rU x rV,  v += c * (g r0.rU * uU - g r0.rV * uV) / sqrt(K + |c|^4)
// which is the reference formula with unitVec(r) = r*u.  A strip walk keeps the node quantities in registers:
// a strip record of width W describes one node row of W+1 adjacent node columns, the W spanwise edges between
// them and the W streamwise edges that reach the first W of them from the previous row (whose node quantities are
// still in registers).  For W = 1: 2 nodes (11 FP64 instr each) + 2 edges (25 each) = 72 FP64-pipe instructions
// per (target, ring) = 4 reference pair interactions = 18 per pair (the flat kernel needs 43); W = 2..4 bring it
// to 66.5 / 64.7 / 63.75.
//
// Merging needs both copies of an edge to carry the same core radius; the pack kernels check this bitwise and raise a
// device flag otherwise: streamwise copies that differ select the DUAL form of this kernel (below), anything else the
// flat kernel on the reference enumeration (capi.cu: sweep_shared) -- results never depend on the assumption.
#pragma once
#include "vlc_device.cuh"

namespace vlc {

// Strip record of width W (W ring columns = W+1 node columns, one node row), all doubles:
//   [0 .. 3(W+1))            nodes N_0 .. N_W of this row                     (padded to an even count)
//   then for k = 0 .. W-1    spanwise edge N_k -> N_{k+1}:  g*(r0)[3], g*|r0|^2, K       (5)
//                            streamwise edge Nprev_k -> N_k: g*(r0)[3], g*|r0|^2, K      (5)
//   then for k = 0 .. W-1    4 more doubles, used by the DUAL form only (below)
// W = 1: 16 + 4 doubles (A, B, edge A->B, edge A_prev->A).  Wider strips amortise the node work over more edges:
// FP64 instructions per ring = (11 (W+1) + 50 W) / W = 72, 66.5, 64.7, 63.75 for W = 1..4.
//
// DUAL form (round-1 review, task 5): the two copies of a shared STREAMWISE edge -- vf(1) of ring (r, j) and vf(3) of
// ring (r, j-1) -- carry different core radii whenever streamwiseCoreVec is not uniform (the reference keeps both:
// classdef.f90:3841, and rotor_dissipate_wake's vf(3)%rVc <- vf(1)%rVc, :4371-4372).  Such an edge cannot be
// merged into one strength, but it still shares its two nodes, its cross product and the end-point term:
//     v += c * (r0.rU uU - r0.rV uV) * (gA / sqrt(KA + |c|^4) + gB / sqrt(KB + |c|^4))
// = 33 instead of 25 FP64 instructions (two reciprocal square roots); nodes and spanwise edges are unchanged, so a ring
// costs (11 (W+1) + 58 W) / W = 71.75 at W = 4 -- against 172 on the flat enumeration the whole set fell back to in
// round 1.  The streamwise slot then holds r0[3], |r0|^2, gA and the extra four doubles KA, gB, KB, 0 (gB carries the
// sign of the reversed copy).  Which form a buffer holds is decided ON THE DEVICE by the set's flag (pack.cuh:
// check_rings_kernel): 0 = merged, 2 = dual, odd = not a lattice / spanwise copies differ -> flat enumeration.
__host__ __device__ constexpr int lat_nodes_pad(int W) { return (3 * (W + 1) + 1) / 2 * 2; }
__host__ __device__ constexpr int lat_core_doubles(int W) { return lat_nodes_pad(W) + 10 * W; }  // what the merged form reads
__host__ __device__ constexpr int lat_rec_doubles(int W) { return lat_core_doubles(W) + 4 * W; }  // record stride
#ifndef VLC_LAT_TILE_DIV
#define VLC_LAT_TILE_DIV 1
#endif
__host__ __device__ constexpr int lat_tile(int W) { return (W <= 2 ? 64 : 32) / VLC_LAT_TILE_DIV; }  // records per shared-memory tile
// A source split (chunk) is a multiple of the GRANULE, a quarter of a tile: the last tile of a chunk may be partial.  Small
// sweeps (a few thousand targets x a few hundred records: the reference's own test cases) need more CTAs than whole tiles give.
__host__ __device__ constexpr int lat_granule(int W) { return lat_tile(W) / 4; }

struct NodeQ {
  double rx, ry, rz, u;  // r = P - X, u = 1/|r|
};

__device__ __forceinline__ NodeQ node_eval(double px, double py, double pz, double ax, double ay, double az) {
  NodeQ n;
  n.rx = px - ax;
  n.ry = py - ay;
  n.rz = pz - az;
  const double d = fma(n.rz, n.rz, fma(n.ry, n.ry, n.rx * n.rx));
  // d == 0 (target on the node) gives NaN; every edge that uses it has c == 0 and is guarded.  Unlike libMath.f90:257
  // (unitVec = 0 when |r| <= eps) u is NOT zeroed for 0 < |r| <= 2.2e-16: for a filament longer than ~1 such a target
  // passes c2 > eps^2 and keeps the r/|r| piece the reference drops; the difference is <= eps/rVc^2 in absolute terms
  // (tests/test_gpu_parity.py::test_targets_within_eps_of_a_node_of_a_long_filament)
  n.u = rsqrt_fp64<false>(d);
  return n;
}

// One (target, edge U->V) interaction: 25 FP64-pipe instructions.
__device__ __forceinline__ void edge_accumulate(const NodeQ& a, const NodeQ& b, double gx, double gy, double gz,
                                                double L2g, double K, double& vx, double& vy, double& vz) {
  const double cx = fma(a.ry, b.rz, -(a.rz * b.ry));
  const double cy = fma(a.rz, b.rx, -(a.rx * b.rz));
  const double cz = fma(a.rx, b.ry, -(a.ry * b.rx));
  const double c2 = fma(cz, cz, fma(cy, cy, cx * cx));
  const double a1 = fma(gz, a.rz, fma(gy, a.ry, gx * a.rx));  // g r0.rU
  const double a2 = a1 - L2g;                                  // g r0.rV
  const double w = rsqrt_edge(fma(c2, c2, K));  // 3 FP64 instructions + an FP32 seed off the pipe (vlc_device.cuh)
  double sc = fma(-a2, b.u, a1 * a.u) * w;
  // classdef.f90:498 `if (r1Xr2Abs2 > eps*eps)`: c2 >= 0 orders like its bit pattern, eps^2 = 2^-104 =
  // 0x3970000000000000: one 64-bit integer compare + a select of sc's high word (guard_scale), nothing extra on the FP64 pipe.  (Predicating the three
  // accumulates instead is if-converted by ptxas into three selects -- measured, worse.)  NaN/Inf in sc (target on a
  // node) never reach the accumulators because those edges have c == 0 exactly and are selected away here.
  guard_scale(sc, c2);
  vx = fma(cx, sc, vx);
  vy = fma(cy, sc, vy);
  vz = fma(cz, sc, vz);
}

// One (target, streamwise edge U->V) interaction whose two copies carry different core radii: 33 FP64-pipe instructions.
__device__ __forceinline__ void edge_accumulate_dual(const NodeQ& a, const NodeQ& b, double ex, double ey, double ez, double L2,
                                                     double gA, double KA, double gB, double KB, double& vx, double& vy,
                                                     double& vz) {
  const double cx = fma(a.ry, b.rz, -(a.rz * b.ry));
  const double cy = fma(a.rz, b.rx, -(a.rx * b.rz));
  const double cz = fma(a.rx, b.ry, -(a.ry * b.rx));
  const double c2 = fma(cz, cz, fma(cy, cy, cx * cx));
  const double a1 = fma(ez, a.rz, fma(ey, a.ry, ex * a.rx));  // r0.rU
  const double a2 = a1 - L2;                                   // r0.rV
  const double wA = rsqrt_fp64<false>(fma(c2, c2, KA));
  const double wB = rsqrt_fp64<false>(fma(c2, c2, KB));
  const double t = fma(-a2, b.u, a1 * a.u);
  double sc = t * fma(gB, wB, gA * wA);
  guard_scale(sc, c2);
  vx = fma(cx, sc, vx);
  vy = fma(cy, sc, vy);
  vz = fma(cz, sc, vz);
}

// flag dispatch: the kernel runs when (*flag & mask) == want (flag == nullptr: always)
template <int W, int T, int THREADS, int STAGES, int MINB, bool DUAL = false>
__global__ void __launch_bounds__(THREADS, MINB)
bs_lattice_kernel(const double* __restrict__ lat,  // strip records of width W, padded to a multiple of lat_tile(W)
                  long long chunk,                 // records per split (multiple of the granule)
                  long long n_pad,                 // total padded records
                  const double* __restrict__ P, long long m,
                  double* __restrict__ out,        // [y0 + gridDim.y][3 m]: this launch fills slots y0 .. y0 + gridDim.y - 1
                  const int* __restrict__ flag, int mask, int want,
                  int y0) {                        // first source split of this launch (0 unless the splits are shared out
                                                   // between devices: collocation-point stage of a group, capi.cu)
  if (flag != nullptr && (*flag & mask) != want) return;  // uniform: another form of the set does the work
  constexpr int RD = lat_rec_doubles(W), NP = lat_nodes_pad(W), TILE = lat_tile(W);
  constexpr int RL = DUAL ? RD : lat_core_doubles(W);  // doubles of a record this form reads
#if defined(__CUDA_EMUL__)  // host build of the tests (tests/native/kernels_emul.cpp): one emulated thread at a time, which
  // stages its own tiles (VLC_PRODUCER) into a static buffer with memcpy standing in for the bulk copy
  alignas(128) static unsigned char smem_raw[(size_t)STAGES * TILE * RD * 8 + STAGES * 8];
#else
  extern __shared__ __align__(128) unsigned char smem_raw[];
#endif
  double* buf = reinterpret_cast<double*>(smem_raw);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + (size_t)STAGES * TILE * RD * 8);

  const int tid = threadIdx.x;
  const long long s_begin = ((long long)blockIdx.y + y0) * chunk;
  long long s_end = s_begin + chunk;
  if (s_end > n_pad) s_end = n_pad;
  const long long len = s_end > s_begin ? s_end - s_begin : 0;  // a multiple of the granule (even)
  const int ntiles = (int)((len + TILE - 1) / TILE);
  const double* gsrc = lat + s_begin * RD;
  auto tile_records = [&](int t) -> int {  // TILE, except for the last tile of a chunk that is not a whole number of tiles
    const long long left = len - (long long)t * TILE;
    return left < TILE ? (int)left : TILE;
  };

  if (VLC_PRODUCER(tid)) {
#pragma unroll
    for (int s = 0; s < STAGES; ++s) mbar_init(&bars[s], 1);
    fence_mbar_init();
  }
  __syncthreads();
  if (VLC_PRODUCER(tid)) {
#pragma unroll
    for (int s = 0; s < STAGES; ++s)
      if (s < ntiles) {
        const uint32_t bytes = (uint32_t)tile_records(s) * (RD * 8);
        mbar_expect_tx(&bars[s], bytes);
        tma_bulk_g2s(buf + (size_t)s * TILE * RD, gsrc + (size_t)s * TILE * RD, bytes, &bars[s]);
      }
  }

  const long long t0 = (long long)blockIdx.x * (THREADS * T) + tid;
  double px[T], py[T], pz[T], vx[T], vy[T], vz[T];
  NodeQ prev[T][W];
  // Nprev of the first record of this chunk = the nodes of the record before it (same strip unless the record
  // starts a strip, in which case its streamwise strengths are 0 and any finite nodes will do).
  const double* r0 = lat + (s_begin > 0 ? (s_begin - 1) : 0) * RD;
#pragma unroll
  for (int k = 0; k < T; ++k) {
    const long long t = t0 + (long long)k * THREADS;
    const bool ok = t < m;
    px[k] = ok ? P[3 * t + 0] : 0.0;
    py[k] = ok ? P[3 * t + 1] : 0.0;
    pz[k] = ok ? P[3 * t + 2] : 0.0;
    vx[k] = vy[k] = vz[k] = 0.0;
#pragma unroll
    for (int i = 0; i < W; ++i) prev[k][i] = node_eval(px[k], py[k], pz[k], r0[3 * i], r0[3 * i + 1], r0[3 * i + 2]);
  }

  for (int tile = 0; tile < ntiles; ++tile) {
    const int stage = tile % STAGES;
    const uint32_t phase = (uint32_t)(tile / STAGES) & 1u;
    mbar_wait(&bars[stage], phase);
    const double* sbase = buf + (size_t)stage * TILE * RD;
    const int jn = tile_records(tile);
#pragma unroll 1
    for (int j2 = 0; j2 < jn; j2 += 2)  // two records per trip (jn is even): the unroll the fixed-length loop had
#pragma unroll
    for (int jj = 0; jj < 2; ++jj) {
      const int j = j2 + jj;
      const double2* sb = reinterpret_cast<const double2*>(sbase + (size_t)j * RD);
      double q[RL];  // the record, in registers (constant indices after unrolling)
#pragma unroll
      for (int i = 0; i < RL / 2; ++i) {
        const double2 v = sb[i];
        q[2 * i] = v.x;
        q[2 * i + 1] = v.y;
      }
#pragma unroll
      for (int k = 0; k < T; ++k) {
        // nodes are evaluated one column ahead of the edges that need them (keeps few node quantities live)
        NodeQ cur = node_eval(px[k], py[k], pz[k], q[0], q[1], q[2]);
#pragma unroll
        for (int i = 0; i < W; ++i) {
          const NodeQ nxt = node_eval(px[k], py[k], pz[k], q[3 * i + 3], q[3 * i + 4], q[3 * i + 5]);
          const double* e = &q[NP + 10 * i];
          if (DUAL) {  // streamwise Nprev_i -> N_i with the core radii of both copies
            const double* x = &q[NP + 10 * W + 4 * i];
            edge_accumulate_dual(prev[k][i], cur, e[5], e[6], e[7], e[8], e[9], x[0], x[1], x[2], vx[k], vy[k], vz[k]);
          } else {
            edge_accumulate(prev[k][i], cur, e[5], e[6], e[7], e[8], e[9], vx[k], vy[k], vz[k]);  // streamwise Nprev_i -> N_i
          }
          edge_accumulate(cur, nxt, e[0], e[1], e[2], e[3], e[4], vx[k], vy[k], vz[k]);         // spanwise   N_i -> N_{i+1}
          prev[k][i] = cur;
          cur = nxt;
        }
      }
    }
    __syncthreads();
    if (VLC_PRODUCER(tid) && tile + STAGES < ntiles) {
      const uint32_t bytes = (uint32_t)tile_records(tile + STAGES) * (RD * 8);
      mbar_expect_tx(&bars[stage], bytes);
      tma_bulk_g2s(buf + (size_t)stage * TILE * RD, gsrc + (size_t)(tile + STAGES) * TILE * RD, bytes, &bars[stage]);
    }
  }

  double* o = out + ((size_t)blockIdx.y + (size_t)y0) * 3 * (size_t)m;
#pragma unroll
  for (int k = 0; k < T; ++k) {
    const long long t = t0 + (long long)k * THREADS;
    if (t < m) {
      o[3 * t + 0] = vx[k];
      o[3 * t + 1] = vy[k];
      o[3 * t + 2] = vz[k];
    }
  }
}

// Fixed-order sum of partial slots, selecting the slots of the form that actually ran.  Slot layout:
//   [0, na) merged lattice (+ tail strips) | [na, na+nr) flat remainder | [na+nr, na+nr+nd) dual lattice (+ tail strips) |
//   [na+nr+nd, na+nr+nd+nf) flat fallback
//   *flag == 0 -> [0, na+nr)     *flag == 2 -> [na, na+nr+nd)     *flag odd -> the last nf slots
__global__ void bs_reduce_select_kernel(const double* __restrict__ part, const int* __restrict__ flag, int na, int nr, int nd,
                                        int nf, long long len, double* __restrict__ V) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= len) return;
  const int f = *flag;
  const int first = (f & 1) ? na + nr + nd : (f == 2 ? na : 0);
  const int count = (f & 1) ? nf : (f == 2 ? nr + nd : na + nr);
  // the slots are ADDED in order (the result does not depend on how the loop is written); the loads of eight slots are
  // issued together -- with up to 256 slots of a small sweep the one-load-one-add loop was a chain of DRAM latencies
  // (13.6 us per sweep of K&P, r03j)
  const double* p = part + (size_t)first * len + i;
  double a = 0.0;
  int s = 0;
  for (; s + 8 <= count; s += 8) {
    double v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = p[(size_t)(s + k) * len];
#pragma unroll
    for (int k = 0; k < 8; ++k) a += v[k];
  }
  for (; s < count; ++s) a += p[(size_t)s * len];
  V[i] = a;
}

}  // namespace vlc
