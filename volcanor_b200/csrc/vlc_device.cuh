// vlc_device.cuh -- device-side building blocks of the Biot-Savart hot path (sm_100a, FP64).
//
// The pair interaction restates vf_vind (reference src/classdef.f90:476-503) in a form
// that keeps the FP64 pipe busy:
//   reference : v = (c * inv4pi * r0.(r1/|r1| - r2/|r2|)) / sqrt((rVc |r0|)^4 + |c|^4),  c = r1 x r2,
//               0 when |c|^2 <= eps^2
//   here      : v = c * [ (G r0.r1) * w1 - (G r0.r2) * w2 ],
//               w_k = rsqrt(|r_k|^2 * (K + |c|^4)),  K = (rVc^2 |r0|^2)^2,  G = gam/(4 pi)
// G r0, G |r0|^2 and K are per-SOURCE quantities computed once by the pack kernels, so a pair
// costs 2 reciprocal square roots instead of 3 square roots + 9 divides.  r1 = P - p1 and
// r2 = P - p2 are formed exactly as the reference does, so a target that coincides bitwise
// with a filament end point still gives c == 0 exactly and is skipped by the same guard
// (this is how wake nodes skip the filaments they belong to).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace vlc {

// libMath.f90:11 eps = epsilon(1._dp); classdef.f90:498 guard is c2 > eps*eps = 2^-104.
__device__ __constant__ const double kEps = 2.220446049250313e-16;
#define VLC_EPS2 4.930380657631324e-32 /* 2^-104 exactly */

// A packed source filament: 12 doubles = 96 B = 6 x 16 B (LDS.128 broadcast friendly).
//   [0..2] p1   [3..5] p2   [6..8] G*r0, r0 = p2 - p1   [9] G*|r0|^2   [10] K = (rVc^2 |r0|^2)^2   [11] G = gam/4pi
constexpr int kSrcDoubles = 12;
constexpr int kSrcBytes = kSrcDoubles * 8;

// Reciprocal square root on the FP64 pipe: MUFU.RSQ64H seed y0 (only the high word of x is used: measured
// relative error d <= 9.2e-7 = 2^-20.06 on B200, tests/test_gpu_parity.py::test_rsqrt_seed_accuracy)
// refined by Newton.
//   FULL (third order: y1 = y0 + y0*e*(1/2 + 3/8 e), e = 1 - x y0^2): 5 FP64 instr, error ~2.5 d^3 ~ 2e-18 ->
//        results limited by rounding (measured 1.1e-16).  DEFAULT.
//   FAST (second order): 4 FP64 instr, y1 = y0 (1 - 1.5 d^2 + ...): relative error <= 1.5 d^2 = 1.3e-12,
//        always low; bs_sweep_kernel centres it with one multiply by (1 + 0.75 d_max^2) per target, leaving
//        <= 6.4e-13 per pair (1.5e-13 measured on a whole sweep).  Inside the 1e-12 tolerance but with
//        little margin, hence opt-in only (vlc_set_precision).
// No special-case branch: inputs are finite > 0 whenever the result is used (the c2 guard predicates
// the accumulation otherwise).
constexpr double kSeedRelErr = 9.5367431640625e-07;  // 2^-20 >= measured max 9.18e-7
// Operand sourcing matters on this pipe (tools/micro/fp64_operands.cu, profiles/r01h_fp64_operands.md): a DFMA whose
// three operands are three DIFFERENT registers issues every 3 cycles per scheduler, one with at most two different
// registers (an immediate, or the same register in two slots) every 2.  The last refinement step is therefore written
// y + y*(e*p) -- DMUL(e, p) then DFMA(y, ep, y), two different registers each -- instead of fma(p, e*y, y): the same
// five instructions, the same single rounding of the result, one issue cycle less per reciprocal square root.
template <bool FAST>
__device__ __forceinline__ double rsqrt_fp64(double x) {
  double y;
#if defined(__CUDA_EMUL__)  // host build of the tests (tests/native/emul): a model of MUFU.RSQ64H -- only the high word of x
  // is read and the result carries about 23 bits (relative error <= 2^-21 + 2^-23, the device's measured 2^-20.06 is of
  // the same size), so that the refinement below is exercised with a seed as coarse as the real one
  {
    long long xb = __double_as_longlong(x) & ~0xFFFFFFFFLL;
    double xh;
    std::memcpy(&xh, &xb, sizeof xh);
    y = 1.0 / std::sqrt(xh);
    long long yb = __double_as_longlong(y) & ~0x1FFFFFFFLL;
    std::memcpy(&y, &yb, sizeof y);
  }
#else
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
#endif
  const double t = y * y;
  const double e = fma(-x, t, 1.0);
  if (FAST) return fma(y, 0.5 * e, y);
  const double p = fma(0.375, e, 0.5);
  const double ep = e * p;
  return fma(y, ep, y);
}

// Reciprocal square root for the EDGE term w = rsqrt(K + |c|^4) of the shared-node kernel (bs_lattice.cuh): its relative
// error enters a pair's contribution as it is (no cancellation follows it -- unlike the node quantity u = 1/|r|, whose
// error is amplified by r/L in r0.(r1 u1 - r2 u2) and therefore keeps the full refinement above), so 5e-14 is enough
// for the 1e-12 per-target bar with a factor 20 to spare.  That allows ONE second-order Newton step if the seed carries
// ~22.7 bits instead of MUFU.RSQ64H's 20, and such a seed is available off the FP64 pipe (round-1 review, item 3; ncu
// r01j: 46 % of the issue slots idle, the ALU / XU pipes almost unused):
//   x = 2^(2k+p) * 1.f, p in {-1, 0}  ->  m = 2^p * 1.f in [0.5, 2) as an FP32 number, built with two integer instructions
//       from the top 23 fraction bits of x and the low bit of its exponent;
//   ym = MUFU.RSQ(m) (FP32, error 2^-22.9; the truncation of x adds <= 2^-24);
//   y = ym * 2^-k assembled bitwise as an FP64 number (three shifts, one add), yh = y/2 by an exponent decrement;
//   result = y + yh*(1 - x y^2): THREE FP64 instructions (DMUL, DFMA with an immediate, DFMA) instead of five.
// Measured over 2e6 log-uniform x in [1e-62, 1e60] (numpy model, tests/test_kernels_emul.py on the same source): seed
// error <= 2^-22.7, result error <= 3.2e-14.  x must be a positive normal number (the callers' c2 guard discards every
// other case: K + c2^2 >= 2^-208 whenever the result is used).
// VLC_EDGE_RSQRT: 0 = rsqrt_fp64<false> (round 1; DEFAULT), 1 = this.
// MEASURED (r02b, profiles/r02b_edge_rsqrt.md): 1 is SLOWER on B200 -- 274.2 ms instead of 259.6 ms per sweep at 1e6
// filaments although the loop holds 956 instead of 1020 FP64 instructions: the 8 integer / shift instructions per seed
// (557 instead of 266 non-FP64 instructions per loop iteration) cost about one issue cycle each, i.e. 0.44 of an FP64
// instruction, more than the two FP64 instructions they save.  The "idle" 46 % of the issue slots are not free: with two
// warps per scheduler the FP64 unit is fed only while nothing else competes for the dispatch port.  Kept as an option
// for the record; every non-FP64 instruction REMOVED from the loop is worth the same 0.44.
#ifndef VLC_EDGE_RSQRT
#define VLC_EDGE_RSQRT 0
#endif
__device__ __forceinline__ double rsqrt_edge(double x) {
#if VLC_EDGE_RSQRT == 0
  return rsqrt_fp64<false>(x);
#else
  const unsigned hi = (unsigned)__double2hiint(x), lo = (unsigned)__double2loint(x);
#if defined(__CUDA_EMUL__)
  const unsigned v = (hi << 3) | (lo >> 29);
#else
  const unsigned v = __funnelshift_l(lo, hi, 3);                 // [e8..e0 | fraction bits 51..29]
#endif
  const unsigned fb = (v & 0x00FFFFFFu) | 0x3F000000u;            // exponent field 126 + e0: m in [0.5, 2)
  float ym;
#if defined(__CUDA_EMUL__)
  {
    float m;
    std::memcpy(&m, &fb, sizeof m);
    ym = 1.0f / std::sqrt(m);
  }
  unsigned yb;
  std::memcpy(&yb, &ym, sizeof yb);
#else
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(ym) : "f"(__uint_as_float(fb)));
  const unsigned yb = __float_as_uint(ym);
#endif
  // k = ((e11 | 1) - 1023)/2;  high word of ym as FP64 = (yb >> 3) + (896 << 20);  minus k << 20:
  const unsigned B = (hi & 0x7FF00000u) | 0x00100000u;
  const unsigned hy = (yb >> 3) + 0x57F80000u - (B >> 1);
  const unsigned ly = yb << 29;
  const double y = __hiloint2double((int)hy, (int)ly);
  const double yh = __hiloint2double((int)(hy - 0x00100000u), (int)ly);
  const double t = y * y;
  const double e = fma(-x, t, 1.0);
  return fma(yh, e, y);
#endif
}

// sc <- (c2 > eps^2) ? sc : 0 without touching the FP64 pipe.  VLC_GUARD_HI (default): only the HIGH word of sc is
// selected (one SEL instead of two): a guarded sc becomes {lo, 0} = a denormal < 2^-1042 (whatever sc was, NaN and
// Inf included), and every |c_i| <= 2^-52 there (c2 <= 2^-104), so |c_i * sc| < 2^-1094 rounds to zero in the three
// accumulates that consume it: the same velocities (up to the sign of an exact zero), one instruction fewer (worth 0.2 %: the kernels are bound by
// register-operand bandwidth of the FP64 unit, not by issue slots, profiles/r01h_fp64_operands.md).
// (Predicating the seed instruction instead -- `@p rsqrt.approx` into a zeroed pair -- is if-converted by ptxas into
// MUFU + FSEL + MOV: no gain, tried.)
#ifndef VLC_GUARD_HI
#define VLC_GUARD_HI 1  // 0 = select both words (r01e and earlier)
#endif
__device__ __forceinline__ void guard_scale(double& sc, double c2) {
#if defined(__CUDA_EMUL__)  // host build of the tests: the same integer compare and high-word select in C++
  if (!(__double_as_longlong(c2) > 0x3970000000000000LL)) {
    long long b = __double_as_longlong(sc) & 0xFFFFFFFFLL;
    std::memcpy(&sc, &b, sizeof sc);
  }
#elif VLC_GUARD_HI
  asm("{\n\t"
      ".reg .pred p;\n\t"
      ".reg .b32 lo, hi;\n\t"
      "setp.gt.s64 p, %1, 0x3970000000000000;\n\t"
      "mov.b64 {lo, hi}, %0;\n\t"
      "selp.b32 hi, hi, 0, p;\n\t"
      "mov.b64 %0, {lo, hi};\n\t"
      "}"
      : "+d"(sc)
      : "l"(__double_as_longlong(c2)));
#else
  asm("{\n\t"
      ".reg .pred p;\n\t"
      "setp.gt.s64 p, %1, 0x3970000000000000;\n\t"
      "selp.f64 %0, %0, 0d0000000000000000, p;\n\t"
      "}"
      : "+d"(sc)
      : "l"(__double_as_longlong(c2)));
#endif
}

struct Src {
  // r0g = G*(p2 - p1), L2g = G*|p2 - p1|^2 with G = gam/(4 pi): the strength is folded into the two
  // per-source quantities that enter the result linearly, which saves one FP64 multiply per pair.
  double p1x, p1y, p1z, p2x, p2y, p2z, r0gx, r0gy, r0gz, L2g, K, spare;
};

__device__ __forceinline__ Src load_src(const double* __restrict__ s) {
  // 6 x 16-byte shared loads; every lane reads the same address -> broadcast.
  const double2* s2 = reinterpret_cast<const double2*>(s);
  double2 a = s2[0], b = s2[1], c = s2[2], d = s2[3], e = s2[4], f = s2[5];
  Src r;
  r.p1x = a.x; r.p1y = a.y; r.p1z = b.x; r.p2x = b.y; r.p2y = c.x; r.p2z = c.y;
  r.r0gx = d.x; r.r0gy = d.y; r.r0gz = e.x; r.L2g = e.y; r.K = f.x; r.spare = f.y;
  return r;
}

// One pair interaction, accumulated into (vx,vy,vz).  43 (FULL) / 41 (FAST) FP64-pipe instructions:
// 6 sub, 6 cross, 3+3+3 squared norms, 3 dot, 1 a2, 1 den, 2 products, 2 rsqrt (5|4 each), 2 combine,
// 3 accumulate.
template <bool FAST>
__device__ __forceinline__ void pair_accumulate(const Src& s, double px, double py, double pz,
                                                double& vx, double& vy, double& vz) {
  const double r1x = px - s.p1x, r1y = py - s.p1y, r1z = pz - s.p1z;
  const double r2x = px - s.p2x, r2y = py - s.p2y, r2z = pz - s.p2z;
  const double cx = fma(r1y, r2z, -(r1z * r2y));
  const double cy = fma(r1z, r2x, -(r1x * r2z));
  const double cz = fma(r1x, r2y, -(r1y * r2x));
  const double c2 = fma(cz, cz, fma(cy, cy, cx * cx));
  const double d1 = fma(r1z, r1z, fma(r1y, r1y, r1x * r1x));
  const double d2 = fma(r2z, r2z, fma(r2y, r2y, r2x * r2x));
  const double a1 = fma(s.r0gz, r1z, fma(s.r0gy, r1y, s.r0gx * r1x));  // G * r0.r1
  const double a2 = a1 - s.L2g;                                         // G * r0.r2 = G*(r0.r1 - |r0|^2)
  const double den = fma(c2, c2, s.K);
  const double w1 = rsqrt_fp64<FAST>(d1 * den);
  const double w2 = rsqrt_fp64<FAST>(d2 * den);
  double sc = fma(a1, w1, -(a2 * w2));
  // classdef.f90:498 `if (r1Xr2Abs2 > eps*eps)`: c2 >= 0, so its bit pattern orders like an integer;
  // eps^2 = 2^-104 = 0x3970000000000000.  One 64-bit integer compare + one select keep the guard off
  // the FP64 pipe.
  guard_scale(sc, c2);
  vx = fma(cx, sc, vx);
  vy = fma(cy, sc, vy);
  vz = fma(cz, sc, vz);
}

// ---- mbarrier + 1-D TMA bulk copy (cp.async.bulk -> SASS UBLKCP) ----
// the thread that issues the bulk copies of a CTA (host build of the tests: every emulated thread stages its own tiles)
#if defined(__CUDA_EMUL__)
#define VLC_PRODUCER(tid) true
#else
#define VLC_PRODUCER(tid) ((tid) == 0)
#endif
#if !defined(__CUDA_EMUL__)
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes,
                                             uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(smem_dst)),
      "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
#else  // __CUDA_EMUL__: no barriers, the bulk copy is a memcpy that has completed when it returns
inline void mbar_init(uint64_t*, uint32_t) {}
inline void fence_mbar_init() {}
inline void mbar_expect_tx(uint64_t*, uint32_t) {}
inline void mbar_wait(uint64_t*, uint32_t) {}
inline void tma_bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t*) { std::memcpy(smem_dst, gmem_src, bytes); }
#endif  // !__CUDA_EMUL__

}  // namespace vlc
