// TEST INFRASTRUCTURE: a stand-in for <cuda_runtime.h> so that g++ can compile the product's O(N) record kernels
// (volcanor_b200/csrc/wake_records.cuh, pack.cuh) for the host and tests/native/kernels_emul.cpp can run them thread by
// thread -- index logic and arithmetic of the actual kernel bodies checked against the oracle without a GPU.  The
// product never sees this file (it is on the include path of tests/native/Makefile only).
#pragma once
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __constant__ static
#define __shared__ static
#define __CUDA_EMUL__ 1
#define __align__(n) alignas(n)
static inline void __syncthreads() {}

struct emul_dim3 {
  unsigned x = 1, y = 1, z = 1;
};
// one emulated thread at a time: the launcher sets these before each call of the kernel body
static emul_dim3 threadIdx, blockIdx, blockDim, gridDim;

struct alignas(16) double2 {
  double x, y;
};
static inline double2 make_double2(double x, double y) { return double2{x, y}; }

// round-to-nearest IEEE operations (the file is compiled with -ffp-contract=off)
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __ddiv_rn(double a, double b) { return a / b; }
static inline double __dsqrt_rn(double a) { return std::sqrt(a); }
static inline long long __double_as_longlong(double a) {
  long long r;
  std::memcpy(&r, &a, sizeof r);
  return r;
}
static inline int __double2hiint(double a) { return (int)(__double_as_longlong(a) >> 32); }
static inline int __double2loint(double a) { return (int)(__double_as_longlong(a) & 0xFFFFFFFFLL); }
static inline double __hiloint2double(int hi, int lo) {
  const unsigned long long b = ((unsigned long long)(unsigned)hi << 32) | (unsigned)lo;
  double r;
  std::memcpy(&r, &b, sizeof r);
  return r;
}
static inline int atomicOr(int* p, int v) {  // one emulated thread at a time
  const int old = *p;
  *p = old | v;
  return old;
}
static inline size_t __cvta_generic_to_shared(const void* p) { return (size_t)p; }

// run kernel(args...) for every thread of a 1-D / 2-D grid, serially, in reverse order (the kernels must not depend on
// the order: a thread reads only what no thread of the same launch writes)
template <class K, class... A>
static void emul_launch(unsigned gx, unsigned gy, unsigned bx, K kernel, A... args) {
  gridDim.x = gx; gridDim.y = gy; blockDim.x = bx;
  for (long long by = (long long)gy - 1; by >= 0; --by)
    for (long long b = (long long)gx - 1; b >= 0; --b)
      for (long long t = (long long)bx - 1; t >= 0; --t) {
        blockIdx.x = (unsigned)b; blockIdx.y = (unsigned)by; threadIdx.x = (unsigned)t;
        kernel(args...);
      }
}
