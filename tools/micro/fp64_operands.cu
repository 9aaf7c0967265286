// fp64_operands.cu -- what does one FP64 instruction cost on sm_100a as a function of where its operands come from?
// Register-resident chains, 8 independent per thread, all SMs full; prints TFLOP/s-equivalent (DFMA = 2 flop, DMUL /
// DADD counted as 2 as well so that every line reads "instructions per second x 2").  profiles/r01h_fp64_operands.md.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/fp64_operands tools/micro/fp64_operands.cu
#include <cstdio>
#include <cuda_runtime.h>
constexpr int K = 8, U = 16;
#define INIT                                                      \
  double a[K], b[K], c[K];                                        \
  _Pragma("unroll") for (int k = 0; k < K; ++k) {                 \
    a[k] = 1e-3 * (k + 1) + threadIdx.x * 1e-9;                   \
    b[k] = 0.5 + 1e-2 * k + threadIdx.x * 1e-7;                   \
    c[k] = 1e-4 * (threadIdx.x + 1) + 1e-5 * k;                   \
  }
#define FINI                                                      \
  double s = 0.0;                                                 \
  _Pragma("unroll") for (int k = 0; k < K; ++k) s += a[k] + b[k] + c[k]; \
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
// 0: a = a*const + loopinv   (one changing register operand)
__global__ void p0(int iters, double* out) { INIT; const double cc = c[0];
  for (int it = 0; it < iters; ++it) _Pragma("unroll") for (int u = 0; u < U; ++u) _Pragma("unroll") for (int k = 0; k < K; ++k) a[k] = fma(a[k], 1.0000001, cc);
  FINI }
// 1: three distinct changing registers
__global__ void p1(int iters, double* out) { INIT;
  for (int it = 0; it < iters; ++it) _Pragma("unroll") for (int u = 0; u < U / 4; ++u) {
    _Pragma("unroll") for (int k = 0; k < K; ++k) a[k] = fma(b[k], c[k], a[k]);
    _Pragma("unroll") for (int k = 0; k < K; ++k) b[k] = fma(c[k], a[k], b[k]);
    _Pragma("unroll") for (int k = 0; k < K; ++k) c[k] = fma(a[k], b[k], c[k]);
    _Pragma("unroll") for (int k = 0; k < K; ++k) a[k] = fma(c[k], b[k], -a[k]); }
  FINI }
// 2: three registers, slot A shared by consecutive instructions (reuse cache candidate)
__global__ void p2(int iters, double* out) { INIT;
  for (int it = 0; it < iters; ++it) _Pragma("unroll") for (int u = 0; u < U / 2; ++u) {
    const double s0 = b[u % K];
    _Pragma("unroll") for (int k = 0; k < K; ++k) a[k] = fma(s0, c[k], a[k]);
    const double s1 = a[u % K];
    _Pragma("unroll") for (int k = 0; k < K; ++k) c[k] = fma(s1, a[k], c[k]); }
  FINI }
// 3: same register in two slots: a = b*b + a
__global__ void p3(int iters, double* out) { INIT;
  for (int it = 0; it < iters; ++it) _Pragma("unroll") for (int u = 0; u < U / 2; ++u) {
    _Pragma("unroll") for (int k = 0; k < K; ++k) a[k] = fma(b[k], b[k], a[k]);
    _Pragma("unroll") for (int k = 0; k < K; ++k) b[k] = fma(a[k], a[k], -b[k]); }
  FINI }
// 4: DMUL, two distinct changing registers
__global__ void p4(int iters, double* out) { INIT;
  for (int it = 0; it < iters; ++it) _Pragma("unroll") for (int u = 0; u < U / 2; ++u) {
    _Pragma("unroll") for (int k = 0; k < K; ++k) a[k] = b[k] * c[k];
    _Pragma("unroll") for (int k = 0; k < K; ++k) c[k] = a[k] * b[k]; }
  FINI }
// 5: two changing registers + immediate: a = b*1.5 + a
__global__ void p5(int iters, double* out) { INIT;
  for (int it = 0; it < iters; ++it) _Pragma("unroll") for (int u = 0; u < U / 2; ++u) {
    _Pragma("unroll") for (int k = 0; k < K; ++k) a[k] = fma(b[k], 0.75, a[k]);
    _Pragma("unroll") for (int k = 0; k < K; ++k) b[k] = fma(a[k], 0.25, -b[k]); }
  FINI }
// 6: y + y*e  (register y in slots A and C): a = a*b + a
__global__ void p6(int iters, double* out) { INIT;
  for (int it = 0; it < iters; ++it) _Pragma("unroll") for (int u = 0; u < U / 2; ++u) {
    _Pragma("unroll") for (int k = 0; k < K; ++k) a[k] = fma(a[k], c[k], a[k]);
    _Pragma("unroll") for (int k = 0; k < K; ++k) c[k] = fma(c[k], b[k], c[k]); }
  FINI }
// 7: alternate a 3-register DFMA with a 1-register DMUL (x*x): is the register bandwidth averaged over instructions?
__global__ void p7(int iters, double* out) { INIT;
  for (int it = 0; it < iters; ++it) _Pragma("unroll") for (int u = 0; u < U / 2; ++u) {
    _Pragma("unroll") for (int k = 0; k < K; ++k) { a[k] = fma(b[k], c[k], a[k]); b[k] = b[k] * b[k]; } }
  FINI }
// 8: DADD two distinct registers
__global__ void p8(int iters, double* out) { INIT;
  for (int it = 0; it < iters; ++it) _Pragma("unroll") for (int u = 0; u < U / 2; ++u) {
    _Pragma("unroll") for (int k = 0; k < K; ++k) a[k] = b[k] + c[k];
    _Pragma("unroll") for (int k = 0; k < K; ++k) c[k] = a[k] - b[k]; }
  FINI }
typedef void (*kern_t)(int, double*);
int main() {
  cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
  const int blocks = pr.multiProcessorCount * 8, threads = 256, iters = 20000;
  double* out; cudaMalloc(&out, sizeof(double) * blocks * threads);
  kern_t ks[] = {p0, p1, p2, p3, p4, p5, p6, p7, p8};
  const char* names[] = {"0 DFMA a*imm+loopinv (1 changing reg)", "1 DFMA three distinct changing regs", "2 DFMA three regs, slot A shared by consecutive instr",
                         "3 DFMA b*b+a (same reg in two slots)", "4 DMUL two distinct regs", "5 DFMA b*imm+a (two regs + immediate)",
                         "6 DFMA a*c+a (same reg in slots A and C)", "7 DFMA(3 regs) alternating with DMUL x*x", "8 DADD two distinct regs"};
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int i = 0; i < 9; ++i) {
    ks[i]<<<blocks, threads>>>(iters / 4, out);
    cudaEventRecord(e0); ks[i]<<<blocks, threads>>>(iters, out); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double inst = (double)blocks * threads * iters * K * U;
    printf("%-55s %8.2f ms  %6.2f T(2*instr)/s\n", names[i], ms, inst * 2 / (ms * 1e-3) / 1e12);
  }
  printf("cudaGetLastError: %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
