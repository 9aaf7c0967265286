// bs_sweep.cuh -- the targets x sources Biot-Savart sweep kernel (K1/K2/K3/K5/K9 of SURVEY 2.1).
//
// Replaces the OpenMP target loops of src/libCommon.f90:132-146, :190-195 and the RHS / force loops
// of src/main.f90:528-573, :632-656: every target sums gam * vf_vind over a packed source set.
//
// Decomposition: blockIdx.x = tile of THREADS*T targets (T targets per thread, held in registers),
// blockIdx.y = contiguous chunk of the source list (source split, used when targets are few).
// Each CTA streams its chunk through shared memory in TILE-filament tiles with 1-D TMA bulk copies
// (cp.async.bulk + mbarrier, STAGES-deep ring); all lanes read the same filament -> LDS.128 broadcast.
// Partial sums of the source splits go to part[split][3m]; a fixed-order reduce kernel adds them, so
// results are deterministic for a given (m, n, nsplit).
#pragma once
#include "vlc_device.cuh"

namespace vlc {

template <int T, int THREADS, int TILE, int STAGES, int MINB, bool FAST>
__global__ void __launch_bounds__(THREADS, MINB)
bs_sweep_kernel(const double* __restrict__ src,   // packed sources, padded to a multiple of TILE
                long long chunk,                  // sources per split (multiple of TILE / 4: the last tile of a chunk may be partial)
                long long n_src_padded,           // total padded sources (multiple of TILE)
                const double* __restrict__ P,     // targets (3, m) interleaved
                long long m,
                double* __restrict__ out,         // [gridDim.y][3 m]
                const int* __restrict__ flag,     // optional device flag: run only when (*flag & mask) == want (capi.cu: sweep_shared)
                int mask, int want)
{
  if (flag != nullptr && (*flag & mask) != want) return;
#if defined(__CUDA_EMUL__)  // host build of the tests (tests/native/kernels_emul.cpp), see bs_lattice.cuh
  alignas(128) static unsigned char smem_raw[(size_t)STAGES * TILE * kSrcBytes + STAGES * 8];
#else
  extern __shared__ __align__(128) unsigned char smem_raw[];
#endif
  double* buf = reinterpret_cast<double*>(smem_raw);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + (size_t)STAGES * TILE * kSrcBytes);

  const int tid = threadIdx.x;
  const long long s_begin = (long long)blockIdx.y * chunk;
  long long s_end = s_begin + chunk;
  if (s_end > n_src_padded) s_end = n_src_padded;
  const long long len = s_end > s_begin ? s_end - s_begin : 0;  // a multiple of the granule TILE / 4 (even)
  const int ntiles = (int)((len + TILE - 1) / TILE);
  const double* gsrc = src + s_begin * kSrcDoubles;
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
        const uint32_t bytes = (uint32_t)tile_records(s) * kSrcBytes;
        mbar_expect_tx(&bars[s], bytes);
        tma_bulk_g2s(buf + (size_t)s * TILE * kSrcDoubles, gsrc + (size_t)s * TILE * kSrcDoubles, bytes, &bars[s]);
      }
  }

  // T targets per thread, strided by THREADS so global loads/stores of a warp stay close together.
  const long long t0 = (long long)blockIdx.x * (THREADS * T) + tid;
  double px[T], py[T], pz[T], vx[T], vy[T], vz[T];
#pragma unroll
  for (int k = 0; k < T; ++k) {
    const long long t = t0 + (long long)k * THREADS;
    const bool ok = t < m;
    px[k] = ok ? P[3 * t + 0] : 0.0;
    py[k] = ok ? P[3 * t + 1] : 0.0;
    pz[k] = ok ? P[3 * t + 2] : 0.0;
    vx[k] = vy[k] = vz[k] = 0.0;
  }

  for (int tile = 0; tile < ntiles; ++tile) {
    const int stage = tile % STAGES;
    const uint32_t phase = (uint32_t)(tile / STAGES) & 1u;
    mbar_wait(&bars[stage], phase);
    const double* sb = buf + (size_t)stage * TILE * kSrcDoubles;
    const int jn = tile_records(tile);
#pragma unroll 1
    for (int j2 = 0; j2 < jn; j2 += 2)  // two sources per trip (jn is even): the unroll the fixed-length loop had
#pragma unroll
    for (int jj = 0; jj < 2; ++jj) {
      const int j = j2 + jj;
      const Src s = load_src(sb + j * kSrcDoubles);
#pragma unroll
      for (int k = 0; k < T; ++k) pair_accumulate<FAST>(s, px[k], py[k], pz[k], vx[k], vy[k], vz[k]);
    }
    __syncthreads();  // every warp is done with this stage before it is refilled
    if (VLC_PRODUCER(tid) && tile + STAGES < ntiles) {
      const uint32_t bytes = (uint32_t)tile_records(tile + STAGES) * kSrcBytes;
      mbar_expect_tx(&bars[stage], bytes);
      tma_bulk_g2s(buf + (size_t)stage * TILE * kSrcDoubles, gsrc + (size_t)(tile + STAGES) * TILE * kSrcDoubles, bytes,
                   &bars[stage]);
    }
  }

  // FAST mode: the second-order Newton step leaves every rsqrt low by a factor in [1 - 1.5 d^2, 1]
  // (d = seed error <= kSeedRelErr); centre that bias once per target.
  constexpr double kCentre = FAST ? (1.0 + 0.75 * kSeedRelErr * kSeedRelErr) : 1.0;
  double* o = out + (size_t)blockIdx.y * 3 * (size_t)m;
#pragma unroll
  for (int k = 0; k < T; ++k) {
    const long long t = t0 + (long long)k * THREADS;
    if (t < m) {
      o[3 * t + 0] = FAST ? vx[k] * kCentre : vx[k];
      o[3 * t + 1] = FAST ? vy[k] * kCentre : vy[k];
      o[3 * t + 2] = FAST ? vz[k] * kCentre : vz[k];
    }
  }
}

// Fixed-order sum of the source-split partials: V[i] = ((part[0][i] + part[1][i]) + ...).
__global__ void bs_reduce_kernel(const double* __restrict__ part, int nsplit, long long len, double* __restrict__ V) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= len) return;
  double a = part[i];  // slots added in order; loads of eight slots in flight (see bs_reduce_select_kernel)
  int s = 1;
  for (; s + 8 <= nsplit; s += 8) {
    double v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = part[(size_t)(s + k) * len + i];
#pragma unroll
    for (int k = 0; k < 8; ++k) a += v[k];
  }
  for (; s < nsplit; ++s) a += part[(size_t)s * len + i];
  V[i] = a;
}

}  // namespace vlc
