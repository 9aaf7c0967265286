// kernels_emul.cpp -- TEST INFRASTRUCTURE: the product's O(N) record kernels (volcanor_b200/csrc/wake_records.cuh, the
// arithmetic of pfwake.cuh) compiled by g++ against a stand-in for <cuda_runtime.h> (tests/native/emul) and run thread by
// thread, serially and in REVERSE thread order, with the launch shapes volcanor_b200/csrc/capi.cu uses.  What runs is
// the kernel body itself -- its index decoding, its guards, its unfused arithmetic -- so tests/test_kernels_emul.py can
// hold the kernels against the oracle bit for bit without a GPU (cos / sin / atan2 / acos are then libm's on both
// sides).  Nothing in the product links or loads this file.  Build: tests/native/Makefile (g++ -O2 -ffp-contract=off).
#include <algorithm>
#include <cmath>
#include <vector>

#include "../../volcanor_b200/csrc/wake_records.cuh"
#include "../../volcanor_b200/csrc/bs_lattice.cuh"
#include "../../volcanor_b200/csrc/bs_sweep.cuh"

namespace {
inline unsigned blocks_for(long long n, int t) { return (unsigned)((n + t - 1) / t); }
}  // namespace

extern "C" {

// = vlc_rotor_age_wake
void emul_age_wake(int nb, int ns, int nNwake, int nFwake, int rowNear, int rowFar, double dt, double omegaSlow, double* waN,
                   double* waF) {
  const long long nact = std::max(0, nNwake - rowNear + 1), nfar = std::max(0, nFwake - rowFar + 1);
  const long long n = (long long)nb * (ns * nact + nfar);
  if (n <= 0) return;
  emul_launch(blocks_for(n, 256), 1, 256, vlc::rec_age_kernel, nb, ns, nNwake, nFwake, rowNear, rowFar, dt, dt * omegaSlow, waN, waF);
}

// = vlc_rotor_dissipate_wake
void emul_dissipate_wake(int nb, int ns, int nNwake, int nFwake, int rowNear, int rowFar, double apparentViscCoeff,
                         double decayCoeff, double dt, double kinematicVisc, double* waN, double* waF) {
  const double growTerm = 4.0 * 1.2564 * apparentViscCoeff * kinematicVisc * dt;
  const double decayFactor = std::exp(-decayCoeff * dt);
  const long long nact = std::max(0, nNwake - rowNear + 1), nfar = std::max(0, nFwake - rowFar + 1);
  const long long n = (long long)nb * (ns * nact + nfar);
  if (n <= 0) return;
  emul_launch(blocks_for(n, 256), 1, 256, vlc::rec_dissipate_kernel, nb, ns, nNwake, nFwake, rowNear, rowFar, growTerm, decayFactor,
              waN, waF);
  const long long n4 = (long long)nb * ns * (nact - 1);
  if (n4 > 0) emul_launch(blocks_for(n4, 256), 1, 256, vlc::rec_dissipate_vf4_kernel, nb, ns, nNwake, rowNear, waN);
}

// = vlc_rotor_strain_wake
void emul_strain_wake(int nb, int nFwake, int rowFar, double* waF) {
  const int nfar = nFwake - rowFar + 1;
  if (nfar <= 0) return;
  emul_launch(blocks_for((long long)nb * nfar, 128), 1, 128, vlc::rec_strain_kernel, nb, nFwake, rowFar, waF);
}

// = vlc_rotor_burst_wake
void emul_burst_wake(int nb, int nFwake, int rowFar, double skewLimit, double largeCoreRadius, double* waF) {
  const long long n = (long long)nb * std::max(0, nFwake - rowFar);
  if (n <= 0) return;
  emul_launch(blocks_for(n, 128), 1, 128, vlc::rec_burst_kernel, nb, nFwake, rowFar, skewLimit, largeCoreRadius, waF);
}

// = vlc_rotor_calc_skew
void emul_calc_skew(int nb, int nbConvect, int axisym, int ns, int nNwake, int rowNear, double* waN) {
  const long long n = (long long)nb * ns * std::max(0, nNwake - rowNear + 1);
  if (n <= 0) return;
  emul_launch(blocks_for(n, 128), 1, 128, vlc::rec_skew_kernel, nb, nbConvect, axisym, ns, nNwake, rowNear, waN);
}

// = vlc_rotor_updatePrescribedWake; T9 = 9 doubles per blade (column-major, blade 0 unused), rotate = flag per blade
int emul_updatePrescribedWake(int nb, int nbConvect, int axisym, int nFwake, int rowFar, int prescWakeGenNt, double deltaPsi,
                              const double* hub, const double* T9, const int* rotate, const double* waF, double* wapF,
                              double* helix) {
  const int rowStart = prescWakeGenNt == 0 ? rowFar : nFwake - prescWakeGenNt;
  if (nFwake <= 0 || rowStart < 1 || rowStart > nFwake) return 2;
  std::vector<vlc::AxiT> Ts(nb);
  for (int ib = 0; ib < nb; ++ib) {
    for (int k = 0; k < 9; ++k) Ts[ib].T[k] = T9[9 * ib + k];
    Ts[ib].rotate = rotate[ib];
  }
  std::vector<vlc::pf::Fit> fits(nb);
  const bool copies = axisym == 1 && nb > 1;
  emul_launch(blocks_for(nbConvect, 32), 1, 32, vlc::pf_fit_kernel, nbConvect, nFwake, rowStart, nFwake - rowStart + 1, deltaPsi,
              hub[2], waF, helix, fits.data());
  emul_launch(blocks_for((long long)nb * 240, 128), 1, 128, vlc::pf_helix_kernel, nb, nbConvect, axisym,
              (const vlc::pf::Fit*)fits.data(), copies ? (const vlc::AxiT*)Ts.data() : (const vlc::AxiT*)nullptr, hub[0], hub[1],
              hub[2], wapF, helix);
  return 0;
}

// = vlc_rotor_assignshed
void emul_assignshed(int edge, int nb, int nc, int ns, int nNwake, int rowNear, const double* wiP, double* waN) {
  emul_launch(blocks_for((long long)nb * ns, 128), 1, 128, vlc::rec_assignshed_kernel, edge, nb, nc, ns, nNwake, rowNear, wiP, waN);
}

// = vlc_rotor_convectwake (without the prescribed far wake: emul_updatePrescribedWake follows it like in g_convect)
void emul_convectwake(int predicted, int nb, int nbConvect, int axisym, int duct, int ns, int nNwake, int nFwake, int rowNear,
                      int rowFar, double dt, const double* hub, const double* T9, const int* rotate, const double* velN,
                      const double* velF, double* waN, double* waF) {
  const long long nfar = std::max(0, nFwake - rowFar + 1), nact = std::max(0, nNwake - rowNear + 1);
  {
    const long long n = (long long)nbConvect * ((long long)(ns + 1) * nNwake + nfar);
    emul_launch(blocks_for(n, 256), 1, 256, vlc::rec_convect_kernel, predicted, nbConvect, ns, nNwake, nFwake, rowNear, rowFar, dt,
                velN, velF, waN, waF);
  }
  {
    const long long n = (long long)nbConvect * (ns * nact + std::max(0LL, nfar - 1));
    if (n > 0) {
      emul_launch(blocks_for(n, 256), 1, 256, vlc::rec_continuity_kernel, 0, nbConvect, ns, nNwake, nFwake, rowNear, rowFar, waN, waF);
      if (!predicted && duct == 1)
        emul_launch(blocks_for(n, 256), 1, 256, vlc::rec_continuity_kernel, 1, nbConvect, ns, nNwake, nFwake, rowNear, rowFar, waN,
                    waF);
    }
  }
  if (axisym == 1 && nb > 1) {
    std::vector<vlc::AxiT> Ts(nb);
    for (int ib = 0; ib < nb; ++ib) {
      for (int k = 0; k < 9; ++k) Ts[ib].T[k] = T9[9 * ib + k];
      Ts[ib].rotate = rotate[ib];
    }
    const long long n = (long long)(nb - 1) * (ns * nact + nfar);
    if (n > 0)
      emul_launch(blocks_for(n, 128), 1, 128, vlc::rec_axisym_kernel, nb, ns, nNwake, nFwake, rowNear, rowFar,
                  (const vlc::AxiT*)Ts.data(), hub[0], hub[1], hub[2], waN, waF);
  }
}

// = vlc_rotor_rollup: waN is replaced by its shifted copy like the swap of the two device buffers
void emul_rollup(int nb, int ns, int nNwake, int nFwake, int rowFar, int rollupStart, int rollupEnd, int sgnPositive,
                 int suppressFwake, double* waN, double* waF) {
  int rowFarNext = rowFar - 1;
  if (nFwake > 0 && rowFarNext == 0) {
    emul_launch(blocks_for(nb * vlc::kFw, 64), 1, 64, vlc::rec_shiftFwake_kernel, nb, nFwake, waF);
    rowFarNext = 1;
  }
  emul_launch(blocks_for(nb, 32), 1, 32, vlc::rec_rollup_kernel, nb, ns, nNwake, nFwake, rollupStart, rollupEnd, sgnPositive,
              suppressFwake, rowFarNext, (const double*)waN, waF);
  const size_t total = (size_t)nb * nNwake * ns * vlc::kVr;
  std::vector<double> alt(total);
  emul_launch(blocks_for((long long)total, 256), 1, 256, vlc::rec_shiftwake_kernel, (long long)nb * ns, nNwake, (const double*)waN,
              alt.data());
  std::memcpy(waN, alt.data(), total * sizeof(double));
}

// ---- source packing (pack.cuh) + the pair arithmetic of the sweeps (vlc_device.cuh: pair_accumulate with an exact seed
// in place of MUFU.RSQ64H), summed record by record in enumeration order: what a flat sweep computes, up to summation order

// = pack_bound (what = 3) / pack_chord (what = 4) of capi.cu; rec holds (2*nc*ns + ns)*nb records of 12 doubles
void emul_pack_wing_subset(int what, int nb, int nc, int ns, const double* wiP, double* rec) {
  const long long per_blade = 2LL * nc * ns + ns, cnt = 2LL * nc * ns, wiP_blade = (long long)nc * ns * vlc::kWp;
  const int mask = what == 3 ? 0xA : 0x5;
  const double te_sign = what == 3 ? -1.0 : 1.0;
  emul_launch(blocks_for(cnt, 256), (unsigned)nb, 256, vlc::pack_rings_kernel, wiP, vlc::kWp, nc, 0, nc, ns, mask, 2, 1.0, 0, rec,
              wiP_blade, per_blade);
  emul_launch(blocks_for(ns, 128), (unsigned)nb, 128, vlc::pack_rings_kernel, wiP, vlc::kWp, nc, nc - 1, 1, ns, 0x2, 1, te_sign, 0,
              rec + (size_t)cnt * vlc::kSrcDoubles, wiP_blade, per_blade);
}

// = the prescribed-wake part of pack_rotor: 240 records per blade
void emul_pack_pfwake(int nb, const double* wapF, double* rec) {
  emul_launch(blocks_for(240, 128), (unsigned)nb, 128, vlc::pack_fwake_kernel, wapF, 0, 240, rec, (long long)240 * vlc::kFw,
              (long long)240);
}

void emul_vind_records(long long n, const double* rec, long long m, const double* P, double* V) {
  for (long long t = 0; t < m; ++t) {
    double vx = 0.0, vy = 0.0, vz = 0.0;
    for (long long k = 0; k < n; ++k) {
      const vlc::Src s = vlc::load_src(rec + k * vlc::kSrcDoubles);
      vlc::pair_accumulate<false>(s, P[3 * t], P[3 * t + 1], P[3 * t + 2], vx, vy, vz);
    }
    V[3 * t] = vx;
    V[3 * t + 1] = vy;
    V[3 * t + 2] = vz;
  }
}

// ---- the dominant kernel: bs_lattice_kernel<W, T, 128, 3, .> on the strip records pack_rings_shared_kernel<W> makes from
// one blade's near-wake ring records (rows i0 .. i0+nrows-1 of waN(nNwake, ns), ns a multiple of W), one split; plus the
// flat remainder (the last column's outer streamwise edges, pack_rings_kernel mask 0x4 with the wake rule) through the
// pair arithmetic.  V = what vind_bywake gives for a blade without a far wake, regrouped node by node and edge by edge.
}  // extern "C"

// strips of width W over ring columns col0 .. (clipped at ns by fill_strip_record): pack + null padding + sweep with nsplit
// source splits (grid y, chunks of whole granules as sweep_shared cuts them); the partial slots are appended to parts
template <int W, int T>
static int lattice_strips(const double* waN, int nNwake, int ns, int i0, int nrows, int col0, int nstrips, int nsplit, long long m,
                          const double* P, int* unmergeable, std::vector<double>& parts, int* nslots) {
  constexpr int RD = vlc::lat_rec_doubles(W), TILE = vlc::lat_tile(W), THREADS = 128;
  const long long nrec = (long long)nstrips * (nrows + 1), npad = (nrec + TILE - 1) / TILE * TILE;
  std::vector<double> lat((size_t)npad * RD);
  emul_launch(blocks_for(nrec, 128), 1, 128, vlc::pack_rings_shared_kernel<W>, waN, vlc::kVr, nNwake, i0, nrows, ns, col0, nstrips,
              lat.data(), unmergeable, 0LL);
  if (npad > nrec)
    emul_launch(blocks_for(npad - nrec, 128), 1, 128, vlc::pack_null_lat_kernel<W>, npad - nrec, lat.data() + (size_t)nrec * RD);
  if (*unmergeable & 1) return 1;
  // chunks are multiples of the granule (a quarter tile), as sweep_shared cuts a tuned or a small sweep: with nsplit that does
  // not divide the tiles the last tile of a chunk is partial
  constexpr int GRAN = vlc::lat_granule(W);
  const long long units = npad / GRAN, per = (units + nsplit - 1) / nsplit, chunk = per * GRAN;
  const int real_split = (int)((units + per - 1) / per);
  const size_t len = 3 * (size_t)m, at = parts.size();
  parts.resize(at + (size_t)real_split * len, 0.0);
  // both forms are launched, as on the device; the set's flag (0 merged / 2 dual) decides which one does the work
  std::vector<double> other((size_t)real_split * len, std::nan(""));
  const bool dual = (*unmergeable == 2);
  emul_launch(blocks_for(m, THREADS * T), (unsigned)real_split, THREADS, vlc::bs_lattice_kernel<W, T, THREADS, 3, 1>,
              (const double*)lat.data(), chunk, npad, P, m, dual ? other.data() : parts.data() + at,
              (const int*)unmergeable, 3, 0, 0);
  emul_launch(blocks_for(m, THREADS), (unsigned)real_split, THREADS, vlc::bs_lattice_kernel<W, 1, THREADS, 3, 1, true>,
              (const double*)lat.data(), chunk, npad, P, m, dual ? parts.data() + at : other.data(),
              (const int*)unmergeable, 3, 2, 0);
  for (double v : other)
    if (v == v) return 4;  // the form that must not run wrote something
  *nslots += real_split;
  return 0;
}

static int lattice_dispatch(int W, int T, const double* waN, int nNwake, int ns, int i0, int nrows, int col0, int nstrips,
                            int nsplit, long long m, const double* P, int* unmergeable, std::vector<double>& parts, int* nslots) {
#define X(WW, TT) \
  if (W == WW && T == TT) \
    return lattice_strips<WW, TT>(waN, nNwake, ns, i0, nrows, col0, nstrips, nsplit, m, P, unmergeable, parts, nslots);
  X(1, 1) X(1, 3) X(2, 2) X(3, 1) X(3, 2) X(4, 1) X(4, 2)
#undef X
  return 3;
}

extern "C" {

// The strip plan of capi.cu (plan_strips): tailW = 0: ceil(ns / W) strips of width W (the last one partial when ns is not a
// multiple of W); tailW = ns mod W > 0: floor(ns / W) strips of width W + one tail strip of width tailW swept with
// kLatBestT[tailW] targets per thread (sweep_shared's second lattice launch).  Launch sequence of pack_rotor_sources /
// sweep_shared for one blade: check_rings_kernel, the strip packs, the lattice launches with nsplit source splits, the flat
// remainder (one slot), bs_reduce_select_kernel over the slots of the mergeable path.  1 = not a lattice (flag raised).
int emul_lattice_vind_split(int W, int T, int tailW, int nsplit, const double* waN, int nNwake, int ns, int i0, int nrows,
                            long long m, const double* P, double* V) {
  static const int bestT[4] = {0, 3, 2, 2};
  if (tailW < 0 || tailW > 3 || (tailW && (ns <= W || ns % W != tailW)) || nsplit < 1) return 2;
  const int nmain = tailW ? ns / W : (ns + W - 1) / W;
  int unmergeable = 0, nslots = 0;
  emul_launch(blocks_for((long long)nrows * ns, 256), 1, 256, vlc::check_rings_kernel, waN, vlc::kVr, nNwake, i0, nrows, ns,
              &unmergeable, 0LL);
  std::vector<double> parts;
  int rc = lattice_dispatch(W, T, waN, nNwake, ns, i0, nrows, 0, nmain, nsplit, m, P, &unmergeable, parts, &nslots);
  if (rc) return rc;
  if (tailW && (rc = lattice_dispatch(tailW, bestT[tailW], waN, nNwake, ns, i0, nrows, nmain * W, 1, 1, m, P, &unmergeable, parts,
                                      &nslots)))
    return rc;
  const size_t len = 3 * (size_t)m;
  std::vector<double> rem((size_t)nrows * vlc::kSrcDoubles);
  emul_launch(blocks_for(nrows, 128), 1, 128, vlc::pack_rings_kernel, waN + (size_t)vlc::kVr * nNwake * (ns - 1), vlc::kVr, nNwake, i0,
              nrows, 1, 0x4, 1, 1.0, 1, rem.data(), 0LL, 0LL);
  std::vector<double> remv(len);
  emul_vind_records(nrows, rem.data(), m, P, remv.data());
  // slot layout of sweep_shared: [merged | remainder | dual | flat]; the lattice slots above belong to the form the flag names
  std::vector<double> all;
  if (unmergeable == 2) {
    all = remv;
    all.insert(all.end(), parts.begin(), parts.end());
    emul_launch(blocks_for((long long)len, 256), 1, 256, vlc::bs_reduce_select_kernel, (const double*)all.data(),
                (const int*)&unmergeable, 0, 1, nslots, 0, (long long)len, V);
  } else {
    all = parts;
    all.insert(all.end(), remv.begin(), remv.end());
    emul_launch(blocks_for((long long)len, 256), 1, 256, vlc::bs_reduce_select_kernel, (const double*)all.data(),
                (const int*)&unmergeable, nslots, 1, 0, 0, (long long)len, V);
  }
  return unmergeable == 2 ? -2 : 0;  // -2: done, by the dual form
}

int emul_lattice_vind_plan(int W, int T, int tailW, const double* waN, int nNwake, int ns, int i0, int nrows, long long m,
                           const double* P, double* V) {
  return emul_lattice_vind_split(W, T, tailW, 1, waN, nNwake, ns, i0, nrows, m, P, V);
}

// Device-side dispatch of sweep_shared for one blade, every launch made whatever the flag says (as on the device, where the
// host never reads it): check_rings + strip pack raise *flag_out when the records are not a lattice; the W = 4 lattice launch
// and the flat remainder run with want = 0, the flat enumeration (pack_rings_kernel, 4 filaments per ring, wake rule) with
// want = 1 and nsplit_flat source splits; bs_reduce_select_kernel sums the slots of the path that ran.  The partial buffer
// starts as NaN: a selected slot that nobody wrote shows up in V.
int emul_lattice_vind_dispatch(int nsplit_flat, const double* waN, int nNwake, int ns, int i0, int nrows, long long m,
                               const double* P, double* V, int* flag_out) {
  constexpr int W = 4, T = 2, THREADS = 128, FT = 4, FTILE = 128;
  constexpr int RD = vlc::lat_rec_doubles(W), TILE = vlc::lat_tile(W);
  if (nsplit_flat < 1 || m <= 0) return 2;
  int flag = 0;
  emul_launch(blocks_for((long long)nrows * ns, 256), 1, 256, vlc::check_rings_kernel, waN, vlc::kVr, nNwake, i0, nrows, ns, &flag, 0LL);
  const int nstrips = (ns + W - 1) / W;
  const long long nrec = (long long)nstrips * (nrows + 1), npad = (nrec + TILE - 1) / TILE * TILE;
  std::vector<double> lat((size_t)npad * RD);
  emul_launch(blocks_for(nrec, 128), 1, 128, vlc::pack_rings_shared_kernel<W>, waN, vlc::kVr, nNwake, i0, nrows, ns, 0, nstrips,
              lat.data(), &flag, 0LL);
  if (npad > nrec)
    emul_launch(blocks_for(npad - nrec, 128), 1, 128, vlc::pack_null_lat_kernel<W>, npad - nrec, lat.data() + (size_t)nrec * RD);
  auto flat_records = [&](const double* base, int cols, int mask, int per_ring, std::vector<double>& rec) -> long long {
    const long long n = (long long)per_ring * nrows * cols, pad = std::max(1LL, (n + FTILE - 1) / FTILE) * FTILE;
    rec.assign((size_t)pad * vlc::kSrcDoubles, 0.0);
    emul_launch(blocks_for(n, 256), 1, 256, vlc::pack_rings_kernel, base, vlc::kVr, nNwake, i0, nrows, cols, mask, per_ring, 1.0, 1,
                rec.data(), 0LL, 0LL);
    if (pad > n) emul_launch(blocks_for(pad - n, 256), 1, 256, vlc::pack_null_kernel, pad - n, rec.data() + (size_t)n * vlc::kSrcDoubles);
    return pad;
  };
  std::vector<double> rem, flat;
  const long long rem_pad = flat_records(waN + (size_t)vlc::kVr * nNwake * (ns - 1), 1, 0x4, 1, rem);
  const long long flat_pad = flat_records(waN, ns, 0xF, 4, flat);
  const long long ftiles = flat_pad / FTILE, fchunk_tiles = (ftiles + nsplit_flat - 1) / nsplit_flat;
  const int nb = (int)((ftiles + fchunk_tiles - 1) / fchunk_tiles);
  const size_t len = 3 * (size_t)m;
  // slots: [0] merged lattice, [1] flat remainder, [2] dual lattice, [3 ..] flat enumeration
  std::vector<double> parts((size_t)(3 + nb) * len, std::nan(""));
  emul_launch(blocks_for(m, THREADS * T), 1, THREADS, vlc::bs_lattice_kernel<W, T, THREADS, 3, 1>, (const double*)lat.data(), npad, npad,
              P, m, parts.data(), (const int*)&flag, 3, 0, 0);
  emul_launch(blocks_for(m, THREADS), 1, THREADS, vlc::bs_lattice_kernel<W, 1, THREADS, 3, 1, true>, (const double*)lat.data(), npad,
              npad, P, m, parts.data() + 2 * len, (const int*)&flag, 3, 2, 0);
  emul_launch(blocks_for(m, THREADS * FT), 1, THREADS, vlc::bs_sweep_kernel<FT, THREADS, FTILE, 3, 1, false>, (const double*)rem.data(),
              rem_pad, rem_pad, P, m, parts.data() + len, (const int*)&flag, 1, 0);
  emul_launch(blocks_for(m, THREADS * FT), (unsigned)nb, THREADS, vlc::bs_sweep_kernel<FT, THREADS, FTILE, 3, 1, false>,
              (const double*)flat.data(), fchunk_tiles * FTILE, flat_pad, P, m, parts.data() + 3 * len, (const int*)&flag, 1, 1);
  emul_launch(blocks_for((long long)len, 256), 1, 256, vlc::bs_reduce_select_kernel, (const double*)parts.data(), (const int*)&flag, 1,
              1, 1, nb, (long long)len, V);
  *flag_out = flag;
  return 0;
}

int emul_lattice_vind(int W, int T, const double* waN, int nNwake, int ns, int i0, int nrows, long long m, const double* P,
                      double* V) {
  if (ns % W != 0) return 2;
  return emul_lattice_vind_plan(W, T, 0, waN, nNwake, ns, i0, nrows, m, P, V);
}

// ---- the flat sweep: pack_flat_kernel (tier 1 arrays -> records, padded with null filaments to whole tiles) ->
// bs_sweep_kernel<4, 128, 128, 3, ., FAST> with nsplit source splits -> bs_reduce_kernel (fixed-order sum of the partials)
int emul_flat_sweep(int fast, int nsplit, long long n, const double* p1, const double* p2, const double* rvc, const double* gam,
                    const unsigned char* wake_flag, long long m, const double* P, double* V) {
  constexpr int T = 4, THREADS = 128, TILE = 128;
  if (nsplit < 1 || n < 0 || m <= 0) return 2;
  const long long npad = std::max(1LL, (n + TILE - 1) / TILE) * TILE;
  std::vector<double> rec((size_t)npad * vlc::kSrcDoubles);
  emul_launch(blocks_for(npad, 256), 1, 256, vlc::pack_flat_kernel, n, npad, p1, p2, rvc, gam, wake_flag, rec.data());
  // chunks are multiples of the granule (a quarter tile), as plan_flat cuts a tuned or a small sweep: the last tile of a chunk
  // may be partial
  constexpr int GRAN = TILE / 4;
  const long long units = npad / GRAN, per = (units + nsplit - 1) / nsplit, chunk = per * GRAN;
  nsplit = (int)((units + per - 1) / per);
  std::vector<double> part((size_t)nsplit * 3 * (size_t)m);
  if (fast)
    emul_launch(blocks_for(m, THREADS * T), (unsigned)nsplit, THREADS, vlc::bs_sweep_kernel<T, THREADS, TILE, 3, 1, true>,
                (const double*)rec.data(), chunk, npad, P, m, part.data(), (const int*)nullptr, 0, 0);
  else
    emul_launch(blocks_for(m, THREADS * T), (unsigned)nsplit, THREADS, vlc::bs_sweep_kernel<T, THREADS, TILE, 3, 1, false>,
                (const double*)rec.data(), chunk, npad, P, m, part.data(), (const int*)nullptr, 0, 0);
  emul_launch(blocks_for(3 * m, 256), 1, 256, vlc::bs_reduce_kernel, (const double*)part.data(), nsplit, 3 * m, V);
  return 0;
}

// rsqrt_fp64<FAST> of vlc_device.cuh (seed modelled, refinement = the product's source)
void emul_rsqrt(int fast, long long n, const double* x, double* y) {
  for (long long i = 0; i < n; ++i) y[i] = fast ? vlc::rsqrt_fp64<true>(x[i]) : vlc::rsqrt_fp64<false>(x[i]);
}

// = vlc_rotor_wakevel_lincomb on one array (the caller passes the near- or the far-wake arrays)
void emul_lincomb(long long n, int nterms, const double* s0, const double* s1, const double* s2, const double* s3,
                  const double* coef, double divisor, double* dst) {
  if (n <= 0) return;
  double cf[4] = {0, 0, 0, 0};
  for (int k = 0; k < nterms; ++k) cf[k] = coef[k];
  emul_launch(blocks_for(n, 256), 1, 256, vlc::rec_lincomb_kernel, n, nterms, s0, s1, s2, s3, cf[0], cf[1], cf[2], cf[3], divisor, dst);
}

}  // extern "C"
