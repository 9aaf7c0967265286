// pfwake_host.cpp -- TEST INFRASTRUCTURE: the host + device routines of volcanor_b200/csrc/pfwake.cuh (prescribed far
// wake: the fit of pFwake_update, the helix filaments, the axisymmetric copies; wake burst) compiled by g++ and driven with the loops
// the two CUDA kernels run (pf_fit_kernel: one "thread" per convected blade; pf_helix_kernel: one per (blade, filament),
// here in reverse order: the threads are independent), so tests/test_prescribed_wake.py can check arithmetic and index
// logic against the oracle bit for bit without a GPU.  Nothing in the product links or loads this file.
// Build: tests/native/Makefile (g++ -O2 -ffp-contract=off).
#include <vector>

#include "../../volcanor_b200/csrc/pfwake.cuh"

extern "C" {

// = vlc_rotor_updatePrescribedWake on host arrays in the device layouts: waF (nFwake x 13 per blade, blade-major), wapF
// (240 x 13 per blade), helix (2 per blade), T (9 per blade, column-major; blade 0 unused), rotate (flag per blade)
int pf_host_update(int nb, int nbConvect, int axisym, int nFwake, int rowFar, int prescWakeGenNt, double deltaPsi,
                   const double* hub, const double* T, const int* rotate, const double* waF, double* wapF, double* helix) {
  const int rowStart = prescWakeGenNt == 0 ? rowFar : nFwake - prescWakeGenNt;
  if (rowStart < 1 || rowStart > nFwake) return 2;
  std::vector<vlc::pf::Fit> fits(nb);
  for (int ib = 0; ib < nbConvect; ++ib)
    vlc::pf::fit(waF + (size_t)vlc::pf::kFwRec * ((size_t)(rowStart - 1) + (size_t)nFwake * ib), nFwake - rowStart + 1, deltaPsi,
                 hub[2], helix + 2 * ib, &fits[ib]);
  for (int q = nb * vlc::pf::kNpf - 1; q >= 0; --q) {
    const int ib = q / vlc::pf::kNpf, i = q % vlc::pf::kNpf;
    const bool copy = axisym == 1 && ib > 0;
    vlc::pf::blade_filament(ib, i, nbConvect, axisym, fits.data(), copy ? T + 9 * ib : nullptr, copy ? rotate[ib] : 0, hub, wapF,
                            helix);
  }
  return 0;
}

// = rec_burst_kernel, pairs visited in reverse order (the decisions read only end points; every write stores the same value)
void pf_host_burst(int nb, int nFwake, int rowFar, double skewLimit, double largeCoreRadius, double* waF) {
  const int npair = nFwake - rowFar;
  for (int q = nb * (npair > 0 ? npair : 0) - 1; q >= 0; --q) {
    const int ib = q / npair, irow = rowFar + q % npair;
    double* f0 = waF + (size_t)vlc::pf::kFwRec * ((size_t)(irow - 1) + (size_t)nFwake * ib);
    if (vlc::pf::burst_pair(f0, f0 + vlc::pf::kFwRec, skewLimit)) {
      f0[vlc::pf::kFwRec + vlc::pf::kRvc] = largeCoreRadius;
      f0[vlc::pf::kRvc] = largeCoreRadius;
    }
  }
}

// = rec_skew_kernel
void pf_host_skew(int nb, int nbConvect, int axisym, int ns, int nNwake, int rowNear, double* waN) {
  const int nact = nNwake - rowNear + 1;
  const size_t blade = (size_t)nNwake * ns * 50;
  for (long long q = (long long)nb * ns * (nact > 0 ? nact : 0) - 1; q >= 0; --q) {
    const int i = rowNear + (int)(q % nact), j = (int)((q / nact) % ns) + 1, ib = (int)(q / ((long long)nact * ns));
    const int src = vlc::pf::source_blade(ib, nbConvect, axisym);
    if (src < 0) continue;
    const size_t at = 50 * ((size_t)(i - 1) + (size_t)nNwake * (j - 1));
    waN[blade * ib + at + 49] = vlc::pf::ring_skew(waN + blade * src + at);
  }
}

}  // extern "C"
