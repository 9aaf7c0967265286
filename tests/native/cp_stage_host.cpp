// cp_stage_host.cpp -- TEST INFRASTRUCTURE: the host + device routines of volcanor_b200/csrc/cp_stage.cuh (per-section
// loads, blade sums, right-hand-side entries, map_gam) compiled by g++ and driven with the loops the CUDA kernels run
// (one "thread" per section / entry), so tests/test_cp_stage_host.py can check the arithmetic and the index logic against
// the oracle bit for bit without a GPU.  Nothing in the product links or loads this file.
// Build: tests/native/Makefile (g++ -O2 -ffp-contract=off).
#include <cmath>
#include <vector>

#include "../../volcanor_b200/csrc/cp_stage.cuh"

extern "C" {

// = cp_loads_kernel for blades 0..nbConvect-1 (wiP, sec, loads are the blade-major device layouts)
void cp_host_loads(int nbConvect, int nc, int ns, double density, double dt, double Omega, int spanwiseLiftSwitch, double* wiP,
                   const double* sec, double* loads) {
  for (int ib = 0; ib < nbConvect; ++ib) {
    double* w = wiP + (size_t)vlc::cp::kRec * nc * ns * ib;
    const double* s = sec + (size_t)vlc::cp::sec_doubles(ns) * ib;
    double* l = loads + (size_t)vlc::cp::loads_doubles(ns) * ib;
    // the four phases of cp_loads_kernel, each over its panels / sections in REVERSE order (any order within a phase: a
    // phase reads only what an earlier phase wrote)
    std::vector<double> scr((size_t)vlc::cp::kScr * nc * ns, std::nan(""));
    for (int q = nc * ns - 1; q >= 0; --q) vlc::cp::loads_panel_resvel(nc, ns, q % nc + 1, q / nc + 1, w, s, scr.data());
    for (int is = ns; is >= 1; --is) vlc::cp::loads_section_dirs(nc, ns, is, w, s, Omega, l, scr.data());
    for (int q = nc * ns - 1; q >= 0; --q)
      vlc::cp::loads_panel_forces(nc, ns, q % nc + 1, q / nc + 1, w, density, dt, Omega, spanwiseLiftSwitch, l, scr.data());
    for (int is = ns; is >= 1; --is) vlc::cp::loads_section_sums(nc, ns, is, s, density, l, scr.data());
    vlc::cp::blade_sum_loads(ns, l);
  }
}

// = cp_rhs_kernel
void cp_host_rhs(int N, int npb, int nbConvect, int axisym, const double* wiP, double* RHS) {
  for (int i = 0; i < N; ++i) RHS[i] = vlc::cp::rhs_entry(i, npb, nbConvect, axisym, wiP);
}

// = cp_map_gam_kernel
void cp_host_map_gam(int nb, int npb, int nbConvect, int axisym, const double* gamVec, double* wiP) {
  for (int i = 0; i < nb * npb; ++i) {
    const int src = vlc::cp::map_gam_source(i / npb, i % npb, npb, nbConvect, axisym);
    if (src >= 0) wiP[(size_t)vlc::cp::kRec * i + vlc::cp::kGam] = gamVec[src];
  }
}

int cp_host_sec_doubles(int ns) { return vlc::cp::sec_doubles(ns); }
int cp_host_loads_doubles(int ns) { return vlc::cp::loads_doubles(ns); }

}  // extern "C"
