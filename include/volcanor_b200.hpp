// volcanor_b200.hpp -- header-only C++ mirror of the reference's procedure interface over the C ABI
// (include/volcanor_b200.h).  Names, argument meaning and error behaviour follow the Fortran procedures they
// replace so that host code reads like the reference's call sites:
//
//   reference (Fortran)                                       here (C++)
//   rotor%vind_bywing(P)            classdef.f90:4424    ->   rotor.vind_bywing(P, m, V)
//   rotor%vind_bywake(P [, 'P'])    classdef.f90:4459    ->   rotor.vind_bywake(P, m, V, predicted)
//   rotor%vind_bywing_boundVortices classdef.f90:4445    ->   rotor.vind_bywing_boundVortices(P, m, V)
//   vind_onNwake_byRotor(rotor, Nwake [, 'P'])  libCommon.f90:114  ->  vind_onNwake_byRotor(rotor, Nwake, rows, cols, ld, out, predicted)
//   vind_onFwake_byRotor(rotor, Fwake [, 'P'])  libCommon.f90:173  ->  vind_onFwake_byRotor(rotor, Fwake, rows, out, predicted)
//   rotor%calcAIC()                 classdef.f90:4151    ->   rotor.calcAIC(AIC_out)
//   matmulAX(AIC_inv, RHS)          libMath.f90:105      ->   rotor.solve(RHS, gamVec)
//   rotor%assignshed('LE'|'TE')     classdef.f90:4297    ->   rotor.assignshed("LE")          (device copies, tier 2b)
//   rotor%age_wake(dt)              classdef.f90:4331    ->   rotor.age_wake(dt, omegaSlow)
//   rotor%dissipate_wake(dt, nu)    classdef.f90:4356    ->   rotor.dissipate_wake(dt, nu)
//   rotor%strain_wake()             classdef.f90:4410    ->   rotor.strain_wake()
//   rotor%convectwake(iter, dt, c)  classdef.f90:4786    ->   rotor.convectwake(dt, 'C' | 'P')
//   rotor%rollup()                  classdef.f90:4515    ->   rotor.rollup()
//
// Errors: the reference aborts with `error stop '<msg>'`; here every failed call throws vlc::Error carrying the
// library's message (there is no CPU fallback to catch it with).  Arrays are column-major (3, m) doubles exactly as
// Fortran passes them; wake / wing state is handed over as arrays of doubles in the reference's record layout
// (vr_class = 50, Fwake_class = 13, wingpanel_class = 104 doubles).
#ifndef VOLCANOR_B200_HPP
#define VOLCANOR_B200_HPP

#include <cstdint>
#include <stdexcept>
#include <string>

#include "volcanor_b200.h"

namespace vlc {

struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string& msg) : std::runtime_error(msg), code(c) {}
};

class Context {
 public:
  explicit Context(int device = 0) {
    const int rc = vlc_create(device, &h_);
    if (rc != VLC_OK) throw Error(rc, std::string("vlc_create: ") + vlc_last_error(nullptr));
  }
  ~Context() { vlc_destroy(h_); }
  Context(const Context&) = delete;
  Context& operator=(const Context&) = delete;
  vlc_ctx* handle() const { return h_; }
  void check(int rc) const {
    if (rc != VLC_OK) throw Error(rc, vlc_last_error(h_));
  }
  // tier 1: flat filament sets
  void set_sources(int set, std::int64_t n, const double* p1, const double* p2, const double* rvc, const double* gam,
                   const std::uint8_t* wake_flag = nullptr) {
    check(vlc_set_sources(h_, set, n, p1, p2, rvc, gam, wake_flag));
  }
  void vind(int set, std::int64_t m, const double* P, double* V) { check(vlc_vind(h_, set, m, P, V)); }
  // the wake sweeps of the convection driver for all rotors at once (main.f90:800-838, :889-911), on device copies
  void wake_sweep(bool predicted, bool addInitWakeVel = false) { check(vlc_wake_sweep(h_, predicted, addInitWakeVel)); }
  // program gridgen (src/gridgen.f90)
  void gridgen(int nx, int ny, int nz, const double* xyzMin, const double* xyzMax, const double* vel,
               std::int64_t nVrWing, const double* vrWing, std::int64_t nVrNwake, const double* vrNwake,
               std::int64_t nVfNwakeTE, const double* vfNwakeTE, const double* gamNwakeTE, std::int64_t nVfFwake,
               const double* vfFwake, const double* gamFwake, double* gridCentre, double* velCentre) {
    check(vlc_gridgen(h_, nx, ny, nz, xyzMin, xyzMax, vel, nVrWing, vrWing, nVrNwake, vrNwake, nVfNwakeTE, vfNwakeTE,
                      gamNwakeTE, nVfFwake, vfFwake, gamFwake, gridCentre, velCentre));
  }

 private:
  vlc_ctx* h_ = nullptr;
};

// One rotor of the reference's global rotor(:) array (0-based index ir).
class Rotor {
 public:
  Rotor(Context& c, int ir, int nb, int nc, int ns, int nNwake, int nFwake, int surfaceType = 1)
      : c_(c), ir_(ir), nb_(nb), nc_(nc), ns_(ns), nNwake_(nNwake), nFwake_(nFwake) {
    c_.check(vlc_rotor_define(c_.handle(), ir, nb, nc, ns, nNwake, nFwake, surfaceType));
  }
  int N() const { return nc_ * ns_ * nb_; }
  // state hand-over (what gpu_sync_rotor of fortran/libGPU.f90 does with transfer(...))
  void set_rows(int rowNear, int rowFar) { c_.check(vlc_rotor_set_rows(c_.handle(), ir_, rowNear, rowFar)); }
  void put_wing(int ib, const double* wiP) { c_.check(vlc_rotor_put_wing(c_.handle(), ir_, ib, wiP)); }
  void put_wing_gam(int ib, const double* gam) { c_.check(vlc_rotor_put_wing_gam(c_.handle(), ir_, ib, gam)); }
  void put_nwake(int ib, const double* waN, bool predicted = false) {
    c_.check(vlc_rotor_put_nwake(c_.handle(), ir_, ib, predicted, waN));
  }
  void put_fwake(int ib, const double* waF, bool predicted = false) {
    c_.check(vlc_rotor_put_fwake(c_.handle(), ir_, ib, predicted, waF));
  }
  void put_pfwake(int ib, const double* wapF, bool predicted = false) {
    c_.check(vlc_rotor_put_pfwake(c_.handle(), ir_, ib, predicted, wapF));
  }
  // the reference's type-bound procedures, batched over m points
  void vind_bywing(const double* P, std::int64_t m, double* V) { c_.check(vlc_rotor_vind_bywing(c_.handle(), ir_, m, P, V)); }
  void vind_bywake(const double* P, std::int64_t m, double* V, bool predicted = false) {
    c_.check(vlc_rotor_vind_bywake(c_.handle(), ir_, predicted, m, P, V));
  }
  void vind_bywing_boundVortices(const double* P, std::int64_t m, double* V) {
    c_.check(vlc_rotor_vind_bywing_boundVortices(c_.handle(), ir_, m, P, V));
  }
  void vind_bywing_chordwiseVortices(const double* P, std::int64_t m, double* V) {
    c_.check(vlc_rotor_vind_bywing_chordwiseVortices(c_.handle(), ir_, m, P, V));
  }
  void vind(const double* P, std::int64_t m, double* V, bool predicted = false) {
    c_.check(vlc_rotor_vind(c_.handle(), ir_, predicted, m, P, V));
  }
  void calcAIC(double* AIC_out = nullptr) { c_.check(vlc_rotor_calcAIC(c_.handle(), ir_, AIC_out)); }
  void solve(const double* RHS, double* gamVec) { c_.check(vlc_rotor_solve(c_.handle(), ir_, RHS, gamVec)); }
  void get_AIC_inv(double* AIC_inv) { c_.check(vlc_rotor_get_AIC_inv(c_.handle(), ir_, AIC_inv)); }
  // tier 2b: the reference's wake mutators on the library's device copies (device-resident time stepping)
  void set_wake_params(int nbConvect, int axisymmetrySwitch, int ductSwitch, int suppressFwakeSwitch, int rollupStart,
                       int rollupEnd, double rollupSign, double apparentViscCoeff, double decayCoeff, double initWakeVel) {
    c_.check(vlc_rotor_set_wake_params(c_.handle(), ir_, nbConvect, axisymmetrySwitch, ductSwitch, suppressFwakeSwitch,
                                       rollupStart, rollupEnd, rollupSign, apparentViscCoeff, decayCoeff, initWakeVel));
  }
  void set_frame(const double* shaftAxis, const double* hubCoords) {
    c_.check(vlc_rotor_set_frame(c_.handle(), ir_, shaftAxis, hubCoords));
  }
  void assignshed(const std::string& edge) {
    if (edge != "LE" && edge != "TE") throw Error(VLC_ERR_ARG, "ERROR: Wrong option for edge");  // classdef.f90:4322
    c_.check(vlc_rotor_assignshed(c_.handle(), ir_, edge == "LE" ? 0 : 1));
  }
  void age_wake(double dt, double omegaSlow) { c_.check(vlc_rotor_age_wake(c_.handle(), ir_, dt, omegaSlow)); }
  void dissipate_wake(double dt, double kinematicVisc) { c_.check(vlc_rotor_dissipate_wake(c_.handle(), ir_, dt, kinematicVisc)); }
  void strain_wake() { c_.check(vlc_rotor_strain_wake(c_.handle(), ir_)); }
  void calc_skew() { c_.check(vlc_rotor_calc_skew(c_.handle(), ir_)); }
  void burst_wake(double skewLimit, double largeCoreRadius) { c_.check(vlc_rotor_burst_wake(c_.handle(), ir_, skewLimit, largeCoreRadius)); }
  void wake_to_predicted() { c_.check(vlc_rotor_wake_to_predicted(c_.handle(), ir_)); }
  void convectwake(double dt, char wakeType) {
    if (wakeType != 'C' && wakeType != 'P') throw Error(VLC_ERR_ARG, "ERROR: Wrong character flag for convectwake()");
    c_.check(vlc_rotor_convectwake(c_.handle(), ir_, dt, wakeType == 'P'));
  }
  // rotor%updatePrescribedWake(dt, wakeType) classdef.f90:5170-5218; deltaPsi = omegaSlow*dt
  void updatePrescribedWake(double deltaPsi, int prescWakeGenNt, char wakeType) {
    if (wakeType != 'C' && wakeType != 'P') throw Error(VLC_ERR_ARG, "ERROR: Wrong character flag for updatePrescribedWake()");
    c_.check(vlc_rotor_updatePrescribedWake(c_.handle(), ir_, deltaPsi, prescWakeGenNt, wakeType == 'P'));
  }
  void rollup() { c_.check(vlc_rotor_rollup(c_.handle(), ir_)); }
  void wakevel_op(int op) { c_.check(vlc_rotor_wakevel_op(c_.handle(), ir_, op)); }
  void wakevel_copy(int dst, int src) { c_.check(vlc_rotor_wakevel_copy(c_.handle(), ir_, dst, src)); }
  void wakevel_lincomb(int dst, int nterms, const int* src, const double* coef, double divisor) {
    c_.check(vlc_rotor_wakevel_lincomb(c_.handle(), ir_, dst, nterms, src, coef, divisor));
  }
  void get_nwake(int ib, double* waN, bool predicted = false) { c_.check(vlc_rotor_get_nwake(c_.handle(), ir_, ib, predicted, waN)); }
  void get_fwake(int ib, double* waF, bool predicted = false) { c_.check(vlc_rotor_get_fwake(c_.handle(), ir_, ib, predicted, waF)); }
  void put_pfwake_helix(int ib, const double* helix, bool predicted = false) {
    c_.check(vlc_rotor_put_pfwake_helix(c_.handle(), ir_, ib, predicted, helix));
  }
  void get_pfwake(int ib, double* wapF, double* helix = nullptr, bool predicted = false) {
    c_.check(vlc_rotor_get_pfwake(c_.handle(), ir_, ib, predicted, wapF, helix));
  }
  // tier 2c: the collocation-point stage on the device copies of the wing records (main.f90:548-670)
  void calc_RHS(double* velCP_out = nullptr, double* RHS_out = nullptr) {
    c_.check(vlc_rotor_calc_RHS(c_.handle(), ir_, velCP_out, RHS_out));
  }
  void solve_map_gam(double* gamVec_out = nullptr) { c_.check(vlc_rotor_solve_map_gam(c_.handle(), ir_, gamVec_out)); }
  void put_sections(int ib, const double* sec) { c_.check(vlc_rotor_put_sections(c_.handle(), ir_, ib, sec)); }
  void calc_velCPTotal() { c_.check(vlc_rotor_calc_velCPTotal(c_.handle(), ir_)); }
  void calc_force(double density, double dt, double Omega, int spanwiseLiftSwitch = 0) {  // + calc_secAlpha
    c_.check(vlc_rotor_calc_force(c_.handle(), ir_, density, dt, Omega, spanwiseLiftSwitch));
  }
  void get_loads(int ib, double* loads) { c_.check(vlc_rotor_get_loads(c_.handle(), ir_, ib, loads)); }
  void get_wing(int ib, double* wiP) { c_.check(vlc_rotor_get_wing(c_.handle(), ir_, ib, wiP)); }
  int sections_doubles() const { return 10 * ns_ + 6; }
  int loads_doubles() const { return 12 + 25 * ns_; }
  Context& context() const { return c_; }
  int index() const { return ir_; }

 private:
  Context& c_;
  int ir_, nb_, nc_, ns_, nNwake_, nFwake_;
};

// libCommon.f90:114-171.  Nwake points at element (1,1) of the slice waN(rowNear:nNwakeEnd, :), ld = nNwake.
inline void vind_onNwake_byRotor(Rotor& rotor, const double* Nwake, int rows, int cols, int ld, double* vindArray,
                                 bool predicted = false) {
  rotor.context().check(vlc_vind_onNwake_byRotor(rotor.context().handle(), rotor.index(), Nwake, rows, cols, ld, predicted,
                                                 vindArray));
}
// libCommon.f90:173-211
inline void vind_onFwake_byRotor(Rotor& rotor, const double* Fwake, int rows, double* vindArray, bool predicted = false) {
  rotor.context().check(vlc_vind_onFwake_byRotor(rotor.context().handle(), rotor.index(), Fwake, rows, predicted, vindArray));
}

}  // namespace vlc
#endif
