/*
 * vlc_case.h -- CPU ORACLE, case driver (test infrastructure, NOT product code).
 *
 * Plain-C restatement of the reference's program flow (src/main.f90) and of the subset of
 * rotor_init / kinematics / loads (src/classdef.f90) that the reference's own golden results
 * exercise, so that the oracle's wake sweeps, convection, core growth and roll-up are pinned
 * END TO END by the reference's CT/CL histories
 * (tests/katzNplotkin-AR04.case/referenceResults/r01ForceNonDim.csv.ref) and by the force KATs of
 * tests/rotor1x2_test.f90 / wing1x2_test.f90 / wing1x3_test.f90.
 *
 * The five hot-path call sites of the driver go through a table of function pointers
 * (orc_hooks_t).  The default table is the CPU oracle itself; the GPU parity tests install a
 * table that forwards to the C ABI of include/volcanor_b200.h -- i.e. the same driver, with its
 * OpenMP vind loops replaced by the library, which is exactly the substitution the Fortran shim
 * performs.
 *
 * Scope: surfaceType 1 (lifting) rotors and wings, geometryFile '0' or a PLOT3D grid passed in
 * memory, forceCalcSwitch 0, fdScheme 0-5, slowStart 0-3, wake dissipation / strain,
 * axisymmetry, far-wake roll-up and truncation, wake burst, prescribed far wake.  Not restated (unused by every shipped case):
 * image surfaces, non-lifting STL bodies, camber files, C81 tables, blade/body dynamics, custom
 * trajectories.  (Restated although unused by every shipped case, parity unpinned: wake strain, wake burst, the
 * prescribed far wake.)
 */
#ifndef VLC_CASE_H
#define VLC_CASE_H

#include "vlc_oracle.h"

#ifdef __cplusplus
extern "C" {
#endif

/* config.nml (libCommon.f90:51-108); missing keys are 0. */
typedef struct {
  int nt, nr;
  double dt, density, velSound, kinematicVisc;
  int ntSub, ntSubInit, rotorForcePlot, wakeDissipation, wakeStrain, wakeBurst, wakeSuppress;
  int slowStart, slowStartNt, fdScheme, initWakeVelNt;
} orc_config_t;

/* Hot-path call sites of the driver.  Every function returns 0 on success. */
typedef struct orc_hooks {
  void *user;
  /* rotor(jr)%vind_bywing (what 0) / %vind_bywake[,'P'] (1) / both (2) / %vind_bywing_boundVortices (3),
   * batched over m points: replaces the per-point calls inside the OpenMP loops main.f90:124-167,
   * :247-271, :528-573, :632-656. */
  int (*vind_points)(void *user, int jr, int what, int predicted, long m, const double *P, double *V);
  /* vind_onNwake_byRotor / vind_onFwake_byRotor (libCommon.f90:114-211) */
  int (*vind_onNwake)(void *user, int jr, const double *Nwake, int rows, int cols, int ld, int predicted,
                      double *out);
  int (*vind_onFwake)(void *user, int jr, const double *Fwake, int rows, int predicted, double *out);
  /* rotor%calcAIC() (classdef.f90:4151): fill AIC (N x N); AIC_inv may be left untouched if solve is set */
  int (*calcAIC)(void *user, int ir, double *AIC, double *AIC_inv);
  /* gamVec = matmulAX(AIC_inv, RHS) (main.f90:190, :596) */
  int (*solve)(void *user, int ir, const double *RHS, double *gamVec);
  /* OPTIONAL, both or neither (NULL = this file's CPU restatement): device-resident stepping.  wake_prestep replaces
   * main.f90:466-506 (assignshed 'LE', age_wake, dissipate_wake of every rotor); wake_convect replaces :800-1440 (the
   * wake sweeps, the fdScheme switch with its velocity bookkeeping, strain_wake, rollup, assignshed 'TE').  When they
   * are set the driver never reads or writes its own wake records or wake velocity arrays: the hook owner holds them
   * (tests/native/case_gpu_hooks.c in resident mode: the C ABI's tier 2b). */
  int (*wake_prestep)(void *user, int iter);
  int (*wake_convect)(void *user, int iter);
  /* OPTIONAL, each on its own (NULL = the call sites above): the collocation-point stage by the hook owner (the C ABI's
   * tier 2c).  cp_rhs_solve replaces main.f90:548-603 of every rotor inside the time loop (ntSub = 0): the driver has
   * written the kinematic velCP into its wing records (:528-547) and saved gamVecPrev; the hook leaves velCP in the
   * records, RHS and gamVec in the rotors' vectors and owns the new circulation (the driver maps gamVec into its own
   * records afterwards without marking the wing as changed).  cp_forces replaces main.f90:630-663 + calc_secAlpha +
   * calc_force of rotor ir down to the blade sums: it leaves the wing records, the sectional arrays and the blade
   * forces of every blade in the driver's arrays; the driver then adds the blades (sumBladeToNetForces). */
  int (*cp_rhs_solve)(void *user);
  int (*cp_forces)(void *user, int ir);
} orc_hooks_t;

typedef struct orc_case orc_case_t;

orc_case_t *orc_case_new(int nr);
void orc_case_free(orc_case_t *c);
/* key = namelist variable name of config.nml; returns 0 if the key is known */
int orc_case_set_config(orc_case_t *c, const char *key, double value);
/* key = namelist variable name of geomXX.nml (ir 0-based); vectors take n values.
 * "grid" passes a PLOT3D grid (3, nc+1, ns+1) as read by rotor_plot3dtoblade (classdef.f90:3957). */
int orc_case_set_geom(orc_case_t *c, int ir, const char *key, int n, const double *values);
void orc_case_set_hooks(orc_case_t *c, const orc_hooks_t *hooks); /* NULL = CPU oracle */
/* Only the two wake-stage hooks, with their own `user`; the five call sites keep what orc_case_set_hooks installed
 * (call it first).  NULL functions restore the inline stages. */
void orc_case_set_stage_hooks(orc_case_t *c, void *user, int (*wake_prestep)(void *, int), int (*wake_convect)(void *, int));
/* Only the two collocation-point hooks, with their own `user` (call orc_case_set_hooks first).  NULL restores the call sites. */
void orc_case_set_cp_hooks(orc_case_t *c, void *user, int (*cp_rhs_solve)(void *), int (*cp_forces)(void *, int));
/* main.f90:1-382: init rotors, pitch, AIC, initial solution, initial forces.  Returns 0 / error code. */
int orc_case_init(orc_case_t *c);
/* one pass of the time loop main.f90:400-1452 (iter = 1..nt) */
int orc_case_step(orc_case_t *c);
int orc_case_iter(const orc_case_t *c);
const orc_config_t *orc_case_config(const orc_case_t *c);
orc_rotor_t *orc_case_rotor(orc_case_t *c, int ir);
const char *orc_case_error(const orc_case_t *c);
/* the columns force2file writes to rNNForceNonDim.csv (libPostprocess.f90:824-837):
 * out[0..8] = CL/CT, CD/CQ, CLu, CDi, CD0, CDu, CFx, CFy, CFz */
void orc_case_force_nondim(orc_case_t *c, int ir, double out[9]);
/* pair interactions (vf_vind evaluations in the reference's enumeration) of the last step */
double orc_case_pairs_last_step(const orc_case_t *c);

/* the wake stages as separate calls on the driver's own (CPU) state: the CPU backend of the staged orchestration in
 * tests/native/case_gpu_hooks.c, whose result must equal this file's inline time loop bit for bit */
int orc_case_wake_sweep(orc_case_t *c, int predicted);
void orc_rotor_wake_to_predicted(orc_rotor_t *r);
int orc_rotor_wakevel_op(orc_rotor_t *r, int op);
/* velocity arrays by id, as in the C ABI (vlc_rotor_wakevel_copy / _lincomb) */
enum { ORC_VEL = 0, ORC_VEL_1 = 1, ORC_VEL_PREDICTED = 2, ORC_VEL_STEP = 3, ORC_VEL_2 = 4, ORC_VEL_3 = 5 };
int orc_rotor_wakevel_copy(orc_rotor_t *r, int dst, int src);
int orc_rotor_wakevel_lincomb(orc_rotor_t *r, int dst, int nterms, const int *src, const double *coef, double divisor);

/* pieces exposed for the KAT tests */
double orc_pwl_interp1d(int n, const double *x, const double *y, double q); /* libMath.f90:476-517 */
int orc_case_init_rotors(orc_case_t *c); /* main.f90:31-58 only: rotor%init + initial pitch */
void orc_blade_rot_pitch(orc_blade_t *b, double theta);
double orc_rotor_gettheta(const orc_rotor_t *r, double psi, int ib);
void orc_rotor_dirLiftDrag(orc_rotor_t *r);
void orc_rotor_calc_secAlpha(orc_rotor_t *r);
void orc_rotor_calc_force(orc_rotor_t *r, double density, double dt);
/* the tail of rotor_calc_force: copies for an axisymmetric rotor + sumBladeToNetForces (classdef.f90:4623-4671) */
void orc_rotor_sum_forces(orc_rotor_t *r);
/* out[0..7] = radius root_cut chord Omega nonDimforceDenominator nNwake wakeTruncateNt prescWakeNt (params2file) */
void orc_rotor_get_file_params(const orc_rotor_t *r, double out[8]);
/* out[0..3] = Omega, spanwiseLiftSwitch, axisymmetrySwitch, nbConvect */
void orc_rotor_get_force_params(const orc_rotor_t *r, double out[4]);
double *orc_blade_sec(orc_rotor_t *r, int ib, const char *name); /* sectional arrays by name */

#ifdef __cplusplus
}
#endif
#endif
