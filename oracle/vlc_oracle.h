/*
 * vlc_oracle.h -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * Plain-C restatement of the Biot-Savart hot path of cibinjoseph/VOLCANOR
 * (Fortran 2008 + OpenMP).  Every function cites the reference file:line it
 * follows.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may link or call this; the product library
 * (volcanor_b200/) never does.
 *
 * Parity pin: the reference cannot be compiled here (no Fortran compiler), so
 * this restatement is pinned against the reference's own known-answer tests:
 *   tests/wing1x3_test.f90:83-85 (AIC, 15 digits), tests/wing1x2_test.f90:164,
 *   tests/rotor1x2_test.f90:182-218 (AIC + gamVec), the CT/CL histories in
 *   tests/katzNplotkin-AR04.case/referenceResults/ (see tests/test_oracle_*.py).
 *
 * Memory layout deliberately equals the reference's derived types
 * (src/classdef.f90:57-104, 106-179, 198-220): arrays of these structs are
 * bit-compatible with `transfer(blade%waN, buf)` on the Fortran side.
 */
#ifndef VLC_ORACLE_H
#define VLC_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

/* libMath.f90:11  eps = epsilon(1._dp) */
#define ORC_EPS 2.220446049250313e-16

/* classdef.f90:57-79  vf_class: 12 doubles = 96 B. fc(3,2): fc[k][xyz], k=0 -> fc(:,1) */
typedef struct {
  double fc[2][3];
  double l0, lc, rVc0, rVc, age, ageAzimuthal;
} orc_vf_t;

/* classdef.f90:81-104  vr_class: 4 vf + gam + skew = 50 doubles = 400 B */
typedef struct {
  orc_vf_t vf[4];
  double gam, skew;
} orc_vr_t;

/* classdef.f90:106-179  wingpanel_class: 104 doubles = 832 B */
typedef struct {
  orc_vr_t vr;
  double gamPrev, gamTrapz;
  double PC[4][3]; /* PC(3,4): PC[n][xyz] */
  double CP[3], nCap[3], tauCapChord[3], tauCapSpan[3];
  double velCP[3], velCPTotal[3], velCPm[3];
  double normalForce[3], normalForceUnsteady[3], chordwiseResVel[3];
  double velPitch, delP, delPUnsteady, delDiConstant, delDiUnsteady;
  double meanChord, meanSpan, panelArea, rHinge, alpha;
} orc_wingpanel_t;

/* classdef.f90:198-220  Fwake_class: vf + gam = 13 doubles = 104 B */
typedef struct {
  orc_vf_t vf;
  double gam;
} orc_fwake_t;

#define ORC_NPFWAKE 240 /* classdef.f90:225 */

/* The hot-path subset of blade_class (classdef.f90:238-358). Arrays are
 * column-major like Fortran: wiP(i,j) -> wiP[(i-1) + nc*(j-1)], waN(i,j) ->
 * waN[(i-1) + nNwake*(j-1)], vel*(d,i,j) -> v[(d-1) + 3*((i-1) + nNwake*(j-1))]. */
typedef struct {
  int nc, ns, nNwake, nFwake;
  orc_wingpanel_t *wiP;
  orc_vr_t *waN, *waNPredicted;
  orc_fwake_t *waF, *waFPredicted;
  orc_fwake_t wapF[ORC_NPFWAKE], wapFPredicted[ORC_NPFWAKE];
  double pfHelixPitch[2], pfHelixRadius[2]; /* pFwake_class%helixPitch / helixRadius of wapF [0] and wapFPredicted [1] */
  double *velNwake, *velNwake1, *velNwakePredicted, *velNwakeStep; /* (3,nNwake,ns+1) */
  double *velFwake, *velFwake1, *velFwakePredicted, *velFwakeStep; /* (3,nFwake) */
  double *velNwake2, *velNwake3, *velFwake2, *velFwake3;           /* fdScheme 4 / 5 histories (classdef.f90:3733-3824) */
  /* ---- case-driver state (vlc_case.c; classdef.f90:238-312) ---- */
  double theta, psi, pivotLE, preconeAngle, flap, dflap;
  double flapOrigin[3];
  double forceInertial[3], lift[3], drag[3], liftUnsteady[3];
  double xAxis[3], yAxis[3], zAxis[3];
  double xAxisAzi[3], yAxisAzi[3], zAxisAzi[3];
  double xAxisAziFlap[3], yAxisAziFlap[3], zAxisAziFlap[3];
  double *secChord, *secArea, *secAlpha, *secCL, *secCLu, *secCD, *secMflapArm; /* (ns) */
  double *secForceInertial, *secLift, *secDrag, *secLiftDir, *secDragDir, *secLiftUnsteady; /* (3,ns) */
  double *secTauCapChord, *secTauCapSpan, *secNormalVec, *secCP, *secChordwiseResVel;       /* (3,ns) */
  int spanwiseLiftSwitch;
} orc_blade_t;

/* The hot-path subset of rotor_class (classdef.f90:360-467). */
typedef struct {
  int nb, nc, ns, nNwake, nFwake, nbConvect, nNwakeEnd, nFwakeEnd;
  int rowNear, rowFar;
  int surfaceType, axisymmetrySwitch, ductSwitch, suppressFwakeSwitch;
  int rollupStart, rollupEnd;
  int prescWakeNt;
  double Omega, omegaSlow;
  double shaftAxis[3], hubCoords[3];
  double controlPitch[3];
  double apparentViscCoeff, decayCoeff;
  orc_blade_t *blade;
  double *AIC, *AIC_inv; /* (N,N) column-major, N = nc*ns*nb */
  double *gamVec, *gamVecPrev, *RHS;
  /* ---- case-driver state (vlc_case.c; classdef.f90:360-416) ---- */
  int propConvention, spanSpacing, chordSpacing, spanwiseLiftSwitch, symmetricTau, forceCalcSwitch;
  int wakeTruncateNt, prescWakeAfterTruncNt, prescWakeGenNt;
  double radius, root_cut, chord, preconeAngle, thetaTwist, pivotLE, flapHinge, psi, psiStart;
  double pts[3], cgCoords[3], fromCoords[3], velBody[3], omegaBody[3];
  double xAxisBody[3], yAxisBody[3], zAxisBody[3];
  double dragUnitVec[3], sideUnitVec[3], liftUnitVec[3];
  double spanwiseCore, *streamwiseCoreVec; /* (ns+1) */
  double rollupStartRadius, rollupEndRadius, initWakeVel, skewLimit;
  double nonDimforceDenominator;
  double forceInertial[3], lift[3], liftPrev[3], drag[3], liftUnsteady[3];
  /* generation counters bumped by the case driver whenever it changes the wing / the 'C' wake / the 'P' wake, so a
   * shim can skip uploads of unchanged state (tests/native/case_gpu_hooks.c) */
  unsigned long gen_wing, gen_wake[2];
} orc_rotor_t;

/* ---- libMath.f90 ---- */
double orc_norm2(const double a[3]);                              /* intrinsic norm2 */
void orc_unitVec(const double a[3], double u[3]);                 /* libMath.f90:249-262 */
void orc_cross(const double a[3], const double b[3], double c[3]);/* libMath.f90:202-212 */
void orc_getTransformAxis(double theta, const double axisVec[3], double T[9]); /* :695-726; T column-major */
int orc_inv2(int n, const double *A, double *Ainv);               /* libMath.f90:48-83 (DGETRF+DGETRI restated) */
void orc_matmulAX(int m, int n, const double *A, const double *X, double *AX); /* :105-122 */

/* ---- pair kernel: classdef.f90:476-503, 527-542 ---- */
void orc_vf_vind(const orc_vf_t *f, const double P[3], double v[3]);
/* classdef.f90:998-1066 pFwake_update and :5170-5218 rotor_updatePrescribedWake (PARITY UNPINNED: no shipped case and no
 * reference test enables the prescribed far wake; restated from the source only) */
int orc_pfwake_update(orc_fwake_t *pf, double *helixPitch, double *helixRadius, const orc_fwake_t *waF, int nFwake,
                      const double hubCoords[3], const double shaftAxis[3], double deltaPsi);
int orc_rotor_updatePrescribedWake(orc_rotor_t *r, double dt, char wakeType);
void orc_rotor_get_presc(const orc_rotor_t *r, int out[3]); /* prescWakeNt, prescWakeAfterTruncNt, prescWakeGenNt */
void orc_rotor_set_pfHelix(orc_rotor_t *r, int ib, int predicted, const double in[2]);
void orc_rotor_get_pfHelix(const orc_rotor_t *r, int ib, int predicted, double out[2]); /* helixPitch, helixRadius */
void orc_vr_vind(const orc_vr_t *r, const double P[3], double v[3]);

/* ---- source loops: classdef.f90:1342-1513, 4424-4479 ---- */
void orc_blade_vind_bywing(const orc_blade_t *b, const double P[3], double v[3]);
void orc_blade_vind_bywing_boundVortices(const orc_blade_t *b, const double P[3], double v[3]);
void orc_blade_vind_bywing_chordwiseVortices(const orc_blade_t *b, const double P[3], double v[3]);
void orc_blade_vind_boundVortex(const orc_blade_t *b, int ic, int is, const double P[3], double v[3]);
void orc_blade_vind_bywake(const orc_blade_t *b, int rowNear, int rowFar, const double P[3], int predicted, double v[3]);
void orc_rotor_vind_bywing(const orc_rotor_t *r, const double P[3], double v[3]);
void orc_rotor_vind_bywing_boundVortices(const orc_rotor_t *r, const double P[3], double v[3]);
void orc_rotor_vind_bywake(const orc_rotor_t *r, const double P[3], int predicted, double v[3]);

/* ---- target sweeps: libCommon.f90:114-258 ---- */
/* Nwake points at element (1,1) of the slice; ld = leading dimension (nNwake of the parent array).
 * out is (3, rows, cols+1) column-major. */
void orc_vind_onNwake_byRotor(const orc_rotor_t *src, const orc_vr_t *Nwake, int rows, int cols, int ld, int predicted, double *out);
void orc_vind_onFwake_byRotor(const orc_rotor_t *src, const orc_fwake_t *Fwake, int rows, int predicted, double *out);
void orc_vel_order2_Nwake(const double *vn, const double *vnp1, int rows, int cols, double *out); /* all (3,rows,cols) */
void orc_vel_order2_Fwake(const double *vn, const double *vnp1, int rows, double *out);

/* ---- vr/fwake helpers: classdef.f90:505-521, 569-624, 933-959 ---- */
void orc_vr_assignP(orc_vr_t *r, int n, const double P[3]);
void orc_vr_shiftdP(orc_vr_t *r, int n, const double d[3]);
void orc_vf_calclength(orc_vf_t *f, int isOriginal);

/* ---- wake state updates ---- */
void orc_blade_convectwake(orc_blade_t *b, int rowNear, int rowFar, double dt, char wakeType, int ductSwitch); /* :1515-1575 */
void orc_blade_wake_continuity(orc_blade_t *b, int rowNear, int rowFar, char wakeType, int ductSwitch);        /* :1609-1702 */
void orc_rotor_convectwake(orc_rotor_t *r, int iter, double dt, char wakeType);                                 /* :4786-4830 */
void orc_rotor_assignshed(orc_rotor_t *r, const char *edge);                                                    /* :4297-4325 */
void orc_rotor_age_wake(orc_rotor_t *r, double dt);                                                             /* :4331-4354 */
void orc_rotor_dissipate_wake(orc_rotor_t *r, double dt, double kinematicViscosity);                            /* :4356-4408 */
void orc_rotor_strain_wake(orc_rotor_t *r);                                                                     /* :4410-4422 */
void orc_rotor_burst_wake(orc_rotor_t *r); /* :4911-4917, :2306-2339 (far wake; parity unpinned: no shipped case) */
void orc_rotor_calc_skew(orc_rotor_t *r); /* :4919-4936 (output for skew2file; parity unpinned) */
double orc_vr_skew(const orc_vr_t *v);
int orc_burst_pair(const orc_fwake_t *f0, const orc_fwake_t *f1, double skewLimit);
void orc_rotor_shiftwake(orc_rotor_t *r);                                                                       /* :4481-4498 */
void orc_rotor_shiftFwake(orc_rotor_t *r);                                                                      /* :4500-4513 */
void orc_rotor_rollup(orc_rotor_t *r);                                                                          /* :4515-4605 */

/* ---- AIC: classdef.f90:4151-4196 ---- */
int orc_rotor_calcAIC(orc_rotor_t *r);
void orc_rotor_map_gam(orc_rotor_t *r);

/* ---- flat helpers used by the tests and the CPU baseline ---- */
/* Sum of gam*vf_vind over a flat list of n filaments (reference arithmetic per pair),
 * fil: n records of orc_vf_t, gam[n]; skip[n] != 0 applies the wake rule
 * |gam| > eps (classdef.f90:1452,1466).  V (3,m) is overwritten.  OpenMP over targets
 * like libCommon.f90:132-139 when built with -fopenmp. */
void orc_vind_flat(long n, const orc_vf_t *fil, const double *gam, const unsigned char *skip,
                   long m, const double *P, double *V);
/* long-double accumulation + long-double kernel: the "truth" used to scale tolerances. */
void orc_vind_flat_ld(long n, const orc_vf_t *fil, const double *gam, const unsigned char *skip,
                      long m, const double *P, double *V, double *Vabs);
int orc_num_threads(void);
/* program gridgen (src/gridgen.f90:62-145) */
void orc_gridgen(int nx, int ny, int nz, const double xyzMin[3], const double xyzMax[3], const double vel[3],
                 long nVrWing, const orc_vr_t *vrWing, long nVrNwake, const orc_vr_t *vrNwake, long nVfNwakeTE,
                 const orc_vf_t *vfNwakeTE, const double *gamNwakeTE, long nVfFwake, const orc_vf_t *vfFwake,
                 const double *gamFwake, double *gridCentre, double *velCentre);

/* ---- allocation / raw views for the Python test harness ---- */
orc_rotor_t *orc_rotor_new(int nb, int nc, int ns, int nNwake, int nFwake);
void orc_rotor_free(orc_rotor_t *r);
double *orc_rotor_wiP(orc_rotor_t *r, int ib);
double *orc_rotor_waN(orc_rotor_t *r, int ib, int predicted);
double *orc_rotor_waF(orc_rotor_t *r, int ib, int predicted);
double *orc_rotor_wapF(orc_rotor_t *r, int ib, int predicted);
double *orc_rotor_vel(orc_rotor_t *r, int ib, int which);
double *orc_rotor_AIC(orc_rotor_t *r, int inverse);
double *orc_rotor_vec(orc_rotor_t *r, int which);
void orc_rotor_dims(const orc_rotor_t *r, int *out);
void orc_rotor_gens(const orc_rotor_t *r, unsigned long out[3]); /* gen_wing, gen_wake C, gen_wake P */
void orc_rotor_set_rows(orc_rotor_t *r, int rowNear, int rowFar);
void orc_rotor_get_params(const orc_rotor_t *r, double *out18);
void orc_rotor_set_params(orc_rotor_t *r, int surfaceType, int axisymmetrySwitch, int nbConvect, double Omega,
                          double omegaSlow, const double *shaftAxis, const double *hubCoords, double theta0,
                          double apparentViscCoeff, double decayCoeff, int rollupStart, int rollupEnd);
/* what: 0 vind_bywing, 1 vind_bywake, 2 both, 3 vind_bywing_boundVortices */
void orc_rotor_vind_points(const orc_rotor_t *r, int what, int predicted, long m, const double *P, double *V);

#ifdef __cplusplus
}
#endif
#endif
