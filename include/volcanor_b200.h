/*
 * volcanor_b200.h -- C ABI of the B200-native Biot-Savart hot path of VOLCANOR.
 *
 * This is the drop-in boundary.  The reference (cibinjoseph/VOLCANOR, Fortran 2008 + OpenMP,
 * one statically linked program) has no FFI of its own; the entry points below are what an
 * `iso_c_binding` shim (fortran/libGPU.f90) binds so that the driver's
 * call sites stay unchanged.  Each entry point names the reference procedure it replaces.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, no torch / C++ types.
 *   - every function returns 0 on success, non-zero on failure; vlc_last_error(ctx) gives the
 *     message (the Fortran shim turns it into `error stop`, like the reference's own
 *     `error stop` sites, e.g. libCommon.f90:168, libMath.f90:73).
 *   - arrays are column-major exactly as Fortran passes them: a `real(dp) :: P(3, m)` is
 *     `const double P[3*m]` with xyz fastest.
 *   - "records" are the reference's derived types viewed as arrays of doubles
 *     (`transfer(blade%waN, buf)`): vr_class = 50 doubles (classdef.f90:81-104), Fwake_class = 13
 *     doubles (:198-220), wingpanel_class = 104 doubles (:106-179).  See VLC_*_DOUBLES.
 *   - host-pointer entry points copy in and out synchronously; `_dev` entry points take device
 *     pointers and are asynchronous on the context's stream (vlc_set_stream / vlc_sync).
 *   - there is no CPU fallback: without a CUDA device vlc_create fails.
 *   - the caller is single-threaded per context (the reference calls these sites from the
 *     master thread, outside OpenMP regions).
 */
#ifndef VOLCANOR_B200_H
#define VOLCANOR_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct vlc_ctx vlc_ctx;

#define VLC_VF_DOUBLES 12         /* vf_class        classdef.f90:57-79   */
#define VLC_VR_DOUBLES 50         /* vr_class        classdef.f90:81-104  */
#define VLC_FWAKE_DOUBLES 13      /* Fwake_class     classdef.f90:198-220 */
#define VLC_WINGPANEL_DOUBLES 104 /* wingpanel_class classdef.f90:106-179 */
#define VLC_NPFWAKE 240           /* pFwake_class    classdef.f90:225     */
#define VLC_MAX_SETS 8

enum {
  VLC_OK = 0,
  VLC_ERR_CUDA = 1,     /* CUDA runtime / cuSOLVER failure */
  VLC_ERR_ARG = 2,      /* bad argument */
  VLC_ERR_STATE = 3,    /* call order (e.g. solve before calcAIC) */
  VLC_ERR_SINGULAR = 4, /* 'Matrix is numerically singular!' libMath.f90:73 */
  VLC_ERR_NODEVICE = 5
};

/* ---- context ---------------------------------------------------------------------------- */
int vlc_create(int device, vlc_ctx** out);
int vlc_destroy(vlc_ctx* ctx);
/*
 * Several GPUs of one box behind ONE handle (SURVEY 8b "library owns ... NCCL communicators", 8e).  The reference is a
 * single process whose wake loops call vind_onNwake_byRotor target by target (main.f90:814-841 ->
 * libCommon.f90:114-171); its call sites cannot hand out slices.  vlc_create_multi returns a context that LOOKS like a
 * single-GPU one: every entry point that changes state is replicated on all n_devices members (each holds the whole
 * wake; the O(N) mutators run redundantly), every host-pointer sweep (vlc_vind, vlc_rotor_vind_*,
 * vlc_vind_on{N,F}wake_byRotor, vlc_gridgen) gives each member a contiguous slice of the targets against ALL sources,
 * and vlc_wake_sweep of the resident path all-gathers the velocity slices -- ncclAllGather over NVLink, 24 bytes per
 * target, once per predictor and once per corrector stage -- so that the replicas stay bit-identical.  Readers
 * (vlc_rotor_get_*, ..._out arguments) are served by the first member.  Entry points that take DEVICE pointers address
 * one device and return VLC_ERR_STATE on such a handle.  devices == NULL means 0..n_devices-1.  A list that repeats a
 * device (group logic on a single-GPU box) and VLC_GROUP_NCCL=0 exchange the slices with peer copies instead of NCCL.
 * One worker thread per member issues its launches, so the caller stays single-threaded.
 */
int vlc_create_multi(int n_devices, const int* devices, vlc_ctx** out);
/*
 * The same partition with ONE PROCESS PER GPU (torchrun, or an MPI build of the driver): rank 0 obtains a 128-byte id
 * (ncclGetUniqueId), the launcher distributes it, every rank joins with its own single-GPU context.  From then on
 * vlc_wake_sweep and the host-pointer sweeps are COLLECTIVE: every rank calls them with the same arguments, sweeps its
 * slice, and all ranks end up with the complete result (ncclAllGather inside the library).  world = 1 leaves the
 * context as it is.  The communicator is destroyed by vlc_destroy.
 */
int vlc_comm_unique_id(void* id128);
int vlc_comm_init_rank(vlc_ctx* ctx, int world, int rank, const void* id128);
/* world, rank of this handle in the target partition; transport: 0 = single GPU, 1 = NCCL, 2 = peer copies */
int vlc_comm_info(const vlc_ctx* ctx, int* world, int* rank, int* transport);
const char* vlc_last_error(const vlc_ctx* ctx); /* ctx may be NULL: error of the last failed vlc_create */
const char* vlc_version(void);
/* Run on the caller's cudaStream_t (cuda_stream == NULL is the legacy default stream, which is what
 * torch.cuda.current_stream().cuda_stream returns for the default stream); use_own != 0 switches back to
 * the context's own non-blocking stream (the default after vlc_create). */
int vlc_set_stream(vlc_ctx* ctx, void* cuda_stream, int use_own);
int vlc_sync(vlc_ctx* ctx);
/* sm_count, compute capability major/minor, bytes of device memory */
int vlc_device_info(vlc_ctx* ctx, int* sm_count, int* cc_major, int* cc_minor, int64_t* mem_bytes);
/* Launch shape of the sweep kernel: targets_per_thread in {1,2,3,4} (0 = auto),
 * nsplit = source splits (0 = auto).  Only affects speed and summation order, never the pair formula. */
int vlc_set_tuning(vlc_ctx* ctx, int targets_per_thread, int nsplit);
/* Arithmetic of the two reciprocal square roots per pair (everything else is identical):
 *   0 = full: MUFU.RSQ64H seed + third-order Newton step, pair error ~1e-16 (default);
 *   1 = fast: second-order step, rsqrt error <= 6.4e-13 (2 FP64 instructions fewer per pair, ~6 % faster).
 *       OUTSIDE the 1e-12 per-target tolerance: the error enters the two end-point terms of a pair separately and
 *       their difference cancels away from the filament (measured 1.4e-12 of a target's own scale) -> opt-in only,
 *       never used for a parity or benchmark claim (DESIGN.md "Precision modes"). */
int vlc_set_precision(vlc_ctx* ctx, int mode);
/* Kernels launched by this context since creation (for bench.py's gpu_launches). */
int64_t vlc_launch_count(const vlc_ctx* ctx);

/* ---- tier 1: flat filament sets --------------------------------------------------------- */
/*
 * Load source set `set` (0 <= set < VLC_MAX_SETS) with n straight vortex filaments.
 * p1, p2: (3, n) end points = vf%fc(:,1), vf%fc(:,2); rvc: core radius vf%rVc; gam: circulation of
 * the ring / far-wake element that owns the filament (negative for the horseshoe correction,
 * classdef.f90:1461).  wake_flag (may be NULL = all 0): non-zero applies the wake rule
 * `abs(gam) > eps` (classdef.f90:1452, :1466, :1472); wing filaments are never skipped (:1350-1355).
 */
int vlc_set_sources(vlc_ctx* ctx, int set, int64_t n, const double* p1, const double* p2, const double* rvc,
                    const double* gam, const uint8_t* wake_flag);
int vlc_set_sources_dev(vlc_ctx* ctx, int set, int64_t n, const double* d_p1, const double* d_p2,
                        const double* d_rvc, const double* d_gam, const uint8_t* d_wake_flag);
int64_t vlc_num_sources(const vlc_ctx* ctx, int set);
/*
 * V(:, t) = sum_k gam_k * vf_vind(filament_k, P(:, t))   -- vf_vind = classdef.f90:476-503.
 * One call replaces one OpenMP target loop (libCommon.f90:132-146, :190-195; main.f90:528-573).
 * P, V: (3, m).  V is overwritten.
 */
int vlc_vind(vlc_ctx* ctx, int set, int64_t m, const double* P, double* V);
int vlc_vind_dev(vlc_ctx* ctx, int set, int64_t m, const double* d_P, double* d_V);
/* Only sources [first, first+count) of the set; first must be a multiple of vlc_source_tile(). */
int vlc_vind_range_dev(vlc_ctx* ctx, int set, int64_t first, int64_t count, int64_t m, const double* d_P,
                       double* d_V);
int vlc_source_tile(void);

/* ---- tier 2: the reference's rotor-level call sites ---------------------------------------- */
/* Declare rotor ir (0-based) -- sizes as rotor_class (classdef.f90:363): nb blades, nc x ns wing
 * panels, nNwake near-wake rows, nFwake far-wake rows; surfaceType as rotor%surfaceType (+-1 lifting,
 * +-2 non-lifting: vind_bywing returns 0, classdef.f90:4437-4441 + :544-554). */
int vlc_rotor_define(vlc_ctx* ctx, int ir, int nb, int nc, int ns, int nNwake, int nFwake, int surfaceType);
/* Forget every rotor declared so far (their device buffers are kept for re-use).  The sweeps of tiers 2b / 2c sum over
 * ALL declared rotors, like the reference's loops over rotor(1:nr): a context that is re-used for another configuration
 * starts from here (the shim's gpu_init calls it before declaring the rotors of config.nml). */
int vlc_rotors_clear(vlc_ctx* ctx);
/* rotor%rowNear, rotor%rowFar (1-based, as in the reference, main.f90:412-417). */
int vlc_rotor_set_rows(vlc_ctx* ctx, int ir, int rowNear, int rowFar);
/* Upload blade ib (0-based) state in reference record layout. predicted != 0 -> waNPredicted etc.
 * The pointers address the WHOLE arrays (waN(1,1), waF(1)); only the active rows rowNear..nNwake / rowFar..nFwake of
 * the last vlc_rotor_set_rows travel (no sweep reads the others), so set the rows first. */
int vlc_rotor_put_wing(vlc_ctx* ctx, int ir, int ib, const double* wiP /* nc*ns x 104 */);
int vlc_rotor_put_nwake(vlc_ctx* ctx, int ir, int ib, int predicted, const double* waN /* nNwake*ns x 50 */);
int vlc_rotor_put_fwake(vlc_ctx* ctx, int ir, int ib, int predicted, const double* waF /* nFwake x 13 */);
int vlc_rotor_put_pfwake(vlc_ctx* ctx, int ir, int ib, int predicted, const double* wapF /* 240 x 13 */);
/* Only the circulations of the wing rings: gam (nc*ns) for blade ib = rotor_map_gam (classdef.f90:4181). */
int vlc_rotor_put_wing_gam(vlc_ctx* ctx, int ir, int ib, const double* gam);

/* = rotor%vind_bywing(P) classdef.f90:4424, batched over m points */
int vlc_rotor_vind_bywing(vlc_ctx* ctx, int ir, int64_t m, const double* P, double* V);
/* = rotor%vind_bywake(P [, 'P']) classdef.f90:4459 */
int vlc_rotor_vind_bywake(vlc_ctx* ctx, int ir, int predicted, int64_t m, const double* P, double* V);
/* = rotor%vind_bywing_boundVortices(P) classdef.f90:4445 */
int vlc_rotor_vind_bywing_boundVortices(vlc_ctx* ctx, int ir, int64_t m, const double* P, double* V);
/* = sum over blades of blade%vind_bywing_chordwiseVortices(P) classdef.f90:1398-1418: (vf(1) + vf(3))*gam of every wing
 * ring plus vf(2)*gam of the trailing-edge row -- the complement of the bound vortices: bywing = boundVortices +
 * chordwiseVortices.  (The reference evaluates it only for `velInduced`, whose consumers are commented out,
 * classdef.f90:1733-1735, :1802-1806; provided so that SURVEY 8a row a6 is complete.) */
int vlc_rotor_vind_bywing_chordwiseVortices(vlc_ctx* ctx, int ir, int64_t m, const double* P, double* V);
/* = vind_bywing(P) + vind_bywake(P[, 'P']) of rotor ir in one sweep */
int vlc_rotor_vind(vlc_ctx* ctx, int ir, int predicted, int64_t m, const double* P, double* V);
/*
 * = vind_onNwake_byRotor(rotor(ir), Nwake[, 'P'])  libCommon.f90:114-171.
 * Nwake: slice of Nwake_class records, element (i, j) at Nwake + 50*((i-1) + ld*(j-1)),
 * rows x cols; vindArray: (3, rows, cols+1).
 */
int vlc_vind_onNwake_byRotor(vlc_ctx* ctx, int ir, const double* Nwake, int rows, int cols, int ld, int predicted,
                             double* vindArray);
/* = vind_onFwake_byRotor(rotor(ir), Fwake[, 'P'])  libCommon.f90:173-211; vindArray (3, rows). */
int vlc_vind_onFwake_byRotor(vlc_ctx* ctx, int ir, const double* Fwake, int rows, int predicted,
                             double* vindArray);

/* = rotor%calcAIC() classdef.f90:4151-4179: assembles AIC on the device from the uploaded wing
 * (CP, nCap, vortex rings), LU-factors it (cuSOLVER getrf) and forms AIC_inv once (getrs on the identity: the
 * reference's inv2 = DGETRF+DGETRI, libMath.f90:48-83, classdef.f90:4178).
 * AIC_out (N x N column-major, N = nc*ns*nb) may be NULL. */
int vlc_rotor_calcAIC(vlc_ctx* ctx, int ir, double* AIC_out);
/* gamVec = matmulAX(AIC_inv, RHS) (main.f90:190, :596; libMath.f90:105-122): one GEMV with the stored inverse. */
int vlc_rotor_solve(vlc_ctx* ctx, int ir, const double* RHS, double* gamVec);
/* AIC_inv (N x N) if the caller wants the explicit inverse the reference stores. */
int vlc_rotor_get_AIC_inv(vlc_ctx* ctx, int ir, double* AIC_inv);

/* ---- tier 2b: the reference's wake mutators on the device copies (device-resident time stepping) ------------- */
/*
 * With these the wake records uploaded once by vlc_rotor_put_nwake / _put_fwake never travel again: every procedure
 * of the reference's time loop that changes the wake (main.f90:466-506, :800-1440) has a device twin that works on the
 * library's copies of waN / waF / waNPredicted / waFPredicted in the reference's own record layout, and the velocity
 * arrays velNwake, velNwake1, velNwakePredicted, velNwakeStep, velFwake... (classdef.f90:285-292) live on the device
 * too.  The driver keeps its row counters (vlc_rotor_set_rows) and its wing (vlc_rotor_put_wing: the shed edge and its
 * circulation are read from it).  Arithmetic and statement order follow the reference; results are bit-identical to
 * the CPU restatement for the same inputs (tests/test_gpu_resident.py).  nNwakeEnd = nNwake and nFwakeEnd = nFwake
 * (classdef.f90:3054-3055; the prescribed-wake generator that would change them is out of scope).
 */
/* rotor_class members the mutators read: nbConvect (classdef.f90:3041-3045), axisymmetrySwitch, ductSwitch,
 * suppressFwakeSwitch, rollupStart / rollupEnd (1-based columns, :3135-3136), rollupSign = Omega*controlPitch(1) (only
 * its sign is used, :4541), apparentViscCoeff, decayCoeff, initWakeVel. */
int vlc_rotor_set_wake_params(vlc_ctx* ctx, int ir, int nbConvect, int axisymmetrySwitch, int ductSwitch,
                              int suppressFwakeSwitch, int rollupStart, int rollupEnd, double rollupSign,
                              double apparentViscCoeff, double decayCoeff, double initWakeVel);
/* rotor%shaftAxis, rotor%hubCoords of the current step (they move with the body, main.f90:455-463). */
int vlc_rotor_set_frame(vlc_ctx* ctx, int ir, const double* shaftAxis, const double* hubCoords);
/* = rotor%assignshed('LE' | 'TE') classdef.f90:4297-4325; edge 0 = 'LE', 1 = 'TE'.  Reads the uploaded wing. */
int vlc_rotor_assignshed(vlc_ctx* ctx, int ir, int edge);
/* = rotor%age_wake(dt) classdef.f90:4331-4354 (omegaSlow = rotor%omegaSlow) */
int vlc_rotor_age_wake(vlc_ctx* ctx, int ir, double dt, double omegaSlow);
/* = rotor%dissipate_wake(dt, kinematicVisc) classdef.f90:4356-4408 */
int vlc_rotor_dissipate_wake(vlc_ctx* ctx, int ir, double dt, double kinematicVisc);
/* = rotor%strain_wake() classdef.f90:4410-4422 */
int vlc_rotor_strain_wake(vlc_ctx* ctx, int ir);
/* = rotor%burst_wake() classdef.f90:4911-4917 (blade_burst_wake :2306-2339; the driver calls it every wakeBurst-th step,
 * main.f90:490-497): where successive far-wake filaments of the current wake kink by skewLimit or more (skew =
 * |angle - pi|/pi), both get the core radius largeCoreRadius (= rotor%chord). */
int vlc_rotor_burst_wake(vlc_ctx* ctx, int ir, double skewLimit, double largeCoreRadius);
/* = rotor%calc_skew() classdef.f90:4919-4936 (the driver calls it every skewPlotSwitch-th step before skew2file,
 * main.f90:499-504): vr%skew of the active near-wake rings of the current wake; vlc_rotor_get_nwake brings it back. */
int vlc_rotor_calc_skew(vlc_ctx* ctx, int ir);
/* waNPredicted(rowNear:, :) = waN(rowNear:, :), waFPredicted(rowFar:) = waF(rowFar:) of the convected blades
 * (main.f90:869-872, :1028-1030) */
int vlc_rotor_wake_to_predicted(vlc_ctx* ctx, int ir);
/* = rotor%convectwake(iter, dt, 'C' | 'P') classdef.f90:4786-4830 with the device's velNwake / velFwake: shift the
 * corners (blade_convectwake :1515-1575, the predictor's loop quirk included), re-stitch (wake_continuity
 * :1609-1702), axisymmetric copy + rotate of blades 2..nb (:4801-4823). */
int vlc_rotor_convectwake(vlc_ctx* ctx, int ir, double dt, int predicted);
/* = rotor%updatePrescribedWake(dt, 'C' | 'P') classdef.f90:5170-5218, the statement that ends rotor%convectwake when
 * prescWakeNt > 0 and iter > prescWakeNt (:4826-4828): per convected blade pFwake_update (:998-1066) -- a helix of 240
 * filaments fitted to the far-wake rows rowStart..nFwake of the device's waF / waFPredicted (rowStart = rowFar when
 * prescWakeGenNt == 0, else nFwake - prescWakeGenNt), relaxed against the blade's previous fit, which the device keeps
 * per record set -- then the axisymmetric copies of blade 1's (pFwake_rot_wake_axis :1068-1086).  deltaPsi = omegaSlow*dt.
 * From then on the helix is a source of the sweeps of that record set.  VLC_ERR_ARG when the shaft is not along z (the
 * reference's error stop), VLC_ERR_STATE without a far-wake row to fit. */
int vlc_rotor_updatePrescribedWake(vlc_ctx* ctx, int ir, double deltaPsi, int prescWakeGenNt, int predicted);
/* Device -> host copy of a blade's prescribed wake records (wake plots, libPostprocess.f90:266-284; restart files);
 * helix, when not NULL, receives (helixPitch, helixRadius) of that blade and record set. */
int vlc_rotor_get_pfwake(vlc_ctx* ctx, int ir, int ib, int predicted, double* wapF /* 240 x 13 */, double* helix /* 2 or NULL */);
/* Host -> device: (helixPitch, helixRadius) of a blade's prescribed wake, the state the relaxation of the next
 * vlc_rotor_updatePrescribedWake starts from (zero after vlc_rotor_define, like pFwake_class classdef.f90:228-229) -- with
 * vlc_rotor_put_pfwake what a driver resuming from a restart file sends. */
int vlc_rotor_put_pfwake_helix(vlc_ctx* ctx, int ir, int ib, int predicted, const double* helix /* 2 */);
/* = rotor%rollup() classdef.f90:4515-4605 (shiftFwake :4500-4513 when the far wake is full, then shiftwake :4481-4498);
 * the driver calls it when rowNear == 1 (main.f90:1424-1425). */
int vlc_rotor_rollup(vlc_ctx* ctx, int ir);
/* The wake sweeps of the convection driver for ALL rotors at once (main.f90:800-838; 'P': :889-911, :1057-1081):
 * vel{N,F}wake[Predicted](active rows) of every convected blade <- sum over source rotors of
 * vind_on{N,F}wake_byRotor; addInitWakeVel != 0 adds -/+ initWakeVel*shaftAxis with the reference's signs. */
int vlc_wake_sweep(vlc_ctx* ctx, int predicted, int addInitWakeVel);
/* The same in three parts, for one process per GPU (targets are independent, libCommon.f90:132-139): every rank keeps
 * the whole wake (the O(N) mutators above run redundantly), sweeps only ITS slice of the target list against all
 * sources, the ranks all-gather the velocity slices (NCCL, 24 bytes per target), and every rank scatters the complete
 * list into its velocity arrays -- one exchange per wake sweep, i.e. one per predictor and one per corrector stage.
 *   vlc_wake_sweep_count  : M = wake-node + far-wake targets of all convected blades at the current row counters
 *   vlc_wake_sweep_slice  : velocities of targets [first, first+count) into d_vel(3, M) (device) at those positions
 *   vlc_wake_sweep_scatter: d_vel(3, M) -> velNwake / velFwake [Predicted] (+- initWakeVel), as vlc_wake_sweep does
 * vlc_wake_sweep(ctx, p, a) == count; slice(0, M) into an internal buffer; scatter. */
int vlc_wake_sweep_count(vlc_ctx* ctx, int64_t* M);
int vlc_wake_sweep_slice(vlc_ctx* ctx, int predicted, int64_t first, int64_t count, double* d_vel);
int vlc_wake_sweep_scatter(vlc_ctx* ctx, int predicted, int addInitWakeVel, const double* d_vel);
/* Bookkeeping of the velocity arrays between the sweeps (convected blades, whole arrays like the reference). */
enum {
  VLC_VEL_FIRST_STEP = 0,    /* main.f90:1013-1020  vel1 = vel                                         */
  VLC_VEL_AB2 = 1,           /* main.f90:1031-1041  velStep = vel; vel = 0.5*(3*vel - vel1)            */
  VLC_VEL_AM2 = 2,           /* main.f90:1094-1099  vel = (velPredicted + velStep)*0.5                 */
  VLC_VEL_SHIFT_HISTORY = 3, /* main.f90:1103-1107  vel1 = velStep                                     */
  VLC_VEL_ORDER2 = 4,        /* main.f90:927-940    vel(active) = vel_order2(vel, velPredicted)        */
  VLC_VEL_COPY_TO_STEP = 5   /* velStep = vel; fdScheme 2 (explicit Adams-Bashforth, main.f90:975-988) is
                                AB2, FIRST_STEP, COPY_TO_STEP: velStep = vel1 = vel = 0.5*(3*vel - vel1)   */
};
int vlc_rotor_wakevel_op(vlc_ctx* ctx, int ir, int op);
/* The same bookkeeping in general form, for the multistep schemes fdScheme 4 / 5 (main.f90:1117-1404), which keep two
 * more histories per blade (velNwake2 / 3, velFwake2 / 3, classdef.f90:3733-3824; here allocated on first use).
 * Arrays by id, near and far wake together, whole arrays of the convected blades: */
enum {
  VLC_VEL_ARRAY = 0, VLC_VEL_ARRAY_1 = 1, VLC_VEL_ARRAY_PREDICTED = 2, VLC_VEL_ARRAY_STEP = 3, VLC_VEL_ARRAY_2 = 4,
  VLC_VEL_ARRAY_3 = 5
};
/* dst = src */
int vlc_rotor_wakevel_copy(vlc_ctx* ctx, int ir, int dst, int src);
/* dst = (coef[0]*src[0] + ... + coef[nterms-1]*src[nterms-1]) / divisor, 1 <= nterms <= 4, terms added left to right, e.g.
 * the third-order predictor vel = (23*vel - 16*vel2 + 5*vel1)/12 (main.f90:1160-1163) is dst = VLC_VEL_ARRAY,
 * src = {ARRAY, ARRAY_2, ARRAY_1}, coef = {23, -16, 5}, divisor = 12.  dst may be one of the sources. */
int vlc_rotor_wakevel_lincomb(vlc_ctx* ctx, int ir, int dst, int nterms, const int* src, const double* coef, double divisor);
/* Read the device copies back (plots, restart files, tests): whole arrays of blade ib in the reference layout.
 * which: 0 = velNwake/velFwake, 1 = ...1, 2 = ...Predicted, 3 = ...Step, 4 = ...2, 5 = ...3; either pointer may be NULL. */
int vlc_rotor_get_nwake(vlc_ctx* ctx, int ir, int ib, int predicted, double* waN /* nNwake*ns x 50 */);
int vlc_rotor_get_fwake(vlc_ctx* ctx, int ir, int ib, int predicted, double* waF /* nFwake x 13 */);
int vlc_rotor_put_wakevel(vlc_ctx* ctx, int ir, int ib, int which, const double* velN, const double* velF);
int vlc_rotor_get_wakevel(vlc_ctx* ctx, int ir, int ib, int which, double* velN /* 3 x nNwake x (ns+1) */,
                          double* velF /* 3 x nFwake */);

/* ---- tier 2c: the collocation-point stage on the device copies of the wing records ---------------------------- */
/*
 * The rest of a time step between the two wake stages (main.f90:522-670): velocity at the collocation points, the
 * right-hand side, the solve, map_gam and -- when the driver wants forces -- velCPTotal and the circulation-based
 * sectional loads of forceCalcSwitch = 0.  The driver moves the wing and writes the KINEMATIC part of velCP into the
 * wing records (main.f90:528-547: -velBody - omegaBody x r - omegaSlow*shaftAxis x r - flap term) before
 * vlc_rotor_put_wing; after that neither the collocation-point velocities nor the right-hand side cross the bus.
 * Needs vlc_rotor_set_wake_params (nbConvect, axisymmetrySwitch).  The O(nc*ns) arithmetic is unfused IEEE in the
 * reference's statement order: given the same velCPTotal the loads are bit-identical to the CPU restatement
 * (tests/test_cp_stage_host.py without a GPU on the same source, tests/test_zz_gpu_cp_stage.py through this ABI).
 */
/* main.f90:548-603 for rotor ir: wingpanel%velCP of its convected blades += for jr = 1..nr: rotor(jr)%vind_bywake(CP)
 * [+ rotor(jr)%vind_bywing(CP) if jr /= ir]; RHS = -dot(velCP, nCap), blade 1's entries repeated for an axisymmetric
 * rotor.  velCP stays in the device records, RHS on the device for vlc_rotor_solve_map_gam.  Optional host copies:
 * velCP_out (3, nbConvect*nc*ns) in wiP order, RHS_out (nc*ns*nb); either may be NULL. */
int vlc_rotor_calc_RHS(vlc_ctx* ctx, int ir, double* velCP_out, double* RHS_out);
/* Sub-iterations (switches%ntSub > 0, main.f90:522-615): every pass of ntSubLoop starts velCP again from its kinematic
 * part (:528-547), which the records keep in velCPm (:545-546).  Pass i >= 1 of the loop is
 *   vlc_rotor_reset_velCP of every rotor; vlc_rotor_calc_RHS of every rotor; vlc_rotor_solve_map_gam of every rotor
 * (the wings of the other rotors then carry the circulation of pass i-1, as in the reference); the driver compares
 * gamVec with the previous pass (:606-613). */
int vlc_rotor_reset_velCP(vlc_ctx* ctx, int ir);
/* gamVec = matmulAX(AIC_inv, RHS) (main.f90:596; one GEMV with the inverse of vlc_rotor_calcAIC) followed by
 * rotor%map_gam() (classdef.f90:4181-4196) on the device records.  gamVec_out (nc*ns*nb) may be NULL. */
int vlc_rotor_solve_map_gam(vlc_ctx* ctx, int ir, double* gamVec_out);
/* Section frames of blade ib as the driver holds them after moving the wing, one block of 10*ns + 6 doubles:
 * secTauCapChord(3,ns) | secNormalVec(3,ns) | secCP(3,ns) | secArea(ns) | yAxisAziFlap(3) | zAxisAziFlap(3)
 * (blade_class, classdef.f90:238-358). */
int vlc_rotor_put_sections(vlc_ctx* ctx, int ir, int ib, const double* sec);
/* main.f90:630-663 for rotor ir: velCPTotal = velCP - sum over rotors of vind_bywing_boundVortices(CP)
 * + rotor(ir)%vind_bywing(CP); blade 1's values for every blade of an axisymmetric rotor. */
int vlc_rotor_calc_velCPTotal(vlc_ctx* ctx, int ir);
/* = rotor%calc_secAlpha() + rotor%calc_force(density, dt) for forceCalcSwitch = 0 (classdef.f90:4766-4784, :4607-4671
 * down to the blade sums of sumSecToNetForces :2368-2380; the sum over blades, sumBladeToNetForces :4954-4988, stays
 * with the driver).  Updates gamPrev, gamTrapz, delP, delPUnsteady, normalForce, normalForceUnsteady, chordwiseResVel
 * in the device records.  Omega = rotor%Omega (its sign sets invertGammaSign and the lift direction). */
int vlc_rotor_calc_force(vlc_ctx* ctx, int ir, double density, double dt, double Omega, int spanwiseLiftSwitch);
/* Loads of blade ib, one block of 12 + 25*ns doubles: forceInertial(3) lift(3) drag(3) liftUnsteady(3) |
 * secChordwiseResVel secDragDir secLiftDir secForceInertial secLift secDrag secLiftUnsteady (3,ns each) |
 * secAlpha secCL secCD secCLu (ns each). */
int vlc_rotor_get_loads(vlc_ctx* ctx, int ir, int ib, double* loads);
/* The device copy of blade ib's wing records (nc*ns x 104), e.g. to carry gamPrev / delP back into the driver's wiP. */
int vlc_rotor_get_wing(vlc_ctx* ctx, int ir, int ib, double* wiP);

/* ---- tier 3: device-resident wake state (node-indexed SoA) -------------------------------- */
/*
 * A "lattice" is one blade's near wake kept on the device as TE nodes: nodes(3, nrows+1, ns+1)
 * where node row r (0-based) is the leading edge of ring row r and the trailing edge of row r-1;
 * ring (r, j) has corners 1=(r,j) 2=(r+1,j) 3=(r+1,j+1) 4=(r,j+1) (vr_assignP, classdef.f90:569-592),
 * so the re-stitch of blade_wake_continuity (classdef.f90:1609-1702) is implicit.
 * Declared in wake_state section of DESIGN.md; entry points below operate on flat device arrays so
 * that bench.py / the multi-GPU driver can shard targets and all-gather node slices.
 */
/* x(3,n) += v(3,n) * dt   -- vr_shiftdP / Fwake_shiftdP with U*dt (classdef.f90:1531, :1545) */
int vlc_convect_dev(vlc_ctx* ctx, int64_t n, double* d_x, const double* d_v, double dt);
/* Adams-Bashforth predictor velocity: out = 0.5*(3 v - v1)  (main.f90:1032-1034) */
int vlc_ab2_dev(vlc_ctx* ctx, int64_t n, const double* d_v, const double* d_v1, double* d_out);
/* Adams-Moulton corrector velocity: out = (vp + vs) * 0.5   (main.f90:1094-1096) */
int vlc_am2_dev(vlc_ctx* ctx, int64_t n, const double* d_vp, const double* d_vs, double* d_out);
/* Predictor-corrector (fdScheme 1) velocity: vel_order2_Nwake / vel_order2_Fwake (libCommon.f90:213-258) on
 * (3, rows, cols) arrays (cols = 1 for the far wake); out must not alias the inputs. */
int vlc_vel_order2_dev(vlc_ctx* ctx, int rows, int cols, const double* d_vn, const double* d_vnp1, double* d_out);
/* rVc <- sqrt(rVc^2 + 4*1.2564*apparentViscCoeff*nu*dt); gam <- gam*exp(-decayCoeff*dt)
 * (rotor_dissipate_wake classdef.f90:4356-4408, vr_decay :662-668).  Either pointer may be NULL. */
int vlc_dissipate_dev(vlc_ctx* ctx, int64_t n_rvc, double* d_rvc, int64_t n_gam, double* d_gam,
                      double apparentViscCoeff, double kinematicVisc, double decayCoeff, double dt);
/* far-wake strain: lc = |p1-p2|, rVc = rVc0*sqrt(l0/lc)  (classdef.f90:4410-4422, :505-521) */
int vlc_strain_dev(vlc_ctx* ctx, int64_t n, const double* d_p1, const double* d_p2, const double* d_l0,
                   const double* d_rvc0, double* d_rvc);
/*
 * Build the packed source set `set` from a near-wake lattice + far-wake chain on the device, in the
 * reference's enumeration order (blade_vind_bywake, classdef.f90:1450-1469): rings (j, i) x 4 filaments,
 * then (if nfar > 0) the horseshoe correction -vf2 of the last row, then the far filaments.
 *   nodes (3, nrows+1, ns+1); gam (nrows, ns); rvc4 (4, nrows, ns) = vf(1..4)%rVc of every ring (kept per
 *   ring, not per edge, because the reference lets the two copies of a shared edge differ);
 *   far: p(3, nfar+1) chain with filament i = fc(:,2)=p(:,i) -> fc(:,1)=p(:,i+1), gamF(nfar), rvcF(nfar).
 *   append != 0 adds to the set instead of replacing it (several blades / rotors).
 */
int vlc_pack_lattice_dev(vlc_ctx* ctx, int set, int append, int nrows, int ns, const double* d_nodes,
                         const double* d_gam, const double* d_rvc4, int nfar, const double* d_far_nodes,
                         const double* d_gamF, const double* d_rvcF);
/* Same with HOST arrays (copied in synchronously): the caller-facing form of the lattice path. */
int vlc_pack_lattice(vlc_ctx* ctx, int set, int append, int nrows, int ns, const double* nodes, const double* gam,
                     const double* rvc4, int nfar, const double* far_nodes, const double* gamF, const double* rvcF);
/* Sets built by vlc_pack_lattice_dev also hold a SHARED-NODE form: every lattice node is evaluated once
 * per target instead of once per adjacent ring filament and every interior edge once with the merged strength of the
 * two rings that share it -- the same sum as the reference's ring-by-ring enumeration (classdef.f90:1450-1456), 18
 * instead of 43 FP64 instructions per reference pair.  If the two copies of a shared edge carry different core radii
 * the device falls back to the flat enumeration by itself.  on = 0 forces the flat enumeration (default on = 1). */
int vlc_set_shared_nodes(vlc_ctx* ctx, int on);
/* Launch shape of the shared-node kernel: strip_width W in 1..4 ring columns per strip record (wider strips
 * amortise the node work: 72 / 66.5 / 64.7 / 63.75 FP64 instructions per ring), targets_per_thread in 1..3;
 * 0 = default.  Takes effect at the next pack.  Only speed and summation order change.  A rotor's lattice whose column
 * count is not a multiple of 4 is covered by width-4 strips plus one narrower tail strip when the wake is large (>= 2e4
 * rings: the tail's extra launches then cost less than padded columns); strip_width = 5 asks for that cover at any size. */
int vlc_set_lattice_tuning(vlc_ctx* ctx, int strip_width, int targets_per_thread);
/* out[0..5]: out[0] = filaments in the reference's enumeration, out[1] = strip records and out[2] = remainder filaments of
 * the shared-node form (0 if the set has none), out[3] = 1 if the next sweep will use the shared-node kernel with merged
 * edge strengths, 2 if its dual form (rotor records whose streamwise edge copies carry different core radii: a non-uniform
 * streamwiseCoreVec, classdef.f90:3841, :4371-4372), 0 if the flat kernel on the reference's enumeration (reads the device
 * flag: synchronises), -1 for a flat-only set, out[4] = strip width of the records,
 * out[5] = width of the tail strips (0 = none): a rotor's lattice with ns columns, ns not a multiple of 4, is covered by
 * floor(ns/4) strips of width 4 plus one strip of width ns mod 4 when that is cheaper than padding the last strip. */
int vlc_set_info(vlc_ctx* ctx, int set, int64_t* out);
/* Same six numbers for the packed [wing | wake] set of rotor ir (packs it if needed): tells whether the uploaded
 * records describe a lattice the shared-node kernel can use (out[3] = 1) or the flat enumeration is used (0). */
int vlc_rotor_info(vlc_ctx* ctx, int ir, int predicted, int64_t* out);
/* rotor_dissipate_wake on a lattice (classdef.f90:4364-4393): vf1 grows, vf3 <- vf1, gam decays, vf2 grows,
 * vf4(i) <- vf2(i-1) for i > first row. */
int vlc_dissipate_lattice_dev(vlc_ctx* ctx, int nrows, int ns, double* d_rvc4, double* d_gam,
                              double apparentViscCoeff, double kinematicVisc, double decayCoeff, double dt);

/* Convected nodes of a lattice <-> contiguous target list, in the order of vind_onNwake_byRotor
 * (libCommon.f90:133-145: column-major over (row, col) incl. the extra last column):
 *   gather : P(3, nrows, ns+1) = nodes(:, 2:nrows+1, :)        scatter : the inverse copy. */
int vlc_lattice_targets_dev(vlc_ctx* ctx, int nrows, int ns, const double* d_nodes, double* d_P);
int vlc_lattice_scatter_dev(vlc_ctx* ctx, int nrows, int ns, double* d_nodes, const double* d_P);

/* ---- the Eulerian grid tool ------------------------------------------------------------------ */
/*
 * = program gridgen (src/gridgen.f90:62-145): velocity at the (nx-1)(ny-1)(nz-1) cell centres of the Cartesian grid
 * xyzMin..xyzMax induced by the filaments of one filamentsNNNNN.dat (written by filaments2file,
 * src/libPostprocess.f90:363-473), plus the free stream `vel`.  vrWing / vrNwake: vr_class records (50 doubles);
 * vfNwakeTE / vfFwake: vf_class records (12 doubles) with their circulation arrays.  The file's arithmetic is kept,
 * including its quirk that the wing loop assigns instead of accumulating (:122: only the last wing ring counts)
 * and that no |gam| > eps rule applies here.  gridCentre (may be NULL) and velCentre: (3, nx-1, ny-1, nz-1).
 * Uses source set VLC_MAX_SETS-1 as scratch.
 */
int vlc_gridgen(vlc_ctx* ctx, int nx, int ny, int nz, const double* xyzMin, const double* xyzMax, const double* vel,
                int64_t nVrWing, const double* vrWing, int64_t nVrNwake, const double* vrNwake, int64_t nVfNwakeTE,
                const double* vfNwakeTE, const double* gamNwakeTE, int64_t nVfFwake, const double* vfFwake,
                const double* gamFwake, double* gridCentre, double* velCentre);
/* The same for cells [first, first + count) of the cell list (x fastest, then y, then z): gridCentre (may be NULL) and
 * velCentre are (3, count).  One process per GPU takes a contiguous slice each (the cells are independent,
 * gridgen.f90:116-139) and the slices are gathered by the caller (volcanor_b200/gridgen.py under torchrun). */
int vlc_gridgen_slice(vlc_ctx* ctx, int nx, int ny, int nz, const double* xyzMin, const double* xyzMax, const double* vel,
                      int64_t nVrWing, const double* vrWing, int64_t nVrNwake, const double* vrNwake, int64_t nVfNwakeTE,
                      const double* vfNwakeTE, const double* gamNwakeTE, int64_t nVfFwake, const double* vfFwake,
                      const double* gamFwake, int64_t first, int64_t count, double* gridCentre, double* velCentre);

/* ---- measurement helper ------------------------------------------------------------------ */
/* Device-side timing across the library's own stream(s): vlc_event_record(slot) records a CUDA event (slot 0..7) on the
 * stream of the context -- of every member for a handle made by vlc_create_multi; vlc_event_elapsed_ms waits for slot_b
 * and returns the time from slot_a to slot_b, the LARGEST over the members of a group (the step is over when the slowest
 * device is done). */
int vlc_event_record(vlc_ctx* ctx, int slot);
int vlc_event_elapsed_ms(vlc_ctx* ctx, int slot_a, int slot_b, double* ms);
/* Roofline bookkeeping of the dominant kernels, measured with one CUDA event pair per launch on the launching stream:
 * reset > 0 starts collecting (and clears), reset < 0 stops.  Output arrays (any may be NULL): launches[2]; ms[4],
 * pairs[4], fp64_instr[4].  Entries [0] bs_lattice_kernel and [1] bs_sweep_kernel: the dominant launches since the last
 * reset, their summed device time in ms, the reference pair interactions they stand for (lattice: targets x 4 filaments x
 * rings; flat: targets x filaments) and the FP64-pipe instructions they issued (lattice: (11(W+1) + 50W) per (target, strip
 * record); flat: 43 per pair).  Entries [2], [3]: the same for the WHOLE sweeps those launches belong to (dominant kernel
 * + tail strips + flat remainder + reduce, window = first launch to end of the reduce): the remainder runs on a
 * low-priority side stream and its CTAs may share SMs with the dominant kernel, inside that kernel's own window.  Reads
 * the first member of a group handle (its slice of the targets). */
int vlc_sweep_stats(vlc_ctx* ctx, int reset, int64_t* launches, double* ms, double* pairs, double* fp64_instr);
/* memset of a 256 MiB scratch buffer on the context's stream(s): evicts the previous step's data from the 126 MB L2. */
int vlc_l2_flush(vlc_ctx* ctx);
/* Register-resident DFMA chains on every SM: returns sustained FP64 FMA rate in flop/s (DFMA = 2)
 * and the elapsed milliseconds; used by bench.py as the measured FP64 roofline denominator
 * (MEASURED_PEAKS.json has no FP64 entry). */
int vlc_measure_fp64_peak(vlc_ctx* ctx, int iters, double* flops_per_s, double* ms);
/* Same measurement with a chosen operand pattern: 0 = the chains above (one changing register operand per DFMA, the
 * other two a constant and a loop-invariant register: the pipe's best case), 1 = three distinct, changing register
 * operands per DFMA (a = b*c + a; b = c*a + b; ...), which is what the instructions of a real kernel present to the
 * register file.  bench.py reports pattern 1 next to the peak as the practical ceiling of register-fed FP64 code. */
int vlc_measure_fp64_rate(vlc_ctx* ctx, int iters, int pattern, double* flops_per_s, double* ms);
/* Device time of the last sweep launched by this context, from CUDA events recorded on the context's stream around
 * the dominant kernel (bs_lattice_kernel or bs_sweep_kernel) and around the whole sweep (incl. remainder + reduce).
 * Waits for that sweep to finish. */
int vlc_last_sweep_ms(vlc_ctx* ctx, double* ms_kernel, double* ms_total);
/* Evaluate the raw MUFU.RSQ64H seed and the full / fast refined reciprocal square roots of x[0..n)
 * (host arrays) -- lets the tests measure the error bounds the precision modes rely on. */
int vlc_probe_rsqrt(vlc_ctx* ctx, int64_t n, const double* x, double* seed, double* full, double* fast);

#ifdef __cplusplus
}
#endif
#endif /* VOLCANOR_B200_H */
