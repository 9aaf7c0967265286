/*
 * case_gpu_hooks.c -- TEST INFRASTRUCTURE: the C twin of fortran/libGPU.f90.
 *
 * Installs, into the oracle's restatement of the reference driver (oracle/vlc_case.c), a hook table whose five
 * hot-path call sites go straight to the C ABI of include/volcanor_b200.h -- upload the rotor's records
 * (gpu_sync_rotor), call the batched entry point, hand the velocities back -- with no Python in the loop.
 * tests/test_gpu_case.py uses it to run the reference's cases natively through the boundary (parity + per-step
 * wall time of the whole driver loop, i.e. what a Fortran user would see).
 *
 * Build: tests/native/Makefile (gcc; links libvlc_oracle.so and libvolcanor_b200.so).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/volcanor_b200.h"
#include "../../oracle/vlc_case.h"

#define MAX_ROTORS 64
typedef struct {
  orc_case_t *cas;
  vlc_ctx *ctx;
  int nr;
  long uploads, skipped;
  int last_rc;
  /* what the library currently holds: generation of the wing and of the 'C' / 'P' wake, and the row counters */
  unsigned long have_wing[MAX_ROTORS], have_wake[MAX_ROTORS][2];
  int have_rows[MAX_ROTORS][2][2];
  /* resident mode (tier 2b of the C ABI): the wake lives on the device; only the wing travels */
  int resident, resident_started;
  long wing_uploads;
  const struct resident_ops *ops; /* who executes the wake stages: the C ABI (GPU) or the oracle's own mutators (CPU) */
  const struct cp_ops *cp;        /* who executes the collocation-point stage (tier 2c): the C ABI or its CPU emulation */
  struct cp_emu *emu;             /* CPU emulation only: the "library's" copies per rotor */
  long cp_rhs_calls, cp_force_calls;
  /* one process per GPU (world > 1): this rank sweeps targets [rank*per, (rank+1)*per) and `exchange` all-gathers the
   * velocity slices in xbuf (device, (3, world*per_max)); set by case_gpu_hooks_set_sharding */
  int world, rank;
  double *xbuf;
  long xbuf_targets;
  int (*exchange)(void *arg, long per);
  void *exchange_arg;
  long exchanges;
} gpu_user_t;

/* The wake stages of one time step, as the orchestration below calls them.  Two backends: `gpu_ops` forwards each to the
 * C ABI's tier 2b; `cpu_ops` to the CPU restatement acting on the driver's own arrays -- with it the SAME orchestration
 * runs without a GPU and must reproduce the driver's inline time loop bit for bit (tests/test_staged_hooks.py). */
typedef struct resident_ops {
  int (*begin)(gpu_user_t *u);
  int (*sync)(gpu_user_t *u, int ir);
  int (*assignshed)(gpu_user_t *u, int ir, int edge);
  int (*age_wake)(gpu_user_t *u, int ir, double dt, double omegaSlow);
  int (*dissipate_wake)(gpu_user_t *u, int ir, double dt, double kinematicVisc);
  int (*strain_wake)(gpu_user_t *u, int ir);
  int (*wake_to_predicted)(gpu_user_t *u, int ir);
  int (*convectwake)(gpu_user_t *u, int ir, int iter, double dt, int predicted);
  int (*rollup)(gpu_user_t *u, int ir);
  int (*wake_sweep)(gpu_user_t *u, int predicted, int addInitWakeVel);
  int (*wakevel_op)(gpu_user_t *u, int ir, int op);
  /* the wake sweep in three parts for one process per GPU (vlc_wake_sweep_count / _slice / _scatter) + stream sync */
  int (*sweep_count)(gpu_user_t *u, int64_t *M);
  int (*sweep_slice)(gpu_user_t *u, int predicted, int64_t first, int64_t count, double *xbuf);
  int (*sweep_scatter)(gpu_user_t *u, int predicted, int addInitWakeVel, const double *xbuf);
  int (*stream_sync)(gpu_user_t *u);
  /* general velocity bookkeeping of the multistep schemes (fdScheme 4 / 5): array ids VLC_VEL_ARRAY_* */
  int (*wakevel_copy)(gpu_user_t *u, int ir, int dst, int src);
  int (*wakevel_lincomb)(gpu_user_t *u, int ir, int dst, int nterms, const int *src, const double *coef, double divisor);
  int (*burst_wake)(gpu_user_t *u, int ir); /* rotor%burst_wake(), every wakeBurst-th step (main.f90:490-497) */
} resident_ops_t;

#define CK(expr)                        \
  do {                                  \
    int rc_ = (expr);                   \
    if (rc_) return u->last_rc = rc_;   \
  } while (0)

/* resident mode: row counters, frame and (when the driver moved it or changed its circulation) the wing */
static int sync_wing(gpu_user_t *u, int jr) {
  orc_rotor_t *r = orc_case_rotor(u->cas, jr);
  CK(vlc_rotor_set_rows(u->ctx, jr, r->rowNear, r->rowFar));
  CK(vlc_rotor_set_frame(u->ctx, jr, r->shaftAxis, r->hubCoords));
  if (r->gen_wing != u->have_wing[jr]) {
    for (int ib = 0; ib < r->nb; ++ib) CK(vlc_rotor_put_wing(u->ctx, jr, ib, orc_rotor_wiP(r, ib)));
    u->have_wing[jr] = r->gen_wing;
    u->wing_uploads++;
  }
  return 0;
}

/* resident mode, first step: the wake records as rotor_init left them (core radii, the first shed edge) go up once,
 * whole arrays, current and predicted */
static int resident_begin(gpu_user_t *u) {
  for (int jr = 0; jr < u->nr; ++jr) {
    orc_rotor_t *r = orc_case_rotor(u->cas, jr);
    CK(vlc_rotor_set_wake_params(u->ctx, jr, r->nbConvect, r->axisymmetrySwitch, r->ductSwitch, r->suppressFwakeSwitch,
                                 r->rollupStart, r->rollupEnd, r->Omega * r->controlPitch[0], r->apparentViscCoeff,
                                 r->decayCoeff, r->initWakeVel));
    CK(vlc_rotor_set_rows(u->ctx, jr, 1, 1));
    for (int ib = 0; ib < r->nb; ++ib)
      for (int s = 0; s < 2; ++s) {
        if (r->nNwake > 0) CK(vlc_rotor_put_nwake(u->ctx, jr, ib, s, orc_rotor_waN(r, ib, s)));
        if (r->nFwake > 0) CK(vlc_rotor_put_fwake(u->ctx, jr, ib, s, orc_rotor_waF(r, ib, s)));
      }
    u->uploads++;
  }
  u->resident_started = 1;
  return 0;
}

/* = gpu_sync_rotor of fortran/libGPU.f90: row counters + wing + near/far wake records of every blade.  State the
 * driver has not touched since the last transfer (generation counters of oracle/vlc_case.c; a Fortran shim would set
 * the same flags where main.f90 calls convectwake / assignshed / dissipate_wake / map_gam / move) is not sent again. */
static int sync_rotor(gpu_user_t *u, int jr, int predicted) {
  orc_rotor_t *r = orc_case_rotor(u->cas, jr);
  int d[10], rc;
  unsigned long g[3];
  orc_rotor_dims(r, d);
  orc_rotor_gens(r, g);
  const int nb = d[0], nNwake = d[3], nFwake = d[4], s = predicted ? 1 : 0;
  if (u->resident && u->resident_started) return sync_wing(u, jr); /* the device's wake is the wake */
  if ((rc = vlc_rotor_set_rows(u->ctx, jr, d[5], d[6]))) return rc;
  const int wing_new = (g[0] != u->have_wing[jr]);
  const int wake_new = (g[1 + s] != u->have_wake[jr][s]) || u->have_rows[jr][s][0] != d[5] || u->have_rows[jr][s][1] != d[6];
  if (!wing_new && !wake_new) {
    u->skipped++;
    return 0;
  }
  for (int ib = 0; ib < nb; ++ib) {
    if (wing_new && (rc = vlc_rotor_put_wing(u->ctx, jr, ib, orc_rotor_wiP(r, ib)))) return rc;
    if (wake_new) {
      if (nNwake > 0 && (rc = vlc_rotor_put_nwake(u->ctx, jr, ib, predicted, orc_rotor_waN(r, ib, predicted)))) return rc;
      if (nFwake > 0 && (rc = vlc_rotor_put_fwake(u->ctx, jr, ib, predicted, orc_rotor_waF(r, ib, predicted)))) return rc;
      if (r->prescWakeNt > 0 && (rc = vlc_rotor_put_pfwake(u->ctx, jr, ib, predicted, orc_rotor_wapF(r, ib, predicted)))) return rc;
    }
  }
  u->have_wing[jr] = g[0];
  if (wake_new) {
    u->have_wake[jr][s] = g[1 + s];
    u->have_rows[jr][s][0] = d[5];
    u->have_rows[jr][s][1] = d[6];
  }
  u->uploads++;
  return 0;
}

static int h_vind_points(void *user, int jr, int what, int predicted, long m, const double *P, double *V) {
  gpu_user_t *u = (gpu_user_t *)user;
  int rc = sync_rotor(u, jr, predicted);
  if (rc) return u->last_rc = rc;
  switch (what) {
    case 0: rc = vlc_rotor_vind_bywing(u->ctx, jr, m, P, V); break;
    case 1: rc = vlc_rotor_vind_bywake(u->ctx, jr, predicted, m, P, V); break;
    case 2: rc = vlc_rotor_vind(u->ctx, jr, predicted, m, P, V); break;
    default: rc = vlc_rotor_vind_bywing_boundVortices(u->ctx, jr, m, P, V); break;
  }
  return u->last_rc = rc;
}

static int h_onN(void *user, int jr, const double *Nwake, int rows, int cols, int ld, int predicted, double *out) {
  gpu_user_t *u = (gpu_user_t *)user;
  int rc = sync_rotor(u, jr, predicted);
  if (rc) return u->last_rc = rc;
  return u->last_rc = vlc_vind_onNwake_byRotor(u->ctx, jr, Nwake, rows, cols, ld, predicted, out);
}

static int h_onF(void *user, int jr, const double *Fwake, int rows, int predicted, double *out) {
  gpu_user_t *u = (gpu_user_t *)user;
  int rc = sync_rotor(u, jr, predicted);
  if (rc) return u->last_rc = rc;
  return u->last_rc = vlc_vind_onFwake_byRotor(u->ctx, jr, Fwake, rows, predicted, out);
}

static int h_calcAIC(void *user, int ir, double *AIC, double *AIC_inv) {
  gpu_user_t *u = (gpu_user_t *)user;
  (void)AIC_inv; /* the explicit inverse is not needed: vlc_rotor_solve uses the LU factors */
  int rc = sync_rotor(u, ir, 0);
  if (rc) return u->last_rc = rc;
  return u->last_rc = vlc_rotor_calcAIC(u->ctx, ir, AIC);
}

static int h_solve(void *user, int ir, const double *RHS, double *gamVec) {
  gpu_user_t *u = (gpu_user_t *)user;
  return u->last_rc = vlc_rotor_solve(u->ctx, ir, RHS, gamVec);
}

/* ---- backends ---- */
static int g_assignshed(gpu_user_t *u, int ir, int edge) { return vlc_rotor_assignshed(u->ctx, ir, edge); }
static int g_age(gpu_user_t *u, int ir, double dt, double om) { return vlc_rotor_age_wake(u->ctx, ir, dt, om); }
static int g_dissipate(gpu_user_t *u, int ir, double dt, double nu) { return vlc_rotor_dissipate_wake(u->ctx, ir, dt, nu); }
static int g_strain(gpu_user_t *u, int ir) { return vlc_rotor_strain_wake(u->ctx, ir); }
static int g_to_pred(gpu_user_t *u, int ir) { return vlc_rotor_wake_to_predicted(u->ctx, ir); }
/* convectwake on the device, its last statement included: the prescribed far wake (classdef.f90:4826-4828), when the case
 * uses one, is regenerated from the device's own far rows (vlc_rotor_updatePrescribedWake) -- nothing crosses the bus */
static int g_convect(gpu_user_t *u, int ir, int iter, double dt, int p) {
  int rc = vlc_rotor_convectwake(u->ctx, ir, dt, p);
  orc_rotor_t *r = orc_case_rotor(u->cas, ir);
  if (rc || !(r->prescWakeNt > 0 && iter > r->prescWakeNt)) return rc;
  return vlc_rotor_updatePrescribedWake(u->ctx, ir, r->omegaSlow * dt, r->prescWakeGenNt, p);
}
static int g_rollup(gpu_user_t *u, int ir) { return vlc_rotor_rollup(u->ctx, ir); }
static int g_sweep(gpu_user_t *u, int p, int addInit) { return vlc_wake_sweep(u->ctx, p, addInit); }
static int g_sweep_count(gpu_user_t *u, int64_t *M) { return vlc_wake_sweep_count(u->ctx, M); }
static int g_sweep_slice(gpu_user_t *u, int p, int64_t first, int64_t count, double *x) {
  return vlc_wake_sweep_slice(u->ctx, p, first, count, x);
}
static int g_sweep_scatter(gpu_user_t *u, int p, int addInit, const double *x) { return vlc_wake_sweep_scatter(u->ctx, p, addInit, x); }
static int g_stream_sync(gpu_user_t *u) { return vlc_sync(u->ctx); }
static int g_velop(gpu_user_t *u, int ir, int op) { return vlc_rotor_wakevel_op(u->ctx, ir, op); }
static int g_velcopy(gpu_user_t *u, int ir, int dst, int src) { return vlc_rotor_wakevel_copy(u->ctx, ir, dst, src); }
static int g_vellin(gpu_user_t *u, int ir, int dst, int n, const int *src, const double *coef, double div) {
  return vlc_rotor_wakevel_lincomb(u->ctx, ir, dst, n, src, coef, div);
}
static int g_burst(gpu_user_t *u, int ir) {
  const orc_rotor_t *r = orc_case_rotor(u->cas, ir);
  return vlc_rotor_burst_wake(u->ctx, ir, r->skewLimit, r->chord);
}
static const resident_ops_t gpu_ops = {resident_begin, sync_wing, g_assignshed, g_age, g_dissipate, g_strain,
                                       g_to_pred, g_convect, g_rollup, g_sweep, g_velop,
                                       g_sweep_count, g_sweep_slice, g_sweep_scatter, g_stream_sync, g_velcopy, g_vellin,
                                       g_burst};

#define ROT(u, ir) orc_case_rotor((u)->cas, (ir))
static int c_begin(gpu_user_t *u) { u->resident_started = 1; return 0; }
static int c_sync(gpu_user_t *u, int ir) { (void)u; (void)ir; return 0; }
static int c_assignshed(gpu_user_t *u, int ir, int edge) {
  if (ROT(u, ir)->nNwake > 0) orc_rotor_assignshed(ROT(u, ir), edge ? "TE" : "LE");
  return 0;
}
static int c_age(gpu_user_t *u, int ir, double dt, double om) {
  (void)om; /* orc_rotor_age_wake reads rotor%omegaSlow itself */
  if (ROT(u, ir)->nNwake > 0) orc_rotor_age_wake(ROT(u, ir), dt);
  return 0;
}
static int c_dissipate(gpu_user_t *u, int ir, double dt, double nu) {
  if (ROT(u, ir)->nNwake > 0) orc_rotor_dissipate_wake(ROT(u, ir), dt, nu);
  return 0;
}
static int c_strain(gpu_user_t *u, int ir) {
  if (ROT(u, ir)->nNwake > 0) orc_rotor_strain_wake(ROT(u, ir));
  return 0;
}
static int c_to_pred(gpu_user_t *u, int ir) {
  if (ROT(u, ir)->nNwake > 0) orc_rotor_wake_to_predicted(ROT(u, ir));
  return 0;
}
static int c_convect(gpu_user_t *u, int ir, int iter, double dt, int p) {
  if (ROT(u, ir)->nNwake > 0) orc_rotor_convectwake(ROT(u, ir), iter, dt, p ? 'P' : 'C');
  return 0;
}
static int c_rollup(gpu_user_t *u, int ir) { orc_rotor_rollup(ROT(u, ir)); return 0; }
static int c_sweep(gpu_user_t *u, int p, int addInit) { (void)addInit; return orc_case_wake_sweep(u->cas, p); }
static int c_velop(gpu_user_t *u, int ir, int op) { return ROT(u, ir)->nNwake > 0 ? orc_rotor_wakevel_op(ROT(u, ir), op) : 0; }
/* CPU emulation of vlc_wake_sweep_count / _slice / _scatter: the target list in the library's order (per rotor, per
 * convected blade: wake nodes of columns 1..ns+1 with the rows rowNear..nNwake fastest, then the far rows rowFar..nFwake).
 * `visit` walks it and hands out the address of each target's velocity in velNwake / velFwake [Predicted]. */
static int64_t c_walk(gpu_user_t *u, int p, void (*visit)(double *vel, int64_t q, void *arg), void *arg) {
  int64_t q = 0;
  for (int ir = 0; ir < u->nr; ++ir) {
    orc_rotor_t *r = ROT(u, ir);
    if (r->nNwake <= 0) continue;
    const int nact = r->nNwake - r->rowNear + 1 > 0 ? r->nNwake - r->rowNear + 1 : 0;
    const int nfar = r->nFwake - r->rowFar + 1 > 0 ? r->nFwake - r->rowFar + 1 : 0;
    for (int ib = 0; ib < r->nbConvect; ++ib) {
      orc_blade_t *b = &r->blade[ib];
      double *vn = p ? b->velNwakePredicted : b->velNwake, *vf = p ? b->velFwakePredicted : b->velFwake;
      for (int j = 1; j <= r->ns + 1; ++j)
        for (int i = r->rowNear; i < r->rowNear + nact; ++i, ++q)
          if (visit) visit(vn + 3 * ((size_t)(i - 1) + (size_t)r->nNwake * (j - 1)), q, arg);
      for (int i = r->rowFar; i < r->rowFar + nfar; ++i, ++q)
        if (visit) visit(vf + 3 * (size_t)(i - 1), q, arg);
    }
  }
  return q;
}
typedef struct { double *x; int64_t first, count; } c_slice_t;
static void c_take(double *vel, int64_t q, void *arg) { /* own slice into the exchange buffer, then poison the array entry */
  c_slice_t *s = (c_slice_t *)arg;
  if (q >= s->first && q < s->first + s->count) memcpy(s->x + 3 * q, vel, 3 * sizeof(double));
  vel[0] = vel[1] = vel[2] = 0.0 / 0.0 * (double)(q + 1); /* NaN: only the scatter below may make it a number again */
}
static void c_put(double *vel, int64_t q, void *arg) { memcpy(vel, ((c_slice_t *)arg)->x + 3 * q, 3 * sizeof(double)); }
static int c_sweep_count(gpu_user_t *u, int64_t *M) { *M = c_walk(u, 0, NULL, NULL); return 0; }
static int c_sweep_slice(gpu_user_t *u, int p, int64_t first, int64_t count, double *x) {
  int rc = orc_case_wake_sweep(u->cas, p); /* the whole sweep by the oracle (incl. the initWakeVel terms) ... */
  if (rc) return rc;
  c_slice_t s = {x, first, count};
  c_walk(u, p, c_take, &s); /* ... of which this rank keeps only its slice */
  return 0;
}
static int c_sweep_scatter(gpu_user_t *u, int p, int addInit, const double *x) {
  (void)addInit; /* already inside the oracle's sweep */
  c_slice_t s = {(double *)x, 0, 0};
  c_walk(u, p, c_put, &s);
  return 0;
}
static int c_stream_sync(gpu_user_t *u) { (void)u; return 0; }
static int c_velcopy(gpu_user_t *u, int ir, int dst, int src) {
  return ROT(u, ir)->nNwake > 0 ? orc_rotor_wakevel_copy(ROT(u, ir), dst, src) : 0;
}
static int c_vellin(gpu_user_t *u, int ir, int dst, int n, const int *src, const double *coef, double div) {
  return ROT(u, ir)->nNwake > 0 ? orc_rotor_wakevel_lincomb(ROT(u, ir), dst, n, src, coef, div) : 0;
}
static int c_burst(gpu_user_t *u, int ir) {
  if (ROT(u, ir)->nNwake > 0) orc_rotor_burst_wake(ROT(u, ir));
  return 0;
}
static const resident_ops_t cpu_ops = {c_begin, c_sync, c_assignshed, c_age, c_dissipate, c_strain,
                                       c_to_pred, c_convect, c_rollup, c_sweep, c_velop,
                                       c_sweep_count, c_sweep_slice, c_sweep_scatter, c_stream_sync, c_velcopy, c_vellin,
                                       c_burst};

/* One wake sweep of the staged orchestration.  One process: the backend's whole sweep.  One process per GPU (world > 1):
 * every rank holds the whole wake; it sweeps its slice of the targets, the slices are all-gathered (the one exchange of
 * this stage), every rank scatters the complete list. */
static int staged_sweep(gpu_user_t *u, int p, int addInit) {
  const resident_ops_t *o = u->ops;
  if (u->world <= 1) return o->wake_sweep(u, p, addInit);
  int64_t M = 0;
  int rc = o->sweep_count(u, &M);
  if (rc || M <= 0) return rc;
  const long per = (long)((M + u->world - 1) / u->world);
  if (per * u->world > u->xbuf_targets) return VLC_ERR_ARG;
  const int64_t first = (int64_t)u->rank * per, count = first >= M ? 0 : (first + per > M ? M - first : per);
  if ((rc = o->sweep_slice(u, p, first < M ? first : 0, count, u->xbuf))) return rc;
  if ((rc = o->stream_sync(u))) return rc;
  if ((rc = u->exchange(u->exchange_arg, per))) return rc;
  u->exchanges++;
  return o->sweep_scatter(u, p, addInit, u->xbuf);
}

/* main.f90:466-506 as separate stages: assignshed('LE'), age_wake, dissipate_wake of every rotor, in the driver's order */
static int h_wake_prestep(void *user, int iter) {
  gpu_user_t *u = (gpu_user_t *)user;
  const resident_ops_t *o = u->ops;
  const orc_config_t *cfg = orc_case_config(u->cas);
  if (!u->resident_started) CK(o->begin(u));
  for (int ir = 0; ir < u->nr; ++ir) CK(o->sync(u, ir));
  for (int ir = 0; ir < u->nr; ++ir) CK(o->assignshed(u, ir, 0));
  for (int ir = 0; ir < u->nr; ++ir) CK(o->age_wake(u, ir, cfg->dt, orc_case_rotor(u->cas, ir)->omegaSlow));
  if (cfg->wakeDissipation == 1)
    for (int ir = 0; ir < u->nr; ++ir) CK(o->dissipate_wake(u, ir, cfg->dt, cfg->kinematicVisc));
  if (cfg->wakeBurst != 0 && iter % cfg->wakeBurst == 0) /* main.f90:490-497 */
    for (int ir = 0; ir < u->nr; ++ir)
      if (orc_case_rotor(u->cas, ir)->nNwake > 0) CK(o->burst_wake(u, ir));
  return 0;
}

/* main.f90:800-1440 as separate stages: the wake sweeps, the fdScheme switch (0: explicit Euler, 1: predictor-corrector,
 * 3: Adams-Bashforth / Adams-Moulton), strain_wake, rollup, assignshed('TE') */
static int h_wake_convect(void *user, int iter) {
  gpu_user_t *u = (gpu_user_t *)user;
  const resident_ops_t *o = u->ops;
  const orc_config_t *cfg = orc_case_config(u->cas);
  const double dt = cfg->dt;
  const int addInit = iter < cfg->initWakeVelNt;
  const int nr = u->nr;
  for (int ir = 0; ir < nr; ++ir) CK(o->sync(u, ir)); /* the solve changed the wing's circulation */
  CK(staged_sweep(u, 0, addInit));
  switch (cfg->fdScheme) {
    case 0: /* :846-859 */
      for (int ir = 0; ir < nr; ++ir) CK(o->convectwake(u, ir, iter, dt, 0));
      break;
    case 1: /* :861-949 */
      for (int ir = 0; ir < nr; ++ir) {
        CK(o->wake_to_predicted(u, ir));
        CK(o->convectwake(u, ir, iter, dt, 1));
      }
      CK(staged_sweep(u, 1, addInit));
      for (int ir = 0; ir < nr; ++ir) {
        CK(o->wakevel_op(u, ir, VLC_VEL_ORDER2));
        CK(o->convectwake(u, ir, iter, dt, 0));
      }
      break;
    case 2: /* :951-1000 explicit Adams-Bashforth: velStep = vel1 = vel = 0.5*(3*vel - vel1), then convect */
      for (int ir = 0; ir < nr; ++ir) {
        if (iter == 1) {
          CK(o->convectwake(u, ir, iter, dt, 0));
          CK(o->wakevel_op(u, ir, VLC_VEL_FIRST_STEP));
        } else {
          CK(o->wakevel_op(u, ir, VLC_VEL_AB2));
          CK(o->wakevel_op(u, ir, VLC_VEL_FIRST_STEP));
          CK(o->wakevel_op(u, ir, VLC_VEL_COPY_TO_STEP));
          CK(o->convectwake(u, ir, iter, dt, 0));
        }
      }
      break;
    case 3: /* :1002-1115 */
      if (iter == 1) {
        for (int ir = 0; ir < nr; ++ir) {
          CK(o->convectwake(u, ir, iter, dt, 0));
          CK(o->wakevel_op(u, ir, VLC_VEL_FIRST_STEP));
        }
      } else {
        for (int ir = 0; ir < nr; ++ir) {
          CK(o->wake_to_predicted(u, ir));
          CK(o->wakevel_op(u, ir, VLC_VEL_AB2));
          CK(o->convectwake(u, ir, iter, dt, 1));
        }
        CK(staged_sweep(u, 1, addInit));
        for (int ir = 0; ir < nr; ++ir) {
          CK(o->wakevel_op(u, ir, VLC_VEL_AM2));
          CK(o->convectwake(u, ir, iter, dt, 0));
          CK(o->wakevel_op(u, ir, VLC_VEL_SHIFT_HISTORY));
        }
      }
      break;
    case 4:   /* :1117-1248 Adams-Bashforth / Adams-Moulton, third order (its `iter == 0` start branch never runs) */
    case 5: { /* :1250-1404 fourth order: steps 1, 2, 3 fill vel1, vel2, vel3 */
      const int order = cfg->fdScheme == 4 ? 3 : 4;
      const int start = (order == 3) ? (iter == 2 ? 2 : 0) : (iter <= 3 ? iter : 0);
      static const int hist[4] = {0, VLC_VEL_ARRAY_1, VLC_VEL_ARRAY_2, VLC_VEL_ARRAY_3};
      static const int p3[3] = {VLC_VEL_ARRAY, VLC_VEL_ARRAY_2, VLC_VEL_ARRAY_1};
      static const double pc3[3] = {23.0, -16.0, 5.0};
      static const int c3[3] = {VLC_VEL_ARRAY_PREDICTED, VLC_VEL_ARRAY_STEP, VLC_VEL_ARRAY_2};
      static const double cc3[3] = {5.0, 8.0, -1.0};
      static const int p4[4] = {VLC_VEL_ARRAY, VLC_VEL_ARRAY_3, VLC_VEL_ARRAY_2, VLC_VEL_ARRAY_1};
      static const double pc4[4] = {55.0, -59.0, 37.0, -9.0};
      static const int c4[4] = {VLC_VEL_ARRAY_PREDICTED, VLC_VEL_ARRAY_STEP, VLC_VEL_ARRAY_3, VLC_VEL_ARRAY_2};
      static const double cc4[4] = {9.0, 19.0, -5.0, 1.0};
      if (start) {
        for (int ir = 0; ir < nr; ++ir) {
          CK(o->convectwake(u, ir, iter, dt, 0));
          CK(o->wakevel_copy(u, ir, hist[start], VLC_VEL_ARRAY));
        }
      } else {
        for (int ir = 0; ir < nr; ++ir) {
          CK(o->wake_to_predicted(u, ir));
          CK(o->wakevel_copy(u, ir, VLC_VEL_ARRAY_STEP, VLC_VEL_ARRAY));
          if (order == 3) CK(o->wakevel_lincomb(u, ir, VLC_VEL_ARRAY, 3, p3, pc3, 12.0));
          else CK(o->wakevel_lincomb(u, ir, VLC_VEL_ARRAY, 4, p4, pc4, 24.0));
          CK(o->convectwake(u, ir, iter, dt, 1));
        }
        CK(staged_sweep(u, 1, addInit));
        for (int ir = 0; ir < nr; ++ir) {
          if (order == 3) CK(o->wakevel_lincomb(u, ir, VLC_VEL_ARRAY, 3, c3, cc3, 12.0));
          else CK(o->wakevel_lincomb(u, ir, VLC_VEL_ARRAY, 4, c4, cc4, 24.0));
          CK(o->convectwake(u, ir, iter, dt, 0));
          CK(o->wakevel_copy(u, ir, VLC_VEL_ARRAY_1, VLC_VEL_ARRAY_2));
          if (order == 3) {
            CK(o->wakevel_copy(u, ir, VLC_VEL_ARRAY_2, VLC_VEL_ARRAY_STEP));
          } else {
            CK(o->wakevel_copy(u, ir, VLC_VEL_ARRAY_2, VLC_VEL_ARRAY_3));
            CK(o->wakevel_copy(u, ir, VLC_VEL_ARRAY_3, VLC_VEL_ARRAY_STEP));
          }
        }
      }
    } break;
    default: return u->last_rc = VLC_ERR_ARG;
  }
  if (cfg->wakeStrain == 1) /* :1409-1416 */
    for (int ir = 0; ir < nr; ++ir) CK(o->strain_wake(u, ir));
  for (int ir = 0; ir < nr; ++ir) { /* :1419-1439 */
    const orc_rotor_t *r = orc_case_rotor(u->cas, ir);
    if (r->nNwake <= 0) continue;
    if (r->rowNear == 1) CK(o->rollup(u, ir));
    CK(o->assignshed(u, ir, 1));
  }
  return 0;
}

/* ------------------------------------------------------------------ tier 2c: the collocation-point stage
 * = gpu_cp_rhs_solve / gpu_cp_forces of fortran/libGPU.f90.  The stage's calls, as the orchestration below makes them;
 * two backends like the wake stages: `gpu_cp_ops` forwards to the C ABI, `cpu_cp_ops` emulates the library on the CPU
 * (its own copies of the wing records, the g++ build of cp_stage.cuh for the record arithmetic, the oracle for the
 * sweeps), so that the orchestration and its write-backs run without a GPU and must reproduce the driver's inline
 * statement bit for bit (tests/test_staged_hooks.py). */
typedef struct cp_ops {
  int (*sync)(gpu_user_t *u, int ir); /* the library's copy of rotor ir becomes current (wing; wake unless resident) */
  int (*calc_RHS)(gpu_user_t *u, int ir, double *velCP, double *RHS);
  int (*solve_map_gam)(gpu_user_t *u, int ir, double *gamVec);
  int (*put_sections)(gpu_user_t *u, int ir, int ib, const double *sec);
  int (*calc_velCPTotal)(gpu_user_t *u, int ir);
  int (*calc_force)(gpu_user_t *u, int ir, double density, double dt, double Omega, int spanwiseLiftSwitch);
  int (*get_loads)(gpu_user_t *u, int ir, int ib, double *loads);
  int (*get_wing)(gpu_user_t *u, int ir, int ib, double *wiP);
} cp_ops_t;

static int gc_sync(gpu_user_t *u, int ir) { return sync_rotor(u, ir, 0); }
static int gc_rhs(gpu_user_t *u, int ir, double *v, double *rhs) { return vlc_rotor_calc_RHS(u->ctx, ir, v, rhs); }
static int gc_solve(gpu_user_t *u, int ir, double *g) { return vlc_rotor_solve_map_gam(u->ctx, ir, g); }
static int gc_put_sec(gpu_user_t *u, int ir, int ib, const double *s) { return vlc_rotor_put_sections(u->ctx, ir, ib, s); }
static int gc_veltot(gpu_user_t *u, int ir) { return vlc_rotor_calc_velCPTotal(u->ctx, ir); }
static int gc_force(gpu_user_t *u, int ir, double rho, double dt, double Om, int sl) {
  return vlc_rotor_calc_force(u->ctx, ir, rho, dt, Om, sl);
}
static int gc_get_loads(gpu_user_t *u, int ir, int ib, double *l) { return vlc_rotor_get_loads(u->ctx, ir, ib, l); }
static int gc_get_wing(gpu_user_t *u, int ir, int ib, double *w) { return vlc_rotor_get_wing(u->ctx, ir, ib, w); }
static const cp_ops_t gpu_cp_ops = {gc_sync, gc_rhs, gc_solve, gc_put_sec, gc_veltot, gc_force, gc_get_loads, gc_get_wing};

/* CPU emulation of the library's tier 2c state (tests/native/cp_stage_host.cpp = cp_stage.cuh built by g++) */
void cp_host_loads(int nbConvect, int nc, int ns, double density, double dt, double Omega, int spanwiseLiftSwitch, double *wiP,
                   const double *sec, double *loads);
void cp_host_rhs(int N, int npb, int nbConvect, int axisym, const double *wiP, double *RHS);
void cp_host_map_gam(int nb, int npb, int nbConvect, int axisym, const double *gamVec, double *wiP);
typedef struct cp_emu {
  double *wiP, *sec, *loads, *rhs; /* what vlc_ctx holds per rotor: records of all blades, section frames, loads, RHS */
} cp_emu_t;
enum { WP_CP = 64, WP_VELCP = 76, WP_VELCPTOTAL = 79, WP_GAMPREV = 50, WP_NFORCE = 85, WP_DELP = 95 };
#define SEC_N(ns) (10 * (ns) + 6)
#define LOADS_N(ns) (12 + 25 * (ns))

static int ec_sync(gpu_user_t *u, int ir) { /* = vlc_rotor_put_wing of every blade */
  orc_rotor_t *r = ROT(u, ir);
  const size_t per = (size_t)r->nc * r->ns * VLC_WINGPANEL_DOUBLES;
  for (int ib = 0; ib < r->nb; ++ib) memcpy(u->emu[ir].wiP + per * ib, orc_rotor_wiP(r, ib), per * sizeof(double));
  return 0;
}
/* one source swept at the collocation points of the emulated records and added to / subtracted from a record field */
static void ec_sweep_acc(gpu_user_t *u, int ir, int jr, int what, int field, double sign) {
  orc_rotor_t *r = ROT(u, ir);
  const long m = (long)r->nbConvect * r->nc * r->ns;
  double *w = u->emu[ir].wiP, *P = (double *)malloc(sizeof(double) * 3 * (size_t)m), *V = (double *)malloc(sizeof(double) * 3 * (size_t)m);
  for (long q = 0; q < m; ++q) memcpy(P + 3 * q, w + VLC_WINGPANEL_DOUBLES * q + WP_CP, 3 * sizeof(double));
  orc_rotor_vind_points(ROT(u, jr), what, 0, m, P, V);
  for (long q = 0; q < m; ++q)
    for (int k = 0; k < 3; ++k) {
      double *x = w + VLC_WINGPANEL_DOUBLES * q + field + k;
      *x = sign > 0 ? *x + V[3 * q + k] : *x - V[3 * q + k];
    }
  free(P);
  free(V);
}
static int ec_rhs(gpu_user_t *u, int ir, double *velCP, double *RHS) {
  orc_rotor_t *r = ROT(u, ir);
  const int npb = r->nc * r->ns, N = npb * r->nb;
  for (int jr = 0; jr < u->nr; ++jr) {
    ec_sweep_acc(u, ir, jr, 1, WP_VELCP, +1.0);
    if (jr != ir) ec_sweep_acc(u, ir, jr, 0, WP_VELCP, +1.0);
  }
  cp_host_rhs(N, npb, r->nbConvect, r->axisymmetrySwitch, u->emu[ir].wiP, u->emu[ir].rhs);
  if (velCP)
    for (long q = 0; q < (long)r->nbConvect * npb; ++q)
      memcpy(velCP + 3 * q, u->emu[ir].wiP + VLC_WINGPANEL_DOUBLES * q + WP_VELCP, 3 * sizeof(double));
  if (RHS) memcpy(RHS, u->emu[ir].rhs, sizeof(double) * (size_t)N);
  return 0;
}
static int ec_solve(gpu_user_t *u, int ir, double *gamVec) {
  orc_rotor_t *r = ROT(u, ir);
  const int npb = r->nc * r->ns, N = npb * r->nb;
  double *g = (double *)malloc(sizeof(double) * (size_t)N);
  orc_matmulAX(N, N, r->AIC_inv, u->emu[ir].rhs, g);
  cp_host_map_gam(r->nb, npb, r->nbConvect, r->axisymmetrySwitch, g, u->emu[ir].wiP);
  if (gamVec) memcpy(gamVec, g, sizeof(double) * (size_t)N);
  free(g);
  return 0;
}
static int ec_put_sec(gpu_user_t *u, int ir, int ib, const double *s) {
  memcpy(u->emu[ir].sec + (size_t)SEC_N(ROT(u, ir)->ns) * ib, s, sizeof(double) * (size_t)SEC_N(ROT(u, ir)->ns));
  return 0;
}
static void ec_axisym_field(orc_rotor_t *r, double *w, int field, int n) { /* = cp_axisym_field_kernel */
  const int npb = r->nc * r->ns;
  for (int ib = 1; ib < r->nb; ++ib)
    for (int q = 0; q < npb; ++q)
      memcpy(w + VLC_WINGPANEL_DOUBLES * ((size_t)q + (size_t)npb * ib) + field, w + VLC_WINGPANEL_DOUBLES * (size_t)q + field,
             sizeof(double) * (size_t)n);
}
static int ec_veltot(gpu_user_t *u, int ir) {
  orc_rotor_t *r = ROT(u, ir);
  double *w = u->emu[ir].wiP;
  for (long q = 0; q < (long)r->nbConvect * r->nc * r->ns; ++q)
    memcpy(w + VLC_WINGPANEL_DOUBLES * q + WP_VELCPTOTAL, w + VLC_WINGPANEL_DOUBLES * q + WP_VELCP, 3 * sizeof(double));
  for (int jr = 0; jr < u->nr; ++jr) ec_sweep_acc(u, ir, jr, 3, WP_VELCPTOTAL, -1.0);
  ec_sweep_acc(u, ir, ir, 0, WP_VELCPTOTAL, +1.0);
  if (r->axisymmetrySwitch == 1) ec_axisym_field(r, w, WP_VELCPTOTAL, 3);
  return 0;
}
static int ec_force(gpu_user_t *u, int ir, double rho, double dt, double Om, int sl) {
  orc_rotor_t *r = ROT(u, ir);
  double *w = u->emu[ir].wiP, *l = u->emu[ir].loads;
  const int ns = r->ns, nld = LOADS_N(ns);
  cp_host_loads(r->nbConvect, r->nc, ns, rho, dt, Om, sl, w, u->emu[ir].sec, l);
  if (r->axisymmetrySwitch == 1) {
    ec_axisym_field(r, w, WP_DELP, 2);
    ec_axisym_field(r, w, WP_GAMPREV, 1);
    ec_axisym_field(r, w, WP_NFORCE, 6);
    for (int ib = 1; ib < r->nb; ++ib) /* = cp_axisym_loads_kernel: everything but secChordwiseResVel */
      for (int k = 0; k < nld; ++k)
        if (k < 12 || k >= 12 + 3 * ns) l[(size_t)nld * ib + k] = l[k];
  }
  return 0;
}
static int ec_get_loads(gpu_user_t *u, int ir, int ib, double *l) {
  memcpy(l, u->emu[ir].loads + (size_t)LOADS_N(ROT(u, ir)->ns) * ib, sizeof(double) * (size_t)LOADS_N(ROT(u, ir)->ns));
  return 0;
}
static int ec_get_wing(gpu_user_t *u, int ir, int ib, double *wq) {
  const size_t per = (size_t)ROT(u, ir)->nc * ROT(u, ir)->ns * VLC_WINGPANEL_DOUBLES;
  memcpy(wq, u->emu[ir].wiP + per * ib, per * sizeof(double));
  return 0;
}
static const cp_ops_t cpu_cp_ops = {ec_sync, ec_rhs, ec_solve, ec_put_sec, ec_veltot, ec_force, ec_get_loads, ec_get_wing};

/* main.f90:548-603 of every rotor through the stage's calls.  The driver has written the kinematic velCP into its
 * records and saved gamVecPrev; it maps gamVec into its own records afterwards. */
static int h_cp_rhs_solve(void *user) {
  gpu_user_t *u = (gpu_user_t *)user;
  const cp_ops_t *o = u->cp;
  for (int ir = 0; ir < u->nr; ++ir) CK(o->sync(u, ir));
  for (int ir = 0; ir < u->nr; ++ir) {
    orc_rotor_t *r = ROT(u, ir);
    const int npb = r->nc * r->ns;
    double *v = (double *)malloc(sizeof(double) * 3 * (size_t)r->nbConvect * npb);
    int rc = o->calc_RHS(u, ir, v, r->RHS);
    if (!rc)
      for (int ib = 0; ib < r->nbConvect; ++ib) /* velCP back into the driver's records: blade_calc_force reads it */
        for (int q = 0; q < npb; ++q)
          memcpy(orc_rotor_wiP(r, ib) + (size_t)VLC_WINGPANEL_DOUBLES * q + WP_VELCP, v + 3 * ((size_t)q + (size_t)npb * ib),
                 3 * sizeof(double));
    free(v);
    CK(rc);
  }
  for (int ir = 0; ir < u->nr; ++ir) CK(o->solve_map_gam(u, ir, ROT(u, ir)->gamVec));
  u->cp_rhs_calls++;
  return 0;
}

/* main.f90:630-663 + calc_secAlpha + calc_force of rotor ir down to the blade sums */
static int h_cp_forces(void *user, int ir) {
  gpu_user_t *u = (gpu_user_t *)user;
  const cp_ops_t *o = u->cp;
  const orc_config_t *cfg = orc_case_config(u->cas);
  orc_rotor_t *r = ROT(u, ir);
  const int ns = r->ns;
  for (int jr = 0; jr < u->nr; ++jr) CK(o->sync(u, jr)); /* the bound vortices of every rotor are sources */
  double *sec = (double *)malloc(sizeof(double) * (size_t)SEC_N(ns)), *l = (double *)malloc(sizeof(double) * (size_t)LOADS_N(ns));
  int rc = 0;
  for (int ib = 0; ib < r->nb && !rc; ++ib) {
    const orc_blade_t *b = &r->blade[ib];
    memcpy(sec, b->secTauCapChord, sizeof(double) * 3 * (size_t)ns);
    memcpy(sec + 3 * ns, b->secNormalVec, sizeof(double) * 3 * (size_t)ns);
    memcpy(sec + 6 * ns, b->secCP, sizeof(double) * 3 * (size_t)ns);
    memcpy(sec + 9 * ns, b->secArea, sizeof(double) * (size_t)ns);
    memcpy(sec + 10 * ns, b->yAxisAziFlap, 3 * sizeof(double));
    memcpy(sec + 10 * ns + 3, b->zAxisAziFlap, 3 * sizeof(double));
    rc = o->put_sections(u, ir, ib, sec);
  }
  if (!rc) rc = o->calc_velCPTotal(u, ir);
  if (!rc) rc = o->calc_force(u, ir, cfg->density, cfg->dt, r->Omega, r->spanwiseLiftSwitch);
  for (int ib = 0; ib < r->nb && !rc; ++ib) {
    orc_blade_t *b = &r->blade[ib];
    if (ib >= r->nbConvect && r->axisymmetrySwitch != 1) continue;
    if ((rc = o->get_wing(u, ir, ib, orc_rotor_wiP(r, ib)))) break;
    if ((rc = o->get_loads(u, ir, ib, l))) break;
    double *s3[7] = {ib < r->nbConvect ? b->secChordwiseResVel : NULL, b->secDragDir, b->secLiftDir, b->secForceInertial,
                     b->secLift, b->secDrag, b->secLiftUnsteady};
    double *s1[4] = {b->secAlpha, b->secCL, b->secCD, b->secCLu};
    memcpy(b->forceInertial, l, 3 * sizeof(double));
    memcpy(b->lift, l + 3, 3 * sizeof(double));
    memcpy(b->drag, l + 6, 3 * sizeof(double));
    memcpy(b->liftUnsteady, l + 9, 3 * sizeof(double));
    for (int k = 0; k < 7; ++k)
      if (s3[k]) memcpy(s3[k], l + 12 + 3 * ns * k, sizeof(double) * 3 * (size_t)ns);
    for (int k = 0; k < 4; ++k) memcpy(s1[k], l + 12 + 21 * ns + ns * k, sizeof(double) * (size_t)ns);
  }
  free(sec);
  free(l);
  u->cp_force_calls++;
  return u->last_rc = rc;
}

/* Declares every rotor to the library (gpu_init of the Fortran shim) and installs the hook table.
 * The case must have its rotors initialised (orc_case_init_rotors).  Returns an opaque handle (free with
 * case_gpu_hooks_free) or NULL. */
void *case_gpu_hooks_install(orc_case_t *cas, vlc_ctx *ctx, int nr) {
  if (nr > MAX_ROTORS) return NULL;
  gpu_user_t *u = (gpu_user_t *)calloc(1, sizeof(gpu_user_t));
  u->cas = cas;
  u->ctx = ctx;
  u->nr = nr;
  vlc_rotors_clear(ctx); /* the context may have served another case: its rotors would be sources of this one's sweeps */
  for (int ir = 0; ir < nr; ++ir) {
    int d[10];
    orc_rotor_dims(orc_case_rotor(cas, ir), d);
    if (vlc_rotor_define(ctx, ir, d[0], d[1], d[2], d[3], d[4], 1)) {
      free(u);
      return NULL;
    }
  }
  orc_hooks_t h;
  memset(&h, 0, sizeof h);
  h.user = u;
  h.vind_points = h_vind_points;
  h.vind_onNwake = h_onN;
  h.vind_onFwake = h_onF;
  h.calcAIC = h_calcAIC;
  h.solve = h_solve;
  orc_case_set_hooks(cas, &h);
  return u;
}

/* Same, in resident mode: the two wake stages of the time loop run on the device too (tier 2b) and the wake records
 * never travel after the first step. */
void *case_gpu_hooks_install_resident(orc_case_t *cas, vlc_ctx *ctx, int nr) {
  gpu_user_t *u = (gpu_user_t *)case_gpu_hooks_install(cas, ctx, nr);
  if (!u) return NULL;
  u->resident = 1;
  u->ops = &gpu_ops;
  orc_hooks_t h;
  memset(&h, 0, sizeof h);
  h.user = u;
  h.vind_points = h_vind_points;
  h.vind_onNwake = h_onN;
  h.vind_onFwake = h_onF;
  h.calcAIC = h_calcAIC;
  h.solve = h_solve;
  h.wake_prestep = h_wake_prestep;
  h.wake_convect = h_wake_convect;
  orc_case_set_hooks(cas, &h);
  return u;
}

/* The same staged orchestration with the CPU backend: nothing touches the library; only the two wake-stage hooks are
 * installed (the five hot-path call sites keep the CPU oracle).  `ctx` is not used. */
void *case_cpu_staged_hooks_install(orc_case_t *cas, int nr) {
  if (nr > MAX_ROTORS) return NULL;
  gpu_user_t *u = (gpu_user_t *)calloc(1, sizeof(gpu_user_t));
  u->cas = cas;
  u->nr = nr;
  u->ops = &cpu_ops;
  orc_case_set_hooks(cas, NULL);
  orc_case_set_stage_hooks(cas, u, h_wake_prestep, h_wake_convect);
  return u;
}

/* Adds the collocation-point stage (tier 2c) to an installed table: the RHS / solve / map_gam of the time loop and the
 * force evaluation go through the stage's calls -- to the C ABI when the handle drives a library context, to the CPU
 * emulation otherwise. */
int case_hooks_enable_cp(void *handle) {
  gpu_user_t *u = (gpu_user_t *)handle;
  if (u->ctx) {
    u->cp = &gpu_cp_ops;
    for (int ir = 0; ir < u->nr; ++ir) { /* nbConvect and the axisymmetry switch (resident mode sets them again, identically) */
      const orc_rotor_t *r = ROT(u, ir);
      int rc = vlc_rotor_set_wake_params(u->ctx, ir, r->nbConvect, r->axisymmetrySwitch, r->ductSwitch, r->suppressFwakeSwitch,
                                         r->rollupStart, r->rollupEnd, r->Omega * r->controlPitch[0], r->apparentViscCoeff,
                                         r->decayCoeff, r->initWakeVel);
      if (rc) return u->last_rc = rc;
    }
  } else {
    u->cp = &cpu_cp_ops;
    u->emu = (cp_emu_t *)calloc((size_t)u->nr, sizeof(cp_emu_t));
    for (int ir = 0; ir < u->nr; ++ir) {
      const orc_rotor_t *r = ROT(u, ir);
      const size_t npb = (size_t)r->nc * r->ns;
      u->emu[ir].wiP = (double *)calloc(npb * r->nb * VLC_WINGPANEL_DOUBLES, sizeof(double));
      u->emu[ir].sec = (double *)calloc((size_t)SEC_N(r->ns) * r->nb, sizeof(double));
      u->emu[ir].loads = (double *)calloc((size_t)LOADS_N(r->ns) * r->nb, sizeof(double));
      u->emu[ir].rhs = (double *)calloc(npb * r->nb, sizeof(double));
    }
  }
  orc_case_set_cp_hooks(u->cas, u, h_cp_rhs_solve, h_cp_forces);
  return 0;
}
long case_hooks_cp_rhs_calls(void *handle) { return ((gpu_user_t *)handle)->cp_rhs_calls; }
long case_hooks_cp_force_calls(void *handle) { return ((gpu_user_t *)handle)->cp_force_calls; }

/* resident mode: bring the device's wake (records and velocity arrays) back into the driver's arrays */
int case_gpu_hooks_download_wake(void *handle) {
  gpu_user_t *u = (gpu_user_t *)handle;
  for (int jr = 0; jr < u->nr; ++jr) {
    orc_rotor_t *r = orc_case_rotor(u->cas, jr);
    for (int ib = 0; ib < r->nb; ++ib) {
      for (int s = 0; s < 2; ++s) {
        if (r->nNwake > 0) CK(vlc_rotor_get_nwake(u->ctx, jr, ib, s, orc_rotor_waN(r, ib, s)));
        if (r->nFwake > 0) CK(vlc_rotor_get_fwake(u->ctx, jr, ib, s, orc_rotor_waF(r, ib, s)));
        if (r->prescWakeNt > 0) CK(vlc_rotor_get_pfwake(u->ctx, jr, ib, s, orc_rotor_wapF(r, ib, s), NULL));
      }
      for (int w = 0; w < 4; ++w)
        CK(vlc_rotor_get_wakevel(u->ctx, jr, ib, w, r->nNwake > 0 ? orc_rotor_vel(r, ib, w) : NULL,
                                 r->nFwake > 0 ? orc_rotor_vel(r, ib, 4 + w) : NULL));
      if (orc_case_config(u->cas)->fdScheme >= 4) /* the two extra histories of the multistep schemes */
        for (int w = 4; w < 6; ++w)
          CK(vlc_rotor_get_wakevel(u->ctx, jr, ib, w, r->nNwake > 0 ? orc_rotor_vel(r, ib, 4 + w) : NULL,
                                   r->nFwake > 0 ? orc_rotor_vel(r, ib, 6 + w) : NULL));
    }
  }
  return 0;
}

/* One process per GPU: rank / world, the device exchange buffer (3 x xbuf_targets doubles, xbuf_targets >= world *
 * ceil(M_max / world)) and the all-gather callback (NCCL through torch.distributed in tests/multi_gpu_case.py). */
void case_gpu_hooks_set_sharding(void *handle, int world, int rank, double *xbuf, long xbuf_targets,
                                 int (*exchange)(void *, long), void *arg) {
  gpu_user_t *u = (gpu_user_t *)handle;
  u->world = world;
  u->rank = rank;
  u->xbuf = xbuf;
  u->xbuf_targets = xbuf_targets;
  u->exchange = exchange;
  u->exchange_arg = arg;
}
long case_gpu_hooks_exchanges(void *handle) { return ((gpu_user_t *)handle)->exchanges; }
long case_gpu_hooks_wing_uploads(void *handle) { return ((gpu_user_t *)handle)->wing_uploads; }
long case_gpu_hooks_uploads(void *handle) { return ((gpu_user_t *)handle)->uploads; }
long case_gpu_hooks_skipped(void *handle) { return ((gpu_user_t *)handle)->skipped; }
int case_gpu_hooks_last_rc(void *handle) { return ((gpu_user_t *)handle)->last_rc; }
void case_gpu_hooks_free(void *handle) {
  gpu_user_t *u = (gpu_user_t *)handle;
  if (u && u->emu) {
    for (int ir = 0; ir < u->nr; ++ir) {
      free(u->emu[ir].wiP);
      free(u->emu[ir].sec);
      free(u->emu[ir].loads);
      free(u->emu[ir].rhs);
    }
    free(u->emu);
  }
  free(u);
}
