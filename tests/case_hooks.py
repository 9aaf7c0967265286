"""Hook tables for the oracle's case driver (oracle/vlc_case.h: orc_hooks_t) -- test infrastructure.

The driver (restatement of the reference's `program main`) reaches its five hot-path call sites through function
pointers.  `gpu_hooks` forwards them to the C ABI of include/volcanor_b200.h exactly the way the iso_c_binding shim
does for the Fortran driver: upload the rotor's records, call the batched entry point, hand the
velocities back.  `loopback_hooks` forwards to the CPU oracle through the same Python callbacks (plumbing check).
"""
import ctypes as C

import numpy as np

from oracle import pyoracle

VR, FW = 50, 13


def _as_array(ptr, n):
    return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_double)), shape=(n,))


def _make(case, vind_points, on_nwake, on_fwake, calc_aic, solve):
    H = pyoracle.OrcHooks
    errors = []

    def guard(fn):
        def wrapped(*a):
            try:
                return fn(*a) or 0
            except Exception as e:  # an exception must not cross the C boundary
                errors.append(e)
                return 9
        return wrapped

    def _vp(user, jr, what, predicted, m, P, V):
        out = vind_points(jr, what, bool(predicted), _as_array(P, 3 * m).reshape(m, 3))
        _as_array(V, 3 * m)[:] = np.asarray(out).reshape(-1)

    def _on(user, jr, Nwake, rows, cols, ld, predicted, out):
        res = on_nwake(jr, Nwake, rows, cols, ld, bool(predicted))
        _as_array(out, 3 * rows * (cols + 1))[:] = np.asarray(res).reshape(-1)

    def _of(user, jr, Fwake, rows, predicted, out):
        res = on_fwake(jr, Fwake, rows, bool(predicted))
        _as_array(out, 3 * rows)[:] = np.asarray(res).reshape(-1)

    def _aic(user, ir, AIC, AIC_inv):
        r = case.rotor(ir)
        A, Ainv = calc_aic(ir)
        _as_array(AIC, r.N * r.N)[:] = np.asarray(A).T.reshape(-1)       # column-major
        if Ainv is not None:
            _as_array(AIC_inv, r.N * r.N)[:] = np.asarray(Ainv).T.reshape(-1)

    def _solve(user, ir, RHS, gam):
        N = case.rotor(ir).N
        _as_array(gam, N)[:] = solve(ir, _as_array(RHS, N).copy())

    hooks = H(None, H.VIND_POINTS(guard(_vp)), H.VIND_ONN(guard(_on)), H.VIND_ONF(guard(_of)), H.CALC_AIC(guard(_aic)),
              H.SOLVE(guard(_solve)))
    hooks.errors = errors
    return hooks


def loopback_hooks(case):
    """Same arithmetic as the default (C) hooks, through the Python callback plumbing."""
    lib = case.lib

    def vind_points(jr, what, predicted, P):
        return case.rotor(jr).vind_points(what, P, predicted)

    def on_nwake(jr, Nwake, rows, cols, ld, predicted):
        out = np.empty((cols + 1, rows, 3))
        lib.orc_vind_onNwake_byRotor(case.rotor(jr).h, Nwake, rows, cols, ld, int(predicted), out.ctypes.data)
        return out

    def on_fwake(jr, Fwake, rows, predicted):
        out = np.empty((rows, 3))
        lib.orc_vind_onFwake_byRotor(case.rotor(jr).h, Fwake, rows, int(predicted), out.ctypes.data)
        return out

    def calc_aic(ir):
        r = case.rotor(ir)
        assert r.calcAIC() == 0
        return r.AIC().copy(), r.AIC(inverse=True).copy()

    def solve(ir, rhs):
        r = case.rotor(ir)
        out = np.empty(r.N)
        lib.orc_matmulAX(r.N, r.N, lib.orc_rotor_AIC(r.h, 1), rhs.ctypes.data, out.ctypes.data)
        return out

    return _make(case, vind_points, on_nwake, on_fwake, calc_aic, solve)


def gpu_hooks(case, ctx):
    """Every hot-path call site goes through the C ABI (include/volcanor_b200.h, tier 2)."""
    nr = case.nr
    case.init_rotors()      # the rotors (and their sizes) exist only after rotor%init (main.f90:31-40)
    rotors = [case.rotor(ir) for ir in range(nr)]
    for ir, r in enumerate(rotors):
        ctx.rotor_define(ir, r.nb, r.nc, r.ns, r.nNwake, r.nFwake, 1)
    stats = {"calls": 0}

    def upload(jr, predicted):
        """What the Fortran shim does before a sweep: transfer(blade%waN, buf) etc. and the row counters."""
        r = rotors[jr]
        d = r.dims()
        ctx.rotor_set_rows(jr, d["rowNear"], d["rowFar"])
        for ib in range(r.nb):
            ctx.rotor_put_wing(jr, ib, r.wiP(ib))
            if r.nNwake > 0:
                ctx.rotor_put_nwake(jr, ib, r.waN(ib, predicted), predicted)
            if r.nFwake > 0:
                ctx.rotor_put_fwake(jr, ib, r.waF(ib, predicted), predicted)
        stats["calls"] += 1

    def vind_points(jr, what, predicted, P):
        upload(jr, predicted)
        if what == 0:
            return ctx.rotor_vind_bywing(jr, P)
        if what == 1:
            return ctx.rotor_vind_bywake(jr, P, predicted)
        if what == 2:
            return ctx.rotor_vind(jr, P, predicted)
        return ctx.rotor_vind_bywing_boundVortices(jr, P)

    def on_nwake(jr, Nwake, rows, cols, ld, predicted):
        upload(jr, predicted)
        return ctx.vind_onNwake_byRotor_ptr(jr, Nwake, rows, cols, ld, predicted)

    def on_fwake(jr, Fwake, rows, predicted):
        upload(jr, predicted)
        return ctx.vind_onFwake_byRotor_ptr(jr, Fwake, rows, predicted)

    def calc_aic(ir):
        upload(ir, False)
        return ctx.rotor_calcAIC(ir, rotors[ir].N), None

    def solve(ir, rhs):
        return ctx.rotor_solve(ir, rhs)

    h = _make(case, vind_points, on_nwake, on_fwake, calc_aic, solve)
    h.stats = stats
    return h


class OracleBackedContext:
    """Stand-in for volcanor_b200.Context with the same tier-2 methods, computing with the CPU oracle from the
    UPLOADED copies only.  Lets the CPU suite check the shim logic of gpu_hooks (what is uploaded when, row
    counters, predicted sets) without a GPU: a forgotten upload shows up as a diverging history."""

    def __init__(self):
        self.rotors = {}

    def rotor_define(self, ir, nb, nc, ns, nNwake, nFwake, surfaceType=1):
        self.rotors[ir] = pyoracle.Rotor(nb, nc, ns, nNwake, nFwake)
        self.rotors[ir].set_params(surfaceType=surfaceType)

    def rotor_set_rows(self, ir, rowNear, rowFar):
        self.rotors[ir].set_rows(rowNear, rowFar)

    def rotor_put_wing(self, ir, ib, wiP):
        self.rotors[ir].wiP(ib)[...] = wiP

    def rotor_put_nwake(self, ir, ib, waN, predicted=False):
        self.rotors[ir].waN(ib, predicted)[...] = waN

    def rotor_put_fwake(self, ir, ib, waF, predicted=False):
        self.rotors[ir].waF(ib, predicted)[...] = waF

    def rotor_vind_bywing(self, ir, P):
        return self.rotors[ir].vind_points(0, P)

    def rotor_vind_bywake(self, ir, P, predicted=False):
        return self.rotors[ir].vind_points(1, P, predicted)

    def rotor_vind(self, ir, P, predicted=False):
        return self.rotors[ir].vind_points(2, P, predicted)

    def rotor_vind_bywing_boundVortices(self, ir, P):
        return self.rotors[ir].vind_points(3, P)

    def vind_onNwake_byRotor_ptr(self, ir, ptr, rows, cols, ld, predicted=False):
        r = self.rotors[ir]
        out = np.empty((cols + 1, rows, 3))
        r.lib.orc_vind_onNwake_byRotor(r.h, ptr, rows, cols, ld, int(predicted), out.ctypes.data)
        return out

    def vind_onFwake_byRotor_ptr(self, ir, ptr, rows, predicted=False):
        r = self.rotors[ir]
        out = np.empty((rows, 3))
        r.lib.orc_vind_onFwake_byRotor(r.h, ptr, rows, int(predicted), out.ctypes.data)
        return out

    def rotor_calcAIC(self, ir, N, want_matrix=True):
        r = self.rotors[ir]
        assert r.calcAIC() == 0
        return np.asfortranarray(r.AIC().copy())

    def rotor_solve(self, ir, rhs):
        r = self.rotors[ir]
        out = np.empty(r.N)
        r.lib.orc_matmulAX(r.N, r.N, r.lib.orc_rotor_AIC(r.h, 1), rhs.ctypes.data, out.ctypes.data)
        return out
