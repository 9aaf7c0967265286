"""TEST INFRASTRUCTURE: a stand-in for volcanor_b200.Context that emulates tier 2c of the C ABI on the CPU -- its own
copies of the wing records, the g++ build of volcanor_b200/csrc/cp_stage.cuh (tests/native/cp_stage_host.cpp) for the
record arithmetic, the oracle's rotors for the sweeps and the solve -- so that the BODIES of the GPU tests in
tests/test_zz_gpu_cp_stage.py (what they upload, what they compare, with which tolerances) also run without a GPU
(tests/test_cp_stage_host.py).  The Python twin of the CPU emulation in tests/native/case_gpu_hooks.c."""
import numpy as np

WP = 104


class EmulatedCpContext:
    def __init__(self, host_lib):
        self.lib_host = host_lib
        self.rot = {}          # ir -> dict(nb, nc, ns, wiP (nb, ns*nc, 104), sec, loads, rhs, src (oracle rotor), params)
        self.launch_count = 0

    # ---- what the tests call on a Context
    def rotor_define(self, ir, nb, nc, ns, nNwake, nFwake, surfaceType=1):
        self.rot[ir] = dict(nb=nb, nc=nc, ns=ns, wiP=np.zeros((nb, ns * nc, WP)), sec=np.zeros((nb, 10 * ns + 6)),
                            loads=np.zeros((nb, 12 + 25 * ns)), rhs=None, src=None, nbConvect=nb, axisym=0,
                            have_sec=[False] * nb)

    def rotor_set_wake_params(self, ir, nbConvect, axisymmetrySwitch, *rest):
        self.rot[ir].update(nbConvect=nbConvect, axisym=axisymmetrySwitch)

    def rotor_set_rows(self, ir, rowNear, rowFar):
        pass

    def rotor_put_wing(self, ir, ib, wiP):
        self.rot[ir]["wiP"][ib] = np.asarray(wiP).reshape(-1, WP)

    def rotor_put_nwake(self, ir, ib, waN, predicted=False):
        pass                   # the sweeps read the oracle rotor attached with attach_sources

    def rotor_put_fwake(self, ir, ib, waF, predicted=False):
        pass

    def attach_sources(self, ir, oracle_rotor):
        """The emulation sweeps with the oracle: rotor ir's sources (wake, and wing geometry) are this oracle rotor's; the
        wing circulations are kept in sync by rotor_solve_map_gam."""
        self.rot[ir]["src"] = oracle_rotor

    def rotor_calcAIC(self, ir, N, want_matrix=True):
        return None            # the oracle rotor's AIC_inv (rot.calcAIC()) is used by rotor_solve_map_gam

    def _sweep_acc(self, ir, jr, what, field, sign):
        r = self.rot[ir]
        m = r["nbConvect"] * r["ns"] * r["nc"]
        w = r["wiP"].reshape(-1, WP)
        V = self.rot[jr]["src"].vind_points(what, w[:m, 64:67].copy())
        w[:m, field:field + 3] = w[:m, field:field + 3] + V if sign > 0 else w[:m, field:field + 3] - V
        self.launch_count += 1

    def rotor_calc_RHS(self, ir, m, N, want_velCP=True, want_RHS=True):
        r = self.rot[ir]
        for jr in sorted(self.rot):
            self._sweep_acc(ir, jr, 1, 76, +1)
            if jr != ir:
                self._sweep_acc(ir, jr, 0, 76, +1)
        rhs = np.zeros(N)
        w = np.ascontiguousarray(r["wiP"].reshape(-1))
        self.lib_host.cp_host_rhs(N, r["ns"] * r["nc"], r["nbConvect"], r["axisym"], w.ctypes.data, rhs.ctypes.data)
        r["rhs"] = rhs
        return r["wiP"].reshape(-1, WP)[:m, 76:79].copy(), rhs.copy()

    def rotor_solve_map_gam(self, ir, N, want_gamVec=True):
        import volcanor_b200 as vb
        r = self.rot[ir]
        if r["rhs"] is None:
            raise vb.VlcError("vlc_rotor_solve_map_gam before vlc_rotor_calc_RHS")
        g = np.ascontiguousarray(r["src"].AIC(inverse=True) @ r["rhs"])
        w = np.ascontiguousarray(r["wiP"].reshape(-1))
        self.lib_host.cp_host_map_gam(r["nb"], r["ns"] * r["nc"], r["nbConvect"], r["axisym"], g.ctypes.data, w.ctypes.data)
        r["wiP"] = w.reshape(r["nb"], -1, WP)
        r["rhs"] = None
        return g

    def rotor_put_sections(self, ir, ib, sec):
        self.rot[ir]["sec"][ib] = sec
        self.rot[ir]["have_sec"][ib] = True

    def _axisym_field(self, r, field, n):
        for ib in range(1, r["nb"]):
            r["wiP"][ib][:, field:field + n] = r["wiP"][0][:, field:field + n]

    def rotor_calc_velCPTotal(self, ir):
        r = self.rot[ir]
        m = r["nbConvect"] * r["ns"] * r["nc"]
        w = r["wiP"].reshape(-1, WP)
        w[:m, 79:82] = w[:m, 76:79]
        for jr in sorted(self.rot):
            self._sweep_acc(ir, jr, 3, 79, -1)
        self._sweep_acc(ir, ir, 0, 79, +1)
        if r["axisym"] == 1:
            self._axisym_field(r, 79, 3)

    def rotor_calc_force(self, ir, density, dt, Omega, spanwiseLiftSwitch=0):
        import volcanor_b200 as vb
        r = self.rot[ir]
        if not dt > 0.0:
            raise vb.VlcError("dt must be positive")
        if not all(r["have_sec"][:r["nbConvect"]]):
            raise vb.VlcError("vlc_rotor_calc_force before vlc_rotor_put_sections of every convected blade")
        w = np.ascontiguousarray(r["wiP"].reshape(-1))
        sec = np.ascontiguousarray(r["sec"].reshape(-1))
        loads = np.ascontiguousarray(r["loads"].reshape(-1))
        self.lib_host.cp_host_loads(r["nbConvect"], r["nc"], r["ns"], float(density), float(dt), float(Omega),
                                    int(spanwiseLiftSwitch), w.ctypes.data, sec.ctypes.data, loads.ctypes.data)
        r["wiP"], r["loads"] = w.reshape(r["nb"], -1, WP), loads.reshape(r["nb"], -1)
        if r["axisym"] == 1:
            for field, n in ((95, 2), (50, 1), (85, 6)):
                self._axisym_field(r, field, n)
            ns = r["ns"]
            for ib in range(1, r["nb"]):
                keep = r["loads"][ib][12:12 + 3 * ns].copy()        # secChordwiseResVel is not copied
                r["loads"][ib] = r["loads"][0]
                r["loads"][ib][12:12 + 3 * ns] = keep
        self.launch_count += 1

    def rotor_get_loads(self, ir, ib, ns):
        import volcanor_b200 as vb
        a = self.rot[ir]["loads"][ib].copy()
        out = {"forceInertial": a[0:3], "lift": a[3:6], "drag": a[6:9], "liftUnsteady": a[9:12]}
        for k, name in enumerate(vb.Context.LOADS_SEC3):
            out[name] = a[12 + 3 * ns * k: 12 + 3 * ns * (k + 1)].reshape(ns, 3)
        for k, name in enumerate(vb.Context.LOADS_SEC1):
            out[name] = a[12 + 21 * ns + ns * k: 12 + 21 * ns + ns * (k + 1)]
        return out

    def rotor_get_wing(self, ir, ib, nc, ns):
        return self.rot[ir]["wiP"][ib].reshape(ns, nc, WP).copy()
