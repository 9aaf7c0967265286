"""Shared helpers of the parity tests (test infrastructure)."""
import numpy as np

VF, VR, FW = 12, 50, 13


def lattice_to_rotor(lat, rotor, ib, predicted=False):
    """Fill blade ib of an oracle Rotor (reference record layout) from a synthetic lattice:
    ring (i, j) corners per vr_assignP (classdef.f90:569-592), rVc per filament, gam; far chain."""
    S, R = lat.S, lat.R
    waN = rotor.waN(ib, predicted)
    nd = lat.nodes
    c = [nd[:-1, :-1], nd[:-1, 1:], nd[1:, 1:], nd[1:, :-1]]
    for f in range(4):
        waN[:, :, VF * f + 0:VF * f + 3] = c[f]
        waN[:, :, VF * f + 3:VF * f + 6] = c[(f + 1) % 4]
        waN[:, :, VF * f + 8] = lat.rvc4[:, :, f]
        waN[:, :, VF * f + 9] = lat.rvc4[:, :, f]
    waN[:, :, 48] = lat.gam
    if lat.F > 0:
        waF = rotor.waF(ib, predicted)
        waF[:, 0:3] = lat.far_nodes[1:]
        waF[:, 3:6] = lat.far_nodes[:-1]
        waF[:, 8] = lat.rvcF
        waF[:, 9] = lat.rvcF
        waF[:, 12] = lat.gamF


def scaled_err(V, Vref, Vabs):
    """PER-TARGET parity measure: max_t ( max_k |V_tk - Vref_tk| / max_k sum_i |term_i,tk| ).

    Every target is measured against ITS OWN velocity scale sum|terms|, so a target far from the wake --
    whose scale is orders of magnitude below the batch maximum -- cannot hide a dropped edge or a wrong merged
    strength behind the large targets of the batch (round-1 review: the batch-scaled form could)."""
    V, Vref, Vabs = (np.asarray(a, dtype=np.float64).reshape(-1, 3) for a in (V, Vref, Vabs))
    if V.shape[0] == 0:
        return 0.0
    scale = np.maximum(Vabs.max(axis=1), 1e-300)
    return float(np.max(np.abs(V - Vref).max(axis=1) / scale))


def scaled_err_batch(V, Vref, Vabs):
    """Batch-scaled form (round 1): max |V - Vref| / max sum|terms| -- kept as a diagnostic next to scaled_err."""
    return float(np.max(np.abs(V - Vref)) / max(np.max(Vabs), 1e-300))
