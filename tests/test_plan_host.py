"""CPU tests of the launch planner (volcanor_b200/csrc/plan.hpp, compiled for the host by tests/native/plan_host.cpp): how
a sweep is cut into source splits for a given number of target tiles and CTA slots.  The split decides the machine fill
and a target's summation order; the numbers below are the launch shapes of the measured runs (profiles/r02m_*,
profiles/r02v_small_cases.md)."""
import ctypes as C
import subprocess
from pathlib import Path

import pytest

HERE = Path(__file__).resolve().parent / "native"
SLOTS = 148 * 2   # bs_lattice_kernel<4,2>: two CTAs per SM on a B200


@pytest.fixture(scope="module")
def plib():
    subprocess.run(["make", "-C", str(HERE)], check=True, capture_output=True)
    lib = C.CDLL(str(HERE / "libplan_host.so"))
    LL = C.c_longlong
    lib.plan_small_split.argtypes = [LL, LL, LL, C.c_int]
    lib.plan_wave_split.argtypes = [LL, LL, LL, LL, C.c_int]
    lib.plan_cut.argtypes = [LL, LL, C.c_int, C.POINTER(C.c_int), C.POINTER(LL)]
    return lib


def _cut(lib, n_pad, unit, nsplit):
    ns, chunk = C.c_int(), C.c_longlong()
    lib.plan_cut(n_pad, unit, nsplit, C.byref(ns), C.byref(chunk))
    return ns.value, chunk.value


def test_headline_sweep_keeps_its_launch_shape(plib):
    """10^6 filaments x 258 176 targets on one B200: 62 560 strip records of width 4 (1 955 tiles of 32), 1 009 target tiles
    of 256: not a small sweep; 12 splits = 40.9 waves, the shape of every profile of the round."""
    ttiles, tiles = (258176 + 255) // 256, 62560 // 32
    assert plib.plan_small_split(ttiles, tiles, SLOTS, 4) == 0
    s = plib.plan_wave_split(ttiles, tiles, SLOTS, 347, 4)
    assert s == 12 and plib.plan_wave_split(ttiles, tiles, SLOTS, 347, 1) == 12
    ns, chunk = _cut(plib, 62560, 8, s)
    assert ns == 12 and chunk % 8 == 0 and (ns - 1) * chunk < 62560 <= ns * chunk
    waves = ttiles * ns / SLOTS
    assert waves / -(-waves // 1) > 0.99


@pytest.mark.parametrize("ttiles,tiles,slots,per_tile", [(15, 28, 296, 4), (28, 14, 592, 4), (2, 40, 740, 1), (1, 391, 296, 4),
                                                         (4, 350, 444, 4), (1, 1, 296, 4), (1, 1, 740, 1), (57, 9, 296, 4)])
def test_small_sweeps_are_split_for_parallelism(plib, ttiles, tiles, slots, per_tile):
    """A sweep that cannot fill two waves with chunks of >= 4 tiles (the reference's own cases): chunks go below 4 tiles --
    down to a quarter tile for the lattice kernel --, never more than 256 splits, never more CTAs than about two waves
    unless the chunks are already minimal, and the cut covers every record exactly once."""
    s = plib.plan_small_split(ttiles, tiles, slots, per_tile)
    units = tiles * per_tile
    if units <= 1:
        assert s == 0
        return
    assert 1 <= s <= min(units, 256)
    assert s >= min(max(1, tiles // 4), units)      # at least the parallelism the old ">= 4 tiles per chunk" rule gave
    unit = 8
    ns, chunk = _cut(plib, units * unit, unit, s)
    assert ns == s and chunk % unit == 0 and (ns - 1) * chunk < units * unit <= ns * chunk
    assert ttiles * ns <= 2 * slots or chunk == unit or ns == 256


def test_flat_sweep_of_the_headline_size_is_split_into_whole_waves(plib):
    """The same-work sweep of bench.py (flat kernel, T = 4: 505 target tiles of 512, 7 815 source tiles of 128, 3 CTAs per
    SM): 7 splits = 7.96 waves.  (A planner bug of r02t left such sweeps at ONE split = 1.14 waves: 806 instead of 678 ms.)"""
    slots = 148 * 3
    assert plib.plan_small_split(505, 7815, slots, 1) == 0
    s = plib.plan_wave_split(505, 7815, slots, 347, 1)
    assert s == 7
    waves = 505 * s / slots
    assert waves / -(-waves // 1) > 0.99


def test_quarter_tile_units_reach_whole_waves_where_tiles_cannot(plib):
    """A late caradonna step: 95 target tiles x 175 strip-record tiles on 296 slots.  In whole tiles the best fit is 3 splits
    (285 CTAs, 0.963 of one wave); in quarter tiles 28 splits (2 660 CTAs = 8.99 waves) -- measured 1.5 % faster over the case
    (profiles/r02v_small_cases.md, scan of fixed splits)."""
    assert plib.plan_wave_split(95, 175, SLOTS, 256, 1) == 3
    s = plib.plan_wave_split(95, 175, SLOTS, 256, 4)
    waves = 95 * s / SLOTS
    assert waves >= 8 and waves / -(-waves // 1) > 0.995
    ns, chunk = _cut(plib, 175 * 32, 8, s)
    assert ns == s and chunk >= 4 * 32


def test_a_sweep_that_fills_the_machine_is_not_small(plib):
    for ttiles, tiles in [(1009, 1955), (113, 391), (95, 175), (600, 8)]:
        assert plib.plan_small_split(ttiles, tiles, SLOTS, 4) == 0, (ttiles, tiles)


def test_wave_search_respects_the_partial_buffer_cap_and_chunk_floor(plib):
    assert plib.plan_wave_split(1009, 1955, SLOTS, 5, 4) <= 5          # caller's cap (2 GiB of partial sums)
    assert plib.plan_wave_split(1009, 7, SLOTS, 256, 4) == 1            # fewer than 8 tiles: one chunk
    for tiles in (16, 64, 391, 1955, 100000):
        for per_tile in (1, 4):
            s = plib.plan_wave_split(113, tiles, SLOTS, 256, per_tile)
            assert 1 <= s <= min(256, max(1, tiles // 4))


@pytest.mark.parametrize("n_pad,unit,nsplit", [(62560, 32, 12), (62560, 8, 12), (128, 16, 5), (64, 64, 3), (4096, 128, 7), (32, 8, 100)])
def test_cut_covers_all_records_in_unit_multiples(plib, n_pad, unit, nsplit):
    ns, chunk = _cut(plib, n_pad, unit, nsplit)
    assert 1 <= ns <= nsplit and chunk % unit == 0
    assert (ns - 1) * chunk < n_pad <= ns * chunk


def test_planner_invariants_on_random_shapes(plib):
    """Whatever the sweep looks like: 1 <= splits <= 256, the cut covers every record exactly once in multiples of the unit,
    a wave split never cuts below 4 tiles per chunk, a small split never asks for more splits than there are units."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=400, deadline=None)
    @given(ttiles=st.integers(1, 20000), tiles=st.integers(1, 200000), occ=st.integers(1, 6), per_tile=st.sampled_from([1, 4]),
           cap=st.integers(1, 4096))
    def check(ttiles, tiles, occ, per_tile, cap):
        slots = 148 * occ
        unit = 32 // per_tile if per_tile == 4 else 32
        n_pad = tiles * 32
        s = plib.plan_small_split(ttiles, tiles, slots, per_tile)
        if s:
            assert 1 <= s <= min(256, tiles * per_tile)
            ns, chunk = _cut(plib, n_pad, 32 // per_tile, s)
            assert ns == s and chunk % (32 // per_tile) == 0 and (ns - 1) * chunk < n_pad <= ns * chunk
        w = plib.plan_wave_split(ttiles, tiles, slots, cap, per_tile)
        assert 1 <= w <= max(1, min(256, cap, tiles // 4))
        ns, chunk = _cut(plib, n_pad, unit, w)
        assert ns == w and chunk % unit == 0 and (ns - 1) * chunk < n_pad <= ns * chunk
        assert chunk >= min(n_pad, 4 * 32) - 32 or w == 1

    check()
