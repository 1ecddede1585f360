"""The lane-per-pair EPA tier and the device GJK, compiled for the CPU from the CUDA source
(tests/host_sim) and compared bit for bit with the oracle.  Runs without a GPU: it checks the
kernel's control flow (work-list fetch, visibility-mask walk, literal edge list, dump / resume
between tiers), not the device build — the `-m gpu` parity tests do that."""
import numpy as np
import pytest

from collision_rs_b200 import workloads as W
from host_sim.sim import Sim
from oracle.binding import Settings

FULL = 0
OVERFLOW = 7


@pytest.fixture(scope="module")
def sim():
    return Sim()


# 0: the lane kernel (one warp per 32 pairs); 2..4: the warp-team kernel with that many warps
@pytest.fixture(params=[0, 2, 3, 4], ids=lambda w: f"team{w}")
def team_w(request):
    return request.param


def _check(sim, oracle, w, caps, settings=None, allow_overflow=0.0, team_w=0, **kw):
    kw["team_w"] = team_w
    args = (w["l_shape"], w["r_shape"], w["l_t"], w["r_t"])
    got, gst, tiers = sim.intersection_full(w["shapes"], w["verts"], *args, settings=settings, caps=caps, **kw)
    exp, est = oracle.intersection(FULL, w["shapes"], w["verts"], *args, settings=settings, nthreads=8)
    over = gst == OVERFLOW
    assert over.mean() <= allow_overflow, (over.mean(), tiers)
    assert not (est[~over] == OVERFLOW).any()
    assert np.array_equal(gst[~over], est[~over])
    for f in ("strategy", "normal", "depth", "point", "toi"):
        a = np.ascontiguousarray(got[f][~over]).view(np.uint32)
        b = np.ascontiguousarray(exp[f][~over]).view(np.uint32)
        assert np.array_equal(a, b), f
    return tiers, gst


def test_polyhedron_pairs_default_caps(sim, oracle, team_w):
    w = W.cfg3_polyhedron_pairs(20_000, n_shapes=256)
    tiers, gst = _check(sim, oracle, w, team_w=team_w, caps=0, allow_overflow=0.02)
    assert tiers[0] > 15_000 and 0 < tiers[1] < tiers[0] // 3


def test_polyhedron_pairs_small_caps(sim, oracle, team_w):
    w = W.cfg3_polyhedron_pairs(20_000, n_shapes=256, seed=31)
    tiers, gst = _check(sim, oracle, w, team_w=team_w, caps=1, allow_overflow=0.02)
    assert tiers[1] > 0


def test_tiny_caps_exercise_dump_and_resume(sim, oracle, team_w):
    # (12, 8, 8) + (24, 14, 8): most pairs leave tier 0 mid-flight and are resumed by tier 1
    w = W.cfg1_cuboid_pairs(20_000)
    tiers, gst = _check(sim, oracle, w, team_w=team_w, caps=2, allow_overflow=0.2)
    assert tiers[1] > 1000
    # without room for dumps every overflowing pair restarts from its simplex in tier 1
    _check(sim, oracle, w, team_w=team_w, caps=2, allow_overflow=0.2, dump_cap0=0)
    _check(sim, oracle, w, team_w=team_w, caps=2, allow_overflow=0.2, dump_cap0=100)


def test_cuboid_pairs(sim, oracle, team_w):
    for spread in (0.75, 0.05):
        w = W.cfg1_cuboid_pairs(10_000, spread=spread, seed=5)
        _check(sim, oracle, w, team_w=team_w, caps=0)
        _check(sim, oracle, w, team_w=team_w, caps=0, prefetch=False)


def test_settings_are_honoured(sim, oracle, team_w):
    w = W.cfg3_polyhedron_pairs(5_000, n_shapes=128, seed=9)
    for st in (Settings(np.float32(1e-6), np.float32(1e-6), np.float32(1e-3), 100),
               Settings(np.float32(1e-6), np.float32(1e-6), np.float32(1e-7), 12),
               Settings(np.float32(1e-6), np.float32(1e-6), np.float32(1e-5), 1)):
        _check(sim, oracle, w, team_w=team_w, caps=0, settings=st, allow_overflow=0.02)


def test_mixed_kinds_with_curved_shapes(sim, oracle, team_w):
    # curved pairs run EPA to the iteration cap and outgrow both lane tiers (the CUDA build sends
    # them to the group tiers): everything that fits must still match
    w = W.cfg2_mixed_pairs(10_000, spread=1.0)
    _check(sim, oracle, w, team_w=team_w, caps=0, allow_overflow=0.6)


def test_team_literal_horizon_path(sim, oracle):
    # the bit-matrix horizon falls back to a literal replay of remove_or_add_edge when a directed
    # edge occurs twice (about one pair in 10 M): force that path for every pair
    sim.force_literal_horizon(True)
    try:
        w = W.cfg3_polyhedron_pairs(10_000, n_shapes=128, seed=3)
        _check(sim, oracle, w, team_w=4, caps=0, allow_overflow=0.02)
        w = W.cfg1_cuboid_pairs(10_000)
        _check(sim, oracle, w, team_w=2, caps=2, allow_overflow=0.2)
    finally:
        sim.force_literal_horizon(False)
