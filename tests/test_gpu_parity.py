"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle on identical inputs.

Bar (north_star): Option/boolean outcomes and pair sets bit-exact; contact normal / depth /
point / toi within 1e-4 relative.  The kernels follow the reference's unfused f32 operation
order, so these tests assert the STRONGER property: every f32 bit-identical (NaN == NaN).
"""
import numpy as np
import pytest

from collision_rs_b200 import abi as host_abi
from collision_rs_b200 import workloads as W
from collision_rs_b200.host import (Cb2Error, CollisionStrategy, ConvexPolyhedron, Cuboid,
                                    Decomposed, GJK3, MultiContext, ReferencePanic, Sphere,
                                    SweepAndPrune3)
from helpers import transform_3d

pytestmark = pytest.mark.gpu

EPS = np.float32(1.1920929e-7)


def bits_equal(a, b):
    """bitwise equality of f32 arrays, treating any NaN == any NaN"""
    a = np.ascontiguousarray(a, np.float32)
    b = np.ascontiguousarray(b, np.float32)
    same = a.view(np.uint32) == b.view(np.uint32)
    same |= np.isnan(a) & np.isnan(b)
    return same


def assert_contacts_equal(got, gst, exp, est, what=""):
    assert np.array_equal(gst, est), (
        f"{what}: status mismatch at {np.nonzero(gst != est)[0][:10]}: "
        f"gpu {gst[gst != est][:10]} oracle {est[gst != est][:10]}")
    m = est == 1
    for f in ("normal", "depth", "point", "toi"):
        ok = bits_equal(got[f][m], exp[f][m])
        assert ok.all(), f"{what}: {f} differs on {int((~ok).sum())} values"
    assert np.array_equal(got["strategy"], exp["strategy"])


def D(x, y, z, angle_z=0.0):
    t = transform_3d(x, y, z, angle_z)
    return Decomposed(tuple(t["d"][0]), tuple(t["q"][0]), float(t["scale"][0]))


# ---- the reference's own 3-D tests, through the reference-shaped API on the GPU ----------------
def test_gjk_exact_3d(ctx):  # gjk/mod.rs:634-642
    gjk = GJK3.new(ctx)
    shape, t = Cuboid(1, 1, 1), D(0, 0, 0)
    assert gjk.intersection(CollisionStrategy.FullResolution, shape, t, shape, t) is not None
    assert gjk.distance(shape, t, shape, t) is None


def test_gjk_sphere(ctx):  # gjk/mod.rs:645-653
    gjk = GJK3.new(ctx)
    shape, t = Sphere(1.0), D(0, 0, 0)
    assert gjk.intersection(CollisionStrategy.FullResolution, shape, t, shape, t) is not None
    assert gjk.distance(shape, t, shape, t) is None


def test_gjk_3d_hit(ctx):  # gjk/mod.rs:700-720
    gjk = GJK3.new(ctx)
    left, right = Cuboid(10, 10, 10), Cuboid(10, 10, 10)
    c = gjk.intersection(CollisionStrategy.FullResolution, left, D(15, 0, 0), right, D(7, 2, 0))
    assert c is not None
    assert c.normal == (-1, 0, 0)
    assert np.float32(2) - c.penetration_depth <= EPS
    assert np.allclose(c.contact_point, (10, 1, 5), rtol=0, atol=5e-6)
    c = gjk.intersection(CollisionStrategy.CollisionOnly, left, D(15, 0, 0), right, D(7, 2, 0))
    assert c is not None and c.strategy == CollisionStrategy.CollisionOnly
    assert c.normal == (0, 0, 0) and c.penetration_depth == 0
    assert gjk.intersection(CollisionStrategy.FullResolution, left, D(15, 0, 0), right,
                            D(-15, 0, 0)) is None


def test_gjk_distance_3d(ctx):  # gjk/mod.rs:743-759
    gjk = GJK3.new(ctx)
    left, right = Cuboid(10, 10, 10), Cuboid(10, 10, 10)
    assert gjk.distance(left, D(15, 0, 0), right, D(7, 2, 0)) is None
    assert gjk.distance(left, D(15, 0, 0), right, D(1, 0, 0)) == np.float32(4.0)


def test_gjk_time_of_impact_3d(ctx):  # gjk/mod.rs:795-825
    gjk = GJK3.new(ctx)
    left, right = Cuboid(10, 11, 10), Cuboid(10, 15, 10)
    c = gjk.intersection_time_of_impact(left, (D(0, 0, 0), D(30, 0, 0)), right,
                                        (D(15, 0, 0), D(15, 0, 0)))
    assert c is not None
    assert abs(c.time_of_impact - np.float32(0.1666667)) <= 2 * EPS
    assert c.normal == (-1, 0, 0)
    assert np.float32(0) - c.penetration_depth <= EPS
    assert c.contact_point == (10, 0, 0)
    assert gjk.intersection_time_of_impact(left, (D(0, 0, 0), D(0, 0, 0)), right,
                                           (D(15, 0, 0), D(15, 0, 0))) is None


def test_polyhedron_pair_single(ctx, oracle):
    # a tetrahedron against a cuboid through the single-pair API
    gjk = GJK3.new(ctx)
    tet = ConvexPolyhedron([[1, 0, 0], [0, 1, 0], [0, 0, 1], [0, 0, 0]])
    c = gjk.intersection(CollisionStrategy.FullResolution, tet, D(0.2, 0.1, 0.0, 0.3),
                         Cuboid(1, 1, 1), D(0.5, 0.2, 0.1))
    assert c is not None and c.penetration_depth > 0


def test_sweep_prune_reference_tests(ctx):  # sweep_prune.rs:317-361 (z padded to (0,1))
    def run(boxes):
        sap = SweepAndPrune3.new(ctx)
        items = [((x0, y0, 0), (x1, y1, 1)) for (x0, y0, x1, y1) in boxes]
        return sap.find_collider_pairs(items), items, sap.get_sweep_axis()

    assert run([(8, 8, 10, 11), (12, 13, 18, 18)])[0] == []
    assert run([(12, 13, 18, 18), (8, 8, 10, 11)])[0] == []
    assert run([(8, 8, 10, 11), (9, 10, 18, 18)])[0] == [(0, 1)]
    pairs, items, axis = run([(9, 10, 18, 18), (8, 8, 10, 11)])
    assert pairs == [(0, 1)] and axis == 0
    assert items[0][0][0] == 8  # slice was re-sorted in place


# ---- batch parity against the oracle -------------------------------------------------------------------
def _upload(ctx, w):
    ctx.upload_shapes(w["shapes"], w.get("verts"))


@pytest.mark.parametrize("spread", [0.75, 1.5])
def test_cfg1_cuboid_full_resolution(ctx, oracle, spread):
    w = W.cfg1_cuboid_pairs(10_000, spread=spread)
    _upload(ctx, w)
    args = (w["l_shape"], w["r_shape"], w["l_t"], w["r_t"])
    exp, est = oracle.intersection(0, w["shapes"], None, *args, nthreads=8)
    got, gst = ctx.intersection(0, *args)
    assert_contacts_equal(got, gst, exp, est, "cfg1 FullResolution")
    assert 0.2 < (est == 1).mean() < 0.99
    exp, est = oracle.intersection(1, w["shapes"], None, *args, nthreads=8)
    got, gst = ctx.intersection(1, *args)
    assert_contacts_equal(got, gst, exp, est, "cfg1 CollisionOnly")


def test_cfg2_mixed_collision_only_distance_full(ctx, oracle):
    w = W.cfg2_mixed_pairs(60_000)
    _upload(ctx, w)
    args = (w["l_shape"], w["r_shape"], w["l_t"], w["r_t"])
    exp, est = oracle.intersection(1, w["shapes"], None, *args, nthreads=8)
    got, gst = ctx.intersection(1, *args)
    assert_contacts_equal(got, gst, exp, est, "cfg2 CollisionOnly")
    ed, est = oracle.distance(w["shapes"], None, *args, nthreads=8)
    gd, gst = ctx.distance(*args)
    assert np.array_equal(gst, est), f"distance status differs on {(gst != est).sum()}"
    assert bits_equal(gd[est == 1], ed[est == 1]).all()
    assert set(np.unique(est)) >= {0, 1, 2, 3}  # the panic paths are exercised
    # FullResolution on curved shapes: EPA runs to the 100-iteration cap (103 verts / 202 faces)
    exp, est = oracle.intersection(0, w["shapes"], None, *args, nthreads=8)
    got, gst = ctx.intersection(0, *args)
    assert_contacts_equal(got, gst, exp, est, "cfg2 FullResolution")


def test_cfg3_polyhedra_full_resolution(ctx, oracle):
    w = W.cfg3_polyhedron_pairs(8_000, n_shapes=512)
    _upload(ctx, w)
    args = (w["l_shape"], w["r_shape"], w["l_t"], w["r_t"])
    exp, est = oracle.intersection(0, w["shapes"], w["verts"], *args, nthreads=8)
    got, gst = ctx.intersection(0, *args)
    assert_contacts_equal(got, gst, exp, est, "cfg3 FullResolution")
    ed, est = oracle.distance(w["shapes"], w["verts"], *args, nthreads=8)
    gd, gst = ctx.distance(*args)
    assert np.array_equal(gst, est)
    assert bits_equal(gd[est == 1], ed[est == 1]).all()


def test_cfg3_odd_vertex_counts(ctx, oracle):
    # ragged polyhedra: 1..70 vertices incl. counts that are not multiples of the sub-warp width
    w = W.cfg3_polyhedron_pairs(4_000, n_shapes=256, seed=5,
                                sizes=(1, 2, 3, 4, 5, 7, 8, 9, 15, 16, 17, 31, 33, 63, 70))
    _upload(ctx, w)
    args = (w["l_shape"], w["r_shape"], w["l_t"], w["r_t"])
    exp, est = oracle.intersection(0, w["shapes"], w["verts"], *args, nthreads=8)
    got, gst = ctx.intersection(0, *args)
    assert_contacts_equal(got, gst, exp, est, "ragged polyhedra")


def test_mixed_table_all_classes(ctx, oracle):
    """One shape table with spheres, cuboids and polyhedra of 4..1500 vertices: every size class
    of the classified dispatch (thread-per-pair, staged sub-warp groups of 4/8/16, unstaged 32)."""
    rng = np.random.default_rng(31)
    ps, pv = W.polyhedron_table(96, seed=9, sizes=(4, 20, 40, 90, 200, 400, 700, 1500))
    prim = W.mixed_shapes(rng, 64)
    shapes = np.concatenate([ps, prim])
    n = 6_000
    l = rng.integers(0, len(shapes), n).astype(np.uint32)
    r = rng.integers(0, len(shapes), n).astype(np.uint32)
    lt, rt = W.random_xforms(rng, n, 0.8), W.random_xforms(rng, n, 0.8)
    lt["scale"][::5] = rng.uniform(0.5, 1.5, len(lt["scale"][::5])).astype(np.float32)
    rt["scale"][::131] = 0.0
    ctx.upload_shapes(shapes, pv)
    for strategy in (0, 1):
        exp, est = oracle.intersection(strategy, shapes, pv, l, r, lt, rt, nthreads=8)
        got, gst = ctx.intersection(strategy, l, r, lt, rt)
        assert_contacts_equal(got, gst, exp, est, f"mixed table strategy {strategy}")
    assert (est == 4).sum() > 0 and (est == 1).sum() > 1000


def test_non_manifold_polytopes_are_retried(ctx, oracle):
    """~5 curved pairs per million grow a non-manifold EPA polytope far past 2V-4 faces (this
    sample peaks at 1050 faces): the fast kernels flag them and k_retry_overflow redoes them."""
    w = W.cfg2_mixed_pairs(500_000, seed=1, spread=0.8)
    _upload(ctx, w)
    args = (w["l_shape"], w["r_shape"], w["l_t"], w["r_t"])
    exp, est, stats = oracle.intersection(0, w["shapes"], None, *args, nthreads=8, want_stats=True)
    assert stats[:, 2].max() > 256
    got, gst = ctx.intersection(0, *args)
    assert_contacts_equal(got, gst, exp, est, "non-manifold retry")
    assert (gst == 7).sum() == 0


def test_cfg5_time_of_impact(ctx, oracle):
    w = W.cfg5_toi_pairs(40_000)
    _upload(ctx, w)
    args = (w["l_shape"], w["r_shape"], w["l_t0"], w["l_t1"], w["r_t0"], w["r_t1"])
    exp, est = oracle.time_of_impact(w["shapes"], None, *args, iter_cap=256, nthreads=8)
    got, gst = ctx.time_of_impact(*args, iter_cap=256)
    assert_contacts_equal(got, gst, exp, est, "cfg5 TOI")
    assert (est == 5).sum() > 0  # NONCONVERGED limit cycles exist in this sample


def test_mixed_toi_with_spheres(ctx, oracle):
    w = W.cfg2_mixed_pairs(20_000, seed=7, spread=2.5)
    _upload(ctx, w)
    rng = np.random.default_rng(8)
    l1, r1 = w["l_t"].copy(), w["r_t"].copy()
    l1["d"] += (w["r_t"]["d"] - w["l_t"]["d"]) * rng.uniform(0.5, 1.5, (len(l1), 1)).astype(np.float32)
    args = (w["l_shape"], w["r_shape"], w["l_t"], l1, w["r_t"], r1)
    exp, est = oracle.time_of_impact(w["shapes"], None, *args, iter_cap=300, nthreads=8)
    got, gst = ctx.time_of_impact(*args, iter_cap=300)
    assert_contacts_equal(got, gst, exp, est, "mixed TOI")


def test_non_unit_scale_and_zero_scale(ctx, oracle):
    w = W.cfg2_mixed_pairs(20_000, seed=11, spread=1.0)
    rng = np.random.default_rng(12)
    w["l_t"]["scale"] = rng.uniform(0.3, 2.0, len(w["l_t"])).astype(np.float32)
    w["r_t"]["scale"] = rng.uniform(0.3, 2.0, len(w["r_t"])).astype(np.float32)
    w["r_t"]["scale"][::97] = 0.0          # inverse_transform_vector(..).unwrap() panics
    w["l_t"]["scale"][5::193] = 1e-8       # |scale| <= f32::EPSILON also counts as zero
    w["l_t"]["scale"][7::211] = -1.5       # negative scale is legal
    _upload(ctx, w)
    args = (w["l_shape"], w["r_shape"], w["l_t"], w["r_t"])
    exp, est = oracle.intersection(0, w["shapes"], None, *args, nthreads=8)
    got, gst = ctx.intersection(0, *args)
    assert_contacts_equal(got, gst, exp, est, "scaled FullResolution")
    assert (est == 4).sum() > 100
    ed, est = oracle.distance(w["shapes"], None, *args, nthreads=8)
    gd, gst = ctx.distance(*args)
    assert np.array_equal(gst, est)
    assert bits_equal(gd[est == 1], ed[est == 1]).all()


def test_settings_are_honoured(ctx, oracle):
    from collision_rs_b200.abi import Settings as S
    from oracle.binding import Settings as OS
    w = W.cfg2_mixed_pairs(10_000, seed=21)
    _upload(ctx, w)
    args = (w["l_shape"], w["r_shape"], w["l_t"], w["r_t"])
    for (dt, ct, et, mi) in [(1e-3, 1e-3, 1e-2, 7), (1e-6, 1e-6, 1e-5, 1), (1e-6, 1e-6, 1e-5, 0)]:
        exp, est = oracle.intersection(0, w["shapes"], None, *args, settings=OS(dt, ct, et, mi),
                                       nthreads=8)
        got, gst = ctx.intersection(0, *args, settings=S(dt, ct, et, mi))
        assert_contacts_equal(got, gst, exp, est, f"settings {mi}")
        ed, est = oracle.distance(w["shapes"], None, *args, settings=OS(dt, ct, et, mi), nthreads=8)
        gd, gst = ctx.distance(*args, settings=S(dt, ct, et, mi))
        assert np.array_equal(gst, est) and bits_equal(gd[est == 1], ed[est == 1]).all()
    with pytest.raises(Cb2Error) as e:
        ctx.intersection(0, *args, settings=S(1e-6, 1e-6, 1e-5, 251))
    assert e.value.code == -5


def test_more_than_a_hundred_iterations(ctx, oracle):
    """GJK::new_with_settings takes any max_iterations (gjk/mod.rs:71-84); the device accepts up to 250.
    Some curved pairs need more than the crate's default 100 iterations: their polytopes (more than
    104 vertices) outgrow every shared-memory tier and finish in the second-chance kernel over global
    scratch."""
    from collision_rs_b200.abi import Settings as S
    from oracle.binding import Settings as OS
    w = W.cfg2_mixed_pairs(20_000, seed=33, spread=1.0)
    _upload(ctx, w)
    args = (w["l_shape"], w["r_shape"], w["l_t"], w["r_t"])
    for mi in (160, 250):
        exp, est, stats = oracle.intersection(0, w["shapes"], None, *args, settings=OS(1e-6, 1e-6, 1e-5, mi),
                                              nthreads=8, want_stats=True)
        assert (stats[:, 1] > 100).sum() >= 20   # pairs that need more than the default cap
        got, gst = ctx.intersection(0, *args, settings=S(1e-6, 1e-6, 1e-5, mi))
        assert_contacts_equal(got, gst, exp, est, f"max_iterations {mi}")
        ed, est = oracle.distance(w["shapes"], None, *args, settings=OS(1e-6, 1e-6, 1e-5, mi), nthreads=8)
        gd, gst = ctx.distance(*args, settings=S(1e-6, 1e-6, 1e-5, mi))
        assert np.array_equal(gst, est) and bits_equal(gd[est == 1], ed[est == 1]).all()


def test_hazard_vectors(ctx, oracle):
    # Appendix B vectors: the GPU must classify them exactly like the oracle
    import test_oracle_hazards as H
    for c in H.B1:
        shapes = np.concatenate([H.cuboid(*H.dims(c["ld"])), H.cuboid(*H.dims(c["rd"]))])
        ctx.upload_shapes(shapes)
        out, st = ctx.time_of_impact(H.I0, H.I1, H.xf(c["l0"], c["lq"]), H.xf(c["l1"], c["lq"]),
                                     H.xf(c["r0"], c["rq"]), H.xf(c["r1"], c["rq"]), iter_cap=500)
        assert st[0] == 5
    gjk = GJK3.new(ctx)
    with pytest.raises(ReferencePanic) as e:
        gjk.distance(Sphere(float.fromhex("0x1.b38cfap-1")),
                     Decomposed(tuple(float.fromhex(x) for x in
                                      ["0x1.eeb542p-1", "-0x1.39f1bep-1", "0x1.5da646p-1"]),
                                tuple(float.fromhex(x) for x in
                                      ["-0x1.1937eap-1", "0x1.6e74c0p-2", "0x1.620826p-4",
                                       "-0x1.80186cp-1"])),
                     Sphere(float.fromhex("0x1.a94cf4p-1")),
                     Decomposed(tuple(float.fromhex(x) for x in
                                      ["0x1.19c170p-3", "-0x1.45ed10p-2", "0x1.25f46cp+0"]),
                                tuple(float.fromhex(x) for x in
                                      ["0x1.0b81a6p-2", "-0x1.2b66e8p-1", "0x1.6d93c8p-1",
                                       "0x1.21975ep-2"])))
    assert e.value.status == 3


def test_empty_and_tiny_batches(ctx):
    w = W.cfg1_cuboid_pairs(3)
    _upload(ctx, w)
    z32 = np.zeros(0, np.uint32)
    out, st = ctx.intersection(0, z32, z32, w["l_t"][:0], w["r_t"][:0])
    assert len(out) == 0 and len(st) == 0
    out, st = ctx.intersection(0, w["l_shape"][:1], w["r_shape"][:1], w["l_t"][:1], w["r_t"][:1])
    assert len(st) == 1
    with pytest.raises(Cb2Error):
        ctx.intersection(0, np.array([99], np.uint32), w["r_shape"][:1], w["l_t"][:1], w["r_t"][:1])


def test_world_aabbs(ctx, oracle):
    w = W.cfg3_polyhedron_pairs(5_000, n_shapes=128)
    m = W.cfg2_mixed_pairs(5_000, seed=3)
    # one table with all three kinds
    shapes = np.concatenate([w["shapes"], m["shapes"][:5000]])
    ctx.upload_shapes(shapes, w["verts"])
    rng = np.random.default_rng(1)
    idx = rng.integers(0, len(shapes), 20_000).astype(np.uint32)
    t = W.random_xforms(rng, len(idx), 5.0)
    t["scale"] = rng.uniform(0.5, 2.0, len(idx)).astype(np.float32)
    exp = oracle.world_aabbs(shapes, w["verts"], idx, t, nthreads=8)
    got = ctx.world_aabbs(idx, t)
    assert bits_equal(got["min"], exp["min"]).all() and bits_equal(got["max"], exp["max"]).all()


# ---- broad phase ----------------------------------------------------------------------------------------
def _check_sap(ctx, oracle, boxes, axis):
    perm_e, pairs_e, cnt, axis_e = oracle.sap3(boxes, axis=axis)
    perm_g, pairs_g, axis_g = ctx.sap3(boxes, axis=axis)
    assert np.array_equal(perm_g, perm_e)
    assert len(pairs_g) == cnt
    assert np.array_equal(pairs_g, pairs_e)  # same pairs in the same Vec order
    assert axis_g == axis_e


@pytest.mark.parametrize("axis", [0, 1, 2])
def test_sap_random_boxes(ctx, oracle, axis):
    n = 30_000
    w = W.cfg4_aabbs(n, extent=200.0 * (n / 1e6) ** (1 / 3))
    _check_sap(ctx, oracle, w["aabbs"], axis)


def test_sap_ties_signed_zero_and_wide_boxes(ctx, oracle):
    rng = np.random.default_rng(3)
    n = 5_000
    w = W.cfg4_aabbs(n, seed=9, extent=40.0)
    a = w["aabbs"]
    # quantise so that many min (and min+max) keys tie -> exercises stability
    a["min"] = np.round(a["min"])
    a["max"] = np.round(a["max"]) + 1
    a["min"][::7, 0] = -0.0
    a["min"][3::7, 0] = 0.0
    a["min"][::501] = -100.0   # a few boxes spanning the whole scene
    a["max"][::501] = 100.0
    a["max"][10::97] = a["min"][10::97]  # zero-width boxes (strict test rejects touching)
    for axis in (0, 2):
        _check_sap(ctx, oracle, a, axis)


def test_sap_long_segments_of_the_pair_list(ctx, oracle):
    """The pair list is ordered segment by segment (one segment per later box): a box that spans the
    scene collects a long segment — block-sorted up to 4096 pairs, the radix-sort path beyond."""
    for n, stride in ((3_000, 997), (9_000, 2_999), (2_000, 1)):
        a = W.cfg4_aabbs(n, seed=n, extent=40.0)["aabbs"]
        if stride > 1:
            a["min"][::stride] = -100.0
            a["max"][::stride] = 100.0
        else:  # every box overlaps every other: segment k holds k pairs (33 .. 1999: thread and block paths)
            a["min"] -= 50.0
            a["max"] += 50.0
        for axis in (0, 1):
            _check_sap(ctx, oracle, a, axis)
        exp = oracle.brute_force_pairs(a)
        assert np.array_equal(ctx.brute_force(a), exp)


def test_sap_small_n(ctx, oracle):
    w = W.cfg4_aabbs(64, extent=8.0)
    for n in (0, 1, 2, 3, 33):
        perm_g, pairs_g, axis_g = ctx.sap3(w["aabbs"][:n], axis=1)
        perm_e, pairs_e, cnt, axis_e = oracle.sap3(w["aabbs"][:n], axis=1)
        assert np.array_equal(perm_g, perm_e[:n]) and np.array_equal(pairs_g, pairs_e)
        assert axis_g == axis_e == (1 if n <= 1 else 0)


def test_sap_nan_is_rejected_loudly(ctx):
    w = W.cfg4_aabbs(100, extent=10.0)
    w["aabbs"]["min"][5, 0] = np.nan
    with pytest.raises(Cb2Error) as e:
        ctx.sap3(w["aabbs"], axis=0)
    assert e.value.code == -6


def _odd_boxes(seed=21, n=6_000):
    """cfg-4 boxes with NaN extents on the two grid axes and inverted (min > max) extents on every
    axis sprinkled in: the reference tests them literally (a NaN box pairs with nothing, an inverted
    box can still pass the strict test), sweep_prune.rs:100-118, aabb3.rs:246-253."""
    rng = np.random.default_rng(seed)
    a = W.cfg4_aabbs(n, seed=seed, extent=200.0 * (n / 1e6) ** (1 / 3))["aabbs"]
    k = rng.permutation(n)
    a["min"][k[0:40], 1] = np.nan
    a["max"][k[40:80], 2] = np.nan
    a["max"][k[80:100], 1] = np.nan
    a["min"][k[100:120], 2] = -np.nan
    for lo, hi, ax in ((120, 170, 1), (170, 220, 2), (220, 270, 0)):  # inverted extents
        i = k[lo:hi]
        mn, mx = a["min"][i, ax].copy(), a["max"][i, ax].copy()
        a["min"][i, ax], a["max"][i, ax] = mx, mn
    i = k[270:290]  # inverted AND wide: still passes the strict test against many boxes
    a["min"][i, 1], a["max"][i, 1] = 1e3, -1e3
    a["min"][k[290:300], 2] = np.inf
    a["max"][k[300:310], 1] = -np.inf
    return a


def test_sap_nan_and_inverted_extents_off_the_sweep_axis(ctx, oracle):
    a = _odd_boxes()
    _check_sap(ctx, oracle, a, 0)
    pairs = oracle.sap3(a, axis=0)[1]
    assert len(pairs) > 1000
    # the inverted-and-wide boxes do produce pairs in the reference: the test is not vacuous
    perm = oracle.sap3(a, axis=0)[0]
    odd = np.flatnonzero(a["min"][perm, 1] > a["max"][perm, 1])
    assert np.isin(pairs, odd).any()


def test_brute_force_accepts_nan_boxes(ctx, oracle):
    a = _odd_boxes(seed=4, n=3_000)
    a["min"][7, 0] = np.nan   # BruteForce never sorts: NaN anywhere is just "intersects nothing"
    a["max"][11, 0] = np.nan
    a["min"][13, 0] = -np.nan
    exp = oracle.brute_force_pairs(a)
    got = ctx.brute_force(a)
    assert len(exp) > 500 and np.array_equal(got, exp)


# ---- support cells (csrc/cb2_cells.cuh) against the plain vertex scan ---------------------------
def _ctx_with_env(**env):
    """A second context created under environment toggles the C ABI reads in cb2_create."""
    import os
    from collision_rs_b200 import host
    old = {k: os.environ.get(k) for k in env}
    os.environ.update({k: str(v) for k, v in env.items()})
    try:
        return host.Context(0)
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def test_support_cells_match_full_scan_on_a_million_pairs(ctx):
    """The cube-map candidate lists must return the very vertex the O(V) scan returns: compare
    the whole pipeline with CB2_NO_CELLS=1 (every polyhedron support = full scan) on 1M pairs
    (~25M support queries; the oracle can only afford a few thousand pairs)."""
    w = W.cfg3_polyhedron_pairs(1_000_000, seed=4321)
    args = (w["l_shape"], w["r_shape"], w["l_t"], w["r_t"])
    ctx.upload_shapes(w["shapes"], w["verts"])
    got, gst = ctx.intersection(0, *args)
    gd, gdst = ctx.distance(*args)
    plain = _ctx_with_env(CB2_NO_CELLS=1)
    try:
        plain.upload_shapes(w["shapes"], w["verts"])
        exp, est = plain.intersection(0, *args)
        ed, edst = plain.distance(*args)
    finally:
        plain.close()
    assert_contacts_equal(got, gst, exp, est, "cells vs full scan")
    assert np.array_equal(gdst, edst)
    assert bits_equal(gd[edst == 1], ed[edst == 1]).all()


def test_intersection_bodies_matches_the_per_pair_call(ctx):
    """cb2_gjk3_intersection_bodies (host buffers, one body table, 8 bytes per pair streamed) must give
    what cb2_gjk3_intersection gives on the expanded per-pair arrays, for both strategies and across
    several pipeline chunks; a body index outside the table fails the call."""
    w = W.cfg3_polyhedron_pairs(16, n_shapes=256)
    rng = np.random.default_rng(99)
    nb, n = 50_000, 300_000
    body_shape = rng.integers(0, 256, nb).astype(np.uint32)
    body_t = W.random_xforms(rng, nb, 0.75)
    pairs = rng.integers(0, nb, (n, 2)).astype(np.uint32)
    small = _ctx_with_env(CB2_CHUNK_PAIRS=40_000)  # 300 k pairs = many chunks, ramps included
    try:
        small.upload_shapes(w["shapes"], w["verts"])
        for strategy in (0, 1):
            got, gst = small.intersection_bodies(strategy, body_shape, body_t, pairs)
            exp, est = small.intersection(strategy, body_shape[pairs[:, 0]], body_shape[pairs[:, 1]],
                                          body_t[pairs[:, 0]], body_t[pairs[:, 1]])
            assert (est == 1).sum() > n // 2
            assert_contacts_equal(got, gst, exp, est, f"bodies vs per pair, strategy {strategy}")
        bad = pairs.copy()
        bad[1234, 1] = nb
        with pytest.raises(Cb2Error):
            small.intersection_bodies(0, body_shape, body_t, bad)
        # and the context is usable afterwards
        got, gst = small.intersection_bodies(1, body_shape, body_t, pairs[:1000])
        exp, est = small.intersection(1, body_shape[pairs[:1000, 0]], body_shape[pairs[:1000, 1]],
                                      body_t[pairs[:1000, 0]], body_t[pairs[:1000, 1]])
        assert_contacts_equal(got, gst, exp, est, "after a failed call")
    finally:
        small.close()


def test_split_support_on_tied_candidate_lists(ctx):
    """The EPA groups split a support cell's candidate list over the lanes of a side (chunks of four,
    greater dot wins, equal dots go to the lower part).  Lattice hulls make exact ties and long
    lists — narrow (u8) records with more than eight candidates and wide (u16) ones — so the split
    scan must still return the first strictly greatest vertex: compare FullResolution contacts with
    CB2_NO_CELLS=1 (every support a full scan on one lane, no split) on 60 k pairs."""
    rng = np.random.default_rng(2718)
    verts, shapes = [], []
    for k, cnt in enumerate((60, 120, 200, 250, 300, 400, 600, 90, 150, 36)):
        span = 2 + k % 3  # coarse lattices: many exactly equal dots along the axes and diagonals
        p = np.unique(rng.integers(-span, span + 1, (cnt * 3, 3)), axis=0).astype(np.float32)
        p = p[rng.permutation(len(p))][:cnt] * np.float32(0.25 + 0.1 * k)
        s = np.zeros(1, host_abi.SHAPE_DT)
        s["kind"] = host_abi.POLYHEDRON
        s["vtx_off"] = sum(len(v) for v in verts)
        s["vtx_cnt"] = len(p)
        verts.append(p)
        shapes.append(s)
    shapes = np.concatenate(shapes)
    pv = np.concatenate(verts)
    n = 60_000
    l = rng.integers(0, len(shapes), n).astype(np.uint32)
    r = rng.integers(0, len(shapes), n).astype(np.uint32)
    lt, rt = W.random_xforms(rng, n, 0.8), W.random_xforms(rng, n, 0.8)
    # axis-aligned poses keep the ties exact in world space as well
    lt["q"][: n // 2] = (1.0, 0.0, 0.0, 0.0)
    rt["q"][: n // 2] = (1.0, 0.0, 0.0, 0.0)
    ctx.upload_shapes(shapes, pv)
    got, gst = ctx.intersection(0, l, r, lt, rt)
    plain = _ctx_with_env(CB2_NO_CELLS=1)
    try:
        plain.upload_shapes(shapes, pv)
        exp, est = plain.intersection(0, l, r, lt, rt)
    finally:
        plain.close()
    assert (est == 1).sum() > n // 10
    assert_contacts_equal(got, gst, exp, est, "split support vs full scan")


def test_degenerate_polyhedra(ctx, oracle):
    """Hulls the cell construction must not trip over: interior points, duplicated vertices,
    coplanar and collinear point sets, a single point, tiny and huge coordinates."""
    rng = np.random.default_rng(77)
    verts, shapes = [], []

    def add(p):
        s = np.zeros(1, host_abi.SHAPE_DT)
        s["kind"] = host_abi.POLYHEDRON
        s["vtx_off"] = sum(len(v) for v in verts)
        s["vtx_cnt"] = len(p)
        verts.append(np.asarray(p, np.float32))
        shapes.append(s)

    sph = rng.standard_normal((60, 3))
    sph /= np.linalg.norm(sph, axis=1, keepdims=True)
    add(np.concatenate([sph, 0.3 * sph[:20], np.zeros((3, 3))]))            # interior points
    add(np.concatenate([sph[:30], sph[:30]]))                                # every vertex twice
    add(np.c_[rng.uniform(-1, 1, (40, 2)), np.zeros(40)])                    # coplanar
    add(np.c_[np.linspace(-1, 1, 20), np.zeros(20), np.zeros(20)])           # collinear
    add(np.zeros((12, 3)))                                                   # all at the origin
    add(sph[:48] * 1e-12)                                                    # tiny
    add(sph[:48] * 1e12)                                                     # huge
    cube = np.array([[x, y, z] for x in (-1, 1) for y in (-1, 1) for z in (-1, 1)], float)
    add(np.concatenate([cube, cube * 0.999999, rng.uniform(-1, 1, (30, 3))]))  # near-ties
    add(rng.integers(-2, 3, (300, 3)).astype(float))                         # lattice: exact ties, u16
    shapes = np.concatenate(shapes)
    pv = np.concatenate(verts)
    n = 3_000
    l = rng.integers(0, len(shapes), n).astype(np.uint32)
    r = rng.integers(0, len(shapes), n).astype(np.uint32)
    lt, rt = W.random_xforms(rng, n, 0.6), W.random_xforms(rng, n, 0.6)
    ctx.upload_shapes(shapes, pv)
    exp, est = oracle.intersection(1, shapes, pv, l, r, lt, rt, nthreads=8)
    got, gst = ctx.intersection(1, l, r, lt, rt)
    assert_contacts_equal(got, gst, exp, est, "degenerate hulls CollisionOnly")
    # FullResolution: flat / collinear hulls make EPA grow non-manifold polytopes without bound
    # (the reference's Vec reaches 32k faces here).  The device keeps at most 65536 faces and 16384
    # horizon edges in its second-chance scratch; past that it must say CB2_SCRATCH_OVERFLOW, never a
    # wrong contact.
    exp, est, stats = oracle.intersection(0, shapes, pv, l, r, lt, rt, nthreads=8, want_stats=True)
    got, gst = ctx.intersection(0, l, r, lt, rt)
    fits = (stats[:, 2] <= 65536) & ((stats[:, 3] >> 16) <= 16384)
    assert (gst[~fits] == 7).all()
    assert fits.sum() > 0.99 * n
    assert_contacts_equal(got[fits], gst[fits], exp[fits], est[fits], "degenerate hulls FullResolution")
    ed, est = oracle.distance(shapes, pv, l, r, lt, rt, nthreads=8)
    gd, gst = ctx.distance(l, r, lt, rt)
    assert np.array_equal(gst, est)
    assert bits_equal(gd[est == 1], ed[est == 1]).all()


def test_grid_sap_matches_tiled_sweep_on_a_million_boxes(ctx):
    """The grid-accelerated sweep must return the byte-identical permutation and pair list as the
    plain windowed sweep (CB2_SAP_LEGACY=1), at cfg 4's full size and with a heavy-tailed mix of
    box sizes (so the large-box path and multi-cell boxes are exercised)."""
    import os
    rng = np.random.default_rng(5)
    cases = [W.cfg4_aabbs(1_000_000)["aabbs"]]
    n = 200_000
    c = rng.uniform(0, 100, (n, 3)).astype(np.float32)
    h = (0.2 * rng.pareto(1.5, (n, 3)) + 0.05).astype(np.float32)  # a few boxes span the domain
    b = np.zeros(n, host_abi.AABB_DT)
    b["min"], b["max"] = c - h, c + h
    cases.append(b)
    for boxes in cases:
        for axis in (0, 2):
            perm, pairs, ax = ctx.sap3(boxes, axis=axis)
            os.environ["CB2_SAP_LEGACY"] = "1"
            try:
                perm_e, pairs_e, ax_e = ctx.sap3(boxes, axis=axis)
            finally:
                del os.environ["CB2_SAP_LEGACY"]
            assert np.array_equal(perm, perm_e) and ax == ax_e
            assert pairs.shape == pairs_e.shape and np.array_equal(pairs, pairs_e)


def test_sap_shards_concatenate_to_the_full_pair_list(ctx, oracle):
    """cb2_sap3_find_collider_pairs_shard_dev: for every world size the shards, each
    computed on its own (here: the same) GPU, concatenated in rank order are byte-identical to the
    single-GPU list; heavy-tailed box sizes exercise the large-box path too."""
    import torch
    from collision_rs_b200.sharding import pairs_of_shard
    rng = np.random.default_rng(6)
    n = 150_000
    c = rng.uniform(0, 90, (n, 3)).astype(np.float32)
    h = (0.2 * rng.pareto(1.5, (n, 3)) + 0.05).astype(np.float32)
    heavy = np.zeros(n, host_abi.AABB_DT)
    heavy["min"], heavy["max"] = c - h, c + h
    dev = torch.device("cuda:0")
    s = torch.cuda.current_stream().cuda_stream
    for boxes in (W.cfg4_aabbs(n, extent=200.0 * (n / 1e6) ** (1 / 3))["aabbs"], heavy):
        perm_e, pairs_e, _ = ctx.sap3(boxes, axis=0)
        d_boxes = torch.from_numpy(boxes.view(np.uint8).reshape(-1).copy()).to(dev)
        perm = torch.zeros(n, dtype=torch.int32, device=dev)
        cap = len(pairs_e) + 16
        out = torch.zeros(2 * cap, dtype=torch.int32, device=dev)
        for world in (1, 2, 3, 8):
            parts = []
            for r in range(world):
                rc, cnt = ctx.sap3_shard_dev(d_boxes.data_ptr(), n, 0, r, world, perm.data_ptr(),
                                             out.data_ptr(), cap, s)
                assert rc == 0
                torch.cuda.synchronize()
                got = out[:2 * cnt].cpu().numpy().view(np.uint32).reshape(-1, 2)
                assert np.array_equal(got, pairs_of_shard(pairs_e, n, r, world))
                assert np.array_equal(perm.cpu().numpy().view(np.uint32), perm_e)
                parts.append(got)
            assert np.array_equal(np.concatenate(parts), pairs_e)
    # a shard that does not fit reports its own required capacity
    rc, cnt = ctx.sap3_shard_dev(d_boxes.data_ptr(), n, 0, 0, 2, perm.data_ptr(), out.data_ptr(), 8, s)
    assert rc == -4 and cnt == len(pairs_of_shard(pairs_e, n, 0, 2))


def test_sap_unordered_is_the_same_pair_set(ctx):
    import torch
    n = 200_000
    boxes = W.cfg4_aabbs(n, extent=200.0 * (n / 1e6) ** (1 / 3))["aabbs"]
    perm_e, pairs_e, _ = ctx.sap3(boxes, axis=1)
    dev = torch.device("cuda:0")
    d_boxes = torch.from_numpy(boxes.view(np.uint8).reshape(-1).copy()).to(dev)
    perm = torch.zeros(n, dtype=torch.int32, device=dev)
    cap = len(pairs_e) + 8
    out = torch.zeros(2 * cap, dtype=torch.int32, device=dev)
    rc, cnt = ctx.sap3_unordered_dev(d_boxes.data_ptr(), n, 1, perm.data_ptr(), out.data_ptr(), cap,
                                     torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert rc == 0 and cnt == len(pairs_e)
    got = out[:2 * cnt].cpu().numpy().view(np.uint32).reshape(-1, 2)
    assert np.array_equal(perm.cpu().numpy().view(np.uint32), perm_e)
    key = lambda p: (p[:, 1].astype(np.uint64) << np.uint64(32)) | p[:, 0].astype(np.uint64)
    assert np.array_equal(np.sort(key(got)), key(pairs_e))   # the reference order IS ascending in this key


# ---- composite shapes: *_complex (gjk/mod.rs:412-588) ----------------------------------------------
def test_complex_callers_match_the_oracle(ctx, oracle):
    from helpers_complex import composite_workload
    cw = composite_workload(20_000, seed=11)
    ctx.upload_shapes(cw["shapes"], cw["verts"])
    ctx.upload_composites(cw["comp_off"], cw["comp_shape"], cw["comp_local"])
    oargs = (cw["shapes"], cw["verts"], cw["comp_off"], cw["comp_shape"], cw["comp_local"],
             cw["l_comp"], cw["r_comp"], cw["l_t0"], cw["r_t0"], cw["l_t1"], cw["r_t1"])
    for strategy in (0, 1):
        exp, est = oracle.complex(0, strategy, *oargs, nthreads=8)
        got, gst = ctx.intersection_complex(strategy, cw["l_comp"], cw["r_comp"], cw["l_t0"], cw["r_t0"])
        assert_contacts_equal(got, gst, exp, est, f"intersection_complex strategy {strategy}")
        assert (est == 1).sum() > 1000
        exp, est = oracle.complex(2, strategy, *oargs, nthreads=8)
        got, gst = ctx.time_of_impact_complex(strategy, cw["l_comp"], cw["r_comp"], cw["l_t0"],
                                              cw["l_t1"], cw["r_t0"], cw["r_t1"])
        assert_contacts_equal(got, gst, exp, est, f"toi_complex strategy {strategy}")
        assert (est == 1).sum() > 1000
    ed, est = oracle.complex(1, 0, *oargs, nthreads=8)
    gd, gst = ctx.distance_complex(cw["l_comp"], cw["r_comp"], cw["l_t0"], cw["r_t0"])
    assert np.array_equal(gst, est)
    assert bits_equal(gd[est == 1], ed[est == 1]).all() and (est == 1).sum() > 500


def test_complex_empty_composites_and_errors(ctx, oracle):
    shapes = W.cuboid_shapes(np.random.default_rng(0), 3)
    ctx.upload_shapes(shapes)
    ident = np.zeros(3, host_abi.XFORM_DT)
    ident["scale"] = 1
    ident["q"][:, 0] = 1
    # composite 0 is empty, composite 1 has one part, composite 2 has two
    ctx.upload_composites(np.array([0, 0, 1, 3], np.uint32), np.array([0, 1, 2], np.uint32), ident)
    t = ident[:3].copy()
    l = np.array([0, 1, 0], np.uint32)
    r = np.array([1, 2, 0], np.uint32)
    out, st = ctx.intersection_complex(0, l, r, t, t)
    assert list(st) == [0, 1, 0]            # an empty side yields None; coincident cuboids collide
    d, st = ctx.distance_complex(l, r, t, t)
    # empty -> None; two coincident cuboids: whatever the reference does (it panics in
    # get_closest_point_on_face), the oracle says
    _, est = oracle.complex(1, 0, shapes, None, np.array([0, 0, 1, 3], np.uint32),
                            np.array([0, 1, 2], np.uint32), ident, l, r, t, t)
    assert st[0] == 0 and st[2] == 0 and np.array_equal(st, est)
    with pytest.raises(Cb2Error):
        ctx.intersection_complex(0, np.array([7], np.uint32), np.array([0], np.uint32), t[:1], t[:1])
    # a new shape table drops the composite table (its part indices refer to the old one)
    ctx.upload_shapes(shapes[:1])
    with pytest.raises(Cb2Error):
        ctx.intersection_complex(0, l, r, t, t)


def test_every_primitive3_variant(ctx, oracle):
    """Particle, Quad, Cylinder, Capsule and Cube next to Sphere / Cuboid / ConvexPolyhedron
    (primitive3.rs:153-176): intersection (both strategies), distance, time of impact and world
    bounds, bit for bit against the oracle; zero scale panics only where the reference's
    support_point inverts the transform (a Particle never does)."""
    rng = np.random.default_rng(99)
    ps, pv = W.polyhedron_table(16, seed=3, sizes=(12, 40))
    shapes = np.concatenate([W.all_primitive3_shapes(rng, 140), ps])
    n = 60_000
    l = rng.integers(0, len(shapes), n).astype(np.uint32)
    r = rng.integers(0, len(shapes), n).astype(np.uint32)
    lt, rt = W.random_xforms(rng, n, 0.9), W.random_xforms(rng, n, 0.9)
    lt["scale"][::9] = rng.uniform(0.5, 1.5, len(lt["scale"][::9])).astype(np.float32)
    rt["scale"][::211] = 0.0
    lt["scale"][::307] = 0.0
    ctx.upload_shapes(shapes, pv)
    exp, est = oracle.intersection(1, shapes, pv, l, r, lt, rt, nthreads=8)
    got, gst = ctx.intersection(1, l, r, lt, rt)
    assert_contacts_equal(got, gst, exp, est, "Primitive3 mix, CollisionOnly")
    # FullResolution: a Particle inside a Cylinder makes EPA grow non-manifold polytopes without
    # bound (the reference's Vec reaches 533k faces and 100 s for ONE pair), so those pairs are
    # checked on a handful only.  Past the 65536 faces / 16384 horizon edges of the second-chance
    # scratch the device must answer CB2_SCRATCH_OVERFLOW, never a wrong contact; every other pair
    # of the mix (one of them reaches 20562 faces) must be computed.
    kk = shapes["kind"]
    slow = ((kk[l] == 3) & (kk[r] == 5)) | ((kk[l] == 5) & (kk[r] == 3))
    few = np.nonzero(slow)[0][:12]
    for sel in (np.nonzero(~slow)[0], few):
        e0, s0, stats = oracle.intersection(0, shapes, pv, l[sel], r[sel], lt[sel], rt[sel],
                                            nthreads=8, want_stats=True)
        g0, gs0 = ctx.intersection(0, l[sel], r[sel], lt[sel], rt[sel])
        fits = (stats[:, 2] <= 65536) & ((stats[:, 3] >> 16) <= 16384)
        assert (gs0[~fits] == 7).all() and fits.sum() > 0.5 * len(sel)
        if sel is not few:
            assert fits.all() and not (gs0 == 7).any()
        assert_contacts_equal(g0[fits], gs0[fits], e0[fits], s0[fits], "Primitive3 mix, FullResolution")
    kinds = shapes["kind"]
    particle_zero = ((kinds[l] == 3) & (lt["scale"] == 0)) | ((kinds[r] == 3) & (rt["scale"] == 0))
    assert particle_zero.sum() > 5 and (est == 4).sum() > 50
    # a zero-scale Particle alone must not panic
    alone = particle_zero & ~(((kinds[l] != 3) & (lt["scale"] == 0)) | ((kinds[r] != 3) & (rt["scale"] == 0)))
    assert alone.sum() > 0 and (est[alone] != 4).all()
    ed, est = oracle.distance(shapes, pv, l, r, lt, rt, nthreads=8)
    gd, gst = ctx.distance(l, r, lt, rt)
    assert np.array_equal(gst, est)
    assert bits_equal(gd[est == 1], ed[est == 1]).all()
    m = 20_000
    lt1, rt1 = lt[:m].copy(), rt[:m].copy()
    lt1["d"] += rng.uniform(-1, 1, (m, 3)).astype(np.float32)
    rt1["d"] += ((lt["d"][:m] - rt["d"][:m]) * rng.uniform(0.5, 1.5, (m, 1))).astype(np.float32)
    exp, est = oracle.time_of_impact(shapes, pv, l[:m], r[:m], lt[:m], lt1, rt[:m], rt1, nthreads=8)
    got, gst = ctx.time_of_impact(l[:m], r[:m], lt[:m], lt1, rt[:m], rt1)
    assert_contacts_equal(got, gst, exp, est, "Primitive3 mix, time of impact")
    eb = oracle.world_aabbs(shapes, pv, l, lt, nthreads=8)
    gb = ctx.world_aabbs(l, lt)
    assert gb.tobytes() == eb.tobytes()


# ---- HalfEdge polyhedra: ConvexPolyhedron::new_with_faces, hill-climb support ----------------------
def _he_table(n_shapes, seed, sizes=(8, 20, 60, 150)):
    """Random convex hulls given BOTH ways: shape 2k is HalfEdge mode (kind 8, oriented hull
    triangles), shape 2k+1 the same vertices in VertexOnly mode."""
    from test_oracle_golden import oriented_hull_faces
    from helpers import polyhedron, polyhedron_he
    rng = np.random.default_rng(seed)
    shapes, verts, faces = [], [], []
    voff = foff = 0
    for k in range(n_shapes):
        v = rng.standard_normal((sizes[k % len(sizes)], 3))
        v /= np.linalg.norm(v, axis=1, keepdims=True)
        v = (v * rng.uniform(0.5, 1.0, 3)).astype(np.float32)
        f = oriented_hull_faces(v)
        shapes.append(polyhedron_he(voff, len(v), foff, len(f)))
        shapes.append(polyhedron(voff, len(v)))
        verts.append(v)
        faces.append(f)
        voff += len(v)
        foff += len(f)
    return np.concatenate(shapes), np.concatenate(verts), np.concatenate(faces)


def test_half_edge_polyhedra(ctx, oracle):
    shapes, verts, faces = _he_table(24, seed=21)
    prim = W.mixed_shapes(np.random.default_rng(2), 8)
    shapes = np.concatenate([shapes, prim])
    oracle.build_half_edges(shapes, verts, faces)
    ctx.upload_shapes(shapes, verts, faces)
    rng = np.random.default_rng(22)
    n = 40_000
    l = rng.integers(0, len(shapes), n).astype(np.uint32)
    r = rng.integers(0, len(shapes), n).astype(np.uint32)
    lt, rt = W.random_xforms(rng, n, 0.7), W.random_xforms(rng, n, 0.7)
    lt["scale"][::6] = rng.uniform(0.5, 1.5, len(lt["scale"][::6])).astype(np.float32)
    for strategy in (0, 1):
        exp, est = oracle.intersection(strategy, shapes, verts, l, r, lt, rt, nthreads=8)
        got, gst = ctx.intersection(strategy, l, r, lt, rt)
        assert_contacts_equal(got, gst, exp, est, f"HalfEdge polyhedra, strategy {strategy}")
    assert (est == 1).sum() > 5000
    ed, est = oracle.distance(shapes, verts, l, r, lt, rt, nthreads=8)
    gd, gst = ctx.distance(l, r, lt, rt)
    assert np.array_equal(gst, est) and bits_equal(gd[est == 1], ed[est == 1]).all()
    eb = oracle.world_aabbs(shapes, verts, l, lt, nthreads=8)
    assert ctx.world_aabbs(l, lt).tobytes() == eb.tobytes()


def test_half_edge_nan_direction_is_a_hang(ctx, oracle):
    """A NaN support direction makes the reference's hill climb spin forever (best_dot != previous
    is always true for NaN): both sides must report status 5."""
    shapes, verts, faces = _he_table(2, seed=5, sizes=(20,))
    oracle.build_half_edges(shapes, verts, faces)
    ctx.upload_shapes(shapes, verts, faces)
    t = np.zeros(1, host_abi.XFORM_DT)
    t["scale"] = 1
    t["q"][:, 0] = 1
    tn = t.copy()
    tn["d"][0, 0] = np.nan  # NaN displacement -> NaN seed direction
    l = np.array([0], np.uint32)
    r = np.array([2], np.uint32)
    exp, est = oracle.intersection(0, shapes, verts, l, r, t, tn)
    got, gst = ctx.intersection(0, l, r, t, tn)
    assert est[0] == 5 and gst[0] == 5
    ed, est = oracle.distance(shapes, verts, l, r, t, tn)
    gd, gst = ctx.distance(l, r, t, tn)
    assert est[0] == gst[0]


def test_half_edge_upload_errors(ctx):
    from helpers import polyhedron_he
    pts = np.array([[0, 0, 0], [1, 0, 0], [2, 0, 0], [0, 1, 0]], np.float32)
    faces = np.array([(0, 1, 2), (0, 1, 3), (1, 2, 3), (0, 2, 3)], np.uint32)  # collinear face
    with pytest.raises(Cb2Error):
        ctx.upload_shapes(polyhedron_he(0, 4, 0, 4), pts, faces)
    with pytest.raises(Cb2Error):
        ctx.upload_shapes(polyhedron_he(0, 4, 0, 4), pts)  # HalfEdge shape without a face table


def test_brute_force_broad_phase(ctx, oracle):
    """BruteForce::find_collider_pairs (brute_force.rs:20-39): the O(n^2) double loop's pair list,
    in the caller's indices and order, out of the grid machinery."""
    rng = np.random.default_rng(17)
    for n, extent in ((1, 5.0), (2, 1.0), (3000, 30.0)):
        c = rng.uniform(0, extent, (n, 3)).astype(np.float32)
        h = rng.uniform(0.2, 1.5, (n, 3)).astype(np.float32)
        b = np.zeros(n, host_abi.AABB_DT)
        b["min"], b["max"] = c - h, c + h
        b["min"][::7] = np.round(b["min"][::7])  # ties and touching boxes
        exp = oracle.brute_force_pairs(b)
        got = ctx.brute_force(b)
        assert got.shape == exp.shape and np.array_equal(got, exp)
    # at scale: the same pair set as SweepAndPrune3 mapped back through its permutation
    boxes = W.cfg4_aabbs(300_000, extent=200.0 * 0.3 ** (1 / 3))["aabbs"]
    perm, pairs, _ = ctx.sap3(boxes)
    orig = np.sort(perm[pairs.astype(np.int64)], axis=1)
    key = np.lexsort((orig[:, 1], orig[:, 0]))
    got = ctx.brute_force(boxes)
    assert np.array_equal(got, orig[key])


def test_sweep_and_prune2(ctx, oracle):
    """SweepAndPrune2 (the alias the reference's own tests use, sweep_prune.rs:317-361): the four
    reference cases through cb2_sap2_find_collider_pairs, then random 2-D boxes against the 3-D
    oracle on the lifted boxes."""
    def run(boxes, axis=0):
        perm, pairs, ax = ctx.sap2(np.asarray(boxes, np.float32), axis)
        return [tuple(int(x) for x in p) for p in pairs], perm, ax

    assert run([(8, 8, 10, 11), (12, 13, 18, 18)])[0] == []
    assert run([(12, 13, 18, 18), (8, 8, 10, 11)])[0] == []
    assert run([(8, 8, 10, 11), (9, 10, 18, 18)])[0] == [(0, 1)]
    pairs, perm, ax = run([(9, 10, 18, 18), (8, 8, 10, 11)])
    assert pairs == [(0, 1)] and ax == 0 and list(perm) == [1, 0]
    rng = np.random.default_rng(4)
    n = 20_000
    c = rng.uniform(0, 150, (n, 2)).astype(np.float32)
    h = rng.uniform(0.3, 1.5, (n, 2)).astype(np.float32)
    b2 = np.concatenate([c - h, c + h], axis=1)
    b3 = np.zeros(n, host_abi.AABB_DT)
    b3["min"][:, :2], b3["max"][:, :2] = c - h, c + h
    b3["max"][:, 2] = 1.0
    for axis in (0, 1):
        perm, pairs, ax = ctx.sap2(b2, axis)
        perm_e, pairs_e, cnt, ax_e = oracle.sap3(b3, axis=axis)
        assert np.array_equal(perm, perm_e) and np.array_equal(pairs, pairs_e) and ax == ax_e
    with pytest.raises(Cb2Error):
        ctx.sap2(b2, axis=2)


# ---- GJK::intersect / GJK::get_contact_manifold on their own (gjk/mod.rs:101-146, 326-344) -----------
def test_intersect_returns_the_reference_simplex(ctx, oracle):
    w = W.cfg3_polyhedron_pairs(30_000, n_shapes=256, seed=12)
    ctx.upload_shapes(w["shapes"], w["verts"])
    args = (w["l_shape"], w["r_shape"], w["l_t"], w["r_t"])
    sx, st = ctx.intersect(*args)
    sx_e, st_e = oracle.intersect_batch(w["shapes"], w["verts"], *args, nthreads=8)
    assert np.array_equal(st, st_e) and 0.5 < (st == 1).mean() < 0.95
    for f in ("v", "sup_a", "sup_b"):   # every point of every tetrahedron, in the reference's Vec order
        assert bits_equal(sx[f], sx_e[f]).all(), f
    # a mixed table: every kind the oracle has a support function for
    w = W.cfg2_mixed_pairs(20_000, spread=1.0)
    ctx.upload_shapes(w["shapes"], w["verts"])
    args = (w["l_shape"], w["r_shape"], w["l_t"], w["r_t"])
    sx, st = ctx.intersect(*args)
    sx_e, st_e = oracle.intersect_batch(w["shapes"], w["verts"], *args, nthreads=8)
    assert np.array_equal(st, st_e)
    for f in ("v", "sup_a", "sup_b"):
        assert bits_equal(sx[f], sx_e[f]).all(), f


def test_get_contact_manifold_completes_intersect(ctx, oracle):
    """intersect + get_contact_manifold == intersection(FullResolution), bit for bit; a simplex of
    fewer than four points gives None (epa3d.rs:38-40)."""
    for w in (W.cfg3_polyhedron_pairs(30_000, n_shapes=256, seed=13), W.cfg2_mixed_pairs(8_000, spread=1.0)):
        ctx.upload_shapes(w["shapes"], w["verts"])
        args = (w["l_shape"], w["r_shape"], w["l_t"], w["r_t"])
        sx, st = ctx.intersect(*args)
        held = np.where(st == 1, 4, 0).astype(np.uint32)
        held[::11] = np.minimum(held[::11], 3)   # some callers hold a smaller simplex: None
        out, st2 = ctx.get_contact_manifold(sx, *args, simplex_len=held)
        exp, est = oracle.manifold_batch(sx, w["shapes"], w["verts"], *args, simplex_len=held, nthreads=8)
        assert np.array_equal(st2, est)
        for f in ("strategy", "normal", "depth", "point", "toi"):
            assert bits_equal(out[f], exp[f]).all(), f
        full, fst = ctx.intersection(0, *args)
        keep = held == np.where(st == 1, 4, 0)
        assert np.array_equal(st2[keep], fst[keep])
        assert np.array_equal(out[keep].view(np.uint8), full[keep].view(np.uint8))


def test_gjk3_intersect_and_manifold_wrappers(ctx):
    # gjk/mod.rs:700-725 (test_gjk_3d_hit) through the reference-shaped single-pair methods
    gjk = GJK3.new(ctx)
    left, right = Cuboid(10.0, 10.0, 10.0), Cuboid(10.0, 10.0, 10.0)
    lt, rt = D(15, 0, 0), D(7, 2, 0)
    simplex = gjk.intersect(left, lt, right, rt)
    assert simplex is not None and len(simplex) == 4
    contact = gjk.get_contact_manifold(simplex, left, lt, right, rt)
    whole = gjk.intersection(CollisionStrategy.FullResolution, left, lt, right, rt)
    assert contact is not None and contact.penetration_depth == whole.penetration_depth
    assert tuple(contact.normal) == tuple(whole.normal)
    assert gjk.get_contact_manifold(simplex[:3], left, lt, right, rt) is None
    assert gjk.intersect(left, lt, right, D(-15, 0, 0)) is None


def test_shape_index_out_of_range(ctx, oracle):
    """A Rust slice index would panic: host-buffer calls fail as a whole (CB2_ERR_ARG), device-buffer
    calls mark the pair CB2_BAD_INDEX and read nothing through the index."""
    import torch
    w = W.cfg1_cuboid_pairs(4_096)
    ctx.upload_shapes(w["shapes"], None)
    l_shape = w["l_shape"].copy()
    l_shape[[5, 777]] = len(w["shapes"]) + 3
    r_shape = w["r_shape"].copy()
    r_shape[2048] = 0xFFFFFFFF
    for call in (lambda: ctx.intersection(0, l_shape, r_shape, w["l_t"], w["r_t"]),
                 lambda: ctx.distance(l_shape, r_shape, w["l_t"], w["r_t"]),
                 lambda: ctx.intersect(l_shape, r_shape, w["l_t"], w["r_t"])):
        with pytest.raises(Cb2Error) as e:
            call()
        assert e.value.code == -3
    got, gst = ctx.intersection(0, w["l_shape"], w["r_shape"], w["l_t"], w["r_t"])  # the ctx recovers
    exp, est = oracle.intersection(0, w["shapes"], None, w["l_shape"], w["r_shape"], w["l_t"], w["r_t"], nthreads=4)
    assert np.array_equal(gst, est)
    dev = torch.device("cuda", 0)
    d = {k: torch.from_numpy(v.view(np.uint8).reshape(-1).copy()).to(dev)
         for k, v in (("l", l_shape), ("r", r_shape), ("lt", w["l_t"]), ("rt", w["r_t"]))}
    n = len(l_shape)
    d_out = torch.zeros(n * 36, dtype=torch.uint8, device=dev)
    d_st = torch.zeros(n, dtype=torch.uint8, device=dev)
    ctx.intersection_dev(0, d["l"].data_ptr(), d["r"].data_ptr(), d["lt"].data_ptr(), d["rt"].data_ptr(), n,
                         d_out.data_ptr(), d_st.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    st = d_st.cpu().numpy()
    bad = np.zeros(n, bool)
    bad[[5, 777, 2048]] = True
    assert (st[bad] == 11).all() and np.array_equal(st[~bad], est[~bad])


def test_multi_context_shards_one_call_over_every_gpu(oracle):
    """cb2_create_multi: one host-buffer call, every visible GPU takes a contiguous slice (with one
    GPU the same code runs with a single slice)."""
    import torch
    devices = list(range(min(torch.cuda.device_count(), 8)))
    m = MultiContext(devices)
    assert m.device_count == len(devices)
    w = W.cfg3_polyhedron_pairs(40_003, n_shapes=256, seed=21)   # not a multiple of the device count
    m.upload_shapes(w["shapes"], w["verts"])
    args = (w["l_shape"], w["r_shape"], w["l_t"], w["r_t"])
    got, gst = m.intersection(0, *args)
    exp, est = oracle.intersection(0, w["shapes"], w["verts"], *args, nthreads=8)
    assert np.array_equal(gst, est)
    for f in ("strategy", "normal", "depth", "point", "toi"):
        assert bits_equal(got[f], exp[f]).all(), f
    d, dst = m.distance(*args)
    de, dste = oracle.distance(w["shapes"], w["verts"], *args, nthreads=8)
    assert np.array_equal(dst, dste) and bits_equal(d, de).all()
    b = W.cfg4_aabbs(20_000, extent=200.0 * (20_000 / 1e6) ** (1 / 3))["aabbs"]
    perm, pairs, axis = m.sap3(b, axis=1)
    perm_e, pairs_e, cnt, axis_e = oracle.sap3(b, axis=1)
    assert np.array_equal(perm, perm_e) and np.array_equal(pairs, pairs_e) and axis == axis_e
    bad = w["l_shape"].copy()
    bad[-1] = 1 << 30   # lands in the last device's slice
    with pytest.raises(Cb2Error) as e:
        m.intersection(0, bad, w["r_shape"], w["l_t"], w["r_t"])
    assert e.value.code == -3
    m.close()
