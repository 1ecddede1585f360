"""GPU parity of the 2-D narrow phase (GJK2 / EPA2): the CUDA path through the
C ABI against the 2-D CPU oracle on identical inputs, every f32 bit-identical (NaN == NaN)."""
import numpy as np
import pytest

from collision_rs_b200 import abi, workloads as W
from collision_rs_b200.host import (Basis2, Circle, CollisionStrategy, ConvexPolygon, Decomposed2, GJK2,
                                    Line2, Particle2, Rectangle, ReferencePanic, Square)
from test_gpu_parity import EPS, assert_contacts_equal, bits_equal

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def o2():
    from oracle.binding import Oracle2D
    return Oracle2D()


def T(x, y, angle=0.0, scale=1.0):
    return Decomposed2((x, y), Basis2.from_angle(angle), scale)


# ---- the reference's own 2-D tests, through the reference-shaped API on the GPU ---------------------
def test_gjk_exact(ctx):  # gjk/mod.rs:623-631
    gjk = GJK2.new(ctx)
    shape, t = Rectangle(1, 1), T(0, 0)
    assert gjk.intersection(CollisionStrategy.FullResolution, shape, t, shape, t) is not None
    assert gjk.distance(shape, t, shape, t) is None


def test_gjk_miss(ctx):  # gjk/mod.rs:656-674
    gjk = GJK2.new(ctx)
    left, right = Rectangle(10, 10), Rectangle(10, 10)
    assert not gjk.intersect(left, T(15, 0), right, T(-15, 0))
    assert gjk.intersection(CollisionStrategy.FullResolution, left, T(15, 0), right, T(-15, 0)) is None


def test_gjk_hit(ctx):  # gjk/mod.rs:677-698
    gjk = GJK2.new(ctx)
    left, right = Rectangle(10, 10), Rectangle(10, 10)
    assert gjk.intersect(left, T(15, 0), right, T(7, 2))
    c = gjk.intersection(CollisionStrategy.FullResolution, left, T(15, 0), right, T(7, 2))
    assert c is not None
    assert c.normal == (-1, 0)
    assert np.float32(2) - c.penetration_depth <= EPS
    assert c.contact_point == (10, 1)


def test_gjk_distance_2d(ctx):  # gjk/mod.rs:723-740
    gjk = GJK2.new(ctx)
    left, right = Rectangle(10, 10), Rectangle(10, 10)
    assert gjk.distance(left, T(15, 0), right, T(0, 0)) == np.float32(5.0)
    assert gjk.distance(left, T(15, 0), right, T(7, 2)) is None


def test_gjk_time_of_impact_2d(ctx):  # gjk/mod.rs:762-792
    gjk = GJK2.new(ctx)
    left, right = Rectangle(10, 20), Rectangle(10, 11)
    c = gjk.intersection_time_of_impact(left, (T(0, 0), T(30, 0)), right, (T(15, 0), T(15, 0)))
    assert c is not None
    assert abs(c.time_of_impact - np.float32(0.1666667)) <= 4 * EPS
    assert c.normal == (-1, 0)
    assert np.float32(0) - c.penetration_depth <= EPS
    assert c.contact_point == (10, 0)
    assert gjk.intersection_time_of_impact(left, (T(0, 0), T(0, 0)), right, (T(15, 0), T(15, 0))) is None


def test_line_rectangle_intersect(ctx):  # primitive/line.rs:53-62
    gjk = GJK2.new(ctx)
    assert gjk.intersect(Line2((0, -1), (0, 1)), T(1, 0), Rectangle(1, 0.2), T(1.1, 0))


def test_epa_3_contact_through_gjk(ctx):  # epa2d.rs:243-262 (reached through GJK2::intersection)
    c = GJK2.new(ctx).intersection(CollisionStrategy.FullResolution, Rectangle(10, 10), T(15, 0),
                                   Rectangle(10, 10), T(7, 2))
    assert c.normal == (-1, 0) and np.float32(2) - c.penetration_depth <= EPS


def test_reference_panics_surface(ctx):
    gjk = GJK2.new(ctx)
    with pytest.raises(ReferencePanic) as e:
        gjk.intersection(CollisionStrategy.FullResolution, Circle(1.0), T(0, 0, 0, 0.0), Circle(1.0), T(1, 0))
    assert e.value.status == 4
    with pytest.raises(ReferencePanic) as e:
        gjk.distance(Circle(1.0), Decomposed2((0, 0), Basis2((1, 1, 1, 1))), Circle(1.0), T(3, 0))
    assert e.value.status == 10
    # a Particle never inverts its transform (primitive2.rs:96): no panic
    assert gjk.distance(Particle2(), T(0, 0, 0, 0.0), Circle(1.0), T(3, 0)) == np.float32(2.0)


# ---- random batches against the oracle ------------------------------------------------------------------
def _upload(ctx, w):
    ctx.upload_shapes2(w["shapes"], w["verts"])


@pytest.mark.parametrize("kinds,spread,scale", [
    ((0, 1, 2, 3, 4, 5), 1.0, 1.0),   # every Primitive2 variant
    ((3, 4), 0.75, 1.0),              # rectangles only (cfg 1's 2-D twin)
    ((5,), 0.6, 1.0),                 # polygons on both sides of the 10-vertex switch
    ((2,), 0.8, 1.0),                 # circles: EPA2 runs to the iteration cap
    ((0, 1, 2, 3, 4, 5), 1.5, 1.75),  # non-unit scale
])
def test_intersection_matches_oracle(ctx, o2, kinds, spread, scale):
    w = W.cfg2d_mixed_pairs(200_000, seed=101 + len(kinds), spread=spread, n_shapes=512, kinds=kinds,
                            scale=scale)
    _upload(ctx, w)
    for strategy in (0, 1):
        got, gst = ctx.intersection2(strategy, w["l_shape"], w["r_shape"], w["l_t"], w["r_t"])
        exp, est = o2.intersection(strategy, w["shapes"], w["verts"], w["l_shape"], w["r_shape"], w["l_t"],
                                   w["r_t"], nthreads=8)
        assert_contacts_equal(got, gst, exp, est, f"2-D intersection kinds={kinds} strategy={strategy}")
    assert 0.05 < (est == 1).mean() < 0.98


@pytest.mark.parametrize("kinds,spread", [((0, 1, 2, 3, 4, 5), 1.5), ((3, 4), 1.5), ((2, 3), 1.5), ((5,), 1.2)])
def test_distance_matches_oracle(ctx, o2, kinds, spread):
    w = W.cfg2d_mixed_pairs(200_000, seed=202 + len(kinds), spread=spread, n_shapes=512, kinds=kinds)
    _upload(ctx, w)
    got, gst = ctx.distance2(w["l_shape"], w["r_shape"], w["l_t"], w["r_t"])
    exp, est = o2.distance(w["shapes"], w["verts"], w["l_shape"], w["r_shape"], w["l_t"], w["r_t"], nthreads=8)
    assert np.array_equal(gst, est)
    assert bits_equal(got[est == 1], exp[est == 1]).all()
    assert (est == 1).mean() > 0.2


@pytest.mark.parametrize("kinds", [(0, 1, 2, 3, 4, 5), (3, 4), (5,)])
def test_time_of_impact_matches_oracle(ctx, o2, kinds):
    w = W.cfg2d_toi_pairs(200_000, seed=303 + len(kinds), n_shapes=512, kinds=kinds)
    _upload(ctx, w)
    got, gst = ctx.time_of_impact2(w["l_shape"], w["r_shape"], w["l_t0"], w["l_t1"], w["r_t0"], w["r_t1"],
                                   iter_cap=256)
    exp, est = o2.time_of_impact(w["shapes"], w["verts"], w["l_shape"], w["r_shape"], w["l_t0"], w["l_t1"],
                                 w["r_t0"], w["r_t1"], iter_cap=256, nthreads=8)
    assert_contacts_equal(got, gst, exp, est, f"2-D time of impact kinds={kinds}")
    assert (est == 1).mean() > 0.1


def test_custom_settings_and_iteration_caps(ctx, o2):
    from oracle.binding import Settings as OS
    w = W.cfg2d_mixed_pairs(50_000, seed=404, spread=0.8, n_shapes=256, kinds=(2, 3, 5))
    _upload(ctx, w)
    for mi, tol in ((0, 1e-5), (1, 1e-5), (7, 1e-3), (100, 1e-7)):
        s = abi.Settings(np.float32(1e-6), np.float32(1e-6), np.float32(tol), mi)
        so = OS(np.float32(1e-6), np.float32(1e-6), np.float32(tol), mi)
        got, gst = ctx.intersection2(0, w["l_shape"], w["r_shape"], w["l_t"], w["r_t"], s)
        exp, est = o2.intersection(0, w["shapes"], w["verts"], w["l_shape"], w["r_shape"], w["l_t"], w["r_t"],
                                   settings=so, nthreads=8)
        assert_contacts_equal(got, gst, exp, est, f"max_iterations={mi}")


def test_hazard_inputs_match_oracle(ctx, o2):
    """scale ~ 0, singular / NaN rotation matrices, NaN and huge displacements, coincident bodies."""
    w = W.cfg2d_mixed_pairs(4096, seed=505, spread=0.5, n_shapes=256)
    lt, rt = w["l_t"], w["r_t"]
    lt["scale"][0:64] = 0.0
    rt["scale"][64:128] = 1e-8
    lt["m"][128:192] = (1, 1, 1, 1)
    rt["m"][192:256] = (0, 0, 0, 0)
    lt["m"][256:320, 0] = np.nan
    rt["d"][320:384, 1] = np.nan
    lt["d"][384:448] = 3e38
    rt["d"][448:512] = lt["d"][448:512]  # coincident centres: the (1,1) seed direction
    lt["scale"][512:576] = -2.0
    _upload(ctx, w)
    for strategy in (0, 1):
        got, gst = ctx.intersection2(strategy, w["l_shape"], w["r_shape"], lt, rt)
        exp, est = o2.intersection(strategy, w["shapes"], w["verts"], w["l_shape"], w["r_shape"], lt, rt)
        assert_contacts_equal(got, gst, exp, est, "hazards/intersection")
    assert set(np.unique(est)) >= {0, 1, 4, 10}
    got, gst = ctx.distance2(w["l_shape"], w["r_shape"], lt, rt)
    exp, est = o2.distance(w["shapes"], w["verts"], w["l_shape"], w["r_shape"], lt, rt)
    assert np.array_equal(gst, est) and bits_equal(got[est == 1], exp[est == 1]).all()
    got, gst = ctx.time_of_impact2(w["l_shape"], w["r_shape"], lt, rt, rt, lt, iter_cap=64)
    exp, est = o2.time_of_impact(w["shapes"], w["verts"], w["l_shape"], w["r_shape"], lt, rt, rt, lt, iter_cap=64)
    assert_contacts_equal(got, gst, exp, est, "hazards/toi")


def test_empty_batch_and_bad_arguments(ctx):
    from collision_rs_b200.host import Cb2Error
    w = W.cfg2d_mixed_pairs(16, seed=606, n_shapes=8)
    _upload(ctx, w)
    e = np.zeros(0, np.uint32)
    out, st = ctx.intersection2(0, e, e, np.zeros(0, abi.XFORM2_DT), np.zeros(0, abi.XFORM2_DT))
    assert len(out) == 0 and len(st) == 0
    bad = w["l_shape"].copy()
    bad[3] = 8  # index past the table: a Rust slice index would panic
    with pytest.raises(Cb2Error):
        ctx.intersection2(0, bad, w["r_shape"], w["l_t"], w["r_t"])
    with pytest.raises(Cb2Error):
        ctx.intersection2(2, w["l_shape"], w["r_shape"], w["l_t"], w["r_t"])


def test_device_api_and_million_pair_properties(ctx, o2):
    """1 M pairs through the device-pointer API: FullResolution and CollisionOnly agree on every
    Option; a sampled slice is bit-identical to the oracle."""
    import torch
    n = 1_000_000
    w = W.cfg2d_mixed_pairs(n, seed=707, spread=1.0, n_shapes=4096)
    _upload(ctx, w)
    dev = torch.device("cuda:0")
    dv = lambda a: torch.from_numpy(a.view(np.uint8).reshape(-1).copy()).to(dev)
    d = {k: dv(w[k]) for k in ("l_shape", "r_shape", "l_t", "r_t")}
    out = torch.zeros(n * abi.CONTACT2_DT.itemsize, dtype=torch.uint8, device=dev)
    st_full = torch.zeros(n, dtype=torch.uint8, device=dev)
    st_only = torch.zeros(n, dtype=torch.uint8, device=dev)
    s = torch.cuda.current_stream().cuda_stream
    ctx.intersection2_dev(1, d["l_shape"].data_ptr(), d["r_shape"].data_ptr(), d["l_t"].data_ptr(),
                          d["r_t"].data_ptr(), n, out.data_ptr(), st_only.data_ptr(), s)
    ctx.intersection2_dev(0, d["l_shape"].data_ptr(), d["r_shape"].data_ptr(), d["l_t"].data_ptr(),
                          d["r_t"].data_ptr(), n, out.data_ptr(), st_full.data_ptr(), s)
    torch.cuda.synchronize()
    assert torch.equal(st_full, st_only)
    got = out.cpu().numpy().view(abi.CONTACT2_DT)
    gst = st_full.cpu().numpy()
    sl = slice(123_456, 123_456 + 50_000)
    exp, est = o2.intersection(0, w["shapes"], w["verts"], w["l_shape"][sl], w["r_shape"][sl], w["l_t"][sl],
                               w["r_t"][sl], nthreads=8)
    assert_contacts_equal(got[sl], gst[sl], exp, est, "1M slice")
    hit = gst == 1
    nrm = np.linalg.norm(got["normal"][hit], axis=1)
    finite = np.isfinite(nrm)
    assert np.allclose(nrm[finite], 1.0, atol=1e-5)


def test_complex_callers_match_the_oracle(ctx, o2):
    """GJK2::intersection_complex / distance_complex / intersection_complex_time_of_impact
    (gjk/mod.rs:412-588): expansion with Decomposed<Vector2, Basis2>::concat, the GJK2 kernels on the
    flat part pairs, the loop-order reduction — bit-identical to the oracle."""
    from helpers_complex import composite_workload2
    cw = composite_workload2(30_000, seed=9)
    ctx.upload_shapes2(cw["shapes"], cw["verts"])
    ctx.upload_composites2(cw["comp_off"], cw["comp_shape"], cw["comp_local"])
    oc = (cw["shapes"], cw["verts"], cw["comp_off"], cw["comp_shape"], cw["comp_local"], cw["l_comp"],
          cw["r_comp"], cw["l_t0"], cw["r_t0"], cw["l_t1"], cw["r_t1"])
    for strategy in (0, 1):
        got, gst = ctx.intersection_complex2(strategy, cw["l_comp"], cw["r_comp"], cw["l_t0"], cw["r_t0"])
        exp, est = o2.complex(0, strategy, *oc, nthreads=8)
        assert_contacts_equal(got, gst, exp, est, f"2-D intersection_complex strategy={strategy}")
        got, gst = ctx.time_of_impact_complex2(strategy, cw["l_comp"], cw["r_comp"], cw["l_t0"], cw["l_t1"],
                                               cw["r_t0"], cw["r_t1"], iter_cap=256)
        exp, est = o2.complex(2, strategy, *oc, iter_cap=256, nthreads=8)
        assert_contacts_equal(got, gst, exp, est, f"2-D toi_complex strategy={strategy}")
    got, gst = ctx.distance_complex2(cw["l_comp"], cw["r_comp"], cw["l_t0"], cw["r_t0"])
    exp, est = o2.complex(1, 0, *oc, nthreads=8)
    assert np.array_equal(gst, est) and bits_equal(got[est == 1], exp[est == 1]).all()
    assert (est == 1).mean() > 0.05


def test_world_aabbs2_match_oracle(ctx, o2):
    rng = np.random.default_rng(11)
    shapes, verts = W.shapes2_table(512, 77)
    ctx.upload_shapes2(shapes, verts)
    n = 100_000
    shape = rng.integers(0, len(shapes), n).astype(np.uint32)
    t = W.random_xforms2(rng, n, 5.0)
    t["scale"] = rng.uniform(0.3, 3.0, n).astype(np.float32)
    t["scale"][::97] = -1.5
    t["d"][::131, 0] = np.nan
    got = ctx.world_aabbs2(shape, t)
    exp = o2.world_aabbs(shapes, verts, shape, t)
    assert bits_equal(got, exp).all()


def test_device_chain_bodies_to_boxes_to_pairs_to_contacts(ctx, o2, oracle):
    """The 2-D pipeline without a host trip: cb2_world_aabbs2_dev -> cb2_sap2_find_collider_pairs_dev ->
    cb2_gjk2_intersection_bodies_dev, every stage bit-identical to the oracle."""
    import torch
    from oracle.binding import AABB_DT
    rng = np.random.default_rng(21)
    shapes, verts = W.shapes2_table(256, 91)
    ctx.upload_shapes2(shapes, verts)
    n = 60_000
    body_shape = rng.integers(0, len(shapes), n).astype(np.uint32)
    body_t = W.random_xforms2(rng, n, 120.0)
    dev = torch.device("cuda:0")
    dv = lambda a: torch.from_numpy(a.view(np.uint8).reshape(-1).copy()).to(dev)
    d_shape, d_t = dv(body_shape), dv(body_t)
    s = torch.cuda.current_stream().cuda_stream
    boxes = torch.zeros(n * 4, dtype=torch.float32, device=dev)
    ctx.world_aabbs2_dev(d_shape.data_ptr(), d_t.data_ptr(), n, boxes.data_ptr(), s)
    exp_boxes = o2.world_aabbs(shapes, verts, body_shape, body_t)
    torch.cuda.synchronize()
    assert bits_equal(boxes.cpu().numpy().reshape(-1, 4), exp_boxes).all()
    perm = torch.zeros(n, dtype=torch.int32, device=dev)
    cap = 16 * n
    pairs = torch.zeros(2 * cap, dtype=torch.int32, device=dev)
    rc, cnt = ctx.sap2_dev(boxes.data_ptr(), n, 0, perm.data_ptr(), pairs.data_ptr(), cap, s)
    torch.cuda.synchronize()
    lifted = np.zeros(n, AABB_DT)
    lifted["min"][:, :2], lifted["max"][:, :2] = exp_boxes[:, :2], exp_boxes[:, 2:]
    lifted["max"][:, 2] = 1.0
    perm_e, pairs_e, _, _ = oracle.sap3(lifted, 0)
    assert rc == 0 and cnt == len(pairs_e) and cnt > 1000
    got_pairs = pairs[:2 * cnt].cpu().numpy().view(np.uint32).reshape(-1, 2)
    assert np.array_equal(perm.cpu().numpy().view(np.uint32), perm_e) and np.array_equal(got_pairs, pairs_e)
    # bodies in sorted order, like the reference's permuted slice
    p64 = perm.long()
    s_shape = d_shape.view(torch.int32)[p64].contiguous()
    s_t = d_t.view(-1, 32)[p64].contiguous()
    out = torch.zeros(cnt * abi.CONTACT2_DT.itemsize, dtype=torch.uint8, device=dev)
    st = torch.zeros(cnt, dtype=torch.uint8, device=dev)
    ctx.intersection2_bodies_dev(0, s_shape.data_ptr(), s_t.data_ptr(), pairs.data_ptr(), cnt, out.data_ptr(),
                                 st.data_ptr(), s)
    torch.cuda.synchronize()
    bs, bt = body_shape[perm_e], body_t[perm_e]
    a, b = pairs_e[:, 0], pairs_e[:, 1]
    exp, est = o2.intersection(0, shapes, verts, np.ascontiguousarray(bs[a]), np.ascontiguousarray(bs[b]),
                               np.ascontiguousarray(bt[a]), np.ascontiguousarray(bt[b]), nthreads=8)
    assert_contacts_equal(out.cpu().numpy().view(abi.CONTACT2_DT), st.cpu().numpy(), exp, est, "2-D chain")
    assert (est == 1).mean() > 0.2
