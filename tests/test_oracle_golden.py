"""Pins the CPU oracle against the reference's OWN known-answer tests (3-D, f32).

Every case below is lifted from a `#[test]` in /root/reference (cited per test).
The Rust crate itself cannot be built in this environment, so these vectors are
what anchors the restatement.  Equality semantics follow
the reference assertion: `assert_eq!` -> exact, `assert_ulps_eq!` -> 4 ulps or
|diff| <= f32::EPSILON.
"""
import numpy as np
import pytest

from helpers import aabb, cuboid, polyhedron, sphere, transform_3d

EPS = np.float32(1.1920929e-7)


def ulps_eq(a, b, max_ulps=4):
    a = np.asarray(a, np.float32).ravel()
    b = np.asarray(b, np.float32).ravel()
    ok = np.abs(a - b) <= EPS
    ia = a.view(np.int32).astype(np.int64)
    ib = b.view(np.int32).astype(np.int64)
    same_sign = np.signbit(a) == np.signbit(b)
    ok |= same_sign & (np.abs(ia - ib) <= max_ulps)
    return bool(ok.all())


IDENT = transform_3d(0, 0, 0, 0)


# ---- primitive/sphere.rs:114-152 --------------------------------------------------------
@pytest.mark.parametrize("d,exp,angle", [
    ((1, 0, 0), (10, 0, 0), 0.0),
    ((1, 1, 1), (5.773502691896258,) * 3, 0.0),
    ((1, 0, 0), (10, 0.0000009536743, 0), -np.pi / 4),
])
def test_sphere_support_1_2_3(oracle, d, exp, angle):
    p, st = oracle.support_point(sphere(10.0), None, transform_3d(0, 0, 0, angle), d)
    assert st == 1
    assert ulps_eq(p, exp)


def test_sphere_support_3_bit_pattern(oracle):
    # survey §4: quaternion-rotation rounding gives exactly 9.536743e-7 = 2^-20
    p, _ = oracle.support_point(sphere(10.0), None, transform_3d(0, 0, 0, -np.pi / 4), (1, 0, 0))
    assert p[1] == np.float32(2.0 ** -20)


def test_sphere_support_4(oracle):
    p, _ = oracle.support_point(sphere(10.0), None, transform_3d(0, 10, 0, 0), (1, 0, 0))
    assert tuple(p) == (10.0, 10.0, 0.0)


def test_sphere_and_cuboid_bound(oracle):
    # sphere.rs:154-160 test_sphere_bound; cuboid.rs:231-234 test_rectangle_bound
    from oracle.binding import SHAPE_DT
    shapes = np.concatenate([sphere(10.0), cuboid(10, 10, 10)])
    b = oracle.world_aabbs(shapes, None, np.array([0, 1], np.uint32),
                           np.concatenate([IDENT, IDENT]))
    assert tuple(b[0]["min"]) == (-10, -10, -10) and tuple(b[0]["max"]) == (10, 10, 10)
    assert tuple(b[1]["min"]) == (-5, -5, -5) and tuple(b[1]["max"]) == (5, 5, 5)


# ---- primitive/polyhedron.rs:605-651 -------------------------------------------------------
TETRA = np.array([[1, 0, 0], [0, 1, 0], [0, 0, 1], [0, 0, 0]], np.float32)


def test_polytope_support_vertex_only(oracle):
    s = polyhedron(0, 4)
    p, _ = oracle.support_point(s, TETRA, IDENT, (1, 0, 0))
    assert tuple(p) == (1, 0, 0)
    p, _ = oracle.support_point(s, TETRA, IDENT, (0, 1, 0))
    assert tuple(p) == (0, 1, 0)


def test_polytope_bound(oracle):
    b = oracle.world_aabbs(polyhedron(0, 4), TETRA, np.array([0], np.uint32), IDENT)
    assert tuple(b[0]["min"]) == (0, 0, 0) and tuple(b[0]["max"]) == (1, 1, 1)


# ---- primitive/util.rs:229-250 (2-D vectors embedded with z = 0) -------------------------------
@pytest.mark.parametrize("start,end,exp", [
    ((3, -1, 0), (-1, 3, 0), (1, 1, 0)),
    ((2, -2, 0), (2, 2, 0), (2, 0, 0)),
    ((2, 4, 0), (2, 2, 0), (2, 2, 0)),
    ((2, -2, 0), (2, -4, 0), (2, -2, 0)),
])
def test_closest_point_on_edge(oracle, start, end, exp):
    assert ulps_eq(oracle.closest_point_on_edge(start, end, (0, 0, 0)), exp)


# ---- tests/plane.rs:5-47 (f64 in the reference; these inputs are exact in f32) ----------------
def test_plane_from_points(oracle):
    n, d = oracle.plane_from_points((5, 0, 5), (5, 5, 5), (5, 0, -1))
    # Plane::from_abcd(-1,0,0,5) == {n:(-1,0,0), d:5}
    assert tuple(n) == (-1, 0, 0) and d == 5
    assert oracle.plane_from_points((0, 5, -5), (0, 5, 0), (0, 5, 5)) is None


# ---- tests/aabb.rs:246-283 ---------------------------------------------------------------------
@pytest.mark.parametrize("bmin,bmax,exp", [
    ((15, 15, 15), (25, 25, 25), False),
    ((5, 5, 5), (15, 15, 15), True),
    ((5, 11, 11), (15, 13, 13), False),
    ((11, 5, 11), (15, 13, 13), False),
    ((11, 11, 5), (15, 13, 13), False),
])
def test_aabb3_intersects(oracle, bmin, bmax, exp):
    a = np.array([0, 0, 0, 10, 10, 10], np.float32)
    b = np.array(list(bmin) + list(bmax), np.float32)
    assert oracle.aabb3_intersects(a, b) is exp


def test_aabb3_touching_is_not_intersecting(oracle):
    # strict inequalities, aabb3.rs:246-253
    a = np.array([0, 0, 0, 10, 10, 10], np.float32)
    b = np.array([10, 0, 0, 20, 10, 10], np.float32)
    assert oracle.aabb3_intersects(a, b) is False


# ---- gjk/simplex/simplex3d.rs:234-373 -------------------------------------------------------------
def test_check_side_outside_ab(oracle):
    s, v = oracle.check_side([(8, -10, 0), (-1, -10, 0), (3, 5, 0)], False, False)
    assert len(s) == 2 and tuple(s[0]) == (-1, -10, 0)
    assert ulps_eq(v, (-375, 100, 0))


def test_check_side_outside_ac(oracle):
    s, v = oracle.check_side([(2, -10, 0), (-7, -10, 0), (-3, 5, 0)], False, False)
    assert len(s) == 2 and tuple(s[0]) == (2, -10, 0)
    assert ulps_eq(v, (300, 100, 0))


def test_check_side_above(oracle):
    s, v = oracle.check_side([(5, -10, -1), (-4, -10, -1), (0, 5, -1)], False, False)
    assert len(s) == 3 and tuple(s[0]) == (5, -10, -1)
    assert ulps_eq(v, (0, 0, 135))


def test_check_side_below(oracle):
    s, v = oracle.check_side([(5, -10, 1), (-4, -10, 1), (0, 5, 1)], False, False)
    assert len(s) == 3 and tuple(s[0]) == (-4, -10, 1)
    assert ulps_eq(v, (0, 0, -135))


def test_check_origin_empty_and_point(oracle):
    hit, s, v = oracle.reduce_to_closest_feature(np.zeros((0, 3)))
    assert not hit and len(s) == 0 and tuple(v) == (0, 0, 0)
    hit, s, v = oracle.reduce_to_closest_feature([(8, -10, 0)])
    assert not hit and len(s) == 1 and tuple(v) == (0, 0, 0)


def test_check_origin_line(oracle):
    hit, s, v = oracle.reduce_to_closest_feature([(8, -10, 0), (-1, -10, 0)])
    assert not hit and len(s) == 2 and tuple(v) == (0, 810, 0)


def test_check_origin_triangle(oracle):
    hit, s, v = oracle.reduce_to_closest_feature([(5, -10, -1), (-4, -10, -1), (0, 5, -1)])
    assert not hit and len(s) == 3 and tuple(v) == (0, 0, 135)


def test_check_origin_tetrahedron_outside_abc(oracle):
    hit, s, v = oracle.reduce_to_closest_feature(
        [(8, -10, -1), (-1, -10, -1), (3, 5, -1), (3, -3, 5)])
    assert not hit and len(s) == 3
    assert tuple(s[0]) == (-1, -10, -1)
    assert tuple(v) == (-90, 24, 32)


def test_check_origin_tetrahedron_outside_acd(oracle):
    f = np.float32
    hit, s, v = oracle.reduce_to_closest_feature(
        [(8, 0.1, -1), (-1, 0.1, -1), (3, 15, -1), (3, 7, 5)])
    assert not hit and len(s) == 3
    assert tuple(s[2]) == (3, 7, 5)
    assert tuple(s[1]) == (f(-1), f(0.1), f(-1))
    assert tuple(s[0]) == (f(8), f(0.1), f(-1))
    assert tuple(v) == (f(0), f(-54), f(62.1))


def test_check_origin_tetrahedron_outside_adb(oracle):
    hit, s, v = oracle.reduce_to_closest_feature(
        [(2, -10, -1), (-7, -10, -1), (-3, 5, -1), (-3, -3, 5)])
    assert not hit and len(s) == 3
    assert tuple(s[2]) == (-3, -3, 5)
    assert tuple(s[1]) == (2, -10, -1)
    assert tuple(s[0]) == (-3, 5, -1)
    assert tuple(v) == (90, 30, 40)


def test_check_origin_tetrahedron_inside(oracle):
    hit, s, _ = oracle.reduce_to_closest_feature(
        [(3, -3, -1), (-3, -3, -1), (0, 3, -1), (0, 0, 5)])
    assert hit and len(s) == 4


# ---- epa/epa3d.rs:230-341 -------------------------------------------------------------------------
def test_remove_or_add_edge(oracle):
    assert oracle.remove_or_add_edge([(1, 2), (6, 5)], (4, 3)) == [(1, 2), (6, 5), (4, 3)]
    assert oracle.remove_or_add_edge([(1, 2), (6, 5)], (2, 1)) == [(6, 5)]


TET = [(3, -3, -1), (-3, -3, -1), (0, 3, -1), (0, 0, 5)]


def test_face_impl(oracle):
    idx, nd, closest, nv = oracle.polytope(TET)
    assert idx.tolist() == [[3, 2, 1], [3, 1, 0], [3, 0, 2], [2, 0, 1]]
    exp = [(-0.8728715, 0.43643576, 0.21821788, 1.0910894),
           (0.0, -0.89442724, 0.44721362, 2.236068),
           (0.8728715, 0.43643576, 0.21821788, 1.0910894),
           (0.0, 0.0, -1.0, 1.0)]
    for got, e in zip(nd, exp):
        assert ulps_eq(got, e)
    # test_polytope_closest_to_origin
    assert closest == 3 and nv == 4


def test_polytope_add(oracle):
    idx, nd, closest, nv = oracle.polytope(TET, add=[(0, 0, -2)])
    assert nv == 5 and len(idx) == 6
    assert idx[3].tolist() == [4, 2, 0]
    assert idx[4].tolist() == [4, 0, 1]
    assert idx[5].tolist() == [4, 1, 2]


L10 = cuboid(10, 10, 10)


def test_epa_3d(oracle):
    simplex = [(18, -12, 0), (-2, 8, 0), (-2, -12, 0), (8, -2, -10)]
    st, c = oracle.epa3_process(simplex, L10, transform_3d(15, 0, 0), L10, transform_3d(7, 2, 0))
    assert st == 1
    assert tuple(c["normal"]) == (-1, 0, 0)
    assert np.float32(2) - c["depth"] <= EPS


# ---- gjk/mod.rs:634-825 ---------------------------------------------------------------------------
def _pair(oracle, fn, l, lt, r, rt, **kw):
    shapes = np.concatenate([l, r])
    return fn(shapes, None, np.array([0], np.uint32), np.array([1], np.uint32), lt, rt, **kw)


def test_gjk_exact_3d(oracle):
    s = cuboid(1, 1, 1)
    out, st = oracle.intersection(0, np.concatenate([s, s]), None, np.array([0], np.uint32),
                                  np.array([1], np.uint32), IDENT, IDENT)
    assert st[0] == 1
    d, st = _pair(oracle, oracle.distance, s, IDENT, s, IDENT)
    assert st[0] == 0


def test_gjk_sphere(oracle):
    s = sphere(1.0)
    out, st = oracle.intersection(0, np.concatenate([s, s]), None, np.array([0], np.uint32),
                                  np.array([1], np.uint32), IDENT, IDENT)
    assert st[0] == 1
    d, st = _pair(oracle, oracle.distance, s, IDENT, s, IDENT)
    assert st[0] == 0


def test_gjk_3d_hit(oracle):
    lt, rt = transform_3d(15, 0, 0), transform_3d(7, 2, 0)
    st, simplex = oracle.gjk3_intersect(L10, lt, L10, rt)
    assert st == 1 and len(simplex) == 4
    # survey §4: the simplex GJK returns here is exactly the hard-coded one of test_epa_3d
    assert simplex.tolist() == [[18, -12, 0], [-2, 8, 0], [-2, -12, 0], [8, -2, -10]]
    out, st = oracle.intersection(0, np.concatenate([L10, L10]), None, np.array([0], np.uint32),
                                  np.array([1], np.uint32), lt, rt)
    assert st[0] == 1
    c = out[0]
    assert c["strategy"] == 0
    assert tuple(c["normal"]) == (-1, 0, 0)
    assert np.float32(2) - c["depth"] <= EPS
    assert ulps_eq(c["point"], (10, 1, 5))
    # CollisionOnly => Contact::new(CollisionOnly): zero payload
    out, st = oracle.intersection(1, np.concatenate([L10, L10]), None, np.array([0], np.uint32),
                                  np.array([1], np.uint32), lt, rt)
    assert st[0] == 1 and out[0]["strategy"] == 1 and out[0]["depth"] == 0
    assert tuple(out[0]["normal"]) == (0, 0, 0)


def test_gjk_3d_miss(oracle):
    # 3-D analogue of test_gjk_miss (gjk/mod.rs:655-674)
    out, st = oracle.intersection(0, np.concatenate([L10, L10]), None, np.array([0], np.uint32),
                                  np.array([1], np.uint32), transform_3d(15, 0, 0),
                                  transform_3d(-15, 0, 0))
    assert st[0] == 0


def test_gjk_distance_3d(oracle):
    lt = transform_3d(15, 0, 0)
    d, st = _pair(oracle, oracle.distance, L10, lt, L10, transform_3d(7, 2, 0))
    assert st[0] == 0
    d, st = _pair(oracle, oracle.distance, L10, lt, L10, transform_3d(1, 0, 0))
    assert st[0] == 1 and d[0] == np.float32(4.0)


def test_gjk_time_of_impact_3d(oracle):
    l, r = cuboid(10, 11, 10), cuboid(10, 15, 10)
    shapes = np.concatenate([l, r])
    i0, i1 = np.array([0], np.uint32), np.array([1], np.uint32)
    l0, l1, rt = transform_3d(0, 0, 0), transform_3d(30, 0, 0), transform_3d(15, 0, 0)
    out, st = oracle.time_of_impact(shapes, None, i0, i1, l0, l1, rt, rt)
    assert st[0] == 1
    c = out[0]
    assert ulps_eq(c["toi"], 0.1666667)
    assert tuple(c["normal"]) == (-1, 0, 0)
    assert np.float32(0) - c["depth"] <= EPS
    assert tuple(c["point"]) == (10, 0, 0)
    out, st = oracle.time_of_impact(shapes, None, i0, i1, l0, l0, rt, rt)
    assert st[0] == 0


# ---- algorithm/broad_phase/sweep_prune.rs:317-361 (2-D boxes embedded with z in (0,1)) ---------------
def _boxes(*b):
    a = np.concatenate([aabb((x0, y0, 0), (x1, y1, 1)) for (x0, y0, x1, y1) in b])
    return a


@pytest.mark.parametrize("boxes,exp", [
    ([(8, 8, 10, 11), (12, 13, 18, 18)], []),
    ([(12, 13, 18, 18), (8, 8, 10, 11)], []),
    ([(8, 8, 10, 11), (9, 10, 18, 18)], [(0, 1)]),
    ([(9, 10, 18, 18), (8, 8, 10, 11)], [(0, 1)]),
])
def test_sweep_prune(oracle, boxes, exp):
    perm, pairs, cnt, axis = oracle.sap3(_boxes(*boxes))
    assert cnt == len(exp)
    assert [tuple(p) for p in pairs.tolist()] == exp
    assert axis == 0
    # the slice ends up sorted by min.x
    xs = [boxes[i][0] for i in perm]
    assert xs == sorted(xs)


# ---- the remaining Primitive3 variants (capsule.rs:213-265, cylinder.rs:243-295, quad.rs:143-153)
def _norm(v):
    v = np.asarray(v, np.float32)
    # InnerSpace::normalize = v * (1 / |v|), f32 one operation at a time
    m = np.sqrt(np.float32(np.float32(v[0] * v[0] + v[1] * v[1]) + v[2] * v[2]))
    return (v * np.float32(np.float32(1) / m)).astype(np.float32)


def test_capsule_and_cylinder_bounds(oracle):
    # capsule.rs:214-220 test_capsule_aabb; cylinder.rs:244-250 test_cylinder_aabb
    from helpers import shape
    shapes = np.concatenate([shape(6, (2.0, 1.0, 0)), shape(5, (2.0, 1.0, 0)), shape(3),
                             shape(4, (2.0, 4.0, 0)), shape(7, (10.0, 0, 0))])
    t = np.concatenate([IDENT] * 5)
    b = oracle.world_aabbs(shapes, None, np.arange(5, dtype=np.uint32), t)
    assert tuple(b[0]["min"]) == (-1, -3, -1) and tuple(b[0]["max"]) == (1, 3, 1)
    assert tuple(b[1]["min"]) == (-1, -2, -1) and tuple(b[1]["max"]) == (1, 2, 1)
    assert tuple(b[2]["min"]) == (0, 0, 0) and tuple(b[2]["max"]) == (0, 0, 0)      # Aabb3::zero()
    assert tuple(b[3]["min"]) == (-1, -2, 0) and tuple(b[3]["max"]) == (1, 2, 0)
    assert tuple(b[4]["min"]) == (-5, -5, -5) and tuple(b[4]["max"]) == (5, 5, 5)


@pytest.mark.parametrize("d,t,exp", [
    ((1, 0, 0), (0, 0, 0, 0.0), (1, 2, 0)),                                  # test_capsule_support_1
    (_norm((0.5, -1, 0)), (0, 0, 0, 0.0), (0.447_213_65, -2.894_427_3, 0)),  # _2
    ((0, 1, 0), (0, 0, 0, 0.0), (0, 3, 0)),                                  # _3
    ((1, 0, 0), (10, 0, 0, 0.0), (11, 2, 0)),                                # _4
    ((1, 0, 0), (0, 0, 0, np.pi), (1, -2, 0)),                               # _5
])
def test_capsule_support(oracle, d, t, exp):
    from helpers import shape
    p, st = oracle.support_point(shape(6, (2.0, 1.0, 0)), None, transform_3d(*t), d)
    assert st == 1 and ulps_eq(p, exp)


@pytest.mark.parametrize("d,t,exp", [
    ((1, 0, 0), (0, 0, 0, 0.0), (1, 2, 0)),                 # test_cylinder_support_1
    (_norm((0.5, -1, 0)), (0, 0, 0, 0.0), (1, -2, 0)),      # _2
    ((0, 1, 0), (0, 0, 0, 0.0), (0, 2, 0)),                 # _3
    ((1, 0, 0), (10, 0, 0, 0.0), (11, 2, 0)),               # _4
    ((1, 0, 0), (0, 0, 0, np.pi), (1, -2, 0)),              # _5
])
def test_cylinder_support(oracle, d, t, exp):
    from helpers import shape
    p, st = oracle.support_point(shape(5, (2.0, 1.0, 0)), None, transform_3d(*t), d)
    assert st == 1 and ulps_eq(p, exp)


def test_quad_cuboid_intersect(oracle):
    # quad.rs:143-153 test_plane_cuboid_intersect: Quad(2,2) at z=1 against Cuboid(1,1,1) at z=1.1
    from helpers import shape
    shapes = np.concatenate([shape(4, (2.0, 2.0, 0)), cuboid(1, 1, 1)])
    out, st = oracle.intersection(1, shapes, None, np.array([0], np.uint32), np.array([1], np.uint32),
                                  transform_3d(0, 0, 1.0), transform_3d(0, 0, 1.1))
    assert st[0] == 1


def test_particle_support_ignores_direction_and_scale(oracle):
    # primitive3.rs:164: transform.transform_point(origin) — no inverse transform, so no panic
    # even when scale is 0
    from helpers import shape
    p, st = oracle.support_point(shape(3), None, transform_3d(1, 2, 3, 0.3, scale=0.0), (0, 1, 0))
    assert st == 1 and tuple(p) == (1, 2, 3)


def test_cube_is_a_cuboid(oracle):
    from helpers import shape
    for d in ((1, 0.2, -0.3), (-1, -1, 1), (0, 0, 0)):
        a, _ = oracle.support_point(shape(7, (3.0, 0, 0)), None, transform_3d(1, 0, 0, 0.7), d)
        b, _ = oracle.support_point(cuboid(3, 3, 3), None, transform_3d(1, 0, 0, 0.7), d)
        assert np.array_equal(a, b)


# ---- HalfEdge polyhedra: polyhedron.rs:604-660 (test_polytope_half_edge, test_polytope_bound) ----
TETRA_FACES = np.array([(1, 3, 2), (3, 1, 0), (2, 0, 1), (0, 2, 3)], np.uint32)


def test_polytope_half_edge(oracle):
    from helpers import polyhedron_he
    s = polyhedron_he(0, 4, 0, 4)
    oracle.build_half_edges(s, TETRA, TETRA_FACES)
    for d in ((1, 0, 0), (0, 1, 0)):
        p, st = oracle.support_point(s, TETRA, IDENT, d)
        assert st == 1 and tuple(p) == d
        q, _ = oracle.support_point(polyhedron(0, 4), TETRA, IDENT, d)
        assert tuple(q) == d
    b = oracle.world_aabbs(s, TETRA, np.array([0], np.uint32), IDENT)
    assert tuple(b[0]["min"]) == (0, 0, 0) and tuple(b[0]["max"]) == (1, 1, 1)


def oriented_hull_faces(pts):
    """Triangles of the convex hull, all wound counter-clockwise seen from outside (the half-edge
    rings only close on a consistently oriented closed mesh)."""
    from scipy.spatial import ConvexHull
    hull = ConvexHull(np.asarray(pts, np.float64))
    faces = hull.simplices.copy()
    p = np.asarray(pts, np.float64)
    n = np.cross(p[faces[:, 1]] - p[faces[:, 0]], p[faces[:, 2]] - p[faces[:, 0]])
    flip = np.einsum("ij,ij->i", n, hull.equations[:, :3]) < 0
    faces[flip] = faces[flip][:, [0, 2, 1]]
    return faces.astype(np.uint32)


def test_half_edge_hill_climb_vs_scan_on_random_hulls(oracle):
    """The hill climb reaches the same DOT as the scan on a convex hull (it may stop on another
    vertex when dots tie), and a NaN direction never returns in the reference -> status 5."""
    from scipy.spatial import ConvexHull
    from helpers import polyhedron_he
    rng = np.random.default_rng(8)
    pts = rng.standard_normal((60, 3))
    pts /= np.linalg.norm(pts, axis=1, keepdims=True)
    pts = pts.astype(np.float32)
    faces = oriented_hull_faces(pts)
    s = polyhedron_he(0, len(pts), 0, len(faces))
    oracle.build_half_edges(s, pts, faces)
    for _ in range(300):
        d = rng.standard_normal(3).astype(np.float32)
        p, st = oracle.support_point(s, pts, IDENT, d)
        q, _ = oracle.support_point(polyhedron(0, len(pts)), pts, IDENT, d)
        assert st == 1
        assert np.float32(np.dot(p, d)) == np.float32(np.dot(q, d)) or np.allclose(np.dot(p, d), np.dot(q, d), rtol=1e-6)
    p, st = oracle.support_point(s, pts, IDENT, (np.nan, 0, 0))
    assert st == 5


def test_half_edge_rejects_degenerate_faces(oracle):
    from helpers import polyhedron_he
    pts = np.array([[0, 0, 0], [1, 0, 0], [2, 0, 0], [0, 1, 0]], np.float32)
    faces = np.array([(0, 1, 2), (0, 1, 3), (1, 2, 3), (0, 2, 3)], np.uint32)  # first face collinear
    with pytest.raises(ValueError):
        oracle.build_half_edges(polyhedron_he(0, 4, 0, 4), pts, faces)


def test_threaded_sweep_is_the_serial_sweep(oracle):
    """sap3_mt (the oracle's sweep split over threads, used for the 1 M-box GPU comparison) returns
    the serial function's permutation and pair list, ties and wide boxes included."""
    from collision_rs_b200 import workloads as W
    b = W.cfg4_aabbs(40_000, extent=200.0 * (40_000 / 1e6) ** (1 / 3))["aabbs"]
    q = b.copy()
    q["min"] = np.round(q["min"])
    q["max"] = np.round(q["max"]) + 1
    q["min"][::1501] = -100.0
    q["max"][::1501] = 100.0
    for boxes in (b, q):
        for axis in (0, 2):
            p1, r1, c1, a1 = oracle.sap3(boxes, axis=axis)
            p2, r2, c2, a2 = oracle.sap3_mt(boxes, axis=axis, nthreads=5)
            assert c1 == c2 and a1 == a2 and np.array_equal(p1, p2) and np.array_equal(r1, r2)
