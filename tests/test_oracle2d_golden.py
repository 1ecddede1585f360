"""Pins the 2-D CPU oracle (oracle/collision_oracle2d.cpp) against the reference's OWN 2-D
known-answer tests.  Every case is lifted from a `#[test]` in /root/reference/src (cited per test);
`assert_eq!` -> exact, `assert_ulps_eq!` -> 4 ulps or |diff| <= f32::EPSILON."""
import numpy as np
import pytest

from helpers import circle, line2, particle2, polygon, rectangle, square, transform_2d
from test_oracle_golden import EPS, ulps_eq


@pytest.fixture(scope="module")
def o2():
    from oracle.binding import Oracle2D
    return Oracle2D()


# ---- gjk/simplex/simplex2d.rs:107-195 ----------------------------------------------------------------
def test_check_origin_empty_and_single(o2):
    hit, s, d = o2.reduce_to_closest_feature([], (1, 0))
    assert not hit and len(s) == 0 and tuple(d) == (1, 0)
    hit, s, d = o2.reduce_to_closest_feature([(40, 0)], (1, 0))
    assert not hit and len(s) == 1 and tuple(d) == (1, 0)


def test_check_origin_edge(o2):
    hit, s, d = o2.reduce_to_closest_feature([(40, 10), (-10, 10)], (1, 0))
    assert not hit and len(s) == 2
    assert 0.0 - d[0] <= EPS and d[1] < 0


def test_check_origin_triangle_outside_ac(o2):
    hit, s, d = o2.reduce_to_closest_feature([(40, 10), (-10, 10), (0, 3)], (1, 0))
    assert not hit and len(s) == 2 and d[0] < 0 and d[1] < 0


def test_check_origin_triangle_outside_ab(o2):
    hit, s, d = o2.reduce_to_closest_feature([(40, 10), (10, 10), (3, -3)], (1, 0))
    assert not hit and len(s) == 2 and d[0] < 0 and d[1] > 0


def test_check_origin_triangle_hit(o2):
    hit, s, d = o2.reduce_to_closest_feature([(40, 10), (-10, 10), (0, -3)], (1, 0))
    assert hit and len(s) == 3 and tuple(d) == (1, 0)


def test_closest_point_to_origin(o2):
    p, s = o2.closest_point_to_origin([(40, 10), (-10, 10), (0, 3)])
    assert len(s) == 2 and ulps_eq(p, (0, 3))
    p, s = o2.closest_point_to_origin([(40, 10), (-10, 10), (0, -3)])
    assert len(s) == 3 and ulps_eq(p, (0, 0))
    p, s = o2.closest_point_to_origin([(40, 10), (-10, 10)])
    assert len(s) == 2 and ulps_eq(p, (0, 10))


# ---- primitive/util.rs:160-241 --------------------------------------------------------------------
TRIANGLE = np.array([(-1, 1), (0, -1), (1, 0)], np.float32)


@pytest.mark.parametrize("d,exp,angle", [
    ((0, 1), (-1, 1), 0.0), ((-1, 0), (-1, 1), 0.0), ((0, -1), (0, -1), 0.0), ((1, 0), (1, 0), 0.0),
    ((10, 1), (1, 0), 0.0), ((2, -100), (0, -1), 0.0),
    ((0, 1), (np.float32(0.70710678),) * 2, np.float32(np.pi) / np.float32(4)),
])
def test_max_point(o2, d, exp, angle):
    p, st = o2.support_point(polygon(0, 3), TRIANGLE, transform_2d(0, 0, angle), d)
    assert st == 1 and ulps_eq(p, exp)


def test_max_point_disp(o2):
    p, _ = o2.support_point(polygon(0, 3), TRIANGLE, transform_2d(0, 8, 0), (0, 1))
    assert tuple(p) == (-1, 9)


def test_closest_point_2d(o2):
    z = (0, 0)
    assert ulps_eq(o2.closest_point_on_edge((3, -1), (-1, 3), z), (1, 1))
    assert ulps_eq(o2.closest_point_on_edge((2, -2), (2, 2), z), (2, 0))
    assert ulps_eq(o2.closest_point_on_edge((2, 4), (2, 2), z), (2, 2))
    assert ulps_eq(o2.closest_point_on_edge((2, -2), (2, -4), z), (2, -2))


# ---- primitive/circle.rs:106-128 --------------------------------------------------------------------
@pytest.mark.parametrize("d,exp,angle", [
    ((1, 0), (10, 0), 0.0),
    ((1, 1), (7.0710677, 7.0710677), 0.0),
    ((1, 0), (10, 0), -np.float32(np.pi) / np.float32(4)),
])
def test_circle_far(o2, d, exp, angle):
    p, st = o2.support_point(circle(10.0), None, transform_2d(0, 0, angle), d)
    assert st == 1 and ulps_eq(p, exp)


def test_circle_far_4(o2):
    p, _ = o2.support_point(circle(10.0), None, transform_2d(0, 10, 0), (1, 0))
    assert tuple(p) == (10, 10)


# ---- primitive/polygon.rs:210-243: the neighbour walk on the 17-gon ------------------------------------
GON17 = np.array([(5, 5), (4, 6), (3, 7), (2, 6), (1, 6), (0, 5), (-1, 4), (-3, 3), (-6, 1), (-5, 0),
                  (-4, -1), (-2, -3), (0, -7), (1, -8), (2, -5), (3, 0), (4, 3)], np.float32)


@pytest.mark.parametrize("d,exp", [((-1, 0), (-5, 0)), ((0, -1), (1, -8)), ((0, 1), (3, 7)),
                                   ((1, 0), (5, 5))])
def test_polygon_support_point(o2, d, exp):
    p, st = o2.support_point(polygon(0, 17), GON17, transform_2d(0, 0, 0), d)
    assert st == 1 and tuple(p) == exp


# ---- epa/epa2d.rs:166-263 -------------------------------------------------------------------------------
def test_closest_edge_0_1_2(o2):
    assert o2.closest_edge([])[0] == 0
    assert o2.closest_edge([(10, 10)])[0] == 0
    assert o2.closest_edge([(10, 10), (-10, 5)])[0] == 0


def test_closest_edge_3(o2):
    st, n, dist, idx = o2.closest_edge([(10, 10), (-10, 5), (2, -5)])
    assert st == 1 and idx == 2
    assert ulps_eq(dist, 2.5607374) and ulps_eq(n[0], -0.6401844) and ulps_eq(n[1], -0.7682213)


L, LT, R_, RT = rectangle(10, 10), transform_2d(15, 0, 0), rectangle(10, 10), transform_2d(7, 2, 0)


@pytest.mark.parametrize("simplex", [[], [(-2, 8)], [(-2, 8), (18, -12)]])
def test_epa_0_1_2(o2, simplex):
    st, _ = o2.epa2_process(simplex, L, LT, R_, RT)
    assert st == 0


def test_epa_3(o2):
    st, c = o2.epa2_process([(-2, 8), (18, -12), (-2, -12)], L, LT, R_, RT)
    assert st == 1
    assert tuple(c["normal"]) == (-1, 0)
    assert 2.0 - c["depth"] <= EPS


# ---- gjk/mod.rs:623-779 (the GJK2 cases) ------------------------------------------------------------------
def _one(a):
    return np.ascontiguousarray(a)


def _intersection(o2, strategy, l, lt, r, rt, verts=None):
    shapes = np.concatenate([l, r])
    out, st = o2.intersection(strategy, shapes, verts, np.array([0], np.uint32), np.array([1], np.uint32),
                              _one(lt), _one(rt))
    return st[0], out[0]


def _distance(o2, l, lt, r, rt):
    shapes = np.concatenate([l, r])
    out, st = o2.distance(shapes, None, np.array([0], np.uint32), np.array([1], np.uint32), _one(lt), _one(rt))
    return st[0], out[0]


def test_gjk_exact(o2):
    s, t = rectangle(1, 1), transform_2d(0, 0, 0)
    assert _intersection(o2, 0, s, t, s, t)[0] == 1
    assert _distance(o2, s, t, s, t)[0] == 0


def test_gjk_miss(o2):
    l, lt, r, rt = rectangle(10, 10), transform_2d(15, 0, 0), rectangle(10, 10), transform_2d(-15, 0, 0)
    assert o2.gjk2_intersect(l, lt, r, rt)[0] == 0
    assert _intersection(o2, 0, l, lt, r, rt)[0] == 0


def test_gjk_hit(o2):
    st, simplex = o2.gjk2_intersect(L, LT, R_, RT)
    assert st == 1 and len(simplex) == 3
    st, c = _intersection(o2, 0, L, LT, R_, RT)
    assert st == 1
    assert tuple(c["normal"]) == (-1, 0)
    assert 2.0 - c["depth"] <= EPS
    assert tuple(c["point"]) == (10, 1)
    # CollisionOnly: Contact::new(CollisionOnly), contact.rs:54-56
    st, c = _intersection(o2, 1, L, LT, R_, RT)
    assert st == 1 and c["strategy"] == 1 and c["depth"] == 0


def test_gjk_distance_2d(o2):
    st, d = _distance(o2, L, LT, R_, transform_2d(0, 0, 0))
    assert st == 1 and d == 5.0
    assert _distance(o2, L, LT, R_, RT)[0] == 0


def test_gjk_time_of_impact_2d(o2):
    l, r = rectangle(10, 20), rectangle(10, 11)
    shapes = np.concatenate([l, r])
    t0, t1, rt = transform_2d(0, 0, 0), transform_2d(30, 0, 0), transform_2d(15, 0, 0)
    i0, i1 = np.array([0], np.uint32), np.array([1], np.uint32)
    out, st = o2.time_of_impact(shapes, None, i0, i1, t0, t1, rt, rt)
    c = out[0]
    assert st[0] == 1
    assert ulps_eq(c["toi"], 0.1666667)
    assert tuple(c["normal"]) == (-1, 0)
    assert 0.0 - c["depth"] <= EPS
    assert tuple(c["point"]) == (10, 0)
    out, st = o2.time_of_impact(shapes, None, i0, i1, t0, t0, rt, rt)
    assert st[0] == 0


# ---- primitive/line.rs:53-62 ---------------------------------------------------------------------------------
def test_line_rectangle_intersect(o2):
    st, _ = o2.gjk2_intersect(line2(0, -1, 0, 1), transform_2d(1, 0, 0), rectangle(1, 0.2),
                              transform_2d(1.1, 0, 0))
    assert st == 1


# ---- restated-only behaviour (no reference test): Square = Rectangle(dim, dim), Particle, panics -------------
def test_square_equals_rectangle_and_particle(o2):
    t = transform_2d(0.25, -0.5, 0.3)
    for d in [(1, 0), (0.3, -0.9), (-1, 1)]:
        assert tuple(o2.support_point(square(3.0), None, t, d)[0]) == tuple(
            o2.support_point(rectangle(3.0, 3.0), None, t, d)[0])
        assert tuple(o2.support_point(particle2(), None, t, d)[0]) == (np.float32(0.25), np.float32(-0.5))


def test_panic_statuses(o2):
    t = transform_2d(0, 0, 0, scale=0.0)
    assert o2.support_point(circle(1.0), None, t, (1, 0))[1] == 4  # inverse_transform_vector().unwrap()
    t = transform_2d(0, 0, 0)
    t["m"] = (1, 1, 1, 1)  # singular Basis2: Matrix2::invert() is None
    assert o2.support_point(circle(1.0), None, t, (1, 0))[1] == 10
    nan = np.float32(np.nan)
    assert o2.closest_edge([(nan, nan), (nan, 1), (1, nan)])[0] == 9  # assert_ulps_ne!(inf, distance)


# ---- bounds: circle.rs:131-134, rectangle.rs:198-201, polygon.rs:245-270, util.rs:146-158, tests/aabb.rs:508-519
def test_bounds_and_aabb2_transform(o2):
    from oracle.binding import SHAPE2_DT
    ident = transform_2d(0, 0, 0)
    tri = np.array([(-1, 1), (0, -1), (1, 0)], np.float32)
    verts = np.concatenate([GON17, tri])
    shapes = np.concatenate([circle(10.0), rectangle(10, 10), polygon(0, 17), polygon(17, 3), rectangle(10, 20),
                             particle2(), line2(3, -1, -2, 4)])
    t = np.concatenate([ident] * 7)
    b = o2.world_aabbs(shapes, verts, np.arange(7, dtype=np.uint32), t)
    assert tuple(b[0]) == (-10, -10, 10, 10)      # test_circle_bound
    assert tuple(b[1]) == (-5, -5, 5, 5)          # test_rectangle_bound
    assert tuple(b[2]) == (-6, -8, 5, 7)          # polygon test_bound
    assert tuple(b[3]) == (-1, -1, 1, 1)          # util.rs test_get_bound
    assert tuple(b[5]) == (0, 0, 0, 0) and tuple(b[6]) == (-2, -1, 3, 4)
    # tests/aabb.rs test_aabb2_transform: Aabb2((0,0),(10,20)) moved by (5,3) -> ((5,3),(15,23)); the same box
    # as a Rectangle(10, 20) centred at (5, 10)
    moved = o2.world_aabbs(shapes, verts, np.array([4], np.uint32), transform_2d(5 + 5, 3 + 10, 0))
    assert tuple(moved[0]) == (5, 3, 15, 23)
