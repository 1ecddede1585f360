"""Committed fixtures (tests/golden/*.npz, made by tests/golden/make_golden.py): inputs and the
pinned oracle's outputs.  CPU: the oracle still reproduces them bit for bit.  GPU (-m gpu): the
CUDA path, through the C ABI, reproduces them too — nothing outside the repository is read."""
import os

import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    with np.load(os.path.join(GOLDEN, name + ".npz")) as z:
        return {k: z[k] for k in z.files}


def same_records(got, gst, exp, est, fields):
    assert np.array_equal(gst, est)
    m = est == 1
    for f in fields:
        a, b = np.ascontiguousarray(got[f][m]), np.ascontiguousarray(exp[f][m])
        eq = (a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b))
        assert eq.all(), f
    return True


CF = ("normal", "depth", "point", "toi")


def same_floats(got, gst, exp, est):
    assert np.array_equal(gst, est)
    m = est == 1
    return np.array_equal(got[m].view(np.uint32), exp[m].view(np.uint32))


class _OracleRunner:
    """adapts oracle.binding to the call shape of collision_rs_b200.host.Context"""

    def __init__(self):
        from oracle.binding import Oracle, Oracle2D
        self.o, self.o2 = Oracle(), Oracle2D()

    def run3(self, g):
        v = g.get("verts")
        a = (g["l_shape"], g["r_shape"], g["l_t"], g["r_t"])
        return (lambda s: self.o.intersection(s, g["shapes"], v, *a),
                lambda: self.o.distance(g["shapes"], v, *a))

    def toi3(self, g):
        return self.o.time_of_impact(g["shapes"], None, g["l_shape"], g["r_shape"], g["l_t0"], g["l_t1"],
                                     g["r_t0"], g["r_t1"], iter_cap=int(g["iter_cap"][0]))

    def sap(self, boxes, axis):
        perm, pairs, _, nxt = self.o.sap3(boxes, axis)
        return perm, pairs, nxt

    def run2(self, g):
        v = g.get("verts")
        a = (g["l_shape"], g["r_shape"], g["l_t"], g["r_t"])
        return (lambda s: self.o2.intersection(s, g["shapes"], v, *a),
                lambda: self.o2.distance(g["shapes"], v, *a))

    def toi2(self, g):
        return self.o2.time_of_impact(g["shapes"], g.get("verts"), g["l_shape"], g["r_shape"], g["l_t0"],
                                      g["l_t1"], g["r_t0"], g["r_t1"], iter_cap=int(g["iter_cap"][0]))


class _GpuRunner:
    def __init__(self, ctx):
        self.ctx = ctx

    def run3(self, g):
        self.ctx.upload_shapes(g["shapes"], g.get("verts"))
        a = (g["l_shape"], g["r_shape"], g["l_t"], g["r_t"])
        return (lambda s: self.ctx.intersection(s, *a), lambda: self.ctx.distance(*a))

    def toi3(self, g):
        self.ctx.upload_shapes(g["shapes"], None)
        return self.ctx.time_of_impact(g["l_shape"], g["r_shape"], g["l_t0"], g["l_t1"], g["r_t0"], g["r_t1"],
                                       iter_cap=int(g["iter_cap"][0]))

    def sap(self, boxes, axis):
        return self.ctx.sap3(boxes, axis=axis)

    def run2(self, g):
        self.ctx.upload_shapes2(g["shapes"], g.get("verts"))
        a = (g["l_shape"], g["r_shape"], g["l_t"], g["r_t"])
        return (lambda s: self.ctx.intersection2(s, *a), lambda: self.ctx.distance2(*a))

    def toi2(self, g):
        self.ctx.upload_shapes2(g["shapes"], g.get("verts"))
        return self.ctx.time_of_impact2(g["l_shape"], g["r_shape"], g["l_t0"], g["l_t1"], g["r_t0"], g["r_t1"],
                                        iter_cap=int(g["iter_cap"][0]))


def _check_all(r):
    for name in ("cfg1_cuboids", "cfg2_mixed", "cfg3_polyhedra"):
        g = load(name)
        inter, dist = r.run3(g)
        assert same_records(*inter(0), g["full"], g["full_status"], CF), name
        if "only" in g:
            got, gst = inter(1)
            assert np.array_equal(gst, g["only_status"]) and got.tobytes() == g["only"].tobytes()
        assert same_floats(*dist(), g["dist"], g["dist_status"]), name
        assert 0.05 < (g["full_status"] == 1).mean() < 0.99
    g = load("cfg5_toi")
    assert same_records(*r.toi3(g), g["toi"], g["toi_status"], CF)
    g = load("cfg4_sap")
    for axis in (0, 2):
        perm, pairs, nxt = r.sap(g["aabbs"], axis)
        assert np.array_equal(perm, g[f"perm_axis{axis}"]) and np.array_equal(pairs, g[f"pairs_axis{axis}"])
        assert nxt == 0
    g = load("gjk2_mixed")
    inter, dist = r.run2(g)
    assert same_records(*inter(0), g["full"], g["full_status"], CF)
    assert same_floats(*dist(), g["dist"], g["dist_status"])
    g = load("gjk2_toi")
    assert same_records(*r.toi2(g), g["toi"], g["toi_status"], CF)


def test_oracle_reproduces_the_committed_fixtures():
    _check_all(_OracleRunner())


def test_fixture_dtypes_match_the_abi():
    from collision_rs_b200 import abi
    g = load("cfg3_polyhedra")
    assert g["shapes"].dtype == abi.SHAPE_DT and g["l_t"].dtype == abi.XFORM_DT
    assert g["full"].dtype == abi.CONTACT_DT
    g = load("gjk2_mixed")
    assert g["shapes"].dtype == abi.SHAPE2_DT and g["l_t"].dtype == abi.XFORM2_DT
    assert g["full"].dtype == abi.CONTACT2_DT


@pytest.mark.gpu
def test_cuda_path_reproduces_the_committed_fixtures(ctx):
    _check_all(_GpuRunner(ctx))
