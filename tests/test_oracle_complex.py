"""Oracle checks for the composite-shape callers (gjk/mod.rs:412-588).  The reference has no
3-D test for them, so the pin is structural: the oracle's *_complex functions must equal an
independent numpy re-expression of the double loop — Transform::concat done in numpy f32, the
single-primitive oracle routines on the expanded part pairs, and the reference's early-return /
max_by / min_by / f32::min rules applied in Python."""
import numpy as np

from collision_rs_b200 import workloads as W
from collision_rs_b200.abi import XFORM_DT
from helpers_complex import composite_workload, concat_np, reduce_reference


def test_concat_identity_and_translation(oracle):
    # concat with the identity is the identity; pure translations add
    rng = np.random.default_rng(1)
    a = W.random_xforms(rng, 50, 2.0)
    ident = np.zeros(50, XFORM_DT)
    ident["scale"] = 1
    ident["q"][:, 0] = 1
    c = concat_np(a, ident)
    assert np.array_equal(c["q"], a["q"]) and np.array_equal(c["d"], a["d"])
    t1, t2 = ident.copy(), ident.copy()
    t1["d"] = rng.uniform(-1, 1, (50, 3)).astype(np.float32)
    t2["d"] = rng.uniform(-1, 1, (50, 3)).astype(np.float32)
    assert np.array_equal(concat_np(t1, t2)["d"], t2["d"] + t1["d"])


def test_complex_equals_expanded_double_loop(oracle):
    cw = composite_workload(4000, seed=3)
    for op, strategy in ((0, 0), (0, 1), (1, 0), (2, 0), (2, 1)):
        got, gst = oracle.complex(op, strategy, cw["shapes"], cw["verts"], cw["comp_off"],
                                  cw["comp_shape"], cw["comp_local"], cw["l_comp"], cw["r_comp"],
                                  cw["l_t0"], cw["r_t0"], cw["l_t1"], cw["r_t1"], nthreads=4)
        exp, est = reduce_reference(oracle, op, strategy, cw)
        assert np.array_equal(gst, est), (op, strategy, np.nonzero(gst != est)[0][:5])
        m = est == 1
        assert m.sum() > 50
        if op == 1:
            assert np.array_equal(got[m].view(np.uint32), exp[m].view(np.uint32))
        else:
            assert got[m].tobytes() == exp[m].tobytes()


def test_2d_complex_equals_expanded_double_loop():
    """The same structural pin for the 2-D instantiation (Decomposed<Vector2, Basis2>::concat)."""
    from collision_rs_b200.abi import CONTACT2_DT
    from helpers_complex import composite_workload2, concat2_np
    from oracle.binding import Oracle2D
    o2 = Oracle2D()
    cw = composite_workload2(4000, seed=4)
    for op, strategy in ((0, 0), (0, 1), (1, 0), (2, 0), (2, 1)):
        got, gst = o2.complex(op, strategy, cw["shapes"], cw["verts"], cw["comp_off"], cw["comp_shape"],
                              cw["comp_local"], cw["l_comp"], cw["r_comp"], cw["l_t0"], cw["r_t0"], cw["l_t1"],
                              cw["r_t1"], nthreads=4)
        exp, est = reduce_reference(o2, op, strategy, cw, concat2_np, CONTACT2_DT)
        assert np.array_equal(gst, est), (op, strategy, np.nonzero(gst != est)[0][:5])
        m = est == 1
        assert m.sum() > 50
        if op == 1:
            assert np.array_equal(got[m].view(np.uint32), exp[m].view(np.uint32))
        else:
            assert got[m].tobytes() == exp[m].tobytes()
