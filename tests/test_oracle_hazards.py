"""Panic / hang regression vectors and parity-hazard checks
(Appendix A).  These inputs were predicted by the survey's scratch restatement, not by
running the Rust crate (impossible here): the oracle must classify them with the
documented non-OK status, and the CUDA kernels must agree with the oracle (test_gpu_*)."""
import numpy as np

from helpers import cuboid, sphere
from oracle.binding import XFORM_DT

H = float.fromhex


def xf(disp, quat):
    t = np.zeros(1, XFORM_DT)
    t["scale"] = 1.0
    t["q"] = [H(x) for x in quat]
    t["d"] = [H(x) for x in disp]
    return t


def dims(d):
    return [H(x) for x in d]


I0, I1 = np.array([0], np.uint32), np.array([1], np.uint32)

# B1: intersection_time_of_impact never terminates (gjk/mod.rs:204 has no cap)
B1 = [
    dict(ld=["0x1.8ff378p-1", "0x1.128098p+0", "0x1.091164p+0"],
         l0=["-0x1.5be462p+0", "0x1.49d336p+1", "0x1.7a0904p-4"],
         lq=["-0x1.3db56ep-1", "0x1.f48ab2p-2", "0x1.fa1e82p-2", "-0x1.73a5d2p-2"],
         l1=["0x1.a648d4p+1", "-0x1.1c6ed4p+2", "0x1.543606p+0"],
         rd=["0x1.3f251ap-1", "0x1.f089f0p+0", "0x1.e75f2ep+0"],
         r0=["0x1.34b072p+1", "-0x1.ed4a2ep+0", "0x1.59ac48p-1"],
         rq=["0x1.9a841cp-1", "0x1.4b95f4p-2", "0x1.4e78f8p-3", "-0x1.e66134p-2"],
         r1=["0x1.3a53bap+1", "-0x1.1e4a98p+0", "0x1.9c0198p+0"]),
    dict(ld=["0x1.b11df8p-1", "0x1.a605a8p-1", "0x1.ff9520p+0"],
         l0=["0x1.830e6ap-1", "-0x1.956c4cp-1", "-0x1.d9dda4p+0"],
         lq=["0x1.18c014p-1", "0x1.f9157cp-2", "-0x1.cb34c0p-2", "-0x1.0282fcp-1"],
         l1=["0x1.461044p+0", "0x1.3df638p-1", "-0x1.2dcd8cp-1"],
         rd=["0x1.a2102ap+0", "0x1.d7eacep-1", "0x1.08283ep+0"],
         r0=["0x1.38f82cp+1", "0x1.16b128p+1", "-0x1.41e6a6p-4"],
         rq=["-0x1.a5dd38p-2", "0x1.a00000p-1", "0x1.d51182p-3", "-0x1.5f3fecp-2"],
         r1=["0x1.3dac5ep+1", "0x1.7041acp+0", "-0x1.248562p-1"]),
]


def test_b1_toi_nonconverged(oracle):
    for c in B1:
        shapes = np.concatenate([cuboid(*dims(c["ld"])), cuboid(*dims(c["rd"]))])
        for cap in (64, 1024, 20000):
            out, st, it = oracle.time_of_impact(
                shapes, None, I0, I1, xf(c["l0"], c["lq"]), xf(c["l1"], c["lq"]),
                xf(c["r0"], c["rq"]), xf(c["r1"], c["rq"]), iter_cap=cap, want_iters=True)
            assert st[0] == 5 and it[0] == cap  # CB2_NONCONVERGED at any cap (limit cycle)


def test_b2_distance_assert_range(oracle):
    shapes = np.concatenate([sphere(H("0x1.b38cfap-1")), sphere(H("0x1.a94cf4p-1"))])
    lt = xf(["0x1.eeb542p-1", "-0x1.39f1bep-1", "0x1.5da646p-1"],
            ["-0x1.1937eap-1", "0x1.6e74c0p-2", "0x1.620826p-4", "-0x1.80186cp-1"])
    rt = xf(["0x1.19c170p-3", "-0x1.45ed10p-2", "0x1.25f46cp+0"],
            ["0x1.0b81a6p-2", "-0x1.2b66e8p-1", "0x1.6d93c8p-1", "0x1.21975ep-2"])
    d, st = oracle.distance(shapes, None, I0, I1, lt, rt)
    assert st[0] == 3  # CB2_PANIC_ASSERT_RANGE, simplex3d.rs:150


def test_b3_distance_unwrap_plane(oracle):
    cases = [
        (cuboid(*dims(["0x1.0d485ap-1", "0x1.8533aep+0", "0x1.043536p+0"])),
         xf(["-0x1.7b8216p-1", "-0x1.f585a6p-2", "-0x1.4156c0p+0"],
            ["0x1.18d876p-3", "0x1.fb2fe0p-2", "-0x1.b71d44p-1", "-0x1.241b1ap-6"]),
         sphere(H("0x1.19d93ep-2")),
         xf(["0x1.de5256p-1", "-0x1.3cf80ap-3", "0x1.5475d4p+0"],
            ["-0x1.24fc1ap-2", "0x1.d30012p-1", "0x1.2cab12p-3", "-0x1.0457dcp-2"])),
        (cuboid(*dims(["0x1.93b8e2p+0", "0x1.709a64p-1", "0x1.b50accp+0"])),
         xf(["-0x1.43cec4p+0", "0x1.7ea348p+0", "-0x1.798a2ap+0"],
            ["0x1.4d0bfep-1", "0x1.3fb6b6p-1", "0x1.f80710p-3", "0x1.6c074cp-2"]),
         cuboid(*dims(["0x1.7c7ed0p+0", "0x1.b26e92p+0", "0x1.36e52cp-1"])),
         xf(["-0x1.576652p+0", "-0x1.f2e99cp-5", "0x1.0e14e2p+0"],
            ["-0x1.c6b154p-2", "-0x1.a678f2p-1", "0x1.a56762p-3", "-0x1.20f832p-2"])),
    ]
    for l, lt, r, rt in cases:
        d, st = oracle.distance(np.concatenate([l, r]), None, I0, I1, lt, rt)
        assert st[0] == 2  # CB2_PANIC_UNWRAP_PLANE, simplex3d.rs:146


def test_zero_scale_is_inv_scale_panic(oracle):
    # inverse_transform_vector(..).unwrap() with ulps_eq(scale, 0), util.rs:18
    shapes = np.concatenate([cuboid(1, 1, 1), cuboid(1, 1, 1)])
    lt = xf(["0x0p+0"] * 3, ["0x1p+0", "0x0p+0", "0x0p+0", "0x0p+0"])
    rt = lt.copy()
    rt["scale"] = 1e-8
    out, st = oracle.intersection(0, shapes, None, I0, I1, lt, rt)
    assert st[0] == 4
    d, st = oracle.distance(shapes, None, I0, I1, lt, rt)
    assert st[0] == 4
    out, st = oracle.time_of_impact(shapes, None, I0, I1, lt, lt, rt, rt)
    assert st[0] == 4


def test_sphere_sphere_max_polytope(oracle):
    # Sphere(1) vs itself runs EPA to the 100-iteration cap: 103 vertices, 202 faces
    shapes = np.concatenate([sphere(1.0), sphere(1.0)])
    t = xf(["0x0p+0"] * 3, ["0x1p+0", "0x0p+0", "0x0p+0", "0x0p+0"])
    out, st, stats = oracle.intersection(0, shapes, None, I0, I1, t, t, want_stats=True)
    assert st[0] == 1
    # stats = {gjk its, epa its, peak faces, verts | peak horizon edges << 16}
    assert stats[0, 1] == 100 and stats[0, 2] == 202 and (stats[0, 3] & 0xFFFF) == 103


def test_cuboid_corner_tie_break_is_ordered_scan(oracle):
    # a sign-based shortcut is NOT bit-exact; the ordered 8-corner scan picks the
    # lower-index corner when rounded dots tie.
    s = cuboid(2 * 0.923, 2 * 0.883, 2 * 0.544)
    t = xf(["0x0p+0"] * 3, ["0x1p+0", "0x0p+0", "0x0p+0", "0x0p+0"])
    p, st = oracle.support_point(s, None, t, (-1.163, -1.36e-9, -0.488))
    h = s["p"][0] / np.float32(2)
    assert tuple(p) == (-h[0], h[1], -h[2])  # corner (-,+,-), not (-,-,-)


def test_sap_order_and_axis(oracle):
    # output order is (j asc, i asc) over SORTED positions; ties on min fall back to max;
    # stable for full ties; axis is used for the sort, then reset to 0 (F7)
    from helpers import aabb
    boxes = np.concatenate([
        aabb((5, 0, 0), (9, 9, 9)),     # 0
        aabb((0, 5, 0), (9, 9, 9)),     # 1
        aabb((0, 1, 0), (8, 9, 9)),     # 2
        aabb((0, 1, 0), (8, 9, 9)),     # 3 (full tie with 2 on axis 0 -> stable)
        aabb((-0.0, 0, 0), (7, 9, 9)),  # 4 (-0.0 == +0.0 under partial_cmp)
    ])
    perm, pairs, cnt, axis = oracle.sap3(boxes, axis=0)
    assert perm.tolist() == [4, 2, 3, 1, 0]
    assert cnt == 10 and axis == 0
    assert [tuple(p) for p in pairs.tolist()] == [
        (0, 1), (0, 2), (1, 2), (0, 3), (1, 3), (2, 3), (0, 4), (1, 4), (2, 4), (3, 4)]
    perm, pairs, cnt, axis = oracle.sap3(boxes, axis=1)
    assert perm.tolist() == [0, 4, 2, 3, 1] and axis == 0
    # n <= 1: nothing sorted, axis unchanged (sweep_prune.rs:83-85)
    perm, pairs, cnt, axis = oracle.sap3(boxes[:1], axis=2)
    assert cnt == 0 and axis == 2 and perm.tolist() == [0]
