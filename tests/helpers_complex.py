"""Independent re-expression of the composite-shape loops for the oracle tests, and the shared
composite workload for the GPU parity tests."""
import numpy as np

from collision_rs_b200 import workloads as W
from collision_rs_b200.abi import CONTACT_DT, XFORM_DT

f32 = np.float32


def _cross(a, b):
    return np.stack([a[:, 1] * b[:, 2] - a[:, 2] * b[:, 1], a[:, 2] * b[:, 0] - a[:, 0] * b[:, 2],
                     a[:, 0] * b[:, 1] - a[:, 1] * b[:, 0]], axis=1).astype(f32)


def _rotate(q, x):
    qv, s = q[:, 1:], q[:, :1]
    tmp = (_cross(qv, x) + (x * s).astype(f32)).astype(f32)
    return ((_cross(qv, tmp) * f32(2)).astype(f32) + x).astype(f32)


def concat_np(a, b):
    """Decomposed::concat in numpy f32, one operation at a time (no fused multiply-add)."""
    out = np.zeros(len(a), XFORM_DT)
    out["scale"] = (a["scale"] * b["scale"]).astype(f32)
    aq, bq = a["q"], b["q"]
    s, x, y, z = aq[:, 0], aq[:, 1], aq[:, 2], aq[:, 3]
    S, X, Y, Z = bq[:, 0], bq[:, 1], bq[:, 2], bq[:, 3]
    m = lambda p, q: (p * q).astype(f32)
    out["q"][:, 0] = ((m(s, S) - m(x, X)).astype(f32) - m(y, Y)).astype(f32) - m(z, Z)
    out["q"][:, 1] = ((m(s, X) + m(x, S)).astype(f32) + m(y, Z)).astype(f32) - m(z, Y)
    out["q"][:, 2] = ((m(s, Y) + m(y, S)).astype(f32) + m(z, X)).astype(f32) - m(x, Z)
    out["q"][:, 3] = ((m(s, Z) + m(z, S)).astype(f32) + m(x, Y)).astype(f32) - m(y, X)
    out["d"] = (_rotate(aq, (b["d"] * a["scale"][:, None]).astype(f32)) + a["d"]).astype(f32)
    return out


def composite_workload(n, seed=0, n_comp=200, spread=1.2):
    """Composites of 1-4 parts (spheres, cuboids, small polyhedra) with random local transforms
    (some non-unit scales), composite pairs with start/end model transforms."""
    rng = np.random.default_rng(seed)
    ps, pv = W.polyhedron_table(24, seed=seed + 100, sizes=(12, 20, 40))
    prim = W.mixed_shapes(rng, 40)
    shapes = np.concatenate([ps, prim])
    counts = rng.integers(1, 5, n_comp)
    comp_off = np.zeros(n_comp + 1, np.uint32)
    comp_off[1:] = np.cumsum(counts)
    n_parts = int(comp_off[-1])
    comp_shape = rng.integers(0, len(shapes), n_parts).astype(np.uint32)
    comp_local = W.random_xforms(rng, n_parts, 0.6)
    comp_local["scale"][::7] = rng.uniform(0.6, 1.4, len(comp_local["scale"][::7])).astype(f32)
    l_comp = rng.integers(0, n_comp, n).astype(np.uint32)
    r_comp = rng.integers(0, n_comp, n).astype(np.uint32)
    l_t0, r_t0 = W.random_xforms(rng, n, spread), W.random_xforms(rng, n, spread)
    l_t1, r_t1 = l_t0.copy(), r_t0.copy()
    l_t1["d"] += rng.uniform(-1, 1, (n, 3)).astype(f32)
    r_t1["d"] += ((l_t0["d"] - r_t0["d"]) * rng.uniform(0.5, 1.5, (n, 1))).astype(f32)
    return dict(shapes=shapes, verts=pv, comp_off=comp_off, comp_shape=comp_shape,
                comp_local=comp_local, l_comp=l_comp, r_comp=r_comp, l_t0=l_t0, r_t0=r_t0,
                l_t1=l_t1, r_t1=r_t1)


def expand(cw):
    """Flat part pairs in the reference's loop order + the owning composite pair of each."""
    off = cw["comp_off"].astype(np.int64)
    own, la, rb = [], [], []
    for i, (lc, rc) in enumerate(zip(cw["l_comp"], cw["r_comp"])):
        for a in range(off[lc], off[lc + 1]):
            for b in range(off[rc], off[rc + 1]):
                own.append(i)
                la.append(a)
                rb.append(b)
    return np.asarray(own), np.asarray(la), np.asarray(rb)


def concat2_np(a, b):
    """Decomposed<Vector2, Basis2>::concat in numpy f32: scale product, Matrix2 product column by
    column (row-dot matrix-vector products), rot.rotate(other.disp * scale) + disp."""
    from collision_rs_b200.abi import XFORM2_DT
    out = np.zeros(len(a), XFORM2_DT)
    m = lambda p, q: (p * q).astype(f32)
    am = a["m"]

    def matvec(vx, vy):
        return ((m(am[:, 0], vx) + m(am[:, 2], vy)).astype(f32), (m(am[:, 1], vx) + m(am[:, 3], vy)).astype(f32))

    out["scale"] = m(a["scale"], b["scale"])
    out["m"][:, 0], out["m"][:, 1] = matvec(b["m"][:, 0], b["m"][:, 1])
    out["m"][:, 2], out["m"][:, 3] = matvec(b["m"][:, 2], b["m"][:, 3])
    dx, dy = matvec(m(b["d"][:, 0], a["scale"]), m(b["d"][:, 1], a["scale"]))
    out["d"][:, 0] = (dx + a["d"][:, 0]).astype(f32)
    out["d"][:, 1] = (dy + a["d"][:, 1]).astype(f32)
    return out


def composite_workload2(n, seed=0, n_comp=200, spread=1.2):
    """2-D composites of 1-4 parts over every Primitive2 kind, random local transforms (some
    non-unit scales), composite pairs with start/end model transforms."""
    rng = np.random.default_rng(seed)
    shapes, verts = W.shapes2_table(64, seed + 100)
    counts = rng.integers(1, 5, n_comp)
    comp_off = np.zeros(n_comp + 1, np.uint32)
    comp_off[1:] = np.cumsum(counts)
    n_parts = int(comp_off[-1])
    comp_shape = rng.integers(0, len(shapes), n_parts).astype(np.uint32)
    comp_local = W.random_xforms2(rng, n_parts, 0.6)
    comp_local["scale"][::7] = rng.uniform(0.6, 1.4, len(comp_local["scale"][::7])).astype(f32)
    l_comp = rng.integers(0, n_comp, n).astype(np.uint32)
    r_comp = rng.integers(0, n_comp, n).astype(np.uint32)
    l_t0, r_t0 = W.random_xforms2(rng, n, spread), W.random_xforms2(rng, n, spread)
    l_t1, r_t1 = l_t0.copy(), r_t0.copy()
    l_t1["d"] += rng.uniform(-1, 1, (n, 2)).astype(f32)
    r_t1["d"] += ((l_t0["d"] - r_t0["d"]) * rng.uniform(0.5, 1.5, (n, 1))).astype(f32)
    return dict(shapes=shapes, verts=verts, comp_off=comp_off, comp_shape=comp_shape, comp_local=comp_local,
                l_comp=l_comp, r_comp=r_comp, l_t0=l_t0, r_t0=r_t0, l_t1=l_t1, r_t1=r_t1)


def reduce_reference(oracle, op, strategy, cw, concat_np=concat_np, CONTACT_DT=CONTACT_DT):
    own, la, rb = expand(cw)
    ls, rs = cw["comp_shape"][la], cw["comp_shape"][rb]
    lt0 = concat_np(cw["l_t0"][own], cw["comp_local"][la])
    rt0 = concat_np(cw["r_t0"][own], cw["comp_local"][rb])
    n = len(cw["l_comp"])
    if op == 1:
        sub, sst = oracle.distance(cw["shapes"], cw["verts"], ls, rs, lt0, rt0, nthreads=4)
        out, st = np.zeros(n, f32), np.zeros(n, np.uint8)
    else:
        if op == 0:
            sub, sst = oracle.intersection(strategy, cw["shapes"], cw["verts"], ls, rs, lt0, rt0,
                                           nthreads=4)
        else:
            lt1 = concat_np(cw["l_t1"][own], cw["comp_local"][la])
            rt1 = concat_np(cw["r_t1"][own], cw["comp_local"][rb])
            sub, sst = oracle.time_of_impact(cw["shapes"], cw["verts"], ls, rs, lt0, lt1, rt0, rt1,
                                             nthreads=4)
        out, st = np.zeros(n, CONTACT_DT), np.zeros(n, np.uint8)
        out["strategy"] = strategy
    starts = np.searchsorted(own, np.arange(n))
    ends = np.searchsorted(own, np.arange(n), side="right")
    for i in range(n):
        have = False
        for s in range(starts[i], ends[i]):
            ss = sst[s]
            if op == 1:
                if ss != 1:
                    st[i], have = ss, False
                    break
                out[i] = sub[s] if not have else (sub[s] if (sub[s] < out[i] or out[i] != out[i]) else out[i])
                have, st[i] = True, 1
                continue
            if ss == 0:
                continue
            if ss != 1:
                st[i], have = ss, False
                break
            if strategy == 1:
                out[i], have, st[i] = sub[s], True, 1
                out[i]["strategy"] = 1
                break
            if not have:
                out[i], have, st[i] = sub[s], True, 1
            elif op == 0:
                a, b = out[i]["depth"], sub[s]["depth"]
                if a != a or b != b:
                    st[i], have = 8, False
                    break
                if not a > b:
                    out[i] = sub[s]
            else:
                if out[i]["toi"] > sub[s]["toi"]:
                    out[i] = sub[s]
        if not have:
            out[i] = 0
            if op != 1:
                out[i]["strategy"] = strategy
    return out, st
