"""BASELINE.json's full sizes on the GPU, through properties that do not need the (slow) oracle:
determinism (two runs are byte-identical although work is distributed through atomic queues),
batch-split invariance (a pair's record does not depend on which other pairs share its batch),
CollisionOnly/FullResolution agreement on every Option, unit normals, sorted / truly-overlapping /
duplicate-free broad-phase output, and a sampled slice against the oracle."""
import numpy as np
import pytest

from collision_rs_b200 import abi, workloads as W

pytestmark = pytest.mark.gpu


def _dev(arr):
    import torch
    return torch.from_numpy(arr.view(np.uint8).reshape(-1).copy()).to("cuda:0")


def _run_intersection(ctx, strategy, d, n, lo=0, hi=None):
    import torch
    hi = n if hi is None else hi
    m = hi - lo
    out = torch.zeros(m * abi.CONTACT_DT.itemsize, dtype=torch.uint8, device="cuda:0")
    st = torch.zeros(m, dtype=torch.uint8, device="cuda:0")
    s = torch.cuda.current_stream().cuda_stream
    ctx.intersection_dev(strategy, d["l_shape"].data_ptr() + 4 * lo, d["r_shape"].data_ptr() + 4 * lo,
                         d["l_t"].data_ptr() + 32 * lo, d["r_t"].data_ptr() + 32 * lo, m, out.data_ptr(),
                         st.data_ptr(), s)
    torch.cuda.synchronize()
    return out, st


def test_cfg3_four_million_polyhedron_pairs(ctx, oracle):
    import torch
    n = 4_000_000
    w = W.cfg3_polyhedron_pairs(n)
    ctx.upload_shapes(w["shapes"], w["verts"])
    d = {k: _dev(w[k]) for k in ("l_shape", "r_shape", "l_t", "r_t")}
    out1, st1 = _run_intersection(ctx, 0, d, n)
    out2, st2 = _run_intersection(ctx, 0, d, n)
    assert torch.equal(st1, st2) and torch.equal(out1, out2)                      # deterministic
    cut = 1_500_001
    oa, sa = _run_intersection(ctx, 0, d, n, 0, cut)
    ob, sb = _run_intersection(ctx, 0, d, n, cut, n)
    assert torch.equal(torch.cat([sa, sb]), st1) and torch.equal(torch.cat([oa, ob]), out1)   # split invariant
    _, st_only = _run_intersection(ctx, 1, d, n)
    assert torch.equal(st_only, st1)                                              # same Option either way
    assert int((st1 > 1).sum()) == 0
    got = out1.cpu().numpy().view(abi.CONTACT_DT)
    gst = st1.cpu().numpy()
    hit = gst == 1
    assert 0.8 < hit.mean() < 0.9
    nrm = np.linalg.norm(got["normal"][hit].astype(np.float64), axis=1)
    ok = np.isfinite(nrm)
    assert ok.mean() > 0.999999 and np.allclose(nrm[ok], 1.0, atol=1e-5)
    assert (got["depth"][hit][ok] > -1e-4).all()
    sl = slice(2_345_678, 2_345_678 + 20_000)                                     # and a slice against the oracle
    exp, est = oracle.intersection(0, w["shapes"], w["verts"], w["l_shape"][sl], w["r_shape"][sl], w["l_t"][sl],
                                   w["r_t"][sl], nthreads=8)
    assert np.array_equal(gst[sl], est)
    m = est == 1
    for f in ("normal", "depth", "point"):
        assert np.array_equal(got[f][sl][m].view(np.uint32), exp[f][m].view(np.uint32)), f


def test_cfg5_eight_million_time_of_impact_pairs(ctx, oracle):
    import torch
    n = 8_000_000
    w = W.cfg5_toi_pairs(n)
    ctx.upload_shapes(w["shapes"])
    d = {k: _dev(w[k]) for k in ("l_shape", "r_shape", "l_t0", "l_t1", "r_t0", "r_t1")}
    s = torch.cuda.current_stream().cuda_stream

    def run(lo, hi):
        m = hi - lo
        out = torch.zeros(m * abi.CONTACT_DT.itemsize, dtype=torch.uint8, device="cuda:0")
        st = torch.zeros(m, dtype=torch.uint8, device="cuda:0")
        ctx.time_of_impact_dev(d["l_shape"].data_ptr() + 4 * lo, d["r_shape"].data_ptr() + 4 * lo,
                               d["l_t0"].data_ptr() + 32 * lo, d["l_t1"].data_ptr() + 32 * lo,
                               d["r_t0"].data_ptr() + 32 * lo, d["r_t1"].data_ptr() + 32 * lo, m, out.data_ptr(),
                               st.data_ptr(), s)
        torch.cuda.synchronize()
        return out, st

    out1, st1 = run(0, n)
    out2, st2 = run(0, n)
    assert torch.equal(st1, st2) and torch.equal(out1, out2)
    oa, sa = run(0, 3_000_001)
    ob, sb = run(3_000_001, n)
    assert torch.equal(torch.cat([sa, sb]), st1) and torch.equal(torch.cat([oa, ob]), out1)
    got = out1.cpu().numpy().view(abi.CONTACT_DT)
    gst = st1.cpu().numpy()
    hit = gst == 1
    assert 0.6 < hit.mean() < 0.9 and set(np.unique(gst)) <= {0, 1, 2, 3, 5}   # 2, 3: the reference's own panics
    toi = got["toi"][hit]
    assert (toi >= 0).all() and (toi <= 1).all()                                  # lambda is clipped to [0, 1]
    sl = slice(5_000_000, 5_020_000)
    exp, est = oracle.time_of_impact(w["shapes"], None, w["l_shape"][sl], w["r_shape"][sl], w["l_t0"][sl],
                                     w["l_t1"][sl], w["r_t0"][sl], w["r_t1"][sl], iter_cap=1024, nthreads=8)
    assert np.array_equal(gst[sl], est)
    m = est == 1
    for f in ("normal", "depth", "point", "toi"):
        a, b = got[f][sl][m], exp[f][m]
        assert ((a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b))).all(), f


def test_cfg4_one_million_boxes_pipeline(ctx):
    """SweepAndPrune3 on 1 M boxes: permutation sorts the keys, pairs are sorted (b, then a), unique,
    all truly overlap (strictly), and the chained narrow phase finds a contact for every pair (the
    bodies are the boxes themselves)."""
    import torch
    n = 1_000_000
    w = W.cfg4_aabbs(n)
    perm, pairs, axis = ctx.sap3(w["aabbs"], axis=0)
    assert axis == 0 and np.array_equal(np.sort(perm), np.arange(n, dtype=np.uint32))
    mn, mx = w["aabbs"]["min"][perm], w["aabbs"]["max"][perm]
    assert (np.diff(mn[:, 0]) >= 0).all()
    tie = np.diff(mn[:, 0]) == 0
    assert (np.diff(mx[:, 0])[tie] >= 0).all() and (np.diff(perm.astype(np.int64))[tie & (np.diff(mx[:, 0]) == 0)] > 0).all()
    a, b = pairs[:, 0].astype(np.int64), pairs[:, 1].astype(np.int64)
    assert (a < b).all()
    key = (b << 32) | a
    assert (np.diff(key) > 0).all()                                               # Vec order, no duplicates
    assert ((mx[a] > mn[b]) & (mn[a] < mx[b])).all()                              # Aabb3::intersects, strict
    # completeness on a sample: every overlapping partner of 300 random boxes is reported
    rng = np.random.default_rng(0)
    have = set(key.tolist())
    for i in rng.integers(0, n, 300):
        hit = np.nonzero(((mx[i] > mn) & (mn[i] < mx)).all(1))[0]
        for j in hit:
            if j != i:
                lo, hi = (i, j) if i < j else (j, i)
                assert ((int(hi) << 32) | int(lo)) in have
    # chained narrow phase on the device
    ctx.upload_shapes(w["shapes"])
    body_shape = torch.from_numpy(perm.astype(np.int32)).to("cuda:0")
    body_t = _dev(np.ascontiguousarray(w["body_t"][perm]))
    d_pairs = _dev(np.ascontiguousarray(pairs))
    m = len(pairs)
    out = torch.zeros(m * abi.CONTACT_DT.itemsize, dtype=torch.uint8, device="cuda:0")
    st = torch.zeros(m, dtype=torch.uint8, device="cuda:0")
    ctx.intersection_bodies_dev(0, body_shape.data_ptr(), body_t.data_ptr(), d_pairs.data_ptr(), m, out.data_ptr(),
                                st.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert int((st == 1).sum()) == m


def test_cfg4_one_million_boxes_against_the_oracle(ctx, oracle):
    """The whole pair list of cfg 4 — permutation and ~3.9 M pairs in the reference's Vec order —
    byte for byte against the oracle (its sweep split over the host threads: 1e10 box tests)."""
    n = 1_000_000
    b = W.cfg4_aabbs(n)["aabbs"]
    perm, pairs, axis = ctx.sap3(b, axis=0)
    perm_e, pairs_e, cnt, axis_e = oracle.sap3_mt(b, axis=0)
    assert cnt == len(pairs) and 3_900_000 < cnt < 4_000_000
    assert np.array_equal(perm, perm_e)
    assert np.array_equal(pairs, pairs_e)
    assert axis == axis_e == 0
