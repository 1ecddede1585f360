"""Generates the committed fixtures in tests/golden/*.npz.

The reference is a Rust crate and cannot run in this image (no rustc / cargo), so these fixtures are
NOT outputs of the reference itself: they are outputs of the CPU oracle (oracle/*.cpp) at the commit
where it is pinned by every known-answer test the reference holds for the path
(tests/test_oracle_golden.py, tests/test_oracle2d_golden.py, tests/test_oracle_hazards.py).  They
freeze that pinned behaviour: tests/test_golden_fixtures.py checks that the oracle still reproduces
them bit for bit (no drift through compiler / flag / refactoring changes) and, on the GPU box, that
the CUDA path does too — without depending on anything outside the repository.

Inputs AND outputs are stored (small: 1-2 k records per case), so the fixtures do not depend on
numpy's generator staying stable either.   Run:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from collision_rs_b200 import workloads as W  # noqa: E402
from oracle.binding import Oracle, Oracle2D  # noqa: E402


def save(name, **arrays):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **{k: v for k, v in arrays.items() if v is not None})
    print(f"{name}: {os.path.getsize(path) / 1024:.0f} KiB")


def main():
    o, o2 = Oracle(), Oracle2D()
    # cfg 1: cuboid pairs, FullResolution + CollisionOnly + distance
    w = W.cfg1_cuboid_pairs(1500, spread=1.0)
    a = (w["l_shape"], w["r_shape"], w["l_t"], w["r_t"])
    c0, s0 = o.intersection(0, w["shapes"], None, *a)
    c1, s1 = o.intersection(1, w["shapes"], None, *a)
    d, sd = o.distance(w["shapes"], None, *a)
    save("cfg1_cuboids", **w, full=c0, full_status=s0, only=c1, only_status=s1, dist=d, dist_status=sd)
    # cfg 2: mixed Sphere/Cuboid (curved shapes: EPA to the cap, the distance panics)
    w = W.cfg2_mixed_pairs(1500)
    a = (w["l_shape"], w["r_shape"], w["l_t"], w["r_t"])
    c0, s0 = o.intersection(0, w["shapes"], None, *a)
    d, sd = o.distance(w["shapes"], None, *a)
    save("cfg2_mixed", **w, full=c0, full_status=s0, dist=d, dist_status=sd)
    # cfg 3: polyhedron pairs over a small table
    w = W.cfg3_polyhedron_pairs(1500, n_shapes=48)
    a = (w["l_shape"], w["r_shape"], w["l_t"], w["r_t"])
    c0, s0 = o.intersection(0, w["shapes"], w["verts"], *a)
    d, sd = o.distance(w["shapes"], w["verts"], *a)
    save("cfg3_polyhedra", **w, full=c0, full_status=s0, dist=d, dist_status=sd)
    # cfg 4: SweepAndPrune3 on 2000 boxes, axis 0 and 2
    w = W.cfg4_aabbs(2000, extent=200.0 * (2000 / 1e6) ** (1 / 3))
    p0, pr0, _, ax0 = o.sap3(w["aabbs"], 0)
    p2, pr2, _, ax2 = o.sap3(w["aabbs"], 2)
    save("cfg4_sap", aabbs=w["aabbs"], perm_axis0=p0, pairs_axis0=pr0, perm_axis2=p2, pairs_axis2=pr2,
         next_axis=np.array([ax0, ax2], np.uint32))
    # cfg 5: time of impact
    w = W.cfg5_toi_pairs(1500)
    c, s = o.time_of_impact(w["shapes"], None, w["l_shape"], w["r_shape"], w["l_t0"], w["l_t1"], w["r_t0"],
                            w["r_t1"], iter_cap=256)
    save("cfg5_toi", **w, toi=c, toi_status=s, iter_cap=np.array([256], np.uint32))
    # 2-D: every Primitive2 variant
    w = W.cfg2d_mixed_pairs(2000, seed=901, n_shapes=96)
    a = (w["l_shape"], w["r_shape"], w["l_t"], w["r_t"])
    c0, s0 = o2.intersection(0, w["shapes"], w["verts"], *a)
    d, sd = o2.distance(w["shapes"], w["verts"], *a)
    wt = W.cfg2d_toi_pairs(2000, seed=902, n_shapes=96)
    ct, st = o2.time_of_impact(wt["shapes"], wt["verts"], wt["l_shape"], wt["r_shape"], wt["l_t0"], wt["l_t1"],
                               wt["r_t0"], wt["r_t1"], iter_cap=256)
    save("gjk2_mixed", **w, full=c0, full_status=s0, dist=d, dist_status=sd)
    save("gjk2_toi", **wt, toi=ct, toi_status=st, iter_cap=np.array([256], np.uint32))


if __name__ == "__main__":
    main()
