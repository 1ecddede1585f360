"""One untimed pass of every BASELINE config at its full size, for ncu (few launches, each kernel of
the hot path once or twice): cfg 3 (4 M polyhedron pairs, FullResolution), cfg 2 (1 M mixed pairs:
CollisionOnly + distance), cfg 5 (8 M time-of-impact pairs), cfg 4 (1 M boxes: SAP, then the narrow
phase on its pairs), the 2-D narrow phase (4 M mixed Primitive2 pairs).
    ncu --metrics ... python tools/profile_pass.py [cfg3 cfg2 cfg5 cfg4 gjk2]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from collision_rs_b200 import workloads as W  # noqa: E402
from collision_rs_b200.abi import CONTACT2_DT, CONTACT_DT  # noqa: E402
from collision_rs_b200.host import Context  # noqa: E402

which = sys.argv[1:] or ["cfg3", "cfg2", "cfg5", "cfg4", "gjk2"]
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
ctx = Context(0)
s = torch.cuda.current_stream().cuda_stream


def mark():
    """section marker in the launch stream (profiles/summarize_r2.py splits on it)"""
    torch.full((1,), 1.0, device=dev)


def dv(arr):
    return torch.from_numpy(arr.view(np.uint8).reshape(-1).copy()).to(dev)


def bufs(n, rec=CONTACT_DT.itemsize):
    return (torch.empty(n * rec, dtype=torch.uint8, device=dev), torch.empty(n, dtype=torch.uint8, device=dev))


if "cfg3" in which:
    mark()
    n = int(os.environ.get("CB2_PROFILE_PAIRS", 4_000_000))
    w = W.cfg3_polyhedron_pairs(n)
    ctx.upload_shapes(w["shapes"], w["verts"])
    d = {k: dv(w[k]) for k in ("l_shape", "r_shape", "l_t", "r_t")}
    out, st = bufs(n)
    for _ in range(2):  # the second pass is the one to read (workspaces exist, tables are warm)
        ctx.intersection_dev(0, d["l_shape"].data_ptr(), d["r_shape"].data_ptr(), d["l_t"].data_ptr(),
                             d["r_t"].data_ptr(), n, out.data_ptr(), st.data_ptr(), s)
    torch.cuda.synchronize()
if "cfg2" in which:
    mark()
    n = 1_000_000
    w = W.cfg2_mixed_pairs(n)
    ctx.upload_shapes(w["shapes"])
    d = {k: dv(w[k]) for k in ("l_shape", "r_shape", "l_t", "r_t")}
    out, st = bufs(n)
    dist = torch.empty(n, dtype=torch.float32, device=dev)
    for _ in range(2):
        ctx.intersection_dev(1, d["l_shape"].data_ptr(), d["r_shape"].data_ptr(), d["l_t"].data_ptr(),
                             d["r_t"].data_ptr(), n, out.data_ptr(), st.data_ptr(), s)
        ctx.distance_dev(d["l_shape"].data_ptr(), d["r_shape"].data_ptr(), d["l_t"].data_ptr(),
                         d["r_t"].data_ptr(), n, dist.data_ptr(), st.data_ptr(), s)
    torch.cuda.synchronize()
if "cfg5" in which:
    mark()
    n = 8_000_000
    w = W.cfg5_toi_pairs(n)
    ctx.upload_shapes(w["shapes"])
    d = {k: dv(w[k]) for k in ("l_shape", "r_shape", "l_t0", "l_t1", "r_t0", "r_t1")}
    out, st = bufs(n)
    for _ in range(2):
        ctx.time_of_impact_dev(d["l_shape"].data_ptr(), d["r_shape"].data_ptr(), d["l_t0"].data_ptr(),
                               d["l_t1"].data_ptr(), d["r_t0"].data_ptr(), d["r_t1"].data_ptr(), n,
                               out.data_ptr(), st.data_ptr(), s)
    torch.cuda.synchronize()
if "cfg4" in which:
    mark()
    n = 1_000_000
    w = W.cfg4_aabbs(n)
    ctx.upload_shapes(w["shapes"])
    boxes = dv(w["aabbs"])
    perm = torch.empty(n, dtype=torch.int32, device=dev)
    cap = 8 * n
    pairs = torch.empty(2 * cap, dtype=torch.int32, device=dev)
    for _ in range(2):
        rc, npairs = ctx.sap3_dev(boxes.data_ptr(), n, 0, perm.data_ptr(), pairs.data_ptr(), cap, s)
    p64 = perm.long()
    body_shape = dv(np.arange(n, dtype=np.uint32)).view(torch.int32)[p64].contiguous()
    body_t = dv(w["body_t"]).view(-1, 32)[p64].contiguous()
    out, st = bufs(npairs)
    for _ in range(2):
        ctx.intersection_bodies_dev(0, body_shape.data_ptr(), body_t.data_ptr(), pairs.data_ptr(), npairs,
                                    out.data_ptr(), st.data_ptr(), s)
    torch.cuda.synchronize()
if "gjk2" in which:
    mark()
    n = 4_000_000
    w = W.cfg2d_mixed_pairs(n, seed=11, spread=1.0, kinds=(0, 1, 2, 3, 4, 5))
    ctx.upload_shapes2(w["shapes"], w["verts"])
    d = {k: dv(w[k]) for k in ("l_shape", "r_shape", "l_t", "r_t")}
    out, st = bufs(n, CONTACT2_DT.itemsize)
    for _ in range(2):
        ctx.intersection2_dev(0, d["l_shape"].data_ptr(), d["r_shape"].data_ptr(), d["l_t"].data_ptr(),
                              d["r_t"].data_ptr(), n, out.data_ptr(), st.data_ptr(), s)
    torch.cuda.synchronize()
print("profile pass done:", which)
