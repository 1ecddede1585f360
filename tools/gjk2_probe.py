"""Times the 2-D narrow phase kernels (device-resident inputs, CUDA events) on one GPU."""
import sys
import numpy as np
import torch

sys.path.insert(0, ".")
from collision_rs_b200 import abi, workloads as W
from collision_rs_b200.host import Context

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
ctx = Context(0)
dev = torch.device("cuda:0")
dv = lambda a: torch.from_numpy(a.view(np.uint8).reshape(-1).copy()).to(dev)
s = torch.cuda.current_stream().cuda_stream


def timed(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for name, kinds, spread in (("mixed", (0, 1, 2, 3, 4, 5), 1.0), ("rectangles", (3, 4), 0.75),
                            ("polygons", (5,), 0.6), ("circles", (2,), 0.8)):
    w = W.cfg2d_mixed_pairs(n, seed=11, spread=spread, kinds=kinds)
    ctx.upload_shapes2(w["shapes"], w["verts"])
    d = {k: dv(w[k]) for k in ("l_shape", "r_shape", "l_t", "r_t")}
    out = torch.zeros(n * abi.CONTACT2_DT.itemsize, dtype=torch.uint8, device=dev)
    dist = torch.zeros(n, dtype=torch.float32, device=dev)
    st = torch.zeros(n, dtype=torch.uint8, device=dev)
    a = (d["l_shape"].data_ptr(), d["r_shape"].data_ptr(), d["l_t"].data_ptr(), d["r_t"].data_ptr(), n)
    ms_full = timed(lambda: ctx.intersection2_dev(0, *a, out.data_ptr(), st.data_ptr(), s))
    hits = float((st == 1).float().mean())
    ms_only = timed(lambda: ctx.intersection2_dev(1, *a, out.data_ptr(), st.data_ptr(), s))
    ms_dist = timed(lambda: ctx.distance2_dev(*a, dist.data_ptr(), st.data_ptr(), s))
    print(f"{name:11s} n={n} hits={hits:.3f}  FullResolution {ms_full:8.3f} ms ({n / ms_full / 1e3:8.1f} M pairs/s)"
          f"  CollisionOnly {ms_only:7.3f} ms ({n / ms_only / 1e3:8.1f} M/s)  distance {ms_dist:8.3f} ms"
          f" ({n / ms_dist / 1e3:8.1f} M/s)", flush=True)
w = W.cfg2d_toi_pairs(n, seed=12)
ctx.upload_shapes2(w["shapes"], w["verts"])
d = {k: dv(w[k]) for k in ("l_shape", "r_shape", "l_t0", "l_t1", "r_t0", "r_t1")}
out = torch.zeros(n * abi.CONTACT2_DT.itemsize, dtype=torch.uint8, device=dev)
st = torch.zeros(n, dtype=torch.uint8, device=dev)
ms = timed(lambda: ctx.time_of_impact2_dev(d["l_shape"].data_ptr(), d["r_shape"].data_ptr(), d["l_t0"].data_ptr(),
                                           d["l_t1"].data_ptr(), d["r_t0"].data_ptr(), d["r_t1"].data_ptr(), n,
                                           out.data_ptr(), st.data_ptr(), s))
print(f"toi mixed   n={n} hits={float((st == 1).float().mean()):.3f}  {ms:8.3f} ms ({n / ms / 1e3:8.1f} M pairs/s)")
