import os, sys
import numpy as np, torch
sys.path.insert(0, os.getcwd())
from collision_rs_b200 import workloads as W
from collision_rs_b200.abi import CONTACT_DT
from collision_rs_b200.host import Context
n = 1_000_000
w = W.cfg3_polyhedron_pairs(n)
dev = torch.device("cuda:0")
dv = lambda a: torch.from_numpy(a.view(np.uint8).reshape(-1).copy()).to(dev)
d = {k: dv(w[k]) for k in ("l_shape", "r_shape", "l_t", "r_t")}
out = torch.empty(n * CONTACT_DT.itemsize, dtype=torch.uint8, device=dev)
st = torch.empty(n, dtype=torch.uint8, device=dev)
s = torch.cuda.current_stream().cuda_stream
for legacy in ("0", "1"):
    os.environ["CB2_LEGACY_EPA"] = legacy
    ctx = Context(0)
    ctx.upload_shapes(w["shapes"], w["verts"])
    f = lambda: ctx.intersection_dev(0, d["l_shape"].data_ptr(), d["r_shape"].data_ptr(), d["l_t"].data_ptr(), d["r_t"].data_ptr(), n, out.data_ptr(), st.data_ptr(), s)
    for _ in range(2): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): f()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(f"legacy={legacy}: {ms:.3f} ms per 1M pairs -> {n/ms/1e3:.1f} M pairs/s", flush=True)
    ctx.close()
