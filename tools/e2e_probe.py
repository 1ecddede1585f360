"""End-to-end (host buffers) timing of cb2_gjk3_intersection on cfg 3 for a few chunk sizes."""
import os, sys, time, ctypes as C
import numpy as np, torch
sys.path.insert(0, os.getcwd())
from collision_rs_b200 import workloads as W
from collision_rs_b200.abi import CONTACT_DT, Settings
n = 4_000_000
w = W.cfg3_polyhedron_pairs(n)
def pinned(arr):
    return torch.from_numpy(arr.view(np.uint8).reshape(-1)).pin_memory()
h_in = {k: pinned(w[k]) for k in ("l_shape", "r_shape", "l_t", "r_t")}
h_out = torch.empty(n * CONTACT_DT.itemsize, dtype=torch.uint8).pin_memory()
h_st = torch.empty(n, dtype=torch.uint8).pin_memory()
S = Settings.default()
for chunk in [int(x) for x in sys.argv[1:]] or [1 << 19]:
    os.environ["CB2_CHUNK_PAIRS"] = str(chunk)
    from collision_rs_b200.host import Context
    ctx = Context(0)
    ctx.upload_shapes(w["shapes"], w["verts"])
    def step():
        ctx._ck(ctx.lib.cb2_gjk3_intersection(ctx.h, 0, C.byref(S), h_in["l_shape"].data_ptr(), h_in["r_shape"].data_ptr(),
            h_in["l_t"].data_ptr(), h_in["r_t"].data_ptr(), n, h_out.data_ptr(), h_st.data_ptr()))
    step(); step()
    t = time.perf_counter()
    for _ in range(4): step()
    dt = (time.perf_counter() - t) / 4
    print(f"{os.environ.get('CB2_LIB_PATH','default').split('_')[-1]} chunk {chunk}: {dt*1e3:.1f} ms/step -> {n/dt/1e6:.1f} M pairs/s", flush=True)
    ctx.close()
