"""One traced host-buffer call (CB2_TRACE=1): prints the per-chunk timeline of cb2_gjk3_intersection."""
import os, sys, time, ctypes as C
os.environ["CB2_TRACE"] = "1"
import numpy as np, torch
sys.path.insert(0, os.getcwd())
from collision_rs_b200 import workloads as W
from collision_rs_b200.abi import CONTACT_DT, Settings
from collision_rs_b200.host import Context
n = 4_000_000
w = W.cfg3_polyhedron_pairs(n)
pinned = lambda arr: torch.from_numpy(arr.view(np.uint8).reshape(-1)).pin_memory()
h_in = {k: pinned(w[k]) for k in ("l_shape", "r_shape", "l_t", "r_t")}
h_out = torch.empty(n * CONTACT_DT.itemsize, dtype=torch.uint8).pin_memory()
h_st = torch.empty(n, dtype=torch.uint8).pin_memory()
S = Settings.default()
ctx = Context(0)
ctx.upload_shapes(w["shapes"], w["verts"])
for i in range(3):
    print(f"--- call {i}", file=sys.stderr, flush=True)
    t = time.perf_counter()
    ctx._ck(ctx.lib.cb2_gjk3_intersection(ctx.h, 0, C.byref(S), h_in["l_shape"].data_ptr(), h_in["r_shape"].data_ptr(),
            h_in["l_t"].data_ptr(), h_in["r_t"].data_ptr(), n, h_out.data_ptr(), h_st.data_ptr()))
    print(f"wall {1e3 * (time.perf_counter() - t):.2f} ms", file=sys.stderr, flush=True)
