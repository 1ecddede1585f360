"""FullResolution on cfg 2's mixed Sphere / Cuboid pairs (curved pairs start in the last EPA tier),
device-resident: python tools/curved_probe.py [n]"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.getcwd())
from collision_rs_b200 import workloads as W
from collision_rs_b200.abi import CONTACT_DT
from collision_rs_b200.host import Context
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
w = W.cfg2_mixed_pairs(n)
dev = torch.device("cuda", 0)
ctx = Context(0)
ctx.upload_shapes(w["shapes"], w.get("verts"))
def dv(a): return torch.from_numpy(a.view(np.uint8).reshape(-1).copy()).to(dev)
d = {k: dv(w[k]) for k in ("l_shape", "r_shape", "l_t", "r_t")}
out = torch.empty(n * CONTACT_DT.itemsize, dtype=torch.uint8, device=dev)
st = torch.empty(n, dtype=torch.uint8, device=dev)
s = torch.cuda.current_stream().cuda_stream
def step():
    ctx.intersection_dev(0, d["l_shape"].data_ptr(), d["r_shape"].data_ptr(), d["l_t"].data_ptr(), d["r_t"].data_ptr(), n,
                         out.data_ptr(), st.data_ptr(), stream=s)
for _ in range(3): step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): step()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
sts = st.cpu().numpy()
print(f"{os.environ.get('CB2_LIB_PATH','default').split('_')[-1]}: cfg2 mixed FullResolution {n} pairs {ms:.3f} ms -> {n/ms/1e3:.1f} M pairs/s, hits {int((sts==1).sum())}, status7 {int((sts==7).sum())}")
