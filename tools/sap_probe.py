"""SAP-only loop for profiling: cfg4 (1M boxes), device-resident."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.getcwd())
from collision_rs_b200 import workloads as W
from collision_rs_b200.host import Context
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
ctx = Context(0)
w = W.cfg4_aabbs(n, extent=200.0 * (n / 1e6) ** (1 / 3))  # cfg 4's density at any n
dev = torch.device("cuda", 0)
boxes = torch.from_numpy(w["aabbs"].view(np.uint8).reshape(-1).copy()).to(dev)
perm = torch.empty(n, dtype=torch.int32, device=dev)
cap = 8 * n
pairs = torch.empty(2 * cap, dtype=torch.int32, device=dev)
s = torch.cuda.current_stream().cuda_stream
for _ in range(3):
    rc, cnt = ctx.sap3_dev(boxes.data_ptr(), n, 0, perm.data_ptr(), pairs.data_ptr(), cap, s)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    rc, cnt = ctx.sap3_dev(boxes.data_ptr(), n, 0, perm.data_ptr(), pairs.data_ptr(), cap, s)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print(f"n={n} pairs={cnt} {ms:.3f} ms/call -> {n / ms / 1e3:.1f} M AABBs/s (ordered Vec)")
e0.record()
for _ in range(reps):
    rc, cnt2 = ctx.sap3_unordered_dev(boxes.data_ptr(), n, 0, perm.data_ptr(), pairs.data_ptr(), cap, s)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print(f"n={n} pairs={cnt2} {ms:.3f} ms/call -> {n / ms / 1e3:.1f} M AABBs/s (pair set)")
