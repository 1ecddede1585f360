"""Time one GPU's shard of SweepAndPrune3 for several shard counts (all on cuda:0; shard in the middle)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.getcwd())
from collision_rs_b200 import workloads as W
from collision_rs_b200.host import Context
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
ctx = Context(0)
w = W.cfg4_aabbs(n, extent=200.0 * (n / 1e6) ** (1 / 3))
dev = torch.device("cuda", 0)
boxes = torch.from_numpy(w["aabbs"].view(np.uint8).reshape(-1).copy()).to(dev)
perm = torch.empty(n, dtype=torch.int32, device=dev)
cap = 8 * n
pairs = torch.empty(2 * cap, dtype=torch.int32, device=dev)
s = torch.cuda.current_stream().cuda_stream
for shards in (1, 2, 4, 8):
    tot = 0
    worst = 0.0
    for shard in range(shards):
        for _ in range(2):
            rc, cnt = ctx.sap3_shard_dev(boxes.data_ptr(), n, 0, shard, shards, perm.data_ptr(), pairs.data_ptr(), cap, s)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            rc, cnt = ctx.sap3_shard_dev(boxes.data_ptr(), n, 0, shard, shards, perm.data_ptr(), pairs.data_ptr(), cap, s)
        e1.record()
        torch.cuda.synchronize()
        worst = max(worst, e0.elapsed_time(e1) / 5)
        tot += cnt
    print(f"n={n} shards={shards}: slowest shard {worst:.3f} ms, pairs over all shards {tot}")
