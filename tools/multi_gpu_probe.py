"""Multi-GPU broad phase under torchrun: rank 0 broadcasts the cfg-4 AABB table over
NCCL, every rank computes its shard of SweepAndPrune3::find_collider_pairs and feeds it straight to
its own narrow phase; prints per-rank device times and checks shard counts sum to the full list."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, ".")
from collision_rs_b200 import abi, workloads as W
from collision_rs_b200.host import Context

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
ctx = Context(local)
w = W.cfg4_aabbs(n) if rank == 0 else None
boxes = torch.zeros(n * abi.AABB_DT.itemsize, dtype=torch.uint8, device=dev)
body_t = torch.zeros(n * abi.XFORM_DT.itemsize, dtype=torch.uint8, device=dev)
dims = torch.zeros(n * abi.SHAPE_DT.itemsize, dtype=torch.uint8, device=dev)
if rank == 0:
    boxes.copy_(torch.from_numpy(w["aabbs"].view(np.uint8).reshape(-1).copy()))
    body_t.copy_(torch.from_numpy(w["body_t"].view(np.uint8).reshape(-1).copy()))
    dims.copy_(torch.from_numpy(w["shapes"].view(np.uint8).reshape(-1).copy()))
if world > 1:  # the only collective: the tables, once
    for t in (boxes, body_t, dims):
        dist.broadcast(t, 0)
ctx.upload_shapes(dims.cpu().numpy().view(abi.SHAPE_DT))
stream = torch.cuda.current_stream()
s = stream.cuda_stream
perm = torch.zeros(n, dtype=torch.int32, device=dev)
cap = 8 * n
pairs = torch.zeros(2 * cap, dtype=torch.int32, device=dev)
cnt = 0


def sap():
    global cnt
    rc, cnt = ctx.sap3_shard_dev(boxes.data_ptr(), n, 0, rank, world, perm.data_ptr(), pairs.data_ptr(), cap, s)
    assert rc == 0


def timed(fn, reps=5, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


ms_sap = timed(sap)
p64 = perm.long()
body_shape = torch.arange(n, dtype=torch.int32, device=dev)[p64].contiguous()
body_ts = body_t.view(-1, 32)[p64].contiguous()
out = torch.zeros(max(cnt, 1) * abi.CONTACT_DT.itemsize, dtype=torch.uint8, device=dev)
st = torch.zeros(max(cnt, 1), dtype=torch.uint8, device=dev)
ms_np = timed(lambda: ctx.intersection_bodies_dev(0, body_shape.data_ptr(), body_ts.data_ptr(), pairs.data_ptr(),
                                                  cnt, out.data_ptr(), st.data_ptr(), s))
res = torch.tensor([ms_sap, ms_np, float(cnt), float((st == 1).sum())], dtype=torch.float64, device=dev)
allres = [torch.zeros_like(res) for _ in range(world)]
if world > 1:
    dist.all_gather(allres, res)
else:
    allres = [res]
if rank == 0:
    a = torch.stack(allres).cpu().numpy()
    tot = int(a[:, 2].sum())
    print(f"world={world} boxes={n} pairs={tot} per-rank pairs={a[:, 2].astype(int).tolist()}")
    print(f"  sap shard ms per rank {np.round(a[:, 0], 3).tolist()}  -> {n / a[:, 0].max() / 1e3:.1f} M AABBs/s")
    print(f"  narrow phase on own shard ms per rank {np.round(a[:, 1], 3).tolist()} -> "
          f"{tot / a[:, 1].max() / 1e3:.1f} M pairs/s, contacts {int(a[:, 3].sum())}")
# cfg 5: GJK3::intersection_time_of_impact, 8 M cuboid pairs per GPU (64 M over 8), independent shards
from collision_rs_b200.abi import CONTACT_DT
m = int(sys.argv[2]) if len(sys.argv) > 2 else 8_000_000
w5 = W.cfg5_toi_pairs(m, shard=rank)
ctx.upload_shapes(w5["shapes"])
d5 = {k: torch.from_numpy(w5[k].view(np.uint8).reshape(-1).copy()).to(dev)
      for k in ("l_shape", "r_shape", "l_t0", "l_t1", "r_t0", "r_t1")}
out5 = torch.zeros(m * CONTACT_DT.itemsize, dtype=torch.uint8, device=dev)
st5 = torch.zeros(m, dtype=torch.uint8, device=dev)
ms_toi = timed(lambda: ctx.time_of_impact_dev(d5["l_shape"].data_ptr(), d5["r_shape"].data_ptr(),
                                              d5["l_t0"].data_ptr(), d5["l_t1"].data_ptr(),
                                              d5["r_t0"].data_ptr(), d5["r_t1"].data_ptr(), m,
                                              out5.data_ptr(), st5.data_ptr(), s), reps=3, warm=2)
r5 = torch.tensor([ms_toi, float((st5 == 1).sum()), float((st5 == 5).sum())], dtype=torch.float64, device=dev)
all5 = [torch.zeros_like(r5) for _ in range(world)]
if world > 1:
    dist.all_gather(all5, r5)
else:
    all5 = [r5]
if rank == 0:
    a5 = torch.stack(all5).cpu().numpy()
    print(f"  cfg5 time of impact: {m} pairs per GPU, ms per rank {np.round(a5[:, 0], 3).tolist()} -> "
          f"{m * world / a5[:, 0].max() / 1e3:.1f} M pairs/s over {world} GPU(s), hits {int(a5[:, 1].sum())}, "
          f"nonconverged (cap 1024) {int(a5[:, 2].sum())}")
if world > 1:
    dist.destroy_process_group()
