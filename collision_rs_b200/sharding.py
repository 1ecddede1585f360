"""Multi-GPU plumbing for the narrow phase: pairs are independent, so each rank
owns a contiguous range of the pair stream; the shape / vertex tables are broadcast once; the
fixed-stride contact + status records are gathered in rank order (no counts exchange needed).
No collective sits on the data path.  Works over any torch.distributed backend: NCCL on the
GPUs (device tensors), gloo in the CPU tests (world_size 2)."""
import numpy as np
import torch
import torch.distributed as dist

from .abi import CONTACT_DT, SHAPE_DT


def shard_range(n, rank, world):
    """Contiguous [begin, end) of rank's pairs; the first n % world ranks hold one more."""
    base, rem = divmod(int(n), int(world))
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def _bytes_tensor(arr, device):
    return torch.from_numpy(np.ascontiguousarray(arr).view(np.uint8).reshape(-1).copy()).to(device)


def broadcast_shape_table(shapes, verts, src=0, device="cpu", group=None):
    """Rank `src` passes (shapes: SHAPE_DT array, verts: (V,3) f32 or None); every rank returns
    the same pair.  Other ranks may pass None."""
    rank = dist.get_rank(group)
    hdr = torch.zeros(2, dtype=torch.int64, device=device)
    if rank == src:
        hdr[0] = len(shapes)
        hdr[1] = 0 if verts is None else len(verts)
    dist.broadcast(hdr, src, group=group)
    ns, nv = int(hdr[0]), int(hdr[1])
    if rank == src:
        ts = _bytes_tensor(np.ascontiguousarray(shapes, SHAPE_DT), device)
        tv = _bytes_tensor(np.zeros((0, 3), np.float32) if verts is None
                           else np.ascontiguousarray(verts, np.float32), device)
    else:
        ts = torch.zeros(ns * SHAPE_DT.itemsize, dtype=torch.uint8, device=device)
        tv = torch.zeros(nv * 12, dtype=torch.uint8, device=device)
    if ns:
        dist.broadcast(ts, src, group=group)
    if nv:
        dist.broadcast(tv, src, group=group)
    out_s = ts.cpu().numpy().view(SHAPE_DT).copy()
    out_v = tv.cpu().numpy().view(np.float32).reshape(-1, 3).copy() if nv else None
    return out_s, out_v


def gather_contacts(contacts, status, n_total, device="cpu", group=None):
    """All-gather the per-shard (contacts: CONTACT_DT array, status: u8 array) into the records
    of the whole pair stream, in pair order, on every rank.  Shards follow shard_range()."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    b, e = shard_range(n_total, rank, world)
    assert len(contacts) == e - b and len(status) == e - b
    # equal-size slots (the largest shard) keep this a plain all_gather
    slot = shard_range(n_total, 0, world)[1]
    c = torch.zeros(slot * CONTACT_DT.itemsize, dtype=torch.uint8, device=device)
    s = torch.zeros(slot, dtype=torch.uint8, device=device)
    c[:len(contacts) * CONTACT_DT.itemsize] = _bytes_tensor(contacts, device)
    s[:len(status)] = _bytes_tensor(status, device)
    cs = [torch.zeros_like(c) for _ in range(world)]
    ss = [torch.zeros_like(s) for _ in range(world)]
    dist.all_gather(cs, c, group=group)
    dist.all_gather(ss, s, group=group)
    out_c = np.zeros(n_total, CONTACT_DT)
    out_s = np.zeros(n_total, np.uint8)
    for r in range(world):
        rb, re = shard_range(n_total, r, world)
        out_c[rb:re] = cs[r].cpu().numpy()[:(re - rb) * CONTACT_DT.itemsize].view(CONTACT_DT)
        out_s[rb:re] = ss[r].cpu().numpy()[:re - rb]
    return out_c, out_s


# ---- broad phase: shard the sweep, not the sort ------------------------------------------
def broadcast_aabbs(aabbs, src=0, device="cpu", group=None):
    """Rank `src` passes the AABB table (AABB_DT array); every rank returns it."""
    from .abi import AABB_DT
    rank = dist.get_rank(group)
    hdr = torch.zeros(1, dtype=torch.int64, device=device)
    if rank == src:
        hdr[0] = len(aabbs)
    dist.broadcast(hdr, src, group=group)
    n = int(hdr[0])
    t = (_bytes_tensor(np.ascontiguousarray(aabbs, AABB_DT), device) if rank == src
         else torch.zeros(n * AABB_DT.itemsize, dtype=torch.uint8, device=device))
    if n:
        dist.broadcast(t, src, group=group)
    return t.cpu().numpy().view(AABB_DT).copy()


def pairs_of_shard(pairs, n_boxes, rank, world):
    """The slice of a full (a, b)-pair list (reference order: b ascending, then a) that belongs to
    `rank`: pairs whose later box b sits at a sorted position in shard_range(n_boxes, rank, world).
    This is what cb2_sap3_find_collider_pairs_shard_dev computes directly on each GPU."""
    lo, hi = shard_range(n_boxes, rank, world)
    pairs = np.asarray(pairs).reshape(-1, 2)
    b = pairs[:, 1]
    return pairs[(b >= lo) & (b < hi)]


def gather_pairs(pairs, device="cpu", group=None):
    """All-gather the per-rank pair lists (ragged) and concatenate them in rank order: the
    reference's Vec.  One count exchange, then equal-size padded slots."""
    world = dist.get_world_size(group)
    pairs = np.ascontiguousarray(pairs, np.uint32).reshape(-1, 2)
    cnt = torch.tensor([len(pairs)], dtype=torch.int64, device=device)
    cnts = [torch.zeros_like(cnt) for _ in range(world)]
    dist.all_gather(cnts, cnt, group=group)
    cnts = [int(c[0]) for c in cnts]
    slot = max(max(cnts), 1)
    buf = torch.zeros(slot * 2, dtype=torch.int32, device=device)
    if len(pairs):
        buf[:2 * len(pairs)] = torch.from_numpy(pairs.reshape(-1).view(np.int32).copy()).to(device)
    bufs = [torch.zeros_like(buf) for _ in range(world)]
    dist.all_gather(bufs, buf, group=group)
    out = [bufs[r].cpu().numpy().view(np.uint32)[:2 * cnts[r]].reshape(-1, 2) for r in range(world)]
    return np.concatenate(out) if out else np.zeros((0, 2), np.uint32)
