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


def gather_contacts_dev(contacts, status, n_total, group=None):
    """Device-resident all-gather of the per-shard records: `contacts` (uint8 tensor, 36 B per pair)
    and `status` (uint8 tensor) of this rank's shard_range() slice stay on their device, travel over
    the backend's collective (NCCL over NVLink on the GPUs) and come back as the records of the
    whole pair stream, in pair order, on every rank — no host bounce.  Returns (contacts, status)
    tensors on the input device."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    rec = CONTACT_DT.itemsize
    b, e = shard_range(n_total, rank, world)
    assert contacts.numel() == (e - b) * rec and status.numel() == e - b
    slot = shard_range(n_total, 0, world)[1]  # the largest shard: equal slots keep it one all_gather
    dev = contacts.device
    if (e - b) == slot:
        c, s = contacts.view(-1), status.view(-1)
    else:
        c = torch.zeros(slot * rec, dtype=torch.uint8, device=dev)
        s = torch.zeros(slot, dtype=torch.uint8, device=dev)
        c[:contacts.numel()] = contacts.view(-1)
        s[:status.numel()] = status.view(-1)
    big_c = torch.empty(world * slot * rec, dtype=torch.uint8, device=dev)
    big_s = torch.empty(world * slot, dtype=torch.uint8, device=dev)
    if dist.get_backend(group) == "nccl":
        dist.all_gather_into_tensor(big_c, c, group=group)
        dist.all_gather_into_tensor(big_s, s, group=group)
    else:
        dist.all_gather(list(big_c.view(world, -1).unbind(0)), c, group=group)
        dist.all_gather(list(big_s.view(world, -1).unbind(0)), s, group=group)
    if n_total % world == 0:
        return big_c, big_s
    # ragged tail: drop the padding of the shorter shards
    keep_c = [big_c[r * slot * rec:(r * slot + (shard_range(n_total, r, world)[1] - shard_range(n_total, r, world)[0])) * rec]
              for r in range(world)]
    keep_s = [big_s[r * slot:r * slot + (shard_range(n_total, r, world)[1] - shard_range(n_total, r, world)[0])]
              for r in range(world)]
    return torch.cat(keep_c), torch.cat(keep_s)


def bind_to_gpu_numa(device_index):
    """Pin this process to the CPUs of the NUMA node its GPU hangs off (sysfs local_cpulist), before
    any pinned host buffer is allocated: first-touch then places the staging buffers next to the GPU
    and eight ranks stop sharing one socket's memory controller.  Returns the CPU set or None."""
    import os
    bdf = None
    try:
        pr = torch.cuda.get_device_properties(device_index)
        bdf = f"{int(pr.pci_domain_id):04x}:{int(pr.pci_bus_id):02x}:{int(pr.pci_device_id):02x}.0"
    except Exception:
        pass
    if not bdf:
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[device_index]) if vis else device_index
            bdf = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(idx)).busId
            bdf = bdf.decode() if isinstance(bdf, bytes) else bdf
        except Exception:
            return None
    bdf = bdf.lower()
    if len(bdf.split(":")[0]) == 8:  # nvml prints an 8-digit domain, sysfs a 4-digit one
        bdf = bdf[4:]
    try:
        txt = open(f"/sys/bus/pci/devices/{bdf}/local_cpulist").read().strip()
    except OSError:
        return None
    cpus = set()
    for tok in txt.split(","):
        if "-" in tok:
            a, b = tok.split("-")
            cpus.update(range(int(a), int(b) + 1))
        elif tok:
            cpus.add(int(tok))
    allowed = os.sched_getaffinity(0)
    cpus &= allowed if allowed else cpus
    if not cpus:
        return None
    os.sched_setaffinity(0, cpus)
    return sorted(cpus)


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
