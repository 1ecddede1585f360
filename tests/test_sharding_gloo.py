"""world_size-2 gloo tests (CPU) of the multi-GPU plumbing: pair sharding, shape-table
broadcast, contact gather.  The per-shard compute is the CPU oracle here (test infrastructure);
on the GPUs the same plumbing carries the CUDA path's records (bench.py --gpus N)."""
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_partitions_exactly():
    from collision_rs_b200.sharding import shard_range
    for n in (0, 1, 7, 8, 1000, 4_000_001):
        for world in (1, 2, 3, 8):
            edges = [shard_range(n, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(edges[r][1] == edges[r + 1][0] for r in range(world - 1))
            sizes = [e - b for b, e in edges]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from collision_rs_b200 import workloads as W
    from collision_rs_b200.sharding import broadcast_shape_table, gather_contacts, shard_range
    from oracle.binding import Oracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        w = W.cfg3_polyhedron_pairs(n, n_shapes=64)
        # only rank 0 owns the shape table; the others receive it
        shapes, verts = broadcast_shape_table(w["shapes"] if rank == 0 else None,
                                              w["verts"] if rank == 0 else None, src=0)
        assert np.array_equal(shapes, w["shapes"]) and np.array_equal(verts, w["verts"])
        b, e = shard_range(n, rank, world)
        o = Oracle()
        c, s = o.intersection(0, shapes, verts, w["l_shape"][b:e], w["r_shape"][b:e],
                              w["l_t"][b:e], w["r_t"][b:e], nthreads=2)
        allc, alls = gather_contacts(c, s, n)
        # the tensor-in / tensor-out gather bench.py times over NCCL (no host bounce on the GPUs)
        from collision_rs_b200.sharding import gather_contacts_dev
        import torch
        tc, ts = gather_contacts_dev(torch.from_numpy(c.view(np.uint8).reshape(-1).copy()),
                                     torch.from_numpy(s.copy()), n)
        assert tc.numpy().tobytes() == allc.tobytes() and np.array_equal(ts.numpy(), alls)
        if rank == 0:
            ec, es = o.intersection(0, shapes, verts, w["l_shape"], w["r_shape"], w["l_t"],
                                    w["r_t"], nthreads=2)
            ok = np.array_equal(alls, es) and allc.tobytes() == ec.tobytes()
            q.put(bool(ok))
    finally:
        dist.destroy_process_group()


def test_broadcast_shard_gather_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    n = 1001  # odd: ragged shards
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True


def _sap_worker(rank, world, port, n, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from collision_rs_b200 import workloads as W
    from collision_rs_b200.sharding import broadcast_aabbs, gather_pairs, pairs_of_shard
    from oracle.binding import Oracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        w = W.cfg4_aabbs(n, extent=12.0)
        boxes = broadcast_aabbs(w["aabbs"] if rank == 0 else None, src=0)
        assert boxes.tobytes() == w["aabbs"].tobytes()
        o = Oracle()
        perm, pairs, _, _ = o.sap3(boxes, 0)
        # every rank sorts redundantly (identical perm) and keeps its own slice of the pair list
        mine = pairs_of_shard(pairs, n, rank, world)
        allp = gather_pairs(mine)
        if rank == 0:
            q.put(bool(np.array_equal(allp, pairs) and 0 < len(mine) < len(pairs)))
    finally:
        dist.destroy_process_group()


def test_sap_shards_concatenate_to_the_reference_vec_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_sap_worker, args=(r, 2, port, 3001, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True


def test_pairs_of_shard_partitions_any_world():
    from collision_rs_b200.sharding import pairs_of_shard
    rng = np.random.default_rng(5)
    n = 1000
    b = np.sort(rng.integers(1, n, 5000))
    a = (rng.random(5000) * b).astype(np.uint32)
    pairs = np.stack([a, b.astype(np.uint32)], axis=1)
    for world in (1, 2, 3, 8):
        parts = [pairs_of_shard(pairs, n, r, world) for r in range(world)]
        assert np.array_equal(np.concatenate(parts), pairs)
